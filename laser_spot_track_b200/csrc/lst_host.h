// lst_host.h -- host-side logic of the tracking path (no CUDA): parameter checks, window
// placement, template constants, and the recurrence resolver that lets a whole batch of frames
// be matched in parallel while reproducing the sequential loop of LST4:793-846 exactly.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/lst_b200.h"
#include "lst_math.h"

namespace lst {

struct Rect {
    int x, y, w, h;
    bool operator==(const Rect &o) const { return x == o.x && y == o.y && w == o.w && h == o.h; }
    bool operator!=(const Rect &o) const { return !(*this == o); }
};

// matching order inside analyseSlice: mark1..mark4, then the spot (LST4:1476, 1909-2051)
static const int kMatchOrder[5] = {1, 2, 3, 4, 0};
static const int kBlurRadius = 7;   // kRadius of GaussianBlur.makeGaussianKernel(2, 0.02): 13 taps

int validate_params(const lst_params &p, std::string *err);
void fill_default_params(lst_params *p);

Rect clip_rect(int x0, int y0, int w, int h, int width, int height);
Rect template_rect(const lst_params &p, float ref_x, float ref_y);
// bounds: bit0 left, bit1 upper, bit2 right, bit3 bottom (the *Bound flags of LST4:1666-1695)
Rect window_for(const lst_params &p, const Rect &rect, int idis_x, int idis_y, int s, bool shrink, int *bounds);

// GaussianBlur.makeGaussianKernel(sigma=2, accuracy=0.02, maxRadius>=50) [ImageJ, 3P-recall]
void make_gaussian_kernel(float kern[kBlurRadius], float ksum[kBlurRadius]);

// cv::meanStdDev + the template constants of common_matchTemplate; tpl8 has row pitch tw.
TplConst template_consts(const uint8_t *tpl8, int tw, int th, int method = 5);
TplConst template_consts3(const uint8_t *planes, int tw, int th, int method = 5);   // three planes, one after the other
// The template packed for IDP.4A: word[v*kw + k] holds template bytes T[v][4k .. 4k+3] (0 beyond tw);
// kw = words per row, rounded up to a multiple of 4 so that rows load as 128-bit words.
int packed_words(int tw);
void build_packed_template(const uint8_t *tpl8, int tw, int th, std::vector<uint32_t> *out);

// The device (or, in CPU tests, a stand-in) as seen by the resolver.
struct Executor {
    virtual ~Executor() {}
    virtual int compute(const lst_task *tasks, int n, lst_task_result *results) = 0;
    // calcDisplacement for n frames: dis10[n][10] -> out5[n][5]
    virtual int solve(const double *dis10, int n, double *out5) = 0;
};

struct ResolveInput {
    lst_params p;
    Rect rect[5];
    float ref_xy[10];
    double spot0[2];
    int *spec_stride = nullptr;   // nullable: the caller's sampling stride of the speculation, adapted per batch
};

// Resolves n frames starting from displacement state dis_io (updated to the state after the
// last frame).  Returns LST_OK or an error from the executor.
int resolve_batch(const ResolveInput &in, double dis_io[10], int64_t first_frame_index, int n,
                  Executor &ex, lst_record *out, lst_stats *st);

}  // namespace lst
