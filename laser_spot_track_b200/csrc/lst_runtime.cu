// lst_runtime.cu -- device context, the CUDA executor behind the recurrence resolver, slice ingest (two-half rings on a copy
// stream: search regions as strided 3-D DMA copies from pinned memory, packed by reader threads from pageable memory and
// from slice files, whole slices on request), the asynchronous front end and the C ABI of include/lst_b200.h.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/lst_b200.h"
#include "lst_device.h"
#include "lst_host.h"
#include "lst_jpeg.h"
#include "lst_tiff.h"

using namespace lst;

// NVTX ranges (tracing): one per tracking call, batch and device pass; they cost nothing without a tool attached
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// bumped whenever pinned memory is released: the contexts drop their host-pointer -> device-alias caches
static std::atomic<uint64_t> g_pin_generation{0};

static thread_local std::string g_create_error;
static const int kRoiMarginMax = 32;   // ROI ingest: the region ring is sized for this margin
static const int kRetryWholeSlices = -100;   // internal: a window left the regions read from a slice file

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct TimedLaunch {
    cudaEvent_t e0, e1;            // around the match kernel
    cudaEvent_t ea, eb, ec;        // before lst_preprocess, before the match kernel, after lst_epilogue
    double macs;
};

struct lst_ctx : public Executor {
    lst_params p;
    int device = 0;
    std::string err;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    // reference / templates
    bool have_ref = false;
    Rect rect[5];
    float ref_xy[10];
    double spot0[2] = {0, 0};
    double dis[10] = {0};
    DevTemplates h_tpls;
    DevTemplates *d_tpls = nullptr;
    uint32_t *d_packed[kMaxTpl] = {nullptr};
    uint8_t *d_raw[kMaxTpl] = {nullptr};
    uint8_t *d_toep[kMaxTpl] = {nullptr};   // tensor-core engine: Toeplitz tables
    size_t toep_cap[kMaxTpl] = {0};
    int ncc_engine = 0;                     // 0 = IDP.4A (lst_ncc_match), 1 = tensor cores (lst_ncc_tc)
    size_t packed_cap[kMaxTpl] = {0}, raw_cap[kMaxTpl] = {0};
    // per-pass scratch
    int task_cap = 0;
    DevTask *d_tasks = nullptr, *h_tasks = nullptr;
    uint32_t *d_mm = nullptr;
    unsigned long long *d_best = nullptr;
    float *d_nbhd = nullptr;      // [task][12]: 3x3 scores around the peak + flag, left by lst_ncc_tc
    DevResult *d_results = nullptr, *h_results = nullptr;
    uint8_t *d_win8 = nullptr;
    size_t win8_cap = 0;
    float *d_map = nullptr;
    size_t map_cap = 0;
    // surface table (where each task's pixels live) + slice ring (whole-slice ingest)
    int frame_cap = 0;
    Surface *d_surf = nullptr, *h_surf = nullptr;
    uint8_t *d_ring[2] = {nullptr, nullptr};
    size_t ring_cap = 0;   // bytes per half
    // ROI ingest: pinned host slices stay on the host, only the search regions are gathered
    uint8_t *d_roi[2] = {nullptr, nullptr};
    size_t roi_cap = 0;    // bytes per half
    const void **d_gptr[2] = {nullptr, nullptr}, **h_gptr[2] = {nullptr, nullptr};
    int gptr_cap = 0;
    RoiDesc roi[2][5];     // per half: the rectangles gathered for that batch
    bool roi_active = false;   // the batch being resolved uses ROI surfaces
    int roi_half = 0, roi_nb = 0;
    int roi_margin[5] = {8, 8, 8, 8, 8};   // current margin per region; adapts to the excursion of that template's windows in the last batch
    int roi_margin_fixed = -1;   // LST_ROI_MARGIN pins it
    int roi_align_bytes = 16;
    std::unordered_map<const void *, const void *> mapped_cache;   // pinned host ptr -> device-visible ptr
    uint64_t mapped_generation = 0;    // g_pin_generation the cache was filled under
    // slices read from files (lst_track_files): pinned staging with the layout of the ROI ring
    uint8_t *h_fstage[2] = {nullptr, nullptr};
    size_t fstage_cap = 0;
    // JPEG slice files: quantised coefficients of the MCUs under the search regions (+ the slices' quantisation tables), staged by
    // the reader threads and reconstructed on the device (lst_jpeg_reconstruct); LST_JPEG_DEVICE=0 keeps the host reconstruction
    int16_t *h_jcoef[2] = {nullptr, nullptr}, *d_jcoef[2] = {nullptr, nullptr};
    uint16_t *h_jqt[2] = {nullptr, nullptr}, *d_jqt[2] = {nullptr, nullptr};
    size_t jcoef_cap = 0;          // int16 elements per half
    int jqt_cap = 0;               // slices per half
    bool jpeg_device = true;
    uint8_t *h_fwhole = nullptr;   // whole slices of a sub-batch (fallback / whole-slice search)
    size_t fwhole_cap = 0;
    bool file_mode = false;        // the batch being resolved has no whole slice in memory to fall back to
    bool swap_bytes = false;       // the slices being tracked hold big-endian samples
    int spec_stride = 8;           // sampling stride of the resolver's speculation (adapts per batch)
    int file_threads = 8;
    // solve buffers
    int solve_cap = 0;
    double *d_dis10 = nullptr, *d_out5 = nullptr, *h_dis10 = nullptr, *h_out5 = nullptr;
    // stats
    lst_stats stats;
    bool timing = false;
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> event_pool;
    std::vector<int> order;   // device slot -> caller's task index (tasks are grouped by template)
    std::vector<int> first_task_of_frame;   // whole-slice search: the task whose 8-bit window the later tasks of a surface share
    size_t scratch_limit = (size_t)6 << 30;

    size_t bpp() const { return p.pixel_type == LST_PIX_U8 ? 1 : (p.pixel_type == LST_PIX_U16 ? 2 : 4); }
    size_t frame_bytes() const { return (size_t)p.width * p.height * bpp(); }

    int fail(int code, const char *fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    int cuda_fail(cudaError_t e, const char *what) { return fail(LST_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e)); }

    int compute(const lst_task *tasks, int n, lst_task_result *results) override;
    int solve(const double *dis10, int n, double *out5) override;
    int run_group(const lst_task *tasks, int n, lst_task_result *results, bool preprocess_only, bool have_win8,
                  int tpl_override, float *map_out_host);
    // LST_RANGE_FRAME: per-slice {min, max} keys of the batch being resolved, and what the caller announced for the next call
    uint32_t *d_frame_mm = nullptr, *h_frame_mm = nullptr;
    DevTask *d_ftasks = nullptr, *h_ftasks = nullptr;
    int frame_mm_cap = 0;
    std::vector<double> user_ranges;     // announced with lst_set_frame_ranges for the next call
    std::vector<double> cur_ranges;      // the ranges of the tracking call in progress, slice `range_first` first
    int64_t range_first = 0;
    int track_depth = 0;
    // RGB stack with matchIntensity = false: the channels are blurred separately (ColorProcessor) and OpenCV gets a CV_8UC3 image
    bool rgb3() const { return p.pixel_type == LST_PIX_RGB && !p.match_intensity; }
    bool frame_range_mode() const
    {
        return p.range_mode == LST_RANGE_FRAME && !((p.pixel_type == LST_PIX_U8 || p.pixel_type == LST_PIX_RGB) && !p.match_intensity);
    }
    int upload_template3(int slot, const uint8_t *planes, int tw, int th);
    size_t last_plane_bytes = 0;   // distance between the channel planes in d_win8 after the last pass (rgb3)
    int prepare_frame_ranges(int nb, const double *lo_hi);
    // asynchronous front end (lst_submit / lst_fetch / lst_override): batches tracked in order by a worker thread
    struct AsyncJob {
        std::vector<const void *> ptrs;
        int64_t first = 0;
        std::vector<lst_record> out;
        size_t fetched = 0;
        int rc = LST_OK;
        std::string err;
        int state = 0;       // 0 queued, 1 running, 2 done
    };
    std::deque<std::unique_ptr<AsyncJob>> aq;
    std::mutex amx;
    std::condition_variable acv_work, acv_done;
    std::thread aworker;
    bool astop = false;
    int queue_depth = 2;
    bool async_busy()
    {
        std::lock_guard<std::mutex> lk(amx);
        for (auto &j : aq)
            if (j->state != 2) return true;
        return false;
    }
    void async_loop();
    int preproc_engine = 0;      // 0 = auto (lst_preprocess_win when the window fits), 1 = always lst_preprocess (LST_PREPROC=old)
    int ensure_tasks(int n);
    int ensure_frames(int n);
    int ensure_scratch(size_t win_bytes, size_t map_elems);
    int upload_template(int slot, const uint8_t *tpl8, int tw, int th, int method_override = -1);
    void preproc_params(PreprocParams *pp, int max_w, int max_h) const;
    void drain_timers();
};

#define CU(call)                                             \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
    } while (0)
#define CUC(ctx, call)                                              \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return (ctx)->cuda_fail(e_, #call);  \
    } while (0)

template <typename T>
static cudaError_t regrow(T **p, size_t count)
{
    if (*p) cudaFree(*p);
    *p = nullptr;
    return cudaMalloc((void **)p, count * sizeof(T));
}
template <typename T>
static cudaError_t regrow_host(T **p, size_t count)
{
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    return cudaMallocHost((void **)p, count * sizeof(T));
}

int lst_ctx::ensure_tasks(int n)
{
    if (n <= task_cap) return LST_OK;
    CU(cudaStreamSynchronize(s_compute));
    int cap = std::max(n, task_cap * 2);
    CU(regrow(&d_tasks, cap));
    CU(regrow(&d_mm, (size_t)cap * 2));
    CU(regrow(&d_best, cap));
    CU(regrow(&d_nbhd, (size_t)cap * 12));
    CU(regrow(&d_results, cap));
    CU(regrow_host(&h_tasks, cap));
    CU(regrow_host(&h_results, cap));
    task_cap = cap;
    return LST_OK;
}

int lst_ctx::ensure_frames(int n)
{
    if (n <= frame_cap) return LST_OK;
    CU(cudaStreamSynchronize(s_compute));
    int cap = std::max(n, frame_cap * 2);
    CU(regrow(&d_surf, cap));
    CU(regrow_host(&h_surf, cap));
    frame_cap = cap;
    return LST_OK;
}

static Surface whole_slice(const void *ptr, int width, int height)
{
    Surface s;
    s.ptr = ptr;
    s.pitch_px = width;
    s.org_x = 0;
    s.org_y = 0;
    s.rows = height;
    return s;
}

int lst_ctx::ensure_scratch(size_t win_bytes, size_t map_elems)
{
    if (win_bytes > win8_cap) {
        CU(cudaStreamSynchronize(s_compute));
        size_t cap = std::max(win_bytes, win8_cap + win8_cap / 2);
        CU(regrow(&d_win8, cap));
        win8_cap = cap;
    }
    if (map_elems > map_cap) {
        CU(cudaStreamSynchronize(s_compute));
        CU(regrow(&d_map, map_elems));
        map_cap = map_elems;
    }
    return LST_OK;
}

void lst_ctx::preproc_params(PreprocParams *pp, int max_w, int max_h) const
{
    memset(pp, 0, sizeof(*pp));
    pp->pixel_type = p.pixel_type;
    pp->frame_w = p.width;
    pp->frame_h = p.height;
    pp->quant_mode = p.quant_mode;
    pp->rgb_shift = -1;
    if ((p.pixel_type == LST_PIX_U8 || p.pixel_type == LST_PIX_RGB) && !p.match_intensity) {
        // ByteProcessor blur (and ColorProcessor blur, channel by channel) rounds back with (int)(v+0.5f), clamped: == ROUND255 on [0,255]
        pp->range_per_task = 0;
        pp->lo = 0.0f;
        pp->hi = 255.0f;
        pp->quant_mode = LST_QUANT_ROUND255;
    } else if (p.range_mode == LST_RANGE_FIXED) {
        pp->range_per_task = 0;
        pp->lo = (float)p.range_lo;
        pp->hi = (float)p.range_hi;
    } else if (p.range_mode == LST_RANGE_FRAME) {
        pp->range_per_task = 2;
        pp->frame_mm = d_frame_mm;
    } else if (p.range_mode == LST_RANGE_DISPLAY && (p.pixel_type == LST_PIX_U8 || p.pixel_type == LST_PIX_RGB)) {
        // 8-bit and RGB images keep their 0..255 display range through convertToGray32
        pp->range_per_task = 0;
        pp->lo = 0.0f;
        pp->hi = 255.0f;
    } else {
        pp->range_per_task = 1;
    }
    int tw = (max_w + 7) & ~7;
    pp->tile_w = tw < 192 ? ((tw + 15) & ~15) : 192;   // whole window rows for the named configurations (S <= 192); multiple of 16
    pp->minmax_in_kernel = (pp->range_per_task == 1 && max_w <= pp->tile_w && (long)max_w * max_h <= 256L * 256) ? 1 : 0;
    pp->swap_bytes = swap_bytes ? 1 : 0;
}

int lst_ctx::upload_template(int slot, const uint8_t *tpl8, int tw, int th, int method_override)
{
    TplConst c = template_consts(tpl8, tw, th, method_override >= 0 ? method_override : p.method);
    std::vector<uint32_t> sh;
    build_packed_template(tpl8, tw, th, &sh);
    if (sh.size() > packed_cap[slot]) {
        CU(regrow(&d_packed[slot], sh.size()));
        packed_cap[slot] = sh.size();
    }
    if ((size_t)tw * th > raw_cap[slot]) {
        CU(regrow(&d_raw[slot], (size_t)tw * th));
        raw_cap[slot] = (size_t)tw * th;
    }
    CU(cudaMemcpyAsync(d_packed[slot], sh.data(), sh.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s_compute));
    CU(cudaMemcpyAsync(d_raw[slot], tpl8, (size_t)tw * th, cudaMemcpyHostToDevice, s_compute));
    h_tpls.c[slot] = c;
    h_tpls.packed[slot] = d_packed[slot];
    h_tpls.raw[slot] = d_raw[slot];
    std::vector<uint8_t> toep;
    if (ncc_engine == 1 && tc_supported(tw, th)) {
        toep.resize(tc_table_bytes(tw, th));
        tc_build_table(tpl8, tw, th, toep.data());
        if (toep.size() > toep_cap[slot]) {
            CU(regrow(&d_toep[slot], toep.size()));
            toep_cap[slot] = toep.size();
        }
        CU(cudaMemcpyAsync(d_toep[slot], toep.data(), toep.size(), cudaMemcpyHostToDevice, s_compute));
        h_tpls.toeplitz[slot] = d_toep[slot];
        stats.h2d_bytes += (int64_t)toep.size();
    }
    CU(cudaMemcpyAsync(d_tpls, &h_tpls, sizeof(DevTemplates), cudaMemcpyHostToDevice, s_compute));
    CU(cudaStreamSynchronize(s_compute));   // sh / h_tpls are pageable: finish before they go away
    stats.h2d_bytes += (int64_t)(sh.size() * 4 + (size_t)tw * th + sizeof(DevTemplates));
    return LST_OK;
}

// 3-channel template (rgb3): the three blurred channel planes one after the other; matched by lst_ncc_rgb only
int lst_ctx::upload_template3(int slot, const uint8_t *planes, int tw, int th)
{
    const size_t bytes = 3 * (size_t)tw * th;
    if (bytes > raw_cap[slot]) {
        CU(regrow(&d_raw[slot], bytes));
        raw_cap[slot] = bytes;
    }
    CU(cudaMemcpyAsync(d_raw[slot], planes, bytes, cudaMemcpyHostToDevice, s_compute));
    h_tpls.c[slot] = template_consts3(planes, tw, th, p.method);
    h_tpls.raw[slot] = d_raw[slot];
    h_tpls.packed[slot] = nullptr;
    h_tpls.toeplitz[slot] = nullptr;
    CU(cudaMemcpyAsync(d_tpls, &h_tpls, sizeof(DevTemplates), cudaMemcpyHostToDevice, s_compute));
    CU(cudaStreamSynchronize(s_compute));
    stats.h2d_bytes += (int64_t)(bytes + sizeof(DevTemplates));
    return LST_OK;
}

void lst_ctx::drain_timers()
{
    for (TimedLaunch &t : timed) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
            stats.ncc_kernel_ms += ms;
            stats.ncc_kernel_launches += 1;
            stats.ncc_algorithmic_macs += t.macs;
        }
        if (cudaEventElapsedTime(&ms, t.ea, t.eb) == cudaSuccess) stats.preprocess_ms += ms;
        if (cudaEventElapsedTime(&ms, t.eb, t.e0) == cudaSuccess) stats.box_sums_ms += ms;
        if (cudaEventElapsedTime(&ms, t.e1, t.ec) == cudaSuccess) stats.epilogue_ms += ms;
        for (cudaEvent_t e : {t.e0, t.e1, t.ea, t.eb, t.ec}) event_pool.push_back(e);
    }
    timed.clear();
}

// order-preserving key of a slice's min or max as the blur kernels read it back (px_key / key_to_float of lst_kernels.cu)
static uint32_t host_range_key(int pixel_type, double v)
{
    if (pixel_type == LST_PIX_F32) {
        const float f = (float)v;
        uint32_t u;
        memcpy(&u, &f, 4);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
    const double hi = pixel_type == LST_PIX_U16 ? 65535.0 : 255.0;
    return (uint32_t)std::min(std::max(v, 0.0), hi);
}

// LST_RANGE_FRAME: fills d_frame_mm for the nb slices of the batch whose whole-slice surfaces are d_surf[0 .. nb) -- from
// the caller's ranges, or by reducing every slice on the device (lst_minmax over the whole slice).
int lst_ctx::prepare_frame_ranges(int nb, const double *lo_hi)
{
    if (nb > frame_mm_cap) {
        CU(cudaStreamSynchronize(s_compute));
        const int cap = std::max(nb, frame_mm_cap * 2);
        CU(regrow(&d_frame_mm, (size_t)cap * 2));
        CU(regrow_host(&h_frame_mm, (size_t)cap * 2));
        CU(regrow(&d_ftasks, cap));
        CU(regrow_host(&h_ftasks, cap));
        frame_mm_cap = cap;
    }
    for (int i = 0; i < nb; i++) {
        h_frame_mm[2 * i] = lo_hi ? host_range_key(p.pixel_type, lo_hi[2 * i]) : 0xffffffffu;
        h_frame_mm[2 * i + 1] = lo_hi ? host_range_key(p.pixel_type, lo_hi[2 * i + 1]) : 0u;
    }
    if (launch_upload(d_frame_mm, h_frame_mm, sizeof(uint32_t) * 2 * nb, s_compute) < 0) return fail(LST_ERR_CUDA, "range upload failed");
    stats.kernel_launches += 1;
    if (!lo_hi) {
        for (int i = 0; i < nb; i++) {
            DevTask &d = h_ftasks[i];
            memset(&d, 0, sizeof(d));
            d.frame = i;
            d.w = p.width;
            d.h = p.height;
        }
        if (launch_upload(d_ftasks, h_ftasks, sizeof(DevTask) * nb, s_compute) < 0) return fail(LST_ERR_CUDA, "task upload failed");
        PreprocParams pp;
        preproc_params(&pp, p.width, p.height);
        stats.kernel_launches += 1 + launch_minmax(d_ftasks, nb, p.height, d_surf, pp, d_frame_mm, s_compute);
    }
    return LST_OK;
}

// One device pass over a group of windows that fits the scratch.
//   preprocess_only: stop after the 8-bit windows (template extraction, lst_preprocess)
//   have_win8:       the 8-bit windows are already in scratch (lst_match_single)
int lst_ctx::run_group(const lst_task *tasks, int n, lst_task_result *results, bool preprocess_only, bool have_win8,
                       int tpl_override, float *map_out_host)
{
    NvtxRange nvtx_pass(preprocess_only ? "lst pass (blur only)" : "lst pass");
    int rc = ensure_tasks(n);
    if (rc != LST_OK) return rc;
    size_t win_off = 0, sum_off = 0;
    int max_w = 0, max_h = 0, max_rw = 0, max_rh = 0, max_tw = 0, max_th = 0;
    double macs = 0;
    // device order: grouped by template, so that CTAs resident on one SM read the same packed
    // template through L1 (order[i] = index of the caller's task in device slot i)
    order.resize(n);
    {
        int cnt[kMaxTpl + 1] = {0};
        for (int i = 0; i < n; i++) cnt[std::min(std::max(tasks[i].tpl, 0), kMaxTpl - 1) + 1]++;
        for (int k = 0; k < kMaxTpl; k++) cnt[k + 1] += cnt[k];
        for (int i = 0; i < n; i++) order[cnt[std::min(std::max(tasks[i].tpl, 0), kMaxTpl - 1)]++] = i;
    }
    // The tensor-core engine reads the 8-bit windows by TMA as one 2-D byte image: every window of the pass gets
    // the same row pitch (the widest window's) and starts at a multiple of it.
    const bool rgb = rgb3() && !have_win8;   // three channel planes per window, matched by lst_ncc_rgb
    if (rgb && map_out_host) return fail(LST_ERR_UNSUPPORTED, "no score map output for the 3-channel match");
    const bool tc_layout = ncc_engine == 1 && !preprocess_only && !map_out_host && !rgb;
    int pass_pitch = 0;
    if (tc_layout)
        for (int i = 0; i < n; i++) {
            const lst_task &t = tasks[i];
            Rect c = have_win8 ? Rect{0, 0, t.ww, t.wh} : clip_rect(t.wx, t.wy, t.ww, t.wh, p.width, p.height);
            pass_pitch = std::max(pass_pitch, (c.w + 15) & ~15);
        }
    first_task_of_frame.assign(first_task_of_frame.size(), -1);
    for (int i = 0; i < n; i++) {
        const lst_task &t = tasks[order[i]];
        DevTask &d = h_tasks[i];
        memset(&d, 0, sizeof(d));
        Rect c = have_win8 ? Rect{0, 0, t.ww, t.wh} : clip_rect(t.wx, t.wy, t.ww, t.wh, p.width, p.height);
        d.frame = t.frame;
        if (roi_active && !have_win8) {
            // pixels come from the gathered ROI of (slice, template) when the window lies inside it,
            // else straight from the pinned host slice (zero-copy; rare: speculation margin exceeded)
            const RoiDesc &rr = roi[roi_half][t.tpl];
            if (c.x >= rr.x && c.y >= rr.y && c.x + c.w <= rr.x + rr.w && c.y + c.h <= rr.y + rr.h) {
                d.frame = t.frame * 5 + t.tpl;
            } else {
                if (file_mode) return kRetryWholeSlices;   // no slice in memory: the caller re-reads this batch whole
                d.frame = 5 * roi_nb + t.frame;
                stats.roi_fallback_tasks++;
                stats.h2d_bytes += (int64_t)c.w * c.h * (int64_t)bpp() * 2;
            }
        }
        d.tpl = tpl_override >= 0 ? tpl_override : t.tpl;
        d.pad1 = t.frame;
        d.x0 = c.x;
        d.y0 = c.y;
        d.w = c.w;
        d.h = c.h;
        d.pitch = tc_layout ? pass_pitch : (c.w + 15) & ~15;
        d.s_used = t.s_used;
        if (c.w < 2 * kBlurRadius || c.h < 2 * kBlurRadius)
            if (!have_win8) return fail(LST_ERR_UNSUPPORTED, "window %dx%d smaller than the blur support", c.w, c.h);
        if (!preprocess_only) {
            const TplConst &tc = h_tpls.c[d.tpl];
            d.rw = c.w - tc.tw + 1;
            d.rh = c.h - tc.th + 1;
            if (d.rw < 1 || d.rh < 1) return fail(LST_ERR_INVALID, "window %dx%d smaller than the template", c.w, c.h);
            max_rw = std::max(max_rw, d.rw);
            max_rh = std::max(max_rh, d.rh);
            max_tw = std::max(max_tw, tc.tw);
            max_th = std::max(max_th, tc.th);
            macs += (double)d.rw * d.rh * tc.tw * tc.th;
        }
        d.sum_off = sum_off;
        // Whole-slice search (sArea = 0, LST4:1316-1320): the five templates of a slice are matched against the same image --
        // it is blurred once, the other four tasks point at that 8-bit window (pad0 = 1: the blur kernels skip the task).
        int twin = -1;
        if (p.s_area == 0 && !have_win8 && !preprocess_only && d.frame >= 0) {
            if ((size_t)d.frame >= first_task_of_frame.size()) first_task_of_frame.resize((size_t)d.frame + 1, -1);
            const int j = first_task_of_frame[d.frame];
            if (j < 0) {
                first_task_of_frame[d.frame] = i;
            } else {
                const DevTask &e = h_tasks[j];
                if (e.x0 == d.x0 && e.y0 == d.y0 && e.w == d.w && e.h == d.h && e.pitch == d.pitch) twin = j;
            }
        }
        if (twin >= 0) {
            d.win_off = h_tasks[twin].win_off;
            d.pad0 = 1;
        } else {
            d.win_off = win_off;
            win_off += tc_layout ? (size_t)d.pitch * d.h : (((size_t)d.pitch * d.h + 255) & ~(size_t)255);
        }
        sum_off += (size_t)d.rw * d.rh;
        max_w = std::max(max_w, d.w);
        max_h = std::max(max_h, d.h);
    }
    const size_t plane_bytes = rgb ? ((win_off + 255) & ~(size_t)255) : 0;
    last_plane_bytes = plane_bytes;
    rc = ensure_scratch(rgb ? 3 * plane_bytes + 64 : win_off + 64, map_out_host ? sum_off + 32 : 0);
    if (rc != LST_OK) return rc;
    int launches = 0;
    {
        int up = launch_upload(d_tasks, h_tasks, sizeof(DevTask) * n, s_compute);
        if (up < 0) return fail(LST_ERR_CUDA, "task upload failed");
        launches += up;
    }
    stats.h2d_bytes += (int64_t)sizeof(DevTask) * n;
    launches += launch_init_pass(n, d_mm, d_best, d_nbhd, s_compute);
    TimedLaunch tl;
    const bool timed_pass = timing && !preprocess_only && !have_win8;
    if (timed_pass) {
        for (cudaEvent_t *e : {&tl.e0, &tl.e1, &tl.ea, &tl.eb, &tl.ec}) {
            if (!event_pool.empty()) {
                *e = event_pool.back();
                event_pool.pop_back();
            } else {
                CU(cudaEventCreate(e));
            }
        }
        CU(cudaEventRecord(tl.ea, s_compute));
    }
    if (!have_win8) {
        NvtxRange nvtx_blur("lst_preprocess");
        PreprocParams pp;
        preproc_params(&pp, max_w, max_h);
        if (rgb) {
            for (int k = 0; k < 3; k++) {   // planes in the order R, G, B (matchTemplate sums over the channels: the order is immaterial)
                pp.rgb_shift = 16 - 8 * k;
                const int l3 = launch_preprocess(d_tasks, n, max_w, max_h, d_surf, pp, d_mm, d_win8 + k * plane_bytes, s_compute);
                if (l3 < 0) return fail(LST_ERR_CUDA, "lst_preprocess launch failed");
                launches += l3;
            }
        } else if (preproc_engine == 0 && preprocess_win_eligible(p.pixel_type, max_w, max_h)) {
            int l2 = launch_preprocess_win(d_tasks, n, max_w, max_h, d_surf, pp, d_win8, s_compute);
            if (l2 < 0) return fail(LST_ERR_CUDA, "lst_preprocess_win launch failed");
            launches += l2;
        } else {
            if (pp.range_per_task == 1 && !pp.minmax_in_kernel) launches += launch_minmax(d_tasks, n, max_h, d_surf, pp, d_mm, s_compute);
            launches += launch_preprocess(d_tasks, n, max_w, max_h, d_surf, pp, d_mm, d_win8, s_compute);
        }
    }
    if (!preprocess_only) {
        MatchParams mp;
        size_t smem = 0;
        plan_match(max_rw, max_rh, max_tw, max_th, n, &mp, &smem);
        mp.sub_pixel = p.sub_pixel;
        mp.templ_size = p.templ_size;
        mp.thresh = p.match_threshold;
        mp.plane_bytes = plane_bytes;
        if (smem > 200 * 1024 && !rgb) return fail(LST_ERR_UNSUPPORTED, "match tile needs %zu bytes of shared memory", smem);
        if (timed_pass) {
            tl.macs = macs;
            CU(cudaEventRecord(tl.eb, s_compute));
            CU(cudaEventRecord(tl.e0, s_compute));
        }
        if (rgb) {
            for (int i = 0; i < n; i++)
                if (h_tpls.c[h_tasks[i].tpl].channels != 3) return fail(LST_ERR_STATE, "3-channel match without a 3-channel template");
            const int l3 = launch_ncc_rgb(d_tasks, n, max_rh, max_w, max_tw, max_th, d_tpls, d_win8, plane_bytes, d_best, s_compute);
            if (l3 < 0) return fail(LST_ERR_UNSUPPORTED, "3-channel match: window %dx%d / template %dx%d too large", max_w, max_h, max_tw, max_th);
            launches += l3;
        }
        bool use_tc = tc_layout && tc_supported(max_tw, max_th);
        if (use_tc)
            for (int i = 0; i < n && use_tc; i++) use_tc = h_tpls.toeplitz[h_tasks[i].tpl] != nullptr;
        int tcl = -1;
        if (use_tc)
            tcl = launch_ncc_tc(d_tasks, n, max_rw, max_rh, max_tw, max_th, pass_pitch, (long long)(win_off / (size_t)pass_pitch), d_tpls,
                                d_win8, d_best, d_nbhd, s_compute);
        if (use_tc && tcl < 0) {
            static bool warned = false;
            if (!warned) fprintf(stderr, "lst_b200: tensor-core match engine not applicable to this window/template size; using IDP.4A\n");
            warned = true;
        }
        if (rgb)
            ;
        else if (tcl >= 0)
            launches += tcl;
        else
            launches += launch_ncc_match(d_tasks, n, max_rw, max_rh, max_tw, max_th, d_tpls, d_win8, d_best,
                                         map_out_host ? d_map : nullptr, mp, s_compute);
        if (timed_pass) CU(cudaEventRecord(tl.e1, s_compute));
        launches += launch_epilogue(d_tasks, n, d_tpls, d_win8, d_best, d_nbhd, d_results, mp, s_compute);
        if (timed_pass) {
            CU(cudaEventRecord(tl.ec, s_compute));
            timed.push_back(tl);
            for (int i = 0; i < n; i++)
                if (!h_tasks[i].pad0) stats.preprocess_px += (double)h_tasks[i].w * h_tasks[i].h;
            stats.preprocess_launches += 1;
        }
        CU(cudaMemcpyAsync(h_results, d_results, sizeof(DevResult) * n, cudaMemcpyDeviceToHost, s_compute));
        stats.d2h_bytes += (int64_t)sizeof(DevResult) * n;
    }
    CU(cudaGetLastError());
    {
        double t0 = now_ms();
        CU(cudaStreamSynchronize(s_compute));
        stats.device_wait_ms += now_ms() - t0;
    }
    stats.kernel_launches += launches;
    if (timing) drain_timers();
    if (!preprocess_only && results) {
        for (int i = 0; i < n; i++) {
            const DevResult &r = h_results[i];
            if (!r.peak_check)
                return fail(LST_ERR_CUDA, "internal: epilogue peak score differs from the argmax score (task %d)", i);
            lst_task_result &o = results[order[i]];
            o.x = r.x;
            o.y = r.y;
            o.score = r.score;
            o.ix = r.ix;
            o.iy = r.iy;
            o.accepted = r.accepted;
            o.reserved = 0;
        }
    }
    if (map_out_host && !preprocess_only) {
        // single task: the map starts at sum_off 0
        CU(cudaMemcpy(map_out_host, d_map, sizeof(float) * (size_t)h_tasks[0].rw * h_tasks[0].rh, cudaMemcpyDeviceToHost));
    }
    return LST_OK;
}

int lst_ctx::compute(const lst_task *tasks, int n, lst_task_result *results)
{
    // split into groups whose scratch fits the limit
    int i = 0;
    while (i < n) {
        size_t bytes = 0;
        int j = i;
        while (j < n) {
            const lst_task &t = tasks[j];
            size_t pitch = (t.ww + 15) & ~15;
            size_t need = (pitch * t.wh + 256) * (rgb3() ? 3 : 1);
            if (j > i && bytes + need > scratch_limit) break;
            bytes += need;
            j++;
        }
        int rc = run_group(tasks + i, j - i, results + i, false, false, -1, nullptr);
        if (rc != LST_OK) return rc;
        i = j;
    }
    return LST_OK;
}

int lst_ctx::solve(const double *dis10, int n, double *out5)
{
    if (n > solve_cap) {
        CU(cudaStreamSynchronize(s_compute));
        int cap = std::max(n, solve_cap * 2);
        CU(regrow(&d_dis10, (size_t)cap * 10));
        CU(regrow(&d_out5, (size_t)cap * 5));
        CU(regrow_host(&h_dis10, (size_t)cap * 10));
        CU(regrow_host(&h_out5, (size_t)cap * 5));
        solve_cap = cap;
    }
    memcpy(h_dis10, dis10, sizeof(double) * 10 * n);
    if (launch_upload(d_dis10, h_dis10, sizeof(double) * 10 * n, s_compute) < 0) return fail(LST_ERR_CUDA, "upload failed");
    stats.kernel_launches += 1;
    stats.kernel_launches += launch_solve(d_dis10, n, ref_xy, p.mark_dist, spot0, d_out5, s_compute);
    CU(cudaMemcpyAsync(h_out5, d_out5, sizeof(double) * 5 * n, cudaMemcpyDeviceToHost, s_compute));
    CU(cudaGetLastError());
    {
        double t0 = now_ms();
        CU(cudaStreamSynchronize(s_compute));
        stats.device_wait_ms += now_ms() - t0;
    }
    memcpy(out5, h_out5, sizeof(double) * 5 * n);
    stats.h2d_bytes += (int64_t)sizeof(double) * 10 * n;
    stats.d2h_bytes += (int64_t)sizeof(double) * 5 * n;
    return LST_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int lst_abi_version(void) { return LST_ABI_VERSION; }

int lst_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *lst_last_error(const lst_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int lst_create(const lst_params *params, int device, lst_ctx **out)
{
    if (!params || !out) {
        g_create_error = "null argument";
        return LST_ERR_INVALID;
    }
    *out = nullptr;
    std::string e;
    int rc = validate_params(*params, &e);
    if (rc != LST_OK) {
        g_create_error = e;
        return rc;
    }
    int n = lst_device_count();
    if (n <= 0) {
        g_create_error = "no CUDA device: this library has no CPU fallback";
        return LST_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        g_create_error = "bad device index";
        return LST_ERR_INVALID;
    }
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) {
        g_create_error = cudaGetErrorString(ce);
        return LST_ERR_CUDA;
    }
    lst_ctx *c = new lst_ctx();
    c->p = *params;
    if (c->p.max_batch == 0) c->p.max_batch = 512;
    if (c->p.match_threshold == 0.0) {   // matchThreshold[] of LST4:147
        static const double kThr[6] = {0.1, 0.1, 0.05, 0.05, 0.2, 0.2};
        c->p.match_threshold = kThr[c->p.method];
    }
    c->device = device;
    memset(&c->stats, 0, sizeof(c->stats));
    memset(&c->h_tpls, 0, sizeof(c->h_tpls));
    auto bail = [&](cudaError_t err2, const char *what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err2);
        lst_destroy(c);
        return LST_ERR_CUDA;
    };
    if ((ce = cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking)) != cudaSuccess) return bail(ce, "stream");
    {   // the ingest stream outranks the compute stream: its few CTAs must not queue behind a full grid
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if ((ce = cudaStreamCreateWithPriority(&c->s_copy, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail(ce, "stream");
    }
    for (int i = 0; i < 2; i++)
        if ((ce = cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming)) != cudaSuccess) return bail(ce, "event");
    if ((ce = cudaMalloc((void **)&c->d_tpls, sizeof(DevTemplates))) != cudaSuccess) return bail(ce, "cudaMalloc");
    {
        // engine of the match kernel: params.ncc_engine, overridable by LST_NCC_ENGINE=dp4a|tc for A/B runs
        int eng = params->ncc_engine;
        const char *e = getenv("LST_NCC_ENGINE");
        if (e && !strcmp(e, "dp4a")) eng = 1;
        if (e && !strcmp(e, "tc")) eng = 2;
        c->ncc_engine = eng == 1 ? 0 : 1;      // internal: 1 = tensor cores wherever lst_ncc_tc applies
        tc_init_device();
        if (const char *m = getenv("LST_ROI_MARGIN")) {
            c->roi_margin_fixed = std::max(0, atoi(m));
            for (int k = 0; k < 5; k++) c->roi_margin[k] = c->roi_margin_fixed;
        }
        if (const char *a = getenv("LST_ROI_ALIGN")) c->roi_align_bytes = std::max(4, atoi(a));
        if (const char *b = getenv("LST_PREPROC")) c->preproc_engine = !strcmp(b, "old") ? 1 : 0;   // A/B runs of the blur kernels
    }
    float kern[7], ksum[7];
    make_gaussian_kernel(kern, ksum);
    if (set_blur_kernel(kern, ksum) != 0) return bail(cudaGetLastError(), "blur kernel upload");
    if (c->p.pixel_type == LST_PIX_RGB) {
        double w[3] = {c->p.rgb_weights[0], c->p.rgb_weights[1], c->p.rgb_weights[2]};
        if (w[0] == 0.0 && w[1] == 0.0 && w[2] == 0.0) w[0] = w[1] = w[2] = 1.0 / 3.0;   // ColorProcessor default
        if (set_rgb_weights(w) != 0) return bail(cudaGetLastError(), "RGB weights upload");
    }
    *out = c;
    return LST_OK;
}

void lst_destroy(lst_ctx *c)
{
    if (!c) return;
    if (c->aworker.joinable()) {
        {
            std::lock_guard<std::mutex> lk(c->amx);
            c->astop = true;
        }
        c->acv_work.notify_all();
        c->aworker.join();
    }
    cudaSetDevice(c->device);
    if (c->s_compute) cudaStreamSynchronize(c->s_compute);
    if (c->s_copy) cudaStreamSynchronize(c->s_copy);
    for (int i = 0; i < kMaxTpl; i++) {
        cudaFree(c->d_packed[i]);
        cudaFree(c->d_toep[i]);
        cudaFree(c->d_raw[i]);
    }
    cudaFree(c->d_tpls);
    cudaFree(c->d_tasks);
    cudaFree(c->d_mm);
    cudaFree(c->d_best);
    cudaFree(c->d_nbhd);
    cudaFree(c->d_results);
    cudaFree(c->d_win8);
    cudaFree(c->d_map);
    cudaFree(c->d_surf);
    cudaFree(c->d_frame_mm);
    cudaFree(c->d_ftasks);
    for (int i = 0; i < 2; i++) {
        cudaFree(c->d_roi[i]);
        cudaFree(c->d_gptr[i]);
        cudaFreeHost(c->h_gptr[i]);
        cudaFree(c->d_jcoef[i]);
        cudaFree(c->d_jqt[i]);
        if (c->h_jcoef[i]) cudaFreeHost(c->h_jcoef[i]);
        if (c->h_jqt[i]) cudaFreeHost(c->h_jqt[i]);
    }
    cudaFree(c->d_ring[0]);
    cudaFree(c->d_ring[1]);
    cudaFree(c->d_dis10);
    cudaFree(c->d_out5);
    cudaFreeHost(c->h_tasks);
    cudaFreeHost(c->h_results);
    cudaFreeHost(c->h_surf);
    cudaFreeHost(c->h_frame_mm);
    cudaFreeHost(c->h_ftasks);
    cudaFreeHost(c->h_dis10);
    cudaFreeHost(c->h_out5);
    for (TimedLaunch &t : c->timed) {
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++)
        if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
    if (c->s_compute) cudaStreamDestroy(c->s_compute);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    delete c;
}

static int ensure_ring(lst_ctx *c, size_t bytes_per_half)
{
    if (bytes_per_half <= c->ring_cap) return LST_OK;
    CUC(c, cudaStreamSynchronize(c->s_compute));
    CUC(c, cudaStreamSynchronize(c->s_copy));
    for (int i = 0; i < 2; i++) CUC(c, regrow(&c->d_ring[i], bytes_per_half));
    c->ring_cap = bytes_per_half;
    return LST_OK;
}

int lst_set_reference(lst_ctx *c, const void *ref_frame, const float ref_xy[10], lst_reference_info *out)
{
    if (!c || !ref_frame || !ref_xy) return LST_ERR_INVALID;
    if (c->async_busy()) return c->fail(LST_ERR_STATE, "batches submitted with lst_submit are still in flight");
    CUC(c, cudaSetDevice(c->device));
    int rc = ensure_ring(c, c->frame_bytes());
    if (rc != LST_OK) return rc;
    rc = c->ensure_frames(1);
    if (rc != LST_OK) return rc;
    CUC(c, cudaMemcpyAsync(c->d_ring[0], ref_frame, c->frame_bytes(), cudaMemcpyHostToDevice, c->s_compute));
    c->stats.h2d_bytes += (int64_t)c->frame_bytes();
    c->h_surf[0] = whole_slice(c->d_ring[0], c->p.width, c->p.height);
    CUC(c, cudaMemcpyAsync(c->d_surf, c->h_surf, sizeof(Surface), cudaMemcpyHostToDevice, c->s_compute));
    if (c->frame_range_mode()) {
        std::vector<double> ur;
        ur.swap(c->user_ranges);            // consumed by this call
        rc = c->prepare_frame_ranges(1, ur.size() >= 2 ? ur.data() : nullptr);
        if (rc != LST_OK) return rc;
    }
    lst_task tasks[5];
    Rect crop[5];
    for (int k = 0; k < 5; k++) {
        c->ref_xy[2 * k] = ref_xy[2 * k];
        c->ref_xy[2 * k + 1] = ref_xy[2 * k + 1];
        c->rect[k] = template_rect(c->p, ref_xy[2 * k], ref_xy[2 * k + 1]);   // Roi.getBounds(): unclipped
        crop[k] = clip_rect(c->rect[k].x, c->rect[k].y, c->rect[k].w, c->rect[k].h, c->p.width, c->p.height);
        if (crop[k].w < 2 * kBlurRadius || crop[k].h < 2 * kBlurRadius)
            return c->fail(LST_ERR_INVALID, "template %d falls outside the slice", k);
        memset(&tasks[k], 0, sizeof(lst_task));
        tasks[k].frame = 0;
        tasks[k].tpl = k;
        tasks[k].wx = crop[k].x;
        tasks[k].wy = crop[k].y;
        tasks[k].ww = crop[k].w;
        tasks[k].wh = crop[k].h;
    }
    rc = c->run_group(tasks, 5, nullptr, true, false, -1, nullptr);
    if (rc != LST_OK) return rc;
    for (int k = 0; k < 5; k++) {
        const int planes = c->rgb3() ? 3 : 1;
        const size_t plane_px = (size_t)crop[k].w * crop[k].h;
        std::vector<uint8_t> t8(planes * plane_px);
        int slot = 0;
        while (c->order[slot] != k) slot++;
        const DevTask &d = c->h_tasks[slot];
        for (int pl = 0; pl < planes; pl++)
            CUC(c, cudaMemcpy2D(t8.data() + pl * plane_px, crop[k].w, c->d_win8 + pl * c->last_plane_bytes + d.win_off, d.pitch, crop[k].w,
                                crop[k].h, cudaMemcpyDeviceToHost));
        c->stats.d2h_bytes += (int64_t)t8.size();
        rc = planes == 3 ? c->upload_template3(k, t8.data(), crop[k].w, crop[k].h) : c->upload_template(k, t8.data(), crop[k].w, crop[k].h);
        if (rc != LST_OK) return rc;
    }
    for (int i = 0; i < 10; i++) c->dis[i] = 0.0;
    double o5[5];
    calc_displacement(c->ref_xy, c->dis, c->p.mark_dist, 0, c->spot0, o5);   // LST4:702: latches spotX0/spotY0
    c->have_ref = true;
    if (out) {
        memset(out, 0, sizeof(*out));
        for (int k = 0; k < 5; k++) {
            out->rect[k][0] = c->rect[k].x;
            out->rect[k][1] = c->rect[k].y;
            out->rect[k][2] = c->rect[k].w;
            out->rect[k][3] = c->rect[k].h;
            out->mideal[k] = c->h_tpls.c[k].mideal;
        }
        out->ref_record.frame = -1;
        out->ref_record.status = LST_STATUS_OK;
        out->ref_record.reject = -1;
        out->ref_record.dX_pix = o5[0];
        out->ref_record.dY_pix = o5[1];
        out->ref_record.x_abs = o5[2];
        out->ref_record.y_abs = o5[3];
        out->ref_record.dL = o5[4];
    }
    return LST_OK;
}

static void fill_resolve_input(const lst_ctx *c, ResolveInput *in)
{
    in->p = c->p;
    for (int k = 0; k < 5; k++) in->rect[k] = c->rect[k];
    memcpy(in->ref_xy, c->ref_xy, sizeof(in->ref_xy));
    in->spot0[0] = c->spot0[0];
    in->spot0[1] = c->spot0[1];
    in->spec_stride = const_cast<int *>(&c->spec_stride);
}

static int ensure_roi(lst_ctx *c, size_t bytes_per_half, int n_ptrs)
{
    if (bytes_per_half > c->roi_cap) {
        CUC(c, cudaStreamSynchronize(c->s_compute));
        CUC(c, cudaStreamSynchronize(c->s_copy));
        for (int i = 0; i < 2; i++) CUC(c, regrow(&c->d_roi[i], bytes_per_half));
        c->roi_cap = bytes_per_half;
    }
    if (n_ptrs > c->gptr_cap) {
        CUC(c, cudaStreamSynchronize(c->s_copy));
        for (int i = 0; i < 2; i++) {
            CUC(c, regrow(&c->d_gptr[i], n_ptrs));
            CUC(c, regrow_host(&c->h_gptr[i], n_ptrs));
        }
        c->gptr_cap = n_ptrs;
    }
    return LST_OK;
}

// Device-visible address of a pinned host pointer, or nullptr when the memory is pageable.
// cudaPointerGetAttributes costs microseconds; slices of a stack are usually re-submitted from
// the same pinned buffers (rings), so successful lookups are cached per context.
static const void *mapped_host_ptr_uncached(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return at.devicePointer;
}

static const void *mapped_host_ptr(lst_ctx *c, const void *p)
{
    const uint64_t gen = g_pin_generation.load(std::memory_order_acquire);
    if (gen != c->mapped_generation) {       // some pinned block was freed or unregistered: an alias may be stale
        c->mapped_cache.clear();
        c->mapped_generation = gen;
    }
    auto it = c->mapped_cache.find(p);
    if (it != c->mapped_cache.end()) return it->second;
    const void *d = mapped_host_ptr_uncached(p);
    if (d) {
        if (c->mapped_cache.size() > (1u << 16)) c->mapped_cache.clear();
        c->mapped_cache[p] = d;
    }
    return d;
}

// The five regions a batch can need, from the displacement state known when the copy is enqueued:
// each search window grown by `margin` pixels (the state may drift while earlier batches resolve),
// x aligned to 64 bytes so rows move as whole PCIe-friendly segments, clipped to the slice.
// Returns the bytes of one ring half for nb slices.
static size_t plan_rois(const lst_ctx *c, const double dis[10], int nb, RoiDesc out[5])
{
    const int bpp = (int)c->bpp();
    const int ax = c->roi_align_bytes / bpp;   // x origin/extent granularity in pixels
    size_t off = 0;
    for (int k = 0; k < 5; k++) {
        Rect w = window_for(c->p, c->rect[k], jint(dis[2 * k]), jint(dis[2 * k + 1]), c->p.s_area, false, nullptr);
        int x0 = w.x - c->roi_margin[k], y0 = w.y - c->roi_margin[k];
        int x1 = w.x + w.w + c->roi_margin[k], y1 = w.y + w.h + c->roi_margin[k];
        x0 = x0 < 0 ? 0 : (x0 / ax) * ax;
        x1 = ((x1 + ax - 1) / ax) * ax;
        if (x1 > c->p.width) x1 = c->p.width;
        if (y0 < 0) y0 = 0;
        if (y1 > c->p.height) y1 = c->p.height;
        out[k].x = x0;
        out[k].y = y0;
        out[k].w = x1 - x0 > 0 ? x1 - x0 : 0;
        out[k].h = y1 - y0 > 0 ? y1 - y0 : 0;
        out[k].off = off;
        out[k].slice_bytes = (size_t)out[k].w * out[k].h * bpp;
        off += ((out[k].slice_bytes * (size_t)nb) + 255) & ~(size_t)255;
    }
    return off;
}

// The tracking loop body over n_frames slices given as one strided block (base) or as pointers
// (ptrs); host memory unless on_device.  Host slices in pinned memory use the ROI ingest (only the
// search regions cross PCIe, as strided 3-D DMA copies on the copy stream while the previous batch
// computes; a gather kernel reading mapped host memory when the slices are irregularly placed or
// ingest_mode == 2); ingest_mode == 1 uses whole-slice cudaMemcpyAsync into a ring.
// Slices that are not handed to the copy engine directly: files, or host memory whose search regions the
// reader threads pack into pinned staging (pageable memory, e.g. Java arrays held through JNI; ingest_mode 3).
struct FileSource {
    const char *const *paths = nullptr;
    const int32_t *pages = nullptr;   // nullable: page 0 of every file
    bool big_endian = false;
    const void *mem_base = nullptr;   // memory source: slice i at mem_ptrs[i] or mem_base + i * mem_stride
    size_t mem_stride = 0;
    const void *const *mem_ptrs = nullptr;
    bool is_mem() const { return paths == nullptr; }
    const uint8_t *mem(int64_t i) const { return mem_ptrs ? (const uint8_t *)mem_ptrs[i] : (const uint8_t *)mem_base + (size_t)i * mem_stride; }
};

// Reads the regions `roi` of the slices [f0, f0 + nb) of a file source into `stage` (layout of the ROI
// ring), several files at a time.  whole: the entire slice i goes to stage + i * frame_bytes instead.
// JPEG slices reconstructed on the device: what the reader threads stage for a batch
struct JpegBatchPlan {
    JpegDevBatch dev;            // geometry, regions, offsets (coefficients in int16 elements, pixels in bytes)
    size_t coef_elems = 0;       // int16 elements of the whole batch
    int16_t *h_coef = nullptr;   // pinned staging of this half
    uint16_t *h_qt = nullptr;
};

static int read_slices(lst_ctx *c, const FileSource &fs, int64_t f0, int nb, const RoiDesc roi[5], bool whole, uint8_t *stage,
                       std::string *err, const JpegBatchPlan *jp = nullptr)
{
    const int nthreads = std::max(1, std::min(c->file_threads, nb));
    std::vector<std::string> errs(nthreads);
    std::vector<int> rcs(nthreads, LST_OK);
    const int want_bits = c->p.pixel_type == LST_PIX_U8 ? 8 : (c->p.pixel_type == LST_PIX_U16 ? 16 : 32);
    const size_t fb = c->frame_bytes();
    auto work = [&](int ti) {
        const int i0 = (int)((int64_t)nb * ti / nthreads), i1 = (int)((int64_t)nb * (ti + 1) / nthreads);
        for (int i = i0; i < i1; i++) {
            if (fs.is_mem()) {
                const uint8_t *src = fs.mem(f0 + i);
                const size_t bpp = c->bpp(), rowb = (size_t)c->p.width * bpp;
                if (whole) {
                    memcpy(stage + (size_t)i * fb, src, fb);
                } else {
                    for (int k = 0; k < 5; k++) {
                        const size_t nbytes = (size_t)roi[k].w * bpp;
                        uint8_t *dst = stage + roi[k].off + (size_t)i * roi[k].slice_bytes;
                        const uint8_t *s0 = src + (size_t)roi[k].y * rowb + (size_t)roi[k].x * bpp;
                        for (int r = 0; r < roi[k].h; r++) memcpy(dst + (size_t)r * nbytes, s0 + (size_t)r * rowb, nbytes);
                    }
                }
                continue;
            }
            FileMap fm;
            TiffLayout L;
            int rc = fm.open(fs.paths[f0 + i], &errs[ti]);
            if (rc == LST_OK && jp) {
                // entropy decoding only: the quantised coefficients of the MCUs under the regions go to the staging buffer
                static thread_local std::vector<uint8_t> file;
                if (file.size() < fm.size) file.resize(fm.size);
                uint64_t got = 0;
                while (got < fm.size) {
                    const ssize_t k = pread(fm.fd, file.data() + got, fm.size - got, (off_t)got);
                    if (k <= 0) break;
                    got += (uint64_t)k;
                }
                if (got < fm.size) {
                    errs[ti] = std::string("short read of ") + fs.paths[f0 + i];
                    rc = LST_ERR_INVALID;
                } else {
                    JpegRegion regs[5];
                    int16_t *dst[5];
                    for (int k = 0; k < 5; k++) {
                        const JpegDevRegion &r = jp->dev.r[k];
                        regs[k] = r.m;
                        if (r.w <= 0 || r.h <= 0) regs[k].nmx = regs[k].nmy = 0;
                        dst[k] = jp->h_coef + r.coef_off + (size_t)i * r.coef_slice;
                    }
                    rc = jpeg_decode_coeffs(file.data(), fm.size, jp->dev.g, regs, dst, 5, jp->h_qt + (size_t)i * 192, &errs[ti]);
                    if (rc != LST_OK) errs[ti] = std::string(fs.paths[f0 + i]) + ": " + errs[ti];
                }
                if (rc != LST_OK) {
                    rcs[ti] = rc;
                    return;
                }
                continue;
            }
            if (rc == LST_OK) rc = tiff_parse(fm.data, fm.size, fs.pages ? fs.pages[f0 + i] : 0, &L, &errs[ti]);
            if (rc == LST_OK && (L.width != c->p.width || L.height != c->p.height || L.bits != want_bits ||
                                 L.rgb != (c->p.pixel_type == LST_PIX_RGB) || (L.bits > 8 && L.big_endian != fs.big_endian))) {
                errs[ti] = std::string("slice file differs from the stack's geometry or byte order: ") + fs.paths[f0 + i];
                rc = LST_ERR_INVALID;   // all stack images must have the same size and type (LST4:393)
            }
            if (rc == LST_OK) {
                if (whole) {
                    rc = tiff_copy_rect(fm, L, 0, 0, L.width, L.height, stage + (size_t)i * fb, &errs[ti]);
                } else {
                    TiffRect q[5];
                    int nq = 0;
                    for (int k = 0; k < 5; k++)
                        if (roi[k].w > 0 && roi[k].h > 0)
                            q[nq++] = TiffRect{roi[k].x, roi[k].y, roi[k].w, roi[k].h, stage + roi[k].off + (size_t)i * roi[k].slice_bytes};
                    rc = tiff_copy_rects(fm, L, q, nq, &errs[ti]);
                }
            }
            if (rc != LST_OK) {
                rcs[ti] = rc;
                return;
            }
        }
    };
    std::vector<std::thread> th;
    for (int ti = 1; ti < nthreads; ti++) th.emplace_back(work, ti);
    work(0);
    for (auto &t : th) t.join();
    for (int ti = 0; ti < nthreads; ti++)
        if (rcs[ti] != LST_OK) {
            if (err) *err = errs[ti];
            return rcs[ti];
        }
    return LST_OK;
}

static int track_impl(lst_ctx *c, const void *base, size_t stride, const void *const *ptrs, bool on_device,
                      int64_t first_index, int32_t n_frames, lst_record *out, const FileSource *fs = nullptr);

// Slices of a file source tracked through whole-slice reads (whole-slice search, or a batch whose
// windows left the regions that had been read): sub-batches of whole slices in pinned memory go
// through the ordinary host path.
static int track_files_whole(lst_ctx *c, const FileSource &fs, int64_t f0, int nb, int64_t first_index, lst_record *out)
{
    const size_t fb = c->frame_bytes();
    if (fs.is_mem()) {   // the slices are in memory already: the ordinary whole-slice path reads them in place
        const int save_mode = c->p.ingest_mode;
        c->p.ingest_mode = 1;
        int rc = fs.mem_ptrs ? track_impl(c, nullptr, 0, fs.mem_ptrs + f0, false, first_index, nb, out)
                             : track_impl(c, (const uint8_t *)fs.mem_base + (size_t)f0 * fs.mem_stride, fs.mem_stride, nullptr, false,
                                          first_index, nb, out);
        c->p.ingest_mode = save_mode;
        return rc;
    }
    const int sub = (int)std::max<size_t>(1, std::min<size_t>((size_t)nb, ((size_t)256 << 20) / fb));
    if ((size_t)sub * fb > c->fwhole_cap) {
        if (c->h_fwhole) cudaFreeHost(c->h_fwhole);
        c->h_fwhole = nullptr;
        c->fwhole_cap = 0;
        CUC(c, cudaHostAlloc((void **)&c->h_fwhole, (size_t)sub * fb, cudaHostAllocDefault));
        c->fwhole_cap = (size_t)sub * fb;
    }
    for (int i = 0; i < nb; i += sub) {
        const int m = std::min(sub, nb - i);
        std::string err;
        int rc = read_slices(c, fs, f0 + i, m, nullptr, true, c->h_fwhole, &err);
        if (rc != LST_OK) return c->fail(rc, "%s", err.c_str());
        rc = track_impl(c, c->h_fwhole, fb, nullptr, false, first_index + i, m, out + i);
        if (rc != LST_OK) return rc;
    }
    return LST_OK;
}

static int track_impl(lst_ctx *c, const void *base, size_t stride, const void *const *ptrs, bool on_device,
                      int64_t first_index, int32_t n_frames, lst_record *out, const FileSource *fs)
{
    if (!c || !out || n_frames < 0 || (!base && !ptrs && !fs)) return LST_ERR_INVALID;
    if (!c->have_ref) return c->fail(LST_ERR_STATE, "lst_set_reference must be called before tracking");
    CUC(c, cudaSetDevice(c->device));
    if (n_frames == 0) return LST_OK;
    NvtxRange nvtx_call("lst_track");
    struct HostTimer {
        lst_ctx *c;
        double t0;
        ~HostTimer() { c->stats.host_ms += now_ms() - t0; }
    } host_timer{c, now_ms()};
    const size_t fb = c->frame_bytes();
    if (stride == 0) stride = fb;
    const int B = c->p.max_batch;
    // LST_RANGE_FRAME: the outermost call takes over the ranges the caller announced (nested calls -- fallbacks that hand
    // a sub-range of the slices down -- find them by frame index)
    struct DepthGuard {
        lst_ctx *c;
        ~DepthGuard()
        {
            if (--c->track_depth == 0) c->cur_ranges.clear();
        }
    } depth_guard{c};
    if (c->track_depth++ == 0) {
        c->cur_ranges.clear();
        c->cur_ranges.swap(c->user_ranges);
        if (c->cur_ranges.size() < 2 * (size_t)n_frames) c->cur_ranges.clear();
        c->range_first = first_index;
    }
    const bool frame_mode = c->frame_range_mode();
    const bool have_ranges = frame_mode && !c->cur_ranges.empty();
    const bool need_whole = frame_mode && !have_ranges;      // the slices must be reduced on the device: whole slices only
    if (need_whole && fs && !fs->is_mem())
        return c->fail(LST_ERR_UNSUPPORTED, "LST_RANGE_FRAME with slice files: give the slice ranges with lst_set_frame_ranges");
    if (fs && c->p.s_area == 0) return track_files_whole(c, *fs, 0, n_frames, first_index, out);   // whole-slice search
    if (!fs && !on_device && !need_whole && c->p.s_area > 0 && c->p.ingest_mode != 1 && c->p.ingest_mode != 2) {
        // Host-packed ROI ingest: reader threads copy the rows of the search regions into pinned staging and
        // one contiguous DMA per region follows.  Taken for pageable memory (the copy engine cannot gather
        // from it; whole-slice copies would move 14x the bytes) and on request (ingest_mode 3, LST_HOST_PACK).
        static const int force = [] {
            const char *e = getenv("LST_HOST_PACK");
            return e ? atoi(e) : 0;
        }();
        const void *first = ptrs ? ptrs[0] : base;
        if (c->p.ingest_mode == 3 || force || !mapped_host_ptr(c, first)) {
            FileSource ms;
            ms.mem_base = base;
            ms.mem_stride = stride;
            ms.mem_ptrs = ptrs;
            unsigned hc = std::thread::hardware_concurrency();
            c->file_threads = (int)std::min(16u, std::max(1u, hc));
            if (const char *e = getenv("LST_FILE_THREADS")) c->file_threads = std::max(1, atoi(e));
            return track_impl(c, base, stride, ptrs, false, first_index, n_frames, out, &ms);
        }
    }
    ResolveInput in;
    fill_resolve_input(c, &in);
    auto src_of = [&](int64_t i) -> const void * { return ptrs ? ptrs[i] : (const uint8_t *)base + (size_t)i * stride; };
    // Batch schedule.  Host slices: the ingest of batch b+1 overlaps the kernels of batch b, so only the
    // first batch's ingest is exposed -- start small (128 slices) and double up to max_batch.
    std::vector<int64_t> bstart;
    {
        int64_t f = 0;
        int sz = on_device ? B : std::min(B, 128);
        while (f < n_frames) {
            bstart.push_back(f);
            f += sz;
            sz = std::min(B, sz * 2);
        }
        bstart.push_back(n_frames);
    }
    const int nbatches = (int)bstart.size() - 1;

    bool use_roi = !on_device && !need_whole && c->p.s_area > 0 && (fs || c->p.ingest_mode != 1);
    std::vector<const void *> mapped;
    std::string file_err[2];          // declared first: the reader threads of file_job write them and are joined first
    std::future<int> file_job[2];
    JpegGeom jgeom;
    memset(&jgeom, 0, sizeof(jgeom));
    bool jpeg_stack = false;          // the slices are JPEG files reconstructed on the device
    if (use_roi && !fs) {
        mapped.resize(n_frames);
        for (int64_t i = 0; i < n_frames && use_roi; i++) {
            mapped[i] = mapped_host_ptr(c, src_of(i));
            if (!mapped[i]) use_roi = false;     // pageable memory: whole-slice copies
        }
    }
    int rc = c->ensure_frames(use_roi ? 6 * B : B);
    if (rc != LST_OK) return rc;
    // the ROI ring is sized once from an upper bound of plan_rois (both halves are live at once)
    size_t roi_block_max = 0;
    if (use_roi) {
        const int ax = c->roi_align_bytes / (int)c->bpp();
        for (int k = 0; k < 5; k++) {
            const int mg = std::max(c->roi_margin[k], kRoiMarginMax);   // the margin adapts up to this
            size_t w = (size_t)c->rect[k].w + 2 * c->p.s_area + 2 * mg + 2 * ax;
            size_t h = (size_t)c->rect[k].h + 2 * c->p.s_area + 2 * mg;
            w = std::min<size_t>(w, (size_t)c->p.width + ax);
            h = std::min<size_t>(h, (size_t)c->p.height);
            roi_block_max += ((w * h * c->bpp()) + 255) & ~(size_t)255;
        }
        rc = ensure_roi(c, (roi_block_max + 256) * (size_t)std::min<int64_t>(B, n_frames), B);
        if (rc != LST_OK) return rc;
        // JPEG files: the reader threads do the entropy decoding, the device the reconstruction of the regions.  The
        // coefficient staging is sized once per call from the same upper bound as the ROI ring.
        if (fs && !fs->is_mem() && c->jpeg_device) {
            FileMap fm;
            std::string e;
            if (fm.open(fs->paths[0], &e) == LST_OK && jpeg_signature(fm.data, fm.size) && jpeg_geometry(fm.data, fm.size, &jgeom, &e) == LST_OK &&
                jgeom.width == c->p.width && jgeom.height == c->p.height && (jgeom.ncomp == 3) == (c->p.pixel_type == LST_PIX_RGB))
                jpeg_stack = true;
        }
        if (jpeg_stack) {
            const int ax = c->roi_align_bytes / (int)c->bpp(), mw = 8 * jgeom.hmax, mh = 8 * jgeom.vmax;
            size_t elems = 0;
            for (int k = 0; k < 5; k++) {
                const int mg = std::max(c->roi_margin[k], kRoiMarginMax);
                size_t w = std::min<size_t>((size_t)c->rect[k].w + 2 * c->p.s_area + 2 * mg + 2 * ax, (size_t)c->p.width + ax);
                size_t h = std::min<size_t>((size_t)c->rect[k].h + 2 * c->p.s_area + 2 * mg, (size_t)c->p.height);
                JpegRegion bound;
                bound.mx0 = bound.my0 = 0, bound.nmx = (int)(w / mw + 3), bound.nmy = (int)(h / mh + 3);
                if (jpeg_reconstruct_smem(jgeom, bound) > 200 * 1024) jpeg_stack = false;   // a region may outgrow one CTA: host path
                elems += (size_t)bound.nmx * bound.nmy * (size_t)jpeg_blocks_per_mcu(jgeom) * 64;
            }
            const int nbmax = (int)std::min<int64_t>(B, n_frames);
            elems *= (size_t)nbmax;
            if (jpeg_stack && (elems > c->jcoef_cap || nbmax > c->jqt_cap)) {
                CUC(c, cudaStreamSynchronize(c->s_copy));
                CUC(c, cudaStreamSynchronize(c->s_compute));
                const size_t cap = std::max(elems, c->jcoef_cap);
                const int qcap = std::max(nbmax, c->jqt_cap);
                for (int i = 0; i < 2; i++) {
                    if (c->h_jcoef[i]) cudaFreeHost(c->h_jcoef[i]);
                    if (c->h_jqt[i]) cudaFreeHost(c->h_jqt[i]);
                    c->h_jcoef[i] = nullptr, c->h_jqt[i] = nullptr;
                    c->jcoef_cap = 0, c->jqt_cap = 0;
                    CUC(c, cudaHostAlloc((void **)&c->h_jcoef[i], cap * sizeof(int16_t), cudaHostAllocDefault));
                    CUC(c, cudaHostAlloc((void **)&c->h_jqt[i], (size_t)qcap * 192 * sizeof(uint16_t), cudaHostAllocDefault));
                    CUC(c, regrow(&c->d_jcoef[i], cap));
                    CUC(c, regrow(&c->d_jqt[i], (size_t)qcap * 192));
                }
                c->jcoef_cap = cap;
                c->jqt_cap = qcap;
            }
        }
        if (fs && !jpeg_stack && c->fstage_cap < c->roi_cap) {
            for (int i = 0; i < 2; i++) {
                if (c->h_fstage[i]) cudaFreeHost(c->h_fstage[i]);
                c->h_fstage[i] = nullptr;
            }
            c->fstage_cap = 0;
            for (int i = 0; i < 2; i++) CUC(c, cudaHostAlloc((void **)&c->h_fstage[i], c->roi_cap, cudaHostAllocDefault));
            c->fstage_cap = c->roi_cap;
        }
    }
    if (!on_device && !use_roi) {
        rc = ensure_ring(c, (size_t)std::min<int64_t>(B, n_frames) * fb);
        if (rc != LST_OK) return rc;
    }

    auto enqueue_copy = [&](int b) -> int {
        int64_t f0 = bstart[b];
        int nb = (int)(bstart[b + 1] - f0);
        const int half = b & 1;
        if (use_roi) {
            size_t need = plan_rois(c, c->dis, nb, c->roi[half]);
            if (need > c->roi_cap) return c->fail(LST_ERR_STATE, "internal: ROI ring smaller than planned");
            if (fs) {
                // A reader thread maps the slice files, copies the rows of the five regions into the pinned
                // staging half and then queues the (contiguous) copies to the device itself, so that
                // they overlap the kernels of the batch being resolved.  Joined before the batch is used.
                RoiDesc rois[5];
                for (int k = 0; k < 5; k++) {
                    rois[k] = c->roi[half][k];
                    c->stats.h2d_bytes += (int64_t)(rois[k].slice_bytes * nb);
                }
                lst_ctx *cc = c;
                const FileSource src = *fs;
                std::string *errp = &file_err[half];
                // JPEG files: the reader threads do the entropy decoding, the device the reconstruction of the regions
                JpegBatchPlan jp;
                bool jpeg_dev = jpeg_stack;
                if (jpeg_dev) {
                    jp.dev.g = jgeom;
                    const int bpm = jpeg_blocks_per_mcu(jgeom);
                    size_t off = 0;
                    for (int k = 0; k < 5; k++) {
                        JpegDevRegion &r = jp.dev.r[k];
                        memset(&r, 0, sizeof(r));
                        if (rois[k].w <= 0 || rois[k].h <= 0) continue;
                        r.x = rois[k].x, r.y = rois[k].y, r.w = rois[k].w, r.h = rois[k].h;
                        r.m = jpeg_region(jgeom, r.x, r.y, r.w, r.h);
                        r.coef_slice = (unsigned long long)r.m.nmx * r.m.nmy * bpm * 64;
                        r.coef_off = off;
                        off += (size_t)r.coef_slice * nb;
                        r.out_off = rois[k].off;
                        r.out_slice = rois[k].slice_bytes;
                        if (jpeg_reconstruct_smem(jgeom, r.m) > 200 * 1024) return c->fail(LST_ERR_STATE, "internal: JPEG region larger than planned");
                    }
                    jp.coef_elems = off;
                    if (off > c->jcoef_cap || nb > c->jqt_cap) return c->fail(LST_ERR_STATE, "internal: JPEG coefficient staging smaller than planned");
                }
                if (jpeg_dev) {
                    jp.h_coef = c->h_jcoef[half];
                    jp.h_qt = c->h_jqt[half];
                    c->stats.h2d_bytes += (int64_t)(jp.coef_elems * sizeof(int16_t) + (size_t)nb * 192 * sizeof(uint16_t));
                    for (int k = 0; k < 5; k++) c->stats.h2d_bytes -= (int64_t)(rois[k].slice_bytes * nb);   // no pixels travel
                    c->stats.kernel_launches += 1;
                    file_job[half] = std::async(std::launch::async, [cc, src, f0, nb, half, rois, errp, jp]() -> int {
                        int r = read_slices(cc, src, f0, nb, rois, false, nullptr, errp, &jp);
                        if (r != LST_OK) return r;
                        if (cudaSetDevice(cc->device) != cudaSuccess) return LST_ERR_CUDA;
                        if (cudaMemcpyAsync(cc->d_jcoef[half], jp.h_coef, jp.coef_elems * sizeof(int16_t), cudaMemcpyHostToDevice, cc->s_copy) != cudaSuccess ||
                            cudaMemcpyAsync(cc->d_jqt[half], jp.h_qt, (size_t)nb * 192 * sizeof(uint16_t), cudaMemcpyHostToDevice, cc->s_copy) != cudaSuccess)
                            return LST_ERR_CUDA;
                        if (launch_jpeg_reconstruct(cc->d_jcoef[half], cc->d_jqt[half], jp.dev, nb, cc->d_roi[half], cc->s_copy) < 0) return LST_ERR_CUDA;
                        return cudaEventRecord(cc->ev_copy[half], cc->s_copy) == cudaSuccess ? LST_OK : LST_ERR_CUDA;
                    });
                    return LST_OK;
                }
                file_job[half] = std::async(std::launch::async, [cc, src, f0, nb, half, rois, errp]() -> int {
                    int r = read_slices(cc, src, f0, nb, rois, false, cc->h_fstage[half], errp);
                    if (r != LST_OK) return r;
                    if (cudaSetDevice(cc->device) != cudaSuccess) return LST_ERR_CUDA;
                    for (int k = 0; k < 5; k++) {
                        const size_t bytes = (size_t)rois[k].slice_bytes * nb;
                        if (bytes == 0) continue;
                        if (cudaMemcpyAsync(cc->d_roi[half] + rois[k].off, cc->h_fstage[half] + rois[k].off, bytes,
                                            cudaMemcpyHostToDevice, cc->s_copy) != cudaSuccess)
                            return LST_ERR_CUDA;
                    }
                    return cudaEventRecord(cc->ev_copy[half], cc->s_copy) == cudaSuccess ? LST_OK : LST_ERR_CUDA;
                });
                return LST_OK;
            }
            // Runs of equally spaced slices move with one strided 3-D DMA per region (copy engine: no SM
            // is taken from the kernels of the previous batch); irregularly placed slices (or
            // ingest_mode == 2) are gathered by a kernel reading the mapped host memory.
            const size_t rowb = (size_t)c->p.width * c->bpp();
            std::vector<std::pair<int, int>> runs;   // [first, last) of constant positive stride
            bool dma = c->p.ingest_mode != 2;
            auto at = [&](int i) { return (const uint8_t *)src_of(f0 + i); };
            for (int i = 0; i < nb && dma;) {
                int j = i + 1;
                if (j < nb) {
                    const ptrdiff_t st = at(j) - at(i);
                    if (st > 0 && (size_t)st % rowb == 0 && (size_t)st / rowb >= (size_t)c->p.height) {
                        j++;
                        while (j < nb && at(j) - at(j - 1) == st) j++;
                    }
                }
                runs.push_back({i, j});
                i = j;
                if ((int)runs.size() > 64 + nb / 8) dma = false;   // too fragmented: one kernel is cheaper
            }
            if (dma) {
                for (auto &rn : runs) {
                    const int i0 = rn.first, len = rn.second - rn.first;
                    const uint8_t *p0 = (const uint8_t *)src_of(f0 + i0);
                    size_t ysize = (size_t)c->p.height;
                    if (len > 1) ysize = (size_t)((const uint8_t *)src_of(f0 + i0 + 1) - p0) / rowb;
                    for (int k = 0; k < 5; k++) {
                        const RoiDesc &r = c->roi[half][k];
                        if (r.w <= 0 || r.h <= 0) continue;
                        cudaMemcpy3DParms pr;
                        memset(&pr, 0, sizeof(pr));
                        pr.srcPtr = make_cudaPitchedPtr((void *)p0, rowb, rowb, ysize);
                        pr.srcPos = make_cudaPos((size_t)r.x * c->bpp(), (size_t)r.y, 0);
                        pr.dstPtr = make_cudaPitchedPtr(c->d_roi[half] + r.off + (size_t)i0 * r.slice_bytes,
                                                        (size_t)r.w * c->bpp(), (size_t)r.w * c->bpp(), (size_t)r.h);
                        pr.dstPos = make_cudaPos(0, 0, 0);
                        pr.extent = make_cudaExtent((size_t)r.w * c->bpp(), (size_t)r.h, (size_t)len);
                        pr.kind = cudaMemcpyHostToDevice;
                        CUC(c, cudaMemcpy3DAsync(&pr, c->s_copy));
                    }
                }
            } else {
                for (int i = 0; i < nb; i++) c->h_gptr[half][i] = mapped[f0 + i];
                CUC(c, cudaMemcpyAsync(c->d_gptr[half], c->h_gptr[half], sizeof(void *) * nb, cudaMemcpyHostToDevice, c->s_copy));
                c->stats.kernel_launches += launch_gather_rois(c->d_gptr[half], nb, c->roi[half], c->p.width, (int)c->bpp(),
                                                               c->d_roi[half], c->s_copy);
                CUC(c, cudaGetLastError());
                c->stats.h2d_bytes += (int64_t)(sizeof(void *) * nb);
            }
            for (int k = 0; k < 5; k++) c->stats.h2d_bytes += (int64_t)(c->roi[half][k].slice_bytes * nb);
        } else {
            uint8_t *dst = c->d_ring[half];
            if (!ptrs && stride == fb) {
                CUC(c, cudaMemcpyAsync(dst, src_of(f0), (size_t)nb * fb, cudaMemcpyHostToDevice, c->s_copy));
            } else {
                for (int i = 0; i < nb; i++)
                    CUC(c, cudaMemcpyAsync(dst + (size_t)i * fb, src_of(f0 + i), fb, cudaMemcpyHostToDevice, c->s_copy));
            }
            c->stats.h2d_bytes += (int64_t)nb * (int64_t)fb;
        }
        CUC(c, cudaEventRecord(c->ev_copy[half], c->s_copy));
        return LST_OK;
    };

    if (!on_device) {
        rc = enqueue_copy(0);
        if (rc != LST_OK) return rc;
    }
    for (int b = 0; b < nbatches; b++) {
        int64_t f0 = bstart[b];
        int nb = (int)(bstart[b + 1] - f0);
        const int half = b & 1;
        int nsurf = nb;
        if (!on_device) {
            if (b + 1 < nbatches) {   // double buffering: next batch's ingest overlaps this batch's kernels
                rc = enqueue_copy(b + 1);
                if (rc != LST_OK) return rc;
            }
            if (fs && file_job[half].valid()) {
                rc = file_job[half].get();
                if (rc != LST_OK) {
                    if (file_job[half ^ 1].valid()) file_job[half ^ 1].wait();
                    return c->fail(rc, "%s", file_err[half].empty() ? "reading slice files failed" : file_err[half].c_str());
                }
            }
            CUC(c, cudaStreamWaitEvent(c->s_compute, c->ev_copy[half], 0));
        }
        if (use_roi) {
            for (int i = 0; i < nb; i++) {
                for (int k = 0; k < 5; k++) {
                    const RoiDesc &r = c->roi[half][k];
                    Surface &sf = c->h_surf[i * 5 + k];
                    sf.ptr = c->d_roi[half] + r.off + (size_t)i * r.slice_bytes;
                    sf.pitch_px = r.w;
                    sf.org_x = r.x;
                    sf.org_y = r.y;
                    sf.rows = r.h;
                }
                c->h_surf[5 * nb + i] = whole_slice(fs ? nullptr : mapped[f0 + i], c->p.width, fs ? 0 : c->p.height);   // zero-copy fallback
            }
            nsurf = 6 * nb;
        } else if (!on_device) {
            for (int i = 0; i < nb; i++) c->h_surf[i] = whole_slice(c->d_ring[half] + (size_t)i * fb, c->p.width, c->p.height);
        } else {
            for (int i = 0; i < nb; i++) c->h_surf[i] = whole_slice(src_of(f0 + i), c->p.width, c->p.height);
        }
        if (launch_upload(c->d_surf, c->h_surf, sizeof(Surface) * nsurf, c->s_compute) < 0)
            return c->fail(LST_ERR_CUDA, "surface upload failed");
        c->stats.kernel_launches += 1;
        c->stats.h2d_bytes += (int64_t)sizeof(Surface) * nsurf;
        if (frame_mode) {
            const int64_t r0 = first_index + f0 - c->range_first;
            const bool ok = have_ranges && r0 >= 0 && (size_t)(2 * (r0 + nb)) <= c->cur_ranges.size();
            if (!ok && use_roi) return c->fail(LST_ERR_STATE, "internal: slice ranges missing for a region-ingest batch");
            rc = c->prepare_frame_ranges(nb, ok ? &c->cur_ranges[2 * r0] : nullptr);
            if (rc != LST_OK) return rc;
        }
        c->roi_active = use_roi;
        c->roi_half = half;
        c->roi_nb = nb;
        c->file_mode = fs != nullptr;
        {
            NvtxRange nvtx_batch("lst batch");
            rc = resolve_batch(in, c->dis, first_index + f0, nb, *c, out + f0, &c->stats);
        }
        c->roi_active = false;
        c->file_mode = false;
        if (rc == kRetryWholeSlices && fs) {
            // a window left the regions that were read: this batch again from whole slices.  The read-ahead
            // of the next batch is drained first (the nested call reuses the copy stream and events) and
            // re-planned afterwards from the state this batch leaves.
            if (file_job[half ^ 1].valid()) file_job[half ^ 1].wait();
            CUC(c, cudaStreamSynchronize(c->s_copy));
            CUC(c, cudaStreamSynchronize(c->s_compute));
            const int save_mode = c->p.ingest_mode;
            c->p.ingest_mode = 1;
            rc = track_files_whole(c, *fs, f0, nb, first_index + f0, out + f0);
            c->p.ingest_mode = save_mode;
            if (rc != LST_OK) return rc;
            c->stats.roi_fallback_tasks += nb;
            if (b + 1 < nbatches) {
                rc = enqueue_copy(b + 1);
                if (rc != LST_OK) return rc;
            }
            continue;
        }
        if (rc != LST_OK) {
            if (fs)
                for (int h2 = 0; h2 < 2; h2++)
                    if (file_job[h2].valid()) file_job[h2].wait();
            if (c->err.empty()) c->fail(rc, "resolver failed (%d)", rc);
            return rc;
        }
        if (use_roi && c->roi_margin_fixed < 0) {
            // adapt the margins: how far did each template's windows of this batch stray from where its
            // region was planned?  The next plan gets that excursion + 3 px (4..kRoiMarginMax) per region:
            // the marks of a rigid holder barely move while the spot may travel several pixels per batch.
            int exc[5] = {0, 0, 0, 0, 0};
            for (int i = 0; i < nb; i++) {
                const lst_record &r = out[f0 + i];
                for (int k = 0; k < 5; k++) {
                    if (r.tm[k].ww == 0) continue;
                    const RoiDesc &rd = c->roi[half][k];
                    int inner_x = r.tm[k].wx - rd.x, inner_y = r.tm[k].wy - rd.y;
                    int outer_x = rd.x + rd.w - (r.tm[k].wx + r.tm[k].ww), outer_y = rd.y + rd.h - (r.tm[k].wy + r.tm[k].wh);
                    int slack = std::min(std::min(inner_x, inner_y), std::min(outer_x, outer_y));   // < 0: fell outside
                    exc[k] = std::max(exc[k], c->roi_margin[k] - slack);
                }
            }
            for (int k = 0; k < 5; k++) c->roi_margin[k] = std::min(kRoiMarginMax, std::max(4, exc[k] + 3));
        }
    }
    return LST_OK;
}

int lst_track(lst_ctx *c, const void *frames, size_t frame_stride_bytes, int64_t first_frame_index, int32_t n_frames,
              lst_record *out)
{
    return track_impl(c, frames, frame_stride_bytes, nullptr, false, first_frame_index, n_frames, out);
}
int lst_track_ptrs(lst_ctx *c, const void *const *frame_ptrs, int64_t first_frame_index, int32_t n_frames, lst_record *out)
{
    return track_impl(c, nullptr, 0, frame_ptrs, false, first_frame_index, n_frames, out);
}
int lst_track_device(lst_ctx *c, const void *dev_frames, size_t frame_stride_bytes, int64_t first_frame_index,
                     int32_t n_frames, lst_record *out)
{
    return track_impl(c, dev_frames, frame_stride_bytes, nullptr, true, first_frame_index, n_frames, out);
}

int lst_track_device_ptrs(lst_ctx *c, const void *const *dev_frame_ptrs, int64_t first_frame_index, int32_t n_frames,
                          lst_record *out)
{
    return track_impl(c, nullptr, 0, dev_frame_ptrs, true, first_frame_index, n_frames, out);
}

// ---- output side -------------------------------------------------------------------
int64_t lst_results_rows(const lst_record *records, int64_t n_records, double time0, double time_step,
                         int64_t first_row_number, double *rows, int64_t *src_index)
{
    if (!records || !rows || n_records < 0) return LST_ERR_INVALID;
    int64_t n = 0;
    double seconds = time0;
    for (int64_t i = 0; i < n_records; i++) {
        const lst_record &r = records[i];
        if (r.status != LST_STATUS_OK) continue;   // `continue` of LST4:845: no row, no time step
        seconds += time_step;
        double *o = rows + n * 7;
        o[0] = seconds;
        o[1] = (double)(first_row_number + n);
        o[2] = r.dX_pix;
        o[3] = r.dY_pix;
        o[4] = r.x_abs;
        o[5] = r.y_abs;
        o[6] = r.dL;
        if (src_index) src_index[n] = i;
        n++;
    }
    return n;
}

// ---- slices as files -------------------------------------------------------------
int lst_tiff_probe(const char *path, int32_t page, lst_tiff_info *out, char *err, int32_t err_cap)
{
    if (!path || !out) return LST_ERR_INVALID;
    FileMap fm;
    TiffLayout L;
    std::string e;
    int rc = fm.open(path, &e);
    if (rc == LST_OK) rc = tiff_parse(fm.data, fm.size, page, &L, &e);
    if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s", e.c_str());
    if (rc != LST_OK) return rc;
    out->width = L.width;
    out->height = L.height;
    out->bits = L.rgb ? 24 : L.bits;        // 24 = ColorProcessor pixels (one int32 0x00RRGGBB each), ImageJ's bit depth for RGB
    out->sample_format = L.sample_format;
    out->big_endian = L.big_endian ? 1 : 0;
    out->rows_per_strip = L.rows_per_strip;
    out->n_strips = (int32_t)L.strip_off.size();
    out->imagej_images = L.imagej_images;
    out->first_strip_offset = (int64_t)L.strip_off[0];
    out->compression = L.compression;
    out->predictor = L.predictor;
    return LST_OK;
}

int lst_tiff_read_rect(const char *path, int32_t page, int32_t x, int32_t y, int32_t w, int32_t h, void *out, char *err,
                       int32_t err_cap)
{
    if (!path || !out) return LST_ERR_INVALID;
    FileMap fm;
    TiffLayout L;
    std::string e;
    int rc = fm.open(path, &e);
    if (rc == LST_OK) rc = tiff_parse(fm.data, fm.size, page, &L, &e);
    if (rc == LST_OK) rc = tiff_copy_rect(fm, L, x, y, w, h, (uint8_t *)out, &e);
    if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s", e.c_str());
    if (rc != LST_OK) return rc;
    uint8_t *px = (uint8_t *)out;
    const size_t nbytes = (size_t)w * h * (L.bits / 8);
    if (L.big_endian && L.bits == 16)
        for (size_t i = 0; i + 1 < nbytes; i += 2) std::swap(px[i], px[i + 1]);
    else if (L.big_endian && L.bits == 32)
        for (size_t i = 0; i + 3 < nbytes; i += 4) {
            std::swap(px[i], px[i + 3]);
            std::swap(px[i + 1], px[i + 2]);
        }
    return LST_OK;
}

static int check_slice_file(lst_ctx *c, const char *path, int page, FileMap *fm, TiffLayout *L)
{
    std::string e;
    int rc = fm->open(path, &e);
    if (rc == LST_OK) rc = tiff_parse(fm->data, fm->size, page, L, &e);
    if (rc != LST_OK) return c->fail(rc, "%s: %s", path, e.c_str());
    const int want_bits = c->p.pixel_type == LST_PIX_U8 ? 8 : (c->p.pixel_type == LST_PIX_U16 ? 16 : 32);
    if (L->width != c->p.width || L->height != c->p.height || L->bits != want_bits || L->rgb != (c->p.pixel_type == LST_PIX_RGB))
        return c->fail(LST_ERR_INVALID, "%s is %dx%d %s%d-bit, the context was created for %dx%d %s%d-bit", path, L->width, L->height,
                       L->rgb ? "RGB " : "", L->bits, c->p.width, c->p.height, c->p.pixel_type == LST_PIX_RGB ? "RGB " : "", want_bits);
    return LST_OK;
}

int lst_set_reference_file(lst_ctx *c, const char *path, int32_t page, const float ref_xy[10], lst_reference_info *out)
{
    if (!c || !path || !ref_xy) return LST_ERR_INVALID;
    FileMap fm;
    TiffLayout L;
    int rc = check_slice_file(c, path, page, &fm, &L);
    if (rc != LST_OK) return rc;
    std::vector<uint8_t> px(c->frame_bytes());
    std::string e;
    rc = tiff_copy_rect(fm, L, 0, 0, L.width, L.height, px.data(), &e);
    if (rc != LST_OK) return c->fail(rc, "%s: %s", path, e.c_str());
    if (L.big_endian && L.bits == 16) {
        for (size_t i = 0; i + 1 < px.size(); i += 2) std::swap(px[i], px[i + 1]);
    } else if (L.big_endian && L.bits == 32) {
        for (size_t i = 0; i + 3 < px.size(); i += 4) {
            std::swap(px[i], px[i + 3]);
            std::swap(px[i + 1], px[i + 2]);
        }
    }
    return lst_set_reference(c, px.data(), ref_xy, out);
}

int lst_track_files(lst_ctx *c, const char *const *paths, const int32_t *pages, int64_t first_frame_index, int32_t n_frames,
                    lst_record *out)
{
    if (!c || !paths || !out || n_frames < 0) return LST_ERR_INVALID;
    if (n_frames == 0) return LST_OK;
    // the first file fixes the byte order of the stack (every file is checked against it while reading)
    FileSource fs;
    fs.paths = paths;
    fs.pages = pages;
    {
        FileMap fm;
        TiffLayout L;
        int rc = check_slice_file(c, paths[0], pages ? pages[0] : 0, &fm, &L);
        if (rc != LST_OK) return rc;
        fs.big_endian = L.big_endian;
    }
    {
        unsigned hc = std::thread::hardware_concurrency();
        c->file_threads = (int)std::min(16u, std::max(1u, hc));   // pread scales with the cores; the copy engine does the rest
        if (const char *e = getenv("LST_FILE_THREADS")) c->file_threads = std::max(1, atoi(e));
        if (const char *e = getenv("LST_JPEG_DEVICE")) c->jpeg_device = atoi(e) != 0;   // 0: JPEG slices are reconstructed by the reader threads
    }
    c->swap_bytes = fs.big_endian && c->p.pixel_type != LST_PIX_U8;
    int rc = track_impl(c, nullptr, 0, nullptr, false, first_frame_index, n_frames, out, &fs);
    c->swap_bytes = false;
    return rc;
}

// ---- several GPUs in one process ------------------------------------------------------------------
struct lst_multi {
    std::vector<lst_ctx *> ctx;
    std::string err;
    double dis[10] = {0};
    int64_t retracked = 0;
    int lead_in = 2;
};

// Do two displacement states give the same records from here on?  The windows depend on (int)dis only (LST4:1327-1328).
// With search-area doubling the un-truncated displacement of mark1 also enters: when mark1 is lost and found again its
// shift against the *previous* dis moves the other four windows (LST4:1757: int += double), so there the two mark1
// components must agree exactly -- a fraction that differs could move a window by a pixel.
static bool same_windows(const double a[10], const double b[10], bool doubling)
{
    for (int i = 0; i < 10; i++)
        if (jint(a[i]) != jint(b[i])) return false;
    if (doubling && (a[2] != b[2] || a[3] != b[3])) return false;
    return true;
}

int lst_multi_create(const lst_params *params, const int32_t *devices, int32_t n_devices, lst_multi **out)
{
    if (!params || !devices || n_devices < 1 || !out) return LST_ERR_INVALID;
    *out = nullptr;
    lst_multi *m = new lst_multi();
    if (const char *e = getenv("LST_MULTI_LEAD_IN")) m->lead_in = std::max(0, atoi(e));   // 0: chunks start from the state at call time (tests of the re-track path)
    for (int i = 0; i < n_devices; i++) {
        lst_ctx *c = nullptr;
        int rc = lst_create(params, devices[i], &c);
        if (rc != LST_OK) {
            lst_multi_destroy(m);
            return rc;   // lst_last_error(NULL) holds the message
        }
        m->ctx.push_back(c);
    }
    *out = m;
    return LST_OK;
}

void lst_multi_destroy(lst_multi *m)
{
    if (!m) return;
    for (lst_ctx *c : m->ctx) lst_destroy(c);
    delete m;
}

const char *lst_multi_last_error(const lst_multi *m) { return m ? m->err.c_str() : g_create_error.c_str(); }

int lst_multi_set_reference(lst_multi *m, const void *ref_frame, const float ref_xy[10], lst_reference_info *out)
{
    if (!m) return LST_ERR_INVALID;
    for (size_t i = 0; i < m->ctx.size(); i++) {
        int rc = lst_set_reference(m->ctx[i], ref_frame, ref_xy, i == 0 ? out : nullptr);
        if (rc != LST_OK) {
            m->err = m->ctx[i]->err;
            return rc;
        }
    }
    for (int i = 0; i < 10; i++) m->dis[i] = 0.0;
    return LST_OK;
}

int lst_multi_get_state(const lst_multi *m, double dis[10])
{
    if (!m || !dis) return LST_ERR_INVALID;
    memcpy(dis, m->dis, sizeof(m->dis));
    return LST_OK;
}
int lst_multi_set_state(lst_multi *m, const double dis[10])
{
    if (!m || !dis) return LST_ERR_INVALID;
    memcpy(m->dis, dis, sizeof(m->dis));
    return LST_OK;
}
int64_t lst_multi_retracked(const lst_multi *m) { return m ? m->retracked : 0; }

static int multi_track(lst_multi *m, const void *const *frame_ptrs, bool on_device, int64_t first_frame_index, int32_t n_frames,
                       lst_record *out);

int lst_multi_track(lst_multi *m, const void *const *frame_ptrs, int64_t first_frame_index, int32_t n_frames,
                    lst_record *out)
{
    return multi_track(m, frame_ptrs, false, first_frame_index, n_frames, out);
}

int lst_multi_track_device(lst_multi *m, const void *const *dev_frame_ptrs, int64_t first_frame_index, int32_t n_frames,
                           lst_record *out)
{
    return multi_track(m, dev_frame_ptrs, true, first_frame_index, n_frames, out);
}

static int multi_track(lst_multi *m, const void *const *frame_ptrs, bool on_device, int64_t first_frame_index, int32_t n_frames,
                       lst_record *out)
{
    if (!m || !frame_ptrs || !out || n_frames < 0) return LST_ERR_INVALID;
    auto track = [on_device](lst_ctx *c, const void *const *p, int64_t first, int32_t n, lst_record *o) {
        return on_device ? lst_track_device_ptrs(c, p, first, n, o) : lst_track_ptrs(c, p, first, n, o);
    };
    const int G = (int)m->ctx.size();
    if (n_frames == 0) return LST_OK;
    // contiguous chunks, sizes differing by at most one slice
    std::vector<int> lo(G + 1, 0);
    for (int r = 0; r < G; r++) lo[r + 1] = lo[r] + n_frames / G + (r < n_frames % G ? 1 : 0);
    std::vector<int> rcs(G, LST_OK);
    std::vector<std::vector<double>> guess(G, std::vector<double>(10)), end(G, std::vector<double>(10));
    auto work = [&](int r) {
        lst_ctx *c = m->ctx[r];
        const int n = lo[r + 1] - lo[r];
        lst_set_state(c, m->dis);
        if (r > 0 && n > 0) {   // refine the speculated start state on the last slices of the previous chunk
            int li = std::min(m->lead_in, lo[r]);
            if (li > 0) {
                std::vector<lst_record> tmp(li);
                rcs[r] = track(c, frame_ptrs + lo[r] - li, first_frame_index + lo[r] - li, li, tmp.data());
                if (rcs[r] != LST_OK) return;
            }
        }
        lst_get_state(c, guess[r].data());
        if (n > 0) rcs[r] = track(c, frame_ptrs + lo[r], first_frame_index + lo[r], n, out + lo[r]);
        lst_get_state(c, end[r].data());
    };
    {
        std::vector<std::thread> th;
        for (int r = 1; r < G; r++) th.emplace_back(work, r);
        work(0);
        for (auto &t : th) t.join();
    }
    for (int r = 0; r < G; r++)
        if (rcs[r] != LST_OK) {
            m->err = m->ctx[r]->err;
            return rcs[r];
        }
    // boundary verification in order (host side, 10 doubles per boundary)
    double verified[10];
    memcpy(verified, end[0].data(), sizeof(verified));
    if (lo[1] - lo[0] == 0) memcpy(verified, m->dis, sizeof(verified));
    for (int r = 1; r < G; r++) {
        const int n = lo[r + 1] - lo[r];
        if (n == 0) continue;
        lst_ctx *c = m->ctx[r];
        bool any_ok = false;
        if (!same_windows(verified, guess[r].data(), c->p.doubling != 0)) {
            lst_set_state(c, verified);
            int rc = track(c, frame_ptrs + lo[r], first_frame_index + lo[r], n, out + lo[r]);
            if (rc != LST_OK) {
                m->err = c->err;
                return rc;
            }
            lst_get_state(c, end[r].data());
            m->retracked++;
            for (int j = lo[r]; j < lo[r + 1]; j++) any_ok |= out[j].status == LST_STATUS_OK;
        } else {
            // same windows => same matches; only the carried state of leading skipped slices differs
            for (int j = lo[r]; j < lo[r + 1]; j++) {
                if (out[j].status == LST_STATUS_OK) {
                    any_ok = true;
                    break;
                }
                memcpy(out[j].dis, verified, sizeof(verified));
            }
        }
        if (any_ok) memcpy(verified, end[r].data(), sizeof(verified));
    }
    memcpy(m->dis, verified, sizeof(verified));
    for (lst_ctx *c : m->ctx) lst_set_state(c, verified);
    return LST_OK;
}

int lst_multi_get_stats(const lst_multi *m, lst_stats *out)
{
    if (!m || !out) return LST_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    for (lst_ctx *c : m->ctx) {
        lst_stats s;
        int rc = lst_get_stats(c, &s);
        if (rc != LST_OK) return rc;
        out->frames += s.frames;
        out->tasks += s.tasks;
        out->respeculated += s.respeculated;
        out->passes += s.passes;
        out->kernel_launches += s.kernel_launches;
        out->h2d_bytes += s.h2d_bytes;
        out->d2h_bytes += s.d2h_bytes;
        out->ncc_kernel_ms += s.ncc_kernel_ms;
        out->ncc_kernel_launches += s.ncc_kernel_launches;
        out->ncc_algorithmic_macs += s.ncc_algorithmic_macs;
        out->roi_fallback_tasks += s.roi_fallback_tasks;
        out->device_wait_ms += s.device_wait_ms;
        out->host_ms += s.host_ms;
        out->preprocess_ms += s.preprocess_ms;
        out->box_sums_ms += s.box_sums_ms;
        out->epilogue_ms += s.epilogue_ms;
        out->preprocess_px += s.preprocess_px;
        out->preprocess_launches += s.preprocess_launches;
    }
    return LST_OK;
}

int lst_multi_reset_stats(lst_multi *m)
{
    if (!m) return LST_ERR_INVALID;
    for (lst_ctx *c : m->ctx) lst_reset_stats(c);
    return LST_OK;
}

int lst_set_frame_ranges(lst_ctx *c, const double *lo_hi, int32_t n)
{
    if (!c || n < 0 || (n > 0 && !lo_hi)) return LST_ERR_INVALID;
    c->user_ranges.assign(lo_hi, lo_hi + 2 * (size_t)n);
    return LST_OK;
}

// ---- asynchronous front end -------------------------------------------------------------------------
void lst_ctx::async_loop()
{
    for (;;) {
        AsyncJob *job = nullptr;
        {
            std::unique_lock<std::mutex> lk(amx);
            acv_work.wait(lk, [&] {
                if (astop) return true;
                for (auto &j : aq)
                    if (j->state == 0) return true;
                return false;
            });
            if (astop) return;
            for (auto &j : aq)
                if (j->state == 0) {
                    job = j.get();
                    break;
                }
            job->state = 1;
        }
        // batches run strictly one after the other: the recurrence (LST4:1327-1328) carries the state across them
        const int rc = lst_track_ptrs(this, job->ptrs.data(), job->first, (int32_t)job->ptrs.size(), job->out.data());
        {
            std::lock_guard<std::mutex> lk(amx);
            job->rc = rc;
            if (rc != LST_OK) job->err = err;
            job->state = 2;
        }
        acv_done.notify_all();
    }
}

int lst_set_queue_depth(lst_ctx *c, int32_t queue_depth)
{
    if (!c || queue_depth < 1 || queue_depth > 64) return LST_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->amx);
    c->queue_depth = queue_depth;
    return LST_OK;
}

int lst_submit(lst_ctx *c, const void *const *frame_ptrs, int64_t first_frame_index, int32_t n_frames)
{
    if (!c || !frame_ptrs || n_frames <= 0) return LST_ERR_INVALID;
    if (!c->have_ref) return c->fail(LST_ERR_STATE, "lst_submit before lst_set_reference");
    std::unique_lock<std::mutex> lk(c->amx);
    int pending = 0;
    for (auto &j : c->aq)
        if (j->state != 2) pending++;
    // finished batches nobody fetched count as well: they hold their records until lst_fetch takes them
    if (pending >= c->queue_depth || (int)c->aq.size() >= c->queue_depth + 2) return LST_WOULD_BLOCK;
    std::unique_ptr<lst_ctx::AsyncJob> job(new lst_ctx::AsyncJob());
    job->ptrs.assign(frame_ptrs, frame_ptrs + n_frames);
    job->first = first_frame_index;
    job->out.resize((size_t)n_frames);
    c->aq.push_back(std::move(job));
    if (!c->aworker.joinable()) c->aworker = std::thread([c] { c->async_loop(); });
    lk.unlock();
    c->acv_work.notify_one();
    return LST_OK;
}

int lst_fetch(lst_ctx *c, lst_record *out, int32_t max_records, int32_t *n_out, int32_t wait)
{
    if (!c || !out || !n_out || max_records < 0) return LST_ERR_INVALID;
    *n_out = 0;
    std::unique_lock<std::mutex> lk(c->amx);
    for (;;) {
        while (!c->aq.empty() && c->aq.front()->state == 2 && *n_out < max_records) {
            lst_ctx::AsyncJob &j = *c->aq.front();
            if (j.rc != LST_OK) {      // the batch failed: what came before it has been delivered; report it once
                if (*n_out > 0) return LST_OK;
                const int rc = j.rc;
                c->err = j.err;
                c->aq.pop_front();
                return rc;
            }
            const size_t take = std::min(j.out.size() - j.fetched, (size_t)(max_records - *n_out));
            memcpy(out + *n_out, j.out.data() + j.fetched, take * sizeof(lst_record));
            j.fetched += take;
            *n_out += (int32_t)take;
            if (j.fetched == j.out.size()) c->aq.pop_front();
        }
        if (*n_out > 0 || max_records == 0) return LST_OK;
        if (c->aq.empty()) return LST_OK;                  // nothing in flight: zero records is the answer
        if (!wait) return LST_WOULD_BLOCK;
        c->acv_done.wait(lk);
    }
}

int lst_override(lst_ctx *c, int64_t frame_index, const double dis[10], int64_t *resume_from)
{
    if (!c || !dis) return LST_ERR_INVALID;
    {
        std::unique_lock<std::mutex> lk(c->amx);
        // batches that have not started never will; the one that runs is waited for
        for (auto it = c->aq.begin(); it != c->aq.end();) it = (*it)->state == 0 ? c->aq.erase(it) : it + 1;
        c->acv_done.wait(lk, [&] {
            for (auto &j : c->aq)
                if (j->state == 1) return false;
            return true;
        });
        // records after frame_index are void: the recurrence restarts from the caller's state
        for (auto it = c->aq.begin(); it != c->aq.end();) {
            lst_ctx::AsyncJob &j = **it;
            if (j.first > frame_index) {
                it = c->aq.erase(it);
                continue;
            }
            const size_t keep = (size_t)std::min<int64_t>((int64_t)j.out.size(), frame_index - j.first + 1);
            j.out.resize(keep);
            j.ptrs.resize(keep);
            if (j.fetched >= j.out.size()) it = c->aq.erase(it);
            else ++it;
        }
    }
    if (resume_from) *resume_from = frame_index + 1;
    return lst_set_state(c, dis);
}

int lst_get_state(const lst_ctx *c, double dis[10])
{
    if (!c || !dis) return LST_ERR_INVALID;
    memcpy(dis, c->dis, sizeof(double) * 10);
    return LST_OK;
}
int lst_set_state(lst_ctx *c, const double dis[10])
{
    if (!c || !dis) return LST_ERR_INVALID;
    memcpy(c->dis, dis, sizeof(double) * 10);
    return LST_OK;
}

int lst_match_single(lst_ctx *c, const uint8_t *src8, int32_t sw, int32_t sh, const uint8_t *tpl8, int32_t tw,
                     int32_t th, int32_t method, int32_t sub_pix, double out[3], float *score_map)
{
    if (!c || !src8 || !tpl8 || !out) return LST_ERR_INVALID;
    if (method < 0 || method > 5) return c->fail(LST_ERR_INVALID, "matching method must be 0..5");
    if (tw < 1 || th < 1 || tw > 256 || th > 256 || sw < tw || sh < th) return c->fail(LST_ERR_INVALID, "bad sizes");
    CUC(c, cudaSetDevice(c->device));
    int rc = c->upload_template(kMaxTpl - 1, tpl8, tw, th, method);
    if (rc != LST_OK) return rc;
    size_t pitch = (sw + 15) & ~15;
    rc = c->ensure_scratch(pitch * sh + 512, score_map ? (size_t)sw * sh + 64 : 0);
    if (rc != LST_OK) return rc;
    CUC(c, cudaMemsetAsync(c->d_win8, 0, pitch * sh, c->s_compute));
    CUC(c, cudaMemcpy2DAsync(c->d_win8, pitch, src8, sw, sw, sh, cudaMemcpyHostToDevice, c->s_compute));
    c->stats.h2d_bytes += (int64_t)sw * sh;
    lst_task t;
    memset(&t, 0, sizeof(t));
    t.ww = sw;
    t.wh = sh;
    lst_task_result r;
    int save_sub = c->p.sub_pixel;
    c->p.sub_pixel = sub_pix;
    rc = c->run_group(&t, 1, &r, false, true, kMaxTpl - 1, score_map);
    c->p.sub_pixel = save_sub;
    if (rc != LST_OK) return rc;
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.score;
    return LST_OK;
}

int lst_match_test(lst_ctx *c, const uint8_t *src8, int32_t sw, int32_t sh, int32_t method, double *out)
{
    double o[3];
    int rc = lst_match_single(c, src8, sw, sh, src8, sw, sh, method, 0, o, nullptr);
    if (rc == LST_OK && out) *out = o[2];
    return rc;
}

int lst_preprocess(lst_ctx *c, const void *frame, int32_t x0, int32_t y0, int32_t w, int32_t h, uint8_t *out8)
{
    if (!c || !frame || !out8) return LST_ERR_INVALID;
    CUC(c, cudaSetDevice(c->device));
    Rect r = clip_rect(x0, y0, w, h, c->p.width, c->p.height);
    if (r.w != w || r.h != h) return c->fail(LST_ERR_INVALID, "rectangle outside the slice");
    int rc = ensure_ring(c, c->frame_bytes());
    if (rc != LST_OK) return rc;
    rc = c->ensure_frames(1);
    if (rc != LST_OK) return rc;
    CUC(c, cudaMemcpyAsync(c->d_ring[1], frame, c->frame_bytes(), cudaMemcpyHostToDevice, c->s_compute));
    c->h_surf[0] = whole_slice(c->d_ring[1], c->p.width, c->p.height);
    CUC(c, cudaMemcpyAsync(c->d_surf, c->h_surf, sizeof(Surface), cudaMemcpyHostToDevice, c->s_compute));
    if (c->frame_range_mode()) {
        std::vector<double> ur;
        ur.swap(c->user_ranges);
        rc = c->prepare_frame_ranges(1, ur.size() >= 2 ? ur.data() : nullptr);
        if (rc != LST_OK) return rc;
    }
    lst_task t;
    memset(&t, 0, sizeof(t));
    t.wx = x0;
    t.wy = y0;
    t.ww = w;
    t.wh = h;
    rc = c->run_group(&t, 1, nullptr, true, false, -1, nullptr);
    if (rc != LST_OK) return rc;
    for (int pl = 0; pl < (c->rgb3() ? 3 : 1); pl++)   // 3-channel match: the planes R, G, B one after the other
        CUC(c, cudaMemcpy2D(out8 + (size_t)pl * w * h, w, c->d_win8 + pl * c->last_plane_bytes + c->h_tasks[0].win_off, c->h_tasks[0].pitch, w, h,
                            cudaMemcpyDeviceToHost));
    return LST_OK;
}

int lst_calc_displacement(const float ref_xy[10], const double dis[10], double mark_dist, const double spot0[2],
                          double out5[5])
{
    if (!ref_xy || !dis || !out5) return LST_ERR_INVALID;
    double s0[2] = {spot0 ? spot0[0] : 0.0, spot0 ? spot0[1] : 0.0};
    calc_displacement(ref_xy, dis, mark_dist, spot0 != nullptr, s0, out5);
    return LST_OK;
}

int lst_alloc_pinned(size_t bytes, void **out)
{
    if (!out) return LST_ERR_INVALID;
    return cudaMallocHost(out, bytes) == cudaSuccess ? LST_OK : LST_ERR_NOMEM;
}
int lst_free_pinned(void *p)
{
    g_pin_generation.fetch_add(1, std::memory_order_acq_rel);
    return cudaFreeHost(p) == cudaSuccess ? LST_OK : LST_ERR_CUDA;
}
int lst_host_register(void *p, size_t bytes)
{
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? LST_OK : LST_ERR_CUDA;
}
int lst_host_unregister(void *p)
{
    g_pin_generation.fetch_add(1, std::memory_order_acq_rel);
    return cudaHostUnregister(p) == cudaSuccess ? LST_OK : LST_ERR_CUDA;
}

int lst_device_alloc(lst_ctx *c, size_t bytes, void **out)
{
    if (!c || !out) return LST_ERR_INVALID;
    CUC(c, cudaSetDevice(c->device));
    CUC(c, cudaMalloc(out, bytes));
    return LST_OK;
}
int lst_device_free(lst_ctx *c, void *p)
{
    if (!c) return LST_ERR_INVALID;
    CUC(c, cudaSetDevice(c->device));
    CUC(c, cudaFree(p));
    return LST_OK;
}
int lst_device_upload(lst_ctx *c, void *dst_dev, const void *src_host, size_t bytes)
{
    if (!c) return LST_ERR_INVALID;
    CUC(c, cudaSetDevice(c->device));
    CUC(c, cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
    return LST_OK;
}

int lst_get_stats(const lst_ctx *c, lst_stats *out)
{
    if (!c || !out) return LST_ERR_INVALID;
    *out = c->stats;
    return LST_OK;
}
int lst_reset_stats(lst_ctx *c)
{
    if (!c) return LST_ERR_INVALID;
    memset(&c->stats, 0, sizeof(c->stats));
    return LST_OK;
}
int lst_set_kernel_timing(lst_ctx *c, int32_t on)
{
    if (!c) return LST_ERR_INVALID;
    c->timing = on != 0;
    return LST_OK;
}

int lst_measure_alu_peaks(int device, double ms_per_test, double out[2])
{
    if (!out) return LST_ERR_INVALID;
    if (lst_device_count() <= device || device < 0) return LST_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return LST_ERR_CUDA;
    return measure_alu_peaks(ms_per_test, out) == 0 ? LST_OK : LST_ERR_CUDA;
}

// ---- host-only test seams -------------------------------------------------------------------------
int lst_host_window_for(const lst_params *p, const int32_t rect[4], double dis_x, double dis_y, int32_t out[4])
{
    if (!p || !rect || !out) return LST_ERR_INVALID;
    Rect r{rect[0], rect[1], rect[2], rect[3]};
    Rect w = window_for(*p, r, jint(dis_x), jint(dis_y), p->s_area, false, nullptr);
    out[0] = w.x;
    out[1] = w.y;
    out[2] = w.w;
    out[3] = w.h;
    return LST_OK;
}

int lst_host_template_rect(const lst_params *p, float ref_x, float ref_y, int32_t out[4])
{
    if (!p || !out) return LST_ERR_INVALID;
    Rect r = template_rect(*p, ref_x, ref_y);
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.w;
    out[3] = r.h;
    return LST_OK;
}

namespace {
struct CallbackExecutor : public Executor {
    lst_compute_fn fn;
    void *user;
    ResolveInput *in;
    int compute(const lst_task *tasks, int n, lst_task_result *results) override { return fn(user, tasks, n, results); }
    int solve(const double *dis10, int n, double *out5) override
    {
        for (int i = 0; i < n; i++) {
            double s0[2] = {in->spot0[0], in->spot0[1]};
            calc_displacement(in->ref_xy, dis10 + 10 * i, in->p.mark_dist, 1, s0, out5 + 5 * i);
        }
        return LST_OK;
    }
};
}  // namespace

int lst_host_resolve(const lst_params *p, const int32_t rect[5][4], const float ref_xy[10], const double spot0[2],
                     double dis_io[10], int64_t first_frame_index, int32_t n_frames, lst_compute_fn compute, void *user,
                     lst_record *out, lst_stats *stats_io)
{
    if (!p || !rect || !ref_xy || !spot0 || !dis_io || !compute || !out) return LST_ERR_INVALID;
    ResolveInput in;
    in.p = *p;
    for (int k = 0; k < 5; k++) in.rect[k] = Rect{rect[k][0], rect[k][1], rect[k][2], rect[k][3]};
    memcpy(in.ref_xy, ref_xy, sizeof(in.ref_xy));
    in.spot0[0] = spot0[0];
    in.spot0[1] = spot0[1];
    CallbackExecutor ex;
    ex.fn = compute;
    ex.user = user;
    ex.in = &in;
    int B = p->max_batch > 0 ? p->max_batch : 512;
    for (int f0 = 0; f0 < n_frames; f0 += B) {
        int nb = std::min(B, n_frames - f0);
        // the callback sees batch-local frame indices; shift them to caller-global ones
        struct Shift {
            lst_compute_fn fn;
            void *user;
            int f0;
        } sh{compute, user, f0};
        struct ShiftExecutor : public Executor {
            Shift *s;
            CallbackExecutor *base;
            int compute(const lst_task *tasks, int n, lst_task_result *results) override
            {
                std::vector<lst_task> t(tasks, tasks + n);
                for (auto &x : t) x.frame += s->f0;
                return s->fn(s->user, t.data(), n, results);
            }
            int solve(const double *d, int n, double *o) override { return base->solve(d, n, o); }
        } sx;
        sx.s = &sh;
        sx.base = &ex;
        int rc = resolve_batch(in, dis_io, first_frame_index + f0, nb, sx, out + f0, stats_io);
        if (rc != LST_OK) return rc;
    }
    return LST_OK;
}

}  // extern "C"
