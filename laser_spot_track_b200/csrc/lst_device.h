// lst_device.h -- structures and launchers shared by the CUDA kernels (lst_kernels.cu) and the
// runtime (lst_runtime.cu).  Internal; the public boundary is include/lst_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "lst_math.h"
#include "lst_jpeg_math.h"

namespace lst {

// One (frame, template) search window resident in the per-pass device scratch.
struct DevTask {
    int32_t frame;        // index into the pass's surface table
    int32_t tpl;          // template slot (0..4; match_single uses slot 5)
    int32_t x0, y0;       // window origin in the frame (already clipped to the frame)
    int32_t w, h;         // window size
    int32_t pitch;        // row pitch of the 8-bit window in scratch (bytes, multiple of 16)
    int32_t rw, rh;       // result-map size: w - tw + 1, h - th + 1
    int32_t s_used;       // search area used by the accept test (searchWidth = 2*s_used)
    int32_t pad0;         // 1: the 8-bit window at win_off belongs to an earlier task of the pass (blur kernels skip this one)
    int32_t pad1;         // slice index inside the batch (LST_RANGE_FRAME: row of PreprocParams.frame_mm)
    uint64_t win_off;     // byte offset of the 8-bit window in the win8 scratch
    uint64_t sum_off;     // element offset of this task's rw*rh score map (lst_match_single only)
};

// Where a task's pixels live: a whole slice in HBM, a gathered ROI in HBM, or a pinned host slice
// mapped into the device address space (zero-copy fallback of the ROI ingest).
struct Surface {
    const void *ptr;
    int32_t pitch_px;     // row pitch in pixels
    int32_t org_x, org_y; // slice coordinates of the surface's first pixel
    int32_t rows;         // rows of the surface: [ptr, ptr + rows*pitch) is readable (the blur kernel's 128-bit loads stay inside)
};

// One region of interest gathered from every slice of a batch (same rectangle for all slices).
// Ring layout: for each of the five regions a dense array [slice][h][w].
struct RoiDesc {
    int32_t x, y, w, h;   // slice coordinates
    uint64_t off;         // byte offset of this region's array in the ring half
    uint64_t slice_bytes; // w*h*bytes_per_pixel: distance between consecutive slices
};

struct DevResult {
    double x, y, score;   // coord_res of doMatch_coord_res (LST4:2502)
    int32_t ix, iy;
    int32_t accepted;
    int32_t peak_check;   // 1 if the epilogue's recomputed peak score equals the argmax score
    float nb[9];          // 3x3 neighbourhood of the peak (diagnostic)
    float pad;
};

static const int kMaxTpl = 6;   // five tracked templates + one scratch slot for lst_match_single

// Per-context template table as the kernels see it.
struct DevTemplates {
    TplConst c[kMaxTpl];
    const uint32_t *packed[kMaxTpl];    // [th][kw] words, see build_packed_template
    const uint8_t *raw[kMaxTpl];        // [th][tw] bytes
    const uint8_t *toeplitz[kMaxTpl];   // tensor-core engine: [th][2][nd*32][16] bytes, see tc_build_table
};

struct PreprocParams {
    int pixel_type;        // LST_PIX_*
    int frame_w, frame_h;  // slice size in pixels
    int range_per_task;    // 0: fixed lo/hi below, 1: lo/hi = min/max of the task's crop, 2: lo/hi = frame_mm keys of slice DevTask.pad1
    float lo, hi;
    int quant_mode;        // LST_QUANT_*
    int tile_w;            // columns per CTA of the blur kernel (whole window when it fits)
    int minmax_in_kernel;  // the blur kernel finds the crop min/max itself (single column block)
    int swap_bytes;        // the slices hold big-endian samples (files read by lst_track_files)
    const uint32_t *frame_mm;   // LST_RANGE_FRAME: {min key, max key} per slice of the batch (px_key order)
    int band_rows;         // lst_preprocess: output rows per CTA (0 = the whole window height); tall windows (whole-slice search) are split
    int rgb_shift;         // RGB slices: -1 = grey value of the pixel (matchIntensity), 16 / 8 / 0 = that channel alone (3-channel match)
};

struct MatchParams {
    int chunk_words;       // result columns per CTA, in 4-column words (a CTA owns 32 rows x this)
    int smem_pitch;        // window tile pitch in shared memory (odd multiple of 16 bytes)
    int v_pitch;           // pitch of the vertical-sum arrays (odd)
    int max_th;            // tallest template of the pass (sizes the window tile)
    int sub_pixel;
    int templ_size;
    double thresh;
    unsigned long long plane_bytes;   // 3-channel match: distance between the channel planes of the 8-bit windows (0 = one channel)
};

// launchers (all asynchronous on `stream`); each returns the number of kernels it launched or <0
int launch_init_pass(int n_tasks, uint32_t *mm_keys, unsigned long long *best, float *nbhd, cudaStream_t stream);
int launch_minmax(const DevTask *tasks, int n_tasks, int max_h, const Surface *surfaces, const PreprocParams &pp,
                  uint32_t *mm_keys, cudaStream_t stream);
int launch_preprocess(const DevTask *tasks, int n_tasks, int max_w, int max_h, const Surface *surfaces,
                      const PreprocParams &pp, const uint32_t *mm_keys, uint8_t *win8, cudaStream_t stream);
// Second-generation blur kernel (lst_preprocess_win.cuh): one CTA keeps the whole raw crop of a window in shared memory.
bool preprocess_win_eligible(int pixel_type, int max_w, int max_h);
int launch_preprocess_win(const DevTask *tasks, int n_tasks, int max_w, int max_h, const Surface *surfaces, const PreprocParams &pp,
                          uint8_t *win8, cudaStream_t stream);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda); rank <= 3, no swizzle, zero fill
bool encode_tensor_map(CUtensorMap *map, CUtensorMapDataType type, int rank, const void *base, const uint64_t *dims,
                       const uint64_t *strides_bytes, const uint32_t *box);
// ROI ingest by kernel (irregularly placed slices): copy 5 rectangles of each of n_frames pinned host
// slices (device-mapped pointers) into the ROI ring.
int launch_gather_rois(const void *const *mapped_frames, int n_frames, const RoiDesc rois[5], int frame_w, int bpp,
                       uint8_t *dst, cudaStream_t stream);
int launch_ncc_match(const DevTask *tasks, int n_tasks, int max_rw, int max_rh, int tw, int th,
                     const DevTemplates *tpls, const uint8_t *win8, unsigned long long *best, float *score_map,
                     const MatchParams &mp, cudaStream_t stream);
// 3-channel match (RGB stacks with matchIntensity = false): exact correlation summed over the channel planes, per-channel
// window sums, common_matchTemplate's cn = 3 normalisation, first-maximum argmax.  The planes of a window lie plane_bytes apart.
int launch_ncc_rgb(const DevTask *tasks, int n_tasks, int max_rh, int max_w, int max_tw, int max_th, const DevTemplates *tpls,
                   const uint8_t *win8, size_t plane_bytes, unsigned long long *best, cudaStream_t stream);
// JPEG slice files: the reconstruction stage of the decoder on the device (lst_jpeg.h).  For every (region, slice) of a batch
// the quantised coefficients the reader threads staged become pixels of the ROI ring: dequantisation + integer inverse DCT into
// shared-memory sample planes, then fancy chroma upsampling + YCbCr -> RGB (or the luma samples of a grey file).
struct JpegDevRegion {
    int x, y, w, h;                          // output rectangle in slice pixels (w == 0: nothing)
    JpegRegion m;                            // the MCUs staged for it
    unsigned long long coef_off, coef_slice; // int16 elements: the region's blocks of slice 0, distance between slices
    unsigned long long out_off, out_slice;   // bytes: the region in the ROI ring, distance between slices
};
struct JpegDevBatch {
    JpegGeom g;
    JpegDevRegion r[5];
};
size_t jpeg_reconstruct_smem(const JpegGeom &g, const JpegRegion &m);
int launch_jpeg_reconstruct(const int16_t *coef, const uint16_t *qt, const JpegDevBatch &batch, int n_slices, uint8_t *roi_out,
                            cudaStream_t stream);
int launch_epilogue(const DevTask *tasks, int n_tasks, const DevTemplates *tpls, const uint8_t *win8,
                    const unsigned long long *best, const float *nbhd, DevResult *results, const MatchParams &mp,
                    cudaStream_t stream);
int launch_solve(const double *dis10, int n, const float *ref_xy, double mark_dist, const double *spot0,
                 double *out5, cudaStream_t stream);
// Small host->device uploads on the compute stream go through a kernel reading pinned (mapped) host
// memory instead of the H2D copy engine: the engine's queue is busy for milliseconds with the bulk
// ingest of the *next* batch, and a 300 KB task list must not wait behind it.
int launch_upload(void *dst_dev, const void *src_pinned_host, size_t bytes, cudaStream_t stream);
int set_blur_kernel(const float kern[7], const float ksum[7]);
// RGB -> grey weighting factors; like ColorProcessor.setWeightingFactors they are global (per device)
int set_rgb_weights(const double w[3]);
// picks band/tile split and the shared-memory pitch for lst_ncc_match
void plan_match(int max_rw, int max_rh, int tw, int th, int n_tasks, MatchParams *mp, size_t *smem_bytes);
int measure_alu_peaks(double ms_per_test, double out[3]);

// Tensor-core engine of the match (lst_ncc_tc.cu): Toeplitz table of a template and the launcher of the
// persistent kernel (window sums, correlation, normalisation and argmax fused).
// table[v][kc][n][j] = T[v][32*d + 16*kc + j - nn], row n = (nd-1-d)*32 + nn (0 outside the template), kc < 2, j < 16.
int tc_nd(int tw);                                      // nd = ceil((31 + tw) / 32) band blocks per template row
size_t tc_table_bytes(int tw, int th);
void tc_build_table(const uint8_t *tpl8, int tw, int th, uint8_t *out);
bool tc_supported(int tw, int th);
// every window of the pass has row pitch `pitch` and starts at a multiple of it in win8 (total_rows rows in all)
int launch_ncc_tc(const DevTask *tasks, int n_tasks, int max_rw, int max_rh, int tw, int th, int pitch, long long total_rows,
                  const DevTemplates *tpls, const uint8_t *win8, unsigned long long *best, float *nbhd, cudaStream_t stream);
double measure_imma_peak(double ms_per_test);   // int8 tensor-core MAC/s (tcgen05.mma kind::i8)
void tc_init_device();

#if defined(__CUDACC__)
// order-preserving float <-> uint32 keys (argmax by atomicMax on 64-bit keys)
__device__ __forceinline__ uint32_t f32_key(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t k)
{
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
// 64-bit argmax key: larger key = better score, ties -> smaller index (minMaxLoc: first in row-major
// order).  want_min flips the score order for the methods that look for the minimum.
__device__ __forceinline__ unsigned long long pack_key(float score, uint32_t idx, bool want_min = false)
{
    score = score + 0.0f;   // -0 -> +0: minMaxLoc compares values, not bit patterns
    uint32_t k = f32_key(score);
    if (want_min) k = ~k;
    return ((unsigned long long)k << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ float key_score(unsigned long long key, bool want_min = false)
{
    uint32_t k = (uint32_t)(key >> 32);
    return key_f32(want_min ? ~k : k);
}
#endif

}  // namespace lst
