// lst_kernels.cu -- hand-written sm_100a kernels of the tracking hot path.
//
//   lst_minmax       per-window min/max of the raw crop (FloatProcessor.findMinAndMax of the crop)
//   lst_preprocess   crop -> float -> ImageJ GaussianBlur(2,2,0.02) -> 8-bit   (LST4:1346-1347, 1456-1473, 2529)
//   lst_ncc_match    exact integer cross-correlation (IDP.4A) + window sums of I, I^2 (cv::integral in
//                    matchTemplate) + TM_CCOEFF_NORMED normalisation + first-maximum argmax
//                                                                              (LST4:2560 matchTemplate, 2670 minMaxLoc)
//   lst_epilogue     3x3 sub-pixel fit + accept test per window                (LST4:2681-2717, 1225-1255)
//   lst_solve        calcDisplacement per frame                                (LST4:2318-2365)
//
// Compiled with -fmad=false: every float/double operation rounds once, like the Java and OpenCV
// code it reproduces.  All heavy arithmetic is integer (IDP.4A.U8.U8), which is exact.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>

#include "../../include/lst_b200.h"
#include "lst_device.h"

namespace lst {

#define KR 7   // kRadius of the sigma=2 / accuracy=0.02 Gaussian (13 taps)

__constant__ float c_kern[KR];
__constant__ float c_ksum[KR];

static void prefer_max_shared_everywhere();

int set_blur_kernel(const float kern[7], const float ksum[7])
{
    prefer_max_shared_everywhere();
    if (cudaMemcpyToSymbol(c_kern, kern, sizeof(float) * KR) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_ksum, ksum, sizeof(float) * KR) != cudaSuccess) return -1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------

template <int PIX>
__device__ __forceinline__ float load_px(const void *base, size_t idx)
{
    if (PIX == LST_PIX_U8) return (float)((const uint8_t *)base)[idx];
    if (PIX == LST_PIX_U16) return (float)((const uint16_t *)base)[idx];
    return ((const float *)base)[idx];
}
template <int PIX>
__device__ __forceinline__ uint32_t px_key(const void *base, size_t idx)
{
    if (PIX == LST_PIX_U8) return ((const uint8_t *)base)[idx];
    if (PIX == LST_PIX_U16) return ((const uint16_t *)base)[idx];
    return f32_key(((const float *)base)[idx]);
}
template <int PIX>
__device__ __forceinline__ float key_to_float(uint32_t k)
{
    if (PIX == LST_PIX_F32) return key_f32(k);
    return (float)k;
}

// ------------------------------------------------------------------------------------------------
// small uploads by kernel (see lst_device.h)
// ------------------------------------------------------------------------------------------------
__global__ void lst_upload(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, size_t nwords)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

int launch_upload(void *dst_dev, const void *src_pinned_host, size_t bytes, cudaStream_t stream)
{
    size_t nwords = (bytes + 3) / 4;   // buffers are allocated in multiples of 4 bytes
    if (nwords == 0) return 0;
    const void *mapped = nullptr;
    if (cudaHostGetDevicePointer((void **)&mapped, (void *)src_pinned_host, 0) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    int blocks = (int)((nwords + 255) / 256);
    if (blocks > 296) blocks = 296;
    lst_upload<<<blocks, 256, 0, stream>>>((uint32_t *)dst_dev, (const uint32_t *)mapped, nwords);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// pass init
// ------------------------------------------------------------------------------------------------
__global__ void lst_init_pass(int n_tasks, uint32_t *mm_keys, unsigned long long *best)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tasks) {
        mm_keys[2 * i] = 0xffffffffu;
        mm_keys[2 * i + 1] = 0u;
        best[i] = 0ull;
    }
}

int launch_init_pass(int n_tasks, uint32_t *mm_keys, unsigned long long *best, cudaStream_t stream)
{
    lst_init_pass<<<(n_tasks + 255) / 256, 256, 0, stream>>>(n_tasks, mm_keys, best);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// min / max of the raw crop
// ------------------------------------------------------------------------------------------------
template <int PIX>
__global__ void lst_minmax(const DevTask *__restrict__ tasks, const Surface *__restrict__ surfaces,
                           uint32_t *__restrict__ mm_keys)
{
    const DevTask t = tasks[blockIdx.y];
    const Surface sf = surfaces[t.frame];
    const void *fr = sf.ptr;
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int r = blockIdx.x; r < t.h; r += gridDim.x) {
        size_t row = (size_t)(t.y0 - sf.org_y + r) * sf.pitch_px + (t.x0 - sf.org_x);
        for (int c = threadIdx.x; c < t.w; c += blockDim.x) {
            uint32_t k = px_key<PIX>(fr, row + c);
            lo = min(lo, k);
            hi = max(hi, k);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm_keys[2 * blockIdx.y], lo);
        atomicMax(&mm_keys[2 * blockIdx.y + 1], hi);
    }
}

int launch_minmax(const DevTask *tasks, int n_tasks, int max_h, const Surface *surfaces, const PreprocParams &pp,
                  uint32_t *mm_keys, cudaStream_t stream)
{
    int launches = 0;
    int gx = max_h < 8 ? 1 : (max_h + 31) / 32;   // 32 rows per CTA
    if (gx > 256) gx = 256;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 grid(gx, n);
        if (pp.pixel_type == LST_PIX_U8)
            lst_minmax<LST_PIX_U8><<<grid, 128, 0, stream>>>(tasks + base, surfaces, mm_keys + 2 * base);
        else if (pp.pixel_type == LST_PIX_U16)
            lst_minmax<LST_PIX_U16><<<grid, 128, 0, stream>>>(tasks + base, surfaces, mm_keys + 2 * base);
        else
            lst_minmax<LST_PIX_F32><<<grid, 128, 0, stream>>>(tasks + base, surfaces, mm_keys + 2 * base);
        launches++;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------------
// ROI ingest: gather the five search regions of every slice of a batch from pinned host memory
// ------------------------------------------------------------------------------------------------
// The slices stay in the caller's pinned memory; only the rectangles the windows can fall in
// cross PCIe (256 KB instead of 4.1 MB per 1080p slice).  blockIdx.y = slice*5 + roi,
// blockIdx.x = group of rows; a row is copied with 16-byte loads when the addresses allow.
struct GatherArgs {
    RoiDesc roi[5];
    int frame_w, bpp;
};

#define GATHER_ROWS 8

// Persistent: a small fixed grid (2 CTAs per SM) walks all (slice, roi, row group) items, so the
// gather keeps enough PCIe reads in flight without taking the SMs away from the match kernels of
// the previous batch that run concurrently on the compute stream.
__global__ void __launch_bounds__(256)
lst_gather_rois(const void *const *__restrict__ frames, int n_frames, GatherArgs a, int groups, uint8_t *__restrict__ dst)
{
    const long total = (long)n_frames * 5 * groups;
    for (long item = blockIdx.x; item < total; item += gridDim.x) {
        const int g = (int)(item % groups);
        const int fk = (int)(item / groups);
        const int f = fk / 5, k = fk - f * 5;
        const RoiDesc r = a.roi[k];
        const int row0 = g * GATHER_ROWS;
        if (row0 >= r.h) continue;
        const int nrows = min(GATHER_ROWS, r.h - row0);
        const uint8_t *src = (const uint8_t *)frames[f];
        uint8_t *out = dst + r.off + (size_t)f * r.slice_bytes;
        const size_t src_pitch = (size_t)a.frame_w * a.bpp, row_bytes = (size_t)r.w * a.bpp;
        const size_t src0 = ((size_t)r.y * a.frame_w + r.x) * a.bpp;
        const bool vec = (((uintptr_t)src + src0) % 16 == 0) && (src_pitch % 16 == 0) && (row_bytes % 16 == 0) &&
                         (((uintptr_t)out) % 16 == 0);
        if (vec) {
            const int per_row = (int)(row_bytes / 16);
            for (int i = threadIdx.x; i < per_row * nrows; i += blockDim.x) {
                int rr = i / per_row, cc = i - rr * per_row;
                const uint4 v = *(const uint4 *)(src + src0 + (size_t)(row0 + rr) * src_pitch + (size_t)cc * 16);
                *(uint4 *)(out + (size_t)(row0 + rr) * row_bytes + (size_t)cc * 16) = v;
            }
        } else {
            for (int i = threadIdx.x; i < (int)row_bytes * nrows; i += blockDim.x) {
                int rr = i / (int)row_bytes, cc = i - rr * (int)row_bytes;
                out[(size_t)(row0 + rr) * row_bytes + cc] = src[src0 + (size_t)(row0 + rr) * src_pitch + cc];
            }
        }
    }
}

int launch_gather_rois(const void *const *mapped_frames, int n_frames, const RoiDesc rois[5], int frame_w, int bpp,
                       uint8_t *dst, cudaStream_t stream)
{
    GatherArgs a;
    int max_h = 0;
    for (int k = 0; k < 5; k++) {
        a.roi[k] = rois[k];
        max_h = rois[k].h > max_h ? rois[k].h : max_h;
    }
    a.frame_w = frame_w;
    a.bpp = bpp;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int groups = (max_h + GATHER_ROWS - 1) / GATHER_ROWS;
    long total = (long)n_frames * 5 * groups;
    int grid = (int)(total < 2L * sms ? total : 2L * sms);
    lst_gather_rois<<<grid, 256, 0, stream>>>(mapped_frames, n_frames, a, groups, dst);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// crop -> float -> blur -> 8-bit
// ------------------------------------------------------------------------------------------------
// One line of ImageJ's GaussianBlur.convolveLine for output index i of a line of length n
// (n >= 2*KR), float32, Java operation order.  IN(i) reads the line.
#define LST_BLUR_AT(res, i, n, IN)                                            \
    {                                                                         \
        res = IN(i) * c_kern[0];                                              \
        if ((i) < KR) res = res + c_ksum[(i)] * IN(0);                        \
        if ((i) + KR >= (n)) res = res + c_ksum[(n) - (i)-1] * IN((n)-1);     \
        _Pragma("unroll") for (int k_ = 1; k_ < KR; k_++)                     \
        {                                                                     \
            float v_ = 0.0f;                                                  \
            if ((i)-k_ >= 0) v_ = v_ + IN((i)-k_);                            \
            if ((i) + k_ < (n)) v_ = v_ + IN((i) + k_);                       \
            res = res + c_kern[k_] * v_;                                      \
        }                                                                     \
    }

// A CTA owns a column block (up to tile_w columns, the whole window for the named configurations) of
// one window over its full height and walks it in strips of PRE_STRIP rows.  X-pass rows live in a
// ring of PRE_RING rows indexed by (row mod PRE_RING), so the 12 rows two consecutive strips share
// are computed once.  When the range convention needs the crop's min/max and the window is a
// single column block, the CTA finds them itself first (no separate kernel).
#define PRE_STRIP 32
#define PRE_RING (PRE_STRIP + 2 * (KR - 1))   // 44

template <int PIX>
__global__ void __launch_bounds__(256)
lst_preprocess(const DevTask *__restrict__ tasks, const Surface *__restrict__ surfaces, PreprocParams pp,
               const uint32_t *__restrict__ mm_keys, uint8_t *__restrict__ win8)
{
    extern __shared__ float smem_f[];
    __shared__ uint32_t s_mm[2];
    const DevTask t = tasks[blockIdx.y];
    const int TW = pp.tile_w;
    const int tiles_x = (t.w + TW - 1) / TW;
    if ((int)blockIdx.x >= tiles_x) return;
    const int cx0 = blockIdx.x * TW;
    const int cx1 = min(cx0 + TW, t.w);
    const int rx0 = max(cx0 - (KR - 1), 0), rx1 = min(cx1 + (KR - 1), t.w);
    const int RWP = TW + 2 * (KR - 1) + 1;     // raw strip pitch (floats)
    const int XWP = TW + 1;                    // x-pass ring pitch
    float *raw = smem_f;                                            // [PRE_STRIP + KR - 1][RWP]
    float *xp = smem_f + (size_t)(PRE_STRIP + KR - 1) * RWP;        // [PRE_RING][XWP]
    const Surface sf = surfaces[t.frame];
    const void *fr = sf.ptr;
    const int nrw = rx1 - rx0, ncw = cx1 - cx0;
    const size_t src00 = (size_t)(t.y0 - sf.org_y) * sf.pitch_px + (t.x0 - sf.org_x);

    float lo = pp.lo, hi = pp.hi;
    if (pp.range_per_task) {
        if (pp.minmax_in_kernel) {
            if (threadIdx.x == 0) {
                s_mm[0] = 0xffffffffu;
                s_mm[1] = 0u;
            }
            __syncthreads();
            uint32_t klo = 0xffffffffu, khi = 0u;
            for (int r = threadIdx.x >> 5; r < t.h; r += 8) {
                const size_t rowoff = src00 + (size_t)r * sf.pitch_px;
#pragma unroll 8
                for (int c = threadIdx.x & 31; c < t.w; c += 32) {
                    uint32_t k = px_key<PIX>(fr, rowoff + c);
                    klo = min(klo, k);
                    khi = max(khi, k);
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                klo = min(klo, __shfl_xor_sync(0xffffffffu, klo, o));
                khi = max(khi, __shfl_xor_sync(0xffffffffu, khi, o));
            }
            if ((threadIdx.x & 31) == 0) {
                atomicMin(&s_mm[0], klo);
                atomicMax(&s_mm[1], khi);
            }
            __syncthreads();
            lo = key_to_float<PIX>(s_mm[0]);
            hi = key_to_float<PIX>(s_mm[1]);
        } else {
            lo = key_to_float<PIX>(mm_keys[2 * blockIdx.y]);
            hi = key_to_float<PIX>(mm_keys[2 * blockIdx.y + 1]);
        }
    }
    const float scale = (pp.quant_mode == LST_QUANT_ROUND255 ? 255.0f : 256.0f) / (hi - lo);
    uint8_t *dst = win8 + t.win_off;
    const int nsegx = (ncw + 3) >> 2;

    int xdone = 0;   // x-pass rows [0, xdone) are in the ring (the newest PRE_RING of them)
    for (int cy0 = 0; cy0 < t.h; cy0 += PRE_STRIP) {
        const int cy1 = min(cy0 + PRE_STRIP, t.h);
        const int need = min(cy1 + (KR - 1), t.h);          // x-pass rows needed up to here
        const int newr = need - xdone;                      // <= PRE_STRIP + KR - 1
        // raw rows [xdone, need), columns [rx0, rx1)
        for (int r = threadIdx.x >> 5; r < newr; r += 8) {
            const size_t rowoff = src00 + (size_t)(xdone + r) * sf.pitch_px + rx0;
#pragma unroll 8
            for (int c = threadIdx.x & 31; c < nrw; c += 32) raw[r * RWP + c] = load_px<PIX>(fr, rowoff + c);
        }
        __syncthreads();
        // X pass of the new rows.  Interior runs of 4 columns (no edge replication) take 16 loads per
        // 4 results; the few runs next to the crop's left/right border take the general form.  The two
        // kinds are walked by separate loops so that no warp executes both paths.
        {
            int sg_lo = cx0 >= KR ? 0 : (KR - cx0 + 3) >> 2;
            int lim = min(t.w - (KR + 4), cx1 - 4) - cx0;              // i0 - cx0 <= lim
            int sg_hi = lim < 0 ? -1 : lim >> 2;                        // last interior run
            if (sg_hi >= nsegx) sg_hi = nsegx - 1;
            const int nint = sg_hi >= sg_lo ? sg_hi - sg_lo + 1 : 0;
            for (int idx = threadIdx.x; idx < nint * newr; idx += blockDim.x) {
                const int r = idx / nint, sg = sg_lo + (idx - r * nint);
                const float *line = raw + r * RWP - rx0;   // line[i] = crop row pixel i
                const int i0 = cx0 + 4 * sg;
                float *dstx = xp + ((xdone + r) % PRE_RING) * XWP + 4 * sg;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = line[i0 - (KR - 1) + j];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float res = v[KR - 1 + j] * c_kern[0];
#pragma unroll
                    for (int k = 1; k < KR; k++) res = res + c_kern[k] * (v[KR - 1 + j - k] + v[KR - 1 + j + k]);
                    dstx[j] = res;
                }
            }
            const int nedge = nsegx - nint;                              // runs [0, sg_lo) and (sg_hi, nsegx)
            for (int idx = threadIdx.x; idx < nedge * 4 * newr; idx += blockDim.x) {
                const int r = idx / (nedge * 4), e = idx - r * (nedge * 4);
                const int es = e >> 2, j = e & 3;
                const int sg = nint == 0 ? es : (es < sg_lo ? es : sg_hi + 1 + (es - sg_lo));
                const int i = cx0 + 4 * sg + j;
                if (i < cx1) {
                    const float *line = raw + r * RWP - rx0;
                    float res;
#define IN_X(ii) line[(ii)]
                    LST_BLUR_AT(res, i, t.w, IN_X)
#undef IN_X
                    xp[((xdone + r) % PRE_RING) * XWP + 4 * sg + j] = res;
                }
            }
        }
        xdone = need;
        __syncthreads();
        // Y pass + quantise of rows [cy0, cy1): a thread produces 4 consecutive rows of one column
        // (consecutive threads own consecutive columns: conflict-free, coalesced stores).
        const int nch = cy1 - cy0;
        const int nsegy = (nch + 3) >> 2;
        for (int idx = threadIdx.x; idx < ncw * nsegy; idx += blockDim.x) {
            const int sg = idx / ncw, cc = idx - sg * ncw;
            const float *col = xp + cc;                // x-pass value at crop row i: col[(i % PRE_RING) * XWP]
            const int i0 = cy0 + 4 * sg;
            float out[4];
            const int nout = min(4, cy1 - i0);
            if (i0 >= KR && i0 + 3 + KR < t.h && nout == 4) {
                float v[16];
                int slot = (i0 - (KR - 1)) % PRE_RING;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    v[j] = col[slot * XWP];
                    slot = slot + 1 == PRE_RING ? 0 : slot + 1;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float res = v[KR - 1 + j] * c_kern[0];
#pragma unroll
                    for (int k = 1; k < KR; k++) res = res + c_kern[k] * (v[KR - 1 + j - k] + v[KR - 1 + j + k]);
                    out[j] = res;
                }
            } else {
                for (int j = 0; j < nout; j++) {
                    const int i = i0 + j;
                    float res;
#define IN_Y(ii) col[((ii) % PRE_RING) * XWP]
                    LST_BLUR_AT(res, i, t.h, IN_Y)
#undef IN_Y
                    out[j] = res;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j < nout) {
                    float value = out[j] - lo;
                    if (value < 0.0f) value = 0.0f;
                    float q = value * scale;
                    if (pp.quant_mode == LST_QUANT_ROUND255) q = q + 0.5f;
                    int iv = __float2int_rz(q);            // Java (int): NaN -> 0, saturating
                    iv = iv > 255 ? 255 : (iv < 0 ? 0 : iv);
                    dst[(size_t)(i0 + j) * t.pitch + (cx0 + cc)] = (uint8_t)iv;
                }
            }
        }
        // zero the pitch padding so the match kernel can load whole 16-byte words
        if (cx1 == t.w && t.pitch > t.w) {
            int padw = t.pitch - t.w;
            for (int idx = threadIdx.x; idx < padw * nch; idx += blockDim.x) {
                int r = idx / padw, c = idx - r * padw;
                dst[(size_t)(cy0 + r) * t.pitch + t.w + c] = 0;
            }
        }
        __syncthreads();   // the next strip overwrites raw and the oldest ring rows
    }
}

static size_t preprocess_smem(const PreprocParams &pp)
{
    return ((size_t)(PRE_STRIP + KR - 1) * (pp.tile_w + 2 * (KR - 1) + 1) + (size_t)PRE_RING * (pp.tile_w + 1)) * sizeof(float);
}

int launch_preprocess(const DevTask *tasks, int n_tasks, int max_w, int max_h, const Surface *surfaces,
                      const PreprocParams &pp, const uint32_t *mm_keys, uint8_t *win8, cudaStream_t stream)
{
    int tiles = (max_w + pp.tile_w - 1) / pp.tile_w;
    size_t smem = preprocess_smem(pp);
    (void)max_h;
    int launches = 0;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 grid(tiles, n);
        const uint32_t *mm = mm_keys + 2 * base;
        if (pp.pixel_type == LST_PIX_U8)
            lst_preprocess<LST_PIX_U8><<<grid, 256, smem, stream>>>(tasks + base, surfaces, pp, mm, win8);
        else if (pp.pixel_type == LST_PIX_U16)
            lst_preprocess<LST_PIX_U16><<<grid, 256, smem, stream>>>(tasks + base, surfaces, pp, mm, win8);
        else
            lst_preprocess<LST_PIX_F32><<<grid, 256, smem, stream>>>(tasks + base, surfaces, pp, mm, win8);
        launches++;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------------
// exact NCC + window sums + argmax (one kernel)
// ------------------------------------------------------------------------------------------------
// Result position x = 4q + j.  The correlation of one template row is
//     sum_k IDP.4A(Wj[q + k], T[k]),   Wj[m] = window bytes 4m+j .. 4m+j+3,
// where T[k] are the packed template words (build_packed_template) and Wj is made from two
// adjacent aligned window words with one funnel shift (SHF, ALU pipe -- it runs beside the
// IDP.4A pipe).  Every load is an aligned word; one shifted word serves 4 template words.
//
// A CTA owns a band of 32 result rows x a chunk of result columns of one window.  It stages the
// window rows it needs in shared memory (pitch = odd multiple of 16 B, so 128-bit row loads of 32
// lanes that own 32 consecutive rows are conflict-free), builds the vertical sliding sums of I and
// I^2 for its band (what cv::integral provides to matchTemplate), and then every thread takes
// (tile, row) tasks: a tile is Q = 8, 4 or 1 window words = 32, 16 or 4 result columns, so that
// the last columns of an odd-sized result map cost an eighth of a tile instead of a whole one.
// Template words are warp-uniform read-only loads served by L1.
//
// Normalisation: an exact integer form of the score, A / sqrt(B*C) with
//     A = N*corr - wsum*tsum,  B = N*wsq - wsum^2,  C = N*tsq - tsum^2,
// is evaluated in float32 as a filter (relative error < 1e-6); only positions within 1e-5 of the
// thread's running maximum -- the only ones that can win the argmax -- are evaluated with
// OpenCV's double-precision formula (ncc_score), whose float32 result is what is compared and
// reported.  Nearly flat windows (B <= ~N/2, where OpenCV's "avoid rounding errors" branch may
// apply) always take the exact path.

#define NCC_THREADS 128
#define NCC_BAND 32
#define NCC_EPS 1e-5f
#define NCC_ROWTASK_MAX 16

struct NccCtx {
    const uint8_t *s_win;      // window tile in shared memory
    const uint16_t *s_v;       // vertical sums of I   [NCC_BAND][vpitch]
    const uint32_t *s_v2;      // vertical sums of I^2 [NCC_BAND][vpitch]
    const uint4 *g_tpl;        // packed template [th][kw/4] uint4 (global, L1)
    int SP, VP, th, tw, kw, rw;
    const TplConst *tc;        // all constants of the template (methods other than 5 use them)
    long long N, tsum, C_unused;
    float inv_sqrt_c;
    double inv_area, mean, norm;
    size_t map_off;            // sum_off of the task (score map)
    float *score_map;
    int y;                     // absolute result row
    int x_chunk0;              // first result column of the CTA's chunk
};

// One (tile, row) task: Q window words -> 4Q result columns starting at word `w0` of the chunk.
template <int Q, bool MAP>
__device__ __forceinline__ void ncc_tile(const NccCtx &c, bool valid, int yy, int w0, unsigned long long &best)
{
    uint32_t acc[4][Q];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int q = 0; q < Q; q++) acc[j][q] = 0u;
    const uint8_t *wbase = c.s_win + (size_t)yy * c.SP + 4 * w0;
    constexpr int NL = Q + 4;              // window words [k, k + Q + 4) feed one chunk of 4 template words
    constexpr int NW = (NL + 3) / 4;       // as 128-bit loads
    // Per template row: kw/4 chunks of 4 packed template words (one LDG.128, warp-uniform, L1) against
    // NL window words (LDS.128).  Shift j uses Wj[i] = funnelshift(ww[i], ww[i+1], 8j); 16*Q IDP.4A per chunk.
    const int nchunk = c.kw >> 2;
    const uint4 *trow = c.g_tpl;
    const uint8_t *wrow = wbase;
    for (int v = 0; v < c.th; v++, wrow += c.SP) {
        const uint8_t *wp = wrow;
        for (int kc = 0; kc < nchunk; kc++, wp += 16, trow++) {
            uint32_t ww[4 * NW];
            if (Q == 1) {   // single-word tiles start at any word: 32-bit loads
#pragma unroll
                for (int i = 0; i < NL; i++) ww[i] = *(const uint32_t *)(wp + 4 * i);
            } else {        // 4- and 8-word tiles start at multiples of 4 words: 128-bit loads
#pragma unroll
                for (int i = 0; i < NW; i++) {
                    const uint4 t4 = *(const uint4 *)(wp + 16 * i);
                    ww[4 * i] = t4.x;
                    ww[4 * i + 1] = t4.y;
                    ww[4 * i + 2] = t4.z;
                    ww[4 * i + 3] = t4.w;
                }
            }
            const uint4 tq = __ldg(trow);
            const uint32_t tk[4] = {tq.x, tq.y, tq.z, tq.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t wj[Q + 3];
#pragma unroll
                for (int i = 0; i < Q + 3; i++) wj[i] = j == 0 ? ww[i] : __funnelshift_r(ww[i], ww[i + 1], 8 * j);
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
#pragma unroll
                    for (int q = 0; q < Q; q++) acc[j][q] = __dp4a(wj[kk + q], tk[kk], acc[j][q]);
            }
        }
    }
    // Normalisation.  The 4Q correlations move to a thread-local array so that the loops below are
    // real loops (a few dozen instructions instead of 4Q inlined copies: the instruction cache matters).
    // Pass 1: float32 filter score of every position, warp-wide maximum.
    // Pass 2: positions within NCC_EPS of that maximum (and nearly flat windows) are evaluated with
    // OpenCV's double-precision formula; only those can be the first maximum of the map.
    uint32_t corr[4 * Q];
#pragma unroll
    for (int p = 0; p < 4 * Q; p++) corr[p] = acc[p & 3][p >> 2];
    const int xr = 4 * w0;                       // column relative to the chunk
    const int x0 = c.x_chunk0 + xr;              // absolute result column
    const int np = valid ? min(4 * Q, c.rw - x0) : 0;
    const uint16_t *vr = c.s_v + (size_t)yy * c.VP + xr;
    const uint32_t *vr2 = c.s_v2 + (size_t)yy * c.VP + xr;
    uint32_t s0 = 0, sq0 = 0;
    if (np > 0) {
#pragma unroll 4
        for (int u = 0; u < c.tw; u++) {
            s0 += vr[u];
            sq0 += vr2[u];
        }
    }
    const uint32_t rowidx = (uint32_t)c.y * (uint32_t)c.rw;
    float wmax = -3.0e38f;
    if (!MAP && c.tc->method == 5) {
        uint32_t s = s0, sq = sq0;
        float tmax = -3.0e38f;
#pragma unroll 1
        for (int p = 0; p < np; p++) {
            const long long A = c.N * (long long)corr[p] - (long long)s * c.tsum;
            const long long B = c.N * (long long)sq - (long long)s * (long long)s;
            if (2 * B > c.N + (c.N >> 5)) tmax = fmaxf(tmax, (float)A * rsqrtf((float)B) * c.inv_sqrt_c);
            s += (uint32_t)vr[p + c.tw] - (uint32_t)vr[p];
            sq += vr2[p + c.tw] - vr2[p];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        wmax = tmax;
    }
    {
        uint32_t s = s0, sq = sq0;
#pragma unroll 1
        for (int p = 0; p < np; p++) {
            bool exact = MAP || c.tc->method != 5;   // the float32 filter is derived for TM_CCOEFF_NORMED only
            if (!exact) {
                const long long A = c.N * (long long)corr[p] - (long long)s * c.tsum;
                const long long B = c.N * (long long)sq - (long long)s * (long long)s;
                if (2 * B <= c.N + (c.N >> 5)) exact = true;   // nearly flat: OpenCV's special branch may apply
                else exact = (float)A * rsqrtf((float)B) * c.inv_sqrt_c >= wmax - NCC_EPS;
            }
            if (exact) {
                const float sc = match_score(*c.tc, (double)corr[p], (double)s, (double)sq);
                if (MAP) c.score_map[c.map_off + rowidx + (uint32_t)(x0 + p)] = sc;
                const unsigned long long key = pack_key(sc, rowidx + (uint32_t)(x0 + p), method_wants_min(c.tc->method));
                best = key > best ? key : best;
            }
            s += (uint32_t)vr[p + c.tw] - (uint32_t)vr[p];
            sq += vr2[p + c.tw] - vr2[p];
        }
    }
}

template <bool MAP>
__global__ void __launch_bounds__(NCC_THREADS, 4)
lst_ncc_match(const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls,
              const uint8_t *__restrict__ win8, unsigned long long *__restrict__ best,
              float *__restrict__ score_map, MatchParams mp)
{
    extern __shared__ __align__(16) uint8_t smem_b[];
    __shared__ unsigned long long s_best[NCC_THREADS / 32];
    const DevTask t = tasks[blockIdx.y];
    const TplConst tc = tpls->c[t.tpl];
    const int tw = tc.tw, th = tc.th;
    const int words_total = (t.rw + 3) / 4;                       // result columns in 4-column words
    const int chunks = (words_total + mp.chunk_words - 1) / mp.chunk_words;
    const int bands = (t.rh + NCC_BAND - 1) / NCC_BAND;
    if ((int)blockIdx.x >= chunks * bands) return;
    const int chunk = blockIdx.x % chunks, band = blockIdx.x / chunks;
    if (tc.flat && tc.method == 5) {   // templSdv^2 < DBL_EPSILON: OpenCV fills the result with 1 -> first position wins
        if (blockIdx.x == 0) {
            if (MAP)
                for (int i = threadIdx.x; i < t.rw * t.rh; i += blockDim.x) score_map[t.sum_off + i] = 1.0f;
            if (threadIdx.x == 0) atomicMax(&best[blockIdx.y], pack_key(1.0f, 0u));
        }
        return;
    }
    const int word0 = chunk * mp.chunk_words;
    const int nwords = min(mp.chunk_words, words_total - word0);
    const int y0 = band * NCC_BAND;
    const int nrows = min(NCC_BAND, t.rh - y0);
    const int SP = mp.smem_pitch, VP = mp.v_pitch;

    uint8_t *s_win = smem_b;                                              // [(NCC_BAND+th-1)][SP]
    uint32_t *s_v2 = (uint32_t *)(smem_b + (size_t)(NCC_BAND + mp.max_th - 1) * SP);   // [NCC_BAND][VP]
    uint16_t *s_v = (uint16_t *)(s_v2 + (size_t)NCC_BAND * VP);           // [NCC_BAND][VP]
    const int wrows = nrows + th - 1;
    {
        const int vec = SP / 16;
        const int xb0 = word0 * 4;                       // first window byte column of this chunk
        const uint8_t *src = win8 + t.win_off;
        for (int i = threadIdx.x; i < wrows * vec; i += blockDim.x) {
            int r = i / vec, cc = (i - r * vec) * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            int gx = xb0 + cc, gy = y0 + r;
            if (gx < t.pitch && gy < t.h) v = *(const uint4 *)(src + (size_t)gy * t.pitch + gx);
            *(uint4 *)(s_win + (size_t)r * SP + cc) = v;
        }
    }
    __syncthreads();
    {
        // vertical sliding sums for the columns this chunk's positions touch: [0, 4*nwords + tw - 1)
        const int ncols = min(4 * nwords + tw, VP);
        for (int cc = threadIdx.x; cc < ncols; cc += blockDim.x) {
            uint32_t s = 0, q = 0;
            for (int v = 0; v < th; v++) {
                uint32_t p = s_win[(size_t)v * SP + cc];
                s += p;
                q += p * p;
            }
            s_v[cc] = (uint16_t)s;
            s_v2[cc] = q;
            for (int y = 1; y < nrows; y++) {
                uint32_t a = s_win[(size_t)(y + th - 1) * SP + cc], b = s_win[(size_t)(y - 1) * SP + cc];
                s += a - b;
                q += a * a - b * b;
                s_v[(size_t)y * VP + cc] = (uint16_t)s;
                s_v2[(size_t)y * VP + cc] = q;
            }
        }
    }
    __syncthreads();

    NccCtx c;
    c.s_win = s_win;
    c.s_v = s_v;
    c.s_v2 = s_v2;
    c.g_tpl = (const uint4 *)tpls->packed[t.tpl];
    c.tc = &tpls->c[t.tpl];
    c.SP = SP;
    c.VP = VP;
    c.th = th;
    c.tw = tw;
    c.kw = tc.kw;
    c.rw = t.rw;
    c.N = (long long)tw * th;
    c.tsum = (long long)tc.tsum;
    c.inv_sqrt_c = tc.inv_sqrt_c;
    c.inv_area = tc.inv_area;
    c.mean = tc.mean;
    c.norm = tc.norm;
    c.map_off = t.sum_off;
    c.score_map = score_map;
    c.x_chunk0 = word0 * 4;

    // tiles of the chunk: n8 x (8 words), n4 x (4 words), n1 x (1 word)
    const int n8 = nwords >> 3, n4 = (nwords & 7) >> 2, n1 = nwords & 3;
    unsigned long long mybest = 0ull;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (nrows <= NCC_ROWTASK_MAX) {
        // Short band (the last rows of a map whose height is not a multiple of 32): lanes own 32
        // consecutive single-word tiles of ONE row instead of 32 rows of one tile, so a leftover row
        // costs one cheap warp task rather than a whole band of mostly masked lanes.
        const int groups = (nwords + 31) >> 5;
        for (int task = warp; task < nrows * groups; task += NCC_THREADS / 32) {
            const int row = task / groups, word = (task - row * groups) * 32 + lane;
            const bool valid = word < nwords;
            c.y = y0 + row;
            ncc_tile<1, MAP>(c, valid, row, valid ? word : 0, mybest);
        }
    } else {
        // a warp owns one tile x 32 rows (rows beyond the band are masked, not skipped: the filter's
        // warp-wide maximum needs all 32 lanes)
        const int ntile = n8 + n4 + n1;
        for (int tile = 0; tile < ntile; tile++) {
            // tiles are ordered by decreasing cost; the first four go to warps 0..3, later (cheaper)
            // ones to the warps that got the cheapest tiles: longest-processing-time-first balance
            const int owner = tile < 4 ? tile : 3 - ((tile - 4) & 3);
            if (owner != warp) continue;
            const bool valid = lane < nrows;
            const int yy = valid ? lane : 0;
            c.y = y0 + yy;
            if (tile < n8) ncc_tile<8, MAP>(c, valid, yy, tile * 8, mybest);
            else if (tile < n8 + n4) ncc_tile<4, MAP>(c, valid, yy, n8 * 8 + (tile - n8) * 4, mybest);
            else ncc_tile<1, MAP>(c, valid, yy, n8 * 8 + n4 * 4 + (tile - n8 - n4), mybest);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, mybest, o);
        mybest = other > mybest ? other : mybest;
    }
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = mybest;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < NCC_THREADS / 32; w++) mybest = s_best[w] > mybest ? s_best[w] : mybest;
        atomicMax(&best[blockIdx.y], mybest);
    }
}

static int ncc_pitch(int chunk_words, int kw)
{
    // widest read: an 8-word tile at word (chunk_words - 8) loading words [k, k + 12) for k <= kw - 4,
    // i.e. up to word chunk_words + kw; a 4-word tile reads up to chunk_words - 4 + kw - 4 + 8: the same
    int p = 4 * chunk_words + 4 * (kw + 1);
    p = (p + 15) & ~15;
    if (((p / 16) & 1) == 0) p += 16;   // odd multiple of 16 B: conflict-free 128-bit row loads
    return p;
}
static int ncc_vpitch(int chunk_words, int tw)
{
    int v = 4 * chunk_words + tw;
    return v | 1;                        // odd: conflict-free when 32 lanes own 32 rows
}
static size_t ncc_smem(int chunk_words, int tw, int th)
{
    int kw = ((tw + 3) / 4 + 3) & ~3;
    return (size_t)(NCC_BAND + th - 1) * ncc_pitch(chunk_words, kw) + (size_t)NCC_BAND * ncc_vpitch(chunk_words, tw) * 6 + 16;
}

void plan_match(int max_rw, int max_rh, int tw, int th, int n_tasks, MatchParams *mp, size_t *smem_bytes)
{
    // One CTA = 32 result rows x chunk_words*4 result columns, 4 warps, one 8-word tile per warp at a
    // time.  Pick the chunk width that keeps the most warps busy per SM: wide chunks give every warp of
    // the CTA a tile but cost shared memory (window rows + vertical sums), i.e. resident CTAs.  Small
    // templates fit a whole result row in ~50 KB (4 CTAs/SM, 16 warps); 128x128 templates need ~90 KB
    // for 4 tiles (2 CTAs/SM, 8 warps), which still beats narrow chunks that idle 3 of 4 warps.
    const int words_total = (max_rw + 3) / 4;
    const int cw_max = (words_total + 7) & ~7;
    const size_t smem_sm = 220 * 1024;
    int best_cw = 8;
    double best_score = -1.0;
    for (int cw = 8; cw <= cw_max && cw <= 64; cw += 8) {
        size_t sm = ncc_smem(cw, tw, th);
        if (sm > 200 * 1024) break;
        int ctas = (int)(smem_sm / (sm + 1024));
        if (ctas > 4) ctas = 4;                       // registers: 4 CTAs of 128 threads
        if (ctas < 1) break;
        int tiles = (std::min(cw, words_total) + 7) / 8;
        int busy = ctas * (tiles < 4 ? tiles : 4);
        // prefer more busy warps; among equals the wider chunk (less redundant window staging)
        double score = busy + 0.01 * cw;
        if (score > best_score) {
            best_score = score;
            best_cw = cw;
        }
    }
    int cw = best_cw;
    // too few CTAs to fill the machine (single windows): split further
    auto ctas_total = [&](int w) { return (long)n_tasks * ((words_total + w - 1) / w) * ((max_rh + NCC_BAND - 1) / NCC_BAND); };
    while (cw > 8 && ctas_total(cw) < 148L * 4) cw = (((cw + 1) / 2) + 7) & ~7;
    mp->chunk_words = cw;
    mp->smem_pitch = ncc_pitch(cw, ((tw + 3) / 4 + 3) & ~3);
    mp->v_pitch = ncc_vpitch(cw, tw);
    mp->max_th = th;
    *smem_bytes = ncc_smem(cw, tw, th);
}

int launch_ncc_match(const DevTask *tasks, int n_tasks, int max_rw, int max_rh, int tw, int th,
                     const DevTemplates *tpls, const uint8_t *win8, unsigned long long *best, float *score_map,
                     const MatchParams &mp, cudaStream_t stream)
{
    const int words_total = (max_rw + 3) / 4;
    const int chunks = (words_total + mp.chunk_words - 1) / mp.chunk_words;
    const int bands = (max_rh + NCC_BAND - 1) / NCC_BAND;
    size_t smem = ncc_smem(mp.chunk_words, tw, th);
    int launches = 0;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 grid(chunks * bands, n);
        if (score_map)
            lst_ncc_match<true><<<grid, NCC_THREADS, smem, stream>>>(tasks + base, tpls, win8, best + base, score_map, mp);
        else
            lst_ncc_match<false><<<grid, NCC_THREADS, smem, stream>>>(tasks + base, tpls, win8, best + base, score_map, mp);
        launches++;
    }
    return launches;
}


// ------------------------------------------------------------------------------------------------
// per-window epilogue: 3x3 neighbourhood, sub-pixel fit, accept test
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288)
lst_epilogue(const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls, const uint8_t *__restrict__ win8,
             const unsigned long long *__restrict__ best, DevResult *__restrict__ results, MatchParams mp)
{
    // nine warps: warp w recomputes the score of neighbour (w%3-1, w/3-1) of the peak exactly
    // (correlation and window sums in integers, normalisation by ncc_score as in the match kernel)
    __shared__ float nb[9];
    const DevTask t = tasks[blockIdx.x];
    const TplConst tc = tpls->c[t.tpl];
    const unsigned long long key = best[blockIdx.x];
    const uint32_t idx = 0xffffffffu - (uint32_t)(key & 0xffffffffull);
    const int ix = (int)(idx % (uint32_t)t.rw), iy = (int)(idx / (uint32_t)t.rw);
    const float peak = key_score(key, method_wants_min(tc.method));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nx = ix + (warp % 3) - 1, ny = iy + (warp / 3) - 1;
    float sc = 0.0f;
    if (nx >= 0 && nx < t.rw && ny >= 0 && ny < t.rh) {
        const uint8_t *w = win8 + t.win_off + (size_t)ny * t.pitch + nx;
        const uint8_t *tp = tpls->raw[t.tpl];
        uint32_t acc = 0, s = 0, q = 0;
        for (int e = lane; e < tc.tw * tc.th; e += 32) {
            int v = e / tc.tw, u = e - v * tc.tw;
            uint32_t px = w[(size_t)v * t.pitch + u];
            acc += px * (uint32_t)tp[e];
            s += px;
            q += px * px;
        }
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        sc = match_score(tc, (double)acc, (double)s, (double)q);
    }
    if (lane == 0) nb[warp] = sc;
    __syncthreads();
    if (threadIdx.x == 0) {
        DevResult r;
        r.ix = ix;
        r.iy = iy;
        r.score = (double)peak;
        r.peak_check = (nb[4] + 0.0f == peak) ? 1 : 0;
        int border = (ix == 0 || ix == t.rw - 1 || iy == 0 || iy == t.rh - 1);
        double dx = 0.0, dy = 0.0;
        if (mp.sub_pixel) subpixel_fit(nb, border, &dx, &dy);
        r.x = (double)ix + dx;
        r.y = (double)iy + dy;
        r.accepted = test_match_result(tc.method, r.score, tc.mideal, mp.thresh, r.x, r.y, 2 * t.s_used, mp.templ_size);
        for (int i = 0; i < 9; i++) r.nb[i] = nb[i];
        r.pad = 0.0f;
        results[blockIdx.x] = r;
    }
}

int launch_epilogue(const DevTask *tasks, int n_tasks, const DevTemplates *tpls, const uint8_t *win8,
                    const unsigned long long *best, DevResult *results, const MatchParams &mp, cudaStream_t stream)
{
    lst_epilogue<<<n_tasks, 288, 0, stream>>>(tasks, tpls, win8, best, results, mp);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// calcDisplacement per frame
// ------------------------------------------------------------------------------------------------
struct SolveArgs {
    float ref[10];
    double mark_dist;
    double spot0[2];
};

__global__ void lst_solve(const double *__restrict__ dis10, int n, SolveArgs a, double *__restrict__ out5)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d[10], o[5], s0[2] = {a.spot0[0], a.spot0[1]};
#pragma unroll
    for (int k = 0; k < 10; k++) d[k] = dis10[(size_t)i * 10 + k];
    calc_displacement(a.ref, d, a.mark_dist, 1, s0, o);
#pragma unroll
    for (int k = 0; k < 5; k++) out5[(size_t)i * 5 + k] = o[k];
}

int launch_solve(const double *dis10, int n, const float *ref_xy, double mark_dist, const double *spot0,
                 double *out5, cudaStream_t stream)
{
    SolveArgs a;
    for (int i = 0; i < 10; i++) a.ref[i] = ref_xy[i];
    a.mark_dist = mark_dist;
    a.spot0[0] = spot0[0];
    a.spot0[1] = spot0[1];
    lst_solve<<<(n + 127) / 128, 128, 0, stream>>>(dis10, n, a, out5);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// One shared-memory carve-out for every kernel of the library.  An SM cannot host CTAs of kernels
// that were given different L1/shared splits at the same time; the ROI gather (no shared memory)
// runs on the copy stream *while* the match kernels (200+ KB of shared memory per SM) run on the
// compute stream, so all kernels ask for the maximum-shared split and can co-reside.
static void prefer_max_shared_everywhere()
{
    // function attributes belong to the current device: this runs once per lst_create
    cudaFuncSetAttribute(lst_preprocess<LST_PIX_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(lst_preprocess<LST_PIX_U16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(lst_preprocess<LST_PIX_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(lst_ncc_match<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(lst_ncc_match<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int mx = (int)cudaSharedmemCarveoutMaxShared;
    cudaFuncSetAttribute(lst_init_pass, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_upload, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_minmax<LST_PIX_U8>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_minmax<LST_PIX_U16>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_minmax<LST_PIX_F32>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_preprocess<LST_PIX_U8>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_preprocess<LST_PIX_U16>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_preprocess<LST_PIX_F32>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_ncc_match<false>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_ncc_match<true>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_epilogue, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_solve, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(lst_gather_rois, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// ALU peak microbenchmarks (roofline denominators: SURVEY.md 8d asks for a measured IDP.4A / FFMA peak)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lst_peak_dp4a(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[16];
    const uint32_t b0 = seed + threadIdx.x, b1 = seed * 3u + blockIdx.x, b2 = seed ^ 0x01020304u, b3 = ~seed;
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                a[i] = __dp4a(r & 1 ? b0 : b2, b1, a[i]);
                a[i + 1] = __dp4a(r & 1 ? b2 : b0, b3, a[i + 1]);
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 0x12345678u) out[0] = s;   // never true in practice; keeps the loop alive
}

__global__ void __launch_bounds__(256) lst_peak_ffma(float *out, int iters, float seed)
{
    float a[16];
    float b0 = seed + threadIdx.x * 1e-3f, b1 = seed * 0.5f;
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fmaf_rn(a[i], b0, b1);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 123.456f) out[0] = s;
}

int measure_alu_peaks(double ms_per_test, double out[3])
{
    uint32_t *d = nullptr;
    if (cudaMalloc(&d, 256) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    for (int which = 0; which < 2; which++) {
        int iters = 2000;
        double best = 0.0;
        for (int rep = 0; rep < 6; rep++) {
            cudaEventRecord(e0);
            if (which == 0) lst_peak_dp4a<<<blocks, threads>>>(d, iters, 12345u + rep);
            else lst_peak_ffma<<<blocks, threads>>>((float *)d, iters, 1.0001f);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) {
                cudaFree(d);
                return -1;
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            double ops = (double)blocks * threads * (double)iters * 64.0 * (which == 0 ? 4.0 : 1.0);
            double rate = ops / (ms * 1e-3);
            if (rep > 0 && rate > best) best = rate;
            if (rep == 0 && ms > 0) {   // scale the loop so one launch lasts ~ms_per_test
                double f = ms_per_test / ms;
                if (f > 64) f = 64;
                if (f < 0.25) f = 0.25;
                iters = (int)(iters * f);
                if (iters < 100) iters = 100;
            }
        }
        out[which] = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    out[2] = measure_imma_peak(ms_per_test);
    return 0;
}

}  // namespace lst
