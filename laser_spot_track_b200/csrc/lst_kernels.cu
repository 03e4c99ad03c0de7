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
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>

#include "../../include/lst_b200.h"
#include "lst_device.h"

namespace lst {

#define KR 7   // kRadius of the sigma=2 / accuracy=0.02 Gaussian (13 taps)

__constant__ float c_kern[KR];
__constant__ float c_ksum[KR];
__constant__ double c_rgbw[3] = {1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0};   // ColorProcessor weighting factors

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

int set_rgb_weights(const double w[3])
{
    return cudaMemcpyToSymbol(c_rgbw, w, sizeof(double) * 3) == cudaSuccess ? 0 : -1;
}

// TypeConverter.convertRGBToByte [ImageJ, 3P-recall]: (byte)(r*rw + g*gw + b*bw + 0.5) in double
__device__ __forceinline__ uint32_t rgb_gray(uint32_t c)
{
    const double r = (double)((c >> 16) & 0xffu), g = (double)((c >> 8) & 0xffu), b = (double)(c & 0xffu);
    return (uint32_t)(int)(r * c_rgbw[0] + g * c_rgbw[1] + b * c_rgbw[2] + 0.5) & 0xffu;
}

template <int PIX>
__device__ __forceinline__ float load_px(const void *base, size_t idx)
{
    if (PIX == LST_PIX_RGB) return (float)rgb_gray(((const uint32_t *)base)[idx]);
    if (PIX == LST_PIX_U8) return (float)((const uint8_t *)base)[idx];
    if (PIX == LST_PIX_U16) return (float)((const uint16_t *)base)[idx];
    return ((const float *)base)[idx];
}
template <int PIX>
__device__ __forceinline__ uint32_t px_key(const void *base, size_t idx)
{
    if (PIX == LST_PIX_RGB) return rgb_gray(((const uint32_t *)base)[idx]);
    if (PIX == LST_PIX_U8) return ((const uint8_t *)base)[idx];
    if (PIX == LST_PIX_U16) return ((const uint16_t *)base)[idx];
    return f32_key(((const float *)base)[idx]);
}
// the same for a slice whose samples are stored big-endian (TIFF "MM" files, lst_track_files)
template <int PIX>
__device__ __forceinline__ uint32_t px_key_swapped(const void *base, size_t idx)
{
    if (PIX == LST_PIX_RGB) return rgb_gray(((const uint32_t *)base)[idx]);
    if (PIX == LST_PIX_U8) return ((const uint8_t *)base)[idx];
    if (PIX == LST_PIX_U16) return __byte_perm((uint32_t)((const uint16_t *)base)[idx], 0, 0x3201);
    return f32_key(__uint_as_float(__byte_perm(((const uint32_t *)base)[idx], 0, 0x0123)));
}
template <int PIX>
__device__ __forceinline__ float key_to_float(uint32_t k)
{
    if (PIX == LST_PIX_F32) return key_f32(k);
    return (float)k;   // u8, u16, grey value of an RGB pixel
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
__global__ void lst_init_pass(int n_tasks, uint32_t *mm_keys, unsigned long long *best, float *nbhd)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tasks) {
        mm_keys[2 * i] = 0xffffffffu;
        mm_keys[2 * i + 1] = 0u;
        best[i] = 0ull;
        nbhd[12 * i + 9] = 0.0f;   // "the match kernel left the 3x3 neighbourhood of the peak here"
    }
}

int launch_init_pass(int n_tasks, uint32_t *mm_keys, unsigned long long *best, float *nbhd, cudaStream_t stream)
{
    lst_init_pass<<<(n_tasks + 255) / 256, 256, 0, stream>>>(n_tasks, mm_keys, best, nbhd);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// min / max of the raw crop
// ------------------------------------------------------------------------------------------------
template <int PIX>
__device__ __forceinline__ void mm_word(uint32_t wlo, uint32_t whi, uint32_t &lo, uint32_t &hi);
template <int PIX>
__device__ __forceinline__ void mm_finish(uint32_t &lo, uint32_t &hi);

template <int PIX>
__global__ void lst_minmax(const DevTask *__restrict__ tasks, const Surface *__restrict__ surfaces,
                           uint32_t *__restrict__ mm_keys, int swap_bytes)
{
    const DevTask t = tasks[blockIdx.y];
    if (t.pad0) return;
    const Surface sf = surfaces[t.frame];
    const void *fr = sf.ptr;
    uint32_t lo = 0xffffffffu, hi = 0u;
    constexpr int ES = PIX == LST_PIX_U8 ? 1 : (PIX == LST_PIX_U16 ? 2 : 4);
    for (int r = blockIdx.x; r < t.h; r += gridDim.x) {
        size_t row = (size_t)(t.y0 - sf.org_y + r) * sf.pitch_px + (t.x0 - sf.org_x);
        int c_first = 0;
        if (PIX == LST_PIX_U8 || PIX == LST_PIX_U16) {
            // whole 16-byte vectors of the row with packed min / max (whole slices: the rows start aligned); the rest pixel by pixel
            const uint8_t *rp = (const uint8_t *)fr + row * ES;
            if (((uintptr_t)rp & 15u) == 0) {
                const int nvec = (t.w * ES) >> 4;
                uint32_t plo = 0xffffffffu, phi = 0u;
                for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
                    const uint4 q = __ldg((const uint4 *)rp + v);
                    uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (PIX == LST_PIX_U16 && swap_bytes) w4[j] = __byte_perm(w4[j], 0, 0x2301);
                        mm_word<PIX>(w4[j], w4[j], plo, phi);
                    }
                }
                mm_finish<PIX>(plo, phi);
                lo = min(lo, plo);
                hi = max(hi, phi);
                c_first = (nvec << 4) / ES;
            }
        }
        for (int c = c_first + threadIdx.x; c < t.w; c += blockDim.x) {
            uint32_t k = swap_bytes ? px_key_swapped<PIX>(fr, row + c) : px_key<PIX>(fr, row + c);
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
            lst_minmax<LST_PIX_U8><<<grid, 128, 0, stream>>>(tasks + base, surfaces, mm_keys + 2 * base, pp.swap_bytes);
        else if (pp.pixel_type == LST_PIX_U16)
            lst_minmax<LST_PIX_U16><<<grid, 128, 0, stream>>>(tasks + base, surfaces, mm_keys + 2 * base, pp.swap_bytes);
        else if (pp.pixel_type == LST_PIX_RGB)
            lst_minmax<LST_PIX_RGB><<<grid, 128, 0, stream>>>(tasks + base, surfaces, mm_keys + 2 * base, pp.swap_bytes);
        else
            lst_minmax<LST_PIX_F32><<<grid, 128, 0, stream>>>(tasks + base, surfaces, mm_keys + 2 * base, pp.swap_bytes);
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
// ImageJ's GaussianBlur.convolveLine, float32, Java operation order, for a run of NOUT consecutive
// outputs whose inputs v[0 .. NOUT + 2*KR - 3] start KR-1 samples before the first output.  Samples
// outside the line must be 0 in v: the reference adds only the in-range one of the two mirrored taps
// (0 + x == x), and -- for the KR-1 outputs next to an end -- the replicated end sample times the
// tail sum of the kernel *before* the tap loop.  i0 = line index of the first output, n = line length.
template <int NOUT, bool EDGE>
__device__ __forceinline__ void blur_run(const float *v, float *out, int i0, int n, float first, float last)
{
#pragma unroll
    for (int j = 0; j < NOUT; j++) {
        float res = v[KR - 1 + j] * c_kern[0];
        if (EDGE) {
            const int i = i0 + j;
            if (i < KR) res = res + c_ksum[i] * first;
            if (i + KR >= n && i < n) res = res + c_ksum[n - i - 1] * last;
        }
#pragma unroll
        for (int k = 1; k < KR; k++) res = res + c_kern[k] * (v[KR - 1 + j - k] + v[KR - 1 + j + k]);
        out[j] = res;
    }
}

// ---- raw pixel staging --------------------------------------------------------------------------
// Packed min/max of raw pixel words (u8 x 4, u16 x 2 or one float key per 32-bit word); wlo / whi are
// the word with the bytes that do not count forced to 0xff / 0x00.
template <int PIX>
__device__ __forceinline__ void mm_word(uint32_t wlo, uint32_t whi, uint32_t &lo, uint32_t &hi)
{
    if (PIX == LST_PIX_U8) {
        lo = __vminu4(lo, wlo);
        hi = __vmaxu4(hi, whi);
    } else if (PIX == LST_PIX_U16) {
        lo = __vminu2(lo, wlo);
        hi = __vmaxu2(hi, whi);
    } else {
        const uint32_t k = PIX == LST_PIX_RGB ? rgb_gray(whi) : f32_key(__uint_as_float(whi));
        lo = min(lo, k);
        hi = max(hi, k);
    }
}
template <int PIX>
__device__ __forceinline__ void mm_finish(uint32_t &lo, uint32_t &hi)   // packed accumulators -> keys
{
    if (PIX == LST_PIX_U8) {
        lo = min(min(lo & 0xffu, (lo >> 8) & 0xffu), min((lo >> 16) & 0xffu, lo >> 24));
        hi = max(max(hi & 0xffu, (hi >> 8) & 0xffu), max((hi >> 16) & 0xffu, hi >> 24));
    } else if (PIX == LST_PIX_U16) {
        lo = min(lo & 0xffffu, lo >> 16);
        hi = max(hi & 0xffffu, hi >> 16);
    }
}
// bytes [s, e) of a 32-bit word as a mask (s, e clamped to 0..4)
__device__ __forceinline__ uint32_t byte_range_mask(int s, int e)
{
    s = max(s, 0);
    e = min(e, 4);
    return e > s ? ((0xffffffffu >> (8 * (4 - (e - s)))) << (8 * s)) : 0u;
}

// Rows [row0, row0 + nrows) of the crop, byte columns [lo0, lo0 + 16*vpr) (lo0 may be negative), as 16-byte
// vectors; a warp takes a row, a lane a vector.  The crop starts at an arbitrary pixel of the
// surface, so a vector is fetched as the two aligned 128-bit words around it and realigned with
// funnel shifts (the shift is the same for the whole row); bytes outside the crop row are masked
// to 0.  Only when the aligned words would leave the surface is the vector assembled pixel by
// pixel.  STORE: the pixels go to sraw[r][.] as floats; MINMAX: the in-crop pixels are folded into
// the packed accumulators.
template <int PIX, bool STORE, bool MINMAX>
__device__ __forceinline__ void stage_rows(const uint8_t *__restrict__ crop00, size_t row_pitch_bytes, int row0, int nrows,
                                           int lo0, int vpr, int row_bytes, const uint8_t *surf_lo, const uint8_t *surf_hi,
                                           float *sraw, int rwp, uint32_t &mlo, uint32_t &mhi, bool swap, int rgb_shift = -1)
{
    constexpr int ES = PIX == LST_PIX_U8 ? 1 : (PIX == LST_PIX_U16 ? 2 : 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < nrows; r += nwarps) {
        const uint8_t *grow = crop00 + (size_t)(row0 + r) * row_pitch_bytes;
        const uint8_t *p0 = grow + lo0;
        const uint32_t a = (uint32_t)((uintptr_t)p0 & 15u);
        const uint32_t bs = 8u * (a & 3u);
        for (int q = lane; q < vpr; q += 32) {
            const int lo = lo0 + 16 * q;                       // crop byte of this vector's first byte
            uint32_t o[4] = {0u, 0u, 0u, 0u};
            if (lo + 16 > 0 && lo < row_bytes) {
                const uint8_t *pa = p0 + 16 * q - a;
                if (pa >= surf_lo && pa + 32 <= surf_hi) {
                    const uint4 v0 = __ldg((const uint4 *)pa), v1 = __ldg((const uint4 *)pa + 1);
                    uint32_t w0 = v0.x, w1 = v0.y, w2 = v0.z, w3 = v0.w, w4 = v1.x, w5 = v1.y, w6 = v1.z, w7 = v1.w;
                    if (a & 8u) w0 = w2, w1 = w3, w2 = w4, w3 = w5, w4 = w6, w5 = w7;
                    if (a & 4u) w0 = w1, w1 = w2, w2 = w3, w3 = w4, w4 = w5;
                    o[0] = __funnelshift_r(w0, w1, bs);
                    o[1] = __funnelshift_r(w1, w2, bs);
                    o[2] = __funnelshift_r(w2, w3, bs);
                    o[3] = __funnelshift_r(w3, w4, bs);
                    if (PIX != LST_PIX_U8 && PIX != LST_PIX_RGB && swap) {   // big-endian samples (TIFF "MM")
#pragma unroll
                        for (int j = 0; j < 4; j++) o[j] = __byte_perm(o[j], 0, PIX == LST_PIX_U16 ? 0x2301 : 0x0123);
                    }
                    if (lo >= 0 && lo + 16 <= row_bytes) {
                        if (MINMAX) {
#pragma unroll
                            for (int j = 0; j < 4; j++) mm_word<PIX>(o[j], o[j], mlo, mhi);
                        }
                    } else {
                        const int b0 = -lo, b1 = row_bytes - lo;   // valid bytes of the vector: [b0, b1)
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t m = byte_range_mask(b0 - 4 * j, b1 - 4 * j);
                            o[j] &= m;
                            if (MINMAX) {
                                if (PIX != LST_PIX_F32 && PIX != LST_PIX_RGB) mm_word<PIX>(o[j] | ~m, o[j], mlo, mhi);
                                else if (m) mm_word<PIX>(o[j], o[j], mlo, mhi);
                            }
                        }
                    }
                } else {
#pragma unroll 1
                    for (int e = 0; e < 16 / ES; e++) {
                        const int bo = lo + e * ES;            // crop byte offset of this pixel
                        if (bo >= 0 && bo < row_bytes) {
                            uint32_t v;
                            if (PIX == LST_PIX_U8) v = grow[bo];
                            else if (PIX == LST_PIX_U16) v = *(const uint16_t *)(grow + bo);
                            else v = *(const uint32_t *)(grow + bo);
                            if (PIX != LST_PIX_U8 && PIX != LST_PIX_RGB && swap) v = PIX == LST_PIX_U16 ? __byte_perm(v, 0, 0x3201) : __byte_perm(v, 0, 0x0123);
                            const uint32_t sh = v << (8 * ((e * ES) & 3));
                            const int wi = (e * ES) >> 2;
                            o[0] |= wi == 0 ? sh : 0u, o[1] |= wi == 1 ? sh : 0u, o[2] |= wi == 2 ? sh : 0u, o[3] |= wi == 3 ? sh : 0u;
                            if (MINMAX) {
                                const uint32_t rep = PIX == LST_PIX_U8 ? v * 0x01010101u : (PIX == LST_PIX_U16 ? v * 0x00010001u : v);
                                mm_word<PIX>(rep, rep, mlo, mhi);
                            }
                        }
                    }
                }
            }
            if (STORE) {
                float *d = sraw + (size_t)r * rwp + q * (16 / ES);
                if (PIX == LST_PIX_U8) {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        ((float4 *)d)[j] = make_float4((float)(o[j] & 0xffu), (float)((o[j] >> 8) & 0xffu),
                                                       (float)((o[j] >> 16) & 0xffu), (float)(o[j] >> 24));
                } else if (PIX == LST_PIX_U16) {
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        ((float4 *)d)[j] = make_float4((float)(o[2 * j] & 0xffffu), (float)(o[2 * j] >> 16),
                                                       (float)(o[2 * j + 1] & 0xffffu), (float)(o[2 * j + 1] >> 16));
                } else if (PIX == LST_PIX_RGB) {
                    if (rgb_shift >= 0)      // one channel of the ColorProcessor (GaussianBlur blurs the channels separately)
                        *(float4 *)d = make_float4((float)((o[0] >> rgb_shift) & 0xffu), (float)((o[1] >> rgb_shift) & 0xffu),
                                                   (float)((o[2] >> rgb_shift) & 0xffu), (float)((o[3] >> rgb_shift) & 0xffu));
                    else
                        *(float4 *)d = make_float4((float)rgb_gray(o[0]), (float)rgb_gray(o[1]), (float)rgb_gray(o[2]), (float)rgb_gray(o[3]));
                } else {
                    *(float4 *)d = make_float4(__uint_as_float(o[0]), __uint_as_float(o[1]), __uint_as_float(o[2]), __uint_as_float(o[3]));
                }
            }
        }
    }
}

// A CTA owns a column block (up to tile_w columns, the whole window for the named configurations) of
// one window over its full height and walks it in strips of PRE_STRIP rows:
//   sraw [PRE_STRIP + KR-1][RWP]  the strip's new crop rows as floats, columns from cx0 - 8, zero outside
//                                 the crop (so every run takes the same code path), fetched with
//                                 128-bit loads (stage_rows);
//   xp   [PRE_STRIP + 2(KR-1)][XWP] X-pass rows cy0-(KR-1) .. cy1+(KR-1)-1; the 2(KR-1) rows two strips
//                                 share are moved to the front between strips, not recomputed.
// X pass: a warp takes a run of 8 columns, its lanes consecutive rows (six 128-bit loads -> 8
// results, two 128-bit stores; both pitches are odd multiples of 16 bytes: conflict-free).  Y pass:
// a thread takes 8 rows of one column (20 loads at constant offsets -> 8 results), consecutive
// threads own consecutive columns.  When the range convention needs the crop's min/max and the
// window is a single column block, the CTA finds them itself first with the same 128-bit loads (no
// separate kernel).
#define PRE_STRIP 32
#define PRE_RING (PRE_STRIP + 2 * (KR - 1))   // 44
#define PRE_XRUN 8
#define PRE_YRUN 8

__host__ __device__ inline int pre_raw_pitch(int tile_w) { int v = tile_w + 32; return ((v >> 2) & 1) ? v : v + 4; }
__host__ __device__ inline int pre_xp_pitch(int tile_w) { return ((tile_w >> 2) & 1) ? tile_w : tile_w + 4; }

template <int PIX>
__global__ void __launch_bounds__(512, 2)
lst_preprocess(const DevTask *__restrict__ tasks, const Surface *__restrict__ surfaces, PreprocParams pp,
               const uint32_t *__restrict__ mm_keys, uint8_t *__restrict__ win8)
{
    constexpr int ES = PIX == LST_PIX_U8 ? 1 : (PIX == LST_PIX_U16 ? 2 : 4);
    extern __shared__ __align__(16) float smem_f[];
    __shared__ uint32_t s_mm[2];
    const DevTask t = tasks[blockIdx.y];
    if (t.pad0) return;                        // shares the 8-bit window of an earlier task (whole-slice search)
    const int TW = pp.tile_w;                  // multiple of 16
    const int tiles_x = (t.w + TW - 1) / TW;
    if ((int)blockIdx.x >= tiles_x) return;
    const int NT = blockDim.x;
    const int cx0 = blockIdx.x * TW;
    const int cx1 = min(cx0 + TW, t.w);
    const int ncw = cx1 - cx0;
    const int nrun = (ncw + PRE_XRUN - 1) / PRE_XRUN;
    const int RWP = pre_raw_pitch(TW), XWP = pre_xp_pitch(TW);
    const int RB = cx0 - 8;                    // crop column held by sraw[.][0]
    const int VPR = ((PRE_XRUN * nrun + 16) * ES + 15) >> 4;   // 16-byte vectors of a crop row the X pass needs
    float *xp = smem_f;                                  // [PRE_RING][XWP]
    float *sraw = smem_f + (size_t)PRE_RING * XWP;       // [PRE_STRIP + KR - 1][RWP]
    const Surface sf = surfaces[t.frame];
    const size_t row_pitch_bytes = (size_t)sf.pitch_px * ES;
    const uint8_t *surf_lo = (const uint8_t *)sf.ptr, *surf_hi = surf_lo + (size_t)sf.rows * row_pitch_bytes;
    const uint8_t *crop00 = surf_lo + ((size_t)(t.y0 - sf.org_y) * sf.pitch_px + (t.x0 - sf.org_x)) * ES;
    const int row_bytes = t.w * ES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = NT >> 5;

    float lo = pp.lo, hi = pp.hi;
    if (pp.range_per_task == 2) {              // the whole slice's min/max (LST_RANGE_FRAME)
        lo = key_to_float<PIX>(pp.frame_mm[2 * t.pad1]);
        hi = key_to_float<PIX>(pp.frame_mm[2 * t.pad1 + 1]);
    } else if (pp.range_per_task) {
        if (pp.minmax_in_kernel) {
            if (threadIdx.x == 0) {
                s_mm[0] = 0xffffffffu;
                s_mm[1] = 0u;
            }
            __syncthreads();
            uint32_t klo = 0xffffffffu, khi = 0u;
            // vectors aligned to the crop's first pixel: only the last one of a row is partial
            stage_rows<PIX, false, true>(crop00, row_pitch_bytes, 0, t.h, 0, (row_bytes + 15) >> 4, row_bytes, surf_lo, surf_hi,
                                         nullptr, 0, klo, khi, pp.swap_bytes != 0);
            mm_finish<PIX>(klo, khi);
            for (int o = 16; o > 0; o >>= 1) {
                klo = min(klo, __shfl_xor_sync(0xffffffffu, klo, o));
                khi = max(khi, __shfl_xor_sync(0xffffffffu, khi, o));
            }
            if (lane == 0) {
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
    const float qadd = pp.quant_mode == LST_QUANT_ROUND255 ? 0.5f : 0.0f;
    uint8_t *dst = win8 + t.win_off;
    // Y-pass items (8-row run sg, column cc) of this thread: idx = tid, tid + NT, ... without a division per item
    const int ysg0 = threadIdx.x / ncw, ycc0 = threadIdx.x - ysg0 * ncw;
    const int ydsg = NT / ncw, ydcc = NT - ydsg * ncw;

    // Tall windows (whole-slice search) are split into bands of band_rows output rows, one CTA (blockIdx.z) each; a band
    // below the first starts KR-1 X-pass rows early, so its first strip is that much shorter (sraw holds PRE_STRIP + KR-1 rows).
    const int y_begin = pp.band_rows > 0 ? (int)blockIdx.z * pp.band_rows : 0;
    if (y_begin >= t.h) return;
    const int y_end = pp.band_rows > 0 ? min(y_begin + pp.band_rows, t.h) : t.h;
    int xdone = max(0, y_begin - (KR - 1));   // x-pass rows [.., xdone) have been computed
    int prev_len = 0;
    for (int cy0 = y_begin; cy0 < y_end; cy0 += prev_len) {
        const int cy1 = min(cy0 + (cy0 == y_begin && y_begin > 0 ? PRE_STRIP - (KR - 1) : PRE_STRIP), y_end);
        const int need = min(cy1 + (KR - 1), t.h);          // x-pass rows needed up to here
        const int newr = need - xdone;                      // <= PRE_STRIP + KR - 1
        const int xrow0 = cy0 - (KR - 1);                   // crop row held by xp[0][.]
        // raw rows [xdone, need), columns from RB, zero outside the crop
        {
            uint32_t d0 = 0, d1 = 0;
            stage_rows<PIX, true, false>(crop00, row_pitch_bytes, xdone, newr, RB * ES, VPR, row_bytes, surf_lo, surf_hi, sraw,
                                         RWP, d0, d1, pp.swap_bytes != 0, PIX == LST_PIX_RGB ? pp.rgb_shift : -1);
        }
        // the X-pass rows this strip shares with the previous one move to the front
        if (cy0 > y_begin) {
            const int n4 = XWP >> 2;
            for (int r = warp; r < 2 * (KR - 1); r += nwarps)
                for (int c = lane; c < n4; c += 32)
                    ((float4 *)(xp + (size_t)r * XWP))[c] = ((const float4 *)(xp + (size_t)(prev_len + r) * XWP))[c];
        }
        __syncthreads();
        // X pass of the new rows
        for (int run = warp; run < nrun; run += nwarps) {
            const int i0 = cx0 + PRE_XRUN * run;
            const bool interior = i0 >= KR && i0 + PRE_XRUN - 1 + KR < t.w;
            for (int r = lane; r < newr; r += 32) {
                const float *line = sraw + (size_t)r * RWP;
                float v[PRE_XRUN + 16];
#pragma unroll
                for (int q = 0; q < (PRE_XRUN + 16) / 4; q++) {
                    const float4 f = ((const float4 *)(line + PRE_XRUN * run))[q];      // crop columns i0 - 8 + 4q ..
                    v[4 * q] = f.x, v[4 * q + 1] = f.y, v[4 * q + 2] = f.z, v[4 * q + 3] = f.w;
                }
                float o[PRE_XRUN];
                if (interior)
                    blur_run<PRE_XRUN, false>(v + 8 - (KR - 1), o, i0, t.w, 0.0f, 0.0f);
                else
                    blur_run<PRE_XRUN, true>(v + 8 - (KR - 1), o, i0, t.w, line[0 - RB > 0 ? 0 - RB : 0],
                                             line[t.w - 1 - RB < RWP ? t.w - 1 - RB : 0]);
                float4 *dx = (float4 *)(xp + (size_t)(xdone + r - xrow0) * XWP + PRE_XRUN * run);
#pragma unroll
                for (int q = 0; q < PRE_XRUN / 4; q++) dx[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            }
        }
        xdone = need;
        __syncthreads();
        // Y pass + quantise of rows [cy0, cy1)
        const int nch = cy1 - cy0;
        const int nsegy = (nch + PRE_YRUN - 1) / PRE_YRUN;
        for (int sg = ysg0, cc = ycc0; sg < nsegy;) {
            const int i0 = cy0 + PRE_YRUN * sg;
            const int nout = min(PRE_YRUN, cy1 - i0);
            const float *col = xp + (size_t)(PRE_YRUN * sg) * XWP + cc;     // x-pass value at crop row i0 - (KR-1)
            float v[PRE_YRUN + 2 * (KR - 1)], o[PRE_YRUN];
            if (i0 >= KR && i0 + PRE_YRUN - 1 + KR < t.h) {
#pragma unroll
                for (int j = 0; j < PRE_YRUN + 2 * (KR - 1); j++) v[j] = col[j * XWP];
                blur_run<PRE_YRUN, false>(v, o, i0, t.h, 0.0f, 0.0f);
            } else {
#pragma unroll
                for (int j = 0; j < PRE_YRUN + 2 * (KR - 1); j++) {
                    const int i = i0 - (KR - 1) + j;
                    v[j] = (i >= 0 && i < t.h) ? col[j * XWP] : 0.0f;
                }
                blur_run<PRE_YRUN, true>(v, o, i0, t.h, xp[(size_t)(0 - xrow0 > 0 ? 0 - xrow0 : 0) * XWP + cc],
                                         xp[(size_t)(t.h - 1 - xrow0 < PRE_RING ? t.h - 1 - xrow0 : 0) * XWP + cc]);
            }
            uint8_t *d = dst + (size_t)i0 * t.pitch + (cx0 + cc);
            uint32_t b[PRE_YRUN];
#pragma unroll
            for (int j = 0; j < PRE_YRUN; j++) {
                // (int)(max(v - lo, 0) * scale + 0.5) clamped to 0..255, Java casts: NaN -> 0, saturating
                const float q = fmaxf(o[j] - lo, 0.0f) * scale + qadd;
                b[j] = min(__float2uint_rz(q), 255u);
            }
            if (nout == PRE_YRUN) {
#pragma unroll
                for (int j = 0; j < PRE_YRUN; j++, d += t.pitch) *d = (uint8_t)b[j];
            } else {
#pragma unroll
                for (int j = 0; j < PRE_YRUN; j++, d += t.pitch)
                    if (j < nout) *d = (uint8_t)b[j];
            }
            cc += ydcc;
            sg += ydsg;
            if (cc >= ncw) cc -= ncw, sg++;
        }
        // zero the pitch padding so the match kernel can load whole 16-byte words
        if (cx1 == t.w && t.pitch > t.w) {
            int padw = t.pitch - t.w;
            for (int idx = threadIdx.x; idx < padw * nch; idx += NT) {
                int r = idx / padw, c = idx - r * padw;
                dst[(size_t)(cy0 + r) * t.pitch + t.w + c] = 0;
            }
        }
        __syncthreads();   // the next strip overwrites sraw and moves the ring
        prev_len = cy1 - cy0;
    }
}

static size_t preprocess_smem(const PreprocParams &pp)
{
    return ((size_t)PRE_RING * pre_xp_pitch(pp.tile_w) + (size_t)(PRE_STRIP + KR - 1) * pre_raw_pitch(pp.tile_w)) * sizeof(float);
}

static int preprocess_threads(const PreprocParams &pp, int max_w)
{
    int w = max_w < pp.tile_w ? max_w : pp.tile_w;
    int nt = (2 * w + 31) & ~31;     // Y pass: columns x 4 row runs per strip = two rounds; X pass: 32 rows x w/8 runs = two rounds
    return nt < 128 ? 128 : (nt > 512 ? 512 : nt);
}

int launch_preprocess(const DevTask *tasks, int n_tasks, int max_w, int max_h, const Surface *surfaces,
                      const PreprocParams &pp, const uint32_t *mm_keys, uint8_t *win8, cudaStream_t stream)
{
    int tiles = (max_w + pp.tile_w - 1) / pp.tile_w;
    size_t smem = preprocess_smem(pp);
    const int nt = preprocess_threads(pp, max_w);
    // whole slices: bands of 256 rows, so that a pass of a few tall windows still fills the GPU (12 extra X-pass rows per band)
    PreprocParams pb = pp;
    pb.band_rows = max_h > 512 ? 256 : 0;
    const int bands = pb.band_rows > 0 ? (max_h + pb.band_rows - 1) / pb.band_rows : 1;
    int launches = 0;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 grid(tiles, n, bands);
        const uint32_t *mm = mm_keys + 2 * base;
        if (pp.pixel_type == LST_PIX_U8)
            lst_preprocess<LST_PIX_U8><<<grid, nt, smem, stream>>>(tasks + base, surfaces, pb, mm, win8);
        else if (pp.pixel_type == LST_PIX_U16)
            lst_preprocess<LST_PIX_U16><<<grid, nt, smem, stream>>>(tasks + base, surfaces, pb, mm, win8);
        else if (pp.pixel_type == LST_PIX_RGB)
            lst_preprocess<LST_PIX_RGB><<<grid, nt, smem, stream>>>(tasks + base, surfaces, pb, mm, win8);
        else
            lst_preprocess<LST_PIX_F32><<<grid, nt, smem, stream>>>(tasks + base, surfaces, pb, mm, win8);
        launches++;
    }
    return launches;
}

}  // namespace lst
#include "lst_preprocess_win.cuh"
namespace lst {

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
// 3-channel match (RGB stacks, matchIntensity = false: LST4:2525-2535 case 24 hands OpenCV a CV_8UC3 image)
// ------------------------------------------------------------------------------------------------
// matchTemplate sums the correlation over the channels and normalises with per-channel window sums.  A rarely used mode, so
// one direct kernel: a CTA owns RGB_BAND result rows x RGB_TILE result columns of one window, keeps the window rows of the
// three planes it needs (as 32-bit words) and the three template planes (rows padded with zeros to whole words) in shared
// memory; a thread walks the template for its positions four pixels at a time -- the window word realigned with a funnel
// shift, then IDP.4A for the correlation, the window sum (against 0x01 bytes) and the sum of squares (the word against its
// masked self), all exact integers -- scores them with match_score3 and the CTA reduces the first maximum.
#define RGB_BAND 4
#define RGB_TILE 128
#define RGB_THREADS 256

__global__ void __launch_bounds__(RGB_THREADS)
lst_ncc_rgb(const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls, const uint8_t *__restrict__ win8,
            unsigned long long plane_bytes, unsigned long long *__restrict__ best, int ww_pitch, int max_th, int max_kw, int ntiles)
{
    extern __shared__ __align__(16) uint32_t smem_rgb[];
    __shared__ unsigned long long s_best[RGB_THREADS / 32];
    const DevTask t = tasks[blockIdx.y];
    const TplConst tc = tpls->c[t.tpl];
    const int tw = tc.tw, th = tc.th, kw = (tw + 3) >> 2;
    const int band = blockIdx.x / ntiles, tile = blockIdx.x - band * ntiles;
    const int y0 = band * RGB_BAND, x0 = tile * RGB_TILE;
    if (y0 >= t.rh || x0 >= t.rw) return;
    const int nrows = min(RGB_BAND, t.rh - y0), ncols = min(RGB_TILE, t.rw - x0);
    const int wrows = nrows + th - 1;
    const int wwords = (ncols + tw - 1 + 3) / 4 + 1;                  // window words per row (+1: the funnel shift reads one ahead)
    uint32_t *s_win = smem_rgb;                                       // [3][RGB_BAND + max_th - 1][ww_pitch] words
    const size_t plane_smem = (size_t)(RGB_BAND + max_th - 1) * ww_pitch;
    uint32_t *s_tpl = s_win + 3 * plane_smem;                         // [3][th][kw] words, zero padded
    uint32_t *s_ones = s_tpl + 3 * (size_t)max_th * max_kw;           // [kw]: 0x01 in the bytes of a template row
    for (int c = 0; c < 3; c++) {
        const uint8_t *src = win8 + (size_t)c * plane_bytes + t.win_off + x0;      // x0 is a multiple of 4, rows 16-byte aligned
        for (int i = threadIdx.x; i < wrows * wwords; i += blockDim.x) {
            const int r = i / wwords, k = i - r * wwords;
            const int col = x0 + 4 * k;
            uint32_t v = 0;
            if (col < t.pitch) v = *(const uint32_t *)(src + (size_t)(y0 + r) * t.pitch + 4 * k);    // pitch padding is zero
            s_win[c * plane_smem + (size_t)r * ww_pitch + k] = v;
        }
        const uint8_t *tp = tpls->raw[t.tpl] + (size_t)c * tw * th;
        for (int i = threadIdx.x; i < th * kw; i += blockDim.x) {
            const int v = i / kw, k = i - v * kw;
            uint32_t word = 0;
            for (int b = 0; b < 4; b++)
                if (4 * k + b < tw) word |= (uint32_t)tp[v * tw + 4 * k + b] << (8 * b);
            s_tpl[((size_t)c * th + v) * kw + k] = word;
        }
    }
    for (int k = threadIdx.x; k < kw; k += blockDim.x) {
        uint32_t m = 0;
        for (int b = 0; b < 4; b++)
            if (4 * k + b < tw) m |= 1u << (8 * b);
        s_ones[k] = m;
    }
    __syncthreads();
    unsigned long long mybest = 0ull;
    const bool want_min = method_wants_min(tc.method);
    for (int pos = threadIdx.x; pos < nrows * ncols; pos += blockDim.x) {
        const int r = pos / ncols, x = pos - r * ncols;
        const uint32_t sh = 8u * (uint32_t)(x & 3);
        unsigned long long corr = 0, q = 0;      // a channel's sums fit 32 bits (tw * th <= 66000), the three together may not
        uint32_t s3[3];
        for (int c = 0; c < 3; c++) {
            const uint32_t *wrow = s_win + c * plane_smem + (size_t)r * ww_pitch + (x >> 2);
            const uint32_t *trow = s_tpl + (size_t)c * th * kw;
            uint32_t sc = 0, cc = 0, qc = 0;
            for (int v = 0; v < th; v++, wrow += ww_pitch, trow += kw) {
                uint32_t lo = wrow[0];
                for (int k = 0; k < kw; k++) {
                    const uint32_t hi = wrow[k + 1];
                    const uint32_t px = __funnelshift_r(lo, hi, sh);          // window bytes x + 4k .. x + 4k + 3 of this row
                    const uint32_t ones = s_ones[k];
                    cc = __dp4a(px, trow[k], cc);
                    sc = __dp4a(px, ones, sc);
                    qc = __dp4a(px, px & (ones * 255u), qc);
                    lo = hi;
                }
            }
            s3[c] = sc;
            corr += cc;
            q += qc;
        }
        const double ws[3] = {(double)s3[0], (double)s3[1], (double)s3[2]};
        const float score = match_score3(tc, (double)corr, ws, (double)q);
        const unsigned long long key = pack_key(score, (uint32_t)((y0 + r) * t.rw + x0 + x), want_min);
        mybest = key > mybest ? key : mybest;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, mybest, o);
        mybest = other > mybest ? other : mybest;
    }
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = mybest;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < RGB_THREADS / 32; w++) mybest = s_best[w] > mybest ? s_best[w] : mybest;
        atomicMax(&best[blockIdx.y], mybest);
    }
}

int launch_ncc_rgb(const DevTask *tasks, int n_tasks, int max_rh, int max_w, int max_tw, int max_th, const DevTemplates *tpls,
                   const uint8_t *win8, size_t plane_bytes, unsigned long long *best, cudaStream_t stream)
{
    const int max_rw = max_w - 1 + 1;                                  // an upper bound of the map width (templates are >= 1 wide)
    const int tile_cols = max_rw < RGB_TILE ? max_rw : RGB_TILE;
    const int ww_pitch = ((tile_cols + max_tw - 1 + 3) / 4 + 1) | 1;   // odd pitch in words: the rows of a band fall into different banks
    const int max_kw = (max_tw + 3) / 4;
    const size_t smem = 4 * (3 * (size_t)(RGB_BAND + max_th - 1) * ww_pitch + 3 * (size_t)max_th * max_kw + max_kw) + 16;
    if (smem > 200 * 1024 || (long)max_tw * max_th > 66000) return -1;
    if (cudaFuncSetAttribute(lst_ncc_rgb, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
    const int bands = (max_rh + RGB_BAND - 1) / RGB_BAND;
    const int ntiles = (max_rw + RGB_TILE - 1) / RGB_TILE;
    int launches = 0;
    for (int base = 0; base < n_tasks; base += 65535) {
        const int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        lst_ncc_rgb<<<dim3(bands * ntiles, n), RGB_THREADS, smem, stream>>>(tasks + base, tpls, win8, plane_bytes, best + base, ww_pitch, max_th,
                                                                            max_kw, ntiles);
        launches++;
    }
    return cudaPeekAtLastError() == cudaSuccess ? launches : -1;
}

// ------------------------------------------------------------------------------------------------
// JPEG slice files: reconstruction of the search regions on the device (SURVEY 8f-N2 "frame decode -> device")
// ------------------------------------------------------------------------------------------------
// One CTA per (region, slice).  Phase 1: a thread takes a block -- 64 quantised coefficients of the staging buffer, the
// slice's quantisation table -- and runs jpeg_idct_islow (the IJG integer IDCT, the same code the host decoder uses) into the
// region's sample planes in shared memory.  Phase 2: a thread takes an output pixel: luma sample, chroma through the fancy
// upsampling filters (the staged MCUs include the two pixels of context around the rectangle; the edge rules use the image's
// own downsampled size), YCbCr -> RGB, and writes the ColorProcessor pixel (or the grey sample) into the ROI ring.
__global__ void __launch_bounds__(256)
lst_jpeg_reconstruct(const int16_t *__restrict__ coef, const uint16_t *__restrict__ qt, const JpegDevBatch B, uint8_t *__restrict__ out)
{
    extern __shared__ __align__(16) uint8_t jp_smem[];
    const JpegDevRegion R = B.r[blockIdx.x];
    if (R.w <= 0 || R.h <= 0) return;
    const JpegGeom g = B.g;
    const int slice = blockIdx.y;
    const int mw = 8 * g.hmax, mh = 8 * g.vmax, bpm = jpeg_blocks_per_mcu(g), nluma = g.ncomp == 1 ? 1 : g.hmax * g.vmax;
    const int ys = R.m.nmx * mw, cs = R.m.nmx * 8;                          // row strides of the luma / chroma planes
    uint8_t *Yp = jp_smem;
    uint8_t *Cbp = Yp + (size_t)R.m.nmy * mh * ys;
    uint8_t *Crp = Cbp + (size_t)R.m.nmy * 8 * cs;
    const int16_t *cf = coef + R.coef_off + (size_t)slice * R.coef_slice;
    const uint16_t *q = qt + (size_t)slice * 192;
    const int nblocks = R.m.nmx * R.m.nmy * bpm;
    for (int blk = threadIdx.x; blk < nblocks; blk += blockDim.x) {
        const int mcu = blk / bpm, b = blk - mcu * bpm;
        const int mcy = mcu / R.m.nmx, mcx = mcu - mcy * R.m.nmx;
        if (b < nluma) {
            const int by = b / g.hmax, bx = b - by * g.hmax;
            jpeg_idct_islow(cf + (size_t)blk * 64, q, Yp + (size_t)((mcy * g.vmax + by) * 8) * ys + (mcx * g.hmax + bx) * 8, ys);
        } else {
            const int c = b - nluma + 1;
            jpeg_idct_islow(cf + (size_t)blk * 64, q + 64 * c, (c == 1 ? Cbp : Crp) + (size_t)(mcy * 8) * cs + mcx * 8, cs);
        }
    }
    __syncthreads();
    const int y_org = R.m.my0 * mh, x_org = R.m.mx0 * mw;                   // slice pixel of the planes' first sample
    const int cy_org = R.m.my0 * 8, cx_org = R.m.mx0 * 8;                   // chroma sample of the chroma planes' first sample
    const bool h2 = g.hmax == 2, v2 = g.vmax == 2;
    const int ds_w = jpeg_ds_w(g), ds_h = jpeg_ds_h(g);
    uint8_t *o = out + R.out_off + (size_t)slice * R.out_slice;
    for (int p = threadIdx.x; p < R.w * R.h; p += blockDim.x) {
        const int y = p / R.w, x = p - y * R.w;
        const int X = R.x + x, Y = R.y + y;
        const int c0 = Yp[(size_t)(Y - y_org) * ys + (X - x_org)];
        if (g.ncomp == 1) {
            o[p] = (uint8_t)c0;
            continue;
        }
        const int cy = v2 ? Y >> 1 : Y, fy = v2 ? jpeg_far_row(Y, ds_h) : Y;
        const uint8_t *bn = Cbp + (size_t)(cy - cy_org) * cs, *bf = Cbp + (size_t)(fy - cy_org) * cs;
        const uint8_t *rn = Crp + (size_t)(cy - cy_org) * cs, *rf = Crp + (size_t)(fy - cy_org) * cs;
        int c1, c2;
        if (!h2) c1 = bn[X - cx_org], c2 = rn[X - cx_org];
        else if (!v2) c1 = jpeg_up_h2v1(bn, cx_org, ds_w, X), c2 = jpeg_up_h2v1(rn, cx_org, ds_w, X);
        else c1 = jpeg_up_h2v2(bn, bf, cx_org, ds_w, X), c2 = jpeg_up_h2v2(rn, rf, cx_org, ds_w, X);
        ((uint32_t *)o)[p] = g.ycc ? jpeg_ycc_to_rgb(c0, c1, c2) : ((uint32_t)c0 << 16) | ((uint32_t)c1 << 8) | (uint32_t)c2;
    }
}

size_t jpeg_reconstruct_smem(const JpegGeom &g, const JpegRegion &m)
{
    const size_t luma = (size_t)m.nmy * 8 * g.vmax * m.nmx * 8 * g.hmax;
    return luma + (g.ncomp == 3 ? 2 * (size_t)m.nmy * 8 * m.nmx * 8 : 0) + 16;
}

int launch_jpeg_reconstruct(const int16_t *coef, const uint16_t *qt, const JpegDevBatch &batch, int n_slices, uint8_t *roi_out,
                            cudaStream_t stream)
{
    size_t smem = 0;
    for (int k = 0; k < 5; k++)
        if (batch.r[k].w > 0 && batch.r[k].h > 0) smem = std::max(smem, jpeg_reconstruct_smem(batch.g, batch.r[k].m));
    if (smem == 0 || n_slices <= 0) return 0;
    if (smem > 200 * 1024) return -1;
    if (cudaFuncSetAttribute(lst_jpeg_reconstruct, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
    int launches = 0;
    for (int base = 0; base < n_slices; base += 65535) {
        const int n = n_slices - base < 65535 ? n_slices - base : 65535;
        JpegDevBatch b = batch;
        for (int k = 0; k < 5; k++) {
            b.r[k].coef_off += (unsigned long long)base * b.r[k].coef_slice;
            b.r[k].out_off += (unsigned long long)base * b.r[k].out_slice;
        }
        lst_jpeg_reconstruct<<<dim3(5, n), 256, smem, stream>>>(coef, qt + (size_t)base * 192, b, roi_out);
        launches++;
    }
    return cudaPeekAtLastError() == cudaSuccess ? launches : -1;
}

// ------------------------------------------------------------------------------------------------
// per-window epilogue: 3x3 neighbourhood, sub-pixel fit, accept test
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288)
lst_epilogue(const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls, const uint8_t *__restrict__ win8,
             const unsigned long long *__restrict__ best, const float *__restrict__ nbhd, DevResult *__restrict__ results,
             MatchParams mp)
{
    // nine warps: warp w recomputes the score of neighbour (w%3-1, w/3-1) of the peak exactly
    // (correlation and window sums in integers, normalisation by ncc_score as in the match kernel) --
    // unless the match kernel already left the nine scores (lst_ncc_tc, window owned by one CTA)
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
    const bool inside = nx >= 0 && nx < t.rw && ny >= 0 && ny < t.rh;
    const bool have = nbhd[12 * (size_t)blockIdx.x + 9] != 0.0f;
    if (have) {
        if (inside) sc = nbhd[12 * (size_t)blockIdx.x + warp];
    } else if (inside && tc.channels == 3) {
        unsigned long long acc = 0, q = 0;
        uint32_t s3[3] = {0, 0, 0};
        for (int c = 0; c < 3; c++) {
            const uint8_t *w = win8 + (size_t)c * mp.plane_bytes + t.win_off + (size_t)ny * t.pitch + nx;
            const uint8_t *tp = tpls->raw[t.tpl] + (size_t)c * tc.tw * tc.th;
            uint32_t s = 0, a = 0, qq = 0;
            for (int e = lane; e < tc.tw * tc.th; e += 32) {
                int v = e / tc.tw, u = e - v * tc.tw;
                uint32_t px = w[(size_t)v * t.pitch + u];
                a += px * (uint32_t)tp[e];
                s += px;
                qq += px * px;
            }
            s3[c] = s;
            acc += a;
            q += qq;
        }
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
            for (int c = 0; c < 3; c++) s3[c] += __shfl_xor_sync(0xffffffffu, s3[c], o);
        }
        const double ws[3] = {(double)s3[0], (double)s3[1], (double)s3[2]};
        sc = match_score3(tc, (double)acc, ws, (double)q);
    } else if (inside) {
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
                    const unsigned long long *best, const float *nbhd, DevResult *results, const MatchParams &mp,
                    cudaStream_t stream)
{
    lst_epilogue<<<n_tasks, 288, 0, stream>>>(tasks, tpls, win8, best, nbhd, results, mp);
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
    // function attributes belong to the current device: this runs once per lst_create.  A failure is reported (a kernel that
    // was refused its shared memory would fail at launch with a less telling message) and does not stop the others.
    auto set = [](const void *fn, cudaFuncAttribute attr, int value, const char *what) {
        const cudaError_t e = cudaFuncSetAttribute(fn, attr, value);
        if (e != cudaSuccess) {
            fprintf(stderr, "lst_b200: cudaFuncSetAttribute(%s, %s) failed: %s\n", what,
                    attr == cudaFuncAttributeMaxDynamicSharedMemorySize ? "MaxDynamicSharedMemorySize" : "PreferredSharedMemoryCarveout",
                    cudaGetErrorString(e));
            cudaGetLastError();
        }
    };
    const cudaFuncAttribute dyn = cudaFuncAttributeMaxDynamicSharedMemorySize, carve = cudaFuncAttributePreferredSharedMemoryCarveout;
    const int mx = (int)cudaSharedmemCarveoutMaxShared;
#define LST_SET_BOTH(fn, bytes)                     \
    set((const void *)(fn), dyn, (bytes), #fn);     \
    set((const void *)(fn), carve, mx, #fn)
#define LST_SET_CARVE(fn) set((const void *)(fn), carve, mx, #fn)
    LST_SET_BOTH(lst_preprocess<LST_PIX_U8>, 200 * 1024);
    LST_SET_BOTH(lst_preprocess<LST_PIX_U16>, 200 * 1024);
    LST_SET_BOTH(lst_preprocess<LST_PIX_F32>, 200 * 1024);
    LST_SET_BOTH(lst_preprocess<LST_PIX_RGB>, 200 * 1024);
    LST_SET_BOTH(lst_preprocess_win<LST_PIX_U8>, 200 * 1024);
    LST_SET_BOTH(lst_preprocess_win<LST_PIX_U16>, 200 * 1024);
    LST_SET_BOTH(lst_preprocess_win<LST_PIX_F32>, 200 * 1024);
    LST_SET_BOTH(lst_ncc_match<false>, 220 * 1024);
    LST_SET_BOTH(lst_ncc_match<true>, 220 * 1024);
    LST_SET_CARVE(lst_init_pass);
    LST_SET_CARVE(lst_upload);
    LST_SET_CARVE(lst_minmax<LST_PIX_U8>);
    LST_SET_CARVE(lst_minmax<LST_PIX_U16>);
    LST_SET_CARVE(lst_minmax<LST_PIX_F32>);
    LST_SET_CARVE(lst_minmax<LST_PIX_RGB>);
    LST_SET_CARVE(lst_epilogue);
    LST_SET_CARVE(lst_solve);
    LST_SET_CARVE(lst_gather_rois);
#undef LST_SET_BOTH
#undef LST_SET_CARVE
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
