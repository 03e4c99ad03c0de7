// lst_kernels.cu -- hand-written sm_100a kernels of the tracking hot path.
//
//   lst_minmax       per-window min/max of the raw crop (FloatProcessor.findMinAndMax of the crop)
//   lst_preprocess   crop -> float -> ImageJ GaussianBlur(2,2,0.02) -> 8-bit   (LST4:1346-1347, 1456-1473, 2529)
//   lst_box_sums     window sums of I and I^2 under the template footprint    (cv::integral in matchTemplate)
//   lst_ncc_match    exact integer cross-correlation (IDP.4A) + TM_CCOEFF_NORMED normalisation + argmax
//                                                                              (LST4:2560 matchTemplate, 2670 minMaxLoc)
//   lst_epilogue     3x3 sub-pixel fit + accept test per window                (LST4:2681-2717, 1225-1255)
//   lst_solve        calcDisplacement per frame                                (LST4:2318-2365)
//
// Compiled with -fmad=false: every float/double operation rounds once, like the Java and OpenCV
// code it reproduces.  All heavy arithmetic is integer (IDP.4A.U8.U8), which is exact.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lst_b200.h"
#include "lst_device.h"

namespace lst {

#define KR 7   // kRadius of the sigma=2 / accuracy=0.02 Gaussian (13 taps)

__constant__ float c_kern[KR];
__constant__ float c_ksum[KR];

int set_blur_kernel(const float kern[7], const float ksum[7])
{
    if (cudaMemcpyToSymbol(c_kern, kern, sizeof(float) * KR) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_ksum, ksum, sizeof(float) * KR) != cudaSuccess) return -1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
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
__global__ void lst_minmax(const DevTask *__restrict__ tasks, const void *const *__restrict__ frames, int frame_w,
                           uint32_t *__restrict__ mm_keys)
{
    const DevTask t = tasks[blockIdx.y];
    const void *fr = frames[t.frame];
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int r = blockIdx.x; r < t.h; r += gridDim.x) {
        size_t row = (size_t)(t.y0 + r) * frame_w + t.x0;
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

int launch_minmax(const DevTask *tasks, int n_tasks, int max_h, const void *const *frames, const PreprocParams &pp,
                  uint32_t *mm_keys, cudaStream_t stream)
{
    int launches = 0;
    int gx = max_h < 8 ? 1 : (max_h + 31) / 32;   // 32 rows per CTA
    if (gx > 256) gx = 256;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 grid(gx, n);
        if (pp.pixel_type == LST_PIX_U8)
            lst_minmax<LST_PIX_U8><<<grid, 128, 0, stream>>>(tasks + base, frames, pp.frame_w, mm_keys + 2 * base);
        else if (pp.pixel_type == LST_PIX_U16)
            lst_minmax<LST_PIX_U16><<<grid, 128, 0, stream>>>(tasks + base, frames, pp.frame_w, mm_keys + 2 * base);
        else
            lst_minmax<LST_PIX_F32><<<grid, 128, 0, stream>>>(tasks + base, frames, pp.frame_w, mm_keys + 2 * base);
        launches++;
    }
    return launches;
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

template <int PIX>
__global__ void __launch_bounds__(256)
lst_preprocess(const DevTask *__restrict__ tasks, const void *const *__restrict__ frames, PreprocParams pp,
               const uint32_t *__restrict__ mm_keys, uint8_t *__restrict__ win8)
{
    extern __shared__ float smem_f[];
    const DevTask t = tasks[blockIdx.y];
    const int TW = pp.tile_w, TH = pp.tile_h;
    const int tiles_x = (t.w + TW - 1) / TW, tiles_y = (t.h + TH - 1) / TH;
    if ((int)blockIdx.x >= tiles_x * tiles_y) return;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int cx0 = tx * TW, cy0 = ty * TH;
    const int cx1 = min(cx0 + TW, t.w), cy1 = min(cy0 + TH, t.h);
    const int rx0 = max(cx0 - (KR - 1), 0), rx1 = min(cx1 + (KR - 1), t.w);
    const int ry0 = max(cy0 - (KR - 1), 0), ry1 = min(cy1 + (KR - 1), t.h);
    const int RWP = TW + 2 * (KR - 1) + 1;     // raw tile pitch (floats)
    const int XWP = TW + 1;                    // x-pass tile pitch
    float *raw = smem_f;                                   // [(TH+12)][RWP]
    float *xp = smem_f + (size_t)(TH + 2 * (KR - 1)) * RWP;   // [(TH+12)][XWP]
    const void *fr = frames[t.frame];
    const int nrw = rx1 - rx0, nrh = ry1 - ry0;

    for (int idx = threadIdx.x; idx < nrw * nrh; idx += blockDim.x) {
        int r = idx / nrw, c = idx - r * nrw;
        size_t src = (size_t)(t.y0 + ry0 + r) * pp.frame_w + (t.x0 + rx0 + c);
        raw[r * RWP + c] = load_px<PIX>(fr, src);
    }
    float lo = pp.lo, hi = pp.hi;
    if (pp.range_per_task) {
        lo = key_to_float<PIX>(mm_keys[2 * blockIdx.y]);
        hi = key_to_float<PIX>(mm_keys[2 * blockIdx.y + 1]);
    }
    __syncthreads();

    // X pass: rows [ry0,ry1), columns [cx0,cx1)
    const int ncw = cx1 - cx0;
    for (int idx = threadIdx.x; idx < ncw * nrh; idx += blockDim.x) {
        int r = idx / ncw, c = idx - r * ncw;
        const float *line = raw + r * RWP - rx0;   // line[i] = crop row pixel i
        int i = cx0 + c;
        float res;
#define IN_X(ii) line[(ii)]
        LST_BLUR_AT(res, i, t.w, IN_X)
#undef IN_X
        xp[r * XWP + c] = res;
    }
    __syncthreads();

    // Y pass + quantise: rows [cy0,cy1), columns [cx0,cx1)
    float scale = (pp.quant_mode == LST_QUANT_ROUND255 ? 255.0f : 256.0f) / (hi - lo);
    const int nch = cy1 - cy0;
    uint8_t *dst = win8 + t.win_off;
    for (int idx = threadIdx.x; idx < ncw * nch; idx += blockDim.x) {
        int r = idx / ncw, c = idx - r * ncw;
        const float *col = xp + c - ry0 * XWP;     // col[i*XWP] = x-pass value at crop row i
        int i = cy0 + r;
        float res;
#define IN_Y(ii) col[(ii)*XWP]
        LST_BLUR_AT(res, i, t.h, IN_Y)
#undef IN_Y
        float value = res - lo;
        if (value < 0.0f) value = 0.0f;
        float q = value * scale;
        if (pp.quant_mode == LST_QUANT_ROUND255) q = q + 0.5f;
        int iv = __float2int_rz(q);            // Java (int): NaN -> 0, saturating
        iv = iv > 255 ? 255 : (iv < 0 ? 0 : iv);
        dst[(size_t)i * t.pitch + (cx0 + c)] = (uint8_t)iv;
    }
    // zero the pitch padding so the match kernel can load whole 16-byte words
    if (cx1 == t.w && t.pitch > t.w) {
        int padw = t.pitch - t.w;
        for (int idx = threadIdx.x; idx < padw * nch; idx += blockDim.x) {
            int r = idx / padw, c = idx - r * padw;
            dst[(size_t)(cy0 + r) * t.pitch + t.w + c] = 0;
        }
    }
}

static size_t preprocess_smem(const PreprocParams &pp)
{
    size_t rows = pp.tile_h + 2 * (KR - 1);
    return rows * (size_t)(pp.tile_w + 2 * (KR - 1) + 1 + pp.tile_w + 1) * sizeof(float);
}

int launch_preprocess(const DevTask *tasks, int n_tasks, int max_w, int max_h, const void *const *frames,
                      const PreprocParams &pp, const uint32_t *mm_keys, uint8_t *win8, cudaStream_t stream)
{
    int tiles = ((max_w + pp.tile_w - 1) / pp.tile_w) * ((max_h + pp.tile_h - 1) / pp.tile_h);
    size_t smem = preprocess_smem(pp);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(lst_preprocess<LST_PIX_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(lst_preprocess<LST_PIX_U16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(lst_preprocess<LST_PIX_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    int launches = 0;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 grid(tiles, n);
        const uint32_t *mm = mm_keys + 2 * base;
        if (pp.pixel_type == LST_PIX_U8)
            lst_preprocess<LST_PIX_U8><<<grid, 256, smem, stream>>>(tasks + base, frames, pp, mm, win8);
        else if (pp.pixel_type == LST_PIX_U16)
            lst_preprocess<LST_PIX_U16><<<grid, 256, smem, stream>>>(tasks + base, frames, pp, mm, win8);
        else
            lst_preprocess<LST_PIX_F32><<<grid, 256, smem, stream>>>(tasks + base, frames, pp, mm, win8);
        launches++;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------------
// window sums under the template footprint
// ------------------------------------------------------------------------------------------------
#define BOX_W 32
#define BOX_H 32

__global__ void __launch_bounds__(256)
lst_box_sums(const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls, const uint8_t *__restrict__ win8,
             uint32_t *__restrict__ wsum, uint32_t *__restrict__ wsq, int max_tw, int max_th)
{
    extern __shared__ uint32_t smem_u[];
    const DevTask t = tasks[blockIdx.y];
    const int tw = tpls->c[t.tpl].tw, th = tpls->c[t.tpl].th;
    const int tiles_x = (t.rw + BOX_W - 1) / BOX_W, tiles_y = (t.rh + BOX_H - 1) / BOX_H;
    if ((int)blockIdx.x >= tiles_x * tiles_y) return;
    const int bx = (blockIdx.x % tiles_x) * BOX_W, by = (blockIdx.x / tiles_x) * BOX_H;
    const int ow = min(BOX_W, t.rw - bx), oh = min(BOX_H, t.rh - by);   // outputs in this tile
    const int cw = ow + tw - 1, ch = oh + th - 1;                       // window pixels needed
    const int CP = BOX_W + max_tw - 1;                                  // pitch of V / V2 (words)
    const int TP = (BOX_W + max_tw - 1 + 3) & ~3;                       // pitch of the byte tile
    uint32_t *V = smem_u;                                               // [BOX_H][CP]
    uint32_t *V2 = V + BOX_H * CP;
    uint8_t *tile = (uint8_t *)(V2 + BOX_H * CP);                       // [(BOX_H+max_th-1)][TP]
    const uint8_t *src = win8 + t.win_off + (size_t)by * t.pitch + bx;
    for (int idx = threadIdx.x; idx < cw * ch; idx += blockDim.x) {
        int r = idx / cw, c = idx - r * cw;
        tile[r * TP + c] = src[(size_t)r * t.pitch + c];
    }
    __syncthreads();
    // vertical sliding sums: V[y][c] = sum_{v<th} tile[y+v][c]
    for (int c = threadIdx.x; c < cw; c += blockDim.x) {
        uint32_t s = 0, q = 0;
        for (int v = 0; v < th; v++) {
            uint32_t p = tile[v * TP + c];
            s += p;
            q += p * p;
        }
        V[c] = s;
        V2[c] = q;
        for (int y = 1; y < oh; y++) {
            uint32_t a = tile[(y + th - 1) * TP + c], b = tile[(y - 1) * TP + c];
            s += a - b;
            q += a * a - b * b;
            V[y * CP + c] = s;
            V2[y * CP + c] = q;
        }
    }
    __syncthreads();
    // horizontal: out[y][x] = sum_{u<tw} V[y][x+u]; each thread does a run of 8 outputs
    const int segs = (ow + 7) / 8;
    uint32_t *osum = wsum + t.sum_off, *osq = wsq + t.sum_off;
    for (int idx = threadIdx.x; idx < oh * segs; idx += blockDim.x) {
        int y = idx / segs, x0 = (idx - y * segs) * 8;
        int x1 = min(x0 + 8, ow);
        const uint32_t *vr = V + y * CP, *vr2 = V2 + y * CP;
        uint32_t s = 0, q = 0;
        for (int u = 0; u < tw; u++) {
            s += vr[x0 + u];
            q += vr2[x0 + u];
        }
        size_t o = (size_t)(by + y) * t.rw + bx;
        osum[o + x0] = s;
        osq[o + x0] = q;
        for (int x = x0 + 1; x < x1; x++) {
            s += vr[x + tw - 1] - vr[x - 1];
            q += vr2[x + tw - 1] - vr2[x - 1];
            osum[o + x] = s;
            osq[o + x] = q;
        }
    }
}

static size_t box_smem(int tw, int th)
{
    size_t cp = BOX_W + tw - 1, tp = (BOX_W + tw - 1 + 3) & ~3;
    return 2 * BOX_H * cp * sizeof(uint32_t) + (size_t)(BOX_H + th - 1) * tp;
}

int launch_box_sums(const DevTask *tasks, int n_tasks, int max_rw, int max_rh, int tw, int th,
                    const DevTemplates *tpls, const uint8_t *win8, uint32_t *wsum, uint32_t *wsq, cudaStream_t stream)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(lst_box_sums, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    int tiles = ((max_rw + BOX_W - 1) / BOX_W) * ((max_rh + BOX_H - 1) / BOX_H);
    size_t smem = box_smem(tw, th);
    int launches = 0;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 grid(tiles, n);
        lst_box_sums<<<grid, 256, smem, stream>>>(tasks + base, tpls, win8, wsum, wsq, tw, th);
        launches++;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------------
// exact NCC + argmax
// ------------------------------------------------------------------------------------------------
// Result position x = 4q + j.  With the template stored four times, shifted by j bytes and
// zero padded (build_shifted_template), the correlation of one template row is
//     sum_k IDP.4A(Wword[q + k], Tj[k])
// over aligned 32-bit words of the window row: no unaligned loads and no byte shuffles.
// A thread owns 8 consecutive q (32 result columns) of one result row; lanes of a warp own
// consecutive rows, so template words are warp-uniform (shared-memory broadcast) and window
// rows are conflict-free because the tile pitch is an odd multiple of 16 bytes.
__device__ __forceinline__ unsigned long long pack_key(float score, uint32_t idx)
{
    score = score + 0.0f;   // -0 -> +0: minMaxLoc compares values, not bit patterns
    return ((unsigned long long)f32_key(score) << 32) | (unsigned long long)(0xffffffffu - idx);
}

#define NCC_THREADS 256

__global__ void __launch_bounds__(NCC_THREADS)
lst_ncc_match(const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls,
              const uint8_t *__restrict__ win8, const uint32_t *__restrict__ wsum, const uint32_t *__restrict__ wsq,
              unsigned long long *__restrict__ best, float *__restrict__ score_map, MatchParams mp)
{
    extern __shared__ __align__(16) uint8_t smem_b[];
    __shared__ unsigned long long s_best[NCC_THREADS / 32];
    const DevTask t = tasks[blockIdx.y];
    const TplConst tc = tpls->c[t.tpl];
    const int tw = tc.tw, th = tc.th, k4 = tc.k4;
    const int tiles_x = (t.rw + 31) / 32;
    const int chunks = (tiles_x + mp.tiles_per_cta - 1) / mp.tiles_per_cta;
    const int bands = (t.rh + mp.band_rows - 1) / mp.band_rows;
    if ((int)blockIdx.x >= chunks * bands) return;
    const int chunk = blockIdx.x % chunks, band = blockIdx.x / chunks;
    const int tile0 = chunk * mp.tiles_per_cta;
    const int ntiles = min(mp.tiles_per_cta, tiles_x - tile0);
    const int y0 = band * mp.band_rows;
    const int nrows = min(mp.band_rows, t.rh - y0);
    const int SP = mp.smem_pitch;

    // shared memory: the window tile [(nrows+th-1)][SP].  The shifted template ([th][k4] uint4) is read
    // with warp-uniform read-only loads and lives in L1: it costs no shared memory, so large
    // templates (128x128 -> 67 KB) do not squeeze the window tile.
    const uint4 *__restrict__ g_tpl = (const uint4 *)tpls->shifted[t.tpl];
    uint8_t *s_win = smem_b;
    {
        const int wrows = nrows + th - 1;
        const int vec = SP / 16;
        const int xb0 = tile0 * 32;                      // first window byte column of this chunk
        const uint8_t *src = win8 + t.win_off;
        for (int i = threadIdx.x; i < wrows * vec; i += blockDim.x) {
            int r = i / vec, c = (i - r * vec) * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            int gx = xb0 + c, gy = y0 + r;
            if (gx < t.pitch && gy < t.h) v = *(const uint4 *)(src + (size_t)gy * t.pitch + gx);
            *(uint4 *)(s_win + (size_t)r * SP + c) = v;
        }
    }
    __syncthreads();

    unsigned long long mybest = 0ull;
    const int ntask = ntiles * nrows;
    const int k4main = k4 & ~3;
    for (int tt = threadIdx.x; tt < ntask; tt += blockDim.x) {
        const int tile = tt / nrows, yy = tt - tile * nrows;
        uint32_t acc[4][8];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int q = 0; q < 8; q++) acc[j][q] = 0u;
        const uint8_t *wbase = s_win + (size_t)yy * SP + tile * 32;
        for (int v = 0; v < th; v++) {
            const uint8_t *wrow = wbase + (size_t)v * SP;
            const uint4 *trow = g_tpl + v * k4;
            int k = 0;
            for (; k < k4main; k += 4) {
                const uint4 wa = *(const uint4 *)(wrow + 4 * k);
                const uint4 wb = *(const uint4 *)(wrow + 4 * k + 16);
                const uint4 wc = *(const uint4 *)(wrow + 4 * k + 32);
                const uint32_t ww[12] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w};
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    const uint4 tj = __ldg(trow + k + kk);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t wv = ww[kk + q];
                        acc[0][q] = __dp4a(wv, tj.x, acc[0][q]);
                        acc[1][q] = __dp4a(wv, tj.y, acc[1][q]);
                        acc[2][q] = __dp4a(wv, tj.z, acc[2][q]);
                        acc[3][q] = __dp4a(wv, tj.w, acc[3][q]);
                    }
                }
            }
            for (; k < k4; k++) {
                const uint4 tj = __ldg(trow + k);
                uint32_t ww[8];
#pragma unroll
                for (int q = 0; q < 8; q++) ww[q] = *(const uint32_t *)(wrow + 4 * (k + q));
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    acc[0][q] = __dp4a(ww[q], tj.x, acc[0][q]);
                    acc[1][q] = __dp4a(ww[q], tj.y, acc[1][q]);
                    acc[2][q] = __dp4a(ww[q], tj.z, acc[2][q]);
                    acc[3][q] = __dp4a(ww[q], tj.w, acc[3][q]);
                }
            }
        }
        // normalise (double, OpenCV order) and keep the first maximum
        const int y = y0 + yy;
        const int xbase = (tile0 + tile) * 32;
        const size_t rowoff = t.sum_off + (size_t)y * t.rw;
#pragma unroll
        for (int q = 0; q < 8; q++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int x = xbase + 4 * q + j;
                if (x < t.rw) {
                    const uint32_t s = wsum[rowoff + x], sq = wsq[rowoff + x];
                    const float sc = ncc_score((double)acc[j][q], (double)s, (double)sq, tc.inv_area, tc.mean, tc.norm, tc.flat);
                    if (score_map) score_map[rowoff + x] = sc;
                    const unsigned long long key = pack_key(sc, (uint32_t)(y * t.rw + x));
                    mybest = key > mybest ? key : mybest;
                }
            }
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

void plan_match(int max_rw, int max_rh, int tw, int th, int n_tasks, MatchParams *mp, size_t *smem_bytes)
{
    const int k4 = (tw + 2) / 4 + 1;
    const int tiles_x = (max_rw + 31) / 32;
    const size_t budget = 100 * 1024;   // <= ~100 KB per CTA keeps two CTAs resident per SM
    auto pitch_of = [&](int tiles) {
        int p = 32 * tiles + 4 * k4;
        p = (p + 15) & ~15;
        if (((p / 16) & 1) == 0) p += 16;   // odd multiple of 16 B: conflict-free 128-bit row loads
        return p;
    };
    auto smem_of = [&](int tiles, int rows) { return (size_t)(rows + th - 1) * pitch_of(tiles); };
    auto ctas_of = [&](int tiles, int rows) {
        return (long)n_tasks * ((tiles_x + tiles - 1) / tiles) * ((max_rh + rows - 1) / rows);
    };
    // Start from one CTA per window; split rows (in multiples of 32: lanes own rows) and then
    // column tiles until the tile fits and there are enough CTAs to fill 148 SMs a few times.
    int tiles = tiles_x, rows = max_rh;
    const long want_ctas = 148L * 4;
    for (;;) {
        bool too_big = smem_of(tiles, rows) > budget;
        bool too_few = ctas_of(tiles, rows) < want_ctas && (long)tiles * rows > 2 * NCC_THREADS;
        if (!too_big && !too_few) break;
        if (rows > 32) rows = (((rows + 1) / 2) + 31) & ~31;
        else if (tiles > 1) tiles = (tiles + 1) / 2;
        else break;
    }
    mp->band_rows = rows;
    mp->tiles_per_cta = tiles;
    mp->smem_pitch = pitch_of(tiles);
    *smem_bytes = smem_of(tiles, rows);
}

int launch_ncc_match(const DevTask *tasks, int n_tasks, int max_rw, int max_rh, int tw, int th,
                     const DevTemplates *tpls, const uint8_t *win8, const uint32_t *wsum, const uint32_t *wsq,
                     unsigned long long *best, float *score_map, const MatchParams &mp, cudaStream_t stream)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(lst_ncc_match, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_set = true;
    }
    const int tiles_x = (max_rw + 31) / 32;
    const int chunks = (tiles_x + mp.tiles_per_cta - 1) / mp.tiles_per_cta;
    const int bands = (max_rh + mp.band_rows - 1) / mp.band_rows;
    size_t smem = (size_t)(mp.band_rows + th - 1) * mp.smem_pitch;
    (void)tw;
    int launches = 0;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 grid(chunks * bands, n);
        lst_ncc_match<<<grid, NCC_THREADS, smem, stream>>>(tasks + base, tpls, win8, wsum, wsq, best + base, score_map, mp);
        launches++;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------------
// per-window epilogue: 3x3 neighbourhood, sub-pixel fit, accept test
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288)
lst_epilogue(const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls, const uint8_t *__restrict__ win8,
             const uint32_t *__restrict__ wsum, const uint32_t *__restrict__ wsq,
             const unsigned long long *__restrict__ best, DevResult *__restrict__ results, MatchParams mp)
{
    __shared__ float nb[9];
    const DevTask t = tasks[blockIdx.x];
    const TplConst tc = tpls->c[t.tpl];
    const unsigned long long key = best[blockIdx.x];
    const uint32_t idx = 0xffffffffu - (uint32_t)(key & 0xffffffffull);
    const int ix = (int)(idx % (uint32_t)t.rw), iy = (int)(idx / (uint32_t)t.rw);
    const float peak = key_f32((uint32_t)(key >> 32));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nx = ix + (warp % 3) - 1, ny = iy + (warp / 3) - 1;
    float sc = 0.0f;
    if (nx >= 0 && nx < t.rw && ny >= 0 && ny < t.rh) {
        const uint8_t *w = win8 + t.win_off + (size_t)ny * t.pitch + nx;
        const uint8_t *tp = tpls->raw[t.tpl];
        uint32_t acc = 0;
        for (int e = lane; e < tc.tw * tc.th; e += 32) {
            int v = e / tc.tw, u = e - v * tc.tw;
            acc += (uint32_t)w[(size_t)v * t.pitch + u] * (uint32_t)tp[e];
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        size_t o = t.sum_off + (size_t)ny * t.rw + nx;
        sc = ncc_score((double)acc, (double)wsum[o], (double)wsq[o], tc.inv_area, tc.mean, tc.norm, tc.flat);
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
        r.accepted = test_match_result(r.score, tc.mideal, mp.thresh, r.x, r.y, 2 * t.s_used, mp.templ_size);
        for (int i = 0; i < 9; i++) r.nb[i] = nb[i];
        r.pad = 0.0f;
        results[blockIdx.x] = r;
    }
}

int launch_epilogue(const DevTask *tasks, int n_tasks, const DevTemplates *tpls, const uint8_t *win8,
                    const uint32_t *wsum, const uint32_t *wsq, const unsigned long long *best, DevResult *results,
                    const MatchParams &mp, cudaStream_t stream)
{
    lst_epilogue<<<n_tasks, 288, 0, stream>>>(tasks, tpls, win8, wsum, wsq, best, results, mp);
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

int measure_alu_peaks(double ms_per_test, double out[2])
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
    return 0;
}

}  // namespace lst
