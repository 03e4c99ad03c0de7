// lst_preprocess_win.cuh -- crop -> float -> ImageJ GaussianBlur(2,2,0.02) -> 8-bit, second generation
// (LST4:1346-1347, 1456-1473, 2529); included by lst_kernels.cu (shares c_kern / c_ksum / blur_run).
//
// Same arithmetic as lst_preprocess (blur_run: Java float order, one rounding per operation), different data
// movement -- the first kernel spent two thirds of its instructions outside the blur:
//   * the whole raw crop of a window is brought into shared memory once, as raw pixels (no float copy): 128-bit
//     loads realigned with funnel shifts whose word selection is a branch taken once per row, the crop's min/max
//     (the range convention) folded into the same pass with packed min/max instructions.  (TMA cannot
//     do this staging: cp.async.bulk.tensor faults unless the box starts on a 16-byte boundary of the global
//     tensor, and a crop starts at any pixel -- measured with tools/scratch probes, see profiles/README.md);
//   * X pass: a warp owns a run of 16 columns (8 for float stacks), its lanes 32 consecutive rows; the raw
//     pixels are unpacked in registers (PRMT + FADD of 2^23: exact, on the FP32 pipe);
//   * Y pass: a thread owns 4 adjacent columns x 4 rows (sixteen 128-bit loads for 16 pixels) and writes each
//     row of four 8-bit results as one 32-bit word;
//   * rows are scheduled in rounds of 32 X-pass rows, so every round has the same, full, amount of work.
// Windows wider than 240 pixels or taller than 256 and RGB stacks stay with lst_preprocess.
#pragma once

namespace lst {

#define PT_MAXTHREADS 320                 // 2 CTAs per SM with ~96 registers each
// X-pass rows per round (64 when two CTAs of that size fit an SM, else 32); the ring holds a round plus the 2 (KR-1)
// rows two rounds share, plus 4: a partial group of four output rows reads (and discards) up to three rows more
__host__ __device__ inline int pt_ring_rows(int round_rows) { return round_rows + 2 * (KR - 1) + 4; }

__device__ __forceinline__ uint32_t pt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int PIX>
struct PtPix {
    static constexpr int ES = PIX == LST_PIX_U8 ? 1 : (PIX == LST_PIX_U16 ? 2 : 4);
    static constexpr int XRUN = PIX == LST_PIX_F32 ? 8 : 16;   // outputs per X-pass item
    static constexpr int NLOAD = XRUN + 16;                    // pixels an item loads: columns i0-8 .. i0+XRUN+7
};

__device__ __forceinline__ unsigned long long pt_pack(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void pt_unpack2(unsigned long long v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long pt_mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long pt_add2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// 16 raw bytes -> floats (exact): the pixel bytes go into the mantissa of 2^23, which is then subtracted
template <int PIX>
__device__ __forceinline__ void pt_unpack(const uint4 q, bool swap, float *out)
{
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    if (PIX == LST_PIX_U8) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            out[4 * j + 0] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7440)) - 8388608.0f;
            out[4 * j + 1] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7441)) - 8388608.0f;
            out[4 * j + 2] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7442)) - 8388608.0f;
            out[4 * j + 3] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7443)) - 8388608.0f;
        }
    } else if (PIX == LST_PIX_U16) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            // big-endian samples (TIFF "MM"): the two bytes of a sample change places
            out[2 * j + 0] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, swap ? 0x7401 : 0x7410)) - 8388608.0f;
            out[2 * j + 1] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, swap ? 0x7423 : 0x7432)) - 8388608.0f;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) out[j] = __uint_as_float(swap ? __byte_perm(w[j], 0, 0x0123) : w[j]);
    }
}

// one raw pixel of the crop as float (edge terms of the blur)
template <int PIX>
__device__ __forceinline__ float pt_pixel(const uint8_t *row, int col, bool swap)
{
    if (PIX == LST_PIX_U8) return (float)row[col];
    if (PIX == LST_PIX_U16) {
        uint32_t v = ((const uint16_t *)row)[col];
        if (swap) v = __byte_perm(v, 0, 0x3201);
        return (float)v;
    }
    uint32_t v = ((const uint32_t *)row)[col];
    if (swap) v = __byte_perm(v, 0, 0x0123);
    return __uint_as_float(v);
}

// The 16 bytes [a, a + 16) of the 32 bytes {v0, v1}; sh = 8 * (a & 3)
__device__ __forceinline__ void pw_realign(const uint4 v0, const uint4 v1, uint32_t a, uint32_t (&o)[4])
{
    const uint32_t sh = 8u * (a & 3u);
    switch (a >> 2) {        // the same for every vector of a row (and of the window when the pitch is a multiple of 16 bytes)
    case 0:
        o[0] = __funnelshift_r(v0.x, v0.y, sh), o[1] = __funnelshift_r(v0.y, v0.z, sh);
        o[2] = __funnelshift_r(v0.z, v0.w, sh), o[3] = __funnelshift_r(v0.w, v1.x, sh);
        break;
    case 1:
        o[0] = __funnelshift_r(v0.y, v0.z, sh), o[1] = __funnelshift_r(v0.z, v0.w, sh);
        o[2] = __funnelshift_r(v0.w, v1.x, sh), o[3] = __funnelshift_r(v1.x, v1.y, sh);
        break;
    case 2:
        o[0] = __funnelshift_r(v0.z, v0.w, sh), o[1] = __funnelshift_r(v0.w, v1.x, sh);
        o[2] = __funnelshift_r(v1.x, v1.y, sh), o[3] = __funnelshift_r(v1.y, v1.z, sh);
        break;
    default:
        o[0] = __funnelshift_r(v0.w, v1.x, sh), o[1] = __funnelshift_r(v1.x, v1.y, sh);
        o[2] = __funnelshift_r(v1.y, v1.z, sh), o[3] = __funnelshift_r(v1.z, v1.w, sh);
        break;
    }
}

// Shared memory: raw[h][bw] pixels (crop columns -8 .. bw-9, zero outside the crop; row pitch an odd multiple of
// 16 bytes) | xp[pt_ring_rows(round_rows)][xwp] floats (X-pass rows; xwp/4 odd).
template <int PIX>
__global__ void __launch_bounds__(PT_MAXTHREADS, 2)
lst_preprocess_win(const DevTask *__restrict__ tasks, const Surface *__restrict__ surfaces, PreprocParams pp, int bw, int bh, int xwp,
                   int round_rows, uint8_t *__restrict__ win8)
{
    constexpr int ES = PtPix<PIX>::ES, XRUN = PtPix<PIX>::XRUN, NLOAD = PtPix<PIX>::NLOAD;
    extern __shared__ __align__(128) uint8_t pt_smem[];
    __shared__ uint32_t s_mm[2];
    const DevTask t = tasks[blockIdx.x];
    if (t.pad0) return;                        // shares the 8-bit window of an earlier task (whole-slice search)
    const int w = t.w, h = t.h;
    const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int rp = bw * ES;                                   // raw row pitch in bytes
    uint8_t *raw = pt_smem;
    float *xp = (float *)(pt_smem + (((size_t)bh * rp + 127) & ~(size_t)127));
    const bool swap = pp.swap_bytes != 0;
    if (tid == 0) {
        s_mm[0] = 0xffffffffu;
        s_mm[1] = 0u;
    }

    // ---- the raw crop -> shared memory (and its min / max on the way) ----
    // Vector q of row r holds crop bytes [16q - 8 ES, 16q - 8 ES + 16) of that row.  It is fetched as the (at most) two
    // aligned 128-bit words around it -- an aligned word that overlaps the crop row cannot leave the allocation -- and
    // realigned with funnel shifts; bytes outside the crop row become 0 (the blur adds the replicated end pixels through
    // the kernel-tail sums).  Items (r, q) go to the threads in linear order without a division per item.
    uint32_t klo = 0xffffffffu, khi = 0u;
    {
        const Surface sf = surfaces[t.frame];
        const size_t pitch_b = (size_t)sf.pitch_px * ES;
        const uint8_t *crop00 = (const uint8_t *)sf.ptr + ((size_t)(t.y0 - sf.org_y) * sf.pitch_px + (t.x0 - sf.org_x)) * ES;
        const int row_bytes = w * ES;
        const int nv = rp >> 4;
        int r = tid / nv, q = tid - r * nv;
        const int dr = NT / nv, dq = NT - dr * nv;
        const bool want_mm = pp.range_per_task == 1;
        while (r < h) {
            const uint8_t *g = crop00 + (size_t)r * pitch_b;
            const int lo = 16 * q - 8 * ES;                               // crop byte of this vector's first byte
            uint32_t o[4] = {0u, 0u, 0u, 0u};
            if (lo + 16 > 0 && lo < row_bytes) {
                const uint8_t *s = g + lo;
                const uint32_t a = (uint32_t)((uintptr_t)s & 15u);
                const uint4 *pa = (const uint4 *)(s - a);
                const int c0 = lo - (int)a;                                // crop byte of the first aligned word's first byte
                uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
                if (c0 + 16 > 0) v0 = __ldg(pa);
                if (a != 0 && c0 + 16 < row_bytes) v1 = __ldg(pa + 1);
                pw_realign(v0, v1, a, o);
                if (lo >= 0 && lo + 16 <= row_bytes) {
                    if (PIX != LST_PIX_U8 && swap && want_mm) {
                        uint32_t sw[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) sw[j] = __byte_perm(o[j], 0, PIX == LST_PIX_U16 ? 0x2301 : 0x0123);
#pragma unroll
                        for (int j = 0; j < 4; j++) mm_word<PIX>(sw[j], sw[j], klo, khi);
                    } else if (want_mm) {
#pragma unroll
                        for (int j = 0; j < 4; j++) mm_word<PIX>(o[j], o[j], klo, khi);
                    }
                } else {
                    const int b0 = -lo, b1 = row_bytes - lo;               // valid bytes of the vector: [b0, b1)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t m = byte_range_mask(b0 - 4 * j, b1 - 4 * j);
                        o[j] &= m;
                        if (want_mm) {
                            const uint32_t v = (PIX != LST_PIX_U8 && swap) ? __byte_perm(o[j], 0, PIX == LST_PIX_U16 ? 0x2301 : 0x0123) : o[j];
                            if (PIX != LST_PIX_F32) mm_word<PIX>(v | ~m, v, klo, khi);
                            else if (m) mm_word<PIX>(v, v, klo, khi);
                        }
                    }
                }
            }
            *(uint4 *)(raw + (size_t)r * rp + 16 * q) = make_uint4(o[0], o[1], o[2], o[3]);
            q += dq;
            r += dr;
            if (q >= nv) q -= nv, r++;
        }
    }
    __syncthreads();

    // ---- range of the float -> 8-bit conversion ----
    float lo = pp.lo, hi = pp.hi;
    if (pp.range_per_task == 2) {              // the whole slice's min/max (LST_RANGE_FRAME)
        lo = key_to_float<PIX>(pp.frame_mm[2 * t.pad1]);
        hi = key_to_float<PIX>(pp.frame_mm[2 * t.pad1 + 1]);
    } else if (pp.range_per_task) {
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
    }
    const float scale = (pp.quant_mode == LST_QUANT_ROUND255 ? 255.0f : 256.0f) / (hi - lo);
    const float qadd = pp.quant_mode == LST_QUANT_ROUND255 ? 0.5f : 0.0f;
    uint8_t *dst = win8 + t.win_off;
    const int pitch = t.pitch;

    const int nrun = (w + XRUN - 1) / XRUN;
    const int nquad = pitch >> 2;                 // the pitch padding is written (as zeros) with the last quads
    // Y-pass items (4-row group rg, quad q): idx = tid, tid + NT, ... without a division per item
    const int yrg0 = tid / nquad, yq0 = tid - yrg0 * nquad;
    const int ydrg = NT / nquad, ydq = NT - ydrg * nquad;

    int ydone = 0;                                 // output rows [0, ydone) are written
    const int ring_live = round_rows + 2 * (KR - 1);
    for (int x0r = 0; x0r < h; x0r += round_rows) {
        const int x1r = min(x0r + round_rows, h);
        const int base = x0r == 0 ? 0 : x0r - 2 * (KR - 1);      // crop row held by xp[0][.]
        // ---- X pass of rows [x0r, x1r): a warp takes (run, block of 32 rows) items, lane = row ----
        const int nblk = (x1r - x0r + 31) >> 5;
        for (int it = warp; it < nrun * nblk; it += nwarps) {
            const int blk = it / nrun, run = it - blk * nrun;
            const int i0 = XRUN * run;
            const int xr = x0r + 32 * blk + lane;
            if (xr < x1r) {
                const uint8_t *rr = raw + (size_t)xr * rp;
                float v[NLOAD];
#pragma unroll
                for (int q = 0; q < NLOAD * ES / 16; q++)
                    pt_unpack<PIX>(*(const uint4 *)(rr + (size_t)i0 * ES + 16 * q), swap, v + q * (16 / ES));   // crop columns i0-8 ..
                float o[XRUN];
                // samples outside the crop are 0 in raw[]; next to an end of the line the replicated end pixels come in
                // through the kernel-tail sums
                if (i0 >= KR && i0 + XRUN - 1 + KR < w) blur_run<XRUN, false>(v + 8 - (KR - 1), o, i0, w, 0.0f, 0.0f);
                else blur_run<XRUN, true>(v + 8 - (KR - 1), o, i0, w, pt_pixel<PIX>(rr, 8, swap), pt_pixel<PIX>(rr, 8 + w - 1, swap));
                float4 *dx = (float4 *)(xp + (size_t)(xr - base) * xwp + i0);
#pragma unroll
                for (int q = 0; q < XRUN / 4; q++) dx[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            }
        }
        __syncthreads();
        // ---- Y pass + quantise of the rows that have all their X-pass rows: [ydone, yend) ----
        const int yend = x1r == h ? h : x1r - (KR - 1);
        const int nrg = (yend - ydone + 3) >> 2;
        for (int rg = yrg0, q = yq0; rg < nrg;) {
            const int i0 = ydone + 4 * rg;
            const int c0 = 4 * q;
            const float *col = xp + (ptrdiff_t)(i0 - (KR - 1) - base) * xwp + c0;      // X-pass row i0 - (KR-1)
            float4 vv[4 + 2 * (KR - 1)];
            float o[4][4];
            const bool interior = i0 >= KR && i0 + 3 + KR < h;
            if (interior) {
#pragma unroll
                for (int j = 0; j < 4 + 2 * (KR - 1); j++) vv[j] = *(const float4 *)(col + (ptrdiff_t)j * xwp);
            } else {
#pragma unroll
                for (int j = 0; j < 4 + 2 * (KR - 1); j++) {
                    const int i = i0 - (KR - 1) + j;
                    vv[j] = (i >= 0 && i < h) ? *(const float4 *)(col + (ptrdiff_t)j * xwp) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
            }
            float4 fi = make_float4(0, 0, 0, 0), la = fi;
            if (!interior) {
                fi = *(const float4 *)(xp + (size_t)(0 - base > 0 ? 0 - base : 0) * xwp + c0);               // row 0 is here when it is needed
                la = *(const float4 *)(xp + (size_t)(h - 1 - base < ring_live ? h - 1 - base : 0) * xwp + c0);   // and so is row h-1
            }
            if (interior) {
                // Two adjacent columns per instruction where the rounding allows it: the pair sums and the products are packed
                // (FADD2 / FMUL2: two IEEE-rounded results each); the accumulation stays scalar -- ptxas contracts a packed
                // multiply followed by a packed add into FFMA2 (one rounding) whatever the PTX says, and Java rounds twice.
#pragma unroll
                for (int hp = 0; hp < 2; hp++) {
                    unsigned long long P[4 + 2 * (KR - 1)];
#pragma unroll
                    for (int j = 0; j < 4 + 2 * (KR - 1); j++) P[j] = hp == 0 ? pt_pack(vv[j].x, vv[j].y) : pt_pack(vv[j].z, vv[j].w);
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) {
                        float r0, r1;
                        pt_unpack2(pt_mul2(pt_pack(c_kern[0], c_kern[0]), P[KR - 1 + jj]), r0, r1);
#pragma unroll
                        for (int k = 1; k < KR; k++) {
                            float p0, p1;
                            pt_unpack2(pt_mul2(pt_pack(c_kern[k], c_kern[k]), pt_add2(P[KR - 1 + jj - k], P[KR - 1 + jj + k])), p0, p1);
                            r0 = r0 + p0;
                            r1 = r1 + p1;
                        }
                        o[2 * hp][jj] = r0;
                        o[2 * hp + 1][jj] = r1;
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float vc[4 + 2 * (KR - 1)];
#pragma unroll
                    for (int j = 0; j < 4 + 2 * (KR - 1); j++) vc[j] = e == 0 ? vv[j].x : (e == 1 ? vv[j].y : (e == 2 ? vv[j].z : vv[j].w));
                    blur_run<4, true>(vc, o[e], i0, h, e == 0 ? fi.x : (e == 1 ? fi.y : (e == 2 ? fi.z : fi.w)),
                                      e == 0 ? la.x : (e == 1 ? la.y : (e == 2 ? la.z : la.w)));
                }
            }
            // (int)((v - lo) * scale + 0.5) clamped to 0..255 as Java casts do (NaN -> 0, saturating); a negative argument
            // converts to 0 like the reference's max(v - lo, 0) would.  cvt.pack saturates two integers to bytes and packs them.
            // the last quads: columns right of the window are pitch padding, written as 0 (one mask per quad)
            const uint32_t keep = c0 + 3 < w ? 0xffffffffu : (c0 < w ? 0xffffffffu >> (8 * (c0 + 4 - w)) : 0u);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (i0 + j < yend) {
                    int b[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) b[e] = __float2int_rz((o[e][j] - lo) * scale + qadd);
                    uint32_t hi16, word;
                    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi16) : "r"(b[3]), "r"(b[2]), "r"(0));
                    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(word) : "r"(b[1]), "r"(b[0]), "r"(hi16));
                    *(uint32_t *)(dst + (size_t)(i0 + j) * pitch + c0) = word & keep;
                }
            }
            q += ydq;
            rg += ydrg;
            if (q >= nquad) q -= nquad, rg++;
        }
        ydone = yend;
        __syncthreads();
        // the X-pass rows the next round shares with this one move to the front
        if (x1r < h) {
            const int n4 = xwp >> 2;
            const int from = x1r - 2 * (KR - 1) - base;        // slot of crop row x1r - 12
            for (int r = warp; r < 2 * (KR - 1); r += nwarps)
                for (int c = lane; c < n4; c += 32) ((float4 *)(xp + (size_t)r * xwp))[c] = ((const float4 *)(xp + (size_t)(from + r) * xwp))[c];
            __syncthreads();
        }
    }
}

// shared-memory geometry of a pass: box width in pixels (row pitch an odd multiple of 16 bytes), X-pass row pitch
static inline int pt_box_w(int pixel_type, int max_w)
{
    const int es = pixel_type == LST_PIX_U8 ? 1 : (pixel_type == LST_PIX_U16 ? 2 : 4);
    const int xrun = pixel_type == LST_PIX_F32 ? 8 : 16;
    int need = ((max_w + xrun - 1) / xrun) * xrun + 16;       // crop columns -8 .. round_up(w, XRUN) + 7
    int bytes = (need * es + 15) & ~15;
    if (((bytes >> 4) & 1) == 0) bytes += 16;
    return bytes / es;
}
static inline int pt_xp_pitch(int max_w)
{
    int v = (max_w + 15) & ~15;
    return ((v >> 2) & 1) ? v : v + 4;
}
static inline size_t pt_smem_bytes(int pixel_type, int max_w, int max_h, int round_rows)
{
    const int es = pixel_type == LST_PIX_U8 ? 1 : (pixel_type == LST_PIX_U16 ? 2 : 4);
    const size_t rawb = ((size_t)max_h * pt_box_w(pixel_type, max_w) * es + 127) & ~(size_t)127;
    return rawb + (size_t)pt_ring_rows(round_rows) * pt_xp_pitch(max_w) * sizeof(float);
}


// n windows, all eligible (preprocess_win_eligible): one CTA per window.
int launch_preprocess_win(const DevTask *tasks, int n_tasks, int max_w, int max_h, const Surface *surfaces, const PreprocParams &pp,
                          uint8_t *win8, cudaStream_t stream)
{
    // 64-row rounds halve the barriers and the ring moves of a window; taken when two such CTAs still fit an SM
    const int round_rows = (max_h > 32 && pt_smem_bytes(pp.pixel_type, max_w, max_h, 64) <= 112 * 1024) ? 64 : 32;
    const size_t smem = pt_smem_bytes(pp.pixel_type, max_w, max_h, round_rows);
    const int xrun = pp.pixel_type == LST_PIX_F32 ? 8 : 16;
    int nt = 32 * ((max_w + xrun - 1) / xrun);
    nt = nt < 128 ? 128 : (nt > PT_MAXTHREADS ? PT_MAXTHREADS : nt);
    const int xwp = pt_xp_pitch(max_w), bw = pt_box_w(pp.pixel_type, max_w);
    if (pp.pixel_type == LST_PIX_U8)
        lst_preprocess_win<LST_PIX_U8><<<n_tasks, nt, smem, stream>>>(tasks, surfaces, pp, bw, max_h, xwp, round_rows, win8);
    else if (pp.pixel_type == LST_PIX_U16)
        lst_preprocess_win<LST_PIX_U16><<<n_tasks, nt, smem, stream>>>(tasks, surfaces, pp, bw, max_h, xwp, round_rows, win8);
    else if (pp.pixel_type == LST_PIX_F32)
        lst_preprocess_win<LST_PIX_F32><<<n_tasks, nt, smem, stream>>>(tasks, surfaces, pp, bw, max_h, xwp, round_rows, win8);
    else
        return -1;
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

bool preprocess_win_eligible(int pixel_type, int max_w, int max_h)
{
    if (pixel_type != LST_PIX_U8 && pixel_type != LST_PIX_U16 && pixel_type != LST_PIX_F32) return false;
    if (max_w > 240 || max_h > 256 || max_w < 2 * KR || max_h < 2 * KR) return false;
    return pt_smem_bytes(pixel_type, max_w, max_h, 32) <= 200 * 1024;
}

}  // namespace lst
