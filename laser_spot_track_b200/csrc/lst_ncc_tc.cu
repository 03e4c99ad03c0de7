// lst_ncc_tc.cu -- tensor-core engine of the template match: the exact integer cross-correlation as
// an implicit (Toeplitz) GEMM on tcgen05.mma kind::i8 (u8 x u8 -> s32, exact), accumulators in TMEM,
// fused with the window sums, the normalisation and the argmax in ONE persistent kernel.
//
//   out[y][x] = sum_v sum_c W[y+v][c] * T[v][c-x]                       (matchTemplate, LST4:2560)
//
// Work item = one tile (128 result rows x up to 4 blocks of 32 result columns) of one window.  One CTA
// per SM walks the items of a pass (item = blockIdx.x, + gridDim.x, ...) with these roles:
//
//   warp 8   window loader   (lane 0) the tile's window rows come in by TMA (cp.async.bulk.tensor.2d, one box of
//                            16 bytes x rows per 16-byte column chunk) as [chunk][row][16 B]: 8 consecutive
//                            rows of one chunk are a 128-byte core matrix *for any starting row*, so the A
//                            operand "rows y+v" is a start address advanced by v*16 bytes in the descriptor;
//   warp 8   table feeder    (lane 1) for template row v the Toeplitz band B_v = [D_nd-1; ..; D_0], D_d[n][k] =
//                            T[v][32d + k - n] (32 x 32 bytes each), streams from global memory (L2) through a
//                            ring with cp.async.bulk completing on mbarriers;
//   warp 10  MMA issuer      for every 32-byte K step j of the window one MMA covers all the column blocks
//                            b = j-nd+1 .. j it contributes to (N = 32..128), D = 128 TMEM lanes x 128 columns,
//                            two accumulator sets so that the MMAs of item i+1 run under the epilogue of item i;
//   warps 0-7 workers        (a) vertical sliding sums of I and I^2 per window column, straight from the staged
//                            window, into shared memory ([column][row], what cv::integral gives matchTemplate);
//                            (b) epilogue: lane = result row walks x, sliding the horizontal sums, float32
//                            filter score, OpenCV's double-precision normalisation for the candidates only,
//                            first-maximum argmax; when the window is one tile they also copy what the nine
//                            scores around the peak need (correlations from TMEM, three rows of vertical sums)
//                            into a small stash before they release the accumulator;
//   warp 9   neighbourhood   turns a stash into the nine exact scores for the sub-pixel fit (LST4:2681-2717)
//                            while the workers are already busy with the next item.
//
// No window sum ever goes through global memory; the same arithmetic as the IDP.4A engine (lst_ncc_match),
// so both engines return identical bits.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "../../include/lst_b200.h"
#include "lst_device.h"

namespace lst {

#define TC_M 128          // result rows per tile (TMEM lanes)
#define TC_NB 32          // result columns per column block
#define TC_MAXNB 4        // column blocks per tile (128 TMEM columns per accumulator set)
#define TC_MAXSTAGES 8
#define TC_MAXK 10        // K steps of a tile: nbc + nd - 1 <= 4 + 7 - 1
#define TC_WORKERS 256    // warps 0-7
#ifndef TC_NISS
#define TC_NISS 2
#endif
#define TC_NISS_DOC         // MMA issuer warps: they share the template rows of every ring stage
#define TC_NBW 9                           // the neighbourhood warp
// 12 warps = 3 per scheduler: a 13th would cap every thread at 128 registers (the register file is per scheduler).
#define TC_THREADS (320 + 32 * TC_NISS)   // + warp 8: window loader (lane 0) and table feeder (lane 1), warp 9: neighbourhood warp, warps 10..: MMA issuers (warp 10 owns the TMEM allocation)
#define TC_EPS 1e-5f
#ifndef TC_VUNROLL
#define TC_VUNROLL 4
#endif
constexpr int kVUnroll = TC_VUNROLL;   // unrolling of the vertical sliding sums
#ifndef TC_CVT
#define TC_CVT 0
#endif

int tc_nd(int tw) { return (tw + 31 + 31) / 32; }                       // Toeplitz sub-blocks per template row
size_t tc_table_bytes(int tw, int th) { return (size_t)th * tc_nd(tw) * 1024; }
bool tc_supported(int tw, int th) { return tw >= 1 && th >= 1 && tw <= 181 && th <= 181; }   // s32 accumulators

// table[v][kc][n][j] = T[v][32*d + 16*kc + j - nn] for row n = (nd-1-d)*32 + nn, kc < 2, j < 16 (0 outside the template)
void tc_build_table(const uint8_t *tpl8, int tw, int th, uint8_t *out)
{
    const int nd = tc_nd(tw);
    for (int v = 0; v < th; v++)
        for (int kc = 0; kc < 2; kc++)
            for (int n = 0; n < nd * 32; n++)
                for (int j = 0; j < 16; j++) {
                    const int d = nd - 1 - n / 32, nn = n % 32;
                    const int u = 32 * d + 16 * kc + j - nn;
                    out[(((size_t)v * 2 + kc) * (nd * 32) + n) * 16 + j] = (u >= 0 && u < tw) ? tpl8[v * tw + u] : 0;
                }
}

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 2-D tiled TMA load: box {16 bytes, rows} of the pass's 8-bit window buffer -> [row][16 B] in shared memory
__device__ __forceinline__ void tma_2d(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TC_WORKERS) : "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1):
// 8 rows x 16 B core matrices; LBO = bytes between the two K halves, SBO = bytes between 8-row groups
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = u8, K-major
__device__ __forceinline__ uint32_t instr_desc_u8(int m, int n)
{
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// The MMA warp runs its loops convergent (all 32 lanes, uniform values -> uniform registers, no R2UR / waterfall
// loop per MMA); only the instruction itself is predicated on the elected lane.
__device__ __forceinline__ void mma_i8_pred(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                            uint32_t leader)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pred(uint64_t *bar, uint32_t leader)
{
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(leader)
        : "memory");
}
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred;
}
// int64 -> float32 without I2F.S64 (a slow conversion-unit instruction): |v| as two 32-bit halves through the
// pipelined 32-bit conversions; relative error <= 2^-23
// Exact float32 (round to nearest, what I2F gives) of an unsigned integer below 2^46 WITHOUT a conversion instruction: the
// two 23-bit halves go into the mantissa of 2^23 (LOP3), the magic number is subtracted (FADD, exact) and the halves are
// combined by one FMA, whose single rounding is the rounding of the integer itself.  I2F.S64 runs on the conversion unit at a
// fraction of the FP32 rate, and the filter pass needs two of them per position.
__device__ __forceinline__ float u46_to_f32_alu(unsigned long long a)
{
    const uint32_t lo = (uint32_t)a & 0x7fffffu, hi = (uint32_t)(a >> 23);
    const float fl = __uint_as_float(lo | 0x4B000000u) - 8388608.0f;
    const float fh = __uint_as_float(hi | 0x4B000000u) - 8388608.0f;
    return __fmaf_rn(fh, 8388608.0f, fl);
}
__device__ __forceinline__ float i64_to_f32(long long v)
{
#if TC_CVT == 3
    const long long sgn = v >> 63;
    const float f = u46_to_f32_alu((unsigned long long)((v ^ sgn) - sgn));       // |v| < 2^46 for every supported template
    return __uint_as_float(__float_as_uint(f) | ((uint32_t)sgn & 0x80000000u));
#elif TC_CVT == 0
    return (float)v;
#elif TC_CVT == 1
    const long long sgn = v >> 63;
    const unsigned long long a = (unsigned long long)((v ^ sgn) - sgn);
    const float f = (float)(uint32_t)(a >> 32) * 4294967296.0f + (float)(uint32_t)a;
    return __uint_as_float(__float_as_uint(f) | ((uint32_t)(sgn) & 0x80000000u));
#else
    return (float)(int)(v >> 32) * 4294967296.0f + (float)(uint32_t)v;
#endif
}
__device__ __forceinline__ float u64_to_f32(unsigned long long a)
{
#if TC_CVT == 3
    return u46_to_f32_alu(a);
#elif TC_CVT == 0
    return (float)(long long)a;
#else
    return (float)(uint32_t)(a >> 32) * 4294967296.0f + (float)(uint32_t)a;
#endif
}
// waits of the two feeder threads: they have the highest warp ids of their schedulers, which the arbiter
// prefers -- a hot spin would take issue slots from the workers, so they sleep between polls
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(64);
    }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr)
{
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return r;
}
__device__ __forceinline__ float rsqrt_fast(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ uint32_t dcol_addr(uint32_t tmem, int acc, int q4, int col)
{
    return tmem + ((uint32_t)(32 * q4) << 16) + (uint32_t)(acc * 128 + col);      // TMEM address: lane << 16 | column
}
__device__ __forceinline__ void o_flag(float *nbhd, int task) { nbhd[(size_t)task * 12 + 9] = 1.0f; }

// The exact (double-precision, OpenCV order) score of one candidate position, out of line: the
// epilogue visits 16 positions with unrolled code and only a few of them get here, so the long
// division / square-root sequences must not be replicated (instruction cache).
__device__ __noinline__ float tc_exact_score(uint32_t corr, uint32_t s, uint32_t sq, const TplConst *tc)
{
    return match_score(*tc, (double)corr, (double)s, (double)sq);
}


// Differences of the vertical sums tw columns apart for 16 consecutive columns from xb (what the horizontal sliding sums of
// a result row add per step): 128-bit loads when the alignment allows (al: tw is a multiple of 8).
__device__ __forceinline__ void tc_load_diffs(const uint16_t *v1, const uint32_t *v2, int xb, int tw, bool al, int (&d1)[16], int (&d2)[16])
{
    uint32_t o1[8], o2[16];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint4 t4 = *(const uint4 *)(v2 + xb + 4 * j);
        o2[4 * j] = t4.x, o2[4 * j + 1] = t4.y, o2[4 * j + 2] = t4.z, o2[4 * j + 3] = t4.w;
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint4 t4 = *(const uint4 *)(v1 + xb + 8 * j);
        o1[4 * j] = t4.x, o1[4 * j + 1] = t4.y, o1[4 * j + 2] = t4.z, o1[4 * j + 3] = t4.w;
    }
    if (al) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint4 t4 = *(const uint4 *)(v2 + xb + tw + 4 * j);
            d2[4 * j] = (int)(t4.x - o2[4 * j]), d2[4 * j + 1] = (int)(t4.y - o2[4 * j + 1]);
            d2[4 * j + 2] = (int)(t4.z - o2[4 * j + 2]), d2[4 * j + 3] = (int)(t4.w - o2[4 * j + 3]);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint4 t4 = *(const uint4 *)(v1 + xb + tw + 8 * j);
            const uint32_t n4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                d1[8 * j + 2 * e] = (int)(n4[e] & 0xffffu) - (int)(o1[4 * j + e] & 0xffffu);
                d1[8 * j + 2 * e + 1] = (int)(n4[e] >> 16) - (int)(o1[4 * j + e] >> 16);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            d2[i] = (int)(v2[xb + tw + i] - o2[i]);
            d1[i] = (int)v1[xb + tw + i] - (int)((i & 1) ? (o1[i >> 1] >> 16) : (o1[i >> 1] & 0xffffu));
        }
    }
}

// the float32 filter score A / sqrt(B) of one position (TM_CCOEFF_NORMED up to the template's constant factor); the two
// passes of the epilogue must get the same bits, so both call this
__device__ __forceinline__ float tc_filter(uint32_t N, int ntsum, uint32_t corr, uint32_t S, uint32_t Q, float *Bf_out)
{
    // 32 x 32 -> 64-bit products (IMAD.WIDE): every factor fits 31 bits
    const long long A = (long long)(int)N * (long long)(int)corr + (long long)(int)S * (long long)ntsum;
    const long long B = (long long)((unsigned long long)N * Q) + (long long)(int)S * (long long)(-(int)S);
    const float Bf = u64_to_f32((unsigned long long)B);
    *Bf_out = Bf;
    return i64_to_f32(A) * rsqrt_fast(Bf);
}

// Second pass of the TM_CCOEFF_NORMED epilogue over one 16-column chunk of this lane's row (out of line: one or two warps
// of an item get here).  Every valid position whose filter score reaches thr -- and every nearly flat one, which OpenCV's
// double formula must decide -- is scored exactly; returns the lane's best key.
__device__ __noinline__ unsigned long long tc_second_pass(const uint16_t *v1, const uint32_t *v2, uint32_t taddr, int xb, int np, int tw,
                                                          uint32_t N, int ntsum, float Bflat, float thr, uint32_t idx0, const TplConst *tcp)
{
    uint32_t corr[16];
    tmem_ld16(taddr, corr);
    uint32_t S = 0, Q = 0;
#pragma unroll 1
    for (int u = 0; u < tw; u++) {
        S += v1[xb + u];
        Q += v2[xb + u];
    }
    unsigned long long mybest = 0ull;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float Bf;
        const float f = tc_filter(N, ntsum, corr[i], S, Q, &Bf);
        if (i < np && (Bf <= Bflat || f >= thr)) {
            const float sc = tc_exact_score(corr[i], S, Q, tcp);
            const unsigned long long key = pack_key(sc, idx0 + (uint32_t)i, false);
            mybest = key > mybest ? key : mybest;
        }
        S += (uint32_t)v1[xb + tw + i] - (uint32_t)v1[xb + i];
        Q += v2[xb + tw + i] - v2[xb + i];
    }
    return mybest;
}

// Light second pass over one chunk (TM_CCOEFF_NORMED): the filter scores again, from the sums the first pass had at the start
// of the chunk; the positions that can be the maximum go to the CTA's candidate list instead of being scored on the spot.
#define TC_MAXCAND 8
__device__ __forceinline__ void tc_collect(const uint16_t *v1, const uint32_t *v2, uint32_t taddr, int xb, int np, int tw, bool al, uint32_t N,
                                           int ntsum, float Bflat, float thr, uint32_t idx0, uint32_t S, uint32_t Q, uint32_t *s_ncand,
                                           uint4 *s_cand)
{
    uint32_t corr[16];
    tmem_ld16(taddr, corr);
    int d1[16], d2[16];
    tc_load_diffs(v1, v2, xb, tw, al, d1, d2);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float Bf;
        const float f = tc_filter(N, ntsum, corr[i], S, Q, &Bf);
        if (i < np && (Bf <= Bflat || f >= thr)) {
            const uint32_t slot = atomicAdd(s_ncand, 1u);
            if (slot < TC_MAXCAND) s_cand[slot] = make_uint4(idx0 + (uint32_t)i, corr[i], S, Q);
        }
        S += (uint32_t)d1[i];
        Q += (uint32_t)d2[i];
    }
}

#ifdef TC_ISS_SLEEP
#define MBAR_WAIT_ISS mbar_wait_sleep
#else
#define MBAR_WAIT_ISS mbar_wait
#endif

struct TcParams {
    int n_items;       // tasks * tiles_x * tiles_y
    int tiles_x, tiles_y;
    int nbc;           // column blocks per tile
    int tile_rows;     // result rows per tile (<= 128; the MMA always computes 128)
    int rows_alloc;    // window rows of an A buffer: 127 + max_th, rounded up to a multiple of 8
    int rows_load;     // window rows loaded per tile: tile_rows + max_th - 1
    int rows_box1;     // rows of the first TMA box (<= 256); a second box takes the rest
    int kc_a;          // 16-byte column chunks of an A buffer
    int n_abuf;        // 1 or 2 window buffers
    int vcols;         // columns of the vertical sums: 32 * nbc + max_tw
    int p1, p2;        // row pitches of the vertical sums V1 (u16 elements) / V2 (u32 elements): rows 16-byte aligned,
                       // pitch in words = 4 mod 8, so that 32 lanes owning 32 rows load 128-bit words conflict-free
    int stages, vb;    // ring: stages, template rows per stage
    int stage_bytes;
    int single_tile;   // every window of the pass is one tile: the kernel also leaves the 3x3 scores around the peak
    unsigned long long *dbg;   // optional per-CTA phase clocks (LST_TC_DEBUG)
    int mode;          // experiments (LST_TC_MODE): 1 = workers skip their work, 2 = no MMAs issued, 4 = workers skip the vertical sums, 8 = skip the epilogue
};

// what every role needs to know about a work item
struct TcItem {
    int task, tpl;
    int x0, y0;        // tile origin in the result map
    int rw_t, rh_t;    // valid result columns / rows of this tile
    int nbe;           // column blocks with valid columns
    int nk;            // 32-byte K steps of the window the tile needs
    int tw, th, nd;
    int row_base;      // first row of the window in the pass's window buffer (win_off / pitch)
    int rw;            // result-map width of the task
};

// returns 0 = nothing to do for this tile, 1 = match it, 2 = flat template (method 5): the map is all ones
__device__ __forceinline__ int tc_decode(int item, const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls,
                                         const TcParams &tp, TcItem *it)
{
    const int tpt = tp.tiles_x * tp.tiles_y;
    const int task = item / tpt, tile = item - task * tpt;
    const int ty = tile / tp.tiles_x, tx = tile - ty * tp.tiles_x;
    const DevTask *t = tasks + task;
    const int rw = t->rw, rh = t->rh, tpl = t->tpl;
    it->task = task;
    it->tpl = tpl;
    it->x0 = tx * tp.nbc * TC_NB;
    it->y0 = ty * tp.tile_rows;
    if (it->x0 >= rw || it->y0 >= rh) return 0;
    const TplConst *c = &tpls->c[tpl];
    if (c->flat && c->method == 5) return tile == 0 ? 2 : 0;
    it->tw = c->tw;
    it->th = c->th;
    it->nd = (it->tw + 62) >> 5;
    it->rw = rw;
    it->rw_t = min(tp.nbc * TC_NB, rw - it->x0);
    it->rh_t = min(tp.tile_rows, rh - it->y0);
    it->nbe = (it->rw_t + TC_NB - 1) / TC_NB;
    it->nk = ((it->rw_t + it->tw - 2) >> 5) + 1;
    it->row_base = (int)(t->win_off / (uint64_t)t->pitch);
    return 1;
}

// What the workers leave for the neighbourhood warp: the peak, the nine correlations around it and rows by-1 .. by+1 of
// the vertical sums at columns bx-1 .. bx+tw (0 outside the window).
#define TC_STASH_COLS 192                 // >= tw + 2 for every supported template (tw <= 181)
struct TcStash {
    int task, tpl, by, bx, rw, rh, tw, need_best;    // need_best: the peak's exact score has not been computed yet
    uint32_t corr[12];
    uint32_t v2[3][TC_STASH_COLS];
    uint16_t v1[3][TC_STASH_COLS];
};

// What the MMA issuer needs for one item.
struct TcIssue {
    uint64_t a_desc0;
    uint32_t a_kstep, b_hi32, b_lbo, b_vstep, dcol, idesc0;
    int th, g, iss, stages, vb, stage_bytes, nd, nbe, nk, issue;
    uint8_t *sB;
    uint64_t *r_full, *r_empty, *row0;
};

// Issues the MMAs of one item (the template rows this issuer owns).  K step j feeds the column blocks b_lo = max(0, j-nd+1)
// .. b_hi = min(nbe-1, j); its band starts (nd-1-(j-b_lo)) blocks into the table row.  In the first template row column
// block j gets its first contribution from K step j and must be overwritten, the blocks left of it accumulate.
// ND, NBE > 0: the band shape is a compile-time constant, so the column ranges, band offsets and instruction descriptors
// of the K steps are immediates and an MMA costs two adds; ND = 0: any shape.  NK < ND + NBE - 1: the last K steps would only
// feed result columns beyond the map (a 113-column map in four 32-column blocks: window columns 160.. reach columns >= 113
// only) and are not issued.  Returns the cycles spent waiting for the ring.
template <int ND, int NBE, int NK = ND + NBE - 1>
__device__ __forceinline__ long long tc_issue_item(const TcIssue &q)
{
    const int nd = ND > 0 ? ND : q.nd, nbe = ND > 0 ? NBE : q.nbe, nk = ND > 0 ? NK : q.nk;
    constexpr int KMAX = ND > 0 ? NK : TC_MAXK;
    const uint32_t a_hi32 = (uint32_t)(q.a_desc0 >> 32);
    long long waited = 0;
    int v = 0, gg = q.g;
    for (; v < q.th; gg++) {
        const int s = gg % q.stages;
        const int nv = min(q.vb, q.th - v);
        const long long t3 = clock64();
        MBAR_WAIT_ISS(&q.r_full[s], (gg / q.stages) & 1);
        waited += clock64() - t3;
        tc_fence_after();
        uint32_t b_v = (uint32_t)smem_desc(smem_u32(q.sB + (size_t)s * q.stage_bytes), q.b_lbo, 128);
        for (int vv = 0; vv < nv; vv++, v++, b_v += q.b_vstep) {
            if (vv % TC_NISS != q.iss) continue;
            const uint32_t a_lo = (uint32_t)q.a_desc0 + (uint32_t)v;
            if (!q.issue) {
                if (v == 0 && TC_NISS > 1) mbar_arrive(q.row0);
            } else if (v == 0) {
#pragma unroll
                for (int j = 0; j < KMAX; j++) {
                    if (j < nk) {
                        const int b_lo = max(0, j - nd + 1), b_hi = min(nbe - 1, j);
                        const uint64_t a_d = ((uint64_t)a_hi32 << 32) | (a_lo + (uint32_t)j * q.a_kstep);
                        const uint64_t b_d = ((uint64_t)q.b_hi32 << 32) | (b_v + (uint32_t)((nd - 1 - (j - b_lo)) * 32));
                        if (j < nbe) {
                            const uint32_t nold = (uint32_t)(j - b_lo);       // blocks that already hold a partial sum
                            if (nold > 0) mma_i8(q.dcol + (uint32_t)(32 * b_lo), a_d, b_d, q.idesc0 | (nold << 19), 1u);
                            mma_i8(q.dcol + (uint32_t)(32 * j), a_d, b_d + (uint64_t)(nold * 32u), q.idesc0 | (1u << 19), 0u);
                        } else {
                            mma_i8(q.dcol + (uint32_t)(32 * b_lo), a_d, b_d, q.idesc0 | ((uint32_t)(b_hi - b_lo + 1) << 19), 1u);
                        }
                    }
                }
                if (TC_NISS > 1) {
                    tc_fence_before();
                    mbar_arrive(q.row0);
                }
            } else {
#pragma unroll
                for (int j = 0; j < KMAX; j++) {
                    if (j < nk) {
                        const int b_lo = max(0, j - nd + 1), b_hi = min(nbe - 1, j);
                        mma_i8(q.dcol + (uint32_t)(32 * b_lo), ((uint64_t)a_hi32 << 32) | (a_lo + (uint32_t)j * q.a_kstep),
                               ((uint64_t)q.b_hi32 << 32) | (b_v + (uint32_t)((nd - 1 - (j - b_lo)) * 32)),
                               q.idesc0 | ((uint32_t)(b_hi - b_lo + 1) << 19), 1u);
                    }
                }
            }
        }
        tc_commit(&q.r_empty[s]);     // the ring slot is free once these MMAs have read it
    }
    return waited;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
lst_ncc_tc(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
           const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls, unsigned long long *__restrict__ best,
           float *__restrict__ nbhd, const TcParams tp)
{
    extern __shared__ __align__(128) uint8_t smem_tc[];
    __shared__ __align__(8) uint64_t a_full[2], a_free[2], acc_full[2], acc_free[2], row0_done[2], r_full[TC_MAXSTAGES], r_empty[TC_MAXSTAGES];
    __shared__ uint32_t s_tmem;
    __shared__ unsigned long long s_best[8];
    __shared__ uint32_t s_fmax;    // orderable key of the largest filter score any warp has seen in the current item
    __shared__ __align__(8) uint64_t nb_full[2], nb_free[2];
    __shared__ __align__(16) TcStash s_stash[2];
    __shared__ uint32_t s_ncand;   // positions of the current item that can be its maximum (may exceed TC_MAXCAND)
    __shared__ __align__(16) uint4 s_cand[TC_MAXCAND];   // {map index, correlation, sum I, sum I^2}
    __shared__ __align__(16) TcItem s_item[2];     // the item whose window is in A buffer [i]; task < 0 = no more items

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t a_bytes = (size_t)tp.kc_a * tp.rows_alloc * 16;
    uint8_t *sA = smem_tc;                                                  // [n_abuf][kc_a][rows_alloc][16]
    uint8_t *sB = sA + (size_t)tp.n_abuf * a_bytes;                         // [stages][stage_bytes]
    uint32_t *V2 = (uint32_t *)(sB + (size_t)tp.stages * tp.stage_bytes);   // [tile_rows][p2]  sums of I^2 over th rows, per window column
    uint16_t *V1 = (uint16_t *)(V2 + (size_t)tp.tile_rows * tp.p2);         // [tile_rows][p1]  sums of I

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_free[i], TC_NISS + 1);       // the MMAs that read the buffer (every issuer) + the workers' vertical sums
            mbar_init(&acc_full[i], TC_NISS);
            mbar_init(&row0_done[i], 1);
            mbar_init(&acc_free[i], 1);
            mbar_init(&nb_full[i], 1);
            mbar_init(&nb_free[i], 1);
        }
        for (int s = 0; s < TC_MAXSTAGES; s++) {
            mbar_init(&r_full[s], 1);
            mbar_init(&r_empty[s], TC_NISS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 10) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == 8) {
        // ---- window loader: the only role that decodes items (divisions, dependent global loads); the others read the
        //      decoded item next to its window once the window has landed ----
        if (lane == 0) {
            int k = 0;
            for (int item = blockIdx.x;; item += gridDim.x) {
                TcItem it;
                int kind = -1;
                if (item < tp.n_items) {
                    kind = tc_decode(item, tasks, tpls, tp, &it);
                    if (kind == 2) atomicMax(&best[it.task], pack_key(1.0f, 0u));   // templSdv^2 < DBL_EPSILON: all ones, first wins
                    if (kind != 1) continue;
                }
                const int buf = tp.n_abuf == 2 ? (k & 1) : 0, u = tp.n_abuf == 2 ? (k >> 1) : k;
                if (u > 0) mbar_wait_sleep(&a_free[buf], (u - 1) & 1);
                if (kind < 0) {
                    s_item[buf].task = -1;
                    mbar_arrive(&a_full[buf]);
                    break;
                }
                s_item[buf] = it;
                const int nch = 2 * it.nk;                       // chunks this tile needs
                mbar_expect_tx(&a_full[buf], (uint32_t)(nch * tp.rows_load * 16));
                uint8_t *dst = sA + (size_t)buf * a_bytes;
                const int r0 = it.row_base + it.y0;
                for (int kc = 0; kc < nch; kc++) {
                    tma_2d(dst + (size_t)kc * tp.rows_alloc * 16, &map1, it.x0 + 16 * kc, r0, &a_full[buf]);
                    if (tp.rows_load > tp.rows_box1)
                        tma_2d(dst + ((size_t)kc * tp.rows_alloc + tp.rows_box1) * 16, &map2, it.x0 + 16 * kc, r0 + tp.rows_box1,
                               &a_full[buf]);
                }
                k++;
            }
        } else if (lane == 1) {
            // ---- Toeplitz table feeder: the other side of the same branch, so the two lanes block on their barriers and
            //      progress independently of each other (independent thread scheduling) ----
            int g = 0;
            for (int k = 0;; k++) {
                const int buf = tp.n_abuf == 2 ? (k & 1) : 0, ua = tp.n_abuf == 2 ? (k >> 1) : k;
                mbar_wait_sleep(&a_full[buf], ua & 1);
                const int tpl_k = s_item[buf].task < 0 ? -1 : s_item[buf].tpl;
                if (tpl_k < 0) break;
                struct { int tpl, th, nd; } it = {tpl_k, s_item[buf].th, s_item[buf].nd};
                const uint8_t *table = tpls->toeplitz[it.tpl];
                const uint32_t row_bytes = (uint32_t)it.nd * 1024u;
                for (int v = 0; v < it.th; v += tp.vb, g++) {
                    const int s = g % tp.stages;
                    if (g >= tp.stages) mbar_wait_sleep(&r_empty[s], ((g / tp.stages) - 1) & 1);
                    const uint32_t bytes = row_bytes * (uint32_t)min(tp.vb, it.th - v);
                    mbar_expect_tx(&r_full[s], bytes);
                    bulk_g2s(sB + (size_t)s * tp.stage_bytes, table + (size_t)v * row_bytes, bytes, &r_full[s]);
                }
            }
        }
    } else if (warp >= 10) {
        // ---- MMA issuers ----
        // One thread issues everything, so the work per MMA must be a handful of instructions: per item the
        // descriptors of the K steps are set up once in registers (the loops over j are fully unrolled); per template
        // row the A descriptors advance by one row (+1 in 16-byte units) and the B descriptors by one table row --
        // 32-bit adds on the low words, the start-address field never carries out of its 14 bits.
        // The loop over the items runs convergent: all lanes decode the item and its shape goes through a warp
        // reduction, whose result ptxas knows to be uniform -- the descriptors then live in uniform registers and a
        // UTCIMMA costs a couple of uniform adds instead of five R2UR transfers.  One elected lane issues an item's MMAs.
        // A single thread cannot issue faster than one MMA per ~125 cycles (descriptor moves into uniform registers), so
        // TC_NISS warps share the work: issuer i takes the template rows i, i + TC_NISS, .. of every ring stage.  The
        // accumulation is integer, so the order in which the tensor core takes their MMAs does not matter -- except that
        // template row 0 (issuer 0), which overwrites the accumulators, must come first (row0_done).
        {
            const int iss = warp - 10;
            const uint32_t tmem_u = (uint32_t)__reduce_max_sync(0xffffffffu, tmem);
            const uint32_t idesc0 = instr_desc_u8(TC_M, 0);
            int k = 0, g = 0;
            long long m_wa = 0, m_wc = 0, m_wr = 0, m_tot = 0;
            const uint32_t b_hi32 = (uint32_t)(smem_desc(0, 0, 128) >> 32);
            for (;;) {
                // all lanes wait for the window, read the item beside it and make its shape warp-uniform
                {
                    const int buf_w = tp.n_abuf == 2 ? (k & 1) : 0, ua_w = tp.n_abuf == 2 ? (k >> 1) : k;
                    MBAR_WAIT_ISS(&a_full[buf_w], ua_w & 1);
                }
                const TcItem &sit = s_item[tp.n_abuf == 2 ? (k & 1) : 0];
                const uint32_t pk = __reduce_max_sync(0xffffffffu, sit.task < 0 ? 0u : 1u | ((uint32_t)sit.nd << 2) | ((uint32_t)sit.nbe << 5) |
                                                                                          ((uint32_t)sit.nk << 8) | ((uint32_t)sit.th << 12));
                if ((pk & 3u) != 1u) break;
                const int nd = (int)((pk >> 2) & 7u), nbe = (int)((pk >> 5) & 7u), nk = (int)((pk >> 8) & 15u), th = (int)(pk >> 12);
                const int buf = tp.n_abuf == 2 ? (k & 1) : 0, ua = tp.n_abuf == 2 ? (k >> 1) : k;
                const int acc = k & 1, uc = k >> 1;
                const uint64_t a_desc0 = smem_desc(smem_u32(sA + (size_t)buf * a_bytes), (uint32_t)tp.rows_alloc * 16, 128);
                const uint32_t a_hi32 = (uint32_t)(a_desc0 >> 32);
                const uint32_t b_lbo = (uint32_t)nd * 32 * 16;                  // the second K half of a band: nd*32 rows further
                const uint32_t b_vstep = (uint32_t)(nd * 64);                  // one template row of the table, in 16-byte units
                const uint32_t dcol = tmem_u + (uint32_t)(acc * 128);
                const uint32_t a_kstep = (uint32_t)(2 * tp.rows_alloc);        // two 16-byte column chunks per K step
                const int nsteps = (th + tp.vb - 1) / tp.vb;
                if (elect_one()) {
                    const long long t0 = clock64();
                    mbar_wait(&a_full[buf], ua & 1);
                    const long long t1 = clock64();
                    if (uc > 0) MBAR_WAIT_ISS(&acc_free[acc], (uc - 1) & 1);
                    if (iss != 0) MBAR_WAIT_ISS(&row0_done[acc], uc & 1);
                    const long long t2 = clock64();
                    m_wa += t1 - t0;
                    m_wc += t2 - t1;
                    tc_fence_after();
                    TcIssue q;
                    q.a_desc0 = a_desc0, q.a_kstep = a_kstep, q.b_hi32 = b_hi32, q.b_lbo = b_lbo, q.b_vstep = b_vstep, q.dcol = dcol;
                    q.idesc0 = idesc0, q.th = th, q.g = g, q.iss = iss, q.stages = tp.stages, q.vb = tp.vb, q.stage_bytes = tp.stage_bytes;
                    q.sB = sB, q.r_full = r_full, q.r_empty = r_empty, q.row0 = &row0_done[acc], q.nd = nd, q.nbe = nbe, q.nk = nk;
                    q.issue = !(tp.mode & 2);
                    // the loops are specialised for the band shapes of the named configurations (compile-time column ranges)
                    long long wr;
                    if (nd == 3 && nbe == 4 && nk == 5) wr = tc_issue_item<3, 4, 5>(q);      // 48-column templates, 113-column maps (config B)
                    else if (nd == 3 && nbe == 4) wr = tc_issue_item<3, 4>(q);     // templates of 33..64 columns, four column blocks
                    else if (nd == 2 && nbe == 3 && nk == 3) wr = tc_issue_item<2, 3, 3>(q);   // 32-column templates, 65-column maps (config A)
                    else if (nd == 2 && nbe == 3) wr = tc_issue_item<2, 3>(q);     // templates up to 32 columns, three blocks
                    else if (nd == 3 && nbe == 2) wr = tc_issue_item<3, 2>(q);     // 64-column templates, two blocks (config C tiles)
                    else if (nd == 5 && nbe == 2) wr = tc_issue_item<5, 2>(q);     // 128-column templates, two blocks (config D tiles)
                    else wr = tc_issue_item<0, 0>(q);
                    m_wr += wr;
                    tc_commit(&a_free[buf]);
                    tc_commit(&acc_full[acc]);
                    m_tot += clock64() - t0;
                }
                __syncwarp();
                g += nsteps;
                k++;
            }
            // whichever lane was elected holds the clocks
            m_wa = __reduce_max_sync(0xffffffffu, (unsigned)min(m_wa, 0xffffffffLL)), m_wc = __reduce_max_sync(0xffffffffu, (unsigned)min(m_wc, 0xffffffffLL));
            m_wr = __reduce_max_sync(0xffffffffu, (unsigned)min(m_wr, 0xffffffffLL)), m_tot = __reduce_max_sync(0xffffffffu, (unsigned)min(m_tot, 0xffffffffLL));
            if (tp.dbg && lane == 0 && iss == 0) {
                unsigned long long *d = tp.dbg + (size_t)blockIdx.x * 8 + 4;
                d[0] = (unsigned long long)m_wa;
                d[1] = (unsigned long long)m_wc;
                d[2] = (unsigned long long)m_wr;
                d[3] = (unsigned long long)m_tot;
            }
        }
    } else if (warp == TC_NBW) {
        // ---- neighbourhood warp: the nine exact scores around the peak of a single-tile window ----
        if (tp.single_tile && nbhd) {
            for (int k = 0;; k++) {
                const TcStash *st = &s_stash[k & 1];
                mbar_wait_sleep(&nb_full[k & 1], (k >> 1) & 1);
                if (st->task < 0) break;        // the workers have seen the end of the items
                const int by = st->by, bx = st->bx, tw = st->tw;
                // Window sums of a row: the lanes add the tw + 2 vertical sums of columns bx-1 .. bx+tw (full sum F), then
                // S(bx-1) = F - c[tw] - c[tw+1], S(bx) = F - c[0] - c[tw+1], S(bx+1) = F - c[0] - c[1].
                uint32_t myS = 0, myQ = 0;
#pragma unroll
                for (int dy = 0; dy < 3; dy++) {
                    uint32_t f1 = 0, f2 = 0;
                    for (int c = lane; c < tw + 2; c += 32) f1 += st->v1[dy][c], f2 += st->v2[dy][c];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        f1 += __shfl_xor_sync(0xffffffffu, f1, o);
                        f2 += __shfl_xor_sync(0xffffffffu, f2, o);
                    }
                    const uint32_t a1 = st->v1[dy][0], b1 = st->v1[dy][1], c1 = st->v1[dy][tw], d1 = st->v1[dy][tw + 1];
                    const uint32_t a2 = st->v2[dy][0], b2 = st->v2[dy][1], c2 = st->v2[dy][tw], d2 = st->v2[dy][tw + 1];
                    if (lane == 3 * dy) myS = f1 - c1 - d1, myQ = f2 - c2 - d2;
                    if (lane == 3 * dy + 1) myS = f1 - a1 - d1, myQ = f2 - a2 - d2;
                    if (lane == 3 * dy + 2) myS = f1 - a1 - b1, myQ = f2 - a2 - b2;
                }
                const int dy = lane / 3 - 1, dx = lane - 3 * (lane / 3) - 1;
                if (lane < 9 && by + dy >= 0 && by + dy < st->rh && bx + dx >= 0 && bx + dx < st->rw) {
                    const float sc = tc_exact_score(st->corr[lane], myS, myQ, &tpls->c[st->tpl]);
                    nbhd[(size_t)st->task * 12 + lane] = sc;
                    if (lane == 4 && st->need_best) best[st->task] = pack_key(sc, (uint32_t)(by * st->rw + bx), false);
                }
                const int task = st->task;
                __syncwarp();
                if (lane == 0) {
                    o_flag(nbhd, task);
                    mbar_arrive(&nb_free[k & 1]);
                }
            }
        }
    } else {
        // ---- workers: vertical sums, then the epilogue of the item ----
        const int tid = threadIdx.x;
        const int q4 = warp & 3, hh = warp >> 2;         // TMEM lane quarter (rows 32*q4 ..), column half
        const int row = 32 * q4 + lane;
        const int P1 = tp.p1, P2 = tp.p2;                // row pitches of V1 (in u16) and V2 (in u32)
        long long clk_wait = 0, clk_v = 0, clk_e = 0, clk_e1 = 0, clk_e2 = 0, clk_e3 = 0, clk_e4 = 0, clk_e5 = 0, clk_p1 = 0;
        const bool do_nb = tp.single_tile && nbhd != nullptr;
        int k = 0;
        for (;;) {
            const int buf = tp.n_abuf == 2 ? (k & 1) : 0, ua = tp.n_abuf == 2 ? (k >> 1) : k;
            const int acc = k & 1, uc = k >> 1;
            const long long c0 = clock64();
            mbar_wait(&a_full[buf], ua & 1);
            const long long c1 = clock64();
            const TcItem it = s_item[buf];      // decoded by the window loader
            if (it.task < 0) {
                if (do_nb && tid == 0) {        // tell the neighbourhood warp
                    if (k >= 2) mbar_wait(&nb_free[k & 1], ((k >> 1) - 1) & 1);
                    s_stash[k & 1].task = -1;
                    mbar_arrive(&nb_full[k & 1]);
                }
                break;
            }
            const TplConst *tcp = &tpls->c[it.tpl];
            const int tw = it.tw, th = it.th;
            // (a) vertical sliding sums: a thread owns two adjacent window columns (16-bit loads) over a segment of the
            //     tile's rows; V1 / V2 are [row][column], so a pair is one 32-bit / one 64-bit store
            if (!(tp.mode & 5)) {
                const int ncv = it.rw_t + tw - 1;
                // work item = (16-byte column chunk, row segment, pair within the chunk): the lanes of a warp read one or two
                // chunks at different rows -- chunk columns are a multiple of 128 bytes apart (all in the same four banks),
                // rows 16 bytes
                const int nchunk = (ncv + 15) >> 4;
                int nseg = TC_WORKERS / (8 * nchunk);
                nseg = nseg < 1 ? 1 : (nseg > 6 ? 6 : nseg);
                if (it.rh_t < 8 * nseg) nseg = it.rh_t >= 8 ? it.rh_t / 8 : 1;
                int seg_rows = (it.rh_t + nseg - 1) / nseg;
                if ((seg_rows & 7) == 0 && nseg > 1) seg_rows++;          // segments start in different banks
                const uint8_t *A = sA + (size_t)buf * a_bytes;
                for (int w = tid; w < 8 * nchunk * nseg; w += TC_WORKERS) {
                    const int rest = w >> 3, ch = rest / nseg, sg = rest - ch * nseg, c = 16 * ch + 2 * (w & 7);
                    const int r0 = sg * seg_rows, r1 = min(r0 + seg_rows, it.rh_t);
                    if (r0 >= r1 || c >= ncv) continue;
                    const uint8_t *p = A + ((size_t)(c >> 4) * tp.rows_alloc) * 16 + (c & 15);
                    uint32_t s0 = 0, s1 = 0, q0 = 0, q1 = 0;
                    const uint8_t *pv = p + (size_t)r0 * 16;
#pragma unroll 8
                    for (int v = 0; v < th; v++) {
                        const uint32_t px = *(const uint16_t *)(pv + v * 16);
                        const uint32_t a0 = px & 0xffu, a1 = px >> 8;
                        s0 += a0, s1 += a1;
                        q0 += a0 * a0, q1 += a1 * a1;
                    }
                    uint32_t *o1 = (uint32_t *)(V1 + (size_t)r0 * P1 + c);
                    uint2 *o2 = (uint2 *)(V2 + (size_t)r0 * P2 + c);
                    *o1 = s0 | (s1 << 16);
                    *o2 = make_uint2(q0, q1);
                    const uint8_t *pn = p + (size_t)(r0 + th) * 16, *po = p + (size_t)r0 * 16;
#pragma unroll kVUnroll
                    for (int r = r0 + 1; r < r1; r++) {
                        const uint32_t wn = *(const uint16_t *)pn, wo = *(const uint16_t *)po;
                        pn += 16, po += 16;
                        o1 = (uint32_t *)((uint16_t *)o1 + P1);
                        o2 = (uint2 *)((uint32_t *)o2 + P2);
                        const int a0 = wn & 0xffu, a1 = wn >> 8, b0 = wo & 0xffu, b1 = wo >> 8;
                        const int d0 = a0 - b0, d1 = a1 - b1;
                        s0 += (uint32_t)d0, s1 += (uint32_t)d1;
                        q0 += (uint32_t)(d0 * (a0 + b0)), q1 += (uint32_t)(d1 * (a1 + b1));
                        *o1 = s0 | (s1 << 16);
                        *o2 = make_uint2(q0, q1);
                    }
                }
            }
            if (tid == 0) s_fmax = 0u, s_ncand = 0u;
            worker_bar();
            if (tid == 0) mbar_arrive(&a_free[buf]);
            const long long c2 = clock64();
            mbar_wait(&acc_full[acc], uc & 1);
            tc_fence_after();
            const long long c3 = clock64();
            // (b) epilogue: rows 32*q4 .. +31 (TMEM lanes of this warp), column blocks of this warp's half
            const bool row_ok = row < it.rh_t;
            const int vrow = row_ok ? row : 0;
            const int y = it.y0 + row;
            const uint32_t N = (uint32_t)(tw * th);
            const int ntsum = -(int)tcp->tsum;
            const float Bflat = (float)(((long long)N + (long long)(N >> 5)) >> 1);   // 2*B <= N + N/32: a nearly flat window
            const int method = tcp->method;
            const bool want_min = method_wants_min(method);
            const float eps_g = method == 5 ? TC_EPS / tcp->inv_sqrt_c : 0.0f;   // the filter compares A / sqrt(B); score = that * inv_sqrt_c
            const int nbh = (it.nbe + 1) >> 1;
            const int nb0 = hh == 0 ? 0 : nbh, nb1 = hh == 0 ? nbh : it.nbe;
            const bool al = ((tw & 7) == 0);               // the columns tw further on are 16-byte aligned as well
            unsigned long long mybest = 0ull;
            float runmax = -3.0e38f, cm0 = -3.0e38f, cm1 = -3.0e38f, cm2 = -3.0e38f, cm3 = -3.0e38f;   // first-pass maxima: lane, lane's chunks
            uint32_t sc0 = 0, sc1 = 0, sc2 = 0, sc3 = 0, qc0 = 0, qc1 = 0, qc2 = 0, qc3 = 0;           // window sums at the start of the lane's chunks
            const bool epi_active = nb0 < nb1 && 32 * q4 < it.rh_t && !(tp.mode & 9);
            if (epi_active) {
                const uint16_t *v1 = V1 + (size_t)vrow * P1;
                const uint32_t *v2 = V2 + (size_t)vrow * P2;
                // horizontal sums under the template footprint at the first column of this warp's range
                uint32_t S = 0, Q = 0;
                {
                    const int xs = TC_NB * nb0;
                    int u = 0;
                    for (; u + 8 <= tw; u += 8) {           // xs is a multiple of 32: 16-byte aligned vectors
                        const uint4 a4 = *(const uint4 *)(v2 + xs + u), b4 = *(const uint4 *)(v2 + xs + u + 4);
                        const uint4 c4 = *(const uint4 *)(v1 + xs + u);
                        Q += (a4.x + a4.y) + (a4.z + a4.w) + (b4.x + b4.y) + (b4.z + b4.w);
                        S += (c4.x & 0xffffu) + (c4.x >> 16) + (c4.y & 0xffffu) + (c4.y >> 16) + (c4.z & 0xffffu) + (c4.z >> 16) +
                             (c4.w & 0xffffu) + (c4.w >> 16);
                    }
                    for (; u < tw; u++) {
                        S += v1[xs + u];
                        Q += v2[xs + u];
                    }
                }
                clk_e1 += clock64() - c3;
                const uint32_t rowidx = (uint32_t)y * (uint32_t)it.rw + (uint32_t)it.x0;
                const int xend = min(TC_NB * nb1, it.rw_t);
                int ci = 0;
#pragma unroll 1
                for (int xb = TC_NB * nb0; xb < xend; xb += 16, ci++) {
                    uint32_t corr[16];
                    tmem_ld16(dcol_addr(tmem, acc, q4, xb), corr);
                    const int np = row_ok ? min(16, it.rw_t - xb) : 0;
                    int d1[16], d2[16];
                    tc_load_diffs(v1, v2, xb, tw, al, d1, d2);
                    if (method == 5) {
                        // First pass: the filter score f = A / sqrt(B) of the 16 positions with nothing but the arithmetic -- no
                        // validity or flatness selects, no exact scores: only the lane's maximum per chunk is kept.  A lane whose
                        // chunk is partial (np < 16) or holds a nearly flat footprint (2B <= N + N/32; OpenCV's double formula
                        // decides those) redoes its maximum below.
                        const uint32_t S0 = S, Q0 = Q;
                        float tmax = -3.0e38f, bmin = 3.0e38f;
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            float Bf;
                            const float f = tc_filter(N, ntsum, corr[i], S, Q, &Bf);
                            tmax = fmaxf(tmax, f);          // a NaN (B = 0) is ignored here and caught by bmin
                            bmin = fminf(bmin, Bf);
                            S += (uint32_t)d1[i];
                            Q += (uint32_t)d2[i];
                        }
                        float cmv = tmax;                   // what the second pass compares with the threshold
                        if (!row_ok) {                      // a lane below the map: nothing to offer (and no reason to redo anything)
                            tmax = cmv = -3.0e38f;
                        } else if (np < 16 || bmin <= Bflat) {
                            tmax = -3.0e38f;
                            bool special = false;           // a valid, nearly flat position: always scored exactly
                            uint32_t Sr = S0, Qr = Q0;
#pragma unroll
                            for (int i = 0; i < 16; i++) {
                                float Bf;
                                const float f = tc_filter(N, ntsum, corr[i], Sr, Qr, &Bf);
                                if (i < np) {
                                    if (Bf <= Bflat) special = true;      // exact: integers below 2^24 when it matters
                                    else tmax = fmaxf(tmax, f);
                                }
                                Sr += (uint32_t)d1[i];
                                Qr += (uint32_t)d2[i];
                            }
                            cmv = special ? 3.0e38f : tmax;
                        }
                        runmax = fmaxf(runmax, tmax);
                        if (ci == 0) cm0 = cmv, sc0 = S0, qc0 = Q0;
                        else if (ci == 1) cm1 = cmv, sc1 = S0, qc1 = Q0;
                        else if (ci == 2) cm2 = cmv, sc2 = S0, qc2 = Q0;
                        else cm3 = cmv, sc3 = S0, qc3 = Q0;
                    } else {
                        // the other five methods: every position is scored with OpenCV's formula
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            if (i < np) {
                                const float sc = tc_exact_score(corr[i], S, Q, tcp);
                                const unsigned long long key = pack_key(sc, rowidx + (uint32_t)(xb + i), want_min);
                                mybest = key > mybest ? key : mybest;
                            }
                            S += (uint32_t)d1[i];
                            Q += (uint32_t)d2[i];
                        }
                    }
                }
            }
            clk_p1 += clock64() - c3;
            uint32_t ncand = TC_MAXCAND + 1;       // the other methods: every position was scored in the first pass
            if (method == 5) {
                // Second pass: only positions within eps of the largest filter score of the whole tile can be its maximum.  The
                // warps that hold such positions put them on the candidate list; usually there is one.
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) runmax = fmaxf(runmax, __shfl_xor_sync(0xffffffffu, runmax, o));
                if (lane == 0 && runmax > -3.0e38f) atomicMax(&s_fmax, f32_key(runmax));
                worker_bar();
                const uint32_t kmax = *(volatile uint32_t *)&s_fmax;
                const float thr = kmax ? key_f32(kmax) - eps_g : -3.0e38f;
                const uint16_t *v1 = V1 + (size_t)vrow * P1;
                const uint32_t *v2 = V2 + (size_t)vrow * P2;
                const uint32_t rowidx = (uint32_t)y * (uint32_t)it.rw + (uint32_t)it.x0;
                const int xend = min(TC_NB * nb1, it.rw_t);
                if (epi_active) {
                    int ci = 0;
#pragma unroll 1
                    for (int xb = TC_NB * nb0; xb < xend; xb += 16, ci++) {
                        const float cmv = ci == 0 ? cm0 : (ci == 1 ? cm1 : (ci == 2 ? cm2 : cm3));
                        if (!__any_sync(0xffffffffu, cmv >= thr)) continue;
                        const uint32_t Sc = ci == 0 ? sc0 : (ci == 1 ? sc1 : (ci == 2 ? sc2 : sc3));
                        const uint32_t Qc = ci == 0 ? qc0 : (ci == 1 ? qc1 : (ci == 2 ? qc2 : qc3));
                        const int np = row_ok ? min(16, it.rw_t - xb) : 0;
                        tc_collect(v1, v2, dcol_addr(tmem, acc, q4, xb), xb, np, tw, al, N, ntsum, Bflat, thr, rowidx + (uint32_t)xb, Sc, Qc,
                                   &s_ncand, s_cand);
                    }
                }
                worker_bar();
                ncand = *(volatile uint32_t *)&s_ncand;
                if (ncand > TC_MAXCAND && epi_active) {
                    // too many for the list (flat or periodic windows): every lane scores its own candidates
                    int ci = 0;
#pragma unroll 1
                    for (int xb = TC_NB * nb0; xb < xend; xb += 16, ci++) {
                        const float cmv = ci == 0 ? cm0 : (ci == 1 ? cm1 : (ci == 2 ? cm2 : cm3));
                        if (!__any_sync(0xffffffffu, cmv >= thr)) continue;
                        const int np = row_ok ? min(16, it.rw_t - xb) : 0;
                        const unsigned long long key = tc_second_pass(v1, v2, dcol_addr(tmem, acc, q4, xb), xb, np, tw, N, ntsum, Bflat, thr,
                                                                      rowidx + (uint32_t)xb, tcp);
                        mybest = key > mybest ? key : mybest;
                    }
                }
            }
            const long long c3b = clock64();
            // One candidate in a window that is this tile alone: it is the argmax whatever its exact score, and the
            // neighbourhood warp computes that score anyway (the centre of the 3x3) -- the workers score nothing.
            const bool deferred = do_nb && ncand == 1;
            if (ncand <= TC_MAXCAND && !deferred) {
                if (warp == 0) {          // the listed candidates, one per lane
                    unsigned long long key = 0ull;
                    if (lane < (int)ncand) {
                        const uint4 cd = s_cand[lane];
                        key = pack_key(tc_exact_score(cd.y, cd.z, cd.w, tcp), cd.x, false);
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                        key = other > key ? other : key;
                    }
                    mybest = key;
                }
            } else if (ncand > TC_MAXCAND) {
                for (int o = 16; o > 0; o >>= 1) {
                    unsigned long long other = __shfl_xor_sync(0xffffffffu, mybest, o);
                    mybest = other > mybest ? other : mybest;
                }
            }
            if (lane == 0) s_best[warp] = mybest;
            if (do_nb && tid == 0 && k >= 2) mbar_wait(&nb_free[k & 1], ((k >> 1) - 1) & 1);    // the stash of item k-2 has been read
            worker_bar();
            const long long c3c = clock64();
            clk_e2 += c3b - c3;
            clk_e3 += c3c - c3b;
            unsigned long long b = s_best[0];
#pragma unroll
            for (int w = 1; w < 8; w++) b = s_best[w] > b ? s_best[w] : b;
            if (deferred) b = (unsigned long long)(0xffffffffu - s_cand[0].x);      // index only; the score comes later
            if (tp.single_tile) {
                // The window is this tile alone: b is its final argmax, and what the 3x3 neighbourhood of the sub-pixel fit
                // (LST4:2681-2717) needs is still on chip.  The workers copy it into the stash -- the nine correlations from TMEM
                // (the warps that own rows by-1 .. by+1), rows by-1 .. by+1 of the vertical sums at columns bx-1 .. bx+tw -- and
                // go on; the neighbourhood warp scores the nine positions, so lst_epilogue does not recompute nine correlations.
                if (tid == 0 && !deferred) best[it.task] = b;
                if (do_nb) {
                    TcStash *st = &s_stash[k & 1];
                    const uint32_t idx = 0xffffffffu - (uint32_t)(b & 0xffffffffull);
                    int by = __float2int_rz(__fmul_rn((float)idx, __frcp_rn((float)it.rw)));      // idx < 2^24: off by one at most
                    int bx = (int)idx - by * it.rw;
                    if (bx >= it.rw) by++, bx -= it.rw;
                    if (bx < 0) by--, bx += it.rw;
                    if (tid == 0) {
                        st->task = it.task, st->tpl = it.tpl, st->by = by, st->bx = bx;
                        st->rw = it.rw, st->rh = it.rh_t, st->tw = tw, st->need_best = deferred ? 1 : 0;
                    }
                    if (!(tp.mode & 64) && hh == 0 && by + 1 >= 32 * q4 && by - 1 <= 32 * q4 + 31) {       // warp-uniform: this quarter holds some of the rows
                        uint32_t cv[3];
#pragma unroll
                        for (int d = 0; d < 3; d++) cv[d] = tmem_ld1(dcol_addr(tmem, acc, q4, min(max(bx - 1 + d, 0), TC_NB * TC_MAXNB - 1)));
                        const int dr = row - by + 1;
                        if (dr >= 0 && dr < 3) {
#pragma unroll
                            for (int d = 0; d < 3; d++) st->corr[3 * dr + d] = cv[d];
                        }
                    }
                    const int ncs = tw + 2, vmax = it.rw + tw - 1;
                    for (int e = tid; e < 3 * ncs && !(tp.mode & 32); e += TC_WORKERS) {
                        const int dr = e >= 2 * ncs ? 2 : (e >= ncs ? 1 : 0), c = e - dr * ncs;
                        const int r = by - 1 + dr, col = bx - 1 + c;
                        const bool ok = r >= 0 && r < it.rh_t && col >= 0 && col < vmax;
                        st->v1[dr][c] = ok ? V1[(size_t)r * P1 + col] : (uint16_t)0;
                        st->v2[dr][c] = ok ? V2[(size_t)r * P2 + col] : 0u;
                    }
                }
            } else if (tid == 0) {
                atomicMax(&best[it.task], b);
            }
            const long long c3d = clock64();
            tc_fence_before();
            worker_bar();
            const long long c3e = clock64();
            clk_e4 += c3d - c3c;
            clk_e5 += c3e - c3d;
            if (tid == 0) {
                mbar_arrive(&acc_free[acc]);
                if (do_nb) mbar_arrive(&nb_full[k & 1]);
            }
            const long long c4 = clock64();
            clk_wait += (c1 - c0) + (c3 - c2);
            clk_v += c2 - c1;
            clk_e += c4 - c3;
            k++;
        }
        if (tp.dbg && lane == 0) tp.dbg[(size_t)(1536 + blockIdx.x) * 8 + warp] = (unsigned long long)clk_p1;
        if (tp.dbg && tid == 0) {
            unsigned long long *d = tp.dbg + (size_t)blockIdx.x * 8;
            d[0] = (unsigned long long)clk_wait;
            d[1] = (unsigned long long)clk_v;
            d[2] = (unsigned long long)clk_e;
            d[3] = (unsigned long long)k;
            unsigned long long *d2 = tp.dbg + (size_t)(1024 + blockIdx.x) * 8;
            d2[0] = (unsigned long long)clk_e1;
            d2[1] = (unsigned long long)clk_e2;
            d2[2] = (unsigned long long)clk_e3;
            d2[3] = (unsigned long long)clk_e4;
            d2[4] = (unsigned long long)clk_e5;
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 10) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
    }
}

// ---- host side: tile plan, tensor maps, launch -----------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return (EncodeTiledFn)p;
    }();
    return fn;
}

bool encode_tensor_map(CUtensorMap *map, CUtensorMapDataType type, int rank, const void *base, const uint64_t *dims,
                       const uint64_t *strides_bytes, const uint32_t *box)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || rank < 1 || rank > 3) return false;
    cuuint64_t d[3], st[2];
    cuuint32_t b[3], es[3] = {1u, 1u, 1u};
    for (int i = 0; i < rank; i++) d[i] = dims[i], b[i] = box[i];
    for (int i = 0; i + 1 < rank; i++) st[i] = strides_bytes[i];
    return enc(map, type, (cuuint32_t)rank, (void *)base, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The pass's 8-bit windows as one 2-D byte image [total_rows][pitch]; box = {16 bytes, rows}.
static bool make_window_map(CUtensorMap *map, const uint8_t *win8, int pitch, long long total_rows, int box_rows)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)total_rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)win8, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static const size_t kTcSmemMax = 227 * 1024 - 9216;    // dynamic part; the static barriers, the two neighbourhood stashes and the 128-byte alignment take the rest

// Picks the tile shape and the shared-memory split for a pass; returns the dynamic shared-memory size, 0 = does not fit.
static size_t tc_plan(int max_rw, int max_rh, int tw, int th, TcParams *tp)
{
    static const int env_nbc = [] { const char *e = getenv("LST_TC_NBC"); return e ? atoi(e) : 0; }();
    static const int env_abuf = [] { const char *e = getenv("LST_TC_ABUF"); return e ? atoi(e) : 0; }();
    const int nd = tc_nd(tw);
    const int nb_total = (max_rw + TC_NB - 1) / TC_NB;
    tp->tile_rows = max_rh < TC_M ? max_rh : TC_M;
    tp->rows_alloc = (TC_M - 1 + th + 7) & ~7;     // chunk columns start 128-byte aligned (TMA destination)
    tp->rows_load = tp->tile_rows + th - 1;
    tp->rows_box1 = tp->rows_load < 256 ? tp->rows_load : 256;
    const size_t row_bytes = (size_t)nd * 1024;
    for (int nbc = nb_total < TC_MAXNB ? nb_total : TC_MAXNB; nbc >= 1; nbc = nbc > 2 ? 2 : nbc - 1) {
        if (env_nbc > 0 && nbc > env_nbc) continue;
        for (int nab = 2; nab >= 1; nab--) {
            if (env_abuf > 0 && nab > env_abuf) continue;
            // K steps (32 window columns each) a tile can need: the band of the widest template over the tile's valid columns
            const int nk = std::min(nbc + nd - 1, ((std::min(TC_NB * nbc, max_rw) + tw - 2) >> 5) + 1);
            tp->nbc = nbc;
            tp->kc_a = 2 * nk;
            tp->n_abuf = nab;
            tp->vcols = TC_NB * nbc + tw;
            tp->p2 = (tp->vcols + 3) & ~3;
            if ((tp->p2 & 7) != 4) tp->p2 += 4;
            tp->p1 = (tp->vcols + 7) & ~7;
            if ((tp->p1 & 15) != 8) tp->p1 += 8;
            const size_t a = (size_t)tp->kc_a * tp->rows_alloc * 16 * nab;
            const size_t v = (size_t)tp->tile_rows * ((size_t)tp->p2 * 4 + (size_t)tp->p1 * 2) + 16;
            if (a + v + 2 * TC_NISS * row_bytes > kTcSmemMax) continue;
            // ring: template rows per stage ~6 KB, as many stages as fit (the table streams from L2: ~1 us of latency to cover)
            static const int env_vb = [] { const char *e = getenv("LST_TC_VB"); return e ? atoi(e) : 0; }();
            int vb = (env_vb > 0 ? env_vb : (int)(6144 / row_bytes)) / TC_NISS * TC_NISS;
            if (vb < TC_NISS) vb = TC_NISS;
            int stages = (int)((kTcSmemMax - a - v) / (vb * row_bytes));
            if (stages < 2) {
                vb = TC_NISS;
                stages = (int)((kTcSmemMax - a - v) / (vb * row_bytes));
            }
            if (stages > TC_MAXSTAGES) stages = TC_MAXSTAGES;
            if (stages < 2) continue;
            tp->vb = vb;
            tp->stages = stages;
            tp->stage_bytes = (int)(vb * row_bytes);
            return a + (size_t)stages * tp->stage_bytes + v;
        }
    }
    return 0;
}

static std::mutex g_tc_mutex;
struct TcDeviceState {
    int sms = 0;
    bool attr_ok = false;
    unsigned long long *dbg = nullptr;
};
static TcDeviceState g_tc_dev[64];
static int g_tc_debug = -1;

void tc_init_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_tc_mutex);
    if (g_tc_debug < 0) g_tc_debug = getenv("LST_TC_DEBUG") ? 1 : 0;
    if (dev < 0 || dev >= 64) return;
    TcDeviceState &d = g_tc_dev[dev];
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(lst_ncc_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemMax);
    if (e != cudaSuccess) fprintf(stderr, "lst_ncc_tc: MaxDynamicSharedMemorySize: %s\n", cudaGetErrorString(e));
    d.attr_ok = e == cudaSuccess;
    e = cudaFuncSetAttribute(lst_ncc_tc, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) fprintf(stderr, "lst_ncc_tc: PreferredSharedMemoryCarveout: %s\n", cudaGetErrorString(e));
    if (g_tc_debug && !d.dbg) cudaMalloc((void **)&d.dbg, 2048 * 8 * sizeof(unsigned long long));
    cudaGetLastError();
}

// All windows of the pass lie in win8 with the same row pitch, window k at row win_off / pitch (run_group lays them
// out so); total_rows = rows of the whole buffer.  Returns the number of kernels launched, < 0 = not applicable.
int launch_ncc_tc(const DevTask *tasks, int n_tasks, int max_rw, int max_rh, int tw, int th, int pitch, long long total_rows,
                  const DevTemplates *tpls, const uint8_t *win8, unsigned long long *best, float *nbhd, cudaStream_t stream)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !g_tc_dev[dev].attr_ok) return -1;
    const TcDeviceState &ds = g_tc_dev[dev];
    TcParams tp;
    memset(&tp, 0, sizeof(tp));
    const size_t smem = tc_plan(max_rw, max_rh, tw, th, &tp);
    if (smem == 0) return -1;
    tp.tiles_x = ((max_rw + TC_NB - 1) / TC_NB + tp.nbc - 1) / tp.nbc;
    tp.tiles_y = (max_rh + tp.tile_rows - 1) / tp.tile_rows;
    tp.single_tile = tp.tiles_x * tp.tiles_y == 1 ? 1 : 0;
    const long long items = (long long)n_tasks * tp.tiles_x * tp.tiles_y;
    if (items > 0x7fffffffLL) return -1;
    tp.n_items = (int)items;
    tp.dbg = g_tc_debug ? ds.dbg : nullptr;
    {
        static const int mode = [] { const char *e = getenv("LST_TC_MODE"); return e ? atoi(e) : 0; }();
        tp.mode = mode;
    }
    CUtensorMap map1, map2;
    if (!make_window_map(&map1, win8, pitch, total_rows, tp.rows_box1)) return -1;
    const int rows2 = tp.rows_load - tp.rows_box1;
    if (!make_window_map(&map2, win8, pitch, total_rows, rows2 > 0 ? rows2 : 1)) return -1;
    const int grid = (int)(items < ds.sms ? items : ds.sms);
    if (grid < 1) return 0;
    if (tp.dbg) cudaMemsetAsync(tp.dbg, 0, (size_t)2048 * 64, stream);
    lst_ncc_tc<<<grid, TC_THREADS, smem, stream>>>(map1, map2, tasks, tpls, best, nbhd, tp);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "lst_ncc_tc launch failed: %s (grid %d, smem %zu)\n", cudaGetErrorString(e), grid, smem);
        return -1;
    }
    static int dbg_prints = 0;
    if (tp.dbg && n_tasks >= 1024 && dbg_prints++ < 2) {
        cudaStreamSynchronize(stream);
        static unsigned long long h[2048 * 8];
        cudaMemcpy(h, tp.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        double a[8] = {0}, e[5] = {0};
        for (int i = 0; i < grid; i++) {
            for (int j = 0; j < 8; j++) a[j] += (double)h[8 * i + j];
            for (int j = 0; j < 5; j++) e[j] += (double)h[8 * (1024 + i) + j];
        }
        const double k = a[3];
        {
            double w8[8] = {0};
            for (int i = 0; i < grid; i++)
                for (int j = 0; j < 8; j++) w8[j] += (double)h[8 * (1536 + i) + j];
            fprintf(stderr, "lst_ncc_tc first pass per worker warp (cycles per item): %.0f %.0f %.0f %.0f | %.0f %.0f %.0f %.0f\n", w8[0] / k, w8[1] / k,
                    w8[2] / k, w8[3] / k, w8[4] / k, w8[5] / k, w8[6] / k, w8[7] / k);
        }
        fprintf(stderr,
                "lst_ncc_tc (cycles per item, mean over %d CTAs, %.0f items): workers wait %.0f vsums %.0f epilogue %.0f | mma wait-window %.0f "
                "wait-acc %.0f wait-table %.0f total %.0f | nbc %d abuf %d stages %d x %d B smem %zu | epilogue warp 0: init %.0f, to end of chunks %.0f, "
                "barrier %.0f, stash %.0f, barrier %.0f\n",
                grid, k, a[0] / k, a[1] / k, a[2] / k, a[4] / k, a[5] / k, a[6] / k, a[7] / k, tp.nbc, tp.n_abuf, tp.stages, tp.stage_bytes,
                smem, e[0] / k, e[1] / k, e[2] / k, e[3] / k, e[4] / k);
    }
    return 1;
}

// ---- int8 tensor-core peak (roofline denominator of lst_ncc_tc) ------------------------------------
// One CTA per SM issues `iters` x 2 tcgen05.mma kind::i8 of M128 N256 K32 from (uninitialised) shared
// memory into two TMEM accumulators: the operand values do not matter for the rate.
__global__ void __launch_bounds__(128, 1) lst_peak_imma(int iters)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_t;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (4096 + 8192) / 16; i += blockDim.x) ((uint4 *)sm)[i] = make_uint4(0x01010101u, 0x02020202u, 0, 0x03030303u);
    fence_async_smem();
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_t)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_t;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc_u8(128, 256);
        const uint64_t a = smem_desc(smem_u32(sm), 128 * 16, 128), b = smem_desc(smem_u32(sm + 4096), 256 * 16, 128);
        for (int i = 0; i < iters; i++) {
            mma_i8(tmem, a, b, idesc, 1u);
            mma_i8(tmem + 256, a, b, idesc, 1u);
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

double measure_imma_peak(double ms_per_test)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 4000;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        lst_peak_imma<<<sms, 128, 4096 + 8192>>>(iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) {
            cudaGetLastError();
            best = 0.0;
            break;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double macs = (double)sms * iters * 2.0 * 128.0 * 256.0 * 32.0;
        if (rep > 0 && macs / (ms * 1e-3) > best) best = macs / (ms * 1e-3);
        if (rep == 0 && ms > 0) {
            double f = ms_per_test / ms;
            f = f > 32 ? 32 : (f < 0.25 ? 0.25 : f);
            iters = (int)(iters * f);
            if (iters < 200) iters = 200;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace lst
