// lst_ncc_tc.cu -- tensor-core engine of the template match: the exact integer cross-correlation as
// an implicit (Toeplitz) GEMM on tcgen05.mma kind::i8 (u8 x u8 -> s32, exact), accumulators in TMEM.
//
//   out[y][x] = sum_v sum_c W[y+v][c] * T[v][c-x]
//
// For a tile of 128 result rows x (up to 4 blocks of) 32 result columns:
//   A operand  = the 8-bit window in shared memory as [16-byte column chunk][row][16 B]: 8 consecutive
//                rows of one chunk are a 128-byte core matrix *for any starting row*, so "rows y+v" is
//                just a start address advanced by v*16 bytes in the shared-memory descriptor;
//   B operand  = for template row v the Toeplitz block B_v[n][k] = T[v][k-n] (32 x (31+tw) bytes), the
//                same for every 32-column block; all B_v of a template form a table in global memory
//                (built once per template) and stream through a 4-stage shared-memory ring with
//                cp.async.bulk (TMA engine) completing on mbarriers;
//   D          = 128 lanes x 32 columns of TMEM per column block, accumulated over v and the K steps.
// One elected thread issues the MMAs, one thread feeds the ring, four warps run the epilogue:
// window sums from vertical sliding sums, float32 filter, OpenCV's double-precision normalisation
// for the candidates, first-maximum argmax -- the same arithmetic as the IDP.4A engine
// (lst_ncc_match), so both engines return identical bits.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/lst_b200.h"
#include "lst_device.h"

namespace lst {

#define TC_M 128          // result rows per CTA (TMEM lanes)
#define TC_NB 32          // result columns per MMA (N)
#define TC_MAXNB 4        // column blocks per CTA (128 TMEM columns)
#define TC_STAGES 3
#define TC_THREADS 160    // warps 0-3 epilogue, warp 4: one thread feeds the ring (TMA) and issues the MMAs; TMEM alloc
#define TC_EPS 1e-5f

int tc_ksteps(int tw) { return (TC_NB - 1 + tw + 31) / 32; }
size_t tc_table_bytes(int tw, int th) { return (size_t)th * 2 * tc_ksteps(tw) * TC_NB * 16; }
bool tc_supported(int tw, int th) { return tw >= 1 && th >= 1 && tw <= 181 && th <= 181; }   // s32 accumulators

void tc_build_table(const uint8_t *tpl8, int tw, int th, uint8_t *out)
{
    const int kc2 = 2 * tc_ksteps(tw);
    for (int v = 0; v < th; v++)
        for (int kc = 0; kc < kc2; kc++)
            for (int n = 0; n < TC_NB; n++)
                for (int j = 0; j < 16; j++) {
                    int u = 16 * kc + j - n;
                    out[(((size_t)v * kc2 + kc) * TC_NB + n) * 16 + j] = (u >= 0 && u < tw) ? tpl8[v * tw + u] : 0;
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
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1):
// 8 rows x 16 B core matrices; LBO = bytes between the two K halves, SBO = bytes between 8-row groups
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = u8, K-major, N = 32, M = 128
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- window sums for the tensor-core engine ------------------------------------------------------
// The IDP.4A engine derives the sums of I and I^2 under the template footprint inside its CTAs; the
// tensor-core engine wants its shared memory for operands and its warps for the epilogue, so the
// sums are produced once per window by this kernel and read back (L2) by the epilogue.
// CTA = 32 result rows of one window: vertical sliding sums per column into shared memory, then
// horizontal sliding sums with one row per lane.
#define BS_ROWS 32

__global__ void __launch_bounds__(256)
lst_box_sums(const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls, const uint8_t *__restrict__ win8,
             uint32_t *__restrict__ wsum, uint32_t *__restrict__ wsq, int wp)
{
    extern __shared__ __align__(16) uint8_t smem_bs[];
    const DevTask t = tasks[blockIdx.y];
    const int tw = tpls->c[t.tpl].tw, th = tpls->c[t.tpl].th;
    const int y0 = blockIdx.x * BS_ROWS;
    if (y0 >= t.rh) return;
    const int nrows = min(BS_ROWS, t.rh - y0);
    const int trows = nrows + th - 1;
    uint32_t *P2 = (uint32_t *)smem_bs;                 // [BS_ROWS][wp]  sums of squares, then prefix
    uint32_t *P1 = P2 + (size_t)BS_ROWS * wp;           // [BS_ROWS][wp]  sums, then prefix
    uint8_t *tile = (uint8_t *)(P1 + (size_t)BS_ROWS * wp);   // [trows][pitch]
    const int pitch = t.pitch;
    {
        const uint8_t *src = win8 + t.win_off + (size_t)y0 * pitch;
        const int vec = pitch / 16;
        for (int i = threadIdx.x; i < trows * vec; i += blockDim.x) {
            int r = i / vec, c = (i - r * vec) * 16;
            *(uint4 *)(tile + (size_t)r * pitch + c) = *(const uint4 *)(src + (size_t)r * pitch + c);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < t.w; c += blockDim.x) {
        uint32_t s = 0, q = 0;
        for (int v = 0; v < th; v++) {
            uint32_t p = tile[(size_t)v * pitch + c];
            s += p;
            q += p * p;
        }
        P1[c] = s;
        P2[c] = q;
        for (int r = 1; r < nrows; r++) {
            uint32_t a = tile[(size_t)(r + th - 1) * pitch + c], b = tile[(size_t)(r - 1) * pitch + c];
            s += a - b;
            q += a * a - b * b;
            P1[(size_t)r * wp + c] = s;
            P2[(size_t)r * wp + c] = q;
        }
    }
    __syncthreads();
    // Horizontal pass: lanes = the rows of the band, each warp a run of result columns.  A lane sums the
    // first footprint of its run directly (tw loads) and slides from there: out[x+1] = out[x] + V[x+tw] - V[x].
    // Rows are wp words apart (odd: conflict-free); a warp's store is 32 consecutive words of the
    // output layout [band of 32 rows][x][row in band] -- the epilogue of lst_ncc_tc owns one row per
    // lane and walks x, so its loads are 128 bytes per warp as well.
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        const int per = (t.rw + nwarps - 1) / nwarps;
        const int x0 = warp * per, x1 = min(x0 + per, t.rw);
        if (lane < nrows && x0 < x1) {
            const uint32_t *p1 = P1 + (size_t)lane * wp, *p2 = P2 + (size_t)lane * wp;
            uint32_t a1 = 0, a2 = 0;
#pragma unroll 8
            for (int u = 0; u < tw; u++) {
                a1 += p1[x0 + u];
                a2 += p2[x0 + u];
            }
            uint32_t *o1 = wsum + t.sum_off + (size_t)blockIdx.x * t.rw * BS_ROWS + lane;
            uint32_t *o2 = wsq + t.sum_off + (size_t)blockIdx.x * t.rw * BS_ROWS + lane;
            for (int x = x0;; x++) {
                o1[(size_t)x * BS_ROWS] = a1;
                o2[(size_t)x * BS_ROWS] = a2;
                if (x + 1 >= x1) break;
                a1 += p1[x + tw] - p1[x];
                a2 += p2[x + tw] - p2[x];
            }
        }
    }
}

// The exact (double-precision, OpenCV order) score of one candidate position, out of line: the
// epilogue visits 32 positions with unrolled code and only a few of them get here, so the long
// division / square-root sequences must not be replicated 32 times (instruction cache).
__device__ __noinline__ float tc_exact_score(uint32_t corr, const uint32_t *ws, const uint32_t *wq, const TplConst *tc)
{
    const uint32_t s = __ldg(ws), sq = __ldg(wq);
    return match_score(*tc, (double)corr, (double)s, (double)sq);
}
__device__ __forceinline__ unsigned long long tc_exact_key(uint32_t corr, const uint32_t *ws, const uint32_t *wq,
                                                           const TplConst *tc, uint32_t idx)
{
    return pack_key(tc_exact_score(corr, ws, wq, tc), idx, method_wants_min(tc->method));
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr)
{
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return r;
}

struct TcParams {
    int rows_a;       // window rows staged per CTA: TC_M + max_th - 1
    int kc_a;         // 16-byte column chunks staged per CTA
    int ks;           // K steps (32 bytes) per column block
    int nb_per_cta;   // column blocks per CTA
    int vb;           // template rows (Toeplitz blocks) per ring stage
    int group;        // column blocks per MMA group: the epilogue of a group overlaps the MMAs of the next
    unsigned long long *dbg;   // optional: per-CTA phase clocks (LST_TC_DEBUG)
};

__global__ void __launch_bounds__(TC_THREADS, 4)
lst_ncc_tc(const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls, const uint8_t *__restrict__ win8,
           const uint32_t *__restrict__ wsum, const uint32_t *__restrict__ wsq, unsigned long long *__restrict__ best,
           float *__restrict__ nbhd, TcParams tp)
{
    extern __shared__ __align__(128) uint8_t smem_tc[];
    __shared__ __align__(8) uint64_t bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_done[TC_MAXNB];
    __shared__ uint32_t s_tmem;
    __shared__ unsigned long long s_best[4];
    __shared__ uint32_t s_fmax;    // orderable key of the largest filter score any warp of the CTA has seen so far

    const long long clk0 = clock64();
    const DevTask t = tasks[blockIdx.y];
    const TplConst tc = tpls->c[t.tpl];
    const int tw = tc.tw, th = tc.th;
    const int nb_total = (t.rw + TC_NB - 1) / TC_NB;
    const int xchunks = (nb_total + tp.nb_per_cta - 1) / tp.nb_per_cta;
    const int ytiles = (t.rh + TC_M - 1) / TC_M;
    if ((int)blockIdx.x >= xchunks * ytiles) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (tc.flat && tc.method == 5) {   // templSdv^2 < DBL_EPSILON: OpenCV fills the result with 1 -> first position wins
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(&best[blockIdx.y], pack_key(1.0f, 0u));
        return;
    }
    const int xc = blockIdx.x % xchunks, yt = blockIdx.x / xchunks;
    const int nb0 = xc * tp.nb_per_cta;
    const int nbc = min(tp.nb_per_cta, nb_total - nb0);     // column blocks of this CTA
    const int x0 = nb0 * TC_NB, y0 = yt * TC_M;
    const int ROWS = tp.rows_a, KCA = tp.kc_a, KS = tp.ks;
    const uint32_t bv_bytes = (uint32_t)(2 * KS * TC_NB * 16);   // one Toeplitz block B_v
    const int VB = tp.vb;                                        // blocks per ring stage
    const uint32_t st_bytes = bv_bytes * (uint32_t)VB;
    const int nsteps = (th + VB - 1) / VB;

    uint8_t *sA = smem_tc;                                          // [KCA][ROWS][16]
    uint8_t *sB = sA + (size_t)KCA * ROWS * 16;                     // [TC_STAGES][VB][2*KS][32][16]

    // ---- stage the window: rows [y0, y0 + ROWS), byte columns [x0, x0 + 16*KCA), zero filled ----
    {
        const uint8_t *src = win8 + t.win_off;
        const int total = KCA * ROWS;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int kc = i / ROWS, r = i - kc * ROWS;      // consecutive threads -> consecutive rows: conflict-free stores
            const int gx = x0 + 16 * kc, gy = y0 + r;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (gy < t.h && gx < t.pitch) v = *(const uint4 *)(src + (size_t)gy * t.pitch + gx);
            *(uint4 *)(sA + ((size_t)kc * ROWS + r) * 16) = v;
        }
        fence_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        s_fmax = 0u;
        for (int g = 0; g < TC_MAXNB; g++) mbar_init(&bar_done[g], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const long long clk1 = clock64();
    long long clk2 = clk1, clk3 = clk1;

    if (warp == 4) {
        // ---- MMA issuer ----
        // One thread issues everything, so the loop body must be a handful of instructions: the
        // descriptors are 64-bit integers whose low field is (address >> 4), advanced by plain adds.
        if (lane == 0) {
            const uint32_t idesc = instr_desc_u8(TC_M, TC_NB);
            const uint64_t a_desc0 = smem_desc(smem_u32(sA), (uint32_t)ROWS * 16, 128);
            const uint64_t b_desc0 = smem_desc(smem_u32(sB), TC_NB * 16, 128);
            const uint64_t a_kstep = (uint64_t)(2 * ROWS);            // two 16-byte column chunks, in 16-byte units
            const uint64_t b_kstep = (uint64_t)(2 * TC_NB);
            const uint64_t b_vstep = (uint64_t)(bv_bytes >> 4), b_sstep = (uint64_t)(st_bytes >> 4);
            // ring producer and MMA issue in one thread: the copy for step st+1 goes out before the MMAs
            // of step st; its slot was last read by the MMAs of step st-2 (3 stages), long finished
            const uint8_t *table = tpls->toeplitz[t.tpl];
            // The column blocks are worked off in groups of tp.group: all template rows for the blocks of group
            // g, then the next group (the Toeplitz table streams through the ring once per group), so that the
            // epilogue warps score group g while the tensor core is busy with group g + 1.
            const int G = tp.group, ngroups = (nbc + G - 1) / G, total = ngroups * nsteps;
            auto feed = [&](int gst) {                  // gst = global ring step: group * nsteps + step
                const int s = gst % TC_STAGES, st = gst % nsteps;
                const int nv = min(VB, th - st * VB);
                if (gst >= TC_STAGES) mbar_wait(&bar_empty[s], ((gst / TC_STAGES) - 1) & 1);
                mbar_expect_tx(&bar_full[s], bv_bytes * (uint32_t)nv);
                bulk_g2s(sB + (size_t)s * st_bytes, table + (size_t)st * st_bytes, bv_bytes * (uint32_t)nv, &bar_full[s]);
            };
            feed(0);
            int gst = 0;
            for (int g = 0; g < ngroups; g++) {
                const int nbg0 = g * G, nbg1 = min(nbc, nbg0 + G);
                uint32_t acc = 0;
                int v = 0;
                for (int st = 0; st < nsteps; st++, gst++) {
                    const int s = gst % TC_STAGES;
                    const int nv = min(VB, th - st * VB);
                    if (gst + 1 < total) feed(gst + 1);
                    mbar_wait(&bar_full[s], (gst / TC_STAGES) & 1);
                    tc_fence_after();
                    uint64_t b_v = b_desc0 + (uint64_t)s * b_sstep;
                    for (int vv = 0; vv < nv; vv++, v++, b_v += b_vstep) {
                        uint64_t a_nb = a_desc0 + (uint64_t)v + (uint64_t)nbg0 * a_kstep;            // rows v .. v+127
                        for (int nb = nbg0; nb < nbg1; nb++, a_nb += a_kstep) {
                            uint64_t a_d = a_nb, b_d = b_v;
                            const uint32_t d = tmem + (uint32_t)(nb * TC_NB);
                            mma_i8(d, a_d, b_d, idesc, acc);              // first k-step: overwrites D only for v == 0
#pragma unroll 1
                            for (int ks = 1; ks < KS; ks++) {
                                a_d += a_kstep;
                                b_d += b_kstep;
                                mma_i8(d, a_d, b_d, idesc, 1u);
                            }
                        }
                        acc = 1u;
                    }
                    tc_commit(&bar_empty[s]);     // the ring slot is free once these MMAs have read it
                }
                tc_commit(&bar_done[g]);
            }
        }
    } else {
        // ---- epilogue warps: rows 32*warp .. +31 of the tile (TMEM lanes of this warp) ----
        const int yy = 32 * warp + lane;                  // row of the tile owned by this lane
        const int y = y0 + yy;
        const bool row_ok = y < t.rh;
        const uint32_t N = (uint32_t)(tw * th), tsum = tc.tsum;
        const long long Nflat = (long long)N + (long long)(N >> 5);
        // sums of this lane's row: [band][x][row in band] (see lst_box_sums), band = y / 32
        // (lanes whose row lies below the result map read the task's first row instead: their band may not
        // exist in the sums buffer, and the last task's would lie beyond the allocation)
        const size_t rowoff = row_ok ? t.sum_off + (size_t)(y0 / 32 + warp) * t.rw * 32 + lane : t.sum_off;
        unsigned long long mybest = 0ull;
        // While the tensor core works, pull this warp's window sums (written by lst_box_sums, long
        // evicted from L2 by the time they are read) back into L2: one 128-byte line per (array, column).
        if (y0 + 32 * warp < t.rh) {
            const int ncols = min(nbc * TC_NB, t.rw - x0);
            const size_t line0 = t.sum_off + (size_t)(y0 / 32 + warp) * t.rw * 32 + (size_t)x0 * 32;
            for (int c = lane; c < 2 * ncols; c += 32) {
                const uint32_t *pf = (c < ncols ? wsum : wsq) + line0 + (size_t)(c < ncols ? c : c - ncols) * 32;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            }
        }
        for (int nb = 0; nb < nbc; nb++) {
            if (nb % tp.group == 0) {
                mbar_wait(&bar_done[nb / tp.group], 0);
                tc_fence_after();
                if (nb == 0) clk2 = clock64();
            }
            uint32_t corr[32];
            tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(nb * TC_NB), corr);
            const int xb = x0 + TC_NB * nb;
            const int np = row_ok ? min(32, t.rw - xb) : 0;
            const int xs = xb < t.rw ? xb : 0;     // a block right of this (smaller) window's map: stay inside the buffer
            const uint32_t *ws = wsum + rowoff + (size_t)xs * 32, *wq = wsq + rowoff + (size_t)xs * 32;
            // pass 1: float32 filter score of the 32 positions of this lane's row (all loads of a half
            // are issued before they are used), warp-wide maximum.  3e38 marks "nearly flat: always exact".
            float sf[32];
            float tmax = -3.0e38f;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t sv[16], qv[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int p = 16 * h + i;
                    const int pc = p < np ? p : 0;            // clamped: the load is always legal
                    sv[i] = __ldg(ws + pc * 32);
                    qv[i] = __ldg(wq + pc * 32);
                }
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int p = 16 * h + i;
                    // 32 x 32 -> 64-bit products (IMAD.WIDE.U32): every factor fits 32 bits
                    const long long A = (long long)((unsigned long long)N * corr[p]) - (long long)((unsigned long long)sv[i] * tsum);
                    const long long B = (long long)((unsigned long long)N * qv[i]) - (long long)((unsigned long long)sv[i] * sv[i]);
                    float f;
                    if (p >= np) f = -3.0e38f;
                    else if (2 * B <= Nflat || tc.method != 5) f = 3.0e38f;   // other methods: every position exact
                    else {
                        f = (float)A * rsqrtf((float)B) * tc.inv_sqrt_c;
                        tmax = fmaxf(tmax, f);
                    }
                    sf[p] = f;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
            // pass 2: OpenCV's double-precision score for the candidates only.  A candidate must come within
            // TC_EPS of the largest filter score seen anywhere in the CTA so far (every position within TC_EPS
            // of the final maximum passes this, whenever it is visited), not just of this block's.
            const uint32_t rowidx = (uint32_t)y * (uint32_t)t.rw;
            if (lane == 0 && tmax > -3.0e38f) atomicMax(&s_fmax, f32_key(tmax));
            __syncwarp();
            const float thr = fmaxf(tmax, key_f32(*(volatile uint32_t *)&s_fmax)) - TC_EPS;
#pragma unroll
            for (int p = 0; p < 32; p++) {
                if (sf[p] >= thr && p < np) {
                    const unsigned long long key =
                        tc_exact_key(corr[p], ws + p * 32, wq + p * 32, &tpls->c[t.tpl], rowidx + (uint32_t)(xb + p));
                    mybest = key > mybest ? key : mybest;
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, mybest, o);
            mybest = other > mybest ? other : mybest;
        }
        if (lane == 0) s_best[warp] = mybest;
        clk3 = clock64();
        if (tp.dbg && threadIdx.x == 0 && blockIdx.y < 4096) {
            unsigned long long *d = tp.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 4;
            d[0] = (unsigned long long)(clk1 - clk0);
            d[1] = (unsigned long long)(clk2 - clk1);
            d[2] = (unsigned long long)(clk3 - clk2);
            d[3] = (unsigned long long)clk0;
        }
    }
    tc_fence_before();
    __syncthreads();
    unsigned long long b = s_best[0];
    for (int w = 1; w < 4; w++) b = s_best[w] > b ? s_best[w] : b;
    if (threadIdx.x == 0) atomicMax(&best[blockIdx.y], b);
    if (nbhd && gridDim.x == 1) {
        // The window is this CTA's alone, so b is its final argmax and the 3x3 neighbourhood the sub-pixel
        // fit needs (LST4:2681-2717) is still in TMEM: the lanes that own rows by-1..by+1 score the three
        // columns bx-1..bx+1 exactly, and lst_epilogue does not have to recompute nine correlations.
        if (warp < 4) {
            tc_fence_after();
            const uint32_t idx = 0xffffffffu - (uint32_t)(b & 0xffffffffull);
            const int by = (int)(idx / (uint32_t)t.rw), bx = (int)(idx - (uint32_t)by * (uint32_t)t.rw);
            const int y = 32 * warp + lane, dyi = y - by;
            const size_t rowoff = t.sum_off + (size_t)warp * t.rw * 32 + lane;
            float *o = nbhd + (size_t)blockIdx.y * 12;
            for (int dxi = -1; dxi <= 1; dxi++) {
                const int x = bx + dxi;
                if (x < 0 || x >= t.rw) continue;
                const uint32_t cv = tmem_ld1(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)x);
                if (dyi >= -1 && dyi <= 1 && y < t.rh)
                    o[(dyi + 1) * 3 + (dxi + 1)] = tc_exact_score(cv, wsum + rowoff + (size_t)x * 32, wsq + rowoff + (size_t)x * 32,
                                                                  &tpls->c[t.tpl]);
            }
            if (threadIdx.x == 0) o[9] = 1.0f;
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
    }
}

static size_t tc_smem(int tw, int th, TcParams *tp, int max_rw)
{
    const int nb_total = (max_rw + TC_NB - 1) / TC_NB;
    tp->nb_per_cta = nb_total < TC_MAXNB ? nb_total : TC_MAXNB;
    tp->ks = tc_ksteps(tw);
    tp->rows_a = TC_M + th - 1;
    tp->kc_a = 2 * (tp->nb_per_cta - 1) + 2 * tp->ks;
    const size_t bv = (size_t)2 * tp->ks * TC_NB * 16;
    static const int group_env = [] {
        const char *e = getenv("LST_TC_GROUP");
        return e ? atoi(e) : 0;
    }();
    tp->group = group_env > 0 ? group_env : 2;
    if (tp->group > tp->nb_per_cta) tp->group = tp->nb_per_cta;
    tp->vb = (int)(6144 / bv);                        // ~6 KB per ring stage: 3 stages + the window stay under 56 KB
    if (tp->vb < 1) tp->vb = 1;
    if (tp->vb > th) tp->vb = th;
    return (size_t)tp->kc_a * tp->rows_a * 16 + (size_t)TC_STAGES * tp->vb * bv + 64;
}

void tc_init_device()
{
    cudaError_t e = cudaFuncSetAttribute(lst_ncc_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256);
    if (e != cudaSuccess) fprintf(stderr, "lst_ncc_tc: MaxDynamicSharedMemorySize: %s\n", cudaGetErrorString(e));
    cudaFuncSetAttribute(lst_ncc_tc, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(lst_box_sums, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256);
    cudaFuncSetAttribute(lst_box_sums, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
}

int launch_box_sums(const DevTask *tasks, int n_tasks, int max_w, int max_rh, int th, const DevTemplates *tpls,
                    const uint8_t *win8, uint32_t *wsum, uint32_t *wsq, cudaStream_t stream)
{
    const int wp = ((max_w + 15) & ~15) | 1;
    const size_t bs_smem = (size_t)2 * BS_ROWS * wp * 4 + (size_t)(BS_ROWS + th - 1) * ((max_w + 15) & ~15) + 16;
    if (bs_smem > 226 * 1024) return -1;
    int launches = 0;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 bgrid((max_rh + BS_ROWS - 1) / BS_ROWS, n);
        lst_box_sums<<<bgrid, 256, bs_smem, stream>>>(tasks + base, tpls, win8, wsum, wsq, wp);
        launches++;
    }
    return launches;
}

int launch_ncc_tc(const DevTask *tasks, int n_tasks, int max_rw, int max_rh, int tw, int th,
                  const DevTemplates *tpls, const uint8_t *win8, const uint32_t *wsum, const uint32_t *wsq,
                  unsigned long long *best, float *nbhd, cudaStream_t stream)
{
    TcParams tp;
    size_t smem = tc_smem(tw, th, &tp, max_rw);
    if (smem > 226 * 1024) return -1;
    static unsigned long long *dbg = nullptr;
    static int dbg_on = -1;
    if (dbg_on < 0) {
        dbg_on = getenv("LST_TC_DEBUG") ? 1 : 0;
        if (dbg_on) cudaMalloc((void **)&dbg, 4096 * 16 * 4 * sizeof(unsigned long long));
    }
    tp.dbg = dbg_on ? dbg : nullptr;
    const int nb_total = (max_rw + TC_NB - 1) / TC_NB;
    const int xchunks = (nb_total + tp.nb_per_cta - 1) / tp.nb_per_cta;
    const int ytiles = (max_rh + TC_M - 1) / TC_M;
    int launches = 0;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 grid(xchunks * ytiles, n);
        lst_ncc_tc<<<grid, TC_THREADS, smem, stream>>>(tasks + base, tpls, win8, wsum, wsq, best + base,
                                                       nbhd ? nbhd + (size_t)base * 12 : nullptr, tp);
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess)
            fprintf(stderr, "lst_ncc_tc launch failed: %s (grid %d x %d, smem %zu)\n", cudaGetErrorString(e), grid.x, grid.y, smem);
        launches++;
    }
    if (dbg_on && n_tasks >= 1024) {
        cudaStreamSynchronize(stream);
        const int ncta = xchunks * ytiles * (n_tasks < 4096 ? n_tasks : 4096);
        unsigned long long *h = (unsigned long long *)malloc((size_t)ncta * 32);
        cudaMemcpy(h, dbg, (size_t)ncta * 32, cudaMemcpyDeviceToHost);
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < ncta; i++) {
            a += (double)h[4 * i];
            b += (double)h[4 * i + 1];
            c += (double)h[4 * i + 2];
        }
        fprintf(stderr, "lst_ncc_tc phases (cycles, mean over %d CTAs): stage %.0f  mma-wait %.0f  epilogue %.0f  smem %zu\n", ncta,
                a / ncta, b / ncta, c / ncta, smem);
        free(h);
    }
    return launches;
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
