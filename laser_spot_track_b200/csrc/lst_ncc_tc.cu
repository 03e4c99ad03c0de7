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
#include <string.h>

#include "../../include/lst_b200.h"
#include "lst_device.h"

namespace lst {

#define TC_M 128          // result rows per CTA (TMEM lanes)
#define TC_NB 32          // result columns per MMA (N)
#define TC_MAXNB 4        // column blocks per CTA (128 TMEM columns)
#define TC_STAGES 4
#define TC_THREADS 192    // warps 0-3 epilogue, warp 4 MMA issue + TMEM alloc, warp 5 TMA producer
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

struct TcParams {
    int rows_a;       // window rows staged per CTA: TC_M + max_th - 1
    int kc_a;         // 16-byte column chunks staged per CTA
    int ks;           // K steps (32 bytes) per column block
    int vp;           // pitch of the per-warp vertical-sum scratch (odd)
    int nb_per_cta;   // column blocks per CTA
};

__global__ void __launch_bounds__(TC_THREADS, 1)
lst_ncc_tc(const DevTask *__restrict__ tasks, const DevTemplates *__restrict__ tpls, const uint8_t *__restrict__ win8,
           unsigned long long *__restrict__ best, TcParams tp)
{
    extern __shared__ __align__(128) uint8_t smem_tc[];
    __shared__ __align__(8) uint64_t bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_done;
    __shared__ uint32_t s_tmem;
    __shared__ unsigned long long s_best[4];

    const DevTask t = tasks[blockIdx.y];
    const TplConst tc = tpls->c[t.tpl];
    const int tw = tc.tw, th = tc.th;
    const int nb_total = (t.rw + TC_NB - 1) / TC_NB;
    const int xchunks = (nb_total + tp.nb_per_cta - 1) / tp.nb_per_cta;
    const int ytiles = (t.rh + TC_M - 1) / TC_M;
    if ((int)blockIdx.x >= xchunks * ytiles) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (tc.flat) {   // templSdv^2 < DBL_EPSILON: OpenCV fills the result with 1 -> first position wins
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(&best[blockIdx.y], pack_key(1.0f, 0u));
        return;
    }
    const int xc = blockIdx.x % xchunks, yt = blockIdx.x / xchunks;
    const int nb0 = xc * tp.nb_per_cta;
    const int nbc = min(tp.nb_per_cta, nb_total - nb0);     // column blocks of this CTA
    const int x0 = nb0 * TC_NB, y0 = yt * TC_M;
    const int ROWS = tp.rows_a, KCA = tp.kc_a, KS = tp.ks;
    const uint32_t bv_bytes = (uint32_t)(2 * KS * TC_NB * 16);

    uint8_t *sA = smem_tc;                                          // [KCA][ROWS][16]
    uint8_t *sB = sA + (size_t)KCA * ROWS * 16;                     // [TC_STAGES][2*KS][32][16]
    uint8_t *sV = sB + (size_t)TC_STAGES * bv_bytes;                // 4 warps x ([32][vp] u32 + [32][vp] u16)

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
        mbar_init(&bar_done, 1);
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

    if (warp == 5) {
        // ---- TMA producer: stream the Toeplitz blocks B_v through the ring ----
        if (lane == 0) {
            const uint8_t *table = tpls->toeplitz[t.tpl];
            for (int v = 0; v < th; v++) {
                const int s = v % TC_STAGES;
                if (v >= TC_STAGES) mbar_wait(&bar_empty[s], ((v / TC_STAGES) - 1) & 1);
                mbar_expect_tx(&bar_full[s], bv_bytes);
                bulk_g2s(sB + (size_t)s * bv_bytes, table + (size_t)v * bv_bytes, bv_bytes, &bar_full[s]);
            }
        }
    } else if (warp == 4) {
        // ---- MMA issuer ----
        if (lane == 0) {
            const uint32_t idesc = instr_desc_u8(TC_M, TC_NB);
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            for (int v = 0; v < th; v++) {
                const int s = v % TC_STAGES;
                mbar_wait(&bar_full[s], (v / TC_STAGES) & 1);
                tc_fence_after();
                for (int nb = 0; nb < nbc; nb++) {
                    for (int ks = 0; ks < KS; ks++) {
                        // window columns 32*nb + 32*ks .. +31 of rows v .. v+127  x  B_v k-step ks
                        const uint32_t a_addr = a_base + (uint32_t)(((2 * nb + 2 * ks) * ROWS + v) * 16);
                        const uint32_t b_addr = b_base + (uint32_t)(s * bv_bytes + (2 * ks) * TC_NB * 16);
                        mma_i8(tmem + (uint32_t)(nb * TC_NB), smem_desc(a_addr, (uint32_t)ROWS * 16, 128),
                               smem_desc(b_addr, TC_NB * 16, 128), idesc, (v > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                tc_commit(&bar_empty[s]);     // the ring slot is free once these MMAs have read it
            }
            tc_commit(&bar_done);
        }
    } else {
        // ---- epilogue warps: rows 32*warp .. +31 of the tile ----
        const int vp = tp.vp;
        uint32_t *v2 = (uint32_t *)(sV + (size_t)warp * ((size_t)32 * vp * 6));
        uint16_t *v1 = (uint16_t *)(v2 + (size_t)32 * vp);
        const int yy = 32 * warp + lane;                  // row of the tile owned by this lane
        const int y = y0 + yy;
        const bool row_ok = y < t.rh;
        const long long N = (long long)tw * th, tsum = (long long)tc.tsum;
        unsigned long long mybest = 0ull;
        bool waited = false;
        for (int nb = 0; nb < nbc; nb++) {
            // vertical sliding sums of I and I^2 for rows 32*warp..+31, window columns 32*nb .. 32*nb + 31 + tw - 1
            const int ncols = TC_NB + tw - 1;
            __syncwarp();
            for (int cc = lane; cc < ncols; cc += 32) {
                const int col = TC_NB * nb + cc;          // byte column inside the staged tile
                const uint8_t *colp = sA + ((size_t)(col >> 4) * ROWS) * 16 + (col & 15);
                uint32_t s = 0, q = 0;
                const int r0 = 32 * warp;
                for (int v = 0; v < th; v++) {
                    uint32_t p = colp[(size_t)(r0 + v) * 16];
                    s += p;
                    q += p * p;
                }
                v1[cc] = (uint16_t)s;
                v2[cc] = q;
                for (int r = 1; r < 32; r++) {
                    uint32_t a = colp[(size_t)(r0 + r + th - 1) * 16], b = colp[(size_t)(r0 + r - 1) * 16];
                    s += a - b;
                    q += a * a - b * b;
                    v1[r * vp + cc] = (uint16_t)s;
                    v2[r * vp + cc] = q;
                }
            }
            __syncwarp();
            if (!waited) {
                mbar_wait(&bar_done, 0);
                tc_fence_after();
                waited = true;
            }
            uint32_t corr[32];
            tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(nb * TC_NB), corr);
            const int xb = x0 + TC_NB * nb;
            const int np = row_ok ? min(32, t.rw - xb) : 0;
            const uint16_t *vr = v1 + lane * vp;
            const uint32_t *vr2 = v2 + lane * vp;
            uint32_t s0 = 0, sq0 = 0;
            if (np > 0)
                for (int u = 0; u < tw; u++) {
                    s0 += vr[u];
                    sq0 += vr2[u];
                }
            // pass 1: float32 filter, warp-wide maximum (see lst_ncc_match)
            float tmax = -3.0e38f;
            {
                uint32_t s = s0, sq = sq0;
#pragma unroll
                for (int p = 0; p < 32; p++) {
                    if (p < np) {
                        const long long A = N * (long long)corr[p] - (long long)s * tsum;
                        const long long B = N * (long long)sq - (long long)s * (long long)s;
                        if (2 * B > N + (N >> 5)) tmax = fmaxf(tmax, (float)A * rsqrtf((float)B) * tc.inv_sqrt_c);
                        s += (uint32_t)vr[p + tw] - (uint32_t)vr[p];
                        sq += vr2[p + tw] - vr2[p];
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
            // pass 2: exact double-precision score of the candidates
            {
                uint32_t s = s0, sq = sq0;
                const uint32_t rowidx = (uint32_t)y * (uint32_t)t.rw;
#pragma unroll
                for (int p = 0; p < 32; p++) {
                    if (p < np) {
                        const long long A = N * (long long)corr[p] - (long long)s * tsum;
                        const long long B = N * (long long)sq - (long long)s * (long long)s;
                        bool exact;
                        if (2 * B <= N + (N >> 5)) exact = true;
                        else exact = (float)A * rsqrtf((float)B) * tc.inv_sqrt_c >= tmax - TC_EPS;
                        if (exact) {
                            const float sc = ncc_score((double)corr[p], (double)s, (double)sq, tc.inv_area, tc.mean, tc.norm, 0);
                            const unsigned long long key = pack_key(sc, rowidx + (uint32_t)(xb + p));
                            mybest = key > mybest ? key : mybest;
                        }
                        s += (uint32_t)vr[p + tw] - (uint32_t)vr[p];
                        sq += vr2[p + tw] - vr2[p];
                    }
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, mybest, o);
            mybest = other > mybest ? other : mybest;
        }
        if (lane == 0) s_best[warp] = mybest;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long b = s_best[0];
        for (int w = 1; w < 4; w++) b = s_best[w] > b ? s_best[w] : b;
        atomicMax(&best[blockIdx.y], b);
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
    tp->vp = (TC_NB + tw) | 1;
    return (size_t)tp->kc_a * tp->rows_a * 16 + (size_t)TC_STAGES * 2 * tp->ks * TC_NB * 16 + (size_t)4 * 32 * tp->vp * 6 + 64;
}

void tc_init_device()
{
    cudaError_t e = cudaFuncSetAttribute(lst_ncc_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256);
    if (e != cudaSuccess) fprintf(stderr, "lst_ncc_tc: MaxDynamicSharedMemorySize: %s\n", cudaGetErrorString(e));
    cudaFuncSetAttribute(lst_ncc_tc, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
}

int launch_ncc_tc(const DevTask *tasks, int n_tasks, int max_rw, int max_rh, int tw, int th,
                  const DevTemplates *tpls, const uint8_t *win8, unsigned long long *best, cudaStream_t stream)
{
    TcParams tp;
    size_t smem = tc_smem(tw, th, &tp, max_rw);
    if (smem > 227 * 1024) return -1;
    const int nb_total = (max_rw + TC_NB - 1) / TC_NB;
    const int xchunks = (nb_total + tp.nb_per_cta - 1) / tp.nb_per_cta;
    const int ytiles = (max_rh + TC_M - 1) / TC_M;
    int launches = 0;
    for (int base = 0; base < n_tasks; base += 65535) {
        int n = n_tasks - base < 65535 ? n_tasks - base : 65535;
        dim3 grid(xchunks * ytiles, n);
        lst_ncc_tc<<<grid, TC_THREADS, smem, stream>>>(tasks + base, tpls, win8, best + base, tp);
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess)
            fprintf(stderr, "lst_ncc_tc launch failed: %s (grid %d x %d, smem %zu)\n", cudaGetErrorString(e), grid.x, grid.y, smem);
        launches++;
    }
    return launches;
}

}  // namespace lst
