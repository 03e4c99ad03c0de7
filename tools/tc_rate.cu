// tc_rate.cu -- microbenchmark: issue rate of tcgen05.mma kind::i8 M128 for small N, operands in shared
// memory (SS mode), with the A descriptor advancing like lst_ncc_tc's (row shift per template row).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_rate tools/tc_rate.cu
#include "../laser_spot_track_b200/csrc/lst_ncc_tc.cu"
using namespace lst;

template <int N>
__global__ void __launch_bounds__(128) rate_kernel(int iters, int rows, int a_mode)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_t;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_t)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_t;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc_u8(128, N);
        const uint64_t a0 = smem_desc(smem_u32(sm), rows * 16, 128);
        const uint64_t b0 = smem_desc(smem_u32(sm + 10 * rows * 16), N * 16, 128);
        const int nblk = 128 / N;
        for (int i = 0; i < iters; i++) {
            // one "template row": nblk column blocks x 3 k-steps, A rows shifted by (i % 48)
            uint64_t a_v = a0 + (a_mode ? (uint64_t)(i % 48) : 0);
            for (int nb = 0; nb < nblk; nb++) {
                uint64_t a = a_v + (uint64_t)(a_mode ? nb * (N / 16) * rows : 0), b = b0;
                for (int ks = 0; ks < 3; ks++) {
                    mma_i8(tmem + nb * N, a, b, idesc, 1u);
                    a += a_mode ? 2 * rows : 0;
                    b += 2 * N;
                }
            }
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
    }
}

template <int N>
static void run(int ctas_per_sm, int a_mode)
{
    const int rows = 176, iters = 4000;
    size_t smem = (size_t)10 * rows * 16 + 6 * N * 16 + 4096;
    if (ctas_per_sm == 1) smem = 120 * 1024;   // force one CTA per SM
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        rate_kernel<N><<<148 * ctas_per_sm, 128, smem>>>(iters, rows, a_mode);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double mmas = (double)iters * (128 / N) * 3 * ctas_per_sm;          // per SM
    const double cyc = best * 1e-3 * 1.965e9;
    printf("N=%3d ctas/sm=%d a_mode=%d: %.3f ms, %.1f cycles per MMA per SM, %.1f cycles per 128 columns x 3 k-steps (%s)\n", N,
           ctas_per_sm, a_mode, best, cyc / mmas, cyc / mmas * (128 / N) * 3, cudaGetErrorString(e));
}

int main()
{
    for (int a_mode = 0; a_mode < 2; a_mode++)
        for (int c : {1, 4}) {
            run<32>(c, a_mode);
            run<64>(c, a_mode);
            run<128>(c, a_mode);
        }
    return 0;
}
