// tc_rate.cu -- microbenchmark of tcgen05.mma kind::i8 M128 issued the way lst_ncc_tc issues it: operands in shared
// memory (SS), the A descriptor advancing by one row per template row, the B band of a K step covering 1-4 column
// blocks.  It answers two questions of the kernel's design:
//   (1) what an MMA of N = 32 / 64 / 96 / 128 costs when consecutive MMAs accumulate into the SAME TMEM range,
//   (2) what the kernel's per-template-row pattern (N = 32@0, 64@0, 96@0, 96@32, 64@64: overlapping but different
//       ranges back to back) costs, in the kernel's order and with the K step as the outer loop (runs of R rows).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_rate tools/tc_rate.cu
#include "../laser_spot_track_b200/csrc/lst_ncc_tc.cu"
using namespace lst;

// pattern 0: same range, N = n0 every time; pattern 1: the five K steps of a 48-wide template over 4 column blocks, row
// by row; pattern 2: the same MMAs with the K step outermost over runs of `run` template rows; pattern 3: as 1, but
// every MMA N = 128 over the full range (what zero padded bands would cost)
__global__ void __launch_bounds__(128) rate_kernel(int rows_total, int rows_alloc, int pattern, int n0, int run, long long *cycles)
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
    if (warp == 0 && elect_one()) {
        const uint32_t idesc0 = instr_desc_u8(128, 0);
        const uint64_t a0 = smem_desc(smem_u32(sm), rows_alloc * 16, 128);
        const uint64_t b0 = smem_desc(smem_u32(sm + 12 * rows_alloc * 16), 3 * 32 * 16, 128);     // band of nd = 3 blocks
        const uint64_t bfull = smem_desc(smem_u32(sm + 12 * rows_alloc * 16), 128 * 16, 128);
        const uint32_t a_k = 2 * rows_alloc;
        const int blo[5] = {0, 0, 0, 1, 2}, nblk[5] = {1, 2, 3, 3, 2}, boff[5] = {2, 1, 0, 0, 0};
        const long long t0 = clock64();
        if (pattern == 0) {
            for (int v = 0; v < rows_total; v++)
#pragma unroll
                for (int j = 0; j < 5; j++) mma_i8(tmem, a0 + (uint64_t)(v % 48) + (uint64_t)(j * a_k), b0, idesc0 | ((uint32_t)(n0 / 32) << 19), 1u);
        } else if (pattern == 1) {
            for (int v = 0; v < rows_total; v++)
#pragma unroll
                for (int j = 0; j < 5; j++)
                    mma_i8(tmem + 32 * blo[j], a0 + (uint64_t)(v % 48) + (uint64_t)(j * a_k), b0 + (uint64_t)(boff[j] * 32),
                           idesc0 | ((uint32_t)nblk[j] << 19), 1u);
        } else if (pattern == 2) {
            for (int v0 = 0; v0 < rows_total; v0 += run)
#pragma unroll
                for (int j = 0; j < 5; j++)
                    for (int v = v0; v < v0 + run; v++)
                        mma_i8(tmem + 32 * blo[j], a0 + (uint64_t)(v % 48) + (uint64_t)(j * a_k), b0 + (uint64_t)(boff[j] * 32),
                               idesc0 | ((uint32_t)nblk[j] << 19), 1u);
        } else {
            for (int v = 0; v < rows_total; v++)
#pragma unroll
                for (int j = 0; j < 5; j++) mma_i8(tmem, a0 + (uint64_t)(v % 48) + (uint64_t)(j * a_k), bfull, idesc0 | (4u << 19), 1u);
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        if (blockIdx.x == 0) *cycles = clock64() - t0;
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
    }
}

static void run(const char *what, int pattern, int n0, int runlen)
{
    const int rows_alloc = 176, rows_total = 4800;
    const size_t smem = 150 * 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long *d_c, c = 0;
    cudaMalloc((void **)&d_c, 8);
    for (int rep = 0; rep < 2; rep++) rate_kernel<<<148, 128, smem>>>(rows_total, rows_alloc, pattern, n0, runlen, d_c);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    printf("%-58s %7.1f cycles per MMA, %7.1f per template row of 5 MMAs (%s)\n", what, (double)c / (rows_total * 5.0),
           (double)c / rows_total, cudaGetErrorString(e));
    cudaFree(d_c);
}

int main()
{
    run("same range, N = 32", 0, 32, 0);
    run("same range, N = 64", 0, 64, 0);
    run("same range, N = 96", 0, 96, 0);
    run("same range, N = 128 (B band of 96 rows re-read)", 0, 128, 0);
    run("kernel pattern 32@0 64@0 96@0 96@32 64@64, row by row", 1, 0, 0);
    run("kernel pattern, K step outermost over runs of 2 rows", 2, 0, 2);
    run("kernel pattern, K step outermost over runs of 4 rows", 2, 0, 4);
    run("kernel pattern, K step outermost over runs of 8 rows", 2, 0, 8);
    run("kernel pattern, K step outermost over runs of 48 rows", 2, 0, 48);
    run("every MMA N = 128 over the full range", 3, 0, 0);
    return 0;
}
