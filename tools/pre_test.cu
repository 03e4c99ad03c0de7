// pre_test.cu -- stand-alone check of the window-resident blur kernel (lst_preprocess_win) against the first-generation kernel
// (lst_preprocess, itself byte-exact against the oracle in tests/test_gpu_parity.py) on random windows of a dense
// stack of slices, plus a timing loop of both on the bulk pass of config B.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -Xcompiler -ffp-contract=off \
//        -o tools/pre_test tools/pre_test.cu laser_spot_track_b200/csrc/lst_host.cpp laser_spot_track_b200/csrc/lst_ncc_tc.cu
#include <vector>

#include "../laser_spot_track_b200/csrc/lst_host.h"
#include "../laser_spot_track_b200/csrc/lst_kernels.cu"
using namespace lst;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

static uint32_t rng_state = 777u;
static uint32_t rnd()
{
    rng_state = rng_state * 1664525u + 1013904223u;
    return rng_state >> 8;
}

struct Case {
    int pix, W, H, nslices, n, w, h, mixed, range_per_task, quant, swap;
};

static int run_case(const Case &c, int time_iters)
{
    const int es = c.pix == LST_PIX_U8 ? 1 : (c.pix == LST_PIX_U16 ? 2 : 4);
    const size_t slice_bytes = (size_t)c.W * c.H * es;
    std::vector<uint8_t> stack(slice_bytes * c.nslices);
    for (int s = 0; s < c.nslices; s++)
        for (int y = 0; y < c.H; y++)
            for (int x = 0; x < c.W; x++) {
                const double g = 200.0 * exp(-((x - c.W * 0.4) * (x - c.W * 0.4) + (y - c.H * 0.55) * (y - c.H * 0.55)) / (2.0 * 30 * 30));
                const size_t i = ((size_t)s * c.H + y) * c.W + x;
                if (c.pix == LST_PIX_U8) stack[i] = (uint8_t)std::min(255.0, 20 + g + rnd() % 13);
                else if (c.pix == LST_PIX_U16) {
                    uint16_t v = (uint16_t)(3000 + 200 * g + rnd() % 900);
                    if (c.swap) v = (uint16_t)((v >> 8) | (v << 8));
                    ((uint16_t *)stack.data())[i] = v;
                } else
                    ((float *)stack.data())[i] = (float)(0.1 + g / 300.0 + (rnd() % 1000) * 1e-4);
            }
    uint8_t *d_stack;
    CK(cudaMalloc((void **)&d_stack, stack.size()));
    CK(cudaMemcpy(d_stack, stack.data(), stack.size(), cudaMemcpyHostToDevice));
    std::vector<Surface> surf(c.nslices);
    for (int s = 0; s < c.nslices; s++) surf[s] = Surface{d_stack + slice_bytes * s, c.W, 0, 0, c.H};
    Surface *d_surf;
    CK(cudaMalloc((void **)&d_surf, sizeof(Surface) * c.nslices));
    CK(cudaMemcpy(d_surf, surf.data(), sizeof(Surface) * c.nslices, cudaMemcpyHostToDevice));

    std::vector<DevTask> tasks(c.n), tasks2(c.n);
    int max_w = 0, max_h = 0;
    for (int i = 0; i < c.n; i++) {
        DevTask &t = tasks[i];
        memset(&t, 0, sizeof(t));
        t.frame = i % c.nslices;
        t.w = c.w - (c.mixed && i % 3 == 1 ? 5 : 0);
        t.h = c.h - (c.mixed && i % 3 == 2 ? 7 : 0);
        // origins anywhere, including flush against every border of the slice
        t.x0 = c.mixed && i % 5 == 0 ? 0 : (c.mixed && i % 5 == 1 ? c.W - t.w : (int)(rnd() % (uint32_t)(c.W - t.w + 1)));
        t.y0 = c.mixed && i % 7 == 0 ? 0 : (c.mixed && i % 7 == 1 ? c.H - t.h : (int)(rnd() % (uint32_t)(c.H - t.h + 1)));
        max_w = std::max(max_w, t.w);
        max_h = std::max(max_h, t.h);
    }
    const int pitch = (max_w + 15) & ~15;
    size_t off = 0;
    for (int i = 0; i < c.n; i++) {
        tasks[i].pitch = pitch;
        tasks[i].win_off = off;
        off += (size_t)pitch * tasks[i].h;
        tasks2[i] = tasks[i];
        tasks2[i].pad0 = 0;
        tasks2[i].pad1 = tasks[i].frame;
    }
    DevTask *d_tasks, *d_tasks2;
    uint8_t *d_w1, *d_w2;
    uint32_t *d_mm;
    CK(cudaMalloc((void **)&d_tasks, sizeof(DevTask) * c.n));
    CK(cudaMalloc((void **)&d_tasks2, sizeof(DevTask) * c.n));
    CK(cudaMalloc((void **)&d_w1, off + 64));
    CK(cudaMalloc((void **)&d_w2, off + 64));
    CK(cudaMalloc((void **)&d_mm, 8 * (size_t)c.n));
    CK(cudaMemcpy(d_tasks, tasks.data(), sizeof(DevTask) * c.n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tasks2, tasks2.data(), sizeof(DevTask) * c.n, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_w1, 0xEE, off + 64));
    CK(cudaMemset(d_w2, 0xDD, off + 64));

    PreprocParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.pixel_type = c.pix;
    pp.frame_w = c.W, pp.frame_h = c.H;
    pp.range_per_task = c.range_per_task;
    pp.lo = c.pix == LST_PIX_U8 ? 0.0f : 3000.0f;
    pp.hi = c.pix == LST_PIX_U8 ? 255.0f : 50000.0f;
    pp.quant_mode = c.quant;
    int tw = (max_w + 7) & ~7;
    pp.tile_w = tw < 192 ? ((tw + 15) & ~15) : 192;
    pp.minmax_in_kernel = (pp.range_per_task && max_w <= pp.tile_w && (long)max_w * max_h <= 256L * 256) ? 1 : 0;
    pp.swap_bytes = c.swap;

    // old kernel
    {
        std::vector<uint32_t> mm(2 * (size_t)c.n);
        for (int i = 0; i < c.n; i++) mm[2 * i] = 0xffffffffu, mm[2 * i + 1] = 0u;
        CK(cudaMemcpy(d_mm, mm.data(), 8 * (size_t)c.n, cudaMemcpyHostToDevice));
    }
    if (pp.range_per_task && !pp.minmax_in_kernel) launch_minmax(d_tasks, c.n, max_h, d_surf, pp, d_mm, 0);
    launch_preprocess(d_tasks, c.n, max_w, max_h, d_surf, pp, d_mm, d_w1, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("case pix %d slice %dx%d x%d, %d windows %dx%d mixed %d range %d quant %d swap %d: old kernel %s", c.pix, c.W, c.H, c.nslices, c.n,
           c.w, c.h, c.mixed, c.range_per_task, c.quant, c.swap, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;

    // new kernel
    if (!preprocess_win_eligible(c.pix, max_w, max_h)) {
        printf(", new kernel: not eligible\n");
        return 1;
    }
    int rc = launch_preprocess_win(d_tasks, c.n, max_w, max_h, d_surf, pp, d_w2, 0);
    e = cudaDeviceSynchronize();
    printf(", new kernel rc %d %s\n", rc, cudaGetErrorString(e));
    if (rc < 0 || e != cudaSuccess) return 1;
    std::vector<uint8_t> w1(off), w2(off);
    CK(cudaMemcpy(w1.data(), d_w1, off, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(w2.data(), d_w2, off, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int i = 0; i < c.n; i++) {
        const DevTask &t = tasks[i];
        for (int y = 0; y < t.h; y++)
            for (int x = 0; x < pitch; x++) {
                const uint8_t a = x < t.w ? w1[t.win_off + (size_t)y * pitch + x] : 0, b = w2[t.win_off + (size_t)y * pitch + x];
                if (a != b) {
                    if (bad < 8) printf("  window %d (%dx%d at %d,%d) pixel (%d,%d): old %u new %u\n", i, t.w, t.h, t.x0, t.y0, x, y, a, b);
                    bad++;
                }
            }
    }
    printf("  %ld differing bytes\n", bad);
    if (time_iters > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int which = 0; which < 2; which++) {
            for (int i = 0; i < time_iters + 3; i++) {
                if (i == 3) cudaEventRecord(e0);
                if (which == 0) launch_preprocess(d_tasks, c.n, max_w, max_h, d_surf, pp, d_mm, d_w1, 0);
                else launch_preprocess_win(d_tasks, c.n, max_w, max_h, d_surf, pp, d_w2, 0);
            }
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            double px = 0;
            for (int i = 0; i < c.n; i++) px += (double)tasks[i].w * tasks[i].h;
            printf("  %s: %.4f ms per launch, %.2f Gpx/s, %.2f TFLOP/s of blur arithmetic (38 flop/px)\n", which ? "lst_preprocess_win" : "lst_preprocess    ",
                   ms / time_iters, px / (ms / time_iters * 1e-3) * 1e-9, 38.0 * px / (ms / time_iters * 1e-3) * 1e-12);
        }
    }
    cudaFree(d_stack), cudaFree(d_surf), cudaFree(d_tasks), cudaFree(d_tasks2), cudaFree(d_w1), cudaFree(d_w2), cudaFree(d_mm);
    return bad != 0;
}

int main(int argc, char **argv)
{
    float kern[7], ksum[7];
    make_gaussian_kernel(kern, ksum);
    if (set_blur_kernel(kern, ksum) != 0) {
        fprintf(stderr, "set_blur_kernel failed (no device?)\n");
        return 2;
    }
    if (argc > 1 && !strcmp(argv[1], "time")) {
        run_case(Case{LST_PIX_U16, 1920, 1080, 16, 4704, 160, 160, 0, 1, 0, 0}, 20);
        return 0;
    }
    int fails = 0;
    // pix, W, H, slices, n, w, h, mixed, range_per_task, quant, swap
    fails += run_case(Case{LST_PIX_U16, 640, 480, 3, 40, 160, 160, 0, 1, 0, 0}, 0);
    fails += run_case(Case{LST_PIX_U16, 640, 480, 3, 60, 160, 160, 1, 1, 0, 0}, 0);
    fails += run_case(Case{LST_PIX_U16, 640, 480, 2, 30, 160, 160, 1, 0, 1, 1}, 0);
    fails += run_case(Case{LST_PIX_U8, 640, 480, 4, 60, 96, 96, 1, 0, 0, 0}, 0);
    fails += run_case(Case{LST_PIX_U8, 640, 480, 4, 60, 96, 96, 1, 1, 1, 0}, 0);
    fails += run_case(Case{LST_PIX_F32, 512, 400, 2, 30, 150, 139, 1, 1, 0, 0}, 0);
    fails += run_case(Case{LST_PIX_U16, 320, 304, 2, 20, 48, 48, 0, 1, 0, 0}, 0);       // templates
    fails += run_case(Case{LST_PIX_U16, 320, 304, 2, 30, 37, 21, 1, 1, 0, 0}, 0);       // small odd windows
    fails += run_case(Case{LST_PIX_U16, 1024, 512, 2, 20, 240, 256, 1, 1, 0, 0}, 0);    // largest windows the kernel takes
    fails += run_case(Case{LST_PIX_U8, 1024, 512, 2, 20, 233, 250, 1, 1, 0, 0}, 0);
    fails += run_case(Case{LST_PIX_U8, 641, 479, 3, 40, 96, 96, 1, 1, 0, 0}, 0);        // odd pitch: the alignment changes from row to row
    fails += run_case(Case{LST_PIX_U16, 333, 301, 3, 40, 120, 100, 1, 1, 0, 0}, 0);
    run_case(Case{LST_PIX_U16, 1920, 1080, 16, 4704, 160, 160, 0, 1, 0, 0}, 20);        // timing: the bulk pass of config B
    run_case(Case{LST_PIX_U8, 640, 480, 16, 5000, 96, 96, 0, 0, 0, 0}, 20);             // config A
    printf(fails ? "FAILED: %d cases\n" : "ALL OK\n", fails);
    return fails != 0;
}
