// tc_test.cu -- stand-alone check of the persistent tensor-core match kernel (lst_ncc_tc) against a brute-force
// CPU evaluation of the same arithmetic (integer correlation and window sums, match_score in double, first
// maximum in row-major order), for a list of window / template geometries, plus a timing loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -Xcompiler -ffp-contract=off \
//        -o tools/tc_test tools/tc_test.cu laser_spot_track_b200/csrc/lst_host.cpp
#include <vector>

#include "../laser_spot_track_b200/csrc/lst_host.h"
#include "../laser_spot_track_b200/csrc/lst_ncc_tc.cu"
using namespace lst;

static uint32_t rng_state = 12345u;
static uint32_t rnd()
{
    rng_state = rng_state * 1664525u + 1013904223u;
    return rng_state >> 8;
}

static uint32_t h_f32_key(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static unsigned long long h_pack_key(float score, uint32_t idx, bool want_min)
{
    score = score + 0.0f;
    uint32_t k = h_f32_key(score);
    if (want_min) k = ~k;
    return ((unsigned long long)k << 32) | (unsigned long long)(0xffffffffu - idx);
}

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e_ = (x);                                                       \
        if (e_ != cudaSuccess) {                                                    \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(2);                                                                \
        }                                                                           \
    } while (0)

struct Geo {
    int w, h, tw, th, n, method, mixed;
};

static int run_case(const Geo &g, int verify, int time_iters)
{
    const int ntpl = 5;
    std::vector<std::vector<uint8_t>> tpl(ntpl);
    DevTemplates ht;
    memset(&ht, 0, sizeof(ht));
    std::vector<uint8_t *> d_tab(ntpl);
    for (int k = 0; k < ntpl; k++) {
        tpl[k].resize((size_t)g.tw * g.th);
        const double cx = g.tw * (0.3 + 0.1 * k), cy = g.th * (0.6 - 0.07 * k), sg = g.tw / 6.0 + k;
        for (int v = 0; v < g.th; v++)
            for (int u = 0; u < g.tw; u++) {
                double r2 = (u - cx) * (u - cx) + (v - cy) * (v - cy);
                int val = (int)(30 + 200 * exp(-r2 / (2 * sg * sg))) + (int)(rnd() % 7);
                tpl[k][(size_t)v * g.tw + u] = (uint8_t)(val > 255 ? 255 : val);
            }
        if (k == 4 && g.mixed) memset(tpl[k].data(), 77, tpl[k].size());     // a flat template: the map is all ones
        ht.c[k] = template_consts(tpl[k].data(), g.tw, g.th, g.method);
        std::vector<uint8_t> tab(tc_table_bytes(g.tw, g.th));
        tc_build_table(tpl[k].data(), g.tw, g.th, tab.data());
        CK(cudaMalloc((void **)&d_tab[k], tab.size()));
        CK(cudaMemcpy(d_tab[k], tab.data(), tab.size(), cudaMemcpyHostToDevice));
        ht.toeplitz[k] = d_tab[k];
    }
    DevTemplates *d_tpls;
    CK(cudaMalloc((void **)&d_tpls, sizeof(ht)));
    CK(cudaMemcpy(d_tpls, &ht, sizeof(ht), cudaMemcpyHostToDevice));

    // windows: mixed sizes when asked (every 3rd window smaller), one pitch for the pass
    std::vector<DevTask> tasks(g.n);
    int pitch = (g.w + 15) & ~15;
    size_t rows = 0;
    int max_rw = 0, max_rh = 0;
    for (int i = 0; i < g.n; i++) {
        DevTask &t = tasks[i];
        memset(&t, 0, sizeof(t));
        t.tpl = i % ntpl;
        t.w = g.w - ((g.mixed && i % 3 == 1) ? 5 : 0);
        t.h = g.h - ((g.mixed && i % 3 == 2) ? 3 : 0);
        t.pitch = pitch;
        t.rw = t.w - g.tw + 1;
        t.rh = t.h - g.th + 1;
        t.win_off = rows * pitch;
        rows += t.h;
        max_rw = std::max(max_rw, t.rw);
        max_rh = std::max(max_rh, t.rh);
    }
    std::vector<uint8_t> win(rows * pitch + 64, 0);
    for (int i = 0; i < g.n; i++) {
        const DevTask &t = tasks[i];
        uint8_t *wp = win.data() + t.win_off;
        const int kind = g.mixed ? (i % 7) : 0;
        // scene: the template pasted somewhere over a noisy ramp
        const int px = (int)(rnd() % (uint32_t)t.rw), py = (int)(rnd() % (uint32_t)t.rh);
        for (int y = 0; y < t.h; y++)
            for (int x = 0; x < t.pitch; x++) {
                int val = 40 + (x + 2 * y) / 8 + (int)(rnd() % 9);
                if (kind == 5) val = 90;                    // flat window
                if (kind == 6) val = (x / 3 + y / 5) & 1 ? 200 : 20;   // periodic: many equal maxima
                wp[(size_t)y * t.pitch + x] = x < t.w ? (uint8_t)val : (uint8_t)(rnd() & 255);   // junk in the padding must not matter
            }
        if (kind != 5 && kind != 6)
            for (int v = 0; v < g.th; v++)
                for (int u = 0; u < g.tw; u++) {
                    int val = tpl[t.tpl][(size_t)v * g.tw + u] + (int)(rnd() % 5) - 2;
                    wp[(size_t)(py + v) * t.pitch + px + u] = (uint8_t)(val < 0 ? 0 : (val > 255 ? 255 : val));
                }
    }
    DevTask *d_tasks;
    uint8_t *d_win;
    unsigned long long *d_best;
    float *d_nbhd;
    CK(cudaMalloc((void **)&d_tasks, sizeof(DevTask) * g.n));
    CK(cudaMalloc((void **)&d_win, win.size()));
    CK(cudaMalloc((void **)&d_best, 8 * (size_t)g.n));
    CK(cudaMalloc((void **)&d_nbhd, 48 * (size_t)g.n));
    CK(cudaMemcpy(d_tasks, tasks.data(), sizeof(DevTask) * g.n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_win, win.data(), win.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_best, 0, 8 * (size_t)g.n));
    CK(cudaMemset(d_nbhd, 0, 48 * (size_t)g.n));
    int rc = launch_ncc_tc(d_tasks, g.n, max_rw, max_rh, g.tw, g.th, pitch, (long long)rows, d_tpls, d_win, d_best, d_nbhd, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("case w%d h%d t%dx%d n%d method %d mixed %d: launch rc %d, sync %s\n", g.w, g.h, g.tw, g.th, g.n, g.method, g.mixed, rc,
           cudaGetErrorString(e));
    if (rc < 0 || e != cudaSuccess) return 1;
    std::vector<unsigned long long> best(g.n);
    std::vector<float> nbhd(12 * (size_t)g.n);
    CK(cudaMemcpy(best.data(), d_best, 8 * (size_t)g.n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(nbhd.data(), d_nbhd, 48 * (size_t)g.n, cudaMemcpyDeviceToHost));
    int bad = 0;
    const bool want_min = method_wants_min(g.method);
    const int nver = verify < g.n ? verify : g.n;
    for (int i = 0; i < nver; i++) {
        const DevTask &t = tasks[i];
        const TplConst &c = ht.c[t.tpl];
        const uint8_t *wp = win.data() + t.win_off;
        std::vector<float> sc((size_t)t.rw * t.rh);
        unsigned long long want = 0;
        for (int y = 0; y < t.rh; y++)
            for (int x = 0; x < t.rw; x++) {
                uint32_t corr = 0, s = 0, q = 0;
                for (int v = 0; v < g.th; v++) {
                    const uint8_t *wr = wp + (size_t)(y + v) * t.pitch + x, *tr = tpl[t.tpl].data() + (size_t)v * g.tw;
                    for (int u = 0; u < g.tw; u++) {
                        uint32_t p = wr[u];
                        corr += p * tr[u];
                        s += p;
                        q += p * p;
                    }
                }
                float f = match_score(c, (double)corr, (double)s, (double)q);
                sc[(size_t)y * t.rw + x] = f;
                unsigned long long key = h_pack_key(f, (uint32_t)(y * t.rw + x), want_min);
                if (key > want) want = key;
            }
        if (want != best[i]) {
            uint32_t wi = 0xffffffffu - (uint32_t)want, gi = 0xffffffffu - (uint32_t)best[i];
            if (bad < 10)
                printf("  task %d (tpl %d, %dx%d): best key %016llx (x %u y %u) want %016llx (x %u y %u)\n", i, t.tpl, t.w, t.h, best[i],
                       gi % t.rw, gi / t.rw, want, wi % t.rw, wi / t.rw);
            bad++;
            continue;
        }
        if (nbhd[12 * (size_t)i + 9] != 0.0f) {
            uint32_t wi = 0xffffffffu - (uint32_t)want;
            int bx = wi % t.rw, by = wi / t.rw;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int x = bx + dx, y = by + dy;
                    if (x < 0 || y < 0 || x >= t.rw || y >= t.rh) continue;
                    float gv = nbhd[12 * (size_t)i + (dy + 1) * 3 + dx + 1], wv = sc[(size_t)y * t.rw + x];
                    if (memcmp(&gv, &wv, 4) != 0) {
                        if (bad < 10) printf("  task %d: nbhd (%d,%d) got %.9g want %.9g\n", i, dx, dy, gv, wv);
                        bad++;
                    }
                }
        }
    }
    printf("  verified %d tasks: %d mismatches%s\n", nver, bad, nbhd[9] != 0.0f ? " (3x3 scores checked)" : "");
    if (time_iters > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = 0; i < 3; i++) launch_ncc_tc(d_tasks, g.n, max_rw, max_rh, g.tw, g.th, pitch, (long long)rows, d_tpls, d_win, d_best, d_nbhd, 0);
        cudaEventRecord(e0);
        for (int i = 0; i < time_iters; i++)
            launch_ncc_tc(d_tasks, g.n, max_rw, max_rh, g.tw, g.th, pitch, (long long)rows, d_tpls, d_win, d_best, d_nbhd, 0);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double macs = 0;
        for (int i = 0; i < g.n; i++) macs += (double)tasks[i].rw * tasks[i].rh * g.tw * g.th;
        printf("  time: %.4f ms per launch of %d windows, %.1f TOP/s algorithmic, %.2f Mwindows/s\n", ms / time_iters, g.n,
               2.0 * macs / (ms / time_iters * 1e-3) * 1e-12, g.n / (ms / time_iters * 1e-3) * 1e-6);
    }
    cudaFree(d_tasks);
    cudaFree(d_win);
    cudaFree(d_best);
    cudaFree(d_nbhd);
    cudaFree(d_tpls);
    for (int k = 0; k < ntpl; k++) cudaFree(d_tab[k]);
    return bad != 0;
}

int main(int argc, char **argv)
{
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");
    if (argc > 1 && !strcmp(argv[1], "time")) {
        tc_init_device();
        run_case(Geo{160, 160, 48, 48, 4704, 5, 0}, 0, 20);
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "timeA")) {
        tc_init_device();
        run_case(Geo{96, 96, 32, 32, 5000, 5, 0}, 0, 20);
        return 0;
    }
    tc_init_device();
    int fails = 0;
    // w, h, tw, th, n, method, mixed
    fails += run_case(Geo{96, 96, 32, 32, 40, 5, 0}, 40, 0);          // config A
    fails += run_case(Geo{160, 160, 48, 48, 300, 5, 0}, 24, 0);       // config B
    fails += run_case(Geo{160, 160, 48, 48, 70, 5, 1}, 70, 0);        // mixed sizes, flat template / windows, ties
    if (!quick) {
        fails += run_case(Geo{75, 61, 21, 17, 33, 5, 1}, 33, 0);      // odd sizes
        fails += run_case(Geo{272, 272, 48, 48, 12, 5, 0}, 6, 0);     // doubled search area: 225 x 225 map, four tiles
        fails += run_case(Geo{512, 512, 128, 128, 6, 5, 0}, 2, 0);    // config D
        fails += run_case(Geo{600, 420, 64, 64, 4, 5, 0}, 2, 0);      // config C like: many tiles per window
        fails += run_case(Geo{200, 190, 181, 150, 6, 5, 0}, 6, 0);    // widest template
        for (int m = 0; m < 5; m++) fails += run_case(Geo{96, 96, 32, 32, 10, m, 0}, 10, 0);
    }
    run_case(Geo{160, 160, 48, 48, 4704, 5, 0}, 0, 20);               // timing: the bulk pass of config B
    run_case(Geo{96, 96, 32, 32, 5000, 5, 0}, 0, 20);
    if (!quick) run_case(Geo{512, 512, 128, 128, 100, 5, 0}, 0, 5);
    printf(fails ? "FAILED: %d cases\n" : "ALL OK\n", fails);
    return fails != 0;
}
