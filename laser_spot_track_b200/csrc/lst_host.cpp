// lst_host.cpp -- host logic of the tracking path (see lst_host.h).  No CUDA in this file.
#include "lst_host.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace lst {

void fill_default_params(lst_params *p)
{
    memset(p, 0, sizeof(*p));
    p->struct_size = (int32_t)sizeof(lst_params);
    p->pixel_type = LST_PIX_U8;
    p->method = 5;            // matchMethodDefault  LST4:82
    p->templ_size = 120;      // templSizeDefault    LST4:84
    p->s_area = 50;           // searchAreaDefault   LST4:83
    p->sub_pixel = 1;         // LST4:89
    p->match_intensity = 1;   // LST4:88
    p->range_mode = LST_RANGE_DISPLAY;
    p->quant_mode = LST_QUANT_ROUND255;
    p->auto_skip = 1;
    p->max_s_area = 500;      // LST4:131
    p->mark_dist = 100.0;     // LST4:85
    p->range_lo = 0.0;
    p->range_hi = 255.0;
    p->match_threshold = 0.2; // matchThreshold[5]  LST4:147
}

int validate_params(const lst_params &p, std::string *err)
{
    auto fail = [&](int code, const char *m) {
        if (err) *err = m;
        return code;
    };
    if (p.struct_size != (int32_t)sizeof(lst_params)) return fail(LST_ERR_INVALID, "lst_params.struct_size mismatch");
    if (p.width <= 0 || p.height <= 0 || p.width > 32768 || p.height > 32768)
        return fail(LST_ERR_INVALID, "bad slice size");
    if (p.pixel_type != LST_PIX_U8 && p.pixel_type != LST_PIX_U16 && p.pixel_type != LST_PIX_F32 && p.pixel_type != LST_PIX_RGB)
        return fail(LST_ERR_UNSUPPORTED, "Unsupported image type");   // IJ.error text of LST4:2537
    if (p.method < 0 || p.method > 5) return fail(LST_ERR_INVALID, "matching method must be 0..5 (LST4:2381)");
    if (p.pixel_type == LST_PIX_U16 && !p.match_intensity)
        return fail(LST_ERR_UNSUPPORTED,
                    "16-bit stacks need matchIntensity=true (the reference dereferences a null Mat, LST4:2510-2523)");
    // 2*kRadius: below this ImageJ's convolveLine takes its short-line branches; 256: u32 window sums
    if (p.templ_size < 2 * kBlurRadius || p.templ_size > 256)
        return fail(LST_ERR_UNSUPPORTED, "templSize must be in [14, 256]");
    if (p.s_area < 0 || p.s_area > 8192) return fail(LST_ERR_INVALID, "bad sArea");
    if (p.templ_size > p.width || p.templ_size > p.height) return fail(LST_ERR_INVALID, "template larger than the slice");
    if (p.range_mode < 0 || p.range_mode > LST_RANGE_FRAME || p.quant_mode < 0 || p.quant_mode > 1)
        return fail(LST_ERR_INVALID, "bad range/quant mode");
    if (p.max_batch < 0) return fail(LST_ERR_INVALID, "bad max_batch");
    if (p.ingest_mode < 0 || p.ingest_mode > 3) return fail(LST_ERR_INVALID, "bad ingest_mode");
    if (p.ncc_engine < 0 || p.ncc_engine > 2) return fail(LST_ERR_INVALID, "bad ncc_engine");
    return LST_OK;
}

Rect clip_rect(int x0, int y0, int w, int h, int width, int height)
{
    long x1 = std::min<long>((long)x0 + w, width), y1 = std::min<long>((long)y0 + h, height);
    int cx = std::max(x0, 0), cy = std::max(y0, 0);
    Rect r;
    r.x = cx;
    r.y = cy;
    r.w = (int)std::max<long>(x1 - cx, 0);
    r.h = (int)std::max<long>(y1 - cy, 0);
    return r;
}

// LST4:506-515: float rect_half_size = templSize/2.0f;
// new Roi((int)(ref - half), (int)(ref - half), (int)(2*half), (int)(2*half)); rect = roi.getBounds()
Rect template_rect(const lst_params &p, float ref_x, float ref_y)
{
    volatile float half = (float)p.templ_size / 2.0f;
    volatile float fx = ref_x - half, fy = ref_y - half, fs = 2.0f * half;
    Rect r;
    r.x = jint((double)fx);
    r.y = jint((double)fy);
    r.w = jint((double)fs);
    r.h = r.w;
    return r;
}

// LST4:1325-1345 (clamp by shift) and LST4:1671-1695 (doubling: clamp by shrink + bound flags)
Rect window_for(const lst_params &p, const Rect &rect, int idis_x, int idis_y, int s, bool shrink, int *bounds)
{
    Rect w;
    int b = 0;
    if (p.s_area == 0) {      // LST4:1316-1320: whole slice
        w.x = 0;
        w.y = 0;
        w.w = p.width;
        w.h = p.height;
        if (bounds) *bounds = 0;
        return w;
    }
    w.x = rect.x + idis_x - s;
    w.y = rect.y + idis_y - s;
    w.w = rect.w + 2 * s;
    w.h = rect.h + 2 * s;
    if (w.x < 0) {
        w.x = 0;
        b |= 1;
    }
    if (w.y < 0) {
        w.y = 0;
        b |= 2;
    }
    if (w.x + w.w > p.width) {
        if (shrink) w.w = p.width - w.x;
        else w.x = p.width - w.w;
        b |= 4;
    }
    if (w.y + w.h > p.height) {
        if (shrink) w.h = p.height - w.y;
        else w.y = p.height - w.h;
        b |= 8;
    }
    if (bounds) *bounds = b;
    return w;
}

void make_gaussian_kernel(float kern[kBlurRadius], float ksum[kBlurRadius])
{
    const double sigma = 2.0, accuracy = 0.02;
    const int max_radius = 50;
    int k_radius = (int)ceil(sigma * sqrt(-2.0 * log(accuracy))) + 1;   // == 7
    if (k_radius > kBlurRadius) k_radius = kBlurRadius;
    for (int i = 0; i < k_radius; i++) kern[i] = (float)exp(-0.5 * i * i / sigma / sigma);
    if (k_radius < max_radius && k_radius > 3) {
        double sqrt_slope = 1.7976931348623157e308;
        int r = k_radius;
        while (r > k_radius / 2) {
            r--;
            double a = sqrt((double)kern[r]) / (k_radius - r);
            if (a < sqrt_slope) sqrt_slope = a;
            else break;
        }
        for (int r1 = r + 2; r1 < k_radius; r1++)
            kern[r1] = (float)((k_radius - r1) * (k_radius - r1) * sqrt_slope * sqrt_slope);
    }
    double sum = kern[0];
    for (int i = 1; i < k_radius; i++) sum += 2.0 * (double)kern[i];
    double rsum = 0.5 + 0.5 * (double)kern[0] / sum;
    for (int i = 0; i < k_radius; i++) {
        double v = (double)kern[i] / sum;
        kern[i] = (float)v;
        rsum -= v;
        ksum[i] = (float)rsum;
    }
}

TplConst template_consts(const uint8_t *tpl8, int tw, int th, int method)
{
    TplConst c;
    memset(&c, 0, sizeof(c));
    uint64_t s = 0, q = 0;
    for (int i = 0; i < tw * th; i++) {
        s += tpl8[i];
        q += (uint64_t)tpl8[i] * tpl8[i];
    }
    double n = (double)(tw * th);
    c.tw = tw;
    c.th = th;
    c.channels = 1;
    c.kw = packed_words(tw);
    c.tsum = (uint32_t)s;
    c.tsq_lo = (uint32_t)q;
    c.inv_area = 1.0 / n;
    c.mean = (double)s / n;
    double var = (double)q / n - c.mean * c.mean;
    if (var < 0.0) var = 0.0;
    double sdv = sqrt(var);
    double norm = sdv * sdv;
    c.flat = norm < 2.220446049250313e-16;
    norm = sqrt(norm);
    norm = norm / sqrt(c.inv_area);
    c.norm = norm;
    {
        double C = n * (double)q - (double)s * (double)s;   // exact: both terms < 2^53
        c.inv_sqrt_c = C > 0.0 ? (float)(1.0 / sqrt(C)) : 0.0f;
    }
    {   // templSum2 and the templNorm of the non-CCOEFF methods, in common_matchTemplate's order
        double sum2 = sdv * sdv + c.mean * c.mean;
        c.norm_raw = sqrt(sum2) / sqrt(c.inv_area);
        c.sum2 = sum2 / c.inv_area;
    }
    // *_mideal = doMatch_test(tpl, method == 0 ? 2 : method): the template matched against itself (LST4:566)
    c.method = method == 0 ? 2 : method;
    c.mideal = (double)match_score(c, (double)q, (double)s, (double)q);
    c.method = method;
    return c;
}

// planes: three tw x th planes one after the other (the channels of a ColorProcessor template after its per-channel blur)
TplConst template_consts3(const uint8_t *planes, int tw, int th, int method)
{
    TplConst c;
    memset(&c, 0, sizeof(c));
    const int n_px = tw * th;
    const double n = (double)n_px;
    c.tw = tw;
    c.th = th;
    c.kw = packed_words(tw);
    c.channels = 3;
    c.inv_area = 1.0 / n;
    uint64_t s[3], q[3], qsum = 0;
    double sdv2 = 0.0, mean2 = 0.0;
    for (int k = 0; k < 3; k++) {
        s[k] = q[k] = 0;
        for (int i = 0; i < n_px; i++) {
            const uint64_t v = planes[(size_t)k * n_px + i];
            s[k] += v;
            q[k] += v * v;
        }
        qsum += q[k];
        c.mean3[k] = (double)s[k] / n;                       // meanStdDev per channel
        double var = (double)q[k] / n - c.mean3[k] * c.mean3[k];
        if (var < 0.0) var = 0.0;
        const double sdv = sqrt(var);
        sdv2 += sdv * sdv;                                   // templSdv[0]^2 + templSdv[1]^2 + templSdv[2]^2 (+ 0)
        mean2 += c.mean3[k] * c.mean3[k];
    }
    c.tsum = (uint32_t)(s[0] + s[1] + s[2]);
    c.tsq_lo = (uint32_t)qsum;
    c.mean = 0.0;
    c.flat = sdv2 < 2.220446049250313e-16;
    c.norm = sqrt(sdv2) / sqrt(c.inv_area);
    {   // templSum2 = templNorm + sum of the squared means, then / invArea; the non-CCOEFF templNorm = sqrt(templSum2 before the division)
        double sum2 = sdv2;
        for (int k = 0; k < 3; k++) sum2 += c.mean3[k] * c.mean3[k];
        c.norm_raw = sqrt(sum2) / sqrt(c.inv_area);
        c.sum2 = sum2 / c.inv_area;
    }
    (void)mean2;
    const double ws[3] = {(double)s[0], (double)s[1], (double)s[2]};
    c.method = method == 0 ? 2 : method;                     // *_mideal = doMatch_test(tpl, method == 0 ? 2 : method), LST4:566
    c.mideal = (double)match_score3(c, (double)qsum, ws, (double)qsum);
    c.method = method;
    return c;
}

int packed_words(int tw) { return ((tw + 3) / 4 + 3) & ~3; }

void build_packed_template(const uint8_t *tpl8, int tw, int th, std::vector<uint32_t> *out)
{
    int kw = packed_words(tw);
    out->assign((size_t)th * kw, 0u);
    for (int v = 0; v < th; v++)
        for (int u = 0; u < tw; u++) (*out)[(size_t)v * kw + u / 4] |= (uint32_t)tpl8[v * tw + u] << (8 * (u & 3));
}

// ------------------------------------------------------------------------------------------
// recurrence resolver
// ------------------------------------------------------------------------------------------
namespace {

struct Entry {
    Rect win;
    int s_used;
    int prev;   // previous entry of the same (frame, template) slot, -1 = none
    lst_task_result r;
};

// Results are kept per (frame, template) slot as a chain in one flat vector (newest first): no
// per-slot allocations on the hot path -- a batch is a few thousand slots with 1-2 entries each.
struct Resolver {
    const ResolveInput &in;
    Executor &ex;
    int n;
    std::vector<Entry> entries;
    std::vector<int> head;        // [frame*5 + tpl] -> newest entry, -1 = none
    std::vector<lst_task> pending;
    std::vector<int> pend_prev;   // chain of pending tasks of the same slot
    std::vector<int> pend_head;   // [frame*5 + tpl] -> newest pending task, -1 = none
    std::vector<lst_task_result> res;
    lst_stats *st;

    Resolver(const ResolveInput &i, Executor &e, int nn, lst_stats *s)
        : in(i), ex(e), n(nn), head((size_t)nn * 5, -1), pend_head((size_t)nn * 5, -1), st(s)
    {
        entries.reserve((size_t)nn * 8);
        pending.reserve((size_t)nn * 5);
        pend_prev.reserve((size_t)nn * 5);
    }

    const Entry *find(int j, int k, const Rect &w, int s_used) const
    {
        for (int i = head[(size_t)j * 5 + k]; i >= 0; i = entries[i].prev)
            if (entries[i].win == w && entries[i].s_used == s_used) return &entries[i];
        return nullptr;
    }
    const Entry *latest(int j, int k) const
    {
        int i = head[(size_t)j * 5 + k];
        return i < 0 ? nullptr : &entries[i];
    }
    const Entry *need(int j, int k, const Rect &w, int s_used)
    {
        const Entry *e = find(j, k, w, s_used);
        if (e) return e;
        const size_t slot = (size_t)j * 5 + k;
        for (int pi = pend_head[slot]; pi >= 0; pi = pend_prev[pi]) {
            const lst_task &t = pending[pi];
            if (t.wx == w.x && t.wy == w.y && t.ww == w.w && t.wh == w.h && t.s_used == s_used) return nullptr;
        }
        pend_prev.push_back(pend_head[slot]);
        pend_head[slot] = (int)pending.size();
        lst_task t;
        t.frame = j;
        t.tpl = k;
        t.wx = w.x;
        t.wy = w.y;
        t.ww = w.w;
        t.wh = w.h;
        t.s_used = s_used;
        t.reserved = 0;
        pending.push_back(t);
        return nullptr;
    }

    int run_pending(bool first_pass)
    {
        if (pending.empty()) return LST_OK;
        res.resize(pending.size());
        int rc = ex.compute(pending.data(), (int)pending.size(), res.data());
        if (rc != LST_OK) return rc;
        for (size_t i = 0; i < pending.size(); i++) {
            const size_t slot = (size_t)pending[i].frame * 5 + pending[i].tpl;
            Entry e;
            e.win = Rect{pending[i].wx, pending[i].wy, pending[i].ww, pending[i].wh};
            e.s_used = pending[i].s_used;
            e.prev = head[slot];
            e.r = res[i];
            head[slot] = (int)entries.size();
            entries.push_back(e);
            pend_head[slot] = -1;
        }
        if (st) {
            st->tasks += (int64_t)pending.size();
            if (!first_pass) st->respeculated += (int64_t)pending.size();
            st->passes += 1;
        }
        pending.clear();
        pend_prev.clear();
        return LST_OK;
    }

    static void fill_tm(lst_template_match *tm, const Entry &e)
    {
        tm->x = e.r.x;
        tm->y = e.r.y;
        tm->score = e.r.score;
        tm->ix = e.r.ix;
        tm->iy = e.r.iy;
        tm->wx = e.win.x;
        tm->wy = e.win.y;
        tm->ww = e.win.w;
        tm->wh = e.win.h;
        tm->s_used = e.s_used;
        tm->accepted = e.r.accepted;
    }

    // One frame of analyseSlice given the displacement state before it.
    // exact: the state is validated; a missing result makes the frame incomplete (return 2).
    // !exact: prediction; missing results fall back to the latest stale result of that template.
    // return 0 = all five accepted (dis_out updated), 1 = frame rejected (dis_out = dis_in), 2 = incomplete.
    int eval_frame(int j, const double dis_in[10], bool exact, lst_record *rec, double dis_out[10])
    {
        const lst_params &p = in.p;
        int idx[5], idy[5];
        Rect wins[5];
        double dis[10];
        for (int k = 0; k < 5; k++) {
            idx[k] = jint(dis_in[2 * k]);
            idy[k] = jint(dis_in[2 * k + 1]);
            wins[k] = window_for(p, in.rect[k], idx[k], idy[k], p.s_area, false, nullptr);
        }
        memcpy(dis, dis_in, sizeof(dis));
        memcpy(dis_out, dis_in, sizeof(dis));
        if (rec) {   // prediction walks only need the state
            memset(rec, 0, sizeof(*rec));
            rec->status = LST_STATUS_OK;
            rec->reject = -1;
        }
        bool incomplete = false;
        for (int kk = 0; kk < 5; kk++) {
            int k = kMatchOrder[kk];
            const Entry *e = need(j, k, wins[k], p.s_area);
            if (!e) {
                incomplete = true;
                const Entry *stale = exact ? nullptr : latest(j, k);
                if (!stale) {
                    for (int k2 = kk + 1; k2 < 5; k2++) need(j, kMatchOrder[k2], wins[kMatchOrder[k2]], p.s_area);
                    return 2;
                }
                e = stale;
            }
            Entry cur = *e;
            if (!cur.r.accepted && p.doubling && p.s_area != 0 && (k == 1 || k == 0)) {
                // search-area doubling: mark1 LST4:1666-1735, spot LST4:2058-2128
                int s_new = p.s_area, flags = 0;
                bool found = false;
                while (!found && flags != 15) {
                    if (s_new > (1 << 24)) break;
                    s_new *= 2;
                    if (k == 0 && p.auto_skip && p.max_s_area != 0 && s_new > p.max_s_area) break;   // LST4:2066
                    int b = 0;
                    Rect dw = window_for(p, in.rect[k], idx[k], idy[k], s_new, true, &b);
                    flags |= b;
                    const Entry *de = need(j, k, dw, s_new);
                    if (!de) {
                        incomplete = true;
                        if (exact) return 2;
                        break;
                    }
                    cur = *de;
                    found = cur.r.accepted != 0;
                }
                if (found && k == 1) {
                    // mark1 recovered: shift the other four windows, LST4:1757-1880
                    double xs = cur.r.x + cur.win.x - in.rect[1].x - dis_in[2];
                    double ys = cur.r.y + cur.win.y - in.rect[1].y - dis_in[3];
                    for (int o = 0; o < 5; o++) {
                        if (o == 1) continue;
                        Rect &w = wins[o];
                        w.x = jint((double)w.x + xs);   // int += double
                        w.y = jint((double)w.y + ys);
                        if (w.x < 0) w.x = 0;
                        if (w.y < 0) w.y = 0;
                        if (w.x + w.w > p.width) w.x = p.width - w.w;
                        if (w.y + w.h > p.height) w.y = p.height - w.h;
                    }
                }
            }
            if (rec) fill_tm(&rec->tm[k], cur);
            if (!cur.r.accepted) {
                if (rec) {
                    rec->status = LST_STATUS_SKIP;
                    rec->reject = k;
                }
                return incomplete ? 2 : 1;
            }
            dis[2 * k] = cur.r.x + cur.win.x - in.rect[k].x;       // LST4:1891-1892, 2182-2183
            dis[2 * k + 1] = cur.r.y + cur.win.y - in.rect[k].y;
        }
        memcpy(dis_out, dis, sizeof(dis));
        return incomplete ? 2 : 0;
    }
};

}  // namespace

int resolve_batch(const ResolveInput &in, double dis_io[10], int64_t first_frame_index, int n,
                  Executor &ex, lst_record *out, lst_stats *st)
{
    if (n <= 0) return LST_OK;
    Resolver R(in, ex, n, st);
    double state[10];
    memcpy(state, dis_io, sizeof(state));
    // Speculation.  The windows of frame j are placed from the integer-truncated displacements left by
    // frame j-1, which are unknown until the frames before it are matched.  Long batches are therefore
    // speculated in two steps: (1a) every K-th frame is matched with the window origins of the state at
    // batch start -- the target is still inside those windows unless it moved by more than sArea, so
    // the positions found are the true ones; (1b) the states between the samples are interpolated
    // linearly and every frame is matched with the windows those predicted states place.  Short
    // batches skip (1a) and use the batch-start state for every frame.  Whatever the prediction, the
    // exact walk below accepts a result only if its window is the one the true state places.
    // The stride adapts when the caller keeps it between batches (ResolveInput::spec_stride): halved when
    // more than 10 % of the windows needed a fix-up pass, doubled (up to 64) when fewer than 2 % did.
    static const int kStrideEnv = [] {
        const char *e = getenv("LST_SPEC_STRIDE");
        int v = e ? atoi(e) : -1;
        return v;
    }();
    const bool adaptive = kStrideEnv < 0 && in.spec_stride != nullptr;
    int kStride = kStrideEnv >= 0 ? kStrideEnv : (in.spec_stride ? *in.spec_stride : 8);
    if (kStride < 2 && kStrideEnv < 0) kStride = 8;
    const int64_t respec_before = st ? st->respeculated : 0;
    int passes = 0;
    if (kStride >= 2 && n >= 4 * kStride) {
        std::vector<int> samp;
        for (int i = kStride - 1; i < n; i += kStride) samp.push_back(i);
        if (samp.back() != n - 1) samp.push_back(n - 1);
        double nd[10];
        for (int i : samp) R.eval_frame(i, state, false, nullptr, nd);
        int rc = R.run_pending(true);
        if (rc != LST_OK) return rc;
        passes++;
        // states after the sample frames (valid when all five matches were accepted)
        std::vector<double> sst(samp.size() * 10);
        std::vector<char> ok(samp.size(), 0);
        for (size_t m = 0; m < samp.size(); m++) {
            int r = R.eval_frame(samp[m], state, false, nullptr, &sst[m * 10]);
            ok[m] = r == 0;
        }
        R.pending.clear();      // only look-ups were meant; anything queued here is re-queued below if needed
        R.pend_prev.clear();
        std::fill(R.pend_head.begin(), R.pend_head.end(), -1);
        // predicted state before frame j = state after frame j-1: cubic Hermite interpolation through the
        // valid samples (anchor 0 = the batch-start state at frame -1), finite-difference tangents
        std::vector<int> at;            // anchor frame indices
        std::vector<const double *> ap; // anchor states
        at.push_back(-1);
        ap.push_back(state);
        for (size_t m = 0; m < samp.size(); m++)
            if (ok[m]) {
                at.push_back(samp[m]);
                ap.push_back(&sst[m * 10]);
            }
        const int na = (int)at.size();
        int seg = 0;                    // anchors seg, seg+1 bracket prev
        for (int j = 0; j < n; j++) {
            const int prev = j - 1;
            while (seg + 1 < na && at[seg + 1] <= prev) seg++;
            double pred[10];
            if (seg + 1 >= na || at[seg] == prev) {
                memcpy(pred, ap[seg], sizeof(pred));      // on an anchor, or past the last one: hold
            } else {
                const double t1 = at[seg], t2 = at[seg + 1], h = t2 - t1, u = (prev - t1) / h;
                const double u2 = u * u, u3 = u2 * u;
                const double h00 = 2 * u3 - 3 * u2 + 1, h10 = u3 - 2 * u2 + u, h01 = -2 * u3 + 3 * u2, h11 = u3 - u2;
                for (int c = 0; c < 10; c++) {
                    const double p1 = ap[seg][c], p2 = ap[seg + 1][c];
                    const double d = (p2 - p1) / h;
                    const double m1 = seg > 0 ? (p2 - ap[seg - 1][c]) / (t2 - at[seg - 1]) : d;
                    const double m2 = seg + 2 < na ? (ap[seg + 2][c] - p1) / (at[seg + 2] - t1) : d;
                    pred[c] = h00 * p1 + h10 * h * m1 + h01 * p2 + h11 * h * m2;
                }
            }
            R.eval_frame(j, pred, false, nullptr, nd);
        }
        rc = R.run_pending(true);
        if (rc != LST_OK) return rc;
        passes++;
    } else {
        double nd[10];
        for (int j = 0; j < n; j++) R.eval_frame(j, state, false, nullptr, nd);
        int rc = R.run_pending(true);
        if (rc != LST_OK) return rc;
        passes++;
    }
    int first = 0;
    while (first < n) {
        double walk[10];
        memcpy(walk, state, sizeof(walk));
        bool exact = true;
        for (int j = first; j < n; j++) {
            lst_record rec;
            double nd[10];
            if (exact) {
                int rc = R.eval_frame(j, walk, true, &rec, nd);
                if (rc != 2) {
                    rec.frame = first_frame_index + j;
                    memcpy(rec.dis, nd, sizeof(nd));
                    out[j] = rec;
                    memcpy(walk, nd, sizeof(nd));
                    memcpy(state, nd, sizeof(nd));
                    first = j + 1;
                    continue;
                }
                exact = false;
            }
            R.eval_frame(j, walk, false, nullptr, nd);   // prediction: queue what the chain will probably need
            memcpy(walk, nd, sizeof(nd));
        }
        if (first >= n) break;
        if (R.pending.empty()) return LST_ERR_STATE;   // cannot happen: the first unresolved frame always queues
        int rc = R.run_pending(false);
        if (rc != LST_OK) return rc;
        passes++;
    }
    // calcDisplacement for every accepted frame (device epilogue in the product path)
    std::vector<double> d10, o5;
    std::vector<int> idx;
    for (int j = 0; j < n; j++) {
        out[j].passes = passes;
        if (out[j].status == LST_STATUS_OK) {
            idx.push_back(j);
            d10.insert(d10.end(), out[j].dis, out[j].dis + 10);
        }
    }
    if (!idx.empty()) {
        o5.resize(idx.size() * 5);
        int rc = ex.solve(d10.data(), (int)idx.size(), o5.data());
        if (rc != LST_OK) return rc;
        for (size_t i = 0; i < idx.size(); i++) {
            lst_record &r = out[idx[i]];
            r.dX_pix = o5[i * 5 + 0];
            r.dY_pix = o5[i * 5 + 1];
            r.x_abs = o5[i * 5 + 2];
            r.y_abs = o5[i * 5 + 3];
            r.dL = o5[i * 5 + 4];
        }
    }
    memcpy(dis_io, state, sizeof(state));
    if (st) st->frames += n;
    if (adaptive && st && n >= 16) {
        const double frac = (double)(st->respeculated - respec_before) / (5.0 * n);
        if (frac > 0.10 && kStride > 2) *in.spec_stride = kStride / 2;
        else if (frac < 0.02 && kStride < 64 && n >= 4 * kStride) *in.spec_stride = kStride * 2;
    }
    return LST_OK;
}

}  // namespace lst
