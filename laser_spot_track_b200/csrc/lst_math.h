// lst_math.h -- arithmetic shared bit-for-bit by the host code and the CUDA kernels.
//
// Everything here must round exactly like Java / OpenCV double and float arithmetic: one IEEE
// rounding per operation, no fused multiply-add.  The .cu files are compiled with -fmad=false
// and the host files with -ffp-contract=off, so plain C++ expressions are safe in both.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LST_HD __host__ __device__ __forceinline__
#else
#define LST_HD inline
#endif

namespace lst {

// Constants of cv::matchTemplate's normalisation for one template (OpenCV imgproc
// templmatch.cpp common_matchTemplate, reached from LST4:2560), 1 channel, TM_CCOEFF_NORMED.
struct TplConst {
    double inv_area;    // 1/(tw*th)
    double mean;        // templMean[0]
    double norm;        // sqrt(templSdv^2)/sqrt(invArea)
    int32_t flat;       // templSdv^2 < DBL_EPSILON  => whole result is 1
    int32_t tw, th;
    int32_t kw;         // 32-bit words per packed template row, multiple of 4 (see build_packed_template)
    uint32_t tsum, tsq_lo;
    double mideal;      // doMatch_test of the template (LST4:2724-2767)
    float inv_sqrt_c;   // 1/sqrt(N*tsq - tsum^2): float32 filter of the match kernel
    int32_t method;     // OpenCV match method 0..5 (LST4:2547-2552); 5 = TM_CCOEFF_NORMED
    double sum2;        // templSum2 = sum of t^2 as common_matchTemplate derives it
    double norm_raw;    // sqrt(E[t^2])/sqrt(invArea): templNorm of the non-CCOEFF methods
    // 3-channel templates (RGB stacks matched with matchIntensity = false, LST4:2525-2535 case 24): channels = 3, mean3 = the
    // per-channel templMean; norm / norm_raw / sum2 / flat then hold the sums over the channels as common_matchTemplate forms them
    int32_t channels;   // 1 or 3
    int32_t pad_;
    double mean3[3];
};

// One position of common_matchTemplate for TM_CCOEFF_NORMED with an exact integer numerator:
//   corr = sum I*T, wsum = sum I, wsq = sum I^2 over the template footprint.
// Operation order follows templmatch.cpp; the result is what OpenCV stores: (float)num.
LST_HD float ncc_score(double corr, double wsum, double wsq, double inv_area, double mean, double norm, int flat)
{
    if (flat) return 1.0f;
    double t = wsum;
    double wnd_mean2 = (t * t) * inv_area;
    double num = corr - t * mean;
    double diff2 = wsq - wnd_mean2;
    if (diff2 < 0.0) diff2 = 0.0;
    double thr = 10.0 * 1.1920928955078125e-07 * wsq;   // 10 * FLT_EPSILON * wndSum2
    if (thr > 0.5) thr = 0.5;
    double tt = (diff2 <= thr) ? 0.0 : sqrt(diff2) * norm;
    double a = fabs(num);
    double r;
    if (a < tt) r = num / tt;
    else if (a < tt * 1.125) r = num > 0.0 ? 1.0 : -1.0;
    else r = 0.0;
    return (float)r;
}

// One position of common_matchTemplate (templmatch.cpp) for any of the six methods of the plugin's dialog
// (LST4:2381): 0 SQDIFF, 1 SQDIFF_NORMED, 2 CCORR, 3 CCORR_NORMED, 4 CCOEFF, 5 CCOEFF_NORMED, from the
// exact integer correlation and window sums.  Double arithmetic in OpenCV's order, stored as float32.
LST_HD float match_score(const TplConst &c, double corr, double wsum, double wsq)
{
    const int method = c.method;
    if (method == 5) return ncc_score(corr, wsum, wsq, c.inv_area, c.mean, c.norm, c.flat);
    const int num_type = (method == 2 || method == 3) ? 0 : (method == 4 ? 1 : 2);
    const bool normed = method == 1 || method == 3;
    double num = corr, wnd_mean2 = 0.0, wnd_sum2 = 0.0;
    if (num_type == 1) {
        double t = wsum;
        wnd_mean2 = (t * t) * c.inv_area;
        num = num - t * c.mean;
    }
    if (normed || num_type == 2) {
        wnd_sum2 = wsq;
        if (num_type == 2) {
            num = wnd_sum2 - 2.0 * num + c.sum2;
            if (num < 0.0) num = 0.0;
        }
    }
    if (normed) {
        double diff2 = wnd_sum2 - wnd_mean2;
        if (diff2 < 0.0) diff2 = 0.0;
        double thr = 10.0 * 1.1920928955078125e-07 * wnd_sum2;
        if (thr > 0.5) thr = 0.5;
        double tt = (diff2 <= thr) ? 0.0 : sqrt(diff2) * c.norm_raw;
        double a = fabs(num);
        if (a < tt) num = num / tt;
        else if (a < tt * 1.125) num = num > 0.0 ? 1.0 : -1.0;
        else num = method != 1 ? 0.0 : 1.0;
    }
    return (float)num;
}

// The same for a 3-channel image (cn = 3 in common_matchTemplate): corr = sum over the channels of sum I_c*T_c, wsum[k] = the
// window sum of channel k, wsq = sum over the channels of the window sums of I_c^2.  Operation order as in templmatch.cpp:
// wndMean2 and num accumulate channel by channel, wndMean2 is scaled by invArea afterwards.
LST_HD float match_score3(const TplConst &c, double corr, const double wsum[3], double wsq)
{
    const int method = c.method;
    if (method == 5 && c.flat) return 1.0f;
    const int num_type = (method == 2 || method == 3) ? 0 : ((method == 4 || method == 5) ? 1 : 2);
    const bool normed = method == 1 || method == 3 || method == 5;
    double num = corr, wnd_mean2 = 0.0, wnd_sum2 = 0.0;
    if (num_type == 1) {
        for (int k = 0; k < 3; k++) {
            const double t = wsum[k];
            wnd_mean2 += t * t;
            num -= t * c.mean3[k];
        }
        wnd_mean2 *= c.inv_area;
    }
    if (normed || num_type == 2) {
        wnd_sum2 = wsq;
        if (num_type == 2) {
            num = wnd_sum2 - 2.0 * num + c.sum2;
            if (num < 0.0) num = 0.0;
        }
    }
    if (normed) {
        double diff2 = wnd_sum2 - wnd_mean2;
        if (diff2 < 0.0) diff2 = 0.0;
        double thr = 10.0 * 1.1920928955078125e-07 * wnd_sum2;
        if (thr > 0.5) thr = 0.5;
        const double tn = method == 5 ? c.norm : c.norm_raw;
        double tt = (diff2 <= thr) ? 0.0 : sqrt(diff2) * tn;
        double a = fabs(num);
        if (a < tt) num = num / tt;
        else if (a < tt * 1.125) num = num > 0.0 ? 1.0 : -1.0;
        else num = method != 1 ? 0.0 : 1.0;
    }
    return (float)num;
}

// methods 0 and 1 look for the minimum of the map, the others for the maximum (LST4:2671-2679)
LST_HD bool method_wants_min(int method) { return method == 0 || method == 1; }

// 3x3 quadratic sub-pixel fit, LST4:2681-2717.  nb = 3x3 neighbourhood (row-major, centre nb[4]).
// FloatIndexer.get returns a Java float: the four-term sum of fxy and the differences of fx / fy are float32
// operations (one rounding each) that are promoted only by the division by the double constant; fxx and fyy are
// promoted by the 2.0 * term from the start.  border != 0 => peak on the map border => (0,0).
LST_HD void subpixel_fit(const float *nb, int border, double *dx, double *dy)
{
    *dx = 0.0;
    *dy = 0.0;
    if (border) return;
    double c = (double)nb[4];
    double fxx = (double)nb[3] - 2.0 * c + (double)nb[5];
    double fyy = (double)nb[1] - 2.0 * c + (double)nb[7];
    float sxy = nb[8] + nb[0];
    sxy = sxy - nb[2];
    sxy = sxy - nb[6];
    double fxy = (double)sxy / 4.0;
    float dfx = nb[5] - nb[3], dfy = nb[7] - nb[1];
    double fx = (double)dfx / 2.0;
    double fy = (double)dfy / 2.0;
    double denom = fxy * fxy - fxx * fyy;
    if (denom == 0.0) return;
    double ddx = (fyy * fx - fxy * fy) / denom;
    double ddy = (fxx * fy - fxy * fx) / denom;
    if (fabs(ddx) > 1.0 || fabs(ddy) > 1.0) return;
    *dx = ddx;
    *dy = ddy;
}

// testMatchResult, LST4:1225-1255 (all six methods).
LST_HD int test_match_result(int method, double result, double ref, double thresh, double x, double y, int search_width,
                             int tpl_size)
{
    int ok = 1;
    double d1 = 0.05 * tpl_size, d2 = 0.05 * search_width;
    double dist = d1 < d2 ? d1 : d2;
    switch (method) {
        case 0: if (result / ref > thresh) ok = 0; break;
        case 1: if (result > thresh) ok = 0; break;
        case 2: if (fabs((result - ref) / ref) > thresh) ok = 0; break;
        case 3: if (fabs(result - ref) > thresh) ok = 0; break;
        case 4: if (fabs(result - ref) / ref > thresh) ok = 0; break;
        default: if (fabs(result - ref) > thresh) ok = 0; break;
    }
    if (search_width != 0 && ((x < dist) || (y < dist) || (x > search_width - dist) || (y > search_width - dist))) ok = 0;
    return ok;
}

// calcDisplacement, LST4:2318-2365.  ref = refX/refY (Java floats) of spot, mark1..4;
// dis = disX/disY of spot, mark1..4.  have_spot0 == 0 reproduces the firstPoint latch.
// out = dX_pix, dY_pix, x_abs, y_abs, dL.  spot0_io is updated when latching.
LST_HD void calc_displacement(const float *ref, const double *dis, double mark_dist, int have_spot0,
                              double *spot0_io, double *out)
{
    double dX_pix = dis[0] - dis[2];
    double dY_pix = dis[1] - dis[3];
    double m1x = (double)ref[2] + dis[2], m1y = (double)ref[3] + dis[3];
    double m2x = (double)ref[4] + dis[4], m2y = (double)ref[5] + dis[5];
    double m3x = (double)ref[6] + dis[6], m3y = (double)ref[7] + dis[7];
    double m4x = (double)ref[8] + dis[8], m4y = (double)ref[9] + dis[9];
    double a1x = m4x - m1x, a1y = m4y - m1y;
    double b1x = m2x - m1x, b1y = m2y - m1y;
    double a2x = m3x - m2x, a2y = m3y - m2y;
    double b2x = m3x - m4x, b2y = m3y - m4y;
    double rx = (double)ref[0] + dis[0] - m1x, ry = (double)ref[1] + dis[1] - m1y;
    double AX = a1x * (b1y - b2y) + a1y * (b2x - b1x);
    double BX = -a1x * b1y + a1y * b1x - rx * (b1y - b2y) - ry * (b2x - b1x);
    double CX = -rx * b1y + ry * b1x;
    double AY = b1x * (a1y - a2y) + b1y * (a2x - a1x);
    double BY = -b1x * a1y + b1y * a1x - rx * (a1y - a2y) - ry * (a2x - a1x);
    double CY = -rx * a1y + ry * a1x;
    double x_abs, y_abs;
    if (AX == 0.0) {
        x_abs = CX / BX * mark_dist;
        y_abs = CY / BY * mark_dist;
    } else {
        double DX = sqrt(BX * BX + 4.0 * AX * CX);
        double DY = sqrt(BY * BY + 4.0 * AY * CY);
        x_abs = (DX - BX) / AX / 2.0 * mark_dist;
        y_abs = (-DY - BY) / AY / 2.0 * mark_dist;
    }
    if (!have_spot0) {
        spot0_io[0] = x_abs;
        spot0_io[1] = y_abs;
    }
    double dX = x_abs - spot0_io[0];
    double dY = y_abs - spot0_io[1];
    out[0] = dX_pix;
    out[1] = dY_pix;
    out[2] = x_abs;
    out[3] = y_abs;
    out[4] = sqrt(dX * dX + dY * dY);
}

// Java (int) cast of a double (LST4:1327): truncation toward zero, saturating, NaN -> 0.
LST_HD int jint(double v)
{
    if (!(v == v)) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (int)(-2147483647 - 1);
    return (int)v;
}

}  // namespace lst
