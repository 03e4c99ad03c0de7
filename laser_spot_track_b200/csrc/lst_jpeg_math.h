// lst_jpeg_math.h -- the sample reconstruction of a JPEG image, shared bit for bit by the host decoder (lst_jpeg.cpp) and
// the device stage (lst_jpeg_reconstruct in lst_kernels.cu): the IJG library's default integer algorithms -- jidctint.c
// (accurate integer inverse DCT, dequantisation folded in), jdsample.c ("fancy" triangle-filter upsampling) and jdcolor.c
// (YCbCr -> RGB in 16-bit fixed point).  Integer arithmetic only: one defined result, the same on both sides.
#pragma once
#include <stdint.h>

#include "lst_math.h"   // LST_HD

namespace lst {

// what the reconstruction of a region needs to know about the image; the same for every slice of a batch
struct JpegGeom {
    int32_t width, height;
    int32_t ncomp;          // 1 (grey) or 3
    int32_t hmax, vmax;     // luma sampling factors (1 or 2); the chroma components are 1 x 1
    int32_t ycc;            // 3 components: 1 = YCbCr -> RGB on output, 0 = the components are R, G, B
};
inline bool operator==(const JpegGeom &a, const JpegGeom &b)
{
    return a.width == b.width && a.height == b.height && a.ncomp == b.ncomp && a.hmax == b.hmax && a.vmax == b.vmax && a.ycc == b.ycc;
}
LST_HD int jpeg_blocks_per_mcu(const JpegGeom &g) { return g.ncomp == 1 ? 1 : g.hmax * g.vmax + 2; }
LST_HD int jpeg_ds_w(const JpegGeom &g) { return (g.width + g.hmax - 1) / g.hmax; }     // downsampled_width of a chroma component
LST_HD int jpeg_ds_h(const JpegGeom &g) { return (g.height + g.vmax - 1) / g.vmax; }

// the MCUs that cover a pixel rectangle plus the two pixels of chroma context around it
struct JpegRegion {
    int32_t mx0, my0, nmx, nmy;
};
LST_HD JpegRegion jpeg_region(const JpegGeom &g, int x, int y, int w, int h)
{
    const int mw = 8 * g.hmax, mh = 8 * g.vmax;
    const int mcus_x = (g.width + mw - 1) / mw, mcus_y = (g.height + mh - 1) / mh;
    int mx0 = (x - 2) / mw, mx1 = (x + w + 1) / mw, my0 = (y - 2) / mh, my1 = (y + h + 1) / mh;
    if (x - 2 < 0) mx0 = 0;
    if (y - 2 < 0) my0 = 0;
    if (mx1 > mcus_x - 1) mx1 = mcus_x - 1;
    if (my1 > mcus_y - 1) my1 = mcus_y - 1;
    JpegRegion r;
    r.mx0 = mx0, r.my0 = my0, r.nmx = mx1 - mx0 + 1, r.nmy = my1 - my0 + 1;
    return r;
}

// ---- jidctint.c ------------------------------------------------------------------------------------
LST_HD long long jpeg_descale(long long x, int n) { return (x + ((long long)1 << (n - 1))) >> n; }
// range_limit[x & RANGE_MASK] of the IDCT: x modulo 1024 as a signed value, recentred by 128, clamped to 0..255
LST_HD uint8_t jpeg_idct_limit(long long x)
{
    int v = (int)(x & 1023);
    if (v >= 512) v -= 1024;
    v += 128;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
// one 8-point pass: even part in t10..t13, odd part in o0..o3 (the caller combines and descales them)
LST_HD void jpeg_idct_1d(const long long in[8], long long &t10, long long &t11, long long &t12, long long &t13, long long &o0,
                         long long &o1, long long &o2, long long &o3)
{
    long long z2 = in[2], z3 = in[6];
    long long z1 = (z2 + z3) * 4433;                     // FIX_0_541196100
    const long long tmp2 = z1 + z3 * (-15137);           // FIX_1_847759065
    const long long tmp3 = z1 + z2 * 6270;               // FIX_0_765366865
    z2 = in[0], z3 = in[4];
    const long long tmp0 = (z2 + z3) * 8192;             // << CONST_BITS (13)
    const long long tmp1 = (z2 - z3) * 8192;
    t10 = tmp0 + tmp3, t13 = tmp0 - tmp3, t11 = tmp1 + tmp2, t12 = tmp1 - tmp2;
    long long a0 = in[7], a1 = in[5], a2 = in[3], a3 = in[1];
    z1 = a0 + a3;
    z2 = a1 + a2;
    z3 = a0 + a2;
    long long z4 = a1 + a3;
    const long long z5 = (z3 + z4) * 9633;               // FIX_1_175875602
    a0 *= 2446;                                          // FIX_0_298631336
    a1 *= 16819;                                         // FIX_2_053119869
    a2 *= 25172;                                         // FIX_3_072711026
    a3 *= 12299;                                         // FIX_1_501321110
    z1 *= -7373;                                         // FIX_0_899976223
    z2 *= -20995;                                        // FIX_2_562915447
    z3 *= -16069;                                        // FIX_1_961570560
    z4 *= -3196;                                         // FIX_0_390180644
    z3 += z5;
    z4 += z5;
    o0 = a0 + z1 + z3;
    o1 = a1 + z2 + z4;
    o2 = a2 + z2 + z3;
    o3 = a3 + z1 + z4;
}
// coef: 64 quantised coefficients in natural order; q: the quantisation table in natural order; out: 8 rows of 8 samples
LST_HD void jpeg_idct_islow(const int16_t *coef, const uint16_t *q, uint8_t *out, int stride)
{
    long long ws[64];
    for (int c = 0; c < 8; c++) {
        long long in[8];
        for (int r = 0; r < 8; r++) in[r] = (long long)coef[8 * r + c] * (long long)q[8 * r + c];
        long long t10, t11, t12, t13, o0, o1, o2, o3;
        jpeg_idct_1d(in, t10, t11, t12, t13, o0, o1, o2, o3);
        const int sh = 13 - 2;                           // CONST_BITS - PASS1_BITS
        ws[8 * 0 + c] = jpeg_descale(t10 + o3, sh);
        ws[8 * 7 + c] = jpeg_descale(t10 - o3, sh);
        ws[8 * 1 + c] = jpeg_descale(t11 + o2, sh);
        ws[8 * 6 + c] = jpeg_descale(t11 - o2, sh);
        ws[8 * 2 + c] = jpeg_descale(t12 + o1, sh);
        ws[8 * 5 + c] = jpeg_descale(t12 - o1, sh);
        ws[8 * 3 + c] = jpeg_descale(t13 + o0, sh);
        ws[8 * 4 + c] = jpeg_descale(t13 - o0, sh);
    }
    for (int r = 0; r < 8; r++) {
        long long t10, t11, t12, t13, o0, o1, o2, o3;
        jpeg_idct_1d(ws + 8 * r, t10, t11, t12, t13, o0, o1, o2, o3);
        const int sh = 13 + 2 + 3;                       // CONST_BITS + PASS1_BITS + 3
        uint8_t *p = out + (size_t)r * stride;
        p[0] = jpeg_idct_limit(jpeg_descale(t10 + o3, sh));
        p[7] = jpeg_idct_limit(jpeg_descale(t10 - o3, sh));
        p[1] = jpeg_idct_limit(jpeg_descale(t11 + o2, sh));
        p[6] = jpeg_idct_limit(jpeg_descale(t11 - o2, sh));
        p[2] = jpeg_idct_limit(jpeg_descale(t12 + o1, sh));
        p[5] = jpeg_idct_limit(jpeg_descale(t12 - o1, sh));
        p[3] = jpeg_idct_limit(jpeg_descale(t13 + o0, sh));
        p[4] = jpeg_idct_limit(jpeg_descale(t13 - o0, sh));
    }
}

// ---- jdsample.c: "fancy" (triangle filter) upsampling, one output sample ----------------------------
// cx = chroma column of the output column X (X >> 1); `row` points at the chroma row with index 0 = chroma column `c_org`,
// ds_w = downsampled_width of the component.  h2v1: out[2i] = (3 in[i] + in[i-1] + 1) >> 2, out[2i+1] = (3 in[i] + in[i+1]
// + 2) >> 2; the first and the last output sample of a row are copies of the first / last input sample.
LST_HD int jpeg_up_h2v1(const uint8_t *row, int c_org, int ds_w, int X)
{
    const int cx = X >> 1, i = cx - c_org;
    if (X & 1) return cx == ds_w - 1 ? row[i] : (3 * row[i] + row[i + 1] + 2) >> 2;
    return cx == 0 ? row[i] : (3 * row[i] + row[i - 1] + 1) >> 2;
}
// h2v2: column sums 3 * nearer row + further row, then the same horizontal filter on the sums with rounding 8 / 7
LST_HD int jpeg_up_h2v2(const uint8_t *near_row, const uint8_t *far_row, int c_org, int ds_w, int X)
{
    const int cx = X >> 1, i = cx - c_org;
    const int cur = 3 * near_row[i] + far_row[i];
    if (X & 1) return cx == ds_w - 1 ? (cur * 4 + 7) >> 4 : (cur * 3 + 3 * near_row[i + 1] + far_row[i + 1] + 7) >> 4;
    return cx == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + 3 * near_row[i - 1] + far_row[i - 1] + 8) >> 4;
}
// the chroma row that is further from output row Y (the row above the first / below the last chroma row is that row again)
LST_HD int jpeg_far_row(int Y, int ds_h)
{
    const int cy = Y >> 1;
    int fy = (Y & 1) ? cy + 1 : cy - 1;
    return fy < 0 ? 0 : (fy > ds_h - 1 ? ds_h - 1 : fy);
}

// ---- jdcolor.c: YCbCr -> RGB, SCALEBITS = 16 --------------------------------------------------------
LST_HD int jpeg_clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
LST_HD uint32_t jpeg_ycc_to_rgb(int y, int cb, int cr)      // -> 0x00RRGGBB
{
    const long long xb = cb - 128, xr = cr - 128;
    const int r = jpeg_clamp255(y + (int)((91881 * xr + 32768) >> 16));                    // FIX(1.40200)
    const int g = jpeg_clamp255(y + (int)(((-22554 * xb + 32768) + (-46802 * xr)) >> 16));   // FIX(0.34414), FIX(0.71414)
    const int b = jpeg_clamp255(y + (int)((116130 * xb + 32768) >> 16));                   // FIX(1.77200)
    return ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
}

}  // namespace lst
