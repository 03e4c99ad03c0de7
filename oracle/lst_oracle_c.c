/* CPU oracle helper -- TEST INFRASTRUCTURE ONLY (see oracle/lst_oracle.py header).
 *
 * Exact integer cross-correlation of an 8-bit window with an 8-bit template:
 *     out[y][x] = sum_{v<th} sum_{u<tw} win[y+v][x+u] * tpl[v][u]
 * This is the quantity cv::crossCorr (OpenCV imgproc templmatch.cpp, reached from
 * Laser_Spot_Track4.java:2560 through javacv 1.5.8) approximates with a float32 DFT.
 * PARITY UNPINNED: the reference ships no vectors for it; tests pin it against a
 * pure-numpy evaluation and against cv2.matchTemplate (within OpenCV's DFT noise).
 */
#include <stdint.h>

void lst_oracle_xcorr_u8(const uint8_t *win, int sw, int sh, int wstride,
                         const uint8_t *tpl, int tw, int th, int64_t *out)
{
    int rw = sw - tw + 1, rh = sh - th + 1;
    for (int y = 0; y < rh; ++y) {
        for (int x = 0; x < rw; ++x) {
            int64_t acc = 0;
            for (int v = 0; v < th; ++v) {
                const uint8_t *wr = win + (long)(y + v) * wstride + x;
                const uint8_t *tr = tpl + (long)v * tw;
                uint32_t row = 0; /* tw*255*255 < 2^32 for tw <= 66051 */
                for (int u = 0; u < tw; ++u) row += (uint32_t)wr[u] * (uint32_t)tr[u];
                acc += row;
            }
            out[(long)y * rw + x] = acc;
        }
    }
}
