// lst_jpeg.cpp -- baseline JPEG decoder for slice files, bit-compatible with the IJG library's defaults (see lst_jpeg.h).
// Restated from the published algorithms (ITU T.81 for the entropy coding; jidctint.c / jdsample.c / jdcolor.c of
// libjpeg 6b for the sample reconstruction); pinned against cv2.imread in tests/test_jpeg.py.
#include "lst_jpeg.h"

#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "../../include/lst_b200.h"

namespace lst {

namespace {

int fail(std::string *err, int rc, const char *msg)
{
    if (err) *err = msg;
    return rc;
}

// zigzag position -> natural (row-major) position of a coefficient
const uint8_t kNatural[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                              41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                              30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    bool ok = false;
    uint8_t bits[17];
    uint8_t vals[256];
    int32_t maxcode[18];
    int32_t valoff[17];
    uint16_t look[512];   // 9-bit lookahead: (code length << 8) | symbol, 0 = longer code
    int16_t fast[512];    // AC tables: code and magnitude bits together within 9 bits: (value << 8) | (run << 4) | total bits
    uint16_t skip[2048];  // AC tables, blocks that are only parsed: 11-bit lookahead -> bits 0-4 code + magnitude bits to drop,
                          // bits 5-8 run, bit 9 end of block, bit 15 entry valid (code no longer than 11 bits)

    bool build()
    {
        int32_t code = 0;
        int k = 0;
        memset(look, 0, sizeof(look));
        for (int l = 1; l <= 16; l++) {
            valoff[l] = k - code;
            for (int i = 0; i < bits[l]; i++, k++, code++) {
                if (k >= 256 || code >= (1 << l)) return false;
                if (l <= 9) {
                    const int first = code << (9 - l), cnt = 1 << (9 - l);
                    for (int e = 0; e < cnt; e++) look[first + e] = (uint16_t)((l << 8) | vals[k]);
                }
            }
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        for (int i = 0; i < 512; i++) {
            fast[i] = 0;
            const int e = look[i];
            if (!e) continue;
            const int len = e >> 8, rs = e & 255, run = rs >> 4, ss = rs & 15;
            if (ss == 0 || len + ss > 9) continue;
            int v = (i >> (9 - len - ss)) & ((1 << ss) - 1);       // the magnitude bits that follow the code
            if (v < (1 << (ss - 1))) v += -(1 << ss) + 1;
            if (v >= -128 && v <= 127) fast[i] = (int16_t)(v * 256 + run * 16 + len + ss);
        }
        {
            memset(skip, 0, sizeof(skip));
            int32_t c2 = 0;
            int k2 = 0;
            for (int l = 1; l <= 11; l++) {
                for (int i = 0; i < bits[l]; i++, k2++, c2++) {
                    const int rs = vals[k2], run = rs >> 4, ss = rs & 15;
                    const uint16_t e = (uint16_t)(0x8000u | (uint16_t)(l + ss) | (uint16_t)(run << 5) | (uint16_t)((rs == 0) << 9));
                    const int first = c2 << (11 - l), cnt = 1 << (11 - l);
                    for (int j = 0; j < cnt; j++) skip[first + j] = e;
                }
                c2 <<= 1;
            }
        }
        ok = true;
        return true;
    }
};

struct Comp {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int ds_w = 0, ds_h = 0;       // downsampled_width / downsampled_height of the component
    int stride = 0, rows = 0;     // the plane: whole MCUs
    int pred = 0;
};

struct Header {
    int width = 0, height = 0, ncomp = 0;
    Comp comp[3];
    uint16_t qt[4][64];           // natural order
    bool have_qt[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    int restart_interval = 0;
    int hmax = 1, vmax = 1, mcu_w = 8, mcu_h = 8, mcus_x = 0, mcus_y = 0;
    bool adobe = false;
    int adobe_transform = 0;
    bool ycc = true;              // 3 components: YCbCr -> RGB conversion on output
    uint64_t scan_pos = 0;        // first byte of the entropy-coded data
};

inline int be16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

int parse_header(const uint8_t *d, uint64_t n, Header *H, std::string *err)
{
    if (!jpeg_signature(d, n)) return fail(err, LST_ERR_INVALID, "not a JPEG file");
    uint64_t o = 2;
    bool have_sof = false;
    for (;;) {
        if (o + 4 > n) return fail(err, LST_ERR_INVALID, "JPEG: truncated before the scan");
        if (d[o] != 0xff) return fail(err, LST_ERR_INVALID, "JPEG: marker expected");
        while (o < n && d[o] == 0xff) o++;          // fill bytes
        if (o >= n) return fail(err, LST_ERR_INVALID, "JPEG: truncated");
        const int m = d[o++];
        if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) continue;
        if (m == 0xd9) return fail(err, LST_ERR_INVALID, "JPEG: no scan");
        if (o + 2 > n) return fail(err, LST_ERR_INVALID, "JPEG: truncated");
        const int len = be16(d + o);
        if (len < 2 || o + (uint64_t)len > n) return fail(err, LST_ERR_INVALID, "JPEG: bad segment length");
        const uint8_t *s = d + o + 2;
        const int sl = len - 2;
        if (m == 0xc0 || m == 0xc1) {
            if (sl < 6) return fail(err, LST_ERR_INVALID, "JPEG: bad frame header");
            if (s[0] != 8) return fail(err, LST_ERR_UNSUPPORTED, "JPEG: only 8-bit samples are supported");
            H->height = be16(s + 1);
            H->width = be16(s + 3);
            H->ncomp = s[5];
            if (H->ncomp != 1 && H->ncomp != 3) return fail(err, LST_ERR_UNSUPPORTED, "JPEG: only 1- and 3-component images are supported");
            if (sl < 6 + 3 * H->ncomp || H->width <= 0 || H->height <= 0) return fail(err, LST_ERR_INVALID, "JPEG: bad frame header");
            if ((int64_t)H->width * H->height > ((int64_t)1 << 28)) return fail(err, LST_ERR_UNSUPPORTED, "JPEG: image larger than 2^28 pixels");
            for (int c = 0; c < H->ncomp; c++) {
                H->comp[c].id = s[6 + 3 * c];
                H->comp[c].h = s[7 + 3 * c] >> 4;
                H->comp[c].v = s[7 + 3 * c] & 15;
                H->comp[c].tq = s[8 + 3 * c] & 3;
            }
            have_sof = true;
        } else if (m == 0xc2 || m == 0xc3 || (m >= 0xc5 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc)) {
            return fail(err, LST_ERR_UNSUPPORTED, "JPEG: only baseline / extended sequential Huffman coding is supported (this file is progressive, lossless or arithmetic coded)");
        } else if (m == 0xcc) {
            return fail(err, LST_ERR_UNSUPPORTED, "JPEG: arithmetic coding is not supported");
        } else if (m == 0xc4) {
            int p = 0;
            while (p < sl) {
                if (p + 17 > sl) return fail(err, LST_ERR_INVALID, "JPEG: bad Huffman table");
                const int tc = s[p] >> 4, th = s[p] & 15;
                if (tc > 1 || th > 3) return fail(err, LST_ERR_INVALID, "JPEG: bad Huffman table id");
                Huff &T = tc ? H->ac[th] : H->dc[th];
                int cnt = 0;
                T.bits[0] = 0;
                for (int i = 1; i <= 16; i++) cnt += (T.bits[i] = s[p + i]);
                if (cnt > 256 || p + 17 + cnt > sl) return fail(err, LST_ERR_INVALID, "JPEG: bad Huffman table");
                memset(T.vals, 0, sizeof(T.vals));
                memcpy(T.vals, s + p + 17, (size_t)cnt);
                if (!T.build()) return fail(err, LST_ERR_INVALID, "JPEG: bad Huffman code lengths");
                p += 17 + cnt;
            }
        } else if (m == 0xdb) {
            int p = 0;
            while (p < sl) {
                const int pq = s[p] >> 4, tq = s[p] & 15;
                if (tq > 3 || pq > 1 || p + 1 + 64 * (pq + 1) > sl) return fail(err, LST_ERR_INVALID, "JPEG: bad quantisation table");
                for (int i = 0; i < 64; i++) H->qt[tq][kNatural[i]] = (uint16_t)(pq ? be16(s + p + 1 + 2 * i) : s[p + 1 + i]);
                H->have_qt[tq] = true;
                p += 1 + 64 * (pq + 1);
            }
        } else if (m == 0xdd) {
            if (sl < 2) return fail(err, LST_ERR_INVALID, "JPEG: bad restart interval");
            H->restart_interval = be16(s);
        } else if (m == 0xee) {
            if (sl >= 12 && !memcmp(s, "Adobe", 5)) {
                H->adobe = true;
                H->adobe_transform = s[11];
            }
        } else if (m == 0xda) {
            if (!have_sof) return fail(err, LST_ERR_INVALID, "JPEG: scan before the frame header");
            if (sl < 1 || s[0] != H->ncomp || sl < 1 + 2 * H->ncomp + 3)
                return fail(err, LST_ERR_UNSUPPORTED, "JPEG: only one interleaved scan with all components is supported");
            for (int c = 0; c < H->ncomp; c++) {
                if (s[1 + 2 * c] != H->comp[c].id) return fail(err, LST_ERR_UNSUPPORTED, "JPEG: scan components out of frame order");
                H->comp[c].td = s[2 + 2 * c] >> 4;
                H->comp[c].ta = s[2 + 2 * c] & 15;
                if (H->comp[c].td > 3 || H->comp[c].ta > 3 || !H->dc[H->comp[c].td].ok || !H->ac[H->comp[c].ta].ok)
                    return fail(err, LST_ERR_INVALID, "JPEG: scan refers to an undefined Huffman table");
                if (!H->have_qt[H->comp[c].tq]) return fail(err, LST_ERR_INVALID, "JPEG: undefined quantisation table");
            }
            H->scan_pos = o + (uint64_t)len;
            break;
        }
        o += (uint64_t)len;
    }
    // sampling: 4:4:4, 4:2:2 (h2v1), 4:2:0 (h2v2) -- the ratios whose upsampling libjpeg 6b and libjpeg-turbo share
    if (H->ncomp == 1) {
        H->comp[0].h = H->comp[0].v = 1;    // a single-component scan is never interleaved: one block per MCU
    } else {
        if (H->comp[1].h != 1 || H->comp[1].v != 1 || H->comp[2].h != 1 || H->comp[2].v != 1 ||
            !((H->comp[0].h == 1 && H->comp[0].v == 1) || (H->comp[0].h == 2 && H->comp[0].v == 1) || (H->comp[0].h == 2 && H->comp[0].v == 2)))
            return fail(err, LST_ERR_UNSUPPORTED, "JPEG: chroma sampling other than 4:4:4, 4:2:2 and 4:2:0 is not supported");
        // colour space as jdapimin.c default_decompress_parms decides it for three components
        if (H->adobe) H->ycc = H->adobe_transform != 0;
        else H->ycc = !(H->comp[0].id == 'R' && H->comp[1].id == 'G' && H->comp[2].id == 'B');
    }
    H->hmax = H->comp[0].h;
    H->vmax = H->comp[0].v;
    H->mcu_w = 8 * H->hmax;
    H->mcu_h = 8 * H->vmax;
    H->mcus_x = (H->width + H->mcu_w - 1) / H->mcu_w;
    H->mcus_y = (H->height + H->mcu_h - 1) / H->mcu_h;
    for (int c = 0; c < H->ncomp; c++) {
        Comp &C = H->comp[c];
        C.ds_w = (H->width * C.h + H->hmax - 1) / H->hmax;
        C.ds_h = (H->height * C.v + H->vmax - 1) / H->vmax;
        C.stride = H->mcus_x * C.h * 8;
        C.rows = H->mcus_y * C.v * 8;
    }
    return LST_OK;
}

// ---- entropy decoding (ITU T.81 F.2.2) ----------------------------------------------------------
struct BitReader {
    const uint8_t *d;
    uint64_t n, pos;
    uint64_t acc = 0;
    int nbits = 0;
    bool marker = false;   // a marker was reached: zeros are supplied from here on (as libjpeg does)

    inline void fill()
    {
        if (!marker && pos + 8 <= n && nbits <= 56) {
            uint64_t w;
            memcpy(&w, d + pos, 8);
            const uint64_t inv = ~w;
            if (!((inv - 0x0101010101010101ull) & ~inv & 0x8080808080808080ull)) {     // no 0xff among the next eight bytes
                w = __builtin_bswap64(w);
                const int k = (64 - nbits) >> 3;
                acc |= (w >> (64 - 8 * k)) << (64 - nbits - 8 * k);
                pos += (uint64_t)k;
                nbits += 8 * k;
                return;
            }
        }
        while (nbits <= 56) {
            uint32_t b = 0;
            if (!marker) {
                if (pos >= n) {
                    marker = true;
                } else {
                    b = d[pos];
                    if (b == 0xff) {
                        if (pos + 1 < n && d[pos + 1] == 0) pos += 2;
                        else marker = true, b = 0;
                    } else {
                        pos++;
                    }
                }
            }
            acc |= (uint64_t)b << (56 - nbits);
            nbits += 8;
        }
    }
    inline uint32_t peek(int k) { return (uint32_t)(acc >> (64 - k)); }
    inline void skip(int k) { acc <<= k, nbits -= k; }
    // a Huffman code (<= 16 bits) and its magnitude bits (<= 15) never need more than 31 bits
    inline void need32()
    {
        if (nbits < 32) fill();
    }
    inline int decode(const Huff &T)
    {
        const uint32_t e = T.look[peek(9)];
        if (e) {
            skip((int)(e >> 8));
            return (int)(e & 255u);
        }
        int l = 10;
        int32_t code = (int32_t)peek(10);
        while (l <= 16 && code > T.maxcode[l]) {
            l++;
            code = (int32_t)peek(l);
        }
        if (l > 16) {
            skip(16);
            return 0;     // corrupt data: a zero-length code, as libjpeg substitutes
        }
        skip(l);
        return T.vals[(code + T.valoff[l]) & 255];
    }
    inline int receive_extend(int s)
    {
        const int v = (int)peek(s);
        skip(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
    // restart marker: drop the padding bits and step over RSTn
    void restart()
    {
        acc = 0, nbits = 0;
        while (pos + 1 < n && !(d[pos] == 0xff && d[pos + 1] >= 0xd0 && d[pos + 1] <= 0xd7)) pos++;
        if (pos + 1 < n) pos += 2;
        marker = false;
    }
};

// One block of the scan: the DC difference feeds the prediction; STORE: the coefficients go to coef[64] in natural order,
// else the AC symbols are only dropped, code and magnitude bits at once.
template <bool STORE>
inline void decode_block(BitReader &br, const Huff &DC, const Huff &AC, int &pred, int16_t *coef)
{
    br.need32();
    int s = br.decode(DC);
    if (s) {
        if (s > 15) s = 15;
        pred += br.receive_extend(s);
    }
    if (!STORE) {
        for (int k = 1; k < 64; k++) {
            br.need32();
            const uint32_t e = AC.skip[br.peek(11)];
            if (e & 0x8000u) {
                br.skip((int)(e & 31u));
                if (e & 0x200u) break;
                k += (int)((e >> 5) & 15u);
                continue;
            }
            const int rs = br.decode(AC);
            if (rs & 15) {
                k += rs >> 4;
                br.skip(rs & 15);
            } else {
                if ((rs >> 4) != 15) break;
                k += 15;
            }
        }
        return;
    }
    memset(coef, 0, sizeof(int16_t) * 64);
    coef[0] = (int16_t)pred;
    for (int k = 1; k < 64; k++) {
        br.need32();
        const int f = AC.fast[br.peek(9)];
        if (f) {                     // code and value in one look-up
            k += (f >> 4) & 15;
            br.skip(f & 15);
            if (k < 64) coef[kNatural[k]] = (int16_t)(f >> 8);
            continue;
        }
        const int rs = br.decode(AC);
        const int rr = rs >> 4, ss = rs & 15;
        if (ss) {
            k += rr;
            const int v = br.receive_extend(ss);
            if (k < 64) coef[kNatural[k]] = (int16_t)v;
        } else {
            if (rr != 15) break;
            k += 15;
        }
    }
}

struct Scratch {
    std::vector<uint8_t> plane[3];
    std::vector<uint8_t> need;     // per MCU
};

}  // namespace

int jpeg_parse(const uint8_t *data, uint64_t size, int page, TiffLayout *out, std::string *err)
{
    if (page != 0) return fail(err, LST_ERR_INVALID, "a JPEG file holds one image");
    Header H;
    int rc = parse_header(data, size, &H, err);
    if (rc != LST_OK) return rc;
    TiffLayout L;
    L.jpeg = true;
    L.rgb = H.ncomp == 3;
    L.width = H.width;
    L.height = H.height;
    L.bits = L.rgb ? 32 : 8;
    L.sample_format = 1;
    L.big_endian = false;
    L.rows_per_strip = H.height;
    L.compression = 7;             // TIFF's code for JPEG
    L.row_bytes = (uint64_t)H.width * (L.bits / 8);
    L.file_size = size;
    L.strip_off.push_back(H.scan_pos);
    L.strip_len.push_back(size - H.scan_pos);
    *out = L;
    return LST_OK;
}

int jpeg_copy_rects(const uint8_t *data, uint64_t size, const TiffRect *rects, int n, std::string *err)
{
    Header H;
    int rc = parse_header(data, size, &H, err);
    if (rc != LST_OK) return rc;
    static thread_local Scratch S;
    try {
        for (int c = 0; c < H.ncomp; c++) {
            const size_t need = (size_t)H.comp[c].stride * H.comp[c].rows;
            if (S.plane[c].size() < need) S.plane[c].resize(need);
        }
        S.need.assign((size_t)H.mcus_x * H.mcus_y, 0);
    } catch (const std::bad_alloc &) {
        return fail(err, LST_ERR_INVALID, "JPEG: out of memory for the sample planes");
    }
    // the MCUs whose samples the rectangles need: two pixels of margin cover the chroma context of the upsampling filters
    int last_row = -1;
    for (int i = 0; i < n; i++) {
        const TiffRect &r = rects[i];
        if (r.w <= 0 || r.h <= 0) continue;
        if (r.x < 0 || r.y < 0 || r.x + r.w > H.width || r.y + r.h > H.height) return fail(err, LST_ERR_INVALID, "rectangle outside the image");
        const int mx0 = std::max(0, (r.x - 2) / H.mcu_w), mx1 = std::min(H.mcus_x - 1, (r.x + r.w + 1) / H.mcu_w);
        const int my0 = std::max(0, (r.y - 2) / H.mcu_h), my1 = std::min(H.mcus_y - 1, (r.y + r.h + 1) / H.mcu_h);
        for (int my = my0; my <= my1; my++) memset(&S.need[(size_t)my * H.mcus_x + mx0], 1, (size_t)(mx1 - mx0 + 1));
        last_row = std::max(last_row, my1);
    }
    if (last_row < 0) return LST_OK;

    BitReader br{data, size, H.scan_pos};
    for (int c = 0; c < H.ncomp; c++) H.comp[c].pred = 0;
    int16_t coef[64];
    int to_restart = H.restart_interval;
    for (int my = 0; my <= last_row; my++) {
        for (int mx = 0; mx < H.mcus_x; mx++) {
            if (H.restart_interval) {
                if (to_restart == 0) {
                    br.restart();
                    for (int c = 0; c < H.ncomp; c++) H.comp[c].pred = 0;
                    to_restart = H.restart_interval;
                }
                to_restart--;
            }
            const bool need = S.need[(size_t)my * H.mcus_x + mx] != 0;
            for (int c = 0; c < H.ncomp; c++) {
                Comp &C = H.comp[c];
                const Huff &DC = H.dc[C.td], &AC = H.ac[C.ta];
                for (int by = 0; by < C.v; by++)
                    for (int bx = 0; bx < C.h; bx++) {
                        if (need) decode_block<true>(br, DC, AC, C.pred, coef);
                        else decode_block<false>(br, DC, AC, C.pred, nullptr);
                        if (need)
                            jpeg_idct_islow(coef, H.qt[C.tq],
                                       S.plane[c].data() + (size_t)((my * C.v + by) * 8) * C.stride + (size_t)(mx * C.h + bx) * 8, C.stride);
                    }
            }
        }
    }

    // output: upsample + colour-convert the pixels of every rectangle
    for (int i = 0; i < n; i++) {
        const TiffRect &r = rects[i];
        if (r.w <= 0 || r.h <= 0) continue;
        if (H.ncomp == 1) {
            for (int y = 0; y < r.h; y++)
                memcpy(r.dst + (size_t)y * r.w, S.plane[0].data() + (size_t)(r.y + y) * H.comp[0].stride + r.x, (size_t)r.w);
            continue;
        }
        const Comp &Cb = H.comp[1], &Cr = H.comp[2];
        const bool h2 = H.hmax == 2, v2 = H.vmax == 2;
        for (int y = 0; y < r.h; y++) {
            const int Y = r.y + y;
            const uint8_t *yrow = S.plane[0].data() + (size_t)Y * H.comp[0].stride;
            const int cy = v2 ? Y >> 1 : Y, fy = v2 ? jpeg_far_row(Y, Cb.ds_h) : Y;
            const uint8_t *bn = S.plane[1].data() + (size_t)cy * Cb.stride, *bf = S.plane[1].data() + (size_t)fy * Cb.stride;
            const uint8_t *rn = S.plane[2].data() + (size_t)cy * Cr.stride, *rf = S.plane[2].data() + (size_t)fy * Cr.stride;
            uint32_t *o = (uint32_t *)(r.dst + (size_t)y * r.w * 4);
            for (int x = 0; x < r.w; x++) {
                const int X = r.x + x;
                int c0 = yrow[X], c1, c2;
                if (!h2) c1 = bn[X], c2 = rn[X];
                else if (!v2) c1 = jpeg_up_h2v1(bn, 0, Cb.ds_w, X), c2 = jpeg_up_h2v1(rn, 0, Cr.ds_w, X);
                else c1 = jpeg_up_h2v2(bn, bf, 0, Cb.ds_w, X), c2 = jpeg_up_h2v2(rn, rf, 0, Cr.ds_w, X);
                o[x] = H.ycc ? jpeg_ycc_to_rgb(c0, c1, c2) : ((uint32_t)c0 << 16) | ((uint32_t)c1 << 8) | (uint32_t)c2;   // int32 0x00RRGGBB
            }
        }
    }
    return LST_OK;
}

int jpeg_geometry(const uint8_t *data, uint64_t size, JpegGeom *g, std::string *err)
{
    Header H;
    int rc = parse_header(data, size, &H, err);
    if (rc != LST_OK) return rc;
    g->width = H.width, g->height = H.height, g->ncomp = H.ncomp, g->hmax = H.hmax, g->vmax = H.vmax, g->ycc = H.ycc ? 1 : 0;
    return LST_OK;
}

int jpeg_decode_coeffs(const uint8_t *data, uint64_t size, const JpegGeom &g, const JpegRegion *regs, int16_t *const *coef, int n,
                       uint16_t *qt_out, std::string *err)
{
    Header H;
    int rc = parse_header(data, size, &H, err);
    if (rc != LST_OK) return rc;
    JpegGeom mine;
    mine.width = H.width, mine.height = H.height, mine.ncomp = H.ncomp, mine.hmax = H.hmax, mine.vmax = H.vmax, mine.ycc = H.ycc ? 1 : 0;
    if (!(mine == g)) return fail(err, LST_ERR_INVALID, "JPEG slice differs from the stack's size, components or chroma sampling");
    if (n > 8) return fail(err, LST_ERR_INVALID, "too many regions");
    for (int c = 0; c < 3; c++)
        for (int i = 0; i < 64; i++) qt_out[64 * c + i] = c < H.ncomp ? H.qt[H.comp[c].tq][i] : (uint16_t)0;
    static thread_local std::vector<uint8_t> mask;        // per MCU: the regions it belongs to
    mask.assign((size_t)H.mcus_x * H.mcus_y, 0);
    int last_row = -1;
    for (int k = 0; k < n; k++) {
        const JpegRegion &r = regs[k];
        if (r.nmx <= 0 || r.nmy <= 0) continue;
        if (r.mx0 < 0 || r.my0 < 0 || r.mx0 + r.nmx > H.mcus_x || r.my0 + r.nmy > H.mcus_y) return fail(err, LST_ERR_INVALID, "region outside the image");
        for (int my = r.my0; my < r.my0 + r.nmy; my++)
            for (int mx = r.mx0; mx < r.mx0 + r.nmx; mx++) mask[(size_t)my * H.mcus_x + mx] |= (uint8_t)(1u << k);
        last_row = std::max(last_row, r.my0 + r.nmy - 1);
    }
    const int bpm = jpeg_blocks_per_mcu(g);
    int16_t mcu[6 * 64];
    BitReader br{data, size, H.scan_pos};
    for (int c = 0; c < H.ncomp; c++) H.comp[c].pred = 0;
    int to_restart = H.restart_interval;
    for (int my = 0; my <= last_row; my++) {
        for (int mx = 0; mx < H.mcus_x; mx++) {
            if (H.restart_interval) {
                if (to_restart == 0) {
                    br.restart();
                    for (int c = 0; c < H.ncomp; c++) H.comp[c].pred = 0;
                    to_restart = H.restart_interval;
                }
                to_restart--;
            }
            const uint8_t m = mask[(size_t)my * H.mcus_x + mx];
            int blk = 0;
            for (int c = 0; c < H.ncomp; c++) {
                Comp &C = H.comp[c];
                const Huff &DC = H.dc[C.td], &AC = H.ac[C.ta];
                for (int b = 0; b < C.v * C.h; b++, blk++) {
                    if (m) decode_block<true>(br, DC, AC, C.pred, mcu + 64 * blk);
                    else decode_block<false>(br, DC, AC, C.pred, nullptr);
                }
            }
            for (int k = 0; m && k < n; k++)
                if (m & (1u << k)) {
                    const JpegRegion &r = regs[k];
                    memcpy(coef[k] + ((size_t)(my - r.my0) * r.nmx + (mx - r.mx0)) * bpm * 64, mcu, sizeof(int16_t) * 64 * bpm);
                }
        }
    }
    return LST_OK;
}

}  // namespace lst
