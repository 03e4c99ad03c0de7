// lst_tiff.cpp -- minimal reader for the uncompressed grey-scale TIFF files of a time-lapse (see lst_tiff.h).
#include "lst_tiff.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <vector>

#include "../../include/lst_b200.h"
#include "lst_jpeg.h"

namespace lst {

namespace {

struct Reader {
    const uint8_t *d;
    uint64_t n;
    bool be;
    bool ok(uint64_t off, uint64_t len) const { return off <= n && len <= n - off; }
    uint16_t u16(uint64_t o) const { return be ? (uint16_t)((d[o] << 8) | d[o + 1]) : (uint16_t)(d[o] | (d[o + 1] << 8)); }
    uint32_t u32(uint64_t o) const
    {
        return be ? ((uint32_t)d[o] << 24) | ((uint32_t)d[o + 1] << 16) | ((uint32_t)d[o + 2] << 8) | d[o + 3]
                  : ((uint32_t)d[o + 3] << 24) | ((uint32_t)d[o + 2] << 16) | ((uint32_t)d[o + 1] << 8) | d[o];
    }
};

int fail(std::string *err, int code, const char *msg)
{
    if (err) *err = msg;
    return code;
}

const int kTypeSize[13] = {0, 1, 1, 2, 4, 8, 1, 1, 2, 4, 8, 4, 8};

// value i of an IFD entry of type BYTE/SHORT/LONG (inline when it fits four bytes)
bool entry_value(const Reader &r, uint64_t entry, uint32_t i, uint32_t *out)
{
    const uint16_t type = r.u16(entry + 2);
    const uint32_t count = r.u32(entry + 4);
    if (type == 0 || type > 12 || i >= count) return false;
    const int ts = kTypeSize[type];
    uint64_t base = entry + 8;
    if ((uint64_t)ts * count > 4) {
        base = r.u32(entry + 8);
        if (!r.ok(base, (uint64_t)ts * count)) return false;
    }
    const uint64_t o = base + (uint64_t)ts * i;
    if (type == 1 || type == 6 || type == 7) *out = r.d[o];
    else if (type == 3 || type == 8) *out = r.u16(o);
    else if (type == 4 || type == 9) *out = r.u32(o);
    else return false;
    return true;
}

}  // namespace

static int png_parse(const uint8_t *d, uint64_t n, int page, TiffLayout *out, std::string *err)
{
    if (page != 0) return fail(err, LST_ERR_INVALID, "a PNG file holds one image");
    TiffLayout L;
    L.png = true;
    L.big_endian = true;   // PNG samples are in network byte order
    L.file_size = n;
    uint64_t o = 8;
    bool have_hdr = false, done = false;
    while (!done && o + 12 <= n) {
        const uint32_t len = ((uint32_t)d[o] << 24) | ((uint32_t)d[o + 1] << 16) | ((uint32_t)d[o + 2] << 8) | d[o + 3];
        const uint8_t *ty = d + o + 4;
        if (o + 12 + (uint64_t)len > n) return fail(err, LST_ERR_INVALID, "PNG chunk outside the file");
        if (!memcmp(ty, "IHDR", 4)) {
            if (len < 13) return fail(err, LST_ERR_INVALID, "bad PNG header");
            const uint8_t *h = d + o + 8;
            L.width = (int)(((uint32_t)h[0] << 24) | ((uint32_t)h[1] << 16) | ((uint32_t)h[2] << 8) | h[3]);
            L.height = (int)(((uint32_t)h[4] << 24) | ((uint32_t)h[5] << 16) | ((uint32_t)h[6] << 8) | h[7]);
            L.bits = h[8];
            if (h[9] != 0) return fail(err, LST_ERR_UNSUPPORTED, "only grey-scale PNG is supported (RGB / palette / alpha are not on this path)");
            if (L.bits != 8 && L.bits != 16) return fail(err, LST_ERR_UNSUPPORTED, "only 8- and 16-bit PNG samples are supported");
            if (h[12] != 0) return fail(err, LST_ERR_UNSUPPORTED, "interlaced PNG is not supported");
            have_hdr = true;
        } else if (!memcmp(ty, "IDAT", 4)) {
            L.strip_off.push_back(o + 8);
            L.strip_len.push_back(len);
        } else if (!memcmp(ty, "IEND", 4)) {
            done = true;
        }
        o += 12 + (uint64_t)len;
    }
    if (!have_hdr || L.width <= 0 || L.height <= 0 || L.strip_off.empty()) return fail(err, LST_ERR_INVALID, "PNG without header or image data");
    L.sample_format = 1;
    L.row_bytes = (uint64_t)L.width * (L.bits / 8);
    L.rows_per_strip = L.height;
    L.compression = 8;
    *out = L;
    return LST_OK;
}

int tiff_parse(const uint8_t *data, uint64_t size, int page, TiffLayout *out, std::string *err)
{
    static const uint8_t kPngSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (data && size >= 8 && !memcmp(data, kPngSig, 8)) return png_parse(data, size, page, out, err);
    if (jpeg_signature(data, size)) return jpeg_parse(data, size, page, out, err);
    if (!data || size < 8 || page < 0) return fail(err, LST_ERR_INVALID, "not a TIFF file (too short)");
    Reader r{data, size, false};
    if (data[0] == 'I' && data[1] == 'I') r.be = false;
    else if (data[0] == 'M' && data[1] == 'M') r.be = true;
    else return fail(err, LST_ERR_INVALID, "not a TIFF file (byte-order mark)");
    const uint16_t magic = r.u16(2);
    if (magic == 43) return fail(err, LST_ERR_UNSUPPORTED, "BigTIFF is not supported");
    if (magic != 42) return fail(err, LST_ERR_INVALID, "not a TIFF file (magic)");
    uint64_t ifd = r.u32(4);
    int walked = 0;
    // walk the IFD chain to the requested page; an ImageJ stack has one IFD and is addressed by offset
    for (;;) {
        if (!r.ok(ifd, 2)) return fail(err, LST_ERR_INVALID, "TIFF directory outside the file");
        const uint16_t ne = r.u16(ifd);
        if (!r.ok(ifd + 2, (uint64_t)ne * 12 + 4)) return fail(err, LST_ERR_INVALID, "TIFF directory truncated");
        const uint32_t next = r.u32(ifd + 2 + (uint64_t)ne * 12);
        if (walked == page || next == 0) break;
        ifd = next;
        walked++;
    }
    TiffLayout L;
    L.big_endian = r.be;
    L.file_size = size;
    const uint16_t ne = r.u16(ifd);
    uint32_t compression = 1, photometric = 1, spp = 1, planar = 1, rps = 0xffffffffu;
    uint64_t strips_entry = 0, counts_entry = 0, desc_off = 0;
    uint32_t predictor = 1;
    uint32_t nstrips = 0, desc_len = 0, bits = 1, fmt = 1;
    bool tiled = false;
    for (uint16_t e = 0; e < ne; e++) {
        const uint64_t en = ifd + 2 + (uint64_t)e * 12;
        const uint16_t tag = r.u16(en);
        uint32_t v = 0;
        switch (tag) {
            case 256: if (entry_value(r, en, 0, &v)) L.width = (int)v; break;
            case 257: if (entry_value(r, en, 0, &v)) L.height = (int)v; break;
            case 258: if (entry_value(r, en, 0, &v)) bits = v; break;
            case 259: if (entry_value(r, en, 0, &v)) compression = v; break;
            case 262: if (entry_value(r, en, 0, &v)) photometric = v; break;
            case 270:
                desc_len = r.u32(en + 4);
                desc_off = desc_len > 4 ? r.u32(en + 8) : en + 8;
                break;
            case 273:
                strips_entry = en;
                nstrips = r.u32(en + 4);
                break;
            case 277: if (entry_value(r, en, 0, &v)) spp = v; break;
            case 279: counts_entry = en; break;
            case 317: if (entry_value(r, en, 0, &v)) predictor = v; break;
            case 278: if (entry_value(r, en, 0, &v)) rps = v; break;
            case 284: if (entry_value(r, en, 0, &v)) planar = v; break;
            case 322: case 323: case 324: case 325: tiled = true; break;
            case 339: if (entry_value(r, en, 0, &v)) fmt = v; break;
            default: break;
        }
    }
    (void)planar;
    if (L.width <= 0 || L.height <= 0) return fail(err, LST_ERR_INVALID, "TIFF without image size");
    if (tiled) return fail(err, LST_ERR_UNSUPPORTED, "tiled TIFF is not supported");
    if (compression != 1 && compression != 5 && compression != 8 && compression != 32946 && compression != 32773)
        return fail(err, LST_ERR_UNSUPPORTED, "TIFF compression scheme is not supported (decode the file and use lst_track_ptrs)");
    if (predictor != 1 && !(predictor == 2 && bits != 32))
        return fail(err, LST_ERR_UNSUPPORTED, "TIFF predictor is not supported");
    if (compression != 1 && !counts_entry) return fail(err, LST_ERR_INVALID, "compressed TIFF without strip byte counts");
    if (spp != 1) return fail(err, LST_ERR_UNSUPPORTED, "only one sample per pixel (RGB is not on this path, LST4:172)");
    if (photometric != 1) return fail(err, LST_ERR_UNSUPPORTED, "only BlackIsZero grey-scale TIFF is supported");
    if (!((bits == 8 && fmt == 1) || (bits == 16 && fmt == 1) || (bits == 32 && fmt == 3)))
        return fail(err, LST_ERR_UNSUPPORTED, "only 8/16-bit unsigned and 32-bit float samples are supported");
    if (!strips_entry || nstrips == 0) return fail(err, LST_ERR_INVALID, "TIFF without strip offsets");
    L.bits = (int)bits;
    L.sample_format = (int)fmt;
    L.row_bytes = (uint64_t)L.width * (bits / 8);
    if (rps == 0 || rps > (uint32_t)L.height) rps = (uint32_t)L.height;
    L.rows_per_strip = (int)rps;
    const uint32_t want_strips = ((uint32_t)L.height + rps - 1) / rps;
    if (nstrips < want_strips) return fail(err, LST_ERR_INVALID, "TIFF strip table shorter than the image");
    L.strip_off.resize(want_strips);
    for (uint32_t i = 0; i < want_strips; i++) {
        uint32_t v = 0;
        if (!entry_value(r, strips_entry, i, &v)) return fail(err, LST_ERR_INVALID, "bad TIFF strip table");
        L.strip_off[i] = v;
    }
    if (desc_len >= 8 && r.ok(desc_off, desc_len) && memcmp(data + desc_off, "ImageJ=", 7) == 0) {
        std::string d((const char *)data + desc_off, desc_len);
        size_t p = d.find("images=");
        if (p != std::string::npos) L.imagej_images = atoi(d.c_str() + p + 7);
    }
    if (walked < page) {
        // fewer directories than pages: slices of an ImageJ stack follow the first one back to back
        if (L.imagej_images > page && want_strips == 1) {
            L.strip_off[0] += (uint64_t)page * L.row_bytes * (uint64_t)L.height;
        } else {
            return fail(err, LST_ERR_INVALID, "TIFF page index beyond the last directory");
        }
    }
    L.compression = (int)compression;
    L.predictor = (int)predictor;
    L.strip_len.resize(want_strips);
    for (uint32_t i = 0; i < want_strips; i++) {
        const uint64_t rows = i + 1 < want_strips ? rps : (uint64_t)L.height - (uint64_t)i * rps;
        uint32_t v = 0;
        L.strip_len[i] = rows * L.row_bytes;
        if (compression != 1) {
            if (!entry_value(r, counts_entry, i, &v)) return fail(err, LST_ERR_INVALID, "bad TIFF strip byte counts");
            L.strip_len[i] = v;
        }
        if (!r.ok(L.strip_off[i], L.strip_len[i])) return fail(err, LST_ERR_INVALID, "TIFF pixel data outside the file");
    }
    if (walked < page && compression != 1) return fail(err, LST_ERR_INVALID, "TIFF page index beyond the last directory");
    *out = L;
    return LST_OK;
}

int FileMap::open(const char *path, std::string *err)
{
    close();
    int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) {
        if (err) *err = std::string("cannot open ") + path;
        return LST_ERR_INVALID;
    }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) {
        ::close(fd);
        if (err) *err = std::string("cannot stat ") + path;
        return LST_ERR_INVALID;
    }
    void *p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) {
        ::close(fd);
        if (err) *err = std::string("cannot map ") + path;
        return LST_ERR_INVALID;
    }
    this->fd = fd;
    data = (const uint8_t *)p;
    size = (uint64_t)st.st_size;
    return LST_OK;
}

void FileMap::close()
{
    if (data) munmap((void *)data, (size_t)size);
    if (fd >= 0) ::close(fd);
    fd = -1;
    data = nullptr;
    size = 0;
}

namespace {

bool read_at(int fd, uint8_t *d, uint64_t off, size_t len)
{
    while (len > 0) {
        ssize_t k = pread(fd, d, len, (off_t)off);
        if (k <= 0) return false;
        d += k;
        off += (uint64_t)k;
        len -= (size_t)k;
    }
    return true;
}

// PackBits (TIFF 6.0 section 9)
bool unpack_bits(const uint8_t *src, size_t n, uint8_t *dst, size_t want)
{
    size_t i = 0, o = 0;
    while (i < n && o < want) {
        const int8_t c = (int8_t)src[i++];
        if (c >= 0) {
            const size_t k = (size_t)c + 1;
            if (i + k > n || o + k > want) return false;
            memcpy(dst + o, src + i, k);
            i += k;
            o += k;
        } else if (c != -128) {
            const size_t k = (size_t)(1 - (int)c);
            if (i >= n || o + k > want) return false;
            memset(dst + o, src[i++], k);
            o += k;
        }
    }
    return o == want;
}

// TIFF LZW (section 13): MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, the code
// width grows one code early ("early change"), as every current writer does.
bool unpack_lzw(const uint8_t *src, size_t n, uint8_t *dst, size_t want)
{
    struct Ent {
        int32_t prev;
        uint16_t len;
        uint8_t first, last;
    };
    std::vector<Ent> tab(4096);
    for (int i = 0; i < 256; i++) tab[i] = Ent{-1, 1, (uint8_t)i, (uint8_t)i};
    int next = 258, width = 9, prev = -1;
    uint64_t acc = 0;
    int nbits = 0;
    size_t i = 0, o = 0;
    while (o < want) {
        while (nbits < width && i < n) {
            acc = (acc << 8) | src[i++];
            nbits += 8;
        }
        if (nbits < width) break;
        const int code = (int)((acc >> (nbits - width)) & ((1u << width) - 1));
        nbits -= width;
        if (code == 257) break;
        if (code == 256) {
            next = 258;
            width = 9;
            prev = -1;
            continue;
        }
        int cur = code;
        if (prev < 0) {
            if (code >= 256) return false;
        } else {
            if (code > next || next >= 4096) return false;
            // new entry = string(prev) + first char of string(cur) (or of string(prev) when cur is the entry itself)
            const uint8_t fc = code < next ? tab[code].first : tab[prev].first;
            tab[next] = Ent{prev, (uint16_t)(tab[prev].len + 1), tab[prev].first, fc};
            next++;
        }
        const size_t len = tab[cur].len;
        if (o + len > want) {
            // the last string may run past the strip's rows when the writer padded: copy what fits
            std::vector<uint8_t> tmp(len);
            size_t k = len;
            for (int e = cur; e >= 0; e = tab[e].prev) tmp[--k] = tab[e].last;
            memcpy(dst + o, tmp.data(), want - o);
            o = want;
            break;
        }
        size_t k = o + len;
        for (int e = cur; e >= 0; e = tab[e].prev) dst[--k] = tab[e].last;
        o += len;
        prev = cur;
        if (next + 1 >= (1 << width) && width < 12) width++;
    }
    return o == want;
}

// One strip, decoded and with the predictor undone, as rows of L.row_bytes in the file's byte order.
bool decode_strip(const FileMap &fm, const TiffLayout &L, int strip, std::vector<uint8_t> *raw, std::vector<uint8_t> *out)
{
    const int rows = std::min(L.rows_per_strip, L.height - strip * L.rows_per_strip);
    const size_t want = (size_t)rows * L.row_bytes;
    out->resize(want);
    raw->resize(L.strip_len[strip]);
    if (!read_at(fm.fd, raw->data(), L.strip_off[strip], raw->size())) return false;
    bool ok = false;
    if (L.compression == 32773) ok = unpack_bits(raw->data(), raw->size(), out->data(), want);
    else if (L.compression == 5) ok = unpack_lzw(raw->data(), raw->size(), out->data(), want);
    else {
        uLongf got = (uLongf)want;
        ok = uncompress(out->data(), &got, raw->data(), (uLong)raw->size()) == Z_OK && got == want;
    }
    if (!ok) return false;
    if (L.predictor == 2) {
        for (int r = 0; r < rows; r++) {
            uint8_t *p = out->data() + (size_t)r * L.row_bytes;
            if (L.bits == 8) {
                for (int x = 1; x < L.width; x++) p[x] = (uint8_t)(p[x] + p[x - 1]);
            } else {   // 16-bit samples in the file's byte order
                const int hi = L.big_endian ? 0 : 1, lo = 1 - hi;
                uint32_t acc = ((uint32_t)p[hi] << 8) | p[lo];
                for (int x = 1; x < L.width; x++) {
                    acc = (acc + (((uint32_t)p[2 * x + hi] << 8) | p[2 * x + lo])) & 0xffffu;
                    p[2 * x + hi] = (uint8_t)(acc >> 8);
                    p[2 * x + lo] = (uint8_t)acc;
                }
            }
        }
    }
    return true;
}

}  // namespace

int tiff_copy_rects(const FileMap &fm, const TiffLayout &L, const TiffRect *rects, int n, std::string *err)
{
    for (int k = 0; k < n; k++) {
        const TiffRect &q = rects[k];
        if (q.x < 0 || q.y < 0 || q.w < 0 || q.h < 0 || q.x + q.w > L.width || q.y + q.h > L.height)
            return fail(err, LST_ERR_INVALID, "rectangle outside the TIFF image");
    }
    if (L.jpeg) {
        // the whole file is read (pread, not the mapping: page faults would serialise the reader threads) and the scan parsed
        // from its start down to the last MCU row a rectangle needs
        static thread_local std::vector<uint8_t> file;
        if (file.size() < fm.size) file.resize(fm.size);
        uint64_t got = 0;
        while (got < fm.size) {
            const ssize_t k = pread(fm.fd, file.data() + got, fm.size - got, (off_t)got);
            if (k <= 0) return fail(err, LST_ERR_INVALID, "short read of a JPEG file");
            got += (uint64_t)k;
        }
        return jpeg_copy_rects(file.data(), fm.size, rects, n, err);
    }
    if (L.png) {
        // one zlib stream over all IDAT chunks; rows are filtered against the row above, so the image is
        // inflated from the top down to the last row any rectangle needs
        int last = -1;
        for (int k = 0; k < n; k++)
            if (rects[k].w > 0 && rects[k].h > 0) last = std::max(last, rects[k].y + rects[k].h - 1);
        if (last < 0) return LST_OK;
        const size_t bpp = (size_t)L.bits / 8, rb = (size_t)L.row_bytes;
        std::vector<uint8_t> rows(2 * (rb + 1), 0);
        uint8_t *cur = rows.data(), *prv = rows.data() + rb + 1;
        z_stream z;
        memset(&z, 0, sizeof(z));
        if (inflateInit(&z) != Z_OK) return fail(err, LST_ERR_INVALID, "zlib initialisation failed");
        size_t chunk = 0;
        bool ok = true;
        for (int y = 0; y <= last && ok; y++) {
            z.next_out = cur;
            z.avail_out = (uInt)(rb + 1);
            while (z.avail_out > 0 && ok) {
                if (z.avail_in == 0) {
                    if (chunk >= L.strip_off.size()) {
                        ok = false;
                        break;
                    }
                    z.next_in = (Bytef *)(fm.data + L.strip_off[chunk]);
                    z.avail_in = (uInt)L.strip_len[chunk];
                    chunk++;
                }
                const int zr = inflate(&z, Z_NO_FLUSH);
                if (zr != Z_OK && zr != Z_STREAM_END) ok = false;
                if (zr == Z_STREAM_END && z.avail_out > 0) ok = false;
            }
            if (!ok) break;
            uint8_t *p = cur + 1;
            const uint8_t *up = prv + 1;
            switch (cur[0]) {
                case 0: break;
                case 1: for (size_t i = bpp; i < rb; i++) p[i] = (uint8_t)(p[i] + p[i - bpp]); break;
                case 2: for (size_t i = 0; i < rb; i++) p[i] = (uint8_t)(p[i] + up[i]); break;
                case 3:
                    for (size_t i = 0; i < rb; i++) p[i] = (uint8_t)(p[i] + (((i >= bpp ? p[i - bpp] : 0) + up[i]) >> 1));
                    break;
                case 4:
                    for (size_t i = 0; i < rb; i++) {
                        const int a = i >= bpp ? p[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
                        const int pa = abs(b - c), pb = abs(a - c), pc = abs(a + b - 2 * c);
                        p[i] = (uint8_t)(p[i] + (pa <= pb && pa <= pc ? a : (pb <= pc ? b : c)));
                    }
                    break;
                default: ok = false; break;
            }
            for (int k = 0; k < n && ok; k++) {
                const TiffRect &q = rects[k];
                if (q.w > 0 && y >= q.y && y < q.y + q.h) memcpy(q.dst + (size_t)(y - q.y) * q.w * bpp, p + (size_t)q.x * bpp, (size_t)q.w * bpp);
            }
            std::swap(cur, prv);
        }
        inflateEnd(&z);
        if (!ok) return fail(err, LST_ERR_INVALID, "corrupt PNG image data");
        return LST_OK;
    }
    if (L.compression == 1) {
        for (int k = 0; k < n; k++) {
            int rc = tiff_copy_rect(fm, L, rects[k].x, rects[k].y, rects[k].w, rects[k].h, rects[k].dst, err);
            if (rc != LST_OK) return rc;
        }
        return LST_OK;
    }
    // compressed: every strip that holds rows of a rectangle is decoded once
    const size_t bpp = (size_t)L.bits / 8;
    std::vector<uint8_t> raw, strip;
    for (int s = 0; s < (int)L.strip_off.size(); s++) {
        const int r0 = s * L.rows_per_strip, r1 = std::min(L.height, r0 + L.rows_per_strip);
        bool needed = false;
        for (int k = 0; k < n && !needed; k++) needed = rects[k].h > 0 && rects[k].w > 0 && rects[k].y < r1 && rects[k].y + rects[k].h > r0;
        if (!needed) continue;
        if (!decode_strip(fm, L, s, &raw, &strip)) return fail(err, LST_ERR_INVALID, "corrupt compressed TIFF strip");
        for (int k = 0; k < n; k++) {
            const TiffRect &q = rects[k];
            const size_t nb = (size_t)q.w * bpp;
            for (int y = std::max(q.y, r0); y < std::min(q.y + q.h, r1); y++)
                memcpy(q.dst + (size_t)(y - q.y) * nb, strip.data() + (size_t)(y - r0) * L.row_bytes + (size_t)q.x * bpp, nb);
        }
    }
    return LST_OK;
}

int tiff_copy_rect(const FileMap &fm, const TiffLayout &L, int x0, int y0, int w, int h, uint8_t *dst, std::string *err)
{
    if (x0 < 0 || y0 < 0 || w < 0 || h < 0 || x0 + w > L.width || y0 + h > L.height)
        return fail(err, LST_ERR_INVALID, "rectangle outside the TIFF image");
    if (L.compression != 1) {
        TiffRect q{x0, y0, w, h, dst};
        return tiff_copy_rects(fm, L, &q, 1, err);
    }
    const size_t bpp = (size_t)L.bits / 8, nb = (size_t)w * bpp;
    auto rd = [&](uint8_t *d, uint64_t off, size_t len) -> bool {
        while (len > 0) {
            ssize_t k = pread(fm.fd, d, len, (off_t)off);
            if (k <= 0) return false;
            d += k;
            off += (uint64_t)k;
            len -= (size_t)k;
        }
        return true;
    };
    if (w == L.width && L.strip_off.size() == 1) {   // whole rows of a single strip: one read
        if (!rd(dst, L.offset_of_row(y0), nb * (size_t)h)) return fail(err, LST_ERR_INVALID, "short read from the slice file");
        return LST_OK;
    }
    for (int r = 0; r < h; r++)
        if (!rd(dst + (size_t)r * nb, L.offset_of_row(y0 + r) + (size_t)x0 * bpp, nb))
            return fail(err, LST_ERR_INVALID, "short read from the slice file");
    return LST_OK;
}

}  // namespace lst
