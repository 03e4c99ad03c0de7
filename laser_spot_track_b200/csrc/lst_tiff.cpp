// lst_tiff.cpp -- minimal reader for the uncompressed grey-scale TIFF files of a time-lapse (see lst_tiff.h).
#include "lst_tiff.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/lst_b200.h"

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

int tiff_parse(const uint8_t *data, uint64_t size, int page, TiffLayout *out, std::string *err)
{
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
    uint64_t strips_entry = 0, desc_off = 0;
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
    if (compression != 1) return fail(err, LST_ERR_UNSUPPORTED, "compressed TIFF is not supported (decode it and use lst_track_ptrs)");
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
    for (uint32_t i = 0; i < want_strips; i++) {
        const uint64_t rows = i + 1 < want_strips ? rps : (uint64_t)L.height - (uint64_t)i * rps;
        if (!r.ok(L.strip_off[i], rows * L.row_bytes)) return fail(err, LST_ERR_INVALID, "TIFF pixel data outside the file");
    }
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

int tiff_copy_rect(const FileMap &fm, const TiffLayout &L, int x0, int y0, int w, int h, uint8_t *dst, std::string *err)
{
    if (x0 < 0 || y0 < 0 || w < 0 || h < 0 || x0 + w > L.width || y0 + h > L.height)
        return fail(err, LST_ERR_INVALID, "rectangle outside the TIFF image");
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
