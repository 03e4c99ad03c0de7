// lst_tiff.h -- the slices of a time-lapse as files: the reference's virtual stack
// hands analyseSlice one decoded file per slice (LST4:808-817, 831) and opens each file twice.  Here a
// slice file is mapped once and only the rows of the five search regions are touched.
//
// Supported: classic TIFF (II or MM), one sample per pixel, 8/16-bit unsigned or 32-bit float, strips
// of any height, uncompressed or compressed with PackBits, LZW or Deflate (with or without the
// horizontal predictor, as libtiff / OpenCV write them): only the strips that cover the requested rows
// are decoded; page p of a multi-page file by walking the IFD chain, or -- for stacks written by
// ImageJ (one IFD, "images=N" in the ImageDescription, slices back to back) -- by offset.  Anything
// else (JPEG-in-TIFF, tiles, RGB, BigTIFF, WhiteIsZero) is LST_ERR_UNSUPPORTED: the caller decodes such
// files itself and uses lst_track_ptrs.  JPEG files (baseline, grey or YCbCr) are decoded by lst_jpeg.cpp.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace lst {

struct TiffLayout {
    int width = 0, height = 0;
    int bits = 0;            // 8, 16, 32
    int sample_format = 1;   // 1 unsigned integer, 3 IEEE float
    bool big_endian = false;
    int rows_per_strip = 0;
    std::vector<uint64_t> strip_off;   // file offset of every strip
    std::vector<uint64_t> strip_len;   // stored (compressed) bytes of every strip
    int compression = 1;     // 1 none, 5 LZW, 8 / 32946 Deflate, 32773 PackBits
    int predictor = 1;       // 2: horizontal differencing of the samples
    uint64_t row_bytes = 0;
    uint64_t file_size = 0;
    int imagej_images = 0;   // "images=N" of an ImageJ ImageDescription (0 = none)
    bool png = false;        // a PNG file (grey 8/16-bit, not interlaced): strip_off/strip_len list its IDAT chunks
    bool jpeg = false;       // a JPEG file (lst_jpeg.h): strip 0 is the entropy-coded scan
    bool rgb = false;        // the pixels are ColorProcessor pixels (int32 0x00RRGGBB, bits = 32): 3-component JPEG
    uint64_t offset_of_row(int y) const { return strip_off[(size_t)(y / rows_per_strip)] + (uint64_t)(y % rows_per_strip) * row_bytes; }
};

// Parses page `page` of a TIFF image that lies in memory (a mapping of the file); a PNG signature is
// recognised and handed to the PNG parser (grey-scale 8/16-bit, not interlaced; samples are big-endian
// by definition, so they take the same device-side swap as "MM" TIFFs).  Returns LST_OK,
// LST_ERR_UNSUPPORTED or LST_ERR_INVALID with a message.
int tiff_parse(const uint8_t *data, uint64_t size, int page, TiffLayout *out, std::string *err);

// A read-only mapping of a file (used for the directory, which may lie anywhere in the file) plus the
// descriptor: pixel rows are fetched with pread, which -- unlike page faults on a mapping -- does not
// serialise the reader threads on the process's address-space lock.
struct FileMap {
    const uint8_t *data = nullptr;
    uint64_t size = 0;
    int fd = -1;
    int open(const char *path, std::string *err);
    void close();
    ~FileMap() { close(); }
};

// Copies the rectangle [x0, x0+w) x [y0, y0+h) (pixels) of the image into dst (row pitch w*bytes per pixel).
int tiff_copy_rect(const FileMap &fm, const TiffLayout &L, int x0, int y0, int w, int h, uint8_t *dst, std::string *err);

// Several rectangles of one image at once (every strip of a compressed file is decoded at most once).
struct TiffRect {
    int x, y, w, h;
    uint8_t *dst;
};
int tiff_copy_rects(const FileMap &fm, const TiffLayout &L, const TiffRect *rects, int n, std::string *err);

}  // namespace lst
