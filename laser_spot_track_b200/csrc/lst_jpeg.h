// lst_jpeg.h -- JPEG slice files of a time-lapse (the virtual stack of LST4:808-817, 831 hands
// analyseSlice one decoded file per slice, and camera time-lapses are JPEG: the EXIF path of LST4:1199-1223).
//
// ImageJ opens a JPEG through the JDK's decoder, which is the IJG library (libjpeg 6b) with its default settings:
// the accurate integer IDCT (jidctint.c, JDCT_ISLOW), "fancy" triangle-filter chroma upsampling (jdsample.c) and the
// fixed-point YCbCr -> RGB tables of jdcolor.c.  All three are integer algorithms with one defined result, restated
// here; libjpeg-turbo (what cv2.imread uses in this image) is bit-compatible with 6b for these settings, so the
// decoder is pinned bit for bit against cv2.imread (tests/test_jpeg.py).  A 3-component file gives a ColorProcessor
// (one int32 0x00RRGGBB per pixel -> LST_PIX_RGB), a 1-component file an 8-bit slice (ImageJ converts grey JPEGs to
// 8 bits, Opener.convertGrayJpegTo8Bits).
//
// Supported: baseline / extended-sequential Huffman JPEG (SOF0, SOF1), 8-bit samples, one interleaved scan, 1 component
// or 3 components (YCbCr by JFIF / Adobe convention, or RGB with the Adobe transform flag 0) sampled 4:4:4, 4:2:2
// (h2v1) or 4:2:0 (h2v2), restart intervals.  Progressive, arithmetic-coded, lossless, 12-bit, CMYK and other
// sampling ratios: LST_ERR_UNSUPPORTED.
//
// The entropy-coded scan has to be parsed from its start (DC prediction, no random access), but only the blocks of
// the MCUs that cover the requested rectangles go through the IDCT, and parsing stops after the last MCU row needed.
#pragma once
#include <stdint.h>
#include <string>

#include "lst_jpeg_math.h"
#include "lst_tiff.h"

namespace lst {

inline bool jpeg_signature(const uint8_t *d, uint64_t n) { return d && n >= 4 && d[0] == 0xff && d[1] == 0xd8 && d[2] == 0xff; }

// Fills width / height / bits (8, or 32 with rgb = true) and marks the layout as JPEG.
int jpeg_parse(const uint8_t *data, uint64_t size, int page, TiffLayout *out, std::string *err);

// The rectangles of the decoded image: 8-bit samples, or little-endian int32 0x00RRGGBB pixels (row pitch w * 4).
int jpeg_copy_rects(const uint8_t *data, uint64_t size, const TiffRect *rects, int n, std::string *err);

// ---- the device path: entropy decoding here, reconstruction on the GPU (lst_jpeg_reconstruct) ---------------------------
// Entropy decoding is a bit-serial dependency chain through the whole scan and stays on host threads (one file per thread);
// what follows it -- dequantisation, inverse DCT, chroma upsampling, colour conversion -- is independent per block / pixel
// and runs on the device for the MCUs of the search regions: the reader threads stage quantised coefficients instead of pixels.
int jpeg_geometry(const uint8_t *data, uint64_t size, JpegGeom *g, std::string *err);
// The quantised coefficients (64 int16 per block, natural order) of the MCUs of region k go to coef[k]: MCU-major
// ((my - my0) * nmx + (mx - mx0)), within an MCU the luma blocks row-major, then Cb, Cr.  qt_out: 3 x 64 quantisation
// values (component order, natural order).  The file must have the geometry g (LST_ERR_INVALID otherwise).
int jpeg_decode_coeffs(const uint8_t *data, uint64_t size, const JpegGeom &g, const JpegRegion *regs, int16_t *const *coef, int n,
                       uint16_t *qt_out, std::string *err);

}  // namespace lst
