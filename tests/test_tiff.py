"""the slice-file reader behind lst_track_files.  CPU part: the TIFF parser (host only)."""
import numpy as np
import pytest

from laser_spot_track_b200 import LstError, tiff_probe
from laser_spot_track_b200 import native as N

import tiffio


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("big_endian", [False, True])
@pytest.mark.parametrize("rps", [None, 7])
def test_probe_layout(built_lib, tmp_path, dtype, big_endian, rps):
    img = (np.arange(40 * 52).reshape(40, 52) % 251).astype(dtype)
    p = tmp_path / "a.tif"
    tiffio.write_tiff(p, [img], big_endian=big_endian, rows_per_strip=rps)
    info = tiff_probe(str(p))
    assert (info.width, info.height, info.bits) == (52, 40, np.dtype(dtype).itemsize * 8)
    assert info.sample_format == (3 if dtype == np.float32 else 1)
    assert bool(info.big_endian) == big_endian
    assert info.n_strips == (1 if rps is None else 6) and info.rows_per_strip == (40 if rps is None else 7)
    raw = open(p, "rb").read()
    first_row = np.frombuffer(raw, dtype=np.dtype(dtype).newbyteorder(">" if big_endian else "<"), count=52,
                              offset=info.first_strip_offset)
    assert np.array_equal(first_row, img[0])


def test_probe_pages_and_imagej_stack(built_lib, tmp_path):
    pages = [np.full((20, 30), 10 * i + 1, np.uint16) for i in range(4)]
    p = tmp_path / "multi.tif"
    tiffio.write_tiff(p, pages)
    offs = [tiff_probe(str(p), i).first_strip_offset for i in range(4)]
    assert len(set(offs)) == 4 and offs == sorted(offs)
    with pytest.raises(LstError):
        tiff_probe(str(p), 4)
    q = tmp_path / "stack.tif"
    tiffio.write_tiff(q, pages, big_endian=True, imagej_stack=True)
    i0, i3 = tiff_probe(str(q), 0), tiff_probe(str(q), 3)
    assert i0.imagej_images == 4 and i3.first_strip_offset - i0.first_strip_offset == 3 * 20 * 30 * 2
    with pytest.raises(LstError):
        tiff_probe(str(q), 4)


def test_probe_rejects_what_the_path_cannot_read(built_lib, tmp_path):
    img = np.zeros((16, 16), np.uint8)
    for kw in (dict(compression=7, fake_compression=True), dict(photometric=0)):      # 7: JPEG-in-TIFF
        p = tmp_path / "x.tif"
        tiffio.write_tiff(p, [img], **kw)
        with pytest.raises(LstError) as e:
            tiff_probe(str(p))
        assert e.value.code == N.LST_ERR_UNSUPPORTED
    (tmp_path / "junk.tif").write_bytes(b"not a tiff at all")
    with pytest.raises(LstError) as e:
        tiff_probe(str(tmp_path / "junk.tif"))
    assert e.value.code == N.LST_ERR_INVALID
    with pytest.raises(LstError):
        tiff_probe(str(tmp_path / "missing.tif"))
    # truncated pixel data
    p = tmp_path / "t.tif"
    tiffio.write_tiff(p, [np.zeros((64, 64), np.uint16)])
    raw = open(p, "rb").read()
    info = tiff_probe(str(p))
    cut = bytearray(raw)
    # move the strip offset beyond the end of the file
    import struct
    idx = raw.find(struct.pack("<HHII", 273, 4, 1, info.first_strip_offset))
    struct.pack_into("<I", cut, idx + 8, len(raw) - 100)
    p.write_bytes(bytes(cut))
    with pytest.raises(LstError):
        tiff_probe(str(p))


def _test_images(dtype):
    rng = np.random.default_rng(11)
    h, w = 97, 131
    yy, xx = np.mgrid[0:h, 0:w]
    top = 255 if dtype == np.uint8 else 65535
    smooth = ((np.sin(xx / 9.0) + np.cos(yy / 7.0) + 2.0) / 4.0 * top).astype(dtype)           # compresses well
    noisy = rng.integers(0, top + 1, (h, w)).astype(dtype)                                       # does not
    flat = np.full((h, w), 7, dtype)                                                             # long runs
    return [smooth, noisy, flat]


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("scheme", ["lzw", "deflate", "packbits"])
def test_compressed_strips_written_by_libtiff(built_lib, tmp_path, dtype, scheme):
    """Slice files as libtiff (through cv2.imwrite) writes them: LZW / Deflate with the horizontal predictor,
    PackBits, several strips.  Whole images and rectangles that straddle strips decode to the pixels."""
    import cv2
    from laser_spot_track_b200 import tiff_read
    code = {"lzw": 5, "deflate": 8, "packbits": 32773}[scheme]
    for k, img in enumerate(_test_images(dtype)):
        p = str(tmp_path / f"{scheme}{k}.tif")
        assert cv2.imwrite(p, img, [cv2.IMWRITE_TIFF_COMPRESSION, code])
        info = tiff_probe(p)
        assert info.compression == code and (info.width, info.height) == (img.shape[1], img.shape[0])
        assert np.array_equal(tiff_read(p), img)
        for rect in [(0, 0, 14, 14), (17, 40, 90, 33), (100, 60, 31, 37), (0, 96, 131, 1)]:
            x, y, w, h = rect
            assert np.array_equal(tiff_read(p, rect=rect), img[y:y + h, x:x + w]), (scheme, k, rect)


@pytest.mark.parametrize("big_endian", [False, True])
@pytest.mark.parametrize("compression,predictor", [(8, 1), (8, 2), (32773, 1), (32773, 2)])
def test_compressed_strips_either_byte_order(built_lib, tmp_path, big_endian, compression, predictor):
    from laser_spot_track_b200 import tiff_read
    for dtype in (np.uint8, np.uint16):
        for k, img in enumerate(_test_images(dtype)):
            p = str(tmp_path / f"c{k}.tif")
            tiffio.write_tiff(p, [img], big_endian=big_endian, rows_per_strip=13, compression=compression, predictor=predictor)
            assert np.array_equal(tiff_read(p), img)
            assert np.array_equal(tiff_read(p, rect=(5, 11, 100, 30)), img[11:41, 5:105])


def test_corrupt_compressed_strip_is_an_error(built_lib, tmp_path):
    from laser_spot_track_b200 import tiff_read
    img = _test_images(np.uint16)[0]
    p = tmp_path / "bad.tif"
    tiffio.write_tiff(p, [img], compression=5, fake_compression=True)      # says LZW, holds raw bytes
    tiff_probe(str(p))
    with pytest.raises(LstError):
        tiff_read(str(p))


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_grey_png_slices(built_lib, tmp_path, dtype):
    """Grey 8/16-bit PNG slices (cv2.imwrite -> libpng: every row filter type occurs): whole image and
    rectangles; RGB / interlaced PNG is rejected."""
    import cv2
    from laser_spot_track_b200 import tiff_read
    for k, img in enumerate(_test_images(dtype)):
        for level in (1, 9):
            p = str(tmp_path / f"g{k}_{level}.png")
            assert cv2.imwrite(p, img, [cv2.IMWRITE_PNG_COMPRESSION, level])
            info = tiff_probe(p)
            assert (info.width, info.height, info.bits) == (img.shape[1], img.shape[0], img.dtype.itemsize * 8)
            assert np.array_equal(tiff_read(p), img)
            assert np.array_equal(tiff_read(p, rect=(17, 40, 90, 33)), img[40:73, 17:107])
            assert np.array_equal(tiff_read(p, rect=(0, 0, 14, 14)), img[:14, :14])
    rgb = str(tmp_path / "rgb.png")
    assert cv2.imwrite(rgb, np.zeros((20, 20, 3), np.uint8))
    with pytest.raises(LstError) as e:
        tiff_probe(rgb)
    assert e.value.code == N.LST_ERR_UNSUPPORTED
    with pytest.raises(LstError):
        tiff_probe(p, 1)
