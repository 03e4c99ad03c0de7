"""SURVEY 8f-N2, JPEG slice files (host side, no GPU): the decoder of lst_jpeg.cpp restates the IJG library's default
reconstruction (integer IDCT jidctint.c, fancy upsampling jdsample.c, YCbCr tables jdcolor.c) -- what the JDK decoder
behind ImageJ's opener runs.  libjpeg-turbo, which cv2.imread uses here, is bit-compatible with it for those settings,
so every pixel must EQUAL cv2.imread's."""
import os

import numpy as np
import pytest

from laser_spot_track_b200 import LstError, tiff_probe, tiff_read

cv2 = pytest.importorskip("cv2")

SAMPLING = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
            "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}


def _scene(rng, h, w):
    base = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 3)
    return cv2.add(base, rng.integers(0, 40, (h, w, 3), dtype=np.uint8))


def _packed(bgr):
    return (bgr[:, :, 2].astype(np.uint32) << 16) | (bgr[:, :, 1].astype(np.uint32) << 8) | bgr[:, :, 0]


@pytest.mark.parametrize("shape", [(240, 320), (251, 333), (17, 23), (96, 8)])
@pytest.mark.parametrize("sampling", ["444", "422", "420"])
def test_colour_jpeg_equals_libjpeg(built_lib, tmp_path, shape, sampling):
    """Every quality / restart-interval / Huffman-table variant, whole image and rectangles (odd sizes exercise the
    partial MCUs and the edge rules of the upsampling filters)."""
    rng = np.random.default_rng(shape[0] * 7 + int(sampling))
    img = _scene(rng, *shape)
    h, w = shape
    for q in (35, 90, 100):
        for rst in (0, 5):
            for opt in (0, 1):
                p = str(tmp_path / f"c_{q}_{rst}_{opt}.jpg")
                assert cv2.imwrite(p, img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[sampling],
                                            cv2.IMWRITE_JPEG_RST_INTERVAL, rst, cv2.IMWRITE_JPEG_OPTIMIZE, opt])
                want = _packed(cv2.imread(p, cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION))
                info = tiff_probe(p)
                assert (info.width, info.height, info.bits, info.compression) == (w, h, 24, 7)
                got = tiff_read(p)
                assert got.dtype == np.uint32 and np.array_equal(got, want), (q, rst, opt)
                for (x, y, ww, hh) in [(0, 0, min(w, 9), min(h, 9)), (w // 3, h // 4, w // 2, h // 3), (w - 5, h - 6, 5, 6)]:
                    assert np.array_equal(tiff_read(p, rect=(x, y, ww, hh)), want[y:y + hh, x:x + ww]), (q, rst, opt, x, y)


def test_grey_jpeg_is_an_8_bit_slice(built_lib, tmp_path):
    rng = np.random.default_rng(3)
    g = cv2.cvtColor(_scene(rng, 203, 301), cv2.COLOR_BGR2GRAY)
    p = str(tmp_path / "g.jpg")
    assert cv2.imwrite(p, g, [cv2.IMWRITE_JPEG_QUALITY, 85])
    want = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert want.ndim == 2
    info = tiff_probe(p)
    assert (info.width, info.height, info.bits) == (301, 203, 8)
    assert np.array_equal(tiff_read(p), want)
    assert np.array_equal(tiff_read(p, rect=(100, 50, 77, 31)), want[50:81, 100:177])


def test_jpeg_full_hd_rectangles(built_lib, tmp_path):
    """Config-B sized slice: the five search regions of a 1920x1080 4:2:0 camera frame."""
    rng = np.random.default_rng(11)
    img = _scene(rng, 1080, 1920)
    p = str(tmp_path / "hd.jpg")
    assert cv2.imwrite(p, img, [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING["420"]])
    want = _packed(cv2.imread(p, cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION))
    for (x, y) in [(392, 182), (1352, 182), (1352, 722), (392, 722), (872, 452)]:
        assert np.array_equal(tiff_read(p, rect=(x, y, 176, 176)), want[y:y + 176, x:x + 176])


def test_unsupported_jpeg_variants_are_refused(built_lib, tmp_path):
    rng = np.random.default_rng(5)
    img = _scene(rng, 64, 64)
    p = str(tmp_path / "prog.jpg")
    assert cv2.imwrite(p, img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(LstError) as e:
        tiff_probe(p)
    assert "progressive" in str(e.value)
    q = str(tmp_path / "s411.jpg")
    assert cv2.imwrite(q, img, [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411])
    with pytest.raises(LstError):
        tiff_probe(q)
    t = str(tmp_path / "trunc.jpg")
    data = open(str(tmp_path / "prog.jpg"), "rb").read()
    open(t, "wb").write(data[:40])
    with pytest.raises(LstError):
        tiff_probe(t)
    with pytest.raises(LstError):
        tiff_read(str(tmp_path / "prog.jpg"))
    assert not os.path.exists(str(tmp_path / "nothing.jpg"))
