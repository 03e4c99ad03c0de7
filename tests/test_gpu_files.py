"""SURVEY 8f-N2 on the GPU: slices read from TIFF files (lst_track_files) give the records the oracle
computes from the decoded arrays, for either byte order, strip layouts, multi-page files and the
fallback to whole-slice reads."""
import numpy as np
import pytest

from laser_spot_track_b200 import LaserSpotTrack4, LstError
from laser_spot_track_b200.synth import StackConfig, SyntheticStack
from oracle import lst_oracle as O

import tiffio
import util

pytestmark = pytest.mark.gpu


def _oracle(cfg, st, frames, **kw):
    op = util.oracle_params(cfg, **kw)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    return op, otr, otr.run(frames[1:])


def _tracker(cfg, op, **kw):
    bits = {"u8": 8, "u16": 16, "f32": 32}[cfg.dtype]
    return LaserSpotTrack4(cfg.width, cfg.height, bits, templSize=op.templ_size, sArea=op.s_area, markDist=op.mark_dist,
                           subPixel=op.sub_pixel, matchIntensity=op.match_intensity, **kw)


@pytest.mark.parametrize("dtype,big_endian,rps", [("u16", False, None), ("u16", True, 7), ("u8", False, 5),
                                                   ("f32", True, None), ("f32", False, 64)])
def test_track_files_equals_oracle(built_lib, oracle_clib, tmp_path, dtype, big_endian, rps):
    cfg = StackConfig(400, 300, dtype, 32, 24, 30)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 30)]
    op, otr, want = _oracle(cfg, st, frames, match_intensity=True)
    paths = []
    for i, f in enumerate(frames):
        p = tmp_path / f"s{i:04d}.tif"
        tiffio.write_tiff(p, [f], big_endian=big_endian, rows_per_strip=rps)
        paths.append(str(p))
    with _tracker(cfg, op, max_batch=8) as t:
        info = t.set_reference_file(paths[0], st.ref_xy)
        assert list(info.mideal) == otr.mideal
        got = t.track_files(paths[1:])
        util.compare_records(got, want, label=f"files {dtype} be={big_endian} rps={rps}")
        s = t.stats()
        assert s.roi_fallback_tasks == 0
        assert 0 < s.h2d_bytes < 0.8 * len(want) * frames[0].nbytes     # only the search regions travel


def test_track_pages_of_one_file(built_lib, oracle_clib, tmp_path):
    cfg = StackConfig(400, 300, "u16", 32, 24, 20)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 20)]
    op, otr, want = _oracle(cfg, st, frames, match_intensity=True)
    for name, kw in (("pages.tif", dict()), ("ijstack.tif", dict(imagej_stack=True, big_endian=True))):
        p = str(tmp_path / name)
        tiffio.write_tiff(p, frames, **kw)
        with _tracker(cfg, op, max_batch=6) as t:
            t.set_reference_file(p, st.ref_xy, page=0)
            got = t.track_files([p] * 19, pages=list(range(1, 20)))
            util.compare_records(got, want, label=name)


def test_track_files_windows_leave_the_regions(built_lib, oracle_clib, tmp_path):
    """A jump larger than the margin read around the search regions: the batch is re-read whole."""
    cfg = StackConfig(400, 300, "u16", 32, 24, 24)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 24)]
    for i in range(9, 24):
        frames[i] = np.roll(frames[i], (15, -17), axis=(0, 1))
    op, otr, want = _oracle(cfg, st, frames, match_intensity=True)
    assert sum(w.status == 0 for w in want) >= 20
    paths = []
    for i, f in enumerate(frames):
        p = tmp_path / f"j{i:04d}.tif"
        tiffio.write_tiff(p, [f], big_endian=True)
        paths.append(str(p))
    with _tracker(cfg, op, max_batch=5) as t:
        t.set_reference_file(paths[0], st.ref_xy)
        got = t.track_files(paths[1:])
        util.compare_records(got, want, label="files with a jump")
        assert t.stats().roi_fallback_tasks > 0


def test_track_files_whole_slice_search_and_errors(built_lib, oracle_clib, tmp_path):
    cfg = StackConfig(200, 160, "u8", 24, 0, 6)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 6)]
    op, otr, want = _oracle(cfg, st, frames, match_intensity=False)
    paths = []
    for i, f in enumerate(frames):
        p = tmp_path / f"w{i}.tif"
        tiffio.write_tiff(p, [f])
        paths.append(str(p))
    with _tracker(cfg, op) as t:
        t.set_reference_file(paths[0], st.ref_xy)
        util.compare_records(t.track_files(paths[1:]), want, label="files, sArea=0")
        # a slice of another size: "All stack images must have the same size and type" (LST4:393)
        bad = tmp_path / "bad.tif"
        tiffio.write_tiff(bad, [np.zeros((100, 100), np.uint8)])
        with pytest.raises(LstError):
            t.track_files([str(bad)])
        with pytest.raises(LstError):
            t.track_files([str(tmp_path / "missing.tif")])
        comp = tmp_path / "lzw.tif"
        tiffio.write_tiff(comp, [frames[1]], compression=5)
        with pytest.raises(LstError):
            t.track_files([str(comp)])


@pytest.mark.parametrize("scheme", ["lzw", "deflate", "packbits"])
def test_track_compressed_slice_files(built_lib, oracle_clib, tmp_path, scheme):
    """Compressed slice files as libtiff writes them (cv2.imwrite): the reader threads decode only the strips
    that cover the search regions; records equal the oracle's on the decoded arrays."""
    import cv2
    cfg = StackConfig(400, 300, "u16", 32, 24, 16)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 16)]
    op, otr, want = _oracle(cfg, st, frames, match_intensity=True)
    code = {"lzw": 5, "deflate": 8, "packbits": 32773}[scheme]
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"z{i:03d}.tif")
        assert cv2.imwrite(p, f, [cv2.IMWRITE_TIFF_COMPRESSION, code])
        paths.append(p)
    with _tracker(cfg, op, max_batch=6) as t:
        t.set_reference_file(paths[0], st.ref_xy)
        util.compare_records(t.track_files(paths[1:]), want, label=f"{scheme} files")


def test_track_png_slice_files(built_lib, oracle_clib, tmp_path):
    """16-bit grey PNG slices (network byte order: swapped on the device like big-endian TIFFs)."""
    import cv2
    cfg = StackConfig(400, 300, "u16", 32, 24, 12)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 12)]
    op, otr, want = _oracle(cfg, st, frames, match_intensity=True)
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"p{i:03d}.png")
        assert cv2.imwrite(p, f)
        paths.append(p)
    with _tracker(cfg, op, max_batch=5) as t:
        t.set_reference_file(paths[0], st.ref_xy)
        util.compare_records(t.track_files(paths[1:]), want, label="png files")


@pytest.mark.parametrize("device_stage", ["1", "0"])
@pytest.mark.parametrize("sampling,intensity", [("420", True), ("444", True), ("422", False)])
def test_track_jpeg_slices(built_lib, oracle_clib, tmp_path, monkeypatch, sampling, intensity, device_stage):
    """a time-lapse of JPEG files (the virtual stack of LST4:808-831).  A 3-component JPEG is a ColorProcessor
    slice; the oracle runs on the frames cv2.imread decodes (libjpeg-turbo, bit-compatible with the JDK's IJG decoder), the
    library reads the files itself -- only the MCUs under the search regions are reconstructed: on the device from the
    quantised coefficients the reader threads stage (default), or by the reader threads (LST_JPEG_DEVICE=0).  matchIntensity
    true (convertToGray32) and false (3-channel match)."""
    import cv2
    monkeypatch.setenv("LST_JPEG_DEVICE", device_stage)
    samp = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
            "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422}[sampling]
    cfg = StackConfig(400, 300, "u8", 32, 24, 16)
    st = SyntheticStack(cfg)
    paths, frames = [], []
    for i in range(16):
        g = st.frame(i).astype(np.uint8)
        bgr = np.stack([255 - g // 2, g, (g.astype(np.uint16) * 3 // 4 + 17).astype(np.uint8)], axis=-1)
        p = str(tmp_path / f"t{i:04d}.jpg")
        assert cv2.imwrite(p, bgr, [cv2.IMWRITE_JPEG_QUALITY, 93, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samp,
                                    cv2.IMWRITE_JPEG_RST_INTERVAL, 4 if i % 2 else 0])
        dec = cv2.imread(p, cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)
        frames.append((dec[:, :, 2].astype(np.uint32) << 16) | (dec[:, :, 1].astype(np.uint32) << 8) | dec[:, :, 0])
        paths.append(p)
    op, otr, want = _oracle(cfg, st, frames, match_intensity=intensity)
    assert sum(w.status == 0 for w in want) >= 12
    with LaserSpotTrack4(cfg.width, cfg.height, 24, templSize=cfg.templ_size, sArea=cfg.s_area, matchIntensity=intensity,
                         max_batch=6) as t:
        info = t.set_reference_file(paths[0], st.ref_xy)
        assert list(info.mideal) == otr.mideal
        got = t.track_files(paths[1:])
        util.compare_records(got, want, label=f"JPEG {sampling} intensity={intensity}")
        assert t.stats().roi_fallback_tasks == 0


@pytest.mark.parametrize("device_stage", ["1", "0"])
def test_track_grey_jpeg_slices(built_lib, oracle_clib, tmp_path, monkeypatch, device_stage):
    """1-component JPEG files are 8-bit slices (ImageJ converts grey JPEGs to 8 bits)."""
    import cv2
    monkeypatch.setenv("LST_JPEG_DEVICE", device_stage)
    cfg = StackConfig(400, 300, "u8", 32, 24, 12)
    st = SyntheticStack(cfg)
    paths, frames = [], []
    for i in range(12):
        p = str(tmp_path / f"g{i:04d}.jpg")
        assert cv2.imwrite(p, st.frame(i).astype(np.uint8), [cv2.IMWRITE_JPEG_QUALITY, 95])
        frames.append(cv2.imread(p, cv2.IMREAD_UNCHANGED))
        paths.append(p)
    op, otr, want = _oracle(cfg, st, frames, match_intensity=True)
    with _tracker(cfg, op, max_batch=5) as t:
        t.set_reference_file(paths[0], st.ref_xy)
        util.compare_records(t.track_files(paths[1:]), want, label="grey JPEG")
    with LaserSpotTrack4(cfg.width, cfg.height, 24, templSize=cfg.templ_size, sArea=cfg.s_area) as t:
        with pytest.raises(LstError):
            t.set_reference_file(paths[0], st.ref_xy)       # an 8-bit file in an RGB context


@pytest.mark.parametrize("sampling", ["420", "422", "444"])
def test_jpeg_regions_at_the_image_border(built_lib, oracle_clib, tmp_path, sampling):
    """Search windows that reach the border of a 250 x 187 image (partial MCUs on the right and at the bottom): the device
    reconstruction has to apply the upsampling filters' edge rules at the image's own first / last chroma sample, not at the
    border of the staged MCUs.  Any wrong pixel changes a blurred window and with it a score: records bit-exact."""
    import cv2
    samp = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
            "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422}[sampling]
    cfg = StackConfig(250, 187, "u8", 24, 60, 8)
    st = SyntheticStack(cfg)
    paths, frames = [], []
    for i in range(8):
        g = st.frame(i).astype(np.uint8)
        bgr = np.stack([255 - g // 2, g, (g.astype(np.uint16) * 3 // 4 + 17).astype(np.uint8)], axis=-1)
        p = str(tmp_path / f"b{i:04d}.jpg")
        assert cv2.imwrite(p, bgr, [cv2.IMWRITE_JPEG_QUALITY, 93, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samp, cv2.IMWRITE_JPEG_RST_INTERVAL, 3])
        dec = cv2.imread(p, cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)
        frames.append((dec[:, :, 2].astype(np.uint32) << 16) | (dec[:, :, 1].astype(np.uint32) << 8) | dec[:, :, 0])
        paths.append(p)
    op, otr, want = _oracle(cfg, st, frames, match_intensity=True)
    assert any(t.wy == 0 for t in want[0].tm) and any(t.wy + t.wh == cfg.height for t in want[0].tm)
    with LaserSpotTrack4(cfg.width, cfg.height, 24, templSize=cfg.templ_size, sArea=cfg.s_area, max_batch=4) as t:
        t.set_reference_file(paths[0], st.ref_xy)
        util.compare_records(t.track_files(paths[1:]), want, label=f"JPEG border {sampling}")


def test_jpeg_windows_leave_the_regions(built_lib, oracle_clib, tmp_path):
    """JPEG slices with a jump larger than the margin staged around the search regions: the coefficient staging of that batch
    does not hold the MCUs the windows need, so the batch is decoded whole (host reconstruction) and tracked again."""
    import cv2
    cfg = StackConfig(400, 300, "u8", 32, 24, 20)
    st = SyntheticStack(cfg)
    paths, frames = [], []
    for i in range(20):
        g = st.frame(i).astype(np.uint8)
        if i >= 9:
            g = np.roll(g, (15, -17), axis=(0, 1))
        bgr = np.stack([255 - g // 2, g, (g.astype(np.uint16) * 3 // 4 + 17).astype(np.uint8)], axis=-1)
        p = str(tmp_path / f"w{i:04d}.jpg")
        assert cv2.imwrite(p, bgr, [cv2.IMWRITE_JPEG_QUALITY, 93, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420])
        dec = cv2.imread(p, cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)
        frames.append((dec[:, :, 2].astype(np.uint32) << 16) | (dec[:, :, 1].astype(np.uint32) << 8) | dec[:, :, 0])
        paths.append(p)
    op, otr, want = _oracle(cfg, st, frames, match_intensity=True)
    assert sum(w.status == 0 for w in want) >= 15
    with LaserSpotTrack4(cfg.width, cfg.height, 24, templSize=cfg.templ_size, sArea=cfg.s_area, max_batch=5) as t:
        t.set_reference_file(paths[0], st.ref_xy)
        got = t.track_files(paths[1:])
        util.compare_records(got, want, label="JPEG files with a jump")
        assert t.stats().roi_fallback_tasks > 0
