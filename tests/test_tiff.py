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
    for kw in (dict(compression=5), dict(photometric=0)):
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
