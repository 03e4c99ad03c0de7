"""Minimal TIFF writer for the tests: grey-scale, either byte order, any strip height, multi-page (IFD
chain) or ImageJ-style stack (one IFD, slices back to back); strips uncompressed, PackBits (32773) or
Deflate (8), optionally with the horizontal predictor (LZW files come from cv2.imwrite in the tests)."""
import struct
import zlib

import numpy as np


def _ifd(bo, entries, next_off):
    out = struct.pack(bo + "H", len(entries))
    for tag, typ, cnt, val in sorted(entries):
        out += struct.pack(bo + "HHI", tag, typ, cnt)
        if typ == 3 and cnt == 1:
            out += struct.pack(bo + "HH", val, 0)
        else:
            out += struct.pack(bo + "I", val)
    return out + struct.pack(bo + "I", next_off)


def _packbits(data: bytes) -> bytes:
    out, i, n = bytearray(), 0, len(data)
    while i < n:
        j = i
        while j + 1 < n and data[j + 1] == data[i] and j - i < 127:
            j += 1
        if j > i:                                   # run of j - i + 1 equal bytes
            out += bytes([(257 - (j - i + 1)) & 0xff, data[i]])
            i = j + 1
        else:                                       # literal run
            j = i
            while j < n and j - i < 128 and not (j + 1 < n and data[j + 1] == data[j]):
                j += 1
            j = max(j, i + 1)
            out += bytes([j - i - 1]) + data[i:j]
            i = j
    return bytes(out)


def write_tiff(path, pages, big_endian=False, rows_per_strip=None, imagej_stack=False, compression=1,
               photometric=1, predictor=1, fake_compression=False):
    """pages: list of 2-D arrays (uint8 / uint16 / float32), all of one shape and dtype."""
    bo = ">" if big_endian else "<"
    pages = [np.ascontiguousarray(p) for p in pages]
    h, w = pages[0].shape
    dt = pages[0].dtype
    bits = dt.itemsize * 8
    fmt = 3 if dt.kind == "f" else 1
    rps = h if rows_per_strip is None else rows_per_strip
    nstr = (h + rps - 1) // rps
    def encode(p):
        """list of stored strips of one page"""
        strips = []
        for i in range(nstr):
            rows = p[i * rps:min(h, (i + 1) * rps)]
            if predictor == 2:
                d = rows.astype(np.int64)
                d[:, 1:] = d[:, 1:] - d[:, :-1]
                rows = (d % (1 << bits)).astype(dt)
            b = rows.astype(dt.newbyteorder(bo)).tobytes()
            if not fake_compression:
                if compression == 32773:
                    b = _packbits(b)
                elif compression in (8, 32946):
                    b = zlib.compress(b)
            strips.append(b)
        return strips
    raw = [encode(p) for p in pages]
    buf = bytearray((b"MM" if big_endian else b"II") + struct.pack(bo + "HI", 42, 0))
    desc = (f"ImageJ=1.53t\nimages={len(pages)}\nslices={len(pages)}\n".encode() + b"\0") if imagej_stack else None
    ifd_pos_patch = 4
    n_ifds = 1 if imagej_stack else len(pages)
    for pi in range(n_ifds):
        # pixel data first (strips contiguous), then tables, then the IFD
        offs, cnts = [], []
        for pj, strips in enumerate(raw if imagej_stack else [raw[pi]]):
            for b in strips:
                if pj == 0:
                    offs.append(len(buf))
                    cnts.append(len(b))
                buf += b
        if len(buf) % 2:
            buf += b"\0"
        entries = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, bits), (259, 3, 1, compression),
                   (262, 3, 1, photometric), (277, 3, 1, 1), (278, 4, 1, rps), (339, 3, 1, fmt)]
        if predictor != 1:
            entries.append((317, 3, 1, predictor))
        if nstr == 1:
            entries += [(273, 4, 1, offs[0]), (279, 4, 1, cnts[0])]
        else:
            o1 = len(buf)
            buf += struct.pack(bo + f"{nstr}I", *offs)
            o2 = len(buf)
            buf += struct.pack(bo + f"{nstr}I", *cnts)
            entries += [(273, 4, nstr, o1), (279, 4, nstr, o2)]
        if desc is not None:
            od = len(buf)
            buf += desc
            if len(buf) % 2:
                buf += b"\0"
            entries.append((270, 2, len(desc), od))
        ifd_off = len(buf)
        struct.pack_into(bo + "I", buf, ifd_pos_patch, ifd_off)
        ifd = _ifd(bo, entries, 0)
        buf += ifd
        ifd_pos_patch = ifd_off + len(ifd) - 4
    with open(path, "wb") as f:
        f.write(bytes(buf))
