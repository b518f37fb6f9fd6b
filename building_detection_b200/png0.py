"""Level-0 ("stored") PNG codec for 8-bit grey masks -- the hand-off format between the reference's stages:
predict.py:115 and model_fuse.py:350 write their masks with cv.imwrite(..., [IMWRITE_PNG_COMPRESSION, 0]) and
buildAPI.py:122-123 base64-encodes the result file.  A stored PNG is the raw rows behind a fixed skeleton
(signature, IHDR, one zlib stream of stored deflate blocks with a filter byte per row, IEND), so encoding is a
strided copy plus two checksums (zlib.crc32 / zlib.adler32 run at memory speed in C) and decoding is a strided view:
no deflate on either side.  The chunking differs from libpng's (one IDAT instead of 8 KB pieces); every PNG reader,
cv.imread included, decodes both to the same pixels (tests/test_host_io.py).  Other PNGs fall back to cv.imdecode."""
from __future__ import annotations

import struct
import zlib

import numpy as np

_SIG = b"\x89PNG\r\n\x1a\n"
_BLOCK = 65535  # largest stored deflate block


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(data, zlib.crc32(tag)) & 0xFFFFFFFF)


def encode_gray(mask, threads=0):
    """(H,W) uint8 -> PNG bytes (colour type 0, bit depth 8, compression level 0, filter 0 on every row).  Uses the
    multi-threaded native encoder (csrc/png0.cpp, host code of libbd_b200.so: 3 GB/s on 8 cores against 0.1 GB/s for
    cv.imwrite at level 0); ``encode_gray_py`` is the same byte stream from numpy + zlib checksums."""
    m = np.ascontiguousarray(mask, np.uint8)
    if m.ndim != 2:
        raise ValueError(f"expected an (H,W) uint8 mask, got {m.shape}")
    import ctypes as C

    from . import runtime as R
    L = R.lib()
    h, w = m.shape
    n = L.bd_png0_size(h, w)
    out = np.empty(n, np.uint8)
    ln = C.c_size_t()
    if L.bd_png0_encode(m.ctypes.data_as(C.c_void_p), h, w, out.ctypes.data_as(C.c_void_p), n, C.byref(ln), int(threads)) != 0:
        raise ValueError("bd_png0_encode failed")
    return out[:ln.value].tobytes()


def encode_gray_py(mask):
    """encode_gray without the native library (numpy copies + zlib.crc32 / zlib.adler32): identical bytes."""
    m = np.ascontiguousarray(mask, np.uint8)
    if m.ndim != 2:
        raise ValueError(f"expected an (H,W) uint8 mask, got {m.shape}")
    h, w = m.shape
    raw = np.empty((h, w + 1), np.uint8)
    raw[:, 0] = 0  # filter type None
    raw[:, 1:] = m
    raw = raw.reshape(-1)
    n = raw.size
    nblk = max(1, -(-n // _BLOCK))
    # stored blocks: 1 header byte (BFINAL on the last) + LEN + ~LEN, then the bytes
    body = np.empty(2 + n + 5 * nblk + 4, np.uint8)
    body[0], body[1] = 0x78, 0x01
    full = n // _BLOCK
    pos = 2
    if full:
        blocks = body[pos:pos + full * (_BLOCK + 5)].reshape(full, _BLOCK + 5)
        blocks[:, 0] = 0
        blocks[:, 1:5] = np.frombuffer(struct.pack("<HH", _BLOCK, 0), np.uint8)
        blocks[:, 5:] = raw[:full * _BLOCK].reshape(full, _BLOCK)
        pos += full * (_BLOCK + 5)
    rest = n - full * _BLOCK
    if rest or not full:
        body[pos] = 1
        body[pos + 1:pos + 5] = np.frombuffer(struct.pack("<HH", rest, rest ^ 0xFFFF), np.uint8)
        body[pos + 5:pos + 5 + rest] = raw[full * _BLOCK:]
        pos += 5 + rest
    else:
        body[pos - (_BLOCK + 5)] = 1  # the last full block is final
    body[pos:pos + 4] = np.frombuffer(struct.pack(">I", zlib.adler32(raw) & 0xFFFFFFFF), np.uint8)
    pos += 4
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)
    return _SIG + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", body[:pos].tobytes()) + _chunk(b"IEND", b"")


def decode_gray(data):
    """PNG bytes -> (H,W) uint8.  Fast path for 8-bit grey, non-interlaced, stored-only deflate with filter 0 rows
    (what encode_gray and cv.imwrite at compression 0 produce); anything else is handed to cv.imdecode."""
    buf = memoryview(data)
    if bytes(buf[:8]) != _SIG:
        raise ValueError("not a PNG file")
    pos, idat, ihdr = 8, [], None
    while pos + 8 <= len(buf):
        (ln,), tag = struct.unpack(">I", buf[pos:pos + 4]), bytes(buf[pos + 4:pos + 8])
        if tag == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", buf[pos + 8:pos + 8 + 13])
        elif tag == b"IDAT":
            idat.append(buf[pos + 8:pos + 8 + ln])
        elif tag == b"IEND":
            break
        pos += 12 + ln
    if ihdr is None:
        raise ValueError("PNG without IHDR")
    w, h, depth, ctype, _comp, _filt, interlace = ihdr
    out = None
    if depth == 8 and ctype == 0 and interlace == 0 and idat:
        z = np.frombuffer(b"".join(idat) if len(idat) > 1 else idat[0], np.uint8)
        out = _inflate_stored(z, h * (w + 1))
        if out is not None:
            rows = out.reshape(h, w + 1)
            if rows[:, 0].any():
                out = None  # filtered rows: general decoder
            else:
                return np.ascontiguousarray(rows[:, 1:])
    import cv2 as cv
    img = cv.imdecode(np.frombuffer(bytes(buf), np.uint8), cv.IMREAD_UNCHANGED)
    if img is None:
        raise ValueError("undecodable PNG")
    return img


def _inflate_stored(z, expect):
    """zlib stream made of stored blocks only -> bytes, or None when a compressed block shows up."""
    if z.size < 6 or (int(z[0]) & 0x0F) != 8:
        return None
    out = np.empty(expect, np.uint8)
    pos, o = 2, 0
    while True:
        if pos + 5 > z.size:
            return None
        hdr = int(z[pos])
        if hdr & 0x06:  # BTYPE != 00
            return None
        ln = int(z[pos + 1]) | (int(z[pos + 2]) << 8)
        if (ln ^ 0xFFFF) != (int(z[pos + 3]) | (int(z[pos + 4]) << 8)) or o + ln > expect or pos + 5 + ln > z.size:
            return None
        out[o:o + ln] = z[pos + 5:pos + 5 + ln]
        o += ln
        pos += 5 + ln
        if hdr & 1:
            break
    if o != expect or pos + 4 > z.size:
        return None
    if struct.unpack(">I", z[pos:pos + 4].tobytes())[0] != (zlib.adler32(out) & 0xFFFFFFFF):
        raise ValueError("PNG data checksum mismatch")
    return out
