"""Depth-image input format (host only): the 16-bit PNGs the reference loads with stbi_load_16 (Application.cpp:28-29).

The PNG encoder / decoder below are an independent restatement of the PNG specification in numpy + zlib (Python's
stdlib), so the library's reader and writer are each checked against something that is not themselves."""
import ctypes as C
import struct
import zlib

import numpy as np
import pytest

from voxelhashing_demo_b200 import lib as L


def _chunk(t: bytes, d: bytes) -> bytes:
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def _filter_rows(rows: np.ndarray, bpp: int, ftypes) -> bytes:
    """rows: [H, rowbytes] uint8 -> filtered scanlines, filter type per row from ftypes (cycled)."""
    out = bytearray()
    prev = np.zeros(rows.shape[1], np.int32)
    for y, r in enumerate(rows.astype(np.int32)):
        ft = ftypes[y % len(ftypes)]
        out.append(ft)
        line = bytearray(len(r))
        for i in range(len(r)):
            a = r[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[ft]
            line[i] = (r[i] - pred) & 0xFF
        out += line
        prev = r
    return bytes(out)


def encode_png(samples: np.ndarray, bit_depth: int, ctype: int, ftypes=(0,), idat_split: int = 0, interlace: int = 0) -> bytes:
    """samples: [H, W, channels] of uint8 / uint16."""
    h, w, ch = samples.shape
    raw = samples.astype(">u2").tobytes() if bit_depth == 16 else samples.astype(np.uint8).tobytes()
    rows = np.frombuffer(raw, np.uint8).reshape(h, w * ch * bit_depth // 8)
    z = zlib.compress(_filter_rows(rows, ch * bit_depth // 8, ftypes), 6)
    parts = [z] if not idat_split else [z[i:i + idat_split] for i in range(0, len(z), idat_split)]
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bit_depth, ctype, 0, 0, interlace))
            + _chunk(b"tEXt", b"Comment\x00synthetic") + b"".join(_chunk(b"IDAT", p) for p in parts) + _chunk(b"IEND", b""))


def decode_png16_grey(data: bytes) -> np.ndarray:
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w = 8, b"", None
    while pos < len(data):
        n, t = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(t + body) & 0xFFFFFFFF
        if t == b"IHDR":
            w, h, bd, ct, _, _, il = struct.unpack(">IIBBBBB", body)
            assert (bd, ct, il) == (16, 0, 0)
        elif t == b"IDAT":
            idat += body
        pos += 12 + n
    raw = zlib.decompress(idat)
    row = 2 * w
    out = np.zeros((h, row), np.int32)
    prev = np.zeros(row, np.int32)
    for y in range(h):
        ft = raw[y * (row + 1)]
        s = raw[y * (row + 1) + 1:(y + 1) * (row + 1)]
        cur = np.zeros(row, np.int32)
        for i in range(row):
            a = cur[i - 2] if i >= 2 else 0
            b = prev[i]
            c = prev[i - 2] if i >= 2 else 0
            cur[i] = (s[i] + (0, a, b, (a + b) >> 1, _paeth(a, b, c))[ft]) & 0xFF
        out[y] = cur
        prev = cur
    return (out[:, 0::2] << 8 | out[:, 1::2]).astype(np.uint16)


def lib_read(lib, path):
    p = C.POINTER(C.c_uint16)()
    w, h = C.c_int(), C.c_int()
    rc = lib.vh_depth_read(str(path).encode(), C.byref(p), C.byref(w), C.byref(h))
    if rc != L.VH_OK:
        return rc, lib.vh_depth_last_error().decode()
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value)).copy()
    lib.vh_depth_free(p)
    return L.VH_OK, a


@pytest.fixture(scope="module")
def lib(built_library):
    return L.load_library()


def _image(h=37, w=53, seed=5):
    rng = np.random.default_rng(seed)
    img = (5000 * (1.5 + np.add.outer(np.linspace(0, 1, h), np.linspace(0, 0.7, w)))).astype(np.uint16)
    img[rng.random((h, w)) < 0.1] = 0                      # holes
    img[3:6, 2:w - 2] = rng.integers(0, 65536, (3, w - 4))  # high-entropy patch: every byte value, both bytes
    return img


@pytest.mark.parametrize("ftypes", [(0,), (1,), (2,), (3,), (4,), (4, 1, 3, 2, 0)])
def test_reads_16bit_grey_png_with_every_filter(lib, tmp_path, ftypes):
    img = _image()
    f = tmp_path / "d.png"
    f.write_bytes(encode_png(img[:, :, None], 16, 0, ftypes, idat_split=97))
    rc, got = lib_read(lib, f)
    assert rc == L.VH_OK and got.dtype == np.uint16 and np.array_equal(got, img)


def test_other_sample_formats_follow_stb_conventions(lib, tmp_path):
    img = _image(19, 23)
    f = tmp_path / "x.png"
    f.write_bytes(encode_png((img >> 8).astype(np.uint8)[:, :, None], 8, 0, (1, 4)))             # 8-bit grey -> v * 257
    assert np.array_equal(lib_read(lib, f)[1], (img >> 8) * 257)
    rgb = np.stack([img, img // 2, img // 3], -1)
    f.write_bytes(encode_png(rgb, 16, 2, (3, 4)))                                                # 16-bit RGB -> first channel
    assert np.array_equal(lib_read(lib, f)[1], img)
    f.write_bytes(encode_png(np.stack([img, 65535 - img], -1), 16, 4, (2,)))                     # grey + alpha
    assert np.array_equal(lib_read(lib, f)[1], img)
    f.write_bytes(encode_png(np.concatenate([rgb >> 8, np.full(img.shape + (1,), 255)], -1).astype(np.uint8), 8, 6, (4,)))   # RGBA8
    assert np.array_equal(lib_read(lib, f)[1], (img >> 8) * 257)


def test_pgm(lib, tmp_path):
    img = _image(11, 17)
    f = tmp_path / "d.pgm"
    f.write_bytes(b"P5\n# depth, 5000 per metre\n17 11\n65535\n" + img.astype(">u2").tobytes())
    assert np.array_equal(lib_read(lib, f)[1], img)
    f.write_bytes(b"P5 17 11 255\n" + (img >> 8).astype(np.uint8).tobytes())
    assert np.array_equal(lib_read(lib, f)[1], img >> 8)


@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4])
def test_writer_roundtrip_and_independent_decode(lib, tmp_path, filt):
    img = _image(29, 31, seed=filt)
    f = tmp_path / "w.png"
    assert lib.vh_depth_write_png(str(f).encode(), img.ctypes.data, 31, 29, filt) == L.VH_OK
    assert np.array_equal(decode_png16_grey(f.read_bytes()), img)        # the spec restatement can read it
    assert np.array_equal(lib_read(lib, f)[1], img)


def test_rejects_damaged_and_unsupported_files(lib, tmp_path):
    good = encode_png(_image(9, 9)[:, :, None], 16, 0)
    f = tmp_path / "bad.png"
    cases = {
        "not a PNG": b"JFIF" + good[4:],
        "CRC": good[:40] + bytes([good[40] ^ 1]) + good[41:],
        "truncated": good[:-20],
        "interlaced": encode_png(_image(9, 9)[:, :, None], 16, 0, interlace=1),
        "palette": good[:25].replace(b"\x10\x00\x00\x00\x00", b"\x08\x03\x00\x00\x00"),
    }
    for name, blob in cases.items():
        f.write_bytes(blob)
        rc, msg = lib_read(lib, f)
        assert rc == L.VH_ERR_INVALID and msg, name
    assert lib_read(lib, tmp_path / "missing.png")[0] == L.VH_ERR_INVALID
    assert lib.vh_depth_write_png(str(f).encode(), None, 4, 4, 0) == L.VH_ERR_INVALID
