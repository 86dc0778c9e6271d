"""CPU tests of the sequence driver's image input (PNG via zlib, PGM), against files written by python cv2 -- the
stand-in for cv::imread / loadPGM of the reference driver (stereo_vision.cu:661-662, image.h:134-161)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def test_png_flavours_roundtrip(svb, tmp_path, kitti_gray):
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (77, 131, 3), dtype=np.uint8)
    rgb[:, :40] = rgb[:, :1]  # smooth areas make the encoder pick different scanline filters
    rgb[30:50] = np.arange(131, dtype=np.uint8)[None, :, None]
    for level in (0, 1, 9):
        p = tmp_path / ("rgb%d.png" % level)
        cv2.imwrite(str(p), rgb, [cv2.IMWRITE_PNG_COMPRESSION, level])
        got = svb.image_read(p)
        assert got.shape == (77, 131, 4) and np.array_equal(got[..., :3], rgb) and (got[..., 3] == 255).all()
    rgba = rng.integers(0, 256, (33, 57, 4), dtype=np.uint8)
    cv2.imwrite(str(tmp_path / "rgba.png"), rgba)
    assert np.array_equal(svb.image_read(tmp_path / "rgba.png"), rgba)
    g16 = rng.integers(0, 65536, (21, 45)).astype(np.uint16)
    cv2.imwrite(str(tmp_path / "g16.png"), g16)
    assert np.array_equal(svb.image_read(tmp_path / "g16.png")[..., 0], (g16 >> 8).astype(np.uint8))
    # a full KITTI-size gray frame
    L = kitti_gray["L0"]
    cv2.imwrite(str(tmp_path / "L.png"), L)
    assert np.array_equal(svb.image_read(tmp_path / "L.png")[..., 0], L)
    # colour frame as the KITTI sequences store it, compared with what sv.py feeds the library (BGR -> BGRA)
    col = np.stack([L, np.roll(L, 3, 1), np.roll(L, 5, 0)], -1)
    cv2.imwrite(str(tmp_path / "col.png"), col)
    want = cv2.cvtColor(cv2.imread(str(tmp_path / "col.png")), cv2.COLOR_BGR2BGRA)
    assert np.array_equal(svb.image_read(tmp_path / "col.png"), want)


def test_pgm_with_comment_and_errors(svb, tmp_path):
    rng = np.random.default_rng(1)
    g = rng.integers(0, 256, (19, 23), dtype=np.uint8)
    (tmp_path / "c.pgm").write_bytes(b"P5\n# CREATOR: GIMP PNM Filter Version 1.1\n23 19\n255\n" + g.tobytes())
    assert np.array_equal(svb.image_read(tmp_path / "c.pgm")[..., 0], g)
    (tmp_path / "plain.pgm").write_bytes(b"P5 23 19 255\n" + g.tobytes())
    assert np.array_equal(svb.image_read(tmp_path / "plain.pgm")[..., 0], g)
    (tmp_path / "short.pgm").write_bytes(b"P5\n23 19\n255\n" + g.tobytes()[:100])
    with pytest.raises(svb.SvbError):
        svb.image_read(tmp_path / "short.pgm")
    (tmp_path / "bad.png").write_bytes(b"not a png at all, really not")
    with pytest.raises(svb.SvbError):
        svb.image_read(tmp_path / "bad.png")
    with pytest.raises(svb.SvbError):
        svb.image_read(tmp_path / "missing.png")


def test_corrupt_files_fail_cleanly(svb, tmp_path):
    """Bit flips, truncations and hostile headers: svb_image_read returns an error (or a decoded image) and never crashes, never throws
    across the C boundary, never allocates what the payload cannot fill (run under ASan / UBSan as well when the parser changes)."""
    import struct
    import zlib

    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (24, 31, 3), dtype=np.uint8)
    cv2.imwrite(str(tmp_path / "ok.png"), img)
    good = (tmp_path / "ok.png").read_bytes()
    target = tmp_path / "fuzz.png"
    outcomes = {"ok": 0, "error": 0}

    def attempt(data, name=target):
        name.write_bytes(data)
        try:
            out = svb.image_read(name)
            assert out.ndim == 3 and out.shape[2] in (1, 4)
            outcomes["ok"] += 1
        except svb.SvbError:
            outcomes["error"] += 1

    for k in range(400):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(8, len(b)))] ^= 1 << int(rng.integers(0, 8))
        attempt(bytes(b))
    for cut in range(0, len(good), 7):
        attempt(good[:cut])

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)

    sig = good[:8]
    tiny = zlib.compress(b"\x00" * 40)
    for w, h, depth, ctype in ((16384, 16384, 16, 6), (0xFFFFFFFF, 0xFFFFFFFF, 8, 2), (0, 5, 8, 0), (5, 0, 8, 0), (7, 3, 8, 3), (4, 4, 1, 0), (3, 3, 8, 5)):
        ihdr = struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)
        attempt(sig + chunk(b"IHDR", ihdr) + chunk(b"IDAT", tiny) + chunk(b"IEND", b""))
    # palette image whose indices point past a short palette, bad filter byte, chunk length past the end of the file
    pal = zlib.compress(b"".join(b"\x00" + bytes([200] * 7) for _ in range(3)))
    attempt(sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 7, 3, 8, 3, 0, 0, 0)) + chunk(b"PLTE", b"\x01\x02\x03") + chunk(b"IDAT", pal) + chunk(b"IEND", b""))
    badf = zlib.compress(b"".join(b"\x09" + bytes(7) for _ in range(3)))
    attempt(sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 7, 3, 8, 0, 0, 0, 0)) + chunk(b"IDAT", badf) + chunk(b"IEND", b""))
    attempt(sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 7, 3, 8, 0, 0, 0, 0)) + struct.pack(">I", 0x7FFFFFFF) + b"IDAT" + b"xx")
    assert outcomes["error"] > 50 and outcomes["ok"] > 0  # flips in ancillary bytes still decode

    pgm = tmp_path / "fuzz.pgm"
    for header in (b"P5\n2147483647 2147483647\n255\n", b"P5\n-3 4\n255\n", b"P5\n3 4\n65535\n", b"P5\n3\n", b"P5", b"", b"P5\n# only a comment", b"P2\n3 4\n255\n"):
        pgm.write_bytes(header + bytes(12))
        with pytest.raises(svb.SvbError):
            svb.image_read(pgm)
