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
