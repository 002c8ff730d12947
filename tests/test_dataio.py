"""CPU: the data formats either side of the hot path (mc-cnn_b200/dataio.py)."""
import os

import numpy as np
import pytest

import mccnn_b200  # noqa: F401
from mccnn_b200 import dataio

pytest.importorskip("PIL.Image")


@pytest.mark.parametrize("dtype", ["float32", "int32", "int64"])
def test_bin_dim_type_roundtrip(tmp_path, dtype):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((3, 1, 5, 7)) * 100).astype(dtype)
    f = str(tmp_path / "x.bin")
    dataio.tofile(f, x)
    assert open(f + ".type").read() == dtype
    assert open(f + ".dim").read().split() == ["3", "1", "5", "7"]           # one dim per line (main.lua:355-358)
    y = dataio.fromfile(f)
    assert y.dtype == np.dtype(dtype) and y.shape == x.shape and np.array_equal(x, y)


def test_empty_and_errors(tmp_path):
    f = str(tmp_path / "e.bin")
    dataio.tofile(f, None)
    assert dataio.fromfile(f).size == 0                                       # .dim == "0" (main.lua:359-361)
    g = str(tmp_path / "g.bin")
    dataio.tofile(g, np.zeros((2, 3), np.float32))
    open(g + ".dim", "w").write("2\n4")
    with pytest.raises(ValueError):
        dataio.fromfile(g)
    open(g + ".dim", "w").write("2\n3")
    open(g + ".type", "w").write("float64")
    with pytest.raises(ValueError):
        dataio.fromfile(g)
    with pytest.raises(ValueError):
        dataio.tofile(g, np.zeros(3, np.float64))


def test_png16_semantics(tmp_path):
    """adcensus.cu:1670-1706: value = disparity * 256 truncated, < 1e-5 -> 0 (invalid), read back / 256"""
    d = np.array([[0.0, 5e-6, 1.0, 1.5], [2.00390625, 100.999, 227.99609375, 255.99]], np.float32)
    f = str(tmp_path / "d.png")
    dataio.write_png16(d, f)
    from PIL import Image

    raw = np.asarray(Image.open(f)).astype(np.int64)
    want = np.array([[0, 0, 256, 384], [513, int(np.float32(100.999) * np.float32(256)), 58367, int(np.float32(255.99) * np.float32(256))]])
    assert np.array_equal(raw, want)
    back = dataio.read_png16(f)
    assert back.dtype == np.float32 and np.array_equal(back, want.astype(np.float32) / 256)
    assert np.abs(back - d)[d >= 1e-5].max() < 1 / 256 + 1e-6
    dataio.write_png16(np.full((2, 2), 300.0, np.float32), f)                 # past uint16: clamped, not wrapped
    assert np.asarray(Image.open(f)).max() == 65535
    Image.fromarray(np.zeros((2, 2), np.uint8)).save(f)
    with pytest.raises(ValueError):
        dataio.read_png16(f)


def test_pfm_roundtrip_and_header(tmp_path):
    rng = np.random.default_rng(1)
    d = rng.random((4, 6)).astype(np.float32) * 50
    f = str(tmp_path / "d.pfm")
    dataio.write_pfm(np.flipud(d), f)                                         # main.lua:1215 flips before writing
    head = open(f, "rb").read(24)
    assert head.startswith(b"Pf\n6 4\n-0.003922\n")
    assert np.array_equal(np.flipud(dataio.read_pfm(f)), d)
    open(f, "wb").write(b"PF\n1 1\n-1\n")
    with pytest.raises(ValueError):
        dataio.read_pfm(f)


def test_bad_pixel_rate():
    actual = np.array([[0, 10, 20, 30], [40, 0, 50, 60]], np.float32)        # 0 = no ground truth
    pred = np.array([[99, 10, 23.5, 27], [40, 99, 46.9, 60]], np.float32)
    # errors on the 6 valid pixels: 0, 3.5, 3, 0, 3.1, 0
    assert dataio.bad_pixel_rate(pred, actual, 3.0) == pytest.approx(2 / 6)  # strictly greater (main.lua:1234)
    assert dataio.bad_pixel_rate(pred, actual, 1.0) == pytest.approx(3 / 6)
    assert np.isnan(dataio.bad_pixel_rate(pred, np.zeros_like(actual), 3.0))
    assert dataio.ERR_AT["kitti"] == 3.0 and dataio.ERR_AT["mb"] == 1.0     # main.lua:400, 453


def test_kitti_submission_image_and_file(tmp_path):
    pred = np.arange(12, dtype=np.float32).reshape(3, 4) + 1
    img = dataio.kitti_submission_image(pred, 5, 4)
    assert img.shape == (5, 4) and np.all(img[:2] == 0) and np.array_equal(img[2:], pred)   # bottom rows (main.lua:1204)
    f = dataio.write_kitti_submission(pred, 5, 4, str(tmp_path / "out"), 7, "kitti2015")
    assert f.endswith("disp_0/000007_10.png")
    assert np.array_equal(dataio.read_png16(f), img)


# ---- pinned against the reference's own Python (fixtures: oracle/make_dataio_golden.py ran /root/reference/preprocess_mb.py's code)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataio")


def test_fromfile_reads_what_the_reference_tofile_wrote():
    """tests/golden/dataio/*.bin(+.dim,.type) were written by preprocess_mb.py:99-106 itself"""
    want = np.load(os.path.join(GOLD, "expected.npz"))
    for name in ("x_f32", "te_i32", "meta_i32"):
        got = dataio.fromfile(os.path.join(GOLD, name + ".bin"))
        assert got.dtype == want[name].dtype and got.shape == want[name].shape
        assert np.array_equal(got, want[name]), name
    assert dataio.fromfile(os.path.join(GOLD, "none.bin")).size == 0          # '.dim' holding the single line 0


def test_tofile_writes_the_same_bytes_as_the_reference(tmp_path):
    want = np.load(os.path.join(GOLD, "expected.npz"))
    for name in ("x_f32", "te_i32", "meta_i32"):
        f = str(tmp_path / (name + ".bin"))
        dataio.tofile(f, want[name])
        for ext in ("", ".dim", ".type"):
            assert open(f + ext, "rb").read() == open(os.path.join(GOLD, name + ".bin" + ext), "rb").read(), name + ext
    f = str(tmp_path / "none.bin")
    dataio.tofile(f, None)
    assert open(f + ".dim", "rb").read() == open(os.path.join(GOLD, "none.bin.dim"), "rb").read()


def test_read_pfm_agrees_with_the_reference_reader():
    """disp.pfm as parsed by preprocess_mb.py:13-57 (load_pfm returns the rows top-down: np.flipud of the file order)"""
    ref = np.load(os.path.join(GOLD, "disp_pfm_as_read_by_reference.npy"))
    got = dataio.read_pfm(os.path.join(GOLD, "disp.pfm"))
    assert np.array_equal(np.flipud(got), ref)


def test_bin_layout_is_what_samples_load_bin_reads(tmp_path):
    """samples/load_bin.py: np.memmap(name, float32, shape=(1, D, H, W)) -- plain C-order float32, no header"""
    D, H, W = 3, 4, 5
    vol = np.arange(D * H * W, dtype=np.float32).reshape(1, D, H, W)
    f = str(tmp_path / "left.bin")
    vol.tofile(f)
    back = np.memmap(f, dtype=np.float32, shape=(1, D, H, W))
    assert np.array_equal(back, vol)
