"""The ascii .t7 reader / writer (mccnn_b200/t7.py) against hand-written files in Torch7's published layout (File.lua
writeObject, THDiskFile auto-spacing; parity unpinned -- no Torch7 here to produce a file) and through round trips, and the
mapping of the reference's saved nets ({net_te, [net_te2,] opt}, main.lua:587-600 / 893-901) onto tower / head layers."""
import numpy as np
import pytest

from mccnn_b200 import t7

# torch.save(f, {torch.FloatTensor{{1,2,3},{4,5,6}}, 'ab'}, 'ascii') as File.lua lays it out
HAND = b"""3
1
2
1
1
4
2
3
V 1
17
torch.FloatTensor
2
2 3
3 1
1
4
3
3
V 1
18
torch.FloatStorage
6
1 2 3 4 5 6
1
2
2
2
ab
"""

# {w = <column 2 of a 2x3 CudaTensor>, same = <the same tensor object>, t = true, n = nil-valued keys are never written,
#  x = -0.5, e = <empty CudaTensor>}; legacy class header without the 'V ' string for the storage
HAND2 = b"""3
1
5
2
1
w
4
2
3
V 1
16
torch.CudaTensor
1
2
3
2
4
3
17
torch.CudaStorage
6
0.100000001 1.5 -2 3.25 1e-30 7
2
4
same
4
2
2
1
t
5
1
2
1
x
1
-0.5
2
1
e
4
4
3
V 1
16
torch.CudaTensor
0
1
0
"""


def test_hand_written_file_parses():
    o = t7.loads(HAND)
    assert list(o) == [1, 2] and o[2] == "ab"
    assert o[1].dtype == np.float32 and np.array_equal(o[1], [[1, 2, 3], [4, 5, 6]])
    assert o.array()[1] == "ab"


def test_writer_reproduces_the_hand_written_file():
    assert t7.dumps(t7.loads(HAND), cuda=False) == HAND


def test_strided_views_shared_references_and_legacy_header():
    o = t7.loads(HAND2)
    assert np.array_equal(o["w"], np.array([1.5, 1e-30], np.float32))             # offset 2, stride 3
    assert o["same"] is o["w"] and o["t"] is True and o["x"] == -0.5
    assert o["e"].size == 0 and o["e"].dtype == np.float32


def test_round_trip_keeps_values_and_ties():
    rng = np.random.default_rng(0)
    st = rng.standard_normal(24).astype(np.float32)
    a, b = st.reshape(4, 6), st.reshape(4, 6)[1:3, ::2]                           # two tensors over one storage
    d = rng.standard_normal(5)                                                    # DoubleTensor
    obj = {1: a, 2: b, "d": d, "i": np.arange(4, dtype=np.int64), "s": "x\ny z", "nested": [1.5, False, {"k": 3}],
           "big": float(2 ** 60), "nan": float("nan"), "inf": float("-inf")}
    back = t7.loads(t7.dumps(obj))
    assert np.array_equal(back[1], a) and np.array_equal(back[2], b) and np.shares_memory(back[1], back[2])
    assert back["d"].dtype == np.float64 and np.array_equal(back["d"], d)         # %.17g is exact
    assert back["i"].dtype == np.int64 and back["s"] == "x\ny z"
    n = back["nested"]
    assert n.array()[:2] == [1.5, False] and n[3]["k"] == 3
    assert back["big"] == 2.0 ** 60 and np.isnan(back["nan"]) and back["inf"] == float("-inf")
    back[1][0, 0] = 42.0                                                          # ties are real views
    assert t7.dumps(back) != t7.dumps(obj)


def test_fp32_text_is_exact():
    x = np.frombuffer(np.random.default_rng(1).bytes(4 * 4096), dtype=np.float32)
    x = x[np.isfinite(x)]
    back = t7.loads(t7.dumps({1: x}))[1]
    assert np.array_equal(back.view(np.uint32), x.view(np.uint32))                # %.9g round-trips every finite fp32


def test_cycles_terminate():
    a = t7.T7Table()
    a["self"] = a
    back = t7.loads(t7.dumps(a))
    assert back["self"] is back


def _layers(rng, l1, fm, n_in=1, nh2=None, l2=0):
    tower = [(rng.standard_normal((fm, n_in if i == 0 else fm, 3, 3)).astype(np.float32),
              rng.standard_normal(fm).astype(np.float32)) for i in range(l1)]
    head = None
    if nh2:
        dims = [2 * fm] + [nh2] * l2 + [1]
        head = [(rng.standard_normal((dims[i + 1], dims[i])).astype(np.float32),
                 rng.standard_normal(dims[i + 1]).astype(np.float32)) for i in range(l2 + 1)]
    return tower, head


@pytest.mark.parametrize("arch", ["fast", "slow"])
def test_net_file_round_trip(tmp_path, arch):
    rng = np.random.default_rng(7)
    tower, head = _layers(rng, 4, 16, nh2=32 if arch == "slow" else None, l2=3)
    f = str(tmp_path / "net.t7")
    t7.save_net(f, tower, head, {"fm": 16, "l1": 4, "a": "train_all", "at": 0})
    raw = t7.load(f)
    seq = raw[1]
    names = [m.typename for m in seq["modules"].array()]
    if arch == "fast":                                                            # main.lua:726-749
        assert names == ["cudnn.SpatialConvolution", "cudnn.ReLU"] * 3 + ["cudnn.SpatialConvolution", "nn.Normalize2",
                                                                          "nn.StereoJoin"]
    else:                                                                         # main.lua:682-695
        assert names == ["cudnn.SpatialConvolution", "cudnn.ReLU"] * 4
        assert [m.typename for m in raw[2]["modules"].array()] == ["nn.SpatialConvolution1_fw", "cudnn.ReLU"] * 3 + [
            "nn.SpatialConvolution1_fw", "cudnn.Sigmoid"]
        assert raw[2]["modules"][1]["bias"].shape == (1, 32, 1, 1)                # SpatialConvolution1_fw.lua:8
    net = t7.load_net(f)
    assert net.arch == arch and net.opt["a"] == "train_all" and net.opt["fm"] == 16
    assert len(net.tower) == 4 and all(np.array_equal(w, w0) and np.array_equal(b, b0)
                                       for (w, b), (w0, b0) in zip(net.tower, tower))
    if arch == "slow":
        assert len(net.head) == 4 and all(np.array_equal(w, w0) and np.array_equal(b, b0) and b.ndim == 1
                                          for (w, b), (w0, b0) in zip(net.head, head))
    else:
        assert net.head is None


def test_flat_conv_weight_is_reshaped():
    """older cudnn.torch keeps the weight as (fm, cin*kH*kW): the module's plane / kernel fields give the shape"""
    rng = np.random.default_rng(2)
    tower, _ = _layers(rng, 2, 16)
    obj = t7.make_net_object(tower)
    for m in obj[1]["modules"].array():
        if m.typename == "cudnn.SpatialConvolution":
            m.fields["weight"] = m["weight"].reshape(m["weight"].shape[0], -1)
    net = t7.net_from_object(t7.loads(t7.dumps(obj)))
    assert all(np.array_equal(w, w0) for (w, _), (w0, _) in zip(net.tower, tower))


@pytest.mark.parametrize("bad, msg", [
    (b"", "end of file"),
    (b"9\n", "unknown type tag"),
    (b"6\n1\n", "Lua function"),
    (b"2\n5\nab\n", "past the end"),
    (HAND[:-5], "end of file|past the end"),
    (HAND + b"1\n", "trailing"),
    (HAND.replace(b"1 2 3 4 5 6", b"1 2 3 4 5"), "values expected|malformed|integer|number"),
    (HAND.replace(b"\n2 3\n3 1\n1\n", b"\n2 3\n3 1\n2\n"), "does not fit"),
])
def test_malformed_files_fail_loudly(bad, msg):
    with pytest.raises(t7.T7Error, match=msg):
        t7.loads(bad)


def test_binary_file_is_refused(tmp_path):
    f = tmp_path / "bin.t7"
    f.write_bytes(b"\x03\x00\x00\x00\x01\x00\x00\x00")
    with pytest.raises(t7.T7Error, match="ASCII"):
        t7.load(str(f))


def test_wrong_net_shapes_are_refused():
    with pytest.raises(t7.T7Error, match="net_te"):
        t7.net_from_object(t7.T7Table({1: t7.T7Table()}))
    obj = t7.make_net_object(_layers(np.random.default_rng(0), 2, 16)[0])
    obj[1]["modules"][1].fields["bias"] = np.zeros(3, np.float32)
    with pytest.raises(t7.T7Error, match="bias"):
        t7.net_from_object(obj)
