"""C0: readers of the reference's two Caffe files (weights.caffemodel read by caffe.Net in
evaluation.py:17-22, mean.binaryproto read through BlobProto in evaluation.py:25-31).  No real file
is available offline, so files in both on-disk layouts (LayerParameter / legacy V1LayerParameter,
BlobShape / legacy num-channels-height-width) are synthesised and read back."""
import warnings

import numpy as np
import pytest

from vanishing_points_2017_b200 import caffe_proto, cnn


def test_binaryproto_round_trip_both_shape_layouts(tmp_path):
    rs = np.random.RandomState(0)
    mean = rs.uniform(0, 255, (1, 1, 500, 500)).astype(np.float32)
    for legacy in (True, False):
        path = str(tmp_path / ("mean_%d.binaryproto" % legacy))
        caffe_proto.write_binaryproto(path, mean, legacy_shape=legacy)
        got = cnn.read_mean_blob(path)
        assert got.shape == (1, 1, 500, 500) and got.dtype == np.float32
        np.testing.assert_array_equal(got, mean)


def test_known_answer_bytes():
    # BlobProto{num:1 channels:1 height:1 width:2 data:[1.0, -2.5]} hand-encoded:
    # 08 01 | 10 01 | 18 01 | 20 02 | 2a 08 <0000803f 000020c0>
    raw = bytes.fromhex("08011001180120022a080000803f000020c0")
    np.testing.assert_array_equal(caffe_proto.parse_blob(raw), np.array([1.0, -2.5], np.float32).reshape(1, 1, 1, 2))
    # unpacked repeated float (wire type 5) and a multi-byte varint dimension (300 = ac 02)
    raw = bytes.fromhex("2d0000803f2d00000040")
    np.testing.assert_array_equal(caffe_proto.parse_blob(raw), np.array([1.0, 2.0], np.float32))
    big = caffe_proto.encode_blob(np.zeros((300, 2), np.float32))
    assert bytes.fromhex("ac02") in big[:8]
    assert caffe_proto.parse_blob(big).shape == (300, 2)
    with pytest.raises(ValueError):
        caffe_proto.parse_blob(bytes.fromhex("08021001180120022a080000803f000020c0"))     # shape says 4 values


@pytest.mark.parametrize("v1", [False, True])
def test_caffemodel_round_trip(tmp_path, v1):
    rs = np.random.RandomState(1)
    shapes = dict(zip(cnn.LAYER_NAMES, cnn.LAYER_SHAPES))
    shapes["fc6"] = (64, 57600)              # keep the synthetic file small; restored below
    layers, want = [("data", "Input", [])], {}
    for name in cnn.LAYER_NAMES:
        w = rs.standard_normal(shapes[name]).astype(np.float32)
        b = rs.standard_normal(shapes[name][0]).astype(np.float32)
        want[name] = (w, b)
        layers.append((name, "Convolution" if name.startswith("conv") else "InnerProduct", [w, b]))
        layers.append(("relu_" + name, "ReLU", []))
    path = str(tmp_path / "weights.caffemodel")
    caffe_proto.write_caffemodel(path, layers, v1=v1)
    got = caffe_proto.read_caffemodel(path)
    assert sorted(got) == sorted(cnn.LAYER_NAMES)            # layers without blobs are skipped
    for name in cnn.LAYER_NAMES:
        np.testing.assert_array_equal(got[name][0].reshape(shapes[name]), want[name][0])
        np.testing.assert_array_equal(got[name][1].reshape(-1), want[name][1])
    # through the loader init_caffe uses (shape check of the real architecture: fc6 is wrong on purpose)
    with pytest.raises(ValueError):
        cnn.load_weights(path)
    with pytest.raises(ValueError):
        caffe_proto.write_caffemodel(path, layers[:3], v1=v1)
        cnn.load_weights(path)


def test_random_fillers_warn_unless_asked_for():
    with pytest.warns(RuntimeWarning):
        cnn.load_weights(None)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        ws, bs = cnn.load_weights(None, allow_random=True)
    assert [w.shape for w in ws] == cnn.LAYER_SHAPES
