"""Readers for the two Caffe files the reference's CNN glue opens: the trained
`weights.caffemodel` (caffe.Net(model_def, model_weights, caffe.TEST), reference
evaluation.py:17-22) and `mean.binaryproto` (caffe.proto.caffe_pb2.BlobProto +
caffe.io.blobproto_to_array, evaluation.py:25-31).

Caffe and protobuf's generated classes are not needed: both files are plain
protocol-buffer wire format and only a handful of fields matter.  Field numbers
follow BVLC/caffe `src/caffe/proto/caffe.proto` (tag rc5, the version the
reference pins in README.md:5; the source is not vendored in the reference):

  BlobProto        num=1 channels=2 height=3 width=4 (legacy 4-D shape), data=5 (packed float),
                   diff=6, shape=7 (BlobShape), double_data=8 (packed double)
  BlobShape        dim=1 (packed int64)
  NetParameter     name=1, layers=2 (V1LayerParameter, pre-2015 files), layer=100 (LayerParameter)
  LayerParameter   name=1 type=2 blobs=7
  V1LayerParameter name=4 blobs=6

The writers exist for the round-trip tests (no real caffemodel is available
offline) and to let users export weights for the original code.
"""
import struct

import numpy as np

_VARINT, _I64, _LEN, _I32 = 0, 1, 2, 5


def _varint(buf, pos):
    out = shift = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _fields(buf):
    """Yield (field number, wire type, value) of one message; LEN values are memoryviews."""
    buf = memoryview(buf)
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == _VARINT:
            val, pos = _varint(buf, pos)
        elif wt == _I64:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == _I32:
            val, pos = buf[pos:pos + 4], pos + 4
        elif wt == _LEN:
            n, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + n], pos + n
        else:
            raise ValueError("unsupported wire type %d (groups are not used by caffe.proto)" % wt)
        if pos > end:
            raise ValueError("truncated field %d" % num)
        yield num, wt, val


def _packed_varints(view):
    out, pos = [], 0
    while pos < len(view):
        v, pos = _varint(view, pos)
        out.append(v)
    return out


def parse_blob(buf):
    """BlobProto -> float32 ndarray in Caffe blob layout (caffe.io.blobproto_to_array: the shape
    message if present, else the legacy (num, channels, height, width))."""
    legacy = {}
    shape = None
    chunks, single = [], []
    dchunks, dsingle = [], []
    for num, wt, val in _fields(buf):
        if num in (1, 2, 3, 4) and wt == _VARINT:
            legacy[num] = val
        elif num == 5:
            if wt == _LEN:
                chunks.append(np.frombuffer(val, dtype="<f4"))
            else:                                   # unpacked repeated float
                single.append(struct.unpack("<f", val)[0])
        elif num == 8:
            if wt == _LEN:
                dchunks.append(np.frombuffer(val, dtype="<f8"))
            else:
                dsingle.append(struct.unpack("<d", val)[0])
        elif num == 7 and wt == _LEN:
            dims = []
            for n2, w2, v2 in _fields(val):
                if n2 == 1:
                    dims += _packed_varints(v2) if w2 == _LEN else [v2]
            shape = tuple(int(d) for d in dims)
    if single:
        chunks.append(np.asarray(single, dtype=np.float32))
    if dsingle:
        dchunks.append(np.asarray(dsingle, dtype=np.float64))
    if chunks:
        data = np.concatenate(chunks).astype(np.float32)
    elif dchunks:
        data = np.concatenate(dchunks).astype(np.float32)
    else:
        data = np.zeros(0, dtype=np.float32)
    if shape is None:
        shape = tuple(int(legacy.get(k, 1)) for k in (1, 2, 3, 4)) if legacy else (data.size,)
    if int(np.prod(shape)) != data.size:
        raise ValueError("BlobProto: shape %r does not match %d values" % (shape, data.size))
    return data.reshape(shape)


def read_binaryproto(path):
    """mean.binaryproto -> ndarray (reference evaluation.py:25-31 returns it as (1,1,H,W))."""
    with open(path, "rb") as fh:
        return parse_blob(fh.read())


def parse_caffemodel(buf):
    """NetParameter -> {layer name: [blob ndarrays]} for every layer that carries blobs."""
    layers = {}
    for num, wt, val in _fields(buf):
        if wt != _LEN or num not in (2, 100):
            continue
        name_field, blob_field = (1, 7) if num == 100 else (4, 6)
        name, blobs = None, []
        for n2, w2, v2 in _fields(val):
            if n2 == name_field and w2 == _LEN:
                name = bytes(v2).decode("utf-8", "replace")
            elif n2 == blob_field and w2 == _LEN:
                blobs.append(parse_blob(v2))
        if name is not None and blobs:
            layers[name] = blobs
    return layers


def read_caffemodel(path):
    with open(path, "rb") as fh:
        return parse_caffemodel(fh.read())


# ---- writers (tests, export) ---------------------------------------------------
def _enc_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_len(num, payload):
    return _enc_varint((num << 3) | _LEN) + _enc_varint(len(payload)) + bytes(payload)


def encode_blob(arr, legacy_shape=False):
    arr = np.ascontiguousarray(arr, dtype="<f4")
    out = b""
    if legacy_shape:
        dims = (1,) * (4 - arr.ndim) + arr.shape
        for k, d in zip((1, 2, 3, 4), dims):
            out += _enc_varint((k << 3) | _VARINT) + _enc_varint(int(d))
    else:
        out += _enc_len(7, _enc_len(1, b"".join(_enc_varint(int(d)) for d in arr.shape)))
    return out + _enc_len(5, arr.tobytes())


def write_binaryproto(path, arr, legacy_shape=True):
    with open(path, "wb") as fh:
        fh.write(encode_blob(arr, legacy_shape))


def write_caffemodel(path, layers, v1=False):
    """layers: iterable of (name, type, [blobs]).  v1: the pre-2015 V1LayerParameter layout."""
    out = _enc_len(1, b"vp_net")
    for name, typ, blobs in layers:
        if v1:
            body = _enc_len(4, name.encode())
            for b in blobs:
                body += _enc_len(6, encode_blob(b, legacy_shape=True))
            out += _enc_len(2, body)
        else:
            body = _enc_len(1, name.encode()) + _enc_len(2, typ.encode())
            for b in blobs:
                body += _enc_len(7, encode_blob(b))
            out += _enc_len(100, body)
    with open(path, "wb") as fh:
        fh.write(out)
