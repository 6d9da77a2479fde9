"""Stage 2 parity: tcgen05 GEMM / implicit-GEMM CNN (through the C ABI) vs the
torch-CPU fp32 oracle with identical weights."""
import numpy as np
import pytest
import torch

from oracle import cnn_oracle, sphere_oracle as so
from vanishing_points_2017_b200 import synth

pytestmark = pytest.mark.gpu

# BASELINE.json north_star: CNN logits within 1e-2 relative (bf16 operands, fp32 accumulation)
LOGIT_TOL = 1e-2


@pytest.fixture(scope="module")
def cnn():
    from vanishing_points_2017_b200 import cnn as m
    return m


def bf16_bits(x):
    t = torch.from_numpy(x).to(torch.bfloat16)
    return t.view(torch.int16).numpy().view(np.uint16), t.to(torch.float32).numpy()


@pytest.mark.parametrize("m,n,k,bn,ksplit", [(128, 128, 64, 128, 1), (128, 256, 256, 256, 1), (300, 384, 1024, 128, 1),
                                             (77, 192, 192, 192, 1), (1, 400, 4096, 80, 1), (515, 96, 192, 96, 1),
                                             # split-K as the fully connected layers use it (fc6: K = 57600, 9 splits)
                                             (102, 512, 57600, 256, 9), (102, 4096, 4096, 256, 8), (7, 400, 4096, 80, 16),
                                             (300, 256, 640, 128, 3), (102, 256, 704, 256, 4)])
def test_gemm_tcgen05(cnn, m, n, k, bn, ksplit):
    rs = np.random.RandomState(m + n + k)
    a_bits, a = bf16_bits(rs.standard_normal((m, k)).astype(np.float32))
    b_bits, b = bf16_bits(rs.standard_normal((n, k)).astype(np.float32))
    bias = rs.standard_normal(n).astype(np.float32)
    out = cnn.debug_gemm(a_bits, b_bits, bias=bias, relu=True, bn=bn, ksplit=ksplit)
    ref = np.maximum(a.astype(np.float64) @ b.astype(np.float64).T + bias, 0)
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-3 * np.sqrt(k))


def sphere_images(n, seed=0):
    imgs = []
    for i in range(n):
        sc = synth.make_scene(5000 + seed + i, 150 + 40 * i)
        imgs.append(so.votes_to_image(so.sphere_votes(sc["lines"], 500)) if i % 2 == 0
                    else so.curve_image(so.sphere_curve_counts(sc["lines"], 500), 0.1))
    return np.stack(imgs)


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.mark.parametrize("scale", [1.0, 3.0])
def test_cnn_logits_vs_fp32_oracle(cnn, scale):
    ws, bs = cnn_oracle.random_weights(seed=0, scale=scale)
    net = cnn.Net(__import__("vanishing_points_2017_b200")._lib.default_context(), ws, bs)
    imgs = sphere_images(3)
    sig, logits = net.forward_batch(imgs, want_logits=True)
    rsig, rlogits = cnn_oracle.forward(imgs, ws, bs)
    assert sig.shape == (3, 20, 20) and sig.dtype == np.float32
    assert rel_err(logits, rlogits) < LOGIT_TOL, rel_err(logits, rlogits)
    # |d sigmoid| <= |d logit| / 4
    np.testing.assert_allclose(sig, rsig, atol=0.25 * LOGIT_TOL * float(np.max(np.abs(rlogits))) + 1e-6)


def test_caffe_forward_drop_in_with_mean(cnn):
    net = cnn.init_caffe(None, None, 0)          # seeded train_val.prototxt fillers
    ws, bs = cnn_oracle.random_weights(seed=0)
    img = sphere_images(1, seed=7)[0]
    mean = np.random.RandomState(1).uniform(0, 20, (1, 1, 500, 500)).astype(np.float32)
    out = cnn.caffe_forward(net, img, mean)
    ref, _ = cnn_oracle.forward(img[None], ws, bs, mean=mean)
    assert out.shape == (20, 20)
    np.testing.assert_allclose(out, ref[0], atol=2e-3)
    out0 = cnn.caffe_forward(net, img, cnn.read_mean_blob(None))
    ref0, _ = cnn_oracle.forward(img[None], ws, bs)
    np.testing.assert_allclose(out0, ref0[0], atol=2e-3)


def test_batch_is_deterministic_and_order_independent(cnn):
    net = cnn.init_caffe(None, None, 0)
    imgs = sphere_images(5, seed=3)
    a = net.forward_batch(imgs)
    b = net.forward_batch(imgs[::-1].copy())[::-1]
    np.testing.assert_array_equal(a, b)


def test_persistent_and_one_tile_gemm_kernels_give_the_same_bits(cnn, monkeypatch):
    """The convolutions run through the persistent tcgen05 kernel (two accumulators in tensor memory) by default and
    through the one-tile-per-CTA kernel with VPK_GEMM_PERSIST=0: same MMA order per tile, same epilogue arithmetic."""
    net = cnn.init_caffe(None, None, 0)
    imgs = sphere_images(7, seed=11)
    a = net.forward_batch(imgs)
    monkeypatch.setenv("VPK_GEMM_PERSIST", "0")
    b = net.forward_batch(imgs)
    monkeypatch.delenv("VPK_GEMM_PERSIST")
    np.testing.assert_array_equal(a, b)
