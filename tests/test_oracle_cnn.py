"""CNN oracle: layer shapes / parameter count of cnn/deploy.prototxt."""
import numpy as np

from oracle import cnn_oracle


def test_layer_shapes_and_param_count():
    ws, bs = cnn_oracle.random_weights(0)
    assert sum(w.size for w in ws) + sum(b.size for b in bs) == 256_664_656
    img = np.random.RandomState(0).randint(0, 256, (1, 500, 500)).astype(np.uint8)
    sig, logits, layers = cnn_oracle.forward(img, ws, bs, return_layers=True)
    want = {"conv1": (96, 123, 123), "pool1": (96, 61, 61), "conv2": (256, 61, 61), "pool2": (256, 30, 30),
            "conv3": (384, 30, 30), "conv4": (384, 30, 30), "conv5": (256, 30, 30), "pool5": (256, 15, 15),
            "fc6": (4096,), "fc7": (4096,)}
    for k, shp in want.items():
        assert layers[k].shape[1:] == shp, k
    assert sig.shape == (1, 20, 20) and logits.shape == (1, 400)
    assert np.all((sig > 0) & (sig < 1))
