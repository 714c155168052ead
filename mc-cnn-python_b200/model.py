"""Drop-in for the reference's ``src/model.py``: the MC-CNN-fast Siamese tower ``NET``.

The reference builds a TensorFlow graph over a placeholder ``x`` and exposes ``.features``
(model.py:9-65).  Here ``x`` is the image batch itself ([N,H,W,1], [H,W,1] or [H,W]; NumPy or CUDA
tensor) and ``.features`` runs the hand-written CUDA forward pass through the C ABI
(mccnn_features): ``num_conv_layers`` 3x3 VALID cross-correlations with 64 maps, ReLU after all but
the last, channel L2-normalisation.  Like the reference's graph, NET applies NO padding
(compute_features pads, pf:20-25), so features are [N, H-2n, W-2n, 64].
"""
import numpy as np

try:
    from . import process_functional as _pf
except ImportError:
    import process_functional as _pf


class NET(object):

    def __init__(self, x, weights_path='DEFAULT', input_patch_size=11, num_conv_layers=5,
                 num_conv_feature_maps=64, conv_kernel_size=3, batch_size=128):
        self.X = x
        self.batch_size = batch_size
        self.input_patch_size = input_patch_size
        self.num_conv_layers = int(num_conv_layers)
        self.num_conv_feature_maps = int(num_conv_feature_maps)
        self.conv_kernel_size = int(conv_kernel_size)
        assert self.num_conv_layers >= 2, "num of conv layers: at least 2 (model.py:44)"
        assert self.num_conv_feature_maps == 64 and self.conv_kernel_size == 3, \
            "the CUDA path implements the fast architecture's 64 maps / 3x3 kernels (model.py:14-16)"
        self.WEIGHTS_PATH = 'pretrain.npy' if weights_path == 'DEFAULT' else weights_path   # model.py:25-28
        # tf.get_variable's default initialiser, seeded (model.py:100-101)
        self._weights = _pf.resolve_weights(None, num_layers=self.num_conv_layers)
        self._features = None

    # -- weights ---------------------------------------------------------------------------------
    def set_weights(self, weights, biases):
        """HWIO weight arrays [3,3,1,64], [3,3,64,64]... and bias arrays [64]."""
        self._weights = _pf.DeviceWeights(weights, biases)
        self._features = None

    def restore(self, checkpoint_prefix):
        """What ``tf.train.Saver().restore(sess, checkpoint)`` does at pf:32/:43, without TensorFlow."""
        self._weights = _pf.resolve_weights(checkpoint_prefix, num_layers=self.num_conv_layers)
        assert self._weights.num_layers == self.num_conv_layers
        self._features = None

    def load_initial_weights(self, session=None):
        """model.py:67-77: load {var_name: array} from the .npy dict at WEIGHTS_PATH."""
        d = np.load(self.WEIGHTS_PATH, encoding='bytes', allow_pickle=True).item()
        d = {(k.decode() if isinstance(k, bytes) else k): v for k, v in d.items()}
        ws = [d['conv%d/weights:0' % i] for i in range(1, self.num_conv_layers + 1)]
        bs = [d['conv%d/biases:0' % i] for i in range(1, self.num_conv_layers + 1)]
        self.set_weights(ws, bs)

    def save_weights(self, session=None, file_name='pretrain.npy'):
        """model.py:79-85."""
        d = {}
        for i, (w, b) in enumerate(zip(self._weights.w, self._weights.b), 1):
            d['conv%d/weights:0' % i] = w.cpu().numpy()
            d['conv%d/biases:0' % i] = b.cpu().numpy()
        np.save(file_name, d)

    # -- forward ---------------------------------------------------------------------------------
    @property
    def features(self):
        if self._features is None:
            self._features = self.forward(self.X)
        return self._features

    def forward(self, x):
        torch = _pf._torch()
        like = x
        t = _pf._to_dev(x)
        if t.ndim == 2:
            batch = t[None]
        elif t.ndim == 3:
            assert t.shape[2] == 1, "grayscale input expected (ic = 1, model.py:40)"
            batch = t[None, :, :, 0]
        else:
            assert t.ndim == 4 and t.shape[3] == 1, "expected [N,H,W,1]"
            batch = t[:, :, :, 0]
        outs = [_pf.net_forward(batch[i].contiguous(), self._weights, 0) for i in range(batch.shape[0])]
        out = torch.stack(outs, 0)
        if t.ndim < 4:
            out = out[0]
        return _pf._ret(out, like)
