"""Max abs error of the CUDA feature net against the float64 golden features (glorot weights and the shipped checkpoint),
for the operand format in force (MCCNN_CONV_TF32=1: TF32 split, default: FP16 split), plus a scaled-input stress."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module("mc-cnn-python_b200")
pf = pkg.process_functional
import oracle as O
O.build()
g = np.load(os.path.join(ROOT, "tests", "golden", "features_ckpt.npz"))
ck = np.load(os.path.join(ROOT, "tests", "golden", "checkpoint_tensors.npz"))
img = g["image"]
ws, bs = pf.glorot_uniform_weights(seed=int(g["glorot_seed"]))
fl, _ = pf.compute_features(img[..., None], img[..., None], 11, 11, (ws, bs))
print("glorot weights:      max |err| vs float64 golden %.3e" % np.abs(fl - g["features_glorot"]).max())
cw = [ck["conv%d_weights" % i] for i in range(1, 6)]; cb = [ck["conv%d_biases" % i] for i in range(1, 6)]
fl, _ = pf.compute_features(img[..., None], img[..., None], 11, 11, (cw, cb))
print("shipped checkpoint:  max |err| vs float64 golden %.3e" % np.abs(fl - g["features"]).max())
rng = np.random.default_rng(1)
for scale in (1.0, 1e-3, 30.0):
    big = (rng.standard_normal((64, 256)) * scale).astype(np.float32)
    f, _ = pf.compute_features(big[..., None], big[..., None], 11, 11, (cw, cb))
    ref = O.net_forward(big, cw, cb)
    print("checkpoint, input x %-6g max |err| vs C oracle %.3e  finite %s" % (scale, np.abs(f - ref).max(), np.isfinite(f).all()))
