import importlib, os, sys
sys.path.insert(0, "/root/repo")
import torch
from bench import synth_pair
pkg = importlib.import_module("mc-cnn-python_b200")
m = pkg.StereoMatcher(1024, 1024, 192, stages=("features",))
li, ri = synth_pair(1024, 1024, 37, seed=0); m.set_images(li, ri)
for _ in range(3): m.run()
print({k: round(v, 3) for k, v in m.run_timed().items()})
