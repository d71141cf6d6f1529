"""One-off parity soak on a GPU box: the full-size randomised comparison of tests/test_gpu_parity.py with other seeds and more checked instances.
usage: python tools/soak_parity.py [seeds...]   (default 1 2 3; H1 batch 4096 all gaits + G1 batch 8192)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T
from oracle import pyoracle
pyoracle.build()
seeds = [int(a) for a in sys.argv[1:]] or [1, 2, 3]
for s in seeds:
    t0 = time.time(); T._full_size_randomized(T.MODEL, 4096, s, all_gaits=True, n_check=128); print("h1 seed", s, "ok %.1fs" % (time.time() - t0), flush=True)
    t0 = time.time(); T._full_size_randomized(os.path.join(ROOT, "configs", "g1.model"), 8192, 10 + s, all_gaits=True, n_check=96); print("g1 seed", 10 + s, "ok %.1fs" % (time.time() - t0), flush=True)
print("soak passed")
