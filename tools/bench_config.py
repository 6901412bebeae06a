"""Dev tool (run under gpurun): per-kernel time of one configuration through the C ABI with profiling events.
usage: python tools/bench_config.py <config> <batch> [ftype]   e.g.  large384 128 bf16w"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import pkg, gf, model_path  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "large384"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ft = sys.argv[3] if len(sys.argv) > 3 else "f16"
eng = pkg.engine
L = eng.lib()
t0 = time.time()
path = model_path(cfg, ft)
print(f"model file ready in {time.time() - t0:.1f}s", flush=True)
m = eng.vit_model_load(path, 0, B)
imgs = gf.synthetic_images(B, m.img_size, seed=5)
for _ in range(3):
    eng.vit_predict(m, imgs, 5)
L.vitb200_profile_enable(m.handle, 1)
steps = 5
t0 = time.time()
for _ in range(steps):
    eng.vit_predict(m, imgs, 5)
wall = (time.time() - t0) / steps
names = ["patch", "qkv", "proj", "fc1", "fc2", "head", "attention", "layernorm"]
tot = 0.0
for k, n in enumerate(names):
    ms, nl, fl = C.c_double(), C.c_int(), C.c_double()
    L.vitb200_profile_read(m.handle, k, C.byref(ms), C.byref(nl), C.byref(fl))
    per = ms.value / steps
    tot += per
    tf = (fl.value * nl.value / steps) / (per * 1e-3) / 1e12 if per > 0 and fl.value > 0 else float("nan")
    print(f"{n:10s} {per:8.3f} ms/step  {nl.value // steps:4d} launches  {tf:8.1f} TFLOP/s")
print(f"sum of tracked kernels {tot:.3f} ms/step -> {B / tot * 1e3:.0f} img/s (device);  host wall incl. H2D {wall * 1e3:.1f} ms/step")
