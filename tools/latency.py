"""Dev tool (run under gpurun): single-image / small-batch latency of vit_predict through the C ABI (host buffers, graph replay)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import pkg, gf, model_path  # noqa: E402

eng = pkg.engine
for cfg in ("tiny", "base"):
    m = eng.vit_model_load(model_path(cfg, "f16"), 0, 16)
    for B in (1, 4, 16):
        imgs = gf.synthetic_images(B, m.img_size, seed=1)
        for _ in range(5):
            eng.vit_predict(m, imgs, 5)
        t = []
        for _ in range(30):
            t0 = time.perf_counter()
            eng.vit_predict(m, imgs, 5)
            t.append(time.perf_counter() - t0)
        print(f"{cfg} batch {B:2d}: median {np.median(t) * 1e3:.3f} ms, min {min(t) * 1e3:.3f} ms per call ({m.last_launch_count()} kernels)")
    m.close()
