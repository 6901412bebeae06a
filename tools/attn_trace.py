"""Dev tool (run under gpurun): dump and print the attention kernel's phase timeline (clock64 stamps of CTA 0).
usage: VITB200_ATTN_TRACE=gpurun_out/attn_trace.txt python tools/attn_trace.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import pkg, gf, model_path  # noqa: E402

path = os.environ.setdefault("VITB200_ATTN_TRACE", os.path.join(ROOT, "gpurun_out", "attn_trace.txt"))
eng = pkg.engine
m = eng.vit_model_load(model_path("base", "f16"), 0, 256)
imgs = gf.synthetic_images(256, m.img_size, seed=3)
os.environ["VITB200_GRAPH"] = "0"
eng.vit_predict(m, imgs, 5)
t = np.loadtxt(path).astype(np.int64)
t0 = t[0][t[0] > 0].min()
names = ["wait_s", "s_ready", "pass1_end", "p_arrive", "o_ready", "o_read"]
for i in range(12):
    row = t[i]
    def r(j):
        return int(row[j] - t0) if row[j] else -1
    print(f"prob {i:2d} | WG0 " + " ".join(f"{n}={r(k)}" for k, n in enumerate(names)))
    print(f"        | WG1 " + " ".join(f"{n}={r(8 + k)}" for k, n in enumerate(names)))
    print(f"        | WG0 chunks done " + " ".join(str(r(24 + k)) for k in range(7)))
    print(f"        | MMA t0: p_seen={r(16)} o_free={r(17)} issued={r(18)} | t1: p_seen={r(20)} o_free={r(21)} issued={r(22)}")
