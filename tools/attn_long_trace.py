"""Dev tool (run under gpurun): phase timeline of the two-sweep attention kernel (ViT-L/16-384 geometry, 577 tokens), clock64 stamps of
CTA 0's first 16 query tiles.  usage: python tools/attn_long_trace.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = os.environ.setdefault("VITB200_ATTN_TRACE", os.path.join(ROOT, "gpurun_out", "attn_long_trace.txt"))
from tests.util import pkg  # noqa: E402

eng = pkg.engine
B, N, H = 32, 577, 16
rng = np.random.default_rng(0)
qkv = rng.normal(0, 1, (B * N, 3 * H * 64)).astype(np.float16)
eng.test_attention(qkv, B, N, H, eng.ATTN_TC_LONG)
t = np.loadtxt(path).astype(np.int64)
t0 = t[0][t[0] > 0].min()
names = ["start", "sweepA_end", "max_xchg", "sweepB_end", "sum_xchg", "o_ready", "stored"]
for i in range(12):
    row = t[i]
    r = lambda j: int(row[j] - t0) if row[j] else -1
    for w in range(2):
        print(f"tile {i:2d} WG{w} " + " ".join(f"{n}={r(8 * w + k)}" for k, n in enumerate(names)) + " | B blocks done " + " ".join(str(r(16 + 4 * w + k)) for k in range(4)))
