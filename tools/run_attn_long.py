"""Dev tool (run under gpurun / ncu): one launch of the two-sweep attention kernel on random QKV (ViT-L/16-384 geometry)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import pkg  # noqa: E402

eng = pkg.engine
B, N, H = int(os.environ.get("ATT_B", 32)), 577, 16
rng = np.random.default_rng(0)
qkv = rng.normal(0, 1, (B * N, 3 * H * 64)).astype(np.float16)
out = eng.test_attention(qkv, B, N, H, eng.ATTN_TC_LONG)
print("ok", float(np.abs(out).mean()))
