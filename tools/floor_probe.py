"""Dev tool (CPU): parity noise floor of the restatement's variants against its own bit-exact mode (== the reference).
usage: python tools/floor_probe.py <config> <n_images> [variants...]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import gf, model_path  # noqa: E402
from oracle import restatement as rs  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "tiny"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
variants = [int(v) for v in sys.argv[3:]] or [1, 2, 3, 4]
vf = gf.read(model_path(cfg, "f16"))
om = rs.OracleModel(vf, gf.tensor_specs)
rs.set_threads(8)
imgs = gf.synthetic_images(n, vf.img_size, seed=4242)
L = rs.lib()
L.vo_set_variant.argtypes = [C.c_int]
L.vo_set_variant(0)
_, l0 = om.forward_batch(imgs)
for v in variants:
    L.vo_set_variant(v)
    _, lv = om.forward_batch(imgs)
    re = np.abs(lv - l0).max(1) / np.abs(l0).max(1)
    print(f"{cfg} variant {v}: median {np.median(re):.3e} p90 {np.quantile(re, 0.9):.3e} max {re.max():.3e} mean {re.mean():.3e}", flush=True)
L.vo_set_variant(0)
