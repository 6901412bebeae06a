"""Dev tool (run under compute-sanitizer on the GPU box): one small pass through every kernel family --
micro forward (N = 17, single-block attention, graph off), tiny forward (N = 197, two query tiles, TMA-store epilogue),
the two-sweep attention (N = 577) and the mma.sync attention, the GPU preprocessing path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["VITB200_GRAPH"] = "0"
from tests.util import pkg, gf, model_path  # noqa: E402

eng = pkg.engine
for cfg, n in (("micro", 3), ("micro14", 2), ("tiny", 2)):
    m = eng.vit_model_load(model_path(cfg, "f16"), 0, 4)
    imgs = gf.synthetic_images(n, m.img_size, seed=5)
    probs, idx, val, logits = eng.vit_predict(m, imgs, 5, want_logits=True)
    assert np.isfinite(logits).all()
    print(cfg, "ok", idx[:, 0])
    if cfg == "micro":
        u8 = [np.random.default_rng(1).integers(0, 256, (50, 70, 3), dtype=np.uint8)]
        eng.vit_image_preprocess_predict(m, u8)
        print("preprocess ok")
    m.close()
rng = np.random.default_rng(0)
for N, H, k in ((577, 2, eng.ATTN_TC_LONG), (225, 1, eng.ATTN_TC_LONG), (577, 1, eng.ATTN_MMA), (197, 2, eng.ATTN_TC)):
    qkv = rng.normal(0, 1, (2 * N, 3 * H * 64)).astype(np.float16)
    out = eng.test_attention(qkv, 2, N, H, k)
    assert np.isfinite(out).all()
    print("attention", N, H, k, "ok")
# round 2: the split-precision attention entry point, the hi-lo GEMM epilogue, the q8_0 integer tensor-core prototype
N, H = 197, 2
x = rng.normal(0, 1, (2 * N, 3 * H * 64)).astype(np.float32)
out = eng.test_attention_hilo(x, 2, N, H)
assert np.isfinite(out).all()
print("attention hi-lo ok")
A = rng.normal(0, 1, (300, 192)).astype(np.float16)
W = rng.normal(0, 0.05, (576, 192)).astype(np.float16)
for epi in (0, 1, 2, 4, 6):
    out = eng.test_gemm(300, 576, 192, epi, A, W, np.zeros(576, np.float32), resid=np.zeros((300, 576), np.float32) if epi == 2 else None)
    assert np.isfinite(out).all()
print("gemm epilogues ok")
xq = rng.normal(0, 1, (300, 256)).astype(np.float32)
wb = pkg.convert.quantize_q8_0_reference((rng.normal(0, 0.05, (132, 256))).astype(np.float32).reshape(-1))
y, _, _, _ = eng.test_gemm_q8(xq, wb, np.zeros(132, np.float32))
assert np.isfinite(y).all()
print("q8_0 gemm ok")
