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
