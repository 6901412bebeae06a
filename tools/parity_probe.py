"""Dev tool (run under gpurun): parity numbers of the golden configs, without assertions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import pkg, gf, model_path  # noqa: E402

eng = pkg.engine
for cfg in sys.argv[1:] or ["micro", "tiny", "base"]:
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{cfg}_f16.npz"))
    m = eng.vit_model_load(model_path(cfg, "f16"), 0, 4)
    imgs = gf.synthetic_images(int(g["n_images"]), m.img_size, seed=int(g["image_seed"]))
    probs, idx, val, logits = eng.vit_predict(m, imgs, 5, want_logits=True)
    re = np.abs(logits - g["logits"]).max(1) / np.abs(g["logits"]).max(1)
    l2 = np.linalg.norm(logits - g["logits"], axis=1) / np.linalg.norm(g["logits"], axis=1)
    print(cfg, "max-norm rel", re, "L2 rel", l2, "prob abs", np.abs(probs - g["probs"]).max(), "top5 same",
          (np.argsort(-g["logits"], 1)[:, :5] == idx).all(1))
    m.close()
