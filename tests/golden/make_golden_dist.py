"""Generate the per-image parity fixtures of the headline configuration from the UNMODIFIED reference (oracle/_ref) and the
restatement's noise-floor variant.  Run in the dev container:  python tests/golden/make_golden_dist.py [base|large]

  base_f16_b64.npz       64 seeded images of the batch-256 workload (vit_base_patch16_224 f16): reference logits + probabilities, and
                         `floor` = per-image max|dlogit| / max|ref logit| of the restatement run with double-precision accumulation and
                         the reference's rounding points (vo_set_variant(1)) against the reference itself -- the distance between two
                         CORRECT implementations that are not bit-identical, on exactly these images.
  large384_bf16w_b8.npz  8 images of vit_large_patch16_384 with bf16-representable weights in the f32 container: reference logits."""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import gf, model_path  # noqa: E402
from oracle import ref, restatement as rs  # noqa: E402


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


def base(n=64, seed=777):
    path = model_path("base", "f16")
    imgs = gf.synthetic_images(n, 224, seed=seed)
    m = ref.RefModel(path)
    probs, logits = m.predict_batch(imgs, n_threads=8)
    m.close()
    vf = gf.read(path)
    om = rs.OracleModel(vf, gf.tensor_specs)
    rs.set_threads(8)
    L = rs.lib()
    L.vo_set_variant.argtypes = [C.c_int]
    L.vo_set_variant(0)
    _, l0 = om.forward(imgs[0])
    assert np.array_equal(l0, logits[0]), "restatement is not bit-exact against the reference on this host"
    L.vo_set_variant(1)
    floor = np.empty(n, np.float64)
    for i in range(n):
        _, lv = om.forward(imgs[i])
        floor[i] = np.abs(lv - logits[i]).max() / np.abs(logits[i]).max()
        print("base floor", i, floor[i], flush=True)
    L.vo_set_variant(0)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "base_f16_b64.npz")
    np.savez_compressed(out, logits=logits, probs=probs, floor=floor, image_seed=seed, n_images=n, model_sha256=sha256(path))
    print("wrote", out, "floor median %.3e p90 %.3e max %.3e" % (np.median(floor), np.quantile(floor, 0.9), floor.max()))


def large(n=8, seed=555):
    path = model_path("large384", "bf16w")
    imgs = gf.synthetic_images(n, 384, seed=seed)
    m = ref.RefModel(path)
    probs, logits = m.predict_batch(imgs, n_threads=8)
    m.close()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "large384_bf16w_b8.npz")
    np.savez_compressed(out, logits=logits, probs=probs, image_seed=seed, n_images=n, model_sha256=sha256(path))
    print("wrote", out, "top1", logits.argmax(1))


if __name__ == "__main__":
    which = sys.argv[1:] or ["base", "large"]
    if "large" in which:
        large()
    if "base" in which:
        base()
