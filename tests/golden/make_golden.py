"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libvitref.so, built from
/root/reference by oracle/Makefile).  Run in the dev container:  python tests/golden/make_golden.py

Each fixture holds, for a seeded synthetic model file (vit.cpp_b200/ggml_file.py) and seeded synthetic
images, the reference's pre-softmax logits and its probabilities (vit_predict, reference
vit.cpp:1004-1075), plus the sha256 of the model file so a drifting numpy RNG is detected."""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import gf, model_path  # noqa: E402
from oracle import ref  # noqa: E402

CASES = [  # (config, ftype tag, n_images)
    ("micro", "f16", 4), ("micro14", "f16", 4), ("micro", "f32", 2), ("micro", "q8_0", 2),
    ("tiny", "f16", 4), ("tiny", "q8_0", 2), ("base", "f16", 4), ("base", "q8_0", 2),
    ("micro", "q4_0", 2), ("micro", "q4_1", 2), ("micro", "q5_0", 2), ("micro", "q5_1", 2),
    ("tiny", "q4_0", 2), ("tiny", "q5_1", 2), ("base", "q4_1", 2), ("base", "q5_0", 2),
]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


VITSTR_CASES = [("vitstr_micro", "f16", 3), ("vitstr_tiny", "f16", 2)]  # the reference's ViTSTR extension (25 x classes per image)


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    only = set(sys.argv[1:])  # optional: regenerate just these "<config>_<ftype>" fixtures
    for cfg, ft, n in VITSTR_CASES:
        if only and f"{cfg}_{ft}" not in only:
            continue
        path = model_path(cfg, ft)
        m = ref.VitstrRefModel(path)
        imgs = gf.synthetic_gray_images(n, m.img, seed=1234)
        probs, logits = m.predict_batch(imgs, n_threads=8)
        np.savez_compressed(os.path.join(out_dir, f"{cfg}_{ft}.npz"), logits=logits, probs=probs,
                            image_seed=1234, n_images=n, model_sha256=sha256(path))
        print(cfg, ft, "top1 of tokens 0..4", logits.argmax(-1)[:, :5].tolist(), "sha", sha256(path)[:12])
        m.close()
    for cfg, ft, n in CASES:
        if only and f"{cfg}_{ft}" not in only:
            continue
        path = model_path(cfg, ft)
        m = ref.RefModel(path)
        imgs = gf.synthetic_images(n, m.img, seed=1234)
        probs, logits = m.predict_batch(imgs, n_threads=8)
        np.savez_compressed(os.path.join(out_dir, f"{cfg}_{ft}.npz"), logits=logits, probs=probs,
                            image_seed=1234, n_images=n, model_sha256=sha256(path))
        print(cfg, ft, "top1", logits.argmax(1), "sha", sha256(path)[:12])
        m.close()


if __name__ == "__main__":
    main()
