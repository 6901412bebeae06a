"""Dev tool (GPU box with N GPUs): the single-process data-parallel model SURVEY.md 8(e) specifies -- ONE host thread, one engine per
GPU, `vitb200_forward_sharded_async` over a global batch of 256 images per GPU in pinned host memory -- timed end to end (H2D of the
shards, forward, D2H of probabilities + top-5) next to what `bench.py` measures under torchrun (one process per GPU).
usage: python tools/bench_sharded.py [--gpus N] [--steps K] [--warmup W]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import gf, model_path, pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
eng = pkg.engine
L = eng.lib()
G = args.gpus or torch.cuda.device_count()
B = 256
models = [eng.vit_model_load(model_path("base", "f16"), device=g, max_batch=B) for g in range(G)]
hs = (C.c_void_p * G)(*[m.handle for m in models])
nc = models[0].num_classes
imgs = [torch.from_numpy(np.concatenate([gf.synthetic_images(B, 224, seed=1234 + 17 * g + j) for g in range(G)])).pin_memory() for j in range(2)]
probs = [torch.empty(G * B, nc).pin_memory() for _ in range(2)]
idx = [torch.empty(G * B, 5, dtype=torch.int32).pin_memory() for _ in range(2)]
val = [torch.empty(G * B, 5).pin_memory() for _ in range(2)]
L.vitb200_forward_sharded_async.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
L.vitb200_sync_all.argtypes = [C.c_void_p, C.c_int]


def step(i):
    j = i & 1
    if L.vitb200_forward_sharded_async(hs, G, imgs[j].data_ptr(), G * B, probs[j].data_ptr(), None, idx[j].data_ptr(), val[j].data_ptr(), 5):
        raise RuntimeError(L.vitb200_last_error().decode())


for i in range(args.warmup):
    step(i)
L.vitb200_sync_all(hs, G)
t0 = time.perf_counter()
for i in range(args.steps):
    step(i)
L.vitb200_sync_all(hs, G)
dt = time.perf_counter() - t0
# blocking form (what vit_predict_sharded calls): no overlap between consecutive global batches
t1 = time.perf_counter()
for i in range(args.steps):
    if L.vitb200_forward_sharded(hs, G, imgs[i & 1].data_ptr(), G * B, probs[i & 1].data_ptr(), None, idx[i & 1].data_ptr(), val[i & 1].data_ptr(), 5):
        raise RuntimeError(L.vitb200_last_error().decode())
dt_block = time.perf_counter() - t1
print(json.dumps({"what": "single-process vitb200_forward_sharded over %d GPUs, 256 images per GPU per step, host (pinned) buffers" % G,
                  "n_gpus": G, "steps": args.steps, "images_per_s_pipelined": G * B * args.steps / dt, "ms_per_step_pipelined": dt / args.steps * 1e3,
                  "images_per_s_blocking": G * B * args.steps / dt_block, "ms_per_step_blocking": dt_block / args.steps * 1e3,
                  "top1_first_image": int(idx[0][0, 0])}))
