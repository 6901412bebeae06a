"""Dev experiment (GPU): does running the batch as two half-batch forwards on two streams (two engines, zero kernel changes) let the
memory-bound kernels (LayerNorm, patchify) of one half overlap the tensor-bound GEMMs of the other?  Device-timed, K steps each."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import gf, model_path, pkg  # noqa: E402

eng = pkg.engine
L = eng.lib()
B, K, W = 256, 20, 3
path = model_path("base", "f16")
imgs = [torch.from_numpy(gf.synthetic_images(B, 224, seed=1234 + j)).cuda() for j in range(2)]


def run(parts, pdl=True):
    hb = B // parts
    models = [eng.vit_model_load(path, 0, hb) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    probs = [torch.empty(hb, 1000, device="cuda") for _ in range(parts)]
    idx = [torch.empty(hb, 5, dtype=torch.int32, device="cuda") for _ in range(parts)]
    val = [torch.empty(hb, 5, device="cuda") for _ in range(parts)]

    def step(i):
        for p in range(parts):
            src = imgs[i & 1][p * hb:(p + 1) * hb]
            rc = L.vitb200_forward_device(models[p].handle, src.data_ptr(), hb, probs[p].data_ptr(), None, idx[p].data_ptr(), val[p].data_ptr(), 5,
                                          C.c_void_p(streams[p].cuda_stream))
            if rc:
                raise RuntimeError(L.vitb200_last_error().decode())
    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    for i in range(K):
        step(i)
    for s in streams:
        ev = torch.cuda.Event()
        ev.record(s)
        main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    top1 = int(idx[0][0, 0])
    for m in models:
        m.close()
    return ms, top1


for parts in (1, 2, 4, 1, 2):
    ms, t1 = run(parts)
    print(json.dumps({"parts": parts, "ms_per_step": ms, "images_per_s": B / ms * 1e3, "top1": t1}), flush=True)
