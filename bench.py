#!/usr/bin/env python
"""bench.py -- images/sec of the ViT-B/16 224^2 forward path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own ggml CPU path (oracle/_ref) on the host cores

A "step" is one forward pass (pixels f32 HWC -> probabilities + top-5) over one batch of 256 synthetic images per
GPU (BASELINE.json configs[1]: vit_base_patch16_224 f16, batch=256, 1xB200).  Images shard across ranks, weights are
replicated, there is no collective on the data path => "scaling": "weak".

`value`  : whole-job images/s with inputs already resident in HBM (CUDA events on the launching stream, max over ranks).
`e2e`    : same metric through the reference-facing C-ABI call with HOST (pinned) buffers: H2D of the 154 MB batch and
           D2H of probabilities + top-5 inside the timed region.
`roofline`: tensor-core roofline of the dominant kernel (the fc1 / fc2 tcgen05 GEMM), timed live with CUDA events
           around its launches inside the same timed steps; peak = MEASURED_PEAKS.json (sustained bf16/f16 GEMM).
`cpu_baseline`: the reference (oracle/_ref, unmodified ggml CPU path) timed on this box's host cores, bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 256
MODEL_CFG = "base"
FLOPS_PER_IMAGE = 35.128e9  # BASELINE.md section 3 / SURVEY.md 8(d): 2 * 17.564 GMAC, ViT-B/16 224^2


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each GEMM of the step, from the committed
    `ncu --set full` capture (profiles/traffic.json, written by tools/summarize_ncu.py); None when no capture exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "hbm_gbs": float(d.get("hbm_gbs", 6650.0)),
                "which": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "which": "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained, 6.65 TB/s)"}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_reference_rate(n_images, threads_list):
    """images/s of the reference vit_predict (oracle/_ref, reference vit.cpp:1004) on host cores; model load excluded."""
    from tests.util import gf, model_path
    from oracle import ref
    m = ref.RefModel(model_path(MODEL_CFG, "f16"))
    imgs = gf.synthetic_images(n_images, 224, seed=4321)
    m.predict(imgs[0], threads_list[0])  # warm (page in weights)
    best = None
    for nt in threads_list:
        t = time.perf_counter()
        for i in range(n_images):
            m.predict(imgs[i], nt)
        dt = time.perf_counter() - t
        rate = n_images / dt
        if best is None or rate > best[0]:
            best = (rate, nt, dt)
    m.close()
    return best


def host_threads():
    n = os.cpu_count() or 8
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return n


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    nproc = host_threads()
    cand = sorted({min(nproc, 64), max(1, min(nproc, 64) // 2), max(1, min(nproc, 64) // 4)}, reverse=True)
    per_step = 4
    # choose the better thread count on a short probe, then time exactly `steps` steps of `per_step` images
    rate, nt, _ = cpu_reference_rate(2, cand)
    from tests.util import gf, model_path
    from oracle import ref
    m = ref.RefModel(model_path(MODEL_CFG, "f16"))
    imgs = gf.synthetic_images(per_step, 224, seed=99)
    for _ in range(min(args.warmup, 1)):
        m.predict(imgs[0], nt)
    t = time.perf_counter()
    for _ in range(args.steps):
        for i in range(per_step):
            m.predict(imgs[i], nt)
    dt = time.perf_counter() - t
    val = args.steps * per_step / dt
    line = {"impl": "reference", "metric": "images/sec ViT-B/16 224^2 forward", "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "vit_base_patch16_224 f16 legacy-ggml file, reference ggml CPU vit_predict, %d images per step (bounded sample of the batch-256 workload)" % per_step,
                       "threads": nt, "host_cores": nproc},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": nt, "kind": "reference",
                             "sample": "%d steps x %d images through oracle/_ref vit_predict" % (args.steps, per_step)},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from tests.util import pkg, gf, model_path

    eng = pkg.engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the vit.cpp_b200 forward path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B = args.batch
    W = max(args.warmup, 3)
    L = eng.lib()

    # synthetic weights (legacy-ggml file through the product loader) + synthetic inputs, per rank
    if rank == 0:
        path = model_path(MODEL_CFG, "f16")
    if world > 1:
        dist.barrier()
    path = model_path(MODEL_CFG, "f16")
    model = eng.vit_model_load(path, device=local_rank, max_batch=B)
    # two distinct input batches (2 x 154 MB > 126 MB L2), alternated so no step finds its input in L2
    host_imgs = [torch.from_numpy(gf.synthetic_images(B, 224, seed=1234 + 17 * rank + j)).pin_memory() for j in range(2)]
    dev_imgs = [h.cuda(non_blocking=True) for h in host_imgs]
    d_probs = torch.empty(B, model.num_classes, device="cuda")
    d_idx = torch.empty(B, 5, dtype=torch.int32, device="cuda")
    d_val = torch.empty(B, 5, device="cuda")
    # a non-default torch stream: its handle is what the C ABI launches on (NULL would mean "engine's own stream"),
    # and torch.cuda.Event timings on it see exactly those launches
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step_device(i):
        rc = L.vitb200_forward_device(model.handle, dev_imgs[i & 1].data_ptr(), B, d_probs.data_ptr(), None, d_idx.data_ptr(),
                                      d_val.data_ptr(), 5, stream)
        if rc != 0:
            raise RuntimeError(L.vitb200_last_error().decode())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-data throughput (region A: the product path, kernel schedule replayed as a CUDA graph) --------
    for i in range(W + 2):  # first two calls per input buffer run eagerly / capture the graph
        step_device(i)
    sync_all()
    def timed_region():
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step_device(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), sampler.result()

    ms_local, clocks = timed_region()
    # a run that saw a hardware / thermal slowdown is rejected and re-measured once (sw_power_cap is kept and reported)
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    again = torch.tensor([1 if bad & set(clocks.get("reasons", [])) else 0], device="cuda", dtype=torch.int32)
    if world > 1:
        dist.all_reduce(again, op=dist.ReduceOp.MAX)  # every rank repeats if any rank was throttled (the barriers are collective)
    if int(again.item()):
        first = clocks
        sync_all()
        ms_local, clocks = timed_region()
        clocks["remeasured_after"] = first.get("reasons", [])
    launches_per_step = model.last_launch_count()
    # ---- region B: the same K steps launched eagerly with CUDA-event pairs around every tracked kernel (roofline) --------
    L.vitb200_profile_enable(model.handle, 1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(args.steps):
        step_device(i)
    p1.record()
    torch.cuda.synchronize()
    ms_profiled = p0.elapsed_time(p1)
    prof = {}
    for kind, name in enumerate(["patch", "qkv", "proj", "fc1", "fc2", "head", "attention", "layernorm"]):
        ms, cnt, fl = C.c_double(), C.c_int(), C.c_double()
        if L.vitb200_profile_read(model.handle, kind, C.byref(ms), C.byref(cnt), C.byref(fl)) == 0 and cnt.value > 0:
            prof[name] = {"ms_total": ms.value, "launches": cnt.value, "flops_per_launch": fl.value}
    L.vitb200_profile_enable(model.handle, 0)
    t = torch.tensor([ms_local], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (H2D + forward + D2H) ------------------------------
    # Every step copies its 154 MB input batch from pinned host memory and reads probabilities + top-5 back into pinned
    # host memory; vitb200_forward_async double-buffers so step i+1's H2D overlaps step i's kernels.
    h_probs = [torch.empty(B, model.num_classes).pin_memory() for _ in range(2)]
    h_idx = [torch.empty(B, 5, dtype=torch.int32).pin_memory() for _ in range(2)]
    h_val = [torch.empty(B, 5).pin_memory() for _ in range(2)]

    def step_e2e(i):
        j = i & 1
        rc = L.vitb200_forward_async(model.handle, host_imgs[j].data_ptr(), B, h_probs[j].data_ptr(), None, h_idx[j].data_ptr(),
                                     h_val[j].data_ptr(), 5)
        if rc != 0:
            raise RuntimeError(L.vitb200_last_error().decode())
    for i in range(2):
        step_e2e(i)
    L.vitb200_sync(model.handle)
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    L.vitb200_sync(model.handle)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    top1_check = int(h_idx[0][0, 0])

    # optional tail of the north-star design: gather the top-k pairs on rank 0 over NCCL (not on the timed path)
    if world > 1:
        gathered = [torch.empty_like(d_idx) for _ in range(world)] if rank == 0 else None
        dist.gather(d_idx, gathered, dst=0)

    if rank == 0:
        peaks = measured_peaks()
        # dominant kernel: the GEMM family member with the largest share of the step
        dom = max((k for k in prof if k in ("qkv", "proj", "fc1", "fc2")), key=lambda k: prof[k]["ms_total"], default=None)
        roof = None
        if dom:
            d = prof[dom]
            avg_s = d["ms_total"] * 1e-3 / d["launches"]
            ach = d["flops_per_launch"] / avg_s / 1e12
            roof = {"bound": "tensor", "kernel": f"gemm_tcgen05_kernel ({dom})", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tflops"], "traffic": ncu_traffic().get(dom), "traffic_unit": "bytes per launch (ncu dram read+write)",
                    "peak_source": peaks["which"],
                    "avg_launch_ms": avg_s * 1e3, "share_of_step": d["ms_total"] / ms_profiled,
                    "timed_in": "a second pass of the same %d steps, launched eagerly with a CUDA-event pair around every tracked kernel "
                                "(%.3f ms/step vs %.3f ms/step for the graph-replayed pass that `value` reports)" % (args.steps, ms_profiled / args.steps, ms_local / args.steps)}
        line = {
            "metric": "images/sec ViT-B/16 224^2 forward", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": "vit_base_patch16_224 f16 (legacy-ggml model file), batch=%d per GPU, 224x224x3 f32 HWC synthetic images" % B,
                       "global_batch": B * world, "parallelism": "dp%d (images sharded, weights replicated, no collective on the data path)" % world,
                       "l2": "two alternating 154 MB input batches + >1 GB of activations per step exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * 224 * 224 * 4,
                    "d2h_bytes_per_step": B * (model.num_classes * 4 + 5 * 4 + 5 * 4)},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": roof,
            "model_tflops": value * FLOPS_PER_IMAGE / 1e12 / world,
            "model_frac_of_peak": value * FLOPS_PER_IMAGE / 1e12 / world / peaks["tflops"],
            "kernels": {k: {"ms_per_step": v["ms_total"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                            "tflops": (v["flops_per_launch"] * v["launches"] / (v["ms_total"] * 1e-3) / 1e12) if v["flops_per_launch"] else None}
                        for k, v in prof.items()},
            "top1_sample": top1_check,
        }
        if world == 1 and not args.no_cpu_baseline:
            nproc = host_threads()
            cand = sorted({min(nproc, 64), max(1, min(nproc, 64) // 2), max(1, min(nproc, 64) // 4)}, reverse=True)
            rate, nt, dt = cpu_reference_rate(8, cand)
            line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": nt, "kind": "reference",
                                    "sample": "8 images of the same ViT-B/16 f16 model through oracle/_ref vit_predict (%.1f s), host has %d cores" % (dt, nproc)}
        print(json.dumps(line), flush=True)
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
