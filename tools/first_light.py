"""Bring-up script for the GPU box: runs each stage in its own process (a trapped kernel poisons the CUDA
context) and prints error statistics instead of asserting.  Usage: python tools/first_light.py [stage ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def stage_gemm():
    from tests.util import pkg
    from oracle import restatement as rs
    eng = pkg.engine
    rng = np.random.default_rng(0)
    cases = [(128, 256, 64, 4), (128, 256, 64, 0), (128, 128, 64, 4), (128, 256, 128, 4), (256, 512, 192, 4),
             (300, 576, 192, 0), (300, 576, 192, 2), (1000, 768, 768, 2), (1000, 3072, 768, 1), (197, 1000, 768, 4),
             (50432, 2304, 768, 0)]
    for (M, N, K, epi) in cases:
        A = (rng.standard_normal((M, K)) * 1.0).astype(np.float16)
        W = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
        bias = rng.standard_normal(N).astype(np.float32)
        resid = rng.standard_normal((M, N)).astype(np.float32) if epi == 2 else None
        t = time.time()
        try:
            out = eng.test_gemm(M, N, K, epi, A, W, bias, resid)
        except Exception as e:
            print(f"GEMM M={M} N={N} K={K} epi={epi}: FAILED {e}", flush=True)
            return
        dt = time.time() - t
        Ms = min(M, 2048)
        ref = A[:Ms].astype(np.float32) @ W.astype(np.float32).T + bias
        if epi == 2:
            ref = ref + resid[:Ms]
        if epi == 0:
            ref = ref.astype(np.float16).astype(np.float32)
        if epi == 1:
            ref = rs.gelu_table(ref)
        d = np.abs(out[:Ms] - ref)
        bad = np.argwhere(d > 2e-3 * (1 + np.abs(ref)))
        print(f"GEMM M={M} N={N} K={K} epi={epi}: max|d|={d.max():.3e} mean|d|={d.mean():.3e} max|ref|={np.abs(ref).max():.2f} "
              f"nbad={len(bad)} first_bad={bad[:4].tolist()} ({dt:.2f}s)", flush=True)
        if len(bad):
            r, c = bad[0]
            print("   out", out[r, c:c + 8], "\n   ref", ref[r, c:c + 8], flush=True)
            # structure of the error: which rows / cols are wrong
            print("   bad rows (first 16):", np.unique(bad[:, 0])[:16].tolist(), " bad cols (first 16):", np.unique(bad[:, 1])[:16].tolist(), flush=True)


def _taps(cfg, B, layer):
    from tests.util import pkg, gf, model_path
    from oracle import restatement as rs
    eng = pkg.engine
    path = model_path(cfg, "f16")
    vf = gf.read(path)
    om = rs.OracleModel(vf, gf.tensor_specs)
    rs.set_threads(8)
    imgs = gf.synthetic_images(B, vf.img_size, seed=7)
    m = eng.vit_model_load(path, 0, max(B, 4))
    probs, logits, taps = eng.vit_predict_debug(m, imgs, layer)
    names = list(eng.TAP_SHAPES)
    for b in range(B):
        p_o, l_o, t_o = om.forward(imgs[b], layer, tuple(names))
        for n in names:
            a, r = taps[n][b], t_o[n]
            d = np.abs(a - r)
            print(f"  {cfg} img{b} L{layer} {n:9s} max|d|={d.max():.3e} rel={d.max() / (np.abs(r).max() + 1e-30):.3e} max|ref|={np.abs(r).max():.3f}", flush=True)
        d = np.abs(logits[b] - l_o)
        print(f"  {cfg} img{b} LOGITS rel={d.max() / np.abs(l_o).max():.3e} top5 gpu={np.argsort(-logits[b])[:5].tolist()} ref={np.argsort(-l_o)[:5].tolist()} "
              f"|dp|max={np.abs(probs[b] - p_o).max():.2e}", flush=True)


def stage_taps_micro():
    _taps("micro", 2, 0)
    _taps("micro", 1, 1)


def stage_taps_micro14():
    _taps("micro14", 2, 0)


def stage_taps_tiny():
    _taps("tiny", 2, 0)
    _taps("tiny", 1, 11)


def stage_golden():
    from tests.util import pkg, gf, model_path
    eng = pkg.engine
    for cfg in ("micro", "micro14", "tiny", "base"):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"{cfg}_f16.npz"))
        path = model_path(cfg, "f16")
        m = eng.vit_model_load(path, 0, 8)
        imgs = gf.synthetic_images(int(g["n_images"]), m.img_size, seed=int(g["image_seed"]))
        probs, idx, val, logits = eng.vit_predict(m, imgs, 5, want_logits=True)
        for b in range(imgs.shape[0]):
            rel = np.abs(logits[b] - g["logits"][b]).max() / np.abs(g["logits"][b]).max()
            print(f"  golden {cfg} img{b}: logit rel err {rel:.3e}  top5 gpu={idx[b].tolist()} ref={np.argsort(-g['logits'][b])[:5].tolist()} "
                  f"|dp|max={np.abs(probs[b] - g['probs'][b]).max():.2e}", flush=True)
        m.close()


def stage_speed():
    import ctypes as C
    import torch
    from tests.util import pkg, gf, model_path
    eng = pkg.engine
    path = model_path("base", "f16")
    B = 256
    m = eng.vit_model_load(path, 0, B)
    imgs = torch.from_numpy(gf.synthetic_images(B, 224, seed=1)).cuda()
    probs = torch.empty(B, 1000, device="cuda")
    idx = torch.empty(B, 5, dtype=torch.int32, device="cuda")
    val = torch.empty(B, 5, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    s = ts.cuda_stream
    L = eng.lib()

    def run():
        rc = L.vitb200_forward_device(m.handle, imgs.data_ptr(), B, probs.data_ptr(), None, idx.data_ptr(), val.data_ptr(), 5, s)
        assert rc == 0, L.vitb200_last_error()
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"  base f16 B={B}: {ms:.2f} ms/forward = {B / ms * 1e3:.0f} img/s, launches={m.last_launch_count()}", flush=True)


STAGES = {"gemm": stage_gemm, "taps_micro": stage_taps_micro, "taps_micro14": stage_taps_micro14, "taps_tiny": stage_taps_tiny,
          "golden": stage_golden, "speed": stage_speed}

if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--run":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    for name in (sys.argv[1:] or list(STAGES)):
        print(f"=== stage {name}", flush=True)
        t = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", name], timeout=420,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            lines = [l for l in r.stdout.splitlines() if not l.startswith("vit_model_load") and "ggml ctx size" not in l]
            print("\n".join(lines[-120:]))
            print(f"=== stage {name} exit={r.returncode} ({time.time() - t:.1f}s)", flush=True)
        except subprocess.TimeoutExpired as ex:
            print((ex.stdout or "")[-4000:] if isinstance(ex.stdout, str) else ex.stdout)
            print(f"=== stage {name} TIMEOUT", flush=True)
