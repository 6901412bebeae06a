"""Offline converters for ViT checkpoints -- the timm-free counterpart of the reference's convert-pth-to-ggml.py and quantize.cpp.

    python -m convert <checkpoint.pth|.safetensors> <out.gguf> [--ftype 0|1] [--heads H] [--labels labels.json] [--container legacy|gguf]
    python -m convert --quantize <in.gguf> <out.gguf> q8_0

* `state_dict_to_model_file` follows reference convert-pth-to-ggml.py:105-158 step by step (header, id2label table, tensors in
  state_dict order, `norm_pre*` skipped, 1-D tensors + pos_embed + cls_token kept f32, the conv bias reshaped to (1, D, 1, 1)),
  but takes the weights from a plain `state_dict` -- a torch checkpoint (`torch.load`, needs torch) or a .safetensors file (read
  here with numpy only) -- instead of instantiating the model through timm.  Hyper-parameters are inferred from the tensor shapes
  (timm `VisionTransformer` names, reference vit.cpp:518-579); the head count is not recoverable from shapes and defaults to
  hidden / 64 (true for every timm ViT the reference lists).
* `quantize_model_file` restates reference quantize.cpp for q8_0 (every 2-D `*weight` tensor, quantize_row_q8_0_reference,
  ggml-quants.c): byte-identical to the reference binary's output (pinned by the CPU test-suite).

Host-side tooling: nothing here touches the GPU or the test infrastructure.
"""
from __future__ import annotations

import json
import re
import struct
import sys
from typing import Dict, Optional

import numpy as np

try:  # package import (vit_cpp_b200.convert) or plain script next to ggml_file.py
    from . import ggml_file as gf
except ImportError:  # pragma: no cover
    import ggml_file as gf

_ST_DTYPES = {"F32": np.float32, "F16": np.float16, "F64": np.float64, "I64": np.int64, "I32": np.int32, "U8": np.uint8, "BF16": "bf16"}


def read_safetensors(path: str) -> Dict[str, np.ndarray]:
    """Minimal safetensors reader (8-byte little-endian header length, JSON header, raw tensor bytes); BF16 is widened to f32."""
    with open(path, "rb") as f:
        n = struct.unpack("<Q", f.read(8))[0]
        if n > (1 << 28):
            raise ValueError("safetensors header is implausibly large")
        hdr = json.loads(f.read(n).decode("utf-8"))
        data = f.read()
    out = {}
    for name, info in hdr.items():
        if name == "__metadata__":
            continue
        b, e = info["data_offsets"]
        dt = _ST_DTYPES.get(info["dtype"])
        if dt is None:
            raise ValueError(f"tensor '{name}': unsupported safetensors dtype {info['dtype']}")
        if dt == "bf16":
            arr = gf.bf16_bits_to_f32(np.frombuffer(data, np.uint16, (e - b) // 2, b))
        else:
            arr = np.frombuffer(data, dt, (e - b) // np.dtype(dt).itemsize, b)
        out[name] = arr.reshape(info["shape"])
    return out


def write_safetensors(path: str, tensors: Dict[str, np.ndarray]) -> None:
    """Writer for tests (f32 / f16 tensors, insertion order)."""
    hdr, off, blobs = {}, 0, []
    for name, a in tensors.items():
        a = np.ascontiguousarray(a)
        dt = {np.dtype(np.float32): "F32", np.dtype(np.float16): "F16"}[a.dtype]
        hdr[name] = {"dtype": dt, "shape": list(a.shape), "data_offsets": [off, off + a.nbytes]}
        off += a.nbytes
        blobs.append(a.tobytes())
    js = json.dumps(hdr, separators=(",", ":")).encode()
    js += b" " * ((8 - len(js) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(js)))
        f.write(js)
        for b in blobs:
            f.write(b)


def load_state_dict(path: str) -> Dict[str, np.ndarray]:
    if path.endswith(".safetensors"):
        return read_safetensors(path)
    import torch  # checkpoints written by torch.save
    sd = torch.load(path, map_location="cpu", weights_only=True)
    for key in ("state_dict", "model"):
        if isinstance(sd, dict) and key in sd and isinstance(sd[key], dict):
            sd = sd[key]
    return {k: v.detach().to(torch.float32).numpy() for k, v in sd.items() if hasattr(v, "detach")}


def infer_hparams(sd: Dict[str, np.ndarray], heads: Optional[int] = None):
    """(hidden, layers, heads, classes, patch, img, in_chans) from timm VisionTransformer tensor shapes."""
    pw = sd["patch_embed.proj.weight"]                       # (D, C, P, P)
    hidden, in_chans, patch = int(pw.shape[0]), int(pw.shape[1]), int(pw.shape[2])
    n_tok = int(sd["pos_embed"].shape[1])                    # (1, N, D), N = (img / patch)^2 + 1
    grid = int(round((n_tok - 1) ** 0.5))
    if grid * grid + 1 != n_tok:
        raise ValueError(f"pos_embed has {n_tok} tokens: not a square patch grid plus one class token")
    layers = 1 + max(int(m.group(1)) for k in sd for m in [re.match(r"blocks\.(\d+)\.", k)] if m)
    classes = int(sd["head.weight"].shape[0])
    heads = heads or hidden // 64
    if hidden % heads:
        raise ValueError(f"hidden size {hidden} is not divisible by {heads} heads")
    return hidden, layers, heads, classes, patch, grid * patch, in_chans


def state_dict_to_model_file(sd: Dict[str, np.ndarray], out_path: str, ftype: int = 1, heads: Optional[int] = None,
                             id2label: Optional[Dict[int, str]] = None, container: str = "legacy") -> None:
    assert ftype in (0, 1)
    hidden, layers, heads, classes, patch, img, in_chans = infer_hparams(sd, heads)
    id2label = id2label or {i: f"LABEL_{i}" for i in range(classes)}   # convert-pth-to-ggml.py:95-96 fallback names
    vf = gf.VitFile(hidden, layers, heads, classes, patch, img, ftype, dict(id2label))
    for k, v in sd.items():
        if k.startswith("norm_pre"):                                     # convert-pth-to-ggml.py:127-130
            continue
        a = np.asarray(v)
        if k == "patch_embed.proj.bias":
            a = a.reshape(1, a.shape[0], 1, 1)                           # convert-pth-to-ggml.py:150-151
        # convert-pth-to-ggml.py:143-147; the loader additionally hard-codes the patch kernel as F16 (vit.cpp:515), so an f32
        # file must still carry it as f16 (the reference's own --ftype 0 output does not load)
        f16 = (ftype == 1 and np.asarray(v).ndim != 1 and k not in ("pos_embed", "cls_token")) or k == "patch_embed.proj.weight"
        vf.tensors[k] = a.astype(np.float16 if f16 else np.float32)
        vf.tensor_ftype[k] = 1 if f16 else 0
    expect = 4 + 12 * layers + 4                                         # vit.cpp:697
    if len(vf.tensors) != expect:
        raise ValueError(f"state_dict has {len(vf.tensors)} tensors after filtering, a {layers}-layer ViT needs {expect}")
    if container == "gguf":
        gf.write_gguf(out_path, vf, "keep")
    else:
        write_legacy(out_path, vf)


def write_legacy(path: str, vf) -> None:
    """A VitFile (f32 / f16 / raw q8_0 tensors) -> legacy-ggml container, tensor order as stored."""
    with open(path, "wb") as f:
        f.write(struct.pack("i", gf.GGML_FILE_MAGIC))
        for v in (vf.hidden_size, vf.num_hidden_layers, vf.num_attention_heads, vf.num_classes, vf.patch_size, vf.img_size, vf.ftype):
            f.write(struct.pack("i", v))
        f.write(struct.pack("i", len(vf.id2label)))
        for key, value in vf.id2label.items():
            s = value.encode("utf-8")
            f.write(struct.pack("ii", key, len(s)))
            f.write(s)
        for name, arr in vf.tensors.items():
            if name.endswith(".q8_0_raw"):
                continue
            ft = vf.tensor_ftype[name]
            shape = arr.shape
            if name == "patch_embed.proj.bias" and arr.ndim == 1:
                shape = (1, arr.shape[0], 1, 1)                          # convert-pth-to-ggml.py:150-151
            nm = name.encode("utf-8")
            f.write(struct.pack("iii", len(shape), len(nm), ft))
            for d in reversed(shape):
                f.write(struct.pack("i", d))
            f.write(nm)
            if ft == 8:
                f.write(vf.tensors[name + ".q8_0_raw"].tobytes())
            else:
                np.ascontiguousarray(arr, np.float16 if ft == 1 else np.float32).tofile(f)


def quantize_q8_0_reference(x: np.ndarray) -> np.ndarray:
    """quantize_row_q8_0_reference (ggml-quants.c): per 32 values d = amax / 127 (stored f16), q = roundf(x * (1/d)), half away from
    zero.  x: float32 [n], n % 32 == 0.  Returns the raw block bytes (34 per block: f16 d, int8 q[32])."""
    x = np.ascontiguousarray(x, np.float32).reshape(-1, 32)
    amax = np.abs(x).max(axis=1)
    d = (amax / np.float32(127.0)).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
    x0 = (x * inv[:, None]).astype(np.float32)
    q = (np.sign(x0) * np.floor(np.abs(x0) + np.float32(0.5))).astype(np.int8)   # roundf
    out = np.empty((x.shape[0], 34), np.uint8)
    out[:, :2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = q.view(np.uint8)
    return out.reshape(-1)


def quantize_model_file(src: str, dst: str, fmt: str = "q8_0") -> None:
    """reference quantize.cpp:26-330 for itype 8: every 2-D tensor whose name ends in `weight` is quantised row-wise, everything
    else is copied; the header's ftype becomes the quantisation type."""
    if fmt != "q8_0":
        raise ValueError("only q8_0 is restated here; the other block formats come from the reference's quantize binary")
    vf = gf.read(src)
    out = gf.VitFile(vf.hidden_size, vf.num_hidden_layers, vf.num_attention_heads, vf.num_classes, vf.patch_size, vf.img_size, 8,
                     dict(sorted(vf.id2label.items())))                  # quantize.cpp re-writes the labels from a std::map: key order
    for name, arr in vf.tensors.items():
        ft = vf.tensor_ftype[name]
        if name.endswith("weight") and arr.ndim == 2:                    # quantize.cpp:196-214
            out.tensors[name + ".q8_0_raw"] = quantize_q8_0_reference(arr.astype(np.float32))
            out.tensors[name] = arr
            out.tensor_ftype[name] = 8
        else:
            out.tensors[name] = arr
            out.tensor_ftype[name] = ft
    write_legacy(dst, out)


def main(argv):
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("fmt", nargs="?", default="q8_0")
    ap.add_argument("--quantize", action="store_true")
    ap.add_argument("--ftype", type=int, choices=[0, 1], default=1)
    ap.add_argument("--heads", type=int, default=None)
    ap.add_argument("--labels", default=None, help="JSON file {class id: label}")
    ap.add_argument("--container", choices=["legacy", "gguf"], default="legacy")
    a = ap.parse_args(argv)
    if a.quantize:
        quantize_model_file(a.src, a.dst, a.fmt)
        return
    labels = {int(k): v for k, v in json.load(open(a.labels)).items()} if a.labels else None
    state_dict_to_model_file(load_state_dict(a.src), a.dst, a.ftype, a.heads, labels, a.container)


if __name__ == "__main__":
    main(sys.argv[1:])
