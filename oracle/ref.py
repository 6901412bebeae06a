"""ctypes wrapper around oracle/_ref/libvitref.so -- the UNMODIFIED reference, compiled from
/root/reference by oracle/Makefile.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libvitref.so")
QUANTIZE_BIN = os.path.join(_HERE, "_ref", "quantize")
VIT_REF_BIN = os.path.join(_HERE, "_ref", "vit_ref")

_lib = None


class _quiet:
    """Redirect the process-level stdout/stderr (fd 1, 2) to /dev/null: the reference prints hparams, progress dots and
    preprocess chatter from C (vit.cpp:310-352, 226-231), which must not pollute bench.py's one-line JSON."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        sys.stderr.flush()
        self.null = os.open(os.devnull, os.O_WRONLY)
        self.saved = (os.dup(1), os.dup(2))
        os.dup2(self.null, 1)
        os.dup2(self.null, 2)

    def __exit__(self, *a):
        os.dup2(self.saved[0], 1)
        os.dup2(self.saved[1], 2)
        for fd in (self.saved[0], self.saved[1], self.null):
            os.close(fd)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.vitref_load.restype = C.c_void_p
        L.vitref_load.argtypes = [C.c_char_p]
        L.vitref_hparams.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.vitref_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.vitref_preprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vitref_load_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_int64]
        L.vitref_free.argtypes = [C.c_void_p]
        L.vitref_round_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        _lib = L
    return _lib


class RefModel:
    """Reference vit_model + vit_state (reference vit.h:72-89), loaded by the reference loader."""

    def __init__(self, path: str):
        with _quiet():
            self._h = lib().vitref_load(path.encode())
        if not self._h:
            raise RuntimeError(f"reference vit_model_load failed for {path}")
        hp = (C.c_int32 * 8)()
        lib().vitref_hparams(self._h, hp)
        (self.hidden, self.layers, self.heads, self.classes, self.patch, self.img, self.ftype) = list(hp)[:7]

    def predict(self, img_hwc: np.ndarray, n_threads: int = 4):
        """One reference vit_predict (vit.cpp:1004).  Returns (probs, logits) float32[num_classes]."""
        img = np.ascontiguousarray(img_hwc, dtype=np.float32)
        assert img.shape == (self.img, self.img, 3), img.shape
        probs = np.empty(self.classes, np.float32)
        logits = np.empty(self.classes, np.float32)
        with _quiet():
            rc = lib().vitref_predict(self._h, img.ctypes.data, n_threads, probs.ctypes.data, logits.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"reference vit_predict returned {rc}")
        return probs, logits

    def predict_batch(self, imgs: np.ndarray, n_threads: int = 4):
        ps, ls = [], []
        for i in range(imgs.shape[0]):
            p, l = self.predict(imgs[i], n_threads)
            ps.append(p)
            ls.append(l)
        return np.stack(ps), np.stack(ls)

    def preprocess(self, rgb_u8: np.ndarray, bilinear: bool = False) -> np.ndarray:
        rgb = np.ascontiguousarray(rgb_u8, dtype=np.uint8)
        ny, nx, _ = rgb.shape
        out = np.empty((self.img, self.img, 3), np.float32)
        with _quiet():
            rc = lib().vitref_preprocess(self._h, rgb.ctypes.data, nx, ny, int(bilinear), out.ctypes.data)
        if rc != 0:
            raise RuntimeError("reference preprocess failed")
        return out

    def close(self):
        if self._h:
            lib().vitref_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_image(path: str) -> np.ndarray:
    nx, ny = C.c_int(), C.c_int()
    if lib().vitref_load_image(path.encode(), C.byref(nx), C.byref(ny), None, 0) != 0:
        raise RuntimeError(f"stb_image failed on {path}")
    out = np.empty((ny.value, nx.value, 3), np.uint8)
    lib().vitref_load_image(path.encode(), C.byref(nx), C.byref(ny), out.ctypes.data, out.nbytes)
    return out


def round_f16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    lib().vitref_round_f16(x.ctypes.data, y.ctypes.data, x.size)
    return y


# ------------------------------------------------------------------------------------------------
# The reference's ViTSTR extension (extensions/vitstr.cpp), its own shared object (oracle/Makefile, oracle/vitstr_harness.cpp)
VITSTR_LIB_PATH = os.path.join(_HERE, "_ref", "libvitstrref.so")
_vitstr_lib = None


def vitstr_available() -> bool:
    return os.path.exists(VITSTR_LIB_PATH)


def vitstr_lib():
    global _vitstr_lib
    if _vitstr_lib is None:
        L = C.CDLL(VITSTR_LIB_PATH)
        L.vitstrref_load.restype = C.c_void_p
        L.vitstrref_load.argtypes = [C.c_char_p]
        L.vitstrref_hparams.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.vitstrref_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.vitstrref_free.argtypes = [C.c_void_p]
        _vitstr_lib = L
    return _vitstr_lib


class VitstrRefModel:
    """The extension's vit_model + vit_state, loaded by its own loader (vitstr.cpp:276-683)."""

    SEQ_LEN = 25  # vitstr.cpp:865

    def __init__(self, path: str):
        with _quiet():
            self._h = vitstr_lib().vitstrref_load(path.encode())
        if not self._h:
            raise RuntimeError(f"reference ViTSTR loader rejected {path}")
        hp = (C.c_int32 * 7)()
        vitstr_lib().vitstrref_hparams(self._h, hp)
        self.hidden, self.layers, self.heads, self.classes, self.patch, self.img, self.ftype = list(hp)

    def predict(self, img_gray: np.ndarray, n_threads: int = 4):
        """One reference forward on a pre-processed grayscale image [S,S] f32.  Returns (probs, logits) float32[25, classes]."""
        img = np.ascontiguousarray(img_gray, dtype=np.float32)
        assert img.size == self.img * self.img, img.shape
        probs = np.empty((self.SEQ_LEN, self.classes), np.float32)
        logits = np.empty((self.SEQ_LEN, self.classes), np.float32)
        with _quiet():
            rc = vitstr_lib().vitstrref_predict(self._h, img.ctypes.data, n_threads, probs.ctypes.data, logits.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"reference ViTSTR vit_predict returned {rc}")
        return probs, logits

    def predict_batch(self, imgs: np.ndarray, n_threads: int = 4):
        ps, ls = zip(*(self.predict(imgs[i], n_threads) for i in range(imgs.shape[0])))
        return np.stack(ps), np.stack(ls)

    def close(self):
        if self._h:
            vitstr_lib().vitstrref_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
