"""Helpers for the GPU parity tests: padded-NHWC packing and cached plan files."""
import os

import numpy as np

import adas_b200  # noqa: F401
from adas_b200 import plan



def to_padded(x_nchw: np.ndarray, C: int) -> np.ndarray:
    """[B,c,H,W] float -> [B*(H+2)*(W+2), C] fp16 with zero halo / zero extra channels."""
    B, c, H, W = x_nchw.shape
    out = np.zeros((B, H + 2, W + 2, C), np.float16)
    out[:, 1:-1, 1:-1, :c] = x_nchw.transpose(0, 2, 3, 1).astype(np.float16)
    return out.reshape(-1, C)


def from_padded(buf: np.ndarray, B: int, H: int, W: int, coff: int, c: int) -> np.ndarray:
    """[B*(H+2)*(W+2), ld] -> [B,c,H,W] float32 interior."""
    v = buf.reshape(B, H + 2, W + 2, -1)[:, 1:-1, 1:-1, coff:coff + c]
    return v.astype(np.float32).transpose(0, 3, 1, 2)


def halo_is_zero(buf: np.ndarray, B: int, H: int, W: int) -> bool:
    v = buf.reshape(B, H + 2, W + 2, -1).astype(np.float32)
    return not (v[:, 0].any() or v[:, -1].any() or v[:, :, 0].any() or v[:, :, -1].any())


def cached_plan(kind: str, seed: int = 0, **kw):
    """Build (once per process tree) the synthetic plan + return (path, Weights-like state_dict)."""
    CACHE = plan.cache_dir()
    import zlib
    prof = zlib.crc32(repr((plan.SYNTH_PROFILES.get("ufldv2" if kind == "ufldv1" else kind), plan.PLAN_VERSION)).encode()) & 0xffff      # a changed operating point is a new plan
    tag = kind + "_" + "_".join(f"{k}{v}" for k, v in sorted(kw.items())) + f"_s{seed}_{prof:04x}"
    path = os.path.join(CACHE, tag + ".b200w")
    variant = kw.get("scale", kw.get("backbone"))             # calibrated BatchNorm statistics exist for the tested variants
    W = plan.synth_weights("ufldv2" if kind == "ufldv1" else kind, seed, variant=variant)
    if kind == "yolov8":
        pb = plan.build_yolov8(W, **kw)
    elif kind == "yolov5":
        pb = plan.build_yolov5(W, **kw)
    elif kind == "ufldv1":
        pb = plan.build_ufldv1(W, **kw)
    else:
        pb = plan.build_ufldv2(W, **kw)
    if not os.path.isfile(path):
        pb.write(path + ".tmp")
        os.replace(path + ".tmp", path)
    return path, W.state_dict, pb
