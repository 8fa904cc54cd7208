"""Shared helpers of the parity tests: cached synthetic windows and a uniform way to drive the product
(libsosba.so, CUDA) and the oracle (liborc_parity.so, CPU) through the same ctypes binding."""
import hashlib
import os
import pickle

import numpy as np

from sosba_loader import load_package

load_package()
from sos_slam_b200 import binding, problem, synth  # noqa: E402

_CACHE = {}


def scene(**kw):
    key = tuple(sorted(kw.items()))
    if key in _CACHE:
        return _CACHE[key]
    tag = hashlib.sha1(repr(key).encode()).hexdigest()[:16]
    path = os.path.join("/tmp", f"sosba_scene_{tag}.pkl")
    sc = None
    if os.path.exists(path):
        try:
            with open(path, "rb") as f:
                sc = pickle.load(f)
        except Exception:
            sc = None
    if sc is None:
        sc = synth.make_scene(**kw)
        try:
            with open(path + ".tmp", "wb") as f:
                pickle.dump(sc, f)
            os.replace(path + ".tmp", path)
        except Exception:
            pass
    _CACHE[key] = sc
    return sc


SMALL = dict(w=320, h=240, nf=5, n_points=400, seed=3)
CONFIG_B = dict(w=640, h=480, nf=8, n_points=2000, seed=1234)       # BASELINE.json configs[1]
KITTI = dict(w=1232, h=368, nf=12, n_points=4000, seed=5, forward_motion=True)  # configs[3] shape
# configs[2] shape: EuRoC 752x480 with the intrinsics of tests/EuRoC/camera0.txt (relative K of Undistort::readFromFile, pre-rectified)
EUROC = dict(w=752, h=480, nf=8, n_points=2000, seed=11, fx=0.6099 * 752, fy=0.9527 * 480, cx=0.4890 * 752 - 0.5, cy=0.5185 * 480 - 0.5)
TUMVI = dict(w=512, h=512, nf=8, n_points=2000, seed=12)   # configs[4] shape: 512x512, 4 pyramid levels


def open_handle(lib, sc, threads=1):
    cfg = lib.config_default(sc.w, sc.h)
    cfg.num_threads = threads
    cfg.max_frames = sc.nf + 2
    h = binding.Handle(lib, cfg)
    for i, img in enumerate(sc.images):
        h.frame_make_images(i, img)
    return h


def upload(h, sc, HM=None, bM=None, pts=None, res=None, calib_delta=(0.0, 0.0, 0.0, 0.0)):
    val, val0 = problem.calib_of(sc, calib_delta)
    P, keep = h.make_problem(problem.frames_of(sc), val, val0, pts or problem.points_of(sc), res or problem.residuals_of(sc), HM, bM)
    h.ba_upload(P)
    return P, keep


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def trace_points(h, sc, case):
    """ImmaturePoint construction of every candidate of `case` (synth.trace_case) through handle `h` -> one SoA dict."""
    parts = []
    for hst in range(sc.nf):
        m = case["host"] == hst
        if m.any():
            parts.append(h.immature_init(hst, case["u"][m], case["v"][m]))
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
