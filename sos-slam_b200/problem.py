"""Scene -> the flat arrays the C ABI takes (frames / points / residuals dictionaries)."""
from __future__ import annotations

import numpy as np

from .synth import Scene, SCALE_F, SCALE_C


def frames_of(scene: Scene, frame_energy_th: float = 8 * 8 * 8) -> list:
    return [{"evalPT": scene.evalPT[i][:3, :4], "state": scene.state[i], "state_zero": scene.state_zero[i],
             "ab_exposure": float(scene.ab_exposure[i]), "frame_energy_th": frame_energy_th,
             "frame_id": int(scene.frame_id[i]), "slot": i} for i in range(scene.nf)]


def calib_of(scene: Scene, delta=(0.0, 0.0, 0.0, 0.0)):
    """CalibHessian::value (unscaled) and value_zero; `delta` = value - value_zero."""
    fx, fy, cx, cy = scene.K
    value = np.array([fx / SCALE_F, fy / SCALE_F, cx / SCALE_C, cy / SCALE_C], np.float64)
    return value, value - np.asarray(delta, np.float64)


def points_of(scene: Scene, prior=None, idepth_zero=None) -> dict:
    P = scene.n_points
    iz = scene.pt_idepth if idepth_zero is None else idepth_zero
    return {"u": scene.pt_u, "v": scene.pt_v, "idepth": scene.pt_idepth, "idepth_zero": iz,
            "color": scene.pt_color, "weights": scene.pt_weights, "host": scene.pt_host,
            "priorF": np.zeros(P, np.float32) if prior is None else prior,
            "deltaF": (scene.pt_idepth - iz).astype(np.float32)}


def residuals_of(scene: Scene) -> dict:
    R = scene.n_residuals
    return {"point": scene.res_point, "target": scene.res_target, "state": np.zeros(R, np.uint8),
            "is_linearized": np.zeros(R, np.uint8), "is_active": np.zeros(R, np.uint8), "is_new": np.ones(R, np.uint8)}


def shard_points(res_point: np.ndarray, n_points: int, world: int) -> list:
    """Block-partition points over `world` ranks balanced by residual count.
    Returns per rank (p0, p1): the half-open point range it owns."""
    counts = np.bincount(res_point, minlength=n_points).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        bounds.append(int(np.searchsorted(csum, target, side="left")))
    bounds.append(n_points)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_scene_arrays(pts: dict, res: dict, p0: int, p1: int):
    """The sub-problem of points [p0,p1) with their residuals, re-indexed from 0."""
    sp = {k: (v[p0:p1] if k not in ("color", "weights") else v[p0:p1]) for k, v in pts.items()}
    m = (res["point"] >= p0) & (res["point"] < p1)
    sr = {k: v[m] for k, v in res.items()}
    sr["point"] = (sr["point"] - p0).astype(np.int32)
    return sp, sr
