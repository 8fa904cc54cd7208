"""Generates tests/golden/ba_tiny.npz: inputs and outputs of the ORACLE (parity build) on a tiny seeded window.

The reference ships no golden vectors for this path and cannot be built in this image, so these vectors
pin the oracle and the CUDA path against drift; they are not reference outputs.  Re-run after an intentional change of
the parity definition:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sosba_loader import load_package  # noqa: E402

load_package()
from sos_slam_b200 import binding, problem, synth  # noqa: E402

TINY = dict(w=160, h=120, nf=4, n_points=120, seed=5)


def run(lib, sc):
    cfg = lib.config_default(sc.w, sc.h)
    cfg.max_frames = sc.nf + 2
    h = binding.Handle(lib, cfg)
    for i, img in enumerate(sc.images):
        h.frame_make_images(i, img)
    val, val0 = problem.calib_of(sc, (1e-4, -1e-4, 2e-4, 1e-4))
    P, keep = h.make_problem(problem.frames_of(sc), val, val0, problem.points_of(sc), problem.residuals_of(sc))
    h.ba_upload(P)
    out = {}
    lvl1 = h.frame_get_level(0, 1)
    out["pyr1_dI"], out["pyr1_abs"] = lvl1
    h.reset_oob()
    lo = h.linearize_all(False)
    out["lin_counts"] = np.array([lo["n_in"], lo["n_oob"], lo["n_outlier"]], np.int32)
    out["lin_energy"] = np.float64(lo["energy"])
    out["lin_th"] = np.float32(lo["new_frame_energy_th"])
    st = h.get_state()
    out["new_state"], out["new_energy"] = st["new_state"], st["new_energy"]
    out["J"] = h.get_jacobians(False)
    h.apply_res()
    acc = h.accumulate()
    for k in ("HA", "bA", "HL", "bL", "Hsc", "bsc"):
        out[k] = acc[k]
    out["resIn"] = np.array([acc["resInA"], acc["resInL"]], np.int32)
    x, Hf, bf = h.solve_system()
    out["x"] = x
    out["step"] = h.resubstitute(x)
    P2, keep2 = h.make_problem(problem.frames_of(sc), val, val0, problem.points_of(sc), problem.residuals_of(sc))
    o = h.optimize(P2, 6)
    res = h.problem_result(P2, keep2)
    out["opt_iterations"] = np.int32(o["iterations"])
    out["opt_energy"] = np.array([o["energy_initial"], o["energy_final"]])
    out["opt_state"], out["opt_idepth"] = res["state"], res["idepth"]
    out["opt_res_state"] = h.get_state()["state"]
    h.close()
    return out


def main():
    sc = synth.make_scene(**TINY)
    orc = binding.Lib(os.path.join(ROOT, "oracle", "_build", "liborc_parity.so"), "orc")
    out = run(orc, sc)
    inputs = {"in_images": np.stack(sc.images).astype(np.float32), "in_K": sc.K, "in_evalPT": sc.evalPT, "in_state": sc.state,
              "in_state_zero": sc.state_zero, "in_ab_exposure": sc.ab_exposure, "in_frame_id": sc.frame_id, "in_pt_host": sc.pt_host,
              "in_pt_u": sc.pt_u, "in_pt_v": sc.pt_v, "in_pt_idepth": sc.pt_idepth, "in_pt_color": sc.pt_color, "in_pt_weights": sc.pt_weights,
              "in_res_point": sc.res_point, "in_res_target": sc.res_target}
    path = os.path.join(ROOT, "tests", "golden", "ba_tiny.npz")
    np.savez_compressed(path, **inputs, **{"out_" + k: v for k, v in out.items()})
    print(path, os.path.getsize(path) // 1024, "KiB;", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
