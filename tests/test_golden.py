"""tests/golden/ba_tiny.npz (made by tools/make_golden.py from the oracle's parity build): the oracle must reproduce it
bit for bit (CPU), the CUDA path must match it to the parity definition of DESIGN.md §2 (GPU).  The window is rebuilt from
the arrays stored in the fixture, not from the generator."""
import os

import numpy as np
import pytest

from _scenes import relerr
from sosba_loader import load_package

load_package()
from sos_slam_b200 import binding, synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "ba_tiny.npz"))


def _run(lib):
    imgs = G["in_images"]
    nf, hgt, wid = imgs.shape
    cfg = lib.config_default(wid, hgt)
    cfg.max_frames = nf + 2
    h = binding.Handle(lib, cfg)
    for i in range(nf):
        h.frame_make_images(i, imgs[i])
    K = G["in_K"]
    val = np.array([K[0] / 50.0, K[1] / 50.0, K[2] / 50.0, K[3] / 50.0])
    val0 = val - np.array([1e-4, -1e-4, 2e-4, 1e-4])
    frames = [{"evalPT": G["in_evalPT"][i][:3, :4], "state": G["in_state"][i], "state_zero": G["in_state_zero"][i],
               "ab_exposure": float(G["in_ab_exposure"][i]), "frame_energy_th": 512.0, "frame_id": int(G["in_frame_id"][i]), "slot": i} for i in range(nf)]
    P = len(G["in_pt_u"])
    R = len(G["in_res_point"])
    pts = {"u": G["in_pt_u"], "v": G["in_pt_v"], "idepth": G["in_pt_idepth"], "idepth_zero": G["in_pt_idepth"], "color": G["in_pt_color"],
           "weights": G["in_pt_weights"], "host": G["in_pt_host"], "priorF": np.zeros(P, np.float32), "deltaF": np.zeros(P, np.float32)}
    res = {"point": G["in_res_point"], "target": G["in_res_target"], "state": np.zeros(R, np.uint8), "is_linearized": np.zeros(R, np.uint8),
           "is_active": np.zeros(R, np.uint8), "is_new": np.ones(R, np.uint8)}
    Pr, keep = h.make_problem(frames, val, val0, pts, res)
    h.ba_upload(Pr)
    out = {}
    out["pyr1_dI"], out["pyr1_abs"] = h.frame_get_level(0, 1)
    h.reset_oob()
    lo = h.linearize_all(False)
    out["lin_counts"] = np.array([lo["n_in"], lo["n_oob"], lo["n_outlier"]], np.int32)
    out["lin_energy"], out["lin_th"] = lo["energy"], np.float32(lo["new_frame_energy_th"])
    st = h.get_state()
    out["new_state"], out["new_energy"] = st["new_state"], st["new_energy"]
    out["J"] = h.get_jacobians(False)
    h.apply_res()
    acc = h.accumulate()
    out.update({k: acc[k] for k in ("HA", "bA", "HL", "bL", "Hsc", "bsc")})
    out["resIn"] = np.array([acc["resInA"], acc["resInL"]], np.int32)
    x, Hf, bf = h.solve_system()
    out["x"], out["Hf"] = x, Hf
    out["step"] = h.resubstitute(G["out_x"])
    P2, keep2 = h.make_problem(frames, val, val0, pts, res)
    o = h.optimize(P2, 6)
    r = h.problem_result(P2, keep2)
    out["opt_iterations"], out["opt_energy"] = o["iterations"], np.array([o["energy_initial"], o["energy_final"]])
    out["opt_state"], out["opt_idepth"], out["opt_res_state"] = r["state"], r["idepth"], h.get_state()["state"]
    h.close()
    return out


def test_oracle_reproduces_golden(orc):
    o = _run(orc)
    for k in ("pyr1_dI", "pyr1_abs", "lin_counts", "lin_th", "new_state", "new_energy", "J", "HA", "bA", "HL", "bL", "Hsc", "bsc", "resIn", "x",
              "step", "opt_state", "opt_idepth", "opt_res_state", "opt_energy"):
        assert np.array_equal(np.asarray(o[k]), G["out_" + k]), k
    assert o["lin_energy"] == float(G["out_lin_energy"]) and o["opt_iterations"] == int(G["out_opt_iterations"])


@pytest.mark.gpu
def test_cuda_matches_golden(gpu):
    o = _run(gpu)
    for k in ("pyr1_dI", "pyr1_abs", "lin_counts", "lin_th", "new_state", "resIn"):
        assert np.array_equal(np.asarray(o[k]), G["out_" + k]), k
    live = G["out_new_state"] != 1
    assert np.array_equal(o["new_energy"][live], G["out_new_energy"][live])
    assert np.array_equal(o["J"][live], G["out_J"][live])
    assert abs(o["lin_energy"] - float(G["out_lin_energy"])) <= 1e-9 * abs(float(G["out_lin_energy"]))
    for k in ("HA", "bA", "HL", "bL", "Hsc", "bsc"):
        assert relerr(o[k], G["out_" + k]) < 1e-4, k
    d = o["x"] - G["out_x"]
    Hf = o["Hf"]
    assert np.sqrt(abs(d @ Hf @ d)) <= 1e-3 * np.sqrt(abs(G["out_x"] @ Hf @ G["out_x"]))
    assert np.allclose(o["step"], G["out_step"], rtol=2e-4, atol=2e-6 * np.abs(G["out_step"]).max())
    assert o["opt_iterations"] == int(G["out_opt_iterations"])
    assert np.allclose(o["opt_energy"], G["out_opt_energy"], rtol=2e-3)
    assert int((o["opt_res_state"] != G["out_opt_res_state"]).sum()) <= 2
    upd = np.abs(G["out_opt_state"] - G["in_state"]).max(axis=0) + 1e-12
    assert (np.abs(o["opt_state"] - G["out_opt_state"]).max(axis=0) <= 5e-3 * upd).all()
    assert np.allclose(o["opt_idepth"], G["out_opt_idepth"], rtol=2e-3, atol=1e-5)


# ---- front-end rows: tests/golden/frontend_tiny.npz, made by tools/make_golden_frontend.py ----------
import _frontend_case as fc  # noqa: E402

FG = np.load(os.path.join(ROOT, "tests", "golden", "frontend_tiny.npz"))


def _check_frontend(o):
    for k in fc.EXACT:
        assert np.array_equal(np.asarray(o[k]), FG["out_" + k], equal_nan=True), k
    for k, tol in fc.CLOSE:
        assert relerr(o[k], FG["out_" + k]) < tol, (k, relerr(o[k], FG["out_" + k]))


def test_oracle_reproduces_frontend_golden(orc):
    """raw frame -> irradiance -> pyramid, pixel selection, immature points (constructor, two traces, activation) and the
    loop-closure alignment: the oracle reproduces the stored vectors bit for bit."""
    o = fc.run(orc, FG)
    _check_frontend(o)
    for k, _ in fc.CLOSE:
        assert np.array_equal(np.asarray(o[k]), FG["out_" + k]), k
    assert FG["out_trace_counts"][1][[0, 1, 2, 3]].min() > 0 and (FG["out_act_result"] == 1).sum() > 50


@pytest.mark.gpu
def test_cuda_matches_frontend_golden(gpu):
    _check_frontend(fc.run(gpu, FG))
