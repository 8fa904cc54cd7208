"""ctypes binding of the C ABI in include/sosba.h.

`Lib(path, prefix)` binds one shared library exporting `<prefix>_*` with the sosba.h signatures; the
product is `prefix="sosba"` (libsosba.so, CUDA).  The parity tests bind the CPU oracle with the same
class (`prefix="orc"`), which is why nothing in here names the oracle.  numpy arrays in, numpy arrays
out; every call checks the status code and raises `SosbaError`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PRECALC_FLOATS = 32
J_FLOATS = 74
PC_RTLL0, PC_TTLL0, PC_KRKI, PC_KT, PC_AFF, PC_B0, PC_DIST = 0, 9, 12, 21, 24, 26, 27
RES_IN, RES_OOB, RES_OUTLIER = 0, 1, 2

f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)


class SosbaError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("pyr_levels", C.c_int32), ("max_frames", C.c_int32),
                ("num_threads", C.c_int32), ("gamma_weights_pixel_select", C.c_int32),
                ("min_opt_iterations", C.c_int32), ("reserved0", C.c_int32),
                ("huber_th", C.c_float), ("outlier_th_sum_component", C.c_float),
                ("affine_opt_mode_a", C.c_float), ("affine_opt_mode_b", C.c_float),
                ("coarse_cutoff_th", C.c_float), ("idepth_fix_prior", C.c_float),
                ("idepth_fix_prior_marg_fac", C.c_float), ("frame_energy_th_const_weight", C.c_float),
                ("frame_energy_th_n", C.c_float), ("frame_energy_th_fac_median", C.c_float),
                ("overall_energy_th_weight", C.c_float), ("initial_calib_hessian", C.c_float),
                ("initial_rot_prior", C.c_float), ("initial_trans_prior", C.c_float),
                ("initial_aff_a_prior", C.c_float), ("initial_aff_b_prior", C.c_float),
                ("marg_weight_fac", C.c_float), ("th_opt_iterations", C.c_float)]


class Window(C.Structure):
    _fields_ = [("nf", C.c_int32), ("reserved0", C.c_int32), ("frame_slot", i32p), ("precalc", f32p),
                ("adHost", f64p), ("adTarget", f64p), ("adHTdeltaF", f32p), ("frame_energy_th", f32p),
                ("calib", C.c_float * 4), ("cDeltaF", C.c_float * 4), ("cPrior", C.c_double * 4),
                ("frame_prior", f64p), ("frame_delta_prior", f64p), ("frame_delta", f64p)]


class Points(C.Structure):
    _fields_ = [("n", C.c_int32), ("reserved0", C.c_int32), ("u", f32p), ("v", f32p), ("idepth", f32p),
                ("idepth_zero", f32p), ("color", f32p), ("weights", f32p), ("host", i32p), ("priorF", f32p),
                ("deltaF", f32p)]


class Residuals(C.Structure):
    _fields_ = [("n", C.c_int32), ("reserved0", C.c_int32), ("point", i32p), ("target", i32p), ("state", u8p),
                ("is_linearized", u8p), ("is_active", u8p), ("is_new", u8p), ("state_energy", f32p)]


class LinearizeOut(C.Structure):
    _fields_ = [("energy", C.c_double), ("new_frame_energy_th", C.c_float), ("n_in", C.c_int32),
                ("n_oob", C.c_int32), ("n_outlier", C.c_int32), ("n_removed", C.c_int32), ("reserved0", C.c_int32)]


class FrameState(C.Structure):
    _fields_ = [("camToWorld_evalPT", C.c_double * 12), ("state", C.c_double * 10), ("state_zero", C.c_double * 10),
                ("ab_exposure", C.c_float), ("frame_energy_th", C.c_float), ("frame_id", C.c_int32), ("slot", C.c_int32)]


class BAProblem(C.Structure):
    _fields_ = [("nf", C.c_int32), ("reserved0", C.c_int32), ("frames", C.POINTER(FrameState)),
                ("calib_value", C.c_double * 4), ("calib_value_zero", C.c_double * 4), ("points", Points),
                ("residuals", Residuals), ("HM", f64p), ("bM", f64p), ("idepth_out", f32p)]


class Immature(C.Structure):
    _fields_ = [("n", C.c_int32), ("reserved0", C.c_int32), ("host", i32p), ("u", f32p), ("v", f32p), ("color", f32p),
                ("weights", f32p), ("gradH", f32p), ("energy_th", f32p), ("idepth_min", f32p), ("idepth_max", f32p),
                ("quality", f32p), ("last_trace_status", u8p), ("last_trace_uv", f32p),
                ("last_trace_pixel_interval", f32p)]


class InitPoints(C.Structure):
    _fields_ = [("n", C.c_int32), ("reserved0", C.c_int32), ("u", f32p), ("v", f32p), ("idepth_new", f32p), ("iR", f32p), ("energy", f32p),
                ("outlierTH", f32p), ("isGood", u8p), ("energy_new", f32p), ("isGood_new", u8p), ("maxstep", f32p), ("lastHessian_new", f32p),
                ("JbBuffer_new", f32p)]


class ActivationWindow(C.Structure):
    _fields_ = [("nf", C.c_int32), ("min_obs", C.c_int32), ("frame_slot", i32p), ("RTll", f32p), ("tTll", f32p), ("aff", f32p),
                ("calib", C.c_float * 4), ("reserved0", C.c_int32), ("reserved1", C.c_int32)]


ACT_SKIP, ACT_ACTIVATED, ACT_DELETE = 0, 1, -1
IPS_GOOD, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED = range(6)


TRACK_MAX_PASSES = 8


class TrackHypothesis(C.Structure):
    _fields_ = [("q", C.c_double * 4), ("t", C.c_double * 3), ("aff_g2l", C.c_double * 2), ("min_res_for_abort", C.c_double * 5),
                ("last_residuals", C.c_double * 5), ("flow_indicators", C.c_double * 3), ("ok", C.c_int32), ("n_passes", C.c_int32),
                ("pass_lvl", C.c_int32 * TRACK_MAX_PASSES), ("pass_iterations", C.c_int32 * TRACK_MAX_PASSES),
                ("pass_accept", C.c_uint64 * TRACK_MAX_PASSES), ("pass_tie", C.c_uint64 * TRACK_MAX_PASSES), ("pass_residual", C.c_double * TRACK_MAX_PASSES),
                ("pass_cutoff_repeat", C.c_float * TRACK_MAX_PASSES)]


class ScaleHypothesis(C.Structure):
    _fields_ = [("scale", C.c_float), ("error", C.c_float), ("last_residuals", C.c_double * 5), ("n_passes", C.c_int32),
                ("pass_lvl", C.c_int32 * TRACK_MAX_PASSES), ("pass_iterations", C.c_int32 * TRACK_MAX_PASSES),
                ("pass_accept", C.c_uint64 * TRACK_MAX_PASSES), ("pass_tie", C.c_uint64 * TRACK_MAX_PASSES), ("reserved0", C.c_int32)]


class StepOut(C.Structure):
    _fields_ = [("energy", C.c_double), ("new_frame_energy_th", C.c_float), ("n_in", C.c_int32), ("n_oob", C.c_int32), ("n_outlier", C.c_int32),
                ("sum_a", C.c_double), ("sum_b", C.c_double), ("sum_t", C.c_double), ("sum_r", C.c_double), ("sum_id", C.c_double),
                ("sum_nid", C.c_double), ("num_id", C.c_double)]


class OptimizeOut(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("res_in_a", C.c_int32), ("energy_initial", C.c_double),
                ("energy_final", C.c_double), ("rmse", C.c_float), ("n_removed", C.c_int32),
                ("last_x_norm", C.c_double), ("reserved0", C.c_int32), ("reserved1", C.c_int32)]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


class Lib:
    """One loaded shared library exporting `<prefix>_*`."""

    def __init__(self, path: str, prefix: str = "sosba"):
        if not os.path.exists(path):
            raise SosbaError(f"{path} is missing - run `python -c 'import __graft_entry__ as g; g.build()'`")
        self.path = path
        self.prefix = prefix
        self.dll = C.CDLL(path, mode=C.RTLD_GLOBAL if prefix == "sosba" else C.RTLD_LOCAL)
        self._bind()

    def f(self, name):
        return getattr(self.dll, f"{self.prefix}_{name}")

    def has(self, name):
        return hasattr(self.dll, f"{self.prefix}_{name}")

    def _bind(self):
        self.f("last_error").restype = C.c_char_p
        self.f("config_default").restype = None
        for n in ("create", "frame_make_images", "frame_get_level", "window_set", "window_update", "points_set",
                  "points_update", "residuals_set", "reset_oob", "linearize_all", "apply_res", "fix_linearization",
                  "residuals_get_state", "residuals_get_jacobians", "residuals_get_aux", "points_get_stats",
                  "accumulate", "points_get_acc", "solve_system", "resubstitute", "marginalize_points",
                  "tracker_make_k", "tracker_set_ref", "tracker_calc_res_pose", "tracker_calc_gs_pose",
                  "scale_set_stereo", "scale_calc_res", "scale_calc_gs", "optimize", "ba_upload", "ba_iterate",
                  "ba_download", "ba_optimize", "pyr_levels", "immature_init", "trace_immature", "optimize_immature", "undistort_set", "frame_make_images_raw", "loop_set_points", "loop_calc_res", "loop_calc_gs", "pixel_selector_set", "pixel_select", "init_calc_res_and_gs", "immature_pool_set", "immature_pool_trace", "immature_pool_get"):
            self.f(n).restype = C.c_int
        self.f("destroy").restype = None
        if self.has("launch_count"):
            self.f("launch_count").restype = C.c_int64

    def config_default(self, w, h) -> Config:
        cfg = Config()
        self.f("config_default")(C.byref(cfg), C.c_int32(w), C.c_int32(h))
        return cfg

    def last_error(self) -> str:
        s = self.f("last_error")()
        return s.decode() if s else ""


class Handle:
    def __init__(self, lib: Lib, cfg: Config, device: int = 0):
        self.lib = lib
        self.cfg = cfg
        self.h = C.c_void_p()
        self._keep = {}
        self._ck(lib.f("create")(C.byref(cfg), C.c_int32(device), C.byref(self.h)), "create")
        self.levels = int(lib.f("pyr_levels")(self.h))
        self.nf = 0
        self.P = 0
        self.R = 0

    def close(self):
        if self.h:
            self.lib.f("destroy")(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise SosbaError(f"{self.lib.prefix}_{what} failed: rc={rc} {self.lib.last_error()}")

    def level_size(self, lvl):
        return self.cfg.w >> lvl, self.cfg.h >> lvl

    @property
    def D(self):
        return 4 + 8 * self.nf

    # ---- stream / sync / comm (product only) ----
    def set_stream(self, stream_ptr: int):
        self._ck(self.lib.f("set_stream")(self.h, C.c_void_p(stream_ptr)), "set_stream")

    def synchronize(self):
        self._ck(self.lib.f("synchronize")(self.h), "synchronize")

    def launch_count(self) -> int:
        return int(self.lib.f("launch_count")(self.h))

    def lib_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self._ck(self.lib.f("comm_unique_id")(buf), "comm_unique_id")
        return bytes(buf)

    def comm_uses_peer_memory(self) -> bool:
        return bool(self.lib.has("comm_uses_peer_memory") and self.lib.f("comm_uses_peer_memory")(self.h))

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self.lib.f("comm_init")(self.h, buf, C.c_int32(rank), C.c_int32(world)), "comm_init")

    # ---- a1 ----
    def frame_make_images(self, slot, color, B=None):
        color = _f32(color)
        assert color.size == self.cfg.w * self.cfg.h
        Bc = None if B is None else _f32(B)
        self._ck(self.lib.f("frame_make_images")(self.h, C.c_int32(slot), _p(color, f32p), _p(Bc, f32p)), "frame_make_images")

    def frame_get_level(self, slot, lvl):
        w, h = self.level_size(lvl)
        dI = np.empty((h, w, 3), np.float32)
        ab = np.empty((h, w), np.float32)
        self._ck(self.lib.f("frame_get_level")(self.h, C.c_int32(slot), C.c_int32(lvl), _p(dI, f32p), _p(ab, f32p)), "frame_get_level")
        return dI, ab

    # ---- uploads ----
    def _window_struct(self, win: dict):
        nf = int(win["nf"])
        k = {"frame_slot": _i32(win["frame_slot"]), "precalc": _f32(win["precalc"]), "adHost": _f64(win["adHost"]),
             "adTarget": _f64(win["adTarget"]), "adHTdeltaF": _f32(win["adHTdeltaF"]),
             "frame_energy_th": _f32(win["frame_energy_th"]), "frame_prior": _f64(win["frame_prior"]),
             "frame_delta_prior": _f64(win["frame_delta_prior"]), "frame_delta": _f64(win["frame_delta"])}
        assert k["precalc"].size == nf * nf * PRECALC_FLOATS and k["adHost"].size == nf * nf * 64
        W = Window()
        W.nf = nf
        W.frame_slot = _p(k["frame_slot"], i32p)
        W.precalc = _p(k["precalc"], f32p)
        W.adHost = _p(k["adHost"], f64p)
        W.adTarget = _p(k["adTarget"], f64p)
        W.adHTdeltaF = _p(k["adHTdeltaF"], f32p)
        W.frame_energy_th = _p(k["frame_energy_th"], f32p)
        W.calib = (C.c_float * 4)(*[float(x) for x in win["calib"]])
        W.cDeltaF = (C.c_float * 4)(*[float(x) for x in win["cDeltaF"]])
        W.cPrior = (C.c_double * 4)(*[float(x) for x in win["cPrior"]])
        W.frame_prior = _p(k["frame_prior"], f64p)
        W.frame_delta_prior = _p(k["frame_delta_prior"], f64p)
        W.frame_delta = _p(k["frame_delta"], f64p)
        return W, k

    def window_set(self, win: dict):
        W, keep = self._window_struct(win)
        self._ck(self.lib.f("window_set")(self.h, C.byref(W)), "window_set")
        self.nf = int(win["nf"])

    def window_update(self, win: dict):
        W, keep = self._window_struct(win)
        self._ck(self.lib.f("window_update")(self.h, C.byref(W)), "window_update")

    @staticmethod
    def _points_struct(pts: dict):
        n = int(len(pts["u"]))
        k = {"u": _f32(pts["u"]), "v": _f32(pts["v"]), "idepth": _f32(pts["idepth"]),
             "idepth_zero": _f32(pts["idepth_zero"]), "color": _f32(pts["color"]), "weights": _f32(pts["weights"]),
             "host": _i32(pts["host"]), "priorF": _f32(pts.get("priorF", np.zeros(n))),
             "deltaF": _f32(pts.get("deltaF", np.zeros(n)))}
        S = Points()
        S.n = n
        for name in ("u", "v", "idepth", "idepth_zero", "color", "weights", "priorF", "deltaF"):
            setattr(S, name, _p(k[name], f32p))
        S.host = _p(k["host"], i32p)
        return S, k

    @staticmethod
    def _residuals_struct(res: dict):
        n = int(len(res["point"]))
        k = {"point": _i32(res["point"]), "target": _i32(res["target"]),
             "state": _u8(res.get("state", np.zeros(n))), "is_linearized": _u8(res.get("is_linearized", np.zeros(n))),
             "is_active": _u8(res.get("is_active", np.zeros(n))), "is_new": _u8(res.get("is_new", np.ones(n))),
             "state_energy": _f32(res.get("state_energy", np.zeros(n)))}
        S = Residuals()
        S.n = n
        S.point = _p(k["point"], i32p)
        S.target = _p(k["target"], i32p)
        for name in ("state", "is_linearized", "is_active", "is_new"):
            setattr(S, name, _p(k[name], u8p))
        S.state_energy = _p(k["state_energy"], f32p)
        return S, k

    def points_set(self, pts: dict):
        S, keep = self._points_struct(pts)
        self._ck(self.lib.f("points_set")(self.h, C.byref(S)), "points_set")
        self.P = S.n

    def points_update(self, idepth=None, idepth_zero=None, deltaF=None):
        a = None if idepth is None else _f32(idepth)
        b = None if idepth_zero is None else _f32(idepth_zero)
        c = None if deltaF is None else _f32(deltaF)
        self._ck(self.lib.f("points_update")(self.h, _p(a, f32p), _p(b, f32p), _p(c, f32p)), "points_update")

    def residuals_set(self, res: dict):
        S, keep = self._residuals_struct(res)
        self._ck(self.lib.f("residuals_set")(self.h, C.byref(S)), "residuals_set")
        self.R = S.n

    # ---- a3-a5, a13 ----
    def reset_oob(self):
        self._ck(self.lib.f("reset_oob")(self.h), "reset_oob")

    def linearize_all(self, fix=False) -> dict:
        out = LinearizeOut()
        self._ck(self.lib.f("linearize_all")(self.h, C.c_int32(1 if fix else 0), C.byref(out)), "linearize_all")
        return {n: getattr(out, n) for n, _ in LinearizeOut._fields_ if n != "reserved0"}

    def apply_res(self):
        self._ck(self.lib.f("apply_res")(self.h), "apply_res")

    def fix_linearization(self, ids):
        ids = _i32(ids)
        self._ck(self.lib.f("fix_linearization")(self.h, _p(ids, i32p), C.c_int32(ids.size)), "fix_linearization")

    def get_state(self) -> dict:
        R = self.R
        o = {"state": np.empty(R, np.uint8), "new_state": np.empty(R, np.uint8), "energy": np.empty(R, np.float32),
             "new_energy": np.empty(R, np.float32), "new_energy_wo": np.empty(R, np.float32),
             "is_active": np.empty(R, np.uint8), "is_linearized": np.empty(R, np.uint8)}
        self._ck(self.lib.f("residuals_get_state")(self.h, _p(o["state"], u8p), _p(o["new_state"], u8p), _p(o["energy"], f32p),
                                                   _p(o["new_energy"], f32p), _p(o["new_energy_wo"], f32p),
                                                   _p(o["is_active"], u8p), _p(o["is_linearized"], u8p)), "residuals_get_state")
        return o

    def get_jacobians(self, committed: bool) -> np.ndarray:
        J = np.zeros((self.R, J_FLOATS), np.float32)
        self._ck(self.lib.f("residuals_get_jacobians")(self.h, C.c_int32(1 if committed else 0), _p(J, f32p)), "residuals_get_jacobians")
        return J

    def get_aux(self) -> dict:
        R = self.R
        o = {"JpJdF": np.zeros((R, 8), np.float32), "res_toZeroF": np.zeros((R, 8), np.float32),
             "projectedTo": np.zeros((R, 8, 2), np.float32), "centerProjectedTo": np.zeros((R, 3), np.float32)}
        self._ck(self.lib.f("residuals_get_aux")(self.h, _p(o["JpJdF"], f32p), _p(o["res_toZeroF"], f32p),
                                                 _p(o["projectedTo"], f32p), _p(o["centerProjectedTo"], f32p)), "residuals_get_aux")
        return o

    def points_get_stats(self):
        mrb = np.zeros(self.P, np.float32)
        ngr = np.zeros(self.P, np.int32)
        self._ck(self.lib.f("points_get_stats")(self.h, _p(mrb, f32p), _p(ngr, i32p)), "points_get_stats")
        return mrb, ngr

    # ---- a6-a11 ----
    def accumulate(self) -> dict:
        D = self.D
        o = {k: np.zeros((D, D)) for k in ("HA", "HL", "Hsc")}
        o.update({k: np.zeros(D) for k in ("bA", "bL", "bsc")})
        ra, rl = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.f("accumulate")(self.h, _p(o["HA"], f64p), _p(o["bA"], f64p), _p(o["HL"], f64p), _p(o["bL"], f64p),
                                          _p(o["Hsc"], f64p), _p(o["bsc"], f64p), C.byref(ra), C.byref(rl)), "accumulate")
        o["resInA"], o["resInL"] = ra.value, rl.value
        return o

    def points_get_acc(self) -> dict:
        P = self.P
        o = {"HddA": np.zeros(P, np.float32), "bdA": np.zeros(P, np.float32), "HcdA": np.zeros((P, 4), np.float32),
             "HddL": np.zeros(P, np.float32), "bdL": np.zeros(P, np.float32), "HcdL": np.zeros((P, 4), np.float32),
             "HdiF": np.zeros(P, np.float32), "bdSumF": np.zeros(P, np.float32)}
        self._ck(self.lib.f("points_get_acc")(self.h, *[_p(o[k], f32p) for k in ("HddA", "bdA", "HcdA", "HddL", "bdL", "HcdL", "HdiF", "bdSumF")]),
                 "points_get_acc")
        return o

    def solve_system(self, HM=None, bM=None):
        D = self.D
        x, Hf, bf = np.zeros(D), np.zeros((D, D)), np.zeros(D)
        hm = None if HM is None else _f64(HM)
        bm = None if bM is None else _f64(bM)
        self._ck(self.lib.f("solve_system")(self.h, _p(hm, f64p), _p(bm, f64p), _p(x, f64p), _p(Hf, f64p), _p(bf, f64p)), "solve_system")
        return x, Hf, bf

    def resubstitute(self, x):
        x = _f64(x)
        step = np.zeros(self.P, np.float32)
        self._ck(self.lib.f("resubstitute")(self.h, _p(x, f64p), _p(step, f32p)), "resubstitute")
        return step

    def marginalize_points(self, ids):
        ids = _i32(ids)
        D = self.D
        H, b = np.zeros((D, D)), np.zeros(D)
        n = C.c_int32(0)
        self._ck(self.lib.f("marginalize_points")(self.h, _p(ids, i32p), C.c_int32(ids.size), _p(H, f64p), _p(b, f64p), C.byref(n)), "marginalize_points")
        return H, b, n.value

    # ---- tracker / scale ----
    def tracker_make_k(self, calib):
        c = (C.c_float * 4)(*[float(x) for x in calib])
        self._ck(self.lib.f("tracker_make_k")(self.h, c), "tracker_make_k")

    def tracker_set_ref(self, lvl, u, v, idepth, color):
        u, v, idepth, color = _f32(u), _f32(v), _f32(idepth), _f32(color)
        self._ck(self.lib.f("tracker_set_ref")(self.h, C.c_int32(lvl), C.c_int32(u.size), _p(u, f32p), _p(v, f32p), _p(idepth, f32p), _p(color, f32p)), "tracker_set_ref")

    # ---- a16: makeCoarseDepthL0 / trackNewestCoarse / optimizeScale, resident on the device
    def tracker_make_coarse_depth(self, ref_slot, center_projected_to, HdiF):
        """-> pc_n per level"""
        c, hd = _f32(center_projected_to).reshape(-1, 3), _f32(HdiF)
        out = np.zeros(8, np.int32)
        self._ck(self.lib.f("tracker_make_coarse_depth")(self.h, C.c_int32(ref_slot), C.c_int32(hd.size), _p(c, f32p), _p(hd, f32p), _p(out, i32p)),
                 "tracker_make_coarse_depth")
        return out[:self.levels].copy()

    def tracker_get_ref(self, lvl):
        n = C.c_int32(0)
        self._ck(self.lib.f("tracker_get_ref")(self.h, C.c_int32(lvl), C.byref(n), None, None, None, None), "tracker_get_ref")
        a = [np.zeros(n.value, np.float32) for _ in range(4)]
        self._ck(self.lib.f("tracker_get_ref")(self.h, C.c_int32(lvl), C.byref(n), *[_p(x, f32p) for x in a]), "tracker_get_ref")
        return tuple(a)

    def distance_map(self, KRKi, Kt, host, u, v, idepth):
        """CoarseDistanceMap::makeDistanceMap -> (h1, w1) float32 map"""
        KRKi, Kt, host, u, v, idepth = _f32(KRKi).reshape(-1, 9), _f32(Kt).reshape(-1, 3), _i32(host), _f32(u), _f32(v), _f32(idepth)
        w1, h1 = self.cfg.w >> 1, self.cfg.h >> 1
        out = np.zeros((h1, w1), np.float32)
        self._ck(self.lib.f("distance_map")(self.h, C.c_int32(KRKi.shape[0]), _p(KRKi, f32p), _p(Kt, f32p), C.c_int32(host.size), _p(host, i32p), _p(u, f32p),
                                            _p(v, f32p), _p(idepth, f32p), _p(out, f32p)), "distance_map")
        return out

    def tracker_scale_coarse_depth(self, scale):
        self._ck(self.lib.f("tracker_scale_coarse_depth")(self.h, C.c_float(scale)), "tracker_scale_coarse_depth")

    def tracker_track(self, new_slot, ref_ab_exposure, new_ab_exposure, ref_aff_g2l, coarsest_lvl, hyps):
        """hyps: list of dicts {q (x, y, z, w), t, aff_g2l (a, b), min_res_for_abort (5, optional)} -> list of result dicts."""
        n = len(hyps)
        arr = (TrackHypothesis * max(n, 1))()
        for i, hy in enumerate(hyps):
            arr[i].q = (C.c_double * 4)(*[float(x) for x in hy["q"]])
            arr[i].t = (C.c_double * 3)(*[float(x) for x in hy["t"]])
            arr[i].aff_g2l = (C.c_double * 2)(*[float(x) for x in hy.get("aff_g2l", (0.0, 0.0))])
            arr[i].min_res_for_abort = (C.c_double * 5)(*[float(x) for x in hy.get("min_res_for_abort", [float("nan")] * 5)])
        ra = (C.c_double * 2)(float(ref_aff_g2l[0]), float(ref_aff_g2l[1]))
        self._ck(self.lib.f("tracker_track")(self.h, C.c_int32(new_slot), C.c_float(ref_ab_exposure), C.c_float(new_ab_exposure), ra,
                                             C.c_int32(coarsest_lvl), C.c_int32(n), arr), "tracker_track")
        out = []
        for i in range(n):
            a = arr[i]
            k = a.n_passes
            out.append(dict(ok=bool(a.ok), q=np.array(a.q[:]), t=np.array(a.t[:]), aff_g2l=np.array(a.aff_g2l[:]), last_residuals=np.array(a.last_residuals[:]),
                            flow_indicators=np.array(a.flow_indicators[:]), n_passes=k, pass_lvl=list(a.pass_lvl[:k]), pass_iterations=list(a.pass_iterations[:k]),
                            pass_accept=list(a.pass_accept[:k]), pass_tie=list(a.pass_tie[:k]), pass_residual=list(a.pass_residual[:k]), pass_cutoff_repeat=list(a.pass_cutoff_repeat[:k])))
        return out

    def scale_optimize(self, stereo_slot, coarsest_lvl, scales):
        n = len(scales)
        arr = (ScaleHypothesis * max(n, 1))()
        for i, sc in enumerate(scales):
            arr[i].scale = float(sc)
        self._ck(self.lib.f("scale_optimize")(self.h, C.c_int32(stereo_slot), C.c_int32(coarsest_lvl), C.c_int32(n), arr), "scale_optimize")
        return [dict(scale=arr[i].scale, error=arr[i].error, last_residuals=np.array(arr[i].last_residuals[:]), n_passes=arr[i].n_passes,
                     pass_lvl=list(arr[i].pass_lvl[:arr[i].n_passes]), pass_iterations=list(arr[i].pass_iterations[:arr[i].n_passes]),
                     pass_accept=list(arr[i].pass_accept[:arr[i].n_passes]), pass_tie=list(arr[i].pass_tie[:arr[i].n_passes])) for i in range(n)]

    # ---- CoarseInitializer::calcResAndGS (FullSystem/CoarseInitializer.cpp:450-673)
    def init_calc_res_and_gs(self, lvl, ref_slot, new_slot, refToNew34, aff, tlog, pts, alphaW=150.0 * 150.0, alphaK=2.5 * 2.5, couplingWeight=1.0):
        """pts: dict(u, v, idepth_new, iR, energy [n,2], outlierTH, isGood) -> dict(H, b, Hsc, bsc, res3, energy_new, isGood_new, maxstep,
        lastHessian_new, JbBuffer_new)."""
        T = _f64(refToNew34).reshape(12)
        n = len(pts["u"])
        i_ = {k: _f32(pts[k]) for k in ("u", "v", "idepth_new", "iR", "energy", "outlierTH")}
        good = _u8(pts["isGood"])
        o = dict(energy_new=np.zeros((n, 2), np.float32), isGood_new=np.zeros(n, np.uint8), maxstep=np.zeros(n, np.float32),
                 lastHessian_new=_f32(pts.get("lastHessian_new", np.zeros(n))).copy(), JbBuffer_new=_f32(pts.get("JbBuffer_new", np.zeros((n, 10)))).copy())
        ip = InitPoints(n=n, reserved0=0, u=_p(i_["u"], f32p), v=_p(i_["v"], f32p), idepth_new=_p(i_["idepth_new"], f32p), iR=_p(i_["iR"], f32p),
                        energy=_p(i_["energy"], f32p), outlierTH=_p(i_["outlierTH"], f32p), isGood=_p(good, u8p), energy_new=_p(o["energy_new"], f32p),
                        isGood_new=_p(o["isGood_new"], u8p), maxstep=_p(o["maxstep"], f32p), lastHessian_new=_p(o["lastHessian_new"], f32p),
                        JbBuffer_new=_p(o["JbBuffer_new"], f32p))
        H = np.zeros((8, 8), np.float32); b = np.zeros(8, np.float32); Hsc = np.zeros((8, 8), np.float32); bsc = np.zeros(8, np.float32)
        r3 = np.zeros(3, np.float32)
        a2 = (C.c_float * 2)(float(aff[0]), float(aff[1]))
        t3 = (C.c_float * 3)(*[float(x) for x in tlog])
        self._ck(self.lib.f("init_calc_res_and_gs")(self.h, C.c_int32(lvl), C.c_int32(ref_slot), C.c_int32(new_slot), _p(T, f64p), a2, t3,
                                                    C.c_float(alphaW), C.c_float(alphaK), C.c_float(couplingWeight), C.byref(ip), _p(H, f32p), _p(b, f32p),
                                                    _p(Hsc, f32p), _p(bsc, f32p), _p(r3, f32p)), "init_calc_res_and_gs")
        o.update(H=H, b=b, Hsc=Hsc, bsc=bsc, res3=r3)
        return o

    # ---- pixel selection (FullSystem/PixelSelector2.cpp)
    def pixel_selector_set(self, random_pattern, current_potential=3):
        rp = _u8(random_pattern)
        assert rp.size == self.cfg.w * self.cfg.h
        self._ck(self.lib.f("pixel_selector_set")(self.h, _p(rp, u8p), C.c_int32(current_potential)), "pixel_selector_set")

    def pixel_select(self, slot, density, recursions_left=1, th_factor=1.0, cap=None, want_map=True):
        """-> dict(n, u, v, type, map, potential): makeMaps' return value, the raster-order list of selected pixels, the full
        status map (float, 0/1/2/4) and currentPotential after the call."""
        cap = cap if cap is not None else self.cfg.w * self.cfg.h // 4
        u = np.zeros(cap, np.int32); v = np.zeros(cap, np.int32); t = np.zeros(cap, np.float32)
        m = np.zeros((self.cfg.h, self.cfg.w), np.float32) if want_map else None
        n = C.c_int32(0); pot = C.c_int32(0)
        self._ck(self.lib.f("pixel_select")(self.h, C.c_int32(slot), C.c_float(density), C.c_int32(recursions_left), C.c_float(th_factor),
                                            C.c_int32(cap), C.byref(n), _p(u, i32p), _p(v, i32p), _p(t, f32p), _p(m, f32p), C.byref(pot)), "pixel_select")
        return dict(n=n.value, u=u[:n.value], v=v[:n.value], type=t[:n.value], map=m, potential=pot.value)

    # ---- loop-closure direct alignment (LoopClosure/PoseEstimator.cpp:75-284)
    def loop_set_points(self, xyz, color):
        xyz, color = _f64(xyz), _f32(color)
        self._ck(self.lib.f("loop_set_points")(self.h, C.c_int32(xyz.size // 3), _p(xyz, f64p), _p(color, f32p)), "loop_set_points")

    def loop_calc_res(self, lvl, slot, refToNew34, affLL, cutoff):
        T = _f64(refToNew34).reshape(12)
        a = (C.c_float * 2)(float(affLL[0]), float(affLL[1]))
        out6 = np.zeros(6)
        cnt = np.zeros(3, np.int32)
        self._ck(self.lib.f("loop_calc_res")(self.h, C.c_int32(lvl), C.c_int32(slot), _p(T, f64p), a, C.c_float(cutoff), _p(out6, f64p), _p(cnt, i32p)),
                 "loop_calc_res")
        return out6, cnt

    def loop_calc_gs(self, lvl, a, b0):
        H = np.zeros((8, 8))
        b = np.zeros(8)
        self._ck(self.lib.f("loop_calc_gs")(self.h, C.c_int32(lvl), C.c_float(a), C.c_float(b0), _p(H, f64p), _p(b, f64p)), "loop_calc_gs")
        return H, b

    # ---- pre-pyramid image path (util/Undistort.cpp:194-227, 361-458)
    def undistort_set(self, w_org, h_org, remapX=None, remapY=None, G=None, vignette_inv=None):
        rx = None if remapX is None else _f32(remapX)
        ry = None if remapY is None else _f32(remapY)
        g = None if G is None else _f32(G)
        vi = None if vignette_inv is None else _f32(vignette_inv)
        self._ck(self.lib.f("undistort_set")(self.h, C.c_int32(w_org), C.c_int32(h_org), _p(rx, f32p), _p(ry, f32p), _p(g, f32p),
                                             C.c_int32(0 if g is None else g.size), _p(vi, f32p)), "undistort_set")

    def frame_make_images_raw(self, slot, raw, factor=1.0, B=None, want_image=False):
        raw = np.ascontiguousarray(raw)
        assert raw.dtype in (np.uint8, np.uint16)
        b = None if B is None else _f32(B)
        out = np.zeros((self.cfg.h, self.cfg.w), np.float32) if want_image else None
        self._ck(self.lib.f("frame_make_images_raw")(self.h, C.c_int32(slot), raw.ctypes.data_as(C.c_void_p), C.c_int32(8 * raw.dtype.itemsize),
                                                     C.c_float(factor), _p(b, f32p), _p(out, f32p)), "frame_make_images_raw")
        return out

    # ---- immature points (ImmaturePoint.cpp:28-60, 70-415; FullSystem.cpp:311-361)
    def immature_init(self, host_slot, u, v):
        """-> dict of the constructor's outputs for candidate pixels (u, v) of the frame in `host_slot`; idepth interval,
        quality and status are the fresh point's (0 / NaN, 10000, UNINITIALIZED)."""
        u, v = _i32(u), _i32(v)
        n = u.size
        out = dict(u=u.astype(np.float32), v=v.astype(np.float32), color=np.zeros((n, 8), np.float32), weights=np.zeros((n, 8), np.float32),
                   gradH=np.zeros((n, 4), np.float32), energy_th=np.zeros(n, np.float32), idepth_min=np.zeros(n, np.float32),
                   idepth_max=np.full(n, np.nan, np.float32), quality=np.full(n, 10000, np.float32),
                   status=np.full(n, IPS_UNINITIALIZED, np.uint8), uv=np.zeros((n, 2), np.float32), pixel_interval=np.zeros(n, np.float32))
        self._ck(self.lib.f("immature_init")(self.h, C.c_int32(host_slot), C.c_int32(n), _p(u, i32p), _p(v, i32p), _p(out["color"], f32p),
                                             _p(out["weights"], f32p), _p(out["gradH"], f32p), _p(out["energy_th"], f32p)), "immature_init")
        return out

    def trace_immature(self, frame_slot, host, KRKi, Kt, aff, pts):
        """traceOn of every point in `pts` (the dict of immature_init, updated in place) against `frame_slot`;
        host [n] indexes KRKi [nh,3,3] / Kt [nh,3] / aff [nh,2].  -> counts[6] per ImmaturePointStatus."""
        host, KRKi, Kt, aff = _i32(host), _f32(KRKi), _f32(Kt), _f32(aff)
        for k in ("u", "v", "color", "weights", "gradH", "energy_th", "idepth_min", "idepth_max", "quality", "uv", "pixel_interval"):
            pts[k] = _f32(pts[k])
        pts["status"] = _u8(pts["status"])
        ip = Immature(n=host.size, reserved0=0, host=_p(host, i32p), u=_p(pts["u"], f32p), v=_p(pts["v"], f32p), color=_p(pts["color"], f32p),
                      weights=_p(pts["weights"], f32p), gradH=_p(pts["gradH"], f32p), energy_th=_p(pts["energy_th"], f32p),
                      idepth_min=_p(pts["idepth_min"], f32p), idepth_max=_p(pts["idepth_max"], f32p), quality=_p(pts["quality"], f32p),
                      last_trace_status=_p(pts["status"], u8p), last_trace_uv=_p(pts["uv"], f32p),
                      last_trace_pixel_interval=_p(pts["pixel_interval"], f32p))
        counts = np.zeros(6, np.int32)
        self._ck(self.lib.f("trace_immature")(self.h, C.c_int32(frame_slot), C.c_int32(KRKi.size // 9), _p(KRKi, f32p), _p(Kt, f32p), _p(aff, f32p),
                                              C.byref(ip), _p(counts, i32p)), "trace_immature")
        return counts

    def _immature_struct(self, host, pts):
        host = _i32(host)
        for k in ("u", "v", "color", "weights", "gradH", "energy_th", "idepth_min", "idepth_max", "quality", "uv", "pixel_interval"):
            pts[k] = _f32(pts[k])
        pts["status"] = _u8(pts["status"])
        ip = Immature(n=host.size, reserved0=0, host=_p(host, i32p), u=_p(pts["u"], f32p), v=_p(pts["v"], f32p), color=_p(pts["color"], f32p),
                      weights=_p(pts["weights"], f32p), gradH=_p(pts["gradH"], f32p), energy_th=_p(pts["energy_th"], f32p),
                      idepth_min=_p(pts["idepth_min"], f32p), idepth_max=_p(pts["idepth_max"], f32p), quality=_p(pts["quality"], f32p),
                      last_trace_status=_p(pts["status"], u8p), last_trace_uv=_p(pts["uv"], f32p),
                      last_trace_pixel_interval=_p(pts["pixel_interval"], f32p))
        return ip, host

    def immature_pool_set(self, host, pts):
        ip, keep = self._immature_struct(host, pts)
        self._ck(self.lib.f("immature_pool_set")(self.h, C.byref(ip)), "immature_pool_set")

    def immature_pool_trace(self, frame_slot, KRKi, Kt, aff, want_counts=True):
        KRKi, Kt, aff = _f32(KRKi), _f32(Kt), _f32(aff)
        counts = np.zeros(6, np.int32) if want_counts else None
        self._ck(self.lib.f("immature_pool_trace")(self.h, C.c_int32(frame_slot), C.c_int32(KRKi.size // 9), _p(KRKi, f32p), _p(Kt, f32p), _p(aff, f32p),
                                                   _p(counts, i32p)), "immature_pool_trace")
        return counts

    def immature_pool_get(self, host, pts):
        """reads idepth_min/max, quality, status, uv, pixel_interval of the resident pool back into `pts` (in place)."""
        ip, keep = self._immature_struct(host, pts)
        self._ck(self.lib.f("immature_pool_get")(self.h, C.byref(ip)), "immature_pool_get")

    def optimize_immature(self, frame_slot, RTll, tTll, aff, calib, host, pts, min_obs=1):
        """optimizeImmaturePoint (FullSystemOptPoint.cpp:47-192) of every point of `pts` (dict of immature_init, after tracing).
        RTll [nf,nf,3,3], tTll [nf,nf,3], aff [nf,nf,2] per (host, target); calib = fxl fyl cxl cyl.
        -> result [n] int8 (ACT_*), idepth [n], res_state [n, nf]."""
        frame_slot, RTll, tTll, aff, host = _i32(frame_slot), _f32(RTll), _f32(tTll), _f32(aff), _i32(host)
        nf, n = frame_slot.size, host.size
        arr = {k: _f32(pts[k]) for k in ("u", "v", "color", "weights", "energy_th", "idepth_min", "idepth_max")}
        win = ActivationWindow(nf=nf, min_obs=min_obs, frame_slot=_p(frame_slot, i32p), RTll=_p(RTll, f32p), tTll=_p(tTll, f32p), aff=_p(aff, f32p),
                               calib=(C.c_float * 4)(*[float(x) for x in calib]))
        ip = Immature(n=n, reserved0=0, host=_p(host, i32p), u=_p(arr["u"], f32p), v=_p(arr["v"], f32p), color=_p(arr["color"], f32p),
                      weights=_p(arr["weights"], f32p), energy_th=_p(arr["energy_th"], f32p), idepth_min=_p(arr["idepth_min"], f32p),
                      idepth_max=_p(arr["idepth_max"], f32p))
        result = np.zeros(n, np.int8)
        idepth = np.zeros(n, np.float32)
        res_state = np.zeros((n, nf), np.uint8)
        self._ck(self.lib.f("optimize_immature")(self.h, C.byref(win), C.byref(ip), _p(result, C.POINTER(C.c_int8)), _p(idepth, f32p), _p(res_state, u8p)),
                 "optimize_immature")
        return result, idepth, res_state

    def tracker_calc_res_pose(self, lvl, slot, refToNew34, affLL, cutoff):
        T = _f64(refToNew34).reshape(12)
        a = (C.c_float * 2)(float(affLL[0]), float(affLL[1]))
        out6 = np.zeros(6)
        cnt = np.zeros(3, np.int32)
        self._ck(self.lib.f("tracker_calc_res_pose")(self.h, C.c_int32(lvl), C.c_int32(slot), _p(T, f64p), a, C.c_float(cutoff), _p(out6, f64p), _p(cnt, i32p)),
                 "tracker_calc_res_pose")
        return out6, cnt

    def tracker_calc_gs_pose(self, lvl, a, b0):
        H, b = np.zeros((8, 8)), np.zeros(8)
        self._ck(self.lib.f("tracker_calc_gs_pose")(self.h, C.c_int32(lvl), C.c_float(a), C.c_float(b0), _p(H, f64p), _p(b, f64p)), "tracker_calc_gs_pose")
        return H, b

    def scale_set_stereo(self, T10_34, K1):
        T = _f64(T10_34).reshape(12)
        k = (C.c_float * 4)(*[float(x) for x in K1])
        self._ck(self.lib.f("scale_set_stereo")(self.h, _p(T, f64p), k), "scale_set_stereo")

    def scale_calc_res(self, lvl, slot, scale, cutoff):
        out6 = np.zeros(6)
        cnt = np.zeros(3, np.int32)
        self._ck(self.lib.f("scale_calc_res")(self.h, C.c_int32(lvl), C.c_int32(slot), C.c_float(scale), C.c_float(cutoff), _p(out6, f64p), _p(cnt, i32p)), "scale_calc_res")
        return out6, cnt

    def scale_calc_gs(self, lvl, scale):
        H, b = C.c_float(0), C.c_float(0)
        self._ck(self.lib.f("scale_calc_gs")(self.h, C.c_int32(lvl), C.c_float(scale), C.byref(H), C.byref(b)), "scale_calc_gs")
        return H.value, b.value

    # ---- composed GN loop ----
    def make_problem(self, frames: list, calib_value, calib_value_zero, pts: dict, res: dict, HM=None, bM=None):
        """frames: list of dicts {evalPT (3x4 or 4x4), state, state_zero, ab_exposure, frame_energy_th, frame_id, slot}."""
        nf = len(frames)
        arr = (FrameState * nf)()
        for i, f in enumerate(frames):
            T = np.asarray(f["evalPT"], np.float64)[:3, :4].reshape(12)
            arr[i].camToWorld_evalPT = (C.c_double * 12)(*T)
            arr[i].state = (C.c_double * 10)(*np.asarray(f["state"], np.float64))
            arr[i].state_zero = (C.c_double * 10)(*np.asarray(f["state_zero"], np.float64))
            arr[i].ab_exposure = float(f.get("ab_exposure", 1.0))
            arr[i].frame_energy_th = float(f.get("frame_energy_th", 8 * 8 * 8))
            arr[i].frame_id = int(f.get("frame_id", i))
            arr[i].slot = int(f.get("slot", i))
        P = BAProblem()
        P.nf = nf
        P.frames = arr
        P.calib_value = (C.c_double * 4)(*[float(x) for x in calib_value])
        P.calib_value_zero = (C.c_double * 4)(*[float(x) for x in calib_value_zero])
        ps, pk = self._points_struct(pts)
        rs, rk = self._residuals_struct(res)
        P.points, P.residuals = ps, rs
        hm = None if HM is None else _f64(HM)
        bm = None if bM is None else _f64(bM)
        P.HM, P.bM = _p(hm, f64p), _p(bm, f64p)
        idout = np.zeros(ps.n, np.float32)
        P.idepth_out = _p(idout, f32p)
        keep = (arr, pk, rk, hm, bm, idout)
        self.nf, self.P, self.R = nf, ps.n, rs.n
        return P, keep

    @staticmethod
    def problem_result(P, keep) -> dict:
        arr = keep[0]
        nf = P.nf
        return {"evalPT": np.array([list(arr[i].camToWorld_evalPT) for i in range(nf)]).reshape(nf, 3, 4),
                "state": np.array([list(arr[i].state) for i in range(nf)]),
                "state_zero": np.array([list(arr[i].state_zero) for i in range(nf)]),
                "frame_energy_th": np.array([arr[i].frame_energy_th for i in range(nf)], np.float32),
                "calib_value": np.array(list(P.calib_value)), "idepth": keep[5].copy()}

    def optimize(self, P, max_iterations: int) -> dict:
        out = OptimizeOut()
        self._ck(self.lib.f("optimize")(self.h, C.byref(P), C.c_int32(max_iterations), C.byref(out)), "optimize")
        return {n: getattr(out, n) for n, _ in OptimizeOut._fields_ if n != "reserved1"}

    def ba_optimize(self, max_iterations: int) -> dict:
        out = OptimizeOut()
        self._ck(self.lib.f("ba_optimize")(self.h, C.c_int32(max_iterations), C.byref(out)), "ba_optimize")
        return {n: getattr(out, n) for n, _ in OptimizeOut._fields_ if n != "reserved1"}

    def profile_enable(self, on):
        """True / 1: start a new measurement, False / 0: pause, 2: resume without dropping the recorded brackets"""
        self._ck(self.lib.f("profile_enable")(self.h, C.c_int32(int(on))), "profile_enable")

    def profile_read(self):
        ms, n = C.c_double(0), C.c_int32(0)
        self._ck(self.lib.f("profile_read")(self.h, C.byref(ms), C.byref(n)), "profile_read")
        return ms.value, n.value

    def trace_enable(self, on: bool):
        self._ck(self.lib.f("trace_enable")(self.h, C.c_int32(1 if on else 0)), "trace_enable")

    def trace_read(self, kernel: str, skip_first: int = 0):
        """-> (mean ns of a launch inside the programmatic launch chain, launches) for the kernels whose name starts with `kernel`"""
        ns, n = C.c_double(0), C.c_int32(0)
        self._ck(self.lib.f("trace_read")(self.h, kernel.encode(), C.c_int32(skip_first), C.byref(ns), C.byref(n)), "trace_read")
        return ns.value, n.value

    def frame_make_images_dev(self, slot, color_ptr: int, B_ptr: int = 0):
        self._ck(self.lib.f("frame_make_images_dev")(self.h, C.c_int32(slot), C.c_void_p(color_ptr), C.c_void_p(B_ptr or None)), "frame_make_images_dev")

    def ba_upload(self, P):
        self._ck(self.lib.f("ba_upload")(self.h, C.byref(P)), "ba_upload")
        self.nf_ba = int(P.nf)

    def ba_system(self):
        """-> dict(H_top, b_top, H_sc, b_sc, resInA, resInL): the device part of solveSystemF before a caller-side solve"""
        D = 4 + 8 * self.nf_ba
        o = {"H_top": np.zeros((D, D)), "b_top": np.zeros(D), "H_sc": np.zeros((D, D)), "b_sc": np.zeros(D)}
        a, l = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.f("ba_system")(self.h, _p(o["H_top"], f64p), _p(o["b_top"], f64p), _p(o["H_sc"], f64p), _p(o["b_sc"], f64p), C.byref(a), C.byref(l)), "ba_system")
        o["resInA"], o["resInL"] = a.value, l.value
        return o

    def ba_step(self, x) -> dict:
        x = _f64(x)
        out = StepOut()
        self._ck(self.lib.f("ba_step")(self.h, _p(x, f64p), C.byref(out)), "ba_step")
        return {n: getattr(out, n) for n, _ in StepOut._fields_}

    def ba_iterate(self, n: int) -> int:
        nres = C.c_int32(0)
        self._ck(self.lib.f("ba_iterate")(self.h, C.c_int32(n), C.byref(nres)), "ba_iterate")
        return nres.value

    def ba_download(self, P):
        self._ck(self.lib.f("ba_download")(self.h, C.byref(P)), "ba_download")
