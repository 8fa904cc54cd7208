#!/usr/bin/env python
"""bench.py — photometric residuals/sec and GN-iteration ms of the windowed bundle-adjustment hot path.

Workload (BASELINE.json configs[1]): synthetic 640x480, 8-keyframe window, 2000 active points, BA only.
One STEP = one FullSystem::optimize(6) of that window (FullSystemOptimize.cpp:305-489): resetOOB,
linearizeAll + applyRes, 6 x {solveSystemF, doStepFromBackup, linearizeAll, applyRes}, new evaluation point,
linearizeAll(fix) — 8 linearisation passes and 6 solves, forced (min_opt_iterations) so every step does the
same work.  1 photometric residual = 1 of the 8 pattern samples of a PointFrameResidual.

  value  : residuals/s with the window resident in HBM (sosba_ba_optimize), CUDA events per step, L2 flushed
           between steps, max over ranks.
  e2e    : the same step through the public C ABI with HOST buffers (sosba_frame_make_images of the newest
           keyframe + sosba_optimize: H2D of image, points, residuals, frame states; D2H of the result).
  --impl reference : the CPU restatement of the reference path (oracle/, -O3 -mavx2 -mfma, all host threads;
           the reference itself cannot be built here — see DESIGN.md) on the same window, same step.
N > 1 (weak scaling): N x 2000 points, point-sharded, one all-reduce of the block tables per GN iteration.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "photometric_residuals_per_sec"
UNIT = "residuals/s"
ITERS = 6
BYTES_LINEARIZE = 616          # algorithmic bytes per PointFrameResidual of the fused linearize+applyRes launch of the GN loop (DESIGN.md §4.1;
                               # the un-fused API variant, SURVEY.md §8d: 787)
WORKLOAD = "synthetic 640x480 stereo, 8-KF window, 2000 active points, BA only (BASELINE.json configs[1])"


def load():
    from sosba_loader import load_package
    pkg = load_package()
    from sos_slam_b200 import binding, problem, synth
    return pkg, binding, problem, synth


WORKLOADS = {"configB": WORKLOAD,
             "euroc": "EuRoC-shaped 752x480 stereo, 8-KF window, 2000 active points, BA only, IMU off (BASELINE.json configs[2] shape)",
             "kitti": "KITTI-shaped 1232x368 stereo, 12-KF window, 4000 active points, BA only (BASELINE.json configs[3] shape)",
             "tumvi": "TUM-VI-shaped 512x512, 8-KF window, 2000 active points, BA only, hot path only (BASELINE.json configs[4] shape)",
             "euroc_imu": "EuRoC-shaped 752x480 stereo+IMU, 8-KF window, 2000 active points: visual system on the device, IMU/KKT widening and "
                          "solve on the host with a stand-in IMU Hessian of the reference's shape (BASELINE.json configs[2])"}
_workload = "configB"
PROF_EVERY = 4


def get_scene(synth, n_points_factor=1):
    from _scenes import CONFIG_B, EUROC, KITTI, TUMVI, scene
    sc = scene(**{"configB": CONFIG_B, "euroc": EUROC, "euroc_imu": EUROC, "kitti": KITTI, "tumvi": TUMVI}[_workload])
    if n_points_factor > 1:
        sc = synth.replicate_points(sc, n_points_factor)
    return sc


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons of one GPU, sampled DURING the timed region (NVML in a thread, every 5 ms;
    falls back to `nvidia-smi -lms` when pynvml is unavailable)."""

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.th = None
        self.nvml = None
        self.proc = None
        self.rows = []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                for bit, name in self.NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.th.join(timeout=1)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 5 ms"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


L2_NOTE = "GPU arm: L2 flushed between steps (256 MiB fill), per-step CUDA events on the launch stream; CPU arm: host wall clock"


def config_of(sc, factor, scaling):
    """The same dictionary in both arms (the driver compares them)."""
    wl = WORKLOAD if factor == 1 else WORKLOAD + f" x{factor} points"
    return {"workload": wl, "points": int(sc.n_points), "residuals": int(sc.n_residuals), "gn_iterations_per_step": ITERS,
            "linearizations_per_step": ITERS + 2, "scaling": scaling, "l2": L2_NOTE}


def lm_timings(h, sc, synth, reps=10, make_images=False):
    """Host wall per call of makeCoarseDepthL0, trackNewestCoarse (1 and 32 hypotheses) and optimizeScale (the 7 start values of
    FullSystem::optimizeScale) on the window's newest keyframe; the same calls time libsosba and the CPU port."""
    from _track_case import hypotheses, ref_affine, stereo_frame
    if make_images:
        for i, img in enumerate(sc.images):
            h.frame_make_images(i, img)
    rng = np.random.default_rng(9)
    nref = sc.nf - 1
    # every active point hosted in an older keyframe projects into the newest one: (u, v, idepth) of the centre projections
    m = sc.pt_host != nref
    Tn = np.linalg.inv(sc.camToWorld_true[nref])
    fx, fy, cx, cy = [float(x) for x in sc.K]
    cpt = []
    for hst in np.unique(sc.pt_host[m]):
        q = m & (sc.pt_host == hst)
        z = 1.0 / sc.pt_idepth_true[q]
        P = np.stack([(sc.pt_u[q] - cx) / fx * z, (sc.pt_v[q] - cy) / fy * z, z, np.ones_like(z)], 0)
        Pn = (Tn @ sc.camToWorld_true[hst] @ P)[:3]
        cpt.append(np.stack([fx * Pn[0] / Pn[2] + cx, fy * Pn[1] / Pn[2] + cy, 1.0 / Pn[2]], 1))
    cpt = np.concatenate(cpt).astype(np.float32)
    ok = (cpt[:, 0] > 3) & (cpt[:, 1] > 3) & (cpt[:, 0] < sc.w - 4) & (cpt[:, 1] < sc.h - 4) & (cpt[:, 2] > 0)
    cpt = cpt[ok]
    hdi = rng.uniform(1e-4, 1e-2, len(cpt)).astype(np.float32)
    K = sc.K.astype(np.float32)
    h.tracker_make_k(K)
    n = h.tracker_make_coarse_depth(nref, cpt, hdi)
    t0 = time.perf_counter()
    for _ in range(reps):
        h.tracker_make_coarse_depth(nref, cpt, hdi)
    cd_ms = 1e3 * (time.perf_counter() - t0) / reps
    _, hyps = hypotheses(sc, n_extra=30, seed=2)
    ref_aff, ref_exp, new_exp = ref_affine(sc)
    out = {}
    for tag, hy in (("1", hyps[:1]), ("32", hyps)):
        h.tracker_track(sc.nf - 2, ref_exp, new_exp, ref_aff, h.levels - 1, hy)
        t0 = time.perf_counter()
        for _ in range(reps):
            r = h.tracker_track(sc.nf - 2, ref_exp, new_exp, ref_aff, h.levels - 1, hy)
        out["track_%s_hyp_ms" % tag] = 1e3 * (time.perf_counter() - t0) / reps
        out["track_%s_lm_iterations" % tag] = int(sum(sum(x["pass_iterations"]) for x in r))
    T10, img1 = stereo_frame(sc, 1234 if (sc.w, sc.h) == (640, 480) else 0)
    h.scale_set_stereo(T10, K)
    h.frame_make_images(sc.nf, img1)
    starts = [0.1, 0.2, 0.5, 1, 2, 5, 10]
    h.scale_optimize(sc.nf, h.levels - 1, starts)
    t0 = time.perf_counter()
    for _ in range(reps):
        rs = h.scale_optimize(sc.nf, h.levels - 1, starts)
    out.update({"make_coarse_depth_ms": cd_ms, "reference_points_per_level": [int(x) for x in n], "splatted_points": int(len(cpt)),
                "optimize_scale_7_starts_ms": 1e3 * (time.perf_counter() - t0) / reps,
                "optimize_scale_lm_iterations": int(sum(sum(x["pass_iterations"]) for x in rs)),
                "note": "host wall per call incl. H2D of the inputs and D2H of the results; hypotheses = perturbed true pose, identity, 30 random"})
    return out


def forced_cfg(lib, sc, threads=1):
    cfg = lib.config_default(sc.w, sc.h)
    cfg.num_threads = threads
    cfg.max_frames = sc.nf + 2
    cfg.min_opt_iterations = 1000   # never break early: every step runs exactly ITERS Gauss-Newton iterations
    return cfg


def _ref_factor(args):
    """Problem size of the CPU arm = the GPU arm's: weak scaling multiplies the points by the GPU count, strong keeps them."""
    return args.points_factor or (max(1, args.gpus) if args.scaling == "weak" else 1)


def _time_cpu_steps(binding, problem, lib, sc, threads, steps, warmup, budget_s=None):
    """`steps` FullSystem::optimize(6) calls of the oracle speed build on window `sc` with `threads` IndexThreadReduce
    workers.  The caller's problem struct is built once (as the GPU arm's e2e does) and its in/out members are restored
    before every step; the step itself = makeImages of the newest keyframe + optimize()."""
    import ctypes
    cfg = forced_cfg(lib, sc, threads=threads)
    h = binding.Handle(lib, cfg)
    for i, img in enumerate(sc.images):
        h.frame_make_images(i, img)
    val, val0 = problem.calib_of(sc)
    P, keep = h.make_problem(problem.frames_of(sc), val, val0, problem.points_of(sc), problem.residuals_of(sc))
    frames0 = bytes(ctypes.string_at(ctypes.addressof(keep[0]), ctypes.sizeof(keep[0])))
    calib0 = list(P.calib_value)

    def step():
        ctypes.memmove(ctypes.addressof(keep[0]), frames0, len(frames0))
        for i in range(4):
            P.calib_value[i] = calib0[i]
        h.frame_make_images(sc.nf - 1, sc.images[-1])
        return h.optimize(P, ITERS)

    for _ in range(warmup):
        step()
    nres, n, dt = 0, 0, 0.0
    while n < steps:
        t0 = time.perf_counter()
        out = step()
        dt += time.perf_counter() - t0
        nres += out["reserved0"] * 8 * (out["iterations"] + 2)
        n += 1
        if budget_s is not None and dt > budget_s and n >= 3:
            break
    h.close()
    return nres, dt, n


def run_reference(args):
    """The reference's CPU path (oracle restatement, speed build) on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg, binding, problem, synth = load()
    path = os.path.join(ROOT, "oracle", "_build", "liborc_speed.so")
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    lib = binding.Lib(path, "orc")
    factor = _ref_factor(args)
    sc = get_scene(synth, factor)
    cores = os.cpu_count() or 1
    nres, dt, n = _time_cpu_steps(binding, problem, lib, sc, cores, args.steps, args.warmup)
    v = nres / dt
    # the reference's own threading: NUM_THREADS = 6 workers (util/NumType.h:37), bounded sample
    nres6, dt6, n6 = _time_cpu_steps(binding, problem, lib, sc, 6, min(args.steps, 10), 1, budget_s=10.0)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / n, "gn_iter_ms": 1e3 * dt / n / ITERS, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(sc, factor, args.scaling),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} full optimize() steps of the workload, oracle speed build (-O3 -mavx2 -mfma), {cores} IndexThreadReduce workers",
                             "reference_threading": {"workers": 6, "value": nres6 / dt6, "ms_per_step": 1e3 * dt6 / n6, "steps": n6,
                                                     "note": "NUM_THREADS = 6 as the reference hard-codes it (util/NumType.h:37)"}},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- configs[2]: the IMU configuration ---------------------------------------------------------------------------------------
class ImuStandIn:
    """Caller-side half of EnergyFunctional::solveSystemF with IMU (EnergyFunctional.cpp:1052-1171) for the benchmark: the widening
    8 -> 29 states per frame + scale (expandHbtoFitImu :256-286), an IMU Hessian, the KKT rows of the spline constraints (6 per
    consecutive frame pair, 3 for the last, :323-330), the removal of the scale column when scale optimisation is on (:1113), Jacobi
    scaling, LDL^T, and the unpacking of x_dso (:1150-1171).  The IMU Hessian and the constraint Jacobians are a STAND-IN with the
    reference's sparsity and magnitudes (bias random walk + a banded positive definite 29x29 block per frame, unit-scaled constraint
    rows): getImuHessian needs the FrameHessian spline state and stays in the reference.  It measures the data path
    of the configuration, not IMU physics."""

    def __init__(self, nf, seed=0):
        rng = np.random.default_rng(seed)
        self.nf, self.dim = nf, 4 + 1 + 29 * nf
        H = np.zeros((self.dim, self.dim))
        b = np.zeros(self.dim)
        for f in range(1, nf):
            c, p = 5 + 29 * f, 5 + 29 * (f - 1)
            A = rng.standard_normal((29, 29)) * 3.0
            H[c:c + 29, c:c + 29] += A @ A.T + 50.0 * np.eye(29)
            W = 1e3 * np.eye(6)
            H[p + 8:p + 14, p + 8:p + 14] += W; H[c + 8:c + 14, c + 8:c + 14] += W
            H[p + 8:p + 14, c + 8:c + 14] -= W; H[c + 8:c + 14, p + 8:p + 14] -= W
            b[c:c + 29] += rng.standard_normal(29) * 0.1
        self.H_imu, self.b_imu = H, b
        rows = []
        for f in range(1, nf):
            n = 6 if f < nf - 1 else 3
            J = np.zeros((n, self.dim))
            c, p = 5 + 29 * f, 5 + 29 * (f - 1)
            J[:3, p + 3:p + 6] = -np.eye(3); J[:3, c + 3:c + 6] = np.eye(3); J[:3, c + 14:c + 17] = -0.05 * np.eye(3)
            if n == 6:
                J[3:, p:p + 3] = 10 * np.eye(3); J[3:, c:c + 3] = -20 * np.eye(3); J[3:, c + 29:c + 32] = 10 * np.eye(3); J[3:, c + 17:c + 20] = -0.05 * np.eye(3)
            rows.append(J)
        self.J_cst = np.concatenate(rows)
        self.r_cst = rng.standard_normal(len(self.J_cst)) * 1e-4
        # scale optimisation on: the scale column is not a state; frame 0 has no valid spline: only its 8 + 6 bias states stay (:1113-1123)
        self.keep = np.array([i for i in range(self.dim) if i != 4 and not (5 + 14 <= i < 5 + 29)])
        self.start = np.cumsum([4] + [14] + [29] * (nf - 2))        # first compacted column of every frame

    def expand(self, H, b):
        nf, He, be = self.nf, np.zeros((self.dim, self.dim)), np.zeros(self.dim)
        idx = np.concatenate([np.arange(4)] + [5 + 29 * f + np.arange(8) for f in range(nf)])
        He[np.ix_(idx, idx)] = H
        be[idx] = b
        return He, be

    def solve(self, sys, lam=1e-5):
        H, b = self.expand(sys["H_top"], sys["b_top"])
        Hs, bs = self.expand(sys["H_sc"], sys["b_sc"])
        H += self.H_imu; b = b + self.b_imu
        H[np.diag_indices_from(H)] *= (1 + lam)
        H -= Hs * float(np.float32(1.0) / np.float32(1 + lam)); b = b - bs
        H, b = H[np.ix_(self.keep, self.keep)], b[self.keep]
        J = self.J_cst[:, self.keep]
        n, c = len(b), len(J)
        K = np.zeros((n + c, n + c)); K[:n, :n] = H; K[:n, n:] = J.T; K[n:, :n] = J
        rhs = np.concatenate([b, self.r_cst])
        S = 1.0 / np.sqrt(np.diag(K) + 10.0)
        x = S * np.linalg.solve(S[:, None] * K * S[None, :], S * rhs)
        xd = np.zeros(4 + 8 * self.nf)
        xd[:4] = x[:4]
        for f in range(self.nf):
            xd[4 + 8 * f:12 + 8 * f] = x[self.start[f]:self.start[f] + 8]
        return xd


def run_imu_workload(args):
    """configs[2]: one step = FullSystem::optimize(6) with the solve on the host (sosba_ba_system -> widened KKT solve -> sosba_ba_step
    per iteration).  Host wall clock per step for both arms (the host solve is part of the step)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ref = args.impl == "reference"
    if ref and rank != 0:
        return
    pkg, binding, problem, synth = load()
    factor = args.points_factor or (max(1, args.gpus) if args.scaling == "weak" else 1)
    sc = get_scene(synth, factor)
    imu = ImuStandIn(sc.nf)
    dist = None
    if ref:
        lib = binding.Lib(os.path.join(ROOT, "oracle", "_build", "liborc_speed.so"), "orc")
        cfg = forced_cfg(lib, sc, threads=os.cpu_count() or 1)
        h = binding.Handle(lib, cfg)
    else:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        lib = pkg.load()
        cfg = forced_cfg(lib, sc)
        h = binding.Handle(lib, cfg, device=local)
    for i, img in enumerate(sc.images):
        h.frame_make_images(i, img)
    val, val0 = problem.calib_of(sc)
    pts, res = problem.points_of(sc), problem.residuals_of(sc)
    if not ref and world > 1:
        uid = [h.lib_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.comm_init(uid[0], rank, world)
        p0, p1 = problem.shard_points(sc.res_point, sc.n_points, world)[rank]
        pts, res = problem.shard_scene_arrays(pts, res, p0, p1)
    P, keep = h.make_problem(problem.frames_of(sc), val, val0, pts, res)

    def step():
        h.ba_upload(P)
        h.reset_oob()
        lo = h.linearize_all(False)
        h.apply_res()
        nres, t_solve = lo["n_in"] + lo["n_oob"] + lo["n_outlier"], 0.0
        for _ in range(ITERS):
            sy = h.ba_system()
            t0 = time.perf_counter()
            x = imu.solve(sy)
            t_solve += time.perf_counter() - t0
            so = h.ba_step(x)
        return nres * 8 * (ITERS + 1), so["energy"], t_solve

    def barrier():
        if dist is not None and world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    steps = args.steps if not ref else min(args.steps, 20)
    t0 = time.perf_counter()
    n_res, t_solve = 0, 0.0
    for _ in range(steps):
        n, e, ts = step()
        n_res += n; t_solve += ts
    barrier()
    dt = time.perf_counter() - t0
    if not ref and world > 1:
        import torch
        t = torch.tensor([dt, float(n_res)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dt, n_res = float(tmax[0]), float(tsum[1])
    if rank == 0:
        v = n_res / dt
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus if ref else world, "steps": steps, "warmup": max(args.warmup, 3),
                "ms_per_step": 1e3 * dt / steps, "gn_iter_ms": 1e3 * dt / steps / ITERS, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32 (system f64)", "data": "synthetic", "config": config_of(sc, factor, args.scaling),
                "host_solve_ms_per_step": 1e3 * t_solve / steps, "system_dim": int(imu.dim - 1 + len(imu.J_cst)),
                "timing": "host wall clock: the widened solve runs on the host inside every iteration (2 synchronisations per iteration)",
                "final_energy": e,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0 if ref else int(sum(np.asarray(a).nbytes for a in pts.values()) + sum(np.asarray(a).nbytes for a in res.values())),
                        "d2h_bytes_per_step": 0 if ref else int(ITERS * 2 * 8 * ((4 + 8 * sc.nf) ** 2 + 4 + 8 * sc.nf))}}
        if ref:
            line["impl"] = "reference"
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": f"{steps} steps, oracle speed build + the same host solve"}
        print(json.dumps(line), flush=True)
    h.close()
    if dist is not None and world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sosba")
    ap.add_argument("--points-factor", type=int, default=0, help="scaling sweep: multiply the 2000 points (default: = gpus)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = N x the workload's points (2000 per GPU), strong = the workload's points split over the N GPUs")
    ap.add_argument("--no-other", action="store_true", help="skip the per-call timings of the rows next to the BA loop")
    ap.add_argument("--workload", default="configB", choices=sorted(WORKLOADS), help="configB = the benchmark workload (BASELINE.json configs[1]); others are extra data points")
    args = ap.parse_args()
    global _workload, WORKLOAD
    _workload = args.workload
    WORKLOAD = WORKLOADS[_workload]
    if _workload == "euroc_imu":
        return run_imu_workload(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsosba has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg, binding, problem, synth = load()
    lib = pkg.load()
    factor = args.points_factor or (max(1, world) if args.scaling == "weak" else 1)
    sc = get_scene(synth, factor)
    cfg = forced_cfg(lib, sc)
    h = binding.Handle(lib, cfg, device=local)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    for i, img in enumerate(sc.images):
        h.frame_make_images(i, img)
    val, val0 = problem.calib_of(sc)
    pts, res = problem.points_of(sc), problem.residuals_of(sc)
    if world > 1:
        uid = [None]
        if rank == 0:
            uid[0] = h.lib_unique_id()
        dist.broadcast_object_list(uid, src=0)
        h.comm_init(uid[0], rank, world)
        p0, p1 = problem.shard_points(sc.res_point, sc.n_points, world)[rank]
        pts, res = problem.shard_scene_arrays(pts, res, p0, p1)
    P, keep = h.make_problem(problem.frames_of(sc), val, val0, pts, res)
    h.ba_upload(P)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: resident window ------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        out = h.ba_optimize(ITERS)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    h.profile_enable(1)
    l0 = h.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    nres_local = 0
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)           # evict the window from L2 (126 MB) between steps
        # the per-launch CUDA-event brackets of the roofline kernel ride on every PROF_EVERY-th timed step: a bracket breaks the
        # programmatic launch chain on both sides of the kernel (about 4 us per launch), which is not part of the product
        h.profile_enable(2 if k % PROF_EVERY == 0 else 0)
        ev[k][0].record(stream)
        out = h.ba_optimize(ITERS)
        ev[k][1].record(stream)
        nres_local += out["reserved0"] * 8 * (out["iterations"] + 2)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = h.launch_count() - l0
    lin_ms, lin_n = h.profile_read()
    h.profile_enable(0)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    tot_ms = sum(step_ms)
    clocks = sampler.stop()
    t = torch.tensor([tot_ms, float(nres_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tot_ms, nres = float(tmax[0]), float(tsum[1])
    else:
        nres = float(nres_local)
    value = nres / (tot_ms * 1e-3)

    # ---- GN-iteration ms: the loop body alone, warm L2 ------------------------------------------------
    n_it = 30
    h.ba_iterate(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    h.ba_iterate(n_it)
    e1.record(stream)
    barrier()
    gn_ms = e0.elapsed_time(e1) / n_it

    # ---- e2e: host buffers through the public C ABI ---------------------------------------------------
    frames = problem.frames_of(sc)
    h2d = sc.images[-1].nbytes + sum(np.asarray(v).nbytes for v in pts.values()) + sum(np.asarray(v).nbytes for v in res.values()) + len(frames) * 8 * 32
    d2h = len(frames) * 8 * 32 + 4 * len(pts["u"])

    # The caller's problem struct (sosba_ba_problem over its own host arrays) is built once, as FullSystem would keep
    # it; every step restores the in/out members (frame states, calibration) and makes the two calls a user makes.
    import ctypes
    Pn, kn = h.make_problem(frames, val, val0, pts, res)
    frames0 = bytes(ctypes.string_at(ctypes.addressof(kn[0]), ctypes.sizeof(kn[0])))
    calib0 = list(Pn.calib_value)

    # the step's largest input, the newest keyframe's image, sits in pinned host memory (the e2e contract) as a camera driver's
    # DMA buffer would: the library reads it in place instead of staging a pageable copy
    img_pin = torch.from_numpy(sc.images[-1].copy()).pin_memory()
    img_new = img_pin.numpy()
    raw8 = np.clip(np.rint(sc.images[-1]), 0, 255).astype(np.uint8)      # the camera frame as it arrives (8f rank 2 input)
    raw_pin = torch.from_numpy(raw8.copy()).pin_memory()
    raw8 = raw_pin.numpy()

    def e2e_step(raw=False):
        ctypes.memmove(ctypes.addressof(kn[0]), frames0, len(frames0))
        for i in range(4):
            Pn.calib_value[i] = calib0[i]
        if raw:
            h.frame_make_images_raw(sc.nf - 1, raw8)
        else:
            h.frame_make_images(sc.nf - 1, img_new)
        return h.optimize(Pn, ITERS)

    def e2e_loop(raw):
        for _ in range(3):
            e2e_step(raw)
        barrier()
        ms, nres = 0.0, 0
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            tw = time.perf_counter()
            a.record(stream)
            o = e2e_step(raw)
            b.record(stream)
            torch.cuda.synchronize()
            ms += max(a.elapsed_time(b), 1e3 * (time.perf_counter() - tw))   # device timeline and host wall agree; take the larger
            nres += o["reserved0"] * 8 * (o["iterations"] + 2)
        t = torch.tensor([ms, float(nres)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return float(tmax[0]), float(tsum[1])
        return ms, float(nres)

    e2e_ms, e2e_res = e2e_loop(False)
    e2e_value = e2e_res / (e2e_ms * 1e-3)
    # the same step fed with the 8-bit camera frame (passthrough rectification, no photometric calibration, factor 1):
    # reported next to the headline e2e, which keeps the float irradiance image of FrameHessian::makeImages as its input
    h.undistort_set(sc.w, sc.h)
    raw_ms, raw_res = e2e_loop(True)
    e2e_raw = {"value": raw_res / (raw_ms * 1e-3), "unit": UNIT, "ms_per_step": raw_ms / args.steps,
               "h2d_bytes_per_step": int(h2d - sc.images[-1].nbytes + raw8.nbytes),
               "note": "newest keyframe enters as the 8-bit camera frame (sosba_frame_make_images_raw) instead of the float irradiance image"}
    h.frame_make_images(sc.nf - 1, sc.images[-1])

    # ---- the other kernels of the path: per-call time, host buffers in ------------
    other = None
    if rank == 0 and not args.no_other:
        try:
            reps = 30
            t0 = time.perf_counter()
            for _ in range(reps):
                h.frame_make_images(sc.nf, sc.images[-1])      # a1: H2D of the irradiance image + the 4-level pyramid
            h.synchronize()
            pyr_ms = 1e3 * (time.perf_counter() - t0) / reps
            # 8f rank 2: raw 8-bit frame in, photometric + geometric undistortion and the pyramid on the device
            uc = synth.undistort_case(sc.w, sc.h)
            h.undistort_set(uc["w_org"], uc["h_org"], uc["remapX"], uc["remapY"], uc["G"], uc["vignette_inv"])
            h.frame_make_images_raw(sc.nf, uc["raw"])
            h.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                h.frame_make_images_raw(sc.nf, uc["raw"])
            h.synchronize()
            raw_ms = 1e3 * (time.perf_counter() - t0) / reps
            h.frame_make_images(sc.nf, sc.images[-1])
            # 8f rank 3: PixelSelector::makeMaps on the newest keyframe, density 1500 (setting_desiredImmatureDensity)
            h.pixel_selector_set(np.random.default_rng(3141592).integers(0, 256, sc.w * sc.h).astype(np.uint8), 3)
            h.pixel_select(sc.nf - 1, 1500.0, want_map=False)
            sel_ms = []
            for k in range(10):
                t0 = time.perf_counter()
                sel = h.pixel_select(k % sc.nf, 1500.0, want_map=False)
                sel_ms.append(1e3 * (time.perf_counter() - t0))
            rng = np.random.default_rng(5)
            n_ref = 10000
            K = sc.K.astype(np.float32)
            h.tracker_make_k(K)
            u = rng.integers(2, sc.w - 2, n_ref).astype(np.float32)
            v = rng.integers(2, sc.h - 2, n_ref).astype(np.float32)
            idp = rng.uniform(0.35, 0.65, n_ref).astype(np.float32)
            col = sc.images[0][v.astype(int), u.astype(int)].astype(np.float32)
            h.tracker_set_ref(0, u, v, idp, col)
            T = (np.linalg.inv(sc.camToWorld_true[1]) @ sc.camToWorld_true[0])[:3, :4]
            h.scale_set_stereo(T, K)
            h.tracker_calc_res_pose(0, 1, T, (1.0, 0.0), 20.0)
            h.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                h.tracker_calc_res_pose(0, 1, T, (1.0, 0.0), 20.0)      # a14 (returns the Vec6 to the host: one sync per call, as the LM loop needs)
                h.tracker_calc_gs_pose(0, 1.0, 0.0)                      # a15
            trk_ms = 1e3 * (time.perf_counter() - t0) / reps
            t0 = time.perf_counter()
            for _ in range(reps):
                h.scale_calc_res(0, 1, 1.0, 20.0)                        # a17
                h.scale_calc_gs(0, 1.0)
            scl_ms = 1e3 * (time.perf_counter() - t0) / reps
            # a16 / a17: the direct-alignment control loops resident on the device (one launch, one synchronisation per call)
            lm = lm_timings(h, sc, synth, reps=10)
            # 8f rank 1: traceNewCoarse of 2000 immature points per older keyframe into the newest one (host buffers in and out)
            case = synth.trace_case(sc, sc.nf - 1, n_per_host=2000, seed=3)
            ip_parts = [h.immature_init(hst, case["u"][case["host"] == hst], case["v"][case["host"] == hst]) for hst in range(sc.nf - 1)]
            ip = {k: np.concatenate([q[k] for q in ip_parts]) for k in ip_parts[0]}
            trace_ms, trace_counts = [], None
            for _ in range(5):
                ipc = {k: x.copy() for k, x in ip.items()}
                t0 = time.perf_counter()
                trace_counts = h.trace_immature(sc.nf - 1, case["host"], case["KRKi"], case["Kt"], case["aff"], ipc)
                trace_ms.append(1e3 * (time.perf_counter() - t0))
            # the same trace with the points resident in HBM (sosba_immature_pool_*): only the per-host tables go up
            h.immature_pool_set(case["host"], {k: x.copy() for k, x in ip.items()})
            h.immature_pool_trace(sc.nf - 1, case["KRKi"], case["Kt"], case["aff"])
            pool_ms = []
            for _ in range(5):
                h.immature_pool_set(case["host"], {k: x.copy() for k, x in ip.items()})
                h.synchronize()
                t0 = time.perf_counter()
                h.immature_pool_trace(sc.nf - 1, case["KRKi"], case["Kt"], case["aff"])
                pool_ms.append(1e3 * (time.perf_counter() - t0))
            okp = np.isfinite(ipc["idepth_max"])
            sub = {k: x[okp] for k, x in ipc.items()}
            win = synth.activation_case(sc)
            act_ms = []
            for _ in range(5):
                t0 = time.perf_counter()
                act = h.optimize_immature(np.arange(sc.nf), win["RTll"], win["tTll"], win["aff"], win["calib"], case["host"][okp], sub)
                act_ms.append(1e3 * (time.perf_counter() - t0))
            other = {"optimize_immature_ms": float(np.median(act_ms)), "optimize_immature_points": int(okp.sum()),
                     "optimize_immature_activated": int((act[0] == 1).sum()), "pixel_select_ms": float(np.median(sel_ms)), "pixel_select_n": int(sel["n"]), "make_images_raw_ms": raw_ms,
                     "make_images_raw_note": f"{uc['w_org']}x{uc['h_org']} 8-bit raw frame in: H2D + response/vignette + rectification + 4 levels, host wall per call",
                     "make_images_ms": pyr_ms, "make_images_note": f"{sc.w}x{sc.h}, H2D + 4 levels, host wall per call",
                     "trace_immature_ms": float(np.median(trace_ms)), "trace_immature_resident_ms": float(np.median(pool_ms)), "trace_immature_points": int(case["host"].size),
                     "trace_immature_counts": [int(x) for x in trace_counts],
                     "trace_note": "first trace (unbounded interval: the longest epipolar search), host SoA in and out, host wall per call",
                     "tracker_calcRes_plus_calcGS_ms": trk_ms, "scale_calcRes_plus_calcGS_ms": scl_ms,
                     "tracker_note": f"level 0, {n_ref} reference points, results returned to the host each call (host wall)",
                     "direct_alignment_loops": lm}
        except Exception as ex:   # the BA numbers above do not depend on this block
            other = {"error": str(ex)}

    # ---- point shards against one GPU on the same window (pytest -m gpu runs on a 1-GPU box, so the check lives here) ----
    shard_parity = None
    if world > 1:
        def final_of(hh, Pp, kk):
            o = hh.optimize(Pp, ITERS)
            return o, hh.problem_result(Pp, kk)
        Ps, ks = h.make_problem(frames, val, val0, pts, res)
        o_s, r_s = final_of(h, Ps, ks)
        t = torch.tensor(np.concatenate([r_s["state"].ravel(), r_s["frame_energy_th"].astype(np.float64), [o_s["iterations"], o_s["energy_final"], o_s["res_in_a"]]]), device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ranks_identical = bool((hi == lo).all().item())
        if rank == 0:
            h1 = binding.Handle(lib, cfg, device=local)
            for i, img in enumerate(sc.images):
                h1.frame_make_images(i, img)
            P1, k1 = h1.make_problem(frames, val, val0, problem.points_of(sc), problem.residuals_of(sc))
            o_1, r_1 = final_of(h1, P1, k1)
            h1.close()
            upd = np.abs(r_1["state"] - sc.state).max(axis=0) + 1e-12
            dev = float(np.max(np.abs(r_s["state"] - r_1["state"]).max(axis=0) / upd))
            p0, p1 = problem.shard_points(sc.res_point, sc.n_points, world)[0]
            did = float(np.max(np.abs(r_s["idepth"] - r_1["idepth"][p0:p1]) / (np.abs(r_1["idepth"][p0:p1]) + 1e-3)))
            shard_parity = {"ranks_bit_identical": ranks_identical, "iterations": [int(o_s["iterations"]), int(o_1["iterations"])],
                            "res_in_a": [int(o_s["res_in_a"]), int(o_1["res_in_a"])], "n_removed": [int(o_s["n_removed"]), int(o_1["n_removed"])],
                            "energy_final_rel": abs(o_s["energy_final"] - o_1["energy_final"]) / abs(o_1["energy_final"]),
                            "state_dev_over_update": dev, "idepth_rel_rank0": did,
                            "frame_energy_th_rel": float(np.max(np.abs(r_s["frame_energy_th"] - r_1["frame_energy_th"]) / np.abs(r_1["frame_energy_th"]))),
                            "reference": "the same window optimised on one GPU (rank 0, all points)"}
            shard_parity["ok"] = bool(ranks_identical and o_s["iterations"] == o_1["iterations"] and abs(o_s["res_in_a"] - o_1["res_in_a"]) <= 2
                                      and shard_parity["energy_final_rel"] < 1e-3 and dev < 5e-3 and did < 2e-3 and shard_parity["frame_energy_th_rel"] < 1e-3)
        barrier()

    # ---- the same launches inside the programmatic launch chain: device-side timeline (globaltimer stamps in the kernels) ----
    timeline = None
    try:
        h.trace_enable(True)
        tl = {}
        for _ in range(3):
            flush.fill_(7)
            h.ba_optimize(ITERS)
            for kname in ("k_linearize", "k_accumulate_fused", "k_stitch_xchg", "k_solve", "k_step (points)"):
                ns, n = h.trace_read(kname, 1 if kname == "k_linearize" else 0)
                if n:
                    tl.setdefault(kname, []).append(ns * 1e-3)
        h.trace_enable(False)
        timeline = {k.replace(" (points)", ""): float(np.mean(v)) for k, v in tl.items()}
    except Exception as e:   # the timeline is a diagnostic, never a reason to lose the bench line
        timeline = {"error": str(e)}
    barrier()

    # ---- roofline of the dominant kernel (linearize) ---------------------------------------------------
    peak, peak_src = measured_peak()
    R_lin = out["reserved0"]
    lin_us = 1e3 * lin_ms / max(lin_n, 1)
    achieved = BYTES_LINEARIZE * R_lin / (lin_us * 1e-6) / 1e9 if lin_n else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_linearize.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_linearize", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "launch_us": lin_us, "launches_timed": lin_n,
                "algorithmic_bytes_per_launch": BYTES_LINEARIZE * R_lin,
                "note": "fused linearize+applyRes launches (7 of the 8 linearisations of a step), CUDA events around each launch on every "
                        f"{PROF_EVERY}th timed step; the bracket itself costs the launch latency the programmatic launch chain otherwise hides, "
                        "so in_chain_us (device timeline, first CTA past its dependency wait .. last CTA done) is given next to it; 8.4 MB per "
                        "launch = one partial wave: latency-bound at this size (DESIGN.md section 5; points sweep in profiles/)"}
    if timeline and "k_linearize" in timeline:
        roofline["in_chain_us"] = timeline["k_linearize"]
        roofline["in_chain_frac"] = BYTES_LINEARIZE * R_lin / (timeline["k_linearize"] * 1e-6) / 1e9 / peak

    # ---- CPU baseline (rank 0, bounded sample) ---------------------------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu:
        opath = os.path.join(ROOT, "oracle", "_build", "liborc_speed.so")
        if os.path.exists(opath):
            olib = binding.Lib(opath, "orc")
            cores = os.cpu_count() or 1
            sc1 = get_scene(synth, 1)
            oh = binding.Handle(olib, forced_cfg(olib, sc1, threads=cores))
            for i, img in enumerate(sc1.images):
                oh.frame_make_images(i, img)
            v1, v10 = problem.calib_of(sc1)
            n, tcpu, rcpu = 0, 0.0, 0
            while n < 3 or (tcpu < 8.0 and n < 60):
                Pn, kn = oh.make_problem(problem.frames_of(sc1), v1, v10, problem.points_of(sc1), problem.residuals_of(sc1))
                t0 = time.perf_counter()
                o = oh.optimize(Pn, ITERS)
                dt = time.perf_counter() - t0
                if n >= 1:
                    tcpu += dt; rcpu += o["reserved0"] * 8 * (o["iterations"] + 2)
                n += 1
            cpu_raw_ms = None
            try:
                uc = synth.undistort_case(sc1.w, sc1.h)
                oh.undistort_set(uc["w_org"], uc["h_org"], uc["remapX"], uc["remapY"], uc["G"], uc["vignette_inv"])
                oh.frame_make_images_raw(sc1.nf, uc["raw"])
                t0 = time.perf_counter()
                for _ in range(5):
                    oh.frame_make_images_raw(sc1.nf, uc["raw"])
                cpu_raw_ms = 1e3 * (time.perf_counter() - t0) / 5
            except Exception:
                pass
            cpu_sel_ms = None
            try:
                oh.pixel_selector_set(np.random.default_rng(3141592).integers(0, 256, sc1.w * sc1.h).astype(np.uint8), 3)
                oh.pixel_select(sc1.nf - 1, 1500.0, want_map=False)
                ts = []
                for k in range(5):
                    t0 = time.perf_counter()
                    oh.pixel_select(k % sc1.nf, 1500.0, want_map=False)
                    ts.append(1e3 * (time.perf_counter() - t0))
                cpu_sel_ms = float(np.median(ts))
            except Exception:
                pass
            # the 8f rank-1 row on the host cores: traceNewCoarse is a serial loop in the reference (FullSystem.cpp:311-361)
            cpu_trace = None
            try:
                case = synth.trace_case(sc1, sc1.nf - 1, n_per_host=2000, seed=3)
                parts = [oh.immature_init(hst, case["u"][case["host"] == hst], case["v"][case["host"] == hst]) for hst in range(sc1.nf - 1)]
                ip = {k: np.concatenate([q[k] for q in parts]) for k in parts[0]}
                t0 = time.perf_counter()
                oh.trace_immature(sc1.nf - 1, case["host"], case["KRKi"], case["Kt"], case["aff"], ip)
                t_tr = time.perf_counter() - t0
                okp = np.isfinite(ip["idepth_max"])
                win = synth.activation_case(sc1)
                sub = {k: x[okp] for k, x in ip.items()}
                t0 = time.perf_counter()
                oh.optimize_immature(np.arange(sc1.nf), win["RTll"], win["tTll"], win["aff"], win["calib"], case["host"][okp], sub)
                t_act = time.perf_counter() - t0
                cpu_trace = {"pixel_select_ms": cpu_sel_ms, "make_images_raw_ms": cpu_raw_ms, "trace_immature_ms": 1e3 * t_tr, "optimize_immature_ms": 1e3 * t_act, "cores": 1,
                             "note": "same inputs as other_kernels; single thread (the reference's trace loop is serial, its activation loop threaded)"}
            except Exception as ex:
                cpu_trace = {"error": str(ex)}
            oh.close()
            try:
                if cpu_trace is not None and "error" not in cpu_trace:
                    cpu_trace["direct_alignment_loops"] = lm_timings(oh2 := binding.Handle(olib, forced_cfg(olib, sc1, threads=1)), sc1, synth, reps=2, make_images=True)
                    oh2.close()
            except Exception as ex:
                cpu_trace["direct_alignment_loops"] = {"error": str(ex)}
            cpu = {"other_kernels": cpu_trace, "value": rcpu / tcpu, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n - 1} optimize() steps of the same window ({tcpu:.1f} s), oracle speed build, {cores} IndexThreadReduce workers",
                   "ms_per_step": 1e3 * tcpu / (n - 1)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": tot_ms / args.steps, "gn_iter_ms": gn_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": config_of(sc, factor, args.scaling),
                "sharding": {"parallelism": f"points/{world}", "points_per_rank": int(len(pts["u"])),
                             "exchange": "none" if world == 1 else ("stitch fused with a peer-memory push over NVLink, one launch per GN iteration (k_xchg.cu)"
                                                                    if h.comm_uses_peer_memory() else "nccl all-reduce of the block tables")},
                "shard_parity": shard_parity,
                "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                                          "ms_per_step": e2e_ms / args.steps},
                "e2e_raw_frame": e2e_raw, "gpu_launches": int(launches), "roofline": roofline, "device_timeline_us": timeline, "cpu_baseline": cpu, "other_kernels": other,
                "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps, "final_rmse": out["rmse"]}
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
