"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: bit-exact for integer bookkeeping (residual states, counts, removal sets, pyramid
levels); per-residual floats rel 1e-5; block accumulators / stitched H,b rel 1e-4 of the Frobenius norm;
solved step rel 1e-3; tracker H,b rel 1e-4; immature points (constructor + traceOn) bit-exact in every field."""
import numpy as np
import pytest

from _scenes import EUROC, TUMVI, CONFIG_B, KITTI, SMALL, open_handle, relerr, scene, trace_points, upload

pytestmark = pytest.mark.gpu


def both(gpu, orc, sc, **kw):
    out = []
    for lib in (gpu, orc):
        h = open_handle(lib, sc)
        P, keep = upload(h, sc, **kw)
        out.append((h, P, keep))
    return out


# ---- a1 -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(640, 480), (752, 480), (1232, 368), (512, 512), (322, 242)])
def test_pyramid_bit_exact(gpu, orc, shape):
    w, h = shape
    rng = np.random.default_rng(w * 7 + h)
    img = (rng.uniform(0, 255, (h, w)) * 0.3 + 0.7 * 127 * (1 + np.sin(np.arange(w)[None, :] * 0.07) * np.cos(np.arange(h)[:, None] * 0.05))).astype(np.float32)
    B = np.linspace(0, 255, 256).astype(np.float32) ** 1.02
    for useB in (False, True):
        res = []
        for lib in (gpu, orc):
            cfg = lib.config_default(w, h)
            cfg.max_frames = 2
            from sos_slam_b200 import binding
            hd = binding.Handle(lib, cfg)
            hd.frame_make_images(1, img, B if useB else None)
            res.append([hd.frame_get_level(1, l) for l in range(hd.levels)])
            res[-1].append(hd.levels)
            hd.close()
        assert res[0][-1] == res[1][-1]
        for l in range(res[0][-1]):
            assert np.array_equal(res[0][l][0], res[1][l][0]), f"dI level {l}"
            assert np.array_equal(res[0][l][1], res[1][l][1]), f"absSquaredGrad level {l}"


# ---- a3/a4/a5 -----------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B, KITTI], ids=["small", "configB", "kitti"])
def test_linearize_apply(gpu, orc, cfg):
    sc = scene(**cfg)
    (hg, Pg, kg), (ho, Po, ko) = both(gpu, orc, sc)
    for h in (hg, ho):
        h.reset_oob()
    lg, lo = hg.linearize_all(False), ho.linearize_all(False)
    for k in ("n_in", "n_oob", "n_outlier", "n_removed"):
        assert lg[k] == lo[k], (k, lg, lo)
    assert lg["new_frame_energy_th"] == lo["new_frame_energy_th"]
    assert abs(lg["energy"] - lo["energy"]) <= 1e-9 * abs(lo["energy"])
    sg, so = hg.get_state(), ho.get_state()
    assert np.array_equal(sg["new_state"], so["new_state"])
    live = so["new_state"] != 1
    assert np.array_equal(sg["new_energy"][live], so["new_energy"][live])
    assert np.array_equal(sg["new_energy_wo"], so["new_energy_wo"])
    Jg, Jo = hg.get_jacobians(False), ho.get_jacobians(False)
    assert np.array_equal(Jg[live], Jo[live])          # same op order, no FMA: bit-exact
    ag, ao = hg.get_aux(), ho.get_aux()
    assert np.array_equal(ag["projectedTo"][live], ao["projectedTo"][live])
    assert np.array_equal(ag["centerProjectedTo"][live], ao["centerProjectedTo"][live])
    for h in (hg, ho):
        h.apply_res()
    sg, so = hg.get_state(), ho.get_state()
    for k in ("state", "is_active"):
        assert np.array_equal(sg[k], so[k]), k
    act = so["is_active"] == 1
    assert np.array_equal(hg.get_jacobians(True)[act], ho.get_jacobians(True)[act])
    ag, ao = hg.get_aux(), ho.get_aux()
    assert np.allclose(ag["JpJdF"][act], ao["JpJdF"][act], rtol=1e-5, atol=1e-6 * np.abs(ao["JpJdF"][act]).max())
    # second linearisation from the committed state and the new threshold
    lg, lo = hg.linearize_all(True), ho.linearize_all(True)
    for k in ("n_in", "n_oob", "n_outlier", "n_removed"):
        assert lg[k] == lo[k], (k, lg, lo)
    mg, ng = hg.points_get_stats()
    mo, no = ho.points_get_stats()
    assert np.array_equal(ng, no)
    assert np.allclose(mg, mo, rtol=1e-6)
    hg.close(); ho.close()


# ---- a6-a11 -------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B, KITTI], ids=["small", "configB", "kitti"])
def test_accumulate_solve(gpu, orc, cfg):
    sc = scene(**cfg)
    D = 4 + 8 * sc.nf
    rng = np.random.default_rng(11)
    A = rng.normal(size=(D, D))
    HM = (A @ A.T) * 50.0
    bM = rng.normal(size=D) * 10.0
    (hg, Pg, kg), (ho, Po, ko) = both(gpu, orc, sc, calib_delta=(1e-4, -2e-4, 3e-4, 1e-4))
    for h in (hg, ho):
        h.reset_oob(); h.linearize_all(False); h.apply_res()
    ag, ao = hg.accumulate(), ho.accumulate()
    assert ag["resInA"] == ao["resInA"] and ag["resInL"] == ao["resInL"]
    for k in ("HA", "bA", "HL", "bL", "Hsc", "bsc"):
        assert relerr(ag[k], ao[k]) < 1e-4, (k, relerr(ag[k], ao[k]))
    assert np.allclose(ag["HA"], ag["HA"].T, rtol=0, atol=1e-9 * np.abs(ag["HA"]).max())
    assert np.allclose(ag["Hsc"], ag["Hsc"].T, rtol=0, atol=1e-9 * np.abs(ag["Hsc"]).max())
    pg, po = hg.points_get_acc(), ho.points_get_acc()
    for k in pg:
        assert np.allclose(pg[k], po[k], rtol=2e-5, atol=2e-6 * max(1e-30, np.abs(po[k]).max())), k
    for (hm, bm) in ((None, None), (HM, bM)):
        xg, Hg, bg = hg.solve_system(hm, bm)
        xo, Ho, bo = ho.solve_system(hm, bm)
        assert relerr(Hg, Ho) < 1e-4 and relerr(bg, bo) < 1e-4
        # compare the step in the H-norm (the system is ill-conditioned along the gauge)
        d = xg - xo
        assert np.sqrt(abs(d @ Ho @ d)) <= 1e-3 * np.sqrt(abs(xo @ Ho @ xo)), (relerr(xg, xo))
        assert relerr(xg, xo) < 5e-3
        x = xo
        stg, sto = hg.resubstitute(x), ho.resubstitute(x)
        assert np.allclose(stg, sto, rtol=2e-4, atol=2e-6 * np.abs(sto).max())
    hg.close(); ho.close()


def test_linearized_and_marginalize(gpu, orc):
    """fixLinearizationF (a13), addPoint<1> with deltas, marginalizePointsF (addPoint<2> + SC(false))."""
    sc = scene(**SMALL)
    (hg, Pg, kg), (ho, Po, ko) = both(gpu, orc, sc, calib_delta=(2e-4, 1e-4, -1e-4, 3e-4))
    counts = np.bincount(sc.res_point, minlength=sc.n_points)
    starts = np.concatenate([[0], np.cumsum(counts)])
    pts = np.arange(0, sc.n_points, 7, dtype=np.int32)
    rids = np.concatenate([np.arange(starts[p], starts[p + 1]) for p in pts]).astype(np.int32)
    for h in (hg, ho):
        h.reset_oob(); h.linearize_all(False); h.apply_res()
    act = ho.get_state()["is_active"]
    rids = rids[act[rids] == 1]
    for h in (hg, ho):
        h.fix_linearization(rids)
    ag, ao = hg.get_aux(), ho.get_aux()
    assert np.array_equal(ag["res_toZeroF"][rids], ao["res_toZeroF"][rids])
    assert np.array_equal(hg.get_state()["is_linearized"], ho.get_state()["is_linearized"])
    ag, ao = hg.accumulate(), ho.accumulate()
    assert ag["resInA"] == ao["resInA"] and ag["resInL"] == ao["resInL"] and ao["resInL"] == len(rids)
    for k in ("HA", "bA", "HL", "bL", "Hsc", "bsc"):
        assert relerr(ag[k], ao[k]) < 1e-4, (k, relerr(ag[k], ao[k]))
    Hg, bg, ng = hg.marginalize_points(pts)
    Ho, bo, no = ho.marginalize_points(pts)
    assert ng == no
    assert relerr(Hg, Ho) < 1e-4 and relerr(bg, bo) < 1e-4
    hg.close(); ho.close()


# ---- the composed Gauss-Newton loop -------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B, KITTI, EUROC, TUMVI], ids=["small", "configB", "kitti-12kf", "euroc-752x480", "tumvi-512x512"])
def test_optimize(gpu, orc, cfg):
    """FullSystem::optimize(6) on every window shape BASELINE.json names: 640x480 (configs[1]), 752x480 with the EuRoC intrinsics
    (configs[2], IMU off), 1232x368 with 12 keyframes (configs[3]: the k_solve<7> instantiation, D = 100), 512x512 (configs[4])."""
    sc = scene(**cfg)
    res = []
    for lib in (gpu, orc):
        h = open_handle(lib, sc)
        from sos_slam_b200 import problem
        val, val0 = problem.calib_of(sc)
        P, keep = h.make_problem(problem.frames_of(sc), val, val0, problem.points_of(sc), problem.residuals_of(sc))
        out = h.optimize(P, 6)
        res.append((out, h.problem_result(P, keep), h.get_state()))
        h.close()
    (og, rg, sg), (oo, ro, so) = res
    assert og["iterations"] == oo["iterations"]
    assert og["energy_initial"] == pytest.approx(oo["energy_initial"], rel=1e-9)
    assert og["energy_final"] == pytest.approx(oo["energy_final"], rel=2e-3)
    assert og["energy_final"] < 0.2 * og["energy_initial"]
    assert abs(og["res_in_a"] - oo["res_in_a"]) <= 2
    assert int((sg["state"] != so["state"]).sum()) <= max(2, sg["state"].size // 500)
    # the two runs take the same Gauss-Newton path; what differs is fp32 summation order inside H, b
    upd = np.abs(ro["state"] - sc.state).max(axis=0) + 1e-12     # per state component (t, r, a, b live on different scales)
    dev = np.abs(rg["state"] - ro["state"]).max(axis=0)
    print("state deviation / update per component:", dev / upd)
    assert (dev <= 5e-3 * upd).all(), (dev, upd)
    assert np.allclose(rg["idepth"], ro["idepth"], rtol=2e-3, atol=1e-5)
    assert np.allclose(rg["frame_energy_th"], ro["frame_energy_th"], rtol=1e-3)
    # the newest frame moves to its optimised pose: same bar as the states above (5e-3 of the update; translation state x SCALE_XI_TRANS)
    assert np.allclose(rg["evalPT"], ro["evalPT"], rtol=1e-5, atol=max(5e-6, 5e-3 * 0.5 * float(upd[:6].max())))


# ---- a14/a15/a17 ----------------------------------------------------------------------------------
def _ref_points(sc, h, lvl, n, rng):
    w, hh = sc.w >> lvl, sc.h >> lvl
    u = rng.integers(2, w - 2, n).astype(np.float32)
    v = rng.integers(2, hh - 2, n).astype(np.float32)
    idepth = rng.uniform(0.35, 0.65, n).astype(np.float32)
    dI, _ = h.frame_get_level(0, lvl)
    color = dI[v.astype(int), u.astype(int), 0].copy()
    return u, v, idepth, color


@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B], ids=["small", "configB"])
def test_tracker_and_scale(gpu, orc, cfg):
    from sos_slam_b200 import synth
    sc = scene(**cfg)
    hg, ho = open_handle(gpu, sc), open_handle(orc, sc)
    T = (np.linalg.inv(sc.camToWorld_true[1]) @ sc.camToWorld_true[0])[:3, :4]
    T10 = synth.se3_exp([0.1, 0.002, -0.001, 0.001, -0.002, 0.0005])[:3, :4]
    K = sc.K.astype(np.float32)
    for h in (hg, ho):
        h.tracker_make_k(K)
        h.scale_set_stereo(T10, K * np.array([1.01, 0.99, 1.0, 1.0], np.float32))
    for lvl in range(hg.levels):
        rng = np.random.default_rng(100 + lvl)
        n = max(37, 9000 >> (2 * lvl))
        pts = _ref_points(sc, ho, lvl, n, rng)
        for h in (hg, ho):
            h.tracker_set_ref(lvl, *pts)
        for cutoff in (20.0, 6.0):
            og, cg = hg.tracker_calc_res_pose(lvl, 1, T, (1.02, -3.0), cutoff)
            oo, co = ho.tracker_calc_res_pose(lvl, 1, T, (1.02, -3.0), cutoff)
            assert np.array_equal(cg, co), (lvl, cg, co)
            assert np.allclose(og, oo, rtol=2e-5, atol=1e-6), (lvl, og, oo)
            Hg, bg = hg.tracker_calc_gs_pose(lvl, 1.02, 0.004)
            Ho, bo = ho.tracker_calc_gs_pose(lvl, 1.02, 0.004)
            assert relerr(Hg, Ho) < 1e-4 and relerr(bg, bo) < 1e-4
        for s in (1.0, 0.8):
            og, cg = hg.scale_calc_res(lvl, 2, s, 20.0)
            oo, co = ho.scale_calc_res(lvl, 2, s, 20.0)
            assert np.array_equal(cg, co), (lvl, cg, co)
            assert np.allclose(og, oo, rtol=2e-5, atol=1e-6)
            Hg, bg = hg.scale_calc_gs(lvl, s)
            Ho, bo = ho.scale_calc_gs(lvl, s)
            assert Hg == pytest.approx(Ho, rel=1e-4) and bg == pytest.approx(bo, rel=1e-4, abs=1e-6 * abs(Ho))
    hg.close(); ho.close()


# ---- edge cases -----------------------------------------------------------------------------------
def test_edge_cases(gpu, orc):
    from sos_slam_b200 import problem
    sc = scene(**SMALL)
    # (1) points without residuals + a ragged residual list; (2) a window seen from far away: everything OOB
    pts = problem.points_of(sc)
    res = problem.residuals_of(sc)
    keep = (res["point"] % 3) != 0          # every third point loses all its residuals
    res = {k: v[keep] for k, v in res.items()}
    for far in (False, True):
        out = []
        for lib in (gpu, orc):
            h = open_handle(lib, sc)
            frames = problem.frames_of(sc)
            if far:
                for f in frames[1:]:
                    f["evalPT"] = f["evalPT"].copy()
                    f["evalPT"][:3, 3] += 50.0
            val, val0 = problem.calib_of(sc)
            P, k = h.make_problem(frames, val, val0, pts, res)
            h.ba_upload(P)
            h.reset_oob()
            lo = h.linearize_all(False)
            h.apply_res()
            acc = h.accumulate()
            st = h.get_state()
            out.append((lo, acc, st))
            h.close()
        (lg, ag, sg), (lo, ao, so) = out
        for k in ("n_in", "n_oob", "n_outlier"):
            assert lg[k] == lo[k]
        assert np.array_equal(sg["state"], so["state"])
        assert ag["resInA"] == ao["resInA"]
        if far:
            assert lo["n_in"] <= 3
        for k in ("HA", "HL", "Hsc"):
            assert relerr(ag[k], ao[k]) < 1e-4
    # (3) empty residual set
    for lib in (gpu,):
        h = open_handle(lib, sc)
        empty = {k: v[:0] for k, v in problem.residuals_of(sc).items()}
        val, val0 = problem.calib_of(sc)
        P, k = h.make_problem(problem.frames_of(sc), val, val0, pts, empty)
        h.ba_upload(P)
        lo = h.linearize_all(False)
        assert lo["n_in"] == 0 and lo["energy"] == 0.0 and lo["new_frame_energy_th"] == 12 * 12 * 8
        acc = h.accumulate()
        assert acc["resInA"] == 0 and not acc["HA"].any() and not acc["Hsc"].any()
        h.close()


def test_bad_arguments_are_rejected(gpu):
    """The C ABI turns bad caller input into SOSBA_E_ARG instead of reading out of bounds: a negative point index, a target
    outside the window, residuals that are not point-major, and null mandatory arrays."""
    import ctypes as C
    from sos_slam_b200 import binding, problem
    sc = scene(**SMALL)
    h = open_handle(gpu, sc)
    upload(h, sc)
    res = problem.residuals_of(sc)
    for field, idx, val in (("point", 0, -1), ("point", 5, 10 ** 6), ("target", 3, -1), ("target", 3, sc.nf)):
        bad = {k: v.copy() for k, v in res.items()}
        bad[field][idx] = val
        with pytest.raises(binding.SosbaError) as e:
            h.residuals_set(bad)
        assert "rc=-1" in str(e.value)
    bad = {k: v.copy() for k, v in res.items()}
    bad["point"][[0, -1]] = bad["point"][[-1, 0]]
    with pytest.raises(binding.SosbaError):
        h.residuals_set(bad)
    S, keep = h._residuals_struct(res)
    S.point = None
    assert gpu.f("residuals_set")(h.h, C.byref(S)) == -1
    S, keep = h._points_struct(problem.points_of(sc))
    S.color = None
    assert gpu.f("points_set")(h.h, C.byref(S)) == -1
    # the handle is still usable
    h.points_set(problem.points_of(sc))
    h.residuals_set(res)
    assert h.linearize_all(False)["n_in"] > 0
    h.close()


# ---- next row: immature points ------------------------------------------------
@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B, KITTI], ids=["small", "configB", "kitti"])
def test_immature_trace_bit_exact(gpu, orc, cfg):
    """ImmaturePoint::ImmaturePoint + traceOn (ImmaturePoint.cpp:28-60, 70-415) through three consecutive
    traceNewCoarse passes (FullSystem.cpp:311-361): unbounded interval, then two refinements.  Every field of every
    point (colours, weights, gradH, energyTH, interval, quality, status, uv, pixel interval) and the status counters are
    identical bit for bit to the oracle's."""
    from sos_slam_b200 import synth
    sc = scene(**cfg)
    hg, ho = open_handle(gpu, sc), open_handle(orc, sc)
    first = sc.nf - 3
    base = synth.trace_case(sc, first, n_per_host=500, seed=21)
    keep = base["host"] < first
    sel = dict(host=base["host"][keep], u=base["u"][keep], v=base["v"][keep])
    pg, po = trace_points(hg, sc, sel), trace_points(ho, sc, sel)
    for k in pg:
        assert np.array_equal(pg[k], po[k], equal_nan=True), k
    seen = np.zeros(6, np.int64)
    for new_frame in (first, first + 1, first + 2):
        case = synth.trace_case(sc, new_frame, n_per_host=1, seed=21, pose_noise=2e-3 if new_frame == first + 2 else 0.0)
        cg = hg.trace_immature(new_frame, sel["host"], case["KRKi"], case["Kt"], case["aff"], pg)
        co = ho.trace_immature(new_frame, sel["host"], case["KRKi"], case["Kt"], case["aff"], po)
        assert np.array_equal(cg, co), (new_frame, cg, co)
        assert cg.sum() == sel["host"].size and np.array_equal(cg, np.bincount(pg["status"], minlength=6))
        for k in pg:
            assert np.array_equal(pg[k], po[k], equal_nan=True), (new_frame, k, int(np.sum(pg[k] != po[k])))
        seen += cg
    assert np.all(seen[:4] > 0) and (cfg is SMALL or seen[4] > 0), seen     # GOOD, OOB, OUTLIER, SKIPPED (+ BADCONDITION on the larger images)

    # activation: optimizeImmaturePoint + linearizeResidual (FullSystemOptPoint.cpp:47-192, ImmaturePoint.cpp:475-545) of
    # every point that was traced successfully at least once -- result code, depth and residual states bit for bit
    ok = np.isfinite(pg["idepth_max"])
    sub = {k: x[ok] for k, x in pg.items()}
    win = synth.activation_case(sc)
    outs = [h.optimize_immature(np.arange(sc.nf), win["RTll"], win["tTll"], win["aff"], win["calib"], sel["host"][ok], sub) for h in (hg, ho)]
    for name, x, y in zip(("result", "idepth", "res_state"), outs[0], outs[1]):
        assert np.array_equal(x, y, equal_nan=True), (name, int(np.sum(x != y)))
    res = outs[0][0]
    assert (res == 1).sum() > ok.sum() // 4 and (res != 1).sum() > 0, np.bincount(res + 1, minlength=3)
    assert np.all(outs[0][2][np.arange(ok.sum()), sel["host"][ok]] == 255)
    hg.close(); ho.close()


def test_immature_edge_cases(gpu, orc):
    """n = 0, a candidate outside the selector margin, a host index out of range, a point that is already OOB."""
    from sos_slam_b200 import binding, synth
    sc = scene(**SMALL)
    hg = open_handle(gpu, sc)
    empty = hg.immature_init(0, np.zeros(0, np.int32), np.zeros(0, np.int32))
    case = synth.trace_case(sc, sc.nf - 1, n_per_host=4)
    assert hg.trace_immature(sc.nf - 1, np.zeros(0, np.int32), case["KRKi"], case["Kt"], case["aff"], empty).sum() == 0
    with pytest.raises(binding.SosbaError):
        hg.immature_init(0, np.array([0], np.int32), np.array([10], np.int32))
    pts = hg.immature_init(0, np.array([50, 60], np.int32), np.array([40, 44], np.int32))
    with pytest.raises(binding.SosbaError):
        hg.trace_immature(sc.nf - 1, np.array([0, sc.nf], np.int32), case["KRKi"], case["Kt"], case["aff"], pts)
    pts["status"][0] = binding.IPS_OOB
    before = {k: v.copy() for k, v in pts.items()}
    c = hg.trace_immature(sc.nf - 1, np.array([0, 0], np.int32), case["KRKi"], case["Kt"], case["aff"], pts)
    assert c[binding.IPS_OOB] >= 1 and c.sum() == 2
    for k in pts:   # an OOB point is returned untouched (ImmaturePoint.cpp:74-75)
        assert np.array_equal(pts[k][0], before[k][0], equal_nan=True), k
    hg.close()


# ---- next row: pre-pyramid image path -----------------------------------------
@pytest.mark.parametrize("mode", ["full8", "full16", "nocalib", "passthrough"])
@pytest.mark.parametrize("shape", [(640, 480), (1232, 368)])
def test_undistort_raw_bit_exact(gpu, orc, mode, shape):
    """Raw frame -> PhotometricUndistorter::processFrame + Undistort::undistort (util/Undistort.cpp:194-227, 361-458) ->
    makeImages: the undistorted irradiance and every pyramid level identical bit for bit to the oracle's."""
    from sos_slam_b200 import binding, synth
    w, h = shape
    bits = 16 if mode == "full16" else 8
    c = synth.undistort_case(w, h, bits=bits) if mode != "passthrough" else synth.undistort_case(w, h, w_org=w, h_org=h)
    G = None if mode == "nocalib" else c["G"]
    V = None if mode == "nocalib" else c["vignette_inv"]
    rx, ry = (None, None) if mode == "passthrough" else (c["remapX"], c["remapY"])
    B = (np.linspace(0, 255, 256) ** 1.01).astype(np.float32)
    outs = []
    for lib in (gpu, orc):
        cfg = lib.config_default(w, h)
        cfg.max_frames = 1
        hd = binding.Handle(lib, cfg)
        hd.undistort_set(c["w_org"], c["h_org"], rx, ry, G, V)
        img = hd.frame_make_images_raw(0, c["raw"], factor=0.9, B=B, want_image=True)
        outs.append((img, [hd.frame_get_level(0, l) for l in range(hd.levels)]))
        hd.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


# ---- next row: loop-closure direct alignment ----------------------------------
@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B], ids=["small", "configB"])
def test_loop_pose(gpu, orc, cfg):
    """PoseEstimator::calcRes / calcGSSSE (LoopClosure/PoseEstimator.cpp:147-284, 75-145) on every level: counts
    (numTermsInE / numTermsInWarped / numSaturated) exact, energy and flow indicators 2e-5, H and b 1e-4."""
    from sos_slam_b200 import binding, synth
    sc = scene(**cfg)
    hg, ho = open_handle(gpu, sc), open_handle(orc, sc)
    K = sc.K.astype(np.float32)
    lv = [np.asarray(ho.frame_get_level(0, l)[0], np.float32).reshape(sc.h >> l, sc.w >> l, 3)[..., 0] for l in range(ho.levels)]
    case = synth.loop_case(sc, 0, ho.levels, lv, n=6000)
    T = (np.linalg.inv(sc.camToWorld_true[sc.nf - 1]) @ sc.camToWorld_true[0] @ synth.se3_exp(np.array([0.001, -0.0008, 0.0005, 0.0005, -0.0003, 0.0004])))[:3, :4]
    with pytest.raises(binding.SosbaError):
        hg.tracker_make_k(K); hg.loop_calc_res(0, sc.nf - 1, T, (1.0, 0.0), 20.0)       # no points yet
    for h in (hg, ho):
        h.tracker_make_k(K)
        h.loop_set_points(case["xyz"], case["color"])
    for lvl in range(hg.levels):
        for cutoff in (20.0, 6.0):
            og, cg = hg.loop_calc_res(lvl, sc.nf - 1, T, (1.02, -1.5), cutoff)
            oo, co = ho.loop_calc_res(lvl, sc.nf - 1, T, (1.02, -1.5), cutoff)
            assert np.array_equal(cg, co), (lvl, cg, co)
            assert cg[1] > 500
            assert np.allclose(og, oo, rtol=2e-5, atol=1e-6), (lvl, og, oo)
            Hg, bg = hg.loop_calc_gs(lvl, 1.02, 0.0)
            Ho, bo = ho.loop_calc_gs(lvl, 1.02, 0.0)
            assert relerr(Hg, Ho) < 1e-4 and relerr(bg, bo) < 1e-4
    with pytest.raises(binding.SosbaError):
        hg.tracker_calc_gs_pose(0, 1.0, 0.0)                 # the warped buffers belong to the loop variant
    hg.close(); ho.close()


# ---- next row: pixel selection --------------------------------------------------
@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B, KITTI], ids=["small", "configB", "kitti"])
def test_pixel_select_bit_exact(gpu, orc, cfg):
    """PixelSelector::makeMaps / makeHists / select (PixelSelector2.cpp:69-422): status map, raster-order list, returned
    count and currentPotential identical to the oracle's -- at fixed potentials without sub-sampling, and through the
    recursion + random sub-sampling of makeMaps for several densities, with the selector state carried from call to call."""
    sc = scene(**cfg)
    hg, ho = open_handle(gpu, sc), open_handle(orc, sc)
    rp = np.random.default_rng(3141592).integers(0, 256, sc.w * sc.h).astype(np.uint8)
    for pot in (1, 2, 3, 7):
        outs = []
        for h in (hg, ho):
            h.pixel_selector_set(rp, pot)
            outs.append(h.pixel_select(1, 1e9, recursions_left=0, cap=sc.w * sc.h))
        g, o = outs
        assert g["n"] == o["n"] and g["potential"] == o["potential"], (pot, g["n"], o["n"])
        assert np.array_equal(g["map"], o["map"]), (pot, int(np.sum(g["map"] != o["map"])))
        for k in ("u", "v", "type"):
            assert np.array_equal(g[k], o[k]), (pot, k)
        assert g["n"] > 50
    for h in (hg, ho):
        h.pixel_selector_set(rp, 3)
    for slot, density in ((0, 1500.0), (1, 1500.0), (2, 6000.0), (3, 200.0), (0, 30000.0), (1, 1500.0)):
        g, o = hg.pixel_select(slot, density, cap=sc.w * sc.h), ho.pixel_select(slot, density, cap=sc.w * sc.h)
        assert (g["n"], g["potential"]) == (o["n"], o["potential"]), (slot, density, g["n"], o["n"], g["potential"], o["potential"])
        assert np.array_equal(g["map"], o["map"])
        for k in ("u", "v", "type"):
            assert np.array_equal(g[k], o[k]), (slot, density, k)
        assert g["n"] == np.count_nonzero(g["map"])
    hg.close(); ho.close()


@pytest.mark.parametrize("kind", ["stripes", "mixed"])
def test_pixel_select_degenerate_directions(gpu, orc, kind):
    """Gradients exactly orthogonal to direction 0 (an image area that varies along x only): whether a block selects a pixel
    then depends on its direction, the per-block counts of the first pass are not final, and the running count of the
    reference even stalls on such blocks (no selection -> same randomPattern entry -> same direction).  "mixed": a
    degenerate band inside a textured image (the fix-up rounds settle); "stripes": the whole image (sequential fallback)."""
    from sos_slam_b200 import binding
    w, h = 320, 240
    xs = np.arange(w, dtype=np.float32)
    img = np.tile(100 + 60 * ((xs // 6) % 2) + 25 * ((xs // 45) % 2), (h, 1)).astype(np.float32)   # vertical step edges: dy == 0 exactly
    if kind == "mixed":
        tex = scene(**SMALL).images[0]
        img[:, 96:] = tex[:, 96:]
    rp = np.random.default_rng(5).integers(0, 256, w * h).astype(np.uint8)
    outs = []
    for lib in (gpu, orc):
        cfg = lib.config_default(w, h)
        cfg.max_frames = 1
        hd = binding.Handle(lib, cfg)
        hd.frame_make_images(0, img)
        hd.pixel_selector_set(rp, 2)
        outs.append([hd.pixel_select(0, 1e9, recursions_left=0, cap=w * h), hd.pixel_select(0, 800.0)])
        hd.close()
    for g, o in zip(*outs):
        assert g["n"] == o["n"] and g["potential"] == o["potential"], (g["n"], o["n"])
        assert np.array_equal(g["map"], o["map"]), int(np.sum(g["map"] != o["map"]))
    assert outs[0][0]["n"] > (500 if kind == "mixed" else 0)


# ---- next row (second half): CoarseInitializer::calcResAndGS ---------------------
@pytest.mark.parametrize("big_t", [False, True], ids=["alphaW", "alphaOpt0"])
@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B], ids=["small", "configB"])
def test_init_calc_res_and_gs(gpu, orc, cfg, big_t):
    """CoarseInitializer::calcResAndGS (CoarseInitializer.cpp:450-673) on every level, both alpha regimes: the per-point
    outputs (isGood_new, energy_new, maxstep, lastHessian_new, JbBuffer_new) bit for bit, H / b / Hsc / bsc 1e-4, E 2e-5."""
    from sos_slam_b200 import binding, synth
    sc = scene(**cfg)
    xi = np.array([0.002, -0.001, 0.0015, 0.0004, -0.0006, 0.0003]) * (40.0 if big_t else 1.0)
    T = synth.se3_exp(xi)[:3, :4]
    tlog = xi[:3].astype(np.float32)
    hs = []
    for lib in (gpu, orc):
        h = open_handle(lib, sc)
        h.tracker_make_k(sc.K.astype(np.float32))
        hs.append(h)
    hg, ho = hs
    for lvl in range(hg.levels):
        pts = synth.init_case(sc, lvl, n=max(100, 9000 >> (2 * lvl)))
        pts["JbBuffer_new"] = np.full((len(pts["u"]), 10), 0.25, np.float32)      # rows of points that are bad on entry must survive
        pts["lastHessian_new"] = np.full(len(pts["u"]), 3.0, np.float32)
        g = hg.init_calc_res_and_gs(lvl, 0, 1, T, (0.02, -1.0), tlog, pts)
        o = ho.init_calc_res_and_gs(lvl, 0, 1, T, (0.02, -1.0), tlog, pts)
        for k in ("isGood_new", "energy_new", "maxstep", "lastHessian_new", "JbBuffer_new"):
            assert np.array_equal(g[k], o[k]), (lvl, k, int(np.sum(g[k] != o[k])))
        for k in ("H", "b", "Hsc", "bsc"):
            assert relerr(g[k], o[k]) < 1e-4, (lvl, k, relerr(g[k], o[k]))
        assert np.allclose(g["res3"], o["res3"], rtol=2e-5)
        assert np.all(g["JbBuffer_new"][pts["isGood"] == 0] == 0.25) and 0.3 < g["isGood_new"].mean() < 0.95
    with pytest.raises(binding.SosbaError):
        bad = dict(pts); bad["u"] = pts["u"].copy(); bad["u"][0] = 0.0
        hg.init_calc_res_and_gs(0, 0, 1, T, (0.0, 0.0), tlog, bad)
    hg.close(); ho.close()


def test_immature_pool_resident(gpu, orc):
    """The resident pool (sosba_immature_pool_set / _trace / _get) gives exactly what the per-call path gives: three traces
    on the device without the points leaving HBM, then one read-back, against the oracle's per-call results."""
    from sos_slam_b200 import synth
    sc = scene(**CONFIG_B)
    hg, ho = open_handle(gpu, sc), open_handle(orc, sc)
    first = sc.nf - 3
    base = synth.trace_case(sc, first, n_per_host=700, seed=33)
    keep = base["host"] < first
    sel = dict(host=base["host"][keep], u=base["u"][keep], v=base["v"][keep])
    pg, po = trace_points(hg, sc, sel), trace_points(ho, sc, sel)
    hg.immature_pool_set(sel["host"], pg)
    for new_frame in (first, first + 1, first + 2):
        case = synth.trace_case(sc, new_frame, n_per_host=1, seed=33)
        cg = hg.immature_pool_trace(new_frame, case["KRKi"], case["Kt"], case["aff"])
        co = ho.trace_immature(new_frame, sel["host"], case["KRKi"], case["Kt"], case["aff"], po)
        assert np.array_equal(cg, co), (new_frame, cg, co)
    untouched = {k: v.copy() for k, v in pg.items()}
    hg.immature_pool_get(sel["host"], pg)
    for k in pg:
        assert np.array_equal(pg[k], po[k], equal_nan=True), k
    assert not np.array_equal(untouched["status"], pg["status"])          # the host copy only changes at pool_get
    hg.immature_pool_set(sel["host"][:0], {k: v[:0] for k, v in pg.items()})
    assert hg.immature_pool_trace(first, case["KRKi"], case["Kt"], case["aff"]).sum() == 0
    hg.close(); ho.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_pixel_select_random_images(gpu, orc, seed):
    """Noisy random 112x80 images (not a multiple of the 32-pixel threshold blocks, pyramid forced to 3 levels): maps and
    counts identical to the oracle for potentials 1..4 and through makeMaps."""
    from sos_slam_b200 import binding
    w, h = 112, 80
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:h, 0:w]
    img = (128 + 60 * np.sin(xs * rng.uniform(0.1, 0.4) + ys * rng.uniform(-0.3, 0.3)) + rng.normal(0, 12, (h, w))).astype(np.float32)
    rp = rng.integers(0, 256, w * h).astype(np.uint8)
    hs = []
    for lib in (gpu, orc):
        cfg = lib.config_default(w, h)
        cfg.max_frames = 1
        cfg.pyr_levels = 3
        hd = binding.Handle(lib, cfg)
        hd.frame_make_images(0, img)
        hs.append(hd)
    for pot in (1, 2, 3, 4):
        outs = []
        for hd in hs:
            hd.pixel_selector_set(rp, pot)
            outs.append(hd.pixel_select(0, 1e9, recursions_left=0, cap=w * h))
        assert outs[0]["n"] == outs[1]["n"] and np.array_equal(outs[0]["map"], outs[1]["map"]), (pot, outs[0]["n"], outs[1]["n"])
    for hd in hs:
        hd.pixel_selector_set(rp, 3)
    for density in (60.0, 400.0, 2500.0):
        g, o = hs[0].pixel_select(0, density, cap=w * h), hs[1].pixel_select(0, density, cap=w * h)
        assert (g["n"], g["potential"]) == (o["n"], o["potential"]) and np.array_equal(g["map"], o["map"]), density
    for hd in hs:
        hd.close()


# ---- a16 / a17: the control loops of the direct alignment, resident on the device ---------------------------------------
def _lm_case(orc, cfg):
    """reference lists input (from the oracle's optimised window, so that both sides get identical arrays)"""
    from _track_case import coarse_depth_input
    sc = scene(**cfg)
    ho = open_handle(orc, sc)
    cpt, hdi = coarse_depth_input(ho, sc)
    # three and four points on one pixel: the order-sensitive float sums of the splat
    cpt = np.concatenate([cpt, cpt[:3] * 0 + cpt[5], cpt[7:9] * 0 + cpt[5], cpt[20:23] * 0 + cpt[30]])
    hdi = np.concatenate([hdi, hdi[:3] * 1.7, hdi[7:9] * 0.3, hdi[20:23] * 0.9])
    return sc, ho, cpt, hdi


@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B, KITTI], ids=["small", "configB", "kitti"])
def test_make_coarse_depth_bit_exact(gpu, orc, cfg):
    """CoarseTracker::makeCoarseDepthL0 (CoarseTracker.cpp:56-230): pc_u / pc_v / pc_idepth / pc_color of every level identical
    to the oracle's (count, order, every bit), scaleCoarseDepthL0 included."""
    sc, ho, cpt, hdi = _lm_case(orc, cfg)
    hg = open_handle(gpu, sc)
    ng = hg.tracker_make_coarse_depth(sc.nf - 1, cpt, hdi)
    no = ho.tracker_make_coarse_depth(sc.nf - 1, cpt, hdi)
    assert np.array_equal(ng, no) and ng[0] > len(hdi) // 2
    for scale in (None, 1.7):
        if scale:
            hg.tracker_scale_coarse_depth(scale); ho.tracker_scale_coarse_depth(scale)
        for l in range(hg.levels):
            for a, b in zip(hg.tracker_get_ref(l), ho.tracker_get_ref(l)):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (l, scale)
    # an empty reference (no IN residual on the newest keyframe): empty lists, and tracking it is not an error
    assert not hg.tracker_make_coarse_depth(sc.nf - 1, np.zeros((0, 3), np.float32), np.zeros(0, np.float32)).any()
    with pytest.raises(Exception):
        hg.tracker_make_coarse_depth(sc.nf - 1, np.array([[sc.w + 5.0, 3.0, 1.0]], np.float32), np.ones(1, np.float32))
    hg.close(); ho.close()


@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B], ids=["small", "configB"])
def test_track_newest_coarse(gpu, orc, cfg):
    """CoarseTracker::trackNewestCoarse (CoarseTracker.cpp:366-552), a batch of hypotheses in one launch: level schedule,
    iteration counts, accept / reject sequence, cutoff doubling, abort and return value identical to the oracle; residuals
    2e-5, flow indicators 1e-4, final pose and affine 1e-5 relative to the step the loop made."""
    from _track_case import hypotheses, quat_to_T, ref_affine
    sc, ho, cpt, hdi = _lm_case(orc, cfg)
    hg = open_handle(gpu, sc)
    Ttrue, hyps = hypotheses(sc, n_extra=6, seed=4)
    hyps.append(dict(hyps[0], min_res_for_abort=[0.01] * 5))
    hyps.append(dict(hyps[1], aff_g2l=(0.3, -20.0)))
    ref_aff, ref_exp, new_exp = ref_affine(sc)
    outs = []
    for h in (hg, ho):
        h.tracker_make_k(sc.K.astype(np.float32))
        h.tracker_make_coarse_depth(sc.nf - 1, cpt, hdi)
        outs.append(h.tracker_track(sc.nf - 2, ref_exp, new_exp, ref_aff, h.levels - 1, hyps))
    n_ok = n_exact = 0
    for hy, g, o in zip(hyps, *outs):
        Tg, To, T0 = quat_to_T(g["q"], g["t"]), quat_to_T(o["q"], o["t"]), quat_to_T(hy["q"], hy["t"])
        moved = max(np.abs(To - T0).max(), 1e-3)
        n_ok += g["ok"]
        same = g["pass_lvl"] == o["pass_lvl"] and g["pass_iterations"] == o["pass_iterations"] and g["pass_accept"] == o["pass_accept"]
        if not same:
            # The only licence to differ: an accept / reject decision taken on two mean energies closer than 2e-5 relative (flagged
            # by either side; below the noise of the reference's own float sums, whose normal equations no parallel sum reproduces
            # bit for bit).  Everything before that iteration is identical and the loop still ends at the same minimum.
            p = next(i for i in range(min(g["n_passes"], o["n_passes"])) if (g["pass_lvl"][i], g["pass_iterations"][i], g["pass_accept"][i]) !=
                     (o["pass_lvl"][i], o["pass_iterations"][i], o["pass_accept"][i]))
            assert g["pass_lvl"][:p + 1] == o["pass_lvl"][:p + 1] and g["pass_cutoff_repeat"][:p + 1] == o["pass_cutoff_repeat"][:p + 1]
            diff = g["pass_accept"][p] ^ o["pass_accept"][p]
            k = (diff & -diff).bit_length() - 1 if diff else min(g["pass_iterations"][p], o["pass_iterations"][p]) - 1
            assert ((g["pass_tie"][p] | o["pass_tie"][p]) >> k) & 1, (p, k, g, o)
            assert g["ok"] == o["ok"]
            np.testing.assert_allclose(g["last_residuals"], o["last_residuals"], rtol=2e-3, equal_nan=True)
            assert np.abs(Tg - To).max() <= 2e-2 * moved
            continue
        n_exact += 1
        assert g["ok"] == o["ok"] and g["n_passes"] == o["n_passes"] and g["pass_cutoff_repeat"] == o["pass_cutoff_repeat"]
        np.testing.assert_allclose(g["pass_residual"], o["pass_residual"], rtol=2e-5)
        np.testing.assert_allclose(g["last_residuals"], o["last_residuals"], rtol=2e-5, equal_nan=True)
        np.testing.assert_allclose(g["flow_indicators"], o["flow_indicators"], rtol=1e-4, atol=1e-7)
        assert np.abs(Tg - To).max() <= 1e-5 * moved + 1e-9, (np.abs(Tg - To).max(), moved)
        np.testing.assert_allclose(g["aff_g2l"], o["aff_g2l"], rtol=1e-4, atol=1e-4)
    assert n_exact >= len(hyps) - 2      # near-ties are the exception
    assert n_ok >= len(hyps) - 3 and not outs[0][-2]["ok"]
    dT = np.linalg.inv(quat_to_T(outs[0][0]["q"], outs[0][0]["t"])) @ Ttrue
    assert np.abs(dT - np.eye(4)).max() < 3e-3
    hg.close(); ho.close()


def test_optimize_scale(gpu, orc):
    """ScaleOptimizer::optimizeScale (ScaleOptimizer.cpp:120-230) from the 7 start values of FullSystem::optimizeScale
    (FullSystem.cpp:1135) in one launch: schedule and accept sequence identical, scale 1e-5, error 2e-5."""
    from _track_case import stereo_frame
    sc, ho, cpt, hdi = _lm_case(orc, SMALL)
    hg = open_handle(gpu, sc)
    K = sc.K.astype(np.float32)
    T10, img1 = stereo_frame(sc, SMALL["seed"])
    starts = [0.1, 0.2, 0.5, 1, 2, 5, 10]
    outs = []
    for h in (hg, ho):
        h.tracker_make_k(K)
        h.tracker_make_coarse_depth(sc.nf - 1, cpt, hdi)
        h.scale_set_stereo(T10, K)
        h.frame_make_images(sc.nf, img1)
        outs.append(h.scale_optimize(sc.nf, h.levels - 1, starts))
    n_exact = 0
    for s0, g, o in zip(starts, *outs):
        same = g["pass_lvl"] == o["pass_lvl"] and g["pass_iterations"] == o["pass_iterations"] and g["pass_accept"] == o["pass_accept"]
        if not same:    # only after a decision between two mean energies closer than 2e-5 relative (see test_track_newest_coarse)
            p = next(i for i in range(min(g["n_passes"], o["n_passes"])) if (g["pass_lvl"][i], g["pass_iterations"][i], g["pass_accept"][i]) !=
                     (o["pass_lvl"][i], o["pass_iterations"][i], o["pass_accept"][i]))
            diff = g["pass_accept"][p] ^ o["pass_accept"][p]
            k = (diff & -diff).bit_length() - 1 if diff else min(g["pass_iterations"][p], o["pass_iterations"][p]) - 1
            assert ((g["pass_tie"][p] | o["pass_tie"][p]) >> k) & 1, (s0, p, k, g, o)
            assert g["error"] == pytest.approx(o["error"], rel=2e-3)
            continue
        n_exact += 1
        # the 1-DoF loop runs in float (inc = -b / H): one ulp in H or b moves the scale by ~1e-7 per iteration
        assert g["scale"] == pytest.approx(o["scale"], rel=2e-5) and g["error"] == pytest.approx(o["error"], rel=5e-5), (s0, g, o)
    assert n_exact >= len(starts) - 2
    best = min((g for g in outs[0] if g["error"] > 0), key=lambda g: g["error"])
    assert abs(best["scale"] - 1.0) < 0.03
    hg.close(); ho.close()


@pytest.mark.parametrize("cfg", [SMALL, CONFIG_B, KITTI], ids=["small", "configB", "kitti"])
def test_distance_map_bit_exact(gpu, orc, cfg):
    """CoarseDistanceMap::makeDistanceMap + growDistBFS (CoarseTracker.cpp:789-916): fwdWarpedIDDistFinal identical to the oracle's
    breadth-first flood (second half of SURVEY.md 8f rank 3), full window and sparse inputs, empty input."""
    from _track_case import distance_map_case
    sc = scene(**cfg)
    hg, ho = open_handle(gpu, sc), open_handle(orc, sc)
    KRKi, Kt, host, u, v, idp = distance_map_case(sc)
    for k in (len(host), 5, 1, 0):
        dg = hg.distance_map(KRKi, Kt, host[:k], u[:k], v[:k], idp[:k])
        do = ho.distance_map(KRKi, Kt, host[:k], u[:k], v[:k], idp[:k])
        assert np.array_equal(dg, do), (k, int((dg != do).sum()))
    assert (do == 1000).all() and (dg == 1000).all()
    hg.close(); ho.close()


def _host_solve(sys, lam=1e-5):
    """EnergyFunctional::solveSystemF (EnergyFunctional.cpp:1094-1148) without IMU and without marginalisation prior, on the host."""
    H = sys["H_top"].copy()
    b = sys["b_top"] - sys["b_sc"]
    H[np.diag_indices_from(H)] *= (1 + lam)
    H -= sys["H_sc"] * float(np.float32(1.0) / np.float32(1 + lam))
    S = 1.0 / np.sqrt(np.diag(H) + 10.0)
    return S * np.linalg.solve(S[:, None] * H * S[None, :], S * b)


@pytest.mark.parametrize("cfg", [SMALL, EUROC], ids=["small", "euroc-752x480"])
def test_split_body_with_caller_side_solve(gpu, orc, cfg):
    """sosba_ba_system / sosba_ba_step: the loop body of FullSystem::optimize with the solve on the caller's side (what an IMU
    configuration needs, EnergyFunctional.cpp:1052-1171).  (1) the stitched systems match the oracle's (1e-4); (2) fed with the same
    x, the step + linearizeAll + applyRes leave identical residual states and energies (1e-6), step sums 1e-5; (3) three such
    bodies with a host LDL^T land where the device-resident loop (sosba_ba_iterate) lands."""
    sc = scene(**cfg)
    hs = []
    for lib in (gpu, orc, gpu):
        h = open_handle(lib, sc)
        upload(h, sc)
        h.reset_oob(); h.linearize_all(False); h.apply_res()
        hs.append(h)
    hg, ho, hres = hs
    for it in range(3):
        sg, so = hg.ba_system(), ho.ba_system()
        assert sg["resInA"] == so["resInA"] and sg["resInL"] == so["resInL"]
        for k in ("H_top", "H_sc"):
            assert relerr(sg[k], so[k]) < 1e-4, (it, k, relerr(sg[k], so[k]))
        for k in ("b_top", "b_sc"):
            assert relerr(sg[k], so[k]) < 2e-4, (it, k)
        assert np.array_equal(sg["H_top"], sg["H_top"].T) and np.array_equal(sg["H_sc"], sg["H_sc"].T)
        x = _host_solve(so)
        og, oo = hg.ba_step(x), ho.ba_step(x)
        assert (og["n_in"], og["n_oob"], og["n_outlier"]) == (oo["n_in"], oo["n_oob"], oo["n_outlier"])
        assert og["energy"] == pytest.approx(oo["energy"], rel=1e-6) and og["new_frame_energy_th"] == pytest.approx(oo["new_frame_energy_th"], rel=1e-6)
        for k in ("sum_a", "sum_b", "sum_t", "sum_r", "sum_id", "sum_nid", "num_id"):
            assert og[k] == pytest.approx(oo[k], rel=2e-5, abs=1e-30), k
        assert np.array_equal(hg.get_state()["state"], ho.get_state()["state"])
    # the device-resident loop takes the same three steps (its LDL^T runs on the device)
    hres.ba_iterate(3)
    stg, str_ = hg.get_state(), hres.get_state()
    assert int((stg["state"] != str_["state"]).sum()) <= max(2, stg["state"].size // 500)
    assert np.abs(stg["energy"] - str_["energy"]).max() <= 2e-3 * np.abs(str_["energy"]).max()
    for h in hs:
        h.close()
