"""The front-end rows (raw frame, pixel selection, immature points, loop-closure alignment) driven on a small
seeded window.  Used by tools/make_golden_frontend.py (oracle -> fixture) and tests/test_golden.py (oracle and CUDA path
against the fixture).  Every input is stored in the fixture."""
import numpy as np

from sosba_loader import load_package

load_package()
from sos_slam_b200 import binding  # noqa: E402


def make_inputs():
    """Seeded inputs, all stored in the fixture so the tests do not depend on the generators.  256x192 is the smallest size
    with the three pyramid levels the selector reads; the images are stored as 8-bit integers (exactly representable)."""
    from sos_slam_b200 import synth
    sc = synth.make_scene(w=256, h=192, nf=4, n_points=150, seed=11)
    rng = np.random.default_rng(2024)
    uc = synth.undistort_case(sc.w, sc.h, w_org=sc.w + 16, h_org=sc.h + 12, seed=4)
    return {"in_images": np.stack([np.clip(np.rint(i), 0, 255).astype(np.uint8) for i in sc.images]), "in_K": sc.K, "in_evalPT": sc.evalPT,
            "in_pt_host": sc.pt_host, "in_pt_u": sc.pt_u, "in_pt_v": sc.pt_v, "in_pt_idepth": sc.pt_idepth, "in_pt_color": sc.pt_color,
            "fin_random_pattern": rng.integers(0, 256, sc.w * sc.h).astype(np.uint8), "fin_raw": uc["raw"], "fin_G": uc["G"],
            **{"fin_init_" + k: v for k, v in synth.init_case(sc, 0, n=400, seed=6).items()}}


def run(lib, F):
    G = F
    imgs = F["in_images"].astype(np.float32)
    nf, hgt, wid = imgs.shape
    cfg = lib.config_default(wid, hgt)
    cfg.max_frames = nf + 2
    h = binding.Handle(lib, cfg)
    for i in range(nf):
        h.frame_make_images(i, imgs[i])
    K = np.asarray(G["in_K"], np.float64)
    Kf = K.astype(np.float32)
    out = {}
    # 8f-2 raw frame -> irradiance -> pyramid (slot nf)
    # (the rectification map and the vignette are + - * / of integers in float32: regenerated, not stored)
    from sos_slam_b200 import synth
    import zlib
    uc = synth.undistort_case(wid, hgt, w_org=int(F["fin_raw"].shape[1]), h_org=int(F["fin_raw"].shape[0]), seed=4)
    h.undistort_set(int(F["fin_raw"].shape[1]), int(F["fin_raw"].shape[0]), uc["remapX"], uc["remapY"], F["fin_G"], uc["vignette_inv"])
    img = h.frame_make_images_raw(nf, F["fin_raw"], want_image=True)
    out["und_image_crc"] = np.uint32(zlib.crc32(np.ascontiguousarray(img).tobytes()))
    out["und_image_row"] = img[hgt // 2].copy()
    out["und_pyr2_dI"], out["und_pyr2_abs"] = h.frame_get_level(nf, 2)
    # 8f-3 pixel selection on frames 0 and 1 (selector state carried over)
    h.pixel_selector_set(F["fin_random_pattern"], 3)
    sel = [h.pixel_select(f, 150.0, cap=wid * hgt) for f in (0, 1)]
    out["sel_n"] = np.array([s["n"] for s in sel], np.int32)
    out["sel_potential"] = np.array([s["potential"] for s in sel], np.int32)
    out["sel_map0"] = sel[0]["map"].astype(np.uint8)
    out["sel_map1"] = sel[1]["map"].astype(np.uint8)
    # 8f-1 immature points on the selected pixels (inside the pattern margin), traced into frames 2 and 3, then activated
    parts, hosts = [], []
    for f, s in zip((0, 1), sel):
        ok = (s["u"] >= 6) & (s["u"] < wid - 7) & (s["v"] >= 6) & (s["v"] < hgt - 7)
        parts.append(h.immature_init(f, s["u"][ok], s["v"][ok]))
        hosts.append(np.full(int(ok.sum()), f, np.int32))
    pts = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    host = np.concatenate(hosts)
    out["imm_color"], out["imm_weights"], out["imm_gradH"] = pts["color"].copy(), pts["weights"].copy(), pts["gradH"].copy()
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]])
    counts = []
    for new in (2, 3):
        KRKi = np.zeros((nf, 3, 3), np.float32); Kt = np.zeros((nf, 3), np.float32); aff = np.zeros((nf, 2), np.float32)
        for f in range(nf):
            T = np.linalg.inv(G["in_evalPT"][new]) @ G["in_evalPT"][f]
            KRKi[f] = (Km.astype(np.float32) @ T[:3, :3].astype(np.float32)) @ np.linalg.inv(Km).astype(np.float32)
            Kt[f] = Km.astype(np.float32) @ T[:3, 3].astype(np.float32)
            aff[f] = (1.0, 0.0)
        counts.append(h.trace_immature(new, host, KRKi, Kt, aff, pts))
    out["trace_counts"] = np.stack(counts)
    for k in ("idepth_min", "idepth_max", "quality", "status", "uv", "pixel_interval"):
        out["trace_" + k] = pts[k].copy()
    ok = np.isfinite(pts["idepth_max"])
    RT = np.zeros((nf, nf, 3, 3), np.float32); tT = np.zeros((nf, nf, 3), np.float32); af = np.zeros((nf, nf, 2), np.float32)
    for a in range(nf):
        for b in range(nf):
            T = np.linalg.inv(G["in_evalPT"][b]) @ G["in_evalPT"][a]
            RT[a, b] = T[:3, :3]; tT[a, b] = T[:3, 3]; af[a, b] = (1.0, 0.0)
    res, idepth, states = h.optimize_immature(np.arange(nf), RT, tT, af, Kf, host[ok], {k: x[ok] for k, x in pts.items()})
    out["act_result"], out["act_idepth"], out["act_states"] = res, idepth, states
    # 8f-4 loop-closure alignment: the points of the BA window hosted in frame 0, against frame 3
    h.tracker_make_k(Kf)
    m = G["in_pt_host"] == 0
    u, v, d = G["in_pt_u"][m].astype(np.float64), G["in_pt_v"][m].astype(np.float64), G["in_pt_idepth"][m].astype(np.float64)
    xyz = np.stack([(u - K[2]) / K[0] / d, (v - K[3]) / K[1] / d, 1.0 / d], 1)
    color = np.repeat(G["in_pt_color"][m][:, 4:5], h.levels, 1).astype(np.float32)     # centre pattern sample on every level
    h.loop_set_points(xyz, color)
    T = (np.linalg.inv(G["in_evalPT"][3]) @ G["in_evalPT"][0])[:3, :4]
    o6, cnt = h.loop_calc_res(0, 3, T, (1.0, 0.0), 20.0)
    H, b = h.loop_calc_gs(0, 1.0, 0.0)
    out["loop_out6"], out["loop_counts"], out["loop_H"], out["loop_b"] = o6, cnt, H, b
    # 8f-4 CoarseInitializer::calcResAndGS on level 0, frame 0 -> frame 1, both alpha regimes
    ipts = {k[len("fin_init_"):]: F[k] for k in F.keys() if k.startswith("fin_init_")}
    for tag, scale in (("a", 1.0), ("b", 40.0)):
        xi = np.array([0.002, -0.001, 0.0015, 0.0004, -0.0006, 0.0003]) * scale
        from sos_slam_b200 import synth as _s
        Ti = _s.se3_exp(xi)[:3, :4]
        r = h.init_calc_res_and_gs(0, 0, 1, Ti, (0.02, -1.0), xi[:3].astype(np.float32), ipts)
        for k in ("isGood_new", "energy_new", "maxstep", "lastHessian_new", "JbBuffer_new"):
            out[f"init{tag}_{k}"] = r[k]
        out[f"init{tag}_H"], out[f"init{tag}_b"], out[f"init{tag}_Hsc"], out[f"init{tag}_bsc"], out[f"init{tag}_res3"] = r["H"], r["b"], r["Hsc"], r["bsc"], r["res3"]
    h.close()
    return out


EXACT = ("und_image_crc", "und_image_row", "und_pyr2_dI", "und_pyr2_abs", "sel_n", "sel_potential", "sel_map0", "sel_map1", "imm_color", "imm_weights", "imm_gradH",
         "trace_counts", "trace_idepth_min", "trace_idepth_max", "trace_quality", "trace_status", "trace_uv", "trace_pixel_interval", "act_result",
         "act_idepth", "act_states", "loop_counts") + tuple(f"init{t}_{k}" for t in "ab" for k in ("isGood_new", "energy_new", "maxstep",
                                                                                                   "lastHessian_new", "JbBuffer_new"))
CLOSE = (("loop_out6", 2e-5), ("loop_H", 1e-4), ("loop_b", 1e-4)) + tuple((f"init{t}_{k}", 1e-4) for t in "ab" for k in ("H", "b", "Hsc", "bsc", "res3"))
