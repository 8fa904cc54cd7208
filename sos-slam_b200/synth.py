"""Synthetic photometric-BA windows of the shapes BASELINE.json names.

A smooth textured surface z = Z(X, Y) is ray-cast from nf cameras on an arc; per-frame affine
brightness I_i = exp(a_i) * T + b_i.  Points are picked at gradient maxima of every frame but the
newest (the reference activates points only in older keyframes, FullSystem.cpp:375-531), each with a
residual to every other frame its 8-pixel pattern projects into.  States are perturbed from ground
truth so Gauss-Newton has work to do.  Pure numpy; workload generation only (no algorithm of the
path lives here).
"""
from __future__ import annotations

import dataclasses

import numpy as np

PATTERN = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], dtype=np.int32)
SCALE_A, SCALE_B = 10.0, 1000.0
SCALE_XI_TRANS, SCALE_XI_ROT = 0.5, 1.0
SCALE_F = SCALE_C = 50.0


def so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * (K @ K)


def se3_exp(xi):
    """xi = [upsilon, omega] (Sophus order) -> 4x4."""
    u, w = np.asarray(xi[:3], np.float64), np.asarray(xi[3:], np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    R = so3_exp(w)
    if th < 1e-10:
        V = R
    else:
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * K + (th - np.sin(th)) / th**3 * (K @ K)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ u
    return T


@dataclasses.dataclass
class Scene:
    w: int
    h: int
    nf: int
    K: np.ndarray                 # fx fy cx cy
    images: list                  # nf float32 (h, w)
    camToWorld_true: np.ndarray   # (nf,4,4)
    evalPT: np.ndarray            # (nf,4,4) estimated linearisation-point poses
    state: np.ndarray             # (nf,10) unscaled
    state_zero: np.ndarray        # (nf,10)
    aff_true: np.ndarray          # (nf,2) a,b
    ab_exposure: np.ndarray       # (nf,)
    frame_id: np.ndarray          # (nf,)
    # points
    pt_host: np.ndarray
    pt_u: np.ndarray
    pt_v: np.ndarray
    pt_idepth_true: np.ndarray
    pt_idepth: np.ndarray
    pt_color: np.ndarray          # (P,8)
    pt_weights: np.ndarray        # (P,8)
    # residuals (point-major)
    res_point: np.ndarray
    res_target: np.ndarray

    @property
    def n_points(self):
        return int(self.pt_u.shape[0])

    @property
    def n_residuals(self):
        return int(self.res_point.shape[0])


class _Surface:
    def __init__(self, rng, z0=2.0):
        self.z0 = z0
        nz = 6
        self.zf = rng.uniform(-1.5, 1.5, size=(nz, 2))
        self.zp = rng.uniform(0, 2 * np.pi, size=nz)
        self.za = rng.uniform(0.02, 0.07, size=nz)
        nt = 48
        # wavelengths between 12 and 160 px at (z0, f=500) -> world wavelengths lam_px * z0 / 500
        lam = np.exp(rng.uniform(np.log(12.0), np.log(160.0), size=nt)) * z0 / 500.0
        ang = rng.uniform(0, 2 * np.pi, size=nt)
        self.tf = (2 * np.pi / lam)[:, None] * np.stack([np.cos(ang), np.sin(ang)], 1)
        self.tp = rng.uniform(0, 2 * np.pi, size=nt)
        amp = rng.uniform(0.5, 1.0, size=nt) * np.sqrt(lam / lam.max()) ** 0.5
        self.ta = amp * (42.0 / np.sqrt(0.5 * np.sum(amp * amp)))  # texture std ~42 around 127.5

    def Z(self, X, Y):
        z = np.full_like(X, self.z0)
        for f, p, a in zip(self.zf, self.zp, self.za):
            z = z + a * np.sin(f[0] * X + f[1] * Y + p)
        return z

    def T(self, X, Y):
        t = np.full_like(X, 127.5)
        for f, p, a in zip(self.tf, self.tp, self.ta):
            t = t + a * np.sin(f[0] * X + f[1] * Y + p)
        return t

    def cast(self, c2w, dirs_c):
        """dirs_c: (...,3) camera rays with z=1.  Returns camera-frame depth (z_c) and world X,Y."""
        R, t = c2w[:3, :3], c2w[:3, 3]
        d = dirs_c @ R.T
        s = (self.z0 - t[2]) / d[..., 2]
        for _ in range(8):
            X = t[0] + s * d[..., 0]
            Y = t[1] + s * d[..., 1]
            s = (self.Z(X, Y) - t[2]) / d[..., 2]
        X = t[0] + s * d[..., 0]
        Y = t[1] + s * d[..., 1]
        return s, X, Y  # dirs have z_c = 1 so s is the camera-frame depth


def render_view(seed, w, h, K, camToWorld, aff=(0.0, 0.0)) -> np.ndarray:
    """The surface of make_scene(seed=...) seen from an arbitrary pose (e.g. camera 1 of a stereo pair, or a frame between
    keyframes): float32 irradiance image, I = exp(a) * texture + b like the window's frames."""
    fx, fy, cx, cy = [float(x) for x in K]
    surf = _Surface(np.random.default_rng(seed))
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    dirs = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs)], -1)
    z, X, Y = surf.cast(np.asarray(camToWorld, np.float64), dirs)
    img = np.clip(np.exp(aff[0]) * np.clip(surf.T(X, Y), 8.0, 247.0) + aff[1], 0.0, 255.0)
    return np.ascontiguousarray(img.astype(np.float32))


def make_scene(w=640, h=480, nf=8, n_points=2000, seed=1234, fx=None, fy=None, cx=None, cy=None,
               pose_noise=2e-3, idepth_noise=0.03, forward_motion=False, dtype=np.float32) -> Scene:
    rng = np.random.default_rng(seed)
    fx = 500.0 * w / 640.0 if fx is None else fx
    fy = fx if fy is None else fy
    cx = (w - 1) / 2.0 if cx is None else cx
    cy = (h - 1) / 2.0 if cy is None else cy
    surf = _Surface(rng)

    # camera arc: <= ~5 deg rotation, <= 0.3 units baseline over the window
    c2w = np.zeros((nf, 4, 4))
    for i in range(nf):
        s = i / max(nf - 1, 1) - 0.5
        if forward_motion:
            xi = np.array([0.05 * s, 0.01 * np.sin(3 * s), 0.5 * s, 0.004 * s, 0.01 * s, 0.002 * s])
        else:
            xi = np.array([0.30 * s, 0.04 * np.sin(3 * s), 0.05 * s * s, 0.01 * s, -0.06 * s, 0.02 * s])
        c2w[i] = se3_exp(xi)

    aff = np.stack([rng.normal(0, 0.05, nf), rng.normal(0, 5.0, nf)], 1)
    aff[0] = 0.0
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    dirs = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs)], -1)
    images, depths = [], []
    for i in range(nf):
        z, X, Y = surf.cast(c2w[i], dirs)
        img = np.clip(np.exp(aff[i, 0]) * np.clip(surf.T(X, Y), 8.0, 247.0) + aff[i, 1], 0.0, 255.0)
        images.append(np.ascontiguousarray(img.astype(np.float32)))
        depths.append(z)

    # --- points: per host frame, strongest-gradient pixel of each cell of a regular grid
    hosts = list(range(nf - 1)) if nf > 1 else [0]
    per_host = int(np.ceil(n_points / len(hosts)))
    margin = 12
    pt_host, pt_u, pt_v, pt_id, pt_col, pt_wgt = [], [], [], [], [], []
    c_out = 50.0 * 50.0
    for hi in hosts:
        img = images[hi].astype(np.float32)
        gx = np.zeros_like(img)
        gy = np.zeros_like(img)
        gx[:, 1:-1] = 0.5 * (img[:, 2:] - img[:, :-2])
        gy[1:-1, :] = 0.5 * (img[2:, :] - img[:-2, :])
        g2 = gx * gx + gy * gy
        aw, ah = w - 2 * margin, h - 2 * margin
        ncx = max(1, int(np.round(np.sqrt(per_host * aw / ah))))
        ncy = max(1, int(np.ceil(per_host / ncx)))
        cw, ch = aw / ncx, ah / ncy
        cnt = 0
        for j in range(ncy):
            for i in range(ncx):
                if cnt >= per_host or len(pt_u) >= n_points:
                    break
                x0, x1 = int(margin + i * cw), int(margin + (i + 1) * cw)
                y0, y1 = int(margin + j * ch), int(margin + (j + 1) * ch)
                if x1 <= x0 or y1 <= y0:
                    continue
                blk = g2[y0:y1, x0:x1]
                k = int(np.argmax(blk))
                v, u = y0 + k // blk.shape[1], x0 + k % blk.shape[1]
                pu = u + PATTERN[:, 0]
                pv = v + PATTERN[:, 1]
                pt_host.append(hi)
                pt_u.append(float(u))
                pt_v.append(float(v))
                pt_id.append(1.0 / depths[hi][v, u])
                pt_col.append(img[pv, pu])
                # ImmaturePoint.cpp:50-52: weights = sqrt(c / (c + |grad|^2))
                pt_wgt.append(np.sqrt(c_out / (c_out + g2[pv, pu])))
                cnt += 1
    P = len(pt_u)
    pt_host = np.array(pt_host, np.int32)
    pt_u = np.array(pt_u, np.float32)
    pt_v = np.array(pt_v, np.float32)
    pt_id_true = np.array(pt_id, np.float32)
    pt_col = np.array(pt_col, np.float32).reshape(P, 8)
    pt_wgt = np.array(pt_wgt, np.float32).reshape(P, 8)

    # --- residuals: every other frame the whole pattern projects into (true geometry, 6 px margin)
    res_point, res_target = [], []
    w2c = np.linalg.inv(c2w)
    Kinv_pts = np.stack([(pt_u - cx) / fx, (pt_v - cy) / fy, np.ones(P)], 1) / pt_id_true[:, None]  # camera-frame 3D
    vis = np.zeros((P, nf), bool)
    for t in range(nf):
        for hi in hosts:
            m = pt_host == hi
            if t == hi or not m.any():
                continue
            T = w2c[t] @ c2w[hi]
            Xc = Kinv_pts[m] @ T[:3, :3].T + T[:3, 3]
            uu = fx * Xc[:, 0] / Xc[:, 2] + cx
            vv = fy * Xc[:, 1] / Xc[:, 2] + cy
            ok = (Xc[:, 2] > 0.1) & (uu > 6) & (vv > 6) & (uu < w - 7) & (vv < h - 7)
            idx = np.nonzero(m)[0]
            vis[idx[ok], t] = True
    for p in range(P):
        for t in np.nonzero(vis[p])[0]:
            res_point.append(p)
            res_target.append(int(t))
    res_point = np.array(res_point, np.int32)
    res_target = np.array(res_target, np.int32)

    # --- estimates: perturbed poses / idepths; older frames carry a non-zero state - state_zero
    evalPT = np.zeros_like(c2w)
    state = np.zeros((nf, 10))
    for i in range(nf):
        noise = np.zeros(6) if i == 0 else rng.normal(0, pose_noise, 6) * np.array([1, 1, 1, 0.5, 0.5, 0.5])
        est = se3_exp(noise) @ c2w[i]
        if 0 < i < nf - 1:
            st = rng.normal(0, 5e-4, 6)
            state[i, :6] = st
            sc = np.concatenate([SCALE_XI_TRANS * st[:3], SCALE_XI_ROT * st[3:]])
            evalPT[i] = np.linalg.inv(se3_exp(sc)) @ est  # PRE_camToWorld = exp(state_scaled) * evalPT = est
        else:
            evalPT[i] = est
        a_est = aff[i, 0] + (0 if i == 0 else rng.normal(0, 0.01))
        b_est = aff[i, 1] + (0 if i == 0 else rng.normal(0, 1.0))
        state[i, 6] = a_est / SCALE_A
        state[i, 7] = b_est / SCALE_B
    state_zero = state.copy()
    state_zero[:, :6] = 0.0
    for i in range(1, nf - 1):  # affine linearisation point slightly off the current estimate
        state_zero[i, 6] = state[i, 6] - 1e-4
        state_zero[i, 7] = state[i, 7] + 2e-5
    pt_idepth = (pt_id_true * (1.0 + rng.normal(0, idepth_noise, P))).astype(np.float32)

    return Scene(w=w, h=h, nf=nf, K=np.array([fx, fy, cx, cy], np.float64), images=images, camToWorld_true=c2w,
                 evalPT=evalPT, state=state, state_zero=state_zero, aff_true=aff, ab_exposure=np.ones(nf, np.float32),
                 frame_id=np.arange(nf, dtype=np.int32), pt_host=pt_host, pt_u=pt_u, pt_v=pt_v,
                 pt_idepth_true=pt_id_true, pt_idepth=pt_idepth, pt_color=pt_col, pt_weights=pt_wgt,
                 res_point=res_point, res_target=res_target)


def replicate_points(scene: Scene, factor: int, seed: int = 7) -> Scene:
    """Scaling sweep: `factor` x the points, each copy jittered to a nearby pixel of the
    same host so the gather pattern stays realistic."""
    if factor <= 1:
        return scene
    rng = np.random.default_rng(seed)
    P = scene.n_points
    hosts, us, vs, ids, idt, cols, wgts = [], [], [], [], [], [], []
    rp, rt = [], []
    counts = np.bincount(scene.res_point, minlength=P)
    starts = np.concatenate([[0], np.cumsum(counts)])
    for k in range(factor):
        du = rng.integers(-3, 4, size=P).astype(np.float32) if k else np.zeros(P, np.float32)
        dv = rng.integers(-3, 4, size=P).astype(np.float32) if k else np.zeros(P, np.float32)
        hosts.append(scene.pt_host)
        us.append(scene.pt_u + du)
        vs.append(scene.pt_v + dv)
        ids.append(scene.pt_idepth)
        idt.append(scene.pt_idepth_true)
        col = np.empty_like(scene.pt_color)
        for hi in np.unique(scene.pt_host):
            m = scene.pt_host == hi
            pu = (scene.pt_u[m] + du[m]).astype(np.int64)[:, None] + PATTERN[None, :, 0]
            pv = (scene.pt_v[m] + dv[m]).astype(np.int64)[:, None] + PATTERN[None, :, 1]
            col[m] = scene.images[hi][pv, pu]
        cols.append(col)
        wgts.append(scene.pt_weights)
        rp.append(scene.res_point + k * P)
        rt.append(scene.res_target)
    # keep point-major order: copy k's points come after copy k-1's points; residuals likewise
    return dataclasses.replace(
        scene, pt_host=np.concatenate(hosts), pt_u=np.concatenate(us), pt_v=np.concatenate(vs),
        pt_idepth=np.concatenate(ids), pt_idepth_true=np.concatenate(idt), pt_color=np.concatenate(cols),
        pt_weights=np.concatenate(wgts), res_point=np.concatenate(rp).astype(np.int32),
        res_target=np.concatenate(rt).astype(np.int32))


def trace_case(scene: Scene, new_frame: int, n_per_host: int = 200, seed: int = 11, pose_noise: float = 0.0):
    """Inputs of FullSystem::traceNewCoarse (FullSystem.cpp:311-361) for tracing candidates of every other frame of the
    window into `new_frame`: integer candidate pixels per host (random, inside the selector's margin) and the per-host
    KRKi / Kt / affine brightness transfer built from the scene's estimated poses.
    -> dict(host [n], u [n], v [n], KRKi [nf,3,3], Kt [nf,3], aff [nf,2]); host indexes frames (the row of `new_frame`
    itself is the identity and has no points)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = [float(x) for x in scene.K]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
    Ki = np.linalg.inv(K)
    hosts, us, vs = [], [], []
    KRKi = np.zeros((scene.nf, 3, 3), np.float32)
    Kt = np.zeros((scene.nf, 3), np.float32)
    aff = np.zeros((scene.nf, 2), np.float32)
    w2c_new = np.linalg.inv(scene.evalPT[new_frame])
    for hst in range(scene.nf):
        T = w2c_new @ scene.evalPT[hst]
        if pose_noise:
            T = se3_exp(rng.normal(0, pose_noise, 6)) @ T
        KRKi[hst] = (K.astype(np.float32) @ T[:3, :3].astype(np.float32)) @ Ki.astype(np.float32)
        Kt[hst] = K.astype(np.float32) @ T[:3, 3].astype(np.float32)
        # AffLight::fromToVecExposure (util/NumType.h): a = e^{a_t - a_h} * t_exp / h_exp, b = b_t - a * b_h
        a = np.exp(scene.aff_true[new_frame, 0] - scene.aff_true[hst, 0]) * scene.ab_exposure[new_frame] / scene.ab_exposure[hst]
        aff[hst] = (a, scene.aff_true[new_frame, 1] - a * scene.aff_true[hst, 1])
        if hst == new_frame:
            continue
        hosts.append(np.full(n_per_host, hst, np.int32))
        us.append(rng.integers(6, scene.w - 7, n_per_host).astype(np.int32))
        vs.append(rng.integers(6, scene.h - 7, n_per_host).astype(np.int32))
    return dict(host=np.concatenate(hosts), u=np.concatenate(us), v=np.concatenate(vs), KRKi=KRKi, Kt=Kt, aff=aff)


def activation_case(scene: Scene):
    """Per (host, target) FrameFramePrecalc::PRE_RTll / PRE_tTll / PRE_aff_mode (HessianBlocks.cpp:184-214) of the
    window at the scene's estimated poses -> dict(RTll [nf,nf,3,3], tTll [nf,nf,3], aff [nf,nf,2], calib [4])."""
    nf = scene.nf
    RTll = np.zeros((nf, nf, 3, 3), np.float32)
    tTll = np.zeros((nf, nf, 3), np.float32)
    aff = np.zeros((nf, nf, 2), np.float32)
    for hst in range(nf):
        for tgt in range(nf):
            T = np.linalg.inv(scene.evalPT[tgt]) @ scene.evalPT[hst]
            RTll[hst, tgt] = T[:3, :3]
            tTll[hst, tgt] = T[:3, 3]
            a = np.exp(scene.aff_true[tgt, 0] - scene.aff_true[hst, 0]) * scene.ab_exposure[tgt] / scene.ab_exposure[hst]
            aff[hst, tgt] = (a, scene.aff_true[tgt, 1] - a * scene.aff_true[hst, 1])
    return dict(RTll=RTll, tTll=tTll, aff=aff, calib=np.asarray(scene.K, np.float32))


def undistort_case(w, h, w_org=None, h_org=None, seed=3, bits=8, k1=-0.28, k2=0.07):
    """A synthetic camera for the pre-pyramid image path (util/Undistort.cpp): raw frame (w_org x h_org, `bits` bits), a
    rectification map built like Undistort's (radial-tangential model, entries without a source set to -1, :854-884), an
    inverse response G and an inverse vignette.  -> dict(raw, remapX, remapY, G, vignette_inv, w_org, h_org)."""
    rng = np.random.default_rng(seed)
    w_org = w_org or w + 112
    h_org = h_org or h + 32
    ys, xs = np.mgrid[0:h_org, 0:w_org].astype(np.float64)
    tex = 110 + 70 * np.sin(xs * 0.045) * np.cos(ys * 0.06) + 35 * np.sin((xs + 2 * ys) * 0.013) + rng.normal(0, 6, (h_org, w_org))
    top = (1 << bits) - 1
    raw = np.clip(tex * (top / 255.0), 0, top).astype(np.uint8 if bits == 8 else np.uint16)
    depth = 256 if bits == 8 else 65536
    G = (255.0 * (np.arange(depth) / (depth - 1.0)) ** 1.18).astype(np.float32)          # inverse response, G[0] = 0, G[top] = 255
    r2 = ((xs - w_org / 2) ** 2 + (ys - h_org / 2) ** 2) / (w_org / 2) ** 2
    vignette_inv = (1.0 / (1.0 - 0.35 * r2 + 0.05 * r2 * r2)).astype(np.float32)
    # output pinhole K chosen so that the corners fall outside the raw image (exercises the -1 entries)
    fx_o, fy_o, cx_o, cy_o = 0.62 * w_org, 0.62 * w_org, w_org / 2 - 0.5, h_org / 2 - 0.5
    fx, fy, cx, cy = 0.50 * w, 0.50 * w, w / 2 - 0.5, h / 2 - 0.5
    yo, xo = np.mgrid[0:h, 0:w].astype(np.float32)
    xn = (xo - np.float32(cx)) / np.float32(fx)
    yn = (yo - np.float32(cy)) / np.float32(fy)
    rr = xn * xn + yn * yn
    fac = 1 + np.float32(k1) * rr + np.float32(k2) * rr * rr
    ix = (np.float32(fx_o) * (xn * fac) + np.float32(cx_o)).astype(np.float32)
    iy = (np.float32(fy_o) * (yn * fac) + np.float32(cy_o)).astype(np.float32)
    ok = (ix > 0) & (iy > 0) & (ix < w_org - 1) & (iy < h_org - 1)
    remapX = np.where(ok, ix, np.float32(-1)).astype(np.float32)
    remapY = np.where(ok, iy, np.float32(-1)).astype(np.float32)
    return dict(raw=raw, remapX=remapX, remapY=remapY, G=G, vignette_inv=vignette_inv, w_org=w_org, h_org=h_org)


def loop_case(scene: Scene, matched: int, levels: int, level_images, n: int = 3000, seed: int = 9):
    """PoseEstimator inputs built like LoopHandler::addKeyFrame (LoopClosure/LoopHandler.cpp:186-209): 3D points of the
    matched frame in its camera frame and their colour on every pyramid level (bilinear tap of channel 0 at the level's
    pixel).  level_images[l] = (h_l, w_l) float32 intensity of the matched frame at level l.
    -> dict(xyz [n,3] float64, color [n, levels] float32)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = [float(x) for x in scene.K]
    sel = np.flatnonzero(scene.pt_host == matched)
    if sel.size == 0:
        sel = np.arange(scene.n_points)
    idx = rng.choice(sel, n, replace=True)
    u = scene.pt_u[idx].astype(np.float64) + rng.integers(-3, 4, n)
    v = scene.pt_v[idx].astype(np.float64) + rng.integers(-3, 4, n)
    u = np.clip(u, 4, scene.w - 5); v = np.clip(v, 4, scene.h - 5)
    d = scene.pt_idepth[idx].astype(np.float64) * rng.uniform(0.99, 1.01, n)
    xyz = np.stack([(u - cx) / fx / d, (v - cy) / fy / d, 1.0 / d], 1)
    color = np.zeros((n, levels), np.float32)
    for l in range(levels):
        img = np.asarray(level_images[l], np.float32)
        ul = ((u + 0.5) / (1 << l) - 0.5).astype(np.float32)
        vl = ((v + 0.5) / (1 << l) - 0.5).astype(np.float32)
        ix = ul.astype(np.int32); iy = vl.astype(np.int32)
        dx = ul - ix; dy = vl - iy
        dxdy = dx * dy
        color[:, l] = dxdy * img[iy + 1, ix + 1] + (dy - dxdy) * img[iy + 1, ix] + (dx - dxdy) * img[iy, ix + 1] + (1 - dx - dy + dxdy) * img[iy, ix]
    return dict(xyz=xyz, color=color)


def init_case(scene: Scene, lvl: int, n: int = 2500, seed: int = 13):
    """CoarseInitializer::Pnt arrays of one level in the state trackFrame leaves them between iterations: integer pixels of
    the first frame (level coordinates), inverse depths around 1 (the initializer's normalisation), some points already
    marked bad.  -> dict(u, v, idepth_new, iR, energy [n,2], outlierTH, isGood)."""
    rng = np.random.default_rng(seed)
    wl, hl = scene.w >> lvl, scene.h >> lvl
    u = rng.integers(5, wl - 6, n).astype(np.float32)
    v = rng.integers(5, hl - 6, n).astype(np.float32)
    return dict(u=u, v=v, idepth_new=rng.uniform(0.7, 1.3, n).astype(np.float32), iR=rng.uniform(0.8, 1.2, n).astype(np.float32),
                energy=np.stack([rng.uniform(0, 400, n), rng.uniform(0, 0.1, n)], 1).astype(np.float32),
                outlierTH=np.full(n, 8 * 12 * 12, np.float32) * rng.choice([1.0, 0.02], n, p=[0.9, 0.1]).astype(np.float32),
                isGood=(rng.uniform(0, 1, n) > 0.1).astype(np.uint8))
