"""Second, independent restatement (numpy) of parts of the reference path, written from the reference
sources rather than from oracle/ — used by the CPU tests to cross-check the C++ oracle.  TEST INFRASTRUCTURE.

  pyramid          FrameHessian::makeImages            src/FullSystem/HessianBlocks.cpp:121-176
  window_tables    FrameFramePrecalc::set              src/FullSystem/HessianBlocks.cpp:431-461
                   EnergyFunctional::setAdjointsF      src/OptimizationBackend/EnergyFunctional.cpp:42-103
                   EnergyFunctional::setDeltaF         src/OptimizationBackend/EnergyFunctional.cpp:163-194
  linearize        PointFrameResidual::linearize       src/FullSystem/Residuals.cpp:77-271
  dense_system     what accumulateAF/SCF + stitch compute, as plain fp64 dense algebra
                   (AccumulatedTopHessian.cpp:35-147,231-301; AccumulatedSCHessian.cpp:32-158)

float32 arithmetic is kept in float32 with the reference's operation order (numpy does not contract to FMA),
so `linearize` is expected to agree with the oracle bit for bit when both are given the same window tables.
"""
import numpy as np

F = np.float32
PATTERN = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], dtype=np.int32)  # settings.cpp:307-309
SCALE_XI_TRANS, SCALE_XI_ROT, SCALE_A, SCALE_B, SCALE_F, SCALE_C = 0.5, 1.0, 10.0, 1000.0, 50.0, 50.0
PRECALC_FLOATS = 32
RES_IN, RES_OOB, RES_OUTLIER = 0, 1, 2


# ---- a1 ------------------------------------------------------------------------------------------
def pyramid(img, levels, B=None, gamma=1):
    """Returns [(dI (h,w,3) float32, absSquaredGrad (h,w) float32)] per level; first/last row gradients 0."""
    out = []
    I = np.asarray(img, F)
    for lvl in range(levels):
        if lvl > 0:
            p = out[-1][0][..., 0]
            h, w = p.shape[0] // 2, p.shape[1] // 2
            p = p[:2 * h, :2 * w]
            I = F(0.25) * (((p[0::2, 0::2] + p[0::2, 1::2]) + p[1::2, 0::2]) + p[1::2, 1::2])
        h, w = I.shape
        flat = I.reshape(-1)
        dx = np.zeros(w * h, F)
        dy = np.zeros(w * h, F)
        idx = np.arange(w, w * (h - 1))
        dx[idx] = F(0.5) * (flat[idx + 1] - flat[idx - 1])     # flat index: crosses row boundaries at the first/last column
        dy[idx] = F(0.5) * (flat[idx + w] - flat[idx - w])
        dx[~np.isfinite(dx)] = 0
        dy[~np.isfinite(dy)] = 0
        ab = dx * dx + dy * dy
        if B is not None and gamma == 1:
            c = (flat + F(0.5)).astype(np.int32)
            c = np.clip(c, 5, 250)
            gw = np.asarray(B, F)[c + 1] - np.asarray(B, F)[c]
            ab2 = ab * (gw * gw)
            ab[idx] = ab2[idx]
        dI = np.stack([flat, dx, dy], axis=1).reshape(h, w, 3)
        out.append((dI, ab.reshape(h, w)))
    return out


# ---- SE3 (thirdparty/Sophus/sophus/se3.hpp, so3.hpp) ------------------------------------------------
def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], np.float64)


def se3_exp(xi):
    u, w = np.asarray(xi[:3], np.float64), np.asarray(xi[3:6], np.float64)
    th = np.linalg.norm(w)
    K = hat(w)
    if th < 1e-10:
        R = np.eye(3) + K
        V = R
    else:
        R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * (K @ K)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ u
    return T


def se3_inv(T):
    o = np.eye(4)
    o[:3, :3] = T[:3, :3].T
    o[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return o


def se3_adj(T):
    R, t = T[:3, :3], T[:3, 3]
    A = np.zeros((6, 6))
    A[:3, :3] = R
    A[3:, 3:] = R
    A[:3, 3:] = hat(t) @ R
    return A


def to44(T34):
    T = np.eye(4)
    T[:3, :4] = np.asarray(T34, np.float64).reshape(3, 4)
    return T


def aff_from_to(expF, expT, aF, bF, aT, bT):
    """AffLight::fromToVecExposure (util/NumType.h:157-168)."""
    expF, expT = F(expF), F(expT)
    if expF == 0 or expT == 0:
        expF = expT = F(1)
    a = np.exp(aT - aF) * float(expT) / float(expF)
    return a, bT - a * bF


def _mm3(A, B):
    """float32 3x3 product with the coefficient order ((a0 b0 + a1 b1) + a2 b2)."""
    A, B = np.asarray(A, F), np.asarray(B, F)
    C = np.zeros((3, 3), F)
    for i in range(3):
        for j in range(3):
            C[i, j] = (A[i, 0] * B[0, j] + A[i, 1] * B[1, j]) + A[i, 2] * B[2, j]
    return C


def window_tables(frames, calib_value, calib_value_zero, cfg_priors=None):
    """frames: list of dicts as problem.frames_of.  Returns the dict Handle.window_set takes."""
    nf = len(frames)
    ev = [to44(f["evalPT"]) for f in frames]
    st = [np.asarray(f["state"], np.float64) for f in frames]
    st0 = [np.asarray(f["state_zero"], np.float64) for f in frames]
    scaled = [np.array([SCALE_XI_TRANS * s[0], SCALE_XI_TRANS * s[1], SCALE_XI_TRANS * s[2], SCALE_XI_ROT * s[3], SCALE_XI_ROT * s[4],
                        SCALE_XI_ROT * s[5], SCALE_A * s[6], SCALE_B * s[7]]) for s in st]
    c2w = [se3_exp(scaled[i][:6]) @ ev[i] for i in range(nf)]
    w2c = [se3_inv(T) for T in c2w]
    cal = np.array([SCALE_F * calib_value[0], SCALE_F * calib_value[1], SCALE_C * calib_value[2], SCALE_C * calib_value[3]])
    calf = cal.astype(F)
    fx, fy, cx, cy = calf
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], F)
    Ki = np.array([[F(1) / fx, 0, -cx / fx], [0, F(1) / fy, -cy / fy], [0, 0, 1]], F)
    precalc = np.zeros((nf, nf, PRECALC_FLOATS), F)
    adHost = np.zeros((nf * nf, 8, 8))
    adTarget = np.zeros((nf * nf, 8, 8))
    adHTdeltaF = np.zeros((nf * nf, 8), F)
    for h in range(nf):
        for t in range(nf):
            l0 = se3_inv(ev[t]) @ ev[h]
            l = w2c[t] @ c2w[h]
            pc = precalc[h, t]
            pc[0:9] = l0[:3, :3].astype(F).reshape(9)
            pc[9:12] = l0[:3, 3].astype(F)
            R = l[:3, :3].astype(F)
            tt = l[:3, 3].astype(F)
            pc[12:21] = _mm3(_mm3(K, R), Ki).reshape(9)
            for i in range(3):
                pc[21 + i] = (K[i, 0] * tt[0] + K[i, 1] * tt[1]) + K[i, 2] * tt[2]
            a, b = aff_from_to(frames[h]["ab_exposure"], frames[t]["ab_exposure"], scaled[h][6], scaled[h][7], scaled[t][6], scaled[t][7])
            pc[24], pc[25] = F(a), F(b)
            pc[26] = F(st0[h][7] * SCALE_B)
            pc[27] = F(np.linalg.norm(l[:3, 3]))
            # adjoints (EnergyFunctional.cpp:56-83)
            AH, AT = np.eye(8), np.eye(8)
            adjT = se3_adj(se3_inv(ev[t])).T
            AH[:6, :6] = adjT
            AT[:6, :6] = -adjT
            a0, _ = aff_from_to(frames[h]["ab_exposure"], frames[t]["ab_exposure"], st0[h][6] * SCALE_A, st0[h][7] * SCALE_B,
                                st0[t][6] * SCALE_A, st0[t][7] * SCALE_B)
            a0 = float(F(a0))
            AT[6, 6] = -a0
            AH[6, 6] = a0
            AT[7, 7] = -1
            AH[7, 7] = a0
            for M in (AH, AT):
                M[0:3] *= SCALE_XI_TRANS
                M[3:6] *= SCALE_XI_ROT
                M[6] *= SCALE_A
                M[7] *= SCALE_B
            idx = h + t * nf
            adHost[idx], adTarget[idx] = AH, AT
            dh = (st[h] - st0[h])[:8].astype(F)
            dt = (st[t] - st0[t])[:8].astype(F)
            AHf, ATf = AH.astype(F), AT.astype(F)
            sh = np.zeros(8, F)
            stt = np.zeros(8, F)
            for k in range(8):
                sh = sh + dh[k] * AHf[k]
                stt = stt + dt[k] * ATf[k]
            adHTdeltaF[idx] = sh + stt
    pri = cfg_priors or {}
    frame_prior = np.zeros((nf, 8))
    for i, f in enumerate(frames):
        if int(f.get("frame_id", i)) == 0:   # FrameHessian::getPrior for the first frame (HessianBlocks.h:301-323)
            frame_prior[i, 0:3] = pri.get("initial_trans_prior", 1e10)
            frame_prior[i, 3:6] = pri.get("initial_rot_prior", 1e11)
            frame_prior[i, 6] = pri.get("initial_aff_a_prior", 1e14)
            frame_prior[i, 7] = pri.get("initial_aff_b_prior", 1e14)
    return {"nf": nf, "frame_slot": np.array([f.get("slot", i) for i, f in enumerate(frames)], np.int32),
            "precalc": precalc.reshape(-1), "adHost": adHost.reshape(-1), "adTarget": adTarget.reshape(-1),
            "adHTdeltaF": adHTdeltaF.reshape(-1),
            "frame_energy_th": np.array([f.get("frame_energy_th", 512.0) for f in frames], F),
            "calib": calf, "cDeltaF": (np.asarray(calib_value) - np.asarray(calib_value_zero)).astype(F),
            "cPrior": np.full(4, pri.get("initial_calib_hessian", 5e9)),
            "frame_prior": frame_prior.reshape(-1), "frame_delta_prior": np.array([s[:8] for s in st]).reshape(-1),
            "frame_delta": np.array([(s - s0)[:8] for s, s0 in zip(st, st0)]).reshape(-1)}


# ---- a3 ------------------------------------------------------------------------------------------
def _interp33(img, x, y):
    """getInterpolatedElement33 (util/globalFuncs.h:68-82); img (h,w,3); x,y float32 arrays."""
    ix = x.astype(np.int32)
    iy = y.astype(np.int32)
    dx = x - ix.astype(F)
    dy = y - iy.astype(F)
    dxdy = dx * dy
    t11, t01, t10, t00 = img[iy + 1, ix + 1], img[iy + 1, ix], img[iy, ix + 1], img[iy, ix]
    w11, w01, w10, w00 = dxdy, dy - dxdy, dx - dxdy, F(1) - dx - dy + dxdy
    return w11[..., None] * t11 + w01[..., None] * t01 + w10[..., None] * t10 + w00[..., None] * t00


def linearize(win, pts, res, images_dI, w, h, frame_energy_th=None, huber=9.0, outlier_sum=2500.0):
    """All residuals with state IN at entry.  Returns dict(new_state, new_energy, new_energy_wo, J (R,74), projectedTo, center)."""
    nf = win["nf"]
    precalc = np.asarray(win["precalc"], F).reshape(nf, nf, PRECALC_FLOATS)
    fxl, fyl, cxl, cyl = [F(x) for x in win["calib"]]
    fxli, fyli = F(1) / fxl, F(1) / fyl
    th_f = np.asarray(win["frame_energy_th"] if frame_energy_th is None else frame_energy_th, F)
    rp = np.asarray(res["point"])
    rt = np.asarray(res["target"])
    rh = np.asarray(pts["host"])[rp]
    R = len(rp)
    pc = precalc[rh, rt]
    R0, t0, KRKi, Kt = pc[:, 0:9], pc[:, 9:12], pc[:, 12:21], pc[:, 21:24]
    aff0, aff1, b0 = pc[:, 24], pc[:, 25], pc[:, 26]
    pu, pv = np.asarray(pts["u"], F)[rp], np.asarray(pts["v"], F)[rp]
    idz, idp = np.asarray(pts["idepth_zero"], F)[rp], np.asarray(pts["idepth"], F)[rp]
    color, weights = np.asarray(pts["color"], F).reshape(-1, 8)[rp], np.asarray(pts["weights"], F).reshape(-1, 8)[rp]
    wM3, hM3 = F(w - 3), F(h - 3)
    with np.errstate(all="ignore"):
        # centre projection (ResidualProjections.h:43-73)
        Kl0, Kl1, Kl2 = (pu + F(0) - cxl) * fxli, (pv + F(0) - cyl) * fyli, np.ones(R, F)
        ptp = [((R0[:, 3 * i] * Kl0 + R0[:, 3 * i + 1] * Kl1) + R0[:, 3 * i + 2] * Kl2) + t0[:, i] * idz for i in range(3)]
        dres = F(1) / ptp[2]
        nid = idz * dres
        u, v = ptp[0] * dres, ptp[1] * dres
        Ku, Kv = u * fxl + cxl, v * fyl + cyl
        ok_c = (dres > 0) & (Ku > F(1.1)) & (Kv > F(1.1)) & (Ku < wM3) & (Kv < hM3)
        J = np.zeros((R, 74), F)
        d_d_x = dres * (t0[:, 0] - t0[:, 2] * u) * F(1) * fxl
        d_d_y = dres * (t0[:, 1] - t0[:, 2] * v) * F(1) * fyl
        cx2 = dres * (R0[:, 6] * u - R0[:, 0])
        cx3 = fxl * dres * (R0[:, 7] * u - R0[:, 1]) * fyli
        cx0, cx1 = Kl0 * cx2, Kl1 * cx3
        cy2 = fyl * dres * (R0[:, 6] * v - R0[:, 3]) * fxli
        cy3 = dres * (R0[:, 7] * v - R0[:, 4])
        cy0, cy1 = Kl0 * cy2, Kl1 * cy3
        SF, SC = F(SCALE_F), F(SCALE_C)
        Jpdc0 = np.stack([(cx0 + u) * SF, cx1 * SF, (cx2 + F(1)) * SC, cx3 * SC], 1)
        Jpdc1 = np.stack([cy0 * SF, (cy1 + v) * SF, cy2 * SC, (cy3 + F(1)) * SC], 1)
        z = np.zeros(R, F)
        Jpdxi0 = np.stack([nid * fxl, z, -nid * u * fxl, -u * v * fxl, (F(1) + u * u) * fxl, -v * fxl], 1)
        Jpdxi1 = np.stack([z, nid * fyl, -nid * v * fyl, -(F(1) + v * v) * fyl, u * v * fyl, u * fyl], 1)
        J[:, 8:14], J[:, 14:20], J[:, 20:24], J[:, 24:28] = Jpdxi0, Jpdxi1, Jpdc0, Jpdc1
        J[:, 28], J[:, 29] = d_d_x, d_d_y
        # pattern
        up = pu[:, None] + PATTERN[None, :, 0].astype(F)
        vp = pv[:, None] + PATTERN[None, :, 1].astype(F)
        q = [((KRKi[:, 3 * i, None] * up + KRKi[:, 3 * i + 1, None] * vp) + KRKi[:, 3 * i + 2, None] * F(1)) + Kt[:, i, None] * idp[:, None] for i in range(3)]
        Kup, Kvp = q[0] / q[2], q[1] / q[2]
        ok_p = (Kup > F(1.1)) & (Kvp > F(1.1)) & (Kup < wM3) & (Kvp < hM3)
        hit = np.zeros((R, 8, 3), F)
        for t in range(nf):
            m = (rt[:, None] == t) & ok_p
            if m.any():
                hit[m] = _interp33(images_dI[t], Kup[m], Kvp[m])
        ok_p &= np.isfinite(hit[..., 0])
        ok = ok_c & ok_p.all(axis=1)
        residual = hit[..., 0] - (aff0[:, None] * color + aff1[:, None]).astype(F)
        drdA = color - b0[:, None]
        c = F(outlier_sum)
        wgt = np.sqrt(c / (c + (hit[..., 1] * hit[..., 1] + hit[..., 2] * hit[..., 2])))
        wgt = F(0.5) * (wgt + weights)
        hth = F(huber)
        hw = np.where(np.abs(residual) < hth, F(1), hth / np.abs(residual)).astype(F)
        e_terms = wgt * wgt * hw * residual * residual * (F(2) - hw)
        hw = np.where(hw < 1, np.sqrt(hw), hw).astype(F)
        hw = hw * wgt
        hx, hy = hit[..., 1] * hw, hit[..., 2] * hw

        def seq(a):
            s = np.zeros(R, F)
            for i in range(8):
                s = s + a[:, i]
            return s
        energy = seq(e_terms)
        J[:, 0:8] = residual * hw
        J[:, 30:38], J[:, 38:46] = hx, hy
        J[:, 46:54], J[:, 54:62] = drdA * hw, hw
        J00, J11, J10 = seq(hx * hx), seq(hy * hy), seq(hx * hy)
        J[:, 62], J[:, 63], J[:, 64], J[:, 65] = J00, J10, J10, J11
        J[:, 66], J[:, 67], J[:, 68], J[:, 69] = seq(drdA * hw * hx), seq(drdA * hw * hy), seq(hw * hx), seq(hw * hy)
        A00, A01, A11 = seq(drdA * drdA * hw * hw), seq(drdA * hw * hw), seq(hw * hw)
        J[:, 70], J[:, 71], J[:, 72], J[:, 73] = A00, A01, A01, A11
        wJI2 = seq(hw * hw * (hx * hx + hy * hy))
    th = np.maximum(th_f[rh], th_f[rt])
    outl = (energy > th) | (wJI2 < 2)
    new_state = np.where(ok, np.where(outl, RES_OUTLIER, RES_IN), RES_OOB).astype(np.uint8)
    new_energy = np.where(outl, th, energy).astype(F)
    return {"new_state": new_state, "new_energy": new_energy, "new_energy_wo": np.where(ok, energy, F(-1)).astype(F), "J": J,
            "projectedTo": np.stack([Kup, Kvp], -1), "center": np.stack([Ku, Kv, nid], 1), "host": rh}


# ---- a6-a10 as dense fp64 algebra --------------------------------------------------------------------
def dense_system(win, pts, res, J, active, n_points):
    """Full Gauss-Newton normal equations over [calib(4) | frames(8 nf) | idepth(P)] from per-residual J records
    (RawResidualJacobian order), then the split the reference makes: top (H, b) and Schur (Hsc, bsc)."""
    nf = win["nf"]
    D = 4 + 8 * nf
    adH = np.asarray(win["adHost"]).reshape(nf * nf, 8, 8)
    adT = np.asarray(win["adTarget"]).reshape(nf * nf, 8, 8)
    rp, rt = np.asarray(res["point"]), np.asarray(res["target"])
    rh = np.asarray(pts["host"])[rp]
    J = np.asarray(J, np.float64)
    H = np.zeros((D, D))
    b = np.zeros(D)
    Hcd = np.zeros((n_points, D))
    Hdd = np.zeros(n_points)
    bd = np.zeros(n_points)
    for r in np.nonzero(active)[0]:
        j = J[r]
        resF, Jxi, Jc, Jd = j[0:8], j[8:20].reshape(2, 6), j[20:28].reshape(2, 4), j[28:30]
        JI, Jab = j[30:46].reshape(2, 8), j[46:62].reshape(2, 8)
        Jrel = np.zeros((8, 13))                       # d r_i / d [calib, xi_rel, a, b, idepth]
        Jrel[:, 0:4] = JI.T @ Jc
        Jrel[:, 4:10] = JI.T @ Jxi
        Jrel[:, 10:12] = Jab.T
        Jrel[:, 12] = JI.T @ Jd
        idx = rh[r] + rt[r] * nf
        Jg = np.zeros((8, D))
        Jg[:, 0:4] = Jrel[:, 0:4]
        Jg[:, 4 + 8 * rh[r]:12 + 8 * rh[r]] += Jrel[:, 4:12] @ adH[idx].T
        Jg[:, 4 + 8 * rt[r]:12 + 8 * rt[r]] += Jrel[:, 4:12] @ adT[idx].T
        H += Jg.T @ Jg
        b += Jg.T @ resF
        p = rp[r]
        Hcd[p] += Jg.T @ Jrel[:, 12]
        Hdd[p] += Jrel[:, 12] @ Jrel[:, 12]
        bd[p] += Jrel[:, 12] @ resF
    return H, b, Hcd, Hdd, bd


def schur(Hcd, Hdd, bd, prior=0.0):
    Hd = Hdd + prior
    m = Hdd > 0
    Hdi = np.where(m, 1.0 / np.maximum(Hd, 1e-10), 0.0)
    Hsc = (Hcd.T * Hdi) @ Hcd
    bsc = Hcd.T @ (Hdi * bd)
    return Hsc, bsc


# ---- immature points ----------------------------------------------------------------
# Independent restatement of ImmaturePoint::ImmaturePoint (ImmaturePoint.cpp:28-60) and ImmaturePoint::traceOn
# (ImmaturePoint.cpp:70-415) in float32 numpy; the epipolar search is vectorised over steps x pattern (positions by a
# sequential float32 cumsum, like the reference's ptx += dx), the Gauss-Newton refinement is a plain loop.
F = np.float32
IPS_GOOD, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED = range(6)


def _bilin(dI, x, y):
    """getInterpolatedElement33 (globalFuncs.h:68-82) on dI (h, w, 3) at float32 positions (any shape) -> (..., 3)."""
    x = np.asarray(x, F); y = np.asarray(y, F)
    ix = x.astype(np.int32); iy = y.astype(np.int32)
    dx = x - ix.astype(F); dy = y - iy.astype(F)
    dxdy = dx * dy
    w11, w01, w10, w00 = dxdy, dy - dxdy, dx - dxdy, F(1) - dx - dy + dxdy
    return (w11[..., None] * dI[iy + 1, ix + 1] + w01[..., None] * dI[iy + 1, ix] + w10[..., None] * dI[iy, ix + 1] + w00[..., None] * dI[iy, ix])


def immature_init_ref(dI, u, v, outlier_th_sum=F(2500), overall_weight=F(1)):
    """dI (h, w, 3) float32; integer pixels u, v -> color (n,8), weights (n,8), gradH (n,4), energyTH (n,)."""
    n = len(u)
    color = np.zeros((n, 8), F); weights = np.zeros((n, 8), F); G = np.zeros((n, 4), F)
    for idx, (px, py) in enumerate(PATTERN):
        x = u + px; y = v + py
        tl, tr, bl, br = dI[y, x, 0], dI[y, x + 1, 0], dI[y + 1, x, 0], dI[y + 1, x + 1, 0]
        # integer position: dx = dy = 0 in getInterpolatedElement33BiLin (globalFuncs.h:161-182)
        z, o = F(0), F(1)
        top = z * tr + o * tl; bot = z * br + o * bl; left = z * bl + o * tl; right = z * br + o * tr
        c0 = z * right + o * left; gx = right - left; gy = bot - top
        color[:, idx] = c0
        G[:, 0] += gx * gx; G[:, 1] += gx * gy; G[:, 2] += gy * gx; G[:, 3] += gy * gy
        weights[:, idx] = np.sqrt(outlier_th_sum / (outlier_th_sum + (gx * gx + gy * gy)))
    eth = np.full(n, F(8) * F(144) * (overall_weight * overall_weight), F)
    return color, weights, G, eth


def trace_on_ref(dI, p, KRKi, Kt, aff, huber=F(9)):
    """One point.  p: dict(u, v, color[8], weights[8], gradH[4], energy_th, idepth_min, idepth_max, quality, status);
    returns the updated dict (plus uv, pixel_interval)."""
    p = dict(p)
    if p["status"] == IPS_OOB:
        return p
    h, w = dI.shape[:2]
    KRKi = np.asarray(KRKi, F).reshape(3, 3); Kt = np.asarray(Kt, F); aff = np.asarray(aff, F)

    def leave(status, uv=(-1, -1), interval=0):
        p["status"] = status; p["uv"] = (F(uv[0]), F(uv[1])); p["pixel_interval"] = F(interval)
        return p

    def inside(a, b):
        return a > 4 and b > 4 and a < w - 5 and b < h - 5

    max_search = F(w + h) * F(0.027)
    pr = (KRKi[:, 0] * F(p["u"]) + KRKi[:, 1] * F(p["v"])) + KRKi[:, 2]
    pmin = pr + Kt * F(p["idepth_min"])
    uMin, vMin = pmin[0] / pmin[2], pmin[1] / pmin[2]
    if not inside(uMin, vMin):
        return leave(IPS_OOB)
    finite_max = bool(np.isfinite(p["idepth_max"]))
    if finite_max:
        pmax = pr + Kt * F(p["idepth_max"])
        uMax, vMax = pmax[0] / pmax[2], pmax[1] / pmax[2]
        if not inside(uMax, vMax):
            return leave(IPS_OOB)
        dist = np.sqrt((uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax))
        if dist < F(1.5):
            return leave(IPS_SKIPPED, ((uMax + uMin) * F(0.5), (vMax + vMin) * F(0.5)), dist)
    else:
        dist = max_search
        pmax = pr + Kt * F(0.01)
        uMax, vMax = pmax[0] / pmax[2], pmax[1] / pmax[2]
        ex, ey = uMax - uMin, vMax - vMin
        d = F(1) / np.sqrt(ex * ex + ey * ey)
        uMax = uMin + dist * ex * d; vMax = vMin + dist * ey * d
        if not inside(uMax, vMax):
            return leave(IPS_OOB)
    if not (p["idepth_min"] < 0 or (pmin[2] > F(0.75) and pmin[2] < F(1.5))):
        return leave(IPS_OOB)
    dx, dy = uMax - uMin, vMax - vMin
    G = np.asarray(p["gradH"], F)
    a = (dx * G[0] + dy * G[2]) * dx + (dx * G[1] + dy * G[3]) * dy
    b = (dy * G[0] + (-dx) * G[2]) * dy + (dy * G[1] + (-dx) * G[3]) * (-dx)
    with np.errstate(all="ignore"):
        err_px = F(0.2) + F(0.2) * (a + b) / a
    if err_px * F(2) > dist and finite_max:
        return leave(IPS_BADCONDITION, ((uMax + uMin) * F(0.5), (vMax + vMin) * F(0.5)), dist)
    if err_px > 10:
        err_px = F(10)
    dx = dx / dist; dy = dy / dist
    if dist > max_search:
        dist = max_search
    nsteps = int(F(1.9999) + dist / F(1))
    shift = uMin * F(1000) - np.floor(uMin * F(1000))
    x0, y0 = uMin - shift * dx, vMin - shift * dy
    pat = np.asarray(PATTERN, F)
    rx = KRKi[0, 0] * pat[:, 0] + KRKi[0, 1] * pat[:, 1]
    ry = KRKi[1, 0] * pat[:, 0] + KRKi[1, 1] * pat[:, 1]
    if not (np.isfinite(dx) and np.isfinite(dy)):
        return leave(IPS_OOB)
    nsteps = min(nsteps, 99)
    xs = np.cumsum(np.concatenate([[x0], np.full(nsteps - 1, dx, F)]).astype(F), dtype=F)
    ys = np.cumsum(np.concatenate([[y0], np.full(nsteps - 1, dy, F)]).astype(F), dtype=F)
    color = np.asarray(p["color"], F); wts = np.asarray(p["weights"], F)
    ref = aff[0] * color + aff[1]
    hit = _bilin(dI, xs[:, None] + rx[None, :], ys[:, None] + ry[None, :])[..., 0]
    r = hit - ref[None, :]
    hw = np.where(np.abs(r) < huber, F(1), huber / np.maximum(np.abs(r), F(1e-30))).astype(F)
    term = np.where(np.isfinite(hit), hw * r * r * (F(2) - hw), F(1e5)).astype(F)
    energies = np.zeros(nsteps, F)
    for idx in range(8):
        energies = energies + term[:, idx]
    best = int(np.argmin(energies))        # first minimum, like the strict '<' of the reference
    bestU, bestV, bestE = xs[best], ys[best], energies[best]
    far = np.abs(np.arange(nsteps) - best) > 2
    second = energies[far].min() if far.any() else F(1e10)
    second = min(second, F(1e10))
    with np.errstate(all="ignore"):
        new_q = F(second) / bestE
    if new_q < p["quality"] or nsteps > 10:
        p["quality"] = new_q
    uBak, vBak, back = bestU, bestV, F(0)
    bestE = F(1e5)
    for _ in range(3):
        H, bb, E = F(1), F(0), F(0)
        for idx in range(8):
            hit3 = _bilin(dI, F(bestU + rx[idx]), F(bestV + ry[idx]))
            if not np.isfinite(hit3[0]):
                E += F(1e5)
                continue
            res = hit3[0] - (aff[0] * color[idx] + aff[1])
            dd = dx * hit3[1] + dy * hit3[2]
            hw1 = F(1) if abs(res) < huber else huber / abs(res)
            H += hw1 * dd * dd
            bb += hw1 * res * dd
            E += wts[idx] * wts[idx] * hw1 * res * res * (F(2) - hw1)
        if E > bestE:
            back = back * F(0.5)
            bestU = uBak + back * dx; bestV = vBak + back * dy
        else:
            step = F(-1) * bb / H
            step = min(max(step, F(-0.5)), F(0.5))
            if not np.isfinite(step):
                step = F(0)
            uBak, vBak, back = bestU, bestV, step
            bestU = bestU + step * dx; bestV = bestV + step * dy
            bestE = E
        if abs(back) < F(0.1):
            break
    if not (bestE < F(p["energy_th"]) * F(1.2)):
        return leave(IPS_OOB if p["status"] == IPS_OUTLIER else IPS_OUTLIER)
    with np.errstate(all="ignore"):
        if dx * dx > dy * dy:
            lo = (pr[2] * (bestU - err_px * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU - err_px * dx))
            hi = (pr[2] * (bestU + err_px * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU + err_px * dx))
        else:
            lo = (pr[2] * (bestV - err_px * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV - err_px * dy))
            hi = (pr[2] * (bestV + err_px * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV + err_px * dy))
    if lo > hi:
        lo, hi = hi, lo
    p["idepth_min"], p["idepth_max"] = F(lo), F(hi)
    if not (np.isfinite(lo) and np.isfinite(hi)) or hi < 0:
        return leave(IPS_OUTLIER)
    return leave(IPS_GOOD, (bestU, bestV), F(2) * err_px)


ACT_SKIP, ACT_ACTIVATED, ACT_DELETE = 0, 1, -1


def _linearize_residual_ref(dI, p, R, t, affLL, calib, slack, tr, acc, idepth, huber=F(9)):
    """ImmaturePoint::linearizeResidual (ImmaturePoint.cpp:475-545).  tr: dict(state, energy, new_state, new_energy);
    acc: [Hdd, bd] accumulated in place (partial sums of an aborted pattern stay, as in the reference)."""
    if tr["state"] == 1:
        tr["new_state"] = 1
        return np.float64(tr["energy"])
    h, w = dI.shape[:2]
    fx, fy, cx, cy = [F(x) for x in calib]
    fxi, fyi = F(1) / fx, F(1) / fy
    E = F(0)
    for idx, (px, py) in enumerate(PATTERN):
        Kl = np.array([(F(p["u"]) + F(px) - cx) * fxi, (F(p["v"]) + F(py) - cy) * fyi, F(1)], F)
        ptp = ((R[:, 0] * Kl[0] + R[:, 1] * Kl[1]) + R[:, 2] * Kl[2]) + t * idepth
        with np.errstate(all="ignore"):
            dres = F(1) / ptp[2]
        if not dres > 0:
            tr["new_state"] = 1
            return np.float64(tr["energy"])
        u, v = ptp[0] * dres, ptp[1] * dres
        Ku, Kv = u * fx + cx, v * fy + cy
        if not (Ku > F(1.1) and Kv > F(1.1) and Ku < F(w - 3) and Kv < F(h - 3)):
            tr["new_state"] = 1
            return np.float64(tr["energy"])
        hit = _bilin(dI, Ku, Kv)
        if not np.isfinite(hit[0]):
            tr["new_state"] = 1
            return np.float64(tr["energy"])
        res = hit[0] - (affLL[0] * F(p["color"][idx]) + affLL[1])
        hw = F(1) if abs(res) < huber else huber / abs(res)
        wt = F(p["weights"][idx])
        E = E + wt * wt * hw * res * res * (F(2) - hw)
        dxi, dyi = hit[1] * fx, hit[2] * fy
        dd = dxi * dres * (t[0] - t[2] * u) + dyi * dres * (t[1] - t[2] * v)
        hw = hw * (wt * wt)
        acc[0] = acc[0] + (hw * dd) * dd
        acc[1] = acc[1] + (hw * res) * dd
    lim = F(p["energy_th"]) * F(slack)
    if E > lim:
        E = lim
        tr["new_state"] = 2
    else:
        tr["new_state"] = 0
    tr["new_energy"] = np.float64(E)
    return np.float64(E)


def optimize_immature_ref(dIs, p, host, RTll, tTll, aff, calib, min_obs=1):
    """FullSystem::optimizeImmaturePoint (FullSystemOptPoint.cpp:47-192) for one point.  dIs: list of (h, w, 3) images per
    frame.  -> (result, idepth, states[nf])."""
    nf = len(dIs)
    targets = [f for f in range(nf) if f != host]
    trs = {f: dict(state=0, energy=np.float64(0), new_state=2, new_energy=np.float64(0)) for f in targets}

    def sweep(slack, idepth):
        acc = [F(0), F(0)]
        E = F(0)
        for f in targets:
            E = F(np.float64(E) + _linearize_residual_ref(dIs[f], p, RTll[host, f], tTll[host, f], aff[host, f], calib, slack, trs[f], acc, idepth))
        return E, acc[0], acc[1]

    def commit():
        for f in targets:
            trs[f]["state"] = trs[f]["new_state"]; trs[f]["energy"] = trs[f]["new_energy"]

    def states():
        s = np.full(nf, 255, np.uint8)
        for f in targets:
            s[f] = trs[f]["state"]
        return s

    cur = (F(p["idepth_max"]) + F(p["idepth_min"])) * F(0.5)
    lastE, lastH, lastb = sweep(1000, cur)
    commit()
    if not np.isfinite(lastE) or lastH < 100:
        return ACT_SKIP, cur, states()
    lam = F(0.1)
    for _ in range(3):
        H = lastH * (F(1) + lam)
        step = F((1.0 / np.float64(H)) * np.float64(lastb))
        new = cur - step
        newE, newH, newb = sweep(1, new)
        if not np.isfinite(lastE) or newH < 100:
            return ACT_SKIP, cur, states()
        if newE < lastE:
            cur, lastH, lastb, lastE = new, newH, newb, newE
            commit()
            lam = F(np.float64(lam) * 0.5)
        else:
            lam = lam * F(5)
        if np.float64(abs(step)) < 0.0001 * np.float64(cur):
            break
    st = states()
    if not np.isfinite(cur) or int(np.sum(st == 0)) < min_obs or not np.isfinite(p["energy_th"]):
        return ACT_DELETE, cur, st
    return ACT_ACTIVATED, cur, st


# ---- pre-pyramid image path ---------------------------------------------------------
def undistort_ref(raw, remapX, remapY, G=None, vignette_inv=None, factor=1.0):
    """PhotometricUndistorter::processFrame (util/Undistort.cpp:194-227) + Undistort::undistort (:361-458), float32."""
    h_org, w_org = raw.shape
    if G is None:
        data = (F(factor) * raw.astype(F)).astype(F)
    else:
        data = np.asarray(G, F)[raw.astype(np.int64)]
        if vignette_inv is not None:
            data = (data * np.asarray(vignette_inv, F).reshape(h_org, w_org)).astype(F)
    if remapX is None:
        return data
    xx = np.asarray(remapX, F); yy = np.asarray(remapY, F)
    ok = xx >= 0
    xs = np.where(ok, xx, F(0)); ys = np.where(ok, yy, F(0))
    xi = xs.astype(np.int32); yi = ys.astype(np.int32)
    fx = xs - xi.astype(F); fy = ys - yi.astype(F)
    fxy = fx * fy
    out = fxy * data[yi + 1, xi + 1] + (fy - fxy) * data[yi + 1, xi] + (fx - fxy) * data[yi, xi + 1] + (F(1) - fx - fy + fxy) * data[yi, xi]
    return np.where(ok, out, F(0)).astype(F)


# ---- loop-closure direct alignment ------------------------------------------------
def loop_calc_res_ref(dI, K_lvl, xyz, color_lvl, refToNew, affLL, cutoff, lvl, huber=F(9)):
    """PoseEstimator::calcRes (LoopClosure/PoseEstimator.cpp:147-284), vectorised float32.  dI (h, w, 3) level image of
    the new frame, K_lvl = fx fy cx cy of the level.  -> out6 (float64), counts, warped buffers (dict, un-padded)."""
    h, w = dI.shape[:2]
    fx, fy, cx, cy = [F(x) for x in K_lvl]
    R = np.asarray(refToNew, np.float64)[:3, :3].astype(F); t = np.asarray(refToNew, np.float64)[:3, 3].astype(F)
    x, y, z = [np.asarray(xyz, np.float64)[:, k].astype(F) for k in range(3)]
    pt = [(R[k, 0] * x + R[k, 1] * y) + R[k, 2] * z + t[k] for k in range(3)]
    with np.errstate(all="ignore"):
        u, v = pt[0] / pt[2], pt[1] / pt[2]
        new_id = F(1) / pt[2]
    Ku, Kv = fx * u + cx, fy * v + cy
    sT = sRT = sN = F(0)
    if lvl == 0:
        for i in range(0, len(x), 32):
            Ku0, Kv0 = fx * (x[i] / z[i]) + cx, fy * (y[i] / z[i]) + cy
            pT = (x[i] + t[0], y[i] + t[1], F(1) + t[2])
            KuT, KvT = fx * (pT[0] / pT[2]) + cx, fy * (pT[1] / pT[2]) + cy
            pT2 = (x[i] - t[0], y[i] - t[1], F(1) - t[2])
            KuT2, KvT2 = fx * (pT2[0] / pT2[2]) + cx, fy * (pT2[1] / pT2[2]) + cy
            rp = [(R[k, 0] * x[i] + R[k, 1] * y[i]) + R[k, 2] * F(1) for k in range(3)]
            p3 = (rp[0] - t[0], rp[1] - t[1], rp[2] - t[2])
            Ku3, Kv3 = fx * (p3[0] / p3[2]) + cx, fy * (p3[1] / p3[2]) + cy
            sT += (KuT - Ku0) * (KuT - Ku0) + (KvT - Kv0) * (KvT - Kv0)
            sT += (KuT2 - Ku0) * (KuT2 - Ku0) + (KvT2 - Kv0) * (KvT2 - Kv0)
            sRT += (Ku[i] - Ku0) * (Ku[i] - Ku0) + (Kv[i] - Kv0) * (Kv[i] - Kv0)
            sRT += (Ku3 - Ku0) * (Ku3 - Ku0) + (Kv3 - Kv0) * (Kv3 - Kv0)
            sN += F(2)
    ok = (Ku > 2) & (Kv > 2) & (Ku < w - 3) & (Kv < h - 3) & (new_id > 0)
    Kus, Kvs = np.where(ok, Ku, F(3)), np.where(ok, Kv, F(3))
    hit = _bilin(dI, Kus, Kvs)
    ok &= np.isfinite(hit[:, 0])
    ref = np.asarray(color_lvl, F)
    res = hit[:, 0] - (F(affLL[0]) * ref + F(affLL[1])).astype(F)
    ares = np.abs(res)
    hw = np.where(ares < huber, F(1), huber / np.maximum(ares, F(1e-30))).astype(F)
    sat = ok & (ares > F(cutoff))
    inw = ok & ~sat
    maxE = F(2) * huber * F(cutoff) - huber * huber
    terms = np.where(sat, maxE, hw * res * res * (F(2) - hw)).astype(F)
    E = F(0)
    for k in np.flatnonzero(ok):        # sequential float sum, the order of the reference loop
        E += terms[k]
    nE, nW, nS = int(ok.sum()), int(inw.sum()), int(sat.sum())
    out6 = np.array([E, nE, sT / (sN + F(0.1)) if True else 0, 0, sRT / (sN + F(0.1)), F(nS) / F(max(nE, 1))], np.float64)
    buf = dict(idepth=new_id[inw], u=u[inw], v=v[inw], dx=hit[inw, 1], dy=hit[inw, 2], residual=res[inw], weight=hw[inw], ref=ref[inw])
    return out6, np.array([nE, nW, nS], np.int32), buf


def pose_gs_ref(buf, fx, fy, a, b0):
    """PoseEstimator::calcGSSSE (:75-145) in float64 from the warped buffers (n = count padded to a multiple of 4)."""
    dx = buf["dx"].astype(np.float64) * fx; dy = buf["dy"].astype(np.float64) * fy
    u, v, idp = [buf[k].astype(np.float64) for k in ("u", "v", "idepth")]
    J = np.stack([idp * dx, idp * dy, -idp * (u * dx + v * dy), -((u * v) * dx + dy * (1 + v * v)), (u * v) * dy + dx * (1 + u * u), u * dy - v * dx,
                  a * (b0 - buf["ref"].astype(np.float64)), -np.ones_like(u), buf["residual"].astype(np.float64)], 1)
    A = (J * buf["weight"].astype(np.float64)[:, None]).T @ J
    n = (len(u) + 3) // 4 * 4
    sc = np.array([1, 1, 1, 0.5, 0.5, 0.5, 10, 1000], np.float64)
    H = A[:8, :8] / n * sc[None, :] * sc[:, None]
    b = A[:8, 8] / n * sc
    return H, b


# ---- pixel selection ----------------------------------------------------------------
# Independent formulation of PixelSelector (FullSystem/PixelSelector2.cpp): makeHists with numpy block reductions, and
# select() NOT as the reference's nested scan but through what that scan computes -- per pot-block the first arg-max of
# |g . dir2| over pixels above the level-0 threshold; per 2pot-block a level-1 pick only if none of its pot-blocks picked;
# per 4pot-block a level-2 pick only if nothing below it picked or updated.  Only the running count n2 is sequential.
SEL_DIRS = np.array([[0, 1.0], [0.3827, 0.9239], [0.1951, 0.9808], [0.9239, 0.3827], [0.7071, 0.7071], [0.3827, -0.9239], [0.8315, 0.5556],
                     [0.8315, -0.5556], [0.5556, -0.8315], [0.9808, 0.1951], [0.9239, -0.3827], [0.7071, -0.7071], [0.5556, 0.8315],
                     [0.9808, -0.1951], [1.0, 0.0], [0.1951, -0.9808]], np.float32)


def sel_hists_ref(absg0):
    h, w = absg0.shape
    w32, h32 = w // 32, h // 32
    ths = np.zeros((h32, w32), F)
    g = np.minimum(np.sqrt(absg0.astype(F)).astype(np.int32), 48)
    valid = np.zeros((h, w), bool)
    valid[1:h - 1, 1:w - 1] = True
    for by in range(h32):
        for bx in range(w32):
            sl = (slice(32 * by, 32 * by + 32), slice(32 * bx, 32 * bx + 32))
            vals = g[sl][valid[sl]]
            hist = np.bincount(vals, minlength=49)
            th = int(F(len(vals)) * F(0.5) + F(0.5))
            q = 90
            for i in range(49):
                th -= hist[i]
                if th < 0:
                    q = i
                    break
            ths[by, bx] = q + 7
    pad = np.pad(ths, 1)
    cnt = np.pad(np.ones_like(ths), 1)
    ssum = sum(pad[1 + dy:1 + dy + h32, 1 + dx:1 + dx + w32] for dy in (-1, 0, 1) for dx in (-1, 0, 1))
    snum = sum(cnt[1 + dy:1 + dy + h32, 1 + dx:1 + dx + w32] for dy in (-1, 0, 1) for dx in (-1, 0, 1))
    sm = (ssum.astype(F) / snum.astype(F))
    return ths, (sm * sm).astype(F)


def sel_select_ref(dI0, absg, ths_smoothed, random_pattern, pot, th_factor=1.0):
    """-> (map (h, w) float32, (n2, n3, n4)).  dI0 (h, w, 3); absg = [level0, level1, level2] absSquaredGrad images."""
    h, w = dI0.shape[:2]
    w32 = w // 32
    flat = np.concatenate([ths_smoothed.reshape(-1), np.zeros(100, F)])
    ys, xs = np.mgrid[0:h, 0:w]
    th0 = flat[(xs >> 5) + (ys >> 5) * w32].astype(F)
    th1 = (th0 * F(0.75)).astype(F)
    th2 = (th1 * (F(0.75) * F(0.75))).astype(F)
    tf = F(th_factor)
    inner = (xs >= 4) & (xs < w - 5) & (ys >= 4) & (ys <= h - 4)
    ag1 = absg[1][(ys * F(0.5) + F(0.25)).astype(np.int32), (xs * F(0.5) + F(0.25)).astype(np.int32)]
    ag2 = absg[2][(ys.astype(F) * F(0.25) + 0.125).astype(np.int32), (xs.astype(F) * F(0.25) + 0.125).astype(np.int32)]
    c0 = inner & (absg[0] > th0 * tf)
    c1 = inner & (ag1 > th1 * tf)
    c2 = inner & (ag2 > th2 * tf)
    gx, gy = dI0[..., 1].astype(F), dI0[..., 2].astype(F)
    out = np.zeros((h, w), F)
    n2 = n3 = n4 = 0

    def pick(y0, x0, size, cand, d):
        sl = (slice(y0, min(y0 + size, h)), slice(x0, min(x0 + size, w)))
        m = cand[sl]
        if not m.any():
            return None
        val = np.abs(gx[sl] * SEL_DIRS[d, 0] + gy[sl] * SEL_DIRS[d, 1]).astype(F)
        val = np.where(m, val, F(-1))
        return sl, val

    for y4 in range(0, h, 4 * pot):
        for x4 in range(0, w, 4 * pot):
            d4 = random_pattern[n2] & 0xF
            trig4 = False
            order4 = []          # (scan position, value, idx) of the level-2 candidates in the reference's visiting order
            for y3 in range(y4, min(y4 + 4 * pot, h), 2 * pot):
                for x3 in range(x4, min(x4 + 4 * pot, w), 2 * pot):
                    d3 = random_pattern[n2] & 0xF
                    trig3 = False
                    best3 = (F(0), -1)
                    for y2 in range(y3, min(y3 + 2 * pot, h), pot):
                        for x2 in range(x3, min(x3 + 2 * pot, w), pot):
                            d2 = random_pattern[n2] & 0xF
                            r = pick(y2, x2, pot, c0, d2)
                            sel2 = False
                            if r is not None:
                                sl, val = r
                                if val.max() > 0:
                                    k = int(np.argmax(val))            # first maximum in row-major order
                                    yy, xx = np.unravel_index(k, val.shape)
                                    out[sl[0].start + yy, sl[1].start + xx] = 1
                                    n2 += 1
                                    sel2 = True
                            if sel2:
                                trig3 = True
                                trig4 = True
                                continue
                            if trig3:
                                continue
                            r = pick(y2, x2, pot, c1, d3)
                            if r is not None:
                                sl, val = r
                                if val.max() > best3[0]:
                                    k = int(np.argmax(val))
                                    yy, xx = np.unravel_index(k, val.shape)
                                    best3 = (val.max(), (sl[0].start + yy, sl[1].start + xx))
                                    trig4 = True
                            if not trig4:
                                r = pick(y2, x2, pot, c2, d4)
                                if r is not None:
                                    sl, val = r
                                    order4.append((sl, val))
                    if not trig3 and best3[1] != -1:
                        out[best3[1]] = 2
                        n3 += 1
            if not trig4:
                best = (F(0), None)
                for sl, val in order4:
                    if val.max() > best[0]:
                        k = int(np.argmax(val))
                        yy, xx = np.unravel_index(k, val.shape)
                        best = (val.max(), (sl[0].start + yy, sl[1].start + xx))
                if best[1] is not None:
                    out[best[1]] = 4
                    n4 += 1
    return out, (n2, n3, n4)


def sel_make_maps_ref(dI0, absg, random_pattern, potential, density, recursions_left=1, th_factor=1.0):
    """PixelSelector::makeMaps (PixelSelector2.cpp:146-282) -> (map, numHaveSub, currentPotential)."""
    _, ths_s = sel_hists_ref(absg[0])
    while True:
        m, n = sel_select_ref(dI0, absg, ths_s, random_pattern, potential, th_factor)
        have = F(n[0] + n[1] + n[2])
        with np.errstate(all="ignore"):
            quotia = F(density) / have
            K = have * F(potential + 1) * F(potential + 1)
            ideal = int(np.sqrt(K / F(density)) - F(1))
        ideal = max(ideal, 1)
        if recursions_left > 0 and quotia > 1.25 and potential > 1:
            potential = potential - 1 if ideal >= potential else ideal
            recursions_left -= 1
            continue
        if recursions_left > 0 and quotia < 0.25:
            potential = potential + 1 if ideal <= potential else ideal
            recursions_left -= 1
            continue
        break
    sub = int(have)
    if quotia < 0.95:
        char_th = int(F(255) * quotia) & 0xFF
        flat = m.reshape(-1)
        nz = np.flatnonzero(flat)
        drop = random_pattern[:len(nz)] > char_th
        flat[nz[drop]] = 0
        sub -= int(drop.sum())
    return m, sub, ideal


# ---- CoarseInitializer::calcResAndGS (second half) -----------------------------------
def init_calc_res_and_gs_ref(dIref, dInew, K_lvl, refToNew, aff, tlog, pts, alphaW=F(150 * 150), alphaK=F(2.5 * 2.5), coupling=F(1), huber=F(9)):
    """CoarseInitializer.cpp:450-673, vectorised over the points (the pattern loop stays sequential, a per-point `alive` mask
    plays the role of `break`).  Per-point outputs in float32 with the reference's expression order; H, b in float64."""
    h, w = dInew.shape[:2]
    fx, fy, cx, cy = [F(x) for x in K_lvl]
    T = np.asarray(refToNew, np.float64)
    R = T[:3, :3].astype(F); t = T[:3, 3].astype(F)
    Ki = np.array([[F(1) / fx, 0, -cx / fx], [0, F(1) / fy, -cy / fy], [0, 0, 1]], F)
    RKi = np.zeros((3, 3), F)
    for i in range(3):
        for j in range(3):
            RKi[i, j] = (R[i, 0] * Ki[0, j] + R[i, 1] * Ki[1, j]) + R[i, 2] * Ki[2, j]
    a0, a1 = F(np.exp(np.float64(F(aff[0])))), F(aff[1])
    u0, v0, idn = [np.asarray(pts[k], F) for k in ("u", "v", "idepth_new")]
    n = len(u0)
    good_in = np.asarray(pts["isGood"]).astype(bool)
    alive = good_in.copy()
    energy = np.zeros(n, F); maxstep = np.full(n, F(1e10), F)
    Jb = np.zeros((n, 10), F)
    Jall = np.zeros((n, 8, 9), F)
    with np.errstate(all="ignore"):
        for idx, (px, py) in enumerate(PATTERN):
            x = u0 + F(px); y = v0 + F(py)
            pt = [((RKi[k, 0] * x + RKi[k, 1] * y) + RKi[k, 2]) + t[k] * idn for k in range(3)]
            uu, vv = pt[0] / pt[2], pt[1] / pt[2]
            Ku, Kv = fx * uu + cx, fy * vv + cy
            nid = idn / pt[2]
            ok = (Ku > 1) & (Kv > 1) & (Ku < w - 2) & (Kv < h - 2) & (nid > 0)
            hit = _bilin(dInew, np.where(ok, Ku, F(2)), np.where(ok, Kv, F(2)))
            rl = _bilin(dIref, x, y)[:, 0]
            ok &= np.isfinite(rl) & np.isfinite(hit[:, 0])
            alive &= ok
            res = hit[:, 0] - a0 * rl - a1
            hw = np.where(np.abs(res) < huber, F(1), huber / np.abs(res)).astype(F)
            energy = np.where(alive, energy + hw * res * res * (F(2) - hw), energy).astype(F)
            dxdd = (t[0] - t[2] * uu) / pt[2]; dydd = (t[1] - t[2] * vv) / pt[2]
            hw = np.where(hw < 1, np.sqrt(hw), hw).astype(F)
            dxi = hw * hit[:, 1] * fx; dyi = hw * hit[:, 2] * fy
            J = np.stack([nid * dxi, nid * dyi, -nid * (uu * dxi + vv * dyi), -uu * vv * dxi - (F(1) + vv * vv) * dyi,
                          (F(1) + uu * uu) * dxi + uu * vv * dyi, -vv * dxi + uu * dyi, -hw * a0 * rl, -hw * F(1), hw * res], 1).astype(F)
            dd = (dxi * dxdd + dyi * dydd).astype(F)
            mx = dxdd * fx; my = dydd * fy
            ms = F(1) / np.sqrt(mx * mx + my * my)
            maxstep = np.where(alive & (ms < maxstep), ms, maxstep).astype(F)
            upd = np.concatenate([J[:, :8] * dd[:, None], (J[:, 8] * dd)[:, None], (dd * dd)[:, None]], 1).astype(F)
            Jb = np.where(alive[:, None], Jb + upd, Jb).astype(F)
            Jall[:, idx] = np.where(alive[:, None], J, 0)
    en_in = np.asarray(pts["energy"], F).reshape(n, 2)
    good = alive & ~(energy > np.asarray(pts["outlierTH"], F) * F(20))
    energy_new = en_in.copy()
    energy_new[good, 0] = energy[good]
    energy_new[good, 1] = ((idn - F(1)) * (idn - F(1)))[good]
    E = float(np.sum(np.where(good, energy, en_in[:, 0]).astype(np.float64)))
    Jg = Jall[good].astype(np.float64).reshape(-1, 9)
    A = Jg.T @ Jg
    tsq = float(T[0, 3] ** 2 + T[1, 3] ** 2 + T[2, 3] ** 2)
    alphaEnergy = F(alphaW * (0.0 + tsq * n))
    if alphaEnergy > alphaK * n:
        alphaOpt, alphaEnergy = F(0), F(alphaK * n)
    else:
        alphaOpt = F(alphaW)
    lastH = np.asarray(pts.get("lastHessian_new", np.zeros(n)), F).copy()
    lastH[good] = Jb[good, 9]
    Jb2 = Jb.copy()
    iR = np.asarray(pts["iR"], F)
    Jb2[:, 8] = Jb2[:, 8] + alphaOpt * (idn - F(1))
    Jb2[:, 9] = Jb2[:, 9] + alphaOpt
    if alphaOpt == 0:
        Jb2[:, 8] = Jb2[:, 8] + coupling * (idn - iR)
        Jb2[:, 9] = Jb2[:, 9] + coupling
    Jb2[:, 9] = F(1) / (F(1) + Jb2[:, 9])
    Jb_out = np.where(good[:, None], Jb2, Jb).astype(F)
    Jb_out[~good_in] = np.asarray(pts.get("JbBuffer_new", np.zeros((n, 10))), F)[~good_in]
    Js = Jb2[good].astype(np.float64)
    Asc = (Js[:, :9] * Js[:, 9:10]).T @ Js[:, :9]
    H = A[:8, :8].copy(); b = A[:8, 8].copy()
    for k in range(3):
        H[k, k] += float(alphaOpt) * n
        b[k] += float(tlog[k]) * float(alphaOpt) * n
    return dict(H=H, b=b, Hsc=Asc[:8, :8], bsc=Asc[:8, 8], res3=np.array([E, alphaEnergy, 2 * n]), energy_new=energy_new, isGood_new=good.astype(np.uint8),
                maxstep=maxstep, lastHessian_new=lastH, JbBuffer_new=Jb_out)


# ---- coarse tracker / scale optimizer (rows a14, a15, a17) --------------------------------------------------
def align_calc_res_ref(dI, K_lvl, pc, R, t, lvl, cutoff, kind="pose", affLL=(1.0, 0.0), scale=1.0, Ki_lvl=None, huber=F(9)):
    """CoarseTracker::calcResPose (CoarseTracker.cpp:612-764) and ScaleOptimizer::calcResScale (ScaleOptimizer.cpp:273-437),
    vectorised float32.  dI (h, w, 3) level image of the new frame; K_lvl = fx fy cx cy of the camera the points are
    projected INTO; Ki_lvl = fx fy cx cy of the reference camera (defaults to K_lvl); pc = (u, v, idepth, color).
    -> out6 (float64), counts, warped buffers (dict, un-padded; rx* instead of idepth/u/v for kind == "scale")."""
    h, w = dI.shape[:2]
    fx, fy, cx, cy = [F(x) for x in K_lvl]
    kfx, kfy, kcx, kcy = [F(x) for x in (K_lvl if Ki_lvl is None else Ki_lvl)]
    Ki = np.array([[F(1) / kfx, 0, -kcx / kfx], [0, F(1) / kfy, -kcy / kfy], [0, 0, 1]], F)
    R = np.asarray(R, np.float64).astype(F); t = np.asarray(t, np.float64).astype(F)
    RKi = np.zeros((3, 3), F)
    for i in range(3):
        for j in range(3):
            RKi[i, j] = (R[i, 0] * Ki[0, j] + R[i, 1] * Ki[1, j]) + R[i, 2] * Ki[2, j]
    s = F(scale)
    M = (s * RKi).astype(F) if kind == "scale" else RKi
    Mk = (s * Ki).astype(F) if kind == "scale" else Ki
    x, y, idp, col = [np.asarray(a, F) for a in pc]

    def proj(A, sign):
        p = [((A[k, 0] * x + A[k, 1] * y) + A[k, 2]) + F(sign) * t[k] * idp if sign >= 0 else ((A[k, 0] * x + A[k, 1] * y) + A[k, 2]) - t[k] * idp
             for k in range(3)]
        return p

    with np.errstate(all="ignore"):
        pt = proj(M, 1)
        u, v = pt[0] / pt[2], pt[1] / pt[2]
        Ku, Kv = fx * u + cx, fy * v + cy
        nid = idp / pt[2]
        rxs = [(((RKi[k, 0] * x + RKi[k, 1] * y) + RKi[k, 2]) / idp).astype(F) for k in range(3)]
        sT = sRT = sN = F(0)
        if lvl == 0:
            pT, pT2, p3 = proj(Mk, 1), proj(Mk, -1), proj(M, -1)
            KuT, KvT = fx * (pT[0] / pT[2]) + cx, fy * (pT[1] / pT[2]) + cy
            KuT2, KvT2 = fx * (pT2[0] / pT2[2]) + cx, fy * (pT2[1] / pT2[2]) + cy
            Ku3, Kv3 = fx * (p3[0] / p3[2]) + cx, fy * (p3[1] / p3[2]) + cy
            for i in range(0, len(x), 32):
                sT += (KuT[i] - x[i]) * (KuT[i] - x[i]) + (KvT[i] - y[i]) * (KvT[i] - y[i])
                sT += (KuT2[i] - x[i]) * (KuT2[i] - x[i]) + (KvT2[i] - y[i]) * (KvT2[i] - y[i])
                sRT += (Ku[i] - x[i]) * (Ku[i] - x[i]) + (Kv[i] - y[i]) * (Kv[i] - y[i])
                sRT += (Ku3[i] - x[i]) * (Ku3[i] - x[i]) + (Kv3[i] - y[i]) * (Kv3[i] - y[i])
                sN += F(2)
    ok = (Ku > 2) & (Kv > 2) & (Ku < w - 3) & (Kv < h - 3) & (nid > 0)
    hit = _bilin(dI, np.where(ok, Ku, F(3)), np.where(ok, Kv, F(3)))
    ok &= np.isfinite(hit[:, 0])
    res = hit[:, 0] - ((F(affLL[0]) * col + F(affLL[1])).astype(F) if kind == "pose" else col)
    ares = np.abs(res)
    hw = np.where(ares < huber, F(1), huber / np.maximum(ares, F(1e-30))).astype(F)
    sat = ok & (ares > F(cutoff))
    inw = ok & ~sat
    maxE = F(2) * huber * F(cutoff) - huber * huber
    terms = np.where(sat, maxE, hw * res * res * (F(2) - hw)).astype(F)
    E = F(0)
    for k in np.flatnonzero(ok):
        E += terms[k]
    nE, nW, nS = int(ok.sum()), int(inw.sum()), int(sat.sum())
    out6 = np.array([E, nE, sT / (sN + F(0.1)), 0, sRT / (sN + F(0.1)), F(nS) / F(max(nE, 1))], np.float64)
    buf = dict(dx=hit[inw, 1], dy=hit[inw, 2], residual=res[inw], weight=hw[inw], ref=col[inw])
    if kind == "pose":
        buf.update(idepth=nid[inw], u=u[inw], v=v[inw])
    else:
        buf.update(rx1=rxs[0][inw], rx2=rxs[1][inw], rx3=rxs[2][inw])
    return out6, np.array([nE, nW, nS], np.int32), buf


def scale_gs_ref(buf, fx1, fy1, scale, t):
    """ScaleOptimizer::calcGSSSEScale (ScaleOptimizer.cpp:232-271) in float64: H = sum w J0^2 / n', b = sum w J0 r / n'."""
    dxfx = buf["dx"].astype(np.float64) * fx1; dyfy = buf["dy"].astype(np.float64) * fy1
    rx1, rx2, rx3 = [buf[k].astype(np.float64) for k in ("rx1", "rx2", "rx3")]
    deno = 1.0 / (scale * rx3 + t[2]) ** 2
    J0 = dxfx * (deno * (rx1 * t[2] - rx3 * t[0])) + dyfy * (deno * (rx2 * t[2] - rx3 * t[1]))
    w = buf["weight"].astype(np.float64)
    n = (len(w) + 3) // 4 * 4
    return float(np.sum(w * J0 * J0) / n), float(np.sum(w * J0 * buf["residual"].astype(np.float64)) / n)


# ---- a16: makeCoarseDepthL0 / trackNewestCoarse / optimizeScale (independent numpy formulations) -----------------
def make_coarse_depth_ref(dI_levels, cpt, HdiF):
    """CoarseTracker::makeCoarseDepthL0 (CoarseTracker.cpp:56-230) with array operations: np.add.at splat, reshape-sum pooling,
    shifted-array dilation (neighbours outside the map count as empty), row-major compaction.
    dI_levels: per level (h, w, 3).  -> per level (pc_u, pc_v, pc_idepth, pc_color)."""
    L = len(dI_levels)
    h0, w0 = dI_levels[0].shape[:2]
    cpt = np.asarray(cpt, F).reshape(-1, 3)
    u = (cpt[:, 0] + F(0.5)).astype(np.int64); v = (cpt[:, 1] + F(0.5)).astype(np.int64)
    wgt = np.sqrt((1e-3 / (np.asarray(HdiF, F).astype(np.float64) + 1e-12)).astype(F))      # sqrtf of the double quotient cast to float
    idp = [np.zeros((h0, w0), F)]; ws = [np.zeros((h0, w0), F)]
    for k in range(len(u)):           # float += in point order (collisions are order-sensitive)
        idp[0][v[k], u[k]] += cpt[k, 2] * wgt[k]
        ws[0][v[k], u[k]] += wgt[k]
    for l in range(1, L):
        hl, wl = dI_levels[l].shape[:2]
        def pool(a):
            a = a[:2 * hl, :2 * wl]
            return ((a[0::2, 0::2] + a[0::2, 1::2]) + a[1::2, 0::2]) + a[1::2, 1::2]
        idp.append(pool(idp[l - 1]).astype(F)); ws.append(pool(ws[l - 1]).astype(F))
    out = []
    for l in range(L):
        hl, wl = dI_levels[l].shape[:2]
        bak = ws[l].copy()
        offs = [(1, 1), (-1, -1), (1, -1), (-1, 1)] if l < 2 else [(0, 1), (0, -1), (1, 0), (-1, 0)]     # (dy, dx) in the reference's order
        flat_b, flat_i = bak.ravel(), idp[l].ravel().copy()
        N = hl * wl
        idx = np.arange(wl, N - wl)
        s = np.zeros(len(idx), F); num = np.zeros(len(idx), F); numn = np.zeros(len(idx), F)
        for dy, dx in offs:
            j = idx + dy * wl + dx      # FLAT neighbour index like the reference (wraps across rows at the borders)
            okj = (j >= 0) & (j < N)
            jj = np.where(okj, j, 0)
            m = okj & (flat_b[jj] > 0)
            s = np.where(m, s + flat_i[jj], s).astype(F); num = np.where(m, num + flat_b[jj], num).astype(F); numn = np.where(m, numn + F(1), numn).astype(F)
        fill = (flat_b[idx] <= 0) & (numn > 0)
        new_i = idp[l].ravel().copy(); new_w = ws[l].ravel().copy()
        with np.errstate(all="ignore"):
            new_i[idx[fill]] = (s[fill] / numn[fill]).astype(F); new_w[idx[fill]] = (num[fill] / numn[fill]).astype(F)
        I = new_i.reshape(hl, wl); W = new_w.reshape(hl, wl)
        ys, xs = np.mgrid[2:hl - 2, 2:wl - 2]
        Wc = W[2:hl - 2, 2:wl - 2]; Ic = I[2:hl - 2, 2:wl - 2]
        with np.errstate(all="ignore"):
            nid = np.where(Wc > 0, Ic / np.where(Wc > 0, Wc, F(1)), F(-1)).astype(F)
        col = dI_levels[l][2:hl - 2, 2:wl - 2, 0]
        keep = (Wc > 0) & np.isfinite(col) & (nid > 0)
        out.append((xs[keep].astype(F), ys[keep].astype(F), nid[keep], col[keep].astype(F)))
    return out


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


def track_newest_coarse_ref(dI_new_levels, K, pcs, T0, aff0, ref_exp, new_exp, ref_aff, coarsest, min_res_abort=None, cutoff_th=20.0, huber=F(9)):
    """CoarseTracker::trackNewestCoarse (CoarseTracker.cpp:366-552), affine modes 0/0, poses as 4x4 matrices, numpy solves.
    -> dict(ok, T, aff, last_residuals, flow, passes=[(lvl, iterations, accept_mask, residual)])."""
    maxIt = [10, 20, 50, 50, 50]
    T = np.array(T0, np.float64); aff = np.array(aff0, np.float64)
    last = np.full(5, np.nan); flow = np.full(3, 1000.0)
    mra = np.full(5, np.nan) if min_res_abort is None else np.asarray(min_res_abort, np.float64)
    passes = []
    haveRepeated = False

    def lvlK(l):
        return np.array([K[0] / (1 << l), K[1] / (1 << l), (K[2] + 0.5) / (1 << l) - 0.5, (K[3] + 0.5) / (1 << l) - 0.5], np.float32)

    def affLL(a):
        aa, bb = aff_from_to(ref_exp, new_exp, ref_aff[0], ref_aff[1], a[0], a[1])
        return aa, bb

    def calc_res(l, Tm, a, cutoff):
        aa, bb = affLL(a)
        return align_calc_res_ref(dI_new_levels[l], lvlK(l), pcs[l], Tm[:3, :3], Tm[:3, 3], l, cutoff, "pose", (F(aa), F(bb)), huber=huber)

    lvl = coarsest
    while lvl >= 0:
        rep = F(1)
        resOld, _, buf = calc_res(lvl, T, aff, F(cutoff_th) * rep)
        while resOld[5] > 0.6 and rep < 50:
            rep = rep * F(2)
            resOld, _, buf = calc_res(lvl, T, aff, F(cutoff_th) * rep)
        k = lvlK(lvl)
        H, b = pose_gs_ref(buf, float(k[0]), float(k[1]), float(F(affLL(aff)[0])), float(F(ref_aff[1])))
        lam = F(0.01)
        acc_mask, it = 0, 0
        while it < maxIt[lvl]:
            Hl = H.copy()
            Hl[np.arange(8), np.arange(8)] *= (1 + float(lam))
            inc = np.linalg.solve(Hl, -b)
            extrap = F(1)
            if lam < F(0.001):
                extrap = F(np.sqrt(np.sqrt(np.float64(F(0.001)) / np.float64(lam))))
            inc = inc * float(extrap)
            incS = inc * np.array([1, 1, 1, 0.5, 0.5, 0.5, 10, 1000.0])
            if not np.isfinite(incS.sum()):
                incS[:] = 0
            Tn = se3_exp(incS[:6]) @ T
            an = aff + incS[6:8]
            resNew, _, bufn = calc_res(lvl, Tn, an, F(cutoff_th) * rep)
            accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1])
            if accept:
                H, b = pose_gs_ref(bufn, float(k[0]), float(k[1]), float(F(affLL(an)[0])), float(F(ref_aff[1])))
                resOld, aff, T = resNew, an, Tn
                lam = lam * F(0.5)
                acc_mask |= 1 << it
            else:
                lam = lam * F(4)
                if lam < F(0.001):
                    lam = F(0.001)
            it += 1
            if not (np.linalg.norm(inc) > 1e-3):
                break
        last[lvl] = F(np.sqrt(F(resOld[0] / resOld[1])))
        flow = resOld[2:5].copy()
        passes.append((lvl, it, acc_mask, float(last[lvl])))
        if last[lvl] > 1.5 * mra[lvl]:
            return dict(ok=False, T=np.array(T0, np.float64), aff=np.array(aff0, np.float64), last_residuals=last, flow=flow, passes=passes)
        if rep > 1 and not haveRepeated:
            haveRepeated = True
        else:
            lvl -= 1
    ok = True
    aa, bb = affLL(aff)
    if abs(np.log(F(aa))) > 1.5 or abs(F(bb)) > 200:
        ok = False
    return dict(ok=ok, T=T, aff=aff, last_residuals=last, flow=flow, passes=passes)


def optimize_scale_ref(dI_stereo_levels, K0, K1, pcs, T10, scale0, coarsest, cutoff_th=20.0):
    """ScaleOptimizer::optimizeScale (ScaleOptimizer.cpp:120-230). -> dict(scale, error, last_residuals, passes)."""
    maxIt = [10, 20, 50, 50, 50]
    last = np.full(5, np.nan)
    s_cur = F(scale0)
    passes = []
    haveRepeated = False
    T10 = np.asarray(T10, np.float64)

    def lvlK(Kc, l):
        return np.array([Kc[0] / (1 << l), Kc[1] / (1 << l), (Kc[2] + 0.5) / (1 << l) - 0.5, (Kc[3] + 0.5) / (1 << l) - 0.5], np.float32)

    def calc_res(l, s, cutoff):
        return align_calc_res_ref(dI_stereo_levels[l], lvlK(K1, l), pcs[l], T10[:3, :3], T10[:3, 3], l, cutoff, "scale", scale=s, Ki_lvl=lvlK(K0, l))

    lvl = coarsest
    while lvl >= 0:
        rep = F(1)
        resOld, _, buf = calc_res(lvl, s_cur, F(cutoff_th) * rep)
        while resOld[5] > 0.6 and rep < 50:
            rep = rep * F(2)
            resOld, _, buf = calc_res(lvl, s_cur, F(cutoff_th) * rep)
        k1 = lvlK(K1, lvl)
        H, b = scale_gs_ref(buf, float(k1[0]), float(k1[1]), float(s_cur), T10[:3, 3])
        H, b = F(H), F(b)
        lam = F(0.01)
        acc_mask, it = 0, 0
        while it < maxIt[lvl]:
            Hl = F(H * (F(1) + lam))
            with np.errstate(all="ignore"):
                inc = F(-b / Hl)
            extrap = F(1)
            if lam < F(0.001):
                extrap = F(np.sqrt(np.sqrt(np.float64(F(0.001)) / np.float64(lam))))
            inc = F(inc * extrap)
            if not np.isfinite(inc) or abs(inc) > s_cur:
                inc = F(0)
            s_new = F(s_cur + inc)
            resNew, _, bufn = calc_res(lvl, s_new, F(cutoff_th) * rep)
            accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1])
            if accept:
                H, b = scale_gs_ref(bufn, float(k1[0]), float(k1[1]), float(s_new), T10[:3, 3])
                H, b = F(H), F(b)
                resOld, s_cur = resNew, s_new
                lam = lam * F(0.5)
                acc_mask |= 1 << it
            else:
                lam = lam * F(4)
                if lam < F(0.001):
                    lam = F(0.001)
            it += 1
            if not (inc > 1e-3):
                break
        last[lvl] = F(np.sqrt(F(resOld[0] / resOld[1])))
        passes.append((lvl, it, acc_mask))
        if rep > 1 and not haveRepeated:
            haveRepeated = True
        else:
            lvl -= 1
    return dict(scale=float(s_cur), error=float(last[0]), last_residuals=last, passes=passes)


def distance_map_ref(w1, h1, KRKi, Kt, host, u, v, idepth):
    """CoarseDistanceMap::makeDistanceMap + growDistBFS (CoarseTracker.cpp:789-916) as a level-synchronous PULL: at step k every
    pixel still above k takes k if one of its (8 for odd k, 4 for even k) neighbours holds k - 1 and is not a border pixel."""
    KRKi = np.asarray(KRKi, F).reshape(-1, 9); Kt = np.asarray(Kt, F).reshape(-1, 3)
    M, t = KRKi[host], Kt[host]
    u, v, idp = np.asarray(u, F), np.asarray(v, F), np.asarray(idepth, F)
    p = [((M[:, 3 * k] * u + M[:, 3 * k + 1] * v) + M[:, 3 * k + 2] * F(1)) + t[:, k] * idp for k in range(3)]
    with np.errstate(all="ignore"):
        fu, fv = p[0] / p[2] + F(0.5), p[1] / p[2] + F(0.5)
    okf = np.isfinite(fu) & np.isfinite(fv) & (np.abs(fu) < 1e9) & (np.abs(fv) < 1e9)
    uu = np.where(okf, fu, -1).astype(np.int64); vv = np.where(okf, fv, -1).astype(np.int64)      # C cast: truncation toward zero
    ok = okf & (uu > 0) & (vv > 0) & (uu < w1) & (vv < h1)
    d = np.full((h1, w1), 1000, np.float32)
    d[vv[ok], uu[ok]] = 0
    border = np.zeros((h1, w1), bool)
    border[0, :] = border[-1, :] = border[:, 0] = border[:, -1] = True
    for k in range(1, 40):
        src = (d == k - 1) & ~border
        nb = [(0, 1), (0, -1), (1, 0), (-1, 0)] + ([(1, 1), (1, -1), (-1, -1), (-1, 1)] if k % 2 else [])
        reach = np.zeros_like(src)
        for dy, dx in nb:
            sh = np.zeros_like(src)
            ys, yd = (slice(0, h1 - dy), slice(dy, h1)) if dy >= 0 else (slice(-dy, h1), slice(0, h1 + dy))
            xs, xd = (slice(0, w1 - dx), slice(dx, w1)) if dx >= 0 else (slice(-dx, w1), slice(0, w1 + dx))
            sh[yd, xd] = src[ys, xs]
            reach |= sh
        d[reach & (d > k)] = k
    return d


# ---- the solver of k_solve.cu, restated ---------------------------------------------------------------------
def _adjugate4(A):
    """adjugate and determinant of a symmetric 4x4 block from its 2x2 minors, the formula sheet of panel_rows<true> / panel_pipeline"""
    a00, a10, a11, a20, a21, a22, a30, a31, a32, a33 = A[0, 0], A[1, 0], A[1, 1], A[2, 0], A[2, 1], A[2, 2], A[3, 0], A[3, 1], A[3, 2], A[3, 3]
    s0 = a00 * a11 - a10 * a10; s1 = a00 * a21 - a10 * a20; s2 = a00 * a31 - a10 * a30
    s3 = a10 * a21 - a11 * a20; s4 = a10 * a31 - a11 * a30; s5 = a20 * a31 - a21 * a30
    c5 = a22 * a33 - a32 * a32; c4 = a21 * a33 - a31 * a32; c3 = a21 * a32 - a31 * a22
    c2 = a20 * a33 - a30 * a32; c1 = a20 * a32 - a30 * a22
    det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * s5
    i00 = a11 * c5 - a21 * c4 + a31 * c3; i01 = -a10 * c5 + a20 * c4 - a30 * c3
    i02 = a31 * s5 - a32 * s4 + a33 * s3; i03 = -a21 * s5 + a22 * s4 - a32 * s3
    i11 = a00 * c5 - a20 * c2 + a30 * c1; i12 = -a30 * s5 + a32 * s2 - a33 * s1
    i13 = a20 * s5 - a22 * s2 + a32 * s1; i22 = a30 * s4 - a31 * s2 + a33 * s0
    i23 = -a20 * s4 + a21 * s2 - a32 * s0; i33 = a20 * s3 - a21 * s1 + a22 * s0
    adj = np.array([[i00, i01, i02, i03], [i01, i11, i12, i13], [i02, i12, i22, i23], [i03, i13, i23, i33]])
    return adj, det


def block_ldlt_solve(Hf, bf):
    """The solve of k_solve.cu on an assembled system (EnergyFunctional.cpp:1141-1149 solves the same one with Eigen::LDLT):
    Jacobi scaling 1/sqrt(diag + 10); pivot order = descending |scaled diagonal| (ties by index); right-looking block LDL^T
    with 4x4 pivot blocks inverted through their adjugate (a singular block leaves its columns alone); the right-hand side
    rides along as an extra row and ends as A11^-1 (L^-1 P b) per block; back substitution block by block; x = S * P^T w."""
    D = len(bf)
    assert D % 4 == 0
    Hf = np.tril(Hf) + np.tril(Hf, -1).T      # Eigen::LDLT and the kernel's assembly both read entry (max(i,j), min(i,j))
    S = 1.0 / np.sqrt(np.diag(Hf) + 10)
    A = S[:, None] * Hf * S[None, :]
    key = np.abs(np.diag(A))
    perm = np.array(sorted(range(D), key=lambda i: (-key[i], i)))
    M = np.zeros((D + 1, D))
    M[:D] = A[np.ix_(perm, perm)]
    M[D] = (S * bf)[perm]
    W = np.zeros((D + 1, D))
    for k0 in range(0, D, 4):
        adj, det = _adjugate4(M[k0:k0 + 4, k0:k0 + 4])
        rdet = 1.0 / det if det != 0 and np.isfinite(1.0 / det) else 0.0
        rows = np.arange(k0 + 4, D + 1)
        A21 = M[rows, k0:k0 + 4].copy()
        Wr = (A21 @ adj) * rdet
        W[rows, k0:k0 + 4] = Wr
        cols = np.arange(k0 + 4, D)
        M[np.ix_(rows, cols)] -= Wr @ M[cols, k0:k0 + 4].T
    w = W[D].copy()
    for k0 in range(D - 4, -1, -4):
        w[:k0] -= W[k0:k0 + 4, :k0].T @ w[k0:k0 + 4]     # rows of the block carry their coupling to every earlier column
    x = np.zeros(D)
    x[perm] = w
    return S * x
