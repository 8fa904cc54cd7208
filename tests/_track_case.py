"""Shared scenario of the direct-alignment control-loop tests: the window is optimised, the IN
residuals on the newest keyframe give makeCoarseDepthL0 its input, the previous keyframe's image plays the newly arrived frame
(its true relative pose is known), a shifted copy of the newest keyframe plays camera 1 of the stereo pair."""
import numpy as np

from _scenes import open_handle, upload
from sos_slam_b200 import problem, synth


def rot_to_quat(R):
    """(x, y, z, w) of a rotation matrix (Shepperd)."""
    R = np.asarray(R, np.float64)
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        w, x, y, z = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(3)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        w = (R[k, j] - R[j, k]) / s
        x, y, z = q
    v = np.array([x, y, z, w])
    return v / np.linalg.norm(v)


def quat_to_T(q, t):
    x, y, z, w = q
    T = np.eye(4)
    T[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    T[:3, 3] = t
    return T


def coarse_depth_input(h, sc, iters=3):
    """optimize the window on handle `h`; -> (centerProjectedTo [n,3], HdiF [n]) of the IN residuals on the newest keyframe."""
    P, keep = upload(h, sc)
    h.ba_optimize(iters)
    st, aux, acc = h.get_state(), h.get_aux(), h.points_get_acc()
    sel = (sc.res_target == sc.nf - 1) & (st["state"] == 0) & (st["is_active"] == 1)
    return aux["centerProjectedTo"][sel].copy(), acc["HdiF"][sc.res_point[sel]].copy()


def hypotheses(sc, n_extra=0, seed=0):
    """pose guesses for tracking frame nf-2 against the reference nf-1: perturbed truth, identity, and random rotations."""
    Ttrue = np.linalg.inv(sc.camToWorld_true[sc.nf - 2]) @ sc.camToWorld_true[sc.nf - 1]
    rng = np.random.default_rng(seed)
    guesses = [synth.se3_exp([0.004, -0.003, 0.002, 0.002, -0.003, 0.001]) @ Ttrue, np.eye(4)]
    for _ in range(n_extra):
        guesses.append(synth.se3_exp(np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.02, 3)])) @ Ttrue)
    return Ttrue, [dict(q=rot_to_quat(T[:3, :3]), t=T[:3, 3].copy(), aff_g2l=(0.0, 0.0)) for T in guesses]


def ref_affine(sc):
    """lastRef->aff_g2l() and the exposures of reference (nf-1) and new frame (nf-2)."""
    return (float(sc.aff_true[sc.nf - 1, 0]), float(sc.aff_true[sc.nf - 1, 1])), float(sc.ab_exposure[sc.nf - 1]), float(sc.ab_exposure[sc.nf - 2])


def stereo_frame(sc, seed, baseline=0.08):
    """(tfmF0ToF1 3x4, camera-1 image of the newest keyframe) rendered from the scene's surface (synth.render_view; `seed` = the
    scene's).  The translation has a (small) z component like every calibrated rig: with t_z == 0 exactly, the zero rows that
    pad the warped buffers to a multiple of 4 give 1 / (s * 0 + t_z)^2 = inf and inf * 0 = NaN in calcGSSSEScale
    (ScaleOptimizer.cpp:249-262) -- in the reference too."""
    T10 = synth.se3_exp([-baseline, 0.001, 0.0015, 0.001, -0.002, 0.0005])
    c2w1 = sc.camToWorld_true[sc.nf - 1] @ np.linalg.inv(T10)
    return T10[:3, :4], synth.render_view(seed, sc.w, sc.h, sc.K, c2w1)


def distance_map_case(sc, seed=0):
    """inputs of CoarseDistanceMap::makeDistanceMap for the newest keyframe: every active point of the other keyframes with
    KRKi = K[1] R Ki[0], Kt = K[1] t of its host (FullSystem.cpp:417-424), plus a few points that project outside / onto the border."""
    nf = sc.nf
    fx, fy, cx, cy = [np.float32(x) for x in sc.K]
    K1 = np.array([[fx * np.float32(0.5), 0, (cx + np.float32(0.5)) / 2 - np.float32(0.5)], [0, fy * np.float32(0.5), (cy + np.float32(0.5)) / 2 - np.float32(0.5)], [0, 0, 1]], np.float32)
    Ki0 = np.array([[1 / fx, 0, -cx / fx], [0, 1 / fy, -cy / fy], [0, 0, 1]], np.float32)
    KRKi, Kt = [], []
    for hst in range(nf):
        T = np.linalg.inv(sc.camToWorld_true[nf - 1]) @ sc.camToWorld_true[hst]
        KRKi.append((K1 @ T[:3, :3].astype(np.float32) @ Ki0).astype(np.float32).ravel())
        Kt.append((K1 @ T[:3, 3].astype(np.float32)).astype(np.float32))
    m = sc.pt_host != nf - 1
    host, u, v, idp = sc.pt_host[m].astype(np.int32), sc.pt_u[m].copy(), sc.pt_v[m].copy(), sc.pt_idepth[m].copy()
    rng = np.random.default_rng(seed)
    extra = 12
    host = np.concatenate([host, np.zeros(extra, np.int32)])
    u = np.concatenate([u, rng.uniform(-200, sc.w + 200, extra).astype(np.float32)])
    v = np.concatenate([v, rng.uniform(-200, sc.h + 200, extra).astype(np.float32)])
    idp = np.concatenate([idp, np.full(extra, 0.5, np.float32)])
    # a corner region without points far from everything keeps 1000; a point exactly on the border row does not spread
    u[-1], v[-1], idp[-1] = 2.0 * 7 + 0.6, 0.4, 0.0
    return np.array(KRKi), np.array(Kt), host, u, v, idp
