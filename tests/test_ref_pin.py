"""Pins the oracle to REFERENCE TEXT: a few reference translation units whose arithmetic
is spelled out in scalar / SSE code compile unmodified against minimal Eigen / Sophus / boost stand-ins (oracle/ref_stub/,
recipe oracle/ref_build.sh -> oracle/_ref/libref_units.so):

  OptimizationBackend/MatrixAccumulators.h   AccumulatorApprox, Accumulator9, Accumulator11, AccumulatorXX, AccumulatorX
  OptimizationBackend/ScaleAccumulator.h     ScaleAccumulator
  util/globalFuncs.h                         getInterpolatedElement33 / 31 / 33BiLin / (float image)
  util/settings.cpp                          the setting_* defaults and the residual pattern

The oracle's restatements (orc_accum.h, orc_sample.h, orc_config_default, patternP) must reproduce them BIT FOR BIT on random
streams long enough to cross both shiftUp tiers (> 1000 and > 10^6 additions).  Rows a6, a8, a15, a17 and f4 of SURVEY §8
accumulate through these classes, a3 / a14 / f1 sample through these functions.  On a box with /root/reference the library is
(re)built and its absence is a failure; on the GPU box the prebuilt file travels with the snapshot."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_units.so")
ORC_SO = os.path.join(ROOT, "oracle", "_build", "liborc_parity.so")
fp = C.POINTER(C.c_float)


def _p(a):
    return a.ctypes.data_as(fp)


@pytest.fixture(scope="module")
def libs(built):
    if os.path.isdir("/root/reference"):
        r = subprocess.run(["bash", os.path.join(ROOT, "oracle", "ref_build.sh")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert os.path.exists(REF_SO), "oracle/_ref/libref_units.so missing although /root/reference is present"
    elif not os.path.exists(REF_SO):
        pytest.skip("no /root/reference and no prebuilt oracle/_ref/libref_units.so")
    ref, orc = C.CDLL(REF_SO), C.CDLL(ORC_SO)
    ref.ref_acc11_run.restype = C.c_float
    orc.orc_pin_acc11_run.restype = C.c_float
    return ref, orc


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("n", [7, 1000, 1001, 2500, 1_100_000])
def test_accumulator_approx(libs, n):
    """AccumulatorApprox::update / updateTopRight / updateBotRight / shiftUp / finish (MatrixAccumulators.h:744-1170)."""
    ref, orc = libs
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((n, 35)) * rng.uniform(0.01, 30.0, (1, 35))).astype(np.float32)
    out = []
    for fn in (ref.ref_approx_run, orc.orc_pin_approx_run):
        H = np.zeros(169, np.float32)
        num = C.c_double()
        fn(n, _p(x), _p(H), C.byref(num))
        out.append((H, num.value))
    assert out[0][1] == out[1][1] == n
    assert np.isfinite(out[0][0]).all() and np.abs(out[0][0]).max() > 0
    assert np.array_equal(_bits(out[0][0]), _bits(out[1][0]))


@pytest.mark.parametrize("mode", [0, 1, 3], ids=["updateSSE", "updateSSE_eighted", "updateSingleWeighted"])
@pytest.mark.parametrize("n", [8, 4004, 4_100_000])
def test_accumulator9(libs, mode, n):
    """Accumulator9 (MatrixAccumulators.h:1172-1687): the 9x9 normal equations of calcGSSSEPose / calcResAndGS."""
    ref, orc = libs
    if mode == 3 and n > 2_000_000:
        n = 1_100_000
    rng = np.random.default_rng(100 * mode + n % 97)
    J = (rng.standard_normal((9, n)) * rng.uniform(0.01, 20.0, (9, 1))).astype(np.float32)
    w = rng.uniform(0.0, 1.0, n).astype(np.float32)
    out = []
    for fn in (ref.ref_acc9_run, orc.orc_pin_acc9_run):
        H = np.zeros(81, np.float32)
        num = C.c_double()
        fn(mode, n, _p(J), _p(w), _p(H), C.byref(num))
        out.append((H, num.value))
    assert out[0][1] == out[1][1] == n
    assert np.array_equal(_bits(out[0][0]), _bits(out[1][0]))


@pytest.mark.parametrize("n", [5, 1001, 1_200_000])
def test_accumulator11_xx_x_scale(libs, n):
    """Accumulator11::updateSingle, AccumulatorXX<8,8> / <8,4>, AccumulatorX<8> (tier logic; the outer-product expression is
    evaluated by the stand-in Eigen) and ScaleAccumulator::updateSSE_oneed."""
    ref, orc = libs
    rng = np.random.default_rng(n)
    v = rng.uniform(0, 50, n).astype(np.float32)
    assert np.float32(ref.ref_acc11_run(n, _p(v))).view(np.uint32) == np.float32(orc.orc_pin_acc11_run(n, _p(v))).view(np.uint32)
    m = min(n, 300_000)
    L = rng.standard_normal((m, 8)).astype(np.float32)
    R8 = rng.standard_normal((m, 8)).astype(np.float32)
    R4 = rng.standard_normal((m, 4)).astype(np.float32)
    w = rng.uniform(0, 2, m).astype(np.float32)
    for name, R, k in (("accxx88", R8, 64), ("accxx84", R4, 32)):
        a, b = np.zeros(k, np.float32), np.zeros(k, np.float32)
        getattr(ref, f"ref_{name}_run")(m, _p(L), _p(R), _p(w), _p(a))
        getattr(orc, f"orc_pin_{name}_run")(m, _p(L), _p(R), _p(w), _p(b))
        assert np.array_equal(_bits(a), _bits(b)), name
    a, b = np.zeros(8, np.float32), np.zeros(8, np.float32)
    ref.ref_accx8_run(m, _p(L), _p(w), _p(a))
    orc.orc_pin_accx8_run(m, _p(L), _p(w), _p(b))
    assert np.array_equal(_bits(a), _bits(b))
    n4 = max(4, n // 4 * 4)
    J0 = rng.standard_normal(n4).astype(np.float32)
    J1 = rng.standard_normal(n4).astype(np.float32)
    ww = rng.uniform(0, 1, n4).astype(np.float32)
    out = []
    for fn in (ref.ref_scaleacc_run, orc.orc_pin_scaleacc_run):
        H = np.zeros(4, np.float32)
        num = C.c_double()
        fn(n4, _p(J0), _p(J1), _p(ww), _p(H), C.byref(num))
        out.append((H, num.value))
    assert out[0][1] == out[1][1] == n4
    assert np.array_equal(_bits(out[0][0]), _bits(out[1][0]))


@pytest.mark.parametrize("which", [0, 1, 2, 3], ids=["33", "31", "33BiLin", "float"])
def test_bilinear_samplers(libs, which):
    """util/globalFuncs.h:36-52, 68-82, 122-136, 161-182 on a random Vector3f image."""
    ref, orc = libs
    rng = np.random.default_rng(7 + which)
    w, h, n = 96, 64, 20000
    img3 = rng.uniform(0, 255, (h, w, 3)).astype(np.float32)
    img1 = np.ascontiguousarray(img3[:, :, 0])
    x = rng.uniform(0, w - 1.001, n).astype(np.float32)
    y = rng.uniform(0, h - 1.001, n).astype(np.float32)
    x[:50] = np.floor(x[:50])     # integer positions: dx = 0
    y[25:75] = np.floor(y[25:75])
    a, b = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    ref.ref_interp(which, _p(img3), _p(img1), w, n, _p(x), _p(y), _p(a))
    orc.orc_pin_interp(which, _p(img3), _p(img1), w, n, _p(x), _p(y), _p(b))
    assert np.abs(a).max() > 1
    assert np.array_equal(_bits(a), _bits(b))


def test_settings_and_pattern(libs, orc):
    """util/settings.cpp defaults after settingsDefault(preset 0, mode 1) (main.cpp:27-90 only touches the affine modes, which
    settings.cpp already initialises to the mode-1 values' neighbours) against sosba_config_default; the residual pattern."""
    ref, orcdll = libs
    f = np.zeros(18, np.float32)
    i = np.zeros(4, np.int32)
    ref.ref_settings(_p(f), i.ctypes.data_as(C.POINTER(C.c_int)))
    cfg = orc.config_default(640, 480)
    names = ["huber_th", "outlier_th_sum_component", "affine_opt_mode_a", "affine_opt_mode_b", "coarse_cutoff_th", "idepth_fix_prior",
             "idepth_fix_prior_marg_fac", "frame_energy_th_const_weight", "frame_energy_th_n", "frame_energy_th_fac_median",
             "overall_energy_th_weight", "initial_calib_hessian", "initial_rot_prior", "initial_trans_prior", "initial_aff_a_prior",
             "initial_aff_b_prior", "marg_weight_fac", "th_opt_iterations"]
    for k, name in enumerate(names):
        want = np.float32(f[k])
        if name.startswith("affine_opt_mode"):
            want = np.float32(0)      # settingsDefault(mode = 1): setting_affineOptModeA = setting_affineOptModeB = 0 (main.cpp:75-80)
        assert np.float32(getattr(cfg, name)) == want, (name, getattr(cfg, name), f[k])
    assert cfg.gamma_weights_pixel_select == i[0] and cfg.min_opt_iterations == i[1] and i[2] == 6 and i[3] == 8
    a, b = np.zeros(16, np.int32), np.zeros(16, np.int32)
    ref.ref_pattern(a.ctypes.data_as(C.POINTER(C.c_int)))
    orcdll.orc_pin_pattern(b.ctypes.data_as(C.POINTER(C.c_int)))
    assert np.array_equal(a, b) and a.reshape(8, 2).tolist()[4] == [0, 0]
