"""Prints the LM pass records of libsosba and the oracle side by side for the hypotheses of the parity test."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _scenes import SMALL, CONFIG_B, open_handle, scene
from _track_case import coarse_depth_input, hypotheses, ref_affine
from sosba_loader import load_package
pkg = load_package()
from sos_slam_b200 import binding
gpu = pkg.load()
orc = binding.Lib(os.path.join(ROOT, "oracle", "_build", "liborc_parity.so"), "orc")
sc = scene(**(CONFIG_B if len(sys.argv) > 1 else SMALL))
ho, hg = open_handle(orc, sc), open_handle(gpu, sc)
cpt, hdi = coarse_depth_input(ho, sc)
Ttrue, hyps = hypotheses(sc, n_extra=6, seed=4)
ref_aff, ref_exp, new_exp = ref_affine(sc)
outs = []
for h in (hg, ho):
    h.tracker_make_k(sc.K.astype(np.float32))
    h.tracker_make_coarse_depth(sc.nf - 1, cpt, hdi)
    outs.append(h.tracker_track(sc.nf - 2, ref_exp, new_exp, ref_aff, h.levels - 1, hyps))
for i, (g, o) in enumerate(zip(*outs)):
    print(i, "gpu", g["pass_lvl"], g["pass_iterations"], [bin(x) for x in g["pass_accept"]], [bin(x) for x in g["pass_tie"]], g["pass_residual"])
    print(i, "orc", o["pass_lvl"], o["pass_iterations"], [bin(x) for x in o["pass_accept"]], [bin(x) for x in o["pass_tie"]], o["pass_residual"])
from _track_case import quat_to_T
print("pose agreement (max |Tg - To|, moved, aff diff):")
for i, (hy, g, o) in enumerate(zip(hyps, *outs)):
    Tg, To, T0 = quat_to_T(g["q"], g["t"]), quat_to_T(o["q"], o["t"]), quat_to_T(hy["q"], hy["t"])
    print(i, f"{np.abs(Tg - To).max():.3e} rot {np.abs(Tg[:3,:3] - To[:3,:3]).max():.3e} moved {np.abs(To - T0).max():.3e} aff {np.abs(g['aff_g2l'] - o['aff_g2l'])}", "t", Tg[:3, 3] - To[:3, 3])
