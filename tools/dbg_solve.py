"""debug: k_solve x vs numpy on the returned (H_final, b_final) for several window sizes"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _scenes import open_handle, scene, upload
from sosba_loader import load_package
pkg = load_package()
lib = pkg.load()
for nf in (3, 4, 5, 6, 7, 8, 9, 10, 12):
    sc = scene(w=320, h=240, nf=nf, n_points=60 * nf, seed=3)
    h = open_handle(lib, sc)
    upload(h, sc)
    h.reset_oob(); h.linearize_all(False); h.apply_res()
    x, Hf, bf = h.solve_system()
    S = 1.0 / np.sqrt(np.diag(Hf) + 10)
    xn = S * np.linalg.solve(S[:, None] * Hf * S[None, :], S * bf)
    d = x - xn
    err = np.sqrt(abs(d @ Hf @ d)) / np.sqrt(abs(xn @ Hf @ xn))
    bad = np.nonzero(np.abs(d) > 1e-6 * np.abs(xn).max())[0]
    print(f"nf={nf} D={4+8*nf} H-norm rel err {err:.3e}  max|x| {np.abs(xn).max():.3e}  bad idx {bad[:12]} n_bad {len(bad)}")
    h.close()
