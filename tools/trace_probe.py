"""Immature-point row on the GPU: constructor, two traceNewCoarse passes, activation, on the
configs[1] window with 2000 candidates per older keyframe.  Prints host-wall times per call (host buffers in and out);
run under ncu for the kernel times (profiles/README.md)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _scenes import CONFIG_B, KITTI, open_handle, scene  # noqa: E402
from sosba_loader import load_package  # noqa: E402

pkg = load_package()
from sos_slam_b200 import synth  # noqa: E402


def main():
    cfg = KITTI if "kitti" in sys.argv else CONFIG_B
    per_host = 2000
    sc = scene(**cfg)
    h = open_handle(pkg.load(), sc)
    first = sc.nf - 2
    case = synth.trace_case(sc, first, n_per_host=per_host, seed=3)
    keep = case["host"] < first
    host, u, v = case["host"][keep], case["u"][keep], case["v"][keep]
    t0 = time.perf_counter()
    parts = [h.immature_init(int(x), u[host == x], v[host == x]) for x in np.unique(host)]
    t_init = time.perf_counter() - t0
    pts = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    for rep in range(3):
        q = {k: x.copy() for k, x in pts.items()}
        for nfm in (first, first + 1):
            c = synth.trace_case(sc, nfm, n_per_host=1)
            t0 = time.perf_counter()
            cnt = h.trace_immature(nfm, host, c["KRKi"], c["Kt"], c["aff"], q)
            print(f"trace into frame {nfm}: {host.size} points {1e3 * (time.perf_counter() - t0):.3f} ms counts {cnt.tolist()}")
        ok = np.isfinite(q["idepth_max"])
        sub = {k: x[ok] for k, x in q.items()}
        win = synth.activation_case(sc)
        t0 = time.perf_counter()
        res, idepth, st = h.optimize_immature(np.arange(sc.nf), win["RTll"], win["tTll"], win["aff"], win["calib"], host[ok], sub)
        print(f"activation: {ok.sum()} points {1e3 * (time.perf_counter() - t0):.3f} ms activated {(res == 1).sum()} skip {(res == 0).sum()} delete {(res == -1).sum()}")
    print(f"immature_init: {host.size} points in {len(parts)} calls {1e3 * t_init:.3f} ms")
    h.close()


if __name__ == "__main__":
    main()
