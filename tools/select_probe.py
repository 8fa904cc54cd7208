"""Pixel selection on the GPU vs the single-thread CPU port: host-wall per makeMaps call."""
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
from sos_slam_b200 import binding  # noqa: E402


def main():
    cfg = KITTI if "kitti" in sys.argv else CONFIG_B
    sc = scene(**cfg)
    orc = binding.Lib(os.path.join(ROOT, "oracle", "_build", "liborc_speed.so"), "orc")
    rp = np.random.default_rng(3141592).integers(0, 256, sc.w * sc.h).astype(np.uint8)
    for name, lib in (("gpu", pkg.load()), ("cpu port", orc)):
        h = open_handle(lib, sc)
        h.pixel_selector_set(rp, 3)
        h.pixel_select(0, 1500.0, want_map=False)
        ts = []
        for k in range(10):
            t0 = time.perf_counter()
            r = h.pixel_select(k % sc.nf, 1500.0, want_map=False)
            ts.append(1e3 * (time.perf_counter() - t0))
        print(f"{name}: makeMaps {sc.w}x{sc.h} density 1500: median {np.median(ts):.3f} ms, selected {r['n']}, potential {r['potential']}")
        h.close()


if __name__ == "__main__":
    main()
