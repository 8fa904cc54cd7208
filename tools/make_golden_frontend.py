"""Generates tests/golden/frontend_tiny.npz: the front-end rows (raw frame, pixel selection, immature points, loop-closure
alignment) run by the ORACLE (parity build) on a small seeded window (256x192, 4 frames).  Like ba_tiny.npz these vectors pin
the oracle and the CUDA path against drift; they are not reference outputs.
    python tools/make_golden_frontend.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _frontend_case as fc  # noqa: E402
from sos_slam_b200 import binding  # noqa: E402


def main():
    F = fc.make_inputs()
    orc = binding.Lib(os.path.join(ROOT, "oracle", "_build", "liborc_parity.so"), "orc")
    out = fc.run(orc, F)
    path = os.path.join(ROOT, "tests", "golden", "frontend_tiny.npz")
    np.savez_compressed(path, **F, **{"out_" + k: v for k, v in out.items()})
    print(path, os.path.getsize(path) // 1024, "KiB")
    print({k: (np.asarray(v).shape, np.asarray(v).dtype.name) for k, v in out.items()})
    print("selected", out["sel_n"], "potential", out["sel_potential"], "trace counts", out["trace_counts"].tolist(), "activation",
          np.bincount(out["act_result"].astype(int) + 1, minlength=3), "loop counts", out["loop_counts"])


if __name__ == "__main__":
    main()
