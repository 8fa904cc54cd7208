"""Point-sharded optimize on WORLD_SIZE GPUs (one process per GPU; fused stitch + peer-memory exchange, and the NCCL
fallback) against the single-GPU result.
Launched by tests/test_gpu_multi.py:  python -m torch.distributed.run --nproc-per-node N tools/mgpu_check.py [scene]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _scenes import CONFIG_B, SMALL, scene  # noqa: E402
from sosba_loader import load_package  # noqa: E402

pkg = load_package()
from sos_slam_b200 import binding, problem  # noqa: E402


def run(lib, sc, local, shard=None, uid=None, rank=0, world=1, iters=6):
    cfg = lib.config_default(sc.w, sc.h)
    cfg.max_frames = sc.nf + 2
    h = binding.Handle(lib, cfg, device=local)
    for i, img in enumerate(sc.images):
        h.frame_make_images(i, img)
    pts, res = problem.points_of(sc), problem.residuals_of(sc)
    mode = "single"
    if shard is not None:
        h.comm_init(uid, rank, world)
        mode = "peer-memory" if h.comm_uses_peer_memory() else "nccl"
        pts, res = problem.shard_scene_arrays(pts, res, *shard)
    val, val0 = problem.calib_of(sc)
    P, keep = h.make_problem(problem.frames_of(sc), val, val0, pts, res)
    out = h.optimize(P, iters)
    r = h.problem_result(P, keep)
    st = h.get_state()["state"]
    h.close()
    out["exchange"] = mode
    return out, r, st


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = pkg.load()
    sc = scene(**(CONFIG_B if len(sys.argv) > 1 and sys.argv[1] == "configB" else SMALL))
    uid = [None]
    if rank == 0:
        import ctypes
        buf = (ctypes.c_uint8 * 128)()
        assert lib.f("comm_unique_id")(buf) == 0
        uid[0] = bytes(buf)
    dist.broadcast_object_list(uid, src=0)
    def new_uid():
        u = [None]
        if rank == 0:
            import ctypes
            b = (ctypes.c_uint8 * 128)()
            assert lib.f("comm_unique_id")(b) == 0
            u[0] = bytes(b)
        dist.broadcast_object_list(u, src=0)
        return u[0]

    shards = problem.shard_points(sc.res_point, sc.n_points, world)
    out, r, st = run(lib, sc, local, shards[rank], uid[0], rank, world)
    assert out["exchange"] == "peer-memory" or os.environ.get("SOSBA_ALLOW_NO_P2P"), "peer memory could not be mapped on this box"
    # the same window with the NCCL all-reduce of the un-stitched tables (the fallback): both must lead to the same optimised window
    os.environ["SOSBA_COMM_NCCL"] = "1"
    out_n, r_n, st_n = run(lib, sc, local, shards[rank], new_uid(), rank, world)
    del os.environ["SOSBA_COMM_NCCL"]
    assert out_n["exchange"] == "nccl"
    dn = float(np.abs(r_n["state"] - r["state"]).max())
    agree = (out_n["iterations"] == out["iterations"] and dn <= 1e-9 * max(1.0, float(np.abs(r["state"]).max()))
             and np.allclose(r_n["frame_energy_th"], r["frame_energy_th"], rtol=1e-6) and np.array_equal(st_n, st))
    if rank == 0:
        print(f"exchange {out['exchange']} vs {out_n['exchange']}: iterations {out['iterations']} / {out_n['iterations']}, max state difference {dn:.2e}, "
              f"residual states equal {np.array_equal(st_n, st)}")
    assert agree, "peer-memory exchange and NCCL all-reduce disagree"
    # every rank must hold the same frame states / thresholds / iteration count
    t = torch.tensor(np.concatenate([r["state"].ravel(), r["frame_energy_th"].astype(np.float64), [out["iterations"], out["energy_final"]]]), device="cuda")
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    d = (hi - lo).cpu().numpy()
    ns = r["state"].size
    # the exchange sums the ranks' parts in rank order and the stitch has no atomics: every rank holds the same bits
    same = bool(np.all(d == 0))
    if not same:
        print(f"[rank {rank}] rank disagreement: state {np.abs(d[:ns]).max():.3e} th {np.abs(d[ns:ns + sc.nf]).max():.3e} iterations {d[-2]} energy {d[-1]:.3e}", flush=True)
    assert same, "ranks disagree on the optimised window"
    ok = True
    if rank == 0:
        o1, r1, st1 = run(lib, sc, local)
        upd = np.abs(r1["state"] - sc.state).max(axis=0) + 1e-12
        dev = np.abs(r["state"] - r1["state"]).max(axis=0)
        p0, p1 = shards[0]
        print(f"world {world}: iterations {out['iterations']} vs {o1['iterations']}, energy_final {out['energy_final']:.6f} vs {o1['energy_final']:.6f}, "
              f"res_in_a {out['res_in_a']} vs {o1['res_in_a']}, state dev/update {np.max(dev / upd):.2e}, th {r['frame_energy_th'][-1]} vs {r1['frame_energy_th'][-1]}")
        ok = (out["iterations"] == o1["iterations"] and abs(out["energy_final"] - o1["energy_final"]) <= 1e-3 * abs(o1["energy_final"])
              and abs(out["res_in_a"] - o1["res_in_a"]) <= 2 and (dev <= 5e-3 * upd).all()
              and np.allclose(r["frame_energy_th"], r1["frame_energy_th"], rtol=1e-3)
              and np.allclose(r["idepth"], r1["idepth"][p0:p1], rtol=2e-3, atol=1e-5))
        print("MGPU_CHECK", "OK" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
