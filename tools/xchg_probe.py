"""Timing probe of the fused stitch + peer exchange: N ranks, weak-scaled configs[1] window, `ba_iterate` loop bodies only.
  SOSBA_XCHG_DEBUG=1 python -m torch.distributed.run --nproc-per-node N tools/xchg_probe.py [iterations]
Every rank prints where its k_stitch_xchg launches spent their time (globaltimer stamps, sosba_destroy)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _scenes import CONFIG_B, scene  # noqa: E402
from sosba_loader import load_package  # noqa: E402

pkg = load_package()
from sos_slam_b200 import binding, problem, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    factor = int(sys.argv[2]) if len(sys.argv) > 2 else world
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = pkg.load()
    sc = synth.replicate_points(scene(**CONFIG_B), factor)
    cfg = lib.config_default(sc.w, sc.h)
    cfg.max_frames = sc.nf + 2
    cfg.min_opt_iterations = 1000
    h = binding.Handle(lib, cfg, device=local)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    for i, img in enumerate(sc.images):
        h.frame_make_images(i, img)
    pts, res = problem.points_of(sc), problem.residuals_of(sc)
    if world > 1:
        uid = [h.lib_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.comm_init(uid[0], rank, world)
        p0, p1 = problem.shard_points(sc.res_point, sc.n_points, world)[rank]
        pts, res = problem.shard_scene_arrays(pts, res, p0, p1)
    val, val0 = problem.calib_of(sc)
    P, keep = h.make_problem(problem.frames_of(sc), val, val0, pts, res)
    h.ba_upload(P)
    for _ in range(3):
        h.ba_optimize(6)
    h.ba_iterate(5)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    h.ba_iterate(n_it)
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"[rank {rank}] {n_it} loop bodies: {1e3 * e0.elapsed_time(e1) / n_it:.2f} us per iteration", flush=True)
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
