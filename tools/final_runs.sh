set -x
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_end.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-other > gpurun_out/ncu_b.log 2>&1
for K in k_linearize_t k_solve k_accumulate_fused k_step k_stitch_xchg; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -o gpurun_out/r2_end_full_$K -f python bench.py --steps 2 --warmup 1 --no-cpu --no-other > /dev/null 2>&1
done
timeout 400 python bench.py --steps 200 --warmup 3 > gpurun_out/r2_bench_end.json 2> gpurun_out/r2_bench_end.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref_end.json 2>/dev/null
for W in euroc kitti tumvi; do timeout 200 python bench.py --workload $W --steps 100 --warmup 3 --no-cpu --no-other > gpurun_out/r2_bench_${W}_end.json 2>/dev/null; done
timeout 200 python bench.py --workload euroc_imu --steps 50 --warmup 3 > gpurun_out/r2_bench_euroc_imu_end.json 2>/dev/null
timeout 200 python bench.py --workload euroc_imu --impl reference --steps 10 --warmup 3 > gpurun_out/r2_bench_euroc_imu_ref_end.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-other --points-factor 64 > gpurun_out/r2_bench_points_x64_end.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-other --points-factor 8 > gpurun_out/r2_bench_points_x8_end.json 2>/dev/null
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "linearize or optimize or solve or accumulate" > gpurun_out/r2_sanitizer_end.txt 2>&1
tail -5 gpurun_out/r2_sanitizer_end.txt
