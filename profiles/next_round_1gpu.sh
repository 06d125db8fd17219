#!/bin/bash
# First single-GPU call of the next round (see profiles/next_round.md).  Run from the repo root on the GPU box:
#   gpurun --timeout 2400 -- 'bash profiles/next_round_1gpu.sh'
# Everything is written under gpurun_out/next/ ; nothing here is a benchmark number (sanitizer / pytest runs).
set -u
mkdir -p gpurun_out/next
python __graft_entry__.py > gpurun_out/next/build.log 2>&1
# 1. the rows and modes built without a GPU at the end of round 1 (non-strict xfail in the default run)
timeout 900 python -m pytest tests/test_zz_gpu_new_rows.py tests/test_zz_gpu_shard_mode.py -q --runxfail -p no:cacheprovider \
    > gpurun_out/next/new_rows.log 2>&1
echo "new rows rc=$?" >> gpurun_out/next/summary.txt
# 2. the verified suite (kernels touched since: LSD tie vote, wsum alignment, compute_radius refactor, project templates)
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/next/gpu_suite.log 2>&1
echo "gpu suite rc=$?" >> gpurun_out/next/summary.txt
# 3. sanitizer, FULL logs this time
timeout 900 compute-sanitizer --tool racecheck python tests/sanitize_workload.py > gpurun_out/next/racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/next/summary.txt
timeout 600 compute-sanitizer --tool memcheck python tests/sanitize_workload.py > gpurun_out/next/memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/next/summary.txt
# 4. one bench line to see that nothing moved (10 M, 1 GPU)
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/next/bench_10m.json 2> gpurun_out/next/bench_10m.err
echo "bench rc=$?" >> gpurun_out/next/summary.txt
tail -3 gpurun_out/next/new_rows.log gpurun_out/next/gpu_suite.log; tail -4 gpurun_out/next/racecheck.log; cat gpurun_out/next/summary.txt
