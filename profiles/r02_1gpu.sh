#!/bin/bash
# Round-2 single-GPU check: parity suite, one 10 M bench line, ncu --set full of the blend kernels of one step.
#   gpurun --timeout 1500 -- 'bash profiles/r02_1gpu.sh [tag]'
set -u
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "gpu suite rc=$?" > $OUT/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_10m.json 2> $OUT/bench_10m.err
echo "bench rc=$?" >> $OUT/summary.txt
# last step only: 3 warm-up steps + 1 timed step, 2 blend launches per step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blend --launch-skip 6 --launch-count 2 -f -o $OUT/blend \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu.log 2>&1
echo "ncu rc=$?" >> $OUT/summary.txt
ncu -i $OUT/blend.ncu-rep --page raw --csv > $OUT/blend_raw.csv 2>/dev/null
tail -3 $OUT/gpu_suite.log; cat $OUT/bench_10m.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms'])"; cat $OUT/summary.txt
