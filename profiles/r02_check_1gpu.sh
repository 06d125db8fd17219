#!/bin/bash
# quick regression: GPU suite, smoke(), one bench line
set -u
OUT=gpurun_out/${1:-k1}
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "gpu suite rc=$?" > $OUT/summary.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/summary.txt
timeout 900 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_10m.json 2> $OUT/bench_10m.err
echo "bench rc=$?" >> $OUT/summary.txt
tail -n 3 $OUT/gpu_suite.log; tail -n 2 $OUT/smoke.log
python -c "import json; d=json.loads(open('$OUT/bench_10m.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms'], d['kernel_ms_bin_sort'])"
cat $OUT/summary.txt
