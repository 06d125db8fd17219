#!/bin/bash
set -u
OUT=gpurun_out/${1:-ab1}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -x -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "gpu parity rc=$?" > $OUT/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_10m.json 2> $OUT/bench_10m.err
LGR_CONTRIB_BITS=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_10m_nobits.json 2> $OUT/bench_10m_nobits.err
timeout 600 python bench.py --steps 50 --warmup 5 --workload 100k --no-e2e --no-cpu-baseline > $OUT/bench_100k.json 2> $OUT/bench_100k.err
tail -n 3 $OUT/gpu_suite.log
for f in $OUT/bench_*.json; do python -c "import sys,json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms'].items()})"; done
cat $OUT/summary.txt
