#!/bin/bash
# Round-2 single-GPU call B: parity suite, the full default bench line, A/B of the ranked binning, morton order, racecheck.
set -u
TAG=${1:-b1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "gpu suite rc=$?" > $OUT/summary.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_10m.json 2> $OUT/bench_10m.err
echo "bench rc=$?" >> $OUT/summary.txt
LGR_RANKED_BIN=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_10m_unranked.json 2> $OUT/bench_10m_unranked.err
echo "bench unranked rc=$?" >> $OUT/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --order morton > $OUT/bench_10m_morton.json 2> $OUT/bench_10m_morton.err
echo "bench morton rc=$?" >> $OUT/summary.txt
LGR_RANKED_BIN=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --order morton > $OUT/bench_10m_morton_unranked.json 2> $OUT/bench_10m_morton_unranked.err
echo "bench morton unranked rc=$?" >> $OUT/summary.txt
timeout 900 compute-sanitizer --tool racecheck python tests/sanitize_workload.py > $OUT/racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/summary.txt
timeout 600 compute-sanitizer --tool memcheck python tests/sanitize_workload.py > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/summary.txt
tail -3 $OUT/gpu_suite.log
for f in $OUT/bench_10m*.json; do python -c "import sys,json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms'].items()}, (d.get('e2e') or {}).get('ms_per_step'))"; done
tail -4 $OUT/racecheck.log; tail -3 $OUT/memcheck.log; cat $OUT/summary.txt
