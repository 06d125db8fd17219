#!/bin/bash
# Round-2 final single-GPU evidence: full GPU suite (parity recorder on), bench lines for the three single-GPU workloads,
# ncu launch list + ncu --set full of one step of the 10 M workload, LoG's own loop (when scratch/reference holds a checkout).
#   gpurun --timeout 3000 -- 'bash profiles/r02_final_1gpu.sh f1'
set -u
TAG=${1:-f1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
rm -f gpurun_out/parity.json
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "gpu suite rc=$?" > $OUT/summary.txt
cp gpurun_out/parity.json $OUT/parity.json 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_10m.json 2> $OUT/bench_10m.err
echo "bench 10m rc=$?" >> $OUT/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_10m.json 2> $OUT/bench_reference_10m.err
echo "bench reference rc=$?" >> $OUT/summary.txt
for wl in 100k 1k; do
  timeout 600 python bench.py --steps 50 --warmup 5 --workload $wl > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  echo "bench $wl rc=$?" >> $OUT/summary.txt
done
timeout 600 python bench.py --steps 20 --warmup 3 --order morton --no-e2e --no-cpu-baseline > $OUT/bench_10m_morton.json 2> $OUT/bench_10m_morton.err
# ncu: launch list of 2 steps, then --set full of the last step (plain launches: the graph replays the same kernels)
LGR_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?" >> $OUT/summary.txt
LGR_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(project_|tile_|clamp_|bin_|blend_)' --launch-skip 27 --launch-count 9 -f -o $OUT/step \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?" >> $OUT/summary.txt
ncu -i $OUT/step.ncu-rep --page raw --csv > $OUT/step_raw.csv 2>/dev/null
if [ -d scratch/reference/LoG ]; then
  LGR_REFERENCE_ROOT=scratch/reference timeout 900 python profiles/log_loop_gpu.py --iters 20 --out $OUT/log_loop.json > $OUT/log_loop.log 2>&1
  echo "log loop rc=$?" >> $OUT/summary.txt
fi
tail -n 4 $OUT/gpu_suite.log
for f in $OUT/bench_1*.json; do python -c "import sys,json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms'].items()}, (d.get('e2e') or {}).get('ms_per_step'))"; done
cut -c1-400 $OUT/bench_reference_10m.json
cat $OUT/summary.txt
