#!/bin/bash
# Round-2 single-GPU call C: full GPU suite (parity recorder on), bench lines for the three single-GPU workloads (device-sized
# calls + CUDA graph), LoG's own training loop on the GPU (needs a checkout of the reference under scratch/reference).
set -u
TAG=${1:-c1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
rm -f gpurun_out/parity.json
timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "gpu suite rc=$?" > $OUT/summary.txt
cp gpurun_out/parity.json $OUT/parity.json 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_10m.json 2> $OUT/bench_10m.err
echo "bench rc=$?" >> $OUT/summary.txt



timeout 600 ncu --set full --clock-control none --import-source on -k regex:blend --launch-skip 6 --launch-count 2 -f -o $OUT/blend \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu.log 2>&1
echo "ncu rc=$?" >> $OUT/summary.txt
ncu -i $OUT/blend.ncu-rep --page raw --csv > $OUT/blend_raw.csv 2>/dev/null
for wl in 100k 1k; do
  timeout 600 python bench.py --steps 50 --warmup 5 --workload $wl > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  echo "bench $wl rc=$?" >> $OUT/summary.txt
  LGR_GRAPH=0 LGR_SYNC_FREE=0 timeout 600 python bench.py --steps 50 --warmup 5 --workload $wl --no-e2e --no-cpu-baseline > $OUT/bench_${wl}_hostsized.json 2> $OUT/bench_${wl}_hostsized.err
done
if [ -d scratch/reference/LoG ]; then
  LGR_REFERENCE_ROOT=scratch/reference timeout 900 python profiles/log_loop_gpu.py --iters 20 --out $OUT/log_loop.json > $OUT/log_loop.log 2>&1
  echo "log loop rc=$?" >> $OUT/summary.txt
fi
tail -n 3 $OUT/gpu_suite.log
for f in $OUT/bench_*.json; do python -c "import sys,json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms'].items()}, (d.get('e2e') or {}).get('ms_per_step'), d['config'].get('launch'))"; done
tail -n 30 $OUT/log_loop.json 2>/dev/null; tail -n 5 $OUT/log_loop.log 2>/dev/null | cut -c1-300
cat $OUT/summary.txt
