#!/bin/bash
set -u
OUT=gpurun_out/${1:-d1}
mkdir -p $OUT
LGR_GRAPH=0 LGR_SYNC_FREE=0 timeout 600 python bench.py --steps 5 --warmup 3 --workload big300k --no-e2e --no-cpu-baseline > $OUT/bench_big300k.json 2> $OUT/bench_big300k.err
python -c "import json; d=json.loads(open('$OUT/bench_big300k.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms'], d['kernel_ms_bin_sort'], d['config'])"
tail -n 3 $OUT/bench_big300k.err
LGR_REFERENCE_ROOT=scratch/reference timeout 900 python profiles/log_loop_gpu.py --iters 10 --out $OUT/log_loop.json > $OUT/log_loop.log 2>&1
python -c "import json; d=json.load(open('$OUT/log_loop.json')); print(d['base_stage_ms_per_iter_stock'], d['base_stage_phase_ms_stock'], d['base_stage_kernel_ms_one_iter'], d.get('base_stage_rendered_rows'))"
