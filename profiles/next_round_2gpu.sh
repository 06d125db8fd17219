#!/bin/bash
# Second call: shard mode against band mode on 2 GPUs (then repeat with --gpus 4 / 8 by editing N).
#   gpurun --gpus 2 --timeout 1500 -- 'N=2 bash profiles/next_round_2gpu.sh'
set -u
N=${N:-2}
mkdir -p gpurun_out/next
run() {  # $1 = label, env in front
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/next/scale_${N}_$1.json 2> gpurun_out/next/scale_${N}_$1.err
  echo "$1 N=$N rc=$?" >> gpurun_out/next/summary.txt
}
LGR_MULTI=band run band
LGR_MULTI=shard run shard
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_zz_gpu_shard_multirank.py -q -m gpu --runxfail -p no:cacheprovider > gpurun_out/next/multirank.log 2>&1
echo "multirank rc=$?" >> gpurun_out/next/summary.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/next/scale_*_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 3), 'ms/step; e2e', (d.get('e2e') or {}).get('ms_per_step'), d.get('phase_ms_rank0'))
    except Exception as e:
        print(f, 'unreadable:', e)
PY
cat gpurun_out/next/summary.txt
