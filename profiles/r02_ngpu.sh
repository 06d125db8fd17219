#!/bin/bash
# Round-2 multi-GPU call: shard mode (default) and band mode on N GPUs, multirank tests.
#   gpurun --gpus N --timeout 1500 -- 'N=8 bash profiles/r02_ngpu.sh tag'
set -u
N=${N:-2}
TAG=${1:-n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {  # $1 = label, rest = extra bench args ; env in front
  local label=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 3 "$@" > $OUT/scale_${N}_$label.json 2> $OUT/scale_${N}_$label.err
  echo "$label N=$N rc=$?" >> $OUT/summary.txt
}
run shard
if [ "${NOGRAPH:-1}" = "1" ]; then LGR_GRAPH=0 run shard_nograph --no-e2e; fi
if [ "${HOSTSIZED:-1}" = "1" ]; then LGR_GRAPH=0 LGR_SYNC_FREE=0 run shard_hostsized --no-e2e; fi
if [ "${BAND:-1}" = "1" ]; then LGR_MULTI=band run band --no-e2e; fi
if [ "${TESTS:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_zz_gpu_shard_multirank.py -q -m gpu --runxfail -p no:cacheprovider > $OUT/multirank.log 2>&1
  echo "multirank rc=$?" >> $OUT/summary.txt
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "gather_fused or no_binned" > $OUT/new_tests.log 2>&1
  echo "new tests rc=$?" >> $OUT/summary.txt
  tail -n 3 $OUT/multirank.log; tail -n 3 $OUT/new_tests.log
fi
if [ "${BIG:-0}" = "1" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --gpus $N --steps 5 --warmup 3 --workload 50m4k --no-e2e > $OUT/scale_${N}_50m4k.json 2> $OUT/scale_${N}_50m4k.err
  echo "50m4k N=$N rc=$?" >> $OUT/summary.txt
fi
python - <<PY
import json, glob
for f in sorted(glob.glob('$OUT/scale_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 3), 'ms/step; e2e', (d.get('e2e') or {}).get('ms_per_step'), d.get('phase_ms_rank0'))
        print('   kernels', {k: round(v, 3) for k, v in d['kernel_ms'].items()}, {k: round(v, 3) for k, v in (d.get('kernel_ms_exchange') or {}).items()})
        print('   parity', {k: v for k, v in (d.get('parity') or {}).items() if k != 'what'})
    except Exception as e:
        print(f, 'unreadable:', e)
PY
for f in $OUT/*.err; do tail -n 3 $f; done | tail -n 30
cat $OUT/summary.txt
