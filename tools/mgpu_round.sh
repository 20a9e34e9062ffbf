#!/bin/bash
# One multi-GPU box visit: correctness of the sharded paths, then the three multi-GPU bench workloads.
# usage (under gpurun --gpus N): bash tools/mgpu_round.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_${N}_check.log 2>&1; echo "check rc=$?"; tail -2 gpurun_out/mgpu_${N}_check.log
$TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/mgpu_${N}_bench.json 2> gpurun_out/mgpu_${N}_bench.err; echo "bench rc=$?"
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f'gpurun_out/mgpu_{n}_bench.json') if l.startswith('{')][-1])
    for k, v in d.get('zslab_1024', {}).items():
        if isinstance(v, dict):
            print('zslab_1024', k, n, 'GPUs:', round(v['value'], 1), 'Gvox/s, distribute', round(v['distribute_ms'], 2), 'ms, resample', round(v['resample_ms'], 2), 'ms', v.get('path'), v.get('largest_receive_share'))
except Exception as e:
    print('zslab unreadable', e)
PY
for f in bench; do python - "$f" "$N" <<'PY'
import json, sys
f, n = sys.argv[1], sys.argv[2]
try:
    lines = [l for l in open(f'gpurun_out/mgpu_{n}_{f}.json') if l.startswith('{')]
    d = json.loads(lines[-1])
    print(f, n, 'GPUs:', round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 3), 'e2e', (d.get('e2e') or {}).get('value'))
except Exception as e:
    print(f, 'unreadable', e)
PY
done
