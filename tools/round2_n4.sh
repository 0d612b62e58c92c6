#!/bin/bash
set -u
mkdir -p gpurun_out
R=${1:-r2n4}
N=${2:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 \
    > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
tail -c 600 gpurun_out/${R}_bench.err
python - "$R" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/%s_bench.json'%sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print('n_gpus', d['n_gpus'], 'broadcast', d['weight_broadcast'])
    for k,v in d['workloads'].items():
        print(k, v['scaling'], 'value %.4g'%v['value'], 'e2e %.4g'%v['e2e']['value'], 'region %.2f'%v['timed_region_s'])
except Exception as e: print('parse failed', e)
PY
