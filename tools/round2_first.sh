#!/bin/bash
set -u
mkdir -p gpurun_out
R=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${R}_info.txt 2>&1
# gate: the golden-vector parity of the tensor-core path must pass before anything long runs (a hung kernel costs GPU minutes)
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=120 -k "tensor_core_kernels_match_reference or conv_taps" > gpurun_out/${R}_gate.log 2>&1
if [ $? -ne 0 ]; then echo "GATE FAILED"; tail -40 gpurun_out/${R}_gate.log; exit 1; fi
tail -2 gpurun_out/${R}_gate.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=240 > gpurun_out/${R}_pytest.log 2>&1; tail -5 gpurun_out/${R}_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 600 gpurun_out/${R}_bench.err
python - "$R" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/'+sys.argv[1]+'_bench.json').read().strip().splitlines()[-1])
    for k,v in d['workloads'].items():
        print(k, 'value %.3g'%v['value'], 'e2e %.3g'%v['e2e']['value'], 'region %.2fs'%v['timed_region_s'], 'clk', v['clocks'])
        if 'kernels' in v: print('   ', {n:round(x['ms_per_launch']*1e3,1) for n,x in v['kernels'].items()})
        if 'e2e' in v and 'predict_stream' in v['e2e']: print('    ps %.3g sync %.3g'%(v['e2e']['predict_stream']['value'], v['e2e']['synchronous_per_step']['value']))
except Exception as e: print('bench parse failed', e)
PY
