#!/bin/bash
set -u
mkdir -p gpurun_out
R=${1:-r2i}
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run() { name=$1; shift; timeout 240 $B "$@" > gpurun_out/${R}_bench_${name}.json 2> gpurun_out/${R}_bench_${name}.err; python - "$R" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/%s_bench_%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
    for k,v in d['workloads'].items():
        print(sys.argv[2], k, 'value %.4g'%v['value'], 'e2e %.4g'%v['e2e']['value'], 'clk', v['clocks']['sm_mhz'], v['clocks']['reasons'], {n:(round(x['ms_per_launch']*1e3,1), int(x.get('ctas',0))) for n,x in v.get('kernels',{}).items()})
except Exception as e: print(sys.argv[2], 'bench parse failed', e); print(open('gpurun_out/%s_bench_%s.err'%(sys.argv[1],sys.argv[2])).read()[-800:])
PY
}
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x --timeout=90 -k "tensor_core_kernels_match_reference or pileup_1024 or tiles_agree or warpgroup or concurrent_streams or deep" > gpurun_out/${R}_gate.log 2>&1
if [ $? -ne 0 ]; then echo "GATE FAILED"; tail -40 gpurun_out/${R}_gate.log; exit 1; fi
tail -2 gpurun_out/${R}_gate.log
timeout 200 python tools/diag.py trace lstm_tile=64 2>&1 | tail -3
run p_default --workloads pileup
