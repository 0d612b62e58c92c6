#!/bin/bash
set -u
mkdir -p gpurun_out
R=${1:-r2c}
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=120 -k "tensor_core_kernels_match_reference or conv_taps" > gpurun_out/${R}_gate.log 2>&1
if [ $? -ne 0 ]; then echo "GATE FAILED"; tail -40 gpurun_out/${R}_gate.log; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout=200 -k "pair" > gpurun_out/${R}_pytest_pair.log 2>&1; tail -8 gpurun_out/${R}_pytest_pair.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > gpurun_out/${R}_bench_${name}.json 2> gpurun_out/${R}_bench_${name}.err; python - "$R" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/%s_bench_%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
    for k,v in d['workloads'].items():
        print(sys.argv[2], k, 'value %.4g'%v['value'], 'e2e %.4g'%v['e2e']['value'], 'clk', v['clocks']['sm_mhz'], v['clocks']['reasons'], {n:round(x['ms_per_launch']*1e3,1) for n,x in v.get('kernels',{}).items()})
except Exception as e: print(sys.argv[2], 'bench parse failed', e); print(open('gpurun_out/%s_bench_%s.err'%(sys.argv[1],sys.argv[2])).read()[-600:])
PY
}
run p_default --workloads pileup
run p_pair2 --workloads pileup --opt lstm2_impl=1
run p_pair12 --workloads pileup --opt lstm2_impl=1 --opt lstm1_impl=1
for n in 6 20 29; do C3B_PROJ_CTAS=$n run p_pair12_proj$n --workloads pileup --opt lstm2_impl=1 --opt lstm1_impl=1; done
timeout 900 python -m pytest tests -m gpu -q --timeout=240 > gpurun_out/${R}_pytest.log 2>&1; tail -4 gpurun_out/${R}_pytest.log
