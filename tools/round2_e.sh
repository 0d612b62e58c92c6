#!/bin/bash
set -u
mkdir -p gpurun_out
R=${1:-r2e}
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run() { name=$1; shift; timeout 240 $B "$@" > gpurun_out/${R}_bench_${name}.json 2> gpurun_out/${R}_bench_${name}.err; python - "$R" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/%s_bench_%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
    for k,v in d['workloads'].items():
        print(sys.argv[2], k, 'value %.4g'%v['value'], 'e2e %.4g'%v['e2e']['value'], 'ps %.4g'%v['e2e'].get('predict_stream',{}).get('value',0), 'sync %.4g'%v['e2e'].get('synchronous_per_step',{}).get('value',0), 'win %.4g'%v['e2e'].get('forward_windows',{}).get('value',0), 'clk', v['clocks']['sm_mhz'], v['clocks']['reasons'], {n:(round(x['ms_per_launch']*1e3,1), int(x.get('ctas',0))) for n,x in v.get('kernels',{}).items()})
except Exception as e: print(sys.argv[2], 'bench parse failed', e); print(open('gpurun_out/%s_bench_%s.err'%(sys.argv[1],sys.argv[2])).read()[-800:])
PY
}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=90 -k "tensor_core_kernels_match_reference or conv_taps" > gpurun_out/${R}_gate.log 2>&1
if [ $? -ne 0 ]; then echo "GATE FAILED"; tail -40 gpurun_out/${R}_gate.log; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x --timeout=60 -k "pair" > gpurun_out/${R}_gate_pair.log 2>&1; P=$?; tail -3 gpurun_out/${R}_gate_pair.log
if [ $P -ne 0 ]; then echo "PAIR GATE FAILED"; tail -30 gpurun_out/${R}_gate_pair.log; exit 1; fi
timeout 120 python tools/pair_trace.py 1024 > gpurun_out/${R}_pair_trace.txt 2>&1; tail -16 gpurun_out/${R}_pair_trace.txt
for extra in "$@"; do :; done
run p_pair2 --workloads pileup --opt lstm2_impl=1
run p_pair12 --workloads pileup --opt lstm2_impl=1 --opt lstm1_impl=1
