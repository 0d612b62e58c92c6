#!/bin/bash
# Gated run: every stage that exercises a new kernel has a short timeout; a failure skips what depends on it.
set -u
mkdir -p gpurun_out
R=${1:-r2d}
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
echo "gate ok"
timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -q -x --timeout=60 -k "pair_lstm2_kernel" > gpurun_out/${R}_gate_pair2.log 2>&1; P2=$?; tail -2 gpurun_out/${R}_gate_pair2.log
timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -q -x --timeout=60 -k "pair_lstm1" > gpurun_out/${R}_gate_pair1.log 2>&1; P1=$?; tail -12 gpurun_out/${R}_gate_pair1.log
if [ $P2 -eq 0 ] && [ $P1 -eq 0 ]; then timeout 120 python tools/pair_trace.py 1024 > gpurun_out/${R}_pair_trace.txt 2>&1; cat gpurun_out/${R}_pair_trace.txt | tail -16; fi
run p_default --workloads pileup
if [ $P2 -eq 0 ]; then run p_pair2 --workloads pileup --opt lstm2_impl=1; fi
if [ $P2 -eq 0 ] && [ $P1 -eq 0 ]; then
  run p_pair12 --workloads pileup --opt lstm2_impl=1 --opt lstm1_impl=1
  for n in 29; do C3B_PROJ_CTAS=$n run p_pair12_proj$n --workloads pileup --opt lstm2_impl=1 --opt lstm1_impl=1; done
fi
run fa --workloads fa
DESEL=""
if [ $P1 -ne 0 ]; then DESEL="--deselect tests/test_gpu_round2.py::test_pair_lstm1_and_lstm2_kernels_match_reference_goldens --deselect tests/test_gpu_round2.py::test_pair_lstm1_kernel_full_size_deep_and_ragged"; fi
timeout 600 python -m pytest tests -m gpu -q --timeout=120 $DESEL > gpurun_out/${R}_pytest.log 2>&1; tail -4 gpurun_out/${R}_pytest.log
