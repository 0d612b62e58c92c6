#!/bin/bash
set -u
mkdir -p gpurun_out
R=${1:-r2x}
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=90 -k "tensor_core_kernels_match_reference or conv_taps or large_batch or concurrent_streams" > gpurun_out/${R}_gate.log 2>&1
if [ $? -ne 0 ]; then echo "GATE FAILED"; tail -40 gpurun_out/${R}_gate.log; exit 1; fi
tail -2 gpurun_out/${R}_gate.log
C3B_PCONV_PAIR=1 timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=90 -k "tensor_core_kernels_match_reference or conv_taps or large_batch or concurrent_streams" > gpurun_out/${R}_gate_pair.log 2>&1
if [ $? -ne 0 ]; then echo "PAIR GATE FAILED"; tail -40 gpurun_out/${R}_gate_pair.log; exit 1; fi
tail -2 gpurun_out/${R}_gate_pair.log
for pair in 1 0; do
  echo "=== pair=$pair"
  C3B_PCONV_PAIR=$pair C3B_DEBUG_PCONV=1 timeout -k 10 150 python tools/diag.py ptrace convs=${CONVS:-34678} 2>&1 | grep -E "^---|^macro|^  ring|^\[pconv\]" | awk '!seen[$0]++'
done > gpurun_out/${R}_ptrace.log 2>&1
for pair in 1; do
  C3B_PCONV_PAIR=$pair timeout -k 10 300 python bench.py --workloads fa --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench_fa_pair${pair}.json 2> gpurun_out/${R}_bench_fa_pair${pair}.err
  echo "pair=$pair rc=$?"; tail -c 300 gpurun_out/${R}_bench_fa_pair${pair}.err
  python tools/bench_summary.py gpurun_out/${R}_bench_fa_pair${pair}.json 2>&1 | grep -E "^\| (fa|conv|ingest|spp|tail)" 
done
