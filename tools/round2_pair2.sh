#!/bin/bash
# CTA-pair convolution iteration: gated parity, CTA-0 traces, FA bench (pair form on)
set -u
mkdir -p gpurun_out
R=${1:-r2q}
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=90 -k "tensor_core_kernels_match_reference or conv_taps or large_batch or concurrent_streams" > gpurun_out/${R}_gate.log 2>&1
if [ $? -ne 0 ]; then echo "GATE FAILED"; tail -40 gpurun_out/${R}_gate.log; exit 1; fi
tail -2 gpurun_out/${R}_gate.log
C3B_DEBUG_PCONV=1 timeout -k 10 150 python tools/diag.py ptrace convs=${CONVS:-134678} 2>&1 | grep -E "^---|^macro|^\[pconv\]" | awk '!seen[$0]++' > gpurun_out/${R}_ptrace.log 2>&1
for pair in ${PAIRS:-1}; do
  C3B_PCONV_PAIR=$pair timeout -k 10 300 python bench.py --workloads fa --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench_fa_pair${pair}.json 2> gpurun_out/${R}_bench_fa_pair${pair}.err
  echo "pair=$pair rc=$?"; tail -c 300 gpurun_out/${R}_bench_fa_pair${pair}.err
  python tools/bench_summary.py gpurun_out/${R}_bench_fa_pair${pair}.json 2>&1 | grep -E "^\| (fa|conv|ingest|spp|tail)" 
done
