#!/bin/bash
set -u
mkdir -p gpurun_out
R=${1:-r2y}
K="tensor_core_kernels_match_reference or conv_taps or large_batch or concurrent_streams"
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=90 -k "$K" > gpurun_out/${R}_gate.log 2>&1
if [ $? -ne 0 ]; then echo "GATE FAILED"; tail -40 gpurun_out/${R}_gate.log; exit 1; fi
tail -2 gpurun_out/${R}_gate.log
C3B_PCONV_PAIR=1 C3B_LSTM2X_RELAXED=1 timeout -k 10 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x --timeout=120 > gpurun_out/${R}_gate_pair.log 2>&1
if [ $? -ne 0 ]; then echo "PAIR GATE FAILED"; tail -40 gpurun_out/${R}_gate_pair.log; exit 1; fi
tail -2 gpurun_out/${R}_gate_pair.log
for pair in 1; do
  echo "=== pair=$pair"
  C3B_PCONV_PAIR=$pair C3B_DEBUG_PCONV=1 timeout -k 10 150 python tools/diag.py ptrace convs=${CONVS:-34678} 2>&1 | grep -E "^---|^macro|^  ring|^\[pconv\]" | awk '!seen[$0]++'
done > gpurun_out/${R}_ptrace.log 2>&1
for cfg in "1 0" "1 4"; do
  set -- $cfg
  C3B_PCONV_PAIR=$1 C3B_PCONV_MT=$2 timeout -k 10 300 python bench.py --workloads fa --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench_fa_pair$1_mt$2.json 2> gpurun_out/${R}_bench_fa_pair$1_mt$2.err
  echo "pair=$1 mt=$2 rc=$?"; tail -c 300 gpurun_out/${R}_bench_fa_pair$1_mt$2.err
  python tools/bench_summary.py gpurun_out/${R}_bench_fa_pair$1_mt$2.json 2>&1 | grep -E "^\| (fa|conv|ingest|spp|tail)" 
done
for rel in 0 1; do
  C3B_LSTM2X_RELAXED=$rel timeout -k 10 300 python bench.py --workloads pileup --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench_p_rel${rel}.json 2> gpurun_out/${R}_bench_p_rel${rel}.err
  echo "relaxed=$rel rc=$?"
  python tools/bench_summary.py gpurun_out/${R}_bench_p_rel${rel}.json 2>&1 | grep -E "^\| (pileup|lstm|proj|ingest|tail)" 
done
