#!/bin/bash
# CTA-pair convolution check: gated parity tests, then FA bench with and without the pair form, then sanitizer on FA.
set -u
mkdir -p gpurun_out
R=${1:-r2p}
C3B_DEBUG_PCONV=1 timeout -k 10 120 python tools/sanitize_case.py f 16 > gpurun_out/${R}_case16.log 2>&1
echo "case16 rc=$?"; grep -E "^\[pconv\]|^ok" gpurun_out/${R}_case16.log | sort | uniq -c | head -20
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=90 -k "tensor_core_kernels_match_reference or conv_taps or large_batch or concurrent_streams" > gpurun_out/${R}_gate.log 2>&1
if [ $? -ne 0 ]; then echo "GATE FAILED"; tail -40 gpurun_out/${R}_gate.log; exit 1; fi
tail -2 gpurun_out/${R}_gate.log
for pair in 1 0; do
  C3B_PCONV_PAIR=$pair timeout -k 10 300 python bench.py --workloads fa --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench_fa_pair${pair}.json 2> gpurun_out/${R}_bench_fa_pair${pair}.err
  echo "pair=$pair rc=$?"; tail -c 300 gpurun_out/${R}_bench_fa_pair${pair}.err
  python tools/bench_summary.py gpurun_out/${R}_bench_fa_pair${pair}.json 2>&1 | grep -E "^\| (fa|conv|ingest|spp|tail)" 
done
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 300 $SAN --tool $tool --print-limit 5 python tools/sanitize_case.py f 16 > gpurun_out/${R}_san_fa_${tool}.log 2>&1
  echo "$tool [full-alignment, pair convs] :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok \(' gpurun_out/${R}_san_fa_${tool}.log | tr '\n' ' ')"
done
