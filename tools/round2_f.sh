#!/bin/bash
# Round-2 evidence run: full GPU test suite, smoke, the driver's bench command (default options), sanitizer on the pair kernel,
# ncu launch lists + full captures.
set -u
mkdir -p gpurun_out
R=${1:-r2f}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=90 -k "tensor_core_kernels_match_reference or conv_taps" > gpurun_out/${R}_gate.log 2>&1
if [ $? -ne 0 ]; then echo "GATE FAILED"; tail -40 gpurun_out/${R}_gate.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q --timeout=120 > gpurun_out/${R}_pytest.log 2>&1; tail -4 gpurun_out/${R}_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 400 gpurun_out/${R}_bench.err
python tools/bench_summary.py gpurun_out/${R}_bench.json > gpurun_out/${R}_bench_summary.md 2>&1; head -12 gpurun_out/${R}_bench_summary.md
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 300 $SAN --tool $tool --print-limit 5 python tools/sanitize_case.py p 256 lstm_tile=64 > gpurun_out/${R}_san_pairdefault_${tool}.log 2>&1
  echo "$tool [default pileup path, pair LSTM2] :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok \(' gpurun_out/${R}_san_pairdefault_${tool}.log | tr '\n' ' ')"
  timeout 300 $SAN --tool $tool --print-limit 5 python tools/sanitize_case.py p 256 lstm1_impl=1 > gpurun_out/${R}_san_pair12_${tool}.log 2>&1
  echo "$tool [pair LSTM1 + LSTM2] :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok \(' gpurun_out/${R}_san_pair12_${tool}.log | tr '\n' ' ')"
  timeout 300 $SAN --tool $tool --print-limit 5 python tools/sanitize_case.py f 16 > gpurun_out/${R}_san_fa_${tool}.log 2>&1
  echo "$tool [full-alignment] :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok \(' gpurun_out/${R}_san_fa_${tool}.log | tr '\n' ' ')"
done
bash tools/profile_round2.sh ${R}
