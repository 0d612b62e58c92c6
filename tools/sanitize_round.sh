#!/bin/bash
# compute-sanitizer evidence (run under gpurun): memcheck / racecheck / synccheck / initcheck on one small forward of each
# network, LSTM tiles 16/32/64 and 1/2 epilogue warpgroups.  Logs -> gpurun_out/${R}_san_*.log; one summary line each.
set -u
R=${1:-r2}
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {   # name tool args...
    local name=$1 tool=$2; shift 2
    timeout 300 $SAN --tool $tool --print-limit 20 python tools/sanitize_case.py "$@" > gpurun_out/${R}_san_${name}_${tool}.log 2>&1
    local rc=$?
    echo "${name} ${tool} rc=${rc} :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok \(' gpurun_out/${R}_san_${name}_${tool}.log | tr '\n' ' ')"
}
for tool in memcheck racecheck synccheck; do
    run p_t16 $tool p 128 lstm_tile=16
    run p_t64 $tool p 256 lstm_tile=64
    run p_t32_wg1 $tool p 128 lstm_tile=32 lstm_wg=1
    run f_b16 $tool f 16
done
run p_t64 initcheck p 256 lstm_tile=64
run f_b16 initcheck f 16
