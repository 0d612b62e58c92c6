#!/bin/bash
# pconv cycle traces (CTA 0) with / without the pair form and with the streamed-weight waits removed (timing experiment)
set -u
mkdir -p gpurun_out
R=${1:-r2t}
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  echo "=== pair=$1 skipw=$2"
  C3B_PCONV_PAIR=$1 C3B_PCONV_SKIPW=$2 C3B_DEBUG_PCONV=1 timeout -k 10 150 python tools/diag.py ptrace convs=${CONVS:-13467} 2>&1 | grep -E "^---|^macro|^\[pconv\]" | awk '!seen[$0]++'
done > gpurun_out/${R}_ptrace.log 2>&1
tail -5 gpurun_out/${R}_ptrace.log
