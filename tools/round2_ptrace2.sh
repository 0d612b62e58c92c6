#!/bin/bash
set -u
mkdir -p gpurun_out
R=${1:-r2w}
for cfg in "0 0 0" "0 0 1" "0 1 0" "1 0 1" "1 1 0"; do
  set -- $cfg
  echo "=== pair=$1 skipw=$2 tapouter=$3"
  C3B_PCONV_PAIR=$1 C3B_PCONV_SKIPW=$2 C3B_PCONV_TAPOUTER=$3 timeout -k 10 150 python tools/diag.py ptrace convs=${CONVS:-47} 2>&1 | grep -E "^---|^macro" 
done > gpurun_out/${R}_ptrace.log 2>&1
tail -5 gpurun_out/${R}_ptrace.log
