#!/bin/bash
# A/B of the count kernel variants (MODE bit 0: fixed-trip search, bit 1: batched indel bookkeeping): bit-exactness and device time
set -u
mkdir -p gpurun_out
R=${1:-r2q}
for mode in 0 1 2 3; do
  C3B_PLP_MODE=$mode timeout -k 10 100 python tools/plp_diag.py diag 2>&1 | tail -1
  C3B_PLP_MODE=$mode timeout -k 10 100 python tools/plp_diag.py prof 1048576 10 2>&1 | tail -1
done | tee gpurun_out/${R}_ab.log
C3B_PLP_MODE=3 timeout -k 10 200 python -m pytest tests/test_gpu_pileup_counts.py -m gpu -q --timeout=90 2>&1 | tail -3 | tee -a gpurun_out/${R}_ab.log
