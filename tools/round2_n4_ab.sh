#!/bin/bash
# A/B of the count kernel's ILP (reads resolved side by side): bit-exactness and device time for ILP = 1, 2, 4
set -u
mkdir -p gpurun_out
R=${1:-r2o}
for ilp in 1 2 4; do
  C3B_PLP_ILP=$ilp timeout -k 10 100 python tools/plp_diag.py diag 2>&1 | tail -1
  C3B_PLP_ILP=$ilp timeout -k 10 100 python tools/plp_diag.py prof 1048576 12 2>&1 | tail -1
done | tee gpurun_out/${R}_ab.log
C3B_PLP_ILP=4 timeout -k 10 200 python -m pytest tests/test_gpu_pileup_counts.py -m gpu -q --timeout=90 2>&1 | tail -3 | tee -a gpurun_out/${R}_ab.log
