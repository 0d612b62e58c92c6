#!/bin/bash
# same-box A/B of the convolution kernels: round-2 baseline pconv (libclair3b200_v1.so) vs the current one, without / with CTA pairs
set -u
mkdir -p gpurun_out
R=${1:-r2ab}
cp clair3_b200/libclair3b200.so /tmp/lib_v2.so
for cfg in "v1 0" "v2 0" "v2 1" "v1 0" "v2 1"; do
  set -- $cfg
  if [ "$1" = "v1" ]; then cp clair3_b200/libclair3b200_v1.so clair3_b200/libclair3b200.so; else cp /tmp/lib_v2.so clair3_b200/libclair3b200.so; fi
  C3B_PCONV_PAIR=$2 timeout -k 10 300 python bench.py --workloads fa --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_$1_pair$2.json 2> gpurun_out/${R}_$1_pair$2.err
  echo "lib=$1 pair=$2 rc=$? :: $(tail -c 200 gpurun_out/${R}_$1_pair$2.err | tr '\n' ' ')"
done
cp /tmp/lib_v2.so clair3_b200/libclair3b200.so
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=90 -k "tensor_core_kernels_match_reference or conv_taps" 2>&1 | tail -2
