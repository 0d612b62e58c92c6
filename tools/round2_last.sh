#!/bin/bash
# last GPU check of round 2: counter tests at the shipped kernel, smoke, the driver's bench command
set -u
mkdir -p gpurun_out
R=${1:-r2r}
timeout -k 10 100 python tools/plp_diag.py prof 1048576 10 2>&1 | tail -1 | tee gpurun_out/${R}_plp_prof.log
timeout -k 10 200 python -m pytest tests/test_gpu_pileup_counts.py tests/test_gpu_parity.py -m gpu -q --timeout=90 > gpurun_out/${R}_pytest.log 2>&1; tail -3 gpurun_out/${R}_pytest.log
timeout -k 10 150 python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; tail -3 gpurun_out/${R}_smoke.log
timeout -k 10 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 600 gpurun_out/${R}_bench.err
python tools/bench_summary.py gpurun_out/${R}_bench.json > gpurun_out/${R}_bench_summary.md 2>&1; head -10 gpurun_out/${R}_bench_summary.md; tail -3 gpurun_out/${R}_bench_summary.md
