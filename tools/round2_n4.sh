#!/bin/bash
# Round-2 closing run: the new pileup feature counter (diag + tests), the whole GPU suite, smoke, the driver's bench command,
# ncu launch list + one full capture of the counter.
set -u
mkdir -p gpurun_out
R=${1:-r2n}
timeout -k 10 120 python tools/plp_diag.py diag > gpurun_out/${R}_plp_diag.log 2>&1; tail -25 gpurun_out/${R}_plp_diag.log
timeout -k 10 240 python -m pytest tests/test_gpu_pileup_counts.py -m gpu -q --timeout=90 > gpurun_out/${R}_pytest_plp.log 2>&1; tail -15 gpurun_out/${R}_pytest_plp.log
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout=120 --deselect tests/test_gpu_pileup_counts.py > gpurun_out/${R}_pytest.log 2>&1; tail -6 gpurun_out/${R}_pytest.log
timeout -k 10 150 python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; tail -4 gpurun_out/${R}_smoke.log
timeout -k 10 480 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 600 gpurun_out/${R}_bench.err
python tools/bench_summary.py gpurun_out/${R}_bench.json > gpurun_out/${R}_bench_summary.md 2>&1; head -10 gpurun_out/${R}_bench_summary.md; tail -4 gpurun_out/${R}_bench_summary.md
timeout -k 10 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${R}_launches_plp.csv \
    python tools/plp_diag.py prof > gpurun_out/${R}_ncu_launch_plp.log 2>&1
timeout -k 10 180 ncu --set full --clock-control none --import-source on -k regex:"plp_count_tile" -s 1 -c 1 \
    -o gpurun_out/${R}_prof_plp python tools/plp_diag.py prof > gpurun_out/${R}_ncu_full_plp.log 2>&1
ls -la gpurun_out | grep ${R}_ | head -30
