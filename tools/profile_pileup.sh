#!/bin/bash
# Pileup-only part of tools/profile_round.sh (run under gpurun): validation, default benches, ncu launch list + full capture.
set -u
mkdir -p gpurun_out
R=${1:-r1}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest.log 2>&1; tail -3 gpurun_out/${R}_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
timeout 400 python bench.py > gpurun_out/${R}_bench_pileup.json 2> gpurun_out/${R}_bench_pileup.err
timeout 400 python bench.py --workload fa --steps 100 > gpurun_out/${R}_bench_fa.json 2> gpurun_out/${R}_bench_fa.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/${R}_launches_pileup.csv \
    python bench.py --steps 8 --warmup 8 --no-cpu-baseline > gpurun_out/${R}_ncu_launch_p.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"lstm_tc_kernel|igemm_kernel|heads_kernel" -s 60 -c 5 \
    -o gpurun_out/${R}_prof_pileup python bench.py --steps 4 --warmup 8 --streams 1 --no-cpu-baseline > gpurun_out/${R}_ncu_full_p.log 2>&1
ls -la gpurun_out | grep ${R}_ | head -30
