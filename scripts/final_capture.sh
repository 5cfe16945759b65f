#!/bin/bash
# End-of-round evidence at N = 1: bench line, ncu launch list, ncu --set full of the inner-BnB / ICP kernels (default bunny run, and the dense
# inner-BnB shape on the spanner pair).  The reports are exported to raw CSV on the box (gpurun_out/ is capped at 64 MiB).   TAG=r2z bash scripts/final_capture.sh
T=${TAG:-r2z}
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo bench rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_raw.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${T}_ncu_list.log 2>&1; echo list rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"inner_bnb|icp_kernel|expand_bounds|gather_peak" -c 16 -f -o /tmp/${T}_full python scripts/profile_target.py > gpurun_out/${T}_ncu_full.log 2>&1; echo full rc=$?
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_full_raw.csv 2>/dev/null
PROFILE_GOLDEN=spanner_s0.02_mse3e-4 GOICP_BNB_VARIANT=q5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:inner_bnb -c 14 -f -o /tmp/${T}_full_dense python scripts/profile_target.py > gpurun_out/${T}_ncu_full_dense.log 2>&1; echo dense rc=$?
ncu -i /tmp/${T}_full_dense.ncu-rep --page raw --csv > gpurun_out/${T}_full_dense_raw.csv 2>/dev/null
if [ -n "$WITH_TESTS" ]; then (time timeout 800 python -m pytest tests -q -m gpu) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log; fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
ls -la gpurun_out/${T}_*; du -sh gpurun_out
