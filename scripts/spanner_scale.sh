#!/bin/bash
# BASELINE config 4 (test/spanner_goicp.toml as written) under torchrun at N GPUs of one node.   bash scripts/spanner_scale.sh N   (TAG=r2w)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
if [ "$N" = 1 ]; then TR=python; fi
timeout 500 $TR bench.py --gpus $N --workload spanner_goicp_toml --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-samples 1 > gpurun_out/${TAG:-r2w}_spanner_n$N.json 2> gpurun_out/${TAG:-r2w}_spanner_n$N.err; echo spanner rc=$?
tail -1 gpurun_out/${TAG:-r2w}_spanner_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['n_gpus','value','ms_per_step','seconds_bnb_kernels_per_step','rounds_per_step','bound_evals_per_step','bound_evals_executed_per_step','rot_pops','trans_pops','sse','exit_path']}); print(d['roofline']['frac'], d['e2e']['seconds_per_step'])"
