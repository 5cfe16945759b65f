#!/bin/bash
# bench lines under torchrun at N GPUs of one node: bunny config, sweep_100k strict, sweep_100k tolerance numerics.   bash scripts/multi_gpu_bench.sh N   (TAG=r2s)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG:-r2s}_bench_n$N.json 2> gpurun_out/${TAG:-r2s}_bench_n$N.err; echo bench rc=$?
timeout 400 $TR bench.py --gpus $N --workload sweep_100k --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG:-r2s}_sweep100k_n$N.json 2> gpurun_out/${TAG:-r2s}_sweep100k_n$N.err; echo sweep rc=$?
timeout 400 $TR bench.py --gpus $N --workload sweep_100k --numerics 3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG:-r2s}_sweep100k_fast_n$N.json 2> gpurun_out/${TAG:-r2s}_sweep100k_fast_n$N.err; echo sweepfast rc=$?
for f in bench sweep100k sweep100k_fast; do tail -1 gpurun_out/${TAG:-r2s}_${f}_n$N.json | cut -c1-230; done
