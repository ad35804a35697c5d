#!/bin/bash
# GPU pass W (2 GPUs): the headline bench under torchrun, N=2 (weak scaling over frames)
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-lifting > gpurun_out/w_bench_n2.json 2> gpurun_out/w_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/w_bench_n2.err | cut -c1-300; cut -c1-1200 gpurun_out/w_bench_n2.json
