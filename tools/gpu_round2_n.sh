#!/bin/bash
# GPU pass N: chain kernel (one CTA per unit, ring of slots in shared memory): parity, sweeps
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_batch or multi_attribute" --timeout=200 --timeout-method=thread > gpurun_out/n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n_pytest.log
tail -5 gpurun_out/n_pytest.log
GANG_SWEEP="128:4:CHAIN=0,128:4,128:4:CHAIN=16,128:4:CHAIN=8,128:8:CHAIN=24,128:16:CHAIN=24,148:4:CHAIN=24" timeout -k 10 500 python tools/gang_sweep.py > gpurun_out/n_sweep_textured.log 2>&1
cat gpurun_out/n_sweep_textured.log | tail -12
GANG_SWEEP="128:4:CHAIN=0,128:4,128:4:CHAIN=16" timeout -k 10 300 python tools/gang_sweep.py 0 0 4 > gpurun_out/n_sweep_smooth.log 2>&1
cat gpurun_out/n_sweep_smooth.log | tail -8
