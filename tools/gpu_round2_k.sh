#!/bin/bash
# GPU pass K: gang launches (many units per call): parity, then throughput sweeps
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_batch or multi_attribute or shell" --timeout=150 --timeout-method=thread > gpurun_out/k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/k_pytest.log
tail -5 gpurun_out/k_pytest.log
GANG_SWEEP="32:1,64:2,128:4,128:8,128:16,128:32" timeout -k 10 400 python tools/gang_sweep.py > gpurun_out/k_sweep_textured.log 2>&1
cat gpurun_out/k_sweep_textured.log | tail -12
GANG_SWEEP="32:1,128:4,128:16" timeout -k 10 200 python tools/gang_sweep.py 0 0 4 > gpurun_out/k_sweep_smooth.log 2>&1
cat gpurun_out/k_sweep_smooth.log | tail -8
