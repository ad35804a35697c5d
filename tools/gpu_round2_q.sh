#!/bin/bash
# GPU pass Q: neighbour values by component row (fewer exchanges); ncu full capture of gang launches
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_batch or multi_attribute or golden or shell or qp_structures" --timeout=200 --timeout-method=thread > gpurun_out/q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log
GANG_SWEEP="1:1,128:4" timeout -k 10 400 python tools/gang_sweep.py > gpurun_out/q_sweep_textured.log 2>&1
cat gpurun_out/q_sweep_textured.log | tail -3
GANG_SWEEP="1:1,128:4" timeout -k 10 300 python tools/gang_sweep.py 0 0 4 > gpurun_out/q_sweep_smooth.log 2>&1
cat gpurun_out/q_sweep_smooth.log | tail -3
export GANG_NOREF=1 GANG_STEPS=0 GANG_SWEEP="16:4"
timeout -k 10 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_block_warp_gang --launch-skip 38 --launch-count 5 -f -o gpurun_out/q_gang_full python tools/gang_sweep.py 16 24 2 > gpurun_out/q_ncu_full.log 2>&1
tail -3 gpurun_out/q_ncu_full.log
ls -la gpurun_out/q_*
