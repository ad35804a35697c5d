#!/bin/bash
# GPU pass O: one 64-bit state word per block (no release fences), binary-search RDOQ code
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_batch or multi_attribute or shell or golden or full_size or qp_structures or dups or fuzz" --timeout=200 --timeout-method=thread > gpurun_out/o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/o_pytest.log
tail -5 gpurun_out/o_pytest.log
GANG_SWEEP="1:1,32:1,128:4" timeout -k 10 400 python tools/gang_sweep.py > gpurun_out/o_sweep_textured.log 2>&1
cat gpurun_out/o_sweep_textured.log | tail -6
GANG_SWEEP="1:1,128:4" timeout -k 10 300 python tools/gang_sweep.py 0 0 4 > gpurun_out/o_sweep_smooth.log 2>&1
cat gpurun_out/o_sweep_smooth.log | tail -4
GANG_SWEEP="1:1,128:4" timeout -k 10 300 python tools/gang_sweep.py 32 32 4 > gpurun_out/o_sweep_tex32.log 2>&1
cat gpurun_out/o_sweep_tex32.log | tail -4
