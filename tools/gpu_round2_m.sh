#!/bin/bash
# GPU pass M: chunked tickets + hand-over through shared memory: parity, then sweeps
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_batch or multi_attribute or shell or golden or full_size or qp_structures or dups" --timeout=200 --timeout-method=thread > gpurun_out/m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -5 gpurun_out/m_pytest.log
GANG_SWEEP="1:1,1:1:CHUNKED=0,32:1,128:4,128:4:CHUNKED=0,128:4:GANG_CTAS=1,128:4:GANG_CTAS=2" timeout -k 10 500 python tools/gang_sweep.py > gpurun_out/m_sweep_textured.log 2>&1
cat gpurun_out/m_sweep_textured.log | tail -12
GANG_SWEEP="1:1,1:1:CHUNKED=0,128:4,128:4:CHUNKED=0,128:4:GANG_CTAS=2" timeout -k 10 300 python tools/gang_sweep.py 0 0 4 > gpurun_out/m_sweep_smooth.log 2>&1
cat gpurun_out/m_sweep_smooth.log | tail -8
