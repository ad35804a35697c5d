#!/bin/bash
# GPU pass Z: whole GPU suite + smoke on the final tree; recolour timing
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout=600 --timeout-method=thread > gpurun_out/z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z_pytest.log
tail -3 gpurun_out/z_pytest.log
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; tail -1 gpurun_out/z_smoke.log
GANG_SWEEP="160:10" timeout -k 10 300 python tools/gang_sweep.py > gpurun_out/z_sweep.log 2>&1; tail -1 gpurun_out/z_sweep.log
