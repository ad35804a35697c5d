#!/bin/bash
# GPU pass L: gang launches, knob sweep (CTAs per unit, poll interval, machine share)
mkdir -p gpurun_out
cd /root/repo
GANG_SWEEP="128:4,128:4:GANG_CTAS=1,128:4:GANG_CTAS=2,128:4:POLL_NS=200,128:4:POLL_NS=1000,128:4:BLOCK_SHARE=60,128:4:GANG_CTAS=1;BLOCK_SHARE=60,128:4:GANG_CTAS=1;POLL_NS=200" timeout -k 10 500 python tools/gang_sweep.py > gpurun_out/l_sweep_textured.log 2>&1
cat gpurun_out/l_sweep_textured.log | tail -12
GANG_SWEEP="128:4,128:4:GANG_CTAS=1,128:4:GANG_CTAS=2,128:4:BLOCK_SHARE=60,128:4:GANG_CTAS=2;BLOCK_SHARE=60" timeout -k 10 300 python tools/gang_sweep.py 0 0 4 > gpurun_out/l_sweep_smooth.log 2>&1
cat gpurun_out/l_sweep_smooth.log | tail -8
