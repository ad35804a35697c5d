#!/bin/bash
# GPU pass T: tickets one block ahead; lazy quantisation weights on the LoD handle (predlift3m)
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_batch or multi_attribute or golden or shell or lod or lift or whole_codec" --timeout=300 --timeout-method=thread > gpurun_out/t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_pytest.log
tail -4 gpurun_out/t_pytest.log
GANG_SWEEP="1:1,160:10" timeout -k 10 400 python tools/gang_sweep.py > gpurun_out/t_sweep_textured.log 2>&1
cat gpurun_out/t_sweep_textured.log | tail -3
GANG_SWEEP="1:1,160:10" timeout -k 10 300 python tools/gang_sweep.py 0 0 4 > gpurun_out/t_sweep_smooth.log 2>&1
cat gpurun_out/t_sweep_smooth.log | tail -3
timeout -k 10 600 python bench.py --workload predlift3m --steps 3 --warmup 3 > gpurun_out/t_predlift3m.json 2> gpurun_out/t_predlift3m.err
echo "predlift3m rc=$?"; tail -2 gpurun_out/t_predlift3m.err; cut -c1-300 gpurun_out/t_predlift3m.json
