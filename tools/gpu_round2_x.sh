#!/bin/bash
# GPU pass X: ncu launch lists of the lifting path and of recolouring; the whole GPU suite; the headline bench
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/x_launches_lift.csv python tools/ncu_lift_recolour.py lift > gpurun_out/x_ncu_lift.log 2>&1
tail -2 gpurun_out/x_ncu_lift.log
timeout -k 10 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/x_launches_recolour.csv python tools/ncu_lift_recolour.py recolour > gpurun_out/x_ncu_recolour.log 2>&1
tail -2 gpurun_out/x_ncu_recolour.log
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout=600 --timeout-method=thread > gpurun_out/x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/x_pytest.log
tail -3 gpurun_out/x_pytest.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/x_bench.err | cut -c1-300; cut -c1-300 gpurun_out/x_bench.json
