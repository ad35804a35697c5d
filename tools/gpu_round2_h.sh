#!/bin/bash
# GPU pass H: hand-over through shared memory (batched tickets): parity + A/B
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 700 python -m pytest tests -m gpu -x -q --timeout=300 --timeout-method=thread > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -3 gpurun_out/h_pytest.log
QP="python tools/quick_profile.py 1000000 lidar"
{
  for tex in "0 0" "16 24" "32 32"; do
    for ho in 1 0; do
      echo "=== texture $tex handover $ho"
      PCCB200_HANDOVER=$ho QP_FULL=0 timeout -k 10 120 $QP $tex 2>&1 | grep -E "enc |dec multi" | cut -c1-200
    done
  done
} > gpurun_out/h_profile.log 2>&1
cat gpurun_out/h_profile.log
timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-lifting --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/h_bench.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'single',d['single_frame']['ms'],'smooth',d['smooth_frame'] and (d['smooth_frame']['value'], d['smooth_frame']['single_frame_ms']),'dec',d['decoder'])
PY
