#!/bin/bash
# GPU pass F: coefficient / word prefetch: parity, profiles; share of the machine held by the persistent grids
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 700 python -m pytest tests -m gpu -x -q --timeout=300 --timeout-method=thread > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -3 gpurun_out/f_pytest.log
QP="python tools/quick_profile.py 1000000 lidar"
{
  for tex in "0 0" "16 24" "32 32"; do
    echo "=== texture $tex"
    PCCB200_DEBUG=1 QP_FULL=0 timeout -k 10 120 $QP $tex 2>&1 | grep -E "launches|enc |dec " | cut -c1-330
  done
} > gpurun_out/f_profile.log 2>&1
cat gpurun_out/f_profile.log
for sh in 100 70 40; do
  PCCB200_BLOCK_SHARE=$sh timeout -k 10 300 python bench.py --steps 3 --warmup 3 --frames 32 --no-lifting --no-cpu-baseline > gpurun_out/f_bench_s$sh.json 2> gpurun_out/f_bench_s$sh.err
  echo "bench share=$sh rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/f_bench_s$sh.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'single',d['single_frame']['ms'],'smooth',d['smooth_frame'] and (d['smooth_frame']['value'], d['smooth_frame']['single_frame_ms']),'dec',d['decoder'])
PY
done
