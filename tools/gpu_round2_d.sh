#!/bin/bash
# GPU pass D: dual-stream look-back with prefetch: parity, profiles, frames sweep, ncu
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 700 python -m pytest tests -m gpu -x -q --timeout=300 --timeout-method=thread > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -4 gpurun_out/d_pytest.log
QP="python tools/quick_profile.py 1000000 lidar"
{
  for tex in "0 0" "16 24" "32 32"; do
    echo "=== texture $tex"
    QP_FULL=0 timeout -k 10 120 $QP $tex
  done
} > gpurun_out/d_profile.log 2>&1
grep -E "===|enc |dec " gpurun_out/d_profile.log | cut -c1-175
for f in 16 32; do
  timeout -k 10 300 python bench.py --steps 3 --warmup 3 --frames $f --no-lifting --no-cpu-baseline > gpurun_out/d_bench_f$f.json 2> gpurun_out/d_bench_f$f.err
  echo "bench frames=$f rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/d_bench_f$f.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'single',d['single_frame']['ms'],'smooth',d['smooth_frame'] and (d['smooth_frame']['value'], d['smooth_frame']['single_frame_ms']),'dec',d['decoder'])
PY
done
# ncu: launch list of two fused textured encodes, then a full capture of the three largest stages
timeout -k 10 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
  --log-file gpurun_out/d_launches_textured.csv python tools/one_frame.py 1000000 2 textured > gpurun_out/d_ncu_list.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
  --log-file gpurun_out/d_launches_smooth.csv python tools/one_frame.py 1000000 2 smooth > gpurun_out/d_ncu_list2.log 2>&1
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:k_block_warp -s 28 -c 3 \
  -o gpurun_out/d_block_warp_textured -f python tools/one_frame.py 1000000 2 textured > gpurun_out/d_ncu_full.log 2>&1
tail -2 gpurun_out/d_ncu_full.log
ls -la gpurun_out | grep " d_"
