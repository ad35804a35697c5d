#!/bin/bash
# GPU pass S: the other BASELINE configurations (configs[2]-[4]) with their reference arms
mkdir -p gpurun_out
cd /root/repo
for w in predlift3m lift10m raht30m; do
  timeout -k 10 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/s_$w.json 2> gpurun_out/s_$w.err
  echo "$w rc=$?"; tail -2 gpurun_out/s_$w.err; cut -c1-400 gpurun_out/s_$w.json
  timeout -k 10 400 python bench.py --workload $w --impl reference --steps 1 --warmup 1 > gpurun_out/s_${w}_ref.json 2> gpurun_out/s_${w}_ref.err
  echo "$w ref rc=$?"; cut -c1-300 gpurun_out/s_${w}_ref.json
done
