#!/bin/bash
# GPU pass C: parity (incl. multi-attribute pass), profiles, bench both arms
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 600 python -m pytest tests -m gpu -x -q --timeout=300 --timeout-method=thread > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -4 gpurun_out/c_pytest.log
QP="python tools/quick_profile.py 1000000 lidar"
{
  for tex in "0 0" "16 24" "32 32"; do
    echo "=== texture $tex"
    QP_FULL=0 timeout -k 10 120 $QP $tex
  done
} > gpurun_out/c_profile.log 2>&1
grep -E "===|enc |dec " gpurun_out/c_profile.log | cut -c1-175
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/c_bench.err; head -c 3000 gpurun_out/c_bench.json
timeout -k 10 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c_bench_ref.json 2> gpurun_out/c_bench_ref.err
echo "ref rc=$?"; head -c 1500 gpurun_out/c_bench_ref.json
