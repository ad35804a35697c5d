#!/bin/bash
# GPU pass R: the whole GPU test suite, smoke, the headline bench (N=1) and its reference arm
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout=300 --timeout-method=thread > gpurun_out/r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r_pytest.log
tail -4 gpurun_out/r_pytest.log
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r_smoke.log 2>&1; tail -2 gpurun_out/r_smoke.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r_bench.err; cut -c1-1500 gpurun_out/r_bench.json
timeout -k 10 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r_bench_ref.json 2> gpurun_out/r_bench_ref.err
echo "ref rc=$?"; cut -c1-600 gpurun_out/r_bench_ref.json
