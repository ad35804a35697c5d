#!/bin/bash
# last pass: whole GPU suite and the headline bench on the final tree
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 400 python -m pytest tests -m gpu -x -q --timeout=300 --timeout-method=thread > gpurun_out/last_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/last_pytest.log
tail -3 gpurun_out/last_pytest.log
timeout -k 10 300 python bench.py --steps 3 --warmup 3 > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/last_bench.json').read().strip().splitlines()[-1]);print(d['value'], d['e2e']['value'], d['parity_checked'], d['lifting_path'])"
