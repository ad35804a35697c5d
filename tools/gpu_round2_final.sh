#!/bin/bash
# final: the headline bench on the final tree
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/final_bench.err | cut -c1-300; python -c "
import json;d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1]);print(d['value'], d['e2e']['value'], d['decoder'], d['parity_checked'], d['recolouring'].get('ms'))"
