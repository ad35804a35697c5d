#!/bin/bash
# developer tool: throughput against the number of frames (independent chains) in flight
for f in ${FRAMES:-4 8 16}; do
  timeout 300 python bench.py --frames $f --steps 3 --warmup 3 --no-cpu-baseline > /tmp/b.json 2>/tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('frames', $f, 'value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'single', round(d['single_frame']['ms'],1))" || tail -3 /tmp/b.err
done
