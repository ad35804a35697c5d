#!/bin/bash
# developer tool: effect of the persistent grid size on the chain-bound encoder
for g in 16 37 74 148 296; do
  echo "== PCCB200_BLOCK_GRID=$g"
  PCCB200_BLOCK_GRID=$g timeout 120 python tools/quick_profile.py 1000000 2>&1 | grep -E "enc default|dec default|dec nopred" | head -3
done
