#!/bin/bash
# GPU pass I: hand-over through shared memory, second attempt: parity + A/B
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout -k 10 600 python -m pytest tests -m gpu -x -q --timeout=100 --timeout-method=thread > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_pytest.log
tail -3 gpurun_out/i_pytest.log | cut -c1-200
QP="python tools/quick_profile.py 1000000 lidar"
{
  for tex in "0 0" "16 24" "32 32"; do
    for ho in 1 0; do
      echo "=== texture $tex handover $ho"
      PCCB200_HANDOVER=$ho QP_FULL=0 timeout -k 10 60 $QP $tex 2>&1 | grep -E "enc |dec multi" | cut -c1-200
    done
  done
} > gpurun_out/i_profile.log 2>&1
cat gpurun_out/i_profile.log
