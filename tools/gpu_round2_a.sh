#!/bin/bash
# first GPU pass of round 2: parity of the wavefront path, then A/B timings
mkdir -p gpurun_out
cd /root/repo; timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; tail -2 gpurun_out/a_smoke.log
timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=240 --timeout-method=thread > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
QP="python tools/quick_profile.py 1000000 lidar"
{
  for tex in "0 0" "16 24" "32 32"; do
    echo "=== texture $tex: legacy (Morton order, per-stage, round-1 kernel)"
    PCCB200_BLOCK_KERNEL=warp QP_FULL=0 timeout -k 10 300 $QP $tex
    echo "=== texture $tex: new kernel, Morton order"
    PCCB200_WAVE_ORDER=morton QP_FULL=0 timeout -k 10 300 $QP $tex
    echo "=== texture $tex: wave default (window 2, patience 64)"
    PCCB200_DEBUG=1 QP_FULL=1 timeout -k 10 300 $QP $tex
    for w in 0 4; do
      echo "=== texture $tex: wave window $w"
      PCCB200_WAVE_WINDOW=$w QP_FULL=0 timeout -k 10 300 $QP $tex
    done
    for pat in 8 512; do
      echo "=== texture $tex: wave patience $pat"
      PCCB200_WAVE_PATIENCE=$pat QP_FULL=0 timeout -k 10 300 $QP $tex
    done
  done
} > gpurun_out/a_profile.log 2>&1
tail -40 gpurun_out/a_profile.log
