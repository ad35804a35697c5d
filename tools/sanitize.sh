#!/bin/bash
# developer tool (run under gpurun): compute-sanitizer memcheck + racecheck on a small encode/decode
cat > /tmp/san.py <<'PY'
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pcc_attr_b200 as pb
from pcc_testlib import cloud_shell, make_params, make_qpset, sort_cloud, oracle_raht
xyz, attrs = cloud_shell(6000, bits=7, seed=3, dups=True)
params, qpset = make_params(), make_qpset(qp=30)
p = pb.RahtParams.from_buffer_copy(bytes(params)); q = pb.QpSet.from_buffer_copy(bytes(qpset))
rec, coef = pb.attr_raht_encode(p, q, xyz, attrs)
dec = pb.attr_raht_decode(p, q, xyz, coef)
mort, a_s, order = sort_cloud(xyz, attrs)
orec, ocoef = oracle_raht(1, params, qpset, mort, a_s)
assert np.array_equal(coef, ocoef) and np.array_equal(dec, rec)
print("sanitizer run: results exact")
PY
for tool in ${TOOLS:-memcheck racecheck}; do
  echo "== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py 2>&1 | tail -6
done
