#!/bin/bash
# developer tool (run under gpurun): compute-sanitizer over small RAHT encodes /
# decodes (single attribute, two attributes in one pass) and a lifting encode
# with distance subsampling (k_subsample_cells).  racecheck looks at shared
# memory only: it is run with PCCB200_HANDOVER=0 (the shared-memory hand-over of
# k_block_warp is a flag protocol between warps without a barrier, by design).
cat > /tmp/san.py <<'PY'
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pcc_attr_b200 as pb
from pcc_testlib import cloud_shell, make_params, make_qpset, make_lod_params, sort_cloud, oracle_raht
from pcc_attr_b200.synth import texture
xyz, attrs = cloud_shell(6000, bits=7, seed=3, dups=True)
attrs = texture(attrs, 24, 5)
refl = texture(attrs[:, :1].copy(), 16, 6)
params, qpset = make_params(), make_qpset(qp=30)
p = pb.RahtParams.from_buffer_copy(bytes(params)); q = pb.QpSet.from_buffer_copy(bytes(qpset))
rec, coef = pb.attr_raht_encode(p, q, xyz, attrs)
dec = pb.attr_raht_decode(p, q, xyz, coef)
mort, a_s, order = sort_cloud(xyz, attrs)
orec, ocoef = oracle_raht(1, params, qpset, mort, a_s)
assert np.array_equal(coef, ocoef) and np.array_equal(dec, rec)
recs, coefs = pb.attr_raht_encode_multi(p, [q, q], xyz, [attrs, refl])
assert np.array_equal(coefs[0], ocoef)
decs = pb.attr_raht_decode_multi(p, [q, q], xyz, coefs)
assert np.array_equal(decs[0], recs[0]) and np.array_equal(decs[1], recs[1])
lp = pb.LodParams.from_buffer_copy(bytes(make_lod_params(levels=6)))
lq = pb.QpSet.from_buffer_copy(bytes(make_qpset(qp=30, fixed_point_qp_offset=24)))
vals, lrec, lcp = pb.attr_lift_encode(lp, lq, xyz[:4000], attrs[:4000], lcp_enabled=1)
ldec = pb.attr_lift_decode(lp, lq, xyz[:4000], vals, lcp=lcp)
assert np.array_equal(ldec, lrec)
print("sanitizer run: results exact")
PY
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  echo "== compute-sanitizer --tool $tool"
  if [ $tool = racecheck ]; then export PCCB200_HANDOVER=0; else unset PCCB200_HANDOVER; fi
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py 2>&1 | tail -8
  echo "exit code: $?"
done
