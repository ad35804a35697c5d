#!/bin/bash
# developer tool (run under gpurun): compute-sanitizer over small RAHT encodes /
# decodes (single attribute, two attributes in one pass) and a lifting encode
# decodes (single attribute, two attributes in one pass, a gang of units in one
# batch call) and a lifting encode with distance subsampling (k_subsample_cells).
# racecheck looks at shared memory only (the block kernels hand values over
# through global memory with relaxed / acquire-release accesses).
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
units = [cloud_shell(3000 + 700 * u, bits=6 + u % 2, seed=20 + u) for u in range(5)]
uat = [[texture(a, 20, 30 + u), texture(a[:, :1].copy(), 12, 40 + u)] for u, (x, a) in enumerate(units)]
os.environ["PCCB200_GANG"] = "3"
brec, bcoef = pb.attr_raht_encode_multi_batch(p, [q, q], [x for x, _ in units], uat)
for u, (x, _) in enumerate(units):
    r1, c1 = pb.attr_raht_encode_multi(p, [q, q], x, uat[u])
    assert all(np.array_equal(bcoef[u][s], c1[s]) and np.array_equal(brec[u][s], r1[s]) for s in range(2))
bdec = pb.attr_raht_decode_multi_batch(p, [q, q], [x for x, _ in units], bcoef)
assert all(np.array_equal(bdec[u][s], brec[u][s]) for u in range(5) for s in range(2))
lp = pb.LodParams.from_buffer_copy(bytes(make_lod_params(levels=6)))
lq = pb.QpSet.from_buffer_copy(bytes(make_qpset(qp=30, fixed_point_qp_offset=24)))
vals, lrec, lcp = pb.attr_lift_encode(lp, lq, xyz[:4000], attrs[:4000], lcp_enabled=1)
ldec = pb.attr_lift_decode(lp, lq, xyz[:4000], vals, lcp=lcp)
assert np.array_equal(ldec, lrec)
print("sanitizer run: results exact")
PY
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py 2>&1 | tail -8
  echo "exit code: $?"
done
