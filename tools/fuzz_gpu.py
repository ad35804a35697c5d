"""Developer tool (run under gpurun): randomised check of the CUDA path, through
the C ABI, against the oracle over the parameter space (same generators as the
CPU fuzzers).  Usage: python tools/fuzz_gpu.py [raht cases] [lod cases] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))
import numpy as np  # noqa: E402
import pcc_attr_b200 as pb  # noqa: E402
from fuzz_cpu import random_case  # noqa: E402
from pcc_testlib import *  # noqa: E402,F401,F403


def raht_cases(cases, rng):
    bad = 0
    for i in range(cases):
        xyz, attrs, pkw, qkw, qpo = random_case(rng)
        params, qs = make_params(**pkw), make_qpset(**qkw)
        p = pb.RahtParams.from_buffer_copy(bytes(params))
        q = pb.QpSet.from_buffer_copy(bytes(qs))
        mort, a_s, order = sort_cloud(xyz, attrs)
        qo = qpo[order] if qpo is not None else None
        orec, ocoef = oracle_raht(1, params, qs, mort, a_s, qpoffs=qo)
        grec, gcoef = pb.raht_forward(p, q, mort, a_s, qpoffs=qo)
        ok = np.array_equal(gcoef, ocoef) and np.array_equal(grec, orec)
        ok = ok and np.array_equal(pb.raht_inverse(p, q, mort, ocoef, qpoffs=qo), orec)
        # attribute level (sort, gather, clip, write back) and the symbol stream
        bd = qkw["bitdepth"]
        rec, coef = pb.attr_raht_encode(p, q, xyz, attrs, bitdepth=bd, qpoffs=qpo)
        out = np.empty_like(orec)
        out[order] = np.clip(orec, 0, (1 << bd) - 1)
        ok = ok and np.array_equal(coef, ocoef) and np.array_equal(rec, out)
        rec2, runs, vals, ctx, tail = pb.attr_raht_encode_symbols(p, q, xyz, attrs, bitdepth=bd, qpoffs=qpo)
        o = oracle_coeff_symbols(ocoef)
        ok = ok and np.array_equal(rec2, out) and np.array_equal(runs, o[0]) and np.array_equal(vals, o[1]) and tail == o[3]
        if not ok:
            bad += 1
            print("MISMATCH raht case", i, "n", len(xyz), "A", attrs.shape[1], pkw, qkw, "qpo", qpo is not None, flush=True)
    return bad


def lod_cases(cases, rng):
    bad = 0
    for i in range(cases):
        kind = rng.integers(0, 3)
        n = int(rng.integers(1, 5000))
        if kind == 0:
            xyz, attrs = cloud_shell(n, bits=int(rng.integers(4, 11)), seed=int(rng.integers(1 << 30)),
                                     dups=bool(rng.integers(0, 2)))
        elif kind == 1:
            xyz, attrs = cloud_lidar(max(n, 50), seed=int(rng.integers(1 << 30)))
        else:
            xyz, attrs = cloud_random(n, int(rng.integers(2, 22)), seed=int(rng.integers(1 << 30)),
                                      dup_frac=float(rng.choice([0.0, 0.2])))
        levels = int(rng.integers(1, 14))
        lifting = bool(rng.integers(0, 2))
        kw = dict(levels=levels, decimation=int(rng.integers(0, 3)), period=int(rng.integers(2, 9)),
                  dist2=int(rng.integers(0, 4)), k=int(rng.integers(1, 4)),
                  inter_range=int(rng.choice([1, 8, 128, 1100000])), distribution=int(rng.integers(0, 2)),
                  bias=tuple(int(x) for x in rng.integers(1, 4, 3)))
        if not lifting:
            kw.update(intra_range=int(rng.choice([0, 4, 128])), skip_layers=int(rng.integers(0, levels + 1)),
                      blending=int(rng.integers(0, 2)))
        lp = make_lod_params(**kw)
        glp = pb.LodParams.from_buffer_copy(bytes(lp))
        op, oi, on = oracle_lod_build(lp, xyz)
        gp, gi, gn = pb.lod_build(glp, xyz)
        ok = np.array_equal(gn, on) and np.array_equal(gi, oi) and np.array_equal(gp, op)
        if ok and lifting and len(xyz) > 1:
            a = int(rng.choice([1, 3]))
            at = attrs[:, :a].copy()
            lcp = int(rng.integers(0, 2))
            qs = make_qpset(qp=int(rng.integers(4, 52)), chroma_offset=int(rng.integers(-4, 5)) if a == 3 else 0,
                            fixed_point_qp_offset=24)
            gq = pb.QpSet.from_buffer_copy(bytes(qs))
            ov, orr, ol = oracle_lift_encode(lp, qs, lcp, xyz, at)
            gv, gr, gl = pb.attr_lift_encode(glp, gq, xyz, at, lcp_enabled=lcp)
            ok = np.array_equal(gv, ov) and np.array_equal(gr, orr) and (not (a == 3 and lcp) or np.array_equal(gl, ol))
            gd = pb.attr_lift_decode(glp, gq, xyz, ov, lcp=ol if (a == 3 and lcp) else None)
            ok = ok and np.array_equal(gd, orr)
        if not ok:
            bad += 1
            print("MISMATCH lod case", i, "n", len(xyz), kw, "lifting", lifting, flush=True)
    return bad


def main():
    nr = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    nl = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    rng = np.random.default_rng(seed)
    b1 = raht_cases(nr, rng)
    b2 = lod_cases(nl, rng)
    print(f"raht {nr} cases, {b1} mismatches; lod/lifting {nl} cases, {b2} mismatches (seed {seed})")
    return 1 if b1 or b2 else 0


if __name__ == "__main__":
    sys.exit(main())
