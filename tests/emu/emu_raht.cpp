// emu_raht.cpp — TEST INFRASTRUCTURE ONLY: the RAHT kernel bodies of the
// product (raht_core.cuh / raht_pipeline.cuh) compiled for the host and run
// as loops (see exec_host.h).  Built by tests/emu/Makefile into libemu.so.
#include "exec_host.h"
#include "lift_pipeline.cuh"
#include "lod_pipeline.cuh"
#include "spherical.cuh"
#include "symbols.cuh"
#include "dist2.cuh"
#include "raht_pipeline.cuh"
#include "recolour.cuh"

extern "C" int
emu_raht(int forward, const pccb200_raht_params* pp, const pccb200_qpset* qs,
         const int32_t* qpo, const int64_t* morton, int32_t* attrs, int A,
         int N, int32_t* coeffs)
{
  HostExec ex;
  return pccb200::raht_run(ex, *pp, *qs, forward != 0, morton, attrs, qpo,
                           coeffs, int64_t(N), A, N);
}

extern "C" uint32_t emu_isqrt(uint64_t x) { return pccb200::isqrt64(x); }
extern "C" uint64_t emu_irsqrt(uint64_t x) { return pccb200::irsqrt64(x); }
extern "C" int64_t emu_morton_addr(int32_t x, int32_t y, int32_t z) { return pccb200::morton_addr(x, y, z); }
extern "C" uint64_t emu_morton3d_add(uint64_t a, uint64_t b) { return pccb200::morton3d_add(a, b); }
extern "C" int64_t emu_quantize(int qp, int64_t x) { return pccb200::make_quantizer(qp).quantize(x); }
extern "C" int64_t emu_scale(int qp, int64_t x) { return pccb200::make_quantizer(qp).scale(x); }
extern "C" int64_t emu_fixed_mul(int64_t a, int64_t b) { return pccb200::fx_mul(a, b); }
extern "C" int64_t emu_div_approx(int64_t a, uint64_t b, int32_t s) { return pccb200::div_approx(a, b, s); }

extern "C" int
emu_lod_build(const pccb200_lod_params* lp, const int32_t* xyz, int n, pccb200_predictor* preds,
              uint32_t* indexes, uint32_t* npl, int32_t* lodCount)
{
  HostExec ex;
  int cnt = 0;
  int rc = pccb200::lod_run(ex, *lp, xyz, n, preds, indexes, npl, &cnt);
  *lodCount = cnt;
  return rc;
}

extern "C" int
emu_attr_lift(int forward, const pccb200_lod_params* lod, const pccb200_qpset* qs, int lcpEnabled,
              const int32_t* qpo, const int32_t* xyz, int32_t* attrs, int A, int n, int bitdepth,
              int32_t* values, int8_t* lcp)
{
  HostExec ex;
  int8_t lcpLocal[PCCB200_MAX_LODS + 1] = {};
  if (!forward && lcp)
    for (int l = 0; l < lod->num_detail_levels; l++)
      lcpLocal[l] = lcp[l];
  std::vector<int32_t> out(size_t(n) * A);
  int rc = pccb200::attr_lift_run(ex, forward != 0, *lod, *qs, lcpEnabled != 0, qpo, xyz, attrs,
                                  out.data(), A, n, bitdepth, values, lcpLocal);
  if (rc)
    return rc;
  std::copy(out.begin(), out.end(), attrs);
  if (forward && lcp)
    for (int l = 0; l < lod->num_detail_levels; l++)
      lcp[l] = lcpLocal[l];
  return 0;
}

// spherical.cuh (host build): conversion + optional offsetAndScale
extern "C" int
emu_xyz_to_rpl(const int32_t* origin, const int32_t* theta, int numTheta, const int32_t* weight,
               const int32_t* minPos, const int32_t* xyz, int64_t n, int32_t* out, int32_t* bbox)
{
  HostExec ex;
  pccb200::run_xyz_to_rpl(ex, origin, theta, numTheta, xyz, n, out, bbox, minPos, weight);
  return 0;
}
extern "C" int emu_iatan2(int y, int x) { return pccb200::iatan2_q20(y, x); }

// symbols.cuh (host build)
extern "C" int
emu_coeff_symbols(const int32_t* coeffs, int A, int n, int32_t* runs, int32_t* values, uint8_t* ctx,
                  int32_t* tail)
{
  HostExec ex;
  int count = 0, t = 0;
  pccb200::run_coeff_symbols(ex, coeffs, n, A, n, runs, values, ctx, &count, &t);
  *tail = t;
  return count;
}

// dist2.cuh (host build)
extern "C" int
emu_estimate_dist2(const int32_t* xyz, int n, int period, int range, float percentile)
{
  HostExec ex;
  return pccb200::run_estimate_dist2(ex, xyz, n, period, range, percentile);
}

// the other quantisation-weight derivations of lifting.cuh (host build)
extern "C" int
emu_quant_weights_fixed(const pccb200_predictor* preds, int n, const uint32_t* npl, int lodCount,
                        const int32_t* neighWeight, uint64_t* qw)
{
  HostExec ex;
  return pccb200::run_quant_weights(ex, preds, n, npl, lodCount, qw, neighWeight);
}
extern "C" int
emu_quant_weights_scalable(const uint32_t* npl, int lodCount, uint64_t numPoints, int minLog2, int n,
                           uint64_t* qw)
{
  HostExec ex;
  return pccb200::run_quant_weights_scalable(ex, npl, lodCount, numPoints, minLog2, n, qw);
}


// recolour.cuh (host build)
extern "C" int
emu_recolour(const pccb200_recolour_params* rp, const int32_t* srcXyz, const int32_t* srcAttr, int A,
             int nSrc, double scale, const int32_t* off, const int32_t* tgtXyz, int nTgt,
             int bitdepth, int32_t* out)
{
  HostExec ex;
  return pccb200::recolour_run(ex, *rp, srcXyz, srcAttr, A, nSrc, scale, off, tgtXyz, nTgt, bitdepth,
                               out);
}
