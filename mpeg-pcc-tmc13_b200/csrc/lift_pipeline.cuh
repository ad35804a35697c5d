// lift_pipeline.cuh — the lifting attribute coder without its entropy coding
// (AttributeEncoder::encode{Colors,Reflectances}Lift,
// tmc3/AttributeEncoder.cpp:1379-1494,1543-1648, and the decoder's
// decode{Colors,Reflectances}Lift, tmc3/AttributeDecoder.cpp:678-857) as one
// executor-generic schedule: LoD build -> quantisation weights -> forward
// lifting -> last-component prediction + quantisation -> inverse lifting ->
// rounding, clip and write-back in point order.
#pragma once

#include "lifting.cuh"
#include "lod_pipeline.cuh"

namespace pccb200 {

// xyz [n*3], attrsIn [n*A] (encoder), qpoIn [n*2] or null: point order,
// executor memory.  values [n*A]: coding order (out when forward).
// attrsOut [n*A]: reconstruction in point order.  lcp: host array of
// PCCB200_MAX_LODS + 1 entries.
template<class Exec>
int
attr_lift_run(Exec& ex, bool forward, const pccb200_lod_params& lod, const pccb200_qpset& qpset,
              bool lcpEnabled, const int32_t* qpoIn, const int32_t* xyz, const int32_t* attrsIn,
              int32_t* attrsOut, int A, int n, int bitdepth, int32_t* values, int8_t* lcp)
{
  pccb200_predictor* preds = ex.template alloc<pccb200_predictor>(n);
  uint32_t* idx = ex.template alloc<uint32_t>(n);
  uint32_t npl[PCCB200_MAX_LODS];
  int lodCount = 0;
  int rc = lod_run(ex, lod, xyz, n, preds, idx, npl, &lodCount);
  if (rc != PCCB200_OK)
    return rc;
  uint64_t* qw = ex.template alloc<uint64_t>(n);
  rc = run_quant_weights(ex, preds, n, npl, lodCount, qw);
  if (rc != PCCB200_OK)
    return rc;
  int64_t* coef = ex.template alloc<int64_t>(size_t(n) * A);
  int32_t* qpo = nullptr;
  if (qpoIn) {
    qpo = ex.template alloc<int32_t>(size_t(n) * 2);
    ex.foreach(n, GatherQpoFn{qpoIn, idx, qpo});
  }
  if (forward) {
    ex.foreach(n, GatherAttrShiftFn{attrsIn, idx, A, coef});
    rc = run_lift(ex, true, preds, qw, n, npl, lodCount, coef, A);
    if (rc != PCCB200_OK)
      return rc;
  }
  rc = run_lift_quant(ex, forward, qpset, qpo, qw, n, npl, lodCount, lod.num_detail_levels, coef,
                      A, lcpEnabled, lcp, values);
  if (rc != PCCB200_OK)
    return rc;
  rc = run_lift(ex, false, preds, qw, n, npl, lodCount, coef, A);
  if (rc != PCCB200_OK)
    return rc;
  ex.foreach(n, ScatterReconFn{coef, idx, A, (1 << bitdepth) - 1, attrsOut});
  return PCCB200_OK;
}

}  // namespace pccb200
