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

// The attribute-independent part of a lifting call: levels of detail,
// predictors, quantisation weights (executor memory).  The reference keeps
// it across the attributes of a slice (AttributeEncoder::_lods,
// tmc3/AttributeEncoder.h:183; reuse rule AttributeLods::isReusable,
// tmc3/AttributeCommon.cpp:76-140).
struct LodState {
  int n;
  int numDetailLevels;
  pccb200_predictor* preds;  // n, predictor order
  uint32_t* idx;             // n, predictor order -> point index
  uint64_t* qw;              // n, quantisation weights
  uint32_t npl[PCCB200_MAX_LODS];
  int lodCount;
};

// fills st (preds / idx / qw must point to n entries each)
template<class Exec>
int
lod_state_build(Exec& ex, const pccb200_lod_params& lod, const int32_t* xyz, int n, LodState& st,
                bool withWeights = true)
{
  st.n = n;
  st.numDetailLevels = lod.num_detail_levels;
  st.lodCount = 0;
  int rc = lod_run(ex, lod, xyz, n, st.preds, st.idx, st.npl, &st.lodCount);
  if (rc != PCCB200_OK)
    return rc;
  // (the quantisation weights belong to the lifting transform; a caller that
  // may only need the predictors -- the predicting transform, whose coding loop
  // stays on the host -- asks for them later)
  if (!withWeights)
    return PCCB200_OK;
  return run_quant_weights(ex, st.preds, n, st.npl, st.lodCount, st.qw);
}

// One attribute on prepared levels of detail.  attrsIn [n*A] (encoder),
// qpoIn [n*2] or null: point order, executor memory.  values [n*A]: coding
// order (out when forward).  attrsOut [n*A]: reconstruction in point order.
// lcp: host array of PCCB200_MAX_LODS + 1 entries.
template<class Exec>
int
attr_lift_on_lods(Exec& ex, bool forward, const LodState& st, const pccb200_qpset& qpset,
                  bool lcpEnabled, const int32_t* qpoIn, const int32_t* attrsIn,
                  int32_t* attrsOut, int A, int bitdepth, int32_t* values, int8_t* lcp)
{
  const int n = st.n;
  int64_t* coef = ex.template alloc<int64_t>(size_t(n) * A);
  int32_t* qpo = nullptr;
  if (qpoIn) {
    qpo = ex.template alloc<int32_t>(size_t(n) * 2);
    ex.foreach(n, GatherQpoFn{qpoIn, st.idx, qpo});
  }
  int rc;
  if (forward) {
    ex.foreach(n, GatherAttrShiftFn{attrsIn, st.idx, A, coef});
    rc = run_lift(ex, true, st.preds, st.qw, n, st.npl, st.lodCount, coef, A);
    if (rc != PCCB200_OK)
      return rc;
  }
  rc = run_lift_quant(ex, forward, qpset, qpo, st.qw, n, st.npl, st.lodCount, st.numDetailLevels,
                      coef, A, lcpEnabled, lcp, values);
  if (rc != PCCB200_OK)
    return rc;
  rc = run_lift(ex, false, st.preds, st.qw, n, st.npl, st.lodCount, coef, A);
  if (rc != PCCB200_OK)
    return rc;
  ex.foreach(n, ScatterReconFn{coef, st.idx, A, (1 << bitdepth) - 1, attrsOut});
  return PCCB200_OK;
}

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
  LodState st;
  st.preds = ex.template alloc<pccb200_predictor>(n);
  st.idx = ex.template alloc<uint32_t>(n);
  st.qw = ex.template alloc<uint64_t>(n);
  int rc = lod_state_build(ex, lod, xyz, n, st);
  if (rc != PCCB200_OK)
    return rc;
  return attr_lift_on_lods(ex, forward, st, qpset, lcpEnabled, qpoIn, attrsIn, attrsOut, A,
                           bitdepth, values, lcp);
}

}  // namespace pccb200
