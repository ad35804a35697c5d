/* lift_oracle.c — TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Sequential plain-C restatement of the lifting transform's weight and
 * predict / update passes:
 *   PCCComputeQuantizationWeights   tmc3/PCCTMC3Common.h:828-854
 *   PCCLiftPredict                  tmc3/PCCTMC3Common.h:716-770
 *   PCCLiftUpdate                   tmc3/PCCTMC3Common.h:774-824
 *   LoD loops                       tmc3/AttributeEncoder.cpp:1408-1415,1476-1482
 * Parity is PINNED against the compiled reference (oracle/_ref) by
 * tests/test_oracle_vs_reference.py::test_live_lifting.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pcc_attr_b200.h"
#include "pcc_arith_oracle.h"

void
oracle_quant_weights(const pccb200_predictor* preds, int n, uint64_t* qw)
{
  for (int i = 0; i < n; i++)
    qw[i] = 1u << 8;
  for (int i = n - 1; i >= 0; i--) {
    const pccb200_predictor* p = &preds[i];
    uint64_t w = qw[i];
    for (uint32_t j = 0; j < p->neighbor_count; j++)
      qw[p->predictor_index[j]] +=
        orc_div_exp2_half_inf_u((uint64_t)p->weight[j] * w, 8);
  }
}

static void
lift_predict(const pccb200_predictor* preds, int start, int end, int direct,
             int64_t* attr, int A)
{
  for (int idx = end - 1; idx >= start; idx--) {
    const pccb200_predictor* p = &preds[idx];
    for (int k = 0; k < A; k++) {
      int64_t acc = 0;
      for (uint32_t j = 0; j < p->neighbor_count; j++)
        acc += (int64_t)p->weight[j] * attr[(size_t)p->predictor_index[j] * A + k];
      acc = orc_div_exp2_half_inf(acc, 8);
      if (direct)
        attr[(size_t)idx * A + k] -= acc;
      else
        attr[(size_t)idx * A + k] += acc;
    }
  }
}

static void
lift_update(const pccb200_predictor* preds, const uint64_t* qw, int start,
            int end, int direct, int64_t* attr, int A)
{
  uint64_t* updW = (uint64_t*)calloc((size_t)start + 1, sizeof(uint64_t));
  uint64_t* upd = (uint64_t*)calloc((size_t)start * A + 1, sizeof(uint64_t));
  for (int idx = end - 1; idx >= start; idx--) {
    const pccb200_predictor* p = &preds[idx];
    for (uint32_t j = 0; j < p->neighbor_count; j++) {
      uint64_t w = orc_div_exp2_half_inf_u((uint64_t)p->weight[j] * qw[idx], 8);
      uint32_t nb = p->predictor_index[j];
      updW[nb] += w;
      for (int k = 0; k < A; k++)
        upd[(size_t)nb * A + k] += w * (uint64_t)attr[(size_t)idx * A + k];
    }
  }
  for (int i = 0; i < start; i++) {
    uint32_t sumW = (uint32_t)updW[i]; /* truncation as in the reference */
    if (!sumW)
      continue;
    for (int k = 0; k < A; k++) {
      int64_t u = orc_div_approx((int64_t)upd[(size_t)i * A + k], sumW, 0);
      if (direct)
        attr[(size_t)i * A + k] += u;
      else
        attr[(size_t)i * A + k] -= u;
    }
  }
  free(updW);
  free(upd);
}

void
oracle_lift(int forward, const pccb200_predictor* preds, const uint64_t* qw,
            int n, const uint32_t* numPointsInLod, int lodCount, int64_t* attr,
            int A)
{
  (void)n;
  if (forward) {
    for (int l = lodCount - 1; l >= 1; l--) {
      int s = (int)numPointsInLod[l - 1], e = (int)numPointsInLod[l];
      lift_predict(preds, s, e, 1, attr, A);
      lift_update(preds, qw, s, e, 1, attr, A);
    }
  } else {
    for (int l = 1; l < lodCount; l++) {
      int s = (int)numPointsInLod[l - 1], e = (int)numPointsInLod[l];
      lift_update(preds, qw, s, e, 0, attr, A);
      lift_predict(preds, s, e, 0, attr, A);
    }
  }
}

/* ---- lifting quantisation and last-component prediction -----------------
 *   per-coefficient loop        tmc3/AttributeEncoder.cpp:1424-1473 (colour),
 *                               :1597-1625 (reflectance); decoder
 *                               tmc3/AttributeDecoder.cpp:711-749, 815-837
 *   computeLastComponentPredictionCoeff   tmc3/AttributeEncoder.cpp:1498-1539
 */
static void
lift_quantizers(const pccb200_qpset* qs, int layer, int off0, int off1, orc_quantizer q[2])
{
  int qp0 = qs->layers[layer][0] + off0;
  if (qp0 < 4) qp0 = 4;
  if (qp0 > qs->max_qp) qp0 = qs->max_qp;
  int qp1 = qs->layers[layer][1] + off1 + qp0;
  if (qp1 < 4) qp1 = 4;
  if (qp1 > qs->max_qp) qp1 = qs->max_qp;
  q[0] = orc_mkquant(qp0 + qs->fixed_point_qp_offset);
  q[1] = orc_mkquant(qp1 + qs->fixed_point_qp_offset);
}

void
oracle_lcp_coeffs(const int64_t* coeffs, int n, const uint32_t* npl, int lodCount,
                  int numDetailLevels, int8_t* out)
{
  int64_t s12 = 0, s11 = 0;
  int lod = 0;
  for (int l = 0; l < numDetailLevels; l++)
    out[l] = 0;
  for (int i = 0; i < n; i++) {
    int32_t mult = (int32_t)(uint32_t)((uint64_t)coeffs[3 * (size_t)i + 1] * (uint64_t)coeffs[3 * (size_t)i + 2]);
    int32_t mult2 = (int32_t)(uint32_t)((uint64_t)coeffs[3 * (size_t)i + 1] * (uint64_t)coeffs[3 * (size_t)i + 1]);
    s12 += mult;
    s11 += mult2;
    if (lod >= lodCount || (uint32_t)i != npl[lod] - 1)
      continue;
    int scale = 0;
    if (s12 && s11) {
      int sign = ((s12 < 0) ^ (s11 < 0)) ? -1 : 1;
      scale = (int)(((s12 << 2) + sign * (s11 >> 1)) / s11);
    }
    s12 = s11 = 0;
    out[lod] = (int8_t)(scale < -8 ? -8 : (scale > 8 ? 8 : scale));
    lod++;
  }
  for (; lod < numDetailLevels; lod++)
    out[lod] = lod ? out[lod - 1] : 0;
}

/* forward == 1: attrs in = lifting coefficients, values out, attrs out =
 * reconstructed coefficients; forward == 0: values in, attrs out. */
void
oracle_lift_quant(int forward, const pccb200_qpset* qs, const int32_t* qpo, const uint64_t* qw,
                  int n, const uint32_t* npl, int lodCount, int64_t* attrs, int A,
                  const int8_t* lcp, int32_t* values)
{
  int quantLayer = 0, lod = 0;
  int lcpCoeff = (lcp && A == 3) ? lcp[0] : 0;
  for (int i = 0; i < n; i++) {
    if (quantLayer < lodCount && (uint32_t)i == npl[quantLayer])
      quantLayer = quantLayer + 1 < qs->num_layers ? quantLayer + 1 : qs->num_layers - 1;
    if (lod < lodCount && (uint32_t)i == npl[lod]) {
      lod++;
      if (lcp && A == 3)
        lcpCoeff = lcp[lod];
    }
    orc_quantizer q[2];
    lift_quantizers(qs, quantLayer, qpo ? qpo[2 * i] : 0, qpo ? qpo[2 * i + 1] : 0, q);
    const int64_t iqw = (int64_t)orc_irsqrt(qw[i]);
    const int64_t qwt = (int64_t)((qw[i] * (uint64_t)iqw + (1ull << 39)) >> 40);
    int64_t* a = &attrs[(size_t)i * A];
    int32_t* v = &values[(size_t)i * A];
    if (A == 1) {
      if (forward)
        v[0] = (int32_t)orc_quantize(q[0], a[0] * qwt);
      a[0] = orc_div_exp2_half_inf(orc_scale(q[0], v[0]) * iqw, 40);
      continue;
    }
    if (forward)
      v[0] = (int32_t)orc_quantize(q[0], a[0] * qwt);
    a[0] = orc_div_exp2_half_inf(orc_scale(q[0], v[0]) * iqw, 40);
    if (forward)
      v[1] = (int32_t)orc_quantize(q[1], a[1] * qwt);
    int64_t scaled = orc_scale(q[1], v[1]);
    a[1] = orc_div_exp2_half_inf(scaled * iqw, 40);
    if (forward)
      a[2] -= (lcpCoeff * a[1]) >> 2;
    scaled *= lcpCoeff;
    scaled >>= 2;
    if (forward)
      v[2] = (int32_t)orc_quantize(q[1], a[2] * qwt);
    scaled += orc_scale(q[1], v[2]);
    a[2] = orc_div_exp2_half_inf(scaled * iqw, 40);
  }
}

/* computeQuantizationWeights, tmc3/PCCTMC3Common.h:895-921 (predicting transform) */
void
oracle_quant_weights_fixed(const pccb200_predictor* preds, int n, const int32_t neigh_weight[3],
                           uint64_t* qw)
{
  for (int i = 0; i < n; i++)
    qw[i] = 1u << 8;
  for (int i = n - 1; i >= 0; i--) {
    const pccb200_predictor* p = &preds[i];
    uint64_t w = qw[i];
    for (uint32_t j = 0; j < p->neighbor_count; j++)
      qw[p->predictor_index[j]] += orc_div_exp2_half_inf_u((uint64_t)(int64_t)neigh_weight[j] * w, 8);
  }
}

/* computeQuantizationWeightsScalable, tmc3/PCCTMC3Common.h:858-891 */
void
oracle_quant_weights_scalable(const uint32_t* npl, int lod_count, uint64_t num_points,
                              int min_geom_node_size_log2, uint64_t* qw)
{
  for (int l = 0; l < lod_count; l++) {
    uint32_t s = l ? npl[l - 1] : 0, e = npl[l];
    uint64_t v = (num_points / npl[l]) << 8;
    for (uint32_t i = s; i < e; i++)
      qw[i] = (!min_geom_node_size_log2 && l == lod_count - 1) ? (uint64_t)1 << 8 : v;
  }
}

