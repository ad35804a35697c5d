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
