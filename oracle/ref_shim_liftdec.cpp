// ref_shim_liftdec.cpp — TEST INFRASTRUCTURE ONLY (oracle).
//
// Decodes the arithmetic-coded residual payload written by the reference's
// lifting encoder with the reference's own translation-unit-local
// PCCResidualsDecoder (tmc3/AttributeDecoder.cpp:53-176), walking it exactly
// as decodeColorsLift / decodeReflectancesLift do
// (tmc3/AttributeDecoder.cpp:711-749, 815-837): run lengths of all-zero
// entries, then the values.  #includes the reference's .cpp from where it lies.
// standard headers first: opening `private` must not reach libstdc++
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <queue>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#define protected public
#define private public
#include "AttributeDecoder.cpp"
#undef protected
#undef private

#include "pcc_attr_b200.h"

using namespace pcc;

// shared with ref_shim_liftenc.cpp / ref_shim.cpp
void
tmc13ref_fill_aps(const pccb200_lod_params* lp, AttributeParameterSet& aps)
{
  aps.lod_decimation_type = LodDecimationMethod(lp->lod_decimation_type);
  aps.canonical_point_order_flag = false;
  aps.max_points_per_sort_log2_plus1 = 0;
  aps.num_pred_nearest_neighbours_minus1 = lp->num_pred_nearest_neighbours - 1;
  aps.num_detail_levels_minus1 = lp->num_detail_levels - 1;
  aps.dist2 = lp->dist2;
  aps.inter_lod_search_range = lp->inter_lod_search_range;
  aps.intra_lod_search_range = lp->intra_lod_search_range;
  aps.intra_lod_prediction_skip_layers = lp->intra_lod_prediction_skip_layers;
  aps.predictionWithDistributionEnabled = lp->prediction_with_distribution != 0;
  aps.lodNeighBias = {lp->lod_neigh_bias[0], lp->lod_neigh_bias[1], lp->lod_neigh_bias[2]};
  aps.pred_weight_blending_enabled_flag = lp->pred_weight_blending != 0;
  aps.scalable_lifting_enabled_flag = false;
  aps.lodSamplingPeriod.assign(
    lp->lod_sampling_period, lp->lod_sampling_period + PCCB200_MAX_LODS);
}

extern "C" int
tmc13ref_lift_decode_values(
  const uint8_t* buf, int len, int n, int numAttrs, int32_t* valuesOut)
{
  AttributeBrickHeader abh{};
  SequenceParameterSet sps{};
  AttributeContexts ctxtMem;
  ctxtMem.reset();
  PCCResidualsDecoder decoder(abh, ctxtMem);
  decoder.start(sps, reinterpret_cast<const char*>(buf), len);
  int zeroRunRem = 0;
  for (int i = 0; i < n; i++) {
    if (--zeroRunRem < 0)
      zeroRunRem = decoder.decodeRunLength();
    int32_t values[3] = {};
    if (!zeroRunRem) {
      if (numAttrs == 3)
        decoder.decode(values);
      else
        values[0] = decoder.decode();
    }
    for (int k = 0; k < numAttrs; k++)
      valuesOut[i * numAttrs + k] = values[k];
  }
  decoder.stop();
  return 0;
}

// The symbol stream as the reference's RAHT decoder reads it
// (tmc3/AttributeDecoder.cpp:553-565 one component, :641-654 three): every
// decodeRunLength() result and every decoded value, in order.
extern "C" int
tmc13ref_decode_symbol_stream(
  const uint8_t* buf, int len, int n, int numAttrs, int32_t* runsOut, int32_t* valuesOut,
  int32_t* tailRunOut)
{
  AttributeBrickHeader abh{};
  SequenceParameterSet sps{};
  AttributeContexts ctxtMem;
  ctxtMem.reset();
  PCCResidualsDecoder decoder(abh, ctxtMem);
  decoder.start(sps, reinterpret_cast<const char*>(buf), len);
  int count = 0;
  int zeroRunRem = 0;
  int lastRun = 0;
  bool pendingRun = false;
  for (int i = 0; i < n; i++) {
    if (--zeroRunRem < 0) {
      zeroRunRem = decoder.decodeRunLength();
      lastRun = zeroRunRem;
      pendingRun = true;
    }
    if (!zeroRunRem) {
      int32_t values[3] = {};
      if (numAttrs == 3)
        decoder.decode(values);
      else
        values[0] = decoder.decode();
      runsOut[count] = lastRun;
      for (int k = 0; k < numAttrs; k++)
        valuesOut[count * numAttrs + k] = values[k];
      count++;
      pendingRun = false;
    }
  }
  *tailRunOut = pendingRun ? lastRun : 0;
  decoder.stop();
  return count;
}
