// ref_shim_liftenc.cpp — TEST INFRASTRUCTURE ONLY (oracle).
//
// Drives the reference's own lifting-transform ENCODER bodies
//   AttributeEncoder::encodeColorsLift        tmc3/AttributeEncoder.cpp:1379-1494
//   AttributeEncoder::encodeReflectancesLift  tmc3/AttributeEncoder.cpp:1543-1648
// which are protected members using the translation-unit-local
// PCCResidualsEncoder.  To reach them without copying or modifying anything,
// this TU #includes the reference's AttributeEncoder.cpp (from where it lies)
// with `protected` / `private` opened.  The arithmetic-coded payload it
// produces is decoded again by ref_shim_liftdec.cpp to recover the quantised
// coefficient values.
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
#include "AttributeEncoder.cpp"
#undef protected
#undef private

#include <chrono>
#include <cstring>

#include "pcc_attr_b200.h"

using namespace pcc;

void tmc13ref_fill_aps(const pccb200_lod_params* lp, AttributeParameterSet& aps);

extern "C" int tmc13ref_lift_decode_values(
  const uint8_t* buf, int len, int n, int numAttrs, int32_t* valuesOut);

extern "C" double
tmc13ref_lift_encode(
  const pccb200_lod_params* lp,
  const pccb200_qpset* qs,
  int lcpEnabled,
  const int32_t* xyz,
  const int32_t* attrs,   // n x numAttrs, input order
  int n,
  int numAttrs,
  int bitdepth,
  int32_t* valuesOut,     // n x numAttrs, predictor order
  int32_t* reconOut,      // n x numAttrs, input order
  int8_t* lcpOut)         // num_detail_levels entries (colour only)
{
  AttributeParameterSet aps{};
  tmc13ref_fill_aps(lp, aps);
  aps.attr_encoding = AttributeEncoding::kLiftingTransform;
  aps.last_component_prediction_enabled_flag = lcpEnabled != 0;
  aps.max_num_direct_predictors = 0;
  aps.direct_avg_predictor_disabled_flag = false;
  AttributeBrickHeader abh{};
  abh.attr_dist2_delta = 0;
  AttributeDescription desc{};
  desc.bitdepth = bitdepth;
  desc.attr_num_dimensions_minus1 = numAttrs - 1;
  SequenceParameterSet sps{};

  QpSet qpSet;
  for (int i = 0; i < qs->num_layers; i++)
    qpSet.layers.push_back(Qps{qs->layers[i][0], qs->layers[i][1]});
  qpSet.maxQp = qs->max_qp;
  qpSet.fixedPointQpOffset = qs->fixed_point_qp_offset;

  PCCPointSet3 cloud;
  cloud.addRemoveAttributes(numAttrs == 3, numAttrs == 1);
  cloud.resize(n);
  for (int i = 0; i < n; i++) {
    cloud[i] = point_t{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    if (numAttrs == 3)
      cloud.setColor(i, Vec3<attr_t>(attrs[3 * i], attrs[3 * i + 1], attrs[3 * i + 2]));
    else
      cloud.setReflectance(i, attr_t(attrs[i]));
  }
  AttributeInterPredParams ip;
  ip.frameDistance = 1;
  ip.enableAttrInterPred = false;
  ip.attrInterIntraSliceRDO = false;

  auto t0 = std::chrono::steady_clock::now();
  AttributeEncoder enc;
  enc._abh = &abh;
  enc._lods.generate(aps, abh, n - 1, 0, cloud, ip);
  AttributeContexts ctxtMem;
  ctxtMem.reset();
  PCCResidualsEncoder encoder(aps, abh, ctxtMem);
  encoder.start(sps, n);
  if (numAttrs == 3)
    enc.encodeColorsLift(desc, aps, qpSet, cloud, encoder);
  else
    enc.encodeReflectancesLift(desc, aps, qpSet, cloud, encoder, ip);
  int len = encoder.stop();
  auto t1 = std::chrono::steady_clock::now();

  tmc13ref_lift_decode_values(
    reinterpret_cast<const uint8_t*>(encoder.arithmeticEncoder.buffer()), len, n, numAttrs,
    valuesOut);
  for (int i = 0; i < n; i++) {
    if (numAttrs == 3) {
      auto c = cloud.getColor(i);
      for (int k = 0; k < 3; k++)
        reconOut[3 * i + k] = c[k];
    } else {
      reconOut[i] = cloud.getReflectance(i);
    }
  }
  if (lcpOut && numAttrs == 3)
    for (int l = 0; l < lp->num_detail_levels; l++)
      lcpOut[l] = l < int(abh.attrLcpCoeffs.size()) ? abh.attrLcpCoeffs[l] : 0;
  return std::chrono::duration<double>(t1 - t0).count();
}
