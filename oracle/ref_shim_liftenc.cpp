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

//============================================================================
// Symbol stream of the RAHT attribute coder (row N1).
//
// tmc13ref_raht_encode_payload runs the reference's own
// encode{Colors,Reflectances}TransformRaht (tmc3/AttributeEncoder.cpp:1215-1377:
// sort, transform, the coefficient walk, write-back) and returns the
// arithmetic-coded payload.  tmc13ref_symbols_payload feeds a symbol stream
// (zero runs + values) to the same PCCResidualsEncoder: mode 0 through its
// encodeRunLength / encode members, mode 1 through encodeSymbol with context
// selectors supplied by the caller — equal payloads prove the stream (and the
// selectors) are the reference's.

static void
fillRaht(const pccb200_raht_params* p, AttributeParameterSet& aps)
{
  auto& rp = aps.rahtPredParams;
  rp.raht_prediction_enabled_flag = p->prediction_enabled != 0;
  rp.integer_haar_enable_flag = p->integer_haar != 0;
  rp.raht_prediction_threshold0 = p->prediction_threshold0;
  rp.raht_prediction_threshold1 = p->prediction_threshold1;
  rp.raht_subnode_prediction_enabled_flag = p->subnode_prediction_enabled != 0;
  rp.raht_prediction_search_range = p->prediction_search_range;
  rp.predWeightParent.assign(p->pred_weight_parent, p->pred_weight_parent + 19);
  rp.predWeightChild.assign(p->pred_weight_child, p->pred_weight_child + 12);
  aps.raht_extension = p->raht_extension != 0;
  aps.attr_encoding = AttributeEncoding::kRAHTransform;
}

static QpSet
mkQpSet2(const pccb200_qpset* q)
{
  QpSet qs;
  for (int i = 0; i < q->num_layers; i++)
    qs.layers.push_back(Qps{q->layers[i][0], q->layers[i][1]});
  qs.maxQp = q->max_qp;
  qs.fixedPointQpOffset = q->fixed_point_qp_offset;
  for (int l = 0; l < q->num_ac_coeff_qp_layers; l++) {
    std::vector<Qps> layer;
    for (int c = 0; c < 7; c++)
      layer.push_back(Qps{q->ac_coeff_qps[l][c][0], q->ac_coeff_qps[l][c][1]});
    qs.rahtAcCoeffQps.push_back(layer);
  }
  return qs;
}

extern "C" int
tmc13ref_raht_encode_payload(
  const pccb200_raht_params* params,
  const pccb200_qpset* qs,
  const int32_t* xyz,
  const int32_t* attrs,  // n x numAttrs, input order
  int n,
  int numAttrs,
  int bitdepth,
  uint8_t* buf,
  int cap,
  int32_t* reconOut)  // n x numAttrs, input order (or null)
{
  AttributeParameterSet aps{};
  fillRaht(params, aps);
  AttributeBrickHeader abh{};
  AttributeDescription desc{};
  desc.bitdepth = bitdepth;
  desc.attr_num_dimensions_minus1 = numAttrs - 1;
  SequenceParameterSet sps{};
  QpSet qpSet = mkQpSet2(qs);

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

  AttributeEncoder enc;
  enc._abh = &abh;
  AttributeContexts ctxtMem;
  ctxtMem.reset();
  PCCResidualsEncoder encoder(aps, abh, ctxtMem);
  encoder.start(sps, n);
  if (numAttrs == 3)
    enc.encodeColorsTransformRaht(desc, aps, qpSet, cloud, encoder, ip);
  else
    enc.encodeReflectancesTransformRaht(desc, aps, qpSet, cloud, encoder, ip);
  int len = encoder.stop();
  if (len > cap)
    return -len;
  memcpy(buf, encoder.arithmeticEncoder.buffer(), len);
  if (reconOut)
    for (int i = 0; i < n; i++) {
      if (numAttrs == 3) {
        auto c = cloud.getColor(i);
        for (int k = 0; k < 3; k++)
          reconOut[3 * i + k] = c[k];
      } else {
        reconOut[i] = cloud.getReflectance(i);
      }
    }
  return len;
}

extern "C" int
tmc13ref_symbols_payload(
  int mode,
  const int32_t* runs,
  const int32_t* values,  // count x numAttrs
  const uint8_t* ctx,     // count (mode 1, three components)
  int count,
  int tailRun,
  int numAttrs,
  int n,
  uint8_t* buf,
  int cap)
{
  AttributeParameterSet aps{};
  aps.attr_encoding = AttributeEncoding::kRAHTransform;
  AttributeBrickHeader abh{};
  SequenceParameterSet sps{};
  AttributeContexts ctxtMem;
  ctxtMem.reset();
  PCCResidualsEncoder encoder(aps, abh, ctxtMem);
  encoder.start(sps, n);
  for (int s = 0; s < count; s++) {
    encoder.encodeRunLength(runs[s]);
    const int32_t* v = &values[s * numAttrs];
    if (mode == 0) {
      if (numAttrs == 3)
        encoder.encode(v[0], v[1], v[2]);
      else
        encoder.encode(v[0]);
    } else if (numAttrs == 3) {
      // the caller's selectors instead of the ones encode() would derive
      const int b0 = ctx[s] & 1, b1 = (ctx[s] >> 1) & 1, b2 = (ctx[s] >> 2) & 1,
                b3 = (ctx[s] >> 3) & 1;
      const int mag0 = abs(v[0]), mag1 = abs(v[1]), mag2 = abs(v[2]);
      encoder.encodeSymbol(mag1, 0, 0, 1);
      encoder.encodeSymbol(mag2, 1 + b0, 1 + b1, 1);
      encoder.encodeSymbol(b0 && b2 ? mag0 - 1 : mag0, 3 + (b0 << 1) + b2, 3 + (b1 << 1) + b3, 0);
      if (mag0)
        encoder.arithmeticEncoder.encode(v[0] < 0);
      if (mag1)
        encoder.arithmeticEncoder.encode(v[1] < 0);
      if (mag2)
        encoder.arithmeticEncoder.encode(v[2] < 0);
    } else {
      encoder.encodeSymbol(abs(v[0]) - 1, 0, 0, 0);
      encoder.arithmeticEncoder.encode(v[0] < 0);
    }
  }
  if (tailRun)
    encoder.encodeRunLength(tailRun);
  int len = encoder.stop();
  if (len > cap)
    return -len;
  memcpy(buf, encoder.arithmeticEncoder.buffer(), len);
  return len;
}

//============================================================================
// estimateDist2 (tmc3/AttributeEncoder.cpp:1683-1720), a free function of the
// included translation unit
extern "C" int
tmc13ref_estimate_dist2(
  const int32_t* xyz, int n, int samplingPeriod, int searchRange, float percentile)
{
  PCCPointSet3 cloud;
  cloud.resize(n);
  for (int i = 0; i < n; i++)
    cloud[i] = point_t{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  return estimateDist2(cloud, samplingPeriod, searchRange, percentile);
}
