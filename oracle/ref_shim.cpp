// ref_shim.cpp — TEST INFRASTRUCTURE ONLY (oracle).
//
// Thin extern "C" wrapper around the UNMODIFIED reference sources where they
// lie under /root/reference (never copied into this repository).  Compiled by
// oracle/Makefile together with the reference's own translation units into
// oracle/_ref/libtmc13_ref.so.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.
//
// Wrapped reference entry points:
//   pcc::regionAdaptiveHierarchicalTransform         tmc3/RAHT.cpp:1997
//   pcc::regionAdaptiveHierarchicalInverseTransform  tmc3/RAHT.cpp:2037
//   pcc::mortonAddr + std::sort(MortonCodeWithIndex) tmc3/AttributeEncoder.cpp:1316-1321
//   pcc::isqrt / pcc::irsqrt                         tmc3/misc.cpp:138-225
//   pcc::Quantizer                                   tmc3/quantization.h:53-102
//   pcc::PCCComputeQuantizationWeights / PCCLiftPredict / PCCLiftUpdate
//                                                    tmc3/PCCTMC3Common.h:716-854
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "RAHT.h"
#include "AttributeCommon.h"
#include "PCCTMC3Common.h"
#include "PCCMisc.h"
#include "quantization.h"
#include "FixedPoint.h"

#include "pcc_attr_b200.h"

using namespace pcc;

namespace {
RahtPredictionParams
mkParams(const pccb200_raht_params* p)
{
  RahtPredictionParams rp;
  rp.raht_prediction_enabled_flag = p->prediction_enabled != 0;
  rp.integer_haar_enable_flag = p->integer_haar != 0;
  rp.raht_prediction_threshold0 = p->prediction_threshold0;
  rp.raht_prediction_threshold1 = p->prediction_threshold1;
  rp.raht_subnode_prediction_enabled_flag = p->subnode_prediction_enabled != 0;
  rp.raht_prediction_search_range = p->prediction_search_range;
  rp.predWeightParent.assign(p->pred_weight_parent, p->pred_weight_parent + 19);
  rp.predWeightChild.assign(p->pred_weight_child, p->pred_weight_child + 12);
  return rp;
}

QpSet
mkQpSet(const pccb200_qpset* q)
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

AttributeInterPredParams
mkIntra()
{
  AttributeInterPredParams ip;
  ip.frameDistance = 1;
  ip.enableAttrInterPred = false;
  ip.attrInterIntraSliceRDO = false;
  return ip;
}
}  // namespace

extern "C" {

// forward != 0: encoder (attrs in/out, coeffs out); else decoder.
// Returns elapsed seconds of the reference call alone (steady_clock).
double
tmc13ref_raht(
  int forward,
  const pccb200_raht_params* params,
  const pccb200_qpset* qpset,
  const int32_t* pointQpOffsets,
  const int64_t* morton,
  int32_t* attrs,
  int numAttrs,
  int n,
  int32_t* coeffs)
{
  auto rp = mkParams(params);
  auto qs = mkQpSet(qpset);
  auto ip = mkIntra();
  std::vector<Qps> qpo(n, Qps{0, 0});
  if (pointQpOffsets)
    for (int i = 0; i < n; i++)
      qpo[i] = Qps{pointQpOffsets[2 * i], pointQpOffsets[2 * i + 1]};
  std::vector<int64_t> mc(morton, morton + n);
  auto t0 = std::chrono::steady_clock::now();
  if (forward)
    regionAdaptiveHierarchicalTransform(
      rp, qs, qpo.data(), mc.data(), attrs, numAttrs, n, coeffs,
      params->raht_extension != 0, ip);
  else
    regionAdaptiveHierarchicalInverseTransform(
      rp, qs, qpo.data(), mc.data(), attrs, numAttrs, n, coeffs,
      params->raht_extension != 0, ip);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// mortonAddr + sort, as the attribute coders do it.  Returns seconds.
double
tmc13ref_morton_sort(const int32_t* xyz, int n, int64_t* keys, int32_t* order)
{
  auto t0 = std::chrono::steady_clock::now();
  std::vector<MortonCodeWithIndex> packed(n);
  for (int i = 0; i < n; i++) {
    packed[i].mortonCode =
      mortonAddr(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    packed[i].index = i;
  }
  std::sort(packed.begin(), packed.end());
  for (int i = 0; i < n; i++) {
    keys[i] = packed[i].mortonCode;
    order[i] = packed[i].index;
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// The whole timed region of the attribute coder's RAHT path, minus entropy
// coding: sort, gather, transform, clip, write back
// (tmc3/AttributeEncoder.cpp:1306-1375 / AttributeDecoder.cpp:613-674).
double
tmc13ref_attr_raht(
  int forward,
  const pccb200_raht_params* params,
  const pccb200_qpset* qpset,
  const int32_t* pointQpOffsets,
  const int32_t* xyz,
  int32_t* attrs,
  int numAttrs,
  int n,
  int bitdepth,
  int32_t* coeffs)
{
  auto rp = mkParams(params);
  auto qs = mkQpSet(qpset);
  auto ip = mkIntra();
  auto t0 = std::chrono::steady_clock::now();
  std::vector<MortonCodeWithIndex> packed(n);
  for (int i = 0; i < n; i++) {
    packed[i].mortonCode =
      mortonAddr(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    packed[i].index = i;
  }
  std::sort(packed.begin(), packed.end());
  std::vector<int64_t> mc(n);
  std::vector<int> a(size_t(n) * numAttrs);
  std::vector<Qps> qpo(n, Qps{0, 0});
  for (int i = 0; i < n; i++) {
    mc[i] = packed[i].mortonCode;
    int src = packed[i].index;
    for (int k = 0; k < numAttrs; k++)
      a[size_t(i) * numAttrs + k] = attrs[size_t(src) * numAttrs + k];
    if (pointQpOffsets)
      qpo[i] = Qps{pointQpOffsets[2 * src], pointQpOffsets[2 * src + 1]};
  }
  if (forward)
    regionAdaptiveHierarchicalTransform(
      rp, qs, qpo.data(), mc.data(), a.data(), numAttrs, n, coeffs,
      params->raht_extension != 0, ip);
  else
    regionAdaptiveHierarchicalInverseTransform(
      rp, qs, qpo.data(), mc.data(), a.data(), numAttrs, n, coeffs,
      params->raht_extension != 0, ip);
  const int64_t clipMax = (1 << bitdepth) - 1;
  for (int i = 0; i < n; i++) {
    int dst = packed[i].index;
    for (int k = 0; k < numAttrs; k++)
      attrs[size_t(dst) * numAttrs + k] =
        int32_t(PCCClip(int64_t(a[size_t(i) * numAttrs + k]), 0, clipMax));
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// scalar helpers, for known-answer tests of the device arithmetic
uint32_t tmc13ref_isqrt(uint64_t x) { return isqrt(x); }
uint64_t tmc13ref_irsqrt(uint64_t x) { return irsqrt(x); }
int64_t tmc13ref_morton_addr(int32_t x, int32_t y, int32_t z) { return mortonAddr(x, y, z); }
uint64_t tmc13ref_morton3d_add(uint64_t a, uint64_t b) { return morton3dAdd(a, b); }
int64_t tmc13ref_quantize(int qp, int64_t x) { return Quantizer(qp).quantize(x); }
int64_t tmc13ref_scale(int qp, int64_t x) { return Quantizer(qp).scale(x); }
int64_t tmc13ref_fixed_mul(int64_t a, int64_t b)
{
  FixedPoint x, y;
  x.val = a;
  y.val = b;
  x *= y;
  return x.val;
}
int64_t tmc13ref_div_approx(int64_t a, uint64_t b, int32_t log2Scale) { return divApprox(a, b, log2Scale); }

// Lifting: quantisation weights and forward / inverse lifting over all LoDs,
// driven exactly as tmc3/AttributeEncoder.cpp:1391,1408-1415,1476-1482.
static void
mkPredictors(
  const pccb200_predictor* preds, int n, std::vector<PCCPredictor>& out)
{
  out.resize(n);
  for (int i = 0; i < n; i++) {
    auto& p = out[i];
    p.neighborCount = preds[i].neighbor_count;
    p.predMode = 0;
    for (int j = 0; j < 3; j++) {
      p.neighbors[j].predictorIndex = preds[i].predictor_index[j];
      p.neighbors[j].weight = preds[i].weight[j];
      p.neighbors[j].pointIndex = 0;
      p.neighbors[j].interFrameRef = false;
    }
  }
}

double
tmc13ref_quant_weights(const pccb200_predictor* preds, int n, uint64_t* qw)
{
  std::vector<PCCPredictor> predictors;
  mkPredictors(preds, n, predictors);
  std::vector<uint64_t> w;
  auto t0 = std::chrono::steady_clock::now();
  PCCComputeQuantizationWeights(predictors, w);
  auto t1 = std::chrono::steady_clock::now();
  std::copy(w.begin(), w.end(), qw);
  return std::chrono::duration<double>(t1 - t0).count();
}

void
tmc13ref_quant_weights_fixed(
  const pccb200_predictor* preds, int n, const int32_t neighWeight[3], uint64_t* qw)
{
  std::vector<PCCPredictor> predictors;
  mkPredictors(preds, n, predictors);
  std::vector<uint64_t> w;
  computeQuantizationWeights(
    predictors, w, Vec3<int32_t>(neighWeight[0], neighWeight[1], neighWeight[2]));
  std::copy(w.begin(), w.end(), qw);
}

void
tmc13ref_quant_weights_scalable(
  const pccb200_predictor* preds, int n, const uint32_t* npl, int lodCount, uint64_t numPoints,
  int minGeomNodeSizeLog2, uint64_t* qw)
{
  std::vector<PCCPredictor> predictors;
  mkPredictors(preds, n, predictors);
  std::vector<uint32_t> perLod(npl, npl + lodCount);
  std::vector<uint64_t> w;
  computeQuantizationWeightsScalable(predictors, perLod, numPoints, minGeomNodeSizeLog2, w);
  std::copy(w.begin(), w.end(), qw);
}

double
tmc13ref_lift(
  int forward,
  const pccb200_predictor* preds,
  const uint64_t* qw,
  int n,
  const uint32_t* numPointsInLod,
  int lodCount,
  int64_t* attrs,
  int numAttrs)
{
  std::vector<PCCPredictor> predictors;
  mkPredictors(preds, n, predictors);
  std::vector<uint64_t> weights(qw, qw + n);
  double secs = 0;
  if (numAttrs == 3) {
    std::vector<Vec3<int64_t>> v(n);
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++)
        v[i][k] = attrs[size_t(i) * 3 + k];
    auto t0 = std::chrono::steady_clock::now();
    if (forward) {
      for (int l = lodCount - 2; l >= 0; --l) {
        int s = numPointsInLod[l], e = numPointsInLod[l + 1];
        PCCLiftPredict(predictors, s, e, true, v);
        PCCLiftUpdate(predictors, weights, s, e, true, v);
      }
    } else {
      for (int l = 0; l < lodCount - 1; ++l) {
        int s = numPointsInLod[l], e = numPointsInLod[l + 1];
        PCCLiftUpdate(predictors, weights, s, e, false, v);
        PCCLiftPredict(predictors, s, e, false, v);
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    secs = std::chrono::duration<double>(t1 - t0).count();
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++)
        attrs[size_t(i) * 3 + k] = v[i][k];
  } else {
    std::vector<int64_t> v(attrs, attrs + n);
    auto t0 = std::chrono::steady_clock::now();
    if (forward) {
      for (int l = lodCount - 2; l >= 0; --l) {
        int s = numPointsInLod[l], e = numPointsInLod[l + 1];
        PCCLiftPredict(predictors, s, e, true, v);
        PCCLiftUpdate(predictors, weights, s, e, true, v);
      }
    } else {
      for (int l = 0; l < lodCount - 1; ++l) {
        int s = numPointsInLod[l], e = numPointsInLod[l + 1];
        PCCLiftUpdate(predictors, weights, s, e, false, v);
        PCCLiftPredict(predictors, s, e, false, v);
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    secs = std::chrono::duration<double>(t1 - t0).count();
    std::copy(v.begin(), v.end(), attrs);
  }
  return secs;
}

// AttributeLods::generate (tmc3/AttributeCommon.cpp:45-72), parameters set
// as tmc3/encoder.cpp:777-818 leaves them.  Returns seconds.
double
tmc13ref_lod_build(
  const pccb200_lod_params* lp,
  const int32_t* xyz,
  int n,
  pccb200_predictor* predsOut,
  uint32_t* indexesOut,
  uint32_t* numPointsInLodOut,
  int32_t* lodCountOut)
{
  AttributeParameterSet aps{};
  aps.attr_encoding = lp->pred_weight_blending
    ? AttributeEncoding::kPredictingTransform
    : AttributeEncoding::kLiftingTransform;
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
  AttributeBrickHeader abh{};
  abh.attr_dist2_delta = 0;
  AttributeInterPredParams ip = mkIntra();

  PCCPointSet3 cloud;
  cloud.resize(n);
  for (int i = 0; i < n; i++)
    cloud[i] = point_t{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};

  AttributeLods lods;
  auto t0 = std::chrono::steady_clock::now();
  lods.generate(aps, abh, n - 1, 0, cloud, ip);
  auto t1 = std::chrono::steady_clock::now();

  for (int i = 0; i < n; i++) {
    const auto& p = lods.predictors[i];
    predsOut[i].neighbor_count = p.neighborCount;
    for (int j = 0; j < 3; j++) {
      predsOut[i].predictor_index[j] = j < int(p.neighborCount) ? p.neighbors[j].predictorIndex : 0;
      predsOut[i].weight[j] = j < int(p.neighborCount) ? uint32_t(p.neighbors[j].weight) : 0;
    }
    indexesOut[i] = lods.indexes[i];
  }
  *lodCountOut = int(lods.numPointsInLod.size());
  for (size_t i = 0; i < lods.numPointsInLod.size() && i < PCCB200_MAX_LODS; i++)
    numPointsInLodOut[i] = lods.numPointsInLod[i];
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
