// lifting.cuh — the lifting transform's data-parallel passes.
//
//   quantisation weights  PCCComputeQuantizationWeights  tmc3/PCCTMC3Common.h:828-854
//   prediction step       PCCLiftPredict                 tmc3/PCCTMC3Common.h:716-770
//   update step           PCCLiftUpdate                  tmc3/PCCTMC3Common.h:774-824
//   LoD loops             tmc3/AttributeEncoder.cpp:1408-1415 (forward),
//                         tmc3/AttributeEncoder.cpp:1476-1482 (inverse)
//
// The reference walks predictors one by one; because the lifting transform
// only ever predicts from strictly coarser levels of detail
// (tmc3/encoder.cpp:777-780), every LoD is one data-parallel launch, and the
// scatter-adds of the update step and of the quantisation weights are 64-bit
// integer atomics: addition modulo 2^64 is order independent, so the results
// are bit-identical to the sequential walk.  A predictor that references its
// own LoD is reported through an error word and executed by a single ordered
// thread instead.
#pragma once

#include <vector>

#include "pcc_arith.cuh"
#include "raht_core.cuh"

namespace pccb200 {

PCC_HD void
atomic_add_u64(uint64_t* p, uint64_t v)
{
#if defined(__CUDA_ARCH__)
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
#else
  *p += v;
#endif
}

struct FillU64Fn {
  uint64_t* p;
  uint64_t v;
  PCC_HD void operator()(int64_t i) const { p[i] = v; }
};

// does any predictor in [start, start + n) reference an index >= start ?
struct LodCheckFn {
  const pccb200_predictor* preds;
  int64_t start;
  int* flag;
  PCC_HD void operator()(int64_t i) const
  {
    const pccb200_predictor& p = preds[start + i];
    for (uint32_t j = 0; j < p.neighbor_count; j++)
      if (int64_t(p.predictor_index[j]) >= start)
        atomic_or_i32(flag, 1);
  }
};

struct QuantWeightLodFn {
  const pccb200_predictor* preds;
  uint64_t* qw;
  int64_t start;
  PCC_HD void operator()(int64_t i) const
  {
    const pccb200_predictor& p = preds[start + i];
    const uint64_t w = qw[start + i];
    for (uint32_t j = 0; j < p.neighbor_count; j++)
      atomic_add_u64(&qw[p.predictor_index[j]],
                     div_exp2_round_half_inf_u(uint64_t(p.weight[j]) * w, 8));
  }
};

// sequential walk of a range, last to first (used when a LoD references itself)
struct QuantWeightSeqFn {
  const pccb200_predictor* preds;
  uint64_t* qw;
  int64_t start, end;
  PCC_HD void operator()(int64_t) const
  {
    for (int64_t i = end - 1; i >= start; i--) {
      const pccb200_predictor& p = preds[i];
      const uint64_t w = qw[i];
      for (uint32_t j = 0; j < p.neighbor_count; j++)
        qw[p.predictor_index[j]] +=
          div_exp2_round_half_inf_u(uint64_t(p.weight[j]) * w, 8);
    }
  }
};

struct LiftPredictFn {
  const pccb200_predictor* preds;
  int64_t* attr;
  int64_t start;
  int A;
  int direct;
  PCC_HD void operator()(int64_t i) const
  {
    const int64_t idx = start + i;
    const pccb200_predictor& p = preds[idx];
    for (int k = 0; k < A; k++) {
      int64_t acc = 0;
      for (uint32_t j = 0; j < p.neighbor_count; j++)
        acc += int64_t(p.weight[j]) * attr[int64_t(p.predictor_index[j]) * A + k];
      acc = div_exp2_round_half_inf(acc, 8);
      if (direct)
        attr[idx * A + k] -= acc;
      else
        attr[idx * A + k] += acc;
    }
  }
};

struct LiftUpdateScatterFn {
  const pccb200_predictor* preds;
  const uint64_t* qw;
  const int64_t* attr;
  uint64_t* updW;  // [start]
  uint64_t* upd;   // [start * A]
  int64_t start;
  int A;
  PCC_HD void operator()(int64_t i) const
  {
    const int64_t idx = start + i;
    const pccb200_predictor& p = preds[idx];
    const uint64_t q = qw[idx];
    for (uint32_t j = 0; j < p.neighbor_count; j++) {
      const uint64_t w = div_exp2_round_half_inf_u(uint64_t(p.weight[j]) * q, 8);
      const int64_t nb = p.predictor_index[j];
      atomic_add_u64(&updW[nb], w);
      for (int k = 0; k < A; k++)
        atomic_add_u64(&upd[nb * A + k], w * uint64_t(attr[idx * A + k]));
    }
  }
};

struct LiftUpdateApplyFn {
  const uint64_t* updW;
  const uint64_t* upd;
  int64_t* attr;
  int A;
  int direct;
  PCC_HD void operator()(int64_t i) const
  {
    const uint32_t sumW = uint32_t(updW[i]);  // the reference truncates to 32 bits
    if (!sumW)
      return;
    for (int k = 0; k < A; k++) {
      int64_t u = div_approx(int64_t(upd[i * A + k]), sumW, 0);
      if (direct)
        attr[i * A + k] += u;
      else
        attr[i * A + k] -= u;
    }
  }
};

// executor-generic drivers (numPointsInLod is a host array)

template<class Exec>
int
run_quant_weights(Exec& ex, const pccb200_predictor* preds, int64_t n,
                  const uint32_t* numPointsInLod, int lodCount, uint64_t* qw)
{
  ex.phase(5);
  ex.foreach(n, FillU64Fn{qw, uint64_t(1) << 8});
  int* dFlags = ex.template alloc<int>(size_t(lodCount) + 1);
  ex.zero(dFlags, (size_t(lodCount) + 1) * sizeof(int));
  int64_t prevEnd = 0;
  for (int l = 0; l < lodCount; l++) {
    int64_t s = l ? numPointsInLod[l - 1] : 0;
    int64_t e = numPointsInLod[l];
    if (s != prevEnd || e < s || e > n)
      return PCCB200_ERR_INVALID_ARG;
    prevEnd = e;
    ex.foreach(e - s, LodCheckFn{preds, s, dFlags + l});
  }
  if (prevEnd != n)
    return PCCB200_ERR_INVALID_ARG;
  std::vector<int> flags(size_t(lodCount) + 1);
  ex.download(flags.data(), dFlags, flags.size() * sizeof(int));
  for (int l = lodCount - 1; l >= 0; l--) {
    int64_t s = l ? numPointsInLod[l - 1] : 0;
    int64_t e = numPointsInLod[l];
    if (flags[l])
      ex.foreach(1, QuantWeightSeqFn{preds, qw, s, e});
    else
      ex.foreach(e - s, QuantWeightLodFn{preds, qw, s});
  }
  return PCCB200_OK;
}

template<class Exec>
int
run_lift(Exec& ex, bool forward, const pccb200_predictor* preds, const uint64_t* qw,
         int64_t n, const uint32_t* numPointsInLod, int lodCount, int64_t* attr, int A)
{
  if (lodCount < 1 || int64_t(numPointsInLod[lodCount - 1]) != n)
    return PCCB200_ERR_INVALID_ARG;
  ex.phase(5);
  // the lifting passes require strictly-coarser references
  int* dFlag = ex.template alloc<int>(1);
  ex.zero(dFlag, sizeof(int));
  for (int l = 1; l < lodCount; l++) {
    int64_t s = numPointsInLod[l - 1], e = numPointsInLod[l];
    if (e < s || e > n)
      return PCCB200_ERR_INVALID_ARG;
    ex.foreach(e - s, LodCheckFn{preds, s, dFlag});
  }
  int flag = 0;
  ex.download(&flag, dFlag, sizeof(int));
  if (flag)
    return PCCB200_ERR_UNSUPPORTED;

  int64_t maxStart = lodCount > 1 ? numPointsInLod[lodCount - 2] : 0;
  uint64_t* updW = ex.template alloc<uint64_t>(size_t(maxStart));
  uint64_t* upd = ex.template alloc<uint64_t>(size_t(maxStart) * A);
  auto update = [&](int64_t s, int64_t e, bool direct) {
    ex.zero(updW, size_t(s) * sizeof(uint64_t));
    ex.zero(upd, size_t(s) * A * sizeof(uint64_t));
    ex.foreach(e - s, LiftUpdateScatterFn{preds, qw, attr, updW, upd, s, A});
    ex.foreach(s, LiftUpdateApplyFn{updW, upd, attr, A, direct});
  };
  if (forward) {
    for (int l = lodCount - 1; l >= 1; l--) {
      int64_t s = numPointsInLod[l - 1], e = numPointsInLod[l];
      ex.foreach(e - s, LiftPredictFn{preds, attr, s, A, 1});
      update(s, e, true);
    }
  } else {
    for (int l = 1; l < lodCount; l++) {
      int64_t s = numPointsInLod[l - 1], e = numPointsInLod[l];
      update(s, e, false);
      ex.foreach(e - s, LiftPredictFn{preds, attr, s, A, 0});
    }
  }
  return PCCB200_OK;
}

}  // namespace pccb200
