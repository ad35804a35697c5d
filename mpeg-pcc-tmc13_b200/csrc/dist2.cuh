// dist2.cuh — estimateDist2 (tmc3/AttributeEncoder.cpp:1683-1720, called per
// slice from tmc3/encoder.cpp:1199-1206; SURVEY 8(f) row N3, first half): the
// encoder's estimate of the LoD sampling distance, abh.attr_dist2_delta.
//
// Every samplingPeriod-th point looks for its nearest neighbour (squared L2)
// among the searchRange points before and after it in coding order; the
// percentile of those distances is turned into the smallest shift s with
// 3 << 2s >= distance.  The reference selects the percentile with
// nth_element; only the shift is used, so the device counts, per sample, the
// smallest shift that covers its distance (22 counters, order independent)
// and the host reads the percentile off the cumulative counts.
#pragma once

#include <math.h>

#include "pcc_arith.cuh"

namespace pccb200 {

constexpr int kDist2MaxShift = 20;

PCC_HD void
atomic_inc_u32(uint32_t* p)
{
#if defined(__CUDA_ARCH__)
  atomicAdd(p, 1u);
#else
  ++*p;
#endif
}

// smallest s in [0, 20] with (3 << 2s) >= d2, as the reference's final loop
PCC_HD int
dist2_shift(int64_t d2)
{
  int s = 0;
  while ((int64_t(3) << (s << 1)) < d2 && s < kDist2MaxShift)
    ++s;
  return s;
}

struct Dist2SampleFn {
  const int32_t* xyz;
  int n;
  int samplingPeriod;
  int searchRange;
  uint32_t* hist;  // kDist2MaxShift + 1 counters: samples whose distance needs shift s
  PCC_HD void operator()(int64_t sample) const
  {
    const int index = int(sample) * samplingPeriod;
    const int k0 = index - searchRange > 0 ? index - searchRange : 0;
    const int k1 = index + searchRange < n - 1 ? index + searchRange : n - 1;
    int64_t d2 = INT64_MAX;
    const int64_t px = xyz[3 * size_t(index)], py = xyz[3 * size_t(index) + 1],
                  pz = xyz[3 * size_t(index) + 2];
    for (int k = k0; k <= k1; k++) {
      if (k == index)
        continue;
      const int64_t dx = px - xyz[3 * size_t(k)], dy = py - xyz[3 * size_t(k) + 1],
                    dz = pz - xyz[3 * size_t(k) + 2];
      const int64_t d = dx * dx + dy * dy + dz * dz;
      d2 = d < d2 ? d : d2;
    }
    atomic_inc_u32(&hist[dist2_shift(d2)]);
  }
};

// returns the shift, or -1 for a percentile outside [0, 1)
template<class Exec>
int
run_estimate_dist2(Exec& ex, const int32_t* dXyz, int n, int samplingPeriod, int searchRange,
                   float percentileEstimate)
{
  if (n < 2)
    return 0;
  const int samples = (n + samplingPeriod - 1) / samplingPeriod;
  // the reference's index arithmetic: size_t * float in single precision
  const int p = int(floorf(float(size_t(samples)) * percentileEstimate));
  if (p < 0 || p >= samples)
    return -1;
  uint32_t* dHist = ex.template alloc<uint32_t>(kDist2MaxShift + 1);
  ex.zero(dHist, (kDist2MaxShift + 1) * sizeof(uint32_t));
  ex.foreach(samples, Dist2SampleFn{dXyz, n, samplingPeriod, searchRange, dHist});
  uint32_t hist[kDist2MaxShift + 1];
  ex.download(hist, dHist, sizeof(hist));
  // the p-th smallest distance needs the first shift whose cumulative count exceeds p
  uint64_t cum = 0;
  for (int s = 0; s <= kDist2MaxShift; s++) {
    cum += hist[s];
    if (cum > uint64_t(p))
      return s;
  }
  return kDist2MaxShift;
}

}  // namespace pccb200
