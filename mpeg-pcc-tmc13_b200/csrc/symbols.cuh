// symbols.cuh — symbol preparation for the entropy coder:
// the step right after the RAHT forward transform.  The reference walks the
// planar coefficients once, counting positions whose components are all zero
// and handing (run length, values) to the arithmetic coder
// (tmc3/AttributeEncoder.cpp:1279-1291 one component, :1346-1362 three), which
// derives its context selectors from the magnitudes
// (PCCResidualsEncoder::encode, :271-296).  The arithmetic coder itself is
// serial and stays on the host; the run-length extraction is a stream
// compaction and the selectors are per-symbol arithmetic, so the device hands
// over the symbol stream instead of N x A mostly-zero coefficients.
#pragma once

#include "pcc_arith.cuh"

namespace pccb200 {

struct SymbolPred {
  const int32_t* coef;  // component k at coef + k * stride
  int64_t stride;
  int A;
  PCC_HD bool operator()(int64_t i) const
  {
    int32_t any = 0;
    for (int k = 0; k < A; k++)
      any |= coef[k * stride + i];
    return any != 0;
  }
};

// context selectors of the three-component symbol coder:
// b0 = |v1| == 0, b1 = |v1| <= 1, b2 = |v2| == 0, b3 = |v2| <= 1
PCC_HD uint8_t
symbol_ctx(int32_t v1, int32_t v2)
{
  const uint32_t m1 = v1 < 0 ? uint32_t(-int64_t(v1)) : uint32_t(v1);
  const uint32_t m2 = v2 < 0 ? uint32_t(-int64_t(v2)) : uint32_t(v2);
  return uint8_t((m1 == 0) | (m1 <= 1) << 1 | (m2 == 0) << 2 | (m2 <= 1) << 3);
}

struct SymbolEmit {
  const int32_t* coef;
  int64_t stride;
  int A;
  int32_t* pos;     // position of symbol `rank`
  int32_t* values;  // rank * A + k
  uint8_t* ctx;     // rank (A == 3) or null
  PCC_HD void operator()(int64_t rank, int64_t i) const
  {
    pos[rank] = int32_t(i);
    int32_t v[3] = {0, 0, 0};
    for (int k = 0; k < A; k++)
      values[rank * A + k] = v[k] = coef[k * stride + i];
    if (ctx && A == 3)
      ctx[rank] = symbol_ctx(v[1], v[2]);
  }
};

struct SymbolRunFn {  // all-zero positions between consecutive symbols
  const int32_t* pos;
  int32_t* runs;
  PCC_HD void operator()(int64_t r) const { runs[r] = pos[r] - (r ? pos[r - 1] : -1) - 1; }
};

// dRuns[count], dValues[count * A], dCtx[count] (or null); *hostTail = zero
// positions after the last symbol
template<class Exec>
void
run_coeff_symbols(Exec& ex, const int32_t* dCoef, int64_t stride, int A, int n, int32_t* dRuns,
                  int32_t* dValues, uint8_t* dCtx, int* hostCount, int* hostTail)
{
  int32_t* pos = ex.template alloc<int32_t>(size_t(n) + 1);
  int* dCount = ex.template alloc<int>(1);
  ex.compact(n, SymbolPred{dCoef, stride, A}, SymbolEmit{dCoef, stride, A, pos, dValues, dCtx},
             dCount);
  int count = 0;
  ex.download(&count, dCount, sizeof(int));
  int last = -1;
  if (count > 0) {
    ex.foreach(count, SymbolRunFn{pos, dRuns});
    ex.download(&last, pos + count - 1, sizeof(int32_t));
  }
  *hostCount = count;
  *hostTail = n - 1 - last;
}

}  // namespace pccb200
