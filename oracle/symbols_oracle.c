/* symbols_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the coefficient walk in front of the arithmetic coder:
 *   one component    tmc3/AttributeEncoder.cpp:1279-1291
 *   three components tmc3/AttributeEncoder.cpp:1346-1362
 * and of the context selectors PCCResidualsEncoder::encode derives from the
 * magnitudes (tmc3/AttributeEncoder.cpp:271-296).  Pinned against the reference's
 * own bitstream by tests/test_oracle_vs_reference.py::test_live_symbols. */
#include <stdint.h>
#include <stdlib.h>

/* coeffs: A x n planar.  Returns the number of symbols. */
int
oracle_coeff_symbols(const int32_t* coeffs, int A, int n, int32_t* runs, int32_t* values,
                     uint8_t* ctx, int32_t* tail_run)
{
  int count = 0;
  int zero_run = 0;
  for (int i = 0; i < n; i++) {
    int32_t v[3] = {0, 0, 0};
    int any = 0;
    for (int d = 0; d < A; d++) {
      v[d] = coeffs[(int64_t)n * d + i];
      any |= v[d] != 0;
    }
    if (!any) {
      ++zero_run;
      continue;
    }
    runs[count] = zero_run; /* encoder.encodeRunLength(zeroRun) */
    for (int d = 0; d < A; d++)
      values[(int64_t)count * A + d] = v[d]; /* encoder.encode(...) */
    if (ctx && A == 3) {
      int64_t mag1 = llabs((int64_t)v[1]), mag2 = llabs((int64_t)v[2]);
      int b0 = mag1 == 0, b1 = mag1 <= 1, b2 = mag2 == 0, b3 = mag2 <= 1;
      ctx[count] = (uint8_t)(b0 | b1 << 1 | b2 << 2 | b3 << 3);
    }
    count++;
    zero_run = 0;
  }
  *tail_run = zero_run; /* if (zeroRun) encoder.encodeRunLength(zeroRun) */
  return count;
}
