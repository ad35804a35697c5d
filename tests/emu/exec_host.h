// exec_host.h — TEST INFRASTRUCTURE ONLY.
//
// Host "executor" for the kernel bodies in mpeg-pcc-tmc13_b200/csrc: runs
// every per-item functor as a plain in-order loop so that the kernels' logic
// (stage planning, node construction, coefficient addressing, the dataflow
// protocol) can be unit-tested in a container without a GPU.  It is never
// linked into the product library; the product path has no CPU fallback.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "raht_core.cuh"

struct HostExec {
  std::vector<void*> blocks;
  ~HostExec()
  {
    for (void* p : blocks)
      free(p);
  }
  template<class T>
  T* alloc(size_t n)
  {
    void* p = malloc((n ? n : 1) * sizeof(T));
    memset(p, 0xCD, (n ? n : 1) * sizeof(T));  // poison: catch reads of unset data
    blocks.push_back(p);
    return static_cast<T*>(p);
  }
  void phase(int) {}
  void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
  void fill(void* p, int byte, size_t bytes) { memset(p, byte, bytes); }
  void upload(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }
  void download(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }
  template<class F>
  void foreach(int64_t n, const F& f)
  {
    for (int64_t i = 0; i < n; i++)
      f(i);
  }
  template<class F>
  void ordered(int64_t n, const F& f)
  {
    for (int64_t i = 0; i < n; i++)
      f(i);
  }
  void morton_sort(const int32_t* xyz, int64_t n, int64_t* keys, int32_t* order)
  {
    std::vector<int32_t> idx(n);
    std::vector<int64_t> k(n);
    for (int64_t i = 0; i < n; i++) {
      idx[i] = int32_t(i);
      k[i] = pccb200::morton_addr(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }
    std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return k[a] < k[b]; });
    for (int64_t i = 0; i < n; i++) {
      keys[i] = k[idx[i]];
      order[i] = idx[i];
    }
  }
  void exclusive_scan(int* data, int64_t n)
  {
    int acc = 0;
    for (int64_t i = 0; i < n; i++) {
      const int v = data[i];
      data[i] = acc;
      acc += v;
    }
  }
  template<class P, class E>
  void compact(int64_t n, const P& pred, const E& emit, int* total = nullptr)
  {
    int64_t rank = 0;
    for (int64_t i = 0; i < n; i++)
      if (pred(i))
        emit(rank++, i);
    if (total)
      *total = int(rank);
  }
  template<class Fn>
  void subsample_distance(const Fn& fn, int nCells)
  {
    ordered(nCells, fn);
  }
  // one top-down stage, in Morton order: single-child fast path (PrepFn), the
  // thread-per-block body for the rest, then the zero-run hand-over
  template<class Fn>
  void block_stage(const Fn& fn, int64_t nBlocks, int* tzNext)
  {
    if (fn.P.n == 0) {
      fn(0);
    } else {
      foreach(nBlocks, pccb200::PrepFn{fn.cfg, fn.S, fn.P, fn.predInLvl, fn.tz});
      ordered(nBlocks, pccb200::SkipSinglesFn<Fn>{fn});
    }
    if (tzNext)
      pccb200::TzCarryFn{fn.tz, nullptr, int(nBlocks), tzNext}(0);
  }
};
