// pcc_capi.cu — the C ABI (include/pcc_attr_b200.h) over the CUDA kernels.
//
// Every entry point: validate, stage the caller's host arrays into HBM, run
// the kernels on the context's stream, copy results back, synchronise.  No
// CPU implementation exists behind this ABI: without a usable sm_100 device
// every compute entry point returns PCCB200_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "exec_cuda.cuh"
#include "lifting.cuh"
#include "lift_pipeline.cuh"
#include "lod_pipeline.cuh"
#include "morton_sort.cuh"
#include "pcc_attr_b200.h"
#include "raht_pipeline.cuh"
#include "raht_wave.cuh"
#include "spherical.cuh"
#include "symbols.cuh"
#include "dist2.cuh"
#include "recolour.cuh"

namespace pccb200 {

std::atomic<uint64_t> g_launchCount{0};

namespace {

thread_local std::string t_lastError;

// One lane = one CUDA stream with its own workspace arena, ticket word and
// event pool.  Every API call runs on one lane; calls from different host
// threads (slices, attributes, frames are independent work units,
// tmc3/encoder.cpp:545-568,1052) run on different lanes and overlap on the
// device.  Lanes are created on demand up to kMaxLanes.
struct Lane {
  cudaStream_t stream = nullptr;
  Arena arena;
  unsigned long long* ticket = nullptr;
  cudaEvent_t tail = nullptr;  // "everything queued so far" marker (pccb200_time_end)
  Profiler prof;
  bool busy = false;
};

constexpr int kMaxLanes = 32;

struct Context {
  std::mutex mu;
  std::condition_variable cv;
  int device = 0;
  bool ready = false;
  int numSMs = 0;
  std::atomic<int> active{0};
  std::vector<std::unique_ptr<Lane>> lanes;
  // profiling totals over all lanes
  bool profEnabled = false;
  double profMs[PCCB200_NUM_PHASES] = {};
  uint64_t profLaunches[PCCB200_NUM_PHASES] = {};
  // whole-device timing (pccb200_time_begin / _end)
  cudaStream_t timeStream = nullptr;
  cudaEvent_t timeBegin = nullptr, timeEnd = nullptr;
};

Context&
ctx()
{
  static Context c;
  return c;
}

int
fail(int code, const std::string& msg)
{
  t_lastError = msg;
  return code;
}

// returns PCCB200_OK or an error; must hold c.mu
int
ensure_ready(Context& c)
{
  if (c.ready)
    return PCCB200_OK;
  // (one hardware queue per lane needs CUDA_DEVICE_MAX_CONNECTIONS=32 in the
  // application's environment before its CUDA context exists: see the header)
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return fail(PCCB200_ERR_NO_DEVICE,
                std::string("no CUDA device: ") + cudaGetErrorString(e));
  }
  if (c.device >= count)
    return fail(PCCB200_ERR_NO_DEVICE, "device index out of range");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, c.device);
  if (e != cudaSuccess)
    return fail(PCCB200_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(PCCB200_ERR_NO_DEVICE,
                std::string("kernels are built for sm_100a only; device is ") + prop.name);
  e = cudaSetDevice(c.device);
  if (e != cudaSuccess)
    return fail(PCCB200_ERR_CUDA, cudaGetErrorString(e));
  c.numSMs = prop.multiProcessorCount;
  c.ready = true;
  return PCCB200_OK;
}

// must hold c.mu; creates the lane's device objects
int
make_lane(Context& c, Lane& l)
{
  cudaError_t e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking);
  if (e != cudaSuccess)
    return fail(PCCB200_ERR_CUDA, cudaGetErrorString(e));
  e = cudaMalloc(&l.ticket, 256);
  if (e != cudaSuccess)
    return fail(PCCB200_ERR_NOMEM, cudaGetErrorString(e));
  e = cudaEventCreateWithFlags(&l.tail, cudaEventDisableTiming);
  if (e != cudaSuccess)
    return fail(PCCB200_ERR_CUDA, cudaGetErrorString(e));
  return PCCB200_OK;
}

void
destroy_lanes(Context& c)
{
  for (auto& l : c.lanes) {
    cudaStreamSynchronize(l->stream);
    l->arena.release();
    cudaFree(l->ticket);
    if (l->tail)
      cudaEventDestroy(l->tail);
    cudaStreamDestroy(l->stream);
  }
  c.lanes.clear();
}

template<class T>
T*
to_device(DeviceExec& ex, const T* host, size_t n)
{
  T* d = ex.alloc<T>(n);
  if (n)
    PCC_CUDA_CHECK(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, ex.stream));
  return d;
}

template<class T>
void
to_host(DeviceExec& ex, T* host, const T* dev, size_t n)
{
  if (n)
    PCC_CUDA_CHECK(cudaMemcpyAsync(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, ex.stream));
}

// runs body(ex) on a free lane, CUDA errors mapped to a status
template<class Body>
int
with_device(Body body)
{
  Context& c = ctx();
  Lane* lane = nullptr;
  {
    std::unique_lock<std::mutex> lock(c.mu);
    int rc = ensure_ready(c);
    if (rc != PCCB200_OK)
      return rc;
    for (;;) {
      for (auto& l : c.lanes)
        if (!l->busy) {
          lane = l.get();
          break;
        }
      if (lane)
        break;
      if (int(c.lanes.size()) < kMaxLanes) {
        cudaSetDevice(c.device);
        std::unique_ptr<Lane> l(new Lane);
        rc = make_lane(c, *l);
        if (rc != PCCB200_OK)
          return rc;
        lane = l.get();
        c.lanes.push_back(std::move(l));
        break;
      }
      c.cv.wait(lock);
    }
    lane->busy = true;
    lane->prof.enabled = c.profEnabled;
    ++c.active;
  }
  int rc;
  try {
    PCC_CUDA_CHECK(cudaSetDevice(c.device));
    lane->arena.reset(c.active.load() == 1);
    DeviceExec ex;
    ex.stream = lane->stream;
    ex.arena = &lane->arena;
    ex.numSMs = c.numSMs;
    ex.ticket = lane->ticket;
    ex.prof = &lane->prof;
    ex.activeCalls = &c.active;
    rc = body(ex);
    PCC_CUDA_CHECK(cudaStreamSynchronize(lane->stream));
    lane->prof.resolve();
    if (rc != PCCB200_OK && t_lastError.empty())
      t_lastError = "invalid argument";
  } catch (const CudaError& e) {
    cudaGetLastError();
    rc = fail(e.code == cudaErrorMemoryAllocation ? PCCB200_ERR_NOMEM : PCCB200_ERR_CUDA,
              std::string(e.what) + ": " + cudaGetErrorString(e.code));
  } catch (const std::bad_alloc&) {
    rc = fail(PCCB200_ERR_NOMEM, "host allocation failed");
  } catch (const std::exception& e) {
    rc = fail(PCCB200_ERR_CUDA, std::string("internal error: ") + e.what());
  } catch (...) {
    rc = fail(PCCB200_ERR_CUDA, "internal error");
  }
  if (rc != PCCB200_OK)  // nothing of a failed call may still be running on the lane
    cudaStreamSynchronize(lane->stream);
  {
    std::lock_guard<std::mutex> lock(c.mu);
    for (int i = 0; i < PCCB200_NUM_PHASES; i++) {
      c.profMs[i] += lane->prof.ms[i];
      c.profLaunches[i] += lane->prof.launches[i];
      lane->prof.ms[i] = 0;
      lane->prof.launches[i] = 0;
    }
    lane->busy = false;
    c.active--;
  }
  c.cv.notify_one();
  return rc;
}

// runs fn(i) for i in [0, n) on up to maxThreads host threads; returns the
// first non-zero status
template<class Fn>
int
parallel_for(int n, int maxThreads, Fn fn)
{
  if (n == 1 || maxThreads <= 1) {
    for (int i = 0; i < n; i++) {
      int rc = fn(i);
      if (rc)
        return rc;
    }
    return 0;
  }
  std::atomic<int> next{0}, status{0};
  std::string err;
  std::mutex errMu;
  auto worker = [&] {
    for (;;) {
      int i = next.fetch_add(1);
      if (i >= n)
        return;
      int rc;
      try {
        rc = fn(i);
      } catch (...) {  // a worker thread must not let anything escape
        rc = fail(PCCB200_ERR_CUDA, "internal error in a slice worker");
      }
      if (rc) {
        int expected = 0;
        if (status.compare_exchange_strong(expected, rc)) {
          std::lock_guard<std::mutex> g(errMu);
          err = t_lastError;
        }
      }
    }
  };
  int nt = n < maxThreads ? n : maxThreads;
  std::vector<std::thread> ts;
  for (int t = 1; t < nt; t++)
    ts.emplace_back(worker);
  worker();
  for (auto& t : ts)
    t.join();
  if (status.load())
    t_lastError = err;
  return status.load();
}

int
raht_common(bool forward, const pccb200_raht_params* params, const pccb200_qpset* qpset,
            const int32_t* qpo, const int64_t* morton, int32_t* attrs, int A, int n,
            int32_t* coeffs)
{
  if (!params || !qpset || !morton || !attrs || !coeffs || n <= 0 || A < 1 || A > 3)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    int64_t* dKeys = to_device(ex, morton, size_t(n));
    int32_t* dAttrs = forward ? to_device(ex, attrs, size_t(n) * A)
                              : ex.alloc<int32_t>(size_t(n) * A);
    int32_t* dQpo = qpo ? to_device(ex, qpo, size_t(n) * 2) : nullptr;
    int32_t* dCoef = forward ? ex.alloc<int32_t>(size_t(n) * A)
                             : to_device(ex, coeffs, size_t(n) * A);
    int rc = raht_run(ex, *params, *qpset, forward, dKeys, dAttrs, dQpo, dCoef,
                      int64_t(n), A, n);
    if (rc != PCCB200_OK)
      return fail(rc, rc == PCCB200_ERR_UNSORTED ? "morton codes not ascending"
                                                 : "invalid parameters");
    to_host(ex, attrs, dAttrs, size_t(n) * A);
    if (forward)
      to_host(ex, coeffs, dCoef, size_t(n) * A);
    return PCCB200_OK;
  });
}

// one slice: sort, gather, transform, clip + scatter back.  All array
// pointers are device pointers addressing the slice's first point; coefficient
// component k sits at dCoef + k * coefStride.
int
attr_raht_slice(DeviceExec& ex, bool forward, const pccb200_raht_params* params,
                const pccb200_qpset* qpset, const int32_t* dQpoIn, const int32_t* dXyz,
                const int32_t* dAttrsIn, int32_t* dAttrsOut, int A, int bitdepth, int n,
                int32_t* dCoef, int64_t coefStride)
{
  int64_t* dKeys = ex.alloc<int64_t>(size_t(n));
  int32_t* dOrder = ex.alloc<int32_t>(size_t(n));
  int32_t* dAttrs = ex.alloc<int32_t>(size_t(n) * A);
  int32_t* dQpo = dQpoIn ? ex.alloc<int32_t>(size_t(n) * 2) : nullptr;
  const int32_t clipMax = (1 << bitdepth) - 1;
  device_morton_sort(ex, dXyz, n, dKeys, dOrder);
  const unsigned g = grid_for(n, ex.numSMs);
  ex.phase(kPhaseGather);
  if (forward) {
    DeviceExec::Scope sc(ex);
    k_gather_rows<int32_t><<<g, 256, 0, ex.stream>>>(dAttrsIn, dOrder, n, A, dAttrs);
    g_launchCount++;
  }
  if (dQpoIn) {
    DeviceExec::Scope sc(ex);
    k_gather_rows<int32_t><<<g, 256, 0, ex.stream>>>(dQpoIn, dOrder, n, 2, dQpo);
    g_launchCount++;
  }
  int rc = raht_run(ex, *params, *qpset, forward, dKeys, dAttrs, dQpo, dCoef, coefStride, A, n);
  if (rc != PCCB200_OK)
    return fail(rc, "invalid parameters");
  ex.phase(kPhaseGather);
  {
    DeviceExec::Scope sc(ex);
    k_scatter_rows_clip<<<g, 256, 0, ex.stream>>>(dAttrs, dOrder, n, A, clipMax, dAttrsOut);
    g_launchCount++;
  }
  PCC_CUDA_CHECK(cudaGetLastError());
  return PCCB200_OK;
}

// one slice, several attributes on the same positions in one pass: one sort,
// one tree, one dependency chain.  Pointers as in attr_raht_slice, one per set.
// In two halves, so that the descents of several slices can be issued together
// (a gang, attr_raht_batch_common): _begin sorts, gathers, builds the tree and
// either runs the descent or leaves it prepared in ms.defer; _end runs the
// duplicate tail (if deferred) and scatters the reconstruction back.
// PCCB200_ERR_UNSUPPORTED: this parameter combination has no fused path.
struct MultiSlice {
  int32_t* dOrder = nullptr;
  int32_t* dAttrs = nullptr;
  int AT = 0;
  RahtDeferred<DeviceExec> defer;
};

int
attr_raht_slice_multi_begin(DeviceExec& ex, bool forward, const pccb200_raht_params* params,
                            int numSets, const pccb200_qpset* const* qpsets, const int32_t* dXyz,
                            const int32_t* const* dAttrsIn, const int* A, int n,
                            int32_t* const* dCoef, const int64_t* coefStride, MultiSlice& ms,
                            bool deferDescent)
{
  int AT = 0;
  for (int s = 0; s < numSets; s++)
    AT += A[s];
  int64_t* dKeys = ex.alloc<int64_t>(size_t(n));
  ms.dOrder = ex.alloc<int32_t>(size_t(n));
  ms.dAttrs = ex.alloc<int32_t>(size_t(n) * AT);
  ms.AT = AT;
  device_morton_sort(ex, dXyz, n, dKeys, ms.dOrder);
  const unsigned g = grid_for(n, ex.numSMs);
  ex.phase(kPhaseGather);
  RahtSetIO io[2];
  for (int s = 0, base = 0; s < numSets; base += A[s], s++) {
    if (forward) {
      DeviceExec::Scope sc(ex);
      k_gather_rows_strided<<<g, 256, 0, ex.stream>>>(dAttrsIn[s], ms.dOrder, n, A[s], ms.dAttrs, AT, base);
      g_launchCount++;
    }
    io[s] = RahtSetIO{qpsets[s], A[s], dCoef[s], coefStride[s]};
  }
  int rc = raht_run_sets(ex, *params, numSets, io, forward, dKeys, ms.dAttrs, nullptr, n,
                         deferDescent ? &ms.defer : nullptr);
  if (rc != PCCB200_OK)
    return rc == PCCB200_ERR_UNSUPPORTED ? rc : fail(rc, "invalid parameters");
  return PCCB200_OK;
}

void
attr_raht_slice_multi_end(DeviceExec& ex, int numSets, const int* A, const int* bitdepth, int n,
                          int32_t* const* dAttrsOut, MultiSlice& ms)
{
  if (ms.defer.pending) {
    ex.phase(kPhaseTail);
    ex.foreach(ms.defer.nLeaves, ms.defer.tail);
    ms.defer.pending = false;
  }
  const unsigned g = grid_for(n, ex.numSMs);
  ex.phase(kPhaseGather);
  for (int s = 0, base = 0; s < numSets; base += A[s], s++) {
    DeviceExec::Scope sc(ex);
    k_scatter_rows_clip_strided<<<g, 256, 0, ex.stream>>>(ms.dAttrs, ms.AT, base, ms.dOrder, n, A[s],
                                                          (1 << bitdepth[s]) - 1, dAttrsOut[s]);
    g_launchCount++;
  }
  PCC_CUDA_CHECK(cudaGetLastError());
}

int
attr_raht_slice_multi(DeviceExec& ex, bool forward, const pccb200_raht_params* params,
                      int numSets, const pccb200_qpset* const* qpsets, const int32_t* dXyz,
                      const int32_t* const* dAttrsIn, int32_t* const* dAttrsOut, const int* A,
                      const int* bitdepth, int n, int32_t* const* dCoef, const int64_t* coefStride)
{
  MultiSlice ms;
  int rc = attr_raht_slice_multi_begin(ex, forward, params, numSets, qpsets, dXyz, dAttrsIn, A, n,
                                       dCoef, coefStride, ms, false);
  if (rc != PCCB200_OK)
    return rc;
  attr_raht_slice_multi_end(ex, numSets, A, bitdepth, n, dAttrsOut, ms);
  return PCCB200_OK;
}

int
check_slices(const void* params, const void* qpset, const void* xyz, const void* attrs,
             const void* coeffs, int A, int bitdepth, const int64_t* sliceOffsets,
             int numSlices)
{
  if (!params || !qpset || !xyz || !attrs || !coeffs || !sliceOffsets || numSlices <= 0
      || A < 1 || A > 3 || bitdepth < 1 || bitdepth > 16)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  for (int s = 0; s < numSlices; s++) {
    int64_t len = sliceOffsets[s + 1] - sliceOffsets[s];
    if (len <= 0 || len > INT32_MAX)
      return fail(PCCB200_ERR_INVALID_ARG, "empty or oversized slice");
  }
  return PCCB200_OK;
}

constexpr int kMaxSliceThreads = 16;

// host pointers: every slice is staged, transformed and returned on its own
// lane; slices overlap on the device
int
attr_raht_common(bool forward, const pccb200_raht_params* params, const pccb200_qpset* qpset,
                 const int32_t* qpo, const int32_t* xyz, int32_t* attrs, int A,
                 int bitdepth, const int64_t* sliceOffsets, int numSlices,
                 int32_t* coeffs)
{
  int rc = check_slices(params, qpset, xyz, attrs, coeffs, A, bitdepth, sliceOffsets, numSlices);
  if (rc != PCCB200_OK)
    return rc;
  const int64_t total = sliceOffsets[numSlices];
  return parallel_for(numSlices, kMaxSliceThreads, [&](int s) -> int {
    const int64_t o = sliceOffsets[s];
    const int n = int(sliceOffsets[s + 1] - o);
    return with_device([&](DeviceExec& ex) -> int {
      int32_t* dXyz = to_device(ex, xyz + 3 * o, size_t(n) * 3);
      int32_t* dAttrsIn = forward ? to_device(ex, attrs + o * A, size_t(n) * A) : nullptr;
      int32_t* dQpoIn = qpo ? to_device(ex, qpo + 2 * o, size_t(n) * 2) : nullptr;
      int32_t* dCoef = ex.alloc<int32_t>(size_t(n) * A);
      if (!forward)
        for (int k = 0; k < A; k++)
          PCC_CUDA_CHECK(cudaMemcpyAsync(dCoef + size_t(k) * n, coeffs + k * total + o,
                                         size_t(n) * sizeof(int32_t), cudaMemcpyHostToDevice,
                                         ex.stream));
      int32_t* dOut = ex.alloc<int32_t>(size_t(n) * A);
      int rc2 = attr_raht_slice(ex, forward, params, qpset, dQpoIn, dXyz, dAttrsIn, dOut, A,
                                bitdepth, n, dCoef, n);
      if (rc2 != PCCB200_OK)
        return rc2;
      to_host(ex, attrs + o * A, dOut, size_t(n) * A);
      if (forward)
        for (int k = 0; k < A; k++)
          to_host(ex, coeffs + k * total + o, dCoef + size_t(k) * n, size_t(n));
      return PCCB200_OK;
    });
  });
}

int
attr_raht_common_dev(bool forward, const pccb200_raht_params* params,
                     const pccb200_qpset* qpset, const int32_t* dQpo, const int32_t* dXyz,
                     int32_t* dAttrs, int A, int bitdepth, const int64_t* sliceOffsets,
                     int numSlices, int32_t* dCoef)
{
  int rc = check_slices(params, qpset, dXyz, dAttrs, dCoef, A, bitdepth, sliceOffsets, numSlices);
  if (rc != PCCB200_OK)
    return rc;
  const int64_t total = sliceOffsets[numSlices];
  return parallel_for(numSlices, kMaxSliceThreads, [&](int s) -> int {
    const int64_t o = sliceOffsets[s];
    const int n = int(sliceOffsets[s + 1] - o);
    return with_device([&](DeviceExec& ex) -> int {
      // in place on the caller's attribute buffer: the gather reads the slice
      // before the final scatter overwrites it
      return attr_raht_slice(ex, forward, params, qpset, dQpo ? dQpo + 2 * o : nullptr,
                             dXyz + 3 * o, dAttrs + o * A, dAttrs + o * A, A, bitdepth, n,
                             dCoef + o, total);
    });
  });
}

int
check_multi(const void* params, int numSets, const pccb200_qpset* const* qpsets, const void* xyz,
            const void* const* attrs, const int32_t* A, const int32_t* bitdepth,
            const void* const* coeffs, int64_t n)
{
  if (!params || !qpsets || !xyz || !attrs || !A || !bitdepth || !coeffs || n <= 0
      || n > INT32_MAX || numSets < 1 || numSets > 2)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  int total = 0;
  for (int s = 0; s < numSets; s++) {
    if (!qpsets[s] || !attrs[s] || !coeffs[s] || A[s] < 1 || A[s] > 3 || bitdepth[s] < 1
        || bitdepth[s] > 16)
      return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad attribute description");
    total += A[s];
  }
  if (total > 4)
    return fail(PCCB200_ERR_INVALID_ARG, "more than four components in one pass");
  return PCCB200_OK;
}

// Several attributes of one slice.  Fused pass where the parameters allow it,
// otherwise one pass per attribute (same results either way).
int
attr_raht_multi_common(bool forward, bool device, const pccb200_raht_params* params, int numSets,
                       const pccb200_qpset* const* qpsets, const int32_t* xyz,
                       int32_t* const* attrs, const int32_t* A, const int32_t* bitdepth, int n,
                       int32_t* const* coeffs)
{
  int rc = check_multi(params, numSets, qpsets, xyz, reinterpret_cast<const void* const*>(attrs),
                       A, bitdepth, reinterpret_cast<const void* const*>(coeffs), n);
  if (rc != PCCB200_OK)
    return rc;
  rc = with_device([&](DeviceExec& ex) -> int {
    const int32_t* dXyz = device ? xyz : to_device(ex, xyz, size_t(n) * 3);
    int32_t* dIn[2];
    int32_t* dOut[2];
    int32_t* dCoef[2];
    int64_t stride[2];
    for (int s = 0; s < numSets; s++) {
      stride[s] = n;
      if (device) {
        dIn[s] = dOut[s] = attrs[s];
        dCoef[s] = coeffs[s];
      } else {
        dIn[s] = forward ? to_device(ex, attrs[s], size_t(n) * A[s]) : nullptr;
        dOut[s] = ex.alloc<int32_t>(size_t(n) * A[s]);
        dCoef[s] = forward ? ex.alloc<int32_t>(size_t(n) * A[s])
                           : to_device(ex, coeffs[s], size_t(n) * A[s]);
      }
    }
    int rc2 = attr_raht_slice_multi(ex, forward, params, numSets, qpsets, dXyz, dIn, dOut, A,
                                    bitdepth, n, dCoef, stride);
    if (rc2 != PCCB200_OK)
      return rc2;
    if (!device)
      for (int s = 0; s < numSets; s++) {
        to_host(ex, attrs[s], dOut[s], size_t(n) * A[s]);
        if (forward)
          to_host(ex, coeffs[s], dCoef[s], size_t(n) * A[s]);
      }
    return PCCB200_OK;
  });
  if (rc != PCCB200_ERR_UNSUPPORTED)
    return rc;
  // no fused path for these parameters: attribute by attribute
  for (int s = 0; s < numSets; s++) {
    const int64_t offs[2] = {0, n};
    rc = device ? attr_raht_common_dev(forward, params, qpsets[s], nullptr, xyz, attrs[s], A[s],
                                       bitdepth[s], offs, 1, coeffs[s])
                : attr_raht_common(forward, params, qpsets[s], nullptr, xyz, attrs[s], A[s],
                                   bitdepth[s], offs, 1, coeffs[s]);
    if (rc != PCCB200_OK)
      return rc;
  }
  return PCCB200_OK;
}

// Many coding units (the slices of a frame, or whole frames: independent
// point sets, each with the same attributes) in one call.  The units are dealt
// to the lanes in gangs: a lane sorts / builds the tree of each unit of its
// gang in turn and then issues the top-down passes of all of them together,
// one launch per descent step (WaveDescent::run_gang).  A textured unit keeps
// only a few warps busy (its blocks form one chain): the number of chains in
// flight is what the device's throughput follows, and with gangs it is no
// longer limited by the number of hardware queues.
constexpr int kMaxGang = 32;
constexpr int kBatchLanes = 16;  // lanes a batch call spreads over (measured: 8, 16 and 32 lanes
                                 // give the same throughput; fewer lanes = fewer host threads
                                 // and fewer launches)

int
attr_raht_batch_common(bool forward, bool device, const pccb200_raht_params* params, int numSets,
                       const pccb200_qpset* const* qpsets, int numUnits,
                       const int32_t* const* xyz, int32_t* const* attrs, const int32_t* A,
                       const int32_t* bitdepth, const int32_t* n, int32_t* const* coeffs)
{
  if (!xyz || !attrs || !coeffs || !n || numUnits <= 0)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  for (int u = 0; u < numUnits; u++) {
    if (!xyz[u])
      return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
    int rc = check_multi(params, numSets, qpsets, xyz[u],
                         reinterpret_cast<const void* const*>(attrs + size_t(u) * numSets), A,
                         bitdepth, reinterpret_cast<const void* const*>(coeffs + size_t(u) * numSets),
                         n[u]);
    if (rc != PCCB200_OK)
      return rc;
  }
  // the fused pass exists for these parameters?  (same test as raht_run_sets)
  bool fused = true;
  if (numSets > 1)
    for (int u = 0; u < numUnits && fused; u++) {
      fused = n[u] >= 2;
      for (int s = 0; s < numSets && fused; s++)
        fused = WaveDescent<DeviceExec>::enabled(make_config(*params, *qpsets[s], forward, A[s], false));
    }
  if (!fused)  // unit by unit, attribute by attribute (results do not depend on the route)
    return parallel_for(numUnits, kMaxSliceThreads, [&](int u) -> int {
      return attr_raht_multi_common(forward, device, params, numSets, qpsets, xyz[u],
                                    attrs + size_t(u) * numSets, A, bitdepth, n[u],
                                    coeffs + size_t(u) * numSets);
    });
  // units per gang: by default the units are spread over all lanes first
  // (PCCB200_GANG: A/B knob, read per call)
  const char* eg = getenv("PCCB200_GANG");
  const int envGang = eg ? atoi(eg) : 0;
  int gang = envGang > 0 ? envGang : (numUnits + kBatchLanes - 1) / kBatchLanes;
  gang = gang > kMaxGang ? kMaxGang : gang;
  const int numGangs = (numUnits + gang - 1) / gang;
  return parallel_for(numGangs, kMaxLanes, [&](int g) -> int {
    const int u0 = g * gang;
    const int u1 = u0 + gang < numUnits ? u0 + gang : numUnits;
    const int m = u1 - u0;
    return with_device([&](DeviceExec& ex) -> int {
      std::vector<MultiSlice> ms(m);
      std::vector<int32_t*> dOut(size_t(m) * 2), dCoef(size_t(m) * 2);
      for (int i = 0; i < m; i++) {
        const int u = u0 + i;
        const int nu = n[u];
        const int32_t* dXyz = device ? xyz[u] : to_device(ex, xyz[u], size_t(nu) * 3);
        int32_t* dIn[2] = {nullptr, nullptr};
        int64_t stride[2] = {nu, nu};
        for (int s = 0; s < numSets; s++) {
          int32_t* a = attrs[size_t(u) * numSets + s];
          int32_t* c = coeffs[size_t(u) * numSets + s];
          if (device) {
            dIn[s] = dOut[2 * i + s] = a;
            dCoef[2 * i + s] = c;
          } else {
            dIn[s] = forward ? to_device(ex, a, size_t(nu) * A[s]) : nullptr;
            dOut[2 * i + s] = ex.alloc<int32_t>(size_t(nu) * A[s]);
            dCoef[2 * i + s] = forward ? ex.alloc<int32_t>(size_t(nu) * A[s])
                                       : to_device(ex, c, size_t(nu) * A[s]);
          }
        }
        int rc2 = attr_raht_slice_multi_begin(ex, forward, params, numSets, qpsets, dXyz, dIn, A, nu,
                                              &dCoef[2 * i], stride, ms[i], true);
        if (rc2 != PCCB200_OK)
          return rc2;
      }
      std::vector<WaveDescent<DeviceExec>::Job*> jobs;
      for (int i = 0; i < m; i++)
        if (ms[i].defer.pending)
          jobs.push_back(&ms[i].defer.job);
      if (!jobs.empty())
        WaveDescent<DeviceExec>::run_gang(ex, jobs.data(), int(jobs.size()));
      for (int i = 0; i < m; i++) {
        const int u = u0 + i;
        attr_raht_slice_multi_end(ex, numSets, A, bitdepth, n[u], &dOut[2 * i], ms[i]);
        if (!device)
          for (int s = 0; s < numSets; s++) {
            to_host(ex, attrs[size_t(u) * numSets + s], dOut[2 * i + s], size_t(n[u]) * A[s]);
            if (forward)
              to_host(ex, coeffs[size_t(u) * numSets + s], dCoef[2 * i + s], size_t(n[u]) * A[s]);
          }
      }
      return PCCB200_OK;
    });
  });
}

}  // namespace
}  // namespace pccb200

using namespace pccb200;

extern "C" {

void
pccb200_raht_set_prediction_weights(pccb200_raht_params* p, const int32_t w[5])
{
  const int child[12] = {4, 4, 3, 4, 3, 3, 4, 4, 4, 4, 4, 4};
  const int parent[19] = {0, 1, 1, 1, 2, 2, 2, 2, 2, 1, 2, 1, 1, 2, 2, 2, 2, 2, 2};
  for (int i = 0; i < 12; i++)
    p->pred_weight_child[i] = w[child[i]];
  for (int i = 0; i < 19; i++)
    p->pred_weight_parent[i] = w[parent[i]];
}

void
pccb200_raht_params_default(pccb200_raht_params* p)
{
  p->prediction_enabled = 1;
  p->integer_haar = 0;
  p->prediction_threshold0 = 2;
  p->prediction_threshold1 = 6;
  p->subnode_prediction_enabled = 1;
  p->prediction_search_range = 50000;
  p->raht_extension = 1;
  const int32_t w[5] = {9, 3, 1, 5, 2};
  pccb200_raht_set_prediction_weights(p, w);
}

int
pccb200_abi_version(void)
{
  return PCCB200_ABI_VERSION;
}

int
pccb200_set_device(int device)
{
  Context& c = ctx();
  std::lock_guard<std::mutex> lock(c.mu);
  if (device < 0)
    return fail(PCCB200_ERR_INVALID_ARG, "negative device index");
  if (c.ready && device != c.device) {
    if (c.active.load())
      return fail(PCCB200_ERR_INVALID_ARG, "calls in flight on the current device");
    cudaSetDevice(c.device);
    destroy_lanes(c);
    if (c.timeStream) {
      cudaEventDestroy(c.timeBegin);
      cudaEventDestroy(c.timeEnd);
      cudaStreamDestroy(c.timeStream);
      c.timeStream = nullptr;
    }
    c.ready = false;
  }
  c.device = device;
  return ensure_ready(c);
}

const char*
pccb200_last_error(void)
{
  return t_lastError.c_str();
}

uint64_t
pccb200_kernel_launch_count(void)
{
  return g_launchCount.load();
}

int
pccb200_morton_sort(const int32_t* xyz, int32_t n, int64_t* keys_out, int32_t* order_out)
{
  if (!xyz || !keys_out || !order_out || n <= 0)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    int32_t* dXyz = to_device(ex, xyz, size_t(n) * 3);
    int64_t* dKeys = ex.alloc<int64_t>(n);
    int32_t* dOrder = ex.alloc<int32_t>(n);
    device_morton_sort(ex, dXyz, n, dKeys, dOrder);
    to_host(ex, keys_out, dKeys, size_t(n));
    to_host(ex, order_out, dOrder, size_t(n));
    return PCCB200_OK;
  });
}

int
pccb200_raht_forward(const pccb200_raht_params* params, const pccb200_qpset* qpset,
                     const int32_t* point_qp_offsets, const int64_t* morton,
                     int32_t* attrs_inout, int32_t num_attrs, int32_t n,
                     int32_t* coeffs_out)
{
  return raht_common(true, params, qpset, point_qp_offsets, morton, attrs_inout, num_attrs, n,
                     coeffs_out);
}

int
pccb200_raht_inverse(const pccb200_raht_params* params, const pccb200_qpset* qpset,
                     const int32_t* point_qp_offsets, const int64_t* morton,
                     int32_t* attrs_out, int32_t num_attrs, int32_t n,
                     const int32_t* coeffs_in)
{
  return raht_common(false, params, qpset, point_qp_offsets, morton, attrs_out, num_attrs, n,
                     const_cast<int32_t*>(coeffs_in));
}

int
pccb200_attr_raht_encode(const pccb200_raht_params* params, const pccb200_qpset* qpset,
                         const int32_t* point_qp_offsets, const int32_t* xyz,
                         int32_t* attrs_inout, int32_t num_attrs, int32_t n, int32_t bitdepth,
                         int32_t* coeffs_out)
{
  const int64_t offs[2] = {0, n};
  return attr_raht_common(true, params, qpset, point_qp_offsets, xyz, attrs_inout, num_attrs,
                          bitdepth, offs, 1, coeffs_out);
}

int
pccb200_attr_raht_decode(const pccb200_raht_params* params, const pccb200_qpset* qpset,
                         const int32_t* point_qp_offsets, const int32_t* xyz,
                         int32_t* attrs_out, int32_t num_attrs, int32_t n, int32_t bitdepth,
                         const int32_t* coeffs_in)
{
  const int64_t offs[2] = {0, n};
  return attr_raht_common(false, params, qpset, point_qp_offsets, xyz, attrs_out, num_attrs,
                          bitdepth, offs, 1, const_cast<int32_t*>(coeffs_in));
}

int
pccb200_attr_raht_encode_slices(const pccb200_raht_params* params, const pccb200_qpset* qpset,
                                const int32_t* point_qp_offsets, const int32_t* xyz,
                                int32_t* attrs_inout, int32_t num_attrs, int32_t bitdepth,
                                const int64_t* slice_offsets, int32_t num_slices,
                                int32_t* coeffs_out)
{
  return attr_raht_common(true, params, qpset, point_qp_offsets, xyz, attrs_inout, num_attrs,
                          bitdepth, slice_offsets, num_slices, coeffs_out);
}

int
pccb200_attr_raht_encode_slices_dev(const pccb200_raht_params* params,
                                    const pccb200_qpset* qpset,
                                    const int32_t* d_point_qp_offsets, const int32_t* d_xyz,
                                    int32_t* d_attrs_inout, int32_t num_attrs, int32_t bitdepth,
                                    const int64_t* slice_offsets, int32_t num_slices,
                                    int32_t* d_coeffs_out)
{
  return attr_raht_common_dev(true, params, qpset, d_point_qp_offsets, d_xyz, d_attrs_inout,
                              num_attrs, bitdepth, slice_offsets, num_slices, d_coeffs_out);
}

int
pccb200_attr_raht_decode_slices_dev(const pccb200_raht_params* params,
                                    const pccb200_qpset* qpset,
                                    const int32_t* d_point_qp_offsets, const int32_t* d_xyz,
                                    int32_t* d_attrs_out, int32_t num_attrs, int32_t bitdepth,
                                    const int64_t* slice_offsets, int32_t num_slices,
                                    const int32_t* d_coeffs_in)
{
  return attr_raht_common_dev(false, params, qpset, d_point_qp_offsets, d_xyz, d_attrs_out,
                              num_attrs, bitdepth, slice_offsets, num_slices,
                              const_cast<int32_t*>(d_coeffs_in));
}

int
pccb200_attr_raht_encode_multi(const pccb200_raht_params* params, int32_t num_sets,
                               const pccb200_qpset* const* qpsets, const int32_t* xyz,
                               int32_t* const* attrs_inout, const int32_t* num_attrs,
                               const int32_t* bitdepths, int32_t n, int32_t* const* coeffs_out)
{
  return attr_raht_multi_common(true, false, params, num_sets, qpsets, xyz, attrs_inout,
                                num_attrs, bitdepths, n, coeffs_out);
}

int
pccb200_attr_raht_decode_multi(const pccb200_raht_params* params, int32_t num_sets,
                               const pccb200_qpset* const* qpsets, const int32_t* xyz,
                               int32_t* const* attrs_out, const int32_t* num_attrs,
                               const int32_t* bitdepths, int32_t n,
                               const int32_t* const* coeffs_in)
{
  return attr_raht_multi_common(false, false, params, num_sets, qpsets, xyz, attrs_out,
                                num_attrs, bitdepths, n,
                                const_cast<int32_t* const*>(coeffs_in));
}

int
pccb200_attr_raht_encode_multi_dev(const pccb200_raht_params* params, int32_t num_sets,
                                   const pccb200_qpset* const* qpsets, const int32_t* d_xyz,
                                   int32_t* const* d_attrs_inout, const int32_t* num_attrs,
                                   const int32_t* bitdepths, int32_t n,
                                   int32_t* const* d_coeffs_out)
{
  return attr_raht_multi_common(true, true, params, num_sets, qpsets, d_xyz, d_attrs_inout,
                                num_attrs, bitdepths, n, d_coeffs_out);
}

int
pccb200_attr_raht_decode_multi_dev(const pccb200_raht_params* params, int32_t num_sets,
                                   const pccb200_qpset* const* qpsets, const int32_t* d_xyz,
                                   int32_t* const* d_attrs_out, const int32_t* num_attrs,
                                   const int32_t* bitdepths, int32_t n,
                                   const int32_t* const* d_coeffs_in)
{
  return attr_raht_multi_common(false, true, params, num_sets, qpsets, d_xyz, d_attrs_out,
                                num_attrs, bitdepths, n,
                                const_cast<int32_t* const*>(d_coeffs_in));
}

int
pccb200_attr_raht_encode_multi_batch(const pccb200_raht_params* params, int32_t num_sets,
                                     const pccb200_qpset* const* qpsets, int32_t num_units,
                                     const int32_t* const* xyz, int32_t* const* attrs_inout,
                                     const int32_t* num_attrs, const int32_t* bitdepths,
                                     const int32_t* n, int32_t* const* coeffs_out)
{
  return attr_raht_batch_common(true, false, params, num_sets, qpsets, num_units, xyz,
                                attrs_inout, num_attrs, bitdepths, n, coeffs_out);
}

int
pccb200_attr_raht_decode_multi_batch(const pccb200_raht_params* params, int32_t num_sets,
                                     const pccb200_qpset* const* qpsets, int32_t num_units,
                                     const int32_t* const* xyz, int32_t* const* attrs_out,
                                     const int32_t* num_attrs, const int32_t* bitdepths,
                                     const int32_t* n, const int32_t* const* coeffs_in)
{
  return attr_raht_batch_common(false, false, params, num_sets, qpsets, num_units, xyz, attrs_out,
                                num_attrs, bitdepths, n, const_cast<int32_t* const*>(coeffs_in));
}

int
pccb200_attr_raht_encode_multi_batch_dev(const pccb200_raht_params* params, int32_t num_sets,
                                         const pccb200_qpset* const* qpsets, int32_t num_units,
                                         const int32_t* const* d_xyz,
                                         int32_t* const* d_attrs_inout, const int32_t* num_attrs,
                                         const int32_t* bitdepths, const int32_t* n,
                                         int32_t* const* d_coeffs_out)
{
  return attr_raht_batch_common(true, true, params, num_sets, qpsets, num_units, d_xyz,
                                d_attrs_inout, num_attrs, bitdepths, n, d_coeffs_out);
}

int
pccb200_attr_raht_decode_multi_batch_dev(const pccb200_raht_params* params, int32_t num_sets,
                                         const pccb200_qpset* const* qpsets, int32_t num_units,
                                         const int32_t* const* d_xyz, int32_t* const* d_attrs_out,
                                         const int32_t* num_attrs, const int32_t* bitdepths,
                                         const int32_t* n, const int32_t* const* d_coeffs_in)
{
  return attr_raht_batch_common(false, true, params, num_sets, qpsets, num_units, d_xyz,
                                d_attrs_out, num_attrs, bitdepths, n,
                                const_cast<int32_t* const*>(d_coeffs_in));
}


// Device-side timing across all lanes: begin() records an event that every
// lane's stream waits for; end() records one event that waits for every
// lane's stream and returns the elapsed milliseconds between the two.
int
pccb200_time_begin(void)
{
  Context& c = ctx();
  std::lock_guard<std::mutex> lock(c.mu);
  int rc = ensure_ready(c);
  if (rc != PCCB200_OK)
    return rc;
  cudaSetDevice(c.device);
  if (!c.timeStream) {
    cudaStreamCreateWithFlags(&c.timeStream, cudaStreamNonBlocking);
    cudaEventCreate(&c.timeBegin);
    cudaEventCreate(&c.timeEnd);
  }
  cudaEventRecord(c.timeBegin, c.timeStream);
  for (auto& l : c.lanes)
    cudaStreamWaitEvent(l->stream, c.timeBegin, 0);
  return PCCB200_OK;
}

int
pccb200_time_end(double* ms_out)
{
  Context& c = ctx();
  std::lock_guard<std::mutex> lock(c.mu);
  if (!c.timeStream || !ms_out)
    return fail(PCCB200_ERR_INVALID_ARG, "pccb200_time_begin not called");
  cudaSetDevice(c.device);
  for (auto& l : c.lanes) {
    cudaEventRecord(l->tail, l->stream);
    cudaStreamWaitEvent(c.timeStream, l->tail, 0);
  }
  cudaEventRecord(c.timeEnd, c.timeStream);
  cudaEventSynchronize(c.timeEnd);
  float ms = 0;
  if (cudaEventElapsedTime(&ms, c.timeBegin, c.timeEnd) != cudaSuccess)
    return fail(PCCB200_ERR_CUDA, "cudaEventElapsedTime failed");
  *ms_out = ms;
  return PCCB200_OK;
}

void
pccb200_profile_enable(int enable)
{
  Context& c = ctx();
  std::lock_guard<std::mutex> lock(c.mu);
  c.profEnabled = enable != 0;
}

void
pccb200_profile_reset(void)
{
  Context& c = ctx();
  std::lock_guard<std::mutex> lock(c.mu);
  for (int i = 0; i < PCCB200_NUM_PHASES; i++) {
    c.profMs[i] = 0;
    c.profLaunches[i] = 0;
  }
}

void
pccb200_profile_read(double ms_out[PCCB200_NUM_PHASES], uint64_t launches_out[PCCB200_NUM_PHASES])
{
  Context& c = ctx();
  std::lock_guard<std::mutex> lock(c.mu);
  for (int i = 0; i < PCCB200_NUM_PHASES; i++) {
    ms_out[i] = c.profMs[i];
    launches_out[i] = c.profLaunches[i];
  }
}

int
pccb200_lod_build(const pccb200_lod_params* params, const int32_t* xyz, int32_t n,
                  pccb200_predictor* preds_out, uint32_t* indexes_out,
                  uint32_t* num_points_in_lod_out, int32_t* lod_count_out)
{
  if (!params || !xyz || !preds_out || !indexes_out || !num_points_in_lod_out || !lod_count_out
      || n <= 0)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    int32_t* dXyz = to_device(ex, xyz, size_t(n) * 3);
    pccb200_predictor* dP = ex.alloc<pccb200_predictor>(n);
    uint32_t* dIdx = ex.alloc<uint32_t>(n);
    int cnt = 0;
    int rc = lod_run(ex, *params, dXyz, n, dP, dIdx, num_points_in_lod_out, &cnt);
    if (rc != PCCB200_OK)
      return fail(rc, "invalid LoD parameters");
    *lod_count_out = cnt;
    to_host(ex, preds_out, dP, size_t(n));
    to_host(ex, indexes_out, dIdx, size_t(n));
    return PCCB200_OK;
  });
}

int
pccb200_quant_weights(const pccb200_predictor* preds, int32_t n,
                      const uint32_t* num_points_in_lod, int32_t lod_count, uint64_t* qw_out)
{
  if (!preds || !num_points_in_lod || !qw_out || n <= 0 || lod_count <= 0)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    pccb200_predictor* dP = to_device(ex, preds, size_t(n));
    uint64_t* dQw = ex.alloc<uint64_t>(n);
    int rc = run_quant_weights(ex, dP, n, num_points_in_lod, lod_count, dQw);
    if (rc != PCCB200_OK)
      return fail(rc, "numPointsInLod does not partition [0, n)");
    to_host(ex, qw_out, dQw, size_t(n));
    return PCCB200_OK;
  });
}

static int
lift_common(bool forward, const pccb200_predictor* preds, const uint64_t* qw, int32_t n,
            const uint32_t* num_points_in_lod, int32_t lod_count, int64_t* attrs, int32_t A)
{
  if (!preds || !qw || !num_points_in_lod || !attrs || n <= 0 || lod_count <= 0
      || (A != 1 && A != 3))
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    pccb200_predictor* dP = to_device(ex, preds, size_t(n));
    uint64_t* dQw = to_device(ex, qw, size_t(n));
    int64_t* dA = to_device(ex, attrs, size_t(n) * A);
    int rc = run_lift(ex, forward, dP, dQw, n, num_points_in_lod, lod_count, dA, A);
    if (rc != PCCB200_OK)
      return fail(rc, rc == PCCB200_ERR_UNSUPPORTED
                        ? "a predictor references its own level of detail"
                        : "numPointsInLod does not partition [0, n)");
    to_host(ex, attrs, dA, size_t(n) * A);
    return PCCB200_OK;
  });
}

int
pccb200_lift_forward(const pccb200_predictor* preds, const uint64_t* qw, int32_t n,
                     const uint32_t* num_points_in_lod, int32_t lod_count,
                     int64_t* attrs_inout, int32_t num_attrs)
{
  return lift_common(true, preds, qw, n, num_points_in_lod, lod_count, attrs_inout, num_attrs);
}

int
pccb200_lift_inverse(const pccb200_predictor* preds, const uint64_t* qw, int32_t n,
                     const uint32_t* num_points_in_lod, int32_t lod_count,
                     int64_t* attrs_inout, int32_t num_attrs)
{
  return lift_common(false, preds, qw, n, num_points_in_lod, lod_count, attrs_inout, num_attrs);
}

static int
lift_quant_common(bool forward, const pccb200_qpset* qpset, const int32_t* qpo,
                  const uint64_t* qw, int32_t n, const uint32_t* npl, int32_t lodCount,
                  int32_t numDetailLevels, int64_t* attrs, int32_t A, int32_t lcpEnabled,
                  int32_t* values, int8_t* lcp)
{
  if (!qpset || !qw || !npl || !attrs || !values || n <= 0 || (lcpEnabled && A == 3 && !lcp))
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    uint64_t* dQw = to_device(ex, qw, size_t(n));
    int32_t* dQpo = qpo ? to_device(ex, qpo, size_t(n) * 2) : nullptr;
    int64_t* dA = forward ? to_device(ex, attrs, size_t(n) * A) : ex.alloc<int64_t>(size_t(n) * A);
    int32_t* dV = forward ? ex.alloc<int32_t>(size_t(n) * A) : to_device(ex, values, size_t(n) * A);
    int rc = run_lift_quant(ex, forward, *qpset, dQpo, dQw, n, npl, lodCount, numDetailLevels, dA,
                            A, lcpEnabled != 0, lcp, dV);
    if (rc != PCCB200_OK)
      return fail(rc, "invalid lifting quantisation parameters");
    to_host(ex, attrs, dA, size_t(n) * A);
    if (forward)
      to_host(ex, values, dV, size_t(n) * A);
    return PCCB200_OK;
  });
}

int
pccb200_lift_quantize(const pccb200_qpset* qpset, const int32_t* point_qp_offsets,
                      const uint64_t* qw, int32_t n, const uint32_t* num_points_in_lod,
                      int32_t lod_count, int32_t num_detail_levels, int64_t* attrs_inout,
                      int32_t num_attrs, int32_t lcp_enabled, int32_t* values_out,
                      int8_t* lcp_coeffs_out)
{
  return lift_quant_common(true, qpset, point_qp_offsets, qw, n, num_points_in_lod, lod_count,
                           num_detail_levels, attrs_inout, num_attrs, lcp_enabled, values_out,
                           lcp_coeffs_out);
}

int
pccb200_lift_dequantize(const pccb200_qpset* qpset, const int32_t* point_qp_offsets,
                        const uint64_t* qw, int32_t n, const uint32_t* num_points_in_lod,
                        int32_t lod_count, int32_t num_detail_levels, const int32_t* values_in,
                        int32_t num_attrs, const int8_t* lcp_coeffs, int64_t* attrs_out)
{
  return lift_quant_common(false, qpset, point_qp_offsets, qw, n, num_points_in_lod, lod_count,
                           num_detail_levels, attrs_out, num_attrs, lcp_coeffs != nullptr,
                           const_cast<int32_t*>(values_in), const_cast<int8_t*>(lcp_coeffs));
}

// the whole lifting attribute coder minus entropy coding, on the device
static int
attr_lift_common(bool forward, const pccb200_lod_params* lod, const pccb200_qpset* qpset,
                 int32_t lcpEnabled, const int32_t* qpo, const int32_t* xyz, int32_t* attrs,
                 int32_t A, int32_t n, int32_t bitdepth, int32_t* values, int8_t* lcp,
                 bool device = false)
{
  if (!lod || !qpset || !xyz || !attrs || !values || n <= 0 || (A != 1 && A != 3)
      || bitdepth < 1 || bitdepth > 16)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  int8_t lcpLocal[PCCB200_MAX_LODS + 1] = {};
  if (!forward && lcpEnabled && A == 3) {
    if (!lcp)
      return fail(PCCB200_ERR_INVALID_ARG, "lcp coefficients missing");
    for (int l = 0; l < lod->num_detail_levels && l < PCCB200_MAX_LODS; l++)
      lcpLocal[l] = lcp[l];
  }
  int rc = with_device([&](DeviceExec& ex) -> int {
    // device: xyz / qpo / attrs / values are device pointers; attrs is coded in
    // place (the gather reads it before the final write-back overwrites it)
    const int32_t* dXyz = device ? xyz : to_device(ex, xyz, size_t(n) * 3);
    const int32_t* dIn = !forward ? nullptr : device ? attrs : to_device(ex, attrs, size_t(n) * A);
    const int32_t* dQpoIn = !qpo ? nullptr : device ? qpo : to_device(ex, qpo, size_t(n) * 2);
    int32_t* dV = device ? values
                         : forward ? ex.alloc<int32_t>(size_t(n) * A)
                                   : to_device(ex, values, size_t(n) * A);
    int32_t* dOut = device ? attrs : ex.alloc<int32_t>(size_t(n) * A);
    int rc2 = attr_lift_run(ex, forward, *lod, *qpset, lcpEnabled != 0, dQpoIn, dXyz, dIn, dOut, A,
                            n, bitdepth, dV, lcpLocal);
    if (rc2 != PCCB200_OK)
      return fail(rc2, rc2 == PCCB200_ERR_UNSUPPORTED
                         ? "a predictor references its own level of detail"
                         : "invalid lifting parameters");
    if (!device) {
      to_host(ex, attrs, dOut, size_t(n) * A);
      if (forward)
        to_host(ex, values, dV, size_t(n) * A);
    }
    return PCCB200_OK;
  });
  if (rc == PCCB200_OK && forward && lcp)
    for (int l = 0; l < lod->num_detail_levels && l < PCCB200_MAX_LODS; l++)
      lcp[l] = lcpLocal[l];
  return rc;
}

}  // extern "C"

struct pccb200_lod_handle_s {
  int device;
  pccb200_lod_params params;
  pccb200::LodState st;
  void* block;  // one device allocation behind preds / idx / qw
  // the lifting quantisation weights are computed by the first lifting call on
  // the handle (a handle made for the predicting transform never needs them)
  std::mutex qwMu;
  bool qwReady = false;
};

namespace {

int
attr_lift_lod_common(bool forward, pccb200_lod_handle h, const pccb200_qpset* qpset,
                     int32_t lcpEnabled, const int32_t* qpo, int32_t* attrs, int32_t A,
                     int32_t bitdepth, int32_t* values, int8_t* lcp)
{
  if (!h || !qpset || !attrs || !values || (A != 1 && A != 3) || bitdepth < 1 || bitdepth > 16)
    return pccb200::fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  if (h->device != pccb200::ctx().device)
    return pccb200::fail(PCCB200_ERR_INVALID_ARG, "handle belongs to another device");
  const int n = h->st.n;
  const int levels = h->params.num_detail_levels;
  int8_t lcpLocal[PCCB200_MAX_LODS + 1] = {};
  if (!forward && lcpEnabled && A == 3) {
    if (!lcp)
      return pccb200::fail(PCCB200_ERR_INVALID_ARG, "lcp coefficients missing");
    for (int l = 0; l < levels && l < PCCB200_MAX_LODS; l++)
      lcpLocal[l] = lcp[l];
  }
  int rc = pccb200::with_device([&](pccb200::DeviceExec& ex) -> int {
    int32_t* dIn = forward ? pccb200::to_device(ex, attrs, size_t(n) * A) : nullptr;
    int32_t* dQpoIn = qpo ? pccb200::to_device(ex, qpo, size_t(n) * 2) : nullptr;
    int32_t* dV = forward ? ex.alloc<int32_t>(size_t(n) * A)
                          : pccb200::to_device(ex, values, size_t(n) * A);
    int32_t* dOut = ex.alloc<int32_t>(size_t(n) * A);
    {
      std::lock_guard<std::mutex> g(h->qwMu);
      if (!h->qwReady) {
        int rcq = pccb200::run_quant_weights(ex, h->st.preds, n, h->st.npl, h->st.lodCount, h->st.qw);
        if (rcq != PCCB200_OK)
          return pccb200::fail(rcq, "invalid levels of detail");
        PCC_CUDA_CHECK(cudaStreamSynchronize(ex.stream));  // other lanes read them from now on
        h->qwReady = true;
      }
    }
    int rc2 = pccb200::attr_lift_on_lods(ex, forward, h->st, *qpset, lcpEnabled != 0, dQpoIn, dIn,
                                         dOut, A, bitdepth, dV, lcpLocal);
    if (rc2 != PCCB200_OK)
      return pccb200::fail(rc2, rc2 == PCCB200_ERR_UNSUPPORTED
                                  ? "a predictor references its own level of detail"
                                  : "invalid lifting parameters");
    pccb200::to_host(ex, attrs, dOut, size_t(n) * A);
    if (forward)
      pccb200::to_host(ex, values, dV, size_t(n) * A);
    return PCCB200_OK;
  });
  if (rc == PCCB200_OK && forward && lcp)
    for (int l = 0; l < levels && l < PCCB200_MAX_LODS; l++)
      lcp[l] = lcpLocal[l];
  return rc;
}

}  // namespace

extern "C" {

int
pccb200_lod_create(const pccb200_lod_params* params, const int32_t* xyz, int32_t n,
                   pccb200_lod_handle* handle_out)
{
  if (!params || !xyz || !handle_out || n <= 0)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  *handle_out = nullptr;
  pccb200_lod_handle h = new (std::nothrow) pccb200_lod_handle_s();
  if (!h)
    return fail(PCCB200_ERR_NOMEM, "host allocation failed");
  h->params = *params;
  h->block = nullptr;
  int rc = with_device([&](DeviceExec& ex) -> int {
    h->device = ctx().device;
    const size_t szP = (size_t(n) * sizeof(pccb200_predictor) + 255) & ~size_t(255);
    const size_t szQ = (size_t(n) * sizeof(uint64_t) + 255) & ~size_t(255);
    const size_t szI = (size_t(n) * sizeof(uint32_t) + 255) & ~size_t(255);
    PCC_CUDA_CHECK(cudaMalloc(&h->block, szP + szQ + szI));
    char* b = static_cast<char*>(h->block);
    h->st.preds = reinterpret_cast<pccb200_predictor*>(b);
    h->st.qw = reinterpret_cast<uint64_t*>(b + szP);
    h->st.idx = reinterpret_cast<uint32_t*>(b + szP + szQ);
    int32_t* dXyz = to_device(ex, xyz, size_t(n) * 3);
    int rc2 = lod_state_build(ex, *params, dXyz, n, h->st, false);
    if (rc2 != PCCB200_OK)
      return fail(rc2, "invalid LoD parameters");
    return PCCB200_OK;
  });
  if (rc != PCCB200_OK) {
    if (h->block)
      cudaFree(h->block);
    delete h;
    return rc;
  }
  *handle_out = h;
  return PCCB200_OK;
}

void
pccb200_lod_destroy(pccb200_lod_handle handle)
{
  if (!handle)
    return;
  if (handle->block) {
    cudaSetDevice(handle->device);
    cudaFree(handle->block);
  }
  delete handle;
}

int
pccb200_lod_reusable(pccb200_lod_handle h, const pccb200_lod_params* p)
{
  if (!h || !p)
    return 0;
  const pccb200_lod_params& a = h->params;
  // the order of AttributeLods::isReusable (tmc3/AttributeCommon.cpp:76-140)
  if (a.num_pred_nearest_neighbours != p->num_pred_nearest_neighbours
      || a.inter_lod_search_range != p->inter_lod_search_range
      || a.intra_lod_search_range != p->intra_lod_search_range
      || a.num_detail_levels != p->num_detail_levels)
    return 0;
  for (int k = 0; k < 3; k++)
    if (a.lod_neigh_bias[k] != p->lod_neigh_bias[k])
      return 0;
  if (a.lod_decimation_type != p->lod_decimation_type || a.dist2 != p->dist2)
    return 0;
  for (int l = 0; l < PCCB200_MAX_LODS; l++)
    if (a.lod_sampling_period[l] != p->lod_sampling_period[l])
      return 0;
  if (a.intra_lod_prediction_skip_layers != p->intra_lod_prediction_skip_layers
      || a.pred_weight_blending != p->pred_weight_blending
      || a.prediction_with_distribution != p->prediction_with_distribution)
    return 0;
  return 1;
}

int
pccb200_lod_info(pccb200_lod_handle h, int32_t* n_out, int32_t* lod_count_out,
                 uint32_t* num_points_in_lod_out)
{
  if (!h)
    return fail(PCCB200_ERR_INVALID_ARG, "null handle");
  if (n_out)
    *n_out = h->st.n;
  if (lod_count_out)
    *lod_count_out = h->st.lodCount;
  if (num_points_in_lod_out)
    for (int l = 0; l < PCCB200_MAX_LODS; l++)
      num_points_in_lod_out[l] = l < h->st.lodCount ? h->st.npl[l] : 0;
  return PCCB200_OK;
}

int
pccb200_attr_lift_encode_lod(pccb200_lod_handle handle, const pccb200_qpset* qpset,
                             int32_t lcp_enabled, const int32_t* point_qp_offsets,
                             int32_t* attrs_inout, int32_t num_attrs, int32_t bitdepth,
                             int32_t* values_out, int8_t* lcp_coeffs_out)
{
  return attr_lift_lod_common(true, handle, qpset, lcp_enabled, point_qp_offsets, attrs_inout,
                              num_attrs, bitdepth, values_out, lcp_coeffs_out);
}

int
pccb200_attr_lift_decode_lod(pccb200_lod_handle handle, const pccb200_qpset* qpset,
                             int32_t lcp_enabled, const int32_t* point_qp_offsets,
                             int32_t* attrs_out, int32_t num_attrs, int32_t bitdepth,
                             const int32_t* values_in, const int8_t* lcp_coeffs)
{
  return attr_lift_lod_common(false, handle, qpset, lcp_enabled, point_qp_offsets, attrs_out,
                              num_attrs, bitdepth, const_cast<int32_t*>(values_in),
                              const_cast<int8_t*>(lcp_coeffs));
}

int
pccb200_attr_lift_encode(const pccb200_lod_params* lod, const pccb200_qpset* qpset,
                         int32_t lcp_enabled, const int32_t* point_qp_offsets, const int32_t* xyz,
                         int32_t* attrs_inout, int32_t num_attrs, int32_t n, int32_t bitdepth,
                         int32_t* values_out, int8_t* lcp_coeffs_out)
{
  return attr_lift_common(true, lod, qpset, lcp_enabled, point_qp_offsets, xyz, attrs_inout,
                          num_attrs, n, bitdepth, values_out, lcp_coeffs_out);
}

int
pccb200_attr_lift_decode(const pccb200_lod_params* lod, const pccb200_qpset* qpset,
                         int32_t lcp_enabled, const int32_t* point_qp_offsets, const int32_t* xyz,
                         int32_t* attrs_out, int32_t num_attrs, int32_t n, int32_t bitdepth,
                         const int32_t* values_in, const int8_t* lcp_coeffs)
{
  return attr_lift_common(false, lod, qpset, lcp_enabled, point_qp_offsets, xyz, attrs_out,
                          num_attrs, n, bitdepth, const_cast<int32_t*>(values_in),
                          const_cast<int8_t*>(lcp_coeffs));
}

// Slices of a frame (tmc3/encoder.cpp:545-568): each with its own levels of
// detail, each on its own lane.  lcp: num_slices rows of PCCB200_MAX_LODS.
static int
attr_lift_slices(bool forward, const pccb200_lod_params* lod, const pccb200_qpset* qpset,
                 int32_t lcpEnabled, const int32_t* qpo, const int32_t* xyz, int32_t* attrs,
                 int32_t A, int32_t bitdepth, const int64_t* sliceOffsets, int32_t numSlices,
                 int32_t* values, int8_t* lcp, bool device = false)
{
  if (!sliceOffsets || numSlices <= 0)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  for (int s = 0; s < numSlices; s++) {
    const int64_t len = sliceOffsets[s + 1] - sliceOffsets[s];
    if (len <= 0 || len > INT32_MAX)
      return fail(PCCB200_ERR_INVALID_ARG, "empty or oversized slice");
  }
  return parallel_for(numSlices, kMaxSliceThreads, [&](int s) -> int {
    const int64_t o = sliceOffsets[s];
    return attr_lift_common(forward, lod, qpset, lcpEnabled, qpo ? qpo + 2 * o : nullptr,
                            xyz + 3 * o, attrs + o * A, A, int32_t(sliceOffsets[s + 1] - o),
                            bitdepth, values + o * A,
                            lcp ? lcp + size_t(s) * PCCB200_MAX_LODS : nullptr, device);
  });
}

int
pccb200_attr_lift_encode_slices(const pccb200_lod_params* lod, const pccb200_qpset* qpset,
                                int32_t lcp_enabled, const int32_t* point_qp_offsets,
                                const int32_t* xyz, int32_t* attrs_inout, int32_t num_attrs,
                                int32_t bitdepth, const int64_t* slice_offsets,
                                int32_t num_slices, int32_t* values_out, int8_t* lcp_coeffs_out)
{
  return attr_lift_slices(true, lod, qpset, lcp_enabled, point_qp_offsets, xyz, attrs_inout,
                          num_attrs, bitdepth, slice_offsets, num_slices, values_out,
                          lcp_coeffs_out);
}

int
pccb200_attr_lift_decode_slices(const pccb200_lod_params* lod, const pccb200_qpset* qpset,
                                int32_t lcp_enabled, const int32_t* point_qp_offsets,
                                const int32_t* xyz, int32_t* attrs_out, int32_t num_attrs,
                                int32_t bitdepth, const int64_t* slice_offsets,
                                int32_t num_slices, const int32_t* values_in,
                                const int8_t* lcp_coeffs)
{
  return attr_lift_slices(false, lod, qpset, lcp_enabled, point_qp_offsets, xyz, attrs_out,
                          num_attrs, bitdepth, slice_offsets, num_slices,
                          const_cast<int32_t*>(values_in), const_cast<int8_t*>(lcp_coeffs));
}

int
pccb200_attr_lift_encode_slices_dev(const pccb200_lod_params* lod, const pccb200_qpset* qpset,
                                    int32_t lcp_enabled, const int32_t* d_point_qp_offsets,
                                    const int32_t* d_xyz, int32_t* d_attrs_inout,
                                    int32_t num_attrs, int32_t bitdepth,
                                    const int64_t* slice_offsets, int32_t num_slices,
                                    int32_t* d_values_out, int8_t* lcp_coeffs_out)
{
  return attr_lift_slices(true, lod, qpset, lcp_enabled, d_point_qp_offsets, d_xyz, d_attrs_inout,
                          num_attrs, bitdepth, slice_offsets, num_slices, d_values_out,
                          lcp_coeffs_out, true);
}

int
pccb200_attr_lift_decode_slices_dev(const pccb200_lod_params* lod, const pccb200_qpset* qpset,
                                    int32_t lcp_enabled, const int32_t* d_point_qp_offsets,
                                    const int32_t* d_xyz, int32_t* d_attrs_out, int32_t num_attrs,
                                    int32_t bitdepth, const int64_t* slice_offsets,
                                    int32_t num_slices, const int32_t* d_values_in,
                                    const int8_t* lcp_coeffs)
{
  return attr_lift_slices(false, lod, qpset, lcp_enabled, d_point_qp_offsets, d_xyz, d_attrs_out,
                          num_attrs, bitdepth, slice_offsets, num_slices,
                          const_cast<int32_t*>(d_values_in), const_cast<int8_t*>(lcp_coeffs),
                          true);
}

//----------------------------------------------------------------------------
// spherical coordinates (spherical.cuh)

static int
spherical_common(const int32_t* origin, const int32_t* theta, int32_t numTheta,
                 const int32_t* weight, const int32_t* minPos, const int32_t* xyz, int64_t n,
                 int32_t* out, int32_t* bbox)
{
  if (!origin || !theta || !xyz || !out || !bbox || numTheta < 1 || n < 0
      || n > int64_t(INT32_MAX) / 3)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    const int32_t* dXyz = to_device(ex, xyz, size_t(n) * 3);
    const int32_t* dTheta = to_device(ex, theta, size_t(numTheta));
    int32_t* dOut = ex.alloc<int32_t>(size_t(n) * 3);
    run_xyz_to_rpl(ex, origin, dTheta, numTheta, dXyz, n, dOut, bbox, minPos, weight);
    to_host(ex, out, dOut, size_t(n) * 3);
    return PCCB200_OK;
  });
}

int
pccb200_xyz_to_rpl(const int32_t laser_origin[3], const int32_t* laser_theta, int32_t num_theta,
                   const int32_t* xyz, int64_t n, int32_t* rpl_out, int32_t bbox_out[6])
{
  return spherical_common(laser_origin, laser_theta, num_theta, nullptr, nullptr, xyz, n,
                          rpl_out, bbox_out);
}

int
pccb200_attr_spherical_positions(const int32_t laser_origin[3], const int32_t* laser_theta,
                                 int32_t num_theta, const int32_t axis_weight[3],
                                 const int32_t* min_pos, const int32_t* xyz, int64_t n,
                                 int32_t* pos_out, int32_t bbox_out[6])
{
  if (!axis_weight)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return spherical_common(laser_origin, laser_theta, num_theta, axis_weight, min_pos, xyz, n,
                          pos_out, bbox_out);
}

int
pccb200_offset_and_scale(const int32_t min_pos[3], const int32_t axis_weight[3],
                         int32_t* pos_inout, int64_t n)
{
  if (!min_pos || !axis_weight || !pos_inout || n < 0 || n > int64_t(INT32_MAX) / 3)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    int32_t* dPos = to_device(ex, pos_inout, size_t(n) * 3);
    OffsetScaleFn os;
    for (int k = 0; k < 3; k++) {
      os.minPos[k] = min_pos[k];
      os.weight[k] = axis_weight[k];
    }
    os.pos = dPos;
    ex.foreach(n, os);
    to_host(ex, pos_inout, dPos, size_t(n) * 3);
    return PCCB200_OK;
  });
}

//----------------------------------------------------------------------------
// symbol preparation for the entropy coder (symbols.cuh)

static void
symbols_to_host(DeviceExec& ex, const int32_t* dRuns, const int32_t* dValues, const uint8_t* dCtx,
                int count, int A, int32_t* runs, int32_t* values, uint8_t* ctx)
{
  to_host(ex, runs, dRuns, size_t(count));
  to_host(ex, values, dValues, size_t(count) * A);
  if (ctx && A == 3)
    to_host(ex, ctx, dCtx, size_t(count));
}

int
pccb200_coeff_symbols(const int32_t* coeffs, int32_t num_attrs, int32_t n,
                      int32_t* zero_runs_out, int32_t* values_out, uint8_t* ctx_out,
                      int32_t* count_out, int32_t* tail_run_out)
{
  if (!coeffs || !zero_runs_out || !values_out || !count_out || !tail_run_out || n <= 0
      || (num_attrs != 1 && num_attrs != 3))
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    const int A = num_attrs;
    const int32_t* dCoef = to_device(ex, coeffs, size_t(n) * A);
    int32_t* dRuns = ex.alloc<int32_t>(size_t(n));
    int32_t* dValues = ex.alloc<int32_t>(size_t(n) * A);
    uint8_t* dCtx = ex.alloc<uint8_t>(size_t(n));
    int count = 0, tail = 0;
    run_coeff_symbols(ex, dCoef, n, A, n, dRuns, dValues, dCtx, &count, &tail);
    symbols_to_host(ex, dRuns, dValues, dCtx, count, A, zero_runs_out, values_out, ctx_out);
    *count_out = count;
    *tail_run_out = tail;
    return PCCB200_OK;
  });
}

int
pccb200_attr_raht_encode_symbols(const pccb200_raht_params* params, const pccb200_qpset* qpset,
                                 const int32_t* point_qp_offsets, const int32_t* xyz,
                                 int32_t* attrs_inout, int32_t num_attrs, int32_t n,
                                 int32_t bitdepth, int32_t* zero_runs_out, int32_t* values_out,
                                 uint8_t* ctx_out, int32_t* count_out, int32_t* tail_run_out)
{
  const int64_t offs[2] = {0, n};
  int rc = check_slices(params, qpset, xyz, attrs_inout, values_out, num_attrs, bitdepth, offs, 1);
  if (rc != PCCB200_OK)
    return rc;
  if (!zero_runs_out || !count_out || !tail_run_out || num_attrs == 2)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    const int A = num_attrs;
    int32_t* dXyz = to_device(ex, xyz, size_t(n) * 3);
    int32_t* dAttrsIn = to_device(ex, attrs_inout, size_t(n) * A);
    int32_t* dQpoIn = point_qp_offsets ? to_device(ex, point_qp_offsets, size_t(n) * 2) : nullptr;
    int32_t* dCoef = ex.alloc<int32_t>(size_t(n) * A);
    int32_t* dOut = ex.alloc<int32_t>(size_t(n) * A);
    int rc2 = attr_raht_slice(ex, true, params, qpset, dQpoIn, dXyz, dAttrsIn, dOut, A, bitdepth,
                              n, dCoef, n);
    if (rc2 != PCCB200_OK)
      return rc2;
    to_host(ex, attrs_inout, dOut, size_t(n) * A);
    int32_t* dRuns = ex.alloc<int32_t>(size_t(n));
    int32_t* dValues = ex.alloc<int32_t>(size_t(n) * A);
    uint8_t* dCtx = ex.alloc<uint8_t>(size_t(n));
    int count = 0, tail = 0;
    run_coeff_symbols(ex, dCoef, n, A, n, dRuns, dValues, dCtx, &count, &tail);
    symbols_to_host(ex, dRuns, dValues, dCtx, count, A, zero_runs_out, values_out, ctx_out);
    *count_out = count;
    *tail_run_out = tail;
    return PCCB200_OK;
  });
}

//----------------------------------------------------------------------------
// estimateDist2 (dist2.cuh)

int
pccb200_estimate_dist2(const int32_t* xyz, int32_t n, int32_t sampling_period,
                       int32_t search_range, float percentile_estimate, int32_t* shift_bits_out)
{
  if (!xyz || !shift_bits_out || n < 0 || sampling_period < 1 || search_range < 0)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  if (n < 2) {
    *shift_bits_out = 0;
    return PCCB200_OK;
  }
  return with_device([&](DeviceExec& ex) -> int {
    const int32_t* dXyz = to_device(ex, xyz, size_t(n) * 3);
    const int s = run_estimate_dist2(ex, dXyz, n, sampling_period, search_range,
                                     percentile_estimate);
    if (s < 0)
      return fail(PCCB200_ERR_INVALID_ARG, "percentile outside [0, 1)");
    *shift_bits_out = s;
    return PCCB200_OK;
  });
}

//----------------------------------------------------------------------------
// the other two quantisation-weight derivations (lifting.cuh)

int
pccb200_quant_weights_fixed(const pccb200_predictor* preds, int32_t n,
                            const uint32_t* num_points_in_lod, int32_t lod_count,
                            const int32_t neigh_weight[3], uint64_t* qw_out)
{
  if (!preds || !num_points_in_lod || !neigh_weight || !qw_out || n <= 0 || lod_count <= 0)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    pccb200_predictor* dP = to_device(ex, preds, size_t(n));
    uint64_t* dQw = ex.alloc<uint64_t>(n);
    int rc = run_quant_weights(ex, dP, n, num_points_in_lod, lod_count, dQw, neigh_weight);
    if (rc != PCCB200_OK)
      return fail(rc, "numPointsInLod does not partition [0, n)");
    to_host(ex, qw_out, dQw, size_t(n));
    return PCCB200_OK;
  });
}

int
pccb200_quant_weights_scalable(const uint32_t* num_points_in_lod, int32_t lod_count,
                               int64_t num_points, int32_t min_geom_node_size_log2, int32_t n,
                               uint64_t* qw_out)
{
  if (!num_points_in_lod || !qw_out || n <= 0 || lod_count <= 0 || num_points < 0)
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    uint64_t* dQw = ex.alloc<uint64_t>(n);
    int rc = run_quant_weights_scalable(ex, num_points_in_lod, lod_count, uint64_t(num_points),
                                        min_geom_node_size_log2, n, dQw);
    if (rc != PCCB200_OK)
      return fail(rc, "numPointsInLod does not partition [0, n)");
    to_host(ex, qw_out, dQw, size_t(n));
    return PCCB200_OK;
  });
}

//----------------------------------------------------------------------------
// recolouring (recolour.cuh)

void
pccb200_recolour_params_default(pccb200_recolour_params* p)
{
  if (!p)
    return;
  // tmc3/TMC3.cpp:1500-1551
  p->dist_offset_fwd = 4.;
  p->dist_offset_bwd = 4.;
  p->max_geometry_dist2_fwd = 1000.;
  p->max_geometry_dist2_bwd = 1000.;
  p->max_attribute_dist2_fwd = 1000.;
  p->max_attribute_dist2_bwd = 1000.;
  p->search_range = 1;
  p->num_neighbours_fwd = 8;
  p->num_neighbours_bwd = 1;
  p->use_dist_weighted_avg_fwd = 1;
  p->use_dist_weighted_avg_bwd = 1;
  p->skip_avg_if_identical_source_point_present_fwd = 1;
  p->skip_avg_if_identical_source_point_present_bwd = 0;
  p->reserved = 0;
}

int
pccb200_recolour(const pccb200_recolour_params* params, const int32_t* source_xyz,
                 const int32_t* source_attrs, int32_t num_attrs, int32_t n_source,
                 double source_to_target_scale, const int32_t tgt_to_src_offset[3],
                 const int32_t* target_xyz, int32_t n_target, int32_t bitdepth,
                 int32_t* target_attrs_out)
{
  if (!params || !source_xyz || !source_attrs || !tgt_to_src_offset || !target_xyz
      || !target_attrs_out || n_source <= 0 || n_target <= 0 || (num_attrs != 1 && num_attrs != 3))
    return fail(PCCB200_ERR_INVALID_ARG, "null pointer or bad size");
  return with_device([&](DeviceExec& ex) -> int {
    int32_t* dSrc = to_device(ex, source_xyz, size_t(n_source) * 3);
    int32_t* dAttr = to_device(ex, source_attrs, size_t(n_source) * num_attrs);
    int32_t* dTgt = to_device(ex, target_xyz, size_t(n_target) * 3);
    int32_t* dOut = ex.alloc<int32_t>(size_t(n_target) * num_attrs);
    int rc = recolour_run(ex, *params, dSrc, dAttr, num_attrs, n_source, source_to_target_scale,
                          tgt_to_src_offset, dTgt, n_target, bitdepth, dOut);
    if (rc != PCCB200_OK)
      return fail(rc, "invalid recolouring parameters (neighbour counts, scale, or a coordinate "
                      "outside [0, 2^21))");
    to_host(ex, target_attrs_out, dOut, size_t(n_target) * num_attrs);
    return PCCB200_OK;
  });
}

}  // extern "C"
