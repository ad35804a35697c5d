// dag_depth.cpp — developer analysis tool (host only, not part of the product
// library): runs the kernel bodies in order on the host (tests/emu/exec_host.h)
// and reports, per transform stage, the number of blocks, of transforming
// (multi-child) blocks, and the depth of the dependency DAG that the dataflow
// kernel has to respect (sub-node prediction reads earlier neighbours'
// children, RAHT.cpp:370-415), plus the critical path of the whole descent
// when stages are allowed to overlap.
//
//   g++ -std=c++17 -O2 -x c++ -Impeg-pcc-tmc13_b200/csrc -Iinclude -Itests/emu tools/dag_depth.cpp -o /tmp/dag_depth
//   /tmp/dag_depth frame.bin     (int32 N, A, then N*3 xyz, then N*A attrs; qp, searchRange as argv)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "exec_host.h"
#include "raht_pipeline.cuh"

using namespace pccb200;

struct DagExec : HostExec {
  std::vector<int> nodeTE;  // same, encoder (with zero-run look-back edges)
  std::vector<int> lcAll;   // classification level of every block so far in coding order
  std::vector<char> hardAll;
  std::vector<int> nodeT;   // per node of the parent stage: time its value exists (whole-call DAG)
  long totalBlocks = 0, totalMulti = 0;
  int stageNo = 0;
  long sumDepth = 0;
  template<class Fn>
  void block_stage(const Fn& fn, int64_t nBlocks, int* tzNext)
  {
    const Stage& S = fn.S;
    const Stage& P = fn.P;
    const RahtConfig& cfg = fn.cfg;
    const int A = cfg.A;
    std::vector<int> nodeTS(S.n, 0);
    if (P.n == 0) {
      HostExec::block_stage(fn, nBlocks, tzNext);
      for (int c = 0; c < S.n; c++) nodeTS[c] = 1;
      nodeT = nodeTS; nodeTE = nodeTS; lcAll.push_back(1); hardAll.push_back(1);
      printf("stage %2d level %2d blocks %8d multi %8d depth %6d  crit %6d\n", stageNo++, S.level, 1, 1, 1, 1);
      return;
    }
    // values of this stage are needed for nothing in the analysis; parent-stage values for the range test
    std::vector<int> depth(nBlocks, 0);
    std::vector<std::vector<int>> depsOf(nBlocks);
    long multi = 0; int maxDepth = 0; int maxT = 0;
    long nDeps = 0, nPred = 0;
    // run stage first so that S.nn etc are final? P.nn is already final. Range test needs P.rec: final.
    for (int p = 0; p < nBlocks; p++) {
      const int c0 = P.first[p], c1 = P.first[p + 1];
      if (c1 - c0 < 2) {
        nodeTS[c0] = nodeT[p];
        continue;
      }
      multi++;
      uint32_t occ = P.occ[p];
      int d = 0;
      int t = nodeT[p];
      bool enablePred = fn.predInLvl != 0;
      int pidx[19];
      int count = 0;
      if (fn.predInLvl) {
        if (P.nn[p] < cfg.thr0) enablePred = false;
        else {
          const int plevel = S.level + 3;
          const int64_t cur = P.key[p] >> plevel;
          const int64_t base = int64_t(morton3d_add(uint64_t(cur), ~uint64_t(0)));
          pidx[0] = p; count = 1;
          for (int i = 1; i < 19; i++) {
            pidx[i] = -1;
            if (!(occ & neigh_mask(i))) continue;
            pidx[i] = find_parent_neighbour(P, p, plevel, cur, base, i, cfg.searchRange);
            count += pidx[i] >= 0;
          }
          if (count < cfg.thr1) enablePred = false;
        }
      }
      if (enablePred) {
        nPred++;
        int64_t lim0 = P.rec[size_t(p) * A];
        for (int i = 0; i < 19; i++) {
          int q = pidx[i];
          if (q < 0) continue;
          depsOf[p].push_back(q);
          t = std::max(t, nodeT[q]);   // needs the neighbour's parent-stage value (even for the range test)
          int64_t v0 = P.rec[size_t(q) * A];
          if (i && (10 * v0 <= 2 * lim0 || 10 * v0 >= 25 * lim0)) continue;
          if (i >= (cfg.subnode ? 7 : 19) && q < p) {
            const int ii = i - 7;
            int sh = occu_shift(ii);
            uint32_t nocc = P.occ[q];
            uint32_t cmask = (ii < 9 ? (nocc >> sh) : (nocc << sh)) & neigh_mask(i) & occ & 0xff;
            if (cmask) {
              nDeps++; depsOf[p].push_back(-q - 1);
              int cf = P.first[q], cl = P.first[q + 1];
              if (cl - cf >= 2) d = std::max(d, depth[q]);
              for (int c = cf; c < cl; c++) t = std::max(t, nodeTS[c]);
            }
          }
        }
      }
      depth[p] = d + 1;
      maxDepth = std::max(maxDepth, depth[p]);
      for (int c = c0; c < c1; c++) nodeTS[c] = t + 1;
      maxT = std::max(maxT, t + 1);
    }
    HostExec::block_stage(fn, nBlocks, tzNext);
    // encoder levels: second pass with the final coefficients of this stage
    int maxTE = 0; long nLook = 0; long sumWin = 0;
    {
      std::vector<int> nodeTES(S.n, 0);
      for (int p = 0; p < nBlocks; p++) {
        const int c0 = P.first[p], c1 = P.first[p + 1];
        if (c1 - c0 < 2) { nodeTES[c0] = nodeTE[p]; continue; }
        int t = nodeTE[p];
        for (int q : depsOf[p]) {
          if (q >= 0) t = std::max(t, nodeTE[q]);
          else { int qq = -q - 1; for (int c = P.first[qq]; c < P.first[qq + 1]; c++) t = std::max(t, nodeTES[c]); }
        }
        int lc = t + 1;
        // classification from the final coefficients
        int64_t pos = fn.coefBase + c0 - p;
        int ncoef = c1 - c0 - 1;
        bool soft = false, hard = false, softFirst = false;
        for (int i = 0; i < ncoef; i++) {
          long sm = 0;
          for (int k = 0; k < A; k++) sm += labs(fn.coef[k * fn.coefStride + pos + i]);
          if (sm >= 3) hard = true;
          else if (sm > 0) { soft = true; if (!hard) softFirst = true; }
        }
        int lr = lc;
        if (softFirst) {
          nLook++;
          for (long u = (long)lcAll.size() - 1; u >= 0; u--) {
            lr = std::max(lr, lcAll[u]); sumWin++;
            if (hardAll[u]) break;
          }
        }
        lcAll.push_back(lc); hardAll.push_back(hard);
        for (int c = c0; c < c1; c++) nodeTES[c] = lr + (lr > lc ? 1 : 0);
        maxTE = std::max(maxTE, lr);
      }
      nodeTE = nodeTES;
    }
    printf("   encoder: crit %6d lookbacks %8ld avg window %.1f blocks\n", maxTE, nLook, nLook ? double(sumWin) / nLook : 0.0);
    nodeT = nodeTS;
    totalBlocks += nBlocks; totalMulti += multi; sumDepth += maxDepth;
    printf("stage %2d level %2d blocks %8ld multi %8ld pred %8ld deps %8ld depth %6d  crit %6d\n", stageNo++, S.level,
           (long)nBlocks, multi, nPred, nDeps, maxDepth, maxT);
  }
};

int main(int argc, char** argv)
{
  FILE* f = fopen(argv[1], "rb");
  int32_t hdr[2];
  if (fread(hdr, 4, 2, f) != 2) return 1;
  int N = hdr[0], A = hdr[1];
  std::vector<int32_t> xyz(size_t(N) * 3), attrs(size_t(N) * A);
  if (fread(xyz.data(), 4, xyz.size(), f) != xyz.size()) return 1;
  if (fread(attrs.data(), 4, attrs.size(), f) != attrs.size()) return 1;
  fclose(f);
  int qp = argc > 2 ? atoi(argv[2]) : 34;
  int sr = argc > 3 ? atoi(argv[3]) : 2500;
  int forward = argc > 4 ? atoi(argv[4]) : 1;
  pccb200_raht_params pp = {};
  pp.raht_extension = 1; pp.prediction_enabled = 1; pp.subnode_prediction_enabled = 1;
  pp.prediction_threshold0 = 2; pp.prediction_threshold1 = 6; pp.prediction_search_range = sr;
  const int wp[19] = {3,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1};  // placeholder, overwritten below via file if needed
  (void)wp;
  // defaults of the reference (RAHT.h / cfg): parent weights {9,3,3,3,1,1,1,...}? read from python side instead
  FILE* g = fopen(argv[5], "rb");
  if (!g || fread(&pp, sizeof(pp), 1, g) != 1) return 2;
  fclose(g);
  pp.prediction_search_range = sr;
  pccb200_qpset qs = {};
  qs.num_layers = 1; qs.layers[0][0] = qp; qs.layers[0][1] = -2; qs.max_qp = 51;
  DagExec ex;
  std::vector<int64_t> keys(N); std::vector<int32_t> order(N);
  ex.morton_sort(xyz.data(), N, keys.data(), order.data());
  std::vector<int32_t> sa(size_t(N) * A), coef(size_t(N) * A);
  for (int i = 0; i < N; i++) for (int k = 0; k < A; k++) sa[size_t(i) * A + k] = attrs[size_t(order[i]) * A + k];
  int rc = raht_run(ex, pp, qs, forward != 0, keys.data(), sa.data(), nullptr, coef.data(), int64_t(N), A, N);
  long nz = 0, soft = 0;
  for (int i = 0; i < N; i++) { long s = 0; for (int k = 0; k < A; k++) s += labs(coef[size_t(k) * N + i]); nz += s != 0; soft += (s == 1 || s == 2); }
  printf("rc %d blocks %ld multi %ld sum-of-stage-depths %ld  nonzero %ld soft %ld of %d\n", rc, ex.totalBlocks, ex.totalMulti, ex.sumDepth, nz, soft, N);
  return 0;
}
