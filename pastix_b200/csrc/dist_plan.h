// dist_plan.h — host-side partition of the column blocks over the GPUs of one box.
//
// The reference maps the elimination tree onto processors with blend's proportional mapping
// (blend/src/splitpart.c:752-1012 propMappTree / propMappSubtree): every subtree gets a set of
// candidate processors in proportion to its cost, a subtree whose set shrinks to one processor lives
// there entirely, and the column blocks of the shared top separators are spread over their candidate
// set (1-D distribution, IPARM_DISTRIBUTION_LEVEL = 0).  Contributions to a column block owned by
// another processor are summed locally and shipped once (fan-in, sopalin_compute.c:600-733).
// Here the same rule runs over the SolverMatrix the single-process analysis produced, so the symbolic
// structure (and therefore the panel layout) is identical on every GPU and to the 1-GPU run:
//   parent(c)  = facing cblk of c's first off-diagonal blok
//   cost(c)    = PaStiX's flop model of the cblk (blend_symbol_cost.c:52-88, 382-430)
//   candidates = contiguous rank interval, split among the children in proportion to subtree cost
//   owner(c)   = the only candidate, or (shared cblk) round-robin over the interval along the chain.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace pb200 {

struct DistPlan {
  int nranks = 1;
  std::vector<int> owner;          // per cblk
  std::vector<uint32_t> contrib;   // per cblk: bit p set <=> rank p != owner holds a cblk with a blok facing it
  std::vector<double> load;        // per rank: flops mapped to it
};

// fblok[C+1], fcblk[B], width[C], stride[C], nrow[B], coefind[B]
inline DistPlan dist_plan(int64_t C, const int *fblok, const int *fcblk, const int *width, const int *stride,
                          const int *nrow, const int *coefind, int nranks, bool lu) {
  DistPlan P;
  P.nranks = nranks;
  P.owner.assign(C, 0);
  P.contrib.assign(C, 0u);
  P.load.assign(nranks, 0.0);
  std::vector<double> cost(C), sub(C);
  std::vector<int> parent(C, -1);
  for (int64_t c = 0; c < C; ++c) {
    const double w = width[c], m = stride[c] - width[c];
    double f = w * w * w / 3.0 + m * w * w;
    for (int b = fblok[c] + 1; b < fblok[c + 1]; ++b) f += 2.0 * (double)(stride[c] - coefind[b]) * nrow[b] * w;
    cost[c] = (lu ? 2.0 : 1.0) * f + 1.0;
    if (fblok[c + 1] - fblok[c] > 1) parent[c] = fcblk[fblok[c] + 1];
  }
  // children lists (parent(c) > c always: cblks are numbered in elimination order)
  std::vector<int> nchild(C + 1, 0), cptr(C + 1, 0), child;
  for (int64_t c = 0; c < C; ++c) { sub[c] = cost[c]; if (parent[c] >= 0) nchild[parent[c]]++; }
  for (int64_t c = 0; c < C; ++c) { if (parent[c] >= 0) sub[parent[c]] += sub[c]; }   // ascending order = children first
  for (int64_t c = 0; c < C; ++c) cptr[c + 1] = cptr[c] + nchild[c];
  child.resize(cptr[C]);
  { std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (int64_t c = 0; c < C; ++c) if (parent[c] >= 0) child[fill[parent[c]]++] = (int)c; }
  if (nranks > 1) {
    // top-down over the forest; roots share [0, nranks) in proportion like children of a virtual root
    struct Item { int c; double lo, hi; int depth; };
    std::vector<Item> stack;
    auto split = [&](std::vector<int> kids, double lo, double hi, int depth) {
      std::sort(kids.begin(), kids.end(), [&](int a, int b) { return sub[a] != sub[b] ? sub[a] > sub[b] : a < b; });
      double tot = 0; for (int k : kids) tot += sub[k];
      double pos = lo;
      for (int k : kids) {
        const double wdt = (hi - lo) * sub[k] / tot;
        stack.push_back({k, pos, pos + wdt, depth});
        pos += wdt;
      }
    };
    std::vector<int> roots;
    for (int64_t c = 0; c < C; ++c) if (parent[c] < 0) roots.push_back((int)c);
    split(roots, 0.0, (double)nranks, 0);
    while (!stack.empty()) {
      const Item it = stack.back(); stack.pop_back();
      // integer candidate interval [a, b): ranks whose unit interval is covered by at least half, at least one
      int a = (int)(it.lo + 0.5), b = (int)(it.hi + 0.5);
      a = std::min(std::max(a, 0), nranks - 1);
      if (b <= a) { a = std::min((int)it.lo, nranks - 1); b = a + 1; }
      b = std::min(b, nranks);
      if (b - a == 1) {
        // whole subtree on rank a
        std::vector<int> st2{it.c};
        while (!st2.empty()) {
          const int c = st2.back(); st2.pop_back();
          P.owner[c] = a;
          for (int q = cptr[c]; q < cptr[c + 1]; ++q) st2.push_back(child[q]);
        }
        continue;
      }
      P.owner[it.c] = a + (it.depth % (b - a));
      std::vector<int> kids(child.begin() + cptr[it.c], child.begin() + cptr[it.c + 1]);
      if (!kids.empty()) split(kids, (double)a, (double)b, it.depth + 1);
    }
  }
  for (int64_t c = 0; c < C; ++c) {
    P.load[P.owner[c]] += cost[c];
    for (int b = fblok[c] + 1; b < fblok[c + 1]; ++b) {
      const int fc = fcblk[b];
      if (P.owner[fc] != P.owner[c]) P.contrib[fc] |= 1u << P.owner[c];
    }
  }
  return P;
}

}  // namespace pb200
