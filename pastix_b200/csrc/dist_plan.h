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
//   owner(c)   = the only candidate, or (shared cblk) dealt over the interval.
// Ways of dealing the shared column blocks (PB200_DIST_CHAIN, read by every rank):
//   default cyclic in elimination order over the candidate interval (1-D cyclic columns; with the fan-out of factored
//           panels — engine.cu — the updates of a shared separator are spread over the GPUs that own their targets);
//   "lpt"   (round 1) every shared cblk on its own, heaviest first to the least-loaded GPU of its interval: best flop
//           balance of the owner-computes-all scheme;
//   "group" the chain of shared cblks of one separator (same candidate interval, linked by parent) stays on ONE GPU:
//           hand-offs only where the elimination tree branches, at the price of a coarser flop balance.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace pb200 {

struct DistPlan {
  int nranks = 1;
  std::vector<int> owner;          // per cblk
  std::vector<uint32_t> contrib;   // per cblk: bit p set <=> rank p != owner holds a cblk with a blok facing it
  std::vector<uint32_t> contrib_priv;   // the same, counting only contributors that are NOT shared cblks (fan-out mode:
                                   // the updates of a shared cblk are computed by the owners of their targets)
  std::vector<char> shared;        // per cblk: 1 = column block of a shared top separator (candidate set > 1 GPU)
  std::vector<double> load;        // per rank: flops mapped to it
};

// fblok[C+1], fcblk[B], width[C], stride[C], nrow[B], coefind[B]
inline DistPlan dist_plan(int64_t C, const int *fblok, const int *fcblk, const int *width, const int *stride,
                          const int *nrow, const int *coefind, int nranks, bool lu, const int32_t *ext_owner = nullptr) {
  DistPlan P;
  P.nranks = nranks;
  P.owner.assign(C, 0);
  P.contrib.assign(C, 0u);
  P.contrib_priv.assign(C, 0u);
  P.shared.assign(C, 0);
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
  const char *mode0 = getenv("PB200_DIST_CHAIN");
  if (nranks > 1 && ext_owner != nullptr) {
    // Mapping handed in by the caller (the reference's own: blend's task-to-thread map, pastix_b200.h pb200_options_t).
    // A cblk whose subtree is not owned by a single rank is a shared top-separator block (its updates are computed by the
    // owners of their targets, fan-out), the rest is private (fan-in).
    std::vector<int> uni(C);
    for (int64_t c = 0; c < C; ++c) { P.owner[c] = std::min(std::max((int)ext_owner[c], 0), nranks - 1); uni[c] = P.owner[c]; }
    for (int64_t c = 0; c < C; ++c)        // ascending = children first
      if (parent[c] >= 0 && uni[parent[c]] != uni[c]) uni[parent[c]] = -1;
    for (int64_t c = 0; c < C; ++c) P.shared[c] = uni[c] < 0;
  } else if (nranks > 1 && (mode0 == nullptr || !strcmp(mode0, "cyclic"))) {
    // Default mapping (round 2).  The elimination tree over cblks is far from the balanced binary tree of the nested
    // dissection: a separator split into 120-column cblks is a CHAIN, and sibling separators hang at different depths
    // (C3: a 124-cblk top chain over subtrees of 0.53 / 0.16 / 0.16 of the flops).  Candidate intervals in proportion
    // to cost then leave one GPU with 0.50 and the other with 0.32 of private work.  Instead: keep a set of private
    // subtrees, place them heaviest-first on the least-loaded GPU (LPT), and while the private loads differ by more
    // than PB200_DIST_EPS (default 6 %) of a fair share, open the heaviest subtree — its root cblk becomes SHARED, its
    // children become subtrees.  Shared cblks are dealt cyclically over all GPUs in elimination order (1-D cyclic
    // columns of the dense top); with the fan-out of factored panels (engine.cu) their updates are computed by the
    // owners of the targets, so chain and update work of every top separator are spread over all GPUs.
    double total = 0;
    for (int64_t c = 0; c < C; ++c) if (parent[c] < 0) total += sub[c];
    double eps = 0.06;
    if (const char *e = getenv("PB200_DIST_EPS")) eps = atof(e);
    std::vector<int> items;
    for (int64_t c = 0; c < C; ++c) if (parent[c] < 0) items.push_back((int)c);
    std::vector<int> shared_list;
    std::vector<double> L(nranks, 0.0);
    std::vector<int> where;
    auto by_cost = [&](int x, int y) { return sub[x] != sub[y] ? sub[x] > sub[y] : x < y; };
    for (;;) {
      std::sort(items.begin(), items.end(), by_cost);
      std::fill(L.begin(), L.end(), 0.0);
      where.assign(items.size(), 0);
      for (size_t i = 0; i < items.size(); ++i) {
        int best = 0;
        for (int p = 1; p < nranks; ++p) if (L[p] < L[best]) best = p;
        where[i] = best; L[best] += sub[items[i]];
      }
      const double mx = *std::max_element(L.begin(), L.end()), mn = *std::min_element(L.begin(), L.end());
      if (items.empty() || mx - mn <= eps * total / nranks) break;
      const int k = items[0];                       // heaviest
      if (cptr[k + 1] == cptr[k]) break;            // a leaf cblk: nothing left to open
      shared_list.push_back(k);
      items.erase(items.begin());
      for (int q = cptr[k]; q < cptr[k + 1]; ++q) items.push_back(child[q]);
    }
    for (size_t i = 0; i < items.size(); ++i) {
      std::vector<int> st2{items[i]};
      while (!st2.empty()) {
        const int c = st2.back(); st2.pop_back();
        P.owner[c] = where[i];
        for (int q = cptr[c]; q < cptr[c + 1]; ++q) st2.push_back(child[q]);
      }
    }
    std::sort(shared_list.begin(), shared_list.end());
    int nx = 0;
    for (int p = 1; p < nranks; ++p) if (L[p] < L[nx]) nx = p;
    for (int k : shared_list) { P.shared[k] = 1; P.owner[k] = nx; nx = (nx + 1) % nranks; }
  } else if (nranks > 1) {
    // Legacy mappings of round 1 (PB200_DIST_CHAIN=lpt|group).
    // Pass 1 (top-down): a subtree heavier than one GPU's fair share stays shared over a candidate interval
    // (several such children split the interval in proportion to their cost); lighter subtrees become
    // single-GPU subtrees, placed in pass 2.  Pass 2: subtrees, heaviest first, go to the least-loaded GPU
    // of their interval; pass 3: the shared column blocks, heaviest first, likewise — they are about half
    // of the flops, so they level what the subtrees left uneven.
    double total = 0;
    for (int64_t c = 0; c < C; ++c) if (parent[c] < 0) total += sub[c];
    // A subtree is split over several GPUs only when it is heavier than a fair share by more than a tolerance
    // (PB200_DIST_TOL, default 0: measured on C3, whose cblk tree is 0.53 / 0.16 / 0.16 under a 124-cblk top chain, splitting wins): the halves of a nested dissection are never exactly equal (100 = 50 + 49 + 1
    // planes), and splitting the slightly heavier one shares its whole top separator chain between the GPUs — a
    // serial chain of hand-offs (measured, C3 on 2 GPUs: 91 such levels) — to correct a 2 % imbalance.
    double tolv = 0.0;
    if (const char *e = getenv("PB200_DIST_TOL")) tolv = atof(e);
    const double fair = (total / nranks) * (1.0 + tolv);
    struct Item { int c, a, b; };
    std::vector<Item> stack, subtrees, shared;
    {
      std::vector<int> roots;
      for (int64_t c = 0; c < C; ++c) if (parent[c] < 0) roots.push_back((int)c);
      // virtual root over the forest
      std::vector<int> big, small;
      for (int r : roots) (sub[r] > fair ? big : small).push_back(r);
      double tot = 0; for (int k : big) tot += sub[k];
      double pos = 0;
      for (int k : big) {
        const double wdt = nranks * sub[k] / tot;
        int lo = (int)(pos + 0.5), hi = (int)(pos + wdt + 0.5);
        lo = std::min(lo, nranks - 1); if (hi <= lo) hi = lo + 1; hi = std::min(hi, nranks);
        stack.push_back({k, lo, hi}); pos += wdt;
      }
      for (int k : small) subtrees.push_back({k, 0, nranks});
    }
    while (!stack.empty()) {
      const Item it = stack.back(); stack.pop_back();
      if (it.b - it.a == 1) { subtrees.push_back(it); continue; }
      shared.push_back(it);
      std::vector<int> big, small;
      for (int q = cptr[it.c]; q < cptr[it.c + 1]; ++q) (sub[child[q]] > fair ? big : small).push_back(child[q]);
      std::sort(big.begin(), big.end(), [&](int x, int y) { return sub[x] != sub[y] ? sub[x] > sub[y] : x < y; });
      double tot = 0; for (int k : big) tot += sub[k];
      double pos = it.a;
      for (int k : big) {
        const double wdt = (it.b - it.a) * sub[k] / tot;
        int lo = (int)(pos + 0.5), hi = (int)(pos + wdt + 0.5);
        lo = std::min(std::max(lo, it.a), it.b - 1); if (hi <= lo) hi = lo + 1; hi = std::min(hi, it.b);
        stack.push_back({k, lo, hi}); pos += wdt;
      }
      for (int k : small) subtrees.push_back({k, it.a, it.b});
    }
    for (const Item &it : shared) P.shared[it.c] = 1;
    std::vector<double> L(nranks, 0.0);
    auto least = [&](int a2, int b2) { int best = a2; for (int p = a2 + 1; p < b2; ++p) if (L[p] < L[best]) best = p; return best; };
    auto place_subtree = [&](const Item &it) {
      const int p = least(it.a, it.b);
      L[p] += sub[it.c];
      std::vector<int> st2{it.c};
      while (!st2.empty()) {
        const int c = st2.back(); st2.pop_back();
        P.owner[c] = p;
        for (int q = cptr[c]; q < cptr[c + 1]; ++q) st2.push_back(child[q]);
      }
    };
    const char *mode = getenv("PB200_DIST_CHAIN");
    if (mode && !strcmp(mode, "group")) {
      // the subtrees that have no choice (one candidate) first, then the chains (the big indivisible items, heaviest
      // first to the least-loaded GPU of their interval), then the free subtrees level what is left uneven
      {
        std::vector<Item> rest;
        for (const Item &it : subtrees) { if (it.b - it.a == 1) place_subtree(it); else rest.push_back(it); }
        subtrees.swap(rest);
      }
      std::vector<int> sa(C, -1), sb(C, -1), gid(C, -1);
      for (const Item &it : shared) { sa[it.c] = it.a; sb[it.c] = it.b; }
      std::sort(shared.begin(), shared.end(), [](const Item &x, const Item &y) { return x.c > y.c; });   // parents first
      struct Group { double cost; int a, b, first; };
      std::vector<Group> groups;
      for (const Item &it : shared) {
        const int p = parent[it.c];
        if (p >= 0 && gid[p] >= 0 && sa[p] == it.a && sb[p] == it.b) gid[it.c] = gid[p];
        else { gid[it.c] = (int)groups.size(); groups.push_back({0.0, it.a, it.b, it.c}); }
        groups[gid[it.c]].cost += cost[it.c];
      }
      std::vector<int> order(groups.size());
      for (size_t k = 0; k < order.size(); ++k) order[k] = (int)k;
      std::sort(order.begin(), order.end(), [&](int x, int y) {
        return groups[x].cost != groups[y].cost ? groups[x].cost > groups[y].cost : groups[x].first < groups[y].first; });
      std::vector<int> gowner(groups.size(), 0);
      for (int k : order) { const int p = least(groups[k].a, groups[k].b); L[p] += groups[k].cost; gowner[k] = p; }
      for (const Item &it : shared) P.owner[it.c] = gowner[gid[it.c]];
      shared.clear();
    }
    std::sort(subtrees.begin(), subtrees.end(), [&](const Item &x, const Item &y) { return sub[x.c] != sub[y.c] ? sub[x.c] > sub[y.c] : x.c < y.c; });
    for (const Item &it : subtrees) place_subtree(it);
    if (mode && !strcmp(mode, "lpt")) {
      // round 1: heaviest first to the least-loaded GPU of the interval (best flop balance of the owner-computes
      // scheme; a whole separator chain may land on the one GPU that was behind)
      std::sort(shared.begin(), shared.end(), [&](const Item &x, const Item &y) { return cost[x.c] != cost[y.c] ? cost[x.c] > cost[y.c] : x.c < y.c; });
      for (const Item &it : shared) {
        const int p = least(it.a, it.b);
        L[p] += cost[it.c];
        P.owner[it.c] = p;
      }
    } else {
      // default: the shared cblks of every candidate interval are dealt CYCLICALLY in elimination order, starting at
      // the least-loaded GPU — the 1-D cyclic column distribution of a dense factorization.  With the fan-out of the
      // factored panels every GPU then updates the cblks it owns, so both the panel chain and the update work of a
      // top separator are spread evenly, level after level.
      std::sort(shared.begin(), shared.end(), [](const Item &x, const Item &y) { return x.c < y.c; });
      std::vector<int> next((size_t)nranks * (nranks + 1), -1);
      for (const Item &it : shared) {
        int &nx = next[(size_t)it.a * (nranks + 1) + it.b];
        if (nx < 0) nx = least(it.a, it.b);
        const int p = nx;
        nx = (nx + 1 - it.a) % (it.b - it.a) + it.a;
        L[p] += cost[it.c];
        P.owner[it.c] = p;
      }
    }
  }
  for (int64_t c = 0; c < C; ++c) {
    // flops mapped to each GPU: a private cblk costs its owner everything; a shared one costs its owner the panel
    // (diagonal block + TRSM) and the owners of its targets their updates (fan-out)
    if (!P.shared[c]) P.load[P.owner[c]] += cost[c];
    else {
      const double w = width[c], m = stride[c] - width[c], f = lu ? 2.0 : 1.0;
      P.load[P.owner[c]] += f * (w * w * w / 3.0 + m * w * w);
      for (int b = fblok[c] + 1; b < fblok[c + 1]; ++b)
        P.load[P.owner[fcblk[b]]] += f * 2.0 * (double)(stride[c] - coefind[b]) * nrow[b] * w;
    }
    for (int b = fblok[c] + 1; b < fblok[c + 1]; ++b) {
      const int fc = fcblk[b];
      if (P.owner[fc] != P.owner[c]) {
        P.contrib[fc] |= 1u << P.owner[c];
        if (!P.shared[c]) P.contrib_priv[fc] |= 1u << P.owner[c];
      }
    }
  }
  return P;
}

}  // namespace pb200
