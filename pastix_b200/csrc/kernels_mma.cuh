// kernels_mma.cuh — the double-precision hot path: FP64 tensor-core (DMMA)
// kernels for real and complex double.
//
// A column block (cblk) wider than NBMAX columns is processed as a chain of
// sub-panels J = [c0, c1) of its own columns (right-looking inside the cblk):
//   k_diag_sub     factor the nb x nb block (c0,c0)            (factor_diag,   compute_diag.c:538)
//   k_trsm_mma     rows [c1, stride) x J  <-  * W^{-T}          (factor_trsm1d, compute_trsm.c:128)
//   k_gemm_scatter mode INT: rows [c1, stride) x cols [c1, w) of the cblk's own panel
//                  -= P[., J] * P[c1:w, J]^T                    (trailing update of PASTIX_*_block,
//                                                                compute_diag.c:171-203, 262-307, 486-518)
// and, once the whole cblk is factored,
//   k_gemm_scatter mode EXT: the fused "GEMM + compacted scatter" of the reference
//                  C = P[w:, 0:w] * P[w:, 0:w]^T, tile by tile straight from the DMMA accumulators into
//                  the facing cblks (compute_1dgemm = compute_contrib_compact + add_contrib_local,
//                  sopalin_compute.c:270-598, 865-1032).  The reference's maxbloktab work buffer and
//                  the AXPY scatter (dim_dgeam) never touch memory here.
// LU keeps two panels per cblk, coeftab (L, with the full square A_kk as diagonal blok) and ucoeftab
// (U^T); while a cblk is being factored ucoeftab's diagonal blok holds the transpose of coeftab's, so
// both panels go through the same kernels with the B operand taken from the other panel.  At the end
// ucoeftab's diagonal blok is (LU)^T, which is what DimTrans leaves (compute_diag.c:521-536, 598-603).
#pragma once
#include "mma.cuh"
#include "symbol.cuh"
#include "kernels_factor.cuh"

namespace pb200 {

// per off-diagonal blok: where its rows land as COLUMNS of the facing cblk (built in pb200_create)
struct __align__(16) BlokTgt {
  int64_t tgt;   // poff[fc] + (frow(b) - fcol(fc)) * stride(fc): slab offset of the target column of b's first row
  int tld;       // stride(fc)
  int tw;        // width(fc)
  int cj0;       // frow(b) - fcol(fc)
  int fc;        // facing cblk
};

// extra device-side maps built once per SolverMatrix (pb200_create)
// static scatter maps, one entry per off-diagonal panel row of every cblk (built once, k_build_maps):
// what add_contrib_local derives per contribution from frownum/coefind (sopalin_compute.c:427-435)
struct RowMap { int rb, roff; };            // local off-diagonal blok of the row, offset of the row inside it
struct __align__(16) ColMap {               // the same panel row seen as a COLUMN of the facing cblk
  int64_t ctgt;                             // slab offset of that target column: poff[fc] + cj * tld
  int cb, cj, tw, tld;                      // local blok, column inside fc, width(fc), stride(fc)
  int pad0, pad1;
};
struct DevMap {
  const int64_t *pairbase;      // per cblk: start of its (b2,b1) table in pairoff
  const int *pairoff;           // tri(lb2,lb1): row offset of blok b2's first row inside fcblk(b1), -1 if none
  const BlokTgt *btgt;          // per blok
  const int64_t *rmbase;        // per cblk: first entry of its rows in rm / cm (entry index = rmbase[k] + m - w)
  const RowMap *rm;
  const ColMap *cm;
};
// everything one CTA of k_gemm_scatter needs to start, in ONE load (built once per launch, k_build_tiledesc)
struct __align__(16) TileDesc {
  int64_t poff, pbase, rmrow;   // panel offset of the source cblk; pairbase; rmbase - width (index by panel row)
  int ld, m0, mrows, n0, ncols, k0, k1, mode;
  int rb_lo, nrb, cb_lo, ncb;   // local blok ranges covered by the tile's rows / columns
};

struct GemmTask {
  int cblk;
  int tile0, ntn;       // first tile of this row tile inside the launch, number of column tiles
  int arow0, arow1;     // panel rows of A handled by this row tile: [arow0, min(arow0+TM, arow1))
  int brow0, brow1;     // panel rows forming the columns of C
  int k0, k1;           // panel columns contracted
  int mode;             // 0 = EXT (scatter into facing cblks), 1 = INT (own panel)
  int rbl;              // EXT: last blok (absolute index) touched by this row tile
  int pad;
};
struct SubTask {        // diag / trsm work on sub-panel [c0,c1) of a cblk
  int cblk, tile0, c0, c1;
};

template <class T> struct UpdCfg;
// tile shape of the real-double update kernel (tuning variants: python -m pastix_b200.build -DPB200_UPD_TM=64 ...)
#ifndef PB200_UPD_D_TM
// measured (r01, C3): 128x64/256 thr/2 CTAs 371 ms; 64x64/128 thr/4 CTAs/2 stages 341 ms; 64x64/3 CTAs/3 stages 370 ms;
// 128x64 KC=32 2 stages 360 ms.  Four small CTAs per SM give the one DMMA pipe four independent phase contexts.
#define PB200_UPD_D_TM 64
#define PB200_UPD_D_TN 64
#define PB200_UPD_D_KC 16
#define PB200_UPD_D_STG 2
#define PB200_UPD_D_WM 2
#define PB200_UPD_D_WN 2
#define PB200_UPD_D_CTAS 4
#endif
// operand staging of k_gemm_scatter: 1 = TMA bulk copies (cp.async.bulk, one per panel-column segment, completion on an
// mbarrier; no LSU work and no address arithmetic in the compute warps), 0 = 8-byte cp.async (LDGSTS) by every thread
#ifndef PB200_UPD_TMA
#define PB200_UPD_TMA 1
#endif
template <> struct UpdCfg<double> {
  static constexpr int TM = PB200_UPD_D_TM, TN = PB200_UPD_D_TN, KC = PB200_UPD_D_KC, STG = PB200_UPD_D_STG,
                       WM = PB200_UPD_D_WM, WN = PB200_UPD_D_WN, PADA = 4, PADB = 4;
  static constexpr int NT = WM * WN * 32, CTAS = PB200_UPD_D_CTAS;
  static constexpr bool TMA = PB200_UPD_TMA != 0;
  // a column segment is copied from the 16-byte aligned address at or just below its first element (panel columns are
  // only 8-byte aligned when the stride is odd): TM + 2 elements, the consumer shifts its row index by the parity
  static constexpr int CPY = TM + 2;
};
template <> struct UpdCfg<cdouble> {
  static constexpr int TM = 64, TN = 64, KC = 16, STG = 3, WM = 4, WN = 2, PADA = 2, PADB = 2;
  static constexpr int NT = WM * WN * 32, CTAS = 2;
  static constexpr bool TMA = PB200_UPD_TMA != 0;
  static constexpr int CPY = TM;   // 16-byte elements: always aligned
};
// single precision: the same kernels on TF32 tensor cores with the 3xTF32 split (mma.cuh)
template <> struct UpdCfg<float> {
  static constexpr int TM = 64, TN = 64, KC = 16, STG = 2, WM = 2, WN = 2, PADA = 4, PADB = 4;
  static constexpr int NT = WM * WN * 32, CTAS = 4;
  static constexpr bool TMA = PB200_UPD_TMA != 0;
  static constexpr int CPY = TM + 4;   // 4-byte elements: the aligned address lies up to 3 elements below
};
template <> struct UpdCfg<cfloat> {
  static constexpr int TM = 64, TN = 64, KC = 16, STG = 3, WM = 4, WN = 2, PADA = 2, PADB = 2;
  static constexpr int NT = WM * WN * 32, CTAS = 2;
  static constexpr bool TMA = PB200_UPD_TMA != 0;
  static constexpr int CPY = TM + 2;
};
#ifndef PB200_TABMAX
#define PB200_TABMAX 1536
#endif
#define PB200_COEFMAX 1024

template <class T>
constexpr size_t upd_smem_bytes() {
  using C = UpdCfg<T>;
  return (size_t)C::STG * C::KC * ((C::TM + C::PADA) + (C::TN + C::PADB) + 1) * sizeof(T) +
         (size_t)C::TN * sizeof(ColMap) + (size_t)C::TM * sizeof(RowMap) + (size_t)PB200_TABMAX * 4 + (size_t)C::STG * 8;
}

// fire-and-forget reductions (RED.ADD.F64 at L2): the reference serialises these adds with
// mutex_blok (sopalin_compute.c:563-580)
// (the sign is flipped on the integer pipe: a DADD would compete with the DMMAs for the FP64 pipe)
// (inline PTX on the high word: a plain xor of the bit pattern is recognised by the compiler and comes back as DADD -RZ, -R)
__device__ __forceinline__ double neg_bits(double v) {
  double r;
  asm("{\n.reg .b32 lo, hi;\nmov.b64 {lo, hi}, %1;\nxor.b32 hi, hi, 0x80000000;\nmov.b64 %0, {lo, hi};\n}" : "=d"(r) : "d"(v));
  return r;
}
__device__ __forceinline__ void red_sub(double *p, double v) { atomicAdd(p, neg_bits(v)); }
__device__ __forceinline__ void red_sub(cdouble *p, cdouble v) { atomicAdd(&p->x, neg_bits(v.x)); atomicAdd(&p->y, neg_bits(v.y)); }
__device__ __forceinline__ float neg_bits(float v) { return __int_as_float(__float_as_int(v) ^ (int)0x80000000); }
__device__ __forceinline__ void red_sub(float *p, float v) { atomicAdd(p, neg_bits(v)); }
__device__ __forceinline__ void red_sub(cfloat *p, cfloat v) { atomicAdd(&p->x, neg_bits(v.x)); atomicAdd(&p->y, neg_bits(v.y)); }

// last index i in [0, n) with key[i] <= v (keys ascending, key[0] <= v assumed)
__device__ __forceinline__ int upper_le_s(const int *key, int n, int v) {
  int l = 0, h = n;
  while (l < h) {
    const int mid = (l + h) >> 1;
    if (key[mid] <= v) l = mid + 1; else h = mid;
  }
  return l - 1;
}

template <int BYTES>
__device__ __forceinline__ void cp_async_raw(void *smem_dst, const void *gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(gmem_src), "n"(BYTES));
}

// Programmatic dependent launch (sm_90+): the kernels of the panel chain let the NEXT launch of their stream be scheduled
// while they run (its CTAs become resident, read their static task record and stop at pdl_wait) and themselves wait for
// the previous launch to have completed — memory included — before touching panels.  Without the launch attribute
// (engine.cu, launch_chain) both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class T, int FACTO>
__global__ void __launch_bounds__(UpdCfg<T>::NT, UpdCfg<T>::CTAS)
k_gemm_scatter(DevMap M, T *L, T *U, const T *__restrict__ W, const TileDesc *__restrict__ descs) {
  using C = UpdCfg<T>;
  using RT = typename ST<T>::real;
  constexpr bool CX = ST<T>::is_complex;
  constexpr int TM = C::TM, TN = C::TN, KC = C::KC, STG = C::STG, NT = C::NT;
  constexpr int LDA = TM + C::PADA, LDB = TN + C::PADB;
  constexpr int MI = TM / C::WM / 16, NI = TN / C::WN / 8;
  // LDLt / LDLh: the B operand is read from W, the L*D copy the TRSM kernel leaves beside the panels (the
  // reference's maxbloktab1, compute_trsm.c:86-114; sopalin_compute.c:356-370) — no per-fragment scaling
  constexpr bool SYM_LDL = (FACTO == F_LDLT || FACTO == F_LDLH);
  constexpr bool CONJB = (FACTO == F_LDLH || FACTO == F_LLT);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sA = reinterpret_cast<T *>(smem_raw);
  T *sB = sA + STG * KC * LDA;
  T *sD = sB + STG * KC * LDB;
  ColMap *s_cm = reinterpret_cast<ColMap *>(sD + STG * KC);
  RowMap *s_rm = reinterpret_cast<RowMap *>(s_cm + TN);
  int tile = blockIdx.x, part = 0;
  if (FACTO == F_LU) { part = tile & 1; tile >>= 1; }
  // the whole tile description in one (broadcast) load: no dependent index chasing before the first
  // operand bytes are requested
  pdl_launch_dependents();
  const TileDesc tk = descs[tile];
  pdl_wait();
  const int ld = tk.ld;
  const int m0 = tk.m0, mrows = tk.mrows;
  const int n0 = tk.n0, ncols = tk.ncols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp / C::WN) * (MI * 16), wn0 = (warp % C::WN) * (NI * 8);

  const T *Ap = ((FACTO == F_LU && part == 1) ? U : L) + tk.poff;
  const T *Bp = (SYM_LDL ? W : ((FACTO == F_LU && part == 0) ? U : L)) + tk.poff;
  const int nchunks = (tk.k1 - tk.k0 + KC - 1) / KC;

  int *s_tab = reinterpret_cast<int *>(s_rm + TM);
  const int ktot = tk.k1 - tk.k0;
  Acc<CX, RT> acc[MI][NI];
#pragma unroll
  for (int a = 0; a < MI; ++a)
#pragma unroll
    for (int b = 0; b < NI; ++b) acc[a][b].zero();

  // ---- scatter maps of this tile: static tables copied asynchronously into shared memory (cp.async group 0; they
  // land while the operand pipeline fills and are published by the main loop's first barrier)
  const int rb_lo = tk.rb_lo, cb_lo = tk.cb_lo, ncb = tk.ncb;
  const bool tab_in_smem = (tk.mode == 0) && (tk.nrb * ncb <= PB200_TABMAX);
  if (tk.mode == 0) {
    if (tid < mrows) cp_async_raw<8>(s_rm + tid, M.rm + tk.rmrow + m0 + tid);
    if (tid < ncols) {
      const ColMap *src = M.cm + tk.rmrow + n0 + tid;
      cp_async_raw<16>(s_cm + tid, src);
      cp_async_raw<16>(reinterpret_cast<char *>(s_cm + tid) + 16, reinterpret_cast<const char *>(src) + 16);
    }
    if (tab_in_smem)
      for (int e = tid; e < tk.nrb * ncb; e += NT) {
        const int rb = rb_lo + e / ncb, cb = cb_lo + e % ncb;
        if (rb >= cb) cp_async_raw<4>(s_tab + e, M.pairoff + tk.pbase + (int64_t)rb * (rb + 1) / 2 + cb);
        else s_tab[e] = -1;
      }
  } else {
    if (tid < TM) { s_rm[tid].rb = 0; s_rm[tid].roff = m0 + tid; }
    if (tid < TN) {
      const int n = n0 + tid;
      ColMap cmv; cmv.ctgt = tk.poff + (int64_t)n * ld; cmv.cb = 0; cmv.cj = n; cmv.tw = 0; cmv.tld = ld; cmv.pad0 = cmv.pad1 = 0;
      s_cm[tid] = cmv;
    }
  }

  if constexpr (C::TMA) {
    // ---- TMA-staged operands.  One bulk copy per (operand, panel column) of a chunk (columns 0..KC-1 of A, then of
    // B), completing on the stage's mbarrier.  Rows past the
    // tile's extent are whatever follows in the slab: they only reach accumulator rows / columns that are never
    // written.  Columns past k1 are not copied; the partial 8-column step is zero-filled once.
    static_assert(STG >= 2 && (2 * KC) % (NT / 32) == 0, "two stages at least; the column copies divide evenly among the warps");
    static_assert((LDA * sizeof(T)) % 16 == 0 && (LDB * sizeof(T)) % 16 == 0 && LDA >= C::CPY && LDB >= C::CPY, "bulk copy alignment");
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_tab + PB200_TABMAX);
    constexpr unsigned CPYB = C::CPY * sizeof(T);
    // offset (in elements, below 16 bytes) of the first element of column k0 of each operand, and of the stride
    constexpr int EPM = 16 / (int)sizeof(T) - 1;   // 0, 1 or 3
    const int ldpar = ld & EPM;
    const int64_t eA = tk.poff + (int64_t)tk.k0 * ld + m0, eB = tk.poff + (int64_t)tk.k0 * ld + n0;
    const int parA = (int)(eA & EPM), parB = (int)(eB & EPM);
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < STG; ++s) mbar_init(bar + s, 1);
      mbar_fence_init();
    }
    // columns [ktot, roundup8(ktot)) of the last chunk are zero-filled (generic stores by every thread) when its
    // stage is free: before the pipeline starts if the chunk is part of the prologue, else right before its copies
    // are issued — at least one CTA barrier separates the stores from the fragment loads either way
    const int cl = nchunks - 1, kz0 = ktot - cl * KC, kz1 = (kz0 + 7) & ~7;
    auto zero_tail = [&]() {
      T *a = sA + (cl % STG) * KC * LDA + kz0 * LDA, *b = sB + (cl % STG) * KC * LDB + kz0 * LDB;
      for (int e = tid; e < (kz1 - kz0) * LDA; e += NT) a[e] = ST<T>::zero();
      for (int e = tid; e < (kz1 - kz0) * LDB; e += NT) b[e] = ST<T>::zero();
    };
    if (kz1 > kz0 && cl < STG) zero_tail();
    cp_async_commit();
    __syncthreads();
    // every warp issues its share of the chunk's 2*KC column copies (lanes 0 .. CPW-1, one copy each: the uniform-datapath
    // UBLKCP is issued once per active lane); thread 0 posts the byte count (the mbarrier's tx-count is signed, copies
    // completing before the expect_tx are fine)
    constexpr int CPW = 2 * KC / (NT / 32);
    auto issue = [&](int c) {
      const int stg = c % STG, nk = min(KC, ktot - c * KC);
      if (tid == 0) mbar_arrive_expect_tx(bar + stg, 2u * (unsigned)nk * CPYB);
      const int q = warp * CPW + lane, kk = q & (KC - 1), isb = q / KC;
      if (lane < CPW && kk < nk) {
        const int k = c * KC + kk;
        const int par = ((isb ? parB : parA) + k * ldpar) & EPM;
        const T *src = (isb ? Bp : Ap) + (size_t)(tk.k0 + k) * ld + (isb ? n0 : m0) - par;
        T *dst = (isb ? sB + stg * KC * LDB + kk * LDB : sA + stg * KC * LDA + kk * LDA);
        bulk_g2s(dst, src, CPYB, bar + stg);
      }
    };
    for (int c = 0; c < STG && c < nchunks; ++c) issue(c);
    // this thread's fragment columns are k = (even) + t4 (+4): one parity per operand for the whole tile
    const int t4f = lane & 3;
    const int pa = (parA + t4f * ldpar) & EPM, pb = (parB + t4f * ldpar) & EPM;   // (k - t4 is a multiple of 4)
    cp_async_wait<0>();
    for (int c = 0; c < nchunks; ++c) {
      const int stg = c % STG;
      mbar_wait(bar + stg, (unsigned)((c / STG) & 1));
      const T *a = sA + stg * KC * LDA + pa, *b = sB + stg * KC * LDB + pb;
      const int nk8 = min(KC, (ktot - c * KC + 7) & ~7);
#pragma unroll
      for (int ks = 0; ks < KC; ks += 8) {
        if (ks < nk8) {
          FragA<CX, RT> fa[MI];
          FragB<CX, RT> fb[NI];
#pragma unroll
          for (int x = 0; x < MI; ++x) load_frag_a<T>(fa[x], a, LDA, wm0 + x * 16, ks, lane);
#pragma unroll
          for (int y = 0; y < NI; ++y) load_frag_b<T, CONJB, false>(fb[y], b, LDB, wn0 + y * 8, ks, lane, nullptr);
#pragma unroll
          for (int x = 0; x < MI; ++x)
#pragma unroll
            for (int y = 0; y < NI; ++y) mma_acc(acc[x][y], fa[x], fb[y]);
        }
      }
      __syncthreads();   // stage free again (and, first time round, the scatter maps are published)
      if (c + STG == cl && kz1 > kz0) zero_tail();
      if (c + STG < nchunks) issue(c + STG);
    }
  } else {
  // operand staging: thread (lk0, li) = (tid / TM, tid % TM) copies element li of the columns k = lk0, lk0 + KSTEP, ...
  // of a chunk; everything that does not change from chunk to chunk (row predicate, column-k0 addresses) is
  // resolved once per tile, so that a chunk costs one 64-bit add and one predicate per cp.async
  static_assert(TM == TN && NT % TM == 0 && KC % (NT / TM) == 0, "operand staging assumes square tiles");
  constexpr int KSTEP = NT / TM;
  const int li = tid % TM, lk0 = tid / TM;
  const bool a_ok = li < mrows, b_ok = li < ncols;
  const T *pA = Ap + (size_t)tk.k0 * ld + m0 + (a_ok ? li : 0);
  const T *pB = Bp + (size_t)tk.k0 * ld + n0 + (b_ok ? li : 0);
  auto load_chunk = [&](int c, int stg) {
    const int kb = c * KC + lk0;                       // first column of this thread, relative to k0
    T *a = sA + stg * KC * LDA + lk0 * LDA + li, *b = sB + stg * KC * LDB + lk0 * LDB + li;
    const T *ga = pA + (size_t)kb * ld, *gb = pB + (size_t)kb * ld;
#pragma unroll
    for (int q = 0; q < KC / KSTEP; ++q) {
      const bool kin = kb + q * KSTEP < ktot;
      cp_async_elem<sizeof(T)>(a + q * KSTEP * LDA, (a_ok && kin) ? ga + (size_t)(q * KSTEP) * ld : pA, a_ok && kin);
      cp_async_elem<sizeof(T)>(b + q * KSTEP * LDB, (b_ok && kin) ? gb + (size_t)(q * KSTEP) * ld : pB, b_ok && kin);
    }
  };
#pragma unroll
  for (int s = 0; s < STG - 1; ++s) {
    if (s < nchunks) load_chunk(s, s);
    cp_async_commit();
  }

  // ---- main loop: C(TM x TN) = A(TM x K) * B(TN x K)^T on DMMA
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<STG - 2>();
    __syncthreads();
    if (c + STG - 1 < nchunks) load_chunk(c + STG - 1, (c + STG - 1) % STG);
    cp_async_commit();
    const int stg = c % STG;
    const T *a = sA + stg * KC * LDA, *b = sB + stg * KC * LDB;
#pragma unroll
    for (int ks = 0; ks < KC; ks += 8) {
      FragA<CX, RT> fa[MI];
      FragB<CX, RT> fb[NI];
#pragma unroll
      for (int x = 0; x < MI; ++x) load_frag_a<T>(fa[x], a, LDA, wm0 + x * 16, ks, lane);
#pragma unroll
      for (int y = 0; y < NI; ++y) load_frag_b<T, CONJB, false>(fb[y], b, LDB, wn0 + y * 8, ks, lane, nullptr);
#pragma unroll
      for (int x = 0; x < MI; ++x)
#pragma unroll
        for (int y = 0; y < NI; ++y) mma_acc(acc[x][y], fa[x], fb[y]);
    }
  }
  cp_async_wait<0>();
  if (nchunks == 0) __syncthreads();   // (never: K >= 1) maps are published by the main loop's first barrier
  }

  // ---- epilogue: subtract the tile from its targets straight from the accumulators with
  // fire-and-forget L2 reductions: nothing is read back, so no latency is exposed here.
  const int g = lane >> 2, t4 = lane & 3;
  T *TA = ((FACTO == F_LU && part == 1) ? U : L);   // slab updated by the "normal" write of this part
  // this thread's NI*2 columns: resolved once
  int c_cb[NI][2], c_cj[NI][2], c_tw[NI][2];
  int64_t c_tgt[NI][2];
#pragma unroll
  for (int y = 0; y < NI; ++y)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = wn0 + y * 8 + t4 * 2 + e;
      const bool ok = j < ncols;
      const ColMap cmv = s_cm[ok ? j : 0];
      c_cb[y][e] = ok ? cmv.cb : 0x7fffffff;     // sentinel: never <= a row blok
      c_cj[y][e] = ok ? cmv.cj : 0x7fffffff;
      c_tw[y][e] = cmv.tw;
      c_tgt[y][e] = cmv.ctgt;
    }
  auto value = [&](int x, int y, int hh, int e) -> T {
    if constexpr (CX) return T(acc[x][y].re[hh * 2 + e], acc[x][y].im[hh * 2 + e]);
    else return acc[x][y].re[hh * 2 + e];
  };
  if (tk.mode == 1) {
    // own panel: symmetric variants only keep the lower triangle of the diagonal blok
#pragma unroll
    for (int x = 0; x < MI; ++x)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int i = wm0 + x * 16 + g + hh * 8;
        if (i >= mrows) continue;
        const int roff = s_rm[i].roff;
#pragma unroll
        for (int y = 0; y < NI; ++y)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (c_cj[y][e] != 0x7fffffff && (FACTO == F_LU || roff >= c_cj[y][e])) {
              if constexpr (CX && FACTO == F_LLT) {
                // the reference's trailing update of a complex LLt block is zherk (sopalin_compute.h:178-179): the
                // diagonal entry becomes Re(c_jj) - sum |a_jl|^2, its imaginary part is dropped.  Each diagonal
                // entry belongs to one tile of the launch and nothing else touches it meanwhile: plain store.
                if (roff == c_cj[y][e]) {
                  T *p = TA + c_tgt[y][e] + roff;
                  atomicAdd(&p->x, neg_bits(value(x, y, hh, e).x));
                  p->y = RT(0);
                  continue;
                }
              }
              red_sub(TA + c_tgt[y][e] + roff, value(x, y, hh, e));
            }
      }
    return;
  }
  // facing cblks: (row blok rb, column blok cb) -> row offset of rb inside the cblk facing cb (-1: no such pair);
  // one lookup + one test per element, the column sentinel (cb = INT_MAX) fails rb >= cb
  auto scatter = [&](auto lookup) {
#pragma unroll
    for (int x = 0; x < MI; ++x)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int i = wm0 + x * 16 + g + hh * 8;
        if (i >= mrows) continue;
        const int rb = s_rm[i].rb, roff = s_rm[i].roff;
#pragma unroll
        for (int y = 0; y < NI; ++y)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int cb = c_cb[y][e];
            int ro = (rb >= cb) ? lookup(rb, cb) : -1;
            if (ro < 0) continue;
            ro += roff;
            const T v = value(x, y, hh, e);
            if (FACTO != F_LU || part == 0 || ro >= c_tw[y][e]) {
              red_sub(TA + c_tgt[y][e] + ro, v);
            } else if (rb != cb) {
              // U contribution to a diagonal target: stored transposed into coeftab
              // (sopalin_compute.c:431-435, 572-575); the b1 == b2 square is skipped
              const int j = wn0 + y * 8 + t4 * 2 + e;
              const int tld = s_cm[j].tld;
              red_sub(L + (c_tgt[y][e] - (int64_t)c_cj[y][e] * tld) + (int64_t)ro * tld + c_cj[y][e], v);
            }
          }
      }
  };
  if (tab_in_smem) {
    const int tbase = -rb_lo * ncb - cb_lo;
    scatter([&](int rb, int cb) { return s_tab[tbase + rb * ncb + cb]; });
  } else {
    scatter([&](int rb, int cb) { return M.pairoff[tk.pbase + (int64_t)rb * (rb + 1) / 2 + cb]; });
  }
}

// ---------------------------------------------------------------- panel TRSM on DMMA
// X (rows [c1, stride) x cols J) <- X * W^{-T}, W = lower triangle of the factored nb x nb block.
//   LLt  : W = L_JJ                          non-unit
//   LDLt : W = L_JJ  unit, then X <- X D^{-1} (LDLh: conj(L_JJ))
//   LU   : part 0  X = coeftab rows,  W = lower(ucoeftab JJ) = U_JJ^T, non-unit   (L <- L U^{-1})
//          part 1  X = ucoeftab rows, W = lower(coeftab JJ)  = L_JJ,   unit       (U^T <- U^T L^{-T})
// One CTA owns TRSM_TM rows and all nb columns; each warp owns 16 rows and walks the 8-column blocks
// left to right: T = X_jb - sum_{kb<jb} Y_kb W[jb,kb]^T, Y_jb = T inv(W[jb,jb])^T, both on DMMA.
#define PB200_TRSM_TM 64
template <class T> struct SubCfg;
// widest sub-panel factored in one diag -> TRSM round.  Measured in round 2 (profiles/r02/README.md): with the blocked
// diagonal kernel a whole 121-column cblk in ONE round (128) costs diag 46 us + TRSM 23 us against 2 x (21 + 13) + 13 us
// for two rounds of 64 — no gain on the chain, and the 214 KB / one-CTA-per-SM TRSM of 128 columns loses on the fat
// levels (C3: 302 ms vs 286 ms).  64 stays the default; -DPB200_NBMAX_D=128 builds the single-round variant.
#ifndef PB200_NBMAX_D
#define PB200_NBMAX_D 64
#endif
template <> struct SubCfg<double> { static constexpr int NBMAX = PB200_NBMAX_D, PADW = 4, PADX = 4; };
template <> struct SubCfg<cdouble> { static constexpr int NBMAX = 64, PADW = 2, PADX = 2; };
template <> struct SubCfg<float> { static constexpr int NBMAX = 64, PADW = 4, PADX = 4; };
template <> struct SubCfg<cfloat> { static constexpr int NBMAX = 64, PADW = 2, PADX = 2; };

template <class T>
__host__ __device__ constexpr int trsm_ldw(int nbp) {
  // LDW = nbp + pad with (LDW mod 16) == 4 for 8-byte and (LDW mod 8) == 2 for 16-byte elements
  return nbp + SubCfg<T>::PADW;
}
template <class T>
inline size_t trsm_smem_bytes(int nb) {
  const int nbp = (nb + 7) & ~7;
  return ((size_t)nbp * trsm_ldw<T>(nbp) + (size_t)nbp * (PB200_TRSM_TM + SubCfg<T>::PADX) + (size_t)nbp * 9) * sizeof(T);
}

template <class T, int FACTO>
__global__ void __launch_bounds__(128)
k_trsm_mma(DevSym S, T *L, T *U, T *W, const SubTask *__restrict__ tasks, int ntasks) {
  using RT = typename ST<T>::real;
  constexpr bool CX = ST<T>::is_complex;
  constexpr int TM = PB200_TRSM_TM, LDX = TM + SubCfg<T>::PADX;
  constexpr bool UNIT_SYM = (FACTO == F_LDLT || FACTO == F_LDLH);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int tile = blockIdx.x, part = 0;
  if (FACTO == F_LU) { part = tile & 1; tile >>= 1; }
  const int ti = find_task(tasks, ntasks, tile);
  pdl_launch_dependents();
  const SubTask tk = tasks[ti];
  pdl_wait();
  const int k = tk.cblk, ld = S.stride[k], c0 = tk.c0, nb = tk.c1 - tk.c0, nbp = (nb + 7) & ~7;
  const int LDW = trsm_ldw<T>(nbp);
  const int r_base = tk.c1 + (tile - tk.tile0) * TM, mrows = min(TM, ld - r_base);
  T *Ws = reinterpret_cast<T *>(smem_raw);
  T *Xs = Ws + (size_t)nbp * LDW;
  T *sInv = Xs + (size_t)nbp * LDX;   // [nbp/8][k*8+n] = Inv_jb[n][k]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool unit = UNIT_SYM || (FACTO == F_LU && part == 1);
  T *Xp = ((FACTO == F_LU && part == 1) ? U : L) + S.poff[k];
  const T *Wp = ((FACTO == F_LU && part == 0) ? U : L) + S.poff[k] + (size_t)c0 * (ld + 1);
  const T one = ST<T>::from_real(1.0), zero = ST<T>::zero();

  // stage W (lower triangle of the nb x nb block; the upper part is never read) and the X row tile
  // with cp.async; padding rows/columns are zero-filled, padded diagonal = 1
  T *rdiag = sInv + (size_t)nbp * 8;   // reciprocals of W's diagonal
  {
    const int n = tid;
    if (n < nbp)
      for (int kk = 0; kk <= n; ++kk)
        cp_async_elem<sizeof(T)>(Ws + (size_t)kk * LDW + n, Wp + (size_t)(n < nb ? kk : 0) * ld + (n < nb ? n : 0), n < nb);
    const int i = tid & (TM - 1);
    for (int kk = tid / TM; kk < nbp; kk += 128 / TM) {
      const bool ok = (i < mrows && kk < nb);
      cp_async_elem<sizeof(T)>(Xs + (size_t)kk * LDX + i, Xp + (size_t)(c0 + (ok ? kk : 0)) * ld + r_base + (ok ? i : 0), ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    if (n < nbp) {
      T dv = one;
      if (n < nb && !unit) {
        dv = Ws[(size_t)n * LDW + n];
        if (FACTO == F_LDLH) dv = ST<T>::conj(dv);
        dv = one / dv;
      }
      if (n >= nb) Ws[(size_t)n * LDW + n] = one;
      rdiag[n] = dv;
    }
  }
  __syncthreads();
  // inverses of the 8x8 diagonal blocks: one thread per (block, column)
  for (int e = tid; e < nbp; e += 128) {
    const int jb = e >> 3, c = e & 7;
    const T *Wd = Ws + (size_t)(jb * 8) * LDW + jb * 8;   // Wd[r][q] at q*LDW + r
    const T *rd = rdiag + jb * 8;
    T *inv = sInv + jb * 64;
    T x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = zero;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r == c) x[r] = rd[r];
      else if (r > c) {
        T sacc = zero;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q >= c && q < r) {
            T wv = Wd[(size_t)q * LDW + r];
            if (FACTO == F_LDLH) wv = ST<T>::conj(wv);
            sacc += wv * x[q];
          }
        x[r] = (zero - sacc) * rd[r];
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) inv[c * 8 + r] = x[r];   // Inv[n=r][k=c] stored at [k*8 + n]
  }
  __syncthreads();

  const int r0 = warp * 16, g = lane >> 2, t4 = lane & 3;
  if (r0 < mrows) {
    for (int jb = 0; jb < nbp / 8; ++jb) {
      Acc<CX, RT> a0, a1;
      a0.zero(); a1.zero();
      FragA<CX, RT> fa; FragB<CX, RT> fb;
      int kb = 0;
      for (; kb + 1 < jb; kb += 2) {
        load_frag_a<T>(fa, Xs, LDX, r0, kb * 8, lane);
        load_frag_b<T, FACTO == F_LDLH, false>(fb, Ws, LDW, jb * 8, kb * 8, lane, nullptr);
        mma_acc(a0, fa, fb);
        load_frag_a<T>(fa, Xs, LDX, r0, kb * 8 + 8, lane);
        load_frag_b<T, FACTO == F_LDLH, false>(fb, Ws, LDW, jb * 8, kb * 8 + 8, lane, nullptr);
        mma_acc(a1, fa, fb);
      }
      if (kb < jb) {
        load_frag_a<T>(fa, Xs, LDX, r0, kb * 8, lane);
        load_frag_b<T, FACTO == F_LDLH, false>(fb, Ws, LDW, jb * 8, kb * 8, lane, nullptr);
        mma_acc(a0, fa, fb);
      }
      // T = X_jb - acc, written back in place (C layout), then re-read as an A fragment
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = r0 + g + ((q & 2) ? 8 : 0), kc = jb * 8 + t4 * 2 + (q & 1);
        T *p = Xs + (size_t)kc * LDX + i;
        T s;
        if constexpr (CX) s = T(a0.re[q] + a1.re[q], a0.im[q] + a1.im[q]);
        else s = a0.re[q] + a1.re[q];
        *p = *p - s;
      }
      __syncwarp();
      load_frag_a<T>(fa, Xs, LDX, r0, jb * 8, lane);
      load_frag_b<T, false, false>(fb, sInv + jb * 64, 8, 0, 0, lane, nullptr);
      a0.zero();
      mma_acc(a0, fa, fb);
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = r0 + g + ((q & 2) ? 8 : 0), kc = jb * 8 + t4 * 2 + (q & 1);
        T s;
        if constexpr (CX) s = T(a0.re[q], a0.im[q]);
        else s = a0.re[q];
        Xs[(size_t)kc * LDX + i] = s;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  {
    const T *Dp = L + S.poff[k] + (size_t)c0 * (ld + 1);
    const int i = tid & (TM - 1);
    if (i < mrows)
      for (int kk = tid / TM; kk < nb; kk += 128 / TM) {
        T v = Xs[(size_t)kk * LDX + i];
        if (UNIT_SYM) {
          W[S.poff[k] + (size_t)(c0 + kk) * ld + r_base + i] = v;   // L*D, the B operand of the updates
          v = v / Dp[(size_t)kk * (ld + 1)];
        }
        Xp[(size_t)(c0 + kk) * ld + r_base + i] = v;
      }
  }
}

// ---------------------------------------------------------------- diagonal sub-block
// Factor the nb x nb block at (c0,c0) of the cblk's diagonal blok with the reference's static-pivot
// rule (|pivot| < critere => pivot := critere, nbpivot++; compute_diag.c:133-137, 232-236, 444-448).
// One CTA of 256 threads holds the whole block in REGISTERS, 2-D cyclic over a 16 x 16 thread grid
// (thread (tx,ty) owns rows tx+16a, columns ty+16b), so the right-looking rank-1 updates are pure
// register FMAs; per pivot only the scaled column (and, for LU, the pivot row) passes through shared
// memory, double-buffered: one barrier per pivot.
template <class T> __device__ __forceinline__ bool below_crit(T d, double crit);
template <> __device__ __forceinline__ bool below_crit<double>(double d, double crit) { return fabs(d) < crit; }
template <> __device__ __forceinline__ bool below_crit<cdouble>(cdouble d, double crit) {
  return hypot(d.x, d.y) < crit;   // (squaring the threshold would underflow for tiny values)
}
template <> __device__ __forceinline__ bool below_crit<float>(float d, double crit) { return fabs((double)d) < crit; }
template <> __device__ __forceinline__ bool below_crit<cfloat>(cfloat d, double crit) {
  return hypot((double)d.x, (double)d.y) < crit;
}
__device__ __forceinline__ float shfl_t(unsigned m, float v, int src) { return __shfl_sync(m, v, src); }
__device__ __forceinline__ cfloat shfl_t(unsigned m, cfloat v, int src) {
  return cfloat(__shfl_sync(m, v.x, src), __shfl_sync(m, v.y, src));
}
__device__ __forceinline__ double shfl_t(unsigned m, double v, int src) { return __shfl_sync(m, v, src); }
__device__ __forceinline__ cdouble shfl_t(unsigned m, cdouble v, int src) {
  return cdouble(__shfl_sync(m, v.x, src), __shfl_sync(m, v.y, src));
}

// reciprocal of the pivot (and, for LLt, the pivot's square root) off the slow IEEE sqrt/div paths
template <int FACTO> __device__ __forceinline__ void pivot_inv(double &d, double &inv) {
  if (FACTO == F_LLT) { inv = rsqrt(d); d = d * inv; inv = inv + inv * fma(-d, inv, 1.0); }   // one Newton step on 1/sqrt
  else inv = __drcp_rn(d);
}
template <int FACTO> __device__ __forceinline__ void pivot_inv(float &d, float &inv) {
  if (FACTO == F_LLT) d = sqrtf(d);
  inv = 1.0f / d;
}
template <int FACTO> __device__ __forceinline__ void pivot_inv(cfloat &d, cfloat &inv) {
  if (FACTO == F_LLT) d = ST<cfloat>::sqrt(d);
  inv = cfloat(1.0f, 0.0f) / d;
}
template <int FACTO> __device__ __forceinline__ void pivot_inv(cdouble &d, cdouble &inv) {
  if (FACTO == F_LLT) d = ST<cdouble>::sqrt(d);
  inv = cdouble(1.0, 0.0) / d;
}

// pivot k (column owners: the half-warp with ty == k%16): static-pivot test, scale the column,
// publish it (and the row factor of the symmetric variants) in buffer k&1
template <class T, int FACTO, int KA, int R>
__device__ __forceinline__ void diag_pivot(T (&a)[R][R], int k, int kx, int tx, int lane, double crit,
                                           unsigned long long *nbpivot, T (*colbuf)[16 * R], T (*rowbuf)[16 * R]) {
  const int buf = k & 1;
  T d = a[KA][KA];
  if (tx == kx && below_crit<T>(d, crit)) { d = ST<T>::from_real(crit); atomicAdd(nbpivot, 1ULL); }
  d = shfl_t(0xFFFFu << (lane & 16), d, (lane & 16) | kx);
  T inv;
  pivot_inv<FACTO>(d, inv);
  if (tx == kx) a[KA][KA] = d;
#pragma unroll
  for (int ia = KA; ia < R; ++ia) {
    const int i = tx + 16 * ia;
    if (i > k) {
      const T l = a[ia][KA] * inv;
      a[ia][KA] = l;
      colbuf[buf][i] = l;
      if (FACTO == F_LLT) rowbuf[buf][i] = l;
      else if (FACTO == F_LDLT) rowbuf[buf][i] = d * l;
      else if (FACTO == F_LDLH) rowbuf[buf][i] = d * ST<T>::conj(l);
    }
  }
}

template <class T, int FACTO, int KA, int R>
__device__ __forceinline__ void diag_steps(T (&a)[R][R], int nb, int tx, int ty, int lane, double crit,
                                           unsigned long long *nbpivot, T (*colbuf)[16 * R], T (*rowbuf)[16 * R]) {
  if (KA * 16 >= nb) return;
  const T zero = ST<T>::zero();
  // first pivot of this 16-column group (no look-ahead across groups)
  if (ty == 0) diag_pivot<T, FACTO, KA, R>(a, KA * 16, 0, tx, lane, crit, nbpivot, colbuf, rowbuf);
  if (FACTO == F_LU && tx == 0) {
#pragma unroll
    for (int jb = KA; jb < R; ++jb) {
      const int j = ty + 16 * jb;
      if (j > KA * 16) rowbuf[(KA * 16) & 1][j] = a[KA][jb];
    }
  }
  for (int kx = 0; kx < 16; ++kx) {
    const int k = KA * 16 + kx;
    if (k >= nb) break;
    const int buf = k & 1;
    __syncthreads();
    T cv[R], rv[R];
#pragma unroll
    for (int q = KA; q < R; ++q) {
      const int i = tx + 16 * q, j = ty + 16 * q;
      cv[q] = (i > k) ? colbuf[buf][i] : zero;
      rv[q] = (j > k) ? rowbuf[buf][j] : zero;
    }
    // the block row / block column holding pivot k+1 first, so that its owners can run ahead
#pragma unroll
    for (int ia = KA; ia < R; ++ia) a[ia][KA] = a[ia][KA] - cv[ia] * rv[KA];
    if (FACTO == F_LU) {
#pragma unroll
      for (int jb = KA + 1; jb < R; ++jb) a[KA][jb] = a[KA][jb] - cv[KA] * rv[jb];
    }
    if (kx + 1 < 16 && k + 1 < nb) {
      if (ty == kx + 1) diag_pivot<T, FACTO, KA, R>(a, k + 1, kx + 1, tx, lane, crit, nbpivot, colbuf, rowbuf);
      if (FACTO == F_LU && tx == kx + 1) {
#pragma unroll
        for (int jb = KA; jb < R; ++jb) {
          const int j = ty + 16 * jb;
          if (j > k + 1) rowbuf[(k + 1) & 1][j] = a[KA][jb];
        }
      }
    }
#pragma unroll
    for (int ia = KA + (FACTO == F_LU ? 1 : 0); ia < R; ++ia)
#pragma unroll
      for (int jb = KA + 1; jb < R; ++jb) {
        if (FACTO != F_LU && jb > ia) continue;   // symmetric variants: lower triangle only
        a[ia][jb] = a[ia][jb] - cv[ia] * rv[jb];
      }
  }
  __syncthreads();   // the next group's first pivot reuses buffer parity 0
}

template <class T, int FACTO>
__global__ void __launch_bounds__(256)
k_diag_sub(DevSym S, T *L, T *U, const SubTask *__restrict__ tasks, double crit, unsigned long long *nbpivot) {
  constexpr int R = SubCfg<T>::NBMAX / 16;
  __shared__ T colbuf[2][16 * R];
  __shared__ T rowbuf[2][16 * R];
  const SubTask tk = tasks[blockIdx.x];
  const int c = tk.cblk, ld = S.stride[c], nb = tk.c1 - tk.c0;
  T *A = L + S.poff[c] + (size_t)tk.c0 * (ld + 1);
  const int tid = threadIdx.x, lane = tid & 31, tx = tid & 15, ty = tid >> 4;
  T a[R][R];
#pragma unroll
  for (int ia = 0; ia < R; ++ia)
#pragma unroll
    for (int jb = 0; jb < R; ++jb) {
      const int i = tx + 16 * ia, j = ty + 16 * jb;
      a[ia][jb] = (i < nb && j < nb) ? A[(size_t)j * ld + i] : ST<T>::zero();
    }
#define PB200_DS(KA) if constexpr (R > KA) diag_steps<T, FACTO, KA, R>(a, nb, tx, ty, lane, crit, nbpivot, colbuf, rowbuf);
  PB200_DS(0) PB200_DS(1) PB200_DS(2) PB200_DS(3) PB200_DS(4) PB200_DS(5) PB200_DS(6) PB200_DS(7)
#undef PB200_DS
#pragma unroll
  for (int ia = 0; ia < R; ++ia)
#pragma unroll
    for (int jb = 0; jb < R; ++jb) {
      const int i = tx + 16 * ia, j = ty + 16 * jb;
      if (i < nb && j < nb && (FACTO == F_LU || i >= j)) A[(size_t)j * ld + i] = a[ia][jb];
    }
  if (FACTO == F_LU) {
    // mirror (LU)^T into ucoeftab's diagonal blok through a 16 x 16 shared-memory transpose per register tile
    __shared__ T tr[16][17];
    T *UA = U + S.poff[c] + (size_t)tk.c0 * (ld + 1);
#pragma unroll
    for (int ia = 0; ia < R; ++ia)
#pragma unroll
      for (int jb = 0; jb < R; ++jb) {
        __syncthreads();
        tr[ty][tx] = a[ia][jb];                     // element (16ia+tx, 16jb+ty)
        __syncthreads();
        const int i = ty + 16 * ia, j = tx + 16 * jb;   // element (i,j) sits in tr[tx][ty]
        if (i < nb && j < nb) UA[(size_t)i * ld + j] = tr[tx][ty];
      }
  }
}

// ---------------------------------------------------------------- blocked diagonal sub-block (round 2)
// Same job and same register layout as k_diag_sub (16 x 16 thread grid, 2-D cyclic, the block never leaves the
// registers), but the pivots are taken BS at a time: per step the BS block columns (and, LU, block rows) go through
// shared memory once, EVERY thread factors the BS x BS diagonal block redundantly in registers (no barrier, no
// shuffle between pivots), one thread per row solves its row of the block column against it, and the trailing matrix
// receives one rank-BS update — 2 CTA barriers per BS pivots instead of one per pivot, which is what lets a whole
// 120-column cblk go through ONE diag -> TRSM round.  Blocking as PASTIX_potrf_block / sytrf_block / getrf_block
// (compute_diag.c:171-203, 262-307, 486-518); pivot rule of the unblocked kernels (:133-137, 232-236, 444-448).
template <class T, int FACTO> struct DiagBS { static constexpr int BS = (FACTO == F_LU || ST<T>::is_complex) ? 4 : 8; };

template <class T, int FACTO, int R, int S>
__device__ __forceinline__ void diag_blk_step(T (&a)[R][R], int nb, int tid, int tx, int ty, double crit,
                                              unsigned long long *nbpivot, T (*P)[16 * R], T (*Q)[16 * R],
                                              T (*X)[16 * R], T (*Y)[16 * R], T (*Dout)[DiagBS<T, FACTO>::BS]) {
  constexpr int BS = DiagBS<T, FACTO>::BS, NBP = 16 * R, J0 = S * BS, JB = J0 / 16, O = J0 % 16, J1 = J0 + BS, IA0 = J1 / 16;
  constexpr bool LU = (FACTO == F_LU), LDL = (FACTO == F_LDLT || FACTO == F_LDLH);
  if (J0 >= nb) return;   // uniform
  const T zero = ST<T>::zero();
  // 1. the owners publish the raw block columns (rows >= J0) and, LU, the raw block rows (columns >= J1)
  if (ty >= O && ty < O + BS) {
#pragma unroll
    for (int ia = JB; ia < R; ++ia) {
      const int r = tx + 16 * ia;
      if (r >= J0) P[ty - O][r] = a[ia][JB];
    }
  }
  if (LU && tx >= O && tx < O + BS) {
#pragma unroll
    for (int jb = IA0; jb < R; ++jb) {
      const int c = ty + 16 * jb;
      if (c >= J1) Q[tx - O][c] = a[JB][jb];
    }
  }
  __syncthreads();
  // 2. every thread factors the BS x BS diagonal block (d[r][q] = element (J0 + r, J0 + q))
  T d[BS][BS], invd[BS], piv[BS];
#pragma unroll
  for (int q = 0; q < BS; ++q)
#pragma unroll
    for (int r = 0; r < BS; ++r)
      if (LU || r >= q) d[r][q] = P[q][J0 + r];
#pragma unroll
  for (int k = 0; k < BS; ++k) {
    T pv = d[k][k];
    if (below_crit<T>(pv, crit)) {
      pv = ST<T>::from_real(crit);
      if (tid == 0 && J0 + k < nb) atomicAdd(nbpivot, 1ULL);
    }
    T inv;
    pivot_inv<FACTO>(pv, inv);
    d[k][k] = pv; invd[k] = inv; piv[k] = pv;
    T wk[BS];   // row factor of pivot k inside the block
#pragma unroll
    for (int r = k + 1; r < BS; ++r) {
      const T l = d[r][k] * inv;
      d[r][k] = l;
      if (FACTO == F_LLT) wk[r] = l;
      else if (FACTO == F_LDLT) wk[r] = pv * l;
      else if (FACTO == F_LDLH) wk[r] = pv * ST<T>::conj(l);
    }
#pragma unroll
    for (int c = k + 1; c < BS; ++c)
#pragma unroll
      for (int r = (LU ? k + 1 : c); r < BS; ++r) d[r][c] = d[r][c] - d[r][k] * (LU ? d[k][c] : wk[c]);
  }
  // 3. one thread publishes the factored block
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < BS; ++q)
#pragma unroll
      for (int r = 0; r < BS; ++r)
        if (LU || r >= q) Dout[q][r] = d[r][q];
  }
  // 4. one thread per row below the block (LU: and one per column right of it) solves against the block
  if (J1 < NBP) {
    constexpr int M = NBP - J1;                        // rows (columns) still to come, padding included
    constexpr int G = (2 * M <= 256) ? M : 128;        // LU: first thread of the column group
    if (tid < M) {
      const int r = J1 + tid;
      T x[BS];
#pragma unroll
      for (int q = 0; q < BS; ++q) {
        T acc = P[q][r];
#pragma unroll
        for (int qq = 0; qq < q; ++qq) {
          // LLt: x L11^T = p (symmetric, no conjugate: compute_diag.c:140); LDLt/LDLh: w conj?(L11)^T = p, w = l d;
          // LU: x U11 = p
          const T f = LU ? d[qq][q] : (FACTO == F_LDLH ? ST<T>::conj(d[q][qq]) : d[q][qq]);
          acc = acc - x[qq] * f;
        }
        if (LDL) {
          x[q] = acc;                                  // w = l * d
          const T l = acc * invd[q];
          X[q][r] = l;
          Y[q][r] = (FACTO == F_LDLH) ? piv[q] * ST<T>::conj(l) : piv[q] * l;
        } else {
          x[q] = acc * invd[q];
          X[q][r] = x[q];
        }
      }
    }
    if (LU && tid >= G && tid < G + M) {
      const int c = J1 + (tid - G);
      T u[BS];
#pragma unroll
      for (int q = 0; q < BS; ++q) {
        T acc = Q[q][c];
#pragma unroll
        for (int qq = 0; qq < q; ++qq) acc = acc - d[q][qq] * u[qq];   // L11 unit lower
        u[q] = acc;
        Y[q][c] = acc;
      }
    }
  }
  __syncthreads();
  // 5. the owners take the finished block columns (and rows) back
  if (ty >= O && ty < O + BS) {
#pragma unroll
    for (int ia = JB; ia < R; ++ia) {
      const int r = tx + 16 * ia;
      if (r >= J1) a[ia][JB] = X[ty - O][r];
      else if (r >= J0 && (LU || r - J0 >= ty - O)) a[ia][JB] = Dout[ty - O][r - J0];
    }
  }
  if (LU && tx >= O && tx < O + BS) {
#pragma unroll
    for (int jb = IA0; jb < R; ++jb) {
      const int c = ty + 16 * jb;
      if (c >= J1) a[JB][jb] = Y[tx - O][c];
    }
  }
  // 6. rank-BS update of the trailing matrix, in registers
  if (IA0 < R) {
#pragma unroll
    for (int q = 0; q < BS; ++q) {
      T xv[R], yv[R];
#pragma unroll
      for (int i = IA0; i < R; ++i) {
        const int r = tx + 16 * i, c = ty + 16 * i;
        xv[i] = (r >= J1) ? X[q][r] : zero;
        yv[i] = (c >= J1) ? ((LU || LDL) ? Y[q][c] : X[q][c]) : zero;
      }
#pragma unroll
      for (int ia = IA0; ia < R; ++ia)
#pragma unroll
        for (int jb = IA0; jb < R; ++jb) {
          if (!LU && jb > ia) continue;   // symmetric variants: lower triangle of 16 x 16 register tiles only
          a[ia][jb] = a[ia][jb] - xv[ia] * yv[jb];
        }
    }
  }
}

template <class T, int FACTO, int R>
__global__ void __launch_bounds__(256)
k_diag_blk(DevSym S, T *L, T *U, const SubTask *__restrict__ tasks, double crit, unsigned long long *nbpivot) {
  constexpr int BS = DiagBS<T, FACTO>::BS, NBP = 16 * R;
  constexpr bool LU = (FACTO == F_LU), LDL = (FACTO == F_LDLT || FACTO == F_LDLH);
  __shared__ T P[BS][NBP];
  __shared__ T Q[LU ? BS : 1][LU ? NBP : 1];
  __shared__ T X[BS][NBP];
  __shared__ T Y[(LU || LDL) ? BS : 1][(LU || LDL) ? NBP : 1];
  __shared__ T Dout[BS][BS];
  pdl_launch_dependents();
  const SubTask tk = tasks[blockIdx.x];
  pdl_wait();
  const int c = tk.cblk, ld = S.stride[c], nb = tk.c1 - tk.c0;
  T *A = L + S.poff[c] + (size_t)tk.c0 * (ld + 1);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  T a[R][R];
#pragma unroll
  for (int ia = 0; ia < R; ++ia)
#pragma unroll
    for (int jb = 0; jb < R; ++jb) {
      if (!LU && jb > ia) continue;
      const int i = tx + 16 * ia, j = ty + 16 * jb;
      // padding: identity, so that the blocked steps need no edge cases
      a[ia][jb] = (i < nb && j < nb) ? A[(size_t)j * ld + i] : ((i == j) ? ST<T>::from_real(1.0) : ST<T>::zero());
    }
  typedef T (*Row)[NBP];
#define PB200_DB(SS) if constexpr (SS * BS < NBP) diag_blk_step<T, FACTO, R, SS>(a, nb, tid, tx, ty, crit, nbpivot, P, (Row)Q, X, (Row)Y, Dout);
  PB200_DB(0) PB200_DB(1) PB200_DB(2) PB200_DB(3) PB200_DB(4) PB200_DB(5) PB200_DB(6) PB200_DB(7)
  PB200_DB(8) PB200_DB(9) PB200_DB(10) PB200_DB(11) PB200_DB(12) PB200_DB(13) PB200_DB(14) PB200_DB(15)
  PB200_DB(16) PB200_DB(17) PB200_DB(18) PB200_DB(19) PB200_DB(20) PB200_DB(21) PB200_DB(22) PB200_DB(23)
  PB200_DB(24) PB200_DB(25) PB200_DB(26) PB200_DB(27) PB200_DB(28) PB200_DB(29) PB200_DB(30) PB200_DB(31)
#undef PB200_DB
#pragma unroll
  for (int ia = 0; ia < R; ++ia)
#pragma unroll
    for (int jb = 0; jb < R; ++jb) {
      if (!LU && jb > ia) continue;
      const int i = tx + 16 * ia, j = ty + 16 * jb;
      if (i < nb && j < nb && (LU || i >= j)) A[(size_t)j * ld + i] = a[ia][jb];
    }
  if constexpr (LU) {
    // mirror (LU)^T into ucoeftab's diagonal blok through a 16 x 16 shared-memory transpose per register tile
    __shared__ T tr[16][17];
    T *UA = U + S.poff[c] + (size_t)tk.c0 * (ld + 1);
#pragma unroll
    for (int ia = 0; ia < R; ++ia)
#pragma unroll
      for (int jb = 0; jb < R; ++jb) {
        __syncthreads();
        tr[ty][tx] = a[ia][jb];                     // element (16ia+tx, 16jb+ty)
        __syncthreads();
        const int i = ty + 16 * ia, j = tx + 16 * jb;   // element (i,j) sits in tr[tx][ty]
        if (i < nb && j < nb) UA[(size_t)i * ld + j] = tr[tx][ty];
      }
  }
}

// ---- compact diagonal-block kernel (real types, LLt / LDLt, sub-blocks of at most 64 columns).
// Same reference (factor_diag: PASTIX_potrf_block / PASTIX_sytrf_block, compute_diag.c) and the same 8-pivot steps as
// k_diag_blk, but the block lives in SHARED memory and the steps are a real loop: k_diag_blk keeps the block in
// registers, which forces every step to be unrolled with its own static indices — 5 640 SASS instructions of
// straight-line code executed once per launch by one CTA, and ncu shows that CTA waiting for instructions
// (stall_no_inst) more than for pivots.  Here the whole kernel is a few hundred instructions that stay in the
// instruction cache across the eight steps.
//   per step: warp 0 factors the 8 x 8 diagonal block in the registers of every lane (no shuffle between pivots, static-
//   pivot rule per pivot in order), each lane solves its two rows of the block column against it and publishes them
//   (and l * d for LDLt); then all 256 threads apply the rank-8 update to the trailing lower triangle, 4 x 4 elements
//   per thread from register copies of their rows / columns of the panel.
template <class T, int FACTO>
__global__ void __launch_bounds__(256)
k_diag_cmp(DevSym S, T *L, const SubTask *__restrict__ tasks, double crit, unsigned long long *nbpivot) {
  constexpr int NB = 64, LD = 65, BS = 8;
  constexpr bool LDL = (FACTO == F_LDLT);
  __shared__ T A[NB * LD];                 // column-major, A[c * LD + r]
  __shared__ T Y[BS][NB];                  // LDLt: l * d of the panel columns (the second factor of the update)
  pdl_launch_dependents();
  const SubTask tk = tasks[blockIdx.x];
  pdl_wait();
  const int c = tk.cblk, ld = S.stride[c], nb = tk.c1 - tk.c0;
  T *Ag = L + S.poff[c] + (size_t)tk.c0 * (ld + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T one = ST<T>::from_real(1.0), zero = ST<T>::zero();
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e & (NB - 1), cc = e >> 6;
    // padding: identity, so that the blocked steps need no edge cases
    A[cc * LD + r] = (r < nb && cc < nb && r >= cc) ? Ag[(size_t)cc * ld + r] : (r == cc ? one : zero);
  }
  __syncthreads();
  const int nsteps = (nb + BS - 1) / BS;
  for (int s = 0; s < nsteps; ++s) {
    const int J0 = s * BS, J1 = J0 + BS;
    if (warp == 0) {
      // 1. the BS x BS diagonal block, factored by every lane in registers (d[r][q] = element (J0 + r, J0 + q), r >= q)
      T d[BS][BS], invd[BS], piv[BS];
#pragma unroll
      for (int q = 0; q < BS; ++q)
#pragma unroll
        for (int r = q; r < BS; ++r) d[r][q] = A[(J0 + q) * LD + J0 + r];
#pragma unroll
      for (int k = 0; k < BS; ++k) {
        T pv = d[k][k];
        if (below_crit<T>(pv, crit)) {
          pv = ST<T>::from_real(crit);
          if (lane == 0 && J0 + k < nb) atomicAdd(nbpivot, 1ULL);
        }
        T inv;
        pivot_inv<FACTO>(pv, inv);
        d[k][k] = pv; invd[k] = inv; piv[k] = pv;
        T wk[BS];
#pragma unroll
        for (int r = k + 1; r < BS; ++r) {
          const T l = d[r][k] * inv;
          d[r][k] = l;
          wk[r] = LDL ? pv * l : l;
        }
#pragma unroll
        for (int cc = k + 1; cc < BS; ++cc)
#pragma unroll
          for (int r = cc; r < BS; ++r) d[r][cc] = d[r][cc] - d[r][k] * wk[cc];
      }
      // 2. lanes 0..7 put row `lane` of the factored block back
      if (lane < BS) {
#pragma unroll
        for (int r = 0; r < BS; ++r)
#pragma unroll
          for (int q = 0; q <= r; ++q)
            if (r == lane) A[(J0 + q) * LD + J0 + r] = d[r][q];
      }
      // 3. rows below the block: x L11^T = p (LLt);  w L11^T = p, l = w / d (LDLt) — two rows per lane
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = J1 + lane + 32 * h;
        if (r < NB) {
          T x[BS];
#pragma unroll
          for (int q = 0; q < BS; ++q) {
            T acc = A[(J0 + q) * LD + r];
#pragma unroll
            for (int qq = 0; qq < q; ++qq) acc = acc - x[qq] * d[q][qq];
            if (LDL) {
              x[q] = acc;                              // w = l * d
              const T l = acc * invd[q];
              A[(J0 + q) * LD + r] = l;
              Y[q][r] = piv[q] * l;
            } else {
              x[q] = acc * invd[q];
              A[(J0 + q) * LD + r] = x[q];
            }
          }
        }
      }
    }
    __syncthreads();
    // 4. rank-BS update of the trailing lower triangle: thread (tx, ty) owns rows J1 + tx + 16 i, columns J1 + ty + 16 j
    if (J1 < NB) {
      const int tx = tid & 15, ty = tid >> 4;
      T xr[4][BS], yc[4][BS];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = J1 + tx + 16 * i, cc = J1 + ty + 16 * i;
#pragma unroll
        for (int q = 0; q < BS; ++q) {
          xr[i][q] = r < NB ? A[(J0 + q) * LD + r] : zero;
          yc[i][q] = cc < NB ? (LDL ? Y[q][cc] : A[(J0 + q) * LD + cc]) : zero;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = J1 + tx + 16 * i, cc = J1 + ty + 16 * j;
          if (r < NB && cc <= r) {
            T acc = zero;
#pragma unroll
            for (int q = 0; q < BS; ++q) acc += xr[i][q] * yc[j][q];
            A[cc * LD + r] -= acc;
          }
        }
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e & (NB - 1), cc = e >> 6;
    if (r < nb && cc < nb && r >= cc) Ag[(size_t)cc * ld + r] = A[cc * LD + r];
  }
}

// LU: ucoeftab's diagonal blok <- transpose of coeftab's, for every cblk of the level about to be
// factored (all external contributions have landed in coeftab by then).
template <class T>
__global__ void k_diag_transpose(DevSym S, const T *L, T *U, const int *__restrict__ cblks, int ncblk) {
  __shared__ T tile[32][33];
  for (int cc = blockIdx.y; cc < ncblk; cc += gridDim.y) {
  const int c = cblks[cc];
  const int w = S.width[c], ld = S.stride[c];
  const int nt = (w + 31) / 32;
  const T *A = L + S.poff[c];
  T *B = U + S.poff[c];
  for (int tt = blockIdx.x; tt < nt * nt; tt += gridDim.x) {
    const int bi = (tt / nt) * 32, bj = (tt % nt) * 32;
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
      const int i = bi + threadIdx.x, j = bj + y;
      if (i < w && j < w) tile[y][threadIdx.x] = A[(size_t)j * ld + i];
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
      const int i = bj + threadIdx.x, j = bi + y;   // B(i,j) = A(j,i)
      if (i < w && j < w) B[(size_t)j * ld + i] = tile[threadIdx.x][y];
    }
  }
  }
}

// LU, cblks factored in several sub-panel rounds: the rounds leave U12 only as its transpose in ucoeftab's
// diagonal blok (and L21 only in coeftab's).  Complete both diagonal bloks to the full LU / (LU)^T that
// PASTIX_getrf_block + DimTrans leave (compute_diag.c:486-536), so that coeftab/ucoeftab read back equal
// the reference's element for element.
template <class T>
__global__ void k_diag_complete_lu(DevSym S, T *L, T *U, const int *__restrict__ cblks, int ncblk, int nbmax) {
  __shared__ T tl[32][33], tu[32][33];
  for (int cc = blockIdx.y; cc < ncblk; cc += gridDim.y) {
    const int c = cblks[cc];
    const int w = S.width[c], ld = S.stride[c];
    if (w <= nbmax) continue;
    const int nt = (w + 31) / 32;
    T *A = L + S.poff[c];
    T *B = U + S.poff[c];
    for (int tt = blockIdx.x; tt < nt * nt; tt += gridDim.x) {
      const int bi = (tt / nt) * 32, bj = (tt % nt) * 32;   // target tile: rows bi.., cols bj.. with bi <= bj (upper part)
      if (bi > bj) continue;
      __syncthreads();
      for (int y = threadIdx.y; y < 32; y += blockDim.y) {   // source tile: rows bj.., cols bi.. (lower part)
        const int i = bj + threadIdx.x, j = bi + y;
        if (i < w && j < w) { tl[y][threadIdx.x] = A[(size_t)j * ld + i]; tu[y][threadIdx.x] = B[(size_t)j * ld + i]; }
      }
      __syncthreads();
      for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int i = bi + threadIdx.x, j = bj + y;           // element (i, j), strictly upper only
        if (i < w && j < w && i < j) { A[(size_t)j * ld + i] = tu[threadIdx.x][y]; B[(size_t)j * ld + i] = tl[threadIdx.x][y]; }
      }
    }
  }
}

// ---------------------------------------------------------------- static scatter maps
// one CTA per cblk, threads over its off-diagonal panel rows
// fan-out (multi-GPU): the updates of a SHARED source cblk are computed by the GPUs owning their targets; on this GPU the
// columns of such a source that face a cblk owned elsewhere get the "no column" mark (cb = INT_MAX fails every
// rb >= cb test of the scatter), so a tile only writes what this GPU owns
__global__ void k_build_maps(DevSym S, const BlokTgt *__restrict__ btgt, const int64_t *__restrict__ rmbase,
                             RowMap *rm, ColMap *cm, const int *__restrict__ owner, const char *__restrict__ fanout, int rank) {
  const int k = blockIdx.x;
  const int w = S.width[k], ld = S.stride[k], bf = S.fblok[k] + 1, be = S.fblok[k + 1];
  const int64_t base = rmbase[k];
  for (int m = w + threadIdx.x; m < ld; m += blockDim.x) {
    const int b = upper_le(S.coefind, bf, be, m);
    const int roff = m - S.coefind[b];
    const BlokTgt bt = btgt[b];
    RowMap r; r.rb = b - bf; r.roff = roff;
    ColMap c; c.ctgt = bt.tgt + (int64_t)roff * bt.tld; c.cb = b - bf; c.cj = bt.cj0 + roff; c.tw = bt.tw; c.tld = bt.tld;
    if (fanout != nullptr && fanout[k] && owner[bt.fc] != rank) c.cb = 0x7fffffff;
    c.pad0 = c.pad1 = 0;
    rm[base + m - w] = r;
    cm[base + m - w] = c;
  }
}

// one thread per tile of one k_gemm_scatter launch
template <int TM, int TN>
__global__ void k_build_tiledesc(DevSym S, DevMap M, const GemmTask *__restrict__ tasks, const int *__restrict__ tile2task,
                                 int ntiles, TileDesc *out) {
  const int tile = blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= ntiles) return;
  const GemmTask tk = tasks[tile2task[tile]];
  const int k = tk.cblk, tn = tile - tk.tile0;
  TileDesc d;
  d.poff = S.poff[k]; d.pbase = M.pairbase[k]; d.rmrow = M.rmbase[k] - S.width[k];
  d.ld = S.stride[k];
  d.m0 = tk.arow0; d.mrows = min(TM, tk.arow1 - tk.arow0);
  d.n0 = tk.brow0 + tn * TN; d.ncols = min(TN, tk.brow1 - d.n0);
  d.k0 = tk.k0; d.k1 = tk.k1; d.mode = tk.mode;
  d.rb_lo = d.cb_lo = 0; d.nrb = d.ncb = 1;
  if (tk.mode == 0) {
    const int bf = S.fblok[k] + 1, be = S.fblok[k + 1];
    d.rb_lo = upper_le(S.coefind, bf, be, d.m0) - bf;
    d.nrb = upper_le(S.coefind, bf, be, d.m0 + d.mrows - 1) - bf - d.rb_lo + 1;
    d.cb_lo = upper_le(S.coefind, bf, be, d.n0) - bf;
    d.ncb = upper_le(S.coefind, bf, be, d.n0 + d.ncols - 1) - bf - d.cb_lo + 1;
  }
  out[tile] = d;
}

// ---------------------------------------------------------------- pair table
// pairoff[pairbase[k] + tri(lb2, lb1)] for off-diagonal bloks b1 <= b2 of cblk k: offset, inside the
// panel of fcblk(b1), of the row that faces b2's first row — what add_contrib_local recomputes for
// every contribution (sopalin_compute.c:923-945).  -1: no facing blok.  `napa` is raised when a blok
// is only partially covered (incomplete factorization): those matrices take the generic path.
__global__ void k_build_pairs(DevSym S, const int64_t *__restrict__ pairbase, int *pairoff, int *napa) {
  const int k = blockIdx.x;
  const int b0 = S.fblok[k] + 1, nb = S.fblok[k + 1] - b0;
  const int64_t base = pairbase[k];
  for (int e = threadIdx.x; e < nb * (nb + 1) / 2; e += blockDim.x) {
    // e = tri(lb2, lb1)
    int lb2 = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
    while (lb2 * (lb2 + 1) / 2 > e) --lb2;
    while ((lb2 + 1) * (lb2 + 2) / 2 <= e) ++lb2;
    const int lb1 = e - lb2 * (lb2 + 1) / 2;
    const int b1 = b0 + lb1, b2 = b0 + lb2;
    const int fc = S.fcblk[b1];
    const int r = S.frow[b2], rl = r + S.nrow[b2] - 1;
    const int tb = upper_le(S.frow, S.fblok[fc], S.fblok[fc + 1], r);
    int ro = -1;
    if (tb >= S.fblok[fc] && r < S.frow[tb] + S.nrow[tb]) {
      ro = S.coefind[tb] + (r - S.frow[tb]);
      if (rl >= S.frow[tb] + S.nrow[tb]) *napa = 1;
    } else {
      *napa = 1;
    }
    pairoff[base + e] = ro;
  }
}

}  // namespace pb200
