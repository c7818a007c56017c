/*
 * pastix_b200.h — C ABI of the B200-native sopalin numeric phase.
 *
 * This is the drop-in boundary for ONE path of PaStiX 5.2.2.16: the numeric
 * factorization (API_TASK_NUMFACT) and the up_down triangular solve
 * (API_TASK_SOLVE).  Everything above it (ordering, symbolic factorization,
 * blend analysis, the pastix()/pastix_fortran() task driver) stays the
 * reference's unchanged host code; this library consumes the SolverMatrix
 * that analysis produced.  Plain pointers and sizes only; all index arrays
 * are int64_t (the reference built with -DINTSIZE64; a 32-bit build widens
 * them in the shim).
 *
 * Reference interfaces replaced (paths relative to /root/reference/src):
 *   pb200_create        <- sopalin_init / sopalin_init_smp + CoefMatrix_Allocate
 *                          (sopalin/src/sopalin_init.c:99,865; coefinit.c:104)
 *   pb200_csc_build     <- CscOrdistrib (sopalin/src/csc_intern_build.c:352-570)
 *   pb200_assemble      <- CoefMatrix_Init + Csc2solv_cblk
 *                          (sopalin/src/coefinit.c:237; csc_intern_solve.c:65-125)
 *   pb200_norm1         <- CscNorm1 (sopalin/src/csc_intern_compute.c:120-176),
 *                          feeding the threshold of init_struct_sopalin
 *                          (sopalin/src/sopalin3d.c:586-606)
 *   pb200_factorize     <- {po,sy,he,ge}_sopalin_thread -> sopalin_smp task loop
 *                          (sopalin/src/sopalin3d.c:1388, 666-1262): compute_1d =
 *                          factor_diag + factor_trsm1d + compute_1dgemm/add_contrib_local
 *                          (compute_diag.c:538, compute_trsm.c:128, sopalin_compute.c:747-1032)
 *   pb200_inertia       <- inertia count (sopalin/src/sopalin3d.c:1145-1161)
 *   pb200_solve         <- {po,sy,he,ge}_updo_thread -> up_down_smp
 *                          (sopalin/src/updo.c:67,114-1664; updo_sendrecv.c:496-639)
 *   pb200_get_coeftab / pb200_set_coeftab / pb200_get_cblk
 *                       <- SolverCblk.coeftab / .ucoeftab contents
 *                          (blend/src/solver.h:94-117), so host consumers
 *                          (pastix_getSchur, dumps) keep working
 *   pb200_create_opts   <- the same with SopalinParam.schur (IPARM_SCHUR: the last
 *                          column block is left unfactored, sopalin_compute.c:767-772,
 *                          and ignored by up_down, updo.c:425-428)
 *   pb200_destroy       <- CoefMatrix_Free + sopalin_clean (coefinit.c:479)
 *
 * Error behaviour: every call returns PB200_SUCCESS (0) or a negative code;
 * pb200_last_error() gives the message.  There is NO CPU fallback: without a
 * CUDA device (or with the wrong architecture) pb200_create fails with
 * PB200_ERR_CUDA.
 */
#ifndef PASTIX_B200_H
#define PASTIX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* arithmetic type — values of API_FLOAT (common/src/api.h: API_REALSINGLE..API_COMPLEXDOUBLE) */
#define PB200_REALSINGLE    0
#define PB200_REALDOUBLE    1
#define PB200_COMPLEXSINGLE 2
#define PB200_COMPLEXDOUBLE 3

/* factorization — values of API_FACT (common/src/api.h:381-384) */
#define PB200_FACT_LLT  0
#define PB200_FACT_LDLT 1
#define PB200_FACT_LU   2
#define PB200_FACT_LDLH 3

#define PB200_SUCCESS        0
#define PB200_ERR_BADARG    -1
#define PB200_ERR_CUDA      -2
#define PB200_ERR_NOMEM     -3
#define PB200_ERR_STATE     -4
#define PB200_ERR_STRUCT    -5   /* SolverMatrix violates an invariant we rely on */

/*
 * Flat, read-only view of the reference's SolverMatrix for one process
 * (blend/src/solver.h:94-168).  Field names follow SolverCblk / SolverBlok.
 * cblk arrays hold cblknbr entries, except bloknum which holds cblknbr+1
 * (the sentinel = bloknbr).  Indices are 0-based (baseval 0), rows/columns
 * are numbered in the permuted ordering.
 */
typedef struct pb200_solver_s {
  int64_t        cblknbr;
  int64_t        bloknbr;
  const int64_t *fcolnum;   /* SolverCblk.fcolnum */
  const int64_t *lcolnum;   /* SolverCblk.lcolnum (inclusive) */
  const int64_t *bloknum;   /* SolverCblk.bloknum: first blok (the diagonal one) */
  const int64_t *stride;    /* SolverCblk.stride: leading dimension of the panel */
  const int64_t *frownum;   /* SolverBlok.frownum */
  const int64_t *lrownum;   /* SolverBlok.lrownum (inclusive) */
  const int64_t *cblknum;   /* SolverBlok.cblknum: facing column block */
  const int64_t *coefind;   /* SolverBlok.coefind: row offset inside the panel */
} pb200_solver_t;

typedef struct pb200_handle_s pb200_handle_t;

/* summary filled by pb200_info */
typedef struct pb200_info_s {
  int64_t n;            /* number of unknowns */
  int64_t coefnbr;      /* elements in one factor slab (sum stride*width) */
  int64_t nlevels;      /* elimination-tree levels in the launch schedule */
  int64_t device_bytes; /* bytes of HBM held by the handle */
  int32_t device;       /* CUDA device ordinal */
  int32_t sm_count;
  int32_t cc_major, cc_minor;
} pb200_info_t;

const char *pb200_last_error(void);
const char *pb200_version(void);

/* Build the device-side structures, the level schedule and allocate the
 * factor slab(s) in HBM. `device` < 0 selects the current CUDA device. */
int pb200_create(pb200_handle_t **h, const pb200_solver_t *solver,
                 int flttype, int factotype, int device);
int pb200_destroy(pb200_handle_t *h);

/* Options of pb200_create_opts (zero-initialise; NULL = all defaults).
 *   schur : IPARM_SCHUR == API_YES (api.h).  The reference never factors the column block that holds the last
 *           column (compute_1d returns at once, sopalin/src/sopalin_compute.c:767-772; blend keeps it unsplit,
 *           blend/src/splitpart.c:573-584): it only receives contributions, so after pb200_factorize its panel IS the
 *           Schur complement — what pastix_getSchur copies out of SOLV_COEFTAB (sopalin/src/pastix.c:6434-6475;
 *           read it with pb200_get_cblk(h, cblknbr-1, ...)).  up_down ignores that cblk and every blok facing it in
 *           the down, diagonal and up steps (updo.c:425-428, 639-646, 951-954, 1154-1180; updo_sendrecv.c:518-523):
 *           the interior system is solved and the Schur unknowns keep their right-hand side. */
typedef struct pb200_options_s {
  int32_t schur;
  int32_t reserved0;
  /* owner : optional [cblknbr] column block -> rank map for nranks > 1 (NULL: the mapping computed inside).  The
   *         drop-in passes the reference's OWN proportional mapping when asked (PB200_DIST_MAP=blend): blend maps the
   *         COMP_1D task of every column block to one of its IPARM_THREAD_NBR "local threads" with the same
   *         propMappTree / distribPart machinery that maps them to processes (blend/src/splitpart.c:752-1012,
   *         distribPart.c; result in SolverMatrix.ttsktab, blend/src/solver.h:158-159), and thread t becomes rank
   *         t * nranks / thrdnbr.  A column block whose subtree spans several ranks is treated as a shared
   *         top-separator block (fan-out), everything else as private (fan-in). */
  const int32_t *owner;
  int32_t reserved[4];
} pb200_options_t;
/* pb200_create / pb200_create_dist with options (rank 0 of 1 for a single GPU). */
int pb200_create_opts(pb200_handle_t **h, const pb200_solver_t *solver, int flttype, int factotype,
                      int device, int rank, int nranks, const pb200_options_t *opts);

/* ---- multi-GPU: one process per GPU of one box, `nranks` <= 8 (the reference's distributed mode:
 * dpastix + MPI, one SolverMatrix per process with fan-in targets, blend/src/ftgt.h:68-127).
 * Every process builds the SAME single-process SolverMatrix (the analysis is deterministic) and passes
 * its rank; column blocks are mapped to GPUs by proportional subtree mapping (blend/src/splitpart.c:752-1012)
 * computed inside; contributions to column blocks owned by another GPU are summed in a local fan-in
 * region and pulled by the owner over NVLink peer memory (add_contrib_target sopalin_compute.c:600-733,
 * recv_handle_fanin sopalin_sendrecv.c:182).  After pb200_create_dist the caller exchanges the opaque
 * IPC blobs (pb200_ipc_size() bytes per rank, e.g. MPI_Allgather / torch.distributed.all_gather) and
 * hands all of them, ordered by rank, to pb200_ipc_attach.  pb200_(re)assemble, pb200_factorize and
 * pb200_destroy are then collective calls.  nbpivot is this GPU's share (sum over ranks =
 * IPARM_STATIC_PIVOTING, the reference's MPI_Allreduce sopalin3d.c:1138).  The first pb200_solve /
 * pb200_get_coeftab / pb200_inertia after a factorization copies the other GPUs' factored panels into
 * the local slab, after which every GPU can solve its own share of the right-hand sides. */
int pb200_create_dist(pb200_handle_t **h, const pb200_solver_t *solver, int flttype, int factotype,
                      int device, int rank, int nranks);
int pb200_ipc_size(void);
int pb200_ipc_export(pb200_handle_t *h, void *blob);
int pb200_ipc_attach(pb200_handle_t *h, const void *all_blobs);
int pb200_dist_barrier(pb200_handle_t *h);
/* ONE host process driving the n GPUs of the box — what pastix() does with iparm[IPARM_CUDA_NBR] = n (api.h:115-120;
 * the reference's own multi-GPU knob, read by its StarPU back end).  hs[r] = pb200_create_dist(..., device r, rank r,
 * n); peer access is enabled pairwise and the peers' slabs are addressed directly, no IPC blob.  The collective calls
 * (pb200_assemble*, pb200_reassemble, pb200_factorize) block until every rank has reached them: drive each handle
 * from its own host thread (shim/sopalin_b200_shim.c does).  pb200_destroy_group frees all of them at once. */
int pb200_attach_local(pb200_handle_t **hs, int n);
int pb200_destroy_group(pb200_handle_t **hs, int n);
/* The mapping alone (host only, no GPU): owner[cblknbr] = rank of each column block; optional
 * contrib[cblknbr] = bit mask of the other ranks that contribute to it, load[nranks] = flops mapped. */
int pb200_dist_plan(const pb200_solver_t *solver, int factotype, int nranks,
                    int32_t *owner, uint32_t *contrib, double *load);
int pb200_info(const pb200_handle_t *h, pb200_info_t *info);

/* Panel offsets inside the flat slab: offsets[c] = sum_{k<c} stride_k*width_k,
 * cblknbr+1 entries. The slab passed to get/set_coeftab uses this layout. */
int pb200_panel_offsets(const pb200_handle_t *h, int64_t *offsets);

/* max_j sum_i |a_ij| over the internal CSC (host arrays). */
double pb200_norm1(int flttype, int64_t n, const int64_t *colptr, const void *values);

/* Zero the slab(s) and scatter the permuted CSC (0-based, colptr[n+1]) into the
 * panels. `tvalues` (same pattern, values of A^T) is required for LU and
 * ignored otherwise. The CSC is copied to HBM and kept for pb200_reassemble. */
int pb200_assemble(pb200_handle_t *h, const int64_t *colptr, const int64_t *rows,
                   const void *values, const void *tvalues);
/* Re-run the device-side zero + scatter from the CSC already resident in HBM. */
int pb200_reassemble(pb200_handle_t *h);

/* ---- vector back end of the refinement drivers (replaces the host `struct solver` operations of
 * sopalin/src/raff_functions.c:100-650 under the reference's unchanged raff_gmres.c / raff_grad.c / raff_bicgstab.c).
 * Vectors are n-element arrays of the handle's precision allocated by pb200_vec_alloc (managed memory: kernels use them
 * in HBM, the scalars and pointer tables the drivers allocate through the same call stay host-dereferenceable).
 * Scalars (`alpha`, `result`) are HOST pointers to one element.  Every call is synchronous. */
int pb200_vec_alloc(pb200_handle_t *h, void **p, int64_t bytes);                 /* Pastix_Malloc  */
int pb200_vec_free(pb200_handle_t *h, void *p);                                  /* Pastix_Free    */
int pb200_vec_set(pb200_handle_t *h, void *dst, const void *src_host, int64_t nelem);   /* host -> vector (Pastix_B, Pastix_X) */
int pb200_vec_get(pb200_handle_t *h, void *dst_host, const void *src, int64_t nelem);   /* vector -> host (Pastix_End)         */
int pb200_vec_zero(pb200_handle_t *h, void *dst, int64_t nelem);
int pb200_vec_copy(pb200_handle_t *h, void *dst, const void *src, int64_t nelem);       /* CscCopy  */
int pb200_vec_scal(pb200_handle_t *h, const void *alpha, void *x, int64_t nelem);       /* CscScal: x <- alpha x */
int pb200_vec_axpy(pb200_handle_t *h, const void *alpha, const void *x, void *y, int64_t nelem);   /* CscAXPY: y <- y + alpha x */
/* result = sum_i x_i * (conj_y ? conj(y_i) : y_i)   (CscGradBeta / CscGmresBeta / CscNormFro^2, csc_intern_compute.c:1041-1560);
 * fixed reduction tree: reproducible from run to run */
int pb200_vec_dot(pb200_handle_t *h, int conj_y, const void *x, const void *y, int64_t nelem, void *result);
/* r = A x (b == NULL) or r = b - A x, A = the internal CSC resident in HBM (CscAx / CscbMAx, csc_intern_compute.c:448, 1146);
 * type = CscMatrix.type ('S', 'H', 'U'); trans != 0: A^T x (IPARM_TRANSPOSE_SOLVE).  Gathered row by row: no atomics, the
 * same additions in the same order as the reference's sequential product. */
int pb200_csc_ax(pb200_handle_t *h, char type, int trans, const void *b, const void *x, void *r);
/* d = up_down(s) with the factors in HBM (Pastix_Precond); d may alias s */
int pb200_precond(pb200_handle_t *h, const void *s, void *d);

/* ---- internal CSC built on the device (replaces CscOrdistrib, sopalin/src/csc_intern_build.c:352-570).
 * From the user's CSC exactly as pastix() receives it — Fortran numbering, colptr[n+1], rows[nnz], values[nnz],
 * lower triangle for type 'S'/'H' — and Order.permtab (0-based new index of every unknown) it builds the matrix
 * in the new numbering, symmetrised for 'S'/'H' (mirror entries conjugated for 'H'), every column sorted by row:
 * the reference's CscMatrix flattened (CSC_COL / CSC_ROWTAB / CSC_VALTAB, blend/src/csc.h).
 *   type  : Type[1] of the reference call: 'S', 'H' or 'U'
 *   trans : 0 none; 1 values of A^T on the pattern of A ('U' only: sopar->transcsc); 2 alias of the values
 *           (forcetrans: LU on a symmetric matrix)
 * The result stays in HBM (pb200_assemble_csc consumes it without another upload); pb200_csc_fetch copies it to
 * host arrays of *nnz_out entries (0-based colptr[n+1]; tvalues may be NULL) for the reference's host-side users
 * of the internal CSC (CscNorm1, the refinement SpMV).  One dof per node only. */
typedef struct pb200_csc_s pb200_csc_t;
int pb200_csc_create(pb200_csc_t **out, int flttype, int device);
int pb200_csc_destroy(pb200_csc_t *c);
int pb200_csc_build(pb200_csc_t *c, char type, int64_t n, const int64_t *colptr, const int64_t *rows,
                    const void *values, const int64_t *permtab, int trans, int64_t *nnz_out);
int pb200_csc_fetch(pb200_csc_t *c, int64_t *colptr, int64_t *rows, void *values, void *tvalues);
/* the column pointers alone (what CSC_COLTAB of every column block is filled from, csc_intern_build.c:520-532): the
 * drop-in copies rows / values back only when a host-side reader of the CscMatrix shows up. */
int pb200_csc_fetch_colptr(pb200_csc_t *c, int64_t *colptr);
/* CscNorm1 (sopalin/src/csc_intern_compute.c:120-176) of the CSC in HBM: max_j sum_i |a_ij|, each column summed in
 * storage order like the reference's loop (identical result for real types; complex |.| is the device hypot). */
int pb200_csc_norm1(pb200_csc_t *c, double *norm);
/* pb200_assemble from the CSC that pb200_csc_build left in HBM (same device, same order and precision). */
int pb200_assemble_csc(pb200_handle_t *h, const pb200_csc_t *c);

/* Numeric factorization of the assembled panels.
 *   critere  : static-pivot threshold (|pivot| < critere => pivot := critere)
 *   nbpivot  : OUT number of replaced pivots (-> IPARM_STATIC_PIVOTING)
 *   seconds  : OUT device time of the factorization (-> DPARM_FACT_TIME) */
int pb200_factorize(pb200_handle_t *h, double critere, int64_t *nbpivot, double *seconds);

/* Number of positive diagonal terms of D (real LDLt; -> IPARM_INERTIA). */
int pb200_inertia(pb200_handle_t *h, int64_t *inertia);

/* up_down on `nrhs` right-hand sides held in HOST memory, permuted ordering,
 * column-major with leading dimension ldx (the layout of UpDownVector.sm2xtab,
 * blend/src/updown.h:69-72). Overwritten by the solution.
 *   seconds : OUT device time of the sweeps, copies excluded (-> DPARM_SOLV_TIME) */
int pb200_solve(pb200_handle_t *h, void *x, int64_t ldx, int64_t nrhs, double *seconds);
/* IPARM_TRANSPOSE_SOLVE (api.h): the following pb200_solve / pb200_solve_device calls solve A^T x = b.  Only an LU
 * factorization is affected (A^T = U^T L^T: the sweeps swap the L and U^T panels, updo.c:165-260, 1553-1600); the
 * symmetric factorizations ignore it, like the reference. */
int pb200_set_transpose_solve(pb200_handle_t *h, int transposed);
/* Internal CSC of type 'H' (IPARM_SYM = API_SYM_HER): the next pb200_assemble fills ucoeftab with the conjugate of
 * `tvalues`, as Csc2solv_cblk does for cscmtx->type == 'H' (csc_intern_solve.c:110-116).  pb200_assemble_csc takes the
 * type from the device CSC itself. */
int pb200_set_hermitian(pb200_handle_t *h, int hermitian);
/* Same with x already resident in HBM (device pointer). */
int pb200_solve_device(pb200_handle_t *h, void *x_dev, int64_t ldx, int64_t nrhs, double *seconds);

/* Copy the factor slab(s) device -> host / host -> device. U may be NULL unless LU. */
int pb200_get_coeftab(pb200_handle_t *h, void *L, void *U);
int pb200_set_coeftab(pb200_handle_t *h, const void *L, const void *U);
/* One column block's panel(s), device -> host: stride*width elements laid out like SolverCblk.coeftab / .ucoeftab
 * (blend/src/solver.h:94-117).  U may be NULL. */
int pb200_get_cblk(pb200_handle_t *h, int64_t cblk, void *L, void *U);

/* Declare panels uploaded with pb200_set_coeftab to be already factorized
 * (solve-only use: factors computed elsewhere, e.g. by the reference). */
int pb200_mark_factorized(pb200_handle_t *h);

/* Kernel launches issued by the last factorize / solve call (bench accounting). */
int64_t pb200_last_launches(const pb200_handle_t *h);

/* Measurement support (bench.py roofline): with profiling on, pb200_factorize serialises its
 * launches and times every kernel kind with CUDA events on the launching stream.
 *   kind_ms / kind_launches [4]: 0 diagonal blocks, 1 panel TRSM, 2 fused GEMM+scatter (updates into
 *   facing cblks), 3 in-panel trailing updates / transposes.
 *   gemm_flops: algorithmic flops of kind 2 for one factorization, PaStiX's own GEMM count
 *   (blend/src/blend_symbol_cost.c:382-430). */
int pb200_set_profile(pb200_handle_t *h, int on);
int pb200_get_profile(const pb200_handle_t *h, double *kind_ms, int64_t *kind_launches, double *gemm_flops);

/* FP64 dense-GEMM probe used by bench.py to establish the measured FP64 roof:
 * runs an m x n x k DGEMM tile loop with this library's own MMA micro-kernel
 * and returns achieved GFLOP/s (device timed). */
double pb200_probe_fp64_gflops(int device, int variant);

#ifdef __cplusplus
}
#endif
#endif /* PASTIX_B200_H */
