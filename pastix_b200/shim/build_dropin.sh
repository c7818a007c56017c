#!/bin/bash
# Builds the drop-in host library pastix_b200/lib/libpastix_dropin_<p>.so (p in d z s c):
# the UNMODIFIED reference (PaStiX 5.2.2.16, compiled from its sources where they lie under
# $PASTIX_REFERENCE/src — nothing is copied into this repo) with ONE object replaced:
# sopalin3d.o (x4 factorization variants) -> pastix_b200/shim/sopalin_b200_shim.c, which routes
# API_TASK_NUMFACT / API_TASK_SOLVE to the CUDA layer libpastix_b200.so.
# one more object: raff_functions.o (x4) -> pastix_b200/shim/shim_raff.c (refinement vector back end on the device),
# and ONE symbol: CscOrdistrib (csc_intern_build.c:352) -> pastix_b200/shim/shim_csc.c (internal CSC built on the
# device); the reference's routine stays linked as CscOrdistrib_hostref (objcopy --redefine-sym).
# Recipe = SURVEY.md §8c: -DFORCE_NOMPI, no Scotch/METIS (API_ORDER_PERSONAL + KASS), 64-bit
# PASTIX_INT, -DMULT_SMX (multi-RHS), Fortran-ABI BLAS (only the reference's host-side refinement
# and analysis use it) = the OpenBLAS shipped inside the opencv_python_headless wheel.
# Without the reference tree (GPU box) the prebuilt .so files are kept.
# A precision followed by "32" (d32, z32, ...) builds with the reference's DEFAULT 32-bit PASTIX_INT (-DINTSIZE32,
# common_pastix.h:331-347) into libpastix_dropin_<p>_i32.so: the shim widens the SolverMatrix into the int64 C ABI and
# the internal CSC takes the reference's host CscOrdistrib (shim_csc.c), everything else is identical.
# usage: pastix_b200/shim/build_dropin.sh [precisions...]      (default: d z s c d32)
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
R=${PASTIX_REFERENCE:-/root/reference}/src
OUT="$ROOT/pastix_b200/lib"
[ -d "$R" ] || { echo "reference sources not found at $R — keeping prebuilt drop-in libraries"; exit 0; }
[ -f "$OUT/libpastix_b200.so" ] || { echo "build libpastix_b200.so first (python -m pastix_b200.build)"; exit 1; }
PRECS="${*:-d z s c d32}"
BLASDIR=$(python - <<'PY'
import glob, os, sysconfig
sp = sysconfig.get_paths()["purelib"]
c = glob.glob(os.path.join(sp, "opencv_python_headless.libs", "libopenblas*.so"))
print(os.path.dirname(c[0]) if c else "")
PY
)
[ -n "$BLASDIR" ] || { echo "no Fortran-ABI OpenBLAS found"; exit 1; }
BLASLIB=$(ls "$BLASDIR"/libopenblas*.so | head -1)
INC="-I$HERE -I$R/common/src -I$R/symbol/src -I$R/order/src -I$R/sopalin/src -I$R/blend/src -I$R/fax/src -I$R/kass/src -I$R/perf/src -I$R/sparse-matrix/src -I$ROOT/include"
CC="gcc -O2 -w -std=gnu99 -fcommon -fPIC"

for PP in $PRECS; do
  P=${PP%32}
  INTDEF="-DINTSIZE64"; SUF=""
  if [ "$PP" != "$P" ]; then INTDEF="-DINTSIZE32"; SUF="_i32"; fi
  case $P in
    d) TDEF="-DPREC_DOUBLE";;
    z) TDEF="-DPREC_DOUBLE -DTYPE_COMPLEX";;
    s) TDEF="";;
    c) TDEF="-DTYPE_COMPLEX";;
    *) echo "unknown precision $PP"; exit 1;;
  esac
  LIB="$OUT/libpastix_dropin_$P$SUF.so"
  if [ -f "$LIB" ] && [ "$LIB" -nt "$HERE/sopalin_b200_shim.c" ] && [ "$LIB" -nt "$HERE/shim_hooks.c" ] && [ "$LIB" -nt "$HERE/shim_csc.c" ] && [ "$LIB" -nt "$HERE/shim_raff.c" ] && [ "$LIB" -nt "$HERE/shim_table.h" ] && [ "$LIB" -nt "$0" ] && [ "$LIB" -nt "$ROOT/include/pastix_b200.h" ]; then
    echo "[$PP] up to date"; continue
  fi
  DEF="-DFORCE_NOMPI $TDEF $INTDEF -DMULT_SMX -DX_ARCHi686_pc_linux -DDOF_CONSTANT -DFORCE_NO_CUDA -DVERSION=\"pastix_b200\""
  OBJ="$OUT/_dropin_obj_$P$SUF"; mkdir -p "$OBJ"
  JOBS="$OBJ/jobs.txt"; : > "$JOBS"
  add() { echo "$CC $INC $DEF $3 -c $1 -o $OBJ/$2.o" >> "$JOBS"; }
  for f in common_integer common_error common_memory trace common; do add $R/common/src/$f.c c_$f -DCHOL_SOPALIN; done
  for f in dof dof_io symbol symbol_base symbol_check symbol_cost symbol_draw symbol_io symbol_keep symbol_levf symbol_nonzeros symbol_tree; do add $R/symbol/src/$f.c s_$f -DCHOL_SOPALIN; done
  for f in order order_base order_check order_io; do add $R/order/src/$f.c o_$f -DCHOL_SOPALIN; done
  for f in assemblyGener blend blend_symbol_cost blendctrl bulles cost costfunc distribPart elimin eliminfunc extendVector extrastruct fanboth2 param_blend partbuild queue simu smart_cblk_split solverMatrixGen solverRealloc solver_check solver_io splitfunc splitpart splitpartlocal symbolrand task write_ps blend_distributeOnGPU; do add $R/blend/src/$f.c b_$f -DCHOL_SOPALIN; done
  for f in symbol_compact symbol_costi symbol_fax_graph symbol_fax symbol_faxi_graph symbol_faxi; do add $R/fax/src/$f.c f_$f -DCHOL_SOPALIN; done
  for f in kass compact_graph amalgamate ifax sparRow SF_Direct SF_level find_supernodes KSupernodes sort_row; do add $R/kass/src/$f.c k_$f -DCHOL_SOPALIN; done
  add $R/sparse-matrix/src/pastix_sparse_matrix.c sm_psm -DCHOL_SOPALIN
  for f in bordi sopalin_thread compute_context_nbr coefinit csc_intern_build csc_intern_io csc_intern_solve csc_intern_updown csc_utils cscd_utils cscd_utils_fortran debug_dump ooc pastix pastix_fortran sopalin_init sopalin_option sparse_gemm_cpu tools; do add $R/sopalin/src/$f.c p_$f -DCHOL_SOPALIN; done
  # the reference compiles these four times (src/CMakeLists.txt:40-62); sopalin3d.c is the one we replace
  # raff_functions.c (the host `struct solver` back end of the refinement drivers) is replaced by shim_raff.c
  for f in starpu_submit_tasks csc_intern_compute starpu_updo; do
    add $R/sopalin/src/$f.c p_${f}_po -DCHOL_SOPALIN
    add $R/sopalin/src/$f.c p_${f}_ge -DSOPALIN_LU
    add $R/sopalin/src/$f.c p_${f}_sy -DNOEXTRADEF_SY
    add $R/sopalin/src/$f.c p_${f}_he -DHERMITIAN
  done
  add "$HERE/sopalin_b200_shim.c" x_shim_po -DCHOL_SOPALIN
  add "$HERE/sopalin_b200_shim.c" x_shim_ge -DSOPALIN_LU
  add "$HERE/sopalin_b200_shim.c" x_shim_sy -DNOEXTRADEF_SY
  add "$HERE/sopalin_b200_shim.c" x_shim_he -DHERMITIAN
  add "$HERE/shim_raff.c" x_raff_po -DCHOL_SOPALIN
  add "$HERE/shim_raff.c" x_raff_ge -DSOPALIN_LU
  add "$HERE/shim_raff.c" x_raff_sy -DNOEXTRADEF_SY
  add "$HERE/shim_raff.c" x_raff_he -DHERMITIAN
  add "$HERE/shim_hooks.c" x_shim_hooks -DCHOL_SOPALIN
  add "$HERE/shim_csc.c" x_shim_csc -DCHOL_SOPALIN
  xargs -P "$(nproc)" -I{} sh -c '{} 2>>'"$OBJ"'/err.log || echo "FAIL: {}"' < "$JOBS" | tee "$OBJ/fail.log"
  if [ -s "$OBJ/fail.log" ]; then echo "[$P] compile failures (see $OBJ/err.log)"; tail -20 "$OBJ/err.log"; exit 1; fi
  # the reference's CscOrdistrib stays linked under another name (multi-dof matrices, PB200_HOST_CSC=1); ours takes its place
  SYM=$(nm "$OBJ/p_csc_intern_build.o" | awk '$2=="T" && $3 ~ /CscOrdistrib$/ {print $3}' | head -1)
  [ -n "$SYM" ] || { echo "[$P] CscOrdistrib not found in the reference object"; exit 1; }
  objcopy --redefine-sym "$SYM=CscOrdistrib_hostref" "$OBJ/p_csc_intern_build.o" || exit 1
  if [ "$SYM" != "CscOrdistrib" ]; then objcopy --redefine-sym "CscOrdistrib=$SYM" "$OBJ/x_shim_csc.o" || exit 1; fi
  # the reference's release points (coefinit.c:479 CoefMatrix_Free, solverRealloc.c:217 solverExit) stay linked as
  # *_hostref; shim_hooks.c defines the public names, drops the device state of that SolverMatrix and calls them
  # Csc2updown (csc_intern_updown.c:339) likewise: shim_hooks.c brings the host copy of the internal CSC up to date first
  for FN in CoefMatrix_Free:p_coefinit solverExit:b_solverRealloc Csc2updown:p_csc_intern_updown; do
    F=${FN%%:*}; O=${FN##*:}
    SYM=$(nm "$OBJ/$O.o" | awk -v f="$F" '$2=="T" && $3 ~ (f "$") {print $3}' | head -1)
    [ -n "$SYM" ] || { echo "[$P] $F not found in the reference object $O.o"; exit 1; }
    objcopy --redefine-sym "$SYM=${F}_hostref" "$OBJ/$O.o" || exit 1
    if [ "$SYM" != "$F" ]; then objcopy --redefine-sym "$F=$SYM" "$OBJ/x_shim_hooks.o" || exit 1; fi
  done
  # the host SpMV routines that read the HOST CscMatrix (csc_intern_compute.c: CscbMAx, CscAxPb — static-pivot refinement,
  # host statistics) stay linked as <variant>_*_hostref; sopalin_b200_shim.c defines the public names (host copy of the
  # internal CSC on demand)
  for V in po ge sy he; do
    for F in CscbMAx CscAxPb; do
      SYM=$(nm "$OBJ/p_csc_intern_compute_$V.o" | awk -v f="_$F" '$2=="T" && $3 ~ (f "$") {print $3}' | head -1)
      [ -n "$SYM" ] || { echo "[$P] $F not found in p_csc_intern_compute_$V.o"; exit 1; }
      objcopy --redefine-sym "$SYM=${SYM}_hostref" "$OBJ/p_csc_intern_compute_$V.o" || exit 1
    done
  done
  gcc -shared -o "$LIB" "$OBJ"/*.o -L"$OUT" -lpastix_b200 "$BLASLIB" -lpthread -lm \
      -Wl,--disable-new-dtags -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$BLASDIR" -Wl,-rpath-link,"$BLASDIR" -Wl,--no-undefined 2> "$OBJ/link.log" \
      || { echo "[$P] link failed"; head -30 "$OBJ/link.log"; exit 1; }
  rm -rf "$OBJ"
  echo "[$PP] built $LIB"
done
