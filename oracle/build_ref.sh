#!/bin/bash
# TEST INFRASTRUCTURE — not part of the product path.
#
# Builds the UNMODIFIED reference (PaStiX 5.2.2.16 under /root/reference) into
# oracle/_ref/libpastix_ref_<p>.so for p in {d,z,s,c}, compiling the sources
# where they lie (nothing is copied into the repo).  The recipe is the one
# SURVEY.md §8c proved: -DFORCE_NOMPI, no Scotch/METIS (API_ORDER_PERSONAL +
# KASS), 64-bit PASTIX_INT, the reference's own 4-fold compile of the
# factorization sources (src/CMakeLists.txt:40-62), Fortran-ABI BLAS = the
# OpenBLAS 0.3.15 shipped inside the opencv_python_headless wheel.
# oracle/ref_driver.c (ours) is linked in to expose the internal structures.
#
# usage: oracle/build_ref.sh [precisions...]      (default: d z s c)
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
R=${PASTIX_REFERENCE:-/root/reference}/src
OUT="$HERE/_ref"
[ -d "$R" ] || { echo "reference sources not found at $R — keeping prebuilt $OUT"; exit 0; }
PRECS="${*:-d z s c}"
BLASDIR=$(python - <<'EOF'
import glob, sys, os, sysconfig
sp = sysconfig.get_paths()["purelib"]
c = glob.glob(os.path.join(sp, "opencv_python_headless.libs", "libopenblas*.so"))
print(os.path.dirname(c[0]) if c else "")
EOF
)
[ -n "$BLASDIR" ] || { echo "no Fortran-ABI OpenBLAS found"; exit 1; }
BLASLIB=$(ls "$BLASDIR"/libopenblas*.so | head -1)
mkdir -p "$OUT"
INC="-I$R/common/src -I$R/symbol/src -I$R/order/src -I$R/sopalin/src -I$R/blend/src -I$R/fax/src -I$R/kass/src -I$R/perf/src -I$R/sparse-matrix/src"
CC="gcc -O2 -w -std=gnu99 -fcommon -fPIC"

for P in $PRECS; do
  case $P in
    d) TDEF="-DPREC_DOUBLE";;
    z) TDEF="-DPREC_DOUBLE -DTYPE_COMPLEX";;
    s) TDEF="";;
    c) TDEF="-DTYPE_COMPLEX";;
  esac
  # MULT_SMX: multi-RHS up_down (sopalin_define.h:159); FORCE_NOMPI: nompi.h shim
  DEF="-DFORCE_NOMPI $TDEF -DINTSIZE64 -DMULT_SMX -DX_ARCHi686_pc_linux -DDOF_CONSTANT -DFORCE_NO_CUDA -DVERSION=\"oracle\""
  OBJ="$OUT/obj_$P"; mkdir -p "$OBJ"
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
  for f in sopalin3d starpu_submit_tasks csc_intern_compute raff_functions starpu_updo; do
    add $R/sopalin/src/$f.c p_${f}_po -DCHOL_SOPALIN
    add $R/sopalin/src/$f.c p_${f}_ge -DSOPALIN_LU
    add $R/sopalin/src/$f.c p_${f}_sy -DNOEXTRADEF_SY
    add $R/sopalin/src/$f.c p_${f}_he -DHERMITIAN
  done
  add "$HERE/ref_driver.c" x_ref_driver -DCHOL_SOPALIN
  # compile in parallel; a job only re-runs when its object is missing or older than the script/driver
  if [ -f "$OUT/libpastix_ref_$P.so" ] && [ "$OUT/libpastix_ref_$P.so" -nt "$HERE/ref_driver.c" ] && [ "$OUT/libpastix_ref_$P.so" -nt "$0" ]; then
    echo "[$P] up to date"; continue
  fi
  xargs -P "$(nproc)" -I{} sh -c '{} 2>/dev/null || echo "FAIL: {}"' < "$JOBS" | tee "$OBJ/fail.log"
  if [ -s "$OBJ/fail.log" ]; then echo "[$P] compile failures"; exit 1; fi
  gcc -shared -o "$OUT/libpastix_ref_$P.so" "$OBJ"/*.o "$BLASLIB" -lpthread -lm \
      -Wl,--disable-new-dtags -Wl,-rpath,"$BLASDIR" -Wl,-rpath-link,"$BLASDIR" || exit 1
  echo "[$P] built $OUT/libpastix_ref_$P.so ($(ls "$OBJ"/*.o | wc -l) objects)"
done
