"""GPU parity: the CUDA path, called through the C ABI, against (a) the golden
dumps of the unmodified reference and (b) the C oracle on the same inputs."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, lower_mask, relerr, tol

pytestmark = pytest.mark.gpu


def run_cuda(g, nrhs_b=None):
    from pastix_b200 import Sopalin
    s = Sopalin(g, g["prec"], g["facto"], schur=g["schur"])
    s.assemble(g["colptr"], g["rows"], g["values"], g["tvalues"])
    assert abs(s.norm1(g["colptr"], g["values"]) - g["norm1"]) <= 1e-12 * g["norm1"]
    L0, U0 = s.get_coeftab()
    nb = s.factorize(g["critere"])
    L, U = s.get_coeftab()
    return s, (L0, U0), (L, U), nb


@pytest.mark.parametrize("name", golden_names())
def test_factor_and_solve_match_reference(name):
    g = load_golden(name)
    s, (L0, U0), (L, U), nb = run_cuda(g)
    t = tol(g["prec"])
    # assembly is a pure scatter: bit-exact against the oracle's restatement of Csc2solv_cblk
    from oracle.oracle import Oracle
    o = Oracle(g, g["prec"])
    La, Ua = o.assemble(g["colptr"], g["rows"], g["values"], g["tvalues"], herm=False, lu=(g["facto"] == "lu"))
    assert np.array_equal(L0, La)
    if Ua is not None:
        assert np.array_equal(U0, Ua)
    # static pivoting count is exact
    assert nb == g["nbpivot"]
    m = lower_mask(g) if g["facto"] != "lu" else slice(None)
    assert relerr(L[m], g["L"][m]) <= t, "L panels differ from the reference"
    if g["U"] is not None:
        assert relerr(U, g["U"]) <= t, "U panels differ from the reference"
    # inertia (real LDLt)
    if g["facto"] == "ldlt" and g["prec"] in ("s", "d"):
        assert s.inertia() == g["inertia"]
    # solve: permuted rhs -> solution, against the reference's solution
    from pastix_b200.csc import permute_rhs, unpermute_solution
    x = permute_rhs(g["b"], g["permtab"])
    s.solve(x)
    xs = unpermute_solution(x, g["permtab"])
    # a replaced pivot is ~1e-15: the solution of that (numerically singular) system is not a parity quantity
    assert relerr(xs, g["x"]) <= (50 * t if g["nbpivot"] == 0 else 1e-1)
    if g["schur"]:
        # IPARM_SCHUR: the never-factored last cblk is the Schur complement pastix_getSchur hands out, and the
        # Schur unknowns keep their right-hand side bit for bit
        cb = g["cblknbr"]; w = int(g["lcol"][cb - 1] - g["fcol"][cb - 1] + 1)
        S, Sr = s.get_schur(), g["L"][-w * w:].reshape(w, w, order="F")
        if g["facto"] != "lu":
            S, Sr = np.tril(S), np.tril(Sr)
        assert relerr(S, Sr) <= t
        xp, bp = x.reshape(s.n, -1), permute_rhs(g["b"], g["permtab"]).reshape(s.n, -1)
        assert np.array_equal(xp[s.n - w:], bp[s.n - w:])
    s.close()


@pytest.mark.parametrize("name", ["lap7_8_llt_d", "cd_8_lu_d", "lap7shift_6_ldlt_z"])
def test_solve_only_with_reference_factors(name):
    """up_down alone: upload the reference's own factors, solve on the GPU."""
    from pastix_b200 import Sopalin
    from pastix_b200.csc import permute_rhs, unpermute_solution
    g = load_golden(name)
    s = Sopalin(g, g["prec"], g["facto"])
    s.set_coeftab(g["L"], g["U"], factorized=True)
    x = permute_rhs(g["b"], g["permtab"])
    s.solve(x)
    assert relerr(unpermute_solution(x, g["permtab"]), g["x"]) <= 50 * tol(g["prec"])
    s.close()


@pytest.mark.parametrize("name", ["cd_8_lu_d", "lap7_10_llt_d_bs16"])
def test_get_cblk_equals_slab_slices(name):
    """pb200_get_cblk (one panel, the layout of SolverCblk.coeftab / .ucoeftab) against pb200_get_coeftab."""
    g = load_golden(name)
    s, _, (L, U), _ = run_cuda(g)
    off = s.solver.panel_offsets()
    for c in (0, g["cblknbr"] // 2, g["cblknbr"] - 1):
        w = int(g["lcol"][c] - g["fcol"][c] + 1); ld = int(g["stride"][c])
        if U is not None:
            Lc, Uc = s.get_cblk(c, with_u=True)
            assert np.array_equal(Uc, U[off[c]:off[c + 1]].reshape(ld, w, order="F"))
        else:
            Lc = s.get_cblk(c)
        assert np.array_equal(Lc, L[off[c]:off[c + 1]].reshape(ld, w, order="F"))
    s.close()


@pytest.mark.parametrize("name", ["lap7_8_llt_d", "lap27_6_ldlt_d", "cd_8_lu_d", "cd_6_lu_z", "lap7_10_llt_d_bs16"])
def test_graph_replay_equals_stream_launches(name, monkeypatch):
    """PB200_GRAPH=1 (round-2 experiment): the captured launch sequence, first run (capture + launch) and replay, gives
    the factors of the stream-launched schedule."""
    g = load_golden(name)
    s0, _, (L0, U0), nb0 = run_cuda(g)
    s0.close()
    monkeypatch.setenv("PB200_GRAPH", "1")
    from pastix_b200 import Sopalin
    s = Sopalin(g, g["prec"], g["facto"])
    m = lower_mask(g) if g["facto"] != "lu" else slice(None)
    for it in range(3):
        s.assemble(g["colptr"], g["rows"], g["values"], g["tvalues"])
        assert s.factorize(g["critere"]) == nb0
        L, U = s.get_coeftab()
        assert relerr(L[m], L0[m]) <= tol(g["prec"]), (name, it)
        if U0 is not None:
            assert relerr(U, U0) <= tol(g["prec"]), (name, it)
    s.close()


def test_schur_mode_refuses_the_level_sweeps(monkeypatch):
    """Schur mode is implemented on the persistent up_down only; the A/B switch must fail loudly, not solve wrongly."""
    from pastix_b200 import Sopalin, PastixB200Error
    from pastix_b200.csc import permute_rhs
    g = load_golden("lap7_8_llt_d_schur")
    monkeypatch.setenv("PB200_SOLVE_LEVELS", "1")
    s = Sopalin(g, "d", "llt", schur=True)
    s.assemble(g["colptr"], g["rows"], g["values"])
    s.factorize(g["critere"])
    with pytest.raises(PastixB200Error):
        s.solve(permute_rhs(g["b"], g["permtab"]))
    s.close()


def test_state_errors():
    from pastix_b200 import Sopalin, PastixB200Error
    g = load_golden("lap7_6_llt_s")
    s = Sopalin(g, "s", "llt")
    with pytest.raises(PastixB200Error):
        s.factorize(1e-10)          # not assembled
    s.assemble(g["colptr"], g["rows"], g["values"])
    s.factorize(g["critere"])
    with pytest.raises(PastixB200Error):
        s.factorize(g["critere"])   # already factorized
    s.close()


@pytest.mark.parametrize("name", ["lap7_8_ilu2_llt_d", "lap7_8_llt_d", "cd_6_lu_c", "lap27_6_ldlt_d"])
def test_multi_rhs_solve_matches_single_rhs(name):
    """MULT_SMX semantics (sm2xnbr right-hand sides in one up_down, updo.c): every column of a 9-RHS solve
    equals the single-RHS solve of that column.  On the all-small ILU schedule this exercises the transposed
    right-hand-side path of kernels_small.cuh."""
    from pastix_b200 import Sopalin
    from pastix_b200.csc import permute_rhs
    g = load_golden(name)
    s = Sopalin(g, g["prec"], g["facto"])
    s.assemble(g["colptr"], g["rows"], g["values"], g["tvalues"])
    s.factorize(g["critere"])
    b1 = permute_rhs(g["b"], g["permtab"]).reshape(s.n, -1)[:, :1]
    scale = (1.0 + 0.25 * np.arange(9)).astype(b1.real.dtype)
    X = np.asfortranarray(b1 * scale[None, :]).astype(s.dtype, order="F")
    cols = []
    for k in range(9):
        xk = np.array(X[:, k], copy=True)
        s.solve(xk)
        cols.append(xk)
    s.solve(X)
    for k in range(9):
        assert relerr(X[:, k], cols[k]) <= 50 * tol(g["prec"]), (name, k)
    s.close()


@pytest.mark.parametrize("name", ["lap7_8_llt_d", "lap27_6_ldlt_d", "cd_8_lu_d", "cd_6_lu_c", "lap7_10_llt_d_bs16"])
def test_persistent_updown_equals_level_sweeps(name, monkeypatch):
    """The counter-ordered persistent sweeps (kernels_solve_dag.cuh, two launches per solve) against the
    launch-per-level sweeps (kernels_solve.cuh, PB200_SOLVE_LEVELS=1) on the same factors, 5 right-hand sides
    (one full group of 4 + a partial one)."""
    from pastix_b200 import Sopalin
    from pastix_b200.csc import permute_rhs
    g = load_golden(name)
    b1 = permute_rhs(g["b"], g["permtab"]).reshape(-1, 1)
    out = []
    for levels in (False, True):
        if levels:
            monkeypatch.setenv("PB200_SOLVE_LEVELS", "1")
        s = Sopalin(g, g["prec"], g["facto"])
        s.assemble(g["colptr"], g["rows"], g["values"], g["tvalues"])
        s.factorize(g["critere"])
        X = np.asfortranarray(b1 * (1.0 + 0.5 * np.arange(5))[None, :]).astype(s.dtype, order="F")
        s.solve(X)
        if levels:
            assert s.last_launches() > 2
        else:
            assert s.last_launches() == 2, "persistent path not taken"
        out.append(X)
        s.close()
    assert relerr(out[0], out[1]) <= 50 * tol(g["prec"])


@pytest.mark.parametrize("name", ["lap7_8_llt_d", "lap27_6_ldlt_d", "cd_8_lu_d", "cd_6_lu_c", "lap7_10_llt_d_bs16",
                                  "lap7_6_llt_s", "lap7her_6_ldlh_z", "cd_6_lu_z"])
@pytest.mark.parametrize("nrhs", [1, 5])
def test_second_generation_sweeps_equal_the_first(name, nrhs, monkeypatch):
    """The three generations of the persistent sweeps on the same factors: the default (k_dag3, kernels_solve_dag3.cuh:
    independent tile workers + a diagonal team per SM, for one right-hand side; k_fwd_dag / k_bwd_dag for several),
    k_dag2 (PB200_DAG_V2=1: three stages in flight per CTA, 32-row sub-tiles, NR = 1 and NRMAX instantiations) and
    k_fwd_dag / k_bwd_dag for everything (PB200_DAG_V1=1, 64-row tiles)."""
    from pastix_b200 import Sopalin
    from pastix_b200.csc import permute_rhs
    import os
    if not os.path.exists(os.path.join(os.path.dirname(__file__), "golden", name + ".npz")):
        pytest.skip("golden not present")
    g = load_golden(name)
    b1 = permute_rhs(g["b"], g["permtab"]).reshape(-1, 1)
    out = []
    for var in (None, "PB200_DAG_V2", "PB200_DAG_V1"):
        monkeypatch.delenv("PB200_DAG_V2", raising=False)
        monkeypatch.delenv("PB200_DAG_V1", raising=False)
        if var:
            monkeypatch.setenv(var, "1")
        s = Sopalin(g, g["prec"], g["facto"])
        s.assemble(g["colptr"], g["rows"], g["values"], g["tvalues"])
        s.factorize(g["critere"])
        X = np.asfortranarray(b1 * (1.0 + 0.5 * np.arange(nrhs))[None, :]).astype(s.dtype, order="F")
        s.solve(X)
        assert s.last_launches() == 2, "persistent path not taken"
        out.append(X)
        s.close()
    assert relerr(out[0], out[2]) <= 50 * tol(g["prec"])
    assert relerr(out[1], out[2]) <= 50 * tol(g["prec"])


@pytest.mark.parametrize("name", ["lap7_8_llt_d", "lap27_6_ldlt_d", "cd_8_lu_d", "cd_6_lu_z", "lap7_10_llt_d_bs16", "lap7_6_llt_s",
                                  "lap7sing_6_ldlt_d"])
@pytest.mark.parametrize("var,val", [("PB200_PDL", "0"), ("PB200_PDL", "1"), ("PB200_DIAG_CMP", "1")])
def test_launch_and_diagonal_kernel_variants_give_the_same_factors(name, var, val, monkeypatch):
    """The panel chain with plain stream order (PB200_PDL=0), with EVERY chain launch programmatically dependent
    (PB200_PDL=1; default: only launches of at most half a wave), and with the compact shared-memory diagonal kernel
    (PB200_DIAG_CMP=1, real LLt / LDLt; opt-in) against the default schedule: same factors, same pivot count (the
    static-pivot case included)."""
    g = load_golden(name)
    s0, _, (L0, U0), nb0 = run_cuda(g)
    s0.close()
    monkeypatch.setenv(var, val)
    s1, _, (L1, U1), nb1 = run_cuda(g)
    s1.close()
    m = lower_mask(g) if g["facto"] != "lu" else slice(None)
    assert nb1 == nb0
    assert relerr(L1[m], L0[m]) <= tol(g["prec"]), (name, var, val)
    if U0 is not None:
        assert relerr(U1, U0) <= tol(g["prec"]), (name, var, val)
