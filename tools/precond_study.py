"""SURVEY 8(f) f4, feasibility on the CPU: what would a (block-)Jacobi preconditioner give SPGMR on this problem?

The reference runs CVSPGMR with PREC_NONE (src/ode.c:426) and needs ~2.4 Krylov iterations per Newton iteration;
every iteration is one RHS evaluation + ~13 vector passes, i.e. the whole cost of the model step scales with it.
This script takes a storm state of a synthetic watershed from the reference's own run (oracle/_ref), builds the
Newton matrix A = I - gamma J of that moment explicitly (J by finite differences of the C oracle's RHS, column by
column), scales it the way CVODE does (D A D^-1 with D = diag(ewt)) and counts the GMRES iterations needed to bring
the scaled residual below CVODE's linear tolerance for
    none         (what runs today)
    jacobi       diag(A)
    block        the 3 x 3 (fbr 5 x 5) diagonal block of each element + 2 x 2 of each river segment
left-preconditioned as CVSPGMR does (cvode_spgmr.c:245-334).  Test/tool infrastructure: uses oracle/ only.

    python tools/precond_study.py [size] [nsteps]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mm_pihm_b200  # noqa: E402,F401
from mm_pihm_b200 import watershed as W  # noqa: E402
import oraclelib  # noqa: E402
import reflib  # noqa: E402

RTOL, ATOL, T0 = 1e-3, 1e-4, 2 * 3600.0


def storm_state(tb, nsteps):
    """the reference's state, step size and order after nsteps model steps of the bench window"""
    ne, nr = tb["nelem"], tb["nriver"]
    ref = reflib.RefModel(fbr=False).create_from_tables(tb)
    ref.init_state(tb["y0"]); ref.set_ovlflow(np.zeros((3, ne))); ref.set_cvode_param()
    for k in range(nsteps):
        if k % 15 == 0:
            f = W.storm_forcing(tb, T0 + k * 60.0)
        fr = f.copy(); fr[W.F_WS0SURF] = ref.get_ws()[:ne]
        ref.set_forcing(fr, np.zeros(nr))
        ref.model_step(k)
    st = ref.stats()
    y, ovl = ref.get_y(), ref.get_ovlflow()
    fr = f.copy(); fr[W.F_WS0SURF] = ref.get_ws()[:ne]
    ref.close()
    return y, ovl, fr, st


def jacobian(om, y, ovl, f0):
    """dense-by-columns finite-difference Jacobian of the oracle RHS as a sparse matrix"""
    n = y.size
    rows, cols, vals = [], [], []
    for j in range(n):
        sig = 1e-7 * max(abs(y[j]), 1e-2)
        yp = y.copy(); yp[j] += sig
        om.set_stale_ovlflow(ovl)
        d = (om.ode(yp) - f0) / sig
        nz = np.nonzero(d)[0]
        rows.append(nz); cols.append(np.full(nz.size, j)); vals.append(d[nz])
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def gmres_iters(A, b, Minv, tol, maxit=40):
    """left-preconditioned GMRES (modified Gram-Schmidt, no restart): iterations until the preconditioned
    residual has dropped below tol * |Minv b|"""
    r0 = Minv(b)
    beta = np.linalg.norm(r0)
    V = [r0 / beta]
    H = np.zeros((maxit + 1, maxit))
    for k in range(maxit):
        w = Minv(A @ V[k])
        for i in range(k + 1):
            H[i, k] = w @ V[i]
            w = w - H[i, k] * V[i]
        H[k + 1, k] = np.linalg.norm(w)
        e1 = np.zeros(k + 2); e1[0] = beta
        yk, res, *_ = np.linalg.lstsq(H[:k + 2, :k + 1], e1, rcond=None)
        rn = np.linalg.norm(H[:k + 2, :k + 1] @ yk - e1)
        if rn <= tol * beta:
            return k + 1
        V.append(w / H[k + 1, k])
    return maxit


def main():
    size = sys.argv[1] if len(sys.argv) > 1 else "10k"
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    tb = W.make_named(size)
    ne, nr = tb["nelem"], tb["nriver"]
    t0 = time.time()
    y, ovl, forc, st = storm_state(tb, nsteps)
    q, h = st["qcur"], st["hcur"]
    l0 = 1.0 / sum(1.0 / j for j in range(1, q + 1))            # gamma = h * l0 of fixed-step BDF order q
    h_avg = nsteps * 60.0 / max(st["nst"], 1)
    print(f"{size}: reference after {nsteps} steps: next h = {h:.3f} s, mean h so far {h_avg:.2f} s, order {q}; "
          f"its own run: {st['nli'] / max(st['nni'], 1):.2f} Krylov iterations per Newton iteration")
    om = oraclelib.OracleModel(tb)
    om.set_forcing(forc, np.zeros(nr))
    om.set_stale_ovlflow(ovl)
    f0 = om.ode(y)
    J = jacobian(om, y, ovl, f0)
    n = y.size
    print(f"Jacobian: {J.nnz} non-zeros ({J.nnz / n:.1f} per row), {time.time() - t0:.0f} s")
    ewt = 1.0 / (RTOL * np.abs(y) + ATOL)
    D, Di = sp.diags(ewt), sp.diags(1.0 / ewt)
    idx_blocks = [np.array([i, ne + i, 2 * ne + i]) for i in range(ne)] + \
                 [np.array([3 * ne + r, 3 * ne + nr + r]) for r in range(nr)]
    rng = np.random.default_rng(1)
    rnd = rng.standard_normal(n)
    for label, hh in (("next h", h), ("mean h", h_avg), ("h = 60 s (the cap)", 60.0)):
        gamma = hh * l0
        A = (D @ (sp.identity(n) - gamma * J) @ Di).tocsr()      # the scaled Newton matrix SPGMR sees
        dg = A.diagonal()
        Ad = A.tocsc()
        inv_blocks = [np.linalg.inv(Ad[np.ix_(ix, ix)].toarray()) for ix in idx_blocks]

        def block_solve(v):
            out = np.empty_like(v)
            for ix, B in zip(idx_blocks, inv_blocks):
                out[ix] = B @ v[ix]
            return out
        precs = {"none": lambda v: v, "jacobi": lambda v: v / dg, "block": block_solve}
        rhs = {"gamma f(y) (first Newton residual)": ewt * (gamma * f0), "random": rnd}
        # CVODE: delta = eplifac * tq[4] with eplifac 0.05, tq[4] ~ nlscoef 0.1 / (BDF error constant): the residual
        # has to fall to a few 1e-3 .. 1e-2 of a typical right-hand side (WRMS ~ 0.1 .. 1)
        print(f"gamma = {gamma:.2f} ({label}):")
        for tol in (3e-2, 1e-2, 1e-3):
            for name, b in rhs.items():
                its = {p: gmres_iters(A, b, M, tol) for p, M in precs.items()}
                print(f"   tol {tol:<6g} {name:36s} " + "  ".join(f"{p} {v:2d}" for p, v in its.items()))
    gamma = h_avg * l0
    ev = np.abs(gamma * J.diagonal())
    print(f"|gamma J_ii|: median {np.median(ev):.3f}, 99th percentile {np.percentile(ev, 99):.2f}, max {ev.max():.1f}")


if __name__ == "__main__":
    main()
