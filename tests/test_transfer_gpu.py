"""Pipelined host transfers (csrc/transfer.cu: pihm_b200_forcing_prefetch / _commit, pihm_b200_vec_download_async,
pihm_b200_transfer_wait) against the synchronous calls they overlap (pihm_b200_set_forcing_col,
pihm_b200_vec_download): a model loop that changes its forcing every step and pulls the state every step must see
the same bits either way -- every pulled state, the counters, with and without the locality ordering."""
import numpy as np
import pytest

import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W

pytestmark = pytest.mark.gpu
COLS = (W.F_PCPDRP, W.F_EDIR, W.F_ETT)


def pinned(shape):
    import torch
    return torch.empty(shape, dtype=torch.float64).pin_memory().numpy()


def forcing_columns(tb, k):
    f = W.storm_forcing(tb, 2 * 3600.0 + 240.0 * k)          # a different rain rate every step
    out = pinned((3, tb["nelem"]))
    out[:] = f[list(COLS)]
    out[1] *= 1.0 + 0.01 * k
    return out


@pytest.mark.parametrize("reorder", [1, 0])
@pytest.mark.parametrize("fbr", [False, True])
def test_pipelined_transfers_are_bitwise_the_synchronous_ones(fbr, reorder):
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    nr, nsteps = tb["nriver"], 12
    cols = [forcing_columns(tb, k) for k in range(nsteps + 1)]
    runs = []
    for pipelined in (False, True):
        model = lib.Model(tb, reorder=reorder)
        cv = lib.Cvode(model)
        y = model.N_VNew(tb["y0"])
        model.set_forcing(W.storm_forcing(tb, 2 * 3600.0), np.zeros(nr))
        model.set_diagnostics(True)
        model.set_ws0(y)
        cv.SetCVodeParam(y)
        host = [pinned(model.nsv), pinned(model.nsv)]
        pulled = []
        if pipelined:
            model.forcing_prefetch(COLS, list(cols[0]))
        for k in range(nsteps):
            if pipelined:
                model.forcing_commit()
                model.forcing_prefetch(COLS, list(cols[k + 1]))
            else:
                for j, c in enumerate(COLS):
                    model.set_forcing_col(c, cols[k][j])
            cv.SolveCVode((k + 1) * 60.0, y)
            model.SummaryMB(y, 60.0)
            if pipelined:
                if k > 0:                       # the previous pull overlapped this step: collect it now
                    model.transfer_wait()
                    pulled.append(host[(k - 1) & 1].copy())
                model.download_async(y, host[k & 1])
            else:
                model.L.pihm_b200_vec_download(y.h, host[0].ctypes.data)
                pulled.append(host[0].copy())
        if pipelined:
            model.transfer_wait()
            pulled.append(host[(nsteps - 1) & 1].copy())
        runs.append((pulled, cv.stats(), y.download()))
        assert model.check_nan() == 0
        cv.close(); model.close()
    (pa, sa, ya), (pb_, sb, yb) = runs
    assert len(pa) == len(pb_) == nsteps
    for k in range(nsteps):
        assert np.array_equal(pa[k], pb_[k]), f"state pulled after step {k + 1}: {np.abs(pa[k] - pb_[k]).max():.3e}"
    assert sa == sb and sa["nst"] > nsteps and np.array_equal(ya, yb) and np.array_equal(ya, pa[-1])
    assert not np.array_equal(pa[0], pa[-1])


def test_prefetch_protocol_errors():
    tb = W.make_named("tiny")
    model = lib.Model(tb)
    a = [pinned(tb["nelem"]) for _ in COLS]
    for x in a:
        x[:] = 0.0
    model.forcing_commit()                                   # nothing prefetched: no-op
    model.forcing_prefetch(COLS, a)
    with pytest.raises(RuntimeError):
        model.forcing_prefetch(COLS, a)                      # one prefetch outstanding at a time
    model.forcing_commit()
    model.forcing_prefetch(COLS[:1], a[:1])
    model.forcing_commit()
    model.transfer_wait()                                    # no pull outstanding: returns
    model.close()
