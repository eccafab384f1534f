"""Partitioned RHS on the GPU.
  * one process emulating all ranks on ONE GPU (ghost records staged through
    the host): exercises pihm_b200_create_part, the ghost accessors of the
    kernels and the pack kernel; must equal the unpartitioned RHS bit for bit;
  * the real thing under torch.distributed.run when >= 2 GPUs are visible:
    NCCL halo exchange + all-reduced norms (tests/mgpu_worker.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, partition as PT, watershed as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fbr", [False, True])
@pytest.mark.parametrize("nparts", [2, 5])
def test_partitioned_rhs_emulated_on_one_gpu(fbr, nparts):
    tb = W.make_named("10k", fbr=fbr, dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    y = W.wet_state(tb, seed=4)
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
    single = lib.Model(tb, reorder=1)
    single.set_forcing(forc, np.zeros(nr))
    ref = [single.ODE(0.0, y), single.ODE(0.0, y)]
    parts = PT.partition(tb, nparts)
    models = [lib.Model(p) for p in parts]
    gs = 3 if fbr else 2
    yv = [m.N_VNew(y[p["state_idx"]]) for m, p in zip(models, parts)]
    dv = [m.N_VNew() for m in models]
    for m, p in zip(models, parts):
        m.set_forcing(forc[:, p["elem_gid"]], np.zeros(p["nriver"]))
    for call in range(2):
        packed = [m.halo_pack_host(v) for m, v in zip(models, yv)]
        for q, (mq, pq) in enumerate(zip(models, parts)):
            ge, gr = [], []
            for k, src in enumerate(pq["nbr_rank"]):
                ps = parts[src]
                ks = list(ps["nbr_rank"]).index(q)
                e, r = packed[src]
                ge.append(e[ps["send_e_ptr"][ks] * gs:ps["send_e_ptr"][ks + 1] * gs])
                gr.append(r[ps["send_r_ptr"][ks] * 2:ps["send_r_ptr"][ks + 1] * 2])
            mq.set_ghosts(np.concatenate(ge) if ge else np.zeros(0), np.concatenate(gr) if gr else np.zeros(0))
        dy = np.empty_like(y)
        for m, p, a, b in zip(models, parts, yv, dv):
            m.ode_dev(0.0, a, b)
            dy[p["state_idx"]] = b.download()
        assert np.array_equal(dy, ref[call]), f"call {call}: {np.abs(dy - ref[call]).max():.3e}"
    for m in models:
        m.close()
    single.close()


def exchange_ghosts(models, parts, yv, gs):
    packed = [m.halo_pack_host(v) for m, v in zip(models, yv)]
    for q, (mq, pq) in enumerate(zip(models, parts)):
        ge, gr = [], []
        for k, src in enumerate(pq["nbr_rank"]):
            ps = parts[src]
            ks = list(ps["nbr_rank"]).index(q)
            e, r = packed[src]
            ge.append(e[ps["send_e_ptr"][ks] * gs:ps["send_e_ptr"][ks + 1] * gs])
            gr.append(r[ps["send_r_ptr"][ks] * 2:ps["send_r_ptr"][ks + 1] * 2])
        mq.set_ghosts(np.concatenate(ge) if ge else np.zeros(0), np.concatenate(gr) if gr else np.zeros(0))


@pytest.mark.parametrize("fbr", [False, True])
def test_partitioned_summary_emulated_on_one_gpu(fbr):
    """Summary()/MassBalance() of a partitioned run (SURVEY 8(f) f1 x 8(e)): every rank re-evaluates its
    last RHS call on the ghost records of that call; owned fluxes, mass-balance infil, subrunoff and ws0
    equal the unpartitioned ones bit for bit."""
    nparts = 3
    tb = W.make_named("10k", fbr=fbr, dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    y0 = W.wet_state(tb, seed=4)
    y1 = y0 * (1 + 1e-3 * np.random.default_rng(0).standard_normal(y0.shape))
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y0[:ne], 0))
    single = lib.Model(tb, reorder=1)
    single.set_diagnostics(True)
    single.set_forcing(forc, np.zeros(nr))
    v = single.N_VNew(y0); d = single.N_VNew()
    single.set_ws0(v)
    single.ode_dev(0.0, v, d); single.ode_dev(0.0, v, d)      # the second call sees the first one's bank flows
    v1 = single.N_VNew(y1)
    single.SummaryMB(v1, tb["stepsize"])
    xf_ref, rf_ref = single.get_fluxes()
    sr_ref, ws0_ref = single.get_summary()

    parts = PT.partition(tb, nparts)
    models = [lib.Model(p) for p in parts]
    gs = 3 if fbr else 2
    yv = [m.N_VNew(y0[p["state_idx"]]) for m, p in zip(models, parts)]
    dv = [m.N_VNew() for m in models]
    for m, p, a in zip(models, parts, yv):
        m.set_diagnostics(True)
        m.set_forcing(forc[:, p["elem_gid"]], np.zeros(p["nriver"]))
        m.set_ws0(a)
    for call in range(2):
        exchange_ghosts(models, parts, yv, gs)
        for m, a, b in zip(models, yv, dv):
            m.ode_dev(0.0, a, b)
    xf = np.zeros_like(xf_ref); sr = np.zeros_like(sr_ref); ws0 = np.zeros_like(ws0_ref)
    for m, p in zip(models, parts):
        a1 = m.N_VNew(y1[p["state_idx"]])
        m.SummaryMB(a1, tb["stepsize"])
        no = p["nown_elem"]
        own = p["elem_gid"][:no]
        x, _ = m.get_fluxes()
        s_, w_ = m.get_summary()
        xf[:, own] = x[:, :no]
        sr[own] = s_[:no]
        ws0[p["state_idx"]] = w_
    assert np.array_equal(xf, xf_ref)
    assert np.array_equal(sr, sr_ref) and np.array_equal(ws0, ws0_ref)
    for m in models:
        m.close()
    single.close()


@pytest.mark.parametrize("fbr", [False, True])
def test_nccl_partitioned_run(fbr):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tests", "mgpu_worker.py"), "fbr" if fbr else "pihm", "10k", "20"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert "bitwise: True" in out.stdout
