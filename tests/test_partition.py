"""Multi-GPU host logic on the CPU: the partitioner and its
exchange maps.

  * single process: for 2/3/8 parts the C oracle evaluated on each part's local
    mesh (owned + ghosts, ghost states copied from the global vector) reproduces
    the global RHS bit for bit on the owned unknowns -- pihm and pihm-fbr, with
    the hidden river-edge state (second call);
  * world_size-2 gloo run: every rank owns one part, the ghost records travel
    through torch.distributed send/recv following the send/recv maps exactly as
    the NCCL exchange does, and the assembled dy equals the global one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oraclelib
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import partition as PT, watershed as W


def _global_case(fbr, seed=5):
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    y = W.wet_state(tb, seed=seed)
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
    om = oraclelib.OracleModel(tb)
    om.set_forcing(forc, np.zeros(nr))
    return tb, y, forc, om.ode(y), om.ode(y)


_SUMMARY_CACHE = {}


def _summary_ref(fbr, tb, y, forc, y1):
    """global Summary after two RHS calls on y (cached per variant): (xflux, subrunoff)"""
    if fbr not in _SUMMARY_CACHE:
        om = oraclelib.OracleModel(tb)
        om.set_forcing(forc, np.zeros(tb["nriver"]))
        om.set_ws0(y)
        om.ode(y); om.ode(y)
        sr = om.summary(y1, tb["stepsize"])
        _SUMMARY_CACHE[fbr] = (om.get_fluxes()[0], sr)
    return _SUMMARY_CACHE[fbr]


@pytest.mark.parametrize("fbr", [False, True])
@pytest.mark.parametrize("nparts", [2, 3, 8])
def test_local_meshes_reproduce_global_rhs(fbr, nparts):
    tb, y, forc, dy1, dy2 = _global_case(fbr)
    ne, nr = tb["nelem"], tb["nriver"]
    parts = PT.partition(tb, nparts)
    owned_e = np.concatenate([p["elem_gid"][:p["nown_elem"]] for p in parts])
    owned_r = np.concatenate([p["riv_gid"][:p["nown_riv"]] for p in parts])
    assert sorted(owned_e.tolist()) == list(range(ne))          # a partition: every entity owned once
    assert sorted(owned_r.tolist()) == list(range(nr))
    for p in parts:
        nl, rl = p["nelem"], p["nriver"]
        idx = PT.state_index(nl, rl, fbr, p["elem_gid"], p["riv_gid"], ne, nr)
        ol = oraclelib.OracleModel(p)
        ol.set_forcing(forc[:, p["elem_gid"]], np.zeros(rl))
        own = PT.state_index(p["nown_elem"], p["nown_riv"], fbr, np.arange(nl), np.arange(rl), nl, rl)
        # the helpers of the GPU tests: whole-local-mesh state, owned positions in it, ghost records
        assert np.array_equal(PT.local_state(p, y, extended=True), y[idx])
        assert np.array_equal(PT.owned_in_local(p), own)
        assert np.array_equal(PT.local_state(p, y, extended=True)[own], PT.local_state(p, y))
        ge, gr = PT.ghost_records(p, y)
        gs = 3 if fbr else 2
        no_, ro_ = p["nown_elem"], p["nown_riv"]
        assert np.array_equal(ge.reshape(-1, gs)[:, 0], y[idx][no_:nl])
        assert np.array_equal(ge.reshape(-1, gs)[:, 1], y[idx][2 * nl + no_:3 * nl])
        assert np.array_equal(gr.reshape(-1, 2)[:, 0], y[idx][3 * nl + ro_:3 * nl + rl])
        assert np.array_equal(ol.ode(y[idx])[own], dy1[p["state_idx"]])
        assert np.array_equal(ol.ode(y[idx])[own], dy2[p["state_idx"]])   # stale river-edge flows carried locally
        # Summary()/MassBalance() of the owned elements from the local fluxes = the global ones
        y1 = y * 1.001
        og = _summary_ref(fbr, tb, y, forc, y1)
        ol.set_ws0(y[idx])
        sr_l = ol.summary(y1[idx], tb["stepsize"])
        no = p["nown_elem"]
        assert np.array_equal(ol.get_fluxes()[0][:, :no], og[0][:, p["elem_gid"][:no]])
        assert np.array_equal(sr_l[:no], og[1][p["elem_gid"][:no]])
        # exchange maps are consistent: what p sends to q is what q expects from p
        for k, q in enumerate(p["nbr_rank"]):
            pq = parts[q]
            kq = list(pq["nbr_rank"]).index(p["part"])
            sent = p["elem_gid"][p["send_e_idx"][p["send_e_ptr"][k]:p["send_e_ptr"][k + 1]]]
            off = pq["nown_elem"] + int(pq["recv_e_cnt"][:kq].sum())
            assert np.array_equal(sent, pq["elem_gid"][off:off + pq["recv_e_cnt"][kq]])
            sent_r = p["riv_gid"][p["send_r_idx"][p["send_r_ptr"][k]:p["send_r_ptr"][k + 1]]]
            offr = pq["nown_riv"] + int(pq["recv_r_cnt"][:kq].sum())
            assert np.array_equal(sent_r, pq["riv_gid"][offr:offr + pq["recv_r_cnt"][kq]])


def _exchange(part, y_own, fbr):
    """the halo exchange over gloo: pack owned records, send/recv per neighbour"""
    gs = 3 if fbr else 2
    no, ro = part["nown_elem"], part["nown_riv"]
    blocks = [y_own[:no], y_own[2 * no:3 * no]]
    if fbr:
        blocks.append(y_own[4 * no + 2 * ro:5 * no + 2 * ro])
    rec_e = np.stack(blocks, 1)                                   # [nown, gs]
    rec_r = np.stack([y_own[3 * no:3 * no + ro], y_own[3 * no + ro:3 * no + 2 * ro]], 1)
    ghosts_e = np.zeros((part["nelem"] - no, gs)); ghosts_r = np.zeros((part["nriver"] - ro, 2))
    reqs, bufs = [], []
    offe = offr = 0
    for k, q in enumerate(part["nbr_rank"]):
        se = rec_e[part["send_e_idx"][part["send_e_ptr"][k]:part["send_e_ptr"][k + 1]]]
        sr = rec_r[part["send_r_idx"][part["send_r_ptr"][k]:part["send_r_ptr"][k + 1]]]
        ne_, nr_ = int(part["recv_e_cnt"][k]), int(part["recv_r_cnt"][k])
        out = torch.from_numpy(np.concatenate([se.ravel(), sr.ravel()]).copy())
        inp = torch.zeros(ne_ * gs + nr_ * 2, dtype=torch.float64)
        if out.numel():
            reqs.append(dist.isend(out, int(q)))
        if inp.numel():
            reqs.append(dist.irecv(inp, int(q)))
        bufs.append((inp, offe, ne_, offr, nr_))
        offe += ne_; offr += nr_
    for r in reqs:
        r.wait()
    for inp, oe, ne_, orr, nr_ in bufs:
        a = inp.numpy()
        ghosts_e[oe:oe + ne_] = a[:ne_ * gs].reshape(ne_, gs)
        ghosts_r[orr:orr + nr_] = a[ne_ * gs:].reshape(nr_, 2)
    return ghosts_e, ghosts_r


def _worker(rank, world, port, fbr, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tb, y, forc, dy1, _ = _global_case(fbr)
    part = PT.partition(tb, world, parts=[rank])[0]
    no, ro, nl, rl = part["nown_elem"], part["nown_riv"], part["nelem"], part["nriver"]
    y_own = y[part["state_idx"]]                       # this rank's N_Vector
    ge, gr = _exchange(part, y_own, fbr)
    # local extended state = owned + received ghosts
    surf = np.concatenate([y_own[:no], ge[:, 0]]); gw = np.concatenate([y_own[2 * no:3 * no], ge[:, 1]])
    unsat = np.concatenate([y_own[no:2 * no], np.zeros(nl - no)])
    stg = np.concatenate([y_own[3 * no:3 * no + ro], gr[:, 0]]); rgw = np.concatenate([y_own[3 * no + ro:3 * no + 2 * ro], gr[:, 1]])
    blocks = [surf, unsat, gw, stg, rgw]
    if fbr:
        blocks += [np.concatenate([y_own[3 * no + 2 * ro:4 * no + 2 * ro], np.zeros(nl - no)]),
                   np.concatenate([y_own[4 * no + 2 * ro:], ge[:, 2]])]
    ol = oraclelib.OracleModel(part)
    ol.set_forcing(forc[:, part["elem_gid"]], np.zeros(rl))
    d = ol.ode(np.concatenate(blocks))
    own = PT.state_index(no, ro, fbr, np.arange(nl), np.arange(rl), nl, rl)
    ok = bool(np.array_equal(d[own], dy1[part["state_idx"]]))
    # scalar all-reduce of a WRMS-type sum, as behind every N_VWrmsNorm
    t = torch.tensor([float((d[own] ** 2).sum()), float(len(own))], dtype=torch.float64)
    dist.all_reduce(t)
    ok = ok and abs(t[0].item() - float((dy1 ** 2).sum())) <= 1e-12 * float((dy1 ** 2).sum()) and int(t[1].item()) == len(dy1)
    out_q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fbr", [False, True])
def test_halo_exchange_world_size_2_gloo(fbr):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fbr, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
