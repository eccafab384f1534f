"""Worker of the real multi-GPU test (run under torch.distributed.run, one
rank per GPU): partitioned RHS + integrator over NCCL against a single-GPU
run of the same watershed on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_pihm_b200  # noqa: E402,F401
from mm_pihm_b200 import lib, partition as PT, watershed as W  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")           # control plane only; the data plane is the library's own NCCL comm
    fbr = len(sys.argv) > 1 and sys.argv[1] == "fbr"
    size = sys.argv[2] if len(sys.argv) > 2 else "10k"
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    tb = W.make_named(size, fbr=fbr, dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    part = PT.partition(tb, world, parts=[rank])[0]
    model = lib.Model(part, device=local)
    uid = [lib.Model.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    model.comm_init(rank, world, uid[0])
    nsv_glob = (5 if fbr else 3) * ne + 2 * nr
    assert model.nsv_global == nsv_glob, (model.nsv_global, nsv_glob)

    # ---- RHS parity: partitioned (NCCL halo) == single GPU, bit for bit --------------
    yg = W.wet_state(tb, seed=3)
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(yg[:ne], 0))
    model.set_forcing(forc[:, part["elem_gid"]], np.zeros(part["nriver"]))
    yv = model.N_VNew(yg[part["state_idx"]]); dv = model.N_VNew()
    model.set_diagnostics(True)                # Summary()/MassBalance() on every rank
    model.set_ws0(yv)
    outs = []
    for _ in range(2):                         # second call: stale river-edge flows
        model.ode_dev(0.0, yv, dv)
        outs.append(dv.download())
    y1g = yg * (1 + 1e-3 * np.random.default_rng(0).standard_normal(yg.shape))
    model.SummaryMB(model.N_VNew(y1g[part["state_idx"]]), tb["stepsize"])     # re-evaluates the last call on its ghosts
    no = part["nown_elem"]
    xf_own = model.get_fluxes()[0][:, :no]
    sr_own = model.get_summary()[0][:no]
    gathered = [None] * world
    dist.all_gather_object(gathered, (part["state_idx"], outs, part["elem_gid"][:no], xf_own, sr_own))
    ok = True
    if rank == 0:
        single = lib.Model(tb, device=local, reorder=1)
        single.set_forcing(forc, np.zeros(nr))
        single.set_diagnostics(True)
        sv = single.N_VNew(yg); single.set_ws0(sv)
        ref = [single.ODE(0.0, yg), single.ODE(0.0, yg)]
        single.SummaryMB(single.N_VNew(y1g), tb["stepsize"])
        xf_ref, sr_ref = single.get_fluxes()[0], single.get_summary()[0]
        xf = np.zeros_like(xf_ref); sr = np.zeros_like(sr_ref)
        for _, _, own, x, s_ in gathered:
            xf[:, own] = x
            sr[own] = s_
        same = np.array_equal(xf, xf_ref) and np.array_equal(sr, sr_ref)
        print(f"[mgpu] Summary/MassBalance fluxes: partitioned == single GPU bitwise: {same}", flush=True)
        ok = ok and same
        for k in range(2):
            dy = np.empty(nsv_glob)
            for idx, o, *_ in gathered:
                dy[idx] = o[k]
            same = np.array_equal(dy, ref[k])
            print(f"[mgpu] RHS call {k}: partitioned == single GPU bitwise: {same}", flush=True)
            ok = ok and same
        single.close()

    # ---- integrator: NCCL all-reduced norms, lock step with the single-GPU run -----------
    cv = lib.Cvode(model)
    y = model.N_VNew(tb["y0"][part["state_idx"]])
    model.set_stale_ovlflow(np.zeros((3, part["nelem"])))
    model.set_forcing_col(W.F_WS0SURF, np.zeros(part["nelem"]))
    cv.SetCVodeParam(y)
    for k in range(nsteps):
        if k % 15 == 0:
            f = W.storm_forcing(tb, 2 * 3600.0 + k * 60.0)
            model.set_forcing(f[:, part["elem_gid"]], np.zeros(part["nriver"]))
        model.Summary(y)
        cv.SolveCVode((k + 1) * 60.0, y)
    st = cv.stats()
    gathered = [None] * world
    dist.all_gather_object(gathered, (part["state_idx"], y.download(), st["nst"], st["nfe"], st["nli"]))
    if rank == 0:
        assert len({(g[2], g[3], g[4]) for g in gathered}) == 1, "ranks took different control flow"
        ym = np.empty(nsv_glob)
        for idx, yy, *_ in gathered:
            ym[idx] = yy
        single = lib.Model(tb, device=local, reorder=1)
        cv1 = lib.Cvode(single)
        y1 = single.N_VNew(tb["y0"])
        cv1.SetCVodeParam(y1)
        for k in range(nsteps):
            if k % 15 == 0:
                single.set_forcing(W.storm_forcing(tb, 2 * 3600.0 + k * 60.0), np.zeros(nr))
            single.Summary(y1)
            cv1.SolveCVode((k + 1) * 60.0, y1)
        s1 = cv1.stats()
        yr = y1.download()
        unit = 1e-3 * np.abs(yr) + 1e-4
        err = (np.abs(ym - yr) / unit).max()
        print(f"[mgpu] {world} ranks vs 1 GPU after {nsteps} model steps: max err {err:.3e} x (reltol|y|+abstol); "
              f"nst {st['nst']}/{s1['nst']} nfe {st['nfe']}/{s1['nfe']} nli {st['nli']}/{s1['nli']}", flush=True)
        ok = ok and err <= 30.0 and abs(st["nst"] - s1["nst"]) <= 0.25 * s1["nst"]
        cv1.close(); single.close()
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    cv.close(); model.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"[mgpu] rank {rank} done ok={flag[0]}", flush=True)
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
