"""profiling helper (torchrun, one rank per GPU): back-to-back RHS time of the partitioned mesh per rank"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_pihm_b200  # noqa
from mm_pihm_b200 import lib, watershed as W, partition as PT
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")
size = {2: "2M", 4: "4M", 8: "8M"}[world]
tb = W.make_named(size)
y = W.wet_state(tb, seed=5)
part = PT.partition(tb, world, parts=[rank])[0]
forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:tb["nelem"]], 0))[:, part["elem_gid"]]
m = lib.Model(part, device=rank)
uid = [lib.Model.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
m.comm_init(rank, world, uid[0])
m.set_forcing(forc, np.zeros(part["nriver"]))
yv = m.N_VNew(PT.local_state(part, y)); yd = m.N_VNew()
nrep = int(os.environ.get("NREP", "200"))
for _ in range(10): m.ode_dev(0.0, yv, yd)
m.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(nrep): m.ode_dev(0.0, yv, yd)
m.synchronize()
us = (time.perf_counter() - t0) / nrep * 1e6
out = [None] * world
dist.all_gather_object(out, (rank, round(us, 1), part["nelem"] - part["nown_elem"], len(part["nbr_rank"])))
if rank == 0:
    print(f"[halo_debug={os.environ.get('PIHM_B200_HALO_DEBUG', '0')} paths={m.comm_paths()['halo']}] rhs us per rank (rank, us, ghosts, nbrs):", out, flush=True)
if os.environ.get("HT") == "1":
    # halo protocol timeline (library built with -DPB_HALO_TIMING): averages over the last 128 evaluations
    import ctypes as C
    L = m.L
    L.pihm_b200_debug_halo_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    buf = np.zeros((256, 8), dtype=np.uint64)
    assert L.pihm_b200_debug_halo_times(m.h, None, 1) == 0, "library without PB_HALO_TIMING"
    dist.barrier()
    for _ in range(200): m.ode_dev(0.0, yv, yd)
    m.synchronize()
    L.pihm_b200_debug_halo_times(m.h, buf.ctypes.data_as(C.c_void_p), 0)
    ok = (buf[:, 0] != np.uint64(0xFFFFFFFFFFFFFFFF)) & (buf[:, 7] > 0) & (buf[:, 4] > 0)
    t = buf[ok].astype(np.float64)
    t = t[np.argsort(t[:, 0])][-128:]
    rel = (t - t[:, :1]) / 1e3
    names = ["pre start", "flags raised", "first warp at wait", "last warp at wait", "last warp past wait", "pre end", "main start", "main end"]
    period = np.diff(t[:, 0]).mean() / 1e3
    print(f"[rank {rank}] period {period:.1f} us; " + "; ".join(f"{n} {v:.1f}" for n, v in zip(names, rel.mean(0))), flush=True)
dist.barrier(); m.close(); dist.destroy_process_group()
