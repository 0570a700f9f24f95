"""Multi-GPU host logic on CPU: j-slab partition + combination of per-rank domain statistics over a world_size-2 gloo group.
Each rank runs the (CPU) oracle on its slab -- standing in for its GPU -- and the all-reduced statistics must equal the
single-rank statistics of the whole tile."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from wrfchem_arc_interactions_b200 import partition, synth


def test_jslabs_cover_and_balance():
    rng = np.random.default_rng(0)
    cz = rng.uniform(-0.5, 1.0, (40, 17)).astype(np.float32)
    cz[:10] = -1.0            # a night band: rows are cheap there
    for world in (1, 2, 3, 4, 8):
        slabs = partition.jslabs(cz, world)
        assert slabs[0][0] == 1 and slabs[-1][1] == 40 and len(slabs) == world
        for (a, b), (c, d) in zip(slabs[:-1], slabs[1:]):
            assert b + 1 == c and a <= b
        cost = [17 * (b - a + 1) + 1.6 * (cz[a - 1:b] > 0).sum() for a, b in slabs]
        assert max(cost) <= 1.5 * np.mean(cost) + 17 * 2.6
    with pytest.raises(ValueError):
        partition.jslabs(cz[:3], 4)
    # night rows are cheaper: the first slab of a 2-way split takes more rows than half
    a, b = partition.jslabs(cz, 2)[0]
    assert b > 20


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tempfile
    import torch
    import torch.distributed as dist
    from wrfchem_arc_interactions_b200 import ktables, radiation as R
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dom = synth.make_domain(6, 8, 40, seed=21)
    psw, plw = ktables.write_files(tempfile.mkdtemp())
    orc = O.oracle(); orc.init(dom["p_top"], dom["dims"]["kme"], psw, plw)
    slabs = partition.jslabs(dom["xcoszen"], world)
    jts, jte = slabs[rank]
    dims = dict(dom["dims"]); dims["jts"], dims["jte"] = jts, jte
    flags = R.common_flags(dom)
    sw, lw = R.alloc_outputs(dom, "sw"), R.alloc_outputs(dom, "lw")
    orc.RRTMG_SWRAD(dims, **R.sw_kwargs(dom, sw, **flags)); orc.RRTMG_LWRAD(dims, **R.lw_kwargs(dom, lw, **flags))
    names = [("sw", k) for k in ("swupt", "swuptc", "swuptcln")] + [("lw", k) for k in ("lwupt", "lwuptc", "lwuptcln")]
    st = np.zeros((len(names), 5))
    for f, (w, k) in enumerate(names):
        x = (sw if w == "sw" else lw)[k][jts - 1:jte].astype(np.float64)
        st[f] = [x.sum(), (x * x).sum(), x.size, x.min(), x.max()]
    sums = torch.from_numpy(st[:, :3].copy()); ext = torch.from_numpy(np.stack([-st[:, 3], st[:, 4]], 1))
    dist.all_reduce(sums, op=dist.ReduceOp.SUM); dist.all_reduce(ext, op=dist.ReduceOp.MAX)
    # gather the 2-D fields on rank 0 too (the diagnostic-field gather of SURVEY.md section 8e)
    field = torch.from_numpy(sw["swupt"][jts - 1:jte].copy())
    rows = [torch.zeros(b - a + 1, 6) for a, b in slabs]
    dist.all_gather(rows, field) if all(r.shape == rows[0].shape for r in rows) else None
    if rank == 0:
        q.put((sums.numpy(), ext.numpy(), [tuple(s) for s in slabs]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_statistics_match_single_rank(orc, ktab):
    import torch.multiprocessing as mp
    from wrfchem_arc_interactions_b200 import radiation as R
    from conftest import init, run_pair
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sums, ext, slabs = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dom = synth.make_domain(6, 8, 40, seed=21)
    init(orc, dom, ktab)
    sw, lw = run_pair("sw", orc, dom), run_pair("lw", orc, dom)
    ref = []
    for w, k in [("sw", "swupt"), ("sw", "swuptc"), ("sw", "swuptcln"), ("lw", "lwupt"), ("lw", "lwuptc"), ("lw", "lwuptcln")]:
        x = (sw if w == "sw" else lw)[k].astype(np.float64)
        ref.append([x.sum(), (x * x).sum(), x.size, x.min(), x.max()])
    ref = np.array(ref)
    assert np.allclose(sums, ref[:, :3], rtol=1e-12) and np.allclose(-ext[:, 0], ref[:, 3]) and np.allclose(ext[:, 1], ref[:, 4])
    got = partition.finalize_stats(sums[:, 0], sums[:, 1], sums[:, 2], -ext[:, 0], ext[:, 1])
    x = sw["swupt"].astype(np.float64)
    assert np.isclose(got["mean"][0], x.mean()) and np.isclose(got["sd"][0], x.std(ddof=1)) and np.isclose(got["se"][0], x.std(ddof=1) / np.sqrt(x.size))
    assert slabs[0][0] == 1 and slabs[-1][1] == 8
    # the all-reduced sums feed the decomposition's statistics columns directly (decomposition.stats_from_sums)
    from wrfchem_arc_interactions_b200 import decomposition as D
    cols = D.stats_from_sums(np.column_stack([sums, -ext[:, 0], ext[:, 1]]))
    assert np.allclose(cols["avg"], got["mean"], rtol=1e-14) and np.allclose(cols["stddev"], got["sd"], rtol=1e-9)
    assert np.allclose(cols["standard_error"], got["se"], rtol=1e-9) and np.array_equal(cols["N"], sums[:, 2])


def _gather_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(33)
    nf, nj, ni = 5, 23, 7
    cz = rng.uniform(-0.5, 1.0, (nj, ni)).astype(np.float32)
    cz[:9] = -1.0                                             # unequal slabs: the night band is cheap
    glob = rng.normal(300.0, 40.0, (nf, nj, ni)).astype(np.float32)
    slabs = partition.jslabs(cz, world)
    a, b = slabs[rank]
    g = partition.SlabGather(dist, slabs, rank, nf, ni, torch.device("cpu"))
    mine = [torch.from_numpy(glob[f, a - 1:b].copy()) for f in range(nf)]
    out1 = g(mine).clone()
    out2 = g(mine)                                            # buffers are reused step after step
    ok = bool(np.array_equal(out1.numpy(), glob) and np.array_equal(out2.numpy(), glob))
    # what the gathered fields are for: an order statistic of the whole field, identical on every rank
    med = float(np.sort(out2[0].numpy().ravel())[int(np.floor(np.float32(0.5) * np.float32(nj * ni - 1) + 0.5))])
    q.put((rank, ok, [tuple(s) for s in slabs], med, float(np.sort(glob[0].ravel())[int(np.floor(np.float32(0.5) * np.float32(nj * ni - 1) + 0.5))])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_field_gather_rebuilds_the_global_fields():
    """partition.SlabGather (the all-gather of the 2-D diagnostic fields bench.py runs every step over NCCL) on a
    world-size-2 gloo group with unequal slabs: every rank ends up with the global fields, bit for bit."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, slabs, med, want in res:
        assert ok and med == want
        assert slabs[0][1] - slabs[0][0] != slabs[1][1] - slabs[1][0]          # the slabs really are unequal
