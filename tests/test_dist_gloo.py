"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (ray sharding, gather order, gradient
all-reduce over the duplicated parameter list, parameter broadcast).  No kernel runs here."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vfn_testutil as U
from vfnerf_b200 import dist as vd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- sharding: slices are contiguous, disjoint, cover everything, and slice the draws consistently
        n = 1027
        uv = torch.arange(n * 2, dtype=torch.float32).reshape(n, 2)
        U3 = torch.arange(n * 4, dtype=torch.float32).reshape(n, 4)
        my_uv, my_none, my_U3 = vd.shard_rays([uv, None, U3])
        lo, hi = vd.shard_bounds(n, rank, world)
        assert my_none is None and torch.equal(my_uv, uv[lo:hi]) and torch.equal(my_U3, U3[lo:hi])
        # --- gather: per-rank results come back on rank 0 in ray order, uneven slices included
        rgb = my_uv[:, :1].repeat(1, 3) + 0.5
        dep = my_uv[:, 1:2]
        g_rgb, g_dep = vd.gather_render(rgb, dep, n)
        if rank == 0:
            assert torch.equal(g_rgb, uv[:, :1].repeat(1, 3) + 0.5) and torch.equal(g_dep, uv[:, 1:2])
        else:
            assert g_rgb is None and g_dep is None
        # --- parameters: broadcast makes the replicas identical
        case, z = U.load_golden("small_det")
        st = U.case_state(case, z)
        model = U.make_model(case, st, "cpu")
        with torch.no_grad():
            for p in set(model.parameters()):
                p.add_(float(rank))
        vd.broadcast_parameters(model, src=0)
        w0 = model.vector_field_network.layers[0][0].weight
        assert torch.equal(w0, st["vf_net"]["layers.0.0.weight"])            # rank 0 added 0.0
        # --- gradient all-reduce: average over ranks, every unique parameter exactly once
        for p in set(model.parameters()):
            p.grad = torch.full_like(p, float(rank + 1))
        vd.allreduce_gradients(model, average=True)
        for p in set(model.parameters()):
            assert torch.allclose(p.grad, torch.full_like(p, 1.5))
        # grads still alias nothing weird: a second reduce of the same values is idempotent under averaging
        vd.allreduce_gradients(model, average=True)
        assert torch.allclose(model.density.beta.grad, torch.tensor(1.5))
        # --- flat-gradient mode: the gradient arenas themselves are reduced, the .grad views see the result
        vf_g = model.vector_field_network.arena().enable_flat_grad()
        rn_g = model.rendering_network.arena().enable_flat_grad()
        dn_g = model.density.enable_flat_grad()
        for gflat in (vf_g, rn_g, dn_g):
            gflat.fill_(float(2 * rank + 1))
        vd.allreduce_gradients(model, average=True)
        assert torch.allclose(model.rendering_network.layers[0][0].weight.grad,
                              torch.full_like(model.rendering_network.layers[0][0].weight, 2.0))
        assert torch.allclose(model.density.mean.grad, torch.tensor(2.0)) and torch.allclose(vf_g, torch.full_like(vf_g, 2.0))
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) == "ok" and ret.get(1) == "ok"


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 1024, 816000):
        for w in (1, 2, 3, 8):
            spans = [vd.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
