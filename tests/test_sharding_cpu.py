"""Host-side multi-GPU logic on CPU: (member, IC) sharding and the gloo rendezvous / max-over-ranks pattern bench.py
uses (world_size 2, backend gloo)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from swift_b200.rollout import shard_trajectories, trajectory_seed


@pytest.mark.parametrize("members,n_ic,world", [(12, 8, 1), (12, 64, 8), (12, 16, 2), (5, 3, 4), (3, 1, 8)])
def test_shards_partition_all_trajectories(members, n_ic, world):
    shards = [shard_trajectories(members, n_ic, r, world) for r in range(world)]
    flat = [t for s in shards for t in s]
    assert sorted(flat) == sorted((m, j) for j in range(n_ic) for m in range(members))
    assert len(set(flat)) == members * n_ic
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1          # balanced (the reference: 2 vs 1 members)
    if n_ic % world == 0:                                                           # members of an IC stay together
        for s in shards:
            assert all(sum(1 for m, j in s if j == jj) == members for jj in {j for _, j in s})
    assert len({trajectory_seed(m, j) for m, j in flat}) == len(flat)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_trajectories(12, 8, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)           # "elapsed ms" of this rank
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((gathered, float(t.item())))
    dist.destroy_process_group()


def test_two_rank_gloo_barrier_and_max_reduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 15.0                                                   # max over ranks, as bench.py reports
    assert sorted(gathered[0] + gathered[1]) == sorted((m, j) for j in range(8) for m in range(12))
    assert len(gathered[0]) == len(gathered[1]) == 48


def _store_worker(rank, world, port, path):
    """generate.py's multi-rank output protocol: rank 0 creates the store (run_on_rank0, generate.py:272-282), a
    barrier, then every rank opens it and writes the trajectories of its own shard -- no lock, no collective."""
    import numpy as np
    from swift_b200.store import ForecastStore
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    members, n_ic, steps, h, w = 3, 3, 2, 4, 8
    variables = ["t2m", "z_500", "z_850", "msl"]
    if rank == 0:
        ForecastStore.create(path, variables, n_ic, members, steps, np.arange(h), np.arange(w), layout="trajectory")
    dist.barrier()
    store = ForecastStore.open(path)
    for m, j in shard_trajectories(members, n_ic, rank, world):
        fields = np.full((1, steps + 1, len(variables), h, w), 100.0 * j + 10.0 * m, dtype=np.float32)
        fields += np.arange(steps + 1, dtype=np.float32).reshape(1, -1, 1, 1, 1)
        store.write_trajectories(j, m, fields)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_fill_one_store_without_locks(tmp_path):
    import numpy as np
    from swift_b200.store import ForecastStore
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    path = str(tmp_path / "fc.zarr")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_store_worker, args=(r, 2, port, path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = ForecastStore.open(path).read_all()                     # [ic, member, lead, channel, h, w]
    want = (100.0 * np.arange(3).reshape(3, 1, 1) + 10.0 * np.arange(3).reshape(1, 3, 1) + np.arange(3).reshape(1, 1, 3))
    np.testing.assert_array_equal(got[..., 0, 0, 0], want.astype(np.float32))
    assert (got == got[..., :1, :1, :1]).all()                    # every chunk complete, none torn or missing


def _allreduce_worker(rank, world, port, q):
    """GradientAllReduce bookkeeping on host tensors: every stage's buffers are averaged over the ranks exactly once."""
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from swift_b200.training import GradientAllReduce

    class FakeEngine:                      # the engine's gradient-buffer contract without a GPU
        def __init__(self):
            g = torch.Generator().manual_seed(100 + rank)
            self.grads = {"head": torch.randn(4, 3, generator=g), "l1": torch.randn(5, generator=g),
                          "l0": torch.randn(2, 2, generator=g), "embed": torch.randn(3, generator=g),
                          "cond": torch.randn(6, generator=g)}

        def stage_buffers(self, kind, layer):
            return [self.grads[{"head": "head", "embed": "embed", "cond": "cond"}.get(kind, f"l{layer}")]]

    class FakeModule(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self._train_engine = FakeEngine()

    m = FakeModule()
    local = {k: v.clone() for k, v in m._train_engine.grads.items()}
    red = GradientAllReduce(m)
    for kind, layer in (("head", -1), ("layer", 1), ("layer", 0), ("embed", -1), ("cond", -1)):
        red.hook(kind, layer)
    red.finish()
    gathered = {}
    for k, v in local.items():
        parts = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(parts, v)
        gathered[k] = torch.stack(parts).mean(0)
    ok = all(torch.allclose(m._train_engine.grads[k], gathered[k], atol=1e-7) for k in local)
    same = []
    for k, v in m._train_engine.grads.items():                 # bit-identical on every rank after the reduction
        parts = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(parts, v)
        same.append(all(torch.equal(parts[0], p) for p in parts))
    q.put((rank, ok and all(same), red.bytes))
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == res[1][2] == (12 + 5 + 4 + 3 + 6) * 4


def _exchange_worker(rank, world, port, q):
    """GradientAllReduce.exchange_conditioning on host tensors (gloo): the per-sample inputs of the replicated conditioning stage
    -- gain / bias gradients, the forward scratch [emb | h1 | c | mod], the auxiliary input and dL/dlogvar -- gathered into the
    global batch in rank-major order, gradients pre-scaled by 1 / world."""
    import os
    import types
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from swift_b200.training import GradientAllReduce

    D, L, B = 4, 1, 2
    L2 = 2 * L
    eng = types.SimpleNamespace(geom=types.SimpleNamespace(dim=D, depth=L), grads={})

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self._train_engine = eng

    def mk(r):
        g = torch.Generator().manual_seed(7 + r)
        dgain, dbias = torch.randn(L2, B, D, generator=g), torch.randn(L2, B, D, generator=g)
        emb, h1, c = (torch.randn(B, D, generator=g) for _ in range(3))
        mod = torch.randn(B, L2 * 2 * D, generator=g)
        aux, dlv = torch.randn(B, 1, generator=g), torch.randn(B, generator=g)
        return dgain, dbias, emb, h1, c, mod, aux, dlv

    dgain, dbias, emb, h1, c, mod, aux, dlv = mk(rank)
    scratch = torch.cat([emb.reshape(-1), h1.reshape(-1), c.reshape(-1), mod.reshape(-1), torch.zeros(5)]).contiguous()
    red = GradientAllReduce(M())
    dg, db, sc, aux_all, WB, dlv_all = red.exchange_conditioning(eng, dgain, dbias, scratch.view(torch.uint8), aux, B, dlv)
    parts = [mk(r) for r in range(world)]
    ok = WB == world * B
    ok &= torch.allclose(dg, torch.cat([p[0] for p in parts], 1) / world) and torch.allclose(db, torch.cat([p[1] for p in parts], 1) / world)
    want_sc = torch.cat([torch.cat([p[i] for p in parts], 0).reshape(-1) for i in (2, 3, 4, 5)])
    ok &= torch.equal(sc.view(torch.float32)[:want_sc.numel()], want_sc)
    ok &= torch.equal(aux_all, torch.cat([p[6] for p in parts], 0))
    ok &= torch.allclose(dlv_all, torch.cat([p[7] for p in parts], 0) / world)
    # without a logvar head the call keeps its five-tuple
    ok &= len(red.exchange_conditioning(eng, dgain, dbias, scratch.view(torch.uint8), aux, B)) == 5
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_exchange_conditioning_with_logvar_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
