"""Multi-GPU host logic on CPU: view sharding arithmetic and the gradient / densification-statistics exchange
(SUM, SUM, MAX) with gloo, world_size 2."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from freegaussian_b200.dist import DensificationStats, GradBucket, exchange, shard_views


def test_shard_views_partition():
    for n_views, world in [(32, 8), (8, 8), (5, 2), (1, 4)]:
        shards = [shard_views(n_views, r, world) for r in range(world)]
        assert sorted(v for s in shards for v in s) == list(range(n_views))
        assert all(v % world == r for r, s in enumerate(shards) for v in s)


def _views(n, seed):
    g = torch.Generator().manual_seed(seed)
    radii = torch.randint(0, 40, (4, n), generator=g, dtype=torch.int32)
    radii[radii < 12] = 0
    absgrad = torch.rand(4, n, 2, generator=g)
    grads = [torch.randn(n, 3, generator=g), torch.randn(n, 16, 3, generator=g)]
    return radii, absgrad, grads


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    radii, absgrad, grads = _views(n, 0)
    mine = shard_views(4, rank, world)
    stats = DensificationStats(n, "cpu")
    params = [torch.zeros(n, 3, requires_grad=True), torch.zeros(n, 16, 3, requires_grad=True)]
    bucket = GradBucket(params)
    bucket.attach(params)
    for v in mine:
        stats.accumulate_local(radii[v:v + 1], absgrad[v:v + 1], 100, 200)
        for p, g in zip(params, grads):
            p.grad += g * (v + 1)  # pretend per-view gradient
    exchange([bucket.flat])
    stats.sync()
    if rank == 0:
        torch.save({"g": stats.xys_grad_norm, "c": stats.vis_counts, "m": stats.max_2Dsize,
                    "p0": params[0].grad.clone(), "p1": params[1].grad.clone()}, out)
    dist.destroy_process_group()


def test_exchange_matches_single_process(tmp_path):
    n, world = 257, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, 29531, n, out), nprocs=world, join=True)
    got = torch.load(out)
    radii, absgrad, grads = _views(n, 0)
    ref = DensificationStats(n, "cpu")
    for v in range(4):  # a single process seeing all four views (freegaussian_model.py:369-392 per view)
        ref.accumulate_local(radii[v:v + 1], absgrad[v:v + 1], 100, 200)
    ref.reduce()
    assert torch.allclose(got["g"], ref.xys_grad_norm, atol=1e-6)
    assert torch.equal(got["c"], ref.vis_counts)
    assert torch.equal(got["m"], ref.max_2Dsize)
    assert torch.allclose(got["p0"], grads[0] * 10, atol=1e-5) and torch.allclose(got["p1"], grads[1] * 10, atol=1e-5)


def test_stats_match_reference_formula():
    """One view: identical to after_train_iter (freegaussian_model.py:376-392)."""
    n = 50
    radii, absgrad, _ = _views(n, 3)
    s = DensificationStats(n, "cpu")
    s.accumulate_local(radii[:1], absgrad[:1], 100, 200)
    s.reduce()
    vis = radii[0] > 0
    g = torch.zeros(n); c = torch.ones(n); m = torch.zeros(n)
    g[vis] += absgrad[0][vis].norm(dim=-1)
    c[vis] += 1
    m[vis] = torch.maximum(m[vis], radii[0][vis] / 200.0)
    assert torch.allclose(s.xys_grad_norm, g) and torch.equal(s.vis_counts, c) and torch.allclose(s.max_2Dsize, m)


def _worker_sync(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = DensificationStats(n, "cpu")
    for step in range(3):  # local folds every step, one cross-rank sync at the end
        radii, absgrad, _ = _views(n, step)
        for v in shard_views(4, rank, world):
            stats.accumulate_local(radii[v:v + 1], absgrad[v:v + 1], 100, 200)
    stats.sync()
    if rank == 0:
        torch.save({"g": stats.xys_grad_norm, "c": stats.vis_counts, "m": stats.max_2Dsize}, out)
    dist.destroy_process_group()


def test_deferred_stats_sync_equals_per_step_reduction(tmp_path):
    n, world = 129, 2
    out = str(tmp_path / "sync.pt")
    mp.spawn(_worker_sync, args=(world, 29533, n, out), nprocs=world, join=True)
    got = torch.load(out)
    ref = DensificationStats(n, "cpu")
    for step in range(3):
        radii, absgrad, _ = _views(n, step)
        for v in range(4):
            ref.accumulate_local(radii[v:v + 1], absgrad[v:v + 1], 100, 200)
        ref.reduce()
    assert torch.allclose(got["g"], ref.xys_grad_norm, atol=1e-5)
    assert torch.equal(got["c"], ref.vis_counts) and torch.equal(got["m"], ref.max_2Dsize)


def _network_worker(rank, world, port, out):
    """Each rank holds the same deformation network and a rank-dependent gradient for every parameter."""
    from freegaussian_b200.deform import DeformNetwork

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = DeformNetwork(is_blender=True)
    for k, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(k + 1)) * (rank + 1)
    exchange([p.grad for p in net.parameters()])
    if rank == 0:
        torch.save([p.grad.clone() for p in net.parameters()], out)
    dist.destroy_process_group()


def test_network_gradients_ride_the_same_exchange(tmp_path):
    """The deformation network's weight gradients (28 tensors, 2.4 MB) are summed by the exchange step like any other."""
    world = 2
    out = str(tmp_path / "net.pt")
    mp.spawn(_network_worker, args=(world, 29533, out), nprocs=world, join=True)
    got = torch.load(out)
    assert len(got) == 28
    for k, g in enumerate(got):
        assert torch.equal(g, torch.full_like(g, float(k + 1) * 3))  # rank 0 (x1) + rank 1 (x2)
