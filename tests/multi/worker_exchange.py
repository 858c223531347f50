"""torchrun worker (world size >= 2, one GPU per rank): the view-sharded exchange against a single process.

Every rank builds the same seeded scene.  (1) Single-process truth: the rank renders ALL world*V views itself and
back-propagates.  (2) View-sharded: ``ViewShardedExchange`` installed, the rank renders only its views
``r, r+G, ...`` (``dist.shard_views``) and back-propagates; the gradients autograd returns must equal (1) within the
gradient tolerance, and must be bit-identical on every rank; the densification statistics after ``sync()`` must
equal the single-process ones (``vis_counts`` / ``max_2Dsize`` exactly); a refinement driven by those statistics
with the same seed must then produce identical tensors on every rank (``freegaussian_model.py:404-491``, ``:530``).
Prints one line ``MULTI-OK ...`` on rank 0 when everything holds."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GRAD_TOL = 1e-3


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from freegaussian_b200 import densify
    from freegaussian_b200.dist import DensificationStats, ViewShardedExchange, shard_views
    from freegaussian_b200.knn import k_nearest
    from freegaussian_b200.rendering import rasterization
    from freegaussian_b200.scenes import make_scene
    from util import grad_rel_err

    V = int(os.environ.get("FG_TEST_VIEWS_PER_RANK", "1"))
    N, W, H = int(os.environ.get("FG_TEST_N", "60000")), 320, 192
    knn3 = lambda m: k_nearest(m.to(dev), 3)[0].cpu()  # noqa: E731
    sc = make_scene(N, W, H, n_views=world * V, recipe="trained_like", seed=4, knn3=knn3).to(dev)
    names = ["means", "quats", "scales", "opacities", "sh", "means_next"]
    kw = dict(packed=False, render_mode="RGB+ED", sh_degree=3, absgrad=True)
    g = torch.Generator().manual_seed(1)
    w_r = torch.randn(world * V, H, W, 4, generator=g).to(dev)
    w_a = torch.randn(world * V, H, W, 1, generator=g).to(dev)
    w_f = torch.randn(world * V, H, W, 2, generator=g).to(dev)

    def run(views, exch):
        p = {n: getattr(sc, n).clone().requires_grad_(True) for n in names}
        r, a, m = rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["sh"], sc.viewmats[views], sc.Ks[views],
                                W, H, means_next=p["means_next"], **kw)
        m["means2d"].retain_grad()
        ((r * w_r[views]).sum() + (a * w_a[views]).sum() + (m["flow"] * w_f[views]).sum()).backward()
        st = DensificationStats(N, dev)
        st.accumulate_local(m["radii"], m["means2d"].absgrad, H, W)
        st.sync(reduce=exch is not None)
        grads = {n: p[n].grad.clone() for n in names}
        return grads, st

    full_g, full_s = run(list(range(world * V)), None)               # single-process truth (no exchange installed)
    xc = ViewShardedExchange().install()
    mine = shard_views(world * V, rank, world)
    for _ in range(3):                                                 # several steps: epochs, parity of the publish block
        sh_g, sh_s = run(mine, xc)
    torch.cuda.synchronize()
    worst = {}
    for n in names:
        worst[n] = grad_rel_err(sh_g[n], full_g[n])
        assert worst[n] < GRAD_TOL, (rank, n, worst[n])
        gathered = [torch.empty_like(sh_g[n]) for _ in range(world)]
        dist.all_gather(gathered, sh_g[n].contiguous())
        for other in gathered:
            assert torch.equal(other, gathered[0]), f"{n}: ranks hold different sums"
    assert torch.equal(sh_s.vis_counts, full_s.vis_counts)
    assert torch.equal(sh_s.max_2Dsize, full_s.max_2Dsize)
    assert grad_rel_err(sh_s.xys_grad_norm, full_s.xys_grad_norm) < GRAD_TOL
    # the generic helper for other tensors (network weights) rides the same kernel
    t = torch.full((1000, 7), float(rank + 1), device=dev)
    xc.all_reduce_(t)
    assert bool((t == world * (world + 1) / 2).all())

    # refinement from the reduced statistics, same seed on every rank -> identical Gaussian sets
    params = {"means": sc.means.clone(), "scales": sc.scales.log(), "quats": sc.quats.clone(),
              "opacities": torch.logit(sc.opacities.clamp(1e-4, 1 - 1e-4))[:, None].contiguous(), "sh": sc.sh.clone()}
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in params.items()}
    gen = torch.Generator(device=dev).manual_seed(1234)
    # statistics scaled so that the thresholds of the default schedule select a healthy share of the Gaussians
    res = densify.refine(params, state, sh_s.xys_grad_norm * 1e-3, sh_s.vis_counts, sh_s.max_2Dsize, step=700,
                         num_train_data=100, last_size=(H, W), generator=gen)
    assert res is not None and res.n_after != N and res.n_split > 0
    for k, v in res.params.items():
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([v.shape[0]], device=dev))
        assert all(int(s) == v.shape[0] for s in sizes), "ranks refined to different sizes"
        gathered = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(gathered, v.contiguous())
        for other in gathered:
            assert torch.equal(other, gathered[0]), f"refine: {k} differs between ranks"
    xc.uninstall()
    dist.barrier()
    if rank == 0:
        print(f"MULTI-OK world={world} views_per_rank={V} N={N} multicast={xc.multicast} "
              f"grad_err={ {k: float(f'{v:.2e}') for k, v in worst.items()} } refine {N}->{res.n_after}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
