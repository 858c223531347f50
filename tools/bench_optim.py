"""Times the optimizer step and the refinement at the cfg3 shape (1 M Gaussians) on one GPU:
fg_adam_step (one launch, all groups) against torch.optim.Adam the way the reference runs it (one optimizer
per group, freegaussian_config.py:48-75), default (foreach) and fused=True; refine() per call.
Usage: python tools/bench_optim.py [N]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from freegaussian_b200 import _lib  # noqa: E402
from freegaussian_b200.densify import RefineSchedule, refine  # noqa: E402
from freegaussian_b200.optim import REFERENCE_LRS, GaussianAdam  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    shapes = {"means": (n, 3), "features_dc": (n, 3), "features_rest": (n, 15, 3), "opacities": (n, 1),
              "scales": (n, 3), "quats": (n, 4)}
    params = {k: torch.randn(s, device=dev, generator=g) for k, s in shapes.items()}
    grads = {k: torch.randn(s, device=dev, generator=g) * 1e-3 for k, s in shapes.items()}
    out = {"N": n}

    for label, kw in (("torch_adam_foreach", {}), ("torch_adam_fused", {"fused": True})):
        ps = {k: torch.nn.Parameter(v.clone()) for k, v in params.items()}
        opts = [torch.optim.Adam([ps[k]], lr=REFERENCE_LRS[k], eps=1e-15, **kw) for k in ps]
        for k in ps:
            ps[k].grad = grads[k]

        def ref_step():
            # the reference also rebuilds the SH tensor (model.py:801) and splits its gradient every step
            for o in opts:
                o.step()

        out[label + "_ms"] = timed(ref_step)

    sh = torch.cat([params["features_dc"][:, None], params["features_rest"]], 1).contiguous()
    gsh = torch.cat([grads["features_dc"][:, None], grads["features_rest"]], 1).contiguous()
    opt = GaussianAdam.for_reference_groups(params["means"].clone(), sh, params["opacities"].clone(),
                                            params["scales"].clone(), params["quats"].clone())
    dg = {"means": grads["means"], "sh": gsh, "opacities": grads["opacities"], "scales": grads["scales"],
          "quats": grads["quats"]}
    l0 = _lib.launch_count()
    out["fg_adam_step_ms"] = timed(lambda: opt.step(dg))
    out["fg_adam_launches_per_step"] = (_lib.launch_count() - l0) / 23
    elems = sum(v.numel() for v in params.values())
    out["fg_adam_gbs"] = elems * 28 / (out["fg_adam_step_ms"] * 1e-3) / 1e9  # 16 B read + 12 B written per element

    # refinement at the same size (densify + cull, screen-size tests on)
    rp = {"means": params["means"], "scales": torch.randn(n, 3, device=dev, generator=g) * 1.5 - 4.0,
          "quats": params["quats"], "opacities": torch.randn(n, 1, device=dev, generator=g) * 2, "sh": sh}
    st = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in rp.items()}
    gn = torch.rand(n, device=dev, generator=g) * 0.004
    vc = torch.randint(1, 20, (n,), device=dev, generator=g).float()
    ms = torch.rand(n, device=dev, generator=g) * 0.1
    res = refine(rp, st, gn, vc, ms, 3500, 100, (1920, 1080), RefineSchedule())
    out["refine_n_after"] = res.n_after
    out["refine_ms"] = timed(lambda: refine(rp, st, gn, vc, ms, 3500, 100, (1920, 1080), RefineSchedule()), iters=5)
    moved = (res.n_after + n) * 59 * 4 * 1.5  # params + two moments read (kept rows) and written
    out["refine_note"] = f"{n} -> {res.n_after} rows, params + both Adam moments rebuilt ({moved / 1e9:.2f} GB moved)"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
