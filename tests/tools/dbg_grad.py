import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import render as O
from util import grad_rel_err, small_scene
from freegaussian_b200.rendering import rasterization
W, H = 96, 64
for mode in ("mean", "cov"):
    sc = small_scene(1500, W, H, views=2, seed=17)
    names = ["means", "quats", "scales", "opacities", "sh", "means_next", "quats_next"]
    d = sc.to("cuda")
    gp = {n: getattr(d, n).clone().requires_grad_(True) for n in names}
    op = {n: getattr(sc, n).clone().double().requires_grad_(True) for n in names}
    kw = dict(packed=False, render_mode="RGB+ED", sh_degree=3, absgrad=True, flow_mode=mode)
    ex_g = dict(quats_next=gp["quats_next"]) if mode == "cov" else {}
    ex_o = dict(quats_next=op["quats_next"]) if mode == "cov" else {}
    r, a, m = rasterization(gp["means"], gp["quats"], gp["scales"], gp["opacities"], gp["sh"], d.viewmats, d.Ks, W, H, means_next=gp["means_next"], **ex_g, **kw)
    rr, ra, rm = O.rasterization(op["means"], op["quats"], op["scales"], op["opacities"], op["sh"], sc.viewmats.double(), sc.Ks.double(), W, H, means_next=op["means_next"], **ex_o, **kw)
    gen = torch.Generator().manual_seed(0)
    wr, wf = torch.randn(rr.shape, generator=gen), torch.randn(rm["flow"].shape, generator=gen)
    for scale_f in (1.0, 0.0):
        for p in list(gp.values()) + list(op.values()): p.grad = None
        ((r * wr.cuda()).sum() + scale_f * (m["flow"] * wf.cuda()).sum()).backward(retain_graph=True)
        ((rr * wr.double()).sum() + scale_f * (rm["flow"] * wf.double()).sum()).backward(retain_graph=True)
        print(mode, "flow weight", scale_f, {n: (f"{grad_rel_err(gp[n].grad, op[n].grad):.2e}" if op[n].grad is not None else None) for n in names})
    go, oo = gp["opacities"].grad.cpu().double(), op["opacities"].grad
    err = (go - oo).abs()
    i = int(err.argmax())
    print(" worst opacity idx", i, "err", float(err[i]), "ref", float(oo[i]), "max ref", float(oo.abs().max()), "opac", float(sc.opacities[i]), "flow img max", float(rm["flow"].abs().max()))
