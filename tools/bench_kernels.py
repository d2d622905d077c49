"""Kernel-level timing of the hot-path kernels on one GPU (CUDA events, warm-up, inputs >> L2).
Usage: python tools/bench_kernels.py [N] [--sweep]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from exaconstit_b200 import capi  # noqa: E402
from exaconstit_b200 import voxel  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    return float(np.median(ts)), float(np.min(ts))


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
    sweep = "--sweep" in sys.argv
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    peak = peaks["hbm_gbs"]
    f64 = dict(dtype=torch.float64, device="cuda")
    ne, nn = N ** 3, (N + 1) ** 3
    e2n, coords = voxel.voxel_mesh(N, N, N)
    import refcases
    props = refcases.goldens()["props_cp_voce"]
    ctx = capi.Context(0, 0, props, 298.0, ne, nn, e2n)
    mask = voxel.uniaxial_mask(N, N, N)
    ctx.set_essential_mask(mask)
    xbeg = torch.tensor(coords, **f64)
    vel = torch.tensor(voxel.uniaxial_velocity(coords, nn, 1e-3), **f64)
    jac = torch.empty(ne * 72, **f64)
    dt = 0.1
    ctx.setup_jacobians(xbeg, vel, dt, jac)
    nsv = ctx.nstatev
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(ne, 4, generator=g, **f64)
    q = q / q.norm(dim=1, keepdim=True)
    hist0 = torch.zeros(ne * 8, nsv, **f64)
    hist0[:, 9:13] = q.repeat_interleave(8, dim=0)
    hist0 = hist0.reshape(-1).contiguous()
    ctx.hist_init(hist0)
    s0 = torch.zeros(ne * 48, **f64)
    s1, h1 = torch.empty_like(s0), torch.empty_like(hist0)
    mg = torch.empty(ne * 8 * 36, **f64)
    # advance a few steps so the update works in the plastic regime
    for _ in range(4):
        ctx.model_setup(dt * 2, jac, vel, s0, hist0, s1, h1, mg)
        s0, s1 = s1, s0
        hist0, h1 = h1, hist0
    print("failed points:", ctx.failed_points())
    res = {}
    for mb in ((2, 3) if sweep else (2,)):
        ctx.set_tuning(2, 100 * mb + 10)
        t_med, t_min = timeit(lambda: ctx.model_setup(dt, jac, vel, s0, hist0, s1, h1, mg), iters=5, warm=1)
        res["model_setup_minb%d" % mb] = dict(ms=t_med, qpt_per_s=ne * 8 / t_med * 1e3, GBps=ne * 8 * 928 / t_med / 1e6)
    ctx.set_tuning(2, 210)
    ctx.grad_setup(dt, mg, jac)
    x = torch.randn(3 * nn, **f64)
    y = torch.empty_like(x)
    # 2x: Jacobians rebuilt from the end coordinates (default 20 with 3 CTAs/SM); 1x: J streamed from HBM
    cfgs = [(6, 26)]
    if sweep:
        cfgs = [(6, 26), (5, 26), (3, 20), (11, 27), (8, 27), (8, 28), (6, 28), (4, 29)]
    for ctas, var in cfgs:
        ctx.set_tuning(ctas, var)
        t_med, t_min = timeit(lambda: ctx.grad_mult(x, y), iters=20, warm=3)
        gbs = ne * 3264 / t_med / 1e6
        res["grad_mult_v%d_c%d" % (var, ctas)] = dict(ms=t_med, ms_min=t_min, GBps=gbs, frac=gbs / peak)
    y_jx = y.clone()
    ctx.set_tuning(1, 99)  # stream J
    for ctas, var in ([(2, 10), (1, 12), (2, 15)] if sweep else [(2, 10)]):
        ctx.set_tuning(ctas, var)
        t_med, t_min = timeit(lambda: ctx.grad_mult(x, y), iters=20, warm=3)
        gbs = ne * 3264 / t_med / 1e6
        res["grad_mult_v%d_c%d" % (var, ctas)] = dict(ms=t_med, ms_min=t_min, GBps=gbs, frac=gbs / peak)
    print("JX vs streamed-J max rel diff:", float((y - y_jx).abs().max() / y.abs().max()))
    ctx.set_tuning(2, 10)
    ctx.set_tuning(6, 26)
    # compact tangent records: needs a material update in that format first
    ctx.set_tangent_format(1)
    ctx.model_setup(dt, jac, vel, s0, hist0, s1, h1, mg)
    ctx.grad_setup(dt, mg, jac)
    for ctas, var in ([(6, 30), (5, 30), (4, 31), (3, 32), (8, 33), (11, 33), (3, 34), (8, 35)] if sweep else [(6, 30)]):
        ctx.set_tuning(ctas, var)
        t_med, t_min = timeit(lambda: ctx.grad_mult(x, y), iters=20, warm=3)
        gbs = ne * 3264 / t_med / 1e6
        res["grad_mult_compact_v%d_c%d" % (var, ctas)] = dict(ms=t_med, ms_min=t_min, GBps=gbs, frac=gbs / peak)
    print("compact vs reference-layout max rel diff:", float((y - y_jx).abs().max() / y_jx.abs().max()))
    ctx.set_tuning(6, 30)
    t_med, _ = timeit(lambda: ctx.grad_diag(torch.empty_like(x)))
    res["grad_diag_compact"] = dict(ms=t_med, GBps=ne * 3072 / t_med / 1e6)
    ctx.set_tangent_format(0)
    ctx.model_setup(dt, jac, vel, s0, hist0, s1, h1, mg)
    ctx.grad_setup(dt, mg, jac)
    r = torch.empty_like(x)
    t_med, _ = timeit(lambda: ctx.residual(jac, s1, r))
    res["residual"] = dict(ms=t_med, GBps=ne * 1152 / t_med / 1e6)
    t_med, _ = timeit(lambda: ctx.grad_diag(r))
    res["grad_diag"] = dict(ms=t_med, GBps=ne * 3072 / t_med / 1e6)
    t_med, _ = timeit(lambda: ctx.setup_jacobians(xbeg, vel, dt, jac))
    res["jacobians"] = dict(ms=t_med, GBps=ne * 576 / t_med / 1e6)
    # ---- B-bar residual, element assembly path (config 5) and the HCP material update ----
    ctx_ea = capi.Context(0, 0, props, 298.0, ne, nn, e2n, 1, 1)
    ctx_ea.set_essential_mask(mask)
    t_med, _ = timeit(lambda: ctx_ea.residual(jac, s1, r))
    res["residual_bbar"] = dict(ms=t_med, GBps=ne * 1152 / t_med / 1e6)
    t_med, _ = timeit(lambda: ctx_ea.grad_setup(dt, mg, jac), iters=5, warm=1)
    res["assemble_ea_bbar (incl. memset)"] = dict(ms=t_med, GBps=ne * (4608 + 8 * 45 * 8) / t_med / 1e6)
    y_ref = None
    for ctas, var in [(3, 40), (2, 40), (6, 41), (4, 41), (4, 42), (2, 43), (1, 49)]:
        ctx_ea.set_tuning(ctas, var)
        t_med, t_min = timeit(lambda: ctx_ea.grad_mult(x, y), iters=20, warm=3)
        res["ea_mult_v%d_c%d" % (var, ctas)] = dict(ms=t_med, ms_min=t_min, GBps=ne * 4992 / t_med / 1e6, frac=ne * 4992 / t_med / 1e6 / peak)
        if y_ref is None:
            y_ref = y.clone()
        else:
            assert float((y - y_ref).abs().max() / y_ref.abs().max()) < 1e-13
    ctx_ea.set_tuning(3, 40)
    t_med, _ = timeit(lambda: ctx_ea.grad_diag(r))
    res["ea_diag"] = dict(ms=t_med, GBps=ne * (576 * 8 + 192) / t_med / 1e6)
    ctx_ea.close()
    del ctx_ea
    if N <= 64:
        hprops = refcases.hcp_props()
        ctx_h = capi.Context(2, 2, hprops, 298.0, ne, nn, e2n, 1, 1)
        nsh = ctx_h.nstatev
        hh0 = torch.zeros(ne * 8, nsh, **f64)
        hh0[:, 9:13] = q.repeat_interleave(8, dim=0)
        hh0 = hh0.reshape(-1).contiguous()
        ctx_h.hist_init(hh0)
        sh0 = torch.zeros(ne * 48, **f64)
        sh1, hh1 = torch.empty_like(sh0), torch.empty_like(hh0)
        for _ in range(3):
            ctx_h.model_setup(dt * 3, jac, vel, sh0, hh0, sh1, hh1, mg)
            sh0, sh1 = sh1, sh0
            hh0, hh1 = hh1, hh0
        print("HCP failed points:", ctx_h.failed_points())
        t_med, _ = timeit(lambda: ctx_h.model_setup(dt, jac, vel, sh0, hh0, sh1, hh1, mg), iters=5, warm=1)
        res["model_setup_hcp_kmbald"] = dict(ms=t_med, qpt_per_s=ne * 8 / t_med * 1e3, GBps=ne * 8 * 1120 / t_med / 1e6)
        ctx_h.close()
    t_med, _ = timeit(lambda: y.copy_(x))
    res["copy_vec"] = dict(ms=t_med, GBps=2 * 3 * nn * 8 / t_med / 1e6)
    big = torch.empty(ne * 8 * 36, **f64)
    t_med, _ = timeit(lambda: big.copy_(mg))
    res["copy_big"] = dict(ms=t_med, GBps=2 * big.numel() * 8 / t_med / 1e6)
    for k, v in res.items():
        print(k, json.dumps(v))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(N=N, peak=peak, results=res), open(os.path.join(ROOT, "gpurun_out", "bench_kernels_N%d.json" % N), "w"), indent=1)


if __name__ == "__main__":
    main()
