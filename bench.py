#!/usr/bin/env python
"""Benchmark of the ExaConstit hot path on B200 (see DESIGN.md 'Measurement').

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[3], the one the metric is quoted on; fits one B200):
  128^3 voxel mesh, 2000 Voronoi grains, FCC Voce power-law, PA + PCG (reference settings: identity
  smoother, 1000-iteration cap), uniaxial velocity BCs and dt schedule of test/data/voce_pa.toml.
A "step" is one time step = one SystemDriver::Solve (Newton loop: material update + residual, gradient
setup, PCG with one matrix-free gradient apply per iteration).  metric = Newton iterations per second.
Strong scaling: the same mesh is split into z-slabs over N ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT_SCHEDULE = [0.005, 0.195, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1]
# FCC Voce property vector of the reference's regression suite (test/data/props_cp_voce.txt)
PROPS_VOCE = [8.920e-6, 0.003435984, 1.0e-10, 168.4, 121.4, 75.2, 44.0, 0.02, 1.0, 400.0e-3, 17.0e-3, 122.4e-3, 0.0,
              5.0e9, 17.0e-3, 0.0, -1.0307952]
BC = ([1, 2, 3, 4], [3, 1, 2, 3], [[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0.001]])
ALG_BYTES_PA_APPLY = 3264      # per element
ALG_BYTES_QPT_UPDATE = 928     # per quadrature point


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(n, ngrains, seed):
    from exaconstit_b200 import voxel
    grains = voxel.voronoi_grains(n, n, n, ngrains, seed)
    quats = voxel.random_quats(ngrains, seed + 1)
    return grains, quats


def run_ours(args):
    # libraries (NCCL prints its version banner) must not write to stdout: rank 0's stdout carries ONE JSON line
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from exaconstit_b200 import host
    nranks = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if nranks != args.gpus:
        if nranks == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    nccl_id = None
    if nranks > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(host.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())
    n = args.n
    grains, quats = workload(n, args.grains, 1282000)
    sim = host.VoxelSim((n, n, n), (1.0, 1.0, 1.0), 0, 0, PROPS_VOCE, 298.0, grains, quats, assembly=0,
                        nr=(5e-5, 5e-10, 25), kr=(1e-7, 1e-27, args.krylov_iter), true_jacobi=args.true_jacobi,
                        rank=rank, nranks=nranks, device=local, nccl_id=nccl_id)
    if nranks > 1 and not args.nccl_only:
        sim.enable_peer_collectives(dist)
    for tv in args.tuning:
        c, v = tv.split(":")
        sim.set_tuning(int(c), int(v))
    mask, ess_val = sim.set_bcs(*BC)
    ess_pinned = np.ascontiguousarray(ess_val)
    vel_out = np.zeros(3 * sim.nnodes)

    def barrier():
        torch.cuda.synchronize()
        if nranks > 1:
            dist.barrier()
            torch.cuda.synchronize()

    steps = []
    dts = DT_SCHEDULE
    assert args.warmup + args.steps <= len(dts), "dt schedule has %d steps" % len(dts)
    for i in range(args.warmup):
        steps.append(sim.step(dts[i], bc_changed=(i == 0), ess_val_host=ess_pinned, vel_out_host=vel_out))
    sim.kernel_timing(True)
    sim.kernel_time("grad_mult", reset=True)
    sim.kernel_time("model_setup", reset=True)
    l0 = sim.counter("launches")
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    timed = []
    for i in range(args.warmup, args.warmup + args.steps):
        timed.append(sim.step(dts[i], bc_changed=(i == 0), ess_val_host=ess_pinned, vel_out_host=vel_out))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.counter("launches") - l0
    dev_ms = sum(s["dev_ms"] for s in timed)
    e2e_ms = sum(s["e2e_ms"] for s in timed)
    gm_ms, gm_cnt = sim.kernel_time("grad_mult")
    ms_ms, ms_cnt = sim.kernel_time("model_setup")
    t = torch.tensor([dev_ms, e2e_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if nranks > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms = t.tolist()
    newton = sum(s["newton_iters"] for s in timed)
    pcg = sum(s["pcg_iters"] for s in timed)
    setups = sum(s["model_setups"] for s in timed)
    gmults = sum(s["grad_mults"] for s in timed)
    ne_local = sim.nelems
    peak, peak_src = peaks()
    if rank == 0:
        apply_ms = gm_ms / max(gm_cnt, 1)
        achieved = ne_local * ALG_BYTES_PA_APPLY / (apply_ms * 1e-3) / 1e9 if gm_cnt else None
        k1_ms = ms_ms / max(ms_cnt, 1)
        out = {
            "metric": "newton_steps_per_sec", "value": newton / (dev_ms * 1e-3), "unit": "Newton-steps/s",
            "n_gpus": nranks, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%d^3 voxel, %d Voronoi grains, FCC Voce, PA + PCG (identity smoother, %d-iter cap), "
                                   "uniaxial velocity BC" % (n, args.grains, args.krylov_iter),
                       "mesh": [n, n, n], "partition": "z-slab x%d" % nranks, "collectives": ("nvlink peer-memory kernels" if (nranks > 1 and not args.nccl_only) else ("nccl" if nranks > 1 else "none")), "true_jacobi": bool(args.true_jacobi),
                       "cache": "inputs >> L2 (%.1f GB of quadrature data per rank)" % (ne_local * 8 * (36 + 9 + 2 * 28 + 12) * 8 / 1e9),
                       "newton_iters": newton, "pcg_iters": pcg, "model_setups": setups, "grad_mults": gmults},
            "e2e": {"value": newton / (e2e_ms * 1e-3), "unit": "Newton-steps/s",
                    "h2d_bytes_per_step": int(ess_pinned.nbytes), "d2h_bytes_per_step": int(vel_out.nbytes + 7 * 8)},
            "gpu_launches": int(launches),
            "qpt_updates_per_sec": (ne_local * 8 * nranks) / (k1_ms * 1e-3) if ms_cnt else None,
            "pa_mult_GBps_per_gpu": achieved,
            "roofline": {"kernel": "k_grad_mult_pa_c<2,2,ESS> (PA gradient apply: compact tangent records via tiled TMA, Jacobians rebuilt from coordinates)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": TRAFFIC_NCU.get(n) if nranks == 1 else None,
                         "algorithmic_bytes_per_launch": ne_local * ALG_BYTES_PA_APPLY,
                         # the same launch time against the bytes the kernel really moves (ncu): the algorithmic figure
                         # counts the 36-entry tangent and the Jacobians, which the default kernel no longer reads
                         "dram_frac": (TRAFFIC_NCU[n] / (apply_ms * 1e-3) / 1e9 / peak) if (nranks == 1 and n in TRAFFIC_NCU and gm_cnt) else None,
                         "launches_timed": int(gm_cnt), "avg_launch_ms": apply_ms,
                         "share_of_step": gm_ms / dev_ms},
            "model_setup": {"avg_ms": k1_ms, "calls": int(ms_cnt), "share_of_step": ms_ms / dev_ms,
                            "GBps_algorithmic": ne_local * 8 * ALG_BYTES_QPT_UPDATE / (k1_ms * 1e-3) / 1e9 if ms_cnt else None},
            "clocks": clocks,
            "avg_stress_zz_last": float(timed[-1]["avg_stress"][2]),
            "wall_ms": wall_ms,
        }
        if nranks == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(n, pcg / max(newton, 1), setups / max(newton, 1), budget_s=args.cpu_budget)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    sim.close()
    if nranks > 1:
        dist.destroy_process_group()


# ncu --set full dram bytes (read+write) per K2 launch, filled in from profiles/ (see profiles/README.md)
TRAFFIC_NCU = {}
_tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
if os.path.exists(_tp):
    TRAFFIC_NCU = {int(k): v for k, v in json.load(open(_tp)).items()}


def cpu_sample(n_sample, pcg_iters=8, threads=None):
    """Times the CPU oracle (restatement of the reference algorithm) on an n_sample^3 sub-mesh of the same
    workload: one ModelSetup, one gradient assembly, `pcg_iters` PA applies + CG vector work."""
    from oracle import orc
    from exaconstit_b200 import voxel
    n = n_sample
    ne, nn = n ** 3, (n + 1) ** 3
    e2n, coords = voxel.voxel_mesh(n, n, n)
    G, W = orc.hex8_dshape()
    grains = voxel.voronoi_grains(n, n, n, max(2, 2000 * ne // 128 ** 3), 7)
    quats = voxel.random_quats(int(grains.max()), 8)
    props = np.array(PROPS_VOCE)
    nsv = orc.nhist(0, 0)
    hist0 = np.tile(orc.hist_init(0, 0, props), ne * 8).reshape(ne * 8, nsv)
    hist0[:, 9:13] = np.repeat(quats[grains - 1], 8, axis=0)
    hist0 = hist0.ravel().copy()
    s0 = np.zeros(ne * 48)
    vel = voxel.uniaxial_velocity(coords, nn, 2e-3)
    dt = 0.2
    x = coords.copy()
    velE = orc.gather(e2n, vel)
    # two preparatory updates so the timed one is in the plastic regime
    for _ in range(2):
        x = x + dt * vel
        jac = orc.jacobians(G, orc.gather(e2n, x))
        s0, hist0, _, _ = orc.model_setup(0, 0, props, dt, 298.0, jac, G, velE, s0, hist0)
    x = x + dt * vel
    jac = orc.jacobians(G, orc.gather(e2n, x))
    t = time.perf_counter()
    s1, h1, mg, nfail = orc.model_setup(0, 0, props, dt, 298.0, jac, G, velE, s0, hist0)
    t_ms = time.perf_counter() - t
    t = time.perf_counter()
    c81 = orc.transform_matgrad_4d(mg)
    D = np.zeros(ne * 8 * 81)
    import ctypes as C
    orc.lib().orc_assemble_grad_pa(C.c_long(ne), C.c_double(dt), orc._p(jac), orc._p(W), orc._p(c81), orc._p(D))
    t_ga = time.perf_counter() - t
    xv = np.random.default_rng(0).normal(size=3 * nn)
    t = time.perf_counter()
    for _ in range(pcg_iters):
        xE = orc.gather(e2n, xv)
        yE = np.zeros(ne * 24)
        orc.lib().orc_addmult_grad_pa(C.c_long(ne), orc._p(G), orc._p(D), orc._p(xE), orc._p(yE))
        y = orc.scatter_add(e2n, yE, nn)
        xv = xv + 1e-9 * y  # CG-like vector traffic
        _ = float(xv @ y)
    t_it = (time.perf_counter() - t) / pcg_iters
    return dict(n=n, ne=ne, t_model_setup=t_ms, t_grad_assembly=t_ga, t_pcg_iter=t_it, threads=orc.num_threads())


def cpu_baseline(n_full, pcg_per_newton, setups_per_newton, budget_s=25.0):
    n_s = 32
    s = cpu_sample(n_s)
    scale = (n_full / n_s) ** 3
    t_newton = scale * (s["t_model_setup"] * setups_per_newton + s["t_grad_assembly"] + s["t_pcg_iter"] * pcg_per_newton)
    return {"value": 1.0 / t_newton, "unit": "Newton-steps/s", "cores": s["threads"], "kind": "port",
            "sample": "CPU restatement of the reference algorithm (oracle/, OpenMP) timed on a %d^3 sub-mesh: 1 ModelSetup "
                      "%.2fs, 1 AssembleGradPA %.2fs, PA apply+CG vector work %.3fs/iter; scaled x%.0f elements to %d^3 with "
                      "the GPU run's %.0f PCG iters and %.2f ModelSetups per Newton step"
                      % (n_s, s["t_model_setup"], s["t_grad_assembly"], s["t_pcg_iter"], scale, n_full, pcg_per_newton,
                         setups_per_newton),
            "qpt_updates_per_sec": s["ne"] * 8 / s["t_model_setup"],
            "pa_mult_GBps_reference_layout": s["ne"] * 5760 / s["t_pcg_iter"] / 1e9}


def run_reference(args):
    """Reference arm: the reference's algorithm on the host CPU cores.  The reference itself cannot be built
    here (MFEM/ExaCMech/RAJA/MPI absent), so this is the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    counts = {"pcg_per_newton": 1000.0, "setups_per_newton": 1.5}
    p = os.path.join(ROOT, "profiles", "bench_counts.json")
    if os.path.exists(p):
        counts.update(json.load(open(p)))
    n_s = 32
    samples = []
    for _ in range(args.warmup):
        cpu_sample(16, pcg_iters=2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        samples.append(cpu_sample(n_s, pcg_iters=6))
    wall = time.perf_counter() - t0
    scale = (args.n / n_s) ** 3
    t_newton = np.mean([scale * (s["t_model_setup"] * counts["setups_per_newton"] + s["t_grad_assembly"]
                                 + s["t_pcg_iter"] * counts["pcg_per_newton"]) for s in samples])
    v = 1.0 / t_newton
    out = {"impl": "reference", "metric": "newton_steps_per_sec", "value": v, "unit": "Newton-steps/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "%d^3 voxel, %d Voronoi grains, FCC Voce, PA + PCG (identity smoother, %d-iter cap), "
                                  "uniaxial velocity BC" % (args.n, args.grains, args.krylov_iter), "mesh": [args.n] * 3},
           "cpu_baseline": {"value": v, "unit": "Newton-steps/s", "cores": samples[0]["threads"], "kind": "port",
                            "sample": "each step = oracle on a %d^3 sub-mesh (1 ModelSetup + 1 AssembleGradPA + 6 PCG iterations), "
                                      "scaled x%.0f elements with %.0f PCG iters and %.2f ModelSetups per Newton step"
                                      % (n_s, scale, counts["pcg_per_newton"], counts["setups_per_newton"])},
           "e2e": {"value": v, "unit": "Newton-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=128, help="voxels per edge (default: the 128^3 headline workload)")
    ap.add_argument("--grains", type=int, default=2000)
    ap.add_argument("--krylov-iter", type=int, default=1000)
    ap.add_argument("--true-jacobi", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-only", action="store_true", help="use NCCL for the CG-loop exchanges instead of the peer-memory kernels")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--tuning", action="append", default=[], help="ctas:variant pairs passed to exab200_set_tuning (A/B runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
