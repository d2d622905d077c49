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

# the 40-step schedule of the reference's regression suite (test/data/custom_dt.txt, used by test/data/voce_pa.toml)
DT_SCHEDULE = [0.005, 0.195] + [0.1] * 19 + [0.2] * 6 + [0.4] * 4 + [0.2, 0.6, 0.5, 0.5, 0.75, 0.75, 0.75, 0.75, 1.0]
assert len(DT_SCHEDULE) == 40


def dt_schedule(nsteps):
    """Step sizes of a run of `nsteps` steps: the reference schedule, its last entry repeated if the run outlasts it."""
    return [DT_SCHEDULE[min(i, len(DT_SCHEDULE) - 1)] for i in range(nsteps)]


# BASELINE.json configs 2-4: grain count and Voronoi seed by mesh size; other sizes keep the grain
# density of the 128^3 workload
def grains_for(n, grains=None):
    known = {32: (100, 32100), 64: (500, 64500), 128: (2000, 1282000)}
    g, seed = known.get(n, (max(4, (2000 * n ** 3) // 128 ** 3), 1000 * n + 7))
    return (grains if grains else g), seed

# FCC Voce property vector of the reference's regression suite (test/data/props_cp_voce.txt)
PROPS_VOCE = [8.920e-6, 0.003435984, 1.0e-10, 168.4, 121.4, 75.2, 44.0, 0.02, 1.0, 400.0e-3, 17.0e-3, 122.4e-3, 0.0,
              5.0e9, 17.0e-3, 0.0, -1.0307952]
BC = ([1, 2, 3, 4], [3, 1, 2, 3], [[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0.001]])
ALG_BYTES_PA_APPLY = 3264      # per element
ALG_BYTES_QPT_UPDATE = 928     # per quadrature point
ALG_BYTES_EA_APPLY = 4992      # per element (576 + 24 + 24 doubles)
ALG_BYTES_QPT_UPDATE_HCP = 1120

# Synthetic HCP (Ti-like) KMBalD property vector for BASELINE.json configs[4].  The reference ships NO hcp property
# set, so these values are ours and HCP results are parity-unpinned against the reference (GPU vs
# oracle only).  Order: rho0, cv, tol | c11 c12 c13 c33 c44 | mu_ref, T_ref | c_1 x4 slip families (basal, prismatic,
# pyramidal<a>, pyramidal<c+a>) | tau_a, p, q | gam_wo, gam_ro, wrD | g_0 x4 | s x4 | k1, k2_0, n^-1, gamma_o,
# rho_dd_ref | c/a | Gruneisen, e_ref
_CV_HCP = 2.5e-3
PROPS_HCP = [4.5e-6, _CV_HCP, 1.0e-10, 162.4, 92.0, 69.0, 180.7, 46.7, 44.0, 300.0, 1944.1, 1944.1, 2100.0, 2400.0,
             4.0e-4, 1.0, 1.0, 1.0, 1.0, 3.0e-2, 8.0e-3, 6.0e-3, 1.2e-2, 2.0e-2, 1.0e-1, 1.0e-1, 1.2e-1, 1.5e-1,
             3.0e-4, 5.0e-5, 0.1, 1.0e-2, 9.0e-4, 1.587, 0.0, -_CV_HCP * 300.0]


def bench_config(which, krylov_iter):
    """The two bench lines: 4 = BASELINE.json configs[3] (the headline: FCC Voce, PA + PCG), 5 = configs[4] (HCP KMBalD,
    B-bar + EA + NRLS, cyclic loading: workflows/Stage3/common_simulation_files/options_master.toml:83-106 with the load
    reversals of test/data/voce_full_cyclic.toml:45-74)."""
    if which == 5:
        rev = {1: 1.0, 11: -1.0, 31: 1.0, 51: -1.0, 71: 1.0}
        return dict(name="HCP KMBalD (synthetic Ti-like set), B-bar + EA + PCG (identity smoother, %d-iter cap), Newton with "
                         "line search, cyclic uniaxial velocity BC" % krylov_iter,
                    xtal=2, kin=2, props=PROPS_HCP, assembly=1, integ=1, nl_solver=1, nr=(5e-5, 5e-10, 25),
                    kr=(1e-7, 1e-27, krylov_iter), dts=lambda nsteps: [0.1] * nsteps,
                    bc_at=lambda step: (BC[0], BC[1], [[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0.001 * rev[step]]]) if step in rev else None,
                    alg_bytes=ALG_BYTES_EA_APPLY, alg_qpt=ALG_BYTES_QPT_UPDATE_HCP,
                    kernel="k_ea_mult_p<2,2,ESS> (EA gradient apply: lane-interleaved element matrices via bulk TMA)")
    return dict(name="FCC Voce, PA + PCG (identity smoother, %d-iter cap), uniaxial velocity BC" % krylov_iter,
                xtal=0, kin=0, props=PROPS_VOCE, assembly=0, integ=0, nl_solver=0, nr=(5e-5, 5e-10, 25),
                kr=(1e-7, 1e-27, krylov_iter), dts=dt_schedule, bc_at=lambda step: BC if step == 1 else None,
                alg_bytes=ALG_BYTES_PA_APPLY, alg_qpt=ALG_BYTES_QPT_UPDATE,
                kernel="k_grad_mult_pa_c<2,2,ESS> (PA gradient apply: compact tangent records via tiled TMA, Jacobians "
                       "rebuilt from coordinates)")


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


def workload(n, ngrains, seed, nz=None):
    from exaconstit_b200 import voxel
    grains = voxel.voronoi_grains(n, n, nz or n, ngrains, seed)
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
    ngrains, seed = grains_for(n, args.grains)
    nz = args.nz or n
    grains, quats = workload(n, ngrains, seed, nz)
    cfg = bench_config(args.config, args.krylov_iter)
    sim = host.VoxelSim((n, n, nz), (1.0, 1.0, nz / n), cfg["xtal"], cfg["kin"], cfg["props"], 298.0, grains, quats,
                        assembly=cfg["assembly"], integ=cfg["integ"], nl_solver=cfg["nl_solver"], nr=cfg["nr"], kr=cfg["kr"],
                        true_jacobi=args.true_jacobi, rank=rank, nranks=nranks, device=local, nccl_id=nccl_id)
    p2p = False
    if nranks > 1 and not args.nccl_only:
        p2p = sim.enable_peer_collectives(dist)      # False: a peer mailbox could not be mapped -> NCCL exchanges on all ranks
    if args.deterministic:
        sim.set_deterministic(True)
    for tv in args.tuning:
        c, v = tv.split(":")
        sim.set_tuning(int(c), int(v))
    state = {"ess": None}

    def apply_bcs(step):
        """UpdateEssBdr at the steps the configuration changes its BCs (1-based); returns True if they changed"""
        bc = cfg["bc_at"](step)
        if bc is None:
            return False
        if nz != n:  # slab diagnostic: same axial strain rate as the cube
            bc = (bc[0], bc[1], [[v * nz / n for v in row] for row in bc[2]])
        _, ess_val = sim.set_bcs(*bc)
        state["ess"] = np.ascontiguousarray(ess_val)
        return True

    def do_step(i):
        changed = apply_bcs(i + 1)
        return sim.step(dts[i], bc_changed=changed, ess_val_host=state["ess"], vel_out_host=vel_out)

    vel_out = np.zeros(3 * sim.nnodes)

    def barrier():
        torch.cuda.synchronize()
        if nranks > 1:
            dist.barrier()
            torch.cuda.synchronize()

    steps = []
    dts = cfg["dts"](args.warmup + args.steps)
    for i in range(args.warmup):
        steps.append(do_step(i))
        if not steps[-1]["converged"]:
            raise SystemExit("bench.py: Newton did not converge in warm-up step %d" % (i + 1))
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
        timed.append(do_step(i))
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
        traffic, traffic_src = ncu_traffic(args.config, ne_local)
        achieved = ne_local * cfg["alg_bytes"] / (apply_ms * 1e-3) / 1e9 if gm_cnt else None
        k1_ms = ms_ms / max(ms_cnt, 1)
        out = {
            "metric": "newton_steps_per_sec", "value": newton / (dev_ms * 1e-3), "unit": "Newton-steps/s",
            "n_gpus": nranks, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%d^3 voxel, %d Voronoi grains, %s" % (n, ngrains, cfg["name"]), "baseline_config": args.config,
                       "mesh": [n, n, nz], "partition": "z-slab x%d" % nranks, "collectives": ("nvlink peer-memory (exchange and all-reduces inside the operator / CG kernels)" if p2p else ("nccl" if nranks > 1 else "none")), "true_jacobi": bool(args.true_jacobi), "deterministic_scatter": bool(args.deterministic),
                       "cache": "inputs >> L2 (%.1f GB of quadrature data per rank)" % (ne_local * 8 * (36 + 9 + 2 * 28 + 12) * 8 / 1e9),
                       "newton_iters": newton, "pcg_iters": pcg, "model_setups": setups, "grad_mults": gmults},
            "e2e": {"value": newton / (e2e_ms * 1e-3), "unit": "Newton-steps/s",
                    "h2d_bytes_per_step": int(state["ess"].nbytes), "d2h_bytes_per_step": int(vel_out.nbytes + 7 * 8)},
            "gpu_launches": int(launches),
            "qpt_updates_per_sec": (ne_local * 8 * nranks) / (k1_ms * 1e-3) if ms_cnt else None,
            "pa_mult_GBps_per_gpu": achieved,
            "roofline": {"kernel": cfg["kernel"], "bound": "hbm",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic,
                         "algorithmic_bytes_per_launch": ne_local * cfg["alg_bytes"],
                         # the same launch time against the bytes the kernel really moves (ncu): the algorithmic figure
                         # counts the 36-entry tangent and the Jacobians, which the default kernel no longer reads
                         "dram_frac": (traffic / (apply_ms * 1e-3) / 1e9 / peak) if (traffic and gm_cnt) else None,
                         "traffic_source": traffic_src,
                         "launches_timed": int(gm_cnt), "avg_launch_ms": apply_ms,
                         "share_of_step": gm_ms / dev_ms},
            "model_setup": {"avg_ms": k1_ms, "calls": int(ms_cnt), "share_of_step": ms_ms / dev_ms,
                            "GBps_algorithmic": ne_local * 8 * cfg["alg_qpt"] / (k1_ms * 1e-3) / 1e9 if ms_cnt else None},
            "clocks": clocks,
            "avg_stress_zz_last": float(timed[-1]["avg_stress"][2]),
            "all_steps_converged": bool(all(s["converged"] for s in timed)),
            "wall_ms": wall_ms,
            # where the step goes (rank 0's kernels, CUDA events): the PCG operator apply, the material update, and the rest
            # (CG vector kernels, exchanges, residual / Jacobians / averages, host waits)
            "step_anatomy": {"k2_grad_mult": gm_ms / dev_ms, "k1_model_setup": ms_ms / dev_ms,
                             "other": 1.0 - (gm_ms + ms_ms) / dev_ms, "us_per_pcg_iteration": dev_ms * 1e3 / max(pcg, 1),
                             "us_per_pcg_iteration_outside_k1": (dev_ms - ms_ms) * 1e3 / max(pcg, 1)},
        }
        if nranks == 1 and not args.no_cpu_baseline and args.config == 4:
            out["cpu_baseline"] = cpu_baseline(n, args.krylov_iter, args.warmup, budget_s=args.cpu_budget)
        out["parity_fingerprint"] = check_fingerprint(n, ngrains, args, steps + timed)
        sim.close()
        sim = None
        if nranks == 1 and not args.no_same_config and args.config == 4:
            # the same time steps on the sample meshes the CPU arm may pick: an un-extrapolated GPU/CPU pair
            # (compare with the reference arm's config.newton_steps_per_sec_on_sample at its config.sample_mesh)
            out["config"]["newton_steps_per_sec_on_sample_meshes"] = {
                str(m): gpu_sample_run(host, m, args, local) for m in CPU_SAMPLE_MESHES}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if sim is not None:
        sim.close()
    if nranks > 1:
        dist.destroy_process_group()


def gpu_sample_run(host, m, args, device):
    """e2e Newton-steps/s of the GPU path on an m^3 sample mesh over the timed steps of the schedule (host buffers in and
    out every step, like the main run)."""
    g, seed = grains_for(m)
    grains, quats = workload(m, g, seed)
    sim = host.VoxelSim((m, m, m), (1.0, 1.0, 1.0), 0, 0, PROPS_VOCE, 298.0, grains, quats, assembly=0,
                        nr=(5e-5, 5e-10, 25), kr=(1e-7, 1e-27, args.krylov_iter), true_jacobi=args.true_jacobi, device=device)
    _, ess_val = sim.set_bcs(*BC)
    vel_out = np.zeros(3 * sim.nnodes)
    dts = dt_schedule(args.warmup + args.steps)
    rs = [sim.step(dts[i], bc_changed=(i == 0), ess_val_host=ess_val, vel_out_host=vel_out) for i in range(len(dts))]
    sim.close()
    timed = rs[args.warmup:]
    return sum(r["newton_iters"] for r in timed) / (sum(r["e2e_ms"] for r in timed) * 1e-3)


FINGERPRINT = os.path.join(ROOT, "tests", "golden", "bench_fingerprint.json")


def check_fingerprint(n, ngrains, args, steps):
    if args.nz and args.nz != n:
        return {"checked": False, "why": "slab diagnostic run"}
    """Parity evidence inside the bench: the volume-averaged stress history of this run (any number of ranks) against
    the stored 1-GPU history of the same workload (tools/make_bench_fingerprint.py), relative to the step's largest
    component.  The 1-GPU history itself is tied to the oracle by tests/test_gpu_system.py at 8^3 and 32^3."""
    if not os.path.exists(FINGERPRINT):
        return {"checked": False, "why": "no fingerprint file"}
    fp = json.load(open(FINGERPRINT))
    key = "%d:%d:%d:%d" % (n, ngrains, args.krylov_iter, int(args.true_jacobi))
    if key not in fp:
        return {"checked": False, "why": "no stored history for " + key}
    ref = np.array(fp[key]["avg_stress"])
    got = np.array([s["avg_stress"] for s in steps])
    m = min(len(ref), len(got))
    err = float(np.max(np.abs(got[:m] - ref[:m]).max(axis=1) / np.abs(ref[:m]).max(axis=1)))
    newton_same = [int(s["newton_iters"]) for s in steps[:m]] == [int(x) for x in fp[key]["newton_iters"][:m]]
    return {"checked": True, "steps_compared": m, "max_rel_err": err, "tol": 1e-8, "ok": bool(err <= 1e-8),
            "newton_iters_identical": bool(newton_same), "stored_run": fp[key].get("run", "1 GPU")}


def ncu_traffic(config, nelems_local):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the committed `ncu --set
    full` captures (profiles/k2_traffic.json: bytes per element by configuration and shard size; see profiles/README.md for
    the build they were taken on).  None if no capture covers this shard size."""
    p = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if not os.path.exists(p):
        return None, None
    t = json.load(open(p))
    ent = t.get("config%d" % config, {}).get(str(nelems_local))
    if not ent:
        return None, None
    return ent["dram_bytes_per_launch"], ent.get("source")


WORKLOAD_FMT = ("%d^3 voxel, %d Voronoi grains, FCC Voce, PA + PCG (identity smoother, %d-iter cap), "
                "uniaxial velocity BC")
CPU_SAMPLE_MESHES = (12, 16, 20, 24, 28, 32)


class CpuArm:
    """The CPU restatement of the reference algorithm (oracle/, OpenMP on every host thread this process may use)
    advancing a REAL simulation of the bench workload on an n_s^3 sample mesh: same material, BCs, dt schedule,
    solver settings and grain statistics, one SystemDriver::Solve per step (src/mechanics_driver.cpp:982-998 times the
    same thing).  Nothing here is hand-assembled: iteration counts and times are whatever the solve takes."""

    def __init__(self, n_s, krylov_iter, true_jacobi=False):
        from oracle import orc
        from exaconstit_b200 import voxel
        self.orc = orc
        self.threads = orc.use_all_host_threads()
        self.n = n_s
        g, seed = grains_for(n_s)
        grains = voxel.voronoi_grains(n_s, n_s, n_s, g, seed)
        quats = voxel.random_quats(g, seed + 1)
        self.ngrains = g
        self.sim = orc.SimStepper((n_s,) * 3, (1.0, 1.0, 1.0), 0, 0, PROPS_VOCE, 298.0, grains, quats, [(1,) + BC], assembly=0,
                                  nr=(5e-5, 5e-10, 25), kr=(1e-7, 1e-27, krylov_iter), true_jacobi=true_jacobi)
        self.krylov_iter = krylov_iter

    def run(self, dts):
        out = []
        for dt in dts:
            r = self.sim.step(dt)
            if r["rc"]:
                raise RuntimeError("CPU arm: Newton failed on the %d^3 sample" % self.n)
            out.append(r)
        return out

    def close(self):
        self.sim.close()


def pick_sample_mesh(nsteps, budget_s, krylov_iter):
    """Largest sample mesh whose `nsteps` steps fit `budget_s` of CPU time on this host, from a 4-step calibration run
    on the smallest mesh: a step costs ~ n^3 elements x n CG iterations per solve."""
    n0 = CPU_SAMPLE_MESHES[0]
    arm = CpuArm(n0, krylov_iter)
    t_step = max(r["seconds"] for r in arm.run(dt_schedule(4)))
    arm.close()
    best = n0
    for n in CPU_SAMPLE_MESHES:
        if nsteps * t_step * (n / n0) ** 4 <= budget_s:
            best = n
    return best, t_step


def scale_to_full(timed, n_s, n_full, krylov_iter):
    """Newton-steps/s of the full mesh from the timed steps of the sample: per-element costs scale with the element
    count; the CG iteration count of the un-preconditioned solve grows like the mesh edge until the iteration cap."""
    newton = sum(r["newton_iters"] for r in timed)
    pcg = sum(r["pcg_iters"] for r in timed)
    t_all = sum(r["seconds"] for r in timed)
    t_pcg = sum(r["pcg_seconds"] for r in timed)
    ratio = (n_full / n_s) ** 3
    pcg_per_newton_s = pcg / max(newton, 1)
    pcg_per_newton_full = min(float(krylov_iter), pcg_per_newton_s * n_full / n_s)
    t_newton_full = ratio * ((t_all - t_pcg) / max(newton, 1) + (t_pcg / max(pcg, 1)) * pcg_per_newton_full)
    return {"value": 1.0 / t_newton_full, "newton_iters": newton, "pcg_iters": pcg, "seconds": t_all, "pcg_seconds": t_pcg,
            "element_ratio": ratio, "pcg_per_newton_sample": pcg_per_newton_s, "pcg_per_newton_full": pcg_per_newton_full,
            "newton_steps_per_sec_on_sample": newton / t_all,
            "model_setups": sum(r["model_setups"] for r in timed)}


def sample_text(n_s, n_full, nsteps, first, sc, threads):
    return ("CPU restatement of the reference algorithm (oracle/, OpenMP, %d threads): %d real time steps (steps %d-%d of the "
            "schedule) of the same workload on a %d^3 sample mesh: %d Newton iterations, %d CG iterations, %.1f s (%.1f s in CG); "
            "scaled to %d^3 by x%.0f elements and %.0f CG iterations per Newton iteration (sample: %.0f, grows with the mesh "
            "edge, capped by the solver's iteration limit)"
            % (threads, nsteps, first + 1, first + nsteps, n_s, sc["newton_iters"], sc["pcg_iters"], sc["seconds"],
               sc["pcg_seconds"], n_full, sc["element_ratio"], sc["pcg_per_newton_full"], sc["pcg_per_newton_sample"]))


def cpu_baseline(n_full, krylov_iter, warmup, budget_s=25.0):
    """cpu_baseline leg of our arm (rank 0, N=1): a bounded run of the CPU arm, reported next to the GPU number."""
    nsteps = 3
    n_s, _ = pick_sample_mesh(warmup + nsteps, budget_s, krylov_iter)
    arm = CpuArm(n_s, krylov_iter)
    dts = dt_schedule(warmup + nsteps)
    arm.run(dts[:warmup])
    timed = arm.run(dts[warmup:])
    arm.close()
    sc = scale_to_full(timed, n_s, n_full, krylov_iter)
    return {"value": sc["value"], "unit": "Newton-steps/s", "cores": arm.threads, "kind": "port",
            "sample": sample_text(n_s, n_full, nsteps, warmup, sc, arm.threads),
            "sample_mesh": [n_s] * 3, "newton_steps_per_sec_on_sample": sc["newton_steps_per_sec_on_sample"],
            "qpt_updates_per_sec": None}


def run_reference(args):
    """Reference arm: the reference's algorithm on the host CPU cores.  The reference itself cannot be built
    here (MFEM/ExaCMech/RAJA/MPI absent), so this is the oracle port with all host threads.  Each step is one real time
    step of the workload on a sample mesh sized to the time budget; `value` is computed from exactly the timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ngrains, _ = grains_for(args.n, args.grains)
    nsteps = args.warmup + args.steps
    n_s = args.ref_n or pick_sample_mesh(nsteps, args.ref_budget, args.krylov_iter)[0]
    arm = CpuArm(n_s, args.krylov_iter, args.true_jacobi)
    dts = dt_schedule(nsteps)
    arm.run(dts[:args.warmup])
    t0 = time.perf_counter()
    timed = arm.run(dts[args.warmup:])
    wall = time.perf_counter() - t0
    arm.close()
    sc = scale_to_full(timed, n_s, args.n, args.krylov_iter)
    v = sc["value"]
    out = {"impl": "reference", "metric": "newton_steps_per_sec", "value": v, "unit": "Newton-steps/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD_FMT % (args.n, ngrains, args.krylov_iter), "mesh": [args.n] * 3,
                      "sample_mesh": [n_s] * 3, "sample_grains": arm.ngrains,
                      # value = (Newton iterations timed / seconds timed) / slowdown: everything below is from the timed steps
                      "newton_steps_per_sec_on_sample": sc["newton_steps_per_sec_on_sample"],
                      "slowdown_sample_to_full": sc["newton_steps_per_sec_on_sample"] / v,
                      "newton_iters": sc["newton_iters"], "pcg_iters": sc["pcg_iters"], "model_setups": sc["model_setups"],
                      "pcg_per_newton_sample": sc["pcg_per_newton_sample"], "pcg_per_newton_full": sc["pcg_per_newton_full"],
                      "element_ratio": sc["element_ratio"], "host_threads": arm.threads},
           "cpu_baseline": {"value": v, "unit": "Newton-steps/s", "cores": arm.threads, "kind": "port",
                            "sample": sample_text(n_s, args.n, args.steps, args.warmup, sc, arm.threads)},
           "e2e": {"value": v, "unit": "Newton-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=128, help="voxels per edge (default: the 128^3 headline workload)")
    ap.add_argument("--nz", type=int, default=0, help="element layers in z (default n); a slab of the workload, e.g. one "
                    "rank's share of an 8-GPU run on one GPU (diagnostics, not a bench line)")
    ap.add_argument("--grains", type=int, default=0, help="grain count (default: BASELINE.json's for the mesh size)")
    ap.add_argument("--config", type=int, default=4, choices=[4, 5], help="4: BASELINE.json configs[3], the headline FCC Voce / PA "
                    "workload (default); 5: configs[4], HCP KMBalD + B-bar + EA + NRLS under cyclic loading")
    ap.add_argument("--krylov-iter", type=int, default=0, help="PCG iteration cap (default: 1000, config 5: 2500 as in options_master.toml)")
    ap.add_argument("--true-jacobi", action="store_true")
    ap.add_argument("--deterministic", action="store_true", help="owner-computes scatter: bitwise reproducible operator (A/B of its cost)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-same-config", action="store_true", help="skip the sample-mesh GPU runs (N=1 only)")
    ap.add_argument("--nccl-only", action="store_true", help="use NCCL for the CG-loop exchanges instead of the peer-memory kernels")
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--ref-budget", type=float, default=240.0, help="seconds of CPU work for --impl reference")
    ap.add_argument("--ref-n", type=int, default=0, help="sample mesh edge of --impl reference (default: sized to --ref-budget)")
    ap.add_argument("--tuning", action="append", default=[], help="ctas:variant pairs passed to exab200_set_tuning (A/B runs)")
    args = ap.parse_args(argv)
    if args.steps < 1 or args.warmup < 0 or args.gpus < 1:
        raise SystemExit("bench.py: need --steps >= 1, --warmup >= 0, --gpus >= 1")
    if not args.krylov_iter:
        args.krylov_iter = 2500 if args.config == 5 else 1000
    if args.impl == "reference" and args.config != 4:
        raise SystemExit("bench.py: the reference arm covers the headline configuration (--config 4) only")
    return args


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
