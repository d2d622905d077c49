#!/usr/bin/env python
"""Writes tests/golden/bench_fingerprint.json: the volume-averaged stress history of bench.py's workload from a
1-GPU run, which bench.py (any rank count) compares itself with (`parity_fingerprint` in its JSON line).
  python tools/make_bench_fingerprint.py [--n 128] [--steps 40]      (needs a GPU)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, action="append")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--out", default=bench.FINGERPRINT)
    a = ap.parse_args()
    import torch
    from exaconstit_b200 import host
    torch.cuda.set_device(0)
    fp = json.load(open(a.out)) if os.path.exists(a.out) else {}
    for n in (a.n or [128]):
        g, seed = bench.grains_for(n)
        grains, quats = bench.workload(n, g, seed)
        sim = host.VoxelSim((n, n, n), (1.0, 1.0, 1.0), 0, 0, bench.PROPS_VOCE, 298.0, grains, quats, assembly=0,
                            nr=(5e-5, 5e-10, 25), kr=(1e-7, 1e-27, 1000))
        sim.set_bcs(*bench.BC)
        rs = [sim.step(dt, bc_changed=(i == 0)) for i, dt in enumerate(bench.dt_schedule(a.steps))]
        sim.close()
        assert all(r["converged"] for r in rs)
        fp["%d:%d:1000:0" % (n, g)] = {"run": "1 GPU, %s" % torch.cuda.get_device_name(0),
                                        "avg_stress": [list(map(float, r["avg_stress"])) for r in rs],
                                        "newton_iters": [r["newton_iters"] for r in rs],
                                        "pcg_iters": [r["pcg_iters"] for r in rs]}
        print(n, "zz:", np.array([r["avg_stress"][2] for r in rs])[-3:], "newton", sum(r["newton_iters"] for r in rs))
    json.dump(fp, open(a.out, "w"), indent=0)


if __name__ == "__main__":
    main()
