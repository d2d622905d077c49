"""Diagnostic: distribution of the local-solve evaluation count (history slot 3) after each step of the bench
workload, per point and per warp (32 consecutive points), and the K1 time per call."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from exaconstit_b200 import host  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    torch.cuda.set_device(0)
    grains, quats = bench.workload(n, max(2, 2000 * n ** 3 // 128 ** 3), 1282000)
    sim = host.VoxelSim((n, n, n), (1.0, 1.0, 1.0), 0, 0, bench.PROPS_VOCE, 298.0, grains, quats, assembly=0,
                        nr=(5e-5, 5e-10, 25), kr=(1e-7, 1e-27, 1000))
    sim.set_bcs(*bench.BC)
    sim.kernel_timing(True)
    nsv = sim.nstatev
    for i in range(nsteps):
        sim.kernel_time("model_setup", reset=True)
        st = sim.step(bench.DT_SCHEDULE[i], bc_changed=(i == 0))
        ms, cnt = sim.kernel_time("model_setup")
        h = sim.get("hist").reshape(-1, nsv)
        nf = h[:, 3]
        w = nf.reshape(-1, 32)
        print("step %d dt %.3f newton %d setups %d K1 %.2f ms/call (%.0f M qpt/s) | nfev mean %.2f max %d p99 %.0f | "
              "warp-max mean %.2f | szz %.5f" % (i + 1, bench.DT_SCHEDULE[i], st["newton_iters"], cnt, ms / max(cnt, 1),
                                                 nf.size / (ms / max(cnt, 1)) / 1e3, nf.mean(), nf.max(), np.percentile(nf, 99),
                                                 w.max(axis=1).mean(), st["avg_stress"][2]))
        print("   hist of nfev:", np.bincount(nf.astype(int))[:40].tolist())
    sim.close()


if __name__ == "__main__":
    main()
