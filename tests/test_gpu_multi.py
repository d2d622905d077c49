"""Two-rank NCCL run of the C++ host layer against the single-rank run (needs >= 2 GPUs; launched as
subprocesses through torch.distributed.run so each rank owns one device)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
from exaconstit_b200 import host
import refcases
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.tensor(list(host.nccl_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(idt, 0)
inp, gold = refcases.case_inputs(%(case)r)
sim = host.VoxelSim(inp["n"], inp["length"], inp["xtal"], inp["kin"], inp["props"], inp["temp_k"], inp["grain_ids"],
                    inp["quats"], nr=inp["nr"], kr=inp["kr"], rank=rank, nranks=world, device=local,
                    nccl_id=bytes(idt.cpu().tolist()))
if %(p2p)d:
    sim.enable_peer_collectives(dist)
hist = sim.run(inp["dts"][:%(nsteps)d], inp["bcs"])
if rank == 0:
    print("RESULT " + json.dumps(dict(stress=[h["avg_stress"].tolist() for h in hist],
                                      newton=[h["newton_iters"] for h in hist], pcg=[h["pcg_iters"] for h in hist],
                                      halos=sim.counter("halos"), allreduces=sim.counter("allreduces"))))
sim.close()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("p2p,case,env", [(0, "voce_pa", {}), (1, "voce_pa", {}), (1, "voce_full_cyclic_csm", {}),
                                          (1, "voce_pa", {"EXAHOST_NO_HALO_FUSION": "1"}),
                                          (1, "voce_pa", {"EXAHOST_NO_HALO_FUSION": "1", "EXAHOST_NO_CG_FUSION": "1"}),
                                          (1, "voce_pa", {"EXAB200_HALO_CTAS": "3"})],
                         ids=["nccl", "p2p", "p2p-vgrad", "p2p-separate-halo-kernel", "p2p-separate-kernels", "p2p-3-exchange-ctas"])
def test_two_rank_run_matches_single_rank(tmp_path, p2p, case, env):
    """voce_pa: velocity BCs; voce_full_cyclic_csm: velocity-gradient BCs mixed with a velocity BC (the origin of the
    velocity gradient is a MIN over the ranks' coordinates).  Default peer-memory path = operator apply with the
    interface exchange folded in + one-launch CG vector update; the environment switches select the separate kernels."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    nsteps = 5
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, nsteps=nsteps, p2p=p2p, case=case))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(29611 + p2p + 2 * (case != "voce_pa") + 4 * len(env)), str(script)],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[7:])
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refcases
    from exaconstit_b200 import host
    inp, gold = refcases.case_inputs(case)
    sim = host.VoxelSim(inp["n"], inp["length"], inp["xtal"], inp["kin"], inp["props"], inp["temp_k"], inp["grain_ids"],
                        inp["quats"], nr=inp["nr"], kr=inp["kr"])
    h1 = sim.run(inp["dts"][:nsteps], inp["bcs"])
    sim.close()
    s1 = np.array([h["avg_stress"] for h in h1])
    s2 = np.array(res["stress"])
    assert (np.abs(s1 - s2) / np.abs(s1[:, 2:3])).max() < 1e-8
    assert res["newton"] == [h["newton_iters"] for h in h1]
    assert res["halos"] > 0 and res["allreduces"] > 0
    assert (np.abs(s2[:, 2] - gold[:nsteps, 2]) / np.abs(gold[:nsteps, 2])).max() < 3e-5
