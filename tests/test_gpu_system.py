"""System-level GPU parity: the C++ host layer (Newton + PCG over the CUDA kernels) against the CPU
oracle's simulation and the reference's golden stress history on the reference's own regression case."""
import numpy as np
import pytest

import refcases

pytestmark = pytest.mark.gpu


def _gpu_run(name, nsteps, **kw):
    from exaconstit_b200 import host
    inp, gold = refcases.case_inputs(name)
    sim = host.VoxelSim(inp["n"], inp["length"], inp["xtal"], inp["kin"], inp["props"], inp["temp_k"],
                        inp["grain_ids"], inp["quats"], assembly=inp["assembly"], nr=inp["nr"], kr=inp["kr"], **kw)
    hist = sim.run(inp["dts"][:nsteps], inp["bcs"])
    state = dict(stress=sim.get("stress"), hist=sim.get("hist"), vel=sim.get("vel"))
    launches = sim.counter("launches")
    sim.close()
    return hist, state, gold[:nsteps], inp, launches


@pytest.mark.parametrize("name,nsteps", [("voce_pa", 12), ("voce_ea", 6), ("mtsdd_bcc", 8)])
def test_time_history_matches_oracle_and_golden(orc, name, nsteps):
    hist, state, gold, inp, launches = _gpu_run(name, nsteps)
    inp2 = dict(inp)
    inp2["dts"] = inp["dts"][:nsteps]
    ref = orc.sim_run(want_state=True, **inp2)
    s_gpu = np.array([h["avg_stress"] for h in hist])
    # against the oracle: north-star tolerance 1e-8 relative on the averaged Cauchy stress (loaded
    # component as the scale; both sides stop Newton at rel 5e-5 so each is an O(1e-5)-accurate root of the
    # same equations reached along the same iteration path)
    scale = np.abs(ref["stress"][:, 2:3])
    assert (np.abs(s_gpu - ref["stress"]) / scale).max() < 1e-8
    assert [h["newton_iters"] for h in hist] == list(ref["iters"][:, 0])
    # PCG iteration counts are part of the parity record (identity-preconditioned CG like the reference)
    pc = np.array([h["pcg_iters"] for h in hist])
    assert np.abs(pc - ref["iters"][:, 1]).max() <= 2
    # state variables at the end (skip slot 3 = local solver evaluation count)
    nsv = ref["hist"].size // (1000 * 8)
    hg, hr = state["hist"].reshape(-1, nsv), ref["hist"].reshape(-1, nsv)
    for c in range(nsv):
        if c == 3:
            continue
        sc = max(np.abs(hr[:, c]).max(), 1e-12)
        assert np.abs(hg[:, c] - hr[:, c]).max() / sc < 1e-7, c
    assert np.abs(state["stress"] - ref["stress_qp"]).max() / np.abs(ref["stress_qp"]).max() < 1e-8
    # against the reference's golden file: 6 printed digits
    err = np.abs(s_gpu - gold) / np.abs(gold[:, 2:3])
    assert err.max() < 1.5e-5
    assert launches > 0


def test_true_jacobi_reaches_same_answer_with_fewer_iterations():
    h0, _, gold, _, _ = _gpu_run("voce_pa", 5)
    h1, _, _, _, _ = _gpu_run("voce_pa", 5, true_jacobi=True)
    s0 = np.array([h["avg_stress"] for h in h0])
    s1 = np.array([h["avg_stress"] for h in h1])
    assert (np.abs(s0 - s1) / np.abs(s0[:, 2:3])).max() < 1e-5
    assert sum(h["pcg_iters"] for h in h1) <= sum(h["pcg_iters"] for h in h0)


def test_newton_line_search_variant():
    h, _, gold, _, _ = _gpu_run("voce_pa", 4, nl_solver=1)
    s = np.array([x["avg_stress"] for x in h])
    assert (np.abs(s - gold) / np.abs(gold[:, 2:3])).max() < 1.5e-5
