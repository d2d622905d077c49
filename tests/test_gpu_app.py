"""The application driver on the GPU: `mechanics -opt options.toml` run on inputs in the reference's file formats
must write the reference's output files (src/system_driver.cpp:429-558) with the golden values
(test/data/*_stress.txt etc., 6 printed digits; our path reproduces 5)."""
import os
import subprocess

import numpy as np
import pytest

import app_inputs
import refcases

pytestmark = pytest.mark.gpu


def _run(tmp_path, **kw):
    opt = app_inputs.write_case(str(tmp_path), **kw)
    r = subprocess.run([app_inputs.mechanics_binary(), "-opt", opt], cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout[-2000:]
    return r.stdout


def test_voce_pa_option_file_reproduces_golden_stress_file(tmp_path):
    n = 12
    out = _run(tmp_path, nsteps=n)
    assert "Newton-steps/s" in out
    lines = open(os.path.join(str(tmp_path), "test_stress.txt")).read().strip().split("\n")
    assert len(lines) == n and all(len(l.split(" ")) == 6 for l in lines)   # Vector::Print(file, 6)
    s = np.loadtxt(os.path.join(str(tmp_path), "test_stress.txt"))
    gold = refcases.goldens()["voce_pa_stress"][:n]
    assert (np.abs(s - gold) / np.abs(gold[:, 2:3])).max() < 1.0e-5


def test_constant_strain_rate_ea_with_additional_averages(tmp_path):
    """voce_ea_cs-style run: velocity-gradient BCs, EA assembly, plastic-work / <F> / <D^p> files"""
    n = 6
    _run(tmp_path, nsteps=n, assembly="EA", bcs=app_inputs.BC_CS, extras=True)
    g = refcases.goldens()
    s = np.loadtxt(os.path.join(str(tmp_path), "test_stress.txt"))
    assert (np.abs(s - g["voce_ea_cs_stress"][:n]) / np.abs(g["voce_ea_cs_stress"][:n, 2:3])).max() < 1.0e-5
    F = np.loadtxt(os.path.join(str(tmp_path), "test_def_grad.txt"))
    plw = np.loadtxt(os.path.join(str(tmp_path), "test_pl_work.txt"))
    dp = np.loadtxt(os.path.join(str(tmp_path), "test_dp_tensor.txt"))
    assert F.shape == (n, 9) and plw.shape == (n,) and dp.shape == (n, 6)
    assert abs(F[-1, 8] - np.exp(1e-3 * g["custom_dt"][:n].sum())) < 1e-5     # constant true strain rate along z
    # the reference's own additional-average goldens of this case (6 printed digits)
    assert np.abs(F - g["voce_ea_cs_def_grad"][:n]).max() < 6e-6
    gp, gd = g["voce_ea_cs_pl_work"][:n], g["voce_ea_cs_dp_tensor"][:n]
    assert np.abs(plw - gp).max() / np.abs(gp).max() < 2e-4
    assert np.abs(dp - gd).max() / np.abs(gd).max() < 2e-4


def test_changing_mixed_bcs_fixed_time_steps(tmp_path):
    """cyclic_csm-style run with the reversal moved to step 4: matches the host layer driven directly"""
    from exaconstit_b200 import host
    _run(tmp_path, bcs=app_inputs.BC_CYCLIC_CSM, time="    [Time.Fixed]\n        dt = 0.1\n        t_final = 0.6")
    s = np.loadtxt(os.path.join(str(tmp_path), "test_stress.txt"))
    assert s.shape == (6, 6)
    inp, _ = refcases.case_inputs("voce_full_cyclic_csm")
    bcs = [(1,) + tuple(inp["bcs"][0][1:]), (4,) + tuple(inp["bcs"][1][1:])]
    sim = host.VoxelSim(inp["n"], inp["length"], 0, 0, inp["props"], 298.0, inp["grain_ids"], inp["quats"], nr=inp["nr"],
                        kr=inp["kr"])
    h = sim.run(np.full(6, 0.1), bcs)
    sim.close()
    ref = np.array([x["avg_stress"] for x in h])
    assert (np.abs(s - ref) / np.abs(ref[:, 2:3])).max() < 2e-5      # the file holds 6 significant digits
    assert s[2, 2] > 0 and s[5, 2] < s[3, 2]                          # unloading after the reversal
