"""Input side of the application driver (`mechanics -opt options.toml`, exaconstit_b200/csrc/options.hpp +
mechanics_main.cpp): the option-file reader and the text-input readers on files in the reference's format, without a
GPU (`--check` parses and reads everything, then stops before creating the simulation)."""
import os
import subprocess

import pytest

import app_inputs


def _check(tmp_path, **kw):
    import __graft_entry__ as ge
    ge.build()
    opt = app_inputs.write_case(str(tmp_path), **kw)
    return subprocess.run([app_inputs.mechanics_binary(), "-opt", opt, "--check"], cwd=str(tmp_path), capture_output=True,
                          text=True)


def test_option_file_and_inputs_are_read(tmp_path):
    r = _check(tmp_path, nsteps=7)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "check: xtal 0 slip 0 assembly 1 integ 0 nl_solver 0 nprops 17 nstate 24 ngrains 500 temp 298" in out
    assert "nr 5e-05 5e-10 25 krylov 1e-07 1e-27 1000 nsteps 7 auto 0 cust 1" in out
    assert "bc step 1 ids 1 2 3 4 comps 3 1 2 3 vals 0 0 0 0 0 0 0 0 0 0 0 0.001 vgrad\n" in out
    assert "mesh: 10 x 10 x 10 hexes, 500 grains" in out
    assert "grain checksum 63000" in out            # 8 children per coarse element inherit its grain id
    assert "files test_stress.txt test_pl_work.txt test_def_grad.txt test_dp_tensor.txt" in out


def test_changing_mixed_bcs_and_auto_time(tmp_path):
    r = _check(tmp_path, bcs=app_inputs.BC_CYCLIC_CSM, assembly="EA",
               time="    [Time.Fixed]\n        dt = 0.1\n        t_final = 0.6")
    assert r.returncode == 0, r.stderr
    assert "bc step 1 ids 1 2 3 4 comps 3 -1 -2 -3 vals 0 0 0 0 0 0 0 0 0 0 0 0.001 vgrad 0 0 0 0 0 0 0 0 0.001" in r.stdout
    assert "bc step 4 ids 1 2 3 4 comps 3 -1 -2 -3 vals 0 0 0 0 0 0 0 0 0 0 0 -0.001 vgrad 0 0 0 0 0 0 0 0 -0.001" in r.stdout
    assert "assembly 2" in r.stdout and "nsteps 6 auto 0 cust 0 dt 0.1" in r.stdout
    r = _check(tmp_path, xtal="fcc", slip="mtsdd", props_key="props_cp_mts_in625", assembly="FULL",
               time="    [Time.Auto]\n        dt_start = 0.1\n        dt_min = 0.05\n        dt_scale = 0.333333\n"
                    "        t_final = 10.0\n        auto_dt_file = \"auto_dt_out.txt\"")
    assert r.returncode == 0, r.stderr
    assert "nsteps 200 auto 1 cust 0 dt 0.1 dt_min 0.05 dt_scale 0.333333 t_final 10" in r.stdout
    assert "running the matrix-free PA operator" in r.stdout
    # parity notice: this regime of the KMBalD kinetics is not pinned to the reference
    assert "warning: KMBalD kinetics with p = 0.8, q = 1.4" in r.stdout and "not pinned to the reference" in r.stdout
    r = _check(tmp_path, xtal="bcc", slip="mtsdd", props_key="props_cp_mts")
    assert r.returncode == 0 and "warning" not in r.stdout          # p = q = 1: pinned by mtsdd_bcc / mtsdd_full


@pytest.mark.parametrize("kw,msg", [
    (dict(slip="powervoce", xtal="hcp"), "can not be PowerVoce for HCP"),
    (dict(props_key="props_cp_vocenl"), "Properties.Matl_Props.num_props needs 17"),
    (dict(assembly="MF"), "Solvers.assembly was not provided a valid type."),
    (dict(bcs="    essential_ids = [1, 2]\n    essential_comps = [3, 1]"), "BCs.essential_vals was not provided any values"),
    (dict(bcs=app_inputs.BC_CYCLIC_CSM, time="    [Time.Auto]\n        dt_start = 0.1"), "not compatible with changing boundary"),
])
def test_bad_options_abort_like_the_reference(tmp_path, kw, msg):
    r = _check(tmp_path, **kw)
    assert r.returncode != 0
    assert msg in r.stderr, r.stderr


def test_missing_input_file_aborts(tmp_path):
    r = _check(tmp_path)
    assert r.returncode == 0
    os.remove(os.path.join(str(tmp_path), "grains.txt"))
    r = subprocess.run([app_inputs.mechanics_binary(), "-opt", "options.toml", "--check"], cwd=str(tmp_path),
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Cannot open grain map file" in r.stderr


def test_option_reader_toml_subset_edge_cases(tmp_path):
    """constructs the reference's files use that are easy to get wrong: '#' inside strings, comments inside multi-line
    arrays, trailing commas, integers where floats are expected, underscores in numbers, exponents without a dot"""
    import __graft_entry__ as ge
    ge.build()
    opt = app_inputs.write_case(str(tmp_path))
    txt = open(opt).read()
    txt = txt.replace('avg_stress_fname = "test_stress.txt"', 'avg_stress_fname = "stress#1.txt"   # a comment')
    txt = txt.replace("iter = 25", "iter = 2_5")
    txt = txt.replace("rel_tol = 5e-05", "rel_tol = 5E-5")
    txt = txt.replace("length = [1.0, 1.0, 1.0]", "length = [1, 1,   # integers, and a comment inside the array\n   1,]")
    txt = txt.replace("temperature = 298", "temperature = 3.0e2")
    open(opt, "w").write(txt)
    r = subprocess.run([app_inputs.mechanics_binary(), "-opt", opt, "--check"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "files stress#1.txt" in r.stdout
    assert "nr 5e-05 5e-10 25" in r.stdout and "temp 300" in r.stdout
    # unsupported TOML constructs are rejected loudly, not mis-parsed
    bad = txt.replace('[Mesh.Auto]', '[Mesh.Auto]\n        point = { x = 1, y = 2 }')
    open(opt, "w").write(bad)
    r = subprocess.run([app_inputs.mechanics_binary(), "-opt", opt, "--check"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode != 0 and "inline tables are not supported" in r.stderr
    open(opt, "w").write(txt.replace("[Solvers]", "[[Solvers]]"))
    r = subprocess.run([app_inputs.mechanics_binary(), "-opt", opt, "--check"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode != 0 and "bad table header" in r.stderr
