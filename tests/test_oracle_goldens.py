"""The oracle pinned against the reference's own golden outputs (test/data/*_stress.txt, checked by
test/test_mechanics.py:11-31,49-54).  The goldens print 6 significant digits.

Status ('Oracle and parity status'), with the volume-convected lattice strain found in round 2
(oracle/ecmech_port.hpp Options::vol_convect):
  * FCC/BCC Voce and Voce-NL (voce_pa, voce_ea, voce_full, voce_nl_full, voce_bcc, voce_ea_cs), BCC KMBalD (mtsdd_bcc)
    and the load reversals (voce_full_cyclic*): every printed axial stress within HALF a unit of its 6th digit
    (<= 1e-6 of the peak, 4e-6 for mtsdd_bcc), every other component <= 1e-8 / 9e-8: the printed values are identical,
    and voce_pa / voce_full / voce_nl_full / voce_bcc / mtsdd_bcc PASS the reference's own criterion
    (test/test_mechanics.py:11-31: mean absolute row difference of the 6-digit prints <= 1e-10) -- see
    test_reference_pass_criterion.  The cyclic and EA cases miss it by the solver-noise columns only (1.2e-10 .. 3.9e-10).
  * FCC KMBalD (mtsdd_full): axial 1.0e-6 (0.83 units), shear components up to two units of their last digit (2.8e-8).
"""
import math

import numpy as np
import pytest

import refcases


def _run(orc, name, nsteps):
    inp, gold = refcases.case_inputs(name)
    inp["dts"] = inp["dts"][:nsteps]
    r = orc.sim_run(**inp)
    assert r["rc"] == 0
    assert r["stats"]["failed_points"] == 0
    return r, gold[:nsteps]


def _ulp6(v):
    """one unit of the 6th significant digit of a printed value"""
    return 10.0 ** (math.floor(math.log10(abs(v))) - 5)


def golden_errors(s, gold):
    """(largest axial error in units of the golden's last printed digit, largest axial error / peak axial stress,
    largest shear-component error / peak, largest lateral-component error / peak)"""
    peak = np.abs(gold[:, 2]).max()
    zz_ulp = max(abs(s[i, 2] - gold[i, 2]) / _ulp6(gold[i, 2]) for i in range(len(gold)))
    return (zz_ulp, np.abs(s[:, 2] - gold[:, 2]).max() / peak, np.abs(s[:, 3:] - gold[:, 3:]).max() / peak,
            np.abs(s[:, :2] - gold[:, :2]).max() / peak)


# per case: steps run in the default suite, and the bounds (axial in last-digit units or None, axial / peak, shear / peak)
VOCE = dict(zz_ulp=0.6, zz=1.2e-6, shear=2.0e-8)
KMBALD = dict(zz_ulp=1.0, zz=5.0e-6, shear=6.0e-8)
CASES = [("voce_pa", 40, VOCE), ("voce_ea", 8, VOCE), ("voce_full", 8, VOCE), ("voce_nl_full", 6, VOCE), ("voce_bcc", 40, VOCE),
         ("voce_ea_cs", 40, VOCE), ("mtsdd_bcc", 40, KMBALD), ("mtsdd_full", 40, KMBALD)]


def reference_criterion(s, gold):
    """check_stress of test/test_mechanics.py:11-31 applied to our rows printed the way the reference prints them
    (Vector::Print with the stream's default 6 significant digits): sum of absolute differences over all six columns,
    averaged over the rows; the reference passes a case when this is <= 1e-10."""
    err = 0.0
    for a, t in zip(gold, s):
        for x, y in zip(a, t):
            err += abs(float(x) - float("%.6g" % y))
    return err / len(gold)


REF_PASS = ("voce_pa", "voce_bcc", "mtsdd_bcc")   # whole histories in the default suite


def _check(r, gold, lim):
    zz_ulp, zz, shear, lat = golden_errors(r["stress"], gold)
    if lim["zz_ulp"] is not None:
        assert zz_ulp <= lim["zz_ulp"], zz_ulp
    assert zz <= lim["zz"], zz
    assert shear <= lim["shear"], shear
    assert lat <= 1.5e-8, lat


@pytest.mark.parametrize("name,nsteps,lim", CASES, ids=[c[0] for c in CASES])
def test_monotonic_goldens(orc, name, nsteps, lim):
    r, gold = _run(orc, name, nsteps)
    _check(r, gold, lim)
    if name in REF_PASS:
        # THE REFERENCE'S OWN PASS CRITERION on the whole history (measured 2.1e-11 / 1.9e-11 / 1.1e-11)
        assert nsteps == len(gold) == 40
        assert reference_criterion(r["stress"], gold) <= 1.0e-10
    if name == "voce_pa":
        # iteration counts are part of the parity record (reference = identity-preconditioned CG)
        assert r["stats"]["newton_iters"] < 120


@pytest.mark.slow
@pytest.mark.parametrize("name,lim", [("voce_ea", VOCE), ("voce_full", VOCE), ("voce_nl_full", VOCE)], ids=["voce_ea", "voce_full", "voce_nl_full"])
def test_monotonic_goldens_whole_history(orc, name, lim):
    """the remaining 40-row histories (same material point physics as voce_pa through other operator paths)"""
    r, gold = _run(orc, name, 40)
    _check(r, gold, lim)
    if name != "voce_ea":   # voce_ea: 1.2e-10, the golden's solver-noise columns
        assert reference_criterion(r["stress"], gold) <= 1.0e-10


def test_printed_digits_of_voce_pa(orc):
    """The reference's criterion is identical 6-digit prints (test/test_mechanics.py:11-31).  Count how far we are:
    no printed value may be off by more than one unit of its last digit (values below 1e-7 of the stress scale,
    i.e. the lateral components and the elastic-regime shear, are noise in the golden itself)."""
    r, gold = _run(orc, "voce_pa", 40)
    s = r["stress"]
    off, worst = 0, 0.0
    for i in range(40):
        for c in range(2, 6):
            if abs(gold[i, c]) < 1e-4:   # print resolution finer than 1e-9 GPa = 2e-8 of the axial stress
                continue
            d = abs(float("%.6g" % s[i, c]) - gold[i, c]) / _ulp6(gold[i, c])
            off += d > 0.5
            worst = max(worst, d)
    assert off == 0 and worst <= 0.5, (off, worst)   # all 152 printed values identical (about 40 differed in round 1)


def test_plastic_work_and_dp_goldens(orc):
    """voce_ea also pins the volume-integrated plastic work and the average D^p
    (test/test_mechanics.py:114-117)."""
    g = refcases.goldens()
    r, gold = _run(orc, "voce_ea", 8)
    plw = r["extra"][:8, 0]
    gp = g["voce_ea_pl_work"][:8]
    assert (np.abs(plw - gp)[1:] / np.abs(gp)[1:]).max() < 3e-5      # each row relative to itself (row 1 is zero work)
    dp = r["extra"][:8, 1:7]
    gd = g["voce_ea_dp_tensor"][:8]
    assert np.abs(dp - gd).max() / np.abs(gd).max() < 1.5e-5
    # volume-averaged deformation gradient (voce_ea_def_grad.txt, values ~1 printed with 6 digits)
    assert np.abs(r["extra"][:8, 7:16] - g["voce_ea_def_grad"][:8]).max() < 6e-6


def test_additional_averages_constant_strain_rate(orc):
    """voce_ea_cs_{pl_work,dp_tensor,def_grad}.txt: the additional averages under velocity-gradient BCs"""
    g = refcases.goldens()
    r, gold = _run(orc, "voce_ea_cs", 10)
    gp, gd, gF = g["voce_ea_cs_pl_work"][:10], g["voce_ea_cs_dp_tensor"][:10], g["voce_ea_cs_def_grad"][:10]
    assert (np.abs(r["extra"][:10, 0] - gp)[1:] / np.abs(gp)[1:]).max() < 3e-5
    assert np.abs(r["extra"][:10, 1:7] - gd).max() / np.abs(gd).max() < 1.5e-5
    assert np.abs(r["extra"][:10, 7:16] - gF).max() < 6e-6


def test_cyclic_reversal(orc):
    """voce_full_cyclic_stress.txt across the first load reversal (BC change at step 11 -> SolveInit)."""
    inp, gold = refcases.case_inputs("voce_full_cyclic")
    inp["dts"] = inp["dts"][:14]
    r = orc.sim_run(**inp)
    assert r["rc"] == 0
    err = np.abs(r["stress"][:, 2] - gold[:14, 2]).max() / np.abs(gold[:14, 2]).max()
    assert err < 1.5e-6, err          # 1.5e-5 before the volume-convected strain
    assert np.abs(r["stress"][:, 3:] - gold[:14, 3:]).max() / np.abs(gold[:14, 2]).max() < 3e-8


def test_constant_strain_rate_bcs(orc):
    """voce_ea_cs_stress.txt: velocity-gradient ("constant strain rate") boundary conditions, EA assembly."""
    r, gold = _run(orc, "voce_ea_cs", 40)   # the whole history
    _check(r, gold, VOCE)


@pytest.mark.parametrize("name,nsteps", [("voce_full_cyclic_cs", 13), ("voce_full_cyclic_csm", 70)])
def test_cyclic_constant_strain_rate(orc, name, nsteps):
    """velocity-gradient BCs (all faces / mixed with a velocity BC): across the first load reversal, and the whole
    70-step history with its four reversals (error relative to the peak stress: the golden crosses zero)"""
    inp, gold = refcases.case_inputs(name)
    inp["dts"] = inp["dts"][:nsteps]
    r = orc.sim_run(**inp)
    assert r["rc"] == 0
    err = np.abs(r["stress"][:, 2] - gold[:nsteps, 2]).max() / np.abs(gold[:nsteps, 2]).max()
    assert err < 2.5e-5, err
    assert np.abs(r["stress"][:, 3:] - gold[:nsteps, 3:]).max() / np.abs(gold[:nsteps, 2]).max() < 5e-7


def test_auto_time_stepping(orc):
    """Time.Auto (src/system_driver.cpp:225-274, src/mechanics_driver.cpp:845-848): dt_next = dt * (NR.iter * dt_scale) /
    newton_iterations, floored at dt_min and clipped to t_final.  mtsdd_full_auto_stress.txt (compression, IN625 KMBalD
    set) pins the first row only: its later rows were produced with step sizes that encode the Newton iteration counts
    of the reference's FULL-assembly / BoomerAMG solve (24 iterations in step 2, then 6, 6, 15, ... as recovered from
    the elastic stress increments), which a PA / CG path does not reproduce, and in that parameter regime
    (p = 0.8, q = 1.4, 260 MPa Peierls stress) our restated KMBalD kinetics are UNPINNED: a fixed-dt run reaches
    -724.7 MPa at t = 10 where the golden's last row has -773.1 MPa ('Oracle and parity status')."""
    inp, gold = refcases.case_inputs("mtsdd_full_auto")
    at = inp["auto_time"]
    at["t_final"] = 1.2
    r = orc.sim_run(**inp)
    assert r["rc"] == 0
    dts, nit = r["dts"], r["iters"][:, 0]
    assert abs(dts.sum() - at["t_final"]) < 1e-12 and dts[0] == at["dt_start"]
    t = 0.0
    dt_class = at["dt_start"]
    for k in range(dts.size):
        assert abs(dts[k] - min(dt_class, at["t_final"] - t)) < 1e-14
        t += dts[k]
        dt_class = max(at["dt_min"], dts[k] * (inp["nr"][2] * at["dt_scale"]) / nit[k])
    err = np.abs(r["stress"][0] - gold[0]) / abs(gold[0, 2])
    assert err.max() < 1.5e-5, err.max()
