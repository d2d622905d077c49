"""The reference's regression cases (test/data/*.toml) restated as oracle/product inputs."""
import os

import numpy as np

_G = None


def goldens():
    global _G
    if _G is None:
        _G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exaconstit_goldens.npz"))
    return _G


def refined_grain_ids(n_coarse=5):
    """test/data/voce_pa.toml:122,131-137: 5x5x5 auto mesh, ref_ser=1 -> 10^3 hexes; children inherit
    the parent's grain attribute (src/mechanics_driver.cpp:276-281,308-310); grains.txt is x-fastest."""
    g = goldens()["grains"].reshape(n_coarse, n_coarse, n_coarse)  # [z][y][x]
    fine = np.repeat(np.repeat(np.repeat(g, 2, axis=0), 2, axis=1), 2, axis=2)
    return fine.ravel().astype(np.int32)


# uniaxial symmetric BCs of test/data/voce_pa.toml:40-51
def uniaxial_bcs(rate=0.001):
    return [(1, [1, 2, 3, 4], [3, 1, 2, 3], [[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, rate]])]


CASES = {
    # name: (xtal, kin, props key, assembly, nr, kr, golden key)
    "voce_pa": (0, 0, "props_cp_voce", 0, (5e-5, 5e-10, 25), (1e-7, 1e-27, 1000), "voce_pa_stress"),
    "voce_ea": (0, 0, "props_cp_voce", 1, (5e-5, 5e-10, 25), (1e-7, 1e-27, 1000), "voce_ea_stress"),
    "voce_full": (0, 0, "props_cp_voce", 0, (5e-5, 5e-10, 25), (1e-7, 1e-27, 1000), "voce_full_stress"),
    "voce_nl_full": (0, 1, "props_cp_vocenl", 0, (5e-5, 5e-10, 25), (1e-7, 1e-27, 1000), "voce_full_stress"),
    "voce_bcc": (1, 0, "props_cp_voce", 0, (5e-5, 5e-10, 25), (1e-7, 1e-27, 1000), "voce_bcc_stress"),
    "mtsdd_bcc": (1, 2, "props_cp_mts", 0, (1e-5, 1e-12, 25), (1e-7, 1e-27, 250), "mtsdd_bcc_stress"),
    "mtsdd_full": (0, 2, "props_cp_mts", 0, (1e-5, 1e-12, 25), (1e-7, 1e-27, 250), "mtsdd_full_stress"),
}


def cyclic_bcs(rate=0.001):
    """test/data/voce_full_cyclic.toml:45-74: the z_max velocity flips sign at steps 11, 31, 51, 71."""
    out = []
    for step, sgn in zip([1, 11, 31, 51, 71], [1, -1, 1, -1, 1]):
        out.append((step, [1, 2, 3, 4], [3, 1, 2, 3], [[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, sgn * rate]]))
    return out


def cyclic_cs_bcs(rate=0.001, mixed=False):
    """test/data/voce_full_cyclic_cs.toml:45-79 (all four attributes velocity-gradient driven, comps -3,-1,-2,-3)
    and voce_full_cyclic_csm.toml:65-91 (z_min as a plain velocity BC, the rest velocity-gradient driven)."""
    out = []
    for step, sgn in zip([1, 11, 31, 51, 71], [1, -1, 1, -1, 1]):
        L = [[0, 0, 0], [0, 0, 0], [0, 0, sgn * rate]]
        comps = [3, -1, -2, -3] if mixed else [-3, -1, -2, -3]
        out.append((step, [1, 2, 3, 4], comps, [[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, sgn * rate]], L))
    return out


# Time.Auto of test/data/mtsdd_full_auto.toml:102-120
MTSDD_AUTO_TIME = dict(dt_start=0.1, dt_min=0.05, dt_scale=0.333333, t_final=10.0)


def case_inputs(name):
    g = goldens()
    base = dict(n=(10, 10, 10), length=(1.0, 1.0, 1.0), temp_k=298.0, grain_ids=refined_grain_ids(), quats=g["voce_quats"])
    if name in ("voce_full_cyclic_cs", "voce_full_cyclic_csm"):
        return dict(base, xtal=0, kin=0, props=g["props_cp_voce"], dts=np.full(70, 0.1),
                    bcs=cyclic_cs_bcs(mixed=name.endswith("csm")), assembly=0, nr=(5e-5, 5e-10, 25),
                    kr=(1e-7, 1e-27, 1000)), g[name + "_stress"]
    if name == "voce_ea_cs":
        # voce_ea.toml with constant_strain_rate: comps -3,-1,-2,-3 and essential_vel_grad diag(0,0,1e-3)
        bcs = [(1, [1, 2, 3, 4], [-3, -1, -2, -3], np.zeros((4, 3)), [[0, 0, 0], [0, 0, 0], [0, 0, 0.001]])]
        return dict(base, xtal=0, kin=0, props=g["props_cp_voce"], dts=g["custom_dt"], bcs=bcs, assembly=1,
                    nr=(5e-5, 5e-10, 25), kr=(1e-7, 1e-27, 1000)), g["voce_ea_cs_stress"]
    if name == "mtsdd_full_auto":
        # compression at -1e-3, IN625 property set, automatic time stepping; FULL assembly run as PA here
        return dict(base, xtal=0, kin=2, props=g["props_cp_mts_in625"], dts=np.zeros(0), bcs=uniaxial_bcs(-0.001),
                    assembly=0, nr=(1e-5, 1e-12, 25), kr=(1e-7, 1e-27, 250), auto_time=dict(MTSDD_AUTO_TIME)), \
            g["mtsdd_full_auto_stress"]
    if name == "voce_full_cyclic":
        # Time.Fixed dt = 0.1, t_final = 7.0 (voce_full_cyclic.toml:100-103); FULL assembly run as PA here
        return dict(n=(10, 10, 10), length=(1.0, 1.0, 1.0), xtal=0, kin=0, props=g["props_cp_voce"], temp_k=298.0,
                    grain_ids=refined_grain_ids(), quats=g["voce_quats"], dts=np.full(70, 0.1), bcs=cyclic_bcs(),
                    assembly=0, nr=(5e-5, 5e-10, 25), kr=(1e-7, 1e-27, 1000)), g["voce_full_cyclic_stress"]
    xtal, kin, pk, assembly, nr, kr, gk = CASES[name]
    return dict(n=(10, 10, 10), length=(1.0, 1.0, 1.0), xtal=xtal, kin=kin, props=g[pk], temp_k=298.0,
                grain_ids=refined_grain_ids(), quats=g["voce_quats"], dts=g["custom_dt"], bcs=uniaxial_bcs(),
                assembly=assembly, nr=nr, kr=kr), g[gk]


def golden_rel_err(sim_stress, gold):
    """Error measure in units of the 6th significant digit of the golden's loaded component."""
    s = np.asarray(sim_stress)[: gold.shape[0]]
    scale = np.abs(gold[:, 2:3])
    return np.abs(s - gold) / scale


def hcp_props():
    """Synthetic HCP (Ti-like) KMBalD property vector -- the reference ships none, so this set
    is ours and HCP parity is oracle-vs-GPU only.  Order: rho0, cv, tol | c11 c12 c13 c33 c44 | mu_ref, T_ref |
    c_1 x4 families (basal, prismatic, pyramidal<a>, pyramidal<c+a>) | tau_a, p, q | gam_wo, gam_ro, wrD |
    g_0 x4 | s x4 | k1, k2_0, n^-1, gamma_o, rho_dd_ref | c/a | Gruneisen, e_ref."""
    cv = 2.5e-3
    return np.array([4.5e-6, cv, 1.0e-10,
                     162.4, 92.0, 69.0, 180.7, 46.7,
                     44.0, 300.0,
                     1944.1, 1944.1, 2100.0, 2400.0,
                     4.0e-4, 1.0, 1.0,
                     1.0, 1.0, 3.0e-2,
                     8.0e-3, 6.0e-3, 1.2e-2, 2.0e-2,
                     1.0e-1, 1.0e-1, 1.2e-1, 1.5e-1,
                     3.0e-4, 5.0e-5, 0.1, 1.0e-2, 9.0e-4,
                     1.587,
                     0.0, -cv * 300.0])
