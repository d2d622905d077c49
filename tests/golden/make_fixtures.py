"""Collects the reference's own regression inputs and golden outputs (test/data/*) into one
small fixture so the parity tests never need /root/reference at run time.

Run here (in the build container):  python tests/golden/make_fixtures.py
Source files (read-only): /root/reference/test/data/{props_cp_*.txt, voce_quats.ori, custom_dt.txt,
*_stress.txt, voce_ea{,_cs}_{pl_work,dp_tensor,def_grad}.txt}; they are checked by the reference's
test/test_mechanics.py:11-31,49-54,114-117.
"""
import os

import numpy as np

SRC = "/root/reference/test/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exaconstit_goldens.npz")

d = {}
for name in ["props_cp_voce", "props_cp_vocenl", "props_cp_mts", "props_cp_mts_in625", "custom_dt"]:
    d[name] = np.loadtxt(os.path.join(SRC, name + ".txt")).ravel()
d["voce_quats"] = np.loadtxt(os.path.join(SRC, "voce_quats.ori"))
d["grains"] = np.loadtxt(os.path.join(SRC, "grains.txt")).astype(np.int32)
for name in ["voce_pa", "voce_full", "voce_ea", "voce_bcc", "mtsdd_bcc", "mtsdd_full", "mtsdd_full_auto",
             "voce_full_cyclic", "voce_full_cyclic_cs", "voce_full_cyclic_csm", "voce_ea_cs"]:
    d[name + "_stress"] = np.loadtxt(os.path.join(SRC, name + "_stress.txt"))
for name in ["voce_ea_pl_work", "voce_ea_dp_tensor", "voce_ea_def_grad", "voce_ea_cs_pl_work", "voce_ea_cs_dp_tensor",
             "voce_ea_cs_def_grad"]:
    d[name] = np.loadtxt(os.path.join(SRC, name + ".txt"))
np.savez_compressed(OUT, **d)
print("wrote", OUT, os.path.getsize(OUT), "bytes")
