"""CPU-side checks of the product library: it loads, exports every symbol include/exab200.h declares,
and refuses to run without a Blackwell GPU instead of falling back to anything."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from exaconstit_b200 import capi
    lib = capi.lib()
    header = open(os.path.join(ROOT, "include", "exab200.h")).read()
    declared = set(re.findall(r"\b(exab200_[a-z_0-9]+)\s*\(", header))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_host_library_exports_every_declared_symbol():
    import ctypes as C
    import __graft_entry__ as ge
    ge.build()
    from exaconstit_b200 import host
    lib = host.lib()
    header = open(os.path.join(ROOT, "include", "exahost.h")).read()
    declared = set(re.findall(r"\b(exahost_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 18
    for s in declared:
        assert hasattr(lib, s), s
    assert os.access(os.path.join(ROOT, "exaconstit_b200", "lib", "mechanics"), os.X_OK)


def test_bad_config_is_rejected_before_touching_the_gpu():
    import numpy as np
    from exaconstit_b200 import capi
    with pytest.raises(capi.Exab200Error):
        capi.Context(capi.FCC, capi.POWERVOCE, np.zeros(5), 298.0, 1, 8)  # wrong number of properties
    with pytest.raises(capi.Exab200Error):
        capi.Context(capi.HCP, capi.POWERVOCE, np.zeros(17), 298.0, 1, 8)  # Voce is not defined for HCP


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import refcases
    from exaconstit_b200 import capi
    with pytest.raises(capi.Exab200Error):
        capi.Context(capi.FCC, capi.POWERVOCE, refcases.goldens()["props_cp_voce"], 298.0, 1, 8,
                     np.arange(8, dtype=np.int32))
