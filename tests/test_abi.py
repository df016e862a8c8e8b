"""The C-ABI library loads without a GPU and exports every symbol include/demo_b200.h declares
(no compute calls here)."""
from __future__ import annotations

import os
import re

import pytest

from tests.helpers import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "demo_b200.h")).read()
    return sorted(set(re.findall(r"DEMO_API\s+[\w\s\*]+?\b(demo_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from demo2_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "symbol %s missing from libdemo_b200.so" % n
        assert n in _lib.SIGNATURES, "symbol %s has no ctypes signature" % n
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.demo_version() >= 100
    # host-only size queries work without a device
    assert lib.demo_sqdist_workspace_bytes(836, 836, 1536, 0) > 2 * 836 * 1536 * 4
    assert lib.demo_plan_bytes(836, 836) > 0
    assert lib.demo_eval_workspace_bytes(836, 836, 1536, 10000) > 0


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from demo2_b200 import _lib, metrics
    with pytest.raises(_lib.DemoError):
        metrics.euclidean_distance(np.zeros((2, 4), np.float32), np.zeros((3, 4), np.float32))
    with pytest.raises(_lib.DemoError):
        metrics.eval_func(np.zeros((2, 3), np.float32), [0, 1], [0, 1, 2], [0, 0], [1, 1, 1])
