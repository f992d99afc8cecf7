"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/orb_b200.h declares; wire structs have the reference's layout; no CPU fallback exists."""
import ctypes as C

import numpy as np
import pytest


def test_library_exports_declared_abi(orb):
    L = orb.lib()
    names = orb.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/orb_b200.h but not exported"
    assert L.orb_version() >= 100


def test_cell_layout_is_the_reference_wire_format(orb, oracle):
    # cell.h:9-17 measured offsets
    want = {"id": 0, "nLeafCells": 4, "prevCutAxis": 8, "cutAxis": 12, "foundCut": 16, "cutMarginLeft": 20,
            "cutMarginRight": 24, "lower": 28, "upper": 40}
    for dt in (orb.CELL_DTYPE, oracle.CELL_DTYPE):
        assert dt.itemsize == 52
        for k, off in want.items():
            assert dt.fields[k][1] == off
    assert C.sizeof(orb.BuildStats) % 8 == 0


def test_fails_loudly_without_gpu(orb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(orb.OrbError) as e:
        orb.Orb(1024, 8)
    assert "orb_create failed" in str(e.value)


def test_bad_arguments_are_rejected_before_touching_the_device(orb):
    with pytest.raises(orb.OrbError):
        orb.Orb(1024, 12)        # d must be a power of two (orbit.cpp:42)


def test_clustered_generators_are_deterministic_and_in_box(orb):
    for kind in ("gaussian", "plummer"):
        a = orb.generate_clustered(5000, kind)
        b = orb.generate_clustered(2000, kind, skip=3000)
        for u, v in zip(a, b):
            assert np.array_equal(u[3000:], v)
            assert u.min() >= -0.5 and u.max() <= 0.5
        # clustered: much tighter than uniform around 64 centres
        assert np.std(a[0][::64]) < 0.05
