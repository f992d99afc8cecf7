"""CPU test: oracle vs the unmodified reference binary, run here when /root/reference is present
(oracle/_ref/orbit_ref is rebuilt from the reference's sources by oracle/Makefile)."""
from pathlib import Path

import pytest

HAVE_REF = Path("/root/reference/src/orbit.cpp").exists()


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is only mounted in the build container")
@pytest.mark.parametrize("x,y,parts", [(13, 5, True), (15, 7, False), (12, 12, True), (18, 9, False)])
def test_oracle_trace_equals_live_reference(oracle, tmp_path, x, y, parts):
    oracle.build_oracle()
    assert oracle.REF_BIN.exists()
    ref_trace, ora_trace = tmp_path / "r.trace", tmp_path / "o.trace"
    oracle.run_reference(x, y, 0, trace_path=ref_trace, trace_particles=parts)
    xs, ys, zs = oracle.generate_uniform(1 << x)
    oracle.build(xs, ys, zs, 1 << y, ties=oracle.TIES_HOARE, trace_path=ora_trace, trace_particles=parts)
    assert ref_trace.read_bytes() == ora_trace.read_bytes()


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is only mounted in the build container")
@pytest.mark.parametrize("x,y,parts", [(18, 14, False), (16, 16, True)])
def test_oracle_trace_equals_reference_beyond_its_cell_cap(oracle, tmp_path, x, y, parts):
    """Beyond 2^12 leaf cells the reference's MAX_CELLS = 8096 arrays overflow (constants.h:11); oracle/_ref/orbit_ref_big is
    the same sources with that one constant lifted by the build recipe (oracle/Makefile: refbig).  The cap-free oracle
    must agree with it byte for byte there too - this pins the oracle for BASELINE configs C3..C5."""
    oracle.build_oracle()
    assert oracle.REF_BIN_BIG.exists()
    ref_trace, ora_trace = tmp_path / "r.trace", tmp_path / "o.trace"
    oracle.run_reference(x, y, 0, trace_path=ref_trace, trace_particles=parts, big=True)
    xs, ys, zs = oracle.generate_uniform(1 << x)
    oracle.build(xs, ys, zs, 1 << y, ties=oracle.TIES_HOARE, trace_path=ora_trace, trace_particles=parts)
    assert ref_trace.read_bytes() == ora_trace.read_bytes()
