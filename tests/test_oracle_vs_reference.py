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
