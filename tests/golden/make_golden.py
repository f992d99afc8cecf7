"""Regenerates the golden traces in this directory by running the UNMODIFIED reference
(oracle/_ref/orbit_ref, built by oracle/Makefile from /root/reference/src) in its CPU-only mode
with the service-call tap enabled (oracle/ref_shim/ref_tap.cpp).  Needs /root/reference.

    python tests/golden/make_golden.py

Each file is the gzip of the ORBTRACE stream: per level the Cell array + particle counts
(ServiceCount), every bisection iteration's Cell array + count-left vector, and after the partition
the final cells, child ranges and per-child particle hashes; "p" variants also carry the particle
columns after Init and after every partition.
"""
import gzip
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "tests"))

# (x, y, dump particles)
CONFIGS = [(10, 3, True), (12, 4, True), (11, 11, True), (14, 6, False), (16, 8, False), (20, 10, False)]


def main():
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref"], check=True)
    ref = ROOT / "oracle" / "_ref" / "orbit_ref"
    for x, y, parts in CONFIGS:
        tmp = HERE / f"_tmp_{x}_{y}.trace"
        env = dict(os.environ, ORB_MDL_THREADS="1", ORB_REF_TRACE=str(tmp))
        if parts:
            env["ORB_REF_TRACE_PARTICLES"] = "1"
        else:
            env.pop("ORB_REF_TRACE_PARTICLES", None)
        subprocess.run([str(ref), str(x), str(y), "0"], env=env, check=True, capture_output=True)
        out = HERE / f"ref_{x}_{y}{'p' if parts else ''}.trace.gz"
        with open(tmp, "rb") as f, gzip.GzipFile(out, "wb", mtime=0) as g:
            g.write(f.read())
        tmp.unlink()
        print(out.name, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
