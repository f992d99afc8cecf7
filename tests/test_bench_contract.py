"""The JSON line contract of bench.py: the committed B200 line of the latest profile and a live run of the reference
arm (CPU, runs here) carry every key the driver and the judge read, with consistent metric / unit / workload."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def latest_gpu_line():
    files = sorted(ROOT.glob("profiles/r*_bench.json"))
    assert files, "no committed bench line under profiles/"
    return json.loads(files[-1].read_text().strip().splitlines()[-1]), files[-1].name


def test_committed_gpu_line_has_the_contract_keys():
    d, name = latest_gpu_line()
    assert BASE <= set(d), (name, BASE - set(d))
    assert d["metric"] == "orb_particle_passes_per_s" and d["unit"] == "particle-passes/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "l2" in d["config"]
    assert d["config"]["workload"].startswith("c3:")          # N = 1 default: the largest single-GPU config of BASELINE.json
    assert d["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and d["clocks"]["samples"] > 0
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                      # host buffers in and out cannot beat the resident number
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac"] < 1.0
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("reference", "port")
    # value = particle passes of the job / build time
    assert d["value"] == pytest.approx(d["particle_passes_per_build"] / (d["ms_per_step"] * 1e-3), rel=1e-9)
    # parity is part of the line: the primary leg and every extra leg were compared with the CPU oracle's digests
    assert d["parity"]["status"] == "pass" and d["parity_all_legs"] == "pass"
    assert {"c2", "c4g", "c4p"} <= set(d["legs"])
    for leg in d["legs"].values():
        assert leg["parity"]["status"] == "pass" and leg["clocks"]["samples"] > 0 and leg["roofline"]["frac"] > 0
    assert d["reference_exact_mode"]["parity"]["status"] == "pass"


def test_reference_arm_line():
    if not (ROOT / "oracle" / "_ref" / "orbit_ref").exists() and not (ROOT / "oracle" / "orb_oracle").exists():
        pytest.skip("neither the reference binary nor the oracle is built")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE <= set(d)
    g, _ = latest_gpu_line()
    assert (d["metric"], d["unit"], d["higher_is_better"], d["scaling"]) == (g["metric"], g["unit"], g["higher_is_better"], g["scaling"])
    assert d["config"] == g["config"]                    # both arms name the same workload, key for key
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1 and d["gpu_launches"] == 0
    # Same kind of numerator as the GPU arm (particles in unfound cells summed over bisection iterations).  Not the same
    # number when the reference runs on several threads: its generator state is shared by its threads (init.cu:11-25,
    # unsynchronised), so a multi-threaded run draws a different - partly duplicated - particle set every time.
    passes = d["value"] * d["ms_per_step"] * 1e-3
    assert 0.5 * g["particle_passes_per_build"] < passes < 2.0 * g["particle_passes_per_build"]


def test_reference_arm_same_workload_at_every_gpu_count():
    """N > 1 shards the SAME job (strong scaling), so the reference arm builds the same workload for every --gpus."""
    sys.path.insert(0, str(ROOT))
    import bench

    c1 = bench.config_dict(bench.config_of(bench.default_legs(1)[0], 1), 1)
    for n in (2, 4, 8):
        cn = bench.config_dict(bench.config_of(bench.default_legs(n)[0], n), n)
        assert cn["workload"] == c1["workload"] and cn["particles_total"] == c1["particles_total"]
    assert bench.default_legs(8)[1] == "c5"


def test_cpu_baseline_helper_runs_here():
    """The cpu_baseline leg of the GPU arm is plain CPU work: run it on small workloads - below the reference's cell cap
    (the unmodified binary) and above it (the build with MAX_CELLS lifted)."""
    sys.path.insert(0, str(ROOT))
    import bench

    for y, binary in ((6, "orbit_ref "), (13, "orbit_ref_big")):
        cfg = dict(bench.config_of("c1", 1), x=18, y=y)
        c = bench.cpu_baseline_beside(cfg, budget_s=2.0)
        assert c["value"] and c["value"] > 0 and c["cores"] >= 1 and c["kind"] in ("reference", "port")
        assert "median of" in c["sample"] and c["build_ms"] > 0
        if (ROOT / "oracle" / "_ref" / binary.strip()).exists():
            assert c["kind"] == "reference" and c["binary"].startswith(binary)
