"""bench.py's reference arm (the CPU side of the driver's two-arm comparison) runs without a GPU: its JSON line must
carry the contract's keys.  The B200 arm needs a device; its line is checked for the same keys on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _line(out):
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1 and lines[0].startswith("{"), out[-2000:]   # ONE JSON line, nothing else on stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line(built):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--ref-spp", "1", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _line(r.stdout)
    if "unavailable" in d:
        pytest.skip("neither oracle/_ref nor the oracle port is built here")
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["impl"] == "reference" and d["metric"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_do_nothing(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


@pytest.mark.gpu
def test_b200_arm_prints_the_contract_line(built):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1", "--spp", "8", "--ref-spp", "1",
                        "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _line(r.stdout)
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches"} <= set(d), sorted((BASE_KEYS | {"roofline", "clocks", "gpu_launches"}) - set(d))
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["d2h_bytes_per_step"] == 512 * 512 * 20 and d["e2e"]["h2d_bytes_per_step"] > 0
    rf = d["roofline"]
    # the Cornell scene's BVH is L2-resident: rated against the measured L2 gather peak
    assert rf["bound"] == "l2" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["kernel"] == "TraceClosestKernel" and rf["frac_of_hbm_peak"] > 0
    fam = d["roofline_families"]
    assert set(fam) == {"trace_closest", "trace_any", "sss_walk", "shade"}
    for f in fam.values():
        assert f["bound"] in ("l2", "hbm") and f["units_per_launch"] > 0 and 0 < f["share_of_step"] < 1
    assert d["scaling"] == "strong"
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
