"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints exactly one JSON line on
stdout with the keys the driver reads, and the b200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-power", "6"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "powers/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "powers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--ref-power", "6"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(_have_gpu(), reason="box has a GPU")
def test_b200_arm_needs_a_gpu():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)


def _load_bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_roofline_traffic_comes_from_the_named_capture():
    """`roofline.traffic` is read from the committed same-round ncu --set full summary and names it; a kernel without
    a capture gives null, never a made-up figure."""
    b = _load_bench()
    t = b.ncu_traffic("k_scalar_mul<bls12_377.g1>", {"elements": 1 << 24, "launches": 16, "ms": 1000.0})
    cap = json.load(open(os.path.join(ROOT, b.NCU_FULL)))
    rec = max((r for r in cap["launches"] if "k_scalar_mul<Bls377G1>" in r["Kernel Name"]),
              key=lambda r: int(r["Grid Size"].strip("()").split(",")[0]))
    threads = int(rec["Grid Size"].strip("()").split(",")[0]) * int(rec["Block Size"].strip("()").split(",")[0])
    mb = float(rec["dram__bytes_read.sum"].split()[0]) + float(rec["dram__bytes_write.sum"].split()[0])
    assert rec["dram__bytes_read.sum"].split()[1] == "Mbyte" and rec["dram__bytes_write.sum"].split()[1] == "Mbyte"
    assert abs(t["traffic"] - mb * 1e6 / threads * (1 << 20)) < 1e-3 * t["traffic"]
    assert b.NCU_FULL in t["traffic_source"]
    # only a one-thread G2 launch (beta_g2) is in the capture: null, with the reason
    none = b.ncu_traffic("k_scalar_mul<bls12_377.g2>", {"elements": 1 << 20, "launches": 2, "ms": 100.0})
    assert none["traffic"] is None and "traffic_note" in none


def test_committed_bench_lines_carry_the_contract_keys():
    """The evidence files of the final build are what bench.py printed: one JSON line with every key of the contract."""
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for name, n in (("r02_bench_default_final.json", 1), ("r02_bench_final_n2.json", 2), ("r02_bench_final_n4.json", 4),
                    ("r02_bench_final_n8.json", 8)):
        lines = [l for l in open(os.path.join(ROOT, "profiles", name)) if l.startswith("{")]
        assert len(lines) == 1, name
        d = json.loads(lines[0])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert key in d, (name, key)
        assert d["metric"] == base["metric"] and d["n_gpus"] == n and d["warmup"] >= 3 and d["gpu_launches"] > 0
        assert "2^22" in d["config"]["workload"] and d["scaling"] == "strong" and d["vs_baseline"] is None
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"] * 1.01
        assert abs(d["value"] - (1 << 22) / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
        r = d["roofline"]
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
        assert not d["clocks"]["reasons"] and d["verdict_all_steps"] is True and d["parity_spot_check"] is True
        if n == 1:
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
