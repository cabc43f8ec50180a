"""CPU checks of bench.py's host logic: workload table, the clock sampler's no-GPU behaviour, the reference arm's line."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_workloads_name_existing_configs_and_restate_baseline_configs():
    b = _bench()
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert len(base["configs"]) >= 5
    for name, (cfg, batch, H, W, what) in b.WORKLOADS.items():
        assert os.path.exists(os.path.join(ROOT, cfg)), name
        assert batch > 0 and H % 32 == 0 and W % 32 == 0
        assert what.startswith("configs["), name
    # the size-bucketed variant splits the mixed batch evenly over the three sizes
    assert b.WORKLOADS["hrsc_r50_bucketed"][1] % len(b.MIXED_SIZES) == 0
    assert b.WORKLOADS["hrsc_r50_bucketed"][1] == b.WORKLOADS["hrsc_r50_mixed"][1]
    cfg, spec, batch, H, W, what = b.load_spec("r101_b32")
    assert spec.resnet_depth == 101 and batch == 32


def test_clock_sampler_without_a_gpu_reports_no_samples():
    b = _bench()
    s = b.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
    if out["samples"] == 0:  # this container: neither NVML nor nvidia-smi
        assert out["sm_mhz"] is None and out["reasons"] == []


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` = the restated reference on the host cores (the oracle is the checker being timed)."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images_per_sec_1024x1024" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_claim_stdout_keeps_stdout_for_the_json_line_only():
    code = ("import sys, os; sys.path.insert(0, %r); import importlib.util as u;"
            "sp = u.spec_from_file_location('b', os.path.join(%r, 'bench.py')); b = u.module_from_spec(sp); sp.loader.exec_module(b);"
            "out = b.claim_stdout(); print('noise from a library'); os.write(1, b'raw fd-1 noise\\n');"
            "print('{\"ok\": 1}', file=out, flush=True)") % (ROOT, ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == '{"ok": 1}'
    assert "noise from a library" in r.stderr and "raw fd-1 noise" in r.stderr
