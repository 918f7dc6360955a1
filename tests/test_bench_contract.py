"""bench.py's contract, as far as it can be checked without a GPU: the CPU arm (`--impl reference`) prints exactly ONE
JSON line on stdout with the keys the driver reads, bounded in time whatever --steps/--warmup say, and rank != 0 of a
multi-rank launch prints nothing.  The GPU arm's line is exercised on the B200 box (profiles/r2_bench_*gpu.json)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, env=e)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    t0 = time.perf_counter()
    out = run(["--impl", "reference", "--workload", "ba3000x3", "--steps", "3", "--warmup", "1", "--gpus", "1"],
              env={"ARCTE_BENCH_REF_BUDGET_S": "8"})
    assert out.returncode == 0, out.stderr[-2000:]
    assert time.perf_counter() - t0 < 120
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "arcte_seeds_per_sec" and d["unit"] == "seeds/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert abs(d["ms_per_step"] - 1e3 * d["config"]["seeds"] / d["value"]) < 1e-6 * d["ms_per_step"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] and d["config"]["nodes"] == 3000 and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    out = run(["--impl", "reference", "--workload", "ba3000x3", "--steps", "1", "--warmup", "0", "--gpus", "2"],
              env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_sample_shrinks_with_the_step_count():
    """The whole run fits the budget: more steps -> smaller per-step sample (bench.py run_reference_arm)."""
    a = run(["--impl", "reference", "--workload", "ba20000x3", "--steps", "1", "--warmup", "0"],
            env={"ARCTE_BENCH_REF_BUDGET_S": "4"})
    b = run(["--impl", "reference", "--workload", "ba20000x3", "--steps", "8", "--warmup", "2"],
            env={"ARCTE_BENCH_REF_BUDGET_S": "4"})
    assert a.returncode == 0 and b.returncode == 0
    da, db = json.loads(a.stdout.strip().splitlines()[-1]), json.loads(b.stdout.strip().splitlines()[-1])
    assert 0 < db["cpu_baseline"]["sample_seeds"] <= da["cpu_baseline"]["sample_seeds"] <= 20000
