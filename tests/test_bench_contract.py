"""bench.py's reference arm (`--impl reference`) runs on the CPU: check the JSON contract of that line here;
the B200 arm shares the code that builds the common keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-per-thread", "1", "--ref-per-thread-small", "1", "--ref-pairs", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"] == "importance-sampled histories/sec" and line["unit"] == "histories/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    for key in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["config"]["workload"].startswith("cfg2")
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_workload_is_deterministic():
    sys.path.insert(0, ROOT)
    import bench
    a = [p.dumps() for p in bench.bench_problems()]
    b = [p.dumps() for p in bench.bench_problems()]
    assert a == b and len(a) == 24 == len(bench.config_block("cfg2")["pairs"])
    assert all(p.T == 5 and p.n_lineages() == 50 for p in bench.bench_problems())
    # the pair sample is seeded, not hand-picked, and both arms print the same config object
    import random
    assert bench.config_block("cfg2")["pairs"] == sorted(random.Random(bench.PAIR_SEED).sample(range(len(bench.bench_pairs("cfg2")[2])), 24))
    g3 = bench.bench_problems("cfg3")
    assert len(g3) == 32 and all(p.G == 1024 for p in g3)
