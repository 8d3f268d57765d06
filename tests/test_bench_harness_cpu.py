"""bench.py end to end on a CPU box: the corpus generator, query decoding, recall, self-parity, oracle parity, CPU
baseline, roofline bookkeeping and the JSON contract, with the CPU oracle standing in for the GPU library
(tests/fake_plaid.py) on a tiny corpus.  Both arms.  The numbers mean nothing; the keys, types and the parity
verdicts do."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--docs-total", "3000", "--doclen", "24", "--log2k", "8", "--batch", "4", "--nq", "8", "--top-k", "5",
         "--n-full-scores", "64", "--recall-queries", "6", "--parity-queries", "5", "--docs-per-topic", "100",
         "--pool", "16", "--steps", "3", "--warmup", "1", "--threads", "2"]


def _run(extra):
    env = dict(os.environ, PB_BENCH_LIB="fake_plaid", PB_BENCH_DEVICE="cpu", PB_BENCH_CHUNK_DOCS="1000",
               PYTHONPATH=os.path.join(ROOT, "tests") + os.pathsep + ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + SMALL + extra, env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_b200_arm_contract_on_the_cpu_stand_in():
    d = _run([])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    # the stand-in IS the oracle: parity and self-parity must be perfect, recall is a number in [0, 1]
    assert d["parity"]["ids_identical"] == d["parity"]["queries"] == 5 and d["parity"]["max_abs_score_diff"] == 0.0
    assert d["self_parity"]["ids_identical"] == d["self_parity"]["queries"]
    assert 0.0 <= d["recall_at_k"] <= 1.0 and d["recall_queries"] == 6
    assert d["maxsim"]["frac_of_hbm_peak"] > 0 and "approx16" in d["roofline_all"]


def test_reference_arm_contract():
    d = _run(["--impl", "reference"])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] and d["cpu_baseline"]["cores"] >= 1
