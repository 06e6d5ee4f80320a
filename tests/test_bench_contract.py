"""bench.py on the CPU: the reference arm (the only arm that runs without a GPU) prints ONE JSON line with the
contract's keys and the size it actually measured -- nothing extrapolated -- and the fixtures the GPU arm reads
(single-GPU omega of the bench workloads, ncu traffic of the roofline kernel) are well formed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line_at_its_measured_size():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2",
                          "--warmup", "1", "--cpu-sample-dofs", "1500"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.strip().split("\n") if ln.startswith("{")]
    assert len(lines) == 1
    b = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in b, k
    assert b["impl"] == "reference" and b["metric"] == "converged_omega_solve_time" and b["unit"] == "s"
    assert b["steps"] == 2 and b["warmup"] == 1 and b["higher_is_better"] is False and b["vs_baseline"] is None
    assert b["e2e"] == {"value": b["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert b["cpu_baseline"]["kind"] == "port" and b["cpu_baseline"]["value"] == b["value"]
    assert b["cpu_baseline"]["cores"] >= 1
    # the size in the line is the size that ran; no scale factor anywhere
    assert b["config"]["dofs"] == b["config"]["sample_dofs"] == b["cpu_baseline"]["sample_dofs"]
    assert "scale" not in b["config"] and "scale" not in b["cpu_baseline"]
    assert abs(b["ms_per_step"] - 1e3 * b["value"]) < 1.0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_bench_fixtures_are_well_formed():
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "tests", "golden", "bench_omega.json")) as fh:
        table = json.load(fh)
    sizes = [k for k in table if k.isdigit()]
    assert str(bench.DEFAULT_DOFS // 1) not in ("",) and len(sizes) >= 3
    for k in sizes:
        assert len(table[k]) == 2 and 3000 < table[k][0] < 4000 and 300 < table[k][1] < 400
        om = bench._omega_reference(int(k))
        assert om is not None and abs(om - complex(*table[k])) == 0.0
    assert bench._omega_reference(12345) is None
    # the default workload has both a single-GPU omega and an ncu traffic record
    from helmholtz_x_b200 import synthetic
    n_r, n_t, n_z = synthetic.grid_for_dofs(bench.DEFAULT_DOFS, 1)
    n = n_r * n_t * n_z
    assert bench._omega_reference(n) is not None
    with open(os.path.join(ROOT, "profiles", "ncu_sell_traffic.json")) as fh:
        recs = json.load(fh)
    rec = [r for r in recs if r["n"] == n]
    assert rec and 0.9 < rec[0]["dram_bytes"] / rec[0]["algorithmic_bytes"] < 1.1
    traffic, src = bench._sell_traffic(n, rec[0]["nnz"])
    assert traffic == rec[0]["dram_bytes"] and os.path.exists(os.path.join(ROOT, src.split(" ")[0]))
    assert bench._sell_traffic(n, 17) == (None, None)
