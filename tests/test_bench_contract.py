"""bench.py contract, the part that runs without a GPU: the reference arm (`--impl reference`) times the unmodified
reference package (baseline/_ref or /root/reference; the torch port of its loop when neither is there) on the host
cores and prints ONE JSON line with the driver's keys; ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e)


def test_reference_arm_prints_one_contract_line():
    res = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", "joint_tiny")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [line for line in res.stdout.splitlines() if line.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "map_iterations_per_sec" and d["unit"] == "iter/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 2 and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["config"]["workload"].startswith("joint_tiny")
    assert "joint" in d["config"]["iteration_semantics"]
    cb = d["cpu_baseline"]
    from oracle import ref_runner

    assert cb["kind"] == ("reference" if ref_runner.locate_reference() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "joint_tiny" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    res = run_bench("--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--workload", "joint_tiny",
                    env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_default_workload_is_the_north_star():
    import bench

    sys.argv, old = ["bench.py"], sys.argv
    try:
        args = bench.parse_args()
    finally:
        sys.argv = old
    assert args.workload == "joint1024" and args.collective == "peer" and "joint1024" in bench.JOINT_WORKLOADS
