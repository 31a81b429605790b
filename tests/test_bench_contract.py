"""CPU: the reference arm of bench.py prints one JSON line with the keys the driver reads (same metric / unit as the GPU arm,
impl = reference, cpu_baseline and a zero-copy e2e block), says in `config.sample` what it really ran (a bounded sample mesh,
the reference's unpreconditioned CG) and carries the single-thread port next to the all-cores restatement. The GPU arm itself
needs a device and is exercised by the driver; here only its static description of the CG paths is checked."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-mesh", "12", "10", "8", "--cpu-port-mesh", "5", "5", "5", "--no-ladder"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "schur_cg_gdof_per_s" and d["unit"] == "GDOF/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    cfg = d["config"]
    assert cfg["workload"].startswith("synthetic IAEA-3D refined to 512x512x400")       # names the workload ...
    assert cfg["sample"]["mesh"] == [12, 10, 8] and "BOUNDED SAMPLE" in cfg["sample"]["note"]   # ... and says what really ran
    assert cfg["inner_solver"].startswith("reference (unpreconditioned CG")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "12x10x8" in cb["sample"]
    assert cb["single_thread_port"]["cores"] == 1 and "5x5x5" in cb["single_thread_port"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_every_cg_path_is_described():
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.PATH_KERNELS) >= {0, 2, 3, 5}
    assert "k_xrow" in bench.PATH_KERNELS[3] and "k_slab_back_update" in bench.PATH_KERNELS[5]
