"""bench.py keeps the driver's contract: one JSON line with the agreed keys, for both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    """`--impl reference`: the oracle port on the host cores, no GPU needed."""
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["impl"] == "reference" and d["metric"] == "stylised_images_per_sec_512" and d["unit"] == "images/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_cuda_arm_line():
    """The CUDA arm at a reduced batch (contract only; the numbers are the full bench's business)."""
    d = _run(["--batch", "2", "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--sustain-s", "1.0"], 900)
    assert (BASE_KEYS | {"roofline", "clocks", "configs", "sustained", "gpu_eager_baseline"}) <= set(d)
    assert {"config1_batch6_512", "config2_overall_stats_2048", "config4_single_style", "config5_camelyon_96"} <= set(d["configs"])
    assert d["configs"]["config2_overall_stats_2048"]["allreduce_us"] == 0.0  # single rank: no collective
    assert d["sustained"]["seconds"] >= 0.5 and d["sustained"]["value"] > 0
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["scaling"] == "weak" and d["dtype"] == "f16"
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] and r["peak"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert d["roofline_stats"]["bound"] == "hbm" and d["roofline_adain"]["bound"] == "hbm"
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 2 * 3 * 512 * 512 and e["d2h_bytes_per_step"] == e["h2d_bytes_per_step"]
    assert e["fp32_host_tensors"]["h2d_bytes_per_step"] == 4 * e["h2d_bytes_per_step"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
