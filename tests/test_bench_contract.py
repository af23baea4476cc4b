"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, and the
product arm refuses to run without a CUDA device instead of falling back to CPU math."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "n4096", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(line) == 1
    d = json.loads(line[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["unit"] == "GFLOP/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # "reference" when the reference's own just_inv could be imported (this container), "port" where its file is absent
    assert d["cpu_baseline"]["kind"] == ("reference" if Path("/root/reference/matrix_inv_numpy.py").exists() else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and d["config"]["n"] == 4096 and d["config"]["same_order_as_metric"] is True


def test_product_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600)
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
