"""CPU: bench.py's reference arm prints exactly one JSON line on stdout with the keys the driver's contract names."""
import json
import os
import subprocess
import sys

import pytest

from tests import fixtures as fx


@pytest.mark.skipif(not fx.have_onnx(), reason="fp32 ONNX copies not in this snapshot")
def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(fx.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=fx.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1, out.stdout[:500]               # stdout belongs to the JSON line alone
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "detect+locate frames/sec" and d["unit"] == "frames/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "cpu_baseline"):
        assert key in d, key
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1.0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(fx.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=fx.ROOT, env=env)
    assert out.returncode == 0 and out.stdout == ""
