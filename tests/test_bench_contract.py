"""bench.py's contract, as far as it can be checked without a GPU: the workloads are the ones BASELINE.json names
(pair counts of the reference's generators), and the reference arm (`--impl reference`, the reference's CPU path timed
on the host cores) prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workloads_are_the_baseline_configs():
    assert len(bench.pair_list(500, "sequential")) == 1990            # configs[1]: |i - j| <= 4 (matcher.py:899)
    pairs = np.asarray(bench.bates_pairs(2812))
    assert pairs.shape == (42694, 2)                                  # configs[3]: geotag neighbours on the 38 x 74 grid
    assert (pairs[:, 0] < pairs[:, 1]).all() and pairs.max() == 2811
    order = np.lexsort((pairs[:, 1], pairs[:, 0]))
    assert (order == np.arange(len(pairs))).all()                     # (i, j)-sorted: contiguous shards share frames
    assert bench.FLOP_PER_PAIR_L2 == 2 * 5000 * 5000 * 128 and bench.OP_PER_PAIR_HAMMING == 2 * 5000 * 5000 * 256


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--frames", "8", "--desc", "300",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
