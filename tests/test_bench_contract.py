"""CPU: pieces of bench.py's contract that need no GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def test_committed_dram_traffic_was_captured_with_these_kernel_sources():
  """roofline.traffic / dram_frac come from profiles/traffic_*.json, and bench.py reports them only when the capture's source hash
  equals the hash of csrc/ as it is now.  The committed captures must be of the committed sources (otherwise the bench line would
  silently lose the measured-traffic fields)."""
  for tag in ("cfg3_tsp65", "cfg3_tsp65_cold"):
    d = json.load(open(os.path.join(ROOT, "profiles", "traffic_%s.json" % tag)))
    assert d["src_hash"] == bench.src_hash(), tag
    assert bench.measured_traffic(tag) is not None
    # the steady window moves ~0.8 GB per iteration for 0.74 GB of algorithmic bytes; the cold one re-reads the state per fused pass
    assert 0.5e9 < d["dram_bytes_per_iter"] < 3e9


def test_stale_traffic_capture_is_not_reported(tmp_path, monkeypatch):
  monkeypatch.setattr(bench, "src_hash", lambda: "0" * 16)
  assert bench.measured_traffic("cfg3_tsp65") is None


def test_headline_workload_is_baseline_configs_2():
  egno, ndim, nx, ny, nt, tsp, epsl, stepsz = bench.WORKLOADS["cfg3_tsp65"]
  base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
  assert (egno, ndim, nx, ny, nt) == (1, 2, 256, 256, 65) and tsp == 65
  assert "configs" in base and len(base["configs"]) >= 3
