"""bench.py's secondary measurements must never cost the headline line: `ivf_one_query_secondary` reports a failure in
its entry instead of raising, and leaves the environment knob it toggles as it found it.  (On this CPU-only box the
index constructor fails - the library has no CPU fallback - which is exactly the failure to survive.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_ivf_secondary_reports_failures_instead_of_raising(monkeypatch):
    if torch.cuda.is_available():
        import pytest
        pytest.skip("needs a box without a GPU: the entry would simply be measured")
    import bench
    monkeypatch.setenv("WB_IVF_FUSE_COARSE", "0")
    entry = bench.ivf_one_query_secondary(torch.device("cpu"), 100)
    assert "error" in entry and "parity_check" not in entry
    assert entry["config"]["workload"].startswith("IndexIVFFlat")
    assert os.environ["WB_IVF_FUSE_COARSE"] == "0"
    monkeypatch.delenv("WB_IVF_FUSE_COARSE")
    bench.ivf_one_query_secondary(torch.device("cpu"), 100)
    assert "WB_IVF_FUSE_COARSE" not in os.environ
