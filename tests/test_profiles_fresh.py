"""The ncu counters bench.py reports as `roofline.executed` (profiles/r02/traffic.json) must belong to the kernel
sources and workload generator in the tree: bench.py drops them when the stamped hash differs.  A stale capture is a
skip here (visible in the test summary, not a failure): refresh with profiles/scripts/round2_refresh.sh +
profiles/scripts/make_traffic.py on a B200."""
import json
import os

import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_capture_matches_sources():
    tr = json.load(open(os.path.join(ROOT, "profiles", "r02", "traffic.json")))
    for name in ("amars_lw", "amars_sw", "lbl", "gcm"):
        e = tr[name]
        assert e["solves_per_launch"] > 0 and e["executed_flop_per_solve"] > 0 and e["bytes_per_launch"] > 0
        assert os.path.exists(os.path.join(ROOT, e["capture"])), e["capture"]
    if tr["hash"] != bench.kernel_hash():
        pytest.skip("profiles/r02/traffic.json is stale: kernel sources / workloads changed since the ncu capture "
                    "(bench.py will report roofline.executed = null, stale = true)")
    entry, why = bench.traffic_entry("amars_lw")
    assert entry is not None and why is None
