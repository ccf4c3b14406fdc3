"""Host logic of bench.py that the numbers depend on (no GPU): which nvidia-smi rows count as "during the timed region",
and when ncu's DRAM traffic may be quoted (only for the kernel sources it was captured on)."""
import json

import bench


class FakeProc:
    def terminate(self): pass
    def wait(self, timeout=None): pass
    def kill(self): pass


def sampler(rows):
    s = bench.ClockSampler(0)
    s.proc, s.rows = FakeProc(), rows
    return s


def row(sm, reasons=("Not Active",) * 4):
    return ["0", str(sm), "1965", "500.0", "0x0", *reasons]


def test_clock_rows_inside_the_timed_region():
    rows = [(10.0, row(300)), (10.1, row(1900)), (10.2, row(1965)), (10.3, row(1965)), (11.0, row(200))]
    c = sampler(rows).stop(10.05, 10.25)
    assert c["samples"] == 3 and c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == []


def test_clock_rows_short_region_takes_the_first_row_after_its_start():
    rows = [(10.0, row(300)), (10.5, row(1800)), (10.6, row(210))]
    c = sampler(rows).stop(10.2, 10.21)
    assert c["samples"] == 1 and c["sm_mhz"] == 1800.0


def test_clock_reasons_are_reported():
    rows = [(10.1, row(1500, ("Not Active", "Not Active", "Active", "Active")))]
    c = sampler(rows).stop(10.0, 10.2)
    assert sorted(c["reasons"]) == ["sw_power_cap", "sw_thermal_slowdown"]


def test_no_nvidia_smi():
    s = bench.ClockSampler(0)
    assert s.stop(0, 1)["sm_mhz"] is None


def test_traffic_is_quoted_only_for_the_sources_it_was_captured_on(tmp_path, monkeypatch):
    monkeypatch.setattr(bench, "ROOT", tmp_path)
    (tmp_path / "profiles").mkdir()
    (tmp_path / "flowamok_b200" / "csrc").mkdir(parents=True)
    src = tmp_path / "flowamok_b200" / "csrc" / "k.cu"
    src.write_text("__global__ void k() {}\n")
    h = bench.csrc_hash()
    (tmp_path / "profiles" / "r99_traffic.json").write_text(json.dumps(
        {"workload": "w", "csrc_sha256": h, "dram_bytes": {"k": 1.0}, "source": "x"}))
    assert bench.ncu_traffic("w")["dram_bytes"] == {"k": 1.0}
    assert bench.ncu_traffic("another workload") is None
    src.write_text("__global__ void k() { }\n")          # any edit of the kernel sources invalidates the capture
    assert bench.csrc_hash() != h and bench.ncu_traffic("w") is None


def test_committed_traffic_file_matches_the_committed_sources():
    """profiles/r02_traffic.json must belong to the kernel sources in the tree, or bench.py reports traffic = null"""
    t = bench.ncu_traffic("synthetic-16384x16384-quick")
    assert t is not None, "profiles/*_traffic.json is stale: recapture with ncu --set full and profiles/summarize.py --traffic"
    assert set(t["dram_bytes"]) >= {"k_u_local", "k_u_iter", "k_move", "k_move_fix"}
