"""bench.py host logic that needs no GPU: the clock sampler only reports rows read inside the timed region."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_clock_sampler_filters_to_the_timed_region():
    b = _bench()
    s = b.ClockSampler(0)
    na = "Not Active"
    # clocks.sm, clocks.max.sm, power, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap, host time stamp
    s.rows = [["1200", "1965", "300", na, na, "Active", na, 9.0],   # before the region (idle clocks, a stale reason)
              ["1950", "1965", "990", na, na, na, "Active", 10.05],
              ["1935", "1965", "995", na, na, na, "Active", 10.07],
              ["1965", "1965", "980", na, na, na, na, 10.09],
              ["1300", "1965", "200", "Active", na, na, na, 12.0]]   # long after the region
    s.t_begin, s.t_end = 10.0, 10.1
    r = s.summary()
    assert r["samples"] == 3 and r["sm_mhz"] == 1950.0 and r["sm_max_mhz"] == 1965.0
    assert r["reasons"] == ["sw_power_cap"]
    # no marks (or nothing read inside them): every row counts rather than none
    s.t_begin = s.t_end = None
    assert s.summary()["samples"] == 5
    s.t_begin, s.t_end = 20.0, 21.0
    assert s.summary()["samples"] == 5


def test_workload_config_names_the_baseline_configuration():
    b = _bench()
    c = b.workload_config(37)
    assert "configs[4]" in c["workload"] and c["n_atoms"] == 4_000_000 and c["frames_per_step"] == 37
    assert "larger than L2" in c["l2_policy"]
