"""CPU-side checks of bench.py's bookkeeping: the roofline's `traffic` is READ from the committed ncu summaries (never a
constant in the source), so the kernel names bench.py looks for must exist in those files, and the measured L2 ceilings must load."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def test_bench_finds_ncu_traffic_for_the_kernels_of_the_step():
    import bench
    tr = bench.ncu_traffic()
    for entry in ("acn_render_expert_fwd", "acn_render_expert_bwd", "acn_composite_fwd", "acn_composite_bwd"):
        assert entry in tr, f"{entry}: kernel '{bench.NCU_KERNEL[entry]}' not found in {bench.NCU_FILES}"
        nbytes, src = tr[entry]
        assert nbytes > 1e8 and src == bench.NCU_FILES[0], (entry, nbytes, src)       # the freshest capture has them
    # the two fused kernels: DRAM traffic far BELOW the algorithmic bytes (the table is L2-resident), as DESIGN 4 states
    P = bench.N_RAYS * bench.SAMPLES
    assert tr["acn_render_expert_bwd"][0] < 0.1 * bench.FUSED_BWD_BYTES * P
    assert tr["acn_render_expert_fwd"][0] < 0.15 * (bench.ENC_FWD_BYTES + 16) * P


def test_bench_loads_the_measured_l2_ceilings():
    import bench
    l2 = bench.l2_peaks()
    assert l2 and 100 < l2["gather_g_per_s"] < 1000 and 50 < l2["red_g_per_s"] < 1000


def test_clock_sampler_summarises_only_the_timed_window():
    import bench
    c = bench.ClockSampler(0)
    c.rows = [(10.0, ["1500", "1965", "Not Active", "Not Active", "Not Active", "Not Active"]),
              (11.0, ["1965", "1965", "Not Active", "Not Active", "Not Active", "Active"]),
              (12.0, ["1000", "1965", "Active", "Not Active", "Not Active", "Not Active"])]
    c.t0, c.t1 = 10.5, 11.5
    s = c.summary()
    assert s["samples"] == 1 and s["sm_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"] and s["where"] == "timed region"
    c.t0, c.t1 = 10.2, 10.3          # a region between two samples: the nearest one, and it says so
    s = c.summary()
    assert s["samples"] == 1 and s["sm_mhz"] == 1500.0 and "nearest" in s["where"]
