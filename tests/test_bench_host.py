"""Host-side pieces of bench.py that the timed legs depend on (no GPU): the clock sampler must degrade to an empty record when
neither NVML nor nvidia-smi can be reached, and the roofline denominator must follow the clock regime of the timed region."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_clock_sampler_without_a_gpu_returns_an_empty_record():
    b = _bench()
    s = b.ClockSampler(0, period=0.01)
    s.start_and_wait()
    rec = s.stop()
    assert set(("sm_mhz", "sm_max_mhz", "reasons", "samples", "source")) <= set(rec)
    if rec["samples"] == 0:                      # this container: no driver
        assert rec["sm_mhz"] is None and rec["reasons"] == []


def test_roofline_peak_follows_the_clock_regime():
    b = _bench()
    src = "measured (MEASURED_PEAKS.json, sustained)"
    peak, why = b.pick_peak({"sm_mhz": 1965.0, "sm_max_mhz": 1965.0}, 1367.5, 1653.7, src)
    assert peak == 1653.7 and "burst" in why
    peak, why = b.pick_peak({"sm_mhz": 1245.0, "sm_max_mhz": 1965.0}, 1367.5, 1653.7, src)
    assert peak == 1367.5 and why == src
    assert b.pick_peak(None, 1367.5, 1653.7, src)[0] == 1367.5


def test_launch_summary_parses_an_ncu_csv(tmp_path, capsys):
    spec = importlib.util.spec_from_file_location("ncu_launch_summary", os.path.join(ROOT, "tests", "ncu_launch_summary.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    p = tmp_path / "l.csv"
    p.write_text('==PROF== Connected\n"ID","Kernel Name","Metric Name","Metric Unit","Metric Value"\n'
                 '"0","k_a(int)","gpu__time_duration.sum","us","10.0"\n"1","k_b()","gpu__time_duration.sum","ns","30,000"\n'
                 '"2","k_a(int)","gpu__time_duration.sum","us","20.0"\n')
    m.main(str(p), "hdr")
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "hdr" and out[2].startswith("k_a(int)") and "50.0%" in out[2] and "50.0%" in out[3]
