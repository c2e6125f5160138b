"""CPU checks of bench.py's bookkeeping (no GPU work): the roofline objects of the two models, the algorithmic sample count of
SURVEY.md 8(d), the configs named by BASELINE.json."""
import importlib.util
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def bench():
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec = importlib.util.spec_from_file_location("bench_module", ROOT / "bench.py")
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        sys.argv = argv
    return m


def test_algorithmic_samples_follow_the_survey(bench):
    # 14 hypotheses x 10 views x 36 taps per pixel visit, half the pixels per pass
    assert bench.algorithmic_samples_per_pass(3200, 2130, 10) == 14 * 10 * 36 * (3200 * 2130 // 2) == 17176320000


def test_configs_are_the_baseline_configs(bench):
    base = json.load(open(ROOT / "BASELINE.json"))
    assert base["metric"].startswith("depth maps/s")
    assert bench.UNIT == "depth maps/s"
    c2, c4, c3 = bench.CONFIGS["C2"], bench.CONFIGS["C4"], bench.CONFIGS["C3"]
    assert (c2["width"], c2["height"], c2["n_src"]) == (3200, 2130, 10) and c2["model"] == "pinhole"
    assert (c4["width"], c4["height"], c4["n_src"]) == (4096, 2048, 8) and c4["model"] == "sphere"
    assert c3["scene_views"] == 64 and c3["n_src"] == 10


@pytest.mark.parametrize("cfg,size,ms", [("C2", (3200, 2130), 21.3), ("C4", (3200, 1600), 11.8)])
def test_roofline_object_is_consistent(bench, cfg, size, ms):
    peaks = bench.load_peaks()
    r = bench.roofline_of(bench.CONFIGS[cfg], size[0], size[1], {"photometric": ms}, peaks)
    assert r["bound"] in ("tex", "issue") and r["unit"] and r["peak"] > 0
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert 0.0 < r["frac"] <= 1.0                                    # a utilisation, never above the unit's peak
    assert r["algorithmic_samples_per_launch"] == bench.algorithmic_samples_per_pass(size[0], size[1], bench.CONFIGS[cfg]["n_src"])
    assert r["executed_samples_per_launch"] <= r["algorithmic_samples_per_launch"]
    assert r["hbm"]["achieved_gbs"] < 0.05 * r["hbm"]["peak_gbs"]    # not the bound, and said so
    if cfg == "C4":
        assert r["sfu"]["algorithmic_over_peak"] == pytest.approx(r["sfu"]["algorithmic_gsample_s"] / r["sfu"]["peak_gsample_s"])
    assert bench.roofline_of(bench.CONFIGS[cfg], size[0], size[1], {}, peaks) is None      # no pass timed: no claim
