"""CPU checks of bench.py's bookkeeping (no GPU work): algorithmic byte counts of SURVEY.md 8(d), the measured-peak
lookup, the profiled-traffic lookup, core counting and the reference arm's JSON contract on a tiny sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_survey():
    assert bench.algorithmic_bytes_per_ply(9, 4) == 2028
    assert bench.algorithmic_bytes_per_ply(19, 4) == 8964
    assert bench.algorithmic_bytes_per_ply(9, 1) == 570
    assert bench.algorithmic_bytes_per_ply(19, 1) == 2466
    assert bench.algorithmic_bytes_per_ply(9, 0) == 84 and bench.algorithmic_bytes_per_ply(19, 0) == 300


def test_peak_and_traffic_lookups():
    peak, src = bench.measured_peak_gbs()
    assert 3000 < peak < 9000 and ("measured" in src or "fallback" in src)
    t = bench.profiled_traffic(9, "f32", 32)
    assert t is None or (3e9 < t["dram_bytes_per_launch"] < 5e9 and "profiles/" in t["source"])
    t20 = bench.profiled_traffic(9, "f32", 20)              # scaled to the plies a launch really plays
    assert (t is None) == (t20 is None)
    if t is not None:
        assert abs(t20["dram_bytes_per_launch"] * 32 - t["dram_bytes_per_launch"] * 20) <= 64
    assert bench.profiled_traffic(9, "f64", 32) is None
    assert 1 <= bench.usable_cores() <= (os.cpu_count() or 1)


def test_repetition_count_covers_the_minimum_timed_region():
    """a short driver run (e.g. --steps 20 = 0.5 ms of device time) is repeated until the timed region is long enough to
    be representative; a long one is not repeated"""
    assert bench.pick_repeats(0.0005) * 0.0005 >= bench.MIN_TIMED_MS / 1e3
    assert bench.pick_repeats(0.0005) <= 2 * bench.MIN_TIMED_MS / 0.5
    assert bench.pick_repeats(bench.MIN_TIMED_MS / 1e3) == 1 and bench.pick_repeats(10.0) == 1
    assert bench.PREROLL >= 200


def test_reference_arm_json_contract():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "6", "--warmup", "3"],
                       stdout=subprocess.PIPE, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"] > 0
    assert line["steps"] == 6 and line["warmup"] == 3


def test_developer_tools_parse():
    import ast
    tools = os.path.join(ROOT, "tools")
    for name in sorted(os.listdir(tools)):
        if name.endswith(".py"):
            ast.parse(open(os.path.join(tools, name)).read(), filename=name)
