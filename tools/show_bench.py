"""Pretty-print the JSON line(s) bench.py wrote:  python tools/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f)
    print(" value %.4e %s  ms/step %.5f  repeats %s launches %s" % (d["value"], d["unit"], d["ms_per_step"], d["config"].get("repeats"), d.get("gpu_launches")))
    if "roofline" in d:
        r = d["roofline"]
        print(" roofline frac %.3f achieved %.0f GB/s launch_us %.1f plies/launch %.1f traffic %s" % (r["frac"], r["achieved"], r["launch_us"], r["plies_per_launch"], r["traffic"]))
    print(" clocks", d.get("clocks"))
    e = d["e2e"]
    print(" e2e %.4e  pcie %s GB/s/GPU  d2h %d" % (e["value"], round(e.get("pcie_gbs_per_gpu", 0), 1), e["d2h_bytes_per_step"]), e.get("host_placement"))
    for k, v in e.get("variants", {}).items():
        print("   %-34s %.4e  pcie %.1f GB/s/GPU  d2h %d" % (k, v["value"], v["pcie_gbs_per_gpu"], v["d2h_bytes_per_step"]))
    if "one_launch_per_ply" in d:
        print(" one_launch_per_ply %.4e" % d["one_launch_per_ply"]["value"])
    for k, v in d.get("extra", {}).items():
        print("  extra", k, {kk: (round(vv, 5) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("roofline", "workload")}, "frac %.3f" % v["roofline"]["frac"])
    if "cpu_baseline" in d:
        print(" cpu", d["cpu_baseline"])
