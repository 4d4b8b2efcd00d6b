"""Summarise an ncu report per CUDA source line (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, cur_fn, hdr = None, None, None
agg = collections.OrderedDict()
seen_fn = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        if cur_fn not in seen_fn:
            seen_fn.append(cur_fn)
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit():
        continue
    if filt and filt not in cur_fn:
        continue
    if len(seen_fn) > 1 and cur_fn != [f for f in seen_fn if filt in f][0]:
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:90])
    ie, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    a = agg.setdefault(key, [0, 0])
    a[0] += int(r[ie])
    a[1] += int(r[si])
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print("total executed warp-instr %d, samples %d" % (tot_i, tot_s))
print("%-18s %9s %6s %6s %6s  %s" % ("file:line", "exec", "exec%", "samp", "samp%", "source"))
for (f, ln, src), (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-18s %9d %5.1f%% %6d %5.1f%%  %s" % ("%s:%d" % (f, ln), e, 100.0 * e / tot_i, s, 100.0 * s / tot_s, src))
