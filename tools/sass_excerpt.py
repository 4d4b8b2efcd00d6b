"""SASS excerpt of one kernel of the built library for profiles/: opcode histogram + context around the instructions
that prove the design (TMA bulk copies UBLKCP, ticket ATOMG, shared atomics ATOMS, streaming 16-byte stores, BREV ...).

    python tools/sass_excerpt.py <mangled-name-substring> [regex ...] > profiles/<name>.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gymgo_b200", "_lib", "libgymgo_b200.so")
want = sys.argv[1]
pats = sys.argv[2:] or ["UBLKCP", "ATOMG", "ATOMS", "NANOSLEEP", r"STG\.E\.EF\.128", "SHFL.DOWN"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
name, lines, on = None, [], False
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        on = want in m.group(1) and name is None
        if on:
            name = m.group(1)
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            lines.append((m.group(1), m.group(2).strip()))
if name is None:
    sys.exit("no kernel matching %r" % want)
demangled = subprocess.run(["c++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip()
print("# cuobjdump -sass of %s" % demangled)
print("# (%s), built with the product flags (nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a)." % name)
print("# %d instructions.\n" % len(lines))
hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0] for _, ins in lines)
print("## opcode histogram (static)")
for op, n in hist.most_common(28):
    print("%7d %s" % (n, op))
for key in ("ATOMS", "ATOMG", "UBLKCP", "BREV", "SHFL", "VOTE", "STG", "LDS", "STS"):
    print("# %-7s %d" % (key, sum(1 for _, ins in lines if re.search(r"\b%s\b" % key, ins.split()[0 if not ins.startswith("@") else 1].split(".")[0]))))
for pat in pats:
    hits = [i for i, (_, ins) in enumerate(lines) if re.search(pat, ins)]
    print("\n## %s: %d site(s)" % (pat, len(hits)))
    shown = set()
    for h in hits[:3]:
        lo, hi = max(0, h - 6), min(len(lines), h + 7)
        if any(i in shown for i in range(lo, hi)):
            continue
        shown.update(range(lo, hi))
        for a, ins in lines[lo:hi]:
            print("/*%s*/ %s ;%s" % (a, ins, "      <====" if re.search(pat, ins) else ""))
        print("   ...")
