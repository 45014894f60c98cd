#!/usr/bin/env python
"""Opcode histogram per kernel of the built library (cuobjdump -sass): the Blackwell-specific mnemonics (UTCHMMA / UTCQMMA =
tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, UCGABAR = cluster barrier) first, then the ten most frequent others.

    python tools/sass_histogram.py [lib] > profiles/rNN_sass_opcode_histogram.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "keypoint-detection_b200", "lib", "libkpd_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
SPECIAL = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "LDGSTS", "REDUX", "FFMA2")
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print(f"# {os.path.basename(lib)}: {len(hist)} kernels, {sum(sum(h.values()) for h in hist.values())} SASS instructions")
tot = collections.Counter()
for k in hist: tot.update({o: c for o, c in hist[k].items() if o in SPECIAL})
print("# library totals:", ", ".join(f"{o} {c}" for o, c in sorted(tot.items(), key=lambda x: -x[1])))
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    h = hist[k]
    name = re.sub(r"\(.*", "", demangle(k).replace("(anonymous namespace)::", ""))[:90]
    sp = ", ".join(f"{o} {h[o]}" for o in SPECIAL if h.get(o))
    rest = ", ".join(f"{o} {c}" for o, c in h.most_common(40) if o not in SPECIAL)[:160]
    print(f"{name}\n    total {sum(h.values())} | {sp or '-'}\n    {rest}")
