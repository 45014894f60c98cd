#!/usr/bin/env python
"""Zero-one-principle proof of the 99-exchange median-of-25 network in csrc/kpd_itransform.cu: the exchanges are parsed from the
source and run on all 2^25 binary inputs (an AND / OR per exchange); p[12] must be 1 exactly when at least 13 inputs are 1."""
import os, re, sys
import numpy as np
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "keypoint-detection_b200", "csrc", "kpd_itransform.cu")).read()
body = src[src.index("uint32_t median25("):src.index("return p[12];")]
net = [(int(a), int(b)) for a, b in re.findall(r"IT_CE\(p\[(\d+)\], p\[(\d+)\]\)", body)]
assert len(net) == 99, len(net)
bad, CH = 0, 1 << 22
for c in range(0, 1 << 25, CH):
    idx = np.arange(c, c + CH, dtype=np.uint32)
    v = [((idx >> i) & 1).astype(np.uint8) for i in range(25)]
    for a, b in net:
        v[a], v[b] = v[a] & v[b], v[a] | v[b]
    ones = sum(((idx >> i) & 1).astype(np.uint8) for i in range(25))
    bad += int((v[12] != (ones >= 13)).sum())
print("exchanges", len(net), "mismatches over 2^25 binary inputs:", bad)
sys.exit(1 if bad else 0)
