#!/usr/bin/env python
"""Launch plan of the tcgen05 conv for every distinct conv shape of the two networks (no GPU needed).
Usage: RMR_CONV_V2=1 python tools/conv_plan.py [armor_batch]"""
import ctypes as C
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rm_radar_b200 import _lib  # noqa: E402

kb = int(sys.argv[1]) if len(sys.argv) > 1 else 7
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = re.compile(r"conv\s+umma=1 in\s*(\d+)x\s*(\d+)x\s*(\d+) out\s*(\d+)x\s*(\d+)x\s*(\d+) k(\d) s(\d)")
lib = _lib.load()
seen = set()
batch = 1
print("shape                              v  N  spl halo mtiles ctas t/cta kb/tile sa sb res smemKB tile(w,h,n) bk")
for line in open(os.path.join(root, "profiles", "r1_layers.txt")):
    if line.startswith("== car"):
        batch = 1
    elif line.startswith("== armor"):
        batch = kb
    m = pat.search(line)
    if not m:
        continue
    h, w, cin, ho, wo, cout, k, s = (int(v) for v in m.groups())
    key = (batch, h, w, cin, cout, k, s)
    if key in seen:
        continue
    seen.add(key)
    out = (C.c_int * 16)()
    _lib.check(lib.rmr_conv_plan(batch, h, w, cin, cout, k, s, out))
    v = list(out)
    print(f"n{batch} {h:3d}x{w:3d} c{cin:4d}->{cout:4d} k{k}s{s}    {v[0]} {v[1]:3d} {v[2]:2d}  {v[3]}   {v[4]:5d} {v[5]:4d} {v[6]:4d} {v[7]:5d}   {v[8]:2d} {v[9]:2d}  {v[10]}  {v[11] // 1024:4d}   ({v[12]},{v[13]},{v[14]}) {v[15]}")
