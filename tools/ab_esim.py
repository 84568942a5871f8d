#!/usr/bin/env python
"""Same-box A/B timing of two builds of libv2v_b200.so (paths given), interleaved to cancel drift.

    python tools/ab_esim.py libA.so libB.so
Each library is loaded in a fresh subprocess per round (ctypes cannot unload); prints median ms per variant.
"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:]
STATS = os.environ.get("AB_STATS", "1") == "1"
res = {l: {"none": [], "philox": []} for l in libs}
for rnd in range(3):
    for l in libs:
        for n in ("none", "philox"):
            env = dict(os.environ, V2V_B200_LIB=os.path.abspath(l))
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_esim.py"), "--noise", n, "--clips", "16",
                                  "--iters", "12", "--time"] + (["--stats"] if STATS else []), capture_output=True, text=True, env=env).stdout
            ms = float(out.split("ms=")[1].split()[0])
            res[l][n].append(ms)
for l in libs:
    print(l, {n: sorted(v)[len(v) // 2] for n, v in res[l].items()}, res[l])
