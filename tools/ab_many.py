#!/usr/bin/env python
"""Same-box comparison of several builds: python tools/ab_many.py n world rounds lib1.so lib2.so ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n, world, rounds, libs = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4:]
code = open(os.path.join(ROOT, "tools", "ab_compare.py")).read().split("code = r'''")[1].split("''' % (ROOT, ROOT)")[0] % (ROOT, ROOT)
for r in range(rounds):
    for lib in libs:
        env = dict(os.environ, ARROWHEAD_B200_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, "-c", code, n, world], env=env, capture_output=True, text=True)
        print(f"round {r} {os.path.basename(lib):28s} n={n}/{world}: {out.stdout.strip() or out.stderr[-300:]}", flush=True)
