#!/usr/bin/env python
"""How much of the FP64 pipe can ONE CTA of 8 warps keep busy with the bisection term mix (registers only)?
k_bisect runs two such CTAs per SM and relies on the second one while the first sits at its sweep boundary."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = "import sys; sys.path.insert(0, %r); import arrowhead_b200 as ab; c = ab.get_context(0); print(c.measure_term_peak()[0], c.measure_fp64_peak()[0])" % ROOT
for ctas in ("1", "2", "3", "4"):
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, ARROWHEAD_B200_BURST_CTAS=ctas), capture_output=True, text=True)
    print(f"{ctas} CTA(s) of 256 threads per SM: term burst, DFMA burst [G/s] = {out.stdout.strip() or out.stderr[-200:]}", flush=True)
