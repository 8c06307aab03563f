#!/usr/bin/env python
"""profiles/r2_k_bisect_inner.sass: the tile loop of k_bisect<SymArrow, round 0, 8 slots> from the shipped library
(cuobjdump -sass), with the per-tile instruction census the roofline's "7.25 FP64-pipe instructions per secular term" rests on.
    python tools/sass_inner.py > profiles/r2_k_bisect_inner.sass"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "arrowhead.jl_b200", "libarrowhead_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
name = "_ZN2ah8k_bisectILi0ELi0ELi8ELi1EEE"
start = txt.index("Function : " + name)
end = txt.find("Function : ", start + 20)
fn = txt[start:end]
ins = []
for line in fn.splitlines():
    m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
op = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
# the tile loop: from the consumer's wait on the `full` mbarrier of the stage (second TRYWAIT after the sweep barrier) to the
# backward branch that closes the loop; the branch-free block = up to the first BSSY after the last LDS.128 of the slot parameters
waits = [i for i, (a, t) in enumerate(ins) if "SYNCS.PHASECHK.TRANS64.TRYWAIT" in t]
back = [i for i, (a, t) in enumerate(ins) if op(t).startswith("BRA") and re.search(r"0x([0-9a-f]+)", t) and int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) < a]
# loop head = target of the backward branch that jumps to just before the producer's code
cands = []
for i in back:
    tgt = int(re.search(r"0x([0-9a-f]+)", ins[i][1]).group(1), 16)
    body = [(a, t) for a, t in ins if tgt <= a <= ins[i][0]]
    nd = sum(1 for a, t in body if op(t).startswith("DFMA"))
    if nd >= 160 and any("UBLKCP" in t for a, t in body):
        cands.append((len(body), tgt, ins[i][0], body))
cands.sort()
_, tgt, endaddr, body = cands[0]
census = collections.Counter(op(t).split(".")[0] for a, t in body)
first_fp = next(i for i, (a, t) in enumerate(body) if op(t).startswith(("DADD", "DFMA", "DMUL")))
# branch-free block: from the pole loads to the last FP64 instruction before the first divergent-path BSSY after the block
fp_idx = [i for i, (a, t) in enumerate(body) if op(t).startswith(("DADD", "DFMA", "DMUL", "MUFU"))]
main = body[first_fp: fp_idx[263] + 1] if len(fp_idx) >= 264 else body[first_fp:]
mc = collections.Counter(op(t).split(".")[0] for a, t in main)
print(f"// k_bisect<KIND_ARROW, ROUND 0, NSLOT 8, ACCEL 1>: tile loop, SASS offsets {tgt:#06x}..{endaddr:#06x} of {name}...")
print(f"// shipped library: arrowhead.jl_b200/libarrowhead_b200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false)")
print(f"// One iteration = one tile of 1024 poles = 4 poles x 8 slots = 32 secular terms per thread.")
print(f"// Census of the branch-free block (first pole DADD .. 264th FP64/MUFU instruction, {len(main)} instructions):")
print("//   " + ", ".join(f"{k} {v}" for k, v in mc.most_common()))
fp64 = mc["DFMA"] + mc["DADD"] + mc["DMUL"]
print(f"//   FP64 pipe: DFMA {mc['DFMA']} + DADD {mc['DADD']} + DMUL {mc['DMUL']} = {fp64} = 7.25 x 32 terms;  XU pipe: MUFU.RCP64H {mc['MUFU']} = 1 x 32 terms")
print(f"//   per term: DADD d=D-sigma | DFMA e=1-m*d | DMUL t=d*e | MUFU.RCP64H r0~1/t | DFMA e2=1-t*r0 | DFMA p=e2*e2+e2 | DFMA r1=r0*p+r0 | DMUL/DFMA tile sum t(+)=w2*r1; per 4 terms one DADD acc+=t")
print(f"// Census of the whole loop body incl. the TMA producer (thread 0), the mbarrier waits and the rare shift-pole fix-up path ({len(body)} instructions):")
print("//   " + ", ".join(f"{k} {v}" for k, v in census.most_common()))
print(f"// TMA / mbarrier evidence: " + ", ".join(sorted(set(op(t) for a, t in body if op(t).startswith(("UBLKCP", "SYNCS"))))))
print()
for a, t in body:
    print(f"        /*{a:04x}*/  {t} ;")
