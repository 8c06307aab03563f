"""The C struct layout of include/arrowhead_b200.h as seen by three clients: a plain-C program (gcc), the ctypes mirror
(arrowhead.jl_b200/_lib.py) and the Julia struct in arrowhead.jl_b200/julia/ArrowheadB200.jl.

Julia is not installed in this image, so the Julia side is pinned by its documented rule instead: a `mutable struct` of
isbits fields passed as Ref{T} to ccall has C layout -- fields in declaration order, each aligned to its own size
(Ptr/Int64/Float64: 8, Int32: 4), total size rounded up to the largest alignment.  The field list is parsed out of the
.jl file between the BEGIN/END markers, so an edit there that breaks the layout fails here.

GPU: the same harness dlopen()s the library like Julia's ccall does and solves one problem of each type; every output
must be bit-identical to the ctypes path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import arrowhead_b200 as ab
from arrowhead_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_harness.c")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("abi") / "abi_harness")
    subprocess.run(["gcc", "-O1", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", out, SRC, "-ldl"],
                   check=True, env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    return out


def _c_layout(harness):
    txt = subprocess.run([harness, "layout"], check=True, capture_output=True, text=True).stdout
    lay = {}
    for line in txt.splitlines():
        name, *nums = line.split()
        lay[name] = tuple(int(x) for x in nums)
    return lay


def test_ctypes_mirror_matches_the_c_compiler(harness):
    lay = _c_layout(harness)
    for struct, cname in ((_lib.AhResult, "ah_result"), (_lib.AhStats, "ah_stats")):
        assert C.sizeof(struct) == lay[f"sizeof.{cname}"][0]
        names = [f for f, _ in struct._fields_]
        c_fields = [k.split(".", 1)[1] for k in lay if k.startswith(cname + ".")]
        assert [n.rstrip("_") for n in names] == c_fields                      # same fields, same order (lambda_ = lambda)
        for f, _ in struct._fields_:
            d = getattr(struct, f)
            assert (d.offset, d.size) == lay[f"{cname}.{f.rstrip('_')}"], f


def test_julia_struct_matches_the_c_compiler(harness):
    lay = _c_layout(harness)
    src = open(os.path.join(ROOT, "arrowhead.jl_b200", "julia", "ArrowheadB200.jl"), encoding="utf-8").read()
    block = src[src.index("# BEGIN AhResult"):src.index("# END AhResult")]
    fields = re.findall(r"^\s+(\w+)::([\w{}]+)\s*$", block, flags=re.M)
    assert len(fields) == 18
    size_of = {"Int64": 8, "Float64": 8, "Int32": 4}
    off = 0
    for name, typ in fields:
        sz = 8 if typ.startswith("Ptr{") else size_of[typ]
        off = (off + sz - 1) // sz * sz
        assert (off, sz) == lay[f"ah_result.{name}"], (name, typ)
        off += sz
    assert (off + 7) // 8 * 8 == lay["sizeof.ah_result"][0]
    # the ccall signatures name the argument types of the three drop-in entry points in header order
    for fn, sig in (("ah_symarrow_eigen", "(Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Int64, Int64, Ref{AhResult})"),
                    ("ah_symdpr1_eigen", "(Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Int64, Int64, Ref{AhResult})"),
                    ("ah_halfarrow_svd", "(Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ref{AhResult})")):
        m = re.search(r"ccall\(\(:" + fn + r", LIB\), Cint,\s*(\([^)]*\))", src)
        assert m and re.sub(r"\s+", " ", m.group(1)) == sig, fn


def _parse_solve(txt):
    info, vecs = {}, {}
    for line in txt.splitlines():
        tag, rest = line.split(" ", 1)
        if rest.startswith("k="):
            kv = dict(p.split("=") for p in rest.split())
            info.setdefault(tag, []).append(kv)
        elif rest[0] in "UV":
            head, *vals = rest.split()
            vecs.setdefault((tag, head[0]), []).append([float.fromhex(v) for v in vals])
    return info, vecs


@pytest.mark.gpu
def test_plain_c_client_gets_the_same_bits_as_ctypes(harness, gpu_ctx):
    run = subprocess.run([harness, "solve", ab.build_library()], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    assert "unsorted rc=-2" in run.stdout
    info, vecs = _parse_solve("\n".join(l for l in run.stdout.splitlines() if not l.startswith("unsorted")))
    G = {"symarrow": gpu_ctx.symarrow([2e-3, 1e-7, 0.0, -1e-7, -2e-3], [1e7, 1e7, 1.0, 1e7, 1e7], 1e20),
         "symdpr1": gpu_ctx.symdpr1([10.0, 5.0, 1.0, -2.0], [1.0, 0.5, -0.25, 2.0], 1.5),
         "halfarrow": gpu_ctx.halfarrow([3.0, 2.0, 0.5], [1.0, -0.5, 0.25, 2.0])}
    for tag, g in G.items():
        rows = info[tag]
        assert len(rows) == g.lam.size
        for c, kv in enumerate(rows):
            assert float.fromhex(kv["lambda"]) == g.lam[c] and int(kv["sind"]) == g.sind[c] and int(kv["qout"]) == g.qout[c]
            assert int(kv["status"]) == g.status[c]
            for f in ("kb", "kz", "knu", "krho"):
                assert float.fromhex(kv[f]) == getattr(g, f)[c], (tag, f)
        assert np.array_equal(np.array(vecs[(tag, "U")]).T, g.U)
        if tag == "halfarrow":
            assert np.array_equal(np.array(vecs[(tag, "V")]).T, g.V)
    assert abs(G["symarrow"].lam[1] - 0.001999001249000113) <= 3 * 2.0 ** -52 * 0.002           # runtests.jl:25
