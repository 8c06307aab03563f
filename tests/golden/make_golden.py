#!/usr/bin/env python
"""Extract the known-answer vectors of the reference's own test-suite into a JSON fixture.

Reads /root/reference/test/runtests.jl (only available in the build container, never on the GPU box),
evaluates the Julia literals of the SymArrow / SymDPR1 / tdc known-answer tests (runtests.jl:22-42,
58-78, 91-115) and writes tests/golden/runtests_known_answers.json next to this script.  The reference
is Julia and cannot be executed here, so these published expected values (from the papers the tests
cite) are the only reference outputs that exist; floats are stored as hex strings to stay bit-exact.

    python tests/golden/make_golden.py
"""
import json
import os
import re

SRC = "/root/reference/test/runtests.jl"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "runtests_known_answers.json")
EPS = 2.0 ** -52


def jl_eval(expr: str):
    e = expr.strip()
    e = re.sub(r"(\d)eps\(\)", r"\1*EPS", e)
    e = e.replace("eps()", "EPS")
    return eval(e, {"EPS": EPS, "__builtins__": {}})  # noqa: S307 - literals from the reference test file


def hexlist(v):
    return [float(x).hex() for x in v]


def grab(lines, lineno, pattern):
    m = re.search(pattern, lines[lineno - 1])
    if not m:
        raise SystemExit(f"runtests.jl:{lineno} does not match {pattern!r}: {lines[lineno - 1]!r}")
    return m


def bracket_after(lines, lineno, name):
    """value of `name = [ ... ]` starting at lineno (may span lines)."""
    text = "".join(lines[lineno - 1:lineno + 30])
    start = text.index(name)
    lb = text.index("[", start)
    rb = text.index("]", lb)
    body = re.sub(r"#.*", "", text[lb:rb + 1])
    return jl_eval(body)


def main():
    lines = open(SRC, encoding="utf-8").read().splitlines(keepends=True)
    cases = {}

    def symarrow(tag, la, ll, lv=None):
        m = grab(lines, la, r"SymArrow\(\s*(\[.*?\])\s*,\s*(\[.*?\])\s*,\s*([^,]+),\s*(\d+)\s*\)")
        c = {"kind": "SymArrow", "src": f"test/runtests.jl:{la}",
             "D": hexlist(jl_eval(m.group(1))), "z": hexlist(jl_eval(m.group(2))),
             "a": float(jl_eval(m.group(3))).hex(), "i": int(m.group(4)),
             "values": hexlist(bracket_after(lines, ll, "Λₜ")), "values_src": f"test/runtests.jl:{ll}",
             "values_tol_eps": 3.0}
        if lv:
            c["v4"] = hexlist(bracket_after(lines, lv, "v4"))
            c["v4_src"] = f"test/runtests.jl:{lv}"
            c["v4_tol_eps"] = 3.0
        cases[tag] = c

    def symdpr1(tag, la, ll=None, lv=None, orth_tol=None):
        m = grab(lines, la, r"SymDPR1\(\s*(\[.*?\])\s*,\s*(\[.*?\])\s*,\s*([^)]+)\)")
        c = {"kind": "SymDPR1", "src": f"test/runtests.jl:{la}",
             "D": hexlist(jl_eval(m.group(1))), "u": hexlist(jl_eval(m.group(2))),
             "r": float(jl_eval(m.group(3))).hex()}
        if ll:
            c["values"] = hexlist(bracket_after(lines, ll, "Λₜ"))
            c["values_src"] = f"test/runtests.jl:{ll}"
            c["values_tol_eps"] = 3.0
        if lv:
            c["v4"] = hexlist(bracket_after(lines, lv, "v4"))
            c["v4_src"] = f"test/runtests.jl:{lv}"
            c["v4_tol_eps"] = 3.0
        if orth_tol:
            c["orth_tol_eps"] = orth_tol
        cases[tag] = c

    symarrow("symarrow_ex1", 22, 25, 27)
    symarrow("symarrow_ex2", 31, 34)
    symarrow("symarrow_ex3", 38, 41)
    symdpr1("symdpr1_ex1", 58, 62, 64)
    symdpr1("symdpr1_ex2", 68, 71)
    symdpr1("symdpr1_ex3_gu", 75, orth_tol=3.0)

    grab(lines, 91, r"SymTridiagonal\(abs\.\(collect\(-10\.0:10\)\),ones\(20\)\)")
    cases["tdc_w21"] = {"kind": "tdc", "src": "test/runtests.jl:91",
                        "dv": hexlist([abs(float(x)) for x in range(-10, 11)]),
                        "ev": hexlist([1.0] * 20),
                        "values": hexlist(bracket_after(lines, 93, "Λₜ")),
                        "values_src": "test/runtests.jl:93-113", "norm_tol_eps": 5.0}

    # property-only tests of the reference (no published values): recorded for the parity tests' tolerances
    cases["properties"] = {
        "symarrow_random_residual_tol_eps": 100.0,     # runtests.jl:15-19, GenSymArrow(10,10)
        "symdpr1_random_residual_tol_eps": 300.0,      # runtests.jl:51-55, GenSymDPR1(10)
        "halfarrow_orth_tol_eps": 300.0,               # runtests.jl:83-87, 8x9 exp(20(rand-0.5)) entries
    }
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(cases, f, indent=1, ensure_ascii=False)
    print("wrote", OUT, "with", len(cases), "entries")


if __name__ == "__main__":
    main()
