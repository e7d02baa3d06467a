#!/usr/bin/env python3
"""Generator for the 8x32-bit-limb Montgomery field routines (BN254 Fq and Fr).

Each routine is built as a straight-line list of PTX instructions.  The same
list is (a) emitted as ONE inline-asm block (so the carry flag never crosses an
asm boundary) into ``field_gen.cuh`` and (b) executed by the bit-exact emulator
below, which `tests/test_field_gen.py` checks against Python big-int arithmetic.
There is no GPU in the build container; this is how the carry chains are
validated before they ever run on a B200.

Multiplication is CIOS Montgomery on two interleaved accumulators ("lo-aligned"
A at limb 0 and "hi-aligned" B at limb 1) so every (mad.lo.cc, madc.hi.cc) pair
on the same operands is a 64-bit-aligned multiply-accumulate that ptxas fuses
into one IMAD.WIDE.U32(.X): 8 rounds x (8 a*b_i + 1 m + 8 m*p) = 136 IMAD.WIDE.
"""
from __future__ import annotations

import os
import random
import sys

MASK = 0xFFFFFFFF

FIELDS = {
    "fq": 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    "fr": 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
}


def limbs(x: int, n: int = 8):
    return [(x >> (32 * i)) & MASK for i in range(n)]


class Prog:
    """A straight-line PTX program over named u32 registers."""

    def __init__(self, name, inputs, outputs):
        self.name = name
        self.inputs = inputs  # list of register names bound to asm inputs
        self.outputs = outputs  # list of register names bound to asm outputs
        self.ins = []  # (op, dst, [srcs])
        self.temps = []
        self.preds = []

    def tmp(self, base):
        nm = f"{base}"
        assert nm not in self.temps and nm not in self.inputs
        self.temps.append(nm)
        return nm

    def pred(self, nm):
        self.preds.append(nm)
        return nm

    def op(self, op, dst, *srcs):
        self.ins.append((op, dst, list(srcs)))

    # ---------------- emulation ----------------
    def run(self, invals: dict) -> dict:
        regs = dict(invals)
        cf = 0

        def val(s):
            return s if isinstance(s, int) else regs[s]

        for op, dst, srcs in self.ins:
            v = [val(s) for s in srcs]
            if op == "mov.u32":
                regs[dst] = v[0]
            elif op == "mul.lo.u32":
                regs[dst] = (v[0] * v[1]) & MASK
            elif op == "mul.hi.u32":
                regs[dst] = (v[0] * v[1]) >> 32
            elif op in ("mad.lo.cc.u32", "madc.lo.cc.u32", "madc.lo.u32", "mad.lo.u32"):
                t = ((v[0] * v[1]) & MASK) + v[2] + (cf if op.startswith("madc") else 0)
                regs[dst] = t & MASK
                if ".cc" in op:
                    cf = t >> 32
            elif op in ("mad.hi.cc.u32", "madc.hi.cc.u32", "madc.hi.u32", "mad.hi.u32"):
                t = ((v[0] * v[1]) >> 32) + v[2] + (cf if op.startswith("madc") else 0)
                regs[dst] = t & MASK
                if ".cc" in op:
                    cf = t >> 32
            elif op in ("add.cc.u32", "addc.cc.u32", "addc.u32", "add.u32"):
                t = v[0] + v[1] + (cf if op.startswith("addc") else 0)
                regs[dst] = t & MASK
                if ".cc" in op:
                    cf = t >> 32
            elif op in ("sub.cc.u32", "subc.cc.u32", "subc.u32", "sub.u32"):
                t = v[0] - v[1] - (cf if op.startswith("subc") else 0)
                regs[dst] = t & MASK
                if ".cc" in op:
                    cf = 1 if t < 0 else 0
            elif op == "setp.ne.u32":
                regs[dst] = 1 if v[0] != v[1] else 0
            elif op == "selp.u32":
                regs[dst] = v[0] if v[2] else v[1]
            elif op == "and.b32":
                regs[dst] = v[0] & v[1]
            elif op == "or.b32":
                regs[dst] = v[0] | v[1]
            else:
                raise ValueError(op)
            assert 0 <= regs[dst] <= MASK
        return {o: regs[o] for o in self.outputs}

    # ---------------- emission ----------------
    def emit(self, signature: str, out_exprs, in_exprs) -> str:
        idx = {}
        for i, o in enumerate(self.outputs):
            idx[o] = f"%{i}"
        for i, s in enumerate(self.inputs):
            idx[s] = f"%{i + len(self.outputs)}"

        def fmt(s):
            if isinstance(s, int):
                return f"0x{s:08x}"
            return idx.get(s, s)

        lines = ["{"]
        if self.temps:
            lines.append(".reg .u32 " + ", ".join(self.temps) + ";")
        if self.preds:
            lines.append(".reg .pred " + ", ".join(self.preds) + ";")
        for op, dst, srcs in self.ins:
            lines.append(f"{op} {fmt(dst)}, " + ", ".join(fmt(s) for s in srcs) + ";")
        lines.append("}")
        body = "\n".join(f'        "{l}\\n\\t"' for l in lines)
        outs = ", ".join(f'"=r"({e})' for e in out_exprs)
        ins = ", ".join(f'"r"({e})' for e in in_exprs)
        return (
            f"__device__ __forceinline__ void {signature} {{\n"
            f"    asm(\n{body}\n        : {outs}\n        : {ins});\n}}\n"
        )


def cond_sub(pr: Prog, src, mod, out, tag):
    """out = src - mod if src >= mod else src   (src < 2*mod < 2^256)."""
    pl = limbs(mod)
    d = [pr.tmp(f"{tag}d{i}") for i in range(8)]
    brw = pr.tmp(f"{tag}brw")
    p = pr.pred(f"{tag}p")
    pr.op("sub.cc.u32", d[0], src[0], pl[0])
    for i in range(1, 8):
        pr.op("subc.cc.u32", d[i], src[i], pl[i])
    pr.op("subc.u32", brw, 0, 0)
    pr.op("setp.ne.u32", p, brw, 0)
    for i in range(8):
        pr.op("selp.u32", out[i], src[i], d[i], p)


def gen_mul(field: str, sqr: bool = False) -> Prog:
    mod = FIELDS[field]
    pl = limbs(mod)
    np0 = (-pow(mod, -1, 1 << 32)) & MASK
    a = [f"a{i}" for i in range(8)]
    b = a if sqr else [f"b{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_{'sqr' if sqr else 'mul'}", a + ([] if sqr else b), r)
    X = [pr.tmp(f"x{i}") for i in range(8)]  # two accumulators; roles swap every round
    Y = [pr.tmp(f"y{i}") for i in range(8)]
    m = pr.tmp("m")

    def reduce_round(A, B):
        # A is lo-aligned (limb 0), B hi-aligned (limb 1).  m = A0 * (-p^-1); T += m*p.
        pr.op("mul.lo.u32", m, A[0], np0)
        pr.op("mad.lo.cc.u32", B[0], pl[1], m, B[0])
        pr.op("madc.hi.cc.u32", B[1], pl[1], m, B[1])
        for j in (2, 4, 6):
            pr.op("madc.lo.cc.u32", B[j], pl[j + 1], m, B[j])
            pr.op("madc.hi.cc.u32" if j < 6 else "madc.hi.u32", B[j + 1], pl[j + 1], m, B[j + 1])
        pr.op("mad.lo.cc.u32", A[0], pl[0], m, A[0])
        pr.op("madc.hi.cc.u32", A[1], pl[0], m, A[1])
        for j in (2, 4, 6):
            pr.op("madc.lo.cc.u32", A[j], pl[j], m, A[j])
            pr.op("madc.hi.cc.u32", A[j + 1], pl[j], m, A[j + 1])
        pr.op("addc.u32", B[7], B[7], 0)

    # round 0: A = a_even * b0, B = a_odd * b0
    A, B = X, Y
    for j in (0, 2, 4, 6):
        pr.op("mul.lo.u32", A[j], a[j], b[0])
        pr.op("mul.hi.u32", A[j + 1], a[j], b[0])
    for j in (0, 2, 4, 6):
        pr.op("mul.lo.u32", B[j], a[j + 1], b[0])
        pr.op("mul.hi.u32", B[j + 1], a[j + 1], b[0])
    reduce_round(A, B)

    for i in range(1, 8):
        # shift T right by one limb: new lo-aligned acc = old B (+ old A1 at limb 0),
        # new hi-aligned acc = old A >> 2 limbs, written in place over old A.
        nA, nB = B, A
        oldA = A
        pr.op("add.cc.u32", nA[0], nA[0], oldA[1])
        for j in (0, 2, 4):
            pr.op("madc.lo.cc.u32", nB[j], a[j + 1], b[i], oldA[j + 2])
            pr.op("madc.hi.cc.u32", nB[j + 1], a[j + 1], b[i], oldA[j + 3])
        pr.op("madc.lo.cc.u32", nB[6], a[7], b[i], 0)
        pr.op("madc.hi.u32", nB[7], a[7], b[i], 0)
        pr.op("mad.lo.cc.u32", nA[0], a[0], b[i], nA[0])
        pr.op("madc.hi.cc.u32", nA[1], a[0], b[i], nA[1])
        for j in (2, 4, 6):
            pr.op("madc.lo.cc.u32", nA[j], a[j], b[i], nA[j])
            pr.op("madc.hi.cc.u32", nA[j + 1], a[j], b[i], nA[j + 1])
        pr.op("addc.u32", nB[7], nB[7], 0)
        A, B = nA, nB
        reduce_round(A, B)

    # result = (A >> 32) + B  (< 2p), then one conditional subtraction
    s = [pr.tmp(f"s{i}") for i in range(8)]
    pr.op("add.cc.u32", s[0], A[1], B[0])
    for j in range(1, 7):
        pr.op("addc.cc.u32", s[j], A[j + 1], B[j])
    pr.op("addc.u32", s[7], B[7], 0)
    cond_sub(pr, s, mod, r, "c")
    return pr


def gen_add(field: str) -> Prog:
    mod = FIELDS[field]
    a = [f"a{i}" for i in range(8)]
    b = [f"b{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_add", a + b, r)
    s = [pr.tmp(f"s{i}") for i in range(8)]
    pr.op("add.cc.u32", s[0], a[0], b[0])
    for i in range(1, 7):
        pr.op("addc.cc.u32", s[i], a[i], b[i])
    pr.op("addc.u32", s[7], a[7], b[7])  # a,b < p < 2^254: no carry out
    cond_sub(pr, s, mod, r, "c")
    return pr


def gen_sub(field: str) -> Prog:
    mod = FIELDS[field]
    pl = limbs(mod)
    a = [f"a{i}" for i in range(8)]
    b = [f"b{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_sub", a + b, r)
    d = [pr.tmp(f"d{i}") for i in range(8)]
    brw = pr.tmp("brw")
    pr.op("sub.cc.u32", d[0], a[0], b[0])
    for i in range(1, 8):
        pr.op("subc.cc.u32", d[i], a[i], b[i])
    pr.op("subc.u32", brw, 0, 0)  # 0xffffffff if a < b
    q = [pr.tmp(f"q{i}") for i in range(8)]
    for i in range(8):
        pr.op("and.b32", q[i], brw, pl[i])
    pr.op("add.cc.u32", r[0], d[0], q[0])
    for i in range(1, 7):
        pr.op("addc.cc.u32", r[i], d[i], q[i])
    pr.op("addc.u32", r[7], d[7], q[7])
    return pr


def gen_reduce_once(field: str) -> Prog:
    """r = a - p if a >= p else a  (for a < 2p)."""
    mod = FIELDS[field]
    a = [f"a{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_reduce_once", a, r)
    s = [pr.tmp(f"s{i}") for i in range(8)]
    for i in range(8):  # outputs may share registers with inputs: never read an input after a write
        pr.op("mov.u32", s[i], a[i])
    cond_sub(pr, s, mod, r, "c")
    return pr


ROUTINES = {
    "mul": gen_mul,
    "add": gen_add,
    "sub": gen_sub,
    "reduce_once": gen_reduce_once,
}


def emulate(field: str, what: str, a: int, b: int | None = None) -> int:
    pr = ROUTINES[what](field)
    vals = {f"a{i}": l for i, l in enumerate(limbs(a))}
    if b is not None:
        vals.update({f"b{i}": l for i, l in enumerate(limbs(b))})
    out = pr.run(vals)
    return sum(out[f"r{i}"] << (32 * i) for i in range(8))


def emit_header() -> str:
    out = [
        "// GENERATED by gen_field.py -- do not edit.  BN254 Fq/Fr on 8x32-bit limbs,",
        "// Montgomery form (R = 2^256), every routine one inline-PTX block.",
        "#pragma once",
        "#include <stdint.h>",
        "",
    ]
    def arr(x):
        return "{" + ", ".join(f"0x{l:08x}u" for l in limbs(x)) + "}"

    for field, mod in FIELDS.items():
        F = field.upper()
        out.append(f"#define {F}_MOD_LIMBS {arr(mod)}")
        out.append(f"#define {F}_ONE_LIMBS {arr((1 << 256) % mod)}  // R mod p (Montgomery 1)")
        out.append(f"#define {F}_R2_LIMBS {arr((1 << 512) % mod)}  // R^2 mod p")
        out.append(f"#define {F}_R3_LIMBS {arr((1 << 768) % mod)}  // R^3 mod p")
        out.append(f"#define {F}_PM2_LIMBS {arr(mod - 2)}  // p - 2 (Fermat inverse exponent)")
        out.append(f"#define {F}_HALF_LIMBS {arr((mod - 1) // 2)}  // (p-1)/2")
        out.append(f"#define {F}_NP0_64 0x{(-pow(mod, -1, 1 << 64)) % (1 << 64):016x}ull  // -p^-1 mod 2^64")
    out.append(f"#define FQ_SQRT_EXP_LIMBS {arr((FIELDS['fq'] + 1) // 4)}  // (p+1)/4, p = 3 mod 4")
    out.append("")
    out.append("#ifdef __CUDACC__")
    for field in ("fq", "fr"):
        for what in ("mul", "add", "sub", "reduce_once"):
            pr = ROUTINES[what](field)
            two = what in ("mul", "add", "sub")
            sig = f"{field}_{what}_ptx(uint32_t* r, const uint32_t* a" + (
                ", const uint32_t* b)" if two else ")"
            )
            outs = [f"r[{i}]" for i in range(8)]
            ins = [f"a[{i}]" for i in range(8)] + ([f"b[{i}]" for i in range(8)] if two else [])
            out.append(pr.emit(sig, outs, ins))
    out.append("#endif  // __CUDACC__")
    return "\n".join(out) + "\n"


def selftest(iters: int = 300) -> None:
    rnd = random.Random(1234)
    for field, mod in FIELDS.items():
        rinv = pow(1 << 256, -1, mod)
        edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 253), (1 << 32) - 1, ((1 << 256) % mod)]
        cases = [(x, y) for x in edge for y in edge]
        cases += [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(iters)]
        for x, y in cases:
            assert emulate(field, "mul", x, y) == x * y * rinv % mod, (field, "mul", x, y)
            assert emulate(field, "add", x, y) == (x + y) % mod, (field, "add", x, y)
            assert emulate(field, "sub", x, y) == (x - y) % mod, (field, "sub", x, y)
        # multiplicand a < p, word operand b ANY 256-bit value (used for bytes -> Fr)
        for _ in range(iters):
            x, y = rnd.randrange(mod), rnd.randrange(1 << 256)
            assert emulate(field, "mul", x, y) == x * y * rinv % mod, (field, "mulwide", x, y)
        for x in (mod - 1, (1 << 256) % mod, rnd.randrange(mod)):
            for y in ((1 << 256) - 1, (1 << 256) - 2, 1 << 255, mod, 2 * mod, 5 * mod + 7):
                assert emulate(field, "mul", x, y) == x * y * rinv % mod, (field, "mulwide", x, y)
        for x in edge + [rnd.randrange(2 * mod) for _ in range(iters)] + [mod, mod + 1, 2 * mod - 1]:
            assert emulate(field, "reduce_once", x) == x % mod
    print("gen_field selftest OK")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--selftest":
        selftest()
    else:
        here = os.path.dirname(os.path.abspath(__file__))
        with open(os.path.join(here, "field_gen.cuh"), "w") as f:
            f.write(emit_header())
        print("wrote field_gen.cuh")
