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
            if op.startswith("@"):  # "@pred op": skipped entirely (carry flag included) when the predicate is false
                guard, op = op[1:].split(" ", 1)
                if not regs[guard]:
                    continue
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
            elif op == "xor.b32":
                regs[dst] = v[0] ^ v[1]
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


def gen_mul(field: str, sqr: bool = False, nr: bool = False) -> Prog:
    mod = FIELDS[field]
    pl = limbs(mod)
    np0 = (-pow(mod, -1, 1 << 32)) & MASK
    a = [f"a{i}" for i in range(8)]
    b = a if sqr else [f"b{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_{'sqr' if sqr else 'mul'}{'nr' if nr else ''}", a + ([] if sqr else b), r)
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
    if nr:  # "not reduced": operands and result in [0, 2p) (4p < 2^256), no final subtraction
        for i in range(8):
            pr.op("mov.u32", r[i], s[i])
    else:
        cond_sub(pr, s, mod, r, "c")
    return pr


# ---------------------------------------------------------------------------------------------
# Karatsuba product + separate Montgomery reduction ("mulk").  IMAD.WIDE issues at a quarter of the
# scheduler rate on sm_100 (measured: half the 32-bit IMAD rate), so the integer pipe is the
# bound: one Karatsuba level turns the 64 wide multiplies of the 8x8 product into 3 x 16 = 48 plus
# ~90 additions; the reduction keeps its 64.  NEGATIVE RESULT (B200, profiles/r01_accumulate_variants.txt):
# 5.60e10 mul/s against 6.70e10 for the CIOS routine -- every added IADD3/LOP3 costs an issue cycle
# (cycles ~ 4 x IMAD.WIDE + 2 x IMAD + 1 x everything else), so 16 fewer wide multiplies (64 cycles)
# do not pay for ~110 extra ALU instructions.  Not emitted into field_gen.cuh; kept with its
# emulator check because the separate product / reduction blocks are what lazy reduction would need.
# ---------------------------------------------------------------------------------------------
class Namer:
    def __init__(self, pr, prefix):
        self.pr, self.prefix, self.k = pr, prefix, 0

    def new(self, n=1):
        out = []
        for _ in range(n):
            out.append(self.pr.tmp(f"{self.prefix}{self.k}"))
            self.k += 1
        return out if n > 1 else out[0]


def emit_mulw(pr: Prog, nm: Namer, a, b, n):
    """Schoolbook n x n -> 2n limbs (n even) with two shifting accumulators (lo-aligned A, hi-aligned B),
    the CIOS skeleton of gen_mul without the reduction; every (lo, hi) pair fuses into one IMAD.WIDE."""
    T = []
    A = nm.new(n)
    B = nm.new(n)
    for j in range(0, n, 2):
        pr.op("mul.lo.u32", A[j], a[j], b[0])
        pr.op("mul.hi.u32", A[j + 1], a[j], b[0])
    for j in range(0, n, 2):
        pr.op("mul.lo.u32", B[j], a[j + 1], b[0])
        pr.op("mul.hi.u32", B[j + 1], a[j + 1], b[0])
    T.append(A[0])
    for i in range(1, n):
        oldA, oldB = A, B
        nA0 = nm.new()
        pr.op("add.cc.u32", nA0, oldB[0], oldA[1])
        nB = nm.new(n)
        for j in range(0, n - 2, 2):
            pr.op("madc.lo.cc.u32", nB[j], a[j + 1], b[i], oldA[j + 2])
            pr.op("madc.hi.cc.u32", nB[j + 1], a[j + 1], b[i], oldA[j + 3])
        pr.op("madc.lo.cc.u32", nB[n - 2], a[n - 1], b[i], 0)
        pr.op("madc.hi.u32", nB[n - 1], a[n - 1], b[i], 0)
        nA = nm.new(n)
        pr.op("mad.lo.cc.u32", nA[0], a[0], b[i], nA0)
        pr.op("madc.hi.cc.u32", nA[1], a[0], b[i], oldB[1])
        for j in range(2, n, 2):
            pr.op("madc.lo.cc.u32", nA[j], a[j], b[i], oldB[j])
            pr.op("madc.hi.cc.u32", nA[j + 1], a[j], b[i], oldB[j + 1])
        top = nm.new()
        pr.op("addc.u32", top, nB[n - 1], 0)
        nB = nB[: n - 1] + [top]
        A, B = nA, nB
        T.append(A[0])
    # remaining window: (A >> 32) + B
    hi = nm.new(n)
    pr.op("add.cc.u32", hi[0], A[1], B[0])
    for k in range(1, n - 1):
        pr.op("addc.cc.u32", hi[k], A[k + 1], B[k])
    pr.op("addc.u32", hi[n - 1], B[n - 1], 0)
    return T + hi


def emit_absdiff(pr: Prog, nm: Namer, x, y, n):
    """|x - y| (n limbs) and the mask (all ones if x < y)."""
    d = nm.new(n)
    pr.op("sub.cc.u32", d[0], x[0], y[0])
    for i in range(1, n):
        pr.op("subc.cc.u32", d[i], x[i], y[i])
    m = nm.new()
    pr.op("subc.u32", m, 0, 0)  # 0xffffffff on borrow
    e = nm.new(n)
    for i in range(n):
        pr.op("xor.b32", e[i], d[i], m)
    r = nm.new(n)
    pr.op("sub.cc.u32", r[0], e[0], m)
    for i in range(1, n):
        pr.op("subc.cc.u32" if i < n - 1 else "subc.u32", r[i], e[i], m)
    return r, m


def emit_karatsuba8(pr: Prog, nm: Namer, a, b):
    """8 x 8 -> 16 limbs: z0 = lo*lo, z2 = hi*hi, middle = z0 + z2 - (a_lo - a_hi)(b_lo - b_hi)."""
    z0 = emit_mulw(pr, nm, a[0:4], b[0:4], 4)
    z2 = emit_mulw(pr, nm, a[4:8], b[4:8], 4)
    da, ma = emit_absdiff(pr, nm, a[0:4], a[4:8], 4)
    db, mb = emit_absdiff(pr, nm, b[4:8], b[0:4], 4)  # (a_lo - a_hi)(b_hi - b_lo) = a_lo b_hi + a_hi b_lo - z0 - z2
    z1 = emit_mulw(pr, nm, da, db, 4)
    ms = nm.new()
    pr.op("xor.b32", ms, ma, mb)  # all ones: the product above is negative
    # S = z0 + z2 (9 limbs)
    S = nm.new(9)
    pr.op("add.cc.u32", S[0], z0[0], z2[0])
    for i in range(1, 8):
        pr.op("addc.cc.u32", S[i], z0[i], z2[i])
    pr.op("addc.u32", S[8], 0, 0)
    # M = S + (ms ? -z1 : z1) = S + (z1 ^ ms) + (ms & 1), 9 limbs, never negative
    zx = nm.new(8)
    for i in range(8):
        pr.op("xor.b32", zx[i], z1[i], ms)
    dummy = nm.new()
    pr.op("add.cc.u32", dummy, ms, 1)  # carry = 1 iff ms is all ones
    M = nm.new(9)
    for i in range(8):
        pr.op("addc.cc.u32", M[i], S[i], zx[i])
    pr.op("addc.u32", M[8], S[8], ms)
    # T = z0 + M << 128 + z2 << 256
    T = list(z0[0:4]) + nm.new(12)
    pr.op("add.cc.u32", T[4], z0[4], M[0])
    for i in range(1, 4):
        pr.op("addc.cc.u32", T[4 + i], z0[4 + i], M[i])
    for i in range(4):
        pr.op("addc.cc.u32", T[8 + i], z2[i], M[4 + i])
    pr.op("addc.cc.u32", T[12], z2[4], M[8])
    pr.op("addc.cc.u32", T[13], z2[5], 0)
    pr.op("addc.cc.u32", T[14], z2[6], 0)
    pr.op("addc.u32", T[15], z2[7], 0)
    return T


def emit_redc(pr: Prog, nm: Namer, T, mod, out, tag, nr: bool = False):
    """Montgomery reduction of a 16-limb T < mod * 2^256: out = T / 2^256 mod p (fully reduced).
    Word-serial on the low half with the two shifting accumulators of gen_mul, then + T_hi."""
    pl = limbs(mod)
    np0 = (-pow(mod, -1, 1 << 32)) & MASK
    A = list(T[0:8])
    B = None
    for i in range(8):
        if i == 0:
            t0 = A[0]
            m = nm.new()
            pr.op("mul.lo.u32", m, t0, np0)
            nB = nm.new(8)
            for j in (0, 2, 4, 6):
                pr.op("mul.lo.u32", nB[j], pl[j + 1], m)
                pr.op("mul.hi.u32", nB[j + 1], pl[j + 1], m)
            src = A
        else:
            oldA, oldB = A, B
            t0 = nm.new()
            pr.op("add.cc.u32", t0, oldB[0], oldA[1])
            m = nm.new()
            pr.op("mul.lo.u32", m, t0, np0)  # does not touch the carry flag
            nB = nm.new(8)
            for j in (0, 2, 4):
                pr.op("madc.lo.cc.u32", nB[j], pl[j + 1], m, oldA[j + 2])
                pr.op("madc.hi.cc.u32", nB[j + 1], pl[j + 1], m, oldA[j + 3])
            pr.op("madc.lo.cc.u32", nB[6], pl[7], m, 0)
            pr.op("madc.hi.u32", nB[7], pl[7], m, 0)
            src = [t0] + list(oldB[1:8])
        nA = nm.new(8)
        pr.op("mad.lo.cc.u32", nA[0], pl[0], m, src[0])
        pr.op("madc.hi.cc.u32", nA[1], pl[0], m, src[1])
        for j in (2, 4, 6):
            pr.op("madc.lo.cc.u32", nA[j], pl[j], m, src[j])
            pr.op("madc.hi.cc.u32", nA[j + 1], pl[j], m, src[j + 1])
        top = nm.new()
        pr.op("addc.u32", top, nB[7], 0)
        A, B = nA, nB[:7] + [top]
    # (A >> 32) + B + T_hi  (< 2p), one conditional subtraction
    s = nm.new(8)
    pr.op("add.cc.u32", s[0], A[1], B[0])
    for j in range(1, 7):
        pr.op("addc.cc.u32", s[j], A[j + 1], B[j])
    pr.op("addc.u32", s[7], B[7], 0)
    u = nm.new(8)
    pr.op("add.cc.u32", u[0], s[0], T[8])
    for j in range(1, 7):
        pr.op("addc.cc.u32", u[j], s[j], T[8 + j])
    pr.op("addc.u32", u[7], s[7], T[15])
    if nr:
        for i in range(8):
            pr.op("mov.u32", out[i], u[i])
    else:
        cond_sub(pr, u, mod, out, tag)


def gen_mulk(field: str) -> Prog:
    mod = FIELDS[field]
    a = [f"a{i}" for i in range(8)]
    b = [f"b{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_mulk", a + b, r)
    nm = Namer(pr, "t")
    T = emit_karatsuba8(pr, nm, a, b)
    emit_redc(pr, nm, T, mod, r, "c")
    return pr


def gen_mul2sub(field: str, nr: bool = False) -> Prog:
    """r = (a*b - c*d) / 2^256 mod p with ONE Montgomery reduction (lazy reduction of a difference of
    products: Y3 = R*(Q - X3) - Y1*PPP in XYZZ += affine).  Two plain 8x8 products (64 wide multiplies
    each), a 16-limb subtraction, + p*2^256 when it borrowed (so the value is in [0, p*2^256)), then the
    word-serial reduction (64 wide + 8 narrow): 192 + 8 instead of 2 x (128 + 8) multiplies."""
    mod = FIELDS[field]
    pl = limbs(mod)
    a = [f"a{i}" for i in range(8)]
    b = [f"b{i}" for i in range(8)]
    c = [f"c{i}" for i in range(8)]
    d = [f"d{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_mul2sub{'nr' if nr else ''}", a + b + c + d, r)
    nm = Namer(pr, "t")
    T1 = emit_mulw(pr, nm, a, b, 8)
    T2 = emit_mulw(pr, nm, c, d, 8)
    D = nm.new(16)
    pr.op("sub.cc.u32", D[0], T1[0], T2[0])
    for i in range(1, 16):
        pr.op("subc.cc.u32", D[i], T1[i], T2[i])
    brw = nm.new()
    pr.op("subc.u32", brw, 0, 0)  # 0xffffffff if a*b < c*d
    q = nm.new(8)
    for i in range(8):
        pr.op("and.b32", q[i], brw, pl[i])
    E = nm.new(8)
    pr.op("add.cc.u32", E[0], D[8], q[0])
    for i in range(1, 7):
        pr.op("addc.cc.u32", E[i], D[8 + i], q[i])
    pr.op("addc.u32", E[7], D[15], q[7])  # the carry out cancels the borrow
    emit_redc(pr, nm, D[0:8] + E, mod, r, "c", nr)
    return pr


def emit_sqrw(pr: Prog, nm: Namer, a):
    """8 -> 16 limbs, a^2 = 2 * sum_{i<j} a_i a_j B^(i+j) + sum_i a_i^2 B^(2i): 28 + 8 wide multiplies instead of 64.
    The cross products go row by row into two accumulators by the parity of their position (E: even, O: odd), so that
    inside a row every (lo, hi) pair lands on consecutive 64-bit slots of ONE carry chain and fuses into IMAD.WIDE(.X).
    Limbs nobody has written yet are the literal 0."""
    n = 8
    acc = {0: [0] * (2 * n + 2), 1: [0] * (2 * n + 2)}
    for i in range(n - 1):
        for parity in (0, 1):
            A = acc[parity]
            js = [j for j in range(i + 1, n) if (i + j) % 2 == parity]
            if not js:
                continue
            for idx, j in enumerate(js):
                p = i + j
                last = idx == len(js) - 1
                lo, hi = nm.new(), nm.new()
                pr.op("mad.lo.cc.u32" if idx == 0 else "madc.lo.cc.u32", lo, a[i], a[j], A[p])
                # the top slot of a row cannot overflow when nothing was there before (hi(a_i a_j) + carry < 2^32)
                carry_out = (not last) or A[p + 1] != 0
                pr.op("madc.hi.cc.u32" if carry_out else "madc.hi.u32", hi, a[i], a[j], A[p + 1])
                A[p], A[p + 1] = lo, hi
            if carry_out:  # ripple the carry of the row's top slot upwards
                k = i + js[-1] + 2
                while True:
                    c = nm.new()
                    if A[k] == 0:
                        pr.op("addc.u32", c, 0, 0)
                        A[k] = c
                        break
                    pr.op("addc.cc.u32", c, A[k], 0)
                    A[k] = c
                    k += 1
    E, O = acc[0], acc[1]
    assert E[2 * n] == 0 and O[2 * n] == 0 and E[2 * n + 1] == 0 and O[2 * n + 1] == 0
    # T = E + O
    T = [0] * (2 * n)
    started = False
    for k in range(2 * n):
        if not started and (E[k] == 0 or O[k] == 0):
            T[k] = E[k] if O[k] == 0 else O[k]
            continue
        t = nm.new()
        if not started:
            pr.op("add.cc.u32", t, E[k], O[k])
            started = True
        else:
            pr.op("addc.cc.u32" if k < 2 * n - 1 else "addc.u32", t, E[k], O[k])
        T[k] = t
    # D = 2 T
    D = [0] * (2 * n)
    started = False
    for k in range(2 * n):
        if not started and T[k] == 0:
            continue
        d = nm.new()
        if not started:
            pr.op("add.cc.u32", d, T[k], T[k])
            started = True
        else:
            pr.op("addc.cc.u32" if k < 2 * n - 1 else "addc.u32", d, T[k], T[k])
        D[k] = d
    # + the diagonal, one chain of 8 wide multiply-accumulates
    S = nm.new(2 * n)
    for i in range(n):
        pr.op("mad.lo.cc.u32" if i == 0 else "madc.lo.cc.u32", S[2 * i], a[i], a[i], D[2 * i])
        pr.op("madc.hi.cc.u32" if i < n - 1 else "madc.hi.u32", S[2 * i + 1], a[i], a[i], D[2 * i + 1])
    return S


def gen_sqrnr(field: str) -> Prog:
    """r = a^2 / 2^256 mod p for a in [0, 2p): dedicated squaring (36 wide multiplies) + the word-serial reduction,
    result in [0, 2p) like mulnr.  Experimental (accumulate variant 29): 28 wide multiplies less than mulnr for ~45
    more additions -- about -70 cycles per squaring under the measured issue model."""
    mod = FIELDS[field]
    a = [f"a{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_sqrnr", a, r)
    nm = Namer(pr, "t")
    S = emit_sqrw(pr, nm, a)
    emit_redc(pr, nm, S, mod, r, "c", nr=True)
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


def gen_sub2p(field: str) -> Prog:
    """r = a - b (+ 2p if negative) for a, b in [0, 2p): result in [0, 2p), congruent to a - b."""
    mod = FIELDS[field]
    pl = limbs(2 * mod)
    a = [f"a{i}" for i in range(8)]
    b = [f"b{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_sub2p", a + b, r)
    d = [pr.tmp(f"d{i}") for i in range(8)]
    brw = pr.tmp("brw")
    pr.op("sub.cc.u32", d[0], a[0], b[0])
    for i in range(1, 8):
        pr.op("subc.cc.u32", d[i], a[i], b[i])
    pr.op("subc.u32", brw, 0, 0)
    q = [pr.tmp(f"q{i}") for i in range(8)]
    for i in range(8):
        pr.op("and.b32", q[i], brw, pl[i])
    pr.op("add.cc.u32", r[0], d[0], q[0])
    for i in range(1, 7):
        pr.op("addc.cc.u32", r[i], d[i], q[i])
    pr.op("addc.u32", r[7], d[7], q[7])
    return pr


def gen_sub2p_pred(field: str) -> Prog:
    """sub2p with the correction as PREDICATED additions in place instead of masking 2p: 18 instead of 25 instructions
    (experimental, accumulate variant 33).  Works on temporaries: the outputs are written after the last input is read
    (the asm outputs are not early-clobber)."""
    mod = FIELDS[field]
    pl = limbs(2 * mod)
    a = [f"a{i}" for i in range(8)]
    b = [f"b{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_sub2pp", a + b, r)
    d = [pr.tmp(f"d{i}") for i in range(8)]
    brw = pr.tmp("brw")
    neg = pr.pred("neg")
    pr.op("sub.cc.u32", d[0], a[0], b[0])
    for i in range(1, 8):
        pr.op("subc.cc.u32", d[i], a[i], b[i])
    pr.op("subc.u32", brw, 0, 0)
    pr.op("setp.ne.u32", neg, brw, 0)
    pr.op(f"@{neg} add.cc.u32", d[0], d[0], pl[0])
    for i in range(1, 7):
        pr.op(f"@{neg} addc.cc.u32", d[i], d[i], pl[i])
    pr.op(f"@{neg} addc.u32", d[7], d[7], pl[7])
    for i in range(8):
        pr.op("mov.u32", r[i], d[i])
    return pr


def gen_cneg(field: str) -> Prog:
    """r = flag ? p - a : a for a in (0, p]: the conditional negation of a table point's y as 8 PREDICATED subtractions
    (the sign of a signed digit differs lane by lane, so a branch would run both sides anyway).  a = 0 is excluded by the
    caller (the identity is filtered out before); the result lies in [0, p).  Inputs: a0..a7, b0 = flag."""
    mod = FIELDS[field]
    pl = limbs(mod)
    a = [f"a{i}" for i in range(8)]
    r = [f"r{i}" for i in range(8)]
    pr = Prog(f"{field}_cneg", a + ["b0"], r)
    t = [pr.tmp(f"t{i}") for i in range(8)]
    neg = pr.pred("neg")
    for i in range(8):
        pr.op("mov.u32", t[i], a[i])
    pr.op("setp.ne.u32", neg, "b0", 0)
    pr.op(f"@{neg} sub.cc.u32", t[0], pl[0], t[0])
    for i in range(1, 7):
        pr.op(f"@{neg} subc.cc.u32", t[i], pl[i], t[i])
    pr.op(f"@{neg} subc.u32", t[7], pl[7], t[7])
    for i in range(8):
        pr.op("mov.u32", r[i], t[i])
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
    "mulk": gen_mulk,
    "mul2sub": gen_mul2sub,
    "mulnr": lambda f: gen_mul(f, nr=True),
    "mul2subnr": lambda f: gen_mul2sub(f, nr=True),
    "sub2p": gen_sub2p,
    "sqrnr": gen_sqrnr,
    "sub2pp": gen_sub2p_pred,
    "cneg": gen_cneg,
    "add": gen_add,
    "sub": gen_sub,
    "reduce_once": gen_reduce_once,
}


def emulate(field: str, what: str, a: int, b: int | None = None, c: int | None = None, d: int | None = None) -> int:
    pr = ROUTINES[what](field)
    vals = {f"a{i}": l for i, l in enumerate(limbs(a))}
    if b is not None:
        vals.update({f"b{i}": l for i, l in enumerate(limbs(b))})
    if c is not None:
        vals.update({f"c{i}": l for i, l in enumerate(limbs(c))})
        vals.update({f"d{i}": l for i, l in enumerate(limbs(d))})
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
        for what in ("mul", "add", "sub", "reduce_once"):  # "mulk" is kept in the generator only (measured slower)
            pr = ROUTINES[what](field)
            two = what in ("mul", "add", "sub")
            sig = f"{field}_{what}_ptx(uint32_t* r, const uint32_t* a" + (
                ", const uint32_t* b)" if two else ")"
            )
            outs = [f"r[{i}]" for i in range(8)]
            ins = [f"a[{i}]" for i in range(8)] + ([f"b[{i}]" for i in range(8)] if two else [])
            out.append(pr.emit(sig, outs, ins))
    for what in ("mul2sub", "mul2subnr"):
        pr = ROUTINES[what]("fq")
        out.append(pr.emit(f"fq_{what}_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d)",
                           [f"r[{i}]" for i in range(8)],
                           [f"{v}[{i}]" for v in "abcd" for i in range(8)]))
    for what in ("mulnr", "sub2p"):  # the [0, 2p) forms used inside the bucket accumulation
        pr = ROUTINES[what]("fq")
        out.append(pr.emit(f"fq_{what}_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b)",
                           [f"r[{i}]" for i in range(8)], [f"a[{i}]" for i in range(8)] + [f"b[{i}]" for i in range(8)]))
    # bucket accumulation (msm.cu k_accumulate): dedicated squaring, predicated a - b (+ 2p), conditional negation
    pr = ROUTINES["sqrnr"]("fq")
    out.append(pr.emit("fq_sqrnr_ptx(uint32_t* r, const uint32_t* a)", [f"r[{i}]" for i in range(8)], [f"a[{i}]" for i in range(8)]))
    pr = ROUTINES["sub2pp"]("fq")
    out.append(pr.emit("fq_sub2pp_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b)", [f"r[{i}]" for i in range(8)],
                       [f"a[{i}]" for i in range(8)] + [f"b[{i}]" for i in range(8)]))
    pr = ROUTINES["cneg"]("fq")
    out.append(pr.emit("fq_cneg_ptx(uint32_t* r, const uint32_t* a, uint32_t flag)", [f"r[{i}]" for i in range(8)],
                       [f"a[{i}]" for i in range(8)] + ["flag"]))
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
            assert emulate(field, "mulk", x, y) == x * y * rinv % mod, (field, "mulk", x, y)
            assert emulate(field, "add", x, y) == (x + y) % mod, (field, "add", x, y)
            assert emulate(field, "sub", x, y) == (x - y) % mod, (field, "sub", x, y)
        quad = [(w, x, y, z) for w in edge[:6] for x in edge[:6] for y in (0, 1, mod - 1) for z in (0, mod - 1, 2)]
        quad += [tuple(rnd.randrange(mod) for _ in range(4)) for _ in range(iters)]
        for w, x, y, z in quad:
            assert emulate(field, "mul2sub", w, x, y, z) == (w * x - y * z) * rinv % mod, (field, "mul2sub", w, x, y, z)
        # the [0, 2p) forms: congruent results that stay below 2p for ANY operands below 2p
        edge2 = [0, 1, mod - 1, mod, mod + 1, 2 * mod - 1, 2 * mod - 2, (1 << 254), (1 << 254) + 12345]
        pairs2 = [(x, y) for x in edge2 for y in edge2] + [(rnd.randrange(2 * mod), rnd.randrange(2 * mod)) for _ in range(iters)]
        for x, y in pairs2:
            v = emulate(field, "mulnr", x, y)
            assert v < 2 * mod and v % mod == x * y * rinv % mod, (field, "mulnr", x, y)
            v = emulate(field, "sub2p", x, y)
            assert v < 2 * mod and v % mod == (x - y) % mod, (field, "sub2p", x, y)
            v = emulate(field, "sqrnr", x)
            assert v < 2 * mod and v % mod == x * x * rinv % mod, (field, "sqrnr", x)
            assert emulate(field, "sub2pp", x, y) == emulate(field, "sub2p", x, y), (field, "sub2pp", x, y)
        for x in [(1 << 254) - 1, 2 * mod - 1, 0xFFFFFFFF, sum(0xFFFFFFFF << (64 * i) for i in range(4)) >> 2, sum(0x80000000 << (32 * i) for i in range(7)), sum(0xFFFFFFFF << (32 * i) for i in range(7))]:
            v = emulate(field, "sqrnr", x)
            assert x < 2 * mod and v < 2 * mod and v % mod == x * x * rinv % mod, (field, "sqrnr", x)
        quad2 = [(w, x, y, z) for w in edge2[2:7] for x in edge2[2:7] for y in (0, mod, 2 * mod - 1) for z in (0, 2 * mod - 1, mod + 1)]
        quad2 += [tuple(rnd.randrange(2 * mod) for _ in range(4)) for _ in range(iters)]
        for w, x, y, z in quad2:
            v = emulate(field, "mul2subnr", w, x, y, z)
            assert v < 2 * mod and v % mod == (w * x - y * z) * rinv % mod, (field, "mul2subnr", w, x, y, z)
        # multiplicand a < p, word operand b ANY 256-bit value (used for bytes -> Fr)
        for _ in range(iters):
            x, y = rnd.randrange(mod), rnd.randrange(1 << 256)
            assert emulate(field, "mul", x, y) == x * y * rinv % mod, (field, "mulwide", x, y)
            assert emulate(field, "mulk", x, y) == x * y * rinv % mod, (field, "mulkwide", x, y)
        for x in (mod - 1, (1 << 256) % mod, rnd.randrange(mod)):
            for y in ((1 << 256) - 1, (1 << 256) - 2, 1 << 255, mod, 2 * mod, 5 * mod + 7):
                assert emulate(field, "mul", x, y) == x * y * rinv % mod, (field, "mulwide", x, y)
                assert emulate(field, "mulk", x, y) == x * y * rinv % mod, (field, "mulkwide", x, y)
        for x in [1, 2, mod - 1, mod, (mod - 1) // 2, (1 << 32) - 1, 1 << 32, 1 << 224] + [rnd.randrange(1, mod) for _ in range(iters)]:
            assert emulate(field, "cneg", x, 0) == x and emulate(field, "cneg", x, 1) == mod - x and emulate(field, "cneg", x, 1 << 31) == mod - x
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
