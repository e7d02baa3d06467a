// XYZZ bucket accumulator on the FP64 pipe (field_dfma.cuh): the same EFD madd-2008-s formulas as
// xyzz_madd (ec.cuh), coordinates as 5 x 52-bit double limbs in Montgomery radix 2^260.
//
// The affine table stays in the engine's 8 x u32 radix-2^256 form: 16 X is a representative of the same
// residue in radix 2^260, so a gathered point enters by shifts only; the accumulator leaves through one product
// per coordinate (d5_to_mont256) when a bucket run ends.  Ranges (units of p, tests/test_dfma_field.py checks
// them by interval arithmetic and on random chains): X1, Y1 < 16; ZZ, ZZZ < 1.1; 2^260 = 84.6 p.
// Exceptional cases are exact: P == 0 (mod p) is tested with d5_is_zero_mod_p; P + P goes through the integer
// xyzz_dbl_affine on the canonical table point.
#pragma once
#include "ec.cuh"
#include "field_dfma.cuh"

namespace kzgb {
namespace dfma {

struct XYZZ5 {
    D5 x, y, zz, zzz;
    bool inf;
};

KZD_HD void xyzz5_set_inf(XYZZ5& a) { a.inf = true; }

// canonical radix-2^256 XYZZ -> this domain, every coordinate brought below 1.2 p
KZD_HD void xyzz5_from_xyzz(XYZZ5& r, const XYZZ& s) {
    if (xyzz_is_inf(s)) { r.inf = true; return; }
    const D5 one = {FQ52_ONE_LIMBS};
    D5 t;
    d5_from_u32x8_times16(t, s.x.l); d5_mul(r.x, t, one);
    d5_from_u32x8_times16(t, s.y.l); d5_mul(r.y, t, one);
    d5_from_u32x8_times16(t, s.zz.l); d5_mul(r.zz, t, one);
    d5_from_u32x8_times16(t, s.zzz.l); d5_mul(r.zzz, t, one);
    r.inf = false;
}
// ... and back, canonical ([0, p) limbs; identity = all zero)
KZD_HD void xyzz5_to_xyzz(XYZZ& r, const XYZZ5& s) {
    if (s.inf) { xyzz_set_inf(r); return; }
    d5_to_mont256(r.x.l, s.x);
    d5_to_mont256(r.y.l, s.y);
    d5_to_mont256(r.zz.l, s.zz);
    d5_to_mont256(r.zzz.l, s.zzz);
}

// acc += q   (q canonical affine, (0,0) = identity): 10 products, 9 reductions, no conditional subtraction
KZD_HD void xyzz5_madd(XYZZ5& acc, const Affine& q) {
    if (aff_is_inf(q)) return;
    D5 x2, y2;
    d5_from_u32x8_times16(x2, q.x.l);
    d5_from_u32x8_times16(y2, q.y.l);
    if (acc.inf) {
        const D5 one = {FQ52_ONE_LIMBS};
        acc.x = x2; acc.y = y2; acc.zz = one; acc.zzz = one; acc.inf = false;
        return;
    }
    const double p1[5] = FQ52_KP_1, p6[5] = FQ52_KP_6, p11[5] = FQ52_KP_11, p16[5] = FQ52_KP_16;
    D5 Pp, Rr, PP, PPP, Q, t;
    d5_mul(Pp, x2, acc.zz);           // U2
    d5_mul(Rr, y2, acc.zzz);          // S2
    d5_sub(Pp, Pp, acc.x, p16);       // P = U2 - X1      < 17.2 p
    d5_sub(Rr, Rr, acc.y, p16);       // R = S2 - Y1
    if (d5_is_zero_mod_p(Pp)) {
        if (d5_is_zero_mod_p(Rr)) {
            XYZZ d;
            xyzz_dbl_affine(d, q);
            xyzz5_from_xyzz(acc, d);
        } else {
            acc.inf = true;
        }
        return;
    }
    d5_mul(PP, Pp, Pp);               // < 4.5 p
    d5_mul(PPP, Pp, PP);              // < 1.92 p
    d5_mul(Q, acc.x, PP);             // < 1.85 p
    d5_mul(t, Rr, Rr);                // < 4.5 p
    d5_add(x2, Q, Q);                 // 2 Q (x2 is free now)
    d5_sub2(t, t, PPP, x2, p6);       // X3 = R^2 - PPP - 2 Q < 10.5 p
    d5_sub(Q, Q, t, p11);             // Q - X3 < 12.9 p
    d5_mul2sub(acc.y, Rr, Q, acc.y, PPP, p1);  // Y3 = R (Q - X3) - Y1 PPP < 4.7 p, one reduction
    acc.x = t;
    d5_mul(acc.zz, acc.zz, PP);
    d5_mul(acc.zzz, acc.zzz, PPP);
}

}  // namespace dfma
}  // namespace kzgb
