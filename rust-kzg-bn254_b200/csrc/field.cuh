// BN254 Fq / Fr elements: 8 x u32 little-endian limbs in Montgomery form (R = 2^256).
// The in-memory layout is bit-identical to arkworks' Fp<MontBackend<_,4>,4> (4 x u64 LE
// limbs, Montgomery) that the reference hands to its MSM/FFT calls
// (reference primitives/src/arith.rs:4-55, primitives/src/helpers.rs:158).
//
// Device code uses the generated inline-PTX carry chains (field_gen.cuh); the host path
// (finalisation of a single point, challenge handling) is a 4x64 CIOS on unsigned __int128.
#pragma once
#include <stdint.h>
#include <string.h>
#include "field_gen.cuh"

#ifdef __CUDACC__
#define KZ_HD __host__ __device__ __forceinline__
#define KZ_D __device__ __forceinline__
#else
#define KZ_HD inline
#define KZ_D inline
#endif

namespace kzgb {

enum { FQ = 0, FR = 1 };

template <int F>
struct alignas(16) Fe {
    uint32_t l[8];
};
typedef Fe<FQ> Fq;
typedef Fe<FR> Fr;

template <int F> struct FeConst;
template <> struct FeConst<FQ> {
    KZ_HD static void mod(uint32_t* r) { const uint32_t t[8] = FQ_MOD_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    KZ_HD static void one(uint32_t* r) { const uint32_t t[8] = FQ_ONE_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    KZ_HD static void r2(uint32_t* r) { const uint32_t t[8] = FQ_R2_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    KZ_HD static void r3(uint32_t* r) { const uint32_t t[8] = FQ_R3_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    KZ_HD static void pm2(uint32_t* r) { const uint32_t t[8] = FQ_PM2_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    KZ_HD static void half(uint32_t* r) { const uint32_t t[8] = FQ_HALF_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    static constexpr uint64_t NP0 = FQ_NP0_64;
};
template <> struct FeConst<FR> {
    KZ_HD static void mod(uint32_t* r) { const uint32_t t[8] = FR_MOD_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    KZ_HD static void one(uint32_t* r) { const uint32_t t[8] = FR_ONE_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    KZ_HD static void r2(uint32_t* r) { const uint32_t t[8] = FR_R2_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    KZ_HD static void r3(uint32_t* r) { const uint32_t t[8] = FR_R3_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    KZ_HD static void pm2(uint32_t* r) { const uint32_t t[8] = FR_PM2_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    KZ_HD static void half(uint32_t* r) { const uint32_t t[8] = FR_HALF_LIMBS; for (int i = 0; i < 8; i++) r[i] = t[i]; }
    static constexpr uint64_t NP0 = FR_NP0_64;
};

// ---------------------------------------------------------------------------------
// Host reference path (4 x u64 CIOS).  Also used when this header is compiled by g++.
// ---------------------------------------------------------------------------------
#ifndef __CUDA_ARCH__
namespace hostimpl {
typedef unsigned __int128 u128;
template <int F>
inline void load64(uint64_t* d, const uint32_t* s) { memcpy(d, s, 32); }
template <int F>
inline void mod64(uint64_t* m) { uint32_t t[8]; FeConst<F>::mod(t); memcpy(m, t, 32); }
inline bool geq(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
    return true;
}
inline void sub4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 brw = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - b[i] - brw; r[i] = (uint64_t)t; brw = (t >> 64) & 1; }
}
template <int F>
inline void mul(uint32_t* r32, const uint32_t* a32, const uint32_t* b32) {
    uint64_t a[4], b[4], m[4], t[6] = {0, 0, 0, 0, 0, 0};
    memcpy(a, a32, 32); memcpy(b, b32, 32); mod64<F>(m);
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)t[j] + (u128)a[j] * b[i]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t q = t[0] * FeConst<F>::NP0;
        c = (u128)t[0] + (u128)q * m[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)t[j] + (u128)q * m[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    // t < 2p when a < p (b may be any 256-bit value): a few conditional subtractions to be safe
    while (t[4] != 0 || geq(t, m)) {
        u128 brw = 0;
        for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - m[i] - brw; t[i] = (uint64_t)d; brw = (d >> 64) & 1; }
        t[4] -= (uint64_t)brw;
    }
    memcpy(r32, t, 32);
}
template <int F>
inline void add(uint32_t* r32, const uint32_t* a32, const uint32_t* b32) {
    uint64_t a[4], b[4], m[4], t[4];
    memcpy(a, a32, 32); memcpy(b, b32, 32); mod64<F>(m);
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; t[i] = (uint64_t)c; c >>= 64; }
    if (c || geq(t, m)) sub4(t, t, m);
    memcpy(r32, t, 32);
}
template <int F>
inline void sub(uint32_t* r32, const uint32_t* a32, const uint32_t* b32) {
    uint64_t a[4], b[4], m[4], t[4];
    memcpy(a, a32, 32); memcpy(b, b32, 32); mod64<F>(m);
    if (geq(a, b)) sub4(t, a, b);
    else { sub4(t, b, a); sub4(t, m, t); }
    memcpy(r32, t, 32);
}
template <int F>
inline void reduce_once(uint32_t* r32, const uint32_t* a32) {
    uint64_t a[4], m[4];
    memcpy(a, a32, 32); mod64<F>(m);
    if (geq(a, m)) sub4(a, a, m);
    memcpy(r32, a, 32);
}
}  // namespace hostimpl
#endif

// ---------------------------------------------------------------------------------
// Unified ops
// ---------------------------------------------------------------------------------
template <int F> KZ_HD void fe_mul(Fe<F>& r, const Fe<F>& a, const Fe<F>& b) {
#ifdef __CUDA_ARCH__
    if (F == FQ) fq_mul_ptx(r.l, a.l, b.l); else fr_mul_ptx(r.l, a.l, b.l);
#else
    hostimpl::mul<F>(r.l, a.l, b.l);
#endif
}
template <int F> KZ_HD void fe_sqr(Fe<F>& r, const Fe<F>& a) { fe_mul(r, a, a); }
template <int F> KZ_HD void fe_add(Fe<F>& r, const Fe<F>& a, const Fe<F>& b) {
#ifdef __CUDA_ARCH__
    if (F == FQ) fq_add_ptx(r.l, a.l, b.l); else fr_add_ptx(r.l, a.l, b.l);
#else
    hostimpl::add<F>(r.l, a.l, b.l);
#endif
}
template <int F> KZ_HD void fe_sub(Fe<F>& r, const Fe<F>& a, const Fe<F>& b) {
#ifdef __CUDA_ARCH__
    if (F == FQ) fq_sub_ptx(r.l, a.l, b.l); else fr_sub_ptx(r.l, a.l, b.l);
#else
    hostimpl::sub<F>(r.l, a.l, b.l);
#endif
}
// r = a*b - c*d with ONE Montgomery reduction on the device (gen_field.py gen_mul2sub; Fq only)
KZ_HD void fe_mul2sub(Fe<FQ>& r, const Fe<FQ>& a, const Fe<FQ>& b, const Fe<FQ>& c, const Fe<FQ>& d) {
#ifdef __CUDA_ARCH__
    fq_mul2sub_ptx(r.l, a.l, b.l, c.l, d.l);
#else
    Fe<FQ> t, u;
    fe_mul(t, a, b);
    fe_mul(u, c, d);
    fe_sub(r, t, u);
#endif
}
// r = a mod p for a < 2p
template <int F> KZ_HD void fe_reduce_once(Fe<F>& r, const Fe<F>& a) {
#ifdef __CUDA_ARCH__
    if (F == FQ) fq_reduce_once_ptx(r.l, a.l); else fr_reduce_once_ptx(r.l, a.l);
#else
    hostimpl::reduce_once<F>(r.l, a.l);
#endif
}
template <int F> KZ_HD void fe_dbl(Fe<F>& r, const Fe<F>& a) { fe_add(r, a, a); }
template <int F> KZ_HD void fe_zero(Fe<F>& r) { for (int i = 0; i < 8; i++) r.l[i] = 0; }
template <int F> KZ_HD void fe_one(Fe<F>& r) { FeConst<F>::one(r.l); }
template <int F> KZ_HD bool fe_is_zero(const Fe<F>& a) {
    uint32_t o = 0;
    for (int i = 0; i < 8; i++) o |= a.l[i];
    return o == 0;
}
template <int F> KZ_HD bool fe_eq(const Fe<F>& a, const Fe<F>& b) {
    uint32_t o = 0;
    for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i];
    return o == 0;
}
template <int F> KZ_HD void fe_neg(Fe<F>& r, const Fe<F>& a) {
    Fe<F> z; fe_zero(z);
    fe_sub(r, z, a);  // 0 - 0 = 0 stays canonical
}
template <int F> KZ_HD void fe_cneg(Fe<F>& r, const Fe<F>& a, bool neg) {
    Fe<F> n; fe_neg(n, a);
    for (int i = 0; i < 8; i++) r.l[i] = neg ? n.l[i] : a.l[i];
}
// Montgomery <-> canonical.  to_mont accepts ANY 256-bit value (result fully reduced):
// the unreduced operand goes in as the word-by-word operand b of the CIOS.
template <int F> KZ_HD void fe_to_mont(Fe<F>& r, const Fe<F>& a) {
    Fe<F> r2; FeConst<F>::r2(r2.l);
    fe_mul(r, r2, a);
}
template <int F> KZ_HD void fe_from_mont(Fe<F>& r, const Fe<F>& a) {
    Fe<F> o; fe_zero(o); o.l[0] = 1;
    fe_mul(r, a, o);
}
// a^e for a 256-bit exponent given as 8 LE limbs (plain binary, MSB first)
template <int F> KZ_HD void fe_pow(Fe<F>& r, const Fe<F>& a, const uint32_t* e) {
    Fe<F> acc; fe_one(acc);
    bool started = false;
    for (int i = 7; i >= 0; i--) {
        uint32_t w = e[i];
        for (int b = 31; b >= 0; b--) {
            if (started) fe_sqr(acc, acc);
            if ((w >> b) & 1) {
                if (started) fe_mul(acc, acc, a); else { acc = a; started = true; }
            }
        }
    }
    r = acc;
}
// Fermat inverse (0 -> 0)
template <int F> KZ_HD void fe_inv(Fe<F>& r, const Fe<F>& a) {
    uint32_t e[8]; FeConst<F>::pm2(e);
    fe_pow(r, a, e);
}
// ---------------------------------------------------------------------------------
// Fast modular inverse (0 -> 0): Bernstein-Yang "safegcd" division steps in batches of 30 on
// signed 30-bit limbs, ~20 outer iterations of word-sized work instead of the ~380 dependent
// Montgomery multiplications of the Fermat inverse.  Used where ONE inversion is on the critical
// path of a whole kernel level (batch-affine MSM levels, final XYZZ -> affine on the host).
// Input and output in Montgomery form.
// ---------------------------------------------------------------------------------
namespace safegcd {
struct S30 { int32_t v[9]; };
struct T2x2 { int32_t u, v, q, r; };
constexpr int32_t M30 = (int32_t)((1u << 30) - 1u);

KZ_HD void from_limbs(S30& r, const uint32_t* l) {
    // 8 x 32-bit -> 9 x 30-bit
    for (int i = 0; i < 9; i++) {
        int bit = 30 * i, w = bit >> 5, b = bit & 31;
        uint64_t x = l[w];
        if (w + 1 < 8) x |= (uint64_t)l[w + 1] << 32;
        r.v[i] = (int32_t)((x >> b) & (uint32_t)M30);
    }
}
KZ_HD void to_limbs(uint32_t* l, const S30& a) {  // a normalised: every limb in [0, 2^30)
    for (int i = 0; i < 8; i++) l[i] = 0;
    for (int i = 0; i < 9; i++) {
        int bit = 30 * i, w = bit >> 5, b = bit & 31;
        uint64_t x = (uint64_t)(uint32_t)a.v[i] << b;
        l[w] |= (uint32_t)x;
        if (w + 1 < 8) l[w + 1] |= (uint32_t)(x >> 32);
    }
}
// 30 division steps on the low bits of f, g.  eta = -delta.  Returns the new eta and the
// transition matrix t with  2^30 * [f'; g'] = t * [f; g].
KZ_HD int32_t divsteps_30(int32_t eta, uint32_t f0, uint32_t g0, T2x2& t) {
    uint32_t u = 1, v = 0, q = 0, r = 1;
    uint32_t f = f0, g = g0;
    for (int i = 0; i < 30; i++) {
        uint32_t c1 = (uint32_t)(eta >> 31);   // all ones if delta > 0
        uint32_t c2 = (uint32_t)0 - (g & 1u);  // all ones if g odd
        uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
        g += x & c2; q += y & c2; r += z & c2;
        c1 &= c2;
        eta = (int32_t)(((uint32_t)eta ^ c1) - (c1 + 1u));
        f += g & c1; u += q & c1; v += r & c1;
        g >>= 1; u <<= 1; v <<= 1;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return eta;
}
// [f; g] <- t * [f; g] / 2^30 (exact)
KZ_HD void update_fg(S30& f, S30& g, const T2x2& t) {
    const int32_t u = t.u, v = t.v, q = t.q, r = t.r;  // 32 x 32 -> 64 products
#define KZ_MW(a, b) ((int64_t)(a) * (int64_t)(b))
    int64_t cf = KZ_MW(u, f.v[0]) + KZ_MW(v, g.v[0]);
    int64_t cg = KZ_MW(q, f.v[0]) + KZ_MW(r, g.v[0]);
    cf >>= 30; cg >>= 30;
    for (int i = 1; i < 9; i++) {
        cf += KZ_MW(u, f.v[i]) + KZ_MW(v, g.v[i]);
        cg += KZ_MW(q, f.v[i]) + KZ_MW(r, g.v[i]);
        f.v[i - 1] = (int32_t)cf & M30; cf >>= 30;
        g.v[i - 1] = (int32_t)cg & M30; cg >>= 30;
    }
    f.v[8] = (int32_t)cf;
    g.v[8] = (int32_t)cg;
}
// [d; e] <- t * [d; e] / 2^30 mod m, with d, e kept in (-2m, m)
KZ_HD void update_de(S30& d, S30& e, const T2x2& t, const S30& m, uint32_t m_inv30) {
    const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
    int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
    int32_t md = (t.u & sd) + (t.v & se);
    int32_t me = (t.q & sd) + (t.r & se);
    int64_t cd = KZ_MW(u, d.v[0]) + KZ_MW(v, e.v[0]);
    int64_t ce = KZ_MW(q, d.v[0]) + KZ_MW(r, e.v[0]);
    md -= (int32_t)((m_inv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
    me -= (int32_t)((m_inv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
    cd += KZ_MW(m.v[0], md);
    ce += KZ_MW(m.v[0], me);
    cd >>= 30; ce >>= 30;
    for (int i = 1; i < 9; i++) {
        cd += KZ_MW(u, d.v[i]) + KZ_MW(v, e.v[i]) + KZ_MW(m.v[i], md);
        ce += KZ_MW(q, d.v[i]) + KZ_MW(r, e.v[i]) + KZ_MW(m.v[i], me);
        d.v[i - 1] = (int32_t)cd & M30; cd >>= 30;
        e.v[i - 1] = (int32_t)ce & M30; ce >>= 30;
    }
    d.v[8] = (int32_t)cd;
    e.v[8] = (int32_t)ce;
#undef KZ_MW
}
// r in (-2m, m), optionally negated, -> [0, m)
KZ_HD void normalize(S30& r, int32_t sign, const S30& m) {
    int32_t cond_add = r.v[8] >> 31;
    for (int i = 0; i < 9; i++) r.v[i] += m.v[i] & cond_add;
    int32_t cond_neg = sign >> 31;
    for (int i = 0; i < 9; i++) r.v[i] = (r.v[i] ^ cond_neg) - cond_neg;
    for (int i = 0; i < 8; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
    cond_add = r.v[8] >> 31;
    for (int i = 0; i < 9; i++) r.v[i] += m.v[i] & cond_add;
    for (int i = 0; i < 8; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
}
}  // namespace safegcd

template <int F> KZ_HD void fe_inv_fast(Fe<F>& r, const Fe<F>& a) {
    using namespace safegcd;
    if (fe_is_zero(a)) { fe_zero(r); return; }
    uint32_t ml[8]; FeConst<F>::mod(ml);
    S30 m, f, g, d, e;
    from_limbs(m, ml);
    from_limbs(g, a.l);
    f = m;
    for (int i = 0; i < 9; i++) { d.v[i] = 0; e.v[i] = 0; }
    e.v[0] = 1;
    const uint32_t m_inv30 = ((uint32_t)0 - (uint32_t)FeConst<F>::NP0) & (uint32_t)M30;  // m^-1 mod 2^30
    int32_t eta = -1;
    for (int it = 0; it < 25; it++) {  // 25 * 30 = 750 >= 741 division steps suffice for 256-bit inputs
        T2x2 t;
        eta = divsteps_30(eta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        update_de(d, e, t, m, m_inv30);
        update_fg(f, g, t);
        int32_t nz = 0;
        for (int i = 0; i < 9; i++) nz |= g.v[i];
        if (nz == 0) break;
    }
    // f = +-1, d = +-(a^-1) (a in Montgomery form: d = (a R)^-1)
    normalize(d, f.v[8], m);
    Fe<F> y, r3;
    to_limbs(y.l, d);
    FeConst<F>::r3(r3.l);
    fe_mul(r, y, r3);  // (a R)^-1 * R^3 / R = a^-1 R
}

// canonical(a) > (p-1)/2 ?  (reference primitives/src/helpers.rs:151-173)
template <int F> KZ_HD bool fe_lexicographically_largest(const Fe<F>& a_mont) {
    Fe<F> c; fe_from_mont(c, a_mont);
    uint32_t h[8]; FeConst<F>::half(h);
    for (int i = 7; i >= 0; i--) {
        if (c.l[i] != h[i]) return c.l[i] > h[i];
    }
    return false;
}

#ifdef __CUDACC__
// 32-byte vector load/store (two 128-bit transactions)
template <int F> KZ_D Fe<F> fe_load(const Fe<F>* p) {
    Fe<F> r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
template <int F> KZ_D Fe<F> fe_load_ro(const Fe<F>* p) {
    Fe<F> r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
template <int F> KZ_D void fe_store(Fe<F>* p, const Fe<F>& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
#endif

}  // namespace kzgb
