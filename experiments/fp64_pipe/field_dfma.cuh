// BN254 Fq on the FP64 pipe: 5 limbs of 52 bits, each held EXACTLY in a double, Montgomery radix 2^260.
//
// Why: on sm_100 the 32x32->64 multiply (IMAD.WIDE) issues at half the IMAD rate and the bucket
// accumulation already runs at 97 % of that ceiling with only ~36 % of the issue slots used.
// B200 (unlike B300) keeps a full-rate FP64 pipe, which the integer kernels leave idle.  A DFMA returns 53
// bits of a 52x52-bit product, so a limb product is two DFMAs (Emmart/Zheng/Weems, ARITH 2018):
//
//     hi = fma_rz(a, b, 2^104)                = 2^104 + floor(ab / 2^52) * 2^52      (ulp there is 2^52)
//     lo = fma_rz(a, b, (2^104 + 2^52) - hi)  = 2^52 + (ab mod 2^52)                 (exact)
//
// and the mantissa fields of hi / lo ARE the two halves: they are summed per column as raw 64-bit integers
// (the exponent fields add up to a constant known at compile time, subtracted up front).
// A multiplication is 50 limb products + 5 quotient digits = 105 DFMA + 55 DADD + ~250 integer adds/shifts:
// by itself no faster than the IMAD.WIDE form -- the point is that it runs on OTHER pipes (FP64 + ALU), so
// warps using it can share an SM with warps saturating the IMAD.WIDE pipe (k_accumulate_hybrid, msm.cu).
//
// Ranges: limbs are normalised to [0, 2^52) before every product; values are only bounded by 2^260 = 84.6 p,
// so no conditional subtraction exists anywhere: a (x) b = (ab + qp) / 2^260 < ab / 2^260 + p, and a
// subtraction adds a fixed multiple of p (FQ52_KP_k) chosen from the interval analysis in
// tests/test_dfma_field.py.
//
// Everything here is __host__ __device__: the SAME source is compiled by g++ for tests/native/dfma_test.cpp and
// checked against big-int arithmetic with the FPU in round-toward-zero mode (fma() then equals __fma_rz).
#pragma once
#include <stdint.h>
#include <string.h>
#include "field_dfma_consts.cuh"
#ifndef __CUDA_ARCH__
#include <cmath>
#endif

#ifdef __CUDACC__
#define KZD_HD __host__ __device__ __forceinline__
#else
#define KZD_HD inline
#endif

namespace kzgb {
namespace dfma {

struct D5 {
    double l[5];
};

static constexpr uint64_t MASK52 = (1ull << 52) - 1;
static constexpr uint64_t EXP52 = 0x433ull << 52;   // bit pattern of 2^52
static constexpr uint64_t EXP104 = 0x467ull << 52;  // bit pattern of 2^104
#define KZD_TWO52 4503599627370496.0
#define KZD_C1 0x1p104
#define KZD_C2 (0x1p104 + 0x1p52)
#define KZD_MAGIC (0x1p104 + 0x1p103)

KZD_HD double fma_rz(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rz(a, b, c);
#else
    return std::fma(a, b, c);  // host callers run under fesetround(FE_TOWARDZERO)
#endif
}
KZD_HD double add_rz(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rz(a, b);
#else
    return a + b;  // idem (compiled with -frounding-math)
#endif
}
KZD_HD uint64_t bits_of(double x) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
KZD_HD double from_bits(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
// integer in [0, 2^52) <-> the double holding it
KZD_HD double limb_to_double(uint64_t v) { return from_bits(v | EXP52) - KZD_TWO52; }
KZD_HD uint64_t limb_to_int(double x) { return bits_of(x + KZD_TWO52) & MASK52; }

// number of (i, j) pairs of a 5 x 5 limb product with i + j == c
KZD_HD constexpr int pairs_in_column(int c) { return (c < 0 || c > 8) ? 0 : (c <= 4 ? c + 1 : 9 - c); }
// sum of the exponent fields column c collects during `rounds` 5 x 5 limb-product rounds, negated
KZD_HD constexpr uint64_t column_bias(int c, int rounds) {
    return 0ull - ((uint64_t)(rounds * pairs_in_column(c)) * EXP52 + (uint64_t)(rounds * pairs_in_column(c - 1)) * EXP104);
}

// col_lo += (a b) mod 2^52, col_hi += floor(a b / 2^52)   (plus the exponent fields, see column_bias)
KZD_HD void limb_mac(double a, double b, uint64_t& col_hi, uint64_t& col_lo) {
    double hi = fma_rz(a, b, KZD_C1);
    double lo = fma_rz(a, b, KZD_C2 - hi);
    col_hi += bits_of(hi);
    col_lo += bits_of(lo);
}
// ... with the product subtracted
KZD_HD void limb_msc(double a, double b, uint64_t& col_hi, uint64_t& col_lo) {
    double hi = fma_rz(a, b, KZD_C1);
    double lo = fma_rz(a, b, KZD_C2 - hi);
    col_hi -= bits_of(hi);
    col_lo -= bits_of(lo);
}

// Montgomery reduction of the 10 columns c[] (every column complete, biases removed as the terms arrive):
// five quotient digits, then the upper five columns are carried into r
KZD_HD void d5_reduce_columns(D5& r, uint64_t* c) {
    const double p[5] = FQ52_P_LIMBS;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        double t = limb_to_double(c[i] & MASK52);
        double h = fma_rz(t, FQ52_NP, KZD_C1);
        double q = fma_rz(t, FQ52_NP, KZD_C2 - h) - KZD_TWO52;  // (t * -p^-1) mod 2^52
#pragma unroll
        for (int j = 0; j < 5; j++) limb_mac(q, p[j], c[i + j + 1], c[i + j]);
        c[i + 1] += (uint64_t)((int64_t)c[i] >> 52);  // column i is now a multiple of 2^52 (signed: see d5_mul2sub)
    }
    int64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        int64_t t = (int64_t)c[5 + k] + carry;
        r.l[k] = limb_to_double((uint64_t)t & MASK52);
        carry = t >> 52;
    }
}

// r = a b / 2^260 mod p, r < ab / 2^260 + p.  a, b: normalised limbs.  r may alias a or b.
KZD_HD void d5_mul(D5& r, const D5& a, const D5& b) {
    uint64_t c[10];
#pragma unroll
    for (int k = 0; k < 10; k++) c[k] = column_bias(k, 2);
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) limb_mac(a.l[i], b.l[j], c[i + j + 1], c[i + j]);
    d5_reduce_columns(r, c);
}

// r = (a b - c d + kp * 2^260) / 2^260 mod p with ONE reduction (kp = limbs of k p, k p 2^260 >= c d):
// r < ab / 2^260 + (k + 1) p.  The exponent fields of the added and the subtracted products cancel.
KZD_HD void d5_mul2sub(D5& r, const D5& a, const D5& b, const D5& c_, const D5& d, const double* kp) {
    uint64_t c[10];
#pragma unroll
    for (int k = 0; k < 10; k++) c[k] = column_bias(k, 1);  // only the reduction round is left unbalanced
#pragma unroll
    for (int k = 0; k < 5; k++) c[5 + k] += (uint64_t)kp[k];
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            limb_mac(a.l[i], b.l[j], c[i + j + 1], c[i + j]);
            limb_msc(c_.l[i], d.l[j], c[i + j + 1], c[i + j]);
        }
    d5_reduce_columns(r, c);
}

// r = a - b + kp  (kp = limbs of a multiple of p that is >= b), limbs normalised; all FP64, all exact:
// floor(t / 2^52) is add_rz(t, 1.5 * 2^104) - 1.5 * 2^104 because the sum stays positive
KZD_HD void d5_sub(D5& r, const D5& a, const D5& b, const double* kp) {
    double carry = 0.0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        double t = ((a.l[k] + kp[k]) - b.l[k]) + carry;
        double c52 = add_rz(t, KZD_MAGIC) - KZD_MAGIC;
        r.l[k] = t - c52;
        carry = c52 * 0x1p-52;
    }
}
// r = a - b - c + kp
KZD_HD void d5_sub2(D5& r, const D5& a, const D5& b, const D5& c, const double* kp) {
    double carry = 0.0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        double t = (((a.l[k] + kp[k]) - b.l[k]) - c.l[k]) + carry;
        double c52 = add_rz(t, KZD_MAGIC) - KZD_MAGIC;
        r.l[k] = t - c52;
        carry = c52 * 0x1p-52;
    }
}
// r = a + b
KZD_HD void d5_add(D5& r, const D5& a, const D5& b) {
    double carry = 0.0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        double t = (a.l[k] + b.l[k]) + carry;
        double c52 = add_rz(t, KZD_MAGIC) - KZD_MAGIC;
        r.l[k] = t - c52;
        carry = c52 * 0x1p-52;
    }
}

// a == 0 (mod p) for a < FQ52_MAXK p: a = k p forces k = a_0 p^-1 mod 2^52, so one limb product filters
// out all but 2^-47 of the non-multiples and the rare candidate is compared limb by limb
KZD_HD bool d5_is_zero_mod_p(const D5& a) {
    double h = fma_rz(a.l[0], FQ52_PINV, KZD_C1);
    double k = fma_rz(a.l[0], FQ52_PINV, KZD_C2 - h) - KZD_TWO52;
    if (k >= (double)FQ52_MAXK) return false;
    const double p[5] = FQ52_P_LIMBS;
    uint64_t kk = (uint64_t)k, carry = 0;
    bool same = true;
    for (int j = 0; j < 5; j++) {
        uint64_t t = kk * (uint64_t)p[j] + carry;
        same = same && ((t & MASK52) == limb_to_int(a.l[j]));
        carry = t >> 52;
    }
    return same;
}

// 8 x u32 value X < 2^256 (the engine's radix-2^256 Montgomery form) -> limbs of 16 X: for X = x 2^256 mod p that IS
// a representative of x 2^260 mod p, so table points enter this domain by bit shuffling alone
KZD_HD void d5_from_u32x8_times16(D5& r, const uint32_t* l) {
    uint64_t w0 = l[0] | ((uint64_t)l[1] << 32), w1 = l[2] | ((uint64_t)l[3] << 32);
    uint64_t w2 = l[4] | ((uint64_t)l[5] << 32), w3 = l[6] | ((uint64_t)l[7] << 32);
    r.l[0] = limb_to_double((w0 << 4) & MASK52);
    r.l[1] = limb_to_double(((w0 >> 48) | (w1 << 16)) & MASK52);
    r.l[2] = limb_to_double(((w1 >> 36) | (w2 << 28)) & MASK52);
    r.l[3] = limb_to_double(((w2 >> 24) | (w3 << 40)) & MASK52);
    r.l[4] = limb_to_double(w3 >> 12);
}
// normalised limbs of a value < 2^256 -> 8 x u32
KZD_HD void d5_to_u32x8(uint32_t* l, const D5& a) {
    uint64_t L0 = limb_to_int(a.l[0]), L1 = limb_to_int(a.l[1]), L2 = limb_to_int(a.l[2]);
    uint64_t L3 = limb_to_int(a.l[3]), L4 = limb_to_int(a.l[4]);
    uint64_t w0 = L0 | (L1 << 52), w1 = (L1 >> 12) | (L2 << 40), w2 = (L2 >> 24) | (L3 << 28), w3 = (L3 >> 36) | (L4 << 16);
    l[0] = (uint32_t)w0; l[1] = (uint32_t)(w0 >> 32); l[2] = (uint32_t)w1; l[3] = (uint32_t)(w1 >> 32);
    l[4] = (uint32_t)w2; l[5] = (uint32_t)(w2 >> 32); l[6] = (uint32_t)w3; l[7] = (uint32_t)(w3 >> 32);
}
// x 2^260 (any representative below 2^260) -> canonical x 2^256 mod p as 8 x u32: one product with 2^256 mod p
// lands in [0, 2p), one conditional subtraction
KZD_HD void d5_to_mont256(uint32_t* l, const D5& a) {
    const D5 k = {FQ52_TO256_LIMBS};
    const double p[5] = FQ52_P_LIMBS;
    D5 t;
    d5_mul(t, a, k);
    // t - p, keep it unless it went negative
    double d[5], carry = 0.0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        double s = (t.l[j] - p[j]) + carry;
        double c52 = add_rz(s, KZD_MAGIC) - KZD_MAGIC;
        d[j] = s - c52;
        carry = c52 * 0x1p-52;
    }
    if (carry >= 0.0) {
#pragma unroll
        for (int j = 0; j < 5; j++) t.l[j] = d[j];
    }
    d5_to_u32x8(l, t);
}

}  // namespace dfma
}  // namespace kzgb
