// CPU restatement of the reference's commitment path -- TEST ORACLE and CPU BASELINE only.
// Nothing in the product may link or call this file (see oracle/__init__.py).
//
// "C++ restatement of the arkworks algorithm", NOT arkworks: the reference (Rust, arkworks 0.5,
// rayon) cannot be built in this image.  What is restated, with the reference call sites:
//   * Fp 4x64 Montgomery arithmetic (ark-ff MontBackend; layout corroborated by
//     primitives/src/arith.rs:4-55), inversion by binary extended Euclid as ark-ff does
//   * G1 Jacobian arithmetic and VariableBaseMSM::msm (ark-ec 0.5 `msm_bigint_wnaf`): signed
//     digits, window c = 3 if n < 32 else ceil(log2 n)*69/100 + 2, one bucket array per window,
//     windows processed in parallel (rayon -> std::thread), running-sum, Horner combine
//     (call sites prover/src/kzg.rs:100,121; primitives/src/helpers.rs:332)
//   * radix-2 Fr FFT/IFFT, natural order (ark-poly; primitives/src/polynomial.rs:131-135,242-246)
//   * G1-point IFFT of the SRS, the reference's literal commit_eval_form path (prover/src/kzg.rs:263-285)
//   * to_fr_array, compute_challenge, evaluate_polynomial_in_evaluation_form (n separate
//     inversions), the quotient loop (n more inversions) incl. the z-in-domain case
//     (primitives/src/helpers.rs:40-57,411-535; prover/src/kzg.rs:128-178,237-260)
// Validated against oracle/bn254.py and the reference fixtures in tests/test_oracle_c.py.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>


typedef unsigned __int128 u128;

namespace {

struct Params {
    uint64_t mod[4];
    uint64_t r2[4];   // R^2 mod p
    uint64_t one[4];  // R mod p
    uint64_t np0;     // -p^-1 mod 2^64
};
const Params FQ = {{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                   {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full},
                   {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full},
                   0x87d20782e4866389ull};
const Params FR = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                   {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull},
                   {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full},
                   0xc2e1f593efffffffull};

struct Fp { uint64_t v[4]; };

inline bool geq(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; i--) if (a[i] != b[i]) return a[i] > b[i];
    return true;
}
inline uint64_t sub4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t brw = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - b[i] - brw; r[i] = (uint64_t)t; brw = (uint64_t)(t >> 64) & 1; }
    return brw;
}
inline uint64_t add4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] + b[i] + c; r[i] = (uint64_t)t; c = (uint64_t)(t >> 64); }
    return c;
}
inline bool is_zero(const Fp& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
inline bool eq(const Fp& a, const Fp& b) { return a.v[0] == b.v[0] && a.v[1] == b.v[1] && a.v[2] == b.v[2] && a.v[3] == b.v[3]; }

template <const Params& P>
inline Fp mul(const Fp& a, const Fp& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)t[j] + (u128)a.v[j] * b.v[i]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t q = t[0] * P.np0;
        c = (u128)t[0] + (u128)q * P.mod[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)t[j] + (u128)q * P.mod[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fp r;
    while (t[4] || geq(t, P.mod)) { uint64_t brw = sub4(t, t, P.mod); t[4] -= brw; }
    memcpy(r.v, t, 32);
    return r;
}
template <const Params& P> inline Fp sqr(const Fp& a) { return mul<P>(a, a); }
template <const Params& P> inline Fp add(const Fp& a, const Fp& b) {
    Fp r; uint64_t c = add4(r.v, a.v, b.v);
    if (c || geq(r.v, P.mod)) sub4(r.v, r.v, P.mod);
    return r;
}
template <const Params& P> inline Fp sub(const Fp& a, const Fp& b) {
    Fp r;
    if (sub4(r.v, a.v, b.v)) add4(r.v, r.v, P.mod);
    return r;
}
template <const Params& P> inline Fp neg(const Fp& a) { Fp z = {{0, 0, 0, 0}}; return is_zero(a) ? a : sub<P>(z, a); }
template <const Params& P> inline Fp dbl(const Fp& a) { return add<P>(a, a); }
template <const Params& P> inline Fp one() { Fp r; memcpy(r.v, P.one, 32); return r; }
template <const Params& P> inline Fp to_mont(const Fp& a) { Fp r2; memcpy(r2.v, P.r2, 32); return mul<P>(r2, a); }
template <const Params& P> inline Fp from_mont(const Fp& a) { Fp o = {{1, 0, 0, 0}}; return mul<P>(a, o); }
template <const Params& P> Fp pow_u64(const Fp& a, uint64_t e) {
    Fp r = one<P>(), b = a;
    while (e) { if (e & 1) r = mul<P>(r, b); b = sqr<P>(b); e >>= 1; }
    return r;
}

// Binary extended Euclid on the Montgomery representation (ark-ff `inverse`): returns a^-1 in
// Montgomery form (0 -> 0).
inline void shr1(uint64_t* a) { for (int i = 0; i < 3; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63); a[3] >>= 1; }
inline bool is_one4(const uint64_t* a) { return a[0] == 1 && (a[1] | a[2] | a[3]) == 0; }
// x >>= k for 0 < k < 64
inline void shrk(uint64_t* a, unsigned k) {
    for (int i = 0; i < 3; i++) a[i] = (a[i] >> k) | (a[i + 1] << (64 - k));
    a[3] >>= k;
}
template <const Params& P>
Fp inv(const Fp& a) {
    if (is_zero(a)) return a;
    uint64_t u[4], v[4], b[4], c[4] = {0, 0, 0, 0};
    memcpy(u, a.v, 32); memcpy(v, P.mod, 32); memcpy(b, P.r2, 32);  // b = R^2 so that the result is a^-1 * R
    auto halve = [&](uint64_t* x) {
        if (x[0] & 1) { uint64_t cy = add4(x, x, P.mod); shr1(x); x[3] |= cy << 63; } else shr1(x);
    };
    // same algorithm and the same sequence of values as ark-ff; the runs of halvings of u / v are taken
    // k bits at a time with a count-trailing-zeros (the coefficient still needs one modular halving per bit)
    auto strip = [&](uint64_t* x, uint64_t* coef) {
        while (!(x[0] & 1)) {
            unsigned k = x[0] ? (unsigned)__builtin_ctzll(x[0]) : 63u;
            shrk(x, k);
            for (unsigned i = 0; i < k; i++) halve(coef);
        }
    };
    while (!is_one4(u) && !is_one4(v)) {
        strip(u, b);
        strip(v, c);
        if (geq(u, v)) { sub4(u, u, v); if (sub4(b, b, c)) add4(b, b, P.mod); }
        else { sub4(v, v, u); if (sub4(c, c, b)) add4(c, c, P.mod); }
    }
    Fp r;
    memcpy(r.v, is_one4(u) ? b : c, 32);
    return r;
}

// ------------------------------------------------------------------ G1 (Jacobian, a = 0, b = 3)
struct Aff { Fp x, y; };                       // (0,0) = identity
struct Jac { Fp x, y, z; };                    // z = 0 -> identity
inline bool aff_inf(const Aff& p) { return is_zero(p.x) && is_zero(p.y); }
inline Jac jac_inf() { Jac r; r.x = one<FQ>(); r.y = one<FQ>(); memset(r.z.v, 0, 32); return r; }
inline Jac to_jac(const Aff& p) { if (aff_inf(p)) return jac_inf(); Jac r; r.x = p.x; r.y = p.y; r.z = one<FQ>(); return r; }

Jac jdbl(const Jac& p) {  // dbl-2009-l
    if (is_zero(p.z)) return p;
    Fp A = sqr<FQ>(p.x), B = sqr<FQ>(p.y), C = sqr<FQ>(B);
    Fp t = add<FQ>(p.x, B);
    Fp D = dbl<FQ>(sub<FQ>(sub<FQ>(sqr<FQ>(t), A), C));
    Fp E = add<FQ>(dbl<FQ>(A), A), F = sqr<FQ>(E);
    Jac r;
    r.x = sub<FQ>(F, dbl<FQ>(D));
    Fp c8 = dbl<FQ>(dbl<FQ>(dbl<FQ>(C)));
    r.z = dbl<FQ>(mul<FQ>(p.y, p.z));
    r.y = sub<FQ>(mul<FQ>(E, sub<FQ>(D, r.x)), c8);
    return r;
}
Jac jadd(const Jac& p, const Jac& q) {  // add-2007-bl
    if (is_zero(p.z)) return q;
    if (is_zero(q.z)) return p;
    Fp z1z1 = sqr<FQ>(p.z), z2z2 = sqr<FQ>(q.z);
    Fp u1 = mul<FQ>(p.x, z2z2), u2 = mul<FQ>(q.x, z1z1);
    Fp s1 = mul<FQ>(mul<FQ>(p.y, q.z), z2z2), s2 = mul<FQ>(mul<FQ>(q.y, p.z), z1z1);
    if (eq(u1, u2)) return eq(s1, s2) ? jdbl(p) : jac_inf();
    Fp h = sub<FQ>(u2, u1), i = sqr<FQ>(dbl<FQ>(h)), j = mul<FQ>(h, i);
    Fp rr = dbl<FQ>(sub<FQ>(s2, s1)), v = mul<FQ>(u1, i);
    Jac r;
    r.x = sub<FQ>(sub<FQ>(sqr<FQ>(rr), j), dbl<FQ>(v));
    r.y = sub<FQ>(mul<FQ>(rr, sub<FQ>(v, r.x)), dbl<FQ>(mul<FQ>(s1, j)));
    r.z = mul<FQ>(sub<FQ>(sub<FQ>(sqr<FQ>(add<FQ>(p.z, q.z)), z1z1), z2z2), h);
    return r;
}
Jac jmadd(const Jac& p, const Aff& q) {  // madd-2007-bl
    if (aff_inf(q)) return p;
    if (is_zero(p.z)) return to_jac(q);
    Fp z1z1 = sqr<FQ>(p.z);
    Fp u2 = mul<FQ>(q.x, z1z1), s2 = mul<FQ>(mul<FQ>(q.y, p.z), z1z1);
    if (eq(p.x, u2)) return eq(p.y, s2) ? jdbl(p) : jac_inf();
    Fp h = sub<FQ>(u2, p.x), hh = sqr<FQ>(h), i = dbl<FQ>(dbl<FQ>(hh)), j = mul<FQ>(h, i);
    Fp rr = dbl<FQ>(sub<FQ>(s2, p.y)), v = mul<FQ>(p.x, i);
    Jac r;
    r.x = sub<FQ>(sub<FQ>(sqr<FQ>(rr), j), dbl<FQ>(v));
    r.y = sub<FQ>(mul<FQ>(rr, sub<FQ>(v, r.x)), dbl<FQ>(mul<FQ>(p.y, j)));
    r.z = sub<FQ>(sub<FQ>(sqr<FQ>(add<FQ>(p.z, h)), z1z1), hh);
    return r;
}
Aff to_aff(const Jac& p) {
    Aff r;
    if (is_zero(p.z)) { memset(&r, 0, sizeof r); return r; }
    Fp zi = inv<FQ>(p.z), zi2 = sqr<FQ>(zi);
    r.x = mul<FQ>(p.x, zi2);
    r.y = mul<FQ>(p.y, mul<FQ>(zi2, zi));
    return r;
}
// scalar given as canonical 4x64
Jac jmul(const Jac& p, const uint64_t* k) {
    Jac acc = jac_inf();
    bool started = false;
    for (int i = 3; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            if (started) acc = jdbl(acc);
            if ((k[i] >> b) & 1) { acc = started ? jadd(acc, p) : p; started = true; }
        }
    return acc;
}

template <class F>
void parallel_for(size_t n, int threads, F f) {
    if (threads <= 1 || n <= 1) { for (size_t i = 0; i < n; i++) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    int nt = (int)std::min<size_t>(threads, n);
    for (int t = 0; t < nt; t++)
        th.emplace_back([&]() { for (;;) { size_t i = next.fetch_add(1); if (i >= n) return; f(i); } });
    for (auto& t : th) t.join();
}

// ark-ec 0.5 VariableBaseMSM::msm
Jac msm(const Aff* bases, const Fp* scalars_mont, size_t n, int threads) {
    if (n == 0) return jac_inf();
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    int c = n < 32 ? 3 : (lg * 69 / 100 + 2);
    int W = (254 + c - 1) / c;
    // signed digits, scalar by scalar (make_digits)
    std::vector<int32_t> digits((size_t)n * W);
    parallel_for((n + 4095) / 4096, threads, [&](size_t blk) {
        for (size_t i = blk * 4096; i < std::min(n, (blk + 1) * 4096); i++) {
            Fp s = from_mont<FR>(scalars_mont[i]);
            uint64_t carry = 0;
            for (int w = 0; w < W; w++) {
                int bit = w * c, limb = bit / 64, off = bit % 64;
                uint64_t d = limb < 4 ? s.v[limb] >> off : 0;
                if (off + c > 64 && limb + 1 < 4) d |= s.v[limb + 1] << (64 - off);
                d = (d & (((uint64_t)1 << c) - 1)) + carry;
                int64_t sd = (int64_t)d;
                carry = 0;
                if (d >= ((uint64_t)1 << (c - 1)) && w != W - 1) { sd -= (int64_t)1 << c; carry = 1; }
                digits[i * W + w] = (int32_t)sd;
            }
        }
    });
    std::vector<Jac> wsum(W);
    parallel_for(W, threads, [&](size_t w) {
        std::vector<Jac> buckets((size_t)1 << (c - 1), jac_inf());
        for (size_t i = 0; i < n; i++) {
            int32_t d = digits[i * W + w];
            if (d > 0) buckets[d - 1] = jmadd(buckets[d - 1], bases[i]);
            else if (d < 0) { Aff q = bases[i]; q.y = neg<FQ>(q.y); buckets[-d - 1] = jmadd(buckets[-d - 1], q); }
        }
        Jac run = jac_inf(), tot = jac_inf();
        for (size_t k = buckets.size(); k-- > 0;) { run = jadd(run, buckets[k]); tot = jadd(tot, run); }
        wsum[w] = tot;
    });
    Jac acc = wsum[W - 1];
    for (int w = W - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) acc = jdbl(acc);
        acc = jadd(acc, wsum[w]);
    }
    return acc;
}

// ------------------------------------------------------------------ Fr domain
const uint64_t ROOT28[4] = {0x9bd61b6e725b19f0ull, 0x402d111e41112ed4ull, 0x00e0a7eb8ef62abcull, 0x2a3c09f0a58a7e85ull};
Fp root_of_unity(int k) {  // PRIMITIVE_ROOTS_OF_UNITY[k] (primitives/src/consts.rs:22-52), Montgomery
    Fp w; memcpy(w.v, ROOT28, 32);
    w = to_mont<FR>(w);
    for (int i = 28; i > k; i--) w = sqr<FR>(w);
    return w;
}
int ilog2(size_t n) { int k = 0; while (((size_t)1 << k) < n) k++; return k; }
size_t bitrev(size_t x, int bits) { size_t r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

// in-place radix-2, natural in/out (ark-poly Radix2EvaluationDomain); parallel over butterfly groups
void ntt(Fp* a, size_t n, bool inverse, int threads) {
    int lg = ilog2(n);
    if (n <= 1) return;
    for (size_t i = 0; i < n; i++) { size_t j = bitrev(i, lg); if (j > i) std::swap(a[i], a[j]); }
    Fp w = root_of_unity(lg);
    if (inverse) w = inv<FR>(w);
    std::vector<Fp> tw(n / 2);
    tw[0] = one<FR>();
    for (size_t i = 1; i < n / 2; i++) tw[i] = mul<FR>(tw[i - 1], w);
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t half = len / 2, step = n / len, groups = n / len;
        auto body = [&](size_t gidx) {
            Fp* s = a + gidx * len;
            for (size_t j = 0; j < half; j++) {
                Fp u = s[j], v = mul<FR>(s[j + half], tw[j * step]);
                s[j] = add<FR>(u, v);
                s[j + half] = sub<FR>(u, v);
            }
        };
        if (groups >= (size_t)threads * 4) parallel_for(groups, threads, body);
        else {
            // few large groups: split each group's j-range over threads
            for (size_t gidx = 0; gidx < groups; gidx++) {
                Fp* s = a + gidx * len;
                size_t chunk = (half + threads - 1) / threads;
                parallel_for((half + chunk - 1) / chunk, threads, [&](size_t ci) {
                    for (size_t j = ci * chunk; j < std::min(half, (ci + 1) * chunk); j++) {
                        Fp u = s[j], v = mul<FR>(s[j + half], tw[j * step]);
                        s[j] = add<FR>(u, v);
                        s[j + half] = sub<FR>(u, v);
                    }
                });
            }
        }
    }
    if (inverse) {
        Fp nf = {{(uint64_t)n, 0, 0, 0}};
        Fp ninv = inv<FR>(to_mont<FR>(nf));
        parallel_for((n + 4095) / 4096, threads, [&](size_t blk) {
            for (size_t i = blk * 4096; i < std::min(n, (blk + 1) * 4096); i++) a[i] = mul<FR>(a[i], ninv);
        });
    }
}

// the reference's literal KZG::g1_ifft (prover/src/kzg.rs:263-285): ark-poly ifft over G1Projective
void g1_ifft(const Aff* srs, size_t n, Aff* out, int threads) {
    int lg = ilog2(n);
    std::vector<Jac> a(n);
    for (size_t i = 0; i < n; i++) a[bitrev(i, lg)] = to_jac(srs[i]);
    Fp w = inv<FR>(root_of_unity(lg));
    std::vector<Fp> tw(std::max<size_t>(1, n / 2));
    tw[0] = one<FR>();
    for (size_t i = 1; i < n / 2; i++) tw[i] = mul<FR>(tw[i - 1], w);
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t half = len / 2, step = n / len;
        parallel_for(n / 2, threads, [&](size_t bf) {
            size_t gidx = bf / half, j = bf % half;
            Jac* s = a.data() + gidx * len;
            Fp k = from_mont<FR>(tw[j * step]);
            Jac v = j == 0 ? s[j + half] : jmul(s[j + half], k.v);
            Jac u = s[j], nv = v;
            nv.y = neg<FQ>(nv.y);
            s[j] = jadd(u, v);
            s[j + half] = jadd(u, nv);
        });
    }
    Fp nf = {{(uint64_t)n, 0, 0, 0}};
    Fp ninv = from_mont<FR>(inv<FR>(to_mont<FR>(nf)));
    parallel_for(n, threads, [&](size_t i) { out[i] = to_aff(jmul(a[i], ninv.v)); });  // un-batched into_affine (:280-282)
}

// ------------------------------------------------------------------ blob -> Fr, FS challenge, eval, quotient
Fp fr_from_be(const uint8_t* b) {
    Fp a;
    for (int i = 0; i < 4; i++) {
        uint64_t w = 0;
        for (int k = 0; k < 8; k++) w = (w << 8) | b[8 * (3 - i) + k];
        a.v[i] = w;
    }
    return to_mont<FR>(a);  // reduces any 256-bit value mod r
}
template <const Params& P>
void fp_to_be(const Fp& mont, uint8_t* out) {
    Fp c = from_mont<P>(mont);
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[8 * (3 - i) + k] = (uint8_t)(c.v[i] >> (56 - 8 * k));
}
size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }

std::vector<Fp> to_fr_array_padded(const uint8_t* blob, size_t len) {  // helpers.rs:40-57 + polynomial.rs:41-57
    size_t ne = (len + 31) / 32, n = next_pow2(ne);
    std::vector<Fp> v(n);
    memset(v.data(), 0, n * sizeof(Fp));
    for (size_t i = 0; i < ne; i++) {
        uint8_t ch[32];
        size_t take = std::min<size_t>(32, len - 32 * i);
        memset(ch, 0, 32);
        memcpy(ch, blob + 32 * i, take);
        v[i] = fr_from_be(ch);
    }
    return v;
}
bool lex_largest(const Fp& y_mont) {
    static const uint64_t HALF[4] = {0x9e10460b6c3e7ea3ull, 0xcbc0b548b438e546ull, 0xdc2822db40c0ac2eull, 0x183227397098d014ull};
    Fp c = from_mont<FQ>(y_mont);
    for (int i = 3; i >= 0; i--) if (c.v[i] != HALF[i]) return c.v[i] > HALF[i];
    return false;
}
// SHA-256 (FIPS 180-4), the oracle's OWN plain implementation: the product's sha256.cpp (SHA-NI, AVX-512 multi-buffer) is
// one of the things this oracle checks, so nothing of it is linked here.  Checked against hashlib in tests/test_oracle_c.py.
static inline uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
void oracle_sha256(const uint8_t* data, size_t len, uint8_t out[32]) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t H[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    // the message, a 0x80 byte, zeros up to 56 mod 64, the bit length as a big-endian u64
    const size_t total = ((len + 8) / 64 + 1) * 64;
    uint8_t tail[128];
    const size_t tail_start = len / 64 * 64, tail_len = total - tail_start;
    memset(tail, 0, sizeof tail);
    memcpy(tail, data + tail_start, len - tail_start);
    tail[len - tail_start] = 0x80;
    const uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) tail[tail_len - 1 - i] = (uint8_t)(bits >> (8 * i));
    for (size_t off = 0; off < total; off += 64) {
        const uint8_t* blk = off < tail_start ? data + off : tail + (off - tail_start);
        uint32_t w[64];
        for (int t = 0; t < 16; t++) w[t] = ((uint32_t)blk[4 * t] << 24) | ((uint32_t)blk[4 * t + 1] << 16) | ((uint32_t)blk[4 * t + 2] << 8) | blk[4 * t + 3];
        for (int t = 16; t < 64; t++) {
            const uint32_t s0 = sha_rotr(w[t - 15], 7) ^ sha_rotr(w[t - 15], 18) ^ (w[t - 15] >> 3);
            const uint32_t s1 = sha_rotr(w[t - 2], 17) ^ sha_rotr(w[t - 2], 19) ^ (w[t - 2] >> 10);
            w[t] = w[t - 16] + s0 + w[t - 7] + s1;
        }
        uint32_t v[8];
        memcpy(v, H, sizeof v);
        for (int t = 0; t < 64; t++) {
            const uint32_t S1 = sha_rotr(v[4], 6) ^ sha_rotr(v[4], 11) ^ sha_rotr(v[4], 25);
            const uint32_t ch = (v[4] & v[5]) ^ (~v[4] & v[6]);
            const uint32_t t1 = v[7] + S1 + ch + K[t] + w[t];
            const uint32_t S0 = sha_rotr(v[0], 2) ^ sha_rotr(v[0], 13) ^ sha_rotr(v[0], 22);
            const uint32_t mj = (v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]);
            for (int i = 7; i > 0; i--) v[i] = v[i - 1];
            v[4] += t1;
            v[0] = t1 + S0 + mj;
        }
        for (int i = 0; i < 8; i++) H[i] += v[i];
    }
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(H[i] >> 24); out[4 * i + 1] = (uint8_t)(H[i] >> 16); out[4 * i + 2] = (uint8_t)(H[i] >> 8); out[4 * i + 3] = (uint8_t)H[i]; }
}
void serialize_compressed(const Aff& p, uint8_t out[32]) {  // arkworks layout (helpers.rs:458-460)
    memset(out, 0, 32);
    if (aff_inf(p)) { out[31] = 0x40; return; }
    Fp c = from_mont<FQ>(p.x);
    memcpy(out, c.v, 32);
    if (lex_largest(p.y)) out[31] |= 0x80;
}
Fp compute_challenge(const std::vector<Fp>& evals, const Aff& commitment) {  // helpers.rs:411-472
    size_t n = evals.size();
    std::vector<uint8_t> buf(24 + 8 + 32 * n + 32);
    memcpy(buf.data(), "EIGENDA_FSBLOBVERIFY_V1_", 24);
    for (int i = 0; i < 8; i++) buf[24 + i] = (uint8_t)((uint64_t)n >> (56 - 8 * i));
    for (size_t i = 0; i < n; i++) fp_to_be<FR>(evals[i], &buf[32 + 32 * i]);  // to_byte_array round trip (:448)
    serialize_compressed(commitment, &buf[32 + 32 * n]);
    uint8_t dg[32];
    oracle_sha256(buf.data(), buf.size(), dg);
    return fr_from_be(dg);
}
std::vector<Fp> roots_of_unity(size_t n) {  // helpers.rs:553-610
    std::vector<Fp> r(n);
    Fp w = root_of_unity(ilog2(n));
    r[0] = one<FR>();
    for (size_t i = 1; i < n; i++) r[i] = mul<FR>(r[i - 1], w);
    return r;
}
Fp evaluate(const std::vector<Fp>& f, const Fp& z) {  // helpers.rs:475-535: n separate inversions
    size_t n = f.size();
    std::vector<Fp> roots = roots_of_unity(n);
    for (size_t i = 0; i < n; i++) if (eq(roots[i], z)) return f[i];
    Fp sum; memset(sum.v, 0, 32);
    for (size_t i = 0; i < n; i++) {
        Fp a = mul<FR>(f[i], roots[i]);
        Fp b = sub<FR>(z, roots[i]);
        sum = add<FR>(sum, mul<FR>(a, inv<FR>(b)));
    }
    Fp r = sub<FR>(pow_u64<FR>(z, n), one<FR>());
    Fp nf = {{(uint64_t)n, 0, 0, 0}};
    Fp ninv = inv<FR>(to_mont<FR>(nf));
    return mul<FR>(mul<FR>(sum, r), ninv);
}
std::vector<Fp> quotient(const std::vector<Fp>& f, const Fp& z, const Fp& y) {  // kzg.rs:141-174, 237-260
    size_t n = f.size();
    std::vector<Fp> roots = roots_of_unity(n), q(n);
    for (size_t i = 0; i < n; i++) {
        Fp den = sub<FR>(roots[i], z);
        if (is_zero(den)) {
            Fp acc; memset(acc.v, 0, 32);
            for (size_t k = 0; k < n; k++) {
                if (eq(roots[k], z)) continue;
                Fp fi = sub<FR>(f[k], y);
                Fp num = mul<FR>(fi, roots[k]);
                Fp d = mul<FR>(sub<FR>(z, roots[k]), z);
                acc = add<FR>(acc, mul<FR>(num, inv<FR>(d)));
            }
            q[i] = acc;
        } else {
            q[i] = mul<FR>(sub<FR>(f[i], y), inv<FR>(den));
        }
    }
    return q;
}
Aff commit_evals(const std::vector<Fp>& evals, const Aff* srs, int threads, bool literal) {
    size_t n = evals.size();
    if (literal) {  // kzg.rs:98-100: G1 IFFT of the SRS, then MSM over the Lagrange bases
        std::vector<Aff> lag(n);
        g1_ifft(srs, n, lag.data(), threads);
        return to_aff(msm(lag.data(), evals.data(), n, threads));
    }
    std::vector<Fp> c = evals;
    ntt(c.data(), n, true, threads);
    return to_aff(msm(srs, c.data(), n, threads));
}

}  // namespace

extern "C" {

// bases: n x (x||y) Montgomery, (0,0) = identity; scalars Montgomery; out (x||y) Montgomery
void ref_msm(const uint64_t* bases, const uint64_t* scalars, size_t n, int threads, uint64_t out[8]) {
    Aff r = to_aff(msm((const Aff*)bases, (const Fp*)scalars, n, threads));
    memcpy(out, &r, 64);
}
void ref_ntt(uint64_t* data, size_t n, int inverse, int threads) { ntt((Fp*)data, n, inverse != 0, threads); }
void ref_g1_ifft(const uint64_t* srs, size_t n, int threads, uint64_t* out) { g1_ifft((const Aff*)srs, n, (Aff*)out, threads); }
void ref_to_fr_array(const uint8_t* blob, size_t len, uint64_t* out) {
    std::vector<Fp> v = to_fr_array_padded(blob, len);
    memcpy(out, v.data(), v.size() * 32);
}
// commit_blob (kzg.rs:182-185).  literal != 0 follows the reference literally (G1 IFFT per commit).
void ref_commit_blob(const uint8_t* blob, size_t len, const uint64_t* srs, int threads, int literal, uint64_t out[8]) {
    std::vector<Fp> ev = to_fr_array_padded(blob, len);
    Aff c = commit_evals(ev, (const Aff*)srs, threads, literal != 0);
    memcpy(out, &c, 64);
}
// compute_blob_proof (kzg.rs:288-309)
void ref_blob_proof(const uint8_t* blob, size_t len, const uint64_t commitment[8], const uint64_t* srs, int threads,
                    int literal, uint64_t out[8]) {
    std::vector<Fp> ev = to_fr_array_padded(blob, len);
    Aff c;
    memcpy(&c, commitment, 64);
    Fp z = compute_challenge(ev, c);
    Fp y = evaluate(ev, z);
    std::vector<Fp> q = quotient(ev, z, y);
    Aff p = commit_evals(q, (const Aff*)srs, threads, literal != 0);
    memcpy(out, &p, 64);
}
// compute_proof with caller-supplied z (kzg.rs:215-234); y returned too
void ref_proof_at(const uint64_t* evals, size_t n, const uint64_t z[4], const uint64_t* srs, int threads, uint64_t out[8],
                  uint64_t y_out[4]) {
    std::vector<Fp> ev((const Fp*)evals, (const Fp*)evals + n);
    Fp zz; memcpy(zz.v, z, 32);
    Fp y = evaluate(ev, zz);
    std::vector<Fp> q = quotient(ev, zz, y);
    Aff p = commit_evals(q, (const Aff*)srs, threads, false);
    memcpy(out, &p, 64);
    memcpy(y_out, y.v, 32);
}
void ref_fr_inv(const uint64_t a[4], uint64_t out[4]) { Fp x; memcpy(x.v, a, 32); Fp r = inv<FR>(x); memcpy(out, r.v, 32); }
void ref_fq_inv(const uint64_t a[4], uint64_t out[4]) { Fp x; memcpy(x.v, a, 32); Fp r = inv<FQ>(x); memcpy(out, r.v, 32); }
// evaluate_polynomial_in_evaluation_form (helpers.rs:475-535) and compute_challenge (helpers.rs:411-472) alone
void ref_evaluate(const uint64_t* evals, size_t n, const uint64_t z[4], uint64_t y_out[4]) {
    std::vector<Fp> ev((const Fp*)evals, (const Fp*)evals + n);
    Fp zz; memcpy(zz.v, z, 32);
    Fp y = evaluate(ev, zz);
    memcpy(y_out, y.v, 32);
}
void ref_challenge(const uint8_t* blob, size_t len, const uint64_t commitment[8], uint64_t z_out[4]) {
    std::vector<Fp> ev = to_fr_array_padded(blob, len);
    Aff c; memcpy(&c, commitment, 64);
    Fp z = compute_challenge(ev, c);
    memcpy(z_out, z.v, 32);
}
// read_g1_point_from_bytes_be (helpers.rs:175-226) for n points, `threads` workers as prover/src/srs.rs:86-103;
// returns 0, or 1 + the index of the first point that is not on the curve
size_t ref_srs_decompress(const uint8_t* bytes, size_t n, int threads, uint64_t* out_xy) {
    static const uint64_t SQRT_EXP[4] = {0x4f082305b61f3f52ull, 0x65e05aa45a1c72a3ull, 0x6e14116da0605617ull, 0x0c19139cb84c680aull};  // (p+1)/4
    std::atomic<size_t> bad{0};
    Aff* out = (Aff*)out_xy;
    parallel_for(n, threads, [&](size_t i) {
        const uint8_t* b = bytes + 32 * i;
        uint8_t flag = b[0] & 0xC0;
        if (flag == 0x40) { memset(&out[i], 0, 64); return; }
        uint8_t xb[32]; memcpy(xb, b, 32); xb[0] &= 0x3F;
        Fp xc;
        for (int k = 0; k < 4; k++) { uint64_t w = 0; for (int j = 0; j < 8; j++) w = (w << 8) | xb[8 * (3 - k) + j]; xc.v[k] = w; }
        Fp x = to_mont<FQ>(xc);
        Fp three = add<FQ>(add<FQ>(one<FQ>(), one<FQ>()), one<FQ>());
        Fp rhs = add<FQ>(mul<FQ>(sqr<FQ>(x), x), three);
        Fp y = one<FQ>();
        for (int k = 3; k >= 0; k--) for (int bit = 63; bit >= 0; bit--) { y = sqr<FQ>(y); if ((SQRT_EXP[k] >> bit) & 1) y = mul<FQ>(y, rhs); }
        if (!eq(sqr<FQ>(y), rhs)) { size_t cur = bad.load(); while ((cur == 0 || i + 1 < cur) && !bad.compare_exchange_weak(cur, i + 1)) {} return; }
        bool largest = lex_largest(y);
        if ((flag == 0xC0) != largest && flag != 0) y = neg<FQ>(y);
        out[i].x = x; out[i].y = y;
    });
    return bad.load();
}
int ref_hw_threads() { return (int)std::thread::hardware_concurrency(); }
void ref_sha256(const uint8_t* data, size_t len, uint8_t out[32]) { oracle_sha256(data, len, out); }

}  // extern "C"
