// Test-only shim: the FP64-pipe field / curve code (csrc/field_dfma.cuh, csrc/ec_dfma.cuh) compiled for the HOST.
// Every entry point switches the FPU to round-toward-zero, where fma() equals the kernels' __fma_rz.
// Build: g++ -O1 -frounding-math -ffp-contract=off (tests/test_dfma_field.py).
#include <cfenv>
#include "../../rust-kzg-bn254_b200/csrc/ec_dfma.cuh"
using namespace kzgb;
using namespace kzgb::dfma;

namespace {
struct Rz {
    int old;
    Rz() : old(fegetround()) { fesetround(FE_TOWARDZERO); }
    ~Rz() { fesetround(old); }
};
void load(D5& r, const uint64_t* l) { for (int i = 0; i < 5; i++) r.l[i] = (double)l[i]; }
void store(uint64_t* l, const D5& a) { for (int i = 0; i < 5; i++) l[i] = (uint64_t)a.l[i]; }
const double* kp_limbs(int k) {
#define KP(n) static const double kp##n[5] = FQ52_KP_##n;
    KP(0) KP(1) KP(2) KP(3) KP(4) KP(5) KP(6) KP(7) KP(8) KP(9) KP(10) KP(11) KP(12) KP(13) KP(14) KP(15) KP(16)
    KP(17) KP(18) KP(19) KP(20) KP(21) KP(22) KP(23) KP(24) KP(25) KP(26) KP(27) KP(28) KP(29) KP(30) KP(31)
#undef KP
    static const double* all[32] = {kp0, kp1, kp2, kp3, kp4, kp5, kp6, kp7, kp8, kp9, kp10, kp11, kp12, kp13, kp14, kp15,
                                    kp16, kp17, kp18, kp19, kp20, kp21, kp22, kp23, kp24, kp25, kp26, kp27, kp28, kp29, kp30, kp31};
    return all[k];
}
}  // namespace

extern "C" {
// limbs: 5 x u64, each < 2^52
void t5_mul(uint64_t* r, const uint64_t* a, const uint64_t* b) { Rz z; D5 x, y, o; load(x, a); load(y, b); d5_mul(o, x, y); store(r, o); }
void t5_mul2sub(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d, int k) {
    Rz z; D5 x, y, u, v, o; load(x, a); load(y, b); load(u, c); load(v, d); d5_mul2sub(o, x, y, u, v, kp_limbs(k)); store(r, o);
}
void t5_sub(uint64_t* r, const uint64_t* a, const uint64_t* b, int k) { Rz z; D5 x, y, o; load(x, a); load(y, b); d5_sub(o, x, y, kp_limbs(k)); store(r, o); }
void t5_sub2(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* c, int k) {
    Rz z; D5 x, y, u, o; load(x, a); load(y, b); load(u, c); d5_sub2(o, x, y, u, kp_limbs(k)); store(r, o);
}
void t5_add(uint64_t* r, const uint64_t* a, const uint64_t* b) { Rz z; D5 x, y, o; load(x, a); load(y, b); d5_add(o, x, y); store(r, o); }
int t5_is_zero_mod_p(const uint64_t* a) { Rz z; D5 x; load(x, a); return d5_is_zero_mod_p(x) ? 1 : 0; }
void t5_from_u32x8_times16(uint64_t* r, const uint32_t* l) { Rz z; D5 o; d5_from_u32x8_times16(o, l); store(r, o); }
void t5_to_mont256(uint32_t* l, const uint64_t* a) { Rz z; D5 x; load(x, a); d5_to_mont256(l, x); }

// accumulate n affine points (8 x u32 x, y each, radix-2^256 Montgomery; (0,0) = identity) into an XYZZ5 starting
// from the identity; out = canonical XYZZ (4 x 8 x u32); limbs_out (optional) = the 20 raw limbs after the last add
void t5_accumulate(uint32_t* out, const uint32_t* pts, int n, uint64_t* limbs_out) {
    Rz z;
    XYZZ5 acc; xyzz5_set_inf(acc);
    for (int i = 0; i < n; i++) xyzz5_madd(acc, *(const Affine*)(pts + 16 * i));
    if (limbs_out && !acc.inf) { store(limbs_out, acc.x); store(limbs_out + 5, acc.y); store(limbs_out + 10, acc.zz); store(limbs_out + 15, acc.zzz); }
    xyzz5_to_xyzz(*(XYZZ*)out, acc);
}
// the integer formulas on the same input, for comparison of the group element
void t8_accumulate(uint32_t* out, const uint32_t* pts, int n) {
    XYZZ acc; xyzz_set_inf(acc);
    for (int i = 0; i < n; i++) xyzz_madd(acc, *(const Affine*)(pts + 16 * i));
    *(XYZZ*)out = acc;
}
void t8_to_affine(uint32_t* r, const uint32_t* p) { xyzz_to_affine(*(Affine*)r, *(const XYZZ*)p); }
}
