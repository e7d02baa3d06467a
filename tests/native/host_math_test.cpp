// Test-only shim: exposes the HOST path of csrc/field.cuh + csrc/ec.cuh + sha256 to Python so the
// exact formulas the kernels use (same source, compiled for the host) can be checked on CPU.
#include "../../rust-kzg-bn254_b200/csrc/ec.cuh"
#include "../../rust-kzg-bn254_b200/csrc/sha256.hpp"
using namespace kzgb;
extern "C" {
void t_fq_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) { fe_mul(*(Fq*)r, *(const Fq*)a, *(const Fq*)b); }
void t_fr_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) { fe_mul(*(Fr*)r, *(const Fr*)a, *(const Fr*)b); }
void t_fq_add(uint32_t* r, const uint32_t* a, const uint32_t* b) { fe_add(*(Fq*)r, *(const Fq*)a, *(const Fq*)b); }
void t_fq_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) { fe_sub(*(Fq*)r, *(const Fq*)a, *(const Fq*)b); }
void t_fr_add(uint32_t* r, const uint32_t* a, const uint32_t* b) { fe_add(*(Fr*)r, *(const Fr*)a, *(const Fr*)b); }
void t_fr_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) { fe_sub(*(Fr*)r, *(const Fr*)a, *(const Fr*)b); }
void t_fq_inv(uint32_t* r, const uint32_t* a) { fe_inv(*(Fq*)r, *(const Fq*)a); }
void t_fr_inv(uint32_t* r, const uint32_t* a) { fe_inv(*(Fr*)r, *(const Fr*)a); }
void t_fq_inv_fast(uint32_t* r, const uint32_t* a) { fe_inv_fast(*(Fq*)r, *(const Fq*)a); }
void t_fr_inv_fast(uint32_t* r, const uint32_t* a) { fe_inv_fast(*(Fr*)r, *(const Fr*)a); }
void t_fq_to_mont(uint32_t* r, const uint32_t* a) { fe_to_mont(*(Fq*)r, *(const Fq*)a); }
void t_fr_to_mont(uint32_t* r, const uint32_t* a) { fe_to_mont(*(Fr*)r, *(const Fr*)a); }
void t_fq_from_mont(uint32_t* r, const uint32_t* a) { fe_from_mont(*(Fq*)r, *(const Fq*)a); }
int t_fq_lex_largest(const uint32_t* a) { return fe_lexicographically_largest(*(const Fq*)a) ? 1 : 0; }
// XYZZ ops on (x,y,zz,zzz) Montgomery; affine (x,y) Montgomery, (0,0) = identity
void t_madd(uint32_t* acc, const uint32_t* q) { xyzz_madd(*(XYZZ*)acc, *(const Affine*)q); }
void t_add(uint32_t* acc, const uint32_t* q) { xyzz_add(*(XYZZ*)acc, *(const XYZZ*)q); }
void t_dbl(uint32_t* r, const uint32_t* p) { XYZZ t; xyzz_dbl(t, *(const XYZZ*)p); *(XYZZ*)r = t; }
void t_to_affine(uint32_t* r, const uint32_t* p) { xyzz_to_affine(*(Affine*)r, *(const XYZZ*)p); }
void t_mul_small(uint32_t* r, const uint32_t* p, uint32_t k) { xyzz_mul_small(*(XYZZ*)r, *(const XYZZ*)p, k); }
int t_on_curve(const uint32_t* p) { return aff_on_curve(*(const Affine*)p) ? 1 : 0; }
void t_sha256(const uint8_t* d, size_t n, uint8_t* out) { sha256(d, n, out); }
void t_sha256_chunked(const uint8_t* d, size_t n, size_t step, uint8_t* out) {
    Sha256 s;
    for (size_t i = 0; i < n; i += step) s.update(d + i, (n - i) < step ? (n - i) : step);
    s.finish(out);
}
int t_has_shani() { return sha256_has_shani() ? 1 : 0; }
}
