// SRS ingest: decompression of the gnark big-endian g1.point format and the fixed-base
// window tables that stay resident in HBM.
//   read_g1_point_from_bytes_be   reference primitives/src/helpers.rs:175-226
//   lexicographically_largest     primitives/src/helpers.rs:151-173
//   SRS::new / loader             prover/src/srs.rs:35-188 (one syscall + one sqrt per point on CPU threads)
#include "kzgb_internal.hpp"

namespace kzgb {

__device__ __forceinline__ uint32_t bswap32s(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// err[0] (preset 0xffffffff): ((1 + index of the lowest bad point) << 2) | kind, kind 1 = not on curve, 2 = bad infinity
__global__ void __launch_bounds__(128) k_decompress(const uint8_t* __restrict__ in, uint32_t n, Affine* __restrict__ out,
                                                     uint32_t* __restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(in + (size_t)i * 32);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    uint32_t be[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    Fq x;
    for (int k = 0; k < 8; k++) x.l[7 - k] = bswap32s(be[k]);
    uint32_t flag = x.l[7] >> 30;  // top two bits of byte 0
    x.l[7] &= 0x3fffffffu;
    Affine P;
    if (flag == 1u) {  // 0b01: infinity, remaining bits must be zero (helpers.rs:188-196)
        if (!fe_is_zero(x)) atomicMin(&err[0], ((i + 1) << 2) | 2u);
        aff_set_inf(P);
        aff_store(&out[i], P);
        return;
    }
    fe_to_mont(x, x);  // reduces mod p (from_be_bytes_mod_order, helpers.rs:198-201)
    Fq y2, y, t, three;
    fe_sqr(y2, x); fe_mul(y2, y2, x);
    fe_one(t); fe_dbl(three, t); fe_add(three, three, t);
    fe_add(y2, y2, three);
    const uint32_t e[8] = FQ_SQRT_EXP_LIMBS;
    fe_pow(y, y2, e);  // p = 3 mod 4
    fe_sqr(t, y);
    if (!fe_eq(t, y2)) {
        atomicMin(&err[0], ((i + 1) << 2) | 1u);
        aff_set_inf(P);
        aff_store(&out[i], P);
        return;
    }
    bool largest = fe_lexicographically_largest(y);
    // 0b10: keep the smaller root, 0b11: keep the larger; 0b00 keeps whatever sqrt gave (helpers.rs:210-216)
    if (largest) { if (flag == 2u) fe_neg(y, y); }
    else if (flag == 3u) fe_neg(y, y);
    P.x = x; P.y = y;
    aff_store(&out[i], P);
}

// scratch layout per (w, j): XYZZ point then one Fq prefix product
struct alignas(16) PreEntry { XYZZ p; Fq pref; };

__global__ void __launch_bounds__(128) k_precompute(Affine* __restrict__ table, uint32_t first, uint32_t count,
                                                     uint32_t stride, int c, int W, PreEntry* __restrict__ scratch,
                                                     uint32_t batch) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    uint32_t i = first + j;
    Affine P = aff_load_ro(&table[i]);
    XYZZ Q; xyzz_from_affine(Q, P);
    Fq run; fe_one(run);
    for (int w = 1; w < W; w++) {
        for (int k = 0; k < c; k++) xyzz_dbl(Q, Q);
        PreEntry* e = &scratch[(size_t)(w - 1) * batch + j];
        xyzz_store(&e->p, Q);
        if (!xyzz_is_inf(Q)) fe_mul(run, run, Q.zzz);
        fe_store(&e->pref, run);
    }
    Fq inv; fe_inv(inv, run);
    for (int w = W - 1; w >= 1; w--) {
        PreEntry* e = &scratch[(size_t)(w - 1) * batch + j];
        XYZZ R = xyzz_load(&e->p);
        Affine A;
        if (xyzz_is_inf(R)) {
            aff_set_inf(A);
        } else {
            Fq prev;
            if (w >= 2) prev = fe_load(&scratch[(size_t)(w - 2) * batch + j].pref); else fe_one(prev);
            Fq iz; fe_mul(iz, inv, prev);      // 1 / zzz_w
            fe_mul(inv, inv, R.zzz);
            xyzz_to_affine_with_inv(A, R, iz);
        }
        aff_store(&table[(size_t)w * stride + i], A);
    }
}

__global__ void __launch_bounds__(128) k_validate(const Affine* __restrict__ pts, uint32_t n, uint32_t* __restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine P = aff_load_ro(&pts[i]);
    if (!aff_on_curve(P)) atomicMin(&err[0], i + 1);
}

// out[i] = tau^i * G
__global__ void __launch_bounds__(128) k_synthetic(Affine* __restrict__ out, uint32_t n, Fr tau, uint32_t first) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t e[8] = {first + i, 0, 0, 0, 0, 0, 0, 0};
    Fr s;
    fe_pow(s, tau, e);
    fe_from_mont(s, s);
    Affine G;
    fe_one(G.x);
    fe_dbl(G.y, G.x);
    XYZZ acc; xyzz_set_inf(acc);
    for (int k = 7; k >= 0; k--) {
        for (int b = 31; b >= 0; b--) {
            xyzz_dbl(acc, acc);
            if ((s.l[k] >> b) & 1u) xyzz_madd(acc, G);
        }
    }
    Affine A;
    xyzz_to_affine(A, acc);
    aff_store(&out[i], A);
}

void srs_decompress_launch(const uint8_t* in_be, uint32_t n, Affine* out, uint32_t* err, cudaStream_t st) {
    if (!n) return;
    k_decompress<<<(n + 127) / 128, 128, 0, st>>>(in_be, n, out, err);
    g_launch_count++;
}

size_t srs_precompute_scratch_bytes(uint32_t batch, int W) {
    return (size_t)batch * (size_t)(W > 1 ? W - 1 : 0) * sizeof(PreEntry);
}

void srs_precompute_launch(Affine* table, uint32_t n, uint32_t stride, int c, int W, XYZZ* scratch, uint32_t batch,
                           cudaStream_t st) {
    if (W <= 1) return;
    for (uint32_t first = 0; first < n; first += batch) {
        uint32_t count = (n - first) < batch ? (n - first) : batch;
        k_precompute<<<(count + 127) / 128, 128, 0, st>>>(table, first, count, stride, c, W,
                                                          reinterpret_cast<PreEntry*>(scratch), batch);
        g_launch_count++;
    }
}

void g1_validate_launch(const Affine* pts, uint32_t n, uint32_t* err, cudaStream_t st) {
    if (!n) return;
    k_validate<<<(n + 127) / 128, 128, 0, st>>>(pts, n, err);
    g_launch_count++;
}

void srs_synthetic_launch(Affine* out, uint32_t n, const Fr* tau_mont_host, uint32_t first, cudaStream_t st) {
    if (!n) return;
    k_synthetic<<<(n + 127) / 128, 128, 0, st>>>(out, n, *tau_mont_host, first);
    g_launch_count++;
}

}  // namespace kzgb
