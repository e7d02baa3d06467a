// FP64-pipe experiments (field_dfma.cuh / ec_dfma.cuh): rate microbenchmarks and an on-device self-test.
// Self-contained (no engine context): the entry points take a device ordinal, allocate what they need and
// return.  Nothing on the product path calls into this file; the accumulate variants that use the same
// arithmetic live in msm.cu (accumulate variants 23-26) and are opt-in (KZGB_ACC_VARIANT / kzgb_set_option).
//
//   kind 0  independent DFMA chains                         -> DFMA/s (B200: expected 64 /clk/SM)
//   kind 1  d5_mul on every warp                            -> Fq multiplications/s on the FP64 + ALU pipes
//   kind 2  hybrid: half of the warps of every scheduler run the integer fe_mul (IMAD.WIDE pipe), the other
//           half d5_mul (FP64 pipe); iters_int / iters_dfma multiplications per thread -> total Fq-mul/s
//   kind 3  the integer fe_mul on every warp, same harness   -> baseline for kinds 1 and 2
#include <cuda_runtime.h>
#include "../../include/kzg_bn254_b200.h"
#include "ec_dfma.cuh"
#include "kzgb_internal.hpp"

namespace kzgb {
using namespace dfma;

__global__ void __launch_bounds__(256) k_dfma_peak(double* __restrict__ sink, int iters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    double a = 1.0 + t * 0x1p-30, b = 1.0 - t * 0x1p-31;
    double x0 = a, x1 = b, x2 = a + 1, x3 = b + 1, x4 = a + 2, x5 = b + 2, x6 = a + 3, x7 = b + 3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = __fma_rz(x0, a, b); x1 = __fma_rz(x1, a, b); x2 = __fma_rz(x2, a, b); x3 = __fma_rz(x3, a, b);
            x4 = __fma_rz(x4, a, b); x5 = __fma_rz(x5, a, b); x6 = __fma_rz(x6, a, b); x7 = __fma_rz(x7, a, b);
        }
    }
    sink[t] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// Do DFMA and IMAD.WIDE share an execution pipe?  Warps with role 0 run IMAD.WIDE chains, role 1 DFMA chains,
// role 2 DADD chains, role 3 exit at once.  MIX selects the roles of the two warp halves of every scheduler
// (warps w and w + 4 of a 256-thread block share one): (lo half, hi half).
__device__ __forceinline__ uint32_t run_wide_chain(uint32_t t, int iters) {
    uint32_t a = t * 2654435761u + 12345u, b = t ^ 0x9e3779b9u;
    uint32_t x0 = a, x1 = b, x2 = a + 1, x3 = b + 1, x4 = a + 2, x5 = b + 2, x6 = a + 3, x7 = b + 3;
#define KZ_MW(dst, src, m)                                                                                  \
    asm volatile("{ .reg .u64 t; .reg .u32 lo, hi; mul.wide.u32 t, %1, %2; mov.b64 {lo, hi}, t; "            \
                 "lop3.b32 %0, %0, lo, hi, 0x96; }" : "+r"(dst) : "r"(src), "r"(m))
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            KZ_MW(x0, x1, b); KZ_MW(x2, x3, b); KZ_MW(x4, x5, b); KZ_MW(x6, x7, b);
            KZ_MW(x1, x2, a); KZ_MW(x3, x4, a); KZ_MW(x5, x6, a); KZ_MW(x7, x0, a);
        }
    }
#undef KZ_MW
    return x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}
template <bool ADD>
__device__ __forceinline__ uint32_t run_fp64_chain(uint32_t t, int iters) {
    double a = 1.0 + t * 0x1p-30, b = 1.0 - t * 0x1p-31;
    double x0 = a, x1 = b, x2 = a + 1, x3 = b + 1, x4 = a + 2, x5 = b + 2, x6 = a + 3, x7 = b + 3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (ADD) {
                x0 = __dadd_rz(x0, a); x1 = __dadd_rz(x1, b); x2 = __dadd_rz(x2, a); x3 = __dadd_rz(x3, b);
                x4 = __dadd_rz(x4, a); x5 = __dadd_rz(x5, b); x6 = __dadd_rz(x6, a); x7 = __dadd_rz(x7, b);
            } else {
                x0 = __fma_rz(x0, a, b); x1 = __fma_rz(x1, a, b); x2 = __fma_rz(x2, a, b); x3 = __fma_rz(x3, a, b);
                x4 = __fma_rz(x4, a, b); x5 = __fma_rz(x5, a, b); x6 = __fma_rz(x6, a, b); x7 = __fma_rz(x7, a, b);
            }
        }
    }
    return (uint32_t)__double2ll_rz((x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7) * 0x1p-20);
}
template <int ROLE_LO, int ROLE_HI>
__global__ void __launch_bounds__(256) k_pipe_mix(uint32_t* __restrict__ sink, int iters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    int role = (threadIdx.x >> 7) & 1 ? ROLE_HI : ROLE_LO;
    if (role == 3) return;
    sink[t] = role == 0 ? run_wide_chain(t, iters) : role == 1 ? run_fp64_chain<false>(t, iters) : run_fp64_chain<true>(t, iters);
}

__device__ __forceinline__ void seed_fq(Fq& a, uint32_t t, uint32_t s) {
    for (int k = 0; k < 8; k++) a.l[k] = (t + s) * 2654435761u + 40503u * k * (s + 1);
    a.l[7] &= 0x0fffffffu;  // < 2^252 < p
}

// two independent chains per thread, as k_fqmul_peak (microbench.cu)
__device__ __forceinline__ uint32_t run_int_chain(uint32_t t, int iters) {
    Fq a, b, c, d;
    seed_fq(a, t, 1); seed_fq(b, t, 2); seed_fq(c, t, 3); seed_fq(d, t, 4);
    for (int i = 0; i < iters; i++) {
        fe_mul(a, a, b);
        fe_mul(c, c, d);
        fe_mul(b, b, a);
        fe_mul(d, d, c);
    }
    uint32_t x = 0;
    for (int k = 0; k < 8; k++) x ^= a.l[k] ^ b.l[k] ^ c.l[k] ^ d.l[k];
    return x;
}
__device__ __forceinline__ uint32_t run_dfma_chain(uint32_t t, int iters) {
    Fq s;
    D5 a, b, c, d;
    seed_fq(s, t, 1); d5_from_u32x8_times16(a, s.l);
    seed_fq(s, t, 2); d5_from_u32x8_times16(b, s.l);
    seed_fq(s, t, 3); d5_from_u32x8_times16(c, s.l);
    seed_fq(s, t, 4); d5_from_u32x8_times16(d, s.l);
    for (int i = 0; i < iters; i++) {
        d5_mul(a, a, b);
        d5_mul(c, c, d);
        d5_mul(b, b, a);
        d5_mul(d, d, c);
    }
    double x = 0;
    for (int k = 0; k < 5; k++) x += a.l[k] + b.l[k] + c.l[k] + d.l[k];
    return (uint32_t)__double2ll_rz(x * 0x1p-30);
}

// role of a warp: 0 = integer, 1 = DFMA.  Warps w and w + 4 of a block share a scheduler (warp % 4), so with
// role = (w >> 2) & 1 every scheduler hosts both kinds.
template <int KIND>
__global__ void __launch_bounds__(256) k_fqmul_pipes(uint32_t* __restrict__ sink, int iters_int, int iters_dfma) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    int role = (KIND == 1) ? 1 : (KIND == 3) ? 0 : (int)((threadIdx.x >> 7) & 1);
    sink[t] = role ? run_dfma_chain(t, iters_dfma) : run_int_chain(t, iters_int);
}

// a b 2^-256 mod p both ways for n pseudo-random pairs, plus XYZZ += affine chains of `chain` table-like points
// (multiples of a seed point by repeated addition) both ways; mismatches counted
__global__ void __launch_bounds__(128) k_dfma_selftest(uint32_t n, uint32_t chain, uint32_t* __restrict__ mismatches) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    Fq a, b, r8, r5;
    seed_fq(a, t, 11); seed_fq(b, t, 12);
    if (t % 7 == 0) { for (int k = 0; k < 8; k++) a.l[k] = 0; a.l[0] = t % 3; }  // 0, 1, 2
    fe_mul(r8, a, b);
    D5 x, y;
    d5_from_u32x8_times16(x, a.l);
    d5_from_u32x8_times16(y, b.l);
    d5_mul(x, x, y);            // (16 a)(16 b) / 2^260 = a b 2^-252
    d5_to_mont256(r5.l, x);     // * 2^-4, canonical
    bool bad = false;
    for (int k = 0; k < 8; k++) bad |= (r8.l[k] != r5.l[k]);
    // curve: G = (1, 2) in Montgomery form, points P_j = (j + 1 + t % 5) G built with the integer formulas
    Affine g;
    fe_one(g.x); fe_add(g.y, g.x, g.x);
    XYZZ run; xyzz_set_inf(run);
    for (uint32_t j = 0; j < 1 + t % 5; j++) xyzz_madd(run, g);
    XYZZ acc8; xyzz_set_inf(acc8);
    XYZZ5 acc5; xyzz5_set_inf(acc5);
    if (t % 3 == 1) {  // P + (-P): the accumulator passes through the identity
        Affine m = g;
        fe_neg(m.y, m.y);
        xyzz_madd(acc8, g); xyzz5_madd(acc5, g);
        xyzz_madd(acc8, m); xyzz5_madd(acc5, m);
    }
    for (uint32_t j = 0; j < chain; j++) {
        Affine pj;
        xyzz_to_affine(pj, run);
        if (j % 5 == 3) fe_neg(pj.y, pj.y);
        xyzz_madd(acc8, pj);
        xyzz5_madd(acc5, pj);
        if (j == 0 && t % 3 != 2) { xyzz_madd(acc8, pj); xyzz5_madd(acc5, pj); }  // P + P: the doubling path
        xyzz_madd(run, g);
    }
    XYZZ back;
    xyzz5_to_xyzz(back, acc5);
    Affine a8, a5;
    xyzz_to_affine(a8, acc8);
    xyzz_to_affine(a5, back);
    for (int k = 0; k < 8; k++) bad |= (a8.x.l[k] != a5.x.l[k]) | (a8.y.l[k] != a5.y.l[k]);
    if (bad) atomicAdd(mismatches, 1u);
}

}  // namespace kzgb

using namespace kzgb;

#define DCK(x)                                    \
    do {                                          \
        cudaError_t e_ = (x);                     \
        if (e_ != cudaSuccess) { rc = KZGB_ERR_DEVICE; goto done; } \
    } while (0)

extern "C" int kzgb_dfma_microbench(int device, int kind, int iters_int, int iters_dfma, double* ops_per_second) {
    if (!ops_per_second || kind < 0 || kind > 3 || iters_int < 0 || iters_dfma < 0) return KZGB_ERR_GENERIC;
    int rc = KZGB_OK;
    const int blocks = 148 * 8, threads = 256;
    void* sink = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float best = 1e30f;
    DCK(cudaSetDevice(device));
    DCK(cudaMalloc(&sink, (size_t)blocks * threads * 8));
    DCK(cudaStreamCreate(&st));
    DCK(cudaEventCreate(&e0));
    DCK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 4; rep++) {
        DCK(cudaEventRecord(e0, st));
        if (kind == 0) k_dfma_peak<<<blocks, threads, 0, st>>>((double*)sink, iters_dfma);
        else if (kind == 1) k_fqmul_pipes<1><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters_int, iters_dfma);
        else if (kind == 2) k_fqmul_pipes<2><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters_int, iters_dfma);
        else k_fqmul_pipes<3><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters_int, iters_dfma);
        g_launch_count++;
        DCK(cudaEventRecord(e1, st));
        DCK(cudaEventSynchronize(e1));
        float ms = 0;
        DCK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    DCK(cudaGetLastError());
    {
        double total = (double)blocks * threads;
        double per_thread = kind == 0 ? iters_dfma * 128.0
                          : kind == 1 ? iters_dfma * 4.0
                          : kind == 3 ? iters_int * 4.0
                                      : (iters_int * 4.0 + iters_dfma * 4.0) * 0.5;
        *ops_per_second = per_thread * total / (best * 1e-3);
    }
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (sink) cudaFree(sink);
    return rc;
}

// pipe-sharing probe: milliseconds of k_pipe_mix for the role pair `mix` (0: IMAD.WIDE | idle, 1: DFMA | idle,
// 2: IMAD.WIDE | DFMA, 3: IMAD.WIDE | IMAD.WIDE, 4: DFMA | DFMA, 5: DADD | idle, 6: DADD | DFMA, 7: DADD | IMAD.WIDE);
// every active thread issues iters x 128 operations of its kind.
extern "C" int kzgb_pipe_mix_probe(int device, int mix, int iters, double* ms_out) {
    if (!ms_out || mix < 0 || mix > 7 || iters <= 0) return KZGB_ERR_GENERIC;
    int rc = KZGB_OK;
    const int blocks = 148 * 8, threads = 256;
    void* sink = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float best = 1e30f;
    DCK(cudaSetDevice(device));
    DCK(cudaMalloc(&sink, (size_t)blocks * threads * 4));
    DCK(cudaStreamCreate(&st));
    DCK(cudaEventCreate(&e0));
    DCK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 4; rep++) {
        DCK(cudaEventRecord(e0, st));
        switch (mix) {
            case 0: k_pipe_mix<0, 3><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters); break;
            case 1: k_pipe_mix<1, 3><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters); break;
            case 2: k_pipe_mix<0, 1><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters); break;
            case 3: k_pipe_mix<0, 0><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters); break;
            case 4: k_pipe_mix<1, 1><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters); break;
            case 5: k_pipe_mix<2, 3><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters); break;
            case 6: k_pipe_mix<2, 1><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters); break;
            default: k_pipe_mix<2, 0><<<blocks, threads, 0, st>>>((uint32_t*)sink, iters); break;
        }
        g_launch_count++;
        DCK(cudaEventRecord(e1, st));
        DCK(cudaEventSynchronize(e1));
        float ms = 0;
        DCK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    DCK(cudaGetLastError());
    *ms_out = best;
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (sink) cudaFree(sink);
    return rc;
}

extern "C" int kzgb_dfma_selftest(int device, uint32_t n, uint32_t chain, uint32_t* mismatches) {
    if (!mismatches || n == 0) return KZGB_ERR_GENERIC;
    int rc = KZGB_OK;
    uint32_t* d = nullptr;
    DCK(cudaSetDevice(device));
    DCK(cudaMalloc(&d, 4));
    DCK(cudaMemset(d, 0, 4));
    k_dfma_selftest<<<(n + 127) / 128, 128>>>(n, chain, d);
    g_launch_count++;
    DCK(cudaGetLastError());
    DCK(cudaMemcpy(mismatches, d, 4, cudaMemcpyDeviceToHost));
done:
    if (d) cudaFree(d);
    return rc;
}
