// Fr radix-2 (I)NTT, natural order in and out, omega = PRIMITIVE_ROOTS_OF_UNITY[log2 n]
// (reference primitives/src/consts.rs:22-52).  Replaces ark-poly's
// GeneralEvaluationDomain::<Fr>::{fft,ifft} at primitives/src/polynomial.rs:131-135,242-246 and
// is what lets commit_eval_form (prover/src/kzg.rs:84-104) skip the G1-point IFFT:
// MSM(IFFT_G1(SRS), f) == MSM(SRS, IFFT_Fr(f)).
//
// Decimation-in-frequency passes of up to 10 radix-2 stages each, staged through shared
// memory in limb-major (SoA) layout; the last pass stores to the bit-reversed index so the
// output is in natural order, and folds in 1/n for the inverse.  An Fr element is 32 B = one
// DRAM sector, so the strided tiles and the bit-reversed scatter are sector-efficient.
#include "kzgb_internal.hpp"

namespace kzgb {

static constexpr int MAX_TILE_LOG = 10;
static constexpr int MAX_TILE = 1 << MAX_TILE_LOG;

__global__ void __launch_bounds__(256) k_twiddles(Fr* __restrict__ tw, uint32_t count, Fr omega) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    uint32_t e[8] = {j, 0, 0, 0, 0, 0, 0, 0};
    Fr r;
    fe_pow(r, omega, e);
    fe_store(&tw[j], r);
}

// omega_n^t for the stage whose butterflies span 2^(s+1) elements; idx = t << (logN-1-s)
template <bool INV>
__device__ __forceinline__ Fr twiddle(const Fr* __restrict__ tw, uint32_t idx, uint32_t halfN) {
    if (!INV) return fe_load_ro(&tw[idx]);
    Fr w;
    if (idx == 0) { fe_one(w); return w; }
    w = fe_load_ro(&tw[halfN - idx]);
    fe_neg(w, w);
    return w;
}

template <bool INV>
__global__ void __launch_bounds__(512) k_ntt_pass(const Fr* __restrict__ src, Fr* __restrict__ dst, int logn, int s_hi,
                                                   int s_lo, int g, const Fr* __restrict__ tw, int logN, bool last,
                                                   Fr ninv) {
    __shared__ uint32_t sm[8][MAX_TILE];
    const int b = s_hi - s_lo + 1;
    const uint32_t TE = 1u << (b + g);
    const uint32_t tid = threadIdx.x;  // TE/2 threads
    const uint32_t n = 1u << logn;
    const Fr* in = src + (size_t)blockIdx.y * n;
    Fr* out = dst + (size_t)blockIdx.y * n;
    const uint32_t tile = blockIdx.x;
    const uint32_t low_hi_bits = s_lo - g;
    const uint32_t low_hi = tile & ((1u << low_hi_bits) - 1u);
    const uint32_t top = tile >> low_hi_bits;
    const uint32_t gmask = (1u << g) - 1u;
    const uint32_t halfN = 1u << (logN - 1);

    auto gidx = [&](uint32_t e) -> uint32_t {
        uint32_t mid = e >> g, lp = e & gmask;
        return (top << (s_hi + 1)) | (mid << s_lo) | (low_hi << g) | lp;
    };
    for (uint32_t e = tid; e < TE; e += TE / 2) {
        Fr v = fe_load(&in[gidx(e)]);
#pragma unroll
        for (int k = 0; k < 8; k++) sm[k][e] = v.l[k];
    }
    __syncthreads();
    for (int s = s_hi; s >= s_lo; s--) {
        const int lb = (s - s_lo) + g;
        const uint32_t h = 1u << lb;
        const uint32_t e0 = ((tid >> lb) << (lb + 1)) | (tid & (h - 1u));
        const uint32_t e1 = e0 + h;
        Fr u, v;
#pragma unroll
        for (int k = 0; k < 8; k++) { u.l[k] = sm[k][e0]; v.l[k] = sm[k][e1]; }
        // t = (global index of e0) mod 2^s
        const uint32_t mid0 = e0 >> g;
        const uint32_t t = ((mid0 & ((1u << (s - s_lo)) - 1u)) << s_lo) | (low_hi << g) | (e0 & gmask);
        Fr w = twiddle<INV>(tw, t << (logN - 1 - s), halfN);
        Fr sum, dif;
        fe_add(sum, u, v);
        fe_sub(dif, u, v);
        fe_mul(dif, dif, w);
#pragma unroll
        for (int k = 0; k < 8; k++) { sm[k][e0] = sum.l[k]; sm[k][e1] = dif.l[k]; }
        __syncthreads();
    }
    for (uint32_t e = tid; e < TE; e += TE / 2) {
        Fr v;
#pragma unroll
        for (int k = 0; k < 8; k++) v.l[k] = sm[k][e];
        uint32_t i = gidx(e);
        if (last) {
            if (INV) fe_mul(v, v, ninv);
            i = __brev(i) >> (32 - logn);
        }
        fe_store(&out[i], v);
    }
}

void ntt_twiddles_launch(Fr* tw, int logN, const Fr* omega_mont_host, cudaStream_t st) {
    uint32_t count = logN >= 1 ? (1u << (logN - 1)) : 1u;
    k_twiddles<<<(count + 255) / 256, 256, 0, st>>>(tw, count, *omega_mont_host);
    g_launch_count++;
}

void ntt_launch(Fr* data, int logn, uint32_t batch, bool inverse, const Fr* tw, int logN, const Fr* ninv_mont_host,
                Fr* scratch, cudaStream_t st) {
    if (logn == 0 || batch == 0) return;  // size-1 transform is the identity (1/n = 1)
    const uint32_t n = 1u << logn;
    // plan: P passes; the last (low stages, contiguous tiles, g = 0) gets the larger share
    int P = (logn + MAX_TILE_LOG - 1) / MAX_TILE_LOG;
    int bits[8];
    {
        int rem = logn;
        for (int p = P - 1; p >= 0; p--) {  // fill from the last pass backwards
            int left = p + 1;
            int take = (rem + left - 1) / left;
            bits[p] = take;
            rem -= take;
        }
    }
    int s_hi = logn - 1;
    const Fr* src = data;
    for (int p = 0; p < P; p++) {
        const bool last = (p == P - 1);
        int b = bits[p];
        int s_lo = s_hi - b + 1;
        int g = (!last && b < MAX_TILE_LOG && s_lo >= 1) ? 1 : 0;
        Fr* dst = last ? (P == 1 ? scratch : data) : scratch;
        // P >= 2: pass 0 data->scratch, middle scratch->scratch (tile-local in place), last scratch->data.
        // P == 1: data->scratch (bit-reversed scatter cannot be in place), then copy back.
        uint32_t TE = 1u << (b + g);
        dim3 grid(n / TE, batch);
        if (inverse)
            k_ntt_pass<true><<<grid, TE / 2, 0, st>>>(src, dst, logn, s_hi, s_lo, g, tw, logN, last, *ninv_mont_host);
        else
            k_ntt_pass<false><<<grid, TE / 2, 0, st>>>(src, dst, logn, s_hi, s_lo, g, tw, logN, last, *ninv_mont_host);
        g_launch_count++;
        src = dst;
        s_hi = s_lo - 1;
    }
    if (P == 1)
        cudaMemcpyAsync(data, scratch, (size_t)batch * n * sizeof(Fr), cudaMemcpyDeviceToDevice, st);
}

}  // namespace kzgb
