// Fr radix-2 (I)NTT, natural order in and out, omega = PRIMITIVE_ROOTS_OF_UNITY[log2 n]
// (reference primitives/src/consts.rs:22-52).  Replaces ark-poly's
// GeneralEvaluationDomain::<Fr>::{fft,ifft} at primitives/src/polynomial.rs:131-135,242-246 and
// is what lets commit_eval_form (prover/src/kzg.rs:84-104) skip the G1-point IFFT:
// MSM(IFFT_G1(SRS), f) == MSM(SRS, IFFT_Fr(f)).
//
// Decimation-in-frequency passes; the last pass stores to the bit-reversed index so the output is in natural order and folds
// in 1/n for the inverse; the stage whose twiddles are all 1 multiplies nothing.  Two implementations (ntt_launch below):
// k_ntt_pass -- up to 10 stages per pass through a limb-major shared-memory tile, one butterfly per thread and stage -- and
// k_ntt_warp:
// passes of up to 8 radix-2 stages each.  One WARP owns 256 elements of a pass for its whole
// length: every lane keeps 8 elements in registers (64 limb registers), three of the eight index bits of the group are
// "slot" bits (which register), five are "lane" bits.  A stage whose bit is a slot bit is four butterflies inside the
// thread; a stage whose bit is a lane bit first trades half of each lane's registers with lane ^ 2^k by warp shuffles
// (the bit becomes a slot bit, an already finished bit takes its place among the lane bits), so that every lane again
// does four full butterflies -- no lane ever idles through a multiplication, no shared memory, no block barrier in the
// butterfly network.  Twiddles of the last pass (omega_256^t, the same for every warp) sit in shared memory; the other
// passes index the resident omega_N table.  The last pass reads its contiguous 8 KiB tile with ONE bulk asynchronous copy
// per warp (cp.async.bulk + mbarrier: the copy engine transposes memory order into the (slot, lane) register layout through
// shared memory), stores to the bit-reversed index so the output is in natural order, and folds in 1/n for the inverse;
// its final stage has w = 1 everywhere and multiplies nothing.  An Fr element is 32 B = one DRAM sector, so the strided
// accesses of the upper passes and the bit-reversed scatter are sector-efficient.
// Work: (n/2) log2 n butterflies = (n/2)(log2 n - 1) multiplications (+ n for the inverse's 1/n), 64 B of traffic per
// element and pass.
#include "kzgb_internal.hpp"

namespace kzgb {

static constexpr int NW_THREADS = 128;
static constexpr int MAX_TILE_LOG = 10;
static constexpr int MAX_TILE = 1 << MAX_TILE_LOG;
  // 4 warps, each with its own 256-element group

__global__ void __launch_bounds__(256) k_twiddles(Fr* __restrict__ tw, uint32_t count, Fr omega) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    uint32_t e[8] = {j, 0, 0, 0, 0, 0, 0, 0};
    Fr r;
    fe_pow(r, omega, e);
    fe_store(&tw[j], r);
}

// omega_N^idx (forward) or omega_N^-idx (inverse) from the table tw[j] = omega_N^j, j < N/2
template <bool INV>
__device__ __forceinline__ Fr twiddle(const Fr* __restrict__ tw, uint32_t idx, uint32_t halfN) {
    if (!INV) return fe_load_ro(&tw[idx]);
    Fr w;
    if (idx == 0) { fe_one(w); return w; }
    w = fe_load_ro(&tw[halfN - idx]);
    fe_neg(w, w);
    return w;
}

__device__ __forceinline__ Fr fr_shfl_xor(const Fr& v, int m) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], m);
    return r;
}

// Flat index (over batch x n elements) of element p (8 bits) of the 256-element group wg of a pass that transforms
// index bits [s_lo, s_lo + W): the low W bits of p are the transform bits, the other 8 - W "batch" bits select
// neighbouring sub-transforms (adjacent in memory whenever s_lo > 0).
__device__ __forceinline__ uint64_t ntt_flat_index(uint32_t p, uint64_t wg, int W, int s_lo) {
    const uint32_t pt = p & ((1u << W) - 1u);
    const uint64_t other = (wg << (8 - W)) | (uint64_t)(p >> W);
    const uint64_t lo = other & ((1ull << s_lo) - 1ull);
    const uint64_t hi = other >> s_lo;
    return (hi << (s_lo + W)) | ((uint64_t)pt << s_lo) | lo;
}

// The exchange schedule, fixed at compile time.  Initially slot bits (s0, s1, s2) stand for index bits (5, 6, 7) of the
// group and lane bits (l0..l4) for (0..4).  Stage q works on slot bit j; for q <= 4 it first swaps lane bit k with slot
// bit j (whose old bit is already done).  sb / lb: the arrangement AFTER that swap.
struct NttSched { int j, k; int sb[3]; int lb[5]; };
__device__ constexpr NttSched NTT_SCHED[8] = {
    /* q = 0 */ {1, 0, {2, 0, 1}, {3, 4, 5, 6, 7}},
    /* q = 1 */ {2, 1, {2, 3, 1}, {0, 4, 5, 6, 7}},
    /* q = 2 */ {0, 2, {2, 3, 4}, {0, 1, 5, 6, 7}},
    /* q = 3 */ {1, 3, {5, 3, 4}, {0, 1, 2, 6, 7}},
    /* q = 4 */ {2, 4, {5, 6, 4}, {0, 1, 2, 3, 7}},
    /* q = 5 */ {0, -1, {5, 6, 7}, {0, 1, 2, 3, 4}},
    /* q = 6 */ {1, -1, {5, 6, 7}, {0, 1, 2, 3, 4}},
    /* q = 7 */ {2, -1, {5, 6, 7}, {0, 1, 2, 3, 4}},
};
__device__ __forceinline__ constexpr uint32_t ntt_slot_part(const NttSched& S, int r) {
    return ((uint32_t)(r & 1) << S.sb[0]) | ((uint32_t)((r >> 1) & 1) << S.sb[1]) | ((uint32_t)((r >> 2) & 1) << S.sb[2]);
}
__device__ __forceinline__ uint32_t ntt_lane_part(const NttSched& S, uint32_t L) {
    return ((L & 1u) << S.lb[0]) | (((L >> 1) & 1u) << S.lb[1]) | (((L >> 2) & 1u) << S.lb[2]) | (((L >> 3) & 1u) << S.lb[3]) |
           (((L >> 4) & 1u) << S.lb[4]);
}

template <int Q, bool INV>
__device__ __forceinline__ void ntt_stage(Fr (&x)[8], uint32_t L, uint64_t wg, int W, int s_lo, bool last, const Fr* __restrict__ tw,
                                          const Fr* tws, int logN, uint32_t halfN) {
    constexpr NttSched S = NTT_SCHED[Q];
    constexpr int J = S.j, K = S.k;
    if (K >= 0) {  // Q is a lane bit: trade halves with lane ^ 2^K, Q becomes slot bit J
        const bool hi = (L >> (K < 0 ? 0 : K)) & 1u;
#pragma unroll
        for (int h = 0; h < 4; h++) {
            constexpr int dummy = 0; (void)dummy;
            const int a = ((h >> J) << (J + 1)) | (h & ((1 << J) - 1)), b = a | (1 << J);
            Fr t;
#pragma unroll
            for (int i = 0; i < 8; i++) t.l[i] = hi ? x[a].l[i] : x[b].l[i];
            t = fr_shfl_xor(t, 1 << (K < 0 ? 0 : K));
#pragma unroll
            for (int i = 0; i < 8; i++) { x[a].l[i] = hi ? t.l[i] : x[a].l[i]; x[b].l[i] = hi ? x[b].l[i] : t.l[i]; }
        }
    }
    if (Q < W) {  // (uniform) stage s = s_lo + Q: butterflies between slots a and a | 2^J
        const int s = s_lo + Q;
        const uint32_t lane_p = ntt_lane_part(S, L);
        Fr w[4];
        if (s > 0) {  // all four twiddles first: their loads are in flight together (stage 0: w = 1, nothing to fetch)
#pragma unroll
            for (int h = 0; h < 4; h++) {
                const int a = ((h >> J) << (J + 1)) | (h & ((1 << J) - 1));
                const uint32_t pu = lane_p | ntt_slot_part(S, a);
                const uint32_t tq = pu & ((1u << Q) - 1u);  // transform bits below Q
                if (last) w[h] = tws[tq << (W - 1 - Q)];
                else {
                    const uint64_t other = (wg << (8 - W)) | (uint64_t)(pu >> W);
                    const uint32_t t = (tq << s_lo) | (uint32_t)(other & ((1ull << s_lo) - 1ull));
                    w[h] = twiddle<INV>(tw, t << (logN - 1 - s), halfN);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const int a = ((h >> J) << (J + 1)) | (h & ((1 << J) - 1)), b = a | (1 << J);
            Fr sum, dif;
            fe_add(sum, x[a], x[b]);
            fe_sub(dif, x[a], x[b]);
            if (s > 0) fe_mul(dif, dif, w[h]);
            x[a] = sum;
            x[b] = dif;
        }
    }
}

template <bool INV>
__global__ void __launch_bounds__(NW_THREADS, 4) k_ntt_warp(const Fr* __restrict__ src, Fr* __restrict__ dst, uint64_t E, int logn, int W,
                                                            int s_lo, bool last, const Fr* __restrict__ tw, int logN, Fr ninv) {
    __shared__ Fr tws[128];                                        // last pass: omega_{2^W}^t, t < 2^(W-1)
    __shared__ alignas(128) Fr tile[NW_THREADS / 32][256];         // last pass: the warp's contiguous tile (bulk copy target)
    __shared__ alignas(8) uint64_t mbar[NW_THREADS / 32];
    const uint32_t L = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint64_t wg = (uint64_t)blockIdx.x * (NW_THREADS / 32) + wid;
    const uint32_t halfN = 1u << (logN - 1);
    const bool bulk = last && W == 8 && ((wg + 1) << 8) <= E;      // warp-uniform: the whole tile exists and is contiguous
    if (bulk) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[wid]);
        const uint32_t dsts = (uint32_t)__cvta_generic_to_shared(&tile[wid][0]);
        if (L == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(256u * 32u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dsts), "l"(src + (wg << 8)), "r"(256u * 32u), "r"(bar) : "memory");
        }
    }
    if (last) {
        for (uint32_t t = threadIdx.x; t < (1u << (W - 1)); t += NW_THREADS) tws[t] = twiddle<INV>(tw, t << (logN - W), halfN);
        __syncthreads();
    }
    Fr x[8];
    if (bulk) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[wid]);
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar) : "memory");
        }
#pragma unroll
        for (int r = 0; r < 8; r++) x[r] = tile[wid][(r << 5) | L];
    } else {
#pragma unroll
        for (int r = 0; r < 8; r++) {  // slot r of lane L holds p = (r << 5) | L
            const uint64_t f = ntt_flat_index(((uint32_t)r << 5) | L, wg, W, s_lo);
            if (f < E) x[r] = fe_load(&src[f]); else fe_zero(x[r]);
        }
    }
    ntt_stage<7, INV>(x, L, wg, W, s_lo, last, tw, tws, logN, halfN);
    ntt_stage<6, INV>(x, L, wg, W, s_lo, last, tw, tws, logN, halfN);
    ntt_stage<5, INV>(x, L, wg, W, s_lo, last, tw, tws, logN, halfN);
    ntt_stage<4, INV>(x, L, wg, W, s_lo, last, tw, tws, logN, halfN);
    ntt_stage<3, INV>(x, L, wg, W, s_lo, last, tw, tws, logN, halfN);
    ntt_stage<2, INV>(x, L, wg, W, s_lo, last, tw, tws, logN, halfN);
    ntt_stage<1, INV>(x, L, wg, W, s_lo, last, tw, tws, logN, halfN);
    ntt_stage<0, INV>(x, L, wg, W, s_lo, last, tw, tws, logN, halfN);
    {
        constexpr NttSched S = NTT_SCHED[0];  // the final arrangement
        const uint32_t lane_p = ntt_lane_part(S, L);
        const uint64_t nmask = (1ull << logn) - 1ull;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t p = lane_p | ntt_slot_part(S, r);
            uint64_t f = ntt_flat_index(p, wg, W, s_lo);
            if (f < E) {
                Fr v = x[r];
                if (last) {
                    if (INV) fe_mul(v, v, ninv);
                    const uint32_t i = (uint32_t)(f & nmask);
                    f = (f & ~nmask) | (uint64_t)(__brev(i) >> (32 - logn));
                }
                fe_store(&dst[f], v);
            }
        }
    }
}

void ntt_twiddles_launch(Fr* tw, int logN, const Fr* omega_mont_host, cudaStream_t st) {
    uint32_t count = logN >= 1 ? (1u << (logN - 1)) : 1u;
    k_twiddles<<<(count + 255) / 256, 256, 0, st>>>(tw, count, *omega_mont_host);
    g_launch_count++;
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
        Fr sum, dif;
        fe_add(sum, u, v);
        fe_sub(dif, u, v);
        if (s > 0) {  // stage 0: w = 1 for every butterfly
            Fr w = twiddle<INV>(tw, t << (logN - 1 - s), halfN);
            fe_mul(dif, dif, w);
        }
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


static void ntt_launch_warp(Fr* data, int logn, uint32_t batch, bool inverse, const Fr* tw, int logN, const Fr* ninv_mont_host,
                Fr* scratch, cudaStream_t st) {
    if (logn == 0 || batch == 0) return;  // size-1 transform is the identity (1/n = 1)
    const uint64_t E = (uint64_t)batch << logn;
    // P passes of at most 8 stages, widths as even as possible, the last (low stages) taking the larger share
    const int P = (logn + 7) / 8;
    int bits[8];
    {
        int rem = logn;
        for (int p = P - 1; p >= 0; p--) {
            int left = p + 1;
            int take = (rem + left - 1) / left;
            bits[p] = take;
            rem -= take;
        }
    }
    const uint64_t groups = (E + 255) / 256;
    const unsigned grid = (unsigned)((groups + NW_THREADS / 32 - 1) / (NW_THREADS / 32));
    int s_hi = logn - 1;
    // pass 0: data -> scratch, middle passes in place in scratch, last pass scratch -> data (bit-reversed scatter).
    // P == 1: data -> scratch, then copy back (the scatter cannot be in place).
    const Fr* src = data;
    for (int p = 0; p < P; p++) {
        const bool last = (p == P - 1);
        const int W = bits[p], s_lo = s_hi - W + 1;
        Fr* dst = last ? (P == 1 ? scratch : data) : scratch;
        if (inverse) k_ntt_warp<true><<<grid, NW_THREADS, 0, st>>>(src, dst, E, logn, W, s_lo, last, tw, logN, *ninv_mont_host);
        else k_ntt_warp<false><<<grid, NW_THREADS, 0, st>>>(src, dst, E, logn, W, s_lo, last, tw, logN, *ninv_mont_host);
        g_launch_count++;
        src = dst;
        s_hi = s_lo - 1;
    }
    if (P == 1) cudaMemcpyAsync(data, scratch, E * sizeof(Fr), cudaMemcpyDeviceToDevice, st);
}


static void ntt_launch_tiles(Fr* data, int logn, uint32_t batch, bool inverse, const Fr* tw, int logN, const Fr* ninv_mont_host,
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


// Two implementations of the same passes, bit-identical results (tests run both):
//   0  shared-memory tiles (k_ntt_pass): up to 10 stages per pass, one butterfly per thread and stage, limb-major tile in
//      shared memory, 64 warps per SM
//   1  warp-resident (k_ntt_warp): the register / shuffle / bulk-copy design described at the top of this file, 16 warps per SM
// Measured on B200 (scripts/ntt_bench.py, profiles/r02_ntt.txt): 1024 x 2^16: 11.1 ms (0) vs 12.5 ms (1); one 2^19: 0.147 vs
// 0.151 ms.  Both sit on the instruction-issue model of DESIGN.md 3 (cycles per warp = 4 x IMAD.WIDE + 1 x everything else:
// a butterfly is 136 wide multiplies and ~300 other instructions), not on HBM and not on the multiplier alone; the register-
// resident form pays its 64 limb registers per lane with a quarter of the occupancy and gains nothing back from skipping
// shared memory.  Many SMALL transforms are the exception (4096 x 2^12: 2.40 ms (1) vs 3.30 ms (0)): a single pass, twiddles in
// shared memory, one bulk copy per warp, no block barrier.  Option "ntt_kernel": -1 (default) picks per shape -- (1) for
// batches of >= 2^21 elements in transforms of <= 2^13, (0) otherwise; 0 / 1 force one.
static std::atomic<int> g_ntt_kernel{-1};
void ntt_set_kernel(int which) { g_ntt_kernel.store(which < 0 ? -1 : (which ? 1 : 0)); }

void ntt_launch(Fr* data, int logn, uint32_t batch, bool inverse, const Fr* tw, int logN, const Fr* ninv_mont_host,
                Fr* scratch, cudaStream_t st) {
    const int which = g_ntt_kernel.load();
    const bool warp = which >= 0 ? which == 1 : (logn <= 13 && ((uint64_t)batch << logn) >= (1ull << 21));
    if (warp) ntt_launch_warp(data, logn, batch, inverse, tw, logN, ninv_mont_host, scratch, st);
    else ntt_launch_tiles(data, logn, batch, inverse, tw, logN, ninv_mont_host, scratch, st);
}

}  // namespace kzgb
