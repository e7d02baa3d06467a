// BN254 G1 multi-scalar multiplication: signed-digit Pippenger, hand-written for sm_100a.
// Replaces the reference's `G1Projective::msm` call sites (prover/src/kzg.rs:100,121,
// primitives/src/helpers.rs:332).
//
// Pipeline (all on one stream, no host sync):
//   1. k_digits_hist   scalar -> canonical -> W signed c-bit digits; histogram of bucket keys
//   2. k_scan          exclusive scan of the histogram (bucket offsets)
//   3. k_scatter       counting-sort scatter: point refs grouped by bucket (order inside a
//                      bucket is irrelevant: group addition is commutative and exact)
//   4. k_accumulate    FIXED-SIZE chunks of the sorted list per thread (perfect balance for any
//                      digit distribution); XYZZ += affine gathers from the HBM-resident table
//   5. k_bucket_fix    stitches buckets that span several chunks, zeroes empty buckets
//   6. k_reduce_slices per-slice running sums  sum (k+1) * B_k  + small scalar fix-up
//   7. k_tree_reduce   tree sum of the slice results -> one XYZZ point per bucket set
// The caller copies `sets` XYZZ points (128 B each) back and finishes on the host.
// One launch set can also carry `batch` independent fixed-base MSMs over the same table (MsmPlan.batch_n):
// scalar i belongs to MSM i / batch_n, which owns bucket set i / batch_n -- how runs of small blobs are committed.
#include <cstdlib>
#include "kzgb_internal.hpp"
#include "ec_dfma.cuh"
#include "field_gen_x.cuh"

namespace kzgb {

static constexpr int SM_COUNT = 148;

__device__ __forceinline__ uint32_t scalar_bits(const uint32_t* l, int pos, int c) {
    int w = pos >> 5, b = pos & 31;
    if (w >= 8) return 0;
    uint64_t v = l[w];
    if (w + 1 < 8) v |= (uint64_t)l[w + 1] << 32;
    return (uint32_t)(v >> b) & ((1u << c) - 1u);
}

// One signed digit step.  Returns magnitude in [0, 2^(c-1)], sets neg, updates carry.
__device__ __forceinline__ uint32_t signed_digit(const uint32_t* l, int w, int c, uint32_t& carry, bool& neg) {
    uint32_t d = scalar_bits(l, w * c, c) + carry;
    uint32_t half = 1u << (c - 1);
    if (d > half) { neg = true; carry = 1; return (1u << c) - d; }
    neg = false; carry = 0;
    return d;
}

__global__ void __launch_bounds__(256) k_digits_hist(const Fr* __restrict__ scalars, bool canonical, MsmPlan p,
                                                      Fr* __restrict__ canon, uint32_t* __restrict__ hist) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    Fr s = fe_load_ro(&scalars[i]);
    if (!canonical) fe_from_mont(s, s);
    fe_store(&canon[i], s);
    uint32_t half = 1u << (p.c - 1);
    uint32_t carry = 0;
    for (int w = 0; w < p.W; w++) {
        bool neg;
        uint32_t mag = signed_digit(s.l, w, p.c, carry, neg);
        if (mag) {
            uint32_t key = (p.batch_n ? (i / p.batch_n) * half : (p.sets == 1 ? 0u : (uint32_t)w * half)) + mag - 1u;
            atomicAdd(&hist[key], 1u);
        }
    }
}

// Exclusive scan of the bucket histogram, 32 counters per thread held in registers (8 x 128-bit
// loads in flight), warp-shuffle scan of the thread totals, one smem hop across warps.
// block b scans items [b*SCAN_TILE, (b+1)*SCAN_TILE); block_tot[b] = its total.
static constexpr uint32_t SCAN_PER_THREAD = 32;
static constexpr uint32_t SCAN_THREADS = 1024;
static constexpr uint32_t SCAN_TILE = SCAN_PER_THREAD * SCAN_THREADS;

// pad_mask = 2^L - 1 rounds every bucket's count up to a multiple of 2^L (batch-affine levels pair
// entries (2p, 2p+1) globally, so every bucket must start on a multiple of 2^L; the padding slots hold
// REF_IDENT).
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor,
                                                             uint32_t nb, uint32_t* __restrict__ block_tot, bool single,
                                                             uint32_t pad_mask) {
    __shared__ uint32_t warp_tot[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t base = blockIdx.x * SCAN_TILE + tid * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD];
    if (base + SCAN_PER_THREAD <= nb) {
        const uint4* q = reinterpret_cast<const uint4*>(hist + base);
#pragma unroll
        for (int k = 0; k < 8; k++) { uint4 t = q[k]; v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w; }
    } else {
#pragma unroll
        for (int k = 0; k < (int)SCAN_PER_THREAD; k++) v[k] = (base + k < nb) ? hist[base + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < (int)SCAN_PER_THREAD; k++) v[k] = (v[k] + pad_mask) & ~pad_mask;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < (int)SCAN_PER_THREAD; k++) { uint32_t t = v[k]; v[k] = sum; sum += t; }
    uint32_t inc = sum;
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
        if ((int)lane >= off) inc += o;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = warp_tot[lane], wi = w;
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, wi, off);
            if ((int)lane >= off) wi += o;
        }
        warp_tot[lane] = wi - w;  // exclusive
        if (lane == 31) {
            block_tot[blockIdx.x] = wi;
            if (single) hist[nb] = wi;
        }
    }
    __syncthreads();
    const uint32_t off0 = warp_tot[wid] + (inc - sum);
    if (base + SCAN_PER_THREAD <= nb) {
        uint4* q = reinterpret_cast<uint4*>(hist + base);
        uint4* qc = reinterpret_cast<uint4*>(cursor + base);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint4 t = make_uint4(v[4 * k] + off0, v[4 * k + 1] + off0, v[4 * k + 2] + off0, v[4 * k + 3] + off0);
            q[k] = t;
            if (single) qc[k] = t;
        }
    } else {
#pragma unroll
        for (int k = 0; k < (int)SCAN_PER_THREAD; k++)
            if (base + k < nb) { hist[base + k] = v[k] + off0; if (single) cursor[base + k] = v[k] + off0; }
    }
}

// multi-tile case: exclusive scan of the tile totals (<= 1024 tiles), then add them back
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_totals(uint32_t* __restrict__ block_tot, uint32_t ntiles,
                                                                   uint32_t* __restrict__ hist, uint32_t nb) {
    __shared__ uint32_t sm[SCAN_THREADS];
    const uint32_t tid = threadIdx.x;
    uint32_t v = tid < ntiles ? block_tot[tid] : 0u;
    sm[tid] = v;
    __syncthreads();
    for (uint32_t off = 1; off < SCAN_THREADS; off <<= 1) {
        uint32_t o = tid >= off ? sm[tid - off] : 0u;
        __syncthreads();
        sm[tid] += o;
        __syncthreads();
    }
    if (tid < ntiles) block_tot[tid] = sm[tid] - v;
    if (tid == SCAN_THREADS - 1) hist[nb] = sm[tid];
}
__global__ void __launch_bounds__(256) k_scan_add(uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor, uint32_t nb,
                                                   const uint32_t* __restrict__ block_tot) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    uint32_t v = hist[i] + block_tot[i / SCAN_TILE];
    hist[i] = v;
    cursor[i] = v;
}

__global__ void __launch_bounds__(256) k_scatter(const Fr* __restrict__ canon, MsmPlan p, uint32_t* __restrict__ cursor,
                                                  uint32_t* __restrict__ sorted) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    Fr s = fe_load_ro(&canon[i]);
    uint32_t half = 1u << (p.c - 1);
    uint32_t carry = 0;
    for (int w = 0; w < p.W; w++) {
        bool neg;
        uint32_t mag = signed_digit(s.l, w, p.c, carry, neg);
        if (mag) {
            uint32_t key = (p.batch_n ? (i / p.batch_n) * half : (p.sets == 1 ? 0u : (uint32_t)w * half)) + mag - 1u;
            uint32_t pos = atomicAdd(&cursor[key], 1u);
            uint32_t ref = p.batch_n ? ((uint32_t)w * p.table_stride + p.base_offset + i % p.batch_n)
                                     : (p.sets == 1) ? ((uint32_t)w * p.table_stride + p.base_offset + i) : i;
            sorted[pos] = ref | (neg ? 0x80000000u : 0u);
        }
    }
}

static constexpr uint32_t REF_IDENT = 0xffffffffu;  // padding slot of the sorted list: the identity

__device__ __forceinline__ Affine load_point(const Affine* __restrict__ table, uint32_t ref) {
    if (ref == REF_IDENT) { Affine z; aff_set_inf(z); return z; }
    Affine q = aff_gather_ro(&table[ref & 0x7fffffffu]);
    if (ref & 0x80000000u) fe_neg(q.y, q.y);  // identity (0,0) stays (0,0)
    return q;
}

// Out-of-line multiplication for the accumulation loop (arguments and result in registers): the loop
// body with 10 inlined multiplications is ~35 KB of SASS, more than the 32 KB L1.5 instruction cache.
static __device__ __noinline__ Fq fq_mul_call(Fq a, Fq b) {
    Fq r;
    fe_mul(r, a, b);
    return r;
}
// xyzz_madd (ec.cuh) with called multiplications; exceptional cases delegate to the complete version
__device__ __forceinline__ void xyzz_madd_call(XYZZ& acc, const Affine& q) {
    if (aff_is_inf(q)) return;
    if (xyzz_is_inf(acc)) { acc.x = q.x; acc.y = q.y; fe_one(acc.zz); fe_one(acc.zzz); return; }
    Fq U2 = fq_mul_call(q.x, acc.zz);
    Fq S2 = fq_mul_call(q.y, acc.zzz);
    Fq Pp, Rr;
    fe_sub(Pp, U2, acc.x);
    fe_sub(Rr, S2, acc.y);
    if (fe_is_zero(Pp)) {
        if (fe_is_zero(Rr)) xyzz_dbl_affine(acc, q);
        else xyzz_set_inf(acc);
        return;
    }
    Fq PP = fq_mul_call(Pp, Pp);
    Fq PPP = fq_mul_call(Pp, PP);
    Fq Q = fq_mul_call(acc.x, PP);
    Fq t = fq_mul_call(Rr, Rr);
    fe_sub(t, t, PPP); fe_sub(t, t, Q); fe_sub(t, t, Q);  // X3
    fe_sub(Q, Q, t);
    Q = fq_mul_call(Rr, Q);
    S2 = fq_mul_call(acc.y, PPP);
    fe_sub(acc.y, Q, S2);
    acc.x = t;
    acc.zz = fq_mul_call(acc.zz, PP);
    acc.zzz = fq_mul_call(acc.zzz, PPP);
}

// Source of the entries: SRC_REFS = sorted refs into the table (gather), SRC_POINTS = the affine
// points left by the batch-affine levels (entry `pos` is points[pos]; bucket offsets are the level-0
// offsets >> shift).
template <bool CALL, bool PREFETCH, bool DIRECT, bool LAZY = false, bool RELAXED = false, bool L2PF = false>
__device__ __forceinline__ void accumulate_body(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                                                const Affine* __restrict__ table, uint32_t nb, uint32_t acc_threads,
                                                uint32_t chunk, int shift,
                                                XYZZ* __restrict__ buckets, XYZZ* __restrict__ partial) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= acc_threads) return;
    const uint32_t M = offsets[nb] >> shift;
    uint64_t start64 = (uint64_t)t * chunk;
    if (start64 >= M) return;  // k_bucket_fix only reads slots of chunks that hold entries
    uint32_t start = (uint32_t)start64;
    uint32_t end = (uint32_t)min((uint64_t)M, start64 + chunk);
    // largest b with offsets[b] <= start
    uint32_t lo = 0, hi = nb;  // invariant: offsets[lo] <= start < offsets[hi] (offsets[nb] = M > start)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if ((offsets[mid] >> shift) <= start) lo = mid; else hi = mid;
    }
    uint32_t b = lo;
    uint32_t run_begin = offsets[b] >> shift, next = offsets[b + 1] >> shift;
    XYZZ acc; xyzz_set_inf(acc);
    Affine q;
    auto fetch = [&](uint32_t pos) -> Affine {
        if (DIRECT) return aff_load_ro(&table[pos]);
        return load_point(table, sorted[pos]);
    };
    if (PREFETCH) q = fetch(start);
    for (uint32_t pos = start; pos < end; pos++) {
        Affine qn;
        if (PREFETCH) { if (pos + 1 < end) qn = fetch(pos + 1); }  // next point in flight during this addition
        else q = fetch(pos);
        if (L2PF && !DIRECT && pos + 1 < end) {  // pull the NEXT point into L2 while this addition runs (no registers held)
            uint32_t rn = sorted[pos + 1];
            if (rn != REF_IDENT) asm volatile("prefetch.global.L2 [%0];" ::"l"(&table[rn & 0x7fffffffu]));
        }
        if (pos >= next) {
            bool complete = (run_begin >= start);  // its end (`next`) is <= pos < end
            if (RELAXED) xyzz_relaxed_normalise(acc);
            if (complete) xyzz_store(&buckets[b], acc);
            else xyzz_store(&partial[2 * t], acc);
            xyzz_set_inf(acc);
            do { b++; } while ((offsets[b + 1] >> shift) <= pos);
            run_begin = offsets[b] >> shift; next = offsets[b + 1] >> shift;
        }
        if (RELAXED) xyzz_madd_relaxed(acc, q); else if (LAZY) xyzz_madd(acc, q); else if (CALL) xyzz_madd_call(acc, q); else xyzz_madd_classic(acc, q);
        if (PREFETCH) { if (pos + 1 < end) q = qn; }
    }
    {
        bool complete = (run_begin >= start) && (next <= end);
        if (RELAXED) xyzz_relaxed_normalise(acc);
        if (complete) xyzz_store(&buckets[b], acc);
        else if (run_begin <= start) xyzz_store(&partial[2 * t], acc);
        else xyzz_store(&partial[2 * t + 1], acc);
    }
}

template <int MINB, bool CALL, bool PREFETCH, bool DIRECT>
__global__ void __launch_bounds__(128, MINB) k_accumulate_t(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                                                              const Affine* __restrict__ table, uint32_t nb, uint32_t acc_threads,
                                                              uint32_t chunk, int shift,
                                                              XYZZ* __restrict__ buckets, XYZZ* __restrict__ partial) {
    accumulate_body<CALL, PREFETCH, DIRECT>(sorted, offsets, table, nb, acc_threads, chunk, shift, buckets, partial);
}
// XYZZ += affine with the lazily reduced Y3 (9 Montgomery reductions per addition instead of 10)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_accumulate_lazy(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                                                                 const Affine* __restrict__ table, uint32_t nb, uint32_t acc_threads,
                                                                 uint32_t chunk, int shift,
                                                                 XYZZ* __restrict__ buckets, XYZZ* __restrict__ partial) {
    accumulate_body<false, false, false, true>(sorted, offsets, table, nb, acc_threads, chunk, shift, buckets, partial);
}
// ... and with the accumulator kept in [0, 2p) (no conditional subtraction after any product; normalised when
// it leaves the loop)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_accumulate_relaxed(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                                                                    const Affine* __restrict__ table, uint32_t nb, uint32_t acc_threads,
                                                                    uint32_t chunk, int shift,
                                                                    XYZZ* __restrict__ buckets, XYZZ* __restrict__ partial) {
    accumulate_body<false, false, false, true, true>(sorted, offsets, table, nb, acc_threads, chunk, shift, buckets, partial);
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_accumulate_relaxed_pf(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                                                                       const Affine* __restrict__ table, uint32_t nb, uint32_t acc_threads,
                                                                       uint32_t chunk, int shift,
                                                                       XYZZ* __restrict__ buckets, XYZZ* __restrict__ partial) {
    accumulate_body<false, false, false, true, true, true>(sorted, offsets, table, nb, acc_threads, chunk, shift, buckets, partial);
}
// Same body under a hard register cap instead of a blocks-per-SM hint: at 112 registers four blocks leave
// 8 K registers of every SM free, so the small blocks of the other lanes' sort kernels can be resident
// NEXT TO the accumulation instead of displacing it (ptxas: 112 registers, no spills; 96: 28 B of spills).
template <int MAXR>
__global__ void __maxnreg__(MAXR) k_accumulate_r(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                                                  const Affine* __restrict__ table, uint32_t nb, uint32_t acc_threads,
                                                  uint32_t chunk, int shift,
                                                  XYZZ* __restrict__ buckets, XYZZ* __restrict__ partial) {
    accumulate_body<false, false, false>(sorted, offsets, table, nb, acc_threads, chunk, shift, buckets, partial);
}

// ---- relaxed accumulation with the identity case peeled (experimental, opt-in: accumulate variants 28, 29) --------------
// In accumulate_body the "accumulator is still the identity" case of xyzz_madd_relaxed is if-converted by the compiler:
// every iteration copies the accumulator and materialises (q.x, q.y, 1, 1) before the branch (~80 register moves per
// addition in SASS).  Here the state "accumulator empty" is a flag, the first point of a run is installed by an
// out-of-line path the compiler cannot speculate, and the addition itself never tests for the identity.
// SQR (bit mask): 1 = PP = P^2, 2 = R^2 through the dedicated squaring (gen_field.py sqrnr: 36 instead of 64 wide multiplies
// in the product); 4 = a - b (+ 2p) with predicated additions instead of a masked 2p (sub2pp: 18 instead of 25 instructions);
// 8 = the loop head without a materialised zero point and with the digit's sign applied by predicated subtractions
template <int SQR>
__device__ __forceinline__ bool xyzz_madd_relaxed_nonempty(XYZZ& acc, const Affine& q) {  // false: the sum is the identity
    Fq U2, S2, Pp, Rr, PP, PPP, Q, t;
#define SUB2P(r, a, b) do { if (SQR & 4) fq_sub2pp_ptx(r, a, b); else fq_sub2p_ptx(r, a, b); } while (0)
    fq_mulnr_ptx(U2.l, q.x.l, acc.zz.l);
    fq_mulnr_ptx(S2.l, q.y.l, acc.zzz.l);
    SUB2P(Pp.l, U2.l, acc.x.l);
    SUB2P(Rr.l, S2.l, acc.y.l);
    if (fq_is_zero_mod(Pp)) {
        if (!fq_is_zero_mod(Rr)) return false;
        xyzz_dbl_affine(acc, q);  // canonical result from the canonical q
        return !xyzz_is_inf(acc);
    }
    if (SQR & 1) fq_sqrnr_ptx(PP.l, Pp.l); else fq_mulnr_ptx(PP.l, Pp.l, Pp.l);
    fq_mulnr_ptx(PPP.l, Pp.l, PP.l);
    fq_mulnr_ptx(Q.l, acc.x.l, PP.l);
    fq_mulnr_ptx(acc.zz.l, acc.zz.l, PP.l);
    fq_mulnr_ptx(acc.zzz.l, acc.zzz.l, PPP.l);
    if (SQR & 2) fq_sqrnr_ptx(t.l, Rr.l); else fq_mulnr_ptx(t.l, Rr.l, Rr.l);
    SUB2P(t.l, t.l, PPP.l); SUB2P(t.l, t.l, Q.l); SUB2P(t.l, t.l, Q.l);  // X3
    SUB2P(Q.l, Q.l, t.l);
    fq_mul2subnr_ptx(acc.y.l, Rr.l, Q.l, acc.y.l, PPP.l);
    acc.x = t;
    return true;
#undef SUB2P
}
template <int MINB, int SQR>
__global__ void __launch_bounds__(128, MINB) k_accumulate_relaxed2(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                                                                     const Affine* __restrict__ table, uint32_t nb, uint32_t acc_threads,
                                                                     uint32_t chunk, XYZZ* __restrict__ buckets, XYZZ* __restrict__ partial) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= acc_threads) return;
    const uint32_t M = offsets[nb];
    uint64_t start64 = (uint64_t)t * chunk;
    if (start64 >= M) return;
    uint32_t start = (uint32_t)start64;
    uint32_t end = (uint32_t)min((uint64_t)M, start64 + chunk);
    uint32_t lo = 0, hi = nb;  // invariant: offsets[lo] <= start < offsets[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= start) lo = mid; else hi = mid;
    }
    uint32_t b = lo;
    uint32_t run_begin = offsets[b], next = offsets[b + 1];
    XYZZ acc; xyzz_set_inf(acc);
    bool nonempty = false;
    auto flush = [&](XYZZ* dst) {
        if (nonempty) xyzz_relaxed_normalise(acc); else xyzz_set_inf(acc);
        xyzz_store(dst, acc);
    };
    auto boundary = [&](uint32_t pos) {  // entry `pos` opens a new bucket run: store the finished one
        flush((run_begin >= start) ? &buckets[b] : &partial[2 * t]);
        nonempty = false;
        do { b++; } while (offsets[b + 1] <= pos);
        run_begin = offsets[b]; next = offsets[b + 1];
    };
    for (uint32_t pos = start; pos < end; pos++) {
        Affine q;
        if (SQR & 8) {
            // the padding slot never materialises a zero point, and the sign of the digit is applied by predicated
            // subtractions (fq_cneg_ptx) instead of a divergent branch around a masked negation
            uint32_t ref = sorted[pos];
            if (ref == REF_IDENT) {
                if (pos >= next) boundary(pos);
                continue;
            }
            q = aff_gather_ro(&table[ref & 0x7fffffffu]);
            if (pos >= next) boundary(pos);
            if (aff_is_inf(q)) continue;
            fq_cneg_ptx(q.y.l, q.y.l, ref & 0x80000000u);
        } else {
            q = load_point(table, sorted[pos]);
            if (pos >= next) boundary(pos);
            if (aff_is_inf(q)) continue;
        }
        if (!nonempty) {
            asm volatile("" ::: "memory");  // keep this path a real branch: nothing of it is worth speculating
            acc.x = q.x; acc.y = q.y; fe_one(acc.zz); fe_one(acc.zzz);
            nonempty = true;
            continue;
        }
        nonempty = xyzz_madd_relaxed_nonempty<SQR>(acc, q);
    }
    bool complete = (run_begin >= start) && (next <= end);
    flush(complete ? &buckets[b] : (run_begin <= start) ? &partial[2 * t] : &partial[2 * t + 1]);
}

// ---- FP64-pipe accumulation (experimental, opt-in: accumulate variants 23-26) ----------------------------
// The same fixed-size-chunk walk as accumulate_body with the accumulator held as 5 x 52-bit double limbs
// (ec_dfma.cuh): the products run on the FP64 pipe (DFMA hi/lo halves) and the ALU instead of the IMAD.WIDE pipe
// that bounds k_accumulate_relaxed.  Results are canonical XYZZ like every other variant, so k_bucket_fix and the
// reduction are unchanged.  (The walk is repeated here rather than templated into accumulate_body so that the
// production kernels stay byte-identical while this is being measured.)
__device__ __forceinline__ void accumulate_body_dfma(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                                                     const Affine* __restrict__ table, uint32_t nb, uint32_t acc_threads,
                                                     uint32_t chunk, XYZZ* __restrict__ buckets, XYZZ* __restrict__ partial) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= acc_threads) return;
    const uint32_t M = offsets[nb];
    uint64_t start64 = (uint64_t)t * chunk;
    if (start64 >= M) return;
    uint32_t start = (uint32_t)start64;
    uint32_t end = (uint32_t)min((uint64_t)M, start64 + chunk);
    uint32_t lo = 0, hi = nb;  // invariant: offsets[lo] <= start < offsets[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= start) lo = mid; else hi = mid;
    }
    uint32_t b = lo;
    uint32_t run_begin = offsets[b], next = offsets[b + 1];
    dfma::XYZZ5 acc; dfma::xyzz5_set_inf(acc);
    auto flush = [&](XYZZ* dst) {
        XYZZ out;
        dfma::xyzz5_to_xyzz(out, acc);
        xyzz_store(dst, out);
    };
    for (uint32_t pos = start; pos < end; pos++) {
        Affine q = load_point(table, sorted[pos]);
        if (pos >= next) {
            flush((run_begin >= start) ? &buckets[b] : &partial[2 * t]);
            dfma::xyzz5_set_inf(acc);
            do { b++; } while (offsets[b + 1] <= pos);
            run_begin = offsets[b]; next = offsets[b + 1];
        }
        dfma::xyzz5_madd(acc, q);
    }
    bool complete = (run_begin >= start) && (next <= end);
    flush(complete ? &buckets[b] : (run_begin <= start) ? &partial[2 * t] : &partial[2 * t + 1]);
}
// DFMA_OF4 of every 4 consecutive blocks accumulate on the FP64 pipe, the rest on the IMAD.WIDE pipe (the
// relaxed integer body): blocks of both kinds are resident on every SM, so the two pipes are busy at once.
// DFMA_OF4 = 4: every block on the FP64 pipe.
template <int DFMA_OF4, int MINB>
__global__ void __launch_bounds__(128, MINB) k_accumulate_hybrid(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                                                                   const Affine* __restrict__ table, uint32_t nb, uint32_t acc_threads,
                                                                   uint32_t chunk, XYZZ* __restrict__ buckets, XYZZ* __restrict__ partial) {
    if (DFMA_OF4 == 4 || (int)(blockIdx.x & 3) < DFMA_OF4) accumulate_body_dfma(sorted, offsets, table, nb, acc_threads, chunk, buckets, partial);
    else accumulate_body<false, false, false, true, true>(sorted, offsets, table, nb, acc_threads, chunk, 0, buckets, partial);
}

// Buckets whose entries span more than LONG_SPAN chunks (hot buckets: equal scalars, a top window with a
// few possible digits, adversarial blobs) are queued for k_bucket_fix_long instead of being summed by one
// thread: with 4 waves of chunks a single bucket can own thousands of partial sums.
static constexpr uint32_t LONG_SPAN = 48;
static constexpr int LONG_THREADS = 128;
static constexpr uint32_t LONG_BLOCKS = 296;

__global__ void __launch_bounds__(128) k_bucket_fix(const uint32_t* __restrict__ offsets, uint32_t nbuckets, uint32_t chunk, int shift,
                                                     XYZZ* __restrict__ buckets, const XYZZ* __restrict__ partial,
                                                     uint32_t* __restrict__ long_count, uint32_t* __restrict__ long_list,
                                                     uint32_t long_cap) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    uint32_t s = offsets[b] >> shift, e = offsets[b + 1] >> shift;
    if (s == e) {
        XYZZ inf; xyzz_set_inf(inf);
        xyzz_store(&buckets[b], inf);
        return;
    }
    uint32_t t_lo = s / chunk, t_hi = (e - 1) / chunk;
    if (t_lo == t_hi) return;  // written directly by k_accumulate
    if (t_hi - t_lo >= LONG_SPAN) {
        uint32_t slot = atomicAdd(long_count, 1u);
        if (slot < long_cap) { long_list[slot] = b; return; }  // (cap = every possible long bucket; never exceeded)
    }
    XYZZ acc; xyzz_set_inf(acc);
    for (uint32_t t = t_lo; t <= t_hi; t++) {
        uint32_t slot = ((uint64_t)s <= (uint64_t)t * chunk) ? 0u : 1u;
        XYZZ v = xyzz_load(&partial[2 * t + slot]);
        xyzz_add(acc, v);
    }
    xyzz_store(&buckets[b], acc);
}

// one block per queued bucket (grid-stride over the queue): strided partial sums per thread, then a
// shared-memory tree
__global__ void __launch_bounds__(LONG_THREADS) k_bucket_fix_long(const uint32_t* __restrict__ offsets, uint32_t chunk, int shift,
                                                                   XYZZ* __restrict__ buckets, const XYZZ* __restrict__ partial,
                                                                   const uint32_t* __restrict__ long_count,
                                                                   const uint32_t* __restrict__ long_list, uint32_t long_cap) {
    __shared__ XYZZ sm[LONG_THREADS];
    uint32_t count = *long_count;
    if (count > long_cap) count = long_cap;
    for (uint32_t q = blockIdx.x; q < count; q += gridDim.x) {
        const uint32_t b = long_list[q];
        const uint32_t s = offsets[b] >> shift, e = offsets[b + 1] >> shift;
        const uint32_t t_lo = s / chunk, t_hi = (e - 1) / chunk;
        XYZZ acc; xyzz_set_inf(acc);
        for (uint32_t t = t_lo + threadIdx.x; t <= t_hi; t += LONG_THREADS) {
            uint32_t slot = ((uint64_t)s <= (uint64_t)t * chunk) ? 0u : 1u;
            XYZZ v = xyzz_load(&partial[2 * t + slot]);
            xyzz_add(acc, v);
        }
        sm[threadIdx.x] = acc;
        __syncthreads();
        for (uint32_t w = LONG_THREADS / 2; w >= 1; w >>= 1) {
            if (threadIdx.x < w) {
                XYZZ x = sm[threadIdx.x], y = sm[threadIdx.x + w];
                xyzz_add(x, y);
                sm[threadIdx.x] = x;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) xyzz_store(&buckets[b], sm[0]);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// Batch-affine levels.  The padded sorted list is reduced pairwise, bucket-agnostically:
//   level l:  out[p] = in[2p] + in[2p+1]   for p < M >> (l+1)
// (every bucket starts on a multiple of 2^L, so a pair never straddles two buckets; padding is the
// identity).  Affine + affine needs 1/(x2 - x1): Montgomery's trick over the whole level --
//   k_ba_prefix  per thread: running products e_j of its K denominators (stored), then the product
//                of all OTHER threads' totals in the block (warp shuffles) and the block total
//   k_ba_invert  one block: inverse of every block total (one Fermat inversion per level)
//   k_ba_apply   per thread: u = 1/(own total); walks its pairs backwards: 1/d_j = e_{j-1} * u,
//                u *= d_j; lambda, x3, y3 (3 more multiplications)
// = 6 Fq-mul per addition + ~15/K for the block scan, against 10 for XYZZ += affine.
// Exceptional pairs are exact: identity operands pass through (d = 1), P + P uses d = 2y and the
// tangent slope, P + (-P) gives the identity (d = 1).
// ---------------------------------------------------------------------------------
static constexpr int BA_THREADS = 128;
enum { BA_ADD = 0, BA_DBL = 1, BA_PASS_A = 2, BA_PASS_B = 3, BA_INF = 4 };

struct BaArgs {
    const uint32_t* refs;    // level 0: padded sorted refs; deeper levels: nullptr
    const Affine* in;        // level 0: table; deeper: points of the previous level
    Affine* out;
    Fq* prefix;              // one per pair
    Fq* others;              // one per thread: product of the other threads' totals in its block
    Fq* blk_tot;             // one per block
    Fq* blk_inv;
    const uint32_t* m_ptr;   // padded entry count at level 0 (device)
    int level;
    int K;                   // pairs per thread
};

__device__ __forceinline__ int ba_classify(const Affine& a, const Affine& b, Fq& d) {
    bool ai = aff_is_inf(a), bi = aff_is_inf(b);
    if (ai || bi) { fe_one(d); return ai ? (bi ? BA_INF : BA_PASS_B) : BA_PASS_A; }
    fe_sub(d, b.x, a.x);
    if (!fe_is_zero(d)) return BA_ADD;
    if (fe_eq(a.y, b.y) && !fe_is_zero(a.y)) { fe_dbl(d, a.y); return BA_DBL; }
    fe_one(d);
    return BA_INF;
}

template <bool L0>
__device__ __forceinline__ void ba_load_pair(const BaArgs& g, uint32_t p, Affine& a, Affine& b) {
    if (L0) {
        uint2 r = *reinterpret_cast<const uint2*>(g.refs + 2 * (size_t)p);
        a = load_point(g.in, r.x);
        b = load_point(g.in, r.y);
    } else {
        a = aff_load_ro(&g.in[2 * (size_t)p]);
        b = aff_load_ro(&g.in[2 * (size_t)p + 1]);
    }
}

// denominator only: the y coordinates are fetched just for the exceptional pairs
template <bool L0>
__device__ __forceinline__ int ba_pair_denominator(const BaArgs& g, uint32_t p, Fq& d) {
    const Fq *pax, *pbx;
    bool a_id = false, b_id = false;
    if (L0) {
        uint2 r = *reinterpret_cast<const uint2*>(g.refs + 2 * (size_t)p);
        a_id = r.x == REF_IDENT; b_id = r.y == REF_IDENT;
        pax = &g.in[r.x & 0x7fffffffu].x; pbx = &g.in[r.y & 0x7fffffffu].x;
    } else {
        pax = &g.in[2 * (size_t)p].x; pbx = &g.in[2 * (size_t)p + 1].x;
    }
    if (!a_id && !b_id) {
        Fq ax = fe_load_ro(pax), bx = fe_load_ro(pbx);
        fe_sub(d, bx, ax);
        if (!fe_is_zero(ax) && !fe_is_zero(bx) && !fe_is_zero(d)) return BA_ADD;
    }
    Affine a, b;
    ba_load_pair<L0>(g, p, a, b);
    return ba_classify(a, b, d);
}

__device__ __forceinline__ Fq fq_shfl_up(const Fq& v, int off) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], off);
    return r;
}
__device__ __forceinline__ Fq fq_shfl_down(const Fq& v, int off) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], off);
    return r;
}

// For every thread of a block of NW warps: the product of the totals of all OTHER threads; returns
// the block total too (valid in every thread).  smem: NW Fq.
template <int NW>
__device__ __forceinline__ void block_product_of_others(const Fq& T, Fq* smem, Fq& others, Fq& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Fq pre = T, suf = T;
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        Fq o = fq_shfl_up(pre, off);
        if (lane >= off) fe_mul(pre, pre, o);
    }
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        Fq o = fq_shfl_down(suf, off);
        if (lane + off < 32) fe_mul(suf, suf, o);
    }
    if (lane == 31) smem[wid] = pre;  // warp total
    Fq pe = fq_shfl_up(pre, 1), se = fq_shfl_down(suf, 1);
    Fq one; fe_one(one);
    if (lane == 0) pe = one;
    if (lane == 31) se = one;
    fe_mul(others, pe, se);
    __syncthreads();
    Fq ow;  // product of the other warps' totals
    if (NW <= 4) {
        ow = one;
#pragma unroll 1
        for (int v = 0; v < NW; v++) {
            if (v != wid) { Fq w = smem[v]; fe_mul(ow, ow, w); }
        }
        fe_mul(total, ow, smem[wid]);
    } else {  // NW == 32: second-level shuffle scan over the warp totals (every warp does it redundantly)
        Fq wt = smem[lane < NW ? lane : 0];
        if (lane >= NW) wt = one;
        Fq wp = wt, ws = wt;
#pragma unroll 1
        for (int off = 1; off < 32; off <<= 1) {
            Fq o = fq_shfl_up(wp, off);
            if (lane >= off) fe_mul(wp, wp, o);
        }
#pragma unroll 1
        for (int off = 1; off < 32; off <<= 1) {
            Fq o = fq_shfl_down(ws, off);
            if (lane + off < 32) fe_mul(ws, ws, o);
        }
        Fq wpe = fq_shfl_up(wp, 1), wse = fq_shfl_down(ws, 1);
        if (lane == 0) wpe = one;
        if (lane == 31) wse = one;
        Fq mine; fe_mul(mine, wpe, wse);  // lane v: product of the totals of all warps but v
        // broadcast lane `wid`'s value and the grand total (lane 31's inclusive prefix)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ow.l[i] = __shfl_sync(0xffffffffu, mine.l[i], wid);
            total.l[i] = __shfl_sync(0xffffffffu, wp.l[i], 31);
        }
    }
    fe_mul(others, others, ow);
    __syncthreads();
}

template <bool L0>
__global__ void __launch_bounds__(BA_THREADS) k_ba_prefix(BaArgs g) {
    __shared__ Fq sm[BA_THREADS / 32];
    const uint32_t P = (*g.m_ptr) >> (g.level + 1);
    const uint32_t base = blockIdx.x * (uint32_t)(BA_THREADS * g.K);
    if (base >= P) return;
    Fq e; fe_one(e);
    for (int j = 0; j < g.K; j++) {
        uint32_t p = base + (uint32_t)j * BA_THREADS + threadIdx.x;
        if (p < P) {
            Fq d;
            int cls = ba_pair_denominator<L0>(g, p, d);
            if (cls <= BA_DBL) fe_mul(e, e, d);
            fe_store(&g.prefix[p], e);
        }
    }
    Fq others, total;
    block_product_of_others<BA_THREADS / 32>(e, sm, others, total);
    fe_store(&g.others[blockIdx.x * BA_THREADS + threadIdx.x], others);
    if (threadIdx.x == 0) fe_store(&g.blk_tot[blockIdx.x], total);
}

// one block of 1024 threads: blk_inv[i] = 1 / blk_tot[i] for the active blocks of the level
static constexpr int BA_INV_THREADS = 1024;
static constexpr int BA_INV_MAXQ = 8;  // up to 8192 blocks per level
__global__ void __launch_bounds__(BA_INV_THREADS) k_ba_invert(BaArgs g) {
    __shared__ Fq sm[BA_INV_THREADS / 32];
    __shared__ Fq sm_inv;
    const uint32_t P = (*g.m_ptr) >> (g.level + 1);
    const uint32_t per_block = (uint32_t)(BA_THREADS * g.K);
    const uint32_t nblk = (P + per_block - 1) / per_block;
    const uint32_t q = (nblk + BA_INV_THREADS - 1) / BA_INV_THREADS;  // <= BA_INV_MAXQ (host guarantees)
    const uint32_t first = threadIdx.x * q;
    Fq v[BA_INV_MAXQ];
    Fq T; fe_one(T);
    for (uint32_t k = 0; k < q; k++) {
        if (first + k < nblk) { v[k] = fe_load(&g.blk_tot[first + k]); fe_mul(T, T, v[k]); }
        else fe_one(v[k]);
    }
    Fq others, total;
    block_product_of_others<BA_INV_THREADS / 32>(T, sm, others, total);
    if (threadIdx.x == 0) { Fq inv; fe_inv_fast(inv, total); sm_inv = inv; }
    __syncthreads();
    Fq u;  // 1 / T
    fe_mul(u, others, sm_inv);
    // local batch inversion, backwards: 1/v_k = (v_0..v_{k-1}) * u_k, u_{k-1} = u_k * v_k
    for (int k = (int)q - 1; k >= 0; k--) {
        Fq pre; fe_one(pre);
        for (int i = 0; i < k; i++) fe_mul(pre, pre, v[i]);
        Fq inv; fe_mul(inv, pre, u);
        if (first + k < nblk) fe_store(&g.blk_inv[first + k], inv);
        fe_mul(u, u, v[k]);
    }
}

template <bool L0>
__global__ void __launch_bounds__(BA_THREADS) k_ba_apply(BaArgs g) {
    const uint32_t P = (*g.m_ptr) >> (g.level + 1);
    const uint32_t base = blockIdx.x * (uint32_t)(BA_THREADS * g.K);
    if (base >= P) return;
    Fq u;
    {
        Fq o = fe_load(&g.others[blockIdx.x * BA_THREADS + threadIdx.x]);
        Fq bi = fe_load(&g.blk_inv[blockIdx.x]);
        fe_mul(u, o, bi);
    }
    for (int j = g.K - 1; j >= 0; j--) {
        uint32_t p = base + (uint32_t)j * BA_THREADS + threadIdx.x;
        if (p >= P) continue;
        Affine a, b, r;
        ba_load_pair<L0>(g, p, a, b);
        Fq d;
        int cls = ba_classify(a, b, d);
        if (cls >= BA_PASS_A) {
            if (cls == BA_PASS_A) r = a;
            else if (cls == BA_PASS_B) r = b;
            else aff_set_inf(r);
            aff_store(&g.out[p], r);
            continue;
        }
        Fq inv;
        if (j > 0) {
            Fq e = fe_load(&g.prefix[p - BA_THREADS]);  // e_{j-1}
            fe_mul(inv, e, u);
            fe_mul(u, u, d);
        } else {
            inv = u;
        }
        Fq lam, t;
        if (cls == BA_ADD) {
            fe_sub(t, b.y, a.y);
            fe_mul(lam, t, inv);
        } else {  // tangent: 3 x^2 / (2 y)
            fe_sqr(t, a.x);
            fe_dbl(lam, t); fe_add(t, lam, t);
            fe_mul(lam, t, inv);
        }
        fe_sqr(r.x, lam);
        fe_sub(r.x, r.x, a.x); fe_sub(r.x, r.x, b.x);
        fe_sub(t, a.x, r.x);
        fe_mul(t, lam, t);
        fe_sub(r.y, t, a.y);
        aff_store(&g.out[p], r);
    }
}

// slice j covers buckets [j*slice, (j+1)*slice) of one set; bucket index k (in set) has weight k+1
__global__ void __launch_bounds__(128) k_reduce_slices(const XYZZ* __restrict__ buckets, MsmPlan p,
                                                        XYZZ* __restrict__ slice_sums) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nslices = p.nbuckets / p.slice;
    if (j >= nslices) return;
    uint32_t half = 1u << (p.c - 1);
    uint32_t k0 = (j * p.slice) & (half - 1u);
    XYZZ run, acc;
    xyzz_set_inf(run); xyzz_set_inf(acc);
    for (int i = (int)p.slice - 1; i >= 0; i--) {
        XYZZ v = xyzz_load(&buckets[(size_t)j * p.slice + i]);
        xyzz_add(run, v);
        xyzz_add(acc, run);
    }
    if (k0 != 0 && !xyzz_is_inf(run)) {
        XYZZ m; xyzz_set_inf(m);
        int top = 31 - __clz(k0);
        for (int bit = top; bit >= 0; bit--) {
            xyzz_dbl(m, m);
            if ((k0 >> bit) & 1u) xyzz_add(m, run);
        }
        xyzz_add(acc, m);
    }
    xyzz_store(&slice_sums[j], acc);
}

// out[g] = sum of in[g*tile .. (g+1)*tile), tile a power of two <= 256, blockDim = tile/2 (>=1)
__global__ void k_tree_reduce(const XYZZ* __restrict__ in, XYZZ* __restrict__ out, uint32_t tile) {
    extern __shared__ uint4 smem_raw[];
    XYZZ* sm = reinterpret_cast<XYZZ*>(smem_raw);
    uint32_t tid = threadIdx.x;
    uint32_t hw = tile >> 1;  // == blockDim.x when tile >= 2
    const XYZZ* src = in + (size_t)blockIdx.x * tile;
    if (tile == 1) { if (tid == 0) xyzz_store(&out[blockIdx.x], xyzz_load(&src[0])); return; }
    XYZZ a = xyzz_load(&src[tid]);
    XYZZ b = xyzz_load(&src[tid + hw]);
    xyzz_add(a, b);
    sm[tid] = a;
    __syncthreads();
    for (uint32_t s = hw >> 1; s >= 1; s >>= 1) {
        if (tid < s) {
            XYZZ x = sm[tid];
            XYZZ y = sm[tid + s];
            xyzz_add(x, y);
            sm[tid] = x;
        }
        __syncthreads();
    }
    if (tid == 0) xyzz_store(&out[blockIdx.x], sm[0]);
}

// ---------------------------------------------------------------------------------
static int g_ba_levels = 0, g_ba_min_avg = 64, g_ba_k0 = 0;  // k0 = 0: pairs per thread chosen so each level is one wave
static bool g_tuning_env_read = false;
void msm_set_tuning(int ba_levels, int ba_min_avg_bucket, int ba_k0) {
    g_tuning_env_read = true;
    if (ba_levels >= 0) g_ba_levels = ba_levels > 4 ? 4 : ba_levels;
    if (ba_min_avg_bucket >= 0) g_ba_min_avg = ba_min_avg_bucket;
    if (ba_k0 >= 0) g_ba_k0 = ba_k0;
}
// blocks of k_ba_apply resident on the whole GPU (one wave)
static uint32_t ba_wave_blocks() {
    static uint32_t cached = 0;
    if (!cached) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ba_apply<true>, BA_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 6;
        cached = (uint32_t)per_sm * SM_COUNT;
    }
    return cached;
}
static void tuning_from_env() {
    if (g_tuning_env_read) return;
    g_tuning_env_read = true;
    const char* e;
    if ((e = getenv("KZGB_BA_LEVELS"))) g_ba_levels = atoi(e) > 4 ? 4 : (atoi(e) < 0 ? 0 : atoi(e));
    if ((e = getenv("KZGB_BA_MIN_AVG"))) g_ba_min_avg = atoi(e);
    if ((e = getenv("KZGB_BA_K0")) && atoi(e) >= 0) g_ba_k0 = atoi(e);
}

MsmPlan msm_make_plan(uint32_t n, int c, bool fixed_base, uint32_t table_stride, uint32_t base_offset, uint32_t batch, bool throughput) {
    tuning_from_env();
    MsmPlan p;
    p.c = c;
    p.W = (255 + c - 1) / c;
    p.batch_n = (fixed_base && batch > 0) ? n / batch : 0;  // `batch` independent MSMs over the same table, one bucket set each
    p.sets = fixed_base ? (p.batch_n ? (int)batch : 1) : p.W;
    p.nbuckets = (uint32_t)p.sets << (c - 1);
    p.n = n;
    p.table_stride = fixed_base ? table_stride : 0;
    p.base_offset = fixed_base ? base_offset : 0;
    uint64_t entries = (uint64_t)n * p.W;
    // batch-affine levels pay off when buckets are well filled (padding <= 2^L - 1 slots per bucket)
    p.ba_levels = 0;
    if (g_ba_levels > 0 && entries >= (uint64_t)g_ba_min_avg * p.nbuckets && entries <= ((uint64_t)1 << 27)) p.ba_levels = g_ba_levels;
    uint64_t padded = entries + (uint64_t)p.nbuckets * ((1u << p.ba_levels) - 1u);
    p.max_entries = (uint32_t)padded;
    for (int l = 0; l < 4; l++) {
        uint64_t pairs = padded >> (l + 1);
        uint64_t k;
        if (g_ba_k0 > 0) { k = (uint64_t)g_ba_k0 >> l; if (k < 4) k = 4; }
        else {  // one full wave of the apply kernel, at least 8 pairs per thread
            uint64_t slots = (uint64_t)ba_wave_blocks() * BA_THREADS;
            k = (pairs + slots - 1) / slots;
            if (k < 8) k = 8;
        }
        uint64_t kmin = (pairs + (uint64_t)BA_THREADS * BA_INV_THREADS * BA_INV_MAXQ - 1) / ((uint64_t)BA_THREADS * BA_INV_THREADS * BA_INV_MAXQ);
        if (k < kmin) k = kmin;
        p.ba_k[l] = (uint32_t)k;
    }
    uint64_t tail = padded >> p.ba_levels;
    static int waves = -1;
    if (waves < 0) { const char* e = getenv("KZGB_ACC_WAVES"); waves = e ? atoi(e) : 4; if (waves < 1) waves = 1; }
    uint64_t max_threads = (uint64_t)SM_COUNT * 512 * waves;  // 4 blocks of 128 threads per SM and wave
    uint64_t want = (tail + 15) / 16;
    if (want < 1) want = 1;
    p.acc_threads = (uint32_t)(want < max_threads ? want : max_threads);
    p.chunk = (uint32_t)((tail + p.acc_threads - 1) / p.acc_threads);
    if (p.chunk == 0) p.chunk = 1;
    uint32_t half = 1u << (c - 1);
    // Buckets per bucket-reduce thread.  A thread does 2 XYZZ additions per bucket plus one small scalar
    // multiplication per slice, so longer slices mean less work but a longer serial chain.  Measured on the
    // 2^19 pipeline (profiles/r01_sweep_slice.txt): 4 -> 283, 8 -> 299, 16 -> 304, 32 -> 306 blobs/s (the
    // fat, latency-bound reduce blocks displace accumulation blocks while they are resident); alone, one
    // MSM is fastest at 8 (1.975 ms) and 0.3 ms slower at 32.  So: 32 for pipelines (throughput), 8 for single calls.
    p.slice = half >= 1024 ? 4 : (half >= 4 ? 2 : 1);
    if (half >= 2048) p.slice = 8;
    if (throughput && half >= 8192) p.slice = 32;
    {  // KZGB_SLICE overrides (power of two)
        static int slice_env = -1;
        if (slice_env < 0) { const char* e = getenv("KZGB_SLICE"); slice_env = e ? atoi(e) : 0; }
        if (slice_env > 0 && (slice_env & (slice_env - 1)) == 0 && (uint32_t)slice_env * 256u <= half) p.slice = (uint32_t)slice_env;
    }
    return p;
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static uint32_t ba_max_blocks(const MsmPlan& p) {
    uint32_t mb = 1;
    for (int l = 0; l < p.ba_levels; l++) {
        uint64_t pairs = (uint64_t)p.max_entries >> (l + 1);
        uint64_t per = (uint64_t)BA_THREADS * p.ba_k[l];
        uint32_t nb = (uint32_t)((pairs + per - 1) / per);
        if (nb > mb) mb = nb;
    }
    return mb;
}

size_t msm_workspace_bytes(const MsmPlan& p) {
    size_t b = 0;
    b += align_up((size_t)p.n * sizeof(Fr));
    b += align_up(((size_t)p.nbuckets + 1) * 4);
    b += align_up((size_t)p.nbuckets * 4);
    b += align_up((size_t)p.max_entries * 4);
    b += align_up((size_t)p.nbuckets * sizeof(XYZZ));
    b += align_up((size_t)2 * p.acc_threads * sizeof(XYZZ));
    b += align_up((size_t)(p.nbuckets / p.slice) * sizeof(XYZZ));
    b += align_up((size_t)(p.nbuckets / p.slice) * sizeof(XYZZ));  // tree ping-pong
    b += align_up((size_t)p.sets * sizeof(XYZZ));
    b += align_up(1024 * 4);  // scan tile totals
    b += align_up(((size_t)p.acc_threads / LONG_SPAN + 8) * 4);  // long-bucket queue (+ its counter)
    if (p.ba_levels) {
        uint32_t mb = ba_max_blocks(p);
        b += align_up((size_t)(p.max_entries / 2 + 1) * sizeof(Affine));
        b += align_up((size_t)(p.max_entries / 4 + 1) * sizeof(Affine));
        b += align_up((size_t)(p.max_entries / 2 + 1) * sizeof(Fq));
        b += align_up((size_t)mb * BA_THREADS * sizeof(Fq));
        b += 2 * align_up((size_t)mb * sizeof(Fq));
    }
    return b;
}

void msm_workspace_carve(const MsmPlan& p, void* base, MsmWorkspace* ws) {
    char* c = (char*)base;
    ws->canon = (Fr*)c; c += align_up((size_t)p.n * sizeof(Fr));
    ws->hist = (uint32_t*)c; c += align_up(((size_t)p.nbuckets + 1) * 4);
    ws->cursor = (uint32_t*)c; c += align_up((size_t)p.nbuckets * 4);
    ws->sorted = (uint32_t*)c; c += align_up((size_t)p.max_entries * 4);
    ws->buckets = (XYZZ*)c; c += align_up((size_t)p.nbuckets * sizeof(XYZZ));
    ws->partial = (XYZZ*)c; c += align_up((size_t)2 * p.acc_threads * sizeof(XYZZ));
    ws->slice_sums = (XYZZ*)c; c += 2 * align_up((size_t)(p.nbuckets / p.slice) * sizeof(XYZZ));
    ws->set_sums = (XYZZ*)c; c += align_up((size_t)p.sets * sizeof(XYZZ));
    ws->tile_tot = (uint32_t*)c; c += align_up(1024 * 4);
    ws->long_list = (uint32_t*)c; c += align_up(((size_t)p.acc_threads / LONG_SPAN + 8) * 4);
    ws->ba_pts[0] = ws->ba_pts[1] = nullptr;
    ws->ba_prefix = ws->ba_others = ws->ba_blk_tot = ws->ba_blk_inv = nullptr;
    if (p.ba_levels) {
        uint32_t mb = ba_max_blocks(p);
        ws->ba_pts[0] = (Affine*)c; c += align_up((size_t)(p.max_entries / 2 + 1) * sizeof(Affine));
        ws->ba_pts[1] = (Affine*)c; c += align_up((size_t)(p.max_entries / 4 + 1) * sizeof(Affine));
        ws->ba_prefix = (Fq*)c; c += align_up((size_t)(p.max_entries / 2 + 1) * sizeof(Fq));
        ws->ba_others = (Fq*)c; c += align_up((size_t)mb * BA_THREADS * sizeof(Fq));
        ws->ba_blk_tot = (Fq*)c; c += align_up((size_t)mb * sizeof(Fq));
        ws->ba_blk_inv = (Fq*)c;
    }
}

void msm_launch(const MsmPlan& p, const MsmWorkspace& ws, const Fr* scalars, bool scalars_canonical,
                const Affine* table, cudaStream_t st, cudaEvent_t ev_acc_begin, cudaEvent_t ev_acc_end,
                cudaStream_t st_acc, cudaEvent_t ev_fork, cudaEvent_t ev_join) {
    const int L = p.ba_levels;
    const bool split = st_acc && ev_fork && ev_join;
    cudaStream_t sa = split ? st_acc : st;
    cudaMemsetAsync(ws.hist, 0, ((size_t)p.nbuckets + 1) * 4, st);
    if (L) cudaMemsetAsync(ws.sorted, 0xff, (size_t)p.max_entries * 4, st);  // padding slots = REF_IDENT
    uint32_t gb = (p.n + 255) / 256;
    if (p.n) { k_digits_hist<<<gb, 256, 0, st>>>(scalars, scalars_canonical, p, ws.canon, ws.hist); g_launch_count += 2; }
    g_launch_count += 4;  // scan, accumulate, bucket_fix, reduce_slices
    {
        uint32_t ntiles = (p.nbuckets + SCAN_TILE - 1) / SCAN_TILE;
        k_scan_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(ws.hist, ws.cursor, p.nbuckets, ws.tile_tot, ntiles == 1, (1u << L) - 1u);
        if (ntiles > 1) {
            k_scan_tile_totals<<<1, SCAN_THREADS, 0, st>>>(ws.tile_tot, ntiles, ws.hist, p.nbuckets);
            k_scan_add<<<(p.nbuckets + 255) / 256, 256, 0, st>>>(ws.hist, ws.cursor, p.nbuckets, ws.tile_tot);
            g_launch_count += 2;
        }
    }
    if (p.n) k_scatter<<<gb, 256, 0, st>>>(ws.canon, p, ws.cursor, ws.sorted);
    if (split) { cudaEventRecord(ev_fork, st); cudaStreamWaitEvent(sa, ev_fork, 0); }
    if (ev_acc_begin) cudaEventRecord(ev_acc_begin, sa);
    const Affine* tail_src = table;
    for (int l = 0; l < L; l++) {
        BaArgs g;
        g.refs = l == 0 ? ws.sorted : nullptr;
        g.in = l == 0 ? table : ws.ba_pts[(l - 1) & 1];
        g.out = ws.ba_pts[l & 1];
        g.prefix = ws.ba_prefix; g.others = ws.ba_others; g.blk_tot = ws.ba_blk_tot; g.blk_inv = ws.ba_blk_inv;
        g.m_ptr = ws.hist + p.nbuckets;
        g.level = l;
        g.K = (int)p.ba_k[l];
        uint64_t pairs = (uint64_t)p.max_entries >> (l + 1);
        uint64_t per = (uint64_t)BA_THREADS * p.ba_k[l];
        uint32_t nblk = (uint32_t)((pairs + per - 1) / per);
        if (nblk == 0) nblk = 1;
        if (l == 0) {
            k_ba_prefix<true><<<nblk, BA_THREADS, 0, sa>>>(g);
            k_ba_invert<<<1, BA_INV_THREADS, 0, sa>>>(g);
            k_ba_apply<true><<<nblk, BA_THREADS, 0, sa>>>(g);
        } else {
            k_ba_prefix<false><<<nblk, BA_THREADS, 0, sa>>>(g);
            k_ba_invert<<<1, BA_INV_THREADS, 0, sa>>>(g);
            k_ba_apply<false><<<nblk, BA_THREADS, 0, sa>>>(g);
        }
        g_launch_count += 3;
        tail_src = g.out;
    }
    {
        static int variant = -1;
        if (variant < 0) { const char* e = getenv("KZGB_ACC_VARIANT"); variant = e ? atoi(e) : 0; }
        dim3 g((p.acc_threads + 127) / 128);
#define KZ_ACC(MB, CALL, PF, DIRECT) \
    k_accumulate_t<MB, CALL, PF, DIRECT><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial)
        if (L) {
            KZ_ACC(4, false, false, true);
        } else {
            switch (variant) {  // sweep in profiles/r01_accumulate_variants.txt
                case 1: KZ_ACC(3, false, true, false); break;
                case 3: KZ_ACC(4, true, true, false); break;
                case 9: KZ_ACC(5, false, true, false); break;
                case 15: k_accumulate_lazy<4><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial); break;
                case 19: k_accumulate_relaxed_pf<4><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial); break;
                case 21: k_accumulate_relaxed<5><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial); break;
                case 17: k_accumulate_relaxed<3><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial); break;
                case 11: k_accumulate_r<112><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial); break;
                case 12: k_accumulate_r<96><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial); break;
                case 13: k_accumulate_r<104><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial); break;
                case 14: k_accumulate_r<80><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial); break;
                case 20: KZ_ACC(4, false, false, false); break;  // every product reduced on its own (the default until the lazy Y3)
                // FP64-pipe accumulation (ec_dfma.cuh): 23 = every block, 24/25/26 = 1/2/3 of every 4 blocks, the rest integer
                case 23: k_accumulate_hybrid<4, 3><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 24: k_accumulate_hybrid<1, 3><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 25: k_accumulate_hybrid<2, 3><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 26: k_accumulate_hybrid<3, 3><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 27: k_accumulate_hybrid<2, 4><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                // k_accumulate_relaxed2<blocks/SM, mask>: the production arithmetic with fewer non-multiply instructions (opt-in until
                // timed; modelled cycles per addition in profiles/r01_accumulate_sass_census.txt, production 5560): 28 = identity case
                // peeled (5489), 34 = + predicated subtractions (5453), 31 = + PP squaring (5417), 32 = + R^2 squaring (5472),
                // 29 = + both squarings (5402), 30 = 29 at 3 blocks/SM (5317), 33 = 31 + 34 (5358), 35 = 33 + lean loop head (5327)
                case 28: k_accumulate_relaxed2<4, 0><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 29: k_accumulate_relaxed2<4, 3><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 30: k_accumulate_relaxed2<3, 3><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 31: k_accumulate_relaxed2<4, 1><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 32: k_accumulate_relaxed2<4, 2><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 33: k_accumulate_relaxed2<4, 5><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 34: k_accumulate_relaxed2<4, 4><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 35: k_accumulate_relaxed2<4, 13><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, ws.buckets, ws.partial); break;
                case 16: k_accumulate_lazy<3><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial); break;
                default: k_accumulate_relaxed<4><<<g, 128, 0, sa>>>(ws.sorted, ws.hist, tail_src, p.nbuckets, p.acc_threads, p.chunk, L, ws.buckets, ws.partial); break;
            }
        }
#undef KZ_ACC
    }
    if (ev_acc_end) cudaEventRecord(ev_acc_end, sa);
    if (split) { cudaEventRecord(ev_join, sa); cudaStreamWaitEvent(st, ev_join, 0); }
    {
        const uint32_t long_cap = p.acc_threads / LONG_SPAN + 4;  // a long bucket owns >= LONG_SPAN chunks
        uint32_t* long_count = ws.long_list;                      // [0] = counter, [1..] = queue
        cudaMemsetAsync(long_count, 0, 4, st);
        k_bucket_fix<<<(p.nbuckets + 127) / 128, 128, 0, st>>>(ws.hist, p.nbuckets, p.chunk, L, ws.buckets, ws.partial,
                                                                long_count, ws.long_list + 1, long_cap);
        k_bucket_fix_long<<<LONG_BLOCKS, LONG_THREADS, 0, st>>>(ws.hist, p.chunk, L, ws.buckets, ws.partial, long_count,
                                                                  ws.long_list + 1, long_cap);
        g_launch_count++;
    }
    uint32_t nslices = p.nbuckets / p.slice;
    k_reduce_slices<<<(nslices + 127) / 128, 128, 0, st>>>(ws.buckets, p, ws.slice_sums);
    // tree-reduce each set's slice results down to one point
    uint32_t per_set = nslices / p.sets;  // power of two
    XYZZ* ping = ws.slice_sums;
    XYZZ* pong = (XYZZ*)((char*)ws.slice_sums + align_up((size_t)nslices * sizeof(XYZZ)));
    while (per_set > 1) {
        uint32_t tile = per_set < 256 ? per_set : 256;
        uint32_t groups = (per_set / tile) * p.sets;
        k_tree_reduce<<<groups, tile / 2, (tile / 2) * sizeof(XYZZ), st>>>(ping, pong, tile);
        g_launch_count++;
        XYZZ* t = ping; ping = pong; pong = t;
        per_set /= tile;
    }
    cudaMemcpyAsync(ws.set_sums, ping, (size_t)p.sets * sizeof(XYZZ), cudaMemcpyDeviceToDevice, st);
}

}  // namespace kzgb
