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
#include <cstdlib>
#include "kzgb_internal.hpp"

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
            uint32_t key = (p.sets == 1 ? 0u : (uint32_t)w * half) + mag - 1u;
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

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor,
                                                             uint32_t nb, uint32_t* __restrict__ block_tot, bool single) {
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
            uint32_t key = (p.sets == 1 ? 0u : (uint32_t)w * half) + mag - 1u;
            uint32_t pos = atomicAdd(&cursor[key], 1u);
            uint32_t ref = (p.sets == 1) ? ((uint32_t)w * p.table_stride + p.base_offset + i) : i;
            sorted[pos] = ref | (neg ? 0x80000000u : 0u);
        }
    }
}

__device__ __forceinline__ Affine load_point(const Affine* __restrict__ table, uint32_t ref) {
    Affine q = aff_load_ro(&table[ref & 0x7fffffffu]);
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

template <int MINB, bool CALL, bool PREFETCH>
__global__ void __launch_bounds__(128, MINB) k_accumulate_t(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                                                              const Affine* __restrict__ table, MsmPlan p,
                                                              XYZZ* __restrict__ buckets, XYZZ* __restrict__ partial) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.acc_threads) return;
    const uint32_t nb = p.nbuckets;
    const uint32_t M = offsets[nb];
    uint64_t start64 = (uint64_t)t * p.chunk;
    if (start64 >= M) return;  // k_bucket_fix only reads slots of chunks that hold entries
    uint32_t start = (uint32_t)start64;
    uint32_t end = (uint32_t)min((uint64_t)M, start64 + p.chunk);
    // largest b with offsets[b] <= start
    uint32_t lo = 0, hi = nb;  // invariant: offsets[lo] <= start < offsets[hi] (offsets[nb] = M > start)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= start) lo = mid; else hi = mid;
    }
    uint32_t b = lo;
    uint32_t run_begin = offsets[b], next = offsets[b + 1];
    XYZZ acc; xyzz_set_inf(acc);
    Affine q;
    if (PREFETCH) q = load_point(table, sorted[start]);
    for (uint32_t pos = start; pos < end; pos++) {
        Affine qn;
        if (PREFETCH) { if (pos + 1 < end) qn = load_point(table, sorted[pos + 1]); }  // next point in flight during this addition
        else q = load_point(table, sorted[pos]);
        if (pos >= next) {
            bool complete = (run_begin >= start);  // its end (`next`) is <= pos < end
            if (complete) xyzz_store(&buckets[b], acc);
            else xyzz_store(&partial[2 * t], acc);
            xyzz_set_inf(acc);
            do { b++; } while (offsets[b + 1] <= pos);
            run_begin = offsets[b]; next = offsets[b + 1];
        }
        if (CALL) xyzz_madd_call(acc, q); else xyzz_madd(acc, q);
        if (PREFETCH) { if (pos + 1 < end) q = qn; }
    }
    {
        bool complete = (run_begin >= start) && (next <= end);
        if (complete) xyzz_store(&buckets[b], acc);
        else if (run_begin <= start) xyzz_store(&partial[2 * t], acc);
        else xyzz_store(&partial[2 * t + 1], acc);
    }
}

__global__ void __launch_bounds__(128) k_bucket_fix(const uint32_t* __restrict__ offsets, MsmPlan p,
                                                     XYZZ* __restrict__ buckets, const XYZZ* __restrict__ partial) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.nbuckets) return;
    uint32_t s = offsets[b], e = offsets[b + 1];
    if (s == e) {
        XYZZ inf; xyzz_set_inf(inf);
        xyzz_store(&buckets[b], inf);
        return;
    }
    uint32_t t_lo = s / p.chunk, t_hi = (e - 1) / p.chunk;
    if (t_lo == t_hi) return;  // written directly by k_accumulate
    XYZZ acc; xyzz_set_inf(acc);
    for (uint32_t t = t_lo; t <= t_hi; t++) {
        uint32_t slot = ((uint64_t)s <= (uint64_t)t * p.chunk) ? 0u : 1u;
        XYZZ v = xyzz_load(&partial[2 * t + slot]);
        xyzz_add(acc, v);
    }
    xyzz_store(&buckets[b], acc);
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
MsmPlan msm_make_plan(uint32_t n, int c, bool fixed_base, uint32_t table_stride, uint32_t base_offset) {
    MsmPlan p;
    p.c = c;
    p.W = (255 + c - 1) / c;
    p.sets = fixed_base ? 1 : p.W;
    p.nbuckets = (uint32_t)p.sets << (c - 1);
    p.n = n;
    p.table_stride = fixed_base ? table_stride : 0;
    p.base_offset = fixed_base ? base_offset : 0;
    uint64_t entries = (uint64_t)n * p.W;
    uint64_t max_threads = (uint64_t)SM_COUNT * 512;
    uint64_t want = (entries + 15) / 16;
    if (want < 1) want = 1;
    p.acc_threads = (uint32_t)(want < max_threads ? want : max_threads);
    p.chunk = (uint32_t)((entries + p.acc_threads - 1) / p.acc_threads);
    if (p.chunk == 0) p.chunk = 1;
    uint32_t half = 1u << (c - 1);
    p.slice = half >= 1024 ? 4 : (half >= 4 ? 2 : 1);
    return p;
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

size_t msm_workspace_bytes(const MsmPlan& p) {
    size_t b = 0;
    b += align_up((size_t)p.n * sizeof(Fr));
    b += align_up(((size_t)p.nbuckets + 1) * 4);
    b += align_up((size_t)p.nbuckets * 4);
    b += align_up((size_t)p.n * p.W * 4);
    b += align_up((size_t)p.nbuckets * sizeof(XYZZ));
    b += align_up((size_t)2 * p.acc_threads * sizeof(XYZZ));
    b += align_up((size_t)(p.nbuckets / p.slice) * sizeof(XYZZ));
    b += align_up((size_t)(p.nbuckets / p.slice) * sizeof(XYZZ));  // tree ping-pong
    b += align_up((size_t)p.sets * sizeof(XYZZ));
    b += align_up(1024 * 4);  // scan tile totals
    return b;
}

struct TreeScratch { XYZZ* pong; };

void msm_workspace_carve(const MsmPlan& p, void* base, MsmWorkspace* ws) {
    char* c = (char*)base;
    ws->canon = (Fr*)c; c += align_up((size_t)p.n * sizeof(Fr));
    ws->hist = (uint32_t*)c; c += align_up(((size_t)p.nbuckets + 1) * 4);
    ws->cursor = (uint32_t*)c; c += align_up((size_t)p.nbuckets * 4);
    ws->sorted = (uint32_t*)c; c += align_up((size_t)p.n * p.W * 4);
    ws->buckets = (XYZZ*)c; c += align_up((size_t)p.nbuckets * sizeof(XYZZ));
    ws->partial = (XYZZ*)c; c += align_up((size_t)2 * p.acc_threads * sizeof(XYZZ));
    ws->slice_sums = (XYZZ*)c; c += 2 * align_up((size_t)(p.nbuckets / p.slice) * sizeof(XYZZ));
    ws->set_sums = (XYZZ*)c; c += align_up((size_t)p.sets * sizeof(XYZZ));
    ws->tile_tot = (uint32_t*)c;
}

void msm_launch(const MsmPlan& p, const MsmWorkspace& ws, const Fr* scalars, bool scalars_canonical,
                const Affine* table, cudaStream_t st, cudaEvent_t ev_acc_begin, cudaEvent_t ev_acc_end) {
    cudaMemsetAsync(ws.hist, 0, ((size_t)p.nbuckets + 1) * 4, st);
    uint32_t gb = (p.n + 255) / 256;
    if (p.n) { k_digits_hist<<<gb, 256, 0, st>>>(scalars, scalars_canonical, p, ws.canon, ws.hist); g_launch_count += 2; }
    g_launch_count += 4;  // scan, accumulate, bucket_fix, reduce_slices
    {
        uint32_t ntiles = (p.nbuckets + SCAN_TILE - 1) / SCAN_TILE;
        k_scan_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(ws.hist, ws.cursor, p.nbuckets, ws.tile_tot, ntiles == 1);
        if (ntiles > 1) {
            k_scan_tile_totals<<<1, SCAN_THREADS, 0, st>>>(ws.tile_tot, ntiles, ws.hist, p.nbuckets);
            k_scan_add<<<(p.nbuckets + 255) / 256, 256, 0, st>>>(ws.hist, ws.cursor, p.nbuckets, ws.tile_tot);
            g_launch_count += 2;
        }
    }
    if (p.n) k_scatter<<<gb, 256, 0, st>>>(ws.canon, p, ws.cursor, ws.sorted);
    if (ev_acc_begin) cudaEventRecord(ev_acc_begin, st);
    {
        static int variant = -1;
        if (variant < 0) { const char* e = getenv("KZGB_ACC_VARIANT"); variant = e ? atoi(e) : 0; }
        dim3 g((p.acc_threads + 127) / 128);
#define KZ_ACC(MB, CALL, PF) k_accumulate_t<MB, CALL, PF><<<g, 128, 0, st>>>(ws.sorted, ws.hist, table, p, ws.buckets, ws.partial)
        switch (variant) {
            case 1: KZ_ACC(3, false, true); break;
            case 8: KZ_ACC(4, false, false); break;
            case 9: KZ_ACC(5, false, true); break;
            case 2: KZ_ACC(3, true, true); break;
            case 3: KZ_ACC(4, true, true); break;
            case 4: KZ_ACC(5, true, true); break;
            case 5: KZ_ACC(5, true, false); break;
            case 6: KZ_ACC(6, true, false); break;
            case 7: KZ_ACC(4, true, false); break;
            default: KZ_ACC(4, false, false); break;  // best of the sweep in profiles/r01_accumulate_variants.txt
        }
#undef KZ_ACC
    }
    if (ev_acc_end) cudaEventRecord(ev_acc_end, st);
    k_bucket_fix<<<(p.nbuckets + 127) / 128, 128, 0, st>>>(ws.hist, p, ws.buckets, ws.partial);
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
