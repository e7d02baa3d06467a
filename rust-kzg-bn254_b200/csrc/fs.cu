// Fiat-Shamir challenges on the device for BATCHES OF SMALL BLOBS (verify_blob_kzg_proof_batch):
//   z_i = SHA-256("EIGENDA_FSBLOBVERIFY_V1_" || u64_be(n) || n x Fr_be || serialize_compressed(C_i)) mod r
// (reference primitives/src/helpers.rs:411-472, :382-390), one thread per blob, straight from the
// evaluations that are already resident for the barycentric evaluation.  SHA-256 is sequential per
// message, so this only pays when there are thousands of messages (4096 x 128 KiB in config 5); a
// single 16 MiB transcript stays on a host core with SHA-NI.
// Also derives, per blob, the one inverse the inversion-free evaluation kernels need:
//   tinv_i = 1 / (z_i^n - 1), or z_i / n when z_i is in the domain (see eval_quotient_launch).
#include <cstdlib>
#include "kzgb_internal.hpp"

namespace kzgb {

__constant__ uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// one compression: w[0..15] = the block as big-endian words (destroyed)
__device__ __forceinline__ void sha256_block(uint32_t h[8], uint32_t w[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int t = 0; t < 64; t++) {
        if (t >= 16) {
            uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
            uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            w[t & 15] = w[t & 15] + s0 + w[(t + 9) & 15] + s1;
        }
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA_K[t] + w[t & 15];
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// canonical big-endian words of a Montgomery element: word j = limb 7-j
__device__ __forceinline__ void fr_to_be_words(uint32_t* w, const Fr& mont) {
    Fr c; fe_from_mont(c, mont);
#pragma unroll
    for (int j = 0; j < 8; j++) w[j] = c.l[7 - j];
}

__global__ void __launch_bounds__(64) k_fs_challenges(const Fr* __restrict__ evals, uint32_t n, int logn, uint32_t batch,
                                                       const uint8_t* __restrict__ commit32, Fr ninv,
                                                       Fr* __restrict__ z_out, Fr* __restrict__ tinv_out,
                                                       uint32_t* __restrict__ in_domain_out, bool force_flag) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= batch) return;
    const Fr* f = evals + (size_t)k * n;
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t w[16];
    // item 0: "EIGENDA_FSBLOBVERIFY_V1_" || u64_be(n)
    w[0] = 0x45494745; w[1] = 0x4e44415f; w[2] = 0x4653424c; w[3] = 0x4f425645; w[4] = 0x52494659; w[5] = 0x5f56315f;
    w[6] = 0; w[7] = n;
    // items 1..n: evaluations; item n+1: the commitment (arkworks compressed, taken as bytes)
    fr_to_be_words(w + 8, fe_load_ro(&f[0]));
    sha256_block(h, w);
    uint32_t i = 1;
    for (; i + 1 < n; i += 2) {
        fr_to_be_words(w, fe_load_ro(&f[i]));
        fr_to_be_words(w + 8, fe_load_ro(&f[i + 1]));
        sha256_block(h, w);
    }
    uint32_t cw[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint8_t* p = commit32 + (size_t)k * 32 + 4 * j;
        cw[j] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
    const uint64_t bits = (64ull + 32ull * n) * 8ull;
    if (i < n) {  // n even (>= 2): evaluation n-1 and the commitment share a block, padding gets its own
        fr_to_be_words(w, fe_load_ro(&f[i]));
#pragma unroll
        for (int j = 0; j < 8; j++) w[8 + j] = cw[j];
        sha256_block(h, w);
#pragma unroll
        for (int j = 0; j < 16; j++) w[j] = 0;
        w[0] = 0x80000000u;
    } else {      // n == 1: the commitment opens the last block, padding follows it
#pragma unroll
        for (int j = 0; j < 8; j++) w[j] = cw[j];
#pragma unroll
        for (int j = 8; j < 16; j++) w[j] = 0;
        w[8] = 0x80000000u;
    }
    w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits;
    sha256_block(h, w);
    // hash_to_field_element: digest as a big-endian integer mod r, to Montgomery form
    Fr z;
#pragma unroll
    for (int j = 0; j < 8; j++) z.l[7 - j] = h[j];
    fe_to_mont(z, z);
    fe_store(&z_out[k], z);
    Fr zn = z, one;
    for (int s = 0; s < logn; s++) fe_sqr(zn, zn);
    fe_one(one);
    Fr t;
    const bool in_domain = fe_eq(zn, one);
    if (in_domain) fe_mul(t, z, ninv);
    else { fe_sub(zn, zn, one); fe_inv_fast(t, zn); }
    fe_store(&tinv_out[k], t);
    if (in_domain_out) in_domain_out[k] = (in_domain || force_flag) ? 1u : 0u;
}

// ---- four lanes per blob ----------------------------------------------------------------------------------
// One thread per blob leaves the GPU almost empty (4096 blobs = 128 warps) and every thread also pays for
// the Montgomery -> canonical conversions and the message schedules of its blocks.  The compression rounds
// of one message are sequential, but the schedule W_t + K_t of a block does not depend on the chaining
// state: lanes 4k..4k+3 own blob k, each prepares one of four consecutive blocks (conversion + full
// schedule, 64 words in registers), then lane 4k runs the 4 x 64 rounds, pulling the words of the other
// three lanes with shuffles.  Same digest, ~2.4x fewer issue slots per blob and 4x the warps.
__device__ __forceinline__ void sha256_schedule_k(uint32_t wk[64], uint32_t w[16]) {
#pragma unroll
    for (int t = 0; t < 64; t++) {
        if (t >= 16) {
            uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
            uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            w[t & 15] = w[t & 15] + s0 + w[(t + 9) & 15] + s1;
        }
        wk[t] = w[t & 15] + SHA_K[t];
    }
}

__global__ void __launch_bounds__(128) k_fs_challenges_quad(const Fr* __restrict__ evals, uint32_t n, int logn, uint32_t batch,
                                                             const uint8_t* __restrict__ commit32, Fr ninv,
                                                             Fr* __restrict__ z_out, Fr* __restrict__ tinv_out,
                                                             uint32_t* __restrict__ in_domain_out, bool force_flag) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t k = tid >> 2, sub = tid & 3u;
    const bool valid = k < batch;
    const uint32_t lane = threadIdx.x & 31u, qbase = lane & ~3u;
    const Fr* f = evals + (size_t)(valid ? k : 0) * n;
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t w[16];
    if (valid && sub == 0) {  // block 0: "EIGENDA_FSBLOBVERIFY_V1_" || u64_be(n) || evaluation 0
        w[0] = 0x45494745; w[1] = 0x4e44415f; w[2] = 0x4653424c; w[3] = 0x4f425645; w[4] = 0x52494659; w[5] = 0x5f56315f;
        w[6] = 0; w[7] = n;
        fr_to_be_words(w + 8, fe_load_ro(&f[0]));
        sha256_block(h, w);
    }
    // middle blocks b = 1 .. n/2 - 1: evaluations 2b-1 and 2b
    const uint32_t middle = n >= 2 ? n / 2 - 1 : 0;
    for (uint32_t b0 = 1; b0 <= middle; b0 += 4) {  // uniform trip count across the warp (n is a kernel argument)
        const uint32_t b = b0 + sub;
        const bool mine = valid && b <= middle;
        uint32_t wk[64];
        if (mine) {
            fr_to_be_words(w, fe_load_ro(&f[2 * b - 1]));
            fr_to_be_words(w + 8, fe_load_ro(&f[2 * b]));
            sha256_schedule_k(wk, w);
        } else {
#pragma unroll
            for (int t = 0; t < 64; t++) wk[t] = 0;
        }
        const uint32_t cnt = min(4u, middle - b0 + 1);
        uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], ff = h[5], g = h[6], hh = h[7];
#pragma unroll 1
        for (uint32_t j = 0; j < 4; j++) {
#pragma unroll
            for (int t = 0; t < 64; t++) {
                uint32_t x = __shfl_sync(0xffffffffu, wk[t], qbase + j);
                uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
                uint32_t ch = (e & ff) ^ (~e & g);
                uint32_t t1 = hh + S1 + ch + x;
                uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
                uint32_t mj = (a & bb) ^ (a & c) ^ (bb & c);
                uint32_t t2 = S0 + mj;
                hh = g; g = ff; ff = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
            }
            if (j < cnt) {  // close block b0 + j (only lane 4k keeps a meaningful state)
                h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += ff; h[6] += g; h[7] += hh;
            }
            a = h[0]; bb = h[1]; c = h[2]; d = h[3]; e = h[4]; ff = h[5]; g = h[6]; hh = h[7];
        }
    }
    if (!valid || sub != 0) return;
    uint32_t cw[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint8_t* p = commit32 + (size_t)k * 32 + 4 * j;
        cw[j] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
    const uint64_t bits = (64ull + 32ull * n) * 8ull;
    if (n >= 2) {  // evaluation n-1 and the commitment share a block, padding gets its own
        fr_to_be_words(w, fe_load_ro(&f[n - 1]));
#pragma unroll
        for (int j = 0; j < 8; j++) w[8 + j] = cw[j];
        sha256_block(h, w);
#pragma unroll
        for (int j = 0; j < 16; j++) w[j] = 0;
        w[0] = 0x80000000u;
    } else {       // n == 1: the commitment opens the last block, padding follows it
#pragma unroll
        for (int j = 0; j < 8; j++) w[j] = cw[j];
#pragma unroll
        for (int j = 8; j < 16; j++) w[j] = 0;
        w[8] = 0x80000000u;
    }
    w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits;
    sha256_block(h, w);
    Fr z;
#pragma unroll
    for (int j = 0; j < 8; j++) z.l[7 - j] = h[j];
    fe_to_mont(z, z);
    fe_store(&z_out[k], z);
    Fr zn = z, one;
    for (int s = 0; s < logn; s++) fe_sqr(zn, zn);
    fe_one(one);
    Fr t;
    const bool in_domain = fe_eq(zn, one);
    if (in_domain) fe_mul(t, z, ninv);
    else { fe_sub(zn, zn, one); fe_inv_fast(t, zn); }
    fe_store(&tinv_out[k], t);
    if (in_domain_out) in_domain_out[k] = (in_domain || force_flag) ? 1u : 0u;
}

// ---- one WARP per LONG transcript (16 MiB blobs) -----------------------------------------------------------------
// SHA-256 of one message cannot be spread over threads: block b needs the state after block b - 1, about 64 x 16
// dependent instructions, which makes a 16 MiB transcript ~150-250 ms of one warp against 9 ms of a SHA-NI core.  As
// LATENCY that is useless, as THROUGHPUT it is nearly free: a warp takes 1/2368 of the GPU's issue slots, so dozens of
// transcripts can be hashed next to the MSMs -- which is what a box with 8 GPUs and 32 host threads needs, where the
// hosts' aggregate SHA-256 rate (38 GB/s) is below what the GPUs consume (8 x 350 blobs/s x 16 MiB = 47 GB/s).
// Per iteration the 32 lanes each load one of 32 consecutive 64-byte blocks straight from the resident blob BYTES
// (the transcript hashes the canonical big-endian evaluations, which ARE the blob's bytes; a chunk >= r is reduced in
// place, helpers.rs:32-34) and expand its message schedule W_t + K_t into registers; then the 32 x 64 rounds run on
// the whole warp in lockstep, round inputs pulled from lane j by shuffles -- only the chaining state is sequential.
// Output: the MIDSTATE after tag || u64_be(n) || chunks 0 .. n-2 (the host appends chunk n-1, which shares its block with
// the commitment, and the commitment itself: challenge_finish).  Messages: blobs of exactly 32 n bytes, n = 2^k >= 4.
// done[k] (mapped host memory) is set to 1 when state[8k .. 8k+8) is final; *cancel != 0 makes every warp stop early.
__device__ __forceinline__ void fs_reduce_chunk_be(uint32_t* w) {  // w[0..8): big-endian words of a 256-bit value -> value mod r
    const uint32_t rm[8] = FR_MOD_LIMBS;                           // little-endian limbs of r
    if (w[0] < 0x30644e72u) return;                                // below r's top word: canonical (the common case)
    for (int it = 0; it < 6; it++) {
        bool ge = true;                                            // w >= r ?
        for (int j = 0; j < 8; j++) {
            uint32_t a = w[j], b = rm[7 - j];
            if (a != b) { ge = a > b; break; }
        }
        if (!ge) return;
        uint32_t borrow = 0;
        for (int j = 7; j >= 0; j--) {
            uint64_t d = (uint64_t)w[j] - rm[7 - j] - borrow;
            w[j] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
    }
}
static constexpr int FSL_WARPS = 4;  // messages per block: 4 warps x ~128 registers = the footprint of ONE accumulate block
__global__ void __launch_bounds__(32 * FSL_WARPS) k_fs_midstate_long(const uint8_t* const* __restrict__ blobs, const uint32_t* __restrict__ ns,
                                                                      uint32_t count, uint32_t* __restrict__ state,
                                                                      volatile uint32_t* __restrict__ done, const volatile uint32_t* __restrict__ cancel) {
    const uint32_t k = blockIdx.x * FSL_WARPS + (threadIdx.x >> 5);
    if (k >= count) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = ns[k];
    const uint4* src = reinterpret_cast<const uint4*>(blobs[k]);  // 16-byte units; chunk i = units 2i, 2i + 1
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t w[16];
    auto load_chunk = [&](uint32_t* dst, uint32_t chunk) {
        uint4 u0 = __ldg(src + 2 * (size_t)chunk), u1 = __ldg(src + 2 * (size_t)chunk + 1);
        dst[0] = __byte_perm(u0.x, 0, 0x0123); dst[1] = __byte_perm(u0.y, 0, 0x0123);
        dst[2] = __byte_perm(u0.z, 0, 0x0123); dst[3] = __byte_perm(u0.w, 0, 0x0123);
        dst[4] = __byte_perm(u1.x, 0, 0x0123); dst[5] = __byte_perm(u1.y, 0, 0x0123);
        dst[6] = __byte_perm(u1.z, 0, 0x0123); dst[7] = __byte_perm(u1.w, 0, 0x0123);
        fs_reduce_chunk_be(dst);
    };
    {   // block 0: "EIGENDA_FSBLOBVERIFY_V1_" || u64_be(n) || chunk 0   (every lane redundantly: keeps the warp converged)
        w[0] = 0x45494745; w[1] = 0x4e44415f; w[2] = 0x4653424c; w[3] = 0x4f425645; w[4] = 0x52494659; w[5] = 0x5f56315f;
        w[6] = 0; w[7] = n;
        load_chunk(w + 8, 0);
        sha256_block(h, w);
    }
    const uint32_t middle = n / 2 - 1;  // blocks b = 1 .. middle hold chunks 2b - 1, 2b
    for (uint32_t b0 = 1; b0 <= middle; b0 += 32) {
        if (((b0 >> 5) & 63u) == 0) {  // every 2048 blocks (128 KiB): one lane looks, the warp decides together
            uint32_t stop = lane == 0 ? *cancel : 0u;
            if (__shfl_sync(0xffffffffu, stop, 0)) return;
        }
        const uint32_t b = b0 + lane;
        uint32_t wk[64];
        if (b <= middle) {
            load_chunk(w, 2 * b - 1);
            load_chunk(w + 8, 2 * b);
            sha256_schedule_k(wk, w);
        } else {
#pragma unroll
            for (int t = 0; t < 64; t++) wk[t] = 0;
        }
        const uint32_t cnt = min(32u, middle - b0 + 1);
        uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], ff = h[5], g = h[6], hh = h[7];
#pragma unroll 1
        for (uint32_t j = 0; j < cnt; j++) {
#pragma unroll
            for (int t = 0; t < 64; t++) {
                const uint32_t x = __shfl_sync(0xffffffffu, wk[t], j);
                const uint32_t pre = hh + x, dp = d + pre;  // off the critical path: hh and d are known rounds ahead
                const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
                const uint32_t ch = (e & ff) ^ (~e & g);
                const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
                const uint32_t mj = (a & bb) ^ (a & c) ^ (bb & c);
                const uint32_t t1 = pre + S1 + ch;
                hh = g; g = ff; ff = e; e = dp + S1 + ch; d = c; c = bb; bb = a; a = t1 + S0 + mj;
            }
            h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += ff; h[6] += g; h[7] += hh;
            a = h[0]; bb = h[1]; c = h[2]; d = h[3]; e = h[4]; ff = h[5]; g = h[6]; hh = h[7];
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 8; j++) state[8 * (size_t)k + j] = h[j];
        __threadfence_system();
        done[k] = 1u;
    }
}
// One LANE per transcript: 32 messages per warp, each lane with its own chaining state, schedule and rounds -- the
// formulation whose cost per message is 1/32 of the warp-per-message kernel above (about 1.5 k instructions per 64-byte
// block and lane, 4 x 10^8 warp instructions per 32 transcripts of 16 MiB = 1 % of what the MSMs of those blobs issue),
// at the same latency: the rounds of one message are a dependent chain either way.  Each lane streams its own blob
// (64 B per block, the next block's 64 B requested before the current one is compressed, an L2 prefetch four blocks
// ahead); lanes whose transcript is shorter simply stop earlier.  One warp per block so that the warps spread over SMs.
__global__ void __launch_bounds__(32) k_fs_midstate_lanes(const uint8_t* const* __restrict__ blobs, const uint32_t* __restrict__ ns,
                                                           uint32_t count, uint32_t* __restrict__ state,
                                                           volatile uint32_t* __restrict__ done, const volatile uint32_t* __restrict__ cancel) {
    const uint32_t k = blockIdx.x * 32u + threadIdx.x;
    const bool live = k < count;
    const uint32_t n = live ? ns[k] : 0u;
    const uint4* src = live ? reinterpret_cast<const uint4*>(blobs[k]) : nullptr;  // 16-byte units; chunk i = units 2i, 2i + 1
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t w[16];
    auto swap_chunk = [&](uint32_t* dst, const uint4& u0, const uint4& u1) {
        dst[0] = __byte_perm(u0.x, 0, 0x0123); dst[1] = __byte_perm(u0.y, 0, 0x0123);
        dst[2] = __byte_perm(u0.z, 0, 0x0123); dst[3] = __byte_perm(u0.w, 0, 0x0123);
        dst[4] = __byte_perm(u1.x, 0, 0x0123); dst[5] = __byte_perm(u1.y, 0, 0x0123);
        dst[6] = __byte_perm(u1.z, 0, 0x0123); dst[7] = __byte_perm(u1.w, 0, 0x0123);
        fs_reduce_chunk_be(dst);
    };
    if (live) {  // block 0: "EIGENDA_FSBLOBVERIFY_V1_" || u64_be(n) || chunk 0
        w[0] = 0x45494745; w[1] = 0x4e44415f; w[2] = 0x4653424c; w[3] = 0x4f425645; w[4] = 0x52494659; w[5] = 0x5f56315f;
        w[6] = 0; w[7] = n;
        swap_chunk(w + 8, __ldg(src), __ldg(src + 1));
        sha256_block(h, w);
    }
    const uint32_t middle = live ? n / 2 - 1 : 0u;  // blocks b = 1 .. middle hold chunks 2b - 1, 2b = units 4b - 2 .. 4b + 1
    uint32_t longest = middle;
    for (int off = 16; off >= 1; off >>= 1) longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, off));
    uint4 cur[4];
    if (1u <= middle) {
#pragma unroll
        for (int j = 0; j < 4; j++) cur[j] = __ldg(src + 2 + j);
    }
#pragma unroll 1
    for (uint32_t b = 1; b <= longest; b++) {
        if ((b & 2047u) == 0) {  // every 128 KiB: one lane looks, the warp decides together
            uint32_t stop = threadIdx.x == 0 ? *cancel : 0u;
            if (__shfl_sync(0xffffffffu, stop, 0)) return;
        }
        uint4 nxt[4];
        if (b + 1 <= middle) {
            const uint4* q = src + 4 * (size_t)(b + 1) - 2;
#pragma unroll
            for (int j = 0; j < 4; j++) nxt[j] = __ldg(q + j);
            if (b + 5 <= middle) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 16));
        }
        if (b <= middle) {
            swap_chunk(w, cur[0], cur[1]);
            swap_chunk(w + 8, cur[2], cur[3]);
            sha256_block(h, w);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) cur[j] = nxt[j];
    }
    if (live) {
#pragma unroll
        for (int j = 0; j < 8; j++) state[8 * (size_t)k + j] = h[j];
        __threadfence_system();
        done[k] = 1u;
    }
}
static std::atomic<int> g_fsl_lanes{-1};  // option "device_hash_lanes": 1 one lane per transcript, 0 one warp per transcript, -1 the caller's model decides
void fs_set_midstate_lanes(int on) { g_fsl_lanes.store(on < 0 ? -1 : (on != 0)); }
int fs_midstate_lanes() { return g_fsl_lanes.load(); }
void fs_midstate_long_launch(const uint8_t* const* blobs_dev, const uint32_t* ns_dev, uint32_t count, uint32_t* state_mapped,
                             uint32_t* done_mapped, const uint32_t* cancel_mapped, bool lanes, cudaStream_t st) {
    if (!count) return;
    if (lanes)
        k_fs_midstate_lanes<<<(count + 31) / 32, 32, 0, st>>>(blobs_dev, ns_dev, count, state_mapped, done_mapped, cancel_mapped);
    else
        k_fs_midstate_long<<<(count + FSL_WARPS - 1) / FSL_WARPS, 32 * FSL_WARPS, 0, st>>>(blobs_dev, ns_dev, count, state_mapped, done_mapped, cancel_mapped);
    g_launch_count++;
}

// tests: flag every polynomial as "z in the domain" so the device-side choice takes the generic inverses everywhere
static std::atomic<int> g_fs_force_flag{0};
void fs_set_force_flag(int on) { g_fs_force_flag.store(on != 0); }
static std::atomic<int> g_fs_quad{1};  // 1: four lanes per blob (k_fs_challenges_quad), 0: one thread per blob
void fs_set_quad(int on) { g_fs_quad.store(on != 0); }

void fs_challenges_launch(const Fr* evals, uint32_t n, int logn, uint32_t batch, const uint8_t* commit32_dev,
                          const Fr* ninv_mont_host, Fr* z_out, Fr* tinv_out, cudaStream_t st, uint32_t* in_domain_out) {
    if (!batch) return;
    const int quad = g_fs_quad.load();
    if (quad && n >= 16)
        k_fs_challenges_quad<<<(batch * 4 + 127) / 128, 128, 0, st>>>(evals, n, logn, batch, commit32_dev, *ninv_mont_host, z_out, tinv_out, in_domain_out, g_fs_force_flag.load() != 0);
    else
        k_fs_challenges<<<(batch + 63) / 64, 64, 0, st>>>(evals, n, logn, batch, commit32_dev, *ninv_mont_host, z_out, tinv_out, in_domain_out, g_fs_force_flag.load() != 0);
    g_launch_count++;
}

}  // namespace kzgb
