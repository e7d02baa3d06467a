// Host SHA-256 for the Fiat-Shamir transcripts (reference primitives/src/helpers.rs:382-390,
// 411-472; verifier/src/batch.rs:76-168).  A single message hash is inherently sequential, so it
// stays on the host and is overlapped with the GPU work of the same blob (the 32-byte commitment
// is the LAST thing absorbed, so everything before it is hashed while the commitment MSM runs).
// Uses the SHA-NI extension when the CPU has it.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace kzgb {

struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t total;
    size_t fill;
    Sha256() { reset(); }
    void reset();
    void update(const void* data, size_t len);
    void finish(uint8_t out[32]);
};

void sha256(const void* data, size_t len, uint8_t out[32]);
bool sha256_has_shani();
// Sixteen equal-length messages in lockstep on AVX-512 (one message per 32-bit lane): absorbs nblk 64-byte blocks of
// message m, read from p[m], into st[m].  fix (optional) is called on a COPY of a block whose byte 0 or byte 32 is >= 0x30
// and may rewrite it before it is hashed (the Fiat-Shamir transcript hashes 32-byte values reduced mod r; a value >= r
// starts with a byte >= 0x30).  About twice the bytes per second of a core's SHA-NI; the 16 states are final together.
bool sha256_has_mb16();
void sha256_mb16_blocks(uint32_t st[16][8], const uint8_t* const p[16], size_t nblk, void (*fix)(uint8_t block[64]));

}  // namespace kzgb
