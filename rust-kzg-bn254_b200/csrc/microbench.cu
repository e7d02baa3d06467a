// Integer-pipe roofline microbenchmarks ("the per-SM IMAD/IMAD.WIDE rate on
// sm_100 must be measured ... and that measured figure used as the denominator").
//   mode 0: independent 32-bit IMAD chains                (mad.lo.u32)
//   mode 1: IMAD.WIDE.U32 with a 64-bit accumulate, no carry flags (mad.wide.u32, data-dependent operands)
//   mode 2: carry-chained IMAD.WIDE.U32.X                  (mad.lo.cc / madc.hi.cc pairs, as in fq_mul)
// and the achieved rate of the real 136-IMAD Montgomery multiplication with 2 independent
// chains per thread.
#include "kzgb_internal.hpp"

namespace kzgb {

template <int MODE>
__global__ void __launch_bounds__(256) k_imad_peak(uint32_t* __restrict__ sink, int iters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a = t * 2654435761u + 12345u, b = t ^ 0x9e3779b9u;
    if (MODE == 0) {
        uint32_t x0 = a, x1 = b, x2 = a + 1, x3 = b + 1, x4 = a + 2, x5 = b + 2, x6 = a + 3, x7 = b + 3;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                asm volatile("mad.lo.u32 %0, %0, %8, %9;\n\tmad.lo.u32 %1, %1, %8, %9;\n\t"
                             "mad.lo.u32 %2, %2, %8, %9;\n\tmad.lo.u32 %3, %3, %8, %9;\n\t"
                             "mad.lo.u32 %4, %4, %8, %9;\n\tmad.lo.u32 %5, %5, %8, %9;\n\t"
                             "mad.lo.u32 %6, %6, %8, %9;\n\tmad.lo.u32 %7, %7, %8, %9;"
                             : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7)
                             : "r"(a), "r"(b));
            }
        }
        sink[t] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    } else if (MODE == 1) {
        // pure IMAD.WIDE.U32 issue rate: 32x32->64 products of data-dependent operands, folded into a
        // 32-bit accumulator with one LOP3 (other pipe).  (A mad.wide accumulate chain with loop-invariant
        // operands gets strength-reduced by ptxas to 64-bit adds and then reports the IADD rate.)
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
        sink[t] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    } else {
        uint32_t x0 = a, x1 = b, x2 = a + 1, x3 = b + 1, x4 = a + 2, x5 = b + 2, x6 = a + 3, x7 = b + 3;
        uint32_t y0 = b, y1 = a, y2 = b + 5, y3 = a + 5, y4 = b + 6, y5 = a + 6, y6 = b + 7, y7 = a + 7;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                // two independent 4-pair carry chains = 8 IMAD.WIDE.U32(.X)
                asm volatile(
                    "mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.cc.u32 %1, %16, %17, %1;\n\t"
                    "madc.lo.cc.u32 %2, %16, %17, %2;\n\tmadc.hi.cc.u32 %3, %16, %17, %3;\n\t"
                    "madc.lo.cc.u32 %4, %16, %17, %4;\n\tmadc.hi.cc.u32 %5, %16, %17, %5;\n\t"
                    "madc.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.u32 %7, %16, %17, %7;\n\t"
                    "mad.lo.cc.u32 %8, %17, %16, %8;\n\tmadc.hi.cc.u32 %9, %17, %16, %9;\n\t"
                    "madc.lo.cc.u32 %10, %17, %16, %10;\n\tmadc.hi.cc.u32 %11, %17, %16, %11;\n\t"
                    "madc.lo.cc.u32 %12, %17, %16, %12;\n\tmadc.hi.cc.u32 %13, %17, %16, %13;\n\t"
                    "madc.lo.cc.u32 %14, %17, %16, %14;\n\tmadc.hi.u32 %15, %17, %16, %15;"
                    : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7), "+r"(y0),
                      "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7)
                    : "r"(a), "r"(b));
            }
        }
        sink[t] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7;
    }
}

__global__ void __launch_bounds__(256) k_fqmul_peak(uint32_t* __restrict__ sink, int iters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq a, b, c, d;
    for (int k = 0; k < 8; k++) {
        a.l[k] = t * 2654435761u + k; b.l[k] = (t ^ 0x9e3779b9u) + 7 * k;
        c.l[k] = t + 13 * k; d.l[k] = t * 31 + k;
    }
    a.l[7] &= 0x0fffffffu; b.l[7] &= 0x0fffffffu; c.l[7] &= 0x0fffffffu; d.l[7] &= 0x0fffffffu;
    for (int i = 0; i < iters; i++) {
        fe_mul(a, a, b);
        fe_mul(c, c, d);
        fe_mul(b, b, a);
        fe_mul(d, d, c);
    }
    uint32_t x = 0;
    for (int k = 0; k < 8; k++) x ^= a.l[k] ^ b.l[k] ^ c.l[k] ^ d.l[k];
    sink[t] = x;
}

// ops per thread: mode 0/1: iters*16*8 ; mode 2: iters*8*8 ; fqmul: iters*4
void imad_peak_launch(uint32_t* sink, int iters, int mode, int blocks, int threads, cudaStream_t st) {
    if (mode == 0) k_imad_peak<0><<<blocks, threads, 0, st>>>(sink, iters);
    else if (mode == 1) k_imad_peak<1><<<blocks, threads, 0, st>>>(sink, iters);
    else k_imad_peak<2><<<blocks, threads, 0, st>>>(sink, iters);
    g_launch_count++;
}
void fqmul_peak_launch(uint32_t* sink, int iters, int blocks, int threads, cudaStream_t st) {
    k_fqmul_peak<<<blocks, threads, 0, st>>>(sink, iters);
    g_launch_count++;
}

}  // namespace kzgb
