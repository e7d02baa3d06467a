/*
 * kzg_bn254_b200_bench.h -- measurement hooks of libkzgbn254_b200.so (bench.py, scripts/).  Not part of the
 * drop-in boundary (include/kzg_bn254_b200.h): nothing here computes a result a caller of the reference would ask
 * for; they time the library's own kernels with CUDA events or report what it launched.
 */
#ifndef KZG_BN254_B200_BENCH_H
#define KZG_BN254_B200_BENCH_H

#include "kzg_bn254_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Runs an integer-pipe microbenchmark and returns achieved operations per second.
 * kind 0: IMAD, 1: IMAD.WIDE, 2: carry-chained IMAD.WIDE.X, 3: Fq Montgomery multiplications. */
int kzgb_microbench(kzgb_ctx* ctx, int kind, double* ops_per_second);
/* Times `reps` back-to-back SRS MSMs of n points on device-resident scalars with CUDA events on the
 * context's stream; returns mean milliseconds per MSM of the whole pipeline and of the bucket
 * accumulation kernel alone. */
int kzgb_bench_msm(kzgb_ctx* ctx, size_t n, int reps, double* ms_total, double* ms_accumulate);
/* Times `reps` (I)NTTs of `batch` transforms of size 2^logn back to back on device-resident data
 * (alternating forward / inverse) with CUDA events; returns mean milliseconds per batched call. */
int kzgb_bench_ntt(kzgb_ctx* ctx, int logn, size_t batch, int reps, double* ms_per_call);
/* Number of kernels this library has launched since it was loaded. */
uint64_t kzgb_launch_count(const kzgb_ctx* ctx);
/* Device-side stopwatch (CUDA events) spanning every stream of the context: all work queued between
 * begin and end lies inside the measured interval. */
int kzgb_timer_begin(kzgb_ctx* ctx);
int kzgb_timer_end(kzgb_ctx* ctx, double* ms_out);
/* Bucket-accumulation kernel statistics since the last reset: summed CUDA-event duration of the
 * launches (each bracketed on its own stream), launch count, and point additions performed. */
int kzgb_stats(kzgb_ctx* ctx, double* acc_ms, uint64_t* acc_launches, uint64_t* acc_point_adds, int reset);
/* Timeline of a batch call without a system profiler: between begin and end every MSM of every lane leaves
 * device events (0 sort begins, 1 accumulate begins, 2 accumulate ends, 3 MSM ends) and host time stamps
 * (10 MSM enqueued, 11 lane woke up with the result, 12 result finished on the host).  kzgb_trace_end writes 4 doubles per
 * record -- lane, kind, device ms since begin (-1 for host records), host ms since begin (-1 for device-only records)
 * -- up to capacity_records, and the number of records taken. */
int kzgb_trace_begin(kzgb_ctx* ctx);
int kzgb_trace_end(kzgb_ctx* ctx, double* out, size_t capacity_records, size_t* n_records);

#ifdef __cplusplus
}
#endif
#endif /* KZG_BN254_B200_BENCH_H */
