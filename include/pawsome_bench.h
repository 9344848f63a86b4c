/*
 * pawsome_bench.h — C ABI of libpawsome_bench.so: measurement helpers used by bench.py and tools/ ONLY.
 *
 * Deliberately NOT part of libpawsome_cuda.so (include/pawsome.h): the product library holds the hand-written
 * kernels of the DoG-window + argmax path and nothing else.  These helpers measure the device (the FP32-pipe
 * roofline denominator SURVEY §8d asks to be measured, an issue-port probe) and flush L2 between timed repeats.
 */
#ifndef PAWSOME_BENCH_H
#define PAWSOME_BENCH_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PTB_API __attribute__((visibility("default")))
#else
#define PTB_API
#endif

/* FP32 throughput of `device` in TFLOP/s (2 flops per FMA): packed=0 plain FFMA, packed=1 fma.rn.f32x2,
 * packed=2 add.rn.f32x2 counted as 1 flop per lane-add (is a packed add cheaper than two scalar ones?).
 * Best of `reps` timed launches.  Returns 0 or a negative code. */
PTB_API int ptb_measure_fp32_peak(int device, int packed, int reps, double *tflops);
/* Issue-port probe: time (ms) of a loop of packed FFMA2 with na in {0,4,8} independent integer ops per 8 FFMA2. */
PTB_API int ptb_probe_ffma2_issue(int device, int na, double *ms_out);
/* Overwrite `bytes` of device scratch on `stream` (L2 flush between timed repeats). */
PTB_API int ptb_flush_l2(void *scratch, size_t bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PAWSOME_BENCH_H */
