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

/* Device time (ms, CUDA events on `stream`) of a chain of pt_batch_track_device_async calls issued from C: device-wide
 * synchronise, event,
 * nseg calls (track_fn = address of pt_batch_track_device_async of the loaded product library; segment s tracks Ts[s]
 * steps from bases[s]), event, wait for the second event.  Issuing the three from C keeps interpreter time between
 * the first event and the launch out of the device-timed region (with an idle GPU the first event completes at once
 * and everything the host does before the launch is counted).  The caller synchronises before and after. */
PTB_API int ptb_time_chain(void *track_fn, void *batch, int nseg, const void *const *bases, const int *Ts,
                           size_t step_stride, size_t frame_stride, size_t pitch, void *stream, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* PAWSOME_BENCH_H */
