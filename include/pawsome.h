/*
 * pawsome.h — C ABI of libpawsome_cuda.so
 *
 * B200-native (sm_100a) implementation of ONE hot path of PawsomeTracker.jl:
 * the Difference-of-Gaussians filter over the constant-padded search window
 * followed by the argmax that yields the target's next position, plus the
 * large-window pass used for auto-detection.
 *
 * The reference has no FFI of its own (it is ~270 lines of Julia calling
 * ImageFiltering.jl); each entry point below cites the reference interface
 * (file:line under the reference repository) that a Julia `ccall` shim
 * replaces with it.  See INTEGRATION.md for the shim.
 *
 * Conventions
 *   - plain C types only; every function returns PT_OK (0) or a negative
 *     pt_status; pt_last_error() gives a thread-local message.
 *   - frames are row-major, H rows × W columns, `pitch` = bytes between rows
 *     for u8 frames / elements between rows for f32 frames.  This is the
 *     memory layout of the reference's frame type
 *     (PermutedDimsArray{Gray{N0f8},2,(2,1)} over a W×H Matrix,
 *     src/PawsomeTracker.jl:36).
 *   - (row, col) positions crossing the ABI are 1-based like the reference's
 *     CartesianIndex / NTuple{2,Int} (src/PawsomeTracker.jl:55-62).
 *   - there is NO CPU fallback: every compute entry point runs CUDA kernels
 *     and fails with PT_ERR_CUDA when no device is usable.
 *   - a handle is not thread-safe; distinct handles may be used from distinct
 *     host threads concurrently, on the same or on different devices (reference:
 *     one Tracker per track() call).  The library keeps no mutable process-wide
 *     state except a mutex-guarded pool of released buffers and a one-time,
 *     mutex-guarded per-device initialisation.
 */
#ifndef PAWSOME_H
#define PAWSOME_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define PT_API __attribute__((visibility("default")))
#else
#define PT_API
#endif

typedef enum pt_status {
    PT_OK = 0,
    PT_ERR_ARG = -1,         /* bad argument (the reference throws DimensionMismatch / AssertionError) */
    PT_ERR_CUDA = -2,        /* CUDA runtime error; text in pt_last_error() */
    PT_ERR_NOMEM = -3,
    PT_ERR_STATE = -4,       /* call sequence error, e.g. step before any frame/fill was set */
    PT_ERR_UNSUPPORTED = -5  /* shape outside what the kernels tile (kernel length too large for shared memory) */
} pt_status;

typedef enum pt_pixel {
    PT_PIX_U8 = 0,  /* Gray{N0f8}: value = u8/255 (what VideoIO delivers, src/PawsomeTracker.jl:157) */
    PT_PIX_F32 = 1  /* grayscale Float32 in [0,1] (the layout BASELINE.json's north_star streams) */
} pt_pixel;

typedef struct pt_batch pt_batch;     /* n independent Trackers advanced in lock-step */
typedef struct pt_tracker pt_tracker; /* one Tracker == a pt_batch of 1 */

/* ---- library ------------------------------------------------------------ */
PT_API int pt_version(void);
PT_API const char *pt_last_error(void); /* thread-local, never NULL */
PT_API int pt_device_count(void);       /* number of CUDA devices, or a negative pt_status */
/* Videos per batch that fill `device` evenly: two windows per SM (296 on a B200).  Multiples of it keep every SM
 * busy with two windows; batches between 1.18x and 2x the SM count are balanced by dog_window45_rot. */
PT_API int pt_preferred_batch(int device);

/* ---- scalar helpers (host arithmetic, no device needed) ------------------- */
/* get_sigma — src/PawsomeTracker.jl:30 */
PT_API double pt_sigma(double target_width);
/* length l of Kernel.DoG(σ) = 4⌈σ√2⌉+1 — call site src/PawsomeTracker.jl:43 */
PT_API int pt_kernel_len(double target_width);
/* guess_window_size — src/PawsomeTracker.jl:64-68 */
PT_API int pt_default_window(double target_width);
/* The FP32 1-D factors the kernels multiply with (each length pt_kernel_len):
 * row_p/row_m = narrow/wide Gaussian of the row pass, col_p/col_m = the
 * column-pass factors with the darker_target sign folded in
 * (response = Σ col_p·(row_p⋆P) + Σ col_m·(row_m⋆P)); src/PawsomeTracker.jl:41-43.
 * Any pointer may be NULL.  Returns l or a negative pt_status. */
PT_API int pt_factors_f32(double target_width, int darker_target,
                   float *row_p, float *row_m, float *col_p, float *col_m);

/* ---- batch of trackers ---------------------------------------------------- */
/*
 * Tracker(img, target_width, window_size, darker_target) for n videos of the
 * same geometry — src/PawsomeTracker.jl:39-52.  ws_rows/ws_cols is the
 * (rows, cols) window AFTER fix_window_size (:70-72); radii = ws .÷ 2 (:44).
 * No frame is attached yet and the fill values are unset.
 */
PT_API int pt_batch_create(int n, int H, int W, double target_width, int ws_rows, int ws_cols,
                    int darker_target, int pixel /* pt_pixel */, int device, pt_batch **out);
PT_API void pt_batch_destroy(pt_batch *b);

/* Re-target the window of an existing batch (the auto-detect pass builds a
 * Tracker with window size .÷ 4 and then a normal one, src/PawsomeTracker.jl:102-105). */
PT_API int pt_batch_set_window(pt_batch *b, int ws_rows, int ws_cols);

/* Copy n host frames (frames[v] → video v) into the batch's own HBM frame
 * store through pinned staging + cudaMemcpyAsync.  This is the write side of
 * `read!(vid, trckr.img.data)` (src/PawsomeTracker.jl:166).  pitch: bytes (u8)
 * or elements (f32) between rows of the HOST frames. */
PT_API int pt_batch_set_frames(pt_batch *b, const void *const *frames, size_t pitch);

/* Use frames already resident in HBM (no copy): video v's frame starts at
 * dev_base + v*frame_stride (bytes for u8, elements for f32). */
PT_API int pt_batch_bind_device_frames(pt_batch *b, const void *dev_base, size_t frame_stride, size_t pitch);

/* fillvalue = mode(_img) over the CURRENT frames, StatsBase tie rule, on the
 * device (256-bin histogram kernel) — src/PawsomeTracker.jl:47.  fills_out
 * (n ints, may be NULL) receives the u8 fill per video. */
PT_API int pt_batch_compute_fill(pt_batch *b, int *fills_out);
/* Or set it explicitly (u8 scale 0..255 for both pixel types). */
PT_API int pt_batch_set_fill(pt_batch *b, const int *fills);

/* Set the per-video guess (n×2 int32, 1-based row, col) kept on the device. */
PT_API int pt_batch_set_guess(pt_batch *b, const int32_t *guess_ij);

/*
 * (trckr::Tracker)(guess) for all n videos in ONE launch —
 * src/PawsomeTracker.jl:55-62: window = guess ± radii, DoG response over the
 * window of the constant-padded frame, first maximum in column-major order,
 * mapped to frame coordinates and clamped to [1,H]×[1,W].
 * guess_ij NULL ⇒ each video's guess is its previous result, read on the
 * device (no host round trip).  out_ij (n×2), out_raw_ij (n×2, unclamped) and
 * out_resp (n, the maximum response the reference discards at :59) may be NULL.
 * Synchronises the batch's stream before returning when any out_* is given.
 */
PT_API int pt_batch_step(pt_batch *b, const int32_t *guess_ij,
                  int32_t *out_ij, int32_t *out_raw_ij, float *out_resp);

/*
 * The frame loop of track_one (src/PawsomeTracker.jl:163-169, intended lines
 * :166-167) over T time steps with all frames resident in HBM:
 * frame of (step t, video v) = dev_base + t*step_stride + v*frame_stride.
 * ij[t] = trckr(ij[t-1]); the chain never leaves the device.  out_ij is HOST
 * memory, T×n×2 int32; out_resp T×n floats or NULL.  Starts from the guess on
 * the device (pt_batch_set_guess or the previous result).
 */
PT_API int pt_batch_track_device(pt_batch *b, const void *dev_base, size_t step_stride, size_t frame_stride,
                          size_t pitch, int T, int32_t *out_ij, float *out_resp);

/*
 * The same loop with HOST-resident frames (what the reference has after
 * `read!`): frames[t*n + v].  mode 0 = footprint streaming: only the
 * (2r+l)×(2r+l) footprint around each guess is gathered into pinned staging
 * and copied (the reference touches nothing else of the frame, :56-57);
 * mode 1 = whole frames through double-buffered pinned staging
 * (BASELINE.json north_star's streaming layout).  Host→device copies,
 * kernels and the device→host read of every step's result are all inside this
 * call.  Uses the fills already set.  Starts from the guess on the device.
 */
PT_API int pt_batch_track_host(pt_batch *b, const void *const *frames, int T, size_t pitch, int mode,
                        int32_t *out_ij, float *out_resp);

/* Full response map of video v for a given guess, row-major wr×wc floats
 * (wr = 2(ws_rows÷2)+1 …): parity instrumentation only (the reference's
 * `buff` window view, src/PawsomeTracker.jl:58). */
PT_API int pt_batch_response_map(pt_batch *b, int v, int gi, int gj, float *out_map);

/* Asynchronous variants used by the measurement harness: launch T chained
 * steps on `stream` (a cudaStream_t, NULL = the batch's own stream) without
 * synchronising; results stay in the device trajectory buffer until
 * pt_batch_read_track.  A batch runs on ONE stream at a time: passing a different
 * stream first waits for the batch's earlier work, and destroy / pt_batch_read_track /
 * the next call on the batch wait for the caller's stream. */
PT_API int pt_batch_track_device_async(pt_batch *b, const void *dev_base, size_t step_stride,
                                size_t frame_stride, size_t pitch, int T, void *stream);
PT_API int pt_batch_read_track(pt_batch *b, int T, int32_t *out_ij, float *out_resp);

/* Per-handle tuning / debugging knobs (parity tests pin kernel variants with them).  The process-wide defaults
 * come from PT_* environment variables read once at the first pt_batch_create.  Names: "window45" (0: force the
 * generic kernel), "rect45", "rot" (0 off / 1 auto / 2 always / 3 always, handshake forced to the static schedule), "rot_stride" (slots the empty arc of the rotating schedule advances per step, 0 = its own length), "skew", "r45_chunks" (0 = cost model),
 * "generic_target", "mode_slow", "zero_copy", "host_lanes", "cluster" (0 auto / 1 off / 2, 4, 8 CTAs per lone
 * window), "bulk" (1: TMA staging of the cluster kernel, 0: global loads), "wide" (0: 32-column generic kernel only), "two_phase" (64-column kernel: 0 fused, 1 auto, 2 always row pass and column pass as two launches), "cols_teams" (column kernel of the two-phase path: 0 auto, 1 never two teams per CTA), "cols_ch" (its rows per chunk, 0 = cost model), "crop_gather" (page-locked host frames outside the per-window kernels: 1 footprints copied to device crops once per step, 0 read in place).  Unknown name or value out of range: PT_ERR_ARG. */
PT_API int pt_batch_set_option(pt_batch *b, const char *name, int value);

/* Introspection for the harness: kernels launched by this handle so far, and
 * the name of the window kernel variant the current geometry dispatches to. */
PT_API long long pt_batch_launch_count(const pt_batch *b);
PT_API const char *pt_batch_kernel_name(const pt_batch *b);
/* Name of the kernel the batch's most recent step/track launch actually ran ("" before the first one):
 * e.g. a chained pt_batch_track_device over 256 videos runs dog_window45_rot, a single step dog_window45_argmax. */
PT_API const char *pt_batch_last_kernel(const pt_batch *b);
PT_API void *pt_batch_stream(const pt_batch *b); /* the cudaStream_t the batch launches on */

/* ---- single tracker (thin wrapper over a batch of one) -------------------- */
/* Tracker(img, target_width, window_size, darker_target) — src/PawsomeTracker.jl:39-52 */
PT_API int pt_tracker_create(int H, int W, double target_width, int ws_rows, int ws_cols,
                      int darker_target, int pixel, int device, pt_tracker **out);
PT_API void pt_tracker_destroy(pt_tracker *t);
/* read!(vid, trckr.img.data) — src/PawsomeTracker.jl:166 (whole frame to HBM) */
PT_API int pt_tracker_set_frame(pt_tracker *t, const void *frame, size_t pitch);
/* fillvalue = mode(_img) — :47 */
PT_API int pt_tracker_compute_fill(pt_tracker *t, int *fill_out);
PT_API int pt_tracker_set_fill(pt_tracker *t, int fill);
/* trckr(guess) on the frame in HBM — :55-62 */
PT_API int pt_tracker_step(pt_tracker *t, int gi, int gj, int *oi, int *oj, float *resp);
/* trckr(guess) on a HOST frame, copying only the window footprint — :55-62 */
PT_API int pt_tracker_step_host(pt_tracker *t, const void *frame, size_t pitch, int gi, int gj,
                         int *oi, int *oj, float *resp);
PT_API pt_batch *pt_tracker_batch(pt_tracker *t);

/* ---- arbitrary rectangle (full-frame DoG benchmark shape) ----------------- */
/* DoG response over output rows y0..y0+wr-1, cols x0..x0+wc-1 (0-based, may
 * leave the frame) of video v's current frame + argmax; the same arithmetic as
 * pt_batch_step with an explicit rectangle instead of guess ± radii.
 * Outputs are 1-based, oi/oj clamped. */
PT_API int pt_batch_rect_argmax(pt_batch *b, int v, int y0, int x0, int wr, int wc,
                         int *oi, int *oj, int *raw_i, int *raw_j, float *resp);

/* The same rectangle on the current frame of EVERY video of the batch in one launch
 * (batched full-frame DoG; batched auto-detect, src/PawsomeTracker.jl:99-105).
 * out_ij: [n][4] = (i, j clamped, raw i, raw j), 1-based; out_resp: [n].  With
 * no_readback != 0 the call only enqueues the launch on the batch stream (results stay on the
 * device; used to time the kernel alone) and the output pointers are ignored. */
PT_API int pt_batch_rect_argmax_all(pt_batch *b, int y0, int x0, int wr, int wc,
                             int32_t *out_ij, float *out_resp, int no_readback);

/* ---- diagnostics (src/diagnose.jl) --------------------------------------------- */
/* `imresize!(dia.buffer, img)` (src/diagnose.jl:33; buffer 360x640, :2) for the CURRENT frame in HBM of every
 * video: bilinear, pixel-centre aligned, rounded to u8.  out: HOST memory, n*out_h*out_w bytes.  The overlay
 * (label, dot, trail) and the encoder stay on the host. */
PT_API int pt_batch_downscale(pt_batch *b, int out_h, int out_w, uint8_t *out);

/* ---- page-locked host memory for decoders -------------------------------------- */
/* A frame the decoder writes straight into page-locked memory (the destination of
 * `read!(vid, trckr.img.data)`, src/PawsomeTracker.jl:166) can be read by the kernels over PCIe without a
 * staging copy: pt_batch_track_host recognises such frames and runs the whole frame loop as one launch. */
PT_API int pt_host_alloc(size_t bytes, void **out);
PT_API int pt_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* PAWSOME_H */
