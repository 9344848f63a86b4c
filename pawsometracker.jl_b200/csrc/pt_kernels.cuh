// pt_kernels.cuh — device-side argument block and launch prototypes shared by
// the kernels (pt_kernels.cu, pt_window45.cu) and the C-ABI host layer (pt_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pt {

constexpr int kTileCols = 32;    // output columns per CTA strip (generic kernel)
constexpr int kBatchRows = 32;   // footprint rows staged per batch (generic kernel)
constexpr int kTapChunk = 5;     // taps per unrolled chunk; l = 4⌈σ√2⌉+1 ≡ 1 (mod 4), 65 and 245 divide by 5
constexpr int kGenericThreads = 128;

// Per-batch tuning / debugging knobs.  Defaults come from the PT_* environment variables, read ONCE per
// process (defaults_from_env in pt_api.cu); a batch copies them at creation and pt_batch_set_option changes
// them per handle.  Nothing on a per-call path calls getenv.
struct Cfg {
    int sms = 148;            // SM count of the batch's device (cudaDevAttrMultiProcessorCount at create)
    int window45 = 1;         // 0: never use the l = 65 / 45x45 specialised kernels   (PT_DISABLE_WINDOW45)
    int rect45 = 1;           // 0: never use dog_rect45_march                          (PT_DISABLE_RECT45)
    int rot = 1;              // dog_window45_rot: 0 off, 1 where it pays, 2 always     (PT_W45_ROT)
    int rot_stride = 0;       // slots the empty arc advances per step, 0 = its own length (PT_W45_ROT_STRIDE)
    int skew = 1;             // two-window CTAs: 0 free-running, 1 token, 2 lock       (PT_W45_SKEW)
    int r45_chunks = 0;       // chunks per strip of dog_rect45_march, 0 = cost model   (PT_R45_CHUNKS)
    int generic_target = 592; // CTAs the generic kernel aims for                       (PT_GENERIC_TARGET)
    int mode_slow = 0;        // 1: always run the last-position pass of mode           (PT_MODE_SLOW)
    int zero_copy = 1;        // 0: never read pinned host frames in place              (PT_NO_ZEROCOPY)
    int host_lanes = 0;       // host threads of the pageable footprint path, 0 = auto  (PT_HOST_LANES)
    int cluster = 0;          // lone-window cluster kernel: 0 auto, 1 off, 2/4/8 CTAs per window (PT_W45_CLUSTER)
    int wide = 1;             // 1: dog_rect_argmax_wide (64-column strips) where it fits, 0: always the 32-column kernel (PT_GENERIC_WIDE)
    int crop_gather = 1;      // page-locked host frames, geometries outside the per-window kernels: 1 = copy each footprint into a device crop once per step (gather_footprints), 0 = the filter kernels read the host frames in place (PT_CROP_GATHER)
    int cols_ch = 0;          // two-phase wide path: output rows per column-kernel chunk (multiple of 32), 0 = cost model (PT_WIDE_COLS_CH)
    int cols_teams = 0;       // two-phase wide path, column kernel: 0 auto, 1 always one team of 8 warps per CTA (PT_WIDE_COLS_TEAMS)
    int two_phase = 1;        // wide kernel: 0 fused only, 1 auto, 2 always row pass and column pass as two launches (PT_WIDE_TWO_PHASE)
    int smem_optin = 0;       // largest dynamic shared memory per block the device allows (set at create)
    int bulk = 1;             // cluster kernel staging: 1 = one TMA tile copy per step into shared memory, 0 = global loads (PT_W45_BULK)
};

// One launch = one (trckr::Tracker)(guess) evaluation for every window of the
// batch (reference: src/PawsomeTracker.jl:55-62).
struct WinArgs {
    const void *frames;        // frame of window 0 for this step (u8 or f32)
    size_t frame_stride;       // elements between consecutive windows' frames
    int pitch;                 // elements between rows
    int H, W;                  // frame size
    const float *fill;         // [n] border fill in pixel units (u8 scale or [0,1])
    const int2 *guess;         // [n] 1-based (row, col) = (x, y) fields; unused in rect mode
    int rect_mode;             // 1: window origin is (ry0, rx0) for every window
    int ry0, rx0;              // 0-based origin of the output rectangle (rect mode)
    int rr, rc;                // radii = window_size .÷ 2          (:44)
    int wr, wc;                // output rows / cols of the window  (:56)
    int L, w, Lpad;            // kernel length, half width, length padded to kTapChunk
    const float2 *taps_row;    // [Lpad] (narrow, wide) row-pass taps (pixel scale folded in)
    const float2 *taps_col;    // [Lpad] (narrow, wide) column-pass taps (sign folded in)
    int strips, chunks, CH;    // CTA decomposition of one window
    int wide;                  // 0: 32-column strips (dog_rect_argmax_generic); 1: 64-column strips (dog_rect_argmax_wide)
    unsigned long long *keys;  // [n] packed running argmax (zero between launches)
    unsigned int *counters;    // [n] CTA completion counters (zero between launches)
    unsigned int *tickets;     // [2] work-ticket counter + finished-halves counter of dog_rect45_march (zero between launches)
    int4 *out_pos;             // [n] (i, j clamped; raw_i, raw_j), 1-based
    float *out_resp;           // [n] maximum response
    int2 *next_guess;          // [n] clamped result for the next chained step (may alias guess)
    int4 *traj_pos;            // optional [n] trajectory slot of this step
    float *traj_resp;          // optional [n]
    float *map_out;            // optional response maps, [n][wr*wc] row-major (parity instrumentation)
    // chained steps in one launch (specialised kernel only): step t reads
    // frames + t*step_stride and writes trajectory slot t (traj_* then hold [T][n])
    int T;
    size_t step_stride;
    const void *const *frame_ptrs;  // optional DEVICE array [T][n] of frame base pointers (overrides frames/strides):
                               // used for zero-copy reads of pinned host frames (specialised kernel only)
    unsigned int *xflag;       // [n] hand-off flags   } dog_window45_rot (windows hopping between SMs); zero between launches,
    int2 *xpos;                // [n] hand-off guesses } the kernel leaves them zeroed
    const float *h_taps;       // HOST copy of the taps: [L] row narrow, [L] row wide, [L] col narrow, [L] col wide
    int host_frames;           // frames/strides address page-locked HOST memory (zero-copy over PCIe): changes the cluster policy only
    int cols_teams;            // dog_cols_wide: 0 auto, 1 one team of 8 warps per CTA always (option "cols_teams")
    const int2 *crop_org;      // frames are per-window crops (gather_footprints): origin (row, col) of each crop in its real frame, else null
    int Hreal, Wreal;          // … and the real frame size results are clamped to
    float2 *mid;               // two-phase wide path: row-pass intermediate [n][wr + 2w][strips·64] (null = fused kernel)
};

// Orderable packing of (response, column-major index): max key = largest
// response, ties → smallest column-major index = findmax's "first maximum"
// (src/PawsomeTracker.jl:59).
__device__ __forceinline__ unsigned long long pack_key(float v, unsigned int idx)
{
    v += 0.0f; // canonicalise -0.0
    unsigned int b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float key_value(unsigned long long k)
{
    unsigned int b = (unsigned int)(k >> 32);
    b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
    return __uint_as_float(b);
}
__device__ __forceinline__ unsigned int key_index(unsigned long long k)
{
    return 0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull);
}

// Decode the winning key of window v and publish it: absolute index (:60),
// clamp (:61), response, next guess, optional trajectory slot.
__device__ __forceinline__ void publish_result(const WinArgs &a, int v, unsigned long long key,
                                               int wy0, int wx0)
{
    unsigned int idx = key_index(key);
    int xx = (int)(idx / (unsigned int)a.wr);
    int yy = (int)(idx - (unsigned int)xx * (unsigned int)a.wr);
    int raw_i = wy0 + yy + 1, raw_j = wx0 + xx + 1;
    int Hc = a.H, Wc = a.W;
    if (a.crop_org) { const int2 o = a.crop_org[v]; raw_i += o.x; raw_j += o.y; Hc = a.Hreal; Wc = a.Wreal; }   // crop → frame
    int ci = min(max(raw_i, 1), Hc), cj = min(max(raw_j, 1), Wc);
    float resp = key_value(key);
    int4 p = make_int4(ci, cj, raw_i, raw_j);
    a.out_pos[v] = p;
    a.out_resp[v] = resp;
    if (a.next_guess) a.next_guess[v] = make_int2(ci, cj);
    if (a.traj_pos) { a.traj_pos[v] = p; a.traj_resp[v] = resp; }
}

// Per-device one-time setup (cudaFuncSetAttribute is per device): called by pt_batch_create under a lock,
// once per (process, device), with that device current.
cudaError_t generic_init_device();
cudaError_t window45_init_device();

size_t generic_smem_bytes(int L, int Lpad);
cudaError_t launch_generic(const WinArgs &a, int n, int pixel, cudaStream_t s);

// The same filter with 64-column strips and integer-free inner loops (pt_generic64.cu); needs more shared memory per
// CTA (l up to ≈ 285), the 32-column kernel above covers longer kernels and narrow windows.
int wide_max_kernel_len();
size_t wide_smem_bytes(int L);
size_t wide_cols_smem_bytes(int L, int CH);
size_t wide_mid_elems(int L, int wr, int wc, int n);
cudaError_t wide_init_device();
cudaError_t launch_wide(const WinArgs &a, int n, int pixel, cudaStream_t s);

// Specialised batched kernel: l = 65 (target_width 25), 45×45 window.
bool window45_supported(const WinArgs &a, int pixel);
cudaError_t launch_window45(const WinArgs &a, const Cfg &cfg, int n, int pixel, cudaStream_t s);
// name of the kernel launch_window45 would run for this launch (dog_window45_argmax / _rot / _cluster<C>)
const char *window45_kernel_for(const WinArgs &a, const Cfg &cfg, int n, int pixel);
const char *window45_name();
#ifdef PT_PROBES
void window45_set_debug(long long *dev_buf);   // phase-timestamp buffer [n][T][6] (profiling build only)
#endif

// l = 65 rectangles of any size cut into 45x45 tiles evaluated like windows (auto-detect, full frame, …).
bool rect45_supported(const WinArgs &a, const Cfg &cfg, int pixel);
cudaError_t launch_rect45(const WinArgs &a, const Cfg &cfg, int n, int pixel, cudaStream_t s);
const char *rect45_name();

// fillvalue = mode(frame) (src/PawsomeTracker.jl:47) for n frames.
// hist: [n][kModeScratch] unsigned scratch, zero before the first call and left zeroed by every call.
constexpr int kModeScratch = 832;
cudaError_t launch_mode(const void *frames, size_t frame_stride, int pitch, int H, int W, int n,
                        int pixel, unsigned int *hist, float *fill_out, int *fill_int_out,
                        bool force_slow, cudaStream_t s);

// Footprints of page-locked host frames → device crops, once per step (see gather_footprints in pt_kernels.cu).
cudaError_t launch_gather_footprints(const void *frames, size_t frame_stride, int pitch, int H, int W, int n, int pixel,
                                     const int2 *guess, const float *fill, int rr, int rc, int w, int fr, int cp,
                                     void *crops, size_t crop_stride, int2 *org, int2 *cguess, cudaStream_t s);

// imresize!(dia.buffer, img) (src/diagnose.jl:33): current frame of n videos → [n][oh][ow] u8, bilinear, centre-aligned.
cudaError_t launch_downscale(const void *frames, size_t frame_stride, int pitch, int H, int W, int n, int pixel,
                             int oh, int ow, uint8_t *out, cudaStream_t s);

} // namespace pt
