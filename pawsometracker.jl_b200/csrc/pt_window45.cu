// pt_window45.cu — the batched hot path at the reference's default geometry:
// target_width = 25 → l = 65 (w = 32), window 45×45 (radii 22), footprint
// 109×109.  One CTA evaluates one (video, frame) window at a time; a launch can chain
// T time steps of the same videos (frames resident in HBM, or pinned on the host)
// because windows of different videos never interact: ij[t] = trckr(ij[t-1])
// (src/PawsomeTracker.jl:167) only links consecutive frames of ONE video.
//
// Scheduling: one 512-thread CTA per SM hosts TWO independent windows ("halves"), each
// run by 8 warps that synchronise on their own named barrier.  The warps of the two
// halves are interleaved over the scheduler slots (warps 0-3 and 8-11 → half 0, warps 4-7
// and 12-15 → half 1), so each SM sub-partition holds two warps of either window and the
// slot-order arbitration treats both alike.  (With two separate CTAs per SM the warp
// scheduler favoured one of them: 13.3 K vs 19.8 K cycles per frame, and the slow one set
// the launch time.)  Half h of CTA c walks videos c + S·h, c + S·(h+2), … (S = #CTAs) and
// keeps the guess of its current video in registers from frame to frame: no hand-off.
//
// Per window:
//   stage  109×109 pixels → smem as (pixel − fill), 0 outside the frame
//          (the PaddedView border, src/PawsomeTracker.jl:48, after subtracting
//          the constant fill — legal because ΣDoG = 0)
//   row    both Gaussians for 109 rows × 45 columns.  The factors are
//          symmetric, so each output is g0·x0 + Σ_d g_d·(x_-d + x_+d): one FADD
//          feeds ONE packed FFMA2 that advances (narrow, wide) together.
//          Thread = (row, 5 consecutive columns): 31 warp-tasks balance over 8 warps.
//   col    45×45 outputs, thread = (column, 9 consecutive rows); one FFMA2
//          advances two vertically adjacent outputs; subtraction and the
//          darker_target sign are folded into the column taps.
//   argmax warp shuffles + 8 keys in smem that every thread folds itself: first
//          maximum in column-major order (findmax, :59); clamp (:61).
//
// All taps are kernel parameters (constant bank → uniform registers), the loops
// are fully unrolled, so no tap is ever loaded inside the passes.
#include "pt_kernels.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace pt {

namespace {

constexpr int L = 65, HW = 32;          // kernel length / half width
constexpr int WR = 45, WC = 45;         // window outputs
constexpr int FR = WR + 2 * HW;         // 109 footprint rows
constexpr int FC = WC + 2 * HW;         // 109 footprint cols
constexpr int PIN = 109;                // s_in pitch (floats), odd → row-lanes conflict-free
constexpr int PM = 45;                  // s_mid pitch (float2), odd → 64-bit row-lane stores conflict-free
constexpr int R = 9;                    // column pass: outputs per thread along the filter direction
constexpr int NG = 5;                   // groups of R per 45
constexpr int COL_ITEMS = WC * NG;      // 225 → 8 warp-tasks on 8 warps
constexpr int RR = 5;                   // row pass: outputs per thread; 9 groups per row
// row pass items: FR * (WC / RR) = 981 → 31 warp-tasks: 4 (or 3) per warp, 98.9 % lane fill
#ifndef PT_W45_THREADS
#define PT_W45_THREADS 256
#endif
constexpr int THREADS = PT_W45_THREADS;  // threads per window: 8 warps = 2 per SM sub-partition
constexpr int NWARPS = THREADS / 32;
constexpr int CTA_THREADS = 2 * THREADS;   // two windows per CTA
constexpr size_t HALF_SMEM = ((size_t)(FR * PIN + 1) * sizeof(float) + (size_t)FR * PM * sizeof(float2) + 15) & ~(size_t)15;

// Taps as kernel parameters, laid out for packed FP32 (fma.rn.f32x2 → FFMA2):
//   rt[d]  = (narrow, wide) folded row taps, d = |k − 32| (pixel scale folded in)
//   cpp[q] = (cp[q], cp[q−1]), cmq[q] = (cm[q], cm[q−1]) column taps (sign folded in),
//            zero outside 0..64: one FFMA2 advances two vertically adjacent outputs.
struct Taps45 {
    float2 rt[HW + 1];
    float2 cpp[L + 1], cmq[L + 1];
};

// PRMT selectors of the u8→f32 conversion (byte k of the word under the 2^23 exponent pattern).  Passed as a kernel
// parameter: as constant-bank operands they cost no instruction, whereas immediates made ptxas keep the 2^23 pattern
// as the immediate and re-materialise the selectors with MOVs (one per PRMT).
__constant__ unsigned int c_prmt_sel[4] = {0x7540u, 0x7541u, 0x7542u, 0x7543u};

struct Args45 {
    const void *frames;                 // frame of window 0 at step 0
    size_t frame_stride, step_stride;   // elements
    const void *const *frame_ptrs;      // optional [T][n] frame pointers (zero-copy pinned host frames)
    int pitch, H, W;
    const float *fill;
    const int2 *guess;                  // [n] start guess (1-based)
    int T, n;
    int4 *out_pos; float *out_resp;     // [n] last step
    int2 *next_guess;                   // [n] or null
    int4 *traj_pos; float *traj_resp;   // [T][n] or null
    // geometry inside the 45×45 / l = 65 frame of the kernels (any window ≤ 45×45 and kernel length ≤ 65: the taps
    // are zero-padded to 65, outputs beyond wr × wc are masked): radii, output rows / columns, the footprint rows
    // [f_lo, f_lo + nfr) whose taps are not all zero, groups of 5 output columns, reciprocal of nfr (item → group by
    // one multiplication: (item · inv_nfr) >> 18), lane stride of a row group in the column pass
    int rr, rc, wr, wc;
    int f_lo, nfr, ng, cs;
    unsigned int inv_nfr;
    int host_frames;                    // frames are page-locked host memory (policy only)
    int rstride;                        // dog_window45_rot: slots the empty arc advances per step (0 = its own length)
    int skew;                           // 1: alternate the row passes of the two windows of a CTA (token); 2: lock
    int tm_rows_step, tm_rows_frame;    // dog_window45_cluster, TMA tensor mode: rows of the 2-D frame tensor per step / per video
    unsigned int *xflag;                // [n] hand-off flags of dog_window45_rot (zero between launches)
    int2 *xpos;                         // [n] hand-off guesses
#ifdef PT_PROBES
    long long *dbg;                     // optional [n][T][6]: smid|globaltimer, clock64 at start / stage / row / col / end
#endif
};

// Phase probes exist only in the profiling build (-DPT_PROBES → libpawsome_cuda_probes.so, tools/phase_timing.py);
// the product kernels carry none.
#ifdef PT_PROBES
#define PT_PROBE_BEGIN(a, v, t, tid)                                                              \
    long long *dbg = (a).dbg ? (a).dbg + ((size_t)(v) * (a).T + (t)) * 6 : nullptr;                \
    if (dbg && (tid) == 0) {                                                                      \
        unsigned int smid_;                                                                       \
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));                                        \
        unsigned long long gt_;                                                                   \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                   \
        dbg[0] = ((long long)gt_ << 8) | (long long)(smid_ & 0xFF);                               \
        dbg[1] = clock64();                                                                       \
    }
#define PT_PROBE(k, tid) do { if (dbg && (tid) == 0) dbg[k] = clock64(); } while (0)
#else
#define PT_PROBE_BEGIN(a, v, t, tid)
#define PT_PROBE(k, tid) do { } while (0)
#endif

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
// max over the warp of 64-bit keys with two 32-bit REDUX (instead of five rounds of 64-bit shuffles)
__device__ __forceinline__ unsigned long long warp_max_key(unsigned long long key)
{
    const unsigned int hi = (unsigned int)(key >> 32), lo = (unsigned int)key;
    const unsigned int mh = __reduce_max_sync(0xFFFFFFFFu, hi);
    const unsigned int ml = __reduce_max_sync(0xFFFFFFFFu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | (unsigned long long)ml;
}

__device__ __forceinline__ float2 fmul2(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}

// ---- staging: footprint → smem as (pixel − fill), 0 outside the frame ----------------
// Warp w takes rows w, w+9, …; every load of a thread is issued before the first
// conversion so a frame that is cold in L2/HBM costs one memory round trip.
// NROWS = rows staged into s_in rows 0..NROWS-1 (109 for a whole footprint, 45 for the next batch of a marching strip).
// f32 frames: lane = column (+32q), coalesced 4-byte loads.
// kContig: warp w takes the contiguous rows [w·RPW, (w+1)·RPW) instead of w, w+8, … (the cluster kernel row-filters
// the rows a warp staged without a CTA barrier in between).
// Only rows [f_lo, f_hi) are staged (the per-window kernels skip footprint rows whose taps are all zero).
template <int NROWS, int FCOLS = FC, int PITCH = PIN, bool kContig = false>
__device__ __forceinline__ void stage_rows(const float *frame, int pitch, int H, int W, int fy0, int fx0,
                                           float fill, float *s_in, int warp, int lane, int f_lo = 0, int f_hi = NROWS)
{
    constexpr int RPW = (NROWS + NWARPS - 1) / NWARPS;   // 14 rows per warp for a footprint (last ones masked)
    constexpr int NQ = (FCOLS + 31) / 32;
    float px[RPW][NQ];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int f = kContig ? warp * RPW + r : warp + r * NWARPS;
        const int Y = fy0 + f;
        const bool yok = (f < NROWS) && (f >= f_lo) && (f < f_hi) && (Y >= 0) && (Y < H);
        const float *rowp = frame + (size_t)(yok ? Y : 0) * pitch;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = lane + 32 * q;
            const int X = fx0 + c;
            px[r][q] = (yok && c < FCOLS && X >= 0 && X < W) ? __ldg(rowp + X) : fill;
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int f = kContig ? warp * RPW + r : warp + r * NWARPS;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = lane + 32 * q;
            if (f < NROWS && f >= f_lo && f < f_hi && c < FCOLS) s_in[f * PITCH + c] = px[r][q] - fill;
        }
    }
}

// u8 frames: one aligned 32-bit load = 4 pixels per lane (28 lanes cover a row).  Bytes
// outside the frame are replaced by the fill byte, so they convert to exactly 0.  u8→f32
// uses the 2^23 trick: PRMT builds the bits of (8388608 + px) on the ALU pipe and the
// FADD that subtracts the fill finishes the conversion: (8388608+px) − (8388608+fill) is
// exact.  Each lane rotates its word by lane/8 bytes so the four stores of a warp hit 32
// distinct banks.  Requires frame base, pitch and strides to be multiples of 4 bytes and
// pitch ≥ round_up(W, 4) (checked by window45_supported).
template <int NROWS, bool kInterior, int FCOLS = FC, int PITCH = PIN, bool kContig = false>
__device__ __forceinline__ void stage_rows_u8(const uint8_t *frame, int pitch, int H, int W, int fy0, int fx0,
                                              float fill, float *s_in, int warp, int lane, int f_lo = 0, int f_hi = NROWS)
{
    constexpr int RPW = (NROWS + NWARPS - 1) / NWARPS;
    const int xa = fx0 & ~3, phase = fx0 - xa;             // aligned start, phase 0..3
    const int X = xa + 4 * lane;                            // frame column of this lane's word
    const unsigned int fillw = (unsigned int)fill * 0x01010101u;
    unsigned int keep = 0xFFFFFFFFu;                        // bytes of the word inside [0, W)
    if (!kInterior) {
        keep = 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) if (X + b >= 0 && X + b < W) keep |= 0xFFu << (8 * b);
    }
    const bool wordok = keep != 0u && (4 * lane - phase < FCOLS);
    const int rot = lane >> 3;
    int col[4];
    bool cok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        col[k] = 4 * lane - phase + ((k + rot) & 3);
        cok[k] = col[k] >= 0 && col[k] < FCOLS;
    }
    unsigned int wd[RPW];
    // one running row pointer (bumped by NWARPS rows per load): two integer instructions per load instead of a
    // fresh 64-bit row-times-pitch product
    const uint8_t *rowp = frame + ((long long)(fy0 + (kContig ? warp * RPW : warp)) * pitch + X);   // only dereferenced when valid
    const unsigned long long rstep = (unsigned long long)((kContig ? 1 : NWARPS) * pitch);
    unsigned long long addr = reinterpret_cast<unsigned long long>(rowp);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int f = kContig ? warp * RPW + r : warp + r * NWARPS;
        const int Y = fy0 + f;
        const bool ok = wordok && (f < NROWS) && (f >= f_lo) && (f < f_hi) && (kInterior || ((Y >= 0) && (Y < H)));
        wd[r] = fillw;
        // (an L2::64B prefetch-size hint on this load does not change the PCIe traffic of zero-copy host frames: measured)
        if (ok) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(wd[r]) : "l"(addr));
        addr += rstep;                                       // one 64-bit add per load
    }
    const float cst = 8388608.0f + fill;
    const unsigned int magic = 0x4B000000u;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int f = kContig ? warp * RPW + r : warp + r * NWARPS;
        if (f < NROWS && f >= f_lo && f < f_hi) {
            unsigned int w = kInterior ? wd[r] : ((wd[r] & keep) | (fillw & ~keep));
            w = __funnelshift_r(w, w, 8 * rot);            // byte k of w = pixel (k + rot) & 3 of the word
            float *dst = s_in + f * PITCH;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float val = __uint_as_float(__byte_perm(w, magic, c_prmt_sel[k])) - cst;
                if (cok[k]) dst[col[k]] = val;
            }
        }
    }
}

template <int NROWS, int FCOLS = FC, int PITCH = PIN, bool kContig = false>
__device__ __forceinline__ void stage_rows(const uint8_t *frame, int pitch, int H, int W, int fy0, int fx0,
                                           float fill, float *s_in, int warp, int lane, int f_lo = 0, int f_hi = NROWS)
{
    // interior: every aligned word the rows touch lies inside the frame → no byte masks, no row checks
    constexpr int NWORDS = (FCOLS + 3 + 3) / 4;               // aligned words covering FCOLS columns at any phase (28 for 109)
    const bool interior = (fy0 + f_lo >= 0) && (fy0 + f_hi <= H) && ((fx0 & ~3) >= 0) && ((fx0 & ~3) + 4 * NWORDS <= W);
    if (interior) stage_rows_u8<NROWS, true, FCOLS, PITCH, kContig>(frame, pitch, H, W, fy0, fx0, fill, s_in, warp, lane, f_lo, f_hi);
    else stage_rows_u8<NROWS, false, FCOLS, PITCH, kContig>(frame, pitch, H, W, fy0, fx0, fill, s_in, warp, lane, f_lo, f_hi);
}

template <typename PixT>
__device__ __forceinline__ void stage_tile(const PixT *frame, int pitch, int H, int W, int fy0, int fx0,
                                           float fill, float *s_in, int warp, int lane, int f_lo, int f_hi)
{
    stage_rows<FR>(frame, pitch, H, W, fy0, fx0, fill, s_in, warp, lane, f_lo, f_hi);
}

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned int mbar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned int mbar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned int mbar, unsigned int parity)
{
    unsigned int ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok != 0u;
}
// Wait for the phase with the given parity; a watchdog turns a lost copy into a launch error instead of a hang.
__device__ __forceinline__ void mbar_wait(unsigned int mbar, unsigned int parity)
{
    if (mbar_try_wait(mbar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(mbar, parity)) {
        if (clock64() - t0 > (1ll << 32)) __trap();          // ≈ 2 s at 2 GHz
    }
}
// u8 → f32 conversion of footprint rows out of raw u8 rows in shared memory (filled by TMA): the same word / PRMT /
// 2^23 scheme as stage_rows_u8, LDS instead of LDG.  Warp w converts the contiguous rows [14w, 14w + 14).  raw row 0 ↔
// frame row raw_y0, raw byte 0 ↔ frame column raw_xa (a multiple of 16), SPAN bytes per raw row.  Bytes outside the
// frame were never copied (or arrived as zeros): they are replaced by the fill byte.
template <int SPAN, int FCOLS, int PITCH, bool kInterior>
__device__ __forceinline__ void convert_rows_u8(const uint8_t *raw, int raw_y0, int raw_xa, int H, int W, int fy0, int fx0,
                                                float fill, float *s_in, int warp, int lane)
{
    constexpr int RPW = (FR + NWARPS - 1) / NWARPS;
    const int xw0 = fx0 & ~3, phase = fx0 - xw0;
    const int X = xw0 + 4 * lane;
    const unsigned int fillw = (unsigned int)fill * 0x01010101u;
    unsigned int keep = 0xFFFFFFFFu;
    if (!kInterior) {
        keep = 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) if (X + b >= 0 && X + b < W) keep |= 0xFFu << (8 * b);
    }
    const bool wordok = keep != 0u && (4 * lane - phase < FCOLS);
    const int rot = lane >> 3;
    int col[4];
    bool cok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        col[k] = 4 * lane - phase + ((k + rot) & 3);
        cok[k] = col[k] >= 0 && col[k] < FCOLS;
    }
    unsigned int wd[RPW];
    const int f0 = warp * RPW;
    const uint8_t *src = raw + (fy0 + f0 - raw_y0) * SPAN + (X - raw_xa);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int f = f0 + r;
        const int Y = fy0 + f;
        const bool ok = wordok && (f < FR) && (kInterior || ((Y >= 0) && (Y < H)));
        wd[r] = fillw;
        if (ok) wd[r] = *reinterpret_cast<const unsigned int *>(src + r * SPAN);
    }
    const float cst = 8388608.0f + fill;
    const unsigned int magic = 0x4B000000u;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int f = f0 + r;
        if (f < FR) {
            unsigned int w = kInterior ? wd[r] : ((wd[r] & keep) | (fillw & ~keep));
            w = __funnelshift_r(w, w, 8 * rot);
            float *dst = s_in + f * PITCH;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float val = __uint_as_float(__byte_perm(w, magic, c_prmt_sel[k])) - cst;
                if (cok[k]) dst[col[k]] = val;
            }
        }
    }
}

} // namespace


// ---- row pass over one staged 109×109 tile: item = (footprint row f, group gq of 5 output columns);
// lanes walk rows.  One FADD (symmetric fold) feeds one packed FFMA2 advancing (narrow, wide).
// NROWS rows of s_in (from row 0) → NROWS rows of s_mid (from the pointer given).
// kSkip (kernels shorter than 65): the taps beyond the kernel's half width w are zero — the unrolled tap loop leaves at
// d = 9, 17 or 25 when nothing but zeros follows (one uniform branch per 8 taps).
template <bool kSkip = false>
__device__ __forceinline__ void row_item45(const float *row, float2 *dst, const Taps45 &tp, int w = HW)
{
    float x[RR + 2 * HW];
#pragma unroll
    for (int i = 0; i < RR + 2 * HW; ++i) x[i] = row[i];
    float2 acc[RR];                                  // (narrow, wide) per output
#pragma unroll
    for (int j = 0; j < RR; ++j) acc[j] = fmul2(make_float2(x[j + HW], x[j + HW]), tp.rt[0]);
#pragma unroll
    for (int d = 1; d <= HW; ++d) {
        if (kSkip && (d & 7) == 1 && d > 1 && d > w) break;
#pragma unroll
        for (int j = 0; j < RR; ++j) {
            const float s = x[j + HW - d] + x[j + HW + d];   // exact for u8 frames (integers)
            acc[j] = ffma2(make_float2(s, s), tp.rt[d], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < RR; ++j) dst[j] = acc[j];
}

template <int NROWS = FR>
__device__ __forceinline__ void row_pass45(const float *s_in, float2 *s_mid, int tid, const Taps45 &tp)
{
#pragma unroll 1
    for (int item = tid; item < NROWS * (WC / RR); item += THREADS) {
        const int gq = item / NROWS, f = item - gq * NROWS;
        row_item45(s_in + f * PIN + gq * RR, s_mid + f * PM + gq * RR, tp);
    }
}

// The same over the footprint rows [f_lo, f_lo + nfr) and the first ng column groups only (windows smaller than
// 45×45 / kernels shorter than 65: the other rows meet zero taps only, the other columns are masked).
template <bool kSkip>
__device__ __forceinline__ void row_pass45_rt(const float *s_in, float2 *s_mid, int tid, const Taps45 &tp,
                                              int f_lo, int nfr, int ng, unsigned int inv_nfr)
{
    const int nitems = nfr * ng;
#pragma unroll 1
    for (int item = tid; item < nitems; item += THREADS) {
        const int gq = (int)(((unsigned int)item * inv_nfr) >> 18), f = f_lo + item - gq * nfr;
        row_item45<kSkip>(s_in + f * PIN + gq * RR, s_mid + f * PM + gq * RR, tp, HW - f_lo);
    }
}

// ---- column pass + per-thread argmax over the 45×45 outputs of a tile: item = (column xq, row group h);
// lanes walk columns.  Outputs (2p, 2p+1) share packed accumulators; output 8 stays scalar; narrow and wide
// parts accumulate separately (10 independent dependency chains).  The tile sits at (gy0, gx0) inside an
// output rectangle of wr_tot × wc_tot: outputs beyond it are masked, the key carries the rectangle's
// column-major index.  map_out (optional) receives the responses, row-major with pitch wc_tot.
// kSkip (kernels shorter than 65, i_lo = 32 − w): intermediate rows i < i_lo and i > 72 − i_lo of a thread's column meet
// zero taps only — they are skipped in chunks of 8 (one uniform branch per chunk).
template <bool kSkip = false>
__device__ __forceinline__ unsigned long long col_pass45(const float2 *s_mid, int tid, const Taps45 &tp,
                                                         int gy0, int gx0, int wr_tot, int wc_tot, float *map_out, int cs = 48,
                                                         int i_lo = 0)
{
    // items are laid out cs = 48 (32, 16 for narrow windows) per row group: a half-warp — the unit of a 64-bit
    // shared-memory access — never straddles two row groups, whose addresses differ by an odd multiple of the pitch
    // (30 % excess wavefronts before); lanes and row groups beyond the rectangle leave at once
    const int h = cs == 48 ? tid / 48 : (tid >> (cs == 32 ? 5 : 4)), xq = tid - h * cs;
    if (h >= NG || xq >= WC || gx0 + xq >= wc_tot || gy0 + h * R >= wr_tot) return 0ull;
    const float2 *col = s_mid + (h * R) * PM + xq;
    float2 accP[R / 2], accM[R / 2];
    float acc8p = 0.f, acc8m = 0.f;
#pragma unroll
    for (int p = 0; p < R / 2; ++p) { accP[p] = make_float2(0.f, 0.f); accM[p] = make_float2(0.f, 0.f); }
    const int i_hi = R + 2 * HW - 1 - i_lo;           // last row that meets a non-zero tap (row i meets taps i − 8 … i)
#pragma unroll
    for (int c8 = 0; c8 < (R + 2 * HW + 7) / 8; ++c8) {
        if (kSkip && (8 * c8 + 8 <= i_lo || 8 * c8 > i_hi)) continue;
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
            const int i = 8 * c8 + ii;
            if (i < R + 2 * HW) {
                const float2 m = col[i * PM];
#pragma unroll
                for (int p = 0; p < R / 2; ++p) {
                    const int q = i - 2 * p;                 // tap of the even output; the odd one uses q-1
                    if (q >= 0 && q <= L) accP[p] = ffma2(make_float2(m.x, m.x), tp.cpp[q], accP[p]);
                }
#pragma unroll
                for (int p = 0; p < R / 2; ++p) {
                    const int q = i - 2 * p;
                    if (q >= 0 && q <= L) accM[p] = ffma2(make_float2(m.y, m.y), tp.cmq[q], accM[p]);
                }
                const int q8 = i - (R - 1);
                if (q8 >= 0 && q8 < L) {
                    acc8p = fmaf(m.x, tp.cpp[q8].x, acc8p);
                    acc8m = fmaf(m.y, tp.cmq[q8].x, acc8m);
                }
            }
        }
    }
    float acc[R];
#pragma unroll
    for (int p = 0; p < R / 2; ++p) { acc[2 * p] = accP[p].x + accM[p].x; acc[2 * p + 1] = accP[p].y + accM[p].y; }
    acc[R - 1] = acc8p + acc8m;
    const int gx = gx0 + xq, gyb = gy0 + h * R;
    float bv = acc[0] + 0.0f;
    int bj = 0;
#pragma unroll
    for (int j = 1; j < R; ++j) {
        const float val = acc[j] + 0.0f;
        if (gyb + j < wr_tot && val > bv) { bv = val; bj = j; }   // strict: first maximum within the column segment
    }
    if (map_out) {
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (gyb + j < wr_tot) map_out[(size_t)(gyb + j) * wc_tot + gx] = acc[j] + 0.0f;
    }
    return pack_key(bv, (unsigned int)(gx * wr_tot + gyb + bj));
}

__device__ __forceinline__ void bar_half(int half)
{
    asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "n"(THREADS) : "memory");
}

// kGeom 0: the default geometry (l = 65, 45×45 window) with every bound a compile-time constant; 1: any shorter kernel /
// smaller window, bounds from Args45 (see there); 2: the same for kernels of half width ≤ 24, with the tap loops leaving
// in chunks of 8 where only zero taps follow (tw ≤ 18: 7-16 % faster; with longer kernels nothing is skipped and the
// exits cost 4-6 %, hence a separate instantiation).
template <typename PixT, int kGeom>
__global__ void __launch_bounds__(CTA_THREADS, 1)
dog_window45_argmax(const __grid_constant__ Args45 a, const __grid_constant__ Taps45 tp)
{
    constexpr bool kFull = kGeom == 0, kSkip = kGeom == 2;
    const int g_rr = kFull ? WR / 2 : a.rr, g_rc = kFull ? WC / 2 : a.rc, g_wr = kFull ? WR : a.wr, g_wc = kFull ? WC : a.wc;
    const int g_flo = kFull ? 0 : a.f_lo, g_nfr = kFull ? FR : a.nfr, g_cs = kFull ? 48 : a.cs;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_keys[2][2 * NWARPS];

    // physical warp → (half, logical warp): warps 0-3, 8-11 → half 0; 4-7, 12-15 → half 1
    const int pw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = (pw >> 2) & 1;
    // logical warp 0..7; half 1 is rotated by two sub-partitions so that the warps carrying the
    // third row-pass task (logical 0 and 1) of the two windows sit on different FMA pipes
    const int warp = (((pw & 3) + 2 * half) & 3) + 4 * (pw >> 3);
    const int tid = warp * 32 + lane;
    float *s_in = reinterpret_cast<float *>(smem_raw + half * HALF_SMEM);      // [FR][PIN]
    float2 *s_mid = reinterpret_cast<float2 *>(s_in + FR * PIN + 1);           // [FR][PM] (8-byte aligned: FR*PIN+1 is even)
    unsigned long long *s_key = s_keys[half];
    const int stride = 2 * (int)gridDim.x;

    // Staggering the two windows: left alone the halves drift into lockstep (measured) — both stage at
    // once, leaving the FMA pipes idle, then both run their passes at once.  The row passes are therefore
    // strictly alternated with a token (named barriers 3 and 4 used as semaphores: the 256 threads of one
    // half arrive, the 256 of the other wait): A.row(r) → B.row(r) → A.row(r+1) → …  Each window's
    // stage / reduce / column pass then overlaps the other window's row pass.  NA / NB = number of row
    // passes each half will run in this launch, so the alternation stops cleanly when one half is done.
    const int cntA = ((int)blockIdx.x < a.n) ? (a.n - 1 - (int)blockIdx.x) / stride + 1 : 0;
    const int firstB = (int)blockIdx.x + (int)gridDim.x;
    const int cntB = (firstB < a.n) ? (a.n - 1 - firstB) / stride + 1 : 0;
    const int NA = cntA * a.T, NB = cntB * a.T;
    const bool tokens = a.skew == 1 && NB > 0;
    const bool locked = a.skew == 2 && NB > 0;     // row passes mutually exclusive through a lock in shared memory
    __shared__ int s_rowlock;
    if (threadIdx.x == 0) s_rowlock = 0;
    // reduced geometry: footprint rows outside [f_lo, f_lo + nfr) are never written; the column pass still reads
    // them (against zero taps), so they must hold finite values
    if (!kFull && a.nfr < FR) for (int i = tid; i < (int)(HALF_SMEM / 16); i += THREADS) reinterpret_cast<float4 *>(s_in)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    int round = 0;

  for (int v = (int)blockIdx.x + (int)gridDim.x * half; v < a.n; v += stride) {
    const float fill = a.fill[v];
    int2 g = a.guess[v];

    for (int t = 0; t < a.T; ++t) {
        const unsigned int it = (unsigned int)t;
        PT_PROBE_BEGIN(a, v, t, tid)
        const PixT *frame = a.frame_ptrs
            ? reinterpret_cast<const PixT *>(a.frame_ptrs[(size_t)t * a.n + v])
            : reinterpret_cast<const PixT *>(a.frames) + (size_t)t * a.step_stride + (size_t)v * a.frame_stride;
        const int wy0 = g.x - 1 - g_rr, wx0 = g.y - 1 - g_rc;           // window origin, 0-based
        const int fy0 = wy0 - HW, fx0 = wx0 - HW;                        // footprint origin (of the l = 65 frame)

        // (Fetching the footprint rows of zero-copy HOST frames with one cp.async.bulk per row instead of these lane
        // loads was measured and dropped: NVML counted 11.4 MB of PCIe reads per 256-window step instead of 5.8 MB,
        // the link saturated at 61 GB/s and the step took 147 µs instead of 111 µs.)
        stage_tile<PixT>(frame, a.pitch, a.H, a.W, fy0, fx0, fill, s_in, warp, lane, g_flo, g_flo + g_nfr);

        // ---- warm L2 with everything the NEXT step can touch: its window centre is inside this
        // step's window, so its footprint lies within ±22 px of this one (153 rows × ≤3 lines).
        if (t + 1 < a.T && !a.frame_ptrs) {                      // (host frames are not cached in L2)
            const PixT *nframe = frame + a.step_stride;
            constexpr int PR = FR + WR - 1;                      // 153 rows
            constexpr int NLMAX = (int)(((FC + WC) * sizeof(PixT) + 127) / 128) + 1;   // ≤ 3 lines per row for u8
            const int py0 = fy0 - WR / 2, pxb = (fx0 - WC / 2) * (int)sizeof(PixT);
            const int line0 = pxb >> 7;
            const int nl = ((pxb + (FC + WC - 1) * (int)sizeof(PixT) - 1) >> 7) - line0 + 1;   // exact: 2 or 3 for u8
            const int rowbytes = a.W * (int)sizeof(PixT);
            static_assert(PR <= THREADS, "one thread per prefetched row");
            const int Y = py0 + tid;                              // thread = row: its 2-3 lines share one address
            if (tid < PR && Y >= 0 && Y < a.H) {
                const char *ptr = reinterpret_cast<const char *>(nframe + (size_t)Y * a.pitch) + (line0 << 7);
#pragma unroll
                for (int ln = 0; ln < NLMAX; ++ln) {
                    const int off = (line0 + ln) << 7;
                    if (ln < nl && off >= 0 && off < rowbytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + (ln << 7)));
                }
            }
        }
        if (locked && tid == 0) {
            while (atomicCAS(&s_rowlock, 0, 1) != 0) __nanosleep(40);
        }
        bar_half(half);
        PT_PROBE(2, tid);
        if (tokens) {                                          // wait for the row-pass token
            if (half == 0) { if (round >= 1 && round - 1 < NB) asm volatile("bar.sync 4, %0;" ::"n"(CTA_THREADS) : "memory"); }
            else           { if (round < NA) asm volatile("bar.sync 3, %0;" ::"n"(CTA_THREADS) : "memory"); }
        }

        if (kFull) row_pass45<FR>(s_in, s_mid, tid, tp);
        else row_pass45_rt<kSkip>(s_in, s_mid, tid, tp, a.f_lo, a.nfr, a.ng, a.inv_nfr);
        bar_half(half);
        if (locked && tid == 0) atomicExch(&s_rowlock, 0);
        if (tokens) {                                          // hand the row-pass token to the other window
            if (half == 0) { if (round < NB) asm volatile("bar.arrive 3, %0;" ::"n"(CTA_THREADS) : "memory"); }
            else           { if (round + 1 < NA) asm volatile("bar.arrive 4, %0;" ::"n"(CTA_THREADS) : "memory"); }
        }
        ++round;
        PT_PROBE(3, tid);

        const unsigned long long key = warp_max_key(col_pass45<kSkip>(s_mid, tid, tp, 0, 0, g_wr, g_wc, nullptr, g_cs, g_flo));
        if (lane == 0) s_key[(it & 1) * NWARPS + warp] = key;
        bar_half(half);
        PT_PROBE(4, tid);
        {
            // every warp folds the 8 warp keys itself (lane i reads key i mod 8, two REDUX): no serial section;
            // s_key is double-buffered by iteration parity
            const unsigned long long k = warp_max_key(s_key[(it & 1) * NWARPS + (lane & (NWARPS - 1))]);
            const unsigned int idx = key_index(k);
            const int xx = (int)(idx / (unsigned int)g_wr), yy = (int)(idx - xx * g_wr);
            const int raw_i = wy0 + yy + 1, raw_j = wx0 + xx + 1;                   // absolute index (:60)
            const int ci = min(max(raw_i, 1), a.H), cj = min(max(raw_j, 1), a.W);   // clamp (:61)
            if (tid == 0) {
                const float resp = key_value(k);
                const int4 p = make_int4(ci, cj, raw_i, raw_j);
                if (a.traj_pos) { a.traj_pos[(size_t)t * a.n + v] = p; a.traj_resp[(size_t)t * a.n + v] = resp; }
                if (t == a.T - 1) {
                    a.out_pos[v] = p; a.out_resp[v] = resp;
                    if (a.next_guess) a.next_guess[v] = make_int2(ci, cj);
                }
                PT_PROBE(5, 0);
            }
            g = make_int2(ci, cj);
        }
    }
    bar_half(half);    // s_key / smem of this half are reused by the next video
  }
}


// ---------------------------------------------------------------------------------------------------
// dog_window45_rot — the same per-window work as dog_window45_argmax for batches of S < n < 2·S windows
// (S = #SMs; BASELINE config 3: 256 windows on 148 SMs).  There a static split gives 2S − n SMs one window
// (11.5 K cycles per frame) and the other SMs two (18.6 K each, which sets the launch time) while the
// single-window SMs idle 40 % of the time.  Here the "holes" ROTATE: the 2S window slots (slot σ = SM σ mod S,
// half σ div S) form a ring that holds the n windows in a fixed cyclic order, and the arc of nh = 2S − n empty
// slots advances by its own length every time step, i.e. the nh windows just ahead of the arc hop back over it
// (closed form: rot_window()).  Every window therefore has its SM to itself for nh/n of its steps and all
// windows finish together instead of the lone ones early.  A hop is a hand-off through global memory (guess +
// release flag; the receiving slot is empty and polls), nh of n windows per step; the L2 prefetch issued by the
// old SM serves the new one (L2 is shared).  The row passes of the two halves of an SM exclude each other
// through a lock in shared memory (same effect as the token of dog_window45_argmax, but a half never waits for
// a partner that is empty or late).  Hopping needs all CTAs co-resident.  Instead of a cooperative launch (which
// serialises against every other kernel on the device and cannot share it) the kernel is launched plainly and the
// CTAs agree on the schedule themselves: every CTA counts itself in (xsync[0]); the CTA that completes the count
// proposes "rotate" (all CTAs are resident and stay so until they have run their steps), a CTA that has waited ≈ 30 µs
// without seeing a decision proposes "static"; the first proposal wins (one atomicCAS on xsync[1]) and every CTA
// follows it.  "Static" is the same schedule with the empty slots standing still (nh = 0 hops: window σ stays in
// slot σ), i.e. the split of dog_window45_argmax, which needs no co-residency.  The last CTA to leave zeroes the
// three words.  (Measured: same launch time as the cooperative launch, 184 µs per 20 steps.)
// ---------------------------------------------------------------------------------------------------
// window held by slot `slot` of a ring of R slots at step t (m windows, nh = R − m empty slots), or −1
__device__ __forceinline__ int rot_window(int slot, int t, int R, int m, int nh)
{
    const int a0 = (nh * t) % R;
    int k = slot - a0;
    if (k < 0) k += R;
    if (k >= m) return -1;                   // empty slot at this step
    int u = (nh * t) % m + k;
    if (u >= m) u -= m;
    return u;
}
// the same with the two remainders (nh·t) mod R and (nh·t) mod m carried along by the caller (no division per step)
__device__ __forceinline__ int rot_window_rem(int slot, int remR, int remM, int R, int m)
{
    int k = slot - remR;
    if (k < 0) k += R;
    if (k >= m) return -1;
    int u = remM + k;
    if (u >= m) u -= m;
    return u;
}

template <typename PixT, int kGeom>
__global__ void __launch_bounds__(CTA_THREADS, 1)
dog_window45_rot(const __grid_constant__ Args45 a, const __grid_constant__ Taps45 tp)
{
    constexpr bool kFull = kGeom == 0, kSkip = kGeom == 2;
    const int g_rr = kFull ? WR / 2 : a.rr, g_rc = kFull ? WC / 2 : a.rc, g_wr = kFull ? WR : a.wr, g_wc = kFull ? WC : a.wc;
    const int g_flo = kFull ? 0 : a.f_lo, g_nfr = kFull ? FR : a.nfr, g_cs = kFull ? 48 : a.cs;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_keys[2][2 * NWARPS];
    __shared__ int s_rowlock;
    __shared__ int2 s_handoff[2];

    const int pw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = (pw >> 2) & 1;
    const int warp = (((pw & 3) + 2 * half) & 3) + 4 * (pw >> 3);
    const int tid = warp * 32 + lane;
    float *s_in = reinterpret_cast<float *>(smem_raw + half * HALF_SMEM);
    float2 *s_mid = reinterpret_cast<float2 *>(s_in + FR * PIN + 1);
    unsigned long long *s_key = s_keys[half];
    const int R = 2 * (int)gridDim.x, m = a.n, slot = (int)blockIdx.x + (int)gridDim.x * half;
    __shared__ unsigned int s_mode;
    if (threadIdx.x == 0) {
        s_rowlock = 0;
        unsigned int *xs = a.xflag + a.n;                  // [0] arrived, [1] decision (1 rotate, 2 static), [2] left
        unsigned int mode;
        if (atomicAdd(xs, 1u) + 1u == gridDim.x) atomicCAS(xs + 1, 0u, 1u);
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(mode) : "l"(xs + 1) : "memory");
            if (mode) break;
            if (a.skew < 0 || clock64() - t0 > (1ll << 16)) atomicCAS(xs + 1, 0u, 2u);   // (skew < 0: option rot = 3, tests)
        }
        s_mode = mode;
    }
    if (!kFull && a.nfr < FR) for (int i = tid; i < (int)(HALF_SMEM / 16); i += THREADS) reinterpret_cast<float4 *>(s_in)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int nh = (s_mode == 1u) ? R - m : 0;           // static: the empty slots do not move, nobody hops

    int prev_v = -1;
    int2 g = make_int2(0, 0);
    float fill = 0.f;
    // (nh·t) mod R and (nh·t) mod m for the current step and the next one, advanced by additions (nh < m < R)
    // the arc of empty slots advances by `adv` slots per step (≤ its own length nh): adv windows hop per step
    const int adv = (a.rstride > 0 && a.rstride < nh) ? a.rstride : nh;
    int remR = 0, remM = 0, remR1 = adv, remM1 = adv;
    int v_next = rot_window_rem(slot, 0, 0, R, m);
    for (int t = 0; t < a.T; ++t) {
        const unsigned int it = (unsigned int)t;
        const int v = v_next;
        v_next = rot_window_rem(slot, remR1, remM1, R, m);          // holder of this slot at step t + 1
        remR = remR1; remM = remM1;
        remR1 += adv; if (remR1 >= R) remR1 -= R;
        remM1 += adv; if (remM1 >= m) remM1 -= m;
        if (v < 0) { prev_v = -1; continue; }
        if (v != prev_v) {
            fill = a.fill[v];
            if (t == 0) {
                g = a.guess[v];
            } else {
                // the window hops in from another SM: wait for its previous step (acquire), take the guess, clear the flag
                if (tid == 0) {
                    unsigned int f;
                    for (;;) {                 // poll relaxed (no L1 invalidation per poll), then one acquire
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(a.xflag + v) : "memory");
                        if (f == (unsigned int)t) break;
                        __nanosleep(32);
                    }
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(a.xflag + v) : "memory");
                    s_handoff[half] = a.xpos[v];
                    a.xflag[v] = 0u;
                }
                bar_half(half);
                g = s_handoff[half];
            }
            prev_v = v;
        }
        PT_PROBE_BEGIN(a, v, t, tid)
        const PixT *frame = reinterpret_cast<const PixT *>(a.frames) + (size_t)t * a.step_stride + (size_t)v * a.frame_stride;
        const int wy0 = g.x - 1 - g_rr, wx0 = g.y - 1 - g_rc;
        const int fy0 = wy0 - HW, fx0 = wx0 - HW;

        stage_tile<PixT>(frame, a.pitch, a.H, a.W, fy0, fx0, fill, s_in, warp, lane, g_flo, g_flo + g_nfr);

        if (t + 1 < a.T) {                                       // warm L2 with everything the next step can touch
            const PixT *nframe = frame + a.step_stride;
            constexpr int PR = FR + WR - 1;
            constexpr int NLMAX = (int)(((FC + WC) * sizeof(PixT) + 127) / 128) + 1;
            const int py0 = fy0 - WR / 2, pxb = (fx0 - WC / 2) * (int)sizeof(PixT);
            const int line0 = pxb >> 7;
            const int nl = ((pxb + (FC + WC - 1) * (int)sizeof(PixT) - 1) >> 7) - line0 + 1;
            const int rowbytes = a.W * (int)sizeof(PixT);
            static_assert(PR <= THREADS, "one thread per prefetched row");
            const int Y = py0 + tid;                              // thread = row: its 2-3 lines share one address
            if (tid < PR && Y >= 0 && Y < a.H) {
                const char *ptr = reinterpret_cast<const char *>(nframe + (size_t)Y * a.pitch) + (line0 << 7);
#pragma unroll
                for (int ln = 0; ln < NLMAX; ++ln) {
                    const int off = (line0 + ln) << 7;
                    if (ln < nl && off >= 0 && off < rowbytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + (ln << 7)));
                }
            }
        }
        if (tid == 0) {
            while (atomicCAS(&s_rowlock, 0, 1) != 0) __nanosleep(40);
        }
        bar_half(half);
        PT_PROBE(2, tid);

        if (kFull) row_pass45<FR>(s_in, s_mid, tid, tp);
        else row_pass45_rt<kSkip>(s_in, s_mid, tid, tp, a.f_lo, a.nfr, a.ng, a.inv_nfr);
        bar_half(half);
        if (tid == 0) atomicExch(&s_rowlock, 0);
        PT_PROBE(3, tid);

        const unsigned long long key = warp_max_key(col_pass45<kSkip>(s_mid, tid, tp, 0, 0, g_wr, g_wc, nullptr, g_cs, g_flo));
        if (lane == 0) s_key[(it & 1) * NWARPS + warp] = key;
        bar_half(half);
        PT_PROBE(4, tid);
        {
            // every warp folds the 8 warp keys itself (lane i reads key i mod 8, two REDUX): no serial section;
            // s_key is double-buffered by iteration parity
            const unsigned long long k = warp_max_key(s_key[(it & 1) * NWARPS + (lane & (NWARPS - 1))]);
            const unsigned int idx = key_index(k);
            const int xx = (int)(idx / (unsigned int)g_wr), yy = (int)(idx - xx * g_wr);
            const int raw_i = wy0 + yy + 1, raw_j = wx0 + xx + 1;
            const int ci = min(max(raw_i, 1), a.H), cj = min(max(raw_j, 1), a.W);
            if (tid == 0) {
                const float resp = key_value(k);
                const int4 p = make_int4(ci, cj, raw_i, raw_j);
                if (a.traj_pos) { a.traj_pos[(size_t)t * a.n + v] = p; a.traj_resp[(size_t)t * a.n + v] = resp; }
                if (t == a.T - 1) {
                    a.out_pos[v] = p; a.out_resp[v] = resp;
                    if (a.next_guess) a.next_guess[v] = make_int2(ci, cj);
                } else if (v_next != v) {
                    // the window hops to another SM for the next step: publish the guess, then the flag (release)
                    a.xpos[v] = make_int2(ci, cj);
                    const unsigned int f = (unsigned int)(t + 1);
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.xflag + v), "r"(f) : "memory");
                }
                PT_PROBE(5, 0);
            }
            g = make_int2(ci, cj);
        }
    }
    // the last CTA to leave zeroes the handshake words for the next launch (every CTA has read the decision by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int *xs = a.xflag + a.n;
        if (atomicAdd(xs + 2, 1u) == gridDim.x - 1u) { xs[0] = 0u; xs[1] = 0u; xs[2] = 0u; }
    }
}

// ---------------------------------------------------------------------------------------------------
// Large rectangles at l = 65 (auto-detect window size .÷ 4, src/PawsomeTracker.jl:99-105; the full-frame
// DoG benchmark shape; any non-default window_size at target_width 25).  The output rectangle is cut into
// strips of 45 columns and every strip into `nchunks` runs of 45-row batches.  One half-CTA takes one
// (window, chunk, strip) item and MARCHES down it:
//   batch 0      stage the 109-row footprint, row pass over 109 rows, column pass → 45×45 outputs
//                (exactly a window of dog_window45_argmax);
//   batch b > 0  the last 64 rows of the row-pass intermediate are still valid: move them to the top of
//                s_mid, stage only the 45 NEW footprint rows, row pass over those 45 rows, column pass.
// So inside a chunk the row pass runs once per footprint row (the algorithmic count) instead of 109 rows
// per 45 outputs; with nchunks = number of batches this degenerates to independent 45×45 tiles, which is
// what small rectangles (fewer tiles than SMs) use.  The running argmax stays in registers across the
// batches of an item; items of a rectangle combine through one 64-bit atomicMax per item and the last one
// decodes, clamps and publishes — the merge of the generic kernel.  Half h of CTA c walks items c + S·h,
// then whatever the ticket counter hands it (the token alternation of dog_window45_argmax made no measurable
// difference here — both phases of a marching batch are FMA-heavy — and is not used).
// ---------------------------------------------------------------------------------------------------
struct March45 {
    int n, ntx, nty, nchunks;
};

__device__ __forceinline__ void chunk_range(const March45 &m, int c, int &b0, int &nb)
{
    const int base = m.nty / m.nchunks, extra = m.nty - base * m.nchunks;
    nb = base + (c < extra ? 1 : 0);
    b0 = c * base + min(c, extra);
}

template <typename PixT>
__global__ void __launch_bounds__(CTA_THREADS, 1)
dog_rect45_march(const __grid_constant__ WinArgs a, const __grid_constant__ Taps45 tp, const __grid_constant__ March45 m)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_keys[2][NWARPS];
    __shared__ unsigned int s_ticket[2];

    const int pw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = (pw >> 2) & 1;
    const int warp = (((pw & 3) + 2 * half) & 3) + 4 * (pw >> 3);
    const int tid = warp * 32 + lane;
    float *s_in = reinterpret_cast<float *>(smem_raw + half * HALF_SMEM);
    float2 *s_mid = reinterpret_cast<float2 *>(s_in + FR * PIN + 1);
    unsigned long long *s_key = s_keys[half];
    const int per_win = m.ntx * m.nchunks;
    const unsigned int items = (unsigned int)m.n * (unsigned int)per_win;

    // Items are handed out by a ticket counter: a half that shares its SM runs ≈ 1.6× slower than one
    // that has the SM to itself, so a static split would leave SMs idle at the end.  The first two
    // tickets of a CTA are static (CTA c: c and S + c); the ticket of the NEXT item is fetched while the
    // current one is processed, so its latency is never exposed.
    unsigned int item = (unsigned int)blockIdx.x + (unsigned int)gridDim.x * (unsigned int)half;
    const unsigned int static_items = 2u * gridDim.x;

    while (item < items) {
        unsigned int next = 0u;
        if (tid == 0) next = static_items + atomicAdd(a.tickets, 1u);
        const int v = (int)(item / (unsigned int)per_win), rem = (int)(item - (unsigned int)v * (unsigned int)per_win);
        const int c = rem / m.ntx, tx = rem - c * m.ntx;
        int b0, nb;
        chunk_range(m, c, b0, nb);
        int wy0, wx0;
        if (a.rect_mode) { wy0 = a.ry0; wx0 = a.rx0; }
        else { const int2 g = a.guess[v]; wy0 = g.x - 1 - a.rr; wx0 = g.y - 1 - a.rc; }
        const int gx0 = tx * WC;
        const int fx0 = wx0 + gx0 - HW;
        const float fill = a.fill[v];
        const PixT *frame = reinterpret_cast<const PixT *>(a.frames) + (size_t)v * a.frame_stride;
        float *map = a.map_out ? a.map_out + (size_t)v * a.wr * a.wc : nullptr;
        unsigned long long best = 0ull;

        for (int b = 0; b < nb; ++b) {
            const int gy0 = (b0 + b) * WR;
            constexpr int NKEEP = (2 * HW * WC + THREADS - 1) / THREADS;
            float2 keep[NKEEP];                                          // the 64 still-valid s_mid rows (2880 float2)
            if (b == 0) stage_rows<FR>(frame, a.pitch, a.H, a.W, wy0 + gy0 - HW, fx0, fill, s_in, warp, lane);
            else stage_rows<WR>(frame, a.pitch, a.H, a.W, wy0 + gy0 + HW, fx0, fill, s_in, warp, lane);
            // warm L2 with the 45 new footprint rows of the next batch
            if (b + 1 < nb) {
                constexpr int NL = (int)((FC * sizeof(PixT) + 127) / 128) + 1;
                const int pxb = fx0 * (int)sizeof(PixT);
                const int line0 = pxb >> 7;
                const int nl = ((pxb + FC * (int)sizeof(PixT) - 1) >> 7) - line0 + 1;
                const int rowbytes = a.W * (int)sizeof(PixT);
                const int py0 = wy0 + gy0 + WR + HW;
                for (int e = tid; e < WR * NL; e += THREADS) {
                    const int r = e / NL, ln = e - r * NL;
                    const int Y = py0 + r;
                    const int off = (line0 + ln) << 7;
                    if (ln < nl && Y >= 0 && Y < a.H && off >= 0 && off < rowbytes)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(frame + (size_t)Y * a.pitch) + off));
                }
            }
            if (b > 0) {
#pragma unroll
                for (int i = 0; i < NKEEP; ++i) {
                    const int e = tid + i * THREADS;
                    if (e < 2 * HW * WC) keep[i] = s_mid[WR * PM + e];
                }
            }
            bar_half(half);
            if (b == 0) {
                row_pass45<FR>(s_in, s_mid, tid, tp);
            } else {
#pragma unroll
                for (int i = 0; i < NKEEP; ++i) {
                    const int e = tid + i * THREADS;
                    if (e < 2 * HW * WC) s_mid[e] = keep[i];
                }
                row_pass45<WR>(s_in, s_mid + 2 * HW * PM, tid, tp);
            }
            bar_half(half);
            const unsigned long long key = col_pass45(s_mid, tid, tp, gy0, gx0, a.wr, a.wc, map);
            best = key > best ? key : best;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, best, off);
            best = o > best ? o : best;
        }
        if (lane == 0) s_key[warp] = best;
        if (tid == 0) s_ticket[half] = next;
        bar_half(half);      // also: every column-pass read of s_mid is done before the next item's row pass
        item = s_ticket[half];
        if (tid == 0) {
            unsigned long long k = s_key[0];
#pragma unroll
            for (int i = 1; i < NWARPS; ++i) k = s_key[i] > k ? s_key[i] : k;
            atomicMax(a.keys + v, k);
            __threadfence();
            const unsigned int prev = atomicAdd(a.counters + v, 1u);
            if (prev == (unsigned int)per_win - 1u) {
                __threadfence();
                const unsigned long long win = atomicExch(a.keys + v, 0ull);
                a.counters[v] = 0u;
                publish_result(a, v, win, wy0, wx0);
            }
        }
        // (s_key is rewritten only after the next item's barriers, which thread 0 has to pass too)
    }
    // leave the ticket counter zeroed for the next launch: the last half to run out of work resets it
    // (every half has fetched its final, failing ticket by then)
    if (tid == 0) {
        const unsigned int prev = atomicAdd(a.tickets + 1, 1u);
        if (prev == 2u * gridDim.x - 1u) { a.tickets[0] = 0u; a.tickets[1] = 0u; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------------------------
// dog_window45_cluster<PixT, C> — ONE window spread over a thread-block cluster of C CTAs (C = 2, 4, 8), for
// batches with fewer windows than SMs (a single video above all: the frame loop ij[t] = trckr(ij[t-1]),
// src/PawsomeTracker.jl:167, is a serial chain, so a lone window is pure latency).
//
// The 45 output columns are cut into C slices.  Both passes of a slice are independent of the other slices:
// the column pass of output column x only needs the row-pass intermediate of column x, and that only needs the
// pixels of columns x−32 … x+32.  So the CTAs of a cluster exchange NOTHING but their argmax candidates: each
// warp stores its 64-bit key into the shared memory of every CTA of the cluster (DSMEM), one cluster barrier,
// every CTA folds the 8·C keys and knows the next guess.  The price is redundant staging (each CTA stages
// 109 × (slice + 64) pixels).
//
// Staging (u8 frames, 16-byte aligned rows): the next window centre lies inside the current window, so the
// region any next footprint can touch — 153 rows × (slice + 64 + 44) bytes — is known one step ahead.  It is
// fetched with ONE 2-D TMA tile copy into a double-buffered u8 region in shared memory while the
// current step computes; when the guess is known the footprint is converted u8 → f32 out of shared memory:
// no global-memory latency on the serial chain.  A window that left the prefetched region (possible only when
// the guess was outside the frame and got clamped) re-fetches its own region.  Other frames (f32, unaligned)
// are staged straight from global memory with the L2 prefetch of dog_window45_argmax.
//
// Per step a CTA runs: [stage own 14 rows → row pass of those rows] per warp, independently (no CTA barrier in
// between: a warp's load latency hides behind the other warps' FMAs) → one CTA barrier → column pass → candidates to
// every CTA of the cluster with st.async + mbarrier (no cluster-wide rendezvous; the wait doubles as the CTA's
// second barrier) → fold, next guess.
//
// Layouts: s_in [109][PINS] f32 with PINS chosen so that the (row, group) lanes of a row-pass warp hit distinct
// banks; s_midT [slice column][109] float2 — the column pass walks rows of one column, items ordered
// row-group-fastest: consecutive items are RC·(item) float2 apart (mod 16 bank pairs), i.e. conflict-free for
// any 16 consecutive lanes.
// ---------------------------------------------------------------------------------------------------
template <int C>
struct SliceGeom {
    static constexpr int SW = (WC + C - 1) / C;              // widest slice: 23 / 12 / 6 columns
    static constexpr int RRS = (C == 8) ? 3 : 6;             // row pass: outputs per thread
    static constexpr int NGR = (SW + RRS - 1) / RRS;         // groups per row: 4 / 2 / 2
    static constexpr int SWC = NGR * RRS;                    // computed columns (≥ SW; the surplus is masked)
    static constexpr int SFC = SWC + 2 * HW;                 // staged footprint columns: 88 / 76 / 70
    static constexpr int RPW = (FR + NWARPS - 1) / NWARPS;   // 14 consecutive footprint rows per warp (stage AND row pass)
    // s_in pitch: the lanes of a row-pass warp are (row r < 14, group g) at r·PINS + g·RRS — these pitches make the
    // 28 (C = 4, 8) / 32-at-a-time (C = 2) addresses fall into distinct banks
    static constexpr int PINS = (C == 2) ? 101 : (C == 4) ? 85 : 70;
    static constexpr int NW = (SFC + 3 + 3) / 4;             // aligned words per staged row at any phase
    static constexpr int RC = (C == 8) ? 3 : 5;              // column pass: outputs per thread
    static constexpr int NGC = WR / RC;                      // row groups: 9 / 15
    static constexpr int RGN_ROWS = FR + 2 * (WR / 2);       // 153 rows any next footprint can touch
    static constexpr int SPAN = ((4 * NW + 2 * (WC / 2) + 15 + 15) / 16) * 16;   // bytes per region row (16-byte aligned start)
    static constexpr int RGN_BYTES = ((RGN_ROWS * SPAN + 127) / 128) * 128;   // 128-byte multiple: TMA tile destinations
    static constexpr size_t IN_BYTES = ((size_t)FR * PINS * sizeof(float) + 15) & ~(size_t)15;
    static constexpr size_t MID_BYTES = (size_t)SWC * FR * sizeof(float2);
    static_assert(WR % RC == 0, "row groups must tile the window");
    static_assert(NW <= 32, "one lane per staged word");
    static_assert(SW * NGC <= 256 && RGN_ROWS <= 256, "one item / one region row per thread");
    static size_t smem_bytes(bool bulk) { return (bulk ? 2 * (size_t)RGN_BYTES + 128 : 0) + IN_BYTES + MID_BYTES; }
};
constexpr int CL_THREADS = 256;

// The whole region of a step in ONE instruction: 2-D TMA tile copy (cp.async.bulk.tensor → SASS UTMALDG) out of the
// tensor map that describes the resident frames as [rows][pitch] bytes; elements outside the tensor arrive as zeros
// (and are masked by frame coordinates anyway).
__device__ __forceinline__ void tma_tile_2d(unsigned int dst, const CUtensorMap *tmap, int x, int y, unsigned int mbar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(x), "r"(y), "r"(mbar) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned int cluster_rank()
{
    unsigned int r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned int mapa_u32(unsigned int local_addr, unsigned int rank)
{
    unsigned int remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    return remote;
}
// 8-byte store into the shared memory of CTA `rank` of the cluster that completes 8 bytes on that CTA's mbarrier:
// the receiver only waits on its own mbarrier, no cluster-wide rendezvous.
__device__ __forceinline__ void st_async_u64(unsigned int local_addr, unsigned int local_mbar, unsigned int rank, unsigned long long v)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(mapa_u32(local_addr, rank)), "l"(v), "r"(mapa_u32(local_mbar, rank)) : "memory");
}

// (conversion of a slice out of the prefetched region: convert_rows_u8 with the slice geometry.  Flattening the
// (row, word) pairs over all lanes was measured: no gain, more registers.)
template <int C, bool kInterior>
__device__ __forceinline__ void convert_slice_u8(const uint8_t *rgn, int rgn_y0, int rgn_xa, int H, int W, int fy0, int fxs,
                                                 float fill, float *s_in, int warp, int lane)
{
    using G = SliceGeom<C>;
    convert_rows_u8<G::SPAN, G::SFC, G::PINS, kInterior>(rgn, rgn_y0, rgn_xa, H, W, fy0, fxs, fill, s_in, warp, lane);
}

// Row pass of a slice over the 14 rows THIS WARP staged (no CTA barrier between staging and row pass: warps run
// through both independently, one warp's load latency hides behind another's FMAs): item = (row r of the warp, group g
// of RRS output columns), same folded arithmetic and summation order as row_pass45.  Output → s_midT[column][row].
template <int C>
__device__ __forceinline__ void row_pass_slice(const float *s_in, float2 *s_midT, int warp, int lane, const Taps45 &tp)
{
    using G = SliceGeom<C>;
    constexpr int RRS = G::RRS, RPW = G::RPW;
    const int f0 = warp * RPW;
    const int nrows = min(RPW, FR - f0);                     // 14 (11 for the last warp)
#pragma unroll 1
    for (int item = lane; item < RPW * G::NGR; item += 32) {
        const int g = item / RPW, r = item - g * RPW;
        if (r >= nrows) continue;
        const int f = f0 + r;
        const float *row = s_in + f * G::PINS + g * RRS;
        float x[RRS + 2 * HW];
#pragma unroll
        for (int i = 0; i < RRS + 2 * HW; ++i) x[i] = row[i];
        float2 acc[RRS];
#pragma unroll
        for (int j = 0; j < RRS; ++j) acc[j] = fmul2(make_float2(x[j + HW], x[j + HW]), tp.rt[0]);
#pragma unroll
        for (int d = 1; d <= HW; ++d) {
#pragma unroll
            for (int j = 0; j < RRS; ++j) {
                const float sm = x[j + HW - d] + x[j + HW + d];
                acc[j] = ffma2(make_float2(sm, sm), tp.rt[d], acc[j]);
            }
        }
        float2 *dst = s_midT + (g * RRS) * FR + f;
#pragma unroll
        for (int j = 0; j < RRS; ++j) dst[j * FR] = acc[j];
    }
}

// Column pass of a slice + per-thread argmax: item = (slice column x, group h of RC output rows), h fastest.
// Same per-output operation order as col_pass45 (packed pairs of vertically adjacent outputs, narrow and wide
// parts accumulated separately, added at the end).  `width` columns of the slice are real; the slice's column 0
// is column gx0 of a wr_tot × wc_tot output rectangle whose row 0 is this tile's row gy0.
template <int C>
__device__ __forceinline__ unsigned long long col_pass_slice(const float2 *s_midT, int tid, const Taps45 &tp, int width,
                                                             int gy0, int gx0, int wr_tot, int wc_tot)
{
    using G = SliceGeom<C>;
    constexpr int RC = G::RC, NP = RC / 2;
    if (tid >= G::SW * G::NGC) return 0ull;
    const int x = tid / G::NGC, h = tid - x * G::NGC;
    const float2 *col = s_midT + x * FR + h * RC;
    float2 accP[NP], accM[NP];
    float acc_lp = 0.f, acc_lm = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) { accP[p] = make_float2(0.f, 0.f); accM[p] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int i = 0; i < RC + 2 * HW; ++i) {
        const float2 m = col[i];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int q = i - 2 * p;
            if (q >= 0 && q <= L) accP[p] = ffma2(make_float2(m.x, m.x), tp.cpp[q], accP[p]);
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int q = i - 2 * p;
            if (q >= 0 && q <= L) accM[p] = ffma2(make_float2(m.y, m.y), tp.cmq[q], accM[p]);
        }
        const int ql = i - (RC - 1);
        if (ql >= 0 && ql < L) {
            acc_lp = fmaf(m.x, tp.cpp[ql].x, acc_lp);
            acc_lm = fmaf(m.y, tp.cmq[ql].x, acc_lm);
        }
    }
    float acc[RC];
#pragma unroll
    for (int p = 0; p < NP; ++p) { acc[2 * p] = accP[p].x + accM[p].x; acc[2 * p + 1] = accP[p].y + accM[p].y; }
    acc[RC - 1] = acc_lp + acc_lm;
    const int gx = gx0 + x, gyb = gy0 + h * RC;
    if (x >= width || gx >= wc_tot || gyb >= wr_tot) return 0ull;
    float bv = acc[0] + 0.0f;
    int bj = 0;
#pragma unroll
    for (int j = 1; j < RC; ++j) {
        const float val = acc[j] + 0.0f;
        if (gyb + j < wr_tot && val > bv) { bv = val; bj = j; }
    }
    return pack_key(bv, (unsigned int)(gx * wr_tot + gyb + bj));
}

template <typename PixT, int C, bool kFull>
__global__ void __launch_bounds__(CL_THREADS, 2)
dog_window45_cluster(const __grid_constant__ Args45 a, const __grid_constant__ Taps45 tp, const int use_bulk,
                     const __grid_constant__ CUtensorMap tmap)
{
    using G = SliceGeom<C>;
    // (kFull: default geometry, compile-time bounds; else a shorter kernel — zero-padded taps — or a smaller window)
    const int g_rr = kFull ? WR / 2 : a.rr, g_rc = kFull ? WC / 2 : a.rc, g_wr = kFull ? WR : a.wr, g_wc = kFull ? WC : a.wc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_mbar[2];                // arrival of the prefetched regions
    __shared__ __align__(8) unsigned long long s_xbar[2];                // arrival of the argmax candidates (by step parity)
    __shared__ __align__(8) unsigned long long s_xk[2][NWARPS * C];      // candidates of every warp of every CTA of the cluster

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_rank();
    const int v = (int)blockIdx.x / C;
    const int xs = (g_wc * rank) / C, width = (g_wc * (rank + 1)) / C - xs;  // this CTA's output columns [xs, xs + width)
    constexpr bool kU8 = sizeof(PixT) == 1;
    const bool bulk = kU8 && use_bulk != 0;
    unsigned char *rgn = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);      // 128-byte aligned (TMA tile destination)
    float *s_in = reinterpret_cast<float *>(smem_raw + (bulk ? 2 * G::RGN_BYTES + 128 : 0));
    float2 *s_midT = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(s_in) + G::IN_BYTES);
    const unsigned int mbar0 = smem_u32(&s_mbar[0]);
    const unsigned int xbar0 = smem_u32(&s_xbar[0]);
    const unsigned int rgn0 = smem_u32(rgn);

    if (tid == 0) {
        mbar_init(mbar0, 1u);
        mbar_init(mbar0 + 8u, 1u);
        mbar_init(xbar0, 1u);
        mbar_init(xbar0 + 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_arrive();            // no CTA touches a peer's shared memory before every CTA of the cluster runs
    cluster_wait();

    const float fill = a.fill[v];
    int2 g = a.guess[v];
    // frames: resident in HBM (base + strides) or, zero-copy, a table of page-locked host frames (staged with global
    // loads over PCIe: each CTA reads the rows of its slice)
    const PixT *frame0 = a.frame_ptrs ? nullptr : reinterpret_cast<const PixT *>(a.frames) + (size_t)v * a.frame_stride;
    int rgn_y0[2] = {0, 0}, rgn_xa[2] = {0, 0};
    unsigned int ph[2] = {0u, 0u}, xph[2] = {0u, 0u};

    // request the region around footprint origin (cfy0, cfxs) of step `ts` into buffer `buf`: ONE TMA tile copy
    // (elements outside the tensor arrive as zeros; everything outside the frame is masked at conversion anyway)
    auto issue_region = [&](int buf, int ts, int cfy0, int cfxs) {
        const int y0 = cfy0 - WR / 2, xa = (cfxs - WC / 2) & ~15;
        rgn_y0[buf] = y0; rgn_xa[buf] = xa;
        if (tid == 0) {
            const unsigned int mb = mbar0 + 8u * (unsigned int)buf;
            mbar_arrive_expect_tx(mb, (unsigned int)(G::RGN_ROWS * G::SPAN));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tma_tile_2d(rgn0 + (unsigned int)(buf * G::RGN_BYTES), &tmap, xa, ts * a.tm_rows_step + v * a.tm_rows_frame + y0, mb);
        }
    };

    if (bulk) issue_region(0, 0, g.x - 1 - g_rr - HW, g.y - 1 - g_rc - HW + xs);

    for (int t = 0; t < a.T; ++t) {
        const int par = t & 1;
        PT_PROBE_BEGIN(a, v, t, tid + rank)
        // this step's candidates: 8 bytes from every warp of every CTA of the cluster (this CTA's own included)
        if (tid == 0) mbar_arrive_expect_tx(xbar0 + 8u * par, 8u * NWARPS * C);
        const PixT *frame = a.frame_ptrs ? reinterpret_cast<const PixT *>(a.frame_ptrs[(size_t)t * a.n + v])
                                         : frame0 + (size_t)t * a.step_stride;
        const int wy0 = g.x - 1 - g_rr, wx0 = g.y - 1 - g_rc;
        const int fy0 = wy0 - HW, fxs = wx0 - HW + xs;                       // footprint origin of this slice
        if (bulk) {
            const int xw0 = fxs & ~3;
            const bool covered = fy0 >= rgn_y0[par] && fy0 + FR <= rgn_y0[par] + G::RGN_ROWS &&
                                 xw0 >= rgn_xa[par] && xw0 + 4 * G::NW <= rgn_xa[par] + G::SPAN;
            mbar_wait(mbar0 + 8u * par, ph[par]); ph[par] ^= 1u;
            if (!covered) {                                                  // (uniform) the window left the prefetched region
                __syncthreads();                                             // every thread has seen the completed phase
                issue_region(par, t, fy0, fxs);
                mbar_wait(mbar0 + 8u * par, ph[par]); ph[par] ^= 1u;
            }
            const bool interior = fy0 >= 0 && fy0 + FR <= a.H && xw0 >= 0 && xw0 + 4 * G::NW <= a.W;
            const uint8_t *rb = rgn + par * G::RGN_BYTES;
            if (interior) convert_slice_u8<C, true>(rb, rgn_y0[par], rgn_xa[par], a.H, a.W, fy0, fxs, fill, s_in, warp, lane);
            else convert_slice_u8<C, false>(rb, rgn_y0[par], rgn_xa[par], a.H, a.W, fy0, fxs, fill, s_in, warp, lane);
            // everything the next step can touch → the other buffer (last read by the previous step's conversion)
            if (t + 1 < a.T) issue_region(par ^ 1, t + 1, fy0, fxs);
        } else {
            stage_rows<FR, G::SFC, G::PINS, true>(frame, a.pitch, a.H, a.W, fy0, fxs, fill, s_in, warp, lane);
            if (t + 1 < a.T && !a.frame_ptrs) {                              // warm L2 with the next step's region
                const PixT *nframe = frame + a.step_stride;
                constexpr int NLMAX = (int)(((G::SFC + WC) * sizeof(PixT) + 127) / 128) + 1;
                const int pxb = (fxs - WC / 2) * (int)sizeof(PixT);
                const int line0 = pxb >> 7;
                const int nl = ((pxb + (G::SFC + WC - 1) * (int)sizeof(PixT) - 1) >> 7) - line0 + 1;
                const int rowbytes = a.W * (int)sizeof(PixT);
                const int Y = fy0 - WR / 2 + tid;
                if (tid < G::RGN_ROWS && Y >= 0 && Y < a.H) {
                    const char *ptr = reinterpret_cast<const char *>(nframe + (size_t)Y * a.pitch) + (line0 << 7);
#pragma unroll
                    for (int ln = 0; ln < NLMAX; ++ln) {
                        const int off = (line0 + ln) << 7;
                        if (ln < nl && off >= 0 && off < rowbytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + (ln << 7)));
                    }
                }
            }
        }
        __syncwarp();                        // a warp row-filters exactly the rows it staged: no CTA barrier here
        PT_PROBE(2, tid + rank);

        row_pass_slice<C>(s_in, s_midT, warp, lane, tp);
        __syncthreads();                     // the column pass reads every warp's rows of s_midT
        PT_PROBE(3, tid + rank);

        unsigned long long key = warp_max_key(col_pass_slice<C>(s_midT, tid, tp, width, 0, xs, g_wr, g_wc));
        // Every warp hands its candidate to every CTA of the cluster (its own included) with st.async, which also
        // completes 8 bytes on the receiver's mbarrier; a CTA waits only for ITS 8·C·8 bytes — no cluster-wide
        // rendezvous.  Slots and mbarriers are double-buffered by step parity: a CTA can only send step t+2 after it
        // received every CTA's step t+1 candidates, which each of them sends only after it has read step t's slots.
        if (lane < C) st_async_u64(smem_u32(&s_xk[par][rank * NWARPS + warp]), xbar0 + 8u * par, (unsigned int)lane, key);
        // (this wait is also the CTA's barrier between this step's column pass and the next step's row pass: all eight
        // warps of this CTA have sent their candidates, i.e. finished reading s_midT, before it completes)
        mbar_wait(xbar0 + 8u * par, xph[par]); xph[par] ^= 1u;
        PT_PROBE(4, tid + rank);
        {
            unsigned long long k = s_xk[par][lane % (NWARPS * C)];
            if (NWARPS * C > 32) { const unsigned long long k2 = s_xk[par][32 + lane]; k = k2 > k ? k2 : k; }
            k = warp_max_key(k);
            const unsigned int idx = key_index(k);
            const int xx = (int)(idx / (unsigned int)g_wr), yy = (int)(idx - xx * g_wr);
            const int raw_i = wy0 + yy + 1, raw_j = wx0 + xx + 1;                   // absolute index (:60)
            const int ci = min(max(raw_i, 1), a.H), cj = min(max(raw_j, 1), a.W);   // clamp (:61)
            if (tid == 0 && rank == 0) {
                const float resp = key_value(k);
                const int4 p = make_int4(ci, cj, raw_i, raw_j);
                if (a.traj_pos) { a.traj_pos[(size_t)t * a.n + v] = p; a.traj_resp[(size_t)t * a.n + v] = resp; }
                if (t == a.T - 1) {
                    a.out_pos[v] = p; a.out_resp[v] = resp;
                    if (a.next_guess) a.next_guess[v] = make_int2(ci, cj);
                }
                PT_PROBE(5, 0);
            }
            g = make_int2(ci, cj);
        }
    }
    // A CTA may only exit when no peer can still write into its shared memory: it has received all candidates of the
    // last step by then, and every earlier store was received before that.  Its own outgoing stores target CTAs that
    // are still waiting for them.  One closing barrier keeps the cluster's lifetime simple and costs once per launch.
    cluster_arrive();
    cluster_wait();
}

// How many CTAs share one window for this launch (1 = the per-SM kernels above).  Measured (tools/small_batch_timing.py,
// 1080p, 100 chained steps, µs per step): 8 CTAs per window win for a handful of windows (n = 1 … 9: 2.2-2.3 vs 2.5 with
// 4 CTAs; from n = 16 clusters of 8 no longer pack the GPCs one CTA per SM: 3.2 vs 2.55), 4 CTAs up to n = 32-33
// (2.56; at n = 37 = #SMs/4 the clusters of 4 do not all fit one per SM either: 4.0 vs 3.85 with 2 CTAs), 2 CTAs up to
// #SMs/2 (n = 64 … 74: 3.84 vs 5.75 for the per-SM kernel).
static int cluster_size_for(const WinArgs &a, const Cfg &cfg, int n)
{
    if (cfg.cluster == 1) return 1;
    if (cfg.cluster == 2 || cfg.cluster == 4 || cfg.cluster == 8) return cfg.cluster;
    const int sms = cfg.sms;
    if (a.host_frames) {
        // page-locked host frames at regular strides (zero-copy over PCIe, region prefetched one step ahead), measured
        // with tools/pinned_chain_timing.py, µs per frame: n = 1: 4.2 (2 or 4 CTAs, TMA) vs 7.5 per-SM; n = 4: 6.2-6.7 vs 7.5;
        // n = 8: 6.6 (2 CTAs, global loads) vs 7.7; n = 16: per-SM 9.0 vs 22-26 (the slices' redundant rows cost PCIe bytes)
        return n <= 8 ? 2 : 1;
    }
    if (a.frame_ptrs) {
        // zero-copy host frames: every CTA of a cluster pulls its own slice rows over PCIe (C = 4: 3.5x the bytes of
        // one footprint), which is free for a handful of windows and costs bandwidth for many
        if (16 * n <= sms) return 4;
        if (4 * n <= sms) return 2;
        return 1;
    }
    if (16 * n <= sms) return 8;
    if (4 * n <= sms - 16) return 4;
    if (2 * n <= sms) return 2;
    return 1;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda: the library must load
// on a machine without a driver).  nullptr = unavailable → the cluster kernel stages with global loads instead.
typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder()
{
    static tmap_encode_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (tmap_encode_fn)sym;
        else
            cudaGetLastError();
    });
    return fn;
}

static bool full_geometry(const Args45 &k) { return k.wr == WR && k.wc == WC && k.f_lo == 0 && k.nfr == FR; }

template <typename PixT, int C>
static cudaError_t launch_cluster_t(Args45 k, const Taps45 &tp, int use_bulk, cudaStream_t s)
{
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if (use_bulk) {
        // the resident frames as one 2-D u8 tensor [rows][pitch]: row of (step t, video v, frame row y) =
        // t·rows_step + v·rows_frame + y.  Needs strides that are whole rows; else fall back to row copies.
        tmap_encode_fn enc = tmap_encoder();
        const bool regular = enc && k.pitch > 0 && k.frame_stride % (size_t)k.pitch == 0 && k.step_stride % (size_t)k.pitch == 0;
        bool ok = false;
        if (regular) {
            k.tm_rows_frame = (int)(k.frame_stride / (size_t)k.pitch);
            k.tm_rows_step = (int)(k.step_stride / (size_t)k.pitch);
            const cuuint64_t rows = (cuuint64_t)(k.T - 1) * (cuuint64_t)k.tm_rows_step + (cuuint64_t)(k.n - 1) * (cuuint64_t)k.tm_rows_frame + (cuuint64_t)k.H;
            const cuuint64_t dims[2] = {(cuuint64_t)k.pitch, rows};
            const cuuint64_t strides[1] = {(cuuint64_t)k.pitch};
            const cuuint32_t box[2] = {(cuuint32_t)SliceGeom<C>::SPAN, (cuuint32_t)SliceGeom<C>::RGN_ROWS};
            const cuuint32_t estr[2] = {1u, 1u};
            ok = rows < (1ull << 31) &&
                 enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2u, const_cast<void *>(k.frames), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        if (!ok) use_bulk = 0;               // no tensor map (irregular strides, old driver): stage with global loads
    }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(k.n * C));
    lc.blockDim = dim3(CL_THREADS);
    lc.dynamicSmemBytes = SliceGeom<C>::smem_bytes(use_bulk != 0);
    lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    if (full_geometry(k)) return cudaLaunchKernelEx(&lc, dog_window45_cluster<PixT, C, true>, k, tp, use_bulk, tmap);
    return cudaLaunchKernelEx(&lc, dog_window45_cluster<PixT, C, false>, k, tp, use_bulk, tmap);
}

static cudaError_t launch_cluster(const Args45 &k, const Taps45 &tp, const Cfg &cfg, int C, int pixel, cudaStream_t s)
{
    // the TMA path needs 16-byte aligned rows
    const bool aligned16 = pixel == 0 && !k.frame_ptrs &&
                           ((reinterpret_cast<uintptr_t>(k.frames) | (uintptr_t)k.pitch | (uintptr_t)k.frame_stride |
                             (uintptr_t)k.step_stride) & 15u) == 0;
    // the TMA path needs 16-byte aligned rows; over PCIe (host frames) its 153-row regions pay off for one or two windows only
    const int use_bulk = (cfg.bulk && aligned16 && !(k.host_frames && k.n > 2)) ? 1 : 0;
    if (pixel == 0) {
        if (C == 2) return launch_cluster_t<uint8_t, 2>(k, tp, use_bulk, s);
        if (C == 4) return launch_cluster_t<uint8_t, 4>(k, tp, use_bulk, s);
        return launch_cluster_t<uint8_t, 8>(k, tp, use_bulk, s);
    }
    if (C == 2) return launch_cluster_t<float, 2>(k, tp, 0, s);
    if (C == 4) return launch_cluster_t<float, 4>(k, tp, 0, s);
    return launch_cluster_t<float, 8>(k, tp, 0, s);
}

#define PT_CLUSTER_OPTINS_G(F)                                                          \
    PT_OPTIN((dog_window45_cluster<uint8_t, 2, F>), SliceGeom<2>::smem_bytes(true))     \
    PT_OPTIN((dog_window45_cluster<uint8_t, 4, F>), SliceGeom<4>::smem_bytes(true))     \
    PT_OPTIN((dog_window45_cluster<uint8_t, 8, F>), SliceGeom<8>::smem_bytes(true))     \
    PT_OPTIN((dog_window45_cluster<float, 2, F>), SliceGeom<2>::smem_bytes(false))      \
    PT_OPTIN((dog_window45_cluster<float, 4, F>), SliceGeom<4>::smem_bytes(false))      \
    PT_OPTIN((dog_window45_cluster<float, 8, F>), SliceGeom<8>::smem_bytes(false))
#define PT_CLUSTER_OPTINS PT_CLUSTER_OPTINS_G(true) PT_CLUSTER_OPTINS_G(false)

// ===================================================================================================
// host side
// ===================================================================================================
const char *rect45_name() { return "dog_rect45_march"; }

bool rect45_supported(const WinArgs &a, const Cfg &cfg, int pixel)
{
    if (a.L != L || !cfg.rect45) return false;
    if ((long long)a.wr * a.wc < 24 * 24) return false;     // tiny windows: the generic strip kernel wastes less
    if (pixel == 0) {
        const bool aligned = ((reinterpret_cast<uintptr_t>(a.frames) | (uintptr_t)a.pitch | (uintptr_t)a.frame_stride) & 3u) == 0;
        if (!aligned || a.pitch < ((a.W + 3) & ~3)) return false;
    }
    return true;
}

const char *window45_name() { return "dog_window45_argmax"; }

#ifdef PT_PROBES
static long long *g_dbg = nullptr;
void window45_set_debug(long long *dev_buf) { g_dbg = dev_buf; }
#endif

bool window45_supported(const WinArgs &a, int pixel)
{
    // any kernel length up to 65 (zero-padded taps) and any window up to 45×45 (masked outputs, rows and column
    // groups without work skipped): target_width ≤ 25 with its default window
    if (!(a.L <= L && a.wr <= WR && a.wc <= WC && !a.rect_mode && a.map_out == nullptr)) return false;
    if (pixel == 0 && a.frame_ptrs) return (a.pitch & 3) == 0 && a.pitch >= ((a.W + 3) & ~3);   // caller checked the pointers
    if (pixel == 0 && a.frames) {
        // the u8 staging path loads aligned 32-bit words
        const bool aligned = ((reinterpret_cast<uintptr_t>(a.frames) | (uintptr_t)a.pitch | (uintptr_t)a.frame_stride |
                               (uintptr_t)a.step_stride) & 3u) == 0;
        if (!aligned || a.pitch < ((a.W + 3) & ~3)) return false;
    }
    return true;
}

// The specialised kernels take their taps as kernel parameters (constant bank):
// fold the symmetric row factors (index d = |k − 32|) and pair the column taps.
static void fold_taps(const WinArgs &a, Taps45 &tp)
{
    // a kernel of length a.L ≤ 65 sits centred in the 65 taps of these kernels, zeros around it (adding 0·x changes
    // no sum)
    const int off = HW - a.L / 2;
    const float *rp = a.h_taps, *rm = a.h_taps + a.L, *cp = a.h_taps + 2 * a.L, *cm = a.h_taps + 3 * a.L;
    auto at = [&](const float *t, int k) { k -= off; return (k >= 0 && k < a.L) ? t[k] : 0.f; };
    for (int d = 0; d <= HW; ++d) tp.rt[d] = make_float2(at(rp, HW + d), at(rm, HW + d));
    for (int q = 0; q <= L; ++q) {
        tp.cpp[q] = make_float2(at(cp, q), at(cp, q - 1));
        tp.cmq[q] = make_float2(at(cm, q), at(cm, q - 1));
    }
}

// Once per (process, device), with the device current (pt_batch_create holds the lock): dynamic shared memory
// opt-in of every kernel of this file.
cudaError_t window45_init_device()
{
    const int smem = (int)(2 * HALF_SMEM);
    cudaError_t e;
#define PT_OPTIN(k, bytes)                                                                      \
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));       \
    if (e != cudaSuccess) return e;
#define PT_OPTIN_G(G)                                      \
    PT_OPTIN((dog_window45_argmax<uint8_t, G>), smem)      \
    PT_OPTIN((dog_window45_argmax<float, G>), smem)        \
    PT_OPTIN((dog_window45_rot<uint8_t, G>), smem)         \
    PT_OPTIN((dog_window45_rot<float, G>), smem)
    PT_OPTIN_G(0) PT_OPTIN_G(1) PT_OPTIN_G(2)
#undef PT_OPTIN_G
    PT_OPTIN(dog_rect45_march<uint8_t>, smem)
    PT_OPTIN(dog_rect45_march<float>, smem)
    PT_CLUSTER_OPTINS
#undef PT_OPTIN
    return cudaSuccess;
}

// S < n < 2S with not too many empty slots, frames in HBM, more than one step: rotate the empty slots
// (dog_window45_rot)
static bool uses_rot(const WinArgs &a, const Cfg &cfg, int n)
{
    const int sms = cfg.sms;
    // every step nh = 2S − n windows hop (≈ 2.5 K cycles of hand-off latency each): measured worthwhile up to
    // nh ≈ 0.7·n (n ≥ 1.18·S: 9.3 vs 9.9 µs per step at n = 180, break-even at n = 160); rot = 2 forces it
    const bool few_holes = cfg.rot == 2 || 10 * (2 * sms - n) <= 7 * n;
    return cfg.rot && a.xflag && a.xpos && !a.frame_ptrs && a.T > 1 && n > sms && n < 2 * sms && few_holes;
}

const char *window45_kernel_for(const WinArgs &a, const Cfg &cfg, int n, int pixel)
{
    const int C = cluster_size_for(a, cfg, n);
    if (C == 2) return "dog_window45_cluster<2>";
    if (C == 4) return "dog_window45_cluster<4>";
    if (C == 8) return "dog_window45_cluster<8>";
    return uses_rot(a, cfg, n) ? "dog_window45_rot" : "dog_window45_argmax";
}

cudaError_t launch_window45(const WinArgs &a, const Cfg &cfg, int n, int pixel, cudaStream_t s)
{
    if (!a.h_taps) return cudaErrorInvalidValue;
    Taps45 tp;
    fold_taps(a, tp);
    cudaError_t e;
    Args45 k;
    k.frames = a.frames; k.frame_stride = a.frame_stride; k.step_stride = a.step_stride;
    k.frame_ptrs = a.frame_ptrs;
    k.pitch = a.pitch; k.H = a.H; k.W = a.W; k.fill = a.fill; k.guess = a.guess;
    k.T = a.T > 0 ? a.T : 1;
    k.n = n;
    k.out_pos = a.out_pos; k.out_resp = a.out_resp; k.next_guess = a.next_guess;
    k.traj_pos = a.traj_pos; k.traj_resp = a.traj_resp;
#ifdef PT_PROBES
    k.dbg = g_dbg;
#endif
    k.host_frames = a.host_frames;
    k.rstride = cfg.rot_stride;
    k.rr = a.rr; k.rc = a.rc; k.wr = a.wr; k.wc = a.wc;
    k.f_lo = HW - a.L / 2;                                  // first footprint row that meets a non-zero tap
    k.nfr = a.wr + 2 * (a.L / 2);
    k.ng = (a.wc + RR - 1) / RR;
    k.inv_nfr = (1u << 18) / (unsigned int)k.nfr + 1u;      // (item · inv) >> 18 = item / nfr for every item < 9·nfr + 256
    k.cs = a.wc <= 16 ? 16 : a.wc <= 32 ? 32 : 48;
    k.skew = cfg.rot == 3 ? -1 : cfg.skew;      // rot = 3: dog_window45_rot's handshake settles on the static schedule at once
    k.tm_rows_step = 0; k.tm_rows_frame = 0;
    k.xflag = a.xflag; k.xpos = a.xpos;
    const int C = cluster_size_for(a, cfg, n);
    if (C > 1) {
        e = launch_cluster(k, tp, cfg, C, pixel, s);
        if (e == cudaSuccess) return e;
        cudaGetLastError();                  // cluster launch refused (e.g. MIG slice without clusters): fall through
    }
    const int sms = cfg.sms;
    // one CTA per SM, two windows per CTA; with n ≤ #SMs every window gets its own SM
    const int grid = std::min(sms, n);
    const size_t smem = 2 * HALF_SMEM;
    const int geom = full_geometry(k) ? 0 : (a.L / 2 <= 24 ? 2 : 1);      // see dog_window45_argmax
    if (uses_rot(a, cfg, n)) {
        // plain launch: the CTAs establish co-residency themselves and fall back to the static schedule otherwise
        if (geom == 0) {
            if (pixel == 0) dog_window45_rot<uint8_t, 0><<<grid, CTA_THREADS, smem, s>>>(k, tp);
            else dog_window45_rot<float, 0><<<grid, CTA_THREADS, smem, s>>>(k, tp);
        } else if (geom == 1) {
            if (pixel == 0) dog_window45_rot<uint8_t, 1><<<grid, CTA_THREADS, smem, s>>>(k, tp);
            else dog_window45_rot<float, 1><<<grid, CTA_THREADS, smem, s>>>(k, tp);
        } else {
            if (pixel == 0) dog_window45_rot<uint8_t, 2><<<grid, CTA_THREADS, smem, s>>>(k, tp);
            else dog_window45_rot<float, 2><<<grid, CTA_THREADS, smem, s>>>(k, tp);
        }
        return cudaGetLastError();
    }
    if (geom == 0) {
        if (pixel == 0) dog_window45_argmax<uint8_t, 0><<<grid, CTA_THREADS, smem, s>>>(k, tp);
        else dog_window45_argmax<float, 0><<<grid, CTA_THREADS, smem, s>>>(k, tp);
    } else if (geom == 1) {
        if (pixel == 0) dog_window45_argmax<uint8_t, 1><<<grid, CTA_THREADS, smem, s>>>(k, tp);
        else dog_window45_argmax<float, 1><<<grid, CTA_THREADS, smem, s>>>(k, tp);
    } else {
        if (pixel == 0) dog_window45_argmax<uint8_t, 2><<<grid, CTA_THREADS, smem, s>>>(k, tp);
        else dog_window45_argmax<float, 2><<<grid, CTA_THREADS, smem, s>>>(k, tp);
    }
    return cudaGetLastError();
}

// Decomposition of an output rectangle: how many chunks per strip.  Cost model in units of one marching
// batch on a shared SM (≈ 6.3 µs measured, tools/rect_timing.py): the first batch of an item ≈ 1.75 (it stages
// and row-filters the whole 109-row footprint), later batches 1; an item alone on its SM runs ≈ 1.6× faster than
// a pair.  Measured behaviour of the ticket scheduler: the halves of an SM finish together and fetch together, so
// with few rounds the makespan is ceil(items / 2·SMs) whole items; with many rounds it tends to
// total work / (2·SMs) plus half an item of tail.
int rect45_pick_chunks(int n, int ntx, int nty, int sms, int forced)
{
    if (forced >= 1) return std::min(forced, nty);
    int best_c = nty;
    double best_cost = 1e300;
    for (int c = 1; c <= nty; ++c) {
        const long long items = (long long)n * ntx * c;
        const int nbmax = (nty + c - 1) / c;
        const double item_cost = 1.75 + (nbmax - 1);
        const double work = (double)n * ntx * (1.75 * c + (nty - c));
        const double rounds = (double)items / (2.0 * sms);
        double cost;
        if (items <= sms) cost = item_cost / 1.6;
        else if (rounds <= 4.0) cost = (double)((items + 2LL * sms - 1) / (2LL * sms)) * item_cost;
        else cost = work / (2.0 * sms) + 0.5 * item_cost;
        if (cost < best_cost - 1e-9) { best_cost = cost; best_c = c; }
    }
    return best_c;
}

cudaError_t launch_rect45(const WinArgs &a, const Cfg &cfg, int n, int pixel, cudaStream_t s)
{
    if (!a.h_taps) return cudaErrorInvalidValue;
    Taps45 tp;
    fold_taps(a, tp);
    const int sms = cfg.sms;
    March45 m;
    m.n = n;
    m.nty = (a.wr + WR - 1) / WR;
    m.ntx = (a.wc + WC - 1) / WC;
    m.nchunks = rect45_pick_chunks(n, m.ntx, m.nty, sms, cfg.r45_chunks);
    const long long items = (long long)n * m.ntx * m.nchunks;
    // one CTA per SM; with fewer items than SMs every item gets an SM to itself (second halves stay idle)
    const int grid = (int)std::max<long long>(1, std::min<long long>(sms, items));
    const size_t smem = 2 * HALF_SMEM;
    if (pixel == 0) dog_rect45_march<uint8_t><<<grid, CTA_THREADS, smem, s>>>(a, tp, m);
    else dog_rect45_march<float><<<grid, CTA_THREADS, smem, s>>>(a, tp, m);
    return cudaGetLastError();
}

} // namespace pt
