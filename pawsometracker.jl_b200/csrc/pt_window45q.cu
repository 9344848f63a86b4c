// pt_window45q.cu — second-generation batched kernel for the default geometry
// (target_width 25 → l = 65, window 45×45, u8 frames): FOUR CTAs per window.
//
// Why: with one CTA per window, 256 windows give 148 SMs only one or two CTAs each
// (108×2 + 40×1): phases serialise behind barriers and the load is 86.5 % balanced.
// Here each window is cut into four 12-column output strips → 1024 small CTAs, seven
// co-resident per SM (28 KB smem, 128 threads, ≤72 registers): 1024/148 = 6.9 → 98.8 %
// balance, and the stage/row/col phases of seven different windows overlap on every SM.
//
// The serial chain ij[t] = trckr(ij[t-1]) (src/PawsomeTracker.jl:167) still runs inside
// one launch: after the column pass the four CTAs of a window combine their argmax keys
// through one 64-bit atomicMax + arrival counter per (video, step) in global memory and
// spin (nanosleep) until all four have arrived.  That requires all 4n CTAs to be
// co-resident, which the cooperative launch guarantees (else the launcher declines and
// the one-CTA-per-window kernel is used).
//
// Per frame and CTA (strip of 12 output columns, all 45 output rows):
//   stage  109 rows × 76 columns of the footprint → smem as bf16 pairs of (pixel − fill):
//          integers in [−255, 255] are exact in bf16 (8 significant bits), so the tile costs
//          half the f32 bytes with no rounding at all.  Aligned 32-bit loads (4 px), 2²³-trick
//          conversion (PRMT + the FADD that subtracts the fill), pairs packed with one PRMT.
//          Strip origins are chosen per frame so the first tile column falls on an even
//          frame column ({0,12,24,34} or {−1,11,23,33}; columns outside 0..44 are masked),
//          which keeps byte pairs aligned with bf16 pairs.
//   row    thread = (footprint row, 6 output columns): 35 LDS.32 → 70 inputs (SHL / LOP3 on the
//          ALU pipe), symmetric fold: 32 FADD + 33 packed FFMA2 per output (narrow, wide).
//   col    thread = (output column, 9 rows), one FFMA2 advances two vertically adjacent outputs.
//   argmax warp shuffle → 4 keys → global exchange among the 4 CTAs → decode, clamp.
#include "pt_kernels.cuh"

#include <cstdlib>

namespace pt {

namespace {

constexpr int L = 65, HW = 32;
constexpr int WR = 45, WC = 45;
constexpr int FR = WR + 2 * HW;          // 109 footprint rows
constexpr int NQ = 4;                    // CTAs (strips) per window
constexpr int CS = 12;                   // output columns per strip
constexpr int TC = CS + 2 * HW;          // 76 tile columns
constexpr int PW = 39;                   // tile pitch in 32-bit words (bf16 pairs); odd → row-lanes conflict-free
constexpr int PMQ = 13;                  // s_mid pitch (float2); odd
constexpr int RQ = 6;                    // row-pass outputs per item
constexpr int ROW_ITEMS = FR * (CS / RQ);   // 218
constexpr int R = 9, NGR = 5;            // column pass: 9 rows per item, 5 groups
constexpr int COL_ITEMS = CS * NGR;      // 60
constexpr int THREADS = 128;
constexpr int NW = THREADS / 32;
constexpr int WPR = 20;                  // staged words per row: 76 columns + up to 2 phase bytes → 78 B → 20 words
constexpr int STAGE_ITEMS = FR * WPR;    // 2180
constexpr int STAGE_IT = (STAGE_ITEMS + THREADS - 1) / THREADS;   // 18

struct TapsQ {
    float2 rt[HW + 1];                   // (narrow, wide) folded row taps, d = |k − 32|
    float2 cpp[L + 1], cmq[L + 1];       // column tap pairs (c[q], c[q−1]), zero outside 0..64
};

struct ArgsQ {
    const uint8_t *frames;               // frame of window 0 at step 0
    size_t frame_stride, step_stride;
    const void *const *frame_ptrs;       // optional [T][n] frame pointers (zero-copy pinned host frames)
    int pitch, H, W;
    const float *fill;
    const int2 *guess;
    int T, n;
    unsigned long long *xkeys;           // [n][T] combined argmax keys (zeroed before the launch)
    unsigned int *xcnt;                  // [n][T] arrival counters  (zeroed before the launch)
    int4 *out_pos; float *out_resp;
    int2 *next_guess;
    int4 *traj_pos; float *traj_resp;
    long long *dbg;                      // optional [4n][T][6]: smid|globaltimer, t0, after stage, row, col, exchange
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Stage the strip's tile.  X0 = frame column of tile column 0 (even by construction of the
// strip origin), so phase = X0 mod 4 ∈ {0, 2} and byte pairs coincide with bf16 pairs.
template <bool kInterior>
__device__ __forceinline__ void stage_strip(const uint8_t *frame, int pitch, int H, int W, int fy0, int X0,
                                            float fill, unsigned int *s_tile, int tid)
{
    const int xa = X0 & ~3, hp = (X0 - xa) >> 1;           // aligned start; half-phase 0 or 1 (in bf16 pairs)
    const unsigned int fillw = (unsigned int)fill * 0x01010101u;
    unsigned int wd[STAGE_IT];
#pragma unroll
    for (int i = 0; i < STAGE_IT; ++i) {
        const int e = tid + THREADS * i;
        const int row = e / WPR, k = e - row * WPR;
        const int Y = fy0 + row, X = xa + 4 * k;
        unsigned int w = fillw;
        if (kInterior) {
            if (e < STAGE_ITEMS) w = __ldg(reinterpret_cast<const unsigned int *>(frame + (size_t)Y * pitch + X));
        } else {
            const bool ok = (e < STAGE_ITEMS) && (Y >= 0) && (Y < H) && (X + 3 >= 0) && (X < W);
            if (ok) {
                w = __ldg(reinterpret_cast<const unsigned int *>(frame + (size_t)Y * pitch + X));
                unsigned int keep = 0u;
#pragma unroll
                for (int b = 0; b < 4; ++b) if (X + b >= 0 && X + b < W) keep |= 0xFFu << (8 * b);
                w = (w & keep) | (fillw & ~keep);
            }
        }
        wd[i] = w;
    }
    const float cst = 8388608.0f + fill;
#pragma unroll
    for (int i = 0; i < STAGE_IT; ++i) {
        const int e = tid + THREADS * i;
        const int row = e / WPR, k = e - row * WPR;
        const unsigned int w = wd[i];
        const float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540)) - cst;
        const float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7541)) - cst;
        const float f2 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7542)) - cst;
        const float f3 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7543)) - cst;
        // bf16 = upper half of the f32 (exact: |value| ≤ 255 has ≤ 8 significant bits)
        const unsigned int p01 = __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632);
        const unsigned int p23 = __byte_perm(__float_as_uint(f2), __float_as_uint(f3), 0x7632);
        const int i0 = 2 * k - hp;                         // tile word of bytes (0,1); bytes (2,3) → i0 + 1
        if (e < STAGE_ITEMS) {
            unsigned int *dst = s_tile + row * PW;
            if (i0 >= 0 && i0 < PW - 1) dst[i0] = p01;
            if (i0 + 1 >= 0 && i0 + 1 < PW - 1) dst[i0 + 1] = p23;
        }
    }
}

} // namespace

__global__ void __launch_bounds__(THREADS, 7)
dog_window45_quad(const __grid_constant__ ArgsQ a, const __grid_constant__ TapsQ tp)
{
    __shared__ __align__(16) unsigned int s_tile[FR * PW];            // bf16 pairs, 17,004 B
    __shared__ __align__(16) float2 s_mid[FR * PMQ];                  // 11,336 B
    __shared__ unsigned long long s_key[NW];
    __shared__ unsigned long long s_final;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v = blockIdx.x >> 2, q = blockIdx.x & 3;
    const float fill = a.fill[v];
    int2 g = a.guess[v];

    for (int t = 0; t < a.T; ++t) {
        long long *dbg = a.dbg ? a.dbg + ((size_t)blockIdx.x * a.T + t) * 6 : nullptr;
        if (dbg && tid == 0) {
            unsigned int smid;
            unsigned long long gt;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            dbg[0] = (t == 0) ? (long long)smid : (long long)gt;
            dbg[1] = clock64();
        }
        const uint8_t *frame = a.frame_ptrs ? reinterpret_cast<const uint8_t *>(a.frame_ptrs[(size_t)t * a.n + v])
                                            : a.frames + (size_t)t * a.step_stride + (size_t)v * a.frame_stride;
        const int wy0 = g.x - 1 - (WR / 2), wx0 = g.y - 1 - (WC / 2);   // window origin, 0-based
        const int fy0 = wy0 - HW, fx0 = wx0 - HW;                        // footprint origin
        // strip origin: the first tile column must fall on an even frame column, so the strip
        // starts are {0,12,24,34} when fx0 is even and {−1,11,23,33} when it is odd (both cover
        // window columns 0..44; columns −1 / 45 are masked out of the argmax)
        const int cs = (fx0 & 1) ? (q < 3 ? 12 * q - 1 : 33) : (q < 3 ? 12 * q : 34);
        const int X0 = fx0 + cs;

        const bool interior = (fy0 >= 0) && (fy0 + FR <= a.H) && ((X0 & ~3) >= 0) && ((X0 & ~3) + 4 * WPR <= a.W);
        if (interior) stage_strip<true>(frame, a.pitch, a.H, a.W, fy0, X0, fill, s_tile, tid);
        else stage_strip<false>(frame, a.pitch, a.H, a.W, fy0, X0, fill, s_tile, tid);

        // warm L2 with what the next step can touch (its window centre lies inside this window)
        if (t + 1 < a.T && !a.frame_ptrs) {
            const uint8_t *nframe = frame + a.step_stride;
            constexpr int PR = FR + WR - 1;                      // 153 rows
            const int py0 = fy0 - WR / 2, pxb = X0 - WC / 2 - 1;
            const int line0 = pxb >> 7, nlines = ((pxb + TC + WC + 1) >> 7) - line0 + 1;
            for (int e = tid; e < PR * nlines; e += THREADS) {
                const int r = e / nlines, ln = e - r * nlines;
                const int Y = py0 + r;
                const long long off = ((long long)(line0 + ln)) << 7;
                if (Y >= 0 && Y < a.H && off >= 0 && off < (long long)a.W)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nframe + (size_t)Y * a.pitch + off));
            }
        }
        __syncthreads();
        if (dbg && tid == 0) dbg[2] = clock64();

        // ---- row pass: item = (footprint row f, group gq of 6 output columns); lanes walk rows
#pragma unroll 1
        for (int item = tid; item < ROW_ITEMS; item += THREADS) {
            const int gq = item / FR, f = item - gq * FR;
            const unsigned int *row = s_tile + f * PW + gq * (RQ / 2);
            float x[RQ + 2 * HW];                            // 70 inputs from 35 words
#pragma unroll
            for (int i = 0; i < (RQ + 2 * HW) / 2; ++i) {
                const unsigned int w = row[i];
                x[2 * i] = __uint_as_float(w << 16);
                x[2 * i + 1] = __uint_as_float(w & 0xFFFF0000u);
            }
            float2 acc[RQ];
#pragma unroll
            for (int j = 0; j < RQ; ++j) acc[j] = fmul2(make_float2(x[j + HW], x[j + HW]), tp.rt[0]);
#pragma unroll
            for (int d = 1; d <= HW; ++d) {
#pragma unroll
                for (int j = 0; j < RQ; ++j) {
                    const float s = x[j + HW - d] + x[j + HW + d];   // exact (small integers)
                    acc[j] = ffma2(make_float2(s, s), tp.rt[d], acc[j]);
                }
            }
            float2 *dst = s_mid + f * PMQ + gq * RQ;
#pragma unroll
            for (int j = 0; j < RQ; ++j) dst[j] = acc[j];
        }
        __syncthreads();
        if (dbg && tid == 0) dbg[3] = clock64();

        // ---- column pass: item = (strip column xq, row group h); 60 items on warps 0-1
        unsigned long long key = 0ull;
        if (tid < COL_ITEMS) {
            const int h = tid / CS, xq = tid - h * CS;
            const float2 *col = s_mid + (h * R) * PMQ + xq;
            float2 acc2[R / 2];
            float acc8 = 0.f;
#pragma unroll
            for (int p = 0; p < R / 2; ++p) acc2[p] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < R + 2 * HW; ++i) {
                const float2 m = col[i * PMQ];
#pragma unroll
                for (int p = 0; p < R / 2; ++p) {
                    const int qq = i - 2 * p;
                    if (qq >= 0 && qq <= L) {
                        acc2[p] = ffma2(make_float2(m.x, m.x), tp.cpp[qq], acc2[p]);
                        acc2[p] = ffma2(make_float2(m.y, m.y), tp.cmq[qq], acc2[p]);
                    }
                }
                const int q8 = i - (R - 1);
                if (q8 >= 0 && q8 < L) {
                    acc8 = fmaf(m.x, tp.cpp[q8].x, acc8);
                    acc8 = fmaf(m.y, tp.cmq[q8].x, acc8);
                }
            }
            float acc[R];
#pragma unroll
            for (int p = 0; p < R / 2; ++p) { acc[2 * p] = acc2[p].x; acc[2 * p + 1] = acc2[p].y; }
            acc[R - 1] = acc8;
            float bv = acc[0] + 0.0f;
            int bj = 0;
#pragma unroll
            for (int j = 1; j < R; ++j) {
                const float val = acc[j] + 0.0f;
                if (val > bv) { bv = val; bj = j; }
            }
            const int wcol = cs + xq;                            // window column of this output
            if (wcol >= 0 && wcol < WC) key = pack_key(bv, (unsigned int)(wcol * WR + h * R + bj));
        }
        if (warp < 2) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, key, off);
                key = o > key ? o : key;
            }
            if (lane == 0) s_key[warp] = key;
        }
        __syncthreads();
        if (tid == 0) {
            if (dbg) dbg[4] = clock64();
            const unsigned long long k = s_key[0] > s_key[1] ? s_key[0] : s_key[1];
            const size_t slot = (size_t)v * a.T + t;
            atomicMax(a.xkeys + slot, k);
            __threadfence();
            atomicAdd(a.xcnt + slot, 1u);
            while (ld_acquire_u32(a.xcnt + slot) < (unsigned int)NQ) __nanosleep(64);
            s_final = ld_relaxed_u64(a.xkeys + slot);
            if (dbg) dbg[5] = clock64();
        }
        __syncthreads();
        {
            const unsigned long long k = s_final;
            const unsigned int idx = key_index(k);
            const int xx = (int)(idx / WR), yy = (int)(idx - xx * WR);
            const int raw_i = wy0 + yy + 1, raw_j = wx0 + xx + 1;                   // absolute index (:60)
            const int ci = min(max(raw_i, 1), a.H), cj = min(max(raw_j, 1), a.W);   // clamp (:61)
            if (tid == 0 && q == 0) {
                const float resp = key_value(k);
                const int4 p = make_int4(ci, cj, raw_i, raw_j);
                if (a.traj_pos) { a.traj_pos[(size_t)t * a.n + v] = p; a.traj_resp[(size_t)t * a.n + v] = resp; }
                if (t == a.T - 1) {
                    a.out_pos[v] = p; a.out_resp[v] = resp;
                    if (a.next_guess) a.next_guess[v] = make_int2(ci, cj);
                }
            }
            g = make_int2(ci, cj);
        }
        // s_final is rewritten only after the next frame's three barriers: no hazard
    }
}

const char *window45_quad_name() { return "dog_window45_quad"; }

// Largest batch whose 4n CTAs are all co-resident on the current device (0 = unknown).
int window45_quad_max_windows()
{
    static int cached = -1;
    if (cached >= 0) return cached;
    int dev = 0, sms = 0, per_sm = 0, coop = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dog_window45_quad, THREADS, 0) != cudaSuccess) {
        cudaGetLastError();
        cached = 0;
        return 0;
    }
    cached = (sms * per_sm) / NQ;
    return cached;
}

bool window45_quad_supported(const WinArgs &a, int n, int pixel)
{
    if (pixel != 0 || !getenv("PT_ENABLE_QUAD")) return false;   // experimental: measured slower than dog_window45_argmax
    if (!(a.L == L && a.wr == WR && a.wc == WC && !a.rect_mode && a.map_out == nullptr)) return false;
    if (!a.xkeys || !a.xcnt) return false;
    if ((a.pitch & 3) != 0 || a.pitch < ((a.W + 3) & ~3)) return false;
    if (!a.frame_ptrs) {
        if (((reinterpret_cast<uintptr_t>(a.frames) | (uintptr_t)a.frame_stride | (uintptr_t)a.step_stride) & 3u) != 0) return false;
    }
    return n <= window45_quad_max_windows();
}

cudaError_t launch_window45_quad(const WinArgs &a, int n, cudaStream_t s)
{
    if (!a.h_taps) return cudaErrorInvalidValue;
    TapsQ tp;
    {
        const float *rp = a.h_taps, *rm = a.h_taps + L, *cp = a.h_taps + 2 * L, *cm = a.h_taps + 3 * L;
        auto at = [](const float *t, int k) { return (k >= 0 && k < L) ? t[k] : 0.f; };
        for (int d = 0; d <= HW; ++d) tp.rt[d] = make_float2(rp[HW + d], rm[HW + d]);
        for (int qd = 0; qd <= L; ++qd) {
            tp.cpp[qd] = make_float2(at(cp, qd), at(cp, qd - 1));
            tp.cmq[qd] = make_float2(at(cm, qd), at(cm, qd - 1));
        }
    }
    ArgsQ k;
    k.frames = (const uint8_t *)a.frames; k.frame_stride = a.frame_stride; k.step_stride = a.step_stride;
    k.frame_ptrs = a.frame_ptrs;
    k.pitch = a.pitch; k.H = a.H; k.W = a.W; k.fill = a.fill; k.guess = a.guess;
    k.T = a.T > 0 ? a.T : 1; k.n = n;
    k.xkeys = a.xkeys; k.xcnt = a.xcnt;
    k.out_pos = a.out_pos; k.out_resp = a.out_resp; k.next_guess = a.next_guess;
    k.traj_pos = a.traj_pos; k.traj_resp = a.traj_resp;
    k.dbg = window45_debug_ptr();
    void *params[2] = {(void *)&k, (void *)&tp};
    return cudaLaunchCooperativeKernel((const void *)dog_window45_quad, dim3((unsigned)(NQ * n)), dim3(THREADS), params, 0, s);
}

} // namespace pt
