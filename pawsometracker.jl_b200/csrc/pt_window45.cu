// pt_window45.cu — the batched hot path at the reference's default geometry:
// target_width = 25 → l = 65 (w = 32), window 45×45 (radii 22), footprint
// 109×109.  One CTA per (video, window); a launch can chain T time steps of the
// same videos (frames resident in HBM) because windows of different videos
// never interact: ij[t] = trckr(ij[t-1]) (src/PawsomeTracker.jl:167) stays
// inside the CTA, so there is no grid-wide dependency between steps.
//
// Per frame and CTA:
//   stage  109×109 pixels → smem as (pixel − fill), 0 outside the frame
//          (the PaddedView border, src/PawsomeTracker.jl:48, after subtracting
//          the constant fill — legal because ΣDoG = 0)
//   row    both Gaussians for 109 rows × 45 columns.  The factors are
//          symmetric, so each output is g0·x0 + Σ_d g_d·(x_-d + x_+d): one FADD
//          feeds ONE packed FFMA2 that advances (narrow, wide) together.
//          Thread = (row, 9 consecutive columns), 73 shared loads per 585 ops.
//   col    45×45 outputs, thread = (column, 9 consecutive rows); one FFMA2
//          advances two vertically adjacent outputs; subtraction and the
//          darker_target sign are folded into the column taps.
//   argmax warp shuffles + 9 keys in smem that every thread folds itself: first
//          maximum in column-major order (findmax, :59); clamp (:61).
//
// All taps are kernel parameters (constant bank → uniform registers), the loops
// are fully unrolled, so no tap is ever loaded inside the passes.
#include "pt_kernels.cuh"

#include <cstdlib>

namespace pt {

namespace {

constexpr int L = 65, HW = 32;          // kernel length / half width
constexpr int WR = 45, WC = 45;         // window outputs
constexpr int FR = WR + 2 * HW;         // 109 footprint rows
constexpr int FC = WC + 2 * HW;         // 109 footprint cols
constexpr int PIN = 109;                // s_in pitch (floats), odd → row-lanes conflict-free
constexpr int PM = 45;                  // s_mid pitch (float2), odd → 64-bit row-lane stores conflict-free
constexpr int R = 9;                    // outputs per thread along the filter direction
constexpr int NG = 5;                   // groups of R per 45
constexpr int ROW_ITEMS = FR * NG;      // 545
constexpr int COL_ITEMS = WC * NG;      // 225
constexpr int THREADS = 256;            // 8 warps = 2 per SM sub-partition; row items take 3 rounds, column items 1
constexpr int NWARPS = THREADS / 32;

// Taps as kernel parameters, laid out for packed FP32 (fma.rn.f32x2 → FFMA2):
//   rt[d]  = (narrow, wide) folded row taps, d = |k − 32| (pixel scale folded in)
//   cpp[q] = (cp[q], cp[q−1]), cmq[q] = (cm[q], cm[q−1]) column taps (sign folded in),
//            zero outside 0..64: one FFMA2 advances two vertically adjacent outputs.
struct Taps45 {
    float2 rt[HW + 1];
    float2 cpp[L + 1], cmq[L + 1];
};

struct Args45 {
    const void *frames;                 // frame of window 0 at step 0
    size_t frame_stride, step_stride;   // elements
    const void *const *frame_ptrs;      // optional [T][n] frame pointers (zero-copy pinned host frames)
    int pitch, H, W;
    const float *fill;
    const int2 *guess;                  // [n] start guess (1-based)
    int T;
    int skew_cycles;                    // second-wave CTAs (blockIdx ≥ #SMs) start this many cycles late
    int num_sms;
    int warp_rot;                       // logical-warp rotation of second-wave CTAs (sub-partition balance)
    int4 *out_pos; float *out_resp;     // [n] last step
    int2 *next_guess;                   // [n] or null
    int4 *traj_pos; float *traj_resp;   // [T][n] or null
    int n;
    long long *dbg;                     // optional [n][T][6]: smid, t0, after stage, after row, after col, end (clock64)
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

// ---- staging: footprint → smem as (pixel − fill), 0 outside the frame ----------------
// Warp w takes rows w, w+9, …; every load of a thread is issued before the first
// conversion so a frame that is cold in L2/HBM costs one memory round trip.
template <typename PixT>
__device__ __forceinline__ void stage_tile(const PixT *frame, int pitch, int H, int W, int fy0, int fx0,
                                           float fill, float *s_in, int warp, int lane);

// f32 frames: lane = column (+32q), coalesced 4-byte loads.
template <>
__device__ __forceinline__ void stage_tile<float>(const float *frame, int pitch, int H, int W, int fy0, int fx0,
                                                  float fill, float *s_in, int warp, int lane)
{
    constexpr int RPW = (FR + NWARPS - 1) / NWARPS;      // 13 rows per warp (last ones masked)
    float px[RPW][4];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int f = warp + r * NWARPS;
        const int Y = fy0 + f;
        const bool yok = (f < FR) && (Y >= 0) && (Y < H);
        const float *rowp = frame + (size_t)(yok ? Y : 0) * pitch;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            const int X = fx0 + c;
            px[r][q] = (yok && c < FC && X >= 0 && X < W) ? __ldg(rowp + X) : fill;
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int f = warp + r * NWARPS;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            if (f < FR && c < FC) s_in[f * PIN + c] = px[r][q] - fill;
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
template <>
__device__ __forceinline__ void stage_tile<uint8_t>(const uint8_t *frame, int pitch, int H, int W, int fy0, int fx0,
                                                    float fill, float *s_in, int warp, int lane)
{
    constexpr int RPW = (FR + NWARPS - 1) / NWARPS;
    const int xa = fx0 & ~3, phase = fx0 - xa;             // aligned start, phase 0..3
    const int X = xa + 4 * lane;                            // frame column of this lane's word
    const unsigned int fillw = (unsigned int)fill * 0x01010101u;
    unsigned int keep = 0u;                                 // bytes of the word inside [0, W)
#pragma unroll
    for (int b = 0; b < 4; ++b) if (X + b >= 0 && X + b < W) keep |= 0xFFu << (8 * b);
    const bool wordok = keep != 0u && (4 * lane - phase < FC);
    const int rot = lane >> 3;
    int col[4];
    bool cok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        col[k] = 4 * lane - phase + ((k + rot) & 3);
        cok[k] = col[k] >= 0 && col[k] < FC;
    }
    unsigned int wd[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int f = warp + r * NWARPS;
        const int Y = fy0 + f;
        const bool ok = wordok && (f < FR) && (Y >= 0) && (Y < H);
        wd[r] = fillw;
        if (ok) wd[r] = __ldg(reinterpret_cast<const unsigned int *>(frame + (size_t)Y * pitch + X));
    }
    const float cst = 8388608.0f + fill;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int f = warp + r * NWARPS;
        if (f < FR) {
            unsigned int w = (wd[r] & keep) | (fillw & ~keep);
            w = __funnelshift_r(w, w, 8 * rot);            // byte k of w = pixel (k + rot) & 3 of the word
            float *dst = s_in + f * PIN;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float val = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + k)) - cst;
                if (cok[k]) dst[col[k]] = val;
            }
        }
    }
}

} // namespace

template <typename PixT>
__global__ void __launch_bounds__(THREADS, 2)
dog_window45_argmax(const __grid_constant__ Args45 a, const __grid_constant__ Taps45 tp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_in = reinterpret_cast<float *>(smem_raw);                       // [FR][PIN]
    float2 *s_mid = reinterpret_cast<float2 *>(s_in + FR * PIN + 1);        // [FR][PM] (8-byte aligned: FR*PIN+1 is even)
    __shared__ unsigned long long s_key[2 * NWARPS];

    const int lane = threadIdx.x & 31;
    // Logical warp id.  Row items need 18 warp-tasks on 8 warps, so logical warps 0 and 1 (sub-
    // partitions 0 and 1) carry 3 tasks and the others 2.  The second CTA that lands on an SM
    // (blockIdx ≥ #SMs with the breadth-first block scheduler) rotates its mapping by two
    // sub-partitions so the pair loads all four FMA pipes equally (5+4 tasks each).
    const int warp = ((threadIdx.x >> 5) + (((int)blockIdx.x >= a.num_sms) ? a.warp_rot : 0)) & (NWARPS - 1);
    const int tid = warp * 32 + lane;
    const int v = blockIdx.x;
    const float fill = a.fill[v];
    int2 g = a.guess[v];

    // Two CTAs share an SM and would otherwise run their phases in lockstep (both staging,
    // then both fighting for the FMA pipe).  Starting the second-wave CTA part of a frame
    // late makes one CTA's staging/reduce overlap the other's FMA passes.
    if (a.skew_cycles > 0 && (int)blockIdx.x >= a.num_sms && a.T > 1) {
        const long long t0 = clock64();
        while (clock64() - t0 < a.skew_cycles) { }
    }

    for (int t = 0; t < a.T; ++t) {
        long long *dbg = a.dbg ? a.dbg + ((size_t)v * a.T + t) * 6 : nullptr;
        if (dbg && tid == 0) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            dbg[0] = (t == 0) ? (long long)smid : (long long)gt;   // t>0: wall-clock ns at frame start
            dbg[1] = clock64();
        }
        const PixT *frame = a.frame_ptrs
            ? reinterpret_cast<const PixT *>(a.frame_ptrs[(size_t)t * a.n + v])
            : reinterpret_cast<const PixT *>(a.frames) + (size_t)t * a.step_stride + (size_t)v * a.frame_stride;
        const int wy0 = g.x - 1 - (WR / 2), wx0 = g.y - 1 - (WC / 2);   // window origin, 0-based
        const int fy0 = wy0 - HW, fx0 = wx0 - HW;                        // footprint origin

        stage_tile<PixT>(frame, a.pitch, a.H, a.W, fy0, fx0, fill, s_in, warp, lane);

        // ---- warm L2 with everything the NEXT step can touch: its window centre is inside this
        // step's window, so its footprint lies within ±22 px of this one (153 rows × ≤3 lines).
        if (t + 1 < a.T && !a.frame_ptrs) {                      // (host frames are not cached in L2)
            const PixT *nframe = frame + a.step_stride;
            constexpr int PR = FR + WR - 1;                      // 153 rows
            const int py0 = fy0 - WR / 2, pxb = (fx0 - WC / 2) * (int)sizeof(PixT);
            const int line0 = pxb >> 7, nlines = ((pxb + (FC + WC - 1) * (int)sizeof(PixT) - 1) >> 7) - line0 + 1;
            for (int e = tid; e < PR * nlines; e += THREADS) {
                const int r = e / nlines, ln = e - r * nlines;
                const int Y = py0 + r;
                const long long off = ((long long)(line0 + ln)) << 7;
                if (Y >= 0 && Y < a.H && off >= 0 && off < (long long)a.W * (int)sizeof(PixT)) {
                    const char *ptr = reinterpret_cast<const char *>(nframe + (size_t)Y * a.pitch) + off;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
                }
            }
        }
        __syncthreads();
        if (dbg && tid == 0) dbg[2] = clock64();

        // ---- row pass: item = (footprint row f, column group gq); lanes walk rows
#pragma unroll 1
        for (int item = tid; item < ROW_ITEMS; item += THREADS) {
            const int gq = item / FR, f = item - gq * FR;
            const float *row = s_in + f * PIN + gq * R;
            float x[R + 2 * HW];
#pragma unroll
            for (int i = 0; i < R + 2 * HW; ++i) x[i] = row[i];
            float2 acc[R];                                   // (narrow, wide) per output
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = fmul2(make_float2(x[j + HW], x[j + HW]), tp.rt[0]);
#pragma unroll
            for (int d = 1; d <= HW; ++d) {
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const float s = x[j + HW - d] + x[j + HW + d];   // exact for u8 frames (integers)
                    acc[j] = ffma2(make_float2(s, s), tp.rt[d], acc[j]);
                }
            }
            float2 *dst = s_mid + f * PM + gq * R;
#pragma unroll
            for (int j = 0; j < R; ++j) dst[j] = acc[j];
        }
        __syncthreads();
        if (dbg && tid == 0) dbg[3] = clock64();

        // ---- column pass + running argmax: item = (column xq, row group h); lanes walk columns
        unsigned long long key = 0ull;
        if (tid < COL_ITEMS) {
            const int h = tid / WC, xq = tid - h * WC;
            const float2 *col = s_mid + (h * R) * PM + xq;
            // outputs (2p, 2p+1) share one packed accumulator; output R-1 = 8 stays scalar
            float2 acc2[R / 2];
            float acc8 = 0.f;
#pragma unroll
            for (int p = 0; p < R / 2; ++p) acc2[p] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < R + 2 * HW; ++i) {
                const float2 m = col[i * PM];
#pragma unroll
                for (int p = 0; p < R / 2; ++p) {
                    const int q = i - 2 * p;                 // tap of the even output; the odd one uses q-1
                    if (q >= 0 && q <= L) {
                        acc2[p] = ffma2(make_float2(m.x, m.x), tp.cpp[q], acc2[p]);
                        acc2[p] = ffma2(make_float2(m.y, m.y), tp.cmq[q], acc2[p]);
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
                if (val > bv) { bv = val; bj = j; }       // strict: first maximum within the column segment
            }
            key = pack_key(bv, (unsigned int)(xq * WR + h * R + bj));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, key, off);
            key = o > key ? o : key;
        }
        if (lane == 0) s_key[(t & 1) * NWARPS + warp] = key;
        __syncthreads();
        if (dbg && tid == 0) dbg[4] = clock64();
        {
            // every thread folds the 9 warp keys itself (broadcast loads): no serial section and
            // no second barrier; s_key is double-buffered by step parity
            const unsigned long long *kk = s_key + (t & 1) * NWARPS;
            unsigned long long k = kk[0];
#pragma unroll
            for (int i = 1; i < NWARPS; ++i) k = kk[i] > k ? kk[i] : k;
            const unsigned int idx = key_index(k);
            const int xx = (int)(idx / WR), yy = (int)(idx - xx * WR);
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
            }
            g = make_int2(ci, cj);
        }
        if (dbg && tid == 0) dbg[5] = clock64();
    }
}

const char *window45_name() { return "dog_window45_argmax"; }

static long long *g_dbg = nullptr;
void window45_set_debug(long long *dev_buf) { g_dbg = dev_buf; }

bool window45_supported(const WinArgs &a, int pixel)
{
    if (!(a.L == L && a.wr == WR && a.wc == WC && !a.rect_mode && a.map_out == nullptr)) return false;
    if (pixel == 0 && a.frame_ptrs) return (a.pitch & 3) == 0 && a.pitch >= ((a.W + 3) & ~3);   // caller checked the pointers
    if (pixel == 0 && a.frames) {
        // the u8 staging path loads aligned 32-bit words
        const bool aligned = ((reinterpret_cast<uintptr_t>(a.frames) | (uintptr_t)a.pitch | (uintptr_t)a.frame_stride |
                               (uintptr_t)a.step_stride) & 3u) == 0;
        if (!aligned || a.pitch < ((a.W + 3) & ~3)) return false;
    }
    return true;
}

// The specialised kernel takes its taps as kernel parameters (constant bank):
// fold the symmetric row factors (index d = |k − 32|) and pair the column taps.
static void fold_taps(const WinArgs &a, Taps45 &tp)
{
    const float *rp = a.h_taps, *rm = a.h_taps + L, *cp = a.h_taps + 2 * L, *cm = a.h_taps + 3 * L;
    auto at = [](const float *t, int k) { return (k >= 0 && k < L) ? t[k] : 0.f; };
    for (int d = 0; d <= HW; ++d) tp.rt[d] = make_float2(rp[HW + d], rm[HW + d]);
    for (int q = 0; q <= L; ++q) {
        tp.cpp[q] = make_float2(at(cp, q), at(cp, q - 1));
        tp.cmq[q] = make_float2(at(cm, q), at(cm, q - 1));
    }
}

cudaError_t launch_window45(const WinArgs &a, int n, int pixel, cudaStream_t s)
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
    {
        static int sms = 0, skew = 0, rot = 2;
        if (sms == 0) {
            if (const char *r = getenv("PT_W45_ROT")) rot = atoi(r);
            int dev = 0;
            cudaGetDevice(&dev);
            int v = 0;
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
            const char *env = getenv("PT_W45_SKEW");
            skew = env ? atoi(env) : 0;
            sms = v > 0 ? v : 148;
        }
        k.num_sms = sms; k.skew_cycles = skew; k.warp_rot = rot;
    }
    k.out_pos = a.out_pos; k.out_resp = a.out_resp; k.next_guess = a.next_guess;
    k.traj_pos = a.traj_pos; k.traj_resp = a.traj_resp; k.n = n;
    k.dbg = g_dbg;
    const size_t smem = (size_t)(FR * PIN + 1) * sizeof(float) + (size_t)FR * PM * sizeof(float2);
    if (pixel == 0) {
        e = cudaFuncSetAttribute(dog_window45_argmax<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        dog_window45_argmax<uint8_t><<<n, THREADS, smem, s>>>(k, tp);
    } else {
        e = cudaFuncSetAttribute(dog_window45_argmax<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        dog_window45_argmax<float><<<n, THREADS, smem, s>>>(k, tp);
    }
    return cudaGetLastError();
}

} // namespace pt
