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
//          feeds two FFMAs (narrow, wide) — 98 instead of 130 FP32 ops/output.
//          Thread = (row, 9 consecutive columns), 73 shared loads per 882 ops.
//   col    45×45 outputs, thread = (column, 9 consecutive rows); subtraction
//          and darker_target sign are folded into the column taps; running
//          argmax in registers.
//   argmax warp shuffles + 9-entry smem reduce, first maximum in column-major
//          order (findmax, :59); clamp (:61); next guess stays in smem.
//
// All taps are kernel parameters (constant bank), the loops are fully unrolled,
// so every FFMA takes its tap as a constant operand: no tap loads at all.
#include "pt_kernels.cuh"

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
constexpr int THREADS = 288;            // 9 warps: 2 rounds of row items, 1 round of column items
constexpr int NWARPS = THREADS / 32;

struct Taps45 {
    float rp[HW + 1], rm[HW + 1];       // folded row taps: index d = |k - 32| (pixel scale folded in)
    float cp[L], cm[L];                 // column taps, sign folded in
};

struct Args45 {
    const void *frames;                 // frame of window 0 at step 0
    size_t frame_stride, step_stride;   // elements
    int pitch, H, W;
    const float *fill;
    const int2 *guess;                  // [n] start guess (1-based)
    int T;
    int4 *out_pos; float *out_resp;     // [n] last step
    int2 *next_guess;                   // [n] or null
    int4 *traj_pos; float *traj_resp;   // [T][n] or null
    int n;
};

} // namespace

template <typename PixT> __device__ __forceinline__ float ld_px(const PixT *p);
template <> __device__ __forceinline__ float ld_px<uint8_t>(const uint8_t *p) { return (float)__ldg(p); }
template <> __device__ __forceinline__ float ld_px<float>(const float *p) { return __ldg(p); }

template <typename PixT>
__global__ void __launch_bounds__(THREADS, 2)
dog_window45_argmax(const __grid_constant__ Args45 a, const __grid_constant__ Taps45 tp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_in = reinterpret_cast<float *>(smem_raw);                       // [FR][PIN]
    float2 *s_mid = reinterpret_cast<float2 *>(s_in + FR * PIN + 1);        // [FR][PM] (8-byte aligned: FR*PIN+1 is even)
    __shared__ unsigned long long s_key[NWARPS];
    __shared__ int2 s_guess;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v = blockIdx.x;
    const float fill = a.fill[v];
    int2 g = a.guess[v];

    for (int t = 0; t < a.T; ++t) {
        const PixT *frame = reinterpret_cast<const PixT *>(a.frames) + (size_t)t * a.step_stride + (size_t)v * a.frame_stride;
        const int wy0 = g.x - 1 - (WR / 2), wx0 = g.y - 1 - (WC / 2);   // window origin, 0-based
        const int fy0 = wy0 - HW, fx0 = wx0 - HW;                        // footprint origin

        // ---- stage: warp per row, lane = column (+32q): coalesced byte / float loads
        const bool interior = (fy0 >= 0) && (fx0 >= 0) && (fy0 + FR <= a.H) && (fx0 + FC <= a.W);
        if (interior) {
            for (int f = warp; f < FR; f += NWARPS) {
                const PixT *rowp = frame + (size_t)(fy0 + f) * a.pitch + fx0;
                float *dst = s_in + f * PIN;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = lane + 32 * q;
                    if (c < FC) dst[c] = ld_px<PixT>(rowp + c) - fill;
                }
            }
        } else {
            for (int f = warp; f < FR; f += NWARPS) {
                const int Y = fy0 + f;
                const bool yok = (Y >= 0) && (Y < a.H);
                const PixT *rowp = frame + (size_t)(yok ? Y : 0) * a.pitch;
                float *dst = s_in + f * PIN;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = lane + 32 * q;
                    const int X = fx0 + c;
                    if (c < FC) dst[c] = (yok && X >= 0 && X < a.W) ? ld_px<PixT>(rowp + X) - fill : 0.f;
                }
            }
        }
        __syncthreads();

        // ---- row pass: item = (footprint row f, column group gq); lanes walk rows
#pragma unroll 1
        for (int item = tid; item < ROW_ITEMS; item += THREADS) {
            const int gq = item / FR, f = item - gq * FR;
            const float *row = s_in + f * PIN + gq * R;
            float x[R + 2 * HW];
#pragma unroll
            for (int i = 0; i < R + 2 * HW; ++i) x[i] = row[i];
            float ap[R], am[R];
#pragma unroll
            for (int j = 0; j < R; ++j) { ap[j] = x[j + HW] * tp.rp[0]; am[j] = x[j + HW] * tp.rm[0]; }
#pragma unroll
            for (int d = 1; d <= HW; ++d) {
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const float s = x[j + HW - d] + x[j + HW + d];
                    ap[j] = fmaf(s, tp.rp[d], ap[j]);
                    am[j] = fmaf(s, tp.rm[d], am[j]);
                }
            }
            float2 *dst = s_mid + f * PM + gq * R;
#pragma unroll
            for (int j = 0; j < R; ++j) dst[j] = make_float2(ap[j], am[j]);
        }
        __syncthreads();

        // ---- column pass + running argmax: item = (column xq, row group h); lanes walk columns
        unsigned long long key = 0ull;
        if (tid < COL_ITEMS) {
            const int h = tid / WC, xq = tid - h * WC;
            const float2 *col = s_mid + (h * R) * PM + xq;
            float acc[R];
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = 0.f;
#pragma unroll
            for (int i = 0; i < R + 2 * HW; ++i) {
                const float2 m = col[i * PM];
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const int k = i - j;
                    if (k >= 0 && k < L) {
                        acc[j] = fmaf(m.x, tp.cp[k], acc[j]);
                        acc[j] = fmaf(m.y, tp.cm[k], acc[j]);
                    }
                }
            }
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
        if (lane == 0) s_key[warp] = key;
        __syncthreads();
        if (tid == 0) {
            unsigned long long k = s_key[0];
#pragma unroll
            for (int i = 1; i < NWARPS; ++i) k = s_key[i] > k ? s_key[i] : k;
            const unsigned int idx = key_index(k);
            const int xx = (int)(idx / WR), yy = (int)(idx - xx * WR);
            const int raw_i = wy0 + yy + 1, raw_j = wx0 + xx + 1;
            const int ci = min(max(raw_i, 1), a.H), cj = min(max(raw_j, 1), a.W);
            const float resp = key_value(k);
            const int4 p = make_int4(ci, cj, raw_i, raw_j);
            if (a.traj_pos) { a.traj_pos[(size_t)t * a.n + v] = p; a.traj_resp[(size_t)t * a.n + v] = resp; }
            if (t == a.T - 1) {
                a.out_pos[v] = p; a.out_resp[v] = resp;
                if (a.next_guess) a.next_guess[v] = make_int2(ci, cj);
            }
            s_guess = make_int2(ci, cj);
        }
        __syncthreads();
        g = s_guess;
    }
}

const char *window45_name() { return "dog_window45_argmax"; }

bool window45_supported(const WinArgs &a)
{
    return a.L == L && a.wr == WR && a.wc == WC && !a.rect_mode && a.map_out == nullptr;
}

// The specialised kernel takes its taps as kernel parameters (constant bank):
// fold the symmetric row factors (index d = |k − 32|) from the host copy.
static void fold_taps(const WinArgs &a, Taps45 &tp)
{
    const float *rp = a.h_taps, *rm = a.h_taps + L, *cp = a.h_taps + 2 * L, *cm = a.h_taps + 3 * L;
    for (int d = 0; d <= HW; ++d) { tp.rp[d] = rp[HW + d]; tp.rm[d] = rm[HW + d]; }
    for (int k = 0; k < L; ++k) { tp.cp[k] = cp[k]; tp.cm[k] = cm[k]; }
}

cudaError_t launch_window45(const WinArgs &a, int n, int pixel, cudaStream_t s)
{
    if (!a.h_taps) return cudaErrorInvalidValue;
    Taps45 tp;
    fold_taps(a, tp);
    cudaError_t e;
    Args45 k;
    k.frames = a.frames; k.frame_stride = a.frame_stride; k.step_stride = a.step_stride;
    k.pitch = a.pitch; k.H = a.H; k.W = a.W; k.fill = a.fill; k.guess = a.guess;
    k.T = a.T > 0 ? a.T : 1;
    k.out_pos = a.out_pos; k.out_resp = a.out_resp; k.next_guess = a.next_guess;
    k.traj_pos = a.traj_pos; k.traj_resp = a.traj_resp; k.n = n;
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
