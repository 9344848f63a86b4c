// pt_kernels.cu — generic DoG-window + argmax kernel (any window size, any
// kernel length that fits shared memory) and the mode (fill value) kernels.
//
// What it replaces: imfilter!(…, buff, img, kernel, NoPad(), window_indices)
// followed by findmax over the window view — src/PawsomeTracker.jl:57-59 — for
// every window of a batch in one launch.  Nothing here is derived from the
// reference's code (which is a dense l×l Float64 loop inside ImageFiltering.jl);
// the design is a streaming separable filter:
//
//   grid = (column strips of 32 outputs, row chunks, windows)
//   each CTA marches down its strip in batches of 32 footprint rows:
//     stage    32 × (32+2w) pixels → smem as (pixel − fill)   [0 outside the frame]
//     row pass both Gaussians at once → ring buffer of l+31 rows (float2 per output)
//     col pass 32 output rows whose 2w+1 ring rows are complete, fused with the
//              subtraction, the darker_target sign and a running per-thread argmax
//   block argmax → 64-bit atomicMax per window; the last CTA of a window decodes
//   the key, clamps and publishes (position, response, next guess).
//
// The row pass is computed exactly once per footprint row of the chunk and the
// column pass once per output, i.e. the algorithmic FLOP count of SURVEY §8(d)
// when chunks == 1.  Each thread register-tiles 8 outputs along the filter
// direction so one shared-memory load feeds 16 FMAs.
#include "pt_kernels.cuh"

namespace pt {

template <typename PixT> __device__ __forceinline__ float px_value(const PixT *p);
template <> __device__ __forceinline__ float px_value<uint8_t>(const uint8_t *p) { return (float)__ldg(p); }
template <> __device__ __forceinline__ float px_value<float>(const float *p) { return __ldg(p); }

__device__ __forceinline__ bool key_better(float v, unsigned int idx, float bv, unsigned int bidx)
{
    return v > bv || (v == bv && idx < bidx);
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

constexpr int kGenericWarps = kGenericThreads / 32;

// Stage TB footprint rows of one batch: tile column t ↔ frame column X0 + t, tile row rr ↔ frame row
// Y0 + rr; (pixel − fill) inside the frame, 0 outside (the PaddedView border after the fill shift).
// Generic path: lanes along the row, one element per lane (coalesced).
template <typename PixT>
__device__ __forceinline__ void stage_batch_scalar(const PixT *frame, int pitch, int H, int W, int Y0, int X0, int rows_valid,
                                                   float fill, float *s_in, int PIN, int warp, int lane)
{
    for (int rr = warp; rr < kBatchRows; rr += kGenericWarps) {
        const int Y = Y0 + rr;
        const bool yok = (Y >= 0) && (Y < H) && (rr < rows_valid);
        const PixT *rowp = frame + (size_t)(yok ? Y : 0) * pitch;
        float *dst = s_in + rr * PIN;
        for (int t = lane; t < PIN; t += 32) {
            const int X = X0 + t;
            dst[t] = (yok && X >= 0 && X < W) ? px_value<PixT>(rowp + X) - fill : 0.f;
        }
    }
}

// u8 frames with 4-byte-aligned rows: one 32-bit load = 4 pixels per lane, bytes outside the frame replaced by
// the fill byte (→ exactly 0), u8→f32 by the 2^23 trick (PRMT builds 8388608+px, the FADD that subtracts the
// fill finishes the conversion exactly).  All loads of a thread are issued before the first conversion.
template <int NWI>   // word iterations per row: ceil(words per row / 32)
__device__ __forceinline__ void stage_batch_words(const uint8_t *frame, int pitch, int H, int W, int Y0, int X0, int rows_valid,
                                                  float fill, float *s_in, int PIN, int warp, int lane)
{
    constexpr int RPW = kBatchRows / kGenericWarps;    // 8 rows per warp
    const int xa = X0 & ~3, phase = X0 - xa;
    const unsigned int fillw = (unsigned int)fill * 0x01010101u;
    const float cst = 8388608.0f + fill;
    unsigned int wd[RPW][NWI];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int rr = warp + r * kGenericWarps;
        const int Y = Y0 + rr;
        const bool yok = (Y >= 0) && (Y < H) && (rr < rows_valid);
#pragma unroll
        for (int i = 0; i < NWI; ++i) {
            const int wi = lane + 32 * i;
            const int X = xa + 4 * wi;
            unsigned int wv = fillw;
            if (yok && 4 * wi - phase < PIN && X + 3 >= 0 && X < W) {
                wv = __ldg(reinterpret_cast<const unsigned int *>(frame + (size_t)Y * pitch + X));
                unsigned int keep = 0u;
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) if (X + bb >= 0 && X + bb < W) keep |= 0xFFu << (8 * bb);
                wv = (wv & keep) | (fillw & ~keep);
            }
            wd[r][i] = wv;
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int rr = warp + r * kGenericWarps;
        float *dst = s_in + rr * PIN;
#pragma unroll
        for (int i = 0; i < NWI; ++i) {
            const int wi = lane + 32 * i;
            const unsigned int wv = wd[r][i];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                const int col = 4 * wi - phase + bb;
                const float val = __uint_as_float(__byte_perm(wv, 0x4B000000u, 0x7540 + bb)) - cst;
                if (col >= 0 && col < PIN) dst[col] = val;
            }
        }
    }
}

#ifndef PT_GEN_UNROLL
#define PT_GEN_UNROLL 2      // measured: 13.1 vs 14.2 us per 401x401 window at l = 245, 41 vs 44 us per 1080p frame at l = 65
#endif
#ifndef PT_GEN_MINBLOCKS
#define PT_GEN_MINBLOCKS 4
#endif
#define PT_PRAGMA(x) _Pragma(#x)
#define PT_UNROLL_N(n) PT_PRAGMA(unroll n)

template <typename PixT>
__global__ void __launch_bounds__(kGenericThreads, PT_GEN_MINBLOCKS)
dog_rect_argmax_generic(const WinArgs a)
{
    constexpr int TW = kTileCols, TB = kBatchRows, R = 8, C = kTapChunk;
    static_assert(TW == 32 && TB == 32, "lane/warp mapping assumes 32x32 tiles");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = a.L, w = a.w, Lpad = a.Lpad;
    const int Lq = ((L + 1 + C - 1) / C) * C;             // column pass walks tap pairs q = 0..L
    const int PIN = (TW + Lpad - 1) | 1;                  // odd pitch: lanes walk rows conflict-free
    const int RING = L + TB - 1;
    float4 *s_cq = reinterpret_cast<float4 *>(smem_raw);     // [Lq] (cp[q], cp[q-1], cm[q], cm[q-1]): column tap pairs, one
                                                          // LDS.128 per tap; one FFMA2 advances two vertically adjacent outputs
    float2 *s_trow = reinterpret_cast<float2 *>(s_cq + Lq);  // [Lpad] (narrow, wide) row taps
    constexpr int RP = TW + 1;                            // ring row pitch (float2): odd → row-pass stores conflict-free
    float2 *s_ring = s_trow + Lpad;                       // [RING][RP]
    float *s_in = reinterpret_cast<float *>(s_ring + (size_t)RING * RP); // [TB][PIN]
    __shared__ unsigned long long s_best[kGenericWarps];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v = blockIdx.z, strip = blockIdx.x, chunk = blockIdx.y;

    int wy0, wx0;
    if (a.rect_mode) { wy0 = a.ry0; wx0 = a.rx0; }
    else { int2 g = a.guess[v]; wy0 = g.x - 1 - a.rr; wx0 = g.y - 1 - a.rc; }

    const int c0 = strip * TW;
    const int r0 = chunk * a.CH;
    const int ch = min(a.CH, a.wr - r0);   // output rows of this chunk
    const int sw = min(TW, a.wc - c0);     // output cols of this strip
    const int nfoot = ch + 2 * w;          // footprint rows of this chunk
    const int nb = (nfoot + TB - 1) / TB;

    for (int k = tid; k < Lpad; k += kGenericThreads) s_trow[k] = a.taps_row[k];
    for (int q = tid; q < Lq; q += kGenericThreads) {
        const float2 cur = (q < L) ? a.taps_col[q] : make_float2(0.f, 0.f);
        const float2 prv = (q >= 1 && q - 1 < L) ? a.taps_col[q - 1] : make_float2(0.f, 0.f);
        s_cq[q] = make_float4(cur.x, prv.x, cur.y, prv.y);
    }
    for (int e = tid; e < RING * RP; e += kGenericThreads) s_ring[e] = make_float2(0.f, 0.f);

    const float fill = a.fill[v];
    const PixT *frame = reinterpret_cast<const PixT *>(a.frames) + (size_t)v * a.frame_stride;
    const int X0 = wx0 + c0 - w, Yb = wy0 + r0 - w;
    const int nwords = (PIN + 3 + 3) / 4;                 // words per staged row incl. phase
    const bool words_ok = sizeof(PixT) == 1 && ((reinterpret_cast<uintptr_t>(frame) | (uintptr_t)a.pitch) & 3u) == 0 &&
                          a.pitch >= ((a.W + 3) & ~3) && nwords <= 96;

    float best_v = -INFINITY;
    unsigned int best_i = 0xFFFFFFFFu;

    for (int b = 0; b < nb; ++b) {
        // ---- stage TB footprint rows
        const int rows_valid = nfoot - b * TB;
        if (words_ok) {
            const uint8_t *f8 = reinterpret_cast<const uint8_t *>(frame);
            if (nwords <= 32) stage_batch_words<1>(f8, a.pitch, a.H, a.W, Yb + b * TB, X0, rows_valid, fill, s_in, PIN, warp, lane);
            else if (nwords <= 64) stage_batch_words<2>(f8, a.pitch, a.H, a.W, Yb + b * TB, X0, rows_valid, fill, s_in, PIN, warp, lane);
            else stage_batch_words<3>(f8, a.pitch, a.H, a.W, Yb + b * TB, X0, rows_valid, fill, s_in, PIN, warp, lane);
        } else {
            stage_batch_scalar<PixT>(frame, a.pitch, a.H, a.W, Yb + b * TB, X0, rows_valid, fill, s_in, PIN, warp, lane);
        }
        // warm L2 with the next batch of rows while this one is filtered
        if (b + 1 < nb) {
            const int nl = ((PIN * (int)sizeof(PixT) + 127) >> 7) + 1;
            const long long rowbytes = (long long)a.W * (int)sizeof(PixT);
            const long long xb = ((long long)X0 * (int)sizeof(PixT)) & ~127LL;
            for (int rr = warp; rr < TB; rr += kGenericWarps) {
                const int Y = Yb + (b + 1) * TB + rr;
                for (int ln = lane; ln < nl; ln += 32) {
                    const long long off = xb + ((long long)ln << 7);
                    if (Y >= 0 && Y < a.H && off >= 0 && off < rowbytes)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(frame + (size_t)Y * a.pitch) + off));
                }
            }
        }
        __syncthreads();

        // ---- row pass: lane = footprint row of the batch, warp = group of 8 output columns.  The factors are
        // symmetric, so out = g0·x0 + Σ_d g_d·(x_-d + x_+d): one FADD feeds one packed FFMA2 that advances
        // (narrow, wide) together — 3 issue slots per tap pair instead of 4.  Two register windows slide
        // outwards from the centre (left one to the left, right one to the right), C taps per step; the
        // last w mod C tap pairs are handled one at a time.  Same summation order as row_pass45.
        {
            const float *ctr = s_in + lane * PIN + warp * R + w;      // x0 of output 0
            float2 acc[R];
            float lw[R + C - 1], rw[R + C - 1];
            {
                const float2 g0 = s_trow[w];
#pragma unroll
                for (int j = 0; j < R; ++j) { const float x0 = ctr[j]; acc[j] = make_float2(x0 * g0.x, x0 * g0.y); }
            }
#pragma unroll
            for (int i = 0; i < R - 1; ++i) { lw[C + i] = ctr[i]; rw[i] = ctr[1 + i]; }
            const int nfull = w / C;
            int d0 = 0;
            PT_UNROLL_N(PT_GEN_UNROLL)
            for (int ch = 0; ch < nfull; ++ch, d0 += C) {
                // lw[i] = ctr[-d0-C+i], rw[i] = ctr[d0+1+i]
#pragma unroll
                for (int i = 0; i < C; ++i) { lw[i] = ctr[-d0 - C + i]; rw[R - 1 + i] = ctr[d0 + R + i]; }
#pragma unroll
                for (int t = 1; t <= C; ++t) {
                    const float2 g = s_trow[w + d0 + t];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const float sm = lw[j + C - t] + rw[j + t - 1];
                        acc[j] = ffma2(make_float2(sm, sm), g, acc[j]);
                    }
                }
#pragma unroll
                for (int i = R - 2; i >= 0; --i) lw[C + i] = lw[i];
#pragma unroll
                for (int i = 0; i < R - 1; ++i) rw[i] = rw[i + C];
            }
#pragma unroll 1
            for (; d0 < w; ++d0) {
                // one tap pair: d = d0+1; left values ctr[-d0-1+j], right values ctr[d0+1+j]
                const float nl = ctr[-d0 - 1], nr = ctr[d0 + R];
                const float2 g = s_trow[w + d0 + 1];
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const float left = (j == 0) ? nl : lw[C + j - 1];
                    const float right = (j < R - 1) ? rw[j] : nr;
                    const float sm = left + right;
                    acc[j] = ffma2(make_float2(sm, sm), g, acc[j]);
                }
#pragma unroll
                for (int i = R - 2; i >= 1; --i) lw[C + i] = lw[C + i - 1];
                lw[C] = nl;
#pragma unroll
                for (int i = 0; i < R - 2; ++i) rw[i] = rw[i + 1];
                rw[R - 2] = nr;
            }
            const int f = b * TB + lane;
            float2 *dst = s_ring + (size_t)(f % RING) * RP + warp * R;
#pragma unroll
            for (int j = 0; j < R; ++j) dst[j] = acc[j];
        }
        __syncthreads();

        // ---- column pass: lane = output column, warp = group of 8 output rows; outputs (2p, 2p+1) share one
        // packed accumulator: input row 2p+k meets tap k of the even output and tap k-1 of the odd one
        const int o_base = b * TB - 2 * w;   // first output row whose support is now complete
        if (o_base + TB > 0 && o_base < ch) {
            const int f_start = o_base + warp * R;
            int slot = f_start % RING;
            if (slot < 0) slot += RING;
            const float2 *ring_x = s_ring + lane;
            float2 accP[R / 2], accM[R / 2];      // narrow / wide parts: 8 independent dependency chains
            float wp[R + C - 1], wm[R + C - 1];
#pragma unroll
            for (int p = 0; p < R / 2; ++p) { accP[p] = make_float2(0.f, 0.f); accM[p] = make_float2(0.f, 0.f); }
#pragma unroll
            for (int j = 0; j < R - 1; ++j) {
                const float2 m = ring_x[(size_t)slot * RP];
                wp[j] = m.x; wm[j] = m.y;
                slot = (slot + 1 == RING) ? 0 : slot + 1;
            }
            PT_UNROLL_N(PT_GEN_UNROLL)
            for (int k0 = 0; k0 < Lq; k0 += C) {
#pragma unroll
                for (int t = 0; t < C; ++t) {
                    const float2 m = ring_x[(size_t)slot * RP];
                    wp[R - 1 + t] = m.x; wm[R - 1 + t] = m.y;
                    slot = (slot + 1 == RING) ? 0 : slot + 1;
                }
#pragma unroll
                for (int t = 0; t < C; ++t) {
                    const float4 gq = s_cq[k0 + t];
                    const float2 gp = make_float2(gq.x, gq.y), gm = make_float2(gq.z, gq.w);
#pragma unroll
                    for (int p = 0; p < R / 2; ++p) accP[p] = ffma2(make_float2(wp[2 * p + t], wp[2 * p + t]), gp, accP[p]);
#pragma unroll
                    for (int p = 0; p < R / 2; ++p) accM[p] = ffma2(make_float2(wm[2 * p + t], wm[2 * p + t]), gm, accM[p]);
                }
#pragma unroll
                for (int j = 0; j < R - 1; ++j) { wp[j] = wp[j + C]; wm[j] = wm[j + C]; }
            }
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int o = f_start + j;
                if (o >= 0 && o < ch && lane < sw) {
                    const float val = ((j & 1) ? accP[j / 2].y + accM[j / 2].y : accP[j / 2].x + accM[j / 2].x) + 0.0f;
                    const unsigned int idx = (unsigned int)(c0 + lane) * (unsigned int)a.wr + (unsigned int)(r0 + o);
                    if (key_better(val, idx, best_v, best_i)) { best_v = val; best_i = idx; }
                    if (a.map_out)
                        a.map_out[(size_t)v * a.wr * a.wc + (size_t)(r0 + o) * a.wc + (c0 + lane)] = val;
                }
            }
        }
        // the next batch's staging only touches s_in; its row pass (which
        // overwrites the oldest ring rows read above) runs after the next barrier
    }

    // ---- block argmax → one 64-bit atomicMax per CTA
    unsigned long long key = (best_i == 0xFFFFFFFFu) ? 0ull : pack_key(best_v, best_i);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, key, off);
        key = o > key ? o : key;
    }
    if (lane == 0) s_best[warp] = key;
    __syncthreads();
    if (tid == 0) {
        unsigned long long k = s_best[0];
        for (int i = 1; i < kGenericWarps; ++i) k = s_best[i] > k ? s_best[i] : k;
        atomicMax(a.keys + v, k);
        __threadfence();
        const unsigned int total = (unsigned int)(a.strips * a.chunks);
        const unsigned int prev = atomicAdd(a.counters + v, 1u);
        s_last = (prev == total - 1u);
        if (s_last) {
            __threadfence();
            const unsigned long long win = atomicExch(a.keys + v, 0ull);
            a.counters[v] = 0u;
            publish_result(a, v, win, wy0, wx0);
        }
    }
}

size_t generic_smem_bytes(int L, int Lpad)
{
    const int PIN = (kTileCols + Lpad - 1) | 1;
    const int RING = L + kBatchRows - 1;
    const int Lq = ((L + 1 + kTapChunk - 1) / kTapChunk) * kTapChunk;
    return (size_t)(Lpad + 2 * Lq) * sizeof(float2) + (size_t)RING * (kTileCols + 1) * sizeof(float2) +
           (size_t)kBatchRows * PIN * sizeof(float);
}

// cudaFuncSetAttribute is per device: opt both instantiations in to the largest dynamic shared memory the device
// allows, once per device (pt_batch_create refuses kernel lengths whose footprint exceeds it).
cudaError_t generic_init_device()
{
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    // the opt-in limit covers static + dynamic shared memory: leave room for the kernel's few static bytes
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, dog_rect_argmax_generic<uint8_t>);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(dog_rect_argmax_generic<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&fa, dog_rect_argmax_generic<float>);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(dog_rect_argmax_generic<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

cudaError_t launch_generic(const WinArgs &a, int n, int pixel, cudaStream_t s)
{
    const size_t smem = generic_smem_bytes(a.L, a.Lpad);
    dim3 grid((unsigned)a.strips, (unsigned)a.chunks, (unsigned)n);
    if (pixel == 0) dog_rect_argmax_generic<uint8_t><<<grid, kGenericThreads, smem, s>>>(a);
    else dog_rect_argmax_generic<float><<<grid, kGenericThreads, smem, s>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// mode(frame) — fillvalue = mode(_img), src/PawsomeTracker.jl:47.
// StatsBase.mode returns the value whose count first reaches the final maximum while scanning in
// column-major order; a value reaches its final count at its LAST occurrence, so among the values
// tied for the maximum count the one whose last occurrence comes earliest wins.
//
// Fast path (one pass at HBM speed): counts only.  mode_count_kernel reads 16-byte vectors and
// accumulates into a shared histogram with one private column per LANE (hist[bin][lane]: the 32 lanes
// of a warp always hit 32 different banks, whatever the pixel values — a flat background does not
// serialise); uniform vectors / words cost a single atomic.  mode_decide_kernel picks the
// maximum; only if two or more values tie for it (rare) does the slow pass run: mode_hist_kernel
// also records every value's last column-major position and mode_pick_kernel applies the tie rule.
// Both slow kernels are always enqueued and return at once for videos without a tie, so the host
// never has to look at the counts.
// Scratch `hist`: [n][kModeScratch] unsigned, zero between calls: 0..255 counts + 256..511 last
// positions (slow pass), 512..767 counts (fast pass), 768 tie flag.
// ---------------------------------------------------------------------------
constexpr int kModeSplit = 16;   // CTAs per frame
constexpr int kModeThreads = 256;

template <typename PixT> __device__ __forceinline__ int px_bin(const PixT *p);
template <> __device__ __forceinline__ int px_bin<uint8_t>(const uint8_t *p) { return (int)__ldg(p); }
template <> __device__ __forceinline__ int px_bin<float>(const float *p)
{
    int b = __float2int_rn(__ldg(p) * 255.0f);
    return min(max(b, 0), 255);
}
__device__ __forceinline__ int f32_bin(float x)
{
    int b = __float2int_rn(x * 255.0f);
    return min(max(b, 0), 255);
}

template <typename PixT>
__global__ void __launch_bounds__(kModeThreads)
mode_count_kernel(const void *frames, size_t frame_stride, int pitch, int H, int W, unsigned int *hist)
{
    constexpr int PXV = 16 / (int)sizeof(PixT);            // pixels per 16-byte vector
    __shared__ unsigned int s_cnt[256 * 32];               // [bin][lane]
    const int v = blockIdx.y, lane = threadIdx.x & 31;
    const PixT *frame = reinterpret_cast<const PixT *>(frames) + (size_t)v * frame_stride;
    for (int i = threadIdx.x; i < 256 * 32; i += kModeThreads) s_cnt[i] = 0u;
    __syncthreads();
    const int rows_per = (H + gridDim.x - 1) / gridDim.x;
    const int y0 = blockIdx.x * rows_per, y1 = min(H, y0 + rows_per);
    const int vec_per_row = (W + PXV - 1) / PXV;
    const int total = (y1 - y0) * vec_per_row;
    unsigned int *col = s_cnt + lane;
    auto count_vec = [&](const uint4 &q, int nvalid) {
        if (sizeof(PixT) == 1) {
            // whole vector one value (flat background): one atomic; else word by word: a uniform word costs one
            // atomic, a mixed one four (no run bookkeeping: per pixel that costs more issue slots than it saves)
            const unsigned int rep = __byte_perm(q.x, 0, 0);
            if (nvalid == 16 && q.x == rep && q.y == rep && q.z == rep && q.w == rep) {
                atomicAdd(col + (q.x & 0xFFu) * 32, 16u);
            } else {
                const unsigned int wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const unsigned int w4 = wd[i];
                    const int nv4 = nvalid - 4 * i;                       // valid pixels of this word
                    if (nv4 >= 4 && w4 == __byte_perm(w4, 0, 0)) {
                        atomicAdd(col + (w4 & 0xFFu) * 32, 4u);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < nv4) atomicAdd(col + ((w4 >> (8 * k)) & 0xFFu) * 32, 1u);
                    }
                }
            }
        } else {
            const float fv[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < nvalid) atomicAdd(col + f32_bin(fv[k]) * 32, 1u);
        }
    };
    // four independent 16-byte loads in flight per thread before the first is consumed
    constexpr int U = 4;
    for (int e0 = threadIdx.x; e0 < total; e0 += U * kModeThreads) {
        uint4 q[U];
        int nv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * kModeThreads;
            nv[u] = 0;
            q[u] = make_uint4(0u, 0u, 0u, 0u);
            if (e < total) {
                const int r = e / vec_per_row, c = e - r * vec_per_row;
                q[u] = __ldg(reinterpret_cast<const uint4 *>(frame + (size_t)(y0 + r) * pitch) + c);
                nv[u] = min(PXV, W - c * PXV);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (nv[u] > 0) count_vec(q[u], nv[u]);
    }
    __syncthreads();
    // thread t sums bin t over the 32 lane columns (rotated start: conflict-free)
    unsigned int sum = 0u;
    const int t = threadIdx.x;
#pragma unroll 8
    for (int c = 0; c < 32; ++c) sum += s_cnt[t * 32 + ((c + t) & 31)];
    if (sum) atomicAdd(hist + (size_t)v * kModeScratch + 512 + t, sum);
}

// One CTA of 256 threads per video: maximum count; unique → publish the fill, else raise the tie flag.
__global__ void __launch_bounds__(256)
mode_decide_kernel(unsigned int *hist, int pixel, float *fill_out, int *fill_int_out)
{
    __shared__ unsigned long long s_key[8];
    __shared__ unsigned int s_ties[8];
    const int v = blockIdx.x, i = threadIdx.x;
    unsigned int *h = hist + (size_t)v * kModeScratch;
    const unsigned int cnt = h[512 + i];
    unsigned long long key = ((unsigned long long)cnt << 8) | (unsigned long long)i;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, key, off);
        key = o > key ? o : key;
    }
    if ((i & 31) == 0) s_key[i >> 5] = key;
    __syncthreads();
    unsigned long long k = s_key[0];
#pragma unroll
    for (int t = 1; t < 8; ++t) k = s_key[t] > k ? s_key[t] : k;
    const unsigned int maxc = (unsigned int)(k >> 8);
    const unsigned int tied = __popc(__ballot_sync(0xFFFFFFFFu, cnt == maxc));
    if ((i & 31) == 0) s_ties[i >> 5] = tied;
    __syncthreads();
    unsigned int nt = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) nt += s_ties[t];
    if (i == 0) {
        if (nt == 1) {
            const int bin = (int)(k & 0xFFull);
            fill_int_out[v] = bin;
            fill_out[v] = pixel == 0 ? (float)bin : (float)bin / 255.0f;
            h[768] = 0u;
        } else {
            h[768] = 1u;      // tie for the maximum count: the slow pass decides by last positions
        }
    }
    h[512 + i] = 0u;          // leave the scratch zeroed for the next call
}

template <typename PixT>
__global__ void __launch_bounds__(kModeThreads)
mode_hist_kernel(const void *frames, size_t frame_stride, int pitch, int H, int W, unsigned int *hist, int only_ties)
{
    __shared__ unsigned int s_cnt[256];
    __shared__ unsigned int s_last[256];
    const int v = blockIdx.y;
    if (only_ties && hist[(size_t)v * kModeScratch + 768] == 0u) return;
    const PixT *frame = reinterpret_cast<const PixT *>(frames) + (size_t)v * frame_stride;
    for (int i = threadIdx.x; i < 256; i += kModeThreads) { s_cnt[i] = 0u; s_last[i] = 0u; }
    __syncthreads();
    const int rows_per = (H + gridDim.x - 1) / gridDim.x;
    const int y0 = blockIdx.x * rows_per, y1 = min(H, y0 + rows_per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int y = y0 + warp; y < y1; y += kModeThreads / 32) {
        const PixT *row = frame + (size_t)y * pitch;
        for (int x0 = 0; x0 < W; x0 += 32) {
            const int x = x0 + lane;
            const bool ok = x < W;
            const int bin = ok ? px_bin<PixT>(row + x) : -1;
            const unsigned int pos = (unsigned int)x * (unsigned int)H + (unsigned int)y + 1u; // column-major, 1-based
            // warp-aggregate: one shared-memory atomic per distinct value in the warp
            unsigned int remaining = __ballot_sync(0xFFFFFFFFu, ok);
            while (remaining) {
                const int leader = __ffs(remaining) - 1;
                const int lbin = __shfl_sync(0xFFFFFFFFu, bin, leader);
                const bool mine = (bin == lbin);
                const unsigned int same = __ballot_sync(0xFFFFFFFFu, mine);
                const unsigned int mx = __reduce_max_sync(0xFFFFFFFFu, mine ? pos : 0u);
                if (lane == leader) {
                    atomicAdd(&s_cnt[lbin], (unsigned int)__popc(same));
                    atomicMax(&s_last[lbin], mx);
                }
                remaining &= ~same;
            }
        }
    }
    __syncthreads();
    unsigned int *h = hist + (size_t)v * kModeScratch;
    for (int i = threadIdx.x; i < 256; i += kModeThreads) {
        if (s_cnt[i]) { atomicAdd(h + i, s_cnt[i]); atomicMax(h + 256 + i, s_last[i]); }
    }
}

__global__ void __launch_bounds__(256)
mode_pick_kernel(unsigned int *hist, int pixel, float *fill_out, int *fill_int_out, int only_ties)
{
    __shared__ unsigned long long s_key[8];
    const int v = blockIdx.x, i = threadIdx.x;
    unsigned int *h = hist + (size_t)v * kModeScratch;
    if (only_ties && h[768] == 0u) return;
    const unsigned int cnt = h[i], last = h[256 + i];
    // larger count wins; ties → smaller last position
    unsigned long long key = ((unsigned long long)cnt << 40) | ((unsigned long long)(0xFFFFFFFFu - last) << 8) | (unsigned long long)i;
    if (cnt == 0) key = 0ull;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, key, off);
        key = o > key ? o : key;
    }
    if ((i & 31) == 0) s_key[i >> 5] = key;
    __syncthreads();
    if (i == 0) {
        unsigned long long k = s_key[0];
        for (int t = 1; t < 8; ++t) k = s_key[t] > k ? s_key[t] : k;
        const int bin = (int)(k & 0xFFull);
        fill_int_out[v] = bin;
        fill_out[v] = pixel == 0 ? (float)bin : (float)bin / 255.0f;
    }
    __syncthreads();
    h[i] = 0u; h[256 + i] = 0u; // leave the scratch zeroed for the next call
    if (i == 0) h[768] = 0u;
}

cudaError_t launch_mode(const void *frames, size_t frame_stride, int pitch, int H, int W, int n,
                        int pixel, unsigned int *hist, float *fill_out, int *fill_int_out,
                        bool force_slow, cudaStream_t s)
{
    dim3 grid(kModeSplit, (unsigned)n);
    const size_t es = pixel == 0 ? 1 : 4;
    // the vector path needs 16-byte aligned rows that can be read up to the next multiple of 16 bytes
    const bool vec = !force_slow && (reinterpret_cast<uintptr_t>(frames) & 15u) == 0 && ((size_t)pitch * es) % 16 == 0 &&
                     (frame_stride * es) % 16 == 0 && (size_t)pitch * es >= (((size_t)W * es + 15) & ~(size_t)15);
    cudaError_t e;
    if (vec) {
        if (pixel == 0) mode_count_kernel<uint8_t><<<grid, kModeThreads, 0, s>>>(frames, frame_stride, pitch, H, W, hist);
        else mode_count_kernel<float><<<grid, kModeThreads, 0, s>>>(frames, frame_stride, pitch, H, W, hist);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        mode_decide_kernel<<<n, 256, 0, s>>>(hist, pixel, fill_out, fill_int_out);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (pixel == 0)
        mode_hist_kernel<uint8_t><<<grid, kModeThreads, 0, s>>>(frames, frame_stride, pitch, H, W, hist, vec ? 1 : 0);
    else
        mode_hist_kernel<float><<<grid, kModeThreads, 0, s>>>(frames, frame_stride, pitch, H, W, hist, vec ? 1 : 0);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    mode_pick_kernel<<<n, 256, 0, s>>>(hist, pixel, fill_out, fill_int_out, vec ? 1 : 0);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Diagnostics downscale — `imresize!(dia.buffer, img)` to 360×640 (src/diagnose.jl:2,33): bilinear
// interpolation at pixel-centre aligned sample positions, one thread per output pixel, one launch for the
// current frame of every video.  HBM-bound (reads only the 4 neighbours of each output sample).
// ---------------------------------------------------------------------------
template <typename PixT> __device__ __forceinline__ float px_as_u8scale(const PixT *p);
template <> __device__ __forceinline__ float px_as_u8scale<uint8_t>(const uint8_t *p) { return (float)__ldg(p); }
template <> __device__ __forceinline__ float px_as_u8scale<float>(const float *p) { return __ldg(p) * 255.0f; }

template <typename PixT>
__global__ void __launch_bounds__(256)
downscale_kernel(const void *frames, size_t frame_stride, int pitch, int H, int W, int oh, int ow, uint8_t *out)
{
    const int v = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= ow || y >= oh) return;
    const PixT *frame = reinterpret_cast<const PixT *>(frames) + (size_t)v * frame_stride;
    const float sy = ((float)y + 0.5f) * ((float)H / (float)oh) - 0.5f;
    const float sx = ((float)x + 0.5f) * ((float)W / (float)ow) - 0.5f;
    const float fy = floorf(sy), fx = floorf(sx);
    const float wy = sy - fy, wx = sx - fx;
    const int y0 = min(max((int)fy, 0), H - 1), y1 = min(max((int)fy + 1, 0), H - 1);
    const int x0 = min(max((int)fx, 0), W - 1), x1 = min(max((int)fx + 1, 0), W - 1);
    const float p00 = px_as_u8scale<PixT>(frame + (size_t)y0 * pitch + x0), p01 = px_as_u8scale<PixT>(frame + (size_t)y0 * pitch + x1);
    const float p10 = px_as_u8scale<PixT>(frame + (size_t)y1 * pitch + x0), p11 = px_as_u8scale<PixT>(frame + (size_t)y1 * pitch + x1);
    const float top = p00 + wx * (p01 - p00), bot = p10 + wx * (p11 - p10);
    const float val = top + wy * (bot - top);
    out[((size_t)v * oh + y) * ow + x] = (uint8_t)min(max(__float2int_rn(val), 0), 255);
}

cudaError_t launch_downscale(const void *frames, size_t frame_stride, int pitch, int H, int W, int n, int pixel,
                             int oh, int ow, uint8_t *out, cudaStream_t s)
{
    dim3 grid((unsigned)((ow + 31) / 32), (unsigned)((oh + 7) / 8), (unsigned)n);
    if (pixel == 0) downscale_kernel<uint8_t><<<grid, 256, 0, s>>>(frames, frame_stride, pitch, H, W, oh, ow, out);
    else downscale_kernel<float><<<grid, 256, 0, s>>>(frames, frame_stride, pitch, H, W, oh, ow, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// gather_footprints — for page-locked HOST frames and geometries the chained per-window kernels do not cover (long
// kernels, large windows): copy each window's footprint out of its host frame into a device crop, ONCE per step and in
// 16-byte pieces (the streaming kernels would re-read the kernel-length halos of their column strips over PCIe: 1.4 MB
// instead of 0.4 MB per 401x401 window at l = 245).  The window position is on the device (the chain advances there), so
// the copy is a kernel, not a DMA descriptor.  Crop column 0 is the frame column ox = (footprint origin) rounded down to a
// 16-byte boundary; everything outside the frame is the fill value (the PaddedView border, src/PawsomeTracker.jl:48), so the
// filter kernels see a frame of fr x cp pixels that contains the whole footprint.  Also written: the crop's origin in
// the frame (publish_result translates the argmax back and clamps to the REAL frame, :60-61) and the guess in crop
// coordinates.
// ---------------------------------------------------------------------------------------------------
template <typename PixT>
__global__ void __launch_bounds__(256)
gather_footprints(const void *frames, size_t frame_stride, int pitch, int H, int W, const int2 *guess, const float *fill,
                  int rr, int rc, int w, int fr, int cp, void *crops, size_t crop_stride, int2 *org, int2 *cguess)
{
    constexpr int EPC = 16 / (int)sizeof(PixT);              // elements per 16-byte piece
    const int v = blockIdx.y;
    const int2 g = guess[v];
    const int oy = g.x - 1 - rr - w;
    const int ox = (g.y - 1 - rc - w) & ~(EPC - 1);          // rounds down (two's complement), also when negative
    if (blockIdx.x == 0 && threadIdx.x == 0) { org[v] = make_int2(oy, ox); cguess[v] = make_int2(g.x - oy, g.y - ox); }
    const int nc = cp / EPC;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= fr * nc) return;
    const int y = e / nc, c = e - y * nc;
    const int Y = oy + y, X0 = ox + c * EPC;
    const PixT *frame = reinterpret_cast<const PixT *>(frames) + (size_t)v * frame_stride;
    PixT *dst = reinterpret_cast<PixT *>(crops) + (size_t)v * crop_stride + (size_t)y * cp + c * EPC;
    PixT fv;
    if (sizeof(PixT) == 1) fv = (PixT)(unsigned char)fill[v]; else fv = (PixT)fill[v];
    const bool yok = Y >= 0 && Y < H;
    const PixT *src = frame + (size_t)(yok ? Y : 0) * pitch + X0;
    if (yok && X0 >= 0 && X0 + EPC <= W && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        uint4 q;
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(src));
        *reinterpret_cast<uint4 *>(dst) = q;
    } else {
        PixT tmp[EPC];
#pragma unroll
        for (int k = 0; k < EPC; ++k) {
            const int X = X0 + k;
            tmp[k] = (yok && X >= 0 && X < W) ? src[k] : fv;
        }
#pragma unroll
        for (int k = 0; k < EPC; ++k) dst[k] = tmp[k];
    }
}

cudaError_t launch_gather_footprints(const void *frames, size_t frame_stride, int pitch, int H, int W, int n, int pixel,
                                     const int2 *guess, const float *fill, int rr, int rc, int w, int fr, int cp,
                                     void *crops, size_t crop_stride, int2 *org, int2 *cguess, cudaStream_t s)
{
    const int epc = pixel == 0 ? 16 : 4;
    const long long pieces = (long long)fr * (cp / epc);
    dim3 grid((unsigned)((pieces + 255) / 256), (unsigned)n);
    if (pixel == 0) gather_footprints<uint8_t><<<grid, 256, 0, s>>>(frames, frame_stride, pitch, H, W, guess, fill, rr, rc, w, fr, cp, crops, crop_stride, org, cguess);
    else gather_footprints<float><<<grid, 256, 0, s>>>(frames, frame_stride, pitch, H, W, guess, fill, rr, rc, w, fr, cp, crops, crop_stride, org, cguess);
    return cudaGetLastError();
}

} // namespace pt
