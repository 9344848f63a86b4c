// pt_generic64.cu — dog_rect_argmax_wide: the streaming separable DoG + argmax for ANY kernel length (4K with
// target_width = 100 → l = 245, BASELINE config 4; every target_width ≠ 25) built for instruction efficiency.
//
// What it replaces: imfilter!(…, buff, img, kernel, NoPad(), window_indices) + findmax — src/PawsomeTracker.jl:57-59 —
// like dog_rect_argmax_generic (pt_kernels.cu), with the same decomposition (column strips × row chunks × windows,
// 64-bit atomicMax merge) and the same per-output operation order.  The ncu profile of the older kernel at l = 245
// showed FMA pipe 48 %, issue slots 50 %, and ALU (integer) instructions at 43 % of everything issued: ring
// wrap-around arithmetic per loaded row and register moves of the sliding windows.  Here the inner loops carry no
// integer work at all:
//
//   * 64-column strips, 256 threads (8 warps), batches of 32 footprint rows: half the staging per output column;
//   * sliding windows live in CIRCULAR REGISTER BUFFERS of 16 slots indexed by (position mod 16); the loops are
//     unrolled by exactly 16 taps, so every slot index is a compile-time constant — no register moves;
//   * the ring of row-pass rows has a multiple of 8 rows and every thread starts reading at a row that is a multiple
//     of 8 (the residue of its first needed row is a launch constant δ = (l − 1) mod 8 ∈ {0, 4} → template parameter),
//     so each run of 8 loaded rows is contiguous: immediate offsets, one wrap test per 8 rows;
//   * taps are zero-padded to multiples of 16 (the padded taps multiply finite values by 0).
//   Per 16 column-pass taps: 16 LDS.64 + 16 LDS.128 + 128 FFMA2 (+ 6 integer instructions); per 16 row-pass tap
//   pairs: 32 LDS.32 + 16 LDS.64 + 128 FADD + 128 FFMA2.
#include "pt_kernels.cuh"

namespace pt {

namespace {

constexpr int SW64 = 64;         // output columns per strip
constexpr int TB64 = kBatchRows; // 32 footprint rows per batch
constexpr int THREADS64 = 256, WARPS64 = 8;
constexpr int R64 = 8;           // outputs per thread along the filter direction
constexpr int U64 = 16;          // unroll = circular register buffer size
constexpr int RP64 = SW64 + 1;   // ring row pitch (float2), odd: row-pass stores conflict-free
constexpr int LPAD = 4, RPAD = 12;   // never-used columns left / right of a staged row that the window prefetch may load
constexpr int kMaxLWide = 285;       // longest kernel whose ring + staged batch fit the shared memory of one SM
constexpr int kMaxW16 = 144, kMaxLq16 = 288;

// Taps as a kernel parameter (constant bank): the unrolled loops read them with a warp-uniform index into UNIFORM
// registers, so a packed FFMA2 reads three vector registers (value, accumulator pair) instead of five — with the taps
// in vector registers (loaded from shared memory) the FMA pipe ran at 56 % inside a CTA.
struct WideTaps {
    float4 cq[kMaxLq16];          // (cp[q], cp[q-1], cm[q], cm[q-1]), zero beyond q = L
    float2 trow[kMaxW16 + 1];     // folded row taps (narrow, wide), index d = |k − w|, zero beyond d = w
};

__device__ __forceinline__ float2 ffma2w(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

struct Geom64 {
    int w16;        // row tap pairs padded to a multiple of 16
    int Lq16;       // column tap pairs (q = 0..L) padded to a multiple of 16
    int width;      // staged columns per row: 64 + 2·w16
    int pin;        // s_in pitch (floats, odd): LPAD + width + RPAD
    int nring;      // ring rows (multiple of 8)
    int nwords;     // aligned 32-bit words covering a staged row at any phase
    size_t bytes;   // dynamic shared memory
};

__host__ __device__ inline Geom64 geom64(int L)
{
    Geom64 g;
    const int w = L / 2;
    g.w16 = ((w + U64 - 1) / U64) * U64;
    g.Lq16 = ((L + 1 + U64 - 1) / U64) * U64;
    g.width = SW64 + 2 * g.w16;
    g.pin = (LPAD + g.width + RPAD) | 1;
    const int delta = (L - 1) & 7;
    g.nring = ((L + TB64 - 1 + delta + 7) / 8) * 8;
    g.nwords = (g.width + 3 + 3) / 4;
    g.bytes = (size_t)g.nring * RP64 * sizeof(float2) + (size_t)TB64 * g.pin * sizeof(float);
    return g;
}

// Stage 32 footprint rows × `width` columns: s_in[rr][LPAD + t] = pixel(Y0 + rr, X0 + t) − fill, 0 outside the frame
// (the PaddedView border after the fill shift) and for rows beyond rows_valid.
template <typename PixT>
__device__ __forceinline__ void stage64_scalar(const PixT *frame, int pitch, int H, int W, int Y0, int X0, int rows_valid,
                                               float fill, float *s_in, int pin, int width, int warp, int lane)
{
    for (int rr = warp; rr < TB64; rr += WARPS64) {
        const int Y = Y0 + rr;
        const bool yok = (Y >= 0) && (Y < H) && (rr < rows_valid);
        const PixT *rowp = frame + (size_t)(yok ? Y : 0) * pitch;
        float *dst = s_in + rr * pin + LPAD;
        for (int t = lane; t < width; t += 32) {
            const int X = X0 + t;
            float v = 0.f;
            if (yok && X >= 0 && X < W) v = (float)__ldg(rowp + X) - fill;
            dst[t] = v;
        }
    }
}

// u8 frames with 4-byte aligned rows: aligned 32-bit loads — all of a warp's loads for its 4 rows are issued before the
// first conversion — bytes outside the frame replaced by the fill byte, u8→f32 by the 2^23 trick.
// (A software-pipelined variant — cp.async of the next batch's raw rows into a double buffer during the passes — was
// measured: 14.2 instead of 13.6 µs per 401x401 window at l = 245; the extra barrier and the conversion out of shared
// memory cost more than the hidden L2 latency.)
__device__ __forceinline__ void stage64_words(const uint8_t *frame, int pitch, int H, int W, int Y0, int X0, int rows_valid,
                                              float fill, float *s_in, int pin, int nwords, int width, int warp, int lane)
{
    constexpr int RPW = TB64 / WARPS64;    // 4 rows per warp
    const int xa = X0 & ~3, phase = X0 - xa;
    const unsigned int fillw = (unsigned int)fill * 0x01010101u;
    const float cst = 8388608.0f + fill;
    for (int wi = lane; wi < nwords; wi += 32) {
        const int X = xa + 4 * wi;
        const bool xok = 4 * wi - phase < width && X + 3 >= 0 && X < W;
        unsigned int keep = 0u;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) if (X + bb >= 0 && X + bb < W) keep |= 0xFFu << (8 * bb);
        unsigned int wd[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int rr = warp + r * WARPS64;
            const int Y = Y0 + rr;
            wd[r] = fillw;
            if (xok && (Y >= 0) && (Y < H) && (rr < rows_valid))
                asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(wd[r]) : "l"(frame + (size_t)Y * pitch + X));
        }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int rr = warp + r * WARPS64;
            const unsigned int wv = (wd[r] & keep) | (fillw & ~keep);
            float *dst = s_in + rr * pin + LPAD;
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                const int col = 4 * wi - phase + bb;
                const float val = __uint_as_float(__byte_perm(wv, 0x4B000000u, 0x7540 + bb)) - cst;
                if (col >= 0 && col < width) dst[col] = val;
            }
        }
    }
}


// ---- row pass of one thread: 8 consecutive outputs of one staged row (ctr = x(c), c = centre of output 0)
__device__ __forceinline__ void row_pass64(const float *ctr, const WideTaps &wt, int w16, float2 (&acc)[R64])
{
    float Lb[U64], Rb[U64];
    {
        const float2 g0 = wt.trow[0];
#pragma unroll
        for (int j = 0; j < R64; ++j) { const float x0 = ctr[j]; acc[j] = make_float2(x0 * g0.x, x0 * g0.y); }
    }
#pragma unroll
    for (int k = -6; k <= 4; ++k) Lb[k & 15] = ctr[-k];         // xl(-6 … 4)
#pragma unroll
    for (int k = 1; k <= 11; ++k) Rb[k & 15] = ctr[k];          // xr(1 … 11)
    const float *pl = ctr, *pr = ctr;
#pragma unroll 1
    for (int d0 = 0; d0 < w16; d0 += U64) {
#pragma unroll
        for (int u = 0; u < U64; ++u) {
            const int d = u + 1;                                  // step d0 + d; slots depend on d mod 16 only
            Lb[(d + 4) & 15] = pl[-(d + 4)];                      // xl(d0 + d + 4)
            Rb[(d + 11) & 15] = pr[d + 11];                       // xr(d0 + d + 11)
            const float2 g = wt.trow[d0 + d];                     // warp-uniform index → uniform registers
#pragma unroll
            for (int j = 0; j < R64; ++j) {
                const float sm = Lb[(d - j) & 15] + Rb[(d + j) & 15];
                acc[j] = ffma2w(make_float2(sm, sm), g, acc[j]);
            }
        }
        pl -= U64; pr += U64;
    }
}

// ---- column pass + running argmax for the 32 output rows whose support is complete with batch b of a chunk:
// lane = output column (warps w and w + 4 side by side cover 64), warp & 3 = group of 8 output rows.
// Outputs (2p, 2p+1) share a packed accumulator: the ring row met at tap q by output 2p meets tap q−1 at output
// 2p+1.  Ring rows are numbered n = footprint row − f_first with f_first = o0 − δ a multiple of 8, kept in a
// 16-slot circular register buffer (slot = n mod 16) and fetched 12 rows ahead in contiguous runs of 8.
template <int DELTA, int RP = RP64>
__device__ __forceinline__ void col_pass64(const WinArgs &a, const WideTaps &wt, const Geom64 &G, const float2 *s_ring, int b,
                                           int w, int ch, int sw, int c0, int r0, int v, int warp, int lane,
                                           float &best_v, unsigned int &best_i, int pad = 0)
{
    // ring row f holds footprint row f − pad of the chunk (pad = 0 in the fused kernel; the two-phase column kernel
    // shifts the rows so that whole 32-row output batches complete together)
    const int o_base = b * TB64 - 2 * w - pad;   // first output row whose support is complete with this batch
    if (o_base + TB64 > 0 && o_base < ch && 32 * (warp >> 2) < sw) {
        // warps 0-3 take the left 32 columns (one row group each → one warp per scheduler), warps 4-7 the right
        // 32: a strip of ≤ 32 columns keeps all four schedulers busy with half the work
        const int col = lane + 32 * (warp >> 2);
        const int o0 = o_base + R64 * (warp & 3);
        const int f_first = o0 + pad - DELTA;                         // ≡ 0 (mod 8)
        int s0 = f_first % G.nring;
        if (s0 < 0) s0 += G.nring;
        const float2 *ring_lo = s_ring + col, *ring_hi = s_ring + (size_t)G.nring * RP + col;
        const float2 *run = ring_lo + (size_t)s0 * RP;              // run of 8 rows holding n = 0 … 7
        float2 D[U64];
        float2 accP[R64 / 2], accM[R64 / 2];
#pragma unroll
        for (int p = 0; p < R64 / 2; ++p) { accP[p] = make_float2(0.f, 0.f); accM[p] = make_float2(0.f, 0.f); }
#pragma unroll
        for (int n = 0; n < 8; ++n) D[n] = run[n * RP];
        run += 8 * RP; if (run >= ring_hi) run -= (size_t)G.nring * RP;
#pragma unroll
        for (int n = 8; n < 12; ++n) D[n] = run[(n - 8) * RP];      // run now holds n = 8 … 15
#pragma unroll 1
        for (int q0 = 0; q0 < G.Lq16; q0 += U64) {
#pragma unroll
            for (int u = 0; u < U64; ++u) {
                // fetch row n = q0 + u + 12 (run-relative index (u + 4) mod 8; a new run starts at u = 4 and u = 12)
                if (u == 4 || u == 12) { run += 8 * RP; if (run >= ring_hi) run -= (size_t)G.nring * RP; }
                D[(u + 12) & 15] = run[((u + 4) & 7) * RP];
                const float4 gq = wt.cq[q0 + u];                      // warp-uniform index → uniform registers
                const float2 gp = make_float2(gq.x, gq.y), gm = make_float2(gq.z, gq.w);
#pragma unroll
                for (int p = 0; p < R64 / 2; ++p) {
                    const float2 m = D[(DELTA + 2 * p + u) & 15];
                    accP[p] = ffma2w(make_float2(m.x, m.x), gp, accP[p]);
                }
#pragma unroll
                for (int p = 0; p < R64 / 2; ++p) {
                    const float2 m = D[(DELTA + 2 * p + u) & 15];
                    accM[p] = ffma2w(make_float2(m.y, m.y), gm, accM[p]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R64; ++j) {
            const int o = o0 + j;
            if (o >= 0 && o < ch && col < sw) {
                const float val = ((j & 1) ? accP[j / 2].y + accM[j / 2].y : accP[j / 2].x + accM[j / 2].x) + 0.0f;
                const unsigned int idx = (unsigned int)(c0 + col) * (unsigned int)a.wr + (unsigned int)(r0 + o);
                if (val > best_v || (val == best_v && idx < best_i)) { best_v = val; best_i = idx; }
                if (a.map_out)
                    a.map_out[(size_t)v * a.wr * a.wc + (size_t)(r0 + o) * a.wc + (c0 + col)] = val;
            }
        }
    }
}

} // namespace

// ---- block argmax → one 64-bit atomicMax per CTA; the last CTA of a window decodes, clamps and publishes
__device__ __forceinline__ void merge_and_publish64(const WinArgs &a, int v, int wy0, int wx0, float best_v, unsigned int best_i,
                                                    unsigned long long *s_best, int tid, int warp, int lane, int nwarps = WARPS64)
{
    unsigned long long key = (best_i == 0xFFFFFFFFu) ? 0ull : pack_key(best_v, best_i);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, key, off);
        key = o > key ? o : key;
    }
    if (lane == 0) s_best[warp] = key;
    __syncthreads();
    if (tid == 0) {
        unsigned long long k = s_best[0];
        for (int i = 1; i < nwarps; ++i) k = s_best[i] > k ? s_best[i] : k;
        atomicMax(a.keys + v, k);
        __threadfence();
        const unsigned int total = (unsigned int)(a.strips * a.chunks);
        const unsigned int prev = atomicAdd(a.counters + v, 1u);
        if (prev == total - 1u) {
            __threadfence();
            const unsigned long long win = atomicExch(a.keys + v, 0ull);
            a.counters[v] = 0u;
            publish_result(a, v, win, wy0, wx0);
        }
    }
}

template <typename PixT, int DELTA>
__global__ void __launch_bounds__(THREADS64, 1)
dog_rect_argmax_wide(const __grid_constant__ WinArgs a, const __grid_constant__ WideTaps wt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = a.L, w = a.w;
    const Geom64 G = geom64(L);
    float2 *s_ring = reinterpret_cast<float2 *>(smem_raw);               // [nring][RP64] row-pass rows, slot = footprint row mod nring
    float *s_in = reinterpret_cast<float *>(s_ring + (size_t)G.nring * RP64);   // [32][pin]
    __shared__ unsigned long long s_best[WARPS64];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 1-D grid in longest-first order: the hardware hands CTAs to SMs in block order and one CTA fills an SM, so the
    // last, partial round of a launch decides its length.  All full 64-column strips come first, the (cheaper) narrow
    // last strips of the windows form the tail: 64 windows x 7 strips of a 401-column window run 3.0 instead of 4.0
    // full-strip times.
    int v, strip, chunk;
    {
        const int per_full = a.strips - 1, nc = (int)gridDim.x / a.strips;     // nc = chunks x windows
        const int id = (int)blockIdx.x;
        if (id < per_full * nc) { strip = id % per_full; const int r = id / per_full; chunk = r % a.chunks; v = r / a.chunks; }
        else { strip = a.strips - 1; const int r = id - per_full * nc; chunk = r % a.chunks; v = r / a.chunks; }
    }

    int wy0, wx0;
    if (a.rect_mode) { wy0 = a.ry0; wx0 = a.rx0; }
    else { const int2 g = a.guess[v]; wy0 = g.x - 1 - a.rr; wx0 = g.y - 1 - a.rc; }

    const int c0 = strip * SW64;
    const int r0 = chunk * a.CH;
    const int ch = min(a.CH, a.wr - r0);   // output rows of this chunk
    const int sw = min(SW64, a.wc - c0);   // output cols of this strip
    const int nfoot = ch + 2 * w;          // footprint rows of this chunk
    const int nb = (nfoot + TB64 - 1) / TB64;

    for (int e = tid; e < G.nring * RP64; e += THREADS64) s_ring[e] = make_float2(0.f, 0.f);
    for (int e = tid; e < TB64 * G.pin; e += THREADS64) s_in[e] = 0.f;      // incl. the never-used pad columns

    const float fill = a.fill[v];
    const PixT *frame = reinterpret_cast<const PixT *>(a.frames) + (size_t)v * a.frame_stride;
    const int X0 = wx0 + c0 - G.w16, Yb = wy0 + r0 - w;
    const bool words_ok = sizeof(PixT) == 1 && ((reinterpret_cast<uintptr_t>(frame) | (uintptr_t)a.pitch) & 3u) == 0 &&
                          a.pitch >= ((a.W + 3) & ~3);
    const uint8_t *f8 = reinterpret_cast<const uint8_t *>(frame);
    __syncthreads();

    float best_v = -INFINITY;
    unsigned int best_i = 0xFFFFFFFFu;

    for (int b = 0; b < nb; ++b) {
        // ---- stage 32 footprint rows
        const int rows_valid = nfoot - b * TB64;
        if (words_ok) stage64_words(f8, a.pitch, a.H, a.W, Yb + b * TB64, X0, rows_valid, fill, s_in, G.pin, G.nwords, G.width, warp, lane);
        else stage64_scalar<PixT>(frame, a.pitch, a.H, a.W, Yb + b * TB64, X0, rows_valid, fill, s_in, G.pin, G.width, warp, lane);
        // warm L2 with the next batch of rows while this one is filtered
        if (b + 1 < nb) {
            const int pb = b + 1;
            const int nl = ((G.width * (int)sizeof(PixT) + 127) >> 7) + 1;
            const long long rowbytes = (long long)a.W * (int)sizeof(PixT);
            const long long xb = ((long long)X0 * (int)sizeof(PixT)) & ~127LL;
            for (int rr = warp; rr < TB64; rr += WARPS64) {
                const int Y = Yb + pb * TB64 + rr;
                for (int ln = lane; ln < nl; ln += 32) {
                    const long long off = xb + ((long long)ln << 7);
                    if (Y >= 0 && Y < a.H && off >= 0 && off < rowbytes)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(frame + (size_t)Y * a.pitch) + off));
                }
            }
        }
        __syncthreads();

        // ---- row pass: lane = footprint row of the batch, warp = group of 8 output columns.  Symmetric factors:
        // out_j = g0·x(c+j) + Σ_d g_d·(x(c+j−d) + x(c+j+d)); one FADD feeds one packed FFMA2 advancing (narrow, wide).
        // Left values xl(k) = x(c − k), k ∈ [d−7, d] at step d, right values xr(k) = x(c + k), k ∈ [d, d+7]: both live in
        // 16-slot circular register buffers (slot = k mod 16) and are fetched 4 steps ahead.
        if (warp * R64 < sw) {
            float2 acc[R64];
            row_pass64(s_in + lane * G.pin + LPAD + G.w16 + warp * R64, wt, G.w16, acc);
            const int f = b * TB64 + lane;
            float2 *dst = s_ring + (size_t)(f % G.nring) * RP64 + warp * R64;
#pragma unroll
            for (int j = 0; j < R64; ++j) dst[j] = acc[j];
        }
        __syncthreads();

        // ---- column pass for the output rows whose support is complete with this batch
        col_pass64<DELTA>(a, wt, G, s_ring, b, w, ch, sw, c0, r0, v, warp, lane, best_v, best_i);
        // the next batch's staging only touches s_in; its row pass (which overwrites the oldest ring rows read above)
        // runs after the next barrier
    }

    merge_and_publish64(a, v, wy0, wx0, best_v, best_i, s_best, tid, warp, lane);
}

// ---------------------------------------------------------------------------------------------------
// Two-phase variant for launches that cannot fill the GPU with whole strips (a single 4K video with a 401×401 window at
// l = 245 is 7 strips): splitting a strip into row chunks repeats 2w footprint rows of the row pass per chunk (at 21
// chunks 6× the arithmetic).  Instead
//   dog_rows_wide   row-filters every (32-row batch, 64-column strip) of the window's footprint exactly once — all
//                   batches are independent, 147 CTAs for the 4K window — and writes the (narrow, wide) intermediate
//                   to global memory (L2-resident: 2.3 MB per window);
//   dog_cols_wide   runs the column pass + argmax per (row chunk, strip): it copies the chunk's intermediate rows into
//                   the ring (the redundancy between chunks is now a re-READ of 2w rows from L2, not arithmetic) and
//                   then is the fused kernel's column pass and merge.
// Same per-output operation order as the fused kernel: bit-identical responses.
// ---------------------------------------------------------------------------------------------------
template <typename PixT>
__global__ void __launch_bounds__(THREADS64, 2)
dog_rows_wide(const __grid_constant__ WinArgs a, const __grid_constant__ WideTaps wt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = a.L, w = a.w;
    const Geom64 G = geom64(L);
    float *s_in = reinterpret_cast<float *>(smem_raw);                    // [32][pin]
    // programmatic dependent launch: the column kernel may be scheduled as soon as every CTA of this grid is running (it
    // zeroes its ring meanwhile and waits for this grid's completion — griddepcontrol.wait — before it reads the intermediate)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // … and this kernel is itself a dependent launch of whatever precedes it in the stream (in a chain of steps: the
    // previous step's column kernel, which wrote the guess read below): its launch latency overlaps that kernel's tail
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nfoot = a.wr + 2 * w, nbt = (nfoot + TB64 - 1) / TB64;
    const int id = (int)blockIdx.x;
    const int strip = id % a.strips, r = id / a.strips, b = r % nbt, v = r / nbt;
    int wy0, wx0;
    if (a.rect_mode) { wy0 = a.ry0; wx0 = a.rx0; }
    else { const int2 g = a.guess[v]; wy0 = g.x - 1 - a.rr; wx0 = g.y - 1 - a.rc; }
    const int c0 = strip * SW64, sw = min(SW64, a.wc - c0);
    // (no zero-fill: staging writes every column the row pass consumes, zeros for rows beyond the footprint; the pad
    // columns left and right are only ever loaded ahead into the register windows, never used)
    const float fill = a.fill[v];
    const PixT *frame = reinterpret_cast<const PixT *>(a.frames) + (size_t)v * a.frame_stride;
    const int X0 = wx0 + c0 - G.w16, Yb = wy0 - w;
    const bool words_ok = sizeof(PixT) == 1 && ((reinterpret_cast<uintptr_t>(frame) | (uintptr_t)a.pitch) & 3u) == 0 &&
                          a.pitch >= ((a.W + 3) & ~3);
    const int rows_valid = nfoot - b * TB64;
    if (words_ok) stage64_words(reinterpret_cast<const uint8_t *>(frame), a.pitch, a.H, a.W, Yb + b * TB64, X0, rows_valid, fill,
                                s_in, G.pin, G.nwords, G.width, warp, lane);
    else stage64_scalar<PixT>(frame, a.pitch, a.H, a.W, Yb + b * TB64, X0, rows_valid, fill, s_in, G.pin, G.width, warp, lane);
    __syncthreads();
    const int f = b * TB64 + lane;
    if (warp * R64 < sw && f < nfoot) {
        float2 acc[R64];
        row_pass64(s_in + lane * G.pin + LPAD + G.w16 + warp * R64, wt, G.w16, acc);
        // intermediate: [window][footprint row][strips·64] float2, 16-byte aligned groups of 8
        float4 *dst = reinterpret_cast<float4 *>(a.mid + ((size_t)v * nfoot + f) * (size_t)(a.strips * SW64) + c0 + warp * R64);
#pragma unroll
        for (int j = 0; j < R64; j += 2) dst[j / 2] = make_float4(acc[j].x, acc[j].y, acc[j + 1].x, acc[j + 1].y);
    }
}

constexpr int PFC = 4;                   // batches of intermediate rows in flight ahead of the column pass
constexpr int RPC = SW64;                // ring row pitch of dog_cols_wide (float2): no row-pass stores here, so no odd pitch —
                                         // rows are 16-byte aligned (one cp.async per lane and row) and 224 rows fit an SM twice

// shared memory of dog_cols_wide: the fused kernel's ring + room for the batches in flight — or, when that is more than
// a chunk's rows altogether (short kernels, low chunks), just those rows: 83 KB instead of 125 KB at l = 77 with 64-row
// chunks, so two CTAs share an SM and a 256-window launch is one wave instead of two
__host__ __device__ inline int cols_ring_rows(int L, int CH)
{
    const int w = L / 2, pad = (TB64 - (2 * w) % TB64) % TB64;
    const int all = ((CH + 2 * w + pad + TB64 - 1) / TB64) * TB64;
    const int ring = geom64(L).nring + PFC * TB64;
    return all < ring ? all : ring;
}

template <int DELTA>
__global__ void __launch_bounds__(THREADS64, 2)
dog_cols_wide(const __grid_constant__ WinArgs a, const __grid_constant__ WideTaps wt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = a.L, w = a.w;
    Geom64 G = geom64(L);
    G.nring = cols_ring_rows(L, a.CH);
    float2 *s_ring = reinterpret_cast<float2 *>(smem_raw);               // [nring][RPC], slot = ring row mod nring
    __shared__ unsigned long long s_best[WARPS64];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // (the next step's row kernel may be scheduled)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int v, strip, chunk;
    {
        const int per_full = a.strips - 1, nc = (int)gridDim.x / a.strips;     // longest-first order, as the fused kernel
        const int id = (int)blockIdx.x;
        if (id < per_full * nc) { strip = id % per_full; const int r = id / per_full; chunk = r % a.chunks; v = r / a.chunks; }
        else { strip = a.strips - 1; const int r = id - per_full * nc; chunk = r % a.chunks; v = r / a.chunks; }
    }
    const int c0 = strip * SW64;
    const int r0 = chunk * a.CH;
    const int ch = min(a.CH, a.wr - r0);
    const int sw = min(SW64, a.wc - c0);
    // Ring row f ↔ row r0 + f − pad of the window's footprint, pad chosen so that 2w + pad is a multiple of 32: the
    // supports of output rows [32k, 32k + 32) of the chunk then complete with one batch (one column pass per 32 output
    // rows instead of two partial ones).
    const int pad = (TB64 - (2 * w) % TB64) % TB64;
    const int nrows = ch + 2 * w + pad;    // ring rows of this chunk
    const int nb = (nrows + TB64 - 1) / TB64;
    const int nfw = a.wr + 2 * w;          // footprint rows of the window
    const size_t mp = (size_t)(a.strips * SW64);
    const float2 *mid = a.mid + (size_t)v * nfw * mp + c0;

    // batch b of intermediate rows → ring, asynchronously (cp.async, 16 bytes per lane and row); rows outside
    // the window's footprint are zeros, as the fused kernel stages them
    auto issue_batch = [&](int b) {
#pragma unroll
        for (int q = 0; q < TB64 / WARPS64; ++q) {
            const int f = b * TB64 + warp + q * WARPS64;
            const int fw = r0 + f - pad;                            // row of the window's footprint
            float2 *dst = s_ring + (size_t)(f % G.nring) * RPC + 2 * lane;
            if (f < nrows && fw >= 0 && fw < nfw) {
                const float2 *src = mid + (size_t)fw * mp + 2 * lane;
                const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
            } else {
                dst[0] = make_float2(0.f, 0.f); dst[1] = make_float2(0.f, 0.f);
            }
        }
    };

    // The column pass reads a few ring rows past the support of its outputs (tap counts are padded to multiples of 16,
    // rows are fetched ahead in runs of 8) and multiplies them by zero taps: whatever the shared memory held before this
    // CTA must not be NaN or infinity there → the ring starts zeroed, like the fused kernel's.
    for (int e = tid; e < G.nring * RPC; e += THREADS64) s_ring[e] = make_float2(0.f, 0.f);
    // Everything above touched nothing an earlier kernel writes.  From here on: the row kernel's intermediate — and, since
    // the kernels of a chain of steps run ahead of one another up to this point, the guess the PREVIOUS step's column
    // kernel published — are complete and visible.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int wy0, wx0;
    if (a.rect_mode) { wy0 = a.ry0; wx0 = a.rx0; }
    else { const int2 g = a.guess[v]; wy0 = g.x - 1 - a.rr; wx0 = g.y - 1 - a.rc; }
    __syncthreads();
    float best_v = -INFINITY;
    unsigned int best_i = 0xFFFFFFFFu;
    for (int b = 0; b < PFC; ++b) {
        if (b < nb) issue_batch(b);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int b = 0; b < nb; ++b) {
        asm volatile("cp.async.wait_group %0;" ::"n"(PFC - 1) : "memory");   // batch b has landed (this thread's copies)
        __syncthreads();                     // … everybody's; and the column pass of batch b − 1 has finished reading
        if (b + PFC < nb) issue_batch(b + PFC);      // overwrites ring rows older than anything batch b's pass reads
        asm volatile("cp.async.commit_group;" ::: "memory");
        col_pass64<DELTA, RPC>(a, wt, G, s_ring, b, w, ch, sw, c0, r0, v, warp, lane, best_v, best_i, pad);
    }
    merge_and_publish64(a, v, wy0, wx0, best_v, best_i, s_best, tid, warp, lane);
}

// dog_cols_wide2: the same with TWO teams of 8 warps per CTA working on two consecutive 32-row output batches at once —
// for kernels so long that the ring fills an SM (l = 245: one 8-warp CTA per SM, FMA pipe 66 % busy, 12 % of the warp
// slots).  Intermediate rows arrive in groups of 64; group k + 1 is in flight while the teams run the passes of group k:
// the ring holds the passes' support (nring rows) + 96.
__host__ __device__ inline int cols2_ring_rows(int L, int CH)
{
    const int w = L / 2, pad = (TB64 - (2 * w) % TB64) % TB64;
    const int all = ((CH + 2 * w + pad + 2 * TB64 - 1) / (2 * TB64)) * (2 * TB64);
    const int ring = geom64(L).nring + 3 * TB64;
    return all < ring ? all : ring;
}

template <int DELTA>
__global__ void __launch_bounds__(2 * THREADS64, 1)
dog_cols_wide2(const __grid_constant__ WinArgs a, const __grid_constant__ WideTaps wt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = a.L, w = a.w;
    Geom64 G = geom64(L);
    G.nring = cols2_ring_rows(L, a.CH);
    float2 *s_ring = reinterpret_cast<float2 *>(smem_raw);               // [nring][RPC], slot = ring row mod nring
    __shared__ unsigned long long s_best[2 * WARPS64];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, lane = tid & 31, warp16 = tid >> 5, team = warp16 >> 3, warp = warp16 & 7;
    int v, strip, chunk;
    {
        const int per_full = a.strips - 1, nc = (int)gridDim.x / a.strips;     // longest-first order, as the fused kernel
        const int id = (int)blockIdx.x;
        if (id < per_full * nc) { strip = id % per_full; const int r = id / per_full; chunk = r % a.chunks; v = r / a.chunks; }
        else { strip = a.strips - 1; const int r = id - per_full * nc; chunk = r % a.chunks; v = r / a.chunks; }
    }
    const int c0 = strip * SW64;
    const int r0 = chunk * a.CH;
    const int ch = min(a.CH, a.wr - r0);
    const int sw = min(SW64, a.wc - c0);
    const int pad = (TB64 - (2 * w) % TB64) % TB64;        // 2w + pad is a multiple of 32 (see dog_cols_wide)
    const int nrows = ch + 2 * w + pad;
    const int nb = (nrows + TB64 - 1) / TB64, ngroups = (nb + 1) / 2;
    const int nfw = a.wr + 2 * w;
    const size_t mp = (size_t)(a.strips * SW64);
    const float2 *mid = a.mid + (size_t)v * nfw * mp + c0;

    auto issue_group = [&](int g) {                         // 64 intermediate rows → ring: 4 rows per warp
#pragma unroll
        for (int q = 0; q < 2 * TB64 / (2 * WARPS64); ++q) {
            const int f = g * 2 * TB64 + warp16 + q * 2 * WARPS64;
            const int fw = r0 + f - pad;
            float2 *dst = s_ring + (size_t)(f % G.nring) * RPC + 2 * lane;
            if (f < nrows && fw >= 0 && fw < nfw) {
                const float2 *src = mid + (size_t)fw * mp + 2 * lane;
                const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
            } else {
                dst[0] = make_float2(0.f, 0.f); dst[1] = make_float2(0.f, 0.f);
            }
        }
    };

    for (int e = tid; e < G.nring * RPC; e += 2 * THREADS64) s_ring[e] = make_float2(0.f, 0.f);   // (see dog_cols_wide)
    asm volatile("griddepcontrol.wait;" ::: "memory");       // (see dog_cols_wide: nothing written by an earlier kernel is read above)
    int wy0, wx0;
    if (a.rect_mode) { wy0 = a.ry0; wx0 = a.rx0; }
    else { const int2 g = a.guess[v]; wy0 = g.x - 1 - a.rr; wx0 = g.y - 1 - a.rc; }
    __syncthreads();
    float best_v = -INFINITY;
    unsigned int best_i = 0xFFFFFFFFu;
    // fill-up: nothing is read yet, so every group that fits the ring without wrapping is requested at once
    int issued = min(G.nring / (2 * TB64), ngroups);
    for (int g = 0; g < issued; ++g) issue_group(g);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int k = 0; k < ngroups; ++k) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");     // group k has landed (issued ≥ k + 1 here)
        __syncthreads();                     // … everybody's copies; and the passes of group k − 1 have finished reading
        if (issued <= k + 1 && issued < ngroups) { issue_group(issued); ++issued; }   // rows older than group k's support
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int b = 2 * k + team;
        if (b < nb) col_pass64<DELTA, RPC>(a, wt, G, s_ring, b, w, ch, sw, c0, r0, v, warp, lane, best_v, best_i, pad);
    }
    merge_and_publish64(a, v, wy0, wx0, best_v, best_i, s_best, tid, warp16, lane, 2 * WARPS64);
}

int wide_max_kernel_len() { return kMaxLWide; }
size_t wide_smem_bytes(int L) { return geom64(L).bytes; }
size_t wide_cols_smem_bytes(int L, int CH) { return (size_t)cols_ring_rows(L, CH) * RPC * sizeof(float2); }
size_t wide_cols2_smem_bytes(int L, int CH) { return (size_t)cols2_ring_rows(L, CH) * RPC * sizeof(float2); }
// two teams per CTA where one 8-warp CTA would have the SM to itself anyway and a chunk has at least two output batches
static bool use_cols2(int L, int CH, size_t smem_limit)
{
    return CH >= 2 * TB64 && 2 * (wide_cols_smem_bytes(L, CH) + 1024) > smem_limit && wide_cols2_smem_bytes(L, CH) + 1024 <= smem_limit;
}
// float2 elements of the two-phase intermediate of n windows
size_t wide_mid_elems(int L, int wr, int wc, int n)
{
    const int strips = (wc + SW64 - 1) / SW64;
    return (size_t)n * (size_t)(wr + 2 * (L / 2)) * (size_t)(strips * SW64);
}

cudaError_t wide_init_device()
{
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
#define PT_WIDE_OPTIN(k)                                                                                              \
    e = cudaFuncGetAttributes(&fa, k);                                                                                \
    if (e != cudaSuccess) return e;                                                                                   \
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);        \
    if (e != cudaSuccess) return e;
    PT_WIDE_OPTIN((dog_rect_argmax_wide<uint8_t, 0>))
    PT_WIDE_OPTIN((dog_rect_argmax_wide<uint8_t, 4>))
    PT_WIDE_OPTIN((dog_rect_argmax_wide<float, 0>))
    PT_WIDE_OPTIN((dog_rect_argmax_wide<float, 4>))
    PT_WIDE_OPTIN((dog_rows_wide<uint8_t>))
    PT_WIDE_OPTIN((dog_rows_wide<float>))
    PT_WIDE_OPTIN((dog_cols_wide<0>))
    PT_WIDE_OPTIN((dog_cols_wide<4>))
    PT_WIDE_OPTIN((dog_cols_wide2<0>))
    PT_WIDE_OPTIN((dog_cols_wide2<4>))
#undef PT_WIDE_OPTIN
    return cudaSuccess;
}

// a.strips must count 64-column strips (decompose() in pt_api.cu does that when the wide kernel is chosen)
cudaError_t launch_wide(const WinArgs &a, int n, int pixel, cudaStream_t s)
{
    const size_t smem = wide_smem_bytes(a.L);
    dim3 grid((unsigned)(a.strips * a.chunks * n));
    const int L = a.L, w = L / 2, delta = (L - 1) & 7;
    if ((delta != 0 && delta != 4) || L > kMaxLWide || !a.h_taps) return cudaErrorInvalidValue;   // l = 4k + 1 always (Kernel.DoG)
    // taps as the kernels consume them (a.h_taps: row narrow | row wide | col narrow | col wide, each L)
    WideTaps wt;
    const float *rp = a.h_taps, *rm = a.h_taps + L, *cp = a.h_taps + 2 * L, *cm = a.h_taps + 3 * L;
    auto at = [L](const float *t, int k) { return (k >= 0 && k < L) ? t[k] : 0.f; };
    for (int q = 0; q < kMaxLq16; ++q) wt.cq[q] = make_float4(at(cp, q), at(cp, q - 1), at(cm, q), at(cm, q - 1));
    for (int d = 0; d <= kMaxW16; ++d) wt.trow[d] = d <= w ? make_float2(rp[w + d], rm[w + d]) : make_float2(0.f, 0.f);
    if (a.mid) {
        // two-phase: every footprint batch row-filtered once, then the column pass per (chunk, strip)
        const Geom64 G = geom64(L);
        const int nbt = (a.wr + 2 * w + TB64 - 1) / TB64;
        const size_t smem_rows = (size_t)TB64 * G.pin * sizeof(float), smem_cols = wide_cols_smem_bytes(L, a.CH);
        dim3 grid_rows((unsigned)(a.strips * nbt * n));
        cudaLaunchAttribute pdl[1];
        pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t lr = {};
        lr.gridDim = grid_rows; lr.blockDim = dim3(THREADS64); lr.dynamicSmemBytes = smem_rows; lr.stream = s;
        lr.attrs = pdl; lr.numAttrs = 1;
        cudaError_t e = pixel == 0 ? cudaLaunchKernelEx(&lr, dog_rows_wide<uint8_t>, a, wt) : cudaLaunchKernelEx(&lr, dog_rows_wide<float>, a, wt);
        if (e != cudaSuccess) return e;
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaLaunchConfig_t lc = {};
        lc.gridDim = grid; lc.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = 1;
        if (a.cols_teams != 1 && use_cols2(L, a.CH, (size_t)optin)) {
            lc.blockDim = dim3(2 * THREADS64); lc.dynamicSmemBytes = wide_cols2_smem_bytes(L, a.CH);
            return delta == 0 ? cudaLaunchKernelEx(&lc, dog_cols_wide2<0>, a, wt) : cudaLaunchKernelEx(&lc, dog_cols_wide2<4>, a, wt);
        }
        lc.blockDim = dim3(THREADS64); lc.dynamicSmemBytes = smem_cols;
        return delta == 0 ? cudaLaunchKernelEx(&lc, dog_cols_wide<0>, a, wt) : cudaLaunchKernelEx(&lc, dog_cols_wide<4>, a, wt);
    }
    if (pixel == 0) {
        if (delta == 0) dog_rect_argmax_wide<uint8_t, 0><<<grid, THREADS64, smem, s>>>(a, wt);
        else dog_rect_argmax_wide<uint8_t, 4><<<grid, THREADS64, smem, s>>>(a, wt);
    } else {
        if (delta == 0) dog_rect_argmax_wide<float, 0><<<grid, THREADS64, smem, s>>>(a, wt);
        else dog_rect_argmax_wide<float, 4><<<grid, THREADS64, smem, s>>>(a, wt);
    }
    return cudaGetLastError();
}

} // namespace pt
