// pt_api.cu — the C ABI of libpawsome_cuda.so (include/pawsome.h): handles,
// HBM frame stores, pinned staging, streams, and the host-side loops that
// drive the kernels.  No torch types, no CPU fallback: every compute entry
// point launches CUDA kernels or fails.
#include "../../include/pawsome.h"
#include "pt_kernels.cuh"

#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct pt_batch;
static void sync_batch_streams(pt_batch *b);                   // every stream this batch has work on (not the whole device)

namespace {

thread_local std::string g_err = "";

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(PT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

// ---- scalar arithmetic of the Tracker constructor (src/PawsomeTracker.jl:30,41-45)
double sigma_of(double tw) { return tw / (2.0 * std::sqrt(2.0 * std::log(2.0))); }
int kernel_len_of(double tw) { return 4 * (int)std::ceil(sigma_of(tw) * std::sqrt(2.0)) + 1; }

void gaussian_factor(double sigma, int l, std::vector<double> &g)
{
    const int w = l / 2;
    g.resize(l);
    double s = 0.0;
    for (int k = 0; k < l; ++k) {
        const double x = (double)(k - w);
        g[k] = std::exp(-(x * x) / (2.0 * sigma * sigma));
        s += g[k];
    }
    for (int k = 0; k < l; ++k) g[k] /= s;
}

// FP32 taps exactly as the kernels consume them.  The row pass works on
// (pixel − fill) in raw pixel units, so for u8 frames the N0f8 scale 1/255 is
// folded into the row taps (in double, before rounding to float).
void make_taps(double tw, bool darker, int pixel, std::vector<float> &rp, std::vector<float> &rm,
               std::vector<float> &cp, std::vector<float> &cm)
{
    const int l = kernel_len_of(tw);
    std::vector<double> gp, gm;
    gaussian_factor(sigma_of(tw), l, gp);
    gaussian_factor(sigma_of(tw) * std::sqrt(2.0), l, gm);
    const double scale = pixel == PT_PIX_U8 ? 1.0 / 255.0 : 1.0;
    const double dir = darker ? -1.0 : 1.0;
    rp.resize(l); rm.resize(l); cp.resize(l); cm.resize(l);
    for (int k = 0; k < l; ++k) {
        rp[k] = (float)(gp[k] * scale);
        rm[k] = (float)(gm[k] * scale);
        cp[k] = (float)(dir * gp[k]);
        cm[k] = (float)(-dir * gm[k]);
    }
}

size_t px_size(int pixel) { return pixel == PT_PIX_U8 ? 1 : 4; }

// Process-wide caches of freed buffers.  `track()` builds (and the auto-detect start builds twice) a Tracker per
// call, the reference's own pattern (src/PawsomeTracker.jl:94,103-105); page-locking and unlocking a staging
// buffer costs milliseconds each, more than tracking a short clip, so released buffers are parked here (bounded)
// and handed to the next batch.  Capacities are rounded up to powers of two so they are reusable.
struct BufPool {
    std::mutex m;
    std::multimap<std::pair<int, size_t>, void *> free_;      // (device or -1 for pinned, capacity) → pointer
    size_t held = 0;
    static size_t round_cap(size_t bytes)
    {
        if (bytes > (64u << 20)) return bytes;                 // large buffers are not pooled: exact size
        size_t c = 4096;
        while (c < bytes) c <<= 1;
        return c;
    }
    void *get(int key, size_t cap)
    {
        std::lock_guard<std::mutex> g(m);
        auto it = free_.find(std::make_pair(key, cap));
        if (it == free_.end()) return nullptr;
        void *p = it->second;
        free_.erase(it);
        held -= cap;
        return p;
    }
    bool put(int key, size_t cap, void *p)
    {
        std::lock_guard<std::mutex> g(m);
        if (cap > (64u << 20) || held + cap > (512u << 20)) return false;
        free_.emplace(std::make_pair(key, cap), p);
        held += cap;
        return true;
    }
};
BufPool &pool() { static BufPool *p = new BufPool(); return *p; }     // never destroyed: no CUDA calls at exit

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, pt_batch *owner)
    {
        if (bytes <= cap) return PT_OK;
        if (p) sync_batch_streams(owner);                      // regrow: nothing in flight may still use the old buffer
        release();
        const size_t c = BufPool::round_cap(bytes);
        p = pool().get(-1, c);
        if (!p) CU(cudaHostAlloc(&p, c, cudaHostAllocDefault));
        cap = c;
        return PT_OK;
    }
    void release()
    {
        if (p && !pool().put(-1, cap, p)) cudaFreeHost(p);
        p = nullptr; cap = 0;
    }
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int dev = 0;
    int ensure(size_t bytes, pt_batch *owner)
    {
        if (bytes <= cap) return PT_OK;
        if (p) sync_batch_streams(owner);                      // regrow: nothing in flight may still use the old buffer
        release();
        const size_t c = BufPool::round_cap(bytes);
        CU(cudaGetDevice(&dev));
        p = pool().get(dev, c);
        if (!p) CU(cudaMalloc(&p, c));
        cap = c;
        return PT_OK;
    }
    void release()
    {
        if (p && !pool().put(dev, cap, p)) cudaFree(p);
        p = nullptr; cap = 0;
    }
};

// Range of the allocation containing p, through the driver entry point (no libcuda link:
// the library must still load on a machine without a driver).  false = unknown.
bool alloc_range(const void *p, const char **base, size_t *size)
{
    typedef CUresult (*attr_fn)(void *, CUpointer_attribute, CUdeviceptr);
    static attr_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (attr_fn)sym;
        else
            cudaGetLastError();
    });
    if (!fn) return false;
    CUdeviceptr start = 0;
    size_t sz = 0;
    if (fn(&start, CU_POINTER_ATTRIBUTE_RANGE_START_ADDR, (CUdeviceptr)(uintptr_t)p) != CUDA_SUCCESS) return false;
    if (fn(&sz, CU_POINTER_ATTRIBUTE_RANGE_SIZE, (CUdeviceptr)(uintptr_t)p) != CUDA_SUCCESS) return false;
    *base = (const char *)(uintptr_t)start;
    *size = sz;
    return sz > 0;
}

// Defaults of the per-batch knobs from the PT_* environment variables — read ONCE per process; a batch copies
// them at creation (pt_batch_set_option changes them per handle).  Nothing on a per-call path calls getenv.
const pt::Cfg &defaults_from_env()
{
    static pt::Cfg cfg;
    static std::once_flag once;
    std::call_once(once, [] {
        auto geti = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
        cfg.window45 = getenv("PT_DISABLE_WINDOW45") ? 0 : 1;
        cfg.rect45 = getenv("PT_DISABLE_RECT45") ? 0 : 1;
        cfg.rot = geti("PT_W45_ROT", 1);
        cfg.rot_stride = std::max(0, geti("PT_W45_ROT_STRIDE", 0));
        cfg.skew = geti("PT_W45_SKEW", 1);
        cfg.r45_chunks = std::max(0, geti("PT_R45_CHUNKS", 0));
        cfg.generic_target = std::max(1, geti("PT_GENERIC_TARGET", 4 * 148));
        cfg.mode_slow = getenv("PT_MODE_SLOW") ? 1 : 0;
        cfg.zero_copy = getenv("PT_NO_ZEROCOPY") ? 0 : 1;
        cfg.host_lanes = std::max(0, geti("PT_HOST_LANES", 0));
        cfg.cluster = geti("PT_W45_CLUSTER", 0);
        cfg.bulk = geti("PT_W45_BULK", 1) ? 1 : 0;
        cfg.wide = geti("PT_GENERIC_WIDE", 1) ? 1 : 0;
        cfg.two_phase = geti("PT_WIDE_TWO_PHASE", 1);
        cfg.cols_teams = geti("PT_WIDE_COLS_TEAMS", 0);
        cfg.crop_gather = geti("PT_CROP_GATHER", 1) ? 1 : 0;
        cfg.cols_ch = std::max(0, geti("PT_WIDE_COLS_CH", 0)) / 32 * 32;
    });
    return cfg;
}

// Per-device one-time setup (dynamic shared memory opt-in of every kernel is a per-device attribute; SM count).
constexpr int kMaxDevices = 64;
struct DeviceInfo { bool ready = false; int sms = 0; };
int device_info(int device, DeviceInfo *out)
{
    static std::mutex m;
    static DeviceInfo info[kMaxDevices];
    if (device < 0 || device >= kMaxDevices) return fail(PT_ERR_ARG, "device %d out of range", device);
    std::lock_guard<std::mutex> g(m);
    DeviceInfo &d = info[device];
    if (!d.ready) {
        CU(cudaSetDevice(device));
        int sms = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        CU(pt::generic_init_device());
        CU(pt::wide_init_device());
        CU(pt::window45_init_device());
        d.sms = sms > 0 ? sms : 148;
        d.ready = true;
    }
    *out = d;
    return PT_OK;
}

bool is_pinned_or_device(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

} // namespace

// One worker of the footprint-streaming host loop: owns a contiguous slice of
// the batch's videos, a stream, and pinned staging for crops and results.
struct pt_lane {
    cudaStream_t stream = nullptr;
    PinnedBuf h_crops, h_res;
    DevBuf d_crops;
    int v0 = 0, v1 = 0, index = 0;
    float2 *d_mid = nullptr;             // slice of pt_batch::d_mid for this lane's windows (two-phase wide path)
};

struct pt_batch {
    int n = 0, H = 0, W = 0;
    double tw = 0;
    int ws_r = 0, ws_c = 0, rr = 0, rc = 0, wr = 0, wc = 0;
    int darker = 1, pixel = PT_PIX_U8, device = 0;
    int L = 0, w = 0, Lpad = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    DevBuf d_arena;                      // backs every small d_* pointer below (one allocation per batch)
    float2 *d_taps_row = nullptr, *d_taps_col = nullptr;
    float *d_fill = nullptr;
    int *d_fill_i = nullptr;
    bool fill_set = false;
    std::vector<int> h_fill;
    std::vector<float> h_taps;           // host copy: row narrow | row wide | col narrow | col wide, each L
    int2 *d_guess = nullptr;
    bool guess_set = false;
    // host mirror of the chain state: valid ⇒ h_guess equals what the next step must start from;
    // d_guess_stale ⇒ the device copy is older than the mirror and is refreshed before any device-side use
    std::vector<int32_t> h_guess;
    bool h_guess_valid = false, d_guess_stale = false;
    int center_key[3] = {-1, -1, -1};    // (rr, rc, w) d_center was last filled for
    bool staging_busy = false;           // an async H2D from the guess staging may be in flight …
    cudaStream_t staging_stream = nullptr;   // … on this stream
    int2 *d_center = nullptr;            // crop-mode guess: centre of the footprint
    unsigned long long *d_keys = nullptr;
    unsigned int *d_counters = nullptr;
    unsigned int *d_hist = nullptr;
    unsigned int *d_xflag = nullptr;     // [n] hand-off flags + [3] handshake words of the rotating-slot kernel (zero between launches)
    int2 *d_xpos = nullptr;              // [n] hand-off guesses
    int4 *d_pos = nullptr;
    float *d_resp = nullptr;
    // own HBM frame store (two slots so a whole-frame upload can overlap a step)
    DevBuf d_frames[2];
    size_t own_pitch = 0, own_stride = 0;
    int cur_slot = 0;
    bool have_frames = false;
    // frames bound from caller-owned HBM
    const void *bound_base = nullptr;
    size_t bound_stride = 0, bound_pitch = 0;
    // trajectory buffer for chained steps
    DevBuf d_traj_pos, d_traj_resp, d_map, d_ptrs;
    PinnedBuf h_traj, h_ptrs;
    PinnedBuf h_stage[2], h_out;
    std::vector<pt_lane> lanes;
    long long launches = 0;
    const char *last_kernel = "";        // name of the kernel the most recent launch_step ran
    pt::Cfg cfg;                         // tuning / debugging knobs (defaults_from_env at create, pt_batch_set_option)
    cudaStream_t ext_stream = nullptr;   // caller stream of the most recent pt_batch_track_device_async (may still run)
    PinnedBuf h_small;                   // pinned staging for the small synchronous uploads (fills, centres, taps)
    DevBuf d_mid;                        // row-pass intermediate of the two-phase wide path ([n] windows), grown on demand
    DevBuf d_crop, d_cropmeta;           // per-window footprint crops of page-locked host frames (gather_footprints) + their origin / guess
};

struct pt_tracker {
    pt_batch *b = nullptr;
};

static void sync_batch_streams(pt_batch *b)
{
    if (!b) return;
    if (b->stream) cudaStreamSynchronize(b->stream);
    if (b->copy_stream) cudaStreamSynchronize(b->copy_stream);
    if (b->ext_stream) { cudaStreamSynchronize(b->ext_stream); b->ext_stream = nullptr; }
    for (auto &ln : b->lanes) if (ln.stream) cudaStreamSynchronize(ln.stream);
}

namespace {

int set_device(const pt_batch *b) { CU(cudaSetDevice(b->device)); return PT_OK; }

void configure_window(pt_batch *b, int ws_r, int ws_c)
{
    b->ws_r = ws_r; b->ws_c = ws_c;
    b->rr = ws_r / 2; b->rc = ws_c / 2;           // radii = window_size .÷ 2 (:44)
    b->wr = 2 * b->rr + 1; b->wc = 2 * b->rc + 1; // guess .- radii : guess .+ radii (:56)
}

// Fill the kernel argument block for one step.
pt::WinArgs make_args(pt_batch *b, const void *frames, size_t stride, size_t pitch, int H, int W,
                      const int2 *guess, int nwin)
{
    pt::WinArgs a;
    memset(&a, 0, sizeof a);
    a.frames = frames; a.frame_stride = stride; a.pitch = (int)pitch; a.H = H; a.W = W;
    a.fill = b->d_fill; a.guess = guess;
    a.rect_mode = 0;
    a.rr = b->rr; a.rc = b->rc; a.wr = b->wr; a.wc = b->wc;
    a.L = b->L; a.w = b->w; a.Lpad = b->Lpad;
    a.taps_row = b->d_taps_row; a.taps_col = b->d_taps_col;
    a.keys = b->d_keys; a.counters = b->d_counters; a.tickets = b->d_counters + b->n;
    a.out_pos = b->d_pos; a.out_resp = b->d_resp;
    a.next_guess = nullptr; a.traj_pos = nullptr; a.traj_resp = nullptr; a.map_out = nullptr;
    a.T = 1; a.step_stride = 0;
    a.h_taps = b->h_taps.data();
    a.frame_ptrs = nullptr;
    a.xflag = (nwin == b->n) ? b->d_xflag : nullptr;     // whole-batch launches only (one stream at a time)
    a.xpos = b->d_xpos;
    a.mid = nullptr;
    a.crop_org = nullptr; a.Hreal = H; a.Wreal = W;
    a.host_frames = 0;
    a.cols_teams = b->cfg.cols_teams;
    (void)nwin;
    return a;
}

// the 64-column kernel: windows wider than one 32-column strip whose kernel length fits its shared-memory footprint
bool use_wide_kernel(const pt_batch *b, const pt::WinArgs &a)
{
    return b->cfg.wide && a.wc > pt::kTileCols && a.L <= pt::wide_max_kernel_len() &&
           pt::wide_smem_bytes(a.L) + 512 <= (size_t)b->cfg.smem_optin;
}

void decompose(pt::WinArgs &a, int nwin, int target)
{
    // Strips of 32 output columns; row chunks only where a launch would otherwise leave the GPU
    // under-occupied.  Each extra chunk repeats 2w footprint rows of the row pass, so chunks are
    // added until ~kTarget CTAs exist (measured best for the 1080p full-frame shape) and never below
    // one batch of rows.
    const int tile_cols = a.wide ? 2 * pt::kTileCols : pt::kTileCols;
    a.strips = (a.wc + tile_cols - 1) / tile_cols;
    if (a.wide) target = std::max(1, target / 4);       // one 256-thread CTA per SM (or two) instead of four 128-thread ones
    const int total = nwin * a.strips;
    int chunks = 1;
    if (total < target) {
        chunks = std::max(1, target / total);         // never more CTAs than one resident wave (no tail wave)
        const int maxc = (a.wr + pt::kBatchRows - 1) / pt::kBatchRows;
        chunks = std::max(1, std::min(chunks, maxc));
    }
    a.CH = (a.wr + chunks - 1) / chunks;
    a.chunks = (a.wr + a.CH - 1) / a.CH;
}

// Row chunks of the column-pass launch of the two-phase wide path.  Chunks cost no arithmetic here, only a re-read of
// 2w intermediate rows from L2 each; a CTA's cost ≈ copies (≈ 1000 cycles per 32 rows = 16 KB when every SM pulls from L2:
// 64 windows of 401x401 in 32-row chunks moved 859 MB and took 410 µs) + column pass (32·Lq16 cycles per
// 32 output rows), CTAs are uniform, so the launch costs whole waves of them.
void decompose_cols(pt::WinArgs &a, int nwin, int sms, int forced_ch = 0)
{
    a.strips = (a.wc + 2 * pt::kTileCols - 1) / (2 * pt::kTileCols);
    const int nbo = (a.wr + pt::kBatchRows - 1) / pt::kBatchRows;
    const double colc = 32.0 * (double)(((a.L + 1 + 15) / 16) * 16), cpy = 1000.0;
    double best = 1e300;
    int best_k = 1;
    for (int k = 1; k <= nbo; ++k) {
        const int CH = k * pt::kBatchRows, chunks = (a.wr + CH - 1) / CH;
        const long long ctas = (long long)nwin * a.strips * chunks;
        const double waves = (double)((ctas + sms - 1) / sms);
        const double cost = waves * (cpy * (double)((CH + 2 * a.w + pt::kBatchRows - 1) / pt::kBatchRows) + colc * k);
        if (cost < best - 1e-9) { best = cost; best_k = k; }
    }
    a.CH = best_k * pt::kBatchRows;
    if (forced_ch >= pt::kBatchRows) a.CH = std::min(forced_ch / pt::kBatchRows, nbo) * pt::kBatchRows;
    a.chunks = (a.wr + a.CH - 1) / a.CH;
}

// Two launches (row pass of every footprint batch once, then the column pass) instead of the fused wide kernel?
// Fused, a launch that cannot fill the GPU with whole strips cuts them into row chunks and repeats 2w footprint rows
// of the row pass per chunk; two-phase, CTAs are small and uniform (no wave quantisation of one-CTA-per-SM strips).
bool want_two_phase(const pt_batch *b, const pt::WinArgs &a, int nwin)
{
    if (!a.mid || b->cfg.two_phase == 0 || pt::wide_cols_smem_bytes(a.L, 1 << 20) + 512 > (size_t)b->cfg.smem_optin) return false;
    // measured (tools/config4_timing.py, tools/tw_sweep.py): faster than the fused kernel for every batch size — one
    // 401x401 window at l = 245: 25 vs 74 µs, 64 of them: 690 vs 725 µs, 256 windows at tw = 70: 225 vs 293 µs
    (void)nwin;
    return true;
}

// l = 65 rectangles: the marching tile kernel, unless the rectangle is at most two tiles high (nothing to march: every
// tile row-filters its whole footprint) and the launch has many of them — then the two-phase wide path does less work
// (256 windows of 49x49 at tw = 26: 25.9 vs 38.2 µs per step; 16 of them: 15.1 vs 12.7, the tile kernel stays).
bool takes_rect45(const pt_batch *b, const pt::WinArgs &a, int nwin)
{
    if (!(b->cfg.window45 && pt::rect45_supported(a, b->cfg, b->pixel))) return false;
    const int nty = (a.wr + 44) / 45, ntx = (a.wc + 44) / 45;
    const bool many_flat = nty <= 2 && (long long)nwin * ntx * nty >= 2LL * b->cfg.sms;
    if (many_flat && !a.map_out && b->cfg.two_phase != 0 && use_wide_kernel(b, a) &&
        pt::wide_cols_smem_bytes(a.L, 1 << 20) + 512 <= (size_t)b->cfg.smem_optin)
        return false;
    return true;
}

// Kernel choice shared by every path (so host-footprint, resident and per-step calls round identically).
// Thread-safe: touches no batch state.
cudaError_t launch_windows(const pt_batch *b, pt::WinArgs &a, int nwin, cudaStream_t s)
{
    if (b->cfg.window45 && !a.rect_mode && !a.map_out && !a.crop_org && pt::window45_supported(a, b->pixel))
        return pt::launch_window45(a, b->cfg, nwin, b->pixel, s);
    if (takes_rect45(b, a, nwin))
        return pt::launch_rect45(a, b->cfg, nwin, b->pixel, s);
    a.wide = use_wide_kernel(b, a) ? 1 : 0;
    if (!(a.wide && want_two_phase(b, a, nwin))) a.mid = nullptr;
    if (a.mid) decompose_cols(a, nwin, b->cfg.sms, b->cfg.cols_ch);
    else decompose(a, nwin, b->cfg.generic_target);
    return a.wide ? pt::launch_wide(a, nwin, b->pixel, s) : pt::launch_generic(a, nwin, b->pixel, s);
}

// Grow the two-phase intermediate for nwin windows of a's geometry (no-op for the other kernels); leaves a.mid null
// when the buffer would be unreasonably large.
int ensure_mid(pt_batch *b, pt::WinArgs &a, int nwin, size_t window_offset = 0)
{
    a.mid = nullptr;
    if (b->cfg.two_phase == 0 || (b->cfg.window45 && !a.rect_mode && !a.map_out && !a.crop_org && pt::window45_supported(a, b->pixel)) ||
        takes_rect45(b, a, nwin) || !use_wide_kernel(b, a))
        return PT_OK;
    const size_t per = pt::wide_mid_elems(a.L, a.wr, a.wc, 1);
    const size_t bytes = per * sizeof(float2) * (window_offset + (size_t)nwin);
    if (bytes > ((size_t)2 << 30)) return PT_OK;
    int rc = b->d_mid.ensure(bytes, b);
    if (rc) return rc;
    a.mid = (float2 *)b->d_mid.p + per * window_offset;
    return PT_OK;
}

int launch_step(pt_batch *b, pt::WinArgs &a, int nwin, cudaStream_t s)
{
    int rcm = ensure_mid(b, a, nwin);
    if (rcm) return rcm;
    const bool two = a.mid != nullptr && use_wide_kernel(b, a) && want_two_phase(b, a, nwin);
    const cudaError_t e = launch_windows(b, a, nwin, s);
    if (b->cfg.window45 && !a.rect_mode && !a.map_out && !a.crop_org && pt::window45_supported(a, b->pixel))
        b->last_kernel = pt::window45_kernel_for(a, b->cfg, nwin, b->pixel);
    else if (takes_rect45(b, a, nwin)) b->last_kernel = pt::rect45_name();
    else if (use_wide_kernel(b, a)) b->last_kernel = two ? (b->pixel == PT_PIX_U8 ? "dog_rows_wide<u8>+dog_cols_wide" : "dog_rows_wide<f32>+dog_cols_wide")
                                                        : (b->pixel == PT_PIX_U8 ? "dog_rect_argmax_wide<u8>" : "dog_rect_argmax_wide<f32>");
    else b->last_kernel = b->pixel == PT_PIX_U8 ? "dog_rect_argmax_generic<u8>" : "dog_rect_argmax_generic<f32>";
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    b->launches += two ? 2 : 1;
    return PT_OK;
}

int current_frames(pt_batch *b, const void **base, size_t *stride, size_t *pitch)
{
    if (b->bound_base) { *base = b->bound_base; *stride = b->bound_stride; *pitch = b->bound_pitch; return PT_OK; }
    if (b->have_frames) { *base = b->d_frames[b->cur_slot].p; *stride = b->own_stride; *pitch = b->own_pitch; return PT_OK; }
    return fail(PT_ERR_STATE, "no frame attached: call pt_batch_set_frames or pt_batch_bind_device_frames first");
}

// Small host → device upload ordered on the batch's own stream: through pinned staging (a pageable cudaMemcpy on the
// legacy stream may return before the DMA lands and is not ordered against the non-blocking streams the kernels
// run on), then one stream synchronise so the staging can be reused and every later launch — on any stream — sees it.
int upload_small(pt_batch *b, void *dst, const void *src, size_t bytes)
{
    int rc = b->h_small.ensure(bytes, b);
    if (rc) return rc;
    memcpy(b->h_small.p, src, bytes);
    CU(cudaMemcpyAsync(dst, b->h_small.p, bytes, cudaMemcpyHostToDevice, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return PT_OK;
}

int upload_guess(pt_batch *b, const int32_t *g, cudaStream_t s)
{
    int rc = b->h_out.ensure((size_t)b->n * 32, b);
    if (rc) return rc;
    // staging region [n*16, n*24) of h_out is reserved for guesses
    int32_t *st = reinterpret_cast<int32_t *>((char *)b->h_out.p + (size_t)b->n * 24);
    if (b->staging_busy) {                                     // an earlier async copy may still be reading the staging
        CU(cudaStreamSynchronize(b->staging_stream ? b->staging_stream : s));
        b->staging_busy = false;
    }
    memcpy(st, g, sizeof(int32_t) * 2 * (size_t)b->n);
    CU(cudaMemcpyAsync(b->d_guess, st, sizeof(int2) * (size_t)b->n, cudaMemcpyHostToDevice, s));
    b->guess_set = true;
    b->staging_busy = true;
    b->staging_stream = s;
    if (g != b->h_guess.data()) b->h_guess.assign(g, g + 2 * (size_t)b->n);
    b->h_guess_valid = true;
    b->d_guess_stale = false;
    return PT_OK;
}

// Before any kernel reads d_guess: bring the device copy up to date with the host mirror.
int flush_guess(pt_batch *b, cudaStream_t s = nullptr)
{
    if (!b->d_guess_stale) return PT_OK;
    return upload_guess(b, b->h_guess.data(), s ? s : b->stream);
}

// After a step whose clamped results were read back into out_ij (n×2): they are the next chain state.
void mirror_from_results(pt_batch *b, const int32_t *out_ij)
{
    if (out_ij) { b->h_guess.assign(out_ij, out_ij + 2 * (size_t)b->n); b->h_guess_valid = true; }
    else b->h_guess_valid = false;
    b->d_guess_stale = false;
}

// Copy n host frames into d_frames[slot] on stream s, through pinned staging
// unless the caller's memory is already page-locked.
int upload_frames(pt_batch *b, const void *const *frames, size_t pitch, int slot, cudaStream_t s)
{
    const size_t es = px_size(b->pixel);
    const size_t row_bytes = (size_t)b->W * es;
    const size_t src_pitch_b = pitch * es;
    if (pitch < (size_t)b->W) return fail(PT_ERR_ARG, "pitch %zu smaller than W %d", pitch, b->W);
    int rc = b->d_frames[slot].ensure(b->own_stride * es * (size_t)b->n, b);
    if (rc) return rc;
    char *dbase = (char *)b->d_frames[slot].p;
    const size_t dst_pitch_b = b->own_pitch * es, dst_stride_b = b->own_stride * es;
    const bool pinned = is_pinned_or_device(frames[0]);
    if (pinned) {
        for (int v = 0; v < b->n; ++v)
            CU(cudaMemcpy2DAsync(dbase + dst_stride_b * v, dst_pitch_b, frames[v], src_pitch_b, row_bytes,
                                 (size_t)b->H, cudaMemcpyDefault, s));
        return PT_OK;
    }
    const size_t frame_bytes = row_bytes * (size_t)b->H;
    if (frame_bytes * (size_t)b->n <= (8u << 20) && b->h_stage[0].cap == 0) {
        // small upload (a single Tracker's frame): page-locking a staging buffer costs milliseconds, more than
        // letting the driver stage the pageable copy itself
        for (int v = 0; v < b->n; ++v)
            CU(cudaMemcpy2DAsync(dbase + dst_stride_b * v, dst_pitch_b, frames[v], src_pitch_b, row_bytes,
                                 (size_t)b->H, cudaMemcpyHostToDevice, s));
        return PT_OK;
    }
    // pageable source: pack rows into pinned staging (two halves, alternating) then DMA
    const size_t per_half = std::max<size_t>(1, std::min<size_t>((size_t)b->n, (64u << 20) / std::max<size_t>(frame_bytes, 1)));
    for (int h = 0; h < 2; ++h) { rc = b->h_stage[h].ensure(per_half * frame_bytes, b); if (rc) return rc; }
    int half = 0;
    for (int v0 = 0; v0 < b->n; v0 += (int)per_half, half ^= 1) {
        const int v1 = std::min(b->n, v0 + (int)per_half);
        CU(cudaEventSynchronize(b->ev_copy[half])); // staging half free again?
        char *st = (char *)b->h_stage[half].p;
        for (int v = v0; v < v1; ++v) {
            const char *src = (const char *)frames[v];
            char *dst = st + (size_t)(v - v0) * frame_bytes;
            if (src_pitch_b == row_bytes) memcpy(dst, src, frame_bytes);
            else for (int y = 0; y < b->H; ++y) memcpy(dst + (size_t)y * row_bytes, src + (size_t)y * src_pitch_b, row_bytes);
        }
        for (int v = v0; v < v1; ++v)
            CU(cudaMemcpy2DAsync(dbase + dst_stride_b * v, dst_pitch_b, st + (size_t)(v - v0) * frame_bytes,
                                 row_bytes, row_bytes, (size_t)b->H, cudaMemcpyHostToDevice, s));
        CU(cudaEventRecord(b->ev_copy[half], s));
    }
    return PT_OK;
}

int read_results(pt_batch *b, int32_t *out_ij, int32_t *out_raw, float *out_resp, cudaStream_t s)
{
    const size_t n = (size_t)b->n;
    int rc = b->h_out.ensure(n * 32, b);
    if (rc) return rc;
    int4 *hp = reinterpret_cast<int4 *>(b->h_out.p);
    float *hr = reinterpret_cast<float *>((char *)b->h_out.p + n * 16);
    CU(cudaMemcpyAsync(hp, b->d_pos, n * 16, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(hr, b->d_resp, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    b->staging_busy = false;
    for (size_t v = 0; v < n; ++v) {
        if (out_ij) { out_ij[2 * v] = hp[v].x; out_ij[2 * v + 1] = hp[v].y; }
        if (out_raw) { out_raw[2 * v] = hp[v].z; out_raw[2 * v + 1] = hp[v].w; }
        if (out_resp) out_resp[v] = hr[v];
    }
    return PT_OK;
}

} // namespace

// =============================================================================
extern "C" {

int pt_version(void) { return PT_VERSION; }
const char *pt_last_error(void) { return g_err.c_str(); }

int pt_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

double pt_sigma(double tw) { return sigma_of(tw); }
int pt_preferred_batch(int device)
{
    int ndev = 0, sms = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PT_ERR_ARG, "device %d out of range (have %d)", device, ndev);
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    return 2 * sms;
}

int pt_kernel_len(double tw) { return tw > 0 ? kernel_len_of(tw) : fail(PT_ERR_ARG, "target_width must be > 0"); }
int pt_default_window(double tw) { return 4 * (int)std::ceil(sigma_of(tw)) + 1; }

int pt_factors_f32(double tw, int darker, float *row_p, float *row_m, float *col_p, float *col_m)
{
    if (!(tw > 0)) return fail(PT_ERR_ARG, "target_width must be > 0");
    std::vector<float> rp, rm, cp, cm;
    make_taps(tw, darker != 0, PT_PIX_F32, rp, rm, cp, cm);
    const size_t l = rp.size();
    if (row_p) memcpy(row_p, rp.data(), l * 4);
    if (row_m) memcpy(row_m, rm.data(), l * 4);
    if (col_p) memcpy(col_p, cp.data(), l * 4);
    if (col_m) memcpy(col_m, cm.data(), l * 4);
    return (int)l;
}

int pt_batch_create(int n, int H, int W, double tw, int ws_rows, int ws_cols, int darker, int pixel,
                    int device, pt_batch **out)
{
    if (!out) return fail(PT_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (n < 1 || n > 65535) return fail(PT_ERR_ARG, "n must be in 1..65535 (got %d)", n);
    if (H < 1 || W < 1) return fail(PT_ERR_ARG, "frame size must be positive (got %dx%d)", H, W);
    if (!(tw > 0) || !std::isfinite(tw)) return fail(PT_ERR_ARG, "target_width must be a positive finite number");
    if (ws_rows < 1 || ws_cols < 1) return fail(PT_ERR_ARG, "window_size must be >= 1 (got %dx%d)", ws_rows, ws_cols);
    if (pixel != PT_PIX_U8 && pixel != PT_PIX_F32) return fail(PT_ERR_ARG, "unknown pixel type %d", pixel);
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PT_ERR_ARG, "device %d out of range (have %d)", device, ndev);

    const int L = kernel_len_of(tw);
    const int Lpad = ((L + pt::kTapChunk - 1) / pt::kTapChunk) * pt::kTapChunk;
    DeviceInfo dinfo;
    { const int rc0 = device_info(device, &dinfo); if (rc0) return rc0; }   // once per device: shared-memory opt-ins, SM count
    CU(cudaSetDevice(device));
    int smem_optin = 0;      // (cudaGetDeviceProperties costs ~20 ms per call; one attribute is microseconds)
    CU(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (pt::generic_smem_bytes(L, Lpad) + 256 > (size_t)smem_optin)      // (+ the kernel's static shared memory)
        return fail(PT_ERR_UNSUPPORTED, "kernel length %d (target_width %g) needs %zu B shared memory, device allows %zu",
                    L, tw, pt::generic_smem_bytes(L, Lpad), (size_t)smem_optin);

    pt_batch *b = new (std::nothrow) pt_batch();
    if (!b) return fail(PT_ERR_NOMEM, "out of host memory");
    b->n = n; b->H = H; b->W = W; b->tw = tw; b->darker = darker != 0; b->pixel = pixel; b->device = device;
    b->L = L; b->w = L / 2; b->Lpad = Lpad;
    b->cfg = defaults_from_env();
    b->cfg.sms = dinfo.sms;
    b->cfg.smem_optin = smem_optin;
    configure_window(b, ws_rows, ws_cols);
    b->own_pitch = pixel == PT_PIX_U8 ? (((size_t)W + 15) & ~(size_t)15) : (((size_t)W + 3) & ~(size_t)3);
    b->own_stride = b->own_pitch * (size_t)H;

    std::vector<float> rp, rm, cp, cm;
    make_taps(tw, b->darker, pixel, rp, rm, cp, cm);
    std::vector<float2> trow(Lpad, make_float2(0.f, 0.f)), tcol(Lpad, make_float2(0.f, 0.f));
    for (int k = 0; k < L; ++k) { trow[k] = make_float2(rp[k], rm[k]); tcol[k] = make_float2(cp[k], cm[k]); }
    b->h_taps.resize(4 * (size_t)L);
    for (int k = 0; k < L; ++k) {
        b->h_taps[k] = rp[k]; b->h_taps[L + k] = rm[k]; b->h_taps[2 * L + k] = cp[k]; b->h_taps[3 * L + k] = cm[k];
    }

    int rc = PT_OK;
    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == PT_OK) rc = fail(PT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    };
    cu(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    cu(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int i = 0; i < 2; ++i) {
        cu(cudaEventCreateWithFlags(&b->ev_copy[i], cudaEventDisableTiming), "cudaEventCreate");
        cu(cudaEventCreateWithFlags(&b->ev_done[i], cudaEventDisableTiming), "cudaEventCreate");
    }
    // one arena for all the small per-batch device buffers: one cudaMalloc, one cudaMemset, one cudaFree
    {
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        const size_t o_trow = take(sizeof(float2) * Lpad), o_tcol = take(sizeof(float2) * Lpad);
        const size_t o_fill = take(sizeof(float) * n), o_filli = take(sizeof(int) * n);
        const size_t o_guess = take(sizeof(int2) * n), o_center = take(sizeof(int2) * n);
        const size_t o_keys = take(sizeof(unsigned long long) * n);
        const size_t o_cnt = take(sizeof(unsigned int) * 3 * n);                       // [n] completion + [n][2] ticket scratch
        const size_t o_hist = take(sizeof(unsigned int) * pt::kModeScratch * (size_t)n);
        const size_t o_pos = take(sizeof(int4) * n), o_resp = take(sizeof(float) * n);
        const size_t o_xflag = take(sizeof(unsigned int) * ((size_t)n + 4)), o_xpos = take(sizeof(int2) * n);   // + 3 handshake words
        if (rc == PT_OK && b->d_arena.ensure(off, b) != PT_OK) rc = PT_ERR_CUDA;
        if (rc == PT_OK) {
            char *a0 = (char *)b->d_arena.p;
            b->d_taps_row = (float2 *)(a0 + o_trow); b->d_taps_col = (float2 *)(a0 + o_tcol);
            b->d_fill = (float *)(a0 + o_fill); b->d_fill_i = (int *)(a0 + o_filli);
            b->d_guess = (int2 *)(a0 + o_guess); b->d_center = (int2 *)(a0 + o_center);
            b->d_keys = (unsigned long long *)(a0 + o_keys); b->d_counters = (unsigned int *)(a0 + o_cnt);
            b->d_hist = (unsigned int *)(a0 + o_hist); b->d_pos = (int4 *)(a0 + o_pos); b->d_resp = (float *)(a0 + o_resp);
            b->d_xflag = (unsigned int *)(a0 + o_xflag); b->d_xpos = (int2 *)(a0 + o_xpos);
            cu(cudaMemsetAsync(b->d_arena.p, 0, off, b->stream), "memset");
            // taps: d_taps_row and d_taps_col are adjacent arena slots of equal padded size → one staged upload each
            if (rc == PT_OK && upload_small(b, b->d_taps_row, trow.data(), sizeof(float2) * Lpad) != PT_OK) rc = PT_ERR_CUDA;
            if (rc == PT_OK && upload_small(b, b->d_taps_col, tcol.data(), sizeof(float2) * Lpad) != PT_OK) rc = PT_ERR_CUDA;
        }
    }
    if (rc != PT_OK) { pt_batch_destroy(b); return rc; }
    b->h_fill.assign(n, 0);
    *out = b;
    return PT_OK;
}

void pt_batch_destroy(pt_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    sync_batch_streams(b);               // incl. a caller stream handed to pt_batch_track_device_async
    for (auto &ln : b->lanes) {
        if (ln.stream) cudaStreamDestroy(ln.stream);
        ln.h_crops.release(); ln.h_res.release(); ln.d_crops.release();
    }
    b->h_small.release();
    b->d_arena.release();
    for (int i = 0; i < 2; ++i) {
        b->d_frames[i].release(); b->h_stage[i].release();
        if (b->ev_copy[i]) cudaEventDestroy(b->ev_copy[i]);
        if (b->ev_done[i]) cudaEventDestroy(b->ev_done[i]);
    }
    b->d_traj_pos.release(); b->d_traj_resp.release(); b->d_map.release(); b->h_out.release(); b->d_mid.release(); b->d_crop.release(); b->d_cropmeta.release();
    b->d_ptrs.release(); b->h_traj.release(); b->h_ptrs.release();
    if (b->stream) cudaStreamDestroy(b->stream);
    if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
    delete b;
}

int pt_batch_set_window(pt_batch *b, int ws_rows, int ws_cols)
{
    if (!b) return fail(PT_ERR_ARG, "batch is NULL");
    if (ws_rows < 1 || ws_cols < 1) return fail(PT_ERR_ARG, "window_size must be >= 1 (got %dx%d)", ws_rows, ws_cols);
    configure_window(b, ws_rows, ws_cols);
    return PT_OK;
}

int pt_batch_set_frames(pt_batch *b, const void *const *frames, size_t pitch)
{
    if (!b || !frames) return fail(PT_ERR_ARG, "NULL argument");
    for (int v = 0; v < b->n; ++v) if (!frames[v]) return fail(PT_ERR_ARG, "frames[%d] is NULL", v);
    int rc = set_device(b);
    if (rc) return rc;
    // make sure no step still reads the slot we are about to overwrite
    if (b->ext_stream) { CU(cudaStreamSynchronize(b->ext_stream)); b->ext_stream = nullptr; }
    CU(cudaStreamSynchronize(b->stream));
    rc = upload_frames(b, frames, pitch, b->cur_slot, b->stream);
    if (rc) return rc;
    b->have_frames = true;
    b->bound_base = nullptr;
    return PT_OK;
}

int pt_batch_bind_device_frames(pt_batch *b, const void *dev_base, size_t frame_stride, size_t pitch)
{
    if (!b || !dev_base) return fail(PT_ERR_ARG, "NULL argument");
    if (pitch < (size_t)b->W) return fail(PT_ERR_ARG, "pitch %zu smaller than W %d", pitch, b->W);
    b->bound_base = dev_base; b->bound_stride = frame_stride; b->bound_pitch = pitch;
    return PT_OK;
}

int pt_batch_compute_fill(pt_batch *b, int *fills_out)
{
    if (!b) return fail(PT_ERR_ARG, "batch is NULL");
    int rc = set_device(b);
    if (rc) return rc;
    const void *base; size_t stride, pitch;
    rc = current_frames(b, &base, &stride, &pitch);
    if (rc) return rc;
    cudaError_t e = pt::launch_mode(base, stride, (int)pitch, b->H, b->W, b->n, b->pixel, b->d_hist,
                                    b->d_fill, b->d_fill_i, b->cfg.mode_slow != 0, b->stream);
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "mode kernel launch failed: %s", cudaGetErrorString(e));
    b->launches += 2;
    CU(cudaMemcpyAsync(b->h_fill.data(), b->d_fill_i, sizeof(int) * (size_t)b->n, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    b->fill_set = true;
    if (fills_out) memcpy(fills_out, b->h_fill.data(), sizeof(int) * (size_t)b->n);
    return PT_OK;
}

int pt_batch_set_fill(pt_batch *b, const int *fills)
{
    if (!b || !fills) return fail(PT_ERR_ARG, "NULL argument");
    int rc = set_device(b);
    if (rc) return rc;
    std::vector<float> f(b->n);
    for (int v = 0; v < b->n; ++v) {
        if (fills[v] < 0 || fills[v] > 255) return fail(PT_ERR_ARG, "fill[%d]=%d outside 0..255", v, fills[v]);
        b->h_fill[v] = fills[v];
        f[v] = b->pixel == PT_PIX_U8 ? (float)fills[v] : (float)fills[v] / 255.0f;
    }
    sync_batch_streams(b);               // no step on any of the batch's streams may still read the old fills
    rc = upload_small(b, b->d_fill, f.data(), sizeof(float) * (size_t)b->n); if (rc) return rc;
    rc = upload_small(b, b->d_fill_i, b->h_fill.data(), sizeof(int) * (size_t)b->n); if (rc) return rc;
    b->fill_set = true;
    return PT_OK;
}

int pt_batch_set_guess(pt_batch *b, const int32_t *guess_ij)
{
    if (!b || !guess_ij) return fail(PT_ERR_ARG, "NULL argument");
    // Kept in the host mirror; the device copy is refreshed (flush_guess) by the next call that steps on the
    // device.  The per-frame host path (footprint streaming of pageable frames) never needs it there.
    b->h_guess.assign(guess_ij, guess_ij + 2 * (size_t)b->n);
    b->h_guess_valid = true;
    b->d_guess_stale = true;
    b->guess_set = true;
    return PT_OK;
}

int pt_batch_step(pt_batch *b, const int32_t *guess_ij, int32_t *out_ij, int32_t *out_raw_ij, float *out_resp)
{
    if (!b) return fail(PT_ERR_ARG, "batch is NULL");
    int rc = set_device(b);
    if (rc) return rc;
    if (!b->fill_set) return fail(PT_ERR_STATE, "fill value not set: call pt_batch_compute_fill or pt_batch_set_fill");
    const void *base; size_t stride, pitch;
    rc = current_frames(b, &base, &stride, &pitch);
    if (rc) return rc;
    if (guess_ij) { rc = upload_guess(b, guess_ij, b->stream); if (rc) return rc; }
    else if (!b->guess_set) return fail(PT_ERR_STATE, "no guess on the device: pass guess_ij or call pt_batch_set_guess");
    else { rc = flush_guess(b); if (rc) return rc; }
    pt::WinArgs a = make_args(b, base, stride, pitch, b->H, b->W, b->d_guess, b->n);
    a.next_guess = b->d_guess;
    rc = launch_step(b, a, b->n, b->stream);
    if (rc) return rc;
    b->h_guess_valid = false;
    if (out_ij || out_raw_ij || out_resp) {
        rc = read_results(b, out_ij, out_raw_ij, out_resp, b->stream);
        if (rc == PT_OK) mirror_from_results(b, out_ij);
        return rc;
    }
    return PT_OK;
}

int pt_batch_track_device_async(pt_batch *b, const void *dev_base, size_t step_stride, size_t frame_stride,
                                size_t pitch, int T, void *stream)
{
    if (!b || !dev_base) return fail(PT_ERR_ARG, "NULL argument");
    if (T < 1) return fail(PT_ERR_ARG, "T must be >= 1");
    if (pitch < (size_t)b->W) return fail(PT_ERR_ARG, "pitch %zu smaller than W %d", pitch, b->W);
    int rc = set_device(b);
    if (rc) return rc;
    if (!b->fill_set) return fail(PT_ERR_STATE, "fill value not set");
    if (!b->guess_set) return fail(PT_ERR_STATE, "no guess on the device: call pt_batch_set_guess");
    rc = b->d_traj_pos.ensure(sizeof(int4) * (size_t)b->n * T, b); if (rc) return rc;
    rc = b->d_traj_resp.ensure(sizeof(float) * (size_t)b->n * T, b); if (rc) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : b->stream;
    if (s != b->stream) {
        // A batch runs on one stream at a time: work of earlier calls (own stream or another caller stream) is
        // finished before this one is enqueued, and the caller stream is remembered so that destroy / regrow /
        // pt_batch_read_track can wait for it.
        sync_batch_streams(b);
        b->ext_stream = s;
    } else if (b->ext_stream) {
        CU(cudaStreamSynchronize(b->ext_stream));
        b->ext_stream = nullptr;
    }
    rc = flush_guess(b, s); if (rc) return rc;
    b->h_guess_valid = false;                 // the chain advances on the device only
    const size_t es = px_size(b->pixel);
    {
        // the specialised kernel chains all T steps inside one launch (one CTA per video)
        pt::WinArgs a = make_args(b, dev_base, frame_stride, pitch, b->H, b->W, b->d_guess, b->n);
        if (b->cfg.window45 && pt::window45_supported(a, b->pixel)) {
            a.next_guess = b->d_guess;
            a.traj_pos = (int4 *)b->d_traj_pos.p; a.traj_resp = (float *)b->d_traj_resp.p;
            a.T = T; a.step_stride = step_stride;
            return launch_step(b, a, b->n, s);
        }
    }
    for (int t = 0; t < T; ++t) {
        const void *frames = (const char *)dev_base + (size_t)t * step_stride * es;
        pt::WinArgs a = make_args(b, frames, frame_stride, pitch, b->H, b->W, b->d_guess, b->n);
        a.next_guess = b->d_guess;
        a.traj_pos = (int4 *)b->d_traj_pos.p + (size_t)t * b->n;
        a.traj_resp = (float *)b->d_traj_resp.p + (size_t)t * b->n;
        rc = launch_step(b, a, b->n, s);
        if (rc) return rc;
    }
    return PT_OK;
}

int pt_batch_read_track(pt_batch *b, int T, int32_t *out_ij, float *out_resp)
{
    if (!b) return fail(PT_ERR_ARG, "batch is NULL");
    int rc = set_device(b);
    if (rc) return rc;
    const size_t cnt = (size_t)b->n * T;
    if (b->d_traj_pos.cap < cnt * sizeof(int4)) return fail(PT_ERR_STATE, "no trajectory of %d steps on the device", T);
    rc = b->h_out.ensure(std::max<size_t>(cnt * 20, (size_t)b->n * 32), b);
    if (rc) return rc;
    int4 *hp = (int4 *)b->h_out.p;
    float *hr = (float *)((char *)b->h_out.p + cnt * 16);
    // only this batch's streams are waited for (other handles keep running); pinned destination → true async copies
    cudaStream_t s = b->ext_stream ? b->ext_stream : b->stream;
    CU(cudaMemcpyAsync(hp, b->d_traj_pos.p, cnt * 16, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(hr, b->d_traj_resp.p, cnt * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    b->ext_stream = nullptr;
    b->staging_busy = false;
    for (size_t i = 0; i < cnt; ++i) {
        if (out_ij) { out_ij[2 * i] = hp[i].x; out_ij[2 * i + 1] = hp[i].y; }
        if (out_resp) out_resp[i] = hr[i];
    }
    return PT_OK;
}

int pt_batch_track_device(pt_batch *b, const void *dev_base, size_t step_stride, size_t frame_stride,
                          size_t pitch, int T, int32_t *out_ij, float *out_resp)
{
    int rc = pt_batch_track_device_async(b, dev_base, step_stride, frame_stride, pitch, T, nullptr);
    if (rc) return rc;
    return pt_batch_read_track(b, T, out_ij, out_resp);
}

} // extern "C"

// ---- host-resident frames ----------------------------------------------------
namespace {

// Gather the footprint of one window from a host frame into a crop of pitch
// `cp` elements: everything outside the frame is the fill value, exactly the
// PaddedView the reference filters (src/PawsomeTracker.jl:48).
template <typename T>
void gather_crop(const T *frame, size_t pitch, int H, int W, T fillv, int oy, int ox, int fr, int fc,
                 size_t cp, T *dst)
{
    const int xa = std::max(0, -ox), xb = std::min(fc, W - ox); // crop cols [xa, xb) are inside the frame
    for (int y = 0; y < fr; ++y) {
        T *d = dst + (size_t)y * cp;
        const int Y = oy + y;
        if (Y < 0 || Y >= H || xa >= xb) { std::fill(d, d + fc, fillv); continue; }
        if (xa > 0) std::fill(d, d + xa, fillv);
        memcpy(d + xa, frame + (size_t)Y * pitch + (ox + xa), sizeof(T) * (size_t)(xb - xa));
        if (xb < fc) std::fill(d + xb, d + fc, fillv);
    }
}

struct HostTrack {
    pt_batch *b;
    const void *const *frames;
    int T;
    size_t pitch;
    int32_t *out_ij;
    float *out_resp;
    std::vector<int32_t> guess; // n×2, host copy of the chain state
    std::atomic<int> err{PT_OK};
    std::string err_msg;
};

// One lane = one host thread + one stream: gather → H2D → kernel → D2H → sync,
// chained over T steps for its slice of videos.  Lanes never synchronise with
// one another (videos are independent units — SURVEY §8e).
void lane_worker(HostTrack *ht, pt_lane *ln)
{
    pt_batch *b = ht->b;
    if (cudaSetDevice(b->device) != cudaSuccess) { ht->err = PT_ERR_CUDA; return; }
    const int nl = ln->v1 - ln->v0;
    const int fr = b->wr + 2 * b->w, fc = b->wc + 2 * b->w;
    const size_t es = px_size(b->pixel);
    const size_t cp = b->pixel == PT_PIX_U8 ? (((size_t)fc + 15) & ~(size_t)15) : (((size_t)fc + 3) & ~(size_t)3);
    const size_t crop_elems = cp * (size_t)fr;
    int4 *hres = (int4 *)ln->h_res.p;
    float *hresp = (float *)((char *)ln->h_res.p + (size_t)nl * 16);
    std::vector<int> oy(nl), ox(nl);

    for (int t = 0; t < ht->T && ht->err.load() == PT_OK; ++t) {
        for (int i = 0; i < nl; ++i) {
            const int v = ln->v0 + i;
            const int gi = ht->guess[2 * v], gj = ht->guess[2 * v + 1];
            oy[i] = gi - 1 - b->rr - b->w; ox[i] = gj - 1 - b->rc - b->w;
            const void *f = ht->frames[(size_t)t * b->n + v];
            if (b->pixel == PT_PIX_U8)
                gather_crop<uint8_t>((const uint8_t *)f, ht->pitch, b->H, b->W, (uint8_t)b->h_fill[v], oy[i], ox[i],
                                     fr, fc, cp, (uint8_t *)ln->h_crops.p + crop_elems * i);
            else
                gather_crop<float>((const float *)f, ht->pitch, b->H, b->W, (float)b->h_fill[v] / 255.0f, oy[i], ox[i],
                                   fr, fc, cp, (float *)ln->h_crops.p + crop_elems * i);
        }
        cudaError_t e = cudaMemcpyAsync(ln->d_crops.p, ln->h_crops.p, crop_elems * es * nl, cudaMemcpyHostToDevice, ln->stream);
        if (e == cudaSuccess) {
            pt::WinArgs a = make_args(b, ln->d_crops.p, crop_elems, cp, fr, fc, b->d_center + ln->v0, nl);
            a.fill = b->d_fill + ln->v0;
            a.keys = b->d_keys + ln->v0; a.counters = b->d_counters + ln->v0; a.tickets = b->d_counters + b->n + 2 * ln->v0;
            a.out_pos = hres; a.out_resp = hresp;      // pinned + mapped: the kernel stores the results straight into host memory
            a.mid = ln->d_mid;                         // (two-phase wide path: this lane's slice of the intermediate, or null)
            e = launch_windows(b, a, nl, ln->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ln->stream);
        if (e != cudaSuccess) {
            int expected = PT_OK;
            if (ht->err.compare_exchange_strong(expected, PT_ERR_CUDA)) ht->err_msg = cudaGetErrorString(e);
            return;
        }
        for (int i = 0; i < nl; ++i) {
            const int v = ln->v0 + i;
            // crop row/col 1-based → frame 1-based, then clamp (src/PawsomeTracker.jl:60-61)
            const int raw_i = oy[i] + hres[i].z, raw_j = ox[i] + hres[i].w;
            const int ci = std::min(std::max(raw_i, 1), b->H), cj = std::min(std::max(raw_j, 1), b->W);
            ht->guess[2 * v] = ci; ht->guess[2 * v + 1] = cj;
            int32_t *o = ht->out_ij + ((size_t)t * b->n + v) * 2;
            o[0] = ci; o[1] = cj;
            if (ht->out_resp) ht->out_resp[(size_t)t * b->n + v] = hresp[i];
        }
    }
}

int ensure_lanes(pt_batch *b)
{
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    int want = (int)std::min<unsigned>(std::min<unsigned>(hw, 16u), (unsigned)b->n);
    if (b->cfg.host_lanes >= 1) want = std::min(b->cfg.host_lanes, b->n);
    if ((int)b->lanes.size() != want) {
        for (auto &ln : b->lanes) { if (ln.stream) cudaStreamDestroy(ln.stream); ln.h_crops.release(); ln.h_res.release(); ln.d_crops.release(); }
        b->lanes.assign(want, pt_lane());
        for (int i = 0; i < want; ++i) {
            b->lanes[i].v0 = (int)((long long)b->n * i / want);
            b->lanes[i].v1 = (int)((long long)b->n * (i + 1) / want);
            b->lanes[i].index = i;
            CU(cudaStreamCreateWithFlags(&b->lanes[i].stream, cudaStreamNonBlocking));
        }
    }
    const int fr = b->wr + 2 * b->w, fc = b->wc + 2 * b->w;
    const size_t es = px_size(b->pixel);
    const size_t cp = b->pixel == PT_PIX_U8 ? (((size_t)fc + 15) & ~(size_t)15) : (((size_t)fc + 3) & ~(size_t)3);
    for (auto &ln : b->lanes) {
        const size_t nl = (size_t)(ln.v1 - ln.v0);
        int rc = ln.h_crops.ensure(cp * fr * es * nl, b); if (rc) return rc;
        rc = ln.d_crops.ensure(cp * fr * es * nl, b); if (rc) return rc;
        rc = ln.h_res.ensure(nl * 20, b); if (rc) return rc;
    }
    {
        // two-phase wide path: one intermediate for the whole batch, a slice per lane (crops are fr x fc "frames")
        pt::WinArgs probe = make_args(b, nullptr, 0, cp, fr, fc, b->d_center, b->n);
        int rc = ensure_mid(b, probe, b->n);
        if (rc) return rc;
        const size_t per = pt::wide_mid_elems(b->L, b->wr, b->wc, 1);
        for (auto &ln : b->lanes) ln.d_mid = probe.mid ? probe.mid + per * (size_t)ln.v0 : nullptr;
    }
    if (b->center_key[0] != b->rr || b->center_key[1] != b->rc || b->center_key[2] != b->w) {
        std::vector<int2> c(b->n, make_int2(b->rr + b->w + 1, b->rc + b->w + 1));
        int rc = upload_small(b, b->d_center, c.data(), sizeof(int2) * (size_t)b->n);   // complete before any lane stream starts
        if (rc) return rc;
        b->center_key[0] = b->rr; b->center_key[1] = b->rc; b->center_key[2] = b->w;
    }
    return PT_OK;
}

} // namespace

extern "C" {

int pt_batch_track_host(pt_batch *b, const void *const *frames, int T, size_t pitch, int mode,
                        int32_t *out_ij, float *out_resp)
{
    if (!b || !frames || !out_ij) return fail(PT_ERR_ARG, "NULL argument");
    if (T < 1) return fail(PT_ERR_ARG, "T must be >= 1");
    if (pitch < (size_t)b->W) return fail(PT_ERR_ARG, "pitch %zu smaller than W %d", pitch, b->W);
    if (mode != 0 && mode != 1) return fail(PT_ERR_ARG, "mode must be 0 (footprint) or 1 (whole frames)");
    int rc = set_device(b);
    if (rc) return rc;
    if (!b->fill_set) return fail(PT_ERR_STATE, "fill value not set");
    if (!b->guess_set) return fail(PT_ERR_STATE, "no guess on the device: call pt_batch_set_guess");
    const size_t n = (size_t)b->n;

    if (mode == 1) {
        // whole frames: upload step t+1 on the copy stream into the other slot
        // while step t's kernel runs; one D2H of the result per step.
        b->bound_base = nullptr;
        rc = flush_guess(b); if (rc) return rc;
        CU(cudaStreamSynchronize(b->stream));
        rc = upload_frames(b, frames, pitch, 0, b->copy_stream); if (rc) return rc;
        CU(cudaEventRecord(b->ev_done[0], b->copy_stream));
        for (int t = 0; t < T; ++t) {
            const int slot = t & 1;
            if (t + 1 < T) {
                // slot^1 was last read by step t-1, which we synchronised on below
                rc = upload_frames(b, frames + (size_t)(t + 1) * n, pitch, slot ^ 1, b->copy_stream); if (rc) return rc;
                CU(cudaEventRecord(b->ev_done[slot ^ 1], b->copy_stream));
            }
            CU(cudaStreamWaitEvent(b->stream, b->ev_done[slot], 0));
            pt::WinArgs a = make_args(b, b->d_frames[slot].p, b->own_stride, b->own_pitch, b->H, b->W, b->d_guess, b->n);
            a.next_guess = b->d_guess;
            rc = launch_step(b, a, b->n, b->stream); if (rc) return rc;
            rc = read_results(b, out_ij + (size_t)t * n * 2, nullptr, out_resp ? out_resp + (size_t)t * n : nullptr, b->stream);
            if (rc) return rc;
        }
        b->cur_slot = (T - 1) & 1;
        b->have_frames = true;
        mirror_from_results(b, out_ij + (size_t)(T - 1) * n * 2);
        return PT_OK;
    }

    // footprint streaming, zero-copy variant: when every frame is page-locked (pinned) host memory and
    // the geometry dispatches to the chained kernel, the kernel reads each window's footprint straight
    // from the host frame over PCIe and writes every step's result straight into pinned host memory —
    // the whole frame loop (:163-169) is one launch, with no host gather and no per-step round trip.
    if (b->cfg.zero_copy) {
        pt::WinArgs probe = make_args(b, nullptr, 0, pitch, b->H, b->W, b->d_guess, b->n);
        probe.frame_ptrs = reinterpret_cast<const void *const *>(1);   // "pointer table" marker for the support check
        const bool w45 = b->cfg.window45 && pt::window45_supported(probe, b->pixel);
        bool ok = true;
        const size_t frame_bytes = ((size_t)(b->H - 1) * pitch + (size_t)b->W) * px_size(b->pixel);   // first to last byte of a frame
        const size_t cnt = n * (size_t)T;
        rc = b->h_ptrs.ensure(cnt * sizeof(void *), b); if (rc) return rc;
        const void **hp = (const void **)b->h_ptrs.p;
        // host address range [rb, rb+rs) already known to be pinned, and its device alias rd
        const char *rb = nullptr, *rd = nullptr; size_t rs = 0;
        for (size_t i = 0; ok && i < cnt; ++i) {
            const char *f = (const char *)frames[i];
            if (!f) return fail(PT_ERR_ARG, "frames[%zu] is NULL", i);
            if (!(rb && f >= rb && f + frame_bytes <= rb + rs)) {
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, f) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
                if (at.type != cudaMemoryTypeHost || !at.devicePointer) { ok = false; break; }
                const char *base = nullptr; size_t size = 0;
                if (alloc_range(f, &base, &size) && f >= base && f + frame_bytes <= base + size) {
                    rb = base; rs = size; rd = (const char *)at.devicePointer - (f - base);
                } else {
                    // allocation range unknown (or the frame would run past it): check the frame's last byte too
                    cudaPointerAttributes at2;
                    if (cudaPointerGetAttributes(&at2, f + frame_bytes - 1) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
                    if (at2.type != cudaMemoryTypeHost || !at2.devicePointer) { ok = false; break; }
                    rb = f; rs = frame_bytes; rd = (const char *)at.devicePointer;      // this frame only
                }
            }
            hp[i] = rd + (f - rb);
            if (w45 && b->pixel == PT_PIX_U8 && ((uintptr_t)hp[i] & 3u) != 0) { ok = false; break; }
        }
        if (ok && w45) {
            rc = b->d_ptrs.ensure(cnt * sizeof(void *), b); if (rc) return rc;
            rc = b->h_traj.ensure(cnt * 20, b); if (rc) return rc;
            int4 *hpos = (int4 *)b->h_traj.p;
            float *hresp = (float *)((char *)b->h_traj.p + cnt * 16);
            rc = flush_guess(b); if (rc) return rc;
            // Frames at regular strides (one page-locked buffer holding [T][n] frames — a decoder ring, PinnedArray,
            // FrameFeeder chunks) are handed to the kernels as base + strides, exactly like resident frames: the kernels
            // then prefetch each step's region one step ahead (L2 prefetch / the cluster kernel's TMA tile copy work on
            // mapped host memory too), which takes the PCIe round trip off the serial chain — one 1080p video: 4.2
            // instead of 6.7 µs per frame, 16 videos: 9 instead of 27 (tools/pinned_chain_timing.py).
            const size_t es = px_size(b->pixel);
            const char *p0 = (const char *)hp[0];
            const ptrdiff_t sv = n > 1 ? (const char *)hp[1] - p0 : 0, st = T > 1 ? (const char *)hp[n] - p0 : 0;
            bool regular = sv >= 0 && st >= 0 && sv % (ptrdiff_t)es == 0 && st % (ptrdiff_t)es == 0 &&
                           (b->pixel != PT_PIX_U8 || ((sv | st) & 3) == 0);
            for (size_t t = 0; regular && t < (size_t)T; ++t)
                for (size_t v = 0; v < n; ++v)
                    if ((const char *)hp[t * n + v] != p0 + (ptrdiff_t)t * st + (ptrdiff_t)v * sv) { regular = false; break; }
            pt::WinArgs a = make_args(b, regular ? (const void *)p0 : nullptr, regular ? (size_t)sv / es : 0, pitch, b->H, b->W,
                                      b->d_guess, b->n);
            if (regular) {
                a.step_stride = (size_t)st / es;
                a.host_frames = 1;
            } else {
                CU(cudaMemcpyAsync(b->d_ptrs.p, hp, cnt * sizeof(void *), cudaMemcpyHostToDevice, b->stream));
                a.frame_ptrs = (const void *const *)b->d_ptrs.p;
            }
            a.next_guess = b->d_guess;
            a.traj_pos = hpos; a.traj_resp = hresp;           // pinned + mapped: zero-copy stores, one per step
            a.T = T;
            rc = launch_step(b, a, b->n, b->stream); if (rc) return rc;
            CU(cudaStreamSynchronize(b->stream));
            for (size_t i = 0; i < cnt; ++i) { out_ij[2 * i] = hpos[i].x; out_ij[2 * i + 1] = hpos[i].y; if (out_resp) out_resp[i] = hresp[i]; }
            mirror_from_results(b, out_ij + (size_t)(T - 1) * n * 2);
            return PT_OK;
        }
        // Any other geometry (long kernels, large windows: BASELINE config 4): the streaming kernels read the page-locked
        // frames in place too — one launch (two for the two-phase wide path) per step, all T steps enqueued at once, the
        // chain advancing on the device, results stored straight into pinned memory; no crop gather, no per-step copy, no
        // per-step round trip.  Needs the n frames of a step at one regular stride.
        if (ok && !w45) {
            const size_t es = px_size(b->pixel);
            const ptrdiff_t sv = n > 1 ? (const char *)hp[1] - (const char *)hp[0] : 0;
            bool regular = sv >= 0 && sv % (ptrdiff_t)es == 0;
            for (size_t t = 0; regular && t < (size_t)T; ++t)
                for (size_t v = 0; v < n; ++v)
                    if ((const char *)hp[t * n + v] != (const char *)hp[t * n] + (ptrdiff_t)v * sv) { regular = false; break; }
            if (regular) {
                rc = b->h_traj.ensure(cnt * 20, b); if (rc) return rc;
                int4 *hpos = (int4 *)b->h_traj.p;
                float *hresp = (float *)((char *)b->h_traj.p + cnt * 16);
                rc = flush_guess(b); if (rc) return rc;
                // Each step first copies every window's footprint out of its host frame into a device crop, once and in
                // 16-byte pieces (gather_footprints: the window position is on the device, so the copy is a kernel); the
                // filter kernels then run on the crops out of HBM — reading the host frames directly they would pull the
                // kernel-length halos of their column strips over PCIe again and again (one 401x401 window at l = 245:
                // 1.4 MB instead of 0.4 MB per step; 73 → ≈ 40 µs per frame through track()).
                const int epc = (int)(16 / es);
                const int fr = b->wr + 2 * b->w, fc = b->wc + 2 * b->w;
                const int cpitch = ((fc + epc - 1 + epc - 1) / epc) * epc;          // footprint + alignment phase, whole pieces
                const size_t cstride = (size_t)fr * cpitch;
                const bool crops = b->cfg.crop_gather != 0 && cstride * es * n <= ((size_t)4 << 30);
                int2 *d_org = nullptr, *d_cg = nullptr;
                if (crops) {
                    rc = b->d_crop.ensure(cstride * es * n, b); if (rc) return rc;
                    rc = b->d_cropmeta.ensure(sizeof(int2) * 2 * n, b); if (rc) return rc;
                    d_org = (int2 *)b->d_cropmeta.p; d_cg = d_org + n;
                }
                for (int t = 0; t < T; ++t) {
                    pt::WinArgs a;
                    if (crops) {
                        cudaError_t ge = pt::launch_gather_footprints(hp[(size_t)t * n], (size_t)sv / es, (int)pitch, b->H, b->W, b->n, b->pixel,
                                                                      b->d_guess, b->d_fill, b->rr, b->rc, b->w, fr, cpitch,
                                                                      b->d_crop.p, cstride, d_org, d_cg, b->stream);
                        if (ge != cudaSuccess) return fail(PT_ERR_CUDA, "gather_footprints: %s", cudaGetErrorString(ge));
                        b->launches += 1;
                        a = make_args(b, b->d_crop.p, cstride, (size_t)cpitch, fr, cpitch, d_cg, b->n);
                        a.crop_org = d_org; a.Hreal = b->H; a.Wreal = b->W;
                    } else {
                        a = make_args(b, hp[(size_t)t * n], (size_t)sv / es, pitch, b->H, b->W, b->d_guess, b->n);
                        a.host_frames = 1;
                    }
                    a.next_guess = b->d_guess;
                    a.traj_pos = hpos + (size_t)t * n; a.traj_resp = hresp + (size_t)t * n;
                    rc = launch_step(b, a, b->n, b->stream); if (rc) return rc;
                }
                CU(cudaStreamSynchronize(b->stream));
                for (size_t i = 0; i < cnt; ++i) { out_ij[2 * i] = hpos[i].x; out_ij[2 * i + 1] = hpos[i].y; if (out_resp) out_resp[i] = hresp[i]; }
                mirror_from_results(b, out_ij + (size_t)(T - 1) * n * 2);
                return PT_OK;
            }
        }
    }

    // footprint streaming through pinned staging (pageable host frames, or pinned frames without a regular layout)
    rc = ensure_lanes(b); if (rc) return rc;
    HostTrack ht;
    ht.b = b; ht.frames = frames; ht.T = T; ht.pitch = pitch; ht.out_ij = out_ij; ht.out_resp = out_resp;
    if (b->ext_stream) { CU(cudaStreamSynchronize(b->ext_stream)); b->ext_stream = nullptr; }
    CU(cudaStreamSynchronize(b->stream));          // fill values / centres / earlier steps are complete
    b->staging_busy = false;
    if (b->h_guess_valid) ht.guess = b->h_guess;
    else {
        ht.guess.resize(n * 2);
        CU(cudaMemcpy(ht.guess.data(), b->d_guess, sizeof(int2) * n, cudaMemcpyDeviceToHost));
    }
    std::vector<std::thread> th;
    for (size_t i = 1; i < b->lanes.size(); ++i) th.emplace_back(lane_worker, &ht, &b->lanes[i]);
    lane_worker(&ht, &b->lanes[0]);
    for (auto &t : th) t.join();
    b->launches += (long long)T * (long long)b->lanes.size();
    if (ht.err.load() != PT_OK) return fail(ht.err.load(), "footprint streaming failed: %s", ht.err_msg.c_str());
    // the chain state now lives in the host mirror; the device copy is refreshed lazily (flush_guess)
    // by the next call that steps on the device, as pt_batch_step would expect
    b->h_guess = ht.guess;
    b->h_guess_valid = true;
    b->d_guess_stale = true;
    return PT_OK;
}

int pt_batch_response_map(pt_batch *b, int v, int gi, int gj, float *out_map)
{
    if (!b || !out_map) return fail(PT_ERR_ARG, "NULL argument");
    if (v < 0 || v >= b->n) return fail(PT_ERR_ARG, "video index %d out of range", v);
    int rc = set_device(b);
    if (rc) return rc;
    if (!b->fill_set) return fail(PT_ERR_STATE, "fill value not set");
    const void *base; size_t stride, pitch;
    rc = current_frames(b, &base, &stride, &pitch); if (rc) return rc;
    const size_t cnt = (size_t)b->wr * b->wc;
    rc = b->d_map.ensure(cnt * 4, b); if (rc) return rc;
    const size_t es = px_size(b->pixel);
    pt::WinArgs a = make_args(b, (const char *)base + (size_t)v * stride * es, stride, pitch, b->H, b->W, nullptr, 1);
    a.rect_mode = 1; a.ry0 = gi - 1 - b->rr; a.rx0 = gj - 1 - b->rc;
    a.fill = b->d_fill + v; a.keys = b->d_keys + v; a.counters = b->d_counters + v; a.tickets = b->d_counters + b->n + 2 * v;
    a.out_pos = b->d_pos + v; a.out_resp = b->d_resp + v;
    a.map_out = (float *)b->d_map.p;
    rc = launch_step(b, a, 1, b->stream); if (rc) return rc;
    CU(cudaMemcpyAsync(out_map, b->d_map.p, cnt * 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return PT_OK;
}

int pt_batch_rect_argmax(pt_batch *b, int v, int y0, int x0, int wr, int wc,
                         int *oi, int *oj, int *raw_i, int *raw_j, float *resp)
{
    if (!b) return fail(PT_ERR_ARG, "batch is NULL");
    if (v < 0 || v >= b->n) return fail(PT_ERR_ARG, "video index %d out of range", v);
    if (wr < 1 || wc < 1) return fail(PT_ERR_ARG, "rectangle must be non-empty");
    if ((long long)wr * wc > 0x7FFFFFFFll) return fail(PT_ERR_ARG, "rectangle too large");
    int rc = set_device(b);
    if (rc) return rc;
    if (!b->fill_set) return fail(PT_ERR_STATE, "fill value not set");
    const void *base; size_t stride, pitch;
    rc = current_frames(b, &base, &stride, &pitch); if (rc) return rc;
    const size_t es = px_size(b->pixel);
    pt::WinArgs a = make_args(b, (const char *)base + (size_t)v * stride * es, stride, pitch, b->H, b->W, nullptr, 1);
    a.rect_mode = 1; a.ry0 = y0; a.rx0 = x0; a.wr = wr; a.wc = wc;
    a.fill = b->d_fill + v; a.keys = b->d_keys + v; a.counters = b->d_counters + v; a.tickets = b->d_counters + b->n + 2 * v;
    a.out_pos = b->d_pos + v; a.out_resp = b->d_resp + v;
    rc = launch_step(b, a, 1, b->stream); if (rc) return rc;
    rc = b->h_out.ensure((size_t)b->n * 32, b); if (rc) return rc;
    int4 *hp = reinterpret_cast<int4 *>(b->h_out.p);                 // pinned: truly asynchronous copies
    float *hr = reinterpret_cast<float *>((char *)b->h_out.p + 16);
    CU(cudaMemcpyAsync(hp, b->d_pos + v, sizeof(int4), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(hr, b->d_resp + v, sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    b->staging_busy = false;
    const int4 p = *hp; const float r = *hr;
    if (oi) *oi = p.x; if (oj) *oj = p.y; if (raw_i) *raw_i = p.z; if (raw_j) *raw_j = p.w; if (resp) *resp = r;
    return PT_OK;
}

int pt_batch_rect_argmax_all(pt_batch *b, int y0, int x0, int wr, int wc, int32_t *out_ij, float *out_resp, int no_readback)
{
    if (!b) return fail(PT_ERR_ARG, "batch is NULL");
    if (wr < 1 || wc < 1) return fail(PT_ERR_ARG, "rectangle must be non-empty");
    if ((long long)wr * wc > 0x7FFFFFFFll) return fail(PT_ERR_ARG, "rectangle too large");
    int rc = set_device(b);
    if (rc) return rc;
    if (!b->fill_set) return fail(PT_ERR_STATE, "fill value not set");
    const void *base; size_t stride, pitch;
    rc = current_frames(b, &base, &stride, &pitch); if (rc) return rc;
    pt::WinArgs a = make_args(b, base, stride, pitch, b->H, b->W, nullptr, b->n);
    a.rect_mode = 1; a.ry0 = y0; a.rx0 = x0; a.wr = wr; a.wc = wc;
    rc = launch_step(b, a, b->n, b->stream); if (rc) return rc;
    if (no_readback) return PT_OK;
    rc = b->h_out.ensure((size_t)b->n * 32, b); if (rc) return rc;
    int4 *hp = reinterpret_cast<int4 *>(b->h_out.p);
    float *hr = reinterpret_cast<float *>((char *)b->h_out.p + (size_t)b->n * 16);
    CU(cudaMemcpyAsync(hp, b->d_pos, (size_t)b->n * sizeof(int4), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(hr, b->d_resp, (size_t)b->n * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    b->staging_busy = false;
    for (int v = 0; v < b->n; ++v) {
        if (out_ij) { out_ij[4 * v] = hp[v].x; out_ij[4 * v + 1] = hp[v].y; out_ij[4 * v + 2] = hp[v].z; out_ij[4 * v + 3] = hp[v].w; }
        if (out_resp) out_resp[v] = hr[v];
    }
    return PT_OK;
}

int pt_batch_set_option(pt_batch *b, const char *name, int value)
{
    if (!b || !name) return fail(PT_ERR_ARG, "NULL argument");
    struct Opt { const char *name; int pt::Cfg::* field; int lo, hi; };
    static const Opt opts[] = {
        {"window45", &pt::Cfg::window45, 0, 1},       {"rect45", &pt::Cfg::rect45, 0, 1},
        {"rot", &pt::Cfg::rot, 0, 3},                 {"rot_stride", &pt::Cfg::rot_stride, 0, 1 << 20},                 {"skew", &pt::Cfg::skew, 0, 2},
        {"r45_chunks", &pt::Cfg::r45_chunks, 0, 1 << 20}, {"generic_target", &pt::Cfg::generic_target, 1, 1 << 20},
        {"mode_slow", &pt::Cfg::mode_slow, 0, 1},     {"zero_copy", &pt::Cfg::zero_copy, 0, 1},
        {"host_lanes", &pt::Cfg::host_lanes, 0, 1024}, {"cluster", &pt::Cfg::cluster, 0, 8},
        {"bulk", &pt::Cfg::bulk, 0, 1},
        {"wide", &pt::Cfg::wide, 0, 1},               {"two_phase", &pt::Cfg::two_phase, 0, 2},         {"cols_teams", &pt::Cfg::cols_teams, 0, 1},       {"crop_gather", &pt::Cfg::crop_gather, 0, 1},       {"cols_ch", &pt::Cfg::cols_ch, 0, 1 << 16},
    };
    for (const Opt &o : opts) {
        if (strcmp(o.name, name) != 0) continue;
        if (value < o.lo || value > o.hi) return fail(PT_ERR_ARG, "option %s: value %d outside %d..%d", name, value, o.lo, o.hi);
        if (o.field == &pt::Cfg::cluster && !(value == 0 || value == 1 || value == 2 || value == 4 || value == 8))
            return fail(PT_ERR_ARG, "option cluster: 0 (auto), 1 (off), 2, 4 or 8 CTAs per window");
        b->cfg.*(o.field) = value;
        return PT_OK;
    }
    return fail(PT_ERR_ARG, "unknown option '%s'", name);
}

long long pt_batch_launch_count(const pt_batch *b) { return b ? b->launches : 0; }

const char *pt_batch_kernel_name(const pt_batch *b)
{
    if (!b) return "";
    pt::WinArgs a;
    memset(&a, 0, sizeof a);
    a.wr = b->wr; a.wc = b->wc; a.L = b->L; a.w = b->w; a.h_taps = b->h_taps.data();
    if (b->cfg.window45 && pt::window45_supported(a, b->pixel)) return pt::window45_name();
    if (b->cfg.window45 && b->cfg.rect45 && a.L == 65 && (long long)a.wr * a.wc >= 24 * 24) return pt::rect45_name();
    return b->pixel == PT_PIX_U8 ? "dog_rect_argmax_generic<u8>" : "dog_rect_argmax_generic<f32>";
}

int pt_batch_downscale(pt_batch *b, int out_h, int out_w, uint8_t *out)
{
    if (!b || !out) return fail(PT_ERR_ARG, "NULL argument");
    if (out_h < 1 || out_w < 1) return fail(PT_ERR_ARG, "output size must be positive");
    int rc = set_device(b);
    if (rc) return rc;
    const void *base; size_t stride, pitch;
    rc = current_frames(b, &base, &stride, &pitch); if (rc) return rc;
    const size_t bytes = (size_t)b->n * out_h * out_w;
    rc = b->d_map.ensure(bytes, b); if (rc) return rc;
    cudaError_t e = pt::launch_downscale(base, stride, (int)pitch, b->H, b->W, b->n, b->pixel, out_h, out_w,
                                         (uint8_t *)b->d_map.p, b->stream);
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    CU(cudaMemcpyAsync(out, b->d_map.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return PT_OK;
}

int pt_host_alloc(size_t bytes, void **out)
{
    if (!out || bytes == 0) return fail(PT_ERR_ARG, "bad argument");
    *out = nullptr;
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return PT_OK;
}

int pt_host_free(void *p)
{
    if (p) CU(cudaFreeHost(p));
    return PT_OK;
}

const char *pt_batch_last_kernel(const pt_batch *b) { return b ? b->last_kernel : ""; }

#ifdef PT_PROBES
// Profiling build only (libpawsome_cuda_probes.so): phase timestamps of the window45 kernels → dev_buf [n][T][6] int64.
PT_API int pt_debug_window45_timing(void *dev_buf)
{
    pt::window45_set_debug((long long *)dev_buf);
    return PT_OK;
}
#endif

void *pt_batch_stream(const pt_batch *b) { return b ? (void *)b->stream : nullptr; }

// ---- single tracker ----------------------------------------------------------
int pt_tracker_create(int H, int W, double tw, int ws_rows, int ws_cols, int darker, int pixel, int device,
                      pt_tracker **out)
{
    if (!out) return fail(PT_ERR_ARG, "out is NULL");
    *out = nullptr;
    pt_batch *b = nullptr;
    int rc = pt_batch_create(1, H, W, tw, ws_rows, ws_cols, darker, pixel, device, &b);
    if (rc) return rc;
    pt_tracker *t = new (std::nothrow) pt_tracker();
    if (!t) { pt_batch_destroy(b); return fail(PT_ERR_NOMEM, "out of host memory"); }
    t->b = b;
    *out = t;
    return PT_OK;
}

void pt_tracker_destroy(pt_tracker *t)
{
    if (!t) return;
    pt_batch_destroy(t->b);
    delete t;
}

pt_batch *pt_tracker_batch(pt_tracker *t) { return t ? t->b : nullptr; }

int pt_tracker_set_frame(pt_tracker *t, const void *frame, size_t pitch)
{
    if (!t) return fail(PT_ERR_ARG, "tracker is NULL");
    const void *f[1] = {frame};
    return pt_batch_set_frames(t->b, f, pitch);
}

int pt_tracker_compute_fill(pt_tracker *t, int *fill_out)
{
    if (!t) return fail(PT_ERR_ARG, "tracker is NULL");
    return pt_batch_compute_fill(t->b, fill_out);
}

int pt_tracker_set_fill(pt_tracker *t, int fill)
{
    if (!t) return fail(PT_ERR_ARG, "tracker is NULL");
    return pt_batch_set_fill(t->b, &fill);
}

int pt_tracker_step(pt_tracker *t, int gi, int gj, int *oi, int *oj, float *resp)
{
    if (!t) return fail(PT_ERR_ARG, "tracker is NULL");
    int32_t g[2] = {gi, gj}, o[2] = {0, 0};
    float r = 0.f;
    int rc = pt_batch_step(t->b, g, o, nullptr, &r);
    if (rc) return rc;
    if (oi) *oi = o[0]; if (oj) *oj = o[1]; if (resp) *resp = r;
    return PT_OK;
}

int pt_tracker_step_host(pt_tracker *t, const void *frame, size_t pitch, int gi, int gj, int *oi, int *oj, float *resp)
{
    if (!t || !frame) return fail(PT_ERR_ARG, "NULL argument");
    int32_t g[2] = {gi, gj}, o[2] = {0, 0};
    float r = 0.f;
    int rc = pt_batch_set_guess(t->b, g);
    if (rc) return rc;
    const void *f[1] = {frame};
    rc = pt_batch_track_host(t->b, f, 1, pitch, 0, o, &r);
    if (rc) return rc;
    if (oi) *oi = o[0]; if (oj) *oj = o[1]; if (resp) *resp = r;
    return PT_OK;
}

} // extern "C"
