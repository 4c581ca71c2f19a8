// Context management, the bounding-sphere pyramid over the model grid, and K0 (grid angle).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <new>
#include <thread>
#include <unordered_map>
#include <vector>

#include "gzb_common.cuh"

// ======================================================================== error plumbing
static thread_local char g_global_err[512] = "";

void gzb_set_global_error(const char* msg) {
    strncpy(g_global_err, msg, sizeof(g_global_err) - 1);
    g_global_err[sizeof(g_global_err) - 1] = 0;
}

int gzb_fail(gzb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) {
        strncpy(ctx->err, buf, sizeof(ctx->err) - 1);
        ctx->err[sizeof(ctx->err) - 1] = 0;
    }
    gzb_set_global_error(buf);
    return code;
}

#ifdef GZB_CHECKED
extern "C" const char* gzb_version(void) { return "gonzag_b200 0.2 (sm_100a, CHECKED build)"; }
#else
extern "C" const char* gzb_version(void) { return "gonzag_b200 0.2 (sm_100a)"; }
#endif
extern "C" const char* gzb_last_global_error(void) { return g_global_err; }
extern "C" const char* gzb_last_error(const gzb_ctx* ctx) { return ctx ? ctx->err : g_global_err; }

extern "C" int gzb_device_count(int* n_out) {
    if (!n_out) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_device_count: null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *n_out = 0;
        return gzb_fail(nullptr, GZB_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *n_out = n;
    return GZB_OK;
}

void gzb_trace(gzb_ctx* ctx, cudaStream_t s, const char* what) {
    if (!ctx->trace || ctx->ntrace >= 48) return;
    const int k = ctx->ntrace++;
    if (!ctx->trace_ev[k]) cudaEventCreate(&ctx->trace_ev[k]);
    ctx->trace_name[k] = what;
    cudaEventRecord(ctx->trace_ev[k], s);
}

void gzb_apply_carveout(gzb_ctx* ctx) {
    static int applied[64];
    static bool init = false;
    if (!init) { for (int k = 0; k < 64; ++k) applied[k] = -2; init = true; }
    const int d = ctx->device & 63;
    if (applied[d] == ctx->carveout) return;
    applied[d] = ctx->carveout;
    gzb_carveout_nearest(ctx->carveout);
    gzb_carveout_mesh(ctx->carveout);
    cudaGetLastError();
}

int gzb_use_device(gzb_ctx* ctx) {
    GZB_CUDA(ctx, cudaSetDevice(ctx->device));
    return GZB_OK;
}


// ======================================================================== big pageable copies
// cudaMemcpy from / to pageable memory is bound by ONE host thread moving the data through the
// driver's staging buffer (measured here: 11 GB/s up, 5 GB/s down into freshly allocated pages).
// Large transfers are therefore cut into GZB_BC_THREADS contiguous parts, each moved by its own
// thread through two pinned 8 MiB slots on its own stream (host memcpy of one slot overlapping the
// DMA of the other).  The slots are allocated once per process and kept (gzb_pool_trim frees them).
#define GZB_BC_THREADS 24
// how many host threads the library's host-side helpers (parallel copies, mask packing, record staging) may start:
// 0 = automatic (the machine's hardware threads), else the cap set by gzb_set_host_threads() or GZB_HOST_THREADS --
// with one process per GPU every rank gets its share of the cores instead of all of them (8 x 12 threads on 32 cores)
static int g_host_threads = -1;
int gzb_host_threads(int fallback_cap) {
    if (g_host_threads < 0) {
        const char* e = getenv("GZB_HOST_THREADS");
        g_host_threads = e ? atoi(e) : 0;
        if (g_host_threads < 0) g_host_threads = 0;
    }
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    int want = g_host_threads > 0 ? g_host_threads : hw;
    if (want > hw) want = hw;
    if (want > fallback_cap) want = fallback_cap;
    return want < 1 ? 1 : want;
}
extern "C" int gzb_set_host_threads(int n) {
    g_host_threads = n < 0 ? 0 : n;
    return GZB_OK;
}
#define GZB_BC_SLOT ((size_t)1 << 20)      /* small slots: a thread fills one before its first DMA can start */
#define GZB_BC_MIN ((size_t)24 << 20)
namespace {
struct BigCopy {
    void* slot[GZB_BC_THREADS][2];
    cudaStream_t stream[GZB_BC_THREADS];
    cudaEvent_t ev[GZB_BC_THREADS][2];
    int dev = -1;
    bool ok = false;
};
BigCopy g_bc;
std::mutex g_bc_mu;

bool bigcopy_ready(int dev) {
    if (g_bc.ok && g_bc.dev == dev) return true;
    if (g_bc.ok) return false;               // slots belong to another device: fall back to plain copies
    for (int t = 0; t < GZB_BC_THREADS; ++t) {
        for (int k = 0; k < 2; ++k) {
            if (cudaHostAlloc(&g_bc.slot[t][k], GZB_BC_SLOT, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return false; }
            if (cudaEventCreateWithFlags(&g_bc.ev[t][k], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
        }
        if (cudaStreamCreateWithFlags(&g_bc.stream[t], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
    }
    g_bc.dev = dev;
    g_bc.ok = true;
    return true;
}

// pack > 0 (uploads only): `src` holds integers of `pack` bytes each and dst[k] = (int8_t)src[k] -- the land-sea mask of the
// reference is int64 {0,1} (gonzag/ncio.py:164-191), the context keeps int8: the conversion rides on the staging copy
template <typename T> static void pack_into(int8_t* out, const char* src, size_t first, size_t n) {
    const T* p = (const T*)src + first;
    for (size_t k = 0; k < n; ++k) out[k] = (int8_t)p[k];
}

// Narrowing upload (pack == GZB_BC_NARROW, uploads only): `src` holds float64 values, `bytes` counts SOURCE bytes, and what
// travels is float32 -- as long as every value survives the round trip double -> float -> double unchanged (model grids are
// commonly written by NEMO in single precision and widened by the reader: half the bytes over PCIe, bit-identical values
// after the widening kernel).  The first value that does not (or a NaN) raises the job's `inexact` flag (the caller's:
// contexts may be created from several host threads) and the caller repeats the transfer in full precision.
#define GZB_BC_NARROW 108

static bool narrow_into(float* out, const double* p, size_t n) {
    bool ok = true;
    for (size_t k = 0; k < n; ++k) {
        const float f = (float)p[k];
        ok &= ((double)f == p[k]);
        out[k] = f;
    }
    return ok;
}

// one transfer of a threaded copy: `bytes` from src to dst; pack > 1: see pack_into (then `bytes` counts int8 elements)
struct BcJob {
    char* dst;
    const char* src;
    size_t bytes;
    int pack;
    size_t min_bytes;      // below this the transfer is one plain copy (0 = GZB_BC_MIN)
    std::atomic<int>* inexact;   // GZB_BC_NARROW only: raised when a value does not survive the narrowing, or when the job cannot be narrowed
};

// Thread t of `nth` moves its slice of every job, one after the other, through its two pinned slots: the slots keep
// alternating across jobs, so that several arrays (lat + lon, angle + mask) go up in ONE pipelined pass -- no second
// ramp-up (every thread fills a slot before the first DMA can start) and no second round of thread starts and joins.
void bigcopy_part(int dev, int t, int nth, const BcJob* jobs, int njobs, bool h2d, cudaError_t* err) {
    cudaSetDevice(dev);
    cudaError_t e = cudaSuccess;
    int k = 0;
    bool used[2] = {false, false};
    char* pend_dst[2] = {nullptr, nullptr};
    size_t pend_len[2] = {0, 0};
    for (int q = 0; q < njobs && e == cudaSuccess; ++q) {
        const BcJob& J = jobs[q];
        const size_t part = (((J.bytes + nth - 1) / nth) + 4095) & ~(size_t)4095;
        const size_t a = (size_t)t * part;
        if (a >= J.bytes) continue;
        const size_t bytes = J.bytes - a < part ? J.bytes - a : part;
        const bool narrow = h2d && J.pack == GZB_BC_NARROW;
        char* dst = narrow ? J.dst + a / 2 : J.dst + a;
        const char* src = (J.pack > 1 && !narrow) ? J.src : J.src + a;      // packed sources are addressed by element (first_elem below)
        size_t off = 0;
        while (off < bytes && e == cudaSuccess) {
            if (narrow && J.inexact->load(std::memory_order_relaxed)) break;      // somebody found an inexact value: the transfer will be redone
            const size_t len = bytes - off < GZB_BC_SLOT ? bytes - off : GZB_BC_SLOT;
            if (used[k]) {                           // the slot's previous transfer must be over
                e = cudaEventSynchronize(g_bc.ev[t][k]);
                if (!h2d && e == cudaSuccess) memcpy(pend_dst[k], g_bc.slot[t][k], pend_len[k]);
            }
            if (e != cudaSuccess) break;
            if (narrow) {
                if (!narrow_into((float*)g_bc.slot[t][k], (const double*)(src + off), len / 8)) J.inexact->store(1);
                e = cudaMemcpyAsync(dst + off / 2, g_bc.slot[t][k], len / 2, cudaMemcpyHostToDevice, g_bc.stream[t]);
            } else if (h2d) {
                if (J.pack == 8) pack_into<int64_t>((int8_t*)g_bc.slot[t][k], src, a + off, len);
                else if (J.pack == 4) pack_into<int32_t>((int8_t*)g_bc.slot[t][k], src, a + off, len);
                else if (J.pack == 2) pack_into<int16_t>((int8_t*)g_bc.slot[t][k], src, a + off, len);
                else memcpy(g_bc.slot[t][k], src + off, len);
                e = cudaMemcpyAsync(dst + off, g_bc.slot[t][k], len, cudaMemcpyHostToDevice, g_bc.stream[t]);
            } else {
                e = cudaMemcpyAsync(g_bc.slot[t][k], src + off, len, cudaMemcpyDeviceToHost, g_bc.stream[t]);
                pend_dst[k] = dst + off;
                pend_len[k] = len;
            }
            if (e == cudaSuccess) e = cudaEventRecord(g_bc.ev[t][k], g_bc.stream[t]);
            used[k] = true;
            off += len;
            k ^= 1;
        }
    }
    for (int q = 0; q < 2 && e == cudaSuccess; ++q) {
        const int kk = k ^ q;                    // oldest first
        if (!used[kk]) continue;
        e = cudaEventSynchronize(g_bc.ev[t][kk]);
        if (!h2d && e == cudaSuccess) memcpy(pend_dst[kk], g_bc.slot[t][kk], pend_len[kk]);
    }
    *err = e;
}
}  // namespace

// Copies `bytes` between pageable host memory and the device and RETURNS WHEN DONE.  Falls back to
// one cudaMemcpyAsync + stream synchronisation for small or already pinned buffers.
// A transfer is "plain" (one cudaMemcpyAsync) when it is small or when the host side is not pageable memory.
static bool bigcopy_plain(const void* host, size_t bytes, int pack, size_t min_bytes = 0) {
    if (bytes < (min_bytes ? min_bytes : GZB_BC_MIN) && pack <= 1) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) == cudaSuccess) {
        if (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return true;
    } else {
        cudaGetLastError();
    }
    return false;
}

static int bigcopy_one_plain(gzb_ctx* ctx, void* dst, const void* src, size_t bytes, int pack) {
    if (pack > 1) {          // no staging slots: convert on this thread, then one plain copy
        std::vector<int8_t> tmp(bytes);
        if (pack == 8) pack_into<int64_t>(tmp.data(), (const char*)src, 0, bytes);
        else if (pack == 4) pack_into<int32_t>(tmp.data(), (const char*)src, 0, bytes);
        else pack_into<int16_t>(tmp.data(), (const char*)src, 0, bytes);
        GZB_CUDA(ctx, cudaMemcpyAsync(dst, tmp.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
        GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return GZB_OK;
    }
    // cudaMemcpyDefault: `src` of an upload may itself be device memory (a grid replicated from another GPU)
    GZB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GZB_OK;
}

// Moves every job between pageable host memory and the device in one threaded, pipelined pass and RETURNS WHEN DONE.
int gzb_bigcopy_jobs(gzb_ctx* ctx, const BcJob* jobs_in, int njobs_in, bool h2d, bool sync_before) {
    BcJob jobs[4];
    int njobs = 0;
    for (int q = 0; q < njobs_in && q < 4; ++q) {
        const BcJob& J = jobs_in[q];
        if (J.bytes == 0) continue;
        const void* host = h2d ? (const void*)J.src : (const void*)J.dst;
        if (J.pack == GZB_BC_NARROW && (!h2d || !J.inexact || bigcopy_plain(host, J.bytes, 0))) {
            if (J.inexact) J.inexact->store(1);      // not a threaded pageable upload: the caller takes the full-precision path
        } else if (bigcopy_plain(host, J.bytes, J.pack, J.min_bytes)) {
            const int rc = bigcopy_one_plain(ctx, J.dst, J.src, J.bytes, J.pack);
            if (rc != GZB_OK) return rc;
        } else {
            jobs[njobs++] = J;
        }
    }
    if (njobs == 0) return GZB_OK;
    std::unique_lock<std::mutex> lk(g_bc_mu);
    if (!bigcopy_ready(ctx->device)) {
        lk.unlock();
        for (int q = 0; q < njobs; ++q) {
            if (jobs[q].pack == GZB_BC_NARROW) { jobs[q].inexact->store(1); continue; }
            const int rc = bigcopy_one_plain(ctx, jobs[q].dst, jobs[q].src, jobs[q].bytes, jobs[q].pack);
            if (rc != GZB_OK) return rc;
        }
        return GZB_OK;
    }
    // everything queued so far (producers of `src`, users of `dst`) -- unless the caller knows better
    if (sync_before) GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::thread th[GZB_BC_THREADS];
    cudaError_t errs[GZB_BC_THREADS];
    int want = gzb_host_threads(GZB_BC_THREADS);
    if (want > 2 && want == (int)std::thread::hardware_concurrency()) --want;      // leave one core to the caller
    {   // no more threads than slots' worth of data (a 5 MB track array: 5 threads, not 16)
        size_t total = 0;
        for (int q = 0; q < njobs; ++q) total += jobs[q].bytes;
        const size_t by_data = (total + GZB_BC_SLOT - 1) / GZB_BC_SLOT;
        if ((size_t)want > by_data) want = (int)by_data;
    }
    want = want < 2 ? 2 : want;
    for (int t = 0; t < want; ++t) {
        errs[t] = cudaSuccess;
        th[t] = std::thread(bigcopy_part, ctx->device, t, want, (const BcJob*)jobs, njobs, h2d, &errs[t]);
    }
    for (int t = 0; t < want; ++t) th[t].join();
    for (int t = 0; t < want; ++t)
        if (errs[t] != cudaSuccess) return gzb_fail(ctx, GZB_ERR_CUDA, "parallel copy failed: %s", cudaGetErrorString(errs[t]));
    return GZB_OK;
}

// Copies `bytes` between pageable host memory and the device and RETURNS WHEN DONE.  Falls back to
// one cudaMemcpyAsync + stream synchronisation for small or already pinned buffers.
int gzb_bigcopy(gzb_ctx* ctx, void* dst, const void* src, size_t bytes, bool h2d, bool sync_before = true, int pack = 0,
                size_t min_bytes = 0) {
    if (bytes == 0) return GZB_OK;
    const BcJob J = {(char*)dst, (const char*)src, bytes, pack, min_bytes, nullptr};
    return gzb_bigcopy_jobs(ctx, &J, 1, h2d, sync_before);
}

// ======================================================================== device memory pool
// cudaMalloc / cudaFree were measured to stall for 0.1 - 0.7 s now and then on a box that also holds
// a large pinned + mapped host slab (the zero-copy field records), which dwarfs a 0.1 s end-to-end
// call.  Every device buffer of the library therefore comes from a small process-wide pool: a freed
// block is kept (per device, exact size) and handed to the next request of that size -- a second
// context on the same grid, or the next Model2SatTrack call, allocates nothing.  Bounded by
// gzb_pool_config; gzb_pool_trim gives everything back to the driver.
namespace {
struct PoolBlock { int dev; size_t bytes; void* p; };
std::mutex g_pool_mu;
std::vector<PoolBlock> g_pool_free;
std::unordered_map<void*, std::pair<int, size_t>> g_pool_live;
size_t g_pool_cached = 0;
size_t g_pool_cap = (size_t)16 << 30;
int g_pool_on = 1;
}  // namespace

cudaError_t gzb_pool_malloc(void** p, size_t bytes) {
    if (bytes == 0) bytes = 1;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t k = 0; k < g_pool_free.size(); ++k) {
            if (g_pool_free[k].dev == dev && g_pool_free[k].bytes == bytes) {
                *p = g_pool_free[k].p;
                g_pool_cached -= bytes;
                g_pool_free[k] = g_pool_free.back();
                g_pool_free.pop_back();
                g_pool_live[*p] = {dev, bytes};
                return cudaSuccess;
            }
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {          // out of memory: give the cached blocks back and try once more
        cudaGetLastError();
        gzb_pool_trim();
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        g_pool_live[*p] = {dev, bytes};
    }
    return e;
}

cudaError_t gzb_pool_free(void* p) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto it = g_pool_live.find(p);
        if (it != g_pool_live.end()) {
            const int dev = it->second.first;
            const size_t bytes = it->second.second;
            g_pool_live.erase(it);
            if (g_pool_on && g_pool_cached + bytes <= g_pool_cap) {
                g_pool_free.push_back({dev, bytes, p});
                g_pool_cached += bytes;
                return cudaSuccess;
            }
        }
    }
    return cudaFree(p);
}

// The 4 KiB pinned scratch of a context (status words): pinning / unpinning host memory was
// measured to stall like cudaMalloc / cudaFree do, so the blocks are recycled and only released by
// gzb_pool_trim.
namespace {
std::vector<void*> g_scratch_free;
}
cudaError_t gzb_scratch_alloc(void** p) {
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_scratch_free.empty()) {
            *p = g_scratch_free.back();
            g_scratch_free.pop_back();
            return cudaSuccess;
        }
    }
    return cudaHostAlloc(p, 4096, cudaHostAllocPortable);
}
void gzb_scratch_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool_on && g_scratch_free.size() < 64) g_scratch_free.push_back(p);
    else cudaFreeHost(p);
}

extern "C" int gzb_pool_trim(void) {
    {
        std::lock_guard<std::mutex> lk(g_bc_mu);
        if (g_bc.ok) {
            for (int t = 0; t < GZB_BC_THREADS; ++t) {
                for (int k = 0; k < 2; ++k) { cudaFreeHost(g_bc.slot[t][k]); cudaEventDestroy(g_bc.ev[t][k]); }
                cudaStreamDestroy(g_bc.stream[t]);
            }
            g_bc.ok = false;
            g_bc.dev = -1;
        }
    }
    {
        std::vector<void*> sc;
        {
            std::lock_guard<std::mutex> lk(g_pool_mu);
            sc.swap(g_scratch_free);
        }
        for (void* q : sc) cudaFreeHost(q);
    }
    std::vector<PoolBlock> blocks;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        blocks.swap(g_pool_free);
        g_pool_cached = 0;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (const PoolBlock& b : blocks) {
        cudaSetDevice(b.dev);
        cudaFree(b.p);
    }
    cudaSetDevice(cur);
    cudaGetLastError();
    return GZB_OK;
}

extern "C" int gzb_pool_config(int enabled, uint64_t max_cached_bytes) {
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        g_pool_on = enabled ? 1 : 0;
        g_pool_cap = (size_t)max_cached_bytes;
    }
    if (!enabled) gzb_pool_trim();
    return GZB_OK;
}

extern "C" int gzb_pool_stats(uint64_t* cached_bytes, uint64_t* cached_blocks, uint64_t* live_blocks) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (cached_bytes) *cached_bytes = g_pool_cached;
    if (cached_blocks) *cached_blocks = g_pool_free.size();
    if (live_blocks) *live_blocks = g_pool_live.size();
    return GZB_OK;
}


// ======================================================================== workspace arena
int gzb_arena_reserve(gzb_ctx* ctx, size_t bytes) {
    ctx->arena.off = 0;
    if (bytes <= ctx->arena.cap) return GZB_OK;
    if (ctx->arena.base) {
        GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        GZB_CUDA(ctx, gzb_pool_free(ctx->arena.base));
        ctx->arena.base = nullptr;
        ctx->arena.cap = 0;
    }
    size_t want = bytes + bytes / 4 + (size_t(1) << 20);
    cudaError_t e = gzb_pool_malloc((void**)&ctx->arena.base, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return gzb_fail(ctx, GZB_ERR_NOMEM, "workspace of %zu bytes: %s", want, cudaGetErrorString(e));
    }
    ctx->arena.cap = want;
    return GZB_OK;
}

void* gzb_arena_take(gzb_ctx* ctx, size_t bytes) {
    size_t n = gzb_align(bytes ? bytes : 1);
    if (ctx->arena.off + n > ctx->arena.cap) return nullptr;
    void* p = ctx->arena.base + ctx->arena.off;
    ctx->arena.off += n;
    return p;
}

// ======================================================================== pyramid build
int gzb_build_cert(gzb_ctx* ctx);   // gzb_nearest.cu

// fp32 storage of a sphere computed in fp64: the radius is rounded up and padded by 2.5e-7 so
// that it still bounds the members after the centre itself was rounded to float
__device__ __forceinline__ float4 pack_sphere(double cx, double cy, double cz, double r) {
    return make_float4((float)cx, (float)cy, (float)cz, __double2float_ru(r + 2.5e-7));
}

// One warp per leaf tile: unit vectors of its 32 cells + the tile's bounding sphere.
__global__ void __launch_bounds__(256)
k_build_leaves(GzbGrid g, float* __restrict__ cells, float4* __restrict__ nodes0, float4* __restrict__ cellv,
               int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const GzbLevel L0 = g.lev[0];
    const int64_t ntile = (int64_t)L0.nJ * L0.nI;
    if (tile >= ntile) return;
    const int tJ = (int)(tile / L0.nI), tI = (int)(tile % L0.nI);
    const int64_t j = (int64_t)tJ * GZB_TILE_J + (lane >> 3);
    const int64_t i = (int64_t)tI * GZB_TILE_I + (lane & 7);
    const bool ok = (j < g.Nj) && (i < g.Ni);
    double x = 0.0, y = 0.0, z = 0.0;
    if (ok) {
        const double la = g.lat[j * g.Ni + i] * GZB_D2R;
        const double lo = g.lon[j * g.Ni + i] * GZB_D2R;
        const double cl = cos(la);
        x = cl * cos(lo);
        y = cl * sin(lo);
        z = sin(la);
        if (!(isfinite(la) && isfinite(lo))) atomicOr(bad, 1);
    }
    float* c = cells + tile * 96;
    c[lane] = ok ? (float)x : 1.0e3f;
    c[32 + lane] = ok ? (float)y : 1.0e3f;
    c[64 + lane] = ok ? (float)z : 1.0e3f;
    if (ok && cellv) cellv[j * g.Ni + i] = make_float4((float)x, (float)y, (float)z, 0.0f);
    // centre = normalised mean of the valid cells
    double sx = x, sy = y, sz = z;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    double nrm = sqrt(sx * sx + sy * sy + sz * sz);
    if (!(nrm > 1.0e-6)) {   // degenerate mean: fall back to the first cell of the tile
        sx = __shfl_sync(0xffffffffu, x, 0);
        sy = __shfl_sync(0xffffffffu, y, 0);
        sz = __shfl_sync(0xffffffffu, z, 0);
        nrm = 1.0;
    }
    const double cx = sx / nrm, cy = sy / nrm, cz = sz / nrm;
    double r = 0.0;
    if (ok) {
        const double dx = x - cx, dy = y - cy, dz = z - cz;
        r = sqrt(dx * dx + dy * dy + dz * dz);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    r = r * (1.0 + 1.0e-12) + 1.0e-15;
    if (lane == 0) nodes0[gzb_leaf_index(g, tJ, tI)] = pack_sphere(cx, cy, cz, r);
}

// One warp per node of level `l` (l >= 1): sphere around its (up to 32) children.
__global__ void __launch_bounds__(256)
k_build_level(GzbGrid g, int l, int is_top) {
    const int lane = threadIdx.x & 31;
    const int64_t node = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const GzbLevel L = g.lev[l];
    const int64_t nn = (int64_t)L.nJ * L.nI;
    if (node >= nn) return;
    const int J = (int)(node / L.nI), I = (int)(node % L.nI);
    const float4 chf = g.lev[l - 1].nodes[node * 32 + lane];
    const bool ok = chf.w >= 0.0f;
    const double chx = chf.x, chy = chf.y, chz = chf.z, chw = chf.w;
    double sx = ok ? chx : 0.0, sy = ok ? chy : 0.0, sz = ok ? chz : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    double nrm = sqrt(sx * sx + sy * sy + sz * sz);
    const unsigned okm = __ballot_sync(0xffffffffu, ok);
    if (!(nrm > 1.0e-6)) {
        const int f = __ffs(okm) - 1;
        sx = __shfl_sync(0xffffffffu, chx, f);
        sy = __shfl_sync(0xffffffffu, chy, f);
        sz = __shfl_sync(0xffffffffu, chz, f);
        nrm = 1.0;
    }
    const double cx = sx / nrm, cy = sy / nrm, cz = sz / nrm;
    double r = 0.0;
    if (ok) {
        const double dx = chx - cx, dy = chy - cy, dz = chz - cz;
        r = sqrt(dx * dx + dy * dy + dz * dz) + chw;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    r = r * (1.0 + 1.0e-12) + 1.0e-15;
    if (lane == 0) {
        int64_t idx;
        if (is_top) {
            idx = node;
        } else {
            const GzbLevel P = g.lev[l + 1];
            const int pJ = J / P.bj, pI = I / P.bi;
            idx = ((int64_t)pJ * P.nI + pI) * 32 + (J % P.bj) * P.bi + (I % P.bi);
        }
        L.nodes[idx] = pack_sphere(cx, cy, cz, r);
    }
}

// float32 -> float64 of the narrowed uploads (exact by construction: the host checked every value)
__global__ void __launch_bounds__(256)
k_widen2(const int64_t n, const float* __restrict__ a32, const float* __restrict__ b32, double* __restrict__ a, double* __restrict__ b) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        a[k] = (double)a32[k];
        b[k] = (double)b32[k];
    }
}

__global__ void __launch_bounds__(256)
k_interleave(const int64_t n, const double* __restrict__ lat, const double* __restrict__ lon, double2* __restrict__ ll) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) ll[k] = make_double2(lat[k], lon[k]);
}

__global__ void k_fill_empty(float4* p, int64_t n) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) p[k] = make_float4(0.0f, 0.0f, 0.0f, -1.0f);
}


// ======================================================================== seed look-up table
// order-preserving map double -> uint64 so that atomicMin/atomicMax work on coordinates
__device__ __forceinline__ unsigned long long ord_enc(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
static double ord_dec(unsigned long long u) {
    const unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// extent of the grid: min/max of lat, of lon % 360 (frame A) and of (lon + 180) % 360 (frame B)
// out[0..5] = ord_enc(min lat, max lat, min A, max A, min B, max B)
__global__ void __launch_bounds__(256)
k_extent(const int64_t n, const double* __restrict__ lat, const double* __restrict__ lon,
         unsigned long long* __restrict__ out) {
    __shared__ double sm[6][8];
    double v[6] = {INFINITY, -INFINITY, INFINITY, -INFINITY, INFINITY, -INFINITY};
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double la = lat[k], lo = lon[k];
        if (!(isfinite(la) && isfinite(lo))) continue;
        const double a = gzb_pymod(lo, 360.0), b = gzb_pymod(lo + 180.0, 360.0);
        v[0] = fmin(v[0], la); v[1] = fmax(v[1], la);
        v[2] = fmin(v[2], a);  v[3] = fmax(v[3], a);
        v[4] = fmin(v[4], b);  v[5] = fmax(v[5], b);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v[q], o);
            v[q] = (q & 1) ? fmax(v[q], u) : fmin(v[q], u);
        }
        if (lane == 0) sm[q][warp] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int q = threadIdx.x;
        double r = sm[q][0];
        for (int w = 1; w < 8; ++w) r = (q & 1) ? fmax(r, sm[q][w]) : fmin(r, sm[q][w]);
        if (isfinite(r)) {
            if (q & 1) atomicMax(out + q, ord_enc(r)); else atomicMin(out + q, ord_enc(r));
        }
    }
}

// A bin keeps, among the cells that fall into it, the one NEAREST TO ITS CENTRE (a target in the bin
// is then rarely more than one hill-climb move away from its answer).  Two passes: smallest distance
// code per bin, then the smallest flat index attaining it.
__device__ __forceinline__ unsigned lut_dist_code(const GzbLut& L, const long long bin, const double la, const double lo) {
    const long long bj = bin / L.nlon, bi = bin - bj * L.nlon;
    const double d = 1.0 / L.inv_d;
    const double latc = L.lat0 + ((double)bj + 0.5) * d;
    const double lonc = L.lon0 + ((double)bi + 0.5) * d;
    double dx = gzb_pymod(lo - lonc + 180.0, 360.0) - 180.0;
    dx *= cos(latc * GZB_D2R);
    const double dy = la - latc;
    return __float_as_uint((float)(dx * dx + dy * dy));      // non-negative float: the bit pattern is ordered
}

template <int PASS>
__global__ void __launch_bounds__(256)
k_lut_scatter(const GzbGrid g, const GzbLut L, unsigned* __restrict__ best, unsigned* __restrict__ bins) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.Nj * g.Ni) return;
    const double la = g.lat[k], lo = g.lon[k];
    if (!(isfinite(la) && isfinite(lo))) return;
    const long long bin = gzb_lut_bin(L, la, lo);
    const unsigned code = lut_dist_code(L, bin, la, lo);
    if (PASS == 0) atomicMin(best + bin, code);
    else if (best[bin] == code) atomicMin(bins + bin, (unsigned)k);
}

// empty bins take the value of the first non-empty neighbour at distance `step` (8 directions)
__global__ void __launch_bounds__(256)
k_lut_fill(const GzbLut L, const unsigned* __restrict__ in, unsigned* __restrict__ out, const int step) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (int64_t)L.nlat * L.nlon) return;
    unsigned v = in[k];
    if (v == 0xffffffffu) {
        const int bj = (int)(k / L.nlon), bi = (int)(k - (int64_t)bj * L.nlon);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int dj = (q < 3) ? -1 : (q < 5 ? 0 : 1);
            const int di = (q == 0 || q == 3 || q == 5) ? -1 : ((q == 1 || q == 6) ? 0 : 1);
            const int nj = bj + dj * step;
            int ni = bi + di * step;
            if (nj < 0 || nj >= L.nlat) continue;
            if (L.wrap) {
                ni %= L.nlon;
                if (ni < 0) ni += L.nlon;
            } else if (ni < 0 || ni >= L.nlon) {
                continue;
            }
            const unsigned u = in[(int64_t)nj * L.nlon + ni];
            if (u != 0xffffffffu) { v = u; break; }
        }
    }
    out[k] = v;
}

// Per-cell certificates (GzbGrid::cellv[].w, two bfloat16): lower bounds of the chord from cell c to
// every cell OUTSIDE its 5x5 index neighbourhood, and outside its 3x3 one.  One warp per leaf tile T, lane = cell of T:
//   cells inside T's 5x3-tile window  -> evaluated one by one (tiles too far to matter skipped);
//   cells outside that window         -> cert(T) - |c - centre(T)|   (triangle inequality).
// The inner loop (every cell of the tile against every cell of a window tile) is what a grid context costs on the GPU
// (4.1 of 6.2 ms on ORCA12), so it is kept lean: the candidates are broadcast from shared memory (one 128-bit load
// instead of three shuffles), and which candidates lie inside a lane's 5x5 / 3x3 neighbourhood is one bit test on a mask
// the lane builds once per window tile.
__device__ __forceinline__ unsigned cc_inside_mask(const int A, const int B, const int h) {
    // bit k = (kj, ki) of a 4 x 8 tile set iff |A + kj| <= h and |B + ki| <= h
    unsigned rows = 0, cols = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) rows |= (unsigned)(A + q >= -h && A + q <= h) << (8 * q);
#pragma unroll
    for (int q = 0; q < 8; ++q) cols |= (unsigned)(B + q >= -h && B + q <= h) << q;
    return rows * cols;      // rows holds one bit per byte: the product copies `cols` into the bytes of the rows that are in
}

__global__ void __launch_bounds__(256)
k_cellcert(const GzbGrid g, float4* __restrict__ cellv) {
    __shared__ float4 sq[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const GzbLevel L0 = g.lev[0];
    const int64_t ntile = (int64_t)L0.nJ * L0.nI;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= ntile) return;
    const int tJ = (int)(tile / L0.nI), tI = (int)(tile - (int64_t)tJ * L0.nI);
    const float* __restrict__ cc = g.cells + tile * 96;
    const float x = cc[lane], y = cc[32 + lane], z = cc[64 + lane];
    const int lj = lane >> 3, li = lane & 7;
    const int64_t j = (int64_t)tJ * GZB_TILE_J + lj;
    const int64_t i = (int64_t)tI * GZB_TILE_I + li;
    const bool ok = (j < g.Nj) && (i < g.Ni);
    const int64_t tidx = gzb_leaf_index(g, tJ, tI);
    const float4 nd = L0.nodes[tidx];
    float far;
    {
        const float dx = x - nd.x, dy = y - nd.y, dz = z - nd.z;
        const float a = sqrtf(dx * dx + dy * dy + dz * dz);
        far = g.cert[tidx] - (a * (1.0f + 2.0e-7f) + 5.0e-7f);
    }
    float m2 = INFINITY;     // smallest chord^2 to a window cell outside the 5x5 neighbourhood
    float m3 = INFINITY;     // ... outside the 3x3 neighbourhood (m3 <= m2)
    // nearest tiles first, so that the far ones can be skipped
    const int order_j[5] = {0, -1, 1, -2, 2};
    const int order_i[3] = {0, -1, 1};
    for (int a = 0; a < 5; ++a) {
        for (int b = 0; b < 3; ++b) {
            const int wJ = tJ + order_j[a], wI = tI + order_i[b];
            if (wJ < 0 || wJ >= L0.nJ || wI < 0 || wI >= L0.nI) continue;
            const float4 wn = L0.nodes[gzb_leaf_index(g, wJ, wI)];
            if (wn.w < 0.0f) continue;
            {
                const float dx = x - wn.x, dy = y - wn.y, dz = z - wn.z;
                const float dc = sqrtf(dx * dx + dy * dy + dz * dz) - wn.w;     // lower bound to any cell of W
                const bool skip = !ok || (dc > 0.0f && dc * dc * (1.0f - 1.0e-6f) > m2);
                if (__all_sync(0xffffffffu, skip)) continue;
            }
            const float* __restrict__ wc = g.cells + ((int64_t)wJ * L0.nI + wI) * 96;
            __syncwarp();
            sq[warp][lane] = make_float4(wc[lane], wc[32 + lane], wc[64 + lane], 0.0f);
            __syncwarp();
            // candidate k = (kj, ki) of W sits at index offset (4 dJ + kj - lj, 8 dI + ki - li) from this lane's cell
            const int A = GZB_TILE_J * order_j[a] - lj, B = GZB_TILE_I * order_i[b] - li;
            const unsigned in5 = cc_inside_mask(A, B, 2), in3 = cc_inside_mask(A, B, 1);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float4 u = sq[warp][k];
                const float dx = x - u.x, dy = y - u.y, dz = z - u.z;
                const float d2 = dx * dx + dy * dy + dz * dz;     // padded cells sit at 1e3: never the minimum
                if (!((in5 >> k) & 1u)) m2 = fminf(m2, d2);
                if (!((in3 >> k) & 1u)) m3 = fminf(m3, d2);
            }
        }
    }
    if (ok) {
        float v = fminf(far, sqrtf(m2) - 5.0e-7f) * (1.0f - 1.0e-6f);
        if (!(v > 0.0f)) v = 0.0f;        // also catches NaN
        float v3 = fminf(far, sqrtf(m3) - 5.0e-7f) * (1.0f - 1.0e-6f);
        if (!(v3 > 0.0f)) v3 = 0.0f;
        // both certificates in one word, each truncated (towards zero: still a lower bound) to bfloat16:
        // high half = 3x3 certificate, low half = 5x5 certificate
        cellv[j * g.Ni + i].w = __uint_as_float((__float_as_uint(v3) & 0xffff0000u) | (__float_as_uint(v) >> 16));
    }
}

// host side: geometry of the table from the extent, allocation, scatter and hole filling
static int build_lut(gzb_ctx* ctx, const unsigned long long* ext) {
    GzbGrid& g = ctx->g;
    g.lut.bins = nullptr;
    const double n = (double)g.Nj * (double)g.Ni;
    if (n >= 4294967295.0) return GZB_OK;                    // flat indices must fit 32 bits
    const double lat_min = ord_dec(ext[0]), lat_max = ord_dec(ext[1]);
    ctx->lat_min = lat_min;
    ctx->lat_max = lat_max;
    const double a_min = ord_dec(ext[2]), a_max = ord_dec(ext[3]);
    const double b_min = ord_dec(ext[4]), b_max = ord_dec(ext[5]);
    if (!(lat_max >= lat_min) || !(a_max >= a_min) || !(b_max >= b_min)) return GZB_OK;
    GzbLut L;
    const double span_a = a_max - a_min, span_b = b_max - b_min;
    if (fmin(span_a, span_b) > 300.0) {
        L.wrap = 1; L.lon0 = 0.0; L.lon_span = 360.0;
    } else if (span_a <= span_b) {
        L.wrap = 0; L.lon0 = a_min; L.lon_span = span_a;
    } else {
        L.wrap = 0; L.lon0 = b_min - 180.0; L.lon_span = span_b;
    }
    const double lat_span = lat_max - lat_min;
    L.lat0 = lat_min;
    double d = sqrt(fmax(lat_span, 1.0e-9) * fmax(L.lon_span, 1.0e-9) / n);     // ~ one cell per bin
    if (!(d > 1.0e-7)) d = 1.0e-7;
    for (;;) {
        const double nb = ceil(fmax(lat_span / d, 1.0)) * ceil(fmax(L.lon_span / d, 1.0));
        if (nb <= 134217728.0) break;
        d *= 1.25;
    }
    L.nlat = (int)fmax(1.0, ceil(lat_span / d));
    L.nlon = (int)fmax(1.0, ceil(L.lon_span / d));
    L.inv_d = 1.0 / d;
    const size_t nb = (size_t)L.nlat * (size_t)L.nlon;
    for (int k = 0; k < 2; ++k) {
        cudaError_t e = gzb_pool_malloc((void**)&ctx->d_lut[k], nb * sizeof(unsigned));
        if (e != cudaSuccess) {             // not fatal: the warp walk needs no table
            cudaGetLastError();
            for (int q = 0; q < 2; ++q) { gzb_pool_free(ctx->d_lut[q]); ctx->d_lut[q] = nullptr; }
            return GZB_OK;
        }
    }
    cudaStream_t s = ctx->stream;
    GZB_CUDA(ctx, cudaMemsetAsync(ctx->d_lut[0], 0xff, nb * sizeof(unsigned), s));
    GZB_CUDA(ctx, cudaMemsetAsync(ctx->d_lut[1], 0xff, nb * sizeof(unsigned), s));
    L.bins = nullptr;
    const int64_t cells = g.Nj * g.Ni;
    k_lut_scatter<0><<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(g, L, ctx->d_lut[1], ctx->d_lut[0]);
    GZB_LAUNCH_CHECK(ctx);
    k_lut_scatter<1><<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(g, L, ctx->d_lut[1], ctx->d_lut[0]);
    GZB_LAUNCH_CHECK(ctx);
    int cur = 0;
    const int steps[6] = {1, 1, 2, 4, 8, 16};
    for (int q = 0; q < 6; ++q) {
        k_lut_fill<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(L, ctx->d_lut[cur], ctx->d_lut[cur ^ 1], steps[q]);
        GZB_LAUNCH_CHECK(ctx);
        cur ^= 1;
    }
    L.bins = ctx->d_lut[cur];
    g.lut = L;
    ctx->stats.lut_bins = (int64_t)nb;
    return GZB_OK;
}

static int build_pyramid(gzb_ctx* ctx) {
    GzbGrid& g = ctx->g;
    // level geometry
    int l = 0;
    g.lev[0].nJ = (int)((g.Nj + GZB_TILE_J - 1) / GZB_TILE_J);
    g.lev[0].nI = (int)((g.Ni + GZB_TILE_I - 1) / GZB_TILE_I);
    g.lev[0].bj = GZB_TILE_J;
    g.lev[0].bi = GZB_TILE_I;
    g.lev[0].sJ = GZB_TILE_J;
    g.lev[0].sI = GZB_TILE_I;
    g.lev[0].sbj = 2;
    g.lev[0].sbi = 3;
    while ((int64_t)g.lev[l].nJ * g.lev[l].nI > 32) {
        if (l + 1 >= GZB_MAXLEV) return gzb_fail(ctx, GZB_ERR_ARG, "grid too large for %d pyramid levels", GZB_MAXLEV);
        GzbLevel& P = g.lev[l + 1];
        P.bj = ((l + 1) & 1) ? 8 : 4;
        P.bi = ((l + 1) & 1) ? 4 : 8;
        P.sbj = ((l + 1) & 1) ? 3 : 2;
        P.sbi = ((l + 1) & 1) ? 2 : 3;
        P.nJ = (g.lev[l].nJ + P.bj - 1) / P.bj;
        P.nI = (g.lev[l].nI + P.bi - 1) / P.bi;
        P.sJ = g.lev[l].sJ * P.bj;
        P.sI = g.lev[l].sI * P.bi;
        ++l;
    }
    g.nlev = l + 1;
    ctx->stats.pyramid_levels = g.nlev;
    // storage
    const int64_t ntile = (int64_t)g.lev[0].nJ * g.lev[0].nI;
    GZB_CUDA(ctx, gzb_pool_malloc((void**)&ctx->d_cells, (size_t)ntile * 96 * sizeof(float)));
    g.cells = ctx->d_cells;
    for (int k = 0; k < g.nlev; ++k) {
        int64_t n = (k == g.nlev - 1) ? 32 : (int64_t)g.lev[k + 1].nJ * g.lev[k + 1].nI * 32;
        GZB_CUDA(ctx, gzb_pool_malloc((void**)&ctx->d_nodes[k], (size_t)n * sizeof(float4)));
        g.lev[k].nodes = ctx->d_nodes[k];
        k_fill_empty<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_nodes[k], n);
        GZB_LAUNCH_CHECK(ctx);
        if (k == 0) {
            GZB_CUDA(ctx, gzb_pool_malloc((void**)&ctx->d_cert, (size_t)n * sizeof(float)));
            GZB_CUDA(ctx, cudaMemsetAsync(ctx->d_cert, 0, (size_t)n * sizeof(float), ctx->stream));
            g.cert = ctx->d_cert;
        }
    }
    {   // per-cell vectors + certificates of the per-thread search (optional: skipped when out of memory)
        cudaError_t e = gzb_pool_malloc((void**)&ctx->d_cellv, (size_t)g.Nj * (size_t)g.Ni * sizeof(float4));
        if (e != cudaSuccess) { cudaGetLastError(); ctx->d_cellv = nullptr; }
        g.cellv = ctx->d_cellv;
    }
    unsigned long long* d_ext = (unsigned long long*)(ctx->d_flags + 48);     // 6 words at byte 192, 8-byte aligned
    {
        unsigned long long* h_ext = (unsigned long long*)((char*)ctx->pinned + 256);
        for (int q = 0; q < 6; ++q) h_ext[q] = (q & 1) ? 0ull : ~0ull;
        GZB_CUDA(ctx, cudaMemcpyAsync(d_ext, h_ext, 48, cudaMemcpyHostToDevice, ctx->stream));
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
        k_extent<<<nsm * 8, 256, 0, ctx->stream>>>(g.Nj * g.Ni, g.lat, g.lon, d_ext);
        GZB_LAUNCH_CHECK(ctx);
        GZB_CUDA(ctx, cudaMemcpyAsync(h_ext, d_ext, 48, cudaMemcpyDeviceToHost, ctx->stream));
    }
    int* d_bad = (int*)ctx->d_cert;     // the certificate array is filled last; borrow its first word
    k_build_leaves<<<(unsigned)((ntile + 7) / 8), 256, 0, ctx->stream>>>(g, ctx->d_cells, ctx->d_nodes[0], ctx->d_cellv, d_bad);
    GZB_LAUNCH_CHECK(ctx);
    GZB_CUDA(ctx, cudaMemcpyAsync((char*)ctx->pinned + 128, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(int*)((char*)ctx->pinned + 128)) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_create: grid coordinates must be finite");
    GZB_CUDA(ctx, cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    for (int k = 1; k < g.nlev; ++k) {
        const int64_t nn = (int64_t)g.lev[k].nJ * g.lev[k].nI;
        k_build_level<<<(unsigned)((nn + 7) / 8), 256, 0, ctx->stream>>>(g, k, k == g.nlev - 1 ? 1 : 0);
        GZB_LAUNCH_CHECK(ctx);
    }
    int rc = gzb_build_cert(ctx);
    if (rc != GZB_OK) return rc;
    if (ctx->d_cellv) {
        if ((rc = build_lut(ctx, (const unsigned long long*)((char*)ctx->pinned + 256))) != GZB_OK) return rc;
        k_cellcert<<<(unsigned)((ntile + 7) / 8), 256, 0, ctx->stream>>>(g, ctx->d_cellv);
        GZB_LAUNCH_CHECK(ctx);
    }
    return GZB_OK;
}

// ======================================================================== create / destroy
static int create_impl(int device, int64_t nj, int64_t ni, const double* lat, const double* lon, const double* angle_or_null,
                       const void* mask_or_null, int mask_itemsize, int k_ew_per, gzb_ctx** ctx_out);

extern "C" int gzb_create(int device, int64_t nj, int64_t ni, const double* lat, const double* lon,
                          const double* angle_or_null, const int8_t* mask_or_null, int k_ew_per,
                          gzb_ctx** ctx_out) {
    return create_impl(device, nj, ni, lat, lon, angle_or_null, mask_or_null, 1, k_ew_per, ctx_out);
}

// gzb_create with the land-sea mask in the caller's integer type (HOST memory, 1 / 2 / 4 / 8 bytes per cell): the reference's
// MG.mask is int64; converting it to int8 rides on the upload's staging copies instead of being a pass of its own
extern "C" int gzb_create_imask(int device, int64_t nj, int64_t ni, const double* lat, const double* lon,
                                const double* angle_or_null, const void* mask_or_null, int mask_itemsize, int k_ew_per,
                                gzb_ctx** ctx_out) {
    if (mask_or_null && mask_itemsize != 1 && mask_itemsize != 2 && mask_itemsize != 4 && mask_itemsize != 8)
        return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_create_imask: mask_itemsize must be 1, 2, 4 or 8 (got %d)", mask_itemsize);
    return create_impl(device, nj, ni, lat, lon, angle_or_null, mask_or_null, mask_itemsize, k_ew_per, ctx_out);
}

static int create_impl(int device, int64_t nj, int64_t ni, const double* lat, const double* lon, const double* angle_or_null,
                       const void* mask_or_null, int mask_itemsize, int k_ew_per, gzb_ctx** ctx_out) {
    if (!ctx_out) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_create: null ctx_out");
    *ctx_out = nullptr;
    if (nj < 3 || ni < 3 || !lat || !lon)
        return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_create: need lat/lon of shape (nj>=3, ni>=3)");
    if (nj * ni > ((int64_t)1 << 36) - 2 || nj > 0x3fffffff || ni > 0x3fffffff) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_create: grid too large");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return gzb_fail(nullptr, GZB_ERR_CUDA,
                        "gzb_create: no CUDA device (%s); this library has no CPU fallback",
                        e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev)
        return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_create: device %d out of range (have %d)", device, ndev);
    gzb_ctx* ctx = new (std::nothrow) gzb_ctx();
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_NOMEM, "gzb_create: host allocation failed");
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    memset(&ctx->g, 0, sizeof(ctx->g));
    ctx->err[0] = 0;
    ctx->device = device;
    int rc = GZB_OK;
    do {
        if ((rc = gzb_use_device(ctx)) != GZB_OK) break;
#define CK(call) if ((e = (call)) != cudaSuccess) { rc = gzb_fail(ctx, GZB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e)); break; }
        CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->push_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 9; ++k) CK(cudaEventCreateWithFlags(&ctx->ev_push[k], cudaEventDisableTiming));
        {   // the chain part of K1 is latency-bound and small: it must get SM slots ahead of the bulk kernels
            int least = 0, greatest = 0;
            CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            const char* pr = getenv("GZB_AUX_PRIORITY");          // experiment knob: 0 = same priority as the bulk stream
            CK(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, (pr && pr[0] == '0') ? least : greatest));
        }
        CK(cudaEventCreateWithFlags(&ctx->ev_k1[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_k1[1], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_k1[2], cudaEventDisableTiming));
        ctx->stream = ctx->own_stream;
        bool evfail = false;
        for (int k = 0; k < 4; ++k)
            if ((e = cudaEventCreateWithFlags(&ctx->ev[k], cudaEventDisableTiming)) != cudaSuccess) evfail = true;
        if (evfail) { rc = gzb_fail(ctx, GZB_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e)); break; }
        const size_t n = (size_t)nj * (size_t)ni;
        // GZB_CREATE_TIMING=1: host-side time marks of the phases below on stderr (where a context's milliseconds go)
        const bool ctime = getenv("GZB_CREATE_TIMING") != nullptr;
        const auto tc0 = std::chrono::steady_clock::now();
        auto cmark = [&](const char* what) {
            if (ctime) fprintf(stderr, "[gzb create] %7.2f ms  %s\n",
                               1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - tc0).count(), what);
        };
        CK(gzb_pool_malloc((void**)&ctx->d_lat, n * sizeof(double)));
        CK(gzb_pool_malloc((void**)&ctx->d_lon, n * sizeof(double)));
        CK(gzb_pool_malloc((void**)&ctx->d_ll, n * sizeof(double2)));
        {
            // first as float32 (half the bytes; exact for grids that were written in single precision), staged in the
            // memory of `ll` and widened on the device; in full precision when a value does not survive the narrowing
            bool narrowed = false;
            if (n * sizeof(double) >= GZB_BC_MIN && !getenv("GZB_NO_NARROW_UPLOAD")) {
                float* st = (float*)ctx->d_ll;
                std::atomic<int> inexact{0};
                const BcJob up32[2] = {{(char*)st, (const char*)lat, n * sizeof(double), GZB_BC_NARROW, 0, &inexact},
                                       {(char*)(st + n), (const char*)lon, n * sizeof(double), GZB_BC_NARROW, 0, &inexact}};
                if ((rc = gzb_bigcopy_jobs(ctx, up32, 2, true, true)) != GZB_OK) break;
                if (!inexact.load()) {
                    k_widen2<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((int64_t)n, st, st + n, ctx->d_lat, ctx->d_lon);
                    ctx->stats.kernel_launches++;
                    ctx->stats.h2d_bytes += 2 * n * sizeof(float);
                    narrowed = true;
                    cmark("lat, lon uploaded as float32 (exact)");
                }
            }
            if (!narrowed) {
                const BcJob up[2] = {{(char*)ctx->d_lat, (const char*)lat, n * sizeof(double), 0, 0, nullptr},
                                     {(char*)ctx->d_lon, (const char*)lon, n * sizeof(double), 0, 0, nullptr}};
                if ((rc = gzb_bigcopy_jobs(ctx, up, 2, true, true)) != GZB_OK) break;
                ctx->stats.h2d_bytes += 2 * n * sizeof(double);
                cmark("lat, lon uploaded");
            }
        }
        k_interleave<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((int64_t)n, ctx->d_lat, ctx->d_lon, ctx->d_ll);
        ctx->stats.kernel_launches++;
        if (angle_or_null) CK(gzb_pool_malloc((void**)&ctx->d_angle, n * sizeof(double)));
        if (mask_or_null) CK(gzb_pool_malloc((void**)&ctx->d_mask, n));
        CK(gzb_pool_malloc((void**)&ctx->d_flags, 256));
        CK(cudaMemsetAsync(ctx->d_flags, 0, 256, ctx->stream));
        CK(gzb_scratch_alloc(&ctx->pinned));
        memset(ctx->pinned, 0, 4096);
        ctx->pinned_cap = 4096;
#undef CK
        ctx->g.chk = ctx->d_flags + 4;
        ctx->g.kt = nullptr;
#ifdef GZB_KTIME
        if (gzb_pool_malloc((void**)&ctx->d_kt, 64 * 8) == cudaSuccess) {
            cudaMemsetAsync(ctx->d_kt, 0xff, 32 * 8, ctx->stream);
            cudaMemsetAsync(ctx->d_kt + 32, 0, 32 * 8, ctx->stream);
            ctx->g.kt = ctx->d_kt;
        } else cudaGetLastError();
#endif
        ctx->g.Nj = nj;
        ctx->g.Ni = ni;
        ctx->g.kew = k_ew_per;
        ctx->g.lat = ctx->d_lat;
        ctx->g.lon = ctx->d_lon;
        ctx->g.ll = ctx->d_ll;
        ctx->g.angle = ctx->d_angle;
        ctx->g.mask = ctx->d_mask;
        if ((rc = build_pyramid(ctx)) != GZB_OK) break;
        cmark("search structures enqueued");
        if (ctime && getenv("GZB_CREATE_TIMING")[0] == '2') { cudaStreamSynchronize(ctx->stream); cmark("search structures built (synchronised: GZB_CREATE_TIMING=2)"); }
        // the angle and the mask are not needed by the search structures: their upload runs on the host
        // threads while the GPU finishes the certificates and the seed table
        {
            BcJob up[2];
            int nup = 0;
            if (angle_or_null) {
                up[nup++] = {(char*)ctx->d_angle, (const char*)angle_or_null, n * sizeof(double), 0, 0, nullptr};
                ctx->stats.h2d_bytes += n * sizeof(double);
            }
            if (mask_or_null) {
                up[nup++] = {(char*)ctx->d_mask, (const char*)mask_or_null, n, mask_itemsize, 0, nullptr};
                ctx->stats.h2d_bytes += n;
            }
            if (nup && (rc = gzb_bigcopy_jobs(ctx, up, nup, true, false)) != GZB_OK) break;
        }
        cmark("angle, mask uploaded");
        if ((rc = gzb_build_mask_bits(ctx)) != GZB_OK) break;
        e = cudaStreamSynchronize(ctx->stream);
        cmark("device idle: context ready");
        if (e != cudaSuccess) { rc = gzb_fail(ctx, GZB_ERR_CUDA, "grid upload / pyramid build: %s", cudaGetErrorString(e)); break; }
    } while (0);
    if (rc != GZB_OK) {
        char keep[512];
        strncpy(keep, ctx->err, sizeof(keep));
        gzb_destroy(ctx);
        gzb_set_global_error(keep);
        return rc;
    }
    *ctx_out = ctx;
    return GZB_OK;
}

extern "C" int gzb_destroy(gzb_ctx* ctx) {
    if (!ctx) return GZB_OK;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    if (ctx->stream && ctx->stream != ctx->own_stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->push_stream) cudaStreamSynchronize(ctx->push_stream);
    gzb_pool_free(ctx->d_lat);
    gzb_pool_free(ctx->d_lon);
    gzb_pool_free(ctx->d_ll);
    gzb_pool_free(ctx->d_angle);
    gzb_pool_free(ctx->d_mask);
    gzb_pool_free(ctx->d_cells);
    gzb_pool_free(ctx->d_cert);
    gzb_pool_free(ctx->d_cellv);
    gzb_pool_free(ctx->d_lut[0]);
    gzb_pool_free(ctx->d_lut[1]);
    gzb_pool_free(ctx->d_hdg);
    gzb_pool_free(ctx->d_mbits);
    gzb_pool_free(ctx->d_rowsafe);
    gzb_pool_free(ctx->d_flags);
    for (int k = 0; k < GZB_MAXLEV; ++k) gzb_pool_free(ctx->d_nodes[k]);
    gzb_pool_free(ctx->arena.base);
    gzb_scratch_free(ctx->pinned);
    for (int k = 0; k < 2; ++k)
        if (ctx->stage[k]) cudaFreeHost(ctx->stage[k]);
    for (int k = 0; k < 4; ++k)
        if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->push_stream) cudaStreamDestroy(ctx->push_stream);
    for (int k = 0; k < GZB_MAX_PEERS; ++k) {
        if (ctx->peer_stream[k]) { cudaStreamSynchronize(ctx->peer_stream[k]); cudaStreamDestroy(ctx->peer_stream[k]); }
        if (ctx->ev_peer[k]) cudaEventDestroy(ctx->ev_peer[k]);
    }
    for (int k = 0; k < 9; ++k) if (ctx->ev_push[k]) cudaEventDestroy(ctx->ev_push[k]);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    for (int k = 0; k < 3; ++k)
        if (ctx->ev_k1[k]) cudaEventDestroy(ctx->ev_k1[k]);
    cudaGetLastError();
    delete ctx;
    return GZB_OK;
}

extern "C" int gzb_set_stream(gzb_ctx* ctx, void* cuda_stream) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_set_stream: null ctx");
    ctx->stream = (cudaStream_t)cuda_stream;
    return GZB_OK;
}

extern "C" int gzb_use_own_stream(gzb_ctx* ctx) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_use_own_stream: null ctx");
    ctx->stream = ctx->own_stream;
    return GZB_OK;
}

extern "C" int gzb_sync(gzb_ctx* ctx) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_sync: null ctx");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    if (ctx->time_err_pending) {
        int* h_err = (int*)((char*)ctx->pinned + 64);
        GZB_CUDA(ctx, cudaMemcpyAsync(h_err, ctx->d_flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (ctx->trace && ctx->ntrace > 0) {
        cudaDeviceSynchronize();
        for (int k = 0; k < ctx->ntrace; ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->trace_ev[0], ctx->trace_ev[k]);
            fprintf(stderr, "[gzb trace] %8.1f us  %s\n", 1e3 * ms, ctx->trace_name[k]);
        }
        ctx->ntrace = 0;
    }
    if (ctx->barrier_err_pending)
        GZB_CUDA(ctx, cudaMemcpyAsync((char*)ctx->pinned + 72, ctx->d_flags + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
#ifdef GZB_CHECKED
    GZB_CUDA(ctx, cudaMemcpyAsync((char*)ctx->pinned + 80, ctx->d_flags + 4, 4, cudaMemcpyDeviceToHost, ctx->stream));
#endif
    GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
#ifdef GZB_KTIME
    if (ctx->d_kt) {
        static const char* nm[16] = {"k_near_search", "k_near_slow", "k_chain entry", "k_chain (after the dependency wait | summaries of the points read)",
                                     "k_chain (- | block maps published)", "k_chain (- | look-back done)", "k_chain (- | NP and block break lists written)",
                                     "k_chain (- | end, ordered break list)", "k_tail entry", "k_tail (after the dependency wait | replays done)",
                                     "k_tail (- | validation done)", "k_tail (- | write-back done)", "k_tail (- | end)", "bulk kernel", "k_path_fix", "-"};
        unsigned long long h[64];
        cudaDeviceSynchronize();
        if (cudaMemcpy(h, ctx->d_kt, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
            unsigned long long t0 = ~0ull;
            for (int k = 0; k < 32; ++k) if (h[k] < t0) t0 = h[k];
            if (t0 != ~0ull) {
                for (int k = 0; k < 15; ++k) {
                    if (h[k] == ~0ull && h[32 + k] == 0ull) continue;
                    fprintf(stderr, "[gzb ktime] first %8.2f us   last %8.2f us   %s\n", h[k] == ~0ull ? -1.0 : 1e-3 * (double)(h[k] - t0),
                            h[32 + k] == 0ull ? -1.0 : 1e-3 * (double)(h[32 + k] - t0), nm[k]);
                }
                cudaMemset(ctx->d_kt, 0xff, 32 * 8);
                cudaMemset(ctx->d_kt + 32, 0, 32 * 8);
            }
        } else cudaGetLastError();
    }
#endif
#ifdef GZB_CHECKED
    {
        int* h_c = (int*)((char*)ctx->pinned + 80);
        if (*h_c) {
            const int bits = *h_c;
            *h_c = 0;
            GZB_CUDA(ctx, cudaMemsetAsync(ctx->d_flags + 4, 0, 4, ctx->stream));
            return gzb_fail(ctx, GZB_ERR_STATE, "CHECKED build: a kernel computed an index outside its bound (bits 0x%x, see GZB_CHK)", bits);
        }
    }
#endif
    if (ctx->barrier_err_pending) {
        ctx->barrier_err_pending = 0;
        int* h_b = (int*)((char*)ctx->pinned + 72);
        if (*h_b) {
            *h_b = 0;
            GZB_CUDA(ctx, cudaMemsetAsync(ctx->d_flags + 3, 0, 4, ctx->stream));
            return gzb_fail(ctx, GZB_ERR_STATE, "gzb_peer_barrier: a peer did not arrive within the time limit");
        }
    }
    if (ctx->time_err_pending) {
        ctx->time_err_pending = 0;
        int* h_err = (int*)((char*)ctx->pinned + 64);
        if (*h_err) {
            *h_err = 0;
            GZB_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, 4, ctx->stream));
            return gzb_fail(ctx, GZB_ERR_TIME, "gzb_interp: track time outside the model records [tmod[0], tmod[nrec-1])");
        }
    }
    return GZB_OK;
}

extern "C" int gzb_get_stats(const gzb_ctx* ctx_c, gzb_stats* out) {
    gzb_ctx* ctx = const_cast<gzb_ctx*>(ctx_c);
    if (!ctx || !out) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_get_stats: null argument");
    if (ctx->nearest_stats_pending) {      // left in the persistent device flags by gzb_nearest
        if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
        GZB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->d_flags + 8, 56, cudaMemcpyDeviceToHost, ctx->stream));
        GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const long long* w = (const long long*)ctx->pinned;
        ctx->stats.last_breaks = w[0];
        ctx->stats.last_chain_steps = w[1];
        ctx->stats.last_unresolved = w[2];
        ctx->stats.last_unres_noseed = w[3];
        ctx->stats.last_unres_noconv = w[4];
        ctx->stats.last_unres_nocert = w[5];
        ctx->stats.last_max_chain = w[6];
        ctx->nearest_stats_pending = 0;
    }
    *out = ctx->stats;
    return GZB_OK;
}

extern "C" int gzb_set_option(gzb_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_set_option: null argument");
    if (!strcmp(name, "zero_copy")) { ctx->zero_copy = value < 0 ? 0 : (value > 2 ? 2 : (int)value); return GZB_OK; }
    if (!strcmp(name, "nearest_path")) { ctx->nearest_path = value ? 1 : 0; return GZB_OK; }
    if (!strcmp(name, "edge_exclusion")) { ctx->edge_exclusion = value ? 1 : 0; return GZB_OK; }
    if (!strcmp(name, "heading_table")) {
        ctx->hdg_mode = value < 0 ? 0 : (value > 2 ? 2 : (int)value);
        if (ctx->hdg_mode == 0) { ctx->g.hdg = nullptr; ctx->hdg_built = 0; ctx->stats.heading_table = 0; }
        return GZB_OK;
    }
    if (!strcmp(name, "trace")) { ctx->trace = value ? 1 : 0; ctx->ntrace = 0; return GZB_OK; }
    if (!strcmp(name, "twin_chain")) { ctx->twin_chain = value ? 1 : 0; return GZB_OK; }
    if (!strcmp(name, "fuse_k4")) { ctx->fuse_k4 = value ? 1 : 0; return GZB_OK; }
    if (!strcmp(name, "overlap")) { ctx->overlap = value ? 1 : 0; return GZB_OK; }
    if (!strcmp(name, "bulk_chunks")) { ctx->bulk_chunks = value < 0 ? 0 : (value > 8 ? 8 : (int)value); return GZB_OK; }
    if (!strcmp(name, "k4_early")) { ctx->k4_early = value ? 1 : 0; return GZB_OK; }
    if (!strcmp(name, "pdl")) { ctx->pdl = value ? 1 : 0; return GZB_OK; }
    if (!strcmp(name, "push_blocks")) { ctx->push_blocks = value < 1 ? 1 : (value > 4096 ? 4096 : (int)value); return GZB_OK; }
    if (!strcmp(name, "push_mode")) { ctx->push_mode = value ? 1 : 0; return GZB_OK; }
    if (!strcmp(name, "carveout")) { ctx->carveout = value < 0 ? -1 : (value > 100 ? 100 : (int)value); return GZB_OK; }
    if (!strcmp(name, "chain_first")) { ctx->chain_first = value < 0 ? 0 : (value > 50 ? 50 : (int)value); return GZB_OK; }
    if (!strcmp(name, "bulk_resident")) { ctx->bulk_resident = value < 0 ? 0 : (value > 16 ? 16 : (int)value); return GZB_OK; }
    if (!strcmp(name, "ix64")) { ctx->force_ix64 = value ? 1 : 0; return GZB_OK; }
    if (!strcmp(name, "prefetch")) { ctx->prefetch = value < 0 ? 0 : (value > 4 ? 4 : (int)value); return GZB_OK; }
    if (!strcmp(name, "l2_fetch")) {      // device-wide hint (cudaLimitMaxL2FetchGranularity): 32, 64 or 128 bytes
        if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
        size_t got = 0;
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value);
        if (e != cudaSuccess) { cudaGetLastError(); return gzb_fail(ctx, GZB_ERR_CUDA, "cudaDeviceSetLimit(L2 fetch granularity, %lld): %s", (long long)value, cudaGetErrorString(e)); }
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        ctx->stats.l2_fetch_bytes = (int32_t)got;
        return GZB_OK;
    }
    if (!strcmp(name, "mask_bits")) {
        if (!value) { ctx->mask_bits_ok = 0; return GZB_OK; }
        return gzb_build_mask_bits(ctx);
    }
    if (!strcmp(name, "force_heavy")) { ctx->force_heavy = value < 0 ? 0 : (value > 2 ? 2 : (int)value); return GZB_OK; }
    return gzb_fail(ctx, GZB_ERR_ARG, "gzb_set_option: unknown option '%s'", name);
}

extern "C" int gzb_set_ew_per(gzb_ctx* ctx, int k_ew_per) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_set_ew_per: null ctx");
    if (ctx->g.kew != k_ew_per) {      // the neighbours of the edge columns change: drop the K2 table
        ctx->g.hdg = nullptr;
        ctx->hdg_built = 0;
        ctx->stats.heading_table = 0;
    }
    ctx->g.kew = k_ew_per;
    return GZB_OK;
}

extern "C" int gzb_set_mask(gzb_ctx* ctx, const int8_t* mask, int where) {
    if (!ctx || !mask) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_set_mask: null argument");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    const size_t n = (size_t)ctx->g.Nj * (size_t)ctx->g.Ni;
    if (!ctx->d_mask) GZB_CUDA(ctx, gzb_pool_malloc((void**)&ctx->d_mask, n));
    GZB_CUDA(ctx, cudaMemcpyAsync(ctx->d_mask, mask, n,
                                  where == GZB_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->g.mask = ctx->d_mask;
    const int rc = gzb_build_mask_bits(ctx);
    if (rc != GZB_OK) return rc;
    if (where == GZB_HOST) {
        ctx->stats.h2d_bytes += n;
        GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return GZB_OK;
}

extern "C" int gzb_set_angle(gzb_ctx* ctx, const double* angle_or_null, int where) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_set_angle: null ctx");
    if (!angle_or_null) {
        ctx->g.angle = nullptr;
        return GZB_OK;
    }
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    const size_t n = (size_t)ctx->g.Nj * (size_t)ctx->g.Ni * sizeof(double);
    if (!ctx->d_angle) GZB_CUDA(ctx, gzb_pool_malloc((void**)&ctx->d_angle, n));
    GZB_CUDA(ctx, cudaMemcpyAsync(ctx->d_angle, angle_or_null, n,
                                  where == GZB_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
    if (where == GZB_HOST) {
        ctx->stats.h2d_bytes += n;
        GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->g.angle = ctx->d_angle;
    return GZB_OK;
}


// ======================================================================== longitude statistics
// What IsGlobalLongitudeWise needs (gonzag/utils.py:149-182): with X = lon % 360 and
// XB = degE_to_degWE(X): min/max of X, the columns of the FIRST occurrences of that min and max
// (numpy argmin / argmax), min/max of XB.  Two passes: extrema, then the smallest flat index
// attaining them.  w[0..3] = ord_enc(min X, max X, min XB, max XB); w[4], w[5] = flat indices.
__global__ void __launch_bounds__(256)
k_lon_minmax(const int64_t n, const double* __restrict__ lon, unsigned long long* __restrict__ w) {
    double v[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
    bool nan = false;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double x = gzb_pymod(lon[k], 360.0), xb = gzb_pm180(x);
        if (x != x) nan = true;
        v[0] = fmin(v[0], x); v[1] = fmax(v[1], x);
        v[2] = fmin(v[2], xb); v[3] = fmax(v[3], xb);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v[q], o);
            v[q] = (q & 1) ? fmax(v[q], u) : fmin(v[q], u);
        }
        if ((threadIdx.x & 31) == 0 && isfinite(v[q])) {
            if (q & 1) atomicMax(w + q, ord_enc(v[q])); else atomicMin(w + q, ord_enc(v[q]));
        }
    }
    if (nan) atomicOr((unsigned long long*)(w + 6), 1ull);
}

__global__ void __launch_bounds__(256)
k_lon_argminmax(const int64_t n, const double* __restrict__ lon, unsigned long long* __restrict__ w) {
    const unsigned long long emin = w[0], emax = w[1];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long e = ord_enc(gzb_pymod(lon[k], 360.0));
        if (e == emin) atomicMin(w + 4, (unsigned long long)k);
        if (e == emax) atomicMin(w + 5, (unsigned long long)k);
    }
}

extern "C" int gzb_lon_stats(gzb_ctx* ctx, double* vals4, int64_t* cols2) {
    if (!ctx || !vals4 || !cols2) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_lon_stats: null argument");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    if (gzb_arena_reserve(ctx, 64) != GZB_OK) return GZB_ERR_NOMEM;
    unsigned long long* w = (unsigned long long*)gzb_arena_take(ctx, 64);
    unsigned long long* h = (unsigned long long*)((char*)ctx->pinned + 512);
    h[0] = ~0ull; h[1] = 0ull; h[2] = ~0ull; h[3] = 0ull; h[4] = ~0ull; h[5] = ~0ull; h[6] = 0ull; h[7] = 0ull;
    cudaStream_t s = ctx->stream;
    GZB_CUDA(ctx, cudaMemcpyAsync(w, h, 64, cudaMemcpyHostToDevice, s));
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    const int64_t n = ctx->g.Nj * ctx->g.Ni;
    k_lon_minmax<<<nsm * 8, 256, 0, s>>>(n, ctx->d_lon, w);
    GZB_LAUNCH_CHECK(ctx);
    k_lon_argminmax<<<nsm * 8, 256, 0, s>>>(n, ctx->d_lon, w);
    GZB_LAUNCH_CHECK(ctx);
    GZB_CUDA(ctx, cudaMemcpyAsync(h, w, 64, cudaMemcpyDeviceToHost, s));
    GZB_CUDA(ctx, cudaStreamSynchronize(s));
    if (h[6]) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_lon_stats: NaN longitude");
    for (int q = 0; q < 4; ++q) vals4[q] = ord_dec(h[q]);
    cols2[0] = (int64_t)(h[4] % (unsigned long long)ctx->g.Ni);
    cols2[1] = (int64_t)(h[5] % (unsigned long long)ctx->g.Ni);
    return GZB_OK;
}

// lon -> degE_to_degWE(lon) in place (ModGrid does it when the domain is handled in the
// [-180,180] frame, gonzag/utils.py:441).  The fp32 search structures describe the same points on
// the sphere and stay; the K2 table, which stores longitudes, is dropped.
__global__ void __launch_bounds__(256)
k_lon_pm180(const int64_t n, double* __restrict__ lon, double2* __restrict__ ll) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double x = gzb_pm180(lon[k]);
    lon[k] = x;
    ll[k].y = x;
}

extern "C" int gzb_lon_to_pm180(gzb_ctx* ctx, double* lon_out_or_null) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_lon_to_pm180: null ctx");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    const int64_t n = ctx->g.Nj * ctx->g.Ni;
    k_lon_pm180<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, ctx->d_lon, ctx->d_ll);
    GZB_LAUNCH_CHECK(ctx);
    ctx->g.hdg = nullptr;
    ctx->hdg_built = 0;
    ctx->stats.heading_table = 0;
    if (lon_out_or_null) {
        GZB_CUDA(ctx, cudaMemcpyAsync(lon_out_or_null, ctx->d_lon, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += (uint64_t)n * sizeof(double);
        GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return GZB_OK;
}

// extent of the latitudes (ModGrid.domain_bounds): computed when the context was built
extern "C" int gzb_lat_range(gzb_ctx* ctx, double* lat_min, double* lat_max) {
    if (!ctx || !lat_min || !lat_max) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_lat_range: null argument");
    *lat_min = ctx->lat_min;
    *lat_max = ctx->lat_max;
    return GZB_OK;
}

// ======================================================================== peer (NVLink) outputs
// One process per GPU: a rank exports its product buffer with CUDA IPC, the other ranks map it and
// hand the mapped addresses to gzb_set_peer_outputs; K4 then stores its two series into every
// peer's buffer as it computes them (gzb_point.cuh), which replaces the all-gather that would
// otherwise follow it.
extern "C" int gzb_ipc_export(gzb_ctx* ctx, void* dptr, unsigned char* handle64) {
    if (!ctx || !dptr || !handle64) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_ipc_export: null argument");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    GZB_CUDA(ctx, cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle64, &h, 64);
    return GZB_OK;
}

extern "C" int gzb_ipc_open(gzb_ctx* ctx, const unsigned char* handle64, void** dptr_out) {
    if (!ctx || !handle64 || !dptr_out) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_ipc_open: null argument");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    *dptr_out = nullptr;
    GZB_CUDA(ctx, cudaIpcOpenMemHandle(dptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return GZB_OK;
}

extern "C" int gzb_ipc_close(gzb_ctx* ctx, void* dptr) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_ipc_close: null ctx");
    if (!dptr) return GZB_OK;
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GZB_CUDA(ctx, cudaIpcCloseMemHandle(dptr));
    return GZB_OK;
}

extern "C" int gzb_set_peer_outputs(gzb_ctx* ctx, int n, void* const* vnp_peers, void* const* vbl_peers,
                                    void* vnp_local, void* vbl_local) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_set_peer_outputs: null ctx");
    if (n < 0 || n > GZB_MAX_PEERS) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_set_peer_outputs: 0 <= n <= %d", GZB_MAX_PEERS);
    if (n > 0 && (!vnp_peers || !vbl_peers || !vnp_local || !vbl_local))
        return gzb_fail(ctx, GZB_ERR_ARG, "gzb_set_peer_outputs: null pointer");
    ctx->peer_out.n = n;
    for (int k = 0; k < n; ++k) {
        ctx->peer_out.np[k] = (double*)vnp_peers[k];
        ctx->peer_out.bl[k] = (double*)vbl_peers[k];
    }
    ctx->peer_local_np = (double*)vnp_local;
    ctx->peer_local_bl = (double*)vbl_local;
    return GZB_OK;
}

// first and last row of an (nrows, 2) int64 array (the halo assumption and the last nearest point of
// a segment) to n destinations of 4 int64 each: what the ranks compare at the segment boundaries
struct GzbEdgeDst { long long* p[GZB_MAX_PEERS + 1]; };
__global__ void k_publish_edges(const long long* __restrict__ rows, const long long nrows, const int n, const GzbEdgeDst d) {
    const int k = threadIdx.x >> 2, w = threadIdx.x & 3;
    if (k >= n) return;
    const long long v = (w < 2) ? rows[w] : rows[2 * (nrows - 1) + (w - 2)];
    d.p[k][w] = v;
}

// Cross-GPU barrier over peer memory, stream-ordered: every rank owns bar[world] epoch counters;
// rank r stores `epoch` into bar[r] of every peer (P2P store after a system-wide fence, so that its
// earlier stores into the peers -- K4's results -- are visible first) and waits until its own
// bar[g] has reached `epoch` for every g.  Bounded wait: ~4 s of SM clocks, then *err |= 2.
struct GzbBarDst { long long* p[GZB_MAX_PEERS + 1]; };
__global__ void k_peer_barrier(const GzbBarDst d, const int n, volatile long long* local_bar, const int world,
                               const int my_rank, const long long epoch, int* __restrict__ err) {
    const int t = threadIdx.x;
    __threadfence_system();
    if (t < n) *(volatile long long*)d.p[t] = epoch;
    __threadfence_system();
    if (t < world && t != my_rank) {
        const long long start = clock64();
        while (local_bar[t] < epoch) {
            if (clock64() - start > 8000000000LL) { atomicOr(err, 2); break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

extern "C" int gzb_peer_barrier(gzb_ctx* ctx, int n, void* const* peer_slots, void* local_bar, int world, int my_rank,
                                int64_t epoch) {
    if (!ctx || n < 0 || n > GZB_MAX_PEERS || world < 1 || world > GZB_MAX_PEERS + 1 || !local_bar || (n > 0 && !peer_slots))
        return gzb_fail(ctx, GZB_ERR_ARG, "gzb_peer_barrier: bad arguments");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    GzbBarDst d;
    for (int k = 0; k < n; ++k) d.p[k] = (long long*)peer_slots[k];
    k_peer_barrier<<<1, 32, 0, ctx->stream>>>(d, n, (volatile long long*)local_bar, world, my_rank, (long long)epoch,
                                             ctx->d_flags + 3);
    GZB_LAUNCH_CHECK(ctx);
    ctx->barrier_err_pending = 1;
    return GZB_OK;
}

// The end of a sharded step in ONE kernel (instead of edge publication + barrier + boundary check as three launches, one
// of them a collective): (1) this rank's 4 edge words -- first and last row of its (nrows, 2) nearest-point array: the halo
// assumption and the segment's last nearest point -- go to its own gathered buffer and to every peer's; (2) barrier over peer
// memory as k_peer_barrier: after it every rank's series (stored by K4 / the push kernels over NVLink) and edge words are
// visible in this rank's buffer; (3) *flag |= 1 if some segment's halo assumption differs from the last nearest point of
// the segment before it.
__global__ void k_peer_finish(const long long* __restrict__ rows, const long long nrows, const int nedge, const GzbEdgeDst ed,
                              const GzbBarDst bd, const int nbar, volatile long long* local_bar, const int world,
                              const int my_rank, const long long epoch, const long long* edges, const long long stride,
                              const long long offset, int* __restrict__ flag, int* __restrict__ err) {
    const int t = threadIdx.x;
    {
        const int k = t >> 2, w = t & 3;
        if (k < nedge) ed.p[k][w] = (w < 2) ? rows[w] : rows[2 * (nrows - 1) + (w - 2)];
    }
    __threadfence_system();
    __syncthreads();
    if (t < nbar) *(volatile long long*)bd.p[t] = epoch;
    __threadfence_system();
    if (t < world && t != my_rank) {
        const long long start = clock64();
        while (local_bar[t] < epoch) {
            if (clock64() - start > 8000000000LL) { atomicOr(err, 2); break; }
            __nanosleep(100);
        }
    }
    __threadfence_system();
    __syncthreads();
    const int g = t + 1;
    if (g < world && flag) {
        const volatile long long* cur = edges + (long long)g * stride + offset;
        const volatile long long* prv = edges + (long long)(g - 1) * stride + offset;
        if (cur[0] != prv[2] || cur[1] != prv[3]) atomicOr(flag, 1);
    }
}

extern "C" int gzb_peer_finish(gzb_ctx* ctx, const int64_t* np_rows, int64_t nrows, int nedge, void* const* edge_dst, int nbar,
                               void* const* peer_slots, void* local_bar, int world, int my_rank, int64_t epoch,
                               const int64_t* gathered, int64_t stride_words, int64_t offset_words, int* flag_dev_or_null) {
    if (!ctx || !np_rows || nrows < 1 || nedge < 0 || nedge > GZB_MAX_PEERS + 1 || (nedge > 0 && !edge_dst) || nbar < 0 ||
        nbar > GZB_MAX_PEERS || world < 1 || world > GZB_MAX_PEERS + 1 || !local_bar || (nbar > 0 && !peer_slots) || !gathered)
        return gzb_fail(ctx, GZB_ERR_ARG, "gzb_peer_finish: bad arguments");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    GzbEdgeDst ed;
    GzbBarDst bd;
    for (int k = 0; k < nedge; ++k) ed.p[k] = (long long*)edge_dst[k];
    for (int k = 0; k < nbar; ++k) bd.p[k] = (long long*)peer_slots[k];
    k_peer_finish<<<1, 4 * (GZB_MAX_PEERS + 1), 0, ctx->stream>>>((const long long*)np_rows, nrows, nedge, ed, bd, nbar,
                                                                  (volatile long long*)local_bar, world, my_rank, (long long)epoch,
                                                                  (const long long*)gathered, stride_words, offset_words,
                                                                  flag_dev_or_null, ctx->d_flags + 3);
    GZB_LAUNCH_CHECK(ctx);
    gzb_trace(ctx, ctx->stream, "end-of-step kernel end: edges, barrier, check (main)");
    ctx->barrier_err_pending = 1;
    return GZB_OK;
}

// boundary check of the gathered edge words: flag |= any(halo assumption of rank g != last nearest
// point of rank g-1).  edges: (world, stride) int64 with the 4 edge words at `offset` of every row.
__global__ void k_check_edges(const long long* __restrict__ edges, const long long stride, const long long offset,
                              const int world, int* __restrict__ flag) {
    const int g = threadIdx.x + 1;
    if (g >= world) return;
    const long long* cur = edges + (long long)g * stride + offset;
    const long long* prv = edges + (long long)(g - 1) * stride + offset;
    if (cur[0] != prv[2] || cur[1] != prv[3]) atomicOr(flag, 1);
}

extern "C" int gzb_check_edges(gzb_ctx* ctx, const int64_t* gathered, int64_t stride_words, int64_t offset_words,
                               int world, int* flag_dev) {
    if (!ctx || !gathered || !flag_dev || world < 1 || world > 1024)
        return gzb_fail(ctx, GZB_ERR_ARG, "gzb_check_edges: bad arguments");
    if (world == 1) return GZB_OK;
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    k_check_edges<<<1, ((world + 31) / 32) * 32, 0, ctx->stream>>>((const long long*)gathered, stride_words, offset_words, world,
                                                                  flag_dev);
    GZB_LAUNCH_CHECK(ctx);
    return GZB_OK;
}

extern "C" int gzb_publish_edges(gzb_ctx* ctx, const int64_t* np_rows, int64_t nrows, int n, void* const* dst) {
    if (!ctx || !np_rows || nrows < 1 || n < 0 || n > GZB_MAX_PEERS + 1 || (n > 0 && !dst))
        return gzb_fail(ctx, GZB_ERR_ARG, "gzb_publish_edges: bad arguments");
    if (n == 0) return GZB_OK;
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    GzbEdgeDst d;
    for (int k = 0; k < n; ++k) d.p[k] = (long long*)dst[k];
    k_publish_edges<<<1, 4 * (GZB_MAX_PEERS + 1), 0, ctx->stream>>>((const long long*)np_rows, nrows, n, d);
    GZB_LAUNCH_CHECK(ctx);
    return GZB_OK;
}

// ======================================================================== K0: grid angle
// gonzag/utils.py:226-262.  One thread walks a column strip of GA_ROWS rows with a 3-row
// sliding window, so each lat/lon value is loaded once per strip (plus a 2-row halo) and its
// tan / cos / sin are evaluated once and reused as the j-1, j and j+1 operand.
#define GA_ROWS 32

struct GaRow {
    double t, c, s, lm;   // tan(pi/4 - lat/2), cos(lon), sin(lon), lon % 360
};

__device__ __forceinline__ GaRow ga_load(const double* __restrict__ lat, const double* __restrict__ lon, int64_t k) {
    GaRow r;
    const double la = lat[k], lo = lon[k];
    r.t = tan(GZB_PI / 4.0 - GZB_D2R * la / 2.0);
    r.c = cos(GZB_D2R * lo);
    r.s = sin(GZB_D2R * lo);
    r.lm = gzb_pymod(lo, 360.0);
    return r;
}

__global__ void __launch_bounds__(128)
k_grid_angle(int64_t Nj, int64_t Ni, const double* __restrict__ lat, const double* __restrict__ lon,
             double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ni) return;
    const int64_t j0 = (int64_t)blockIdx.y * GA_ROWS;
    int64_t ja = j0 < 1 ? 1 : j0;
    int64_t jb = j0 + GA_ROWS;
    if (jb > Nj - 1) jb = Nj - 1;
    if (j0 == 0) out[i] = 0.0;
    if (j0 + GA_ROWS >= Nj) out[(Nj - 1) * Ni + i] = 0.0;
    if (ja >= jb) return;
    GaRow dn = ga_load(lat, lon, (ja - 1) * Ni + i);
    GaRow md = ga_load(lat, lon, ja * Ni + i);
    for (int64_t j = ja; j < jb; ++j) {
        const GaRow up = ga_load(lat, lon, (j + 1) * Ni + i);
        double ang = 0.0;
        if (!(fabs(up.lm - dn.lm) < 1.0e-8)) {
            const double px = 0.0 - 2.0 * md.c * md.t;
            const double py = 0.0 - 2.0 * md.s * md.t;
            const double pn = px * px + py * py;
            const double vx = 2.0 * (up.c * up.t - dn.c * dn.t);
            const double vy = 2.0 * (up.s * up.t - dn.s * dn.t);
            double nv = sqrt(pn * (vx * vx + vy * vy));
            nv = fmax(nv, 1.0e-12);
            const double sn = (px * vy - py * vx) / nv;
            const double cs = (px * vx + py * vy) / nv;
            ang = atan2(sn, cs) * 180.0 / GZB_PI;
        }
        out[j * Ni + i] = ang;
        dn = md;
        md = up;
    }
}

extern "C" int gzb_grid_angle(gzb_ctx* ctx, double* angle_out_or_null, int where) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_grid_angle: null ctx");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    const int64_t Nj = ctx->g.Nj, Ni = ctx->g.Ni;
    const size_t bytes = (size_t)Nj * (size_t)Ni * sizeof(double);
    if (!ctx->d_angle) GZB_CUDA(ctx, gzb_pool_malloc((void**)&ctx->d_angle, bytes));
    dim3 grid((unsigned)((Ni + 127) / 128), (unsigned)((Nj + GA_ROWS - 1) / GA_ROWS));
    k_grid_angle<<<grid, 128, 0, ctx->stream>>>(Nj, Ni, ctx->d_lat, ctx->d_lon, ctx->d_angle);
    GZB_LAUNCH_CHECK(ctx);
    ctx->g.angle = ctx->d_angle;
    if (angle_out_or_null) {
        if (where == GZB_HOST) {
            const int rc = gzb_bigcopy(ctx, angle_out_or_null, ctx->d_angle, bytes, false);
            if (rc != GZB_OK) return rc;
            ctx->stats.d2h_bytes += bytes;
        } else {
            GZB_CUDA(ctx, cudaMemcpyAsync(angle_out_or_null, ctx->d_angle, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    if (where == GZB_HOST) GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GZB_OK;
}

// Copy of the angle installed in the context (by gzb_create, gzb_set_angle or gzb_grid_angle): no kernel runs, so a
// host that reads `ModGrid.xangle` lazily sees exactly what K2 uses (gonzag/utils.py:431 computes it once).
extern "C" int gzb_get_angle(gzb_ctx* ctx, double* angle_out, int where) {
    if (!ctx || !angle_out) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_get_angle: null argument");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    if (!ctx->g.angle) return gzb_fail(ctx, GZB_ERR_STATE, "gzb_get_angle: the context holds no angle (no -D)");
    const size_t bytes = (size_t)ctx->g.Nj * (size_t)ctx->g.Ni * sizeof(double);
    if (where == GZB_HOST) {
        const int rc = gzb_bigcopy(ctx, angle_out, ctx->g.angle, bytes, false);
        if (rc != GZB_OK) return rc;
        ctx->stats.d2h_bytes += bytes;
        GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        GZB_CUDA(ctx, cudaMemcpyAsync(angle_out, ctx->g.angle, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return GZB_OK;
}

// Device pointers of the arrays a context was created from (as uploaded: lat, lon row-major float64; the angle
// and the int8 mask when installed, else NULL).  A multi-GPU host replicates a grid with them: one rank uploads,
// the others receive the arrays over NVLink (any collective) and call gzb_create with DEVICE pointers.
extern "C" int gzb_grid_arrays(gzb_ctx* ctx, void** lat, void** lon, void** angle, void** mask) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_grid_arrays: null ctx");
    if (lat) *lat = ctx->d_lat;
    if (lon) *lon = ctx->d_lon;
    if (angle) *angle = (void*)ctx->g.angle;
    if (mask) *mask = (void*)ctx->g.mask;
    return GZB_OK;
}

// ======================================================================== memory helpers
extern "C" int gzb_dev_alloc(gzb_ctx* ctx, uint64_t bytes, void** dptr_out) {
    if (!ctx || !dptr_out) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_dev_alloc: null argument");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    *dptr_out = nullptr;
    cudaError_t e = gzb_pool_malloc(dptr_out, bytes ? (size_t)bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return gzb_fail(ctx, GZB_ERR_NOMEM, "gzb_pool_malloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
    }
    return GZB_OK;
}

extern "C" int gzb_dev_free(gzb_ctx* ctx, void* dptr) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_dev_free: null ctx");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    // the block may be handed out again at once: nothing queued on this context may still use it
    GZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GZB_CUDA(ctx, gzb_pool_free(dptr));
    return GZB_OK;
}

extern "C" int gzb_h2d(gzb_ctx* ctx, void* dst_dev, const void* src_host, uint64_t bytes) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_h2d: null ctx");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    // large: parallel staging when the host buffer is pageable (returns when done).  Medium (a track array of a few MB): the
    // same when the stream is idle -- a plain copy from pageable memory runs at 11 GB/s, and nothing is queued that the
    // staged copy would have to wait for
    bool idle = false;
    if (bytes >= ((size_t)3 << 20) && bytes < GZB_BC_MIN) {
        idle = cudaStreamQuery(ctx->stream) == cudaSuccess;
        if (!idle) cudaGetLastError();
    }
    if (bytes >= GZB_BC_MIN || idle) {
        const int rc = gzb_bigcopy(ctx, dst_dev, src_host, (size_t)bytes, true, true, 0, (size_t)3 << 20);
        if (rc != GZB_OK) return rc;
    } else {
        GZB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->stats.h2d_bytes += bytes;
    return GZB_OK;
}

extern "C" int gzb_d2h(gzb_ctx* ctx, void* dst_host, const void* src_dev, uint64_t bytes) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_d2h: null ctx");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    if (bytes >= GZB_BC_MIN) {
        const int rc = gzb_bigcopy(ctx, dst_host, src_dev, (size_t)bytes, false);
        if (rc != GZB_OK) return rc;
    } else {
        GZB_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ctx->stats.d2h_bytes += bytes;
    return GZB_OK;
}

extern "C" int gzb_d2d(gzb_ctx* ctx, void* dst_dev, const void* src_dev, uint64_t bytes) {
    if (!ctx || (bytes && (!dst_dev || !src_dev))) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_d2d: null argument");
    if (!bytes) return GZB_OK;
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    GZB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_dev, (size_t)bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return GZB_OK;
}

extern "C" int gzb_memset(gzb_ctx* ctx, void* dst_dev, int byte, uint64_t bytes) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_memset: null ctx");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    GZB_CUDA(ctx, cudaMemsetAsync(dst_dev, byte, (size_t)bytes, ctx->stream));
    return GZB_OK;
}

// 0: pageable host memory (or unknown), 1: pinned + mapped host memory, 2: device / managed memory
extern "C" int gzb_pointer_kind(const void* p, int* kind_out) {
    if (!kind_out) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_pointer_kind: null output");
    *kind_out = 0;
    if (!p) return GZB_OK;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return GZB_OK; }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) *kind_out = 2;
    else if (at.type == cudaMemoryTypeHost && at.devicePointer) *kind_out = 1;
    return GZB_OK;
}

extern "C" int gzb_host_alloc(uint64_t bytes, void** hptr_out) {
    if (!hptr_out) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_host_alloc: null output");
    *hptr_out = nullptr;
    cudaError_t e = cudaHostAlloc(hptr_out, bytes ? (size_t)bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return gzb_fail(nullptr, GZB_ERR_NOMEM, "cudaHostAlloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
    }
    return GZB_OK;
}

extern "C" int gzb_host_free(void* hptr) {
    if (!hptr) return GZB_OK;
    cudaError_t e = cudaFreeHost(hptr);
    if (e != cudaSuccess) return gzb_fail(nullptr, GZB_ERR_CUDA, "cudaFreeHost: %s", cudaGetErrorString(e));
    return GZB_OK;
}

extern "C" int gzb_host_register(void* hptr, uint64_t bytes) {
    cudaError_t e = cudaHostRegister(hptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return gzb_fail(nullptr, GZB_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
    }
    return GZB_OK;
}

extern "C" int gzb_host_unregister(void* hptr) {
    cudaError_t e = cudaHostUnregister(hptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return gzb_fail(nullptr, GZB_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
    }
    return GZB_OK;
}

// ======================================================================== host-side mask packing
// The reference keeps its land-sea mask as int64 {0,1} (gonzag/ncio.py:164-191, `MG.mask`); the context wants
// int8.  NumPy's astype over an ORCA12 mask (13.2 M cells, 106 MB read) is one thread at ~7.5 GB/s = 14 ms, a
// third of an end-to-end call: here the conversion is cut over the host threads.  dst[k] = (int8_t)src[k]
// (the low byte, what astype(int8) gives for integer input).  Pure host code: no device, no context.
namespace {
template <typename T> void pack_part(const T* src, int8_t* dst, size_t a, size_t b) {
    for (size_t k = a; k < b; ++k) dst[k] = (int8_t)src[k];
}
}  // namespace

extern "C" int gzb_pack_mask(const void* src, int itemsize, int64_t n, int8_t* dst) {
    if (n < 0 || (n > 0 && (!src || !dst))) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_pack_mask: bad arguments");
    if (itemsize != 1 && itemsize != 2 && itemsize != 4 && itemsize != 8)
        return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_pack_mask: itemsize must be 1, 2, 4 or 8 (got %d)", itemsize);
    if (n == 0) return GZB_OK;
    const int want = gzb_host_threads(16);
    const size_t per_min = (size_t)1 << 20;                        // below ~1 M elements a thread is not worth its start
    const size_t N = (size_t)n;
    int nth = (int)((N + per_min - 1) / per_min);
    nth = nth > want ? want : nth;
    const size_t part = (((N + nth - 1) / nth) + 63) & ~(size_t)63;
    auto run = [&](size_t a, size_t b) {
        switch (itemsize) {
            case 1: pack_part((const int8_t*)src, dst, a, b); break;
            case 2: pack_part((const int16_t*)src, dst, a, b); break;
            case 4: pack_part((const int32_t*)src, dst, a, b); break;
            default: pack_part((const int64_t*)src, dst, a, b); break;
        }
    };
    if (nth <= 1) { run(0, N); return GZB_OK; }
    // a thread that cannot be started (resource limits) leaves its part, and the rest, to the caller's thread
    std::vector<std::thread> th;
    size_t done_to = 0;
    try {
        th.reserve(nth);
        for (int t = 0; t < nth; ++t) {
            const size_t a = (size_t)t * part;
            if (a >= N) break;
            const size_t b = a + part < N ? a + part : N;
            th.emplace_back(run, a, b);
            done_to = b;
        }
    } catch (...) {
    }
    if (done_to < N) run(done_to, N);
    for (auto& t : th) t.join();
    return GZB_OK;
}

// ======================================================================== host-side record staging
// Model records read from a NetCDF-3 file are big-endian views of the memory map; NumPy's byte-swapping
// copy into the pinned staging slab is one thread (~25 ms per ORCA12 record, against a gather kernel that
// needs 40 us for 10 days of track).  dst[k] = src[k] with the bytes of every item reversed when `swap`
// is non-zero (plain copy otherwise), cut over the host threads.  itemsize 1, 2, 4 or 8; src and dst must
// not overlap.  Pure host code: no device, no context.
namespace {
template <typename T> inline T bswap_item(T v);
template <> inline uint16_t bswap_item(uint16_t v) { return __builtin_bswap16(v); }
template <> inline uint32_t bswap_item(uint32_t v) { return __builtin_bswap32(v); }
template <> inline uint64_t bswap_item(uint64_t v) { return __builtin_bswap64(v); }
template <typename T> void swap_part(const char* src, char* dst, size_t a, size_t b) {
    // memcpy in and out: the mapped file data need not be aligned to the item size
    for (size_t k = a; k < b; ++k) {
        T v;
        memcpy(&v, src + k * sizeof(T), sizeof(T));
        v = bswap_item<T>(v);
        memcpy(dst + k * sizeof(T), &v, sizeof(T));
    }
}
}  // namespace

extern "C" int gzb_copy_swap(void* dst, const void* src, int itemsize, int64_t n, int swap) {
    if (n < 0 || (n > 0 && (!src || !dst))) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_copy_swap: bad arguments");
    if (itemsize != 1 && itemsize != 2 && itemsize != 4 && itemsize != 8)
        return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_copy_swap: itemsize must be 1, 2, 4 or 8 (got %d)", itemsize);
    if (n == 0) return GZB_OK;
    const size_t N = (size_t)n;
    const bool do_swap = swap != 0 && itemsize > 1;
    const int want = gzb_host_threads(16);
    const size_t per_min = ((size_t)2 << 20) / (size_t)itemsize;       // 2 MiB per thread at least
    int nth = (int)((N + per_min - 1) / per_min);
    nth = nth > want ? want : nth;
    const size_t part = (((N + nth - 1) / nth) + 63) & ~(size_t)63;
    auto run = [&](size_t a, size_t b) {
        const char* s = (const char*)src;
        char* d = (char*)dst;
        if (!do_swap) { memcpy(d + a * (size_t)itemsize, s + a * (size_t)itemsize, (b - a) * (size_t)itemsize); return; }
        switch (itemsize) {
            case 2: swap_part<uint16_t>(s, d, a, b); break;
            case 4: swap_part<uint32_t>(s, d, a, b); break;
            default: swap_part<uint64_t>(s, d, a, b); break;
        }
    };
    if (nth <= 1) { run(0, N); return GZB_OK; }
    // a thread that cannot be started (resource limits) leaves its part, and the rest, to the caller's thread
    std::vector<std::thread> th;
    size_t done_to = 0;
    try {
        th.reserve(nth);
        for (int t = 0; t < nth; ++t) {
            const size_t a = (size_t)t * part;
            if (a >= N) break;
            const size_t b = a + part < N ? a + part : N;
            th.emplace_back(run, a, b);
            done_to = b;
        }
    } catch (...) {
    }
    if (done_to < N) run(done_to, N);
    for (auto& t : th) t.join();
    return GZB_OK;
}
