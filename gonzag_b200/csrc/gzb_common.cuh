// Shared declarations of libgonzag_b200: context layout, error plumbing and the device
// math helpers that restate the reference's scalar formulas.  The whole library is compiled
// with -fmad=false so that every a*b+c below rounds twice, like NumPy / CPython do.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gonzag_b200.h"

#define GZB_PI 3.141592653589793      // == math.pi
#define GZB_D2R (GZB_PI / 180.0)      // CPython's radians() multiplies by this constant
#define GZB_MAXLEV 12
#define GZB_TILE_J 4
#define GZB_TILE_I 8

// ------------------------------------------------------------------ bounding-sphere pyramid
// Level 0 nodes are leaf tiles of GZB_TILE_J x GZB_TILE_I cells; a level-l node (l >= 1) covers
// bj x bi nodes of level l-1.  Nodes of a level are stored "parent-blocked": the (up to) 32
// children of a parent are contiguous (one warp loads them with one coalesced request), at
// index parent_rowmajor_id*32 + slot.  The top level has <= 32 nodes and is stored row-major.
// The pyramid is fp32: it only PROPOSES candidate cells (with explicit error margins, see
// gzb_nearest.cu); every candidate is then compared with the reference's fp64 haversine.
struct GzbLevel {
    float4* nodes;    // (cx, cy, cz, chord radius, rounded up); radius < 0 marks an empty slot
    int nJ, nI;       // logical node counts
    int bj, bi;       // branching of THIS level's nodes over their children (powers of two)
    int sbj, sbi;     // log2(bj), log2(bi)
    int sJ, sI;       // cells covered by one node in j and i
};

// Lat-lon binned look-up table: bin -> flat index of some grid cell lying in (or next to) the
// bin; it only SEEDS the per-thread hill climb of k_near_fast.  bins == nullptr: no table.
struct GzbLut {
    const unsigned* bins;   // 0xffffffff = no cell anywhere near this bin
    int nlat, nlon;
    int wrap;               // the table covers the whole circle of longitudes
    double lat0, lon0;      // south-west corner (degrees); longitudes are taken modulo 360 from lon0
    double inv_d;           // 1 / bin size (degrees)
    double lon_span;        // covered longitude span (degrees), 360 when wrap
};

struct GzbGrid {
    int64_t Nj, Ni;
    int kew;
    const double* lat;      // row-major (Nj, Ni), as uploaded: the streaming kernels read these
    const double* lon;
    const double2* ll;      // the same coordinates interleaved (lat, lon): one sector per cell for the gathers
    const double* angle;    // may be null
    const int8_t* mask;     // may be null
    const float* cells;     // per leaf tile: x[32] y[32] z[32] (unit vectors), tile row-major
    const float* cert;      // per leaf tile (same index as its node): chord distance from the tile
                            // centre to the nearest cell outside the tile's 5x3-tile neighbourhood
    const float4* cellv;    // per cell, row-major: unit vector (x,y,z) + w = lower bound of the chord
                            // from this cell to any cell outside its 5x5 index neighbourhood
    const double* hdg;      // optional per-cell table of K2 (6 doubles per cell): Mercator ordinate and
                            // longitude (radians, [0,2pi)) of the cell, headings to its N, E, S, W neighbours
    unsigned long long* kt; // KTIME build (-DGZB_KTIME=1): 64 device words of in-kernel time marks (see GZB_KT_MIN); else unused
    int* chk;               // CHECKED build (-DGZB_CHECKED=1): device word that collects violated bounds (see GZB_CHK); else unused
    GzbLut lut;
    int nlev;
    GzbLevel lev[GZB_MAXLEV];
};

// the plain global answers of K1 (before the chain logic): flat index + distance of the nearest cell of
// every point, and what turns them into E(t) (threshold of the last useful pass, edge exclusion)
struct GzbGlobalNearest {
    const long long* G;     // 0x7ffffffffffffffe = not settled yet (the bulk kernels skip the point)
    const double* dG;
    double thr2;
    int edge;
    const long long* ulist;               // the points that were unsettled when the bulk kernels ran ...
    const unsigned long long* ucount;     // ... and how many (device): they go through the repair kernel
};

// Several independent chains in one call (BASELINE config 5: the points of several missions that fall into one window of
// model records): the track arrays are the concatenation of `n` sub-tracks, sub-track s starts at point start[s]
// (start[0] == 0, increasing) and its first nearest-point search is seeded with (sj[s], si[s]) -- no point of one
// sub-track ever seeds a point of another.  n == 1 is the ordinary call.
#define GZB_MAX_SEGS 16
struct GzbSegs {
    int n;
    long long start[GZB_MAX_SEGS];
    long long sj[GZB_MAX_SEGS], si[GZB_MAX_SEGS];
};

struct GzbArena {
    char* base = nullptr;
    size_t cap = 0;
    size_t off = 0;
};

// remote output copies of K4 (see gzb_point.cuh); declared here because the context stores one
#define GZB_MAX_PEERS 15
struct GzbPeerOut {
    int n;
    double* np[GZB_MAX_PEERS];
    double* bl[GZB_MAX_PEERS];
};

struct gzb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;     // chain part of K1 while K2+K3 run on `stream` (gzb_mapping)
    cudaStream_t push_stream = nullptr;    // copies of finished chunks of the two series into the peers' buffers (k_push)
    cudaEvent_t ev_push[9] = {};           // chunk c of the bulk kernel done (0..7); all pushes done (8)
    int push_blocks = 64;                  // option "push_blocks": grid of k_push (blocks of 256 threads)
    int push_mode = 0;                     // option "push_mode": 0 = k_push (stores over NVLink issued by SMs), 1 = one peer copy per peer and series on the copy engines (peer_stream[])
    cudaStream_t peer_stream[GZB_MAX_PEERS] = {};   // push_mode 1: one stream per peer, created on first use
    cudaEvent_t ev_peer[GZB_MAX_PEERS] = {};
    int peer_streams_used = 0;             // how many of them carry copies of the last bulk call
    int push_pending = 0;                  // pushes of the last bulk call are in flight on push_stream (ev_push[8] marks their end)
    int bulk_chunks = 0;                   // option "bulk_chunks": 0 = automatic (one chunk: measured best at 2 and at 8 GPUs), else 1..8
    cudaEvent_t ev_k1[3] = {nullptr, nullptr, nullptr};   // 0: search kernel done; 1: chain part done; 2: the auxiliary stream has passed its wait on 0
    int carveout = -1;                     // option "carveout": preferred shared-memory carve-out (percent of the SM's maximum) of the kernels of a step, -1 = the driver's choice per kernel (see gzb_apply_carveout)
    int chain_first = 6;                   // option "chain_first": n >= 2 = the bulk kernel is released n microseconds after the search kernel ends, behind the warp search (step 212 -> 200 us); 1 = released through the auxiliary stream (no effect); 0 = at once
    int bulk_resident = 0;                 // option "bulk_resident": n > 0 = the fused bulk kernel runs as n persistent blocks per SM (grid-stride), leaving the rest of every SM to the chain kernels; 0 = one block per 128 points
    GzbPeerOut peer_out = {0, {}, {}};   // gzb_set_peer_outputs: where K4 also writes its two series
    double* peer_local_np = nullptr;     // ... when its local outputs are these buffers
    double* peer_local_bl = nullptr;
    int fuse_k4 = 1;                 // gzb_map_interp: 1 K2+K3+K4 in one kernel (mesh and weights never re-read), 0 K2+K3 then K4
    unsigned char* d_rowsafe = nullptr;   // per row: East-West twin shortcut allowed (k_rowsafe), for k_ew_per = rowsafe_kew
    int rowsafe_kew = -2;
    int twin_chain = 1;              // 1: the speculative chain of K1 knows the East-West twins (two-state scan); 0: R = E(t) only
    int trace = 0;                   // option "trace": record a timing event at every gzb_trace() mark, printed by gzb_sync
    int ntrace = 0;
    cudaEvent_t trace_ev[48] = {};
    const char* trace_name[48] = {};
    GzbSegs segs_next = {0, {}, {}, {}};   // gzb_map_interp_multi: the sub-tracks of the NEXT K1 call (n == 0: one chain, seeded by the call's arguments)
    int pdl = 1;                     // K1 chain kernels launched with programmatic dependent launch (griddepcontrol): the next kernel's launch overlaps the tail of the previous one
    int overlap = 1;                 // gzb_mapping: 1 run the chain replays beside K2+K3, 0 strictly in sequence
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    GzbGrid g;
    double* d_lat = nullptr;
    double* d_lon = nullptr;
    double2* d_ll = nullptr;
    double* d_angle = nullptr;
    int8_t* d_mask = nullptr;
    float* d_cells = nullptr;
    float4* d_nodes[GZB_MAXLEV] = {};
    float* d_cert = nullptr;
    float4* d_cellv = nullptr;
    unsigned* d_lut[2] = {nullptr, nullptr};
    unsigned* d_mbits = nullptr;     // land-sea mask as bits in 16x16-cell tiles (K4); d_flags[2] != 0: mask not {0,1}
    int mask_bits_ok = 0;
    int prefetch = 3;                // field gather from device-resident records: 3 (default) = 64-byte L2 fetches (K4 alone 38.9 -> 36.5 us, the fused kernel unchanged); 0 plain loads; 1 / 2 = L2 prefetch once the mesh is known (slower); 4 = prefetch of the whole 3x3 neighbourhood as soon as the nearest point is known (slower: 92 -> 98 us)
    int pf_now = 0;                  // ... resolved per call (off for host-resident slabs)
    int k4_early = 0;                // K4 load order: 1 field loads issued before the ocean test, 0 after it (measured faster)
    double* d_hdg = nullptr;         // heading table of K2 (48 B per cell), built on demand
    int hdg_mode = 1;                // 0 never, 1 when the points mapped on this grid outnumber its cells / 2, 2 at the next K2 call
    int hdg_built = 0;
    double lat_min = 0.0, lat_max = 0.0;   // extent of the grid latitudes (gzb_lat_range)
    int64_t k2_points = 0;           // points that went through K2 on this context
    int nearest_path = 1;            // 1: per-thread hill climb seeded by the look-up table (+ warp fallback); 0: warp walk only
    GzbArena arena;
    void* pinned = nullptr;          // small pinned scratch for status words
    size_t pinned_cap = 0;
    void* stage[2] = {nullptr, nullptr};   // pinned staging for pageable field slabs
    size_t stage_cap = 0;
    int nearest_stats_pending = 0;
    int edge_exclusion = 1;          // 0: plain NearestPoint() semantics (no BilinTrack.nrpt edge rule)
    int force_ix64 = 0;              // option "ix64": 1 = K2-K4 run their 64-bit index variants whatever the grid size (tests)
    int force_heavy = 0;             // 1: heavy-duty quadrant search even if the light test succeeds; 2: skip the light test
    int time_err_pending = 0;
    int barrier_err_pending = 0;     // d_flags[3] != 0: a peer barrier timed out (checked by gzb_sync)
    int* d_flags = nullptr;          // persistent device status words (word 0: K4 time-range error)
    unsigned long long* d_kt = nullptr;   // KTIME build: the 64 time-mark words
    int zero_copy = 1;               // gather from pinned+mapped host slabs instead of copying them
    // corner-box bounding sphere cache (predecessor == (-1,-1)), keyed by np_box_r
    int corner_r = -1;
    double corner[4] = {0, 0, 0, -1};
    gzb_stats stats;
    char err[512];
};

// ------------------------------------------------------------------ host-side error plumbing
void gzb_set_global_error(const char* msg);
int gzb_fail(gzb_ctx* ctx, int code, const char* fmt, ...);

#define GZB_CUDA(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return gzb_fail((ctx), GZB_ERR_CUDA, "%s failed: %s (%s:%d)", #call,         \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                \
    } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched with the ProgrammaticStreamSerialization attribute may
// start while its predecessor in the stream is still finishing; gzb_pdl_wait() at its top blocks until the
// predecessor has completed and its writes are visible (a no-op for a plain launch).  gzb_pdl_trigger() lets the
// dependent launch before the calling grid ends (only used by one-block kernels: a resident dependent holds SM slots).
#ifdef __CUDACC__
__device__ __forceinline__ void gzb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gzb_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---- CHECKED build: GZB_NVCC_EXTRA="-DGZB_CHECKED=1" python -m gonzag_b200.build --force -----------------------------
// compute-sanitizer is not available on the GPU pool, so the library carries its own bounds checks: every index a kernel
// computes into a list, a log or a workspace slot, and every flat grid index before it is dereferenced, is compared with
// its bound; a violation sets a bit of a device word (the access itself is skipped where it can be) and the next gzb_sync
// fails with GZB_ERR_STATE naming the bits.  The product build compiles the checks away.
//   1 unresolved list   2 break list   4 changed-points list   8 replay log / NP index   16 candidate queue
//   32 flat grid index  64 mesh corner index  128 push range
#ifdef GZB_CHECKED
#define GZB_CHK(g, cond, bit) ((cond) ? true : (atomicOr((g).chk, (bit)), false))
#else
#define GZB_CHK(g, cond, bit) (true)
#endif

// ---- KTIME build: GZB_NVCC_EXTRA="-DGZB_KTIME=1" python -m gonzag_b200.build --force -------------------------------------
// In-kernel time marks (%globaltimer) of the kernels of a step: thread 0 of every block folds its time into slot s with
// atomicMin (first block to get there: slots 0..31) or atomicMax (last block to get there: slots 32..63); gzb_sync prints
// the table relative to the earliest mark and resets it.  Shows launch gaps and the phases INSIDE the chain kernels,
// which stream events cannot (and recording events switches the dependent launches off).  Compiled away in the product.
#ifdef __CUDACC__
#ifdef GZB_KTIME
__device__ __forceinline__ unsigned long long gzb_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define GZB_KT_MIN(g, s) do { if (threadIdx.x == 0 && (g).kt) atomicMin((g).kt + (s), gzb_gtime()); } while (0)
#define GZB_KT_MAX(g, s) do { if (threadIdx.x == 0 && (g).kt) atomicMax((g).kt + 32 + (s), gzb_gtime()); } while (0)
#else
#define GZB_KT_MIN(g, s) do { } while (0)
#define GZB_KT_MAX(g, s) do { } while (0)
#endif
#endif

// ---- compile-time experiments (GZB_NVCC_EXTRA of gonzag_b200/build.py); all OFF in the product build ----------
// GZB_EXP_L64=1          every sparse per-cell gather of K2 / K3 / K4 (coordinates, angle, heading table, mask words)
//                        asks L2 for 64 bytes instead of a 128-byte line (ld.global.nc.L2::64B)
// GZB_EXP_HDG_STRIDE=8   heading table rows of 8 doubles (64 B, one fetch) instead of 6 (48 B)
// GZB_EXP_BULK_BLOCKS=n  __launch_bounds__(128, n) on the fused K2+K3+K4 kernel (register cap 65536 / (128 n))
// GZB_EXP_SKIP_DG=0      (ON in the product since round 2: K1 141 -> 133 us, step 227 -> 217 us on C3, parity tests green)
//                        k_near_search does not evaluate the fp64 haversine of a point's nearest cell when that cell is
//                        alone within the fp32 budget AND its fp32 chord is farther than twice that budget from the
//                        chords of both acceptance thresholds (r0, 1.2^2 r0): dG then holds 12720 asin(chord / 2),
//                        which is on the same side of every threshold as the exact value -- the only thing dG is
//                        used for (accept_edge, twin_of, point_sum).  0 = always the exact value.
#ifndef GZB_EXP_SKIP_DG
#define GZB_EXP_SKIP_DG 1
#endif
#ifndef GZB_EXP_L64
#define GZB_EXP_L64 0
#endif
#ifndef GZB_EXP_HDG_STRIDE
#define GZB_EXP_HDG_STRIDE 6
#endif
#ifdef __CUDACC__
#if GZB_EXP_L64
__device__ __forceinline__ double gzb_lds(const double* p) {
    double v;
    asm volatile("ld.global.nc.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 gzb_lds(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L2::64B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned gzb_lds(const unsigned* p) {
    unsigned v;
    asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
#define GZB_LDS_LDG(p) gzb_lds(p)      /* replaces __ldg(p) */
#define GZB_LDS_REF(p) gzb_lds(p)      /* replaces *(p)     */
#else
#define GZB_LDS_LDG(p) __ldg(p)
#define GZB_LDS_REF(p) (*(p))
#endif
#endif

#define GZB_LAUNCH_CHECK(ctx)                                                            \
    do {                                                                                 \
        (ctx)->stats.kernel_launches++;                                                  \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess)                                                          \
            return gzb_fail((ctx), GZB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",     \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                \
    } while (0)

cudaError_t gzb_pool_malloc(void** p, size_t bytes);          // device memory through the process-wide pool
cudaError_t gzb_pool_free(void* p);
int gzb_arena_reserve(gzb_ctx* ctx, size_t bytes);            // grow (contents lost), reset offset
void* gzb_arena_take(gzb_ctx* ctx, size_t bytes);             // 256-byte aligned slice or null
int gzb_use_device(gzb_ctx* ctx);
void gzb_trace(gzb_ctx* ctx, cudaStream_t s, const char* what);   // option "trace": timestamped marks on any stream (debugging aid)
int gzb_build_mask_bits(gzb_ctx* ctx);
// Kernels that ask for different shared-memory / L1 splits cannot share an SM: the SM has to drain before its split changes.
// The bulk kernels use no shared memory (the driver gives them the whole array as L1), the chain kernels of K1 that run
// beside them a few KB per block -- so every launch of a chain kernel waits for SMs to drain (in-kernel time marks: the warp
// search starts ~20 us after the search kernel ends when the bulk kernel got there first; with persistent bulk blocks, only
// when the bulk kernel ends).  One preferred carve-out for all the kernels of a step lets their blocks co-reside at once.
void gzb_apply_carveout(gzb_ctx* ctx);
void gzb_carveout_nearest(int pct);      // gzb_nearest.cu
void gzb_carveout_mesh(int pct);         // gzb_mesh.cu                        // gzb_interp.cu: after every change of the mask

static inline size_t gzb_align(size_t n) { return (n + 255) & ~size_t(255); }

// ------------------------------------------------------------------ device math helpers
#ifdef __CUDACC__

// Python's float `%` (and numpy.remainder): result takes the sign of the divisor.
__device__ __forceinline__ double gzb_pymod(double a, double b) {
    // |a| < b, a != 0: fmod(a, b) == a exactly, so the general path below reduces to these two lines
    if (a > 0.0 && a < b) return a;
    if (a < 0.0 && -a < b) return a + b;
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0.0) != (m < 0.0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

// gonzag/utils.py:24-33 (degE_to_degWE)
__device__ __forceinline__ double gzb_pm180(double x) {
    return copysign(1.0, 180.0 - x) * fmin(x, fabs(x - 360.0));
}

// gonzag/bilin_mapping.py:16-47 (Heading): loxodromic bearing a->b, degrees in [0,360).
// Split in two so that a caller taking several headings from one point evaluates that point's
// Mercator ordinate once (same operations, same bits).
__device__ __forceinline__ double gzb_merc_y(double lat) {      // :29-35
    const double r = fmax(fabs(tan(GZB_PI / 4.0 - (lat * GZB_D2R) / 2.0)), 1.0e-9);
    return -log(r);
}

__device__ __forceinline__ double gzb_heading_y(double ya, double xa, double yb, double xb) {   // xa, xb in radians
    double d = gzb_pymod(xb - xa, 2.0 * GZB_PI);
    if (d >= GZB_PI) d = d - 2.0 * GZB_PI;
    if (d <= -GZB_PI) d = d + 2.0 * GZB_PI;
    const double ang = atan2(d, yb - ya);
    return gzb_pymod(ang * 180.0 / GZB_PI, 360.0);
}

__device__ __forceinline__ double gzb_heading_dev(double lata, double lona, double latb, double lonb) {
    return gzb_heading_y(gzb_merc_y(lata), lona * GZB_D2R, gzb_merc_y(latb), lonb * GZB_D2R);
}

// gonzag/utils.py:296-316 (Haversine), same operation order; cos_plat = cos(plat*to_rad)
__device__ __forceinline__ double gzb_haversine_dev(double plat, double plon, double cos_plat,
                                                    double glat, double glon) {
    double a1 = sin(0.5 * ((glat - plat) * GZB_D2R));
    double a2 = sin(0.5 * ((glon - plon) * GZB_D2R));
    double a3 = cos(glat * GZB_D2R) * cos_plat;
    return 2.0 * 6360.0 * asin(sqrt(a1 * a1 + a3 * a2 * a2));
}

// gonzag/bilin_mapping.py:50-141 (AlfaBeta).  y[0],x[0] = target, 1..4 = mesh corners.
__device__ __forceinline__ void gzb_alfabeta_dev(const double* yin, const double* xin,
                                                 double& alfa, double& beta) {
    double y[5], x[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { y[k] = yin[k]; x[k] = xin[k]; }
    const bool wrap = (fabs(x[1] - x[4]) >= 180.0) || (fabs(x[1] - x[2]) >= 180.0) ||
                      (fabs(x[1] - x[3]) >= 180.0);
    if (wrap) {
#pragma unroll
        for (int k = 0; k < 5; ++k) x[k] = gzb_pm180(x[k]);
    }
    double step = 1000.0, ex = 0.5, ey = 0.5, a = 0.0, b = 0.0;
    int it = 0;
    while ((step > 0.1) && (it < 100)) {
        const double e12 = x[2] - x[1];
        const double e41 = x[1] - x[4];
        const double e23 = x[3] - x[2];
        const double j00 = e12 + (e41 + e23) * b;
        const double j01 = -e41 + (e41 + e23) * a;
        const double j10 = y[2] - y[1] + (y[1] - y[4] + y[3] - y[2]) * b;
        const double j11 = y[4] - y[1] + (y[1] - y[4] + y[3] - y[2]) * a;
        const double det = j00 * j11 - j01 * j10;
        if (det == 0.0) {
            // the reference "gives up" with a no-op and spins to the iteration cap
            it = 100;
            break;
        }
        const double da = (ex * j11 - j01 * ey) / det;
        const double db = (j00 * ey - ex * j10) / det;
        step = sqrt(da * da + db * db);
        a = a + da;
        b = b + db;
        const double ca = 1.0 - a;
        const double cb = 1.0 - b;
        ex = x[0] - (ca * cb * x[1] + a * cb * x[2] + a * b * x[3] + ca * b * x[4]);
        ey = y[0] - (ca * cb * y[1] + a * cb * y[2] + a * b * y[3] + ca * b * y[4]);
        it = it + 1;
    }
    if (it >= 100) { a = 0.5; b = 0.5; }
    if (y[1] == y[2] && y[2] == y[3] && y[3] == y[4]) { a = 0.5; b = 0.5; }
    alfa = a;
    beta = b;
}

// warp-wide minimum of a non-negative float (or +inf): one REDUX on the bit pattern
__device__ __forceinline__ float gzb_warp_min_nonneg(float v) {
    return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v)));
}

// bin of a (finite) position in the look-up table
__device__ __forceinline__ long long gzb_lut_bin(const GzbLut& L, double lat, double lon) {
    const double fj = floor((lat - L.lat0) * L.inv_d);
    const int bj = fj < 0.0 ? 0 : (fj >= (double)L.nlat ? L.nlat - 1 : (int)fj);
    const double x = gzb_pymod(lon - L.lon0, 360.0);
    int bi;
    if (!L.wrap && x > L.lon_span) {
        bi = (x - L.lon_span > 0.5 * (360.0 - L.lon_span)) ? 0 : L.nlon - 1;   // outside: the nearer edge
    } else {
        const double fi = floor(x * L.inv_d);
        bi = fi >= (double)L.nlon ? L.nlon - 1 : (int)fi;
    }
    return (long long)bj * L.nlon + bi;
}

// index of leaf tile (tJ, tI) in the level-0 node array (and in GzbGrid::cert)
__device__ __forceinline__ long long gzb_leaf_index(const GzbGrid& g, int tJ, int tI) {
    if (g.nlev == 1) return (long long)tJ * g.lev[0].nI + tI;
    const GzbLevel& P = g.lev[1];
    const int pJ = tJ >> P.sbj, pI = tI >> P.sbi;
    return ((long long)pJ * P.nI + pI) * 32 + ((tJ - (pJ << P.sbj)) << P.sbi) + (tI - (pI << P.sbi));
}

#endif  // __CUDACC__
