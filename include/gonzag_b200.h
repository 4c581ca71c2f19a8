/*
 * gonzag_b200 -- C ABI of the B200 (sm_100a) implementation of gonzag's
 * gridded -> along-track mapping + interpolation path.
 *
 * The reference (brodeau/gonzag) is pure Python; the interface this library replaces is
 * the Python surface of gonzag/bilin_mapping.py and gonzag/mod2sat.py.  Each entry point
 * below names the reference lines it stands in for.  INTEGRATION.md shows the ctypes stub a
 * gonzag maintainer would add to call them.
 *
 * Conventions
 *   - every function returns an int status (GZB_OK == 0); nothing throws, nothing exits;
 *     gzb_last_error() gives the message of the last failure on a context
 *     (gzb_last_global_error() for failures before a context exists);
 *   - arrays are C-contiguous; grids are row-major (Nj, Ni); indices are int64, weights and
 *     coordinates float64, exactly the dtypes of the reference attributes
 *     BilinTrack.NP (Nt,2), .SM (Nt,4,2), .WB (Nt,4)  (gonzag/bilin_mapping.py:547-549);
 *   - `where` says where the array arguments of a call live: GZB_HOST (caller-owned host
 *     buffers; the call copies in/out and is synchronous on return) or GZB_DEVICE (device
 *     pointers on the context's GPU; the call only enqueues work on the context's stream);
 *   - a context owns the device copy of one model grid and its search structure; it is not
 *     thread-safe (one context per GPU / stream);
 *   - there is no CPU fallback: without a CUDA device gzb_create() fails.
 */
#ifndef GONZAG_B200_H
#define GONZAG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gzb_ctx gzb_ctx;

enum { GZB_OK = 0, GZB_ERR_CUDA = 1, GZB_ERR_ARG = 2, GZB_ERR_NOMEM = 3, GZB_ERR_STATE = 4,
       GZB_ERR_TIME = 5 };
enum { GZB_HOST = 0, GZB_DEVICE = 1 };
enum { GZB_F32 = 0, GZB_F64 = 1 };

/* counters filled by gzb_get_stats() */
typedef struct gzb_stats {
    uint64_t kernel_launches;   /* kernels launched by this context since creation          */
    uint64_t h2d_bytes;         /* bytes copied host->device by this context                */
    uint64_t d2h_bytes;         /* bytes copied device->host by this context                */
    int64_t  last_breaks;       /* last gzb_nearest: points whose seeding had to be replayed */
    int64_t  last_chain_steps;  /* last gzb_nearest: sequential replay steps, all chains     */
    int32_t  pyramid_levels;    /* depth of the bounding-sphere pyramid over the grid       */
    int32_t  l2_fetch_bytes;    /* what the device reports after option "l2_fetch" (0: option never used)  */
    int64_t  lut_bins;          /* bins of the lat-lon seed table (0 = table not built)          */
    int64_t  last_unresolved;   /* last gzb_nearest: points the per-thread search handed to the warp search */
    int64_t  last_unres_noseed; /*   ... of which: no seed in the table                       */
    int64_t  last_unres_noconv; /*   ... hill climb not converged                             */
    int64_t  last_unres_nocert; /*   ... per-cell certificate not sufficient                  */
    int64_t  last_max_chain;    /* last gzb_nearest: longest sequential replay (steps)         */
    int64_t  heading_table;     /* 1 once the per-cell heading table of K2 has been built       */
} gzb_stats;

const char* gzb_version(void);
const char* gzb_last_global_error(void);
int gzb_device_count(int* n_out);

/* ---- context: one model grid resident on one GPU ---------------------------------------
 * Replaces the grid arguments of BilinTrack(Yt,Xt,Ys,Xs,src_grid_local_angle,k_ew_per,...)
 * (gonzag/bilin_mapping.py:533-545) and MG.lat/MG.lon/MG.xangle/MG.EWPer/MG.mask as read by
 * Model2SatTrack (gonzag/mod2sat.py:33-50,124).  lat/lon: (nj,ni) float64 degrees, finite.
 * angle: (nj,ni) float64 degrees or NULL (no -D).  mask: (nj,ni) int8 {0,1} or NULL.
 * The four arrays may live in host memory (pageable or pinned) or in device memory (detected).
 * k_ew_per: -1 none, >=0 overlap of the East-West periodicity. */
int gzb_create(int device, int64_t nj, int64_t ni, const double* lat, const double* lon,
               const double* angle_or_null, const int8_t* mask_or_null, int k_ew_per,
               gzb_ctx** ctx_out);
/* gzb_create with the land-sea mask in the caller's integer type (HOST memory; 1, 2, 4 or 8 bytes per cell, values {0,1}):
 * the reference's MG.mask is int64 (gonzag/ncio.py:164-191); the conversion to int8 rides on the upload's staging copies. */
int gzb_create_imask(int device, int64_t nj, int64_t ni, const double* lat, const double* lon,
                     const double* angle_or_null, const void* mask_or_null, int mask_itemsize, int k_ew_per,
                     gzb_ctx** ctx_out);
int gzb_destroy(gzb_ctx* ctx);
const char* gzb_last_error(const gzb_ctx* ctx);
int gzb_set_stream(gzb_ctx* ctx, void* cuda_stream);      /* a cudaStream_t; NULL = the legacy default stream */
int gzb_use_own_stream(gzb_ctx* ctx);                      /* back to the context's private stream (the default) */
int gzb_sync(gzb_ctx* ctx);
int gzb_get_stats(const gzb_ctx* ctx, gzb_stats* out);
int gzb_set_mask(gzb_ctx* ctx, const int8_t* mask, int where);
int gzb_set_angle(gzb_ctx* ctx, const double* angle_or_null, int where);
int gzb_set_ew_per(gzb_ctx* ctx, int k_ew_per);
/* options: "zero_copy": 1 (default) gzb_interp gathers straight from a pinned + mapped host slab
 * over PCIe when the track is sparse w.r.t. the slab, instead of copying the slab to the GPU;
 * 0 never; 2 whenever the slab is pinned + mapped.
 * "nearest_path": 1 (default) per-thread hill climb seeded by a lat-lon table, proven by per-cell
 * certificates, warp-cooperative search for the rest; 0 warp-cooperative search only (same results).
 * "heading_table": K2 can read the grid-only part of the quadrant test (Mercator ordinate of a
 * cell and the headings to its four neighbours, 48 B per cell) from a table instead of computing
 * it per track point: 0 never, 1 (default) build it once the points mapped on this context
 * outnumber half the grid cells, 2 build it at the next K2 call.  Same results either way.
 * "overlap": 1 (default) gzb_mapping / gzb_map_interp run everything of K1 that follows the
 * per-thread search (warp search of the unsettled points, chain logic, replays) on a second stream
 * beside K2+K3(+K4), which start from the plain global answers, and repair the points whose
 * nearest point turned out different; 0 strictly in sequence.  Same results.
 * "twin_chain": 1 (default) the speculative chain of K1 knows the East-West twins of the overlap
 * columns (two-state scan, far fewer replays); 0 plain speculation + replays.  Same results.
 * "fuse_k4": 1 (default) gzb_map_interp runs K2+K3+K4 as one kernel; 0 K2+K3 then K4.
 * "k4_early", "mask_bits": K4 variants (load order; tiled bit mask on/off).  Same results.
 * "prefetch": variants of the field gather for device-resident slabs, same results: 3 (default) = the eight values are
 * loaded with a 64-byte L2 fetch (ld.global.nc.L2::64B: a plain sparse load fills a whole 128-byte line from DRAM on
 * B200) -- K4 alone 38.9 -> 36.5 us, the fused kernel unchanged; 0 = plain loads; 1 = L2 prefetch of a point's eight field
 * values as soon as its mesh is known (2: the same with 64 registers in K4) -- measured slower; 4 = L2 prefetch of the
 * whole 3x3 neighbourhood (table row, angle, coordinates, mask tile, field rows of both records) as soon as the nearest
 * point is known -- measured slower (fused kernel 92 -> 98 us: it is not DRAM-latency bound).
 * "pdl": 1 (default) the kernels of K1's chain part (k_chain, k_compact, k_walk, k_valid), which follow each
 * other on one stream, are launched as programmatic dependents (griddepcontrol.wait at their top): the
 * launch of each overlaps the tail of the one before it (gzb_nearest 158 -> 143 us on ORCA12); 0 plain
 * launches.  Same results.
 * "l2_fetch": 32 / 64 / 128 = cudaLimitMaxL2FetchGranularity, a DEVICE-WIDE hint for the sparse
 * gathers (not a per-context setting); gzb_stats.l2_fetch_bytes reports what the device answered.
 * "trace": 1 records a timing event at internal marks on both streams; gzb_sync prints them.
 * "edge_exclusion": 1 (default) gzb_nearest applies the first/last row/column rule of
 * BilinTrack.nrpt (gonzag/bilin_mapping.py:582-585); 0 gives plain NearestPoint() results.
 * "force_heavy": 0 (default); 1 = heavy-duty quadrant search as if |angle| > 45 everywhere;
 * 2 = also skip the light test (Iquadran's lforceHD, gonzag/bilin_mapping.py:375).
 * "ix64": 0 (default) K2-K4 use 32-bit row / column / cell indices on every grid of fewer than 2^31 cells (the int64
 * arrays of the reference are widened at the stores, narrowed at the loads: fused kernel 93 -> 82 us on ORCA12);
 * 1 = the 64-bit variants whatever the grid size.  Same results.
 * Scheduling options, all measured on the C3 track (same results in every setting; off by default unless stated):
 * "bulk_chunks": n > 1 cuts the fused bulk kernel of a sharded step into n chunks so that the peers' copies of chunk c
 *   travel while chunk c+1 is computed (0 / 1 = one chunk: at 8 GPUs the arriving stores slow the bulk kernel by 2x);
 * "bulk_resident": n > 0 runs the fused bulk kernel as n persistent blocks per SM (the block scheduler fills SMs one
 *   after the other, so this does NOT leave room on every SM: slower);
 * "chain_first": n >= 2 (default 6, tracks of 100 000 points or more) releases the bulk kernel n microseconds after the search
 *   kernel ends, so that the warp search of K1's chain part is resident first (step 212 -> 200 us; 4 to 24 us measured
 *   alike); 1 = through the auxiliary stream only (the two launches still race); 0 = at once;
 * "carveout": preferred shared-memory carve-out (percent) of all kernels of a step, -1 = the driver's choice (no effect);
 * "push_blocks" (64) grid of the peer-copy kernel, "push_mode" 1 = peer copies on the copy engines (2x slower). */
int gzb_set_option(gzb_ctx* ctx, const char* name, int64_t value);

/* ---- grid heuristics of ModGrid -------------------------------------------
 * gzb_lon_stats: the reductions IsGlobalLongitudeWise() needs (gonzag/utils.py:149-182), on the
 * context's longitudes: vals4 = { min X, max X, min XB, max XB } with X = lon % 360 and
 * XB = degE_to_degWE(X); cols2 = { argmin(X) % ni, argmax(X) % ni } (first occurrences, like
 * numpy).  Host outputs; synchronous.
 * gzb_lon_to_pm180: rewrites the context's longitudes as degE_to_degWE(lon) (what ModGrid does
 * for a regional domain handled in the [-180,180] frame, gonzag/utils.py:441) and optionally
 * copies them to the host array lon_out.
 * gzb_lat_range: min / max of the context's latitudes (ModGrid.domain_bounds, :420-421). */
int gzb_lon_stats(gzb_ctx* ctx, double* vals4, int64_t* cols2);
int gzb_lon_to_pm180(gzb_ctx* ctx, double* lon_out_or_null);
int gzb_lat_range(gzb_ctx* ctx, double* lat_min, double* lat_max);

/* ---- K0: local grid rotation, the `-D` pre-pass --------------------------------------------
 * Replaces GridAngle(xlat, xlon) (gonzag/utils.py:226-262).  Computes the (nj,ni) angle in
 * degrees from the context's lat/lon, installs it as the context's angle and, when
 * angle_out is not NULL, copies it there. */
int gzb_grid_angle(gzb_ctx* ctx, double* angle_out_or_null, int where);
/* Copy of the angle the context holds (installed by gzb_create / gzb_set_angle / gzb_grid_angle); no kernel
 * runs.  What reading `ModGrid.xangle` does (the reference computes the attribute once, gonzag/utils.py:431,
 * on the 0..360 longitudes, before the regional frame switch of :441).  GZB_ERR_STATE without an angle. */
int gzb_get_angle(gzb_ctx* ctx, double* angle_out, int where);

/* ---- K1: nearest points along a track --------------------------------------------------
 * Replaces BilinTrack.nrpt() + NearestPoint() + the edge exclusion
 * (gonzag/bilin_mapping.py:566-593, 144-191).  yt/xt: (nt,) float64.  The previous-hit
 * seeding of consecutive points is reproduced exactly; (seed_j, seed_i) is the pair handed
 * to the first point (the reference uses (np_box_r, np_box_r)).
 * np_out: (nt,2) int64, [j,i], (-1,-1) = not found. */
int gzb_nearest(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt,
                double rd_found_km, int np_box_r, int64_t seed_j, int64_t seed_i,
                int64_t* np_out, int where);

/* ---- K2: source mesh ---------------------------------------------------------------------
 * Replaces BilinTrack.srcm() + IDSourceMesh() + Iquadran() + Iquadran2SrcMesh()
 * (gonzag/bilin_mapping.py:595-608, 453-486, 329-449, 215-262), including the local-angle
 * lookup, the heavy-duty polygon search and the East-West wrap.
 * np: (nt,2) int64.  sm_out: (nt,4,2) int64; all -1 = quadrant not found; all 0 where the
 * nearest point is (-1,-1). */
int gzb_source_mesh(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt,
                    const int64_t* np, int64_t* sm_out, int where);

/* ---- K3: bilinear weights ------------------------------------------------------------------
 * Replaces BilinTrack.wght() + WeightBL() + AlfaBeta()
 * (gonzag/bilin_mapping.py:610-617, 490-527, 50-141).  wb_out: (nt,4) float64. */
int gzb_weights(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt,
                const int64_t* sm, double* wb_out, int where);

/* K2+K3 in one kernel (BilinTrack.srcm() followed by BilinTrack.wght(),
 * gonzag/bilin_mapping.py:562-563): same results as the two calls above, the mesh is written
 * once and never re-read. */
int gzb_mesh_weights(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt,
                     const int64_t* np, int64_t* sm_out, double* wb_out, int where);

/* K1+K2+K3 in one call (what BilinTrack.__init__ does, gonzag/bilin_mapping.py:554-564). */
int gzb_mapping(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt,
                double rd_found_km, int np_box_r, int64_t seed_j, int64_t seed_i,
                int64_t* np_out, int64_t* sm_out, double* wb_out, int where);

/* K1+K2+K3+K4 in one call on a device-resident track (what Model2SatTrack.__init__ does from
 * BilinTrack(...) to the end of its interpolation loop, gonzag/mod2sat.py:49-139).  All track
 * arrays are DEVICE pointers; `slab` holds records k0 .. k0+nk-1 and must be readable by the GPU
 * (device memory, or pinned + mapped host memory) and bracket every track time.  Same results as
 * gzb_mapping followed by gzb_interp; the chain replays of K1 run beside K2-K4. */
int gzb_map_interp(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const double* tt,
                   double rd_found_km, int np_box_r, int64_t seed_j, int64_t seed_i,
                   int64_t nrec, const double* tmod, const void* slab, int dtype, int64_t k0, int64_t nk,
                   double rmissval, int64_t* np_out, int64_t* sm_out, double* wb_out,
                   double* vnp_out, double* vbl_out);

/* The same for SEVERAL independent chains in one call (BASELINE config 5: the points of several altimeter missions that
 * fall into one window of model records).  The track arrays are the concatenation of nseg (<= 16) sub-tracks; sub-track s
 * starts at point seg_start[s] (host array; seg_start[0] == 0, strictly increasing) and its first point is seeded with
 * seg_seed[2s], seg_seed[2s+1] (host array) -- what one BilinTrack per mission would do (gonzag/bilin_mapping.py:570,
 * 579-580).  Same results as one gzb_map_interp per sub-track; the latency-bound chain part of K1 is paid once. */
int gzb_map_interp_multi(gzb_ctx* ctx, int nseg, const int64_t* seg_start, const int64_t* seg_seed, int64_t nt,
                         const double* yt, const double* xt, const double* tt, double rd_found_km, int np_box_r,
                         int64_t nrec, const double* tmod, const void* slab, int dtype, int64_t k0, int64_t nk,
                         double rmissval, int64_t* np_out, int64_t* sm_out, double* wb_out, double* vnp_out,
                         double* vbl_out);

/* K2+K3+K4 in ONE kernel from known nearest points (BilinTrack.srcm() + wght(),
 * gonzag/bilin_mapping.py:595-617, then the interpolation loop, gonzag/mod2sat.py:74-139): the bulk
 * kernel of gzb_map_interp on its own.  Device-resident track, slab as for gzb_map_interp.  Same
 * results as gzb_mesh_weights followed by gzb_interp. */
int gzb_mesh_weights_interp(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const double* tt,
                            const int64_t* np, int64_t nrec, const double* tmod, const void* slab, int dtype,
                            int64_t k0, int64_t nk, double rmissval, int64_t* sm_out, double* wb_out,
                            double* vnp_out, double* vbl_out);

/* ---- K4: fused time-linear + space-bilinear gather -----------------------------------------
 * Replaces the interpolation loop of Model2SatTrack (gonzag/mod2sat.py:74-139).
 * t: (nt,) float64 track times; tmod: (nrec,) float64 model record times, increasing;
 * slab: records [k0, k0+nk) of the model field, (nk,nj,ni) of `dtype` (GZB_F32 / GZB_F64);
 * points whose bracketing records both lie in the slab are written, the others are left
 * untouched, so a caller streams a long series by looping over slabs (never over points).
 * `slab` may be a device pointer, a pinned host pointer or a pageable host pointer whatever
 * `where` says (detected): device slabs are used in place, host slabs are gathered over PCIe
 * (pinned + mapped, sparse track) or streamed through the GPU in 256 MiB chunks.
 * init_outputs != 0 first fills np_out/bl_out with rmissval (do it on the first slab).
 * Needs a mask in the context.  Points outside [tmod[0], tmod[nrec-1]) -> GZB_ERR_TIME
 * (returned by this call when it synchronises, otherwise by the next gzb_sync()). */
int gzb_interp(gzb_ctx* ctx, int64_t nt, const double* t, const int64_t* sm, const double* wb,
               int64_t nrec, const double* tmod, const void* slab, int dtype, int64_t k0,
               int64_t nk, double rmissval, int init_outputs, double* np_out, double* bl_out,
               int where);

/* ---- a14: along-track distance and nearest-point map ---------------------------------------
 * gzb_track_distance: cumulative distance (km) from the first point, Haversine_sclr with the
 * latitude-dependent Earth radius (gonzag/mod2sat.py:146-148, gonzag/utils.py:267-293).
 * gzb_xnp_track: (nj,ni) float64 map holding the last track index that hit each cell,
 * -100 over land, rmissval elsewhere (gonzag/mod2sat.py:177-184).  Needs a mask. */
int gzb_track_distance(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt,
                       double* dist_out, int where);
int gzb_xnp_track(gzb_ctx* ctx, int64_t nt, const int64_t* np, double rmissval,
                  double* xnp_out, int where);
/* same, plus missing_out (nj,ni) uint8 = (xnp == rmissval): the mask of the reference's
 * masked_where(XNPtrack == rmissval, XNPtrack) (gonzag/mod2sat.py:186) without a host pass */
int gzb_xnp_track_masked(gzb_ctx* ctx, int64_t nt, const int64_t* np, double rmissval,
                         double* xnp_out, unsigned char* missing_out, int where);

/* imask of Model2SatTrack (gonzag/mod2sat.py:157-158) on DEVICE arrays: imask_out[t] = 1 iff -100 < bl[t] < 100 and
 * -100 < sat[t] < 100 (int8, the dtype of the reference's attribute `mask`); bad_out (optional, uint8) = its
 * complement, the mask of the reference's three masked_where(imask == 0, .) products (:160-162). */
int gzb_track_mask(gzb_ctx* ctx, int64_t nt, const double* bl, const double* sat, int8_t* imask_out,
                   unsigned char* bad_out_or_null);

/* ---- scalar helpers exported by the reference package, batched ------------------------------
 * gzb_heading  : Heading(lata,lona,latb,lonb)      gonzag/bilin_mapping.py:16-47
 * gzb_alfabeta : AlfaBeta(vy[5], vx[5]) -> (a,b)   gonzag/bilin_mapping.py:50-141
 * gzb_haversine: Haversine(plat,plon,xlat,xlon)     gonzag/utils.py:296-316
 * gzb_haversine_sclr: Haversine_sclr(lat1,lon1,lat2,lon2) gonzag/utils.py:281-293 (latitude-dependent radius)
 * gzb_radius_earth  : RadiusEarth(lat)                     gonzag/utils.py:267-278
 * All arrays HOST, n independent problems per call; they run on `device`. */
int gzb_heading(int device, int64_t n, const double* lata, const double* lona,
                const double* latb, const double* lonb, double* out);
int gzb_alfabeta(int device, int64_t n, const double* vy5, const double* vx5, double* ab_out);
int gzb_haversine(int device, int64_t n, double plat, double plon, const double* xlat,
                  const double* xlon, double* out);
int gzb_haversine_sclr(int device, int64_t n, const double* lat1, const double* lon1,
                       const double* lat2, const double* lon2, double* out);
int gzb_radius_earth(int device, int64_t n, const double* lat, double* out);

/* ---- multi-GPU, one process per GPU: K4 fused with the gather of its results ----------------
 * A rank exports its product buffer (gzb_dev_alloc memory) with gzb_ipc_export; the 64-byte handle
 * travels to the other ranks by any means (torch.distributed here), which map it with gzb_ipc_open.
 * gzb_set_peer_outputs registers, for the local output buffers (vnp_local, vbl_local), n remote
 * copies: from then on gzb_interp / gzb_map_interp calls writing those local buffers with
 * init_outputs = 1 also store every value into the n peer buffers (NVLink stores issued by the
 * gather kernel itself).  n = 0 turns it off.  The caller orders the ranks afterwards (a barrier:
 * peers may read their buffer once every rank's call has completed).
 * gzb_publish_edges copies the first and the last row of an (nrows,2) int64 array (a segment's halo
 * assumption and its last nearest point) into n destinations of 4 int64. */
int gzb_ipc_export(gzb_ctx* ctx, void* dptr, unsigned char* handle64);
int gzb_ipc_open(gzb_ctx* ctx, const unsigned char* handle64, void** dptr_out);
int gzb_ipc_close(gzb_ctx* ctx, void* dptr);
int gzb_set_peer_outputs(gzb_ctx* ctx, int n, void* const* vnp_peers, void* const* vbl_peers,
                         void* vnp_local, void* vbl_local);
int gzb_publish_edges(gzb_ctx* ctx, const int64_t* np_rows, int64_t nrows, int n, void* const* dst);
/* Stream-ordered barrier between the ranks over peer memory (no NCCL call): every rank owns `world`
 * int64 epoch counters (local_bar, zero-initialised, IPC-mapped by the peers); peer_slots[k] is the
 * address, in peer k's counters, of this rank's entry.  Epochs must increase by one per barrier.  A
 * peer that does not arrive within ~4 s is reported by the next gzb_sync (GZB_ERR_STATE). */
int gzb_peer_barrier(gzb_ctx* ctx, int n, void* const* peer_slots, void* local_bar, int world, int my_rank,
                     int64_t epoch);
/* The end of a sharded step in ONE kernel: gzb_publish_edges (np_rows / nrows / nedge / edge_dst), then gzb_peer_barrier
 * (nbar / peer_slots / local_bar / world / my_rank / epoch), then gzb_check_edges on this rank's gathered buffer
 * (gathered / stride_words / offset_words / flag_dev, which may be NULL: no check). */
int gzb_peer_finish(gzb_ctx* ctx, const int64_t* np_rows, int64_t nrows, int nedge, void* const* edge_dst, int nbar,
                    void* const* peer_slots, void* local_bar, int world, int my_rank, int64_t epoch,
                    const int64_t* gathered, int64_t stride_words, int64_t offset_words, int* flag_dev_or_null);
/* *flag_dev |= 1 if, for some rank g > 0, the halo words of row g differ from the last-point words of
 * row g-1 (rows of stride_words int64, the 4 edge words at offset_words). */
int gzb_check_edges(gzb_ctx* ctx, const int64_t* gathered, int64_t stride_words, int64_t offset_words,
                    int world, int* flag_dev);

/* ---- memory helpers so that a host language without a CUDA binding can keep a pipeline on
 * the device (used by the Python layer; not part of the reference's surface) ---------------- */
int gzb_dev_alloc(gzb_ctx* ctx, uint64_t bytes, void** dptr_out);
int gzb_dev_free(gzb_ctx* ctx, void* dptr);
int gzb_h2d(gzb_ctx* ctx, void* dst_dev, const void* src_host, uint64_t bytes);   /* async; >= 24 MiB of pageable memory: parallel staging, returns when done */
int gzb_d2h(gzb_ctx* ctx, void* dst_host, const void* src_dev, uint64_t bytes);   /* idem */
int gzb_d2d(gzb_ctx* ctx, void* dst_dev, const void* src_dev, uint64_t bytes);          /* async, on the context's stream */
int gzb_memset(gzb_ctx* ctx, void* dst_dev, int byte, uint64_t bytes);            /* async */
/* Device buffers of the library come from a process-wide pool (freed blocks are kept, per device
 * and exact size, for the next request of that size: repeated calls on the same grid allocate
 * nothing).  gzb_pool_config(enabled, max bytes kept; default on, 16 GiB), gzb_pool_trim() returns
 * the kept blocks to the driver, gzb_pool_stats reports them. */
int gzb_pool_config(int enabled, uint64_t max_cached_bytes);
int gzb_pool_trim(void);
int gzb_pool_stats(uint64_t* cached_bytes, uint64_t* cached_blocks, uint64_t* live_blocks);
/* Host-only helper (no device, no context): dst[k] = (int8_t)src[k] for an integer land-sea mask of
 * `itemsize` 1, 2, 4 or 8 bytes -- the reference's MG.mask is int64 {0,1} (gonzag/ncio.py:164-191), the
 * context takes int8 -- cut over the host threads (NumPy's single-threaded astype costs a third of an
 * end-to-end call on ORCA12). */
int gzb_pack_mask(const void* src, int itemsize, int64_t n, int8_t* dst);
/* Host-only helper: dst[k] = src[k] for n items of `itemsize` 1, 2, 4 or 8 bytes, bytes of every item reversed
 * when `swap` != 0, cut over the host threads; src and dst must not overlap.  Stages the big-endian records
 * of a NetCDF-3 file (what GetModel2DVar reads, gonzag/ncio.py:194-210) into the pinned slab K4 gathers from. */
int gzb_copy_swap(void* dst, const void* src, int itemsize, int64_t n, int swap);
/* Where does `p` live?  kind_out: 0 pageable host memory (or unknown), 1 pinned + mapped host
 * memory (the GPU can read it in place), 2 device / managed memory. */
int gzb_pointer_kind(const void* p, int* kind_out);
/* Cap on the host threads the library starts for parallel copies, mask packing and record staging (0 = automatic:
 * all hardware threads; also settable with the environment variable GZB_HOST_THREADS).  One process per GPU: give
 * every rank cores / local world size. */
int gzb_set_host_threads(int n);
/* Device pointers of the arrays a context holds as uploaded (lat, lon float64 row-major; angle float64 and mask int8
 * when installed, else NULL): what a multi-GPU host broadcasts so that the other ranks can call gzb_create with
 * DEVICE pointers (gzb_create accepts host or device memory for its four arrays). */
int gzb_grid_arrays(gzb_ctx* ctx, void** lat, void** lon, void** angle, void** mask);
int gzb_host_alloc(uint64_t bytes, void** hptr_out);     /* pinned, portable, mapped */
int gzb_host_free(void* hptr);
int gzb_host_register(void* hptr, uint64_t bytes);
int gzb_host_unregister(void* hptr);

#ifdef __cplusplus
}
#endif
#endif /* GONZAG_B200_H */
