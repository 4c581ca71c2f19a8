// C-ABI wrappers of the mapping stages: host/device placement of the arguments, workspace,
// and the fused K1+K2+K3 entry point.
#include "gzb_common.cuh"

size_t gzb_nearest_ws_bytes(int64_t nt);
int gzb_nearest_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, double rd_found_km,
                    int np_box_r, int64_t seed_j, int64_t seed_i, int64_t* np_out, bool split = false,
                    const long long** chg_out = nullptr, const unsigned long long** nchg_out = nullptr,
                    GzbGlobalNearest* gn_out = nullptr);
int gzb_path_fix_dev(gzb_ctx* ctx, const double* yt, const double* xt, const int64_t* np, int64_t* sm_out,
                     double* wb_out, const long long* list, const unsigned long long* count, const long long* list2,
                     const unsigned long long* count2, const double* tt,
                     int64_t nrec, const double* tmod, const void* slab_dev, int dtype, int64_t k0, int64_t nk,
                     double rmiss, double* vnp_out, double* vbl_out);
int gzb_source_mesh_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* np,
                        int64_t* sm_out);
int gzb_weights_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* sm,
                    double* wb_out);

int gzb_mesh_weights_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* np,
                         int64_t* sm_out, double* wb_out, const GzbGlobalNearest* gn_or_null = nullptr);

int gzb_interp_dev(gzb_ctx* ctx, int64_t nt, const double* t, const int64_t* sm, const double* wb, int64_t nrec,
                   const double* tmod, const void* slab_dev, int dtype, int64_t k0, int64_t nk, double rmiss,
                   double* np_out, double* bl_out, const long long* list, const unsigned long long* count);
int gzb_slab_device_pointer(const void* slab, const void** dev_out, int* is_host);
int gzb_push_join(gzb_ctx* ctx);

int gzb_mesh_weights_interp_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* np,
                                const GzbGlobalNearest* gn_or_null, int64_t* sm_out, double* wb_out, const double* tt, int64_t nrec, const double* tmod,
                                const void* slab_dev, int dtype, int64_t k0, int64_t nk, double rmiss, double* vnp_out,
                                double* vbl_out);

namespace {

// which stages a call runs
enum { ST_NP = 1, ST_SM = 2, ST_WB = 4 };

int run_stages(gzb_ctx* ctx, int stages, int64_t nt, const double* yt, const double* xt, double rd, int r,
               int64_t seed_j, int64_t seed_i, const int64_t* np_in, const int64_t* sm_in, int64_t* np_out,
               int64_t* sm_out, double* wb_out, int where, const char* who) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "%s: null ctx", who);
    if (nt < 0) return gzb_fail(ctx, GZB_ERR_ARG, "%s: nt < 0", who);
    if (where != GZB_HOST && where != GZB_DEVICE) return gzb_fail(ctx, GZB_ERR_ARG, "%s: bad `where`", who);
    if (nt == 0) return GZB_OK;
    if (!yt || !xt) return gzb_fail(ctx, GZB_ERR_ARG, "%s: null track coordinates", who);
    if ((stages & ST_NP) && !np_out) return gzb_fail(ctx, GZB_ERR_ARG, "%s: null np_out", who);
    if ((stages & ST_SM) && (!sm_out || (!(stages & ST_NP) && !np_in))) return gzb_fail(ctx, GZB_ERR_ARG, "%s: null np/sm", who);
    if ((stages & ST_WB) && (!wb_out || (!(stages & ST_SM) && !sm_in))) return gzb_fail(ctx, GZB_ERR_ARG, "%s: null sm/wb", who);
    if ((stages & ST_NP) && !(rd > 0.0)) return gzb_fail(ctx, GZB_ERR_ARG, "%s: rd_found_km must be > 0", who);
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    const size_t n = (size_t)nt;
    size_t need = 0;
    if (stages & ST_NP) need += gzb_nearest_ws_bytes(nt);
    if (where == GZB_HOST) need += 2 * gzb_align(n * 8) + gzb_align(n * 16) + gzb_align(n * 64) + gzb_align(n * 32);
    if (gzb_arena_reserve(ctx, need) != GZB_OK) return GZB_ERR_NOMEM;
    cudaStream_t s = ctx->stream;
    const double* dy = yt;
    const double* dx = xt;
    const int64_t* dnp_in = np_in;
    const int64_t* dsm_in = sm_in;
    int64_t* dnp = np_out;
    int64_t* dsm = sm_out;
    double* dwb = wb_out;
    if (where == GZB_HOST) {
        double* a = (double*)gzb_arena_take(ctx, n * 8);
        double* b = (double*)gzb_arena_take(ctx, n * 8);
        dnp = (int64_t*)gzb_arena_take(ctx, n * 16);
        dsm = (int64_t*)gzb_arena_take(ctx, n * 64);
        dwb = (double*)gzb_arena_take(ctx, n * 32);
        GZB_CUDA(ctx, cudaMemcpyAsync(a, yt, n * 8, cudaMemcpyHostToDevice, s));
        GZB_CUDA(ctx, cudaMemcpyAsync(b, xt, n * 8, cudaMemcpyHostToDevice, s));
        ctx->stats.h2d_bytes += n * 16;
        dy = a;
        dx = b;
        if (!(stages & ST_NP) && (stages & ST_SM)) {
            GZB_CUDA(ctx, cudaMemcpyAsync(dnp, np_in, n * 16, cudaMemcpyHostToDevice, s));
            ctx->stats.h2d_bytes += n * 16;
            dnp_in = dnp;
        }
        if (!(stages & ST_SM) && (stages & ST_WB)) {
            GZB_CUDA(ctx, cudaMemcpyAsync(dsm, sm_in, n * 64, cudaMemcpyHostToDevice, s));
            ctx->stats.h2d_bytes += n * 64;
            dsm_in = dsm;
        }
    }
    int rc;
    if (stages == (ST_NP | ST_SM | ST_WB) && ctx->overlap) {
        // K1's chain replays are a handful of long dependent walks that leave the GPU almost idle:
        // run them on the auxiliary stream while K2+K3 work on the speculative chain, then redo
        // K2+K3 for the few points the replays changed
        const long long* chg = nullptr;
        const unsigned long long* nchg = nullptr;
        GzbGlobalNearest gn;
        if ((rc = gzb_nearest_dev(ctx, nt, dy, dx, rd, r, seed_j, seed_i, dnp, true, &chg, &nchg, &gn)) != GZB_OK) return rc;
        if ((rc = gzb_mesh_weights_dev(ctx, nt, dy, dx, dnp, dsm, dwb, &gn)) != GZB_OK) return rc;
        GZB_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_k1[1], 0));
        if ((rc = gzb_path_fix_dev(ctx, dy, dx, dnp, dsm, dwb, chg, nchg, gn.ulist, gn.ucount, nullptr, 0, nullptr, nullptr,
                                   0, 0, 0, 0.0, nullptr, nullptr)) != GZB_OK) return rc;
    } else {
    if (stages & ST_NP) {
        if ((rc = gzb_nearest_dev(ctx, nt, dy, dx, rd, r, seed_j, seed_i, dnp)) != GZB_OK) return rc;
        dnp_in = dnp;
    }
    if ((stages & ST_SM) && (stages & ST_WB)) {
        if ((rc = gzb_mesh_weights_dev(ctx, nt, dy, dx, dnp_in, dsm, dwb)) != GZB_OK) return rc;
    } else if (stages & ST_SM) {
        if ((rc = gzb_source_mesh_dev(ctx, nt, dy, dx, dnp_in, dsm)) != GZB_OK) return rc;
    } else if (stages & ST_WB) {
        if ((rc = gzb_weights_dev(ctx, nt, dy, dx, dsm_in, dwb)) != GZB_OK) return rc;
    }
    }
    if (where == GZB_HOST) {
        if (stages & ST_NP) {
            GZB_CUDA(ctx, cudaMemcpyAsync(np_out, dnp, n * 16, cudaMemcpyDeviceToHost, s));
            ctx->stats.d2h_bytes += n * 16;
        }
        if (stages & ST_SM) {
            GZB_CUDA(ctx, cudaMemcpyAsync(sm_out, dsm, n * 64, cudaMemcpyDeviceToHost, s));
            ctx->stats.d2h_bytes += n * 64;
        }
        if (stages & ST_WB) {
            GZB_CUDA(ctx, cudaMemcpyAsync(wb_out, dwb, n * 32, cudaMemcpyDeviceToHost, s));
            ctx->stats.d2h_bytes += n * 32;
        }
        GZB_CUDA(ctx, cudaStreamSynchronize(s));
    }
    return GZB_OK;
}

}  // namespace

extern "C" int gzb_nearest(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, double rd_found_km,
                           int np_box_r, int64_t seed_j, int64_t seed_i, int64_t* np_out, int where) {
    return run_stages(ctx, ST_NP, nt, yt, xt, rd_found_km, np_box_r, seed_j, seed_i, nullptr, nullptr, np_out,
                      nullptr, nullptr, where, "gzb_nearest");
}

extern "C" int gzb_source_mesh(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* np,
                               int64_t* sm_out, int where) {
    return run_stages(ctx, ST_SM, nt, yt, xt, 1.0, 0, 0, 0, np, nullptr, nullptr, sm_out, nullptr, where,
                      "gzb_source_mesh");
}

extern "C" int gzb_weights(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* sm,
                           double* wb_out, int where) {
    return run_stages(ctx, ST_WB, nt, yt, xt, 1.0, 0, 0, 0, nullptr, sm, nullptr, nullptr, wb_out, where,
                      "gzb_weights");
}

extern "C" int gzb_mesh_weights(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* np,
                                int64_t* sm_out, double* wb_out, int where) {
    return run_stages(ctx, ST_SM | ST_WB, nt, yt, xt, 1.0, 0, 0, 0, np, nullptr, nullptr, sm_out, wb_out, where,
                      "gzb_mesh_weights");
}

extern "C" int gzb_mapping(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, double rd_found_km,
                           int np_box_r, int64_t seed_j, int64_t seed_i, int64_t* np_out, int64_t* sm_out,
                           double* wb_out, int where) {
    return run_stages(ctx, ST_NP | ST_SM | ST_WB, nt, yt, xt, rd_found_km, np_box_r, seed_j, seed_i, nullptr,
                      nullptr, np_out, sm_out, wb_out, where, "gzb_mapping");
}

// K2 + K3 + K4 in ONE kernel from known nearest points, device-resident track: the bulk kernel of
// gzb_map_interp on its own (mesh and weights of a point stay in registers between the stages).
extern "C" int gzb_mesh_weights_interp(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const double* tt,
                                       const int64_t* np, int64_t nrec, const double* tmod, const void* slab, int dtype,
                                       int64_t k0, int64_t nk, double rmissval, int64_t* sm_out, double* wb_out,
                                       double* vnp_out, double* vbl_out) {
    const char* who = "gzb_mesh_weights_interp";
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "%s: null ctx", who);
    if (nt < 0 || nrec < 2 || nk < 2 || k0 < 0 || k0 + nk > nrec || !tmod || !slab)
        return gzb_fail(ctx, GZB_ERR_ARG, "%s: bad arguments (nt=%lld nrec=%lld k0=%lld nk=%lld)", who, (long long)nt,
                        (long long)nrec, (long long)k0, (long long)nk);
    if (nt == 0) return GZB_OK;
    if (!yt || !xt || !tt || !np || !sm_out || !wb_out || !vnp_out || !vbl_out)
        return gzb_fail(ctx, GZB_ERR_ARG, "%s: null track array", who);
    if (dtype != GZB_F32 && dtype != GZB_F64) return gzb_fail(ctx, GZB_ERR_ARG, "%s: dtype must be GZB_F32 or GZB_F64", who);
    if (!ctx->g.mask) return gzb_fail(ctx, GZB_ERR_STATE, "%s: the context has no land-sea mask (gzb_set_mask)", who);
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    const void* slab_dev = nullptr;
    int slab_is_host = 0;
    if (!gzb_slab_device_pointer(slab, &slab_dev, &slab_is_host))
        return gzb_fail(ctx, GZB_ERR_ARG, "%s: the record slab must be in device memory or in pinned + mapped host memory "
                        "(gzb_host_alloc / gzb_host_register)", who);
    if (slab_is_host) ctx->stats.h2d_bytes += (uint64_t)nt * 8 * 32;
    ctx->pf_now = ctx->prefetch && !slab_is_host;
    const int rc = gzb_mesh_weights_interp_dev(ctx, nt, yt, xt, np, nullptr, sm_out, wb_out, tt, nrec, tmod, slab_dev, dtype, k0, nk,
                                               rmissval, vnp_out, vbl_out);
    return rc != GZB_OK ? rc : gzb_push_join(ctx);
}

// K1 + K2 + K3 + K4 in one call, device-resident track.  The chain replays of K1 (a few long
// dependent walks) run on the auxiliary stream while K2, K3 and K4 work on the speculative chain;
// the points the replays changed (a few hundred on a 10-day track) then go through K2-K4 again.
extern "C" int gzb_map_interp(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const double* tt,
                              double rd_found_km, int np_box_r, int64_t seed_j, int64_t seed_i, int64_t nrec,
                              const double* tmod, const void* slab, int dtype, int64_t k0, int64_t nk,
                              double rmissval, int64_t* np_out, int64_t* sm_out, double* wb_out, double* vnp_out,
                              double* vbl_out) {
    const char* who = "gzb_map_interp";
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "%s: null ctx", who);
    if (nt < 0 || nrec < 2 || nk < 2 || k0 < 0 || k0 + nk > nrec || !tmod || !slab)
        return gzb_fail(ctx, GZB_ERR_ARG, "%s: bad arguments (nt=%lld nrec=%lld k0=%lld nk=%lld)", who, (long long)nt,
                        (long long)nrec, (long long)k0, (long long)nk);
    if (nt == 0) return GZB_OK;
    if (!yt || !xt || !tt || !np_out || !sm_out || !wb_out || !vnp_out || !vbl_out)
        return gzb_fail(ctx, GZB_ERR_ARG, "%s: null track array", who);
    if (dtype != GZB_F32 && dtype != GZB_F64) return gzb_fail(ctx, GZB_ERR_ARG, "%s: dtype must be GZB_F32 or GZB_F64", who);
    if (!(rd_found_km > 0.0)) return gzb_fail(ctx, GZB_ERR_ARG, "%s: rd_found_km must be > 0", who);
    if (!ctx->g.mask) return gzb_fail(ctx, GZB_ERR_STATE, "%s: the context has no land-sea mask (gzb_set_mask)", who);
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    const void* slab_dev = nullptr;
    int slab_is_host = 0;
    if (!gzb_slab_device_pointer(slab, &slab_dev, &slab_is_host))
        return gzb_fail(ctx, GZB_ERR_ARG, "%s: the record slab must be in device memory or in pinned + mapped host memory "
                        "(gzb_host_alloc / gzb_host_register); use gzb_mapping + gzb_interp for pageable slabs", who);
    if (slab_is_host) ctx->stats.h2d_bytes += (uint64_t)nt * 8 * 32;
    ctx->pf_now = ctx->prefetch && !slab_is_host;
    if (gzb_arena_reserve(ctx, gzb_nearest_ws_bytes(nt)) != GZB_OK) return GZB_ERR_NOMEM;
    cudaStream_t s = ctx->stream;
    int rc;
    const long long* chg = nullptr;
    const unsigned long long* nchg = nullptr;
    const bool split = ctx->overlap != 0;
    GzbGlobalNearest gn;
    if ((rc = gzb_nearest_dev(ctx, nt, yt, xt, rd_found_km, np_box_r, seed_j, seed_i, np_out, split, &chg, &nchg, &gn)) != GZB_OK) return rc;
    const GzbGlobalNearest* gnp = split ? &gn : nullptr;      // overlap: K2-K4 start from E(t), the chain runs beside them
    if (ctx->fuse_k4) {
        if ((rc = gzb_mesh_weights_interp_dev(ctx, nt, yt, xt, np_out, gnp, sm_out, wb_out, tt, nrec, tmod, slab_dev, dtype, k0,
                                              nk, rmissval, vnp_out, vbl_out)) != GZB_OK) return rc;
    } else {
        if ((rc = gzb_mesh_weights_dev(ctx, nt, yt, xt, np_out, sm_out, wb_out, gnp)) != GZB_OK) return rc;
        if ((rc = gzb_interp_dev(ctx, nt, tt, sm_out, wb_out, nrec, tmod, slab_dev, dtype, k0, nk, rmissval, vnp_out,
                                 vbl_out, nullptr, nullptr)) != GZB_OK) return rc;
    }
    gzb_trace(ctx, s, "bulk end (main)");
    if (split) {
        GZB_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_k1[1], 0));
        if ((rc = gzb_path_fix_dev(ctx, yt, xt, np_out, sm_out, wb_out, chg, nchg, gn.ulist, gn.ucount, tt, nrec, tmod, slab_dev,
                                   dtype, k0, nk,
                                   rmissval, vnp_out, vbl_out)) != GZB_OK) return rc;
    }
    if ((rc = gzb_push_join(ctx)) != GZB_OK) return rc;      // (no repair kernel ran: nothing joined the chunk copies yet)
    gzb_trace(ctx, s, "path end (main)");
    return GZB_OK;
}

// Several independent chains in one K1-K4 call (BASELINE config 5: the points of several altimeter missions falling into
// one window of model records).  The track arrays are the concatenation of nseg sub-tracks; sub-track s starts at point
// seg_start[s] (seg_start[0] == 0, strictly increasing) and its first point is seeded with seg_seed[2s], seg_seed[2s+1]
// -- what BilinTrack.nrpt would do for each mission on its own (gonzag/bilin_mapping.py:570, 579-580).  Same results
// as one gzb_map_interp per sub-track; the latency-bound chain part of K1 is paid once.
extern "C" int gzb_map_interp_multi(gzb_ctx* ctx, int nseg, const int64_t* seg_start, const int64_t* seg_seed, int64_t nt,
                                    const double* yt, const double* xt, const double* tt, double rd_found_km, int np_box_r,
                                    int64_t nrec, const double* tmod, const void* slab, int dtype, int64_t k0, int64_t nk,
                                    double rmissval, int64_t* np_out, int64_t* sm_out, double* wb_out, double* vnp_out,
                                    double* vbl_out) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_map_interp_multi: null ctx");
    if (nseg < 1 || nseg > GZB_MAX_SEGS || !seg_start || !seg_seed)
        return gzb_fail(ctx, GZB_ERR_ARG, "gzb_map_interp_multi: 1 <= nseg <= %d sub-tracks", GZB_MAX_SEGS);
    if (seg_start[0] != 0) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_map_interp_multi: seg_start[0] must be 0");
    for (int q = 1; q < nseg; ++q)
        if (seg_start[q] <= seg_start[q - 1] || seg_start[q] >= nt)
            return gzb_fail(ctx, GZB_ERR_ARG, "gzb_map_interp_multi: seg_start must increase strictly inside [0, nt)");
    GzbSegs& S = ctx->segs_next;
    S.n = nseg;
    for (int q = 0; q < nseg; ++q) {
        S.start[q] = seg_start[q];
        S.sj[q] = seg_seed[2 * q];
        S.si[q] = seg_seed[2 * q + 1];
    }
    const int rc = gzb_map_interp(ctx, nt, yt, xt, tt, rd_found_km, np_box_r, seg_seed[0], seg_seed[1], nrec, tmod, slab, dtype,
                                  k0, nk, rmissval, np_out, sm_out, wb_out, vnp_out, vbl_out);
    S.n = 0;          // whatever happened: the next call is an ordinary one
    return rc;
}
