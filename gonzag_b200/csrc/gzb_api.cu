// C-ABI wrappers of the mapping stages: host/device placement of the arguments, workspace,
// and the fused K1+K2+K3 entry point.
#include "gzb_common.cuh"

size_t gzb_nearest_ws_bytes(int64_t nt);
int gzb_nearest_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, double rd_found_km,
                    int np_box_r, int64_t seed_j, int64_t seed_i, int64_t* np_out);
int gzb_source_mesh_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* np,
                        int64_t* sm_out);
int gzb_weights_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* sm,
                    double* wb_out);

int gzb_mesh_weights_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* np,
                         int64_t* sm_out, double* wb_out);

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
