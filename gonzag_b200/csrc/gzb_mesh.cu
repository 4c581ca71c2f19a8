// K2 (source mesh) and K3 (bilinear weights): one thread per track point.
//   K2 restates BilinTrack.srcm / IDSourceMesh / Iquadran / Iquadran2SrcMesh / surrounding_j_i
//      (gonzag/bilin_mapping.py:595-608, 453-486, 329-449, 215-262, 194-212)
//   K3 restates BilinTrack.wght / WeightBL / AlfaBeta (gonzag/bilin_mapping.py:610-617, 490-527, 50-141)
// The per-point functions are templates on the index type IX (gzb_point.cuh): `int` for grids of fewer than 2^31 cells.
#include "gzb_common.cuh"
#include "gzb_point.cuh"

// gonzag/bilin_mapping.py:194-212; -1 flags a missing neighbour
template <typename IX>
__device__ __forceinline__ void neighbours(const GzbGrid& g, IX jP, IX iP, IX& jm, IX& jp, IX& im, IX& ip) {
    const IX Nj = (IX)g.Nj, Ni = (IX)g.Ni;
    jm = jP - 1;
    jp = jP + 1;
    im = iP - 1;
    ip = iP + 1;
    if (jp == Nj) jp = -1;
    if (im == -1) {
        if (g.kew >= 0) im = Ni - g.kew - 1;
    }
    if (ip == Ni) {
        ip = -1;
        if (g.kew >= 0) ip = g.kew;
    }
}

// gonzag/bilin_mapping.py:241-260: the three other corners, counter-clockwise from P
template <typename IX>
__device__ __forceinline__ void quad_corners(int iq, IX jP, IX iP, IX jm, IX jp, IX im, IX ip, IX* cj, IX* ci) {
    cj[0] = jP; ci[0] = iP;
    if (iq == 1)      { cj[1] = jP; ci[1] = ip; cj[2] = jp; ci[2] = ip; cj[3] = jp; ci[3] = iP; }
    else if (iq == 2) { cj[1] = jm; ci[1] = iP; cj[2] = jm; ci[2] = ip; cj[3] = jP; ci[3] = ip; }
    else if (iq == 3) { cj[1] = jP; ci[1] = im; cj[2] = jm; ci[2] = im; cj[3] = jm; ci[3] = iP; }
    else              { cj[1] = jp; ci[1] = iP; cj[2] = jp; ci[2] = im; cj[3] = jP; ci[3] = im; }
}

// planar strict point-in-quadrilateral (boundary = outside): the stand-in for shapely's
// Polygon.contains at gonzag/bilin_mapping.py:421-423, identical to
// oracle/gz_oracle.py:point_strictly_in_quad (products and one subtraction, no FMA).
__device__ __forceinline__ bool strictly_in_quad(double py, double px, const double* qy, const double* qx) {
    int wn = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double ay = qy[k], ax = qx[k];
        const double by = qy[(k + 1) & 3], bx = qx[(k + 1) & 3];
        const double side = (by - ay) * (px - ax) - (bx - ax) * (py - ay);
        if (side == 0.0) {
            if (fmin(ay, by) <= py && py <= fmax(ay, by) && fmin(ax, bx) <= px && px <= fmax(ax, bx)) return false;
        }
        if (ax <= px) {
            if (bx > px && side > 0.0) ++wn;
        } else {
            if (bx <= px && side < 0.0) --wn;
        }
    }
    return wn != 0;
}

// coordinates of cell (j, i)
template <typename IX>
__device__ __forceinline__ double2 cell_ll(const GzbGrid& g, const IX j, const IX i) {
    return GZB_LDS_LDG(g.ll + (j * (IX)g.Ni + i));
}

// the grid-only part of Iquadran's light test (gonzag/bilin_mapping.py:377-397): Mercator ordinate
// and longitude of P, loxodromic headings from P to its four neighbours, hN already brought below hE
template <typename IX>
__device__ __forceinline__ void cell_frame(const GzbGrid& g, const IX jP, const IX iP, const IX jm, const IX jp, const IX im,
                                           const IX ip, double& y0, double& x0, double& hN, double& hE,
                                           double& hS, double& hW) {
    GZB_CHK(g, jm >= 0 && jp < g.Nj && im >= 0 && ip < g.Ni && jp >= 0 && ip >= 0, 64);
    const double2 cP = cell_ll<IX>(g, jP, iP), cN = cell_ll<IX>(g, jp, iP);
    const double2 cE = cell_ll<IX>(g, jP, ip), cS = cell_ll<IX>(g, jm, iP);
    const double2 cW = cell_ll<IX>(g, jP, im);
    const double lat0 = cP.x, lonP = cP.y, latN = cN.x, lonN = cN.y, latE = cE.x, lonE = cE.y;
    const double latS = cS.x, lonS = cS.y, latW = cW.x, lonW = cW.y;
    y0 = gzb_merc_y(lat0);
    x0 = gzb_pymod(lonP, 360.0) * GZB_D2R;
    hN = gzb_heading_y(y0, x0, gzb_merc_y(latN), gzb_pymod(lonN, 360.0) * GZB_D2R);
    hE = gzb_heading_y(y0, x0, gzb_merc_y(latE), gzb_pymod(lonE, 360.0) * GZB_D2R);
    hS = gzb_heading_y(y0, x0, gzb_merc_y(latS), gzb_pymod(lonS, 360.0) * GZB_D2R);
    hW = gzb_heading_y(y0, x0, gzb_merc_y(latW), gzb_pymod(lonW, 360.0) * GZB_D2R);
    if (hN > hE) hN = hN - 360.0;
}

// can a quadrant be looked for around (jP, iP)?  (all four neighbours exist)
template <typename IX>
__device__ __forceinline__ bool mesh_usable(const GzbGrid& g, const IX jP, const IX iP, IX& jm, IX& jp, IX& im, IX& ip) {
    const IX Nj = (IX)g.Nj, Ni = (IX)g.Ni;
    if (!(jP >= 0 && jP < Nj && iP >= 0 && iP < Ni)) return false;
    neighbours<IX>(g, jP, iP, jm, jp, im, ip);
    if (jm == -1 || jp == -1 || im == -1 || ip == -1) return false;
    if (im < 0 || jm < 0 || ip > Ni - 1) return false;
    return true;
}

// source mesh of one point: corners P1(nearest)..P4 counter-clockwise, all -1 when no quadrant
// is found, all 0 when the nearest point itself was not found
template <typename IX>
__device__ __forceinline__ void source_mesh_one(const GzbGrid& g, const double yT, const double xT,
                                                const IX jP, const IX iP, const int force_heavy, IX* cj, IX* ci) {
    if (jP == -1 && iP == -1) {   // nearest point not found: the row stays zero
#pragma unroll
        for (int k = 0; k < 4; ++k) { cj[k] = 0; ci[k] = 0; }
        return;
    }
    int iq = -1;
    IX jm = 0, jp = 0, im = 0, ip = 0;
    const bool usable = mesh_usable<IX>(g, jP, iP, jm, jp, im, ip);
    if (usable) {
        // five headings from P (to the target and to the N, E, S, W neighbours): the Mercator
        // ordinate of P is common to all of them; the four that depend on the grid only come from
        // the per-cell table when it has been built (same code, same bits)
        double y0, x0, hN, hE, hS, hW;
        const IX cP = jP * (IX)g.Ni + iP;
        if (g.hdg) {
            const double2* __restrict__ h = reinterpret_cast<const double2*>(g.hdg + (long long)GZB_EXP_HDG_STRIDE * cP);
            const double2 a = GZB_LDS_REF(h), b = GZB_LDS_REF(h + 1), c = GZB_LDS_REF(h + 2);
            y0 = a.x; x0 = a.y; hN = b.x; hE = b.y; hS = c.x; hW = c.y;
        } else {
            cell_frame<IX>(g, jP, iP, jm, jp, im, ip, y0, x0, hN, hE, hS, hW);
        }
        const double ht = gzb_heading_y(y0, x0, gzb_merc_y(yT), gzb_pymod(xT, 360.0) * GZB_D2R);
        const double hz = gzb_pm180(ht);
        if (force_heavy < 2) {
            if (hz > hN && hz <= hE) iq = 1;
            if (ht > hE && ht <= hS) iq = 2;
            if (ht > hS && ht <= hW) iq = 3;
            if (ht > hW && hz <= hN) iq = 4;
        }
        const double ang = g.angle ? GZB_LDS_REF(g.angle + cP) : 0.0;
        if (iq < 1 || fabs(ang) > 45.0 || force_heavy) {
            // heavy-duty: quads 4,3,2,1 in the raw (lat, lon) plane, first strict-interior hit
            for (int cand = 4; cand >= 1; --cand) {
                quad_corners<IX>(cand, jP, iP, jm, jp, im, ip, cj, ci);
                double qy[4], qx[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    GZB_CHK(g, cj[k] >= 0 && cj[k] < g.Nj && ci[k] >= 0 && ci[k] < g.Ni, 64);
                    const double2 c = cell_ll<IX>(g, cj[k], ci[k]);
                    qy[k] = c.x;
                    qx[k] = c.y;
                }
                if (strictly_in_quad(yT, xT, qy, qx)) {
                    iq = cand;
                    break;
                }
            }
        }
    }
    if (iq >= 1) {
        quad_corners<IX>(iq, jP, iP, jm, jp, im, ip, cj, ci);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { cj[k] = -1; ci[k] = -1; }
    }
}

// a nearest-point index read from a caller's int64 array, as an IX: what does not fit 32 bits becomes -2, which is outside
// every grid like the original value (only the pair (-1, -1) means "not found")
template <typename IX> __device__ __forceinline__ IX narrow_np(long long v);
template <> __device__ __forceinline__ long long narrow_np<long long>(long long v) { return v; }
template <> __device__ __forceinline__ int narrow_np<int>(long long v) { return ((long long)(int)v == v) ? (int)v : -2; }

// nearest point of track point t for the bulk kernels: from the NP array, or (gzb_mapping overlap, while
// the chain logic still runs) E(t) derived from the plain global answer -- gonzag/bilin_mapping.py:169-190
// threshold acceptance + :582-585 edge exclusion, as accept_edge() of gzb_nearest.cu
template <typename IX>
__device__ __forceinline__ bool nearest_of(const GzbGrid& g, const long long* __restrict__ NP, const GzbGlobalNearest& gn,
                                           const long long t, IX& jP, IX& iP) {
    if (NP) {
        const longlong2 v = reinterpret_cast<const longlong2*>(NP)[t];
        jP = narrow_np<IX>(v.x);
        iP = narrow_np<IX>(v.y);
        return true;
    }
    const long long flat = gn.G[t];
    if (flat == 0x7ffffffffffffffeLL) return false;      // not settled yet: left to the repair kernel
    const double d = gn.dG[t];
    if (flat != 0x7fffffffffffffffLL && d < gn.thr2) {
        if (sizeof(IX) == 4) {                           // flat < Nj * Ni < 2^31
            const unsigned f = (unsigned)flat, n = (unsigned)g.Ni, q = f / n;
            jP = (IX)q;
            iP = (IX)(f - q * n);
        } else {
            jP = (IX)(flat / g.Ni);
            iP = (IX)(flat - (long long)jP * g.Ni);
        }
    } else {
        jP = -1;
        iP = -1;
    }
    if (gn.edge) {
        if (iP == 0 || (iP == (IX)g.Ni - 1 && g.kew == -1)) { jP = -1; iP = -1; }
        if (jP == 0 || jP == (IX)g.Nj - 1) { jP = -1; iP = -1; }
    }
    return true;
}

// bilinear weights of one point inside its source mesh (gonzag/bilin_mapping.py:490-527)
template <typename IX>
__device__ __forceinline__ void weights_one(const GzbGrid& g, const double yT, const double xT,
                                            const IX* cj, const IX* ci, double* w) {
    w[0] = 1.0; w[1] = 0.0; w[2] = 0.0; w[3] = 0.0;
    if (cj[0] >= 0) {
        double vy[5], vx[5];
        vy[0] = yT;
        vx[0] = xT;
#pragma unroll
        for (int k = 0; k < 4; ++k) {      // Python negative indexing, then clamped (never faults)
            const double2 c = cell_ll<IX>(g, pywrap_ix<IX>(cj[k], (IX)g.Nj), pywrap_ix<IX>(ci[k], (IX)g.Ni));
            vy[k + 1] = c.x;
            vx[k + 1] = c.y;
        }
        double a, b;
        gzb_alfabeta_dev(vy, vx, a, b);
        w[0] = (1.0 - a) * (1.0 - b);
        w[1] = a * (1.0 - b);
        w[2] = a * b;
        w[3] = (1.0 - a) * b;
    }
}

// does this grid take the 32-bit index variants?
static inline bool gzb_ix32(const gzb_ctx* ctx) { return ctx->g.Nj * ctx->g.Ni < 0x7fffffffLL && !ctx->force_ix64; }

// the mesh of a point as the reference's int64 rows (SM[t, k] = [j, i])
template <typename IX>
__device__ __forceinline__ void store_mesh(long long* __restrict__ SM, const long long t, const IX* cj, const IX* ci) {
    longlong2* out = reinterpret_cast<longlong2*>(SM + 8 * t);
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = make_longlong2((long long)cj[k], (long long)ci[k]);
}

template <typename IX>
__global__ void __launch_bounds__(128)
k_source_mesh(const GzbGrid g, const long long nt, const double* __restrict__ yt, const double* __restrict__ xt,
              const long long* __restrict__ NP, long long* __restrict__ SM, const int force_heavy) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    IX cj[4], ci[4];
    source_mesh_one<IX>(g, yt[t], xt[t], narrow_np<IX>(NP[2 * t]), narrow_np<IX>(NP[2 * t + 1]), force_heavy, cj, ci);
    store_mesh<IX>(SM, t, cj, ci);
}

// K3 alone reads a caller's SM: an entry that does not fit the index type lies far outside the grid; it keeps its sign
// (a negative first row means "no mesh") and clamps like the 64-bit rule in weights_one
template <typename IX> __device__ __forceinline__ IX narrow_sm(long long v);
template <> __device__ __forceinline__ long long narrow_sm<long long>(long long v) { return v; }
template <> __device__ __forceinline__ int narrow_sm<int>(long long v) {
    return ((long long)(int)v == v) ? (int)v : (v < 0 ? (int)0x80000000 : 0x7fffffff);
}

template <typename IX>
__global__ void __launch_bounds__(128)
k_weights(const GzbGrid g, const long long nt, const double* __restrict__ yt, const double* __restrict__ xt,
          const long long* __restrict__ SM, double* __restrict__ WB) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const longlong2* sm = reinterpret_cast<const longlong2*>(SM + 8 * t);
    IX cj[4], ci[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const longlong2 p = sm[k];
        cj[k] = narrow_sm<IX>(p.x);
        ci[k] = narrow_sm<IX>(p.y);
    }
    double w[4];
    weights_one<IX>(g, yt[t], xt[t], cj, ci, w);
    double2* out = reinterpret_cast<double2*>(WB + 4 * t);
    out[0] = make_double2(w[0], w[1]);
    out[1] = make_double2(w[2], w[3]);
}

// K2 + K3 in one pass (what gzb_mapping runs): the mesh never travels through HBM between them
template <typename IX>
__global__ void __launch_bounds__(128)
k_mesh_weights(const GzbGrid g, const long long nt, const double* __restrict__ yt, const double* __restrict__ xt,
               const long long* __restrict__ NP, const GzbGlobalNearest gn, long long* __restrict__ SM,
               double* __restrict__ WB, const int force_heavy) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const double yT = yt[t], xT = xt[t];
    IX cj[4], ci[4], jP, iP;
    if (!nearest_of<IX>(g, NP, gn, t, jP, iP)) return;
    source_mesh_one<IX>(g, yT, xT, jP, iP, force_heavy, cj, ci);
    store_mesh<IX>(SM, t, cj, ci);
    double w[4];
    weights_one<IX>(g, yT, xT, cj, ci, w);
    double2* out = reinterpret_cast<double2*>(WB + 4 * t);
    out[0] = make_double2(w[0], w[1]);
    out[1] = make_double2(w[2], w[3]);
}

// per-cell table of cell_frame(): thread per cell
__global__ void __launch_bounds__(128)
k_heading_table(const GzbGrid g, double* __restrict__ hdg) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.Nj * g.Ni) return;
    const long long jP = k / g.Ni, iP = k - jP * g.Ni;
    long long jm = 0, jp = 0, im = 0, ip = 0;
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (mesh_usable<long long>(g, jP, iP, jm, jp, im, ip)) cell_frame<long long>(g, jP, iP, jm, jp, im, ip, v[0], v[1], v[2], v[3], v[4], v[5]);
    double2* out = reinterpret_cast<double2*>(hdg + GZB_EXP_HDG_STRIDE * k);
    out[0] = make_double2(v[0], v[1]);
    out[1] = make_double2(v[2], v[3]);
    out[2] = make_double2(v[4], v[5]);
}

// The table costs one cell_frame() per grid cell; a K2 call costs one per track point.  Build it
// once the points mapped on this grid outnumber half its cells (option "heading_table": 0 never,
// 1 this rule, 2 at the next call).  Out of memory is not an error: K2 computes the frames itself.
static int maybe_build_heading_table(gzb_ctx* ctx, int64_t nt) {
    ctx->k2_points += nt;
    if (ctx->hdg_built || ctx->hdg_mode == 0) return GZB_OK;
    const int64_t cells = ctx->g.Nj * ctx->g.Ni;
    if (ctx->hdg_mode == 1 && ctx->k2_points < cells / 2) return GZB_OK;
    if (!ctx->d_hdg) {
        cudaError_t e = gzb_pool_malloc((void**)&ctx->d_hdg, (size_t)cells * GZB_EXP_HDG_STRIDE * sizeof(double));
        if (e != cudaSuccess) {
            cudaGetLastError();
            ctx->d_hdg = nullptr;
            ctx->hdg_mode = 0;
            return GZB_OK;
        }
    }
    ctx->g.hdg = nullptr;
    k_heading_table<<<(unsigned)((cells + 127) / 128), 128, 0, ctx->stream>>>(ctx->g, ctx->d_hdg);
    GZB_LAUNCH_CHECK(ctx);
    ctx->g.hdg = ctx->d_hdg;
    ctx->hdg_built = 1;
    ctx->stats.heading_table = 1;
    return GZB_OK;
}

int gzb_mesh_weights_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* np,
                         int64_t* sm_out, double* wb_out, const GzbGlobalNearest* gn_or_null) {
    if (nt <= 0) return GZB_OK;
    if (maybe_build_heading_table(ctx, nt) != GZB_OK) return GZB_ERR_CUDA;
    GzbGlobalNearest gn = {nullptr, nullptr, 0.0, 0, nullptr, nullptr};
    const long long* npp = (const long long*)np;
    if (gn_or_null) { gn = *gn_or_null; npp = nullptr; }
    if (gzb_ix32(ctx))
        k_mesh_weights<int><<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(ctx->g, nt, yt, xt, npp, gn, (long long*)sm_out,
                                                                                  wb_out, ctx->force_heavy);
    else
        k_mesh_weights<long long><<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(ctx->g, nt, yt, xt, npp, gn, (long long*)sm_out,
                                                                                        wb_out, ctx->force_heavy);
    GZB_LAUNCH_CHECK(ctx);
    return GZB_OK;
}

// K2 + K3 (+ K4) again for the listed points: the ones whose nearest point the chain replays
// changed after the bulk kernels had used the speculative value.  The list length lives on the
// device.  One kernel for the whole repair: it sits on the critical path behind the replays.
struct InterpArgs {            // K4 arguments of the fused path; slab == nullptr: mapping only
    const double* t_sat;
    long long nrec;
    const double* tmod;
    const void* slab;
    int dtype;
    long long k0, nk;
    double rmiss;
    double* out_np;
    double* out_bl;
    int* err;
    const unsigned* mbits;
    const int* mflag;
    GzbPeerOut peers;
};

template <typename IX>
__global__ void __launch_bounds__(64)
k_path_fix(const GzbGrid g, const double* __restrict__ yt, const double* __restrict__ xt,
           const long long* __restrict__ NP, long long* __restrict__ SM, double* __restrict__ WB,
           const long long* __restrict__ list, const unsigned long long* __restrict__ count,
           const long long* __restrict__ list2, const unsigned long long* __restrict__ count2,
           const int force_heavy, const InterpArgs ia) {
    const long long n1 = (long long)*count;
    const long long n = n1 + (list2 ? (long long)*count2 : 0);
    GZB_KT_MIN(g, 14);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long t = e < n1 ? list[e] : list2[e - n1];
        if (!GZB_CHK(g, t >= 0, 4)) continue;
        const double yT = yt[t], xT = xt[t];
        IX cj[4], ci[4];
        source_mesh_one<IX>(g, yT, xT, narrow_np<IX>(NP[2 * t]), narrow_np<IX>(NP[2 * t + 1]), force_heavy, cj, ci);
        store_mesh<IX>(SM, t, cj, ci);
        double w[4];
        weights_one<IX>(g, yT, xT, cj, ci, w);
        double2* out = reinterpret_cast<double2*>(WB + 4 * t);
        const double2 rw[2] = {make_double2(w[0], w[1]), make_double2(w[2], w[3])};
        out[0] = rw[0];
        out[1] = rw[1];
        if (ia.slab) {      // K4 straight from the registers: the arrays just written are not read back
            if (ia.dtype == GZB_F32)
                interp_point<float, false, true, 0, IX>(g, t, ia.t_sat, SM, WB, ia.nrec, ia.tmod, (const float*)ia.slab, ia.k0,
                                                        ia.nk, ia.rmiss, 1, ia.out_np, ia.out_bl, ia.err, ia.mbits, ia.mflag, ia.peers, cj, ci, rw);
            else
                interp_point<double, false, true, 0, IX>(g, t, ia.t_sat, SM, WB, ia.nrec, ia.tmod, (const double*)ia.slab, ia.k0,
                                                         ia.nk, ia.rmiss, 1, ia.out_np, ia.out_bl, ia.err, ia.mbits, ia.mflag, ia.peers, cj, ci, rw);
        }
    }
    GZB_KT_MAX(g, 14);
}

static int maybe_build_heading_table(gzb_ctx* ctx, int64_t nt);

// K2 + K3 + K4 for every point in one pass (gzb_map_interp): the mesh and the weights go from the
// registers of K3 straight into the gather, SM / WB are written once and never read back
// Prefetch (option "prefetch"): the gather of a point is the last of four dependent DRAM round trips
// (table -> mesh corners -> [Newton solve] -> mask word -> field values).  Its addresses are known as soon as
// the mesh is: ask L2 for the sectors right then, so that they travel while the Newton iteration runs.
// The time bracket is only guessed (mean record spacing); a wrong guess fetches a neighbouring record.
template <typename T, typename IX>
__device__ __forceinline__ void prefetch_point(const GzbGrid& g, const long long t, const InterpArgs& ia,
                                               const IX* cj, const IX* ci) {
    const double now = __ldg(ia.t_sat + t);
    const double ta = __ldg(ia.tmod), tb = __ldg(ia.tmod + ia.nrec - 1);
    if (!(now >= ta) || !(now < tb)) return;
    long long lo = (long long)((now - ta) / (tb - ta) * (double)(ia.nrec - 1));
    lo = lo < 0 ? 0 : (lo > ia.nrec - 2 ? ia.nrec - 2 : lo);
    if (lo < ia.k0 || lo + 1 >= ia.k0 + ia.nk) return;
    const long long cells = g.Nj * g.Ni;
    const T* X1 = (const T*)ia.slab + (lo - ia.k0) * cells;
    const long long ntI = (g.Ni + 15) >> 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long j = pywrap_idx((long long)cj[k], g.Nj), i = pywrap_idx((long long)ci[k], g.Ni);
        const long long c = j * g.Ni + i;
        gzb_l2_prefetch(X1 + c);
        gzb_l2_prefetch(X1 + cells + c);
        if (ia.mbits && (k == 0 || k == 2)) gzb_l2_prefetch(ia.mbits + ((j >> 4) * ntI + (i >> 4)) * 8);
    }
}

// Early prefetch (option "prefetch" = 4): everything a point will gather lies in the 3x3 index neighbourhood of its
// nearest point -- the heading-table row and the angle of the cell, the coordinates of the mesh corners, the mask tile, and
// in both bracketing records the field values -- but the kernel reaches them one dependent round trip after the other
// (table -> quadrant -> corners -> Newton -> mask -> field).  Ask L2 for all of it the moment the nearest point is known:
// the later loads then hit L2.  Costs the third (unused) row of the field and of the coordinates in DRAM traffic.
template <typename T>
__device__ __forceinline__ void prefetch_hood(const GzbGrid& g, const long long t, const InterpArgs& ia, const long long jP,
                                              const long long iP) {
    if (jP < 0 || iP < 0 || jP >= g.Nj || iP >= g.Ni) return;
    const long long c = jP * g.Ni + iP;
    const long long ja = jP > 0 ? jP - 1 : jP, jb = jP + 1 < g.Nj ? jP + 1 : jP;
    const long long ia0 = iP > 0 ? iP - 1 : iP;
    if (g.hdg) {
        const double* h = g.hdg + GZB_EXP_HDG_STRIDE * c;
        gzb_l2_prefetch(h);
        gzb_l2_prefetch(h + 5);
    }
    if (g.angle) gzb_l2_prefetch(g.angle + c);
    gzb_l2_prefetch(g.ll + ja * g.Ni + ia0);
    gzb_l2_prefetch(g.ll + jP * g.Ni + ia0);
    gzb_l2_prefetch(g.ll + jb * g.Ni + ia0);
    if (ia.mbits) gzb_l2_prefetch(ia.mbits + ((jP >> 4) * ((g.Ni + 15) >> 4) + (iP >> 4)) * 8);
    const double now = __ldg(ia.t_sat + t);
    const double ta = __ldg(ia.tmod), tb = __ldg(ia.tmod + ia.nrec - 1);
    if (!(now >= ta) || !(now < tb)) return;
    long long lo = (long long)((now - ta) / (tb - ta) * (double)(ia.nrec - 1));
    lo = lo < 0 ? 0 : (lo > ia.nrec - 2 ? ia.nrec - 2 : lo);
    if (lo < ia.k0 || lo + 1 >= ia.k0 + ia.nk) return;
    const long long cells = g.Nj * g.Ni;
    const T* X1 = (const T*)ia.slab + (lo - ia.k0) * cells;
    gzb_l2_prefetch(X1 + ja * g.Ni + iP); gzb_l2_prefetch(X1 + cells + ja * g.Ni + iP);
    gzb_l2_prefetch(X1 + c);              gzb_l2_prefetch(X1 + cells + c);
    gzb_l2_prefetch(X1 + jb * g.Ni + iP); gzb_l2_prefetch(X1 + cells + jb * g.Ni + iP);
}

// Resident blocks of 128 threads per SM (7 = a cap of 72 registers; uncapped the kernel takes 91).  Measured on the C3 track,
// 32-bit indices: 5 blocks 90.1 us, 6 blocks 84.1, 7 blocks 82.0, 8 blocks 80.6 before the grid-stride form of the loop
// (gpurun_out/ab_exp*_r2n.log), 86.3 / 81.4 / 76.0 / 76.0 with it (ab_exp*_r2w.log) -- the kernel is bound by the latency of
// its dependent gathers, so warps in flight beat the spills they cost.  7, not 8: alone the two are equal, inside a step
// 7 is 1-2 % faster (197 vs 199.6 us; 6 blocks: 195 us, but the kernel alone then takes 81 us).
#ifndef GZB_EXP_BULK_BLOCKS
#define GZB_EXP_BULK_BLOCKS 7
#endif
template <typename T, int PF, typename IX>      // PF: 0 plain, 1 L2 prefetch of the gather, 3 gather with 64-byte L2 fetches, 4 early prefetch
__global__ void __launch_bounds__(128, GZB_EXP_BULK_BLOCKS)
k_mesh_weights_interp(const GzbGrid g, const long long t_first, const long long nt, const double* __restrict__ yt,
                      const double* __restrict__ xt, const long long* __restrict__ NP, const GzbGlobalNearest gn,
                      long long* __restrict__ SM, double* __restrict__ WB, const int force_heavy, const InterpArgs ia) {
    GZB_KT_MIN(g, 13);
    // this launch: points [t_first, nt), one block per 128 points -- or (option "bulk_resident") a few persistent blocks per
    // SM striding over them, so that the chain kernels of K1, which run beside this one, always find room on every SM
    for (long long t = t_first + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (long long)gridDim.x * blockDim.x) {
        const double yT = yt[t], xT = xt[t];
        IX cj[4], ci[4], jP, iP;
        if (!nearest_of<IX>(g, NP, gn, t, jP, iP)) continue;
        if (PF == 4) prefetch_hood<T>(g, t, ia, (long long)jP, (long long)iP);
        source_mesh_one<IX>(g, yT, xT, jP, iP, force_heavy, cj, ci);
        if (PF == 1) prefetch_point<T, IX>(g, t, ia, cj, ci);
        store_mesh<IX>(SM, t, cj, ci);
        double w[4];
        weights_one<IX>(g, yT, xT, cj, ci, w);
        double2* out = reinterpret_cast<double2*>(WB + 4 * t);
        const double2 rw[2] = {make_double2(w[0], w[1]), make_double2(w[2], w[3])};
        out[0] = rw[0];
        out[1] = rw[1];
        interp_point<T, false, true, (PF >= 3 ? 3 : 0), IX>(g, t, ia.t_sat, SM, WB, ia.nrec, ia.tmod, (const T*)ia.slab, ia.k0, ia.nk, ia.rmiss, 1,
                                     ia.out_np, ia.out_bl, ia.err, ia.mbits, ia.mflag, ia.peers, cj, ci, rw);
    }
    GZB_KT_MAX(g, 13);
}

template <typename IX> static void carveout_ix(int v) {
#define GZB_CV(K) cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, v)
    GZB_CV((k_mesh_weights_interp<float, 0, IX>)); GZB_CV((k_mesh_weights_interp<float, 1, IX>));
    GZB_CV((k_mesh_weights_interp<float, 3, IX>)); GZB_CV((k_mesh_weights_interp<float, 4, IX>));
    GZB_CV((k_mesh_weights_interp<double, 0, IX>)); GZB_CV((k_mesh_weights_interp<double, 1, IX>));
    GZB_CV((k_mesh_weights_interp<double, 3, IX>)); GZB_CV((k_mesh_weights_interp<double, 4, IX>));
    GZB_CV(k_mesh_weights<IX>); GZB_CV(k_path_fix<IX>);
#undef GZB_CV
}

void gzb_carveout_mesh(int pct) {
    const int v = pct < 0 ? (int)cudaSharedmemCarveoutDefault : pct;
    carveout_ix<int>(v);
    carveout_ix<long long>(v);
}

static void fill_interp_args(gzb_ctx* ctx, InterpArgs& ia, const double* tt, int64_t nrec, const double* tmod,
                             const void* slab_dev, int dtype, int64_t k0, int64_t nk, double rmiss, double* vnp_out,
                             double* vbl_out) {
    ia.t_sat = tt; ia.nrec = nrec; ia.tmod = tmod; ia.slab = slab_dev; ia.dtype = dtype; ia.k0 = k0; ia.nk = nk;
    ia.rmiss = rmiss; ia.out_np = vnp_out; ia.out_bl = vbl_out; ia.err = ctx->d_flags;
    ia.mbits = ctx->mask_bits_ok ? ctx->d_mbits : nullptr;
    ia.mflag = ctx->d_flags + 2;
    ia.peers.n = 0;
    if (ctx->peer_out.n > 0 && vnp_out == ctx->peer_local_np && vbl_out == ctx->peer_local_bl) ia.peers = ctx->peer_out;
}

// The two series of points [a, b), just written by a bulk launch, into every registered peer buffer: the gather of a
// sharded step, issued chunk by chunk on its own stream so that the NVLink stores of chunk c travel while the bulk kernel
// works on chunk c+1 (the same stores issued from inside the bulk kernel throttle it: +120 us per step at 8 GPUs).
__global__ void __launch_bounds__(256)
k_push(const GzbGrid g, const long long nt, const double* __restrict__ np_loc, const double* __restrict__ bl_loc,
       const GzbPeerOut peers, const long long a, const long long b) {
    for (long long t = a + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < b; t += (long long)gridDim.x * blockDim.x) {
        if (!GZB_CHK(g, t >= 0 && t < nt, 128)) continue;
        const double v = np_loc[t], w = bl_loc[t];
        for (int k = 0; k < peers.n; ++k) {
            peers.np[k][t] = v;
            peers.bl[k][t] = w;
        }
    }
}

// ... and the points the repair kernel rewrote afterwards (two device-side lists)
__global__ void __launch_bounds__(128)
k_push_list(const double* __restrict__ np_loc, const double* __restrict__ bl_loc, const GzbPeerOut peers,
            const long long* __restrict__ list, const unsigned long long* __restrict__ count,
            const long long* __restrict__ list2, const unsigned long long* __restrict__ count2) {
    const long long n1 = (long long)*count;
    const long long n = n1 + (list2 ? (long long)*count2 : 0);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long t = e < n1 ? list[e] : list2[e - n1];
        const double v = np_loc[t], w = bl_loc[t];
        for (int k = 0; k < peers.n; ++k) {
            peers.np[k][t] = v;
            peers.bl[k][t] = w;
        }
    }
}

int gzb_push_join(gzb_ctx* ctx);

int gzb_mesh_weights_interp_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* np,
                                const GzbGlobalNearest* gn_or_null, int64_t* sm_out, double* wb_out, const double* tt, int64_t nrec, const double* tmod,
                                const void* slab_dev, int dtype, int64_t k0, int64_t nk, double rmiss, double* vnp_out,
                                double* vbl_out) {
    if (nt <= 0) return GZB_OK;
    if (maybe_build_heading_table(ctx, nt) != GZB_OK) return GZB_ERR_CUDA;
    InterpArgs ia;
    fill_interp_args(ctx, ia, tt, nrec, tmod, slab_dev, dtype, k0, nk, rmiss, vnp_out, vbl_out);
    const GzbPeerOut peers = ia.peers;
    // with peer outputs the bulk kernel writes locally only; k_push copies the finished series into the peers' buffers on
    // its own stream.  Option "bulk_chunks" > 1 cuts the kernel into chunks so that the copies of chunk c travel while chunk
    // c+1 is computed -- measured at 8 GPUs (gpurun_out/abm_n8_r2t.err) that overlap is a loss: the peers' stores arriving
    // in this GPU's memory (71 MB per step, ~480 GB/s of ingress) slow the latency-bound bulk kernel down by a factor 2
    // (its end moves from 169 to 308 us), so the default is ONE chunk: 0.361 -> 0.338 ms per step.
    int nchunk = ctx->bulk_chunks > 0 ? ctx->bulk_chunks : 1;
    if (nt < 4096 * nchunk) nchunk = 1;
    const bool push = peers.n > 0;
    if (push) ia.peers.n = 0;
    GzbGlobalNearest gn = {nullptr, nullptr, 0.0, 0, nullptr, nullptr};
    const long long* npp = (const long long*)np;
    if (gn_or_null) { gn = *gn_or_null; npp = nullptr; }
    const int pfm = !ctx->pf_now ? 0 : (ctx->prefetch >= 3 ? ctx->prefetch : 1);
    cudaStream_t s = ctx->stream;
    for (int c = 0; c < nchunk; ++c) {
        const long long a = ((nt * c / nchunk) + 127) & ~127LL, b = c + 1 == nchunk ? nt : (((nt * (c + 1) / nchunk) + 127) & ~127LL);
        if (b <= a) continue;
        unsigned nblk = (unsigned)((b - a + 127) / 128);
        if (ctx->bulk_resident > 0) {
            int nsm = 148;
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
            const unsigned cap = (unsigned)(nsm * ctx->bulk_resident);
            if (nblk > cap) nblk = cap;
        }
#define GZB_BULK(T, PF, IX) k_mesh_weights_interp<T, PF, IX><<<nblk, 128, 0, s>>>(ctx->g, a, b, yt, xt, npp, gn, (long long*)sm_out, \
                                                                                wb_out, ctx->force_heavy, ia)
#define GZB_BULK_PF(T, IX) { if (pfm == 4) GZB_BULK(T, 4, IX); else if (pfm == 3) GZB_BULK(T, 3, IX); else if (pfm) GZB_BULK(T, 1, IX); else GZB_BULK(T, 0, IX); }
        if (gzb_ix32(ctx)) { if (dtype == GZB_F32) GZB_BULK_PF(float, int) else GZB_BULK_PF(double, int) }
        else               { if (dtype == GZB_F32) GZB_BULK_PF(float, long long) else GZB_BULK_PF(double, long long) }
#undef GZB_BULK_PF
#undef GZB_BULK
        GZB_LAUNCH_CHECK(ctx);
        if (push && ctx->push_mode == 0) {
            GZB_CUDA(ctx, cudaEventRecord(ctx->ev_push[c], s));
            GZB_CUDA(ctx, cudaStreamWaitEvent(ctx->push_stream, ctx->ev_push[c], 0));
            k_push<<<(unsigned)ctx->push_blocks, 256, 0, ctx->push_stream>>>(ctx->g, nt, vnp_out, vbl_out, peers, a, b);
            GZB_LAUNCH_CHECK(ctx);
            gzb_trace(ctx, ctx->push_stream, "push of a chunk end (push stream)");
        } else if (push) {
            // copy engines: one stream per peer, two copies (the two series) per chunk on each -- the SMs issue nothing
            GZB_CUDA(ctx, cudaEventRecord(ctx->ev_push[c], s));
            for (int k = 0; k < peers.n; ++k) {
                if (!ctx->peer_stream[k]) {
                    GZB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->peer_stream[k], cudaStreamNonBlocking));
                    GZB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_peer[k], cudaEventDisableTiming));
                }
                GZB_CUDA(ctx, cudaStreamWaitEvent(ctx->peer_stream[k], ctx->ev_push[c], 0));
                GZB_CUDA(ctx, cudaMemcpyAsync(peers.np[k] + a, vnp_out + a, (size_t)(b - a) * 8, cudaMemcpyDeviceToDevice, ctx->peer_stream[k]));
                GZB_CUDA(ctx, cudaMemcpyAsync(peers.bl[k] + a, vbl_out + a, (size_t)(b - a) * 8, cudaMemcpyDeviceToDevice, ctx->peer_stream[k]));
            }
        }
    }
    if (push && ctx->push_mode == 0) {      // the caller joins the copies (gzb_push_join): the repair kernel runs beside the last of them
        GZB_CUDA(ctx, cudaEventRecord(ctx->ev_push[8], ctx->push_stream));
        ctx->push_pending = 1;
    } else if (push) {
        for (int k = 0; k < peers.n; ++k) GZB_CUDA(ctx, cudaEventRecord(ctx->ev_peer[k], ctx->peer_stream[k]));
        ctx->peer_streams_used = peers.n;
        ctx->push_pending = 2;
    }
    ctx->time_err_pending = 1;
    return GZB_OK;
}

// mapping only (gzb_mapping) when slab_dev == nullptr, mapping + interpolation (gzb_map_interp) otherwise
int gzb_path_fix_dev(gzb_ctx* ctx, const double* yt, const double* xt, const int64_t* np, int64_t* sm_out,
                     double* wb_out, const long long* list, const unsigned long long* count, const long long* list2,
                     const unsigned long long* count2, const double* tt,
                     int64_t nrec, const double* tmod, const void* slab_dev, int dtype, int64_t k0, int64_t nk,
                     double rmiss, double* vnp_out, double* vbl_out) {
    InterpArgs ia;
    fill_interp_args(ctx, ia, tt, nrec, tmod, slab_dev, dtype, k0, nk, rmiss, vnp_out, vbl_out);
    const GzbPeerOut peers = ia.peers;
    const bool joined = ctx->push_pending && peers.n > 0;
    if (joined) ia.peers.n = 0;       // chunk copies in flight: repair locally beside them, then copy the repaired points
    if (gzb_ix32(ctx))
        k_path_fix<int><<<192, 64, 0, ctx->stream>>>(ctx->g, yt, xt, (const long long*)np, (long long*)sm_out, wb_out, list, count,
                                                    list2, count2, ctx->force_heavy, ia);
    else
        k_path_fix<long long><<<192, 64, 0, ctx->stream>>>(ctx->g, yt, xt, (const long long*)np, (long long*)sm_out, wb_out, list, count,
                                                          list2, count2, ctx->force_heavy, ia);
    GZB_LAUNCH_CHECK(ctx);
    if (joined) {
        if (gzb_push_join(ctx) != GZB_OK) return GZB_ERR_CUDA;
        gzb_trace(ctx, ctx->stream, "repair kernel end, pushes joined (main)");
        k_push_list<<<148, 128, 0, ctx->stream>>>(vnp_out, vbl_out, peers, list, count, list2, count2);      // ~10 k points: one pass
        GZB_LAUNCH_CHECK(ctx);
    }
    if (slab_dev) ctx->time_err_pending = 1;
    return GZB_OK;
}

// the main stream waits for the chunk copies of the last bulk call (no-op when there are none)
int gzb_push_join(gzb_ctx* ctx) {
    if (!ctx->push_pending) return GZB_OK;
    const int mode = ctx->push_pending;
    ctx->push_pending = 0;
    if (mode == 1) {
        GZB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_push[8], 0));
    } else {
        for (int k = 0; k < ctx->peer_streams_used; ++k) GZB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_peer[k], 0));
    }
    return GZB_OK;
}

int gzb_source_mesh_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* np,
                        int64_t* sm_out) {
    if (nt <= 0) return GZB_OK;
    if (maybe_build_heading_table(ctx, nt) != GZB_OK) return GZB_ERR_CUDA;
    if (gzb_ix32(ctx))
        k_source_mesh<int><<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(ctx->g, nt, yt, xt, (const long long*)np,
                                                                                 (long long*)sm_out, ctx->force_heavy);
    else
        k_source_mesh<long long><<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(ctx->g, nt, yt, xt, (const long long*)np,
                                                                                       (long long*)sm_out, ctx->force_heavy);
    GZB_LAUNCH_CHECK(ctx);
    return GZB_OK;
}

int gzb_weights_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, const int64_t* sm,
                    double* wb_out) {
    if (nt <= 0) return GZB_OK;
    if (gzb_ix32(ctx))
        k_weights<int><<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(ctx->g, nt, yt, xt, (const long long*)sm, wb_out);
    else
        k_weights<long long><<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(ctx->g, nt, yt, xt, (const long long*)sm, wb_out);
    GZB_LAUNCH_CHECK(ctx);
    return GZB_OK;
}

// ------------------------------------------------------------------------------------------
// batched scalar helpers (Heading, AlfaBeta, Haversine) for the free functions the reference
// package re-exports (gonzag/__init__.py:13)
__global__ void k_heading(long long n, const double* a, const double* b, const double* c, const double* d, double* o) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) o[k] = gzb_heading_dev(a[k], b[k], c[k], d[k]);
}

__global__ void k_alfabeta(long long n, const double* vy, const double* vx, double* ab) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        double a, b;
        gzb_alfabeta_dev(vy + 5 * k, vx + 5 * k, a, b);
        ab[2 * k] = a;
        ab[2 * k + 1] = b;
    }
}

__global__ void k_haversine(long long n, double plat, double plon, const double* la, const double* lo, double* o) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) o[k] = gzb_haversine_dev(plat, plon, cos(plat * GZB_D2R), la[k], lo[k]);
}

// Haversine_sclr(lat1,lon1,lat2,lon2) and RadiusEarth(lat) (gonzag/utils.py:267-293): the device functions of the
// along-track distance kernels, one problem per thread
__global__ void k_haversine_sclr(long long n, const double* a, const double* b, const double* c, const double* d, double* o) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) o[k] = haversine_sclr_dev(a[k], b[k], c[k], d[k]);
}

__global__ void k_radius_earth(long long n, const double* lat, double* o) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) o[k] = radius_earth_dev(lat[k]);
}

namespace {
struct TmpDev {
    void* p = nullptr;
    ~TmpDev() { if (p) gzb_pool_free(p); }
};
}  // namespace

static int helper_run(int device, size_t in_doubles, size_t out_doubles, const double* const* ins,
                      const size_t* in_sizes, int nin, double* out, void (*launch)(double** din, double* dout)) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return gzb_fail(nullptr, GZB_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    TmpDev buf;
    e = gzb_pool_malloc(&buf.p, (in_doubles + out_doubles + 8) * sizeof(double));
    if (e != cudaSuccess) return gzb_fail(nullptr, GZB_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
    double* base = (double*)buf.p;
    double* din[8];
    size_t off = 0;
    for (int k = 0; k < nin; ++k) {
        din[k] = base + off;
        e = cudaMemcpy(din[k], ins[k], in_sizes[k] * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return gzb_fail(nullptr, GZB_ERR_CUDA, "cudaMemcpy: %s", cudaGetErrorString(e));
        off += in_sizes[k];
    }
    double* dout = base + off;
    launch(din, dout);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, out_doubles * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return gzb_fail(nullptr, GZB_ERR_CUDA, "helper kernel: %s", cudaGetErrorString(e));
    return GZB_OK;
}

static thread_local long long h_n;
static thread_local double h_plat, h_plon;

extern "C" int gzb_heading(int device, int64_t n, const double* lata, const double* lona, const double* latb,
                           const double* lonb, double* out) {
    if (n <= 0) return GZB_OK;
    if (!lata || !lona || !latb || !lonb || !out) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_heading: null argument");
    const double* ins[4] = {lata, lona, latb, lonb};
    const size_t sz[4] = {(size_t)n, (size_t)n, (size_t)n, (size_t)n};
    h_n = n;
    return helper_run(device, 4 * (size_t)n, (size_t)n, ins, sz, 4, out, [](double** d, double* o) {
        k_heading<<<(unsigned)((h_n + 255) / 256), 256>>>(h_n, d[0], d[1], d[2], d[3], o);
    });
}

extern "C" int gzb_alfabeta(int device, int64_t n, const double* vy5, const double* vx5, double* ab_out) {
    if (n <= 0) return GZB_OK;
    if (!vy5 || !vx5 || !ab_out) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_alfabeta: null argument");
    const double* ins[2] = {vy5, vx5};
    const size_t sz[2] = {5 * (size_t)n, 5 * (size_t)n};
    h_n = n;
    return helper_run(device, 10 * (size_t)n, 2 * (size_t)n, ins, sz, 2, ab_out, [](double** d, double* o) {
        k_alfabeta<<<(unsigned)((h_n + 255) / 256), 256>>>(h_n, d[0], d[1], o);
    });
}

extern "C" int gzb_haversine(int device, int64_t n, double plat, double plon, const double* xlat,
                             const double* xlon, double* out) {
    if (n <= 0) return GZB_OK;
    if (!xlat || !xlon || !out) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_haversine: null argument");
    const double* ins[2] = {xlat, xlon};
    const size_t sz[2] = {(size_t)n, (size_t)n};
    h_n = n;
    h_plat = plat;
    h_plon = plon;
    return helper_run(device, 2 * (size_t)n, (size_t)n, ins, sz, 2, out, [](double** d, double* o) {
        k_haversine<<<(unsigned)((h_n + 255) / 256), 256>>>(h_n, h_plat, h_plon, d[0], d[1], o);
    });
}

extern "C" int gzb_haversine_sclr(int device, int64_t n, const double* lat1, const double* lon1, const double* lat2,
                                  const double* lon2, double* out) {
    if (n <= 0) return GZB_OK;
    if (!lat1 || !lon1 || !lat2 || !lon2 || !out) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_haversine_sclr: null argument");
    const double* ins[4] = {lat1, lon1, lat2, lon2};
    const size_t sz[4] = {(size_t)n, (size_t)n, (size_t)n, (size_t)n};
    h_n = n;
    return helper_run(device, 4 * (size_t)n, (size_t)n, ins, sz, 4, out, [](double** d, double* o) {
        k_haversine_sclr<<<(unsigned)((h_n + 255) / 256), 256>>>(h_n, d[0], d[1], d[2], d[3], o);
    });
}

extern "C" int gzb_radius_earth(int device, int64_t n, const double* lat, double* out) {
    if (n <= 0) return GZB_OK;
    if (!lat || !out) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_radius_earth: null argument");
    const double* ins[1] = {lat};
    const size_t sz[1] = {(size_t)n};
    h_n = n;
    return helper_run(device, (size_t)n, (size_t)n, ins, sz, 1, out, [](double** d, double* o) {
        k_radius_earth<<<(unsigned)((h_n + 255) / 256), 256>>>(h_n, d[0], o);
    });
}
