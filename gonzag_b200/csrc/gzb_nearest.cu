// K1: nearest model-grid point for every track point, with the reference's sequential
// "previous hit seeds the next search box" semantics reproduced exactly
// (gonzag/bilin_mapping.py:144-191 NearestPoint, :566-593 BilinTrack.nrpt).
//
// Decomposition:
//   k_global   exact global first-index haversine argmin G(t) of every point.  A warp owns a
//              chunk of consecutive track points and walks it in order:
//                K1a  continuity-seeded local window: the 5x3 leaf tiles (20x24 cells) around
//                     the previous point's tile; 32 cells per step, warp REDUX for the minimum;
//                     a precomputed per-tile certificate (distance from the tile centre to the
//                     nearest cell OUTSIDE the window) proves the local minimum is global;
//                K1b  when the certificate fails (track jump, E-W seam, first point of a
//                     chunk): coarse-to-fine best-first descent of the bounding-sphere pyramid
//                     over the whole grid.
//              Both work on fp32 unit vectors and only PROPOSE cells: every cell whose chord
//              is within the fp32 error margin of the best one is then evaluated with the
//              reference's own fp64 haversine formula (32 candidates at a time, one per lane)
//              and the winner is the lexicographic minimum (distance, row-major index), i.e.
//              numpy.argmin's first occurrence.
//   k_flags    speculative chain: R(t) = E(t) (edge-excluded, threshold-accepted G(t)) whenever
//              G(t) lies inside the search box of E(t-1); the other points are "breaks".
//   k_walk     K1c: one warp per break replays the reference's recurrence R(t) = F_t(R(t-1))
//              with the true box-restricted search (the (2r+1)^2 window, clamped not wrapped,
//              accepted iff d < r0) until the chain rejoins the speculative one; k_valid
//              discards replays whose own starting assumption was overwritten by an earlier one.
#include "gzb_common.cuh"

#define GZB_NOFLAT 0x7fffffffffffffffLL
#define GZB_QCAP 64

// search modes
#define M_GLOBAL 0   // whole grid
#define M_BOX 1      // only cells inside the index box
#define M_EXCL 2     // only cells outside the index box, proxies only (certificate build)

struct WarpSm {
    int fJ[GZB_MAXLEV + 1];
    int fI[GZB_MAXLEV + 1];
    unsigned fmask[GZB_MAXLEV + 1];
    int pad;
    int qpt[GZB_QCAP];
    float qP[GZB_QCAP];
    long long qflat[GZB_QCAP];
    unsigned long long bd[32];   // per owned point: bits of the best exact distance
    long long bf[32];            // ... and the smallest flat index achieving it
    float pb2[32];               // ... and the current candidate bound (chord^2)
};

struct Srch {          // warp-uniform state of the search for one point
    float px, py, pz;  // target unit vector
    float Pmin;        // smallest chord^2 seen (fp32 proxy)
    float B, B2;       // candidate / pruning bound on the chord and its square
    int bJ, bI;        // leaf tile holding the best proxy
};

struct Owner {         // the point a lane owns (exact evaluation needs its coordinates)
    double plat, plon, cpl;
};

struct NearParams {
    double r0;         // rd_found_km
    double thr2;       // fl(1.2*fl(1.2*r0)): acceptance radius of the last useful global pass
    float prox_r0;     // chord^2 bound that contains every cell with d < r0
    double chord_r0;   // chord length of r0 (with margin)
    int r;             // np_box_r
    int edge;          // apply the edge exclusion of BilinTrack.nrpt
    long long seed_j, seed_i;
    const double* corner;   // bounding sphere (cx,cy,cz,rho) of the box of predecessor (-1,-1)
};

// fp32 error budget (see DESIGN.md "K1 numerics"): a stored coordinate is within 6e-8 of the
// fp64 one, so a computed chord D differs from the true chord c by at most 3e-7*D + 3e-7.
// Every cell whose true chord is within 1e-9 relative of the best is kept as a candidate if the
// bound is  B = Dmin*(1+2e-6) + 2e-6  (chord units; 2e-6 is 12.7 m on the ground).
__device__ __forceinline__ void set_bound(Srch& s, float pm) {
    s.Pmin = pm;
    s.B = sqrtf(pm) * (1.0f + 2.0e-6f) + 2.0e-6f;
    s.B2 = s.B * s.B;
}

__device__ __forceinline__ void srch_init(Srch& s, float px, float py, float pz) {
    s.px = px; s.py = py; s.pz = pz;
    s.Pmin = INFINITY; s.B = INFINITY; s.B2 = INFINITY;
    s.bJ = -1; s.bI = -1;
}

// exact evaluation of the queued candidates, 32 at a time, and per-point lexicographic minimum
__device__ __forceinline__ void flush_queue(const GzbGrid& g, WarpSm& w, const int lane, int& qn, const Owner& me) {
    __syncwarp();
    for (int base = 0; base < qn; base += 32) {
        const int e = base + lane;
        bool act = e < qn;
        const int pt = act ? w.qpt[e] : 0;
        const double plat = __shfl_sync(0xffffffffu, me.plat, pt);
        const double plon = __shfl_sync(0xffffffffu, me.plon, pt);
        const double cpl = __shfl_sync(0xffffffffu, me.cpl, pt);
        act = act && (w.qP[e] <= w.pb2[pt]);
        unsigned long long dbits = ~0ull, pre = 0ull;
        long long flat = 0;
        if (act) {
            flat = w.qflat[e];
            const double d = gzb_haversine_dev(plat, plon, cpl, g.lat[flat], g.lon[flat]);
            if (d == d) dbits = (unsigned long long)__double_as_longlong(d); else act = false;
        }
        if (act) pre = w.bd[pt];
        __syncwarp();
        if (act) atomicMin(&w.bd[pt], dbits);
        __syncwarp();
        const bool holder = act && (w.bd[pt] == dbits);
        if (holder && dbits < pre) w.bf[pt] = GZB_NOFLAT;
        __syncwarp();
        if (holder) atomicMin(&w.bf[pt], flat);
        __syncwarp();
    }
    qn = 0;
}

// 32 cells of leaf tile (tJ, tI), one per lane
template <int MODE>
__device__ __forceinline__ void leaf_eval(const GzbGrid& g, WarpSm& w, const int lane, Srch& s, const int pt,
                                          const int tJ, const int tI, const long long j0, const long long j1,
                                          const long long i0, const long long i1, int& qn, const Owner& me) {
    const float* __restrict__ cc = g.cells + ((long long)tJ * g.lev[0].nI + tI) * 96;
    const float x = cc[lane], y = cc[32 + lane], z = cc[64 + lane];
    const long long j = (long long)tJ * GZB_TILE_J + (lane >> 3);
    const long long i = (long long)tI * GZB_TILE_I + (lane & 7);
    // padded cells of edge tiles are stored at (1e3,1e3,1e3): chord^2 ~ 3e6, never a candidate
    bool ok = true;
    if (MODE != M_GLOBAL) {
        const bool inside = (j >= j0) && (j < j1) && (i >= i0) && (i < i1);
        ok = (j < g.Nj) && (i < g.Ni) && (MODE == M_BOX ? inside : !inside);
    }
    const float dx = s.px - x, dy = s.py - y, dz = s.pz - z;
    const float P = dx * dx + dy * dy + dz * dz;
    const float pm = gzb_warp_min_nonneg(ok ? P : INFINITY);
    if (pm < s.Pmin) {
        set_bound(s, pm);
        s.bJ = tJ;
        s.bI = tI;
        if (MODE != M_EXCL && lane == 0) w.pb2[pt] = s.B2;
    }
    if (MODE != M_EXCL) {
        const bool cand = ok && (P <= s.B2);
        const unsigned cm = __ballot_sync(0xffffffffu, cand);
        const int nnew = __popc(cm);
        if (nnew) {
            if (qn + nnew > GZB_QCAP) flush_queue(g, w, lane, qn, me);
            if (cand) {
                const int pos = qn + __popc(cm & ((1u << lane) - 1u));
                w.qpt[pos] = pt;
                w.qP[pos] = P;
                w.qflat[pos] = j * g.Ni + i;
            }
            qn += nnew;
        }
        __syncwarp();
    }
}

// K1b: best-first descent of the pyramid, pruned by the running bound s.B
template <int MODE>
__device__ __forceinline__ void pyramid_search(const GzbGrid& g, WarpSm& w, const int lane, Srch& s, const int pt,
                                               const long long j0, const long long j1, const long long i0,
                                               const long long i1, int& qn, const Owner& me) {
    int depth = 0;
    if (lane == 0) w.fmask[0] = 0xffffffffu;
    __syncwarp();
    while (depth >= 0) {
        const int cl = g.nlev - 1 - depth;   // level of this frame's children
        const GzbLevel L = g.lev[cl];
        int J = 0, I = 0;
        long long base = 0;
        int cJ, cI;
        if (depth == 0) {
            cJ = lane / L.nI;
            cI = lane - cJ * L.nI;
        } else {
            const GzbLevel P = g.lev[cl + 1];
            J = w.fJ[depth];
            I = w.fI[depth];
            base = ((long long)J * P.nI + I) * 32;
            const int a = lane >> P.sbi;
            cJ = (J << P.sbj) + a;
            cI = (I << P.sbi) + (lane - (a << P.sbi));
        }
        const float4 nd = L.nodes[base + lane];
        bool keep = ((w.fmask[depth] >> lane) & 1u) && (nd.w >= 0.0f);
        if (MODE != M_GLOBAL && keep) {
            const long long a0 = (long long)cJ * L.sJ, b0 = (long long)cI * L.sI;
            if (MODE == M_BOX) keep = (a0 < j1) && (a0 + L.sJ > j0) && (b0 < i1) && (b0 + L.sI > i0);
            else keep = !((a0 >= j0) && (a0 + L.sJ <= j1) && (b0 >= i0) && (b0 + L.sI <= i1));
        }
        const float dx = s.px - nd.x, dy = s.py - nd.y, dz = s.pz - nd.z;
        const float dn2 = dx * dx + dy * dy + dz * dz;
        const float lim = nd.w + s.B;
        keep = keep && (dn2 <= lim * lim);
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m == 0u) {
            --depth;
            continue;
        }
        const unsigned key = keep ? ((__float_as_uint(dn2) & 0xffffffe0u) | (unsigned)lane) : 0xffffffffu;
        const int c = (int)(__reduce_min_sync(0xffffffffu, key) & 31u);
        m &= ~(1u << c);
        const int sJ = __shfl_sync(0xffffffffu, cJ, c);
        const int sI = __shfl_sync(0xffffffffu, cI, c);
        __syncwarp();
        if (lane == 0) w.fmask[depth] = m;
        if (cl > 0) {
            ++depth;
            if (lane == 0) {
                w.fJ[depth] = sJ;
                w.fI[depth] = sI;
                w.fmask[depth] = 0xffffffffu;
            }
            __syncwarp();
        } else {
            leaf_eval<MODE>(g, w, lane, s, pt, sJ, sI, j0, j1, i0, i1, qn, me);
            __syncwarp();
        }
    }
}

// K1a: the 5x3-tile window around tile (sJ, sI); returns true when the certificate proves that
// no cell outside the window can tie or beat the best cell found inside it
#define NBJ 2
#define NBI 1
template <int MODE>
__device__ __forceinline__ bool window_search(const GzbGrid& g, WarpSm& w, const int lane, Srch& s, const int pt,
                                              const int sJ, const int sI, const long long j0, const long long j1,
                                              const long long i0, const long long i1, int& qn, const Owner& me) {
    const GzbLevel& L0 = g.lev[0];
    const int nJ = sJ + lane / (2 * NBI + 1) - NBJ;
    const int nI = sI + lane % (2 * NBI + 1) - NBI;
    const bool nv = (lane < (2 * NBJ + 1) * (2 * NBI + 1)) && nJ >= 0 && nJ < L0.nJ && nI >= 0 && nI < L0.nI;
    const long long nidx = nv ? gzb_leaf_index(g, nJ, nI) : 0;
    const int centre = NBJ * (2 * NBI + 1) + NBI;
    float4 nd = make_float4(0.f, 0.f, 0.f, -1.f);
    float gc = 0.f;
    if (nv) nd = L0.nodes[nidx];
    if (lane == centre) gc = g.cert[nidx];
    const float dx = s.px - nd.x, dy = s.py - nd.y, dz = s.pz - nd.z;
    const float dn2 = dx * dx + dy * dy + dz * dz;
    unsigned pending = __ballot_sync(0xffffffffu, nv && nd.w >= 0.0f);
    for (;;) {
        const float lim = nd.w + s.B;
        const bool keep = ((pending >> lane) & 1u) && (dn2 <= lim * lim);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m == 0u) break;
        const unsigned key = keep ? ((__float_as_uint(dn2) & 0xffffffe0u) | (unsigned)lane) : 0xffffffffu;
        const int c = (int)(__reduce_min_sync(0xffffffffu, key) & 31u);
        pending &= ~(1u << c);
        leaf_eval<MODE>(g, w, lane, s, pt, __shfl_sync(0xffffffffu, nJ, c), __shfl_sync(0xffffffffu, nI, c),
                        j0, j1, i0, i1, qn, me);
    }
    const float a = sqrtf(__shfl_sync(0xffffffffu, dn2, centre));
    const float gcert = __shfl_sync(0xffffffffu, gc, centre);
    return (gcert - (a * (1.0f + 2.0e-7f) + 5.0e-7f)) > (s.B + 1.0e-6f);
}

__device__ __forceinline__ void owner_point(const double lat, const double lon, Owner& me, float& px, float& py,
                                            float& pz) {
    const double la = lat * GZB_D2R, lo = lon * GZB_D2R;
    me.plat = lat;
    me.plon = lon;
    me.cpl = cos(la);
    px = (float)(me.cpl * cos(lo));
    py = (float)(me.cpl * sin(lo));
    pz = (float)sin(la);
}

// ------------------------------------------------------------------------------------------
// global nearest point of every track point; a warp owns K (<= 32) consecutive points
__global__ void __launch_bounds__(256)
k_global(const GzbGrid g, const long long nt, const double* __restrict__ yt, const double* __restrict__ xt,
         long long* __restrict__ G, double* __restrict__ dG, const int K) {
    __shared__ WarpSm ws[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSm& w = ws[warp];
    const long long nchunk = (nt + K - 1) / K;
    for (long long ch = (long long)blockIdx.x * 8 + warp; ch < nchunk; ch += (long long)gridDim.x * 8) {
        const long long t0 = ch * K;
        const int n = (int)((nt - t0 < K) ? (nt - t0) : K);
        Owner me;
        float mpx = 0.f, mpy = 0.f, mpz = 0.f;
        me.plat = me.plon = me.cpl = 0.0;
        if (lane < n) owner_point(yt[t0 + lane], xt[t0 + lane], me, mpx, mpy, mpz);
        w.bd[lane] = ~0ull;
        w.bf[lane] = GZB_NOFLAT;
        w.pb2[lane] = INFINITY;
        __syncwarp();
        int qn = 0;
        int sJ = -1, sI = -1;
        for (int q = 0; q < n; ++q) {
            Srch s;
            srch_init(s, __shfl_sync(0xffffffffu, mpx, q), __shfl_sync(0xffffffffu, mpy, q),
                      __shfl_sync(0xffffffffu, mpz, q));
            bool proven = false;
            if (sJ >= 0) proven = window_search<M_GLOBAL>(g, w, lane, s, q, sJ, sI, 0, 0, 0, 0, qn, me);
            if (!proven) pyramid_search<M_GLOBAL>(g, w, lane, s, q, 0, 0, 0, 0, qn, me);
            sJ = s.bJ;
            sI = s.bI;
        }
        flush_queue(g, w, lane, qn, me);
        if (lane < n) {
            G[t0 + lane] = w.bf[lane];
            dG[t0 + lane] = __longlong_as_double((long long)w.bd[lane]);
        }
        __syncwarp();
    }
}

// certificate of every leaf tile: chord from its centre to the nearest cell outside its window
__global__ void __launch_bounds__(256)
k_cert(const GzbGrid g, float* __restrict__ cert) {
    __shared__ WarpSm ws[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GzbLevel L0 = g.lev[0];
    const long long ntile = (long long)L0.nJ * L0.nI;
    const long long tile = (long long)blockIdx.x * 8 + warp;
    if (tile >= ntile) return;
    const int tJ = (int)(tile / L0.nI), tI = (int)(tile - (long long)tJ * L0.nI);
    const long long idx = gzb_leaf_index(g, tJ, tI);
    const float4 nd = L0.nodes[idx];
    Srch s;
    srch_init(s, nd.x, nd.y, nd.z);
    Owner me;
    me.plat = me.plon = me.cpl = 0.0;
    int qn = 0;
    pyramid_search<M_EXCL>(g, ws[warp], lane, s, 0, (long long)(tJ - NBJ) * GZB_TILE_J,
                           (long long)(tJ + NBJ + 1) * GZB_TILE_J, (long long)(tI - NBI) * GZB_TILE_I,
                           (long long)(tI + NBI + 1) * GZB_TILE_I, qn, me);
    if (lane == 0) {
        float gval = INFINITY;
        if (s.Pmin < INFINITY) gval = fmaxf((sqrtf(s.Pmin) - 5.0e-7f) * (1.0f - 1.0e-6f), 0.0f);
        cert[idx] = gval;
    }
}

int gzb_build_cert(gzb_ctx* ctx) {
    const long long ntile = (long long)ctx->g.lev[0].nJ * ctx->g.lev[0].nI;
    k_cert<<<(unsigned)((ntile + 7) / 8), 256, 0, ctx->stream>>>(ctx->g, ctx->d_cert);
    GZB_LAUNCH_CHECK(ctx);
    return GZB_OK;
}

// threshold acceptance of the global hit + the edge exclusion of BilinTrack.nrpt
// (gonzag/bilin_mapping.py:169-190 and :582-585; column 0 is always excluded)
__device__ __forceinline__ void accept_edge(const GzbGrid& g, const long long flat, const double d,
                                            const double thr, long long& j, long long& i, const int edge = 1) {
    if (flat != GZB_NOFLAT && d < thr) {
        j = flat / g.Ni;
        i = flat - j * g.Ni;
    } else {
        j = -1;
        i = -1;
    }
    if (edge) {
        if (i == 0 || (i == g.Ni - 1 && g.kew == -1)) { j = -1; i = -1; }
        if (j == 0 || j == g.Nj - 1) { j = -1; i = -1; }
    }
}

__device__ __forceinline__ bool clamp_box(const GzbGrid& g, const long long pj, const long long pi, const int r,
                                          long long& j0, long long& j1, long long& i0, long long& i1) {
    j0 = pj - r > 0 ? pj - r : 0;
    j1 = pj + r + 1 < g.Nj ? pj + r + 1 : g.Nj;
    i0 = pi - r > 0 ? pi - r : 0;
    i1 = pi + r + 1 < g.Ni ? pi + r + 1 : g.Ni;
    return (j0 < j1) && (i0 < i1);
}

// K1c, step 1: speculative chain + break detection.  One thread per point.
__global__ void __launch_bounds__(1024)
k_flags(const GzbGrid g, const NearParams p, const long long nt, const double* __restrict__ yt,
        const double* __restrict__ xt, const long long* __restrict__ G, const double* __restrict__ dG,
        long long* __restrict__ NP, unsigned char* __restrict__ brk, int* __restrict__ counts) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int isb = 0;
    if (t < nt) {
        long long ej, ei;
        accept_edge(g, G[t], dG[t], p.thr2, ej, ei, p.edge);
        NP[2 * t] = ej;
        NP[2 * t + 1] = ei;
        long long pj, pi;
        if (t == 0) {
            pj = p.seed_j;
            pi = p.seed_i;
        } else {
            accept_edge(g, G[t - 1], dG[t - 1], p.thr2, pj, pi, p.edge);
        }
        if (pj != 0 && pi != 0) {   // Python truthiness of (j_prv and i_prv)
            long long j0, j1, i0, i1;
            if (clamp_box(g, pj, pi, p.r, j0, j1, i0, i1)) {
                const long long f = G[t];
                const long long gj = f / g.Ni, gi = f - gj * g.Ni;
                const bool inside = (f != GZB_NOFLAT) && gj >= j0 && gj < j1 && gi >= i0 && gi < i1;
                if (!inside) {
                    isb = 1;
                    if (pj == -1 && pi == -1 && p.corner[3] >= 0.0) {
                        // the box of a not-found predecessor is the grid corner [0,r)^2: if even
                        // its bounding sphere is farther than r0 the box pass cannot accept
                        const double la = yt[t] * GZB_D2R, lo = xt[t] * GZB_D2R;
                        const double cl = cos(la);
                        const double dx = cl * cos(lo) - p.corner[0], dy = cl * sin(lo) - p.corner[1],
                                     dz = sin(la) - p.corner[2];
                        if (sqrt(dx * dx + dy * dy + dz * dz) - p.corner[3] > p.chord_r0) isb = 0;
                    }
                }
            }
        }
        brk[t] = (unsigned char)isb;
    }
    const int c = __syncthreads_count(isb);
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

// exclusive scan of the per-block break counts (single block)
__global__ void __launch_bounds__(1024)
k_scan_counts(const int* __restrict__ counts, long long* __restrict__ offsets, const long long nblk,
              long long* __restrict__ total) {
    __shared__ long long wsum[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < nblk; base += 1024) {
        const long long k = base + threadIdx.x;
        long long v = k < nblk ? counts[k] : 0;
        long long s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        if (lane == 31) wsum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            long long ww = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long u = __shfl_up_sync(0xffffffffu, ww, o);
                if (lane >= o) ww += u;
            }
            wsum[lane] = ww;
        }
        __syncthreads();
        const long long pre = carry + (warp ? wsum[warp - 1] : 0) + s - v;
        if (k < nblk) offsets[k] = pre;
        __syncthreads();
        if (threadIdx.x == 1023) carry = pre + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// ordered compaction of the break indices
__global__ void __launch_bounds__(1024)
k_compact(const unsigned char* __restrict__ brk, const long long* __restrict__ offsets, const long long nt,
          long long* __restrict__ breaks) {
    __shared__ int wcnt[32];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = (t < nt) ? brk[t] : 0;
    const unsigned m = __ballot_sync(0xffffffffu, b);
    if (lane == 0) wcnt[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
        int ww = wcnt[lane];
        int s = ww;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        wcnt[lane] = s - ww;
    }
    __syncthreads();
    if (b) breaks[offsets[blockIdx.x] + wcnt[warp] + __popc(m & ((1u << lane) - 1u))] = t;
}

// one step of the reference recurrence: R(t) = F_t(pred).  All lanes hold the same point.
__device__ __forceinline__ void chain_step(const GzbGrid& g, WarpSm& w, const int lane, const NearParams& p,
                                           const long long t, const long long pj, const long long pi,
                                           const double* __restrict__ yt, const double* __restrict__ xt,
                                           const long long ej, const long long ei, long long& rj, long long& ri) {
    rj = ej;
    ri = ei;
    if (pj != 0 && pi != 0) {
        long long j0, j1, i0, i1;
        if (clamp_box(g, pj, pi, p.r, j0, j1, i0, i1)) {
            Owner me;
            float px, py, pz;
            owner_point(yt[t], xt[t], me, px, py, pz);
            Srch s;
            srch_init(s, px, py, pz);
            set_bound(s, p.prox_r0);
            if (lane == 0) {
                w.bd[0] = ~0ull;
                w.bf[0] = GZB_NOFLAT;
                w.pb2[0] = s.B2;
            }
            __syncwarp();
            int qn = 0;
            bool proven = false;
            if (pj > 0 && pi > 0)      // continuity: the window around the predecessor's tile
                proven = window_search<M_BOX>(g, w, lane, s, 0, (int)(pj >> 2), (int)(pi >> 3), j0, j1, i0, i1, qn, me);
            if (!proven) pyramid_search<M_BOX>(g, w, lane, s, 0, j0, j1, i0, i1, qn, me);
            flush_queue(g, w, lane, qn, me);
            const long long f = w.bf[0];
            const double d = __longlong_as_double((long long)w.bd[0]);
            __syncwarp();
            if (f != GZB_NOFLAT && d < p.r0) accept_edge(g, f, d, INFINITY, rj, ri, p.edge);
        }
    }
}

// K1c, step 2/4: chain replay, one warp per break.  WRITE=false records where the replay
// rejoins the speculative chain and logs every replayed value tagged with the replay's id
// (one 64-bit word per point: id in the high 27 bits, flat index + 1 in the low 36); WRITE=true
// (valid replays only) copies the logged values, 32 per step, and replays again only if another
// (discarded) replay overwrote part of its log.
#define WLOG_TAG(k) ((unsigned long long)(((k) + 1) & 0x7ffffffLL) << 36)

template <bool WRITE>
__global__ void __launch_bounds__(256)
k_walk(const GzbGrid g, const NearParams p, const long long nt, const double* __restrict__ yt,
       const double* __restrict__ xt, const long long* __restrict__ G, const double* __restrict__ dG,
       const long long* __restrict__ breaks, const long long* __restrict__ nbrk_p,
       long long* __restrict__ stop, long long* __restrict__ first, const unsigned char* __restrict__ valid,
       unsigned long long* __restrict__ wlog, long long* __restrict__ NP, unsigned long long* __restrict__ steps) {
    __shared__ WarpSm ws[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nb = *nbrk_p;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long k = (long long)blockIdx.x * 8 + warp; k < nb; k += nwarps) {
        const long long b = breaks[k];
        if (WRITE) {
            if (!valid[k]) continue;
            const long long e = stop[k];
            if (e == b) {
                if (lane == 0) {
                    NP[2 * b] = first[2 * k];
                    NP[2 * b + 1] = first[2 * k + 1];
                }
                continue;
            }
            // copy from the log
            bool clobbered = false;
            for (long long t0 = b; t0 <= e; t0 += 32) {
                const long long t = t0 + lane;
                bool bad = false;
                if (t <= e) {
                    const unsigned long long v = wlog[t];
                    if ((v >> 36) != (WLOG_TAG(k) >> 36)) {
                        bad = true;
                    } else {
                        const long long f = (long long)(v & 0xfffffffffULL) - 1;
                        const long long j = f < 0 ? -1 : f / g.Ni;
                        NP[2 * t] = j;
                        NP[2 * t + 1] = f < 0 ? -1 : f - j * g.Ni;
                    }
                }
                if (__any_sync(0xffffffffu, bad)) clobbered = true;
            }
            if (!clobbered) continue;
        }
        long long pj, pi;
        if (b == 0) {
            pj = p.seed_j;
            pi = p.seed_i;
        } else {
            accept_edge(g, G[b - 1], dG[b - 1], p.thr2, pj, pi, p.edge);
        }
        long long t = b;
        unsigned long long n = 0;
        for (;;) {
            long long ej, ei, rj, ri;
            accept_edge(g, G[t], dG[t], p.thr2, ej, ei, p.edge);
            chain_step(g, ws[warp], lane, p, t, pj, pi, yt, xt, ej, ei, rj, ri);
            ++n;
            if (WRITE) {
                if (lane == 0) {
                    NP[2 * t] = rj;
                    NP[2 * t + 1] = ri;
                }
                if (t >= stop[k]) break;
            } else {
                if (lane == 0) {
                    wlog[t] = WLOG_TAG(k) | (unsigned long long)(rj < 0 ? 0 : rj * g.Ni + ri + 1);
                    if (t == b) {
                        first[2 * k] = rj;
                        first[2 * k + 1] = ri;
                    }
                }
                if ((rj == ej && ri == ei) || t + 1 >= nt) {
                    if (lane == 0) stop[k] = t;
                    break;
                }
            }
            pj = rj;
            pi = ri;
            ++t;
        }
        if (!WRITE && lane == 0) atomicAdd(steps, n);
    }
}

// K1c, step 3: a replay is valid iff it does not start inside an earlier valid replay
// (whose deviating values would have been its true predecessor).  Single warp, 32 breaks per
// step on the common path where no replay reaches the next break.
__global__ void __launch_bounds__(32)
k_valid(const long long* __restrict__ breaks, const long long* __restrict__ stop,
        const long long* __restrict__ nbrk_p, unsigned char* __restrict__ valid, long long* __restrict__ stats_out) {
    const int lane = threadIdx.x;
    const long long nb = *nbrk_p;
    long long cur_end = -1;
    for (long long base = 0; base < nb; base += 32) {
        const long long k = base + lane;
        const long long b = k < nb ? breaks[k] : GZB_NOFLAT;
        const long long s = k < nb ? stop[k] : GZB_NOFLAT;
        long long nxt = __shfl_down_sync(0xffffffffu, b, 1);
        if (lane == 31) nxt = (base + 32 < nb) ? breaks[base + 32] : GZB_NOFLAT;
        const bool simple = (k >= nb) || (s < nxt);
        const long long b0 = __shfl_sync(0xffffffffu, b, 0);
        if (__all_sync(0xffffffffu, simple) && b0 > cur_end) {
            if (k < nb) valid[k] = 1;
            const int last = (int)((nb - base < 32 ? nb - base : 32) - 1);
            cur_end = __shfl_sync(0xffffffffu, s, last);
        } else {
            for (int l = 0; l < 32; ++l) {
                const long long kb = __shfl_sync(0xffffffffu, b, l);
                const long long ks = __shfl_sync(0xffffffffu, s, l);
                if (base + l < nb) {
                    const bool v = kb > cur_end;
                    if (v) cur_end = ks;
                    if (lane == l) valid[k] = v ? 1 : 0;
                }
            }
        }
    }
    if (lane == 0) {      // (breaks, replay steps) of this call, for gzb_get_stats
        stats_out[0] = nbrk_p[0];
        stats_out[1] = nbrk_p[1];
    }
}

// bounding sphere of the grid corner [0,r)^2 (the search box of a not-found predecessor)
__global__ void __launch_bounds__(256)
k_corner(const GzbGrid g, const int r, double* __restrict__ out) {
    __shared__ double sx[256], sy[256], sz[256];
    __shared__ double c[3];
    const long long rj = r < g.Nj ? r : g.Nj, ri = r < g.Ni ? r : g.Ni;
    const long long n = rj * ri;
    if (n <= 0) {
        if (threadIdx.x == 0) { out[0] = out[1] = out[2] = 0.0; out[3] = -1.0; }
        return;
    }
    double ax = 0, ay = 0, az = 0;
    for (long long k = threadIdx.x; k < n; k += blockDim.x) {
        const long long j = k / ri, i = k - j * ri;
        const double la = g.lat[j * g.Ni + i] * GZB_D2R, lo = g.lon[j * g.Ni + i] * GZB_D2R;
        const double cl = cos(la);
        ax += cl * cos(lo);
        ay += cl * sin(lo);
        az += sin(la);
    }
    sx[threadIdx.x] = ax; sy[threadIdx.x] = ay; sz[threadIdx.x] = az;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sx[threadIdx.x] += sx[threadIdx.x + o];
            sy[threadIdx.x] += sy[threadIdx.x + o];
            sz[threadIdx.x] += sz[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double nrm = sqrt(sx[0] * sx[0] + sy[0] * sy[0] + sz[0] * sz[0]);
        if (!(nrm > 1.0e-6)) { c[0] = 0; c[1] = 0; c[2] = 0; }   // centre at the origin: rho ~ 1
        else { c[0] = sx[0] / nrm; c[1] = sy[0] / nrm; c[2] = sz[0] / nrm; }
    }
    __syncthreads();
    double rm = 0.0;
    for (long long k = threadIdx.x; k < n; k += blockDim.x) {
        const long long j = k / ri, i = k - j * ri;
        const double la = g.lat[j * g.Ni + i] * GZB_D2R, lo = g.lon[j * g.Ni + i] * GZB_D2R;
        const double cl = cos(la);
        const double dx = cl * cos(lo) - c[0], dy = cl * sin(lo) - c[1], dz = sin(la) - c[2];
        rm = fmax(rm, sqrt(dx * dx + dy * dy + dz * dz));
    }
    sx[threadIdx.x] = rm;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sx[threadIdx.x] = fmax(sx[threadIdx.x], sx[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
        out[3] = sx[0] * (1.0 + 1.0e-12) + 1.0e-15;
    }
}

// ------------------------------------------------------------------------------------------
size_t gzb_nearest_ws_bytes(int64_t nt) {
    const size_t n = (size_t)(nt > 0 ? nt : 1);
    const size_t nblk = (n + 1023) / 1024;
    return gzb_align(n * 8) * 2      // G, dG
           + gzb_align(n)            // brk
           + gzb_align(nblk * 4) + gzb_align(nblk * 8)   // counts, offsets
           + gzb_align(n * 8) * 2    // breaks, stop
           + gzb_align(n * 16)       // first
           + gzb_align(n)            // valid
           + gzb_align(n * 8)        // replay log
           + gzb_align(64) * 2;      // scalars, corner
}

// device-level driver: all pointers are device pointers, work is enqueued on ctx->stream
int gzb_nearest_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, double rd_found_km,
                    int np_box_r, int64_t seed_j, int64_t seed_i, int64_t* np_out) {
    if (nt <= 0) return GZB_OK;
    if (np_box_r < 0) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_nearest: np_box_r must be >= 0");
    const size_t n = (size_t)nt;
    const long long nblk = (long long)((n + 1023) / 1024);
    long long* G = (long long*)gzb_arena_take(ctx, n * 8);
    double* dG = (double*)gzb_arena_take(ctx, n * 8);
    unsigned char* brk = (unsigned char*)gzb_arena_take(ctx, n);
    int* counts = (int*)gzb_arena_take(ctx, (size_t)nblk * 4);
    long long* offsets = (long long*)gzb_arena_take(ctx, (size_t)nblk * 8);
    long long* breaks = (long long*)gzb_arena_take(ctx, n * 8);
    long long* stop = (long long*)gzb_arena_take(ctx, n * 8);
    long long* first = (long long*)gzb_arena_take(ctx, n * 16);
    unsigned char* valid = (unsigned char*)gzb_arena_take(ctx, n);
    unsigned long long* wlog = (unsigned long long*)gzb_arena_take(ctx, n * 8);
    long long* scal = (long long*)gzb_arena_take(ctx, 64);
    if (!scal) return gzb_fail(ctx, GZB_ERR_NOMEM, "gzb_nearest: workspace too small");
    double* corner = (double*)(ctx->d_flags + 32);      // persistent: 4 doubles, keyed by np_box_r
    cudaStream_t s = ctx->stream;

    NearParams p;
    volatile double r0 = rd_found_km;
    volatile double t1 = 1.2 * r0;
    volatile double t2 = 1.2 * t1;
    p.r0 = r0;
    p.thr2 = t2;
    const double half = r0 / (2.0 * 6360.0);
    const double sn = sin(half < 1.5 ? half : 1.5);
    p.prox_r0 = (float)(4.0 * sn * sn * (1.0 + 1.0e-5) + 1.0e-20);
    p.chord_r0 = 2.0 * sn * (1.0 + 1.0e-6) + 1.0e-12;
    p.r = np_box_r;
    p.edge = ctx->edge_exclusion;
    p.seed_j = seed_j;
    p.seed_i = seed_i;
    p.corner = corner;

    GZB_CUDA(ctx, cudaMemsetAsync(scal, 0, 64, s));
    if (ctx->corner_r != np_box_r) {
        k_corner<<<1, 256, 0, s>>>(ctx->g, np_box_r, corner);
        GZB_LAUNCH_CHECK(ctx);
        ctx->corner_r = np_box_r;
    }

    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    // chunk length: long enough to amortise the unseeded first point, short enough to fill the GPU
    int K = 32;
    while (K > 4 && (long long)((n + K - 1) / K) < (long long)nsm * 8 * 4) K >>= 1;
    long long gblocks = (long long)(((n + K - 1) / K + 7) / 8);
    const long long gmax = (long long)nsm * 16;
    if (gblocks > gmax) gblocks = gmax;
    k_global<<<(unsigned)gblocks, 256, 0, s>>>(ctx->g, nt, yt, xt, G, dG, K);
    GZB_LAUNCH_CHECK(ctx);
    k_flags<<<(unsigned)nblk, 1024, 0, s>>>(ctx->g, p, nt, yt, xt, G, dG, (long long*)np_out, brk, counts);
    GZB_LAUNCH_CHECK(ctx);
    k_scan_counts<<<1, 1024, 0, s>>>(counts, offsets, nblk, scal);
    GZB_LAUNCH_CHECK(ctx);
    k_compact<<<(unsigned)nblk, 1024, 0, s>>>(brk, offsets, nt, breaks);
    GZB_LAUNCH_CHECK(ctx);
    long long wblocks = (long long)((n + 7) / 8);
    const long long wmax = (long long)nsm * 8;
    if (wblocks > wmax) wblocks = wmax;
    k_walk<false><<<(unsigned)wblocks, 256, 0, s>>>(ctx->g, p, nt, yt, xt, G, dG, breaks, scal, stop, first,
                                                   valid, wlog, (long long*)np_out, (unsigned long long*)(scal + 1));
    GZB_LAUNCH_CHECK(ctx);
    k_valid<<<1, 32, 0, s>>>(breaks, stop, scal, valid, (long long*)(ctx->d_flags + 8));
    GZB_LAUNCH_CHECK(ctx);
    k_walk<true><<<(unsigned)wblocks, 256, 0, s>>>(ctx->g, p, nt, yt, xt, G, dG, breaks, scal, stop, first,
                                                  valid, wlog, (long long*)np_out, (unsigned long long*)(scal + 1));
    GZB_LAUNCH_CHECK(ctx);
    ctx->nearest_stats_pending = 1;      // k_valid left (breaks, replay steps) in the persistent flags
    return GZB_OK;
}
