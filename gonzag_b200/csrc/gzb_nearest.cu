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
//   k_chain    (called "k_flags" in older notes)
//              speculative chain: R(t) = E(t) (edge-excluded, threshold-accepted G(t)) whenever
//              G(t) lies inside the search box of E(t-1); the other points are "breaks".
//   k_walk     K1c: one warp per break replays the reference's recurrence R(t) = F_t(R(t-1))
//              with the true box-restricted search (the (2r+1)^2 window, clamped not wrapped,
//              accepted iff d < r0) until the chain rejoins the speculative one; k_valid
//              discards replays whose own starting assumption was overwritten by an earlier one.
#include "gzb_common.cuh"

#define GZB_NOFLAT 0x7fffffffffffffffLL
#define GZB_UNRESOLVED 0x7ffffffffffffffeLL   /* G[t] until k_near_slow settles the point (bulk kernels skip it) */
#define GZB_QCAP 64
#define GZB_KCAP 64      // kept leaf tiles of one box scan
#define GZB_BOX_MAXT 4096 // largest box (in leaf tiles) scanned flat; larger boxes descend the pyramid

// search modes
#define M_GLOBAL 0   // whole grid
#define M_BOX 1      // only cells inside the index box
#define M_EXCL 2     // only cells outside the index box, proxies only (certificate build)

// flat index -> (j, i); grids of fewer than 2^32 cells (all but the largest) take a 32-bit division
__device__ __forceinline__ void gzb_split(const GzbGrid& g, const long long flat, long long& j, long long& i) {
    if ((unsigned long long)flat <= 0xffffffffull) {
        const unsigned f = (unsigned)flat, n = (unsigned)g.Ni;
        const unsigned q = f / n;
        j = (long long)q;
        i = (long long)(f - q * n);
    } else {
        j = flat / g.Ni;
        i = flat - j * g.Ni;
    }
}

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
    int kt[GZB_KCAP];            // box scan: leaf tiles (row-major id) that can hold a cell within the bound
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
    GzbSegs segs;      // sub-tracks of this call and their seeds (n >= 1, start[0] == 0)
    const double* corner;   // bounding sphere (cx,cy,cz,rho) of the box of predecessor (-1,-1)
    const unsigned char* rowsafe;   // per grid row: 1 = the only exact copies of its first k_ew_per cells are
                                    // their East-West twins (null: no twin shortcut)
};

// sub-track of point t: index of the last start <= t; `end` = last point of that sub-track
__device__ __forceinline__ int seg_of(const GzbSegs& S, const long long t, const long long nt, long long& end) {
    int s = 0;
    for (int q = 1; q < S.n; ++q)
        if (S.start[q] <= t) s = q;
    end = (s + 1 < S.n ? S.start[s + 1] : nt) - 1;
    return s;
}

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
            if (!GZB_CHK(g, flat >= 0 && flat < g.Nj * g.Ni, 32)) flat = 0;
            const double2 c = __ldg(g.ll + flat);
            const double d = gzb_haversine_dev(plat, plon, cpl, c.x, c.y);
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
                if (GZB_CHK(g, pos >= 0 && pos < GZB_QCAP, 16)) {
                    w.qpt[pos] = pt;
                    w.qP[pos] = P;
                    w.qflat[pos] = j * g.Ni + i;
                }
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

// K1c search: the cells of the index box [j0,j1) x [i0,i1) that can lie within the bound of `s`
// (the caller set it to the r0 ball).  A chain replay is a sequence of DEPENDENT searches, so this
// one is organised for latency, not throughput: the bounding spheres of all leaf tiles of the box
// are tested in parallel (4 per lane and per round, loads issued back to back), the few tiles
// that pass are evaluated with their loads batched four tiles at a time, and only then the
// candidates go through the exact fp64 formula (flush_queue, by the caller).  Boxes of more than
// GZB_BOX_MAXT tiles, or with more than GZB_KCAP tiles inside the bound, descend the pyramid.
__device__ __forceinline__ void box_search(const GzbGrid& g, WarpSm& w, const int lane, Srch& s, const int pt,
                                           const long long j0, const long long j1, const long long i0,
                                           const long long i1, int& qn, const Owner& me) {
    const GzbLevel& L0 = g.lev[0];
    const int tJ0 = (int)(j0 >> 2), tJ1 = (int)((j1 - 1) >> 2);
    const int tI0 = (int)(i0 >> 3), tI1 = (int)((i1 - 1) >> 3);
    const int nTI = tI1 - tI0 + 1;
    const long long nTl = (long long)(tJ1 - tJ0 + 1) * nTI;
    bool fallback = nTl > GZB_BOX_MAXT;
    int nk = 0;
    if (!fallback) {
        const int nT = (int)nTl;
        for (int base = 0; base < nT; base += 128) {
            float4 nd[4];
            int tl[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = base + u * 32 + lane;
                tl[u] = -1;
                nd[u] = make_float4(0.f, 0.f, 0.f, -1.f);
                if (k < nT) {
                    const int a = (int)((unsigned)k / (unsigned)nTI);
                    const int tj = tJ0 + a, ti = tI0 + (k - a * nTI);
                    tl[u] = tj * L0.nI + ti;
                    nd[u] = L0.nodes[gzb_leaf_index(g, tj, ti)];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dx = s.px - nd[u].x, dy = s.py - nd[u].y, dz = s.pz - nd[u].z;
                const float dn2 = dx * dx + dy * dy + dz * dz;
                const float lim = nd[u].w + s.B;
                const bool keep = (nd[u].w >= 0.0f) && (dn2 <= lim * lim);
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (m) {
                    const int cnt = __popc(m);
                    if (nk + cnt > GZB_KCAP) fallback = true;
                    else if (keep) w.kt[nk + __popc(m & ((1u << lane) - 1u))] = tl[u];
                    nk += cnt;
                }
            }
            if (fallback) break;
        }
        __syncwarp();
    }
    if (fallback) {
        pyramid_search<M_BOX>(g, w, lane, s, pt, j0, j1, i0, i1, qn, me);
        return;
    }
    if (nk == 0) return;
    // smallest proxy over the kept tiles
    const int lj = lane >> 3, li = lane & 7;
    float pmin = INFINITY;
    for (int b = 0; b < nk; b += 4) {
        float x[4], y[4], z[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            ok[u] = false;
            x[u] = y[u] = z[u] = 0.f;
            if (b + u < nk) {
                const int tile = w.kt[b + u];
                const float* __restrict__ cc = g.cells + (long long)tile * 96;
                x[u] = cc[lane]; y[u] = cc[32 + lane]; z[u] = cc[64 + lane];
                const int tj = (int)((unsigned)tile / (unsigned)L0.nI);
                const long long j = (long long)tj * GZB_TILE_J + lj;
                const long long i = (long long)(tile - tj * L0.nI) * GZB_TILE_I + li;
                ok[u] = (j >= j0) && (j < j1) && (i >= i0) && (i < i1);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float dx = s.px - x[u], dy = s.py - y[u], dz = s.pz - z[u];
            const float P = dx * dx + dy * dy + dz * dz;
            if (ok[u]) pmin = fminf(pmin, P);
        }
    }
    const float pm = gzb_warp_min_nonneg(pmin);
    if (pm < s.Pmin) {
        set_bound(s, pm);
        if (lane == 0) w.pb2[pt] = s.B2;
    }
    __syncwarp();
    // candidates: every box cell of the kept tiles whose proxy is within the (final) bound
    for (int b = 0; b < nk; ++b) {
        const int tile = w.kt[b];
        const float* __restrict__ cc = g.cells + (long long)tile * 96;
        const float x = cc[lane], y = cc[32 + lane], z = cc[64 + lane];
        const int tj = (int)((unsigned)tile / (unsigned)L0.nI);
        const long long j = (long long)tj * GZB_TILE_J + lj;
        const long long i = (long long)(tile - tj * L0.nI) * GZB_TILE_I + li;
        const float dx = s.px - x, dy = s.py - y, dz = s.pz - z;
        const float P = dx * dx + dy * dy + dz * dz;
        const bool cand = (j >= j0) && (j < j1) && (i >= i0) && (i < i1) && (P <= s.B2);
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


// ------------------------------------------------------------------------------------------
// K1, fast path: one THREAD per track point.
//   seed   the lat-lon look-up table gives a cell lying in (or next to) the point's bin;
//   climb  evaluate the 5x5 index neighbourhood of the current cell (fp32 chord^2 from the
//          row-major float4 array), move to its best cell, until the centre itself is the best;
//   prove  (3x3 certificate first: when it already holds, the ring is never loaded)
//          cellv[c].w bounds from below the chord from c to every cell outside its 5x5
//          neighbourhood, so  w - |p - c| > B  proves that no outside cell can tie or beat the
//          best inside one (B = the fp32 candidate bound, as in the warp search);
//   exact  the winner (and any other cell within the fp32 margin) goes through the reference's
//          fp64 haversine; first-index tie-break.
// Points without a seed, not converged in FAST_MAXIT moves, or not proven (duplicated E-W columns,
// strongly anisotropic cells, targets outside the grid) are appended to `ulist` with the last
// cell as a hint for the warp-cooperative search (k_near_slow).
#define FAST_MAXIT 32

// chord^2 (fp32 proxy) from the target to cell (j+DJ, i+DI); c0 points at cell (j, i)
template <bool CHECK, int DJ, int DI>
__device__ __forceinline__ float fast_cell(const GzbGrid& g, const float4* __restrict__ c0, const int j, const int i,
                                           const float px, const float py, const float pz) {
    if (CHECK) {
        if ((unsigned)(j + DJ) >= (unsigned)g.Nj || (unsigned)(i + DI) >= (unsigned)g.Ni) return INFINITY;
    }
    const float4 v = __ldg(c0 + ((long long)DJ * g.Ni + DI));
    const float dx = px - v.x, dy = py - v.y, dz = pz - v.z;
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
}

#define FAST_UPD(P, DJ, DI)                          \
    {                                                \
        const float P__ = (P);                       \
        if (P__ < nb) { nb = P__; mj = (DJ); mi = (DI); } \
    }

// 3x3 step around (j, i): moves (j, i) to the best of the 8 neighbours when it is strictly
// better than the centre (returns 1), else returns 0 with Pc = chord^2 of the centre, cw = its
// certificate, nb8 = the smallest chord^2 among the 8 neighbours.
template <bool CHECK>
__device__ __forceinline__ int fast_step3(const GzbGrid& g, const float px, const float py, const float pz,
                                          int& j, int& i, float& Pc, float& cw, float& nb8) {
    const float4* __restrict__ c0 = g.cellv + ((long long)j * g.Ni + i);
    const float4 vc = __ldg(c0);
    {
        const float dx = px - vc.x, dy = py - vc.y, dz = pz - vc.z;
        Pc = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
        cw = vc.w;
    }
    float nb = INFINITY;
    int mj = 0, mi = 0;
    FAST_UPD((fast_cell<CHECK, -1, -1>(g, c0, j, i, px, py, pz)), -1, -1)
    FAST_UPD((fast_cell<CHECK, -1, 0>(g, c0, j, i, px, py, pz)), -1, 0)
    FAST_UPD((fast_cell<CHECK, -1, 1>(g, c0, j, i, px, py, pz)), -1, 1)
    FAST_UPD((fast_cell<CHECK, 0, -1>(g, c0, j, i, px, py, pz)), 0, -1)
    FAST_UPD((fast_cell<CHECK, 0, 1>(g, c0, j, i, px, py, pz)), 0, 1)
    FAST_UPD((fast_cell<CHECK, 1, -1>(g, c0, j, i, px, py, pz)), 1, -1)
    FAST_UPD((fast_cell<CHECK, 1, 0>(g, c0, j, i, px, py, pz)), 1, 0)
    FAST_UPD((fast_cell<CHECK, 1, 1>(g, c0, j, i, px, py, pz)), 1, 1)
    nb8 = nb;
    if (nb < Pc) { j += mj; i += mi; return 1; }
    return 0;
}

// the 16 cells of the outer ring of the 5x5 neighbourhood of (j, i): moves to the best of them
// when it beats the centre (returns 1), else returns 0 with ring = their smallest chord^2
template <bool CHECK>
__device__ __forceinline__ int fast_ring(const GzbGrid& g, const float px, const float py, const float pz,
                                         int& j, int& i, const float Pc, float& ring) {
    const float4* __restrict__ c0 = g.cellv + ((long long)j * g.Ni + i);
    float nb = INFINITY;
    int mj = 0, mi = 0;
    FAST_UPD((fast_cell<CHECK, -2, -2>(g, c0, j, i, px, py, pz)), -2, -2)
    FAST_UPD((fast_cell<CHECK, -2, -1>(g, c0, j, i, px, py, pz)), -2, -1)
    FAST_UPD((fast_cell<CHECK, -2, 0>(g, c0, j, i, px, py, pz)), -2, 0)
    FAST_UPD((fast_cell<CHECK, -2, 1>(g, c0, j, i, px, py, pz)), -2, 1)
    FAST_UPD((fast_cell<CHECK, -2, 2>(g, c0, j, i, px, py, pz)), -2, 2)
    FAST_UPD((fast_cell<CHECK, -1, -2>(g, c0, j, i, px, py, pz)), -1, -2)
    FAST_UPD((fast_cell<CHECK, -1, 2>(g, c0, j, i, px, py, pz)), -1, 2)
    FAST_UPD((fast_cell<CHECK, 0, -2>(g, c0, j, i, px, py, pz)), 0, -2)
    FAST_UPD((fast_cell<CHECK, 0, 2>(g, c0, j, i, px, py, pz)), 0, 2)
    FAST_UPD((fast_cell<CHECK, 1, -2>(g, c0, j, i, px, py, pz)), 1, -2)
    FAST_UPD((fast_cell<CHECK, 1, 2>(g, c0, j, i, px, py, pz)), 1, 2)
    FAST_UPD((fast_cell<CHECK, 2, -2>(g, c0, j, i, px, py, pz)), 2, -2)
    FAST_UPD((fast_cell<CHECK, 2, -1>(g, c0, j, i, px, py, pz)), 2, -1)
    FAST_UPD((fast_cell<CHECK, 2, 0>(g, c0, j, i, px, py, pz)), 2, 0)
    FAST_UPD((fast_cell<CHECK, 2, 1>(g, c0, j, i, px, py, pz)), 2, 1)
    FAST_UPD((fast_cell<CHECK, 2, 2>(g, c0, j, i, px, py, pz)), 2, 2)
    ring = nb;
    if (nb < Pc) { j += mj; i += mi; return 1; }
    return 0;
}

// outcome of the per-thread search for one point (one 64-bit word):
//   CAND_UNRES        handed to the warp search (k_near_slow), which writes G/dG itself
//   CAND_NONE         non-finite target: no nearest point
//   flat | CAND_MULTI several cells of the 5x5 neighbourhood of `flat` are within the fp32 margin
//   flat              the one candidate
#define CAND_UNRES 0x4000000000000000LL
#define CAND_NONE 0x2000000000000000LL
#define CAND_MULTI 0x1000000000000000LL
#define CAND_FLAT(v) ((v) & 0xfffffffffLL)

// exact (reference formula, fp64) evaluation of the candidate(s) the per-thread search found for a point:
// returns the first-index argmin and its distance bits
__device__ __forceinline__ void exact_from_cand(const GzbGrid& g, const long long c, const double lat, const double lon,
                                                const float px, const float py, const float pz, long long& bf,
                                                unsigned long long& bd) {
    bf = GZB_NOFLAT;
    bd = ~0ull;
    if (c & CAND_NONE) return;
    const double cpl = cos(lat * GZB_D2R);
    const long long f0 = CAND_FLAT(c);
    if (!GZB_CHK(g, f0 >= 0 && f0 < g.Nj * g.Ni, 32)) return;
    if (!(c & CAND_MULTI)) {
        const double2 c0 = __ldg(g.ll + f0);
        const double d = gzb_haversine_dev(lat, lon, cpl, c0.x, c0.y);
        if (d == d) { bd = (unsigned long long)__double_as_longlong(d); bf = f0; }
        return;
    }
    // several cells within the fp32 margin of the best one: same proxy, same bound as the search
    const long long j = f0 / g.Ni, i = f0 - j * g.Ni;
    float B2;
    {
        const float4 v = __ldg(g.cellv + f0);
        const float dx = px - v.x, dy = py - v.y, dz = pz - v.z;
        const float a = sqrtf(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)));
        const float B = a * (1.0f + 2.0e-6f) + 2.0e-6f;
        B2 = B * B;
    }
    for (int dj = -2; dj <= 2; ++dj) {
        const long long jj = j + dj;
        if (jj < 0 || jj >= g.Nj) continue;
        for (int di = -2; di <= 2; ++di) {
            const long long ii = i + di;
            if (ii < 0 || ii >= g.Ni) continue;
            const long long f = jj * g.Ni + ii;
            const float4 v = __ldg(g.cellv + f);
            const float dx = px - v.x, dy = py - v.y, dz = pz - v.z;
            const float P = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
            if (!(P <= B2)) continue;
            const double2 cf = __ldg(g.ll + f);
            const double d = gzb_haversine_dev(lat, lon, cpl, cf.x, cf.y);
            if (!(d == d)) continue;
            const unsigned long long db = (unsigned long long)__double_as_longlong(d);
            if (db < bd) { bd = db; bf = f; }     // row-major scan: the first index keeps a tie
        }
    }
}

__global__ void __launch_bounds__(128, 8)
k_near_search(const GzbGrid g, const long long nt, const double* __restrict__ yt, const double* __restrict__ xt,
              long long* __restrict__ G, double* __restrict__ dG, long long* __restrict__ ulist,
              unsigned* __restrict__ hint, unsigned long long* __restrict__ ucount
#if GZB_EXP_SKIP_DG
              , const float cthr0, const float cthr1      // chords (unit sphere) of r0 and of 1.2^2 r0; NaN = no shortcut
#endif
              ) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    GZB_KT_MIN(g, 0);
    bool unresolved = false;
    unsigned myhint = 0xffffffffu;
#if GZB_EXP_SKIP_DG
    float a_win = -1.0f;                                  // fp32 chord of a proven, single winner
#endif
    if (t < nt) {
        const double lat = yt[t], lon = xt[t];
        long long out = CAND_NONE;
        if (isfinite(lat) && isfinite(lon)) {
            float px, py, pz;
            {
                const double la = lat * GZB_D2R, lo = lon * GZB_D2R;
                double sa, ca, so, co;
                sincos(la, &sa, &ca);
                sincos(lo, &so, &co);
                px = (float)(ca * co);
                py = (float)(ca * so);
                pz = (float)sa;
            }
            const unsigned seed = __ldg(g.lut.bins + gzb_lut_bin(g.lut, lat, lon));
            int reason = 1;      // 1 no seed, 2 not converged, 3 not proven
            if (seed != 0xffffffffu) {
                int j = (int)(seed / (unsigned)g.Ni);
                int i = (int)(seed - (unsigned)j * (unsigned)g.Ni);
                const unsigned nj4 = g.Nj > 4 ? (unsigned)(g.Nj - 4) : 0u, ni4 = g.Ni > 4 ? (unsigned)(g.Ni - 4) : 0u;
                const unsigned nj2 = (unsigned)(g.Nj - 2), ni2 = (unsigned)(g.Ni - 2);     // Nj, Ni >= 3
                float Pc = 0.0f, cw = 0.0f, nb8 = INFINITY, ring = INFINITY;
                bool conv = false, proven3 = false;
                int moves = 0;
                for (;;) {
                    // climb on 3x3 neighbourhoods; the lanes of a warp leave this loop together, so
                    // that the 16 ring loads below are issued once per warp
                    bool inner = (unsigned)(j - 1) < nj2 && (unsigned)(i - 1) < ni2;
                    while (moves < FAST_MAXIT &&
                           (inner ? fast_step3<false>(g, px, py, pz, j, i, Pc, cw, nb8)
                                  : fast_step3<true>(g, px, py, pz, j, i, Pc, cw, nb8))) {
                        ++moves;
                        inner = (unsigned)(j - 1) < nj2 && (unsigned)(i - 1) < ni2;
                    }
                    if (moves >= FAST_MAXIT) break;
                    {   // the centre is the best cell of its 3x3 neighbourhood: its 3x3 certificate may already
                        // prove that nothing farther away in index space can tie or beat it (no ring needed)
                        const float a = sqrtf(Pc);
                        const float B = a * (1.0f + 2.0e-6f) + 2.0e-6f;
                        const float w3 = __uint_as_float(__float_as_uint(cw) & 0xffff0000u);
                        if ((w3 - (a * (1.0f + 2.0e-7f) + 5.0e-7f)) > (B + 1.0e-6f)) {
                            proven3 = true;
                            conv = true;
                            break;
                        }
                    }
                    const bool interior = (unsigned)(j - 2) < nj4 && (unsigned)(i - 2) < ni4;
                    const int moved = interior ? fast_ring<false>(g, px, py, pz, j, i, Pc, ring)
                                               : fast_ring<true>(g, px, py, pz, j, i, Pc, ring);
                    if (!moved) { conv = true; break; }
                    ++moves;
                }
                const long long f = (long long)j * g.Ni + i;
                myhint = (unsigned)f;
                reason = 2;
                if (conv) {
                    reason = 3;
                    const float a = sqrtf(Pc);
                    const float B = a * (1.0f + 2.0e-6f) + 2.0e-6f;
                    const float w5 = __uint_as_float(__float_as_uint(cw) << 16);
                    if (proven3 || (w5 - (a * (1.0f + 2.0e-7f) + 5.0e-7f)) > (B + 1.0e-6f)) {
                        reason = 0;
                        out = f | ((fminf(nb8, ring) <= B * B) ? CAND_MULTI : 0LL);
#if GZB_EXP_SKIP_DG
                        if (!(out & CAND_MULTI)) a_win = a;
#endif
                    }
                }
            }
            if (reason) {
                unresolved = true;
                out = CAND_UNRES;
                atomicAdd(ucount + reason, 1ull);
            }
            if (!unresolved) {
                // settle the point here: the reference's fp64 formula on the candidate(s)
#if GZB_EXP_SKIP_DG
                const float m2 = a_win * 4.0e-6f + 4.0e-6f;           // twice the proven fp32 error of the chord
                if (a_win >= 0.0f && fabsf(a_win - cthr0) > m2 && fabsf(a_win - cthr1) > m2) {
                    G[t] = CAND_FLAT(out);
                    dG[t] = 12720.0 * asin(0.5 * (double)a_win);      // same side of r0 and 1.2^2 r0 as the exact distance
                } else
#endif
                {
                long long bf;
                unsigned long long bd;
                exact_from_cand(g, out, lat, lon, px, py, pz, bf, bd);
                G[t] = bf;
                dG[t] = __longlong_as_double((long long)bd);
                }
            }
        } else {
            G[t] = GZB_NOFLAT;
            dG[t] = __longlong_as_double(-1LL);
        }
    }
    // warp-aggregated append of the unresolved points
    const unsigned um = __ballot_sync(0xffffffffu, unresolved);
    if (um) {
        const int lane = threadIdx.x & 31;
        unsigned long long base = 0;
        if (lane == (__ffs(um) - 1)) base = atomicAdd(ucount, (unsigned long long)__popc(um));
        base = __shfl_sync(0xffffffffu, base, __ffs(um) - 1);
        if (unresolved && GZB_CHK(g, (long long)(base + __popc(um & ((1u << lane) - 1u))) < nt, 1)) {
            ulist[base + __popc(um & ((1u << lane) - 1u))] = t;
            hint[t] = myhint;
            G[t] = GZB_UNRESOLVED;
        }
    }
    GZB_KT_MAX(g, 0);
}

// K1, slow path: the warp-cooperative search (seeded window, then pyramid) for the listed points.
// A warp takes K list entries at a time (K chosen so that every warp of the grid has work).
// small blocks (SLOW_WARPS warps): they must find room on SMs that the bulk kernels of gzb_mapping fill
#define SLOW_WARPS 2
#ifndef GZB_SLOW_MINBLOCKS
#define GZB_SLOW_MINBLOCKS 10     /* resident blocks of 64 threads per SM the compiler must allow (10 = a cap of 102 registers) */
#endif
__global__ void __launch_bounds__(32 * SLOW_WARPS, GZB_SLOW_MINBLOCKS)
k_near_slow(const GzbGrid g, const double* __restrict__ yt, const double* __restrict__ xt,
            long long* __restrict__ G, double* __restrict__ dG, const long long* __restrict__ ulist,
            const unsigned* __restrict__ hint, const unsigned long long* __restrict__ ucount) {
    __shared__ WarpSm ws[SLOW_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSm& w = ws[warp];
    const long long nu = (long long)*ucount;
    GZB_KT_MIN(g, 1);
    if (nu <= 0) return;
    const long long nwarps = (long long)gridDim.x * SLOW_WARPS;
    long long K = (nu + nwarps - 1) / nwarps;
    K = K < 1 ? 1 : (K > 32 ? 32 : K);
    const long long nchunk = (nu + K - 1) / K;
    for (long long ch = (long long)blockIdx.x * SLOW_WARPS + warp; ch < nchunk; ch += nwarps) {
        const long long u0 = ch * K;
        const int n = (int)((nu - u0 < K) ? (nu - u0) : K);
        Owner me;
        float mpx = 0.f, mpy = 0.f, mpz = 0.f;
        me.plat = me.plon = me.cpl = 0.0;
        long long idx = 0;
        int hJ = -1, hI = -1;
        if (lane < n) {
            idx = ulist[u0 + lane];
            owner_point(yt[idx], xt[idx], me, mpx, mpy, mpz);
            const unsigned h = hint[idx];
            if (h != 0xffffffffu) {
                const long long hj = (long long)(h / (unsigned long long)g.Ni);
                hJ = (int)(hj >> 2);
                hI = (int)(((long long)h - hj * g.Ni) >> 3);
            }
        }
        w.bd[lane] = ~0ull;
        w.bf[lane] = GZB_NOFLAT;
        w.pb2[lane] = INFINITY;
        __syncwarp();
        int qn = 0;
        for (int q = 0; q < n; ++q) {
            Srch s;
            srch_init(s, __shfl_sync(0xffffffffu, mpx, q), __shfl_sync(0xffffffffu, mpy, q),
                      __shfl_sync(0xffffffffu, mpz, q));
            const int sJ = __shfl_sync(0xffffffffu, hJ, q), sI = __shfl_sync(0xffffffffu, hI, q);
            bool proven = false;
            if (sJ >= 0) proven = window_search<M_GLOBAL>(g, w, lane, s, q, sJ, sI, 0, 0, 0, 0, qn, me);
            if (!proven) pyramid_search<M_GLOBAL>(g, w, lane, s, q, 0, 0, 0, 0, qn, me);
        }
        flush_queue(g, w, lane, qn, me);
        if (lane < n) {
            G[idx] = w.bf[lane];
            dG[idx] = __longlong_as_double((long long)w.bd[lane]);
        }
        __syncwarp();
    }
    GZB_KT_MAX(g, 1);
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
        gzb_split(g, flat, j, i);
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

// K1c, step 1: the speculative chain and its breaks.
//
// The reference's R(t) = F_t(R(t-1)) depends on the predecessor only through its search box.  Two
// values of R(t) can be written down without any search:
//   E(t)  the global answer (edge-excluded, threshold-accepted G(t)): it is R(t) whenever G(t) lies in
//         the box of the predecessor (DESIGN.md, K1 key fact), or when there is no box at all;
//   T(t)  the East-West twin of G(t), when G(t) sits in one of the first k_ew_per columns, its copy
//         k_ew_per columns before the eastern edge has bit-identical coordinates (so the very same
//         distance) and d < r0: it is R(t) whenever G(t) is NOT in the box but the twin is -- the box
//         pass then finds the twin, at the global minimum distance, and accepts it.  (No third exact
//         copy can precede the twin in the box: it would have to sit in the same row between the two,
//         which `rowsafe` excludes.)
// A track crossing the seam eastwards is a long run of T(t)'s, each one depending on its predecessor
// being T(t-1): a two-state chain.  So every point gets a transition code -- what F_t gives from
// predecessor E(t-1) and from predecessor T(t-1): 0 = E(t), 1 = T(t), 2 = "needs a search" -- and the
// states are resolved for the whole track by a parallel scan of function compositions.  Points that
// come out as 2 are the breaks (speculation goes on from E(t) there); the replays (k_walk) start from
// the speculative predecessor and run until they meet the speculative chain again.
#define FLAGS_BLOCK 256

__device__ __forceinline__ bool twin_of(const GzbGrid& g, const NearParams& p, const long long flat, const double d,
                                        long long& tw) {
    if (g.kew <= 0 || !p.rowsafe || flat == GZB_NOFLAT || !(d < p.r0)) return false;
    long long j, i;
    gzb_split(g, flat, j, i);
    if (i >= g.kew || !p.rowsafe[j]) return false;
    tw = flat + (g.Ni - g.kew);
    const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(g.ll) + flat);
    const longlong2 b = __ldg(reinterpret_cast<const longlong2*>(g.ll) + tw);
    return a.x == b.x && a.y == b.y;
}

// what the chain logic needs to know about one point: G(t) split, E(t), T(t)
struct PointSum {
    int gj, gi;       // G(t), (-1,-1) if none
    int ej, ei;       // E(t)
    int tj, ti;       // T(t); tv == 0: no twin state
    int tv;
};

__device__ __forceinline__ void point_sum(const GzbGrid& g, const NearParams& p, const long long flat, const double d,
                                          PointSum& q) {
    q.gj = q.gi = q.ej = q.ei = q.tj = q.ti = -1;
    q.tv = 0;
    if (flat == GZB_NOFLAT || flat == GZB_UNRESOLVED) return;
    long long j, i;
    gzb_split(g, flat, j, i);
    q.gj = (int)j;
    q.gi = (int)i;
    if (d < p.thr2) { q.ej = q.gj; q.ei = q.gi; }
    if (p.edge) {
        if (q.ei == 0 || (q.ei == g.Ni - 1 && g.kew == -1)) { q.ej = -1; q.ei = -1; }
        if (q.ej == 0 || q.ej == g.Nj - 1) { q.ej = -1; q.ei = -1; }
    }
    if (g.kew > 0 && p.rowsafe && i < g.kew && d < p.r0 && p.rowsafe[j]) {
        const long long tw = flat + (g.Ni - g.kew);
        const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(g.ll) + flat);
        const longlong2 b = __ldg(reinterpret_cast<const longlong2*>(g.ll) + tw);
        if (a.x == b.x && a.y == b.y) {
            q.tv = 1;
            q.tj = q.gj;
            q.ti = q.gi + (int)(g.Ni - g.kew);
            if (p.edge) {        // the twin column is never column 0; the last column only counts without periodicity
                if (q.tj == 0 || q.tj == g.Nj - 1) { q.tj = -1; q.ti = -1; }
            }
        }
    }
}

// what the reference's step gives for point t from predecessor (pj, pi), as far as it can be told
// without a search: 0 = E(t), 1 = T(t), 2 = unknown
__device__ __forceinline__ int step_code(const GzbGrid& g, const NearParams& p, const int pj, const int pi,
                                         const PointSum& q, const double* __restrict__ yt, const double* __restrict__ xt,
                                         const long long t) {
    if (pj == 0 || pi == 0) return 0;           // Python truthiness of (j_prv and i_prv): no box
    const int j0 = pj - p.r > 0 ? pj - p.r : 0;
    const int j1 = (long long)pj + p.r + 1 < g.Nj ? pj + p.r + 1 : (int)g.Nj;
    const int i0 = pi - p.r > 0 ? pi - p.r : 0;
    const int i1 = (long long)pi + p.r + 1 < g.Ni ? pi + p.r + 1 : (int)g.Ni;
    if (!(j0 < j1 && i0 < i1)) return 0;
    if (q.gj >= 0) {
        if (q.gj >= j0 && q.gj < j1 && q.gi >= i0 && q.gi < i1) return 0;
        if (q.tv) {
            const int ti = q.gi + (int)(g.Ni - g.kew);      // the twin cell itself (before the edge rule)
            if (q.gj >= j0 && q.gj < j1 && ti >= i0 && ti < i1) return 1;
        }
    }
    if (pj == -1 && pi == -1 && p.corner[3] >= 0.0) {
        // the box of a not-found predecessor is the grid corner [0,r)^2: if even its bounding sphere
        // is farther than r0 the box pass cannot accept
        const double la = yt[t] * GZB_D2R, lo = xt[t] * GZB_D2R;
        const double cl = cos(la);
        const double dx = cl * cos(lo) - p.corner[0], dy = cl * sin(lo) - p.corner[1], dz = sin(la) - p.corner[2];
        if (sqrt(dx * dx + dy * dy + dz * dz) - p.corner[3] > p.chord_r0) return 0;
    }
    return 2;
}

// a map {0,1} -> {0,1} in two bits (bit s = image of s); composition "first f, then h"
__device__ __forceinline__ unsigned fn_then(const unsigned f, const unsigned h) {
    return ((h >> (f & 1u)) & 1u) | (((h >> ((f >> 1) & 1u)) & 1u) << 1);
}
#define FN_ID 2u

// inclusive scan of compositions over the block; returns this thread's prefix (maps the state that
// enters the block to the state that LEAVES this thread's point); total = the whole block's map.
// Away from the seam every point's map is "both states -> E" (0): a warp of those needs no shuffles.
__device__ __forceinline__ unsigned block_scan_fn(const unsigned mine, unsigned* sh /* [FLAGS_BLOCK/32] */, unsigned& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned f = mine;
    if (!__all_sync(0xffffffffu, mine == 0u)) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, f, o);
            if (lane >= o) f = fn_then(u, f);
        }
    }
    if (lane == 31) sh[warp] = f;
    __syncthreads();
    // the (up to 8) warp maps, scanned by the first lanes of every warp
    unsigned w = lane < FLAGS_BLOCK / 32 ? sh[lane] : FN_ID;
#pragma unroll
    for (int o = 1; o < FLAGS_BLOCK / 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w = fn_then(u, w);
    }
    total = __shfl_sync(0xffffffffu, w, FLAGS_BLOCK / 32 - 1);
    const unsigned before = __shfl_sync(0xffffffffu, w, warp > 0 ? warp - 1 : 0);
    __syncthreads();
    return warp > 0 ? fn_then(before, f) : f;
}

// One pass: transition code of every point (what F_t gives from predecessor E(t-1) and from
// predecessor T(t-1)), composed map of the block, state entering the block by a decoupled look-back
// over the maps the earlier blocks published (a map that sends both states to the same one -- any
// block away from the seam -- ends the look-back, so its depth is almost always one), then the
// speculative chain and the ORDERED list of breaks (per-block lists, put together by the block that finishes last).
// A thread owns CH_PPT consecutive points (one composed map per thread goes through the block scan; a
// 10-day track is ~600 blocks: one wave).  Blocks take their index from a ticket, so a block only ever
// waits for blocks that already started.
#define CH_PPT 4
#define CH_PTS (FLAGS_BLOCK * CH_PPT)
#define CH_LB 4        /* blocks per thread in the last block's ordered copy */
#ifndef GZB_CHAIN_MINBLOCKS
#define GZB_CHAIN_MINBLOCKS 4      /* resident blocks of k_chain per SM the compiler must allow: 4 = a cap of 64 registers (natural: 74, 3 blocks); measured with the bulk kernel at 7 blocks: step 199.6 -> 197.0 us, K1 alone 142 -> 140 (gpurun_out/ab_exp*_r2w.log) */
#endif
__global__ void __launch_bounds__(FLAGS_BLOCK, GZB_CHAIN_MINBLOCKS)
k_chain(const GzbGrid g, const NearParams p, const long long nt, const double* __restrict__ yt,
        const double* __restrict__ xt, const long long* __restrict__ G, const double* __restrict__ dG,
        long long* __restrict__ NP, long long* __restrict__ breaks, long long* __restrict__ blist, int* __restrict__ bcount,
        volatile unsigned* agg, unsigned* __restrict__ ticket, unsigned* __restrict__ done_ctr,
        long long* __restrict__ nbrk_out, long long* __restrict__ chg, unsigned long long* __restrict__ nchg) {
    __shared__ unsigned sh[FLAGS_BLOCK / 32];
    __shared__ unsigned last[FLAGS_BLOCK / 32];
    __shared__ int wcnt[FLAGS_BLOCK / 32];
    __shared__ unsigned s_bid, s_entry;
    __shared__ int sE[FLAGS_BLOCK][2], sT[FLAGS_BLOCK][3];      // E / T of the LAST point of every thread
    GZB_KT_MIN(g, 2);
    gzb_pdl_wait();
    GZB_KT_MIN(g, 3);
    if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = s_bid;
    const long long t0 = (long long)bid * CH_PTS + (long long)threadIdx.x * CH_PPT;
    PointSum q[CH_PPT];
#pragma unroll
    for (int k = 0; k < CH_PPT; ++k) {
        q[k].gj = q[k].gi = q[k].ej = q[k].ei = q[k].tj = q[k].ti = -1;
        q[k].tv = 0;
        if (t0 + k < nt) point_sum(g, p, G[t0 + k], dG[t0 + k], q[k]);
    }
    sE[threadIdx.x][0] = q[CH_PPT - 1].ej; sE[threadIdx.x][1] = q[CH_PPT - 1].ei;
    sT[threadIdx.x][0] = q[CH_PPT - 1].tj; sT[threadIdx.x][1] = q[CH_PPT - 1].ti; sT[threadIdx.x][2] = q[CH_PPT - 1].tv;
    __syncthreads();
    GZB_KT_MAX(g, 3);
    unsigned c[CH_PPT];
    unsigned mine = FN_ID;
    {
        int pe0 = 0, pe1 = 0, pt0 = 0, pt1 = 0, ptv = 0;
        // which of my points start a sub-track (point 0 of the call always does)
        unsigned smask = 0;
        int sseg[CH_PPT];
#pragma unroll
        for (int k = 0; k < CH_PPT; ++k) sseg[k] = 0;
        for (int q = 0; q < p.segs.n; ++q) {
            const long long d = p.segs.start[q] - t0;
            if (d >= 0 && d < CH_PPT) {
                smask |= 1u << (int)d;
#pragma unroll
                for (int k = 0; k < CH_PPT; ++k)
                    if (k == (int)d) sseg[k] = q;
            }
        }
        if (t0 < nt) {
            if (smask & 1u) {
            } else if (threadIdx.x > 0) {
                pe0 = sE[threadIdx.x - 1][0]; pe1 = sE[threadIdx.x - 1][1];
                pt0 = sT[threadIdx.x - 1][0]; pt1 = sT[threadIdx.x - 1][1]; ptv = sT[threadIdx.x - 1][2];
            } else {                 // the predecessor belongs to the previous block
                PointSum pq;
                point_sum(g, p, G[t0 - 1], dG[t0 - 1], pq);
                pe0 = pq.ej; pe1 = pq.ei; pt0 = pq.tj; pt1 = pq.ti; ptv = pq.tv;
            }
        }
#pragma unroll
        for (int k = 0; k < CH_PPT; ++k) {
            c[k] = 0;
            if (t0 + k < nt) {
                int o0, o1;
                if (smask & (1u << k)) {       // first point of a sub-track: the predecessor is its seed
                    o0 = o1 = step_code(g, p, (int)p.segs.sj[sseg[k]], (int)p.segs.si[sseg[k]], q[k], yt, xt, t0 + k);
                } else {
                    o0 = step_code(g, p, pe0, pe1, q[k], yt, xt, t0 + k);
                    o1 = ptv ? step_code(g, p, pt0, pt1, q[k], yt, xt, t0 + k) : o0;     // no twin state at t-1: bit unused
                }
                c[k] = (unsigned)(o0 | (o1 << 2));
                const unsigned mk = (unsigned)((o0 == 1) ? 1 : 0) | ((unsigned)((o1 == 1) ? 1 : 0) << 1);     // 2 propagates as 0
                mine = (k == 0) ? mk : fn_then(mine, mk);
            }
            pe0 = q[k].ej; pe1 = q[k].ei; pt0 = q[k].tj; pt1 = q[k].ti; ptv = q[k].tv;
        }
    }
    unsigned total;
    const unsigned incl = block_scan_fn(mine, sh, total);
    if ((threadIdx.x & 31) == 31) last[threadIdx.x >> 5] = incl;
    GZB_KT_MAX(g, 4);
    if (threadIdx.x == 0) {
        agg[bid] = 0x100u | total;               // publish this block's map
        __threadfence();
        unsigned fmap = FN_ID;                   // map from the state entering block k+1 to the state entering this block
        long long k = (long long)bid - 1;
        while (k >= 0) {
            unsigned v;
            while (((v = agg[k]) & 0x100u) == 0u) __nanosleep(40);
            const unsigned fk = v & 3u;
            fmap = fn_then(fk, fmap);
            if (fk == 0u || fk == 3u) break;     // constant map: nothing before block k matters
            --k;
        }
        s_entry = fmap & 1u;                     // the track starts in state 0
    }
    __syncthreads();
    GZB_KT_MAX(g, 5);
    const unsigned prev = __shfl_up_sync(0xffffffffu, incl, 1);
    unsigned before;          // map from the block's entry state to the state entering this thread's first point
    if (threadIdx.x == 0) before = FN_ID;
    else if ((threadIdx.x & 31) == 0) before = last[(threadIdx.x >> 5) - 1];
    else before = prev;
    unsigned st = (before >> s_entry) & 1u;
    unsigned bmask = 0;
#pragma unroll
    for (int k = 0; k < CH_PPT; ++k) {
        if (t0 + k < nt) {
            const unsigned out = (c[k] >> (2 * st)) & 3u;
            long long rj = q[k].ej, ri = q[k].ei;
            if (out == 1u) {
                rj = q[k].tj;
                ri = q[k].ti;
                if (chg) {      // differs from E(t): K2-K4 of this point are redone (gzb_mapping overlap)
                    const unsigned long long cp = atomicAdd(nchg, 1ull);
                    if (GZB_CHK(g, cp < 3ull * (unsigned long long)nt, 4)) chg[cp] = t0 + k;
                }
            }
            reinterpret_cast<longlong2*>(NP)[t0 + k] = make_longlong2(rj, ri);
            if (out == 2u) bmask |= 1u << k;
            st = out == 1u ? 1u : 0u;
        }
    }
    // ---- ordered break list.  Every block writes its own breaks, in order, into its slot of `blist` (CH_PTS entries per
    // block) and its count into `bcount`; the block that finishes LAST adds up the counts and copies the slots, in block
    // order, into `breaks` -- no block ever waits for another one's count (a look-back over counts serialises a grid that
    // runs as a single wave), and the usual few hundred breaks are copied in one step.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mycnt = __popc(bmask);
    int wincl = mycnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += u;
    }
    if (lane == 31) wcnt[warp] = wincl;
    __syncthreads();
    int wbase = 0, btotal = 0;
#pragma unroll
    for (int w2 = 0; w2 < FLAGS_BLOCK / 32; ++w2) {
        if (w2 < warp) wbase += wcnt[w2];
        btotal += wcnt[w2];
    }
    if (bmask) {
        long long pos = (long long)bid * CH_PTS + wbase + (wincl - mycnt);
#pragma unroll
        for (int k = 0; k < CH_PPT; ++k)
            if (bmask & (1u << k)) {
                if (GZB_CHK(g, pos >= (long long)bid * CH_PTS && pos < (long long)(bid + 1) * CH_PTS, 2)) blist[pos] = t0 + k;
                ++pos;
            }
    }
    __syncthreads();
    GZB_KT_MAX(g, 6);
    if (threadIdx.x == 0) {
        bcount[bid] = btotal;
        __threadfence();
        s_entry = (atomicAdd(done_ctr, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_entry) return;
    __threadfence();
    // the last block: exclusive prefix of the counts, then the ordered copy.  A thread owns CH_LB consecutive blocks and
    // loads their counts in one go (independent loads: one round trip for the usual ~600 blocks instead of one per
    // 256 blocks), the block scans the per-thread sums once per super-chunk of FLAGS_BLOCK * CH_LB blocks.
    const long long nblk = gridDim.x;
    long long carry = 0;
    for (long long c0 = 0; c0 < nblk; c0 += (long long)FLAGS_BLOCK * CH_LB) {
        const long long bk0 = c0 + (long long)threadIdx.x * CH_LB;
        int cnt[CH_LB];
        int mysum = 0;
#pragma unroll
        for (int q = 0; q < CH_LB; ++q) {
            cnt[q] = bk0 + q < nblk ? __ldcg(bcount + bk0 + q) : 0;
            mysum += cnt[q];
        }
        int incl2 = mysum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl2, o);
            if (lane >= o) incl2 += u;
        }
        __syncthreads();
        if (lane == 31) wcnt[warp] = incl2;
        __syncthreads();
        int wb2 = 0, ctot = 0;
#pragma unroll
        for (int w2 = 0; w2 < FLAGS_BLOCK / 32; ++w2) {
            if (w2 < warp) wb2 += wcnt[w2];
            ctot += wcnt[w2];
        }
        if (mysum) {
            long long excl = carry + wb2 + (incl2 - mysum);
#pragma unroll
            for (int q = 0; q < CH_LB; ++q) {
                for (int e = 0; e < cnt[q]; ++e)
                    if (GZB_CHK(g, excl + e >= 0 && excl + e < nt && e < CH_PTS, 2)) breaks[excl + e] = __ldcg(blist + (bk0 + q) * CH_PTS + e);
                excl += cnt[q];
            }
        }
        carry += ctot;
    }
    if (threadIdx.x == 0) nbrk_out[0] = carry;
    GZB_KT_MAX(g, 7);
}

// per row: 1 = no cell strictly between a leading overlap cell and its East-West twin has that cell's
// coordinates, and the row is not a pole (where every cell is the same point)
__global__ void __launch_bounds__(256)
k_rowsafe(const GzbGrid g, unsigned char* __restrict__ rowsafe) {
    const long long j = blockIdx.x;
    const int kew = g.kew < 8 ? g.kew : 8;
    __shared__ long long sy[8], sx[8];
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    const longlong2* row = reinterpret_cast<const longlong2*>(g.ll) + j * g.Ni;
    if (threadIdx.x < kew) {
        const longlong2 c = row[threadIdx.x];
        sy[threadIdx.x] = c.x;
        sx[threadIdx.x] = c.y;
    }
    __syncthreads();
    const long long P = g.Ni - g.kew;
    int mybad = 0;
    for (long long i = threadIdx.x; i < g.Ni; i += blockDim.x) {
        const longlong2 c = row[i];
        if (fabs(__longlong_as_double(c.x)) >= 90.0) mybad = 1;
        for (int q = 0; q < kew; ++q)
            if (i > q && i < q + P && c.x == sy[q] && c.y == sx[q]) mybad = 1;
    }
    if (mybad) atomicOr(&bad, 1);
    __syncthreads();
    if (threadIdx.x == 0) rowsafe[j] = (g.kew > 8 || bad) ? 0 : 1;
}

// K1c, step 2: chain replays, validation and write-back in ONE kernel (k_tail).
//
// replay<WRITE>: one warp replays the reference's recurrence R(t) = F_t(R(t-1)) from break k until the
// chain rejoins the speculative one.  The points are taken 32 at a time: lane l prepares point t0+l (unit
// vector, cos(lat), G, E) in parallel, then the steps run in order with those values shuffled in, so that a
// step costs one box_search and nothing else.  WRITE=false records where the replay rejoins the speculative
// chain and logs every replayed value tagged with the replay's id (one 64-bit word per point: id in the high
// 27 bits, flat index + 1 in the low 36); WRITE=true replays again up to the recorded stop and writes NP
// (only needed when another, discarded, replay overwrote part of this one's log).
#define WLOG_TAG(k) ((unsigned long long)(((k) + 1) & 0x7ffffffLL) << 36)

template <bool WRITE>
__device__ __forceinline__ void replay(const GzbGrid& g, const NearParams& p, WarpSm& w, const int lane, const long long nt,
                                       const double* __restrict__ yt, const double* __restrict__ xt,
                                       const long long* __restrict__ G, const double* __restrict__ dG, const long long k,
                                       const long long b, long long tend, long long* __restrict__ stop,
                                       long long* __restrict__ first, unsigned long long* __restrict__ wlog,
                                       long long* __restrict__ NP, unsigned long long* __restrict__ steps,
                                       long long* __restrict__ chg, unsigned long long* __restrict__ nchg) {
    long long pj, pi;
    long long seg_end;
    const int sg = seg_of(p.segs, b, nt, seg_end);
    if (!WRITE) tend = seg_end;          // a replay never runs into the next sub-track
    if (b == p.segs.start[sg]) {
        pj = p.segs.sj[sg];
        pi = p.segs.si[sg];
    } else {                 // the speculative chain at b-1 (k_chain; an earlier valid replay ending there rejoined it)
        pj = __ldcg(NP + 2 * (b - 1));
        pi = __ldcg(NP + 2 * (b - 1) + 1);
    }
    unsigned long long n = 0;
    bool done = false;
    // the first round prepares 4 points only: most replays rejoin the speculative chain after one or two steps
    int width = 4;
    for (long long t0 = b; !done; t0 += width, width = 32) {
        const int nq = (int)((tend - t0 + 1 < width) ? (tend - t0 + 1) : width);
        Owner me;
        float mpx = 0.f, mpy = 0.f, mpz = 0.f;
        me.plat = me.plon = me.cpl = 0.0;
        long long gf = GZB_NOFLAT, ej = -1, ei = -1, sj = -1, si = -1, tj = -1, ti = -1;
        int tv = 0;
        if (lane < nq) {
            owner_point(yt[t0 + lane], xt[t0 + lane], me, mpx, mpy, mpz);
            gf = G[t0 + lane];
            const double gd = dG[t0 + lane];
            accept_edge(g, gf, gd, p.thr2, ej, ei, p.edge);
            long long tw = 0;
            if (twin_of(g, p, gf, gd, tw)) {          // T(t): see k_chain
                tv = 1;
                accept_edge(g, tw, gd, INFINITY, tj, ti, p.edge);
            }
            if (!WRITE) {                             // the speculative chain, to detect where the replay rejoins it
                sj = __ldcg(NP + 2 * (t0 + lane));
                si = __ldcg(NP + 2 * (t0 + lane) + 1);
            }
        }
        __syncwarp();
        w.bd[lane] = ~0ull;
        w.bf[lane] = GZB_NOFLAT;
        w.pb2[lane] = INFINITY;
        __syncwarp();
        for (int q = 0; q < nq; ++q) {
            const long long t = t0 + q;
            const long long qej = __shfl_sync(0xffffffffu, ej, q), qei = __shfl_sync(0xffffffffu, ei, q);
            const long long qgf = __shfl_sync(0xffffffffu, gf, q);
            long long rj = qej, ri = qei;
            if (pj != 0 && pi != 0) {     // Python truthiness of (j_prv and i_prv)
                long long j0, j1, i0, i1;
                if (clamp_box(g, pj, pi, p.r, j0, j1, i0, i1)) {
                    // the global nearest point inside the box: the box pass finds that very point,
                    // so R(t) = E(t) whatever its distance (DESIGN.md, K1 key fact) -- no search
                    bool inside = false;
                    if (qgf != GZB_NOFLAT) {
                        long long gj, gi;
                        gzb_split(g, qgf, gj, gi);
                        inside = gj >= j0 && gj < j1 && gi >= i0 && gi < i1;
                    }
                    bool twin_in = false;
                    if (!inside && __shfl_sync(0xffffffffu, tv, q)) {
                        long long gj, gi;
                        gzb_split(g, qgf, gj, gi);
                        gi += g.Ni - g.kew;
                        twin_in = gj >= j0 && gj < j1 && gi >= i0 && gi < i1;
                    }
                    if (twin_in) {       // the box pass finds the twin of G(t), at the global minimum distance (< r0)
                        rj = __shfl_sync(0xffffffffu, tj, q);
                        ri = __shfl_sync(0xffffffffu, ti, q);
                    } else if (!inside) {
                        Srch s;
                        srch_init(s, __shfl_sync(0xffffffffu, mpx, q), __shfl_sync(0xffffffffu, mpy, q),
                                  __shfl_sync(0xffffffffu, mpz, q));
                        set_bound(s, p.prox_r0);
                        if (lane == 0) w.pb2[q] = s.B2;
                        __syncwarp();
                        int qn = 0;
                        box_search(g, w, lane, s, q, j0, j1, i0, i1, qn, me);
                        flush_queue(g, w, lane, qn, me);
                        const long long f = w.bf[q];
                        const double d = __longlong_as_double((long long)w.bd[q]);
                        if (f != GZB_NOFLAT && d < p.r0) accept_edge(g, f, d, INFINITY, rj, ri, p.edge);
                    }
                }
            }
            ++n;
            if (WRITE) {
                if (lane == 0 && GZB_CHK(g, t >= 0 && t < nt, 8)) {
                    NP[2 * t] = rj;
                    NP[2 * t + 1] = ri;
                    if (chg) {
                        const unsigned long long cp = atomicAdd(nchg, 1ull);
                        if (GZB_CHK(g, cp < 3ull * (unsigned long long)nt, 4)) chg[cp] = t;
                    }
                }
                if (t >= tend) { done = true; break; }
            } else {
                if (lane == 0 && GZB_CHK(g, t >= 0 && t < nt, 8)) {
                    wlog[t] = WLOG_TAG(k) | (unsigned long long)(rj < 0 ? 0 : rj * g.Ni + ri + 1);
                    if (t == b) {
                        first[2 * k] = rj;
                        first[2 * k + 1] = ri;
                    }
                }
                if ((rj == __shfl_sync(0xffffffffu, sj, q) && ri == __shfl_sync(0xffffffffu, si, q)) || t >= tend) {
                    if (lane == 0) stop[k] = t;
                    done = true;
                    break;
                }
            }
            pj = rj;
            pi = ri;
        }
        __syncwarp();
    }
    if (!WRITE && lane == 0) {
        atomicAdd(steps, n);
        atomicMax(steps + 5, n);      // longest replay
    }
}

// k_tail: (1) every warp of the grid replays its breaks (WRITE=false); (2) the block that finishes LAST (a counter
// behind a __threadfence) decides which replays are valid -- a replay is valid iff it does not start inside an earlier
// valid replay, whose deviating values would have been its true predecessor: an inherently ordered scan over the break
// list, 32 breaks per step on the common path where no replay reaches the next break -- and (3) writes the valid
// replays back: a THREAD per break copies its (short) log into NP and appends the changed points to `chg`; replays
// that are long, or whose log another (discarded) replay overwrote, are redone warp per break (rare).
// One launch instead of compaction + replay + validation + write-back: after k_chain the chain part of K1 is pure
// dependent latency, and every launch boundary in it costs as much as the work between two of them.
#define TAIL_WARPS 4
#define TAIL_THREADS (32 * TAIL_WARPS)
#define VALID_CHUNK 512
#define TAIL_SHORT 8
__global__ void __launch_bounds__(TAIL_THREADS)
k_tail(const GzbGrid g, const NearParams p, const long long nt, const double* __restrict__ yt,
       const double* __restrict__ xt, const long long* __restrict__ G, const double* __restrict__ dG,
       const long long* __restrict__ breaks, long long* __restrict__ scal,
       long long* __restrict__ stop, long long* __restrict__ first, unsigned char* __restrict__ valid,
       unsigned long long* __restrict__ wlog, long long* __restrict__ NP,
       long long* __restrict__ chg, unsigned long long* __restrict__ nchg, unsigned* __restrict__ done_ctr,
       long long* __restrict__ stats_out) {
    __shared__ WarpSm ws[TAIL_WARPS];
    __shared__ long long sb[VALID_CHUNK + 1];
    __shared__ long long ss[VALID_CHUNK];
    __shared__ unsigned char sv[VALID_CHUNK];      // verdicts of the chunk: 1 valid
    __shared__ short s2[VALID_CHUNK];              // entries of the chunk whose replay is redone by a warp
    __shared__ int s_n2;
    __shared__ int s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GZB_KT_MIN(g, 8);
    gzb_pdl_wait();
    GZB_KT_MIN(g, 9);
    const long long nb = scal[0];
    unsigned long long* steps = (unsigned long long*)(scal + 1);
    // ---- (1) replays
    {
        const long long nwarps = (long long)gridDim.x * TAIL_WARPS;
        for (long long k = (long long)blockIdx.x * TAIL_WARPS + warp; k < nb; k += nwarps)
            replay<false>(g, p, ws[warp], lane, nt, yt, xt, G, dG, k, breaks[k], nt - 1, stop, first, wlog, NP, steps, nullptr, nullptr);
    }
    __syncthreads();
    GZB_KT_MAX(g, 9);
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(done_ctr, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- (2) validation and (3) write-back, VALID_CHUNK breaks at a time.  The block stages the chunk's breaks and stops in
    // shared memory; warp 0 decides validity (an ordered scan); then a THREAD per break copies its (short, intact) log into
    // NP -- breaks, stops and verdicts come from shared memory, so a write-back costs two dependent round trips (the log,
    // the slot in `chg`) -- and the replays that are long, or whose log another (discarded) replay overwrote, are listed in
    // shared memory and redone warp per break (rare).  Validity of a chunk is final once the chunk has been scanned.
    long long cur_end = -1;
    for (long long c0 = 0; c0 < nb; c0 += VALID_CHUNK) {
        const int cn = (int)((nb - c0 < VALID_CHUNK) ? (nb - c0) : VALID_CHUNK);
        for (int e = threadIdx.x; e <= cn; e += blockDim.x) {
            sb[e] = (c0 + e < nb) ? breaks[c0 + e] : GZB_NOFLAT;
            if (e < cn) ss[e] = __ldcg(stop + c0 + e);
        }
        if (threadIdx.x == 0) s_n2 = 0;
        __syncthreads();
        if (threadIdx.x < 32) {
            for (int base = 0; base < cn; base += 32) {
                const int e = base + lane;
                const bool in = e < cn;
                const long long b = in ? sb[e] : GZB_NOFLAT;
                const long long s = in ? ss[e] : GZB_NOFLAT;
                const long long nxt = in ? sb[e + 1] : GZB_NOFLAT;
                const bool simple = !in || (s < nxt);
                const long long b0 = __shfl_sync(0xffffffffu, b, 0);
                if (__all_sync(0xffffffffu, simple) && b0 > cur_end) {
                    if (in) sv[e] = 1;
                    const int lastl = ((cn - base < 32) ? (cn - base) : 32) - 1;
                    cur_end = __shfl_sync(0xffffffffu, s, lastl);
                } else {
                    for (int l = 0; l < 32; ++l) {
                        const long long kb = __shfl_sync(0xffffffffu, b, l);
                        const long long ks = __shfl_sync(0xffffffffu, s, l);
                        if (base + l < cn) {
                            const bool v = kb > cur_end;
                            if (v) cur_end = ks;
                            if (lane == l) sv[e] = v ? 1 : 0;
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (c0 + VALID_CHUNK >= nb) GZB_KT_MAX(g, 10);
        for (int e = threadIdx.x; e < cn; e += blockDim.x) {
            const long long k = c0 + e;
            valid[k] = sv[e];
            if (!sv[e]) continue;
            const long long b = sb[e], en = ss[e];
            if (!GZB_CHK(g, b >= 0 && b < nt && en >= b && en < nt, 8)) continue;
            if (en == b) {
                const longlong2 f = __ldcg(reinterpret_cast<const longlong2*>(first) + k);
                reinterpret_cast<longlong2*>(NP)[b] = f;
                if (chg) {
                    const unsigned long long cp = atomicAdd(nchg, 1ull);
                    if (GZB_CHK(g, cp < 3ull * (unsigned long long)nt, 4)) chg[cp] = b;
                }
                continue;
            }
            bool ok = (en - b) < TAIL_SHORT;
            unsigned long long v[TAIL_SHORT];
            if (ok) {
#pragma unroll
                for (int q = 0; q < TAIL_SHORT; ++q) v[q] = (b + q <= en) ? __ldcg(wlog + b + q) : 0ull;
#pragma unroll
                for (int q = 0; q < TAIL_SHORT; ++q)
                    if (b + q <= en && (v[q] >> 36) != (WLOG_TAG(k) >> 36)) ok = false;
            }
            if (!ok) {
                valid[k] = 2;
                s2[atomicAdd(&s_n2, 1)] = e;
                continue;
            }
            const int cnt = (int)(en - b + 1);
            const unsigned long long base = chg ? atomicAdd(nchg, (unsigned long long)cnt) : 0ull;
#pragma unroll
            for (int q = 0; q < TAIL_SHORT; ++q) {
                if (b + q <= en) {
                    const long long f = (long long)(v[q] & 0xfffffffffULL) - 1;
                    long long j = -1, i = -1;
                    if (f >= 0) gzb_split(g, f, j, i);
                    reinterpret_cast<longlong2*>(NP)[b + q] = make_longlong2(j, i);
                    if (chg && GZB_CHK(g, base + q < 3ull * (unsigned long long)nt, 4)) chg[base + q] = b + q;
                }
            }
        }
        __syncthreads();
        if (c0 + VALID_CHUNK >= nb) GZB_KT_MAX(g, 11);
        const int n2 = s_n2;
        for (int q = warp; q < n2; q += TAIL_WARPS) {
            const int e = s2[q];
            replay<true>(g, p, ws[warp], lane, nt, yt, xt, G, dG, c0 + e, sb[e], ss[e], stop, first, wlog, NP, steps, chg, nchg);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {      // (breaks, replay steps, ...) of this call, for gzb_get_stats
        stats_out[0] = scal[0];
        stats_out[1] = (long long)__ldcg(steps);
        for (int q = 2; q < 7; ++q) stats_out[q] = __ldcg(scal + q);   // k_near_search: unresolved (total, no seed, not converged, not proven); longest replay
    }
    GZB_KT_MAX(g, 12);
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

// one thread that lets `ns` nanoseconds pass (see chain_first in gzb_nearest_dev)
__global__ void k_delay(const unsigned ns) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < (unsigned long long)ns);
}

void gzb_carveout_nearest(int pct) {
    const int v = pct < 0 ? (int)cudaSharedmemCarveoutDefault : pct;
    cudaFuncSetAttribute(k_near_search, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(k_near_slow, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(k_chain, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(k_tail, cudaFuncAttributePreferredSharedMemoryCarveout, v);
}

// ------------------------------------------------------------------------------------------
size_t gzb_nearest_ws_bytes(int64_t nt) {
    const size_t n = (size_t)(nt > 0 ? nt : 1);
    const size_t nblk = (n + CH_PTS - 1) / CH_PTS;
    return gzb_align(n * 8) * 2      // G, dG
           + gzb_align((nblk + 3) * 4)   // published block maps + ticket + the done counters of k_chain and k_tail
           + gzb_align(nblk * 4)     // break counts of k_chain's blocks
           + gzb_align(nblk * CH_PTS * 8)   // ... and their break lists
           + gzb_align(n * 8) * 2    // breaks, stop
           + gzb_align(n * 16)       // first
           + gzb_align(n)            // valid
           + gzb_align(n * 8)        // replay log
           + gzb_align(n * 8) + gzb_align(n * 4)   // unresolved list, hints
           + gzb_align(n * 24)       // points whose nearest point differs from E(t): twin states + replay writes
           + gzb_align(64) * 2;      // scalars, corner
}

// device-level driver: all pointers are device pointers, work is enqueued on ctx->stream
// With `split`, the chain part (K1c: scan, compaction, replays) is enqueued on ctx->aux_stream
// behind an event, so that the caller can run K2+K3 on the speculative chain meanwhile; the
// points the replays changed are listed in *chg_out / *nchg_out (device pointers) and ctx->ev_k1[1]
// marks the end of the chain part.
// launch on `s`, optionally as a programmatic dependent of the previous kernel in that stream (option "pdl")
template <typename... P, typename... A>
static void launch_dep(bool pdl, void (*kern)(P...), unsigned grid, unsigned block, cudaStream_t s, A&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.stream = s;
    cudaLaunchAttribute at;
    memset(&at, 0, sizeof(at));
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);     // errors surface in GZB_LAUNCH_CHECK
}

int gzb_nearest_dev(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, double rd_found_km,
                    int np_box_r, int64_t seed_j, int64_t seed_i, int64_t* np_out, bool split,
                    const long long** chg_out, const unsigned long long** nchg_out, GzbGlobalNearest* gn_out) {
    if (nt <= 0) return GZB_OK;
    if (np_box_r < 0) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_nearest: np_box_r must be >= 0");
    gzb_apply_carveout(ctx);
    const size_t n = (size_t)nt;
    const long long nblk = (long long)((n + CH_PTS - 1) / CH_PTS);
    long long* G = (long long*)gzb_arena_take(ctx, n * 8);
    double* dG = (double*)gzb_arena_take(ctx, n * 8);
    unsigned* agg = (unsigned*)gzb_arena_take(ctx, ((size_t)nblk + 3) * 4);      // published block maps + the ticket + the done counters of k_chain / k_tail
    int* bcount = (int*)gzb_arena_take(ctx, (size_t)nblk * 4);
    long long* blist = (long long*)gzb_arena_take(ctx, (size_t)nblk * CH_PTS * 8);
    long long* breaks = (long long*)gzb_arena_take(ctx, n * 8);
    long long* stop = (long long*)gzb_arena_take(ctx, n * 8);
    long long* first = (long long*)gzb_arena_take(ctx, n * 16);
    unsigned char* valid = (unsigned char*)gzb_arena_take(ctx, n);
    unsigned long long* wlog = (unsigned long long*)gzb_arena_take(ctx, n * 8);
    long long* ulist = (long long*)gzb_arena_take(ctx, n * 8);
    unsigned* hint = (unsigned*)gzb_arena_take(ctx, n * 4);
    long long* chg = (long long*)gzb_arena_take(ctx, n * 24);
    long long* scal = (long long*)gzb_arena_take(ctx, 64);
    if (!scal) return gzb_fail(ctx, GZB_ERR_NOMEM, "gzb_nearest: workspace too small");
    unsigned long long* nchg = (unsigned long long*)(scal + 7);
    if (!split) chg = nullptr;
    if (chg_out) *chg_out = chg;
    if (nchg_out) *nchg_out = nchg;
    if (gn_out) {
        gn_out->G = G; gn_out->dG = dG; gn_out->thr2 = 0.0; gn_out->edge = ctx->edge_exclusion;
        gn_out->ulist = ulist; gn_out->ucount = (const unsigned long long*)(scal + 2);
    }
    double* corner = (double*)(ctx->d_flags + 32);      // persistent: 4 doubles, keyed by np_box_r
    cudaStream_t s = ctx->stream;

    NearParams p;
    volatile double r0 = rd_found_km;
    volatile double t1 = 1.2 * r0;
    volatile double t2 = 1.2 * t1;
    p.r0 = r0;
    p.thr2 = t2;
    if (gn_out) gn_out->thr2 = t2;
    const double half = r0 / (2.0 * 6360.0);
    const double sn = sin(half < 1.5 ? half : 1.5);
    p.prox_r0 = (float)(4.0 * sn * sn * (1.0 + 1.0e-5) + 1.0e-20);
    p.chord_r0 = 2.0 * sn * (1.0 + 1.0e-6) + 1.0e-12;
    p.r = np_box_r;
    p.edge = ctx->edge_exclusion;
    if (ctx->segs_next.n > 0) {          // gzb_map_interp_multi: several chains, consumed by this call
        p.segs = ctx->segs_next;
        ctx->segs_next.n = 0;
    } else {
        p.segs.n = 1;
        p.segs.start[0] = 0;
        p.segs.sj[0] = seed_j;
        p.segs.si[0] = seed_i;
    }
    p.corner = corner;
    p.rowsafe = nullptr;
    if (ctx->twin_chain && ctx->g.kew > 0 && ctx->g.kew < ctx->g.Ni / 2) {
        if (!ctx->d_rowsafe) GZB_CUDA(ctx, gzb_pool_malloc((void**)&ctx->d_rowsafe, (size_t)ctx->g.Nj));
        if (ctx->rowsafe_kew != ctx->g.kew) {
            k_rowsafe<<<(unsigned)ctx->g.Nj, 256, 0, ctx->stream>>>(ctx->g, ctx->d_rowsafe);
            GZB_LAUNCH_CHECK(ctx);
            ctx->rowsafe_kew = ctx->g.kew;
        }
        p.rowsafe = ctx->d_rowsafe;
    }

    GZB_CUDA(ctx, cudaMemsetAsync(scal, 0, 64, s));
    GZB_CUDA(ctx, cudaMemsetAsync(agg, 0, ((size_t)nblk + 3) * sizeof(unsigned), s));     // k_chain's ticket and block maps, the two done counters
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
    cudaStream_t sb = s;
    bool sb_has_kernel = false;      // is the operation right before k_chain on its stream a kernel?  (a dependent launch needs one)
    gzb_trace(ctx, s, "K1 start");
    if (ctx->nearest_path && ctx->g.cellv && ctx->g.lut.bins) {
#if GZB_EXP_SKIP_DG
        const float cthr0 = half < 1.5 ? (float)(2.0 * sn) : NAN;
        const float cthr1 = (t2 / (2.0 * 6360.0)) < 1.5 ? (float)(2.0 * sin(t2 / (2.0 * 6360.0))) : NAN;
        k_near_search<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ctx->g, nt, yt, xt, G, dG, ulist, hint,
                                                                 (unsigned long long*)(scal + 2), cthr0, cthr1);
#else
        k_near_search<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ctx->g, nt, yt, xt, G, dG, ulist, hint,
                                                                 (unsigned long long*)(scal + 2));
#endif
        GZB_LAUNCH_CHECK(ctx);
        // from here on everything is latency-bound (the warp search of the few unresolved points, then
        // the chain): with `split` it goes to the auxiliary stream and the caller runs K2-K4 meanwhile on
        // E(t) = the plain global answers, which it derives from G / dG itself, skipping the unresolved points
        if (split) {
            sb = ctx->aux_stream;
            GZB_CUDA(ctx, cudaEventRecord(ctx->ev_k1[0], s));
            GZB_CUDA(ctx, cudaStreamWaitEvent(sb, ctx->ev_k1[0], 0));
            // The bulk kernel the caller launches next on `s` and the warp search below become eligible at the same moment,
            // when the search kernel ends; whichever the GPU starts first fills it, and when that is the bulk kernel (same
            // stream as the search: no cross-stream hop) the chain part starts ~25 us late (in-kernel time marks, KTIME
            // build).  So `s` is made to wait for an event that the auxiliary stream records once it has passed ITS wait:
            // the bulk kernel then needs two hops, the warp search one.
            if (ctx->chain_first == 1) {
                GZB_CUDA(ctx, cudaEventRecord(ctx->ev_k1[2], sb));
                GZB_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_k1[2], 0));
            } else if (ctx->chain_first >= 2 && nt >= 100000) {
                // ... which still leaves the two launches racing (measured).  chain_first = n >= 2: a one-thread kernel on a
                // third stream sleeps n microseconds after the search kernel's end before `s` is released: the warp search's
                // blocks are resident by then and the bulk kernel fills the SMs as they retire.
                GZB_CUDA(ctx, cudaStreamWaitEvent(ctx->push_stream, ctx->ev_k1[0], 0));
                k_delay<<<1, 1, 0, ctx->push_stream>>>((unsigned)ctx->chain_first * 1000u);
                GZB_LAUNCH_CHECK(ctx);
                GZB_CUDA(ctx, cudaEventRecord(ctx->ev_k1[2], ctx->push_stream));
                GZB_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_k1[2], 0));
            }
        }
        gzb_trace(ctx, s, "search end (main)");
        k_near_slow<<<(unsigned)(nsm * 32 / SLOW_WARPS), 32 * SLOW_WARPS, 0, sb>>>(ctx->g, yt, xt, G, dG, ulist, hint,
                                                        (const unsigned long long*)(scal + 2));
        GZB_LAUNCH_CHECK(ctx);
        sb_has_kernel = true;
        gzb_trace(ctx, sb, "slow end");
    } else {
        k_global<<<(unsigned)gblocks, 256, 0, s>>>(ctx->g, nt, yt, xt, G, dG, K);
        GZB_LAUNCH_CHECK(ctx);
        sb_has_kernel = !split;
        if (split) {
            sb = ctx->aux_stream;
            GZB_CUDA(ctx, cudaEventRecord(ctx->ev_k1[0], s));
            GZB_CUDA(ctx, cudaStreamWaitEvent(sb, ctx->ev_k1[0], 0));
        }
    }
    // the chain kernels follow each other on one stream: with "pdl" each is a programmatic dependent of the one
    // before it (gzb_pdl_wait() at its top), so its launch overlaps the tail of its predecessor
    const bool pdl = ctx->pdl != 0 && !ctx->trace;
    launch_dep(pdl && sb_has_kernel, k_chain, (unsigned)nblk, FLAGS_BLOCK, sb, ctx->g, p, (long long)nt, yt, xt, (const long long*)G,
               (const double*)dG, (long long*)np_out, breaks, blist, bcount, (volatile unsigned*)agg, agg + nblk, agg + nblk + 2,
               scal, chg, nchg);
    GZB_LAUNCH_CHECK(ctx);
    gzb_trace(ctx, sb, "chain kernel end");
    // enough warps for the usual few hundred breaks (the replay loop strides over the rest): every block of the grid,
    // empty or not, has to pass the "who is last" counter
    long long wblocks = (long long)((n + TAIL_WARPS - 1) / TAIL_WARPS);
    const long long wmax = (long long)nsm * 2;
    if (wblocks > wmax) wblocks = wmax;
    launch_dep(pdl, k_tail, (unsigned)wblocks, TAIL_THREADS, sb, ctx->g, p, (long long)nt, yt, xt, (const long long*)G,
               (const double*)dG, (const long long*)breaks, scal, stop, first, valid, wlog, (long long*)np_out, chg, nchg,
               agg + nblk + 1, (long long*)(ctx->d_flags + 8));
    GZB_LAUNCH_CHECK(ctx);
    gzb_trace(ctx, sb, "tail end (chain done)");
    if (split) GZB_CUDA(ctx, cudaEventRecord(ctx->ev_k1[1], sb));
    ctx->nearest_stats_pending = 1;      // k_valid left (breaks, replay steps) in the persistent flags
    return GZB_OK;
}
