// K1: nearest model-grid point for every track point, with the reference's sequential
// "previous hit seeds the next search box" semantics reproduced exactly
// (gonzag/bilin_mapping.py:144-191 NearestPoint, :566-593 BilinTrack.nrpt).
//
// Decomposition:
//   K1b k_global   - exact global first-index haversine argmin G(t) for every point, one warp
//                    per point: best-first descent of the bounding-sphere pyramid (coarse to
//                    fine), 32 nodes / 32 cells per step, warp REDUX for the minima; only
//                    cells whose chord^2 proxy ties the best one are evaluated with the
//                    reference's own haversine formula in fp64.
//   K1c k_flags    - speculative chain: R(t) = E(t) (edge-excluded, threshold-accepted G(t))
//                    whenever G(t) lies inside the search box of E(t-1); the other points are
//                    "breaks".
//        k_walk    - one warp per break replays the reference's recurrence
//                    R(t) = F_t(R(t-1)) with the true box-restricted search (K1a: the
//                    (2r+1)^2 window, clamped not wrapped, accepted iff d < r0) until the
//                    chain rejoins the speculative one; k_valid discards replays whose own
//                    starting assumption was overwritten by an earlier replay.
#include "gzb_common.cuh"

#define GZB_NOFLAT 0x7fffffffffffffffLL

struct WarpScratch {
    int fJ[GZB_MAXLEV + 1];
    int fI[GZB_MAXLEV + 1];
    unsigned fmask[GZB_MAXLEV + 1];
    int pad;
    long long cflat[32];
    double cprox[32];
};

struct NearParams {
    double r0;         // rd_found_km
    double thr2;       // fl(1.2*fl(1.2*r0)): acceptance radius of the last useful global pass
    double prox_r0;    // chord^2 bound that contains every cell with d < r0
    double chord_r0;   // chord length of r0 (with margin)
    int r;             // np_box_r
    long long seed_j, seed_i;
    const double* corner;   // bounding sphere (cx,cy,cz,rho) of the box of predecessor (-1,-1)
};

// ------------------------------------------------------------------------------------------
// Warp-cooperative exact nearest-cell search over the whole grid or a clamped index box.
// Returns the minimum reference distance and, among equal distances, the smallest row-major
// flat index (numpy.argmin's first occurrence).  `bound0` is an upper bound on the chord^2 of
// acceptable cells (+inf for the global search).
template <bool BOXED>
__device__ __forceinline__ void warp_search(const GzbGrid& g, WarpScratch& ws, const int lane,
                                            const double plat, const double plon,
                                            const long long j0, const long long j1,
                                            const long long i0, const long long i1,
                                            const double bound0, double& out_d, long long& out_flat) {
    const double la = plat * GZB_D2R, lo = plon * GZB_D2R;
    const double cpl = cos(la);
    const double px = cpl * cos(lo), py = cpl * sin(lo), pz = sin(la);
    double best = bound0;
    double bnd = best * (1.0 + 1.0e-7) + 1.0e-24;
    double my_d = INFINITY;
    long long my_flat = GZB_NOFLAT;
    int ncand = 0;

    auto flush = [&]() {
        if (lane < ncand && ws.cprox[lane] <= bnd) {
            const long long f = ws.cflat[lane];
            const double d = gzb_haversine_dev(plat, plon, cpl, g.lat[f], g.lon[f]);
            if (d < my_d || (d == my_d && f < my_flat)) { my_d = d; my_flat = f; }
        }
        ncand = 0;
        __syncwarp();
    };

    int depth = 0;
    if (lane == 0) ws.fmask[0] = 0xffffffffu;
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
            J = ws.fJ[depth];
            I = ws.fI[depth];
            base = ((long long)J * P.nI + I) * 32;
            const int a = lane / P.bi;
            cJ = J * P.bj + a;
            cI = I * P.bi + (lane - a * P.bi);
        }
        const double4 nd = L.nodes[base + lane];
        bool keep = ((ws.fmask[depth] >> lane) & 1u) && (nd.w >= 0.0);
        if (BOXED && keep) {
            const long long a0 = (long long)cJ * L.sJ, b0 = (long long)cI * L.sI;
            keep = (a0 < j1) && (a0 + L.sJ > j0) && (b0 < i1) && (b0 + L.sI > i0);
        }
        double lb2 = 0.0;
        if (keep) {
            const double dx = px - nd.x, dy = py - nd.y, dz = pz - nd.z;
            const double lb = sqrt(dx * dx + dy * dy + dz * dz) - nd.w;
            lb2 = lb > 0.0 ? lb * lb : 0.0;
            keep = lb2 <= bnd;
        }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m == 0u) {
            --depth;
            continue;
        }
        const unsigned key = keep ? ((__float_as_uint((float)lb2) & 0xffffffe0u) | (unsigned)lane) : 0xffffffffu;
        const int c = (int)(__reduce_min_sync(0xffffffffu, key) & 31u);
        m &= ~(1u << c);
        const int sJ = __shfl_sync(0xffffffffu, cJ, c);
        const int sI = __shfl_sync(0xffffffffu, cI, c);
        __syncwarp();
        if (lane == 0) ws.fmask[depth] = m;
        if (cl > 0) {
            ++depth;
            if (lane == 0) {
                ws.fJ[depth] = sJ;
                ws.fI[depth] = sI;
                ws.fmask[depth] = 0xffffffffu;
            }
            __syncwarp();
        } else {
            // leaf tile (sJ, sI): 32 cells, one per lane
            const double* __restrict__ cc = g.cells + ((long long)sJ * L.nI + sI) * 96;
            const double x = cc[lane], y = cc[32 + lane], z = cc[64 + lane];
            const long long j = (long long)sJ * GZB_TILE_J + (lane >> 3);
            const long long i = (long long)sI * GZB_TILE_I + (lane & 7);
            bool ok = (j < g.Nj) && (i < g.Ni);
            if (BOXED) ok = ok && (j >= j0) && (j < j1) && (i >= i0) && (i < i1);
            const double dx = px - x, dy = py - y, dz = pz - z;
            const double prox = dx * dx + dy * dy + dz * dz;
            const double pm = gzb_warp_min_nonneg(ok ? prox : INFINITY);
            if (pm < best) {
                best = pm;
                bnd = best * (1.0 + 1.0e-7) + 1.0e-24;
            }
            const bool cand = ok && (prox <= bnd);
            const unsigned cm = __ballot_sync(0xffffffffu, cand);
            const int nnew = __popc(cm);
            if (nnew) {
                if (ncand + nnew > 32) flush();
                if (cand) {
                    const int pos = ncand + __popc(cm & ((1u << lane) - 1u));
                    ws.cflat[pos] = j * g.Ni + i;
                    ws.cprox[pos] = prox;
                }
                ncand += nnew;
            }
            __syncwarp();
        }
    }
    flush();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double d2 = __shfl_xor_sync(0xffffffffu, my_d, o);
        const long long f2 = __shfl_xor_sync(0xffffffffu, my_flat, o);
        if (d2 < my_d || (d2 == my_d && f2 < my_flat)) { my_d = d2; my_flat = f2; }
    }
    out_d = my_d;
    out_flat = my_flat;
}

// ------------------------------------------------------------------------------------------
// K1b: global nearest point of every track point.
__global__ void __launch_bounds__(256)
k_global(const GzbGrid g, const long long nt, const double* __restrict__ yt, const double* __restrict__ xt,
         long long* __restrict__ G, double* __restrict__ dG) {
    __shared__ WarpScratch ws[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long base = (long long)blockIdx.x * 32; base < nt; base += (long long)gridDim.x * 32) {
        for (int q = warp; q < 32; q += 8) {
            const long long t = base + q;
            if (t >= nt) break;
            double d;
            long long f;
            warp_search<false>(g, ws[warp], lane, yt[t], xt[t], 0, g.Nj, 0, g.Ni, INFINITY, d, f);
            if (lane == 0) {
                G[t] = f;
                dG[t] = d;
            }
        }
    }
}

// threshold acceptance of the global hit + the edge exclusion of BilinTrack.nrpt
// (gonzag/bilin_mapping.py:169-190 and :582-585; column 0 is always excluded)
__device__ __forceinline__ void accept_edge(const GzbGrid& g, const long long flat, const double d,
                                            const double thr, long long& j, long long& i) {
    if (flat != GZB_NOFLAT && d < thr) {
        j = flat / g.Ni;
        i = flat - j * g.Ni;
    } else {
        j = -1;
        i = -1;
    }
    if (i == 0 || (i == g.Ni - 1 && g.kew == -1)) { j = -1; i = -1; }
    if (j == 0 || j == g.Nj - 1) { j = -1; i = -1; }
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
        accept_edge(g, G[t], dG[t], p.thr2, ej, ei);
        NP[2 * t] = ej;
        NP[2 * t + 1] = ei;
        long long pj, pi;
        if (t == 0) {
            pj = p.seed_j;
            pi = p.seed_i;
        } else {
            accept_edge(g, G[t - 1], dG[t - 1], p.thr2, pj, pi);
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
            long long w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            wsum[lane] = w;
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
        int w = wcnt[lane];
        int s = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        wcnt[lane] = s - w;
    }
    __syncthreads();
    if (b) breaks[offsets[blockIdx.x] + wcnt[warp] + __popc(m & ((1u << lane) - 1u))] = t;
}

// one step of the reference recurrence: R(t) = F_t(pred)
__device__ __forceinline__ void chain_step(const GzbGrid& g, WarpScratch& ws, const int lane, const NearParams& p,
                                           const long long t, const long long pj, const long long pi,
                                           const double* __restrict__ yt, const double* __restrict__ xt,
                                           const long long ej, const long long ei, long long& rj, long long& ri) {
    rj = ej;
    ri = ei;
    if (pj != 0 && pi != 0) {
        long long j0, j1, i0, i1;
        if (clamp_box(g, pj, pi, p.r, j0, j1, i0, i1)) {
            double d;
            long long f;
            warp_search<true>(g, ws, lane, yt[t], xt[t], j0, j1, i0, i1, p.prox_r0, d, f);
            if (f != GZB_NOFLAT && d < p.r0) accept_edge(g, f, d, INFINITY, rj, ri);
        }
    }
}

// K1c, step 2/4: chain replay, one warp per break.  WRITE=false records where the replay
// rejoins the speculative chain; WRITE=true (valid replays only) stores the results.
template <bool WRITE>
__global__ void __launch_bounds__(256)
k_walk(const GzbGrid g, const NearParams p, const long long nt, const double* __restrict__ yt,
       const double* __restrict__ xt, const long long* __restrict__ G, const double* __restrict__ dG,
       const long long* __restrict__ breaks, const long long* __restrict__ nbrk_p,
       long long* __restrict__ stop, long long* __restrict__ first, const unsigned char* __restrict__ valid,
       long long* __restrict__ NP, unsigned long long* __restrict__ steps) {
    __shared__ WarpScratch ws[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nb = *nbrk_p;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long k = (long long)blockIdx.x * 8 + warp; k < nb; k += nwarps) {
        const long long b = breaks[k];
        if (WRITE) {
            if (!valid[k]) continue;
            if (stop[k] == b) {
                if (lane == 0) {
                    NP[2 * b] = first[2 * k];
                    NP[2 * b + 1] = first[2 * k + 1];
                }
                continue;
            }
        }
        long long pj, pi;
        if (b == 0) {
            pj = p.seed_j;
            pi = p.seed_i;
        } else {
            accept_edge(g, G[b - 1], dG[b - 1], p.thr2, pj, pi);
        }
        long long t = b;
        unsigned long long n = 0;
        for (;;) {
            long long ej, ei, rj, ri;
            accept_edge(g, G[t], dG[t], p.thr2, ej, ei);
            chain_step(g, ws[warp], lane, p, t, pj, pi, yt, xt, ej, ei, rj, ri);
            ++n;
            if (WRITE) {
                if (lane == 0) {
                    NP[2 * t] = rj;
                    NP[2 * t + 1] = ri;
                }
                if (t >= stop[k]) break;
            } else {
                if (t == b && lane == 0) {
                    first[2 * k] = rj;
                    first[2 * k + 1] = ri;
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
        const long long* __restrict__ nbrk_p, unsigned char* __restrict__ valid) {
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
    long long* scal = (long long*)gzb_arena_take(ctx, 64);
    double* corner = (double*)gzb_arena_take(ctx, 64);
    if (!corner) return gzb_fail(ctx, GZB_ERR_NOMEM, "gzb_nearest: workspace too small");
    cudaStream_t s = ctx->stream;

    NearParams p;
    volatile double r0 = rd_found_km;
    volatile double t1 = 1.2 * r0;
    volatile double t2 = 1.2 * t1;
    p.r0 = r0;
    p.thr2 = t2;
    const double half = r0 / (2.0 * 6360.0);
    const double sn = sin(half < 1.5 ? half : 1.5);
    p.prox_r0 = 4.0 * sn * sn * (1.0 + 1.0e-6) + 1.0e-24;
    p.chord_r0 = 2.0 * sn * (1.0 + 1.0e-6) + 1.0e-12;
    p.r = np_box_r;
    p.seed_j = seed_j;
    p.seed_i = seed_i;
    p.corner = corner;

    GZB_CUDA(ctx, cudaMemsetAsync(scal, 0, 64, s));
    k_corner<<<1, 256, 0, s>>>(ctx->g, np_box_r, corner);
    GZB_LAUNCH_CHECK(ctx);

    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    long long gblocks = (long long)((n + 31) / 32);
    const long long gmax = (long long)nsm * 8 * 4;
    if (gblocks > gmax) gblocks = gmax;
    k_global<<<(unsigned)gblocks, 256, 0, s>>>(ctx->g, nt, yt, xt, G, dG);
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
                                                   valid, (long long*)np_out, (unsigned long long*)(scal + 1));
    GZB_LAUNCH_CHECK(ctx);
    k_valid<<<1, 32, 0, s>>>(breaks, stop, scal, valid);
    GZB_LAUNCH_CHECK(ctx);
    k_walk<true><<<(unsigned)wblocks, 256, 0, s>>>(ctx->g, p, nt, yt, xt, G, dG, breaks, scal, stop, first,
                                                  valid, (long long*)np_out, (unsigned long long*)(scal + 1));
    GZB_LAUNCH_CHECK(ctx);
    // stats: (breaks, chain steps) -> pinned words, read by gzb_get_stats after a sync
    GZB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, scal, 16, cudaMemcpyDeviceToHost, s));
    return GZB_OK;
}
