// Per-point device code of K4 shared by the gather kernel (gzb_interp.cu) and by the repair kernel
// of the fused path (gzb_mesh.cu): index wrapping, the tiled bit mask, and the fused time-linear +
// space-bilinear interpolation of one track point (gonzag/mod2sat.py:94-135).
#pragma once
#include "gzb_common.cuh"

__device__ __forceinline__ void gzb_l2_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Sparse loads with a 64-byte L2 fetch: a plain ld.global miss fills a whole 128-byte line from DRAM on B200
// (measured, scripts/fetch_gran.cu: 125.6 B of DRAM traffic per isolated 4-byte load; 62.8 B with .L2::64B;
// cudaLimitMaxL2FetchGranularity changes nothing).
template <bool L64> __device__ __forceinline__ float gzb_ldg(const float* p) {
    if (!L64) return *p;
    float v;
    asm volatile("ld.global.nc.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
template <bool L64> __device__ __forceinline__ double gzb_ldg(const double* p) {
    if (!L64) return *p;
    double v;
    asm volatile("ld.global.nc.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// Haversine_sclr / RadiusEarth (gonzag/utils.py:267-293): the along-track distance kernels (gzb_interp.cu)
// and the batched scalar helpers (gzb_mesh.cu) share them.
#define GZB_REQ2 40680631.59076899   /* 6378.137**2 */
#define GZB_RPL2 40408295.989504     /* 6356.752**2 */

__device__ __forceinline__ double radius_earth_dev(double lat) {
    const double p = lat * GZB_D2R;
    const double cp = cos(p), sp = sin(p);
    const double c0 = GZB_REQ2 * cp, d0 = GZB_RPL2 * sp, e0 = 6378.137 * cp, f0 = 6356.752 * sp;
    return sqrt((c0 * c0 + d0 * d0) / (e0 * e0 + f0 * f0));
}

__device__ __forceinline__ double haversine_sclr_dev(double lat1, double lon1, double lat2, double lon2) {
    const double R = radius_earth_dev(0.5 * (lat1 + lat2));
    const double dlat = (lat2 - lat1) * GZB_D2R;
    const double dlon = (lon2 - lon1) * GZB_D2R;
    const double p1 = lat1 * GZB_D2R, p2 = lat2 * GZB_D2R;
    const double s1 = sin(dlat / 2.0), s2 = sin(dlon / 2.0);
    const double a = s1 * s1 + cos(p1) * cos(p2) * (s2 * s2);
    return R * (2.0 * asin(sqrt(a)));
}

__device__ __forceinline__ long long pywrap_idx(long long k, long long n) {
    k = k < 0 ? k + n : k;                      // Python negative indexing (SM == -1 reads the last cell)
    return k < 0 ? 0 : (k >= n ? n - 1 : k);    // anything else out of range is clamped, never faults
}

// Index type of the per-point kernels.  Every grid of fewer than 2^31 cells (ORCA12: 13 M, an ORCA36-class grid: 117 M)
// takes the `int` variants: row / column numbers, flat cell indices and mask-tile numbers are then 32-bit values, which
// halves the integer work of K2-K4 (the reference's int64 arrays are only widened at the stores and narrowed at the
// loads).  `long long` is the general path.  Both give the same indices wherever the narrow one is used.
template <typename IX> __device__ __forceinline__ IX pywrap_ix(IX k, IX n) {
    k = k < 0 ? k + n : k;
    return k < 0 ? 0 : (k >= n ? n - 1 : k);
}
// pywrap_idx() of a 64-bit value read from memory (a mesh corner of a caller's SM array), as an IX: a value that does
// not fit 32 bits lies far outside the grid and clamps to the first / last index exactly as the 64-bit rule does
template <typename IX> __device__ __forceinline__ IX pywrap_from64(long long k, IX n);
template <> __device__ __forceinline__ long long pywrap_from64<long long>(long long k, long long n) { return pywrap_idx(k, n); }
template <> __device__ __forceinline__ int pywrap_from64<int>(long long k, int n) {
    const int lo = (int)k;
    if ((long long)lo == k) return pywrap_ix<int>(lo, n);
    return k < 0 ? 0 : n - 1;
}

template <typename IX>
__device__ __forceinline__ int mask_bit(const unsigned* __restrict__ bits, const IX ntI, const IX j, const IX i) {
    const IX tile = (j >> 4) * ntI + (i >> 4);
    const int b = (int)(((j & 15) << 4) | (i & 15));
    return (int)((GZB_LDS_LDG(bits + ((long long)tile * 8 + (b >> 5))) >> (b & 31)) & 1u);
}

// One thread per track point.  T is the dtype of the model field: the time interpolation
// x1 + ((x2 - x1) / dt) * (t - t1) is carried out in T (NumPy keeps float32 arrays in float32
// when combined with Python floats), the weighted sum in float64, left to right, no FMA.
// Loads are issued as early as their addresses are known (mesh and weights before the bracket
// search, the eight field values together with the mask) so that a point costs two dependent
// DRAM round trips, not four.
// Remote copies of the two output series: the other GPUs' product buffers, mapped through CUDA
// IPC, receive the values straight from the kernel (stores over NVLink, one per point and peer)
// instead of through a collective afterwards.  n == 0: local outputs only.
// (struct GzbPeerOut: gzb_common.cuh)

// REG: the mesh and weights come in registers (cj[4], ci[4], rw[2]) instead of SM / WB.
template <typename T, bool EARLY, bool REG = false, int PF = 0, typename IX = long long>
__device__ __forceinline__ void interp_point(const GzbGrid& g, const long long t, const double* __restrict__ t_sat,
         const long long* __restrict__ SM, const double* __restrict__ WB, const long long nrec,
         const double* __restrict__ tmod, const T* __restrict__ slab, const long long k0, const long long nk,
         const double rmiss, const int init, double* __restrict__ out_np, double* __restrict__ out_bl,
         int* __restrict__ err, const unsigned* __restrict__ mbits, const int* __restrict__ mflag,
         const GzbPeerOut& peers, const IX* rcj = nullptr, const IX* rci = nullptr, const double2* rw = nullptr) {
    const double now = __ldg(t_sat + t);
    const longlong2* sm = reinterpret_cast<const longlong2*>(SM + 8 * t);
    const double2* wb = reinterpret_cast<const double2*>(WB + 4 * t);
    if ((PF == 1 || PF == 2) && !REG && !EARLY) gzb_l2_prefetch(wb);      // option "prefetch": the weights travel with the mesh
    longlong2 p1, p2, p3, p4;
    double2 wa, wc;
    if (REG) {
        wa = rw[0]; wc = rw[1];
    } else if (EARLY) {
        p1 = sm[0]; p2 = sm[1]; p3 = sm[2]; p4 = sm[3];
        wa = wb[0]; wc = wb[1];
    }
    // the two results and whether this call produced them (init: they start as the missing value)
    double r_np = rmiss, r_bl = rmiss;
    bool w_np = init != 0, w_bl = init != 0;
    do {
        if (!(now >= tmod[0]) || !(now < tmod[nrec - 1])) {
            atomicOr(err, 1);
            break;
        }
        // bracket tmod[lo] <= now < tmod[lo+1] (gonzag/mod2sat.py:94-96).  Model records are almost always
        // evenly spaced: guess the index from the mean spacing, verify it against the actual times, and
        // only search (binary, eight dependent loads for 242 records) when the guess is off
        // (the guess is only a guess: single precision and an approximate division are plenty, whatever it says is
        // checked against the record times themselves)
        long long lo = (long long)(__fdividef((float)(now - tmod[0]), (float)(tmod[nrec - 1] - tmod[0])) * (float)(nrec - 1));
        lo = lo < 0 ? 0 : (lo > nrec - 2 ? nrec - 2 : lo);
        if (!(tmod[lo] <= now && now < tmod[lo + 1])) {
            lo = 0;
            long long hi = nrec - 1;                // tmod[lo] <= now < tmod[hi]
            while (hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if (tmod[mid] <= now) lo = mid; else hi = mid;
            }
        }
        if (lo < k0 || lo + 1 >= k0 + nk) break;    // bracket not inside this slab
        if (!EARLY && !REG) {
            p1 = sm[0]; p2 = sm[1]; p3 = sm[2]; p4 = sm[3];
        }
        const IX gNj = (IX)g.Nj, gNi = (IX)g.Ni;
        IX j1, i1, j2, i2, j3, i3, j4, i4;
        if (REG) {
            j1 = pywrap_ix<IX>(rcj[0], gNj); i1 = pywrap_ix<IX>(rci[0], gNi);
            j2 = pywrap_ix<IX>(rcj[1], gNj); i2 = pywrap_ix<IX>(rci[1], gNi);
            j3 = pywrap_ix<IX>(rcj[2], gNj); i3 = pywrap_ix<IX>(rci[2], gNi);
            j4 = pywrap_ix<IX>(rcj[3], gNj); i4 = pywrap_ix<IX>(rci[3], gNi);
        } else {
            j1 = pywrap_from64<IX>(p1.x, gNj); i1 = pywrap_from64<IX>(p1.y, gNi);
            j2 = pywrap_from64<IX>(p2.x, gNj); i2 = pywrap_from64<IX>(p2.y, gNi);
            j3 = pywrap_from64<IX>(p3.x, gNj); i3 = pywrap_from64<IX>(p3.y, gNi);
            j4 = pywrap_from64<IX>(p4.x, gNj); i4 = pywrap_from64<IX>(p4.y, gNi);
        }
        const IX c1 = j1 * gNi + i1, c2 = j2 * gNi + i2, c3 = j3 * gNi + i3, c4 = j4 * gNi + i4;
        const long long cells = g.Nj * g.Ni;
        const T* __restrict__ X1 = slab + (lo - k0) * cells;
        const T* __restrict__ X2 = X1 + cells;
        T a1, b1, a2, b2, a3, b3, a4, b4;
        if (EARLY) {
            a1 = X1[c1]; b1 = X2[c1]; a2 = X1[c2]; b2 = X2[c2];
            a3 = X1[c3]; b3 = X2[c3]; a4 = X1[c4]; b4 = X2[c4];
        }
        if (PF == 1 || PF == 2) {      // option "prefetch": the eight field values travel while the mask word is fetched and tested
            gzb_l2_prefetch(X1 + c1); gzb_l2_prefetch(X2 + c1); gzb_l2_prefetch(X1 + c2); gzb_l2_prefetch(X2 + c2);
            gzb_l2_prefetch(X1 + c3); gzb_l2_prefetch(X2 + c3); gzb_l2_prefetch(X1 + c4); gzb_l2_prefetch(X2 + c4);
        }
        int ocean;
        if (mbits && *mflag == 0) {
            const IX ntI = (gNi + 15) >> 4;
            ocean = mask_bit<IX>(mbits, ntI, j1, i1) + mask_bit<IX>(mbits, ntI, j2, i2) + mask_bit<IX>(mbits, ntI, j3, i3) +
                    mask_bit<IX>(mbits, ntI, j4, i4);
        } else {
            ocean = (int)g.mask[c1] + (int)g.mask[c2] + (int)g.mask[c3] + (int)g.mask[c4];
        }
        if (ocean != 4) break;
        if (!EARLY && !REG) {
            wa = wb[0]; wc = wb[1];
        }
        if (!EARLY) {      // PF == 3: 64-byte L2 fetches for the gather (option "prefetch" = 3)
            a1 = gzb_ldg<PF == 3>(X1 + c1); b1 = gzb_ldg<PF == 3>(X2 + c1); a2 = gzb_ldg<PF == 3>(X1 + c2); b2 = gzb_ldg<PF == 3>(X2 + c2);
            a3 = gzb_ldg<PF == 3>(X1 + c3); b3 = gzb_ldg<PF == 3>(X2 + c3); a4 = gzb_ldg<PF == 3>(X1 + c4); b4 = gzb_ldg<PF == 3>(X2 + c4);
        }
        const T dt = (T)(tmod[lo + 1] - tmod[lo]);
        const T de = (T)(now - tmod[lo]);
        const T v1 = a1 + ((b1 - a1) / dt) * de;
        const T v2 = a2 + ((b2 - a2) / dt) * de;
        const T v3 = a3 + ((b3 - a3) / dt) * de;
        const T v4 = a4 + ((b4 - a4) / dt) * de;
        r_np = (double)v1;
        w_np = true;
        const double sw = ((wa.x + wa.y) + wc.x) + wc.y;
        if (!(fabs(sw - 1.0) > 0.001)) {
            r_bl = ((wa.x * (double)v1 + wa.y * (double)v2) + wc.x * (double)v3) + wc.y * (double)v4;
            w_bl = true;
        }
    } while (0);
    if (w_np) {
        out_np[t] = r_np;
        for (int k = 0; k < peers.n; ++k) peers.np[k][t] = r_np;
    }
    if (w_bl) {
        out_bl[t] = r_bl;
        for (int k = 0; k < peers.n; ++k) peers.bl[k][t] = r_bl;
    }
}
