// K4: fused time-linear + space-bilinear gather (gonzag/mod2sat.py:74-139), plus the cheap
// a14 outputs: cumulative along-track distance (gonzag/mod2sat.py:146-148) and the
// nearest-point map XNPtrack (gonzag/mod2sat.py:177-184).
#include "gzb_common.cuh"
#include "gzb_point.cuh"

// Land-sea mask as bits, tiled 16 x 16 cells (one 32-byte sector per tile): the four corners of a
// source mesh almost always fall in one sector, and consecutive track points reuse it, where the
// row-major byte mask costs two sectors per point.  flag[0] != 0: the mask holds values other
// than {0,1}, the ocean test `sum == 4` must then read the bytes.
__global__ void __launch_bounds__(256)
k_mask_bits(const long long Nj, const long long Ni, const int8_t* __restrict__ mask, unsigned* __restrict__ bits,
            int* __restrict__ flag) {
    // one warp per tile row of 16 cells x 2 rows = one 32-bit word
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // word index
    const long long ntI = (Ni + 15) >> 4, ntJ = (Nj + 15) >> 4;
    if (w >= ntJ * ntI * 8) return;
    const long long tile = w >> 3;
    const int wr = (int)(w & 7);
    const long long tj = tile / ntI, ti = tile - tj * ntI;
    unsigned v = 0;
    bool odd = false;
#pragma unroll 4
    for (int b = 0; b < 32; ++b) {
        const long long j = tj * 16 + wr * 2 + (b >> 4), i = ti * 16 + (b & 15);
        if (j < Nj && i < Ni) {
            const int m = mask[j * Ni + i];
            if (m == 1) v |= 1u << b;
            else if (m != 0) odd = true;
        }
    }
    bits[w] = v;
    if (odd) atomicOr(flag, 1);
}

// resident blocks of 256 threads per SM for the 32-bit index variant of K4 (6 = a cap of 40 registers, no spills; measured
// 34.9 us at 5 and at 6 blocks, 41 us at 7 and 8 where the kernel spills -- gpurun_out/ab_exp*_r2n.log)
#ifndef GZB_K4_BLOCKS
#define GZB_K4_BLOCKS 6
#endif

// `list` == nullptr: thread per point t < nt.  Otherwise only the *count points listed (the ones
// whose mesh changed after a speculative pass, gzb_map_interp), grid-stride.
template <typename T, bool EARLY, int PF, typename IX>      // PF: 0 no prefetch, 1 prefetch, 2 prefetch with 64 registers (4 blocks per SM); IX: gzb_point.cuh
__global__ void __launch_bounds__(256, (EARLY || PF == 2) ? 4 : (sizeof(IX) == 4 ? GZB_K4_BLOCKS : 5))
k_interp(const GzbGrid g, const long long nt, const double* __restrict__ t_sat,
         const long long* __restrict__ SM, const double* __restrict__ WB, const long long nrec,
         const double* __restrict__ tmod, const T* __restrict__ slab, const long long k0, const long long nk,
         const double rmiss, const int init, double* __restrict__ out_np, double* __restrict__ out_bl,
         int* __restrict__ err, const unsigned* __restrict__ mbits, const int* __restrict__ mflag,
         const long long* __restrict__ list, const unsigned long long* __restrict__ count, const GzbPeerOut peers) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (list) {
        const long long n = (long long)*count;
        for (long long e = gt; e < n; e += (long long)gridDim.x * blockDim.x)
            interp_point<T, EARLY, false, PF, IX>(g, list[e], t_sat, SM, WB, nrec, tmod, slab, k0, nk, rmiss, init, out_np, out_bl,
                                                  err, mbits, mflag, peers);
    } else if (gt < nt) {
        interp_point<T, EARLY, false, PF, IX>(g, gt, t_sat, SM, WB, nrec, tmod, slab, k0, nk, rmiss, init, out_np, out_bl, err,
                                              mbits, mflag, peers);
    }
}

int gzb_build_mask_bits(gzb_ctx* ctx) {
    const long long Nj = ctx->g.Nj, Ni = ctx->g.Ni;
    const long long nw = ((Nj + 15) >> 4) * ((Ni + 15) >> 4) * 8;
    ctx->mask_bits_ok = 0;
    if (!ctx->d_mask) return GZB_OK;
    if (!ctx->d_mbits) {
        cudaError_t e = gzb_pool_malloc((void**)&ctx->d_mbits, (size_t)nw * sizeof(unsigned));
        if (e != cudaSuccess) { cudaGetLastError(); ctx->d_mbits = nullptr; return GZB_OK; }   // K4 reads the bytes
    }
    GZB_CUDA(ctx, cudaMemsetAsync(ctx->d_flags + 2, 0, 4, ctx->stream));
    k_mask_bits<<<(unsigned)((nw + 255) / 256), 256, 0, ctx->stream>>>(Nj, Ni, ctx->d_mask, ctx->d_mbits, ctx->d_flags + 2);
    GZB_LAUNCH_CHECK(ctx);
    ctx->mask_bits_ok = 1;
    return GZB_OK;
}

template <typename T>
static int launch_interp(gzb_ctx* ctx, int64_t nt, const double* t, const int64_t* sm, const double* wb,
                         int64_t nrec, const double* tmod, const void* slab, int64_t k0, int64_t nk,
                         double rmiss, int init, double* np_out, double* bl_out, int* err,
                         const long long* list = nullptr, const unsigned long long* count = nullptr) {
    const unsigned nblk = list ? 32u : (unsigned)((nt + 255) / 256);
    const unsigned* mb = ctx->mask_bits_ok ? ctx->d_mbits : nullptr;
    // remote copies only for the buffers they were registered for (gzb_set_peer_outputs)
    GzbPeerOut po;
    po.n = 0;
    if (ctx->peer_out.n > 0 && np_out == ctx->peer_local_np && bl_out == ctx->peer_local_bl && init) po = ctx->peer_out;
    int pf = 0;                 // option "prefetch": only for slabs in device memory
    if (ctx->prefetch) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, slab) == cudaSuccess) pf = at.type == cudaMemoryTypeDevice;
        else cudaGetLastError();
    }
    const bool ix32 = ctx->g.Nj * ctx->g.Ni < 0x7fffffffLL && !ctx->force_ix64;
#define GZB_K4I(EARLY, PF, IX) k_interp<T, EARLY, PF, IX><<<nblk, 256, 0, ctx->stream>>>(ctx->g, nt, t, (const long long*)sm, wb, nrec, tmod, \
        (const T*)slab, k0, nk, rmiss, init, np_out, bl_out, err, mb, ctx->d_flags + 2, list, count, po)
#define GZB_K4(EARLY, PF) { if (ix32) GZB_K4I(EARLY, PF, int); else GZB_K4I(EARLY, PF, long long); }
    if (ctx->k4_early) GZB_K4(true, 0)
    else if (pf && ctx->prefetch >= 3) GZB_K4(false, 3)
    else if (pf && ctx->prefetch == 2) GZB_K4(false, 2)
    else if (pf) GZB_K4(false, 1)
    else GZB_K4(false, 0)
#undef GZB_K4
#undef GZB_K4I
    GZB_LAUNCH_CHECK(ctx);
    return GZB_OK;
}

static int interp_chunk(gzb_ctx* ctx, int dtype, int64_t nt, const double* t, const int64_t* sm, const double* wb,
                        int64_t nrec, const double* tmod, const void* slab, int64_t k0, int64_t nk, double rmiss,
                        int init, double* np_out, double* bl_out, int* err, const long long* list = nullptr,
                        const unsigned long long* count = nullptr) {
    if (dtype == GZB_F32)
        return launch_interp<float>(ctx, nt, t, sm, wb, nrec, tmod, slab, k0, nk, rmiss, init, np_out, bl_out, err, list, count);
    return launch_interp<double>(ctx, nt, t, sm, wb, nrec, tmod, slab, k0, nk, rmiss, init, np_out, bl_out, err, list, count);
}

// device-level K4 on a slab the GPU can read (HBM, or pinned + mapped host memory through
// `slab_dev`); all other pointers are device pointers.  Used by gzb_map_interp.
int gzb_interp_dev(gzb_ctx* ctx, int64_t nt, const double* t, const int64_t* sm, const double* wb, int64_t nrec,
                   const double* tmod, const void* slab_dev, int dtype, int64_t k0, int64_t nk, double rmiss,
                   double* np_out, double* bl_out, const long long* list, const unsigned long long* count) {
    ctx->time_err_pending = 1;          // checked by gzb_sync()
    return interp_chunk(ctx, dtype, nt, t, sm, wb, nrec, tmod, slab_dev, k0, nk, rmiss, 1, np_out, bl_out, ctx->d_flags,
                        list, count);
}

// can the GPU read `slab` directly?  (device / managed memory, or pinned + mapped host memory)
int gzb_slab_device_pointer(const void* slab, const void** dev_out, int* is_host) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, slab);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) { *dev_out = slab; *is_host = 0; return 1; }
    if (at.type == cudaMemoryTypeHost && at.devicePointer) { *dev_out = at.devicePointer; *is_host = 1; return 1; }
    return 0;
}

extern "C" int gzb_interp(gzb_ctx* ctx, int64_t nt, const double* t, const int64_t* sm, const double* wb,
                          int64_t nrec, const double* tmod, const void* slab, int dtype, int64_t k0, int64_t nk,
                          double rmissval, int init_outputs, double* np_out, double* bl_out, int where) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_interp: null ctx");
    if (nt < 0 || nrec < 2 || nk < 0 || k0 < 0 || k0 + nk > nrec || !tmod || !slab || (nt > 0 && (!t || !sm || !wb || !np_out || !bl_out)))
        return gzb_fail(ctx, GZB_ERR_ARG, "gzb_interp: bad arguments (nt=%lld nrec=%lld k0=%lld nk=%lld)",
                        (long long)nt, (long long)nrec, (long long)k0, (long long)nk);
    if (dtype != GZB_F32 && dtype != GZB_F64) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_interp: dtype must be GZB_F32 or GZB_F64");
    if (!ctx->g.mask) return gzb_fail(ctx, GZB_ERR_STATE, "gzb_interp: the context has no land-sea mask (gzb_set_mask)");
    if (nt == 0) return GZB_OK;
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    const size_t elsz = dtype == GZB_F32 ? 4 : 8;
    const size_t rec_bytes = (size_t)ctx->g.Nj * (size_t)ctx->g.Ni * elsz;
    const size_t n = (size_t)nt;

    // where does the field slab live?  (independent of `where`)
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, slab);
    if (e != cudaSuccess) { cudaGetLastError(); at.type = cudaMemoryTypeUnregistered; at.devicePointer = nullptr; }
    bool slab_on_device = (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) || nk == 0;
    const bool slab_pinned = (at.type == cudaMemoryTypeHost);
    // Sparse gather straight from pinned + mapped host memory: a track touches 8 values per point
    // in records of Nj*Ni values, so when the sectors the kernel would read (<= 8 x 32 B per point)
    // are a small fraction of the slab, reading them over PCIe beats copying the slab to HBM.
    const void* slab_dev = slab;
    if (!slab_on_device && slab_pinned && at.devicePointer && ctx->zero_copy &&
        (ctx->zero_copy == 2 || (double)nt * 256.0 * 4.0 < (double)nk * (double)rec_bytes)) {
        slab_dev = at.devicePointer;
        slab_on_device = true;
        ctx->stats.h2d_bytes += (uint64_t)nt * 8 * 32;     // upper bound of what crosses PCIe: 8 sectors per point
    }

    // chunking of a host slab: consecutive chunks share one record
    int64_t C = nk;
    if (!slab_on_device) {
        const size_t budget = (size_t)256 << 20;
        C = (int64_t)(budget / rec_bytes);
        if (C < 2) C = 2;
        if (C > nk) C = nk;
    }
    size_t need = gzb_align((size_t)nrec * 8);
    if (where == GZB_HOST) need += gzb_align(n * 8) * 3 + gzb_align(n * 64) + gzb_align(n * 32);
    if (!slab_on_device) need += 2 * gzb_align((size_t)C * rec_bytes);
    if (gzb_arena_reserve(ctx, need) != GZB_OK) return GZB_ERR_NOMEM;
    cudaStream_t s = ctx->stream;
    int* d_err = ctx->d_flags;      // persistent, zero unless a previous call flagged an error
    const double* d_t = t;
    const long long* d_sm = (const long long*)sm;
    const double* d_wb = wb;
    const double* d_tmod = tmod;
    double* d_np = np_out;
    double* d_bl = bl_out;
    if (where == GZB_HOST) {
        double* a = (double*)gzb_arena_take(ctx, n * 8);
        long long* b = (long long*)gzb_arena_take(ctx, n * 64);
        double* c = (double*)gzb_arena_take(ctx, n * 32);
        double* tm = (double*)gzb_arena_take(ctx, (size_t)nrec * 8);
        d_np = (double*)gzb_arena_take(ctx, n * 8);
        d_bl = (double*)gzb_arena_take(ctx, n * 8);
        GZB_CUDA(ctx, cudaMemcpyAsync(a, t, n * 8, cudaMemcpyHostToDevice, s));
        GZB_CUDA(ctx, cudaMemcpyAsync(b, sm, n * 64, cudaMemcpyHostToDevice, s));
        GZB_CUDA(ctx, cudaMemcpyAsync(c, wb, n * 32, cudaMemcpyHostToDevice, s));
        GZB_CUDA(ctx, cudaMemcpyAsync(tm, tmod, (size_t)nrec * 8, cudaMemcpyHostToDevice, s));
        ctx->stats.h2d_bytes += n * 104 + (size_t)nrec * 8;
        if (!init_outputs) {
            GZB_CUDA(ctx, cudaMemcpyAsync(d_np, np_out, n * 8, cudaMemcpyHostToDevice, s));
            GZB_CUDA(ctx, cudaMemcpyAsync(d_bl, bl_out, n * 8, cudaMemcpyHostToDevice, s));
            ctx->stats.h2d_bytes += n * 16;
        }
        d_t = a; d_sm = b; d_wb = c; d_tmod = tm;
    }

    int rc = GZB_OK;
    if (slab_on_device) {
        rc = interp_chunk(ctx, dtype, nt, d_t, (const int64_t*)d_sm, d_wb, nrec, d_tmod, slab_dev, k0, nk, rmissval,
                          init_outputs, d_np, d_bl, d_err);
        if (rc != GZB_OK) return rc;
    } else {
        char* dbuf[2];
        dbuf[0] = (char*)gzb_arena_take(ctx, (size_t)C * rec_bytes);
        dbuf[1] = (char*)gzb_arena_take(ctx, (size_t)C * rec_bytes);
        if (!dbuf[1]) return gzb_fail(ctx, GZB_ERR_NOMEM, "gzb_interp: workspace too small");
        if (!slab_pinned && ctx->stage_cap < (size_t)C * rec_bytes) {
            for (int k = 0; k < 2; ++k) {
                if (ctx->stage[k]) cudaFreeHost(ctx->stage[k]);
                ctx->stage[k] = nullptr;
                GZB_CUDA(ctx, cudaHostAlloc(&ctx->stage[k], (size_t)C * rec_bytes, cudaHostAllocPortable));
            }
            ctx->stage_cap = (size_t)C * rec_bytes;
        }
        // the copy stream must not run ahead of work already queued on the main stream
        GZB_CUDA(ctx, cudaEventRecord(ctx->ev[0], s));
        GZB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[0], 0));
        bool used[2] = {false, false};
        int init = init_outputs;
        int64_t a = k0;
        int c = 0;
        const int64_t kend = k0 + nk;
        for (;;) {
            const int64_t len = (kend - a < C) ? (kend - a) : C;
            const int b = c & 1;
            const char* src = (const char*)slab + (size_t)(a - k0) * rec_bytes;
            const size_t bytes = (size_t)len * rec_bytes;
            if (used[b]) GZB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[2 + b], 0));
            if (slab_pinned) {
                GZB_CUDA(ctx, cudaMemcpyAsync(dbuf[b], src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            } else {
                if (used[b]) GZB_CUDA(ctx, cudaEventSynchronize(ctx->ev[b]));   // previous H2D out of stage[b] done
                memcpy(ctx->stage[b], src, bytes);
                GZB_CUDA(ctx, cudaMemcpyAsync(dbuf[b], ctx->stage[b], bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            }
            ctx->stats.h2d_bytes += bytes;
            GZB_CUDA(ctx, cudaEventRecord(ctx->ev[b], ctx->copy_stream));
            GZB_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev[b], 0));
            rc = interp_chunk(ctx, dtype, nt, d_t, (const int64_t*)d_sm, d_wb, nrec, d_tmod, dbuf[b], a, len, rmissval,
                              init, d_np, d_bl, d_err);
            if (rc != GZB_OK) return rc;
            GZB_CUDA(ctx, cudaEventRecord(ctx->ev[2 + b], s));
            used[b] = true;
            init = 0;
            ++c;
            if (a + len >= kend) break;
            a += len - 1;
        }
    }
    int* h_err = (int*)((char*)ctx->pinned + 64);
    const bool sync_now = (where == GZB_HOST || !slab_on_device || slab_dev != slab);
    if (sync_now) GZB_CUDA(ctx, cudaMemcpyAsync(h_err, d_err, 4, cudaMemcpyDeviceToHost, s));
    else ctx->time_err_pending = 1;          // checked by gzb_sync()
    if (where == GZB_HOST) {
        GZB_CUDA(ctx, cudaMemcpyAsync(np_out, d_np, n * 8, cudaMemcpyDeviceToHost, s));
        GZB_CUDA(ctx, cudaMemcpyAsync(bl_out, d_bl, n * 8, cudaMemcpyDeviceToHost, s));
        ctx->stats.d2h_bytes += n * 16;
    }
    if (sync_now) {
        GZB_CUDA(ctx, cudaStreamSynchronize(s));
        if (*h_err) {
            *h_err = 0;
            GZB_CUDA(ctx, cudaMemsetAsync(d_err, 0, 4, s));
            return gzb_fail(ctx, GZB_ERR_TIME, "gzb_interp: track time outside the model records [tmod[0], tmod[nrec-1])");
        }
    }
    return GZB_OK;
}

// ------------------------------------------------------------------------------------------
// along-track distance: per-step Haversine_sclr (gonzag/utils.py:267-293, device code in gzb_point.cuh) then a
// prefix sum.  block-local inclusive scan of the step lengths; block totals go to `tot`
__global__ void __launch_bounds__(1024)
k_dist_blocks(const long long nt, const double* __restrict__ yt, const double* __restrict__ xt,
              double* __restrict__ out, double* __restrict__ tot) {
    __shared__ double wsum[32];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v = 0.0;
    if (t < nt && t > 0) v = haversine_sclr_dev(yt[t], xt[t], yt[t - 1], xt[t - 1]);
    double s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += u;
    }
    if (lane == 31) wsum[warp] = s;
    __syncthreads();
    if (warp == 0) {
        double w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    const double incl = s + (warp ? wsum[warp - 1] : 0.0);
    if (t < nt) out[t] = incl;
    if (threadIdx.x == 1023) tot[blockIdx.x] = incl;
}

__global__ void __launch_bounds__(1024)
k_dist_totals(double* __restrict__ tot, const long long nblk) {
    __shared__ double wsum[32];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < nblk; base += 1024) {
        const long long k = base + threadIdx.x;
        const double v = k < nblk ? tot[k] : 0.0;
        double s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        if (lane == 31) wsum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            double w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const double excl = carry + (warp ? wsum[warp - 1] : 0.0) + s - v;
        if (k < nblk) tot[k] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024)
k_dist_add(const long long nt, double* __restrict__ out, const double* __restrict__ tot) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nt && blockIdx.x > 0) out[t] += tot[blockIdx.x];
}

extern "C" int gzb_track_distance(gzb_ctx* ctx, int64_t nt, const double* yt, const double* xt, double* dist_out,
                                  int where) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_track_distance: null ctx");
    if (nt < 0 || (nt > 0 && (!yt || !xt || !dist_out))) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_track_distance: bad arguments");
    if (nt == 0) return GZB_OK;
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    const size_t n = (size_t)nt;
    const long long nblk = (long long)((n + 1023) / 1024);
    size_t need = gzb_align((size_t)nblk * 8);
    if (where == GZB_HOST) need += 3 * gzb_align(n * 8);
    if (gzb_arena_reserve(ctx, need) != GZB_OK) return GZB_ERR_NOMEM;
    cudaStream_t s = ctx->stream;
    double* tot = (double*)gzb_arena_take(ctx, (size_t)nblk * 8);
    const double* dy = yt;
    const double* dx = xt;
    double* dout = dist_out;
    if (where == GZB_HOST) {
        double* a = (double*)gzb_arena_take(ctx, n * 8);
        double* b = (double*)gzb_arena_take(ctx, n * 8);
        dout = (double*)gzb_arena_take(ctx, n * 8);
        GZB_CUDA(ctx, cudaMemcpyAsync(a, yt, n * 8, cudaMemcpyHostToDevice, s));
        GZB_CUDA(ctx, cudaMemcpyAsync(b, xt, n * 8, cudaMemcpyHostToDevice, s));
        ctx->stats.h2d_bytes += n * 16;
        dy = a;
        dx = b;
    }
    k_dist_blocks<<<(unsigned)nblk, 1024, 0, s>>>(nt, dy, dx, dout, tot);
    GZB_LAUNCH_CHECK(ctx);
    if (nblk > 1) {
        k_dist_totals<<<1, 1024, 0, s>>>(tot, nblk);
        GZB_LAUNCH_CHECK(ctx);
        k_dist_add<<<(unsigned)nblk, 1024, 0, s>>>(nt, dout, tot);
        GZB_LAUNCH_CHECK(ctx);
    }
    if (where == GZB_HOST) {
        GZB_CUDA(ctx, cudaMemcpyAsync(dist_out, dout, n * 8, cudaMemcpyDeviceToHost, s));
        ctx->stats.d2h_bytes += n * 8;
        GZB_CUDA(ctx, cudaStreamSynchronize(s));
    }
    return GZB_OK;
}

// ------------------------------------------------------------------------------------------
// imask of Model2SatTrack (gonzag/mod2sat.py:157-158): 1 where the interpolated model value and the satellite value
// both lie in (-100, 100) -- NaN and inf fail, as in the reference's chained comparisons -- plus its complement, the
// mask of the three masked_where(imask == 0, .) products (:160-162)
__global__ void k_track_mask(const long long nt, const double* __restrict__ bl, const double* __restrict__ sat,
                             signed char* __restrict__ imask, unsigned char* __restrict__ bad) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const double a = bl[t], b = sat[t];
    const bool good = (a > -100.0) && (a < 100.0) && (b > -100.0) && (b < 100.0);
    imask[t] = good ? 1 : 0;
    if (bad) bad[t] = good ? 0 : 1;
}

extern "C" int gzb_track_mask(gzb_ctx* ctx, int64_t nt, const double* bl, const double* sat, int8_t* imask_out,
                              unsigned char* bad_out_or_null) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_track_mask: null ctx");
    if (nt < 0 || (nt > 0 && (!bl || !sat || !imask_out))) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_track_mask: bad arguments");
    if (nt == 0) return GZB_OK;
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    k_track_mask<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(nt, bl, sat, (signed char*)imask_out, bad_out_or_null);
    GZB_LAUNCH_CHECK(ctx);
    return GZB_OK;
}

// ------------------------------------------------------------------------------------------
// XNPtrack: last track index per cell (Python's last-writer-wins loop == max index)
__global__ void k_xnp_scatter(const GzbGrid g, const long long nt, const long long* __restrict__ NP,
                              long long* __restrict__ last) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const long long j = pywrap_idx(NP[2 * t], g.Nj), i = pywrap_idx(NP[2 * t + 1], g.Ni);
    atomicMax(last + j * g.Ni + i, t);
}

__global__ void k_xnp_final(const GzbGrid g, const long long* __restrict__ last, const double rmiss,
                            double* __restrict__ out, unsigned char* __restrict__ missing) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.Nj * g.Ni) return;
    const long long l = last[k];
    double v = l >= 0 ? (double)l : rmiss;
    if (g.mask[k] == 0) v = -100.0;
    out[k] = v;
    if (missing) missing[k] = (v == rmiss) ? 1 : 0;      // the mask of masked_where(XNPtrack == rmissval), mod2sat.py:186
}

static int xnp_track_impl(gzb_ctx* ctx, int64_t nt, const int64_t* np, double rmissval, double* xnp_out,
                          unsigned char* missing_out, int where);

extern "C" int gzb_xnp_track(gzb_ctx* ctx, int64_t nt, const int64_t* np, double rmissval, double* xnp_out,
                             int where) {
    return xnp_track_impl(ctx, nt, np, rmissval, xnp_out, nullptr, where);
}

extern "C" int gzb_xnp_track_masked(gzb_ctx* ctx, int64_t nt, const int64_t* np, double rmissval, double* xnp_out,
                                    unsigned char* missing_out, int where) {
    if (!missing_out) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_xnp_track_masked: null missing_out");
    return xnp_track_impl(ctx, nt, np, rmissval, xnp_out, missing_out, where);
}

static int xnp_track_impl(gzb_ctx* ctx, int64_t nt, const int64_t* np, double rmissval, double* xnp_out,
                          unsigned char* missing_out, int where) {
    if (!ctx) return gzb_fail(nullptr, GZB_ERR_ARG, "gzb_xnp_track: null ctx");
    if (nt < 0 || !xnp_out || (nt > 0 && !np)) return gzb_fail(ctx, GZB_ERR_ARG, "gzb_xnp_track: bad arguments");
    if (!ctx->g.mask) return gzb_fail(ctx, GZB_ERR_STATE, "gzb_xnp_track: the context has no land-sea mask");
    if (gzb_use_device(ctx) != GZB_OK) return GZB_ERR_CUDA;
    const size_t n = (size_t)nt, cells = (size_t)ctx->g.Nj * (size_t)ctx->g.Ni;
    size_t need = gzb_align(cells * 8);
    if (where == GZB_HOST) need += gzb_align(n * 16 + 16) + gzb_align(cells * 8) + gzb_align(cells);
    if (gzb_arena_reserve(ctx, need) != GZB_OK) return GZB_ERR_NOMEM;
    cudaStream_t s = ctx->stream;
    long long* last = (long long*)gzb_arena_take(ctx, cells * 8);
    const long long* dnp = (const long long*)np;
    double* dout = xnp_out;
    unsigned char* dmiss = missing_out;
    if (where == GZB_HOST) {
        long long* a = (long long*)gzb_arena_take(ctx, n * 16 + 16);
        dout = (double*)gzb_arena_take(ctx, cells * 8);
        if (missing_out) dmiss = (unsigned char*)gzb_arena_take(ctx, cells);
        if (nt > 0) GZB_CUDA(ctx, cudaMemcpyAsync(a, np, n * 16, cudaMemcpyHostToDevice, s));
        ctx->stats.h2d_bytes += n * 16;
        dnp = a;
    }
    GZB_CUDA(ctx, cudaMemsetAsync(last, 0xff, cells * 8, s));
    if (nt > 0) {
        k_xnp_scatter<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(ctx->g, nt, dnp, last);
        GZB_LAUNCH_CHECK(ctx);
    }
    k_xnp_final<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(ctx->g, last, rmissval, dout, dmiss);
    GZB_LAUNCH_CHECK(ctx);
    if (where == GZB_HOST) {
        GZB_CUDA(ctx, cudaMemcpyAsync(xnp_out, dout, cells * 8, cudaMemcpyDeviceToHost, s));
        ctx->stats.d2h_bytes += cells * 8;
        if (missing_out) {
            GZB_CUDA(ctx, cudaMemcpyAsync(missing_out, dmiss, cells, cudaMemcpyDeviceToHost, s));
            ctx->stats.d2h_bytes += cells;
        }
        GZB_CUDA(ctx, cudaStreamSynchronize(s));
    }
    return GZB_OK;
}
