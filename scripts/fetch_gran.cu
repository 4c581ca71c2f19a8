// Micro-benchmark (measurement helper, not product code): how many DRAM bytes does ONE sparse 4-byte load cost
// on this GPU, per load flavour, and does cudaLimitMaxL2FetchGranularity change it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/fetch_gran scripts/fetch_gran.cu
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --csv ./scripts/fetch_gran [limit_bytes]
// Every thread reads one float from its own pseudo-random 128-byte line of a 4 GiB buffer (no reuse, no
// coalescing): dram__bytes_read / loads = the fetch granularity.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ float load(const float* p) {
    float v;
    if (MODE == 0) asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.cv.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if (MODE == 4) asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if (MODE == 5) asm volatile("ld.global.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if (MODE == 6) asm volatile("ld.global.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else asm volatile("ld.global.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int MODE>
__global__ void gather(const float* __restrict__ buf, unsigned long long nlines, float* out, int per_warp_same) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long h = (per_warp_same ? (t >> 5) : t) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    const unsigned long long line = h % nlines;
    // per_warp_same: the 32 lanes read 32 consecutive floats of one line (a coalesced 128-byte request)
    const float v = load<MODE>(buf + line * 32 + (per_warp_same ? (threadIdx.x & 31) : (t & 7)));
    if (v == 123.456f) out[0] = v;
}

int main(int argc, char** argv) {
    size_t lim = 0;
    if (argc > 1) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
        printf("set limit %s: %s\n", argv[1], cudaGetErrorString(e));
    }
    cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity = %zu\n", lim);
    const size_t bytes = (size_t)4 << 30;
    float *buf, *out;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&out, 4);
    cudaMemset(buf, 0, bytes);
    const unsigned long long nlines = bytes / 128;
    const int n = 1 << 22;      // loads per launch
#define RUN(M) gather<M><<<n / 256, 256>>>(buf, nlines, out, 0); cudaDeviceSynchronize();
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
    gather<0><<<n / 256, 256>>>(buf, nlines, out, 1);      // coalesced 128-byte requests: the reference point
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
