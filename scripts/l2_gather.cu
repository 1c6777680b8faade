// l2_gather.cu - micro-benchmark behind DESIGN.md section 5: how fast can 148 SMs gather random 128-byte
// row-tiles (8 lanes x 16 B, the SpMM access pattern) out of a buffer of W bytes?  W <= ~60 MB is
// L2-resident, W >> 126 MB is HBM-bound.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather l2_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int U>
__global__ void __launch_bounds__(256, 4) gather_kernel(const double2* __restrict__ buf, uint32_t row_mask, int iters, double2* out) {
    const int sub = threadIdx.x & 7;
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    double2 a0 = make_double2(0, 0), a1 = a0;
    uint32_t s = group * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        double2 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = mix(s + 0x9e3779b9u);
            x[u] = __ldg(buf + (size_t)(s & row_mask) * 8 + sub);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u & 1) { a1.x += x[u].x; a1.y += x[u].y; } else { a0.x += x[u].x; a0.y += x[u].y; }
        }
    }
    a0.x += a1.x; a0.y += a1.y;
    if (a0.x == 123.456) out[0] = a0;   // never true: keeps the loads alive
}

int main() {
    const size_t max_rows = size_t(1) << 25;            // 4 GB
    double2* buf; double2* out;
    cudaMalloc(&buf, max_rows * 128); cudaMalloc(&out, 64);
    cudaMemset(buf, 0, max_rows * 128);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int ctas = prop.multiProcessorCount * 4 * 4, threads = 256, iters = 256;
    const double bytes = double(ctas) * threads / 8 * iters * 8 * 128.0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("{\"sms\": %d, \"gather_bytes_per_launch\": %.0f, \"points\": [", prop.multiProcessorCount, bytes);
    bool first = true;
    for (int lg = 16; lg <= 25; ++lg) {
        const uint32_t mask = (uint32_t)((size_t(1) << lg) - 1);
        for (int w = 0; w < 3; ++w) gather_kernel<8><<<ctas, threads>>>(buf, mask, iters, out);
        cudaEventRecord(e0);
        const int reps = 5;
        for (int r = 0; r < reps; ++r) gather_kernel<8><<<ctas, threads>>>(buf, mask, iters, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        printf("%s{\"window_mb\": %.1f, \"ms\": %.3f, \"tb_per_s\": %.2f}", first ? "" : ", ", (double)(size_t(1) << lg) * 128 / 1048576.0, ms, bytes / ms / 1e9);
        first = false;
    }
    printf("]}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
