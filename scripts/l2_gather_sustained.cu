// l2_gather_sustained.cu - does the WIDTH of the gather instruction matter once the GPU is power-capped?
// Same random 128-byte row gathers as l2_gather.cu on a 128 MB window (one C3 panel), issued either as
// 8 lanes x 128-bit loads (the SpMM's current shape) or as 4 lanes x 256-bit loads (LDG.E.256, half the load
// instructions and address arithmetic per byte), each run back to back for ~8 s so that the power cap binds.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather_sustained l2_gather_sustained.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int U>
__global__ void __launch_bounds__(256, 4) gather128(const double2* __restrict__ buf, uint32_t row_mask, int iters, double2* out) {
    const int sub = threadIdx.x & 7;
    uint32_t s = ((blockIdx.x * blockDim.x + threadIdx.x) >> 3) * 2654435761u + 12345u;
    double2 a0 = make_double2(0, 0), a1 = a0;
    for (int it = 0; it < iters; ++it) {
        double2 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { s = mix(s + 0x9e3779b9u); x[u] = __ldg(buf + (size_t)(s & row_mask) * 8 + sub); }
#pragma unroll
        for (int u = 0; u < U; ++u) { if (u & 1) { a1.x += x[u].x; a1.y += x[u].y; } else { a0.x += x[u].x; a0.y += x[u].y; } }
    }
    a0.x += a1.x; a0.y += a1.y;
    if (a0.x == 123.456) out[0] = a0;
}

// 4 lanes per row, 32 bytes per lane; U/2 loads per lane keep the same bytes in flight per lane as gather128<U>
template <int U>
__global__ void __launch_bounds__(256, 4) gather256(const double* __restrict__ buf, uint32_t row_mask, int iters, double2* out) {
    const int sub = threadIdx.x & 3;
    uint32_t s = ((blockIdx.x * blockDim.x + threadIdx.x) >> 2) * 2654435761u + 12345u;
    double a[4] = {0, 0, 0, 0};
    for (int it = 0; it < iters; ++it) {
        double x[U / 2][4];
#pragma unroll
        for (int u = 0; u < U / 2; ++u) {
            s = mix(s + 0x9e3779b9u);
            const double* p = buf + (size_t)(s & row_mask) * 16 + sub * 4;
            asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x[u][0]), "=d"(x[u][1]), "=d"(x[u][2]), "=d"(x[u][3]) : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < U / 2; ++u)
#pragma unroll
            for (int c = 0; c < 4; ++c) a[c] += x[u][c];
    }
    if (a[0] + a[1] + a[2] + a[3] == 123.456) out[0] = make_double2(a[0], a[1]);
}

template <class F>
static void sustained(const char* name, double bytes_per_launch, F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 20; ++w) launch();
    cudaDeviceSynchronize();
    // ~8 s: measure the first and the last second separately
    float first = 0, last = 0; int n_first = 0, n_last = 0;
    double t_total = 0;
    while (t_total < 8000.0) {
        cudaEventRecord(e0);
        for (int r = 0; r < 50; ++r) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (t_total < 1000.0) { first += ms; n_first += 50; }
        if (t_total >= 7000.0) { last += ms; n_last += 50; }
        t_total += ms;
    }
    printf("{\"variant\": \"%s\", \"tb_per_s_first_second\": %.2f, \"tb_per_s_after_7s\": %.2f}\n", name,
           bytes_per_launch * n_first / first / 1e9, bytes_per_launch * n_last / (last > 0 ? last : 1) / 1e9);
    fflush(stdout);
}

int main() {
    const size_t rows = size_t(1) << 20;          // 128 MB window
    double* buf; double2* out;
    cudaMalloc(&buf, rows * 128); cudaMalloc(&out, 64);
    cudaMemset(buf, 0, rows * 128);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const uint32_t mask = (uint32_t)(rows - 1);
    const int ctas = prop.multiProcessorCount * 16, iters = 64;
    const double b128 = double(ctas) * 256 / 8 * iters * 8 * 128.0;      // groups x iters x U rows x 128 B
    const double b256 = double(ctas) * 256 / 4 * iters * 4 * 128.0;
    sustained("8 lanes x 128-bit, U=8", b128, [&] { gather128<8><<<ctas, 256>>>((const double2*)buf, mask, iters, out); });
    sustained("4 lanes x 256-bit, 4 loads/lane", b256, [&] { gather256<8><<<ctas, 256>>>(buf, mask, iters, out); });
    sustained("8 lanes x 128-bit, U=8 (again)", b128, [&] { gather128<8><<<ctas, 256>>>((const double2*)buf, mask, iters, out); });
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
