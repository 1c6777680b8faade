// tma_gather.cu - micro-benchmark: random 128-byte row gathers issued as per-thread bulk async copies
// (cp.async.bulk global -> shared, mbarrier completion) instead of register loads.  Question answered:
// can the TMA path keep more bytes in flight per SM than the register file allows?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather tma_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

constexpr int THREADS = 256;
constexpr int ROW_BYTES = 128;

template <int STAGES>
__global__ void __launch_bounds__(THREADS, 1) tma_gather_kernel(const char* __restrict__ buf, uint32_t row_mask, int iters, double2* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[STAGES];
    char* ring = reinterpret_cast<char*>(smem);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t seed = (blockIdx.x * THREADS + tid) * 2654435761u + 12345u;
    auto issue = [&](int s) {
        if (tid == 0) mbar_expect_tx(&full[s], THREADS * ROW_BYTES);
        seed = mix(seed + 0x9e3779b9u);
        bulk_g2s(ring + ((size_t)s * THREADS + tid) * ROW_BYTES, buf + (size_t)(seed & row_mask) * ROW_BYTES, ROW_BYTES, &full[s]);
    };
    // NOTE: thread 0's expect_tx must be ordered before any complete_tx can flip the phase; the tx-count may go
    // transiently negative, which mbarrier semantics allow within a phase.
    for (int s = 0; s < STAGES; ++s) issue(s);
    double2 acc = make_double2(0, 0);
    const int sub = tid & 7, grp = tid >> 3;
    for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        // consume: every 8-lane group reads 8 of the stage's 256 rows (16 B per lane)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const double2 v = *reinterpret_cast<const double2*>(ring + ((size_t)s * THREADS + grp * 8 + r) * ROW_BYTES + sub * 16);
            acc.x += v.x; acc.y += v.y;
        }
        __syncthreads();
        if (it + STAGES < iters) issue(s);
    }
    if (acc.x == 123.456) out[0] = acc;
}

template <int STAGES>
void run(const char* buf, double2* out, int sms) {
    const int iters = 2048;
    const size_t smem = (size_t)STAGES * THREADS * ROW_BYTES;
    cudaFuncSetAttribute(tma_gather_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const double bytes = double(sms) * THREADS * ROW_BYTES * iters;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int lg = 19; lg <= 21; ++lg) {
        const uint32_t mask = (uint32_t)((size_t(1) << lg) - 1);
        tma_gather_kernel<STAGES><<<sms, THREADS, smem>>>(buf, mask, iters, out);
        cudaEventRecord(e0);
        for (int r = 0; r < 3; ++r) tma_gather_kernel<STAGES><<<sms, THREADS, smem>>>(buf, mask, iters, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
        printf("{\"stages\": %d, \"in_flight_kb_per_sm\": %zu, \"window_mb\": %.0f, \"ms\": %.3f, \"tb_per_s\": %.2f}\n", STAGES, smem / 1024,
               (double)(size_t(1) << lg) * 128 / 1048576.0, ms, bytes / ms / 1e9);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return; }
    }
}

int main() {
    const size_t rows = size_t(1) << 21;
    char* buf; double2* out;
    cudaMalloc(&buf, rows * 128); cudaMalloc(&out, 64);
    cudaMemset(buf, 0, rows * 128);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    run<2>(buf, out, prop.multiProcessorCount);
    run<4>(buf, out, prop.multiProcessorCount);
    run<6>(buf, out, prop.multiProcessorCount);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
