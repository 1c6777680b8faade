// tma_gather4.cu - micro-benchmark: random 128-byte row gathers issued as Blackwell TMA gather4 requests
// (cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4: ONE request fetches FOUR indexed rows of a 2-D
// tensor map = 512 B) with mbarrier completion.  Follow-up to tma_gather.cu (per-thread 128 B bulk copies were
// request-rate bound at 4.4 TB/s): does dividing the request count by four lift the TMA path to the rate of the
// register gathers (14.8 - 18.7 TB/s, l2_gather.cu)?  DESIGN.md section 5.1.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4 tma_gather4.cu   (driver API resolved at run time)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
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
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3, uint64_t* b) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(b)) : "memory");
}

constexpr int THREADS = 256;
constexpr int ROW_BYTES = 128;
constexpr int REQ_BYTES = 4 * ROW_BYTES;

// ISSUERS threads issue one gather4 each per stage (ISSUERS * 512 B per stage); all 256 threads consume.
template <int STAGES, int ISSUERS>
__global__ void __launch_bounds__(THREADS, 1) gather4_kernel(const __grid_constant__ CUtensorMap map, uint32_t row_mask, int iters, double2* out) {
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
        if (tid == 0) mbar_expect_tx(&full[s], ISSUERS * REQ_BYTES);
        if (tid < ISSUERS) {
            int r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { seed = mix(seed + 0x9e3779b9u); r[k] = (int)(seed & row_mask); }
            tma_gather4(ring + ((size_t)s * ISSUERS + tid) * REQ_BYTES, &map, 0, r[0], r[1], r[2], r[3], &full[s]);
        }
    };
    for (int s = 0; s < STAGES; ++s) issue(s);
    double2 acc = make_double2(0, 0);
    const int sub = tid & 7, grp = tid >> 3;               // 32 groups of 8 lanes: a group reads whole 128 B rows
    constexpr int ROWS_PER_STAGE = ISSUERS * 4;
    for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        for (int r = grp; r < ROWS_PER_STAGE; r += 32) {
            const double2 v = *reinterpret_cast<const double2*>(ring + ((size_t)s * ROWS_PER_STAGE + r) * ROW_BYTES + sub * 16);
            acc.x += v.x; acc.y += v.y;
        }
        __syncthreads();
        if (it + STAGES < iters) issue(s);
    }
    if (acc.x == 123.456) out[0] = acc;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int STAGES, int ISSUERS>
void run(const CUtensorMap& map, double2* out, int sms, int box_rows) {
    const int iters = 4096;
    const size_t smem = (size_t)STAGES * ISSUERS * REQ_BYTES;
    cudaFuncSetAttribute(gather4_kernel<STAGES, ISSUERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const double bytes = double(sms) * ISSUERS * REQ_BYTES * iters;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int lg = 19; lg <= 21; ++lg) {
        const uint32_t mask = (uint32_t)((size_t(1) << lg) - 1);
        gather4_kernel<STAGES, ISSUERS><<<sms, THREADS, smem>>>(map, mask, 64, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"box_rows\": %d, \"error\": \"%s\"}\n", box_rows, cudaGetErrorString(e)); exit(2); }
        cudaEventRecord(e0);
        for (int r = 0; r < 3; ++r) gather4_kernel<STAGES, ISSUERS><<<sms, THREADS, smem>>>(map, mask, iters, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
        printf("{\"mode\": \"tma_gather4\", \"box_rows\": %d, \"stages\": %d, \"requests_per_stage\": %d, \"in_flight_kb_per_sm\": %zu, "
               "\"window_mb\": %.0f, \"ms\": %.3f, \"tb_per_s\": %.2f, \"requests_per_us_per_sm\": %.1f}\n",
               box_rows, STAGES, ISSUERS, smem / 1024, (double)(size_t(1) << lg) * 128 / 1048576.0, ms, bytes / ms / 1e9,
               double(ISSUERS) * iters / (ms * 1e3));
        fflush(stdout);
    }
}

int main(int argc, char** argv) {
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1;     // tensor-map box height to try: 1 (expected for gather4) or 4
    const size_t rows = size_t(1) << 21;
    char* buf; double2* out;
    cudaMalloc(&buf, rows * 128); cudaMalloc(&out, 64);
    cudaMemset(buf, 0, rows * 128);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("{\"error\": \"cuTensorMapEncodeTiled not found\"}\n"); return 1;
    }
    CUtensorMap map;
    const cuuint64_t gdim[2] = {16, rows};                 // 16 fp64 columns (one 128 B row-tile) x rows
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, buf, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"box_rows\": %d, \"error\": \"cuTensorMapEncodeTiled -> %d\"}\n", box_rows, (int)r); return 1; }
    const int sms = prop.multiProcessorCount;
    run<2, 64>(map, out, sms, box_rows);
    run<4, 64>(map, out, sms, box_rows);
    run<6, 64>(map, out, sms, box_rows);
    run<3, 128>(map, out, sms, box_rows);
    run<12, 32>(map, out, sms, box_rows);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
