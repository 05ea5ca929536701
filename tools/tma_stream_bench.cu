// Dev micro-benchmark (round-1 open question, profiles/r01_summary_final.md): how fast can 148 persistent CTAs stream a
// [N, K] bf16 matrix from HBM into shared memory, as a function of the TMA request shape?
//   mode 0: 1-D bulk copies of 16 KB contiguous (a tile-major weight layout would allow this)
//   mode 1: 2-D tensor box {64 cols, 128 rows}  = 16 KB, 128 B contiguous per row      (the swap-AB decode GEMM today)
//   mode 2: 3-D tensor box {64 cols, 16 rows, 16 column blocks} = 32 KB, 2 KB per row  (the persistent decode kernel)
//   mode 3: 3-D tensor box {64 cols, 128 rows, 2 column blocks} = 32 KB, 256 B per row
// No compute: a consumer warp just releases the stage.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream_bench
//   tools/tma_stream_bench.cu -lcuda ; run: ./tma_stream_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
    } while (!done);
}

constexpr int kStagesMax = 12;

// every CTA streams a contiguous range of stages (column block fastest, like the k loop of a GEMM tile / a row block of the
// persistent kernel); the stage -> address mapping covers the matrix exactly once
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const uint16_t* W, int mode, int nstages, int stage_bytes,
                                                        long long total_stages, int K, int N, int passes = 1) {
    extern __shared__ __align__(1024) uint8_t ring[];
    __shared__ __align__(8) uint64_t full[kStagesMax], empty[kStagesMax];
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nk64 = K / 64;
    if (threadIdx.x == 0) {            // producer
        unsigned c = 0;
        const long long i0 = total_stages * blockIdx.x / gridDim.x, i1 = total_stages * (blockIdx.x + 1) / gridDim.x;
        for (int pass = 0; pass < passes; ++pass)
        for (long long i = i0; i < i1; ++i, ++c) {
            const int st = c % nstages;
            mbar_wait(&empty[st], ((c / nstages) & 1) ^ 1);
            mbar_expect(&full[st], stage_bytes);
            uint8_t* dst = ring + (size_t)st * stage_bytes;
            if (mode == 0) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
                             "l"((const uint8_t*)W + (size_t)i * stage_bytes), "r"(stage_bytes), "r"(s32(&full[st])) : "memory");
            } else if (mode == 1) {    // stage i = (row tile, column block): column block fastest, like a k loop
                const int rt = (int)(i / nk64), kb = (int)(i % nk64);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(dst)),
                             "l"(&tm), "r"(kb * 64), "r"(rt * 128), "r"(s32(&full[st])) : "memory");
            } else if (mode == 2) {    // 16 rows x 1024 columns
                const int ncc = K / 1024;
                const int blk = (int)(i / ncc), cc = (int)(i % ncc);
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(dst)),
                             "l"(&tm), "r"(0), "r"(blk * 16), "r"(cc * 16), "r"(s32(&full[st])) : "memory");
            } else {                   // 128 rows x 128 columns
                const int n2 = nk64 / 2;
                const int rt = (int)(i / n2), kb2 = (int)(i % n2);
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(dst)),
                             "l"(&tm), "r"(0), "r"(rt * 128), "r"(kb2 * 2), "r"(s32(&full[st])) : "memory");
            }
        }
    } else if (threadIdx.x == 32) {    // consumer: release the stage as soon as it has landed
        unsigned c = 0;
        const long long i0 = total_stages * blockIdx.x / gridDim.x, i1 = total_stages * (blockIdx.x + 1) / gridDim.x;
        for (int pass = 0; pass < passes; ++pass)
        for (long long i = i0; i < i1; ++i, ++c) {
            const int st = c % nstages;
            mbar_wait(&full[st], (c / nstages) & 1);
            mbar_arrive(&empty[st]);
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int N = 28672, K = 4096;            // Mistral-7B gate|up: 235 MB
    // four copies of the matrix, visited round-robin by back-to-back launches: every launch reads 235 MB that are not in L2,
    // and the GPU stays busy (clocks up) for the whole measurement, as in a decode step
    constexpr int kCopies = 4, kIters = 40;
    uint16_t* Wall;
    CK(cudaMalloc(&Wall, (size_t)kCopies * N * K * 2));
    CK(cudaMemset(Wall, 1, (size_t)kCopies * N * K * 2));
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char* names[] = {"1-D bulk, 16 KB contiguous", "2-D box 64 x 128 rows (16 KB, 128 B per row)", "3-D box 64 x 16 rows x 16 col blocks (32 KB, 2 KB per row)",
                           "3-D box 64 x 128 rows x 2 col blocks (32 KB, 256 B per row)"};
    for (int mode = 0; mode < 4; ++mode) {
        int stage_bytes = 16384;
        CUtensorMap tms[kCopies];
        for (int cp = 0; cp < kCopies; ++cp) {
        CUtensorMap& tm = tms[cp];
        uint16_t* W = Wall + (size_t)cp * N * K;
        const cuuint32_t es[3] = {1, 1, 1};
        if (mode == 1 || mode == 0) {
            const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
            const cuuint64_t str[1] = {(cuuint64_t)K * 2};
            const cuuint32_t box[2] = {64, 128};
            enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            const cuuint64_t dims[3] = {64, (cuuint64_t)N, (cuuint64_t)K / 64};
            const cuuint64_t str[2] = {(cuuint64_t)K * 2, 128};
            const cuuint32_t box2[3] = {64, 16, 16}, box3[3] = {64, 128, 2};
            enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, W, dims, str, mode == 2 ? box2 : box3, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            stage_bytes = 32768;
        }
        }
        for (int inflight_kb : {96, 128, 160, 192}) {
            const int nstages = inflight_kb * 1024 / stage_bytes;
            const size_t smem = (size_t)nstages * stage_bytes;
            CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const long long total = (long long)N * K * 2 / stage_bytes;
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                for (int it = 0; it < kIters; ++it)
                    stream_kernel<<<148, 64, smem>>>(tms[it % kCopies], Wall + (size_t)(it % kCopies) * N * K, mode, nstages, stage_bytes, total, K, N);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                ms /= kIters;
                if (rep > 0 && ms < best) best = ms;
            }
            printf("%-62s in flight %3d KB/SM (%2d stages): %7.1f us  %6.0f GB/s\n", names[mode], inflight_kb, nstages, best * 1e3, (double)N * K * 2 / (best * 1e-3) / 1e9);
        }
    }
    // L2-resident variant (round 2): the same 3-D 32 KB boxes over a 33.5 MB matrix (Mistral o_proj) re-read by every launch, so all
    // but the first launch hit the L2.  Answers: can the L2 feed the SMs faster than HBM does (would an L2 prefetch of the next
    // phase's weights during the persistent kernel's grid barriers buy anything)?
    {
        const int N2 = 4096, K2 = 4096;
        CUtensorMap tm;
        const cuuint32_t es[3] = {1, 1, 1};
        const cuuint64_t dims[3] = {64, (cuuint64_t)N2, (cuuint64_t)K2 / 64};
        const cuuint64_t str[2] = {(cuuint64_t)K2 * 2, 128};
        const cuuint32_t box2[3] = {64, 16, 16};
        enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, Wall, dims, str, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int stage_bytes = 32768;
        for (int inflight_kb : {64, 96, 128, 192}) {
            const int nstages = inflight_kb * 1024 / stage_bytes;
            const size_t smem = (size_t)nstages * stage_bytes;
            CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const long long total = (long long)N2 * K2 * 2 / stage_bytes;
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                for (int it = 0; it < 10; ++it) stream_kernel<<<148, 64, smem>>>(tm, Wall, 2, nstages, stage_bytes, total, K2, N2, 50);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                ms /= 500;
                if (rep > 0 && ms < best) best = ms;
            }
            printf("L2-resident 33.5 MB, 3-D box 32 KB, in flight %3d KB/SM (%2d stages): %7.1f us per pass  %6.0f GB/s (50 passes per launch)\n", inflight_kb, nstages,
                   best * 1e3, (double)N2 * K2 * 2 / (best * 1e-3) / 1e9);
        }
    }
    return 0;
}
