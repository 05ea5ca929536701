// BERT post-LN residual blocks as ONE kernel: out = LayerNorm(A . W^T + bias + resid) -> bf16 (sm_100a, tcgen05 / TMEM / TMA).
//
// The reference computes `output.dense(x)`, adds the block input and applies LayerNorm (src/models/embeddings.rs:167-190 and
// :232-241): in round 1 a GEMM wrote the f32 sum [T, H] and a second kernel normalised it -- 50 MB written and read back per
// block at the BASELINE batch, through an epilogue whose thread-per-row f32 stores are the slowest thing in the encoder.  Here a
// CTA owns 128 COMPLETE rows: the H (<= 512) output columns of a row tile are two MMAs of N = H / 2 into one TMEM accumulator
// (H f32 columns of the 512), so the epilogue sees whole rows, computes mean / variance from TMEM (pass 1), and writes the
// normalised bf16 rows (pass 2).  Same arithmetic as gemm_tc_kernel<GEPI_BIAS_RESID_F32> + layernorm_kernel: f32 accumulate, bias,
// bf16 residual, one-pass variance E[x^2] - mean^2, eps from the config.
// Roles as in gemm_tc.cuh: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-17 = epilogue (thread == row of its
// TMEM lane quarter, four warps per quarter share a row's columns).
#pragma once
#include "gemm_tc.cuh"

namespace fl {

struct BertLnGemmArgs {
    int M, H, K;
    const float* bias;       // [H]
    const uint16_t* resid;   // bf16 [M, H]
    const float* lnw;        // [H]
    const float* lnb;        // [H]
    float eps;
    uint16_t* out;           // bf16 [M, H]
};

constexpr int kLnGemmMaxStages = 4;

inline int bert_ln_gemm_stages(int H) { return std::min<int>(kLnGemmMaxStages, 196608 / (kGemmBM * kGemmBK * 2 + H * kGemmBK * 2)); }
inline size_t bert_ln_gemm_smem(int H) { return (size_t)bert_ln_gemm_stages(H) * (kGemmBM * kGemmBK * 2 + H * kGemmBK * 2) + 1024; }

// tmA: activations [M, K], box {64, 128}; tmB: weights [H, K], box {64, H / 2}
static __global__ void __launch_bounds__(kGemmThreads, 1)
bert_gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const BertLnGemmArgs g) {
    const int halfN = g.H >> 1;
    const uint32_t kABytes = kGemmBM * kGemmBK * 2, kBBytes = (uint32_t)g.H * kGemmBK * 2, kStageBytes = kABytes + kBBytes;
    const int nstages = min(kLnGemmMaxStages, (int)(196608u / kStageBytes));
    const uint32_t tmem_cols = g.H <= 128 ? 128u : (g.H <= 256 ? 256u : 512u);

    extern __shared__ uint8_t lsm_raw[];
    uint8_t* gsm = lsm_raw + ((1024u - (smem_u32(lsm_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[kLnGemmMaxStages], empty[kLnGemmMaxStages], acc_full, acc_empty;
    __shared__ uint32_t tmem_base_s;
    __shared__ float part[4][kGemmBM][2];       // per column slice: (sum, sum of squares) of every row of the tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (g.K + kGemmBK - 1) / kGemmBK;
    const int mt = (g.M + kGemmBM - 1) / kGemmBM;

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&acc_full, 1);
        mbar_init(&acc_empty, kGemmEpiWarps);
        mbar_fence_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            pdl_wait();
            asm volatile("fence.proxy.async;" ::: "memory");
            uint32_t c = 0;
            for (int tile = blockIdx.x; tile < mt; tile += gridDim.x) {
                const int m0 = tile * kGemmBM;
                for (int kb = 0; kb < nk; ++kb, ++c) {
                    const int st = c % nstages;
                    uint8_t* sa = gsm + (size_t)st * kStageBytes;
                    mbar_wait(&empty[st], ((c / nstages) & 1) ^ 1);
                    mbar_expect_tx(&full[st], kStageBytes);
                    tma_load_2d(sa, &tmA, kb * kGemmBK, m0, &full[st]);
                    tma_load_2d(sa + kABytes, &tmB, kb * kGemmBK, 0, &full[st]);
                    tma_load_2d(sa + kABytes + kBBytes / 2, &tmB, kb * kGemmBK, halfN, &full[st]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kGemmBM, halfN);
            uint32_t c = 0, ti = 0;
            for (int tile = blockIdx.x; tile < mt; tile += gridDim.x, ++ti) {
                mbar_wait(&acc_empty, (ti & 1) ^ 1);          // the epilogue has drained the accumulator of the previous tile
                tc_fence_after();
                for (int kb = 0; kb < nk; ++kb, ++c) {
                    const int st = c % nstages;
                    mbar_wait(&full[st], (c / nstages) & 1);
                    tc_fence_after();
                    const uint8_t* sa = gsm + (size_t)st * kStageBytes;
                    const uint64_t adesc = umma_smem_desc_sw128(sa), b0 = umma_smem_desc_sw128(sa + kABytes),
                                   b1 = umma_smem_desc_sw128(sa + kABytes + kBBytes / 2);
#pragma unroll
                    for (int k = 0; k < kGemmBK / 16; ++k) {
                        umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                        umma_bf16(tmem_base + (uint32_t)halfN, adesc + (uint64_t)(2 * k), b1 + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&empty[st]);
                }
                umma_commit(&acc_full);
            }
        }
    } else {
        const int q = warp & 3, cslice = (warp - 2) >> 2;
        const int cw = g.H >> 2;                              // columns per slice (multiple of 32)
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < mt; tile += gridDim.x, ++ti) {
            const int rl = q * 32 + lane, row = tile * kGemmBM + rl;
            const bool live = row < g.M;
            const uint16_t* rs = g.resid + (size_t)(live ? row : 0) * g.H;
            mbar_wait(&acc_full, ti & 1);
            tc_fence_after();
            // ---- pass 1: mean / variance of (acc + bias + resid) over the row ----
            float s = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c0 = cslice * cw; c0 < (cslice + 1) * cw; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(g.bias + c0 + j);
                    const uint2 rr = live ? *reinterpret_cast<const uint2*>(rs + c0 + j) : make_uint2(0u, 0u);
                    const float v0 = __uint_as_float(r[j]) + bv.x + bf16lo(rr.x), v1 = __uint_as_float(r[j + 1]) + bv.y + bf16hi(rr.x);
                    const float v2 = __uint_as_float(r[j + 2]) + bv.z + bf16lo(rr.y), v3 = __uint_as_float(r[j + 3]) + bv.w + bf16hi(rr.y);
                    s += (v0 + v1) + (v2 + v3);
                    s2 = fmaf(v0, v0, s2); s2 = fmaf(v1, v1, s2); s2 = fmaf(v2, v2, s2); s2 = fmaf(v3, v3, s2);
                }
            }
            part[cslice][rl][0] = s;
            part[cslice][rl][1] = s2;
            asm volatile("bar.sync 2, %0;" ::"n"(32 * kGemmEpiWarps) : "memory");
            const float ts = (part[0][rl][0] + part[1][rl][0]) + (part[2][rl][0] + part[3][rl][0]);
            const float ts2 = (part[0][rl][1] + part[1][rl][1]) + (part[2][rl][1] + part[3][rl][1]);
            const float mean = ts / (float)g.H;
            const float inv = 1.f / sqrtf(ts2 / (float)g.H - mean * mean + g.eps);
            // ---- pass 2: normalise and store bf16 ----
            uint16_t* orow = g.out + (size_t)(live ? row : 0) * g.H;
#pragma unroll 1
            for (int c0 = cslice * cw; c0 < (cslice + 1) * cw; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 8; e += 4) {
                        const float4 bv = *reinterpret_cast<const float4*>(g.bias + c0 + j + e);
                        const float4 wv = *reinterpret_cast<const float4*>(g.lnw + c0 + j + e);
                        const float4 lb = *reinterpret_cast<const float4*>(g.lnb + c0 + j + e);
                        const uint2 rr = live ? *reinterpret_cast<const uint2*>(rs + c0 + j + e) : make_uint2(0u, 0u);
                        const float v0 = __uint_as_float(r[j + e]) + bv.x + bf16lo(rr.x), v1 = __uint_as_float(r[j + e + 1]) + bv.y + bf16hi(rr.x);
                        const float v2 = __uint_as_float(r[j + e + 2]) + bv.z + bf16lo(rr.y), v3 = __uint_as_float(r[j + e + 3]) + bv.w + bf16hi(rr.y);
                        pk[e >> 1] = pack_bf16x2((v0 - mean) * inv * wv.x + lb.x, (v1 - mean) * inv * wv.y + lb.y);
                        pk[(e >> 1) + 1] = pack_bf16x2((v2 - mean) * inv * wv.z + lb.z, (v3 - mean) * inv * wv.w + lb.w);
                    }
                    if (live) *reinterpret_cast<uint4*>(orow + c0 + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            // the tile's accumulator has been read twice by every warp: hand it back (and keep `part` intact until all have read it)
            tc_fence_before();
            asm volatile("bar.sync 2, %0;" ::"n"(32 * kGemmEpiWarps) : "memory");
            if (lane == 0) mbar_arrive(&acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---- weight-stationary GEMM for the K = hidden projections (q|k|v, intermediate) ---------------------------------------------
// D[M, N] = A[M, K] . B[N, K]^T + bias (-> GELU) -> bf16 with K <= 384.  At hidden 384 a 128 x 192 output tile of gemm_tc_kernel
// pulls 96 KB of activations AND 144 KB of weights through L2 -> shared memory for 18.9 MFLOP: the encoder's big GEMMs are bound by
// that fill rate (~47 GB/s per SM), 4x above their tensor time.  Here a CTA keeps ONE 192-row weight tile (all of K: <= 144 KB)
// resident in shared memory and walks the row tiles of its column block, so a tile costs only its 96 KB of activations: 2.5x less
// traffic per tile.  grid = (148 / nt) * nt CTAs, CTA c owns column block c % nt and row tiles c / nt, + grid / nt, ...
constexpr int kBresBN = 192;
constexpr int kBresMaxNk = 6;        // K <= 384
constexpr int kBresStages = 5;       // activation ring (16 KB stages) next to the resident weights

inline size_t bert_bres_smem(int K) { return (size_t)((K + kGemmBK - 1) / kGemmBK) * kBresBN * kGemmBK * 2 + (size_t)kBresStages * kGemmBM * kGemmBK * 2 + 1024; }

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
bert_gemm_bres_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
    static_assert(EPI == GEPI_BIAS_BF16 || EPI == GEPI_BIAS_GELU_BF16, "weight-stationary variant: bf16 outputs only");
    constexpr int BN = kBresBN;
    constexpr uint32_t kABytes = kGemmBM * kGemmBK * 2, kBBytes = BN * kGemmBK * 2;
    constexpr uint32_t kAccCols = 256, kTmemCols = 512;

    extern __shared__ uint8_t bsm_raw[];
    uint8_t* gsm = bsm_raw + ((1024u - (smem_u32(bsm_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bfull, full[kBresStages], empty[kBresStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (g.K + kGemmBK - 1) / kGemmBK;
    const int mt = (g.M + kGemmBM - 1) / kGemmBM, nt = (g.N + BN - 1) / BN;
    const int groups = gridDim.x / nt, tile_n = blockIdx.x % nt, m_first = blockIdx.x / nt;
    const int n0 = tile_n * BN;
    uint8_t* resB = gsm;                                   // [nk][BN x 128 B]
    uint8_t* ring = gsm + (size_t)nk * kBBytes;            // [kBresStages][128 x 128 B]

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        mbar_init(&bfull, 1);
        for (int s = 0; s < kBresStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], kGemmEpiWarps);
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            // the weights do not depend on the previous kernel: they are requested before the dependency wait
            mbar_expect_tx(&bfull, (uint32_t)nk * kBBytes);
            for (int kb = 0; kb < nk; ++kb) tma_load_2d(resB + (size_t)kb * kBBytes, &tmB, kb * kGemmBK, n0, &bfull);
            pdl_wait();
            asm volatile("fence.proxy.async;" ::: "memory");
            uint32_t c = 0;
            for (int tm = m_first; tm < mt; tm += groups) {
                for (int kb = 0; kb < nk; ++kb, ++c) {
                    const int st = c % kBresStages;
                    mbar_wait(&empty[st], ((c / kBresStages) & 1) ^ 1);
                    mbar_expect_tx(&full[st], kABytes);
                    tma_load_2d(ring + (size_t)st * kABytes, &tmA, kb * kGemmBK, tm * kGemmBM, &full[st]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, BN);
            uint32_t c = 0, ti = 0;
            mbar_wait(&bfull, 0);
            tc_fence_after();
            for (int tm = m_first; tm < mt; tm += groups, ++ti) {
                const uint32_t as = ti & 1;
                mbar_wait(&acc_empty[as], ((ti >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + as * kAccCols;
                for (int kb = 0; kb < nk; ++kb, ++c) {
                    const int st = c % kBresStages;
                    mbar_wait(&full[st], (c / kBresStages) & 1);
                    tc_fence_after();
                    const uint64_t adesc = umma_smem_desc_sw128(ring + (size_t)st * kABytes), bdesc = umma_smem_desc_sw128(resB + (size_t)kb * kBBytes);
#pragma unroll
                    for (int k = 0; k < kGemmBK / 16; ++k)
                        umma_bf16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&empty[st]);
                }
                umma_commit(&acc_full[as]);
            }
        }
    } else {
        const int q = warp & 3, cslice = (warp - 2) >> 2;
        uint32_t ti = 0;
        for (int tm = m_first; tm < mt; tm += groups, ++ti) {
            const uint32_t as = ti & 1;
            const int row = tm * kGemmBM + q * 32 + lane;
            mbar_wait(&acc_full[as], (ti >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = cslice * 32; c0 < BN; c0 += 128) {
                uint32_t r[32];
                tmem_ld32(tmem_base + as * kAccCols + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                if (row < g.M) {
                    const int col = n0 + c0;
                    uint16_t* o = reinterpret_cast<uint16_t*>(g.out) + (size_t)row * g.ldo + col;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        if (col + j < g.N) {
                            uint32_t pk[4];
                            const float4 b0 = g.bias ? *reinterpret_cast<const float4*>(g.bias + col + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float4 b1 = g.bias ? *reinterpret_cast<const float4*>(g.bias + col + j + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                float v0 = __uint_as_float(r[j + e]) + bb[e];
                                float v1 = __uint_as_float(r[j + e + 1]) + bb[e + 1];
                                if (EPI == GEPI_BIAS_GELU_BF16) { v0 = gelu_tanh_f(v0); v1 = gelu_tanh_f(v1); }
                                pk[e >> 1] = pack_bf16x2(v0, v1);
                            }
                            *reinterpret_cast<uint4*>(o + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace fl
