// HBM-bound bf16 weight-streaming GEMV family for decode (1-2 activation rows), sm_100a.
//
// Replaces, per decoder layer, the reference's cuBLAS-through-candle matmuls plus the element-wise kernels around
// them: K2+K3+K4+K5+K6 (RMSNorm -> q/k/v Linear(+bias) -> RoPE -> KV append) in ONE kernel,
// K12+K13 (o_proj + residual), K14+K15 (RMSNorm -> gate/up -> SiLU*up), K16 (down_proj + residual) and
// K17 (final RMSNorm -> lm_head -> f32 logits + arg-max partials).
//
// Design (DESIGN.md section "GEMV"):
//   * y[n] = sum_k W[n,k] x[k];  W is bf16 [N,K] row-major and is read exactly once per step with 128-bit
//     ld.global.nc.L1::no_allocate loads.  One CTA per SM (grid = 148 x ctas_per_sm), each owning a contiguous,
//     even-aligned slice of rows, so every kernel is a single resident wave.
//   * Every thread owns CPT fixed 8-element k-chunks of the row: the activation slice it multiplies with lives in
//     REGISTERS (f32), so the hot loop is LDG.128 + 8 cvt + 8*M FMA per chunk with no shared-memory traffic; a warp
//     reads 512 contiguous bytes of a row, the CTA reads the whole row.
//   * R rows are in flight per thread (R*CPT = 8 x 16 B), warp-shuffle reduce per row, per-warp partials parked in
//     shared memory and summed across warps once per 256-row super-chunk (no barrier inside the streaming loop).
//   * f32 activations, f32 accumulation, f32 residual stream: only the weights are bf16 (oracle = f32 math on
//     bf16-rounded weights).
//   * PDL: the first R rows of weights are requested BEFORE griddepcontrol.wait, overlapping the producer's tail.
#pragma once
#include "common.cuh"

namespace fl {

enum : int { PRO_PLAIN = 0, PRO_RMSNORM = 1 };
enum : int { EPI_STORE = 0, EPI_RESID = 1, EPI_SILU = 2, EPI_QKV = 3 };

// Device-resident step state of one KV cache (so a captured CUDA graph can be replayed as positions advance).
struct StepState {
    int rope_pos;             // RoPE position of token 0 of the current forward call
    int ragged;               // 1 during a fl_forward_slots call: batch row i is cache slot slot[i] at RoPE position rope_seq[i]
    int pad[2];
    int kv_base[kMaxBatch];   // tokens already in the cache per sequence SLOT, BEFORE the current call
    int slot[kMaxBatch];      // ragged calls: cache slot of batch row i (rows of the uniform calls are their own slots)
    int rope_seq[kMaxBatch];  // ragged calls: RoPE position of token 0 of batch row i
};
// cache slot / RoPE position of batch row `seq` (continuous batching: sequences of different lengths share a step)
__device__ __forceinline__ int st_slot(const StepState* s, int seq) { return s->ragged ? s->slot[seq] : seq; }
__device__ __forceinline__ int st_rope(const StepState* s, int seq) { return s->ragged ? s->rope_seq[seq] : s->rope_pos; }

struct GemvArgs {
    const uint16_t* W;   // [N, K] bf16
    int N, K;
    const float* x;      // PRO_PLAIN: activations [M, K]; PRO_RMSNORM: residual stream [M, K]
    const float* norm_w; // PRO_RMSNORM: [K]
    const float* delta;  // PRO_RMSNORM, optional [M, K]: x = resid + delta is normalised and CTA 0 writes the advanced residual
    float* resid_out;    // stream to resid_out (a DIFFERENT buffer: other CTAs are still reading `x`).  Tensor parallelism:
                         // delta = the all-reduced o_proj / down_proj output; this saves a separate residual-add kernel.
    float eps;
    float* out;          // EPI_STORE: [*, ldo]; EPI_RESID: residual [M, N] (+=); EPI_SILU: act [M, N/2]
    int ldo;
    const float* bias;   // optional [N] (EPI_STORE / EPI_QKV), already in the permuted row order of W
    // row bookkeeping: pass row m is global row row_base + m of the flattened [b, t] call
    int row_base, t;
    int last_only;       // EPI_STORE: only rows that are the LAST token of their sequence are stored, at out[seq*ldo + n]
    float* amax_val;     // EPI_STORE optional arg-max partials [M, gridDim.x]
    int* amax_idx;
    // EPI_QKV
    float* q_out;        // [M, nh*d] f32 after bias+RoPE
    uint16_t* kpool;     // this layer's K pages [page][nkv][kKvPage][d] bf16
    uint16_t* vpool;
    const int* page_table;  // [b, pt_stride]
    int pt_stride;
    const StepState* state;
    const float* rope_cos;  // [max_pos, d/2]
    const float* rope_sin;
    int nh, nkv, d, max_pos;
};

constexpr int kGemvSuper = 256;   // rows per cross-warp reduction round
constexpr int kGemvMaxWarps = 16;

__device__ __forceinline__ float dot8(const uint4& w, const float (&x)[8], float acc) {
    acc = fmaf(bf16lo(w.x), x[0], acc);
    acc = fmaf(bf16hi(w.x), x[1], acc);
    acc = fmaf(bf16lo(w.y), x[2], acc);
    acc = fmaf(bf16hi(w.y), x[3], acc);
    acc = fmaf(bf16lo(w.z), x[4], acc);
    acc = fmaf(bf16hi(w.z), x[5], acc);
    acc = fmaf(bf16lo(w.w), x[6], acc);
    acc = fmaf(bf16hi(w.w), x[7], acc);
    return acc;
}

template <int M, int CPT, int PRO, int EPI>
__global__ void __launch_bounds__(512, 1) gemv_kernel(const GemvArgs a) {
    constexpr int R = (CPT == 1) ? 8 : (CPT == 2 ? 4 : 2);
    __shared__ float partial[kGemvSuper * kGemvMaxWarps * M];
    __shared__ float red[M * 32];
    __shared__ float red_amax_v[M * kGemvMaxWarps];
    __shared__ int red_amax_i[M * kGemvMaxWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int K8 = a.K >> 3;
    const int npairs = a.N >> 1;
    const int row_begin = 2 * (int)((int64_t)npairs * blockIdx.x / gridDim.x);
    const int row_end = 2 * (int)((int64_t)npairs * (blockIdx.x + 1) / gridDim.x);

    int c[CPT];
    bool cv[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        c[j] = tid + j * blockDim.x;
        cv[j] = c[j] < K8;
    }
    const uint4* Wv = reinterpret_cast<const uint4*>(a.W);

    uint4 wb[R][CPT];
    auto load_group = [&](int row0) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                if (row0 + r < row_end && cv[j])
                    wb[r][j] = ldg_stream(Wv + (size_t)(row0 + r) * K8 + c[j]);
                else
                    wb[r][j] = make_uint4(0, 0, 0, 0);
            }
    };

    // Weights do not depend on the previous kernel: request the first group now, then wait for the producer.
    load_group(row_begin);
    pdl_launch_dependents();
    pdl_wait();

    // ---- prologue: this thread's slice of the activation row(s) into registers ----
    float xr[M][CPT][8];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            if (cv[j]) {
                const float4* xp = reinterpret_cast<const float4*>(a.x + (size_t)m * a.K + (size_t)c[j] * 8);
                float4 v0 = xp[0], v1 = xp[1];
                xr[m][j][0] = v0.x; xr[m][j][1] = v0.y; xr[m][j][2] = v0.z; xr[m][j][3] = v0.w;
                xr[m][j][4] = v1.x; xr[m][j][5] = v1.y; xr[m][j][6] = v1.z; xr[m][j][7] = v1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) xr[m][j][i] = 0.f;
            }
        }
    if (PRO == PRO_RMSNORM && a.delta != nullptr) {
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int j = 0; j < CPT; ++j)
                if (cv[j]) {
                    const float4* dp = reinterpret_cast<const float4*>(a.delta + (size_t)m * a.K + (size_t)c[j] * 8);
                    const float4 d0 = dp[0], d1 = dp[1];
                    xr[m][j][0] += d0.x; xr[m][j][1] += d0.y; xr[m][j][2] += d0.z; xr[m][j][3] += d0.w;
                    xr[m][j][4] += d1.x; xr[m][j][5] += d1.y; xr[m][j][6] += d1.z; xr[m][j][7] += d1.w;
                    if (blockIdx.x == 0) {   // the residual stream itself is advanced exactly once
                        float4* rp = reinterpret_cast<float4*>(a.resid_out + (size_t)m * a.K + (size_t)c[j] * 8);
                        rp[0] = make_float4(xr[m][j][0], xr[m][j][1], xr[m][j][2], xr[m][j][3]);
                        rp[1] = make_float4(xr[m][j][4], xr[m][j][5], xr[m][j][6], xr[m][j][7]);
                    }
                }
    }
    if (PRO == PRO_RMSNORM) {
        // candle_nn::ops::rms_norm: m = sqrt(sum(x^2)/n + eps); y = x / m * w
        float ss[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < CPT; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) s = fmaf(xr[m][j][i], xr[m][j][i], s);
            s = warp_sum(s);
            if (lane == 0) red[m * 32 + warp] = s;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float s = 0.f;
            for (int w = 0; w < nwarps; ++w) s += red[m * 32 + w];
            ss[m] = sqrtf(s / (float)a.K + a.eps);
        }
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            if (cv[j]) {
                const float4* wp = reinterpret_cast<const float4*>(a.norm_w + (size_t)c[j] * 8);
                float4 w0 = wp[0], w1 = wp[1];
                float nw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int i = 0; i < 8; ++i) xr[m][j][i] = xr[m][j][i] / ss[m] * nw[i];
            }
        }
    }

    float best_v[M];
    int best_i[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        best_v[m] = -INFINITY;
        best_i[m] = -1;
    }

    for (int sc0 = row_begin; sc0 < row_end; sc0 += kGemvSuper) {
        const int sc1 = min(sc0 + kGemvSuper, row_end);
        // ---- streaming loop: no barrier inside ----
        for (int row0 = sc0; row0 < sc1; row0 += R) {
            if (row0 != row_begin) load_group(row0);
            float acc[R][M];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int m = 0; m < M; ++m) acc[r][m] = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < CPT; ++j)
#pragma unroll
                    for (int m = 0; m < M; ++m) acc[r][m] = dot8(wb[r][j], xr[m][j], acc[r][m]);
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    float v = warp_sum(acc[r][m]);
                    if (lane == 0 && row0 + r < sc1) partial[((row0 - sc0 + r) * nwarps + warp) * M + m] = v;
                }
        }
        __syncthreads();
        // ---- cross-warp reduce + fused epilogue, one thread per row pair ----
        const int pairs_here = (sc1 - sc0) >> 1;
        for (int e = tid; e < pairs_here; e += blockDim.x) {
            const int ra = sc0 + 2 * e;
            float ya[M], yb[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                float sa = 0.f, sb = 0.f;
                for (int w = 0; w < nwarps; ++w) {
                    sa += partial[((2 * e) * nwarps + w) * M + m];
                    sb += partial[((2 * e + 1) * nwarps + w) * M + m];
                }
                ya[m] = sa;
                yb[m] = sb;
            }
            if (EPI == EPI_STORE) {
                const float ba = a.bias ? a.bias[ra] : 0.f, bb = a.bias ? a.bias[ra + 1] : 0.f;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float va = ya[m] + ba, vb = yb[m] + bb;
                    int orow = m;
                    bool keep = true;
                    if (a.last_only) {
                        const int rg = a.row_base + m;
                        orow = rg / a.t;
                        keep = (rg % a.t) == a.t - 1;
                    }
                    if (keep) {
                        a.out[(size_t)orow * a.ldo + ra] = va;
                        a.out[(size_t)orow * a.ldo + ra + 1] = vb;
                    }
                    // candle arg-max: max_by(total_cmp) keeps the LAST index among equal maxima
                    if (va >= best_v[m]) { best_v[m] = va; best_i[m] = ra; }
                    if (vb >= best_v[m]) { best_v[m] = vb; best_i[m] = ra + 1; }
                }
            } else if (EPI == EPI_RESID) {
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    float2* p = reinterpret_cast<float2*>(a.out + (size_t)m * a.N + ra);
                    float2 v = *p;
                    v.x += ya[m];
                    v.y += yb[m];
                    *p = v;
                }
            } else if (EPI == EPI_SILU) {
                // rows are interleaved at upload: even row = gate_j, odd row = up_j  ->  act[j] = silu(gate) * up
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float g = ya[m];
                    a.out[(size_t)m * (a.N >> 1) + (ra >> 1)] = g / (1.f + expf(-g)) * yb[m];
                }
            } else {  // EPI_QKV: bias + RoPE (rotate-half) + q store + KV-cache append
                const int d = a.d, half = d >> 1;
                const int hh = ra / d, j = (ra % d) >> 1;
                const float ba = a.bias ? a.bias[ra] : 0.f, bb = a.bias ? a.bias[ra + 1] : 0.f;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const int rg = a.row_base + m;
                    const int seq = rg / a.t, irel = rg % a.t;
                    const float va = ya[m] + ba, vb = yb[m] + bb;
                    const int slot = a.state->kv_base[seq] + irel;
                    const int page = a.page_table[seq * a.pt_stride + slot / kKvPage];
                    if (hh < a.nh + a.nkv) {
                        // rows were permuted at upload so that (ra, ra+1) = elements (j, j + d/2) of one head
                        int pos = a.state->rope_pos + irel;
                        pos = pos < a.max_pos ? pos : a.max_pos - 1;
                        const float cs = a.rope_cos[(size_t)pos * half + j], sn = a.rope_sin[(size_t)pos * half + j];
                        const float o1 = va * cs - vb * sn, o2 = va * sn + vb * cs;
                        if (hh < a.nh) {
                            float* q = a.q_out + ((size_t)m * a.nh + hh) * d;
                            q[j] = o1;
                            q[j + half] = o2;
                        } else {
                            uint16_t* kp = a.kpool + (((size_t)page * a.nkv + (hh - a.nh)) * kKvPage + slot % kKvPage) * d;
                            kp[j] = f32_to_bf16_rne(o1);
                            kp[j + half] = f32_to_bf16_rne(o2);
                        }
                    } else {
                        uint16_t* vp = a.vpool + (((size_t)page * a.nkv + (hh - a.nh - a.nkv)) * kKvPage + slot % kKvPage) * d;
                        const uint32_t packed = (uint32_t)f32_to_bf16_rne(va) | ((uint32_t)f32_to_bf16_rne(vb) << 16);
                        *reinterpret_cast<uint32_t*>(vp + 2 * j) = packed;
                    }
                }
            }
        }
        if (sc1 < row_end) __syncthreads();
    }

    if (EPI == EPI_STORE && a.amax_val != nullptr) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float v = best_v[m];
            int i = best_i[m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xFFFFFFFFu, v, o);
                const int oi = __shfl_xor_sync(0xFFFFFFFFu, i, o);
                if (ov > v || (ov == v && oi > i)) { v = ov; i = oi; }
            }
            if (lane == 0) { red_amax_v[m * kGemvMaxWarps + warp] = v; red_amax_i[m * kGemvMaxWarps + warp] = i; }
        }
        __syncthreads();
        if (tid < M) {
            float v = -INFINITY;
            int i = -1;
            for (int w = 0; w < nwarps; ++w) {
                const float ov = red_amax_v[tid * kGemvMaxWarps + w];
                const int oi = red_amax_i[tid * kGemvMaxWarps + w];
                if (ov > v || (ov == v && oi > i)) { v = ov; i = oi; }
            }
            a.amax_val[tid * gridDim.x + blockIdx.x] = v;
            a.amax_idx[tid * gridDim.x + blockIdx.x] = i;
        }
    }
}

}  // namespace fl
