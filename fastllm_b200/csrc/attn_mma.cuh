// Tensor-core attention over the paged bf16 KV cache for the dense path (3+ activation rows), sm_100a.
//
//   attn_gqa_decode_kernel<D>  : batched decode (one new token per sequence).  One CTA per (split, kv head, sequence) streams
//       its K/V pages ONCE (cp.async 16-byte copies into an XOR-swizzled 2-stage shared-memory ring) and serves all n_rep
//       query heads of the GQA group from the staged copy: the heads are the M rows (padded to 16) of mma.sync m16n8k16,
//       each of the 4 warps owns 16 of the 64 keys of a page, online softmax in f32 per warp, warps merged through shared
//       memory, splits merged by the last-arriving CTA (same partial format as attn_decode_kernel).  HBM-bound: the MMAs
//       replace the per-key shuffle reductions that limited attn_decode_kernel to ~1.2 TB/s at batch 64.
//   attn_prefill_kernel<D>     : multi-token calls (prefill).  One CTA per (64-query tile, q head, sequence): FlashAttention-
//       style loop over the causal range of 64-key pages, S = Q K^T and O += P V on mma.sync, online softmax in registers.
//       Replaces the row-per-CTA decode kernel in prefill (84 % of the Qwen2.5-7B 4k prefill before).
//
// Precision: q and the softmax probabilities are f32 in the oracle.  Both are split hi + lo into two bf16 operands and
// multiplied against the same K / V fragments (as the dense GEMMs do with their activations), K/V are bf16 in the cache
// already, accumulation is f32: results stay inside the kernel tolerance of the f32 oracle.
// Semantics: scores * 1/sqrt(d) after the matmul, causal rows for multi-token calls,
// Mistral/Qwen2 sliding-window rule inside the prefill only (key j banned when j + sw < i, new tokens only; cached keys stay visible), softmax max-subtract/exp/sum/div.
#pragma once
#include "attn_decode.cuh"
#include "dense_ops.cuh"
#include "mma.cuh"

namespace fl {

constexpr int kMmaAttnThreads = 128;
// K|V page stages of the batched-decode kernel.  Measured at batch 64 (Mistral-7B, 2k context): 2 stages x 3 resident CTAs per SM
// = 112 us per layer, 3 stages x 2 CTAs = 121 us -- occupancy (independent latency chains) beats a deeper ring per CTA.
constexpr int kDecStages = 2;
constexpr int kPrefillBM = 64;       // query rows per CTA of the prefill kernel (4 warps x 16)

// element offset of (row, col) in a [rows][D] bf16 tile whose 16-byte chunks are XOR-swizzled by the row (conflict-free ldmatrix)
template <int D>
__device__ __forceinline__ int swz(int row, int col) {
    constexpr int CH = D / 8;
    constexpr int MASK = (CH < 8 ? CH : 8) - 1;
    return row * D + ((((col >> 3) ^ (row & MASK))) << 3) + (col & 7);
}

template <int D>
__device__ __forceinline__ void stage_kv_tile(uint16_t* kd, uint16_t* vd, const uint16_t* ksrc, const uint16_t* vsrc, int tid) {
    constexpr int CH = D / 8;
#pragma unroll
    for (int i = tid; i < kKvPage * CH; i += kMmaAttnThreads) {
        const int r = i / CH, c = i % CH;
        cp_async16(kd + swz<D>(r, c * 8), ksrc + r * D + c * 8);
        cp_async16(vd + swz<D>(r, c * 8), vsrc + r * D + c * 8);
    }
    cp_async_commit();
}

// f32 q rows -> hi/lo bf16 tiles [NR][D] (rows >= valid rows are zero)
__device__ __forceinline__ void store_hi_lo(uint16_t* hi, uint16_t* lo, int off, float x) {
    uint16_t h, l;
    split_hi_lo(x, h, l);
    hi[off] = h;
    lo[off] = l;
}

// writes one output element either as f32 or as the hi/lo bf16 pair the next GEMM consumes
__device__ __forceinline__ void store_out2(const AttnArgs& a, size_t idx, float x0, float x1) {
    if (a.out_hi) {
        uint16_t h0, l0, h1, l1;
        split_hi_lo(x0, h0, l0);
        split_hi_lo(x1, h1, l1);
        *reinterpret_cast<uint32_t*>(a.out_hi + idx) = (uint32_t)h0 | ((uint32_t)h1 << 16);
        *reinterpret_cast<uint32_t*>(a.out_lo + idx) = (uint32_t)l0 | ((uint32_t)l1 << 16);
    } else {
        *reinterpret_cast<float2*>(a.out + idx) = make_float2(x0, x1);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// batched decode: grid (nsplit, nkv, b), 128 threads
// ---------------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kMmaAttnThreads) attn_gqa_decode_kernel(const AttnArgs a) {
    constexpr int TILE = kKvPage * D;
    constexpr int NW = kMmaAttnThreads / 32;
    extern __shared__ __align__(128) uint8_t dsm[];
    uint16_t* qhi = reinterpret_cast<uint16_t*>(dsm);   // [16][D]
    uint16_t* qlo = qhi + 16 * D;
    uint16_t* kv = qlo + 16 * D;                        // [kDecStages][K tile | V tile]
    __shared__ float red_m[NW][16], red_l[NW][16], mrg_w[NW][16], mrg_den[16];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int split = blockIdx.x, nsplit = gridDim.x, kvh = blockIdx.y, seq = blockIdx.z;
    const int n_rep = a.nh / a.nkv;
    pdl_launch_dependents();
    pdl_wait();

    const int cslot = st_slot(a.state, seq);          // ragged batches: every sequence has its own cache slot and length
    const int len = a.state->kv_base[cslot] + 1;
    const int npages = (len + kKvPage - 1) / kKvPage;
    const int per = (npages + nsplit - 1) / nsplit;
    const int p0 = split * per, p1 = min(p0 + per, npages);
    const int* pt = a.page_table + (size_t)cslot * a.pt_stride;

    auto load = [&](int p, int st) {
        const size_t off = ((size_t)pt[p] * a.nkv + kvh) * (size_t)TILE;
        stage_kv_tile<D>(kv + st * 2 * TILE, kv + st * 2 * TILE + TILE, a.kpool + off, a.vpool + off, tid);
    };
    // one cp.async group per page, committed even when there is no page left, so "all but the newest kDecStages - 1 groups" always
    // means "the page about to be consumed has landed"; kDecStages - 1 pages are in flight per CTA while one is consumed
    auto load_or_skip = [&](int p) {
        if (p < p1) load(p, (p - p0) % kDecStages);
        else cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < kDecStages - 1; ++i) load_or_skip(p0 + i);
    for (int i = tid; i < 16 * D; i += kMmaAttnThreads) {
        const int h = i / D, dd = i % D;
        const float q = h < n_rep ? a.q[((size_t)seq * a.nh + kvh * n_rep + h) * D + dd] : 0.f;
        store_hi_lo(qhi, qlo, swz<D>(h, dd), q);
    }

    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int p = p0; p < p1; ++p) {
        const int st = (p - p0) % kDecStages;
        load_or_skip(p + kDecStages - 1);      // into the stage page p - 1 occupied (released by the barrier that ended its iteration)
        cp_async_wait<kDecStages - 1>();
        __syncthreads();
        const uint16_t* kt = kv + st * 2 * TILE;
        const uint16_t* vt = kt + TILE;
        const int key0 = p * kKvPage + warp * 16;
        if (key0 < len) {
            float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
                uint32_t ah[4], al[4], kb[4];
                ldmatrix_x4(ah, qhi + swz<D>(lane & 15, ks * 16 + (lane >> 4) * 8));
                ldmatrix_x4(al, qlo + swz<D>(lane & 15, ks * 16 + (lane >> 4) * 8));
                ldmatrix_x4(kb, kt + swz<D>(warp * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 16 + ((lane >> 3) & 1) * 8));
                mma_bf16_16816(s[0], ah, kb[0], kb[1]);
                mma_bf16_16816(s[0], al, kb[0], kb[1]);
                mma_bf16_16816(s[1], ah, kb[2], kb[3]);
                mma_bf16_16816(s[1], al, kb[2], kb[3]);
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int key = key0 + nt * 8 + tq * 2 + (e & 1);
                    s[nt][e] = key < len ? s[nt][e] * a.qscale : -INFINITY;
                    if (e < 2) mx0 = fmaxf(mx0, s[nt][e]); else mx1 = fmaxf(mx1, s[nt][e]);
                }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 2));
            const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);       // finite: key0 < len is valid for every row
            const float al0 = expf(m0 - mn0), al1 = expf(m1 - mn1);
            m0 = mn0; m1 = mn1;
            float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                s[nt][0] = expf(s[nt][0] - mn0); s[nt][1] = expf(s[nt][1] - mn0);
                s[nt][2] = expf(s[nt][2] - mn1); s[nt][3] = expf(s[nt][3] - mn1);
                ps0 += s[nt][0] + s[nt][1];
                ps1 += s[nt][2] + s[nt][3];
            }
            l0 = l0 * al0 + ps0;       // per-thread partial row sums; the quad is reduced once at the end
            l1 = l1 * al1 + ps1;
            uint32_t ph[4], pl[4];
            ph[0] = pack_bf16x2(s[0][0], s[0][1]); ph[1] = pack_bf16x2(s[0][2], s[0][3]);
            ph[2] = pack_bf16x2(s[1][0], s[1][1]); ph[3] = pack_bf16x2(s[1][2], s[1][3]);
            pl[0] = pack_bf16x2(s[0][0] - bf16lo(ph[0]), s[0][1] - bf16hi(ph[0]));
            pl[1] = pack_bf16x2(s[0][2] - bf16lo(ph[1]), s[0][3] - bf16hi(ph[1]));
            pl[2] = pack_bf16x2(s[1][0] - bf16lo(ph[2]), s[1][1] - bf16hi(ph[2]));
            pl[3] = pack_bf16x2(s[1][2] - bf16lo(ph[3]), s[1][3] - bf16hi(ph[3]));
#pragma unroll
            for (int dt = 0; dt < D / 16; ++dt) {
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, vt + swz<D>(warp * 16 + (lane & 15), dt * 16 + (lane >> 4) * 8));
                float(&oa)[4] = o[2 * dt];
                float(&ob)[4] = o[2 * dt + 1];
                oa[0] *= al0; oa[1] *= al0; oa[2] *= al1; oa[3] *= al1;
                ob[0] *= al0; ob[1] *= al0; ob[2] *= al1; ob[3] *= al1;
                mma_bf16_16816(oa, ph, vb[0], vb[1]);
                mma_bf16_16816(oa, pl, vb[0], vb[1]);
                mma_bf16_16816(ob, ph, vb[2], vb[3]);
                mma_bf16_16816(ob, pl, vb[2], vb[3]);
            }
        }
        __syncthreads();   // everyone is done with stage st before the next iteration refills it
    }

    cp_async_wait<0>();
    // ---- merge the 4 warps (each saw a different quarter of every page) ----
    l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 1); l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 2);
    l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 1); l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 2);
    float* red_o = reinterpret_cast<float*>(kv);      // [NW][16][D] f32 (<= 32 KB, inside the 2-stage ring)
    if (tq == 0) {
        red_m[warp][g] = m0; red_m[warp][g + 8] = m1;
        red_l[warp][g] = l0; red_l[warp][g + 8] = l1;
    }
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        float* r0 = red_o + ((size_t)warp * 16 + g) * D + nt * 8 + tq * 2;
        float* r1 = red_o + ((size_t)warp * 16 + g + 8) * D + nt * 8 + tq * 2;
        *reinterpret_cast<float2*>(r0) = make_float2(o[nt][0], o[nt][1]);
        *reinterpret_cast<float2*>(r1) = make_float2(o[nt][2], o[nt][3]);
    }
    __syncthreads();
    if (tid < 16) {
        float ms = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; ++w) ms = fmaxf(ms, red_m[w][tid]);
        float den = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float wt = red_m[w][tid] == -INFINITY ? 0.f : expf(red_m[w][tid] - ms);
            mrg_w[w][tid] = wt;
            den = fmaf(wt, red_l[w][tid], den);
        }
        mrg_den[tid] = den;
        if (nsplit > 1 && tid < n_rep) {
            float* ml = a.part_ml + (((size_t)seq * a.nh + kvh * n_rep + tid) * nsplit + split) * 2;
            ml[0] = ms;
            ml[1] = den;
        }
    }
    __syncthreads();
    for (int i = tid; i < n_rep * (D / 2); i += kMmaAttnThreads) {
        const int h = i / (D / 2), c = (i % (D / 2)) * 2;
        float x0 = 0.f, x1 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float2 v = *reinterpret_cast<const float2*>(red_o + ((size_t)w * 16 + h) * D + c);
            x0 = fmaf(mrg_w[w][h], v.x, x0);
            x1 = fmaf(mrg_w[w][h], v.y, x1);
        }
        if (nsplit == 1) {
            const float den = mrg_den[h];
            store_out2(a, (size_t)seq * a.nh * D + (size_t)(kvh * n_rep + h) * D + c, x0 / den, x1 / den);
        } else {
            *reinterpret_cast<float2*>(a.part_acc + (((size_t)seq * a.nh + kvh * n_rep + h) * nsplit + split) * D + c) = make_float2(x0, x1);
        }
    }
    if (nsplit == 1) return;

    // ---- last CTA of this (sequence, kv head) merges the splits ----
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int ticket = atomicAdd(&a.counters[seq * a.nkv + kvh], 1);
        s_last = (ticket == nsplit - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float* cm = reinterpret_cast<float*>(kv);             // [n_rep][nsplit] split weights
    float* cden = cm + 16 * nsplit;
    for (int h = warp; h < n_rep; h += NW) {
        const size_t base = ((size_t)seq * a.nh + kvh * n_rep + h) * nsplit;
        float mstar = -INFINITY;
        for (int s = lane; s < nsplit; s += 32) mstar = fmaxf(mstar, __ldcg(a.part_ml + (base + s) * 2));
        mstar = warp_max(mstar);
        float den = 0.f;
        for (int s = lane; s < nsplit; s += 32) {
            const float ms = __ldcg(a.part_ml + (base + s) * 2);
            const float w = (ms == -INFINITY) ? 0.f : expf(ms - mstar);
            cm[h * nsplit + s] = w;
            den = fmaf(w, __ldcg(a.part_ml + (base + s) * 2 + 1), den);
        }
        den = warp_sum(den);
        if (lane == 0) cden[h] = den;
    }
    __syncthreads();
    for (int i = tid; i < n_rep * (D / 2); i += kMmaAttnThreads) {
        const int h = i / (D / 2), c = (i % (D / 2)) * 2;
        const float* src = a.part_acc + ((size_t)seq * a.nh + kvh * n_rep + h) * nsplit * D + c;
        float x0 = 0.f, x1 = 0.f;
        for (int s = 0; s < nsplit; ++s) {
            const float w = cm[h * nsplit + s];
            if (w != 0.f) {                                // empty splits hold unwritten partials
                const float2 v = __ldcg(reinterpret_cast<const float2*>(src + (size_t)s * D));
                x0 = fmaf(w, v.x, x0);
                x1 = fmaf(w, v.y, x1);
            }
        }
        const float den = cden[h];
        store_out2(a, (size_t)seq * a.nh * D + (size_t)(kvh * n_rep + h) * D + c, x0 / den, x1 / den);
    }
    if (tid == 0) a.counters[seq * a.nkv + kvh] = 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// prefill: grid (ceil(t / 64), nh, b), 128 threads; warp w owns query rows [16w, 16w+16) of the tile
// ---------------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kMmaAttnThreads) attn_prefill_kernel(const AttnArgs a) {
    constexpr int TILE = kKvPage * D;
    extern __shared__ __align__(128) uint8_t dsm[];
    uint16_t* qhi = reinterpret_cast<uint16_t*>(dsm);   // [64][D]
    uint16_t* qlo = qhi + kPrefillBM * D;
    uint16_t* kv = qlo + kPrefillBM * D;                // [2 stages][K tile | V tile]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    // heaviest (latest) query tiles first: the causal range grows with the tile index
    const int qt = gridDim.x - 1 - blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
    const int n_rep = a.nh / a.nkv, kvh = head / n_rep;
    pdl_launch_dependents();
    pdl_wait();
    const int i0 = qt * kPrefillBM;
    const int cslot = st_slot(a.state, seq);
    const int kv_base = a.state->kv_base[cslot];
    const int sw = a.sliding_window;
    const int* pt = a.page_table + (size_t)cslot * a.pt_stride;

    // key range of the whole tile
    const int ilast = min(i0 + kPrefillBM, a.t) - 1;
    const int len_tile = kv_base + ilast + 1;
    // the window thins out the NEW tokens only: candle concatenates zeros for the cached columns of its [t, t] mask, so every
    // cached key (< kv_base) stays visible and a call on a non-empty cache walks all pages
    const int start_tile = (sw > 0 && i0 - sw > 0 && kv_base == 0) ? i0 - sw : 0;
    const int p0 = start_tile / kKvPage, p1 = (len_tile + kKvPage - 1) / kKvPage;

    auto load = [&](int p, int st) {
        const size_t off = ((size_t)pt[p] * a.nkv + kvh) * (size_t)TILE;
        stage_kv_tile<D>(kv + st * 2 * TILE, kv + st * 2 * TILE + TILE, a.kpool + off, a.vpool + off, tid);
    };
    if (p0 < p1) load(p0, 0);
    for (int i = tid; i < kPrefillBM * (D / 4); i += kMmaAttnThreads) {
        const int r = i / (D / 4), c = (i % (D / 4)) * 4;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i0 + r < a.t) q = *reinterpret_cast<const float4*>(a.q + (((size_t)seq * a.t + i0 + r) * a.nh + head) * D + c);
        const int off = swz<D>(r, c);
        store_hi_lo(qhi, qlo, off, q.x); store_hi_lo(qhi, qlo, off + 1, q.y);
        store_hi_lo(qhi, qlo, off + 2, q.z); store_hi_lo(qhi, qlo, off + 3, q.w);
    }

    // per-thread rows: g and g + 8 of the warp's 16 (rows past t are clamped: computed, never stored)
    const int ir0 = min(i0 + warp * 16 + g, a.t - 1), ir1 = min(i0 + warp * 16 + g + 8, a.t - 1);
    const int len0 = kv_base + ir0 + 1, len1 = kv_base + ir1 + 1;
    const int st0 = (sw > 0 && ir0 - sw > 0) ? kv_base + ir0 - sw : 0;
    const int st1 = (sw > 0 && ir1 - sw > 0) ? kv_base + ir1 - sw : 0;
    const int warp_len = kv_base + min(i0 + warp * 16 + 15, a.t - 1) + 1;    // keys >= this are masked for the whole warp

    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int p = p0; p < p1; ++p) {
        const int st = (p - p0) & 1;
        if (p + 1 < p1) {
            load(p + 1, st ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint16_t* kt = kv + st * 2 * TILE;
        const uint16_t* vt = kt + TILE;
        const int key0 = p * kKvPage;
        if (key0 < warp_len) {
            float s[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
                uint32_t ah[4], al[4];
                ldmatrix_x4(ah, qhi + swz<D>(warp * 16 + (lane & 15), ks * 16 + (lane >> 4) * 8));
                ldmatrix_x4(al, qlo + swz<D>(warp * 16 + (lane & 15), ks * 16 + (lane >> 4) * 8));
#pragma unroll
                for (int n2 = 0; n2 < 4; ++n2) {
                    uint32_t kb[4];
                    ldmatrix_x4(kb, kt + swz<D>(n2 * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 16 + ((lane >> 3) & 1) * 8));
                    mma_bf16_16816(s[2 * n2], ah, kb[0], kb[1]);
                    mma_bf16_16816(s[2 * n2], al, kb[0], kb[1]);
                    mma_bf16_16816(s[2 * n2 + 1], ah, kb[2], kb[3]);
                    mma_bf16_16816(s[2 * n2 + 1], al, kb[2], kb[3]);
                }
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int key = key0 + nt * 8 + tq * 2 + (e & 1);
                    if (e < 2) {
                        s[nt][e] = ((key >= st0 || key < kv_base) && key < len0) ? s[nt][e] * a.qscale : -INFINITY;
                        mx0 = fmaxf(mx0, s[nt][e]);
                    } else {
                        s[nt][e] = ((key >= st1 || key < kv_base) && key < len1) ? s[nt][e] * a.qscale : -INFINITY;
                        mx1 = fmaxf(mx1, s[nt][e]);
                    }
                }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 2));
            const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
            // a row may not have seen a visible key yet (window start inside a later page): keep exp() finite
            const float mu0 = mn0 == -INFINITY ? 0.f : mn0, mu1 = mn1 == -INFINITY ? 0.f : mn1;
            const float al0 = expf(m0 - mu0), al1 = expf(m1 - mu1);
            m0 = mn0; m1 = mn1;
            float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = expf(s[nt][0] - mu0); s[nt][1] = expf(s[nt][1] - mu0);
                s[nt][2] = expf(s[nt][2] - mu1); s[nt][3] = expf(s[nt][3] - mu1);
                ps0 += s[nt][0] + s[nt][1];
                ps1 += s[nt][2] + s[nt][3];
            }
            l0 = l0 * al0 + ps0;
            l1 = l1 * al1 + ps1;
#pragma unroll
            for (int nt = 0; nt < D / 8; ++nt) {
                o[nt][0] *= al0; o[nt][1] *= al0; o[nt][2] *= al1; o[nt][3] *= al1;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t ph[4], pl[4];
                ph[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]); ph[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
                ph[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]); ph[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
                pl[0] = pack_bf16x2(s[2 * kk][0] - bf16lo(ph[0]), s[2 * kk][1] - bf16hi(ph[0]));
                pl[1] = pack_bf16x2(s[2 * kk][2] - bf16lo(ph[1]), s[2 * kk][3] - bf16hi(ph[1]));
                pl[2] = pack_bf16x2(s[2 * kk + 1][0] - bf16lo(ph[2]), s[2 * kk + 1][1] - bf16hi(ph[2]));
                pl[3] = pack_bf16x2(s[2 * kk + 1][2] - bf16lo(ph[3]), s[2 * kk + 1][3] - bf16hi(ph[3]));
#pragma unroll
                for (int dt = 0; dt < D / 16; ++dt) {
                    uint32_t vb[4];
                    ldmatrix_x4_trans(vb, vt + swz<D>(kk * 16 + (lane & 15), dt * 16 + (lane >> 4) * 8));
                    mma_bf16_16816(o[2 * dt], ph, vb[0], vb[1]);
                    mma_bf16_16816(o[2 * dt], pl, vb[0], vb[1]);
                    mma_bf16_16816(o[2 * dt + 1], ph, vb[2], vb[3]);
                    mma_bf16_16816(o[2 * dt + 1], pl, vb[2], vb[3]);
                }
            }
        }
        __syncthreads();
    }

    l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 1); l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 2);
    l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 1); l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    const int r0 = i0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        const int c = nt * 8 + tq * 2;
        if (r0 < a.t) store_out2(a, (((size_t)seq * a.t + r0) * a.nh + head) * D + c, o[nt][0] * inv0, o[nt][1] * inv0);
        if (r1 < a.t) store_out2(a, (((size_t)seq * a.t + r1) * a.nh + head) * D + c, o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

}  // namespace fl
