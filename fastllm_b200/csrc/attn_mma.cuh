// Tensor-core attention over the paged bf16 KV cache for the dense path (3+ activation rows), sm_100a.
//
//   attn_sk_decode_kernel<D>   : batched decode (one new token per sequence), stream-K: a fixed grid of resident CTAs cuts the
//       step's whole K|V page stream into equal ranges; a producer warp streams the range through a 2-D TMA ring, 4 consumer
//       warps fold 16 keys of every page each into an online softmax (mma.sync m16n8k16, the GQA group's heads are the M rows),
//       pairs that straddle ranges are merged by the last-arriving CTA.  HBM-bound.  Details at the kernel.
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
#include "gemm_tc.cuh"
#include "mma.cuh"

namespace fl {

constexpr int kMmaAttnThreads = 128;
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
// batched decode, stream-K: a fixed grid of resident CTAs splits the batch's whole K|V page stream evenly
// ---------------------------------------------------------------------------------------------------------------------
// attn_gqa_decode_kernel above hands each (split, kv head, sequence) to a CTA: the number of CTAs is a multiple of the pairs, not of
// the machine (batch 64 x 8 kv heads x 3 splits = 1536 CTAs on 444 slots = 3.46 waves; Mixtral's batch 32 = 256 CTAs on 444
// slots), every CTA re-pays set-up, q conversion and merge, and its cp.async ring drains at every CTA boundary.  Here the pages
// of all (sequence, kv head) pairs form ONE list (pair-major) cut into gridDim.x equal ranges:
//   * the producer warp streams the range's pages through a ring of K|V stages with 2-D TMA (one 64-row x 128-byte box per
//     half tile, 128-byte swizzle = the conflict-free ldmatrix layout; SASS UTMALDG) and never stops at a pair boundary;
//   * 4 consumer warps (16 keys of every page each, mma.sync m16n8k16, heads of the GQA group = the M rows) wait on the stage's
//     mbarrier, fold the page into their online softmax, release the stage -- no CTA-wide barrier per page;
//   * a pair that lies inside one range is finished by that CTA alone; a pair that straddles ranges leaves one partial per CTA
//     (at most two per CTA: its first and its last pair) and the last-arriving CTA of the pair merges them in CTA order, so
//     the result is deterministic for a given batch composition.
// Ragged batches (continuous batching) balance the same way: the split is by pages, not by sequences.
constexpr int kSkConsumerWarps = 4;
constexpr int kSkConsumers = kSkConsumerWarps * 32;
constexpr int kSkThreads = kSkConsumers + 32;
constexpr int kSkMaxStages = 4;

// element offset of (row, col) inside a staged 64 x D K or V page.  D >= 64: [D / 64 halves][64 rows][128 bytes] with the TMA
// engine's 128-byte swizzle (16-byte chunk index XOR row & 7); D < 64 (test-size heads): the plain [64][D] box, no swizzle.
template <int D>
__device__ __forceinline__ int kv_off(int row, int col) {
    if constexpr (D >= 64) return (col >> 6) * (kKvPage * 64) + row * 64 + (((((col >> 3) & 7) ^ (row & 7))) << 3) + (col & 7);
    else return row * D + col;
}

struct SkArgs {
    AttnArgs a;
    float* sk_acc;        // [gridDim.x][2][n_rep][D] partial numerators of the pairs a CTA shares with its neighbours
    float* sk_ml;         // [gridDim.x][2][n_rep][2] their running max / denominator
    int b;                // sequences in the step
    int nstages;          // K|V stages of the ring
    int layer_row0;       // row of this layer's page 0 / kv head 0 in the pool-wide tensor maps
};

template <int D>
__global__ void __launch_bounds__(kSkThreads) attn_sk_decode_kernel(const __grid_constant__ CUtensorMap tmk, const __grid_constant__ CUtensorMap tmv,
                                                                    const SkArgs k) {
    const AttnArgs& a = k.a;
    constexpr int TILE = kKvPage * D;
    constexpr int NW = kSkConsumerWarps;
    extern __shared__ __align__(1024) uint8_t sksm[];
    const int n_rep = a.nh / a.nkv;
    // q tiles (hi | lo, [16][D] each) and the cross-warp merge scratch ([NW][n_rep][D] f32) share the front region: the scratch is
    // written when a pair's last page is done, the q tiles are rewritten before the next pair starts
    const int front = max(2 * 16 * D * 2, NW * n_rep * D * 4);
    uint16_t* qhi = reinterpret_cast<uint16_t*>(sksm);
    uint16_t* qlo = qhi + 16 * D;
    float* red_o = reinterpret_cast<float*>(sksm);
    uint16_t* ring = reinterpret_cast<uint16_t*>(sksm + ((front + 1023) & ~1023));
    __shared__ __align__(8) uint64_t full[kSkMaxStages];
    __shared__ __align__(8) uint64_t empty[kSkMaxStages];
    __shared__ float red_m[NW][16], red_l[NW][16], mrg_w[NW][16], mrg_den[16], mrg_star[16];
    __shared__ int prefix[kMaxBatch + 1];        // pages of the sequences before s
    __shared__ int s_len[kMaxBatch], s_slot[kMaxBatch];
    __shared__ int s_last, s_i0, s_i1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int NST = k.nstages;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW);
        }
        mbar_fence_init();
    }
    pdl_launch_dependents();
    // No griddepcontrol.wait yet: the step state, the page table and every K/V row except the one the preceding kernel
    // (dense_qkv_epi_kernel) appends were written by kernels that completed before this one could start (that kernel releases its
    // dependents only after its own wait).  The producer streams all pages that cannot hold the new token right away; q, the new
    // K/V row and the output buffers are touched only after the wait.
    for (int i = tid; i < k.b; i += kSkThreads) {
        const int slot = st_slot(a.state, i);
        const int len = a.state->kv_base[slot] + 1;
        s_slot[i] = slot;
        s_len[i] = len;
        prefix[i + 1] = (len + kKvPage - 1) / kKvPage;
    }
    if (tid == 0) prefix[0] = 0;
    __syncthreads();
    if (warp == 0) {      // inclusive scan of the page counts
        int carry = 0;
        for (int base = 0; base < k.b; base += 32) {
            int v = base + lane < k.b ? prefix[base + lane + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, v, o);
                if (lane >= o) v += t;
            }
            if (base + lane < k.b) prefix[base + lane + 1] = v + carry;
            carry += __shfl_sync(0xFFFFFFFFu, v, 31);
        }
    }
    __syncthreads();
    const long long P = (long long)a.nkv * prefix[k.b];          // pages of the whole step (K and V of a page travel together)
    const int G = (int)min((long long)gridDim.x, P);              // every participating CTA owns at least one page
    const int cta = blockIdx.x;
    if (cta >= G) return;
    const long long flat0 = (long long)cta * P / G, flat1 = (long long)(cta + 1) * P / G;
    // flat index -> (sequence, kv head, page): pair-major, a sequence's nkv pairs are adjacent
    auto locate = [&](long long n, int& s, int& h, int& p) {
        int lo = 0, hi = k.b - 1;
        while (lo < hi) {      // largest s with nkv * prefix[s] <= n
            const int mid = (lo + hi + 1) >> 1;
            if ((long long)a.nkv * prefix[mid] <= n) lo = mid; else hi = mid - 1;
        }
        s = lo;
        const int np = prefix[s + 1] - prefix[s];
        const int rem = (int)(n - (long long)a.nkv * prefix[s]);
        h = rem / np;
        p = rem - h * np;
    };

    // =================================================================================================================
    // producer warp: page-table entries are fetched 32 pages at a time (one per lane), lane 0 issues the TMA requests
    // =================================================================================================================
    if (warp == NW) {
        unsigned int c = 0;
        bool waited = false;
        for (long long base = flat0; base < flat1; base += 32) {
            const long long n = base + lane;
            int row = 0;
            if (n < flat1) {
                int s, h, p;
                locate(n, s, h, p);
                const int pg = a.page_table[(size_t)s_slot[s] * a.pt_stride + p];
                row = k.layer_row0 + (pg * a.nkv + h) * kKvPage;
                if (p == prefix[s + 1] - prefix[s] - 1) row |= 1 << 31;      // the sequence's last page: holds the token appended in this step
            }
            const int cnt = (int)min(32LL, flat1 - base);
            for (int j = 0; j < cnt; ++j, ++c) {
                int r = __shfl_sync(0xFFFFFFFFu, row, j);
                if (r < 0 && !waited) {
                    pdl_wait();
                    waited = true;
                }
                r &= 0x7FFFFFFF;
                if (lane == 0) {
                    const int st = c % NST;
                    mbar_wait(&empty[st], ((c / NST) & 1) ^ 1);
                    mbar_expect_tx(&full[st], 2u * TILE * 2u);
                    uint16_t* kd = ring + (size_t)st * 2 * TILE;
                    if constexpr (D >= 64) {
#pragma unroll
                        for (int hf = 0; hf < D / 64; ++hf) {
                            tma_load_2d(kd + hf * (kKvPage * 64), &tmk, hf * 64, r, &full[st]);
                            tma_load_2d(kd + TILE + hf * (kKvPage * 64), &tmv, hf * 64, r, &full[st]);
                        }
                    } else {
                        tma_load_2d(kd, &tmk, 0, r, &full[st]);
                        tma_load_2d(kd + TILE, &tmv, 0, r, &full[st]);
                    }
                }
            }
        }
        return;
    }

    // =================================================================================================================
    // consumers
    // =================================================================================================================
    auto csync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kSkConsumers) : "memory"); };
    pdl_wait();
    int s, h, p;
    locate(flat0, s, h, p);
    float o[D / 8][4];
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    bool open = false, from_start = false;
    unsigned int c = 0;
    for (long long n = flat0; n < flat1; ++n, ++c) {
        const int np = prefix[s + 1] - prefix[s], len = s_len[s];
        if (!open) {       // first page of this CTA's share of pair (s, h): stage q, reset the running softmax
            csync();       // the previous pair's readers of the merge scratch are done
            for (int i = tid; i < 16 * D; i += kSkConsumers) {
                const int hh = i / D, dd = i % D;
                const float q = hh < n_rep ? a.q[((size_t)s * a.nh + h * n_rep + hh) * D + dd] : 0.f;
                store_hi_lo(qhi, qlo, swz<D>(hh, dd), q);
            }
            csync();
#pragma unroll
            for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
            m0 = m1 = -INFINITY;
            l0 = l1 = 0.f;
            open = true;
            from_start = (p == 0);
        }
        const int st = c % NST;
        mbar_wait(&full[st], (c / NST) & 1);
        const uint16_t* kt = ring + (size_t)st * 2 * TILE;
        const uint16_t* vt = kt + TILE;
        const int key0 = p * kKvPage + warp * 16;
        if (key0 < len) {
            float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
                uint32_t ah[4], al[4], kb[4];
                ldmatrix_x4(ah, qhi + swz<D>(lane & 15, ks * 16 + (lane >> 4) * 8));
                ldmatrix_x4(al, qlo + swz<D>(lane & 15, ks * 16 + (lane >> 4) * 8));
                ldmatrix_x4(kb, kt + kv_off<D>(warp * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 16 + ((lane >> 3) & 1) * 8));
                mma_bf16_16816(sc[0], ah, kb[0], kb[1]);
                mma_bf16_16816(sc[0], al, kb[0], kb[1]);
                mma_bf16_16816(sc[1], ah, kb[2], kb[3]);
                mma_bf16_16816(sc[1], al, kb[2], kb[3]);
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int key = key0 + nt * 8 + tq * 2 + (e & 1);
                    sc[nt][e] = key < len ? sc[nt][e] * a.qscale : -INFINITY;
                    if (e < 2) mx0 = fmaxf(mx0, sc[nt][e]); else mx1 = fmaxf(mx1, sc[nt][e]);
                }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 2));
            const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);       // finite: key0 < len is valid for every row
            const float al0 = expf(m0 - mn0), al1 = expf(m1 - mn1);
            m0 = mn0; m1 = mn1;
            float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                sc[nt][0] = expf(sc[nt][0] - mn0); sc[nt][1] = expf(sc[nt][1] - mn0);
                sc[nt][2] = expf(sc[nt][2] - mn1); sc[nt][3] = expf(sc[nt][3] - mn1);
                ps0 += sc[nt][0] + sc[nt][1];
                ps1 += sc[nt][2] + sc[nt][3];
            }
            l0 = l0 * al0 + ps0;       // per-thread partial row sums; the quad is reduced when the pair is emitted
            l1 = l1 * al1 + ps1;
            uint32_t ph[4], pl[4];
            ph[0] = pack_bf16x2(sc[0][0], sc[0][1]); ph[1] = pack_bf16x2(sc[0][2], sc[0][3]);
            ph[2] = pack_bf16x2(sc[1][0], sc[1][1]); ph[3] = pack_bf16x2(sc[1][2], sc[1][3]);
            pl[0] = pack_bf16x2(sc[0][0] - bf16lo(ph[0]), sc[0][1] - bf16hi(ph[0]));
            pl[1] = pack_bf16x2(sc[0][2] - bf16lo(ph[1]), sc[0][3] - bf16hi(ph[1]));
            pl[2] = pack_bf16x2(sc[1][0] - bf16lo(ph[2]), sc[1][1] - bf16hi(ph[2]));
            pl[3] = pack_bf16x2(sc[1][2] - bf16lo(ph[3]), sc[1][3] - bf16hi(ph[3]));
#pragma unroll
            for (int dt = 0; dt < D / 16; ++dt) {
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, vt + kv_off<D>(warp * 16 + (lane & 15), dt * 16 + (lane >> 4) * 8));
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
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);

        const bool pair_done = (p == np - 1);
        if (pair_done || n == flat1 - 1) {
            // ---- emit: merge the 4 warps (each saw a different quarter of every page) ----
            const bool whole = from_start && pair_done;       // nobody else holds a piece of this pair
            l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 1); l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 2);
            l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 1); l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 2);
            csync();       // every warp is past its last read of the q tiles (the scratch overlays them)
            if (tq == 0) {
                red_m[warp][g] = m0; red_m[warp][g + 8] = m1;
                red_l[warp][g] = l0; red_l[warp][g + 8] = l1;
            }
#pragma unroll
            for (int nt = 0; nt < D / 8; ++nt) {
                if (g < n_rep) *reinterpret_cast<float2*>(red_o + ((size_t)warp * n_rep + g) * D + nt * 8 + tq * 2) = make_float2(o[nt][0], o[nt][1]);
                if (g + 8 < n_rep) *reinterpret_cast<float2*>(red_o + ((size_t)warp * n_rep + g + 8) * D + nt * 8 + tq * 2) = make_float2(o[nt][2], o[nt][3]);
            }
            csync();
            const int part = cta * 2 + (from_start ? 1 : 0);      // this CTA's partial slot for the pair
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
                if (!whole && tid < n_rep) {
                    float* ml = k.sk_ml + ((size_t)part * n_rep + tid) * 2;
                    ml[0] = ms;
                    ml[1] = den;
                }
            }
            csync();
            for (int i = tid; i < n_rep * (D / 2); i += kSkConsumers) {
                const int hh = i / (D / 2), cc = (i % (D / 2)) * 2;
                float x0 = 0.f, x1 = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const float2 v = *reinterpret_cast<const float2*>(red_o + ((size_t)w * n_rep + hh) * D + cc);
                    x0 = fmaf(mrg_w[w][hh], v.x, x0);
                    x1 = fmaf(mrg_w[w][hh], v.y, x1);
                }
                if (whole) {
                    const float den = mrg_den[hh];
                    store_out2(a, (size_t)s * a.nh * D + (size_t)(h * n_rep + hh) * D + cc, x0 / den, x1 / den);
                } else {
                    *reinterpret_cast<float2*>(k.sk_acc + ((size_t)part * n_rep + hh) * D + cc) = make_float2(x0, x1);
                }
            }
            if (!whole) {
                // ---- the pair straddles ranges: the last CTA to arrive merges the partials in CTA order ----
                __threadfence();
                csync();
                if (tid == 0) {
                    const long long ps = (long long)a.nkv * prefix[s] + (long long)h * np, pe = ps + np;
                    const int i0 = (int)(((ps + 1) * G - 1) / P), i1 = (int)((pe * G - 1) / P);      // CTAs holding the first / last page
                    const int ticket = atomicAdd(&a.counters[s * a.nkv + h], 1);
                    s_last = (ticket == i1 - i0);
                    s_i0 = i0;
                    s_i1 = i1;
                }
                csync();
                if (s_last) {
                    __threadfence();
                    const int i0 = s_i0, i1 = s_i1;
                    // partial of CTA j: slot 1 in the CTA where the pair starts (i0), slot 0 in the others
                    for (int hh = warp; hh < n_rep; hh += NW) {
                        float mstar = -INFINITY;
                        for (int j = i0 + lane; j <= i1; j += 32)
                            mstar = fmaxf(mstar, __ldcg(k.sk_ml + ((size_t)(j * 2 + (j == i0)) * n_rep + hh) * 2));
                        mstar = warp_max(mstar);
                        float den = 0.f;
                        for (int j = i0 + lane; j <= i1; j += 32) {
                            const float* ml = k.sk_ml + ((size_t)(j * 2 + (j == i0)) * n_rep + hh) * 2;
                            den = fmaf(expf(__ldcg(ml) - mstar), __ldcg(ml + 1), den);
                        }
                        den = warp_sum(den);
                        if (lane == 0) { mrg_star[hh] = mstar; mrg_den[hh] = den; }
                    }
                    csync();
                    for (int i = tid; i < n_rep * (D / 2); i += kSkConsumers) {
                        const int hh = i / (D / 2), cc = (i % (D / 2)) * 2;
                        float x0 = 0.f, x1 = 0.f;
                        for (int j = i0; j <= i1; ++j) {
                            const size_t pj = (size_t)(j * 2 + (j == i0)) * n_rep + hh;
                            const float w = expf(__ldcg(k.sk_ml + pj * 2) - mrg_star[hh]);
                            const float2 v = __ldcg(reinterpret_cast<const float2*>(k.sk_acc + pj * D + cc));
                            x0 = fmaf(w, v.x, x0);
                            x1 = fmaf(w, v.y, x1);
                        }
                        const float den = mrg_den[hh];
                        store_out2(a, (size_t)s * a.nh * D + (size_t)(h * n_rep + hh) * D + cc, x0 / den, x1 / den);
                    }
                    if (tid == 0) a.counters[s * a.nkv + h] = 0;
                }
            }
            open = false;
        }
        if (++p == np) {
            p = 0;
            if (++h == a.nkv) { h = 0; ++s; }
        }
    }
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
