// Persistent batch-1 decode kernel: ONE cooperative launch runs whole decode steps (all layers + lm_head + arg-max),
// sm_100a.  This is the B200-first answer to "every small kernel pays a ramp and a tail": at batch 1 a decoder layer of
// Mistral-7B is five kernels of 5-36 us of HBM time each, and launch/ramp/tail costs as much as the streaming.
//
//   * grid = one CTA per SM (148), all resident (cooperative launch), 8 consumer warps + 1 producer warp.
//   * WEIGHT STREAM: the producer thread walks the step's static schedule -- for every layer, for every GEMV phase, this
//     CTA's contiguous row slice of the weight matrix, cut into <=32 KB row-aligned chunks -- and keeps a ring of
//     shared-memory stages full with cp.async.bulk (TMA engine) + mbarrier.  The stream is decoupled from the compute
//     phases: while consumers sit in a grid barrier, an epilogue or the attention phase, the ring is already filling with
//     the NEXT phase's weights, so HBM keeps streaming across phase boundaries.
//   * consumers: per chunk, all 8 warps sweep each staged row (LDS.128 weights, f32 activations from shared memory,
//     f32 FMA), warp-shuffle reduce, per-warp partials in shared memory, one cross-warp sum per phase followed by the
//     same fused epilogues as gemv.cuh (RMSNorm prologue; bias+RoPE+KV append; residual add; SiLU*up; logits+arg-max).
//   * phases of a layer are separated by a grid barrier (monotonic atomic counter, release/acquire fences); data that
//     crosses CTAs is read with ld.global.cg (L2) because L1 is not coherent across SMs.
//   * attention: split-K over the sequence, (kv_head, split) items spread over the CTAs, K/V read straight from the paged
//     cache with 128-bit ld.global.cg, online softmax, last-arriver merge (same algorithm as attn_decode.cuh).
//   * `nsteps` decode steps can run inside one launch (greedy feedback on device), so the device-resident loop has no
//     launch gaps at all.
#pragma once
#include "attn_decode.cuh"
#include "common.cuh"
#include "gemv.cuh"

namespace fl {

constexpr int kPkConsumerWarps = 8;
constexpr int kPkConsumers = kPkConsumerWarps * 32;
constexpr int kPkThreads = kPkConsumers + 32;
constexpr int kPkStageBytes = 32768;
constexpr int kPkMaxStages = 6;

struct PkLayer {
    const uint16_t* wqkv;
    const float* bqkv;
    const uint16_t* wo;
    const uint16_t* wgu;
    const uint16_t* wdown;
    const float* ln1;
    const float* ln2;
};

struct PkArgs {
    const PkLayer* layers;
    int L;
    const uint16_t* embed;
    const uint16_t* lm_head;
    const float* final_norm;
    int H, I, V, nh, nkv, d, nqkv, max_pos;
    float eps, qscale;
    const float* rope_cos;
    const float* rope_sin;
    uint16_t* kpool;
    uint16_t* vpool;
    size_t layer_pool_elems;
    const int* page_table;
    StepState* state;
    float* resid;      // [H]
    float* q;          // [nh*d]
    float* attn_out;   // [nh*d]
    float* act;        // [I]
    float* logits;     // [V]
    float* part_acc;   // [nh, nsplit, d]
    float* part_ml;    // [nh, nsplit, 2]
    int* counters;     // [nkv]
    int nsplit;
    float* amax_val;   // [gridDim.x]
    int* amax_idx;
    uint32_t* ids;
    uint32_t* next_ids;
    uint32_t* trace;
    int* trace_pos;
    unsigned int* gbar;   // grid-barrier counter, zero at launch
    int nsteps, feedback;
    int nstages;
    int xs_floats;        // shared-memory activation vector capacity (floats)
    int partial_rows;     // rows of the per-warp partial buffer
    // ---- tensor parallelism inside the kernel: all-reduce over NVLink peer memory (no NCCL call, no extra launch) ----
    int tp, rank;                    // tp == 1: single GPU
    float* peer_part[8];             // rank r's receive area [2 parities][tp sources][H] f32   (peer-mapped, r == rank: local)
    unsigned int* peer_flag[8];      // rank r's flags [tp sources][gridDim.x + 1] u32, monotone exchange epochs
    float* peer_amax[8];             // rank r's arg-max exchange slots [tp sources][2] (value, index bits)
    float* peer_logits[8];           // rank r's full-vocabulary logits [Vfull]
    unsigned int ar_epoch0;          // exchanges completed on this communicator before this launch
    int* comm_err;                   // set to 1 when a peer did not show up within the spin budget
    int flags;            // bit 0: prefetch this CTA's KV pages into L2 at the top of P1
    int lookahead_bytes;  // how far (per CTA) the producer prefetches into L2 beyond the shared-memory ring
    long long* dbg;       // optional: CTA 0 writes %globaltimer at the phase boundaries of layer L/2 (FL_PK_DEBUG=1)
};

// ---- chunk schedule shared by producer and consumers --------------------------------------------------------------
struct PkSlice {
    int row_begin, row_end;   // even-aligned row range of this CTA
    int rpc;                  // rows per chunk (row fits in a stage)
    int nseg;                 // >1: a row is cut into nseg segments of seg_pieces 16-byte pieces
    int seg_pieces;
    int K8;
    int nchunks;
};

__device__ __forceinline__ PkSlice pk_slice(int N, int K, int cta, int ncta) {
    PkSlice s;
    const int npairs = N >> 1;
    s.row_begin = 2 * (int)((long long)npairs * cta / ncta);
    s.row_end = 2 * (int)((long long)npairs * (cta + 1) / ncta);
    s.K8 = K >> 3;
    const int rowbytes = K * 2;
    const int rows = s.row_end - s.row_begin;
    if (rowbytes <= kPkStageBytes) {
        const int fit = kPkStageBytes / rowbytes;
        s.rpc = fit >= 8 ? 8 : (fit >= 4 ? 4 : (fit >= 2 ? 2 : 1));   // power of two: 8 / rpc warps share one row
        s.nseg = 1;
        s.seg_pieces = s.K8;
        s.nchunks = (rows + s.rpc - 1) / s.rpc;
    } else {
        s.rpc = 1;
        s.nseg = (rowbytes + kPkStageBytes - 1) / kPkStageBytes;
        s.seg_pieces = (s.K8 + s.nseg - 1) / s.nseg;
        s.nchunks = rows * s.nseg;
    }
    return s;
}

__device__ __forceinline__ void pk_named_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kPkConsumers) : "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ uint4 ldcg_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void prefetch_l2_bulk(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// L2 eviction-priority hints for the weight stream.  Weights are read exactly once per step, and the L2 lookahead cursor runs
// ahead of the demand loads: under the default (LRU-like) policy the OLDEST lines in L2 are the prefetched-but-not-yet-consumed
// ones, so a deep lookahead evicts exactly the lines it is about to need (measured: 512 KB/CTA lookahead = 1.5x the HBM traffic).
// Demand loads therefore carry evict_first (the line is dead once it is in shared memory) and prefetches evict_last.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk_hint(const void* gsrc, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(gsrc), "r"(bytes), "l"(policy) : "memory");
}

// Weight matrix of global phase g of a step: g = 4*layer + {0:qkv, 1:o, 2:gate|up, 3:down}, g = 4*L: lm_head.
__device__ __forceinline__ void pk_phase(const PkArgs& a, int g, const uint16_t*& W, int& N, int& K) {
    if (g == 4 * a.L) { W = a.lm_head; N = a.V; K = a.H; return; }
    const PkLayer& lw = a.layers[g >> 2];
    switch (g & 3) {
        case 0: W = lw.wqkv; N = a.nqkv; K = a.H; break;
        case 1: W = lw.wo; N = a.H; K = a.nh * a.d; break;
        case 2: W = lw.wgu; N = 2 * a.I; K = a.H; break;
        default: W = lw.wdown; N = a.H; K = a.I; break;
    }
}

// Walks this CTA's chunk schedule (every step, every phase, every chunk) in consumption order.
struct PkCursor {
    int step, g, i, K;
    const uint16_t* W;
    PkSlice s;
    bool valid;
    __device__ void enter(const PkArgs& a, int cta, int ncta) {
        while (true) {
            if (step >= a.nsteps) { valid = false; return; }
            int N;
            pk_phase(a, g, W, N, K);
            s = pk_slice(N, K, cta, ncta);
            i = 0;
            if (s.nchunks > 0) { valid = true; return; }
            if (++g > 4 * a.L) { g = 0; ++step; }
        }
    }
    __device__ void init(const PkArgs& a, int cta, int ncta) { step = 0; g = 0; enter(a, cta, ncta); }
    __device__ void get(const uint16_t*& src, uint32_t& bytes) const {
        if (s.nseg == 1) {
            const int r0 = s.row_begin + i * s.rpc;
            const int nr = min(s.rpc, s.row_end - r0);
            src = W + (size_t)r0 * K;
            bytes = (uint32_t)nr * K * 2;
        } else {
            const int r = s.row_begin + i / s.nseg, seg = i % s.nseg;
            const int p0 = seg * s.seg_pieces;
            const int np = min(s.seg_pieces, s.K8 - p0);
            src = W + (size_t)r * K + (size_t)p0 * 8;
            bytes = (uint32_t)np * 16;
        }
    }
    __device__ void advance(const PkArgs& a, int cta, int ncta) {
        if (++i < s.nchunks) return;
        if (++g > 4 * a.L) { g = 0; ++step; }
        enter(a, cta, ncta);
    }
};

// Grid barrier among the consumer threads of all CTAs (the producer warp never takes part and keeps streaming).
__device__ __forceinline__ void pk_grid_barrier(unsigned int* ctr, unsigned int& epoch, int tid) {
    pk_named_sync();
    epoch += 1;
    if (tid == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        const unsigned int target = epoch * gridDim.x;
        while (ld_acquire_gpu(ctr) < target) { }
        __threadfence();
    }
    pk_named_sync();
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// wait until the peer's flag reaches `epoch`; bounded (about 2 s) so that a missing peer becomes an error, not a hang
__device__ __forceinline__ void pk_wait_flag(const unsigned int* flag, unsigned int epoch, int* err) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        if (clock64() - t0 > 4000000000LL) {
            *err = 1;
            break;
        }
    }
}

template <int D>
__global__ void __launch_bounds__(kPkThreads, 1) decode_persistent_kernel(const PkArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // layout: [ring: nstages * 32 KB][xs: xs_floats f32][partial: partial_rows * 8 f32][sc: 8 * 64 f32][small]
    uint8_t* ring = smem;
    float* xs = reinterpret_cast<float*>(smem + (size_t)a.nstages * kPkStageBytes);
    float* partial = xs + a.xs_floats;
    float* sc = partial + (size_t)a.partial_rows * kPkConsumerWarps;      // [kAttnMaxRep][kKvPage]
    float* red = sc + kAttnMaxRep * kKvPage;                              // [32] block-reduce scratch
    float* alpha_s = red + 32;                                            // [8]
    float* mrun = alpha_s + 8;
    float* lrun = mrun + 8;
    int* s_flag = reinterpret_cast<int*>(lrun + 8);
    __shared__ __align__(8) uint64_t full[kPkMaxStages];
    __shared__ __align__(8) uint64_t empty[kPkMaxStages];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int NS = a.nstages;
    const int nq = a.nh * a.d;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kPkConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // =================================================================================================================
    // PRODUCER: one thread streams every weight chunk this CTA will ever need, in consumption order.
    // =================================================================================================================
    if (warp == kPkConsumerWarps) {
        if (lane != 0) return;
        // Two cursors over the same schedule: `ld` feeds the shared-memory ring, `pf` runs up to lookahead_bytes ahead of
        // it issuing L2 prefetches, so HBM keeps streaming into the 126 MB L2 while the ring is full and the consumers
        // are busy with an epilogue, a grid barrier or the attention phase; the ring then refills at L2 speed.
        unsigned int c = 0;   // running chunk counter -> stage = c % NS, use = c / NS
        PkCursor ld, pf;
        ld.init(a, cta, ncta);
        pf.init(a, cta, ncta);
        int ahead = 0;
        const uint64_t pol_first = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
        const bool hint_ld = (a.flags & 2) != 0, hint_pf = (a.flags & 4) != 0;
        while (ld.valid) {
            while (pf.valid && ahead < a.lookahead_bytes) {
                const uint16_t* src;
                uint32_t bytes;
                pf.get(src, bytes);
                if (hint_pf) prefetch_l2_bulk_hint(src, bytes, pol_last);
                else prefetch_l2_bulk(src, bytes);
                ahead += (int)bytes;
                pf.advance(a, cta, ncta);
            }
            const uint16_t* src;
            uint32_t bytes;
            ld.get(src, bytes);
            if (a.flags & 8) src = ld.W + (size_t)ld.s.row_begin * ld.K;   // DEV experiment: always this CTA's first chunk (L2-resident): consumer-bound rate
            const int st = c % NS;
            mbar_wait(&empty[st], ((c / NS) & 1) ^ 1);
            mbar_expect_tx(&full[st], bytes);
            if (hint_ld) bulk_g2s_hint(ring + (size_t)st * kPkStageBytes, src, bytes, &full[st], pol_first);
            else bulk_g2s(ring + (size_t)st * kPkStageBytes, src, bytes, &full[st]);
            ahead -= (int)bytes;
            ld.advance(a, cta, ncta);
            ++c;
        }
        return;
    }

    // =================================================================================================================
    // CONSUMERS
    // =================================================================================================================
    unsigned int c = 0;       // chunk counter, in lock-step with the producer's
    unsigned int epoch = 0;   // grid-barrier epoch

    // y = W_slice . xs for this CTA's rows.  Within a chunk every warp owns ONE contiguous piece range of ONE row
    // (8 / rpc warps share a row), so there is a single warp-shuffle reduction per warp per 32 KB chunk; the per-warp
    // partial sums land in partial[(row - row_begin) * 8 + slot] and are added up in the phase epilogue.
    auto dot_range = [&](const uint4* wrow, const float4* xlo, const float4* xhi, int lo, int hi) {
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        int p = lo + lane;
        for (; p + 96 < hi; p += 128) {
            const uint4 w0 = wrow[p], w1 = wrow[p + 32], w2 = wrow[p + 64], w3 = wrow[p + 96];
            const float4 a0 = xlo[p], b0 = xhi[p], a1 = xlo[p + 32], b1 = xhi[p + 32];
            const float4 a2 = xlo[p + 64], b2 = xhi[p + 64], a3 = xlo[p + 96], b3 = xhi[p + 96];
            const float x0[8] = {a0.x, a0.y, a0.z, a0.w, b0.x, b0.y, b0.z, b0.w};
            const float x1[8] = {a1.x, a1.y, a1.z, a1.w, b1.x, b1.y, b1.z, b1.w};
            const float x2[8] = {a2.x, a2.y, a2.z, a2.w, b2.x, b2.y, b2.z, b2.w};
            const float x3[8] = {a3.x, a3.y, a3.z, a3.w, b3.x, b3.y, b3.z, b3.w};
            acc0 = dot8(w0, x0, acc0);
            acc1 = dot8(w1, x1, acc1);
            acc2 = dot8(w2, x2, acc2);
            acc3 = dot8(w3, x3, acc3);
        }
        for (; p < hi; p += 32) {
            const uint4 w0 = wrow[p];
            const float4 a0 = xlo[p], b0 = xhi[p];
            const float x0[8] = {a0.x, a0.y, a0.z, a0.w, b0.x, b0.y, b0.z, b0.w};
            acc0 = dot8(w0, x0, acc0);
        }
        return warp_sum((acc0 + acc1) + (acc2 + acc3));
    };
    auto consume = [&](int N, int K) -> PkSlice {
        const PkSlice s = pk_slice(N, K, cta, ncta);
        const float4* xs4 = reinterpret_cast<const float4*>(xs);
        if (s.nseg == 1) {
            // Register-resident activations: within a phase a lane always multiplies the same <= 8 column pieces
            // (p = lo + lane + 32 j; rpc is chosen so that a warp's share of a row is <= 256 pieces), so its slice of x
            // is loaded from shared memory ONCE per phase; the streaming loop is then 1 LDS.128 (weights) + 8 unpack +
            // 8 FMA per 16 bytes -- a third of the shared-memory traffic of reading x per piece.
            constexpr int XR = 8;
            const int wpr = kPkConsumerWarps / s.rpc;            // warps per row
            const int my_row = warp / wpr, sub = warp % wpr;
            const int per = (s.K8 + wpr - 1) / wpr;
            const int lo = sub * per, hi = min(lo + per, s.K8);
            float xr[XR][8];
#pragma unroll
            for (int j = 0; j < XR; ++j) {
                const int p = lo + lane + 32 * j;
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), b0 = a0;
                if (p < hi) { a0 = xs4[p]; b0 = xs4[s.K8 + p]; }
                xr[j][0] = a0.x; xr[j][1] = a0.y; xr[j][2] = a0.z; xr[j][3] = a0.w;
                xr[j][4] = b0.x; xr[j][5] = b0.y; xr[j][6] = b0.z; xr[j][7] = b0.w;
            }
            for (int i = 0; i < s.nchunks; ++i, ++c) {
                const int st = c % NS;
                mbar_wait(&full[st], (c / NS) & 1);
                const int r0 = s.row_begin + i * s.rpc;
                const int nr = min(s.rpc, s.row_end - r0);
                if (my_row < nr) {
                    const uint4* wrow = reinterpret_cast<const uint4*>(ring + (size_t)st * kPkStageBytes + (size_t)my_row * K * 2);
                    uint4 wv[XR];
#pragma unroll
                    for (int j = 0; j < XR; ++j) {
                        const int p = lo + lane + 32 * j;
                        wv[j] = (p < hi) ? wrow[p] : make_uint4(0u, 0u, 0u, 0u);
                    }
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < XR; ++j) acc[j & 3] = dot8(wv[j], xr[j], acc[j & 3]);
                    const float v = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
                    if (lane == 0) partial[(size_t)(r0 + my_row - s.row_begin) * kPkConsumerWarps + sub] = v;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
            }
        } else {
            // rows longer than a stage (e.g. Qwen2.5 down_proj, K = 18944): each chunk is a row segment swept by all warps
            for (int i = 0; i < s.nchunks; ++i, ++c) {
                const int st = c % NS;
                mbar_wait(&full[st], (c / NS) & 1);
                const int r = s.row_begin + i / s.nseg, seg = i % s.nseg;
                const int p0 = seg * s.seg_pieces;
                const int np = min(s.seg_pieces, s.K8 - p0);
                const int sper = (np + kPkConsumerWarps - 1) / kPkConsumerWarps;
                const int slo = warp * sper, shi = min(slo + sper, np);
                const float v = dot_range(reinterpret_cast<const uint4*>(ring + (size_t)st * kPkStageBytes), xs4 + p0, xs4 + s.K8 + p0, slo, shi);
                if (lane == 0) {
                    float* dst = &partial[(size_t)(r - s.row_begin) * kPkConsumerWarps + warp];
                    *dst = (seg == 0) ? v : (*dst + v);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
            }
        }
        pk_named_sync();   // partial[] complete
        return s;
    };
    // sum of the per-warp partials of one row (slots actually written: 8 / rpc, or all 8 for segmented rows)
    auto row_sum = [&](const PkSlice& s, int local_row) {
        const float* pr = partial + (size_t)local_row * kPkConsumerWarps;
        const int nslot = (s.nseg == 1) ? kPkConsumerWarps / s.rpc : kPkConsumerWarps;
        float v = pr[0];
        for (int k = 1; k < nslot; ++k) v += pr[k];
        return v;
    };
    // Row-parallel epilogue under tensor parallelism: resid[slice] += sum over ranks of this rank-local partial output.
    // One-shot exchange over NVLink peer memory: every CTA PUSHES its partial row slice into every rank's receive area
    // (tp x ~100 bytes), publishes a per-CTA flag with release.sys, waits for the same CTA of every peer, then sums the tp
    // partials in rank order (identical on every rank, so the replicas stay bit-identical) and adds them to the residual.
    unsigned int ar_epoch = a.ar_epoch0;
    auto allreduce_resid_add = [&](const PkSlice& s) {
        ar_epoch += 1;
        const int par = ar_epoch & 1;
        const int npair = (s.row_end - s.row_begin) / 2;
        for (int e = tid; e < npair; e += kPkConsumers) {
            const float2 v = make_float2(row_sum(s, 2 * e), row_sum(s, 2 * e + 1));
            const size_t off = (size_t)(par * a.tp + a.rank) * a.H + s.row_begin + 2 * e;
            for (int r = 0; r < a.tp; ++r) *reinterpret_cast<float2*>(a.peer_part[r] + off) = v;
        }
        // no per-thread system fence: the CTA barrier orders every thread's pushes before the flag writers' st.release.sys,
        // which is cumulative at system scope (PTX memory model), so one release per peer publishes the whole slice
        pk_named_sync();
        if (tid < a.tp) {
            st_release_sys(a.peer_flag[tid] + (size_t)a.rank * (ncta + 1) + cta, ar_epoch);
            pk_wait_flag(a.peer_flag[a.rank] + (size_t)tid * (ncta + 1) + cta, ar_epoch, a.comm_err);
        }
        pk_named_sync();
        const float* mine = a.peer_part[a.rank];
        for (int e = tid; e < npair; e += kPkConsumers) {
            float2* p = reinterpret_cast<float2*>(a.resid + s.row_begin + 2 * e);
            float2 v = __ldcg(p);
            for (int r = 0; r < a.tp; ++r) {
                const float2 y = __ldcg(reinterpret_cast<const float2*>(mine + (size_t)(par * a.tp + r) * a.H + s.row_begin + 2 * e));
                v.x += y.x;
                v.y += y.y;
            }
            *p = v;
        }
    };
    auto block_sum = [&](float v) {   // sum over the 256 consumer threads
        v = warp_sum(v);
        if (lane == 0) red[warp] = v;
        pk_named_sync();
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kPkConsumerWarps; ++w) s += red[w];
        pk_named_sync();
        return s;
    };
    // xs = rms_norm(src) * norm_w      (candle_nn::ops::rms_norm, as in gemv.cuh)
    // xs = rms_norm(src) * norm_w      (candle_nn::ops::rms_norm, as in gemv.cuh)
    auto load_x_rmsnorm = [&](const float* src, bool src_is_embed_row, const uint16_t* erow, const float* norm_w, int K,
                              float* resid_out) {
        float ss = 0.f;
        for (int i = tid; i < K / 4; i += kPkConsumers) {
            float4 v;
            if (src_is_embed_row) {
                const uint2 e = __ldg(reinterpret_cast<const uint2*>(erow) + i);
                v = make_float4(bf16lo(e.x), bf16hi(e.x), bf16lo(e.y), bf16hi(e.y));
                if (resid_out) reinterpret_cast<float4*>(resid_out)[i] = v;
            } else {
                v = __ldcg(reinterpret_cast<const float4*>(src) + i);
            }
            reinterpret_cast<float4*>(xs)[(i & 1) * (K >> 3) + (i >> 1)] = v;   // [half][chunk][4]: conflict-free LDS.128
            ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        }
        const float tot = block_sum(ss);
        const float m = sqrtf(tot / (float)K + a.eps);
        for (int i = tid; i < K / 4; i += kPkConsumers) {
            float4 v = reinterpret_cast<float4*>(xs)[(i & 1) * (K >> 3) + (i >> 1)];
            const float4 w = __ldg(reinterpret_cast<const float4*>(norm_w) + i);
            v.x = v.x / m * w.x; v.y = v.y / m * w.y; v.z = v.z / m * w.z; v.w = v.w / m * w.w;
            reinterpret_cast<float4*>(xs)[(i & 1) * (K >> 3) + (i >> 1)] = v;
        }
        pk_named_sync();
    };
    auto load_x_plain = [&](const float* src, int K) {
        for (int i = tid; i < K / 4; i += kPkConsumers)
            reinterpret_cast<float4*>(xs)[(i & 1) * (K >> 3) + (i >> 1)] = __ldcg(reinterpret_cast<const float4*>(src) + i);
        pk_named_sync();
    };

    int dbg_i = 0;
    auto stamp = [&](int l) {
        if (a.dbg != nullptr && cta == 0 && tid == 0 && l == a.L / 2) {
            long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.dbg[dbg_i] = t;
        }
        ++dbg_i;
    };
    const int n_rep = a.nh / a.nkv;
    constexpr int LPT = D / 8, TPW = 32 / LPT;
    const int grp = lane / LPT, gl = lane % LPT;

    for (int step = 0; step < a.nsteps; ++step) {
        const uint32_t tok_id = __ldcg(a.ids);
        const int rope_pos = min(__ldcg(&a.state->rope_pos), a.max_pos - 1);
        const int kv_base = __ldcg(&a.state->kv_base[0]);
        const int slot = kv_base;                       // cache slot of the token being decoded
        const int len = kv_base + 1;
        const int page_of_slot = __ldg(a.page_table + slot / kKvPage);

        for (int l = 0; l < a.L; ++l) {
            const PkLayer& lw = a.layers[l];
            uint16_t* kpool = a.kpool + (size_t)l * a.layer_pool_elems;
            uint16_t* vpool = a.vpool + (size_t)l * a.layer_pool_elems;

            // ---------------- P1: RMSNorm -> q|k|v (+bias) -> RoPE -> q store + KV append ----------------
            dbg_i = 0;
            stamp(l);
            if (a.flags & 1) {   // pull the K/V pages of this CTA's attention item(s) of this layer towards L2 while P1 runs
                const int npages = (len + kKvPage - 1) / kKvPage;
                const int per = (npages + a.nsplit - 1) / a.nsplit;
                for (int item = cta; item < a.nkv * a.nsplit; item += ncta) {
                    const int kvh = item / a.nsplit, split = item % a.nsplit;
                    const int p0 = split * per, p1 = min(p0 + per, npages);
                    constexpr int kLines = kKvPage * D * 2 / 128;          // 128-byte lines per K (or V) page chunk
                    for (int i = tid; i < (p1 - p0) * 2 * kLines; i += kPkConsumers) {
                        const int pg = p0 + i / (2 * kLines), rem = i % (2 * kLines);
                        const size_t base = ((size_t)__ldg(a.page_table + pg) * a.nkv + kvh) * (size_t)(kKvPage * D);
                        const uint16_t* src = (rem < kLines ? kpool : vpool) + base + (size_t)(rem % kLines) * 64;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
                    }
                }
            }
            if (l == 0)
                load_x_rmsnorm(nullptr, true, a.embed + (size_t)tok_id * a.H, lw.ln1, a.H, cta == 0 ? a.resid : nullptr);
            else
                load_x_rmsnorm(a.resid, false, nullptr, lw.ln1, a.H, nullptr);
            stamp(l);
            {
                const PkSlice s = consume(a.nqkv, a.H);
                stamp(l);
                const int d = a.d, half = d >> 1;
                for (int e = tid; e < (s.row_end - s.row_begin) / 2; e += kPkConsumers) {
                    const int ra = s.row_begin + 2 * e;
                    float va = row_sum(s, 2 * e), vb = row_sum(s, 2 * e + 1);
                    if (lw.bqkv) { va += lw.bqkv[ra]; vb += lw.bqkv[ra + 1]; }
                    const int hh = ra / d, j = (ra % d) >> 1;
                    if (hh < a.nh + a.nkv) {
                        const float cs = a.rope_cos[(size_t)rope_pos * half + j], sn = a.rope_sin[(size_t)rope_pos * half + j];
                        const float o1 = va * cs - vb * sn, o2 = va * sn + vb * cs;
                        if (hh < a.nh) {
                            a.q[(size_t)hh * d + j] = o1;
                            a.q[(size_t)hh * d + j + half] = o2;
                        } else {
                            uint16_t* kp = kpool + (((size_t)page_of_slot * a.nkv + (hh - a.nh)) * kKvPage + slot % kKvPage) * d;
                            kp[j] = f32_to_bf16_rne(o1);
                            kp[j + half] = f32_to_bf16_rne(o2);
                        }
                    } else {
                        uint16_t* vp = vpool + (((size_t)page_of_slot * a.nkv + (hh - a.nh - a.nkv)) * kKvPage + slot % kKvPage) * d;
                        *reinterpret_cast<uint32_t*>(vp + 2 * j) = (uint32_t)f32_to_bf16_rne(va) | ((uint32_t)f32_to_bf16_rne(vb) << 16);
                    }
                }
            }
            stamp(l);
            pk_grid_barrier(a.gbar, epoch, tid);
            stamp(l);

            // ---------------- P2: split-K attention over (kv head, split) items + last-arriver merge ----------------
            // Latency-optimised: every 16-lane group (two per warp, 16 per CTA) runs its OWN online softmax over a strided
            // subset of the item's tokens with all of its K/V loads in flight at once and q held in registers (loaded from L2
            // concurrently with K/V); the 16 partial states meet in shared memory after ONE block sync.
            {
                constexpr int HG = 4;                       // query heads per pass (register budget)
                constexpr int NG = kPkConsumerWarps * TPW;  // lane groups per CTA
                constexpr int BATCH = 4;                    // tokens in flight per lane group
                const int npages = (len + kKvPage - 1) / kKvPage;
                const int per = (npages + a.nsplit - 1) / a.nsplit;
                float* gacc = xs;                            // [NG][HG][D]
                float* gml = xs + NG * HG * D;               // [NG][HG][2]
                for (int item = cta; item < a.nkv * a.nsplit; item += ncta) {
                    const int kvh = item / a.nsplit, split = item % a.nsplit;
                    const int tok0 = split * per * kKvPage, tok1 = min(min((split + 1) * per, npages) * kKvPage, len);
                    const int gidx = warp * TPW + grp;
                    for (int h0 = 0; h0 < n_rep; h0 += HG) {
                        float qv[HG][8], acc[HG][8], mr[HG], lr[HG];
#pragma unroll
                        for (int h = 0; h < HG; ++h) {
                            mr[h] = -INFINITY;
                            lr[h] = 0.f;
#pragma unroll
                            for (int i = 0; i < 8; ++i) acc[h][i] = 0.f;
                            if (h0 + h < n_rep) {
                                const float4* qp = reinterpret_cast<const float4*>(a.q + (size_t)(kvh * n_rep + h0 + h) * D + gl * 8);
                                const float4 q0 = __ldcg(qp), q1 = __ldcg(qp + 1);
                                qv[h][0] = q0.x * a.qscale; qv[h][1] = q0.y * a.qscale; qv[h][2] = q0.z * a.qscale; qv[h][3] = q0.w * a.qscale;
                                qv[h][4] = q1.x * a.qscale; qv[h][5] = q1.y * a.qscale; qv[h][6] = q1.z * a.qscale; qv[h][7] = q1.w * a.qscale;
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) qv[h][i] = 0.f;
                            }
                        }
                        // trip count is uniform across the CTA (the shuffles below use the full mask); validity is per lane group
                        for (int tb0 = tok0; tb0 < tok1; tb0 += NG * BATCH) {
                            const int tb = tb0 + gidx;
                            uint4 kw[BATCH], vw[BATCH];
#pragma unroll
                            for (int it = 0; it < BATCH; ++it) {
                                const int tk = tb + it * NG;
                                if (tk < tok1) {
                                    const size_t base = (((size_t)__ldg(a.page_table + tk / kKvPage) * a.nkv + kvh) * kKvPage + tk % kKvPage) * D + gl * 8;
                                    kw[it] = ldcg_u4(kpool + base);
                                    vw[it] = ldcg_u4(vpool + base);
                                } else {
                                    kw[it] = make_uint4(0u, 0u, 0u, 0u);
                                    vw[it] = kw[it];
                                }
                            }
#pragma unroll
                            for (int it = 0; it < BATCH; ++it) {
                                const bool valid = tb + it * NG < tok1;
                                const float kf[8] = {bf16lo(kw[it].x), bf16hi(kw[it].x), bf16lo(kw[it].y), bf16hi(kw[it].y),
                                                     bf16lo(kw[it].z), bf16hi(kw[it].z), bf16lo(kw[it].w), bf16hi(kw[it].w)};
                                const float vf[8] = {bf16lo(vw[it].x), bf16hi(vw[it].x), bf16lo(vw[it].y), bf16hi(vw[it].y),
                                                     bf16lo(vw[it].z), bf16hi(vw[it].z), bf16lo(vw[it].w), bf16hi(vw[it].w)};
#pragma unroll
                                for (int h = 0; h < HG; ++h) {
                                    float sv = kf[0] * qv[h][0];
#pragma unroll
                                    for (int i = 1; i < 8; ++i) sv = fmaf(kf[i], qv[h][i], sv);
#pragma unroll
                                    for (int o = LPT / 2; o > 0; o >>= 1) sv += __shfl_xor_sync(0xFFFFFFFFu, sv, o);
                                    if (valid) {      // uniform inside the lane group
                                        const float m_new = fmaxf(mr[h], sv);
                                        const float al = expf(mr[h] - m_new), pr = expf(sv - m_new);
                                        lr[h] = lr[h] * al + pr;
                                        mr[h] = m_new;
#pragma unroll
                                        for (int i = 0; i < 8; ++i) acc[h][i] = fmaf(acc[h][i], al, pr * vf[i]);
                                    }
                                }
                            }
                        }
                        // the NG partial states meet in shared memory
#pragma unroll
                        for (int h = 0; h < HG; ++h) {
                            float* dst = gacc + ((size_t)gidx * HG + h) * D + gl * 8;
#pragma unroll
                            for (int i = 0; i < 8; ++i) dst[i] = acc[h][i];
                            if (gl == 0) {
                                gml[(gidx * HG + h) * 2] = mr[h];
                                gml[(gidx * HG + h) * 2 + 1] = lr[h];
                            }
                        }
                        pk_named_sync();
                        for (int i = tid; i < HG * D; i += kPkConsumers) {
                            const int h = i / D, dd = i % D;
                            if (h0 + h < n_rep) {
                                float M = -INFINITY;
#pragma unroll
                                for (int g = 0; g < NG; ++g) M = fmaxf(M, gml[(g * HG + h) * 2]);
                                float num = 0.f, den = 0.f;
                                if (M != -INFINITY) {
#pragma unroll
                                    for (int g = 0; g < NG; ++g) {
                                        const float wg = expf(gml[(g * HG + h) * 2] - M);      // exp(-inf) = 0 for idle groups
                                        num = fmaf(wg, gacc[((size_t)g * HG + h) * D + dd], num);
                                        den = fmaf(wg, gml[(g * HG + h) * 2 + 1], den);
                                    }
                                }
                                const size_t pidx = (size_t)(kvh * n_rep + h0 + h) * a.nsplit + split;
                                a.part_acc[pidx * D + dd] = num;
                                if (dd == 0) {
                                    a.part_ml[pidx * 2] = M;
                                    a.part_ml[pidx * 2 + 1] = den;
                                }
                            }
                        }
                        pk_named_sync();      // gacc / gml are reused by the next head group
                    }
                    __threadfence();
                    pk_named_sync();
                    if (tid == 0) *s_flag = (atomicAdd(&a.counters[kvh], 1) == a.nsplit - 1);
                    pk_named_sync();
                    if (*s_flag) {
                        __threadfence();
                        float* cm = xs;                      // [n_rep][nsplit] weights
                        float* cden = cm + kAttnMaxRep * a.nsplit;
                        if (warp < n_rep) {
                            const int h = warp;
                            const size_t base = (size_t)(kvh * n_rep + h) * a.nsplit;
                            float mstar = -INFINITY;
                            for (int s = lane; s < a.nsplit; s += 32) mstar = fmaxf(mstar, __ldcg(a.part_ml + (base + s) * 2));
                            mstar = warp_max(mstar);
                            float den = 0.f;
                            for (int s = lane; s < a.nsplit; s += 32) {
                                const float ms = __ldcg(a.part_ml + (base + s) * 2);
                                const float w = (ms == -INFINITY) ? 0.f : expf(ms - mstar);
                                cm[h * a.nsplit + s] = w;
                                den = fmaf(w, __ldcg(a.part_ml + (base + s) * 2 + 1), den);
                            }
                            den = warp_sum(den);
                            if (lane == 0) cden[h] = den;
                        }
                        pk_named_sync();
                        constexpr int D4 = D / 4;
                        for (int i = tid; i < n_rep * D4; i += kPkConsumers) {
                            const int h = i / D4, c4 = i % D4;
                            const float4* src = reinterpret_cast<const float4*>(a.part_acc + (size_t)(kvh * n_rep + h) * a.nsplit * D) + c4;
                            float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
                            for (int s = 0; s < a.nsplit; ++s) {
                                const float w = cm[h * a.nsplit + s];
                                if (w != 0.f) {
                                    const float4 v = __ldcg(src + (size_t)s * D4);
                                    num.x = fmaf(w, v.x, num.x); num.y = fmaf(w, v.y, num.y); num.z = fmaf(w, v.z, num.z); num.w = fmaf(w, v.w, num.w);
                                }
                            }
                            const float den = cden[h];
                            reinterpret_cast<float4*>(a.attn_out + (size_t)(kvh * n_rep + h) * D)[c4] =
                                make_float4(num.x / den, num.y / den, num.z / den, num.w / den);
                        }
                        if (tid == 0) a.counters[kvh] = 0;
                    }
                    pk_named_sync();
                }
            }
            stamp(l);
            pk_grid_barrier(a.gbar, epoch, tid);
            stamp(l);

            // ---------------- P3: o_proj + residual add ----------------
            load_x_plain(a.attn_out, nq);
            stamp(l);
            {
                const PkSlice s = consume(a.H, nq);
                stamp(l);
                if (a.tp > 1) {
                    allreduce_resid_add(s);
                } else {
                    for (int e = tid; e < (s.row_end - s.row_begin) / 2; e += kPkConsumers) {
                        float2* p = reinterpret_cast<float2*>(a.resid + s.row_begin + 2 * e);
                        float2 v = __ldcg(p);
                        v.x += row_sum(s, 2 * e);
                        v.y += row_sum(s, 2 * e + 1);
                        *p = v;
                    }
                }
            }
            stamp(l);
            pk_grid_barrier(a.gbar, epoch, tid);
            stamp(l);

            // ---------------- P4: RMSNorm -> gate|up -> SiLU(gate) * up ----------------
            load_x_rmsnorm(a.resid, false, nullptr, lw.ln2, a.H, nullptr);
            stamp(l);
            {
                const PkSlice s = consume(2 * a.I, a.H);
                stamp(l);
                for (int e = tid; e < (s.row_end - s.row_begin) / 2; e += kPkConsumers) {
                    const float g = row_sum(s, 2 * e), u = row_sum(s, 2 * e + 1);
                    a.act[(s.row_begin >> 1) + e] = g / (1.f + expf(-g)) * u;
                }
            }
            stamp(l);
            pk_grid_barrier(a.gbar, epoch, tid);
            stamp(l);

            // ---------------- P5: down_proj + residual add ----------------
            load_x_plain(a.act, a.I);
            stamp(l);
            {
                const PkSlice s = consume(a.H, a.I);
                stamp(l);
                if (a.tp > 1) {
                    allreduce_resid_add(s);
                } else {
                    for (int e = tid; e < (s.row_end - s.row_begin) / 2; e += kPkConsumers) {
                        float2* p = reinterpret_cast<float2*>(a.resid + s.row_begin + 2 * e);
                        float2 v = __ldcg(p);
                        v.x += row_sum(s, 2 * e);
                        v.y += row_sum(s, 2 * e + 1);
                        *p = v;
                    }
                }
            }
            stamp(l);
            pk_grid_barrier(a.gbar, epoch, tid);
            stamp(l);
        }

        // ---------------- final RMSNorm -> lm_head -> f32 logits + arg-max (last index wins ties) ----------------
        load_x_rmsnorm(a.resid, false, nullptr, a.final_norm, a.H, nullptr);
        {
            const PkSlice s = consume(a.V, a.H);
            float bv = -INFINITY;
            int bi = -1;
            for (int e = tid; e < (s.row_end - s.row_begin) / 2; e += kPkConsumers) {
                const int ra = a.rank * a.V + s.row_begin + 2 * e;      // index in the FULL vocabulary (vocab-parallel head)
                const float va = row_sum(s, 2 * e), vb = row_sum(s, 2 * e + 1);
                if (a.tp > 1) {
                    for (int r = 0; r < a.tp; ++r) *reinterpret_cast<float2*>(a.peer_logits[r] + ra) = make_float2(va, vb);
                } else {
                    *reinterpret_cast<float2*>(a.logits + ra) = make_float2(va, vb);
                }
                if (va >= bv) { bv = va; bi = ra; }
                if (vb >= bv) { bv = vb; bi = ra + 1; }
            }
            if (a.tp > 1) __threadfence_system();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
                const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
                if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { red[warp] = bv; reinterpret_cast<int*>(red)[8 + warp] = bi; }
            pk_named_sync();
            if (tid == 0) {
                for (int w = 1; w < kPkConsumerWarps; ++w) {
                    const float ov = red[w];
                    const int oi = reinterpret_cast<int*>(red)[8 + w];
                    if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
                }
                a.amax_val[cta] = bv;
                a.amax_idx[cta] = bi;
            }
        }
        pk_grid_barrier(a.gbar, epoch, tid);
        if (a.tp > 1) ar_epoch += 1;      // the arg-max exchange below is exchange number 2L+1 of the step (uniform in every thread)
        if (cta == 0 && warp == 0) {
            float bv = -INFINITY;
            int bi = -1;
            for (int p = lane; p < ncta; p += 32) {
                const float ov = __ldcg(a.amax_val + p);
                const int oi = __ldcg(a.amax_idx + p);
                if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
                const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
                if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
            }
            if (a.tp > 1) {   // exchange the per-rank winners (and, by cumulativity, make every rank's logits slice visible)
                if (lane == 0) {
                    for (int r = 0; r < a.tp; ++r) {
                        a.peer_amax[r][a.rank * 2] = bv;
                        a.peer_amax[r][a.rank * 2 + 1] = __int_as_float(bi);
                    }
                    __threadfence_system();
                }
                __syncwarp();
                if (lane < a.tp) {
                    st_release_sys(a.peer_flag[lane] + (size_t)a.rank * (ncta + 1) + ncta, ar_epoch);
                    pk_wait_flag(a.peer_flag[a.rank] + (size_t)lane * (ncta + 1) + ncta, ar_epoch, a.comm_err);
                }
                __syncwarp();
                bv = -INFINITY;
                bi = -1;
                for (int r = 0; r < a.tp; ++r) {
                    const float ov = __ldcg(a.peer_amax[a.rank] + r * 2);
                    const int oi = __float_as_int(__ldcg(a.peer_amax[a.rank] + r * 2 + 1));
                    if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
                }
            }
            if (lane == 0) {
                a.next_ids[0] = (uint32_t)bi;
                a.state->kv_base[0] = kv_base + 1;
                if (a.feedback) {
                    a.ids[0] = (uint32_t)bi;
                    a.state->rope_pos = rope_pos + 1;
                    if (a.trace != nullptr) {
                        const int tp = *a.trace_pos;
                        a.trace[tp] = (uint32_t)bi;
                        *a.trace_pos = tp + 1;
                    }
                }
            }
        }
        if (step + 1 < a.nsteps) pk_grid_barrier(a.gbar, epoch, tid);
    }
}

}  // namespace fl
