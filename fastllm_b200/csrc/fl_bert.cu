// BERT / MiniLM sentence-encoder path behind fl_embed (host orchestration), sm_100a.
// Reference: MiniLMModel::{new, embed_tokens, forward, mean_pooling, normalize_l2, embed} src/models/embeddings.rs:245-447.
#include <cmath>
#include <memory>

#include "bert.cuh"
#include "bert_model.cuh"
#include "gemm_tc.cuh"
#include "synth.cuh"

namespace fl {

static inline size_t align_up_b(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- TMA descriptors (driver entry point fetched through the runtime: no link-time libcuda dependency) -----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        FL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        FL_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, -2, "cuTensorMapEncodeTiled not available from the driver");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D bf16 tensor [rows, cols] row-major (row stride ld elements), box [box_rows, 64 cols], 128-byte swizzle.
// Output-tile width of the encoder GEMMs.  With hidden = 384 every GEMM of the model has K-major operands of equal depth, so a
// 128 x BN tile moves (128 + BN) * K * 2 bytes from L2 for 2 * 128 * BN * K FLOP: 64 FLOP/B at BN = 128 -- L2->SM bound
// (~7 TB/s chip-wide => ~450 TFLOP/s) long before the tensor pipe.  BN = 192 divides all three widths (1152, 384, 1536), lifts
// the intensity to 77 FLOP/B and still fits two TMEM accumulator stages (2 x 256 columns).
constexpr int kBertBN = 192;

CUtensorMap make_tmap_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kGemmBK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FL_CHECK(r == CUDA_SUCCESS, -2, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

// Weight matrix [N, K] bf16 (K % 64 == 0) as the persistent decode kernel streams it: a 3-D view {64 columns, N rows, K/64
// column blocks} whose box {64, 16, 16} is ONE 32 KB request = 16 rows x 1024 columns, landing in shared memory as
// [column block][row][128 bytes] with the 128-byte swizzle keyed by the row: conflict-free ldmatrix of 16 x 16 tiles.
// Out-of-range rows / column blocks are zero-filled by the TMA engine (no tail handling in the kernel).
CUtensorMap make_tmap_pk(const void* ptr, uint64_t N, uint64_t K) {
    CUtensorMap m;
    const cuuint64_t dims[3] = {64, N, K / 64};
    const cuuint64_t strides[2] = {K * 2, 128};
    const cuuint32_t box[3] = {64, 16, 16};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FL_CHECK(r == CUDA_SUCCESS, -2, "cuTensorMapEncodeTiled (3-D weight stream) failed (" + std::to_string((int)r) + ")");
    return m;
}

// K or V pool of a cache seen as one [rows, head_dim] bf16 matrix (rows = layers x pages x kv heads x 64 tokens): the stream-K decode
// attention fetches a page of one kv head as 64-row boxes of 64 columns (128 bytes, 128-byte swizzle); test-size heads (< 64) as one
// unswizzled [64, head_dim] box.
CUtensorMap make_tmap_kv(const void* ptr, uint64_t rows, uint32_t head_dim) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {head_dim, rows};
    const cuuint64_t strides[1] = {(cuuint64_t)head_dim * 2};
    const cuuint32_t box[2] = {head_dim >= 64 ? 64u : head_dim, 64u};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, head_dim >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FL_CHECK(r == CUDA_SUCCESS, -2, "cuTensorMapEncodeTiled (KV pool) failed (" + std::to_string((int)r) + ")");
    return m;
}

// Every kernel of the encoder is launched with programmatic stream serialization: each calls griddepcontrol.launch_dependents at
// entry and griddepcontrol.wait before it touches anything the previous kernel wrote, so a kernel's prologue (barrier / TMEM
// set-up, descriptor fetch, block scheduling) overlaps its predecessor's tail instead of following it.
static bool g_bert_pdl = !env_flag("FL_NO_PDL");
template <typename... KArgs, typename... Args>
static void blaunch(cudaStream_t st, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (g_bert_pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    FL_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

template <int EPI>
static void launch_gemm(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g) {
    constexpr int BN = kBertBN;
    const size_t smem = gemm_smem_bytes(BN, DUAL_NONE) + (gemm_epi_staged(EPI) ? kEpiStageBytes : 0);
    static bool attr_set = false;
    if (!attr_set) {
        FL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int tiles = ((g.M + kGemmBM - 1) / kGemmBM) * ((g.N + BN - 1) / BN);
    blaunch(st, gemm_tc_kernel<BN, EPI>, dim3(std::min(tiles, kNumSMs)), dim3(kGemmThreads), smem, tmA, tmA, tmB, g);
}

// ---- weights -------------------------------------------------------------------------------------------------------------
void bert_build(BertModel& m) {
    const fl_config& c = m.cfg;
    m.H = c.hidden_size; m.I = c.intermediate_size; m.V = c.vocab_size; m.L = c.num_hidden_layers; m.nh = c.num_attention_heads;
    m.d = m.H / m.nh; m.maxpos = c.max_position_embeddings;
    FL_CHECK(m.d * m.nh == m.H, FL_ERR_INVALID, "hidden_size must be divisible by num_attention_heads");
    FL_CHECK(m.d == kBertD, FL_ERR_UNSUPPORTED, "BERT path is built for head_dim 32 (all-MiniLM-L6-v2); other head sizes not built yet");
    FL_CHECK(m.H % 64 == 0 && m.I % 64 == 0 && m.H <= 512, FL_ERR_UNSUPPORTED, "BERT path needs hidden/intermediate % 64 == 0 and hidden <= 512");
    const size_t H = m.H, I = m.I;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up_b(off + bytes, 256); return o; };
    const size_t o_w = take((size_t)m.V * H * 2), o_p = take((size_t)m.maxpos * H * 2), o_lw = take(H * 4), o_lb = take(H * 4);
    struct LO { size_t wqkv, wo, wi, wo2, bqkv, bo, bi, bo2, l1w, l1b, l2w, l2b; };
    std::vector<LO> lo(m.L);
    for (auto& l : lo) {
        l.wqkv = take(3 * H * H * 2); l.wo = take(H * H * 2); l.wi = take(I * H * 2); l.wo2 = take(H * I * 2);
        l.bqkv = take(3 * H * 4); l.bo = take(H * 4); l.bi = take(I * 4); l.bo2 = take(H * 4);
        l.l1w = take(H * 4); l.l1b = take(H * 4); l.l2w = take(H * 4); l.l2b = take(H * 4);
    }
    m.slab.alloc(off, true);
    uint8_t* b = m.slab.p;
    m.wemb = (uint16_t*)(b + o_w); m.pemb = (uint16_t*)(b + o_p); m.lnw = (float*)(b + o_lw); m.lnb = (float*)(b + o_lb);
    m.layers.resize(m.L);
    for (int i = 0; i < m.L; ++i) {
        BertLayerW& w = m.layers[i];
        const LO& l = lo[i];
        w.wqkv = (uint16_t*)(b + l.wqkv); w.wo = (uint16_t*)(b + l.wo); w.wi = (uint16_t*)(b + l.wi); w.wo2 = (uint16_t*)(b + l.wo2);
        w.bqkv = (float*)(b + l.bqkv); w.bo = (float*)(b + l.bo); w.bi = (float*)(b + l.bi); w.bo2 = (float*)(b + l.bo2);
        w.ln1w = (float*)(b + l.l1w); w.ln1b = (float*)(b + l.l1b); w.ln2w = (float*)(b + l.l2w); w.ln2b = (float*)(b + l.l2b);
    }
    FL_CUDA(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
}

struct BertRoute {
    bool mat;
    void* base;
    int64_t rows, cols, row0;
};

static bool bert_route(BertModel& m, const std::string& name, BertRoute& r) {
    const int64_t H = m.H, I = m.I;
    auto mat = [&](uint16_t* p, int64_t rows, int64_t cols, int64_t row0 = 0) { r = {true, p, rows, cols, row0}; return true; };
    auto vec = [&](float* p, int64_t rows, int64_t row0 = 0) { r = {false, p, rows, 1, row0}; return true; };
    if (name == "embeddings.word_embeddings.weight") return mat(m.wemb, m.V, H);
    if (name == "embeddings.position_embeddings.weight") return mat(m.pemb, m.maxpos, H);
    if (name == "embeddings.LayerNorm.weight") return vec(m.lnw, H);
    if (name == "embeddings.LayerNorm.bias") return vec(m.lnb, H);
    const std::string pre = "encoder.layer.";
    if (name.compare(0, pre.size(), pre) != 0) return false;
    const size_t dot = name.find('.', pre.size());
    if (dot == std::string::npos) return false;
    int li = -1;
    try { li = std::stoi(name.substr(pre.size(), dot - pre.size())); } catch (...) { return false; }
    if (li < 0 || li >= m.L) return false;
    BertLayerW& w = m.layers[li];
    const std::string rest = name.substr(dot + 1);
    if (rest == "attention.self.query.weight") return mat(w.wqkv, H, H, 0);
    if (rest == "attention.self.key.weight") return mat(w.wqkv, H, H, H);
    if (rest == "attention.self.value.weight") return mat(w.wqkv, H, H, 2 * H);
    if (rest == "attention.self.query.bias") return vec(w.bqkv, H, 0);
    if (rest == "attention.self.key.bias") return vec(w.bqkv, H, H);
    if (rest == "attention.self.value.bias") return vec(w.bqkv, H, 2 * H);
    if (rest == "attention.output.dense.weight") return mat(w.wo, H, H);
    if (rest == "attention.output.dense.bias") return vec(w.bo, H);
    if (rest == "attention.output.LayerNorm.weight") return vec(w.ln1w, H);
    if (rest == "attention.output.LayerNorm.bias") return vec(w.ln1b, H);
    if (rest == "intermediate.dense.weight") return mat(w.wi, I, H);
    if (rest == "intermediate.dense.bias") return vec(w.bi, I);
    if (rest == "output.dense.weight") return mat(w.wo2, H, I);
    if (rest == "output.dense.bias") return vec(w.bo2, H);
    if (rest == "output.LayerNorm.weight") return vec(w.ln2w, H);
    if (rest == "output.LayerNorm.bias") return vec(w.ln2b, H);
    return false;
}

static std::vector<std::string> bert_expected(const BertModel& m) {
    std::vector<std::string> v = {"embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
                                  "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"};
    for (int l = 0; l < m.L; ++l) {
        const std::string p = "encoder.layer." + std::to_string(l) + ".";
        for (const char* s : {"attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense",
                              "intermediate.dense", "output.dense", "attention.output.LayerNorm", "output.LayerNorm"}) {
            v.push_back(p + s + ".weight");
            v.push_back(p + s + ".bias");
        }
    }
    return v;
}

static uint16_t h_bf16(float f) {
    uint32_t b;
    std::memcpy(&b, &f, 4);
    if ((b & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((b >> 16) | 0x40u);   // NaN stays NaN (quiet), as half::bf16::from_f32 does
    return (uint16_t)((b + 0x7FFFu + ((b >> 16) & 1u)) >> 16);
}

void bert_put_tensor(BertModel& m, const char* name, int dtype, const int64_t* shape, int rank, const void* host) {
    FL_CHECK(!m.finalized, FL_ERR_STATE, "model already finalized");
    const std::string n(name);
    // tensors the reference never reads (pooler, token-type embeddings, position_ids buffer, anything else in the checkpoint):
    // accepted and ignored, as VarBuilder ignores the extra entries of the loaded HashMap (embeddings.rs:290-298); a MISSING tensor
    // is still an error at finalize
    BertRoute r;
    if (!bert_route(m, n, r)) return;
    FL_CHECK(dtype == FL_DTYPE_F32, FL_ERR_UNSUPPORTED, "BERT tensors must be f32 (the reference loads MiniLM as F32, embeddings.rs:298)");
    const bool ok = r.mat ? (rank == 2 && shape[0] == r.rows && shape[1] == r.cols) : (rank == 1 && shape[0] == r.rows);
    FL_CHECK(ok, FL_ERR_INVALID, "shape mismatch for " + n);
    const float* f = (const float*)host;
    if (r.mat) {
        std::vector<uint16_t> bits((size_t)r.rows * r.cols);
        for (size_t i = 0; i < bits.size(); ++i) bits[i] = h_bf16(f[i]);
        FL_CUDA(cudaMemcpy((uint16_t*)r.base + r.row0 * r.cols, bits.data(), bits.size() * 2, cudaMemcpyHostToDevice));
    } else {
        FL_CUDA(cudaMemcpy((float*)r.base + r.row0, f, (size_t)r.rows * 4, cudaMemcpyHostToDevice));
    }
    m.have.insert(n);
}

void bert_random_init(BertModel& m, uint64_t seed, float stdv) {
    FL_CHECK(!m.finalized, FL_ERR_STATE, "model already finalized");
    for (const std::string& n : bert_expected(m)) {
        BertRoute r;
        FL_CHECK(bert_route(m, n, r), FL_ERR_INVALID, "internal: unroutable " + n);
        const bool is_ln_w = n.size() >= 16 && n.compare(n.size() - 16, 16, "LayerNorm.weight") == 0;
        const uint64_t ts = tensor_seed(seed, n.c_str());
        const RowMap map{r.row0, 0, 0, 0};
        if (r.mat)
            synth_fill_bf16_kernel<<<kNumSMs * 4, 256>>>((uint16_t*)r.base, r.rows, r.cols, r.cols, ts, stdv, map);
        else if (is_ln_w)
            fill_f32_kernel<<<4, 256>>>((float*)r.base + r.row0, r.rows, 1.0f);
        else
            synth_fill_f32_kernel<<<4, 256>>>((float*)r.base, r.rows, ts, stdv, map);
        g_launches.fetch_add(1);
        m.have.insert(n);
    }
    FL_CUDA(cudaGetLastError());
    FL_CUDA(cudaDeviceSynchronize());
}

void bert_finalize(BertModel& m) {
    FL_CHECK(!m.finalized, FL_ERR_STATE, "model already finalized");
    for (const std::string& n : bert_expected(m)) FL_CHECK(m.have.count(n), FL_ERR_STATE, "missing tensor: " + n);
    for (BertLayerW& w : m.layers) {   // weight-side TMA descriptors ([out, in] row-major == K-major B operand)
        w.tm_wqkv = make_tmap_bf16(w.wqkv, 3 * m.H, m.H, m.H, kBertBN);
        w.tm_wo = make_tmap_bf16(w.wo, m.H, m.H, m.H, kBertBN);
        w.tm_wi = make_tmap_bf16(w.wi, m.I, m.H, m.H, kBertBN);
        w.tm_wo2 = make_tmap_bf16(w.wo2, m.H, m.I, m.I, kBertBN);
    }
    m.finalized = true;
}

static void bert_reserve(BertModel& m, int b, int t) {
    const size_t T = (size_t)b * t;
    if (T <= m.cap_tokens && (size_t)b <= m.cap_batch) return;
    const size_t H = m.H, I = m.I;
    m.cap_tokens = std::max(T, m.cap_tokens);
    m.cap_batch = std::max((size_t)b, m.cap_batch);
    const size_t Tc = m.cap_tokens;
    m.x.alloc(Tc * H); m.x1.alloc(Tc * H); m.ctx.alloc(Tc * H); m.qkv.alloc(Tc * 3 * H); m.hbuf.alloc(Tc * I); m.pre.alloc(Tc * H);
    m.ids.alloc(Tc); m.mask.alloc(Tc); m.out.alloc(m.cap_batch * H);
    m.h_ids.alloc(2 * Tc); m.h_out.alloc(m.cap_batch * H);
}

// Enqueue the encoder on m.stream: B1 -> 6 x (B2..B7) -> B8
static void bert_enqueue(BertModel& m, int b, int t, bool has_mask) {
    const int T = b * t, H = m.H, I = m.I;
    cudaStream_t st = m.stream;
    const int rows_per_cta = 8;
    const dim3 row_grid((T + rows_per_cta - 1) / rows_per_cta), row_block(rows_per_cta * 32);
    blaunch(st, embed_ln_kernel, row_grid, row_block, 0, (const uint16_t*)m.wemb, (const uint16_t*)m.pemb, (const float*)m.lnw, (const float*)m.lnb,
            (const uint32_t*)m.ids.p, T, t, H, m.V, m.maxpos, 1e-12f, m.x.p);
    const CUtensorMap tm_x = make_tmap_bf16(m.x.p, T, H, H, kGemmBM), tm_x1 = make_tmap_bf16(m.x1.p, T, H, H, kGemmBM),
                      tm_ctx = make_tmap_bf16(m.ctx.p, T, H, H, kGemmBM), tm_h = make_tmap_bf16(m.hbuf.p, T, I, I, kGemmBM);
    const float eps = m.cfg.norm_eps;
    const float scale = (float)std::sqrt((double)m.d);
    for (int l = 0; l < m.L; ++l) {
        const BertLayerW& w = m.layers[l];
        // B2: fused q|k|v projection + bias -> bf16 [T, 3H]
        launch_gemm<GEPI_BIAS_BF16>(st, tm_x, w.tm_wqkv, GemmArgs{T, 3 * H, H, w.bqkv, nullptr, 0, m.qkv.p, 3 * H, 1, 0});
        // B3+B4: per (sentence, head) softmax(QK^T / sqrt(d)) V, no mask
        if (t <= kBertS)
            blaunch(st, bert_attn_kernel, dim3(m.nh, b), dim3(128), 0, (const uint16_t*)m.qkv.p, t, H, scale, m.ctx.p);
        else
            blaunch(st, bert_attn_long_kernel, dim3(m.nh, b, (t + kBertS - 1) / kBertS), dim3(128), 0, (const uint16_t*)m.qkv.p, t, H, scale, m.ctx.p);
        // B5: attention output dense + bias + residual -> f32, then LayerNorm -> bf16
        launch_gemm<GEPI_BIAS_RESID_F32>(st, tm_ctx, w.tm_wo, GemmArgs{T, H, H, w.bo, m.x.p, H, m.pre.p, H, 1, 0});
        blaunch(st, layernorm_kernel, row_grid, row_block, 0, (const float*)m.pre.p, (const float*)w.ln1w, (const float*)w.ln1b, T, H, eps, m.x1.p);
        // B6: intermediate dense + bias + GELU(tanh) -> bf16 [T, I]
        launch_gemm<GEPI_BIAS_GELU_BF16>(st, tm_x1, w.tm_wi, GemmArgs{T, I, H, w.bi, nullptr, 0, m.hbuf.p, I, 1, 0});
        // B7: output dense + bias + residual -> f32, LayerNorm -> bf16
        launch_gemm<GEPI_BIAS_RESID_F32>(st, tm_h, w.tm_wo2, GemmArgs{T, H, I, w.bo2, m.x1.p, H, m.pre.p, H, 1, 0});
        blaunch(st, layernorm_kernel, row_grid, row_block, 0, (const float*)m.pre.p, (const float*)w.ln2w, (const float*)w.ln2b, T, H, eps, m.x.p);
    }
    // B8: masked mean pooling + L2 normalise -> f32 [b, H]
    blaunch(st, pool_l2_kernel, dim3(b), dim3((H + 31) / 32 * 32), 0, (const uint16_t*)m.x.p, has_mask ? (const uint32_t*)m.mask.p : (const uint32_t*)nullptr,
            t, H, m.out.p);
}

void bert_embed(BertModel& m, const uint32_t* ids, const uint32_t* mask, int b, int t, float* out, float* device_ms) {
    FL_CHECK(m.finalized, FL_ERR_STATE, "model not finalized");
    FL_CHECK(ids && out && b >= 1 && t >= 1, FL_ERR_INVALID, "bad arguments");
    // the reference enforces no max_seq_length (embeddings.rs:285-286); what bounds a sentence is its position table: position ids
    // 0..n index a [max_position_embeddings, H] embedding (embeddings.rs:416, 370-378), and candle's index_select fails past it
    FL_CHECK(t <= m.maxpos, FL_ERR_INVALID, "sentence longer than max_position_embeddings (the position-embedding lookup fails in the reference too)");
    for (int i = 0; i < b * t; ++i) FL_CHECK(ids[i] < (uint32_t)m.V, FL_ERR_INVALID, "token id out of range");
    std::lock_guard<std::mutex> lock(m.mu);      // embed(&self) may be called from several threads: one workspace
    bert_reserve(m, b, t);
    const size_t T = (size_t)b * t;
    std::memcpy(m.h_ids.p, ids, T * 4);
    FL_CUDA(cudaMemcpyAsync(m.ids.p, m.h_ids.p, T * 4, cudaMemcpyHostToDevice, m.stream));
    if (mask) {
        std::memcpy(m.h_ids.p + T, mask, T * 4);
        FL_CUDA(cudaMemcpyAsync(m.mask.p, m.h_ids.p + T, T * 4, cudaMemcpyHostToDevice, m.stream));
    }
    std::unique_ptr<EventPair> ev;
    if (device_ms) {
        ev.reset(new EventPair);
        FL_CUDA(cudaEventRecord(ev->e0, m.stream));
    }
    bert_enqueue(m, b, t, mask != nullptr);
    if (device_ms) FL_CUDA(cudaEventRecord(ev->e1, m.stream));
    FL_CUDA(cudaGetLastError());
    FL_CUDA(cudaMemcpyAsync(m.h_out.p, m.out.p, (size_t)b * m.H * 4, cudaMemcpyDeviceToHost, m.stream));
    FL_CUDA(cudaStreamSynchronize(m.stream));
    std::memcpy(out, m.h_out.p, (size_t)b * m.H * 4);
    if (device_ms) FL_CUDA(cudaEventElapsedTime(device_ms, ev->e0, ev->e1));
}

// Device-resident repeat of the encoder on the ids already uploaded by the last bert_embed (benchmark `value`).
void bert_repeat(BertModel& m, int b, int t, int iters, float* elapsed_ms) {
    std::lock_guard<std::mutex> lock(m.mu);
    FL_CHECK((size_t)b * t <= m.cap_tokens, FL_ERR_STATE, "call fl_embed with this shape first");
    EventPair ev;
    FL_CUDA(cudaEventRecord(ev.e0, m.stream));
    for (int i = 0; i < iters; ++i) bert_enqueue(m, b, t, false);
    FL_CUDA(cudaEventRecord(ev.e1, m.stream));
    FL_CUDA(cudaStreamSynchronize(m.stream));
    FL_CUDA(cudaGetLastError());
    FL_CUDA(cudaEventElapsedTime(elapsed_ms, ev.e0, ev.e1));
}

BertModel::~BertModel() {
    if (stream) cudaStreamDestroy(stream);
}

}  // namespace fl
