// Sampling step of the generate loop (host code, no device work): what the reference does with the logits every
// forward returns -- `LogitsProcessor::new(Default::default(), Some(temperature as f64), None)` and
// `logits_processor.sample(&last_logits)` (src/models/mod.rs:157-158, 308-310, 373-374, 425-428).
//
// The reference delegates to candle-transformers 0.8 `generation::LogitsProcessor`, which uses rand 0.8.5: a `StdRng`
// (ChaCha12) seeded with `seed_from_u64`, candle-nn's `softmax_last_dim`, and `WeightedIndex<f32>`.  A drop-in has to
// reproduce the token stream for a given seed, so each of those pieces is implemented here to the published algorithm
// (summation orders included: both the soft-max denominator and the cumulative weights are sequential f32 sums).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace fl {

// rand_chacha 0.3 ChaCha12Rng behind rand_core's BlockRng: key-stream words handed out in order from block counter 0,
// stream id 0.  Layout: 4 constants | 8 key words | 64-bit counter | 64-bit stream id.
class StdRng {
   public:
    // rand_core 0.6 SeedableRng::seed_from_u64: PCG32 output per 4 seed bytes
    explicit StdRng(uint64_t seed_u64) {
        uint64_t state = seed_u64;
        for (int i = 0; i < 8; ++i) {
            state = state * 6364136223846793005ull + 11634580027462260723ull;
            const uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            const uint32_t rot = (uint32_t)(state >> 59);
            key_[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        }
    }
    explicit StdRng(const uint8_t seed[32]) {   // SeedableRng::from_seed: little-endian key words
        for (int i = 0; i < 8; ++i)
            key_[i] = (uint32_t)seed[4 * i] | (uint32_t)seed[4 * i + 1] << 8 | (uint32_t)seed[4 * i + 2] << 16 | (uint32_t)seed[4 * i + 3] << 24;
    }
    uint32_t next_u32() {
        if (idx_ == 16) refill();
        return buf_[idx_++];
    }

   private:
    static uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
    static void quarter(uint32_t* x, int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    }
    void refill() {
        uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
        for (int i = 0; i < 8; ++i) init[4 + i] = key_[i];
        init[12] = (uint32_t)counter_;
        init[13] = (uint32_t)(counter_ >> 32);
        init[14] = init[15] = 0;
        uint32_t x[16];
        std::memcpy(x, init, sizeof(x));
        for (int r = 0; r < 6; ++r) {   // 12 rounds = 6 column/diagonal double rounds
            quarter(x, 0, 4, 8, 12); quarter(x, 1, 5, 9, 13); quarter(x, 2, 6, 10, 14); quarter(x, 3, 7, 11, 15);
            quarter(x, 0, 5, 10, 15); quarter(x, 1, 6, 11, 12); quarter(x, 2, 7, 8, 13); quarter(x, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; ++i) buf_[i] = x[i] + init[i];
        counter_++;
        idx_ = 0;
    }
    uint32_t key_[8];
    uint32_t buf_[16];
    uint64_t counter_ = 0;
    int idx_ = 16;
};

inline float f32_from_bits(uint32_t b) {
    float f;
    std::memcpy(&f, &b, 4);
    return f;
}
inline uint32_t f32_bits(float f) {
    uint32_t b;
    std::memcpy(&b, &f, 4);
    return b;
}

// f32::total_cmp key: sign-magnitude bits -> monotone signed integer
inline int32_t total_order_key(float f) {
    int32_t b = (int32_t)f32_bits(f);
    return b ^ (int32_t)((uint32_t)(b >> 31) >> 1);
}

// largest order key of v[0..n): a branch-free max reduction the compiler vectorises; the AVX2 clone is picked at run time
#define FL_MAX_KEY_BODY                                  \
    int32_t kb = total_order_key(v[0]);                  \
    for (size_t i = 1; i < n; ++i) {                     \
        const int32_t k = total_order_key(v[i]);         \
        kb = k > kb ? k : kb;                            \
    }                                                    \
    return kb;
static inline int32_t max_key_generic(const float* v, size_t n) { FL_MAX_KEY_BODY }
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) static inline int32_t max_key_avx2(const float* v, size_t n) { FL_MAX_KEY_BODY }
static inline int32_t max_key(const float* v, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    return avx2 ? max_key_avx2(v, n) : max_key_generic(v, n);
}
#else
static inline int32_t max_key(const float* v, size_t n) { return max_key_generic(v, n); }
#endif
#undef FL_MAX_KEY_BODY

// LogitsProcessor::sample_argmax: iter().enumerate().max_by(|(_, u), (_, v)| u.total_cmp(v)) -- max_by keeps the LAST maximum
inline uint32_t sample_argmax(const float* v, size_t n) {
    const int32_t kb = max_key(v, n);
    size_t best = n - 1;
    while (total_order_key(v[best]) != kb) --best;
    return (uint32_t)best;
}

struct SamplerError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class LogitsProcessor {
   public:
    // LogitsProcessor::new(seed, temperature, None): temperature < 1e-7 (or negative "None") => Sampling::ArgMax
    LogitsProcessor(uint64_t seed, double temperature) : rng_(seed), argmax_(!(temperature >= 1e-7)), temperature_(temperature) {}

    uint32_t next_u32() { return rng_.next_u32(); }
    bool is_argmax() const { return argmax_; }
    float inv_temperature() const { return (float)(1.0 / temperature_); }
    // the ONE generator draw a temperature sample consumes, as UniformFloat::<f32> turns it into [0, 1): (u32 >> 9) as the mantissa
    // of a float in [1, 2), minus 1 (device-side sampling takes it from here, so a request's random stream is the same on both paths)
    float draw_unit() { return f32_from_bits((rng_.next_u32() >> 9) | 0x3f800000u) - 1.0f; }

    uint32_t sample(const float* logits, size_t n) {
        if (n == 0) throw SamplerError("sample: empty logits");
        if (argmax_) return sample_argmax(logits, n);
        // (&logits / temperature) = affine(1/T, 0) in f32, then softmax_last_dim: max by fold, exp(s - max), sequential sum, divide
        const float mul = (float)(1.0 / temperature_);
        prs_.resize(n);
        float mx = -INFINITY;
        for (size_t i = 0; i < n; ++i) {
            prs_[i] = logits[i] * mul + 0.0f;
            mx = std::fmax(mx, prs_[i]);   // f32::max: a NaN operand is ignored
        }
        float sum_exp = 0.0f;
        for (size_t i = 0; i < n; ++i) {
            prs_[i] = std::exp(prs_[i] - mx);
            sum_exp += prs_[i];
        }
        for (size_t i = 0; i < n; ++i) prs_[i] /= sum_exp;
        return sample_multinomial();
    }

   private:
    // WeightedIndex::<f32>::new(prs)?.sample(&mut rng)
    uint32_t sample_multinomial() {
        const size_t n = prs_.size();
        cum_.resize(n - 1);
        float total = prs_[0];
        if (!(total >= 0.0f)) throw SamplerError("sample: WeightedError::InvalidWeight (a probability is negative or NaN)");
        for (size_t i = 1; i < n; ++i) {
            if (!(prs_[i] >= 0.0f)) throw SamplerError("sample: WeightedError::InvalidWeight (a probability is negative or NaN)");
            cum_[i - 1] = total;
            total += prs_[i];
        }
        if (total == 0.0f) throw SamplerError("sample: WeightedError::AllWeightsZero");
        // UniformFloat::<f32>::new(0, total): shrink the scale until the largest sample stays below `high`
        if (!std::isfinite(total)) throw SamplerError("sample: Uniform::new called with a non-finite range");
        const float max_rand = f32_from_bits(0x3f800000u | 0x7fffffu) - 1.0f;   // 1 - 2^-23
        float scale = total;
        while (scale * max_rand + 0.0f >= total) scale = f32_from_bits(f32_bits(scale) - 1);
        const float value0_1 = f32_from_bits((rng_.next_u32() >> 9) | 0x3f800000u) - 1.0f;
        const float chosen = value0_1 * scale + 0.0f;
        // partition_point(|w| w <= chosen)
        size_t lo = 0, hi = cum_.size();
        while (lo < hi) {
            const size_t mid = lo + (hi - lo) / 2;
            if (cum_[mid] <= chosen) lo = mid + 1; else hi = mid;
        }
        return (uint32_t)lo;
    }

    StdRng rng_;
    bool argmax_;
    double temperature_;
    std::vector<float> prs_, cum_;
};

}  // namespace fl
