/* Fast host path of the synthetic-weight rule in oracle/synth.py (TEST INFRASTRUCTURE).
 * Same arithmetic as fastllm_b200/csrc/synth.cuh; see oracle/synth.py for the rule.
 * Build: gcc -O3 -fopenmp -shared -fPIC -o oracle/_build/liboracle_synth.so oracle/synth_gen.c
 */
#include <stdint.h>
#include <string.h>

#define GOLDEN 0x9E3779B97F4A7C15ULL

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline uint16_t f32_to_bf16_rne(float f) {
    uint32_t b;
    memcpy(&b, &f, 4);
    return (uint16_t)((b + 0x7FFFu + ((b >> 16) & 1u)) >> 16);
}

void synth_normal_bf16(uint64_t tseed, float std, uint64_t start, uint64_t count, uint16_t* out) {
    const uint32_t unit_bits = 0x37DDB3D7u; /* f32(1/37837.22719439421), checked in tests/test_synth.py */
    float unit;
    memcpy(&unit, &unit_bits, 4);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)count; ++j) {
        uint64_t i = start + (uint64_t)j;
        uint64_t h = mix64(tseed + (i + 1) * GOLDEN);
        int32_t s = (int32_t)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48));
        volatile float v = (float)(s - 131070) * unit; /* volatile: forbid contraction / reassociation */
        v = v * std;
        out[j] = f32_to_bf16_rne(v);
    }
}
