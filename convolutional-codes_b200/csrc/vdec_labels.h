// vdec_labels.h — trellis branch labels, usable on host, on device and in constant expressions.
//
// A label is the n-bit symbol the encoder would emit for (state s, input bit i); the decoders
// compare the received step against it.  Reference: get_transition_metric(),
// AWGN-channel/viterbi-decoder.c:45-63 (identical text in the hard decoder :45-63 and in
// common/encoder.c:93-106).  The reference's parity expression `(0x6996 >> val) & 1` shifts an
// int by a 64-bit count; on x86-64 that evaluates to `(b & 0x10) ? 0 : parity(b)` with
// b = xor of the 8 bytes of (shift_register & polynomial) (SURVEY.md finding 1).  Both that
// de-facto labelling (VDEC_PARITY_REF) and true parity (VDEC_PARITY_LINEAR) are provided.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VDEC_HD __host__ __device__
#else
#define VDEC_HD
#endif

namespace vd {

VDEC_HD constexpr uint32_t fold_byte(uint64_t v) {
    v ^= v >> 32; v ^= v >> 16; v ^= v >> 8;
    return (uint32_t)(v & 0xFFu);
}
VDEC_HD constexpr uint32_t parity_bit(uint64_t masked, bool ref_quirk) {
    const uint32_t b = fold_byte(masked);
    const uint32_t nib = (b ^ (b >> 4)) & 0xFu;
    const uint32_t par = (0x6996u >> nib) & 1u;
    return (ref_quirk && (b & 0x10u)) ? 0u : par;
}
// expected symbol, P0's parity in the most significant of the n bits
VDEC_HD constexpr uint32_t branch_label(uint32_t K, uint32_t n, const uint64_t* polys, bool ref_quirk,
                                        uint32_t state, uint32_t input) {
    const uint64_t sr = ((uint64_t)state << (64u - K)) | ((uint64_t)input << 63);
    uint32_t sym = 0;
    for (uint32_t j = 0; j < n; ++j) sym = (sym << 1) | parity_bit(sr & polys[j], ref_quirk);
    return sym;
}
// rotate a (bits)-wide index left / right by p
VDEC_HD constexpr uint32_t rotl_bits(uint32_t x, uint32_t p, uint32_t bits) {
    p %= bits;
    return p == 0 ? x : (((x << p) | (x >> (bits - p))) & ((1u << bits) - 1u));
}
VDEC_HD constexpr uint32_t rotr_bits(uint32_t x, uint32_t p, uint32_t bits) {
    p %= bits;
    return p == 0 ? x : (((x >> p) | (x << (bits - p))) & ((1u << bits) - 1u));
}

}  // namespace vd
