/*
 * TEST INFRASTRUCTURE — NOT A PRODUCT PATH.
 *
 * CPU oracle for the Viterbi hot path of thomas-emig/convolutional-codes:
 * a type-widened restatement of
 *     AWGN-channel/viterbi-decoder.c              (soft decision, float metrics)
 *     binary-symmetric-channel/viterbi-decoder.c  (hard decision, uint16 metrics)
 * plus the two producers/consumers either side of it that tests need to build
 * inputs (common/encoder.c, common/demapper.c, common/hard-demapper.c).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  The CUDA library never links or calls it.
 *
 * Parity status: PINNED — tests/test_oracle.py checks this file
 * bit-for-bit against the unmodified reference objects (oracle/_ref, built by
 * oracle/Makefile from /root/reference) on every codebook code the reference
 * can run (K <= 8, L + K - 1 <= 255), and tests/golden/ holds vectors dumped
 * from the compiled reference with fixed seeds (oracle/gen_golden.py).
 */
#ifndef VIT_ORACLE_H
#define VIT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_POLY 8

enum { ORC_PARITY_REF = 0,     /* de-facto labelling of the reference's 0x6996 shift expression */
       ORC_PARITY_LINEAR = 1   /* true GF(2) parity (textbook code) */ };

struct orc_code {
    uint32_t symlen_out;               /* n: coded bits per trellis step            (code.h:10) */
    uint32_t constr_len;               /* K: constraint length                      (code.h:11) */
    uint32_t block_len;                /* L: information bits per frame, widened    (code.h:12) */
    uint32_t parity_mode;              /* ORC_PARITY_*                                        */
    uint64_t polynomials[ORC_MAX_POLY];/* left-aligned, bit 63 = newest bit (codebook.c:14-17) */
};

/* parity of (shift_register & polynomial) as the reference computes it */
uint32_t orc_parity(uint64_t masked, uint32_t mode);
/* expected n-bit symbol for (state, input): viterbi-decoder.c:45-63 */
uint32_t orc_expected_symbol(const struct orc_code* c, uint32_t state, uint32_t input);
/* labels[2*s + i] for all S states */
void orc_labels(const struct orc_code* c, uint8_t* labels);

/* soft decoder: in [F][T][M] float, out [F][(T+7)/8] bytes, MSB first, tail bits included */
int orc_decode_soft(const struct orc_code* c, const float* in, uint64_t nframes, uint8_t* out);
/* hard decoder: in [F][T] one symbol byte per step, path_metric[F] = final min metric */
int orc_decode_hard(const struct orc_code* c, const uint8_t* in, uint64_t nframes, uint8_t* out,
                    int32_t* path_metric);

/* encoder: payload [F][(L+7)/8] bytes MSB first -> symbols [F][T], zero tail of K-1 bits */
void orc_encode(const struct orc_code* c, const uint8_t* payload, uint64_t nframes, uint8_t* symbols);
/* constellation table of common/constellations.c for n = 1..3 (2*2^n floats) */
const float* orc_constellation(uint32_t n);
/* soft demapper: iq [N][2] -> metrics [N][M]   (common/demapper.c:43-45,66-70) */
void orc_demap(uint32_t n, const float* iq, uint64_t nsym, float* metrics);
/* hard demapper: snaps iq to the nearest point first (common/hard-demapper.c:66-87) */
void orc_hard_demap(uint32_t n, const float* iq, uint64_t nsym, float* metrics);

#ifdef __cplusplus
}
#endif
#endif
