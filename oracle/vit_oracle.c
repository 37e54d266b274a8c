/*
 * TEST INFRASTRUCTURE — NOT A PRODUCT PATH.  See vit_oracle.h.
 *
 * Type-widened CPU restatement of the reference Viterbi decoders.  The loop
 * order, comparisons, sentinels, clamp, argmin and bit packing follow the
 * reference statement by statement so that results (all T decoded bits and
 * the hard decoder's path metric) are bit-identical wherever the reference
 * itself can run; only the integer widths differ (uint8_t/uint16_t ->
 * uint32_t/uint64_t), and the reference's undefined shift is written in
 * defined C.
 *
 * Build: gcc -O3 (no -march, no -ffast-math: x86-64 baseline has no FMA, which
 * keeps the demapper arithmetic identical to the reference build,
 * CMakeLists.txt:5).
 */
#include "vit_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- parity ------------------------------------------------------------------
 * Reference: AWGN-channel/viterbi-decoder.c:52-57 (same text in
 * binary-symmetric-channel/viterbi-decoder.c:52-57 and common/encoder.c:95-100):
 *     val ^= val>>32; val ^= val>>16; val ^= val>>8; val ^= (val>>4)&0x0F;
 *     res = (0x6996 >> val) & 1;
 * `0x6996` is an int and `val` still carries its upper bits, so the shift count
 * is out of range (UB).  gcc/x86-64 emits a 32-bit shift whose count is taken
 * mod 32: bits 0..3 = xor-fold of the low byte's nibbles, bit 4 = bit 4 of the
 * byte-fold.  If bit 4 is set the 16-bit constant is shifted out -> 0.
 */
uint32_t orc_parity(uint64_t v, uint32_t mode) {
    v ^= v >> 32;
    v ^= v >> 16;
    v ^= v >> 8;
    uint32_t b = (uint32_t)(v & 0xFFu);               /* xor of the 8 bytes */
    uint32_t nib = (b ^ (b >> 4)) & 0x0Fu;
    uint32_t true_parity = (0x6996u >> nib) & 1u;
    if (mode == ORC_PARITY_LINEAR) return true_parity;
    return (b & 0x10u) ? 0u : true_parity;            /* count = (b&0x10)|nib, mod 32 */
}

/* AWGN-channel/viterbi-decoder.c:45-63 — shift register = state left-aligned
 * below the input bit; P0's parity ends up in the MSB of the n-bit symbol. */
uint32_t orc_expected_symbol(const struct orc_code* c, uint32_t state, uint32_t input) {
    uint64_t sr = ((uint64_t)state) << (64u - c->constr_len);
    sr |= ((uint64_t)input) << 63;
    uint32_t sym = 0;
    for (uint32_t n = 0; n < c->symlen_out; ++n) {
        sym |= orc_parity(sr & c->polynomials[n], c->parity_mode);
        if (n != c->symlen_out - 1) sym <<= 1;
    }
    return sym;
}

void orc_labels(const struct orc_code* c, uint8_t* labels) {
    uint32_t S = 1u << (c->constr_len - 1);
    for (uint32_t s = 0; s < S; ++s)
        for (uint32_t i = 0; i < 2; ++i)
            labels[2 * s + i] = (uint8_t)orc_expected_symbol(c, s, i);
}

/* survivor entry, as the reference's struct state (viterbi-decoder.c:12-15), widened */
struct orc_surv { uint32_t prev; uint8_t input; };

/* ---- soft decision -----------------------------------------------------------
 * AWGN-channel/viterbi-decoder.c: decoder_reset :212-222, receive_symbol :108-141,
 * traceback :71-90.  One call decodes `nframes` independent frames. */
int orc_decode_soft(const struct orc_code* c, const float* in, uint64_t nframes, uint8_t* out) {
    const uint32_t K = c->constr_len, S = 1u << (K - 1), T = c->block_len + K - 1;
    const uint32_t M = 1u << c->symlen_out, OB = (T + 7) / 8;
    struct orc_surv* surv = (struct orc_surv*)calloc((size_t)S * T, sizeof(*surv));   /* :183-187 */
    float* pm = (float*)malloc(sizeof(float) * S);
    float* nm = (float*)malloc(sizeof(float) * S);
    if (!surv || !pm || !nm) { free(surv); free(pm); free(nm); return -1; }

    for (uint64_t f = 0; f < nframes; ++f) {
        const float* rx = in + (size_t)f * T * M;
        uint8_t* o = out + (size_t)f * OB;
        /* decoder_reset :212-222 (the survivor matrix is NOT cleared there) */
        pm[0] = 0.0f;
        for (uint32_t s = 1; s < S; ++s) pm[s] = INFINITY;
        memset(o, 0, OB);

        for (uint32_t t = 0; t < T; ++t) {                      /* receive_symbol */
            const float* m = rx + (size_t)t * M;
            for (uint32_t s = 0; s < S; ++s) nm[s] = INFINITY;  /* :110-112 */
            struct orc_surv* row = surv + (size_t)S * t;        /* :114 */
            for (uint32_t s = 0; s < S; ++s) {                  /* :115 */
                for (uint32_t i = 0; i < 2; ++i) {              /* :116 */
                    uint32_t e = orc_expected_symbol(c, s, i);  /* :45-63 */
                    uint32_t ns = (s >> 1) | (i << (K - 2));    /* :65-66 */
                    float cand = pm[s] + m[e];                  /* :68, :121 */
                    if (cand < nm[ns]) {                        /* :122 strict */
                        nm[ns] = cand;
                        row[ns].prev = s;
                        row[ns].input = (uint8_t)i;
                    }
                }
            }
            memcpy(pm, nm, sizeof(float) * S);                  /* :130 */
        }
        /* traceback :71-90 — argmin with strict '<' from +inf: lowest index wins */
        uint32_t cur = 0;
        float best = INFINITY;
        for (uint32_t s = 0; s < S; ++s)
            if (pm[s] < best) { best = pm[s]; cur = s; }
        for (uint32_t t = T; t-- > 0;) {
            const struct orc_surv* row = surv + (size_t)S * t;
            uint8_t bit = row[cur].input;
            cur = row[cur].prev;
            o[t / 8] |= (uint8_t)(bit << (7 - (t % 8)));        /* :88 */
        }
    }
    free(surv); free(pm); free(nm);
    return 0;
}

/* ---- hard decision -----------------------------------------------------------
 * binary-symmetric-channel/viterbi-decoder.c: reset :222-232, receive_symbol
 * :114-150 (clamp :127-130), Hamming metric :69-72, traceback :75-96. */
int orc_decode_hard(const struct orc_code* c, const uint8_t* in, uint64_t nframes, uint8_t* out,
                    int32_t* path_metric) {
    const uint32_t K = c->constr_len, S = 1u << (K - 1), T = c->block_len + K - 1;
    const uint32_t OB = (T + 7) / 8;
    const uint8_t mask = (uint8_t)((1u << c->symlen_out) - 1);                   /* :235 */
    struct orc_surv* surv = (struct orc_surv*)calloc((size_t)S * T, sizeof(*surv));
    uint16_t* pm = (uint16_t*)malloc(sizeof(uint16_t) * S);
    uint16_t* nm = (uint16_t*)malloc(sizeof(uint16_t) * S);
    if (!surv || !pm || !nm) { free(surv); free(pm); free(nm); return -1; }

    for (uint64_t f = 0; f < nframes; ++f) {
        const uint8_t* rx = in + (size_t)f * T;
        uint8_t* o = out + (size_t)f * OB;
        pm[0] = 0;
        for (uint32_t s = 1; s < S; ++s) pm[s] = 0xFF00;                         /* :225-228 */
        memset(o, 0, OB);

        for (uint32_t t = 0; t < T; ++t) {
            uint8_t r = rx[t] & mask;                                            /* :237 */
            for (uint32_t s = 0; s < S; ++s) nm[s] = 0xFF00;                     /* :116-118 */
            struct orc_surv* row = surv + (size_t)S * t;
            for (uint32_t s = 0; s < S; ++s) {
                for (uint32_t i = 0; i < 2; ++i) {
                    uint8_t x = (uint8_t)(orc_expected_symbol(c, s, i) ^ r);     /* :69 */
                    x = (uint8_t)((x & 0x55) + ((x >> 1) & 0x55));               /* :70-72 */
                    x = (uint8_t)((x & 0x33) + ((x >> 2) & 0x33));
                    x = (uint8_t)((x & 0x0F) + ((x >> 4) & 0x0F));
                    uint32_t ns = (s >> 1) | (i << (K - 2));
                    uint16_t cand = (uint16_t)(pm[s] + x);                       /* :127 */
                    if (cand > 0xFF00) cand = 0xFF00;                            /* :128-130 */
                    if (cand < nm[ns]) {
                        nm[ns] = cand;
                        row[ns].prev = s;
                        row[ns].input = (uint8_t)i;
                    }
                }
            }
            memcpy(pm, nm, sizeof(uint16_t) * S);
        }
        uint32_t cur = 0;
        uint16_t best = 0xFF00;                                                  /* :77 */
        for (uint32_t s = 0; s < S; ++s)
            if (pm[s] < best) { best = pm[s]; cur = s; }
        for (uint32_t t = T; t-- > 0;) {
            const struct orc_surv* row = surv + (size_t)S * t;
            uint8_t bit = row[cur].input;
            cur = row[cur].prev;
            o[t / 8] |= (uint8_t)(bit << (7 - (t % 8)));
        }
        if (path_metric) path_metric[f] = (int32_t)best;                         /* :95, :143, :101 */
    }
    free(surv); free(pm); free(nm);
    return 0;
}

/* ---- encoder -----------------------------------------------------------------
 * common/encoder.c:84-119: bits enter MSB-first at bit 63 of a right-shifting
 * register; after L payload bits K-1 zero bits flush the register. */
void orc_encode(const struct orc_code* c, const uint8_t* payload, uint64_t nframes, uint8_t* symbols) {
    const uint32_t K = c->constr_len, L = c->block_len, T = L + K - 1, PB = (L + 7) / 8;
    for (uint64_t f = 0; f < nframes; ++f) {
        const uint8_t* p = payload + (size_t)f * PB;
        uint8_t* o = symbols + (size_t)f * T;
        uint64_t sr = 0;
        for (uint32_t t = 0; t < T; ++t) {
            uint64_t bit = 0;
            if (t < L) bit = (uint64_t)((p[t / 8] >> (7 - (t % 8))) & 1u);        /* :87 */
            sr = (sr >> 1) | (bit << 63);                                         /* :88-89 */
            uint32_t sym = 0;
            for (uint32_t n = 0; n < c->symlen_out; ++n) {                        /* :92-106 */
                sym |= orc_parity(sr & c->polynomials[n], c->parity_mode);
                if (n < c->symlen_out - 1) sym <<= 1;
            }
            o[t] = (uint8_t)sym;
        }
    }
}

/* ---- constellations / demappers ---------------------------------------------
 * Tables are data (common/constellations.c:8-25), 6-digit decimal literals. */
static const float k_c1[] = { 0.707107f,  0.707107f, -0.707107f, -0.707107f };
static const float k_c2[] = { 0.707107f,  0.707107f,  0.707107f, -0.707107f,
                             -0.707107f,  0.707107f, -0.707107f, -0.707107f };
static const float k_c3[] = { 0.408248f,  0.408248f,  0.408248f,  1.224745f,
                             -0.408248f,  0.408248f, -1.224745f,  0.408248f,
                              0.408248f, -0.408248f,  1.224745f, -0.408248f,
                             -0.408248f, -0.408248f, -0.408248f, -1.224745f };

const float* orc_constellation(uint32_t n) {
    switch (n) { case 1: return k_c1; case 2: return k_c2; case 3: return k_c3; default: return 0; }
}

static float orc_ndist(const float* cst) {            /* common/demapper.c:43-45 */
    float dx = cst[0] - cst[2];
    float dy = cst[1] - cst[3];
    return (dx * dx) + (dy * dy);
}

void orc_demap(uint32_t n, const float* iq, uint64_t nsym, float* metrics) {
    const float* cst = orc_constellation(n);
    const uint32_t M = 1u << n;
    const float nd = orc_ndist(cst);
    for (uint64_t j = 0; j < nsym; ++j) {
        for (uint32_t k = 0; k < M; ++k) {                                        /* demapper.c:66-70 */
            float dx = iq[2 * j] - cst[2 * k];
            float dy = iq[2 * j + 1] - cst[2 * k + 1];
            metrics[j * M + k] = ((dx * dx) + (dy * dy)) / nd;
        }
    }
}

void orc_hard_demap(uint32_t n, const float* iq, uint64_t nsym, float* metrics) {
    const float* cst = orc_constellation(n);
    const uint32_t M = 1u << n;
    const float nd = orc_ndist(cst);
    for (uint64_t j = 0; j < nsym; ++j) {
        float best = INFINITY;                                                    /* hard-demapper.c:68-78 */
        uint32_t bi = 0;
        for (uint32_t k = 0; k < M; ++k) {
            float dx = iq[2 * j] - cst[2 * k];
            float dy = iq[2 * j + 1] - cst[2 * k + 1];
            float d = (dx * dx) + (dy * dy);
            if (d < best) { best = d; bi = k; }
        }
        float sx = cst[2 * bi], sy = cst[2 * bi + 1];                             /* :79-80 */
        for (uint32_t k = 0; k < M; ++k) {                                        /* :83-87 */
            float dx = sx - cst[2 * k];
            float dy = sy - cst[2 * k + 1];
            metrics[j * M + k] = ((dx * dx) + (dy * dy)) / nd;
        }
    }
}
