/*
 * TEST INFRASTRUCTURE — NOT A PRODUCT PATH.
 *
 * Thin driver around the UNMODIFIED reference objects.  oracle/Makefile
 * compiles this file together with the reference sources where they lie under
 * /root/reference (nothing is copied into this repo) into the oracle/_ref/ libraries:
 *
 *   libref_awgn.so       REFH_AWGN  + AWGN-channel/viterbi-decoder.c, common/{encoder,mapper,
 *                                     demapper,gaussian,constellations,codebook}.c
 *   libref_awgn_hard.so  same, with common/hard-demapper.c instead of demapper.c
 *   libref_bsc.so        REFH_BSC   + binary-symmetric-channel/viterbi-decoder.c,
 *                                     common/{encoder,codebook}.c
 *
 * The functions below wire the objects exactly as the reference drivers do
 * (AWGN-channel/main.c:80-144,172-190,217-223; binary-symmetric-channel/main.c:57-98,
 * 115-132,162-166) but with a fixed srand() seed and with every decoder input
 * and output captured, so that golden vectors and differential tests can be
 * produced from the real reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "code.h"
#include "codebook.h"
#include "decoder.h"
#include "encoder.h"

#if defined(REFH_AWGN)
#include "demapper.h"
#include "gaussian.h"
#include "mapper.h"
#endif

struct refh_code {            /* mirrors what the python side passes in */
    uint32_t symlen_out, constr_len, block_len;
    uint64_t polynomials[8];
};

struct refh_ctx {
    struct code_param param;
    uint64_t polys[8];
    struct encoder* enc;
    struct decoder* dec;
#if defined(REFH_AWGN)
    struct mapper* map;
    struct demapper* demap;
    float scaling;
    float* in_dump;           /* [F][T][M] decoder inputs */
#else
    int crossover_ppm;        /* binary-symmetric-channel/main.c:64 */
    uint8_t* in_dump;         /* [F][T] decoder inputs (after bit flips) */
    int32_t* pm_dump;
#endif
    uint8_t* out_dump;        /* [F][(T+7)/8] */
    uint64_t in_pos, frame;
    uint32_t T, M, OB;
    int callbacks;
};

static void fill_param(struct refh_ctx* x, const struct refh_code* c) {
    memset(&x->param, 0, sizeof(x->param));
    memcpy(x->polys, c->polynomials, sizeof(x->polys));
    x->param.symlen_out = (uint8_t)c->symlen_out;
    x->param.constr_len = (uint8_t)c->constr_len;
    x->param.block_len = (uint8_t)c->block_len;
    x->param.polynomials = x->polys;
    x->param.userdata = x;
    x->T = c->block_len + c->constr_len - 1;
    x->M = 1u << c->symlen_out;
    x->OB = (x->T + 7) / 8;
}

/* the reference can only run these (uint8_t lengths, uint8_t state loop) */
int refh_supported(const struct refh_code* c) {
    return c->constr_len >= 2 && c->constr_len <= 8 && c->block_len + c->constr_len - 1 <= 255 &&
           c->symlen_out >= 1 && c->symlen_out <= 3;
}

void refh_get_code(uint32_t index, struct refh_code* out) {       /* common/codebook.c:110-120 */
    struct code_param p;
    get_code((uint8_t)index, &p);
    memset(out, 0, sizeof(*out));
    out->symlen_out = p.symlen_out; out->constr_len = p.constr_len; out->block_len = p.block_len;
    for (uint32_t i = 0; i < p.symlen_out; ++i) out->polynomials[i] = p.polynomials[i];
}

#if defined(REFH_AWGN)

static int cb_decoded(uint8_t* data, uint8_t len, void* ud) {
    struct refh_ctx* x = (struct refh_ctx*)ud;
    (void)len;
    memcpy(x->out_dump + x->frame * x->OB, data, x->OB);   /* all T bits incl. tail */
    x->frame++;
    x->callbacks++;
    return -1;
}
static int cb_demapped(float* data, uint8_t len, void* ud) {     /* main.c:108-116 */
    struct refh_ctx* x = (struct refh_ctx*)ud;
    if (x->in_dump) memcpy(x->in_dump + x->in_pos, data, len * sizeof(float));
    x->in_pos += len;
    (void)decoder_input(x->dec, data, len);
    return -1;
}
static int cb_mapped(float* data, uint8_t len, void* ud) {       /* main.c:93-106 */
    struct refh_ctx* x = (struct refh_ctx*)ud;
    for (uint8_t i = 0; i < len; ++i) data[i] += x->scaling * gengauss();
    (void)demapper_input(x->demap, data, len);
    return -1;
}
static int cb_encoded(uint8_t* data, uint8_t len, void* ud) {    /* main.c:80-91 */
    struct refh_ctx* x = (struct refh_ctx*)ud;
    (void)mapper_input(x->map, data, len);
    return -1;
}

/* Full reference link for `nframes` frames: payload rand()%256 -> encoder -> mapper -> +noise
 * -> demapper -> decoder.  Dumps payload [F][(L+7)/8], decoder inputs [F][T][M], outputs. */
int refh_awgn_run(const struct refh_code* c, uint64_t nframes, uint32_t seed, float scaling,
                  uint8_t* payload_dump, float* in_dump, uint8_t* out_dump) {
    if (!refh_supported(c)) return -2;
    struct refh_ctx x; memset(&x, 0, sizeof(x));
    fill_param(&x, c);
    x.scaling = scaling; x.in_dump = in_dump; x.out_dump = out_dump;
    x.enc = encoder_create(); x.map = mapper_create(); x.demap = demapper_create(); x.dec = decoder_create();
    if (!x.enc || !x.map || !x.demap || !x.dec) return -1;
    (void)encoder_init(x.enc, &x.param);   encoder_register_callback(x.enc, cb_encoded);
    (void)mapper_init(x.map, &x.param);    mapper_register_callback(x.map, cb_mapped);
    (void)demapper_init(x.demap, &x.param); demapper_register_callback(x.demap, cb_demapped);
    (void)decoder_init(x.dec, &x.param);   decoder_register_callback(x.dec, cb_decoded);
    srand(seed);
    const uint32_t PB = (c->block_len + 7) / 8;
    uint8_t data[64];
    for (uint64_t f = 0; f < nframes; ++f) {
        for (uint32_t j = 0; j < PB; ++j) data[j] = (uint8_t)(rand() % 256);       /* main.c:217-219 */
        if (payload_dump) memcpy(payload_dump + f * PB, data, PB);
        (void)encoder_input(x.enc, data, (uint8_t)c->block_len);                   /* main.c:223 */
    }
    int ok = (x.frame == nframes) ? 0 : -3;
    encoder_destroy(&x.enc); mapper_destroy(&x.map); demapper_destroy(&x.demap); decoder_destroy(&x.dec);
    return ok;
}

/* Decoder only, on caller-supplied metrics, M floats per call as the driver does (main.c:114). */
int refh_awgn_decode(const struct refh_code* c, const float* in, uint64_t nframes, uint8_t* out_dump) {
    if (!refh_supported(c)) return -2;
    struct refh_ctx x; memset(&x, 0, sizeof(x));
    fill_param(&x, c);
    x.out_dump = out_dump;
    x.dec = decoder_create();
    if (!x.dec) return -1;
    if (decoder_init(x.dec, &x.param) != 0) return -1;
    decoder_register_callback(x.dec, cb_decoded);
    float buf[16];
    for (uint64_t f = 0; f < nframes; ++f)
        for (uint32_t t = 0; t < x.T; ++t) {
            memcpy(buf, in + (f * x.T + t) * x.M, x.M * sizeof(float));
            if (decoder_input(x.dec, buf, (uint8_t)x.M) != 0) return -4;
        }
    int ok = (x.frame == nframes) ? 0 : -3;
    decoder_destroy(&x.dec);
    return ok;
}

/* demapper only: iq [N][2] -> metrics [N][M]  (whichever of demapper.c / hard-demapper.c is linked) */
static int cb_demap_only(float* data, uint8_t len, void* ud) {
    struct refh_ctx* x = (struct refh_ctx*)ud;
    memcpy(x->in_dump + x->in_pos, data, len * sizeof(float));
    x->in_pos += len;
    return -1;
}
int refh_demap(uint32_t n, const float* iq, uint64_t nsym, float* metrics) {
    struct refh_ctx x; memset(&x, 0, sizeof(x));
    struct refh_code c; memset(&c, 0, sizeof(c));
    c.symlen_out = n; c.constr_len = 3; c.block_len = 8;
    fill_param(&x, &c);
    x.in_dump = metrics;
    x.demap = demapper_create();
    if (!x.demap) return -1;
    (void)demapper_init(x.demap, &x.param);
    demapper_register_callback(x.demap, cb_demap_only);
    float xy[2];
    for (uint64_t j = 0; j < nsym; ++j) {
        xy[0] = iq[2 * j]; xy[1] = iq[2 * j + 1];
        (void)demapper_input(x.demap, xy, 2);
    }
    demapper_destroy(&x.demap);
    return 0;
}

#elif defined(REFH_BSC)

static int cb_decoded(uint8_t* data, uint8_t len, int32_t path_metric, void* ud) {   /* main.c:74-98 */
    struct refh_ctx* x = (struct refh_ctx*)ud;
    (void)len;
    memcpy(x->out_dump + x->frame * x->OB, data, x->OB);
    if (x->pm_dump) x->pm_dump[x->frame] = path_metric;
    x->frame++;
    return -1;
}
static int cb_encoded(uint8_t* data, uint8_t len, void* ud) {                        /* main.c:57-72 */
    struct refh_ctx* x = (struct refh_ctx*)ud;
    for (uint8_t i = 0; i < len; ++i)
        for (uint8_t k = 0; k < x->param.symlen_out; ++k)
            if ((rand() % 1000000) < x->crossover_ppm) data[i] ^= (uint8_t)(1 << k);
    if (x->in_dump) memcpy(x->in_dump + x->in_pos, data, len);
    x->in_pos += len;
    (void)decoder_input(x->dec, data, len);
    return -1;
}

int refh_bsc_run(const struct refh_code* c, uint64_t nframes, uint32_t seed, int crossover_ppm,
                 uint8_t* payload_dump, uint8_t* in_dump, uint8_t* out_dump, int32_t* pm_dump) {
    if (!refh_supported(c)) return -2;
    struct refh_ctx x; memset(&x, 0, sizeof(x));
    fill_param(&x, c);
    x.crossover_ppm = crossover_ppm; x.in_dump = in_dump; x.out_dump = out_dump; x.pm_dump = pm_dump;
    x.enc = encoder_create(); x.dec = decoder_create();
    if (!x.enc || !x.dec) return -1;
    (void)encoder_init(x.enc, &x.param); encoder_register_callback(x.enc, cb_encoded);
    (void)decoder_init(x.dec, &x.param); decoder_register_callback(x.dec, cb_decoded);
    srand(seed);
    const uint32_t PB = (c->block_len + 7) / 8;
    uint8_t data[64];
    for (uint64_t f = 0; f < nframes; ++f) {
        for (uint32_t j = 0; j < PB; ++j) data[j] = (uint8_t)(rand() % 256);          /* main.c:162-164 */
        if (payload_dump) memcpy(payload_dump + f * PB, data, PB);
        (void)encoder_input(x.enc, data, (uint8_t)c->block_len);
    }
    int ok = (x.frame == nframes) ? 0 : -3;
    encoder_destroy(&x.enc); decoder_destroy(&x.dec);
    return ok;
}

/* Decoder only: whole frame in one call as the driver does (main.c:70). */
int refh_bsc_decode(const struct refh_code* c, const uint8_t* in, uint64_t nframes, uint8_t* out_dump,
                    int32_t* pm_dump) {
    if (!refh_supported(c)) return -2;
    struct refh_ctx x; memset(&x, 0, sizeof(x));
    fill_param(&x, c);
    x.out_dump = out_dump; x.pm_dump = pm_dump;
    x.dec = decoder_create();
    if (!x.dec) return -1;
    if (decoder_init(x.dec, &x.param) != 0) return -1;
    decoder_register_callback(x.dec, cb_decoded);
    uint8_t buf[256];
    for (uint64_t f = 0; f < nframes; ++f) {
        memcpy(buf, in + f * x.T, x.T);
        if (decoder_input(x.dec, buf, (uint8_t)x.T) != 0) return -4;
    }
    int ok = (x.frame == nframes) ? 0 : -3;
    decoder_destroy(&x.dec);
    return ok;
}

#else
#error "define REFH_AWGN or REFH_BSC"
#endif

/* encoder only (both variants link common/encoder.c): payload -> symbols [F][T] */
static uint8_t* g_sym_dump; static uint64_t g_sym_pos;
static int cb_enc_only(uint8_t* data, uint8_t len, void* ud) {
    (void)ud; memcpy(g_sym_dump + g_sym_pos, data, len); g_sym_pos += len; return -1;
}
int refh_encode(const struct refh_code* c, const uint8_t* payload, uint64_t nframes, uint8_t* symbols) {
    if (!refh_supported(c)) return -2;
    struct refh_ctx x; memset(&x, 0, sizeof(x));
    fill_param(&x, c);
    x.enc = encoder_create();
    if (!x.enc) return -1;
    (void)encoder_init(x.enc, &x.param);
    encoder_register_callback(x.enc, cb_enc_only);
    g_sym_dump = symbols; g_sym_pos = 0;
    const uint32_t PB = (c->block_len + 7) / 8;
    uint8_t data[64];
    for (uint64_t f = 0; f < nframes; ++f) {
        memcpy(data, payload + f * PB, PB);
        (void)encoder_input(x.enc, data, (uint8_t)c->block_len);
    }
    encoder_destroy(&x.enc);
    return 0;
}
