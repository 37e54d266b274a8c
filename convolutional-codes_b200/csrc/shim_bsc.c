/*
 * shim_bsc.c — legacy hard-decision plugin ABI on top of the batched GPU decoder.
 *
 * Exports the six symbols of /root/reference/binary-symmetric-channel/include/decoder.h:17-26 so
 * that binary-symmetric-channel/main.c links against libvdec_bsc.so instead of
 * binary-symmetric-channel/viterbi-decoder.c.  The driver passes a whole
 * frame of T symbol bytes per call (main.c:70); any split is accepted.  The output callback gets
 * the final minimum path metric as well (viterbi-decoder.c:101).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vdec.h"
#include "../../include/vdec_code_param.h"
#include "../../include/vdec_legacy_bsc.h"

struct decoder {
    struct code_param parameter;
    vdec_t* h;
    uint32_t steps, out_bytes;
    uint32_t step_counter;
    uint8_t* frame;                /* [T] received symbols */
    uint8_t* output_data;
    int32_t path_metric;           /* :21 */
    callback_t output_callback;
};

struct decoder* decoder_create(void) {
    return (struct decoder*)calloc(1, sizeof(struct decoder));
}

static void release(struct decoder* d) {
    if (d->h) { vdec_destroy(d->h); d->h = NULL; }
    free(d->frame); d->frame = NULL;
    free(d->output_data); d->output_data = NULL;
}

void decoder_destroy(struct decoder** obj) {
    if (obj == NULL || *obj == NULL) return;
    release(*obj);
    free(*obj);
    *obj = NULL;
}

int decoder_init(struct decoder* obj, struct code_param* param) {
    if (obj == NULL || param == NULL) return -1;
    callback_t keep = obj->output_callback;
    release(obj);
    obj->parameter = *param;
    obj->output_callback = keep;

    struct vdec_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.symlen_out = param->symlen_out;
    cfg.constr_len = param->constr_len;
    cfg.block_len = param->block_len;
    cfg.metric = VDEC_HARD_U16;
    cfg.parity = getenv("VDEC_PARITY_LINEAR") ? VDEC_PARITY_LINEAR : VDEC_PARITY_REF;
    cfg.device = getenv("VDEC_DEVICE") ? atoi(getenv("VDEC_DEVICE")) : 0;
    cfg.max_frames = 1;
    for (uint32_t i = 0; i < cfg.symlen_out && i < VDEC_MAX_POLY; ++i) cfg.polynomials[i] = param->polynomials[i];
    if (vdec_create(&obj->h, &cfg) != 0) return -1;

    obj->steps = vdec_steps(obj->h);
    obj->out_bytes = (uint32_t)vdec_out_bytes_per_frame(obj->h);
    obj->frame = (uint8_t*)malloc(obj->steps);
    obj->output_data = (uint8_t*)malloc(obj->out_bytes);
    if (obj->frame == NULL || obj->output_data == NULL) { release(obj); return -1; }
    decoder_reset(obj);
    return 0;
}

void decoder_register_callback(struct decoder* obj, callback_t output_callback) {
    if (obj) obj->output_callback = output_callback;
}

void decoder_reset(struct decoder* obj) {                    /* :222-232 */
    if (obj == NULL) return;
    obj->step_counter = 0;
    if (obj->output_data) memset(obj->output_data, 0, obj->out_bytes);
}

static int emit_output_buffer(struct decoder* obj) {         /* :98-112 */
    int res = -1;
    if (obj->output_callback != NULL)
        res = obj->output_callback(obj->output_data, obj->parameter.block_len, obj->path_metric,
                                   obj->parameter.userdata);
    if (res == 0) {
        obj->output_data = (uint8_t*)malloc(obj->out_bytes);
        if (obj->output_data == NULL) return -1;
    }
    return 0;
}

int decoder_input(struct decoder* obj, uint8_t* data, uint8_t len) {   /* :234-243 */
    if (obj == NULL || obj->h == NULL || data == NULL) return -1;
    for (uint32_t count = 0; count < len; ++count) {
        obj->frame[obj->step_counter++] = data[count];       /* masking to n bits happens in the kernel */
        if (obj->step_counter == obj->steps) {               /* :141-147 */
            if (vdec_decode_host(obj->h, obj->frame, 1, obj->output_data, &obj->path_metric) != 0) {
                decoder_reset(obj);
                return -1;
            }
            if (emit_output_buffer(obj) != 0) return -1;
            decoder_reset(obj);
        }
    }
    return 0;
}
