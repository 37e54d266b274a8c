/*
 * shim_awgn.c — legacy soft-decision plugin ABI on top of the batched GPU decoder.
 *
 * Exports the six symbols of /root/reference/AWGN-channel/include/decoder.h:17-26 so that
 * AWGN-channel/main.c links against libvdec_awgn.so instead of AWGN-channel/viterbi-decoder.c
 * (see INTEGRATION.md).  Host code is plain C; all arithmetic happens in libvdec.so's CUDA
 * kernels — there is no CPU decode path here.
 *
 * The driver delivers one trellis step (2^n floats) per decoder_input() call
 * (AWGN-channel/main.c:108-116) and expects the output callback synchronously inside the call
 * that completes the frame (main.c:124-127,226-228), so this shim buffers T steps and then runs
 * one frame through vdec_decode_host().  It is latency-bound by construction; throughput users
 * call the batched API (include/vdec.h).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vdec.h"
#include "../../include/vdec_code_param.h"
#include "../../include/vdec_legacy_awgn.h"

struct decoder {
    struct code_param parameter;   /* copied at init (viterbi-decoder.c:169) */
    vdec_t* h;
    uint32_t steps, metrics_per_step, out_bytes;
    uint32_t step_counter;         /* symbol_counter of the reference (:21) */
    float* frame;                  /* [T][M] metrics of the frame being received */
    uint8_t* output_data;          /* malloc'd, handed to the callback (:24) */
    callback_t output_callback;
};

struct decoder* decoder_create(void) {                       /* :144-150 */
    return (struct decoder*)calloc(1, sizeof(struct decoder));
}

static void release(struct decoder* d) {
    if (d->h) { vdec_destroy(d->h); d->h = NULL; }
    free(d->frame); d->frame = NULL;
    free(d->output_data); d->output_data = NULL;
}

void decoder_destroy(struct decoder** obj) {                 /* :152-166 */
    if (obj == NULL || *obj == NULL) return;
    release(*obj);
    free(*obj);
    *obj = NULL;
}

int decoder_init(struct decoder* obj, struct code_param* param) {   /* :168-206 */
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
    cfg.metric = VDEC_SOFT_F32;
    cfg.parity = getenv("VDEC_PARITY_LINEAR") ? VDEC_PARITY_LINEAR : VDEC_PARITY_REF;
    cfg.device = getenv("VDEC_DEVICE") ? atoi(getenv("VDEC_DEVICE")) : 0;
    cfg.max_frames = 1;
    for (uint32_t i = 0; i < cfg.symlen_out && i < VDEC_MAX_POLY; ++i) cfg.polynomials[i] = param->polynomials[i];
    if (vdec_create(&obj->h, &cfg) != 0) return -1;

    obj->steps = vdec_steps(obj->h);
    obj->metrics_per_step = 1u << cfg.symlen_out;
    obj->out_bytes = (uint32_t)vdec_out_bytes_per_frame(obj->h);
    obj->frame = (float*)malloc(sizeof(float) * obj->steps * obj->metrics_per_step);
    obj->output_data = (uint8_t*)malloc(obj->out_bytes);
    if (obj->frame == NULL || obj->output_data == NULL) { release(obj); return -1; }
    decoder_reset(obj);
    return 0;
}

void decoder_register_callback(struct decoder* obj, callback_t output_callback) {   /* :208-210 */
    if (obj) obj->output_callback = output_callback;
}

void decoder_reset(struct decoder* obj) {                    /* :212-222 */
    if (obj == NULL) return;
    obj->step_counter = 0;
    if (obj->output_data) memset(obj->output_data, 0, obj->out_bytes);
}

static int emit_output_buffer(struct decoder* obj) {         /* :92-106 */
    int res = -1;
    if (obj->output_callback != NULL)
        res = obj->output_callback(obj->output_data, obj->parameter.block_len, obj->parameter.userdata);
    if (res == 0) {                                          /* callee took ownership of the buffer */
        obj->output_data = (uint8_t*)malloc(obj->out_bytes);
        if (obj->output_data == NULL) return -1;
    }
    return 0;
}

int decoder_input(struct decoder* obj, float* data, uint8_t len) {   /* :224-236 */
    if (obj == NULL || obj->h == NULL || data == NULL) return -1;
    const uint32_t m = obj->metrics_per_step;
    if (len % m != 0) return -1;
    for (uint32_t count = 0; count < len; count += m) {
        memcpy(obj->frame + (size_t)obj->step_counter * m, data + count, sizeof(float) * m);
        obj->step_counter++;
        if (obj->step_counter == obj->steps) {               /* :133-139 */
            if (vdec_decode_host(obj->h, obj->frame, 1, obj->output_data, NULL) != 0) {
                decoder_reset(obj);
                return -1;
            }
            if (emit_output_buffer(obj) != 0) return -1;
            decoder_reset(obj);
        }
    }
    return 0;
}
