/*
 * vdec_code_param.h — layout of the reference's `struct code_param`
 * (/root/reference/common/include/code.h:9-19), restated so the legacy shims can read the
 * caller's parameters without the reference tree on the include path.  Field order, types and
 * therefore offsets must match the reference exactly: the struct crosses the plugin boundary
 * by pointer (decoder.h:21) and is owned by the caller.
 */
#ifndef VDEC_CODE_PARAM_H
#define VDEC_CODE_PARAM_H

#include <stdint.h>

struct code_param {
    uint8_t   symlen_out;          /* n  */
    uint8_t   constr_len;          /* K  */
    uint8_t   block_len;           /* L  */
    uint64_t* polynomials;         /* n left-aligned generator polynomials, borrowed */
    int32_t*  bit_metrics;         /* stack decoder only — unused here */
    int32_t*  fano_bit_metrics;    /* Fano decoder only — unused here */
    float     metric_weight;       /* stack decoder only — unused here */
    float     fano_metric_weight;  /* Fano decoder only — unused here */
    void*     userdata;            /* passed back to the output callback */
};

#endif
