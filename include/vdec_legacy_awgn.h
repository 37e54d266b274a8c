/*
 * vdec_legacy_awgn.h — the reference's soft-decision decoder plugin ABI, served by the GPU.
 *
 * libvdec_awgn.so exports exactly the six symbols of
 * /root/reference/AWGN-channel/include/decoder.h:17-26 with the same signatures and the
 * same callback contract (decoder.h:9), so AWGN-channel/main.c links against it in place of
 * AWGN-channel/viterbi-decoder.c (CMakeLists.txt:43) without modification.
 *
 * Semantics kept from the reference implementation (AWGN-channel/viterbi-decoder.c):
 *   - decoder_create() returns a zero-filled object (:144-150); decoder_init() may be called
 *     repeatedly and copies *param, reading `polynomials` during the call (:168-206);
 *   - decoder_input() takes `len` floats, a multiple of 2^symlen_out or it returns -1 (:224-236);
 *   - the output callback fires synchronously inside the decoder_input() call that delivers
 *     trellis step T = block_len + constr_len - 1, with (data, block_len, userdata); `data` holds
 *     all T bits MSB-first in a malloc'd buffer of (T+7)/8 bytes; if the callback returns 0 the
 *     callee owns that buffer and a new one is allocated (:92-106); NULL callback == -1;
 *   - decoder_reset() aborts the current frame (:212-222);
 *   - decoder_destroy() frees everything and NULLs the caller's pointer; unlike the reference it
 *     also accepts *obj == NULL (:152-166).
 * Returns: 0 ok, -1 failure (allocation, CUDA error, bad length).
 */
#ifndef VDEC_LEGACY_AWGN_H
#define VDEC_LEGACY_AWGN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int (*callback_t)(uint8_t* data, uint8_t len, void* userdata);   /* decoder.h:9 */

struct code_param;   /* common/include/code.h:9-19 */
struct decoder;      /* opaque, decoder.h:14 */

struct decoder* decoder_create(void);                                     /* decoder.h:17 */
void decoder_destroy(struct decoder** obj);                               /* decoder.h:18 */
int  decoder_init(struct decoder* obj, struct code_param* param);         /* decoder.h:21 */
void decoder_register_callback(struct decoder* obj, callback_t cb);       /* decoder.h:22 */
void decoder_reset(struct decoder* obj);                                  /* decoder.h:23 */
int  decoder_input(struct decoder* obj, float* data, uint8_t len);        /* decoder.h:26 */

#ifdef __cplusplus
}
#endif
#endif
