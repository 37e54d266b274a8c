/*
 * vdec_legacy_bsc.h — the reference's hard-decision decoder plugin ABI, served by the GPU.
 *
 * libvdec_bsc.so exports exactly the six symbols of
 * /root/reference/binary-symmetric-channel/include/decoder.h:17-26 (same names as the AWGN
 * variant, different signatures — hence a separate library), so
 * binary-symmetric-channel/main.c links against it in place of
 * binary-symmetric-channel/viterbi-decoder.c (CMakeLists.txt:21) without modification.
 *
 * Semantics kept from binary-symmetric-channel/viterbi-decoder.c:
 *   - decoder_input() takes `len` symbols, one byte each, masked to the low symlen_out bits
 *     (:234-243) — the header's "bits" is a misnomer; any number of symbols per call;
 *   - the callback receives (data, block_len, path_metric, userdata) where path_metric is the
 *     minimum final uint16 state metric (:75-96,:143,:101); buffer ownership as in the AWGN
 *     variant (:98-112);
 *   - metrics are uint16 clamped at 0xFF00 (:127-130), unreachable = 0xFF00 (:225-228).
 * Returns: 0 ok, -1 failure.
 */
#ifndef VDEC_LEGACY_BSC_H
#define VDEC_LEGACY_BSC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int (*callback_t)(uint8_t* data, uint8_t len, int32_t path_metric, void* userdata);   /* decoder.h:9 */

struct code_param;
struct decoder;

struct decoder* decoder_create(void);                                     /* decoder.h:17 */
void decoder_destroy(struct decoder** obj);                               /* decoder.h:18 */
int  decoder_init(struct decoder* obj, struct code_param* param);         /* decoder.h:21 */
void decoder_register_callback(struct decoder* obj, callback_t cb);       /* decoder.h:22 */
void decoder_reset(struct decoder* obj);                                  /* decoder.h:23 */
int  decoder_input(struct decoder* obj, uint8_t* data, uint8_t len);      /* decoder.h:26 */

#ifdef __cplusplus
}
#endif
#endif
