/*
 * TEST INFRASTRUCTURE — NOT A PRODUCT PATH.
 *
 * LD_PRELOAD shim for the reference's UNMODIFIED drivers: they seed libc rand() with srand(time(NULL))
 * (AWGN-channel/main.c:147, binary-symmetric-channel/main.c:101), which makes their printed BER an unseeded
 * Monte-Carlo estimate.  With this object preloaded time() returns $VDEC_FIXED_TIME (default 20260101), so a driver
 * run is reproducible and tests/test_gpu_legacy.py asserts on a deterministic number.
 */
#include <stdlib.h>
#include <time.h>

time_t time(time_t* out) {
    const char* e = getenv("VDEC_FIXED_TIME");
    const time_t t = (time_t)(e && *e ? atoll(e) : 20260101LL);
    if (out) *out = t;
    return t;
}
