/*
 * vdec_sweep_main.c — the reference drivers' operating-point loops bound to the multi-GPU C entry vdec_sweep().
 *
 * Plain C host code (north_star: "host code stays in C").  It plays the role of AWGN-channel/main.c:147-239 and
 * binary-symmetric-channel/main.c:98-182: pick a code, loop over Eb/N0 points (or channel error rates), print the
 * bit error rate per point in the drivers' wording (main.c:233), and append the vectors in the style of
 * results/awgn_channel.m / results/binary_symmetric_channel.m.  All frames of a point are generated, sent through the
 * channel, demapped, decoded and counted on the GPUs of this node; the only collective is the ncclAllReduce of the
 * three counters inside vdec_sweep().
 *
 *   cc -O2 -I include examples/vdec_sweep_main.c -L convolutional-codes_b200/lib -lvdec -lm -o vdec-sweep
 *   ./vdec-sweep --code k7r12 --labels linear --ebn0 0:8:1 --frames 8000000 --min-frame-errors 200 \
 *                --max-frames 100000000 --out results/awgn_k7r12_linear_soft_b200.m
 *   ./vdec-sweep --code cb0 --channel bsc --out results/bsc_cb0_b200.m        (the 17 rates of the BSC driver)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "vdec.h"

struct named_code { const char* name; uint32_t n, K, L; uint64_t p[3]; };
/* common/codebook.c:14-54 (index 4, the K = 32 WSPR code, is not Viterbi-decodable) + the BASELINE codes */
static const struct named_code kCodes[] = {
    {"cb0", 2, 3, 40, {0xA000000000000000ull, 0x6000000000000000ull, 0}},
    {"cb1", 2, 4, 40, {0xB000000000000000ull, 0xE000000000000000ull, 0}},
    {"cb2", 2, 5, 40, {0xA800000000000000ull, 0xF000000000000000ull, 0}},
    {"cb3", 2, 6, 40, {0xB400000000000000ull, 0xE800000000000000ull, 0}},
    {"cb5", 3, 3, 40, {0xA000000000000000ull, 0xC000000000000000ull, 0x2000000000000000ull}},
    {"k7r12", 2, 7, 1024, {0xF200000000000000ull, 0xB600000000000000ull, 0}},
    {"k7r13", 3, 7, 1024, {0xB600000000000000ull, 0xF200000000000000ull, 0xEA00000000000000ull}},
    {"k9", 2, 9, 1024, {0xB880000000000000ull, 0xF580000000000000ull, 0}},
    {"k15", 2, 15, 1024, {0x99A2000000000000ull, 0xA572000000000000ull, 0}},
};
/* binary-symmetric-channel/main.c:103-109 */
static const double kBscRates[] = {1, 5, 25, 125, 625, 3125, 6250, 12500, 15625, 25000, 50000, 78125, 100000, 200000, 300000, 390625, 400000};

static const char* arg(int argc, char** argv, const char* key, const char* dflt) {
    for (int i = 1; i + 1 < argc; ++i) if (strcmp(argv[i], key) == 0) return argv[i + 1];
    return dflt;
}

static int usage(const char* prog) {
    fprintf(stderr,
            "usage: %s [--code cb0|cb1|cb2|cb3|cb5|k7r12|k7r13|k9|k15] [--labels ref|linear] [--channel awgn|awgn-hard|bsc]\n"
            "          [--ebn0 first:last:step] [--frames N per batch] [--min-frame-errors N] [--max-frames N] [--block-len L]\n"
            "          [--gpus N (0 = all)] [--seed S] [--chunk frames per launch] [--resume file] [--out file.m]\n"
            "Eb/N0 (or, with --channel bsc, the BSC driver's 17 crossover rates) sweep on the GPUs of this node; prints one line per\n"
            "operating point and appends result vectors in the style of the reference's results/*.m to --out.\n", prog);
    return 2;
}

int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i)
        if (strcmp(argv[i], "--help") == 0 || strcmp(argv[i], "-h") == 0) return usage(argv[0]);
    const char* cname = arg(argc, argv, "--code", "k7r12");
    const char* chan = arg(argc, argv, "--channel", "awgn");            /* awgn | awgn-hard | bsc */
    const char* labels = arg(argc, argv, "--labels", "ref");            /* ref = the reference's de-facto labels | linear */
    const char* out = arg(argc, argv, "--out", "");
    const struct named_code* code = NULL;
    for (size_t i = 0; i < sizeof(kCodes) / sizeof(kCodes[0]); ++i) if (strcmp(kCodes[i].name, cname) == 0) code = &kCodes[i];
    if (!code) { fprintf(stderr, "unknown code %s\n", cname); return 2; }

    struct vdec_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.symlen_out = code->n; cfg.constr_len = code->K;
    cfg.block_len = (uint32_t)atol(arg(argc, argv, "--block-len", "0"));
    if (cfg.block_len == 0) cfg.block_len = code->L;
    cfg.parity = strcmp(labels, "linear") == 0 ? VDEC_PARITY_LINEAR : VDEC_PARITY_REF;
    for (uint32_t j = 0; j < code->n; ++j) cfg.polynomials[j] = code->p[j];

    struct vdec_sweep_config sw;
    memset(&sw, 0, sizeof(sw));
    sw.channel = strcmp(chan, "bsc") == 0 ? VDEC_CHANNEL_BSC : strcmp(chan, "awgn-hard") == 0 ? VDEC_CHANNEL_AWGN_HARD : VDEC_CHANNEL_AWGN;
    cfg.metric = sw.channel == VDEC_CHANNEL_BSC ? VDEC_HARD_U16 : VDEC_SOFT_F32;
    sw.seed = strtoull(arg(argc, argv, "--seed", "12345"), NULL, 10);
    sw.frames_per_batch = strtoull(arg(argc, argv, "--frames", "1000000"), NULL, 10);
    sw.min_frame_errors = strtoull(arg(argc, argv, "--min-frame-errors", "0"), NULL, 10);
    sw.max_frames = strtoull(arg(argc, argv, "--max-frames", "0"), NULL, 10);
    sw.chunk_frames = (uint32_t)atol(arg(argc, argv, "--chunk", "0"));
    const char* resume = arg(argc, argv, "--resume", "");
    sw.resume_path = *resume ? resume : NULL;
    int ndev = atoi(arg(argc, argv, "--gpus", "0"));                     /* 0 = every device of the node */

    double points[64];
    uint32_t np = 0;
    if (sw.channel == VDEC_CHANNEL_BSC) {
        for (; np < sizeof(kBscRates) / sizeof(kBscRates[0]); ++np) points[np] = kBscRates[np];
    } else {                                                             /* main.c:150-152: 0, 2, .. 14 dB by default */
        double a = 0, b = 14, st = 2;
        sscanf(arg(argc, argv, "--ebn0", "0:14:2"), "%lf:%lf:%lf", &a, &b, &st);
        for (double x = a; x <= b + 1e-9 && np < 64; x += st) points[np++] = x;
    }

    const int present = vdec_device_count();
    if (ndev <= 0 || ndev > present) ndev = present;
    uint64_t counters[64 * 3];
    {   /* CUDA context, kernel images and NCCL communicators come up once: keep that out of the timed sweep */
        struct vdec_sweep_config warm = sw;
        warm.frames_per_batch = 1024; warm.min_frame_errors = 0; warm.resume_path = NULL;
        if (vdec_sweep(&cfg, &warm, points, 1, NULL, ndev, counters) != 0) { fprintf(stderr, "vdec_sweep failed: %s\n", vdec_last_error()); return 1; }
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (vdec_sweep(&cfg, &sw, points, np, NULL, ndev, counters) != 0) {
        fprintf(stderr, "vdec_sweep failed: %s\n", vdec_last_error());
        return 1;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    double total_bits = 0;
    for (uint32_t i = 0; i < np; ++i) {
        const uint64_t* c = counters + 3 * i;
        total_bits += (double)c[0];
        printf("====================================================\n");
        if (sw.channel == VDEC_CHANNEL_BSC) printf("Channel error rate: %lf\n", points[i] / 1e6);
        else printf("Eb/N0: %4.1lf dB\n", points[i]);
        printf("bits %llu  bit errors %llu  frame errors %llu\n", (unsigned long long)c[0], (unsigned long long)c[1], (unsigned long long)c[2]);
        printf("Overall transmission error rate: %.10lf\n", c[0] ? (double)c[1] / (double)c[0] : 0.0);     /* main.c:233 */
    }
    printf("====================================================\n");
    printf("%u points, %.3e bits on %d GPU(s) in %.2f s (%.1f Gbit/s incl. source, channel, demapper, decoder, error count, per-call buffer set-up)\n",
           np, total_bits, ndev, secs, total_bits / secs / 1e9);
    if (*out) {
        FILE* f = fopen(out, "a");
        if (!f) { perror(out); return 1; }
        char tag[96];
        snprintf(tag, sizeof(tag), "%s_%s_%s", cname, labels, sw.channel == VDEC_CHANNEL_BSC ? "bsc" : sw.channel == VDEC_CHANNEL_AWGN_HARD ? "hard" : "soft");
        fprintf(f, "%% %s, block length %u, %s labels, %d B200 GPU(s), %.2f s, vdec_sweep (ncclAllReduce of the error counters per batch)\n",
                cname, cfg.block_len, labels, ndev, secs);
        fprintf(f, "n_gpus_%s = %d;\n", tag, ndev);
        fprintf(f, "%s_%s = [", sw.channel == VDEC_CHANNEL_BSC ? "ppm" : "snr", tag);
        for (uint32_t i = 0; i < np; ++i) fprintf(f, "%s%g", i ? " " : "", points[i]);
        fprintf(f, "];\nbits_%s = [", tag);
        for (uint32_t i = 0; i < np; ++i) fprintf(f, "%s%llu", i ? " " : "", (unsigned long long)counters[3 * i]);
        fprintf(f, "];\nframe_errors_%s = [", tag);
        for (uint32_t i = 0; i < np; ++i) fprintf(f, "%s%llu", i ? " " : "", (unsigned long long)counters[3 * i + 2]);
        fprintf(f, "];\nber_%s = [", tag);
        for (uint32_t i = 0; i < np; ++i) fprintf(f, "%s%.6e", i ? " " : "", counters[3 * i] ? (double)counters[3 * i + 1] / (double)counters[3 * i] : 0.0);
        fprintf(f, "];\n");
        fclose(f);
    }
    return 0;
}
