// vdec_synth.cu — the stages either side of the decoder, GPU-resident (SURVEY.md §8f "next"):
//   * synthetic link: payload -> encoder -> mapper -> AWGN -> soft demapper   (AWGN-channel/main.c:80-116)
//                     payload -> encoder -> bit flips                           (binary-symmetric-channel/main.c:57-72)
//   * soft demapper on caller-supplied I/Q                                      (common/demapper.c:43-45,66-70)
//   * bit / frame error counting of the drivers' output callbacks              (AWGN-channel/main.c:129-141)
// Random numbers are Philox4x32-10 keyed by `seed` with counter (global frame, step, stream), so a
// frame's data does not depend on how frames are sharded over launches, streams or GPUs.  The
// reference draws from libc rand() with a time seed (main.c:147), so there is no bit-level parity
// target for the noise itself; the deterministic parts (encoder labels, mapper points, demapper
// arithmetic) are bit-exact and tested against the oracle.
#include "vdec_internal.h"

namespace vd {

namespace {

// constellation points: DATA of common/constellations.c:8-25 (unit power, Gray mapped)
__constant__ float c_cst1[4] = {0.707107f, 0.707107f, -0.707107f, -0.707107f};
__constant__ float c_cst2[8] = {0.707107f, 0.707107f, 0.707107f, -0.707107f, -0.707107f, 0.707107f, -0.707107f, -0.707107f};
__constant__ float c_cst3[16] = {0.408248f, 0.408248f, 0.408248f, 1.224745f, -0.408248f, 0.408248f, -1.224745f, 0.408248f,
                                 0.408248f, -0.408248f, 1.224745f, -0.408248f, -0.408248f, -0.408248f, -0.408248f, -1.224745f};

__device__ __forceinline__ const float* constellation(uint32_t n) { return n == 1 ? c_cst1 : (n == 2 ? c_cst2 : c_cst3); }

__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t (&o)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// demapper.c:66-70 with demapper.c:43-45 — explicit round-to-nearest ops, no FMA contraction
__device__ __forceinline__ float ndist_of(const float* c) {
    const float dx = __fsub_rn(c[0], c[2]), dy = __fsub_rn(c[1], c[3]);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
__device__ __forceinline__ float demap_one(const float* c, uint32_t k, float x, float y, float nd) {
    const float dx = __fsub_rn(x, c[2 * k]), dy = __fsub_rn(y, c[2 * k + 1]);
    return __fdiv_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), nd);
}

// payload byte j of global frame F
__device__ __forceinline__ uint32_t payload_word(uint64_t seed, uint64_t F, uint32_t w4) {
    uint32_t o[4];
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)F, (uint32_t)(F >> 32), w4, 0x50415931u, o);
    return o[0];
}

__global__ void payload_kernel(uint64_t nframes, uint64_t first, uint64_t seed, uint32_t pb, uint8_t* __restrict__ payload) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint32_t words = (pb + 3) / 4;
    if (i >= nframes * words) return;
    const uint64_t f = i / words; const uint32_t w = (uint32_t)(i % words);
    const uint32_t v = payload_word(seed, first + f, w);
    for (uint32_t b = 0; b < 4; ++b) {
        const uint32_t j = 4 * w + b;
        if (j < pb) payload[f * pb + j] = (uint8_t)(v >> (8 * b));
    }
}

// encoder state before step t and input bit at step t, from the payload (encoder.c:84-119:
// bits enter MSB-first; steps >= L are the zero tail)
__device__ __forceinline__ uint32_t enc_symbol(const uint8_t* __restrict__ pl, uint32_t t, uint32_t K, uint32_t L,
                                               const uint8_t* __restrict__ labels) {
    uint32_t state = 0;                                  // bit K-2 = previous input ... (viterbi-decoder.c:65-66)
    for (uint32_t j = 1; j < K; ++j) {
        uint32_t bit = 0;
        if (t >= j && (t - j) < L) bit = (pl[(t - j) >> 3] >> (7 - ((t - j) & 7))) & 1u;
        state |= bit << (K - 1 - j);
    }
    uint32_t in = 0;
    if (t < L) in = (pl[t >> 3] >> (7 - (t & 7))) & 1u;
    return labels[2 * state + in];
}

__global__ void synth_awgn_kernel(LaunchGeom g, const uint8_t* __restrict__ labels, uint64_t nframes, uint64_t first,
                                  uint64_t seed, float sigma, const uint8_t* __restrict__ payload,
                                  float* __restrict__ metrics, float* __restrict__ iq) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= nframes * g.T) return;
    const uint64_t f = i / g.T; const uint32_t t = (uint32_t)(i % g.T);
    const uint32_t pb = (g.L + 7) / 8;
    const uint32_t sym = enc_symbol(payload + f * pb, t, g.K, g.L, labels);
    const float* c = constellation(g.n);
    uint32_t o[4];
    const uint64_t F = first + f;
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)F, (uint32_t)(F >> 32), t, 0x4E4F4953u, o);
    // Box-Muller: one pair per symbol (I, Q)   (gaussian.c:18-38 uses the same transform on rand())
    const float u1 = ((float)(o[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);     // (0, 1]
    const float u2 = (float)(o[1] >> 8) * (1.0f / 16777216.0f);              // [0, 1)
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    const float x = __fadd_rn(c[2 * sym], __fmul_rn(sigma, rad * cs));       // main.c:100-102
    const float y = __fadd_rn(c[2 * sym + 1], __fmul_rn(sigma, rad * sn));
    if (iq != nullptr) { iq[2 * i] = x; iq[2 * i + 1] = y; }
    const float nd = ndist_of(c);
    float* dst = metrics + i * g.M;
    if (g.M == 4) {
        float4 v; v.x = demap_one(c, 0, x, y, nd); v.y = demap_one(c, 1, x, y, nd);
        v.z = demap_one(c, 2, x, y, nd); v.w = demap_one(c, 3, x, y, nd);
        *reinterpret_cast<float4*>(dst) = v;
    } else {
        for (uint32_t k = 0; k < g.M; ++k) dst[k] = demap_one(c, k, x, y, nd);
    }
}

__global__ void synth_bsc_kernel(LaunchGeom g, const uint8_t* __restrict__ labels, uint64_t nframes, uint64_t first,
                                 uint64_t seed, uint32_t ppm, const uint8_t* __restrict__ payload, uint8_t* __restrict__ symbols) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= nframes * g.T) return;
    const uint64_t f = i / g.T; const uint32_t t = (uint32_t)(i % g.T);
    const uint32_t pb = (g.L + 7) / 8;
    uint32_t sym = enc_symbol(payload + f * pb, t, g.K, g.L, labels);
    uint32_t o[4];
    const uint64_t F = first + f;
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)F, (uint32_t)(F >> 32), t, 0x42534331u, o);
    for (uint32_t k = 0; k < g.n; ++k)                                       // bsc main.c:62-68
        if ((uint32_t)(((uint64_t)o[k] * 1000000ull) >> 32) < ppm) sym ^= (1u << k);
    symbols[i] = (uint8_t)sym;
}

__global__ void demap_kernel(uint32_t n, const float* __restrict__ iq, uint64_t nsym, float* __restrict__ metrics) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= nsym) return;
    const float* c = constellation(n);
    const float nd = ndist_of(c);
    const float2 p = __ldg(reinterpret_cast<const float2*>(iq) + i);
    const uint32_t M = 1u << n;
    for (uint32_t k = 0; k < M; ++k) metrics[i * M + k] = demap_one(c, k, p.x, p.y, nd);
}

// one thread per frame: compare the first L decoded bits with the payload (main.c:129-141)
__global__ void count_errors_kernel(const uint8_t* __restrict__ dec, const uint8_t* __restrict__ payload, uint64_t nframes,
                                    uint32_t L, uint32_t out_bytes, unsigned long long* __restrict__ counters) {
    const uint64_t f = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t errs = 0;
    if (f < nframes) {
        const uint32_t pb = (L + 7) / 8;
        const uint8_t* d = dec + f * out_bytes; const uint8_t* p = payload + f * pb;
        for (uint32_t j = 0; j < pb; ++j) {
            uint32_t x = (uint32_t)(d[j] ^ p[j]);
            if (j == pb - 1 && (L & 7)) x &= (0xFF00u >> (L & 7)) & 0xFFu;      // only the first L bits count
            errs += __popc(x);
        }
    }
    uint32_t ferr = errs ? 1u : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        errs += __shfl_down_sync(0xffffffffu, errs, off);
        ferr += __shfl_down_sync(0xffffffffu, ferr, off);
    }
    if ((threadIdx.x & 31) == 0 && (errs | ferr)) {
        atomicAdd(counters + 1, (unsigned long long)errs);
        atomicAdd(counters + 2, (unsigned long long)ferr);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(counters, (unsigned long long)nframes * L);
}

// sliding-window stitching: block b contributes bytes [skip, skip + keep) of its output row (the last block
// `last_keep` bytes) to the contiguous decoded stream
__global__ void stitch_kernel(const uint8_t* __restrict__ tmp, uint32_t row_bytes, uint32_t skip, uint64_t nblk, uint64_t keep,
                              uint64_t last_keep, uint8_t* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t total = (nblk - 1) * keep + last_keep;
    if (i >= total) return;
    const uint64_t b = keep ? (i / keep < nblk - 1 ? i / keep : nblk - 1) : 0;
    const uint64_t j = i - b * keep;
    out[i] = tmp[b * row_bytes + skip + j];
}

inline unsigned grid_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

int count_errors_launch(const uint8_t* dec, const uint8_t* payload, uint64_t nframes, uint32_t L, uint32_t out_bytes,
                        uint64_t* counters, cudaStream_t st) {
    if (nframes == 0) return 0;
    count_errors_kernel<<<grid_for(nframes, 256), 256, 0, st>>>(dec, payload, nframes, L, out_bytes,
                                                                reinterpret_cast<unsigned long long*>(counters));
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

int stitch_launch(const uint8_t* tmp, uint32_t row_bytes, uint32_t skip_bytes, uint64_t nblk, uint64_t keep_bytes,
                  uint64_t last_keep_bytes, uint8_t* out, cudaStream_t st) {
    const uint64_t total = (nblk - 1) * keep_bytes + last_keep_bytes;
    if (total == 0) return 0;
    stitch_kernel<<<grid_for(total, 256), 256, 0, st>>>(tmp, row_bytes, skip_bytes, nblk, keep_bytes, last_keep_bytes, out);
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

int demap_launch(uint32_t n, const float* iq, uint64_t nsym, float* metrics, cudaStream_t st) {
    if (n < 1 || n > 3) { set_error("demap: symlen_out %u unsupported (constellations exist for 1..3)", n); return -1; }
    if (nsym == 0) return 0;
    demap_kernel<<<grid_for(nsym, 256), 256, 0, st>>>(n, iq, nsym, metrics);
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

int synth_awgn_launch(const LaunchGeom& g, const uint8_t* labels_dev, uint64_t nframes, uint64_t first, uint64_t seed,
                      float sigma, uint8_t* payload, float* metrics, float* iq, cudaStream_t st) {
    if (g.n < 1 || g.n > 3) { set_error("synth: symlen_out %u unsupported", g.n); return -1; }
    if (nframes == 0) return 0;
    const uint32_t pb = (g.L + 7) / 8;
    payload_kernel<<<grid_for(nframes * ((pb + 3) / 4), 256), 256, 0, st>>>(nframes, first, seed, pb, payload);
    VDEC_CUDA(cudaGetLastError());
    synth_awgn_kernel<<<grid_for(nframes * g.T, 256), 256, 0, st>>>(g, labels_dev, nframes, first, seed, sigma, payload, metrics, iq);
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

int synth_bsc_launch(const LaunchGeom& g, const uint8_t* labels_dev, uint64_t nframes, uint64_t first, uint64_t seed,
                     uint32_t ppm, uint8_t* payload, uint8_t* symbols, cudaStream_t st) {
    if (nframes == 0) return 0;
    const uint32_t pb = (g.L + 7) / 8;
    payload_kernel<<<grid_for(nframes * ((pb + 3) / 4), 256), 256, 0, st>>>(nframes, first, seed, pb, payload);
    VDEC_CUDA(cudaGetLastError());
    synth_bsc_kernel<<<grid_for(nframes * g.T, 256), 256, 0, st>>>(g, labels_dev, nframes, first, seed, ppm, payload, symbols);
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vd
