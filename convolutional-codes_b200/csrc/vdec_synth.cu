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
#include "vdec_demap.cuh"

namespace vd {

namespace {

// constellation points, interleaved (x, y): the table of vdec_demap.cuh's cst_x / cst_y (common/constellations.c:8-25)
__constant__ float c_cst1[4] = {0.707107f, 0.707107f, -0.707107f, -0.707107f};
__constant__ float c_cst2[8] = {0.707107f, 0.707107f, 0.707107f, -0.707107f, -0.707107f, 0.707107f, -0.707107f, -0.707107f};
__constant__ float c_cst3[16] = {0.408248f, 0.408248f, 0.408248f, 1.224745f, -0.408248f, 0.408248f, -1.224745f, 0.408248f,
                                 0.408248f, -0.408248f, 1.224745f, -0.408248f, -0.408248f, -0.408248f, -0.408248f, -1.224745f};

__device__ __forceinline__ const float* constellation(uint32_t n) { return n == 1 ? c_cst1 : (n == 2 ? c_cst2 : c_cst3); }

__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t (&o)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;                  // one IMAD.WIDE per product instead of IMAD + IMAD.HI
        asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(0xD2511F53u));
        asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(0xCD9E8D57u));
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// demapper.c:66-70 with demapper.c:43-45 for a run-time n — explicit round-to-nearest ops, no FMA contraction,
// IEEE division.  The stand-alone demap kernels use this form; the fused paths (synthetic source, decoder feeder) use
// vdec_demap.cuh's three-operation exact division, and the tests compare the two bit for bit.
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

// AWGN link: payload -> encoder -> mapper -> noise [-> demapper].
//   * T even (every BASELINE shape): one thread makes kSynthPairs PAIRS of consecutive steps (t, t+1) of one frame.  One
//     Philox4x32-10 call (4 words) feeds both symbols, the K+1 input bits of the pair are cut out of three payload
//     bytes once, and I/Q leaves as one 16-byte store per pair.  Box-Muller on the SFU: lg2.approx, sqrt.approx,
//     sin/cos.approx (absolute error ~1e-6 of sigma: irrelevant for a noise source, there is no parity target for it).
//   * T odd: one thread per symbol, same random numbers (word pair 2*(t & 1) of Philox block (frame, t >> 1)).
constexpr int kSynthPairs = 4, kSynthThreads = 256;

// symbol of the step whose input bit sits at position 0 of `win` (older bits above it)
__device__ __forceinline__ uint32_t label_of(uint32_t win, uint32_t K, const uint8_t* __restrict__ labels) {
    const uint32_t state = __brev((win >> 1) & ((1u << (K - 1)) - 1u)) >> (33u - K);     // bit t-1 -> position K-2
    return __ldg(labels + 2 * state + (win & 1u));
}
// input bits of a frame up to bit t (t at position 0, t-i at position i; bits outside [0, L) read as zero)
__device__ __forceinline__ uint32_t input_window(const uint8_t* __restrict__ pl, uint32_t t, uint32_t L, uint32_t pb) {
    const int j = (int)(t >> 3);                                  // byte holding bit t
    if ((pb & 3u) == 0 && (L & 7u) == 0) {
        // rows are whole 32-bit words and whole bytes (L = 1024: 128 B): two aligned word loads, bytes in stream order
        const uint32_t* pw = reinterpret_cast<const uint32_t*>(pl);
        const int wi = j >> 2;
        const uint32_t hi = wi >= 1 ? __byte_perm(__ldg(pw + wi - 1), 0, 0x0123) : 0u;       // bytes 4wi-4 .. 4wi-1, big-endian
        const uint32_t lo = (uint32_t)wi < (pb >> 2) ? __byte_perm(__ldg(pw + wi), 0, 0x0123) : 0u;
        // 64-bit stream word: bit index 32(wi-1) at position 63; bit t sits at position 31 - (t & 31)
        return __funnelshift_r(lo, hi, 31u - (t & 31u));
    }
    uint32_t w = 0;                                               // bits 8(j-2) .. 8j+7, bit 8(j-2) at position 23
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const int jj = j - 2 + b;
        uint32_t v = 0;
        if (jj >= 0 && jj < (int)pb) v = (uint32_t)__ldg(pl + jj);
        w = (w << 8) | v;
    }
    const int keep = (int)L - 8 * (j - 2);                        // bits of w that belong to the payload
    if (keep < 24) w = keep <= 0 ? 0u : (w & ~((1u << (24 - keep)) - 1u));
    return w >> (7u - (t & 7u));
}
// Encoder output for step t (encoder.c:84-119: bits enter MSB-first, steps >= L are the zero tail): label of (state
// before step t, input bit t), state bit K-2 = previous input (viterbi-decoder.c:65-66)
__device__ __forceinline__ uint32_t enc_symbol(const uint8_t* __restrict__ pl, uint32_t t, uint32_t K, uint32_t L, uint32_t pb,
                                               const uint8_t* __restrict__ labels) {
    return label_of(input_window(pl, t, L, pb), K, labels);
}

// constellation point `sym` plus Gaussian noise of standard deviation sigma per dimension from two random words
__device__ __forceinline__ void noisy_point(const float* s_c, uint32_t sym, uint32_t w0, uint32_t w1, float sigma, float& x, float& y) {
    // Box-Muller, one pair per symbol (gaussian.c:18-38 uses the same transform on rand()); u1 in (0, 1] keeps all 32
    // random bits where they matter (small u1 = large noise: the tail reaches 6.7 sigma)
    const float u1 = __fmaf_rn(__uint2float_rn(w0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float ang = __fmaf_rn(__uint2float_rn(w1 >> 8), 3.7450702e-7f, -3.14159265f);      // [-pi, pi)
    float lg, rad;                                                                            // u1 >= 2^-33: no denormals in sight
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
    const float r2 = fmaxf(lg * -1.3862943611198906f, 0.0f);                                  // -2 ln u1
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(r2));
    const float sr = sigma * rad;
    x = __fmaf_rn(sr, __cosf(ang), s_c[2 * sym]);                                             // main.c:100-102
    y = __fmaf_rn(sr, __sinf(ang), s_c[2 * sym + 1]);
}

template <int N>
__device__ __forceinline__ void store_metrics(float* __restrict__ dst, float x, float y) {
    constexpr int M = 1 << N;
    float bm[M];
    demap_point<N>(x, y, bm);
    if constexpr (M == 2) *reinterpret_cast<float2*>(dst) = make_float2(bm[0], bm[1]);
    else {
#pragma unroll
        for (int q = 0; q < M / 4; ++q)
            reinterpret_cast<float4*>(dst)[q] = make_float4(bm[4 * q], bm[4 * q + 1], bm[4 * q + 2], bm[4 * q + 3]);
    }
}

// T even.  npairs = nframes * T / 2 of this launch (< 2^31)
template <int N, bool MET, bool IQ>
__global__ void __launch_bounds__(kSynthThreads)
synth_awgn_pairs_kernel(LaunchGeom g, const uint8_t* __restrict__ labels, uint32_t npairs, uint64_t first, uint64_t seed, float sigma,
                        const uint8_t* __restrict__ payload, float* __restrict__ metrics, float* __restrict__ iq) {
    constexpr int M = 1 << N;
    __shared__ float s_c[2 * M];
    if (threadIdx.x < 2 * M) s_c[threadIdx.x] = constellation(N)[threadIdx.x];
    __syncthreads();
    const uint32_t half = g.T >> 1, pb = (g.L + 7) / 8, K = g.K, L = g.L;
    uint32_t p = blockIdx.x * (uint32_t)(kSynthThreads * kSynthPairs) + threadIdx.x;
    if (p >= npairs) return;
    uint32_t f = p / half, tp = p - f * half;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll 1
    for (int r = 0; r < kSynthPairs && p < npairs; ++r) {
        const uint64_t F = first + f;
        uint32_t o[4];
        philox4x32_10(k0, k1, (uint32_t)F, (uint32_t)(F >> 32), tp, 0x4E4F4953u, o);
        const uint32_t t = 2 * tp;
        const uint32_t win = input_window(payload + (size_t)f * pb, t + 1, L, pb);           // bit t+1 at position 0
        float x0, y0, x1, y1;
        noisy_point(s_c, label_of(win >> 1, K, labels), o[0], o[1], sigma, x0, y0);
        noisy_point(s_c, label_of(win, K, labels), o[2], o[3], sigma, x1, y1);
        if constexpr (IQ) reinterpret_cast<float4*>(iq)[p] = make_float4(x0, y0, x1, y1);
        if constexpr (MET) {
            store_metrics<N>(metrics + (size_t)p * (2 * M), x0, y0);
            store_metrics<N>(metrics + (size_t)p * (2 * M) + M, x1, y1);
        }
        p += kSynthThreads; tp += kSynthThreads;
        while (tp >= half) { tp -= half; ++f; }
    }
}

// any T: one thread per symbol
template <int N, bool MET, bool IQ>
__global__ void __launch_bounds__(kSynthThreads)
synth_awgn_kernel(LaunchGeom g, const uint8_t* __restrict__ labels, uint32_t nsym, uint64_t first, uint64_t seed, float sigma,
                  const uint8_t* __restrict__ payload, float* __restrict__ metrics, float* __restrict__ iq) {
    constexpr int M = 1 << N;
    __shared__ float s_c[2 * M];
    if (threadIdx.x < 2 * M) s_c[threadIdx.x] = constellation(N)[threadIdx.x];
    __syncthreads();
    const uint32_t s0 = blockIdx.x * (uint32_t)kSynthThreads + threadIdx.x;
    if (s0 >= nsym) return;
    const uint32_t f = s0 / g.T, t = s0 - f * g.T, pb = (g.L + 7) / 8;
    const uint64_t F = first + f;
    uint32_t o[4];
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)F, (uint32_t)(F >> 32), t >> 1, 0x4E4F4953u, o);
    float x, y;
    noisy_point(s_c, enc_symbol(payload + (size_t)f * pb, t, g.K, g.L, pb, labels), (t & 1u) ? o[2] : o[0], (t & 1u) ? o[3] : o[1], sigma, x, y);
    if constexpr (IQ) reinterpret_cast<float2*>(iq)[s0] = make_float2(x, y);
    if constexpr (MET) store_metrics<N>(metrics + (size_t)s0 * M, x, y);
}

__global__ void synth_bsc_kernel(LaunchGeom g, const uint8_t* __restrict__ labels, uint64_t nframes, uint64_t first,
                                 uint64_t seed, uint32_t ppm, const uint8_t* __restrict__ payload, uint8_t* __restrict__ symbols) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= nframes * g.T) return;
    const uint64_t f = i / g.T; const uint32_t t = (uint32_t)(i % g.T);
    const uint32_t pb = (g.L + 7) / 8;
    uint32_t sym = enc_symbol(payload + f * pb, t, g.K, g.L, pb, labels);
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
    const float nd = ndist_rt((int)n);
    const float2 p = __ldg(reinterpret_cast<const float2*>(iq) + i);
    const uint32_t M = 1u << n;
    for (uint32_t k = 0; k < M; ++k) metrics[i * M + k] = demap_one(c, k, p.x, p.y, nd);
}

// common/hard-demapper.c:62-102: snap the received point to the nearest constellation point (squared distances
// compared with strict '<', so the lowest index wins a tie, :68-77), then emit the soft demapper's expression
// evaluated at the SNAPPED point (:81-87).
__global__ void hard_demap_kernel(uint32_t n, const float* __restrict__ iq, uint64_t nsym, float* __restrict__ metrics) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= nsym) return;
    const float* c = constellation(n);
    const float nd = ndist_rt((int)n);
    const float2 p = __ldg(reinterpret_cast<const float2*>(iq) + i);
    const uint32_t M = 1u << n;
    float best = __int_as_float(0x7f800000);
    uint32_t idx = 0;
    for (uint32_t k = 0; k < M; ++k) {
        const float dx = __fsub_rn(p.x, c[2 * k]), dy = __fsub_rn(p.y, c[2 * k + 1]);
        const float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (d < best) { best = d; idx = k; }
    }
    const float sx = c[2 * idx], sy = c[2 * idx + 1];
    for (uint32_t k = 0; k < M; ++k) metrics[i * M + k] = demap_one(c, k, sx, sy, nd);
}

// Compare the first L decoded bits of every frame with its payload (main.c:129-141).  One warp per frame at a time
// (lanes read consecutive bytes of the two rows: coalesced although the decoded rows are (T+7)/8 bytes apart and hence
// unaligned), kCountFrames frames per warp, one pair of atomics per CTA.
constexpr int kCountThreads = 256, kCountFrames = 8;

__global__ void __launch_bounds__(kCountThreads)
count_errors_kernel(const uint8_t* __restrict__ dec, const uint8_t* __restrict__ payload, uint64_t nframes,
                    uint32_t L, uint32_t out_bytes, unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_err[kCountThreads / 32], s_ferr[kCountThreads / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t pb = (L + 7) / 8;
    const uint32_t last_mask = (L & 7) ? ((0xFF00u >> (L & 7)) & 0xFFu) : 0xFFu;      // only the first L bits count
    const uint64_t f0 = ((uint64_t)blockIdx.x * (kCountThreads / 32) + warp) * kCountFrames;
    uint32_t errs = 0, ferr = 0;
    uint32_t e[kCountFrames];
#pragma unroll
    for (int k = 0; k < kCountFrames; ++k) {                    // all loads of the warp's frames in flight together
        const uint64_t f = f0 + k;
        e[k] = 0;
        if (f < nframes) {
            const uint8_t* d = dec + f * out_bytes; const uint8_t* p = payload + f * pb;
            for (uint32_t j = lane; j < pb; j += 32) {
                uint32_t x = (uint32_t)(__ldg(d + j) ^ __ldg(p + j));
                if (j == pb - 1) x &= last_mask;
                e[k] += __popc(x);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kCountFrames; ++k) {
        const uint32_t t = __reduce_add_sync(0xffffffffu, e[k]);
        errs += t; ferr += t ? 1u : 0u;
    }
    if (lane == 0) { s_err[warp] = errs; s_ferr[warp] = ferr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0, b = 0;
#pragma unroll
        for (int w = 0; w < kCountThreads / 32; ++w) { a += s_err[w]; b += s_ferr[w]; }
        if (a | b) {
            atomicAdd(counters + 1, (unsigned long long)a);
            atomicAdd(counters + 2, (unsigned long long)b);
        }
        if (blockIdx.x == 0) atomicAdd(counters, (unsigned long long)nframes * L);
    }
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
    count_errors_kernel<<<grid_for(nframes, (kCountThreads / 32) * kCountFrames), kCountThreads, 0, st>>>(dec, payload, nframes, L, out_bytes,
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

int demap_launch(uint32_t n, const float* iq, uint64_t nsym, float* metrics, cudaStream_t st, bool hard) {
    if (n < 1 || n > 3) { set_error("demap: symlen_out %u unsupported (constellations exist for 1..3)", n); return -1; }
    if (nsym == 0) return 0;
    if (hard) hard_demap_kernel<<<grid_for(nsym, 256), 256, 0, st>>>(n, iq, nsym, metrics);
    else demap_kernel<<<grid_for(nsym, 256), 256, 0, st>>>(n, iq, nsym, metrics);
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

template <int N>
static int synth_awgn_n(const LaunchGeom& g, const uint8_t* labels_dev, uint32_t nsym, uint64_t first, uint64_t seed, float sigma,
                        const uint8_t* payload, float* metrics, float* iq, cudaStream_t st) {
    if ((g.T & 1u) == 0) {
        const uint32_t npairs = nsym / 2;
        const unsigned grid = grid_for(npairs, kSynthThreads * kSynthPairs);
        if (metrics && iq) synth_awgn_pairs_kernel<N, true, true><<<grid, kSynthThreads, 0, st>>>(g, labels_dev, npairs, first, seed, sigma, payload, metrics, iq);
        else if (metrics) synth_awgn_pairs_kernel<N, true, false><<<grid, kSynthThreads, 0, st>>>(g, labels_dev, npairs, first, seed, sigma, payload, metrics, iq);
        else synth_awgn_pairs_kernel<N, false, true><<<grid, kSynthThreads, 0, st>>>(g, labels_dev, npairs, first, seed, sigma, payload, metrics, iq);
    } else {
        const unsigned grid = grid_for(nsym, kSynthThreads);
        if (metrics && iq) synth_awgn_kernel<N, true, true><<<grid, kSynthThreads, 0, st>>>(g, labels_dev, nsym, first, seed, sigma, payload, metrics, iq);
        else if (metrics) synth_awgn_kernel<N, true, false><<<grid, kSynthThreads, 0, st>>>(g, labels_dev, nsym, first, seed, sigma, payload, metrics, iq);
        else synth_awgn_kernel<N, false, true><<<grid, kSynthThreads, 0, st>>>(g, labels_dev, nsym, first, seed, sigma, payload, metrics, iq);
    }
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

// metrics and/or iq may be null (not both)
int synth_awgn_launch(const LaunchGeom& g, const uint8_t* labels_dev, uint64_t nframes, uint64_t first, uint64_t seed,
                      float sigma, uint8_t* payload, float* metrics, float* iq, cudaStream_t st) {
    if (g.n < 1 || g.n > 3) { set_error("synth: symlen_out %u unsupported", g.n); return -1; }
    if (!metrics && !iq) { set_error("synth: neither metrics nor I/Q requested"); return -1; }
    if (nframes == 0) return 0;
    const uint32_t pb = (g.L + 7) / 8;
    payload_kernel<<<grid_for(nframes * ((pb + 3) / 4), 256), 256, 0, st>>>(nframes, first, seed, pb, payload);
    VDEC_CUDA(cudaGetLastError());
    uint64_t chunk = 0xF0000000ull / g.T;          // symbol indices are 32-bit inside a launch
    if (chunk < 1) chunk = 1;
    for (uint64_t done = 0; done < nframes; done += chunk) {
        const uint64_t nf = nframes - done < chunk ? nframes - done : chunk;
        const uint32_t nsym = (uint32_t)(nf * g.T);
        const uint8_t* pl = payload + done * pb;
        float* met = metrics ? metrics + done * g.T * g.M : nullptr;
        float* q = iq ? iq + done * g.T * 2 : nullptr;
        int rc = g.n == 1 ? synth_awgn_n<1>(g, labels_dev, nsym, first + done, seed, sigma, pl, met, q, st)
               : g.n == 2 ? synth_awgn_n<2>(g, labels_dev, nsym, first + done, seed, sigma, pl, met, q, st)
                          : synth_awgn_n<3>(g, labels_dev, nsym, first + done, seed, sigma, pl, met, q, st);
        if (rc != 0) return -1;
    }
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
