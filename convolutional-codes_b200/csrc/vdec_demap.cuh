// vdec_demap.cuh — the demappers' arithmetic as device code, shared by the stand-alone demap kernels
// (vdec_synth.cu) and by the decoder kernels that take raw I/Q and form their branch metrics in the feeder
// (vdec_regs.cuh).
//
// Reference: common/demapper.c:43-45 (ndist), :66-70 (metric k = ((dx*dx)+(dy*dy))/ndist), compiled by gcc for
// baseline x86-64, i.e. every operation rounded to fp32 separately, no FMA contraction, IEEE division.
// common/hard-demapper.c:62-102 snaps the point to the nearest constellation point first (strict '<': lowest index
// wins a tie) and then computes the same expression from the snapped point.
//
// Constellation points are DATA of common/constellations.c:8-25 (unit power, Gray mapped), indices 1..3 = symlen_out.
#pragma once
#include <type_traits>
#include <utility>

#include "vdec_internal.h"

namespace vd {

VDEC_HD constexpr float cst_x(int n, int k) {
    constexpr float a = 0.707107f, b = 0.408248f, c = 1.224745f;
    if (n == 1) return k == 0 ? a : -a;
    if (n == 2) return (k & 2) ? -a : a;
    constexpr float x3[8] = {b, b, -b, -c, b, c, -b, -b};
    return x3[k & 7];
}
VDEC_HD constexpr float cst_y(int n, int k) {
    constexpr float a = 0.707107f, b = 0.408248f, c = 1.224745f;
    if (n == 1) return k == 0 ? a : -a;
    if (n == 2) return (k & 1) ? -a : a;
    constexpr float y3[8] = {b, c, b, b, -b, -b, -b, -c};
    return y3[k & 7];
}

#if defined(__CUDACC__)
// compile-time loop: f(integral_constant<int, I>) for I in [I0, N)
template <int I0, class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
    (f(std::integral_constant<int, I0 + Is>{}), ...);
}
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) static_for_impl<I>(f, std::make_integer_sequence<int, N - I>{});
}

// ndist = |c0 - c1|^2 as demapper.c:43-45 computes it at run time (fp32, separate roundings).
// Values: n = 1: 0x40800005, n = 2: 0x40000005, n = 3: 0x3f2aaab7.
__device__ __forceinline__ float ndist_rt(int n) {
    const float dx = __fsub_rn(cst_x(n, 0), cst_x(n, 1)), dy = __fsub_rn(cst_y(n, 0), cst_y(n, 1));
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// squared distance of (x, y) to point k, not yet normalised (demapper.c:67-69 without the division)
template <int N, int K>
__device__ __forceinline__ float sqdist(float x, float y) {
    constexpr float cx = cst_x(N, K), cy = cst_y(N, K);
    const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// s / ndist, correctly rounded, in three FMA-pipe operations instead of the ~10-instruction IEEE division sequence:
// q = RN(s * r), e = s - q * ndist (exact in one FMA), result = RN(q + e * r) with r = RN(1 / ndist).
// For the three ndist constants above this equals RN(s / ndist) for EVERY fp32 s in {0} u [1.3e-31, 2.2e38)
// (checked exhaustively on the CPU against the `/` operator, 2^31 values per constant; the exceptions below 1.3e-31 are
// residuals that underflow).  The demapper's s is either exactly 0 or at least 2^-52: a non-zero dx within a factor
// 2 of a constellation coordinate is an exact difference of two floats >= 0.2, hence a multiple of 2^-26, and any
// other dx is larger than 0.2.  Precondition on the caller's I/Q: finite, |I|, |Q| < 1e19.
template <int N>
__device__ __forceinline__ float div_ndist(float s) {
    constexpr uint32_t nd_bits = N == 1 ? 0x40800005u : N == 2 ? 0x40000005u : 0x3f2aaab7u;
    constexpr uint32_t rc_bits = N == 1 ? 0x3e7ffff6u : N == 2 ? 0x3efffff6u : 0x3fbffff2u;
    const float nd = __uint_as_float(nd_bits), rc = __uint_as_float(rc_bits);
    const float q = __fmul_rn(s, rc);
    const float e = __fmaf_rn(-q, nd, s);
    return __fmaf_rn(e, rc, q);
}

// all M = 2^N soft metrics of one received point (the decoder's input for one trellis step)
template <int N>
__device__ __forceinline__ void demap_point(float x, float y, float (&bm)[1 << N]) {
    static_for<0, (1 << N)>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        bm[k] = div_ndist<N>(sqdist<N, k>(x, y));
    });
}
#endif  // __CUDACC__

}  // namespace vd
