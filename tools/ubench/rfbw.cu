// Register-file read bandwidth probe: does a B200 SM sub-partition sustain more than two 32-bit register operand
// reads per cycle?  Each test is a loop of independent chains (8 per thread, 32 warps per SM) of one instruction or
// an alternating pair; `reads` = distinct register operands fetched per instruction (the operand-reuse cache cannot
// help: consecutive instructions use different registers in every slot).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rfbw rfbw.cu && ./rfbw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int NCH = 8, ITERS = 4096;
enum { T_FFMA3, T_FADD2_DISTINCT, T_FADD2_REUSE, T_FFMA3_FMNMX, T_FADD2D_FMNMX, T_FADD2R_FMNMX, T_FADD2D_2FMNMX, T_FADD2D_FADD, T_LOP3_3,
       T_FADD2D_LOP3, T_FADD2D_PRMT, T_ACS_GROUPED, T_ACS_INTERLEAVED, T_ACS_REUSE, NT };
static const char* names[NT] = {"FFMA a=b*c+a (3 distinct reads)", "FADD2 d=d+e (4 distinct reads)", "FADD2 d=d+c (c reused: 2 reads)",
    "FFMA(3)+FMNMX(2) 1:1", "FADD2(4)+FMNMX(2) 1:1", "FADD2(reuse,2)+FMNMX(2) 1:1", "FADD2(4)+2 FMNMX(2)", "FADD2(4)+FADD(2) 1:1", "LOP3 (3 distinct reads)",
    "FADD2(4)+LOP3(3) 1:1", "FADD2(4)+PRMT(2) 1:1",
    "ACS 6 FADD2 then 4 FMNMX + 2 PRMT + 2 LOP3 (grouped)", "ACS same instructions, FMA-pipe and ALU-pipe interleaved", "ACS interleaved, FADD2 second operand reused in pairs"};
static const int ninst[NT] = {1, 1, 1, 2, 2, 2, 3, 2, 1, 2, 2, 14, 14, 14};

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, float cf, uint32_t cu, long long* cyc) {
  float a[NCH], b[NCH], c[NCH];
  uint32_t u[NCH], v[NCH], w[NCH];
  unsigned long long d[NCH], e[NCH], g[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    a[i] = threadIdx.x * 0.5f + i; b[i] = 1.0f + i * 1e-3f + cf; c[i] = cf * (i + 1); u[i] = threadIdx.x + i; v[i] = u[i] * 3 + cu; w[i] = cu + i;
    d[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(b[i]);
    e[i] = ((unsigned long long)__float_as_uint(c[i]) << 32) | __float_as_uint(c[i] + 1.f);
    g[i] = e[i] + i;
  }
  const unsigned long long c2 = ((unsigned long long)__float_as_uint(cf) << 32) | __float_as_uint(cf);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (OP == T_FFMA3) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));
      if (OP == T_FADD2_DISTINCT) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(e[i]));
      if (OP == T_FADD2_REUSE) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(c2));
      if (OP == T_FFMA3_FMNMX) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i])); asm volatile("min.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(c[i])); }
      if (OP == T_FADD2D_FMNMX) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(e[i])); asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i])); }
      if (OP == T_FADD2R_FMNMX) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(c2)); asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i])); }
      if (OP == T_FADD2D_2FMNMX) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(e[i])); asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i])); asm volatile("max.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(c[i])); }
      if (OP == T_FADD2D_FADD) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(e[i])); asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i])); }
      if (OP == T_LOP3_3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(w[i]));
      if (OP == T_FADD2D_LOP3) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(e[i])); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(w[i])); }
      if (OP == T_FADD2D_PRMT) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(e[i])); asm volatile("prmt.b32 %0, %0, %1, 0x7531;" : "+r"(u[i]) : "r"(v[i])); }
      if (OP == T_ACS_GROUPED || OP == T_ACS_INTERLEAVED || OP == T_ACS_REUSE) {
        // the product recipe (vdec_regs.cuh acs_step_f2, sign-based decisions) for two register pairs: 4 state updates
        unsigned long long ee, oo, ff, pp, de, df;
        const unsigned long long q2 = g[i];
        const unsigned long long b0 = e[i], b1 = (OP == T_ACS_REUSE) ? e[i] : e[(i + 1) % NCH], b2 = e[(i + 2) % NCH], b3 = (OP == T_ACS_REUSE) ? e[(i + 2) % NCH] : e[(i + 3) % NCH];
        float n0, n1, n2, n3; uint32_t p0, p1;
#define F2A(dst, x, y) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(dst) : "l"(x), "l"(y))
#define F2S(dst, x, y) asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dst) : "l"(x), "l"(y))
#define LO(x) __uint_as_float((uint32_t)(x))
#define HI(x) __uint_as_float((uint32_t)((x) >> 32))
#define MN(dst, x, y) asm volatile("min.f32 %0, %1, %2;" : "=f"(dst) : "f"(x), "f"(y))
#define PR(dst, x) asm volatile("prmt.b32 %0, %1, %2, 0xFB00;" : "=r"(dst) : "r"((uint32_t)(x)), "r"((uint32_t)((x) >> 32)))
#define LP(acc, p) asm volatile("lop3.b32 %0, %0, %1, 0x01000100, 0xf8;" : "+r"(acc) : "r"(p))
        if (OP == T_ACS_GROUPED) {
          F2A(ee, d[i], b0); F2A(oo, q2, b1); F2A(ff, d[i], b2); F2A(pp, q2, b3); F2S(de, oo, ee); F2S(df, pp, ff);
          MN(n0, LO(ee), LO(oo)); MN(n1, HI(ee), HI(oo)); MN(n2, LO(ff), LO(pp)); MN(n3, HI(ff), HI(pp));
          PR(p0, de); LP(u[i], p0); PR(p1, df); LP(v[i], p1);
        } else {
          // one FMA-pipe instruction, one ALU-pipe instruction, ... (the ALU ones trail by one butterfly half)
          F2A(ee, d[i], b0); F2A(oo, q2, b1);
          F2S(de, oo, ee); MN(n0, LO(ee), LO(oo));
          F2A(ff, d[i], b2); MN(n1, HI(ee), HI(oo));
          F2A(pp, q2, b3); PR(p0, de);
          F2S(df, pp, ff); LP(u[i], p0);
          MN(n2, LO(ff), LO(pp)); MN(n3, HI(ff), HI(pp)); PR(p1, df); LP(v[i], p1);
        }
        d[i] = ((unsigned long long)__float_as_uint(n1) << 32) | __float_as_uint(n0);
        g[i] = ((unsigned long long)__float_as_uint(n3) << 32) | __float_as_uint(n2);
      }
    }
  }
  long long t1 = clock64();
  float s = 0; uint32_t x = 0; unsigned long long y = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) { s += a[i] + b[i] + c[i]; x ^= u[i] ^ v[i] ^ w[i]; y ^= d[i] ^ e[i] ^ g[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + x + (float)y;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(int sms, int threads, float* out, long long* cyc) {
  k<OP><<<sms, threads>>>(out, 0.25f, 7u, cyc);
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  k<OP><<<sms, threads>>>(out, 0.25f, 7u, cyc);
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < sms; ++i) mean += h[i]; mean /= sms;
  double winstr = (double)(threads / 32) * ITERS * NCH * ninst[OP];
  printf("{\"op\": \"%s\", \"threads_per_sm\": %d, \"warp_inst_per_clk_per_sm\": %.3f, \"cycles_per_chain_iteration_per_scheduler\": %.2f}\n",
         names[OP], threads, winstr / mean, mean / ((double)(threads / 128) * ITERS * NCH));
  return 0;
}
template <int OP> struct Loop { static int go(int sms, float* o, long long* c) {
  if (run<OP>(sms, 512, o, c)) return 1; if (run<OP>(sms, 1024, o, c)) return 1; return Loop<OP + 1>::go(sms, o, c); } };
template <> struct Loop<NT> { static int go(int, float*, long long*) { return 0; } };

int main() {
  cudaDeviceProp p; if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) return 1;
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d}\n", p.name, sms);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * 1024);
  return Loop<0>::go(sms, out, cyc);
}
