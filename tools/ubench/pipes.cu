// Micro-benchmarks of the sm_100a issue pipes that bound the channelizer's FIR stage (round 2 design aid):
// cycles per warp-instruction per SM sub-partition for I2F.S16 (conversion pipe), I2FP.F32.S32 (ALU), PRMT/LOP3,
// FFMA2 (packed FMA), scalar FFMA, LDS.64/LDS.128, and mixes of them.  One CTA of NW warps per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
template <int MODE>
__global__ void k(float* out, const uint32_t* in, long long* cyc) {
  __shared__ __align__(16) uint32_t sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i] ;
  __syncthreads();
  uint32_t r0 = in[threadIdx.x], r1 = in[threadIdx.x + 32], r2 = r0 ^ 0x55aa, r3 = r1 ^ 0x1234;
  float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
  float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
  const float2 h = make_float2(1.0001f, 1.0001f);
  float hs[8]; float2 hp[8];
  for (int i = 0; i < 8; i++) { hs[i] = __uint_as_float(in[64 + i + threadIdx.x]) * 1e-30f + 0.999f; hp[i] = make_float2(hs[i], hs[i] + 1e-7f); }
  if (MODE >= 10 && MODE <= 13) { f0 = 1e-3f * threadIdx.x; f1 = 2e-3f; f2 = 3e-3f; f3 = 4e-3f; }
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 8;
  const long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {          // 8 I2F.S16 (+ 8 FADD so the loop carries a dependency)
      float c0, c1, c2, c3, c4, c5, c6, c7;
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(c0) : "h"((short)r0));
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(c1) : "h"((short)r1));
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(c2) : "h"((short)r2));
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(c3) : "h"((short)r3));
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(c4) : "h"((short)(r0 >> 16)));
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(c5) : "h"((short)(r1 >> 16)));
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(c6) : "h"((short)(r2 >> 16)));
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(c7) : "h"((short)(r3 >> 16)));
      f0 += c0; f1 += c1; f2 += c2; f3 += c3; a0.x += c4; a0.y += c5; a1.x += c6; a1.y += c7;
    } else if (MODE == 1) {   // 8 I2FP.F32.S32 (+ 8 FADD)
      float c0, c1, c2, c3, c4, c5, c6, c7;
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(c0) : "r"(r0));
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(c1) : "r"(r1));
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(c2) : "r"(r2));
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(c3) : "r"(r3));
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(c4) : "r"(r0 + it));
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(c5) : "r"(r1 + it));
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(c6) : "r"(r2 + it));
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(c7) : "r"(r3 + it));
      f0 += c0; f1 += c1; f2 += c2; f3 += c3; a0.x += c4; a0.y += c5; a1.x += c6; a1.y += c7;
    } else if (MODE == 2) {   // 8 independent FFMA2
      a0 = __ffma2_rn(h, a0, h); a1 = __ffma2_rn(h, a1, h); a2 = __ffma2_rn(h, a2, h); a3 = __ffma2_rn(h, a3, h);
      a4 = __ffma2_rn(h, a4, h); a5 = __ffma2_rn(h, a5, h); a6 = __ffma2_rn(h, a6, h); a7 = __ffma2_rn(h, a7, h);
    } else if (MODE == 3) {   // 8 independent scalar FFMA
      f0 = fmaf(f0, 1.0001f, h.x); f1 = fmaf(f1, 1.0001f, h.x); f2 = fmaf(f2, 1.0001f, h.x); f3 = fmaf(f3, 1.0001f, h.x);
      a0.x = fmaf(a0.x, 1.0001f, h.x); a0.y = fmaf(a0.y, 1.0001f, h.x); a1.x = fmaf(a1.x, 1.0001f, h.x); a1.y = fmaf(a1.y, 1.0001f, h.x);
    } else if (MODE == 4) {   // 8 LDS.64
      uint32_t x, y;
      #pragma unroll
      for (int q = 0; q < 8; q++) { asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(sbase + q * 1024)); r0 ^= x; r1 ^= y; }
    } else if (MODE == 5) {   // FIR-like mix: 1 LDS.64, 4 I2F.S16, 16 FFMA2
      uint32_t x, y;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(sbase + (it & 7) * 1024));
      float2 u, v;
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(u.x) : "h"((short)x));
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(u.y) : "h"((short)(x >> 16)));
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(v.x) : "h"((short)y));
      asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(v.y) : "h"((short)(y >> 16)));
      a0 = __ffma2_rn(h, u, a0); a1 = __ffma2_rn(h, v, a1); a2 = __ffma2_rn(h, u, a2); a3 = __ffma2_rn(h, v, a3);
      a4 = __ffma2_rn(h, u, a4); a5 = __ffma2_rn(h, v, a5); a6 = __ffma2_rn(h, u, a6); a7 = __ffma2_rn(h, v, a7);
      a0 = __ffma2_rn(h, v, a0); a1 = __ffma2_rn(h, u, a1); a2 = __ffma2_rn(h, v, a2); a3 = __ffma2_rn(h, u, a3);
      a4 = __ffma2_rn(h, v, a4); a5 = __ffma2_rn(h, u, a5); a6 = __ffma2_rn(h, v, a6); a7 = __ffma2_rn(h, u, a7);
    } else if (MODE == 6) {   // same with I2FP path: 1 LDS.64, 2 PRMT-ish + 2 SHF + 4 I2FP, 16 FFMA2
      uint32_t x, y;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(sbase + (it & 7) * 1024));
      float2 u, v;
      int xi = __byte_perm(x, 0, 0x9910), xq = ((int)x) >> 16, yi = __byte_perm(y, 0, 0x9910), yq = ((int)y) >> 16;
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(u.x) : "r"(xi));
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(u.y) : "r"(xq));
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(v.x) : "r"(yi));
      asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(v.y) : "r"(yq));
      a0 = __ffma2_rn(h, u, a0); a1 = __ffma2_rn(h, v, a1); a2 = __ffma2_rn(h, u, a2); a3 = __ffma2_rn(h, v, a3);
      a4 = __ffma2_rn(h, u, a4); a5 = __ffma2_rn(h, v, a5); a6 = __ffma2_rn(h, u, a6); a7 = __ffma2_rn(h, v, a7);
      a0 = __ffma2_rn(h, v, a0); a1 = __ffma2_rn(h, u, a1); a2 = __ffma2_rn(h, v, a2); a3 = __ffma2_rn(h, u, a3);
      a4 = __ffma2_rn(h, v, a4); a5 = __ffma2_rn(h, u, a5); a6 = __ffma2_rn(h, v, a6); a7 = __ffma2_rn(h, u, a7);
    } else if (MODE == 7) {   // 8 PRMT (ALU)
      r0 = __byte_perm(r0, r1, 0x9910); r1 = __byte_perm(r1, r2, 0x9932); r2 = __byte_perm(r2, r3, 0x9910); r3 = __byte_perm(r3, r0, 0x9932);
      r0 = __byte_perm(r0, r1, 0x5410); r1 = __byte_perm(r1, r2, 0x7632); r2 = __byte_perm(r2, r3, 0x5410); r3 = __byte_perm(r3, r0, 0x7632);
    } else if (MODE == 8) {   // 8 LDS.128
      uint32_t x, y, z, w;
      #pragma unroll
      for (int q = 0; q < 8; q++) { asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(sbase + (threadIdx.x & 31) * 8 + q * 1024)); r0 ^= x; r1 ^= y; r2 ^= z; r3 ^= w; }
    } else if (MODE == 10) {  // kernel form: 8 FFMA2 with scalar-broadcast tap, SHARED sample pair, distinct accumulators
      const float2 x = make_float2(f0, f1);
      a0 = __ffma2_rn(make_float2(hs[0], hs[0]), x, a0); a1 = __ffma2_rn(make_float2(hs[1], hs[1]), x, a1);
      a2 = __ffma2_rn(make_float2(hs[2], hs[2]), x, a2); a3 = __ffma2_rn(make_float2(hs[3], hs[3]), x, a3);
      a4 = __ffma2_rn(make_float2(hs[4], hs[4]), x, a4); a5 = __ffma2_rn(make_float2(hs[5], hs[5]), x, a5);
      a6 = __ffma2_rn(make_float2(hs[6], hs[6]), x, a6); a7 = __ffma2_rn(make_float2(hs[7], hs[7]), x, a7);
    } else if (MODE == 11) {  // same with a DIFFERENT sample pair per FFMA2 (no operand reuse possible)
      const float2 x0 = make_float2(f0, f1), x1 = make_float2(f2, f3), x2 = make_float2(f1, f2), x3 = make_float2(f3, f0);
      a0 = __ffma2_rn(make_float2(hs[0], hs[0]), x0, a0); a1 = __ffma2_rn(make_float2(hs[1], hs[1]), x1, a1);
      a2 = __ffma2_rn(make_float2(hs[2], hs[2]), x2, a2); a3 = __ffma2_rn(make_float2(hs[3], hs[3]), x3, a3);
      a4 = __ffma2_rn(make_float2(hs[4], hs[4]), x0, a4); a5 = __ffma2_rn(make_float2(hs[5], hs[5]), x1, a5);
      a6 = __ffma2_rn(make_float2(hs[6], hs[6]), x2, a6); a7 = __ffma2_rn(make_float2(hs[7], hs[7]), x3, a7);
    } else if (MODE == 12) {  // tap duplicated in a register PAIR (no scalar-broadcast form), shared sample
      const float2 x = make_float2(f0, f1);
      a0 = __ffma2_rn(hp[0], x, a0); a1 = __ffma2_rn(hp[1], x, a1); a2 = __ffma2_rn(hp[2], x, a2); a3 = __ffma2_rn(hp[3], x, a3);
      a4 = __ffma2_rn(hp[4], x, a4); a5 = __ffma2_rn(hp[5], x, a5); a6 = __ffma2_rn(hp[6], x, a6); a7 = __ffma2_rn(hp[7], x, a7);
    } else if (MODE == 13) {  // scalar FFMA pairs: 16 FFMA, shared sample
      a0.x = fmaf(hs[0], f0, a0.x); a0.y = fmaf(hs[0], f1, a0.y); a1.x = fmaf(hs[1], f0, a1.x); a1.y = fmaf(hs[1], f1, a1.y);
      a2.x = fmaf(hs[2], f0, a2.x); a2.y = fmaf(hs[2], f1, a2.y); a3.x = fmaf(hs[3], f0, a3.x); a3.y = fmaf(hs[3], f1, a3.y);
      a4.x = fmaf(hs[4], f0, a4.x); a4.y = fmaf(hs[4], f1, a4.y); a5.x = fmaf(hs[5], f0, a5.x); a5.y = fmaf(hs[5], f1, a5.y);
      a6.x = fmaf(hs[6], f0, a6.x); a6.y = fmaf(hs[6], f1, a6.y); a7.x = fmaf(hs[7], f0, a7.x); a7.y = fmaf(hs[7], f1, a7.y);
    } else if (MODE == 9) {   // magic-number unpack mix: 1 LDS.64, 2 LOP3 + 2 PRMT + 2 FADD2, 16 FFMA2
      uint32_t x, y;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(sbase + (it & 7) * 1024));
      float2 u = make_float2(__uint_as_float((x & 0xffffu) ^ 0x4B408000u), __uint_as_float(__byte_perm(x, 0x4B400000u, 0x7632) ^ 0x8000u));
      float2 v = make_float2(__uint_as_float((y & 0xffffu) ^ 0x4B408000u), __uint_as_float(__byte_perm(y, 0x4B400000u, 0x7632) ^ 0x8000u));
      u = __fadd2_rn(u, make_float2(-12615680.0f, -12615680.0f)); v = __fadd2_rn(v, make_float2(-12615680.0f, -12615680.0f));
      a0 = __ffma2_rn(h, u, a0); a1 = __ffma2_rn(h, v, a1); a2 = __ffma2_rn(h, u, a2); a3 = __ffma2_rn(h, v, a3);
      a4 = __ffma2_rn(h, u, a4); a5 = __ffma2_rn(h, v, a5); a6 = __ffma2_rn(h, u, a6); a7 = __ffma2_rn(h, v, a7);
      a0 = __ffma2_rn(h, v, a0); a1 = __ffma2_rn(h, u, a1); a2 = __ffma2_rn(h, v, a2); a3 = __ffma2_rn(h, u, a3);
      a4 = __ffma2_rn(h, v, a4); a5 = __ffma2_rn(h, u, a5); a6 = __ffma2_rn(h, v, a6); a7 = __ffma2_rn(h, u, a7);
    }
  }
  const long long t1 = clock64();
  float s = f0 + f1 + f2 + f3 + a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + a4.x + a5.x + a6.x + a7.x + a4.y + a5.y + a6.y + a7.y;
  s += (float)(r0 ^ r1 ^ r2 ^ r3);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int per_iter, int nw) {
  float* out; uint32_t* in; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, 148 * 8);
  cudaMemset(in, 0x11, 4096 * 4);
  k<MODE><<<148, nw * 32>>>(out, in, cyc);
  k<MODE><<<148, nw * 32>>>(out, in, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
  // warp-instructions per SMSP = ITERS * per_iter * nw / 4
  printf("%-44s nw=%2d  %8.0f cyc  -> %.3f cyc per warp-instr per SMSP (%d instr/iter)  [%s]\n", name, nw, avg,
         avg / ((double)ITERS * per_iter * nw / 4), per_iter, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(in); cudaFree(cyc);
}
int main() {
  for (int nw : {4, 16}) {
    run<0>("I2F.S16 x8 + FADD x8", 16, nw);
    run<1>("I2FP.F32.S32 x8 + FADD x8 (+4 IADD)", 20, nw);
    run<2>("FFMA2 x8", 8, nw);
    run<3>("FFMA x8", 8, nw);
    run<4>("LDS.64 x8 (+8 LOP)", 8, nw);
    run<8>("LDS.128 x8", 8, nw);
    run<7>("PRMT x8", 8, nw);
    run<5>("FIR mix: LDS.64 + 4 I2F.S16 + 16 FFMA2", 21, nw);
    run<6>("FIR mix: LDS.64 + 2 PRMT 2 SHF 4 I2FP + 16 FFMA2", 25, nw);
    run<10>("FFMA2 x8 scalar-bcast tap, shared x", 8, nw);
    run<11>("FFMA2 x8 scalar-bcast tap, 4 different x", 8, nw);
    run<12>("FFMA2 x8 pair tap, shared x", 8, nw);
    run<13>("FFMA x16 scalar, shared x", 16, nw);
    run<9>("FIR mix: LDS.64 + magic (4 LOP/PRMT 2 FADD2) + 16 FFMA2", 25, nw);
  }
  return 0;
}
