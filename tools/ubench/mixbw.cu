// What bandwidth does a do-nothing kernel with the fused channelizer's traffic mix reach on this board?
// Per int16 I/Q sample: 4 bytes read, 8 bytes written (float2).  Variants: vector width and loads in flight.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mixbw tools/ubench/mixbw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int UN, bool STREAM>
__global__ void __launch_bounds__(256) k_mix(const uint4* __restrict__ in, float4* __restrict__ out, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * UN) {
    uint4 w[UN];
    #pragma unroll
    for (int u = 0; u < UN; u++) if (i + u * stride < n4) w[u] = __ldg(in + i + u * stride);
    #pragma unroll
    for (int u = 0; u < UN; u++) {
      if (i + u * stride >= n4) break;
      const uint32_t r[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
      float4 a, b;
      a.x = (float)(short)(r[0] & 0xffff); a.y = (float)((int)r[0] >> 16); a.z = (float)(short)(r[1] & 0xffff); a.w = (float)((int)r[1] >> 16);
      b.x = (float)(short)(r[2] & 0xffff); b.y = (float)((int)r[2] >> 16); b.z = (float)(short)(r[3] & 0xffff); b.w = (float)((int)r[3] >> 16);
      float4* o = out + 2 * (i + u * stride);
      if (STREAM) { __stcs(o, a); __stcs(o + 1, b); } else { o[0] = a; o[1] = b; }
    }
  }
}

// scalar form (what K1 and the fused kernels do): VEC samples per thread per step, UN steps in flight
template <int VEC, int UN>
__global__ void __launch_bounds__(256) k_mix_s(const uint32_t* __restrict__ in, float2* __restrict__ out, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x * VEC;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < n; i += stride * UN) {
    uint32_t w[UN][VEC];
    #pragma unroll
    for (int u = 0; u < UN; u++)
      if (i + u * stride < n) {
        if (VEC == 1) w[u][0] = __ldg(in + i + u * stride);
        else { const uint2 q = __ldg((const uint2*)(in + i + u * stride)); w[u][0] = q.x; w[u][VEC - 1] = q.y; }
      }
    #pragma unroll
    for (int u = 0; u < UN; u++) {
      if (i + u * stride >= n) break;
      float2 a = make_float2((float)(short)(w[u][0] & 0xffff), (float)((int)w[u][0] >> 16));
      if (VEC == 1) out[i + u * stride] = a;
      else {
        float2 b = make_float2((float)(short)(w[u][VEC - 1] & 0xffff), (float)((int)w[u][VEC - 1] >> 16));
        *(float4*)(out + i + u * stride) = make_float4(a.x, a.y, b.x, b.y);
      }
    }
  }
}
template <int VEC, int UN> void run_s(const char* name, const uint32_t* in, float2* out, long long n, int blocks) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 8; rep++) {
    cudaEventRecord(e0);
    k_mix_s<VEC, UN><<<blocks, 256>>>(in, out, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  printf("{\"kernel\": \"%s\", \"blocks\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", name, blocks, best, 12.0 * n / best / 1e6);
}

template <int UN, bool STREAM> void run(const char* name, const uint4* in, float4* out, long long n, int blocks) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 8; rep++) {
    cudaEventRecord(e0);
    k_mix<UN, STREAM><<<blocks, 256>>>(in, out, n / 4);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  printf("{\"kernel\": \"%s\", \"blocks\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", name, blocks, best, 12.0 * n / best / 1e6);
}

int main() {
  const long long n = 614400000;
  uint4* in; float4* out;
  cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 8);
  cudaMemset(in, 1, n * 4);
  for (int bps : {4, 8, 16}) {
    run<1, false>("ldg128 x1", in, out, n, 148 * bps);
    run<2, false>("ldg128 x2", in, out, n, 148 * bps);
    run<4, false>("ldg128 x4", in, out, n, 148 * bps);
    run<4, true>("ldg128 x4, st.cs", in, out, n, 148 * bps);
  }
  for (int bps : {8, 16, 32}) {
    run_s<1, 1>("ld32/st64 x1", (const uint32_t*)in, (float2*)out, n, 148 * bps);
    run_s<1, 4>("ld32/st64 x4", (const uint32_t*)in, (float2*)out, n, 148 * bps);
    run_s<2, 1>("ld64/st128 x1", (const uint32_t*)in, (float2*)out, n, 148 * bps);
    run_s<2, 4>("ld64/st128 x4", (const uint32_t*)in, (float2*)out, n, 148 * bps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
