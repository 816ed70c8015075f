// DFT-as-GEMM on the 5th-generation tensor cores, in our own kernel (north_star: "Tensor cores are evaluated only
// for a DFT-as-GEMM variant at small M, and kept only if ncu shows it winning").
//
// K3 for M = 64 as one tcgen05 GEMM per tile of 128 rows:  [Yr | Yi] = [Ur | Ui] x [[Wr, Wi], [-Wi, Wr]]
// (W = e^{+j 2 pi k p / 64}), operands staged in shared memory in the canonical K-major no-swizzle layout,
// accumulators in TMEM, read back with tcgen05.ld and streamed to global memory.  TF32 has 10 mantissa bits, far
// from the 1e-5 budget, so every operand is split x = hi + lo (hi = x with the low 13 mantissa bits cleared, lo =
// x - hi, exact) and three products are accumulated: hi*hi + hi*lo + lo*hi ("3xTF32").  The imaginary-part
// products use the SAME two B tiles (Wr, Wi) with the instruction descriptor's negate-A bit, so B costs 64 KB.
//
//   MMA shape: M = 128 (rows), N = 64, K = 8 per instruction (32 bytes of TF32); per tile 4 x 8 x 3 = 96 MMAs.
//
// Stand-alone on purpose (tools/, not the library): it answers the evaluation question with our own tcgen05 code
// and its ncu profile; the verdict is in DESIGN.md.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dft64_tc dft64_tc.cu ; run: ./dft64_tc [rows] [mode]
//   mode 3 = 3xTF32 (default), 1 = single TF32 product (accuracy ~1e-3, shows the tensor-pipe floor)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kM = 64;              // channels
constexpr int kRows = 128;          // rows per tile = MMA M
constexpr int kK = 128;             // K = [Ur (64) | Ui (64)]
constexpr int kLBO = 128;           // bytes between core matrices adjacent in K (8 rows x 16 B each, contiguous)
constexpr int kSBO_A = (kK / 4) * kLBO;    // bytes between 8-row groups of A: 32 core matrices of 128 B
constexpr int kSBO_B = (kM / 4) * kLBO;    // B tiles are [N = 64][K = 64]: 16 core matrices per 8-row group
constexpr int kAPart = kRows * kK * 4;     // 64 KB per A part (hi or lo)
constexpr int kBTile = kM * kM * 4;        // 16 KB per B tile (Wr or Wi, hi or lo)
constexpr int kSmem = 2 * kAPart + 4 * kBTile + 64;   // + mbarrier, tmem address

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major, no swizzle: element (mn, k) of a tile whose 8-row groups are `sbo` bytes apart
__device__ __forceinline__ int canon(int mn, int k, int sbo) { return (mn >> 3) * sbo + (k >> 2) * kLBO + (mn & 7) * 16 + (k & 3) * 4; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);               // start address, 16-byte units
  d |= (uint64_t)((kLBO >> 4) & 0x3FFF) << 16;          // leading byte offset (K direction)
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;           // stride byte offset (M/N direction)
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  return d;                                             // base offset 0, layout type 0 = no swizzle
}
// instruction descriptor: D = F32, A = B = TF32, K-major both, N = 64, M = 128
__device__ __forceinline__ uint32_t make_idesc(bool neg_a) {
  uint32_t d = 0;
  d |= 1u << 4;                  // c_format = F32
  d |= 2u << 7;                  // a_format = TF32
  d |= 2u << 10;                 // b_format = TF32
  d |= (neg_a ? 1u : 0u) << 13;  // a_negate
  d |= (uint32_t)(kM >> 3) << 17;      // n_dim = N / 8
  d |= (uint32_t)(kRows >> 4) << 24;   // m_dim = M / 16
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_dft64_tc(const float2* __restrict__ u, float2* __restrict__ y, long long nrows,
                                                     const float* __restrict__ bmat /* [4][64][64] canonical: Wr hi, Wi hi, Wr lo, Wi lo */) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_hi = smem;
  unsigned char* a_lo = smem + kAPart;
  unsigned char* b_s = smem + 2 * kAPart;
  uint64_t* bar = (uint64_t*)(smem + 2 * kAPart + 4 * kBTile);
  uint32_t* tmem_slot = (uint32_t*)(bar + 2);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

  for (int i = t; i < 4 * kBTile / 16; i += 256) ((float4*)b_s)[i] = ((const float4*)bmat)[i];
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const long long ntiles = (nrows + kRows - 1) / kRows;
  // Software pipeline over this CTA's tiles (A is single-buffered in shared memory, D double-buffered in TMEM):
  //   stage A(i) from registers | MMA(i) -> D[i & 1] runs asynchronously while the threads request tile i+1's samples
  //   and drain D[(i-1) & 1] (tcgen05.ld -> global) | wait for MMA(i) before A is overwritten
  uint32_t par[2] = {0u, 0u};
  float4 pre[16];                                    // the next tile's samples: 8 items x 2 float4 per thread
  auto item_of = [&](int it, int& row, int& quad) {
    const int item = it * 256 + t;                   // lanes walk rows first: 8 rows x 16 B = one contiguous core matrix
    row = (item & 7) + ((item >> 7) << 3); quad = (item >> 3) & 15;
  };
  auto prefetch = [&](long long tile) {
    const long long r0 = tile * kRows;
    #pragma unroll
    for (int it = 0; it < 8; it++) {
      int row, quad; item_of(it, row, quad);
      const long long gr = r0 + row;
      pre[2 * it] = make_float4(0.f, 0.f, 0.f, 0.f); pre[2 * it + 1] = pre[2 * it];
      if (tile < ntiles && gr < nrows) {
        const float4* src = (const float4*)(u + gr * kM + quad * 4);
        pre[2 * it] = __ldg(src); pre[2 * it + 1] = __ldg(src + 1);
      }
    }
  };
  auto wait_bar = [&](int b) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
        ::"r"(smem_u32(bar + b)), "r"(par[b]) : "memory");
  };
  // TMEM -> registers -> global for the tile whose accumulators sit in D[b].  Warp w reads lanes 32 (w % 4) .. +31;
  // warps 0-3 take channels 0..31, warps 4-7 channels 32..63 (real part from column c, imaginary from column 64 + c)
  auto epilogue = [&](long long tile, int b) {
    wait_bar(b);
    par[b] ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long r0 = tile * kRows;
    const int row = (warp & 3) * 32 + lane, c0 = (warp >> 2) * 32;
    const uint32_t taddr = tmem + b * 128 + ((uint32_t)((warp & 3) * 32) << 16) + c0;
    uint32_t re[32], im[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(re[0]), "=r"(re[1]), "=r"(re[2]), "=r"(re[3]), "=r"(re[4]), "=r"(re[5]), "=r"(re[6]), "=r"(re[7]), "=r"(re[8]), "=r"(re[9]), "=r"(re[10]), "=r"(re[11]), "=r"(re[12]), "=r"(re[13]), "=r"(re[14]), "=r"(re[15]), "=r"(re[16]), "=r"(re[17]), "=r"(re[18]), "=r"(re[19]), "=r"(re[20]), "=r"(re[21]), "=r"(re[22]), "=r"(re[23]), "=r"(re[24]), "=r"(re[25]), "=r"(re[26]), "=r"(re[27]), "=r"(re[28]), "=r"(re[29]), "=r"(re[30]), "=r"(re[31])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(im[0]), "=r"(im[1]), "=r"(im[2]), "=r"(im[3]), "=r"(im[4]), "=r"(im[5]), "=r"(im[6]), "=r"(im[7]), "=r"(im[8]), "=r"(im[9]), "=r"(im[10]), "=r"(im[11]), "=r"(im[12]), "=r"(im[13]), "=r"(im[14]), "=r"(im[15]), "=r"(im[16]), "=r"(im[17]), "=r"(im[18]), "=r"(im[19]), "=r"(im[20]), "=r"(im[21]), "=r"(im[22]), "=r"(im[23]), "=r"(im[24]), "=r"(im[25]), "=r"(im[26]), "=r"(im[27]), "=r"(im[28]), "=r"(im[29]), "=r"(im[30]), "=r"(im[31])
                 : "r"(taddr + kM));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (r0 + row < nrows) {
      float4* dst = (float4*)(y + (r0 + row) * kM + c0);
      #pragma unroll
      for (int c = 0; c < 32; c += 2)
        dst[c >> 1] = make_float4(__uint_as_float(re[c]), __uint_as_float(im[c]), __uint_as_float(re[c + 1]), __uint_as_float(im[c + 1]));
    }
  };

  prefetch(blockIdx.x);
  long long prev_tile = -1;
  int i = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, i++) {
    const int b = i & 1;
    // ---- stage A(i): split hi/lo, canonical layout; one 16-byte core-matrix row of the real block and one of the
    // imaginary block per item
    #pragma unroll
    for (int it = 0; it < 8; it++) {
      int row, quad; item_of(it, row, quad);
      const float4 v0 = pre[2 * it], v1 = pre[2 * it + 1];
      const float re[4] = {v0.x, v0.z, v1.x, v1.z}, im[4] = {v0.y, v0.w, v1.y, v1.w};
      float rh[4], rl[4], ih[4], il[4];
      #pragma unroll
      for (int c = 0; c < 4; c++) {
        rh[c] = __uint_as_float(__float_as_uint(re[c]) & 0xFFFFE000u); rl[c] = re[c] - rh[c];
        ih[c] = __uint_as_float(__float_as_uint(im[c]) & 0xFFFFE000u); il[c] = im[c] - ih[c];
      }
      const int o_re = canon(row, quad * 4, kSBO_A), o_im = canon(row, 64 + quad * 4, kSBO_A);
      *(float4*)(a_hi + o_re) = make_float4(rh[0], rh[1], rh[2], rh[3]);
      *(float4*)(a_hi + o_im) = make_float4(ih[0], ih[1], ih[2], ih[3]);
      if (MODE == 3) {
        *(float4*)(a_lo + o_re) = make_float4(rl[0], rl[1], rl[2], rl[3]);
        *(float4*)(a_lo + o_im) = make_float4(il[0], il[1], il[2], il[3]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                 // (also: every warp has drained D[b] two tiles ago)
    // ---- MMA(i): one thread issues everything for this tile
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t b_base = smem_u32(b_s);
      // D_r (columns 0..63)  = Ur Wr - Ui Wi ;  D_i (columns 64..127) = Ur Wi + Ui Wr
      #pragma unroll 1
      for (int prod = 0; prod < (MODE == 3 ? 3 : 1); prod++) {
        const uint32_t a_base = smem_u32(prod == 2 ? a_lo : a_hi);            // hi*hi, hi*lo, lo*hi
        const uint32_t bw = b_base + (prod == 1 ? 2 : 0) * kBTile;            // Wr tile (hi or lo); Wi tile follows it
        #pragma unroll 1
        for (int half = 0; half < 2; half++) {                                // 0: D_r, 1: D_i
          const uint32_t d = tmem + b * 128 + half * kM;
          #pragma unroll 1
          for (int blk = 0; blk < 2; blk++) {                                 // 0: Ur block of K, 1: Ui block
            const int use_wi = half ^ blk;                                    // D_r: Ur*Wr, -(Ui*Wi) ; D_i: Ur*Wi, Ui*Wr
            const uint32_t idesc = make_idesc(half == 0 && blk == 1);
            #pragma unroll
            for (int ks = 0; ks < 8; ks++) {                                  // K = 64 per block, 8 per MMA
              const uint64_t ad = make_desc(a_base + (blk * 16 + ks * 2) * kLBO, kSBO_A);
              const uint64_t bd = make_desc(bw + use_wi * kBTile + (ks * 2) * kLBO, kSBO_B);
              mma_tf32(d, ad, bd, idesc, !(prod == 0 && blk == 0 && ks == 0));
            }
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + b)) : "memory");
    }
    // ---- while the tensor core works: request the next tile, drain the previous one
    prefetch(tile + gridDim.x);
    if (prev_tile >= 0) epilogue(prev_tile, b ^ 1);
    prev_tile = tile;
    wait_bar(b);                                       // MMA(i) has read A: the next stage may overwrite it (parity kept for the epilogue)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  if (prev_tile >= 0) epilogue(prev_tile, (i - 1) & 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
  const long long nrows = argc > 1 ? atoll(argv[1]) : 9600000;      // configs[1]: 614.4 M samples / 64
  const int mode = argc > 2 ? atoi(argv[2]) : 3;
  // B tiles, canonical layout of [N = 64][K = 64] (K-major): element (n, k) = W[k][n]
  std::vector<float> bm(4 * 64 * 64, 0.f);
  for (int part = 0; part < 2; part++)        // 0: hi, 1: lo
    for (int which = 0; which < 2; which++)   // 0: Wr, 1: Wi
      for (int n = 0; n < 64; n++)
        for (int k = 0; k < 64; k++) {
          const double a = 2.0 * M_PI * (double)((n * k) % 64) / 64.0;
          const float w = (float)(which ? sin(a) : cos(a));
          uint32_t bits; memcpy(&bits, &w, 4);
          bits &= 0xFFFFE000u;
          float hi; memcpy(&hi, &bits, 4);
          const float val = part ? (w - hi) : hi;
          const int off = (n >> 3) * kSBO_B + (k >> 2) * kLBO + (n & 7) * 16 + (k & 3) * 4;
          bm[(size_t)(part * 2 + which) * 64 * 64 + off / 4] = val;
        }
  std::vector<float2> hu((size_t)std::min<long long>(nrows, 4096) * 64);
  srand(1);
  for (auto& v : hu) { v.x = (float)rand() / RAND_MAX - 0.5f; v.y = (float)rand() / RAND_MAX - 0.5f; }
  float2 *du, *dy; float* db;
  CK(cudaMalloc(&du, (size_t)nrows * 64 * 8)); CK(cudaMalloc(&dy, (size_t)nrows * 64 * 8)); CK(cudaMalloc(&db, bm.size() * 4));
  for (long long r = 0; r < nrows; r += 4096)       // tile the 4096 known rows over the whole input
    CK(cudaMemcpy(du + r * 64, hu.data(), (size_t)std::min<long long>(4096, nrows - r) * 64 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bm.data(), bm.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dy, 0, (size_t)nrows * 64 * 8));
  auto kern = mode == 3 ? k_dft64_tc<3> : k_dft64_tc<1>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  const int grid = 148;
  kern<<<grid, 256, kSmem>>>(du, dy, nrows, db);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 5;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) kern<<<grid, 256, kSmem>>>(du, dy, nrows, db);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  // accuracy against a double-precision DFT on the first rows and on the last tile
  std::vector<float2> hy(4096 * 64);
  const long long chk = std::min<long long>(nrows, 4096);
  CK(cudaMemcpy(hy.data(), dy, (size_t)chk * 64 * 8, cudaMemcpyDeviceToHost));
  double num = 0, den = 0;
  for (long long r = 0; r < std::min<long long>(chk, 300); r++)
    for (int k = 0; k < 64; k++) {
      double sr = 0, si = 0;
      for (int p = 0; p < 64; p++) {
        const double a = 2.0 * M_PI * (double)((k * p) % 64) / 64.0, c = cos(a), s = sin(a);
        const double xr = hu[r * 64 + p].x, xi = hu[r * 64 + p].y;
        sr += xr * c - xi * s; si += xr * s + xi * c;
      }
      const double dr = hy[r * 64 + k].x - sr, di = hy[r * 64 + k].y - si;
      num += dr * dr + di * di; den += sr * sr + si * si;
    }
  // last rows (partial-tile handling, 64-bit addressing)
  std::vector<float2> tail(64);
  CK(cudaMemcpy(tail.data(), dy + (nrows - 1) * 64, 64 * 8, cudaMemcpyDeviceToHost));
  const long long rr = (nrows - 1) % 4096;
  double tnum = 0, tden = 0;
  for (int k = 0; k < 64; k++) {
    double sr = 0, si = 0;
    for (int p = 0; p < 64; p++) {
      const double a = 2.0 * M_PI * (double)((k * p) % 64) / 64.0;
      sr += hu[rr * 64 + p].x * cos(a) - hu[rr * 64 + p].y * sin(a); si += hu[rr * 64 + p].x * sin(a) + hu[rr * 64 + p].y * cos(a);
    }
    tnum += (tail[k].x - sr) * (tail[k].x - sr) + (tail[k].y - si) * (tail[k].y - si); tden += sr * sr + si * si;
  }
  const double samples = (double)nrows * 64;
  printf("{\"kernel\": \"k_dft64_tc\", \"mode\": \"%s\", \"rows\": %lld, \"ms\": %.4f, \"GS_per_s\": %.1f, \"GBps_16B_per_sample\": %.1f, "
         "\"rel_rms_first_rows\": %.3e, \"rel_rms_last_row\": %.3e}\n", mode == 3 ? "3xTF32" : "1xTF32", nrows, ms, samples / ms / 1e6,
         samples * 16 / ms / 1e6, sqrt(num / den), sqrt(tnum / tden));
  return 0;
}
