// Micro-benchmark of the attention kernel's pass-2 chunk in isolation: per warp, per 16-column chunk
//   tcgen05.ld 32x32b.x16 -> 2 x LDS.128 (bias) -> 16 FHADD.BF16 -> 8 FADD2 -> 16 MUFU.EX2 -> 8 FADD2 -> 8 PRMT -> tcgen05.st x8
// clocks per chunk per warp as a function of the number of warps per SM (one CTA per SM, 512 TMEM columns).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o softmax_chunk softmax_chunk.cu && ./softmax_chunk
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
               "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void add_bias2(uint32_t w, float& s0, float& s1) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t}" : "+f"(s0), "+f"(s1) : "r"(w));
}
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}" : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// MODE bits: 1 no MUFU, 2 no bias (LDS + FHADD), 4 no TMEM traffic
template <int MODE> __global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, int iters, int nwarps) {
  __shared__ uint32_t slot;
  __shared__ __align__(16) uint16_t bias[144 * 152];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 144 * 152; i += blockDim.x) bias[i] = 0x3c00 + (i & 255);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32 % 256;
  const uint32_t brow = smem_u32(bias + ((warp & 3) * 32 + lane) * 152);
  const float m = 3.0f;
  float sm0 = 0.f, sm1 = 0.f;
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        uint32_t v[16];
        if (MODE & 4) { for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(sm0) + e + c; }
        else tmem_ld16(base + 16 * c, v);
        uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0;
        if (!(MODE & 2)) { b0 = lds128(brow + 32 * c); b1 = lds128(brow + 32 * c + 16); }
        const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        if (!(MODE & 4)) tmem_wait16(v);
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float s0 = __uint_as_float(v[2 * e]), s1 = __uint_as_float(v[2 * e + 1]);
          if (!(MODE & 2)) add_bias2(bw[e], s0, s1);
          add2(s0, s1, -m, -m);
          const float e0 = (MODE & 1) ? s0 : ex2(s0), e1 = (MODE & 1) ? s1 : ex2(s1);
          add2(sm0, sm1, e0, e1);
          pk[e] = __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632);
        }
        if (MODE & 4) { sm0 += __uint_as_float(pk[0] ^ pk[3] ^ pk[5] ^ pk[7]); sm1 += __uint_as_float(pk[1] ^ pk[2] ^ pk[4] ^ pk[6]); }
        else tmem_st8(base + 256 + 8 * c, pk);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm0 + sm1;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}
template <int MODE> void run(const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int nw : {4, 8, 16, 20}) {
    const int iters = 400;
    k<MODE><<<148, 1024>>>(out, cyc, 4, nw);
    k<MODE><<<148, 1024>>>(out, cyc, iters, nw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("%-34s %2d warps/SM: %7.1f clk per chunk per warp, %6.1f clk per chunk per SM (16 x 32 scores)  -> %5.2f scores/clk/SM\n", name, nw,
           c / (iters * 5), c / (iters * 5) / nw, 512.0 * nw / (c / (iters * 5)));
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("full chunk");
  run<1>("no MUFU");
  run<2>("no bias (LDS + FHADD)");
  run<4>("no TMEM ld / st");
  run<7>("FADD2 + PRMT only");
  return 0;
}
