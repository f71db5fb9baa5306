// Micro-benchmark: issue rates of the instructions the softmax loop is made of (lanes per clock per SM), measured with clock64
// inside one resident CTA per SM:   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP> __device__ __forceinline__ void op8(float (&v)[8], uint32_t (&u)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
    if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
    if (OP == 2 && (i & 1) == 0)
      asm volatile("{\n\t.reg .b64 a;\n\tmov.b64 a, {%0, %1};\n\tfma.rn.f32x2 a, a, a, a;\n\tmov.b64 {%0, %1}, a;\n\t}" : "+f"(v[i]), "+f"(v[i + 1]));
    if (OP == 3 && (i & 1) == 0) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(v[i]), "f"(v[i + 1]));
    if (OP == 4 && (i & 1) == 0) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(u[i]), "r"(u[i + 1]));
    if (OP == 5) asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.bf16 %0, lo, %0;\n\t}" : "+f"(v[i]) : "r"(u[i]));
    if (OP == 6 && (i & 1) == 0)
      asm volatile("{\n\t.reg .b64 a;\n\tmov.b64 a, {%0, %1};\n\tadd.rn.f32x2 a, a, a;\n\tmov.b64 {%0, %1}, a;\n\t}" : "+f"(v[i]), "+f"(v[i + 1]));
    if (OP == 7) asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(v[(i + 1) & 7]));
  }
}
template <int OP> __global__ void k(float* out, long long* cyc, int iters) {
  float v[8]; uint32_t u[8];
  for (int i = 0; i < 8; ++i) { v[i] = threadIdx.x * 1e-3f + i; u[i] = threadIdx.x + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) op8<OP>(v, u);
  __syncthreads();
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, int per8, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<OP><<<148, threads>>>(out, cyc, 16);
  k<OP><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double insts = (double)iters * per8 * threads;       // thread-level instructions per CTA
  printf("%-28s %4d threads: %7.2f thread-instr / clk / SM  (%.2f clk per warp instruction per scheduler)\n", name, threads, insts / c, c / (insts / 32 / 4));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int threads : {512, 1024}) {
    run<0>("MUFU.EX2", 8, threads);
    run<1>("FFMA", 8, threads);
    run<2>("FFMA2 (fma.rn.f32x2)", 4, threads);
    run<6>("FADD2 (add.rn.f32x2)", 4, threads);
    run<3>("F2FP.BF16.PACK_AB", 4, threads);
    run<4>("PRMT", 4, threads);
    run<5>("FHADD.BF16 (f32 += bf16)", 8, threads);
    run<7>("FMNMX", 8, threads);
  }
  return 0;
}
