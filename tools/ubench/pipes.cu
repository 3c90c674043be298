// pipes.cu -- per-SM throughput of the CUDA-core instructions the attention softmax leans on
// (sm_100a).  One CTA per SM, W warps, each running N dependent-free instructions in ILP chains.
// Prints instructions / clk / SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipes.cu -o pipes
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[ILP];
  uint32_t h[ILP];
  unsigned long long d[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = seed + threadIdx.x * 0.001f + i; h[i] = 0x3c003800u + threadIdx.x + i;
    d[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 2) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) % ILP]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) % ILP]));
      if (OP == 5) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) % ILP]), "f"(a[(i + 2) % ILP]));
      if (OP == 6) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(h[(i + 1) % ILP]));
      if (OP == 7) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 8) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 9) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 10) { int x = __float_as_int(a[i]); asm volatile("shl.b32 %0, %0, 23;" : "+r"(x)); a[i] = __int_as_float(x); }
      if (OP == 11) asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(*(unsigned short*)&h[i]) : "f"(a[i]));
      if (OP == 12) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(d[(i + 1) % ILP]));   // 2 FMAs per thread-instr
      if (OP == 13) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(d[(i + 1) % ILP]));
      if (OP == 14) asm volatile("mul.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(d[(i + 1) % ILP]));
    }
  }
  long long t1 = clock64();
  float s = 0; uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { s += a[i]; x ^= h[i]; x ^= (uint32_t)(d[i] >> 32) ^ (uint32_t)d[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int warps) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, warps * 32>>>(out, cyc, 0.5f);
  k<OP><<<148, warps * 32>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  double inst = (double)ITERS * ILP * warps * 32;
  printf("%-28s warps=%2d  %.1f thread-instr/clk/SM  (%.2f warp-instr/clk/SM)\n", name, warps, inst / avg, inst / avg / 32);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<12>("fma.f32x2", w); run<13>("add.f32x2", w); run<14>("mul.f32x2", w);
    if (w == 4) { run<0>("ex2.approx.ftz.f32", 4); run<1>("ex2.approx.f16x2", 4); run<7>("ex2.approx.ftz.bf16x2", 4); run<2>("cvt.rn.f16x2.f32", 4); run<11>("cvt.rn.f16.f32", 4); run<3>("fma.f32", 4); run<4>("max.f32", 4); run<5>("max3.f32", 4); run<6>("fma.f16x2", 4); run<8>("tanh.approx.f32", 4); run<9>("add.f32", 4); run<10>("shl.b32", 4); }
    if (w == 8) { run<0>("ex2.approx.ftz.f32", 8); run<1>("ex2.approx.f16x2", 8); run<7>("ex2.approx.ftz.bf16x2", 8); run<2>("cvt.rn.f16x2.f32", 8); run<11>("cvt.rn.f16.f32", 8); run<3>("fma.f32", 8); run<4>("max.f32", 8); run<5>("max3.f32", 8); run<6>("fma.f16x2", 8); run<8>("tanh.approx.f32", 8); run<9>("add.f32", 8); run<10>("shl.b32", 8); }
    if (w == 16) { run<0>("ex2.approx.ftz.f32", 16); run<1>("ex2.approx.f16x2", 16); run<2>("cvt.rn.f16x2.f32", 16); run<3>("fma.f32", 16); run<4>("max.f32", 16); run<5>("max3.f32", 16); }
  }
  return 0;
}
