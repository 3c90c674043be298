// mma_rate.cu -- tcgen05.mma issue/execute rate vs N, SS and TS forms (sm_100a), one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../whisper.rs_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include "ptx.cuh"
using namespace wb;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS, same D.  mode 1: TS (A from TMEM cols 256..), same D.  mode 2: SS alternating two D buffers
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 0) {
    const uint64_t da = umma_desc_k_sw128(smem_u32(smem));
    const uint64_t db = umma_desc_k_sw128(smem_u32(smem + 16384));
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
    // warm-up
    for (int i = 0; i < 8; ++i) umma_f16_ss(tb, da, db, idesc, 1);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const int k = i & 3;
      if (MODE == 0) umma_f16_ss(tb, da + 2 * k, db + 2 * k, idesc, 1);
      if (MODE == 1) umma_ts(tb, tb + 256 + 8 * k, db + 2 * k, idesc, 1);
      if (MODE == 2) umma_f16_ss(tb + ((i >> 2) & 1) * 128, da + 2 * k, db + 2 * k, idesc, 1);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

template <int N, int MODE>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  const int reps = 512;
  k<N, MODE><<<grid, 128, 49152>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-10s N=%3d grid=%3d  issue %.1f cyc/mma   complete %.1f cyc/mma  (ideal %d)  %s\n", name, N, grid, (double)h[0] / reps,
         (double)h[1] / reps, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<32, 0>("SS", grid); run<64, 0>("SS", grid); run<80, 0>("SS", grid); run<96, 0>("SS", grid); run<128, 0>("SS", grid); run<192, 0>("SS", grid); run<256, 0>("SS", grid);
    run<64, 1>("TS", grid); run<80, 1>("TS", grid); run<128, 1>("TS", grid); run<256, 1>("TS", grid);
    run<64, 2>("SS altD", grid); run<128, 2>("SS altD", grid);
  }
  return 0;
}
