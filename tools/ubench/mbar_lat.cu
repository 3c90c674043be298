// mbar_lat.cu -- the latency of mbarrier.try_wait / test_wait on a COMPLETED phase against a plain ld.shared of the
// same word, as dependent chains on an otherwise idle SM.  B200: 54 / 54 / 46 cycles per operation (incl. ~8 of chain
// arithmetic).  (The word dump: a plain ld.shared of a live mbarrier reads 0 on sm_100a -- the state is not visible
// to generic loads, so a barrier cannot be polled without try_wait / test_wait.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_lat mbar_lat.cu && ./mbar_lat
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t lds64(const void* p) {
  uint64_t v;
  asm volatile("ld.volatile.shared::cta.b64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p;}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{.reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p;}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}

__global__ void k(uint64_t* words, long long* lat, uint32_t zero) {
  __shared__ __align__(8) uint64_t bar[2];
  if (threadIdx.x == 0) {
    int w = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(3) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    words[w++] = lds64(&bar[0]) | ((uint64_t)((volatile uint32_t*)bar)[1] << 32 ^ 0);                                   // 0: init(3)
    for (int ph = 0; ph < 3; ++ph) {
      for (int i = 0; i < 3; ++i) {
        if (i == 1) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(4096) : "memory");
          words[w++] = lds64(&bar[0]);
          asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(4096) : "memory");
        } else {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
        }
        words[w++] = lds64(&bar[0]);
      }
    }
    // latency on a completed phase: barrier 1, count 1, one arrival -> phase 0 complete
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
    uint32_t acc = 0;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < 64; ++i) acc += try_wait(&bar[1], acc & zero) - 1;   // dependent chain
      long long t1 = clock64();
      for (int i = 0; i < 64; ++i) acc += test_wait(&bar[1], acc & zero) - 1;
      long long t2 = clock64();
      for (int i = 0; i < 64; ++i) acc += (uint32_t)(lds64(&bar[1] + (acc & zero)) >> 63) & zero;
      long long t3 = clock64();
      lat[rep * 4 + 0] = t1 - t0;
      lat[rep * 4 + 1] = t2 - t1;
      lat[rep * 4 + 2] = t3 - t2;
      lat[rep * 4 + 3] = acc;
    }
  }
}

int main() {
  uint64_t* w;
  long long* l;
  cudaMalloc(&w, 64 * 8);
  cudaMalloc(&l, 16 * 8);
  cudaMemset(w, 0, 64 * 8);
  k<<<1, 32>>>(w, l, 0);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
  uint64_t hw[64];
  long long hl[16];
  cudaMemcpy(hw, w, sizeof(hw), cudaMemcpyDeviceToHost);
  cudaMemcpy(hl, l, sizeof(hl), cudaMemcpyDeviceToHost);
  const char* names[] = {"init(3)", "ph0 arrive", "ph0 arrive.expect_tx(4096)", "ph0 complete_tx", "ph0 arrive -> phase done",
                         "ph1 arrive", "ph1 arrive.expect_tx", "ph1 complete_tx", "ph1 arrive -> done",
                         "ph2 arrive", "ph2 arrive.expect_tx", "ph2 complete_tx", "ph2 arrive -> done"};
  for (int i = 0; i < 13; ++i) printf("%-32s %016llx\n", names[i], (unsigned long long)hw[i]);
  for (int r = 0; r < 3; ++r)
    printf("rep %d: try_wait %.1f  test_wait %.1f  ld.shared.b64 %.1f cycles per dependent op\n", r, hl[r * 4] / 64.0, hl[r * 4 + 1] / 64.0,
           hl[r * 4 + 2] / 64.0);
  return 0;
}
