// hbm_read.cu -- what a read-only stream reaches on this part, as the yardstick for the decoder's cross-attention
// (a read-only kernel; MEASURED_PEAKS.json's hbm_gbs is a copy: half reads, half writes).
//   (a) contiguous: every warp reads 512 contiguous bytes per instruction, U loads in flight per lane
//   (b) the cross-attention pattern: 128-byte head rows at a 1536-byte pitch (whisper small: d = 768), 8 lanes per row
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_read hbm_read.cu && ./hbm_read
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld_nc(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <int U>
__global__ void __launch_bounds__(256) read_contig(const uint4* __restrict__ p, size_t n16, uint32_t* sink) {
  uint32_t acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld_nc(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

// grid (H, n_seq): CTA (h, s) reads rows t = 0..T-1 of [s][t][h*128 bytes], pitch `pitch` bytes, K and V slabs
template <int U>
__global__ void __launch_bounds__(256, 3) read_heads(const char* __restrict__ k, const char* __restrict__ v, int T, int pitch,
                                                     uint32_t* sink) {
  const int h = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  const int grp = tid >> 3, ch = tid & 7;   // 32 row groups of 8 lanes
  const char* kb = k + ((size_t)s * T) * pitch + h * 128 + ch * 16;
  const char* vb = v + ((size_t)s * T) * pitch + h * 128 + ch * 16;
  uint32_t acc = 0;
  for (int t0 = grp; t0 < T; t0 += U * 32) {
    uint4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * 32;
      a[u] = t < T ? ld_nc(kb + (size_t)t * pitch) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * 32;
      b[u] = t < T ? ld_nc(vb + (size_t)t * pitch) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= a[u].x ^ a[u].y ^ a[u].z ^ a[u].w ^ b[u].x ^ b[u].y ^ b[u].z ^ b[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <class F>
static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t bytes = 4ull << 30;
  char* buf;
  uint32_t* sink;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, bytes);
  const size_t n16 = bytes / 16;
  printf("contiguous read of 4 GiB (best of 5):\n");
  for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
    float a = time_ms([&] { read_contig<4><<<blocks, 256>>>((const uint4*)buf, n16, sink); }, 5);
    float b = time_ms([&] { read_contig<8><<<blocks, 256>>>((const uint4*)buf, n16, sink); }, 5);
    printf("  %5d CTAs x 256: U=4 %.0f GB/s   U=8 %.0f GB/s\n", blocks, bytes / a * 1e-6, bytes / b * 1e-6);
  }
  // whisper small, 32 sequences, 12 layers back to back (each layer a fresh 147 MB: nothing is an L2 hit)
  const int T = 1500, H = 12, S = 32, pitch = 1536, L = 12;
  const size_t slab = (size_t)S * T * pitch;
  printf("cross-attention pattern (12 heads x 32 sequences, 128-byte rows at a 1536-byte pitch, %d layers = %.2f GB):\n", L,
         2.0 * slab * L * 1e-9);
  auto run4 = [&] { for (int l = 0; l < L; ++l) read_heads<4><<<dim3(H, S), 256>>>(buf + 2 * l * slab, buf + (2 * l + 1) * slab, T, pitch, sink); };
  auto run8 = [&] { for (int l = 0; l < L; ++l) read_heads<8><<<dim3(H, S), 256>>>(buf + 2 * l * slab, buf + (2 * l + 1) * slab, T, pitch, sink); };
  float a = time_ms(run4, 5), b = time_ms(run8, 5);
  printf("  U=4 (8 loads per lane in flight): %.1f us per layer, %.0f GB/s   U=8: %.1f us, %.0f GB/s\n", a * 1e3 / L, 2.0 * slab * L / a * 1e-6,
         b * 1e3 / L, 2.0 * slab * L / b * 1e-6);
  return 0;
}
