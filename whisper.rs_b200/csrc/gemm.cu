// gemm.cu -- persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[b][m][n] = epilogue( sum_k A[b][m][k] * W[n][k] )      f16 x f16 -> f32 accumulate
//
// Replaces every galois_matmul / galois_conv_1d_* call site of the reference's encoder
// (src/main.rs:1834, 1856, 1891-1895, 1936, 1956, 1962, 1992, 2013) together with the
// repeat/add bias, scale, GELU, residual add and F16 repack ops that follow each of them, which
// the reference materialises as separate tensors.
//
// Structure (one CTA per SM, 256 threads):
//   warp 0   TMA producer: A tile 128 x 64 and W tile BN x 64 per k-block, SWIZZLE_128B, into a
//            STAGES-deep shared-memory ring guarded by full/empty mbarriers
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16) x 4 per k-block into
//            one of two TMEM accumulator buffers; tcgen05.commit releases smem stages and
//            publishes finished accumulators
//   warp 2   TMEM allocator (2 x BN columns, rounded up to a power of two)
//   warps 4-7 epilogue: tcgen05.ld 32 lanes x 32 columns, fused bias/scale/GELU/residual,
//            vectorised global stores (row-major F16/F32, or V^T scatter); overlaps the next
//            tile's main loop through the second accumulator buffer
// Tiles are scheduled round-robin over the persistent grid with n fastest, so CTAs running
// concurrently share the same A rows through L2 and the weight stays L2-resident.
#include <stdio.h>

#include "ptx.cuh"
#include "wb_kernels.hpp"

namespace wb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;           // one 128-byte swizzle atom of f16
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;       // two warps per TMEM lane quarter, each takes every other 32-column chunk
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);

struct GemmArgs {
  int M_rows, batch, N, K;
  int m_tiles, n_tiles, total_tiles;
  GemmEpilogue e;
};

template <int BN>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 192) ? 5 : (BN >= 128) ? 6 : 8;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                   : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void store_chunk_f16(__half* dst, const float (&v)[32]) {
  uint4* p = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_h2(v[8 * q + 0], v[8 * q + 1]);
    u.y = pack_h2(v[8 * q + 2], v[8 * q + 3]);
    u.z = pack_h2(v[8 * q + 4], v[8 * q + 5]);
    u.w = pack_h2(v[8 * q + 6], v[8 * q + 7]);
    p[q] = u;
  }
}
__device__ __forceinline__ void store_chunk_f32(float* dst, const float (&v)[32]) {
  float4* p = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int q = 0; q < 8; ++q) p[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap w_map,
                        const GemmArgs args) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&a_map);
    prefetch_tmap(&w_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EPI_WARPS);   // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = (args.K + BK - 1) / BK;
  const int tiles_per_batch = args.m_tiles * args.n_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_batch;
        const int r = tile - b * tiles_per_batch;
        const int mt = r / args.n_tiles;
        const int nt = r - mt * args.n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          tma_load_3d(sa, &a_map, &full_bar[stage], kb * BK, mt * BM, b);
          tma_load_2d(sb, &w_map, &full_bar[stage], kb * BK, nt * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 f16 = 32 bytes along K inside the swizzle atom: +2 in (addr >> 4) units
            umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);              // smem stage reusable once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const GemmEpilogue& e = args.e;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int chunk0 = (warp - EPI_WARP0) >> 2;   // 0 or 1: which interleaved half of the column chunks
    const int row_in_tile = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x, ++it) {
      const int b = tile / tiles_per_batch;
      const int r = tile - b * tiles_per_batch;
      const int mt = r / args.n_tiles;
      const int nt = r - mt * args.n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m = mt * BM + row_in_tile;
      const bool row_ok = m < args.M_rows;
      mbar_wait(&tmem_full[acc], acc_phase);
      __syncwarp();   // tcgen05.ld is warp-collective (.sync.aligned): reconverge after the spin
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) {
        const int n0 = nt * BN + c * 32;
        if (n0 >= args.N) break;            // warp-uniform
        uint32_t raw[32];
        tmem_ld_32x32b_x32(t_row + c * 32, raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        const bool full_chunk = (n0 + 32 <= args.N);
        if (!e.transpose_out) {
          if (e.bias) {
            if (full_chunk) {
              const float4* b4 = reinterpret_cast<const float4*>(e.bias + n0);
#pragma unroll
              for (int qd = 0; qd < 8; ++qd) {
                const float4 bb = __ldg(b4 + qd);
                v[4 * qd] += bb.x; v[4 * qd + 1] += bb.y; v[4 * qd + 2] += bb.z; v[4 * qd + 3] += bb.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < args.N) v[j] += __ldg(e.bias + n0 + j);
            }
          }
          if (e.colscale) {
            if (full_chunk) {
              const float4* c4 = reinterpret_cast<const float4*>(e.colscale + n0);
#pragma unroll
              for (int qd = 0; qd < 8; ++qd) {
                const float4 cc = __ldg(c4 + qd);
                v[4 * qd] *= cc.x; v[4 * qd + 1] *= cc.y; v[4 * qd + 2] *= cc.z; v[4 * qd + 3] *= cc.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < args.N) v[j] *= __ldg(e.colscale + n0 + j);
            }
          }
        } else if (row_ok) {
          if (e.bias) {
            const float bb = __ldg(e.bias + m);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bb;
          }
          if (e.colscale) {
            const float cs = __ldg(e.colscale + m);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= cs;
          }
        }
        if (e.scale != 1.0f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= e.scale;
        }
        if (e.gelu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_f16in(v[j]);
        }
        if (!row_ok) continue;
        if (e.transpose_out) {
          // C^T store: element (m, n) -> out[n*out_ld + m]; lanes hold consecutive m => coalesced
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            if (full_chunk || n < args.N) {
              float x = v[j];
              if (e.residual) x += e.residual[(long long)n * e.res_ld + m];
              if (e.out_f16) reinterpret_cast<__half*>(e.out)[(long long)n * e.out_ld + m] = __float2half_rn(x);
              else reinterpret_cast<float*>(e.out)[(long long)n * e.out_ld + m] = x;
            }
          }
          continue;
        }
        if (n0 >= e.vt_col0) {
          // V^T scatter: time contiguous (reference layout [T, Dh, H], src/main.rs:1914-1920)
          const int seg = m / e.vt_T;
          const int t = m - seg * e.vt_T;
          const int nn = n0 - e.vt_col0;   // multiple of 32: the chunk lies inside one head
          __half* dst = e.vt_out + ((long long)(seg * e.vt_heads + (nn >> 6)) * e.vt_head_rows + (nn & 63)) * e.vt_ld + t;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full_chunk || n0 + j < args.N) dst[(long long)j * e.vt_ld] = __float2half_rn(v[j]);
          continue;
        }
        if (e.residual) {
          const float* rp = e.residual + (long long)b * e.res_bstride + (long long)m * e.res_ld + n0;
          if (full_chunk) {
#pragma unroll
            for (int qd = 0; qd < 8; ++qd) {
              const float4 rr = *reinterpret_cast<const float4*>(rp + 4 * qd);
              v[4 * qd] += rr.x; v[4 * qd + 1] += rr.y; v[4 * qd + 2] += rr.z; v[4 * qd + 3] += rr.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < args.N) v[j] += rp[j];
          }
        }
        const long long o = e.out_slab_cols > 0
                                ? (long long)(n0 / e.out_slab_cols) * e.out_slab_stride + (long long)m * e.out_ld +
                                      (n0 % e.out_slab_cols)
                                : (long long)b * e.out_bstride + (long long)m * e.out_ld + n0;
        if (e.out_f16) {
          __half* dst = reinterpret_cast<__half*>(e.out) + o;
          if (full_chunk) {
            store_chunk_f16(dst, v);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < args.N) dst[j] = __float2half_rn(v[j]);
          }
        } else {
          float* dst = reinterpret_cast<float*>(e.out) + o;
          if (full_chunk) {
            store_chunk_f32(dst, v);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < args.N) dst[j] = v[j];
          }
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN>
cudaError_t launch_bn(const GemmProblem& g, const GemmArgs& a, int grid, cudaStream_t st) {
  gemm_f16_tcgen05_kernel<BN><<<grid, NUM_THREADS, Cfg<BN>::SMEM_BYTES, st>>>(g.a_map, g.w_map, a);
  return cudaGetLastError();
}

template <int BN>
bool set_attr(const char** err) {
  cudaError_t e = cudaFuncSetAttribute(gemm_f16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg<BN>::SMEM_BYTES);
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return false;
  }
  return true;
}

}  // namespace

bool gemm_setup_attributes(const char** err) {
  return set_attr<256>(err) && set_attr<192>(err) && set_attr<128>(err) && set_attr<64>(err) && set_attr<32>(err);
}

// widest tile that divides N (wide tiles halve shared-memory traffic per MMA); 192 serves d = 384
int gemm_pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N >= 256) return 256;   // ragged N: last tile masked
  if (N > 64) return 128;
  if (N > 32) return 64;
  return 32;
}

cudaError_t launch_gemm(const GemmProblem& g, int num_sms, cudaStream_t st) {
  GemmArgs a;
  a.M_rows = g.M_rows;
  a.batch = g.batch;
  a.N = g.N;
  a.K = g.K;
  a.m_tiles = (g.M_rows + BM - 1) / BM;
  a.n_tiles = (g.N + g.bn - 1) / g.bn;
  a.total_tiles = a.m_tiles * a.n_tiles * g.batch;
  a.e = g.epi;
  if (a.total_tiles <= 0) return cudaSuccess;
  const int grid = a.total_tiles < num_sms ? a.total_tiles : num_sms;
  switch (g.bn) {
    case 256: return launch_bn<256>(g, a, grid, st);
    case 192: return launch_bn<192>(g, a, grid, st);
    case 128: return launch_bn<128>(g, a, grid, st);
    case 64: return launch_bn<64>(g, a, grid, st);
    case 32: return launch_bn<32>(g, a, grid, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace wb
