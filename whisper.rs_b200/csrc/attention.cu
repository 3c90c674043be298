// attention.cu -- fused softmax attention on tcgen05 for sm_100a.
//
// Replaces galois_flash_attn (src/main.rs:1787-1797, call 1922) and the F16 repack / permute /
// merge ops around it (1898-1929): per head h and query n,
//     out[n] = sum_m softmax_m( K[m].Q[n] / sqrt(Dh) ) * V[m],      non-causal, Dh = 64.
// As in the reference, Q, K, V arrive rounded to F16 and the probabilities are rounded to F16
// before the P.V product; scores, running max and running sum stay in f32.
//
// One CTA = one (segment, head, 128-query tile); 128 threads, thread r owns query row r, which is
// also TMEM lane r, so the row max / row sum need no cross-thread reduction.
//   S  = Q K_j^T      tcgen05.mma 128 x 128 x 64 -> TMEM columns [0,128)
//   P  = exp2(...)    registers -> F16 -> shared memory in the UMMA K-major SWIZZLE_128B layout
//   O_j = P V_j       tcgen05.mma 128 x 64 x 128 -> TMEM columns [128,192), folded into a
//                     register accumulator with the online-softmax rescale
// K_j / V_j tiles are double-buffered through TMA; Q and K are read straight out of the QKV GEMM's
// row-major [tokens][2d] output through a 4-D tensor map (no head-major repack), V from the
// transposed [seg][h][64][Tp] copy the GEMM epilogue scatters (the reference's V layout).
// Two CTAs are resident per SM (112 KB shared memory, 256 TMEM columns each) so one CTA's softmax
// overlaps the other's MMAs.
#include "ptx.cuh"
#include "wb_kernels.hpp"

namespace wb {

namespace {

constexpr int QT = 128;   // queries per CTA
constexpr int KT = 128;   // keys per iteration
constexpr int DH = 64;
constexpr int TILE_QK_BYTES = QT * DH * 2;        // 16 KB
constexpr int TILE_V_HALF_BYTES = DH * 64 * 2;    // 8 KB: [64 dh rows][64 keys]
constexpr int SMEM_Q = 0;
constexpr int SMEM_K = SMEM_Q + TILE_QK_BYTES;                 // 2 stages
constexpr int SMEM_V = SMEM_K + 2 * TILE_QK_BYTES;             // 2 stages x 2 halves
constexpr int SMEM_P = SMEM_V + 2 * 2 * TILE_V_HALF_BYTES;     // 2 sub-tiles [128][64]
constexpr int SMEM_BAR = SMEM_P + 2 * TILE_QK_BYTES;
constexpr int ATTN_SMEM_BYTES = SMEM_BAR + 64;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t TMEM_S = 0, TMEM_O = 128;

struct AttnArgs {
  int B, T, H, n_kt;
  __half* out;
  float scale_log2;   // scale * log2(e)
};

__global__ void __launch_bounds__(128, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap qk_map, const __grid_constant__ CUtensorMap vt_map,
                         const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* bar_kv = bar_q + 1;   // [2]
  uint64_t* bar_s = bar_q + 3;
  uint64_t* bar_o = bar_q + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 5);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int H = a.H, T = a.T;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();   // SWIZZLE_128B tiles need 1024-byte alignment
    prefetch_tmap(&qk_map);
    prefetch_tmap(&vt_map);
    mbar_init(bar_q, 1);
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();   // tcgen05.alloc is warp-collective: reconverge after the tid == 0 branch
    tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto load_kv = [&](int j) {
    const int st = j & 1;
    mbar_arrive_expect_tx(&bar_kv[st], TILE_QK_BYTES + 2 * TILE_V_HALF_BYTES);
    tma_load_4d(smem + SMEM_K + st * TILE_QK_BYTES, &qk_map, &bar_kv[st], 0, H + h, j * KT, b);
    const int vrow = (b * H + h) * DH;
    tma_load_2d(smem + SMEM_V + (st * 2 + 0) * TILE_V_HALF_BYTES, &vt_map, &bar_kv[st], j * KT, vrow);
    tma_load_2d(smem + SMEM_V + (st * 2 + 1) * TILE_V_HALF_BYTES, &vt_map, &bar_kv[st], j * KT + 64, vrow);
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, TILE_QK_BYTES);
    tma_load_4d(smem + SMEM_Q, &qk_map, bar_q, 0, h, qt * QT, b);
    load_kv(0);
    if (a.n_kt > 1) load_kv(1);
  }

  constexpr uint32_t idesc_s = umma_idesc_f16(QT, KT);   // 128 x 128
  constexpr uint32_t idesc_o = umma_idesc_f16(QT, DH);   // 128 x 64
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  const int r = tid;                                      // query row in the tile
  const uint32_t p_row = smem_u32(smem + SMEM_P) + r * 128;
  const int sw = r & 7;

  float m_run = -INFINITY, l_run = 0.0f;
  float o_acc[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) o_acc[c] = 0.0f;

  for (int j = 0; j < a.n_kt; ++j) {
    const int st = j & 1;
    if (tid == 0) {
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(&bar_kv[st], (j >> 1) & 1);
      tc_fence_after();
      const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + SMEM_Q));
      const uint64_t dk = umma_desc_k_sw128(smem_u32(smem + SMEM_K + st * TILE_QK_BYTES));
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_f16_ss(tmem_base + TMEM_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, j & 1);
    __syncwarp();   // tcgen05.ld is warp-collective: reconverge after thread 0's issue branch / the spin
    tc_fence_after();

    // ---- pass 1: row max over the valid keys of this tile
    const int kbase = j * KT;
    const bool ragged = kbase + KT > T;   // CTA-uniform; only the last key tile
    float mx = m_run;
#pragma unroll 1
    for (int c = 0; c < KT / 32; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(tmem_base + lane_base + TMEM_S + c * 32, raw);
      tmem_ld_wait();
      if (!ragged) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (kbase + c * 32 + i < T) mx = fmaxf(mx, __uint_as_float(raw[i]));
      }
    }
    const float alpha = exp2f((m_run - mx) * a.scale_log2);   // 0 on the first tile (m_run = -inf)
    m_run = mx;
    const float moff = mx * a.scale_log2;
    // ---- pass 2: p = exp2(s*c - m*c), row sum, F16 pack into the swizzled P tile
    float psum = 0.0f;
#pragma unroll 1
    for (int c = 0; c < KT / 32; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(tmem_base + lane_base + TMEM_S + c * 32, raw);
      tmem_ld_wait();
      float p[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float e = exp2f(fmaf(__uint_as_float(raw[i]), a.scale_log2, -moff));
        if (ragged && kbase + c * 32 + i >= T) e = 0.0f;
        p[i] = e;
        psum += e;
      }
      // chunk c covers keys [32c, 32c+32): sub-tile c/2, 16-byte chunks (c&1)*4 .. +3
      const uint32_t sub = p_row + (c >> 1) * TILE_QK_BYTES;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int chunk = (c & 1) * 4 + q4;
        const uint32_t addr = sub + ((chunk ^ sw) << 4);
        const uint32_t x0 = pack_h2(p[8 * q4 + 0], p[8 * q4 + 1]), x1 = pack_h2(p[8 * q4 + 2], p[8 * q4 + 3]);
        const uint32_t x2 = pack_h2(p[8 * q4 + 4], p[8 * q4 + 5]), x3 = pack_h2(p[8 * q4 + 6], p[8 * q4 + 7]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x0), "r"(x1), "r"(x2), "r"(x3)
                     : "memory");
      }
    }
    l_run = l_run * alpha + psum;
#pragma unroll
    for (int c = 0; c < DH; ++c) o_acc[c] *= alpha;

    fence_proxy_async_smem();   // P (generic-proxy stores) -> visible to the tensor core's async proxy
    tc_fence_before();          // this thread's TMEM reads of S are done before the next MMA overwrites it
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t dp0 = umma_desc_k_sw128(smem_u32(smem + SMEM_P));
      const uint64_t dp1 = umma_desc_k_sw128(smem_u32(smem + SMEM_P + TILE_QK_BYTES));
      const uint64_t dv0 = umma_desc_k_sw128(smem_u32(smem + SMEM_V + (st * 2 + 0) * TILE_V_HALF_BYTES));
      const uint64_t dv1 = umma_desc_k_sw128(smem_u32(smem + SMEM_V + (st * 2 + 1) * TILE_V_HALF_BYTES));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + TMEM_O, dp0 + 2 * k, dv0 + 2 * k, idesc_o, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + TMEM_O, dp1 + 2 * k, dv1 + 2 * k, idesc_o, 1);
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, j & 1);
    tc_fence_after();
    if (tid == 0 && j + 2 < a.n_kt) load_kv(j + 2);   // stage st (K_j, V_j) is free again
    __syncwarp();
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(tmem_base + lane_base + TMEM_O + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(raw[i]);
    }
    tc_fence_before();   // O reads done before the next iteration's MMAs (ordered by its __syncthreads)
  }

  // ---- normalise and store merged heads: out[(b*T + t)][h*64 + c]  (permute + cpy, 1924-1929)
  const int t = qt * QT + r;
  if (t < T) {
    const float inv = 1.0f / l_run;
    uint4* dst = reinterpret_cast<uint4*>(a.out + ((long long)b * T + t) * (H * DH) + h * DH);
#pragma unroll
    for (int q8 = 0; q8 < DH / 8; ++q8) {
      uint4 u;
      u.x = pack_h2(o_acc[8 * q8 + 0] * inv, o_acc[8 * q8 + 1] * inv);
      u.y = pack_h2(o_acc[8 * q8 + 2] * inv, o_acc[8 * q8 + 3] * inv);
      u.z = pack_h2(o_acc[8 * q8 + 4] * inv, o_acc[8 * q8 + 5] * inv);
      u.w = pack_h2(o_acc[8 * q8 + 6] * inv, o_acc[8 * q8 + 7] * inv);
      dst[q8] = u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace

bool attention_setup_attributes(const char** err) {
  cudaError_t e = cudaFuncSetAttribute(attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       ATTN_SMEM_BYTES);
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return false;
  }
  return true;
}

cudaError_t launch_attention(const AttnProblem& p, cudaStream_t st) {
  AttnArgs a;
  a.B = p.B;
  a.T = p.T;
  a.H = p.H;
  a.n_kt = (p.T + KT - 1) / KT;
  a.out = p.out;
  a.scale_log2 = p.scale * 1.4426950408889634f;
  dim3 grid((p.T + QT - 1) / QT, p.H, p.B);
  attention_tcgen05_kernel<<<grid, 128, ATTN_SMEM_BYTES, st>>>(p.qk_map, p.vt_map, a);
  return cudaGetLastError();
}

}  // namespace wb
