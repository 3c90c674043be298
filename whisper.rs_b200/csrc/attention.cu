// attention.cu -- fused softmax attention on tcgen05 for sm_100a.
//
// Replaces galois_flash_attn (src/main.rs:1787-1797, call 1922) and the F16 repack / permute /
// merge ops around it (1898-1929): per head h and query n,
//     out[n] = sum_m softmax_m( K[m].Q[n] / sqrt(Dh) ) * V[m],      non-causal, Dh = 64.
// As in the reference (ggml semantics), Q, K, V arrive rounded to F16 and the
// probabilities are F16 in the P.V product; scores, the running max and the exponential's
// argument stay in f32 (the reference rounds the argument to F16 for its lookup table), all
// accumulation is f32.
//
// One CTA = one (segment, head, 256-query block): two softmax warpgroups of 128 threads, each
// owning one 128-query tile (thread r = query row r = TMEM lane r, so the row max needs no
// cross-thread reduction), plus one control warp whose single elected thread issues every TMA
// load and every tcgen05.mma.  Per 128-key tile j and warpgroup:
//   S_j = Q K_j^T          tcgen05.mma 128 x 128 x 64  -> that warpgroup's TMEM columns [0,128)
//   P_j = 2^(c S_j - m)    one TMEM pass into registers; exponentials split 2:1 between the MUFU
//                          unit and an FMA-pipe polynomial (the MUFU alone would take twice the
//                          tile's MMA time), packed to F16 into the swizzled K-major P tile
//   O  += P_j [V_j | 1]    tcgen05.mma 128 x 80 x 128 accumulating in TMEM columns [128,208):
//                          the V^T tile carries a row of ones after the 64 head rows, so column
//                          64 of O is the softmax denominator, accumulated by the tensor core
// The output accumulator never leaves TMEM: the running max is applied lazily -- O is rescaled
// (tcgen05.ld / tcgen05.st) only when a row's max grows by more than 2^8, which after the first
// tiles is rare -- so the per-tile CUDA-core work is max + exp + pack only.
// K_j / V_j stream through double-buffered TMA stages shared by both warpgroups (half the L2
// traffic of one tile per CTA); all hand-offs are mbarriers (S ready, P ready, stage free), so
// while one warpgroup's MMAs run the other warpgroup's softmax keeps the CUDA cores busy.
#include "ptx.cuh"
#include "wb_kernels.hpp"

namespace wb {

namespace {

constexpr int QT = 128;   // queries per softmax warpgroup
constexpr int NWG = 2;    // query tiles (warpgroups) per CTA, sharing every K/V tile
constexpr int KT = 128;   // keys per iteration
constexpr int DH = 64;
constexpr int VROWS = ATTN_VT_HEAD_ROWS;          // 64 head rows + ones row + 15 zero rows
constexpr int TILE_QK_BYTES = QT * DH * 2;        // 16 KB
constexpr int TILE_V_HALF_BYTES = VROWS * 64 * 2; // 10 KB: [80 rows][64 keys]
constexpr int TILE_V_BYTES = 2 * TILE_V_HALF_BYTES;
constexpr int SMEM_Q = 0;                                      // NWG tiles
constexpr int SMEM_K = SMEM_Q + NWG * TILE_QK_BYTES;           // 2 stages
constexpr int SMEM_V = SMEM_K + 2 * TILE_QK_BYTES;             // 2 stages x 2 key halves
constexpr int SMEM_P = SMEM_V + 2 * TILE_V_BYTES;              // NWG x 2 sub-tiles [128][64]
constexpr int SMEM_BAR = SMEM_P + NWG * 2 * TILE_QK_BYTES;
constexpr int ATTN_SMEM_BYTES = SMEM_BAR + 128;
static_assert(SMEM_V % 1024 == 0 && SMEM_P % 1024 == 0 && TILE_V_HALF_BYTES % 1024 == 0, "swizzle alignment");
constexpr int ATTN_THREADS = 32 * (4 * NWG + 1);  // softmax warpgroups + one control warp
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TMEM_WG_STRIDE = 256;          // per warpgroup: S at +0 (128 cols), O at +128 (96 cols)
constexpr uint32_t TMEM_S = 0, TMEM_O = 128;
constexpr float RESCALE_LOG2 = 8.0f;              // lazy rescale threshold: P stays below 2^8

struct AttnArgs {
  int B, T, H, n_kt;
  __half* out;
  float scale_log2;   // scale * log2(e)
  long long* dbg;     // optional: clock64() trace of CTA (0,0,0) (tools/prof_attention.py); nullptr in production
};
#define ATTN_TRACE(slot)                                                          \
  do {                                                                            \
    if (trace) a.dbg[(slot)] = clock64();                                         \
  } while (0)

__global__ void __launch_bounds__(ATTN_THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap qk_map, const __grid_constant__ CUtensorMap vt_map,
                         const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);   // [NWG] Q tile landed
  uint64_t* bar_k = bar_q + NWG;      // [2] K stage landed
  uint64_t* bar_v = bar_k + 2;        // [2] V stage landed
  uint64_t* bar_free = bar_v + 2;     // [2] MMAs of iteration parity retired: its K/V stages are reusable
  uint64_t* bar_s = bar_free + 2;     // [NWG] S_j ready (and O accumulated through tile j-1)
  uint64_t* bar_p = bar_s + NWG;      // [NWG] P_j written, S_j consumed (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_p + NWG);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int h = blockIdx.y, b = blockIdx.z;
  const int H = a.H, T = a.T, n_kt = a.n_kt;
  const int q0 = blockIdx.x * (NWG * QT);   // first query of this CTA

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();   // SWIZZLE_128B tiles need 1024-byte alignment
    for (int i = 0; i < NWG; ++i) {
      mbar_init(&bar_q[i], 1);
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], QT);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_k[i], 1);
      mbar_init(&bar_v[i], 1);
      mbar_init(&bar_free[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 4 * NWG) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4 * NWG) {
    // ===================== control warp: every TMA load and every MMA, one thread =====================
    if ((tid & 31) == 0) {
      const bool trace = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      prefetch_tmap(&qk_map);
      prefetch_tmap(&vt_map);
      const int vrow = (b * H + h) * VROWS;
      auto load_k = [&](int j) {
        const int st = j & 1;
        mbar_arrive_expect_tx(&bar_k[st], TILE_QK_BYTES);
        tma_load_4d(smem + SMEM_K + st * TILE_QK_BYTES, &qk_map, &bar_k[st], 0, H + h, j * KT, b);
      };
      auto load_v = [&](int j) {
        const int st = j & 1;
        mbar_arrive_expect_tx(&bar_v[st], TILE_V_BYTES);
        tma_load_2d(smem + SMEM_V + st * TILE_V_BYTES, &vt_map, &bar_v[st], j * KT, vrow);
        tma_load_2d(smem + SMEM_V + st * TILE_V_BYTES + TILE_V_HALF_BYTES, &vt_map, &bar_v[st], j * KT + 64, vrow);
      };
      constexpr uint32_t idesc_s = umma_idesc_f16(QT, KT);      // 128 x 128
      constexpr uint32_t idesc_o = umma_idesc_f16(QT, VROWS);   // 128 x 80
      auto issue_s = [&](int wg, int j) {   // S_j[wg] = Q[wg] K_j^T
        const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + SMEM_Q + wg * TILE_QK_BYTES));
        const uint64_t dk = umma_desc_k_sw128(smem_u32(smem + SMEM_K + (j & 1) * TILE_QK_BYTES));
        const uint32_t d = tmem_base + wg * TMEM_WG_STRIDE + TMEM_S;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_f16_ss(d, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      };
      auto issue_o = [&](int wg, int j) {   // O[wg] += P_j[wg] [V_j | 1]
        const uint32_t pbase = smem_u32(smem + SMEM_P + wg * 2 * TILE_QK_BYTES);
        const uint32_t vbase = smem_u32(smem + SMEM_V + (j & 1) * TILE_V_BYTES);
        const uint64_t dp0 = umma_desc_k_sw128(pbase), dp1 = umma_desc_k_sw128(pbase + TILE_QK_BYTES);
        const uint64_t dv0 = umma_desc_k_sw128(vbase), dv1 = umma_desc_k_sw128(vbase + TILE_V_HALF_BYTES);
        const uint32_t d = tmem_base + wg * TMEM_WG_STRIDE + TMEM_O;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(d, dp0 + 2 * k, dv0 + 2 * k, idesc_o, (j | k) != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(d, dp1 + 2 * k, dv1 + 2 * k, idesc_o, 1);
      };
      for (int wg = 0; wg < NWG; ++wg) {
        mbar_arrive_expect_tx(&bar_q[wg], TILE_QK_BYTES);
        tma_load_4d(smem + SMEM_Q + wg * TILE_QK_BYTES, &qk_map, &bar_q[wg], 0, h, q0 + wg * QT, b);
      }
      load_k(0);
      load_v(0);
      if (n_kt > 1) {
        load_k(1);
        load_v(1);
      }
      mbar_wait(&bar_k[0], 0);
      for (int wg = 0; wg < NWG; ++wg) {
        mbar_wait(&bar_q[wg], 0);
        tc_fence_after();
        issue_s(wg, 0);
        umma_commit(&bar_s[wg]);
      }
      for (int j = 0; j < n_kt; ++j) {
        // V_{j+1} goes into the stage P_{j-1} V_{j-1} read: reusable once iteration j-1's MMAs retired
        if (j >= 1 && j + 1 < n_kt) {
          mbar_wait(&bar_free[(j - 1) & 1], ((j - 1) >> 1) & 1);
          load_v(j + 1);
        }
        for (int wg = 0; wg < NWG; ++wg) {
          ATTN_TRACE(512 + (j * 2 + wg) * 4 + 0);
          mbar_wait(&bar_p[wg], j & 1);   // P_j[wg] in smem, S_j[wg] consumed
          ATTN_TRACE(512 + (j * 2 + wg) * 4 + 1);
          if (wg == 0) mbar_wait(&bar_v[j & 1], (j >> 1) & 1);
          tc_fence_after();
          ATTN_TRACE(512 + (j * 2 + wg) * 4 + 2);
          issue_o(wg, j);
          if (j + 1 < n_kt) {
            if (wg == 0) {
              mbar_wait(&bar_k[(j + 1) & 1], ((j + 1) >> 1) & 1);
              tc_fence_after();
            }
            issue_s(wg, j + 1);
          }
          umma_commit(&bar_s[wg]);   // phase j+1 of this warpgroup
          ATTN_TRACE(512 + (j * 2 + wg) * 4 + 3);
        }
        umma_commit(&bar_free[j & 1]);
        // both warpgroups have consumed S_j (bar_p), so the MMAs that read K_j retired: refill its stage
        if (j + 2 < n_kt) load_k(j + 2);
      }
    }
  } else {
    // ===================== softmax warpgroups: thread r owns query row r (TMEM lane r) =====================
    const int wg = warp >> 2;
    const int r = tid & (QT - 1);
    const uint32_t twg = tmem_base + wg * TMEM_WG_STRIDE + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t p_row = smem_u32(smem + SMEM_P + wg * 2 * TILE_QK_BYTES) + r * 128;
    const int sw = r & 7;
    const float c = a.scale_log2;
    float m_used = -INFINITY;   // row max (raw score units) the exponent offset currently refers to
    const bool trace = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && r == 0;

    for (int j = 0; j < n_kt; ++j) {
      ATTN_TRACE((j * 2 + wg) * 8 + 0);
      mbar_wait(&bar_s[wg], j & 1);   // S_j ready; O accumulated through tile j-1 and idle
      __syncwarp();
      tc_fence_after();
      ATTN_TRACE((j * 2 + wg) * 8 + 1);
      // ---- S row -> registers (one TMEM pass)
      uint32_t s[KT];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
        uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
        tmem_ld_32x32b_x32(twg + TMEM_S + 0, s0);
        tmem_ld_32x32b_x32(twg + TMEM_S + 32, s1);
        tmem_ld_32x32b_x32(twg + TMEM_S + 64, s2);
        tmem_ld_32x32b_x32(twg + TMEM_S + 96, s3);
        tmem_ld_wait();
      }
      ATTN_TRACE((j * 2 + wg) * 8 + 2);
      const int kbase = j * KT;
      if (kbase + KT > T) {   // CTA-uniform; only the last key tile
#pragma unroll
        for (int i = 0; i < KT; ++i)
          if (kbase + i >= T) s[i] = 0xff800000u;   // -inf: masked key
      }
      float mx4[4] = {__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]), __uint_as_float(s[3])};
#pragma unroll
      for (int i = 4; i < KT; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s[i]));   // 4 independent chains
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // ---- lazy running max: rescale O only when this row's max grew by more than 2^8
      const bool grow = (mx - m_used) * c > RESCALE_LOG2;   // true on the first tile (m_used = -inf)
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        const float alpha = grow ? exp2f((m_used - mx) * c) : 1.0f;
#pragma unroll 1
        for (int cc = 0; cc < 3; ++cc) {   // columns [0,96) of O: 64 head values, denominator, zeros
          uint32_t o[32];
          tmem_ld_32x32b_x32(twg + TMEM_O + cc * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x32(twg + TMEM_O + cc * 32, o);
        }
        tmem_st_wait();
      }
      if (grow) m_used = mx;
      const float moff = m_used * c;
      ATTN_TRACE((j * 2 + wg) * 8 + 3);
      // ---- P = 2^(c s - m c) -> F16 pairs straight into the swizzled K-major tile.  The MUFU unit
      // (16 exp/clk/SM) would need 1024 cycles per 128 x 128 tile, twice the tile's MMA time, so every
      // third exponential is evaluated on the FMA pipe instead (ex2_fma).
#pragma unroll
      for (int q16 = 0; q16 < KT / 8; ++q16) {   // 16-byte chunk q16 = keys [8 q16, 8 q16 + 8)
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i0 = 8 * q16 + 2 * u, i1 = i0 + 1;
          const float x0 = fmaf(__uint_as_float(s[i0]), c, -moff);
          const float x1 = fmaf(__uint_as_float(s[i1]), c, -moff);
          const float e0 = (i0 % 3 == 0) ? ex2_fma(x0) : ex2_mufu(x0);
          const float e1 = (i1 % 3 == 0) ? ex2_fma(x1) : ex2_mufu(x1);
          w[u] = pack_h2(e0, e1);
        }
        const uint32_t addr = p_row + (q16 >> 3) * TILE_QK_BYTES + (((q16 & 7) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                     "r"(w[3])
                     : "memory");
      }
      ATTN_TRACE((j * 2 + wg) * 8 + 4);
      fence_proxy_async_smem();   // P (generic-proxy stores) -> visible to the tensor core's async proxy
      tc_fence_before();          // TMEM reads of S / writes of O ordered before the MMAs the control warp issues
      mbar_arrive(&bar_p[wg]);
      ATTN_TRACE((j * 2 + wg) * 8 + 5);
    }
    mbar_wait(&bar_s[wg], n_kt & 1);
    __syncwarp();
    tc_fence_after();
    // ---- normalise and store merged heads: out[(b*T + t)][h*64 + c]  (permute + cpy, 1924-1929)
    uint32_t o0[32], o1[32], o2[32];
    tmem_ld_32x32b_x32(twg + TMEM_O, o0);
    tmem_ld_32x32b_x32(twg + TMEM_O + 32, o1);
    tmem_ld_32x32b_x32(twg + TMEM_O + 64, o2);
    tmem_ld_wait();
    const int t = q0 + wg * QT + r;
    if (t < T) {
      const float inv = 1.0f / __uint_as_float(o2[0]);   // column 64: sum of the F16 probabilities
      uint4* dst = reinterpret_cast<uint4*>(a.out + ((long long)b * T + t) * (H * DH) + h * DH);
#pragma unroll
      for (int q8 = 0; q8 < 4; ++q8) {
        uint4 u;
        u.x = pack_h2(__uint_as_float(o0[8 * q8 + 0]) * inv, __uint_as_float(o0[8 * q8 + 1]) * inv);
        u.y = pack_h2(__uint_as_float(o0[8 * q8 + 2]) * inv, __uint_as_float(o0[8 * q8 + 3]) * inv);
        u.z = pack_h2(__uint_as_float(o0[8 * q8 + 4]) * inv, __uint_as_float(o0[8 * q8 + 5]) * inv);
        u.w = pack_h2(__uint_as_float(o0[8 * q8 + 6]) * inv, __uint_as_float(o0[8 * q8 + 7]) * inv);
        dst[q8] = u;
      }
#pragma unroll
      for (int q8 = 0; q8 < 4; ++q8) {
        uint4 u;
        u.x = pack_h2(__uint_as_float(o1[8 * q8 + 0]) * inv, __uint_as_float(o1[8 * q8 + 1]) * inv);
        u.y = pack_h2(__uint_as_float(o1[8 * q8 + 2]) * inv, __uint_as_float(o1[8 * q8 + 3]) * inv);
        u.z = pack_h2(__uint_as_float(o1[8 * q8 + 4]) * inv, __uint_as_float(o1[8 * q8 + 5]) * inv);
        u.w = pack_h2(__uint_as_float(o1[8 * q8 + 6]) * inv, __uint_as_float(o1[8 * q8 + 7]) * inv);
        dst[4 + q8] = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4 * NWG) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// rows 64..79 of every head block of the V^T buffer: row 64 = 1.0 (softmax denominator column),
// rows 65..79 = 0.  Written once at allocation; the QKV GEMM only ever writes rows 0..63.
__global__ void vt_init_kernel(__half* vt, int n_heads_total, int Tp) {
  const int hb = blockIdx.y;
  __half* base = vt + ((size_t)hb * VROWS + DH) * Tp;
  const size_t n = (size_t)(VROWS - DH) * Tp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    base[i] = __float2half_rn(i < (size_t)Tp ? 1.0f : 0.0f);
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

cudaError_t launch_vt_init(__half* vt, int n_heads_total, int Tp, cudaStream_t st) {
  if (n_heads_total <= 0) return cudaSuccess;
  vt_init_kernel<<<dim3(8, n_heads_total), 256, 0, st>>>(vt, n_heads_total, Tp);
  return cudaGetLastError();
}

cudaError_t launch_attention(const AttnProblem& p, cudaStream_t st) {
  AttnArgs a;
  a.B = p.B;
  a.T = p.T;
  a.H = p.H;
  a.n_kt = (p.T + KT - 1) / KT;
  a.out = p.out;
  a.scale_log2 = p.scale * 1.4426950408889634f;
  a.dbg = p.dbg;
  dim3 grid((p.T + NWG * QT - 1) / (NWG * QT), p.H, p.B);
  attention_tcgen05_kernel<<<grid, ATTN_THREADS, ATTN_SMEM_BYTES, st>>>(p.qk_map, p.vt_map, a);
  return cudaGetLastError();
}

}  // namespace wb
