// gemm2.cu -- CTA-pair (tcgen05 cta_group::2) GEMM for the encoder's large-M products, sm_100a.
//
//   C[b][m][n] = epilogue( sum_k A[b][m][k] * W[n][k] )      f16 x f16 -> f32 accumulate (TMEM)
//
// Same call sites as gemm.cu (galois_matmul / galois_conv_1d_* of src/main.rs:1834, 1856,
// 1891-1895, 1936, 1956, 1962, 1992, 2013 and the bias / scale / GELU / residual / F16-repack ops
// that follow them); gemm.cu stays as the skinny swap-AB kernel of the decoder.
//
// Why a pair: a 128 x 256 tile per SM reads 48 KB of operands per 64-deep k-block, 96 B/clk of
// L2->SM and shared-memory bandwidth at the tensor core's rate -- measured, that (and a
// row-per-thread epilogue) capped the single-CTA kernel near 60 % tensor-pipe activity.  Here
// two CTAs of one cluster own a 256 x BN tile: each loads its own 128 A rows and HALF of the W
// tile (32 KB per k-block), one thread of the leader issues tcgen05.mma.cta_group::2 (M = 256),
// and each CTA's TMEM receives its 128 accumulator rows.  The smaller stages leave shared memory
// for a TMA-store epilogue.
//
// Roles per CTA (384 threads):
//   warp 0      TMA producer (own A rows + own half of W; bytes counted on the LEADER's barrier).  Measured and
//               dropped (round 2): cp.async.bulk.prefetch.tensor of the A boxes 8 k-blocks ahead of the ring, to turn
//               fc2's HBM-streamed A operand into L2 hits -- fc2 732 -> 779 us at whisper medium, 64 segments (every
//               pair working on the same rows issues the same prefetches), all GEMMs 2-6 % slower with it everywhere.
//   warp 1      MMA issuer (leader CTA only); tcgen05.commit multicasts "stage free" and
//               "accumulator ready" to both CTAs
//   warp 2      TMEM allocator (2 x BN columns: the epilogue of tile i overlaps tile i+1's MMAs)
//   warps 4-11  epilogue: tcgen05.ld (thread = row) -> bias / column scale / GELU / residual ->
//               128-byte-swizzled shared-memory box -> cp.async.bulk.tensor store, so every
//               global write (and the residual read, a TMA load prefetched one chunk ahead) is
//               a full coalesced box and rows/columns past the tensor edge are clipped by the
//               TMA unit.  The V^T scatter of the QKV projection keeps direct stores (lanes hold
//               consecutive time steps, so those are coalesced already).
#include "ptx.cuh"
#include "wb_kernels.hpp"

namespace wb {

namespace {

constexpr int BM = 128;            // accumulator rows per CTA (pair tile: 256)
constexpr int BK = 64;             // one 128-byte swizzle atom of f16
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;       // two warps per TMEM lane quarter, interleaved over column chunks
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);
constexpr int EPI_BUF_BYTES = 32 * 128;   // 32 rows x 128 bytes (32 f32 or 64 f16 columns)
constexpr int NBUF = 2;
constexpr int STAGES = 4;

struct Gemm2Args {
  int M_rows, batch, N, K;
  int m_pairs, n_tiles, total_items;
  const float* bias;       // [N] or null
  const float* colscale;   // [N] or null
  float scale;
  int gelu;
  int out_f16;
  int has_res;             // f32 residual tile fetched through res_map, added last
  int res_bcast;           // residual has no batch dimension (positional embedding)
  int out_slab_cols;       // > 0: output column n lives in slab n / C at column n % C (out_map's third dimension)
  __half* vt_out;          // columns >= vt_col0 go to the transposed V buffer (see GemmEpilogue)
  int vt_col0, vt_heads, vt_head_rows, vt_ld, vt_T;
  // LayerNorm folded into the GEMMs around it (see the LN template parameter)
  float2* ln_part_out;         // LN == 1: [row][ln_parts] partial (sum, sum of squares) of (x - center[row]), one slot
                               //          per (N tile, epilogue half): plain stores, summed in slot order by the consumer
  __half* x16_out;             // LN == 1: F16 copy of x - center[row], [row][x16_ld] (the next GEMM's A operand)
  int x16_ld;
  const float2* ln_part_in;    // LN == 2: the partial statistics of this GEMM's A rows
  int ln_parts;                // partials per row
  float* ln_center;            // [row] per-row centre: read by LN == 1, advanced by the row mean by LN == 2 (N tile 0)
  float ln_inv_d, ln_eps;
  long long* dbg;          // optional clock64() trace (CTA 0): [0,256) MMA warp, [256,1024) epilogue warp 4
};
#define G2_TRACE(cond, slot)                                   \
  do {                                                         \
    if ((cond) && (slot) < 1024) args.dbg[(slot)] = clock64(); \
  } while (0)

template <int BN>
struct Cfg {
  static constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_HALF_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int EPI_BYTES = EPI_WARPS * NBUF * EPI_BUF_BYTES;
  static constexpr int BIAS_BYTES = EPI_WARPS * 2 * BN * 4;   // per warp: bias[BN], colscale[BN]
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = RING_BYTES + EPI_BYTES + BIAS_BYTES + BAR_BYTES + 1024 /*align slack*/;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static_assert(B_HALF_BYTES % 1024 == 0, "W half tile must keep the 1024-byte swizzle alignment");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB a CTA may use");
};

struct Item {
  int b, mp, nt;
};
__device__ __forceinline__ Item item_coord(const Gemm2Args& a, int item) {
  const int per_batch = a.m_pairs * a.n_tiles;
  Item t;
  t.b = item / per_batch;
  const int r = item - t.b * per_batch;
  t.mp = r / a.n_tiles;   // n fastest: pairs running concurrently share A rows through L2
  t.nt = r - t.mp * a.n_tiles;
  return t;
}

// Epilogue variants are compile-time (F16 or f32 output, GELU, column scale, f32 residual): as run-time
// flags the compiler predicated the unused paths off instruction by instruction -- the epilogue warps
// still issued them, and at K = 512 the epilogue, not the tensor pipe, set the tile rate.
//
// LN: LayerNorm (galois_norm + repeat/mul/add, src/main.rs:1781-1785, 1882-1886, 1948-1952) folded into the two
// GEMMs around it, so the normalised activations never make a round trip through HBM:
//   LN == 1 (producer, with RES): the epilogue that writes the f32 residual stream x also writes an F16 copy of
//           x - c, c = center[row] (the row's mean at the previous LayerNorm: the residual stream moves slowly, so
//           x - c is nearly centred -- rounding x itself to F16 would lose the deviations of a row whose mean is
//           large against its spread, and sum(x^2) - sum(x)^2 / d would cancel), and leaves the sum and the sum of
//           squares of x - c over its share of the row in its own slot (thread = row; no atomics: the result
//           does not depend on the order the tiles finish in);
//   LN == 2 (consumer): A is that F16 copy and W' = W diag(gamma); the slots are added in slot order to mu' (the
//           mean of x - c) and rstd, and
//           LN(x) W^T + b = rstd * ((x - c) W'^T - mu' * c1) + c2,  c1[n] = sum_k W'[n][k],  c2[n] = sum_k W[n][k] beta[k] + b[n]
//           (c1 arrives in the column-scale slot, c2 in the bias slot); the CTA of N tile 0 moves center[row] on by mu'.
template <int BN, bool F16O, bool GELU, bool CS, bool RES, int LN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm2_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap w_map,
                         const __grid_constant__ CUtensorMap out_map, const __grid_constant__ CUtensorMap res_map,
                         const Gemm2Args args) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_bufs = smem + C::RING_BYTES;
  float* bias_stage = reinterpret_cast<float*>(epi_bufs + C::EPI_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_stage) + C::BIAS_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]  (the leader's copy is the one in use)
  uint64_t* res_bar = tmem_empty + 2;           // [EPI_WARPS][NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + EPI_WARPS * NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&a_map);
    prefetch_tmap(&w_map);
    prefetch_tmap(&out_map);
    if (RES) prefetch_tmap(&res_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);    // the leader's arrive.expect_tx; both CTAs' TMA bytes land here
      mbar_init(&empty_bar[s], 1);   // one multicast tcgen05.commit from the leader
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * EPI_WARPS);   // every epilogue warp of both CTAs
    }
    for (int i = 0; i < EPI_WARPS * NBUF; ++i) mbar_init(&res_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_cg2<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();   // both CTAs' barriers and TMEM exist before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();

  const int num_kb = (args.K + BK - 1) / BK;
  const int item0 = blockIdx.x >> 1, item_step = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = item0; item < args.total_items; item += item_step) {
        const Item it = item_coord(args, item);
        const int m0 = (it.mp * 2 + (int)rank) * BM;
        const int n0 = it.nt * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);   // the pair's MMAs have drained this stage
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          tma_load_3d_cg2(sa, &a_map, full_leader, kb * BK, m0, it.b);
          tma_load_2d_cg2(sa + A_BYTES, &w_map, full_leader, kb * BK, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; the warp stays converged, one elected lane issues) =====================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      for (int item = item0; item < args.total_items; item += item_step, ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (t >> 1) & 1;
        const bool trm = args.dbg != nullptr && blockIdx.x == 0 && lane == 0;
        G2_TRACE(trm, t * 4 + 0);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        G2_TRACE(trm, t * 4 + 1);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16_ss_cg2_elect(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit_cg2_elect(&empty_bar[stage], 3);
          if (kb == num_kb - 1) {
            umma_commit_cg2_elect(&tmem_full[acc], 3);
            G2_TRACE(trm, t * 4 + 2);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue (both CTAs) =====================
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int half = ew >> 2;        // which interleaved half of the column chunks
    uint8_t* my_bufs = epi_bufs + ew * NBUF * EPI_BUF_BYTES;
    uint64_t* my_res_bar = res_bar + ew * NBUF;
    float* my_bias = bias_stage + ew * 2 * BN;
    float* my_cs = my_bias + BN;
    constexpr bool f16o = F16O;
    constexpr int cw = f16o ? 64 : 32;                   // chunk width in columns (128 bytes of output)
    const int n_chunks_tile = BN / cw;
    const int my_nch = (n_chunks_tile - half + 1) / 2;   // chunks half, half+2, ...
    const uint32_t tmem_empty_leader = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    constexpr bool has_cs = CS || LN == 2;   // LN == 2: the column-scale slot carries c1
    const int sw = lane & 7;

    // coordinates of this warp's chunk number `ci` (counted over all of its tiles)
    auto chunk_coord = [&](int ci, int& n0, int& m0, int& b) -> bool {
      const int t = ci / my_nch, k = ci - t * my_nch;
      const int item = item0 + t * item_step;
      if (item >= args.total_items) return false;
      const Item it = item_coord(args, item);
      n0 = it.nt * BN + (half + 2 * k) * cw;
      m0 = (it.mp * 2 + (int)rank) * BM + q * 32;
      b = it.b;
      return true;
    };
    auto issue_residual = [&](int ci) {   // lane 0: TMA-load the f32 residual box of chunk ci into its buffer
      int n0, m0, b;
      if (!chunk_coord(ci, n0, m0, b)) return;
      if (n0 >= args.N || n0 >= args.vt_col0) return;   // nothing will consume it (see the chunk loop)
      const int buf = ci & (NBUF - 1);
      mbar_arrive_expect_tx(&my_res_bar[buf], EPI_BUF_BYTES);
      tma_load_3d(my_bufs + buf * EPI_BUF_BYTES, &res_map, &my_res_bar[buf], n0, m0, args.res_bcast ? 0 : b);
    };

    int ci = 0;   // chunks this warp has started
    if (RES && my_nch > 0 && lane == 0) issue_residual(0);
    int t = 0;
    for (int item = item0; item < args.total_items; item += item_step, ++t) {
      const Item it = item_coord(args, item);
      const int acc = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      const int m_row0 = (it.mp * 2 + (int)rank) * BM + q * 32;
      // ---- this tile's bias / column scale -> registers (issued before the accumulator wait)
      constexpr int NV = (BN + 127) / 128;
      float4 bb[NV], cc[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int j = v * 128 + lane * 4;
        bb[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        cc[v] = make_float4(args.scale, args.scale, args.scale, args.scale);
        if (j < BN && it.nt * BN + j < args.N) {
          if (args.bias) bb[v] = __ldg(reinterpret_cast<const float4*>(args.bias + it.nt * BN + j));
          if (args.colscale) {
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(args.colscale + it.nt * BN + j));
            cc[v] = make_float4(c4.x * args.scale, c4.y * args.scale, c4.z * args.scale, c4.w * args.scale);
          }
        }
      }
      // LN fold: this thread's row of the tile (global row: the batch index is the slowest dimension); the
      // consumer's row statistics are fetched here, under the accumulator wait, like the bias
      const int ln_m = m_row0 + lane;
      const long long ln_row = (long long)it.b * args.M_rows + ln_m;
      float2 ln_st = make_float2(0.0f, 0.0f);
      float ln_c = 0.0f;
      if (LN == 2 && ln_m < args.M_rows) {
        // the row's partial sums, two slots per 16-byte load, every load in flight before the first is used (one L2
        // round trip under the accumulator wait, as the bias), added in slot order
        constexpr int MAXP2 = 5;   // ln_parts = 2 * (d / tile width) <= 10 for d <= 1280
        const float4* pp = reinterpret_cast<const float4*>(args.ln_part_in + ln_row * args.ln_parts);
        const int p2 = args.ln_parts >> 1;
        float4 pv[MAXP2];
#pragma unroll
        for (int p = 0; p < MAXP2; ++p) pv[p] = p < p2 ? __ldcg(pp + p) : make_float4(0.f, 0.f, 0.f, 0.f);   // L2: written by the previous kernel
#pragma unroll
        for (int p = 0; p < MAXP2; ++p) {
          ln_st.x += pv[p].x;
          ln_st.y += pv[p].y;
          ln_st.x += pv[p].z;
          ln_st.y += pv[p].w;
        }
      }
      if (LN == 1 && ln_m < args.M_rows) ln_c = __ldcg(args.ln_center + ln_row);
      const bool tre = args.dbg != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0;
      G2_TRACE(tre, 256 + t * 16 + 0);
      mbar_wait(&tmem_full[acc], acc_phase);
      G2_TRACE(tre, 256 + t * 16 + 1);
      __syncwarp();   // tcgen05.ld is warp-collective: reconverge after the spin
      tc_fence_after();
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int j = v * 128 + lane * 4;
        if (j < BN) {
          *reinterpret_cast<float4*>(my_bias + j) = bb[v];
          *reinterpret_cast<float4*>(my_cs + j) = cc[v];
        }
      }
      __syncwarp();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      float ln_rstd = 0.0f, ln_nmr = 0.0f;   // LN == 2: rstd and -mu * rstd of this row
      float ln_s1 = 0.0f, ln_s2 = 0.0f;      // LN == 1: this thread's share of its row's statistics
      if (LN == 2) {
        const float mu = ln_st.x * args.ln_inv_d;
        const float var = fmaxf(ln_st.y * args.ln_inv_d - mu * mu, 0.0f);
        ln_rstd = rsqrtf(var + args.ln_eps);
        ln_nmr = -mu * ln_rstd;
        if (it.nt == 0 && half == 0 && ln_m < args.M_rows) args.ln_center[ln_row] += mu;   // the next producer's centre
      }
      const uint64_t ln_nc2 = f2_pack(-ln_c, -ln_c);
      const uint64_t ln_rstd2 = f2_pack(ln_rstd, ln_rstd), ln_nmr2 = f2_pack(ln_nmr, ln_nmr);
      uint64_t ln_s1p = 0, ln_s2p = 0;       // LN == 1: two-lane partial sums (FADD2 / FFMA2)
      if (my_nch == 0) {   // narrow tile: this warp has no chunk, but its arrival is counted
        tc_fence_before();
        if (lane == 0) mbar_arrive_cluster(tmem_empty_leader + acc * 8);
      }

      for (int k = 0; k < my_nch; ++k, ++ci) {
        const int c = half + 2 * k;            // chunk index inside the tile
        const int col = c * cw;                // first accumulator column of the chunk
        const int n0 = it.nt * BN + col;
        const bool last_chunk = (k == my_nch - 1);
        const bool dead = n0 >= args.N;        // ragged N: tile columns past the matrix
        const bool to_vt = !dead && n0 >= args.vt_col0;
        const int buf = ci & (NBUF - 1);
        uint8_t* sbuf = my_bufs + buf * EPI_BUF_BYTES;
        const uint32_t srow = smem_u32(sbuf) + lane * 128;

        if (RES && lane == 0) {
          // request the NEXT chunk's residual box now, a whole chunk ahead: its buffer was last read by the TMA store
          // of chunk ci - 1, issued a moment ago (requesting it only after this chunk's store, as round 1 did, left
          // ~400 cycles of lead against a ~1500-cycle L2 / HBM round trip -- the out-proj epilogue, 4 chunks per warp
          // per 8192-cycle tile, was bound by that wait)
          bulk_wait_group_read<0>();
          issue_residual(ci + 1);
        }

        // ---- accumulator chunk -> registers
        uint32_t r0[32], r1[32];
        if (!dead) {
          tmem_ld_32x32b_x32(t_row + col, r0);
          if (f16o) tmem_ld_32x32b_x32(t_row + col + 32, r1);
          tmem_ld_wait();
        }
        G2_TRACE(tre, 256 + t * 16 + 2 + k * 4);   // chunk k: accumulator in registers
        if (last_chunk) {   // this warp is done with the accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tmem_empty_leader + acc * 8);
        }
        if (dead) continue;

        // ---- bias, column scale, GELU (thread = row, registers = columns)
        {
          const float4* b4 = reinterpret_cast<const float4*>(my_bias + col);
          const float4* c4 = reinterpret_cast<const float4*>(my_cs + col);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 bv = b4[g];
            float x0, x1, x2, x3;
            if (LN == 2) {   // rstd * acc + (c2 - mu * rstd * c1)
              const float4 cv = c4[g];   // packed FFMA2: two columns per instruction
              const uint64_t t01 = f2_fma(ln_nmr2, f2_pack(cv.x, cv.y), f2_pack(bv.x, bv.y));
              const uint64_t t23 = f2_fma(ln_nmr2, f2_pack(cv.z, cv.w), f2_pack(bv.z, bv.w));
              f2_unpack(f2_fma(f2_pack(__uint_as_float(r0[4 * g]), __uint_as_float(r0[4 * g + 1])), ln_rstd2, t01), x0, x1);
              f2_unpack(f2_fma(f2_pack(__uint_as_float(r0[4 * g + 2]), __uint_as_float(r0[4 * g + 3])), ln_rstd2, t23), x2, x3);
            } else {
              x0 = __uint_as_float(r0[4 * g]) + bv.x; x1 = __uint_as_float(r0[4 * g + 1]) + bv.y;
              x2 = __uint_as_float(r0[4 * g + 2]) + bv.z; x3 = __uint_as_float(r0[4 * g + 3]) + bv.w;
              if (has_cs) {
                const float4 cv = c4[g];
                x0 *= cv.x; x1 *= cv.y; x2 *= cv.z; x3 *= cv.w;
              }
            }
            if (GELU) {
              gelu_f16in_x2(x0, x1);
              gelu_f16in_x2(x2, x3);
            }
            r0[4 * g] = __float_as_uint(x0); r0[4 * g + 1] = __float_as_uint(x1);
            r0[4 * g + 2] = __float_as_uint(x2); r0[4 * g + 3] = __float_as_uint(x3);
          }
          if (f16o) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 bv = b4[8 + g];
              float x0, x1, x2, x3;
              if (LN == 2) {
                const float4 cv = c4[8 + g];
                const uint64_t t01 = f2_fma(ln_nmr2, f2_pack(cv.x, cv.y), f2_pack(bv.x, bv.y));
                const uint64_t t23 = f2_fma(ln_nmr2, f2_pack(cv.z, cv.w), f2_pack(bv.z, bv.w));
                f2_unpack(f2_fma(f2_pack(__uint_as_float(r1[4 * g]), __uint_as_float(r1[4 * g + 1])), ln_rstd2, t01), x0, x1);
                f2_unpack(f2_fma(f2_pack(__uint_as_float(r1[4 * g + 2]), __uint_as_float(r1[4 * g + 3])), ln_rstd2, t23), x2, x3);
              } else {
                x0 = __uint_as_float(r1[4 * g]) + bv.x; x1 = __uint_as_float(r1[4 * g + 1]) + bv.y;
                x2 = __uint_as_float(r1[4 * g + 2]) + bv.z; x3 = __uint_as_float(r1[4 * g + 3]) + bv.w;
                if (has_cs) {
                  const float4 cv = c4[8 + g];
                  x0 *= cv.x; x1 *= cv.y; x2 *= cv.z; x3 *= cv.w;
                }
              }
              if (GELU) {
                gelu_f16in_x2(x0, x1);
              gelu_f16in_x2(x2, x3);
              }
              r1[4 * g] = __float_as_uint(x0); r1[4 * g + 1] = __float_as_uint(x1);
              r1[4 * g + 2] = __float_as_uint(x2); r1[4 * g + 3] = __float_as_uint(x3);
            }
          }
        }

        if (to_vt) {
          // V^T scatter: time contiguous (reference layout [T, Dh, H], src/main.rs:1914-1920);
          // lanes hold consecutive time steps, so each store instruction writes 64 contiguous bytes
          const int m = m_row0 + lane;
          if (m < args.M_rows) {
            const int seg = m / args.vt_T;
            const int tt = m - seg * args.vt_T;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int nn = n0 - args.vt_col0 + 32 * hh;   // multiple of 32: inside one head
              __half* dst = args.vt_out +
                            ((long long)(seg * args.vt_heads + (nn >> 6)) * args.vt_head_rows + (nn & 63)) * args.vt_ld + tt;
              const uint32_t* rr = hh ? r1 : r0;
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[(long long)j * args.vt_ld] = __float2half_rn(__uint_as_float(rr[j]));
            }
          }
          continue;
        }

        if (RES) {
          // ---- + residual: its box was TMA-loaded into this buffer one chunk ago
          mbar_wait(&my_res_bar[buf], (ci / NBUF) & 1);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 x;
            const uint32_t addr = srow + ((g ^ sw) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                         : "r"(addr)
                         : "memory");
            x.x += __uint_as_float(r0[4 * g]); x.y += __uint_as_float(r0[4 * g + 1]);
            x.z += __uint_as_float(r0[4 * g + 2]); x.w += __uint_as_float(r0[4 * g + 3]);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w)
                         : "memory");
            if (LN == 1) {   // statistics of the centred f32 values; their F16 copy replaces the accumulator registers
              const uint64_t xa = f2_add(f2_pack(x.x, x.y), ln_nc2), xb = f2_add(f2_pack(x.z, x.w), ln_nc2);
              ln_s1p = f2_add(ln_s1p, f2_add(xa, xb));
              ln_s2p = f2_fma(xa, xa, f2_fma(xb, xb, ln_s2p));
              float c0, c1, c2, c3;
              f2_unpack(xa, c0, c1);
              f2_unpack(xb, c2, c3);
              r0[2 * g] = pack_h2(c0, c1);
              r0[2 * g + 1] = pack_h2(c2, c3);
            }
          }
          if (LN == 1 && ln_m < args.M_rows) {   // 32 columns = 64 bytes of this row (two full sectors)
            uint4* dst = reinterpret_cast<uint4*>(args.x16_out + ln_row * args.x16_ld + n0);
#pragma unroll
            for (int v = 0; v < 4; ++v) dst[v] = make_uint4(r0[4 * v], r0[4 * v + 1], r0[4 * v + 2], r0[4 * v + 3]);
          }
        } else {
          // the store that last read this buffer (chunk ci - NBUF) must have finished reading it
          if (lane == 0) bulk_wait_group_read<NBUF - 1>();
          __syncwarp();
          if (f16o) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {   // 16-byte unit g = columns 8g .. 8g+7 of the 64-column chunk
              const uint32_t* rr = (g < 4) ? r0 : r1;
              const int o = (g & 3) * 8;
              const uint32_t w0 = pack_h2(__uint_as_float(rr[o]), __uint_as_float(rr[o + 1]));
              const uint32_t w1 = pack_h2(__uint_as_float(rr[o + 2]), __uint_as_float(rr[o + 3]));
              const uint32_t w2 = pack_h2(__uint_as_float(rr[o + 4]), __uint_as_float(rr[o + 5]));
              const uint32_t w3 = pack_h2(__uint_as_float(rr[o + 6]), __uint_as_float(rr[o + 7]));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((g ^ sw) << 4)), "r"(w0), "r"(w1),
                           "r"(w2), "r"(w3)
                           : "memory");
            }
          } else {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((g ^ sw) << 4)), "r"(r0[4 * g]),
                           "r"(r0[4 * g + 1]), "r"(r0[4 * g + 2]), "r"(r0[4 * g + 3])
                           : "memory");
          }
        }
        G2_TRACE(tre, 256 + t * 16 + 3 + k * 4);   // math + shared-memory box written
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        G2_TRACE(tre, 256 + t * 16 + 4 + k * 4);   // fenced
        if (lane == 0) {
          if (args.out_slab_cols > 0) tma_store_3d(&out_map, sbuf, n0 % args.out_slab_cols, m_row0, n0 / args.out_slab_cols);
          else tma_store_3d(&out_map, sbuf, n0, m_row0, it.b);
          bulk_commit_group();
        }
        __syncwarp();
        G2_TRACE(tre, 256 + t * 16 + 5 + k * 4);   // store issued
      }
      if (LN == 1 && my_nch > 0 && ln_m < args.M_rows) {
        float a0, a1, b0, b1;
        f2_unpack(ln_s1p, a0, a1);
        f2_unpack(ln_s2p, b0, b1);
        args.ln_part_out[ln_row * args.ln_parts + it.nt * 2 + half] = make_float2(ln_s1 + a0 + a1, ln_s2 + b0 + b1);
      }
    }
    if (lane == 0) bulk_wait_group_read<0>();   // shared memory must outlive the stores' reads
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA leaves while its peer may still read its shared memory / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool F16O, bool GELU, bool CS, bool RES, int LN>
cudaError_t launch_one(const GemmProblem& g, const Gemm2Args& a, int grid, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg<BN>::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue under the previous kernel's tail
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  // the kernel's opt-in to > 48 KB of dynamic shared memory, once per instantiation
  static cudaError_t attr_err = cudaFuncSetAttribute(gemm2_f16_tcgen05_kernel<BN, F16O, GELU, CS, RES, LN>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES);
  if (attr_err != cudaSuccess) return attr_err;
  return cudaLaunchKernelEx(&cfg, gemm2_f16_tcgen05_kernel<BN, F16O, GELU, CS, RES, LN>, g.a_map, g.w_map, *g.out_map,
                            g.res_map ? *g.res_map : *g.out_map, a);
}

template <int BN>
cudaError_t launch_bn(const GemmProblem& g, const Gemm2Args& a, int grid, cudaStream_t st) {
  const bool f16 = a.out_f16 != 0, gelu = a.gelu != 0, cs = a.colscale != nullptr || a.scale != 1.0f, res = a.has_res != 0;
  if (f16 && res) return cudaErrorInvalidValue;
  const int ln = a.ln_part_out ? 1 : a.ln_part_in ? 2 : 0;
  if (ln == 1) {   // producer of the folded LayerNorm: f32 residual stream out
    if (f16 || !res || cs) return cudaErrorInvalidValue;
    return gelu ? launch_one<BN, false, true, false, true, 1>(g, a, grid, st)
                : launch_one<BN, false, false, false, true, 1>(g, a, grid, st);
  }
  if (ln == 2) {   // consumer: F16 out, c1 in the column-scale slot
    if (!f16 || res || !a.colscale || a.scale != 1.0f) return cudaErrorInvalidValue;
    return gelu ? launch_one<BN, true, true, false, false, 2>(g, a, grid, st)
                : launch_one<BN, true, false, false, false, 2>(g, a, grid, st);
  }
#define WB_G2_CASE(F, G, C, R) \
  if (f16 == F && gelu == G && cs == C && res == R) return launch_one<BN, F, G, C, R, 0>(g, a, grid, st)
  WB_G2_CASE(true, false, false, false);
  WB_G2_CASE(true, true, false, false);
  WB_G2_CASE(true, false, true, false);
  WB_G2_CASE(true, true, true, false);
  WB_G2_CASE(false, false, false, false);
  WB_G2_CASE(false, true, false, false);
  WB_G2_CASE(false, false, true, false);
  WB_G2_CASE(false, true, true, false);
  WB_G2_CASE(false, false, false, true);
  WB_G2_CASE(false, true, false, true);
  WB_G2_CASE(false, false, true, true);
  WB_G2_CASE(false, true, true, true);
#undef WB_G2_CASE
  return cudaErrorInvalidValue;
}

}  // namespace

// pair tile width: widest of 256 / 192 / 128 / 64 that divides N; 0 = not eligible (use gemm.cu)
int gemm2_pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 0;
}

cudaError_t launch_gemm2(const GemmProblem& g, int num_sms, cudaStream_t st) {
  if (!g.out_map || g.epi.transpose_out || gemm2_pick_bn(g.N) != g.bn) return cudaErrorInvalidValue;
  Gemm2Args a;
  a.M_rows = g.M_rows;
  a.batch = g.batch;
  a.N = g.N;
  a.K = g.K;
  a.m_pairs = (g.M_rows + 2 * BM - 1) / (2 * BM);
  a.n_tiles = (g.N + g.bn - 1) / g.bn;
  a.total_items = a.m_pairs * a.n_tiles * g.batch;
  a.bias = g.epi.bias;
  a.colscale = g.epi.colscale;
  a.scale = g.epi.scale;
  a.gelu = g.epi.gelu;
  a.out_f16 = g.epi.out_f16;
  a.has_res = g.res_map != nullptr;
  a.res_bcast = g.res_bcast;
  a.out_slab_cols = g.epi.out_slab_cols;
  a.vt_out = g.epi.vt_out;
  a.vt_col0 = g.epi.vt_col0;
  a.vt_heads = g.epi.vt_heads;
  a.vt_head_rows = g.epi.vt_head_rows;
  a.vt_ld = g.epi.vt_ld;
  a.vt_T = g.epi.vt_T;
  a.ln_part_out = g.epi.ln_part_out;
  a.x16_out = g.epi.x16_out;
  a.x16_ld = g.epi.x16_ld;
  a.ln_part_in = g.epi.ln_part_in;
  a.ln_parts = g.epi.ln_parts;
  a.ln_center = g.epi.ln_center;
  if (a.ln_part_out && a.ln_parts != 2 * a.n_tiles) return cudaErrorInvalidValue;   // one slot per (N tile, epilogue half)
  if ((a.ln_part_out || a.ln_part_in) && (!a.ln_center || a.ln_parts < 2 || a.ln_parts > 10 || (a.ln_parts & 1)))
    return cudaErrorInvalidValue;
  a.ln_inv_d = g.K > 0 ? 1.0f / (float)g.K : 0.0f;
  a.ln_eps = g.epi.ln_eps;
  a.dbg = g.dbg;
  if (a.total_items <= 0) return cudaSuccess;
  const int max_pairs = num_sms / 2;
  const int grid = 2 * (a.total_items < max_pairs ? a.total_items : max_pairs);
  switch (g.bn) {
    case 256: return launch_bn<256>(g, a, grid, st);
    case 192: return launch_bn<192>(g, a, grid, st);
    case 128: return launch_bn<128>(g, a, grid, st);
    case 64: return launch_bn<64>(g, a, grid, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace wb
