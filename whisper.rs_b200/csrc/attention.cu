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
// One CTA = one (segment, head, 128-query tile), TWO CTAs per SM (each allocates 256 of the SM's 512
// TMEM columns and 88 KB of shared memory): 4 softmax warps (thread r = query row r = TMEM lane r, so
// the row max needs no cross-thread reduction), one warp that issues the tcgen05.mma instructions
// (converged; one elected lane per instruction) and one warp whose elected thread issues the TMA
// loads.  Two independent CTAs per SM instead of two warpgroups in one CTA: each CTA's prologue
// (TMEM allocation, first loads) and tail hide under the other's steady state, the two never run
// in lock-step on the MUFU unit, and 1536 small CTAs pack the 148 SMs better than 768 big ones
// (measured 475 -> 520 TFLOP/s).  Keys advance in steps of 64; per step j:
//   S_j = Q K_j^T          tcgen05.mma 128 x 64 x 64 into one of TWO S buffers in TMEM, issued two
//                          steps ahead, so the next scores are already there when a softmax ends
//   P_j = 2^(c S_j - m)    one TMEM pass into registers (3-input FMNMX max); exponentials split
//                          11 : 5 between the MUFU unit (16 / clk / SM -- alone it would take about
//                          twice the step's MMA time) and an FMA-pipe polynomial; packed to F16 and
//                          written BACK INTO TMEM over the S buffer just consumed (tcgen05.st)
//   O  += P_j [V_j | 1]    tcgen05.mma 128 x 80 x 64 with the A operand read from TMEM (no shared-
//                          memory round trip for P); the V^T tile carries a row of ones after the 64
//                          head rows, so column 64 of O is the softmax denominator, accumulated by
//                          the tensor core from the same F16 probabilities
// The output accumulator never leaves TMEM: the running max is applied lazily -- O is rescaled
// (tcgen05.ld / tcgen05.st, after waiting for the previous P.V to retire) only when a row's max
// grows by more than 2^8, which after the first steps is rare.
// K / V^T arrive as 128-key stages (a stage serves two steps) through a 2-deep TMA ring; every
// hand-off is an mbarrier (S ready x2, P ready x2, O idle, stage full / free).  Measured constraints
// behind the shape (tools/ubench): a tcgen05.mma of M = 128 costs max(~47, N/2) cycles, so the 64-
// and 80-wide tiles run at the instruction floor; the issuing thread is blocked while its MMA
// executes; an mbarrier wait costs ~175 cycles even when already complete; FMNMX3 costs the same as
// FMNMX; ex2.approx.f16x2 gives no MUFU throughput over the f32 form (SASS: two MUFU.EX2.F16 plus a PRMT to merge
// the halves -- five issue slots per key pair against four).  Variants measured slower or
// equal on B200: 128-key tiles with P handed over in halves, 16 softmax warps with half a row per
// thread, software-pipelining the next step's TMEM load and max under the exponentials, two
// warpgroups per CTA with one or two MMA warps (with and without a phase offset).
#include <stdlib.h>

#include "ptx.cuh"
#include "wb_kernels.hpp"

namespace wb {

namespace {

constexpr int QT = 128;   // queries per softmax warpgroup
constexpr int NWG = 1;    // query tiles (softmax warpgroups) per CTA; two CTAs share an SM instead
constexpr int KT = 128;   // keys per TMA stage
constexpr int KS = 64;    // keys per step
constexpr int DH = 64;
constexpr int NS = 2;     // K/V stages (104 KB of shared memory per CTA would not fit twice with 3)
constexpr int VROWS = ATTN_VT_HEAD_ROWS;          // 64 head rows + ones row + 15 zero rows
constexpr int TILE_QK_BYTES = QT * DH * 2;        // 16 KB: 128 rows of 128 bytes
constexpr int TILE_V_HALF_BYTES = VROWS * 64 * 2; // 10 KB: [80 rows][64 keys]
constexpr int TILE_V_BYTES = 2 * TILE_V_HALF_BYTES;
constexpr int SMEM_Q = 0;                                      // NWG tiles
constexpr int SMEM_K = SMEM_Q + NWG * TILE_QK_BYTES;           // NS stages
constexpr int SMEM_V = SMEM_K + NS * TILE_QK_BYTES;            // NS stages x 2 key halves
constexpr int SMEM_BAR = SMEM_V + NS * TILE_V_BYTES;
constexpr int ATTN_SMEM_BYTES = SMEM_BAR + 256;
static_assert(SMEM_V % 1024 == 0 && TILE_V_HALF_BYTES % 1024 == 0, "swizzle alignment");
// after the 8 softmax warps: one MMA-issuing warp PER warpgroup (a thread that issues tcgen05.mma is
// blocked while the instruction executes, and every mbarrier wait costs ~175 cycles even when already
// complete: with a single issuing warp those latencies serialised with the other warpgroup's MMAs and
// set the kernel's pace), then the TMA warp
constexpr int MMA_WARP = 4 * NWG, TMA_WARP = 5 * NWG;
constexpr int ATTN_THREADS = 32 * (5 * NWG + 1);
constexpr uint32_t TMEM_COLS = 256;   // S0 | S1 | O: two CTAs fit the SM's 512 columns
constexpr uint32_t TMEM_WG_STRIDE = 256;          // per warpgroup: S buffers at +0 and +64 (P over them), O at +128
constexpr uint32_t TMEM_S = 0, TMEM_O = 128;
constexpr float RESCALE_LOG2 = 8.0f;              // lazy rescale threshold: P stays below 2^8
// of every 16 key pairs (32 keys), those whose bit is set here run on the FMA pipe (5 of 16).  Measured with the
// packed code (base, 16 segments): 0 pairs 551 TFLOP/s, 1: 558, 2: 571, 3: 575, 5: 580, 6: 578
#ifndef WB_ATTN_POLY_PAIR_MASK
#define WB_ATTN_POLY_PAIR_MASK ((1u << 1) | (1u << 4) | (1u << 7) | (1u << 10) | (1u << 13))
#endif
constexpr uint32_t POLY_PAIR_MASK = WB_ATTN_POLY_PAIR_MASK;

struct AttnArgs {
  int B, T, H, n_kt, n_steps;
  __half* out;
  float scale_log2;   // scale * log2(e)
  long long* dbg;     // optional: clock64() trace of CTA (0,0,0) (tools/attn_trace.py); nullptr in production
};
// clock64() trace points: compiled in only with -DWB_ATTN_TRACE_BUILD (tools/attn_trace.py); as run-time
// predicated code they cost ~15 issue slots per softmax step
#ifdef WB_ATTN_TRACE_BUILD
#define ATTN_TRACE(slot)                                                          \
  do {                                                                            \
    if (trace) a.dbg[(slot)] = clock64();                                         \
  } while (0)
#else
#define ATTN_TRACE(slot) \
  do {                   \
    (void)trace;         \
  } while (0)
#endif

template <bool B>
struct BoolTag {
  static constexpr bool value = B;
};


// Lazy-rescale slow path: O[row][:] *= alpha over the accumulator's 80 columns (64 head values, the
// denominator, zeros).  Rare after the first steps, and deliberately NOT inlined: inlined, its
// registers are live next to the score registers of the steady-state loop.
__device__ __noinline__ void rescale_o(uint32_t taddr_o, float alpha) {
#pragma unroll 1
  for (int cc = 0; cc < 5; ++cc) {
    uint32_t o[16];
    tmem_ld_32x32b_x16(taddr_o + cc * 16, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
    tmem_st_32x32b_x16(taddr_o + cc * 16, o);
  }
  tmem_st_wait();
}

// P for 32 keys: 2^(c s - m c) as F16 pairs, element i (and its pair i + 1) in register i / 2.
// The scale-and-offset runs as packed FFMA2 (two keys per instruction); of every 16 key PAIRS, those
// whose bit is set in POLY_PAIR_MASK evaluate both exponentials on the FMA pipe (packed cubic), the
// rest go through the MUFU unit.
__device__ __forceinline__ void exp_chunk(const uint32_t (&s)[32], uint64_t c2, uint64_t m2, uint32_t (&p)[16],
                                          const Ex2Consts& K) {
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(s[2 * u]), __uint_as_float(s[2 * u + 1])), c2, m2);
    float x0, x1, e0, e1;
    f2_unpack(x2, x0, x1);
    if ((POLY_PAIR_MASK >> u) & 1u) {
      ex2_fma_x2(x0, x1, e0, e1, K);
    } else {
      e0 = ex2_mufu(x0);
      e1 = ex2_mufu(x1);
    }
    p[u] = pack_h2(e0, e1);
  }
}
// max over 32 scores (3-input FMNMX, 4 independent chains)
__device__ __forceinline__ float max_chunk(const uint32_t (&s)[32]) {
  float m[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) m[u] = fmax3(__uint_as_float(s[u]), __uint_as_float(s[4 + u]), __uint_as_float(s[8 + u]));
#pragma unroll
  for (int u = 0; u < 4; ++u) m[u] = fmax3(m[u], __uint_as_float(s[12 + u]), __uint_as_float(s[16 + u]));
#pragma unroll
  for (int u = 0; u < 4; ++u) m[u] = fmax3(m[u], __uint_as_float(s[20 + u]), __uint_as_float(s[24 + u]));
#pragma unroll
  for (int u = 0; u < 4; ++u) m[u] = fmaxf(m[u], __uint_as_float(s[28 + u]));
  return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
}

__global__ void __launch_bounds__(ATTN_THREADS, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap qk_map, const __grid_constant__ CUtensorMap vt_map,
                         const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);   // [NWG] Q tile landed
  uint64_t* bar_kfull = bar_q + NWG;        // [NS] K stage landed
  uint64_t* bar_vfull = bar_kfull + NS;     // [NS] V stage landed
  uint64_t* bar_free = bar_vfull + NS;      // [NS] every MMA reading the stage has retired
  uint64_t* bar_s = bar_free + NS;          // [NWG][2] S buffer written
  uint64_t* bar_p = bar_s + NWG * 2;        // [NWG][2] P written over the S buffer (256 arrivals)
  uint64_t* bar_o = bar_p + NWG * 2;        // [NWG] P.V of the latest step retired: O is idle
  uint64_t* bar_done = bar_o + NWG;         // [NWG] last P.V retired (single phase)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + NWG);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int h = blockIdx.y, b = blockIdx.z;
  const int H = a.H, T = a.T, n_kt = a.n_kt, n_steps = a.n_steps;
  const int q0 = blockIdx.x * (NWG * QT);   // first query of this CTA

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();   // SWIZZLE_128B tiles need 1024-byte alignment
    for (int i = 0; i < NWG; ++i) {
      mbar_init(&bar_q[i], 1);
      mbar_init(&bar_o[i], 1);
      mbar_init(&bar_done[i], 1);
      for (int k = 0; k < 2; ++k) {
        mbar_init(&bar_s[i * 2 + k], 1);
        mbar_init(&bar_p[i * 2 + k], QT / 32);   // one arrival per softmax warp
      }
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&bar_kfull[i], 1);
      mbar_init(&bar_vfull[i], 1);
      mbar_init(&bar_free[i], NWG);   // one commit per MMA warp
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();

  if (warp == TMA_WARP) {
    // ===================== TMA loads, one thread =====================
    if ((tid & 31) == 0) {
      prefetch_tmap(&qk_map);
      prefetch_tmap(&vt_map);
      const int vrow = (b * H + h) * VROWS;
      for (int wg = 0; wg < NWG; ++wg) {
        mbar_arrive_expect_tx(&bar_q[wg], TILE_QK_BYTES);
        tma_load_4d(smem + SMEM_Q + wg * TILE_QK_BYTES, &qk_map, &bar_q[wg], 0, h, q0 + wg * QT, b);
      }
      for (int j = 0; j < n_kt; ++j) {
        const int st = j % NS;
        if (j >= NS) mbar_wait(&bar_free[st], ((j / NS) - 1) & 1);
        mbar_arrive_expect_tx(&bar_kfull[st], TILE_QK_BYTES);
        tma_load_4d(smem + SMEM_K + st * TILE_QK_BYTES, &qk_map, &bar_kfull[st], 0, H + h, j * KT, b);
        mbar_arrive_expect_tx(&bar_vfull[st], TILE_V_BYTES);
        tma_load_2d(smem + SMEM_V + st * TILE_V_BYTES, &vt_map, &bar_vfull[st], j * KT, vrow);
        tma_load_2d(smem + SMEM_V + st * TILE_V_BYTES + TILE_V_HALF_BYTES, &vt_map, &bar_vfull[st], j * KT + 64, vrow);
      }
    }
  } else if (warp >= MMA_WARP && warp < TMA_WARP) {
    // ===================== the MMAs of one warpgroup: the warp runs this converged, one elected lane issues =====================
    const int wg = warp - MMA_WARP;
    const bool trace = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid & 31) == 0;
    constexpr uint32_t idesc_s = umma_idesc_f16(QT, KS);      // 128 x 64
    constexpr uint32_t idesc_o = umma_idesc_f16(QT, VROWS);   // 128 x 80
    const uint32_t q_addr = smem_u32(smem + SMEM_Q), k_addr = smem_u32(smem + SMEM_K), v_addr = smem_u32(smem + SMEM_V);
    auto issue_s = [&](int wg, int j) {   // S_j[wg] = Q[wg] K_j^T  -> S buffer j & 1
      const uint64_t dq = umma_desc_k_sw128(q_addr + wg * TILE_QK_BYTES);
      const uint64_t dk = umma_desc_k_sw128(k_addr + ((j >> 1) % NS) * TILE_QK_BYTES + (j & 1) * (KS * 128));
      const uint32_t d = tmem_base + wg * TMEM_WG_STRIDE + TMEM_S + (j & 1) * KS;
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_f16_ss_elect(d, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
    };
    auto issue_o = [&](int wg, int j) {   // O[wg] += P_j[wg] [V_j | 1]
      const uint64_t dv = umma_desc_k_sw128(v_addr + ((j >> 1) % NS) * TILE_V_BYTES + (j & 1) * TILE_V_HALF_BYTES);
      const uint32_t pa = tmem_base + wg * TMEM_WG_STRIDE + TMEM_S + (j & 1) * KS;   // P_j: 64 F16 = 32 columns
      const uint32_t d = tmem_base + wg * TMEM_WG_STRIDE + TMEM_O;
#pragma unroll
      for (int k = 0; k < KS / 16; ++k) umma_f16_ts_elect(d, pa + 8 * k, dv + 2 * k, idesc_o, (j | k) != 0);
    };
    // prologue: the first two score tiles
    mbar_wait(&bar_kfull[0], 0);
    mbar_wait(&bar_q[wg], 0);
    tc_fence_after();
    issue_s(wg, 0);
    umma_commit_elect(&bar_s[wg * 2 + 0]);
    if (n_steps > 1) {
      issue_s(wg, 1);
      umma_commit_elect(&bar_s[wg * 2 + 1]);
    }
    for (int j = 0; j < n_steps; ++j) {
      const int J = j >> 1, sb = j & 1;
      ATTN_TRACE(512 + (j * 2 + wg) * 4 + 0);
      mbar_wait(&bar_p[wg * 2 + sb], (j >> 1) & 1);   // P_j in TMEM, S_j consumed
      if (sb == 0) mbar_wait(&bar_vfull[J % NS], (J / NS) & 1);
      tc_fence_after();
      ATTN_TRACE(512 + (j * 2 + wg) * 4 + 1);
      issue_o(wg, j);
      umma_commit_elect(&bar_o[wg]);
      if (j == n_steps - 1) umma_commit_elect(&bar_done[wg]);
      ATTN_TRACE(512 + (j * 2 + wg) * 4 + 2);
      if (j + 2 < n_steps) {
        const int J2 = (j + 2) >> 1;
        if (sb == 0) {
          mbar_wait(&bar_kfull[J2 % NS], (J2 / NS) & 1);
          tc_fence_after();
        }
        issue_s(wg, j + 2);   // overwrites S_j / P_j: ordered behind the P.V MMAs above by the tensor pipe
        umma_commit_elect(&bar_s[wg * 2 + sb]);
      }
      ATTN_TRACE(512 + (j * 2 + wg) * 4 + 3);
      // this warpgroup's MMAs on both halves of stage J are issued: free once they (and the other
      // warpgroup's) retire
      if (sb == 1 || j == n_steps - 1) umma_commit_elect(&bar_free[J % NS]);
    }
  } else {
    // ===================== softmax warpgroups: thread r owns query row r (TMEM lane r) =====================
    const int g = warp >> 2;
    const int r = tid & (QT - 1);
    const uint32_t twg = tmem_base + g * TMEM_WG_STRIDE + (uint32_t((warp & 3) * 32) << 16);
    const float c = a.scale_log2;
    float m_used = -INFINITY;   // row max (raw score units) the exponent offset currently refers to
    const bool trace = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && r == 0;
    Ex2Consts K;
    K.load();
    uint64_t c2;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(c));

    // one 64-key step.  MASKED = the last step of the sequence (keys >= T get -inf); kept out of the
    // steady-state instantiation: as a run-time test the compiler turns it into 3 predicated
    // instructions per element on every step.
    // (Tried and measured slower or equal on B200, see DESIGN.md: 16 softmax warps with half a row per
    // thread; software-pipelining the next step's TMEM load and max under the exponentials; requesting the
    // next step's scores into each half of the score registers as soon as its exponentials are done -- no
    // extra registers, but 579 -> 525 TFLOP/s: the barrier wait + fence + tcgen05.ld in mid-step stall the
    // exponential stream more than the early data saves.  Round 2: a SPECULATIVE exponent -- this step's exponentials
    // with the previous steps' offset, the row max computed beside them and used only to decide whether the step must
    // be redone -- takes the max off the dependency chain but holds both P halves in registers (115 -> 164): 485 -> 468
    // TFLOP/s at whisper medium, 64 segments; with 7 of 16 pairs on the FMA pipe as well 445; 7 of 16 without
    // speculation 494, i.e. flat.  Also round 2: P in its OWN TMEM columns (S0 | S1 | P | O = 240 columns, K ring of
    // three stages) so that S_{j+2} is issued as soon as S_j has been read instead of after P_j.V_j -- ncu attributes
    // 17 % of the softmax warps' samples to the "S ready" barrier -- costs an extra barrier round trip per step in both
    // the softmax and the MMA warp and 11 registers: 476 -> 436 TFLOP/s.  Late round 2, all A/B'd against this kernel in
    // the same gpurun call at whisper medium, 64 segments (profiles/r02_attn_*): THREE CTAs per SM -- one S / P buffer,
    // the denominator summed in the softmax threads so that S | O fit 128 TMEM columns, 64-key stages in a ring of three,
    // 112 registers -- makes softmax_j -> P_j V_j -> Q K_{j+1}^T -> softmax_{j+1} one dependent chain per CTA: 480 -> 360.
    // Probing the next step's "S ready" barrier early with mbarrier.test_wait: 485 -> 468 (on an idle SM a completed
    // try_wait is only 50 cycles, tools/ubench/mbar_lat.cu: the ~250 cycles a step spends before its TMEM load are the
    // dependent address / predicate / branch instructions queued behind the other warps, not the barrier).  TWO softmax
    // warpgroups per tile taking alternate steps (4 softmax warps per scheduler, the row max handed over through shared
    // memory and named barriers): 471 -> 432 as long as P overwrote S (the warpgroup then idles ~1000 cycles for
    // P_j V_j -> Q K_{j+2}^T); with P in its own TMEM columns, 64-key K / V^T rings and an "S read" barrier 490 -> 476,
    // the single MMA warp now the pace-setter (4 waits x ~180 cycles + 8 MMAs + commits = the whole step); with TWO
    // MMA-issuing warps that also issue their operand's TMA loads (no TMA warp, no "stage free" barriers) the step drops
    // from ~1330 to ~1200 cycles in the clock trace, but 479 -> 481 TFLOP/s and 698 -> 697 segments/s in the bench: the
    // chip sits at its 985 W power cap either way (SM clock 1.39 GHz of 1.965), so cycles saved come back as a lower
    // clock.  The variant is kept as profiles/r02_attn_variant_two_softmax_warpgroups.patch.)
    auto step = [&](const int j, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const int sb = j & 1;
      ATTN_TRACE((j * 2 + g) * 8 + 0);
      mbar_wait(&bar_s[g * 2 + sb], (j >> 1) & 1);   // S_j ready (issued two steps ago: normally no wait)
      __syncwarp();
      tc_fence_after();
      ATTN_TRACE((j * 2 + g) * 8 + 1);
      const uint32_t my_s = twg + TMEM_S + sb * KS;
      uint32_t s0[32], s1[32];
      tmem_ld_32x32b_x32(my_s, s0);
      tmem_ld_32x32b_x32(my_s + 32, s1);
      tmem_ld_wait();
      if (MASKED) {
        const int kbase = j * KS;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (kbase + i >= T) s0[i] = 0xff800000u;   // -inf: masked key
          if (kbase + 32 + i >= T) s1[i] = 0xff800000u;
        }
      }
      const float mx = fmaxf(max_chunk(s0), max_chunk(s1));
      ATTN_TRACE((j * 2 + g) * 8 + 2);
      // ---- lazy running max: rescale O only when this row's max grew by more than 2^8
      const bool grow = (mx - m_used) * c > RESCALE_LOG2;   // true on the first step (m_used = -inf)
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        mbar_wait(&bar_o[g], (j - 1) & 1);   // P_{j-1} V_{j-1} retired: O is idle until P_j is handed over
        __syncwarp();
        tc_fence_after();
        rescale_o(twg + TMEM_O, grow ? exp2f((m_used - mx) * c) : 1.0f);
      }
      if (grow) m_used = mx;
      const float moff = m_used * c;
      ATTN_TRACE((j * 2 + g) * 8 + 3);
      // ---- P = 2^(c s - m c) as F16 pairs (2 keys per column) over the first 32 columns of the S
      // buffer just read
      {
        uint32_t p[16];
        const uint64_t m2 = f2_pack(-moff, -moff);
        exp_chunk(s0, c2, m2, p, K);
        tmem_st_32x32b_x16(my_s, p);
        exp_chunk(s1, c2, m2, p, K);
        tmem_st_32x32b_x16(my_s + 16, p);
      }
      tmem_st_wait();
      tc_fence_before();   // TMEM accesses ordered before the MMAs the MMA warp issues
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&bar_p[g * 2 + sb]);   // (128 arrivals on one barrier word serialise in the shared-memory atomic unit)
      ATTN_TRACE((j * 2 + g) * 8 + 4);
    };
#pragma unroll 1
    for (int j = 0; j < n_steps - 1; ++j) step(j, BoolTag<false>{});
    if (n_steps * KS > T) step(n_steps - 1, BoolTag<true>{});
    else step(n_steps - 1, BoolTag<false>{});
    // (a parity wait on bar_o could be two phases behind here and return early: own barrier)
    mbar_wait(&bar_done[g], 0);
    __syncwarp();
    tc_fence_after();
    // ---- normalise and store merged heads: out[(b*T + t)][h*64 + c]  (permute + cpy, 1924-1929)
    uint32_t o0[32], o1[32], o2[16];
    tmem_ld_32x32b_x32(twg + TMEM_O, o0);
    tmem_ld_32x32b_x32(twg + TMEM_O + 32, o1);
    tmem_ld_32x32b_x16(twg + TMEM_O + 64, o2);
    tmem_ld_wait();
    const int t = q0 + g * QT + r;
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
  if (warp == MMA_WARP) {
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
  a.n_steps = (p.T + KS - 1) / KS;
  a.out = p.out;
  a.scale_log2 = p.scale * 1.4426950408889634f;


  a.dbg = p.dbg;
  dim3 grid((p.T + NWG * QT - 1) / (NWG * QT), p.H, p.B);
  return launch_pdl(attention_tcgen05_kernel, grid, dim3(ATTN_THREADS), ATTN_SMEM_BYTES, st, p.qk_map, p.vt_map, a);
}

}  // namespace wb
