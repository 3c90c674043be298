// decode_kernels.cu -- decoder-step kernels (sm_100a, CUDA cores; every one is HBM-bound).
//
// The reference declares the decoder's state (WhisperLayerDecoder src/main.rs:694-731, memory_k/v
// 1343-1347, memory_cross_k/v 1350-1354, logits 351-352) but implements no decode step; these
// kernels implement SURVEY.md section 8a rows D1-D6 (upstream whisper.cpp v1.0.3 semantics) on
// the state layout the reference's encoder leaves behind: F16 K/V, cross-K pre-scaled by
// (d/H)^-1/4 (1994-1996), rows of d with each head's 64 values contiguous (128 bytes).
//
// Attention over a cache row-block is laid out so that 8 lanes read one 128-byte head row as
// 8 x 16-byte chunks: every load instruction of a warp covers 4 full rows, fully coalesced.
#include "ptx.cuh"
#include "wb_kernels.hpp"

namespace wb {

namespace {

constexpr int DH = 64;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// streaming 16-byte load (read-only path, do not keep the line in L1)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// (sum, sum of squares) -> the fixed-point form the folded LayerNorm statistics are accumulated in
__device__ __forceinline__ DecLnStat dec_ln_fixed(float s1, float s2) {
  DecLnStat r;
  r.s1 = (unsigned long long)__float2ll_rn(s1 * DEC_LN_S1_SCALE);
  r.s2 = (unsigned long long)__float2ll_rn(s2 * DEC_LN_S2_SCALE);
  return r;
}

__device__ __forceinline__ float block_max(float v, float* sh, int n_warps) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < n_warps; ++i) r = fmaxf(r, sh[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* sh, int n_warps) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.0f;
  for (int i = 0; i < n_warps; ++i) r += sh[i];
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// D1: x[s][i][:] = d_te[tok[s][i]][:] + d_pe[n_past + i][:]
// With `stats` (the LayerNorm-folded single-token step): also leaves the row's exact mean in center[row], the (sum,
// sum of squares) of x - center in stats[row] -- the statistics of layer 0's attn_ln -- and an F16 copy of
// x - center (as in the encoder, gemm2.cu: rounding x itself to F16 would lose the deviations of a row whose mean is
// large against its spread; the producers downstream keep using, and the consumers keep advancing, that centre), and block 0 clears this
// launch's rows of the other `n_clear_slots` statistics slots of the step (their producers accumulate with
// atomics; another sequence group's rows of the same slots may be in use on another stream).
__global__ void __launch_bounds__(128)
embed_kernel(const __half* __restrict__ te, const float* __restrict__ pe, const int* __restrict__ tokens,
             int n_tok, const int* __restrict__ n_past_p, int d, float* __restrict__ x, DecLnStat* __restrict__ stats,
             __half* __restrict__ x16, int n_clear_slots, float* __restrict__ center) {
  pdl_launch_dependents();   // the next kernel of the step may become resident now; it blocks at its own wait
  pdl_wait();                // everything this kernel reads is the previous kernels' output
  __shared__ float red[2][4];
  const int row = blockIdx.x;   // s * n_tok + i
  const int i = row % n_tok;
  const int tok = tokens[row];
  const int pos = *n_past_p + i;
  const __half* e = te + (size_t)tok * d;
  const float* p = pe + (size_t)pos * d;
  float s1 = 0.0f, s2 = 0.0f;
  constexpr int EMB_MAX = 10;   // d <= 1280 on 128 threads
  float vals[EMB_MAX];
#pragma unroll
  for (int u = 0; u < EMB_MAX; ++u) {
    const int c = threadIdx.x + u * 128;
    vals[u] = 0.0f;
    if (c < d) {
      vals[u] = __half2float(e[c]) + p[c];
      x[(size_t)row * d + c] = vals[u];
      s1 += vals[u];
    }
  }
  if (stats) {
    __shared__ float red_m[4];
    s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) red_m[threadIdx.x >> 5] = s1;
    __syncthreads();
    const float mean = ((red_m[0] + red_m[1]) + (red_m[2] + red_m[3])) / (float)d;
    if (threadIdx.x == 0) center[row] = mean;
    s1 = 0.0f;
#pragma unroll
    for (int u = 0; u < EMB_MAX; ++u) {
      const int c = threadIdx.x + u * 128;
      if (c < d) {
        const float w = vals[u] - mean;
        x16[(size_t)row * d + c] = __float2half_rn(w);
        s1 += w;
        s2 = fmaf(w, w, s2);
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = s1;
      red[1][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0)   // this slot's sub-slot 0 holds the whole row; its other sub-slots are cleared below
      stats[row] = dec_ln_fixed((red[0][0] + red[0][1]) + (red[0][2] + red[0][3]), (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]));
    if (blockIdx.x == 0) {
      const int rows = gridDim.x;
      // sub-slots 1 .. of slot 0 and every sub-slot of the n_clear_slots slots behind it, this launch's rows only
      for (int c = threadIdx.x; c < ((1 + n_clear_slots) * DEC_LN_SUB - 1) * rows; c += blockDim.x)
        stats[(size_t)(1 + c / rows) * DEC_LN_ROWS + c % rows] = DecLnStat{0ull, 0ull};
    }
  }
}

// ---------------------------------------------------------------------------------------------
// D2: causal self-attention over the F16 KV cache, one CTA per (head, sequence).
//   qkv   [n_seq*n_tok][3d] F16: Q and K already scaled by Dh^-1/4, V plain (GEMM epilogue)
//   cache [seq][n_text_ctx][d] F16 for this layer; the CTA first appends its head's new K/V rows
//         at positions n_past .. n_past+n_tok-1 (cache append of D2), then attends.
// Probabilities are normalised and rounded to F16 before P.V, as ggml's mul_mat does to its
// second operand when the first (V) is F16.
constexpr int SELF_THREADS = 128;
constexpr int SELF_MAX_CTX = 448 + 64;
constexpr int SELF_U = 4;          // key rows AND value rows per lane requested together
constexpr int SELF_GROUPS = (SELF_THREADS / 32) * 4;   // 8-lane groups per CTA: one cache row each per instruction

// One pass over the cached keys (online softmax, K and V rows requested together, as decode_cross_attn_kernel):
// round 1 ran scores -> shared memory -> block softmax -> P.V, i.e. two dependent L2 round trips and six block
// barriers per token for a few hundred keys; here every 8-lane group keeps a running (max, denominator, 8 output
// dims per lane) over its rows and the 16 groups are merged once.  The probabilities are rounded to F16 before P.V
// relative to the running maximum (the reference rounds the normalised ones: the same F16 grid up to a power of two
// when the maximum is final, within the logit tolerance otherwise).
__global__ void __launch_bounds__(SELF_THREADS)
decode_self_attn_kernel(const __half* __restrict__ qkv, int d, __half* __restrict__ kc, __half* __restrict__ vc,
                        int n_tok, const int* __restrict__ n_past_p, int n_text_ctx, __half* __restrict__ out) {
  pdl_launch_dependents();   // the next kernel of the step may become resident now; it blocks at its own wait
  pdl_wait();                // everything this kernel reads is the previous kernels' output
  __shared__ float g_m[SELF_GROUPS], g_l[SELF_GROUPS];
  __shared__ float g_o[SELF_GROUPS][DH];
  const int h = blockIdx.x, s = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_past = *n_past_p;
  __half* kbase = kc + (size_t)s * n_text_ctx * d + h * DH;
  __half* vbase = vc + (size_t)s * n_text_ctx * d + h * DH;
  // append this head's new K / V rows (16-byte chunks)
  for (int idx = tid; idx < n_tok * 8; idx += SELF_THREADS) {
    const int i = idx >> 3, ch = idx & 7;
    const __half* src = qkv + (size_t)(s * n_tok + i) * 3 * d + h * DH + ch * 8;
    *reinterpret_cast<uint4*>(kbase + (size_t)(n_past + i) * d + ch * 8) = *reinterpret_cast<const uint4*>(src + d);
    *reinterpret_cast<uint4*>(vbase + (size_t)(n_past + i) * d + ch * 8) = *reinterpret_cast<const uint4*>(src + 2 * d);
  }
  __syncthreads();
  const int sub = lane >> 3, ch = lane & 7;   // 4 rows per warp instruction, 8 lanes per 128-byte row
  constexpr int RPI = SELF_GROUPS;            // rows per CTA iteration
  for (int i = 0; i < n_tok; ++i) {
    const int Tk = n_past + i + 1;            // causal: keys 0 .. n_past + i
    float q[8];
    unpack8(*reinterpret_cast<const uint4*>(qkv + (size_t)(s * n_tok + i) * 3 * d + h * DH + ch * 8), q);
    float m = -INFINITY, l = 0.0f, o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.0f;
    for (int t0 = warp * 4; t0 < Tk; t0 += SELF_U * RPI) {
      uint4 kv[SELF_U], vv[SELF_U];
#pragma unroll
      for (int u = 0; u < SELF_U; ++u) {
        const int t = t0 + u * RPI + sub;
        kv[u] = t < Tk ? *reinterpret_cast<const uint4*>(kbase + (size_t)t * d + ch * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < SELF_U; ++u) {
        const int t = t0 + u * RPI + sub;
        vv[u] = t < Tk ? *reinterpret_cast<const uint4*>(vbase + (size_t)t * d + ch * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
      float sc[SELF_U];
      float bm = -INFINITY;
#pragma unroll
      for (int u = 0; u < SELF_U; ++u) {
        const int t = t0 + u * RPI + sub;
        float kf[8];
        unpack8(kv[u], kf);
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(q[j], kf[j], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        sc[u] = t < Tk ? acc : -INFINITY;
        bm = fmaxf(bm, sc[u]);
      }
      const float m_new = fmaxf(m, bm);
      if (m_new > -INFINITY) {   // (group-uniform: the 8 lanes hold the same scores)
        const float alpha = __expf(m - m_new);   // 0 on the group's first rows (m = -inf)
        l *= alpha;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] *= alpha;
#pragma unroll
        for (int u = 0; u < SELF_U; ++u) {
          const float e = __expf(sc[u] - m_new);   // 0 for a row past the end
          l += e;
          const float p = __half2float(__float2half_rn(e));   // P -> F16 before P.V
          float vf[8];
          unpack8(vv[u], vf);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(p, vf[j], o[j]);
        }
        m = m_new;
      }
    }
    const int grp = warp * 4 + sub;
    if (ch == 0) {
      g_m[grp] = m;
      g_l[grp] = l;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g_o[grp][ch * 8 + j] = o[j];
    __syncthreads();
    if (tid < DH) {
      float mx = -INFINITY;
#pragma unroll
      for (int gq = 0; gq < SELF_GROUPS; ++gq) mx = fmaxf(mx, g_m[gq]);
      float sum = 0.0f, r = 0.0f;
#pragma unroll
      for (int gq = 0; gq < SELF_GROUPS; ++gq) {
        const float w = g_m[gq] > -INFINITY ? __expf(g_m[gq] - mx) : 0.0f;
        sum = fmaf(w, g_l[gq], sum);
        r = fmaf(w, g_o[gq][tid], r);
      }
      out[(size_t)(s * n_tok + i) * d + h * DH + tid] = __float2half_rn(sum > 0.0f ? r / sum : 0.0f);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// D3: cross-attention over the encoder memory (memory_cross_k/v), the decoder's dominant HBM
// stream: 2 * T * 128 bytes per (sequence, head, layer) per generated token.
//   q   [n_seq*n_tok][d] F16 (already scaled by Dh^-1/4)
//   K/V rows: kv[(seq*T + t)*ld + h*64 ..], K at `k`, V at `v` (same row pitch)
// grid (H, n_seq, n_split): each CTA handles keys [split*span, ...) for every query token; with
// n_split > 1 it emits (max, sum, o normalised within the split) partials; the CTA of a (row, head) that
// finishes last merges them.
constexpr int CROSS_THREADS = 256;
constexpr int CROSS_MAX_SPAN = 1536;
// CROSS_U (template parameter) = K rows AND V rows per lane requested together (2 U x 16 bytes in flight per lane);
// MINB = CTAs per SM the register budget is held to.  <4, 3> keeps 8 loads per lane in flight on 444 CTA slots;
// <3, 4> trades two of them for 592 slots.  The kernel is bound by the bytes a CTA keeps in flight, so the launch
// picks the variant that holds every (head, sequence) item in ONE wave: whisper medium with 32 sequences has 512
// items -- 2.03 ms per step on 444 slots (a second wave of 68 CTAs), 1.81 ms on 592 (round 2, tools/dec_groups.py).
constexpr int CROSS_GROUPS = (CROSS_THREADS / 32) * 4;   // 8-lane groups per CTA: one key row each per instruction

// One pass over the keys (online softmax): every 8-lane group walks its rows with K and V requested together
// and keeps a running (max, denominator, 8 output dims per lane); the 32 groups are merged once at the end.
// Against two passes (scores -> shared memory, block softmax, then V) there is no mid-kernel barrier during
// which all resident CTAs stop streaming at the same time, and twice the loads are in flight in steady state.
// The probabilities are rounded to F16 before P.V as in the reference, relative to the running maximum.
template <int CROSS_U, int MINB>
__global__ void __launch_bounds__(CROSS_THREADS, MINB)
decode_cross_attn_kernel(const __half* __restrict__ q, int d, const __half* __restrict__ k, const __half* __restrict__ v,
                         long long ld, long long head_stride, int n_tok, int T, int span, __half* __restrict__ out,
                         float* __restrict__ part_o, float* __restrict__ part_ml, int n_split, int* __restrict__ split_cnt) {
  __shared__ float g_m[CROSS_GROUPS], g_l[CROSS_GROUPS];
  __shared__ int s_last;
  __shared__ float g_o[CROSS_GROUPS][DH];
  const int h = blockIdx.x, s = blockIdx.y, sp = blockIdx.z, H = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane >> 3, ch = lane & 7;
  const int t_lo = sp * span, t_hi = min(T, t_lo + span), n = t_hi - t_lo;
  const __half* kb = k + ((size_t)s * T + t_lo) * ld + h * head_stride + ch * 8;
  const __half* vb = v + ((size_t)s * T + t_lo) * ld + h * head_stride + ch * 8;
  pdl_launch_dependents();
  pdl_wait();   // q is the previous kernel's output
  constexpr int RPI = CROSS_GROUPS;   // rows per CTA iteration
  for (int i = 0; i < n_tok; ++i) {
    const int row = s * n_tok + i;
    float qf[8];
    unpack8(*reinterpret_cast<const uint4*>(q + (size_t)row * d + h * DH + ch * 8), qf);
    float m = -INFINITY, l = 0.0f, o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.0f;
    for (int t0 = warp * 4; t0 < n; t0 += CROSS_U * RPI) {
      uint4 kv[CROSS_U], vv[CROSS_U];
#pragma unroll
      for (int u = 0; u < CROSS_U; ++u) {
        const int t = t0 + u * RPI + sub;
        kv[u] = t < n ? ld_nc_v4(kb + (size_t)t * ld) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < CROSS_U; ++u) {
        const int t = t0 + u * RPI + sub;
        vv[u] = t < n ? ld_nc_v4(vb + (size_t)t * ld) : make_uint4(0u, 0u, 0u, 0u);
      }
      float sc[CROSS_U];
      float bm = -INFINITY;
#pragma unroll
      for (int u = 0; u < CROSS_U; ++u) {
        const int t = t0 + u * RPI + sub;
        float kf[8];
        unpack8(kv[u], kf);
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(qf[j], kf[j], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        sc[u] = t < n ? acc : -INFINITY;
        bm = fmaxf(bm, sc[u]);
      }
      const float m_new = fmaxf(m, bm);
      if (m_new > -INFINITY) {   // (group-uniform: the 8 lanes hold the same scores)
        const float alpha = __expf(m - m_new);   // 0 on the group's first rows (m = -inf)
        l *= alpha;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] *= alpha;
#pragma unroll
        for (int u = 0; u < CROSS_U; ++u) {
          const float e = __expf(sc[u] - m_new);   // 0 for a row past the end
          l += e;
          const float p = __half2float(__float2half_rn(e));   // P -> F16 before P.V
          float vf[8];
          unpack8(vv[u], vf);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(p, vf[j], o[j]);
        }
        m = m_new;
      }
    }
    const int grp = warp * 4 + sub;
    if (ch == 0) {
      g_m[grp] = m;
      g_l[grp] = l;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g_o[grp][ch * 8 + j] = o[j];
    __syncthreads();
    if (tid < DH) {
      float mx = -INFINITY;
#pragma unroll
      for (int gq = 0; gq < CROSS_GROUPS; ++gq) mx = fmaxf(mx, g_m[gq]);
      float sum = 0.0f, r = 0.0f;
#pragma unroll
      for (int gq = 0; gq < CROSS_GROUPS; ++gq) {
        const float w = g_m[gq] > -INFINITY ? __expf(g_m[gq] - mx) : 0.0f;
        sum = fmaf(w, g_l[gq], sum);
        r = fmaf(w, g_o[gq][tid], r);
      }
      r = sum > 0.0f ? r / sum : 0.0f;
      if (n_split == 1) {
        out[(size_t)row * d + h * DH + tid] = __float2half_rn(r);
      } else {
        const size_t pi = ((size_t)row * H + h) * n_split + sp;
        part_o[pi * DH + tid] = r;            // normalised within the split
        if (tid == 0) {
          part_ml[pi * 2] = mx;
          part_ml[pi * 2 + 1] = sum;
        }
      }
    }
    if (n_split > 1) {
      // the key range of a (row, head) is split over n_split CTAs so that 148 SMs x 3 resident CTAs stay evenly
      // loaded (384 whole items leave SMs with 2 or 3 of them: 86 %); the CTA that finishes last merges the
      // partials -- no second launch
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const int old = atomicAdd(&split_cnt[row * H + h], 1);
        s_last = old == n_split - 1;
        if (s_last) split_cnt[row * H + h] = 0;   // ready for the next launch
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        if (tid < DH) {
          const size_t p0 = ((size_t)row * H + h) * n_split;
          float mm = -INFINITY;
          for (int q2 = 0; q2 < n_split; ++q2) mm = fmaxf(mm, __ldcg(part_ml + (p0 + q2) * 2));
          float den = 0.0f, acc = 0.0f;
          for (int q2 = 0; q2 < n_split; ++q2) {
            const float w = __ldcg(part_ml + (p0 + q2) * 2 + 1) * __expf(__ldcg(part_ml + (p0 + q2) * 2) - mm);
            den += w;
            acc = fmaf(w, __ldcg(part_o + (p0 + q2) * DH + tid), acc);
          }
          out[(size_t)row * d + h * DH + tid] = __float2half_rn(acc / den);
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// D6 / K13: greedy selection.  Per sequence: arg-max over the logits (first index wins ties),
// top-1 minus top-2 margin, then the loop bookkeeping of the device-side greedy decode:
// record the token unless the sequence already emitted `eot`, feed it to the next step.
struct Top2 {
  float v1, v2;
  int i1;
};

// End-of-step bookkeeping folded into the arg-max kernel (one launch less per token): every block has read the step
// counter before it arrives here, and the block that arrives last moves n_past / step on for the next step.
// n_past_p[0] = n_past, [1] = step, [2] = arrival counter (self-resetting).  n_past_p == nullptr: leave them alone.
__device__ __forceinline__ void step_advance(int* step_p, int* n_past_p, int advance_by) {
  if (!n_past_p) return;
  __threadfence();
  const int old = atomicAdd(&n_past_p[2], 1);
  if (old == (int)gridDim.x - 1) {
    n_past_p[0] += advance_by;
    *step_p += 1;
    n_past_p[2] = 0;
  }
}
__device__ __forceinline__ Top2 top2_merge(Top2 a, Top2 b) {
  Top2 r;
  if (b.v1 > a.v1 || (b.v1 == a.v1 && b.i1 < a.i1)) {
    r.v1 = b.v1;
    r.i1 = b.i1;
    r.v2 = fmaxf(a.v1, b.v2);
  } else {
    r.v1 = a.v1;
    r.i1 = a.i1;
    r.v2 = fmaxf(a.v2, b.v1);
  }
  return r;
}

__global__ void __launch_bounds__(1024)
argmax_kernel(const float* __restrict__ logits, int n_vocab, int* __restrict__ next_tok, float* __restrict__ margin_out,
              int* __restrict__ out_tokens, float* __restrict__ out_margin, int* __restrict__ out_len,
              int* __restrict__ done, int max_new, int* __restrict__ step_p, int eot, int* __restrict__ n_past_p,
              int advance_by) {
  pdl_launch_dependents();   // the next kernel of the step may become resident now; it blocks at its own wait
  pdl_wait();                // everything this kernel reads is the previous kernels' output
  __shared__ Top2 sh[32];
  const int s = blockIdx.x;
  const float* lg = logits + (size_t)s * n_vocab;
  Top2 t{-INFINITY, -INFINITY, 0x7fffffff};
  for (int i = threadIdx.x; i < n_vocab; i += blockDim.x) {
    const float v = lg[i];
    if (v > t.v1) {
      t.v2 = t.v1;
      t.v1 = v;
      t.i1 = i;
    } else if (v > t.v2) {
      t.v2 = v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Top2 u;
    u.v1 = __shfl_xor_sync(0xffffffffu, t.v1, o);
    u.v2 = __shfl_xor_sync(0xffffffffu, t.v2, o);
    u.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
    t = top2_merge(t, u);
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    Top2 r = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = top2_merge(r, sh[w]);
    next_tok[s] = r.i1;
    if (margin_out) margin_out[s] = r.v1 - r.v2;
    if (out_tokens) {
      const int step = *step_p;
      if (!done[s] && step < max_new) {
        out_tokens[(size_t)s * max_new + step] = r.i1;
        if (out_margin) out_margin[(size_t)s * max_new + step] = r.v1 - r.v2;
        out_len[s] = step + 1;
        if (r.i1 == eot) done[s] = 1;
      }
    }
    step_advance(step_p, n_past_p, advance_by);
  }
}

// ---------------------------------------------------------------------------------------------
// Skinny linear layer of a decode step: out[r][n] = epi( sum_k x[r][k] * W[n][k] ) for R <= 32
// activation rows (sequences) -- pure weight streaming, HBM-bound.  A 128-row tcgen05 tile would
// leave 6 .. 24 CTAs on a 148-SM chip here (N = d .. 4d weight rows), so this kernel gives every
// 16 weight rows their own CTA and uses warp-level mma.sync m16n8k16 (A = 16 weight rows, B = the
// activations' 8-row groups): 4 warps split K, each lane streams its weight rows as 16-byte loads
// straight from HBM (4 lanes cover 64 contiguous bytes of a row; the K order inside a 32-wide block
// is permuted identically for A and B so that one 16-byte load feeds two MMAs), partial sums meet
// in shared memory, and the epilogue (bias, column scale, GELU, f32 residual) writes out[r][n].
// For the vocabulary projection it also emits each CTA's per-sequence top-2 (value, index), so the
// arg-max never re-reads the logits (D5 + D6 fused).
struct Top2;
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  // (not volatile: a pure function of its operands, so loads of later k-blocks may be hoisted above it)
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// NW warps split K (4 for K <= 768, 8 beyond); U = 32-wide k-blocks per lane requested together -- the first U
// before the grid dependency resolves.  With K / NW = 32 U every weight of the CTA is in flight before the
// previous kernel has finished and the activations arrive in one batch: one exposed L2 round trip instead
// of one per loop iteration (the step is a chain of ~100 such latency-bound kernels).
constexpr int DL_ROWS = 16;       // weight rows (output features) per CTA

// PIPE (K longer than one round of 32 U NW, few CTAs: fc2): the next round's weights are requested before the current
// round's MMAs, in a second set of registers, so a CTA's rounds overlap instead of each paying a full HBM / L2 round
// trip (4 rounds at d = 1024 / 1280).  Costs ~2 U x 4 registers per thread: only for launches of at most one CTA per SM.
template <int U, int NW, bool PIPE = false>
__global__ void __launch_bounds__(32 * NW)
decode_linear_kernel(const DecodeLinear a) {
  constexpr int DL_THREADS = 32 * NW;
  __shared__ float red[NW][DL_ROWS][33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane & 3, g = lane >> 2;
  const int n0 = blockIdx.x * DL_ROWS;
  const int kw = a.K / NW;                      // this warp's K range (multiple of 16)
  const int k_lo = warp * kw, k_hi = k_lo + kw;
  // rows past N (ragged vocabulary) are clamped for the loads and masked at the store
  const int ra = min(n0 + g, a.N - 1), rb = min(n0 + g + 8, a.N - 1);
  const __half* wa = a.w + (size_t)ra * a.K + 8 * t;
  const __half* wb = a.w + (size_t)rb * a.K + 8 * t;
  const __half* xg = a.x + (size_t)g * a.ldx + 8 * t;   // activation row g of each 8-row group
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
  int k = k_lo;
  // 4 k-blocks of fragments (24 x 16-byte loads per lane) are requested before the first MMA uses
  // them: the kernel is pure HBM / L2 latency otherwise
  // ---- before the grid dependency resolves (programmatic dependent launch: this CTA is resident while the
  // previous kernel of the step still runs): the weights do not depend on it, so this CTA's 16 weight rows are
  // pulled into L2 and the first k-blocks into registers under the previous kernel.  The next kernel of the step
  // may become resident right away (it blocks at its own wait until this grid has completed).
  pdl_launch_dependents();
  {
    const int rows = min(DL_ROWS, a.N - n0);
    const char* wbase = reinterpret_cast<const char*>(a.w + (size_t)n0 * a.K);
    const size_t bytes = (size_t)rows * a.K * 2;
    for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)DL_THREADS * 128) prefetch_l2(wbase + off);
  }
  // the epilogue's per-feature constants are weights too: fetched here, not on the critical path after the MMAs
  const int er = tid >> 2, ef0 = (tid & 3) * 4;          // epilogue mapping: row er, features ef0 .. ef0 + 3
  const bool evec = n0 + ef0 + 3 < a.N;
  float4 e_bias = make_float4(0.f, 0.f, 0.f, 0.f), e_c1 = e_bias, e_cs = make_float4(1.f, 1.f, 1.f, 1.f);
  if (evec) {
    if (a.bias) e_bias = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + ef0));
    if (a.ln_in) e_c1 = __ldg(reinterpret_cast<const float4*>(a.ln_c1 + n0 + ef0));
    if (a.colscale) e_cs = __ldg(reinterpret_cast<const float4*>(a.colscale + n0 + ef0));
  }
  uint4 A0[U], B0[U];
  const bool pre = k + 32 * U <= k_hi;
  if (pre) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      A0[u] = ld_nc_v4(wa + k + 32 * u);
      B0[u] = ld_nc_v4(wb + k + 32 * u);
    }
  }
  pdl_wait();   // the activations (and the residual) are the previous kernel's output
  // residual row segment and row statistics: requested now, used after the MMAs
  float4 e_res = make_float4(0.f, 0.f, 0.f, 0.f);
  DecLnStat e_st{0ull, 0ull};
  float e_center = 0.0f;   // producer: the row's centre (x16 copy and statistics are of x - centre)
  if (a.ln_out && er < a.R) e_center = a.ln_center[er];
  if (a.residual && er < a.R && evec && (a.res_ld & 3) == 0)
    e_res = *reinterpret_cast<const float4*>(a.residual + (size_t)er * a.res_ld + n0 + ef0);
  if (a.ln_in) {   // the row's sub-slots, added as integers (any order gives the same sum)
    const DecLnStat* sp = a.ln_in + min(er, a.R - 1);
#pragma unroll
    for (int u = 0; u < DEC_LN_SUB; ++u) {
      const DecLnStat v = sp[u * DEC_LN_ROWS];
      e_st.s1 += v.s1;
      e_st.s2 += v.s2;
    }
  }
  if (pre) {
    uint4 X[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) X[u][j] = *reinterpret_cast<const uint4*>(xg + (size_t)(8 * j) * a.ldx + k + 32 * u);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mma_16816(acc[j], A0[u].x, B0[u].x, A0[u].y, B0[u].y, X[u][j].x, X[u][j].y);
        mma_16816(acc[j], A0[u].z, B0[u].z, A0[u].w, B0[u].w, X[u][j].z, X[u][j].w);
      }
    k += 32 * U;
  }
  if (PIPE && pre) {
    // rounds 1, 2, ...: ping-pong between two register sets, the request for round i + 1 issued before round i's MMAs
    uint4 A1[U], B1[U];
    auto request = [&](uint4 (&A)[U], uint4 (&B)[U], int kk) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        A[u] = ld_nc_v4(wa + kk + 32 * u);
        B[u] = ld_nc_v4(wb + kk + 32 * u);
      }
    };
    auto consume = [&](const uint4 (&A)[U], const uint4 (&B)[U], int kk) {
      uint4 X[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) X[u][j] = *reinterpret_cast<const uint4*>(xg + (size_t)(8 * j) * a.ldx + kk + 32 * u);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mma_16816(acc[j], A[u].x, B[u].x, A[u].y, B[u].y, X[u][j].x, X[u][j].y);
          mma_16816(acc[j], A[u].z, B[u].z, A[u].w, B[u].w, X[u][j].z, X[u][j].w);
        }
    };
    bool have1 = k + 32 * U <= k_hi;
    if (have1) request(A1, B1, k);
    while (have1) {
      const bool have0 = k + 64 * U <= k_hi;
      if (have0) request(A0, B0, k + 32 * U);
      consume(A1, B1, k);
      k += 32 * U;
      if (!have0) break;
      have1 = k + 64 * U <= k_hi;
      if (have1) request(A1, B1, k + 32 * U);
      consume(A0, B0, k);
      k += 32 * U;
    }
  }
  for (; k + 32 * U <= k_hi; k += 32 * U) {
    uint4 A[U], B[U], X[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      A[u] = ld_nc_v4(wa + k + 32 * u);
      B[u] = ld_nc_v4(wb + k + 32 * u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) X[u][j] = *reinterpret_cast<const uint4*>(xg + (size_t)(8 * j) * a.ldx + k + 32 * u);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mma_16816(acc[j], A[u].x, B[u].x, A[u].y, B[u].y, X[u][j].x, X[u][j].y);
        mma_16816(acc[j], A[u].z, B[u].z, A[u].w, B[u].w, X[u][j].z, X[u][j].w);
      }
  }
  for (; k + 32 <= k_hi; k += 32) {
    const uint4 A = ld_nc_v4(wa + k), B = ld_nc_v4(wb + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 X = *reinterpret_cast<const uint4*>(xg + (size_t)(8 * j) * a.ldx + k);
      mma_16816(acc[j], A.x, B.x, A.y, B.y, X.x, X.y);
      mma_16816(acc[j], A.z, B.z, A.w, B.w, X.z, X.w);
    }
  }
  if (k < k_hi) {   // 16-wide tail (kw = 16 mod 32): lanes read 8 bytes
    const uint2 A = *reinterpret_cast<const uint2*>(a.w + (size_t)ra * a.K + k + 4 * t);
    const uint2 B = *reinterpret_cast<const uint2*>(a.w + (size_t)rb * a.K + k + 4 * t);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint2 X = *reinterpret_cast<const uint2*>(a.x + (size_t)(8 * j + g) * a.ldx + k + 4 * t);
      mma_16816(acc[j], A.x, B.x, A.y, B.y, X.x, X.y);
    }
  }
  // C fragment: c0,c1 = (feature g, rows 2t, 2t+1 of group j), c2,c3 = (feature g+8, same rows)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[warp][g][8 * j + 2 * t] = acc[j][0];
    red[warp][g][8 * j + 2 * t + 1] = acc[j][1];
    red[warp][g + 8][8 * j + 2 * t] = acc[j][2];
    red[warp][g + 8][8 * j + 2 * t + 1] = acc[j][3];
  }
  __syncthreads();
  if (tid >= 128) return;   // (warp-uniform) the epilogue is 32 rows x 4 feature groups
  // ---- epilogue: thread -> (row r = tid / 4, features f = 4 (tid % 4) .. +3): a row's 16 features
  // are written by 4 adjacent lanes as one contiguous run
  const int r = tid >> 2, f0 = (tid & 3) * 4;
  float v[4];
  // LayerNorm folded into this layer (ln_in: statistics of the activation rows; the weights carry gamma):
  //   rstd * (x w^T - mu * c1) + c2, c2 in the bias slot
  float ln_rstd = 1.0f, ln_nmr = 0.0f;
  if (a.ln_in) {
    // exact integer sums -> f64 for the mean / variance (E[x^2] - mu^2 in f64: no cancellation to speak of)
    const double mu = (double)(long long)e_st.s1 * (1.0 / (double)DEC_LN_S1_SCALE) * (double)a.ln_inv_d;
    const double ex2 = (double)(long long)e_st.s2 * (1.0 / (double)DEC_LN_S2_SCALE) * (double)a.ln_inv_d;
    ln_rstd = rsqrtf((float)fmax(ex2 - mu * mu, 0.0) + a.ln_eps);
    ln_nmr = -(float)mu * ln_rstd;
    // mu is the mean of x - centre: the CTA of the first 16 features moves the centre on for the next producer
    if (blockIdx.x == 0 && (tid & 3) == 0 && r < a.R) a.ln_center[r] += (float)mu;
  }
  if (evec && (!a.residual || (a.res_ld & 3) == 0)) {   // whole 4-feature group inside N: the prefetched constants
    const float bb[4] = {e_bias.x, e_bias.y, e_bias.z, e_bias.w}, c1[4] = {e_c1.x, e_c1.y, e_c1.z, e_c1.w};
    const float cs[4] = {e_cs.x, e_cs.y, e_cs.z, e_cs.w}, rs[4] = {e_res.x, e_res.y, e_res.z, e_res.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = f0 + i;
      float x = 0.0f;
#pragma unroll
      for (int w = 0; w < NW; ++w) x += red[w][f][r];
      if (a.ln_in) x = fmaf(x, ln_rstd, ln_nmr * c1[i]);
      x = (x + bb[i]) * cs[i] * a.scale;
      if (a.gelu) x = gelu_f16in(x);
      v[i] = x + rs[i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = f0 + i;
      float x = 0.0f;
#pragma unroll
      for (int w = 0; w < NW; ++w) x += red[w][f][r];
      const int n = n0 + f;
      if (n < a.N) {
        if (a.ln_in) x = fmaf(x, ln_rstd, ln_nmr * a.ln_c1[n]);
        if (a.bias) x += a.bias[n];
        if (a.colscale) x *= a.colscale[n];
        x *= a.scale;
        if (a.gelu) x = gelu_f16in(x);
        if (a.residual && r < a.R) x += a.residual[(size_t)r * a.res_ld + n];
      } else {
        x = -INFINITY;
      }
      v[i] = x;
    }
  }
  if (r < a.R) {
    if (n0 + f0 + 3 < a.N && (a.out_ld & 3) == 0) {
      if (a.out_f16) {
        uint2 u;
        u.x = pack_h2(v[0], v[1]);
        u.y = pack_h2(v[2], v[3]);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.out) + (size_t)r * a.out_ld + n0 + f0) = u;
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + (size_t)r * a.out_ld + n0 + f0) =
            make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n0 + f0 + i < a.N) {
          if (a.out_f16) reinterpret_cast<__half*>(a.out)[(size_t)r * a.out_ld + n0 + f0 + i] = __float2half_rn(v[i]);
          else reinterpret_cast<float*>(a.out)[(size_t)r * a.out_ld + n0 + f0 + i] = v[i];
        }
    }
  }
  if (a.ln_out) {   // producer of the next folded LayerNorm: row statistics of the f32 result + its F16 copy
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] -= e_center;   // (the f32 result itself was stored above)
      if (n0 + f0 + i < a.N) {
        s1 += v[i];
        s2 = fmaf(v[i], v[i], s2);
      }
    }
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
    if (r < a.R) {
      if ((tid & 3) == 0) {   // integer atomics: the sum does not depend on the order the CTAs arrive in
        const DecLnStat f = dec_ln_fixed(s1, s2);
        DecLnStat* dst = a.ln_out + (blockIdx.x & (DEC_LN_SUB - 1)) * DEC_LN_ROWS + r;
        atomicAdd(&dst->s1, f.s1);
        atomicAdd(&dst->s2, f.s2);
      }
      if (n0 + f0 + 3 < a.N) {
        uint2 u;
        u.x = pack_h2(v[0], v[1]);
        u.y = pack_h2(v[2], v[3]);
        *reinterpret_cast<uint2*>(a.x16_out + (size_t)r * a.x16_ld + n0 + f0) = u;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (n0 + f0 + i < a.N) a.x16_out[(size_t)r * a.x16_ld + n0 + f0 + i] = __float2half_rn(v[i]);
      }
    }
  }
  if (a.top2) {   // per-CTA top-2 of every sequence over this CTA's 16 vocabulary entries
    float v1 = -INFINITY, v2 = -INFINITY;
    int i1 = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (v[i] > v1) {
        v2 = v1;
        v1 = v[i];
        i1 = n0 + f0 + i;
      } else if (v[i] > v2) {
        v2 = v[i];
      }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {   // merge the row's 4 lanes (lower index wins ties)
      const float u1 = __shfl_xor_sync(0xffffffffu, v1, o), u2 = __shfl_xor_sync(0xffffffffu, v2, o);
      const int j1 = __shfl_xor_sync(0xffffffffu, i1, o);
      if (u1 > v1 || (u1 == v1 && j1 < i1)) {
        v2 = fmaxf(v1, u2);
        v1 = u1;
        i1 = j1;
      } else {
        v2 = fmaxf(v2, u1);
      }
    }
    if ((tid & 3) == 0 && r < a.R) {
      float* dst = a.top2 + ((size_t)r * gridDim.x + blockIdx.x) * 3;
      dst[0] = v1;
      dst[1] = v2;
      dst[2] = __int_as_float(i1);
    }
  }
}

// D6 from the per-CTA top-2 partials the vocabulary projection leaves ([seq][n_part][3])
__global__ void __launch_bounds__(256)
argmax_partials_kernel(const float* __restrict__ part, int n_part, int* __restrict__ next_tok,
                       float* __restrict__ margin_out, int* __restrict__ out_tokens, float* __restrict__ out_margin,
                       int* __restrict__ out_len, int* __restrict__ done, int max_new, int* __restrict__ step_p,
                       int eot, int* __restrict__ n_past_p, int advance_by) {
  pdl_launch_dependents();   // the next kernel of the step may become resident now; it blocks at its own wait
  pdl_wait();                // everything this kernel reads is the previous kernels' output
  __shared__ Top2 sh[8];
  const int s = blockIdx.x;
  const float* p = part + (size_t)s * n_part * 3;
  Top2 t{-INFINITY, -INFINITY, 0x7fffffff};
  for (int i = threadIdx.x; i < n_part; i += blockDim.x) {
    Top2 u{p[3 * i], p[3 * i + 1], __float_as_int(p[3 * i + 2])};
    t = top2_merge(t, u);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Top2 u;
    u.v1 = __shfl_xor_sync(0xffffffffu, t.v1, o);
    u.v2 = __shfl_xor_sync(0xffffffffu, t.v2, o);
    u.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
    t = top2_merge(t, u);
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    Top2 r = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = top2_merge(r, sh[w]);
    next_tok[s] = r.i1;
    if (margin_out) margin_out[s] = r.v1 - r.v2;
    if (out_tokens) {
      const int step = *step_p;
      if (!done[s] && step < max_new) {
        out_tokens[(size_t)s * max_new + step] = r.i1;
        if (out_margin) out_margin[(size_t)s * max_new + step] = r.v1 - r.v2;
        out_len[s] = step + 1;
        if (r.i1 == eot) done[s] = 1;
      }
    }
    step_advance(step_p, n_past_p, advance_by);
  }
}

__global__ void advance_kernel(int* n_past, int add, int* step) {
  pdl_launch_dependents();   // the next kernel of the step may become resident now; it blocks at its own wait
  pdl_wait();                // everything this kernel reads is the previous kernels' output
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    *n_past += add;
    if (step) *step += 1;
  }
}

}  // namespace

cudaError_t launch_embed(const __half* te, const float* pe, const int* tokens, int n_seq, int n_tok,
                         const int* n_past_dev, int d, float* x, cudaStream_t st, DecLnStat* stats, __half* x16,
                         int n_clear_slots, float* center) {
  if (stats && (n_seq * n_tok > DEC_LN_ROWS || !center)) return cudaErrorInvalidValue;
  if (d > 1280) return cudaErrorInvalidValue;
  return launch_pdl(embed_kernel, dim3(n_seq * n_tok), dim3(128), 0, st, te, pe, tokens, n_tok, n_past_dev, d, x, stats,
                    x16, n_clear_slots, center);
}

cudaError_t launch_decode_self_attn(const __half* qkv, int d, __half* kc, __half* vc, int n_seq, int n_tok,
                                    const int* n_past_dev, int n_text_ctx, int H, __half* out, cudaStream_t st) {
  if (n_text_ctx > SELF_MAX_CTX) return cudaErrorInvalidValue;
  return launch_pdl(decode_self_attn_kernel, dim3(H, n_seq), dim3(SELF_THREADS), 0, st, qkv, d, kc, vc, n_tok, n_past_dev,
                    n_text_ctx, out);
}

// key-range splits per (sequence, head): only when there are too few (sequence, head) pairs to fill the SMs.
// (Splitting to even out 384 pairs over 148 SMs x 3 resident CTAs -- 86 % balanced -- was measured: four splits
// of 375 keys run 360 -> 480 us per step; a CTA needs the long key range to keep its loads in flight.)
int decode_cross_splits(int n_seq, int H, int T, int num_sms) {
  int n_split = 1;
  while (n_seq * H * n_split < 2 * num_sms && n_split < 8 && (T + n_split * 2 - 1) / (n_split * 2) >= 128) n_split *= 2;
  return n_split;
}

cudaError_t launch_decode_cross_attn(const __half* q, int d, const __half* k, const __half* v, long long ld_kv,
                                     long long head_stride, int n_seq, int n_tok, int T, int H, __half* out,
                                     float* part_o, float* part_ml, int n_split, int* split_cnt, cudaStream_t st) {
  const int span = (T + n_split - 1) / n_split;
  if (span > CROSS_MAX_SPAN) return cudaErrorInvalidValue;
  // one wave if possible: 3 CTAs per SM with the deeper load queue, else 4 per SM
  const int n_ctas = H * n_seq * n_split;
  if (n_ctas <= 3 * 148 || n_ctas > 4 * 148)
    return launch_pdl(decode_cross_attn_kernel<4, 3>, dim3(H, n_seq, n_split), dim3(CROSS_THREADS), 0, st, q, d, k, v, ld_kv,
                      head_stride, n_tok, T, span, out, part_o, part_ml, n_split, split_cnt);
  return launch_pdl(decode_cross_attn_kernel<3, 4>, dim3(H, n_seq, n_split), dim3(CROSS_THREADS), 0, st, q, d, k, v, ld_kv,
                    head_stride, n_tok, T, span, out, part_o, part_ml, n_split, split_cnt);
}

cudaError_t launch_argmax(const float* logits, int n_seq, int n_vocab, int* next_tok, float* margin, int* out_tokens,
                          float* out_margin, int* out_len, int* done, int max_new, int* step_dev, int eot,
                          cudaStream_t st, int* n_past_dev, int advance_by) {
  argmax_kernel<<<n_seq, 1024, 0, st>>>(logits, n_vocab, next_tok, margin, out_tokens, out_margin, out_len, done,
                                        max_new, step_dev, eot, n_past_dev, advance_by);
  return cudaGetLastError();
}

cudaError_t launch_decode_linear(const DecodeLinear& a, cudaStream_t st) {
  if (a.R < 1 || a.R > 32 || a.N < 1 || a.K % 64 != 0 || a.ldx % 8 != 0) return cudaErrorInvalidValue;
  const dim3 grid((a.N + DL_ROWS - 1) / DL_ROWS);
  // more than one round per CTA and few CTAs (fc2: N = d): the software-pipelined variant.  Its ~170-230 registers per
  // thread keep the NEXT kernel's CTAs (resident early through programmatic dependent launch, to pull their weights)
  // off the SMs it runs on, so it pays while it occupies less than half of them: small (48 CTAs) 0.767 -> 0.728 ms per
  // step, medium (64) 1.81 -> 1.74, large-v3 (80) 2.14 -> 2.20 (measured, tools/dec_groups.py) -- hence the bound
  const bool pipe = (int)grid.x <= 64;
#define WB_DL(U, NW)                                                                                                  \
  if (a.K % (32 * (U) * (NW)) == 0 || ((U) == 4 && (NW) == 4)) {                                                       \
    if (pipe && a.K >= 2 * 32 * (U) * (NW) && (NW) == 8)                                                               \
      return launch_pdl(decode_linear_kernel<U, NW, true>, grid, dim3(32 * (NW)), 0, st, a);                           \
    return launch_pdl(decode_linear_kernel<U, NW>, grid, dim3(32 * (NW)), 0, st, a);                                   \
  }
  // many CTAs (the vocabulary projection: 3242 of them): waves x per-CTA latency sets the time, so the variant
  // with the fewest registers (most resident CTAs) wins over the one with every load in flight
  if ((int)grid.x > 16 * 148 && a.K % 256 == 0) WB_DL(2, 4);
  if (a.K <= 768) {          // 4 warps, every k-block of the CTA in flight at once when K = 128 U
    if (a.K == 768) WB_DL(6, 4);
    if (a.K == 640) WB_DL(5, 4);
    if (a.K == 384) WB_DL(3, 4);
    if (a.K == 256) WB_DL(2, 4);
    if (a.K == 128) WB_DL(1, 4);
    WB_DL(4, 4);             // 512, and any other K (looped)
  }
  if (a.K % 64 == 0) {       // K > 768: 8 warps (kw multiple of 8... the 16-wide tail needs kw % 16 == 0)
    if (a.K % (8 * 32 * 6) == 0) WB_DL(6, 8);   // 1536, 3072, 4608
    if (a.K % (8 * 32 * 5) == 0) WB_DL(5, 8);   // 1280, 2560, 5120
    if (a.K % (8 * 32 * 4) == 0) WB_DL(4, 8);   // 1024, 2048, 4096
  }
  WB_DL(4, 4);
#undef WB_DL
  return cudaErrorInvalidValue;
}
int decode_linear_parts(int N) { return (N + DL_ROWS - 1) / DL_ROWS; }

cudaError_t launch_argmax_partials(const float* part, int n_part, int n_seq, int* next_tok, float* margin,
                                   int* out_tokens, float* out_margin, int* out_len, int* done, int max_new,
                                   int* step_dev, int eot, cudaStream_t st, int* n_past_dev, int advance_by) {
  return launch_pdl(argmax_partials_kernel, dim3(n_seq), dim3(256), 0, st, part, n_part, next_tok, margin, out_tokens,
                    out_margin, out_len, done, max_new, step_dev, eot, n_past_dev, advance_by);
}

// L2 prefetch of a byte range (4 KB per cp.async.bulk.prefetch): the decoder step issues it on a side branch of the step
// graph for the NEXT layer's cross K / V while the current layer's latency-bound linears leave HBM idle (an
// L2::evict_last cache hint on the prefetch changes nothing)
__global__ void l2_prefetch_kernel(const char* p0, const char* p1, long long n_chunks) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n_chunks; i += (long long)gridDim.x * blockDim.x) {
    const char* p = i < n_chunks ? p0 + i * 4096 : p1 + (i - n_chunks) * 4096;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(4096) : "memory");
  }
}

cudaError_t launch_l2_prefetch(const void* p0, const void* p1, size_t bytes_each, cudaStream_t st) {
  const long long n_chunks = (long long)(bytes_each / 4096);
  if (n_chunks <= 0) return cudaSuccess;
  const int threads = 128;
  int blocks = (int)((2 * n_chunks + threads - 1) / threads);
  if (blocks > 32) blocks = 32;
  l2_prefetch_kernel<<<blocks, threads, 0, st>>>(static_cast<const char*>(p0), static_cast<const char*>(p1), n_chunks);
  return cudaGetLastError();
}

cudaError_t launch_advance(int* n_past_dev, int add, int* step_dev, cudaStream_t st) {
  return launch_pdl(advance_kernel, dim3(1), dim3(32), 0, st, n_past_dev, add, step_dev);
}

}  // namespace wb
