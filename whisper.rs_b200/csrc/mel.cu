// mel.cu -- log-mel front end on the GPU (sm_100a, CUDA cores; HBM-bound byte/float work).
//
// Replaces log_mel_spectrogram + clamp_and_normalize (src/main.rs:1554-1671):
//   frame i = hann[j] * pcm[160 i + j], j < 400, zero past the end (1594-1601)
//   400-point DFT (the reference's recursive 400->200->100->50->25 fft, 1505-1551)
//   power, bin fold p[j] += p[400-j] for j in 1..199 (1603-1610)
//   mel[j][i] = log10(max(sum_k p[k] * filt[j][k], 1e-10)) (1620-1634)
//   whole-clip max, x = max(x, max - 8), x = (x + 4) / 4 (1654-1671)
//
// One CTA handles 16 consecutive frames of one clip.  The 16*160+240 samples the frames share
// are read from HBM once (coalesced, vectorised) into shared memory; each frame's real 400-point
// transform is computed as a 200-point complex FFT of the even/odd-packed samples (25-point DFTs
// held in registers, one shared-memory exchange, 8-point DFTs fused with the untangle step), instead
// of the reference's per-butterfly sinf/cosf recursion; twiddles are f64-evaluated tables rounded to
// f32 or compile-time constants.  The filterbank is
// applied sparsely (each slaney filter touches a few bins; exact zeros contribute nothing), the
// per-clip maximum is reduced with warp shuffles and one atomicMax per warp.
#include "ptx.cuh"
#include "wb_kernels.hpp"

namespace wb {

namespace {

constexpr int F = 16;                        // frames per CTA
constexpr int NFFT = 400, HOP = 160, NBIN = 201;
constexpr int SPAN = HOP * (F - 1) + NFFT;   // 2800 samples
constexpr int MEL_THREADS = 256;             // pass A uses F x 8 = 128 of them (one 25-point DFT per thread); passes B (208
                                             // items) and C (F x n_mel) and the staging copies use all
constexpr int MEL_MAX_NZ = 1024, MEL_MAX_MEL = 128;
// the sample span is stored skewed, sample j at j + 16 (j / 160): frame f starts 160 f samples in, a multiple
// of the 32 banks, so the four frames a warp reads together would otherwise hit the same banks
constexpr int SKEW = 16;
__device__ __forceinline__ int pcm_pos(int j) { return j + SKEW * (j / HOP); }
constexpr int SPAN_SKEWED = SPAN + SKEW * ((SPAN + HOP - 1) / HOP);

// shared-memory strides chosen against bank conflicts (ncu, round 2: a third of the kernel's shared-memory wavefronts
// were conflict replays): pass B's lanes read Y'[q][.] for consecutive q -- 9 float2 per k2 (18 words) puts the 13 values of
// q on 13 different even banks (8 float2 = 16 words put them on two); a frame stride of 232 float2 (= 16 mod 32 words)
// keeps pass A's two frames per half-warp on disjoint banks; pass C's 16 lanes read the same bin of 16 frames -- 205
// words per frame (13 mod 32, odd) spreads them over 16 banks (204 = 12 mod 32 gave 8).
constexpr int A_K2 = 9, A_FRAME = 232, PW_STRIDE = NBIN + 4;
struct MelSmem {
  union {                    // the samples are dead once pass A has windowed them: the power spectrum reuses their space
    float pcm[SPAN_SKEWED];
    float pw[F][PW_STRIDE];
  };
  float2 a[F][A_FRAME];      // pass A output: Y'[k2][n1] at [k2 * A_K2 + n1]
  MelTableBlob tab;          // twiddles, window, filterbank taps (wb_kernels.hpp)
};
constexpr int TAB4 = (int)(sizeof(MelTableBlob) / 16);
constexpr int TAB_IT = (TAB4 + MEL_THREADS - 1) / MEL_THREADS;
constexpr int PCM4 = SPAN / 4;
constexpr int PCM_IT = (PCM4 + MEL_THREADS - 1) / MEL_THREADS;

__device__ __forceinline__ float2 cmul(float2 x, float2 y) {
  return make_float2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
}
__device__ __forceinline__ float2 cadd(float2 x, float2 y) { return make_float2(x.x + y.x, x.y + y.y); }
__device__ __forceinline__ float2 csub(float2 x, float2 y) { return make_float2(x.x - y.x, x.y - y.y); }
__device__ __forceinline__ float2 mul_neg_i(float2 x) { return make_float2(x.y, -x.x); }   // x * (-i)

// order-preserving float <-> int encoding for atomicMax
__device__ __forceinline__ int enc_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float dec_ordered(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// W25^m = e^(-2 pi i m / 25), m = 0..16 (the products b c of pass A), f64-evaluated
__device__ constexpr float W25_TAB[17][2] = {{1.000000000e+00f, -0.000000000e+00f}, {9.685831611e-01f, -2.486898872e-01f}, {8.763066800e-01f, -4.817536741e-01f}, {7.289686274e-01f, -6.845471059e-01f}, {5.358267950e-01f, -8.443279255e-01f}, {3.090169944e-01f, -9.510565163e-01f}, {6.279051953e-02f, -9.980267284e-01f}, {-1.873813146e-01f, -9.822872507e-01f}, {-4.257792916e-01f, -9.048270525e-01f}, {-6.374239897e-01f, -7.705132428e-01f}, {-8.090169944e-01f, -5.877852523e-01f}, {-9.297764859e-01f, -3.681245527e-01f}, {-9.921147013e-01f, -1.253332336e-01f}, {-9.921147013e-01f, 1.253332336e-01f}, {-9.297764859e-01f, 3.681245527e-01f}, {-8.090169944e-01f, 5.877852523e-01f}, {-6.374239897e-01f, 7.705132428e-01f}};

// 5-point DFT in registers, W5 = e^(-2 pi i / 5): x[m] <- sum_a x[a] W5^(a m)
__device__ __forceinline__ void dft5(float2 (&x)[5]) {
  constexpr float C1 = 0.30901699437494742410f, C2 = -0.80901699437494742410f;   // cos(2 pi / 5), cos(4 pi / 5)
  constexpr float S1 = 0.95105651629515357212f, S2 = 0.58778525229247312917f;    // sin(2 pi / 5), sin(4 pi / 5)
  const float2 t1 = cadd(x[1], x[4]), t2 = cadd(x[2], x[3]), t3 = csub(x[1], x[4]), t4 = csub(x[2], x[3]);
  const float2 m1 = make_float2(x[0].x + C1 * t1.x + C2 * t2.x, x[0].y + C1 * t1.y + C2 * t2.y);
  const float2 m2 = make_float2(x[0].x + C2 * t1.x + C1 * t2.x, x[0].y + C2 * t1.y + C1 * t2.y);
  const float2 n1 = make_float2(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y);
  const float2 n2 = make_float2(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y);
  x[0] = cadd(x[0], cadd(t1, t2));
  x[1] = cadd(m1, mul_neg_i(n1));                        // m1 - i n1
  x[4] = make_float2(m1.x - n1.y, m1.y + n1.x);          // m1 + i n1
  x[2] = cadd(m2, mul_neg_i(n2));
  x[3] = make_float2(m2.x - n2.y, m2.y + n2.x);
}

// 8-point DFT in registers, W8 = e^(-2 pi i / 8): y[k] = sum_n x[n] W8^(n k)
__device__ __forceinline__ void dft8(const float2 (&x)[8], float2 (&y)[8]) {
  const float r2 = 0.70710678118654752440f;
  const float2 a0 = cadd(x[0], x[4]), a1 = csub(x[0], x[4]), a2 = cadd(x[2], x[6]), a3 = mul_neg_i(csub(x[2], x[6]));
  const float2 a4 = cadd(x[1], x[5]), a5 = csub(x[1], x[5]), a6 = cadd(x[3], x[7]), a7 = mul_neg_i(csub(x[3], x[7]));
  const float2 e0 = cadd(a0, a2), e2 = csub(a0, a2), e1 = cadd(a1, a3), e3 = csub(a1, a3);
  const float2 o0 = cadd(a4, a6);
  float2 o2 = csub(a4, a6), o1 = cadd(a5, a7), o3 = csub(a5, a7);
  o1 = make_float2((o1.x + o1.y) * r2, (o1.y - o1.x) * r2);      // * W8^1 = (1 - i)/sqrt2
  o2 = mul_neg_i(o2);                                            // * W8^2 = -i
  o3 = make_float2((o3.y - o3.x) * r2, -(o3.x + o3.y) * r2);     // * W8^3 = (-1 - i)/sqrt2
  y[0] = cadd(e0, o0); y[1] = cadd(e1, o1); y[2] = cadd(e2, o2); y[3] = cadd(e3, o3);
  y[4] = csub(e0, o0); y[5] = csub(e1, o1); y[6] = csub(e2, o2); y[7] = csub(e3, o3);
}

// untangle the real transform at bin k from Z[k] and conj(Z[200 - k]), power, bin fold (1603-1610)
__device__ __forceinline__ float bin_power(float2 zk, float2 zo, float2 w, int k) {
  const float2 zc = make_float2(zo.x, -zo.y);                       // conj(Z[200 - k])
  const float2 xe = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
  const float2 d = csub(zk, zc);
  const float2 xo = make_float2(0.5f * d.y, -0.5f * d.x);           // (Zk - Zc) / (2i)
  const float2 X = cadd(xe, cmul(w, xo));
  float p = X.x * X.x + X.y * X.y;
  if (k >= 1 && k <= 199) p += p;                                   // p[j] += p[400 - j]
  return p;
}

// The real 400-point transform of a frame = a 200-point complex FFT of z[n] = x[2n] + i x[2n+1] + an untangle
// step.  With n = 8 n2 + n1 and k = k2 + 25 k1:  W200^(nk) = W25^(n2 k2) W200^(n1 k2) W8^(n1 k1), so
//   pass A  thread (frame, n1): windows its 25 samples z[8 n2 + n1], runs the 25-point DFT over n2 entirely in
//           registers (5 x 5, constant twiddles, no index arithmetic), applies W200^(n1 k2) -> shared memory
//   pass B  thread (frame, q), q = 0..12: the 8-point DFTs over n1 for k2 = q and k2 = 25 - q give
//           Z[q + 25 k1] and Z[200 - (q + 25 k1)] -- exactly the pairs the untangle step combines -- so the
//           power spectrum of 16 bins leaves the thread without Z ever being stored
//   pass C  sparse filterbank, log10, store, running maximum
// One shared-memory round trip for the FFT instead of four, ~2.5x fewer instructions per frame than the
// pass-per-radix version it replaces (which was issue-bound: 5 passes x ~4 k cycles per 16 frames).
template <bool I16>
__global__ void __launch_bounds__(MEL_THREADS, 4)   // 64 registers: four CTAs (32 warps) per SM
mel_frames_kernel(const MelTables tb, const void* __restrict__ pcm_v, size_t n_samples, int n_len,
                  float* __restrict__ mel_out, int* __restrict__ clip_max_enc) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  MelSmem& s = *reinterpret_cast<MelSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int clip = blockIdx.y;
  const int i0 = blockIdx.x * F;
  const size_t base = (size_t)i0 * HOP;

  // ---- tables + the shared sample span (HBM read once per CTA).  Every 16-byte load of both is requested before
  // the first store: as one loop per table (load, store, next iteration) this phase was 15 k cycles of serial
  // L2 / HBM round trips, more than the whole transform
  {
    const float4* blob4 = reinterpret_cast<const float4*>(tb.blob);
    float4* tab4 = reinterpret_cast<float4*>(&s.tab);
    float4 tv[TAB_IT];
#pragma unroll
    for (int u = 0; u < TAB_IT; ++u) {
      const int idx = tid + u * MEL_THREADS;
      if (idx < TAB4) tv[u] = __ldg(blob4 + idx);
    }
    bool vec = false;
    if (!I16) {
      const float* pcm = reinterpret_cast<const float*>(pcm_v) + (size_t)clip * n_samples;
      vec = ((reinterpret_cast<uintptr_t>(pcm + base) & 15) == 0) && (base + SPAN <= n_samples);
      if (vec) {
        const float4* p4 = reinterpret_cast<const float4*>(pcm + base);
        float4 pv[PCM_IT];
#pragma unroll
        for (int u = 0; u < PCM_IT; ++u) {
          const int idx = tid + u * MEL_THREADS;
          if (idx < PCM4) pv[u] = __ldg(p4 + idx);
        }
#pragma unroll
        for (int u = 0; u < PCM_IT; ++u) {
          const int idx = tid + u * MEL_THREADS;   // 4 consecutive samples never straddle a 160-sample block
          if (idx < PCM4) *reinterpret_cast<float4*>(&s.pcm[pcm_pos(4 * idx)]) = pv[u];
        }
      } else {   // the clip's last frames (zero past the end, 1596-1600) or an unaligned clip
        for (int i = tid; i < SPAN; i += MEL_THREADS) s.pcm[pcm_pos(i)] = (base + i < n_samples) ? __ldg(pcm + base + i) : 0.0f;
      }
    } else {
      const int16_t* pcm = reinterpret_cast<const int16_t*>(pcm_v) + (size_t)clip * n_samples;
      // convert_integer_to_float_audio (1673-1679): s / 32768.0 (exact in f32)
      for (int i = tid; i < SPAN; i += MEL_THREADS)
        s.pcm[pcm_pos(i)] = (base + i < n_samples) ? (float)pcm[base + i] * (1.0f / 32768.0f) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < TAB_IT; ++u) {
      const int idx = tid + u * MEL_THREADS;
      if (idx < TAB4) tab4[idx] = tv[u];
    }
  }
  __syncthreads();

  // ---- pass A: window + 25-point DFT over n2 in registers + twiddle W200^(n1 k2)
  if (tid < F * 8) {
    const int f = tid >> 3, n1 = tid & 7;
    float2 z[25];
#pragma unroll
    for (int n2 = 0; n2 < 25; ++n2) {
      const int j = 16 * n2 + 2 * n1;                       // even sample index in the frame
      const float2 pv = *reinterpret_cast<const float2*>(&s.pcm[pcm_pos(f * HOP + j)]);
      const float2 hv = *reinterpret_cast<const float2*>(&s.tab.hann[j]);
      z[n2] = make_float2(hv.x * pv.x, hv.y * pv.y);        // z = x[2n] + i x[2n+1], n = 8 n2 + n1
    }
    // n2 = 5 a + b, k2 = c + 5 e:  W25^(n2 k2) = W5^(a c) W25^(b c) W5^(b e)
    float2 t[5][5];   // t[b][c]
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      float2 x[5] = {z[b], z[5 + b], z[10 + b], z[15 + b], z[20 + b]};
      dft5(x);
#pragma unroll
      for (int c = 0; c < 5; ++c) t[b][c] = x[c];
    }
#pragma unroll
    for (int b = 1; b < 5; ++b)
#pragma unroll
      for (int c = 1; c < 5; ++c) {
        t[b][c] = cmul(t[b][c], make_float2(W25_TAB[b * c][0], W25_TAB[b * c][1]));   // compile-time constants
      }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      float2 x[5] = {t[0][c], t[1][c], t[2][c], t[3][c], t[4][c]};
      dft5(x);
#pragma unroll
      for (int e = 0; e < 5; ++e) {
        const int k2 = c + 5 * e;
        const float2 y = (k2 == 0) ? x[e] : cmul(x[e], s.tab.w200[n1 * k2]);   // n1 k2 <= 7 * 24
        s.a[f][k2 * A_K2 + n1] = y;
      }
    }
  }
  __syncthreads();
  // ---- pass B: 8-point DFTs over n1 for k2 = q and 25 - q, untangle, power, bin fold
  for (int it = tid; it < F * 13; it += MEL_THREADS) {
    const int f = it / 13, q = it - f * 13;
    float2 xa[8], za[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) xa[n1] = s.a[f][q * A_K2 + n1];
    dft8(xa, za);                                           // za[k1] = Z[q + 25 k1]
    float* pw = s.pw[f];
    if (q == 0) {
      // bins 25 k1 pair with 25 (8 - k1); bin 0 pairs with itself and also yields bin 200 (Z[200] = Z[0])
      pw[0] = bin_power(za[0], za[0], s.tab.w400[0], 0);
      pw[200] = bin_power(za[0], za[0], s.tab.w400[200], 200);
#pragma unroll
      for (int k1 = 1; k1 < 8; ++k1) pw[25 * k1] = bin_power(za[k1], za[8 - k1], s.tab.w400[25 * k1], 25 * k1);
    } else {
      float2 xb[8], zb[8];
#pragma unroll
      for (int n1 = 0; n1 < 8; ++n1) xb[n1] = s.a[f][(25 - q) * A_K2 + n1];
      dft8(xb, zb);                                         // zb[k1] = Z[(25 - q) + 25 k1]
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) {
        const int k = q + 25 * k1, kc = 200 - k;            // kc = (25 - q) + 25 (7 - k1)
        pw[k] = bin_power(za[k1], zb[7 - k1], s.tab.w400[k], k);
        pw[kc] = bin_power(zb[7 - k1], za[k1], s.tab.w400[kc], kc);
      }
    }
  }
  __syncthreads();
  // ---- pass C: sparse filterbank, clamp, log10, store mel-major, running max
  float* out = mel_out + (size_t)clip * tb.n_mel * n_len;
  float tmax = -INFINITY;
  for (int it = tid; it < F * tb.n_mel; it += MEL_THREADS) {
    const int j = it / F, f = it - j * F;
    const int i = i0 + f;
    if (i >= n_len) continue;
    // taps in bin order, as the reference's sequential f32 sum (exact zeros contribute nothing)
    const int2 rg = s.tab.frange[j];
    const float* pw = &s.pw[f][rg.x];
    const float* fw = &s.tab.fw[rg.y];
    const int nt = s.tab.fcount[j];
    float sum = 0.0f;
    for (int k = 0; k < nt; ++k) sum = fmaf(pw[k], fw[k], sum);
    sum = fmaxf(sum, 1e-10f);
    const float v = __log10f(sum);   // lg2.approx * log10(2): abs. error ~1e-7, far inside the 1e-4 the front end is held to
    out[(size_t)j * n_len + i] = v;
    tmax = fmaxf(tmax, v);
  }
  tmax = warp_max(tmax);
  if ((tid & 31) == 0 && tmax > -INFINITY) atomicMax(&clip_max_enc[clip], enc_ordered(tmax));
}

// clamp_and_normalize (1654-1671) with the per-clip maximum found above
__global__ void mel_normalize_kernel(float* __restrict__ mel, size_t per_clip, const int* __restrict__ clip_max_enc) {
  const int clip = blockIdx.y;
  const float mmax = dec_ordered(clip_max_enc[clip]) - 8.0f;
  float* p = mel + (size_t)clip * per_clip;
  const size_t n4 = per_clip / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  const bool vec = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    for (size_t i = idx; i < n4; i += stride) {
      float4 v = p4[i];
      v.x = (fmaxf(v.x, mmax) + 4.0f) / 4.0f;
      v.y = (fmaxf(v.y, mmax) + 4.0f) / 4.0f;
      v.z = (fmaxf(v.z, mmax) + 4.0f) / 4.0f;
      v.w = (fmaxf(v.w, mmax) + 4.0f) / 4.0f;
      p4[i] = v;
    }
    for (size_t i = n4 * 4 + idx; i < per_clip; i += stride) p[i] = (fmaxf(p[i], mmax) + 4.0f) / 4.0f;
  } else {
    for (size_t i = idx; i < per_clip; i += stride) p[i] = (fmaxf(p[i], mmax) + 4.0f) / 4.0f;
  }
}

// WB_NORM_SEGMENT: the maximum of one encoder window of a clip's log10 mel (frames past the clip end do not count:
// they are zero-filled after normalisation).  grid (chunks, n_seg); seg_max_enc starts at enc(-1e20).
__global__ void __launch_bounds__(256)
mel_window_max_kernel(const float* __restrict__ mel, int n_mel, int n_len, const int* __restrict__ clip_ids,
                      const long long* __restrict__ offsets, int Tm, int* __restrict__ seg_max_enc) {
  const int seg = blockIdx.y;
  const float* src = mel + (size_t)(clip_ids ? clip_ids[seg] : 0) * n_mel * n_len;
  const long long off = offsets ? offsets[seg] : 0;
  const long long n_in = min((long long)Tm, (long long)n_len - off);   // frames of the window inside the clip
  float m = -INFINITY;
  if (n_in > 0) {
    const long long total = (long long)n_mel * n_in;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long j = i / n_in, t = i - j * n_in;
      m = fmaxf(m, src[(size_t)j * n_len + off + t]);
    }
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > -INFINITY) atomicMax(&seg_max_enc[seg], enc_ordered(m));
}

__global__ void fill_i32_kernel(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

int mel_enc_ordered_host(float f) {
  int i;
  memcpy(&i, &f, 4);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
float mel_dec_ordered_host(int i) {
  i = i >= 0 ? i : i ^ 0x7FFFFFFF;
  float f;
  memcpy(&f, &i, 4);
  return f;
}

cudaError_t launch_mel_frames(const MelTables& t, const void* pcm, int pcm_is_i16, size_t n_samples, int n_clips,
                              int n_len, float* mel_out, int* clip_max_enc, cudaStream_t st) {
  if (n_len <= 0 || n_clips <= 0) return cudaSuccess;
  // opt-in to > 48 KB of dynamic shared memory, once (thread-safe function-local static)
  static const cudaError_t attr_err = [] {
    cudaError_t e = cudaFuncSetAttribute(mel_frames_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(MelSmem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(mel_frames_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MelSmem));
  }();
  if (attr_err != cudaSuccess) return attr_err;
  dim3 grid((n_len + F - 1) / F, n_clips);
  if (pcm_is_i16)
    mel_frames_kernel<true><<<grid, MEL_THREADS, sizeof(MelSmem), st>>>(t, pcm, n_samples, n_len, mel_out, clip_max_enc);
  else
    mel_frames_kernel<false><<<grid, MEL_THREADS, sizeof(MelSmem), st>>>(t, pcm, n_samples, n_len, mel_out, clip_max_enc);
  return cudaGetLastError();
}

cudaError_t launch_mel_normalize(float* mel, int n_clips, size_t per_clip, const int* clip_max_enc, cudaStream_t st) {
  if (n_clips <= 0 || per_clip == 0) return cudaSuccess;
  int bx = (int)((per_clip / 4 + 255) / 256);
  if (bx > 592) bx = 592;   // 4 x 148 SMs, grid-stride
  if (bx < 1) bx = 1;
  mel_normalize_kernel<<<dim3(bx, n_clips), 256, 0, st>>>(mel, per_clip, clip_max_enc);
  return cudaGetLastError();
}

cudaError_t launch_mel_window_max(const float* mel, int n_mel, int n_len, const int* clip_ids, const long long* offsets,
                                  int n_seg, int Tm, int* seg_max_enc, cudaStream_t st) {
  if (n_seg <= 0) return cudaSuccess;
  mel_window_max_kernel<<<dim3(32, n_seg), 256, 0, st>>>(mel, n_mel, n_len, clip_ids, offsets, Tm, seg_max_enc);
  return cudaGetLastError();
}

cudaError_t launch_fill_i32(int* p, int n, int v, cudaStream_t st) {
  fill_i32_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, n, v);
  return cudaGetLastError();
}

}  // namespace wb
