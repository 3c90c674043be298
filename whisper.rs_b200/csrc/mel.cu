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
// transform is computed as a 200-point complex FFT of the even/odd-packed samples (radix 8 x 5 x 5
// passes through shared memory) plus an untangle pass, instead of the reference's per-butterfly
// sinf/cosf recursion; twiddles are f64-evaluated tables rounded to f32.  The filterbank is
// applied sparsely (each slaney filter touches a few bins; exact zeros contribute nothing), the
// per-clip maximum is reduced with warp shuffles and one atomicMax per warp.
#include "ptx.cuh"
#include "wb_kernels.hpp"

namespace wb {

namespace {

constexpr int F = 16;                        // frames per CTA
constexpr int NFFT = 400, HOP = 160, NBIN = 201;
constexpr int SPAN = HOP * (F - 1) + NFFT;   // 2800 samples
constexpr int MEL_THREADS = 256;             // (128 threads per CTA measured slightly slower: 115 vs 108 us)
constexpr int MEL_MAX_NZ = 1024, MEL_MAX_MEL = 128;

struct MelSmem {
  float pcm[SPAN];
  float2 a[F][200];
  union {                    // b is dead once pass 3 has written a: the power spectrum reuses its space
    float2 b[F][200];
    float pw[F][NBIN + 1];
  };
  float2 w200[200];
  float2 w400[NBIN];
  float hann[NFFT];
  float fw[MEL_MAX_NZ];     // non-zero filterbank taps, mel after mel (each bin feeds at most two slaney filters)
  int2 frange[MEL_MAX_MEL]; // per mel: first bin, first tap index in fw (taps = bins frange[j].x .. of the next start)
  int fcount[MEL_MAX_MEL];
};

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

template <bool I16>
__global__ void __launch_bounds__(MEL_THREADS)
mel_frames_kernel(const MelTables tb, const void* __restrict__ pcm_v, size_t n_samples, int n_len,
                  float* __restrict__ mel_out, int* __restrict__ clip_max_enc) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  MelSmem& s = *reinterpret_cast<MelSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int clip = blockIdx.y;
  const int i0 = blockIdx.x * F;
  const size_t base = (size_t)i0 * HOP;

  // ---- tables + the shared sample span (HBM read once per CTA)
  for (int i = tid; i < 200; i += MEL_THREADS) s.w200[i] = tb.tw200[i];
  for (int i = tid; i < NBIN; i += MEL_THREADS) s.w400[i] = tb.tw400[i];
  for (int i = tid; i < NFFT; i += MEL_THREADS) s.hann[i] = tb.hann[i];
  for (int i = tid; i < tb.n_nz; i += MEL_THREADS) s.fw[i] = tb.filt_nz[i];
  for (int i = tid; i < tb.n_mel; i += MEL_THREADS) {
    const int2 rg = tb.filt_range[i];
    s.frange[i] = make_int2(rg.x, tb.filt_start[i]);
    s.fcount[i] = rg.y - rg.x;
  }
  if (I16) {
    const int16_t* pcm = reinterpret_cast<const int16_t*>(pcm_v) + (size_t)clip * n_samples;
    // convert_integer_to_float_audio (1673-1679): s / 32768.0 (exact in f32)
    for (int i = tid; i < SPAN; i += MEL_THREADS)
      s.pcm[i] = (base + i < n_samples) ? (float)pcm[base + i] * (1.0f / 32768.0f) : 0.0f;
  } else {
    const float* pcm = reinterpret_cast<const float*>(pcm_v) + (size_t)clip * n_samples;
    const bool vec = ((reinterpret_cast<uintptr_t>(pcm + base) & 15) == 0) && (base + SPAN <= n_samples);
    if (vec) {
      const float4* p4 = reinterpret_cast<const float4*>(pcm + base);
      float4* s4 = reinterpret_cast<float4*>(s.pcm);
      for (int i = tid; i < SPAN / 4; i += MEL_THREADS) s4[i] = __ldg(p4 + i);
    } else {
      for (int i = tid; i < SPAN; i += MEL_THREADS) s.pcm[i] = (base + i < n_samples) ? __ldg(pcm + base + i) : 0.0f;
    }
  }
  __syncthreads();

  // ---- pass 1: window + radix-8 over n1 (n = 25 n1 + n2), twiddle W200^(n2 k1) -> a[f][k1*25 + n2]
  for (int it = tid; it < F * 25; it += MEL_THREADS) {
    const int f = it / 25, n2 = it - f * 25;
    float2 x[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const int j = 2 * (25 * n1 + n2);                       // even sample index in the frame
      const float2 pv = *reinterpret_cast<const float2*>(&s.pcm[f * HOP + j]);
      const float2 hv = *reinterpret_cast<const float2*>(&s.hann[j]);
      x[n1] = make_float2(hv.x * pv.x, hv.y * pv.y);          // z = x[2n] + i x[2n+1]
    }
    const float r2 = 0.70710678118654752440f;
    float2 a0 = cadd(x[0], x[4]), a1 = csub(x[0], x[4]), a2 = cadd(x[2], x[6]), a3 = mul_neg_i(csub(x[2], x[6]));
    float2 a4 = cadd(x[1], x[5]), a5 = csub(x[1], x[5]), a6 = cadd(x[3], x[7]), a7 = mul_neg_i(csub(x[3], x[7]));
    float2 e0 = cadd(a0, a2), e2 = csub(a0, a2), e1 = cadd(a1, a3), e3 = csub(a1, a3);
    float2 o0 = cadd(a4, a6), o2 = csub(a4, a6), o1 = cadd(a5, a7), o3 = csub(a5, a7);
    o1 = make_float2((o1.x + o1.y) * r2, (o1.y - o1.x) * r2);      // * W8^1 = (1 - i)/sqrt2
    o2 = mul_neg_i(o2);                                            // * W8^2 = -i
    o3 = make_float2((o3.y - o3.x) * r2, -(o3.x + o3.y) * r2);     // * W8^3 = (-1 - i)/sqrt2
    float2 y[8] = {cadd(e0, o0), cadd(e1, o1), cadd(e2, o2), cadd(e3, o3),
                   csub(e0, o0), csub(e1, o1), csub(e2, o2), csub(e3, o3)};
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
      const float2 w = s.w200[n2 * k1];   // n2 k1 <= 24 * 7 < 200
      s.a[f][k1 * 25 + n2] = (k1 == 0) ? y[0] : cmul(y[k1], w);
    }
  }
  __syncthreads();

  // W5^m as W200^(40 m)
  // ---- pass 2: 25-point DFT over n2 = 5a + b, first radix 5 over a, twiddle W25^(b c) -> b[f][k1*25 + b*5 + c]
  for (int it = tid; it < F * 40; it += MEL_THREADS) {
    const int f = it / 40, rr = it - f * 40, k1 = rr / 5, bb = rr - k1 * 5;
    float2 x[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) x[a] = s.a[f][k1 * 25 + 5 * a + bb];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      float2 acc = x[0];
#pragma unroll
      for (int a = 1; a < 5; ++a) acc = cadd(acc, cmul(x[a], s.w200[40 * ((a * c) % 5)]));
      if (c != 0 && bb != 0) acc = cmul(acc, s.w200[8 * bb * c]);   // W25^(b c) = W200^(8 b c), 8 b c <= 128
      s.b[f][k1 * 25 + bb * 5 + c] = acc;
    }
  }
  __syncthreads();
  // ---- pass 3: radix 5 over b -> Z[k1 + 8 (c + 5 e)] in natural order -> a[f][k]
  for (int it = tid; it < F * 40; it += MEL_THREADS) {
    const int f = it / 40, rr = it - f * 40, k1 = rr / 5, c = rr - k1 * 5;
    float2 x[5];
#pragma unroll
    for (int bb = 0; bb < 5; ++bb) x[bb] = s.b[f][k1 * 25 + bb * 5 + c];
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      float2 acc = x[0];
#pragma unroll
      for (int bb = 1; bb < 5; ++bb) acc = cadd(acc, cmul(x[bb], s.w200[40 * ((bb * e) % 5)]));
      s.a[f][k1 + 8 * (c + 5 * e)] = acc;
    }
  }
  __syncthreads();
  // ---- pass 4: untangle the real transform, power, bin fold
  for (int it = tid; it < F * NBIN; it += MEL_THREADS) {
    const int f = it / NBIN, k = it - f * NBIN;
    const float2 zk = s.a[f][k == 200 ? 0 : k];
    float2 zc = s.a[f][k == 0 ? 0 : 200 - k];
    zc.y = -zc.y;                                                  // conj(Z[200 - k])
    const float2 xe = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
    const float2 d = csub(zk, zc);
    const float2 xo = make_float2(0.5f * d.y, -0.5f * d.x);        // (Zk - Zc) / (2i)
    const float2 X = cadd(xe, cmul(s.w400[k], xo));
    float p = X.x * X.x + X.y * X.y;
    if (k >= 1 && k <= 199) p += p;                                // p[j] += p[400 - j]  (1608-1610)
    s.pw[f][k] = p;
  }
  __syncthreads();
  // ---- pass 5: sparse filterbank, clamp, log10, store mel-major, running max
  float* out = mel_out + (size_t)clip * tb.n_mel * n_len;
  float tmax = -INFINITY;
  for (int it = tid; it < F * tb.n_mel; it += MEL_THREADS) {
    const int j = it / F, f = it - j * F;
    const int i = i0 + f;
    if (i >= n_len) continue;
    // taps in bin order, as the reference's sequential f32 sum (exact zeros contribute nothing)
    const int2 rg = s.frange[j];
    const float* pw = &s.pw[f][rg.x];
    const float* fw = &s.fw[rg.y];
    const int nt = s.fcount[j];
    float sum = 0.0f;
    for (int k = 0; k < nt; ++k) sum = fmaf(pw[k], fw[k], sum);
    sum = fmaxf(sum, 1e-10f);
    const float v = log10f(sum);
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

cudaError_t launch_fill_i32(int* p, int n, int v, cudaStream_t st) {
  fill_i32_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, n, v);
  return cudaGetLastError();
}

}  // namespace wb
