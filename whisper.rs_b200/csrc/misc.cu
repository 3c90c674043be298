// misc.cu -- small fused kernels around the GEMMs (sm_100a, CUDA cores, HBM-bound).
#include "ptx.cuh"
#include "wb_kernels.hpp"

namespace wb {

namespace {

// ---------------------------------------------------------------------------------------------
// E0: mel window copy (src/main.rs:1816-1829) fused with the F16 rounding ggml's conv applies to
// its source and with the transpose to token-major rows the conv-as-GEMM reads through TMA:
//   in  mel [clip][n_mel][n_len] f32 (time contiguous)
//   out [seg][Tm + 2][n_mel] f16, rows 0 and Tm+1 stay zero (the conv's zero padding)
// (all mel rows) x 32-frame tiles through shared memory so both the read (along time) and the write
// (along mel) are coalesced.
// clamp_and_normalize (1654-1671) is applied on the way through when the mel still holds log10 values
// (norm_mode 1: the maximum of the segment's clip, 2: the segment's own maximum): x = max(x, max - 8), (x + 4) / 4
// -- the same two f32 operations mel_normalize_kernel applies in place, so both routes give the same bits;
// frames past the clip end stay 0 (the window is zero-filled AFTER normalisation in the reference, 1816-1829).
constexpr int MW_MAX_MEL = 128;
__global__ void __launch_bounds__(256)
mel_window_kernel(const float* __restrict__ mel, int n_mel, int n_len, const int* __restrict__ clip_ids,
                  const long long* __restrict__ offsets, int Tm, __half* __restrict__ out,
                  const int* __restrict__ max_enc, int norm_mode) {
  // a tile = every mel row x 32 frames: the 32 output rows (n_mel halves each) are one contiguous run of memory
  __shared__ float tile[MW_MAX_MEL][33];
  pdl_launch_dependents();   // the next kernel may become resident now; it blocks at its own wait
  pdl_wait();
  const int seg = blockIdx.z;
  const int clip = clip_ids ? clip_ids[seg] : 0;
  const long long off = offsets ? offsets[seg] : 0;
  const float* src = mel + (size_t)clip * n_mel * n_len;
  const int t0 = blockIdx.x * 32;
  const int t_in = t0 + threadIdx.x;
  const long long i_in = off + t_in;
  const bool in_ok = t_in < Tm && i_in < n_len;   // zero past the clip end (1820-1828)
  if (norm_mode) {
    const int e = max_enc[norm_mode == 1 ? clip : seg];
    const float mmax = __int_as_float(e >= 0 ? e : e ^ 0x7FFFFFFF) - 8.0f;   // ordered-int encoding of mel.cu
    for (int j = threadIdx.y; j < n_mel; j += blockDim.y)
      tile[j][threadIdx.x] = in_ok ? (fmaxf(src[(size_t)j * n_len + i_in], mmax) + 4.0f) / 4.0f : 0.0f;
  } else {
    for (int j = threadIdx.y; j < n_mel; j += blockDim.y) tile[j][threadIdx.x] = in_ok ? src[(size_t)j * n_len + i_in] : 0.0f;
  }
  __syncthreads();
  __half* dst = out + (size_t)seg * (Tm + 2) * n_mel;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = t0 + r;
    if (t >= Tm) break;
    for (int j = threadIdx.x; j < n_mel; j += 32) dst[(size_t)(t + 1) * n_mel + j] = __float2half_rn(tile[j][r]);
  }
}

// ---------------------------------------------------------------------------------------------
// E4: galois_norm + repeat/mul/add (src/main.rs:1781-1785, 1882-1886): one warp per row of d,
// the row held in registers (two-pass mean / variance, eps = 1e-5, biased variance), affine,
// F16 store = the rounding the following matmul applies to its activation operand.
constexpr int LN_MAX_V4 = 10;   // d <= 1280

// NV4 = float4 per lane (ceil(d / 128)), RPW = rows per warp (all of a warp's rows are requested
// before the first is reduced).
template <int NV4, int RPW>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, long long in_row_stride, const float* __restrict__ w,
                 const float* __restrict__ b, int rows, int d, __half* __restrict__ out_f16,
                 float* __restrict__ out_f32, float* __restrict__ center_out) {
  pdl_launch_dependents();   // the next kernel may become resident now; it blocks at its own wait
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = warp * RPW;
  if (row0 >= rows) return;
  const int nv4 = d >> 2;   // float4 per row
  float4 v[RPW][NV4];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)min(row0 + r, rows - 1) * in_row_stride);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv4) v[r][i] = xr[idx];
    }
  }
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    if (row0 + r >= rows) break;
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NV4; ++i)
      if (lane + 32 * i < nv4) sum += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
    const float mean = warp_sum(sum) / (float)d;
    if (center_out && lane == 0) center_out[row0 + r] = mean;   // the centre the folded LayerNorms downstream start from
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < NV4; ++i)
      if (lane + 32 * i < nv4) {
        float4& t = v[r][i];
        t.x -= mean; t.y -= mean; t.z -= mean; t.w -= mean;
        sq += (t.x * t.x + t.y * t.y) + (t.z * t.z + t.w * t.w);
      }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)d + 1e-5f);
    const size_t orow = (size_t)(row0 + r) * d;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv4) {
        const float4 ww = __ldg(w4 + idx), bb = __ldg(b4 + idx);
        float4 y;
        y.x = ww.x * (v[r][i].x * rstd) + bb.x;
        y.y = ww.y * (v[r][i].y * rstd) + bb.y;
        y.z = ww.z * (v[r][i].z * rstd) + bb.z;
        y.w = ww.w * (v[r][i].w * rstd) + bb.w;
        if (out_f32) reinterpret_cast<float4*>(out_f32 + orow)[idx] = y;
        if (out_f16) {
          uint2 u;
          u.x = pack_h2(y.x, y.y);
          u.y = pack_h2(y.z, y.w);
          reinterpret_cast<uint2*>(out_f16 + orow)[idx] = u;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// sum|x| probes (the author's checkpoints, src/main.rs:1836-1849): one double per segment.
// grid (ABS_BLOCKS, n_seg): every block leaves its partial sum in its own slot and the block of a segment that
// finishes last adds the slots in slot order -- the result is bit-reproducible from run to run (an atomicAdd of
// the partials is not: the order the blocks finish in changes).
constexpr int ABS_BLOCKS = 64;

__device__ __forceinline__ void abs_sum_finish(double acc, double* __restrict__ part, unsigned* __restrict__ cnt,
                                               double* __restrict__ out) {
  const int seg = blockIdx.y;
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[(size_t)seg * ABS_BLOCKS + blockIdx.x] = sh[0];
    __threadfence();
    const unsigned old = atomicAdd(&cnt[seg], 1u);
    if (old == gridDim.x - 1) {
      __threadfence();
      double tot = 0.0;
      for (int i = 0; i < (int)gridDim.x; ++i) tot += __ldcg(part + (size_t)seg * ABS_BLOCKS + i);
      out[seg] = tot;
      cnt[seg] = 0;   // ready for the next launch
    }
  }
}

__global__ void __launch_bounds__(256)
abs_sum_f32_kernel(const float* __restrict__ x, long long per_seg, long long seg_stride, double* __restrict__ out,
                   double* __restrict__ part, unsigned* __restrict__ cnt) {
  const float* p = x + (size_t)blockIdx.y * seg_stride;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_seg; i += (long long)gridDim.x * blockDim.x)
    acc += (double)fabsf(p[i]);
  abs_sum_finish(acc, part, cnt, out);
}

__global__ void __launch_bounds__(256)
abs_sum_f16_kernel(const __half* __restrict__ x, int rows, int cols, long long row_stride, long long seg_stride,
                   double* __restrict__ out, double* __restrict__ part, unsigned* __restrict__ cnt) {
  const __half* p = x + (size_t)blockIdx.y * seg_stride;
  const long long n = (long long)rows * cols;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    acc += (double)fabsf(__half2float(p[r * row_stride + c]));
  }
  abs_sum_finish(acc, part, cnt, out);
}

}  // namespace

cudaError_t launch_mel_window(const float* mel, int n_mel, int n_len, const int* clip_ids, const long long* offsets,
                              int n_seg, int Tm, __half* out, cudaStream_t st, const int* max_enc, int norm_mode) {
  if (n_mel > MW_MAX_MEL || (norm_mode && !max_enc)) return cudaErrorInvalidValue;
  dim3 grid((Tm + 31) / 32, 1, n_seg);
  return launch_pdl(mel_window_kernel, grid, dim3(32, 8), 0, st, mel, n_mel, n_len, clip_ids, offsets, Tm, out, max_enc,
                    norm_mode);
}

cudaError_t launch_layernorm(const float* x, const float* w, const float* b, int rows, int d, __half* out_f16,
                             float* out_f32, cudaStream_t st, long long in_row_stride, bool pdl, float* center_out) {
  if (d % 4 != 0 || d > 128 * LN_MAX_V4) return cudaErrorInvalidValue;
  if (in_row_stride <= 0) in_row_stride = d;
  const int nv = (d / 4 + 31) / 32;   // float4 per lane
  auto go = [&](auto kernel, int rpw) {
    const int warps = (rows + rpw - 1) / rpw;
    if (pdl) launch_pdl(kernel, dim3((warps + 7) / 8), dim3(256), 0, st, x, in_row_stride, w, b, rows, d, out_f16, out_f32, center_out);
    else kernel<<<(warps + 7) / 8, 256, 0, st>>>(x, in_row_stride, w, b, rows, d, out_f16, out_f32, center_out);
  };
  // one row per warp: measured on B200 (base, 24000 rows of 512), 2 and 4 rows per warp in flight were
  // slower (250 vs 224 us per step over 13 launches) -- the rows mostly come out of L2, where the previous
  // GEMM left them
  if (nv <= 4) go(layernorm_kernel<4, 1>, 1);
  else if (nv <= 8) go(layernorm_kernel<8, 1>, 1);
  else go(layernorm_kernel<LN_MAX_V4, 1>, 1);
  return cudaGetLastError();
}

size_t abs_sum_scratch_bytes(int max_seg) { return (size_t)max_seg * (ABS_BLOCKS * sizeof(double) + sizeof(unsigned)); }

// scratch: abs_sum_scratch_bytes(scratch_segs) bytes, zeroed once at allocation ([scratch_segs][ABS_BLOCKS] partials,
// then the arrival counters)
cudaError_t launch_abs_sum_f32(const float* x, long long per_seg, long long seg_stride, int n_seg, double* out,
                               void* scratch, int scratch_segs, cudaStream_t st) {
  if (n_seg > scratch_segs) return cudaErrorInvalidValue;
  double* part = reinterpret_cast<double*>(scratch);
  unsigned* cnt = reinterpret_cast<unsigned*>(part + (size_t)scratch_segs * ABS_BLOCKS);
  abs_sum_f32_kernel<<<dim3(ABS_BLOCKS, n_seg), 256, 0, st>>>(x, per_seg, seg_stride, out, part, cnt);
  return cudaGetLastError();
}

cudaError_t launch_abs_sum_f16(const __half* x, int rows, int cols, long long row_stride, long long seg_stride,
                               int n_seg, double* out, void* scratch, int scratch_segs, cudaStream_t st) {
  if (n_seg > scratch_segs) return cudaErrorInvalidValue;
  double* part = reinterpret_cast<double*>(scratch);
  unsigned* cnt = reinterpret_cast<unsigned*>(part + (size_t)scratch_segs * ABS_BLOCKS);
  abs_sum_f16_kernel<<<dim3(ABS_BLOCKS, n_seg), 256, 0, st>>>(x, rows, cols, row_stride, seg_stride, out, part, cnt);
  return cudaGetLastError();
}

}  // namespace wb
