// wo_mel.cpp -- restatement of the reference's log-mel front end.  TEST INFRASTRUCTURE ONLY.
//
// Operation-by-operation restatement of src/main.rs:1487-1707 (dft, fft, log_mel_spectrogram,
// clamp_and_normalize, whisper_pcm_to_mel): the same 400 -> 200 -> 100 -> 50 -> 25 recursion
// with O(n^2) DFT leaves, f32 twiddles evaluated per butterfly with cosf/sinf of an f32 angle,
// the same left-to-right f32 evaluation order (this file is compiled with -ffp-contract=off so
// no multiply-add is fused, as rustc does not fuse), per-call heap allocation included.
#include <thread>

#include "wo_common.hpp"

namespace wo {

static const float PI_F32 = 3.14159265358979323846264338327950288f;  // std::f32::consts::PI

// src/main.rs:1487-1502
void dft(const std::vector<float>& inp, std::vector<float>& out) {
  const size_t n = inp.size();
  out.assign(n * 2, 0.0f);
  for (size_t k = 0; k < n; ++k) {
    float re = 0.0f, im = 0.0f;
    for (size_t nv = 0; nv < n; ++nv) {
      float angle = 2.0f * PI_F32 * (float)(k * nv) / (float)n;   // 1495
      re += inp[nv] * cosf(angle);                                 // 1496
      im -= inp[nv] * sinf(angle);                                 // 1497
    }
    out[k * 2] = re;
    out[k * 2 + 1] = im;
  }
}

// src/main.rs:1505-1551
void fft(const std::vector<float>& inp, std::vector<float>& out) {
  const size_t n = inp.size();
  out.assign(n * 2, 0.0f);
  if (n == 1) {
    out[0] = inp[0];
    out[1] = 0.0f;
    return;
  }
  if (n % 2 == 1) {   // 1514-1517
    dft(inp, out);
    return;
  }
  std::vector<float> even, odd;   // 1519-1528
  even.reserve(n / 2);
  odd.reserve(n / 2);
  for (size_t i = 0; i < n; ++i) {
    if (i % 2 == 0) even.push_back(inp[i]);
    else odd.push_back(inp[i]);
  }
  std::vector<float> even_fft(n, 0.0f), odd_fft(n, 0.0f);   // 1530-1531
  fft(even, even_fft);
  fft(odd, odd_fft);
  for (size_t k = 0; k < n / 2; ++k) {   // 1536-1550
    float theta = 2.0f * PI_F32 * (float)k / (float)n;
    float re = cosf(theta);
    float im = -sinf(theta);
    float re_odd = odd_fft[2 * k];
    float im_odd = odd_fft[2 * k + 1];
    out[2 * k] = even_fft[2 * k] + re * re_odd - im * im_odd;
    out[2 * k + 1] = even_fft[2 * k + 1] + re * im_odd + im * re_odd;
    out[2 * (k + n / 2)] = even_fft[2 * k] - re * re_odd + im * im_odd;
    out[2 * (k + n / 2) + 1] = even_fft[2 * k + 1] - re * im_odd - im * re_odd;
  }
}

// src/main.rs:1654-1671
static void clamp_and_normalize(std::vector<float>& mel) {
  double mmax = -1e20;
  for (float v : mel)
    if ((double)v > mmax) mmax = (double)v;
  mmax -= 8.0;
  for (float& v : mel) {
    if ((double)v < mmax) v = (float)mmax;
    v = (v + 4.0f) / 4.0f;
  }
}

// src/main.rs:1554-1652 (speed_up is always false, 1700; n_mel comes from the filterbank
// header instead of the hard-coded WHISPER_N_MEL = 80 of 27/1697 -- SURVEY.md F8)
int pcm_to_mel(orc_ctx* ctx, const float* pcm, size_t n_samples, int n_threads) {
  const int fft_size = 400, fft_step = 160;   // 26, 28
  const Model& m = ctx->model;
  const int n_mel = m.filt_n_mel;
  const int n_fft = 1 + fft_size / 2;         // 1580
  if (m.filt_n_fft != n_fft) return ORC_ERR_UNEXPECTED;
  if (n_threads < 1) n_threads = 1;
  std::vector<float> hann(fft_size);
  const float fft_size_f32 = (float)fft_size;
  for (int i = 0; i < fft_size; ++i)          // 1567-1569
    hann[i] = 0.5f * (1.0f - cosf((2.0f * PI_F32 * (float)i) / fft_size_f32));
  const size_t n_len = n_samples / fft_step;  // 1575
  ctx->mel_n_mel = n_mel;
  ctx->mel_n_len = (int)n_len;
  ctx->mel.assign((size_t)n_mel * n_len, 0.0f);
  float* data = ctx->mel.data();
  const float* filt = m.filters.data();
  auto worker = [&](int ith) {                // 1587-1638
    std::vector<float> fft_in(fft_size, 0.0f), fft_out(2 * fft_size, 0.0f);
    for (size_t i = ith; i < n_len; i += n_threads) {
      const size_t offset = i * fft_step;
      for (int j = 0; j < fft_size; ++j) {    // 1595-1601
        if (offset + j < n_samples) fft_in[j] = hann[j] * pcm[offset + j];
        else fft_in[j] = 0.0f;
      }
      fft(fft_in, fft_out);
      for (int j = 0; j < fft_size; ++j)      // 1603-1606
        fft_out[j] = fft_out[2 * j] * fft_out[2 * j] + fft_out[2 * j + 1] * fft_out[2 * j + 1];
      for (int j = 1; j < fft_size / 2; ++j)  // 1608-1610 (bin fold)
        fft_out[j] += fft_out[fft_size - j];
      for (int j = 0; j < n_mel; ++j) {       // 1620-1634
        float sum = 0.0f;
        for (int k = 0; k < n_fft; ++k) sum += fft_out[k] * filt[j * n_fft + k];
        if (sum < 1e-10f) sum = 1e-10f;
        sum = log10f(sum);
        data[(size_t)j * n_len + i] = sum;
      }
    }
  };
  std::vector<std::thread> works;
  for (int iw = 0; iw < n_threads; ++iw) works.emplace_back(worker, iw);   // 1582-1640
  for (auto& w : works) w.join();                                          // 1642-1644
  clamp_and_normalize(ctx->mel);                                           // 1648
  ctx->chk[ORC_STAGE_MEL * 1000] = abs_sum(ctx->mel.data(), ctx->mel.size());
  return ORC_OK;
}

}  // namespace wo
