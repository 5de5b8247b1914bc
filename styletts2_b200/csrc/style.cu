// Zero-shot reference-style path: log-mel front-end pieces and the 2-D
// convolution stack of StyleEncoder (models.py:27-164, Demo/Inference_LibriTTS.ipynb cell 5).
//
// This path runs once per reference clip (a few GFLOP), not per utterance, so it is plain fp32 SIMT:
// a register-tiled direct conv2d (64 output channels x 8x16 pixels per CTA, input column kept in
// registers and slid over the kernel rows), a depthwise stride-2 conv, the 2x2 mean with the
// reference's odd-width replicate rule, and the framing / power / log stages around the two GEMMs
// (DFT basis and HTK filterbank) that go through st2_linear / st2_linear_tc.
#include "common.cuh"

namespace st2 {
extern long long g_launches;

// ------------------------------------------------------------------------------------------------
// spectral_norm fold (eval mode): sigma = u . (W_mat v);  wt[(ci*KK + k) * Cout + co] = W[co][ci*KK + k] / sigma
__global__ void __launch_bounds__(1024) sn_sigma_kernel(const float* __restrict__ w, const float* __restrict__ u,
                                                        const float* __restrict__ v, int rows, int cols,
                                                        float* __restrict__ sigma) {
  __shared__ double part[32];
  double acc = 0.0;
  const long long n = (long long)rows * cols;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    acc += (double)u[r] * (double)w[i] * (double)v[c];
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
    t = warp_sum_d(t);
    if (threadIdx.x == 0) *sigma = (float)t;
  }
}

__global__ void sn_layout_kernel(const float* __restrict__ w, const float* __restrict__ sigma, int Cout, int n,
                                 float* __restrict__ wt) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * n) return;
  const int co = (int)(i % Cout);
  const int r = (int)(i / Cout);
  wt[i] = __fdiv_rn(w[(long long)co * n + r], *sigma);
}

// ------------------------------------------------------------------------------------------------
// dense conv2d, stride 1: out = (conv(pre(x)) + bias [+ res]) * out_scale
template <int KH, int KW, int CIC>
__global__ void __launch_bounds__(256) conv2d_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                                                     const float* __restrict__ bias, const float* __restrict__ res,
                                                     float* __restrict__ out, int Cin, int H, int W, int Cout, int Ho, int Wo,
                                                     int pad, int pre_act, float slope, float out_scale) {
  constexpr int TH = 8, TW = 16, TC = 64, KK = KH * KW;
  constexpr int PH = TH + KH - 1, PW = TW + KW - 1;
  __shared__ float xs[CIC][PH][PW];
  __shared__ __align__(16) float ws[CIC][KK][TC];
  const int tid = threadIdx.x;
  const int tiles_w = (Wo + TW - 1) / TW;
  const int th0 = (blockIdx.x / tiles_w) * TH, tw0 = (blockIdx.x % tiles_w) * TW;
  const int co0 = blockIdx.y * TC;
  const int b = blockIdx.z;
  const int tx = tid & 15, cg = tid >> 4;
  float acc[TH][4];
#pragma unroll
  for (int r = 0; r < TH; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
  const float* xb = x + (long long)b * Cin * H * W;
  for (int ci0 = 0; ci0 < Cin; ci0 += CIC) {
    __syncthreads();
    for (int i = tid; i < CIC * PH * PW; i += 256) {
      const int c = i / (PH * PW);
      const int rem = i - c * (PH * PW);
      const int r = rem / PW, q = rem - r * PW;
      const int ih = th0 + r - pad, iw = tw0 + q - pad;
      float v = 0.f;
      if (ci0 + c < Cin && ih >= 0 && ih < H && iw >= 0 && iw < W) {
        v = __ldg(xb + ((long long)(ci0 + c) * H + ih) * W + iw);
        if (pre_act) v = v > 0.f ? v : v * slope;
      }
      xs[c][r][q] = v;
    }
    for (int i = tid; i < CIC * KK * TC; i += 256) {
      const int co = i % TC;
      const int rk = i / TC;  // c*KK + k
      const int c = rk / KK;
      float v = 0.f;
      if (ci0 + c < Cin && co0 + co < Cout) v = __ldg(wt + ((long long)ci0 * KK + rk) * Cout + co0 + co);
      (&ws[0][0][0])[i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CIC; ++c) {
#pragma unroll
      for (int kx = 0; kx < KW; ++kx) {
        float col[PH];
#pragma unroll
        for (int r = 0; r < PH; ++r) col[r] = xs[c][r][tx + kx];
#pragma unroll
        for (int ky = 0; ky < KH; ++ky) {
          const float4 wv = *reinterpret_cast<const float4*>(&ws[c][ky * KW + kx][cg * 4]);
#pragma unroll
          for (int r = 0; r < TH; ++r) {
            acc[r][0] = fmaf(col[r + ky], wv.x, acc[r][0]);
            acc[r][1] = fmaf(col[r + ky], wv.y, acc[r][1]);
            acc[r][2] = fmaf(col[r + ky], wv.z, acc[r][2]);
            acc[r][3] = fmaf(col[r + ky], wv.w, acc[r][3]);
          }
        }
      }
    }
  }
  const int ow = tw0 + tx;
  if (ow >= Wo) return;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int co = co0 + cg * 4 + q;
    if (co >= Cout) continue;
    const float bv = bias ? bias[co] : 0.f;
#pragma unroll
    for (int r = 0; r < TH; ++r) {
      const int oh = th0 + r;
      if (oh >= Ho) continue;
      const long long o = (((long long)b * Cout + co) * Ho + oh) * Wo + ow;
      float v = acc[r][q] + bv;
      if (res) v += res[o];
      out[o] = v * out_scale;
    }
  }
}

// depthwise 3x3, stride 2, pad 1 (LearnedDownSample 'half', models.py:37-38); w [9][C] = st2_spectral_norm_fold layout
__global__ void dwconv3x3_s2_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                    float* __restrict__ out, int C, int H, int W, int Ho, int Wo, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ow = (int)(i % Wo);
  const int oh = (int)((i / Wo) % Ho);
  const long long bc = i / ((long long)Wo * Ho);
  const int c = (int)(bc % C);
  const float* xp = x + bc * H * W;
  float acc = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int ih = 2 * oh - 1 + ky;
    if (ih < 0 || ih >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int iw = 2 * ow - 1 + kx;
      if (iw < 0 || iw >= W) continue;
      acc = fmaf(xp[(long long)ih * W + iw], w[(ky * 3 + kx) * C + c], acc);
    }
  }
  out[i] = acc + bias[c];
}

// DownSample('half') (models.py:73-78): replicate the last column when W is odd, then 2x2 mean
__global__ void avgpool_half_kernel(const float* __restrict__ x, float* __restrict__ out, int H, int W, int Ho, int Wo,
                                    long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ow = (int)(i % Wo);
  const int oh = (int)((i / Wo) % Ho);
  const long long bc = i / ((long long)Wo * Ho);
  const float* xp = x + bc * H * W;
  const int w0 = 2 * ow, w1 = min(2 * ow + 1, W - 1);
  const float* r0 = xp + (long long)(2 * oh) * W;
  const float* r1 = r0 + W;
  out[i] = (((r0[w0] + r0[w1]) + r1[w0]) + r1[w1]) * 0.25f;
}

// AdaptiveAvgPool2d(1) + LeakyReLU: [B*C, HW] -> [B*C]   (one warp per row)
__global__ void mean_hw_lrelu_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int hw, float slope) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xp = x + (long long)row * hw;
  float acc = 0.f;
  for (int i = threadIdx.x & 31; i < hw; i += 32) acc += xp[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) {
    const float m = acc / (float)hw;
    out[row] = m > 0.f ? m : m * slope;
  }
}

// ------------------------------------------------------------------------------------------------
// mel front-end: windowed frames (center=True reflect padding folded into the gather), power, log
__global__ void mel_frames_kernel(const float* __restrict__ wave, const float* __restrict__ window, int L, int F, int win, int hop,
                                  int n_fft, float* __restrict__ frames) {
  const int f = blockIdx.x, b = blockIdx.y;
  const float* wp = wave + (long long)b * L;
  float* fp = frames + ((long long)b * F + f) * win;
  const int start = f * hop - n_fft / 2 + (n_fft - win) / 2;  // window centred inside the n_fft frame
  for (int m = threadIdx.x; m < win; m += blockDim.x) {
    int i = start + m;
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    fp[m] = wp[i] * window[m];
  }
}

__global__ void mel_power_kernel(const float* __restrict__ y, int rows, int nf, float* __restrict__ p) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)rows * nf) return;
  const int k = (int)(i % nf);
  const long long r = i / nf;
  const float re = y[r * 2 * nf + k], im = y[r * 2 * nf + nf + k];
  p[i] = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
}

// out[b, m, f] = (log(eps + mel[(b*F + f), m]) - mean) / std
__global__ void logmel_kernel(const float* __restrict__ mel, int B, int F, int M, float eps, float mean, float stdv,
                              float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * F * M) return;
  const int f = (int)(i % F);
  const int m = (int)((i / F) % M);
  const int b = (int)(i / ((long long)F * M));
  const float v = mel[((long long)b * F + f) * M + m];
  out[i] = __fdiv_rn(__fsub_rn(logf(__fadd_rn(eps, v)), mean), stdv);
}

}  // namespace st2

using namespace st2;

extern "C" int st2_spectral_norm_fold(const float* weight_orig, const float* u, const float* v, int Cout, int n, float* wt,
                                      float* sigma_work, void* stream) {
  ST2_REQUIRE(weight_orig && u && v && wt && sigma_work && Cout > 0 && n > 0, "st2_spectral_norm_fold", "bad args");
  cudaStream_t st = (cudaStream_t)stream;
  sn_sigma_kernel<<<1, 1024, 0, st>>>(weight_orig, u, v, Cout, n, sigma_work);
  sn_layout_kernel<<<cdiv((long long)Cout * n, 256), 256, 0, st>>>(weight_orig, sigma_work, Cout, n, wt);
  g_launches += 2;
  ST2_CHECK_LAUNCH("st2_spectral_norm_fold");
  return 0;
}

extern "C" int st2_conv2d(const st2_conv2d_args* a, void* stream) {
  ST2_REQUIRE(a && a->x && a->wt && a->out, "st2_conv2d", "null pointer");
  ST2_REQUIRE(a->B > 0 && a->Cin > 0 && a->Cout > 0 && a->H > 0 && a->W > 0 && a->pad >= 0, "st2_conv2d", "bad shape");
  const int Ho = a->H + 2 * a->pad - a->KH + 1, Wo = a->W + 2 * a->pad - a->KW + 1;
  ST2_REQUIRE(Ho > 0 && Wo > 0, "st2_conv2d", "kernel larger than padded input");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(Ho, 8) * cdiv(Wo, 16), cdiv(a->Cout, 64), a->B);
#define ST2_LAUNCH_CONV2D(KH_, KW_, CIC_)                                                                                  \
  conv2d_kernel<KH_, KW_, CIC_><<<grid, 256, 0, st>>>(a->x, a->wt, a->bias, a->res, a->out, a->Cin, a->H, a->W, a->Cout, Ho, \
                                                      Wo, a->pad, a->pre_act, a->slope, a->out_scale)
  if (a->KH == 3 && a->KW == 3) ST2_LAUNCH_CONV2D(3, 3, 8);
  else if (a->KH == 1 && a->KW == 1) ST2_LAUNCH_CONV2D(1, 1, 16);
  else if (a->KH == 5 && a->KW == 5) ST2_LAUNCH_CONV2D(5, 5, 4);
  else ST2_REQUIRE(false, "st2_conv2d", "kernel size must be 1x1, 3x3 or 5x5 (the sizes StyleEncoder uses)");
#undef ST2_LAUNCH_CONV2D
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_conv2d");
  return 0;
}

extern "C" int st2_dwconv3x3_s2(const float* x, const float* w, const float* bias, float* out, int B, int C, int H, int W,
                                void* stream) {
  ST2_REQUIRE(x && w && bias && out && B > 0 && C > 0 && H > 0 && W > 0, "st2_dwconv3x3_s2", "bad args");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)B * C * Ho * Wo;
  dwconv3x3_s2_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, w, bias, out, C, H, W, Ho, Wo, total);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_dwconv3x3_s2");
  return 0;
}

extern "C" int st2_avgpool_half(const float* x, float* out, int BC, int H, int W, void* stream) {
  ST2_REQUIRE(x && out && BC > 0 && H >= 2 && W >= 1, "st2_avgpool_half", "bad args");
  const int Ho = H / 2, Wo = (W + 1) / 2;
  const long long total = (long long)BC * Ho * Wo;
  avgpool_half_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, out, H, W, Ho, Wo, total);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_avgpool_half");
  return 0;
}

extern "C" int st2_mean_hw_lrelu(const float* x, float* out, int rows, int hw, float slope, void* stream) {
  ST2_REQUIRE(x && out && rows > 0 && hw > 0, "st2_mean_hw_lrelu", "bad args");
  mean_hw_lrelu_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, out, rows, hw, slope);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_mean_hw_lrelu");
  return 0;
}

extern "C" int st2_mel_frames(const float* wave, const float* window, int B, int L, int win, int hop, int n_fft, float* frames,
                              void* stream) {
  ST2_REQUIRE(wave && window && frames && B > 0 && win > 0 && hop > 0 && n_fft >= win, "st2_mel_frames", "bad args");
  ST2_REQUIRE(L > n_fft / 2, "st2_mel_frames", "clip shorter than the reflect padding (n_fft/2)");
  const int F = 1 + L / hop;
  mel_frames_kernel<<<dim3(F, B), 256, 0, (cudaStream_t)stream>>>(wave, window, L, F, win, hop, n_fft, frames);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_mel_frames");
  return 0;
}

extern "C" int st2_mel_power(const float* y, int rows, int nf, float* p, void* stream) {
  ST2_REQUIRE(y && p && rows > 0 && nf > 0, "st2_mel_power", "bad args");
  mel_power_kernel<<<cdiv((long long)rows * nf, 256), 256, 0, (cudaStream_t)stream>>>(y, rows, nf, p);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_mel_power");
  return 0;
}

extern "C" int st2_logmel(const float* mel, int B, int F, int M, float eps, float mean, float stdv, float* out, void* stream) {
  ST2_REQUIRE(mel && out && B > 0 && F > 0 && M > 0 && stdv != 0.f, "st2_logmel", "bad args");
  logmel_kernel<<<cdiv((long long)B * F * M, 256), 256, 0, (cudaStream_t)stream>>>(mel, B, F, M, eps, mean, stdv, out);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_logmel");
  return 0;
}
