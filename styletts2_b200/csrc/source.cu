// Harmonic source (SineGen + SourceModuleHnNSF), STFT(20,5) features and the fused
// exp/sin + inverse STFT output stage of the iSTFTNet generator.
//
// Numerics mirror the reference's CPU arithmetic where the result is ill-conditioned:
// the instantaneous phase reaches 1e4..1e6 rad before sin(), so the low-rate phase is accumulated
// in fp64 and rounded to fp32 exactly as torch.cumsum does on CPU, the scalings are separate fp32
// multiplies, and the x300 linear interpolation uses PyTorch's align_corners=False source-index
// rule with the same FMA contractions ATen's CPU kernel compiles to (verified bit-exact against
// torch in tests/test_source_numerics.py).
#include "common.cuh"

namespace st2 {
extern long long g_launches;

constexpr int NH = 9;  // fundamental + 8 overtones (istftnet.py:313)

// Philox4x32-10 counter-based generator (Salmon et al. 2011) + Box-Muller: the throughput mode draws the
// randn_like noise of the reference (istftnet.py:242, sampler.py:509) on the device without a library call.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;  // (0,1)
  const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincosf(6.28318530717958647692f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}
// 4 standard normals for counter index i of stream (seed, offset)
// `epoch` (read from device memory, so a captured CUDA graph draws fresh noise on every replay) fills the upper counter words.
__device__ __forceinline__ float4 randn4(unsigned long long seed, unsigned long long offset, unsigned long long epoch,
                                         unsigned long long i) {
  const unsigned long long c = offset + i;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)epoch, (uint32_t)(epoch >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

__global__ void rng_advance_kernel(unsigned long long* epoch) { *epoch += 1ull; }

__global__ void randn_kernel(float* __restrict__ out, long long n, unsigned long long seed, unsigned long long offset,
                             const unsigned long long* __restrict__ epoch_p) {
  const unsigned long long epoch = epoch_p ? *epoch_p : 0ull;
  const long long n4 = (n + 3) >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = randn4(seed, offset, epoch, (unsigned long long)i);
    const long long j = i * 4;
    if (j < n) out[j] = v.x;
    if (j + 1 < n) out[j + 1] = v.y;
    if (j + 2 < n) out[j + 2] = v.z;
    if (j + 3 < n) out[j + 3] = v.w;
  }
}

__device__ __forceinline__ float torch_remainder1(float x) {
  // torch `% 1` for floats: fmod, then shift into [0,1) when negative (istftnet.py:152)
  float r = fmodf(x, 1.0f);
  if (r != 0.0f && r < 0.0f) r += 1.0f;
  return r;
}

// phase_lo[b,h,j] = fp32( fp32( fp32( fp32(sum_{i<=j} rad) * 2 ) * pi ) * scale ), rad = ((f0*(h+1))/24000) % 1
__global__ void sine_phase_kernel(const float* __restrict__ f0, int B, int F, float scale, float* __restrict__ phase) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * NH) return;
  const int b = i / NH, h = i - b * NH;
  const float harm = (float)(h + 1);
  double acc = 0.0;
  const float* fr = f0 + (long long)b * F;
  float* pr = phase + (long long)i * F;
  for (int j = 0; j < F; ++j) {
    const float fn = __fmul_rn(fr[j], harm);
    const float rad = torch_remainder1(__fdiv_rn(fn, 24000.0f));
    acc += (double)rad;
    float p = (float)acc;
    p = __fmul_rn(p, 2.0f);
    p = __fmul_rn(p, 3.14159274101257324f);  // float32(np.pi)
    p = __fmul_rn(p, scale);
    pr[j] = p;
  }
}

__global__ void sine_source_kernel(const float* __restrict__ f0, const float* __restrict__ phase, int F, int scale,
                                   const float* __restrict__ noise, const float* __restrict__ lin_w,
                                   const float* __restrict__ lin_b, float* __restrict__ out, unsigned long long seed,
                                   unsigned long long offset, const unsigned long long* __restrict__ epoch_p) {
  const int b = blockIdx.y;
  const long long L = (long long)F * scale;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= L) return;
  const float rscale = (float)(1.0 / (double)scale);
  // area_pixel_compute_source_index (align_corners=False), contracted as ATen's CPU build does
  float src = fmaf(rscale, (float)t + 0.5f, -0.5f);
  if (src < 0.f) src = 0.f;
  int i0 = (int)floorf(src);
  if (i0 > F - 1) i0 = F - 1;
  float l1 = src - (float)i0;
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  const float l0 = 1.0f - l1;
  const int i1 = i0 + (i0 < F - 1 ? 1 : 0);
  const float f0v = f0[(long long)b * F + (int)(t / scale)];  // nearest upsample (istftnet.py:352)
  const float uv = f0v > 10.0f ? 1.0f : 0.0f;
  const float noise_amp = uv > 0.f ? 0.003f : __fdiv_rn(0.1f, 3.0f);
  float nzv[12];
  if (noise) {
    const float* nz = noise + ((long long)b * L + t) * NH;
#pragma unroll
    for (int h = 0; h < NH; ++h) nzv[h] = nz[h];
  } else {  // throughput mode: draw the 9 normals of this sample in place (3 Philox calls)
    const unsigned long long e = ((unsigned long long)b * (unsigned long long)L + (unsigned long long)t) * 3ull;
    const unsigned long long epoch = epoch_p ? *epoch_p : 0ull;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float4 v = randn4(seed, offset, epoch, e + q);
      nzv[4 * q] = v.x; nzv[4 * q + 1] = v.y; nzv[4 * q + 2] = v.z; nzv[4 * q + 3] = v.w;
    }
  }
  float accv = lin_b[0];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float* pr = phase + ((long long)b * NH + h) * F;
    const float ph = fmaf(l0, pr[i0], __fmul_rn(l1, pr[i1]));
    const float sw = __fmul_rn(sinf(ph), 0.1f);
    const float v = __fadd_rn(__fmul_rn(sw, uv), __fmul_rn(noise_amp, nzv[h]));
    accv = fmaf(lin_w[h], v, accv);
  }
  out[(long long)b * L + t] = tanhf(accv);
}

// ------------------------------------------------------------------------------------------
// STFT n_fft=20 hop=5 hann(periodic) center=True reflect  ->  [ |X| (11) ; angle X (11) ]
struct Tab20 {
  float win[20];
  float cs[20];
  float sn[20];
};
__constant__ Tab20 c_tab;
static PerDevice g_tab_once;   // __constant__ memory is per device

static int ensure_tab() {
  if (!g_tab_once.first()) return 0;
  Tab20 t;
  const double PI = 3.14159265358979323846;
  for (int m = 0; m < 20; ++m) {
    t.win[m] = (float)(0.5 - 0.5 * cos(2.0 * PI * m / 20.0));
    t.cs[m] = (float)cos(2.0 * PI * m / 20.0);
    t.sn[m] = (float)sin(2.0 * PI * m / 20.0);
  }
  // exact zeros / ones where the analytic value is exact
  t.cs[5] = 0.f; t.cs[15] = 0.f; t.sn[0] = 0.f; t.sn[10] = 0.f;
  cudaError_t e = cudaMemcpyToSymbol(c_tab, &t, sizeof(t));
  if (e != cudaSuccess) { set_error("stft table", e); g_tab_once.done[g_tab_once.dev()] = false; return (int)e; }
  return 0;
}

__global__ void stft20_kernel(const float* __restrict__ x, int L, int Fr, float* __restrict__ har) {
  const int b = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= Fr) return;
  const float* xr = x + (long long)b * L;
  float xw[20];
#pragma unroll
  for (int m = 0; m < 20; ++m) {
    int i = 5 * f + m - 10;
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    xw[m] = xr[i] * c_tab.win[m];
  }
  float* hb = har + (long long)b * 22 * Fr + f;
#pragma unroll
  for (int k = 0; k <= 10; ++k) {
    float re = 0.f, im = 0.f;
#pragma unroll
    for (int m = 0; m < 20; ++m) {
      const int idx = (k * m) % 20;
      re = fmaf(xw[m], c_tab.cs[idx], re);
      im = fmaf(-xw[m], c_tab.sn[idx], im);
    }
    if (k == 0 || k == 10) im = 0.0f;  // rfft: DC / Nyquist are exactly real (+0 imaginary part)
    hb[(long long)k * Fr] = hypotf(re, im);
    hb[(long long)(11 + k) * Fr] = atan2f(im, re);
  }
}

// ------------------------------------------------------------------------------------------
// conv_post tail + inverse STFT: spec = exp(x[:11]), phase = sin(x[11:]), X = spec*e^{i phase},
// irfft(n=20, backward norm), * window, overlap-add, / window envelope, trim n_fft/2 each side.
constexpr int ISTFT_FB = 64;  // frames of output per CTA -> 320 samples

__global__ void __launch_bounds__(320) istft20_kernel(const float* __restrict__ x, int Fr, float* __restrict__ wav) {
  __shared__ float re_s[ISTFT_FB + 4][11];
  __shared__ float im_s[ISTFT_FB + 4][11];
  const int b = blockIdx.y;
  const int fb0 = blockIdx.x * ISTFT_FB;  // this CTA produces samples n' in [5*fb0, 5*fb0 + 320)
  const int Lout = 5 * (Fr - 1);
  // frames that can touch those samples: n = n'+10, f in [ceil((n-19)/5), floor(n/5)] -> [fb0-1, fb0+ISTFT_FB+1]
  const int fbase = fb0 - 1;
  const float* xb = x + (long long)b * 22 * Fr;
  // the twiddle / window tables are indexed by (k*m) % 20 with m different in every lane: constant memory would serialise
  // those reads, shared memory serves them in one pass
  __shared__ float cs_s[20], sn_s[20], win_s[20];
  if (threadIdx.x < 20) { cs_s[threadIdx.x] = c_tab.cs[threadIdx.x]; sn_s[threadIdx.x] = c_tab.sn[threadIdx.x]; win_s[threadIdx.x] = c_tab.win[threadIdx.x]; }
  // bin-major walk: consecutive threads read consecutive frames of ONE bin row (coalesced)
  for (int i = threadIdx.x; i < (ISTFT_FB + 4) * 11; i += blockDim.x) {
    const int k = i / (ISTFT_FB + 4), fl = i - k * (ISTFT_FB + 4);
    const int f = fbase + fl;
    float re = 0.f, im = 0.f;
    if (f >= 0 && f < Fr) {
      const float mag = expf(xb[(long long)k * Fr + f]);
      const float ph = sinf(xb[(long long)(11 + k) * Fr + f]);
      float s, c;
      sincosf(ph, &s, &c);
      re = mag * c;
      im = mag * s;
    }
    re_s[fl][k] = re;
    im_s[fl][k] = im;
  }
  __syncthreads();
  const int np = 5 * fb0 + threadIdx.x;  // output sample index
  if (np >= Lout) return;
  const int n = np + 10;
  int f_hi = n / 5;
  if (f_hi > Fr - 1) f_hi = Fr - 1;
  int f_lo = (n - 19 + 4) / 5;  // ceil((n-19)/5), n >= 10 so numerator may be negative only slightly
  if (n - 19 < 0) f_lo = 0;
  if (f_lo < 0) f_lo = 0;
  float acc = 0.f, env = 0.f;
  for (int f = f_lo; f <= f_hi; ++f) {
    const int m = n - 5 * f;  // 0..19
    const int fl = f - fbase;
    float v = re_s[fl][0] + ((m & 1) ? -re_s[fl][10] : re_s[fl][10]);
#pragma unroll
    for (int k = 1; k <= 9; ++k) {
      const int idx = (k * m) % 20;
      v = fmaf(2.0f * re_s[fl][k], cs_s[idx], v);
      v = fmaf(-2.0f * im_s[fl][k], sn_s[idx], v);
    }
    const float w = win_s[m];
    acc = fmaf(v * 0.05f, w, acc);
    env = fmaf(w, w, env);
  }
  wav[(long long)b * Lout + np] = acc / env;
}

}  // namespace st2

using namespace st2;

// fp32 waveform -> 16-bit PCM (row f4: the wire format after the path): round-half-even of x * 32767 * gain, saturated.
// Four samples per thread (128-bit load, 64-bit store) when the row is aligned.
__global__ void pcm16_kernel(const float* __restrict__ x, long long n, float gain, short* __restrict__ out) {
  const long long i4 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  auto q = [gain](float v) -> short {
    float r = rintf(v * gain);
    r = fminf(fmaxf(r, -32768.f), 32767.f);
    return (short)(int)r;
  };
  if (i4 + 3 < n && ((reinterpret_cast<size_t>(x) & 15) == 0) && ((reinterpret_cast<size_t>(out) & 7) == 0)) {
    const float4 v = *reinterpret_cast<const float4*>(x + i4);
    short4 o;
    o.x = q(v.x); o.y = q(v.y); o.z = q(v.z); o.w = q(v.w);
    *reinterpret_cast<short4*>(out + i4) = o;
  } else {
    for (long long i = i4; i < n && i < i4 + 4; ++i) out[i] = q(x[i]);
  }
}

extern "C" {

int st2_pcm16(const float* wav, long long n, float gain, short* out, void* stream) {
  ST2_REQUIRE(wav && out && n > 0, "st2_pcm16", "bad args");
  const long long threads = (n + 3) / 4;
  pcm16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wav, n, 32767.0f * gain, out);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_pcm16");
  return 0;
}

int st2_rng_advance(unsigned long long* epoch, void* stream) {
  ST2_REQUIRE(epoch, "st2_rng_advance", "bad args");
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(epoch);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_rng_advance");
  return 0;
}

int st2_randn(float* out, long long n, unsigned long long seed, unsigned long long offset, const unsigned long long* epoch,
              void* stream) {
  ST2_REQUIRE(out && n > 0, "st2_randn", "bad args");
  const long long n4 = (n + 3) / 4;
  const int grid = (int)((n4 + 255) / 256 > 4096 ? 4096 : (n4 + 255) / 256);
  randn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset, epoch);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_randn");
  return 0;
}

int st2_sine_source(const float* f0, int B, int F, int scale, const float* noise, const float* lin_w, const float* lin_b,
                    float* out, float* phase_work, unsigned long long seed, unsigned long long offset,
                    const unsigned long long* epoch, void* stream) {
  ST2_REQUIRE(f0 && lin_w && lin_b && out && phase_work && B > 0 && F > 0 && scale > 0, "st2_sine_source", "bad args");
  cudaStream_t st = (cudaStream_t)stream;
  sine_phase_kernel<<<cdiv(B * NH, 64), 64, 0, st>>>(f0, B, F, (float)scale, phase_work);
  ++g_launches;
  const long long L = (long long)F * scale;
  sine_source_kernel<<<dim3(cdiv(L, 256), B), 256, 0, st>>>(f0, phase_work, F, scale, noise, lin_w, lin_b, out, seed, offset, epoch);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_sine_source");
  return 0;
}

int st2_stft20(const float* x, int B, int L, float* har, void* stream) {
  ST2_REQUIRE(x && har && B > 0 && L >= 20 && L % 5 == 0, "st2_stft20", "bad args");
  if (int rc = ensure_tab()) return rc;
  const int Fr = L / 5 + 1;
  stft20_kernel<<<dim3(cdiv(Fr, 128), B), 128, 0, (cudaStream_t)stream>>>(x, L, Fr, har);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_stft20");
  return 0;
}

int st2_istft20_expsin(const float* x, int B, int Fr, float* wav, void* stream) {
  ST2_REQUIRE(x && wav && B > 0 && Fr >= 2, "st2_istft20_expsin", "bad args");
  if (int rc = ensure_tab()) return rc;
  const int Lout = 5 * (Fr - 1);
  istft20_kernel<<<dim3(cdiv(Lout, 5 * ISTFT_FB), B), 320, 0, (cudaStream_t)stream>>>(x, Fr, wav);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_istft20_expsin");
  return 0;
}

}  // extern "C"
