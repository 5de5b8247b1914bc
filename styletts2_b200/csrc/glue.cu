// Small glue kernels of the inference path: token embedding, duration rounding (the integer
// boundary), alignment expansion as a gather, ADPM2/KDiffusion elementwise updates, error state.
#include "common.cuh"
#include <string.h>

namespace st2 {
extern long long g_launches;

static thread_local char g_err[512] = "";
void set_error(const char* where, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
}
void set_error_msg(const char* where, const char* msg) { snprintf(g_err, sizeof(g_err), "%s: %s", where, msg); }

__global__ void embedding_cl_kernel(const long long* __restrict__ tokens, const float* __restrict__ table,
                                    const int* __restrict__ lengths, int N, int C, float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const bool masked = lengths && n >= lengths[b];
  const long long tok = tokens[(long long)b * N + n];
  out[((long long)b * C + c) * N + n] = masked ? 0.f : table[tok * C + c];
}

__global__ void embedding_sum_rows_kernel(const long long* __restrict__ tokens, const float* __restrict__ word,
                                          const float* __restrict__ pos, const float* __restrict__ type0, int N, int E,
                                          float* __restrict__ out) {
  const int row = blockIdx.x;  // (b,n)
  const int n = row % N;
  const long long tok = tokens[row];
  for (int e = threadIdx.x; e < E; e += blockDim.x)
    out[(long long)row * E + e] = __fadd_rn(__fadd_rn(word[tok * E + e], type0[e]), pos[(long long)n * E + e]);
}

__global__ void durations_kernel(const float* __restrict__ logits, int rows, int N, int J, int last_plus,
                                 const int* __restrict__ lengths, int* __restrict__ pred, float* __restrict__ dur_f) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int len = lengths ? lengths[r / N] : N;   // padded tokens (n >= len) emit no frames
  if ((r % N) >= len) {
    pred[r] = 0;
    if (dur_f) dur_f[r] = 0.f;
    return;
  }
  const float* lp = logits + (long long)r * J;
  double s = 0.0;
  for (int j = 0; j < J; ++j) s += (double)sigmoidf_(lp[j]);
  const float d = (float)s;
  if (dur_f) dur_f[r] = d;
  float rd = rintf(d);  // torch.round: half to even
  if (rd < 1.f) rd = 1.f;
  int v = (int)rd;
  if ((r % N) == len - 1) v += last_plus;   // pred_dur[-1] += 5 of the single-speaker notebook: the LAST REAL token
  pred[r] = v;
}

// one CTA per utterance: exclusive scan of durations (serial, N <= 4096), then frames are filled by
// binary search.  tok[b,t] = index of the token covering frame t (after the optional one-frame delay).
__global__ void frame_tokens_kernel(const int* __restrict__ dur, int N, int T, int shift_right, int* __restrict__ tok,
                                    int* __restrict__ total) {
  extern __shared__ int cum[];  // [N+1]
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < N; ++i) {
      cum[i] = s;
      s += dur[(long long)b * N + i];
    }
    cum[N] = s;
    if (total) total[b] = s;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int tt = shift_right ? (t > 0 ? t - 1 : 0) : t;
    int lo = 0, hi = N - 1;  // largest i with cum[i] <= tt
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (cum[mid] <= tt) lo = mid; else hi = mid - 1;
    }
    tok[(long long)b * T + t] = lo;
  }
}

__global__ void expand_rows_kernel(const float* __restrict__ src, long long src_ld, const int* __restrict__ tok, int N, int T,
                                   int C, float* __restrict__ out, long long out_ld) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int n = tok[(long long)b * T + t];
  const float* sr = src + ((long long)b * N + n) * src_ld;
  float* dr = out + ((long long)b * T + t) * out_ld;
  for (int c = threadIdx.x; c < C; c += blockDim.x) dr[c] = sr[c];
}

// Strided Conv1d (stride S, K = J*S taps) as a stride-1 conv over the polyphase view of its input:
// xp[b, c*S + r, q] = x[b, c, q*S + r - pad] (zero outside), q < Lp.  A block reads S*128 CONSECUTIVE samples of one input
// row (coalesced), redistributes them through shared memory and writes S rows of 128 consecutive outputs.
__global__ void __launch_bounds__(128) polyphase_gather_kernel(const float* __restrict__ x, long long x_bstride, int C, int L, int S, int pad,
                                                               int Lp, float* __restrict__ xp) {
  extern __shared__ float seg[];   // S * 128
  const int b = blockIdx.z, c = blockIdx.y, q0 = blockIdx.x * 128;
  const float* xr = x + (long long)b * x_bstride + (long long)c * L;
  const int base = q0 * S - pad;
  for (int i = threadIdx.x; i < S * 128; i += 128) {
    const int t = base + i;
    seg[i] = (t >= 0 && t < L) ? xr[t] : 0.f;
  }
  __syncthreads();
  const int q = q0 + threadIdx.x;
  if (q < Lp) {
    float* o = xp + ((long long)b * C * S + (long long)c * S) * Lp + q;
    for (int r = 0; r < S; ++r) o[(long long)r * Lp] = seg[threadIdx.x * S + r];
  }
}

__global__ void expand_cl_kernel(const float* __restrict__ src, const int* __restrict__ tok, int C, int N, int T,
                                 float* __restrict__ out, long long out_bstride) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  out[(long long)b * out_bstride + (long long)c * T + t] = src[((long long)b * C + c) * N + tok[(long long)b * T + t]];
}

// KDiffusion.denoise_fn tail + one ADPM2 half step (sampler.py:204-208,499-510); separate roundings
// (no FMA contraction) to stay close to the reference's op-by-op fp32 arithmetic.
__global__ void kdiff_step_kernel(const float* __restrict__ x_eval, const float* __restrict__ x_pred,
                                  const float* __restrict__ x_pred_masked, float cfg_scale, float c_skip, float c_out,
                                  float sigma_eval, const float* __restrict__ x_base, float dt, const float* __restrict__ eps,
                                  float sigma_up, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float xp = x_pred[i];
  if (x_pred_masked) {
    const float m = x_pred_masked[i];
    xp = __fadd_rn(m, __fmul_rn(__fsub_rn(xp, m), cfg_scale));
  }
  const float xe = x_eval[i];
  const float den = __fadd_rn(__fmul_rn(c_skip, xe), __fmul_rn(c_out, xp));
  const float d = __fdiv_rn(__fsub_rn(xe, den), sigma_eval);
  float o = __fadd_rn(x_base[i], __fmul_rn(d, dt));
  if (eps) o = __fadd_rn(o, __fmul_rn(eps[i], sigma_up));
  out[i] = o;
}

__global__ void scale_kernel(const float* __restrict__ x, float a, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fmul_rn(a, x[i]);
}

__global__ void axpby_kernel(const float* __restrict__ x, float a, const float* __restrict__ y, float b, float* __restrict__ out,
                             int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fadd_rn(__fmul_rn(a, x[i]), __fmul_rn(b, y[i]));
}

__global__ void time_embedding_kernel(const float* __restrict__ tp, const float* __restrict__ w, int half, int B, float* __restrict__ out,
                                      long long ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = 2 * half + 1;
  if (i >= B * W) return;
  const int b = i / W, c = i - b * W;
  const float t = tp[b];
  float v;
  if (c == 0) v = t;
  else {
    const int j = c <= half ? c - 1 : c - 1 - half;
    // freqs = ((t * w) * 2) * pi  (modules.py:667)
    const float fr = __fmul_rn(__fmul_rn(__fmul_rn(t, w[j]), 2.0f), 3.14159274101257324f);
    v = c <= half ? sinf(fr) : cosf(fr);
  }
  out[(long long)b * ld + c] = v;
}

}  // namespace st2

using namespace st2;

extern "C" {

const char* st2_last_error(void) { return g_err; }
int st2_abi_version(void) { return ST2_ABI_VERSION; }
long long st2_launch_count(void) { return g_launches; }

int st2_embedding_cl(const long long* tokens, const float* table, const int* lengths, int B, int N, int C, float* out,
                     void* stream) {
  ST2_REQUIRE(tokens && table && out && B > 0 && N > 0 && C > 0, "st2_embedding_cl", "bad args");
  embedding_cl_kernel<<<dim3(cdiv(N, 128), C, B), 128, 0, (cudaStream_t)stream>>>(tokens, table, lengths, N, C, out);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_embedding_cl");
  return 0;
}

int st2_embedding_sum_rows(const long long* tokens, const float* word, const float* pos, const float* type0, int B, int N, int E,
                           float* out, void* stream) {
  ST2_REQUIRE(tokens && word && pos && type0 && out && B > 0 && N > 0 && E > 0, "st2_embedding_sum_rows", "bad args");
  embedding_sum_rows_kernel<<<B * N, 128, 0, (cudaStream_t)stream>>>(tokens, word, pos, type0, N, E, out);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_embedding_sum_rows");
  return 0;
}

int st2_durations(const float* logits, int B, int N, int J, int last_plus, const int* lengths, int* pred_dur, float* dur_f,
                  void* stream) {
  ST2_REQUIRE(logits && pred_dur && B > 0 && N > 0 && J > 0, "st2_durations", "bad args");
  durations_kernel<<<cdiv(B * N, 128), 128, 0, (cudaStream_t)stream>>>(logits, B * N, N, J, last_plus, lengths, pred_dur, dur_f);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_durations");
  return 0;
}

int st2_frame_tokens(const int* dur, int B, int N, int T, int shift_right, int* tok, int* total, void* stream) {
  ST2_REQUIRE(dur && tok && B > 0 && N > 0 && T > 0 && N <= 8192, "st2_frame_tokens", "bad args");
  frame_tokens_kernel<<<B, 256, (N + 1) * sizeof(int), (cudaStream_t)stream>>>(dur, N, T, shift_right, tok, total);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_frame_tokens");
  return 0;
}

int st2_expand_rows(const float* src, long long src_ld, const int* tok, int B, int N, int T, int C, float* out,
                    long long out_ld, void* stream) {
  ST2_REQUIRE(src && tok && out && B > 0 && N > 0 && T > 0 && C > 0, "st2_expand_rows", "bad args");
  expand_rows_kernel<<<dim3(T, B), 128, 0, (cudaStream_t)stream>>>(src, src_ld, tok, N, T, C, out, out_ld);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_expand_rows");
  return 0;
}

int st2_polyphase_gather(const float* x, long long x_bstride, int B, int C, int L, int S, int pad, int Lp, float* xp, void* stream) {
  ST2_REQUIRE(x && xp && B > 0 && C > 0 && L > 0 && S > 0 && S <= 64 && Lp > 0, "st2_polyphase_gather", "bad args");
  polyphase_gather_kernel<<<dim3(cdiv(Lp, 128), C, B), 128, S * 128 * sizeof(float), (cudaStream_t)stream>>>(x, x_bstride, C, L, S, pad, Lp, xp);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_polyphase_gather");
  return 0;
}

int st2_expand_cl(const float* src, const int* tok, int B, int C, int N, int T, float* out, long long out_bstride,
                  void* stream) {
  ST2_REQUIRE(src && tok && out && B > 0 && N > 0 && T > 0 && C > 0, "st2_expand_cl", "bad args");
  expand_cl_kernel<<<dim3(cdiv(T, 128), C, B), 128, 0, (cudaStream_t)stream>>>(src, tok, C, N, T, out, out_bstride);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_expand_cl");
  return 0;
}

int st2_kdiff_step(const float* x_eval, const float* x_pred, const float* x_pred_masked, float cfg_scale, float c_skip,
                   float c_out, float sigma_eval, const float* x_base, float dt, const float* eps, float sigma_up,
                   float* out, int n, void* stream) {
  ST2_REQUIRE(x_eval && x_pred && x_base && out && n > 0, "st2_kdiff_step", "bad args");
  kdiff_step_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x_eval, x_pred, x_pred_masked, cfg_scale, c_skip, c_out,
                                                                    sigma_eval, x_base, dt, eps, sigma_up, out, n);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_kdiff_step");
  return 0;
}

int st2_scale(const float* x, float a, float* out, int n, void* stream) {
  ST2_REQUIRE(x && out && n > 0, "st2_scale", "bad args");
  scale_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, a, out, n);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_scale");
  return 0;
}

int st2_axpby(const float* x, float a, const float* y, float b, float* out, int n, void* stream) {
  ST2_REQUIRE(x && y && out && n > 0, "st2_axpby", "bad args");
  axpby_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, a, y, b, out, n);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_axpby");
  return 0;
}

int st2_time_embedding(const float* t, const float* w, int half, int B, float* out, long long ld, void* stream) {
  ST2_REQUIRE(t && w && out && half > 0 && B > 0, "st2_time_embedding", "bad args");
  time_embedding_kernel<<<cdiv(B * (2 * half + 1), 128), 128, 0, (cudaStream_t)stream>>>(t, w, half, B, out, ld);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_time_embedding");
  return 0;
}

}  // extern "C"
