// Bidirectional LSTM recurrence (fp32).  The input projection x W_ih^T + b is one batched GEMM
// (st2_linear) done beforehand; this file runs the serial part: for every step,
//   gates = gx[:, t] + h_{t-1} W_hh^T ; (i,f,g,o) ; c = f*c + i*g ; h = o*tanh(c)
// v1: one launch per time step covering both directions (grid.y) and all utterances; a CTA owns
// 4 hidden units x 32 utterances, stages its 16 W_hh rows and the 32 previous hidden vectors in
// shared memory and does 128-bit conflict-free reads.  The sequential depth (N token steps,
// T frame steps) makes this latency-bound, not bandwidth-bound.
#include <cooperative_groups.h>

#include "common.cuh"

namespace st2 {
extern long long g_launches;

constexpr int LSTM_UT = 4;   // hidden units per CTA
constexpr int LSTM_BT = 32;  // utterances per CTA

__global__ void __launch_bounds__(LSTM_UT* LSTM_BT) lstm_step_kernel(
    const float* __restrict__ gx, const float* __restrict__ whh, float* __restrict__ out, long long o_bs, long long o_ts,
    long long o_cs, const int* __restrict__ lengths, int B, int L, int H, int step, const float* __restrict__ h_prev,
    float* __restrict__ h_next, float* __restrict__ c_state) {
  extern __shared__ __align__(16) float sm[];
  const int HP = H + 4;
  float* ws = sm;                       // [4 gates][UT][HP]
  float* hs = sm + 4 * LSTM_UT * HP;    // [BT][HP]
  const int dir = blockIdx.y;
  const int j0 = blockIdx.x * LSTM_UT;
  const int b0 = blockIdx.z * LSTM_BT;
  const int tid = threadIdx.x;
  const float* wd = whh + (long long)dir * 4 * H * H;
  // vectorised, unrolled staging: every thread has several independent 128-bit loads in flight
  // (a scalar loop with div/mod serialised ~100 dependent global loads per thread: 50 us per step)
  const int H4 = H >> 2;
  {
    const int n4 = 4 * LSTM_UT * H4;  // float4 count of the W slab
#pragma unroll 4
    for (int i = tid; i < n4; i += LSTM_UT * LSTM_BT) {
      const int k4 = i % H4;
      const int r = i / H4;  // g*UT + u
      const int g = r / LSTM_UT, u = r - g * LSTM_UT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + u < H) v = __ldg(reinterpret_cast<const float4*>(wd + ((long long)g * H + j0 + u) * H) + k4);
      *reinterpret_cast<float4*>(ws + r * HP + 4 * k4) = v;
    }
  }
  const float* hp = h_prev + (long long)dir * B * H;
  {
    const int n4 = LSTM_BT * H4;
#pragma unroll 8
    for (int i = tid; i < n4; i += LSTM_UT * LSTM_BT) {
      const int k4 = i % H4, bl = i / H4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + bl < B) v = *(reinterpret_cast<const float4*>(hp + (long long)(b0 + bl) * H) + k4);
      *reinterpret_cast<float4*>(hs + bl * HP + 4 * k4) = v;
    }
  }
  __syncthreads();
  const int u = tid % LSTM_UT, bl = tid / LSTM_UT;
  const int b = b0 + bl, j = j0 + u;
  if (b >= B || j >= H) return;
  const int len = lengths ? lengths[b] : L;
  float* hn = h_next + (long long)dir * B * H + (long long)b * H + j;
  if (step >= len) {  // padded step: state carried through unchanged, no output
    *hn = hs[bl * HP + j];
    return;
  }
  const int t = dir == 0 ? step : (len - 1 - step);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const float* hr = hs + bl * HP;
  const float* w0 = ws + (0 * LSTM_UT + u) * HP;
  const float* w1 = ws + (1 * LSTM_UT + u) * HP;
  const float* w2 = ws + (2 * LSTM_UT + u) * HP;
  const float* w3 = ws + (3 * LSTM_UT + u) * HP;
#pragma unroll 4
  for (int k = 0; k < H; k += 4) {
    const float4 hv = *reinterpret_cast<const float4*>(hr + k);
    const float4 x0 = *reinterpret_cast<const float4*>(w0 + k);
    const float4 x1 = *reinterpret_cast<const float4*>(w1 + k);
    const float4 x2 = *reinterpret_cast<const float4*>(w2 + k);
    const float4 x3 = *reinterpret_cast<const float4*>(w3 + k);
    a0 = fmaf(hv.x, x0.x, a0); a0 = fmaf(hv.y, x0.y, a0); a0 = fmaf(hv.z, x0.z, a0); a0 = fmaf(hv.w, x0.w, a0);
    a1 = fmaf(hv.x, x1.x, a1); a1 = fmaf(hv.y, x1.y, a1); a1 = fmaf(hv.z, x1.z, a1); a1 = fmaf(hv.w, x1.w, a1);
    a2 = fmaf(hv.x, x2.x, a2); a2 = fmaf(hv.y, x2.y, a2); a2 = fmaf(hv.z, x2.z, a2); a2 = fmaf(hv.w, x2.w, a2);
    a3 = fmaf(hv.x, x3.x, a3); a3 = fmaf(hv.y, x3.y, a3); a3 = fmaf(hv.z, x3.z, a3); a3 = fmaf(hv.w, x3.w, a3);
  }
  const float* g = gx + ((long long)b * L + t) * (8 * H) + (long long)dir * 4 * H + j;
  const float gi = sigmoidf_(g[0] + a0);
  const float gf = sigmoidf_(g[H] + a1);
  const float gg = tanhf(g[2 * H] + a2);
  const float go = sigmoidf_(g[3 * H] + a3);
  float* cp = c_state + (long long)dir * B * H + (long long)b * H + j;
  const float c = gf * (*cp) + gi * gg;
  *cp = c;
  const float h = go * tanhf(c);
  *hn = h;
  out[(long long)b * o_bs + (long long)t * o_ts + (long long)(dir * H + j) * o_cs] = h;
}

// Persistent variant: ONE cooperative launch for the whole sequence.  Each CTA keeps its 16 W_hh rows in shared
// memory for all steps (no per-step re-staging, no per-step launch); the previous hidden state is exchanged through
// L2 with one grid-wide barrier per step (cooperative groups grid.sync()).  grid = (H/4, 2 directions) <= 148 CTAs.
__global__ void __launch_bounds__(LSTM_UT* LSTM_BT) lstm_persistent_kernel(
    const float* __restrict__ gx, const float* __restrict__ whh, float* __restrict__ out, long long o_bs, long long o_ts,
    long long o_cs, const int* __restrict__ lengths, int B, int L, int H, float* __restrict__ h0, float* __restrict__ h1,
    float* __restrict__ c_state, unsigned int* __restrict__ step_bar) {
  extern __shared__ __align__(16) float sm[];
  const int HP = H + 4;
  float* ws = sm;
  float* hs = sm + 4 * LSTM_UT * HP;
  const int dir = blockIdx.y;
  const int j0 = blockIdx.x * LSTM_UT;
  const int tid = threadIdx.x;
  const int H4 = H >> 2;
  const float* wd = whh + (long long)dir * 4 * H * H;
  {
    const int n4 = 4 * LSTM_UT * H4;
#pragma unroll 4
    for (int i = tid; i < n4; i += LSTM_UT * LSTM_BT) {
      const int k4 = i % H4;
      const int r = i / H4;
      const int g = r / LSTM_UT, u = r - g * LSTM_UT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + u < H) v = __ldg(reinterpret_cast<const float4*>(wd + ((long long)g * H + j0 + u) * H) + k4);
      *reinterpret_cast<float4*>(ws + r * HP + 4 * k4) = v;
    }
  }
  const int u = tid % LSTM_UT, bl = tid / LSTM_UT;
  const int j = j0 + u;
  const float* w0 = ws + (0 * LSTM_UT + u) * HP;
  const float* w1 = ws + (1 * LSTM_UT + u) * HP;
  const float* w2 = ws + (2 * LSTM_UT + u) * HP;
  const float* w3 = ws + (3 * LSTM_UT + u) * HP;
  const int nbt = (B + LSTM_BT - 1) / LSTM_BT;
  const unsigned int nct = gridDim.x;                 // CTAs of this direction
  unsigned int* bar = step_bar + dir * 32;            // one 128-byte line per direction
  float creg = 0.f;                                   // cell state lives in a register when one batch tile covers B
  for (int step = 0; step < L; ++step) {
    const float* hprev = ((step & 1) ? h1 : h0) + (long long)dir * B * H;
    float* hnext = ((step & 1) ? h0 : h1) + (long long)dir * B * H;
    for (int bt = 0; bt < nbt; ++bt) {
      const int b0 = bt * LSTM_BT;
      const int b = b0 + bl;
      const bool live = b < B && j < H;
      // everything that does not depend on h_{t-1} is fetched first, under the latency of the h staging
      int len = L, t = 0;
      float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, cprev = creg;
      float* cp = c_state + (long long)dir * B * H + (long long)b * H + j;
      if (live) {
        len = lengths ? lengths[b] : L;
        if (step < len) {
          t = dir == 0 ? step : (len - 1 - step);
          const float* g = gx + ((long long)b * L + t) * (8 * H) + (long long)dir * 4 * H + j;
          g0 = __ldg(g); g1 = __ldg(g + H); g2 = __ldg(g + 2 * H); g3 = __ldg(g + 3 * H);
          if (nbt > 1) cprev = *cp;
        }
      }
      __syncthreads();  // previous users of hs are done (also orders the one-time ws staging)
      {
        const int n4 = LSTM_BT * H4;
#pragma unroll 8
        for (int i = tid; i < n4; i += LSTM_UT * LSTM_BT) {
          const int k4 = i % H4, bb = i / H4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          // L2 (cache-global) loads: written by other CTAs in the previous step, must not hit a stale L1 line
          if (b0 + bb < B) v = __ldcg(reinterpret_cast<const float4*>(hprev + (long long)(b0 + bb) * H) + k4);
          *reinterpret_cast<float4*>(hs + bb * HP + 4 * k4) = v;
        }
      }
      __syncthreads();
      if (live) {
        float* hn = hnext + (long long)b * H + j;
        if (step >= len) {
          *hn = hs[bl * HP + j];
        } else {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          const float* hr = hs + bl * HP;
#pragma unroll 4
          for (int k = 0; k < H; k += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(hr + k);
            const float4 x0 = *reinterpret_cast<const float4*>(w0 + k);
            const float4 x1 = *reinterpret_cast<const float4*>(w1 + k);
            const float4 x2 = *reinterpret_cast<const float4*>(w2 + k);
            const float4 x3 = *reinterpret_cast<const float4*>(w3 + k);
            a0 = fmaf(hv.x, x0.x, a0); a0 = fmaf(hv.y, x0.y, a0); a0 = fmaf(hv.z, x0.z, a0); a0 = fmaf(hv.w, x0.w, a0);
            a1 = fmaf(hv.x, x1.x, a1); a1 = fmaf(hv.y, x1.y, a1); a1 = fmaf(hv.z, x1.z, a1); a1 = fmaf(hv.w, x1.w, a1);
            a2 = fmaf(hv.x, x2.x, a2); a2 = fmaf(hv.y, x2.y, a2); a2 = fmaf(hv.z, x2.z, a2); a2 = fmaf(hv.w, x2.w, a2);
            a3 = fmaf(hv.x, x3.x, a3); a3 = fmaf(hv.y, x3.y, a3); a3 = fmaf(hv.z, x3.z, a3); a3 = fmaf(hv.w, x3.w, a3);
          }
          const float gi = sigmoidf_(g0 + a0);
          const float gf = sigmoidf_(g1 + a1);
          const float gg = tanhf(g2 + a2);
          const float go = sigmoidf_(g3 + a3);
          const float c = gf * cprev + gi * gg;
          if (nbt > 1) *cp = c; else creg = c;
          const float h = go * tanhf(c);
          *hn = h;
          out[(long long)b * o_bs + (long long)t * o_ts + (long long)(dir * H + j) * o_cs] = h;
        }
      }
    }
    // per-direction step barrier (the two directions never exchange data): release our h writes, then wait
    // until all `nct` CTAs of this direction have arrived `step+1` times.  Co-residency of the CTAs is
    // guaranteed by the cooperative launch.
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(bar, 1u);
      const unsigned int want = nct * (unsigned int)(step + 1);
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      } while (seen < want);
    }
    __syncthreads();
  }
}

}  // namespace st2

using namespace st2;

extern "C" int st2_lstm_bidir(const float* gx, const float* whh, float* out, long long o_bs, long long o_ts, long long o_cs,
                              const int* lengths, int B, int L, int H, float* work, void* stream) {
  ST2_REQUIRE(gx && whh && out && work && B > 0 && L > 0 && H > 0 && H % 4 == 0, "st2_lstm_bidir", "bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)2 * B * H;
  float* h0 = work;
  float* h1 = work + n;
  float* c = work + 2 * n;
  unsigned int* step_bar = reinterpret_cast<unsigned int*>(work + 3 * n);   // 64 words: per-direction step barriers
  cudaError_t e = cudaMemsetAsync(work, 0, (3 * n + 64) * sizeof(float), st);
  if (e != cudaSuccess) { set_error("st2_lstm_bidir", e); return (int)e; }
  const size_t smem = (size_t)(4 * LSTM_UT + LSTM_BT) * (H + 4) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(lstm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_done = true;
  }
  ST2_REQUIRE(smem <= 160 * 1024, "st2_lstm_bidir", "hidden size too large");
  // persistent cooperative kernel when the grid fits on the device (H/4 * 2 CTAs), else one launch per step
  static int coop = -1, num_sms = 0;
  if (coop < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(lstm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  }
  if (coop && cdiv(H, LSTM_UT) * 2 <= num_sms) {
    dim3 pgrid(cdiv(H, LSTM_UT), 2);
    void* args[] = {(void*)&gx, (void*)&whh, (void*)&out, (void*)&o_bs, (void*)&o_ts, (void*)&o_cs, (void*)&lengths,
                    (void*)&B,  (void*)&L,   (void*)&H,   (void*)&h0,   (void*)&h1,   (void*)&c,
                    (void*)&step_bar};
    e = cudaLaunchCooperativeKernel((const void*)lstm_persistent_kernel, pgrid, dim3(LSTM_UT * LSTM_BT), args, smem, st);
    if (e != cudaSuccess) { set_error("st2_lstm_bidir (cooperative launch)", e); return (int)e; }
    ++g_launches;
    return 0;
  }
  dim3 grid(cdiv(H, LSTM_UT), 2, cdiv(B, LSTM_BT));
  for (int s = 0; s < L; ++s) {
    const float* hp = (s & 1) ? h1 : h0;
    float* hn = (s & 1) ? h0 : h1;
    lstm_step_kernel<<<grid, LSTM_UT * LSTM_BT, smem, st>>>(gx, whh, out, o_bs, o_ts, o_cs, lengths, B, L, H, s, hp, hn, c);
    ++g_launches;
  }
  ST2_CHECK_LAUNCH("st2_lstm_bidir");
  return 0;
}
