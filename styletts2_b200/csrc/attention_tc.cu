// Tensor-core (tcgen05 / TMEM) multi-head attention for the style denoiser and PL-BERT -- sm_100a.
//
//   out[b, n, h, :] = softmax_m( scale * q[b,n,h,:] . k[b,m,h,:] ) v[b,m,h,:]        head dimension 64, N <= 4096 keys
//   (Modules/diffusion/modules.py:523-535; transformers.AlbertModel's attention for PL-BERT, with a key-padding mask)
//
// The predicted integer durations are downstream, so both contractions run at fp32 accuracy on the 16-bit tensor
// cores with the recipe of linear_tc.cu: every operand is split into two fp16 planes, x = h + l * 2^-11 with h = fp16(x),
// l = fp16((x - h) * 2^11); a product is h*h (accumulator MAIN) + h*l + l*h (accumulator CORR, 2^11 too large, folded in
// by the reader with an exact 2^-11) -- three MMAs, two TMEM accumulators, so the small terms never get truncated
// against the large running sum.
//
// One CTA = (128 query rows, one head, one utterance):
//   S  = Q K^T   M = 128 queries, N = 128 keys per block, K = 64:  4 K-steps x 3 MMAs, accumulators S_main / S_corr
//   P  = exp(scale * (S - rowmax))            one thread per query row = one TMEM lane: row max / row sum need no shuffles
//   O += P V     M = 128 queries, N = 64 (d), K = 128 keys:        8 K-steps x 3 MMAs, accumulators O_main / O_corr
// Keys are processed in blocks of 128.  With more than one block the row maxima are found in a first sweep over S (QK^T is
// recomputed in the second sweep: cheap, and O never needs rescaling inside TMEM); N <= 128 takes a single sweep.
// Operands are staged into the K-major no-swizzle ("interleave") UMMA layout by the row threads: Q / K / P as
// [k-chunk][row][8 x fp16], V transposed on the fly to [key-chunk][d][8 keys] (lanes run over d: coalesced global reads).
// Warps 0-3: row workers (TMEM lane quarter = warp id), warp 4: TMEM allocation + single-thread MMA issue.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace st2 {
extern long long g_launches;

namespace atc {

using namespace st2::ptx;

constexpr int QB = 128, KB = 128, HD = 64;
constexpr int THREADS = 160;
constexpr float LO_SCALE = 2048.0f, LO_UNSCALE = 1.0f / 2048.0f;
constexpr int ROWS16 = 16;                                   // bytes per operand row chunk (8 fp16)
constexpr int QK_LBO = QB * ROWS16;                          // 2048: distance between 8-wide k-chunks of Q / K / P
constexpr int V_LBO = HD * ROWS16;                           // 1024: distance between 8-key chunks of V^T
constexpr int Q_PLANE = (HD / 8) * QK_LBO;                   // 16 KB
constexpr int K_PLANE = (HD / 8) * QK_LBO;                   // 16 KB
constexpr int V_PLANE = (KB / 8) * V_LBO;                    // 16 KB
constexpr int P_PLANE = (KB / 8) * QK_LBO;                   // 32 KB
constexpr int SM_Q = 0, SM_K = SM_Q + 2 * Q_PLANE, SM_V = SM_K + 2 * K_PLANE, SM_P = SM_V + 2 * V_PLANE, SM_BAR = SM_P + 2 * P_PLANE;
constexpr int SM_TOTAL = SM_BAR + 64;
constexpr int T_SMAIN = 0, T_SCORR = 128, T_OMAIN = 256, T_OCORR = 320, TMEM_COLS = 512;

struct Args {
  const float* q; long long q_ld;
  const float* k; const float* v; long long kv_ld;
  float* out; long long out_ld;
  const int* lengths;
  int N, H;
  float scale;
};

__device__ __forceinline__ uint32_t idesc_f16(int n) {
  uint32_t d = 0;
  d |= 1u << 4;                       // D = F32, A = B = F16 (format 0), K-major
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& p0, uint32_t& p1) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((x0 - hf.x) * LO_SCALE, (x1 - hf.y) * LO_SCALE);
  p0 = *reinterpret_cast<const uint32_t*>(&h);
  p1 = *reinterpret_cast<const uint32_t*>(&l);
}
// 8 fp32 -> one 16-byte row of each plane
__device__ __forceinline__ void store_row8(uint8_t* plane0, int plane_bytes, size_t off, const float (&x)[8]) {
  uint32_t a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2(x[2 * i], x[2 * i + 1], a[i], b[i]);
  *reinterpret_cast<uint4*>(plane0 + off) = make_uint4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<uint4*>(plane0 + plane_bytes + off) = make_uint4(b[0], b[1], b[2], b[3]);
}

__global__ void __launch_bounds__(THREADS, 1) attention_tc_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_s = sbase + SM_BAR, bar_o = sbase + SM_BAR + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 16);
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int N = a.N;
  const int klen = a.lengths ? min(a.lengths[b], N) : N;
  const int nkb = max(1, (klen + KB - 1) / KB);
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool worker = tid < QB;
  const int row = qb * QB + tid;                    // query row of this worker == its TMEM lane
  const long long rowbase = (long long)b * N;

  // ---- stage Q (once): worker t -> query row t, eight 8-wide chunks of d
  if (worker) {
    const float* qr = a.q + (rowbase + min(row, N - 1)) * a.q_ld + h * HD;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      float x[8];
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(qr + 8 * c)), v1 = __ldg(reinterpret_cast<const float4*>(qr + 8 * c + 4));
      x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
      if (row >= N) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = 0.f;
      }
      store_row8(smem + SM_Q, Q_PLANE, (size_t)c * QK_LBO + (size_t)tid * ROWS16, x);
    }
  }
  auto stage_k = [&](int kb) {      // worker t -> key kb*128 + t
    const int key = kb * KB + tid;
    const float* kr = a.k + (rowbase + min(key, N - 1)) * a.kv_ld + h * HD;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      float x[8];
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(kr + 8 * c)), v1 = __ldg(reinterpret_cast<const float4*>(kr + 8 * c + 4));
      x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
      if (key >= klen) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = 0.f;
      }
      store_row8(smem + SM_K, K_PLANE, (size_t)c * QK_LBO + (size_t)tid * ROWS16, x);
    }
  };
  auto stage_v = [&](int kb) {      // worker t -> d = t % 64, key chunks (t / 64) * 8 .. + 7; V^T rows of 8 keys
    const int d = tid & 63, c0 = (tid >> 6) * 8;
#pragma unroll 2
    for (int c = c0; c < c0 + 8; ++c) {
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int key = kb * KB + c * 8 + i;
        x[i] = key < klen ? __ldg(a.v + (rowbase + key) * a.kv_ld + h * HD + d) : 0.f;
      }
      store_row8(smem + SM_V, V_PLANE, (size_t)c * V_LBO + (size_t)d * ROWS16, x);
    }
  };
  auto issue_s = [&]() {            // S_main = Qh Kh^T ; S_corr = Qh Kl'^T + Ql' Kh^T
    const uint32_t id = idesc_f16(KB);
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const uint32_t off = (uint32_t)ks * 2 * QK_LBO;
      const uint64_t qh = make_desc(sbase + SM_Q + off, QK_LBO, 128), ql = make_desc(sbase + SM_Q + Q_PLANE + off, QK_LBO, 128);
      const uint64_t kh = make_desc(sbase + SM_K + off, QK_LBO, 128), kl = make_desc(sbase + SM_K + K_PLANE + off, QK_LBO, 128);
      tc_mma(tmem + T_SMAIN, qh, kh, id, ks ? 1u : 0u);
      tc_mma(tmem + T_SCORR, qh, kl, id, ks ? 1u : 0u);
      tc_mma(tmem + T_SCORR, ql, kh, id, 1u);
    }
    tc_commit(bar_s);
  };
  auto issue_o = [&](bool first) {  // O_main += Ph Vh ; O_corr += Ph Vl' + Pl' Vh
    const uint32_t id = idesc_f16(HD);
#pragma unroll
    for (int ks = 0; ks < KB / 16; ++ks) {
      const uint32_t po = (uint32_t)ks * 2 * QK_LBO, vo = (uint32_t)ks * 2 * V_LBO;
      const uint64_t ph = make_desc(sbase + SM_P + po, QK_LBO, 128), pl = make_desc(sbase + SM_P + P_PLANE + po, QK_LBO, 128);
      const uint64_t vh = make_desc(sbase + SM_V + vo, V_LBO, 128), vl = make_desc(sbase + SM_V + V_PLANE + vo, V_LBO, 128);
      const uint32_t acc = (first && ks == 0) ? 0u : 1u;
      tc_mma(tmem + T_OMAIN, ph, vh, id, acc);
      tc_mma(tmem + T_OCORR, ph, vl, id, acc);
      tc_mma(tmem + T_OCORR, pl, vh, id, 1u);
    }
    tc_commit(bar_o);
  };
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  // S row of this worker, 32 columns at a time: s = scale * (main + corr * 2^-11)
  auto read_s32 = [&](int c0, float (&s)[32]) {
    float t[32];
    tmem_ld32(lane_base + T_SMAIN + c0, s);
    tmem_ld32(lane_base + T_SCORR + c0, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) s[j] = fmaf(t[j], LO_UNSCALE, s[j]) * a.scale;
  };

  uint32_t ph_s = 0, ph_o = 0;
  float m = -INFINITY, l = 0.f;
  const bool two_pass = nkb > 1;
  // ---- sweep 1 (only when there are several key blocks): row maxima
  if (two_pass) {
    for (int kb = 0; kb < nkb; ++kb) {
      if (worker) stage_k(kb);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == QB) { tc_fence_after(); issue_s(); }
      mbar_wait(bar_s, ph_s); ph_s ^= 1;
      tc_fence_after();
      if (worker) {
#pragma unroll 1
        for (int c0 = 0; c0 < KB; c0 += 32) {
          float s[32];
          read_s32(c0, s);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (kb * KB + c0 + j < klen) m = fmaxf(m, s[j]);
        }
      }
      tc_fence_before();
      __syncthreads();       // S consumed, K buffer free (its MMAs have completed: bar_s)
    }
  }
  // ---- sweep 2: P and O
  for (int kb = 0; kb < nkb; ++kb) {
    if (kb > 0) { mbar_wait(bar_o, ph_o); ph_o ^= 1; }     // previous PV MMAs done: K / V / P buffers are free
    if (worker) { stage_k(kb); stage_v(kb); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == QB) { tc_fence_after(); issue_s(); }
    mbar_wait(bar_s, ph_s); ph_s ^= 1;
    tc_fence_after();
    if (worker) {
      if (!two_pass) {
#pragma unroll 1
        for (int c0 = 0; c0 < KB; c0 += 32) {
          float s[32];
          read_s32(c0, s);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < klen) m = fmaxf(m, s[j]);
        }
      }
#pragma unroll 1
      for (int c0 = 0; c0 < KB; c0 += 32) {
        float s[32];
        read_s32(c0, s);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float p = (kb * KB + c0 + j < klen) ? expf(s[j] - m) : 0.f;
          s[j] = p;
          l += p;
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = s[8 * cc + j];
          store_row8(smem + SM_P, P_PLANE, (size_t)(c0 / 8 + cc) * QK_LBO + (size_t)tid * ROWS16, x);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == QB) { tc_fence_after(); issue_o(kb == 0); }
  }
  mbar_wait(bar_o, ph_o);
  tc_fence_after();
  if (worker) {     // tcgen05.ld is warp-collective: every lane of the worker warps runs the loads, only the stores are predicated
    const float inv = 1.0f / l;
    float* orow = a.out + (rowbase + min(row, N - 1)) * a.out_ld + h * HD;
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 32) {
      float o[32], t[32];
      tmem_ld32(lane_base + T_OMAIN + c0, o);
      tmem_ld32(lane_base + T_OCORR + c0, t);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 w;
        w.x = fmaf(t[j], LO_UNSCALE, o[j]) * inv; w.y = fmaf(t[j + 1], LO_UNSCALE, o[j + 1]) * inv;
        w.z = fmaf(t[j + 2], LO_UNSCALE, o[j + 2]) * inv; w.w = fmaf(t[j + 3], LO_UNSCALE, o[j + 3]) * inv;
        if (row < N) *reinterpret_cast<float4*>(orow + c0 + j) = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
  }
}

}  // namespace atc
}  // namespace st2

using namespace st2;

extern "C" {

int st2_attention_tc_supported(long long q_ld, long long kv_ld, long long out_ld, int D) {
  return D == atc::HD && (q_ld % 4) == 0 && (kv_ld % 4) == 0 && (out_ld % 4) == 0;
}

int st2_attention_tc(const float* q, long long q_ld, const float* k, const float* v, long long kv_ld, float* out, long long out_ld,
                     const int* lengths, int B, int N, int H, int D, float scale, void* stream) {
  ST2_REQUIRE(q && k && v && out && B > 0 && N > 0 && H > 0, "st2_attention_tc", "bad args");
  ST2_REQUIRE(st2_attention_tc_supported(q_ld, kv_ld, out_ld, D), "st2_attention_tc", "head_features must be 64, row strides multiples of 4");
  ST2_REQUIRE(((reinterpret_cast<size_t>(q) | reinterpret_cast<size_t>(k) | reinterpret_cast<size_t>(v) | reinterpret_cast<size_t>(out)) & 15) == 0,
              "st2_attention_tc", "pointers must be 16-byte aligned");
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_done[dev]) {
    cudaFuncSetAttribute(atc::attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::SM_TOTAL);
    attr_done[dev] = true;
  }
  atc::Args a;
  a.q = q; a.q_ld = q_ld; a.k = k; a.v = v; a.kv_ld = kv_ld; a.out = out; a.out_ld = out_ld; a.lengths = lengths;
  a.N = N; a.H = H; a.scale = scale;
  dim3 grid(cdiv(N, atc::QB), H, B);
  atc::attention_tc_kernel<<<grid, atc::THREADS, atc::SM_TOTAL, (cudaStream_t)stream>>>(a);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_attention_tc");
  return 0;
}

}  // extern "C"
