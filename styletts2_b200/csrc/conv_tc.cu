// Tensor-core (tcgen05 / TMEM) implicit-GEMM Conv1d for the vocoder's AdaIN ResBlocks -- sm_100a.
//
//   D[co (M=128), t (N=256)] = sum_{tap} sum_{ci} W_tap[co, ci] * z[ci, t + tap*dil - pad],   z = snake/lrelu(a*x+b)
//
// Precision recipe (decided with the CPU oracle, DESIGN.md "precision"): fp32 operands are split into
// bf16 hi + bf16 lo and every product is evaluated as hi*hi + hi*lo + lo*hi on the 5th-gen tensor
// cores with fp32 accumulation in TMEM (error ~2^-16 relative per product, 2.8e-5 max-abs on the
// waveform vs 1.3e-3 for single-pass TF32 and 1.2e-2 for plain bf16).
//
// Mapping:
//  * A operand = weights  [128 co x 16 ci] bf16, K-major, no-swizzle "interleave" layout (8-row x 16-byte core
//    matrices, rows contiguous at 16 B pitch).  Pre-arranged in HBM so that one pipeline stage (tap, 32 ci,
//    hi+lo) is ONE contiguous 16 KB block moved by a single 1-D TMA bulk copy (cp.async.bulk) that signals
//    an mbarrier.
//  * B operand = activations [256 t x 16 ci] bf16, K-major, same interleave layout: for each group of 8
//    input channels the frame window is a column of 16-byte rows, so a conv tap is just a descriptor start
//    address shifted by tap*dil rows (16 B granularity) -- the window is staged ONCE per 32-channel block
//    (AdaIN affine + Snake/LeakyReLU + hi/lo split fused into the staging) and re-used by all K taps.
//  * D accumulators live in TMEM (2 x 256 columns, double buffered so the epilogue of tile i overlaps the
//    MMAs of tile i+1); the epilogue reads them with tcgen05.ld, transposes 32x32 blocks through shared
//    memory for coalesced row stores and fuses bias, residual, MRF accumulation and the InstanceNorm
//    partial statistics (count, mean, M2) exactly like the SIMT kernel.
//  * Warp roles: warp 0 = TMEM alloc + single-thread MMA issue, warp 1 = weight TMA producer,
//    warps 2-11 = activation stagers, warps 12-19 = epilogue.
//    Persistent CTAs (one per SM) loop over tiles.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace st2 {
extern long long g_launches;

namespace tc {

constexpr int TN = 256;
constexpr int TM = 128;
constexpr int CB = 16;                          // input channels per pipeline block (2 K-chunks of 8 = one UMMA K step)
constexpr int KCB = CB / 8;                     // K-chunks per block
constexpr int W_STAGES = 6;
constexpr int W_STAGE_BYTES = 2 * KCB * TM * 16;  // (hi|lo) x 2 k-chunks x 128 co x 16 B = 8 KB
constexpr int RW_MAX = 312;                     // TN + (K-1)*dil rounded up to 8, max
constexpr int RWP_MAX = RW_MAX + 2;             // chunk pitch in rows: == 2 (mod 8) -> conflict-free 128-bit staging stores
constexpr int ACT_HALF_BYTES = KCB * RWP_MAX * 16;  // one of hi / lo for one 16-channel block
constexpr int ACT_BUF_BYTES = 2 * ACT_HALF_BYTES;
constexpr int RAW_STAGES = 4;                   // cp.async ring of raw fp32 frame windows: 3 blocks (60 KB) in flight per SM
constexpr int RAW_BYTES = CB * RWP_MAX * 4;     // 20 KB
constexpr int CIN_PAD_MAX = 1120;
constexpr int NUM_STAGERS = 320;                // 10 warps
constexpr int NUM_EPI = 256;                    // 8 warps: two per TMEM lane quarter (each takes half of the columns)
constexpr int THREADS = 64 + NUM_STAGERS + NUM_EPI;  // 640 = 20 warps (register allocation granularity: 4 warps)
constexpr int TPITCH = 36;                      // epilogue transpose row pitch (floats), 16-byte aligned rows

constexpr int SM_W = 0;
constexpr int SM_ACT = SM_W + W_STAGES * W_STAGE_BYTES;
constexpr int SM_RAW = SM_ACT + 2 * ACT_BUF_BYTES;
constexpr int SM_COEF = SM_RAW + RAW_STAGES * RAW_BYTES;
constexpr int SM_EPI = SM_COEF + 4 * CIN_PAD_MAX * 4;
constexpr int SM_BAR = SM_EPI + 8 * (32 * TPITCH + 32) * 4;
constexpr int SM_TOTAL = SM_BAR + 256;

// barrier slots (8 B each) inside SM_BAR
constexpr int B_WFULL = 0, B_WEMPTY = 6, B_AFULL = 12, B_AEMPTY = 14, B_TFULL = 16, B_TEMPTY = 18, B_COUNT = 20;

// Optional per-role cycle trace of CTA 0 (debug/profiling aid; null in production).
__device__ long long* g_trace = nullptr;
constexpr int TRACE_TILES = 16, TRACE_K = 8;  // [role 4][tile 16][8 counters]
__device__ __forceinline__ void trace_put(int role, int it, int k, long long v) {
  if (g_trace && blockIdx.x == 0 && it < TRACE_TILES) g_trace[(role * TRACE_TILES + it) * TRACE_K + k] = v;
}

using namespace st2::ptx;

__device__ __forceinline__ long long mbar_wait_timed(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
  }
  return clock64() - t0;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256.
__device__ __forceinline__ uint32_t make_idesc() {
  uint32_t d = 0;
  d |= 1u << 4;                 // c_format = F32
  d |= 1u << 7;                 // a_format = BF16
  d |= 1u << 10;                // b_format = BF16
  d |= (uint32_t)(TN >> 3) << 17;  // n_dim
  d |= (uint32_t)(TM >> 4) << 24;  // m_dim
  return d;
}


// Output channels of a 128-row tile are spread over the four TMEM lane quarters (an epilogue warp can only read the
// quarter warp_id % 4): quarter q holds channels co0 + q*rq .. + rq-1 with rq = ceil(min(Cout - co0, 128) / 4).  For full
// tiles rq = 32 (identity); for narrow layers (HiFi-GAN C = 64 / 32, conv_post) all eight epilogue warps get work.
__host__ __device__ __forceinline__ int rows_per_quarter(int Cout, int cob) {
  int rem = Cout - cob * TM;
  if (rem > TM) rem = TM;
  return (rem + 3) >> 2;
}

struct TileCoord {
  int b, cob, tq;
};
__device__ __forceinline__ TileCoord tile_coord(int tile, int n_tq, int n_cob) {
  TileCoord c;
  c.tq = tile % n_tq;
  const int r = tile / n_tq;
  c.cob = r % n_cob;
  c.b = r / n_cob;
  return c;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ---------------------------------------------------------------------------------------------
// Stager inner work: AdaIN affine + activation + bf16 hi/lo split + two 128-bit stores for one frame row
// (8 channels of one K-chunk).  Templated on the activation so the per-element code has no branches.
template <int ACT>
__device__ __forceinline__ void stage_row(const float (&x)[8], const float (&pa)[8], const float (&pb)[8], const float (&al)[8],
                                          const float (&ia)[8], float slope, bool inb, uint8_t* hi_dst, uint8_t* lo_dst) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float z = fmaf(x[j], pa[j], pb[j]);
    if (ACT == ST2_ACT_SNAKE) {
      const float sn = __sinf(al[j] * z);
      z = fmaf(ia[j], sn * sn, z);
    } else if (ACT == ST2_ACT_LRELU) {
      z = z > 0.f ? z : z * slope;
    }
    v[j] = inb ? z : 0.f;  // zero padding applies AFTER the activation
  }
  uint32_t hp[4], lp[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    hp[q] = pack_bf16(v[2 * q], v[2 * q + 1]);
    const float h0 = __uint_as_float(hp[q] << 16), h1 = __uint_as_float(hp[q] & 0xFFFF0000u);
    lp[q] = pack_bf16(v[2 * q] - h0, v[2 * q + 1] - h1);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
}

// Epilogue row loop for one 32x32 accumulator block already transposed into T: row r of the block is output
// channel (co_base + r); lane = column.  yp points at (row 0, this lane's column); rv[] holds the 32 residual
// values of this lane's column (prefetched before the TMEM load so their latency is hidden).
template <bool RES, int ACC, bool STATS>
__device__ __forceinline__ void epi_rows(float* T, const float* bsm, float* yp, const float (&rv)[32], long long ystride,
                                         int rmax, bool tv, int lane, float out_div, float acc_div) {
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 8) {
    if (r0 < rmax) {
      float yo[8];
      if (ACC) {
#pragma unroll
        for (int i = 0; i < 8; ++i) yo[i] = (tv && (r0 + i) < rmax) ? yp[i * ystride] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + i;
        float val = T[r * TPITCH + lane] + bsm[r];
        if (RES) val += rv[r];
        if (out_div != 1.0f) val = __fdiv_rn(val, out_div);
        if (ACC == 1) val = yo[i] + val;
        if (ACC == 2) val = __fdiv_rn(yo[i] + val, acc_div);
        if (tv && r < rmax) yp[i * ystride] = val;
        if (STATS) T[r * TPITCH + lane] = val;
      }
      yp += 8 * ystride;
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
conv1d_tc_kernel(const st2_conv_args a, const uint4* __restrict__ wtc, const int ncb, const int RW, const int ntiles,
                 const int n_tq, const int n_cob) {
  // RW = window rows (TN + (K-1)*dil, rounded up to 8); RWP = chunk pitch in rows
  const int RWP = RW + 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + SM_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * B_COUNT);

  if (tid == 0) {
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(BAR(B_WFULL + i), 1); mbar_init(BAR(B_WEMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(B_AFULL + i), NUM_STAGERS);
      mbar_init(BAR(B_AEMPTY + i), 1);
      mbar_init(BAR(B_TFULL + i), 1);
      mbar_init(BAR(B_TEMPTY + i), NUM_EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int K = a.K;

  if (warp == 0) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc();
      const uint32_t lbo_a = TM * 16, lbo_b = (uint32_t)RWP * 16;
      int ws = 0, wph = 0, as = 0, aph = 0;
      auto wait_pumping = [&](uint32_t bar, uint32_t parity) -> long long { return mbar_wait_timed(bar, parity); };
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const long long tt0 = clock64();
        long long w_te = wait_pumping(BAR(B_TEMPTY + buf), ((it >> 1) & 1) ^ 1), w_af = 0, w_wf = 0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * TN;
        uint32_t first = 1;
        for (int cb = 0; cb < ncb; ++cb) {
          w_af += wait_pumping(BAR(B_AFULL + as), aph);
          tc_fence_after();
          const uint32_t act_hi = sbase + SM_ACT + as * ACT_BUF_BYTES;
          const uint32_t act_lo = act_hi + ACT_HALF_BYTES;
          for (int tap = 0; tap < K; ++tap) {
            w_wf += wait_pumping(BAR(B_WFULL + ws), wph);
            tc_fence_after();
            const uint32_t w_hi = sbase + SM_W + ws * W_STAGE_BYTES;
            const uint32_t w_lo = w_hi + W_STAGE_BYTES / 2;
            const uint32_t row_off = (uint32_t)(tap * a.dil) * 16;
            {
              const uint64_t da_hi = make_desc(w_hi, lbo_a, 128), da_lo = make_desc(w_lo, lbo_a, 128);
              const uint64_t db_hi = make_desc(act_hi + row_off, lbo_b, 128), db_lo = make_desc(act_lo + row_off, lbo_b, 128);
              tc_mma(d_tmem, da_hi, db_hi, idesc, first ? 0u : 1u);
              first = 0;
              tc_mma(d_tmem, da_hi, db_lo, idesc, 1u);
              tc_mma(d_tmem, da_lo, db_hi, idesc, 1u);
            }
            tc_commit(BAR(B_WEMPTY + ws));
            if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
          }
          tc_commit(BAR(B_AEMPTY + as));
          if (++as == 2) { as = 0; aph ^= 1; }
        }
        tc_commit(BAR(B_TFULL + buf));
        trace_put(0, it, 0, tt0); trace_put(0, it, 1, clock64()); trace_put(0, it, 2, w_te); trace_put(0, it, 3, w_af);
        trace_put(0, it, 4, w_wf);
      }
    }
  } else if (warp == 1) {
    // ================================================================ weight producer (1-D TMA bulk copies)
    if (lane == 0) {
      int ws = 0, wph = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord tc_ = tile_coord(tile, n_tq, n_cob);
        for (int cb = 0; cb < ncb; ++cb) {
          for (int tap = 0; tap < K; ++tap) {
            mbar_wait(BAR(B_WEMPTY + ws), wph ^ 1);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(wtc) + ((size_t)((tap * n_cob + tc_.cob) * ncb + cb)) * W_STAGE_BYTES;
            mbar_expect_tx(BAR(B_WFULL + ws), W_STAGE_BYTES);
            bulk_g2s(sbase + SM_W + ws * W_STAGE_BYTES, src, W_STAGE_BYTES, BAR(B_WFULL + ws));
            if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp < 2 + NUM_STAGERS / 32) {
    // ================================================================ activation stagers
    // Raw fp32 frame windows travel HBM -> shared memory with cp.async (no registers held while in flight): a ring
    // of RAW_STAGES 16-channel blocks keeps ~60 KB per SM outstanding, which is what it takes to cover HBM latency.
    // Conversion (AdaIN affine + Snake/LeakyReLU + bf16 hi/lo split) reads the landed block from shared memory.
    const int st = tid - 64;  // 0..319
    const int Lin_ = a.Lin, pre_act_ = a.pre_act, Cin_ = a.Cin;
    const float slope_ = a.pre_slope;
    float* coef = reinterpret_cast<float*>(smem + SM_COEF);
    const int cin_pad = ncb * CB;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_blocks = my_tiles * ncb;
    // issue mapping: thread -> one channel of the block (st / 20) and every 20th frame row: pointer increments only
    const int ich = st / 20, ir0 = st - ich * 20;
    int i_tile = -1, i_b = 0, i_g0 = 0;  // producer-side tile state (runs 3 blocks ahead of the conversion)
    const float* i_xb = a.x;
    auto issue = [&](int g) {
      if (g < total_blocks) {
        const int tl = g / ncb, cb = g - tl * ncb;
        if (tl != i_tile) {
          i_tile = tl;
          const TileCoord tc_ = tile_coord(blockIdx.x + tl * gridDim.x, n_tq, n_cob);
          i_b = tc_.b;
          i_xb = a.x + (long long)tc_.b * a.x_bstride;
          i_g0 = tc_.tq * TN - a.pad;
        }
        const int c = cb * CB + ich;
        const int rlo = (c < Cin_) ? max(0, -i_g0) : RW;   // rows [rlo, rhi) are inside the tensor; the rest zero-fill
        const int rhi = min(RW, Lin_ - i_g0);
        const float* src = i_xb + (long long)min(c, Cin_ - 1) * Lin_ + (i_g0 + ir0);
        uint32_t dst = sbase + SM_RAW + (g % RAW_STAGES) * RAW_BYTES + (uint32_t)(ich * RWP + ir0) * 4;
        for (int r = ir0; r < RW; r += 20) {
          const bool ok = (r >= rlo) && (r < rhi);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(ok ? src : a.x), "r"(ok ? 4 : 0) : "memory");
          src += 20;
          dst += 80;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    issue(1);
    issue(2);
    int as = 0, aph = 0;
    int last_b = -1, c_tile = -1, c_b = 0, c_g0 = 0;
    const int kc = st & 1, rg = st >> 1;  // conversion mapping: 2 K-chunks x 160 row groups
    constexpr int NRC = (RW_MAX + 159) / 160;
    for (int g = 0; g < total_blocks; ++g) {
      const int tl = g / ncb, cb = g - tl * ncb;
      if (tl != c_tile) {
        c_tile = tl;
        const TileCoord tc_ = tile_coord(blockIdx.x + tl * gridDim.x, n_tq, n_cob);
        c_b = tc_.b;
        c_g0 = tc_.tq * TN - a.pad;
      }
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(NUM_STAGERS));  // block g landed for everybody; block g-1 fully converted
      issue(g + 3);                                          // reuses the slot of block g-1
      if (c_b != last_b) {
        for (int c = st; c < cin_pad; c += NUM_STAGERS) {
          float pa = 0.f, pb = 0.f, al = 1.f;  // padded channels stage exact zeros
          if (c < Cin_) {
            pa = 1.f;
            if (a.pre_a) { pa = a.pre_a[c_b * Cin_ + c]; pb = a.pre_b[c_b * Cin_ + c]; }
            if (pre_act_ == ST2_ACT_SNAKE) al = a.pre_alpha[c];
          }
          coef[c] = pa; coef[CIN_PAD_MAX + c] = pb; coef[2 * CIN_PAD_MAX + c] = al; coef[3 * CIN_PAD_MAX + c] = 1.0f / al;
        }
        asm volatile("bar.sync 2, %0;" ::"n"(NUM_STAGERS));
        last_b = c_b;
      }
      const int c0 = cb * CB + kc * 8;
      float pa[8], pb[8], al[8], ia[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        pa[j] = coef[c0 + j]; pb[j] = coef[CIN_PAD_MAX + c0 + j]; al[j] = coef[2 * CIN_PAD_MAX + c0 + j];
        ia[j] = coef[3 * CIN_PAD_MAX + c0 + j];
      }
      const float* raw = reinterpret_cast<const float*>(smem + SM_RAW + (g % RAW_STAGES) * RAW_BYTES) + (kc * 8) * RWP;
      mbar_wait(BAR(B_AEMPTY + as), aph ^ 1);
      uint8_t* hi = smem + SM_ACT + as * ACT_BUF_BYTES;
      uint8_t* lo = hi + ACT_HALF_BYTES;
#pragma unroll
      for (int i = 0; i < NRC; ++i) {
        const int r = rg + 160 * i;
        if (r < RW) {
          const int gt = c_g0 + r;
          const bool inb = (gt >= 0) && (gt < Lin_);
          float xv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[j] = raw[j * RWP + r];
          uint8_t* hd = hi + (size_t)(kc * RWP + r) * 16;
          uint8_t* ld = lo + (size_t)(kc * RWP + r) * 16;
          if (pre_act_ == ST2_ACT_SNAKE) stage_row<ST2_ACT_SNAKE>(xv, pa, pb, al, ia, slope_, inb, hd, ld);
          else if (pre_act_ == ST2_ACT_LRELU) stage_row<ST2_ACT_LRELU>(xv, pa, pb, al, ia, slope_, inb, hd, ld);
          else stage_row<ST2_ACT_NONE>(xv, pa, pb, al, ia, slope_, inb, hd, ld);
        }
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(BAR(B_AFULL + as));
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    // ================================================================ epilogue (8 warps)
    const int ewi = warp - (2 + NUM_STAGERS / 32);  // 0..7
    const int ew = warp & 3;                         // TMEM lane quarter this warp may access
    const int half = ewi >> 2;                       // which 128 columns of the tile this warp handles
    float* T = reinterpret_cast<float*>(smem + SM_EPI) + ewi * (32 * TPITCH + 32);
    float* bsm = T + 32 * TPITCH;
    const int y_len_ = a.y_len, res_len_ = a.res_len, acc_ = a.accum_mode, out_act_ = a.out_act, Cout_ = a.Cout, Lq_ = a.Lq;
    const int ytst_ = a.y_tstride, ytoff_ = a.y_toffset, rshift_ = a.res_shift;
    const float out_div_ = a.out_div, acc_div_ = a.accum_div;
    const bool has_stats = a.stats != nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const TileCoord tc_ = tile_coord(tile, n_tq, n_cob);
      const int buf = it & 1;
      const int rq = rows_per_quarter(Cout_, tc_.cob);
      const int co_base = tc_.cob * TM + ew * rq;
      const int t0 = tc_.tq * TN;
      const int ncols = min(TN, Lq_ - t0);
      const int rmax = min(rq, Cout_ - co_base);  // warp-uniform (may be <= 0 for padded channel blocks)
      bsm[lane] = (a.bias && lane < rmax) ? a.bias[co_base + lane] : 0.f;
      const long long ett0 = clock64();
      const long long w_tf = mbar_wait_timed(BAR(B_TFULL + buf), (it >> 1) & 1);
      tc_fence_after();
      float* yb = a.y + (long long)tc_.b * a.y_bstride + (long long)co_base * y_len_;
      const float* rb = a.res ? a.res + (long long)tc_.b * a.res_bstride + (long long)co_base * res_len_ : nullptr;
      float s_n = 0.f, s_mean = 0.f, s_m2 = 0.f;  // running (count, mean, M2) of row (co_base + lane)
      long long tr_ld = 0, tr_st = 0, tr_ss = 0;
      for (int c0 = half * (TN / 2); c0 < (half + 1) * (TN / 2); c0 += 32) {
        float v[32];
        const long long q0 = clock64();
        float rv[32];
        {
          const int tcol = t0 + c0 + lane;
          const bool okc = (c0 + lane) < ncols;
          const float* rp0 = rb ? rb + ((tcol * ytst_ + ytoff_) >> rshift_) : nullptr;
#pragma unroll
          for (int r = 0; r < 32; ++r) rv[r] = (rb && okc && r < rmax) ? __ldg(rp0 + (long long)r * res_len_) : 0.f;
        }
        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TN + c0), v);
        const long long q1 = clock64();
        tr_ld += q1 - q0;
        if (c0 >= ncols || rmax <= 0) continue;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(&T[lane * TPITCH + 4 * q]) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        const int t = t0 + c0 + lane;
        const bool tv = (c0 + lane) < ncols;
        const int oidx = t * ytst_ + ytoff_;
        const int ridx = oidx >> rshift_;
        {
          float* yp = yb + oidx;
          const float* rp = rb ? rb + ridx : nullptr;
          const long long ys = y_len_, rs = res_len_;
          if (out_act_ == ST2_ACT_NONE) {
#define EPI(RES_, ACC_, ST_) epi_rows<RES_, ACC_, ST_>(T, bsm, yp, rv, ys, rmax, tv, lane, out_div_, acc_div_)
            if (has_stats) {
              if (rb) { if (acc_ == 0) EPI(true, 0, true); else if (acc_ == 1) EPI(true, 1, true); else EPI(true, 2, true); }
              else    { if (acc_ == 0) EPI(false, 0, true); else if (acc_ == 1) EPI(false, 1, true); else EPI(false, 2, true); }
            } else {
              if (rb) { if (acc_ == 0) EPI(true, 0, false); else if (acc_ == 1) EPI(true, 1, false); else EPI(true, 2, false); }
              else    { if (acc_ == 0) EPI(false, 0, false); else if (acc_ == 1) EPI(false, 1, false); else EPI(false, 2, false); }
            }
#undef EPI
          } else {  // rare generic path (output activation)
            for (int r = 0; r < rmax; ++r) {
              float val = 0.f;
              if (tv) {
                val = T[r * TPITCH + lane] + bsm[r];
                if (rb) val += rv[r];
                if (out_div_ != 1.0f) val = __fdiv_rn(val, out_div_);
                if (acc_ == 1) val = yp[(long long)r * ys] + val;
                else if (acc_ == 2) val = __fdiv_rn(yp[(long long)r * ys] + val, acc_div_);
                if (out_act_ == ST2_ACT_TANH) val = tanhf(val);
                yp[(long long)r * ys] = val;
              }
              if (has_stats) T[r * TPITCH + lane] = val;
            }
          }
        }
        const long long q2 = clock64();
        tr_st += q2 - q1;
        // ReflectionPad1d((1,0)) duplicate of the q==0 column (istftnet.py:365-366): value differs by its residual
        const bool dup_here = (a.dup_q0_to >= 0) && (t0 == 0) && (c0 == 0);
        float dupv = 0.f;
        if (dup_here && lane < rmax) {
          float val = v[0] + bsm[lane];
          if (rb) val += rb[(long long)lane * a.res_len + (a.dup_q0_to >> a.res_shift)];
          if (a.out_div != 1.0f) val = __fdiv_rn(val, a.out_div);
          float* p = yb + (long long)lane * a.y_len + a.dup_q0_to;
          if (a.accum_mode == 1) val = *p + val;
          else if (a.accum_mode == 2) val = __fdiv_rn(*p + val, a.accum_div);
          if (a.out_act == ST2_ACT_TANH) val = tanhf(val);
          *p = val;
          dupv = val;
        }
        if (has_stats) {
          __syncwarp();
          const int nv = min(32, ncols - c0);
          float w[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 t4 = *reinterpret_cast<const float4*>(&T[lane * TPITCH + 4 * q]);
            w[4 * q] = t4.x; w[4 * q + 1] = t4.y; w[4 * q + 2] = t4.z; w[4 * q + 3] = t4.w;
          }
          float cs = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) cs += (j < nv) ? w[j] : 0.f;
          const float cmean = cs / (float)nv;
          float cm2 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = (j < nv) ? w[j] - cmean : 0.f;
            cm2 = fmaf(d, d, cm2);
          }
          const float nn = s_n + (float)nv;  // Chan merge
          const float delta = cmean - s_mean;
          s_mean += delta * ((float)nv / nn);
          s_m2 += cm2 + delta * delta * (s_n * (float)nv / nn);
          s_n = nn;
          if (dup_here) {  // one extra sample for this row
            const float n2 = s_n + 1.0f;
            const float d2 = dupv - s_mean;
            s_mean += d2 / n2;
            s_m2 += d2 * d2 * (s_n / n2);
            s_n = n2;
          }
        }
        __syncwarp();
        tr_ss += clock64() - q2;
      }
      tc_fence_before();
      mbar_arrive(BAR(B_TEMPTY + buf));
      if (ewi == 0 && lane == 0) {
        trace_put(3, it, 0, ett0); trace_put(3, it, 1, clock64()); trace_put(3, it, 2, w_tf);
        trace_put(3, it, 3, tr_ld); trace_put(3, it, 4, tr_st); trace_put(3, it, 5, tr_ss);
      }
      if (a.stats && lane < rmax) {
        float* sp = a.stats + (((long long)tc_.b * a.Cout + co_base + lane) * a.stats_nparts + a.stats_part_offset + 2 * tc_.tq + half) * 3;
        sp[0] = s_n; sp[1] = s_mean; sp[2] = s_m2;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// fp32 [Cout,Cin,K] -> bf16 hi/lo stage blocks [K][n_cob][ncb][2][4][128][8]
__global__ void conv_tc_weight_layout_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin, int K,
                                             int n_cob, int ncb) {
  const long long total = (long long)K * n_cob * ncb * 2 * KCB * TM * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int j = (int)(r % 8); r /= 8;
    const int col = (int)(r % TM); r /= TM;
    const int kc = (int)(r % KCB); r /= KCB;
    const int hl = (int)(r % 2); r /= 2;
    const int cb = (int)(r % ncb); r /= ncb;
    const int cob = (int)(r % n_cob); r /= n_cob;
    const int tap = (int)r;
    const int rq = rows_per_quarter(Cout, cob);
    const int qq = col >> 5, rr = col & 31;
    const int co = cob * TM + qq * rq + rr, ci = cb * CB + kc * 8 + j;
    float v = 0.f;
    if (rr < rq && co < Cout && ci < Cin) v = w[((long long)co * Cin + ci) * K + tap];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = hl == 0 ? h : __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// ConvTranspose1d weight [Cin,Cout,K] -> S per-phase tensor-core blocks (phase r = J-tap stride-1 conv, see conv.cu)
__global__ void convT_tc_weight_layout_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin, int Cout, int K,
                                              int S, int P, int J, int n_cob, int ncb) {
  const long long per_phase = (long long)J * n_cob * ncb * 2 * KCB * TM * 8;
  const long long total = per_phase * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ph = (int)(i / per_phase);
    long long r = i % per_phase;
    const int j = (int)(r % 8); r /= 8;
    const int col = (int)(r % TM); r /= TM;
    const int kc = (int)(r % KCB); r /= KCB;
    const int hl = (int)(r % 2); r /= 2;
    const int cb = (int)(r % ncb); r /= ncb;
    const int cob = (int)(r % n_cob); r /= n_cob;
    const int kp = (int)r;  // tap of the phase conv
    const int rq = rows_per_quarter(Cout, cob);
    const int qq = col >> 5, rr = col & 31;
    const int co = cob * TM + qq * rq + rr, ci = cb * CB + kc * 8 + j;
    const int kk = (J - 1 - kp) * S + ((ph + P) % S);
    float v = 0.f;
    if (rr < rq && co < Cout && ci < Cin && kk < K) v = w[((long long)ci * Cout + co) * K + kk];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = hl == 0 ? h : __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

static int launch_tc(const st2_conv_args& a, const void* wtc, int max_ctas, cudaStream_t st) {
  const int n_tq = cdiv(a.Lq, TN), n_cob = cdiv(a.Cout, TM), ncb = cdiv(a.Cin, CB);
  const int rw = (TN + (a.K - 1) * a.dil + 7) & ~7;
  const int ntiles = a.B * n_cob * n_tq;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(conv1d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
  }
  int grid = ntiles < num_sms ? ntiles : num_sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  conv1d_tc_kernel<<<grid, THREADS, SM_TOTAL, st>>>(a, (const uint4*)wtc, ncb, rw, ntiles, n_tq, n_cob);
  ++g_launches;
  return 0;
}

}  // namespace tc
}  // namespace st2

using namespace st2;

extern "C" {

long long st2_conv_tc_weight_bytes(int Cout, int Cin, int K) {
  const int n_cob = cdiv(Cout, tc::TM), ncb = cdiv(Cin, tc::CB);
  return (long long)K * n_cob * ncb * tc::W_STAGE_BYTES;
}

int st2_conv_tc_weight_layout(const float* w, void* out, int Cout, int Cin, int K, void* stream) {
  ST2_REQUIRE(w && out && Cout > 0 && Cin > 0 && K > 0, "st2_conv_tc_weight_layout", "bad args");
  const int n_cob = cdiv(Cout, tc::TM), ncb = cdiv(Cin, tc::CB);
  tc::conv_tc_weight_layout_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out, Cout, Cin, K, n_cob, ncb);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_conv_tc_weight_layout");
  return 0;
}

int st2_conv_tc_supported(int Cin, int Cout, int K, int stride, int dil) {
  const int rw = (tc::TN + (K - 1) * dil + 7) & ~7;
  return stride == 1 && rw <= tc::RW_MAX && cdiv(Cin, tc::CB) * tc::CB <= tc::CIN_PAD_MAX;
}

int st2_conv1d_tc(const st2_conv_args* a, const void* wtc, int max_ctas, void* stream) {
  ST2_REQUIRE(a && a->x && wtc && a->y, "st2_conv1d_tc", "null pointer");
  ST2_REQUIRE(st2_conv_tc_supported(a->Cin, a->Cout, a->K, a->stride, a->dil), "st2_conv1d_tc", "unsupported shape");
  ST2_REQUIRE(a->pre_act != ST2_ACT_SNAKE || a->pre_alpha, "st2_conv1d_tc", "snake prologue needs alpha");
  const int n_tq = cdiv(a->Lq, tc::TN);
  ST2_REQUIRE(!a->stats || a->stats_nparts >= a->stats_part_offset + 2 * n_tq, "st2_conv1d_tc", "stats buffer too small (2 partials per 256-column tile)");
  tc::launch_tc(*a, wtc, max_ctas, (cudaStream_t)stream);
  ST2_CHECK_LAUNCH("st2_conv1d_tc");
  return 0;
}

int st2_debug_set_trace(void* buf) {
  long long* p = (long long*)buf;
  cudaError_t e = cudaMemcpyToSymbol(tc::g_trace, &p, sizeof(p));
  if (e != cudaSuccess) { set_error("st2_debug_set_trace", e); return (int)e; }
  return 0;
}

long long st2_convT_tc_weight_bytes(int Cin, int Cout, int K, int S) {
  const int J = (K + S - 1) / S;
  return (long long)S * st2_conv_tc_weight_bytes(Cout, Cin, J);
}

int st2_convT_tc_weight_layout(const float* w, void* out, int Cin, int Cout, int K, int S, int P, void* stream) {
  ST2_REQUIRE(w && out && Cout > 0 && Cin > 0 && K > 0 && S > 0, "st2_convT_tc_weight_layout", "bad args");
  const int J = (K + S - 1) / S;
  const int n_cob = cdiv(Cout, tc::TM), ncb = cdiv(Cin, tc::CB);
  tc::convT_tc_weight_layout_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out, Cin, Cout, K, S, P, J, n_cob, ncb);
  ++g_launches;
  ST2_CHECK_LAUNCH("st2_convT_tc_weight_layout");
  return 0;
}

int st2_conv_transpose1d_tc(const st2_conv_args* a0, const void* wtc, int K, int S, int P, int reflect_left1, void* stream) {
  ST2_REQUIRE(a0 && a0->x && wtc && a0->y, "st2_conv_transpose1d_tc", "null pointer");
  ST2_REQUIRE(K > 0 && S > 0 && P >= 0, "st2_conv_transpose1d_tc", "bad shape");
  const int J = (K + S - 1) / S;
  ST2_REQUIRE(st2_conv_tc_supported(a0->Cin, a0->Cout, J, 1, 1), "st2_conv_transpose1d_tc", "unsupported shape");
  const int parts = 2 * cdiv(a0->Lin, tc::TN);
  ST2_REQUIRE(!a0->stats || a0->stats_nparts >= S * parts, "st2_conv_transpose1d_tc", "stats buffer too small");
  const long long phase_bytes = st2_conv_tc_weight_bytes(a0->Cout, a0->Cin, J);
  for (int r = 0; r < S; ++r) {
    st2_conv_args a = *a0;
    const int cr = (r + P) / S;
    a.K = J;
    a.stride = 1;
    a.dil = 1;
    a.pad = (J - 1) - cr;
    a.Lq = a0->Lin;
    a.y_tstride = S;
    a.y_toffset = r + (reflect_left1 ? 1 : 0);
    a.y_len = a0->Lin * S + (reflect_left1 ? 1 : 0);
    a.stats_part_offset = r * parts;
    a.dup_q0_to = (reflect_left1 && r == 1) ? 0 : -1;
    tc::launch_tc(a, (const uint8_t*)wtc + (size_t)r * phase_bytes, 0, (cudaStream_t)stream);
  }
  ST2_CHECK_LAUNCH("st2_conv_transpose1d_tc");
  return 0;
}

}  // extern "C"
