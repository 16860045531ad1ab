// The tile kernel: a persistent, warp-specialised tcgen05 GEMM whose accumulators never leave
// the SM.  One CTA per SM walks a list of work items; inside an item it streams 128 x BN tiles
// of  S = A * B^T  (A = this rank's 128-row block, B = BN rows of the other modality; BN = 256
// for the S-only modes, 128 for the fused backward) through TMEM and consumes them in registers:
//
//   MODE_FWD  : per-row / per-column online log-sum-exp partials (ClipLoss) or the summed
//               softplus (SigLipLoss).  Nothing N x N is written anywhere (inference / no-grad).
//   MODE_FWDE : the training forward.  Same statistics, and the exponentials it evaluates anyway,
//               E = 2^(S2 - c) (c = the 32 x 64 sub-tile reference), leave the SM as a bf16 block
//               through a swizzled staging tile and a TMA store; the backward rescales that block
//               into the gradient of the logits instead of recomputing S.  SigLIP stores
//               G = sigmoid(z) - delta directly.
//   MODE_GW   : recomputes S and writes the exact gradient tile G (bf16) -- the guard's fallback
//               for the E block (a no-op launch otherwise) and the MRCLIP_BWD=gmat backend.
//   MODE_BWD  : MRCLIP_BWD=fused.  Recomputes S, turns it into G (bf16, written to shared memory
//               in the UMMA K-major swizzled layout) and immediately contracts it with B again:
//               dA[128, DC] += G[128,128] * B[128 cols, DC]   (second tcgen05 GEMM, accumulator
//               stationary in TMEM for the whole item).  O(N*D) memory, 7 GEMM units per step.
//
// Replaces, for the reference (src/open_clip/loss.py): the logits GEMMs :117-124, the two
// F.cross_entropy calls :135-136 and their autograd graph; SigLipLoss._loss :354-363.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (warp-uniform
// loop, one elected lane issues), warps 2..9 = epilogue (two warps per TMEM lane quarter, each
// taking half of the tile's columns, 64 at a time).
#pragma once
#include "peer_sync.cuh"
#include "ptx.cuh"
#include <cuda_bf16.h>
#include <math_constants.h>

namespace mrclip {

constexpr int kBM = 128;           // tile rows  (= TMEM lanes)
constexpr int kBN = 128;           // tile cols of S in the fused backward (BN template value there)
constexpr int kBK = 64;            // K block: 64 bf16 = one 128-byte swizzle row
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// GW:   recompute S, write the bf16 gradient tile G to global.
// FWDE: forward that also writes E = exp2(S2 - c) (bf16, c = the 32x64 sub-tile maximum it already tracks for the
//       column sums) so that the backward never recomputes S: G = E * (2^(c-lse_row) + 2^(c-lse_col)) is formed
//       on the fly inside the gradient GEMMs (gemm_kernel<.., XF>).  SigLIP has no normaliser, so its FWDE writes
//       G = sigmoid(z) - delta directly.
// FWDEU: FWDE that also accumulates, per row and column chunk, u = sum_j 2^(S2_ij - m) * S2_ij (the row-softmax
//        weighted logit sum, merged online like the LSE).  With it d logit_scale of a multi-rank local loss needs no
//        entropy arithmetic in the rescale pass:  s dL_r/ds = <dT_r, T_r> + ln2/(2n) * (R2(r,*) - R2(*,r)),
//        R2(q, r) = sum over rows of q and columns of r of Prow * S2  (not validated on hardware yet; MRCLIP_DS=fwd).
// RANK:  retrieval metrics (reference train.py:465-534, get_clip_metrics): the same S tiles with a rank-of-label epilogue
//        instead of the LSE.  Phase 0 scatters every row's positive logits (columns of the row's class) into a CSR list;
//        phase 1 counts, per row, the negatives above its best positive and, for 32 positives at a time held in
//        registers, the (negative, positive) pairs with negative > positive.  No N x N matrix, no sort.
enum : int { MODE_FWD = 0, MODE_BWD = 1, MODE_GW = 2, MODE_FWDE = 3, MODE_FWDEU = 4, MODE_RANK = 5 };
enum : int { LOSS_CLIP = 0, LOSS_SIGLIP = 1 };

struct TileParams {
  int m_rows;           // rows of A owned by this launch
  int n_cols;           // rows of B (= columns of S)
  int num_kb;           // ceil(d / 64)
  int num_rb;           // ceil(m_rows / 128)
  int tile_begin;       // first column tile handled by this launch
  int tile_end;         // one past the last column tile
  int tiles_per_chunk;  // column tiles per work item
  int num_chunks;       // chunks in [tile_begin, tile_end)
  int chunk_base;       // global index of the first chunk (FWD partial slots)
  int num_dc;           // BWD: number of D chunks
  int num_items;
  int label_offset;     // global column of row 0's positive
  int m_pad, n_pad, d_pad;
  const float* scale;   // device scalar: exp'd logit_scale
  const float* bias;    // device scalar or null (SigLIP)
  float w_own, w_oth;   // weights of the own-direction / other-direction softmax terms
  // FWD (clip)
  float2* row_part;     // [slots][m_pad]   (max2, sum)
  float* col_l;         // [bands][n_pad]
  float* col_c;         // [bands][n_pad/64]  sub-tile reference (max2) of the column sums and of E
  float* diag2;         // [m_pad]  positive logit in log2 units
  // FWD (siglip) / BWD scalar partials
  float2* sc_part;      // [num_items * 8]
  float2* sc_part2;     // [num_items * 8]  FWDE siglip: (sum G*C, sum G)
  // BWD
  const float* lse2_a;  // [m_rows]   own-direction LSE of each A row (log2 units)
  const float* lse2_b;  // [n_pad]    other-direction LSE of each B row, +inf padded
  float* dpart;         // [cs][m_pad][d_pad]
  // GW
  uint16_t* g_out;      // bf16 [m_pad][g_ld]
  long g_ld;
  const int* run_if;    // optional device flag: the kernel returns at once when *run_if == 0
  int ent;              // GW clip: scalar partials become (sum P_own log2 P_own, sum P_oth log2 P_oth)
  float* row_ent;       // FWDEU: [slots][m_pad]  u of each (row, column-chunk half), relative to row_part's max2
  // Multi-rank forward over NVLink peer memory (the all-gather of loss.py:51-57 overlapped with the tiles): the B rows of
  // source rank q were stored into this rank's buffer by q's pack2_push_kernel; a column tile may be loaded once
  // sig_ready[q] has reached *sig_epoch (peer_sync.cuh, CH_TEXT).  The chunk order is rotated so that the tiles on this
  // rank's own columns -- which wait for nobody -- run first, then the sources in ring order.
  // RANK (retrieval metrics)
  const int* rk_row_cls;   // [m_rows] dense class id of each row
  const int* rk_col_cls;   // [n_pad]  dense class id of each column (-1 in the padding)
  const int* rk_col_ord;   // [n_pad]  ordinal of column j among the columns of its class
  const long long* rk_off; // [m_rows] start of row i's positive list
  const int* rk_m;         // [m_rows] number of positives of row i
  float* rk_pos;           // positive logits (raw cosines): rk_pos[rk_off[i] + rk_col_ord[j]] = C_ij
  const float* rk_lmax;    // [m_rows] phase 1: largest positive of the row
  unsigned long long* rk_pairs;   // [m_rows] phase 1: += #{(k negative, t positive in this chunk): C_ik > C_it}
  int* rk_best;            // [m_rows] phase 1 (chunk 0 only): += #{k negative: C_ik > lmax_i}
  int rk_phase, rk_chunk0;
  const int* sig_ready; // this rank's sig block, CH_TEXT row; null: B is complete at launch
  const int* sig_epoch; // this rank's epoch of CH_TEXT
  int src_cols;         // columns per source rank
  int my_src;           // this rank
  int chunk_rot;        // item order: chunk (c + chunk_rot) % num_chunks is the c-th to run
};

template <int LEN, int OFF>
__device__ __forceinline__ void butterfly_step(float (&v)[64], uint32_t lane) {
  const bool upper = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < LEN / 2; ++i) {
    const float keep = upper ? v[i + LEN / 2] : v[i];
    const float send = upper ? v[i] : v[i + LEN / 2];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

__device__ __forceinline__ float log1p_from_exp(float e) {
  // log(1+e) for e in [0,1]; series below 1/8 (lg2.approx is not accurate enough near 1)
  const float series = e * (1.f + e * (-0.5f + e * (0.33333334f + e * (-0.25f + e * 0.2f))));
  const float direct = lg2f(1.f + e) * kLn2;
  return e <= 0.125f ? series : direct;
}

// BN = columns of one S tile: 256 for the S-only modes (FWD, GW: one N=256 MMA per K step, half the
// A-operand traffic per flop), 128 for the fused backward (TMEM must also hold the dA accumulator).
template <int MODE, int LOSS, int DC, int BN>
struct TileCfg {
  static_assert(BN == 128 || BN == 256, "BN must be 128 or 256");
  static constexpr bool kHasDa = (MODE == MODE_BWD);   // second GEMM fused in the same kernel
  static_assert(!kHasDa || BN == 128, "fused backward uses 128-wide S tiles");
  static constexpr int kSBufs = (!kHasDa) ? 2 : (DC == 256 ? 2 : 1);
  static constexpr int kTmemCols = (!kHasDa) ? 2 * BN : 512;
  static constexpr int kDaCol = kSBufs * BN;
  static constexpr int kNSub = (DC == 384) ? 2 : 1;
  static constexpr int kDN = DC / kNSub;
  // one pipeline slot: A(16K)+B(BN*128) for an S step, or one Bt block (<=32K) for a dA step
  static constexpr int kStageBytes = kHasDa ? 32768 : (kBM * kBK * 2 + BN * kBK * 2);
  static constexpr int kStages = kHasDa ? 5 : (BN == 256 ? 4 : 6);
  static constexpr int kGBytes = (MODE == MODE_BWD) ? 2 * 32768 : 0;
  static constexpr int kNumBars = 2 * kStages + 10;
  // GW: each epilogue warp keeps the other-direction LSE of its BN/2 columns in a private smem strip
  // (fetched before the accumulator wait, so the global-load latency is off the critical path)
  static constexpr bool kLseSmem = (MODE == MODE_GW && LOSS == LOSS_CLIP);
  static constexpr int kLseBytes = kLseSmem ? kEpiWarps * (BN / 2) * 4 : 0;
  // FWDE: one 32x64 bf16 staging tile per epilogue warp (128-byte swizzled rows) for the TMA store of E
  static constexpr bool kEOut = (MODE == MODE_FWDE || MODE == MODE_FWDEU);
  static constexpr bool kRowEnt = (MODE == MODE_FWDEU);
  static constexpr int kEBytes = kEOut ? kEpiWarps * 4096 : 0;
  static constexpr int kSmemBytes = kStages * kStageBytes + kGBytes + kLseBytes + kEBytes + kNumBars * 8 + 16 + 1024;
};

template <int MODE, int LOSS, int DC, int BN>
__global__ void __launch_bounds__(kThreads, 1)
tile_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmBt, const TileParams p) {
  using Cfg = TileCfg<MODE, LOSS, DC, BN>;
  constexpr int kStageBytes = Cfg::kStageBytes;
  constexpr int STAGES = Cfg::kStages;
  constexpr int NSB = Cfg::kSBufs;
  constexpr int DN = Cfg::kDN;
  constexpr int NSUB = Cfg::kNSub;
  constexpr uint32_t IDESC_S = make_idesc_bf16(kBM, BN);
  constexpr uint32_t IDESC_D = make_idesc_bf16(kBM, DN);

  if (p.run_if != nullptr && __ldg(p.run_if) == 0) return;   // grid-uniform: nobody touches a barrier or TMEM
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t stage_base = smem_u32(smem);
  const uint32_t g_base = stage_base + STAGES * kStageBytes;
  const uint32_t lse_smem = g_base + Cfg::kGBytes;
  const uint32_t e_smem = lse_smem + Cfg::kLseBytes;   // 1024-byte aligned (stages and G are multiples of 1024)
  uint8_t* bar_ptr = smem + STAGES * kStageBytes + Cfg::kGBytes + Cfg::kLseBytes + Cfg::kEBytes;
  constexpr bool kFwd = (MODE == MODE_FWD || MODE == MODE_FWDE || MODE == MODE_FWDEU);
  const uint32_t bar_base = smem_u32(bar_ptr);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + Cfg::kNumBars * 8);

  auto bar_full = [&](int s) { return bar_base + 8u * s; };
  auto bar_empty = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto bar_sfull = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto bar_sempty = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  auto bar_gfull = [&](int b) { return bar_base + 8u * (2 * STAGES + 4 + b); };
  auto bar_gempty = [&](int b) { return bar_base + 8u * (2 * STAGES + 6 + b); };
  const uint32_t bar_dafull = bar_base + 8u * (2 * STAGES + 8);
  const uint32_t bar_daempty = bar_base + 8u * (2 * STAGES + 9);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (MODE == MODE_BWD || Cfg::kEOut) tma_prefetch_desc(&tmBt);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(bar_full(s), 1);
        mbar_init(bar_empty(s), 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_sfull(b), 1);
        mbar_init(bar_sempty(b), kEpiWarps);
        mbar_init(bar_gfull(b), kEpiWarps);
        mbar_init(bar_gempty(b), 1);
      }
      mbar_init(bar_dafull, 1);
      mbar_init(bar_daempty, kEpiWarps);
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item -> (row block, D chunk, column-tile range)
  auto decode = [&](int item, int& rb, int& dc, int& chunk, int& t0, int& t1) {
    rb = item % p.num_rb;
    int rest = item / p.num_rb;
    if (MODE == MODE_BWD) {
      dc = rest % p.num_dc;
      rest /= p.num_dc;
    } else {
      dc = 0;
    }
    chunk = rest + p.chunk_rot;
    if (chunk >= p.num_chunks) chunk -= p.num_chunks;
    t0 = p.tile_begin + chunk * p.tiles_per_chunk;
    t1 = min(t0 + p.tiles_per_chunk, p.tile_end);
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    // (the whole warp walks the loop so control flow stays uniform; one elected lane issues)
    {
      uint32_t stage = 0, phase = 0;
      auto advance = [&]() {
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      unsigned long long src_ready = 0ull;     // sources whose rows are known to have landed
      const int want_epoch = p.sig_ready != nullptr ? ld_relaxed_gpu(p.sig_epoch) : 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        int rb, dc, chunk, t0, t1;
        decode(item, rb, dc, chunk, t0, t1);
        auto load_s = [&](int t) {
          if (p.sig_ready != nullptr) {
            const int q0 = (t * BN) / p.src_cols, q1 = min((t + 1) * BN - 1, p.n_cols - 1) / p.src_cols;
            for (int q = q0; q <= q1; ++q) {
              if (q == p.my_src || ((src_ready >> q) & 1ull)) continue;
              if (lane == 0) peer_wait_one(p.sig_ready, 0, q, want_epoch);
              __syncwarp();
              fence_proxy_async_all();     // the acquire (generic proxy) before the TMA reads (async proxy) of those rows
              src_ready |= 1ull << q;
            }
          }
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(bar_empty(stage), phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(bar_full(stage), kBM * kBK * 2 + BN * kBK * 2);
              const uint32_t dst = stage_base + stage * kStageBytes;
              tma_load_2d(dst, &tmA, bar_full(stage), kb * kBK, rb * kBM);
              tma_load_2d(dst + kBM * kBK * 2, &tmB, bar_full(stage), kb * kBK, t * BN);
            }
            __syncwarp();
            advance();
          }
        };
        auto load_d = [&](int t) {
          for (int kb2 = 0; kb2 < BN / kBK; ++kb2) {
            for (int sub = 0; sub < NSUB; ++sub) {
              mbar_wait(bar_empty(stage), phase ^ 1);
              if (elect_one()) {
                mbar_expect_tx(bar_full(stage), DN * kBK * 2);
                tma_load_2d(stage_base + stage * kStageBytes, &tmBt, bar_full(stage),
                            t * BN + kb2 * kBK, dc * DC + sub * DN);
              }
              __syncwarp();
              advance();
            }
          }
        };
        if (!Cfg::kHasDa) {
          for (int t = t0; t < t1; ++t) load_s(t);
        } else {
          load_s(t0);
          for (int t = t0; t < t1; ++t) {
            if (t + 1 < t1) load_s(t + 1);
            load_d(t);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // (warp-uniform loop; tcgen05.mma / commit are issued by one elected lane)
    {
      uint32_t stage = 0, phase = 0;
      auto advance = [&]() {
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      uint32_t s_use = 0, g_use = 0, item_count = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        int rb, dc, chunk, t0, t1;
        decode(item, rb, dc, chunk, t0, t1);
        auto mma_s = [&]() {
          const uint32_t buf = s_use % NSB, use = s_use / NSB;
          mbar_wait(bar_sempty(buf), (use & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(bar_full(stage), phase);
            tc_fence_after();
            const uint32_t a = stage_base + stage * kStageBytes;
            const uint64_t adesc = make_kmajor_sw128_desc(a);
            const uint64_t bdesc = make_kmajor_sw128_desc(a + kBM * kBK * 2);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                // +32 bytes per K=16 step inside the swizzle atom == +2 in the (addr >> 4) field
                umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC_S, (kb | k) != 0 ? 1u : 0u);
              }
              umma_commit(bar_empty(stage));
            }
            __syncwarp();
            advance();
          }
          if (elect_one()) umma_commit(bar_sfull(buf));
          __syncwarp();
          ++s_use;
        };
        auto mma_d = [&](bool first) {
          const uint32_t gb = g_use & 1, use = g_use >> 1;
          mbar_wait(bar_gfull(gb), use & 1);
          tc_fence_after();
          if (first) {
            mbar_wait(bar_daempty, (item_count & 1) ^ 1);
            tc_fence_after();
          }
          for (int kb2 = 0; kb2 < BN / kBK; ++kb2) {
            for (int sub = 0; sub < NSUB; ++sub) {
              mbar_wait(bar_full(stage), phase);
              tc_fence_after();
              const uint64_t adesc = make_kmajor_sw128_desc(g_base + gb * 32768 + kb2 * 16384);
              const uint64_t bdesc = make_kmajor_sw128_desc(stage_base + stage * kStageBytes);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                  umma_bf16(tmem_base + Cfg::kDaCol + sub * DN, adesc + 2 * k, bdesc + 2 * k, IDESC_D,
                            (first && kb2 == 0 && k == 0) ? 0u : 1u);
                }
                umma_commit(bar_empty(stage));
              }
              __syncwarp();
              advance();
            }
          }
          if (elect_one()) umma_commit(bar_gempty(gb));
          __syncwarp();
          ++g_use;
        };
        if (!Cfg::kHasDa) {
          for (int t = t0; t < t1; ++t) mma_s();
        } else {
          mma_s();
          for (int t = t0; t < t1; ++t) {
            if (t + 1 < t1) mma_s();
            mma_d(t == t0);
          }
          if (elect_one()) umma_commit(bar_dafull);
          __syncwarp();
          ++item_count;
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue warps
    const uint32_t q = warp & 3;          // TMEM lane quarter this warp may touch
    const uint32_t h = (warp - 2) >> 2;   // which 64-column half of the tile
    const uint32_t row_in_tile = q * 32 + lane;
    const float s = (MODE == MODE_RANK) ? 1.f : __ldg(p.scale);   // (ranks are scale invariant: no scale operand there)
    const float sl = s * kLog2e;
    float bias = 0.f;
    if (LOSS == LOSS_SIGLIP && p.bias != nullptr) bias = __ldg(p.bias);
    const float b2 = bias * kLog2e;
    uint32_t s_use = 0, g_use = 0, item_count = 0;

    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int rb, dc, chunk, t0, t1;
      decode(item, rb, dc, chunk, t0, t1);
      const int grow = rb * kBM + row_in_tile;
      const bool row_valid = grow < p.m_rows;
      const int label = grow + p.label_offset;
      const int label_w0 = rb * kBM + q * 32 + p.label_offset;  // label of lane 0 (warp uniform)

      float m_run = -CUDART_INF_F, l_run = 0.f;  // FWD clip: running row (max2, sum)
      float u_run = 0.f;                         // FWDEU: running sum of 2^(S2 - m_run) * S2
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;  // scalar partials (loss | ds, db)
      float lr2 = CUDART_INF_F;
      if (!kFwd && MODE != MODE_RANK && LOSS == LOSS_CLIP && row_valid) lr2 = __ldg(p.lse2_a + grow);
      // RANK: this row's class, list offset and (phase 1) the 32 positives of the current chunk, +inf padded
      int rk_cls = -2, rk_nbest = 0;
      long long rk_off = 0;
      unsigned long long rk_npairs = 0ull;
      float rk_lmax = CUDART_INF_F, rk_lmin = CUDART_INF_F;
      float rk_l[MODE == MODE_RANK ? 32 : 1];
      if (MODE == MODE_RANK && row_valid) {
        rk_cls = __ldg(p.rk_row_cls + grow);
        rk_off = __ldg(p.rk_off + grow);
        if (p.rk_phase == 1) {
          const int m = __ldg(p.rk_m + grow);
          rk_lmax = p.rk_chunk0 == 0 ? __ldg(p.rk_lmax + grow) : CUDART_INF_F;    // best counted once, in chunk 0
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            rk_l[t] = (p.rk_chunk0 + t < m) ? p.rk_pos[rk_off + p.rk_chunk0 + t] : CUDART_INF_F;
            rk_lmin = fminf(rk_lmin, rk_l[t]);
          }
        }
      }

      for (int t = t0; t < t1; ++t) {
        const uint32_t buf = s_use % NSB, use = s_use / NSB;
        float4 lbv = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t lse_w = lse_smem + (warp - 2) * (BN / 2) * 4;
        if (Cfg::kLseSmem)
          lbv = __ldg(reinterpret_cast<const float4*>(p.lse2_b + t * BN + h * (BN / 2)) + lane);
        mbar_wait(bar_sfull(buf), use & 1);
        tc_fence_after();
        if (Cfg::kLseSmem) {
          __syncwarp();   // every lane is done with the previous tile's strip
          sts_f4(lse_w + 16 * lane, lbv);
          __syncwarp();
        }
#pragma unroll 1
        for (int sb = 0; sb < BN / 128; ++sb) {   // this warp's BN/2 columns, 64 at a time
        uint32_t raw[64];
        {
          const uint32_t taddr = tmem_base + ((q * 32u) << 16) + buf * BN + h * (BN / 2) + sb * 64;
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr, r0);
          tmem_ld_32x32(taddr + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            raw[c] = r0[c];
            raw[c + 32] = r1[c];
          }
        }
        if (sb == BN / 128 - 1) {   // accumulator buffer fully drained into registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_sempty(buf));
        }

        const int col_base = t * BN + h * (BN / 2) + sb * 64;
        const bool ragged = (col_base + 64 > p.n_cols);  // warp uniform
        const bool diag_here = (label_w0 + 31 >= col_base) && (label_w0 < col_base + 64);

        // FWDE: pack 64 fp32 values of this thread's row to bf16, stage them (swizzled) and TMA-store the warp's
        // 32 x 64 tile to e_out[rb*128 + q*32 .., col_base ..]  (tmBt is the store map in this mode)
        auto store_e = [&](const float (&e)[64]) {
          const uint32_t stg = e_smem + (warp - 2) * 4096;
          if (lane == 0) bulk_wait_read0();   // the previous tile of this warp has left the staging buffer
          __syncwarp();
#pragma unroll
          for (int ck = 0; ck < 8; ++ck) {
            uint4 o;
            o.x = pack_bf16x2(e[ck * 8 + 0], e[ck * 8 + 1]);
            o.y = pack_bf16x2(e[ck * 8 + 2], e[ck * 8 + 3]);
            o.z = pack_bf16x2(e[ck * 8 + 4], e[ck * 8 + 5]);
            o.w = pack_bf16x2(e[ck * 8 + 6], e[ck * 8 + 7]);
            sts_u4(stg + lane * 128 + ((static_cast<uint32_t>(ck) ^ (lane & 7)) << 4), o);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmBt, stg, col_base, rb * kBM + q * 32);
            bulk_commit();
          }
        };
        if (Cfg::kEOut && LOSS == LOSS_CLIP) {
          // ---- forward that keeps E: the lean epilogue (sl > 0: logit_scale is exp(.) in the reference, model.py:324)
          //   * the sub-tile reference is taken on the raw accumulators (max commutes with the positive scale), so
          //     the scaling folds into the exponent's FMA, issued as packed fp32 pairs (FFMA2 / FADD2);
          //   * the column sums go through the warp's staging tile in fp32 (two 32-column halves, swizzled, conflict
          //     free: 16 STS.128 + 64 LDS + 64 FADD per lane) instead of a 5-step shuffle butterfly (~250
          //     instructions); they must stay fp32 -- summing the bf16 E instead loses the column LSE of a confident
          //     model (dominant term rounded to 8 bits) and with it the loss.
          float a[64];
#pragma unroll
          for (int c = 0; c < 64; ++c) a[c] = __uint_as_float(raw[c]);
          if (diag_here && row_valid) {   // select chain, not an indexed read: keeps a[] in registers
            const int idx = label - col_base;
            float dv = 0.f;
#pragma unroll
            for (int c = 0; c < 64; ++c) dv = (c == idx) ? a[c] : dv;
            if (idx >= 0 && idx < 64) p.diag2[grow] = dv * sl;
          }
          // masked entries: -inf, or a finite sentinel when the exponent is also multiplied (0 * -inf is NaN)
          const float kMasked = Cfg::kRowEnt ? -1e30f : -CUDART_INF_F;
          if (ragged) {
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (col_base + c >= p.n_cols) a[c] = kMasked;
          }
          if (!row_valid) {
#pragma unroll
            for (int c = 0; c < 64; ++c) a[c] = kMasked;
          }
          // (four independent chains each: a 64-deep dependent max / add chain is ~250 cycles of pure latency per
          //  sub-tile with only two epilogue warps per scheduler to hide it)
          float tm[4] = {a[0], a[1], a[2], a[3]};
#pragma unroll
          for (int c = 4; c < 64; c += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tm[k] = fmaxf(tm[k], a[c + k]);
          }
          const float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3]));
          const float wmax = warp_max(tmax);
          const float cw = (Cfg::kRowEnt ? (wmax <= -1e29f) : (wmax == -CUDART_INF_F)) ? 0.f : wmax * sl;
          const float2 sl2 = make_float2(sl, sl), ncw2 = make_float2(-cw, -cw);
          float2 rs4[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
          float2 us2 = make_float2(0.f, 0.f);
          float v[64];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float2 t = ffma2(make_float2(a[2 * c], a[2 * c + 1]), sl2, ncw2);
            v[2 * c] = ex2f(t.x);
            v[2 * c + 1] = ex2f(t.y);
            rs4[c & 3] = fadd2(rs4[c & 3], make_float2(v[2 * c], v[2 * c + 1]));
            if (Cfg::kRowEnt) us2 = ffma2(make_float2(v[2 * c], v[2 * c + 1]), t, us2);   // sum e * (S2 - cw)
          }
          const float2 rs2 = fadd2(fadd2(rs4[0], rs4[1]), fadd2(rs4[2], rs4[3]));
          const float rowsum = rs2.x + rs2.y;
          const float mnew = fmaxf(m_run, cw);
          if (Cfg::kRowEnt) {
            const float f_old = ex2f(m_run - mnew), f_new = ex2f(cw - mnew);
            l_run = l_run * f_old + rowsum * f_new;
            u_run = u_run * f_old + fmaf(cw, rowsum, us2.x + us2.y) * f_new;
          } else {
            l_run = l_run * ex2f(m_run - mnew) + rowsum * ex2f(cw - mnew);
          }
          m_run = mnew;
          // column sums: lane = row writes 32 fp32 columns (8 swizzled 16-byte chunks), then lane = column reads them
          const uint32_t stg = e_smem + (warp - 2) * 4096;
          if (lane == 0) bulk_wait_read0();   // the previous E tile of this warp has left the staging buffer
          float cs[2];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            __syncwarp();
#pragma unroll
            for (int ck = 0; ck < 8; ++ck)
              sts_f4(stg + lane * 128 + ((static_cast<uint32_t>(ck) ^ (lane & 7)) << 4),
                     make_float4(v[hf * 32 + ck * 4], v[hf * 32 + ck * 4 + 1], v[hf * 32 + ck * 4 + 2],
                                 v[hf * 32 + ck * 4 + 3]));
            __syncwarp();
            float ld[32];                          // all 32 loads issued before the first add consumes one
#pragma unroll
            for (int r8 = 0; r8 < 8; ++r8) {
              const uint32_t base = stg + r8 * 128 + ((((lane >> 2) ^ r8)) << 4) + (lane & 3) * 4;
#pragma unroll
              for (int r = 0; r < 4; ++r) ld[r8 * 4 + r] = __uint_as_float(lds_u32(base + r * 1024));
            }
            float acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = (ld[k] + ld[k + 8]) + (ld[k + 16] + ld[k + 24]);
            cs[hf] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
          }
          store_e(v);   // bf16 E through the same staging tile (its __syncwarp orders it after the column reads)
          const int band = rb * 4 + q;
          p.col_l[(size_t)band * p.n_pad + col_base + lane] = cs[0];
          p.col_l[(size_t)band * p.n_pad + col_base + 32 + lane] = cs[1];
          if (lane == 0) p.col_c[(size_t)band * (p.n_pad / 64) + col_base / 64] = cw;
        } else if (MODE == MODE_RANK) {
          // ---- rank-of-label epilogue (the logit scale is positive, so ranks are taken on the raw cosines)
          const int my_cls = row_valid ? rk_cls : -2;
          if (p.rk_phase == 0) {
#pragma unroll
            for (int c = 0; c < 64; ++c) {
              const int j = col_base + c;
              if (__ldg(p.rk_col_cls + j) == my_cls) p.rk_pos[rk_off + __ldg(p.rk_col_ord + j)] = __uint_as_float(raw[c]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 64; ++c) {
              const int cc = __ldg(p.rk_col_cls + col_base + c);
              const float a = __uint_as_float(raw[c]);
              if (cc < 0 || cc == my_cls || !row_valid) continue;       // padding, positive, dead row
              if (a > rk_lmax) ++rk_nbest;
              if (a > rk_lmin) {
                int cnt = 0;
#pragma unroll
                for (int t = 0; t < 32; ++t) cnt += (a > rk_l[t]) ? 1 : 0;
                rk_npairs += cnt;
              }
            }
          }
        } else if (kFwd && LOSS == LOSS_CLIP) {
          float v[64];
#pragma unroll
          for (int c = 0; c < 64; ++c) v[c] = __uint_as_float(raw[c]) * sl;
          if (diag_here && row_valid) {
            const int idx = label - col_base;
            float dv = 0.f;
#pragma unroll
            for (int c = 0; c < 64; ++c) dv = (c == idx) ? v[c] : dv;
            if (idx >= 0 && idx < 64) p.diag2[grow] = dv;
          }
          if (ragged) {
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (col_base + c >= p.n_cols) v[c] = -CUDART_INF_F;
          }
          if (!row_valid) {
#pragma unroll
            for (int c = 0; c < 64; ++c) v[c] = -CUDART_INF_F;
          }
          float tmax = v[0];
#pragma unroll
          for (int c = 1; c < 64; ++c) tmax = fmaxf(tmax, v[c]);
          float cw = warp_max(tmax);
          if (cw == -CUDART_INF_F) cw = 0.f;
          float rowsum = 0.f;
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            v[c] = ex2f(v[c] - cw);
            rowsum += v[c];
          }
          if (Cfg::kEOut) store_e(v);
          const float mnew = fmaxf(m_run, cw);
          l_run = l_run * ex2f(m_run - mnew) + rowsum * ex2f(cw - mnew);
          m_run = mnew;
          butterfly_step<64, 16>(v, lane);
          butterfly_step<32, 8>(v, lane);
          butterfly_step<16, 4>(v, lane);
          butterfly_step<8, 2>(v, lane);
          butterfly_step<4, 1>(v, lane);
          const int band = rb * 4 + q;
          *reinterpret_cast<float2*>(p.col_l + (size_t)band * p.n_pad + col_base + 2 * lane) =
              make_float2(v[0], v[1]);
          if (lane == 0) p.col_c[(size_t)band * (p.n_pad / 64) + col_base / 64] = cw;
        } else if (kFwd && LOSS == LOSS_SIGLIP) {
          // softplus(-y z) = max(z, 0) + log1p(e) - [positive] z  and  sigmoid(z) = (z >= 0 ? 1 : e) / (1 + e)  with
          // e = 2^-|u|: one ex2 per logit.  When the whole 32 x 64 sub-tile has e <= 1/8 (|z| >= 2.08: the normal
          // case, SigLIP's bias starts at -10) log1p and the reciprocal are short series on the FMA pipe; the exact
          // lg2 / rcp forms (three MUFU per logit, as much XU time as the tile's MMAs) are kept for the rest.
          float part = 0.f;
          float gv[64];             // e, then overwritten by G
          float dsum = 0.f, bsum = 0.f;
          float emax = 0.f;
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            const float u = fmaf(__uint_as_float(raw[c]), sl, b2);
            gv[c] = ex2f(-fabsf(u));
            emax = fmaxf(emax, gv[c]);
          }
          const bool small = __all_sync(0xffffffffu, emax <= 0.125f);
          if (small) {
#pragma unroll
            for (int c = 0; c < 64; ++c) {
              const float a = __uint_as_float(raw[c]);
              const float z = fmaf(a, s, bias);
              const float e = gv[c];
              const float lp = e * (1.f + e * (-0.5f + e * (0.33333334f + e * (-0.25f + e * 0.2f))));
              const float r = 1.f + e * (-1.f + e * (1.f + e * (-1.f + e)));     // 1/(1+e), error < e^5
              float sp = fmaxf(z, 0.f) + lp;
              float g = (z >= 0.f ? 1.f : e) * r;
              const bool dead = (ragged && (col_base + c >= p.n_cols)) || !row_valid;
              if (dead) {
                sp = 0.f;
                g = 0.f;
              }
              part += sp;
              if (Cfg::kEOut) {
                gv[c] = g;
                dsum = fmaf(g, a, dsum);
                bsum += g;
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < 64; ++c) {
              const float a = __uint_as_float(raw[c]);
              const float z = fmaf(a, s, bias);
              const float e = gv[c];
              float sp = fmaxf(z, 0.f) + log1p_from_exp(e);
              float g = (z >= 0.f ? 1.f : e) * rcpf(1.f + e);   // sigmoid(z)
              const bool dead = (ragged && (col_base + c >= p.n_cols)) || !row_valid;
              if (dead) {
                sp = 0.f;
                g = 0.f;
              }
              part += sp;
              if (Cfg::kEOut) {
                gv[c] = g;
                dsum = fmaf(g, a, dsum);
                bsum += g;
              }
            }
          }
          if (diag_here) {   // the positive of this row, if it falls into these 64 columns (select chains only)
            const int idx = label - col_base;
            const bool has = row_valid && idx >= 0 && idx < 64;
            float ad = 0.f;
#pragma unroll
            for (int c = 0; c < 64; ++c) {
              const bool here = has && (c == idx);
              ad = here ? __uint_as_float(raw[c]) : ad;
              if (Cfg::kEOut) gv[c] = here ? gv[c] - 1.f : gv[c];
            }
            if (has) {
              part -= fmaf(ad, s, bias);
              if (Cfg::kEOut) {
                dsum -= ad;
                bsum -= 1.f;
              }
            }
          }
          acc0 += part;
          if (Cfg::kEOut) {
            store_e(gv);
            acc1 += dsum;
            acc2 += bsum;
          }
        } else {
          // ------------------------------------------------------------- BWD / GW: build G
          const uint32_t gb = g_use & 1, guse = g_use >> 1;
          if (MODE == MODE_BWD) mbar_wait(bar_gempty(gb), (guse & 1) ^ 1);
          const uint32_t grow_smem =
              g_base + gb * 32768 + h * 16384 + row_in_tile * 128;
          uint16_t* grow_gmem = (MODE == MODE_GW) ? p.g_out + (size_t)grow * p.g_ld + col_base : nullptr;
          float dsum = 0.f, bsum = 0.f;
#pragma unroll
          for (int ck = 0; ck < 8; ++ck) {
            float g[8];
            if (LOSS == LOSS_CLIP) {
              float4 lb0, lb1;
              if (Cfg::kLseSmem) {
                lb0 = lds_f4(lse_w + (sb * 64 + ck * 8) * 4);
                lb1 = lds_f4(lse_w + (sb * 64 + ck * 8) * 4 + 16);
              } else {
                lb0 = __ldg(reinterpret_cast<const float4*>(p.lse2_b + col_base + ck * 8));
                lb1 = __ldg(reinterpret_cast<const float4*>(p.lse2_b + col_base + ck * 8 + 4));
              }
              const float lb[8] = {lb0.x, lb0.y, lb0.z, lb0.w, lb1.x, lb1.y, lb1.z, lb1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float a = __uint_as_float(raw[ck * 8 + j]);
                const float tt = a * sl;
                const float p_own = ex2f(tt - lr2);
                const float p_oth = ex2f(tt - lb[j]);
                g[j] = p.w_own * p_own + p.w_oth * p_oth;
                if (MODE == MODE_GW && p.ent) {   // negative entropies (lse = +inf on dead rows / columns: P = 0)
                  dsum = fmaf(p_own, p_own > 0.f ? tt - lr2 : 0.f, dsum);
                  bsum = fmaf(p_oth, p_oth > 0.f ? tt - lb[j] : 0.f, bsum);
                } else {
                  dsum = fmaf(p_own, a, dsum);
                  if (MODE == MODE_GW) bsum = fmaf(p_oth, a, bsum);   // other-direction term of d_scale
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float a = __uint_as_float(raw[ck * 8 + j]);
                const float u = fmaf(a, sl, b2);
                float gg = rcpf(1.f + ex2f(-u));
                if (!row_valid) gg = 0.f;
                g[j] = gg;
              }
            }
            if (ragged) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (col_base + ck * 8 + j >= p.n_cols) g[j] = 0.f;
            }
            if (LOSS == LOSS_SIGLIP) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                dsum = fmaf(g[j], __uint_as_float(raw[ck * 8 + j]), dsum);
                bsum += g[j];
              }
            }
            if (diag_here) {
              const int idx = label - col_base - ck * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (j == idx && row_valid) {
                  const float a = __uint_as_float(raw[ck * 8 + j]);
                  if (LOSS == LOSS_CLIP) {
                    g[j] -= (p.w_own + p.w_oth);
                    if (!(MODE == MODE_GW && p.ent)) {
                      dsum -= a;
                      if (MODE == MODE_GW) bsum -= a;
                    }
                  } else {
                    g[j] -= 1.f;
                    dsum -= a;
                    bsum -= 1.f;
                  }
                }
              }
            }
            if (MODE == MODE_BWD) {
              const uint32_t dst = grow_smem + ((static_cast<uint32_t>(ck) ^ (row_in_tile & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                           "r"(pack_bf16x2(g[0], g[1])), "r"(pack_bf16x2(g[2], g[3])),
                           "r"(pack_bf16x2(g[4], g[5])), "r"(pack_bf16x2(g[6], g[7]))
                           : "memory");
            } else if (row_valid) {
              uint4 o;
              o.x = pack_bf16x2(g[0], g[1]);
              o.y = pack_bf16x2(g[2], g[3]);
              o.z = pack_bf16x2(g[4], g[5]);
              o.w = pack_bf16x2(g[6], g[7]);
              *reinterpret_cast<uint4*>(grow_gmem + ck * 8) = o;
            }
          }
          if (MODE == MODE_BWD) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_gfull(gb));
          }
          ++g_use;
          acc0 += dsum;
          acc1 += bsum;
        }
        }  // sub-blocks
        ++s_use;
      }  // tiles

      // ------------------------------------------------------------------ item outputs
      if (MODE == MODE_RANK) {
        if (p.rk_phase == 1 && row_valid) {
          if (rk_npairs) atomicAdd(p.rk_pairs + grow, rk_npairs);
          if (rk_nbest) atomicAdd(p.rk_best + grow, rk_nbest);
        }
      } else if (kFwd && LOSS == LOSS_CLIP) {
        const int slot = (p.chunk_base + chunk) * 2 + h;
        p.row_part[(size_t)slot * p.m_pad + grow] = make_float2(m_run, l_run);
        if (Cfg::kRowEnt) p.row_ent[(size_t)slot * p.m_pad + grow] = u_run;
      } else if (kFwd) {
        const float tot = warp_sum(acc0);
        if (lane == 0) p.sc_part[(size_t)item * kEpiWarps + (warp - 2)] = make_float2(tot, 0.f);
        if (Cfg::kEOut) {   // SigLIP: the d_scale / d_bias partials ride along (second float2 plane)
          const float t1s = warp_sum(acc1);
          const float t2s = warp_sum(acc2);
          if (lane == 0)
            p.sc_part2[(size_t)item * kEpiWarps + (warp - 2)] = make_float2(t1s, t2s);
        }
      } else if (MODE == MODE_GW) {
        const float t0s = warp_sum(acc0);
        const float t1s = warp_sum(acc1);
        if (lane == 0) p.sc_part[(size_t)item * kEpiWarps + (warp - 2)] = make_float2(t0s, t1s);
      } else {
        mbar_wait(bar_dafull, item_count & 1);
        tc_fence_after();
        float* out_row = p.dpart + ((size_t)chunk * p.m_pad + grow) * p.d_pad + dc * DC;
#pragma unroll 1
        for (int c0 = h * (DC / 2); c0 < (int)(h + 1) * (DC / 2); c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + Cfg::kDaCol + c0, r);
          tmem_ld_wait();
          if (row_valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o;
              o.x = __uint_as_float(r[4 * j + 0]);
              o.y = __uint_as_float(r[4 * j + 1]);
              o.z = __uint_as_float(r[4 * j + 2]);
              o.w = __uint_as_float(r[4 * j + 3]);
              *reinterpret_cast<float4*>(out_row + c0 + 4 * j) = o;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_daempty);
        ++item_count;
        if (dc != 0) {
          acc0 = 0.f;
          acc1 = 0.f;
        }
        const float t0s = warp_sum(acc0);
        const float t1s = warp_sum(acc1);
        if (lane == 0) p.sc_part[(size_t)item * kEpiWarps + (warp - 2)] = make_float2(t0s, t1s);
      }
    }  // items
    if (Cfg::kEOut && lane == 0) bulk_wait_read0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mrclip
