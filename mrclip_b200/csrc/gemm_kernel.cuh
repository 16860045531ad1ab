// Gradient contraction  dA[m, d] = sum_k G[m, k] * F[k, d]  as a persistent warp-specialised
// tcgen05 GEMM (128 x 256 output tiles, K streamed in 64-wide blocks through a TMA ring).
// G is the bf16 gradient-of-logits block produced by tile_kernel<MODE_GW>; F is the other
// modality's packed feature matrix, row-major [K, D].  As a UMMA B operand (N = d, K = k) that is
// an "MN-major" tile: TMA boxes of [64 k-rows x 64 d-cols] land as 128-byte swizzled rows and the
// descriptor says N-major, so no transposed copy of the features is ever made.  Split-K partials
// go to dpart[ks][m_pad][d_pad] and are summed (and scaled / cast) by grad_reduce_kernel.
//
// With A_MN (transposed-A variant) the same G block is contracted along its rows instead:
//   dB[k_col, d] = sum_m G[m, k_col] * F'[m, d]   -- A is then an M-major UMMA operand read from
// the very same row-major G through a different TMA box / descriptor; this is what lets one G
// block serve both gradients.
//
// Replaces the autograd matmul-backward GEMMs of loss.py:117-124.
#pragma once
#include "aux_kernels.cuh"
#include "peer_sync.cuh"
#include "ptx.cuh"
#include "tile_kernel.cuh"

namespace mrclip {

constexpr int kGemmBN = 256;
constexpr int kGemmStageBytes = 16384 + 32768;
constexpr int kGemmStages = 4;
constexpr int kGemmBars = 2 * kGemmStages + 4;
constexpr int kGemmSmemBytes = kGemmStages * kGemmStageBytes + kGemmBars * 8 + 16 + 1024;

struct GemmParams {
  int m_rows;        // output rows (valid)
  int num_rb;        // ceil(m_rows / 128)
  int num_dt;        // ceil(d / 256)
  int num_kb;        // total K blocks of 64
  int kb_per_split;  // K blocks per split
  int ksplit;
  int num_items;
  int m_pad, d_pad;
  float* dpart;      // [ksplit][m_pad][d_pad]
  // ksplit == 1: the epilogue scales and stores the result itself (no partials, no grad_reduce pass)
  void* out;         // [m_rows][out_ld] of out_dtype (0 f32, 1 bf16, 2 f16), NULL -> write dpart
  long out_ld;
  int out_dtype, d_valid;
  float coef;
  const float* scale;
  const float* grad_out;   // may be NULL
  // fused reduce-scatter ("push"): output row block q (peer_n rows) goes straight into rank q's receive buffer,
  // slot peer_rank, through the NVLink-mapped pointer peer[q]: fp32 [ranks][peer_n][d_valid] on every rank
  const unsigned long long* peer;
  int peer_n, peer_rank;
  // optional: raise CH_DTEXT on every rank when the last CTA's pushed rows have landed (peer_sync.cuh); sig.ranks == 0: off
  PeerInfo sig;
  // stream-K (gemm2_kernel<.., SK>): the tiles' K blocks form one list of num_rb*num_dt*num_kb units that is cut into
  // equal contiguous ranges, one per CTA pair.  A pair whose range starts inside a tile writes that partial accumulator
  // to sk_part[pair] first thing and bumps sk_flags[tile]; the pair whose range ends inside a tile owns it: it
  // accumulates the tile's first K blocks last, adds the other pairs' partials and runs the epilogue.  No split-K
  // partials of whole matrices, no reduce pass, no wave quantisation.
  float* sk_part;     // [pairs][2 CTAs][128 rows][256 cols] fp32
  int* sk_flags;      // [num_rb * num_dt], zero at launch
};

// destination row of the direct epilogue: local output, or the owner's receive slot over NVLink
__device__ __forceinline__ float* gemm_out_row_f32(const GemmParams& p, int grow) {
  if (p.peer == nullptr) return reinterpret_cast<float*>(p.out) + (size_t)grow * p.out_ld;
  const int q = grow / p.peer_n, lrow = grow - q * p.peer_n;
  return reinterpret_cast<float*>(__ldg(p.peer + q)) + ((size_t)p.peer_rank * p.peer_n + lrow) * p.out_ld;
}

// MN-major (M contiguous) A operand tile: K rows of 128 bytes (64 M-elements), 128B swizzle.
// Two 64-wide M chunks per 128-row A tile, LBO bytes apart; 8 K-rows per swizzle atom (SBO=1024).
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <bool A_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmParams p) {
  constexpr int STAGES = kGemmStages;
  // instruction descriptor: bit 15 = A is MN-major
  constexpr uint32_t IDESC = make_idesc_bf16(kBM, kGemmBN) | (A_MN ? (1u << 15) : 0u) | (1u << 16);  // B is N-major

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t stage_base = smem_u32(smem);
  uint8_t* bar_ptr = smem + STAGES * kGemmStageBytes;
  const uint32_t bar_base = smem_u32(bar_ptr);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + kGemmBars * 8);
  auto bar_full = [&](int s) { return bar_base + 8u * s; };
  auto bar_empty = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto bar_accfull = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto bar_accempty = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(bar_full(s), 1);
        mbar_init(bar_empty(s), 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_accfull(b), 1);
        mbar_init(bar_accempty(b), kEpiWarps);
      }
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& rb, int& dt, int& ks, int& kb0, int& kb1) {
    // dt fastest: the CTAs that share one A (= G) block run side by side, so G streams from HBM once
    // and the other num_dt - 1 readers hit it in L2 (profiles/r1_notes.md: 6.6 GB -> ~2.3 GB per launch)
    dt = item % p.num_dt;
    int rest = item / p.num_dt;
    rb = rest % p.num_rb;
    ks = rest / p.num_rb;
    kb0 = ks * p.kb_per_split;
    kb1 = min(kb0 + p.kb_per_split, p.num_kb);
  };

  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int rb, dt, ks, kb0, kb1;
      decode(item, rb, dt, ks, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_empty(stage), phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(bar_full(stage), kGemmStageBytes);
          const uint32_t dst = stage_base + stage * kGemmStageBytes;
          if (A_MN) {
            // G rows kb*64.. (K), columns rb*128.. (M): two boxes of [64 K-rows x 64 M-cols]
            tma_load_2d(dst, &tmA, bar_full(stage), rb * kBM, kb * kBK);
            tma_load_2d(dst + 8192, &tmA, bar_full(stage), rb * kBM + 64, kb * kBK);
          } else {
            tma_load_2d(dst, &tmA, bar_full(stage), kb * kBK, rb * kBM);
          }
          // F rows kb*64.. (K), columns dt*256.. (N): four boxes of [64 K-rows x 64 N-cols]
#pragma unroll
          for (int c = 0; c < kGemmBN / 64; ++c)
            tma_load_2d(dst + 16384 + c * 8192, &tmB, bar_full(stage), dt * kGemmBN + c * 64, kb * kBK);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0, acc_use = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int rb, dt, ks, kb0, kb1;
      decode(item, rb, dt, ks, kb0, kb1);
      const uint32_t buf = acc_use & 1, use = acc_use >> 1;
      mbar_wait(bar_accempty(buf), (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * kGemmBN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_full(stage), phase);
        tc_fence_after();
        const uint32_t a = stage_base + stage * kGemmStageBytes;
        const uint64_t bdesc = make_mnmajor_sw128_desc(a + 16384, 8192);
        if (A_MN) {
          const uint64_t adesc = make_mnmajor_sw128_desc(a, 8192);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)  // 16 K-rows = 2 swizzle atoms = 2048 bytes per step
              umma_bf16(d_tmem, adesc + 128 * k, bdesc + 128 * k, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit(bar_empty(stage));
          }
        } else {
          const uint64_t adesc = make_kmajor_sw128_desc(a);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 128 * k, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit(bar_empty(stage));
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(bar_accfull(buf));
      __syncwarp();
      ++acc_use;
    }
  } else {
    const uint32_t q = warp & 3;
    const uint32_t h = (warp - 2) >> 2;
    const uint32_t row_in_tile = q * 32 + lane;
    uint32_t acc_use = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int rb, dt, ks, kb0, kb1;
      decode(item, rb, dt, ks, kb0, kb1);
      const uint32_t buf = acc_use & 1, use = acc_use >> 1;
      mbar_wait(bar_accfull(buf), use & 1);
      tc_fence_after();
      const int grow = rb * kBM + row_in_tile;
      float* out_row = p.dpart + ((size_t)ks * p.m_pad + grow) * p.d_pad + dt * kGemmBN;
      float mul = 1.f;
      if (p.out != nullptr) {
        mul = p.coef * __ldg(p.scale);
        if (p.grad_out != nullptr) mul *= __ldg(p.grad_out);
      }
#pragma unroll 1
      for (int c0 = h * 128; c0 < (int)(h + 1) * 128; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + buf * kGemmBN + c0, r);
        tmem_ld_wait();
        if (p.out != nullptr) {
          // direct epilogue: scale, cast, store (row-wise 16/32-byte pieces; the tile's columns may be ragged)
          if (grow < p.m_rows) {
            const int col = dt * kGemmBN + c0;
            if (p.out_dtype == DT_F32) {
              float* o = gemm_out_row_f32(p, grow) + col;
              if (col + 32 <= p.d_valid && (p.out_ld & 3) == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *reinterpret_cast<float4*>(o + 4 * j) =
                      make_float4(__uint_as_float(r[4 * j]) * mul, __uint_as_float(r[4 * j + 1]) * mul,
                                  __uint_as_float(r[4 * j + 2]) * mul, __uint_as_float(r[4 * j + 3]) * mul);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col + j < p.d_valid) o[j] = __uint_as_float(r[j]) * mul;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < p.d_valid)
                  store_from_float(p.out, p.out_dtype, (size_t)grow * p.out_ld + col + j, __uint_as_float(r[j]) * mul);
            }
          }
        } else if (grow < p.m_rows) {
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
      if (lane == 0) mbar_arrive(bar_accempty(buf));
      ++acc_use;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace mrclip
