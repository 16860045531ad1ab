// Gradient contraction on CTA pairs: the gemm_kernel of gemm_kernel.cuh with tcgen05 cta_group::2.
//
// Two CTAs of a cluster (the two SMs of a TPC) own one 256 x 256 output tile: CTA c holds rows c*128.. of A and
// of the accumulator (its own TMEM) and loads only HALF of the B tile (N columns c*128..); the tensor cores of
// both SMs read both halves.  Per 64-wide K block a CTA therefore pulls 16 KB (A) + 16 KB (half B) through L2
// instead of 16 + 32 KB: operand delivery was what bounded the single-CTA kernel, and the
// smaller stage lets the TMA ring hold 6 K blocks instead of 4.
//
// Protocol (one MMA issuer for the pair, in the leader CTA = cluster rank 0):
//   * both producers load into their own shared memory but credit the bytes to the LEADER's full barrier
//     (cp.async.bulk.tensor ... .cta_group::2, barrier address mapped with mapa);
//   * tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to the barriers of both CTAs;
//   * both epilogues drain their own TMEM half and arrive on the leader's "accumulator free" barrier.
#pragma once
#include "gemm_kernel.cuh"

namespace mrclip {

constexpr int kG2StageBytes = 16384 + 16384;
constexpr int kG2Bars = 2 * 6 + 4;
// PUSH (fused reduce-scatter): one ring stage is traded for a 32 x 32 fp32 staging tile per epilogue warp (rows
// padded to 144 B against bank conflicts), from which every lane sends its row as one 128-byte bulk copy -- peer
// memory over NVLink wants full-line packets, 16-byte stores from the registers reached a fraction of the link rate
// PUSH = 2 (MRCLIP_PUSH_DTYPE=bf16, not validated on hardware yet): the same with a bf16 payload -- two 32-column
// accumulator chunks per 128-byte row piece, half the NVLink bytes of the text-gradient exchange.
constexpr int kG2PushRowBytes = 144;
template <int PUSH> struct G2Cfg {
  static constexpr int kStages = PUSH ? 5 : 6;
  static constexpr int kStagingBytes = PUSH ? kEpiWarps * 32 * kG2PushRowBytes : 0;
  static constexpr int kSmemBytes = kStages * kG2StageBytes + kStagingBytes + kG2Bars * 8 + 16 + 1024;
};

// GemmParams::num_rb counts 256-row pair blocks here.
//
// SK (stream-K, local-output GEMMs): whole waves of tiles run as before (pair p: tiles p, p + P, ...); the last < P
// tiles, which would otherwise occupy a full wave (or force split-K partials of the whole matrix and a reduce pass),
// are cut at K-block granularity: their (tile, K block) units form one list split into equal contiguous ranges, one per
// CTA pair (GemmParams::sk_*).  A range that starts inside a tile yields a CONTRIBUTOR segment: the raw fp32
// accumulator goes to sk_part[pair] and sk_flags[tile] is bumped (release).  A range that ends inside a tile yields the
// tile's OWNER segment: its epilogue waits (acquire) until every contributor of the tile has arrived, adds their
// partials and stores the scaled result.  Within a pair the contributor segment always precedes the owner segment, and
// contributors never wait, so the scheme cannot deadlock (all pairs are resident at once: grid <= SM pairs).
template <bool A_MN, int PUSH, bool SK = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const GemmParams p) {
  static_assert(!(SK && PUSH != 0), "stream-K serves the local-output GEMMs");
  constexpr int STAGES = G2Cfg<PUSH>::kStages;
  // M = 256 across the pair; bit 15 = A is MN-major, bit 16 = B is MN-major
  constexpr uint32_t IDESC = make_idesc_bf16(2 * kBM, kGemmBN) | (A_MN ? (1u << 15) : 0u) | (1u << 16);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t stage_base = smem_u32(smem);
  const uint32_t staging_base = stage_base + STAGES * kG2StageBytes;
  uint8_t* bar_ptr = smem + STAGES * kG2StageBytes + G2Cfg<PUSH>::kStagingBytes;
  const uint32_t bar_base = smem_u32(bar_ptr);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + kG2Bars * 8);
  auto bar_full = [&](int s) { return bar_base + 8u * s; };
  auto bar_empty = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto bar_accfull = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto bar_accempty = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

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
        mbar_init(bar_accempty(b), 2 * kEpiWarps);   // the epilogue warps of both CTAs
      }
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc_pair(smem_u32(tmem_slot), 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's barriers exist before anything is credited to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& rb, int& dt, int& ks, int& kb0, int& kb1) {
    dt = item % p.num_dt;   // dt fastest: the pairs that share one A block run side by side (L2 reuse of G)
    int rest = item / p.num_dt;
    rb = rest % p.num_rb;
    ks = rest / p.num_rb;
    kb0 = ks * p.kb_per_split;
    kb1 = min(kb0 + p.kb_per_split, p.num_kb);
  };

  // stream-K, hybrid: the first full_waves * num_pairs tiles run as whole tiles, pair p taking tiles p, p + P, ...
  // (dt fastest, so the pairs that share a G block work on it side by side and it streams from HBM once); only the
  // remaining < P tiles are cut into per-pair ranges of (tile, K block) units.  The ranges come FIRST, so that the
  // owners' short wait for their contributors overlaps the MMAs of the following whole tile.
  const int sk_tiles = p.num_rb * p.num_dt;
  const int sk_full_waves = SK ? sk_tiles / num_pairs : 0;
  const int sk_base_tile = sk_full_waves * num_pairs;
  const long long sk_total = (long long)(sk_tiles - sk_base_tile) * p.num_kb;
  const long long sk_u0 = SK ? (long long)pair * sk_total / num_pairs : 0;
  const long long sk_u1 = SK ? (long long)(pair + 1) * sk_total / num_pairs : 0;
  auto next_segment = [&](long long& u, int& rb, int& dt, int& kb0, int& kb1, int& tile) {
    const int rt = (int)(u / p.num_kb);
    tile = sk_base_tile + rt;
    kb0 = (int)(u - (long long)rt * p.num_kb);
    const long long left = sk_u1 - u;
    kb1 = (left < (long long)(p.num_kb - kb0)) ? kb0 + (int)left : p.num_kb;
    rb = tile / p.num_dt;
    dt = tile - rb * p.num_dt;
    u += kb1 - kb0;
  };
  // one loop header for both schemes; the body starts with MRCLIP_G2_NEXT_ITEM, which declares rb, dt, ks, kb0, kb1, tile
#define MRCLIP_G2_FOR_ITEMS                                                                                    \
  for (long long it_ = SK ? sk_u0 : (long long)pair, fw_ = 0;                                                  \
       SK ? (it_ < sk_u1 || fw_ < sk_full_waves) : (it_ < (long long)p.num_items);)
#define MRCLIP_G2_NEXT_ITEM                                           \
  int rb, dt, ks = 0, kb0, kb1, tile = 0;                             \
  if (SK) {                                                           \
    if (it_ < sk_u1) {                                                \
      next_segment(it_, rb, dt, kb0, kb1, tile);                      \
    } else {                                                          \
      tile = pair + (int)fw_ * num_pairs;                             \
      rb = tile / p.num_dt;                                           \
      dt = tile - rb * p.num_dt;                                      \
      kb0 = 0;                                                        \
      kb1 = p.num_kb;                                                 \
      ++fw_;                                                          \
    }                                                                 \
  } else {                                                            \
    decode((int)it_, rb, dt, ks, kb0, kb1);                           \
    it_ += num_pairs;                                                 \
  }                                                                   \
  (void)ks;                                                           \
  (void)tile;

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    uint32_t stage = 0, phase = 0;
    MRCLIP_G2_FOR_ITEMS {
      MRCLIP_G2_NEXT_ITEM
      const int row0 = rb * 2 * kBM + (int)cta * kBM;          // this CTA's 128 rows of the pair tile
      const int col0 = dt * kGemmBN + (int)cta * (kGemmBN / 2);  // this CTA's half of the B tile
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_empty(stage), phase ^ 1);
        if (elect_one()) {
          const uint32_t full_leader = mapa_shared(bar_full(stage), 0);
          if (cta == 0) mbar_expect_tx(bar_full(stage), 2 * kG2StageBytes);   // bytes of both CTAs
          const uint32_t dst = stage_base + stage * kG2StageBytes;
          if (A_MN) {
            tma_load_2d_pair(dst, &tmA, full_leader, row0, kb * kBK);
            tma_load_2d_pair(dst + 8192, &tmA, full_leader, row0 + 64, kb * kBK);
          } else {
            tma_load_2d_pair(dst, &tmA, full_leader, kb * kBK, row0);
          }
          tma_load_2d_pair(dst + 16384, &tmB, full_leader, col0, kb * kBK);
          tma_load_2d_pair(dst + 16384 + 8192, &tmB, full_leader, col0 + 64, kb * kBK);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA only)
    if (cta == 0) {
      uint32_t stage = 0, phase = 0, acc_use = 0;
      MRCLIP_G2_FOR_ITEMS {
        MRCLIP_G2_NEXT_ITEM
        (void)rb;
        (void)dt;
        const uint32_t buf = acc_use & 1, use = acc_use >> 1;
        mbar_wait(bar_accempty(buf), (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kGemmBN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full(stage), phase);
          tc_fence_after();
          const uint32_t a = stage_base + stage * kG2StageBytes;
          const uint64_t bdesc = make_mnmajor_sw128_desc(a + 16384, 8192);
          const uint64_t adesc = A_MN ? make_mnmajor_sw128_desc(a, 8192) : make_kmajor_sw128_desc(a);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_pair(d_tmem, adesc + (A_MN ? 128 * k : 2 * k), bdesc + 128 * k, IDESC,
                             (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit_pair(bar_empty(stage), 3);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit_pair(bar_accfull(buf), 3);
        __syncwarp();
        ++acc_use;
      }
    }
  } else {
    // ===================================================================== epilogue warps (both CTAs)
    const uint32_t q = warp & 3;
    const uint32_t h = (warp - 2) >> 2;
    const uint32_t row_in_tile = q * 32 + lane;
    uint32_t acc_use = 0;
    MRCLIP_G2_FOR_ITEMS {
      MRCLIP_G2_NEXT_ITEM
      const uint32_t buf = acc_use & 1, use = acc_use >> 1;
      // stream-K roles of this segment
      const bool sk_contrib = SK && kb0 > 0;
      const bool sk_owner = SK && kb0 == 0 && kb1 < p.num_kb;
      int sk_others = 0;                 // owner: how many later pairs hold a part of this tile
      if (sk_owner) {
        const long long tile_end = (long long)(tile - sk_base_tile + 1) * p.num_kb;      // in remainder units
        for (int j = pair + 1; j < num_pairs && (long long)j * sk_total / num_pairs < tile_end; ++j) ++sk_others;
      }
      mbar_wait(bar_accfull(buf), use & 1);
      tc_fence_after();
      const int grow = rb * 2 * kBM + (int)cta * kBM + row_in_tile;
      float* out_row = p.dpart + ((size_t)ks * p.m_pad + grow) * p.d_pad + dt * kGemmBN;
      float mul = 1.f;
      if (p.out != nullptr) {
        mul = p.coef * __ldg(p.scale);
        if (p.grad_out != nullptr) mul *= __ldg(p.grad_out);
      }
      if (sk_owner) {    // every contributor's partial has landed (16 epilogue warps each)
        if (lane == 0) {
          const int want = sk_others * 2 * kEpiWarps;
          const long long t0 = clock64();
          while (ld_acquire_gpu(p.sk_flags + tile) < want) {
            if (clock64() - t0 > MRCLIP_SPIN_LIMIT_CYCLES) {
              printf("mrclip: stream-K owner timed out (pair %d tile %d want %d have %d)\n", pair, tile, want,
                     ld_acquire_gpu(p.sk_flags + tile));
              __trap();
            }
          }
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int c0 = h * 128; c0 < (int)(h + 1) * 128; c0 += (PUSH == 2 ? 64 : 32)) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + buf * kGemmBN + c0, r);
        if (SK && (sk_contrib || sk_owner)) {
          tmem_ld_wait();
          if (sk_contrib) {   // raw accumulator -> this pair's partial slot; nothing else to do for these columns
            float* prow = p.sk_part + (((size_t)pair * 2 + cta) * kBM + row_in_tile) * kGemmBN + c0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(prow + 4 * j) =
                  make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                              __uint_as_float(r[4 * j + 3]));
            continue;
          }
          for (int o = 1; o <= sk_others; ++o) {
            const float* prow = p.sk_part + (((size_t)(pair + o) * 2 + cta) * kBM + row_in_tile) * kGemmBN + c0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(prow + 4 * j));
              r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + v.x);
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + v.y);
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + v.z);
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + v.w);
            }
          }
        }
        if (PUSH == 2) {
          // bf16 payload: 64 accumulator columns -> one 128-byte row piece in the owner's bf16 receive slot
          uint32_t r2[32];
          tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + buf * kGemmBN + c0 + 32, r2);
          tmem_ld_wait();
          const int col = dt * kGemmBN + c0;
          const int qo = grow / p.peer_n, lrow = grow - qo * p.peer_n;
          __nv_bfloat16* orow = nullptr;
          if (grow < p.m_rows)
            orow = reinterpret_cast<__nv_bfloat16*>(__ldg(p.peer + qo)) + ((size_t)p.peer_rank * p.peer_n + lrow) * p.out_ld;
          if (col + 64 <= p.d_valid && (p.out_ld & 7) == 0) {
            const uint32_t my_row = staging_base + ((warp - 2) * 32 + lane) * kG2PushRowBytes;
            bulk_wait_read0();                     // this lane's previous row piece has left the staging tile
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o, o2;
              o.x = pack_bf16x2(__uint_as_float(r[8 * j]) * mul, __uint_as_float(r[8 * j + 1]) * mul);
              o.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * mul, __uint_as_float(r[8 * j + 3]) * mul);
              o.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * mul, __uint_as_float(r[8 * j + 5]) * mul);
              o.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * mul, __uint_as_float(r[8 * j + 7]) * mul);
              o2.x = pack_bf16x2(__uint_as_float(r2[8 * j]) * mul, __uint_as_float(r2[8 * j + 1]) * mul);
              o2.y = pack_bf16x2(__uint_as_float(r2[8 * j + 2]) * mul, __uint_as_float(r2[8 * j + 3]) * mul);
              o2.z = pack_bf16x2(__uint_as_float(r2[8 * j + 4]) * mul, __uint_as_float(r2[8 * j + 5]) * mul);
              o2.w = pack_bf16x2(__uint_as_float(r2[8 * j + 6]) * mul, __uint_as_float(r2[8 * j + 7]) * mul);
              sts_u4(my_row + 16 * j, o);
              sts_u4(my_row + 64 + 16 * j, o2);
            }
            fence_proxy_async_smem();
            if (orow != nullptr) {
              bulk_store(orow + col, my_row, 128);
              bulk_commit();
            }
          } else if (orow != nullptr) {            // ragged D: element stores
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (col + j < p.d_valid) orow[col + j] = __float2bfloat16(__uint_as_float(r[j]) * mul);
              if (col + 32 + j < p.d_valid) orow[col + 32 + j] = __float2bfloat16(__uint_as_float(r2[j]) * mul);
            }
          }
          continue;
        }
        tmem_ld_wait();
        if (PUSH && p.peer != nullptr && dt * kGemmBN + c0 + 32 <= p.d_valid && (p.out_ld & 3) == 0) {
          // push epilogue: scale into the warp's staging tile, then one 128-byte bulk copy per row to the owner
          const uint32_t my_row = staging_base + ((warp - 2) * 32 + lane) * kG2PushRowBytes;
          bulk_wait_read0();                       // this lane's previous row has left the staging tile
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts_f4(my_row + 16 * j,
                   make_float4(__uint_as_float(r[4 * j]) * mul, __uint_as_float(r[4 * j + 1]) * mul,
                               __uint_as_float(r[4 * j + 2]) * mul, __uint_as_float(r[4 * j + 3]) * mul));
          fence_proxy_async_smem();
          if (grow < p.m_rows) {
            bulk_store(gemm_out_row_f32(p, grow) + dt * kGemmBN + c0, my_row, 128);
            bulk_commit();
          }
        } else if (p.out != nullptr) {
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
      if (sk_contrib) {     // this warp's share of the partial is written: tell the tile's owner
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(p.sk_flags + tile, 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(bar_accempty(buf), 0));
      ++acc_use;
    }
    if (PUSH) bulk_wait0();   // every pushed row is complete before the kernel (and the cross-rank barrier) ends
  }
#undef MRCLIP_G2_FOR_ITEMS
#undef MRCLIP_G2_NEXT_ITEM
  tc_fence_before();
  if (PUSH && p.sig.ranks > 1) peer_signal_when_grid_done(p.sig, CH_DTEXT, gridDim.x);   // (contains the __syncthreads)
  __syncthreads();
  cluster_sync_all();      // nobody frees TMEM or exits while the peer may still signal / read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace mrclip
