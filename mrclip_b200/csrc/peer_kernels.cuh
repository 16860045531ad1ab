// Kernels of the multi-GPU step that move data over NVLink peer memory and synchronise through the flags of
// peer_sync.cuh: pack + all-gather by peer stores, the forward reductions that publish their statistics to every
// rank, the owner side of the fused GEMM -> reduce-scatter, and the scalar tails (loss mean, d logit_scale).
#pragma once
#include "aux_kernels.cuh"
#include "peer_sync.cuh"
#include "ptx.cuh"

namespace mrclip {

// ---------------------------------------------------------------------------------------------------------------
// pack + all-gather by peer stores, one launch: both modalities of this rank's n rows are cast to bf16; the image rows
// go to the local operand buffer, the text rows to slot [row0, row0 + rows) of EVERY rank's gathered text buffer
// (16-byte stores, a warp covers 512 contiguous bytes of a row: full NVLink packets); the last block raises CH_TEXT.
// Replaces the autocast casts + the text all_gather of loss.py:51-57 (the image rows of other ranks are never needed).
struct Pack2Params {
  const void* img_src;
  const void* txt_src;
  int img_dtype, txt_dtype;
  int rows, d, ld;
  long img_src_ld, txt_src_ld;
  __nv_bfloat16* img_dst;                 // local [rows, ld]
  const unsigned long long* txt_peers;    // [ranks] base of every rank's gathered text buffer [N, ld]
  __nv_bfloat16* txt_local;               // own gathered text buffer (ranks == 1: the only destination)
  long row0;                              // first global row of this rank
};

__device__ __forceinline__ uint4 load_pack8(const void* src, int dtype, long off, int valid, bool vec) {
  // 8 consecutive elements from element offset `off` (the first `valid` exist) -> 8 bf16; vec: 16-byte loads allowed
  if (vec && valid == 8 && dtype == DT_BF16)
    return *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + off);
  if (vec && valid == 8 && dtype == DT_F32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + off);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + off + 4);
    return make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = j < valid ? load_as_float(src, dtype, (size_t)(off + j)) : 0.f;
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// push_mode: 0 = write this rank's buffers only (one rank, or the rows travel on the copy engines, mrclip_cabi.cu
// push_rows_async: then the CH_TEXT epoch is advanced here, before either consumer starts); 1 = store to every rank and
// raise CH_TEXT.
enum : int { PACK_LOCAL = 0, PACK_PUSH = 1 };
__device__ __forceinline__ void pack_bump_epoch(const PeerInfo& pi, int push_mode) {
  if (push_mode == PACK_LOCAL && pi.ranks > 1 && blockIdx.x == 0 && threadIdx.x == 0) pi.ctl->epoch[CH_TEXT] += 1;
}

__global__ void __launch_bounds__(256)
pack2_push_kernel(const Pack2Params p, const PeerInfo pi, int vec_ok, int push_mode) {
  const int vpr = p.ld / 8;
  const long total = (long)p.rows * vpr;
  pack_bump_epoch(pi, push_mode);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < 2 * total; i += (long)gridDim.x * blockDim.x) {
    const bool is_txt = i >= total;
    const long k = is_txt ? i - total : i;
    const long r = k / vpr;
    const int c = (int)(k - r * vpr) * 8;
    const void* src = is_txt ? p.txt_src : p.img_src;
    const int dt = is_txt ? p.txt_dtype : p.img_dtype;
    const long sld = is_txt ? p.txt_src_ld : p.img_src_ld;
    int valid = p.d - c;
    valid = valid < 0 ? 0 : (valid > 8 ? 8 : valid);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (valid > 0) o = load_pack8(src, dt, r * sld + c, valid, vec_ok != 0);
    if (!is_txt) {
      *reinterpret_cast<uint4*>(p.img_dst + r * p.ld + c) = o;
    } else if (pi.ranks <= 1 || push_mode == PACK_LOCAL) {
      *reinterpret_cast<uint4*>(p.txt_local + (p.row0 + r) * p.ld + c) = o;
    } else {
      for (int q = 0; q < pi.ranks; ++q)
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(__ldg(p.txt_peers + q)) + (p.row0 + r) * p.ld + c) = o;
    }
  }
  if (pi.ranks > 1 && push_mode == PACK_PUSH) peer_signal_when_grid_done(pi, CH_TEXT, gridDim.x);
}

// ---------------------------------------------------------------------------------------------------------------
// Feature hand-off fusion (reference model.py:282-301, :324): pack2_push_kernel for UN-normalised tower
// outputs.  One warp per (modality, row): y = x / max(||x||, 1e-12) in fp32 (F.normalize's definition), cast to bf16,
// stored / pushed like the plain pack; 1/max(||x||, eps) is kept per row for the backward and the logit scale is
// exponentiated once (scale_out[0] = exp(log_scale[0])).  Replaces two F.normalize passes, the autocast casts and
// logit_scale.exp() in front of the loss.
__global__ void __launch_bounds__(256)
packnorm2_push_kernel(const Pack2Params p, const PeerInfo pi, int vec_ok, int push_mode, float* __restrict__ inv_norm,
                      const float* __restrict__ log_scale, float* __restrict__ scale_out) {
  const int lane = threadIdx.x & 31;
  pack_bump_epoch(pi, push_mode);
  const int warps_per_block = blockDim.x >> 5;
  const int vpr = p.ld / 8;
  if (blockIdx.x == 0 && threadIdx.x == 0 && log_scale != nullptr) scale_out[0] = expf(log_scale[0]);
  for (long w = blockIdx.x * (long)warps_per_block + (threadIdx.x >> 5); w < 2L * p.rows; w += (long)gridDim.x * warps_per_block) {
    const bool is_txt = w >= p.rows;
    const long r = is_txt ? w - p.rows : w;
    const void* src = is_txt ? p.txt_src : p.img_src;
    const int dt = is_txt ? p.txt_dtype : p.img_dtype;
    const long sld = is_txt ? p.txt_src_ld : p.img_src_ld;
    float ss = 0.f;
    for (int c = lane; c < p.d; c += 32) {
      const float v = load_as_float(src, dt, (size_t)(r * sld + c));
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) inv_norm[w] = inv;
    for (int v8 = lane; v8 < vpr; v8 += 32) {
      const int c = v8 * 8;
      float f[8];
      if (vec_ok && c + 8 <= p.d && dt == DT_F32) {
        const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + r * sld + c);
        const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + r * sld + c + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
      } else if (vec_ok && c + 8 <= p.d && dt == DT_BF16) {
        const uint4 a = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + r * sld + c);
        const uint32_t u[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[2 * j] = __uint_as_float(u[j] << 16);
          f[2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (c + j < p.d) ? load_as_float(src, dt, (size_t)(r * sld + c + j)) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= inv;
      const uint4 o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      if (!is_txt) {
        *reinterpret_cast<uint4*>(p.img_dst + r * p.ld + c) = o;
      } else if (pi.ranks <= 1 || push_mode == PACK_LOCAL) {
        *reinterpret_cast<uint4*>(p.txt_local + (p.row0 + r) * p.ld + c) = o;
      } else {
        for (int q = 0; q < pi.ranks; ++q)
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(__ldg(p.txt_peers + q)) + (p.row0 + r) * p.ld + c) = o;
      }
    }
  }
  if (pi.ranks > 1 && push_mode == PACK_PUSH) peer_signal_when_grid_done(pi, CH_TEXT, gridDim.x);
}

// Backward of the normalisation, in place on a gradient block: with y = x / ||x|| (the packed bf16 rows) and g = dL/dy,
//   dL/dx = (g - y <g, y>) / ||x||.     One warp per row; g is [rows, d] of MRCLIP_DT_* with leading dimension g_ld.
// Replaces the autograd of F.normalize (model.py:282-301) behind the loss.
__global__ void __launch_bounds__(256)
normalize_bwd_kernel(const __nv_bfloat16* __restrict__ y, long y_ld, const float* __restrict__ inv_norm, int rows, int d,
                     void* __restrict__ g, int g_dtype, long g_ld) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (long r = blockIdx.x * (long)warps_per_block + (threadIdx.x >> 5); r < rows; r += (long)gridDim.x * warps_per_block) {
    float dot = 0.f;
    for (int c = lane; c < d; c += 32)
      dot = fmaf(load_as_float(g, g_dtype, (size_t)(r * g_ld + c)), __bfloat162float(y[r * y_ld + c]), dot);
    dot = warp_sum(dot);
    const float inv = inv_norm[r];
    for (int c = lane; c < d; c += 32) {
      const size_t i = (size_t)(r * g_ld + c);
      store_from_float(g, g_dtype, i, (load_as_float(g, g_dtype, i) - __bfloat162float(y[r * y_ld + c]) * dot) * inv);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Forward reductions that publish straight into every rank's statistics block  stats[ranks][3][N]  (plane 0/1: per-column
// (max2, sum) of the publisher's rows, plane 2: its row LSEs in [:n]).  reduce_rows_pub runs first, reduce_cols_pub
// raises CH_STATS from its last block.  ranks == 1: plain local writes, no flag.
__global__ void reduce_rows_pub_kernel(const float2* __restrict__ row_part, int slots, int m_rows, int m_pad,
                                       const unsigned long long* __restrict__ stats_peers, float* __restrict__ stats_local,
                                       long plane2_off, const PeerInfo pi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m_rows) return;
  float m = -CUDART_INF_F, l = 0.f;
  for (int s = 0; s < slots; ++s) {
    const float2 v = row_part[(size_t)s * m_pad + i];
    lse2_merge(m, l, v.x, v.y);
  }
  const float lse = m + log2f(fmaxf(l, 1e-37f));
  if (pi.ranks <= 1) {
    stats_local[plane2_off + i] = lse;
  } else {
    for (int q = 0; q < pi.ranks; ++q) reinterpret_cast<float*>(__ldg(stats_peers + q))[plane2_off + i] = lse;
  }
}

__global__ void __launch_bounds__(1024)
reduce_cols_pub_kernel(const float* __restrict__ col_l, const float* __restrict__ col_c, int bands, int n_cols, int n_pad,
                       const unsigned long long* __restrict__ stats_peers, float* __restrict__ stats_local,
                       long plane0_off, long plane1_off, const PeerInfo pi) {
  extern __shared__ float wgt[];          // [bands]
  __shared__ float red[32];
  __shared__ float part[16][64];
  const int cb = blockIdx.x, ncb = n_pad / 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -CUDART_INF_F;
  for (int b = tid; b < bands; b += 1024) {
    const float c = col_c[(size_t)b * ncb + cb];
    wgt[b] = c;
    m = fmaxf(m, c);
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < 32; ++w) m = fmaxf(m, red[w]);
  for (int b = tid; b < bands; b += 1024) wgt[b] = exp2f(wgt[b] - m);
  __syncthreads();
  const int c = tid & 63, slice = tid >> 6;
  const int j = cb * 64 + c;
  float l = 0.f;
  for (int b = slice; b < bands; b += 16) l = fmaf(col_l[(size_t)b * n_pad + j], wgt[b], l);
  part[slice][c] = l;
  __syncthreads();
  if (slice == 0 && j < n_cols) {
    for (int k = 1; k < 16; ++k) l += part[k][c];
    if (pi.ranks <= 1) {
      stats_local[plane0_off + j] = m;
      stats_local[plane1_off + j] = l;
    } else {
      for (int q = 0; q < pi.ranks; ++q) {
        float* s = reinterpret_cast<float*>(__ldg(stats_peers + q));
        s[plane0_off + j] = m;
        s[plane1_off + j] = l;
      }
    }
  }
  if (pi.ranks > 1) peer_signal_when_grid_done(pi, CH_STATS, gridDim.x);
}

// Waits for every rank's statistics, then  lse2_col[j] = merge over ranks (padding +inf)  and
// lse2_row_all[q*n + i] = rank q's row LSE.  stats: [ranks][3][N] (this rank's block).
__global__ void merge_stats_kernel(const float* __restrict__ stats, int ranks, int n_per_rank, int n_cols, int n_pad,
                                   float* __restrict__ lse2_col, float* __restrict__ lse2_row_all, const PeerInfo pi) {
  peer_wait_all(pi, CH_STATS);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pad) return;
  if (j >= n_cols) {
    lse2_col[j] = CUDART_INF_F;
    lse2_row_all[j] = CUDART_INF_F;
    return;
  }
  float m = -CUDART_INF_F, l = 0.f;
  for (int w = 0; w < ranks; ++w)
    lse2_merge(m, l, stats[((size_t)w * 3 + 0) * n_cols + j], stats[((size_t)w * 3 + 1) * n_cols + j]);
  lse2_col[j] = m + log2f(fmaxf(l, 1e-37f));
  const int q = j / n_per_rank;
  lse2_row_all[j] = stats[((size_t)q * 3 + 2) * n_cols + (j - q * n_per_rank)];
}

// loss of this rank's rows (clip_loss_kernel) that, in the global-loss modes, also publishes it to every rank
// (scal block: float [2][kPeerMaxRanks] behind the sig block; plane 0 = loss, plane 1 = d logit_scale) and raises
// CH_LOSS; loss_mean_kernel then averages.  Replaces the all_reduce of the reference's identical per-rank global loss.
__global__ void clip_loss_pub_kernel(const float* __restrict__ lse2_row, const float* __restrict__ lse2_col,
                                     const float* __restrict__ diag2, int m_rows, int label_offset,
                                     float* __restrict__ loss_local, float* __restrict__ loss_out, int publish,
                                     const PeerInfo pi) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < m_rows; i += blockDim.x)
    acc += (double)lse2_row[i] + (double)lse2_col[label_offset + i] - 2.0 * (double)diag2[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    const float v = (float)(t * 0.6931471805599453 / (2.0 * m_rows));
    loss_local[0] = v;
    if (publish && pi.ranks > 1) {
      for (int k = 0; k < pi.ranks; ++k) peer_scal(pi, k)[pi.rank] = v;
    } else if (loss_out != nullptr) {
      loss_out[0] = v;
    }
  }
  if (publish && pi.ranks > 1) peer_signal_when_grid_done(pi, CH_LOSS, 1);
}

// out[0] = mean over ranks of scal[plane][0..ranks)   (after the peers' values have landed on `channel`)
__global__ void scal_mean_kernel(int plane, int channel, float* __restrict__ out, const PeerInfo pi) {
  peer_wait_all(pi, channel);
  if (threadIdx.x == 0) {
    const float* scal = local_scal(pi);
    double t = 0.0;
    for (int k = 0; k < pi.ranks; ++k) t += (double)scal[plane * kPeerMaxRanks + k];
    out[0] = (float)(t / pi.ranks);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// R2(me, q) of MODE_FWDEU (row_ent_split_kernel) summed over this rank's rows and delivered where it is needed:
//   r2_row_tot[0] = sum_q R2(me, q)   (local)        r2_in[me] on rank q = R2(me, q)   (plain store over NVLink)
// so that rank q later finds  R2(*, q) = sum_p r2_in[p]  for  s dL_q/ds = <dT_q, T_q> + ln2/(2n)(R2(q,*) - R2(*,q)).
// Blocks accumulate into acc_slots [64][ranks] (zero on entry); the last block reduces, publishes and re-zeroes them,
// so a step whose backward never runs leaves nothing stale behind.  The stores are ordered before CH_STATS, which
// reduce_cols_pub_kernel raises later in the same stream.
__global__ void row_ent_pub_kernel(const float2* __restrict__ row_part, const float* __restrict__ row_ent, int slots,
                                   int slots_per_rank, int ranks, int m_rows, int m_pad, const float* __restrict__ lse2_row,
                                   float* __restrict__ acc_slots, int* __restrict__ done, float* __restrict__ r2_row_tot,
                                   const PeerInfo pi) {
  extern __shared__ float acc[];   // [ranks]
  __shared__ int s_last;
  for (int q = threadIdx.x; q < ranks; q += blockDim.x) acc[q] = 0.f;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float lse = i < m_rows ? lse2_row[i] : 0.f;
  for (int q = 0; q < ranks; ++q) {
    float r = 0.f;
    if (i < m_rows) {
      for (int s = q * slots_per_rank; s < (q + 1) * slots_per_rank && s < slots; ++s) {
        const float m = row_part[(size_t)s * m_pad + i].x;
        const float u = row_ent[(size_t)s * m_pad + i];
        r = fmaf(exp2f(m - lse), u, r);       // m = -inf for an empty slot: weight 0, u = 0
      }
    }
    r = warp_sum(r);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc + q, r);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < ranks; q += blockDim.x) atomicAdd(acc_slots + (size_t)(blockIdx.x & 63) * ranks + q, acc[q]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(done, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float tot = 0.f;
  for (int q = threadIdx.x; q < ranks; q += blockDim.x) {
    float v = 0.f;
    for (int k = 0; k < 64; ++k) {
      v += __ldcg(acc_slots + (size_t)k * ranks + q);
      acc_slots[(size_t)k * ranks + q] = 0.f;
    }
    tot += v;
    if (pi.ranks > 1) peer_r2in(pi, q)[pi.rank] = v;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < ranks; q += blockDim.x) acc[q] = 0.f;
  __syncthreads();
  if (tot != 0.f) atomicAdd(acc, tot);
  __syncthreads();
  if (threadIdx.x == 0) {
    r2_row_tot[0] = acc[0];
    *done = 0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Owner side of the fused reduce-scatter: waits for every rank's pushed tiles (CH_DTEXT), then
// out[r, c] = sum_k slots[k][r][c] (slots fp32 or bf16, four columns per thread when d % 4 == 0) and, optionally,
// <out, feat> spread over dot_slots[64].
template <bool SLOTS_BF16>
__global__ void sum_slots_wait_kernel(const void* __restrict__ slots, int nslots, int rows, int d, void* __restrict__ out,
                                      int out_dtype, long out_ld, const __nv_bfloat16* __restrict__ feat, long feat_ld,
                                      float* __restrict__ dot_out, const PeerInfo pi) {
  peer_wait_all(pi, CH_DTEXT);
  const long slot_stride = (long)rows * d;
  float dot = 0.f;
  if ((d & 3) == 0) {
    const int dq = d >> 2;
    const long total4 = (long)rows * dq;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
      const long r = i / dq;
      const int c = (int)(i - r * dq) << 2;
      const long e = r * d + c;
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < nslots; ++k) {
        if (SLOTS_BF16) {
          const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(slots) + k * slot_stride + e);
          a[0] += __uint_as_float(v.x << 16);
          a[1] += __uint_as_float(v.x & 0xffff0000u);
          a[2] += __uint_as_float(v.y << 16);
          a[3] += __uint_as_float(v.y & 0xffff0000u);
        } else {
          const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(slots) + k * slot_stride + e);
          a[0] += v.x;
          a[1] += v.y;
          a[2] += v.z;
          a[3] += v.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) store_from_float(out, out_dtype, (size_t)(r * out_ld + c + j), a[j]);
      if (feat != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dot = fmaf(a[j], __bfloat162float(feat[r * feat_ld + c + j]), dot);
      }
    }
  } else {
    const long total = (long)rows * d;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      float a = 0.f;
      for (int k = 0; k < nslots; ++k)
        a += SLOTS_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(slots)[k * slot_stride + i])
                        : reinterpret_cast<const float*>(slots)[k * slot_stride + i];
      const long r = i / d;
      const int c = (int)(i - r * d);
      store_from_float(out, out_dtype, (size_t)(r * out_ld + c), a);
      if (feat != nullptr) dot = fmaf(a, __bfloat162float(feat[r * feat_ld + c]), dot);
    }
  }
  if (feat == nullptr) return;     // grid-uniform
  dot = warp_sum(dot);
  __shared__ float part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(dot_out + (blockIdx.x & 63), v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Entropy sums of the rescale pass (emat_transform_kernel<WSUM>: msums [slots][2][ranks], plane 0 = row softmax, plane 1
// = column softmax, split by the rank that owns the column) delivered where d logit_scale needs them:
//   r2_row_tot[0] = sum_q msums[.][0][q]  (stays here)       r2_in[me] on rank q = sum over slots of msums[.][1][q].
// One block; the stores are ordered before CH_DTEXT, which the GEMM push raises later in the same stream.
// Replaces the W-float all_reduce of the column entropies.
__global__ void msums_pub_kernel(const float* __restrict__ msums, int slots, float* __restrict__ r2_row_tot,
                                 const PeerInfo pi) {
  __shared__ float rowp[64];
  const int q = threadIdx.x;
  float r = 0.f, c = 0.f;
  if (q < pi.ranks) {
    for (int k = 0; k < slots; ++k) {
      r += msums[((size_t)k * 2 + 0) * pi.ranks + q];
      c += msums[((size_t)k * 2 + 1) * pi.ranks + q];
    }
    peer_r2in(pi, q)[pi.rank] = c;
  }
  if (q < 64) rowp[q] = q < pi.ranks ? r : 0.f;
  __syncthreads();
  if (q == 0) {
    float t = 0.f;
    for (int k = 0; k < pi.ranks; ++k) t += rowp[k];
    r2_row_tot[0] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// d logit_scale, one block of 128 threads.  With L the local loss, g = grad_out, s = logit_scale, k = ln2 * coef_k:
//   mode 0 (one rank, entropy sums):        ds = g/s * (L + k * sum(msums[0..count)))
//   mode 1 (one rank, <dI, I> dot):         ds = ds_out[0] as accumulated by the GEMM's reduce pass (nothing to do)
//   mode 2 (ranks > 1, forward-side sums):  ds = (sum(dot_slots) + g*k*(r2_row_tot - sum_p r2_in[p])) / s
//                                           (dot_slots = <dT_r, T_r> from the slot sum; R2 in log2 units)
//   mode 3 (ranks > 1, entropy sums):       ds = g/s * (L + k * (r2_row_tot + sum_p r2_in[p]))
// publish: the global-loss modes average ds over the ranks (scal plane 1, CH_DSCALE; scal_mean_kernel follows).
struct DsParams {
  int mode;
  const float* gout;       // [1] or null (= 1)
  const float* scale;      // [1]
  const float* loss_local; // [1]
  float kfac;
  const float* msums;      // mode 0
  int count;
  float* dot_slots;        // mode 2: [64], zeroed here once consumed
  const float* r2_row_tot; // modes 2, 3: [1]
  float* ds_out;           // [1]
  int publish;
};
__global__ void ds_finish_kernel(const DsParams p, const PeerInfo pi) {
  __shared__ float red[2][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float a = 0.f, c = 0.f;
  if (p.mode == 0) {
    for (int i = tid; i < p.count; i += blockDim.x) a += p.msums[i];
  } else if (p.mode == 2 && tid < 64) {
    a = p.dot_slots[tid];
  }
  if (p.mode >= 2 && tid < pi.ranks) c = local_r2in(pi)[tid];
  a = warp_sum(a);
  c = warp_sum(c);
  if (lane == 0) {
    red[0][warp] = a;
    red[1][warp] = c;
  }
  __syncthreads();
  if (tid == 0) {
    float ta = 0.f, tc = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      ta += red[0][w];
      tc += red[1][w];
    }
    const float g = p.gout != nullptr ? p.gout[0] : 1.f, s = p.scale[0];
    float ds;
    if (p.mode == 0) ds = g / s * (p.loss_local[0] + p.kfac * ta);
    else if (p.mode == 1) ds = p.ds_out[0];
    else if (p.mode == 2) ds = (ta + g * p.kfac * (p.r2_row_tot[0] - tc)) / s;
    else ds = g / s * (p.loss_local[0] + p.kfac * (p.r2_row_tot[0] + tc));
    p.ds_out[0] = ds;
    if (p.publish && pi.ranks > 1)
      for (int k = 0; k < pi.ranks; ++k) peer_scal(pi, k)[kPeerMaxRanks + pi.rank] = ds;
  }
  __syncthreads();
  if (p.mode == 2 && tid < 64) p.dot_slots[tid] = 0.f;
  if (p.publish && pi.ranks > 1) peer_signal_when_grid_done(pi, CH_DSCALE, 1);
}

}  // namespace mrclip
