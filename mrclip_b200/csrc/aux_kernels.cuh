// HBM-bound helper kernels around the tile and GEMM kernels: packing features to bf16, the
// reductions that turn per-tile partials into LSE vectors, losses and gradients, the E -> G
// rescale pass of the emat backend with its guard, the peer-store copy and slot sum of the
// NVLink collectives, and the [N,D] -> [D,N] transpose the fused backend needs.
#pragma once
#include "ptx.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

namespace mrclip {

enum : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };

__device__ __forceinline__ float load_as_float(const void* p, int dtype, size_t i) {
  if (dtype == DT_F32) return reinterpret_cast<const float*>(p)[i];
  if (dtype == DT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void store_from_float(void* p, int dtype, size_t i, float v) {
  if (dtype == DT_F32)
    reinterpret_cast<float*>(p)[i] = v;
  else if (dtype == DT_BF16)
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}

// src [rows, d] (any float dtype, leading dim src_ld) -> dst bf16 [rows, dst_ld], zero padded.
__global__ void pack_bf16_kernel(const void* __restrict__ src, int dtype, int rows, int d,
                                 long src_ld, __nv_bfloat16* __restrict__ dst, int dst_ld) {
  const long total = (long)rows * dst_ld;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long r = i / dst_ld;
    const int c = (int)(i - r * dst_ld);
    const float v = (c < d) ? load_as_float(src, dtype, (size_t)(r * src_ld + c)) : 0.f;
    dst[i] = __float2bfloat16_rn(v);
  }
}

// fast path of the same: d, src_ld, dst_ld multiples of 8 and 16-byte aligned bases; 8 elements per thread
__global__ void pack_bf16_vec8_kernel(const void* __restrict__ src, int dtype, int rows, int d, long src_ld,
                                      __nv_bfloat16* __restrict__ dst, int dst_ld) {
  const int vpr = dst_ld / 8;   // vectors per row
  const long total = (long)rows * vpr;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / vpr;
    const int c = (int)(i - r * vpr) * 8;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (c < d) {
      if (dtype == DT_BF16) {
        o = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + r * src_ld + c);
      } else if (dtype == DT_F32) {
        const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + r * src_ld + c);
        const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + r * src_ld + c + 4);
        o.x = pack_bf16x2(a.x, a.y);
        o.y = pack_bf16x2(a.z, a.w);
        o.z = pack_bf16x2(b.x, b.y);
        o.w = pack_bf16x2(b.z, b.w);
      } else {
        const uint4 hv = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(src) + r * src_ld + c);
        const __half2* h2 = reinterpret_cast<const __half2*>(&hv);
        const float2 f0 = __half22float2(h2[0]), f1 = __half22float2(h2[1]);
        const float2 f2 = __half22float2(h2[2]), f3 = __half22float2(h2[3]);
        o.x = pack_bf16x2(f0.x, f0.y);
        o.y = pack_bf16x2(f1.x, f1.y);
        o.z = pack_bf16x2(f2.x, f2.y);
        o.w = pack_bf16x2(f3.x, f3.y);
      }
    }
    *reinterpret_cast<uint4*>(dst + r * dst_ld + c) = o;
  }
}

// src bf16 [rows, src_ld] (cols valid) -> dst bf16 [cols, dst_ld]
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, int rows, int cols,
                                      long src_ld, __nv_bfloat16* __restrict__ dst, long dst_ld) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const int r = r0 + i;
    for (int j = threadIdx.x; j < 64; j += blockDim.x) {
      const int c = c0 + j;
      tile[i][j] = (r < rows && c < cols) ? src[(size_t)r * src_ld + c] : __float2bfloat16_rn(0.f);
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const int c = c0 + i;
    for (int j = threadIdx.x; j < 64; j += blockDim.x) {
      const int r = r0 + j;
      if (c < cols && r < rows) dst[(size_t)c * dst_ld + r] = tile[j][i];
    }
  }
}

__device__ __forceinline__ void lse2_merge(float& m, float& l, float m2, float l2) {
  if (l2 <= 0.f) return;
  const float mn = fmaxf(m, m2);
  l = l * exp2f(m - mn) + l2 * exp2f(m2 - mn);
  m = mn;
}

// per-row merge of the (max2,sum) partial slots -> lse2_row
__global__ void reduce_rows_kernel(const float2* __restrict__ row_part, int slots, int m_rows,
                                   int m_pad, float* __restrict__ lse2_row) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m_rows) return;
  float m = -CUDART_INF_F, l = 0.f;
  for (int s = 0; s < slots; ++s) {
    const float2 v = row_part[(size_t)s * m_pad + i];
    lse2_merge(m, l, v.x, v.y);
  }
  lse2_row[i] = m + log2f(fmaxf(l, 1e-37f));
}

// per-column merge over the 32-row bands -> (max2, sum) of this rank's rows.
// The forward references all 64 columns of a (band, column-block) sub-tile to one value c (col_c), so a block
// that owns one column block first finds g = max_band c and the weights 2^(c - g) (one exp2 per band, in shared
// memory) and then only streams col_l with one FMA per element: 64 columns x 16 band slices per block.
__global__ void __launch_bounds__(1024)
reduce_cols_kernel(const float* __restrict__ col_l, const float* __restrict__ col_c, int bands, int n_cols,
                   int n_pad, float* __restrict__ out_m, float* __restrict__ out_l) {
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
  for (int b = tid; b < bands; b += 1024) wgt[b] = exp2f(wgt[b] - m);   // m is finite: dead sub-tiles store c = 0
  __syncthreads();
  const int c = tid & 63, slice = tid >> 6;
  const int j = cb * 64 + c;
  float l = 0.f;
  for (int b = slice; b < bands; b += 16) l = fmaf(col_l[(size_t)b * n_pad + j], wgt[b], l);
  part[slice][c] = l;
  __syncthreads();
  if (slice == 0 && j < n_cols) {
    for (int k = 1; k < 16; ++k) l += part[k][c];
    out_m[j] = m;
    out_l[j] = l;
  }
}

// merge W per-rank (max2,sum) column partials -> lse2 [n_pad], padded with +inf
__global__ void merge_parts_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l,
                                   int parts, int n_cols, long part_stride, int n_pad,
                                   float* __restrict__ lse2) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pad) return;
  if (j >= n_cols) {
    lse2[j] = CUDART_INF_F;
    return;
  }
  float m = -CUDART_INF_F, l = 0.f;
  for (int w = 0; w < parts; ++w) lse2_merge(m, l, part_m[w * part_stride + j], part_l[w * part_stride + j]);
  lse2[j] = m + log2f(fmaxf(l, 1e-37f));
}

// loss = ln2/(2m) * sum_i (lse2_row[i] + lse2_col[label_i] - 2*diag2[i]); single block
__global__ void clip_loss_kernel(const float* __restrict__ lse2_row, const float* __restrict__ lse2_col,
                                 const float* __restrict__ diag2, int m_rows, int label_offset,
                                 float* __restrict__ loss) {
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
    loss[0] = (float)(t * 0.6931471805599453 / (2.0 * m_rows));
  }
}

// sums float2 partials; out[0] (+)= cx * sum(x), out2[0] (+)= cy * sum(y), times *grad_out if given
__global__ void scalar_reduce_kernel(const float2* __restrict__ part, long count, float cx, float cy,
                                     const float* __restrict__ mul_dev, const float* __restrict__ mul_dev2,
                                     float* __restrict__ out_x, float* __restrict__ out_y,
                                     int accumulate, int fold_y_into_x) {
  __shared__ double rx[32], ry[32];
  double ax = 0.0, ay = 0.0;
  for (long i = threadIdx.x; i < count; i += blockDim.x) {
    const float2 v = part[i];
    ax += v.x;
    ay += v.y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    ax += __shfl_xor_sync(0xffffffffu, ax, o);
    ay += __shfl_xor_sync(0xffffffffu, ay, o);
  }
  if ((threadIdx.x & 31) == 0) {
    rx[threadIdx.x >> 5] = ax;
    ry[threadIdx.x >> 5] = ay;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tx = 0.0, ty = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      tx += rx[w];
      ty += ry[w];
    }
    double mul = 1.0;
    if (mul_dev) mul *= (double)mul_dev[0];
    if (mul_dev2) mul *= (double)mul_dev2[0];
    if (fold_y_into_x) {
      if (out_x) out_x[0] = (accumulate ? out_x[0] : 0.f) + (float)((tx * cx + ty * cy) * mul);
    } else {
      if (out_x) out_x[0] = (accumulate ? out_x[0] : 0.f) + (float)(tx * cx * mul);
      if (out_y) out_y[0] = (accumulate ? out_y[0] : 0.f) + (float)(ty * cy * mul);
    }
  }
}

// dA_out[i, c] = coef * scale * grad_out * sum_cs dpart[cs][i][c]
__global__ void grad_reduce_kernel(const float* __restrict__ dpart, int cs, int m_rows, int d,
                                   int m_pad, int d_pad, float coef, const float* __restrict__ scale,
                                   const float* __restrict__ grad_out, void* __restrict__ out,
                                   int out_dtype, long out_ld, const __nv_bfloat16* __restrict__ dot_feat,
                                   long dot_ld, float* __restrict__ dot_out) {
  float mul = coef * scale[0];
  if (grad_out) mul *= grad_out[0];
  const int dq = d_pad / 4;
  const long total = (long)m_rows * dq;
  float dot = 0.f;   // <out, dot_feat>: with out = dA this is  scale * dLoss/dscale  (homogeneity of S = scale*A.B^T)
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long r = i / dq;
    const int c = (int)(i - r * dq) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < cs; ++k) {
      const float4 v =
          *reinterpret_cast<const float4*>(dpart + ((size_t)k * m_pad + r) * d_pad + c);
      a.x += v.x;
      a.y += v.y;
      a.z += v.z;
      a.w += v.w;
    }
    const float o[4] = {a.x * mul, a.y * mul, a.z * mul, a.w * mul};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < d) {
        store_from_float(out, out_dtype, (size_t)(r * out_ld + c + j), o[j]);
        if (dot_feat) dot = fmaf(o[j], __bfloat162float(dot_feat[r * dot_ld + c + j]), dot);
      }
  }
  if (dot_feat) {
    __shared__ float red[32];
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) atomicAdd(dot_out, t / scale[0]);
    }
  }
}

// ---- E-block guard -------------------------------------------------------------------------------
// The backward rebuilds G_ij = E_ij * (2^(c-lr_i) + 2^(c-lc_j)) from the bf16 E block, c = sub-tile reference.
// An entry whose E flushed to zero (S2_ij < c - 126) still matters only if S2_ij is within ~2^-40 of its row or
// column LSE, which needs  lr_i < c - 80  or  lc_j < c - 80  somewhere in the sub-tile.  cbmin = per 64-column
// block minimum of lse2_col; the flag is raised when any sub-tile violates the bound (then the exact recompute
// kernel rewrites the block with G and the GEMMs skip their transform).
__global__ void emat_cbmin_kernel(const float* __restrict__ lse2_col, int ncb, float* __restrict__ cbmin,
                                  int* __restrict__ flag) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *flag = 0;
  if (w >= ncb) return;
  const float2 v = reinterpret_cast<const float2*>(lse2_col + (size_t)w * 64)[lane];
  float m = fminf(v.x, v.y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) cbmin[w] = m;
}
__global__ void emat_check_kernel(const float* __restrict__ colc, const float* __restrict__ lse2_row, int m_rows,
                                  const float* __restrict__ cbmin, int ncb, float limit, int* __restrict__ flag) {
  __shared__ float s_lrmin;
  const int band = blockIdx.x;
  if (threadIdx.x < 32) {
    const int i = band * 32 + threadIdx.x;
    float m = (i < m_rows) ? lse2_row[i] : CUDART_INF_F;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) s_lrmin = m;
  }
  __syncthreads();
  const float lrmin = s_lrmin;
  bool bad = false;
  for (int cb = threadIdx.x; cb < ncb; cb += blockDim.x) {
    const float c = colc[(size_t)band * ncb + cb];
    bad |= (c - lrmin > limit) || (c - cbmin[cb] > limit);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

// E block -> G block, in place (HBM-bound: one read and one write of the bf16 [m_pad, n_pad] block).
//   G_ij = E_ij * (w_row * 2^(c - lse2_row_i) + w_col * 2^(c - lse2_col_j)),  c = colc[i/32][j/64]
//   G_i,label(i) = w_row * 2^(diag2_i - lse2_row_i) + w_col * 2^(diag2_i - lse2_col_label) - (w_row + w_col)   (exact)
// One block = one 32-row band x 1024 columns: the 1024 column factors are built once in shared memory and
// reused by the 32 rows; a thread owns 8 consecutive columns (16 bytes) of a row.  Skipped when *skip_if != 0
// (the exact recompute fallback has already written G).
//
// WSUM: the pass also accumulates what d(loss)/d(logit_scale) needs, split by the rank that owns the column
// (n_per_rank columns each).  With L_q the local loss of rank q (natural log) and P the two softmaxes,
//   scale * dL_q/dscale = L_q + ln2/(2n) * ( sum_{i in q, all j} Prow_ij log2 Prow_ij + sum_{all i, j in q} Pcol_ij log2 Pcol_ij )
// (the sum_j P = 1 identities absorb the lse and delta terms into L_q), so only the negative entropies are needed:
//   msums[0][r] += sum_{i, j in rank r} w_row * Prow_ij * log2 Prow_ij,     msums[1][r] += same with w_col * Pcol,
// log2 P = c + log2(E) - lse2 recovered from the stored exponential (positives: exact, from diag2).  The error
// of a bf16-rounded E then scales with the entropy, not with |S2|, and nothing cancels against an exact term.
template <bool WSUM>
__global__ void __launch_bounds__(256, 3)
emat_transform_kernel(uint16_t* __restrict__ emat, long ld, int m_rows, int n_pad, const float* __restrict__ colc,
                      int ncb, const float* __restrict__ lse2_row, const float* __restrict__ lse2_col,
                      const float* __restrict__ diag2, int label_offset, float w_row, float w_col,
                      const int* __restrict__ skip_if, float* __restrict__ msums_all, int n_per_rank, int ranks,
                      int msum_slots, int band0) {
  if (skip_if != nullptr && __ldg(skip_if) != 0) return;
  // the per-block sums are spread over msum_slots copies of [2][ranks] (the caller adds them up): thousands of
  // blocks adding into one address serialise in L2 and cost more than the pass itself
  float* msums = WSUM ? msums_all + (size_t)((blockIdx.y * gridDim.x + blockIdx.x) % msum_slots) * 2 * ranks : nullptr;
  __shared__ __align__(16) float cfac[1024];
  __shared__ __align__(16) float lcol[WSUM ? 1024 : 4];
  __shared__ float bacc[2][64];   // per-block sums by column owner (ranks <= 64)
  const int band = blockIdx.y + band0;     // (a launch may cover a sub-range of the 32-row bands)
  const int col0 = blockIdx.x * 1024;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* cw_row = colc + (size_t)band * ncb;
  for (int c = tid; c < 1024; c += 256) {
    const int j = col0 + c;
    float f = 0.f, lc = CUDART_INF_F;
    if (j < n_pad) {
      lc = __ldg(lse2_col + j);
      f = w_col * ex2f(fminf(__ldg(cw_row + (j >> 6)) - lc, 120.f));
    }
    cfac[c] = f;
    // WSUM: w_col * Pcol * log2 Pcol = e * cfac * (log2 e + c - lc) = cfac * (e log2 e) + dfac * e
    if (WSUM) lcol[c] = (f > 0.f) ? f * (__ldg(cw_row + (min(j, n_pad - 1) >> 6)) - lc) : 0.f;
  }
  if (WSUM && tid < 128) bacc[tid >> 6][tid & 63] = 0.f;
  __syncthreads();
  const float wsum = w_row + w_col;
  // WSUM: this thread's 8 columns of segment s belong to rank r0[s] up to (excluding) element ks[s], to rank
  // r0[s] + 1 from there on (n_per_rank >= 8, so at most one boundary falls into a group)
  float acc_lo_r[4] = {0.f, 0.f, 0.f, 0.f}, acc_lo_c[4] = {0.f, 0.f, 0.f, 0.f};
  int r0[4] = {0, 0, 0, 0}, ks[4] = {8, 8, 8, 8};
  if (WSUM) {
#pragma unroll
    for (int seg = 0; seg < 4; ++seg) {
      const int j = col0 + seg * 256 + lane * 8;
      r0[seg] = j / n_per_rank;
      ks[seg] = min(8, (r0[seg] + 1) * n_per_rank - j);
    }
  }
#pragma unroll 1
  for (int rr = warp; rr < 32; rr += 8) {
    const int i = band * 32 + rr;
    const float lr = (i < m_rows) ? __ldg(lse2_row + i) : CUDART_INF_F;
    const int jd = i + label_offset;   // column of this row's positive
    uint16_t* row = emat + (size_t)i * ld;
    uint4 vin[4];   // all four loads of the row are in flight before any of its arithmetic starts
#pragma unroll
    for (int seg = 0; seg < 4; ++seg) {
      const int j = col0 + seg * 256 + lane * 8;
      vin[seg] = (j < n_pad) ? __ldcs(reinterpret_cast<const uint4*>(row + j)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int seg = 0; seg < 4; ++seg) {
      const int c = seg * 256 + lane * 8;
      const int j = col0 + c;
      if (j < n_pad) {
      uint4 v = vin[seg];
      const float cw = __ldg(cw_row + (j >> 6));
      const float rf = w_row * ex2f(fminf(cw - lr, 120.f));
      const float4 f0 = *reinterpret_cast<const float4*>(cfac + c);
      const float4 f1 = *reinterpret_cast<const float4*>(cfac + c + 4);
      const float cf[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
      float e[8], g[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        e[2 * k] = __uint_as_float(w[k] << 16);
        e[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
      }
      const float2 rf2 = make_float2(rf, rf);
#pragma unroll
      for (int k = 0; k < 4; ++k) {   // packed fp32 pairs: g = e * (rf + cf)
        const float2 gg = fmul2(make_float2(e[2 * k], e[2 * k + 1]),
                                fadd2(rf2, make_float2(cf[2 * k], cf[2 * k + 1])));
        g[2 * k] = gg.x;
        g[2 * k + 1] = gg.y;
      }
      v.x = pack_bf16x2(g[0], g[1]);
      v.y = pack_bf16x2(g[2], g[3]);
      v.z = pack_bf16x2(g[4], g[5]);
      v.w = pack_bf16x2(g[6], g[7]);
      // the positive of row i (at most one per thread-row, so this branch is rare): exact probabilities from diag2
      // replace the rescaled entry -- patched into the packed words and taken out of the sums by correction, which
      // keeps the per-element loops free of selects (and every array in registers)
      const bool diag_here = (jd >= j && jd < j + 8 && i < m_rows);
      float trd = 0.f, tcd = 0.f, ed = 0.f, cfd = 0.f, dfd = 0.f;
      int kd = -1;
      if (diag_here) {
        kd = jd - j;
        const float dg = __ldg(diag2 + i);
        const float lpr = dg - lr, lpc = dg - __ldg(lse2_col + jd);
        const float pr = w_row * ex2f(lpr), pc = w_col * ex2f(lpc);
        trd = pr * lpr;
        tcd = pc * lpc;
        const uint32_t g16 = pack_bf16x2(pr + pc - wsum, 0.f) & 0xffffu;
        const int wd = kd >> 1;
        const uint32_t keep = (kd & 1) ? 0x0000ffffu : 0xffff0000u;
        const uint32_t ins = (kd & 1) ? (g16 << 16) : g16;
        v.x = (wd == 0) ? ((v.x & keep) | ins) : v.x;
        v.y = (wd == 1) ? ((v.y & keep) | ins) : v.y;
        v.z = (wd == 2) ? ((v.z & keep) | ins) : v.z;
        v.w = (wd == 3) ? ((v.w & keep) | ins) : v.w;
        if (WSUM) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            ed = (k == kd) ? e[k] : ed;
            cfd = (k == kd) ? cf[k] : cfd;
          }
        }
      }
      if (WSUM) {
        const float4 l0 = *reinterpret_cast<const float4*>(lcol + c);
        const float4 l1 = *reinterpret_cast<const float4*>(lcol + c + 4);
        const float df[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        if (diag_here) {
#pragma unroll
          for (int k = 0; k < 8; ++k) dfd = (k == kd) ? df[k] : dfd;
        }
        // w * P * log2 P with t = e log2 e (0 for flushed / padded entries):
        //   row: rf * (sum t + (c - lr) * sum e)        col: sum (cfac * t + dfac * e)
        const float dr = (rf > 0.f) ? cw - lr : 0.f;
        const float td = ed * fmaxf(lg2f(ed), -200.f);     // the positive's recovered share, removed below
        if (ks[seg] == 8) {
          float2 st2 = make_float2(0.f, 0.f), se2 = st2, sc2 = st2;
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // packed fp32 pairs (FMUL2 / FADD2 / FFMA2)
            const float2 e2 = make_float2(e[2 * k], e[2 * k + 1]);
            const float2 l2 = make_float2(fmaxf(lg2f(e2.x), -200.f), fmaxf(lg2f(e2.y), -200.f));
            const float2 t2 = fmul2(e2, l2);
            st2 = fadd2(st2, t2);
            se2 = fadd2(se2, e2);
            sc2 = ffma2(make_float2(cf[2 * k], cf[2 * k + 1]), t2, sc2);
            sc2 = ffma2(make_float2(df[2 * k], df[2 * k + 1]), e2, sc2);
          }
          const float st = st2.x + st2.y, se = se2.x + se2.y, sc = sc2.x + sc2.y;
          acc_lo_r[seg] += fmaf(rf, fmaf(dr, se - ed, st - td), trd);
          acc_lo_c[seg] += sc - fmaf(cfd, td, dfd * ed) + tcd;
        } else {   // a rank boundary inside these 8 columns (n not a multiple of 8): element by element
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float t = e[k] * fmaxf(lg2f(e[k]), -200.f);
            const float tr = (k == kd) ? trd : rf * fmaf(e[k], dr, t);
            const float tc = (k == kd) ? tcd : fmaf(cf[k], t, df[k] * e[k]);
            if (k < ks[seg]) {
              acc_lo_r[seg] += tr;
              acc_lo_c[seg] += tc;
            } else if (r0[seg] + 1 < ranks) {   // next owner: straight into the block's table (rare path)
              atomicAdd(&bacc[0][r0[seg] + 1], tr);
              atomicAdd(&bacc[1][r0[seg] + 1], tc);
            }
          }
        }
      }
      *reinterpret_cast<uint4*>(row + j) = v;
      }
    }
  }
  if (WSUM) {
    const int n_all = n_per_rank * ranks;
#pragma unroll
    for (int seg = 0; seg < 4; ++seg) {
      const int j_lo = col0 + seg * 256;
      if (j_lo < n_all) {
      const int r_lo = j_lo / n_per_rank, r_hi = min(j_lo + 255, n_all - 1) / n_per_rank;
      if (r_lo == r_hi) {   // the whole 256-column segment belongs to one rank (block-uniform test)
        const float tr = warp_sum(acc_lo_r[seg]), tc = warp_sum(acc_lo_c[seg]);
        if (lane == 0) {
          atomicAdd(&bacc[0][r_lo], tr);
          atomicAdd(&bacc[1][r_lo], tc);
        }
      } else if (col0 + seg * 256 + lane * 8 < n_all) {   // a rank boundary inside the segment: lane by lane
        atomicAdd(&bacc[0][r0[seg]], acc_lo_r[seg]);
        atomicAdd(&bacc[1][r0[seg]], acc_lo_c[seg]);
      }
      }
    }
    __syncthreads();
    if (tid < 2 * ranks) {
      const int which = tid / ranks, r = tid - which * ranks;
      const float v = bacc[which][r];
      if (v != 0.f) atomicAdd(msums + which * ranks + r, v);
    }
  }
}

// Fallback twin of the WSUM accumulation: when the guard flag is set, G was written by tile_kernel<MODE_GW>
// run with TileParams::ent, whose per-item scalar partials (x = sum Prow log2 Prow, y = sum Pcol log2 Pcol over
// one row block x one column chunk) are attributed to the rank owning the chunk's first column.
__global__ void emat_fallback_sums_kernel(const int* __restrict__ run_if, const float2* __restrict__ sc_part,
                                          int num_rb, int num_chunks, int chunk_cols, int n_per_rank, int ranks,
                                          float w_row, float w_col, float* __restrict__ msums) {
  if (__ldg(run_if) == 0) return;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= num_rb * num_chunks) return;
  const int chunk = item / num_rb;
  const int r = min((chunk * chunk_cols) / n_per_rank, ranks - 1);
  float x = 0.f, y = 0.f;
  for (int w = 0; w < 8; ++w) {
    const float2 v = sc_part[(size_t)item * 8 + w];
    x += v.x;
    y += v.y;
  }
  atomicAdd(msums + r, x * w_row);
  atomicAdd(msums + ranks + r, y * w_col);
}

// out[r, c] = sum_k slots[k][r][c]  (the owner's side of the fused reduce-scatter: W partial slots -> gradient)
__global__ void sum_slots_kernel(const float* __restrict__ slots, int nslots, int rows, int d,
                                 void* __restrict__ out, int out_dtype, long out_ld) {
  const long total = (long)rows * d;
  const long slot_stride = total;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < nslots; ++k) a += slots[k * slot_stride + i];
    const long r = i / d;
    const int c = (int)(i - r * d);
    store_from_float(out, out_dtype, (size_t)(r * out_ld + c), a);
  }
}

// Slot sums of the staged paths (MRCLIP_DS=fwd, MRCLIP_PUSH_DTYPE=bf16): out[r, c] = sum_k slots[k][r][c] like
// sum_slots_kernel, four columns per thread (16-byte loads; requires d % 4 == 0, checked by the caller), slots fp32 or
// bf16, optionally with <out, feat> (fp32, before the output cast; feat = packed bf16 [rows, feat_ld]) accumulated into
// dot_out[blockIdx.x % 64] -- the text-gradient side of  s dL_r/ds = <dT_r, T_r> + ...  (tile_kernel.cuh, MODE_FWDEU).
template <bool SLOTS_BF16>
__global__ void sum_slots_vec_kernel(const void* __restrict__ slots, int nslots, int rows, int d,
                                     void* __restrict__ out, int out_dtype, long out_ld,
                                     const __nv_bfloat16* __restrict__ feat, long feat_ld, float* __restrict__ dot_out) {
  const int dq = d >> 2;
  const long total4 = (long)rows * dq;
  const long slot_stride = (long)rows * d;
  float dot = 0.f;
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

// R2(me, q) = sum over my rows i and the columns j owned by rank q of Prow_ij * S2_ij, from the forward's per-slot
// (max2, sum) and u partials (tile_kernel.cuh, MODE_FWDEU): out[(blockIdx.x % 64) * 2 * ranks + q] += block sum.
// A slot is one half of a column chunk; slots_per_rank consecutive slots belong to one owner.
__global__ void row_ent_split_kernel(const float2* __restrict__ row_part, const float* __restrict__ row_ent, int slots,
                                     int slots_per_rank, int ranks, int m_rows, int m_pad,
                                     const float* __restrict__ lse2_row, float* __restrict__ out) {
  extern __shared__ float acc[];   // [ranks]
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
  for (int q = threadIdx.x; q < ranks; q += blockDim.x)
    atomicAdd(out + (size_t)(blockIdx.x & 63) * 2 * ranks + q, acc[q]);
}

// ---- MultiPositiveClipLoss (reference loss.py:626-644, :671-747): class statistics of the packed features -----------
// mean[c] = average of the rows of x (bf16 [N, ld]) whose class is c.  order lists the sample indices grouped by class,
// seg_start[c] / seg_cnt[c] delimit class c inside it (seg_cnt[c] == 0: unused id).  One block per class id, a thread
// per pair of columns; replaces two index_add_ passes over fp32 copies of the features.
__global__ void class_mean_kernel(const __nv_bfloat16* __restrict__ x, int ld, const int* __restrict__ order,
                                  const int* __restrict__ seg_start, const int* __restrict__ seg_cnt, float* __restrict__ mean) {
  const int c = blockIdx.x, cnt = seg_cnt[c];
  if (cnt <= 0) return;
  const int start = seg_start[c];
  const float inv = 1.f / (float)cnt;
  for (int c2 = threadIdx.x; c2 < ld / 2; c2 += blockDim.x) {
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < cnt; ++j) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(x + (size_t)__ldg(order + start + j) * ld + 2 * c2);
      a0 += __uint_as_float(v << 16);
      a1 += __uint_as_float(v & 0xffff0000u);
    }
    *reinterpret_cast<float2*>(mean + (size_t)c * ld + 2 * c2) = make_float2(a0 * inv, a1 * inv);
  }
}

// loss_out[0] += (1/n) sum_i ( delta (lse_row_i - s <I_i, tmean[c_i]>) + (1 - delta) (lse_col_i - s <T_i, imean[c_i]>) ),
// c_i = cls[i]; the mean over the positives of S_ij is s <I_i, mean_{P(i)} T> (natural-log units; lse2 are log2).
// One warp per row.  loss_out must be zero on entry.
__global__ void mpos_forward_kernel(const __nv_bfloat16* __restrict__ img, const __nv_bfloat16* __restrict__ txt, int ld, int n,
                                    int d, const int* __restrict__ cls, const float* __restrict__ tmean,
                                    const float* __restrict__ imean, const float* __restrict__ lse2_row,
                                    const float* __restrict__ lse2_col, const float* __restrict__ scale, float delta,
                                    float* __restrict__ loss_out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float s = scale[0];
  float acc = 0.f;
  for (long i = blockIdx.x * (long)wpb + (threadIdx.x >> 5); i < n; i += (long)gridDim.x * wpb) {
    const size_t c = (size_t)cls[i] * ld;
    float d1 = 0.f, d2 = 0.f;
    for (int k = lane; k < d; k += 32) {
      d1 = fmaf(__bfloat162float(img[i * ld + k]), tmean[c + k], d1);
      d2 = fmaf(__bfloat162float(txt[i * ld + k]), imean[c + k], d2);
    }
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    if (lane == 0)
      acc += delta * (lse2_row[i] * 0.6931471805599453f - s * d1) + (1.f - delta) * (lse2_col[i] * 0.6931471805599453f - s * d2);
  }
  __shared__ float red[32];
  if (lane == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < wpb) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0 && t != 0.f) atomicAdd(loss_out, t / (float)n);
  }
}

// d_img[i] += k (T_i - tmean[c_i]),  d_txt[i] += k (I_i - imean[c_i]),  k = coef * scale * grad_out: the difference between
// the multi-positive gradient of the logits and what the E-block pipeline contracts (G_k + delta_ij - [same class]/|P|).
__global__ void mpos_backward_kernel(void* __restrict__ d_img, int di_dtype, long di_ld, void* __restrict__ d_txt, int dt_dtype,
                                     long dt_ld, const __nv_bfloat16* __restrict__ img, const __nv_bfloat16* __restrict__ txt,
                                     int ld, int n, int d, const int* __restrict__ cls, const float* __restrict__ tmean,
                                     const float* __restrict__ imean, float coef, const float* __restrict__ scale,
                                     const float* __restrict__ grad_out) {
  const float k = coef * scale[0] * (grad_out != nullptr ? grad_out[0] : 1.f);
  const long total = (long)n * d;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long i = e / d;
    const int col = (int)(e - i * d);
    const size_t c = (size_t)cls[i] * ld + col;
    const size_t gi = (size_t)(i * di_ld + col), gt = (size_t)(i * dt_ld + col);
    store_from_float(d_img, di_dtype, gi, load_as_float(d_img, di_dtype, gi) + k * (__bfloat162float(txt[i * ld + col]) - tmean[c]));
    store_from_float(d_txt, dt_dtype, gt, load_as_float(d_txt, dt_dtype, gt) + k * (__bfloat162float(img[i * ld + col]) - imean[c]));
  }
}

// dot_out += <out, feat> / scale over an [rows, d] block (out: any float dtype, feat: packed bf16); one warp per row.
// With out = dA and feat = A this is d(loss)/d(logit_scale), by homogeneity of S = scale * A.B^T.
__global__ void rowdot_kernel(const void* __restrict__ out, int out_dtype, long out_ld, const __nv_bfloat16* __restrict__ feat,
                              long feat_ld, int rows, int d, const float* __restrict__ scale, float* __restrict__ dot_out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float acc = 0.f;
  for (long r = blockIdx.x * (long)wpb + (threadIdx.x >> 5); r < rows; r += (long)gridDim.x * wpb)
    for (int c = lane; c < d; c += 32)
      acc = fmaf(load_as_float(out, out_dtype, (size_t)(r * out_ld + c)), __bfloat162float(feat[r * feat_ld + c]), acc);
  acc = warp_sum(acc);
  __shared__ float red[32];
  if (lane == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < wpb) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(dot_out, t / scale[0]);
  }
}

// retrieval metrics: lmax[i] = largest entry of row i's positive list (one warp per row)
__global__ void rank_lmax_kernel(const float* __restrict__ pos, const long long* __restrict__ off, const int* __restrict__ m,
                                 int rows, float* __restrict__ lmax) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long r = blockIdx.x * (long)wpb + (threadIdx.x >> 5); r < rows; r += (long)gridDim.x * wpb) {
    const float* p = pos + off[r];
    float v = -CUDART_INF_F;
    for (int t = lane; t < m[r]; t += 32) v = fmaxf(v, p[t]);
    v = warp_max(v);
    if (lane == 0) lmax[r] = v;
  }
}

// all-gather by peer stores: copy `bytes` (multiple of 16) from src to dst[k] + offset for every k != skip.
// dsts are NVLink-mapped addresses of the same buffer on every rank; a warp writes 512 contiguous bytes, so the
// link sees full 128-byte packets.  grid.y = destination.
__global__ void push_copy_kernel(const uint4* __restrict__ src, long n16, const unsigned long long* __restrict__ dsts,
                                 long offset_bytes, int skip) {
  const int k = blockIdx.y;
  if (k == skip) return;
  uint4* dst = reinterpret_cast<uint4*>(__ldg(dsts + k) + offset_bytes);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n16; i += (long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

}  // namespace mrclip
