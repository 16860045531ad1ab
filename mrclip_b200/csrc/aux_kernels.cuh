// Small HBM-bound helper kernels around the tile kernel: packing features to bf16, the
// [N,D] -> [D,N] transpose that gives the second GEMM a K-major operand, and the reductions
// that turn per-tile partials into LSE vectors, losses and gradients.
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
// block = 32 columns x 32 band slices (1024 threads); slices are merged through shared memory.
__global__ void reduce_cols_kernel(const float* __restrict__ col_l, const float* __restrict__ col_c,
                                   int bands, int n_cols, int n_pad, float* __restrict__ out_m,
                                   float* __restrict__ out_l) {
  __shared__ float sm[32][33], sl[32][33];
  const int c = threadIdx.x, slice = threadIdx.y;
  const int j = blockIdx.x * 32 + c;
  float m = -CUDART_INF_F, l = 0.f;
  if (j < n_cols) {
    for (int b = slice; b < bands; b += 32) {
      const float ref = col_c[(size_t)b * (n_pad / 64) + j / 64];
      const float v = col_l[(size_t)b * n_pad + j];
      lse2_merge(m, l, ref, v);
    }
  }
  sm[slice][c] = m;
  sl[slice][c] = l;
  __syncthreads();
  if (slice == 0 && j < n_cols) {
    for (int k = 1; k < 32; ++k) lse2_merge(m, l, sm[k][c], sl[k][c]);
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
                                   int out_dtype, long out_ld) {
  float mul = coef * scale[0];
  if (grad_out) mul *= grad_out[0];
  const int dq = d_pad / 4;
  const long total = (long)m_rows * dq;
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
      if (c + j < d) store_from_float(out, out_dtype, (size_t)(r * out_ld + c + j), o[j]);
  }
}

__global__ void fill_kernel(float* __restrict__ p, long n, float v) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    p[i] = v;
}

}  // namespace mrclip
