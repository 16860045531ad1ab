// Standalone GPU self-test / micro-benchmark of the C ABI (no Python, no torch).
//   selftest check            : small/medium shapes vs. an fp64 host computation of the same math
//   selftest time N D [reps]  : CUDA-event timing of fwd / bwd passes at one shape (W=1)
// The fp64 host code below is a test fixture local to this binary; the repository oracle lives
// in oracle/.
#include "../../include/mrclip.h"
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)
#define MR(x)                                                              \
  do {                                                                     \
    int r_ = (x);                                                          \
    if (r_ != 0) {                                                         \
      printf("mrclip error %d: %s (%s:%d)\n", r_, mrclip_last_error(), __FILE__, __LINE__); \
      exit(3);                                                             \
    }                                                                      \
  } while (0)

static float bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  const uint32_t r = 0x7FFFu + ((u >> 16) & 1u);
  u = (u + r) & 0xFFFF0000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

struct Problem {
  int N, D, W, n;
  std::vector<float> I, T;  // [N, D], bf16-representable
};

static Problem make_problem(int N, int D, int W, unsigned seed, float corr) {
  Problem p;
  p.N = N;
  p.D = D;
  p.W = W;
  p.n = N / W;
  p.I.resize((size_t)N * D);
  p.T.resize((size_t)N * D);
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  for (int i = 0; i < N; ++i) {
    double ni = 0, nt = 0;
    for (int k = 0; k < D; ++k) {
      const float a = nd(rng), b = nd(rng);
      p.I[(size_t)i * D + k] = a;
      p.T[(size_t)i * D + k] = corr * a + (1.f - corr) * b;
      ni += (double)a * a;
    }
    for (int k = 0; k < D; ++k) nt += (double)p.T[(size_t)i * D + k] * p.T[(size_t)i * D + k];
    for (int k = 0; k < D; ++k) {
      p.I[(size_t)i * D + k] = bf16_round((float)(p.I[(size_t)i * D + k] / sqrt(ni)));
      p.T[(size_t)i * D + k] = bf16_round((float)(p.T[(size_t)i * D + k] / sqrt(nt)));
    }
  }
  return p;
}

struct DevBufs {
  void *Ibf, *Tbf, *It, *Tt, *ws;
  float *If32, *Tf32, *scale, *bias, *gout;
  int ld, npad;
  size_t ws_bytes;
};

static double rel_err(const std::vector<double>& ref, const std::vector<float>& got) {
  double num = 0, den = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    const double d = ref[i] - (double)got[i];
    num += d * d;
    den += ref[i] * ref[i];
  }
  return sqrt(num / (den > 0 ? den : 1));
}

static int g_fail = 0;
static void report(const char* what, double err, double tol) {
  const bool ok = err <= tol && err == err;
  printf("  %-28s err %.3e (tol %.1e) %s\n", what, err, tol, ok ? "ok" : "FAIL");
  if (!ok) g_fail = 1;
}

static DevBufs setup(const Problem& p, float s, float b) {
  DevBufs d;
  d.ld = mrclip_padded_dim(p.D);
  d.npad = mrclip_padded_cols(p.N);
  const size_t fe = (size_t)p.N * p.D;
  CK(cudaMalloc(&d.If32, fe * 4));
  CK(cudaMalloc(&d.Tf32, fe * 4));
  CK(cudaMemcpy(d.If32, p.I.data(), fe * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.Tf32, p.T.data(), fe * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d.Ibf, (size_t)p.N * d.ld * 2));
  CK(cudaMalloc(&d.Tbf, (size_t)p.N * d.ld * 2));
  CK(cudaMalloc(&d.It, (size_t)d.ld * d.npad * 2));
  CK(cudaMalloc(&d.Tt, (size_t)d.ld * d.npad * 2));
  CK(cudaMemset(d.It, 0, (size_t)d.ld * d.npad * 2));
  CK(cudaMemset(d.Tt, 0, (size_t)d.ld * d.npad * 2));
  d.ws_bytes = mrclip_workspace_bytes(p.n, p.N, p.D);
  CK(cudaMalloc(&d.ws, d.ws_bytes));
  CK(cudaMalloc(&d.scale, 4));
  CK(cudaMalloc(&d.bias, 4));
  CK(cudaMalloc(&d.gout, 4));
  const float one = 1.f;
  CK(cudaMemcpy(d.scale, &s, 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.bias, &b, 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.gout, &one, 4, cudaMemcpyHostToDevice));
  MR(mrclip_pack_bf16(d.If32, MRCLIP_DT_F32, p.N, p.D, p.D, d.Ibf, d.ld, 0));
  MR(mrclip_pack_bf16(d.Tf32, MRCLIP_DT_F32, p.N, p.D, p.D, d.Tbf, d.ld, 0));
  MR(mrclip_transpose_bf16(d.Ibf, p.N, d.ld, d.ld, d.It, d.npad, 0));
  MR(mrclip_transpose_bf16(d.Tbf, p.N, d.ld, d.ld, d.Tt, d.npad, 0));
  CK(cudaDeviceSynchronize());
  return d;
}
static void teardown(DevBufs& d) {
  cudaFree(d.If32); cudaFree(d.Tf32); cudaFree(d.Ibf); cudaFree(d.Tbf); cudaFree(d.It); cudaFree(d.Tt);
  cudaFree(d.ws); cudaFree(d.scale); cudaFree(d.bias); cudaFree(d.gout);
}

static void check_clip(int N, int D, int W, float s, float corr) {
  printf("[clip] N=%d D=%d W=%d scale=%.2f corr=%.2f\n", N, D, W, s, corr);
  Problem p = make_problem(N, D, W, 1234u + N + D, corr);
  DevBufs d = setup(p, s, 0.f);
  const int n = p.n;
  // ---------------- host fp64 reference
  std::vector<double> C((size_t)N * N);
#pragma omp parallel for
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double a = 0;
      for (int k = 0; k < D; ++k) a += (double)p.I[(size_t)i * D + k] * p.T[(size_t)j * D + k];
      C[(size_t)i * N + j] = a;
    }
  std::vector<double> lr(N), lc(N);
  for (int i = 0; i < N; ++i) {
    double m = -1e300;
    for (int j = 0; j < N; ++j) m = fmax(m, s * C[(size_t)i * N + j]);
    double l = 0;
    for (int j = 0; j < N; ++j) l += exp(s * C[(size_t)i * N + j] - m);
    lr[i] = m + log(l);
  }
  for (int j = 0; j < N; ++j) {
    double m = -1e300;
    for (int i = 0; i < N; ++i) m = fmax(m, s * C[(size_t)i * N + j]);
    double l = 0;
    for (int i = 0; i < N; ++i) l += exp(s * C[(size_t)i * N + j] - m);
    lc[j] = m + log(l);
  }
  std::vector<double> dI((size_t)N * D, 0.0), dT((size_t)N * D, 0.0), dS(W, 0.0), L(W, 0.0);
  const double coef = 1.0 / (2.0 * n);
#pragma omp parallel for
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      const double sij = s * C[(size_t)i * N + j];
      const double g = coef * (exp(sij - lr[i]) + exp(sij - lc[j]) - (i == j ? 2.0 : 0.0));
      for (int k = 0; k < D; ++k) dI[(size_t)i * D + k] += s * g * p.T[(size_t)j * D + k];
    }
#pragma omp parallel for
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < N; ++i) {
      const double sij = s * C[(size_t)i * N + j];
      const double g = coef * (exp(sij - lr[i]) + exp(sij - lc[j]) - (i == j ? 2.0 : 0.0));
      for (int k = 0; k < D; ++k) dT[(size_t)j * D + k] += s * g * p.I[(size_t)i * D + k];
    }
  for (int r = 0; r < W; ++r) {
    double acc = 0, ds = 0;
    for (int i = r * n; i < (r + 1) * n; ++i) {
      acc += lr[i] + lc[i] - 2 * s * C[(size_t)i * N + i];
      for (int j = 0; j < N; ++j) {
        ds += coef * (exp(s * C[(size_t)i * N + j] - lr[i]) - (i == j)) * C[(size_t)i * N + j];
        ds += coef * (exp(s * C[(size_t)j * N + i] - lc[i]) - (i == j)) * C[(size_t)j * N + i];
      }
    }
    L[r] = acc * coef;
    dS[r] = ds;
  }
  // ---------------- device
  float *lse2_row_all, *col_m, *col_l, *lse2_col_all, *diag2, *loss, *dscale, *dA;
  CK(cudaMalloc(&lse2_row_all, d.npad * 4));
  CK(cudaMalloc(&col_m, (size_t)W * N * 4));
  CK(cudaMalloc(&col_l, (size_t)W * N * 4));
  CK(cudaMalloc(&lse2_col_all, d.npad * 4));
  CK(cudaMalloc(&diag2, (size_t)N * 4));
  CK(cudaMalloc(&loss, W * 4));
  CK(cudaMalloc(&dscale, W * 4));
  CK(cudaMalloc(&dA, (size_t)N * D * 4));
  {
    std::vector<float> inf(d.npad, INFINITY);
    CK(cudaMemcpy(lse2_row_all, inf.data(), d.npad * 4, cudaMemcpyHostToDevice));
  }
  for (int r = 0; r < W; ++r) {
    mrclip_shape sh = {n, N, D, r * n};
    const char* a = (const char*)d.Ibf + (size_t)r * n * d.ld * 2;
    const int g = mrclip_fwd_col_granule(n, N);
    // exercise the split launch when the granule allows it
    int mid = (N / 2) / g * g;
    if (mid > 0 && mid < N) {
      MR(mrclip_clip_fwd_tiles(a, d.Tbf, sh, d.ld, d.scale, mid, N, d.ws, 0));
      MR(mrclip_clip_fwd_tiles(a, d.Tbf, sh, d.ld, d.scale, 0, mid, d.ws, 0));
    } else {
      MR(mrclip_clip_fwd_tiles(a, d.Tbf, sh, d.ld, d.scale, 0, N, d.ws, 0));
    }
    MR(mrclip_clip_fwd_reduce(sh, d.ws, lse2_row_all + r * n, col_m + (size_t)r * N, col_l + (size_t)r * N,
                              diag2 + r * n, 0));
    CK(cudaDeviceSynchronize());
  }
  MR(mrclip_lse2_merge(col_m, col_l, W, N, N, lse2_col_all, 0));
  for (int r = 0; r < W; ++r)
    MR(mrclip_clip_loss(lse2_row_all + r * n, lse2_col_all, diag2 + r * n, n, r * n, loss + r, 0));
  CK(cudaDeviceSynchronize());
  {
    std::vector<float> h(N), hl(W);
    std::vector<double> ref(N);
    CK(cudaMemcpy(h.data(), lse2_row_all, N * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < N; ++i) ref[i] = lr[i] * 1.4426950408889634;
    report("lse2_row", rel_err(ref, h), 1e-5);
    CK(cudaMemcpy(h.data(), lse2_col_all, N * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < N; ++i) ref[i] = lc[i] * 1.4426950408889634;
    report("lse2_col", rel_err(ref, h), 1e-5);
    CK(cudaMemcpy(h.data(), diag2, N * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < N; ++i) ref[i] = s * C[(size_t)i * N + i] * 1.4426950408889634;
    report("diag2", rel_err(ref, h), 1e-5);
    CK(cudaMemcpy(hl.data(), loss, W * 4, cudaMemcpyDeviceToHost));
    report("loss per rank", rel_err(L, hl), 1e-4);
    printf("    loss[0] = %.6f (ref %.6f)\n", hl[0], L[0]);
  }
  // backward
  std::vector<float> hI((size_t)N * D), hT((size_t)N * D), hds(W);
  for (int r = 0; r < W; ++r) {
    mrclip_shape sh = {n, N, D, r * n};
    const char* ai = (const char*)d.Ibf + (size_t)r * n * d.ld * 2;
    MR(mrclip_clip_bwd(ai, d.Tbf, d.Tt, d.npad, sh, d.ld, lse2_row_all + r * n, lse2_col_all, d.scale, 1.f, 1.f,
                       (float)coef, d.gout, d.ws, dA + (size_t)r * n * D, MRCLIP_DT_F32, D, dscale + r, 0, 0));
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(hI.data(), dA, (size_t)N * D * 4, cudaMemcpyDeviceToHost));
  for (int r = 0; r < W; ++r) {
    mrclip_shape sh = {n, N, D, r * n};
    const char* at = (const char*)d.Tbf + (size_t)r * n * d.ld * 2;
    MR(mrclip_clip_bwd(at, d.Ibf, d.It, d.npad, sh, d.ld, lse2_col_all + r * n, lse2_row_all, d.scale, 1.f, 1.f,
                       (float)coef, d.gout, d.ws, dA + (size_t)r * n * D, MRCLIP_DT_F32, D, dscale + r, 1, 0));
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(hT.data(), dA, (size_t)N * D * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hds.data(), dscale, W * 4, cudaMemcpyDeviceToHost));
  report("dI", rel_err(dI, hI), 1e-2);
  report("dT", rel_err(dT, hT), 1e-2);
  report("d_scale per rank", rel_err(dS, hds), 1e-2);
  printf("    ds[0] = %.6e (ref %.6e)\n", hds[0], dS[0]);
  // ---- gmat backend: G written once per pass, gradients as plain GEMMs
  {
    void* gmat;
    CK(cudaMalloc(&gmat, mrclip_gmat_bytes(n, N)));
    CK(cudaMemset(gmat, 0xff, mrclip_gmat_bytes(n, N)));
    for (int r = 0; r < W; ++r) {
      mrclip_shape sh = {n, N, D, r * n};
      const char* ai = (const char*)d.Ibf + (size_t)r * n * d.ld * 2;
      const char* at = (const char*)d.Tbf + (size_t)r * n * d.ld * 2;
      MR(mrclip_clip_gwrite(ai, d.Tbf, sh, d.ld, lse2_row_all + r * n, lse2_col_all, d.scale, 1.f, 1.f, (float)coef,
                            d.gout, d.ws, gmat, dscale + r, 0, W == 1 ? 1 : 0, 0));
      MR(mrclip_gmat_gemm(0, gmat, sh, d.Tbf, d.ld, (float)coef, d.scale, d.gout, d.ws,
                          dA + (size_t)r * n * D, MRCLIP_DT_F32, D, 0));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hI.data() + (size_t)r * n * D, dA + (size_t)r * n * D, (size_t)n * D * 4, cudaMemcpyDeviceToHost));
      if (W == 1) {
        MR(mrclip_gmat_gemm(1, gmat, sh, d.Ibf, d.ld, (float)coef, d.scale, d.gout, d.ws, dA, MRCLIP_DT_F32, D, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hT.data(), dA, (size_t)N * D * 4, cudaMemcpyDeviceToHost));
        report("gmat dT (transposed G)", rel_err(dT, hT), 1e-2);
      } else {
        MR(mrclip_clip_gwrite(at, d.Ibf, sh, d.ld, lse2_col_all + r * n, lse2_row_all, d.scale, 1.f, 1.f, (float)coef,
                              d.gout, d.ws, gmat, dscale + r, 1, 0, 0));
        MR(mrclip_gmat_gemm(0, gmat, sh, d.Ibf, d.ld, (float)coef, d.scale, d.gout, d.ws,
                            dA + (size_t)r * n * D, MRCLIP_DT_F32, D, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hT.data() + (size_t)r * n * D, dA + (size_t)r * n * D, (size_t)n * D * 4, cudaMemcpyDeviceToHost));
      }
    }
    CK(cudaMemcpy(hds.data(), dscale, W * 4, cudaMemcpyDeviceToHost));
    report("gmat dI", rel_err(dI, hI), 1e-2);
    if (W > 1) report("gmat dT (column pass)", rel_err(dT, hT), 1e-2);
    report("gmat d_scale per rank", rel_err(dS, hds), 1e-2);
    cudaFree(gmat);
  }
  // ---- emat backend: forward keeps E, the GEMMs rebuild G in shared memory; dT = sum over ranks of G_r^T . I_r
  {
    void* emat;
    float *dTp, *tmp_m, *tmp_l, *tmp_lr, *tmp_dg, *msums;
    CK(cudaMalloc(&msums, (size_t)2 * W * W * 4));
    CK(cudaMalloc(&emat, mrclip_gmat_bytes(n, N)));
    CK(cudaMemset(emat, 0xff, mrclip_gmat_bytes(n, N)));
    CK(cudaMalloc(&dTp, (size_t)N * D * 4));
    CK(cudaMalloc(&tmp_m, (size_t)N * 4));
    CK(cudaMalloc(&tmp_l, (size_t)N * 4));
    CK(cudaMalloc(&tmp_lr, (size_t)N * 4));
    CK(cudaMalloc(&tmp_dg, (size_t)N * 4));
    std::vector<double> accT((size_t)N * D, 0.0);
    std::vector<float> hp((size_t)N * D);
    int flags = 0;
    for (int r = 0; r < W; ++r) {
      mrclip_shape sh = {n, N, D, r * n};
      const char* ai = (const char*)d.Ibf + (size_t)r * n * d.ld * 2;
      MR(mrclip_clip_fwd_tiles_e(ai, d.Tbf, sh, d.ld, d.scale, 0, N, d.ws, emat, 0));
      MR(mrclip_clip_fwd_reduce(sh, d.ws, tmp_lr, tmp_m, tmp_l, tmp_dg, 0));
      MR(mrclip_emat_check(sh, d.ws, lse2_row_all + r * n, lse2_col_all, 0));
      const int* flag = mrclip_emat_flag(sh, d.ws);
      MR(mrclip_clip_gwrite_if(ai, d.Tbf, sh, d.ld, lse2_row_all + r * n, lse2_col_all, d.scale, 1.f, 1.f, d.ws, emat,
                               flag, 0));
      CK(cudaMemset(dscale + r, 0, 4));
      MR(mrclip_emat_transform(sh, d.ws, emat, lse2_row_all + r * n, lse2_col_all, diag2 + r * n, d.scale, 1.f, 1.f, flag,
                               msums + 2 * W * r, 1, n, W, 0));
      MR(mrclip_gmat_gemm_dot(0, emat, sh, d.Tbf, d.ld, (float)coef, d.scale, d.gout, d.ws, dA + (size_t)r * n * D,
                              MRCLIP_DT_F32, D, ai, dscale + r, 0));
      MR(mrclip_gmat_gemm(1, emat, sh, ai, d.ld, (float)coef, d.scale, d.gout, d.ws, dTp, MRCLIP_DT_F32, D, 0));
      CK(cudaDeviceSynchronize());
      int hf = 0;
      CK(cudaMemcpy(&hf, flag, 4, cudaMemcpyDeviceToHost));
      flags += hf;
      CK(cudaMemcpy(hI.data() + (size_t)r * n * D, dA + (size_t)r * n * D, (size_t)n * D * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hp.data(), dTp, (size_t)N * D * 4, cudaMemcpyDeviceToHost));
      for (size_t k = 0; k < accT.size(); ++k) accT[k] += hp[k];
    }
    for (size_t k = 0; k < accT.size(); ++k) hT[k] = (float)accT[k];
    report("emat dI", rel_err(dI, hI), 1e-2);
    report("emat dT (sum of rank partials)", rel_err(dT, hT), 1e-2);
    if (W == 1) {
      CK(cudaMemcpy(hds.data(), dscale, W * 4, cudaMemcpyDeviceToHost));
      report("emat d_scale (<dI,I>/s)", rel_err(dS, hds), 1e-2);
      printf("    ds[0] = %.6e (ref %.6e)\n", hds[0], dS[0]);
    }
    {   // d_scale per rank from the split sums: row r of the Prow matrix + column r of the Pcol matrix
      std::vector<float> hm((size_t)2 * W * W);
      CK(cudaMemcpy(hm.data(), msums, hm.size() * 4, cudaMemcpyDeviceToHost));
      std::vector<float> ds2(W);
      for (int r = 0; r < W; ++r) {
        double a = 0;
        for (int c2 = 0; c2 < W; ++c2) a += hm[(size_t)2 * W * r + c2];            // Prow log2 Prow, my rows, all column ranks
        for (int q = 0; q < W; ++q) a += hm[(size_t)2 * W * q + W + r];            // Pcol log2 Pcol, all row ranks, my columns
        // scale * dL_r/dscale = L_r + ln2/(2n) * (negative entropies)
        ds2[r] = (float)((L[r] + 0.6931471805599453 * coef * a) / s);
      }
      {   // the sums themselves against fp64 (single rank view: totals over all ranks)
        double hr = 0, hc = 0, gr = 0, gc = 0;
        for (int i = 0; i < N; ++i)
          for (int j = 0; j < N; ++j) {
            const double sij = s * C[(size_t)i * N + j];
            const double lpr = (sij - lr[i]) * 1.4426950408889634, lpc = (sij - lc[j]) * 1.4426950408889634;
            hr += exp2(lpr) * lpr;
            hc += exp2(lpc) * lpc;
          }
        for (int q = 0; q < W; ++q)
          for (int c2 = 0; c2 < W; ++c2) {
            gr += hm[(size_t)2 * W * q + c2];
            gc += hm[(size_t)2 * W * q + W + c2];
          }
        printf("    sum P log2 P: rows %.6f (ref %.6f)  cols %.6f (ref %.6f)\n", gr, hr, gc, hc);
      }
      report("emat d_scale per rank (split sums)", rel_err(dS, ds2), 1e-2);
      printf("    ds[0] = %.6e (ref %.6e)\n", ds2[0], dS[0]);
    }
    printf("    emat guard flag raised on %d of %d ranks\n", flags, W);
    cudaFree(msums);
    cudaFree(emat); cudaFree(dTp); cudaFree(tmp_m); cudaFree(tmp_l); cudaFree(tmp_lr); cudaFree(tmp_dg);
  }
  cudaFree(lse2_row_all); cudaFree(col_m); cudaFree(col_l); cudaFree(lse2_col_all); cudaFree(diag2);
  cudaFree(loss); cudaFree(dscale); cudaFree(dA);
  teardown(d);
}

static void check_siglip(int N, int D, int W, float s, float b) {
  printf("[siglip] N=%d D=%d W=%d scale=%.2f bias=%.2f\n", N, D, W, s, b);
  Problem p = make_problem(N, D, W, 99u + N + D, 0.5f);
  DevBufs d = setup(p, s, b);
  const int n = p.n;
  std::vector<double> dI((size_t)N * D, 0.0), dT((size_t)N * D, 0.0), dS(W, 0.0), dB(W, 0.0), L(W, 0.0);
  std::vector<double> G((size_t)N * N), C((size_t)N * N);
#pragma omp parallel for
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double a = 0;
      for (int k = 0; k < D; ++k) a += (double)p.I[(size_t)i * D + k] * p.T[(size_t)j * D + k];
      C[(size_t)i * N + j] = a;
      const double z = s * a + b;
      G[(size_t)i * N + j] = (1.0 / (1.0 + exp(-z)) - (i == j ? 1.0 : 0.0)) / n;
    }
  for (int r = 0; r < W; ++r)
    for (int i = r * n; i < (r + 1) * n; ++i)
      for (int j = 0; j < N; ++j) {
        const double z = s * C[(size_t)i * N + j] + b;
        const double y = (i == j) ? 1.0 : -1.0;
        const double x = -y * z;
        L[r] += (fmax(x, 0.0) + log1p(exp(-fabs(x)))) / n;
        dS[r] += G[(size_t)i * N + j] * C[(size_t)i * N + j];
        dB[r] += G[(size_t)i * N + j];
      }
#pragma omp parallel for
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < D; ++k) dI[(size_t)i * D + k] += s * G[(size_t)i * N + j] * p.T[(size_t)j * D + k];
#pragma omp parallel for
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < D; ++k) dT[(size_t)j * D + k] += s * G[(size_t)i * N + j] * p.I[(size_t)i * D + k];
  float *loss, *dscale, *dbias, *dA;
  CK(cudaMalloc(&loss, W * 4));
  CK(cudaMalloc(&dscale, W * 4));
  CK(cudaMalloc(&dbias, W * 4));
  CK(cudaMalloc(&dA, (size_t)N * D * 4));
  std::vector<float> hI((size_t)N * D), hT((size_t)N * D), hs(W), hb(W), hl(W);
  for (int r = 0; r < W; ++r) {
    mrclip_shape sh = {n, N, D, r * n};
    const char* ai = (const char*)d.Ibf + (size_t)r * n * d.ld * 2;
    MR(mrclip_siglip_fwd(ai, d.Tbf, sh, d.ld, d.scale, d.bias, d.ws, loss + r, 0));
    CK(cudaDeviceSynchronize());
    MR(mrclip_siglip_bwd(ai, d.Tbf, d.Tt, d.npad, sh, d.ld, d.scale, d.bias, 1.f / n, d.gout, d.ws,
                         dA + (size_t)r * n * D, MRCLIP_DT_F32, D, dscale + r, dbias + r, 0, 0));
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(hI.data(), dA, (size_t)N * D * 4, cudaMemcpyDeviceToHost));
  for (int r = 0; r < W; ++r) {
    mrclip_shape sh = {n, N, D, r * n};
    const char* at = (const char*)d.Tbf + (size_t)r * n * d.ld * 2;
    MR(mrclip_siglip_bwd(at, d.Ibf, d.It, d.npad, sh, d.ld, d.scale, d.bias, 1.f / n, d.gout, d.ws,
                         dA + (size_t)r * n * D, MRCLIP_DT_F32, D, nullptr, nullptr, 0, 0));
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(hT.data(), dA, (size_t)N * D * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hs.data(), dscale, W * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), dbias, W * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hl.data(), loss, W * 4, cudaMemcpyDeviceToHost));
  report("loss per rank", rel_err(L, hl), 1e-4);
  printf("    loss[0] = %.6f (ref %.6f)\n", hl[0], L[0]);
  report("dI", rel_err(dI, hI), 1e-2);
  report("dT", rel_err(dT, hT), 1e-2);
  report("d_scale", rel_err(dS, hs), 1e-2);
  report("d_bias", rel_err(dB, hb), 1e-2);
  cudaFree(loss); cudaFree(dscale); cudaFree(dbias); cudaFree(dA);
  teardown(d);
}

static void time_shape(int N, int D, int reps) {
  printf("[time] N=%d D=%d reps=%d\n", N, D, reps);
  Problem p;
  p.N = N; p.D = D; p.W = 1; p.n = N;
  p.I.resize((size_t)N * D);
  p.T.resize((size_t)N * D);
  std::mt19937 rng(7);
  std::normal_distribution<float> nd(0.f, 1.f);
  const float sc = 1.f / sqrtf((float)D);
  for (size_t i = 0; i < p.I.size(); ++i) {
    p.I[i] = bf16_round(nd(rng) * sc);
    p.T[i] = bf16_round(0.5f * p.I[i] + 0.5f * nd(rng) * sc);
  }
  DevBufs d = setup(p, 14.285714f, 0.f);
  float *lse2_row, *col_m, *col_l, *lse2_col, *diag2, *loss, *dscale, *dA;
  CK(cudaMalloc(&lse2_row, d.npad * 4));
  CK(cudaMalloc(&col_m, (size_t)N * 4));
  CK(cudaMalloc(&col_l, (size_t)N * 4));
  CK(cudaMalloc(&lse2_col, d.npad * 4));
  CK(cudaMalloc(&diag2, (size_t)N * 4));
  CK(cudaMalloc(&loss, 4));
  CK(cudaMalloc(&dscale, 4));
  CK(cudaMalloc(&dA, (size_t)N * D * 4));
  {
    std::vector<float> inf(d.npad, INFINITY);
    CK(cudaMemcpy(lse2_row, inf.data(), d.npad * 4, cudaMemcpyHostToDevice));
  }
  mrclip_shape sh = {N, N, D, 0};
  cudaEvent_t e0, e1, e2, e3;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
  float tf = 0, tb1 = 0, tb2 = 0;
  for (int it = 0; it < reps + 2; ++it) {
    CK(cudaEventRecord(e0));
    MR(mrclip_clip_fwd_tiles(d.Ibf, d.Tbf, sh, d.ld, d.scale, 0, N, d.ws, 0));
    MR(mrclip_clip_fwd_reduce(sh, d.ws, lse2_row, col_m, col_l, diag2, 0));
    MR(mrclip_lse2_merge(col_m, col_l, 1, N, N, lse2_col, 0));
    MR(mrclip_clip_loss(lse2_row, lse2_col, diag2, N, 0, loss, 0));
    CK(cudaEventRecord(e1));
    MR(mrclip_clip_bwd(d.Ibf, d.Tbf, d.Tt, d.npad, sh, d.ld, lse2_row, lse2_col, d.scale, 1.f, 1.f, 0.5f / N,
                       d.gout, d.ws, dA, MRCLIP_DT_F32, D, dscale, 0, 0));
    CK(cudaEventRecord(e2));
    MR(mrclip_clip_bwd(d.Tbf, d.Ibf, d.It, d.npad, sh, d.ld, lse2_col, lse2_row, d.scale, 1.f, 1.f, 0.5f / N,
                       d.gout, d.ws, dA, MRCLIP_DT_F32, D, dscale, 1, 0));
    CK(cudaEventRecord(e3));
    CK(cudaEventSynchronize(e3));
    float a, b, c;
    cudaEventElapsedTime(&a, e0, e1);
    cudaEventElapsedTime(&b, e1, e2);
    cudaEventElapsedTime(&c, e2, e3);
    if (it >= 2) { tf += a; tb1 += b; tb2 += c; }
  }
  tf /= reps; tb1 /= reps; tb2 /= reps;
  const double flop = 2.0 * N * (double)N * D;
  float hl;
  CK(cudaMemcpy(&hl, loss, 4, cudaMemcpyDeviceToHost));
  printf("  loss %.5f | fwd %.3f ms (%.0f TF/s exec) | bwd passes %.3f + %.3f ms | total %.3f ms -> %.2f Mpairs/s, %.1f TF/s algorithmic (6N^2D)\n",
         hl, tf, flop / tf * 1e-9, tb1, tb2, tf + tb1 + tb2, N / (tf + tb1 + tb2) * 1e-3,
         3 * flop / (tf + tb1 + tb2) * 1e-9);
  {
    void* gmat;
    CK(cudaMalloc(&gmat, mrclip_gmat_bytes(N, N)));
    cudaEvent_t g0, g1, g2, g3;
    cudaEventCreate(&g0); cudaEventCreate(&g1); cudaEventCreate(&g2); cudaEventCreate(&g3);
    float ta = 0, tb = 0, tc = 0;
    for (int it = 0; it < reps + 2; ++it) {
      CK(cudaEventRecord(g0));
      MR(mrclip_clip_gwrite(d.Ibf, d.Tbf, sh, d.ld, lse2_row, lse2_col, d.scale, 1.f, 1.f, 0.5f / N, d.gout, d.ws, gmat,
                            dscale, 0, 1, 0));
      CK(cudaEventRecord(g1));
      MR(mrclip_gmat_gemm(0, gmat, sh, d.Tbf, d.ld, 0.5f / N, d.scale, d.gout, d.ws, dA, MRCLIP_DT_F32, D, 0));
      CK(cudaEventRecord(g2));
      MR(mrclip_gmat_gemm(1, gmat, sh, d.Ibf, d.ld, 0.5f / N, d.scale, d.gout, d.ws, dA, MRCLIP_DT_F32, D, 0));
      CK(cudaEventRecord(g3));
      CK(cudaEventSynchronize(g3));
      float a, b, c;
      cudaEventElapsedTime(&a, g0, g1);
      cudaEventElapsedTime(&b, g1, g2);
      cudaEventElapsedTime(&c, g2, g3);
      if (it >= 2) { ta += a; tb += b; tc += c; }
    }
    ta /= reps; tb /= reps; tc /= reps;
    printf("  gmat: gwrite %.3f ms | gemm %.3f ms (%.0f TF/s) | gemm^T %.3f ms | fwd+gmat total %.3f ms -> %.2f Mpairs/s, %.1f TF/s algorithmic\n",
           ta, tb, flop / tb * 1e-9, tc, tf + ta + tb + tc, N / (tf + ta + tb + tc) * 1e-3,
           3 * flop / (tf + ta + tb + tc) * 1e-9);
    cudaFree(gmat);
  }
  {
    void* emat;
    CK(cudaMalloc(&emat, mrclip_gmat_bytes(N, N)));
    cudaEvent_t g0, g1, g2, g3;
    cudaEventCreate(&g0); cudaEventCreate(&g1); cudaEventCreate(&g2); cudaEventCreate(&g3);
    float ta = 0, tb = 0, tc = 0;
    const int* flag = mrclip_emat_flag(sh, d.ws);
    float* msums;
    CK(cudaMalloc(&msums, 64 * 8));
    const bool wsum_on = getenv("SELFTEST_NO_WSUM") == nullptr;
    for (int it = 0; it < reps + 2; ++it) {
      CK(cudaEventRecord(g0));
      MR(mrclip_clip_fwd_tiles_e(d.Ibf, d.Tbf, sh, d.ld, d.scale, 0, N, d.ws, emat, 0));
      MR(mrclip_clip_fwd_reduce(sh, d.ws, lse2_row, col_m, col_l, diag2, 0));
      MR(mrclip_lse2_merge(col_m, col_l, 1, N, N, lse2_col, 0));
      MR(mrclip_clip_loss(lse2_row, lse2_col, diag2, N, 0, loss, 0));
      CK(cudaEventRecord(g1));
      MR(mrclip_emat_check(sh, d.ws, lse2_row, lse2_col, 0));
      MR(mrclip_clip_gwrite_if(d.Ibf, d.Tbf, sh, d.ld, lse2_row, lse2_col, d.scale, 1.f, 1.f, d.ws, emat, flag, 0));
      MR(mrclip_emat_transform(sh, d.ws, emat, lse2_row, lse2_col, diag2, d.scale, 1.f, 1.f, flag, wsum_on ? msums : nullptr, 64, N, 1, 0));
      CK(cudaEventRecord(g2));
      MR(mrclip_gmat_gemm_dot(0, emat, sh, d.Tbf, d.ld, 0.5f / N, d.scale, d.gout, d.ws, dA, MRCLIP_DT_F32, D, d.Ibf,
                              dscale, 0));
      MR(mrclip_gmat_gemm(1, emat, sh, d.Ibf, d.ld, 0.5f / N, d.scale, d.gout, d.ws, dA, MRCLIP_DT_F32, D, 0));
      CK(cudaEventRecord(g3));
      CK(cudaEventSynchronize(g3));
      float a, b, c;
      cudaEventElapsedTime(&a, g0, g1);
      cudaEventElapsedTime(&b, g1, g2);
      cudaEventElapsedTime(&c, g2, g3);
      if (it >= 2) { ta += a; tb += b; tc += c; }
    }
    ta /= reps; tb /= reps; tc /= reps;
    printf("  emat: fwd+E %.3f ms | check+transform %.3f ms (%.0f) | 2 gemms %.3f ms | total %.3f ms -> %.2f Mpairs/s, %.1f TF/s algorithmic\n",
           ta, tb, flop / tb * 1e-9, tc, ta + tb + tc, N / (ta + tb + tc) * 1e-3, 3 * flop / (ta + tb + tc) * 1e-9);
    cudaFree(emat);
  }
  teardown(d);
}

int main(int argc, char** argv) {
  if (!mrclip_device_ok()) {
    printf("no sm_100 device\n");
    return 1;
  }
  const char* mode = argc > 1 ? argv[1] : "check";
  if (!strcmp(mode, "check")) {
    check_clip(256, 512, 1, 14.285714f, 0.5f);
    check_clip(1024, 768, 1, 14.285714f, 0.5f);
    check_clip(1024, 512, 4, 100.f, 0.1f);
    check_clip(1000, 200, 1, 30.f, 0.2f);
    check_clip(300, 72, 2, 14.285714f, 0.0f);
    check_siglip(1024, 768, 2, 10.f, -10.f);
    check_siglip(520, 264, 1, 10.f, -10.f);
    printf(g_fail ? "SELFTEST FAILED\n" : "SELFTEST PASSED\n");
    return g_fail;
  }
  if (!strcmp(mode, "time")) {
    const int N = argc > 2 ? atoi(argv[2]) : 8192;
    const int D = argc > 3 ? atoi(argv[3]) : 768;
    const int reps = argc > 4 ? atoi(argv[4]) : 5;
    time_shape(N, D, reps);
    return 0;
  }
  printf("usage: selftest check | time N D [reps]\n");
  return 1;
}
