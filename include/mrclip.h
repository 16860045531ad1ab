/*
 * mrclip.h — C ABI of the B200-native contrastive-loss hot path.
 *
 * This is the drop-in boundary for MR-CLIP's distributed contrastive loss
 * (reference: src/open_clip/loss.py).  Plain pointers and sizes only; every
 * pointer is a CUDA device pointer unless stated otherwise; `stream` is a
 * cudaStream_t passed as void*.  Every entry returns 0 on success, a positive
 * cudaError_t on a CUDA failure, a negative value on an argument / driver-API
 * error; mrclip_last_error() returns the message (thread local).  Nothing here
 * synchronises the device, allocates device memory or touches Python.
 *
 * Conventions
 *   - Feature matrices handed to the kernels are bf16, row-major, with leading
 *     dimension `ld = mrclip_padded_dim(d)` (zero padded); mrclip_pack_bf16
 *     produces them from fp32 / bf16 / fp16 inputs.
 *   - "lse2" vectors are log-sum-exp values in log2 units (lse_natural * log2(e)).
 *   - A "row pass" contracts this rank's m_rows rows (A) against all n_cols rows
 *     of the other modality (B).  label_offset is the global column index of
 *     row 0's positive (= rank * m_rows), reference loss.py:94-96.
 */
#ifndef MRCLIP_H_
#define MRCLIP_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRCLIP_DT_F32 0
#define MRCLIP_DT_BF16 1
#define MRCLIP_DT_F16 2

typedef struct mrclip_shape {
  int m_rows;       /* rows owned by this rank (n) */
  int n_cols;       /* rows of the gathered other modality (N = world_size * n) */
  int d;            /* embedding dimension */
  int label_offset; /* global column of local row 0's positive */
} mrclip_shape;

/* library / build identification; mrclip_version() = 10000*major + 100*minor + patch */
int mrclip_version(void);
const char* mrclip_last_error(void);
/* 1 when a CUDA device of compute capability 10.x is present, else 0 (no compute is attempted). */
int mrclip_device_ok(void);

/* leading dimension of packed bf16 rows (d rounded up to 8) and of [*, N] buffers (N rounded to 128) */
int mrclip_padded_dim(int d);
int mrclip_padded_cols(int n_cols);
/* scratch bytes needed by any fwd/bwd entry for this shape (caller allocates, 256 B aligned) */
size_t mrclip_workspace_bytes(int m_rows, int n_cols, int d);
/* column ranges given to mrclip_clip_fwd_tiles must start on a multiple of this many columns */
int mrclip_fwd_col_granule(int m_rows, int n_cols);

/* src [rows, d] of dtype MRCLIP_DT_* with leading dim src_ld  ->  dst bf16 [rows, dst_ld] (zero padded).
 * Replaces the implicit autocast casts in front of loss.py:117-124. */
int mrclip_pack_bf16(const void* src, int src_dtype, int rows, int d, long src_ld, void* dst,
                     int dst_ld, void* stream);
/* src bf16 [rows, src_ld] (cols valid) -> dst bf16 [cols, dst_ld]; feeds the gradient GEMM a K-major operand. */
int mrclip_transpose_bf16(const void* src, int rows, int cols, long src_ld, void* dst, long dst_ld,
                          void* stream);

/* ---- ClipLoss forward (reference loss.py:104-139, get_logits + 2x F.cross_entropy) -------------- */
/* Streams the tiles S = scale * A * B[col_begin:col_end]^T through TMEM and leaves online-LSE
 * partials in `ws`.  May be called several times on disjoint column ranges (e.g. local columns while
 * the all-gather of the rest is in flight); together the calls must cover [0, n_cols). */
int mrclip_clip_fwd_tiles(const void* a_rows, const void* b_all, mrclip_shape shape, int ld,
                          const float* scale, int col_begin, int col_end, void* ws, void* stream);
/* Reduces the partials: lse2_row[m_rows]; per-column (max2,sum) over this rank's rows col_m/col_l[n_cols];
 * diag2[m_rows] = positive logit in log2 units. */
int mrclip_clip_fwd_reduce(mrclip_shape shape, void* ws, float* lse2_row, float* col_m, float* col_l,
                           float* diag2, void* stream);
/* Merges `parts` (max2,sum) partial vectors (element j of part w at [w*part_stride + j]) into
 * lse2_out[mrclip_padded_cols(n_cols)] (+inf in the padding). */
int mrclip_lse2_merge(const float* part_m, const float* part_l, int parts, long part_stride, int n_cols,
                      float* lse2_out, void* stream);
/* loss[0] = (CE_image + CE_text)/2 over this rank's rows (natural log), loss.py:134-137. */
int mrclip_clip_loss(const float* lse2_row, const float* lse2_col, const float* diag2, int m_rows,
                     int label_offset, float* loss, void* stream);

/* ---- ClipLoss backward row pass (autograd of loss.py:117-137 for one operand) -------------------- */
/* d_a[m_rows, d] = coef * scale * grad_out * sum_j G_ij * B_j   with
 *   G_ij = w_own * exp(S_ij - lse_a_i) + w_oth * exp(S_ij - lse_b_j) - (w_own + w_oth) * [j == label_i]
 * d_scale (+)= coef * grad_out * w_own * (sum_ij exp(S_ij - lse_a_i) * C_ij - sum_i C_ii),  C = A * B^T.
 * bt_all is B transposed ([ld, bt_ld] bf16).  lse2_b must be the padded vector from mrclip_lse2_merge.
 * grad_out may be NULL (= 1).  d_scale may be NULL. */
int mrclip_clip_bwd(const void* a_rows, const void* b_all, const void* bt_all, long bt_ld,
                    mrclip_shape shape, int ld, const float* lse2_a, const float* lse2_b,
                    const float* scale, float w_own, float w_oth, float coef, const float* grad_out,
                    void* ws, void* d_a, int out_dtype, long out_ld, float* d_scale,
                    int accumulate_scalars, void* stream);

/* ---- SigLipLoss (reference loss.py:342-448) ------------------------------------------------------ */
/* loss[0] = (1/m_rows) * sum_{i local, j all} softplus(-y_ij (scale * A_i.B_j + bias)), y=+1 iff j==label_i */
int mrclip_siglip_fwd(const void* a_rows, const void* b_all, mrclip_shape shape, int ld,
                      const float* scale, const float* bias, void* ws, float* loss, void* stream);
/* d_a = coef * scale * grad_out * sum_j G_ij B_j,  G = sigmoid(z) - [j==label_i];
 * d_scale (+)= coef*grad_out*sum G*C ; d_bias (+)= coef*grad_out*sum G  (either may be NULL). */
int mrclip_siglip_bwd(const void* a_rows, const void* b_all, const void* bt_all, long bt_ld,
                      mrclip_shape shape, int ld, const float* scale, const float* bias, float coef,
                      const float* grad_out, void* ws, void* d_a, int out_dtype, long out_ld,
                      float* d_scale, float* d_bias, int accumulate_scalars, void* stream);

/* ---- Backward through a materialised bf16 gradient block ("gmat" backend) ------------------------ */
/* The fused row pass above needs no O(n*N) memory but recomputes S once per 384-wide slice of D (TMEM
 * holds 512 fp32 columns).  When mrclip_gmat_bytes(m_rows, n_cols) of scratch is affordable this trio is
 * faster: S is recomputed once, G = dLoss/dS is written as bf16 [m_pad, padded_cols(n_cols)], and each
 * gradient is one plain tcgen05 GEMM against it. */
size_t mrclip_gmat_bytes(int m_rows, int n_cols);
/* G_ij as in mrclip_clip_bwd (unscaled); d_scale as in mrclip_clip_bwd; with both_directions != 0 the
 * other-direction term  coef*grad_out*w_oth*(sum exp(S_ij - lse_b_j) C_ij - sum C_ii)  is added as well
 * (world_size 1, where one G block serves both gradients). */
int mrclip_clip_gwrite(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* lse2_a,
                       const float* lse2_b, const float* scale, float w_own, float w_oth, float coef,
                       const float* grad_out, void* ws, void* gmat, float* d_scale, int accumulate_scalars,
                       int both_directions, void* stream);
int mrclip_siglip_gwrite(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* scale,
                         const float* bias, float coef, const float* grad_out, void* ws, void* gmat, float* d_scale,
                         float* d_bias, int accumulate_scalars, void* stream);
/* transposed == 0: d_out[m_rows, d] = coef*scale*grad_out * G . B      (feat = b_all,  [n_cols, ld] row-major)
 * transposed != 0: d_out[n_cols, d] = coef*scale*grad_out * G^T . A    (feat = a_rows, [m_rows, ld] row-major)
 * The features are consumed as N-major UMMA operands straight from their packed row-major form. */
int mrclip_gmat_gemm(int transposed, const void* gmat, mrclip_shape shape, const void* feat, int ld, float coef,
                     const float* scale, const float* grad_out, void* ws, void* d_out, int out_dtype, long out_ld,
                     void* stream);

/* ---- "emat" backend: nothing N x N is ever recomputed ----------------------------------------------- */
/* The forward of ClipLoss already evaluates E_ij = 2^(S2_ij - c) for its online LSE (c = maximum of the
 * 32 x 64 sub-tile, S2 = S*log2(e)).  mrclip_clip_fwd_tiles_e is mrclip_clip_fwd_tiles that also stores E as bf16
 * into emat [m_pad, mrclip_padded_cols(n_cols)] (mrclip_gmat_bytes) through TMA; the references c stay in ws.
 * The backward then needs no S tiles: mrclip_emat_transform rewrites the block in place (one HBM pass),
 *   G_ij = E_ij * (w_row*2^(c - lse2_row_i) + w_col*2^(c - lse2_col_j)),   positives exactly (fp32, from diag2),
 * and both gradients are plain mrclip_gmat_gemm calls: 3 N x N x D contractions per step instead of 4-5.
 * Replaces loss.py:117-124 + :135-136 and their autograd graph like the other backends. */
int mrclip_clip_fwd_tiles_e(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* scale,
                            int col_begin, int col_end, void* ws, void* emat, void* stream);
/* Guard: bf16 E flushes to zero 126 binary orders below its sub-tile reference.  Raises the device flag
 * mrclip_emat_flag(shape, ws) when some row/column LSE lies more than 80 binary orders below a reference, i.e.
 * when a flushed entry could carry gradient.  mrclip_clip_gwrite_if (exact recompute of G into the same block)
 * runs only while the flag is set (run_if), mrclip_emat_transform only while it is clear (skip_if). */
int mrclip_emat_check(mrclip_shape shape, void* ws, const float* lse2_row, const float* lse2_col, void* stream);
const int* mrclip_emat_flag(mrclip_shape shape, void* ws);
int mrclip_clip_gwrite_if(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* lse2_a,
                          const float* lse2_b, const float* scale, float w_own, float w_oth, void* ws, void* gmat,
                          const int* run_if, void* stream);
/* lse2_row indexes emat rows, lse2_col (padded, +inf) its columns, diag2 its rows.
 * msums (optional, float [msum_slots][2][ranks], zeroed here; the blocks spread their contributions over the slots
 * and the caller adds the slots up): what d(loss)/d(logit_scale) needs, split by the rank that owns the column
 * (n_per_rank columns each):
 *   msums[0][r] = sum_{i, j in rank r} w_row * Prow_ij * log2 Prow_ij,   msums[1][r] = same with w_col * Pcol,
 * log2 P recovered as c + log2(E) - lse2 (positives exact).  With L_q the local loss (natural log) of rank q,
 *   scale * dL_q/dscale = L_q + ln2/(2n) * (sum_r msums_q[0][r] + sum_p msums_p[1][q]);
 * with the guard raised the same sums are taken from the exact recompute's partials (chunk granularity). */
int mrclip_emat_transform(mrclip_shape shape, void* ws, void* emat, const float* lse2_row, const float* lse2_col,
                          const float* diag2, const float* scale, float w_row, float w_col, const int* skip_if,
                          float* msums, int msum_slots, int n_per_rank, int ranks, void* stream);
/* mrclip_gmat_gemm with  dot_out += <d_out, dot_feat> / scale  (dot_feat: bf16 [out_rows, ld]): with d_out = dA and
 * dot_feat = A this is d(loss)/d(scale), by homogeneity of S = scale * A.B^T. */
int mrclip_gmat_gemm_dot(int transposed, const void* gmat, mrclip_shape shape, const void* feat, int ld, float coef,
                         const float* scale, const float* grad_out, void* ws, void* d_out, int out_dtype, long out_ld,
                         const void* dot_feat, float* dot_out, void* stream);
/* ---- fused GEMM -> reduce-scatter over NVLink peer memory ------------------------------------------------
 * The text gradient of rank q is the sum over ranks r of  G_r^T . I_r  restricted to q's rows.  Instead of writing
 * the [n_cols, d] partial locally and calling a reduce-scatter, the transposed GEMM's epilogue stores every output
 * tile straight into its owner's receive buffer:  peer_bufs[q] (device array of `ranks` pointers, each the
 * NVLink-mapped address of rank q's fp32 [ranks][n_per_rank][d] buffer), slot my_rank -- the transfer overlaps the
 * MMA tile by tile.  After a cross-rank barrier (caller) mrclip_sum_slots adds the `ranks` slots on the owner.
 * Replaces torch/distributed/nn/functional.py:343-347 (_AllGather.backward's reduce_scatter) for the text side. */
int mrclip_gmat_gemm_push(const void* gmat, mrclip_shape shape, const void* feat, int ld, float coef,
                          const float* scale, const float* grad_out, void* ws, const unsigned long long* peer_bufs,
                          int n_per_rank, int my_rank, void* stream);
/* All-gather by peer stores: copies `bytes` from src into peer_bufs[k] + dst_offset for every rank k != skip_rank
 * (peer_bufs: device array of NVLink-mapped addresses of the same buffer on every rank).  Replaces the NCCL
 * all-gathers of loss.py:51-57 for the packed features and the LSE statistics; the caller follows it with a
 * cross-rank barrier. */
int mrclip_push_copy(const void* src, size_t bytes, const unsigned long long* peer_bufs, int ranks, size_t dst_offset,
                     int skip_rank, void* stream);
int mrclip_sum_slots(const float* slots, int nslots, int rows, int d, void* out, int out_dtype, long out_ld,
                     void* stream);

/* d logit_scale of a multi-rank local loss (loss.py:117-139 behind mul-backward) without entropy
 * arithmetic in the rescale pass -- selected by MRCLIP_DS=fwd, NOT validated on hardware yet (default off):
 *   s dL_r/ds = <dT_r, T_r> + ln2/(2n) * (R2(r,*) - R2(*,r)),   R2(q, r) = sum_{i in q, j in r} Prow_ij S2_ij.
 * mrclip_clip_fwd_tiles_eu is mrclip_clip_fwd_tiles_e that also keeps u = sum_j 2^(S2_ij - m) S2_ij per row and column
 * chunk half in ws; after mrclip_clip_fwd_reduce, mrclip_row_ent_split turns them into R2(me, q) for every owner q
 * (out_slots: float [64][2][ranks], plane 0, summed over the 64 slots by the caller); mrclip_sum_slots_dot is
 * mrclip_sum_slots that also accumulates <dT_r, T_r> (feat: packed bf16 rows of this rank) into dot_slots[64].
 * mrclip_fwd_row_ent_ok: 1 when every column chunk of the forward plan has a single owner.  The two slot sums of the
 * staged paths (this one and mrclip_sum_slots_bf16) read four columns per thread and need d % 4 == 0. */
int mrclip_fwd_row_ent_ok(int m_rows, int n_cols, int n_per_rank);
int mrclip_clip_fwd_tiles_eu(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* scale,
                             int col_begin, int col_end, void* ws, void* emat, void* stream);
int mrclip_row_ent_split(mrclip_shape shape, void* ws, const float* lse2_row, int n_per_rank, int ranks,
                         float* out_slots, void* stream);
int mrclip_sum_slots_dot(const float* slots, int nslots, int rows, int d, void* out, int out_dtype, long out_ld,
                         const void* feat, long feat_ld, float* dot_slots, void* stream);

/* Fused GEMM -> reduce-scatter with a bf16 payload (MRCLIP_PUSH_DTYPE=bf16, NOT validated on hardware yet, default off):
 * mrclip_gmat_gemm_push whose peer_bufs are bf16 [W, n_per_rank, d] receive buffers (two accumulator chunks per
 * 128-byte row piece: half the NVLink bytes of loss.py's all_gather backward), and the owner-side slot sum over bf16
 * slots, optionally with the <dT_r, T_r> dot of mrclip_sum_slots_dot (feat and dot_slots both NULL or both set). */
int mrclip_gmat_gemm_push_bf16(const void* gmat, mrclip_shape shape, const void* feat, int ld, float coef,
                               const float* scale, const float* grad_out, void* ws, const unsigned long long* peer_bufs,
                               int n_per_rank, int my_rank, void* stream);
int mrclip_sum_slots_bf16(const void* slots, int nslots, int rows, int d, void* out, int out_dtype, long out_ld,
                          const void* feat, long feat_ld, float* dot_slots, void* stream);

/* SigLipLoss has no normaliser: its forward can store G = sigmoid(z) - [j==label_i] directly (then both
 * gradients are plain mrclip_gmat_gemm calls) and leaves the d_scale / d_bias partial sums in ws. */
int mrclip_siglip_fwd_e(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* scale,
                        const float* bias, void* ws, float* loss, void* gmat, void* stream);
int mrclip_siglip_e_scalars(mrclip_shape shape, void* ws, float coef, const float* grad_out, float* d_scale,
                            float* d_bias, int accumulate_scalars, void* stream);

/* ---- whole-step entries ------------------------------------------------------------------------------------------
 * One call launches every kernel of a loss forward (or backward) on `stream`: what ClipLoss.forward / SigLipLoss.forward
 * (loss.py:128-139, :365-448) and their autograd graph do, behind one C call per direction.  On several ranks (one
 * process per GPU) nothing but this library's kernels moves data: the packed text rows, the LSE statistics, the
 * text-gradient tiles and the scalars travel over NVLink peer memory and are synchronised by device-side flags
 * (csrc/peer_sync.cuh) -- no NCCL call, no host synchronisation, so the step can also be captured into a CUDA graph.
 * The text all-gather (loss.py:51-57) overlaps the forward tiles: a rank starts on its own columns and picks up the
 * other ranks' columns as their rows land; the gradient reduce-scatter (torch/distributed/nn/functional.py:343-347) is
 * fused into the dT GEMM's epilogue.
 *
 * mrclip_peer: plumbing of one workspace.  Every buffer named "_peers" is a device array [ranks] of NVLink-mapped
 * addresses of the SAME symmetric buffer on every rank (own rank included).  ctl_block: symmetric,
 * mrclip_peer_block_bytes() bytes, zeroed once; ctl: plain device memory, 128 int32, zeroed once.  ranks <= 1: all
 * pointers may be NULL. */
typedef struct mrclip_peer {
  int ranks, rank;
  const unsigned long long* ctl_block_peers;
  void* ctl_block;
  int* ctl;
  const unsigned long long* txt_peers;    /* bf16 [N, ld] gathered text buffer (the one used by this step) */
  const unsigned long long* stats_peers;  /* float [ranks][3][N] */
  const unsigned long long* recv_peers;   /* [ranks][n][d] receive slots of the fused reduce-scatter (fp32 or bf16) */
  void* recv;                             /* this rank's receive slots */
  int recv_bf16;
  /* HOST copies of ctl_block_peers / txt_peers and a host counter (one int, zero-initialised, owned by the caller): the
   * text all-gather runs on the copy engines (cudaMemcpyAsync to the peers' buffers on a side stream, each followed by
   * a stream memory operation that raises the destination's flag), so the forward kernel keeps every SM.  NULL: the
   * pack kernel stores the rows to the peers itself before the forward starts (no overlap). */
  const unsigned long long* ctl_block_peers_host;
  const unsigned long long* txt_peers_host;
  int* host_epoch;
} mrclip_peer;

typedef struct mrclip_step {
  mrclip_shape shape;      /* m_rows = n (rows of this rank), n_cols = N = ranks * n, label_offset = rank * n */
  int ld;                  /* mrclip_padded_dim(d) */
  int kind;                /* 0 = ClipLoss, 1 = SigLipLoss */
  int local_loss;          /* ClipLoss on several ranks: 0 = loss and d logit_scale are averaged over the ranks */
  void* img_rows;          /* bf16 [n, ld]   packed image rows of this rank (written by the forward) */
  void* txt_all;           /* bf16 [N, ld]   this rank's gathered text buffer (symmetric when ranks > 1) */
  void* ws;                /* mrclip_workspace_bytes(n, N, d) */
  void* emat;              /* mrclip_gmat_bytes(n, N): E / G block (NULL: forward only, nothing kept for a backward) */
  float* stats;            /* float [ranks][3][N]  (symmetric when ranks > 1) */
  float* lse2_row_all;     /* float [mrclip_padded_cols(N)] */
  float* lse2_col_all;     /* float [mrclip_padded_cols(N)] */
  float* msums;            /* float [64][2][ranks] */
  float* small;            /* float [mrclip_step_small_floats()], zeroed once */
  float* inv_norm;         /* float [2][n] (image rows, text rows) for mrclip_step_forward(raw != 0); else may be NULL */
  float* scale_buf;        /* float [1]: exp(logit_scale) written by a raw forward; else may be NULL */
  mrclip_peer peer;
} mrclip_step;

size_t mrclip_peer_block_bytes(void);
size_t mrclip_step_small_floats(void);
/* sizeof(mrclip_step) / sizeof(mrclip_peer) as this library was compiled: lets a foreign-language binding check its
 * struct declarations against the ABI at load time (tests/test_cabi.py does so for the ctypes ones). */
size_t mrclip_step_struct_bytes(void);
size_t mrclip_peer_struct_bytes(void);
/* 1 when mrclip_step_backward will take d logit_scale from forward-side row sums (the forward must then be told:
 * fwd_ds), 0 when it uses the entropy sums of the rescale pass.  Pure function of the shape and mode. */
int mrclip_step_uses_fwd_ds(const mrclip_step* s);
/* img / txt: this rank's [n, d] feature rows (MRCLIP_DT_*, leading dims in elements).  scale (and bias, SigLIP, may be
 * NULL) are device scalars.  need_grad == 0: no E / G block is written.  loss_out: device float [1].
 * raw != 0 (feature hand-off fusion, reference model.py:282-301, :324): img / txt are the towers' UN-normalised outputs
 * and `scale` is the model's log-scale parameter; the pack pre-pass normalises the rows (F.normalize, eps 1e-12, fp32),
 * keeps 1/||x|| in s->inv_norm and exp(scale) in s->scale_buf (which every later kernel of the step, and the backward,
 * uses as the logit scale). */
int mrclip_step_forward(const mrclip_step* s, const void* img, int img_dtype, long img_ld, const void* txt, int txt_dtype,
                        long txt_ld, const float* scale, const float* bias, int need_grad, int raw, float* loss_out,
                        void* stream);
/* Backward of the row normalisation, in place on a gradient block g [rows, d] (MRCLIP_DT_*, leading dim g_ld):
 * g <- (g - y <g, y>) * inv_norm[row], y = the packed bf16 rows [rows, y_ld].  Chains d(loss)/d(normalised features),
 * as mrclip_step_backward leaves it, to the tower outputs (autograd of F.normalize, model.py:282-301). */
int mrclip_normalize_bwd(const void* y, long y_ld, const float* inv_norm, int rows, int d, void* g, int g_dtype, long g_ld,
                         void* stream);
/* coef: 1/(2n) (or 1/(2N) for ClipLoss(local_loss=False, gather_with_grad=False)); 1/n for SigLIP.  grad_out: device
 * float [1] or NULL.  d_img / d_txt: [n, d] of MRCLIP_DT_*; d_scale / d_bias: device float [1] or NULL. */
int mrclip_step_backward(const mrclip_step* s, const float* scale, const float* grad_out, float coef, void* d_img,
                         int d_img_dtype, long d_img_ld, void* d_txt, int d_txt_dtype, long d_txt_ld, float* d_scale,
                         float* d_bias, void* stream);

/* ---- MultiPositiveClipLoss (reference loss.py:626-644, :671-747) on top of the ClipLoss pipeline -------------------
 * With P(i) the samples of row i's class: loss_i = lse_i - mean_{j in P(i)} S_ij and the mean is s <I_i, mean_{P(i)} T>,
 * so the loss needs only the row / column LSEs (the ClipLoss forward) and class means of the packed features.
 * mrclip_class_means: mean[c] (fp32 [n_ids, ld]) = average of the rows of x (bf16 [*, ld]) of class id c; order =
 *   sample indices grouped by class, seg_start[c] / seg_cnt[c] the group of id c (seg_cnt 0: unused id).
 * mrclip_mpos_forward: loss[0] = (1/n) sum_i delta (lse_row_i - s <I_i, tmean[cls_i]>) + (1-delta) (lse_col_i - s <T_i, imean[cls_i]>)
 *   over this rank's n rows (img / txt: its packed rows; lse2_*: log2 units, lse2_col indexed by local row).
 * mrclip_mpos_backward: d_img += k (T_i - tmean[cls_i]), d_txt += k (I_i - imean[cls_i]), k = coef * scale * grad_out --
 *   the part of the gradient that the E-block GEMMs (weights delta / 1-delta) do not carry. */
int mrclip_class_means(const void* x, int ld, const int* order, const int* seg_start, const int* seg_cnt, int n_ids,
                       float* mean, void* stream);
int mrclip_mpos_forward(const void* img_rows, const void* txt_rows, int ld, int n, int d, const int* cls, const float* tmean,
                        const float* imean, const float* lse2_row, const float* lse2_col, const float* scale, float delta,
                        float* loss, void* stream);
int mrclip_mpos_backward(void* d_img, int d_img_dtype, long d_img_ld, void* d_txt, int d_txt_dtype, long d_txt_ld,
                         const void* img_rows, const void* txt_rows, int ld, int n, int d, const int* cls, const float* tmean,
                         const float* imean, float coef, const float* scale, const float* grad_out, void* stream);

/* ---- retrieval metrics (reference open_clip_train/train.py:465-534, get_clip_metrics) ---------------------------------
 * The reference forms the full logit matrix on the CPU, argsorts every row and walks Python loops to find where the
 * samples of the row's class ("positives") rank.  Here the same S = A . B^T tiles carry a rank-of-label epilogue:
 *   mrclip_rank_collect: pos[row_off[i] + col_ord[j]] = <A_i, B_j> for every column j of row i's class
 *   mrclip_rank_lmax:    lmax[i] = largest positive of row i
 *   mrclip_rank_count:   best[i] += #{k negative: C_ik > lmax[i]}   (chunk0 == 0 only; the best positive's 0-based rank)
 *                        pairs[i] += #{(k negative, t in [chunk0, chunk0 + 32) positive): C_ik > pos_i[t]}
 * so that  min rank = best,  mean rank of the positives = (pairs + m (m - 1) / 2) / m  (m = class size), with strict
 * comparisons (a negative that ties a positive ranks after it).  a_rows / b_all: packed bf16 [*, ld]; row_cls [m_rows],
 * row_off [m_rows] (int64), row_m [m_rows]; col_cls / col_ord [mrclip_padded_cols(n_cols)] with class -1 in the padding;
 * pairs: uint64 [m_rows], best: int32 [m_rows], both zeroed by the caller.  shape.label_offset is ignored. */
int mrclip_rank_collect(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const int* row_cls,
                        const int* col_cls, const int* col_ord, const long long* row_off, float* pos, void* stream);
int mrclip_rank_lmax(const float* pos, const long long* row_off, const int* row_m, int rows, float* lmax, void* stream);
int mrclip_rank_count(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const int* row_cls,
                      const int* col_cls, const long long* row_off, const int* row_m, const float* pos, const float* lmax,
                      int chunk0, unsigned long long* pairs, int* best, void* stream);

/* Optional timing of the launch groups inside the whole-step entries (bench.py's roofline): while enabled, CUDA events
 * are recorded on the launching stream around each group; mrclip_prof_report synchronises them and writes
 * "name:total_ms:count;..." into buf.  mrclip_prof_enable(0/1) also discards what was recorded.  Off by default. */
int mrclip_prof_enable(int on);
int mrclip_prof_report(char* buf, size_t cap);

/* number of kernels this library has launched on behalf of the calling process (for bench accounting) */
long mrclip_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MRCLIP_H_ */
