// C-ABI implementation: host-side planning (work decomposition, TMA tensor maps, workspace
// carving) and kernel launches.  See include/mrclip.h for the contract.
#include "../../include/mrclip.h"
#include "aux_kernels.cuh"
#include "gemm2_kernel.cuh"
#include "gemm_kernel.cuh"
#include "peer_kernels.cuh"
#include "tile_kernel.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

using namespace mrclip;

thread_local std::string g_err;
std::atomic<long> g_launches{0};

// Optional per-launch-group timing for bench.py (mrclip_prof_enable): CUDA events on the launching stream around each
// group of the whole-step entries, read back by mrclip_prof_report.  Off by default: no events, no overhead.
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::atomic<bool> g_prof_on{false};
struct ProfScope {
  cudaStream_t st;
  cudaEvent_t b = nullptr;
  ProfScope(const char* name, cudaStream_t s) : st(s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec r;
    r.name = name;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    b = r.b;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
  }
  ~ProfScope() {
    if (b) cudaEventRecord(b, st);
  }
};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail((int)e_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                               \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int num_sms() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}

// ------------------------------------------------------------------ TMA tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// bf16 matrix [rows, cols] with row stride ld elements; box = [box_rows, 64 cols], 128B swizzle
int make_map(CUtensorMap* map, const void* base, long rows, long cols, long ld, int box_rows) {
  // (also used for the E store map: box [32 rows x 64 cols])
  // box is always 64 elements (128 bytes) wide: one swizzle row
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(-2, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(-3, "TMA operand not 16-byte aligned (base %p, ld %ld)", base, ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                  estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-4, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// ------------------------------------------------------------------ plans
constexpr int kSBN = 256;   // S-tile width of the S-only kernels (FWD, GW)

struct FwdPlan {
  int num_rb, num_ct, tiles_per_chunk, total_chunks, m_pad, n_pad, bands;
};
FwdPlan fwd_plan(int m_rows, int n_cols) {
  FwdPlan f;
  f.num_rb = ceil_div(m_rows, kBM);
  f.num_ct = ceil_div(n_cols, kSBN);
  const long total = (long)f.num_rb * f.num_ct;
  long tpc = total / (148L * 8);
  if (tpc < 1) tpc = 1;
  if (tpc > 32) tpc = 32;
  while (tpc & (tpc - 1)) tpc &= tpc - 1;  // power of two, so per-rank column ranges stay chunk aligned
  f.tiles_per_chunk = (int)tpc;
  f.total_chunks = ceil_div(f.num_ct, f.tiles_per_chunk);
  f.m_pad = f.num_rb * kBM;
  f.n_pad = f.num_ct * kSBN;
  f.bands = f.m_pad / 32;
  return f;
}

struct BwdPlan {
  int num_rb, num_ct, dc, num_dc, d_pad, cs, tiles_per_chunk, m_pad, n_pad, num_items;
};
BwdPlan bwd_plan(int m_rows, int n_cols, int ld) {
  BwdPlan b;
  b.num_rb = ceil_div(m_rows, kBM);
  b.num_ct = ceil_div(n_cols, kBN);
  b.num_dc = ceil_div(ld, 384);
  const int dcw = ceil_div(ld, b.num_dc);
  b.dc = dcw <= 256 ? 256 : 384;
  b.d_pad = b.num_dc * b.dc;
  b.m_pad = b.num_rb * kBM;
  b.n_pad = b.num_ct * kBN;
  const int sms = num_sms();
  const int base = b.num_rb * b.num_dc;
  int best_cs = 1;
  double best_eff = 0.0;
  for (int cs = 1; cs <= 32 && cs <= b.num_ct; ++cs) {
    const int tpc = ceil_div(b.num_ct, cs);
    if (cs > 1 && tpc < 4) break;
    const int cs_eff = ceil_div(b.num_ct, tpc);
    const long items = (long)base * cs_eff;
    const double eff = (double)items / ((double)ceil_div((int)items, sms) * sms);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best_cs = cs_eff;
    }
    if (best_eff >= 0.97) break;
  }
  b.tiles_per_chunk = ceil_div(b.num_ct, best_cs);
  b.cs = ceil_div(b.num_ct, b.tiles_per_chunk);
  b.num_items = base * b.cs;
  return b;
}

struct GemmPlan {
  int num_rb, num_dt, num_kb, ksplit, kb_per_split, num_items, m_pad, d_pad;
};
// gradient GEMMs run on CTA pairs (gemm2_kernel, 256 x 256 tiles) unless MRCLIP_GEMM_CTA=1
bool gemm_pairs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MRCLIP_GEMM_CTA");
    v = (e && atoi(e) == 1) ? 0 : 1;
  }
  return v == 1;
}
// out_rows = rows of the gradient being produced, k_len = length of the contraction
GemmPlan gemm_plan(int out_rows, int k_len, int ld, bool force_single_split = false) {
  GemmPlan g;
  const int rows_per_item = gemm_pairs() ? 2 * kBM : kBM;
  g.num_rb = ceil_div(out_rows, rows_per_item);
  g.num_dt = ceil_div(ld, kGemmBN);
  g.num_kb = ceil_div(k_len, kBK);
  g.m_pad = g.num_rb * rows_per_item;
  g.d_pad = g.num_dt * kGemmBN;
  const int sms = gemm_pairs() ? num_sms() / 2 : num_sms();   // schedulable units: CTA pairs or CTAs
  const int base = g.num_rb * g.num_dt;
  int best = 1;
  double best_eff = 0.0;
  // short contractions with many output tiles (the [N, d] text-gradient partial of one rank, K = n): the fp32
  // split-K partials would cost more HBM traffic than the wave quantisation they repair -> one split, and the
  // GEMM's epilogue scales and stores the result itself
  int max_ks = (g.num_kb <= 128 && base >= 3 * sms) ? 1 : 16;
  if (force_single_split) max_ks = 1;
  if (const char* e = getenv("MRCLIP_GEMM_MAX_KSPLIT")) {   // test / experiment knob
    const int v = atoi(e);
    if (v >= 1 && v <= 16) max_ks = v;
  }
  for (int ks = 1; ks <= max_ks; ++ks) {
    const int per = ceil_div(g.num_kb, ks);
    if (ks > 1 && per < 16) break;
    const int ks_eff = ceil_div(g.num_kb, per);
    const long items = (long)base * ks_eff;
    const double eff = (double)items / ((double)ceil_div((int)items, sms) * sms);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = ks_eff;
    }
    if (best_eff >= 0.95) break;
  }
  g.kb_per_split = ceil_div(g.num_kb, best);
  g.ksplit = ceil_div(g.num_kb, g.kb_per_split);
  g.num_items = base * g.ksplit;
  return g;
}

// stream-K scratch at the start of the (aliased) backward view: partial tiles of up to 74 CTA pairs, then the tile flags
constexpr int kSkMaxPairs = 74, kSkMaxTiles = 16384;
constexpr size_t kSkPartBytes = (size_t)kSkMaxPairs * 2 * 128 * 256 * sizeof(float);
constexpr size_t kSkBytes = kSkPartBytes + (size_t)kSkMaxTiles * sizeof(int);
struct WsLayout {
  size_t row_part, col_l, col_c, diag2, sc_part, sc_part2, cbmin, flag, dpart, row_ent, total;
};
WsLayout ws_layout(int m_rows, int n_cols, int d) {
  const int ld = mrclip_padded_dim(d);
  const FwdPlan f = fwd_plan(m_rows, n_cols);
  const BwdPlan b = bwd_plan(m_rows, n_cols, ld);
  WsLayout w;
  size_t off = 0;
  // forward view
  w.row_part = off;
  off += align_up((size_t)f.total_chunks * 2 * f.m_pad * sizeof(float2), 256);
  w.col_l = off;
  off += align_up((size_t)f.bands * f.n_pad * sizeof(float), 256);
  const size_t fwd_end = off;
  // backward view (aliases the forward view; forward partials are dead by then)
  w.dpart = 0;
  size_t bwd_end = align_up((size_t)b.cs * b.m_pad * b.d_pad * sizeof(float), 256);
  {
    const GemmPlan g1 = gemm_plan(m_rows, n_cols, ld);   // dA = G . Bt      (contraction over columns)
    const GemmPlan g2 = gemm_plan(n_cols, m_rows, ld);   // dB = G^T . At    (contraction over rows)
    const size_t e1 = align_up((size_t)g1.ksplit * g1.m_pad * g1.d_pad * sizeof(float), 256);
    const size_t e2 = align_up((size_t)g2.ksplit * g2.m_pad * g2.d_pad * sizeof(float), 256);
    if (e1 > bwd_end) bwd_end = e1;
    if (e2 > bwd_end) bwd_end = e2;
    if (kSkBytes > bwd_end) bwd_end = kSkBytes;      // stream-K: one partial tile per CTA pair + the tile flags
  }
  off = fwd_end > bwd_end ? fwd_end : bwd_end;
  w.sc_part = off;
  const size_t items_f = (size_t)f.num_rb * f.total_chunks;
  const size_t items = items_f > (size_t)b.num_items ? items_f : (size_t)b.num_items;
  off += align_up(items * kEpiWarps * sizeof(float2), 256);
  // regions that must survive from the forward into the backward (never aliased by dpart):
  // the sub-tile references of E, SigLIP's d_scale/d_bias partials, the E-block guard
  w.col_c = off;
  off += align_up((size_t)f.bands * (f.n_pad / 64) * sizeof(float), 256);
  w.sc_part2 = off;
  off += align_up(items_f * kEpiWarps * sizeof(float2), 256);
  w.cbmin = off;
  off += align_up((size_t)(f.n_pad / 64) * sizeof(float), 256);
  w.flag = off;
  off += 256;
  // MODE_FWDEU: u of every (row, slot); appended so that every other offset stays where it was
  w.row_ent = off;
  off += align_up((size_t)f.total_chunks * 2 * f.m_pad * sizeof(float), 256);
  // positive logits of this rank's rows: read by the rescale pass, which may run (banded, on the side stream) while
  // a gradient GEMM already writes its split-K partials over the forward view -> kept outside the aliased region
  w.diag2 = off;
  off += align_up((size_t)f.m_pad * sizeof(float), 256);
  w.total = off;
  return w;
}

template <int MODE, int LOSS, int DC, int BN>
int launch_tile(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mbt, const TileParams& p,
                cudaStream_t st) {
  using Cfg = TileCfg<MODE, LOSS, DC, BN>;
  auto kern = tile_kernel<MODE, LOSS, DC, BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  int grid = p.num_items < num_sms() ? p.num_items : num_sms();
  if (const char* g = getenv("MRCLIP_GRID")) {  // experiment knob: cap the number of persistent CTAs
    const int cap = atoi(g);
    if (cap > 0 && cap < grid) grid = cap;
  }
  if (grid <= 0) return 0;
  kern<<<grid, kThreads, Cfg::kSmemBytes, st>>>(ma, mb, mbt, p);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int check_shape(const mrclip_shape& s, int ld) {
  if (s.m_rows <= 0 || s.n_cols <= 0 || s.d <= 0) return fail(-1, "empty shape (m=%d n=%d d=%d)", s.m_rows, s.n_cols, s.d);
  if (ld < s.d || ld % 8 != 0) return fail(-1, "ld=%d must be a multiple of 8 and >= d=%d", ld, s.d);
  if (s.label_offset < 0) return fail(-1, "negative label_offset");
  return 0;
}

struct FwdSig {   // TileParams::sig_* (multi-rank forward whose B rows arrive over NVLink while it runs)
  const int* ready = nullptr;
  const int* epoch = nullptr;
  int src_cols = 0, my_src = 0;
};
int run_fwd(int loss_kind, const void* a_rows, const void* b_all, const mrclip_shape& sh, int ld,
            const float* scale, const float* bias, int col_begin, int col_end, void* ws, void* emat,
            cudaStream_t st, bool row_ent = false, const FwdSig* sig = nullptr) {
  if (int e = check_shape(sh, ld)) return e;
  const FwdPlan f = fwd_plan(sh.m_rows, sh.n_cols);
  const WsLayout w = ws_layout(sh.m_rows, sh.n_cols, sh.d);
  const int granule = f.tiles_per_chunk * kSBN;
  if (col_begin < 0 || col_end > sh.n_cols || col_begin >= col_end) return fail(-1, "bad column range [%d,%d)", col_begin, col_end);
  if (col_begin % granule != 0) return fail(-1, "col_begin=%d not a multiple of the granule %d", col_begin, granule);
  if (col_end != sh.n_cols && col_end % granule != 0) return fail(-1, "col_end=%d not a multiple of the granule %d", col_end, granule);
  CUtensorMap ma, mb, me;
  if (int e = make_map(&ma, a_rows, sh.m_rows, ld, ld, kBM)) return e;
  if (int e = make_map(&mb, b_all, sh.n_cols, ld, ld, kSBN)) return e;
  me = mb;
  if (emat)   // E / G block [m_pad, n_pad] bf16, stored in 32 x 64 boxes
    if (int e = make_map(&me, emat, f.m_pad, f.n_pad, f.n_pad, 32)) return e;
  TileParams p;
  memset(&p, 0, sizeof p);
  p.m_rows = sh.m_rows;
  p.n_cols = sh.n_cols;
  p.num_kb = ceil_div(ld, kBK);
  p.num_rb = f.num_rb;
  p.tile_begin = col_begin / kSBN;
  p.tile_end = ceil_div(col_end, kSBN);
  p.tiles_per_chunk = f.tiles_per_chunk;
  p.num_chunks = ceil_div(p.tile_end - p.tile_begin, f.tiles_per_chunk);
  p.chunk_base = p.tile_begin / f.tiles_per_chunk;
  p.num_dc = 1;
  p.num_items = p.num_chunks * f.num_rb;
  p.label_offset = sh.label_offset;
  p.m_pad = f.m_pad;
  p.n_pad = f.n_pad;
  p.d_pad = 0;
  p.scale = scale;
  p.bias = bias;
  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  p.row_part = reinterpret_cast<float2*>(wsb + w.row_part);
  p.col_l = reinterpret_cast<float*>(wsb + w.col_l);
  p.col_c = reinterpret_cast<float*>(wsb + w.col_c);
  p.diag2 = reinterpret_cast<float*>(wsb + w.diag2);
  p.sc_part = reinterpret_cast<float2*>(wsb + w.sc_part) + (size_t)p.chunk_base * f.num_rb * kEpiWarps;
  p.sc_part2 = reinterpret_cast<float2*>(wsb + w.sc_part2) + (size_t)p.chunk_base * f.num_rb * kEpiWarps;
  p.row_ent = reinterpret_cast<float*>(wsb + w.row_ent);
  if (sig != nullptr && sig->ready != nullptr && sig->src_cols > 0) {
    p.sig_ready = sig->ready;
    p.sig_epoch = sig->epoch;
    p.src_cols = sig->src_cols;
    p.my_src = sig->my_src;
    // own columns first: rotate by the chunk that holds this rank's first column (whole range only)
    if (col_begin == 0 && col_end == sh.n_cols)
      p.chunk_rot = (int)(((long)sig->my_src * sig->src_cols) / (f.tiles_per_chunk * kSBN)) % p.num_chunks;
  }
  if (emat && row_ent && loss_kind == LOSS_CLIP) return launch_tile<MODE_FWDEU, LOSS_CLIP, 256, kSBN>(ma, mb, me, p, st);
  if (emat) {
    if (loss_kind == LOSS_CLIP) return launch_tile<MODE_FWDE, LOSS_CLIP, 256, kSBN>(ma, mb, me, p, st);
    return launch_tile<MODE_FWDE, LOSS_SIGLIP, 256, kSBN>(ma, mb, me, p, st);
  }
  if (loss_kind == LOSS_CLIP) return launch_tile<MODE_FWD, LOSS_CLIP, 256, kSBN>(ma, mb, mb, p, st);
  return launch_tile<MODE_FWD, LOSS_SIGLIP, 256, kSBN>(ma, mb, mb, p, st);
}

int run_bwd(int loss_kind, const void* a_rows, const void* b_all, const void* bt_all, long bt_ld,
            const mrclip_shape& sh, int ld, const float* lse2_a, const float* lse2_b, const float* scale,
            const float* bias, float w_own, float w_oth, float coef, const float* grad_out, void* ws,
            void* d_a, int out_dtype, long out_ld, float* d_scale, float* d_bias, int accumulate,
            cudaStream_t st) {
  if (int e = check_shape(sh, ld)) return e;
  if (out_dtype < 0 || out_dtype > 2) return fail(-1, "bad out_dtype %d", out_dtype);
  const BwdPlan b = bwd_plan(sh.m_rows, sh.n_cols, ld);
  const WsLayout w = ws_layout(sh.m_rows, sh.n_cols, sh.d);
  if (bt_ld < sh.n_cols || bt_ld % 8 != 0) return fail(-1, "bt_ld=%ld must be a multiple of 8 and >= n_cols", bt_ld);
  CUtensorMap ma, mb, mbt;
  if (int e = make_map(&ma, a_rows, sh.m_rows, ld, ld, kBM)) return e;
  if (int e = make_map(&mb, b_all, sh.n_cols, ld, ld, kBN)) return e;
  if (int e = make_map(&mbt, bt_all, ld, sh.n_cols, bt_ld, b.dc == 384 ? 192 : 256)) return e;
  TileParams p;
  memset(&p, 0, sizeof p);
  p.m_rows = sh.m_rows;
  p.n_cols = sh.n_cols;
  p.num_kb = ceil_div(ld, kBK);
  p.num_rb = b.num_rb;
  p.tile_begin = 0;
  p.tile_end = b.num_ct;
  p.tiles_per_chunk = b.tiles_per_chunk;
  p.num_chunks = b.cs;
  p.chunk_base = 0;
  p.num_dc = b.num_dc;
  p.num_items = b.num_items;
  p.label_offset = sh.label_offset;
  p.m_pad = b.m_pad;
  p.n_pad = b.n_pad;
  p.d_pad = b.d_pad;
  p.scale = scale;
  p.bias = bias;
  p.w_own = w_own;
  p.w_oth = w_oth;
  p.lse2_a = lse2_a;
  p.lse2_b = lse2_b;
  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  p.dpart = reinterpret_cast<float*>(wsb + w.dpart);
  p.sc_part = reinterpret_cast<float2*>(wsb + w.sc_part);
  int e = 0;
  if (loss_kind == LOSS_CLIP)
    e = b.dc == 384 ? launch_tile<MODE_BWD, LOSS_CLIP, 384, 128>(ma, mb, mbt, p, st)
                    : launch_tile<MODE_BWD, LOSS_CLIP, 256, 128>(ma, mb, mbt, p, st);
  else
    e = b.dc == 384 ? launch_tile<MODE_BWD, LOSS_SIGLIP, 384, 128>(ma, mb, mbt, p, st)
                    : launch_tile<MODE_BWD, LOSS_SIGLIP, 256, 128>(ma, mb, mbt, p, st);
  if (e) return e;
  {
    const long total = (long)sh.m_rows * (b.d_pad / 4);
    const int threads = 256;
    long blocks = (total + threads - 1) / threads;
    if (blocks > 148L * 16) blocks = 148L * 16;
    grad_reduce_kernel<<<(int)blocks, threads, 0, st>>>(p.dpart, b.cs, sh.m_rows, sh.d, b.m_pad, b.d_pad, coef,
                                                        scale, grad_out, d_a, out_dtype, out_ld, nullptr, 0, nullptr);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
  }
  if (d_scale || d_bias) {
    const float cx = coef * (loss_kind == LOSS_CLIP ? w_own : 1.f);
    scalar_reduce_kernel<<<1, 1024, 0, st>>>(p.sc_part, (long)b.num_items * kEpiWarps, cx, coef, grad_out,
                                             nullptr, d_scale, d_bias, accumulate, 0);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

template <bool A_MN, int PUSH, bool SK = false>
int launch_gemm2(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t st) {
  auto kern = gemm2_kernel<A_MN, PUSH, SK>;
  constexpr int kG2SmemBytes = G2Cfg<PUSH>::kSmemBytes;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kG2SmemBytes));
    attr_set = true;
  }
  int pairs_avail = num_sms() / 2;
  if (SK && pairs_avail > kSkMaxPairs) pairs_avail = kSkMaxPairs;
  const int pairs = p.num_items < pairs_avail ? p.num_items : pairs_avail;
  if (pairs <= 0) return 0;
  kern<<<2 * pairs, kThreads, kG2SmemBytes, st>>>(ma, mb, p);   // __cluster_dims__(2,1,1)
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

template <bool A_MN>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t st) {
  if (p.peer && p.out_dtype == DT_BF16) {   // bf16 push payload: CTA-pair kernel only
    if (!gemm_pairs()) return fail(-1, "the bf16 push epilogue needs the CTA-pair GEMM (MRCLIP_GEMM_CTA=2)");
    return launch_gemm2<A_MN, 2>(ma, mb, p, st);
  }
  if (gemm_pairs()) return p.peer ? launch_gemm2<A_MN, 1>(ma, mb, p, st) : launch_gemm2<A_MN, 0>(ma, mb, p, st);
  auto kern = gemm_kernel<A_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    attr_set = true;
  }
  int grid = p.num_items < num_sms() ? p.num_items : num_sms();
  if (const char* g = getenv("MRCLIP_GRID")) {
    const int cap = atoi(g);
    if (cap > 0 && cap < grid) grid = cap;
  }
  if (grid <= 0) return 0;
  kern<<<grid, kThreads, kGemmSmemBytes, st>>>(ma, mb, p);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// Recompute S for this rank's row block and write G = dLoss/dS (bf16, unscaled) to gmat [m_pad, n_pad].
int run_gwrite(int loss_kind, const void* a_rows, const void* b_all, const mrclip_shape& sh, int ld,
               const float* lse2_a, const float* lse2_b, const float* scale, const float* bias, float w_own,
               float w_oth, void* ws, void* gmat, const int* run_if, cudaStream_t st) {
  if (int e = check_shape(sh, ld)) return e;
  const FwdPlan f = fwd_plan(sh.m_rows, sh.n_cols);
  const WsLayout w = ws_layout(sh.m_rows, sh.n_cols, sh.d);
  CUtensorMap ma, mb;
  if (int e = make_map(&ma, a_rows, sh.m_rows, ld, ld, kBM)) return e;
  if (int e = make_map(&mb, b_all, sh.n_cols, ld, ld, kSBN)) return e;
  TileParams p;
  memset(&p, 0, sizeof p);
  p.m_rows = sh.m_rows;
  p.n_cols = sh.n_cols;
  p.num_kb = ceil_div(ld, kBK);
  p.num_rb = f.num_rb;
  p.tile_begin = 0;
  p.tile_end = f.num_ct;
  p.tiles_per_chunk = f.tiles_per_chunk;
  p.num_chunks = f.total_chunks;
  p.chunk_base = 0;
  p.num_dc = 1;
  p.num_items = f.total_chunks * f.num_rb;
  p.label_offset = sh.label_offset;
  p.m_pad = f.m_pad;
  p.n_pad = f.n_pad;
  p.scale = scale;
  p.bias = bias;
  p.w_own = w_own;
  p.w_oth = w_oth;
  p.lse2_a = lse2_a;
  p.lse2_b = lse2_b;
  p.g_out = reinterpret_cast<uint16_t*>(gmat);
  p.g_ld = f.n_pad;
  p.run_if = run_if;
  p.ent = run_if != nullptr ? 1 : 0;   // the guarded (fallback) run feeds emat_fallback_sums_kernel
  p.sc_part = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ws) + w.sc_part);
  if (loss_kind == LOSS_CLIP) return launch_tile<MODE_GW, LOSS_CLIP, 256, kSBN>(ma, mb, mb, p, st);
  return launch_tile<MODE_GW, LOSS_SIGLIP, 256, kSBN>(ma, mb, mb, p, st);
}

// d_out[out_rows, d] = mul * sum_k G(.,.) * F[k, d]; transposed=false contracts G's columns (out rows = G rows),
// transposed=true contracts G's rows (out rows = G columns; G read as an M-major operand).  F = feat [k_len, ld].
struct DotArgs {   // optional: dot_out += <d_out, dot_feat> / scale  (d(loss)/d(scale) by homogeneity)
  const void* dot_feat = nullptr;
  float* dot_out = nullptr;
  // optional: fused reduce-scatter, see GemmParams::peer
  const unsigned long long* peer = nullptr;
  int peer_n = 0, peer_rank = 0;
  PeerInfo sig = {nullptr, nullptr, nullptr, 0, 0};   // raise CH_DTEXT when the pushed tiles have landed
};
int run_gmat_gemm(bool transposed, const void* gmat, int g_rows, int g_cols, const void* feat, int d, int ld,
                  float coef, const float* scale, const float* grad_out, void* ws, void* d_out, int out_dtype,
                  long out_ld, const DotArgs& xf, cudaStream_t st) {
  const int out_rows = transposed ? g_cols : g_rows;
  const int k_len = transposed ? g_rows : g_cols;
  const GemmPlan g = gemm_plan(out_rows, k_len, ld, xf.peer != nullptr);
  const long g_ld = mrclip_padded_cols(g_cols);
  CUtensorMap ma, mb;
  if (int e = make_map(&ma, gmat, g_rows, g_cols, g_ld, transposed ? 64 : kBM)) return e;
  if (int e = make_map(&mb, feat, k_len, ld, ld, 64)) return e;
  GemmParams p;
  memset(&p, 0, sizeof p);
  p.m_rows = out_rows;
  p.num_rb = g.num_rb;
  p.num_dt = g.num_dt;
  p.num_kb = g.num_kb;
  p.kb_per_split = g.kb_per_split;
  p.ksplit = g.ksplit;
  p.num_items = g.num_items;
  p.m_pad = g.m_pad;
  p.d_pad = g.d_pad;
  p.dpart = reinterpret_cast<float*>(ws);
  // stream-K (CTA pairs, local output): no split-K partials, no reduce pass.  Opt-in (MRCLIP_STREAMK=1): measured on B200
  // at N=32768 it is 0.05-0.07 ms per GEMM SLOWER than split-K + reduce (profiles/r2_notes.md) -- the short split-K items
  // keep all pairs on the same K range of the feature matrix (L2), whole-K tiles drift apart.
  static const bool sk_on = [] {
    const char* e = getenv("MRCLIP_STREAMK");
    return e && atoi(e) == 1;
  }();
  const long tiles = (long)g.num_rb * g.num_dt;
  const bool streamk = sk_on && gemm_pairs() && xf.peer == nullptr && tiles <= kSkMaxTiles && ws != nullptr;
  if (streamk) {
    const long units = tiles * g.num_kb;
    p.num_items = (int)(units < kSkMaxPairs ? units : kSkMaxPairs);      // = CTA pairs to launch
    p.ksplit = 1;
    p.kb_per_split = g.num_kb;
    p.sk_part = reinterpret_cast<float*>(ws);
    p.sk_flags = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ws) + kSkPartBytes);
    p.out = d_out;
    p.out_ld = out_ld;
    p.out_dtype = out_dtype;
    p.d_valid = d;
    p.coef = coef;
    p.scale = scale;
    p.grad_out = grad_out;
    CUDA_TRY(cudaMemsetAsync(p.sk_flags, 0, (size_t)tiles * sizeof(int), st));
    if (int e = transposed ? launch_gemm2<true, 0, true>(ma, mb, p, st) : launch_gemm2<false, 0, true>(ma, mb, p, st)) return e;
    if (xf.dot_feat != nullptr) {   // <d_out, dot_feat> / scale (d logit_scale by homogeneity), formerly in the reduce pass
      long blocks = ((long)out_rows + 7) / 8;
      if (blocks > 148L * 8) blocks = 148L * 8;
      rowdot_kernel<<<(int)blocks, 256, 0, st>>>(d_out, out_dtype, out_ld, reinterpret_cast<const __nv_bfloat16*>(xf.dot_feat),
                                                 (long)ld, out_rows, d, scale, xf.dot_out);
      g_launches.fetch_add(1);
      CUDA_TRY(cudaGetLastError());
    }
    return 0;
  }
  const bool direct = (g.ksplit == 1 && xf.dot_feat == nullptr);
  if (xf.peer && !direct) return fail(-1, "push epilogue needs a single-split GEMM");
  if (direct) {
    p.peer = xf.peer;
    p.peer_n = xf.peer_n;
    p.peer_rank = xf.peer_rank;
    p.sig = xf.sig;
    p.out = d_out;
    p.out_ld = out_ld;
    p.out_dtype = out_dtype;
    p.d_valid = d;
    p.coef = coef;
    p.scale = scale;
    p.grad_out = grad_out;
  }
  if (int e = transposed ? launch_gemm<true>(ma, mb, p, st) : launch_gemm<false>(ma, mb, p, st)) return e;
  if (direct) return 0;
  const long total = (long)out_rows * (g.d_pad / 4);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  grad_reduce_kernel<<<(int)blocks, 256, 0, st>>>(p.dpart, g.ksplit, out_rows, d, g.m_pad, g.d_pad, coef, scale,
                                                   grad_out, d_out, out_dtype, out_ld,
                                                   reinterpret_cast<const __nv_bfloat16*>(xf.dot_feat), ld, xf.dot_out);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int run_scalar_reduce(const mrclip_shape& sh, void* ws, float cx, float cy, const float* grad_out, float* d_scale,
                      float* d_bias, int accumulate, int fold, cudaStream_t st) {
  if (!d_scale && !d_bias) return 0;
  const FwdPlan f = fwd_plan(sh.m_rows, sh.n_cols);
  const WsLayout w = ws_layout(sh.m_rows, sh.n_cols, sh.d);
  scalar_reduce_kernel<<<1, 1024, 0, st>>>(
      reinterpret_cast<const float2*>(reinterpret_cast<uint8_t*>(ws) + w.sc_part),
      (long)f.num_rb * f.total_chunks * kEpiWarps, cx, cy, grad_out, nullptr, d_scale, d_bias, accumulate, fold);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

int mrclip_version(void) { return 100; }
const char* mrclip_last_error(void) { return g_err.c_str(); }
long mrclip_launch_count(void) { return g_launches.load(); }

int mrclip_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int mrclip_padded_dim(int d) { return (d + 7) / 8 * 8; }
int mrclip_padded_cols(int n_cols) { return (n_cols + kSBN - 1) / kSBN * kSBN; }

size_t mrclip_workspace_bytes(int m_rows, int n_cols, int d) {
  if (m_rows <= 0 || n_cols <= 0 || d <= 0) return 0;
  return ws_layout(m_rows, n_cols, d).total;
}

int mrclip_fwd_col_granule(int m_rows, int n_cols) {
  if (m_rows <= 0 || n_cols <= 0) return kSBN;
  return fwd_plan(m_rows, n_cols).tiles_per_chunk * kSBN;
}

int mrclip_pack_bf16(const void* src, int src_dtype, int rows, int d, long src_ld, void* dst, int dst_ld,
                     void* stream) {
  if (rows <= 0 || d <= 0) return fail(-1, "empty pack");
  if (src_dtype < 0 || src_dtype > 2) return fail(-1, "bad dtype %d", src_dtype);
  if (dst_ld < d) return fail(-1, "dst_ld < d");
  const bool vec = d % 8 == 0 && src_ld % 8 == 0 && dst_ld % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  const long total = vec ? (long)rows * (dst_ld / 8) : (long)rows * dst_ld;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 32) blocks = 148L * 32;
  if (vec)
    pack_bf16_vec8_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, src_dtype, rows, d, src_ld,
                                                                         (__nv_bfloat16*)dst, dst_ld);
  else
    pack_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, src_dtype, rows, d, src_ld,
                                                                    (__nv_bfloat16*)dst, dst_ld);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_transpose_bf16(const void* src, int rows, int cols, long src_ld, void* dst, long dst_ld,
                          void* stream) {
  if (rows <= 0 || cols <= 0) return fail(-1, "empty transpose");
  dim3 grid((cols + 63) / 64, (rows + 63) / 64), block(32, 8);
  transpose_bf16_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, rows, cols,
                                                                  src_ld, (__nv_bfloat16*)dst, dst_ld);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_clip_fwd_tiles(const void* a_rows, const void* b_all, mrclip_shape shape, int ld,
                          const float* scale, int col_begin, int col_end, void* ws, void* stream) {
  return run_fwd(LOSS_CLIP, a_rows, b_all, shape, ld, scale, nullptr, col_begin, col_end, ws, nullptr,
                 (cudaStream_t)stream);
}

int mrclip_clip_fwd_reduce(mrclip_shape sh, void* ws, float* lse2_row, float* col_m, float* col_l,
                           float* diag2, void* stream) {
  if (int e = check_shape(sh, mrclip_padded_dim(sh.d))) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const FwdPlan f = fwd_plan(sh.m_rows, sh.n_cols);
  const WsLayout w = ws_layout(sh.m_rows, sh.n_cols, sh.d);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  reduce_rows_kernel<<<ceil_div(sh.m_rows, 256), 256, 0, st>>>(
      reinterpret_cast<const float2*>(wsb + w.row_part), f.total_chunks * 2, sh.m_rows, f.m_pad, lse2_row);
  {
    const int bands = ceil_div(sh.m_rows, 32);
    static bool attr_set = false;
    if (!attr_set && bands * sizeof(float) > 48 * 1024) {
      CUDA_TRY(cudaFuncSetAttribute(reduce_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    if (bands * sizeof(float) > 200 * 1024) return fail(-1, "m_rows=%d too large for the column reduce", sh.m_rows);
    reduce_cols_kernel<<<ceil_div(sh.n_cols, 64), 1024, bands * sizeof(float), st>>>(
        reinterpret_cast<const float*>(wsb + w.col_l), reinterpret_cast<const float*>(wsb + w.col_c), bands,
        sh.n_cols, f.n_pad, col_m, col_l);
  }
  g_launches.fetch_add(2);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(diag2, wsb + w.diag2, (size_t)sh.m_rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int mrclip_lse2_merge(const float* part_m, const float* part_l, int parts, long part_stride, int n_cols,
                      float* lse2_out, void* stream) {
  if (parts <= 0 || n_cols <= 0) return fail(-1, "empty merge");
  const int n_pad = mrclip_padded_cols(n_cols);
  merge_parts_kernel<<<ceil_div(n_pad, 256), 256, 0, (cudaStream_t)stream>>>(part_m, part_l, parts, n_cols,
                                                                            part_stride, n_pad, lse2_out);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_clip_loss(const float* lse2_row, const float* lse2_col, const float* diag2, int m_rows,
                     int label_offset, float* loss, void* stream) {
  if (m_rows <= 0) return fail(-1, "empty loss");
  clip_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(lse2_row, lse2_col, diag2, m_rows, label_offset,
                                                         loss);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_clip_bwd(const void* a_rows, const void* b_all, const void* bt_all, long bt_ld,
                    mrclip_shape shape, int ld, const float* lse2_a, const float* lse2_b,
                    const float* scale, float w_own, float w_oth, float coef, const float* grad_out,
                    void* ws, void* d_a, int out_dtype, long out_ld, float* d_scale,
                    int accumulate_scalars, void* stream) {
  return run_bwd(LOSS_CLIP, a_rows, b_all, bt_all, bt_ld, shape, ld, lse2_a, lse2_b, scale, nullptr, w_own,
                 w_oth, coef, grad_out, ws, d_a, out_dtype, out_ld, d_scale, nullptr, accumulate_scalars,
                 (cudaStream_t)stream);
}

int mrclip_siglip_fwd(const void* a_rows, const void* b_all, mrclip_shape shape, int ld,
                      const float* scale, const float* bias, void* ws, float* loss, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = run_fwd(LOSS_SIGLIP, a_rows, b_all, shape, ld, scale, bias, 0, shape.n_cols, ws, nullptr, st)) return e;
  const FwdPlan f = fwd_plan(shape.m_rows, shape.n_cols);
  const WsLayout w = ws_layout(shape.m_rows, shape.n_cols, shape.d);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  scalar_reduce_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const float2*>(wsb + w.sc_part),
                                           (long)f.num_rb * f.total_chunks * kEpiWarps,
                                           1.f / (float)shape.m_rows, 0.f, nullptr, nullptr, loss, nullptr, 0, 0);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_siglip_bwd(const void* a_rows, const void* b_all, const void* bt_all, long bt_ld,
                      mrclip_shape shape, int ld, const float* scale, const float* bias, float coef,
                      const float* grad_out, void* ws, void* d_a, int out_dtype, long out_ld,
                      float* d_scale, float* d_bias, int accumulate_scalars, void* stream) {
  return run_bwd(LOSS_SIGLIP, a_rows, b_all, bt_all, bt_ld, shape, ld, nullptr, nullptr, scale, bias, 1.f, 0.f,
                 coef, grad_out, ws, d_a, out_dtype, out_ld, d_scale, d_bias, accumulate_scalars,
                 (cudaStream_t)stream);
}

size_t mrclip_gmat_bytes(int m_rows, int n_cols) {
  if (m_rows <= 0 || n_cols <= 0) return 0;
  return (size_t)ceil_div(m_rows, kBM) * kBM * (size_t)mrclip_padded_cols(n_cols) * 2;
}

int mrclip_clip_gwrite(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* lse2_a,
                       const float* lse2_b, const float* scale, float w_own, float w_oth, float coef,
                       const float* grad_out, void* ws, void* gmat, float* d_scale, int accumulate_scalars,
                       int both_directions, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = run_gwrite(LOSS_CLIP, a_rows, b_all, shape, ld, lse2_a, lse2_b, scale, nullptr, w_own, w_oth, ws, gmat, nullptr, st))
    return e;
  return run_scalar_reduce(shape, ws, coef * w_own, both_directions ? coef * w_oth : 0.f, grad_out, d_scale, nullptr,
                           accumulate_scalars, 1, st);
}

int mrclip_siglip_gwrite(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* scale,
                         const float* bias, float coef, const float* grad_out, void* ws, void* gmat, float* d_scale,
                         float* d_bias, int accumulate_scalars, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = run_gwrite(LOSS_SIGLIP, a_rows, b_all, shape, ld, nullptr, nullptr, scale, bias, 1.f, 0.f, ws, gmat, nullptr, st))
    return e;
  return run_scalar_reduce(shape, ws, coef, coef, grad_out, d_scale, d_bias, accumulate_scalars, 0, st);
}

int mrclip_gmat_gemm(int transposed, const void* gmat, mrclip_shape shape, const void* feat, int ld, float coef,
                     const float* scale, const float* grad_out, void* ws, void* d_out, int out_dtype, long out_ld,
                     void* stream) {
  if (int e = check_shape(shape, ld)) return e;
  if (out_dtype < 0 || out_dtype > 2) return fail(-1, "bad out_dtype %d", out_dtype);
  return run_gmat_gemm(transposed != 0, gmat, shape.m_rows, shape.n_cols, feat, shape.d, ld, coef, scale, grad_out,
                       ws, d_out, out_dtype, out_ld, DotArgs(), (cudaStream_t)stream);
}

/* ---- "emat" backend: the forward keeps E (CLIP) / G (SigLIP); no recompute in the backward ------------ */
int mrclip_clip_fwd_tiles_e(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* scale,
                            int col_begin, int col_end, void* ws, void* emat, void* stream) {
  if (!emat) return fail(-1, "emat is NULL");
  return run_fwd(LOSS_CLIP, a_rows, b_all, shape, ld, scale, nullptr, col_begin, col_end, ws, emat,
                 (cudaStream_t)stream);
}

/* ---- d logit_scale of a multi-rank local loss without entropy arithmetic in the rescale pass (MRCLIP_DS=fwd; not
 *      validated on hardware yet): the forward also keeps u = sum_j 2^(S2 - m) S2 per (row, column-chunk half) ---- */
int mrclip_fwd_row_ent_ok(int m_rows, int n_cols, int n_per_rank) {
  if (m_rows <= 0 || n_cols <= 0 || n_per_rank <= 0 || n_cols % n_per_rank != 0) return 0;
  const FwdPlan f = fwd_plan(m_rows, n_cols);
  return n_per_rank % (f.tiles_per_chunk * kSBN) == 0 ? 1 : 0;   // every column chunk has one owner
}

int mrclip_clip_fwd_tiles_eu(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* scale,
                             int col_begin, int col_end, void* ws, void* emat, void* stream) {
  if (!emat) return fail(-1, "emat is NULL");
  return run_fwd(LOSS_CLIP, a_rows, b_all, shape, ld, scale, nullptr, col_begin, col_end, ws, emat,
                 (cudaStream_t)stream, true);
}

int mrclip_row_ent_split(mrclip_shape sh, void* ws, const float* lse2_row, int n_per_rank, int ranks, float* out_slots,
                         void* stream) {
  if (int e = check_shape(sh, mrclip_padded_dim(sh.d))) return e;
  if (!lse2_row || !out_slots || ranks <= 0 || ranks > 1024 || n_per_rank * ranks != sh.n_cols ||
      !mrclip_fwd_row_ent_ok(sh.m_rows, sh.n_cols, n_per_rank))
    return fail(-1, "row_ent_split: column chunks do not split by owner (n_per_rank=%d, ranks=%d, n_cols=%d)",
                n_per_rank, ranks, sh.n_cols);
  cudaStream_t st = (cudaStream_t)stream;
  const FwdPlan f = fwd_plan(sh.m_rows, sh.n_cols);
  const WsLayout w = ws_layout(sh.m_rows, sh.n_cols, sh.d);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  CUDA_TRY(cudaMemsetAsync(out_slots, 0, (size_t)64 * 2 * ranks * sizeof(float), st));
  const int slots_per_rank = 2 * (n_per_rank / (f.tiles_per_chunk * kSBN));
  row_ent_split_kernel<<<ceil_div(sh.m_rows, 256), 256, ranks * sizeof(float), st>>>(
      reinterpret_cast<const float2*>(wsb + w.row_part), reinterpret_cast<const float*>(wsb + w.row_ent),
      f.total_chunks * 2, slots_per_rank, ranks, sh.m_rows, f.m_pad, lse2_row, out_slots);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_sum_slots_dot(const float* slots, int nslots, int rows, int d, void* out, int out_dtype, long out_ld,
                         const void* feat, long feat_ld, float* dot_slots, void* stream) {
  if (!slots || !out || !feat || !dot_slots || nslots <= 0 || rows <= 0 || d <= 0 || feat_ld < d)
    return fail(-1, "sum_slots_dot: bad arguments");
  if (out_dtype < 0 || out_dtype > 2) return fail(-1, "bad out_dtype %d", out_dtype);
  if (d % 4 != 0) return fail(-1, "sum_slots_dot: d=%d must be a multiple of 4", d);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(dot_slots, 0, 64 * sizeof(float), st));
  const long total = (long)rows * (d / 4);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  sum_slots_vec_kernel<false><<<(int)blocks, 256, 0, st>>>(slots, nslots, rows, d, out, out_dtype, out_ld,
                                                           reinterpret_cast<const __nv_bfloat16*>(feat), feat_ld, dot_slots);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_emat_check(mrclip_shape sh, void* ws, const float* lse2_row, const float* lse2_col, void* stream) {
  if (int e = check_shape(sh, mrclip_padded_dim(sh.d))) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const FwdPlan f = fwd_plan(sh.m_rows, sh.n_cols);
  const WsLayout w = ws_layout(sh.m_rows, sh.n_cols, sh.d);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  const int ncb = f.n_pad / 64;
  float* cbmin = reinterpret_cast<float*>(wsb + w.cbmin);
  int* flag = reinterpret_cast<int*>(wsb + w.flag);
  emat_cbmin_kernel<<<ceil_div(ncb * 32, 256), 256, 0, st>>>(lse2_col, ncb, cbmin, flag);
  emat_check_kernel<<<f.bands, 256, 0, st>>>(reinterpret_cast<const float*>(wsb + w.col_c), lse2_row, sh.m_rows,
                                             cbmin, ncb, 80.f, flag);
  g_launches.fetch_add(2);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

const int* mrclip_emat_flag(mrclip_shape sh, void* ws) {
  if (sh.m_rows <= 0 || sh.n_cols <= 0 || sh.d <= 0 || !ws) return nullptr;
  return reinterpret_cast<const int*>(reinterpret_cast<uint8_t*>(ws) + ws_layout(sh.m_rows, sh.n_cols, sh.d).flag);
}

int mrclip_clip_gwrite_if(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* lse2_a,
                          const float* lse2_b, const float* scale, float w_own, float w_oth, void* ws, void* gmat,
                          const int* run_if, void* stream) {
  return run_gwrite(LOSS_CLIP, a_rows, b_all, shape, ld, lse2_a, lse2_b, scale, nullptr, w_own, w_oth, ws, gmat,
                    run_if, (cudaStream_t)stream);
}

// bands [band0, band0 + nbands) of 32 rows each (nbands <= 0: all of them); zero_msums: clear the sums first
static int run_emat_transform(const mrclip_shape& sh, void* ws, void* emat, const float* lse2_row, const float* lse2_col,
                       const float* diag2, float w_row, float w_col, const int* skip_if, float* msums, int msum_slots,
                       int n_per_rank, int ranks, int band0, int nbands, bool zero_msums, cudaStream_t st) {
  if (int e = check_shape(sh, mrclip_padded_dim(sh.d))) return e;
  if (!emat || !lse2_row || !lse2_col || !diag2) return fail(-1, "emat_transform: NULL argument");
  if (msums && msum_slots <= 0) return fail(-1, "emat_transform: msum_slots must be positive");
  if (msums && (ranks > 64 || (ranks > 1 && n_per_rank < 8)))
    return fail(-1, "emat_transform: split sums need ranks <= 64 and n_per_rank >= 8 (got %d x %d)", ranks, n_per_rank);
  if (msums && (n_per_rank <= 0 || ranks <= 0 || (long)n_per_rank * ranks != sh.n_cols))
    return fail(-1, "emat_transform: n_per_rank * ranks must equal n_cols");
  const FwdPlan f = fwd_plan(sh.m_rows, sh.n_cols);
  const WsLayout w = ws_layout(sh.m_rows, sh.n_cols, sh.d);
  const float* colc = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(ws) + w.col_c);
  if (nbands <= 0) {
    band0 = 0;
    nbands = f.bands;
  }
  if (band0 < 0 || band0 + nbands > f.bands) return fail(-1, "emat_transform: bad band range");
  dim3 grid(ceil_div(f.n_pad, 1024), nbands);
  if (msums) {
    if (zero_msums) CUDA_TRY(cudaMemsetAsync(msums, 0, sizeof(float) * 2 * ranks * msum_slots, st));
    emat_transform_kernel<true><<<grid, 256, 0, st>>>(reinterpret_cast<uint16_t*>(emat), (long)f.n_pad, sh.m_rows,
                                                      f.n_pad, colc, f.n_pad / 64, lse2_row, lse2_col, diag2,
                                                      sh.label_offset, w_row, w_col, skip_if, msums, n_per_rank, ranks,
                                                      msum_slots, band0);
    if (skip_if && band0 + nbands == f.bands) {   // guard raised: the sums come from the exact recompute's partials instead
      const int items = f.num_rb * f.total_chunks;
      emat_fallback_sums_kernel<<<ceil_div(items, 256), 256, 0, st>>>(
          skip_if, reinterpret_cast<const float2*>(reinterpret_cast<const uint8_t*>(ws) + w.sc_part), f.num_rb,
          f.total_chunks, f.tiles_per_chunk * kSBN, n_per_rank, ranks, w_row, w_col, msums);
      g_launches.fetch_add(1);
    }
  } else {
    emat_transform_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<uint16_t*>(emat), (long)f.n_pad, sh.m_rows,
                                                       f.n_pad, colc, f.n_pad / 64, lse2_row, lse2_col, diag2,
                                                       sh.label_offset, w_row, w_col, skip_if, nullptr, 1, 1, 1, band0);
  }
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int mrclip_emat_transform(mrclip_shape sh, void* ws, void* emat, const float* lse2_row, const float* lse2_col,
                                     const float* diag2, const float* scale, float w_row, float w_col, const int* skip_if,
                                     float* msums, int msum_slots, int n_per_rank, int ranks, void* stream) {
  (void)scale;
  return run_emat_transform(sh, ws, emat, lse2_row, lse2_col, diag2, w_row, w_col, skip_if, msums, msum_slots, n_per_rank,
                            ranks, 0, 0, true, (cudaStream_t)stream);
}

int mrclip_gmat_gemm_dot(int transposed, const void* gmat, mrclip_shape shape, const void* feat, int ld, float coef,
                         const float* scale, const float* grad_out, void* ws, void* d_out, int out_dtype, long out_ld,
                         const void* dot_feat, float* dot_out, void* stream) {
  if (int e = check_shape(shape, ld)) return e;
  if (out_dtype < 0 || out_dtype > 2) return fail(-1, "bad out_dtype %d", out_dtype);
  DotArgs xf;
  xf.dot_feat = dot_feat;
  xf.dot_out = dot_out;
  return run_gmat_gemm(transposed != 0, gmat, shape.m_rows, shape.n_cols, feat, shape.d, ld, coef, scale, grad_out,
                       ws, d_out, out_dtype, out_ld, xf, (cudaStream_t)stream);
}

int mrclip_gmat_gemm_push(const void* gmat, mrclip_shape shape, const void* feat, int ld, float coef,
                          const float* scale, const float* grad_out, void* ws, const unsigned long long* peer_bufs,
                          int n_per_rank, int my_rank, void* stream) {
  if (int e = check_shape(shape, ld)) return e;
  if (!peer_bufs || n_per_rank <= 0 || shape.n_cols % n_per_rank != 0 || my_rank < 0 ||
      my_rank >= shape.n_cols / n_per_rank)
    return fail(-1, "gmat_gemm_push: bad peer layout (n_per_rank=%d, rank=%d, n_cols=%d)", n_per_rank, my_rank, shape.n_cols);
  DotArgs xf;
  xf.peer = peer_bufs;
  xf.peer_n = n_per_rank;
  xf.peer_rank = my_rank;
  // d_out only marks the epilogue as direct; rows are routed through peer_bufs (dense fp32, leading dim d)
  return run_gmat_gemm(true, gmat, shape.m_rows, shape.n_cols, feat, shape.d, ld, coef, scale, grad_out, ws,
                       const_cast<unsigned long long*>(peer_bufs), MRCLIP_DT_F32, shape.d, xf, (cudaStream_t)stream);
}

/* mrclip_gmat_gemm_push with a bf16 payload: peer_bufs are bf16 [W, n_per_rank, d] receive buffers (MRCLIP_PUSH_DTYPE=bf16,
 * not validated on hardware yet) */
int mrclip_gmat_gemm_push_bf16(const void* gmat, mrclip_shape shape, const void* feat, int ld, float coef,
                               const float* scale, const float* grad_out, void* ws, const unsigned long long* peer_bufs,
                               int n_per_rank, int my_rank, void* stream) {
  if (int e = check_shape(shape, ld)) return e;
  if (!peer_bufs || n_per_rank <= 0 || shape.n_cols % n_per_rank != 0 || my_rank < 0 ||
      my_rank >= shape.n_cols / n_per_rank)
    return fail(-1, "gmat_gemm_push_bf16: bad peer layout (n_per_rank=%d, rank=%d, n_cols=%d)", n_per_rank, my_rank, shape.n_cols);
  DotArgs xf;
  xf.peer = peer_bufs;
  xf.peer_n = n_per_rank;
  xf.peer_rank = my_rank;
  return run_gmat_gemm(true, gmat, shape.m_rows, shape.n_cols, feat, shape.d, ld, coef, scale, grad_out, ws,
                       const_cast<unsigned long long*>(peer_bufs), MRCLIP_DT_BF16, shape.d, xf, (cudaStream_t)stream);
}

/* mrclip_sum_slots / mrclip_sum_slots_dot over bf16 slots (feat / dot_slots optional) */
int mrclip_sum_slots_bf16(const void* slots, int nslots, int rows, int d, void* out, int out_dtype, long out_ld,
                          const void* feat, long feat_ld, float* dot_slots, void* stream) {
  if (!slots || !out || nslots <= 0 || rows <= 0 || d <= 0) return fail(-1, "sum_slots_bf16: bad arguments");
  if ((feat == nullptr) != (dot_slots == nullptr) || (feat && feat_ld < d)) return fail(-1, "sum_slots_bf16: feat and dot_slots go together");
  if (out_dtype < 0 || out_dtype > 2) return fail(-1, "bad out_dtype %d", out_dtype);
  if (d % 4 != 0) return fail(-1, "sum_slots_bf16: d=%d must be a multiple of 4", d);
  cudaStream_t st = (cudaStream_t)stream;
  if (dot_slots) CUDA_TRY(cudaMemsetAsync(dot_slots, 0, 64 * sizeof(float), st));
  const long total = (long)rows * (d / 4);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  sum_slots_vec_kernel<true><<<(int)blocks, 256, 0, st>>>(slots, nslots, rows, d, out, out_dtype, out_ld,
                                                          reinterpret_cast<const __nv_bfloat16*>(feat), feat_ld, dot_slots);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_push_copy(const void* src, size_t bytes, const unsigned long long* peer_bufs, int ranks, size_t dst_offset,
                     int skip_rank, void* stream) {
  if (!src || !peer_bufs || ranks <= 0 || bytes == 0) return fail(-1, "push_copy: bad arguments");
  if ((bytes & 15) || (dst_offset & 15) || (reinterpret_cast<uintptr_t>(src) & 15))
    return fail(-1, "push_copy: 16-byte alignment required (bytes=%zu, offset=%zu)", bytes, dst_offset);
  const long n16 = (long)(bytes / 16);
  long bx = (n16 + 255) / 256;
  const long cap = (148L * 8 + ranks - 1) / ranks;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  push_copy_kernel<<<dim3((unsigned)bx, (unsigned)ranks), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(src), n16, peer_bufs, (long)dst_offset, skip_rank);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_sum_slots(const float* slots, int nslots, int rows, int d, void* out, int out_dtype, long out_ld,
                     void* stream) {
  if (!slots || !out || nslots <= 0 || rows <= 0 || d <= 0) return fail(-1, "sum_slots: bad arguments");
  if (out_dtype < 0 || out_dtype > 2) return fail(-1, "bad out_dtype %d", out_dtype);
  const long total = (long)rows * d;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  sum_slots_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(slots, nslots, rows, d, out, out_dtype, out_ld);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_siglip_fwd_e(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const float* scale,
                        const float* bias, void* ws, float* loss, void* gmat, void* stream) {
  if (!gmat) return fail(-1, "gmat is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = run_fwd(LOSS_SIGLIP, a_rows, b_all, shape, ld, scale, bias, 0, shape.n_cols, ws, gmat, st)) return e;
  const FwdPlan f = fwd_plan(shape.m_rows, shape.n_cols);
  const WsLayout w = ws_layout(shape.m_rows, shape.n_cols, shape.d);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  scalar_reduce_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const float2*>(wsb + w.sc_part),
                                           (long)f.num_rb * f.total_chunks * kEpiWarps,
                                           1.f / (float)shape.m_rows, 0.f, nullptr, nullptr, loss, nullptr, 0, 0);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

/* d_scale / d_bias of SigLipLoss from the partials mrclip_siglip_fwd_e left in ws */
int mrclip_siglip_e_scalars(mrclip_shape shape, void* ws, float coef, const float* grad_out, float* d_scale,
                            float* d_bias, int accumulate_scalars, void* stream) {
  if (!d_scale && !d_bias) return 0;
  const FwdPlan f = fwd_plan(shape.m_rows, shape.n_cols);
  const WsLayout w = ws_layout(shape.m_rows, shape.n_cols, shape.d);
  scalar_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(reinterpret_cast<uint8_t*>(ws) + w.sc_part2),
      (long)f.num_rb * f.total_chunks * kEpiWarps, coef, coef, grad_out, nullptr, d_scale, d_bias, accumulate_scalars, 0);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"

/* ---- whole-step entries (include/mrclip.h): every kernel of a forward / backward behind one call --------------- */
namespace {
constexpr int kSmallDot = 0, kSmallR2Row = 64, kSmallLoss = 65, kSmallAcc = 128;
constexpr size_t kSmallFloats = 128 + 64 * 64 + 128;
constexpr int kCtlRowEntDone = 16;

PeerInfo make_peer(const mrclip_step* s) {
  PeerInfo pi;
  pi.sig_peers = s->peer.ctl_block_peers;
  pi.sig_local = reinterpret_cast<int*>(s->peer.ctl_block);
  pi.ctl = reinterpret_cast<PeerCtl*>(s->peer.ctl);
  pi.ranks = s->peer.ranks > 1 ? s->peer.ranks : 1;
  pi.rank = s->peer.ranks > 1 ? s->peer.rank : 0;
  return pi;
}

int check_step(const mrclip_step* s, bool backward) {
  if (!s) return fail(-1, "step: NULL descriptor");
  if (int e = check_shape(s->shape, s->ld)) return e;
  const int ranks = s->peer.ranks > 1 ? s->peer.ranks : 1;
  if (s->ld != mrclip_padded_dim(s->shape.d)) return fail(-1, "step: ld=%d must be mrclip_padded_dim(d)=%d", s->ld, mrclip_padded_dim(s->shape.d));
  if ((long)s->shape.m_rows * ranks != s->shape.n_cols) return fail(-1, "step: n_cols=%d must be ranks*m_rows=%d*%d", s->shape.n_cols, ranks, s->shape.m_rows);
  if (!s->img_rows || !s->txt_all || !s->ws || !s->small) return fail(-1, "step: NULL buffer");
  if (s->kind == 0 && (!s->stats || !s->lse2_row_all || !s->lse2_col_all || !s->msums)) return fail(-1, "step: NULL statistics buffer");
  if (ranks > 1) {
    const mrclip_peer& p = s->peer;
    if (ranks > kPeerMaxRanks) return fail(-1, "step: at most %d ranks", kPeerMaxRanks);
    if (p.rank < 0 || p.rank >= ranks || s->shape.label_offset != p.rank * s->shape.m_rows) return fail(-1, "step: bad rank / label_offset");
    if (!p.ctl_block_peers || !p.ctl_block || !p.ctl || !p.txt_peers || (backward && (!p.recv_peers || !p.recv)) ||
        (s->kind == 0 && !p.stats_peers))
      return fail(-1, "step: NULL peer buffer");
    if (!gemm_pairs()) return fail(-1, "step: the multi-rank step needs the CTA-pair GEMM (unset MRCLIP_GEMM_CTA)");
  }
  return 0;
}

struct SideStream;
SideStream* side_stream();
int push_rows_async(const mrclip_step* s, const PeerInfo& pi, cudaStream_t st);
bool step_overlaps(const mrclip_step* s, const PeerInfo& pi);
// overlap: the peer stores of the text rows run on a side stream (push_rows_kernel) while the forward starts on this
// rank's own columns; otherwise the pack kernel itself stores to every rank.  MRCLIP_AG_OVERLAP=0 switches it off.
bool ag_overlap() {
  const char* e = getenv("MRCLIP_AG_OVERLAP");
  return !(e && atoi(e) == 0);
}
int step_pack(const mrclip_step* s, const PeerInfo& pi, const void* img, int img_dtype, long img_ld, const void* txt,
              int txt_dtype, long txt_ld, const float* log_scale, int raw, cudaStream_t st) {
  if (pi.ranks > 1 && s->peer.host_epoch) *s->peer.host_epoch += 1;     // host twin of the CH_TEXT epoch
  if (img_dtype < 0 || img_dtype > 2 || txt_dtype < 0 || txt_dtype > 2) return fail(-1, "step: bad feature dtype");
  if (img_ld < s->shape.d || txt_ld < s->shape.d) return fail(-1, "step: feature leading dimension < d");
  Pack2Params pp;
  memset(&pp, 0, sizeof pp);
  pp.img_src = img;
  pp.txt_src = txt;
  pp.img_dtype = img_dtype;
  pp.txt_dtype = txt_dtype;
  pp.rows = s->shape.m_rows;
  pp.d = s->shape.d;
  pp.ld = s->ld;
  pp.img_src_ld = img_ld;
  pp.txt_src_ld = txt_ld;
  pp.img_dst = reinterpret_cast<__nv_bfloat16*>(s->img_rows);
  pp.txt_peers = s->peer.txt_peers;
  pp.txt_local = reinterpret_cast<__nv_bfloat16*>(s->txt_all);
  pp.row0 = s->shape.label_offset;
  const int vec_ok = (img_ld % 8 == 0 && txt_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(txt) & 15) == 0) ? 1 : 0;
  const bool overlap = step_overlaps(s, pi);
  const int push_mode = (pi.ranks > 1 && !overlap) ? PACK_PUSH : PACK_LOCAL;
  if (raw) {   // un-normalised tower outputs: one warp per row
    if (!s->inv_norm || !s->scale_buf) return fail(-1, "step: raw forward needs inv_norm and scale_buf");
    long blocks = (2L * pp.rows + 7) / 8;
    if (blocks > 148L * 8) blocks = 148L * 8;
    packnorm2_push_kernel<<<(int)blocks, 256, 0, st>>>(pp, pi, vec_ok, push_mode, s->inv_norm, log_scale, s->scale_buf);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return overlap ? push_rows_async(s, pi, st) : 0;
  }
  const long total = 2L * pp.rows * (pp.ld / 8);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 8) blocks = 148L * 8;
  pack2_push_kernel<<<(int)blocks, 256, 0, st>>>(pp, pi, vec_ok, push_mode);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return overlap ? push_rows_async(s, pi, st) : 0;
}

// Second stream + events of the banded backward (rescale pass of band b+1 overlapped with the dI GEMM of band b).
// One set per device, created on first use; fork / join through events only, so the pattern is graph-capturable.
constexpr int kMaxBands = 16;
struct SideStream {
  cudaStream_t st = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, ev[kMaxBands] = {};
  cudaStream_t cp[3] = {};          // copy-engine lanes of the text all-gather (several copies in flight)
  cudaEvent_t cp_join[3] = {};
};
SideStream* side_stream() {
  static std::mutex mu;
  static std::map<int, SideStream*> per_dev;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  auto it = per_dev.find(dev);
  if (it != per_dev.end()) return it->second;
  SideStream* s = new SideStream();
  bool ok = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&s->join, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < kMaxBands; ++i) ok = cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < 3; ++i)
    ok = cudaStreamCreateWithFlags(&s->cp[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&s->cp_join[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    delete s;
    s = nullptr;
  }
  per_dev[dev] = s;
  return s;
}
// Row bands of the backward: the rescale pass is HBM-bound and the dI GEMM tensor-bound, so band b+1 is rescaled on the
// side stream while band b is contracted.  MRCLIP_BANDS overrides (1 = off).
int pick_bands(int n) {
  if (const char* e = getenv("MRCLIP_BANDS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= kMaxBands) return v;
  }
  (void)n;
  return 1;   // measured on B200 (profiles/r2/r2c_*): no gain yet -- the two kernels contend for the SMs' registers
}

// cuStreamWriteValue32: a 32-bit store in stream order, executed by the front end -- no SM, no kernel
typedef CUresult (*WriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
WriteValue32Fn write_value32_fn() {
  static WriteValue32Fn fn = nullptr;
  static bool tried = false;
  if (tried) return fn;
  tried = true;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    fn = reinterpret_cast<WriteValue32Fn>(p);
  return fn;
}
// The all-gather proper, on the copy engines: this rank's packed text rows go to the same place of every other rank's
// buffer, destination by destination in the order in which the destinations will need them (rank r works through the
// sources r, r+1, r+2, ... so source q serves q-1 first, then q-2, ...), each copy followed by a stream memory
// operation that raises that destination's CH_TEXT flag.  The forward kernel, which fills every SM's shared memory,
// meanwhile runs on this rank's own columns.
int push_rows_async(const mrclip_step* s, const PeerInfo& pi, cudaStream_t st) {
  SideStream* ss = side_stream();
  WriteValue32Fn wv = write_value32_fn();
  if (!ss || !wv) return fail(-1, "step: no side stream / stream memory operations");
  CUDA_TRY(cudaEventRecord(ss->fork, st));
  const size_t bytes = (size_t)s->shape.m_rows * s->ld * 2, offset = (size_t)s->shape.label_offset * s->ld * 2;
  const int e = *s->peer.host_epoch;          // already advanced for this step (mrclip_step_forward)
  const int lanes = pi.ranks - 1 < 3 ? pi.ranks - 1 : 3;
  for (int l = 0; l < lanes; ++l) CUDA_TRY(cudaStreamWaitEvent(ss->cp[l], ss->fork, 0));
  ProfScope ps("push_rows(copy engines)", ss->cp[0]);
  for (int k = 1; k < pi.ranks; ++k) {        // destination k-1 needs the rows first; three copies in flight
    const int dest = (pi.rank - k + pi.ranks) % pi.ranks;
    cudaStream_t cs = ss->cp[(k - 1) % lanes];
    CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<void*>(s->peer.txt_peers_host[dest] + offset),
                             reinterpret_cast<const uint8_t*>(s->txt_all) + offset, bytes, cudaMemcpyDeviceToDevice, cs));
    const CUresult r = wv((CUstream)cs,
                          (CUdeviceptr)(s->peer.ctl_block_peers_host[dest] + (size_t)(CH_TEXT * kPeerMaxRanks + pi.rank) * 4),
                          (cuuint32_t)e, 0);
    if (r != CUDA_SUCCESS) return fail(-4, "cuStreamWriteValue32 failed with CUresult %d", (int)r);
  }
  for (int l = 0; l < lanes; ++l) CUDA_TRY(cudaEventRecord(ss->cp_join[l], ss->cp[l]));
  return 0;
}
// the forward's later kernels (and everything after them) are ordered behind the side-stream push
int push_rows_join(const mrclip_step* s, const PeerInfo& pi, cudaStream_t st) {
  if (step_overlaps(s, pi)) {
    SideStream* ss = side_stream();
    const int lanes = pi.ranks - 1 < 3 ? pi.ranks - 1 : 3;
    for (int l = 0; ss && l < lanes; ++l) CUDA_TRY(cudaStreamWaitEvent(st, ss->cp_join[l], 0));
  }
  return 0;
}

// Worth it only when a destination's share is large: below ~4 MB the forward on the local columns is shorter than the
// copies it would hide (c4 at W=8: 0.48 vs 0.42 ms) and the 2 (W-1) driver calls cost more host time than the step has
// (c2 at W=8: 0.47 vs 0.31 ms) -- measured, profiles/r2_notes.md.  MRCLIP_AG_OVERLAP_MIN_BYTES overrides.
size_t ag_overlap_min_bytes() {
  const char* e = getenv("MRCLIP_AG_OVERLAP_MIN_BYTES");
  return e ? (size_t)atoll(e) : (size_t)4 << 20;
}
bool step_overlaps(const mrclip_step* s, const PeerInfo& pi) {
  return pi.ranks > 1 && ag_overlap() && (size_t)s->shape.m_rows * s->ld * 2 >= ag_overlap_min_bytes() && s->peer.txt_peers_host && s->peer.ctl_block_peers_host && s->peer.host_epoch &&
         side_stream() != nullptr && write_value32_fn() != nullptr;
}

int ds_env_entropy() {   // MRCLIP_DS=entropy: d logit_scale from the rescale pass's entropy sums on every shape
  const char* e = getenv("MRCLIP_DS");
  return (e && strcmp(e, "entropy") == 0) ? 1 : 0;
}
}  // namespace

extern "C" {

int mrclip_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  g_prof_on.store(on != 0);
  return 0;
}

int mrclip_prof_report(char* buf, size_t cap) {
  if (!buf || cap == 0) return fail(-1, "prof_report: no buffer");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<double, long>> acc;
  std::vector<std::string> order;
  for (auto& r : g_prof) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) return fail(-1, "prof_report: event not complete");
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) return fail(-1, "prof_report: elapsed time unavailable");
    if (!acc.count(r.name)) order.push_back(r.name);
    acc[r.name].first += ms;
    acc[r.name].second += 1;
  }
  std::string out;
  for (auto& k : order) {
    char line[160];
    snprintf(line, sizeof line, "%s:%.6f:%ld;", k.c_str(), acc[k].first, acc[k].second);
    out += line;
  }
  if (out.size() + 1 > cap) return fail(-1, "prof_report: buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

size_t mrclip_peer_block_bytes(void) { return kPeerBlockBytes; }
size_t mrclip_step_struct_bytes(void) { return sizeof(mrclip_step); }
size_t mrclip_peer_struct_bytes(void) { return sizeof(mrclip_peer); }
size_t mrclip_step_small_floats(void) { return kSmallFloats; }

int mrclip_step_uses_fwd_ds(const mrclip_step* s) {
  if (!s || s->kind != 0 || s->peer.ranks <= 1 || !s->local_loss || ds_env_entropy()) return 0;
  const int n = s->shape.m_rows, N = s->shape.n_cols;
  if ((long)n * N < (1L << 22) || s->shape.d % 4 != 0) return 0;   // below: the bf16 noise of G in <dT_r, T_r> is not averaged out
  return mrclip_fwd_row_ent_ok(n, N, n);
}

int mrclip_step_forward(const mrclip_step* s, const void* img, int img_dtype, long img_ld, const void* txt, int txt_dtype,
                        long txt_ld, const float* scale, const float* bias, int need_grad, int raw, float* loss_out,
                        void* stream) {
  if (int e = check_step(s, false)) return e;
  if (!img || !txt || !scale || !loss_out) return fail(-1, "step_forward: NULL argument");
  if (need_grad && !s->emat) return fail(-1, "step_forward: need_grad without an E block");
  cudaStream_t st = (cudaStream_t)stream;
  const mrclip_shape& sh = s->shape;
  const int n = sh.m_rows, N = sh.n_cols;
  const PeerInfo pi = make_peer(s);
  const int ranks = pi.ranks, rank = pi.rank;
  {
    ProfScope ps("pack_push", st);
    if (int e = step_pack(s, pi, img, img_dtype, img_ld, txt, txt_dtype, txt_ld, scale, raw, st)) return e;
  }
  if (raw) scale = s->scale_buf;     // exp(log-scale), written by the pack pre-pass
  FwdSig sig;
  if (ranks > 1) {
    sig.ready = pi.sig_local + CH_TEXT * kPeerMaxRanks;
    sig.epoch = &pi.ctl->epoch[CH_TEXT];
    sig.src_cols = n;
    sig.my_src = rank;
  }
  const FwdPlan f = fwd_plan(n, N);
  const WsLayout w = ws_layout(n, N, sh.d);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(s->ws);
  if (s->kind == 1) {   // SigLipLoss: softplus sum (+ G block), no statistics
    {
      ProfScope ps("fwd_tiles", st);
      if (int e = run_fwd(LOSS_SIGLIP, s->img_rows, s->txt_all, sh, s->ld, scale, bias, 0, N, s->ws,
                          need_grad ? s->emat : nullptr, st, false, &sig))
        return e;
    }
    if (int e = push_rows_join(s, pi, st)) return e;
    ProfScope ps("fwd_reduce", st);
    scalar_reduce_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const float2*>(wsb + w.sc_part),
                                             (long)f.num_rb * f.total_chunks * kEpiWarps, 1.f / (float)n, 0.f, nullptr,
                                             nullptr, loss_out, nullptr, 0, 0);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  const bool fwd_ds = need_grad && mrclip_step_uses_fwd_ds(s);
  {
    ProfScope ps("fwd_tiles", st);
    if (int e = run_fwd(LOSS_CLIP, s->img_rows, s->txt_all, sh, s->ld, scale, nullptr, 0, N, s->ws,
                        need_grad ? s->emat : nullptr, st, fwd_ds, &sig))
      return e;
  }
  if (int e = push_rows_join(s, pi, st)) return e;
  ProfScope ps("fwd_reduce", st);
  const float2* row_part = reinterpret_cast<const float2*>(wsb + w.row_part);
  const long plane = (long)rank * 3 * N;
  reduce_rows_pub_kernel<<<ceil_div(n, 256), 256, 0, st>>>(row_part, f.total_chunks * 2, n, f.m_pad, s->peer.stats_peers,
                                                          s->stats, plane + 2L * N, pi);
  g_launches.fetch_add(1);
  if (fwd_ds) {
    const int slots_per_rank = 2 * (n / (f.tiles_per_chunk * kSBN));
    row_ent_pub_kernel<<<ceil_div(n, 256), 256, ranks * sizeof(float), st>>>(
        row_part, reinterpret_cast<const float*>(wsb + w.row_ent), f.total_chunks * 2, slots_per_rank, ranks, n, f.m_pad,
        s->stats + plane + 2L * N, s->small + kSmallAcc, s->peer.ctl + kCtlRowEntDone, s->small + kSmallR2Row, pi);
    g_launches.fetch_add(1);
  }
  {
    const int bands = ceil_div(n, 32);
    static bool attr_set = false;
    if (!attr_set && bands * sizeof(float) > 48 * 1024) {
      CUDA_TRY(cudaFuncSetAttribute(reduce_cols_pub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    if (bands * sizeof(float) > 200 * 1024) return fail(-1, "m_rows=%d too large for the column reduce", n);
    reduce_cols_pub_kernel<<<ceil_div(N, 64), 1024, bands * sizeof(float), st>>>(
        reinterpret_cast<const float*>(wsb + w.col_l), reinterpret_cast<const float*>(wsb + w.col_c), bands, N, f.n_pad,
        s->peer.stats_peers, s->stats, plane, plane + N, pi);
    g_launches.fetch_add(1);
  }
  merge_stats_kernel<<<ceil_div(f.n_pad, 256), 256, 0, st>>>(s->stats, ranks, n, N, f.n_pad, s->lse2_col_all, s->lse2_row_all, pi);
  const int publish = (ranks > 1 && !s->local_loss) ? 1 : 0;
  clip_loss_pub_kernel<<<1, 1024, 0, st>>>(s->lse2_row_all + (long)rank * n, s->lse2_col_all,
                                           reinterpret_cast<const float*>(wsb + w.diag2), n, sh.label_offset,
                                           s->small + kSmallLoss, loss_out, publish, pi);
  g_launches.fetch_add(2);
  if (publish) {
    scal_mean_kernel<<<1, 64, 0, st>>>(0, CH_LOSS, loss_out, pi);
    g_launches.fetch_add(1);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_normalize_bwd(const void* y, long y_ld, const float* inv_norm, int rows, int d, void* g, int g_dtype, long g_ld,
                         void* stream) {
  if (!y || !inv_norm || !g || rows <= 0 || d <= 0 || y_ld < d || g_ld < d) return fail(-1, "normalize_bwd: bad arguments");
  if (g_dtype < 0 || g_dtype > 2) return fail(-1, "bad dtype %d", g_dtype);
  long blocks = ((long)rows + 7) / 8;
  if (blocks > 148L * 16) blocks = 148L * 16;
  normalize_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(y), y_ld, inv_norm,
                                                                     rows, d, g, g_dtype, g_ld);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_step_backward(const mrclip_step* s, const float* scale, const float* grad_out, float coef, void* d_img,
                         int d_img_dtype, long d_img_ld, void* d_txt, int d_txt_dtype, long d_txt_ld, float* d_scale,
                         float* d_bias, void* stream) {
  if (int e = check_step(s, true)) return e;
  if (!scale || !d_img || !d_txt || !s->emat) return fail(-1, "step_backward: NULL argument");
  if (d_img_dtype < 0 || d_img_dtype > 2 || d_txt_dtype < 0 || d_txt_dtype > 2) return fail(-1, "step_backward: bad output dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const mrclip_shape& sh = s->shape;
  const int n = sh.m_rows, N = sh.n_cols, d = sh.d, ld = s->ld;
  const PeerInfo pi = make_peer(s);
  const int ranks = pi.ranks, rank = pi.rank;
  const WsLayout w = ws_layout(n, N, d);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(s->ws);
  const __nv_bfloat16* txt_rows = reinterpret_cast<const __nv_bfloat16*>(s->txt_all) + (size_t)rank * n * ld;
  int mode = -1, nbands = 1, band_rows = n;
  SideStream* ss = nullptr;
  if (s->kind == 0) {
    const float* lse_row = s->lse2_row_all + (long)rank * n;
    mode = ranks > 1 ? (mrclip_step_uses_fwd_ds(s) ? 2 : 3) : (((long)n * N >= (1L << 22)) ? 1 : 0);
    float* msums = (d_scale && (mode == 0 || mode == 3)) ? s->msums : nullptr;
    const int* flag = mrclip_emat_flag(sh, s->ws);
    {
      ProfScope ps("emat_guard", st);
      if (int e = mrclip_emat_check(sh, s->ws, lse_row, s->lse2_col_all, stream)) return e;
      if (int e = mrclip_clip_gwrite_if(s->img_rows, s->txt_all, sh, ld, lse_row, s->lse2_col_all, scale, 1.f, 1.f, s->ws,
                                        s->emat, flag, stream))
        return e;
    }
    nbands = pick_bands(n);
    ss = nbands > 1 ? side_stream() : nullptr;
    if (ss == nullptr) nbands = 1;
    if (nbands > 1) {
      band_rows = ceil_div(ceil_div(n, nbands), 256) * 256;
      nbands = ceil_div(n, band_rows);
    }
    if (nbands <= 1) {
      ProfScope ps("emat_transform", st);
      if (int e = run_emat_transform(sh, s->ws, s->emat, lse_row, s->lse2_col_all, reinterpret_cast<const float*>(wsb + w.diag2),
                                     1.f, 1.f, flag, msums, 64, n, ranks, 0, 0, true, st))
        return e;
    } else {
      // fork: the side stream rescales band after band, an event per band tells the GEMMs below when theirs is ready
      if (msums) CUDA_TRY(cudaMemsetAsync(msums, 0, sizeof(float) * 2 * ranks * 64, st));
      CUDA_TRY(cudaEventRecord(ss->fork, st));
      CUDA_TRY(cudaStreamWaitEvent(ss->st, ss->fork, 0));
      ProfScope ps("emat_transform", ss->st);
      for (int b = 0; b < nbands; ++b) {
        const int r0 = b * band_rows, r1 = r0 + band_rows < n ? r0 + band_rows : n;
        if (int e = run_emat_transform(sh, s->ws, s->emat, lse_row, s->lse2_col_all,
                                       reinterpret_cast<const float*>(wsb + w.diag2), 1.f, 1.f, flag, msums, 64, n, ranks,
                                       r0 / 32, ceil_div(r1 - r0, 32), false, ss->st))
          return e;
        CUDA_TRY(cudaEventRecord(ss->ev[b], ss->st));
      }
    }
    if (ranks > 1 && msums) {
      if (nbands > 1) CUDA_TRY(cudaStreamWaitEvent(st, ss->ev[nbands - 1], 0));
      msums_pub_kernel<<<1, 64, 0, st>>>(msums, 64, s->small + kSmallR2Row, pi);
      g_launches.fetch_add(1);
    }
  } else {
    if (int e = mrclip_siglip_e_scalars(sh, s->ws, coef, grad_out, d_scale, d_bias, 0, stream)) return e;
  }
  // dI = G . T_all, band by band (each band waits for its rescale on the side stream)
  auto run_di = [&](const DotArgs& xf0) -> int {
    ProfScope ps("gemm_dI", st);
    const size_t esz = d_img_dtype == MRCLIP_DT_F32 ? 4 : 2;
    const size_t g_ld = (size_t)mrclip_padded_cols(N);
    for (int b = 0; b < nbands; ++b) {
      const int r0 = b * band_rows, r1 = r0 + band_rows < n ? r0 + band_rows : n;
      if (nbands > 1) CUDA_TRY(cudaStreamWaitEvent(st, ss->ev[b], 0));
      DotArgs xf = xf0;
      if (xf.dot_feat) xf.dot_feat = reinterpret_cast<const uint8_t*>(xf0.dot_feat) + (size_t)r0 * ld * 2;
      if (int e = run_gmat_gemm(false, reinterpret_cast<uint8_t*>(s->emat) + (size_t)r0 * g_ld * 2, r1 - r0, N, s->txt_all, d, ld,
                                coef, scale, grad_out, s->ws, reinterpret_cast<uint8_t*>(d_img) + (size_t)r0 * d_img_ld * esz,
                                d_img_dtype, d_img_ld, xf, st))
        return e;
    }
    return 0;
  };
  if (ranks > 1) {
    DotArgs xf;
    xf.peer = s->peer.recv_peers;
    xf.peer_n = n;
    xf.peer_rank = rank;
    xf.sig = pi;
    auto run_dt_push = [&]() -> int {
      ProfScope ps("gemm_dT_push", st);
      return run_gmat_gemm(true, s->emat, n, N, s->img_rows, d, ld, coef, scale, grad_out, s->ws,
                           const_cast<unsigned long long*>(s->peer.recv_peers),
                           s->peer.recv_bf16 ? MRCLIP_DT_BF16 : MRCLIP_DT_F32, d, xf, st);
    };
    if (nbands > 1) {   // large row blocks: hide the rescale pass behind the dI bands, then push
      if (int e = run_di(DotArgs())) return e;
      if (int e = run_dt_push()) return e;
    } else {            // small row blocks: push first, so that the NVLink transfers drain under the dI GEMM
      if (int e = run_dt_push()) return e;
      if (int e = run_di(DotArgs())) return e;
    }
    ProfScope ps("sum_slots", st);
    const bool dot = (s->kind == 0 && d_scale && mode == 2);
    const long total = ((d & 3) == 0) ? (long)n * (d / 4) : (long)n * d;
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    if (s->peer.recv_bf16)
      sum_slots_wait_kernel<true><<<(int)blocks, 256, 0, st>>>(s->peer.recv, ranks, n, d, d_txt, d_txt_dtype, d_txt_ld,
                                                               dot ? txt_rows : nullptr, ld, s->small + kSmallDot, pi);
    else
      sum_slots_wait_kernel<false><<<(int)blocks, 256, 0, st>>>(s->peer.recv, ranks, n, d, d_txt, d_txt_dtype, d_txt_ld,
                                                                dot ? txt_rows : nullptr, ld, s->small + kSmallDot, pi);
    g_launches.fetch_add(1);
  } else {
    DotArgs xf;
    if (s->kind == 0 && d_scale && mode == 1) {
      CUDA_TRY(cudaMemsetAsync(d_scale, 0, sizeof(float), st));
      xf.dot_feat = s->img_rows;
      xf.dot_out = d_scale;
    }
    if (int e = run_di(xf)) return e;
    ProfScope ps("gemm_dT", st);
    if (int e = run_gmat_gemm(true, s->emat, n, N, s->img_rows, d, ld, coef, scale, grad_out, s->ws, d_txt, d_txt_dtype,
                              d_txt_ld, DotArgs(), st))
      return e;
  }
  if (s->kind == 0 && d_scale) {
    ProfScope ps("ds_finish", st);
    DsParams dp;
    memset(&dp, 0, sizeof dp);
    dp.mode = mode;
    dp.gout = grad_out;
    dp.scale = scale;
    dp.loss_local = s->small + kSmallLoss;
    dp.kfac = 0.6931471805599453f * 0.5f / (float)n;
    dp.msums = s->msums;
    dp.count = 64 * 2 * ranks;
    dp.dot_slots = s->small + kSmallDot;
    dp.r2_row_tot = s->small + kSmallR2Row;
    dp.ds_out = d_scale;
    dp.publish = (ranks > 1 && !s->local_loss) ? 1 : 0;
    if (mode != 1 || dp.publish) {
      ds_finish_kernel<<<1, 128, 0, st>>>(dp, pi);
      g_launches.fetch_add(1);
    }
    if (dp.publish) {
      scal_mean_kernel<<<1, 64, 0, st>>>(1, CH_DSCALE, d_scale, pi);
      g_launches.fetch_add(1);
    }
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"

/* ---- retrieval metrics: rank-of-label epilogue on the S tiles (tile_kernel<MODE_RANK>) ---------------------------- */
namespace {
int run_rank(int phase, const void* a_rows, const void* b_all, const mrclip_shape& sh, int ld, const int* row_cls,
             const int* col_cls, const int* col_ord, const long long* row_off, const int* row_m, float* pos,
             const float* lmax, int chunk0, unsigned long long* pairs, int* best, cudaStream_t st) {
  if (int e = check_shape(sh, ld)) return e;
  if (!a_rows || !b_all || !row_cls || !col_cls || !row_off || !pos) return fail(-1, "rank: NULL argument");
  const FwdPlan f = fwd_plan(sh.m_rows, sh.n_cols);
  CUtensorMap ma, mb;
  if (int e = make_map(&ma, a_rows, sh.m_rows, ld, ld, kBM)) return e;
  if (int e = make_map(&mb, b_all, sh.n_cols, ld, ld, kSBN)) return e;
  TileParams p;
  memset(&p, 0, sizeof p);
  p.m_rows = sh.m_rows;
  p.n_cols = sh.n_cols;
  p.num_kb = ceil_div(ld, kBK);
  p.num_rb = f.num_rb;
  p.tile_begin = 0;
  p.tile_end = f.num_ct;
  p.tiles_per_chunk = f.tiles_per_chunk;
  p.num_chunks = f.total_chunks;
  p.num_dc = 1;
  p.num_items = f.total_chunks * f.num_rb;
  p.m_pad = f.m_pad;
  p.n_pad = f.n_pad;
  p.rk_row_cls = row_cls;
  p.rk_col_cls = col_cls;
  p.rk_col_ord = col_ord;
  p.rk_off = row_off;
  p.rk_m = row_m;
  p.rk_pos = pos;
  p.rk_lmax = lmax;
  p.rk_pairs = pairs;
  p.rk_best = best;
  p.rk_phase = phase;
  p.rk_chunk0 = chunk0;
  return launch_tile<MODE_RANK, LOSS_CLIP, 256, kSBN>(ma, mb, mb, p, st);
}
}  // namespace

extern "C" {

int mrclip_rank_collect(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const int* row_cls,
                        const int* col_cls, const int* col_ord, const long long* row_off, float* pos, void* stream) {
  if (!col_ord) return fail(-1, "rank_collect: NULL col_ord");
  return run_rank(0, a_rows, b_all, shape, ld, row_cls, col_cls, col_ord, row_off, nullptr, pos, nullptr, 0, nullptr, nullptr,
                  (cudaStream_t)stream);
}

int mrclip_rank_lmax(const float* pos, const long long* row_off, const int* row_m, int rows, float* lmax, void* stream) {
  if (!pos || !row_off || !row_m || !lmax || rows <= 0) return fail(-1, "rank_lmax: bad arguments");
  long blocks = ((long)rows + 7) / 8;
  if (blocks > 148L * 16) blocks = 148L * 16;
  rank_lmax_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pos, row_off, row_m, rows, lmax);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_rank_count(const void* a_rows, const void* b_all, mrclip_shape shape, int ld, const int* row_cls,
                      const int* col_cls, const long long* row_off, const int* row_m, const float* pos, const float* lmax,
                      int chunk0, unsigned long long* pairs, int* best, void* stream) {
  if (!row_m || !lmax || !pairs || !best || chunk0 < 0) return fail(-1, "rank_count: bad arguments");
  return run_rank(1, a_rows, b_all, shape, ld, row_cls, col_cls, nullptr, row_off, row_m, const_cast<float*>(pos), lmax, chunk0,
                  pairs, best, (cudaStream_t)stream);
}

}  // extern "C"

/* ---- MultiPositiveClipLoss: class means of the packed features and the terms built on them ----------------------- */
extern "C" {

int mrclip_class_means(const void* x, int ld, const int* order, const int* seg_start, const int* seg_cnt, int n_ids,
                       float* mean, void* stream) {
  if (!x || !order || !seg_start || !seg_cnt || !mean || n_ids <= 0 || ld <= 0 || ld % 8 != 0)
    return fail(-1, "class_means: bad arguments");
  class_mean_kernel<<<n_ids, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, order, seg_start,
                                                            seg_cnt, mean);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_mpos_forward(const void* img_rows, const void* txt_rows, int ld, int n, int d, const int* cls, const float* tmean,
                        const float* imean, const float* lse2_row, const float* lse2_col, const float* scale, float delta,
                        float* loss, void* stream) {
  if (!img_rows || !txt_rows || !cls || !tmean || !imean || !lse2_row || !lse2_col || !scale || !loss || n <= 0 || d <= 0 || ld < d)
    return fail(-1, "mpos_forward: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), st));
  long blocks = ((long)n + 7) / 8;
  if (blocks > 148L * 8) blocks = 148L * 8;
  mpos_forward_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(img_rows),
                                                   reinterpret_cast<const __nv_bfloat16*>(txt_rows), ld, n, d, cls, tmean, imean,
                                                   lse2_row, lse2_col, scale, delta, loss);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int mrclip_mpos_backward(void* d_img, int d_img_dtype, long d_img_ld, void* d_txt, int d_txt_dtype, long d_txt_ld,
                         const void* img_rows, const void* txt_rows, int ld, int n, int d, const int* cls, const float* tmean,
                         const float* imean, float coef, const float* scale, const float* grad_out, void* stream) {
  if (!d_img || !d_txt || !img_rows || !txt_rows || !cls || !tmean || !imean || !scale || n <= 0 || d <= 0 || ld < d)
    return fail(-1, "mpos_backward: bad arguments");
  if (d_img_dtype < 0 || d_img_dtype > 2 || d_txt_dtype < 0 || d_txt_dtype > 2) return fail(-1, "mpos_backward: bad dtype");
  long blocks = ((long)n * d + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  mpos_backward_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      d_img, d_img_dtype, d_img_ld, d_txt, d_txt_dtype, d_txt_ld, reinterpret_cast<const __nv_bfloat16*>(img_rows),
      reinterpret_cast<const __nv_bfloat16*>(txt_rows), ld, n, d, cls, tmean, imean, coef, scale, grad_out);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"
