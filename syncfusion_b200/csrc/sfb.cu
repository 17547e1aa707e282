// libsyncfusion_b200.so - C ABI (include/syncfusion_b200.h), weight re-packing, execution plan and launch logic
// for the SyncFusion U-Net v-diffusion sampling loop on sm_100a.  See DESIGN.md for the data layout and the plan.
//
// Reference path being replaced (all in /root/reference): main/generation.py:77-83 and
// main/module_diffusion.py:200-206 call DiffusionModel.sample(); the architecture is exp/model/diffusion.yaml:11-33;
// the arithmetic (audio_diffusion_pytorch / a_unet) is restated in oracle/ (SURVEY.md Appendix A).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <memory>
#include <string>
#include <vector>

#include "../../include/syncfusion_b200.h"
#include "attn_tc.cuh"
#include "d0.cuh"
#include "elementwise.cuh"
#include "encoder.cuh"
#include "gemm_tc.cuh"
#include "postprocess.cuh"
#include "prepare.cuh"
#include "rk_tc.cuh"
#include "sk_tc.cuh"

using namespace sfb;

namespace {

// ------------------------------------------------------------------------------------------------ utilities
struct HostTensor {
  std::vector<float> v;
  std::vector<int64_t> shape;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <typename T> CUtensorMapDataType tmap_dtype();
template <> CUtensorMapDataType tmap_dtype<__nv_bfloat16>() { return CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }
template <> CUtensorMapDataType tmap_dtype<float>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT32; }

// rank-3 map over a contiguous [d2][d1][d0] tensor (d0 innermost), 128B swizzle, zero OOB fill.
template <typename T>
bool make_tmap3(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * sizeof(T), d0 * d1 * sizeof(T)};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return fn(m, tmap_dtype<T>(), 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <typename T>
bool make_tmap2(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint32_t b0, uint32_t b1) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {d0 * sizeof(T)};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t es[2] = {1, 1};
  return fn(m, tmap_dtype<T>(), 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T> T host_cvt(float v);
template <> float host_cvt<float>(float v) {   // round-to-nearest tf32 (10-bit mantissa), matching cvt.rna.tf32.f32
  uint32_t u;
  memcpy(&u, &v, 4);
  if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}
template <> __nv_bfloat16 host_cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Row-major [taps * N][Kw] fp32 -> bf16 tiles [taps][ceil(Kw / 64)][N][64] (K zero padded): one resident UMMA B tile
// (128-byte rows, loaded with a 128B-swizzle TMA box [64, N]) per (tap, 64-wide K atom).  Appends to `out`.
void pack_tiles(std::vector<__nv_bfloat16>& out, const float* Wm, int taps, int N, int Kw) {
  const int KA = (Kw + 63) / 64;
  for (int t = 0; t < taps; ++t)
    for (int a = 0; a < KA; ++a)
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < 64; ++k) {
          const int kk = a * 64 + k;
          out.push_back(__float2bfloat16_rn(kk < Kw ? Wm[((size_t)t * N + n) * Kw + kk] : 0.f));
        }
}

struct GemmW {           // re-packed weights of one contraction
  void* w = nullptr;     // [taps * N][Kw] operand precision
  float* bias = nullptr; // [bias_mod] or null
  int N = 0, K1 = 0, K2 = 0, taps = 1, bias_mod = 0;
};

struct ItemW {
  void *rk1_w = nullptr, *rk2_w = nullptr;   // tile-packed resident weights (rk_tc.cuh): conv1 | conv2 + inject
  int rk1_id = -1, rk2_id = -1;
  float *gn1_g = nullptr, *gn1_b = nullptr, *gn2_g = nullptr, *gn2_b = nullptr;
  GemmW conv1, conv2, inject, qkv, out;
  float *c8_w1 = nullptr, *c8_b1 = nullptr, *c8_w2 = nullptr, *c8_b2 = nullptr, *c8_wi = nullptr, *c8_bi = nullptr;
  bool has_attn = false, has_xattn = false, has_inject = false;
  float *x_ng = nullptr, *x_nb = nullptr, *x_wv = nullptr, *x_wo = nullptr;
  // general M_ctx > 1 cross-attention (K6): q projection with the query LayerNorm affine folded in, full to_kv, to_out
  GemmW xq, xout;
  float* x_wkv = nullptr;
  int xkv_index = -1;
  int mod_off = 0, xb_off = 0;
  float* qkv_ws = nullptr;   // [1536] column sums of the bf16 fused QKV weights (LayerNorm fold, sk_tc.cuh)
};
struct DepthW {
  GemmW down, up;
  void *rk_down_w = nullptr, *rk_up_w = nullptr;
  int rk_down_id = -1, rk_up_id = -1;
  float *c8_dw = nullptr, *c8_db = nullptr, *c8_uw = nullptr;
  float c8_ub = 0.f;
  int up_taps = 1;
  std::vector<ItemW> items[2];
  int skip_off = 0;
};

enum OpKind { OP_D0_DOWN = 0, OP_GN, OP_CONV_C8, OP_INJ_C8, OP_D0_UP, OP_GEMM, OP_LN, OP_ATTN, OP_RK, OP_SK, OP_D0_CONV1, OP_D0_TAIL, OP_XOUT_C8 };
const char* kOpNames[] = {"d0_down", "gn_silu", "conv3_c8", "inject_c8", "d0_up", "gemm", "ln", "attn", "rk", "sk", "d0_conv1", "d0_tail", "xattn_out_c8"};

// ------------------------------------------------------------------------------------------------ wait log (ptx.cuh)
// One host-mapped log per process: the device writes it when a barrier wait outlasts c_wait_bound (then traps); the
// host can still read it after the context is lost.
WaitLog* g_wait_log_host = nullptr;
int wait_log_init() {
  if (g_wait_log_host) return 0;
  void* h = nullptr;
  if (cudaHostAlloc(&h, sizeof(WaitLog), cudaHostAllocMapped) != cudaSuccess) return -1;
  memset(h, 0, sizeof(WaitLog));
  reinterpret_cast<WaitLog*>(h)->cur_tag = 0xFFFFFFFFu;
  void* d = nullptr;
  if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) return -1;
  WaitLog* dp = reinterpret_cast<WaitLog*>(d);
  if (cudaMemcpyToSymbol(g_wait_log, &dp, sizeof dp) != cudaSuccess) return -1;
  g_wait_log_host = reinterpret_cast<WaitLog*>(h);
  return 0;
}
const char* wait_site_file(uint32_t id) {
  switch (id) { case 9: return "sfb.cu"; case 1: return "gemm_tc.cuh"; case 2: return "attn_tc.cuh"; case 3: return "rk_tc.cuh"; case 4: return "sk_tc.cuh"; case 5: return "d0.cuh"; }
  return "?";
}

struct EngineBase {
  sfb_unet_config cfg;
  int device = 0;
  std::string err;
  std::map<std::string, HostTensor> params;
  bool finalized = false;
  int op_limit = -1;
  int64_t launches = 0;
  virtual ~EngineBase() {}
  virtual int finalize() = 0;
  virtual int workspace_bytes(int64_t B, int64_t L, int cfg_on, int64_t rows, int64_t M, size_t* out) = 0;
  virtual int unet_forward(const float* x, const float* sigma, const float* const* channels, int n_channels,
                           const float* embedding, int64_t M, float scale, float* v_out, int64_t B, int64_t L, void* ws,
                           size_t ws_bytes, cudaStream_t st) = 0;
  virtual int sample(const float* x_noisy, int num_steps, const float* const* channels, int n_channels,
                     const float* embedding, int64_t M, float scale, float* x_out, float* traj_x, float* traj_v,
                     const float* teacher_x, int64_t B, int64_t L, void* ws, size_t ws_bytes, cudaStream_t st) = 0;
  virtual int plan_size(int64_t B, int64_t L, int cfg_on, void* ws, size_t ws_bytes, int64_t M = 1) = 0;
  virtual int op_info(int i, char* buf, int len) = 0;
  virtual int profile_report(char* buf, int len) = 0;
  virtual int sk_timeline(int op_index, long long* host_buf, int n) { (void)op_index; (void)host_buf; (void)n; return SFB_ERR_UNSUPPORTED; }
  bool profiling = false;
  virtual std::string describe_wait(const WaitRecord& r) { (void)r; return ""; }
  // Text of the device wait log ("" if no wait timed out): one line per stuck waiter + the raw barrier words.
  std::string wait_log_text() {
    const WaitLog* lg = g_wait_log_host;
    if (!lg || (lg->count == 0 && lg->light_flag == 0)) return "";
    char b[640];
    std::string out;
    WaitRecord cur{};
    cur.tag = lg->cur_tag;
    if (lg->light_flag) {      // SFB_WAIT_LOG=1 (default build): last stuck waiter + the progress marker
      snprintf(b, sizeof b, "barrier wait timed out (device trap): %s:%u cta(%u,%u,%u) thread %u/%u (warp %u); last tensor-core op started: op#%u %s\n"
               "  (build with SFB_WAIT_LOG=2 for one record per stuck waiter incl. barrier slot and state)\n",
               wait_site_file(lg->light_site >> 16), lg->light_site & 0xFFFFu, lg->light_cta & 0xFFFFu, (lg->light_cta >> 16) & 0xFFu,
               lg->light_cta >> 24, lg->light_thread & 0xFFFFu, lg->light_thread >> 16, (lg->light_thread & 0xFFFFu) >> 5, lg->cur_tag,
               describe_wait(cur).c_str());
      out += b;
    }
    if (lg->count) {
      const uint32_t n = std::min<uint32_t>(lg->count, kWaitRecMax);
      snprintf(b, sizeof b, "barrier wait timed out: %u waiter(s) reported (first %u shown); last tensor-core op started: op#%u\n", lg->count, n, lg->cur_tag);
      out += b;
      for (uint32_t i = 0; i < n; ++i) {
        const WaitRecord& r = lg->rec[i];
        snprintf(b, sizeof b, "  [%u] %s:%u op#%u cta(%u,%u,%u) thread %u/%u (warp %u) bar@0x%x parity %u state 0x%08x%08x %s\n", i,
                 wait_site_file(r.site >> 16), r.site & 0xFFFFu, r.tag, r.cta_x, r.cta_yz & 0xFFFFu, r.cta_yz >> 16, r.thread & 0xFFFFu,
                 r.thread >> 16, (r.thread & 0xFFFFu) >> 5, r.bar & 0x7FFFFFFFu, r.bar >> 31, r.state_hi, r.state_lo, describe_wait(r).c_str());
        out += b;
      }
    }
    return out;
  }
  int fail(int code, const char* fmt, ...) {
    char b[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(b, sizeof b, fmt, ap);
    va_end(ap);
    err = b;
    if (code == SFB_ERR_CUDA) { const std::string w = wait_log_text(); if (!w.empty()) err += "\n" + w; }
    return code;
  }
};

#define SFB_CUDA(call)                                                                                 \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail(SFB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------------ GEMM launch
int g_grid_limit = 0;     // test hook (sfb_dbg_set_grid_limit): persistent kernels use at most this many CTAs (0: all SMs)
int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return (g_grid_limit > 0 && g_grid_limit < n) ? g_grid_limit : n;
}
template <typename T, int BN>
void launch_gemm_bn(GemmParams<T> p, int B, cudaStream_t st) {
  p.n_tiles = p.N / BN;
  p.total_tiles = B * p.tiles_per_clip * p.n_tiles;
  const int grid = std::min(p.total_tiles, num_sms() * GemmCfg<BN>::kCtasPerSm);   // persistent CTAs
  launch_pdl(gemm_tc_kernel<T, BN>, grid, kGemmThreads, gemm_smem_bytes<T, BN>(), st, p);
}
template <typename T>
void launch_gemm(const GemmParams<T>& p, int BN, int B, cudaStream_t st) {
  switch (BN) {
    case 32: launch_gemm_bn<T, 32>(p, B, st); break;
    case 64: launch_gemm_bn<T, 64>(p, B, st); break;
    case 128: launch_gemm_bn<T, 128>(p, B, st); break;
    default: launch_gemm_bn<T, 256>(p, B, st); break;
  }
}
template <typename T>
cudaError_t set_kernel_attrs() {
  cudaError_t e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<T, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<T, 32>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<T, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<T, 64>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<T, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<T, 128>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<T, 256>());
  if (e != cudaSuccess) return e;
  e = attn_set_attrs<T>();
  return e;
}
int pick_bn(int N) {
  if (N % 256 == 0 && N >= 1024) return 256;
  if (N % 128 == 0) return 128;
  if (N == 64) return 64;
  if (N == 32) return 32;
  return 0;
}

template <typename T>
bool fill_gemm_maps(GemmParams<T>& p, const void* a1, int K1, int L, int B, const void* a2, int K2, int B2, const void* w,
                    int N, int taps, int BN) {
  constexpr int BK = ElemTraits<T>::kAtomElems;
  if (!make_tmap3<T>(&p.tmA1, a1, K1, L, B, BK, kGemmBM)) return false;
  if (K2 > 0) {
    if (!make_tmap3<T>(&p.tmA2, a2, K2, L, B2, BK, kGemmBM)) return false;
  } else {
    p.tmA2 = p.tmA1;
  }
  if (!make_tmap2<T>(&p.tmW, w, (uint64_t)(K1 + K2), (uint64_t)taps * N, BK, BN)) return false;
  p.rows_per_clip = L;
  p.tiles_per_clip = (L + kGemmBM - 1) / kGemmBM;
  p.N = N;
  p.taps = taps;
  p.k1_chunks = (K1 + BK - 1) / BK;
  p.k2_chunks = K2 > 0 ? (K2 + BK - 1) / BK : 0;
  p.K1 = K1;
  p.a2_bmod = B2 > 0 ? B2 : 1;
  return true;
}

// ------------------------------------------------------------------------------------------------ engine
template <typename T>
struct Engine : EngineBase {
  std::vector<void*> owned;
  std::vector<DepthW> dw;
  // conditioning weights (fp32)
  float *t_w = nullptr, *t_lw = nullptr, *t_lb = nullptr, *t_mw = nullptr, *t_mb = nullptr, *fixed_emb = nullptr;
  float *ft_w = nullptr, *ft_b = nullptr;   // concatenated Linear(SiLU(features)) weights [F_total][MF]
  int F_total = 0, XB_total = 0, n_gn = 0, n_xattn = 0;
  bool no_rk = getenv("SFB_NO_RK") != nullptr;   // debugging aids: force the unfused generic path
  bool no_sk = getenv("SFB_NO_SK") != nullptr;
  bool no_d0_fused = getenv("SFB_NO_D0_FUSED") != nullptr;
  bool d0_tc = !(getenv("SFB_D0_TC") != nullptr && atoi(getenv("SFB_D0_TC")) == 0);   // depth-0 convs on tcgen05 (d0.cuh, bf16 mode); SFB_D0_TC=0: CUDA-core form

  struct Op {
    int kind = 0, depth = 0, stack = 0, item = 0;
    const void* in = nullptr;
    const void* in2 = nullptr;
    void* out_t = nullptr;
    float* out_r = nullptr;
    const float* resid = nullptr;
    double* stats_in = nullptr;
    double* stats_out = nullptr;
    const float *w0 = nullptr, *w1 = nullptr, *w2 = nullptr;
    const float* wx[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // extra vectors of the fused depth-0 ops
    float fscalar = 0.f;
    int L = 0, C = 0, gs = 0, B = 0, taps = 1, in_is_f32 = 0, ctx = 0, k2 = 0;
    int ft_off = -1;      // feature-table column offset (Modulation scale | SkipModulate scale), -1: none
    GemmParams<T> gp;
    AttnParams<T> ap;
    RkParams rp;
    int rk_id = -1;
    SkParams sp;
    int sk_id = -1, sk_ft_is_mod = 0;
    double flops = 0, bytes = 0;   // algorithmic work of an OP_RK op (set at build time)
    const char* ck = "";           // oracle-trace checkpoint this op's output equals (tests/trace.py)
    int BN = 0;
    size_t dbg_off = 0, dbg_bytes = 0;
    int dbg_rows = 0, dbg_cols = 0, dbg_dtype = 0;
  };
  struct WsLayout {
    size_t sigma, embrows, tmp1, tmp2, fourier, h1, h2, feat, ftable, xbias, stats, stats_bytes, veff, xstate, total;
    size_t ctx[SFB_MAX_DEPTH], bufA[SFB_MAX_DEPTH], T1[SFB_MAX_DEPTH], T2[SFB_MAX_DEPTH], qkv[SFB_MAX_DEPTH], o[SFB_MAX_DEPTH];
    size_t rs1[SFB_MAX_DEPTH], rs2[SFB_MAX_DEPTH];   // per-position LayerNorm partial sums [rows, parts <= 8, 2] (sk path)
    size_t foldw, foldw_bytes;                        // scaled inject weights + (ws | wsh) vectors, B copies (LayerNorm fold)
    size_t xln, xq, xo, xkv;                          // general cross-attention (M_ctx > 1): LN(x), q, attention output, per-item k | v
  };
  struct Plan {
    int64_t B = 0, L = 0;
    int cfg_on = 0;
    int64_t M = 1;            // context tokens of the cross-attention (1: collapsed to a bias)
    void* ws = nullptr;
    WsLayout lay;
    std::vector<Op> ops;
  };
  Plan plan;

  ~Engine() override {
    for (GraphEntry& g : graphs) cudaGraphExecDestroy(g.exec);
    if (gs) cudaStreamDestroy(gs);
    if (ge0) cudaEventDestroy(ge0);
    if (ge1) cudaEventDestroy(ge1);
    for (void* p : owned) cudaFree(p);
    for (cudaEvent_t e : prof_ev) cudaEventDestroy(e);
  }

  // ---------------------------------------------------------------- parameter access
  const HostTensor* get(const std::string& name, std::initializer_list<int64_t> shape) {
    auto it = params.find(name);
    if (it == params.end()) {
      fail(SFB_ERR_MISSING, "missing parameter '%s'", name.c_str());
      return nullptr;
    }
    std::vector<int64_t> s(shape);
    if (it->second.shape != s) {
      std::string got;
      for (auto d : it->second.shape) got += std::to_string(d) + ",";
      std::string want;
      for (auto d : s) want += std::to_string(d) + ",";
      fail(SFB_ERR_INVALID, "parameter '%s' has shape [%s], expected [%s]", name.c_str(), got.c_str(), want.c_str());
      return nullptr;
    }
    return &it->second;
  }
  template <typename U>
  U* upload(const std::vector<U>& h) {
    void* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(h.size() * sizeof(U), 16)) != cudaSuccess) return nullptr;
    owned.push_back(d);
    if (cudaMemcpy(d, h.data(), h.size() * sizeof(U), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    return reinterpret_cast<U*>(d);
  }
  float* upload_f(const HostTensor* t) { return t ? upload<float>(t->v) : nullptr; }
  void* upload_T(const std::vector<float>& w) {
    std::vector<T> h(w.size());
    for (size_t i = 0; i < w.size(); ++i) h[i] = host_cvt<T>(w[i]);
    return upload<T>(h);
  }

  // conv weight [Co][Ci][taps] -> [taps*Co][Ci]
  bool pack_conv3(const std::string& base, int C, GemmW& g, std::vector<float>* keep = nullptr) {
    const HostTensor* w = get(base + ".weight", {C, C, 3});
    const HostTensor* b = get(base + ".bias", {C});
    if (!w || !b) return false;
    std::vector<float> r((size_t)3 * C * C);
    for (int co = 0; co < C; ++co)
      for (int ci = 0; ci < C; ++ci)
        for (int t = 0; t < 3; ++t) r[((size_t)t * C + co) * C + ci] = w->v[((size_t)co * C + ci) * 3 + t];
    g.w = upload_T(r);
    g.bias = upload_f(b);
    g.N = C; g.K1 = C; g.K2 = 0; g.taps = 3; g.bias_mod = C;
    if (keep) keep->swap(r);
    return g.w && g.bias;
  }
  static constexpr bool kBF16 = sizeof(T) == 2;   // the resident-weight fused kernels (rk_tc.cuh) exist for bf16 operands

  int finalize() override {
    const sfb_unet_config& c = cfg;
    const int D = c.depth, MF = c.modulation_features, EF = c.embedding_features;
    const int mid = c.attention_heads * c.attention_features;
    if (c.in_channels != 1) return fail(SFB_ERR_UNSUPPORTED, "in_channels must be 1");
    if (c.channels[0] != 8) return fail(SFB_ERR_UNSUPPORTED, "channels[0] must be 8 (depth-0 streaming kernels)");
    if (c.factors[0] != 1) return fail(SFB_ERR_UNSUPPORTED, "factors[0] must be 1");
    if (c.attention_heads != 8 || c.attention_features != 64)
      return fail(SFB_ERR_UNSUPPORTED, "attention must be 8 heads x 64");
    if (c.resnet_groups != 8) return fail(SFB_ERR_UNSUPPORTED, "resnet_groups must be 8");
    if (c.context_channels[0] != 2) return fail(SFB_ERR_UNSUPPORTED, "context_channels[0] must be 2");
    for (int d = 1; d < D; ++d) {
      if (!pick_bn(c.channels[d])) return fail(SFB_ERR_UNSUPPORTED, "channels[%d]=%d must be 32, 64 or a multiple of 128", d, c.channels[d]);
      if (c.context_channels[d] <= 0 || c.context_channels[d] % 8) return fail(SFB_ERR_UNSUPPORTED, "context_channels[%d] must be a positive multiple of 8", d);
      if (!pick_bn(c.factors[d] * c.channels[d - 1])) return fail(SFB_ERR_UNSUPPORTED, "factors[%d]*channels[%d] must be 32, 64 or a multiple of 128", d, d - 1);
    }
    if (c.attentions[0]) return fail(SFB_ERR_UNSUPPORTED, "self-attention at depth 0 unsupported");
    if (set_kernel_attrs<T>() != cudaSuccess) return fail(SFB_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (kBF16 && rk_set_attrs() != cudaSuccess) return fail(SFB_ERR_CUDA, "cudaFuncSetAttribute (rk) failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (kBF16 && sk_set_attrs() != cudaSuccess) return fail(SFB_ERR_CUDA, "cudaFuncSetAttribute (sk) failed: %s", cudaGetErrorString(cudaGetLastError()));

    // time conditioning (A.3)
    t_w = upload_f(get("time.weights", {128}));
    t_lw = upload_f(get("time.linear.weight", {MF, 257}));
    t_lb = upload_f(get("time.linear.bias", {MF}));
    t_mw = upload_f(get("time.mlp.weight", {MF, MF}));
    t_mb = upload_f(get("time.mlp.bias", {MF}));
    fixed_emb = upload_f(get("fixed_embedding.weight", {c.embedding_max_length, EF}));
    if (!t_w || !t_lw || !t_lb || !t_mw || !t_mb || !fixed_emb) return err.empty() ? fail(SFB_ERR_CUDA, "upload failed") : SFB_ERR_MISSING;

    std::vector<float> ftw, ftb;   // concatenated feature linears
    dw.assign(D, DepthW());
    F_total = 0; XB_total = 0; n_xattn = 0;
    for (int d = 0; d < D; ++d) {
      const int C = c.channels[d], Cin = d == 0 ? c.in_channels : c.channels[d - 1], f = c.factors[d], ctx = c.context_channels[d];
      DepthW& W = dw[d];
      char pre[64];
      snprintf(pre, sizeof pre, "d%d.", d);
      const std::string P(pre);
      // ---- Down (A.6)
      {
        const HostTensor* w = get(P + "down.weight", {C, Cin, f});
        const HostTensor* b = get(P + "down.bias", {C});
        if (!w || !b) return SFB_ERR_MISSING;
        if (d == 0) {
          W.c8_dw = upload_f(w); W.c8_db = upload_f(b);
        } else {
          std::vector<float> r((size_t)C * f * Cin);
          for (int co = 0; co < C; ++co)
            for (int ci = 0; ci < Cin; ++ci)
              for (int j = 0; j < f; ++j) r[(size_t)co * f * Cin + j * Cin + ci] = w->v[((size_t)co * Cin + ci) * f + j];
          W.down.w = upload_T(r); W.down.bias = upload_f(b);
          W.down.N = C; W.down.K1 = f * Cin; W.down.taps = 1; W.down.bias_mod = C;
          if (kBF16 && !no_rk && (W.rk_down_id = rk_find(f * Cin, C, 1, 0, 0, C, 1)) >= 0) {
            std::vector<__nv_bfloat16> tiles;
            pack_tiles(tiles, r.data(), 1, C, f * Cin);
            W.rk_down_w = upload<__nv_bfloat16>(tiles);
          }
        }
      }
      // ---- Up (A.6, both modes)
      if (c.upsample_mode == SFB_UPSAMPLE_TRANSPOSE) {
        const HostTensor* w = get(P + "up.weight", {C, Cin, f});
        const HostTensor* b = get(P + "up.bias", {Cin});
        if (!w || !b) return SFB_ERR_MISSING;
        W.up_taps = 1;
        if (d == 0) {
          std::vector<float> r(8);
          for (int ci = 0; ci < 8; ++ci) r[ci] = w->v[ci];
          W.c8_uw = upload<float>(r); W.c8_ub = b->v[0];
        } else {
          std::vector<float> r((size_t)f * Cin * C);
          for (int ci = 0; ci < C; ++ci)
            for (int co = 0; co < Cin; ++co)
              for (int j = 0; j < f; ++j) r[((size_t)j * Cin + co) * C + ci] = w->v[((size_t)ci * Cin + co) * f + j];
          W.up.w = upload_T(r); W.up.bias = upload_f(b);
          W.up.N = f * Cin; W.up.K1 = C; W.up.taps = 1; W.up.bias_mod = Cin;
          if (kBF16 && !no_rk && (W.rk_up_id = rk_find(C, f * Cin, 1, 0, 0, Cin, 1)) >= 0) {
            std::vector<__nv_bfloat16> tiles;
            pack_tiles(tiles, r.data(), 1, f * Cin, C);
            W.rk_up_w = upload<__nv_bfloat16>(tiles);
          }
        }
      } else {
        const HostTensor* w = get(P + "up.conv.weight", {Cin, C, 3});
        const HostTensor* b = get(P + "up.conv.bias", {Cin});
        if (!w || !b) return SFB_ERR_MISSING;
        W.up_taps = 3;
        if (d == 0) {
          std::vector<float> r(24);
          for (int t = 0; t < 3; ++t)
            for (int ci = 0; ci < 8; ++ci) r[t * 8 + ci] = w->v[(size_t)ci * 3 + t];
          W.c8_uw = upload<float>(r); W.c8_ub = b->v[0];
        } else {
          // nearest xf + conv3 at the fine resolution == conv3 at the coarse resolution with N = f*Cin outputs:
          // fine position l*f + j, tap t reads coarse position l + floor((j + t - 1) / f).
          const int N = f * Cin;
          std::vector<float> r((size_t)3 * N * C, 0.f);
          for (int j = 0; j < f; ++j)
            for (int t = 0; t < 3; ++t) {
              const int num = j + t - 1;
              const int o = num < 0 ? -1 : num / f;    // floor division
              for (int co = 0; co < Cin; ++co)
                for (int ci = 0; ci < C; ++ci)
                  r[((size_t)(o + 1) * N + j * Cin + co) * C + ci] += w->v[((size_t)co * C + ci) * 3 + t];
            }
          W.up.w = upload_T(r); W.up.bias = upload_f(b);
          W.up.N = N; W.up.K1 = C; W.up.taps = 3; W.up.bias_mod = Cin;
          if (kBF16 && !no_rk && (W.rk_up_id = rk_find(C, N, 3, 0, 0, Cin, 1)) >= 0) {
            std::vector<__nv_bfloat16> tiles;
            pack_tiles(tiles, r.data(), 3, N, C);
            W.rk_up_w = upload<__nv_bfloat16>(tiles);
          }
        }
      }
      // ---- SkipModulate Linear (A.5)
      {
        const HostTensor* w = get(P + "skip.weight", {Cin, MF});
        const HostTensor* b = get(P + "skip.bias", {Cin});
        if (!w || !b) return SFB_ERR_MISSING;
        W.skip_off = F_total;
        ftw.insert(ftw.end(), w->v.begin(), w->v.end());
        ftb.insert(ftb.end(), b->v.begin(), b->v.end());
        F_total += Cin;
        while (F_total % 4) {   // keep every table segment 16-byte aligned (vector loads in the GEMM epilogue)
          ftw.insert(ftw.end(), MF, 0.f);
          ftb.push_back(0.f);
          ++F_total;
        }
      }
      // ---- items
      for (int s = 0; s < 2; ++s) {
        W.items[s].assign(c.items[d], ItemW());
        for (int i = 0; i < c.items[d]; ++i) {
          ItemW& I = W.items[s][i];
          std::vector<float> keep1, keep2;   // host copies of the conv weights [3 * C][C] for the resident tile packs
          char ip[96];
          snprintf(ip, sizeof ip, "d%d.items_%s.%d.", d, s == 0 ? "down" : "up", i);
          const std::string Q(ip);
          I.gn1_g = upload_f(get(Q + "resnet.gn1.weight", {C}));
          I.gn1_b = upload_f(get(Q + "resnet.gn1.bias", {C}));
          I.gn2_g = upload_f(get(Q + "resnet.gn2.weight", {C}));
          I.gn2_b = upload_f(get(Q + "resnet.gn2.bias", {C}));
          if (!I.gn1_g || !I.gn1_b || !I.gn2_g || !I.gn2_b) return SFB_ERR_MISSING;
          if (d == 0) {
            for (int k = 0; k < 2; ++k) {
              const HostTensor* w = get(Q + (k ? "resnet.conv2.weight" : "resnet.conv1.weight"), {8, 8, 3});
              const HostTensor* b = get(Q + (k ? "resnet.conv2.bias" : "resnet.conv1.bias"), {8});
              if (!w || !b) return SFB_ERR_MISSING;
              std::vector<float> r(192);
              for (int co = 0; co < 8; ++co)
                for (int ci = 0; ci < 8; ++ci)
                  for (int t = 0; t < 3; ++t) r[(t * 8 + ci) * 8 + co] = w->v[(co * 8 + ci) * 3 + t];
              (k ? I.c8_w2 : I.c8_w1) = upload<float>(r);
              (k ? I.c8_b2 : I.c8_b1) = upload_f(b);
            }
          } else {
            if (!pack_conv3(Q + "resnet.conv1", C, I.conv1, &keep1) || !pack_conv3(Q + "resnet.conv2", C, I.conv2, &keep2))
              return err.empty() ? fail(SFB_ERR_CUDA, "weight upload failed") : SFB_ERR_MISSING;
          }
          {  // Modulation Linear(MF -> 2C)
            const HostTensor* w = get(Q + "mod.linear.weight", {2 * C, MF});
            const HostTensor* b = get(Q + "mod.linear.bias", {2 * C});
            if (!w || !b) return SFB_ERR_MISSING;
            I.mod_off = F_total;
            ftw.insert(ftw.end(), w->v.begin(), w->v.end());
            ftb.insert(ftb.end(), b->v.begin(), b->v.end());
            F_total += 2 * C;
          }
          I.has_inject = ctx > 0;
          if (I.has_inject) {
            const HostTensor* w = get(Q + "inject.conv.weight", {C, C + ctx, 1});
            const HostTensor* b = get(Q + "inject.conv.bias", {C});
            if (!w || !b) return SFB_ERR_MISSING;
            if (d == 0) {
              std::vector<float> r((size_t)(8 + ctx) * 8);
              for (int co = 0; co < 8; ++co)
                for (int k = 0; k < 8 + ctx; ++k) r[k * 8 + co] = w->v[(size_t)co * (8 + ctx) + k];
              I.c8_wi = upload<float>(r); I.c8_bi = upload_f(b);
            } else {
              I.inject.w = upload_T(w->v); I.inject.bias = upload_f(b);
              I.inject.N = C; I.inject.K1 = C; I.inject.K2 = ctx; I.inject.taps = 1; I.inject.bias_mod = C;
              // fused item: R1 = GN1+SiLU -> conv1 ; R2 = GN2+SiLU -> conv2 + x -> Modulation -> inject  (rk_tc.cuh)
              const int id1 = rk_find(C, C, 3, 2, 0, C, 0), id2 = rk_find(C, C, 3, 1, 2, C, 1);
              if (kBF16 && !no_rk && id1 >= 0 && id2 >= 0 && ctx % 8 == 0 && ctx <= rk_ctx_capacity(C)) {
                std::vector<__nv_bfloat16> t1, t2;
                pack_tiles(t1, keep1.data(), 3, C, C);
                pack_tiles(t2, keep2.data(), 3, C, C);
                pack_tiles(t2, w->v.data(), 1, C, C + ctx);
                I.rk1_w = upload<__nv_bfloat16>(t1);
                I.rk2_w = upload<__nv_bfloat16>(t2);
                I.rk1_id = id1; I.rk2_id = id2;
              }
            }
          } else {
            return fail(SFB_ERR_UNSUPPORTED, "context_channels[%d] == 0 unsupported", d);
          }
          I.has_attn = c.attentions[d] != 0;
          if (I.has_attn) {
            const HostTensor* g1 = get(Q + "attn.attn.norm.weight", {C});
            const HostTensor* b1 = get(Q + "attn.attn.norm.bias", {C});
            const HostTensor* g2 = get(Q + "attn.attn.norm_ctx.weight", {C});
            const HostTensor* b2 = get(Q + "attn.attn.norm_ctx.bias", {C});
            const HostTensor* wq = get(Q + "attn.attn.to_q.weight", {mid, C});
            const HostTensor* wkv = get(Q + "attn.attn.to_kv.weight", {2 * mid, C});
            const HostTensor* wo = get(Q + "attn.attn.to_out.weight", {C, mid});
            if (!g1 || !b1 || !g2 || !b2 || !wq || !wkv || !wo) return SFB_ERR_MISSING;
            // fold the two LayerNorm affines into the fused QKV projection: W' = W diag(g), bias' = W b
            std::vector<float> r((size_t)3 * mid * C), bb(3 * mid);
            for (int n = 0; n < 3 * mid; ++n) {
              const bool isq = n < mid;
              const float* wr = isq ? &wq->v[(size_t)n * C] : &wkv->v[(size_t)(n - mid) * C];
              const std::vector<float>& g = isq ? g1->v : g2->v;
              const std::vector<float>& bt = isq ? b1->v : b2->v;
              double acc = 0;
              for (int k = 0; k < C; ++k) {
                r[(size_t)n * C + k] = wr[k] * g[k];
                acc += (double)wr[k] * bt[k];
              }
              bb[n] = (float)acc;
            }
            I.qkv.w = upload_T(r); I.qkv.bias = upload<float>(bb);
            {
              std::vector<float> wsum(3 * mid);
              for (int n = 0; n < 3 * mid; ++n) {
                double acc = 0;
                for (int k = 0; k < C; ++k) acc += (double)(float)host_cvt<T>(r[(size_t)n * C + k]);
                wsum[n] = (float)acc;
              }
              I.qkv_ws = upload<float>(wsum);
            }
            I.qkv.N = 3 * mid; I.qkv.K1 = C; I.qkv.taps = 1; I.qkv.bias_mod = 3 * mid;
            I.out.w = upload_T(wo->v); I.out.bias = nullptr;
            I.out.N = C; I.out.K1 = mid; I.out.taps = 1; I.out.bias_mod = C;
          }
          I.has_xattn = c.cross_attentions[d] != 0;
          if (I.has_xattn) {
            // M_ctx = 1 fast path: softmax over one key == 1, so the item is x + W_o V(LN_ctx(e)).
            const HostTensor* g = get(Q + "xattn.attn.norm_ctx.weight", {EF});
            const HostTensor* b = get(Q + "xattn.attn.norm_ctx.bias", {EF});
            const HostTensor* wkv = get(Q + "xattn.attn.to_kv.weight", {2 * mid, EF});
            const HostTensor* wo = get(Q + "xattn.attn.to_out.weight", {C, mid});
            if (!g || !b || !wkv || !wo) return SFB_ERR_MISSING;
            // (norm / to_q of the query side provably do not influence the output when M_ctx = 1; they are
            //  still required to be present so a full state_dict loads without surprises.)
            if (!get(Q + "xattn.attn.to_q.weight", {mid, C}) || !get(Q + "xattn.attn.norm.weight", {C}) ||
                !get(Q + "xattn.attn.norm.bias", {C}))
              return SFB_ERR_MISSING;
            std::vector<float> wv(wkv->v.begin() + (size_t)mid * EF, wkv->v.end());
            I.x_ng = upload_f(g); I.x_nb = upload_f(b); I.x_wv = upload<float>(wv); I.x_wo = upload_f(wo);
            I.xb_off = XB_total;
            XB_total += C;
            if (c.embedding_max_length > 1) {      // general M_ctx path (K6): W_q' = W_q diag(g_q), b_q = W_q beta_q; full to_kv; to_out
              const HostTensor* wq = get(Q + "xattn.attn.to_q.weight", {mid, C});
              const HostTensor* gq = get(Q + "xattn.attn.norm.weight", {C});
              const HostTensor* bq = get(Q + "xattn.attn.norm.bias", {C});
              std::vector<float> r((size_t)mid * C), bb(mid);
              for (int n = 0; n < mid; ++n) {
                double acc = 0;
                for (int k = 0; k < C; ++k) {
                  r[(size_t)n * C + k] = wq->v[(size_t)n * C + k] * gq->v[k];
                  acc += (double)wq->v[(size_t)n * C + k] * bq->v[k];
                }
                bb[n] = (float)acc;
              }
              I.xq.w = upload_T(r); I.xq.bias = upload<float>(bb);
              I.xq.N = mid; I.xq.K1 = C; I.xq.taps = 1; I.xq.bias_mod = mid;
              I.xout.w = upload_T(wo->v); I.xout.bias = nullptr;
              I.xout.N = C; I.xout.K1 = mid; I.xout.taps = 1; I.xout.bias_mod = C;
              I.x_wkv = upload_f(wkv);
              I.xkv_index = n_xattn++;
              if (!I.xq.w || !I.xq.bias || !I.xout.w || !I.x_wkv) return fail(SFB_ERR_CUDA, "upload failed");
            }
          }
        }
      }
    }
    ft_w = upload<float>(ftw);
    ft_b = upload<float>(ftb);
    if (!ft_w || !ft_b) return fail(SFB_ERR_CUDA, "upload failed");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "finalize: %s", cudaGetErrorString(e));
    finalized = true;
    params.clear();
    return SFB_OK;
  }

  // ---------------------------------------------------------------- workspace
  int Ld(int64_t L, int d) const {
    int64_t f = 1;
    for (int i = 0; i <= d; ++i) f *= cfg.factors[i];
    return (int)(L / f);
  }
  void layout(int64_t B, int64_t L, int cfg_on, int64_t rows, int64_t M, WsLayout& w) const {
    const int64_t Beff = cfg_on ? 2 * B : B;
    const int MF = cfg.modulation_features, EF = cfg.embedding_features;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    const int64_t R = std::max<int64_t>(rows, 1);
    w.sigma = take(R * 4);
    w.embrows = take(Beff * M * EF * 4);
    w.tmp1 = take(Beff * M * EF * 4);
    w.tmp2 = take(Beff * M * std::max(EF, 1024) * 4);
    w.fourier = take(R * 257 * 4);
    w.h1 = take(R * MF * 4);
    w.h2 = take(R * MF * 4);
    w.feat = take(R * MF * 4);
    w.ftable = take(R * (size_t)F_total * 4);
    w.xbias = take(Beff * (size_t)std::max(XB_total, 1) * 4);
    int ngn = 0;
    for (int d = 0; d < cfg.depth; ++d) ngn += (M > 1 ? 6 : 4) * cfg.items[d] + 2;
    w.stats_bytes = (size_t)ngn * Beff * 16 * sizeof(double);
    w.stats = take(w.stats_bytes);
    w.veff = take(Beff * L * 4);
    w.xstate = take(B * L * 4);
    for (int d = 0; d < cfg.depth; ++d) {
      const size_t ld = Ld(L, d), C = cfg.channels[d];
      w.ctx[d] = take((size_t)B * ld * cfg.context_channels[d] * sizeof(T));
      w.bufA[d] = take((size_t)Beff * ld * C * 4);
      w.T1[d] = take((size_t)Beff * ld * C * sizeof(T));
      w.T2[d] = take((size_t)Beff * ld * C * sizeof(T));
      w.rs1[d] = w.rs2[d] = 0;
      if (kBF16 && C % 128 == 0) {
        w.rs1[d] = take((size_t)Beff * ld * 8 * 2 * 4);
        w.rs2[d] = take((size_t)Beff * ld * 8 * 2 * 4);
      }
      if (cfg.attentions[d]) {
        w.qkv[d] = take((size_t)Beff * ld * 1536 * sizeof(T));
        w.o[d] = take((size_t)Beff * ld * 512 * sizeof(T));
      } else {
        w.qkv[d] = w.o[d] = 0;
      }
    }
    w.foldw_bytes = 0;
    for (int d = 0; d < cfg.depth; ++d)
      if (depth_sk(d))
        for (int s = 0; s < 2; ++s)
          for (const ItemW& I : dw[d].items[s])
            w.foldw_bytes += fold_item_bytes(I, B);
    w.foldw = take(std::max<size_t>(w.foldw_bytes, 16));
    w.xln = w.xq = w.xo = w.xkv = 0;
    if (M > 1) {
      size_t lc = 0;
      for (int d = 0; d < cfg.depth; ++d) lc = std::max(lc, (size_t)Ld(L, d) * cfg.channels[d]);
      w.xln = take((size_t)Beff * lc * sizeof(T));
      w.xq = take((size_t)Beff * L * 512 * sizeof(T));
      w.xo = take((size_t)Beff * L * 512 * sizeof(T));
      w.xkv = take((size_t)std::max(n_xattn, 1) * Beff * M * 1024 * sizeof(T));
    }
    w.total = off;
  }
  // scaled weight copies [B][N][K] bf16 + vectors [B][2 N] fp32 of one streaming-K inject item
  static size_t fold_w_bytes(const ItemW& I, int64_t copies) { return align_up((size_t)copies * I.inject.N * (I.inject.K1 + I.inject.K2) * 2, 1024); }
  static size_t fold_item_bytes(const ItemW& I, int64_t copies) { return fold_w_bytes(I, copies) + align_up((size_t)copies * 2 * I.inject.N * 4, 1024); }
  int workspace_bytes(int64_t B, int64_t L, int cfg_on, int64_t rows, int64_t M, size_t* out) override {
    if (M < 1 || M > cfg.embedding_max_length) return fail(SFB_ERR_INVALID, "embedding length M=%lld must be in [1, embedding_max_length=%d]", (long long)M, cfg.embedding_max_length);
    if (M > 128) return fail(SFB_ERR_UNSUPPORTED, "cross-attention context longer than 128 tokens");
    if (!finalized) return fail(SFB_ERR_STATE, "finalize first");
    int64_t tf = 1;
    for (int d = 0; d < cfg.depth; ++d) tf *= cfg.factors[d];
    if (B <= 0 || L <= 0 || L % tf) return fail(SFB_ERR_INVALID, "L=%lld must be a positive multiple of %lld", (long long)L, (long long)tf);
    if (L % 8) return fail(SFB_ERR_INVALID, "L must be a multiple of 8");
    WsLayout w;
    layout(B, L, cfg_on, rows, M, w);
    *out = w.total + 1024;
    return SFB_OK;
  }

  // ---------------------------------------------------------------- plan
  uint8_t* wsb = nullptr;
  template <typename U> U* at(size_t off) { return reinterpret_cast<U*>(wsb + off); }
  int stats_next = 0;
  double* new_stats(int Beff) { return at<double>(plan.lay.stats) + (size_t)(stats_next++) * Beff * 16; }

  bool add_gemm(Op& op, const GemmW& g, const void* a1, int K1view, int L, int Beff, const void* a2, int B2) {
    op.kind = OP_GEMM;
    op.BN = pick_bn(g.N);
    op.B = Beff; op.L = L; op.C = g.N; op.k2 = g.K2;
    GemmParams<T>& p = op.gp;
    memset(&p, 0, sizeof p);
    if (!fill_gemm_maps<T>(p, a1, K1view, L, Beff, a2, g.K2, B2, g.w, g.N, g.taps, op.BN)) return false;
    p.bias = g.bias;
    p.bias_mod = g.bias_mod;
    p.gs = 1;
    p.resid = op.resid;
    p.out_r = op.out_r;
    p.out_t = reinterpret_cast<T*>(op.out_t);
    p.stats = op.stats_out;
    return true;
  }
  // Resident-weight fused op (rk_tc.cuh): A / W maps and tile bookkeeping; the caller adds R / T / C maps and vectors.
  bool add_rk(Op& op, int id, const void* a, const void* wt, int L, int Beff) {
    const RkKey k = rk_key(id);
    op.kind = OP_RK; op.rk_id = id; op.B = Beff; op.L = L; op.C = k.N;
    RkParams& p = op.rp;
    memset(&p, 0, sizeof p);
    const int rows_a = k.TAPS == 3 ? 136 : 128;
    const int KA = (k.K1 + 63) / 64, KA2 = k.EPI == 2 ? (k.N + 64) / 64 : 0;
    if (k.XF == 2) { if (!make_tmap3<float>(&p.tmA, a, k.K1, L, Beff, 32, rows_a)) return false; }
    else if (!make_tmap3<__nv_bfloat16>(&p.tmA, a, k.K1, L, Beff, 64, rows_a)) return false;
    if (!make_tmap2<__nv_bfloat16>(&p.tmW, wt, 64, (uint64_t)(k.TAPS * KA + KA2) * k.N, 64, k.N)) return false;
    p.tmR = p.tmA; p.tmT = p.tmA; p.tmC = p.tmA;
    p.L = L; p.tiles_per_clip = (L + 127) / 128; p.total_tiles = Beff * p.tiles_per_clip;
    p.cs_bmod = 1; p.mod_bmod = 1; p.ctx_bmod = 1;
    p.eps = 1e-5f;
    return true;
  }
  bool rk_out_r(Op& op, float* ptr, bool resid) {
    const RkKey k = rk_key(op.rk_id);
    op.out_r = ptr; op.rp.has_out_r = 1; op.rp.has_resid = resid ? 1 : 0;
    if (resid) op.resid = ptr;
    return make_tmap3<float>(&op.rp.tmR, ptr, k.N, op.L, op.B, 32, 128);
  }
  bool rk_out_t(Op& op, void* ptr) {
    const RkKey k = rk_key(op.rk_id);
    op.out_t = ptr; op.rp.has_out_t = 1;
    return make_tmap3<__nv_bfloat16>(&op.rp.tmT, ptr, k.N, op.L, op.B, 64, 128);
  }
  // Streaming-K fused GEMM (sk_tc.cuh).  gs_out: GroupNorm group size of the output statistics (0: none).
  bool sk_ok(const GemmW& g) const { return kBF16 && !no_sk && g.N % 128 == 0 && g.K1 % 64 == 0 && g.w != nullptr; }
  bool add_sk(Op& op, const GemmW& g, const void* a1, int L, int Beff, const void* a2, int B2, int gs_out,
              const void* w_override = nullptr, int w_copies = 1) {
    const int BN = g.N % 256 == 0 ? 256 : 128;
    int id = sk_find(BN, gs_out > 0 ? gs_out : (BN == 256 ? 32 : 16));
    if (id < 0) return false;
    op.kind = OP_SK; op.sk_id = id; op.B = Beff; op.L = L; op.C = g.N; op.k2 = g.K2; op.BN = BN;
    SkParams& p = op.sp;
    memset(&p, 0, sizeof p);
    if (!make_tmap3<__nv_bfloat16>(&p.tmA1, a1, g.K1, L, Beff, 64, g.taps == 3 ? 136 : 128)) return false;
    if (g.K2 > 0) { if (!make_tmap3<__nv_bfloat16>(&p.tmA2, a2, g.K2, L, B2, 64, 128)) return false; }
    else p.tmA2 = p.tmA1;
    if (!make_tmap3<__nv_bfloat16>(&p.tmW, w_override ? w_override : g.w, (uint64_t)(g.K1 + g.K2), (uint64_t)g.taps * g.N,
                                   (uint64_t)w_copies, 64, BN)) return false;
    p.w_bmod = 1;
    p.w_static = (w_override == nullptr && !getenv("SFB_NO_WPRE")) ? 1 : 0;   // per-evaluation scaled copies are written inside the chain
    p.tmR = p.tmA1; p.tmRs = p.tmA1; p.tmT = p.tmA1;
    p.L = L; p.tiles_per_clip = (L + 127) / 128; p.N = g.N; p.n_tiles = g.N / BN;
    p.total_tiles = Beff * p.tiles_per_clip * p.n_tiles;
    p.taps = g.taps; p.k1_chunks = g.K1 / 64; p.k2_chunks = (g.K2 + 63) / 64; p.K1 = g.K1; p.a2_bmod = B2 > 0 ? B2 : 1;
    p.bias = g.bias; p.bias_mod = g.bias_mod > 0 ? g.bias_mod : g.N;
    p.cs_bmod = 1; p.mod_bmod = 1; p.rs_parts = 1;
    p.eps = 1e-5f;
    op.flops = 2.0 * Beff * L * (double)g.N * ((double)g.taps * g.K1 + g.K2);
    op.bytes = (double)Beff * L * (g.K1 + g.K2) * 2;
    return true;
  }
  bool sk_out_r(Op& op, float* ptr, int resid_mode) {
    op.out_r = ptr; op.sp.has_out_r = 1; op.sp.resid_mode = resid_mode;
    if (resid_mode) op.resid = ptr;
    op.bytes += (double)op.B * op.L * op.sp.N * (resid_mode ? 8 : 4);
    return make_tmap3<float>(&op.sp.tmR, ptr, op.sp.N, op.L, op.B, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B) &&
           make_tmap3<float>(&op.sp.tmRs, ptr, op.sp.N, op.L, op.B, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  bool sk_out_t(Op& op, void* ptr) {
    op.out_t = ptr; op.sp.has_out_t = 1;
    op.bytes += (double)op.B * op.L * op.sp.N * 2;
    return make_tmap3<__nv_bfloat16>(&op.sp.tmT, ptr, op.sp.N, op.L, op.B, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  }
  void set_dbg(Op& op, int rows, int cols) {
    if (op.out_r) { op.dbg_off = (uint8_t*)op.out_r - wsb; op.dbg_dtype = 0; op.dbg_bytes = (size_t)rows * cols * 4; }
    else { op.dbg_off = (uint8_t*)op.out_t - wsb; op.dbg_dtype = 1; op.dbg_bytes = (size_t)rows * cols * sizeof(T); }
    op.dbg_rows = rows; op.dbg_cols = cols;
  }

  // ---- streaming-K fused path (bf16, C % 128 == 0): every norm is folded into a GEMM prologue / epilogue -------------
  std::vector<FoldItem> fold_items;   // streaming-K inject items of the current plan (LayerNorm fold)
  FoldItem* fold_items_dev = nullptr;
  size_t fold_next = 0, fold_items_cap = 0;
  int fold_rows = 0;
  void* xt_cur[SFB_MAX_DEPTH] = {};   // bf16 copy of the tensor currently in bufA[d] (written by whoever produced it)
  bool depth_sk(int d) const {
    if (!kBF16 || no_sk || d == 0 || cfg.channels[d] % 128) return false;
    if (!sk_ok(dw[d].down)) return false;
    for (int s = 0; s < 2; ++s)
      for (const ItemW& I : dw[d].items[s]) {
        if (!sk_ok(I.conv1) || !sk_ok(I.conv2) || !sk_ok(I.inject)) return false;
        if (I.has_attn && (!sk_ok(I.qkv) || !sk_ok(I.out))) return false;
      }
    if (d + 1 < cfg.depth && !sk_ok(dw[d + 1].up)) return false;
    return true;
  }
  // General cross-attention item (a10, K6; M_ctx > 1): x + W_o softmax(q k^T / 8) v with q = W_q LN(x) (query LayerNorm
  // affine folded into W_q), k | v = to_kv(LN_ctx(e)) precomputed per call.  Four ops behind the item's last op:
  // LayerNorm pass -> q projection (tcgen05 GEMM) -> attention core (the self-attention kernel with its own K/V source
  // and a masked partial key tile) -> output projection + residual (+ operand copy and GroupNorm sums for the consumer).
  int append_xattn(int d, int s, int i, const ItemW& I, float* A, void* out_t, double* stats_out) {
    const int Beff = plan.cfg_on ? 2 * (int)plan.B : (int)plan.B;
    const int C = cfg.channels[d], L = Ld(plan.L, d), gs = C / 8, rows = Beff * L, M = (int)plan.M;
    T* XLN = at<T>(plan.lay.xln);
    T* XQ = at<T>(plan.lay.xq);
    T* XO = at<T>(plan.lay.xo);
    T* KV = at<T>(plan.lay.xkv) + (size_t)I.xkv_index * Beff * M * 1024;
    auto base = [&](int kind, const char* ck) { Op o; o.kind = kind; o.depth = d; o.stack = s; o.item = i; o.L = L; o.C = C; o.gs = gs; o.B = Beff; o.ck = ck; return o; };
    {
      Op o = base(OP_LN, "xattn_ln"); o.in = A; o.out_t = XLN; o.out_r = nullptr; o.ft_off = -1;
      set_dbg(o, rows, C); plan.ops.push_back(o);
    }
    {
      Op o = base(OP_GEMM, "xattn_q"); o.out_t = XQ;
      if (!add_gemm(o, I.xq, XLN, C, L, Beff, nullptr, 0)) return fail(SFB_ERR_CUDA, "tensor map encode failed (xattn q d%d)", d);
      set_dbg(o, rows, 512); plan.ops.push_back(o);
    }
    {
      Op o = base(OP_ATTN, "xattn"); o.in = XQ; o.out_t = XO;
      constexpr int AE = ElemTraits<T>::kAtomElems;
      if (!make_tmap3<T>(&o.ap.tmQ, XQ, 512, L, Beff, AE, 128) ||
          !make_tmap3<T>(&o.ap.tmKV, KV, 1024, M, Beff, AE, AttnCfg<T>::BKV) ||
          !make_tmap3<T>(&o.ap.tmV, KV, 1024, M, Beff, AE, AttnCfg<T>::BKV,
                         sizeof(T) == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return fail(SFB_ERR_CUDA, "tensor map encode failed (cross-attention d%d)", d);
      o.ap.out = XO; o.ap.n_tokens = L; o.ap.kv_tokens = M; o.ap.q_col0 = 0; o.ap.k_col0 = 0; o.ap.v_col0 = 512;
      o.ap.scale_log2 = 1.4426950408889634f / 8.0f;
      set_dbg(o, rows, 512); plan.ops.push_back(o);
    }
    if (d == 0) {
      Op o = base(OP_XOUT_C8, "xattn_out"); o.in = XO; o.w0 = I.x_wo; o.resid = A; o.out_r = A; o.out_t = out_t; o.stats_out = stats_out;
      set_dbg(o, rows, C); plan.ops.push_back(o);
    } else {
      Op o = base(OP_GEMM, "xattn_out"); o.resid = A; o.out_r = A; o.out_t = out_t; o.stats_out = stats_out;
      if (!add_gemm(o, I.xout, XO, 512, L, Beff, nullptr, 0)) return fail(SFB_ERR_CUDA, "tensor map encode failed (xattn out d%d)", d);
      o.gp.gs = gs;
      set_dbg(o, rows, C); plan.ops.push_back(o);
    }
    return SFB_OK;
  }

  int build_item_sk(int d, int s, int i, double*& cur, bool want_stats) {
    const int64_t B = plan.B;
    const int Beff = plan.cfg_on ? 2 * (int)B : (int)B;
    const int C = cfg.channels[d], L = Ld(plan.L, d), gs = C / 8;
    const ItemW& I = dw[d].items[s][i];
    float* A = at<float>(plan.lay.bufA[d]);
    T* T1 = at<T>(plan.lay.T1[d]);
    T* T2 = at<T>(plan.lay.T2[d]);
    float* RS1 = at<float>(plan.lay.rs1[d]);
    float* RS2 = at<float>(plan.lay.rs2[d]);
    void* P0 = xt_cur[d];
    void* P1 = (P0 == (void*)T1) ? (void*)T2 : (void*)T1;
    const float* xb = (I.has_xattn && plan.M == 1) ? at<float>(plan.lay.xbias) + I.xb_off : nullptr;   // M = 1: collapsed to a bias
    auto base = [&](const char* ck) { Op o; o.kind = OP_SK; o.depth = d; o.stack = s; o.item = i; o.L = L; o.C = C; o.gs = gs; o.B = Beff; o.ck = ck; return o; };
    const int rows = Beff * L;
    double* sB = new_stats(Beff);
    double* sOut = want_stats ? new_stats(Beff) : nullptr;
    const char* err_fmt = "tensor map encode failed (sk %s d%d)";
    {  // conv1: h = conv3(SiLU(GN1(x))) + b  ->  P1 (bf16) + GroupNorm sums of h
      Op o = base("conv1"); o.in = P0; o.stats_in = cur; o.stats_out = sB;
      if (!add_sk(o, I.conv1, P0, L, Beff, nullptr, 0, gs) || !sk_out_t(o, P1)) return fail(SFB_ERR_CUDA, err_fmt, "conv1", d);
      o.sp.xf = 1; o.sp.stats_in = cur; o.sp.gamma = I.gn1_g; o.sp.beta = I.gn1_b; o.sp.stats_out = sB;
      set_dbg(o, rows, C); plan.ops.push_back(o);
    }
    int parts1;
    {  // conv2: r = conv3(SiLU(GN2(h))) + b + x  ->  bufA (fp32), P0 (bf16), per-position LayerNorm sums of r
      Op o = base("conv2"); o.in = P1; o.stats_in = sB;
      if (!add_sk(o, I.conv2, P1, L, Beff, nullptr, 0, 0) || !sk_out_r(o, A, 1) || !sk_out_t(o, P0)) return fail(SFB_ERR_CUDA, err_fmt, "conv2", d);
      o.sp.xf = 1; o.sp.stats_in = sB; o.sp.gamma = I.gn2_g; o.sp.beta = I.gn2_b; o.sp.rowstats_out = RS1;
      parts1 = 2 * o.sp.n_tiles;
      set_dbg(o, rows, C); plan.ops.push_back(o);
    }
    const bool last_is_inject = !I.has_attn;
    int parts2;
    {  // inject: i = W [Mod(LN(r)) | ctx] + b + Mod(LN(r)) (+ cross-attention bias)  ->  bufA (fp32), P1 (bf16)
      Op o = base("inject"); o.in = P0; o.ft_off = I.mod_off; o.sk_ft_is_mod = 1; o.ctx = cfg.context_channels[d];
      // LayerNorm fold: the MMA reads the raw bf16 rows of r against W diag(1 + s) (rebuilt per evaluation by
      // inject_fold_kernel), the context rows are pre-divided by rstd, the epilogue applies rstd (D - mean ws) + W sh.
      uint8_t* fw = at<uint8_t>(plan.lay.foldw) + fold_next;
      float* fvec = reinterpret_cast<float*>(fw + fold_w_bytes(I, B));
      fold_next += fold_item_bytes(I, B);
      FoldItem fi;
      fi.W = reinterpret_cast<const __nv_bfloat16*>(I.inject.w); fi.Wd = reinterpret_cast<__nv_bfloat16*>(fw); fi.vec = fvec;
      fi.C = C; fi.K = I.inject.K1 + I.inject.K2; fi.N = I.inject.N; fi.mod_off = I.mod_off; fi.row0 = fold_rows;
      fold_rows += fi.N;
      fold_items.push_back(fi);
      if (!add_sk(o, I.inject, P0, L, Beff, at<T>(plan.lay.ctx[d]), (int)B, last_is_inject && want_stats ? gs : 0, fw, (int)B) ||
          !sk_out_r(o, A, 2) || !sk_out_t(o, P1)) return fail(SFB_ERR_CUDA, err_fmt, "inject", d);
      o.sp.xf = 3; o.sp.ln_fold = 1; o.sp.ws = fvec; o.sp.addvec = fvec + I.inject.N; o.sp.ws_bstride = 2 * I.inject.N;
      o.sp.rowstats_in = RS1; o.sp.rs_parts = parts1;
      if (last_is_inject) {
        o.stats_out = sOut; o.sp.stats_out = sOut;
        if (xb) { o.sp.rowvec = xb; o.sp.rowvec_stride = XB_total; }
      } else {
        o.sp.rowstats_out = RS2;
      }
      parts2 = 2 * o.sp.n_tiles;
      set_dbg(o, rows, C); plan.ops.push_back(o);
    }
    xt_cur[d] = P1;
    if (I.has_attn) {
      T* QKV = at<T>(plan.lay.qkv[d]);
      T* O = at<T>(plan.lay.o[d]);
      {  // fused pre-norm + QKV projection (both LayerNorm affines are folded into W_qkv)
        Op o = base("qkv"); o.in = P1;
        if (!add_sk(o, I.qkv, P1, L, Beff, nullptr, 0, 0) || !sk_out_t(o, QKV)) return fail(SFB_ERR_CUDA, err_fmt, "qkv", d);
        o.sp.xf = 0; o.sp.ln_fold = 1; o.sp.ws = I.qkv_ws; o.sp.ws_bstride = 0; o.sp.rowstats_in = RS2; o.sp.rs_parts = parts2;
        set_dbg(o, rows, 1536); plan.ops.push_back(o);
      }
      {
        Op o; o.kind = OP_ATTN; o.depth = d; o.stack = s; o.item = i; o.L = L; o.C = C; o.gs = gs; o.B = Beff; o.ck = "attn"; o.in = QKV; o.out_t = O;
        constexpr int AE = ElemTraits<T>::kAtomElems;
        if (!make_tmap3<T>(&o.ap.tmQ, QKV, 1536, L, Beff, AE, 128) ||
            !make_tmap3<T>(&o.ap.tmKV, QKV, 1536, L, Beff, AE, AttnCfg<T>::BKV) ||
            !make_tmap3<T>(&o.ap.tmV, QKV, 1536, L, Beff, AE, AttnCfg<T>::BKV,
                           sizeof(T) == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
          return fail(SFB_ERR_CUDA, "tensor map encode failed (attention d%d)", d);
        o.ap.out = O; o.ap.n_tokens = L; o.ap.kv_tokens = L; o.ap.q_col0 = 0; o.ap.k_col0 = 512; o.ap.v_col0 = 1024; o.ap.scale_log2 = 1.4426950408889634f / 8.0f;
        set_dbg(o, rows, 512); plan.ops.push_back(o);
      }
      {  // out projection + residual (+ cross-attention bias)  ->  bufA (fp32), P0 (bf16), GroupNorm sums
        Op o = base("out"); o.in = O; o.stats_out = sOut;
        if (!add_sk(o, I.out, O, L, Beff, nullptr, 0, want_stats ? gs : 0) || !sk_out_r(o, A, 1) || !sk_out_t(o, P0)) return fail(SFB_ERR_CUDA, err_fmt, "out", d);
        o.sp.stats_out = sOut;
        if (xb) { o.sp.rowvec = xb; o.sp.rowvec_stride = XB_total; }
        set_dbg(o, rows, C); plan.ops.push_back(o);
      }
      xt_cur[d] = P0;
    }
    item_t_out = xt_cur[d];
    cur = sOut;
    if (I.has_xattn && plan.M > 1) {      // the consumer reads the bf16 copy in xt_cur[d] and the statistics in `cur`
      double* sX = want_stats ? new_stats(Beff) : nullptr;
      int rc = append_xattn(d, s, i, I, A, xt_cur[d], sX);
      if (rc) return rc;
      cur = sX;
    }
    return SFB_OK;
  }

  // One [Resnet, Modulation, Inject, Attention?, CrossAttention?] group.  `cur` = stats of the tensor in bufA[d].
  void* item_t_out = nullptr;   // where the last built item left the operand-precision copy of its output
  int build_item(int d, int s, int i, double*& cur, bool want_t, bool want_stats) {
    const int64_t B = plan.B;
    const int Beff = plan.cfg_on ? 2 * (int)B : (int)B;
    const int C = cfg.channels[d], L = Ld(plan.L, d), gs = C / 8, ctx = cfg.context_channels[d];
    if (depth_sk(d)) { if constexpr (kBF16) return build_item_sk(d, s, i, cur, want_stats); }
    const ItemW& I = dw[d].items[s][i];
    float* A = at<float>(plan.lay.bufA[d]);
    T* T1 = at<T>(plan.lay.T1[d]);
    T* T2 = at<T>(plan.lay.T2[d]);
    const float* xb = (I.has_xattn && plan.M == 1) ? at<float>(plan.lay.xbias) + I.xb_off : nullptr;   // M = 1: collapsed to a bias
    auto base = [&](int kind, const char* ck) { Op o; o.kind = kind; o.depth = d; o.stack = s; o.item = i; o.L = L; o.C = C; o.gs = gs; o.B = Beff; o.ck = ck; return o; };
    const int rows = Beff * L;
    const bool last_is_inject = !I.has_attn;
    double* sB = new_stats(Beff);
    double* sOut = want_stats ? new_stats(Beff) : nullptr;
    item_t_out = T2;
    if (I.rk1_id >= 0) {
      // ---- fused path (bf16, C <= 64): two launches per item, every tensor crosses HBM once per launch
      if constexpr (kBF16) {
        {  // R1: h = conv1(SiLU(GN1(x))) + b1 -> T2 (bf16) + GroupNorm sums of h
          Op o = base(OP_RK, "conv1"); o.in = A; o.stats_in = cur; o.stats_out = sB;
          if (!add_rk(o, I.rk1_id, A, I.rk1_w, L, Beff) || !rk_out_t(o, T2)) return fail(SFB_ERR_CUDA, "tensor map encode failed (rk1 d%d)", d);
          o.rp.stats_in = cur; o.rp.stats_out = sB; o.rp.gamma = I.gn1_g; o.rp.beta = I.gn1_b; o.rp.bias = I.conv1.bias;
          o.flops = 2.0 * rows * C * 3.0 * C; o.bytes = (double)rows * C * (4 + 2);
          set_dbg(o, rows, C); plan.ops.push_back(o);
        }
        {  // R2: y = inject(Mod(conv2(SiLU(GN2(h))) + b2 + x)) (+ cross-attention bias) -> bufA in place (+ bf16 copy -> T1)
          Op o = base(OP_RK, "inject"); o.in = T2; o.stats_in = sB; o.ft_off = I.mod_off; o.ctx = ctx;
          if (!add_rk(o, I.rk2_id, T2, I.rk2_w, L, Beff) || !rk_out_r(o, A, true)) return fail(SFB_ERR_CUDA, "tensor map encode failed (rk2 d%d)", d);
          if (last_is_inject && want_t) { if (!rk_out_t(o, T1)) return fail(SFB_ERR_CUDA, "tensor map encode failed (rk2 d%d)", d); item_t_out = T1; }
          if (!make_tmap3<__nv_bfloat16>(&o.rp.tmC, at<T>(plan.lay.ctx[d]), ctx, L, (int)B, ctx, 128, CU_TENSOR_MAP_SWIZZLE_NONE))
            return fail(SFB_ERR_CUDA, "tensor map encode failed (rk2 ctx d%d)", d);
          o.rp.ctx_bmod = (int)B; o.rp.ctx_ch = ctx;
          o.rp.stats_in = sB; o.rp.gamma = I.gn2_g; o.rp.beta = I.gn2_b; o.rp.bias = I.conv2.bias; o.rp.bias2 = I.inject.bias;
          if (last_is_inject) {
            o.stats_out = sOut; o.rp.stats_out = sOut;
            if (xb) { o.rp.rowvec = xb; o.rp.rowvec_stride = XB_total; }
          }
          o.flops = 2.0 * rows * C * (3.0 * C + C + ctx); o.bytes = (double)rows * (C * (2 + 4 + 4 + (o.rp.has_out_t ? 2 : 0)) + ctx * 2);
          set_dbg(o, rows, C); plan.ops.push_back(o);
        }
      }
    } else if (d == 0 && !no_d0_fused && ctx == 2) {
      // ---- fused depth-0 item (d0.cuh): two launches, every tensor crosses HBM once per launch
      {
        Op o = base(OP_D0_CONV1, "conv1"); o.in = A; o.stats_in = cur; o.w0 = I.gn1_g; o.w1 = I.gn1_b; o.wx[0] = I.c8_w1; o.wx[1] = I.c8_b1;
        o.out_t = T2; o.stats_out = sB;
        set_dbg(o, rows, C); plan.ops.push_back(o);
      }
      {
        Op o = base(OP_D0_TAIL, "inject"); o.in = T2; o.stats_in = sB; o.w0 = I.gn2_g; o.w1 = I.gn2_b; o.wx[0] = I.c8_w2; o.wx[1] = I.c8_b2;
        o.resid = A; o.out_r = A; o.ft_off = I.mod_off; o.in2 = at<T>(plan.lay.ctx[d]); o.ctx = ctx; o.wx[2] = I.c8_wi; o.wx[3] = I.c8_bi;
        o.w2 = xb; o.out_t = want_t ? T1 : nullptr; o.stats_out = sOut;
        if (want_t) item_t_out = T1;
        set_dbg(o, rows, C); plan.ops.push_back(o);
      }
    } else {
      {  // gn1 + SiLU
        Op o = base(OP_GN, "gn1"); o.in = A; o.in_is_f32 = 1; o.stats_in = cur; o.w0 = I.gn1_g; o.w1 = I.gn1_b; o.out_t = T1;
        set_dbg(o, rows, C); plan.ops.push_back(o);
      }
      if (d == 0) {
        Op o = base(OP_CONV_C8, "conv1"); o.in = T1; o.w0 = I.c8_w1; o.w1 = I.c8_b1; o.out_t = T2; o.stats_out = sB;
        set_dbg(o, rows, C); plan.ops.push_back(o);
      } else {
        Op o = base(OP_GEMM, "conv1"); o.out_t = T2; o.stats_out = sB;
        if (!add_gemm(o, I.conv1, T1, C, L, Beff, nullptr, 0)) return fail(SFB_ERR_CUDA, "tensor map encode failed (conv1 d%d)", d);
        o.gp.gs = gs; set_dbg(o, rows, C); plan.ops.push_back(o);
      }
      {  // gn2 + SiLU
        Op o = base(OP_GN, "gn2"); o.in = T2; o.in_is_f32 = 0; o.stats_in = sB; o.w0 = I.gn2_g; o.w1 = I.gn2_b; o.out_t = T1;
        set_dbg(o, rows, C); plan.ops.push_back(o);
      }
      if (d == 0) {
        Op o = base(OP_CONV_C8, "conv2"); o.in = T1; o.w0 = I.c8_w2; o.w1 = I.c8_b2; o.resid = A; o.out_r = A;
        set_dbg(o, rows, C); plan.ops.push_back(o);
      } else {
        Op o = base(OP_GEMM, "conv2"); o.resid = A; o.out_r = A;
        if (!add_gemm(o, I.conv2, T1, C, L, Beff, nullptr, 0)) return fail(SFB_ERR_CUDA, "tensor map encode failed (conv2 d%d)", d);
        set_dbg(o, rows, C); plan.ops.push_back(o);
      }
      {  // Modulation
        Op o = base(OP_LN, "mod"); o.in = A; o.out_t = T1; o.out_r = A; o.ft_off = I.mod_off;
        set_dbg(o, rows, C); plan.ops.push_back(o);
      }
      if (d == 0) {
        Op o = base(OP_INJ_C8, "inject"); o.in = T1; o.resid = A; o.in2 = at<T>(plan.lay.ctx[d]); o.ctx = ctx; o.w0 = I.c8_wi; o.w1 = I.c8_bi;
        o.w2 = xb; o.out_r = A; o.out_t = want_t ? T2 : nullptr; o.stats_out = sOut;
        set_dbg(o, rows, C); plan.ops.push_back(o);
      } else {
        Op o = base(OP_GEMM, "inject"); o.resid = A; o.out_r = A;
        o.out_t = (last_is_inject && want_t) ? T2 : nullptr;
        o.stats_out = last_is_inject ? sOut : nullptr;
        if (!add_gemm(o, I.inject, T1, C, L, Beff, at<T>(plan.lay.ctx[d]), (int)B)) return fail(SFB_ERR_CUDA, "tensor map encode failed (inject d%d)", d);
        o.gp.gs = gs;
        if (last_is_inject && xb) { o.gp.rowvec = xb; o.gp.rowvec_stride = XB_total; }
        set_dbg(o, rows, C); plan.ops.push_back(o);
      }
    }
    if (I.has_attn) {
      T* QKV = at<T>(plan.lay.qkv[d]);
      T* O = at<T>(plan.lay.o[d]);
      item_t_out = T2;
      {  // pre-norm (affine folded into W_qkv)
        Op o = base(OP_LN, "attn_ln"); o.in = A; o.out_t = T1; o.out_r = nullptr; o.ft_off = -1;
        set_dbg(o, rows, C); plan.ops.push_back(o);
      }
      {
        Op o = base(OP_GEMM, "qkv"); o.out_t = QKV;
        if (!add_gemm(o, I.qkv, T1, C, L, Beff, nullptr, 0)) return fail(SFB_ERR_CUDA, "tensor map encode failed (qkv d%d)", d);
        set_dbg(o, rows, 1536); plan.ops.push_back(o);
      }
      {
        Op o = base(OP_ATTN, "attn"); o.in = QKV; o.out_t = O;
        constexpr int AE = ElemTraits<T>::kAtomElems;
        if (!make_tmap3<T>(&o.ap.tmQ, QKV, 1536, L, Beff, AE, 128) ||
            !make_tmap3<T>(&o.ap.tmKV, QKV, 1536, L, Beff, AE, AttnCfg<T>::BKV) ||
            !make_tmap3<T>(&o.ap.tmV, QKV, 1536, L, Beff, AE, AttnCfg<T>::BKV,
                           sizeof(T) == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
          return fail(SFB_ERR_CUDA, "tensor map encode failed (attention d%d)", d);
        o.ap.out = O; o.ap.n_tokens = L; o.ap.kv_tokens = L; o.ap.q_col0 = 0; o.ap.k_col0 = 512; o.ap.v_col0 = 1024; o.ap.scale_log2 = 1.4426950408889634f / 8.0f;
        set_dbg(o, rows, 512); plan.ops.push_back(o);
      }
      {
        Op o = base(OP_GEMM, "out"); o.resid = A; o.out_r = A; o.out_t = want_t ? T2 : nullptr; o.stats_out = sOut;
        if (!add_gemm(o, I.out, O, 512, L, Beff, nullptr, 0)) return fail(SFB_ERR_CUDA, "tensor map encode failed (to_out d%d)", d);
        o.gp.gs = gs;
        if (xb) { o.gp.rowvec = xb; o.gp.rowvec_stride = XB_total; }
        set_dbg(o, rows, C); plan.ops.push_back(o);
      }
    }
    cur = sOut;
    if (I.has_xattn && plan.M > 1) {
      double* sX = want_stats ? new_stats(Beff) : nullptr;
      int rc = append_xattn(d, s, i, I, A, want_t ? item_t_out : nullptr, sX);
      if (rc) return rc;
      cur = sX;
    }
    return SFB_OK;
  }

  int build_block(int d) {
    const int64_t B = plan.B;
    const int Beff = plan.cfg_on ? 2 * (int)B : (int)B;
    const int C = cfg.channels[d], L = Ld(plan.L, d), f = cfg.factors[d];
    const int D = cfg.depth;
    const DepthW& W = dw[d];
    float* A = at<float>(plan.lay.bufA[d]);
    double* cur = new_stats(Beff);
    const void* down_in = item_t_out;   // operand copy of the previous depth's last down-stack item
    if (d == 0) {
      Op o; o.kind = OP_D0_DOWN; o.depth = 0; o.stack = 2; o.L = L; o.C = C; o.B = Beff; o.w0 = W.c8_dw; o.w1 = W.c8_db; o.ck = "down";
      o.out_r = A; o.stats_out = cur; set_dbg(o, Beff * L, C); plan.ops.push_back(o);
    } else if (W.rk_down_id >= 0) {
      const int Cin = cfg.channels[d - 1];
      Op o; o.depth = d; o.stack = 2; o.stats_out = cur; o.ck = "down"; o.in = down_in;
      if (!add_rk(o, W.rk_down_id, down_in, W.rk_down_w, L, Beff) || !rk_out_r(o, A, false))
        return fail(SFB_ERR_CUDA, "tensor map encode failed (rk down d%d)", d);
      o.rp.stats_out = cur; o.rp.bias = W.down.bias;
      o.flops = 2.0 * Beff * L * C * (double)f * Cin; o.bytes = (double)Beff * L * (f * Cin * 2 + C * 4);
      set_dbg(o, Beff * L, C); plan.ops.push_back(o);
    } else if (sk_ok(W.down)) {
      Op o; o.depth = d; o.stack = 2; o.stats_out = cur; o.ck = "down"; o.in = down_in;
      if constexpr (kBF16) {
        if (!add_sk(o, W.down, down_in, L, Beff, nullptr, 0, C / 8) || !sk_out_r(o, A, 0))
          return fail(SFB_ERR_CUDA, "tensor map encode failed (sk down d%d)", d);
        if (depth_sk(d)) {
          if (!sk_out_t(o, at<T>(plan.lay.T1[d]))) return fail(SFB_ERR_CUDA, "tensor map encode failed (sk down d%d)", d);
          xt_cur[d] = at<T>(plan.lay.T1[d]);
        }
        o.sp.stats_out = cur;
      }
      set_dbg(o, Beff * L, C); plan.ops.push_back(o);
    } else {
      const int Cin = cfg.channels[d - 1];
      Op o; o.depth = d; o.stack = 2; o.out_r = A; o.stats_out = cur; o.ck = "down";
      if (!add_gemm(o, W.down, down_in, f * Cin, L, Beff, nullptr, 0))
        return fail(SFB_ERR_CUDA, "tensor map encode failed (down d%d)", d);
      o.gp.gs = C / 8; set_dbg(o, Beff * L, C); plan.ops.push_back(o);
    }
    const int n = cfg.items[d];
    const bool has_inner = d + 1 < D;
    for (int i = 0; i < n; ++i) {
      const bool last = i == n - 1;
      int rc = build_item(d, 0, i, cur, /*want_t=*/last && has_inner, /*want_stats=*/!last || !has_inner);
      if (rc) return rc;
    }
    if (has_inner) {
      int rc = build_block(d + 1);   // leaves skip + s * up in bufA[d], stats in inner_out_stats
      if (rc) return rc;
      cur = inner_out_stats;
    }
    for (int i = 0; i < n; ++i) {
      const bool last = i == n - 1;
      int rc = build_item(d, 1, i, cur, /*want_t=*/last, /*want_stats=*/!last);
      if (rc) return rc;
    }
    if (d == 0) {
      Op o; o.kind = OP_D0_UP; o.depth = 0; o.stack = 2; o.L = L; o.C = C; o.B = Beff; o.in = item_t_out; o.ck = "up";
      o.w0 = W.c8_uw; o.fscalar = W.c8_ub; o.taps = W.up_taps; o.ft_off = W.skip_off; o.out_r = at<float>(plan.lay.veff);
      set_dbg(o, Beff, L); plan.ops.push_back(o);
    } else {
      const int Cin = cfg.channels[d - 1];
      float* Ap = at<float>(plan.lay.bufA[d - 1]);
      double* so = new_stats(Beff);
      if (W.rk_up_id >= 0) {
        Op o; o.depth = d; o.stack = 3; o.stats_out = so; o.ft_off = W.skip_off; o.ck = "up"; o.in = item_t_out;
        if (!add_rk(o, W.rk_up_id, item_t_out, W.rk_up_w, L, Beff) || !rk_out_r(o, Ap, true))
          return fail(SFB_ERR_CUDA, "tensor map encode failed (rk up d%d)", d);
        o.rp.stats_out = so; o.rp.bias = W.up.bias;
        o.flops = 2.0 * Beff * L * (double)W.up.N * W.up.taps * C; o.bytes = (double)Beff * L * (C * 2 + W.up.N * 8);
        set_dbg(o, Beff * Ld(plan.L, d - 1), Cin); plan.ops.push_back(o);
      } else if (sk_ok(W.up)) {
        Op o; o.depth = d; o.stack = 3; o.stats_out = so; o.ft_off = W.skip_off; o.ck = "up"; o.in = item_t_out;
        if constexpr (kBF16) {
          if (!add_sk(o, W.up, item_t_out, L, Beff, nullptr, 0, Cin / 8) || !sk_out_r(o, Ap, 1))
            return fail(SFB_ERR_CUDA, "tensor map encode failed (sk up d%d)", d);
          if (depth_sk(d - 1)) {      // the outer depth's up-stack reads the bf16 copy of x
            if (!sk_out_t(o, at<T>(plan.lay.T1[d - 1]))) return fail(SFB_ERR_CUDA, "tensor map encode failed (sk up d%d)", d);
            xt_cur[d - 1] = at<T>(plan.lay.T1[d - 1]);
          }
          o.sp.stats_out = so;
        }
        set_dbg(o, Beff * Ld(plan.L, d - 1), Cin); plan.ops.push_back(o);
      } else {
        Op o; o.depth = d; o.stack = 3; o.resid = Ap; o.out_r = Ap; o.stats_out = so; o.ft_off = W.skip_off; o.ck = "up";
        if (!add_gemm(o, W.up, item_t_out, C, L, Beff, nullptr, 0))
          return fail(SFB_ERR_CUDA, "tensor map encode failed (up d%d)", d);
        o.gp.gs = Cin / 8; set_dbg(o, Beff * Ld(plan.L, d - 1), Cin); plan.ops.push_back(o);
      }
      inner_out_stats = so;
    }
    return SFB_OK;
  }
  double* inner_out_stats = nullptr;
  long long* tl_buf = nullptr;   // sk timeline (debug): 8 roles x 256 stamps
  // op_index >= 0: attach the timeline buffer to that plan op (must be an sk op) ; host_buf != null: read it back
  int sk_timeline(int op_index, long long* host_buf, int n) override {
    if (!tl_buf) { if (cudaMalloc(&tl_buf, 8 * 256 * 8) != cudaSuccess) return fail(SFB_ERR_CUDA, "timeline alloc"); owned.push_back(tl_buf); }
    if (host_buf) {
      SFB_CUDA(cudaDeviceSynchronize());
      SFB_CUDA(cudaMemcpy(host_buf, tl_buf, sizeof(long long) * std::min(n, 8 * 256), cudaMemcpyDeviceToHost));
    }
    for (Op& o : plan.ops) o.sp.dbg = nullptr;
    if (op_index >= 0) {
      if (op_index >= (int)plan.ops.size() || plan.ops[op_index].kind != OP_SK) return fail(SFB_ERR_INVALID, "op %d is not an sk op", op_index);
      SFB_CUDA(cudaMemset(tl_buf, 0, 8 * 256 * 8));
      plan.ops[op_index].sp.dbg = tl_buf;
    }
    return SFB_OK;
  }
  std::vector<cudaEvent_t> prof_ev;   // 2 per op: events around every launch of the most recent U-Net evaluation

  // One line per plan op: "index kind depth stack item ms flops bytes" (algorithmic flops / bytes of the op).
  int profile_report(char* buf, int len) override {
    if (prof_ev.size() < 2 * plan.ops.size()) return fail(SFB_ERR_STATE, "no profiled evaluation recorded");
    std::string out;
    char line[160];
    for (size_t i = 0; i < plan.ops.size(); ++i) {
      const Op& o = plan.ops[i];
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, prof_ev[2 * i], prof_ev[2 * i + 1]) != cudaSuccess) ms = -1.f;
      double flops = 0, bytes = 0;
      const double rows = (double)o.B * o.L;
      if (o.kind == OP_GEMM) {
        const double K = (double)o.gp.taps * o.gp.K1 + (o.gp.k2_chunks ? (double)(o.k2) : 0.0);
        flops = 2.0 * rows * o.gp.N * K;
        bytes = rows * K / o.gp.taps * sizeof(T) + rows * o.gp.N * ((o.gp.resid ? 4 : 0) + (o.gp.out_r ? 4 : 0) + (o.gp.out_t ? sizeof(T) : 0)) +
                (double)o.gp.taps * o.gp.N * (K / o.gp.taps) * sizeof(T);
      } else if (o.kind == OP_RK || o.kind == OP_SK) {
        flops = o.flops;
        bytes = o.bytes;
      } else if (o.kind == OP_ATTN) {
        flops = 4.0 * (double)o.B * 8 * (double)o.L * o.L * 64;
        bytes = rows * (1536 + 512) * sizeof(T);
      } else if (o.kind == OP_GN) {
        bytes = rows * o.C * ((o.in_is_f32 ? 4 : sizeof(T)) + sizeof(T));
      } else if (o.kind == OP_LN) {
        bytes = rows * o.C * (4 + sizeof(T) + (o.out_r ? 4 : 0));
      } else if (o.kind == OP_CONV_C8) {
        flops = 2.0 * rows * 8 * 24;
        bytes = rows * 8 * (sizeof(T) + (o.resid ? 4 : 0) + (o.out_r ? 4 : 0) + (o.out_t ? sizeof(T) : 0));
      } else if (o.kind == OP_INJ_C8) {
        flops = 2.0 * rows * 8 * (8 + o.ctx);
        bytes = rows * (8 * (sizeof(T) + 4) + o.ctx * sizeof(T) + 8 * ((o.out_r ? 4 : 0) + (o.out_t ? sizeof(T) : 0)));
      } else if (o.kind == OP_D0_CONV1) {
        flops = 2.0 * rows * 8 * 24;
        bytes = rows * 8 * (4 + sizeof(T));
      } else if (o.kind == OP_D0_TAIL) {
        flops = 2.0 * rows * 8 * (24 + 8 + o.ctx);
        bytes = rows * (8 * (sizeof(T) + 4 + 4 + (o.out_t ? sizeof(T) : 0)) + o.ctx * sizeof(T));
      } else if (o.kind == OP_XOUT_C8) {
        flops = 2.0 * rows * 8 * 512;
        bytes = rows * (512 * sizeof(T) + 8 * (4 + 4 + (o.out_t ? sizeof(T) : 0)));
      } else if (o.kind == OP_D0_DOWN) {
        flops = 2.0 * rows * 8;
        bytes = rows * (4 + 32);
      } else if (o.kind == OP_D0_UP) {
        flops = 2.0 * rows * 8 * o.taps;
        bytes = rows * (8 * sizeof(T) + 8);
      }
      snprintf(line, sizeof line, "%zu %s %d %d %d %.6f %.6e %.6e\n", i, kOpNames[o.kind], o.depth, o.stack, o.item, ms, flops, bytes);
      out += line;
    }
    if ((int)out.size() + 1 > len) return fail(SFB_ERR_INVALID, "profile buffer too small: need %zu", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return SFB_OK;
  }

  // Names the op and the barrier of a wait-log record (barrier arrays as laid out in sk_tc.cuh / rk_tc.cuh / attn_tc.cuh).
  std::string describe_wait(const WaitRecord& r) override {
    if (r.tag >= plan.ops.size()) return "";
    const Op& o = plan.ops[r.tag];
    char b[256];
    snprintf(b, sizeof b, "| %s d%d s%d i%d '%s' B=%d L=%d C=%d", kOpNames[o.kind], o.depth, o.stack, o.item, o.ck, o.B, o.L, o.C);
    std::string out = b;
    const uint32_t addr = r.bar & 0x7FFFFFFFu;
    if (addr == 0) return out;       // light record: no barrier address
    auto name_of = [&](uint32_t bars_off, const char* const* names, const int* counts, int n) {
      const uint32_t idx = ((addr - bars_off) & 1023u) / 8;     // dynamic shared memory starts 1024-aligned
      uint32_t k = idx;
      for (int j = 0; j < n; ++j) {
        if (k < (uint32_t)counts[j]) { snprintf(b, sizeof b, " bar=%s[%u]", names[j], k); out += b; return; }
        k -= counts[j];
      }
      snprintf(b, sizeof b, " bar=#%u", idx); out += b;
    };
    if (o.kind == OP_SK) {
      const SkParams& q = o.sp;
      const uint32_t fixed = o.BN == 256 ? SkCfg<256>::fixed_bytes(q.epi12) - SkCfg<256>::BAR_BYTES : SkCfg<128>::fixed_bytes(q.epi12) - SkCfg<128>::BAR_BYTES;
      const uint32_t off = q.na * SkCfg<128>::A_BYTES + q.nb * o.BN * 128 + q.nr * SkCfg<128>::R_BYTES + fixed;
      static const char* const names[] = {"a_full", "a_empty", "op_full", "b_full", "b_empty", "acc_full", "acc_empty", "rc_full", "rc_empty"};
      static const int counts[] = {4, 4, 4, 6, 6, 2, 2, 6, 6};
      snprintf(b, sizeof b, " BN=%d taps=%d xf=%d epi12=%d na/nb/nr=%d/%d/%d resid=%d k1c=%d k2c=%d tiles=%d", o.BN, q.taps, q.xf, q.epi12, q.na, q.nb,
               q.nr, q.resid_mode, q.k1_chunks, q.k2_chunks, q.total_tiles);
      out += b;
      name_of(off, names, counts, 9);
    } else if (o.kind == OP_RK) {
      uint32_t off = 0;
      int idn = 0;
#define X(a, b_, c, d, e_, f, g) if (o.rk_id == idn++) off = RkCfg<a, b_, c, d, e_, f, g>::OFF_BAR;
      SFB_RK_LIST(X)
#undef X
      static const char* const names[] = {"w_full", "a_full", "a_empty", "op_full", "acc1_full", "acc1_empty", "a2_full", "acc2_full", "r_full", "r_empty"};
      static const int counts[] = {1, 4, 4, 4, 2, 2, 1, 1, 5, 5};
      snprintf(b, sizeof b, " rk_id=%d tiles=%d", o.rk_id, o.rp.total_tiles); out += b;
      name_of(off, names, counts, 10);
    } else if (o.kind == OP_ATTN) {
      static const char* const names[] = {"q_full", "kv_full", "kv_empty", "s_full", "p_ready", "o_full", "s_free"};
      static const int counts[] = {1, 4, 4, 1, 1, 1, 1};
      name_of((uint32_t)(sizeof(T) == 2 ? attn2_smem_bytes() : attn_smem_bytes<T>()) - 256, names, counts, 7);   // both forms lay their barriers out alike
    }
    return out;
  }

  int ensure_plan(int64_t B, int64_t L, int cfg_on, int64_t rows, void* ws, size_t ws_bytes, int64_t M = 1) {
    if (!finalized) return fail(SFB_ERR_STATE, "finalize first");
    size_t need = 0;
    int rc = workspace_bytes(B, L, cfg_on, rows, M, &need);
    if (rc) return rc;
    if (!ws || ws_bytes < need) return fail(SFB_ERR_INVALID, "workspace too small: %zu < %zu", ws_bytes, need);
    uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(ws), 1024));
    WsLayout lay;
    layout(B, L, cfg_on, rows, M, lay);
    if (plan.ws == base && plan.B == B && plan.L == L && plan.cfg_on == cfg_on && plan.M == M && plan.lay.total == lay.total &&
        !plan.ops.empty())
      return SFB_OK;
    plan = Plan();
    ++plan_gen;
    plan.B = B; plan.L = L; plan.cfg_on = cfg_on; plan.M = M; plan.ws = base; plan.lay = lay;
    wsb = base;
    stats_next = 0;
    fold_items.clear(); fold_next = 0; fold_rows = 0;
    rc = build_block(0);
    if (rc) { plan.ops.clear(); return rc; }
    for (size_t i = 0; i < plan.ops.size(); ++i) {      // plan index of every op, reported by the device wait log (ptx.cuh)
      Op& o = plan.ops[i];
      o.sp.tag = o.rp.tag = o.ap.tag = o.gp.tag = (int)i;
    }
    if constexpr (kBF16) {      // shared-memory ring depths / epilogue width of every streaming-K op (sk_tc.cuh)
      const bool no_epi12 = getenv("SFB_NO_EPI12") != nullptr;
      for (Op& o : plan.ops) {
        if (o.kind != OP_SK) continue;
        SkParams& q = o.sp;
        q.epi12 = (!no_epi12 && q.xf == 0 && q.taps == 1 && q.rowstats_out == nullptr) ? 1 : 0;   // the small-K, epilogue-bound ops
        const bool uses_r = q.resid_mode != 0 || q.has_out_r != 0;
        if (o.BN == 256) sk_pick_rings<256>(q.taps, q.xf, uses_r, q.resid_mode != 0, q.epi12, q.na, q.nb, q.nr);
        else sk_pick_rings<128>(q.taps, q.xf, uses_r, q.resid_mode != 0, q.epi12, q.na, q.nb, q.nr);
      }
    }
    if (!fold_items.empty()) {
      // A FRESH device table per plan: kernels of the previous plan, enqueued asynchronously on the caller's stream by an
      // earlier sample(), may still be reading the old one (a blocking copy on the legacy stream does not order against a
      // non-blocking stream).  Plans are rebuilt only when (B, L, CFG, M, workspace) change; the tables are a few KB.
      void* d = nullptr;
      if (cudaMalloc(&d, fold_items.size() * sizeof(FoldItem)) != cudaSuccess) { plan.ops.clear(); return fail(SFB_ERR_CUDA, "fold table alloc"); }
      owned.push_back(d);
      fold_items_dev = reinterpret_cast<FoldItem*>(d); fold_items_cap = fold_items.size();
      if (cudaMemcpy(fold_items_dev, fold_items.data(), fold_items.size() * sizeof(FoldItem), cudaMemcpyHostToDevice) != cudaSuccess) {
        plan.ops.clear();
        return fail(SFB_ERR_CUDA, "fold table upload");
      }
    }
    return SFB_OK;
  }
  int plan_size(int64_t B, int64_t L, int cfg_on, void* ws, size_t ws_bytes, int64_t M = 1) override {
    int rc = ensure_plan(B, L, cfg_on, B, ws, ws_bytes, M);
    if (rc) return rc;
    return (int)plan.ops.size();
  }
  int op_info(int i, char* buf, int len) override {
    if (i < 0 || i >= (int)plan.ops.size()) return fail(SFB_ERR_INVALID, "op index out of range");
    const Op& o = plan.ops[i];
    snprintf(buf, len, "%s %d %d %d %zu %zu %d %d %d %s", kOpNames[o.kind], o.depth, o.stack, o.item, o.dbg_off, o.dbg_bytes,
             o.dbg_rows, o.dbg_cols, o.dbg_dtype, o.ck[0] ? o.ck : "-");
    return SFB_OK;
  }

  // ---------------------------------------------------------------- execution
  struct StepCtx {
    const float* frow;   // feature-table row(s)
    int bstride, bmod;   // per-clip row stride (0: shared) and modulo
    const float* x;      // [B, L] net input
  };

  template <int NV>
  void launch_ln(const Op& o, const StepCtx& sc, cudaStream_t st) {
    const size_t rows = (size_t)o.B * o.L;
    const int lpr = (o.C / 8 < 32) ? o.C / 8 : 32, rpw = 32 / lpr;
    const size_t warps = (rows + rpw - 1) / rpw;
    const unsigned blocks = (unsigned)((warps + 7) / 8);
    const float* scale = o.ft_off >= 0 ? sc.frow + o.ft_off : nullptr;
    const float* shift = o.ft_off >= 0 ? sc.frow + o.ft_off + o.C : nullptr;
    launch_pdl(ln_mod_kernel<T, NV>, blocks, 256, 0, st, reinterpret_cast<const float*>(o.in), scale, shift, sc.bstride, sc.bmod,
                                                reinterpret_cast<T*>(o.out_t), o.out_r, rows, o.L, o.C, 1e-5f);
  }

  int run_unet(const StepCtx& sc, cudaStream_t st) {
    const int Bx = (int)plan.B;
    SFB_CUDA(cudaMemsetAsync(at<double>(plan.lay.stats), 0, plan.lay.stats_bytes, st));
    ++launches;
    if (!fold_items.empty()) {     // per-evaluation scaled inject weights (LayerNorm fold): one warp per (copy, row)
      const size_t warps = (size_t)fold_rows * sc.bmod;
      launch_pdl(inject_fold_kernel, (unsigned)((warps + 7) / 8), 256, 0, st, fold_items_dev, (int)fold_items.size(), fold_rows, sc.frow,
                                                                     sc.bstride, sc.bmod);
      ++launches;
    }
    int n = (int)plan.ops.size();
    if (op_limit >= 0 && op_limit < n) n = op_limit;
    if (profiling && prof_ev.size() < 2 * plan.ops.size()) {
      const size_t old = prof_ev.size();
      prof_ev.resize(2 * plan.ops.size());
      for (size_t k = old; k < prof_ev.size(); ++k) SFB_CUDA(cudaEventCreate(&prof_ev[k]));
    }
    for (int i = 0; i < n; ++i) {
      const Op& o = plan.ops[i];
      const unsigned lb = (unsigned)((o.L + 255) / 256);
      if (profiling) cudaEventRecord(prof_ev[2 * i], st);
      switch (o.kind) {
        case OP_D0_DOWN:
          launch_pdl(d0_down_kernel, dim3(lb, o.B), 256, 0, st, sc.x, o.w0, o.w1, o.out_r, o.stats_out, o.L, Bx);
          break;
        case OP_GN: {
          const size_t nvec = (size_t)o.L * o.C / 8;
          unsigned bpc = (unsigned)std::min<size_t>(std::max<size_t>((nvec + 1023) / 1024, 1), 8192);
          const size_t sm = (size_t)2 * o.C * sizeof(float);
          if (o.in_is_f32)
            launch_pdl(gn_apply_silu_kernel<float, T>, dim3(bpc, o.B), 256, sm, st, reinterpret_cast<const float*>(o.in), o.stats_in, o.w0, o.w1,
                                                                          reinterpret_cast<T*>(o.out_t), o.L, o.C, o.gs, 1e-5f);
          else
            launch_pdl(gn_apply_silu_kernel<T, T>, dim3(bpc, o.B), 256, sm, st, reinterpret_cast<const T*>(o.in), o.stats_in, o.w0, o.w1,
                                                                      reinterpret_cast<T*>(o.out_t), o.L, o.C, o.gs, 1e-5f);
          break;
        }
        case OP_CONV_C8:
          launch_pdl(conv3_c8_kernel<T>, dim3(lb, o.B), 256, 0, st, reinterpret_cast<const T*>(o.in), o.w0, o.w1, o.resid, o.out_r,
                                                            reinterpret_cast<T*>(o.out_t), o.stats_out, o.L);
          break;
        case OP_INJ_C8:
          launch_pdl(inject_c8_kernel<T, 2>, dim3(lb, o.B), 256, 0, st, reinterpret_cast<const T*>(o.in), o.resid,
                                                               reinterpret_cast<const T*>(o.in2), o.w0, o.w1, o.w2, o.out_r,
                                                               reinterpret_cast<T*>(o.out_t), o.stats_out, o.L, Bx, XB_total);
          break;
        case OP_D0_CONV1:
          if constexpr (sizeof(T) == 2) {
            if (d0_tc) {
              launch_pdl(d0_gn_conv1_tc_kernel, dim3((unsigned)((o.L + kD0TcTile - 1) / kD0TcTile), o.B), 256, 0, st, reinterpret_cast<const float*>(o.in),
                         (const double*)o.stats_in, o.w0, o.w1, o.wx[0], o.wx[1], reinterpret_cast<__nv_bfloat16*>(o.out_t), o.stats_out, o.L, 1e-5f, (int)i);
              break;
            }
          }
          launch_pdl(d0_gn_conv1_kernel<T>, dim3((unsigned)((o.L + d0_positions_per_block() - 1) / d0_positions_per_block()), o.B), 256, 0, st, reinterpret_cast<const float*>(o.in),
                     (const double*)o.stats_in, o.w0, o.w1, o.wx[0], o.wx[1], reinterpret_cast<T*>(o.out_t), o.stats_out, o.L, 1e-5f);
          break;
        case OP_D0_TAIL:
          if constexpr (sizeof(T) == 2) {
            if (d0_tc) {
              launch_pdl(d0_tail_tc_kernel<2>, dim3((unsigned)((o.L + kD0TcTile - 1) / kD0TcTile), o.B), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(o.in),
                         (const double*)o.stats_in, o.w0, o.w1, o.wx[0], o.wx[1], o.resid, sc.frow + o.ft_off, sc.bstride, sc.bmod,
                         reinterpret_cast<const __nv_bfloat16*>(o.in2), Bx, o.wx[2], o.wx[3], o.w2, XB_total, o.out_r, reinterpret_cast<__nv_bfloat16*>(o.out_t),
                         o.stats_out, o.L, 1e-5f, (int)i);
              break;
            }
          }
          launch_pdl(d0_tail_kernel<T, 2>, dim3((unsigned)((o.L + d0_positions_per_block() - 1) / d0_positions_per_block()), o.B), 256, 0, st, reinterpret_cast<const T*>(o.in),
                     (const double*)o.stats_in, o.w0, o.w1, o.wx[0], o.wx[1], o.resid, sc.frow + o.ft_off, sc.bstride, sc.bmod,
                     reinterpret_cast<const T*>(o.in2), Bx, o.wx[2], o.wx[3], o.w2, XB_total, o.out_r, reinterpret_cast<T*>(o.out_t),
                     o.stats_out, o.L, 1e-5f);
          break;
        case OP_D0_UP:
          launch_pdl(d0_up_kernel<T>, dim3((unsigned)((o.L + 1023) / 1024), o.B), 256, 0, st, reinterpret_cast<const T*>(o.in), o.w0, o.fscalar, sc.frow + o.ft_off,
                                                        sc.bstride, sc.bmod, sc.x, o.out_r, o.L, Bx, o.taps);
          break;
        case OP_GEMM: {
          if (o.ft_off >= 0) {
            GemmParams<T> p = o.gp;
            p.colscale = sc.frow + o.ft_off;
            p.cs_bstride = sc.bstride;
            p.cs_bmod = sc.bmod;
            launch_gemm<T>(p, o.BN, o.B, st);
          } else {
            launch_gemm<T>(o.gp, o.BN, o.B, st);
          }
          break;
        }
        case OP_RK: {
          if constexpr (kBF16) {
            RkParams p = o.rp;
            if (o.ft_off >= 0) {
              if (rk_key(o.rk_id).EPI >= 1) { p.mod = sc.frow + o.ft_off; p.mod_bstride = sc.bstride; p.mod_bmod = sc.bmod; }
              else { p.colscale = sc.frow + o.ft_off; p.cs_bstride = sc.bstride; p.cs_bmod = sc.bmod; }
            }
            rk_launch(o.rk_id, p, num_sms(), st);
          }
          break;
        }
        case OP_SK: {
          if constexpr (kBF16) {
            SkParams p = o.sp;
            if (o.ft_off >= 0) {
              if (o.sk_ft_is_mod) { p.mod = sc.frow + o.ft_off; p.mod_bstride = sc.bstride; p.mod_bmod = sc.bmod; p.w_bmod = sc.bmod; }
              else { p.colscale = sc.frow + o.ft_off; p.cs_bstride = sc.bstride; p.cs_bmod = sc.bmod; }
            }
            const int epi = sk_epi_of(p.ln_fold, p.resid_mode, p.colscale != nullptr);
            sk_launch(o.sk_id, epi, p, num_sms(), st);
          }
          break;
        }
        case OP_LN: {
          const int lpr = (o.C / 8 < 32) ? o.C / 8 : 32;
          const int nv = o.C / (8 * lpr);
          if (nv == 1) launch_ln<1>(o, sc, st);
          else if (nv == 2) launch_ln<2>(o, sc, st);
          else launch_ln<4>(o, sc, st);
          break;
        }
        case OP_XOUT_C8:
          launch_pdl(xattn_out_c8_kernel<T>, dim3(lb, o.B), 256, 0, st, reinterpret_cast<const T*>(o.in), o.w0, o.resid, o.out_r,
                     reinterpret_cast<T*>(o.out_t), o.stats_out, o.L);
          break;
        case OP_ATTN: {
          dim3 grid((o.L + 127) / 128, 8, o.B);
          attn_launch<T>(grid, st, o.ap);
          break;
        }
      }
      ++launches;
      if (profiling) cudaEventRecord(prof_ev[2 * i + 1], st);
    }
    SFB_CUDA(cudaGetLastError());
    return SFB_OK;
  }

  // loop-invariant conditioning: time features -> modulation / skip tables; cross-attention biases; onset pyramid layout
  int prepare(const float* sigma_dev, int rows, const float* const* channels, int n_channels, const float* embedding,
              int64_t M, cudaStream_t st) {
    const int64_t B = plan.B, L = plan.L;
    const int Beff = plan.cfg_on ? 2 * (int)B : (int)B;
    const int MF = cfg.modulation_features, EF = cfg.embedding_features;
    if (M != plan.M) return fail(SFB_ERR_STATE, "plan was built for M=%lld", (long long)plan.M);
    if (n_channels < cfg.depth) return fail(SFB_ERR_INVALID, "channels: need %d context tensors, got %d", cfg.depth, n_channels);
    if (!embedding) return fail(SFB_ERR_INVALID, "embedding is required (ClassifierFreeGuidancePlugin)");
    const WsLayout& w = plan.lay;
    auto lin = [&](const float* in, const float* W, const float* bias, float* out, int r, int K, int J, int ldo, int ai, int ao) {
      const unsigned blocks = (unsigned)((J + 7) / 8);
      linear_act_kernel<<<blocks, 256, 0, st>>>(in, W, bias, out, r, K, J, K, ldo, ai, ao);
      ++launches;
    };
    // time features (A.3)
    fourier_embed_kernel<<<rows, 128, 0, st>>>(sigma_dev, t_w, at<float>(w.fourier), rows, 128);
    ++launches;
    lin(at<float>(w.fourier), t_lw, t_lb, at<float>(w.h1), rows, 257, MF, MF, ACT_NONE, ACT_GELU);
    lin(at<float>(w.h1), t_mw, t_mb, at<float>(w.h2), rows, MF, MF, MF, ACT_NONE, ACT_GELU);
    lin(at<float>(w.h2), t_mw, t_mb, at<float>(w.feat), rows, MF, MF, MF, ACT_NONE, ACT_GELU);
    lin(at<float>(w.feat), ft_w, ft_b, at<float>(w.ftable), rows, MF, F_total, F_total, ACT_SILU, ACT_NONE);
    // cross-attention biases (A.4 + A.7, M_ctx = 1)
    // rows [0, B*M) = the caller's embedding tokens, rows [B*M, 2*B*M) = FixedEmbedding tokens 0..M-1 (CFG mask branch)
    build_emb_rows_kernel<<<(unsigned)(Beff * M), 128, 0, st>>>(embedding, fixed_emb, at<float>(w.embrows), (int)(B * M), (int)M, EF);
    ++launches;
    if (M > 1) {
      // general cross-attention: k | v = to_kv(LN_ctx(e)) per item, [Beff, M, 1024] in operand precision (loop invariant)
      const int R2 = (int)(Beff * M);
      for (int d = 0; d < cfg.depth; ++d)
        for (int s = 0; s < 2; ++s)
          for (const ItemW& I : dw[d].items[s]) {
            if (!I.has_xattn) continue;
            if (I.xkv_index < 0) return fail(SFB_ERR_STATE, "cross-attention weights for M > 1 were not built (embedding_max_length = %d)", cfg.embedding_max_length);
            ln_rows_kernel<<<(R2 + 7) / 8, 256, 0, st>>>(at<float>(w.embrows), I.x_ng, I.x_nb, at<float>(w.tmp1), R2, EF, 1e-5f);
            ++launches;
            lin(at<float>(w.tmp1), I.x_wkv, nullptr, at<float>(w.tmp2), R2, EF, 1024, 1024, ACT_NONE, ACT_NONE);
            const size_t n = (size_t)R2 * 1024;
            f32_to_operand_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(at<float>(w.tmp2), at<T>(w.xkv) + (size_t)I.xkv_index * n, n);
            ++launches;
          }
    }
    for (int d = 0; d < cfg.depth && M == 1; ++d)
      for (int s = 0; s < 2; ++s)
        for (const ItemW& I : dw[d].items[s]) {
          if (!I.has_xattn) continue;
          ln_rows_kernel<<<(Beff + 7) / 8, 256, 0, st>>>(at<float>(w.embrows), I.x_ng, I.x_nb, at<float>(w.tmp1), Beff, EF, 1e-5f);
          ++launches;
          lin(at<float>(w.tmp1), I.x_wv, nullptr, at<float>(w.tmp2), Beff, EF, 512, 512, ACT_NONE, ACT_NONE);
          lin(at<float>(w.tmp2), I.x_wo, nullptr, at<float>(w.xbias) + I.xb_off, Beff, 512, cfg.channels[d], XB_total, ACT_NONE, ACT_NONE);
        }
    // onset pyramid NCL f32 -> NLC operand precision
    for (int d = 0; d < cfg.depth; ++d) {
      const int ld = Ld(L, d);
      if (!channels[d]) return fail(SFB_ERR_INVALID, "channels[%d] is null", d);
      ncl_to_nlc_kernel<T><<<dim3((ld + 255) / 256, (unsigned)B), 256, 0, st>>>(channels[d], at<T>(w.ctx[d]), cfg.context_channels[d], ld);
      ++launches;
    }
    SFB_CUDA(cudaGetLastError());
    return SFB_OK;
  }

  int unet_forward(const float* x, const float* sigma, const float* const* channels, int n_channels, const float* embedding,
                   int64_t M, float scale, float* v_out, int64_t B, int64_t L, void* ws, size_t ws_bytes,
                   cudaStream_t st) override {
    const int cfg_on = scale != 1.0f;
    int rc = ensure_plan(B, L, cfg_on, B, ws, ws_bytes, M);
    if (rc) return rc;
    wsb = reinterpret_cast<uint8_t*>(plan.ws);
    launches = 0;
    rc = prepare(sigma, (int)B, channels, n_channels, embedding, M, st);
    if (rc) return rc;
    StepCtx sc{at<float>(plan.lay.ftable), F_total, (int)B, x};
    rc = run_unet(sc, st);
    if (rc) return rc;
    const size_t n = (size_t)B * L;
    cfg_combine_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(at<float>(plan.lay.veff), v_out, n, cfg_on, scale);
    ++launches;
    SFB_CUDA(cudaGetLastError());
    return SFB_OK;
  }

  int sample(const float* x_noisy, int num_steps, const float* const* channels, int n_channels, const float* embedding,
             int64_t M, float scale, float* x_out, float* traj_x, float* traj_v, const float* teacher_x, int64_t B, int64_t L,
             void* ws, size_t ws_bytes, cudaStream_t st) override {
    if (num_steps <= 0) return fail(SFB_ERR_INVALID, "num_steps must be positive");
    const int cfg_on = scale != 1.0f;
    int rc = ensure_plan(B, L, cfg_on, num_steps + 1, ws, ws_bytes, M);
    if (rc) return rc;
    wsb = reinterpret_cast<uint8_t*>(plan.ws);
    launches = 0;
    // LinearSchedule (A.2): torch.linspace(1, 0, N + 1) in fp32, same two-sided formula as ATen.
    std::vector<float> sig(num_steps + 1);   // host copy for the per-step alpha / beta scalars
    {
      const int steps = num_steps + 1;
      const float start = 1.f, end = 0.f;
      const float step = (end - start) / (float)(steps - 1);
      const int half = steps / 2;
      for (int i = 0; i < steps; ++i) sig[i] = i < half ? start + step * (float)i : end - step * (float)(steps - 1 - i);
    }
    const size_t n = (size_t)B * L;
    float* xs = at<float>(plan.lay.xstate);
    bool graphable = use_graph && !profiling && !traj_x && !traj_v && !teacher_x;
    if (graphable) {       // a caller that is itself capturing `st` gets the plain launches recorded into ITS graph
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); graphable = false; }
    }
    // Whole-loop CUDA graph (SFB_GRAPH=0 turns it off): the per-call preparation (tables, cross-attention
    // biases, onset pyramid: the only kernels that read the caller's buffers) runs eagerly, the num_steps x 169 launches
    // of the sampling loop - which touch the workspace only - are captured once per (plan, steps, scale) as plain kernel
    // nodes and replayed with ONE launch: +4 % at the bench shape (the stream's front end no
    // longer meters 8.5 k launches).  Capture and replay run on an internal stream (the caller's may be the legacy default
    // stream, which cannot be captured); events order it after the caller's earlier work and the caller's later work after it.
    cudaStream_t ws_st = st;
    if (graphable) {
      if (!gs) {
        SFB_CUDA(cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking));
        SFB_CUDA(cudaEventCreateWithFlags(&ge0, cudaEventDisableTiming));
        SFB_CUDA(cudaEventCreateWithFlags(&ge1, cudaEventDisableTiming));
      }
      SFB_CUDA(cudaEventRecord(ge0, st));
      SFB_CUDA(cudaStreamWaitEvent(gs, ge0, 0));
      ws_st = gs;
    }
    sigma_linspace_kernel<<<(num_steps + 256) / 256, 256, 0, ws_st>>>(at<float>(plan.lay.sigma), num_steps + 1);
    ++launches;
    rc = prepare(at<float>(plan.lay.sigma), num_steps + 1, channels, n_channels, embedding, M, ws_st);
    if (rc) return rc;
    SFB_CUDA(cudaMemcpyAsync(xs, x_noisy, n * 4, cudaMemcpyDeviceToDevice, ws_st));
    auto loop = [&](cudaStream_t s2) -> int {
      const float hp = 1.5707963267948966f;
      for (int i = 0; i < num_steps; ++i) {
        const float* xe = teacher_x ? teacher_x + (size_t)i * n : xs;
        StepCtx sc{at<float>(plan.lay.ftable) + (size_t)i * F_total, 0, 1, xe};
        const int r = run_unet(sc, s2);
        if (r) return r;
        const float a = cosf(sig[i] * hp), b = sinf(sig[i] * hp), a2 = cosf(sig[i + 1] * hp), b2 = sinf(sig[i + 1] * hp);
        launch_pdl(sampler_update_kernel, (unsigned)((n / 4 + 255) / 256), 256, 0, s2,
            xe, at<float>(plan.lay.veff), xs, traj_x ? traj_x + (size_t)i * n : nullptr, traj_v ? traj_v + (size_t)i * n : nullptr,
            n, cfg_on, scale, a, b, a2, b2);
        ++launches;
      }
      return SFB_OK;
    };
    if (!graphable) {
      rc = loop(st);
      if (rc) return rc;
    } else {
      if (plan_gen != graph_gen) { for (GraphEntry& g : graphs) cudaGraphExecDestroy(g.exec); graphs.clear(); graph_gen = plan_gen; }
      uint32_t sbits;
      memcpy(&sbits, &scale, sizeof sbits);
      const uint64_t key = ((uint64_t)(uint32_t)num_steps << 32) | sbits;
      GraphEntry* hit = nullptr;
      for (GraphEntry& g : graphs) if (g.key == key) hit = &g;
      if (!hit) {
        const int64_t before = launches;
        SFB_CUDA(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
        g_pdl_suppress = getenv("SFB_GRAPH_PDL") == nullptr;      // plain kernel nodes (ptx.cuh); SFB_GRAPH_PDL=1 keeps the programmatic edges
        rc = loop(gs);
        g_pdl_suppress = false;
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(gs, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return fail(SFB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        GraphEntry e;
        e.key = key; e.launches = launches - before;
        const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return fail(SFB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ie));
        if (graphs.size() >= 4) { cudaGraphExecDestroy(graphs.front().exec); graphs.erase(graphs.begin()); }
        graphs.push_back(e);
        hit = &graphs.back();
      } else {
        launches += hit->launches;
      }
      SFB_CUDA(cudaGraphLaunch(hit->exec, gs));
    }
    SFB_CUDA(cudaMemcpyAsync(x_out, xs, n * 4, cudaMemcpyDeviceToDevice, ws_st));
    if (graphable) {
      SFB_CUDA(cudaEventRecord(ge1, gs));
      SFB_CUDA(cudaStreamWaitEvent(st, ge1, 0));
    }
    SFB_CUDA(cudaGetLastError());
    return SFB_OK;
  }
  struct GraphEntry { uint64_t key = 0; cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
  std::vector<GraphEntry> graphs;
  cudaStream_t gs = nullptr;
  cudaEvent_t ge0 = nullptr, ge1 = nullptr;
  bool use_graph = !(getenv("SFB_GRAPH") != nullptr && atoi(getenv("SFB_GRAPH")) == 0);
  uint64_t plan_gen = 0, graph_gen = 0;
};

}  // namespace

// ------------------------------------------------------------------------------------------------ post-processing (f-2)
// torchaudio.functional.resample's tap table, in float32 and in torchaudio's operation order (_get_sinc_resample_kernel
// builds it in the waveform's dtype): the reference's taps carry float32 rounding that a drop-in has to reproduce.
struct ResampleTable { int orig = 0, nw = 0, K = 0, width = 0; float* dev = nullptr; };
static std::mutex g_post_mu;
static std::map<std::pair<int, std::pair<int, int>>, ResampleTable> g_post_tables;   // (device, (orig_freq, new_freq))
static const ResampleTable* resample_table(int device, int orig_freq, int new_freq) {
  std::lock_guard<std::mutex> lk(g_post_mu);
  auto key = std::make_pair(device, std::make_pair(orig_freq, new_freq));
  auto it = g_post_tables.find(key);
  if (it != g_post_tables.end()) return &it->second;
  ResampleTable t;
  const int g = std::gcd(orig_freq, new_freq);
  t.orig = orig_freq / g; t.nw = new_freq / g;
  const int lpw = 6;
  const double base_d = std::min(t.orig, t.nw) * 0.99;
  t.width = (int)std::ceil(lpw * t.orig / base_d);
  t.K = 2 * t.width + t.orig;
  const float base = (float)base_d, pi = (float)M_PI, scale = (float)(base_d / t.orig);
  std::vector<float> h((size_t)t.nw * t.K);
  for (int ph = 0; ph < t.nw; ++ph)
    for (int k = 0; k < t.K; ++k) {
      const float idx = (float)(k - t.width) / (float)t.orig;
      volatile float tt = (float)(-ph) / (float)t.nw + idx;     // volatile: one float32 rounding per operation, no contraction
      tt = tt * base;
      tt = std::fmin(std::fmax((float)tt, (float)-lpw), (float)lpw);
      volatile float w = (float)tt * pi;
      w = w / (float)lpw; w = w / 2.f;
      w = std::cos((float)w); w = w * w;
      tt = tt * pi;
      volatile float kv = ((float)tt == 0.f) ? 1.f : std::sin((float)tt) / (float)tt;
      volatile float ws = w * scale;
      h[(size_t)ph * t.K + k] = kv * ws;
    }
  if (cudaMalloc(&t.dev, h.size() * sizeof(float)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(t.dev, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  return &(g_post_tables[key] = t);
}

// ------------------------------------------------------------------------------------------------ onset encoder (f-1)
struct OnsetEncoder {
  sfb_encoder_config cfg;
  int device = 0;
  std::string err;
  std::map<std::string, HostTensor> params;
  std::vector<void*> owned;
  bool finalized = false;
  struct Block { float *gn1_w, *gn1_b, *c1_w, *c1_b, *gn2_w, *gn2_b, *c2_w, *c2_b, *sc_w, *sc_b; int cin, cout, g1, g2; };
  struct Level { float *dw, *db; int cin, cout, f; std::vector<Block> blocks; };
  Block to_in{};
  std::vector<Level> levels;

  ~OnsetEncoder() { for (void* p : owned) cudaFree(p); }
  int fail(int code, const char* fmt, ...) {
    char b[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(b, sizeof b, fmt, ap);
    va_end(ap);
    err = b;
    return code;
  }
  float* up(const std::string& name, std::initializer_list<int64_t> shape) {
    auto it = params.find(name);
    if (it == params.end()) { fail(SFB_ERR_MISSING, "missing encoder parameter '%s'", name.c_str()); return nullptr; }
    if (it->second.shape != std::vector<int64_t>(shape)) { fail(SFB_ERR_INVALID, "encoder parameter '%s' has the wrong shape", name.c_str()); return nullptr; }
    void* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(it->second.v.size() * 4, 16)) != cudaSuccess) { fail(SFB_ERR_CUDA, "cudaMalloc"); return nullptr; }
    owned.push_back(d);
    if (cudaMemcpy(d, it->second.v.data(), it->second.v.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { fail(SFB_ERR_CUDA, "cudaMemcpy"); return nullptr; }
    return reinterpret_cast<float*>(d);
  }
  bool load_block(const std::string& pre, int cin, int cout, int groups, Block& b) {
    b.cin = cin; b.cout = cout;
    b.g1 = (cin % groups == 0) ? groups : 1;      // ResnetBlock1d: block1 falls back to one group when cin is not divisible
    b.g2 = groups;
    b.gn1_w = up(pre + "block1.groupnorm.weight", {cin}); b.gn1_b = up(pre + "block1.groupnorm.bias", {cin});
    b.c1_w = up(pre + "block1.project.weight", {cout, cin, 3}); b.c1_b = up(pre + "block1.project.bias", {cout});
    b.gn2_w = up(pre + "block2.groupnorm.weight", {cout}); b.gn2_b = up(pre + "block2.groupnorm.bias", {cout});
    b.c2_w = up(pre + "block2.project.weight", {cout, cout, 3}); b.c2_b = up(pre + "block2.project.bias", {cout});
    b.sc_w = b.sc_b = nullptr;
    if (cin != cout) { b.sc_w = up(pre + "to_out.weight", {cout, cin, 1}); b.sc_b = up(pre + "to_out.bias", {cout}); }
    return b.gn1_w && b.gn1_b && b.c1_w && b.c1_b && b.gn2_w && b.gn2_b && b.c2_w && b.c2_b && (cin == cout || (b.sc_w && b.sc_b));
  }
  int finalize() {
    if (finalized) return SFB_OK;
    const sfb_encoder_config& c = cfg;
    if (c.patch_size != 1) return fail(SFB_ERR_UNSUPPORTED, "patch_size must be 1 (exp/model/diffusion.yaml:43)");
    if (c.n_levels < 1 || c.n_levels > SFB_MAX_DEPTH) return fail(SFB_ERR_INVALID, "n_levels out of range");
    if (!load_block("to_in.", c.in_channels, c.channels * c.multipliers[0], 1, to_in)) return err.empty() ? SFB_ERR_MISSING : SFB_ERR_MISSING;
    levels.resize(c.n_levels);
    for (int i = 0; i < c.n_levels; ++i) {
      Level& lv = levels[i];
      lv.cin = c.channels * c.multipliers[i]; lv.cout = c.channels * c.multipliers[i + 1]; lv.f = c.factors[i];
      if (lv.cout > kEncMaxC || lv.cin > kEncMaxC) return fail(SFB_ERR_UNSUPPORTED, "encoder channels > %d", kEncMaxC);
      if (lv.cout % c.resnet_groups) return fail(SFB_ERR_UNSUPPORTED, "channels not divisible by resnet_groups");
      const std::string pre = "downsamples." + std::to_string(i) + ".";
      lv.dw = up(pre + "downsample.weight", {lv.cout, lv.cin, 2 * lv.f + 1}); lv.db = up(pre + "downsample.bias", {lv.cout});
      if (!lv.dw || !lv.db) return SFB_ERR_MISSING;
      lv.blocks.resize(c.num_blocks[i]);
      for (int j = 0; j < c.num_blocks[i]; ++j)
        if (!load_block(pre + "blocks." + std::to_string(j) + ".", lv.cout, lv.cout, c.resnet_groups, lv.blocks[j])) return SFB_ERR_MISSING;
    }
    params.clear();
    finalized = true;
    return SFB_OK;
  }
  int64_t level_len(int64_t L, int i) const {      // output length of level i (Conv1d k = 2f+1, stride f, padding f)
    int64_t l = L;
    for (int k = 0; k <= i; ++k) l = (l + 2 * cfg.factors[k] - (2 * cfg.factors[k] + 1)) / cfg.factors[k] + 1;
    return l;
  }
  int n_gn() const { int n = 2; for (const Level& lv : levels) n += 2 * (int)lv.blocks.size(); return n; }
  size_t stats_bytes(int64_t B) const { return (size_t)n_gn() * B * 8 * 2 * sizeof(double); }
  size_t buf_floats(int64_t B, int64_t L) const {
    size_t m = (size_t)to_in.cout * L;
    for (int i = 0; i < (int)levels.size(); ++i) m = std::max(m, (size_t)levels[i].cout * level_len(L, i));
    return align_up(m * B, 256);
  }
  size_t workspace_bytes(int64_t B, int64_t L) const { return align_up(stats_bytes(B), 1024) + 4 * buf_floats(B, L) * sizeof(float) + 1024; }

  void conv(cudaStream_t st, const float* in, int Cin, int64_t Lin, const float* w, const float* bias, float* out, int Cout, int64_t Lout,
            int K, int stride, int pad, const double* gn_stats, int G, const float* gw, const float* gb, const float* res, int Cres,
            const float* sc_w, const float* sc_b, double* out_stats, int Gout, int64_t B) {
    EncConvParams p;
    p.in = in; p.w = w; p.bias = bias; p.out = out; p.gn_stats = gn_stats; p.gn_w = gw; p.gn_b = gb; p.res = res; p.sc_w = sc_w; p.sc_b = sc_b;
    p.out_stats = out_stats; p.G = G; p.Gout = Gout; p.Cin = Cin; p.Cout = Cout; p.Cres = Cres; p.Lin = (int)Lin; p.Lout = (int)Lout;
    p.K = K; p.stride = stride; p.pad = pad; p.eps = 1e-5f;
    enc_conv_kernel<<<dim3((unsigned)((Lout + 127) / 128), (unsigned)Cout, (unsigned)B), 128, 0, st>>>(p);
  }
  // X [B, cin, L] with statistics sX (g1 groups) -> out [B, cout, L]; next_stats / next_G: statistics the consumer needs
  void resnet(cudaStream_t st, const Block& b, const float* X, const double* sX, float* H, double* sH, float* out, double* next_stats,
              int next_G, int64_t L, int64_t B) {
    conv(st, X, b.cin, L, b.c1_w, b.c1_b, H, b.cout, L, 3, 1, 1, sX, b.g1, b.gn1_w, b.gn1_b, nullptr, 0, nullptr, nullptr, sH, b.g2, B);
    conv(st, H, b.cout, L, b.c2_w, b.c2_b, out, b.cout, L, 3, 1, 1, sH, b.g2, b.gn2_w, b.gn2_b, X, b.cin, b.sc_w, b.sc_b, next_stats, next_G, B);
  }
  int forward(const float* y, int64_t B, int64_t L, float* const* xs_out, int n_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!finalized) return fail(SFB_ERR_STATE, "finalize first");
    if (n_out != (int)levels.size() + 1) return fail(SFB_ERR_INVALID, "xs_out needs %d tensors (to_in output + one per level)", (int)levels.size() + 1);
    if (B <= 0 || L <= 0 || L > INT32_MAX / 4) return fail(SFB_ERR_INVALID, "bad B / L");
    if (!ws || ws_bytes < workspace_bytes(B, L)) return fail(SFB_ERR_INVALID, "encoder workspace too small");
    uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(ws), 1024));
    double* stats = reinterpret_cast<double*>(base);
    float* buf = reinterpret_cast<float*>(base + align_up(stats_bytes(B), 1024));
    const size_t bf = buf_floats(B, L);
    float *A = buf, *H = buf + bf, *R0 = buf + 2 * bf, *R1 = buf + 3 * bf;
    if (cudaMemsetAsync(stats, 0, stats_bytes(B), st) != cudaSuccess) return fail(SFB_ERR_CUDA, "memset");
    int si = 0;
    auto next_stats = [&]() { return stats + (size_t)(si++) * B * 16; };
    // to_in = ResnetBlock1d(in_channels -> channels * m0, one group)
    double* sY = next_stats();
    enc_stats_kernel<<<dim3((unsigned)((L + 1023) / 1024), (unsigned)cfg.in_channels, (unsigned)B), 256, 0, st>>>(y, sY, cfg.in_channels, (int)L, to_in.g1);
    resnet(st, to_in, y, sY, H, next_stats(), xs_out[0], nullptr, 1, L, B);
    const float* cur = xs_out[0];
    int64_t Lc = L;
    for (int i = 0; i < (int)levels.size(); ++i) {
      const Level& lv = levels[i];
      const int64_t Lo = level_len(L, i);
      const int nb = (int)lv.blocks.size();
      float* dst0 = nb == 0 ? xs_out[i + 1] : A;
      double* s_in = nb ? next_stats() : nullptr;
      conv(st, cur, lv.cin, Lc, lv.dw, lv.db, dst0, lv.cout, Lo, 2 * lv.f + 1, lv.f, lv.f, nullptr, 1, nullptr, nullptr, nullptr, 0, nullptr, nullptr,
           s_in, nb ? lv.blocks[0].g1 : 1, B);
      const float* x = dst0;
      for (int j = 0; j < nb; ++j) {
        const bool last = j == nb - 1;
        float* out = last ? xs_out[i + 1] : ((j & 1) ? R1 : R0);
        double* sH = next_stats();
        double* s_next = last ? nullptr : next_stats();
        resnet(st, lv.blocks[j], x, s_in, H, sH, out, s_next, last ? 1 : lv.blocks[j + 1].g1, Lo, B);
        x = out; s_in = s_next;
      }
      cur = xs_out[i + 1]; Lc = Lo;
    }
    return cudaGetLastError() == cudaSuccess ? SFB_OK : fail(SFB_ERR_CUDA, "encoder launch: %s", cudaGetErrorString(cudaGetLastError()));
  }
};

// Fault injection for the wait log: one warp waits on a barrier nobody ever arrives on (tests/test_gpu_waitlog.py).
struct FaultParams { int tag; };
#undef SFB_FILE_ID
#define SFB_FILE_ID 9
__global__ void wait_fault_kernel(const __grid_constant__ FaultParams p) {
  __shared__ __align__(8) uint64_t bar;
  mark_progress(p.tag);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  mbar_wait(&bar, 0);
}

// ================================================================================================== C ABI
struct sfb_handle {
  std::unique_ptr<EngineBase> e;
};

template <typename T>
static int dbg_gemm_t(const void* a1, const void* a2, const void* w, const float* bias, const float* resid, float* out_r,
                      void* out_t, double* stats, int B, int L, int K1, int K2, int N, int taps, int a2_bmod, int bias_mod,
                      int gs, cudaStream_t st) {
  if (set_kernel_attrs<T>() != cudaSuccess) return SFB_ERR_CUDA;
  const int BN = pick_bn(N);
  if (!BN) return SFB_ERR_UNSUPPORTED;
  GemmParams<T> p;
  memset(&p, 0, sizeof p);
  if (!fill_gemm_maps<T>(p, a1, K1, L, B, a2, K2, a2_bmod, w, N, taps, BN)) return SFB_ERR_CUDA;
  p.bias = bias; p.bias_mod = bias_mod > 0 ? bias_mod : N; p.gs = gs > 0 ? gs : 1;
  p.resid = resid; p.out_r = out_r; p.out_t = reinterpret_cast<T*>(out_t); p.stats = stats;
  launch_gemm<T>(p, BN, B, st);
  return cudaGetLastError() == cudaSuccess ? SFB_OK : SFB_ERR_CUDA;
}
template <typename T>
static int dbg_attn_t(const void* qkv, void* out, int B, int N, cudaStream_t st) {
  if (set_kernel_attrs<T>() != cudaSuccess) return SFB_ERR_CUDA;
  AttnParams<T> p;
  memset(&p, 0, sizeof p);
  constexpr int AE = ElemTraits<T>::kAtomElems;
  if (!make_tmap3<T>(&p.tmQ, qkv, 1536, N, B, AE, 128) || !make_tmap3<T>(&p.tmKV, qkv, 1536, N, B, AE, AttnCfg<T>::BKV) ||
      !make_tmap3<T>(&p.tmV, qkv, 1536, N, B, AE, AttnCfg<T>::BKV,
                     sizeof(T) == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
    return SFB_ERR_CUDA;
  p.out = reinterpret_cast<T*>(out); p.n_tokens = N; p.kv_tokens = N; p.q_col0 = 0; p.k_col0 = 512; p.v_col0 = 1024; p.scale_log2 = 1.4426950408889634f / 8.0f;
  long long* tl = nullptr;
  constexpr int kTl = 5 * 32 * 8;
  if (getenv("SFB_ATTN_TIMELINE")) {      // device-clock stamps of CTA 0's softmax warps and MMA thread (tools/attn_timeline.py)
    if (cudaMalloc(&tl, kTl * sizeof(long long)) != cudaSuccess) return SFB_ERR_CUDA;
    cudaMemset(tl, 0, kTl * sizeof(long long));
    p.dbg = tl;
  }
  attn_launch<T>(dim3((N + 127) / 128, 8, B), st, p);
  if (tl) {
    std::vector<long long> h(kTl);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), tl, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(tl);
    const long long t0 = h[0];
    static const char* const nm[] = {"wait_s", "s_ready", "ld_done", "max_done", "exp_done", "o_ready", "p_arrived"};
    for (int j = 0; j < 32 && h[j * 8] != 0; ++j) {
      for (int w = 0; w < 4; ++w) {
        fprintf(stderr, "tile %2d warp %d:", j, w + 2);
        for (int k = 0; k < 7; ++k) fprintf(stderr, " %s %6lld", nm[k], h[(w * 32 + j) * 8 + k] - t0);
        fprintf(stderr, "\n");
      }
      fprintf(stderr, "tile %2d mma   : wait_p %6lld p_ready %6lld\n", j, h[(4 * 32 + j) * 8] - t0, h[(4 * 32 + j) * 8 + 1] - t0);
    }
  }
  return cudaGetLastError() == cudaSuccess ? SFB_OK : SFB_ERR_CUDA;
}
extern "C" {

int sfb_create(const sfb_unet_config* cfg, int device, sfb_handle** out) {
  if (!cfg || !out) return SFB_ERR_INVALID;
  if (cfg->depth < 2 || cfg->depth > SFB_MAX_DEPTH) return SFB_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SFB_ERR_CUDA;   // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return SFB_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SFB_ERR_CUDA;
  if (prop.major != 10) return SFB_ERR_UNSUPPORTED;   // tcgen05 / TMEM kernels are sm_100a only
  if (wait_log_init() != 0) return SFB_ERR_CUDA;
  sfb_handle* h = new sfb_handle();
  if (cfg->precision == SFB_PRECISION_BF16) h->e.reset(new Engine<__nv_bfloat16>());
  else h->e.reset(new Engine<float>());
  h->e->cfg = *cfg;
  h->e->device = device;
  *out = h;
  return SFB_OK;
}

void sfb_destroy(sfb_handle* h) { delete h; }

const char* sfb_last_error(const sfb_handle* h) { return h ? h->e->err.c_str() : "null handle"; }

int sfb_set_param(sfb_handle* h, const char* name, const void* data, int dtype, const int64_t* shape, int ndim) {
  if (!h || !name || !data || !shape || ndim < 0 || ndim > 4) return SFB_ERR_INVALID;
  if (dtype != SFB_DTYPE_F32) return h->e->fail(SFB_ERR_INVALID, "only f32 parameters are accepted");
  if (h->e->finalized) return h->e->fail(SFB_ERR_STATE, "already finalized");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.v.resize(n);
  cudaError_t e = cudaMemcpy(t.v.data(), data, n * sizeof(float), cudaMemcpyDefault);
  if (e != cudaSuccess) return h->e->fail(SFB_ERR_CUDA, "set_param(%s): %s", name, cudaGetErrorString(e));
  h->e->params[name] = std::move(t);
  return SFB_OK;
}

int sfb_finalize(sfb_handle* h) {
  if (!h) return SFB_ERR_INVALID;
  if (h->e->finalized) return SFB_OK;
  cudaSetDevice(h->e->device);
  return h->e->finalize();
}

int sfb_workspace_bytes(sfb_handle* h, int64_t B, int64_t L, int cfg_on, int64_t rows, size_t* out) {
  if (!h || !out) return SFB_ERR_INVALID;
  return h->e->workspace_bytes(B, L, cfg_on, rows, 1, out);
}

int sfb_workspace_bytes_m(sfb_handle* h, int64_t B, int64_t L, int cfg_on, int64_t rows, int64_t M, size_t* out) {
  if (!h || !out) return SFB_ERR_INVALID;
  return h->e->workspace_bytes(B, L, cfg_on, rows, M, out);
}

int sfb_unet_forward(sfb_handle* h, const float* x, const float* sigma, const float* const* channels, int n_channels,
                     const float* embedding, int64_t M, float embedding_scale, float* v_out, int64_t B, int64_t L,
                     void* workspace, size_t workspace_bytes, void* stream) {
  if (!h) return SFB_ERR_INVALID;
  if (!x || !sigma || !channels || !v_out) return h->e->fail(SFB_ERR_INVALID, "null argument");
  if (cudaSetDevice(h->e->device) != cudaSuccess) return h->e->fail(SFB_ERR_CUDA, "cudaSetDevice(%d) failed", h->e->device);
  return h->e->unet_forward(x, sigma, channels, n_channels, embedding, M, embedding_scale, v_out, B, L, workspace,
                            workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int sfb_sample(sfb_handle* h, const float* x_noisy, int num_steps, const float* const* channels, int n_channels,
               const float* embedding, int64_t M, float embedding_scale, float* x_out, float* traj_x, float* traj_v,
               const float* teacher_x, int64_t B, int64_t L, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h) return SFB_ERR_INVALID;
  if (!x_noisy || !channels || !x_out) return h->e->fail(SFB_ERR_INVALID, "null argument");
  if (cudaSetDevice(h->e->device) != cudaSuccess) return h->e->fail(SFB_ERR_CUDA, "cudaSetDevice(%d) failed", h->e->device);
  return h->e->sample(x_noisy, num_steps, channels, n_channels, embedding, M, embedding_scale, x_out, traj_x, traj_v,
                      teacher_x, B, L, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int64_t sfb_last_launch_count(const sfb_handle* h) { return h ? h->e->launches : 0; }

int sfb_dbg_set_op_limit(sfb_handle* h, int n_ops) {
  if (!h) return SFB_ERR_INVALID;
  h->e->op_limit = n_ops;
  return SFB_OK;
}
int sfb_dbg_plan_size(sfb_handle* h, int64_t B, int64_t L, int cfg_on, void* workspace, size_t workspace_bytes) {
  if (!h) return SFB_ERR_INVALID;
  return h->e->plan_size(B, L, cfg_on, workspace, workspace_bytes);
}
int sfb_dbg_plan_size_m(sfb_handle* h, int64_t B, int64_t L, int cfg_on, int64_t M, void* workspace, size_t workspace_bytes) {
  if (!h) return SFB_ERR_INVALID;
  return h->e->plan_size(B, L, cfg_on, workspace, workspace_bytes, M);
}
int sfb_dbg_profile(sfb_handle* h, int enable) {
  if (!h) return SFB_ERR_INVALID;
  h->e->profiling = enable != 0;
  return SFB_OK;
}
int sfb_dbg_profile_report(sfb_handle* h, char* buf, int buf_len) {
  if (!h || !buf) return SFB_ERR_INVALID;
  return h->e->profile_report(buf, buf_len);
}
int sfb_dbg_sk_timeline(sfb_handle* h, int op_index, long long* host_buf, int n) {
  if (!h) return SFB_ERR_INVALID;
  return h->e->sk_timeline(op_index, host_buf, n);
}
struct sfb_encoder { OnsetEncoder e; };

int sfb_encoder_create(const sfb_encoder_config* cfg, int device, sfb_encoder** out) {
  if (!cfg || !out) return SFB_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SFB_ERR_CUDA;   // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return SFB_ERR_CUDA;
  sfb_encoder* h = new sfb_encoder();
  h->e.cfg = *cfg;
  h->e.device = device;
  *out = h;
  return SFB_OK;
}
void sfb_encoder_destroy(sfb_encoder* h) { delete h; }
const char* sfb_encoder_last_error(const sfb_encoder* h) { return h ? h->e.err.c_str() : "null handle"; }
int sfb_encoder_set_param(sfb_encoder* h, const char* name, const void* data, const int64_t* shape, int ndim) {
  if (!h || !name || !data || !shape || ndim < 0 || ndim > 4) return SFB_ERR_INVALID;
  if (h->e.finalized) return h->e.fail(SFB_ERR_STATE, "already finalized");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.v.resize(n);
  if (cudaMemcpy(t.v.data(), data, n * sizeof(float), cudaMemcpyDefault) != cudaSuccess) return h->e.fail(SFB_ERR_CUDA, "set_param(%s) copy failed", name);
  h->e.params[name] = std::move(t);
  return SFB_OK;
}
int sfb_encoder_finalize(sfb_encoder* h) {
  if (!h) return SFB_ERR_INVALID;
  cudaSetDevice(h->e.device);
  return h->e.finalize();
}
int64_t sfb_encoder_level_length(sfb_encoder* h, int64_t L, int level) {
  if (!h || level < -1 || level >= h->e.cfg.n_levels) return -1;
  return level < 0 ? L : h->e.level_len(L, level);
}
int sfb_encoder_workspace_bytes(sfb_encoder* h, int64_t B, int64_t L, size_t* out) {
  if (!h || !out) return SFB_ERR_INVALID;
  if (!h->e.finalized) return h->e.fail(SFB_ERR_STATE, "finalize first");
  *out = h->e.workspace_bytes(B, L);
  return SFB_OK;
}
int sfb_encoder_forward(sfb_encoder* h, const float* y, int64_t B, int64_t L, float* const* xs_out, int n_out, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (!h) return SFB_ERR_INVALID;
  if (!y || !xs_out) return h->e.fail(SFB_ERR_INVALID, "null argument");
  if (cudaSetDevice(h->e.device) != cudaSuccess) return h->e.fail(SFB_ERR_CUDA, "cudaSetDevice failed");
  return h->e.forward(y, B, L, xs_out, n_out, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int64_t sfb_postprocess_out_len(int64_t cut_length, int orig_freq, int new_freq) {
  if (cut_length <= 0 || orig_freq <= 0 || new_freq < 0) return -1;
  if (new_freq == 0 || new_freq == orig_freq) return cut_length;
  const int g = std::gcd(orig_freq, new_freq);
  const int64_t o = orig_freq / g, n = new_freq / g;
  return (n * cut_length + o - 1) / o;       // ceil(new * length / orig), torchaudio's target_length
}

int sfb_postprocess(int device, const float* gen, const float* onsets, int64_t B, int64_t L, int64_t cut_length, int orig_freq,
                    int new_freq, float* out, int64_t out_len, int* first_onset, void* stream) {
  if (!gen || !out || B <= 0 || L <= 0 || cut_length <= 0 || cut_length > L || orig_freq <= 0 || new_freq < 0) return SFB_ERR_INVALID;
  if (onsets && !first_onset) return SFB_ERR_INVALID;
  if (L > INT32_MAX / 2) return SFB_ERR_UNSUPPORTED;
  if (out_len != sfb_postprocess_out_len(cut_length, orig_freq, new_freq)) return SFB_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return SFB_ERR_CUDA;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (onsets) {
    fill_int_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(first_onset, (int)B, (int)L);
    first_onset_kernel<<<dim3((unsigned)((L / 4 + 256) / 256), (unsigned)B), 256, 0, st>>>(onsets, first_onset, (int)L);
  }
  PostParams p;
  p.gen = gen; p.first_onset = onsets ? first_onset : nullptr; p.table = nullptr; p.out = out;
  p.L = (int)L; p.cut = (int)cut_length; p.T = (int)out_len; p.orig = p.nw = p.K = p.width = 0; p.fpb = 1;
  if (new_freq == 0 || new_freq == orig_freq) {
    postprocess_copy_kernel<<<dim3((unsigned)std::min<int64_t>((out_len + 1023) / 1024, 4096), (unsigned)B), 256, 0, st>>>(p);
  } else {
    const ResampleTable* t = resample_table(device, orig_freq, new_freq);
    if (!t) return SFB_ERR_CUDA;
    p.table = t->dev; p.orig = t->orig; p.nw = t->nw; p.K = t->K; p.width = t->width;
    const int max_floats = 40 * 1024;                                // 160 KB of shared memory for the staged span
    if (t->K > max_floats) return SFB_ERR_UNSUPPORTED;              // rate pairs with a huge reduced orig_freq
    p.fpb = std::max(1, std::min(8, (max_floats - t->K) / t->orig + 1));
    const size_t smem = (size_t)((p.fpb - 1) * t->orig + t->K) * sizeof(float);
    static std::mutex mu;
    {
      std::lock_guard<std::mutex> lk(mu);
      if (cudaFuncSetAttribute(postprocess_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(max_floats * sizeof(float))) != cudaSuccess)
        return SFB_ERR_CUDA;
    }
    const int64_t frames = (out_len + t->nw - 1) / t->nw;
    postprocess_resample_kernel<<<dim3((unsigned)((frames + p.fpb - 1) / p.fpb), (unsigned)B), 256, smem, st>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? SFB_OK : SFB_ERR_CUDA;
}

int sfb_dbg_set_grid_limit(sfb_handle* h, int max_ctas) {
  if (!h || max_ctas < 0) return SFB_ERR_INVALID;
  g_grid_limit = max_ctas;
  return SFB_OK;
}
int sfb_dbg_wait_log(sfb_handle* h, char* buf, int buf_len) {
  if (!h || !buf || buf_len <= 0) return SFB_ERR_INVALID;
  const std::string w = h->e->wait_log_text();
  snprintf(buf, (size_t)buf_len, "%s", w.c_str());
  return (int)w.size();
}
int sfb_dbg_fault_inject(sfb_handle* h, void* stream) {
  if (!h) return SFB_ERR_INVALID;
  FaultParams fp{12345};
  wait_fault_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(fp);
  cudaError_t e = cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return h->e->fail(SFB_ERR_CUDA, "fault injection: %s", cudaGetErrorString(e));
  return SFB_OK;
}
int sfb_dbg_op_info(sfb_handle* h, int i, char* buf, int buf_len) {
  if (!h || !buf) return SFB_ERR_INVALID;
  return h->e->op_info(i, buf, buf_len);
}

int sfb_dbg_gemm(int bf16, const void* a1, const void* a2, const void* w, const float* bias, const float* resid,
                 float* out_r, void* out_t, double* stats, int B, int L, int K1, int K2, int N, int taps, int a2_bmod,
                 int bias_mod, int gs, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bf16) return dbg_gemm_t<__nv_bfloat16>(a1, a2, w, bias, resid, out_r, out_t, stats, B, L, K1, K2, N, taps, a2_bmod, bias_mod, gs, st);
  return dbg_gemm_t<float>(a1, a2, w, bias, resid, out_r, out_t, stats, B, L, K1, K2, N, taps, a2_bmod, bias_mod, gs, st);
}

int sfb_dbg_attention(int bf16, const void* qkv, void* out, int B, int N, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return bf16 ? dbg_attn_t<__nv_bfloat16>(qkv, out, B, N, st) : dbg_attn_t<float>(qkv, out, B, N, st);
}

}  // extern "C"
