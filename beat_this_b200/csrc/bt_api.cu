// C ABI (include/beatthis.h): context, packed-parameter registry, chunk planner, workspace
// and the per-wave schedule of the BeatThis forward pass (reference
// beat_this/model/beat_tracker.py:188-192, math restated in SURVEY.md App. A.3).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/beatthis.h"
#include "bt_kernels.h"

using namespace bt;

namespace {

struct Param {
  float* f32 = nullptr;
  void* b16 = nullptr;
  int32_t* i32 = nullptr;
  int64_t n = 0;
};

struct AttnW { const Param *wqkv, *wg, *bg, *wout; };
struct FfW { const Param *w1, *b1, *w2, *b2; };
// every parameter the forward pass touches, resolved once in bt_finalize (no name lookups per launch)
struct ModelW {
  const Param *rope_cos = nullptr, *rope_sin = nullptr, *bn1_scale = nullptr, *bn1_shift = nullptr, *stem_w = nullptr,
              *stem_b = nullptr, *lin_w = nullptr, *lin_b = nullptr, *head_w = nullptr, *head_b = nullptr;
  AttnW fa[3]{}, ta[3]{};
  FfW ff_f[3]{}, ff_t[3]{};
  const Param *conv_w[3] = {nullptr, nullptr, nullptr}, *conv_b[3] = {nullptr, nullptr, nullptr};
  std::vector<AttnW> la;
  std::vector<FfW> lf;
};
// small host -> device tables (offsets, chunk descriptors) travel through a ring of pinned slots: a slot is only
// waited for when it comes round again, kStageSlots uploads later, so no API call blocks on earlier GPU work
constexpr int kStageSlots = 16;
struct StageSlot {
  void* host = nullptr;
  void* dev = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
  bool pending = false;
};

// tensor-core plans for one (wave size, chunk length) geometry
struct AttnPlans { TcGemmPlan *qkv = nullptr, *out = nullptr, *gates = nullptr; TcAttnPlan* attn = nullptr; TcQkvPlan* fqkv = nullptr; };
struct FfPlans { TcGemmPlan *ff1 = nullptr, *ff2 = nullptr; TcFfPlan *fused = nullptr, *fused_op = nullptr; };
struct WavePlans {
  AttnPlans fa[3], ta[3];
  FfPlans ff_f[3], ff_t[3];
  TcGemmPlan* conv[3] = {nullptr, nullptr, nullptr};
  TcGemmPlan* lin = nullptr;
  std::vector<AttnPlans> la;
  std::vector<FfPlans> lf;
};

char g_create_error[512] = "";

}  // namespace

struct bt_ctx {
  int device = 0;
  bt_hparams hp{};
  int dtype = BT_DTYPE_F32;
  bool finalized = false;
  std::map<std::string, Param> params;
  mutable char err[1024] = "";
  int64_t launches = 0;
  bool sync_debug = false;
  bool fuse_ff = true;  // BT_FUSE_FF=0 falls back to norm + two GEMMs for the narrow frontend FFNs
  bool fuse_outproj = true;  // BT_FUSE_OUTPROJ=0: separate attention out-projection GEMM in front of the fused FFN

  // workspace: sized for ws_wave chunks of BT_CHUNK frames; grows on demand up to `wave`
  int wave = 128;
  int ws_wave = 0;
  float *X0 = nullptr, *X1 = nullptr, *GATES = nullptr;
  void *XB = nullptr, *XN = nullptr, *QKV = nullptr, *O = nullptr, *H = nullptr;
  // spectrogram scratch for bt_audio2frames
  float* spect_ws = nullptr;
  int64_t spect_cap = 0;
  // pinned staging + device tables
  StageSlot stage[kStageSlots];
  int stage_next = 0;
  ModelW mw;

  std::map<std::pair<int, int>, WavePlans*> plans;
  std::vector<std::pair<int, int>> plan_order;  // insertion order: oldest geometry is evicted first

  // per-kernel-class device timing (bt_profile_*): one event after every launch; the
  // duration of a launch is the gap to the previous event on the same stream
  bool prof = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct ProfRec { int kind; int ev; int prev; };
  std::vector<ProfRec> prof_recs;
  int prof_prev = -1;
  std::vector<std::string> prof_names;
  std::vector<double> prof_ms;
  std::vector<int64_t> prof_cnt;

  // debug tap
  std::string tap_name;
  float* tap_out = nullptr;
  int64_t tap_cap = 0;
  int64_t tap_count = 0;
};

namespace {

int fail(const bt_ctx* c, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  if (c) vsnprintf(c->err, sizeof(c->err), fmt, ap);
  else vsnprintf(g_create_error, sizeof(g_create_error), fmt, ap);
  va_end(ap);
  return code;
}

#define BT_CUDA(ctx, call)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail(ctx, BT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                  __FILE__, __LINE__);                                                       \
  } while (0)

int prof_event(bt_ctx* c, cudaStream_t st) {
  if (c->ev_used == c->ev_pool.size()) {
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return -1;
    c->ev_pool.push_back(ev);
  }
  const int idx = static_cast<int>(c->ev_used++);
  cudaEventRecord(c->ev_pool[idx], st);
  return idx;
}

void prof_mark(bt_ctx* c, cudaStream_t st) {  // start of an API call: reference point for the first kernel
  if (c->prof) c->prof_prev = prof_event(c, st);
}

void prof_launch(bt_ctx* c, const char* what, cudaStream_t st) {
  int kind = -1;
  for (size_t i = 0; i < c->prof_names.size(); ++i)
    if (c->prof_names[i] == what) { kind = static_cast<int>(i); break; }
  if (kind < 0) {
    kind = static_cast<int>(c->prof_names.size());
    c->prof_names.push_back(what);
    c->prof_ms.push_back(0.0);
    c->prof_cnt.push_back(0);
  }
  const int ev = prof_event(c, st);
  if (ev >= 0 && c->prof_prev >= 0) c->prof_recs.push_back({kind, ev, c->prof_prev});
  c->prof_prev = ev;
}

int check_launch(bt_ctx* c, const char* what, cudaStream_t st) {
  c->launches++;
  if (c->prof) prof_launch(c, what, st);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && c->sync_debug) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(c, BT_ERR_CUDA, "kernel %s failed: %s", what, cudaGetErrorString(e));
  return BT_OK;
}
#define BT_LAUNCHED(c, what, st)                 \
  do {                                           \
    int _r = check_launch(c, what, st);          \
    if (_r != BT_OK) return _r;                  \
  } while (0)

const Param* find_param(const bt_ctx* c, const std::string& name) {
  auto it = c->params.find(name);
  return it == c->params.end() ? nullptr : &it->second;
}

int64_t chunks_for(int64_t T, int64_t* starts, int64_t* lens, int64_t cap) {
  // split_piece (reference inference.py:119-125): starts = arange(-6, T-6, 1488); if T > 1488
  // the last start is moved to T - 1494; chunk = frames [start, start+1500) clipped to the
  // piece, zero padded by max(0,-start) on the left and max(0, min(6, start+1500-T)) on the right.
  if (T <= 0) return 0;
  const int64_t step = BT_CHUNK - 2 * BT_BORDER;
  int64_t n = 0;
  for (int64_t s = -BT_BORDER; s < T - BT_BORDER; s += step) ++n;
  int64_t i = 0;
  for (int64_t s = -BT_BORDER; s < T - BT_BORDER; s += step, ++i) {
    int64_t st = s;
    if (i == n - 1 && T > step) st = T - (BT_CHUNK - BT_BORDER);
    const int64_t lo = std::max<int64_t>(st, 0), hi = std::min<int64_t>(st + BT_CHUNK, T);
    const int64_t left = std::max<int64_t>(0, -st);
    const int64_t right = std::max<int64_t>(0, std::min<int64_t>(BT_BORDER, st + BT_CHUNK - T));
    if (i < cap) {
      if (starts) starts[i] = st;
      if (lens) lens[i] = (hi - lo) + left + right;
    }
  }
  return n;
}

int acquire_stage(bt_ctx* c, size_t bytes, StageSlot** out) {
  StageSlot* sl = &c->stage[c->stage_next];
  c->stage_next = (c->stage_next + 1) % kStageSlots;
  if (sl->pending) {  // kStageSlots uploads ago: long finished unless the caller is that far ahead of the GPU
    BT_CUDA(c, cudaEventSynchronize(sl->ev));
    sl->pending = false;
  }
  if (!sl->ev) BT_CUDA(c, cudaEventCreateWithFlags(&sl->ev, cudaEventDisableTiming));
  if (bytes > sl->cap) {
    if (sl->host) cudaFreeHost(sl->host);
    if (sl->dev) cudaFree(sl->dev);
    sl->host = sl->dev = nullptr;
    sl->cap = 0;
    const size_t cap = std::max<size_t>(bytes * 2, 1 << 16);
    BT_CUDA(c, cudaMallocHost(&sl->host, cap));
    BT_CUDA(c, cudaMalloc(&sl->dev, cap));
    sl->cap = cap;
  }
  *out = sl;
  return BT_OK;
}

int upload_stage(bt_ctx* c, StageSlot* sl, size_t bytes, cudaStream_t st) {
  BT_CUDA(c, cudaMemcpyAsync(sl->dev, sl->host, bytes, cudaMemcpyHostToDevice, st));
  BT_CUDA(c, cudaEventRecord(sl->ev, st));
  sl->pending = true;
  return BT_OK;
}

void free_ws(bt_ctx* c) {
  void** ptrs[] = {reinterpret_cast<void**>(&c->X0), reinterpret_cast<void**>(&c->X1),
                   reinterpret_cast<void**>(&c->GATES), &c->XB, &c->XN, &c->QKV, &c->O, &c->H};
  for (auto p : ptrs) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  c->ws_wave = 0;
}

void destroy_wave_plans(WavePlans* w) {
  auto fa = [](AttnPlans& a) {
    if (a.qkv) tc_gemm_plan_destroy(a.qkv);
    if (a.out) tc_gemm_plan_destroy(a.out);
    if (a.gates) tc_gemm_plan_destroy(a.gates);
    if (a.fqkv) tc_qkv_plan_destroy(a.fqkv);
    if (a.attn) tc_attn_plan_destroy(a.attn);
  };
  auto ff = [](FfPlans& f) {
    if (f.fused) tc_ff_plan_destroy(f.fused);
    if (f.fused_op) tc_ff_plan_destroy(f.fused_op);
    if (f.ff1) tc_gemm_plan_destroy(f.ff1);
    if (f.ff2) tc_gemm_plan_destroy(f.ff2);
  };
  for (int i = 0; i < 3; ++i) {
    fa(w->fa[i]); fa(w->ta[i]); ff(w->ff_f[i]); ff(w->ff_t[i]);
    if (w->conv[i]) tc_gemm_plan_destroy(w->conv[i]);
  }
  if (w->lin) tc_gemm_plan_destroy(w->lin);
  for (auto& a : w->la) fa(a);
  for (auto& f : w->lf) ff(f);
  delete w;
}

void free_plans(bt_ctx* c) {
  for (auto& kv : c->plans) destroy_wave_plans(kv.second);
  c->plans.clear();
  c->plan_order.clear();
}

// elements per chunk of the largest frontend activation: F*L*C is the same for all blocks
int64_t front_elems(const bt_ctx* c) {
  return static_cast<int64_t>(c->hp.spect_dim / 4) * BT_CHUNK * c->hp.stem_dim;
}

int ensure_ws(bt_ctx* c, int need_chunks) {
  // sized once for a full wave (bt_set_wave_chunks; ~46 MB per 1500-frame chunk on the 16-bit path): growing later
  // would mean cudaFree / cudaMalloc -- device-wide synchronisation -- in the middle of a stream of batches
  (void)need_chunks;
  const int want = c->wave;
  if (want <= c->ws_wave) return BT_OK;
  free_ws(c);
  free_plans(c);
  const int64_t G = want;
  const int64_t fe = front_elems(c);                                     // 1.536M
  const int64_t D = c->hp.transformer_dim;
  const int64_t me = static_cast<int64_t>(BT_CHUNK) * D;                 // main tokens * dim
  const int64_t xe = std::max(fe, me);
  const size_t act = c->dtype == BT_DTYPE_H16 ? 2 : 4;
  BT_CUDA(c, cudaMalloc(&c->X0, G * xe * 4));
  BT_CUDA(c, cudaMalloc(&c->X1, G * xe * 4));
  BT_CUDA(c, cudaMalloc(&c->GATES, G * std::max<int64_t>(fe / 32, BT_CHUNK * (D / 32)) * 4));
  BT_CUDA(c, cudaMalloc(&c->XN, G * xe * act));
  BT_CUDA(c, cudaMalloc(&c->QKV, G * 3 * xe * act));
  BT_CUDA(c, cudaMalloc(&c->O, G * xe * act));
  BT_CUDA(c, cudaMalloc(&c->H, G * 4 * xe * act));
  if (c->dtype == BT_DTYPE_H16) {
    BT_CUDA(c, cudaMalloc(&c->XB, G * xe * 2));
  }
  c->ws_wave = want;
  return BT_OK;
}

struct Wave {
  const ChunkSrc* chunks_dev;
  int nb;
  int L;         // padded length: the longest chunk of the wave
  bool varlen;   // some chunk is shorter than L
};

int do_tap(bt_ctx* c, const char* name, const void* buf, int64_t count, bool is_act, cudaStream_t st) {
  if (c->tap_name.empty() || !c->tap_out || c->tap_name != name) return BT_OK;
  if (count > c->tap_cap) return fail(c, BT_ERR_ARG, "tap %s needs %lld floats", name, (long long)count);
  if (is_act && c->dtype == BT_DTYPE_H16) {
    launch_h16_to_f32(buf, c->tap_out, count, st);
    BT_LAUNCHED(c, "tap_convert", st);
  } else {
    BT_CUDA(c, cudaMemcpyAsync(c->tap_out, buf, count * 4, cudaMemcpyDeviceToDevice, st));
  }
  c->tap_count = count;
  return BT_OK;
}

GemmShape plain_shape(int planes, int L, int N, int K, int lda) {
  GemmShape g{};
  g.planes_out = planes; g.L = L; g.N = N; g.Kslab = K; g.nslab = 1; g.plane_mul = 1; g.lda = lda;
  return g;
}

int run_gemm(bt_ctx* c, const void* A, const Param* W, TcGemmPlan* plan, const GemmShape& g,
             const EpiParams& e, const char* what, cudaStream_t st) {
  if (c->dtype == BT_DTYPE_H16) {
    if (launch_gemm_tc(plan, e, st) != 0) return fail(c, BT_ERR_CUDA, "tc gemm launch %s failed", what);
  } else {
    launch_gemm_simt(reinterpret_cast<const float*>(A), W->f32, g, e, st);
  }
  BT_LAUNCHED(c, what, st);
  return BT_OK;
}

EpiParams epi_generic(const Param* bias, int gelu, const float* resid, int ldr, float* out_f32, int ldo32,
                      void* out_act, int ldoa) {
  EpiParams e{};
  e.kind = 0;
  e.bias = bias ? bias->f32 : nullptr;
  e.gelu = gelu;
  e.resid = resid; e.ldr = ldr;
  e.out_f32 = out_f32; e.ldo_f32 = ldo32;
  e.out_act = out_act; e.ldo_act = ldoa;
  return e;
}

// x += attention(x) over `planes` planes of L tokens with dim C (reference roformer.py:114-132).
// freq == true: sequences run over the F planes of each chunk (PartialFTTransformer attnF).
// skip_out: the out-projection + residual are done by the following fused FFN kernel (fused_ff_kernel<C, true>).
int attention_block(bt_ctx* c, float* X, int planes, int L, int C, int F, bool freq, const AttnW& w,
                    AttnPlans* tp, int nb, cudaStream_t st, bool skip_out = false, const ChunkSrc* vl_chunks = nullptr) {
  const bool tc = c->dtype == BT_DTYPE_H16;
  const int heads = C / kHeadDim;
  const int64_t M = static_cast<int64_t>(planes) * L;
  const float inv_sqrt_d = 0.17677669529663687f;  // 1/sqrt(32): SDPA default scale (roformer.py:78-80)
  const bool tc_time = tc && !freq;
  const float qscale = tc_time ? inv_sqrt_d * 1.4426950408889634f : 1.0f;
  int r = BT_OK;
  if (tc && tp && tp->fqkv) {  // narrow frontend attentions: norm + gates + QKV + RoPE in one kernel
    if (launch_fused_qkv(tp->fqkv, X, w.wg->f32, w.bg->f32, c->mw.rope_cos->f32, c->mw.rope_sin->f32,
                         c->QKV, c->GATES, L, F, freq ? 1 : 0, qscale, st) != 0)
      return fail(c, BT_ERR_CUDA, "fused qkv launch failed");
    BT_LAUNCHED(c, C == 32 ? "qkv_fused_c32" : "qkv_fused_c64", st);
  } else {
    // heads 1/2 (unfused fallback of blocks 0/1): gates inside the norm kernel; heads >= 4: padded gates GEMM
    static const int gin_max = getenv("BT_GATES_IN_NORM_MAX") ? atoi(getenv("BT_GATES_IN_NORM_MAX")) : 2;
    const bool gates_in_norm = heads <= gin_max;
    launch_norm(X, c->XN, M, C, tc, st, gates_in_norm ? c->GATES : nullptr, w.wg->f32, w.bg->f32, heads);
    BT_LAUNCHED(c, gates_in_norm ? "norm_gates" : (F > 1 ? "norm_front" : "norm"), st);
    if (!gates_in_norm) {  // gates = sigmoid(to_gates(x_normed)): a [heads -> 32 padded] x C GEMM on the same rows
      GemmShape gg = plain_shape(planes, L, 32, C, C);
      EpiParams eg{};
      eg.kind = 2;
      eg.bias = w.bg->f32;
      eg.heads = heads;
      eg.out_f32 = c->GATES;
      int rg = run_gemm(c, c->XN, w.wg, tp ? tp->gates : nullptr, gg, eg, F > 1 ? "gemm_gates_front" : "gemm_gates", st);
      if (rg != BT_OK) return rg;
    }
    EpiParams e{};
    e.kind = 1;
    e.out_act = c->QKV; e.ldo_act = 3 * C;
    e.rope_cos = c->mw.rope_cos->f32;
    e.rope_sin = c->mw.rope_sin->f32;
    e.C = C; e.heads = heads; e.posmode = freq ? 1 : 0; e.F = F;
    e.qscale = qscale;
    GemmShape g = plain_shape(planes, L, 3 * C, C, C);
    r = run_gemm(c, c->XN, w.wqkv, tp ? tp->qkv : nullptr, g, e, F > 1 ? "gemm_qkv_front" : "gemm_qkv", st);
    if (r != BT_OK) return r;
  }
  if (freq) {
    launch_attn_freq(c->QKV, c->GATES, c->O, nb, F, L, heads, inv_sqrt_d, tc, st);
    BT_LAUNCHED(c, "attn_freq", st);
  } else if (tc) {
    if (launch_attn_time_tc(tp->attn, c->GATES, c->O, st, vl_chunks, planes / nb) != 0) return fail(c, BT_ERR_CUDA, "attn tc launch failed");
    BT_LAUNCHED(c, "attn_time_tc", st);
  } else {
    launch_attn_time_simt(reinterpret_cast<const float*>(c->QKV), c->GATES, reinterpret_cast<float*>(c->O),
                          planes, L, heads, st, vl_chunks, planes / nb);
    BT_LAUNCHED(c, "attn_time_simt", st);
  }
  if (skip_out) return BT_OK;
  GemmShape go = plain_shape(planes, L, C, C, C);
  EpiParams eo = epi_generic(nullptr, 0, X, C, X, C, nullptr, 0);
  return run_gemm(c, c->O, w.wout, tp ? tp->out : nullptr, go, eo, F > 1 ? "gemm_attn_out_front" : "gemm_attn_out", st);
}

// x += ff(x) (reference roformer.py:38-61); optionally also writes a bf16 copy of the result.
int ff_block(bt_ctx* c, float* X, int planes, int L, int C, int mult, const FfW& w, FfPlans* tp, void* copy_act,
             cudaStream_t st, bool with_outproj = false, bool front = false) {
  const bool tc = c->dtype == BT_DTYPE_H16;
  const int64_t M = static_cast<int64_t>(planes) * L;
  if (with_outproj && !(tc && tp && tp->fused_op)) return fail(c, BT_ERR_ARG, "fused out-projection requested without a plan");
  if (tc && tp && tp->fused) {
    if (launch_fused_ff(with_outproj ? tp->fused_op : tp->fused, X, w.b1->f32, w.b2->f32, copy_act, st) != 0)
      return fail(c, BT_ERR_CUDA, "fused ff launch failed");
    BT_LAUNCHED(c, C == 32 ? "ff_fused_c32" : "ff_fused_c64", st);
    return BT_OK;
  }
  launch_norm(X, c->XN, M, C, tc, st);
  BT_LAUNCHED(c, front ? "norm_front" : "norm", st);
  GemmShape g1 = plain_shape(planes, L, mult * C, C, C);
  EpiParams e1 = epi_generic(w.b1, 1, nullptr, 0, nullptr, 0, c->H, mult * C);
  int r = run_gemm(c, c->XN, w.w1, tp ? tp->ff1 : nullptr, g1, e1, front ? "gemm_ff1_front" : "gemm_ff1", st);
  if (r != BT_OK) return r;
  GemmShape g2 = plain_shape(planes, L, C, mult * C, mult * C);
  EpiParams e2 = epi_generic(w.b2, 0, X, C, X, C, copy_act, C);
  return run_gemm(c, c->H, w.w2, tp ? tp->ff2 : nullptr, g2, e2, front ? "gemm_ff2_front" : "gemm_ff2", st);
}

AttnW attn_w(const bt_ctx* c, const std::string& p) {
  return AttnW{find_param(c, p + ".wqkv"), find_param(c, p + ".wg"), find_param(c, p + ".bg"),
               find_param(c, p + ".wout")};
}
FfW ff_w(const bt_ctx* c, const std::string& p) {
  return FfW{find_param(c, p + ".w1"), find_param(c, p + ".b1"), find_param(c, p + ".w2"),
             find_param(c, p + ".b2")};
}

GemmShape conv_shape(int nb, int F, int L, int C) {
  // Conv2d(C -> 2C, k(2 freq, 3 time), stride (2,1), padding (0,1)) over [nb, F, L, C] as an
  // implicit GEMM: slab s = df*3 + dt reads plane 2*p_out + df at time t + dt - 1.
  GemmShape g{};
  g.planes_out = nb * F / 2; g.L = L; g.N = 2 * C; g.Kslab = C; g.nslab = 6; g.plane_mul = 2; g.lda = C;
  for (int df = 0; df < 2; ++df)
    for (int dt = 0; dt < 3; ++dt) { g.plane_add[df * 3 + dt] = df; g.t_shift[df * 3 + dt] = dt - 1; }
  return g;
}
GemmShape lin_shape(int nb, int L, int D, int Fo, int Co) {
  // "b c f t -> b t (c f)" + Linear: slab f reads plane Fo*b + f; W columns permuted to (f, c) on the host
  GemmShape g{};
  g.planes_out = nb; g.L = L; g.N = D; g.Kslab = Co; g.nslab = Fo; g.plane_mul = Fo; g.lda = Co;
  for (int f = 0; f < Fo; ++f) { g.plane_add[f] = f; g.t_shift[f] = 0; }
  return g;
}

int build_plans(bt_ctx* c, int nb, int L, WavePlans** out) {
  auto key = std::make_pair(nb, L);
  auto it = c->plans.find(key);
  if (it != c->plans.end()) { *out = it->second; return BT_OK; }
  // bounded cache: tensor maps are copied into the kernel parameters at launch, so dropping the oldest geometry is
  // safe while its kernels are still in flight
  constexpr size_t kMaxPlans = 48;
  while (c->plans.size() >= kMaxPlans && !c->plan_order.empty()) {
    auto old = c->plans.find(c->plan_order.front());
    if (old != c->plans.end()) { destroy_wave_plans(old->second); c->plans.erase(old); }
    c->plan_order.erase(c->plan_order.begin());
  }
  WavePlans* w = new WavePlans();
  char err[512] = "";
  auto mk = [&](const void* A, const Param* W, const GemmShape& g, int planes_in) -> TcGemmPlan* {
    return tc_gemm_plan_create(A, W->b16, g, planes_in, err, sizeof(err));
  };
  auto mk_attn = [&](AttnPlans& a, const AttnW& aw, int planes, int C, bool freq) -> bool {
    if (c->fuse_ff && (C == 32 || C == 64)) {
      a.fqkv = tc_qkv_plan_create(aw.wqkv->b16, C, static_cast<int64_t>(planes) * L, err, sizeof(err));
      if (!a.fqkv) return false;
    }
    a.qkv = mk(c->XN, aw.wqkv, plain_shape(planes, L, 3 * C, C, C), planes);
    a.out = mk(c->O, aw.wout, plain_shape(planes, L, C, C, C), planes);
    a.gates = mk(c->XN, aw.wg, plain_shape(planes, L, 32, C, C), planes);
    if (!a.qkv || !a.out || !a.gates) return false;
    if (!freq) {
      a.attn = tc_attn_plan_create(c->QKV, planes, L, C / 32, err, sizeof(err));
      if (!a.attn) return false;
    }
    return true;
  };
  auto mk_ff = [&](FfPlans& f, const FfW& fw, int planes, int C, int mult, const Param* wout = nullptr) -> bool {
    if (c->fuse_ff && mult == 4 && (C == 32 || C == 64)) {  // narrow frontend FFNs: one fused kernel
      f.fused = tc_ff_plan_create(fw.w1->b16, fw.w2->b16, C, static_cast<int64_t>(planes) * L, nullptr, nullptr, err, sizeof(err));
      if (f.fused && wout && c->fuse_outproj)  // ... and one with the preceding attention's out-projection in front
        f.fused_op = tc_ff_plan_create(fw.w1->b16, fw.w2->b16, C, static_cast<int64_t>(planes) * L, c->O, wout->b16, err, sizeof(err));
      return f.fused != nullptr && (!(wout && c->fuse_outproj) || f.fused_op != nullptr);
    }
    f.ff1 = mk(c->XN, fw.w1, plain_shape(planes, L, mult * C, C, C), planes);
    f.ff2 = mk(c->H, fw.w2, plain_shape(planes, L, C, mult * C, mult * C), planes);
    return f.ff1 && f.ff2;
  };
  bool ok = true;
  int C = c->hp.stem_dim, F = c->hp.spect_dim / 4;
  for (int i = 0; i < 3 && ok; ++i) {
    const std::string p = "b" + std::to_string(i);
    if (c->hp.partial_transformers) {
      ok = ok && mk_attn(w->fa[i], c->mw.fa[i], nb * F, C, true);
      ok = ok && mk_ff(w->ff_f[i], c->mw.ff_f[i], nb * F, C, 4, c->mw.fa[i].wout);
      ok = ok && mk_attn(w->ta[i], c->mw.ta[i], nb * F, C, false);
      ok = ok && mk_ff(w->ff_t[i], c->mw.ff_t[i], nb * F, C, 4, c->mw.ta[i].wout);
    }
    if (ok) {
      w->conv[i] = mk(c->XB, c->mw.conv_w[i], conv_shape(nb, F, L, C), nb * F);
      ok = w->conv[i] != nullptr;
    }
    C *= 2; F /= 2;
  }
  const int D = c->hp.transformer_dim;
  if (ok) {
    w->lin = mk(c->XN, c->mw.lin_w, lin_shape(nb, L, D, F, C), nb * F);
    ok = w->lin != nullptr;
  }
  w->la.resize(c->hp.n_layers);
  w->lf.resize(c->hp.n_layers);
  for (int l = 0; l < c->hp.n_layers && ok; ++l) {
    const std::string p = "l" + std::to_string(l);
    ok = ok && mk_attn(w->la[l], c->mw.la[l], nb, D, false);
    ok = ok && mk_ff(w->lf[l], c->mw.lf[l], nb, D, c->hp.ff_mult);
  }
  if (!ok) {  // never cache a half-built entry
    destroy_wave_plans(w);
    return fail(c, BT_ERR_CUDA, "tensor-core plan creation failed: %s", err);
  }
  c->plans[key] = w;
  c->plan_order.push_back(key);
  *out = w;
  return BT_OK;
}

// BeatThis.forward for one wave of nb equal-length chunks, scattering the head output.
int run_wave(bt_ctx* c, const float* spect, const Wave& wv, float* beat, float* down, cudaStream_t st) {
  const bool tc = c->dtype == BT_DTYPE_H16;
  const int nb = wv.nb, L = wv.L;
  WavePlans* wp = nullptr;
  if (tc) {
    int r = build_plans(c, nb, L, &wp);
    if (r != BT_OK) return r;
  }
  int r;
  float* X = c->X0;
  float* Xalt = c->X1;
  // chunks shorter than the wave's padded length: the time attentions mask their missing keys and the convolutions
  // see zeros beyond their last frame (everything else works row by row, padding rows are never read back)
  const ChunkSrc* vl = wv.varlen ? wv.chunks_dev : nullptr;
  int C = c->hp.stem_dim, F = c->hp.spect_dim / 4;
  launch_stem(spect, wv.chunks_dev, nb, L, c->mw.bn1_scale->f32, c->mw.bn1_shift->f32, c->mw.stem_w->f32, c->mw.stem_b->f32, X, st);
  BT_LAUNCHED(c, "stem", st);
  if ((r = do_tap(c, "stem", X, static_cast<int64_t>(nb) * F * L * C, false, st)) != BT_OK) return r;
  for (int i = 0; i < 3; ++i) {
    const std::string p = "b" + std::to_string(i);
    const int planes = nb * F;
    const int64_t elems = static_cast<int64_t>(planes) * L * C;
    void* copy_for_conv = tc ? c->XB : nullptr;
    if (c->hp.partial_transformers) {
      // out-projection + residual of an attention move into the following fused FFN kernel when there is a plan
      // for it (C = 32 / 64) and nobody asked to see the intermediate residual stream (debug tap)
      const bool op_f = wp && wp->ff_f[i].fused_op && (c->tap_name.empty() || c->tap_name != p + ".attnF");
      const bool op_t = wp && wp->ff_t[i].fused_op && (c->tap_name.empty() || c->tap_name != p + ".attnT");
      if ((r = attention_block(c, X, planes, L, C, F, true, c->mw.fa[i], wp ? &wp->fa[i] : nullptr, nb, st, op_f)) != BT_OK) return r;
      if ((r = do_tap(c, (p + ".attnF").c_str(), X, elems, false, st)) != BT_OK) return r;
      if ((r = ff_block(c, X, planes, L, C, 4, c->mw.ff_f[i], wp ? &wp->ff_f[i] : nullptr, nullptr, st, op_f, true)) != BT_OK) return r;
      if ((r = do_tap(c, (p + ".ffF").c_str(), X, elems, false, st)) != BT_OK) return r;
      if ((r = attention_block(c, X, planes, L, C, F, false, c->mw.ta[i], wp ? &wp->ta[i] : nullptr, nb, st, op_t, vl)) != BT_OK) return r;
      if ((r = do_tap(c, (p + ".attnT").c_str(), X, elems, false, st)) != BT_OK) return r;
      if ((r = ff_block(c, X, planes, L, C, 4, c->mw.ff_t[i], wp ? &wp->ff_t[i] : nullptr, copy_for_conv, st, op_t, true)) != BT_OK) return r;
      if ((r = do_tap(c, (p + ".ffT").c_str(), X, elems, false, st)) != BT_OK) return r;
    } else if (tc) {
      launch_f32_to_h16(X, c->XB, elems, st);
      BT_LAUNCHED(c, "f32_to_h16", st);
    }
    if (vl) {
      launch_zero_tail(tc ? c->XB : static_cast<void*>(X), tc ? 2 : 4, vl, nb, F, L, C, st);
      BT_LAUNCHED(c, "zero_tail", st);
    }
    // conv C -> 2C (+ folded BN2d + GELU); the last block feeds frontend.linear (activation dtype)
    GemmShape g = conv_shape(nb, F, L, C);
    const bool last = i == 2;
    EpiParams e = epi_generic(c->mw.conv_b[i], 1, nullptr, 0, last ? nullptr : Xalt, 2 * C,
                              last ? c->XN : nullptr, 2 * C);
    if ((r = run_gemm(c, tc ? c->XB : static_cast<const void*>(X), c->mw.conv_w[i],
                      wp ? wp->conv[i] : nullptr, g, e, "gemm_conv", st)) != BT_OK) return r;
    C *= 2; F /= 2;
    if (!last) std::swap(X, Xalt);
    if ((r = do_tap(c, (p + ".conv").c_str(), last ? c->XN : static_cast<const void*>(X),
                    static_cast<int64_t>(nb) * F * L * C, last, st)) != BT_OK) return r;
  }
  const int D = c->hp.transformer_dim;
  {
    GemmShape g = lin_shape(nb, L, D, F, C);
    EpiParams e = epi_generic(c->mw.lin_b, 0, nullptr, 0, X, D, nullptr, 0);
    if ((r = run_gemm(c, c->XN, c->mw.lin_w, wp ? wp->lin : nullptr, g, e, "gemm_frontend_linear", st)) != BT_OK) return r;
    if ((r = do_tap(c, "frontend", X, static_cast<int64_t>(nb) * L * D, false, st)) != BT_OK) return r;
  }
  for (int l = 0; l < c->hp.n_layers; ++l) {
    const std::string p = "l" + std::to_string(l);
    if ((r = attention_block(c, X, nb, L, D, 1, false, c->mw.la[l], wp ? &wp->la[l] : nullptr, nb, st, false, vl)) != BT_OK) return r;
    if ((r = do_tap(c, (p + ".attn").c_str(), X, static_cast<int64_t>(nb) * L * D, false, st)) != BT_OK) return r;
    if ((r = ff_block(c, X, nb, L, D, c->hp.ff_mult, c->mw.lf[l], wp ? &wp->lf[l] : nullptr, nullptr, st)) != BT_OK) return r;
    if ((r = do_tap(c, (p + ".ff").c_str(), X, static_cast<int64_t>(nb) * L * D, false, st)) != BT_OK) return r;
  }
  launch_head(X, D, c->mw.head_w->f32, c->mw.head_b->f32, wv.chunks_dev, nb, L, beat, down, c->hp.sum_head ? 1 : 0, st);
  BT_LAUNCHED(c, "head", st);
  return BT_OK;
}

struct HostChunk { ChunkSrc s; int len; };

// upload the chunk table and run the forward pass in waves of up to ws_wave chunks, longest first.  A wave is padded
// to its longest chunk: a shorter chunk costs its padded share of one wave (<= 0.35 ms on B200) instead of ~90
// launches of its own (~0.7 ms of fixed cost), so chunks of all lengths share waves
int run_chunks(bt_ctx* c, const float* spect_dev, std::vector<HostChunk>& all, float* beat_dev, float* downbeat_dev,
               cudaStream_t st) {
  int r = BT_OK;
  if (all.empty()) return BT_OK;
  if ((r = ensure_ws(c, static_cast<int>(all.size()))) != BT_OK) return r;
  std::stable_sort(all.begin(), all.end(), [](const HostChunk& a, const HostChunk& b) { return a.len > b.len; });
  const size_t bytes = all.size() * sizeof(ChunkSrc);
  StageSlot* sl = nullptr;
  if ((r = acquire_stage(c, bytes, &sl)) != BT_OK) return r;
  ChunkSrc* hs = static_cast<ChunkSrc*>(sl->host);
  for (size_t i = 0; i < all.size(); ++i) hs[i] = all[i].s;
  if ((r = upload_stage(c, sl, bytes, st)) != BT_OK) return r;
  const ChunkSrc* ds = static_cast<const ChunkSrc*>(sl->dev);
  size_t i = 0;
  while (i < all.size()) {
    const size_t j = std::min(all.size(), i + static_cast<size_t>(c->ws_wave));
    Wave wv{ds + i, static_cast<int>(j - i), all[i].len, all[j - 1].len != all[i].len};
    if ((r = run_wave(c, spect_dev, wv, beat_dev, downbeat_dev, st)) != BT_OK) return r;
    i = j;
  }
  return BT_OK;
}

std::vector<std::string> required_params(const bt_hparams& hp) {
  std::vector<std::string> v = {"mel.window", "mel.twiddle", "mel.fb_start", "mel.fb_ptr", "mel.fb_w",
                                "rope.cos", "rope.sin", "stem.bn1_scale", "stem.bn1_shift", "stem.w",
                                "stem.bias", "lin.w", "lin.b", "head.w", "head.b"};
  auto attn = [&](const std::string& p) { for (auto s : {".wqkv", ".wg", ".bg", ".wout"}) v.push_back(p + s); };
  auto ff = [&](const std::string& p) { for (auto s : {".w1", ".b1", ".w2", ".b2"}) v.push_back(p + s); };
  for (int i = 0; i < 3; ++i) {
    const std::string p = "b" + std::to_string(i);
    if (hp.partial_transformers) { attn(p + ".attnF"); ff(p + ".ffF"); attn(p + ".attnT"); ff(p + ".ffT"); }
    v.push_back(p + ".conv.w");
    v.push_back(p + ".conv.bias");
  }
  for (int l = 0; l < hp.n_layers; ++l) { attn("l" + std::to_string(l) + ".attn"); ff("l" + std::to_string(l) + ".ff"); }
  return v;
}

bool is_gemm_weight(const std::string& n) {
  auto ends = [&](const char* s) { size_t k = strlen(s); return n.size() >= k && n.compare(n.size() - k, k, s) == 0; };
  return ends(".wqkv") || ends(".wout") || ends(".wg") || ends(".w1") || ends(".w2") || ends(".conv.w") || n == "lin.w";
}

}  // namespace

// ================================================================================== C ABI
extern "C" {

int bt_version(void) { return 200; }

const char* bt_act_dtype(void) {
#if defined(BT_ACT_BF16)
  return "bf16";
#else
  return "f16";
#endif
}

const char* bt_last_error(const bt_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int64_t bt_num_frames(int64_t n_samples) { return n_samples < 0 ? 0 : 1 + n_samples / BT_HOP; }

int64_t bt_plan_chunks(int64_t T, int64_t* starts, int64_t* lens, int64_t cap) {
  return chunks_for(T, starts, lens, cap);
}

int bt_create(bt_ctx** out, int device_ordinal, const bt_hparams* hp, int compute_dtype) {
  if (!out || !hp) return fail(nullptr, BT_ERR_ARG, "bt_create: null argument");
  *out = nullptr;
  if (compute_dtype != BT_DTYPE_F32 && compute_dtype != BT_DTYPE_H16)
    return fail(nullptr, BT_ERR_ARG, "bt_create: compute_dtype must be BT_DTYPE_F32 or BT_DTYPE_H16");
  if (hp->spect_dim != 128 || hp->head_dim != 32 || hp->stem_dim != 32 || hp->transformer_dim % 64 != 0 ||
      hp->transformer_dim < 64 || hp->transformer_dim > 1024 || hp->n_layers < 1 || hp->ff_mult < 1 ||
      hp->ff_mult > 4)
    return fail(nullptr, BT_ERR_ARG,
                "bt_create: unsupported hyper-parameters (need spect_dim 128, head_dim 32, stem_dim 32, "
                "transformer_dim multiple of 64 in [64,1024], ff_mult <= 4)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, BT_ERR_CUDA, "bt_create: no CUDA device (%s); this library has no CPU fallback",
                cudaGetErrorString(e));
  if (device_ordinal < 0 || device_ordinal >= ndev)
    return fail(nullptr, BT_ERR_ARG, "bt_create: device %d out of range (%d devices)", device_ordinal, ndev);
  if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess)
    return fail(nullptr, BT_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device_ordinal);
  if (prop.major != 10)
    return fail(nullptr, BT_ERR_CUDA, "bt_create: device is sm_%d%d; this library is built for sm_100a only",
                prop.major, prop.minor);
  bt_ctx* c = new bt_ctx();
  c->device = device_ordinal;
  c->hp = *hp;
  c->dtype = compute_dtype;
  const char* dbg = getenv("BT_SYNC_DEBUG");
  c->sync_debug = dbg && dbg[0] == '1';
  const char* ffe = getenv("BT_FUSE_FF");
  c->fuse_ff = !(ffe && ffe[0] == '0');
  const char* foe = getenv("BT_FUSE_OUTPROJ");
  c->fuse_outproj = !(foe && foe[0] == '0');

  if (compute_dtype == BT_DTYPE_H16) {
    char err[512];
    if (tc_init(err, sizeof(err)) != 0) {
      delete c;
      return fail(nullptr, BT_ERR_CUDA, "%s", err);
    }
  }
  *out = c;
  return BT_OK;
}

int bt_set_param(bt_ctx* c, const char* name, const float* data_host, int64_t count) {
  if (!c || !name || !data_host || count <= 0) return fail(c, BT_ERR_ARG, "bt_set_param: bad argument");
  if (c->finalized) return fail(c, BT_ERR_STATE, "bt_set_param after bt_finalize");
  BT_CUDA(c, cudaSetDevice(c->device));
  Param& p = c->params[name];
  if (p.f32) cudaFree(p.f32);
  if (p.i32) cudaFree(p.i32);
  p = Param();
  p.n = count;
  const std::string n(name);
  if (n == "mel.fb_start" || n == "mel.fb_ptr") {
    std::vector<int32_t> tmp(count);
    for (int64_t i = 0; i < count; ++i) tmp[i] = static_cast<int32_t>(lrintf(data_host[i]));
    BT_CUDA(c, cudaMalloc(&p.i32, count * 4));
    BT_CUDA(c, cudaMemcpy(p.i32, tmp.data(), count * 4, cudaMemcpyHostToDevice));
  }
  BT_CUDA(c, cudaMalloc(&p.f32, count * 4));
  BT_CUDA(c, cudaMemcpy(p.f32, data_host, count * 4, cudaMemcpyHostToDevice));
  return BT_OK;
}

int bt_finalize(bt_ctx* c) {
  if (!c) return BT_ERR_ARG;
  if (c->finalized) return BT_OK;
  BT_CUDA(c, cudaSetDevice(c->device));
  const bt_hparams& hp = c->hp;
  for (const auto& n : required_params(hp))
    if (!find_param(c, n)) return fail(c, BT_ERR_PARAM, "bt_finalize: missing parameter '%s'", n.c_str());
  // shape checks for the GEMM weights
  auto expect = [&](const std::string& n, int64_t cnt) -> bool {
    const Param* p = find_param(c, n);
    if (!p || p->n != cnt) {
      fail(c, BT_ERR_PARAM, "parameter '%s' has %lld elements, expected %lld", n.c_str(),
           (long long)(p ? p->n : -1), (long long)cnt);
      return false;
    }
    return true;
  };
  auto chk_attn = [&](const std::string& p, int64_t C) {
    return expect(p + ".wqkv", 3 * C * C) && expect(p + ".wg", 32 * C) && expect(p + ".bg", 32) &&
           expect(p + ".wout", C * C);
  };
  auto chk_ff = [&](const std::string& p, int64_t C, int64_t mult) {
    return expect(p + ".w1", mult * C * C) && expect(p + ".b1", mult * C) && expect(p + ".w2", mult * C * C) &&
           expect(p + ".b2", C);
  };
  int64_t C = hp.stem_dim;
  for (int i = 0; i < 3; ++i) {
    const std::string p = "b" + std::to_string(i);
    if (hp.partial_transformers)
      if (!chk_attn(p + ".attnF", C) || !chk_ff(p + ".ffF", C, 4) || !chk_attn(p + ".attnT", C) ||
          !chk_ff(p + ".ffT", C, 4)) return BT_ERR_PARAM;
    if (!expect(p + ".conv.w", 2 * C * 6 * C) || !expect(p + ".conv.bias", 2 * C)) return BT_ERR_PARAM;
    C *= 2;
  }
  const int64_t D = hp.transformer_dim;
  if (!expect("lin.w", D * C * (hp.spect_dim / 32)) || !expect("lin.b", D) || !expect("head.w", 2 * D) ||
      !expect("head.b", 2) || !expect("stem.w", 32 * 12) || !expect("mel.window", 1024) ||
      !expect("mel.twiddle", 1024) || !expect("mel.fb_start", 128) || !expect("mel.fb_ptr", 129) ||
      !expect("rope.cos", (int64_t)BT_CHUNK * 16) || !expect("rope.sin", (int64_t)BT_CHUNK * 16))
    return BT_ERR_PARAM;
  for (int l = 0; l < hp.n_layers; ++l) {
    const std::string p = "l" + std::to_string(l);
    if (!chk_attn(p + ".attn", D) || !chk_ff(p + ".ff", D, hp.ff_mult)) return BT_ERR_PARAM;
  }
  if (c->dtype == BT_DTYPE_H16) {
    for (auto& kv : c->params) {
      if (!is_gemm_weight(kv.first)) continue;
      BT_CUDA(c, cudaMalloc(&kv.second.b16, kv.second.n * 2));
      launch_f32_to_h16(kv.second.f32, kv.second.b16, kv.second.n, nullptr);
    }
    BT_CUDA(c, cudaDeviceSynchronize());
  }
  {
    ModelW& w = c->mw;
    w.rope_cos = find_param(c, "rope.cos"); w.rope_sin = find_param(c, "rope.sin");
    w.bn1_scale = find_param(c, "stem.bn1_scale"); w.bn1_shift = find_param(c, "stem.bn1_shift");
    w.stem_w = find_param(c, "stem.w"); w.stem_b = find_param(c, "stem.bias");
    w.lin_w = find_param(c, "lin.w"); w.lin_b = find_param(c, "lin.b");
    w.head_w = find_param(c, "head.w"); w.head_b = find_param(c, "head.b");
    for (int i = 0; i < 3; ++i) {
      const std::string p = "b" + std::to_string(i);
      if (hp.partial_transformers) {
        w.fa[i] = attn_w(c, p + ".attnF"); w.ff_f[i] = ff_w(c, p + ".ffF");
        w.ta[i] = attn_w(c, p + ".attnT"); w.ff_t[i] = ff_w(c, p + ".ffT");
      }
      w.conv_w[i] = find_param(c, p + ".conv.w");
      w.conv_b[i] = find_param(c, p + ".conv.bias");
    }
    w.la.clear(); w.lf.clear();
    for (int l = 0; l < hp.n_layers; ++l) {
      w.la.push_back(attn_w(c, "l" + std::to_string(l) + ".attn"));
      w.lf.push_back(ff_w(c, "l" + std::to_string(l) + ".ff"));
    }
  }
  c->finalized = true;
  return BT_OK;
}

void bt_destroy(bt_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_plans(c);
  free_ws(c);
  for (auto& kv : c->params) {
    if (kv.second.f32) cudaFree(kv.second.f32);
    if (kv.second.b16) cudaFree(kv.second.b16);
    if (kv.second.i32) cudaFree(kv.second.i32);
  }
  if (c->spect_ws) cudaFree(c->spect_ws);
  for (auto& sl : c->stage) {
    if (sl.host) cudaFreeHost(sl.host);
    if (sl.dev) cudaFree(sl.dev);
    if (sl.ev) cudaEventDestroy(sl.ev);
  }
  for (auto ev : c->ev_pool) cudaEventDestroy(ev);
  delete c;
}

int bt_set_wave_chunks(bt_ctx* c, int32_t chunks) {
  if (!c || chunks < 1 || chunks > 256) return fail(c, BT_ERR_ARG, "bt_set_wave_chunks: 1..256");
  c->wave = chunks;
  if (c->ws_wave > chunks) {  // shrink: drop the workspace, it is re-created on the next call
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_ws(c);
    free_plans(c);
  }
  return BT_OK;
}

int64_t bt_launch_count(const bt_ctx* c) { return c ? c->launches : 0; }

int bt_profile_enable(bt_ctx* c, int enable) {
  if (!c) return BT_ERR_ARG;
  c->prof = enable != 0;
  c->prof_prev = -1;
  return BT_OK;
}

int bt_profile_collect(bt_ctx* c) {
  if (!c) return BT_ERR_ARG;
  BT_CUDA(c, cudaSetDevice(c->device));
  BT_CUDA(c, cudaDeviceSynchronize());
  for (const auto& r : c->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev_pool[r.prev], c->ev_pool[r.ev]) == cudaSuccess) {
      c->prof_ms[r.kind] += ms;
      c->prof_cnt[r.kind] += 1;
    }
  }
  c->prof_recs.clear();
  c->ev_used = 0;
  c->prof_prev = -1;
  return BT_OK;
}

int bt_profile_reset(bt_ctx* c) {
  if (!c) return BT_ERR_ARG;
  int r = bt_profile_collect(c);
  for (auto& v : c->prof_ms) v = 0.0;
  for (auto& v : c->prof_cnt) v = 0;
  return r;
}

int bt_profile_count(const bt_ctx* c) { return c ? static_cast<int>(c->prof_names.size()) : 0; }

int bt_profile_get(const bt_ctx* c, int index, char* name, int name_cap, double* total_ms, int64_t* launches) {
  if (!c || index < 0 || index >= static_cast<int>(c->prof_names.size())) return BT_ERR_ARG;
  if (name && name_cap > 0) snprintf(name, name_cap, "%s", c->prof_names[index].c_str());
  if (total_ms) *total_ms = c->prof_ms[index];
  if (launches) *launches = c->prof_cnt[index];
  return BT_OK;
}

int bt_debug_request_tap(bt_ctx* c, const char* tap, float* out_dev, int64_t cap) {
  if (!c) return BT_ERR_ARG;
  c->tap_name = tap ? tap : "";
  c->tap_out = out_dev;
  c->tap_cap = cap;
  c->tap_count = 0;
  return BT_OK;
}
int64_t bt_debug_tap_count(const bt_ctx* c) { return c ? c->tap_count : 0; }

int bt_logmel(bt_ctx* c, const float* audio_dev, const int64_t* sample_offsets_host, int32_t n_clips,
              float* spect_dev, const int64_t* frame_offsets_host, void* stream) {
  if (!c) return BT_ERR_ARG;
  for (const char* n : {"mel.window", "mel.twiddle", "mel.fb_start", "mel.fb_ptr", "mel.fb_w"})
    if (!find_param(c, n)) return fail(c, BT_ERR_STATE, "bt_logmel: parameter '%s' not set", n);
  if (n_clips <= 0) return BT_OK;
  if (!audio_dev || !sample_offsets_host || !spect_dev || !frame_offsets_host)
    return fail(c, BT_ERR_ARG, "bt_logmel: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BT_CUDA(c, cudaSetDevice(c->device));
  prof_mark(c, st);
  for (int i = 0; i < n_clips; ++i) {
    const int64_t len = sample_offsets_host[i + 1] - sample_offsets_host[i];
    if (len <= BT_N_FFT / 2)
      return fail(c, BT_ERR_ARG, "bt_logmel: clip %d has %lld samples; reflect padding needs more than %d "
                  "(torch.stft raises for such input as well)", i, (long long)len, BT_N_FFT / 2);
    if (frame_offsets_host[i + 1] - frame_offsets_host[i] != bt_num_frames(len))
      return fail(c, BT_ERR_ARG, "bt_logmel: frame_offsets do not match 1 + len/441 for clip %d", i);
  }
  const size_t bytes = static_cast<size_t>(n_clips + 1) * 8 * 2;
  StageSlot* sl = nullptr;
  int r = acquire_stage(c, bytes, &sl);
  if (r != BT_OK) return r;
  int64_t* h = static_cast<int64_t*>(sl->host);
  memcpy(h, sample_offsets_host, (n_clips + 1) * 8);
  memcpy(h + n_clips + 1, frame_offsets_host, (n_clips + 1) * 8);
  if ((r = upload_stage(c, sl, bytes, st)) != BT_OK) return r;
  const int64_t* d = static_cast<const int64_t*>(sl->dev);
  const int64_t f0 = frame_offsets_host[0];
  const int64_t total = frame_offsets_host[n_clips] - f0;
  if (f0 != 0) return fail(c, BT_ERR_ARG, "bt_logmel: frame_offsets_host[0] must be 0");
  int64_t max_frames = 0;
  for (int i = 0; i < n_clips; ++i) max_frames = std::max(max_frames, frame_offsets_host[i + 1] - frame_offsets_host[i]);
  (void)total;
  launch_logmel(audio_dev, d, d + n_clips + 1, n_clips, max_frames, find_param(c, "mel.window")->f32,
                find_param(c, "mel.twiddle")->f32, find_param(c, "mel.fb_start")->i32,
                find_param(c, "mel.fb_ptr")->i32, find_param(c, "mel.fb_w")->f32, spect_dev, st);
  BT_LAUNCHED(c, "logmel", st);
  return BT_OK;
}

int bt_resample(bt_ctx* c, const float* audio_in_dev, const int64_t* in_offsets_host, int32_t n_clips,
                const float* coef_dev, int32_t L, int32_t M, int32_t K, float* audio_out_dev,
                const int64_t* out_offsets_host, void* stream) {
  if (!c) return BT_ERR_ARG;
  if (n_clips <= 0) return BT_OK;
  if (!audio_in_dev || !in_offsets_host || !coef_dev || !audio_out_dev || !out_offsets_host)
    return fail(c, BT_ERR_ARG, "bt_resample: null argument");
  if (L <= 0 || M <= 0 || K <= 0 || (K & 1)) return fail(c, BT_ERR_ARG, "bt_resample: need L, M > 0 and an even K > 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BT_CUDA(c, cudaSetDevice(c->device));
  prof_mark(c, st);
  int64_t max_out = 0;
  for (int i = 0; i < n_clips; ++i) {
    if (in_offsets_host[i + 1] < in_offsets_host[i] || out_offsets_host[i + 1] < out_offsets_host[i])
      return fail(c, BT_ERR_ARG, "bt_resample: offsets must be non-decreasing");
    max_out = std::max(max_out, out_offsets_host[i + 1] - out_offsets_host[i]);
  }
  const size_t bytes = static_cast<size_t>(n_clips + 1) * 8 * 2;
  StageSlot* sl = nullptr;
  int r = acquire_stage(c, bytes, &sl);
  if (r != BT_OK) return r;
  int64_t* h = static_cast<int64_t*>(sl->host);
  memcpy(h, in_offsets_host, (n_clips + 1) * 8);
  memcpy(h + n_clips + 1, out_offsets_host, (n_clips + 1) * 8);
  if ((r = upload_stage(c, sl, bytes, st)) != BT_OK) return r;
  const int64_t* d = static_cast<const int64_t*>(sl->dev);
  if (launch_resample(audio_in_dev, d, audio_out_dev, d + n_clips + 1, n_clips, max_out, coef_dev, L, M, K, st) != 0)
    return fail(c, BT_ERR_ARG, "bt_resample: ratio %d/%d with %d taps needs too much shared memory", L, M, K);
  BT_LAUNCHED(c, "resample", st);
  return BT_OK;
}

int bt_spect2frames(bt_ctx* c, const float* spect_dev, const int64_t* frame_offsets_host, int32_t n_clips,
                    float* beat_dev, float* downbeat_dev, void* stream) {
  if (!c || !c->finalized) return fail(c, BT_ERR_STATE, "bt_spect2frames: context not finalized");
  if (n_clips <= 0) return BT_OK;
  if (!spect_dev || !frame_offsets_host || !beat_dev || !downbeat_dev)
    return fail(c, BT_ERR_ARG, "bt_spect2frames: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BT_CUDA(c, cudaSetDevice(c->device));
  prof_mark(c, st);
  int r = BT_OK;
  // plan: all chunks of all clips, grouped by chunk length (1500 except for pieces <= 1488 frames)
  std::vector<HostChunk> all;
  std::vector<int64_t> starts, lens;
  for (int i = 0; i < n_clips; ++i) {
    const int64_t T = frame_offsets_host[i + 1] - frame_offsets_host[i];
    if (T < 0) return fail(c, BT_ERR_ARG, "bt_spect2frames: negative clip length");
    if (T == 0) continue;
    const int64_t n = chunks_for(T, nullptr, nullptr, 0);
    starts.resize(n); lens.resize(n);
    chunks_for(T, starts.data(), lens.data(), n);
    for (int64_t j = 0; j < n; ++j) {
      HostChunk hc;
      hc.s.frame_base = frame_offsets_host[i];
      hc.s.out_base = frame_offsets_host[i];
      hc.s.T = static_cast<int32_t>(T);
      hc.s.start = static_cast<int32_t>(starts[j]);
      // keep_first (inference.py:174-184): chunk j owns [start+6, start+len-6) minus what earlier chunks own
      int64_t lo = starts[j] + BT_BORDER;
      if (j > 0) lo = std::max(lo, starts[j - 1] + lens[j - 1] - BT_BORDER);
      const int64_t hi = starts[j] + lens[j] - BT_BORDER;
      hc.s.write_lo = static_cast<int32_t>(lo - starts[j]);
      hc.s.write_hi = static_cast<int32_t>(std::max(lo, hi) - starts[j]);
      hc.s.len = static_cast<int32_t>(lens[j]);
      hc.s.pad_ = 0;
      hc.len = static_cast<int>(lens[j]);
      all.push_back(hc);
    }
  }
  return run_chunks(c, spect_dev, all, beat_dev, downbeat_dev, st);
}

int bt_forward_chunks(bt_ctx* c, const float* chunks_dev, int32_t n_chunks, int32_t chunk_frames, float* beat_dev,
                      float* downbeat_dev, void* stream) {
  if (!c || !c->finalized) return fail(c, BT_ERR_STATE, "bt_forward_chunks: context not finalized");
  if (n_chunks <= 0) return BT_OK;
  if (!chunks_dev || !beat_dev || !downbeat_dev) return fail(c, BT_ERR_ARG, "bt_forward_chunks: null argument");
  if (chunk_frames < 1 || chunk_frames > BT_CHUNK)
    return fail(c, BT_ERR_ARG, "bt_forward_chunks: chunk_frames must be in [1, %d]", BT_CHUNK);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BT_CUDA(c, cudaSetDevice(c->device));
  prof_mark(c, st);
  std::vector<HostChunk> all(n_chunks);
  for (int i = 0; i < n_chunks; ++i) {
    HostChunk& hc = all[i];
    hc.s.frame_base = static_cast<int64_t>(i) * chunk_frames;
    hc.s.out_base = hc.s.frame_base;
    hc.s.T = chunk_frames;
    hc.s.start = 0;
    hc.s.write_lo = 0;
    hc.s.write_hi = chunk_frames;
    hc.s.len = chunk_frames;
    hc.s.pad_ = 0;
    hc.len = chunk_frames;
  }
  return run_chunks(c, chunks_dev, all, beat_dev, downbeat_dev, st);
}

int bt_audio2frames(bt_ctx* c, const float* audio_dev, const int64_t* sample_offsets_host, int32_t n_clips,
                    float* beat_dev, float* downbeat_dev, const int64_t* frame_offsets_host, void* stream) {
  if (!c || !c->finalized) return fail(c, BT_ERR_STATE, "bt_audio2frames: context not finalized");
  if (n_clips <= 0) return BT_OK;
  if (!frame_offsets_host) return fail(c, BT_ERR_ARG, "bt_audio2frames: null argument");
  BT_CUDA(c, cudaSetDevice(c->device));
  const int64_t total = frame_offsets_host[n_clips];
  if (total * 128 > c->spect_cap) {
    if (c->spect_ws) cudaFree(c->spect_ws);
    c->spect_ws = nullptr;
    c->spect_cap = total * 128 * 5 / 4;
    BT_CUDA(c, cudaMalloc(&c->spect_ws, c->spect_cap * 4));
  }
  int r = bt_logmel(c, audio_dev, sample_offsets_host, n_clips, c->spect_ws, frame_offsets_host, stream);
  if (r != BT_OK) return r;
  return bt_spect2frames(c, c->spect_ws, frame_offsets_host, n_clips, beat_dev, downbeat_dev, stream);
}

int bt_peakpick(bt_ctx* c, const float* beat_dev, const float* downbeat_dev, const int64_t* frame_offsets_host,
                int32_t n_clips, double* beat_times_dev, int32_t* n_beats_dev, double* down_times_dev,
                int32_t* n_down_dev, int32_t max_peaks, void* stream) {
  if (!c) return BT_ERR_ARG;
  if (n_clips <= 0) return BT_OK;
  if (!beat_dev || !downbeat_dev || !frame_offsets_host || !beat_times_dev || !n_beats_dev || !down_times_dev ||
      !n_down_dev || max_peaks < 1)
    return fail(c, BT_ERR_ARG, "bt_peakpick: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BT_CUDA(c, cudaSetDevice(c->device));
  prof_mark(c, st);
  const size_t bytes = static_cast<size_t>(n_clips + 1) * 8;
  StageSlot* sl = nullptr;
  int r = acquire_stage(c, bytes, &sl);
  if (r != BT_OK) return r;
  memcpy(sl->host, frame_offsets_host, bytes);
  if ((r = upload_stage(c, sl, bytes, st)) != BT_OK) return r;
  launch_peakpick(beat_dev, downbeat_dev, static_cast<const int64_t*>(sl->dev), n_clips, beat_times_dev,
                  n_beats_dev, down_times_dev, n_down_dev, max_peaks, st);
  BT_LAUNCHED(c, "peakpick", st);
  return BT_OK;
}

int bt_debug_gemm(bt_ctx* c, const float* a_dev, const float* w_dev, float* d_dev, int32_t M, int32_t N, int32_t K,
                  void* stream) {
  if (!c) return BT_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BT_CUDA(c, cudaSetDevice(c->device));
  GemmShape g = plain_shape(1, M, N, K, K);
  EpiParams e = epi_generic(nullptr, 0, nullptr, 0, d_dev, N, nullptr, 0);
  if (c->dtype == BT_DTYPE_H16) {
    void *ab = nullptr, *wb = nullptr;
    BT_CUDA(c, cudaMalloc(&ab, static_cast<size_t>(M) * K * 2));
    BT_CUDA(c, cudaMalloc(&wb, static_cast<size_t>(N) * K * 2));
    launch_f32_to_h16(a_dev, ab, static_cast<int64_t>(M) * K, st);
    launch_f32_to_h16(w_dev, wb, static_cast<int64_t>(N) * K, st);
    char err[512] = "";
    TcGemmPlan* p = tc_gemm_plan_create(ab, wb, g, 1, err, sizeof(err));
    int rc = BT_OK;
    if (!p) rc = fail(c, BT_ERR_CUDA, "%s", err);
    else if (launch_gemm_tc(p, e, st) != 0) rc = fail(c, BT_ERR_CUDA, "tc gemm launch failed");
    cudaError_t se = cudaStreamSynchronize(st);
    if (rc == BT_OK && se != cudaSuccess) rc = fail(c, BT_ERR_CUDA, "tc gemm: %s", cudaGetErrorString(se));
    if (p) tc_gemm_plan_destroy(p);
    cudaFree(ab); cudaFree(wb);
    c->launches++;
    return rc;
  }
  launch_gemm_simt(a_dev, w_dev, g, e, st);
  BT_LAUNCHED(c, "debug_gemm", st);
  return BT_OK;
}

int bt_debug_attention(bt_ctx* c, const float* q_dev, const float* k_dev, const float* v_dev, float* o_dev,
                       int32_t seqs, int32_t L, int32_t heads, void* stream) {
  if (!c) return BT_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BT_CUDA(c, cudaSetDevice(c->device));
  const int C = heads * 32;
  const int64_t M = static_cast<int64_t>(seqs) * L;
  const bool tc = c->dtype == BT_DTYPE_H16;
  const size_t act = tc ? 2 : 4;
  void *qkv = nullptr, *o = nullptr;
  float* gates = nullptr;
  BT_CUDA(c, cudaMalloc(&qkv, M * 3 * C * act));
  BT_CUDA(c, cudaMalloc(&o, M * C * act));
  BT_CUDA(c, cudaMalloc(&gates, M * heads * 4));
  std::vector<float> ones(M * heads, 1.0f);
  BT_CUDA(c, cudaMemcpyAsync(gates, ones.data(), M * heads * 4, cudaMemcpyHostToDevice, st));
  int rc = BT_OK;
  if (tc) {
    launch_pack_qkv_test(q_dev, k_dev, v_dev, qkv, seqs, L, heads, 0.17677669529663687f * 1.4426950408889634f, 1, st);
    char err[512] = "";
    TcAttnPlan* p = tc_attn_plan_create(qkv, seqs, L, heads, err, sizeof(err));
    if (!p) rc = fail(c, BT_ERR_CUDA, "%s", err);
    else {
      launch_attn_time_tc(p, gates, o, st);
      launch_h16_to_f32(o, o_dev, M * C, st);
    }
    cudaError_t se = cudaStreamSynchronize(st);
    if (rc == BT_OK && se != cudaSuccess) rc = fail(c, BT_ERR_CUDA, "tc attention: %s", cudaGetErrorString(se));
    if (p) tc_attn_plan_destroy(p);
  } else {
    launch_pack_qkv_test(q_dev, k_dev, v_dev, qkv, seqs, L, heads, 1.0f, 0, st);
    launch_attn_time_simt(static_cast<const float*>(qkv), gates, o_dev, seqs, L, heads, st);
    cudaError_t se = cudaStreamSynchronize(st);
    if (se != cudaSuccess) rc = fail(c, BT_ERR_CUDA, "simt attention: %s", cudaGetErrorString(se));
  }
  c->launches += 2;
  cudaFree(qkv); cudaFree(o); cudaFree(gates);
  return rc;
}

int bt_debug_attention_time(bt_ctx* c, int32_t seqs, int32_t L, int32_t heads, int32_t variant, int32_t iters,
                            float* ms_per_launch) {
  if (!c || !ms_per_launch || iters < 1) return BT_ERR_ARG;
  if (c->dtype != BT_DTYPE_H16) return fail(c, BT_ERR_ARG, "bt_debug_attention_time needs the 16-bit context");
  BT_CUDA(c, cudaSetDevice(c->device));
  const int C = heads * 32;
  const int64_t M = static_cast<int64_t>(seqs) * L;
  void *qkv = nullptr, *o = nullptr;
  float *gates = nullptr, *src = nullptr;
  BT_CUDA(c, cudaMalloc(&qkv, M * 3 * C * 2));
  BT_CUDA(c, cudaMalloc(&o, M * C * 2));
  BT_CUDA(c, cudaMalloc(&gates, M * heads * 4));
  BT_CUDA(c, cudaMalloc(&src, M * C * 4));
  std::vector<float> h(M * C);
  uint32_t x = 12345u;
  for (auto& v : h) { x = x * 1664525u + 1013904223u; v = (static_cast<float>(x >> 8) / 8388608.0f - 1.0f) * 1.5f; }
  BT_CUDA(c, cudaMemcpy(src, h.data(), M * C * 4, cudaMemcpyHostToDevice));
  std::vector<float> ones(M * heads, 1.0f);
  BT_CUDA(c, cudaMemcpy(gates, ones.data(), M * heads * 4, cudaMemcpyHostToDevice));
  cudaStream_t st = nullptr;
  launch_pack_qkv_test(src, src, src, qkv, seqs, L, heads, 0.17677669529663687f * 1.4426950408889634f, 1, st);
  char err[512] = "";
  TcAttnPlan* p = tc_attn_plan_create(qkv, seqs, L, heads, err, sizeof(err));
  int rc = BT_OK;
  if (!p) rc = fail(c, BT_ERR_CUDA, "%s", err);
  else {
    if (variant >= 0) attn_set_variant(variant);
    unsigned long long prof[40];
    attn_prof_read(prof, true);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2 && rc == BT_OK; ++i)
      if (launch_attn_time_tc(p, gates, o, st) != 0) rc = fail(c, BT_ERR_ARG, "unknown attention variant %d", variant);
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters && rc == BT_OK; ++i) launch_attn_time_tc(p, gates, o, st);
    cudaEventRecord(e1, st);
    cudaError_t se = cudaStreamSynchronize(st);
    if (rc == BT_OK && se != cudaSuccess) rc = fail(c, BT_ERR_CUDA, "attention: %s", cudaGetErrorString(se));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_per_launch = ms / iters;
    if (variant >= 0 && (variant & 64)) {
      attn_prof_read(prof, true);
      for (int w = 0; w < 4; ++w) {
        const unsigned long long* q = prof + 8 * w;
        const double n = static_cast<double>(q[6] ? q[6] : 1);  // warp-tiles
        fprintf(stderr, "attention phases, warp %d, cycles per tile: wait S %.0f | ld S %.0f | max %.0f | exp %.0f | wait PV %.0f | st P %.0f\n",
                w, q[0] / n, q[1] / n, q[2] / n, q[3] / n, q[4] / n, q[5] / n);
      }
      if (prof[34]) fprintf(stderr, "issuer warp, cycles per tile: wait P %.0f | rest %.0f\n", prof[32] / double(prof[34]), prof[33] / double(prof[34]));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    tc_attn_plan_destroy(p);
    if (variant >= 0) attn_set_variant(-1);
  }
  cudaFree(qkv); cudaFree(o); cudaFree(gates); cudaFree(src);
  return rc;
}

}  // extern "C"
