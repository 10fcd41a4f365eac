// Internal kernel-launcher interface shared by the .cu files of libbeatthis_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bt {

constexpr int kHeadDim = 32;
constexpr int kMaxSlabs = 6;

// A GEMM over "planes": output row m = p_out * L + t.  The A operand row for slab s is
//   plane = p_out * plane_mul + plane_add[s],  time = t + t_shift[s]   (zero when time is
// outside [0, L)), columns [0, Kslab) of a row-major [planes_in * L, lda] activation.
// Plain linear: nslab 1.  Conv2d k(2,3) s(2,1) p(0,1): nslab 6 (df,dt).  frontend.linear
// over "b c f t -> b t (c f)": nslab 4 (f).  W is [N, nslab * Kslab] row-major.
struct GemmShape {
  int planes_out;
  int L;
  int N;
  int Kslab;
  int nslab;
  int plane_mul;
  int plane_add[kMaxSlabs];
  int t_shift[kMaxSlabs];
  int lda;
};

// Epilogue description (shared by the fp32 CUDA-core GEMM and the 16-bit tcgen05 GEMM).
struct EpiParams {
  int kind;            // 0 generic, 1 qkv (RoPE + q scaling), 2 attention gates:
                       //   out_f32[m*heads + n] = sigmoid(acc + bias[n]) for n < heads (N padded to 32)
  const float* bias;   // [N] or null
  int gelu;            // exact-erf GELU after bias
  const float* resid;  // fp32 [M, ldr] added after activation (may alias out_f32) or null
  int ldr;
  float* out_f32;      // optional fp32 output [M, ldo_f32]
  int ldo_f32;
  void* out_act;       // optional activation-dtype output [M, ldo_act]
  int ldo_act;
  // kind 1 (qkv): columns [0,C) q, [C,2C) k, [2C,3C) v, head h = (c % C) / 32
  const float* rope_cos;  // [Lmax, 16]
  const float* rope_sin;
  int C;
  int heads;
  int posmode;  // 0: position = m % L (time attention)   1: position = (m / L) % F (freq attention)
  int F;
  float qscale;  // multiplied into q after RoPE
};

struct ChunkSrc;

// ---- fp32 CUDA-core path -------------------------------------------------------------------
void launch_gemm_simt(const float* A, const float* W, const GemmShape& g, const EpiParams& e,
                      cudaStream_t st);
// qkv [seqs*L, 3C] fp32 (q,k roped; q NOT pre-scaled) -> out [seqs*L, C] (gated)
// chunks != nullptr: sequence s belongs to chunk s / seqs_per_chunk and only its first chunks[..].len keys exist
void launch_attn_time_simt(const float* qkv, const float* gates, float* out, int seqs, int L,
                           int heads, cudaStream_t st, const ChunkSrc* chunks = nullptr, int seqs_per_chunk = 1);
// rows [len_b, L) of every plane of chunk b <- 0 (the zero padding a k(2,3) convolution sees beyond the end of a
// chunk that is shorter than the wave's padded length).  buf: [nchunks * F, L, C] of elem_bytes-sized elements.
void launch_zero_tail(void* buf, int elem_bytes, const ChunkSrc* chunks, int nchunks, int F, int L, int C, cudaStream_t st);

// ---- shared small kernels (templated on activation dtype inside) ---------------------------
// frequency-direction attention: tokens m = (b*F + f)*L + t, sequences over f.
void launch_attn_freq(const void* qkv, const float* gates, void* out, int B, int F, int L,
                      int heads, float scale, int act_h16, cudaStream_t st);
// RMSNorm without gamma (folded into the next weight); optionally also the attention gates
// sigmoid(xn . wg[h] + bg[h]) for h < heads (used when heads <= 4; wg is [>=heads, C] fp32).
void launch_norm(const float* x, void* xn, int64_t M, int C, int act_h16, cudaStream_t st, float* gates = nullptr,
                 const float* wg = nullptr, const float* bg = nullptr, int heads = 0);
// per-chunk source description for the stem (chunk gather from per-clip spectrograms)
struct ChunkSrc {
  int64_t frame_base;  // first frame of the clip inside the concatenated spectrogram
  int32_t T;           // frames in the clip
  int32_t start;       // chunk start frame (may be negative)
  int64_t out_base;    // first frame of the clip inside the concatenated outputs
  int32_t write_lo;    // chunk-local frame range [write_lo, write_hi) this chunk owns
  int32_t write_hi;
  int32_t len;         // frames of this chunk (<= the wave's padded length L): rows [len, L) of its planes are padding
  int32_t pad_;
};
void launch_stem(const float* spect, const ChunkSrc* chunks, int nchunks, int L, const float* bn1_scale,
                 const float* bn1_shift, const float* w, const float* bias, float* out,
                 cudaStream_t st);
void launch_head(const float* x, int D, const float* w, const float* b, const ChunkSrc* chunks,
                 int nchunks, int L, float* beat, float* down, int sum_head, cudaStream_t st);
int launch_resample(const float* in, const int64_t* in_off_dev, float* out, const int64_t* out_off_dev, int n_clips,
                    int64_t max_out, const float* coef, int L, int M, int K, cudaStream_t st);
void launch_logmel(const float* audio, const int64_t* sample_off_dev, const int64_t* frame_off_dev,
                   int n_clips, int64_t max_frames, const float* window, const float* twiddle,
                   const int32_t* fb_start, const int32_t* fb_ptr, const float* fb_w, float* spect,
                   cudaStream_t st);
void launch_peakpick(const float* beat, const float* down, const int64_t* frame_off_dev, int n_clips,
                     double* beat_t, int32_t* n_beat, double* down_t, int32_t* n_down,
                     int max_peaks, cudaStream_t st);
void launch_f32_to_h16(const float* in, void* out, int64_t n, cudaStream_t st);
void launch_h16_to_f32(const void* in, float* out, int64_t n, cudaStream_t st);
// [seqs, L, heads*32] fp32 q,k,v -> packed qkv buffer [seqs*L, 3C] of the activation dtype
// (test hook for bt_debug_attention)
void launch_pack_qkv_test(const float* q, const float* k, const float* v, void* qkv, int seqs, int L,
                          int heads, float qscale, int act_h16, cudaStream_t st);

// ---- 16-bit (fp16 or bf16 operands) tcgen05 path ------------------------------------------------------------------------
struct TcGemmPlan;  // cached tensor maps + launch geometry
TcGemmPlan* tc_gemm_plan_create(const void* A_h16, const void* W_h16, const GemmShape& g,
                                int planes_in, char* err, int errlen);
void tc_gemm_plan_destroy(TcGemmPlan*);
int launch_gemm_tc(const TcGemmPlan* plan, const EpiParams& e, cudaStream_t st);

struct TcAttnPlan;
TcAttnPlan* tc_attn_plan_create(const void* qkv_h16, int seqs, int L, int heads, char* err, int errlen);
void tc_attn_plan_destroy(TcAttnPlan*);
void attn_prof_read(unsigned long long* out40, bool reset);  // debug: phase cycle counters of variant bit 64
void attn_set_variant(int v);  // debug: template parameter V of attn_tc48_kernel (kernels_attn.cu)
int launch_attn_time_tc(const TcAttnPlan* plan, const float* gates, void* out_h16, cudaStream_t st,
                        const ChunkSrc* chunks = nullptr, int seqs_per_chunk = 1);

// fused RMSNorm + FFN + residual for C in {32, 64} (frontend), x updated in place (+ optional 16-bit copy)
struct TcFfPlan;
TcFfPlan* tc_ff_plan_create(const void* w1_h16, const void* w2_h16, int C, int64_t M, const void* o_h16,
                            const void* wout_h16, char* err, int errlen);
void tc_ff_plan_destroy(TcFfPlan*);
int launch_fused_ff(const TcFfPlan* plan, float* X, const float* b1, const float* b2, void* xb_out, cudaStream_t st);

// fused RMSNorm + gates + QKV projection + RoPE for C in {32, 64} (frontend attentions)
struct TcQkvPlan;
TcQkvPlan* tc_qkv_plan_create(const void* wqkv_h16, int C, int64_t M, char* err, int errlen);
void tc_qkv_plan_destroy(TcQkvPlan*);
int launch_fused_qkv(const TcQkvPlan* plan, const float* X, const float* wg, const float* bg, const float* rope_cos,
                     const float* rope_sin, void* qkv, float* gates, int L, int F, int posmode, float qscale,
                     cudaStream_t st);

int tc_init(char* err, int errlen);  // resolves cuTensorMapEncodeTiled, sets smem attributes

}  // namespace bt
