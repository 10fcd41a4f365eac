// HBM-bound kernels of the path: log-mel frontend, stem conv, RMSNorm(+gates),
// frequency-direction attention, head + aggregation scatter, peak picking.
#include <cuda_fp16.h>
#include <cstdlib>
#include <mutex>

#include "bt_kernels.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace bt {

// ------------------------------------------------------------------------------------------
// log-mel: reference LogMelSpect.forward (beat_this/preprocessing.py:56-59) =
//   torch.stft(n_fft 1024, hop 441, periodic hann, center reflect, normalized) -> abs ->
//   mel filterbank (slaney, 128 bins, 30..11000 Hz) -> log1p(1000 x).
// Algorithmic HBM bytes: 441 new samples * 4 B read + 128 * 4 B written per frame.
//
// 64 threads per frame, two frames per CTA.  The real 1024-point transform is ONE complex 512-point FFT of
// z[n] = x[2n] + i x[2n+1] followed by the usual untangling step, and 512 = 8 * 8 * 8: three radix-8 passes with the
// eight points of a butterfly in registers,
//   n = 64 n1 + 8 n2 + n3,  k = k1 + 8 k2 + 64 k3:
//   A: thread (n2, n3)  DFT8 over n1, times e^{-2 pi i n2 k1 / 64}          -> T1[k1][n2][n3]
//   B: thread (k1, n3)  DFT8 over n2, times e^{-2 pi i n3 (k1 + 8 k2) / 512} -> T2[n3][k2][k1]
//   C: thread (k2, k1)  DFT8 over n3                                         -> Z[k1 + 8 k2 + 64 k3]
// i.e. two exchanges through shared memory (8-byte accesses, padded pitches: at most the natural two wavefronts per
// warp access) where the radix-2 version of round 1 made ten passes over separate re / im arrays -- that kernel was
// bound by the shared-memory pipe (ncu: l1tex data-pipe wavefronts 97 %), not by HBM.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// in-place 8-point DFT (e^{-2 pi i nk/8}), natural order in and out
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  constexpr float R = 0.70710678118654752f;
  float2 a[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { a[j] = cadd(v[j], v[j + 4]); b[j] = csub(v[j], v[j + 4]); }
  b[1] = make_float2(R * (b[1].x + b[1].y), R * (b[1].y - b[1].x));    // * (1 - i) / sqrt 2
  b[2] = cmul_mi(b[2]);                                                 // * (-i)
  b[3] = make_float2(R * (b[3].y - b[3].x), -R * (b[3].x + b[3].y));   // * (-1 - i) / sqrt 2
  auto dft4 = [](const float2 (&c)[4], float2& y0, float2& y1, float2& y2, float2& y3) {
    const float2 s0 = cadd(c[0], c[2]), s1 = csub(c[0], c[2]), s2 = cadd(c[1], c[3]), s3 = cmul_mi(csub(c[1], c[3]));
    y0 = cadd(s0, s2); y2 = csub(s0, s2); y1 = cadd(s1, s3); y3 = csub(s1, s3);
  };
  dft4(a, v[0], v[2], v[4], v[6]);
  dft4(b, v[1], v[3], v[5], v[7]);
}

constexpr int LM_P1 = 72, LM_P2 = 68;  // pitches (float2) of the two exchange buffers

__global__ void __launch_bounds__(128)
logmel_kernel(const float* __restrict__ audio, const int64_t* __restrict__ sample_off,
              const int64_t* __restrict__ frame_off, const float* __restrict__ window,
              const float2* __restrict__ twiddle, const int32_t* __restrict__ fb_start,
              const int32_t* __restrict__ fb_ptr, const float* __restrict__ fb_w,
              float* __restrict__ spect) {
  __shared__ float2 tw[512];                 // e^{-2 pi i j / 1024}, j < 512
  __shared__ float2 t1[2][8 * LM_P1];        // per frame: T1, later Z (512 entries)
  __shared__ float2 t2[2][8 * LM_P2];
  __shared__ float mag[2][516];
  const int clip = blockIdx.y;
  const int64_t f0 = frame_off[clip];
  const int T = static_cast<int>(frame_off[clip + 1] - f0);
  const int tid = threadIdx.x, half = tid >> 6, lt = tid & 63;
  const int t = 2 * blockIdx.x + half;
  if (2 * static_cast<int>(blockIdx.x) >= T) return;  // whole CTA beyond the clip
  const bool active = t < T;
  const int64_t s0 = sample_off[clip];
  const int64_t len = sample_off[clip + 1] - s0;
  for (int i = tid; i < 512; i += 128) tw[i] = twiddle[i];
  auto TW = [&](int j) -> float2 {  // e^{-2 pi i j / 1024}, 0 <= j < 1024
    const float2 w = tw[j & 511];
    return (j & 512) ? make_float2(-w.x, -w.y) : w;
  };
  float2 v[8];
  float2* T1 = t1[half];
  float2* T2 = t2[half];
  {  // ---- pass A: thread (n2, n3) = lt, points z[64 n1 + lt] ----
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const int n = 2 * (64 * n1 + lt);
      float xs[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int64_t i = 441ll * t + (n + e) - 512;
        if (i < 0) i = -i;                      // reflect (no edge repeat), torch pad_mode="reflect"
        if (i >= len) i = 2 * (len - 1) - i;
        xs[e] = active ? audio[s0 + i] * __ldg(window + n + e) : 0.f;
      }
      v[n1] = make_float2(xs[0], xs[1]);
    }
  }
  __syncthreads();  // twiddle table
  {
    dft8(v);
    const int n2 = lt >> 3;
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) T1[k1 * LM_P1 + lt] = k1 == 0 ? v[0] : cmul(v[k1], TW(16 * n2 * k1));
  }
  __syncthreads();
  {  // ---- pass B: thread (k1, n3) = lt ----
    const int k1 = lt >> 3, n3 = lt & 7;
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) v[n2] = T1[k1 * LM_P1 + n2 * 8 + n3];
    dft8(v);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) T2[n3 * LM_P2 + k2 * 8 + k1] = cmul(v[k2], TW(2 * n3 * (k1 + 8 * k2)));
  }
  __syncthreads();
  {  // ---- pass C: thread (k2, k1) = lt -> Z[lt + 64 k3] (into T1's storage) ----
#pragma unroll
    for (int n3 = 0; n3 < 8; ++n3) v[n3] = T2[n3 * LM_P2 + lt];
    dft8(v);
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) T1[lt + 64 * k3] = v[k3];
  }
  __syncthreads();
  // untangle: X[k] = E[k] + e^{-2 pi i k / 1024} O[k], E = (Z[k] + conj Z[512 - k]) / 2, O = -i (Z[k] - conj Z[512 - k]) / 2;
  // magnitudes of bins 0..512 (normalized=True -> 1 / sqrt(1024))
  for (int k = lt; k <= 512; k += 64) {
    const float2 zk = T1[k & 511], zc = T1[(512 - k) & 511];
    const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
    const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
    const float2 x = cadd(e, cmul(TW(k), o));
    mag[half][k] = sqrtf(x.x * x.x + x.y * x.y) * 0.03125f;
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int mm = 0; mm < 2; ++mm) {
      const int m = lt + 64 * mm;  // mel bin
      const int p0 = fb_ptr[m], p1 = fb_ptr[m + 1];
      const int k0 = fb_start[m];
      float acc = 0.f;
      for (int p = p0; p < p1; ++p) acc = fmaf(mag[half][k0 + (p - p0)], fb_w[p], acc);
      spect[(f0 + t) * 128 + m] = log1pf(1000.0f * acc);
    }
  }
}

void launch_logmel(const float* audio, const int64_t* sample_off_dev, const int64_t* frame_off_dev,
                   int n_clips, int64_t max_frames, const float* window, const float* twiddle,
                   const int32_t* fb_start, const int32_t* fb_ptr, const float* fb_w, float* spect,
                   cudaStream_t st) {
  if (max_frames <= 0 || n_clips <= 0) return;
  dim3 grid(static_cast<unsigned>((max_frames + 1) / 2), static_cast<unsigned>(n_clips));
  logmel_kernel<<<grid, 128, 0, st>>>(audio, sample_off_dev, frame_off_dev, window,
                                      reinterpret_cast<const float2*>(twiddle), fb_start, fb_ptr, fb_w, spect);
}

// ------------------------------------------------------------------------------------------
// Polyphase resampler to 22.05 kHz: device stand-in for soxr.resample (reference inference.py:274-275; method and
// filter design in beat_this_b200/preprocessing.py, parity with soxr unpinned).
//   y[n] = sum_k coef[(n M) mod L][k] * x[floor(n M / L) - K/2 + 1 + k],  zeros outside the clip.
// One CTA = 256 consecutive output samples of one clip; the input span they read is staged in shared memory.
// Algorithmic HBM bytes: 4 B per input sample + 4 B per output sample (the L x K bank stays in L1/L2).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ in, const int64_t* __restrict__ in_off, float* __restrict__ out,
                const int64_t* __restrict__ out_off, const float* __restrict__ coef, int L, int M, int K) {
  extern __shared__ float xs[];
  const int clip = blockIdx.y;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * 256;
  const int64_t s0 = in_off[clip], len = in_off[clip + 1] - s0;
  const int64_t o0 = out_off[clip], nout = out_off[clip + 1] - o0;
  if (n0 >= nout) return;
  const int64_t n_last = min(n0 + 255, nout - 1);
  const int64_t j_lo = (n0 * M) / L - K / 2 + 1;
  const int span = static_cast<int>((n_last * M) / L - K / 2 + K - j_lo + 1);
  for (int i = threadIdx.x; i < span; i += 256) {
    const int64_t j = j_lo + i;
    xs[i] = (j >= 0 && j < len) ? in[s0 + j] : 0.f;
  }
  __syncthreads();
  const int64_t n = n0 + threadIdx.x;
  if (n >= nout) return;
  const int64_t nm = n * M;
  const int base = static_cast<int>(nm / L - K / 2 + 1 - j_lo);
  const float* c = coef + static_cast<int64_t>(nm % L) * K;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = 0;
  for (; k + 4 <= K; k += 4) {
    a0 = fmaf(__ldg(c + k), xs[base + k], a0);
    a1 = fmaf(__ldg(c + k + 1), xs[base + k + 1], a1);
    a2 = fmaf(__ldg(c + k + 2), xs[base + k + 2], a2);
    a3 = fmaf(__ldg(c + k + 3), xs[base + k + 3], a3);
  }
  for (; k < K; ++k) a0 = fmaf(__ldg(c + k), xs[base + k], a0);
  out[o0 + n] = (a0 + a1) + (a2 + a3);
}

int launch_resample(const float* in, const int64_t* in_off_dev, float* out, const int64_t* out_off_dev, int n_clips,
                    int64_t max_out, const float* coef, int L, int M, int K, cudaStream_t st) {
  if (n_clips <= 0 || max_out <= 0) return 0;
  const int64_t span = (255ll * M) / L + K + 2;
  if (span * 4 > 200 * 1024) return -1;  // absurd ratio: the staged input span does not fit in shared memory
  const int smem = static_cast<int>(span * 4);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return -1;
  dim3 grid(static_cast<unsigned>((max_out + 255) / 256), static_cast<unsigned>(n_clips));
  resample_kernel<<<grid, 256, smem, st>>>(in, in_off_dev, out, out_off_dev, coef, L, M, K);
  return 0;
}

// ------------------------------------------------------------------------------------------
// stem: BN1d(128) -> Conv2d(1->32, k(4,3), s(4,1), p(0,1), no bias) -> BN2d -> GELU
// (reference beat_tracker.py:108-126).  BN2d is folded into w/bias on the host; BN1d cannot
// be folded (the conv's time padding is zero *after* BN1d) and is applied to each tap.
// Chunks are gathered straight from the per-clip spectrograms (split_piece/zeropad,
// inference.py:90-135): frames outside the clip are zero *before* BN1d.
// out: [B, 32 f, L, 32 c] fp32.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
stem_kernel(const float* __restrict__ spect, const ChunkSrc* __restrict__ chunks, int L,
            const float* __restrict__ bn1_scale, const float* __restrict__ bn1_shift,
            const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out) {
  __shared__ float ws[32 * 12];
  __shared__ float bs[32];
  for (int i = threadIdx.x; i < 32 * 12; i += 128) ws[i] = w[i];
  if (threadIdx.x < 32) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int t = blockIdx.x * 128 + threadIdx.x;
  const int f = blockIdx.y;
  const int b = blockIdx.z;
  if (t >= L) return;
  const ChunkSrc cs = chunks[b];
  float in[4][3];
#pragma unroll
  for (int dt = 0; dt < 3; ++dt) {
    const int tl = t + dt - 1;
    const bool conv_ok = tl >= 0 && tl < cs.len;  // zero padding of the convolution at the ends of THIS chunk
    const int64_t fr = static_cast<int64_t>(cs.start) + tl;
    const bool clip_ok = fr >= 0 && fr < cs.T;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (conv_ok && clip_ok)
      v = *reinterpret_cast<const float4*>(spect + (cs.frame_base + fr) * 128 + 4 * f);
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int df = 0; df < 4; ++df)
      in[df][dt] = conv_ok ? fmaf(vv[df], bn1_scale[4 * f + df], bn1_shift[4 * f + df]) : 0.f;
  }
  float* op = out + ((static_cast<int64_t>(b) * 32 + f) * L + t) * 32;
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = c4 * 4 + i;
      float a = bs[co];
#pragma unroll
      for (int df = 0; df < 4; ++df)
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) a = fmaf(in[df][dt], ws[co * 12 + df * 3 + dt], a);
      r[i] = gelu_fast(a);  // erf to 1.5e-7 on rcp + ex2 (erff: ~35 instructions, 32 of them per thread here)
    }
    reinterpret_cast<float4*>(op)[c4] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

__global__ void __launch_bounds__(256)
zero_tail_kernel(uint4* __restrict__ buf, const ChunkSrc* __restrict__ chunks, int F, int L, int row_vec) {
  const int plane = blockIdx.x;
  const int len = chunks[plane / F].len;
  const int64_t n = static_cast<int64_t>(L - len) * row_vec;  // 16-byte vectors to clear
  uint4* p = buf + (static_cast<int64_t>(plane) * L + len) * row_vec;
  for (int64_t i = threadIdx.x; i < n; i += 256) p[i] = make_uint4(0u, 0u, 0u, 0u);
}
void launch_zero_tail(void* buf, int elem_bytes, const ChunkSrc* chunks, int nchunks, int F, int L, int C, cudaStream_t st) {
  zero_tail_kernel<<<nchunks * F, 256, 0, st>>>(reinterpret_cast<uint4*>(buf), chunks, F, L, C * elem_bytes / 16);
}

void launch_stem(const float* spect, const ChunkSrc* chunks, int nchunks, int L, const float* bn1_scale,
                 const float* bn1_shift, const float* w, const float* bias, float* out,
                 cudaStream_t st) {
  dim3 grid(ceil_div(L, 128), 32, nchunks);
  stem_kernel<<<grid, 128, 0, st>>>(spect, chunks, L, bn1_scale, bn1_shift, w, bias, out);
}

// ------------------------------------------------------------------------------------------
// RMSNorm (reference roformer.py:22-32: x / max(||x||, 1e-12) * sqrt(dim) * gamma; the
// sqrt(dim)*gamma factor is folded into the consuming weights).  Pure HBM streaming:
// 4 B read + sizeof(TAct) B written per element.  A row is handled by C/4 (<= 32) lanes with
// float4 loads; several rows share a warp when C < 128.
// ------------------------------------------------------------------------------------------
template <typename TAct, int C>
__global__ void __launch_bounds__(256)
norm_kernel(const float* __restrict__ x, TAct* __restrict__ xn, int64_t M, float* __restrict__ gates,
            const float* __restrict__ wg, const float* __restrict__ bg, int heads) {
  constexpr int LPR = C / 4 < 32 ? C / 4 : 32;  // lanes per row
  constexpr int RPW = 32 / LPR;                 // rows per warp
  constexpr int VPL = C / 4 / LPR;              // float4 per lane
  const int lane = threadIdx.x & 31;
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t row = warp * RPW + lane / LPR;
  const int li = lane % LPR;
  const bool ok = row < M;
  const float4* xr = reinterpret_cast<const float4*>(x + (ok ? row : 0) * C);
  float4 v[VPL];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[li + LPR * i];
    ss = fmaf(v[i].x, v[i].x, ss); ss = fmaf(v[i].y, v[i].y, ss);
    ss = fmaf(v[i].z, v[i].z, ss); ss = fmaf(v[i].w, v[i].w, ss);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
  for (int i = 0; i < VPL; ++i) { v[i].x *= inv; v[i].y *= inv; v[i].z *= inv; v[i].w *= inv; }
  if (gates) {
    // attention gates sigmoid(to_gates(x_normed)) (reference roformer.py:127-128) for the few-head
    // frontend attentions (1, 2 or 4 heads): a handful of FMAs per row inside this HBM-bound kernel
    // instead of a separate padded-N GEMM launch over the same rows.
    for (int h = 0; h < heads; ++h) {
      const float4* w4 = reinterpret_cast<const float4*>(wg + h * C);
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float4 w = __ldg(w4 + li + LPR * i);
        a = fmaf(v[i].x, w.x, a); a = fmaf(v[i].y, w.y, a); a = fmaf(v[i].z, w.z, a); a = fmaf(v[i].w, w.w, a);
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (ok && li == 0) gates[row * heads + h] = sigmoidf_(a + __ldg(bg + h));
    }
  }
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    TAct* dst = xn + row * C + 4 * (li + LPR * i);
    if constexpr (sizeof(TAct) == 4) {
      *reinterpret_cast<float4*>(dst) = v[i];
    } else {
      uint2 u;
      u.x = pack_h16x2(v[i].x, v[i].y);
      u.y = pack_h16x2(v[i].z, v[i].w);
      *reinterpret_cast<uint2*>(dst) = u;
    }
  }
}

template <typename TAct>
static void norm_dispatch(const float* x, void* xn, int64_t M, int C, float* gates, const float* wg, const float* bg,
                          int heads, cudaStream_t st) {
  TAct* o = reinterpret_cast<TAct*>(xn);
#define BT_NORM_CASE(c)                                                                                   \
  case c: {                                                                                               \
    constexpr int rpw = (c / 4 < 32) ? 32 / (c / 4) : 1;                                                  \
    norm_kernel<TAct, c><<<static_cast<unsigned>(ceil_div64(M, 8 * rpw)), 256, 0, st>>>(x, o, M, gates, wg, \
                                                                                          bg, heads);     \
  } break;
  switch (C) {
    BT_NORM_CASE(32) BT_NORM_CASE(64) BT_NORM_CASE(128) BT_NORM_CASE(256) BT_NORM_CASE(512) BT_NORM_CASE(1024)
    default: break;  // validated in bt_create
  }
#undef BT_NORM_CASE
}

void launch_norm(const float* x, void* xn, int64_t M, int C, int act_h16, cudaStream_t st, float* gates,
                 const float* wg, const float* bg, int heads) {
  if (act_h16) norm_dispatch<h16>(x, xn, M, C, gates, wg, bg, heads, st);
  else norm_dispatch<float>(x, xn, M, C, gates, wg, bg, heads, st);
}

// ------------------------------------------------------------------------------------------
// frequency-direction attention (PartialFTTransformer attnF, reference
// beat_tracker.py:292-294): sequences of F in {32,16,8} tokens over the frequency axis for
// every (chunk, frame, head).  Token m = (b*F + f)*L + t.  A warp handles 32/F groups
// (b,t,h) at once: lane -> (group g = lane / F, token f = lane % F).  Each lane loads its own
// q/k/v rows with 16-byte loads, K/V are shared through padded shared memory (float4
// broadcast reads).  HBM bytes: 3C in + C out per token (activation dtype).
// ------------------------------------------------------------------------------------------
template <typename TAct>
__device__ __forceinline__ void load_row32(const TAct* p, float (&v)[32]);
template <>
__device__ __forceinline__ void load_row32<float>(const float* p, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 q = reinterpret_cast<const float4*>(p)[i];
    v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
  }
}
template <>
__device__ __forceinline__ void load_row32<h16>(const h16* p, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 q = reinterpret_cast<const uint4*>(p)[i];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) unpack_h16x2(w[j], v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
  }
}

template <typename TAct, int F>
__global__ void __launch_bounds__(128, 4)
attn_freq_kernel(const TAct* __restrict__ qkv, const float* __restrict__ gates, TAct* __restrict__ out,
                 int B, int L, int heads, float scale) {
  constexpr int GPW = 32 / F;
  constexpr int RS = 36;               // padded row stride (floats)
  constexpr int GS = F * RS + 4;       // group stride: skews groups onto different banks
  __shared__ __align__(16) float Ks[4][GPW * GS];
  __shared__ __align__(16) float Vs[4][GPW * GS];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / F, f = lane % F;
  const int64_t ngrp = static_cast<int64_t>(B) * L * heads;
  const int64_t grp0 = (static_cast<int64_t>(blockIdx.x) * 4 + wib) * GPW;
  if (grp0 >= ngrp) return;  // warp-uniform
  const int64_t grp = grp0 + g;
  const bool act = grp < ngrp;
  const int64_t gg = act ? grp : grp0;
  const int h = static_cast<int>(gg % heads);
  const int64_t bt_ = gg / heads;
  const int t = static_cast<int>(bt_ % L);
  const int b = static_cast<int>(bt_ / L);
  const int C = heads * 32;
  const int64_t m = (static_cast<int64_t>(b) * F + f) * L + t;
  const TAct* rp = qkv + m * 3 * C + h * 32;
  float q[32];
  {
    float kv[32];
    load_row32<TAct>(rp + C, kv);
    float* kd = &Ks[wib][g * GS + f * RS];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      reinterpret_cast<float4*>(kd)[i] = make_float4(kv[4 * i], kv[4 * i + 1], kv[4 * i + 2], kv[4 * i + 3]);
    load_row32<TAct>(rp + 2 * C, kv);
    float* vd = &Vs[wib][g * GS + f * RS];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      reinterpret_cast<float4*>(vd)[i] = make_float4(kv[4 * i], kv[4 * i + 1], kv[4 * i + 2], kv[4 * i + 3]);
    load_row32<TAct>(rp, q);
#pragma unroll
    for (int d = 0; d < 32; ++d) q[d] *= scale;
  }
  __syncwarp();
  const float* kb = &Ks[wib][g * GS];
  const float* vb = &Vs[wib][g * GS];
  float s[F];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < F; ++j) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 k4 = reinterpret_cast<const float4*>(kb + j * RS)[i];
      a = fmaf(q[4 * i], k4.x, a); a = fmaf(q[4 * i + 1], k4.y, a);
      a = fmaf(q[4 * i + 2], k4.z, a); a = fmaf(q[4 * i + 3], k4.w, a);
    }
    s[j] = a;
    mx = fmaxf(mx, a);
    asm volatile("" ::: "memory");  // keep ptxas from hoisting every K row into registers (spills)
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < F; ++j) {
    s[j] = __expf(s[j] - mx);
    l += s[j];
  }
  float o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < F; ++j) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v4 = reinterpret_cast<const float4*>(vb + j * RS)[i];
      o[4 * i] = fmaf(s[j], v4.x, o[4 * i]); o[4 * i + 1] = fmaf(s[j], v4.y, o[4 * i + 1]);
      o[4 * i + 2] = fmaf(s[j], v4.z, o[4 * i + 2]); o[4 * i + 3] = fmaf(s[j], v4.w, o[4 * i + 3]);
    }
    asm volatile("" ::: "memory");
  }
  if (act) {
    const float gsc = gates[m * heads + h] / l;
    TAct* op = out + m * C + h * 32;
    if constexpr (sizeof(TAct) == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(op)[i] = make_float4(o[4 * i] * gsc, o[4 * i + 1] * gsc, o[4 * i + 2] * gsc, o[4 * i + 3] * gsc);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_h16x2(o[8 * i] * gsc, o[8 * i + 1] * gsc); u.y = pack_h16x2(o[8 * i + 2] * gsc, o[8 * i + 3] * gsc);
        u.z = pack_h16x2(o[8 * i + 4] * gsc, o[8 * i + 5] * gsc); u.w = pack_h16x2(o[8 * i + 6] * gsc, o[8 * i + 7] * gsc);
        reinterpret_cast<uint4*>(op)[i] = u;
      }
    }
  }
}

// h16 path: the same attention on warp-level tensor-core MMAs (mma.sync m16n8k16, fp32 accumulate).
// A warp stages 32 rows (32/F groups: q | k | v, 64 bytes each) in shared memory with the coalesced
// row-per-lane loads of the SIMT kernel, then works on two 16-row query tiles: S = Q K^T from
// ldmatrix fragments, softmax in the accumulator layout (row reductions over the 4 lanes of a quad),
// P re-packed in registers as the A operand of P V (V through ldmatrix.trans).  F = 32: a tile sees
// all 32 keys; F = 16: a tile is one group; F = 8: a tile holds two groups, the cross blocks are
// masked.  ~40 tensor instructions per warp instead of ~4000 FMAs per lane.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_h16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." BT_H16_MMA_SYNC ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Tile = FT_TT consecutive frames of one chunk, all F frequency planes, all heads: ONE TMA box per (q|k|v, head)
// brings [TT][F][32] fp16 into shared memory (SWIZZLE_64B; the tensor map lists the plane dimension before the frame
// dimension, so the F rows one attention group reads are CONSECUTIVE 64-byte rows and ldmatrix is conflict free --
// with [F][TT] order the eight rows of an ldmatrix phase were 256 B apart: 4-way conflicts, L1 data pipe 82 % in
// ncu) and one box stores the [TT][F][C] output tile.  (Before: every lane fetched its own
// row, L * 3C elements away from its neighbour's -- 32 lines per ld.global, L1 wavefronts 73-87 % in ncu.)
constexpr int FT_TT = 4;

template <int F>
__global__ void __launch_bounds__(128)
attn_freq_mma_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut,
                     const float* __restrict__ gates, int L, int heads, float scale_log2) {
  constexpr int GPW = 32 / F;             // groups (frame, head) per warp
  constexpr int NT = F == 32 ? 4 : 2;     // 8-key tiles a query tile attends to
  constexpr int HEADS = 4 * GPW / FT_TT;  // 1, 2, 4 for F = 32, 16, 8: the four warps cover TT frames x HEADS heads
  constexpr int C = HEADS * 32;
  constexpr int PH_BYTES = F * FT_TT * 64;  // one (part, head) box
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sIn = sbase;                                  // [3 parts][HEADS][F * TT rows][64 B]
  const uint32_t sOut = sIn + 3 * HEADS * PH_BYTES;            // [F * TT rows][C * 2 B], dense
  const uint32_t bar = sOut + F * FT_TT * C * 2;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * FT_TT, b = blockIdx.y;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    fence_barrier_init();
    mbar_expect_tx_a(bar, 3 * HEADS * PH_BYTES);
    for (int part = 0; part < 3; ++part)
      for (int h = 0; h < HEADS; ++h)
        tma_load_3d_a(sIn + (part * HEADS + h) * PH_BYTES, &tmIn, bar, part * C + h * 32, b * F, t0);
  }
  __syncthreads();
  mbar_wait_a(bar, 0);
  // this warp's groups: head h, frames tt_base .. tt_base + GPW - 1; "staged row" r = gl * F + f as before
  const int h = (wib * GPW) / FT_TT;
  const int tt_base = (wib * GPW) % FT_TT;
  auto row_addr = [&](uint32_t base, int r, int chunk) -> uint32_t {  // 16-byte chunk `chunk` of staged row r
    const int row = tt_base * F + r;  // = (tt_base + r / F) * F + r % F
    return base + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
  };
  const uint32_t sQ = sIn + (0 * HEADS + h) * PH_BYTES, sK = sIn + (1 * HEADS + h) * PH_BYTES, sV = sIn + (2 * HEADS + h) * PH_BYTES;
  const int g = lane >> 2, c = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int key_base = F == 32 ? 0 : 16 * mt;
    uint32_t qa[2][4];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
      ldsm_x4(row_addr(sQ, 16 * mt + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4)), qa[kk]);
    float sc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
      uint32_t kb[4];
      ldsm_x4(row_addr(sK, key_base + 8 * j + (lane & 7), lane >> 3), kb);
      mma_h16_16816(sc[j], qa[0], kb[0], kb[1]);
      mma_h16_16816(sc[j], qa[1], kb[2], kb[3]);
    }
    if (F == 8) {  // rows 0-7 belong to the tile's first group (keys of tile 0), rows 8-15 to the second
      sc[1][0] = sc[1][1] = -INFINITY;
      sc[0][2] = sc[0][3] = -INFINITY;
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      sc[j][0] = exp2f((sc[j][0] - mx0) * scale_log2); sc[j][1] = exp2f((sc[j][1] - mx0) * scale_log2);
      sc[j][2] = exp2f((sc[j][2] - mx1) * scale_log2); sc[j][3] = exp2f((sc[j][3] - mx1) * scale_log2);
      l0 += sc[j][0] + sc[j][1];
      l1 += sc[j][2] + sc[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float o[4][4];
#pragma unroll
    for (int jd = 0; jd < 4; ++jd) o[jd][0] = o[jd][1] = o[jd][2] = o[jd][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_h16x2(sc[2 * kk][0], sc[2 * kk][1]);
      pa[1] = pack_h16x2(sc[2 * kk][2], sc[2 * kk][3]);
      pa[2] = pack_h16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
      pa[3] = pack_h16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
      for (int jd = 0; jd < 4; jd += 2) {
        uint32_t vb[4];
        const int q4 = lane >> 3;
        ldsm_x4_trans(row_addr(sV, key_base + 16 * kk + (q4 & 1) * 8 + (lane & 7), jd + (q4 >> 1)), vb);
        mma_h16_16816(o[jd], pa, vb[0], vb[1]);
        mma_h16_16816(o[jd + 1], pa, vb[2], vb[3]);
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = 16 * mt + g + 8 * half;
      const int f = r % F, tt = tt_base + r / F;
      const int t = t0 + tt;
      const int64_t m = (static_cast<int64_t>(b) * F + f) * L + (t < L ? t : L - 1);
      const float gsc = gates[m * HEADS + h] / (half == 0 ? l0 : l1);
      const uint32_t orow = sOut + (tt * F + f) * (C * 2) + h * 64;
#pragma unroll
      for (int jd = 0; jd < 4; ++jd)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(orow + (4 * jd + c) * 4), "r"(pack_h16x2(o[jd][2 * half] * gsc, o[jd][2 * half + 1] * gsc)) : "memory");
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_3d(&tmOut, sOut, 0, b * F, t0);  // frames beyond L are clipped
    bulk_commit();
    bulk_wait_read<0>();
  }
}

template <int F>
static int attn_freq_mma_launch(const void* qkv, const float* gates, void* out, int B, int L, float sl2, cudaStream_t st) {
  constexpr int HEADS = 4 * (32 / F) / FT_TT;
  constexpr int C = HEADS * 32;
  constexpr int SMEM = 3 * HEADS * F * FT_TT * 64 + F * FT_TT * C * 2 + 1024 + 64;
  // tensor maps over the activation buffers, cached per (buffers, geometry)
  struct Key { const void *q, *o; int B, L; CUtensorMap in, outm; };
  static Key cache[4];
  static int n_cached = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  Key* k = nullptr;
  for (int i = 0; i < n_cached; ++i)
    if (cache[i].q == qkv && cache[i].o == out && cache[i].B == B && cache[i].L == L) k = &cache[i];
  if (!k) {
    k = &cache[n_cached < 4 ? n_cached++ : 0];
    char err[256];
    // dimension order (channels, planes, frames): the plane stride is the larger one
    const uint64_t din[3] = {static_cast<uint64_t>(3 * C), static_cast<uint64_t>(B) * F, static_cast<uint64_t>(L)};
    const uint64_t sin_[2] = {static_cast<uint64_t>(L) * 3 * C * 2, static_cast<uint64_t>(3 * C) * 2};
    const uint32_t bin[3] = {32, F, FT_TT};
    const uint64_t dout[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(B) * F, static_cast<uint64_t>(L)};
    const uint64_t sout[2] = {static_cast<uint64_t>(L) * C * 2, static_cast<uint64_t>(C) * 2};
    const uint32_t bout[3] = {static_cast<uint32_t>(C), F, FT_TT};
    if (!make_tmap(&k->in, qkv, 3, din, sin_, bin, 64, err, sizeof(err)) || !make_tmap(&k->outm, out, 3, dout, sout, bout, 0, err, sizeof(err))) {
      fprintf(stderr, "bt: attn_freq tensor map: %s -- using the scalar kernel\n", err);  // never seen; loud if it happens
      k->q = nullptr;
      return -1;
    }
    k->q = qkv; k->o = out; k->B = B; k->L = L;
  }
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(attn_freq_mma_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); attr = true; }
  dim3 grid(ceil_div(L, FT_TT), B);
  attn_freq_mma_kernel<F><<<grid, 128, SMEM, st>>>(k->in, k->outm, gates, L, HEADS, sl2);
  return 0;
}

template <typename TAct>
static void attn_freq_dispatch(const void* qkv, const float* gates, void* out, int B, int F, int L, int heads,
                               float scale, cudaStream_t st) {
  const int64_t ngrp = static_cast<int64_t>(B) * L * heads;
  const TAct* q = reinterpret_cast<const TAct*>(qkv);
  TAct* o = reinterpret_cast<TAct*>(out);
  const unsigned grid = static_cast<unsigned>(ceil_div64(ngrp, 4 * (32 / F)));
  if (F == 32) attn_freq_kernel<TAct, 32><<<grid, 128, 0, st>>>(q, gates, o, B, L, heads, scale);
  else if (F == 16) attn_freq_kernel<TAct, 16><<<grid, 128, 0, st>>>(q, gates, o, B, L, heads, scale);
  else attn_freq_kernel<TAct, 8><<<grid, 128, 0, st>>>(q, gates, o, B, L, heads, scale);
}

void launch_attn_freq(const void* qkv, const float* gates, void* out, int B, int F, int L, int heads,
                      float scale, int act_h16, cudaStream_t st) {
  static const bool simt = getenv("BT_ATTN_FREQ_SIMT") && atoi(getenv("BT_ATTN_FREQ_SIMT")) != 0;
  if (act_h16 && !simt && (F == 32 || F == 16 || F == 8)) {
    const float sl2 = scale * 1.4426950408889634f;
    int rc = -1;
    if (F == 32 && heads == 1) rc = attn_freq_mma_launch<32>(qkv, gates, out, B, L, sl2, st);
    else if (F == 16 && heads == 2) rc = attn_freq_mma_launch<16>(qkv, gates, out, B, L, sl2, st);
    else if (F == 8 && heads == 4) rc = attn_freq_mma_launch<8>(qkv, gates, out, B, L, sl2, st);
    if (rc == 0) return;
  }
  if (act_h16) attn_freq_dispatch<h16>(qkv, gates, out, B, F, L, heads, scale, st);
  else attn_freq_dispatch<float>(qkv, gates, out, B, F, L, heads, scale, st);
}

// ------------------------------------------------------------------------------------------
// head: final RMSNorm (roformer.py:174,180; gamma*sqrt(D) folded into w) -> Linear(D->2) ->
// SumHead (beat = o0 + o1 in fp32, downbeat = o1; beat_tracker.py:315-330) -> scatter into
// the per-clip frame arrays with aggregate_prediction's keep_first rule
// (inference.py:138-185): each chunk owns the chunk-local frames [write_lo, write_hi).
// One warp per token.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ x, int D, const float* __restrict__ w, const float* __restrict__ bias,
            const ChunkSrc* __restrict__ chunks, int nchunks, int L, float* __restrict__ beat,
            float* __restrict__ down, int sum_head) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<int64_t>(nchunks) * L) return;
  const int b = static_cast<int>(row / L), t = static_cast<int>(row % L);
  const ChunkSrc cs = chunks[b];
  if (t < cs.write_lo || t >= cs.write_hi) return;  // warp-uniform
  const float* xr = x + row * D;
  float ss = 0.f, a0 = 0.f, a1 = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float v = xr[i];
    ss = fmaf(v, v, ss);
    a0 = fmaf(v, __ldg(w + i), a0);
    a1 = fmaf(v, __ldg(w + D + i), a1);
  }
  ss = warp_sum(ss); a0 = warp_sum(a0); a1 = warp_sum(a1);
  if (lane == 0) {
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    const float o0 = a0 * inv + bias[0], o1 = a1 * inv + bias[1];
    const int64_t fr = cs.out_base + cs.start + t;
    beat[fr] = sum_head ? o0 + o1 : o0;  // SumHead (beat_tracker.py:315-330) / Head (beat_tracker.py:333-346)
    down[fr] = o1;
  }
}

void launch_head(const float* x, int D, const float* w, const float* b, const ChunkSrc* chunks,
                 int nchunks, int L, float* beat, float* down, int sum_head, cudaStream_t st) {
  const int64_t rows = static_cast<int64_t>(nchunks) * L;
  head_kernel<<<static_cast<unsigned>(ceil_div64(rows, 8)), 256, 0, st>>>(x, D, w, b, chunks, nchunks, L,
                                                                           beat, down, sum_head);
}

// ------------------------------------------------------------------------------------------
// minimal postprocessor (reference model/postprocessor.py:85-136, deduplicate_peaks
// :176-197): peak <=> x[t] == max(x[t-3..t+3]) and x[t] > 0; runs of peaks at most one
// frame from the running mean are merged into the running mean (float64, like the Python
// loop); times = frame / 50; every downbeat snaps to the nearest beat (first argmin); unique.
// One CTA per clip: ordered compaction by ballot/prefix, then the short sequential part.
// ------------------------------------------------------------------------------------------
__device__ int compact_peaks(const float* __restrict__ x, int T, double* __restrict__ frames, int cap,
                             int* s_warp, int* s_base) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) *s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < T; t0 += 256) {
    const int t = t0 + tid;
    bool pk = false;
    if (t < T) {
      const float v = x[t];
      float mx = v;
#pragma unroll
      for (int d = -3; d <= 3; ++d) {
        const int u = t + d;
        if (u >= 0 && u < T) mx = fmaxf(mx, x[u]);
      }
      pk = (v == mx) && (v > 0.f);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pk);
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int off = *s_base;
    for (int w = 0; w < wid; ++w) off += s_warp[w];
    if (pk) {
      const int idx = off + __popc(bal & ((1u << lane) - 1));
      if (idx < cap) frames[idx] = static_cast<double>(t);
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += s_warp[w];
      *s_base += tot;
    }
    __syncthreads();
  }
  return *s_base;
}

__device__ int dedup_to_times(double* p, int n) {
  // deduplicate_peaks(width=1) followed by / fps; in place (output index <= input index)
  if (n == 0) return 0;
  int out = 0;
  double cur = p[0];
  double c = 1.0;
  for (int i = 1; i < n; ++i) {
    const double p2 = p[i];
    if (p2 - cur <= 1.0) {
      c += 1.0;
      cur += (p2 - cur) / c;
    } else {
      p[out++] = cur / 50.0;
      cur = p2;
      c = 1.0;
    }
  }
  p[out++] = cur / 50.0;
  return out;
}

__global__ void __launch_bounds__(256)
peakpick_kernel(const float* __restrict__ beat, const float* __restrict__ down,
                const int64_t* __restrict__ frame_off, double* __restrict__ beat_t, int32_t* __restrict__ n_beat,
                double* __restrict__ down_t, int32_t* __restrict__ n_down, int max_peaks) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int clip = blockIdx.x;
  const int64_t f0 = frame_off[clip];
  const int T = static_cast<int>(frame_off[clip + 1] - f0);
  double* bt_ = beat_t + static_cast<int64_t>(clip) * max_peaks;
  double* dt_ = down_t + static_cast<int64_t>(clip) * max_peaks;
  const int nb_raw = compact_peaks(beat + f0, T, bt_, max_peaks, s_warp, &s_base);
  __syncthreads();
  const int nd_raw = compact_peaks(down + f0, T, dt_, max_peaks, s_warp, &s_base);
  __syncthreads();
  __shared__ int s_nb, s_nd;
  if (threadIdx.x == 0) {
    if (nb_raw > max_peaks || nd_raw > max_peaks) {  // overflow: report the true counts, keep the first max_peaks
      n_beat[clip] = nb_raw;
      n_down[clip] = nd_raw;
      s_nb = -1;
    } else {
      s_nb = dedup_to_times(bt_, nb_raw);
      s_nd = dedup_to_times(dt_, nd_raw);
    }
  }
  __syncthreads();
  const int nb = s_nb;
  if (nb < 0) return;
  int nd = s_nd;
  // every downbeat moves to the nearest beat time (first minimum, postprocessor.py:128-134): one downbeat per thread
  if (nb > 0) {
    for (int i = threadIdx.x; i < nd; i += blockDim.x) {
      const double d = dt_[i];
      int best = 0;
      double bd = fabs(bt_[0] - d);
      for (int j = 1; j < nb; ++j) {
        const double dd = fabs(bt_[j] - d);
        if (dd < bd) { bd = dd; best = j; }
      }
      dt_[i] = bt_[best];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // np.unique: sort + drop duplicates (snapped downbeats are non-decreasing; insertion sort is a no-op then)
    for (int i = 1; i < nd; ++i) {
      const double v = dt_[i];
      int j = i - 1;
      while (j >= 0 && dt_[j] > v) { dt_[j + 1] = dt_[j]; --j; }
      dt_[j + 1] = v;
    }
    int o = 0;
    for (int i = 0; i < nd; ++i)
      if (o == 0 || dt_[i] != dt_[o - 1]) dt_[o++] = dt_[i];
    nd = o;
    n_beat[clip] = nb;
    n_down[clip] = nd;
  }
}

void launch_peakpick(const float* beat, const float* down, const int64_t* frame_off_dev, int n_clips,
                     double* beat_t, int32_t* n_beat, double* down_t, int32_t* n_down, int max_peaks,
                     cudaStream_t st) {
  if (n_clips <= 0) return;
  peakpick_kernel<<<n_clips, 256, 0, st>>>(beat, down, frame_off_dev, beat_t, n_beat, down_t, n_down,
                                           max_peaks);
}

// ------------------------------------------------------------------------------------------ utils
__global__ void f32_to_h16_kernel(const float* __restrict__ in, h16* __restrict__ out, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = to_out<h16>(in[i]);
}
__global__ void h16_to_f32_kernel(const h16* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = to_f32(in[i]);
}
void launch_f32_to_h16(const float* in, void* out, int64_t n, cudaStream_t st) {
  if (n <= 0) return;
  f32_to_h16_kernel<<<static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, st>>>(in, reinterpret_cast<h16*>(out), n);
}
void launch_h16_to_f32(const void* in, float* out, int64_t n, cudaStream_t st) {
  if (n <= 0) return;
  h16_to_f32_kernel<<<static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, st>>>(reinterpret_cast<const h16*>(in), out, n);
}

template <typename TAct>
__global__ void pack_qkv_test_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                     const float* __restrict__ v, TAct* __restrict__ qkv, int seqs, int L,
                                     int heads, float qscale) {
  const int C = heads * 32;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n = static_cast<int64_t>(seqs) * L * C;
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  const int64_t m = i / C;
  qkv[m * 3 * C + c] = to_out<TAct>(q[i] * qscale);
  qkv[m * 3 * C + C + c] = to_out<TAct>(k[i]);
  qkv[m * 3 * C + 2 * C + c] = to_out<TAct>(v[i]);
}
void launch_pack_qkv_test(const float* q, const float* k, const float* v, void* qkv, int seqs, int L, int heads,
                          float qscale, int act_h16, cudaStream_t st) {
  const int64_t n = static_cast<int64_t>(seqs) * L * heads * 32;
  const unsigned grid = static_cast<unsigned>(ceil_div64(n, 256));
  if (act_h16)
    pack_qkv_test_kernel<h16><<<grid, 256, 0, st>>>(q, k, v, reinterpret_cast<h16*>(qkv), seqs, L, heads, qscale);
  else
    pack_qkv_test_kernel<float><<<grid, 256, 0, st>>>(q, k, v, reinterpret_cast<float*>(qkv), seqs, L, heads, qscale);
}

}  // namespace bt
