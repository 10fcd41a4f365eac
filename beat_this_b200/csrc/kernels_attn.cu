// tcgen05 flash attention for head_dim 32 over sequences of up to 1500 frames: the time-direction
// attention of the three frontend blocks and of the 6 main layers (reference roformer.py:73-80 SDPA,
// called from roformer.py:114-132 / beat_tracker.py:290-301).  P and O live in tensor memory.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace bt {

constexpr int AT_BQ = 128;
constexpr int AT_SQ = 8192;
constexpr float AT_TAU = 8.0f;  // log2 units: rescale O only when the row maximum grew by more than this

// Same algorithm with a smaller footprint per CTA so that FOUR CTAs share an SM (the 2-CTA kernel is
// bound by each CTA's own dependency chain: while its softmax warps wait for S / synchronise, only ONE
// other CTA is there to keep the MUFU pipe busy).  CTA = 128 queries, 64-key tiles, 4 softmax warps
// (ONE thread per query row: no partial-maximum exchange) + 1 issuer warp, 128 TMEM columns:
// S [0,64) | O [64,96) | P [96,128).  No ones block (it would need 48 accumulator columns): the row
// sums are accumulated in registers.
constexpr int A6_BKV = 64, A6_NST = 4;
constexpr int A6_SK = 4096, A6_SV = 4096;
constexpr int A6_SMEM = AT_SQ + A6_NST * (A6_SK + A6_SV) + 1024 + 128;
constexpr int A6_THREADS = 160;
constexpr uint32_t A6_TM_O = 64, A6_TM_P = 96;
constexpr int A6_DEFAULT_PP = 0;  // measured (profiles/r2_notes.md): every polynomial share is slower than 0

// which of every 8 score pairs take the polynomial exp2 (spread out so that MUFU and FMA work interleave)
__host__ __device__ constexpr uint32_t attn_poly_mask(int pp) {
  return pp == 0 ? 0x00u : pp == 1 ? 0x08u : pp == 2 ? 0x44u : pp == 3 ? 0x52u : pp == 4 ? 0xAAu : pp == 5 ? 0xB5u : pp == 6 ? 0xBBu : 0xFFu;
}

template <int PP>  // PP of every 8 score pairs: exp2 on the FMA pipe (packed polynomial); the rest on MUFU
__global__ void __launch_bounds__(A6_THREADS, 4)
attn_tc64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const float* __restrict__ gates, h16* __restrict__ out, int L, int heads,
                 const ChunkSrc* __restrict__ chunks, int seqs_per_chunk) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = sbase;
  const uint32_t sK = sQ + AT_SQ;             // [A6_NST] 64 keys x 32 dims, SW64 K-major
  const uint32_t sV = sK + A6_NST * A6_SK;    // [A6_NST] 64 keys x 32 dims, used as MN-major B operand
  const uint32_t bar_q = sV + A6_NST * A6_SV;
  const uint32_t bar_kv = bar_q + 8;          // [A6_NST]
  const uint32_t bar_s = bar_kv + 8 * A6_NST;
  const uint32_t bar_sfree = bar_s + 8;
  const uint32_t bar_p = bar_sfree + 8;
  const uint32_t bar_pv = bar_p + 8;          // [2]
  const uint32_t tmem_slot = bar_pv + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ;
  const int h = blockIdx.y;
  const int seq = blockIdx.z;
  const int C = heads * 32;
  // keys that exist for this sequence: the whole plane, or (waves of chunks of different lengths) its chunk's frames
  const int Lk = chunks ? chunks[seq / seqs_per_chunk].len : L;
  const int nkv = ceil_div(Lk, A6_BKV);
  constexpr int MMA_WARP = 4;
  constexpr int NSOFT = 128;
  constexpr uint32_t POLY_MASK = attn_poly_mask(PP);

  if (warp == MMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    auto init = [](uint32_t bar, uint32_t count) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    };
    init(bar_q, 1);
    for (int i = 0; i < A6_NST; ++i) init(bar_kv + 8 * i, 1);
    init(bar_s, 1);
    init(bar_sfree, NSOFT);
    init(bar_p, NSOFT);
    init(bar_pv, 1); init(bar_pv + 8, 1);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == MMA_WARP) {
    const uint32_t on = elect_one() ? 1u : 0u;  // converged issuer warp, see attn_tc_kernel
    constexpr uint32_t idesc_s = make_idesc_h16(128, 64);
    constexpr uint32_t idesc_o = make_idesc_h16(128, 32) | (1u << 16);  // bit 16: B is MN-major
    auto load_kv = [&](int j) {
      const int st = j % A6_NST;
      mbar_expect_tx_p(on, bar_kv + 8 * st, A6_SK + A6_SV);
      tma_load_3d_p(on, sK + st * A6_SK, &tmKV, bar_kv + 8 * st, C + h * 32, j * A6_BKV, seq);
      tma_load_3d_p(on, sV + st * A6_SV, &tmKV, bar_kv + 8 * st, 2 * C + h * 32, j * A6_BKV, seq);
    };
    auto issue_s = [&](int j) {
      const uint32_t kb = sK + (j % A6_NST) * A6_SK;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        umma_h16_p(on, tmem_base, make_kmajor_desc<64>(sQ + k * 32), make_kmajor_desc<64>(kb + k * 32), idesc_s,
                    k != 0 ? 1u : 0u);
      umma_commit_p(on, bar_s);
    };
    mbar_expect_tx_p(on, bar_q, AT_SQ);
    tma_load_3d_p(on, sQ, &tmQ, bar_q, h * 32, q0, seq);
    for (int j = 0; j < A6_NST && j < nkv; ++j) load_kv(j);
    mbar_wait_a(bar_q, 0);
    mbar_wait_a(bar_kv, 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) {
        mbar_wait_a(bar_sfree, j & 1);
        mbar_wait_a(bar_kv + 8 * ((j + 1) % A6_NST), ((j + 1) / A6_NST) & 1);
        tc_fence_after();
        issue_s(j + 1);
      }
      if (j >= 1 && j - 1 + A6_NST < nkv) {
        mbar_wait_a(bar_pv + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
        load_kv(j - 1 + A6_NST);
      }
      mbar_wait_a(bar_p, j & 1);
      tc_fence_after();
      const uint32_t vb = sV + (j % A6_NST) * A6_SV;
#pragma unroll
      for (int k = 0; k < 4; ++k)  // 16 keys = 8 TMEM columns of P per MMA
        umma_h16_ts_p(on, tmem_base + A6_TM_O, tmem_base + A6_TM_P + k * 8, make_mnmajor_desc_sw64(vb + k * 1024, 0),
                       idesc_o, (j != 0 || k != 0) ? 1u : 0u);
      umma_commit_p(on, bar_pv + 8 * (j & 1));
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    float m_ref = -INFINITY, l = 0.f;
    const uint32_t s_tmem = tmem_base + lane_base;
    const uint32_t o_tmem = tmem_base + lane_base + A6_TM_O;
    const uint32_t p_tmem = tmem_base + lane_base + A6_TM_P;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait_a(bar_s, j & 1);
      tc_fence_after();
      float s[64];
      {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(s_tmem, r0);
        tmem_ld_32x32b_x32(s_tmem + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_a(bar_sfree);
#pragma unroll
        for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(r0[i]); s[32 + i] = __uint_as_float(r1[i]); }
      }
      if (j == nkv - 1) {
        const int lim = Lk - j * A6_BKV;  // keys >= lim are padding
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= lim) s[i] = -INFINITY;
      }
      // row maximum: four independent chains of 3-input maxima (FMNMX3, ALU pipe)
      float mq[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float m = max3f(s[16 * k], s[16 * k + 1], s[16 * k + 2]);
#pragma unroll
        for (int i = 3; i < 15; i += 2) m = max3f(m, s[16 * k + i], s[16 * k + i + 1]);
        mq[k] = fmaxf(m, s[16 * k + 15]);
      }
      const float mx = fmaxf(max3f(mq[0], mq[1], mq[2]), mq[3]);
      const bool need = mx > m_ref + AT_TAU;  // always true for j == 0 (m_ref = -inf)
      const bool any_need = __any_sync(0xffffffffu, need);
      const float a_corr = (need && j > 0) ? ex2_approx(m_ref - mx) : 1.0f;
      if (need) { m_ref = mx; l *= a_corr; }
      // P = exp2(S - m_ref) two scores at a time on packed fp32 (FADD2 / FFMA2: one issue slot per pair).  PP of every
      // 8 pairs take the Cody-Waite + degree-3 polynomial on the FMA pipe instead of MUFU.EX2 (16 /clk/SM): with
      // head_dim 32 there are only 128 tensor FLOPs per exponential, so this kernel is bound by the exponentials.
      uint32_t pk[32];
      uint64_t ls2[2] = {0ull, 0ull};
      const uint64_t m2 = pack_f32x2(m_ref, m_ref);
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const uint64_t x2 = sub_f32x2(pack_f32x2(s[2 * q], s[2 * q + 1]), m2);
        uint64_t p2;
        if ((POLY_MASK >> (q & 7)) & 1u) {
          p2 = ex2_poly_f32x2(x2);
        } else {
          float x0, x1;
          unpack_f32x2(x2, x0, x1);
          p2 = pack_f32x2(ex2_approx(x0), ex2_approx(x1));
        }
        float p0, p1;
        unpack_f32x2(p2, p0, p1);
        pk[q] = pack_h16x2(p0, p1);
        ls2[q & 1] = add_f32x2(ls2[q & 1], p2);
      }
      {
        float a0, a1;
        unpack_f32x2(add_f32x2(ls2[0], ls2[1]), a0, a1);
        l += a0 + a1;
      }
      if (j >= 1) {  // PV_{j-1} complete: P may be overwritten, O holds tiles 0..j-1
        mbar_wait_a(bar_pv + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
        tc_fence_after();
        if (any_need) {  // warp-uniform; rare after the first tiles
          uint32_t r[32];
          tmem_ld_32x32b_x32(o_tmem, r);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < 32; ++d) r[d] = __float_as_uint(__uint_as_float(r[d]) * a_corr);
          tmem_st_32x32b_x16(o_tmem, *reinterpret_cast<uint32_t (*)[16]>(&r[0]));
          tmem_st_32x32b_x16(o_tmem + 16, *reinterpret_cast<uint32_t (*)[16]>(&r[16]));
        }
      }
      tmem_st_32x32b_x16(p_tmem, *reinterpret_cast<uint32_t (*)[16]>(&pk[0]));
      tmem_st_32x32b_x16(p_tmem + 16, *reinterpret_cast<uint32_t (*)[16]>(&pk[16]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(bar_p);
    }
    {
      const int so = (nkv - 1) & 1;
      mbar_wait_a(bar_pv + 8 * so, ((nkv - 1) >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(o_tmem, r);
      tmem_ld_wait();
      const int q = q0 + row;
      if (q < L) {
        const int64_t m = static_cast<int64_t>(seq) * L + q;
        const float gsc = gates[m * heads + h] / l;
        uint4 u[4];
        uint32_t* w = reinterpret_cast<uint32_t*>(u);
#pragma unroll
        for (int d = 0; d < 16; ++d)
          w[d] = pack_h16x2(__uint_as_float(r[2 * d]) * gsc, __uint_as_float(r[2 * d + 1]) * gsc);
        uint4* dst = reinterpret_cast<uint4*>(out + m * C + h * 32);
#pragma unroll
        for (int d = 0; d < 4; ++d) dst[d] = u[d];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc<128>(tmem_base);
}

struct TcAttnPlan {
  CUtensorMap tmQK;
  CUtensorMap tmKV64;  // same tensor, 64-row boxes (attn_tc64_kernel)
  int seqs, L, heads;
};

TcAttnPlan* tc_attn_plan_create(const void* qkv, int seqs, int L, int heads, char* err, int errlen) {
  TcAttnPlan* p = new TcAttnPlan();
  p->seqs = seqs; p->L = L; p->heads = heads;
  const int C = heads * 32;
  const uint64_t dims[3] = {static_cast<uint64_t>(3 * C), static_cast<uint64_t>(L), static_cast<uint64_t>(seqs)};
  const uint64_t strides[2] = {static_cast<uint64_t>(3 * C) * 2, static_cast<uint64_t>(L) * 3 * C * 2};
  const uint32_t box[3] = {32, AT_BQ, 1};
  const uint32_t box64[3] = {32, A6_BKV, 1};
  if (!make_tmap(&p->tmQK, qkv, 3, dims, strides, box, 64, err, errlen) ||
      !make_tmap(&p->tmKV64, qkv, 3, dims, strides, box64, 64, err, errlen)) {
    delete p;
    return nullptr;
  }
  return p;
}
void tc_attn_plan_destroy(TcAttnPlan* p) { delete p; }

int launch_attn_time_tc(const TcAttnPlan* p, const float* gates, void* out, cudaStream_t st, const ChunkSrc* chunks,
                        int seqs_per_chunk) {
  dim3 grid(ceil_div(p->L, AT_BQ), p->heads, p->seqs);
  // BT_ATTN_POLY = number of score pairs out of 8 whose exp2 runs on the FMA pipe (default: measured best)
  static const int pp = getenv("BT_ATTN_POLY") ? atoi(getenv("BT_ATTN_POLY")) : A6_DEFAULT_PP;
  h16* o = reinterpret_cast<h16*>(out);
#define BT_A6_L(P_) attn_tc64_kernel<P_><<<grid, A6_THREADS, A6_SMEM, st>>>(p->tmQK, p->tmKV64, gates, o, p->L, p->heads, chunks, seqs_per_chunk)
  switch (pp) {
    case 0: BT_A6_L(0); break;
    case 1: BT_A6_L(1); break;
    case 2: BT_A6_L(2); break;
    case 3: BT_A6_L(3); break;
    case 4: BT_A6_L(4); break;
    default: return -3;
  }
#undef BT_A6_L
  return 0;
}

int tc_init_attn(char* err, int errlen) {
  cudaError_t r = cudaFuncSetAttribute(attn_tc64_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6_SMEM);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(attn_tc64_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6_SMEM);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(attn_tc64_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6_SMEM);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(attn_tc64_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6_SMEM);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(attn_tc64_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6_SMEM);
  if (r != cudaSuccess) {
    snprintf(err, errlen, "cudaFuncSetAttribute(attn_tc64_kernel) failed: %s", cudaGetErrorString(r));
    return -1;
  }
  return 0;
}

}  // namespace bt
