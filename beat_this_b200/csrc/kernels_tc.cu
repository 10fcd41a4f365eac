// bf16 tensor-core path (BT_DTYPE_BF16) for sm_100a: tcgen05.mma with TMEM accumulators,
// TMA (cp.async.bulk.tensor) operand staging through an mbarrier ring, warp-specialised
// roles.  Kernels:
//   gemm_tc_kernel   -- D = A * W^T over "planes" with shifted slabs (linear layers, the
//                       k(2,3) frontend convolutions as implicit GEMM, frontend.linear),
//                       persistent over output tiles, fused epilogues (epilogue.cuh);
//                       gemm_tc2_kernel = opt-in cta_group::2 variant.
//   attn_tc64_kernel -- flash attention for head_dim 32 over sequences of up to 1500 frames
//                       (time-direction attention of the frontend and the 6 main layers),
//                       P and O in tensor memory, 4 CTAs/SM; attn_tc_kernel<KS,POLY> = the
//                       earlier 2-CTAs/SM form (BT_ATTN_VARIANT=128).
//   fused_ff_kernel  -- persistent [out-projection +] RMSNorm + FFN + residual for C = 32/64.
//   fused_qkv_kernel -- persistent RMSNorm + gates + QKV + RoPE for C = 32/64.
// Every TMA / MMA issuing warp runs converged with predicated single-lane instructions (umma_bf16_p ...).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstring>

#include "epilogue.cuh"

namespace bt {

// Predicated forms for a CONVERGED issuer warp: every lane executes the asm block with warp-uniform
// operands, only the lane with `on != 0` (picked once with elect.sync) performs the operation.
// Keeping the warp converged lets ptxas keep descriptors in uniform registers instead of wrapping
// every tcgen05.mma of a divergent `if (lane == 0)` region in a vote loop (measured: ~85 cycles
// per MMA issue in the divergent form).
__device__ __forceinline__ void umma_bf16_p(uint32_t on, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(on)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_p(uint32_t on, uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(on)
      : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint32_t on, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(on)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_p(uint32_t on, uint32_t smem_dst, const void* tmap, uint32_t bar, int32_t c0,
                                              int32_t c1, int32_t c2) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %6, 0;\n\t"
      "@q cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n\t}" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(on)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_p(uint32_t on, uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes), "r"(on)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_p(uint32_t on, uint32_t smem_dst, const void* tmap, uint32_t bar, int32_t c0,
                                              int32_t c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n\t}" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(on)
      : "memory");
}


// --------------------------------------------------------------------------- tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static bool make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, int swizzle_bytes,
                      char* err, int errlen) {
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx,
                        es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu box %u,%u,%u stride0 %llu",
             static_cast<int>(r), rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
             (unsigned long long)(rank > 2 ? gd[2] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
             (unsigned long long)gs[0]);
    return false;
  }
  return true;
}

// =============================================================================== GEMM
constexpr int TG_BM = 128;
constexpr int TG_EPI_WARPS = 8;
constexpr int TG_THREADS = 64 + 32 * TG_EPI_WARPS;  // warp0 TMA, warp1 MMA, 8 epilogue warps

template <int BN, int BK>
struct TgCfg {
  static constexpr int A_BYTES = TG_BM * BK * 2;
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
  static constexpr int TCOLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  static constexpr int BIAS_BYTES = 16384;  // bias vector (N <= 4096 floats) staged for the epilogue
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + BIAS_BYTES;
  static constexpr int SWZ = BK * 2;  // 128 or 64 byte rows
};

// MC (multicast): launched as clusters of 2 CTAs that work on two M tiles of the SAME N tile in
// lock step; each CTA fetches one half of the W tile and TMA-multicasts it into both CTAs' shared
// memory, which cuts the L2 -> SM operand traffic of a 128x256 tile from 48 to 32 KB per k-block
// (the compute-bound GEMMs run into the L2 bandwidth otherwise).  Stage release is cluster-wide:
// both CTAs' MMAs must have consumed a stage before either producer may overwrite it.
template <int BN, int BK, bool MC>
__global__ void __launch_bounds__(TG_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const GemmShape g, const EpiParams e, int num_tiles, int t_tiles, int n_tiles, int m_tiles) {
  using Cfg = TgCfg<BN, BK>;
  const uint32_t crank = MC ? cluster_ctarank() : 0u;
  // tile walk: non-MC: tile -> (mt, nt).  MC: pair-tile -> (pair, nt), this CTA takes mt = 2*pair + rank.
  const int walk_start = MC ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int walk_step = MC ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  auto tile_mt = [&](int tile) { return MC ? 2 * (tile / n_tiles) + static_cast<int>(crank) : tile / n_tiles; };
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sBias = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool stage_bias = e.kind == 0 && e.bias != nullptr && g.N <= Cfg::BIAS_BYTES / 4;
  if (stage_bias)
    for (int i = threadIdx.x; i < g.N; i += TG_THREADS) sBias[i] = __ldg(e.bias + i);
  const uint32_t bias_smem = stage_bias ? smem_u32(sBias) : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], MC ? 2 : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], BN <= 64 ? TG_EPI_WARPS / 2 : TG_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TCOLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // peer barriers initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int kb_per_slab = g.Kslab / BK;
  const int num_kb = g.nslab * kb_per_slab;

  if (warp == 0) {
    // producer and MMA warps run CONVERGED with predicated single-lane TMA / MMA / commit instructions
    // (see umma_bf16_p): in a divergent `if (lane == 0)` block every tcgen05.mma costs ~85 issue cycles.
    const uint32_t on = elect_one() ? 1u : 0u;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = walk_start; tile < num_tiles; tile += walk_step) {
      const int mt = tile_mt(tile), nt = tile % n_tiles;
      const int p_out = mt / t_tiles;  // beyond the last plane for the dummy half of an odd pair: TMA zero-fills
      const int t0 = (mt - p_out * t_tiles) * TG_BM;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb / kb_per_slab;
        const int k0 = (kb - s * kb_per_slab) * BK;
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t fb = smem_u32(&full[stage]);
        mbar_expect_tx_p(on, fb, Cfg::STAGE_BYTES);
        tma_load_3d_p(on, smem_u32(sA + stage * Cfg::A_BYTES), &tmA, fb, k0, t0 + g.t_shift[s],
                      p_out * g.plane_mul + g.plane_add[s]);
        if constexpr (MC) {  // my half of the W tile, delivered to both CTAs of the pair
          if (on)
            tma_load_2d_mc(sW + stage * Cfg::W_BYTES + crank * (Cfg::W_BYTES / 2), &tmW, &full[stage], s * g.Kslab + k0,
                           nt * BN + static_cast<int>(crank) * (BN / 2), 0x3);
          __syncwarp();
        } else {
          tma_load_2d_p(on, smem_u32(sW + stage * Cfg::W_BYTES), &tmW, fb, s * g.Kslab + k0, nt * BN);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t on = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc = make_idesc_bf16(TG_BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = walk_start; tile < num_tiles; tile += walk_step) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + stage * Cfg::A_BYTES);
        const uint32_t b_base = smem_u32(sW + stage * Cfg::W_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_bf16_p(on, d_tmem, make_kmajor_desc<Cfg::SWZ>(a_base + k * 32), make_kmajor_desc<Cfg::SWZ>(b_base + k * 32),
                      idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if constexpr (MC) {
          if (on) umma_commit_mc(&empty[stage], 0x3);
          __syncwarp();
        } else {
          umma_commit_p(on, smem_u32(&empty[stage]));
        }
        if (kb == num_kb - 1) umma_commit_p(on, smem_u32(&tfull[acc]));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // Epilogue warps.  A warp may only touch TMEM lanes [32*(warp%4), +32); thread = one output row.
    //  BN >= 96: the 8 warps split the columns of every tile (2 warps per lane quarter);
    //  BN <= 64: warps 2-5 take the even tiles of this CTA and warps 6-9 the odd ones.
    // The residual rows of the NEXT 32-column chunk are requested before the current chunk is
    // processed (software pipelining: the epilogue is latency-bound on those loads otherwise,
    // ncu long_scoreboard 64% -- profiles/r1_notes.md).
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int NCH = BN / 32;
    constexpr bool TILE_SPLIT = NCH <= 2;
    constexpr int SPLIT = (NCH + 1) / 2;
    constexpr int MAXC = TILE_SPLIT ? NCH : SPLIT;
    const int c_begin = TILE_SPLIT ? 0 : (half == 0 ? 0 : SPLIT);
    const int nch = TILE_SPLIT ? NCH : (half == 0 ? SPLIT : NCH - SPLIT);
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int iter = 0;
    for (int tile = walk_start; tile < num_tiles; tile += walk_step, ++iter) {
      if (TILE_SPLIT && (iter & 1) != half) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      const int mt = tile_mt(tile), nt = tile % n_tiles;
      const int p_out = mt / t_tiles;
      const int t = (mt - p_out * t_tiles) * TG_BM + row;
      const bool valid = t < g.L && mt < m_tiles;
      const int64_t m = static_cast<int64_t>(p_out) * g.L + t;
      float ra[32], rb[32];
      const bool has_resid = e.kind == 0 && e.resid != nullptr && valid;
      const bool has_rope = e.kind == 1 && valid;
      auto load_resid = [&](int c, float (&dst)[32]) {
        const float4* r4 = reinterpret_cast<const float4*>(e.resid + m * e.ldr + nt * BN + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 q = r4[i];
          dst[4 * i] = q.x; dst[4 * i + 1] = q.y; dst[4 * i + 2] = q.z; dst[4 * i + 3] = q.w;
        }
      };
      if (has_rope) {  // cos[16] | sin[16] of this row's position, reused by every q/k head of the row
        const int pos = e.posmode == 0 ? t : static_cast<int>((m / g.L) % e.F);
        const float4* c4 = reinterpret_cast<const float4*>(e.rope_cos + pos * 16);
        const float4* s4 = reinterpret_cast<const float4*>(e.rope_sin + pos * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = __ldg(c4 + i), b = __ldg(s4 + i);
          ra[4 * i] = a.x; ra[4 * i + 1] = a.y; ra[4 * i + 2] = a.z; ra[4 * i + 3] = a.w;
          ra[16 + 4 * i] = b.x; ra[16 + 4 * i + 1] = b.y; ra[16 + 4 * i + 2] = b.z; ra[16 + 4 * i + 3] = b.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) rb[i] = ra[i];  // either buffer may be handed to the epilogue
      }
      if (has_resid) load_resid(c_begin, ra);  // in flight while we wait for the accumulator
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        if (k < nch) {
          const int c = c_begin + k;
          float (&cur)[32] = (k & 1) ? rb : ra;
          float (&nxt)[32] = (k & 1) ? ra : rb;
          if (has_resid && k + 1 < nch) load_resid(c + 1, nxt);
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c * 32, r);
          tmem_ld_wait();
          if (valid) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            epilogue_apply<bf16, 32>(e, g.L, m, nt * BN + c * 32, v, cur, has_rope || has_resid, bias_smem);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // the peer may still multicast into / arrive on this CTA's smem
  if (warp == 1) tmem_dealloc<Cfg::TCOLS>(tmem_base);
}

// ----------------------------------------------------------------------------- 2-SM GEMM
// cta_group::2: a cluster of two CTAs (an SM pair) computes a 256 x 256 output tile with ONE stream
// of tcgen05.mma instructions issued by the leader CTA.  Each CTA stages its own 128 rows of A and
// one HALF (128 rows) of the W tile and keeps its own 128 x 256 accumulator in TMEM, so every SM
// reads/receives 32 KB of operands per k-block instead of 48 KB -- the 1-SM kernel above is bound by
// the 128 B/clk shared-memory port (96 B/clk of MMA operand reads + 96 B/clk of TMA fills).
//   full[s]  : leader's barrier, 2 arrivals (one producer per CTA) + the bytes of both CTAs' TMA loads
//   empty[s] : per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tfull[a] : per CTA, same multicast commit after the last k-block
//   tempty[a]: leader's barrier, 16 arrivals (8 epilogue warps of each CTA)
constexpr int T2_BN = 256, T2_BK = 64, T2_STAGES = 6;
constexpr int T2_A_BYTES = 128 * T2_BK * 2, T2_W_BYTES = 128 * T2_BK * 2, T2_STAGE_BYTES = T2_A_BYTES + T2_W_BYTES;
constexpr int T2_SMEM = T2_STAGES * T2_STAGE_BYTES + 1024 + 256 + 16384;

__global__ void __launch_bounds__(TG_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmShape g,
                const EpiParams e, int num_tiles, int t_tiles, int n_tiles, int m_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = sbase;
  const uint32_t sW = sA + T2_STAGES * T2_A_BYTES;
  const uint32_t full = sW + T2_STAGES * T2_W_BYTES;  // 8 bytes each
  const uint32_t empty = full + 8 * T2_STAGES;
  const uint32_t tfull = empty + 8 * T2_STAGES;
  const uint32_t tempty = tfull + 16;
  const uint32_t tmem_slot = tempty + 16;
  const uint32_t sBias = full + 256;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const bool stage_bias = e.kind == 0 && e.bias != nullptr && g.N <= 4096;
  if (stage_bias)
    for (int i = threadIdx.x; i < g.N; i += TG_THREADS) st_shared_f32(sBias + 4 * i, __ldg(e.bias + i));
  const uint32_t bias_smem = stage_bias ? sBias : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    auto init = [](uint32_t bar, uint32_t count) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    };
    for (int i = 0; i < T2_STAGES; ++i) { init(full + 8 * i, 2); init(empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { init(tfull + 8 * i, 1); init(tempty + 8 * i, 2 * TG_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int kb_per_slab = g.Kslab / T2_BK;
  const int num_kb = g.nslab * kb_per_slab;
  const int walk_start = static_cast<int>(blockIdx.x >> 1), walk_step = static_cast<int>(gridDim.x >> 1);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = walk_start; tile < num_tiles; tile += walk_step) {
        const int mt = 2 * (tile / n_tiles) + static_cast<int>(crank), nt = tile % n_tiles;
        const int p_out = mt / t_tiles;
        const int t0 = (mt - p_out * t_tiles) * TG_BM;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int s = kb / kb_per_slab;
          const int k0 = (kb - s * kb_per_slab) * T2_BK;
          mbar_wait_a(empty + 8 * stage, phase ^ 1);
          const uint32_t lead_full = mapa_cluster(full + 8 * stage, 0);
          if (leader) mbar_expect_tx_a(full + 8 * stage, 2 * T2_STAGE_BYTES);
          else mbar_arrive_cluster(lead_full);
          tma_load_3d_2sm(sA + stage * T2_A_BYTES, &tmA, lead_full, k0, t0 + g.t_shift[s], p_out * g.plane_mul + g.plane_add[s]);
          tma_load_2d_2sm(sW + stage * T2_W_BYTES, &tmW, lead_full, s * g.Kslab + k0, nt * T2_BN + static_cast<int>(crank) * 128);
          if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, T2_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = walk_start; tile < num_tiles; tile += walk_step) {
        mbar_wait_a(tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * T2_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_a(full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t da = make_kmajor_desc<128>(sA + stage * T2_A_BYTES), db = make_kmajor_desc<128>(sW + stage * T2_W_BYTES);
#pragma unroll
          for (int k = 0; k < T2_BK / 16; ++k) umma_bf16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(empty + 8 * stage, 0x3);
          if (kb == num_kb - 1) umma_commit_2sm(tfull + 8 * acc, 0x3);
          if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int NCH = T2_BN / 32;
    const int c_begin = half * (NCH / 2);
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = walk_start; tile < num_tiles; tile += walk_step) {
      const int mt = 2 * (tile / n_tiles) + static_cast<int>(crank), nt = tile % n_tiles;
      const int p_out = mt / t_tiles;
      const int t = (mt - p_out * t_tiles) * TG_BM + row;
      const bool valid = t < g.L && mt < m_tiles;
      const int64_t m = static_cast<int64_t>(p_out) * g.L + t;
      float ra[32], rb[32];
      const bool has_resid = e.kind == 0 && e.resid != nullptr && valid;
      const bool has_rope = e.kind == 1 && valid;
      auto load_resid = [&](int c, float (&dst)[32]) {
        const float4* r4 = reinterpret_cast<const float4*>(e.resid + m * e.ldr + nt * T2_BN + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 q = r4[i];
          dst[4 * i] = q.x; dst[4 * i + 1] = q.y; dst[4 * i + 2] = q.z; dst[4 * i + 3] = q.w;
        }
      };
      if (has_rope) {
        const int pos = e.posmode == 0 ? t : static_cast<int>((m / g.L) % e.F);
        const float4* c4 = reinterpret_cast<const float4*>(e.rope_cos + pos * 16);
        const float4* s4 = reinterpret_cast<const float4*>(e.rope_sin + pos * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = __ldg(c4 + i), b = __ldg(s4 + i);
          ra[4 * i] = a.x; ra[4 * i + 1] = a.y; ra[4 * i + 2] = a.z; ra[4 * i + 3] = a.w;
          ra[16 + 4 * i] = b.x; ra[16 + 4 * i + 1] = b.y; ra[16 + 4 * i + 2] = b.z; ra[16 + 4 * i + 3] = b.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) rb[i] = ra[i];
      }
      if (has_resid) load_resid(c_begin, ra);
      mbar_wait_a(tfull + 8 * acc, acc_phase);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < NCH / 2; ++k) {
        const int c = c_begin + k;
        float (&cur)[32] = (k & 1) ? rb : ra;
        float (&nxt)[32] = (k & 1) ? ra : rb;
        if (has_resid && k + 1 < NCH / 2) load_resid(c + 1, nxt);
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * T2_BN + c * 32, r);
        tmem_ld_wait();
        if (valid) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          epilogue_apply<bf16, 32>(e, g.L, m, nt * T2_BN + c * 32, v, cur, has_rope || has_resid, bias_smem);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive_a(tempty + 8 * acc);
        else mbar_arrive_cluster(mapa_cluster(tempty + 8 * acc, 0));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm<512>(tmem_base);
}

struct TcGemmPlan {
  CUtensorMap tmA, tmW;
  GemmShape g;
  int BN, BK;
  int num_tiles, t_tiles, n_tiles, m_tiles, grid;
  bool mc;      // 2-CTA cluster, W tile multicast, cta_group::1 MMAs
  bool two_sm;  // 2-CTA cluster, cta_group::2 MMAs (256 x 256 tile per pair)
};

static int g_num_sms = 148;

static int pick_bn(int N) {
  const int cands[6] = {256, 192, 128, 96, 64, 32};
  for (int i = 0; i < 6; ++i)
    if (N % cands[i] == 0) return cands[i];
  return 0;
}

template <int BN, int BK, bool MC>
static int gemm_tc_launch(const TcGemmPlan* p, const EpiParams& e, cudaStream_t st) {
  using Cfg = TgCfg<BN, BK>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t r = cudaFuncSetAttribute(gemm_tc_kernel<BN, BK, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (r != cudaSuccess) return -1;
    attr_set = true;
  }
  if constexpr (MC) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p->grid);
    cfg.blockDim = dim3(TG_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t r = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, BK, MC>, p->tmA, p->tmW, p->g, e, p->num_tiles, p->t_tiles,
                                       p->n_tiles, p->m_tiles);
    return r == cudaSuccess ? 0 : -1;
  } else {
    gemm_tc_kernel<BN, BK, MC><<<p->grid, TG_THREADS, Cfg::SMEM, st>>>(p->tmA, p->tmW, p->g, e, p->num_tiles, p->t_tiles,
                                                                         p->n_tiles, p->m_tiles);
    return 0;
  }
}

TcGemmPlan* tc_gemm_plan_create(const void* A, const void* W, const GemmShape& g, int planes_in, char* err,
                                int errlen) {
  TcGemmPlan* p = new TcGemmPlan();
  p->g = g;
  p->BK = (g.Kslab % 64 == 0) ? 64 : 32;
  p->BN = pick_bn(g.N);
  if (p->BN == 0 || g.Kslab % 32 != 0 || (p->BK == 32 && p->BN > 128)) {
    snprintf(err, errlen, "tc gemm: unsupported shape N=%d Kslab=%d", g.N, g.Kslab);
    delete p;
    return nullptr;
  }
  static int mc_enabled = -1;  // measured: no gain (the 128x256 tiles are bound by SM-side smem bandwidth, not L2)
  if (mc_enabled < 0) { const char* e = getenv("BT_GEMM_MULTICAST"); mc_enabled = (e && e[0] == '1'); }
  static int two_sm_enabled = -1;
  // measured 40 % SLOWER than the 1-SM kernel on the K = 512 / 2048 layer GEMMs (profiles/r1_notes.md): opt-in
  if (two_sm_enabled < 0) { const char* e2 = getenv("BT_GEMM_2SM"); two_sm_enabled = (e2 && e2[0] == '1'); }
  p->two_sm = two_sm_enabled && p->BN == 256 && p->BK == 64;
  p->mc = !p->two_sm && mc_enabled && p->BN == 256 && p->BK == 64;
  const int swz = p->BK * 2;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(g.Kslab), static_cast<uint64_t>(g.L),
                              static_cast<uint64_t>(planes_in)};
    const uint64_t strides[2] = {static_cast<uint64_t>(g.lda) * 2, static_cast<uint64_t>(g.L) * g.lda * 2};
    const uint32_t box[3] = {static_cast<uint32_t>(p->BK), TG_BM, 1};
    if (!make_tmap(&p->tmA, A, 3, dims, strides, box, swz, err, errlen)) { delete p; return nullptr; }
  }
  {
    const uint64_t Ktot = static_cast<uint64_t>(g.Kslab) * g.nslab;
    const uint64_t dims[2] = {Ktot, static_cast<uint64_t>(g.N)};
    const uint64_t strides[1] = {Ktot * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(p->BK), static_cast<uint32_t>((p->mc || p->two_sm) ? p->BN / 2 : p->BN)};
    if (!make_tmap(&p->tmW, W, 2, dims, strides, box, swz, err, errlen)) { delete p; return nullptr; }
  }
  p->t_tiles = ceil_div(g.L, TG_BM);
  p->n_tiles = g.N / p->BN;
  p->m_tiles = p->t_tiles * g.planes_out;
  if (p->mc || p->two_sm) {
    p->num_tiles = ceil_div(p->m_tiles, 2) * p->n_tiles;  // pair tiles
    const int want = 2 * p->num_tiles;
    p->grid = (want < g_num_sms ? want : g_num_sms) & ~1;
  } else {
    p->num_tiles = p->m_tiles * p->n_tiles;
    p->grid = p->num_tiles < g_num_sms ? p->num_tiles : g_num_sms;
  }
  return p;
}
void tc_gemm_plan_destroy(TcGemmPlan* p) { delete p; }

int launch_gemm_tc(const TcGemmPlan* p, const EpiParams& e, cudaStream_t st) {
#define BT_TG_CASE(bn, bk) \
  if (p->BN == bn && p->BK == bk) return gemm_tc_launch<bn, bk, false>(p, e, st);
  if (p->two_sm) {
    static bool attr_set = false;
    if (!attr_set) {
      if (cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM) != cudaSuccess) return -1;
      attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p->grid);
    cfg.blockDim = dim3(TG_THREADS);
    cfg.dynamicSmemBytes = T2_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gemm_tc2_kernel, p->tmA, p->tmW, p->g, e, p->num_tiles, p->t_tiles, p->n_tiles,
                              p->m_tiles) == cudaSuccess ? 0 : -1;
  }
  if (p->mc) return gemm_tc_launch<256, 64, true>(p, e, st);
  BT_TG_CASE(256, 64) BT_TG_CASE(192, 64) BT_TG_CASE(128, 64) BT_TG_CASE(96, 64) BT_TG_CASE(64, 64)
  BT_TG_CASE(32, 64) BT_TG_CASE(128, 32) BT_TG_CASE(96, 32) BT_TG_CASE(64, 32) BT_TG_CASE(32, 32)
#undef BT_TG_CASE
  return -2;
}

// ========================================================================== attention
// One CTA per (sequence, head, 128-query tile), 2 CTAs/SM.  Warps 0-7: softmax, two threads per
// query row (warps w and w+4 share TMEM lane quarter w%4; keys [0,64) / [64,128) of each tile);
// warp 8 (converged, one elected lane issues): TMA producer + MMA issuer.  TMEM (256 columns): S [0,128) | O [128,192) | P [192,256).
//   S_j = Q K_j^T  : A = Q  [128 x 32] bf16 (K-major, SW64), B = K tile [128 keys x 32] (K-major, SW64)
//   O  += P_j [V_j | 1] : A = P_j [128 x 128] bf16 **in tensor memory** (written by the softmax
//                    threads with tcgen05.st, two keys per 32-bit column), B = V tile exactly as the
//                    QKV GEMM stored it (MN-major operand, SW64) + a constant ones block ->
//                    TMEM cols [128,160) = O, col 160 = row sum of P
// Shared memory only carries Q/K/V: the probabilities never touch it.  (Measured by ablation on the
// previous design, profiles/r1_notes.md: the 32 KB/tile of st.shared for P plus the MMA reading it
// back cost 20-30 % of the kernel -- the shared-memory pipe, not MUFU or the tensor pipe, was the
// contended resource.)
// The output accumulates in TMEM over all key tiles (tensor-core accumulate) with FlashAttention-4
// style lazy rescaling: P_j = exp2(S_j - m_ref) with a per-row reference that is only raised -- and O
// rescaled through tcgen05.ld/st -- when the tile maximum exceeds it by more than AT_TAU (P <= 2^AT_TAU
// stays far inside bf16/fp32 range; numerator and row sum share the same scaling).
// q is pre-scaled by log2(e)/sqrt(32) in the QKV epilogue -> exp2 softmax.
// (ex2.approx.f16x2 -- two exponentials per MUFU op -- was measured: 28% SLOWER than fp32
// ex2 + bf16 pack on B200, profiles/r1_notes.md.)
constexpr int AT_BQ = 128, AT_BKV = 128;
constexpr int AT_SQ = 8192, AT_SK = 8192, AT_SV = 8192, AT_SONES = 8192, AT_SMAX = 4096;
constexpr int AT_NST = 4;  // K/V stages: tile j+3 is in flight while tile j is consumed (TMA latency >> one tile otherwise)
constexpr int AT_SMEM = AT_SQ + AT_NST * (AT_SK + AT_SV) + AT_SONES + AT_SMAX + 1024 + 128;
constexpr int AT_POLY_MOD = 0;  // every AT_POLY_MOD-th exponential runs on the FMA pipe instead of MUFU (0: none)
constexpr float AT_TAU = 8.0f;  // log2 units: rescale O only when the row maximum grew by more than this
constexpr uint32_t AT_TM_O = 128, AT_TM_L = 160, AT_TM_P = 192;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x <= AT_TAU on the FMA/ALU pipes (Cody-Waite split + degree-3 polynomial, rel. error
// 8e-5, far below the bf16 rounding of P).  The MUFU pipe (16 ex2/clk/SM) is one bottleneck
// of head_dim-32 attention -- 128 tensor FLOPs per exponential -- so a fixed fraction of the
// exponentials is moved to the FMA pipe (the FlashAttention-4 trick).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -120.0f);
  const float t = x + 12582912.0f;        // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float r = x - (t - 12582912.0f);  // [-0.5, 0.5]
  float p = fmaf(0.05508868f, r, 0.24260405f);
  p = fmaf(p, r, 0.69327623f);
  p = fmaf(p, r, 0.99992895f);
  int y;  // p * 2^round(x): add round(x) (low mantissa bits of t) to the exponent field, one IMAD
  asm("mad.lo.s32 %0, %1, 8388608, %2;" : "=r"(y) : "r"(__float_as_int(t)), "r"(__float_as_int(p)));
  return __int_as_float(y);
}
__device__ __forceinline__ uint32_t tmem_ld_32x32b_x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x1(uint32_t taddr, uint32_t r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
template <int N> __device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_ld_n<8>(uint32_t taddr, uint32_t (&r)[8]) { tmem_ld_32x32b_x8(taddr, r); }
template <> __device__ __forceinline__ void tmem_ld_n<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x32b_x16(taddr, r); }
template <int N> __device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_st_n<8>(uint32_t taddr, const uint32_t (&r)[8]) { tmem_st_32x32b_x8(taddr, r); }
template <> __device__ __forceinline__ void tmem_st_n<16>(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st_32x32b_x16(taddr, r); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: A is read from tensor memory (lane = row, one 32-bit column
// per two K elements), kind::f16 with bf16 operands.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// B operand in MN-major form (N contiguous): rows of 64 B (32 bf16), SWIZZLE_64B, 8-row groups
// 512 B apart (SBO); a second 32-column block of N lives `lbo_bytes` after the first (LBO).
__device__ __forceinline__ uint64_t make_mnmajor_desc_sw64(uint32_t smem_addr, uint32_t lbo_bytes) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

// KS = softmax threads per query row (2: 8 softmax warps, 64 keys each; 4: 16 warps, 32 keys each);
// POLY: every POLY-th exponential on the FMA pipe (0 = all on MUFU).
template <int KS, int POLY>
__global__ void __launch_bounds__(32 * (4 * KS + 1), 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const float* __restrict__ gates, bf16* __restrict__ out,
               int L, int heads) {
  extern __shared__ uint8_t smem_raw[];
  // everything below works on 32-bit shared-space addresses computed once
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = sbase;
  const uint32_t sK = sQ + AT_SQ;              // [AT_NST]
  const uint32_t sV = sK + AT_NST * AT_SK;     // [AT_NST]
  const uint32_t sOnes = sV + AT_NST * AT_SV;  // [128 keys][32 cols] bf16, col 0 = 1: second N block of the PV MMA -> row sums of P
  const uint32_t sMax = sOnes + AT_SONES; // [2 parity][KS parts][128 rows] fp32 partial row maxima
  const uint32_t bar_q = sMax + AT_SMAX;
  const uint32_t bar_kv = bar_q + 8;      // [AT_NST]
  const uint32_t bar_s = bar_kv + 8 * AT_NST;
  const uint32_t bar_sfree = bar_s + 8;   // S_j has been copied to registers: S_{j+1} may overwrite it
  const uint32_t bar_p = bar_sfree + 8;   // P_j sits in TMEM (and O was rescaled if the maximum jumped)
  const uint32_t bar_pv = bar_p + 8;      // [2] PV_j complete: P and K/V stage j&1 free, O holds tiles 0..j
  const uint32_t tmem_slot = bar_pv + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ;
  const int h = blockIdx.y;
  const int seq = blockIdx.z;
  const int C = heads * 32;
  const int nkv = ceil_div(L, AT_BKV);
  constexpr int MMA_WARP = 4 * KS;
  constexpr int NSOFT = 4 * KS * 32;
  constexpr int NK = 128 / KS;  // keys per softmax thread and tile
  constexpr int NO = 32 / KS;   // output columns per softmax thread

  if (warp == MMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQK);
    auto init = [](uint32_t bar, uint32_t count) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    };
    init(bar_q, 1);
    for (int i = 0; i < AT_NST; ++i) init(bar_kv + 8 * i, 1);
    init(bar_s, 1);
    init(bar_sfree, NSOFT);
    init(bar_p, NSOFT);
    init(bar_pv, 1); init(bar_pv + 8, 1);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == MMA_WARP) {
    // the whole warp runs this loop converged; `on` marks the one lane that issues TMA / MMA / commit
    const uint32_t on = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64) | (1u << 16);  // bit 16: B is MN-major
    auto load_kv = [&](int j) {
      const int st = j % AT_NST;
      mbar_expect_tx_p(on, bar_kv + 8 * st, AT_SK + AT_SV);
      tma_load_3d_p(on, sK + st * AT_SK, &tmQK, bar_kv + 8 * st, C + h * 32, j * AT_BKV, seq);
      tma_load_3d_p(on, sV + st * AT_SV, &tmQK, bar_kv + 8 * st, 2 * C + h * 32, j * AT_BKV, seq);
    };
    auto issue_s = [&](int j) {
      const uint32_t kb = sK + (j % AT_NST) * AT_SK;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        umma_bf16_p(on, tmem_base, make_kmajor_desc<64>(sQ + k * 32), make_kmajor_desc<64>(kb + k * 32), idesc_s,
                    k != 0 ? 1u : 0u);
      umma_commit_p(on, bar_s);
    };
    mbar_expect_tx_p(on, bar_q, AT_SQ);
    tma_load_3d_p(on, sQ, &tmQK, bar_q, h * 32, q0, seq);
    for (int j = 0; j < AT_NST && j < nkv; ++j) load_kv(j);
    mbar_wait_a(bar_q, 0);
    mbar_wait_a(bar_kv, 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) {  // S_{j+1} is computed while the softmax warps work on S_j
        mbar_wait_a(bar_sfree, j & 1);  // S_j copied to registers by every softmax thread
        mbar_wait_a(bar_kv + 8 * ((j + 1) % AT_NST), ((j + 1) / AT_NST) & 1);
        tc_fence_after();
        issue_s(j + 1);
      }
      if (j >= 1 && j - 1 + AT_NST < nkv) {  // PV_{j-1} done (issued a whole tile ago) -> refill its K/V stage
        mbar_wait_a(bar_pv + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
        load_kv(j - 1 + AT_NST);
      }
      mbar_wait_a(bar_p, j & 1);  // P_j written to TMEM (and O rescaled if the maximum jumped)
      tc_fence_after();
      const uint32_t vb = sV + (j % AT_NST) * AT_SV;
      const uint32_t lbo = sOnes - vb;
#pragma unroll
      for (int k = 0; k < 8; ++k)  // 16 keys = 8 TMEM columns of P per MMA
        umma_bf16_ts_p(on, tmem_base + AT_TM_O, tmem_base + AT_TM_P + k * 8, make_mnmajor_desc_sw64(vb + k * 1024, lbo),
                       idesc_o, (j != 0 || k != 0) ? 1u : 0u);
      umma_commit_p(on, bar_pv + 8 * (j & 1));
    }
  } else {
    const int quarter = warp & 3;
    const int kq = warp >> 2;  // which part of the keys / of the output columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    if (kq == 0) {  // ones block: logical column 0 of key row `row` is 1.0 (bf16 0x3F80), the rest 0; SW64 swizzle
      const uint32_t orow = sOnes + row * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool one = i == ((row >> 1) & 3);
        st_shared_v4(orow + 16 * i, one ? 0x3F80u : 0u, 0u, 0u, 0u);
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the MMA (async proxy) after bar_p
    }
    float m_ref = -INFINITY;
    const uint32_t s_tmem = tmem_base + lane_base + kq * NK;
    const uint32_t o_tmem = tmem_base + lane_base + AT_TM_O + kq * NO;        // this thread's output columns
    const uint32_t l_tmem = tmem_base + lane_base + AT_TM_L;                  // row sum of P (ones block column)
    const uint32_t p_tmem = tmem_base + lane_base + AT_TM_P + kq * (NK / 2);  // this thread's keys, two per column
    const uint32_t max_row = sMax + row * 4;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait_a(bar_s, j & 1);
      tc_fence_after();
      float s[NK];
      {
        uint32_t r0[32];
        tmem_ld_32x32b_x32(s_tmem, r0);
        if constexpr (NK == 64) {
          uint32_t r1[32];
          tmem_ld_32x32b_x32(s_tmem + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s[32 + (i & (NK - 33))] = __uint_as_float(r1[i]);
        } else {
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive_a(bar_sfree);
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(r0[i]);
      }
      if (j == nkv - 1) {
        const int lim = L - j * AT_BKV - kq * NK;  // keys >= lim are padding
#pragma unroll
        for (int i = 0; i < NK; ++i)
          if (i >= lim) s[i] = -INFINITY;
      }
      float mxs[8];  // independent max chains
#pragma unroll
      for (int k = 0; k < 8; ++k) mxs[k] = fmaxf(s[k], s[8 + k]);
#pragma unroll
      for (int i = 16; i < NK; i += 16) {
#pragma unroll
        for (int k = 0; k < 8; ++k) mxs[k] = fmaxf(mxs[k], fmaxf(s[i + k], s[i + 8 + k]));
      }
      float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                       fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      {  // exchange the partial maxima between the KS threads that share this row
        const uint32_t par = max_row + (j & 1) * (KS * 512);
        st_shared_f32(par + kq * 512, mx);
        named_bar_sync(1 + quarter, 32 * KS);
#pragma unroll
        for (int kk = 1; kk < KS; ++kk) mx = fmaxf(mx, ld_shared_f32(par + ((kq + kk) & (KS - 1)) * 512));
      }
      // lazy rescale: identical decision in all threads of a row (same m_ref, same mx)
      const bool need = mx > m_ref + AT_TAU;  // always true for j == 0 (m_ref = -inf)
      const bool any_need = __any_sync(0xffffffffu, need);
      const float a_corr = (need && j > 0) ? ex2_approx(m_ref - mx) : 1.0f;
      if (need) m_ref = mx;
      uint32_t pk[NK / 2];  // P_j of this thread, two keys per register
#pragma unroll
      for (int c = 0; c < NK / 8; ++c) {
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = s[c * 8 + i] - m_ref;
          p[i] = (POLY > 0 && i % (POLY > 0 ? POLY : 1) == (POLY > 0 ? POLY : 1) - 1) ? ex2_poly(x) : ex2_approx(x);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) pk[c * 4 + i] = pack_bf16x2(p[2 * i], p[2 * i + 1]);
      }
      if (j >= 1) {  // PV_{j-1} complete: P may be overwritten, O holds tiles 0..j-1
        mbar_wait_a(bar_pv + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
        tc_fence_after();
        if (any_need) {  // warp-uniform; rare after the first tiles
          uint32_t r[NO];
          tmem_ld_n<NO>(o_tmem, r);
          uint32_t rs = 0;
          if (kq == 0) rs = tmem_ld_32x32b_x1(l_tmem);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < NO; ++d) r[d] = __float_as_uint(__uint_as_float(r[d]) * a_corr);
          tmem_st_n<NO>(o_tmem, r);
          if (kq == 0) tmem_st_32x32b_x1(l_tmem, __float_as_uint(__uint_as_float(rs) * a_corr));
        }
      }
#pragma unroll
      for (int i = 0; i < NK / 32; ++i)
        tmem_st_32x32b_x16(p_tmem + 16 * i, *reinterpret_cast<uint32_t (*)[16]>(&pk[16 * i]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(bar_p);
    }
    {
      const int so = (nkv - 1) & 1;
      mbar_wait_a(bar_pv + 8 * so, ((nkv - 1) >> 1) & 1);
      tc_fence_after();
      uint32_t r[NO];
      tmem_ld_n<NO>(o_tmem, r);
      const uint32_t rs = tmem_ld_32x32b_x1(l_tmem);
      tmem_ld_wait();
      const int q = q0 + row;
      if (q < L) {
        const int64_t m = static_cast<int64_t>(seq) * L + q;
        const float gsc = gates[m * heads + h] / __uint_as_float(rs);
        uint4 u[NO / 8];
        uint32_t* w = reinterpret_cast<uint32_t*>(u);
#pragma unroll
        for (int d = 0; d < NO / 2; ++d)
          w[d] = pack_bf16x2(__uint_as_float(r[2 * d]) * gsc, __uint_as_float(r[2 * d + 1]) * gsc);
        uint4* dst = reinterpret_cast<uint4*>(out + m * C + h * 32 + kq * NO);
#pragma unroll
        for (int d = 0; d < NO / 8; ++d) dst[d] = u[d];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------ 4-CTAs-per-SM variant
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

template <int POLY>  // every POLY-th exponential on the FMA pipe (0: all on MUFU)
__global__ void __launch_bounds__(A6_THREADS, 4)
attn_tc64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const float* __restrict__ gates, bf16* __restrict__ out, int L, int heads) {
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
  const int nkv = ceil_div(L, A6_BKV);
  constexpr int MMA_WARP = 4;
  constexpr int NSOFT = 128;

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
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 32) | (1u << 16);  // bit 16: B is MN-major
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
        umma_bf16_p(on, tmem_base, make_kmajor_desc<64>(sQ + k * 32), make_kmajor_desc<64>(kb + k * 32), idesc_s,
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
        umma_bf16_ts_p(on, tmem_base + A6_TM_O, tmem_base + A6_TM_P + k * 8, make_mnmajor_desc_sw64(vb + k * 1024, 0),
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
        const int lim = L - j * A6_BKV;  // keys >= lim are padding
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= lim) s[i] = -INFINITY;
      }
      float mxs[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) mxs[k] = fmaxf(s[k], s[8 + k]);
#pragma unroll
      for (int i = 16; i < 64; i += 16) {
#pragma unroll
        for (int k = 0; k < 8; ++k) mxs[k] = fmaxf(mxs[k], fmaxf(s[i + k], s[i + 8 + k]));
      }
      const float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                             fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      const bool need = mx > m_ref + AT_TAU;  // always true for j == 0 (m_ref = -inf)
      const bool any_need = __any_sync(0xffffffffu, need);
      const float a_corr = (need && j > 0) ? ex2_approx(m_ref - mx) : 1.0f;
      if (need) { m_ref = mx; l *= a_corr; }
      uint32_t pk[32];
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = s[c * 8 + i] - m_ref;
          p[i] = (POLY > 0 && i % (POLY > 0 ? POLY : 1) == (POLY > 0 ? POLY : 1) - 1) ? ex2_poly(x) : ex2_approx(x);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          pk[c * 4 + i] = pack_bf16x2(p[2 * i], p[2 * i + 1]);
          ls[i] += p[2 * i] + p[2 * i + 1];
        }
      }
      l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
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
          w[d] = pack_bf16x2(__uint_as_float(r[2 * d]) * gsc, __uint_as_float(r[2 * d + 1]) * gsc);
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

int launch_attn_time_tc(const TcAttnPlan* p, const float* gates, void* out, cudaStream_t st) {
  dim3 grid(ceil_div(p->L, AT_BQ), p->heads, p->seqs);
  // default: 4 CTAs/SM kernel (64-key tiles); BT_ATTN_VARIANT=128 selects the 2-CTAs/SM kernel (128-key tiles)
  static const int variant = getenv("BT_ATTN_VARIANT") ? atoi(getenv("BT_ATTN_VARIANT")) : 64;
  if (variant == 64) {
    static const int poly64 = getenv("BT_ATTN_POLY") ? atoi(getenv("BT_ATTN_POLY")) : 0;
    if (poly64 == 8)
      attn_tc64_kernel<8><<<grid, A6_THREADS, A6_SMEM, st>>>(p->tmQK, p->tmKV64, gates, reinterpret_cast<bf16*>(out), p->L, p->heads);
    else if (poly64 == 4)
      attn_tc64_kernel<4><<<grid, A6_THREADS, A6_SMEM, st>>>(p->tmQK, p->tmKV64, gates, reinterpret_cast<bf16*>(out), p->L, p->heads);
    else
      attn_tc64_kernel<0><<<grid, A6_THREADS, A6_SMEM, st>>>(p->tmQK, p->tmKV64, gates, reinterpret_cast<bf16*>(out), p->L, p->heads);
    return 0;
  }
  static const int poly = getenv("BT_ATTN_POLY") ? atoi(getenv("BT_ATTN_POLY")) : AT_POLY_MOD;
  static const int ks = getenv("BT_ATTN_KS") ? atoi(getenv("BT_ATTN_KS")) : 2;
#define BT_AT_L(K, A)                                                                                              \
  if (ks == K && poly == A) {                                                                                      \
    attn_tc_kernel<K, A><<<grid, 32 * (4 * K + 1), AT_SMEM, st>>>(p->tmQK, gates, reinterpret_cast<bf16*>(out), p->L, \
                                                                  p->heads);                                       \
    return 0;                                                                                                      \
  }
  BT_AT_L(2, 0) BT_AT_L(2, 4) BT_AT_L(4, 0)
#undef BT_AT_L
  return -3;
}

// ==================================================================== fused frontend FFN
// x += W2 gelu(W1 rmsnorm(x) + b1) + b2 for the narrow frontend FFNs (C = 32 / 64, hidden 4C) in ONE
// kernel (reference roformer.py:38-61): the hidden activations never leave the SM.  Unfused, this
// block streams 32 bytes per element through HBM (norm 6 + ff1 10 + ff2 16); fused it is 8.
// CTA = 128 tokens; warps 0-3: one token row per thread (RMSNorm, bias+GELU, residual), warp 4
// (converged, one elected lane issues): TMA (weights) + tcgen05.mma.  Hidden units are processed in chunks of 128:
//   H_h = Xn W1_h^T (N=128, K=C) -> TMEM cols [0,128) -> bias+GELU -> bf16 tile in smem ->
//   OUT (+)= H_h W2_h^T (N=C, K=128) -> TMEM cols [128,128+C).
constexpr int FF_THREADS = 160;
template <int C>
struct FfCfg {
  static constexpr int NH = 4 * C / 128;            // hidden chunks
  static constexpr int A_BYTES = 128 * C * 2;       // normalised tokens, K-major
  static constexpr int W1_BYTES = 4 * C * C * 2;    // all chunks resident
  static constexpr int W2C_BYTES = C * 128 * 2;     // one K-chunk of W2
  static constexpr int H_BYTES = 128 * 128 * 2;
  static constexpr int WO_BYTES = C * C * 2;       // attention out-projection weight (fused_ff_kernel<C, true>)
  static constexpr int SMEM = A_BYTES + W1_BYTES + W2C_BYTES + H_BYTES + WO_BYTES + 5 * C * 4 + 1024 + 128;
  static constexpr int SWZ_A = C * 2 < 128 ? C * 2 : 128;  // 64-byte rows for C=32, 128 for C=64
  // TMEM: H accumulator [0,128) and OUT accumulator.  With a single hidden chunk (C = 32) OUT reuses the H
  // columns (every thread has read H before MMA2 is issued) -> 128 columns, 3 CTAs/SM instead of 2.
  static constexpr int OUT_COL = NH == 1 ? 0 : 128;
  static constexpr int TCOLS = NH == 1 ? 128 : 256;
  static constexpr int CTAS = NH == 1 ? 3 : 2;
  // accumulator of the optional out-projection prologue (O Wo^T): columns that are dead at that point
  static constexpr int D0_COL = NH == 1 ? 64 : 0;
};

// OP = true: the attention out-projection is fused in front (reference roformer.py:134-140 followed by
// roformer.py:38-61): x' = x + O Wo^T is computed per tile by one more MMA (the gated attention output O is
// TMA-loaded into the A-tile buffer, which the normalised x' overwrites afterwards), then the FFN runs on x'.
// Saves the separate out-projection GEMM: one fp32 read + write of the residual stream per element.
template <int C, bool OP>
__global__ void __launch_bounds__(FF_THREADS, FfCfg<C>::CTAS)
fused_ff_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWo,
                float* __restrict__ X, const float* __restrict__ b1, const float* __restrict__ b2,
                bf16* __restrict__ xb_out, int64_t M) {
  // PERSISTENT: each CTA walks over token tiles (stride gridDim.x); W1 (and W2 when it is a single chunk) are
  // fetched once per CTA, barriers / TMEM / bias staging are set up once.  (The one-tile-per-CTA form spent
  // more time on set-up and on re-fetching 16-64 KB of weights per CTA than on its tile.)
  using Cfg = FfCfg<C>;
  constexpr int NH = Cfg::NH;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = sbase;
  const uint32_t sW1 = sA + Cfg::A_BYTES;
  const uint32_t sW2 = sW1 + Cfg::W1_BYTES;
  const uint32_t sH = sW2 + Cfg::W2C_BYTES;
  const uint32_t sWo = sH + Cfg::H_BYTES;
  const uint32_t sB = sWo + Cfg::WO_BYTES;        // b1[4C] | b2[C] fp32
  const uint32_t bar_w1 = sB + 5 * C * 4;
  const uint32_t bar_w2 = bar_w1 + 8;
  const uint32_t bar_a = bar_w2 + 8;
  const uint32_t bar_h = bar_a + 8;
  const uint32_t bar_h2 = bar_h + 8;
  const uint32_t bar_o = bar_h2 + 8;
  const uint32_t bar_of = bar_o + 8;              // O tile landed in the A buffer (OP)
  const uint32_t bar_d0 = bar_of + 8;             // O Wo^T accumulated (OP)
  const uint32_t tmem_slot = bar_d0 + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = static_cast<int>((M + 127) / 128);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    auto init = [](uint32_t bar, uint32_t count) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    };
    init(bar_w1, 1); init(bar_w2, 1); init(bar_a, 128); init(bar_h, 1); init(bar_h2, 128); init(bar_o, 1);
    init(bar_of, 1); init(bar_d0, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 5 * C; i += FF_THREADS)
    st_shared_f32(sB + 4 * i, i < 4 * C ? __ldg(b1 + i) : __ldg(b2 + i - 4 * C));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 4) {
    const uint32_t on = elect_one() ? 1u : 0u;  // converged issuer warp, predicated single-lane TMA / MMA (see umma_bf16_p)
    constexpr uint32_t idesc1 = make_idesc_bf16(128, 128);
    constexpr uint32_t idesc2 = make_idesc_bf16(128, C);
    // weights: W1 [4C, C] all chunks (boxes of 128 rows), W2 [C, 4C] one K-chunk at a time (two 64-wide boxes)
    mbar_expect_tx_p(on, bar_w1, Cfg::W1_BYTES + (OP ? Cfg::WO_BYTES : 0));
    for (int h = 0; h < NH; ++h) tma_load_2d_p(on, sW1 + h * (128 * C * 2), &tmW1, bar_w1, 0, h * 128);
    if constexpr (OP) tma_load_2d_p(on, sWo, &tmWo, bar_w1, 0, 0);
    auto load_o = [&](int tile) {  // gated attention output rows of a tile -> the A buffer (same box / swizzle)
      mbar_expect_tx_p(on, bar_of, Cfg::A_BYTES);
      tma_load_2d_p(on, sA, &tmO, bar_of, 0, tile * 128);
    };
    if constexpr (OP) {
      if (static_cast<int>(blockIdx.x) < ntiles) load_o(blockIdx.x);
    }
    auto load_w2 = [&](int h) {
      mbar_expect_tx_p(on, bar_w2, Cfg::W2C_BYTES);
      for (int a = 0; a < 2; ++a) tma_load_2d_p(on, sW2 + a * (C * 128), &tmW2, bar_w2, h * 128 + a * 64, 0);
    };
    load_w2(0);
    auto issue_mma1 = [&](int h) {
#pragma unroll
      for (int k = 0; k < C / 16; ++k)
        umma_bf16_p(on, tmem_base, make_kmajor_desc<Cfg::SWZ_A>(sA + k * 32),
                    make_kmajor_desc<Cfg::SWZ_A>(sW1 + h * (128 * C * 2) + k * 32), idesc1, k != 0 ? 1u : 0u);
      umma_commit_p(on, bar_h);
    };
    mbar_wait_a(bar_w1, 0);
    int idx = 0;      // chunk counter over all tiles of this CTA: parity of bar_h / bar_h2 / bar_o
    int w2_loads = 0; // completed-or-in-flight W2 chunk loads minus one: parity of bar_w2
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      if constexpr (OP) {  // D0 = O Wo^T into columns that nobody reads at this point
        constexpr uint32_t idesc0 = make_idesc_bf16(128, C);
        mbar_wait_a(bar_of, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < C / 16; ++k)
          umma_bf16_p(on, tmem_base + Cfg::D0_COL, make_kmajor_desc<Cfg::SWZ_A>(sA + k * 32),
                      make_kmajor_desc<Cfg::SWZ_A>(sWo + k * 32), idesc0, k != 0 ? 1u : 0u);
        umma_commit_p(on, bar_d0);
      }
      mbar_wait_a(bar_a, it & 1);  // normalised tile in smem (and every thread is done with the previous tile's TMEM)
      tc_fence_after();
      issue_mma1(0);
      for (int h = 0; h < NH; ++h, ++idx) {
        mbar_wait_a(bar_h2, idx & 1);  // bf16 H_h tile written, accumulator H consumed
        tc_fence_after();
        if (h + 1 < NH) issue_mma1(h + 1);
        if constexpr (OP) {  // the last MMA1 of this tile has completed (its H was read): the A buffer is free
          if (h == NH - 1 && tile + static_cast<int>(gridDim.x) < ntiles) load_o(tile + gridDim.x);
        }
        if (NH > 1 || idx == 0) mbar_wait_a(bar_w2, w2_loads & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_p(on, tmem_base + Cfg::OUT_COL, make_kmajor_desc<128>(sH + (k >> 2) * 16384 + (k & 3) * 32),
                      make_kmajor_desc<128>(sW2 + (k >> 2) * (C * 128) + (k & 3) * 32), idesc2, (h | k) != 0 ? 1u : 0u);
        umma_commit_p(on, bar_o);
        if (NH > 1 && (h + 1 < NH || tile + static_cast<int>(gridDim.x) < ntiles)) {
          mbar_wait_a(bar_o, idx & 1);  // MMA2 finished reading this W2 chunk (and the H tile)
          load_w2((h + 1) % NH);
          ++w2_loads;
        }
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t hrow = sH + row * 128;
    const uint32_t hsw = static_cast<uint32_t>(row & 7) << 4;
    int idx = 0, it = 0;
    constexpr bool PREFETCH = C == 32;  // next tile's row requested while this tile is in the MMAs (register budget: C = 32 only)
    float4 xn[PREFETCH ? C / 4 : 1];
    auto load_x = [&](int tile, float4* dst) {
      const int64_t mm = static_cast<int64_t>(tile) * 128 + row;
      const float4* xr = reinterpret_cast<const float4*>(X + (mm < M ? mm : 0) * C);
#pragma unroll
      for (int i = 0; i < C / 4; ++i) dst[i] = mm < M ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if constexpr (PREFETCH) {
      if (static_cast<int>(blockIdx.x) < ntiles) load_x(blockIdx.x, xn);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int64_t m = static_cast<int64_t>(tile) * 128 + row;
      const bool valid = m < M;
      // ---- RMSNorm of this token (x stays in registers for the residual) ----
      float x[C];
      {
        float4 xq[C / 4];
        if constexpr (PREFETCH) {
#pragma unroll
          for (int i = 0; i < C / 4; ++i) xq[i] = xn[i];
        } else {
          load_x(tile, xq);
        }
#pragma unroll
        for (int i = 0; i < C / 4; ++i) {
          const float4 q = xq[i];
          x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w;
        }
        if constexpr (OP) {  // x' = x + O Wo^T (attention residual)
          mbar_wait_a(bar_d0, it & 1);
          tc_fence_after();
#pragma unroll
          for (int c4 = 0; c4 < C / 32; ++c4) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + lane_base + Cfg::D0_COL + c4 * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) x[c4 * 32 + i] += __uint_as_float(r[i]);
          }
        }
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < C; ++i) ss = fmaf(x[i], x[i], ss);
        const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        constexpr int RB = C * 2;  // bytes per A row
        const uint32_t arow = sA + row * RB;
        const uint32_t sw = C == 32 ? (static_cast<uint32_t>((row >> 1) & 3) << 4) : (static_cast<uint32_t>(row & 7) << 4);
        // the A tile is free: the last MMA1 of the previous tile completed before its bar_h was observed
#pragma unroll
        for (int c = 0; c < C / 8; ++c)
          st_shared_v4(arow + ((c << 4) ^ sw), pack_bf16x2(x[8 * c] * inv, x[8 * c + 1] * inv),
                       pack_bf16x2(x[8 * c + 2] * inv, x[8 * c + 3] * inv), pack_bf16x2(x[8 * c + 4] * inv, x[8 * c + 5] * inv),
                       pack_bf16x2(x[8 * c + 6] * inv, x[8 * c + 7] * inv));
        fence_proxy_async_smem();
        tc_fence_before();  // this thread's TMEM reads of the previous tile are ordered before the next MMAs
        mbar_arrive_a(bar_a);
        if constexpr (PREFETCH) {
          if (tile + static_cast<int>(gridDim.x) < ntiles) load_x(tile + gridDim.x, xn);
        }
      }
      for (int h = 0; h < NH; ++h, ++idx) {
        mbar_wait_a(bar_h, idx & 1);
        tc_fence_after();
        if (h >= 1) {  // the single H tile is free once MMA2_{h-1} has completed (h == 0: waited at the end of the last tile)
          mbar_wait_a(bar_o, (idx - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + lane_base + c4 * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {  // 4 chunks of 8 hidden units
            const float4 ba = ld_shared_v4_f32(sB + 4 * (h * 128 + c4 * 32 + 8 * c));
            const float4 bb = ld_shared_v4_f32(sB + 4 * (h * 128 + c4 * 32 + 8 * c + 4));
            const float bq[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
            float g[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = gelu_tanh_fast(__uint_as_float(r[8 * c + i]) + bq[i]);
            const int cc = c4 * 4 + c;  // 16-byte chunk index inside the 128-wide row: atom = cc >> 3
            st_shared_v4(hrow + (cc >> 3) * 16384 + (((cc & 7) << 4) ^ hsw), pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]),
                         pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive_a(bar_h2);
      }
      mbar_wait_a(bar_o, (idx - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c4 = 0; c4 < C / 32; ++c4) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_base + Cfg::OUT_COL + c4 * 32, r);
        tmem_ld_wait();
        if (valid) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bq = ld_shared_v4_f32(sB + 4 * (4 * C + c4 * 32 + 4 * i));
            v[4 * i] = __uint_as_float(r[4 * i]) + bq.x + x[c4 * 32 + 4 * i];
            v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bq.y + x[c4 * 32 + 4 * i + 1];
            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bq.z + x[c4 * 32 + 4 * i + 2];
            v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bq.w + x[c4 * 32 + 4 * i + 3];
          }
          store_act<float, 32>(X + m * C + c4 * 32, v);
          if (xb_out) store_act<bf16, 32>(xb_out + m * C + c4 * 32, v);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<Cfg::TCOLS>(tmem_base);
}

struct TcFfPlan {
  CUtensorMap tmW1, tmW2, tmO, tmWo;
  int C;
  int64_t M;
  bool outproj;
};

// o_bf16 / wout_bf16 != nullptr: plan for the variant with the attention out-projection fused in front
// (o_bf16: gated attention output [M, C], wout_bf16: [C, C]).
TcFfPlan* tc_ff_plan_create(const void* w1_bf16, const void* w2_bf16, int C, int64_t M, const void* o_bf16,
                            const void* wout_bf16, char* err, int errlen) {
  if (C != 32 && C != 64) { snprintf(err, errlen, "fused ff: C must be 32 or 64"); return nullptr; }
  TcFfPlan* p = new TcFfPlan();
  p->C = C; p->M = M; p->outproj = o_bf16 != nullptr && wout_bf16 != nullptr;
  const uint32_t swz_a = C * 2 < 128 ? C * 2 : 128;
  {  // W1 [4C, C] row-major: box = {C, 128 rows}
    const uint64_t dims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(4 * C)};
    const uint64_t strides[1] = {static_cast<uint64_t>(C) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(C), 128};
    if (!make_tmap(&p->tmW1, w1_bf16, 2, dims, strides, box, swz_a, err, errlen)) { delete p; return nullptr; }
  }
  {  // W2 [C, 4C] row-major: box = {64 K, C rows}
    const uint64_t dims[2] = {static_cast<uint64_t>(4 * C), static_cast<uint64_t>(C)};
    const uint64_t strides[1] = {static_cast<uint64_t>(4 * C) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(C)};
    if (!make_tmap(&p->tmW2, w2_bf16, 2, dims, strides, box, 128, err, errlen)) { delete p; return nullptr; }
  }
  if (p->outproj) {
    {  // O [M, C] row-major: box = {C, 128 tokens}, same swizzle as the hand-written A tile
      const uint64_t dims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(M)};
      const uint64_t strides[1] = {static_cast<uint64_t>(C) * 2};
      const uint32_t box[2] = {static_cast<uint32_t>(C), 128};
      if (!make_tmap(&p->tmO, o_bf16, 2, dims, strides, box, swz_a, err, errlen)) { delete p; return nullptr; }
    }
    {  // Wo [C, C] row-major
      const uint64_t dims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(C)};
      const uint64_t strides[1] = {static_cast<uint64_t>(C) * 2};
      const uint32_t box[2] = {static_cast<uint32_t>(C), static_cast<uint32_t>(C)};
      if (!make_tmap(&p->tmWo, wout_bf16, 2, dims, strides, box, swz_a, err, errlen)) { delete p; return nullptr; }
    }
  } else {
    p->tmO = p->tmW1;  // never dereferenced
    p->tmWo = p->tmW1;
  }
  return p;
}
void tc_ff_plan_destroy(TcFfPlan* p) { delete p; }

int launch_fused_ff(const TcFfPlan* p, float* X, const float* b1, const float* b2, void* xb_out, cudaStream_t st) {
  const unsigned ntiles = static_cast<unsigned>((p->M + 127) / 128);
  const unsigned slots = static_cast<unsigned>(g_num_sms) * (p->C == 32 ? FfCfg<32>::CTAS : FfCfg<64>::CTAS);
  const unsigned grid = ntiles < slots ? ntiles : slots;  // persistent CTAs
  bf16* xb = reinterpret_cast<bf16*>(xb_out);
#define BT_FF_L(CC, OPP)                                                                                          \
  fused_ff_kernel<CC, OPP><<<grid, FF_THREADS, FfCfg<CC>::SMEM, st>>>(p->tmW1, p->tmW2, p->tmO, p->tmWo, X, b1, b2, xb, \
                                                                      p->M)
  if (p->C == 32) { if (p->outproj) BT_FF_L(32, true); else BT_FF_L(32, false); }
  else { if (p->outproj) BT_FF_L(64, true); else BT_FF_L(64, false); }
#undef BT_FF_L
  return 0;
}

// ================================================================ fused frontend QKV projection
// RMSNorm -> gates -> to_qkv GEMM -> RoPE (+ q scaling) for the narrow frontend attentions (C = 32 /
// 64) in one kernel (reference roformer.py:114-123,127-128): replaces norm_kernel + the QKV GEMM
// (16 bytes/element through HBM) by 4 in + 6 out.  CTA = 128 tokens; warps 0-3 one token row per
// thread, warp 4 (converged) TMA (weights) + tcgen05.mma.  N = 3C fits one MMA and 128/256 TMEM columns.
template <int C>
struct QkvCfg {
  static constexpr int A_BYTES = 128 * C * 2;
  static constexpr int W_BYTES = 3 * C * C * 2;
  static constexpr int SMEM = A_BYTES + W_BYTES + 1024 + 128;
  static constexpr int TCOLS = 3 * C <= 128 ? 128 : 256;
  static constexpr int SWZ = C * 2 < 128 ? C * 2 : 128;
};

template <int C>
__global__ void __launch_bounds__(FF_THREADS, (C == 32 ? 3 : 2))
fused_qkv_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ X, const float* __restrict__ wg,
                 const float* __restrict__ bg, const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                 bf16* __restrict__ qkv, float* __restrict__ gates, int64_t M, int L, int F, int posmode, float qscale) {
  using Cfg = QkvCfg<C>;
  constexpr int heads = C / 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = sbase;
  const uint32_t sW = sA + Cfg::A_BYTES;
  const uint32_t bar_w = sW + Cfg::W_BYTES;
  const uint32_t bar_a = bar_w + 8;
  const uint32_t bar_d = bar_a + 8;
  const uint32_t tmem_slot = bar_d + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = static_cast<int>((M + 127) / 128);  // PERSISTENT: tiles blockIdx.x, +gridDim.x, ... (W fetched once)

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmW);
    auto init = [](uint32_t bar, uint32_t count) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    };
    init(bar_w, 1); init(bar_a, 128); init(bar_d, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 4) {
    const uint32_t on = elect_one() ? 1u : 0u;  // converged issuer warp (see umma_bf16_p)
    constexpr uint32_t idesc = make_idesc_bf16(128, 3 * C);
    mbar_expect_tx_p(on, bar_w, Cfg::W_BYTES);
    tma_load_2d_p(on, sW, &tmW, bar_w, 0, 0);
    mbar_wait_a(bar_w, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      mbar_wait_a(bar_a, it & 1);  // normalised tile in smem, previous accumulator read by every thread
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < C / 16; ++k)
        umma_bf16_p(on, tmem_base, make_kmajor_desc<Cfg::SWZ>(sA + k * 32), make_kmajor_desc<Cfg::SWZ>(sW + k * 32), idesc,
                    k != 0 ? 1u : 0u);
      umma_commit_p(on, bar_d);
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    int it = 0;
    float4 xn[C / 4];  // next tile's row, requested while this tile is in the MMA / epilogue
    auto load_x = [&](int tile) {
      const int64_t mm = static_cast<int64_t>(tile) * 128 + row;
      const float4* xr = reinterpret_cast<const float4*>(X + (mm < M ? mm : 0) * C);
#pragma unroll
      for (int i = 0; i < C / 4; ++i) xn[i] = mm < M ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (static_cast<int>(blockIdx.x) < ntiles) load_x(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int64_t m = static_cast<int64_t>(tile) * 128 + row;
    const bool valid = m < M;
    {
      float x[C];
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < C / 4; ++i) {
        const float4 q = xn[i];
        x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w;
        ss = fmaf(q.x, q.x, ss); ss = fmaf(q.y, q.y, ss); ss = fmaf(q.z, q.z, ss); ss = fmaf(q.w, q.w, ss);
      }
      const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int i = 0; i < C; ++i) x[i] *= inv;
      // gates = sigmoid(to_gates(x_normed)) (gamma*sqrt(C) folded into wg)
#pragma unroll
      for (int h = 0; h < heads; ++h) {
        const float4* w4 = reinterpret_cast<const float4*>(wg + h * C);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < C / 4; ++i) {
          const float4 w = __ldg(w4 + i);
          a = fmaf(x[4 * i], w.x, a); a = fmaf(x[4 * i + 1], w.y, a); a = fmaf(x[4 * i + 2], w.z, a); a = fmaf(x[4 * i + 3], w.w, a);
        }
        if (valid) gates[m * heads + h] = sigmoidf_(a + __ldg(bg + h));
      }
      constexpr int RB = C * 2;
      const uint32_t arow = sA + row * RB;
      const uint32_t sw = C == 32 ? (static_cast<uint32_t>((row >> 1) & 3) << 4) : (static_cast<uint32_t>(row & 7) << 4);
#pragma unroll
      for (int c = 0; c < C / 8; ++c)
        st_shared_v4(arow + ((c << 4) ^ sw), pack_bf16x2(x[8 * c], x[8 * c + 1]), pack_bf16x2(x[8 * c + 2], x[8 * c + 3]),
                     pack_bf16x2(x[8 * c + 4], x[8 * c + 5]), pack_bf16x2(x[8 * c + 6], x[8 * c + 7]));
      fence_proxy_async_smem();
      tc_fence_before();  // TMEM reads of the previous tile are ordered before the next MMA
      mbar_arrive_a(bar_a);
      if (tile + static_cast<int>(gridDim.x) < ntiles) load_x(tile + gridDim.x);
    }
    // RoPE row of this token (interleaved pairs, rotary_embedding_torch semantics)
    float cs[16], sn[16];
    {
      const int pos = valid ? (posmode == 0 ? static_cast<int>(m % L) : static_cast<int>((m / L) % F)) : 0;
      const float4* c4 = reinterpret_cast<const float4*>(rope_cos + pos * 16);
      const float4* s4 = reinterpret_cast<const float4*>(rope_sin + pos * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = __ldg(c4 + i), b = __ldg(s4 + i);
        cs[4 * i] = a.x; cs[4 * i + 1] = a.y; cs[4 * i + 2] = a.z; cs[4 * i + 3] = a.w;
        sn[4 * i] = b.x; sn[4 * i + 1] = b.y; sn[4 * i + 2] = b.z; sn[4 * i + 3] = b.w;
      }
    }
    mbar_wait_a(bar_d, it & 1);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 3 * C / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + lane_base + c * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      const int which = (c * 32) / C;  // 0 q, 1 k, 2 v (compile-time after unrolling)
      if (which < 2) {
        const float sc = which == 0 ? qscale : 1.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = v[2 * i], x1 = v[2 * i + 1];
          v[2 * i] = (x0 * cs[i] - x1 * sn[i]) * sc;
          v[2 * i + 1] = (x1 * cs[i] + x0 * sn[i]) * sc;
        }
      }
      if (valid) store_act<bf16, 32>(qkv + m * (3 * C) + c * 32, v);
    }
    }  // tile loop
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<Cfg::TCOLS>(tmem_base);
}

struct TcQkvPlan {
  CUtensorMap tmW;
  int C;
  int64_t M;
};
TcQkvPlan* tc_qkv_plan_create(const void* wqkv_bf16, int C, int64_t M, char* err, int errlen) {
  if (C != 32 && C != 64) { snprintf(err, errlen, "fused qkv: C must be 32 or 64"); return nullptr; }
  TcQkvPlan* p = new TcQkvPlan();
  p->C = C; p->M = M;
  const uint64_t dims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(3 * C)};
  const uint64_t strides[1] = {static_cast<uint64_t>(C) * 2};
  const uint32_t box[2] = {static_cast<uint32_t>(C), static_cast<uint32_t>(3 * C)};
  if (!make_tmap(&p->tmW, wqkv_bf16, 2, dims, strides, box, C * 2 < 128 ? C * 2 : 128, err, errlen)) { delete p; return nullptr; }
  return p;
}
void tc_qkv_plan_destroy(TcQkvPlan* p) { delete p; }
int launch_fused_qkv(const TcQkvPlan* p, const float* X, const float* wg, const float* bg, const float* rope_cos,
                     const float* rope_sin, void* qkv, float* gates, int L, int F, int posmode, float qscale,
                     cudaStream_t st) {
  const unsigned ntiles = static_cast<unsigned>((p->M + 127) / 128);
  const unsigned slots = static_cast<unsigned>(g_num_sms) * (p->C == 32 ? 3u : 2u);
  const unsigned grid = ntiles < slots ? ntiles : slots;  // persistent CTAs
  bf16* q = reinterpret_cast<bf16*>(qkv);
  if (p->C == 32)
    fused_qkv_kernel<32><<<grid, FF_THREADS, QkvCfg<32>::SMEM, st>>>(p->tmW, X, wg, bg, rope_cos, rope_sin, q, gates, p->M, L, F, posmode, qscale);
  else
    fused_qkv_kernel<64><<<grid, FF_THREADS, QkvCfg<64>::SMEM, st>>>(p->tmW, X, wg, bg, rope_cos, rope_sin, q, gates, p->M, L, F, posmode, qscale);
  return 0;
}

int tc_init(char* err, int errlen) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t r = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (r != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      snprintf(err, errlen, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(r));
      return -1;
    }
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(attn_tc_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
  cudaFuncSetAttribute(attn_tc_kernel<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
  cudaFuncSetAttribute(attn_tc64_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6_SMEM);
  cudaFuncSetAttribute(attn_tc64_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6_SMEM);
  cudaFuncSetAttribute(attn_tc64_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6_SMEM);
  cudaError_t r = cudaFuncSetAttribute(attn_tc_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(fused_ff_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfCfg<32>::SMEM);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(fused_ff_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfCfg<64>::SMEM);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(fused_ff_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfCfg<32>::SMEM);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(fused_ff_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfCfg<64>::SMEM);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(fused_qkv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, QkvCfg<32>::SMEM);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(fused_qkv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, QkvCfg<64>::SMEM);
  if (r != cudaSuccess) {
    snprintf(err, errlen, "cudaFuncSetAttribute(attn_tc_kernel / fused_ff_kernel) failed: %s", cudaGetErrorString(r));
    return -1;
  }
  return 0;
}

}  // namespace bt
