// 16-bit tensor-core GEMM for sm_100a: D = A * W^T over "planes" with shifted slabs (linear layers, the
// k(2,3) frontend convolutions as implicit GEMM, frontend.linear), tcgen05.mma with TMEM accumulators,
// TMA (cp.async.bulk.tensor) operand staging through an mbarrier ring, persistent over output tiles,
// warp-specialised roles, fused epilogues (epilogue.cuh).  Reference call sites: every nn.Linear /
// Conv2d of beat_this/model/roformer.py:53-61,103-111 and beat_tracker.py:77,155-166.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace bt {

// --------------------------------------------------------------------------- tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
int g_num_sms = 148;

static bool make_tmap_any(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
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
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(tm, dt, rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu box %u,%u,%u stride0 %llu",
             static_cast<int>(r), rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
             (unsigned long long)(rank > 2 ? gd[2] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
             (unsigned long long)gs[0]);
    return false;
  }
  return true;
}
bool make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int swizzle_bytes, char* err, int errlen) {
  return make_tmap_any(tm, BT_H16_IS_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank,
                       dims, strides_bytes, box, swizzle_bytes, err, errlen);
}
bool make_tmap_f32(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes, char* err, int errlen) {
  return make_tmap_any(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle_bytes, err, errlen);
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

template <int BN, int BK>
__global__ void __launch_bounds__(TG_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const GemmShape g, const EpiParams e, int num_tiles, int t_tiles, int n_tiles, int m_tiles) {
  using Cfg = TgCfg<BN, BK>;
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
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], BN <= 64 ? TG_EPI_WARPS / 2 : TG_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TCOLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int kb_per_slab = g.Kslab / BK;
  const int num_kb = g.nslab * kb_per_slab;

  if (warp == 0) {
    // producer and MMA warps run CONVERGED with predicated single-lane TMA / MMA / commit instructions
    // (see umma_h16_p): in a divergent `if (lane == 0)` block every tcgen05.mma costs ~85 issue cycles.
    const uint32_t on = elect_one() ? 1u : 0u;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / n_tiles, nt = tile % n_tiles;
      const int p_out = mt / t_tiles;
      const int t0 = (mt - p_out * t_tiles) * TG_BM;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb / kb_per_slab;
        const int k0 = (kb - s * kb_per_slab) * BK;
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t fb = smem_u32(&full[stage]);
        mbar_expect_tx_p(on, fb, Cfg::STAGE_BYTES);
        tma_load_3d_p(on, smem_u32(sA + stage * Cfg::A_BYTES), &tmA, fb, k0, t0 + g.t_shift[s],
                      p_out * g.plane_mul + g.plane_add[s]);
        tma_load_2d_p(on, smem_u32(sW + stage * Cfg::W_BYTES), &tmW, fb, s * g.Kslab + k0, nt * BN);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t on = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc = make_idesc_h16(TG_BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
          umma_h16_p(on, d_tmem, make_kmajor_desc<Cfg::SWZ>(a_base + k * 32), make_kmajor_desc<Cfg::SWZ>(b_base + k * 32),
                     idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit_p(on, smem_u32(&empty[stage]));
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      if (TILE_SPLIT && (iter & 1) != half) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      const int mt = tile / n_tiles, nt = tile % n_tiles;
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
            epilogue_apply<h16, 32>(e, g.L, m, nt * BN + c * 32, v, cur, has_rope || has_resid, bias_smem);
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
  if (warp == 1) tmem_dealloc<Cfg::TCOLS>(tmem_base);
}

struct TcGemmPlan {
  CUtensorMap tmA, tmW;
  GemmShape g;
  int BN, BK;
  int num_tiles, t_tiles, n_tiles, m_tiles, grid;
};

static int pick_bn(int N) {
  const int cands[6] = {256, 192, 128, 96, 64, 32};
  for (int i = 0; i < 6; ++i)
    if (N % cands[i] == 0) return cands[i];
  return 0;
}

template <int BN, int BK>
static int gemm_tc_launch(const TcGemmPlan* p, const EpiParams& e, cudaStream_t st) {
  using Cfg = TgCfg<BN, BK>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t r = cudaFuncSetAttribute(gemm_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (r != cudaSuccess) return -1;
    attr_set = true;
  }
  gemm_tc_kernel<BN, BK><<<p->grid, TG_THREADS, Cfg::SMEM, st>>>(p->tmA, p->tmW, p->g, e, p->num_tiles, p->t_tiles,
                                                                   p->n_tiles, p->m_tiles);
  return 0;
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
    const uint32_t box[2] = {static_cast<uint32_t>(p->BK), static_cast<uint32_t>(p->BN)};
    if (!make_tmap(&p->tmW, W, 2, dims, strides, box, swz, err, errlen)) { delete p; return nullptr; }
  }
  p->t_tiles = ceil_div(g.L, TG_BM);
  p->n_tiles = g.N / p->BN;
  p->m_tiles = p->t_tiles * g.planes_out;
  p->num_tiles = p->m_tiles * p->n_tiles;
  p->grid = p->num_tiles < g_num_sms ? p->num_tiles : g_num_sms;
  return p;
}
void tc_gemm_plan_destroy(TcGemmPlan* p) { delete p; }

int launch_gemm_tc(const TcGemmPlan* p, const EpiParams& e, cudaStream_t st) {
#define BT_TG_CASE(bn, bk) \
  if (p->BN == bn && p->BK == bk) return gemm_tc_launch<bn, bk>(p, e, st);
  BT_TG_CASE(256, 64) BT_TG_CASE(192, 64) BT_TG_CASE(128, 64) BT_TG_CASE(96, 64) BT_TG_CASE(64, 64)
  BT_TG_CASE(32, 64) BT_TG_CASE(128, 32) BT_TG_CASE(96, 32) BT_TG_CASE(64, 32) BT_TG_CASE(32, 32)
#undef BT_TG_CASE
  return -2;
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
  if (tc_init_attn(err, errlen) != 0) return -1;
  if (tc_init_fused(err, errlen) != 0) return -1;
  return 0;
}

}  // namespace bt
