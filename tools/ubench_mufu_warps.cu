// How many warps per SM sub-partition does the softmax exponential stream need to fill the MUFU pipe?
// One block per SM of W warps (W = 4 -> one warp per sub-partition); every thread runs the per-tile sequence of
// attn_tc48_kernel (24 x: packed subtract, two ex2, pack to fp16, packed row-sum) on 48 registers.
//   MODE 0: pack / sum right behind their exponentials (what ptxas emits for the unrolled tile loop)
//   MODE 1: pack / sum of the PREVIOUS iteration's exponentials (loop-carried: a full tile of distance)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench_mufu_warps.cu -o tools/_build/ubench_mufu_warps
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packh(float lo, float hi) { __half2 p = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&p); }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float s[48], e[48];
#pragma unroll
  for (int i = 0; i < 48; ++i) { s[i] = -(threadIdx.x * 1e-3f + i * 0.01f); e[i] = 0.f; }
  uint64_t ls[2] = {0ull, 0ull};
  uint32_t acc = 0;
  float m = 0.25f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const uint64_t m2 = pack2(m, m);
    if (MODE == 0) {
#pragma unroll
      for (int q = 0; q < 24; ++q) {
        float x0, x1;
        unpack2(sub2(pack2(s[2 * q], s[2 * q + 1]), m2), x0, x1);
        const float p0 = ex2(x0), p1 = ex2(x1);
        acc ^= packh(p0, p1);
        ls[q & 1] = add2(ls[q & 1], pack2(p0, p1));
      }
    } else {
#pragma unroll
      for (int q = 0; q < 24; ++q) {
        acc ^= packh(e[2 * q], e[2 * q + 1]);
        ls[q & 1] = add2(ls[q & 1], pack2(e[2 * q], e[2 * q + 1]));
      }
#pragma unroll
      for (int q = 0; q < 24; ++q) {
        float x0, x1;
        unpack2(sub2(pack2(s[2 * q], s[2 * q + 1]), m2), x0, x1);
        e[2 * q] = ex2(x0); e[2 * q + 1] = ex2(x1);
      }
    }
    m += 1e-6f;
  }
  const long long t1 = clock64();
  float a0, a1;
  unpack2(add2(ls[0], ls[1]), a0, a1);
  float r = a0 + a1 + __uint_as_float(acc & 0x007fffffu);
#pragma unroll
  for (int i = 0; i < 48; ++i) r += e[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int warps, float* out, long long* cyc) {
  const int iters = 4000;
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("mode %d, %2d warps/SM: %.2f ex2/clk/SM (pipe: 16)\n", MODE, warps, warps * 32.0 * iters * 48 / c);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int w : {4, 8, 12, 16, 20}) run<0>(w, out, cyc);
  for (int w : {4, 8, 12, 16, 20}) run<1>(w, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
