// Instruction-throughput microbenchmark (per SM, 32 resident warps): which pipe do the softmax
// instructions of attn_tc_kernel share?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// tools/ubench_pipes.cu -o tools/_build/ubench_pipes ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, int iters) {
  float a[8];
  uint32_t u[8];
  uint64_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 1e-3f + i; u[i] = threadIdx.x + i;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(w[i]) : "f"(a[i]), "f"(a[i] + 0.5f));
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (OP == 2) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (OP == 3) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (OP == 5) {  // the softmax mix per 2 elements: 2 sub, 2 ex2, 1 pack
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a[(i + 2) & 7]) : "f"(a[(i + 3) & 7]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 2) & 7]));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 2) & 7]));
      }
      if (OP == 6) {  // ex2 + pack only (do they share a pipe?)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      }
      if (OP == 7) asm volatile("max.f32 %0, %0, %1; max.f32 %0, %0, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (OP == 8) asm volatile("shl.b32 %0, %0, 23; add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == 9) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (OP == 10) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (OP == 11) {  // f16x2 softmax mix per 2 elements: 2 sub, 1 pack, 1 ex2.f16x2
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a[(i + 2) & 7]) : "f"(a[(i + 3) & 7]));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 2) & 7]));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      }
      if (OP == 12) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 13) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      // round 2: packed fp32 pairs and 3-input max (sm_100)
      if (OP == 14) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) & 7]));
      if (OP == 15) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(w[i]) : "l"(w[(i + 1) & 7]));
      if (OP == 16) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (OP == 17) asm volatile("mad.lo.s32 %0, %0, 8388608, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == 18) {  // FADD2 next to MUFU: separate pipes?
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) & 7]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      }
      if (OP == 19) {  // FFMA2 next to scalar FFMA: same pipe?
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(w[i]) : "l"(w[(i + 1) & 7]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0; uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += a[i]; x ^= u[i]; x ^= static_cast<uint32_t>(w[i]) ^ static_cast<uint32_t>(w[i] >> 32); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_inner, float* out, long long* cyc) {
  const int iters = 2000;
  k<OP><<<148, 1024>>>(out, cyc, iters);
  k<OP><<<148, 1024>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double thread_ops = 1024.0 * iters * 8 * ops_per_inner;
  printf("%-28s %8.1f thread-instr/clk/SM  (%.0f cycles)\n", name, thread_ops / c, c);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("ex2.approx.ftz.f32", 1, out, cyc);
  run<1>("cvt.rn.bf16x2.f32", 1, out, cyc);
  run<9>("cvt.rn.f16x2.f32", 1, out, cyc);
  run<2>("max.f32", 1, out, cyc);
  run<7>("max3 (2x max.f32 fused?)", 2, out, cyc);
  run<3>("add.f32", 1, out, cyc);
  run<4>("fma.f32", 1, out, cyc);
  run<8>("shl+add.s32", 2, out, cyc);
  run<10>("ex2.approx.f16x2 (instr)", 1, out, cyc);
  run<13>("ex2.approx.ftz.bf16x2 (instr)", 1, out, cyc);
  run<12>("tanh.approx.f32", 1, out, cyc);
  run<11>("2 sub + pack + ex2.f16x2", 4, out, cyc);
  run<6>("ex2 + cvt pair", 2, out, cyc);
  run<5>("2 sub + 2 ex2 + 1 pack", 5, out, cyc);
  run<14>("add.f32x2 (instr)", 1, out, cyc);
  run<15>("fma.f32x2 (instr)", 1, out, cyc);
  run<16>("max.f32 3-input", 1, out, cyc);
  run<17>("mad.lo.s32 x 2^23 + c", 1, out, cyc);
  run<18>("add.f32x2 + ex2", 2, out, cyc);
  run<19>("fma.f32x2 + fma.f32", 2, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
