// fp64_mix.cu -- micro-benchmarks behind the assembly epilogue design: how fast can the
// FP64 pipe of one SM retire the table-based exponential at the occupancy / ILP the kernel has?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/fp64_mix scripts/ubench/fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 512;

template <int V, int MODE>
__global__ void __launch_bounds__(256, 4) k_exp(double *out, const double *in, double c1) {
  __shared__ double tab[64];
  if (threadIdx.x < 64) tab[threadIdx.x] = 1.0 + threadIdx.x / 64.0;
  __syncthreads();
  double acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = in[threadIdx.x + v * 256];
  double sum = 0.0;
  const double L2E64 = 92.33248261689366, SHIFT = 6755399441055744.0;
  const double LN2HI64 = 0.01083042469326756, LN2LO64 = 2.9815858269852933e-12;
  for (int it = 0; it < ITERS; ++it) {
    double r[V], w[V];
    int n[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const double x = -c1 * acc[v];
      const double t = fma(x, L2E64, SHIFT);
      n[v] = __double2loint(t);
      const double nd = t - SHIFT;
      r[v] = fma(nd, -LN2LO64, fma(nd, -LN2HI64, x));
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      double p = fma(r[v], 8.3333333333333332e-03, 4.1666666666666664e-02);
      p = fma(p, r[v], 1.6666666666666666e-01);
      p = fma(p, r[v], 0.5);
      p = fma(p, r[v], 1.0);
      p = fma(p, r[v], 1.0);
      if (MODE == 0) {  // full: table + exponent insertion + range select
        const double ww = tab[n[v] & 63] * p;
        const int hi = __double2hiint(ww) + ((n[v] >> 6) << 20);
        w[v] = (n[v] < -64634) ? 0.0 : __hiloint2double(hi, __double2loint(ww));
      } else if (MODE == 1) {  // no table
        const double ww = 1.25 * p;
        const int hi = __double2hiint(ww) + ((n[v] >> 6) << 20);
        w[v] = (n[v] < -64634) ? 0.0 : __hiloint2double(hi, __double2loint(ww));
      } else {  // FP64 only
        w[v] = p * 1.25;
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      sum += w[v];          // 1 DADD
      acc[v] = acc[v] + 1e-9;  // 1 DADD keeps the chain data dependent on it
    }
  }
  if (sum == 123.456) out[blockIdx.x * 256 + threadIdx.x] = sum;
}

// plain Horner chains, register operands
template <int V>
__global__ void __launch_bounds__(256, 4) k_horner(double *out, double a, double b) {
  double c[V];
#pragma unroll
  for (int v = 0; v < V; ++v) c[v] = threadIdx.x + v;
  for (int it = 0; it < ITERS * 4; ++it)
#pragma unroll
    for (int v = 0; v < V; ++v) c[v] = fma(c[v], a, b);
  double s = 0;
#pragma unroll
  for (int v = 0; v < V; ++v) s += c[v];
  if (s == 123.456) out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double *out, *in;
  cudaMalloc(&out, 1 << 24);
  cudaMalloc(&in, 1 << 24);
  cudaMemset(in, 0, 1 << 24);
  const double clk = p.clockRate * 1e3;  // Hz (nominal max)
  for (int cps = 1; cps <= 4; ++cps) {
    const int ctas = sms * cps * 8;  // 8 waves of cps resident CTAs per SM (residency forced through dynamic smem)
    const int smem = 200 * 1024 / cps;
    cudaFuncSetAttribute(k_horner<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_horner<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_exp<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_exp<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_exp<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_exp<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_exp<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    auto rep = [&](const char *name, float ms, double fp64_per_thread_iter, int iters) {
      const double inst = (double)ctas * 8 * fp64_per_thread_iter * iters;  // warp instructions
      printf("%-28s resident CTAs/SM %d: %.3f ms  FP64 warp-instr/clk/SM %.3f (pipe limit 2.0)\n", name, cps, ms,
             inst / (ms * 1e-3) / clk / sms);
    };
    // residency is fixed by launch bounds; vary grid only -> use cps to scale work, report per-SM rate
    rep("horner V=4 regs", time_ms([&] { k_horner<4><<<ctas, 256, smem>>>(out, 1.0000001, 1e-9); }), 4, ITERS * 4);
    rep("horner V=8 regs", time_ms([&] { k_horner<8><<<ctas, 256, smem>>>(out, 1.0000001, 1e-9); }), 8, ITERS * 4);
    rep("exp V=4 full", time_ms([&] { k_exp<4, 0><<<ctas, 256, smem>>>(out, in, 0.5); }), 4 * 13, ITERS);
    rep("exp V=4 no table", time_ms([&] { k_exp<4, 1><<<ctas, 256, smem>>>(out, in, 0.5); }), 4 * 13, ITERS);
    rep("exp V=4 fp64 only", time_ms([&] { k_exp<4, 2><<<ctas, 256, smem>>>(out, in, 0.5); }), 4 * 13, ITERS);
    rep("exp V=8 full", time_ms([&] { k_exp<8, 0><<<ctas, 256, smem>>>(out, in, 0.5); }), 8 * 13, ITERS);
    rep("exp V=8 fp64 only", time_ms([&] { k_exp<8, 2><<<ctas, 256, smem>>>(out, in, 0.5); }), 8 * 13, ITERS);
  }
  printf("clock %.0f MHz, %d SMs\n", clk / 1e6, sms);
  return 0;
}
