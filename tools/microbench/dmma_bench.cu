// dmma_bench.cu -- throughput/latency of FP64 mma.sync (DMMA) vs DFMA on sm_100a, to decide which pieces of
// the WBC tick (A^T A, blocked factor updates) go to the FP64 tensor path.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#ifdef USE_M16
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
#endif

template <int CH>
__global__ void k_dmma(double* out, int iters) {
  double c0[CH], c1[CH];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
  for (int i = 0; i < CH; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef USE_M16
template <int CH>
__global__ void k_dmma16(double* out, int iters) {
  double c[CH][4], a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-6 * i;
#pragma unroll
  for (int i = 0; i < CH; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

template <int CH>
__global__ void k_dfma(double* out, int iters) {
  double c[CH];
  double a = 1.0000001, b = 1e-9;
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 1 << 14;
  printf("SMs %d clock %d kHz\n", sms, clk);
  // throughput: many warps, 8 chains
  for (int warps = 1; warps <= 16; warps *= 2) {
    float ms = timeit([&] { k_dmma<8><<<sms, warps * 32, 0>>>(out, iters); });
    double inst = (double)iters * 8 * warps * sms;
    printf("DMMA.884  warps/SM %2d chains 8: %.3f ms  %.2f TFLOP/s  %.2f cyc/inst/SM\n", warps, ms, inst * 512 / ms * 1e-9, ms * 1e-3 * clk * 1e3 / (iters * 8.0 * warps));
  }
  { float ms = timeit([&] { k_dmma<1><<<sms, 32, 0>>>(out, iters); });
    printf("DMMA.884 latency (1 warp, dependent chain): %.1f cycles\n", ms * 1e-3 * clk * 1e3 / iters); }
  { float ms = timeit([&] { k_dmma<2><<<sms, 32, 0>>>(out, iters); });
    printf("DMMA.884 1 warp 2 chains: %.1f cycles/inst\n", ms * 1e-3 * clk * 1e3 / iters / 2); }
  { float ms = timeit([&] { k_dmma<4><<<sms, 32, 0>>>(out, iters); });
    printf("DMMA.884 1 warp 4 chains: %.1f cycles/inst\n", ms * 1e-3 * clk * 1e3 / iters / 4); }
#ifdef USE_M16
  for (int warps = 1; warps <= 16; warps *= 4) {
    float ms = timeit([&] { k_dmma16<4><<<sms, warps * 32, 0>>>(out, iters); });
    double inst = (double)iters * 4 * warps * sms;
    printf("DMMA.16816 warps/SM %2d chains 4: %.3f ms  %.2f TFLOP/s %.2f cyc/inst/SM\n", warps, ms, inst * 4096 / ms * 1e-9, ms * 1e-3 * clk * 1e3 / (iters * 4.0 * warps));
  }
  { float ms = timeit([&] { k_dmma16<1><<<sms, 32, 0>>>(out, iters); });
    printf("DMMA.16816 latency: %.1f cycles\n", ms * 1e-3 * clk * 1e3 / iters); }
#endif
  for (int warps = 1; warps <= 32; warps *= 2) {
    float ms = timeit([&] { k_dfma<8><<<sms, warps * 32, 0>>>(out, iters); });
    double inst = (double)iters * 8 * warps * sms;
    printf("DFMA      warps/SM %2d chains 8: %.3f ms  %.2f TFLOP/s  %.2f cyc/inst/SM\n", warps, ms, inst * 64 / ms * 1e-9, ms * 1e-3 * clk * 1e3 / (iters * 8.0 * warps));
  }
  { float ms = timeit([&] { k_dfma<1><<<sms, 32, 0>>>(out, iters); });
    printf("DFMA latency: %.1f cycles\n", ms * 1e-3 * clk * 1e3 / iters); }
  cudaFree(out);
  return 0;
}
