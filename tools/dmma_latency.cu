// dmma_latency.cu -- issue rate of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) against the number of independent accumulator
// chains per warp and the number of warps per SM sub-partition: tells whether a kernel with W resident warps per
// sub-partition and C interleaved chains per warp can saturate the FP64 tensor pipe.  Prints a table and one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k(double *out, int iters) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
float run(double *out, int sms, int warps_per_sm, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    float ms;
    cudaEventRecord(e0);
    k<CH><<<sms, 32 * warps_per_sm>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 4096);
  const int iters = 20000;
  printf("clock %.0f MHz (attribute; the SM may run lower under load)\n", khz * 1e-3);
  printf("%8s %8s %14s %18s\n", "warps/SP", "chains", "TFLOP/s", "cycles/DMMA/warp");
  const int wl[] = {1, 2, 4, 7, 8};
  for (int wi = 0; wi < 5; ++wi) {
    const int wsp = wl[wi], wsm = 4 * wsp;
    float ms[4] = {run<1>(out, p.multiProcessorCount, wsm, iters), run<2>(out, p.multiProcessorCount, wsm, iters),
                   run<4>(out, p.multiProcessorCount, wsm, iters), run<8>(out, p.multiProcessorCount, wsm, iters)};
    const int ch[4] = {1, 2, 4, 8};
    for (int i = 0; i < 4; ++i) {
      const double n = (double)iters * ch[i];   // DMMAs per warp
      const double tf = 512.0 * n * wsm * p.multiProcessorCount / ms[i] * 1e-9;
      const double cyc = ms[i] * 1e-3 * khz * 1e3 / n;
      printf("%8d %8d %14.2f %18.1f\n", wsp, ch[i], tf, cyc);
    }
  }
  // warp -> sub-partition mapping: 4 CTAs of 7 warps per SM against 4 CTAs of 8 warps (same work per warp)
  for (int wpc = 5; wpc <= 8; ++wpc) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      float ms;
      cudaEventRecord(e0);
      k<4><<<p.multiProcessorCount * 4, 32 * wpc>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("4 CTAs/SM x %d warps: %.2f TFLOP/s\n", wpc, 512.0 * iters * 4 * wpc * 4 * p.multiProcessorCount / best * 1e-9);
  }
  return 0;
}
