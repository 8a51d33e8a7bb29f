// dmma_lds.cu -- DMMA.8x8x4 issue rate when every DMMA is fed by L shared-memory loads (LDS.64) of fresh operands:
// how many operand loads per DMMA the SM sustains before the FP64 tensor pipe starves.  4 CTAs x 7 warps per SM,
// conflict-free lane-contiguous loads (the access pattern of das_ns_kernel's fragment-ordered tiles).
#include <cstdio>
#include <cuda_runtime.h>

template <int NL, int ND>   // per step: NL loads, ND DMMAs (4 accumulator chains)
__global__ void __launch_bounds__(224, 4) k(double *out, int iters) {
  __shared__ double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
  __syncthreads();
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  const int lane = threadIdx.x & 31;
  const double *base = sm + lane;
  double v[NL > 0 ? NL : 1];
#pragma unroll
  for (int j = 0; j < (NL > 0 ? NL : 1); ++j) v[j] = 1.0;
  for (int it = 0; it < iters; ++it) {
    const double *b2 = base + ((it & 15) << 6);
#pragma unroll
    for (int j = 0; j < NL; ++j) v[j] = b2[32 * j];
#pragma unroll
    for (int i = 0; i < ND; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i & 3][0]), "+d"(c[i & 3][1]) : "d"(v[(NL > 0) ? (i % (NL > 0 ? NL : 1)) : 0]), "d"(v[(NL > 1) ? ((i + 1) % NL) : 0]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NL, int ND>
void run(double *out, int sms) {
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    float ms;
    cudaEventRecord(e0);
    k<NL, ND><<<sms * 4, 224>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  printf("LDS per step %2d, DMMA per step %2d (%.2f LDS/DMMA): %.2f TFLOP/s\n", NL, ND, (double)NL / ND,
         512.0 * iters * ND * 28 * sms / best * 1e-9);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 4 * 224);
  run<0, 8>(out, p.multiProcessorCount);
  run<4, 8>(out, p.multiProcessorCount);
  run<6, 8>(out, p.multiProcessorCount);
  run<8, 8>(out, p.multiProcessorCount);
  run<10, 8>(out, p.multiProcessorCount);
  run<12, 8>(out, p.multiProcessorCount);
  run<16, 8>(out, p.multiProcessorCount);
  run<6, 4>(out, p.multiProcessorCount);
  return 0;
}
