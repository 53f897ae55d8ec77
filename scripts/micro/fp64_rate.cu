// FP64 issue rate per SM on this part (diagnostic): W warps per CTA, one CTA per SM, every thread
// runs C independent chains of N dependent DADDs (or DFMAs); reports warp-instructions / clk / SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int C, bool FMA, typename T>
__global__ void k(T *out, int n, long long *cyc) {
  T x[C];
  for (int c = 0; c < C; ++c) x[c] = (T)(threadIdx.x + c) * (T)1e-3;
  const T a = (T)1.0000001, b = (T)1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = FMA ? (x[c] * a + b) : (x[c] + b);
  }
  long long t1 = clock64();
  T s = 0;
  for (int c = 0; c < C; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int C, bool FMA, typename T>
void run(const char *name, int warps) {
  T *out; long long *cyc;
  cudaMalloc(&out, sizeof(T) * 148 * 1024); cudaMalloc(&cyc, 8 * 148);
  const int n = 4096;
  k<C, FMA, T><<<148, 32 * warps>>>(out, n, cyc);
  k<C, FMA, T><<<148, 32 * warps>>>(out, n, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%-22s warps/SM %2d chains %d: %.1f cycles per instruction per warp, %.3f warp-instr/clk/SM\n", name, warps, C,
         c / n / C, (double)n * C * warps / c);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 2, 4, 8, 16, 32}) run<1, false, double>("DADD dependent", w);
  for (int w : {1, 4, 8, 16}) run<8, false, double>("DADD 8 chains", w);
  for (int w : {1, 4, 8, 16}) run<8, true, double>("DFMA 8 chains", w);
  for (int w : {1, 4, 16}) run<8, true, float>("FFMA 8 chains", w);
  return 0;
}
