// FP64 pipe microbenchmark for sm_100a: dependent-chain latency and throughput vs warps/SM and ILP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, double* out, long long* cyc) {
  const int iters = 4096;
  chain<ILP><<<148, 32 * warps_per_sm>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  c /= 148;
  const double per_inst = c / (double)(iters * ILP);                 // cycles per DFMA of one warp
  const double sm_rate = 32.0 * warps_per_sm * iters * ILP / c;      // thread-DFMA per clk per SM
  printf("warps/SM %2d ILP %d: %.2f clk per warp-DFMA, %.1f DFMA/clk/SM\n", warps_per_sm, ILP, per_inst, sm_rate);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 148 * 8);
  for (int w : {1, 4, 8, 16, 20, 24, 32}) {
    run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc);
  }
  return 0;
}
