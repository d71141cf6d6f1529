// FP64 throughput microbenchmark for B200 (sm_100a): DFMA pipe vs DMMA (mma.sync.m8n8k4.f64) vs shared-memory-fed DFMA.
// Usage: ./fp64_peak   -> prints TFLOP/s for each variant.  Numbers go into DESIGN.md / profiles/.
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_dfma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
  const double a = 1.0000001 * 0.25, b = 1.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma(c[j][0], c[j][1], a, b);
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA with both operands streamed from shared memory (the naive warp-matmul pattern): 2 LDS.64 per FMA
__global__ void k_dfma_smem(double* out, int iters) {
  __shared__ double sa[1024], sb[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sa[i] = 1.0 + i * 1e-9; sb[i] = 1e-9 * i; }
  __syncthreads();
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const int l = threadIdx.x & 31;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a0 = fma(sa[(l + k) & 1023], sb[(k * 32 + l) & 1023], a0);
      a1 = fma(sa[(l + k + 8) & 1023], sb[(k * 32 + l + 256) & 1023], a1);
      a2 = fma(sa[(l + k + 16) & 1023], sb[(k * 32 + l + 512) & 1023], a2);
      a3 = fma(sa[(l + k + 24) & 1023], sb[(k * 32 + l + 768) & 1023], a3);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

// DMMA with fragments loaded from shared memory each time (1 LDS.64 per operand per mma)
__global__ void k_dmma_smem(double* out, int iters) {
  __shared__ double sa[24 * 28], sb[24 * 28];
  for (int i = threadIdx.x; i < 24 * 28; i += blockDim.x) { sa[i] = 0.25 + i * 1e-9; sb[i] = 1.0 + 1e-9 * i; }
  __syncthreads();
  const int l = threadIdx.x & 31;
  double c[9][2];
  for (int i = 0; i < 9; ++i) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double af[3], bf[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) { af[r] = sa[(8 * r + l / 4) * 28 + 4 * k + (l & 3)]; bf[r] = sb[(8 * r + l / 4) * 28 + 4 * k + (l & 3)]; }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) dmma(c[3 * r + cc][0], c[3 * r + cc][1], af[r], bf[cc]);
    }
  }
  double s = 0; for (int i = 0; i < 9; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  const int blocks = p.multiProcessorCount * 8, threads = 256;
  double* out; CHECK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    int iters = 20000;
    cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("dfma       : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * 8 * iters * (double)blocks * threads / ms * 1e-9);
    iters = 5000;
    cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("dmma       : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32) / ms * 1e-9);
    iters = 2000;
    cudaEventRecord(e0); k_dfma_smem<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("dfma_smem  : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * 32 * iters * (double)blocks * threads / ms * 1e-9);
    iters = 1000;
    cudaEventRecord(e0); k_dmma_smem<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("dmma_smem  : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * 256 * 54 * iters * (double)blocks * (threads / 32) / ms * 1e-9);
  }
  // occupancy sensitivity: DFMA with 1 warp per SMSP
  {
    int iters = 20000;
    cudaEventRecord(e0); k_dfma<<<p.multiProcessorCount, 128>>>(out, iters); cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("dfma 4w/SM : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * 8 * iters * (double)p.multiProcessorCount * 128 / ms * 1e-9);
    iters = 5000;
    cudaEventRecord(e0); k_dmma<<<p.multiProcessorCount, 128>>>(out, iters); cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("dmma 4w/SM : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * 256 * 8 * iters * (double)p.multiProcessorCount * 4 / ms * 1e-9);
  }
  return 0;
}
