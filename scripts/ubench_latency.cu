// Dependent-issue latencies on the device, one warp: DFMA, DADD, DMUL chains, MUFU.RCP64H + Newton, SHFL(64-bit),
// LDS pointer chase, __syncthreads at several CTA sizes.  Diagnostic for the latency-bound LM kernels.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double seed) {
  __shared__ int chain[256];
  __shared__ double sd[64];
  const int t = threadIdx.x;
  if (t < 256) chain[t] = (t * 7 + 3) & 255;
  __syncthreads();
  double a = seed, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  const int N = 512;
  if (t < 32) {
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, b, c);
    t1 = clock64(); if (t == 0) cyc[0] = (t1 - t0) ;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + c;
    t1 = clock64(); if (t == 0) cyc[1] = (t1 - t0);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a * b;
    t1 = clock64(); if (t == 0) cyc[2] = (t1 - t0);
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) a = 1.0 / (a + 2.0);
    t1 = clock64(); if (t == 0) cyc[3] = (t1 - t0);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + c;
    t1 = clock64(); if (t == 0) cyc[4] = (t1 - t0);
    int j = t;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) j = chain[j];
    t1 = clock64(); if (t == 0) cyc[5] = (t1 - t0);
    a += j;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) a = rsqrt(a + 2.0);
    t1 = clock64(); if (t == 0) cyc[6] = (t1 - t0);
    // store -> load through shared memory, same thread
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { sd[t] = a; __syncwarp(); a = sd[(t + 1) & 31] + c; __syncwarp(); }
    t1 = clock64(); if (t == 0) cyc[7] = (t1 - t0);
  }
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < 256; ++i) __syncthreads();
  t1 = clock64(); if (t == 0) cyc[8] = (t1 - t0) * 2;   // per-barrier x 512 to share the divisor below
  out[t] = a;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 8 * 1024); cudaMalloc(&c, 8 * 16);
  const char* names[9] = {"DFMA", "DADD", "DMUL", "1.0/x (IEEE div)", "SHFL64+DADD", "LDS chase", "rsqrt(double)", "STS->LDS+DADD (2 syncwarp)", "__syncthreads"};
  for (int nt : {32, 256, 512, 1024}) {
    lat<<<1, nt>>>(d, c, 1.5);
    long long h[16]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    printf("CTA of %d threads:", nt);
    for (int k = 0; k < 9; ++k) printf("  %s %.1f", names[k], h[k] / 512.0);
    printf("  (cycles per dependent op)\n");
  }
  return 0;
}
