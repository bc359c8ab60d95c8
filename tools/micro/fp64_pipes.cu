// Microbenchmark: FP64 throughput of DFMA vs DMMA (mma.sync f64) on sm_100a, cycles per warp instruction per SM.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096;

__global__ void dfma_kernel(long long* out, double* sink, double x)
{
  double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
    a0 = fma(a0, x, x); a1 = fma(a1, x, x); a2 = fma(a2, x, x); a3 = fma(a3, x, x);
    a4 = fma(a4, x, x); a5 = fma(a5, x, x); a6 = fma(a6, x, x); a7 = fma(a7, x, x);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 1.2345) sink[0] = a0;
}

__global__ void dmma884_kernel(long long* out, double* sink, double x)
{
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double a = x + threadIdx.x, b = x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (c0[0] + c1[0] + c2[1] + c3[1] == 1.2345) sink[0] = c0[0];
}

#ifdef HAVE_M16
__global__ void dmma16816_kernel(long long* out, double* sink, double x)
{
  double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = x + i + threadIdx.x;
  for (int i = 0; i < 4; i++) b[i] = x - i;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c0[0]), "+d"(c0[1]), "+d"(c0[2]), "+d"(c0[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c1[0]), "+d"(c1[1]), "+d"(c1[2]), "+d"(c1[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (c0[0] + c1[0] == 1.2345) sink[0] = c0[0];
}
#endif

int main()
{
  long long* d_out; double* d_sink;
  const int nsm = 148;
  cudaMalloc(&d_out, nsm * 8); cudaMalloc(&d_sink, 8);
  long long h[148];
  for (int warps : {4, 8, 16, 32}) {
    for (int k = 0; k < 3; k++) {
      const char* name; double per_iter;
      for (int rep = 0; rep < 2; rep++) {
        if (k == 0) dfma_kernel<<<nsm, warps * 32>>>(d_out, d_sink, 1.0000001);
        if (k == 1) dmma884_kernel<<<nsm, warps * 32>>>(d_out, d_sink, 1.0000001);
#ifdef HAVE_M16
        if (k == 2) dmma16816_kernel<<<nsm, warps * 32>>>(d_out, d_sink, 1.0000001);
#endif
      }
      cudaDeviceSynchronize();
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      double cyc = 0; for (int i = 0; i < nsm; i++) cyc += h[i]; cyc /= nsm;
      if (k == 0) { name = "DFMA"; per_iter = 8; }
      else if (k == 1) { name = "DMMA m8n8k4"; per_iter = 4; }
      else { name = "DMMA m16n8k16"; per_iter = 2; }
      double cpi = cyc / (ITER * per_iter * warps);
      double fma_per_inst = k == 0 ? 32 : (k == 1 ? 256 : 2048);
      printf("warps/SM %2d  %-14s cycles/warp-inst/SM %.2f  -> %.1f FMA/clk/SM\n", warps, name, cpi, fma_per_inst / cpi);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
