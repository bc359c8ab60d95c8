// Microbenchmark: shared-memory wavefront cost of broadcast-heavy load patterns on sm_100a.
// For each (width, pattern) runs 16 warps/SM of back-to-back independent LDS and reports
// SM cycles per warp-level load instruction (= wavefronts when the LSU data pipe is the limit).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048;
constexpr int UNROLL = 8;

__device__ __forceinline__ int pattern_addr(int pat, int lane, int width)
{
  // returns byte offset (multiple of width)
  const int half = lane >> 4, a = (lane >> 2) & 3, b = lane & 3;
  switch (pat) {
  case 0: return 0;                                  // uniform
  case 1: return half * 432;                         // half-uniform, records 432 B apart
  case 2: return half * 432 + a * width;             // a-pattern (adjacent groups of 4 share)
  case 3: return half * 432 + b * width;             // b-pattern (stride-4 lanes share)
  case 4: return lane * width;                       // all distinct, contiguous
  case 5: return (lane & 1) * 48;                    // 2 addresses alternating lanes
  case 6: return ((lane * 7) & 3) * 48;              // 4 addresses scattered over lanes
  case 7: return (lane >> 3) * 48;                   // quarter-uniform, 4 addresses
  case 8: return half * 432 + b * 2 * width;         // b-pattern with stride 2*width
  case 9: return (lane >> 1) * width;                // pairs share, 16 distinct contiguous
  case 10: return half * 448 + b * width;            // b-pattern, records 448 B apart (bank shift 16)
  case 11: return half * 64 + b * width;             // b-pattern, halves adjacent (contiguous 128 B for w=16)
  case 12: return ((lane * 5) & 7) * 48;             // 8 addresses scattered
  case 13: return half * 432 + ((lane >> 1) & 3) * width; // pairs-of-lanes pattern
  case 14: return ((0x9e3779b9u * (lane + 1)) >> 31) * 48;              // 2 addresses, random lanes
  case 15: return ((0x9e3779b9u * (lane + 1)) >> 30) * 48;              // 4 addresses, random lanes
  case 16: return ((0x9e3779b9u * (lane + 1)) >> 31) * 48 + (lane >> 4) * 96; // 2 random + cell change at half
  case 17: return (lane % 3) * 48;                                      // 3 addresses cyclic
  case 18: return ((0x9e3779b9u * (lane + 1)) >> 31) * 48 + ((0x85ebca6bu * (lane + 1)) >> 31) * 528; // hx,hy random
  case 19: return (lane & 1) * 8;                                       // 2 adjacent words alternating
  case 20: return (lane & 3) * 48;                                      // b-pattern but 48 B apart
  default: return 0;
  }
}

template <int W>
__global__ void bench(int pat, long long* out, double* sink)
{
  extern __shared__ __align__(16) unsigned char sm[];
  for (int i = threadIdx.x; i < 16384 / 8; i += blockDim.x)
    reinterpret_cast<double*>(sm)[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned char* p = sm + warp * 64 * 0 + pattern_addr(pat, lane, W);
  double acc0 = 0, acc1 = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const unsigned char* q = p + u * 1024 + (it & 1) * 8192;
      if (W == 4) {
        acc0 += __int_as_float(*reinterpret_cast<const volatile int*>(q));
      } else if (W == 8) {
        acc0 += *reinterpret_cast<const volatile double*>(q);
      } else {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(q)));
        acc0 += v.x;
        acc1 += v.y;
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0)
    out[blockIdx.x] = t1 - t0;
  if (acc0 + acc1 == 123.456)
    sink[0] = acc0;
}

int main()
{
  long long* d_out;
  double*    d_sink;
  int        nsm = 148;
  cudaMalloc(&d_out, nsm * sizeof(long long));
  cudaMalloc(&d_sink, 8);
  const int threads = 512; // 16 warps per SM
  const char* names[] = {"uniform", "half-uniform(432B)", "a-pattern", "b-pattern", "distinct-contig", "alt-2addr",
                         "scatter-4addr", "quarter-uniform", "b-pattern stride2", "pairs-contig", "b-pattern(448B)",
                         "b-pattern halves adjacent", "scatter-8addr", "pairs-pattern", "rand-2addr", "rand-4addr", "rand-2addr+cell", "cyclic-3addr", "rand hx,hy", "alt adjacent words", "b-pattern 48B"};
  for (int w : {8, 16}) {
    for (int pat = 0; pat < 21; pat++) {
      for (int rep = 0; rep < 2; rep++) {
        if (w == 4) bench<4><<<nsm, threads, 32768>>>(pat, d_out, d_sink);
        if (w == 8) bench<8><<<nsm, threads, 32768>>>(pat, d_out, d_sink);
        if (w == 16) bench<16><<<nsm, threads, 32768>>>(pat, d_out, d_sink);
      }
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      double cyc = 0;
      for (int i = 0; i < nsm; i++) cyc += h[i];
      cyc /= nsm;
      double per = cyc / ((double)ITER * UNROLL * (threads / 32));
      printf("width %2d  %-28s cycles/warp-load/SM %.2f\n", w, names[pat], per);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
