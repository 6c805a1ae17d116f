// Dependent-chain latency of the warp-level primitives a furthest-point-sampling round is built from (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o redux_latency redux_latency.cu && ./redux_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(unsigned *out, long long *cyc, int iters) {
  unsigned x = threadIdx.x * 2654435761u, lane = threadIdx.x & 31;
  __shared__ unsigned long long sm[64];
  sm[threadIdx.x & 63] = x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) x = __reduce_max_sync(0xffffffffu, x ^ lane) + 1u;                    // CREDUX + back to a vector reg
    if (OP == 1) x = __shfl_xor_sync(0xffffffffu, x, 16) + lane;                         // SHFL
    if (OP == 2) x = __popc(__ballot_sync(0xffffffffu, (x + lane) & 1)) + x;             // VOTE
    if (OP == 3) { __syncthreads(); x += 1; }                                             // BAR.SYNC (4 warps)
    if (OP == 4) x = (unsigned)sm[(x + lane) & 63] + 1u;                                  // LDS.64 dependent
    if (OP == 5) {                                                                         // two CREDUX + select (one level)
      const unsigned hi = __reduce_max_sync(0xffffffffu, x ^ lane);
      const unsigned lo = __reduce_max_sync(0xffffffffu, (x ^ lane) == hi ? lane : 0u);
      x = hi + lo;
    }
    if (OP == 6) {                                                                         // 64-bit butterfly max
      unsigned long long k = ((unsigned long long)(x ^ lane) << 32) | lane;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, s);
        k = o > k ? o : k;
      }
      x = (unsigned)(k >> 32) + (unsigned)k;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  unsigned *out;
  long long *cyc, h;
  cudaMalloc(&out, 4096);
  cudaMalloc(&cyc, 8);
  const char *names[] = {"CREDUX.MAX.U32 (+IMAD.U32 back)", "SHFL.BFLY", "VOTE+POPC", "BAR.SYNC 4 warps", "LDS.64 dependent",
                         "2x CREDUX + select (one key level)", "64-bit butterfly max (5 x 2 SHFL)"};
  const int iters = 4096;
#define RUN(OP)                                              \
  chain<OP><<<1, 128>>>(out, cyc, iters);                    \
  chain<OP><<<1, 128>>>(out, cyc, iters);                    \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);            \
  printf("%-40s %7.1f cycles per dependent op\n", names[OP], (double)h / iters);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
  return 0;
}
