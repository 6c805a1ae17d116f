// Pure issue-rate micro-benchmarks for the packed FP32 instructions on sm_100a: how many cycles does one warp
// instruction cost per scheduler, depending on its register operands?  (No memory traffic, all warps busy.)
#include <cuda_runtime.h>
#include <cstdio>

#define FFMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define FADD2(d, a, b) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define FMUL2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}

// MODE 0: FFMA2 acc = s_i(broadcast scalar, distinct per instr) * B_i(pair, distinct) + acc   [5 register reads]
// MODE 1: FFMA2 acc = s(same scalar)  * B_i + acc
// MODE 2: FFMA2 acc = s_i * B(same pair) + acc
// MODE 3: FFMA2 acc = A_i(pair) * B(same pair) + acc          [3 pairs]
// MODE 4: FADD2 acc = acc + B_i (pair + pair)
// MODE 5: FMUL2 d_i = s_i * B(same pair)   (results consumed by an FADD2 once per 16)
// MODE 6: scalar FFMA acc = a_i * b + acc (same b), 32 accumulators
// MODE 7: 5 FFMA2 (mode 2) : 2 FMNMX3 mix
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, const float *in, int iters) {
  const int tid = threadIdx.x;
  float s[16];
  unsigned long long B[16], acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    s[i] = in[tid * 64 + i];
    B[i] = pk(in[tid * 64 + 16 + 2 * i], in[tid * 64 + 17 + 2 * i]);
    acc[i] = pk(0.f, 0.f);
  }
  float m0 = in[tid * 64 + 60], m1 = in[tid * 64 + 61];
  float mm[16];
  int im[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { mm[i] = in[tid * 64 + 48 + (i & 7)] + i; im[i] = __float_as_int(mm[i]); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) FFMA2(acc[i], pk(s[i], s[i]), B[i], acc[i]);
      if (MODE == 1) FFMA2(acc[i], pk(s[0], s[0]), B[i], acc[i]);
      if (MODE == 2) FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]);
      if (MODE == 3) FFMA2(acc[i], B[i], B[0], acc[i]);
      if (MODE == 4) FADD2(acc[i], acc[i], B[i]);
      if (MODE == 5) FMUL2(acc[i], pk(s[i], s[i]), B[0]);
      if (MODE == 6) {
        float lo = __uint_as_float((unsigned)acc[i]), hi = __uint_as_float((unsigned)(acc[i] >> 32));
        lo = __fmaf_rn(s[i], m0, lo);
        hi = __fmaf_rn(s[i], m1, hi);
        acc[i] = pk(lo, hi);
      }
      if (MODE == 7) {
        FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]);
        if (i % 5 == 1 || i % 5 == 3) m0 = fminf(fminf(m0, s[i]), m1);
      }
      if (MODE == 8) mm[i] = fminf(fminf(mm[i], s[i]), s[(i + 1) & 15]);                       // FMNMX3 only
      if (MODE == 9) mm[i] = fminf(mm[i], s[i]);                                              // FMNMX only
      if (MODE == 10) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) mm[i] = fminf(fminf(mm[i], s[i]), s[(i + 1) & 15]); }
      if (MODE == 11) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) mm[i] = fminf(mm[i], s[i]); }
      if (MODE == 12) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) im[i] = min(min(im[i], __float_as_int(s[i])), __float_as_int(s[(i + 1) & 15])); }
      if (MODE == 13) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); mm[i] = fminf(fminf(mm[i], s[i]), s[(i + 1) & 15]); }
      if (MODE == 16) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) mm[i] = fminf(mm[i], __uint_as_float((unsigned)acc[i])); }
      if (MODE == 17) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); mm[i] = fminf(mm[i], __uint_as_float((unsigned)acc[i])); }
      if (MODE == 18) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) mm[i] = fminf(fminf(mm[i], __uint_as_float((unsigned)acc[i])), __uint_as_float((unsigned)(acc[i] >> 32))); }
      if (MODE == 19) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) im[i] = min(im[i], (int)(unsigned)acc[i]); }
      if (MODE == 20) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); mm[i] = fminf(mm[i], __uint_as_float((unsigned)acc[i])); mm[(i + 8) & 15] = fminf(mm[(i + 8) & 15], __uint_as_float((unsigned)(acc[i] >> 32))); }
      if (MODE == 21) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) { asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %0; selp.f32 %0, %1, %0, p;}" : "+f"(mm[i]) : "f"(s[i])); } }
      if (MODE == 22) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) { float v = __uint_as_float((unsigned)acc[i]); asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %0; selp.f32 %0, %1, %0, p;}" : "+f"(mm[i]) : "f"(v)); } }
      if (MODE == 23) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) { asm volatile("min.f32 %0, %0, %1;" : "+f"(mm[i]) : "f"(s[i])); } }
      if (MODE == 24) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(mm[i]) : "f"(s[i]), "f"(s[(i + 1) & 15])); } }
      if (MODE == 25) {  // the integer threshold filter of the kNN main loop: 2 IADD + 2 LOP3 + 1 SHF per FFMA2-pair... per instr here
        FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]);
        const int dx = (int)(unsigned)acc[(i + 5) & 15], dy = (int)(unsigned)(acc[(i + 5) & 15] >> 32);
        const int tx = dx - im[i & 3], ty = dy - im[i & 3];
        unsigned u = (unsigned)(tx | ty | dx);
        u |= (unsigned)dy;
        im[4 + (i & 3)] = __funnelshift_l(u, im[4 + (i & 3)], 1);
      }
      if (MODE == 26) {  // same with the current float filter: FMNMX + FSETP + predicated LOP3
        FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]);
        const float dx = __uint_as_float((unsigned)acc[(i + 5) & 15]), dy = __uint_as_float((unsigned)(acc[(i + 5) & 15] >> 32));
        if (fminf(dx, dy) < mm[i & 3]) im[4 + (i & 3)] |= (1 << i);
      }
      if (MODE == 14) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) { im[i] += __float_as_int(s[i]); } }   // IADD
      if (MODE == 15) { FFMA2(acc[i], pk(s[i], s[i]), B[0], acc[i]); if (i & 1) { im[i] ^= __float_as_int(s[i]) & im[(i + 1) & 15]; } }  // LOP3
    }
  }
  float r = m0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += mm[i] + __int_as_float(im[i]);
#pragma unroll
  for (int i = 0; i < 16; ++i) r += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
  out[blockIdx.x * 256 + tid] = r;
}

template <int MODE>
void run(const char *name, float *out, const float *in, int blocks_per_sm) {
  const int iters = 4096, grid = 148 * blocks_per_sm;
  k<MODE><<<grid, 256>>>(out, in, iters);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<grid, 256>>>(out, in, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double warp_instr_per_smsp = (double)iters * 16 * (blocks_per_sm * 8 / 4.0) * (MODE == 6 ? 2 : 1);
  const double cycles = ms * 1e-3 * 1.965e9;
  printf("%-58s %d warps/sched  %7.3f ms  %.2f cycles per warp-instruction per scheduler\n", name, blocks_per_sm * 2, ms,
         cycles / warp_instr_per_smsp);
}

int main() {
  float *out, *in;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaMalloc(&in, 256 * 64 * sizeof(float));
  cudaMemset(in, 0, 256 * 64 * sizeof(float));
  for (int bps : {3}) {
    run<0>("FFMA2 scalar_i * pair_i + acc (5 reg reads)", out, in, bps);
    run<1>("FFMA2 scalar(same) * pair_i + acc", out, in, bps);
    run<2>("FFMA2 scalar_i * pair(same) + acc", out, in, bps);
    run<3>("FFMA2 pair_i * pair(same) + acc", out, in, bps);
    run<4>("FADD2 acc + pair_i", out, in, bps);
    run<5>("FMUL2 scalar_i * pair(same)", out, in, bps);
    run<6>("scalar FFMA a_i * b(same) + acc (per FFMA)", out, in, bps);
    run<7>("FFMA2 (scalar_i*pair(same)) with 2 FMNMX3 per 5", out, in, bps);
    run<8>("FMNMX3 only (per FMNMX3)", out, in, bps);
    run<9>("FMNMX only (per FMNMX)", out, in, bps);
    run<10>("16 FFMA2 + 8 FMNMX3 (per FFMA2)", out, in, bps);
    run<11>("16 FFMA2 + 8 FMNMX  (per FFMA2)", out, in, bps);
    run<12>("16 FFMA2 + 8 integer min3 (per FFMA2)", out, in, bps);
    run<13>("16 FFMA2 + 16 FMNMX3 (per FFMA2)", out, in, bps);
    run<14>("16 FFMA2 + 8 IADD (per FFMA2)", out, in, bps);
    run<25>("16 x (FFMA2 + int filter: 2 IADD 2 LOP3 1 SHF) (per FFMA2)", out, in, bps);
    run<26>("16 x (FFMA2 + float filter: FMNMX FSETP @LOP3) (per FFMA2)", out, in, bps);
    run<21>("16 FFMA2 + 8 (FSETP+FSEL) static (per FFMA2)", out, in, bps);
    run<22>("16 FFMA2 + 8 (FSETP+FSEL) fresh (per FFMA2)", out, in, bps);
    run<23>("16 FFMA2 + 8 FMNMX2in static, volatile (per FFMA2)", out, in, bps);
    run<24>("16 FFMA2 + 8 FMNMX3 static, volatile (per FFMA2)", out, in, bps);
    run<16>("16 FFMA2 + 8 FMNMX2in on fresh values (per FFMA2)", out, in, bps);
    run<17>("16 FFMA2 + 16 FMNMX2in on fresh values (per FFMA2)", out, in, bps);
    run<18>("16 FFMA2 + 8 FMNMX3 on fresh values (per FFMA2)", out, in, bps);
    run<19>("16 FFMA2 + 8 IMNMX2in on fresh values (per FFMA2)", out, in, bps);
    run<20>("16 FFMA2 + 32 FMNMX2in on fresh values (per FFMA2)", out, in, bps);
    run<15>("16 FFMA2 + 8 LOP3 (per FFMA2)", out, in, bps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
