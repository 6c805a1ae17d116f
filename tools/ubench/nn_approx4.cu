// Experiment for VERDICT r1 item 4: how much would the nn_bidir inner loop gain from
//   (a) a 4-operation APPROXIMATE distance (fma,fma,fma with rx as the initial addend, then + ry: one packed operation
//       less than the reference's (rx + ry) + zz' order, values off by a few ulp) with exact recovery in the finish kernel,
//   (b) integer 3-input min on the value bits (VIMNMX3) instead of FMNMX3,
//   (c) the runner-up bookkeeping an exact recovery needs (per-batch minima, second-best batch, 2-eps test)?
// Same loop shape as the shipped kernel (rows packed in pairs, four rows per iteration, T = 16 columns per lane as
// scalars, column-direction running minima, row-direction min tree + CREDUX + ballot).  Timing only: no results checked.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o nn_approx4 nn_approx4.cu ; run on a B200.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstdio>
#include <vector>

__device__ __forceinline__ float wmin(float v) {
  float m;
  asm("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}
__device__ __forceinline__ float imin3(float a, float b, float c) {  // 3-input signed-integer min on the value bits
  return __int_as_float(__vimin3_s32(__float_as_int(a), __float_as_int(b), __float_as_int(c)));
}

// OPS: 5 = reference order, 4 = approximate.  IMIN: integer mins.  TRACK: 0 none, 1 = shipped batch tracking (which
// batch of 32 rows last lowered a column minimum), 2 = batch minima + runner-up (what an exact recovery of the
// approximate tracker needs).
template <int T, int WPS, int OPS, bool IMIN, int TRACK>
__global__ void __launch_bounds__(128, WPS) kx4(const float4 *__restrict__ xs_g, const float *__restrict__ yg, float *out,
                                                uint4 *rowout, int RB, int reps, float eps2) {
  extern __shared__ float4 xs[];
  float *sprev = reinterpret_cast<float *>(xs + RB + 4);  // [3][T][128]
  for (int r = threadIdx.x; r < RB + 4; r += 128) xs[r] = xs_g[(blockIdx.x * 7 + r) % 4096];
  float y0[T], y1[T], y2[T], ry[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float *p = yg + ((blockIdx.x * 128 + threadIdx.x) % 4096) * 64 + t * 4;
    y0[t] = p[0]; y1[t] = p[1]; y2[t] = p[2]; ry[t] = p[3];
  }
  float cm[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    cm[t] = CUDART_INF_F;
    sprev[t * 128 + threadIdx.x] = CUDART_INF_F;
    sprev[(T + t) * 128 + threadIdx.x] = 0.f;
    sprev[(2 * T + t) * 128 + threadIdx.x] = CUDART_INF_F;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int rep = 0; rep < reps; ++rep) {
    float4 xA = xs[0], xB = xs[1], xC = xs[2], xD = xs[3];
    for (int rb = 0; rb < RB; rb += 32) {
#pragma unroll 2
      for (int rr = 0; rr < 32; rr += 4) {
        const int r = rb + rr;
        const float4 nA = xs[r + 4], nB = xs[r + 5], nC = xs[r + 6], nD = xs[r + 7];
        const float2 X0 = make_float2(xA.x, xA.y), X1 = make_float2(xA.z, xA.w), X2 = make_float2(xB.x, xB.y), RX = make_float2(xB.z, xB.w);
        const float2 Z0 = make_float2(xC.x, xC.y), Z1 = make_float2(xC.z, xC.w), Z2 = make_float2(xD.x, xD.y), RZ = make_float2(xD.z, xD.w);
        float2 P[T], Q[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
          if (OPS == 5) {
            float2 tt = __fmul2_rn(make_float2(y0[t], y0[t]), X0);
            float2 uu = __fmul2_rn(make_float2(y0[t], y0[t]), Z0);
            tt = __ffma2_rn(make_float2(y1[t], y1[t]), X1, tt);
            uu = __ffma2_rn(make_float2(y1[t], y1[t]), Z1, uu);
            tt = __ffma2_rn(make_float2(y2[t], y2[t]), X2, tt);
            uu = __ffma2_rn(make_float2(y2[t], y2[t]), Z2, uu);
            P[t] = __fadd2_rn(__fadd2_rn(make_float2(ry[t], ry[t]), RX), tt);
            Q[t] = __fadd2_rn(__fadd2_rn(make_float2(ry[t], ry[t]), RZ), uu);
          } else {
            float2 tt = __ffma2_rn(make_float2(y0[t], y0[t]), X0, RX);
            float2 uu = __ffma2_rn(make_float2(y0[t], y0[t]), Z0, RZ);
            tt = __ffma2_rn(make_float2(y1[t], y1[t]), X1, tt);
            uu = __ffma2_rn(make_float2(y1[t], y1[t]), Z1, uu);
            tt = __ffma2_rn(make_float2(y2[t], y2[t]), X2, tt);
            uu = __ffma2_rn(make_float2(y2[t], y2[t]), Z2, uu);
            P[t] = __fadd2_rn(tt, make_float2(ry[t], ry[t]));
            Q[t] = __fadd2_rn(uu, make_float2(ry[t], ry[t]));
          }
          if (IMIN)
            cm[t] = imin3(imin3(cm[t], P[t].x, P[t].y), Q[t].x, Q[t].y);
          else
            cm[t] = fminf(fminf(fminf(fminf(cm[t], P[t].x), P[t].y), Q[t].x), Q[t].y);
        }
        float ma, mb, mc, md;
        if (IMIN) {
          ma = imin3(P[0].x, P[1].x, P[2].x); mb = imin3(P[0].y, P[1].y, P[2].y);
          mc = imin3(Q[0].x, Q[1].x, Q[2].x); md = imin3(Q[0].y, Q[1].y, Q[2].y);
#pragma unroll
          for (int t = 3; t + 1 < T; t += 2) {
            ma = imin3(ma, P[t].x, P[t + 1].x); mb = imin3(mb, P[t].y, P[t + 1].y);
            mc = imin3(mc, Q[t].x, Q[t + 1].x); md = imin3(md, Q[t].y, Q[t + 1].y);
          }
          ma = imin3(ma, P[T - 1].x, ma); mb = imin3(mb, P[T - 1].y, mb);
          mc = imin3(mc, Q[T - 1].x, mc); md = imin3(md, Q[T - 1].y, md);
        } else {
          ma = fminf(P[0].x, P[1].x); mb = fminf(P[0].y, P[1].y); mc = fminf(Q[0].x, Q[1].x); md = fminf(Q[0].y, Q[1].y);
#pragma unroll
          for (int t = 2; t < T; t += 2) {
            ma = fminf(fminf(ma, P[t].x), P[t + 1].x); mb = fminf(fminf(mb, P[t].y), P[t + 1].y);
            mc = fminf(fminf(mc, Q[t].x), Q[t + 1].x); md = fminf(fminf(md, Q[t].y), Q[t + 1].y);
          }
        }
        const float wa = wmin(ma), wb = wmin(mb), wc = wmin(mc), wd = wmin(md);
        const unsigned ka = __ballot_sync(0xffffffffu, ma == wa), kb = __ballot_sync(0xffffffffu, mb == wb);
        const unsigned kc = __ballot_sync(0xffffffffu, mc == wc), kd = __ballot_sync(0xffffffffu, md == wd);
        if (lane == 0) {
          uint4 *o = rowout + (blockIdx.x * 4 + (threadIdx.x >> 5)) * 64 + ((r >> 1) & 62);
          o[0] = make_uint4(__float_as_uint(wa), ka, __float_as_uint(wb), kb);
          o[1] = make_uint4(__float_as_uint(wc), kc, __float_as_uint(wd), kd);
        }
        xA = nA; xB = nB; xC = nC; xD = nD;
      }
      if (TRACK == 1) {  // shipped: which batch last lowered the running minimum
#pragma unroll
        for (int t = 0; t < T; ++t) {
          if (cm[t] < sprev[t * 128 + threadIdx.x]) sprev[(T + t) * 128 + threadIdx.x] = (float)rb;
          sprev[t * 128 + threadIdx.x] = cm[t];
        }
      } else if (TRACK == 2) {  // cm[] holds the BATCH minimum: best / best batch / runner-up among the other batches
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float best = sprev[t * 128 + threadIdx.x], second = sprev[(2 * T + t) * 128 + threadIdx.x], bm = cm[t];
          const bool win = bm < best;
          sprev[(2 * T + t) * 128 + threadIdx.x] = fminf(second, win ? best : bm);
          if (win) {
            sprev[t * 128 + threadIdx.x] = bm;
            sprev[(T + t) * 128 + threadIdx.x] = (float)rb;
          }
          cm[t] = CUDART_INF_F;
        }
      }
    }
  }
  float s = eps2;
#pragma unroll
  for (int t = 0; t < T; ++t) s = fminf(s, fminf(cm[t], sprev[t * 128 + threadIdx.x] + sprev[(2 * T + t) * 128 + threadIdx.x]));
  out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int T, int WPS, int OPS, bool IMIN, int TRACK>
void run(const char *name, const float4 *xs, const float *y, float *out, uint4 *rowout) {
  const int RB = 256, reps = 8, grid = 148 * 16;
  const size_t smem = (RB + 4) * sizeof(float4) + 3 * T * 128 * sizeof(float);
  cudaFuncSetAttribute(kx4<T, WPS, OPS, IMIN, TRACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kx4<T, WPS, OPS, IMIN, TRACK><<<grid, 128, smem>>>(xs, y, out, rowout, RB, reps, 1e-6f);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) kx4<T, WPS, OPS, IMIN, TRACK><<<grid, 128, smem>>>(xs, y, out, rowout, RB, reps, 1e-6f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5;
  const double pairs = (double)grid * 128 * T * RB * reps;
  int clk = 1965000;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * clk * 1e3 * 148 * 4 / (pairs / 32);  // SMSP cycles per warp-level pair
  printf("%-58s %8.3f ms  %6.2fe12 pairs/s  %5.1f%% of 74.45 TF  %5.2f cyc/pair  err=%s\n", name, ms, pairs / ms / 1e9,
         pairs * 8 / ms / 1e9 / 74.45 * 100, cyc, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float4 *xs;
  float *y, *out;
  uint4 *rowout;
  cudaMalloc(&xs, 4096 * sizeof(float4));
  cudaMalloc(&y, 4096 * 8 * 8 * sizeof(float));
  cudaMalloc(&out, 148 * 16 * 128 * sizeof(float));
  cudaMalloc(&rowout, 148 * 16 * 4 * 64 * sizeof(uint4));
  std::vector<float> h(4096 * 64);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) % 1000) / 1000.0f + 0.5f;
  cudaMemcpy(xs, h.data(), 4096 * sizeof(float4), cudaMemcpyHostToDevice);
  cudaMemcpy(y, h.data(), 4096 * 64 * sizeof(float), cudaMemcpyHostToDevice);
  run<16, 3, 5, false, 1>("5-op exact, FMNMX3, shipped batch tracking (baseline)", xs, y, out, rowout);
  run<16, 3, 5, true, 1>("5-op exact, VIMNMX3 on the value bits", xs, y, out, rowout);
  run<16, 3, 4, false, 1>("4-op approximate, FMNMX3, shipped tracking", xs, y, out, rowout);
  run<16, 3, 4, true, 1>("4-op approximate, VIMNMX3", xs, y, out, rowout);
  run<16, 3, 4, false, 2>("4-op approximate, FMNMX3, runner-up tracking (exact recovery)", xs, y, out, rowout);
  run<16, 3, 4, true, 2>("4-op approximate, VIMNMX3, runner-up tracking", xs, y, out, rowout);
  run<16, 3, 5, false, 0>("5-op exact, FMNMX3, no batch tracking", xs, y, out, rowout);
  run<16, 3, 4, false, 0>("4-op approximate, FMNMX3, no batch tracking", xs, y, out, rowout);
  run<16, 4, 4, false, 2>("4-op approximate, runner-up tracking, 4 warps/sched", xs, y, out, rowout);
  run<8, 4, 4, false, 2>("4-op approximate, runner-up tracking, T=8", xs, y, out, rowout);
  return 0;
}
