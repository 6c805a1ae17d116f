// Micro-benchmark of the nn_bidir inner loop: which part of the instruction mix limits the packed-FP32 pipe?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o nn_variants nn_variants.cu ; run on a B200.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstdio>
#include <vector>

__device__ __forceinline__ float wmin(float v) {
  float m;
  asm("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}

// VAR 0: full (col-min FMNMX3, row-min tree, CREDUX+ballot+STS)      VAR 1: no cross-lane step
// VAR 2: no row direction at all (col-min only)                       VAR 3: math only (one FMNMX3 per pair-of-pairs)
// VAR 4: like 0 but scalar FFMA/FADD (no packing)
template <int T, int VAR, int WPS>
__global__ void __launch_bounds__(128, WPS) k(const float4 *__restrict__ xs_g, const float *__restrict__ yg, float *out,
                                              uint4 *rowout, int RB, int reps) {
  constexpr int TP = T / 2;
  extern __shared__ float4 xs[];
  for (int r = threadIdx.x; r < RB + 2; r += 128) xs[r] = xs_g[(blockIdx.x * 7 + r) % 4096];
  float2 y0[TP], y1[TP], y2[TP], ry[TP];
#pragma unroll
  for (int q = 0; q < TP; ++q) {
    const float *p = yg + ((blockIdx.x * 128 + threadIdx.x) % 4096) * 64 + q * 8;
    y0[q] = make_float2(p[0], p[1]);
    y1[q] = make_float2(p[2], p[3]);
    y2[q] = make_float2(p[4], p[5]);
    ry[q] = make_float2(p[6], p[7]);
  }
  __syncthreads();
  float cm[T];
#pragma unroll
  for (int t = 0; t < T; ++t) cm[t] = CUDART_INF_F;
  const int lane = threadIdx.x & 31;
  float accm = CUDART_INF_F;
  for (int rep = 0; rep < reps; ++rep) {
    float4 xa = xs[0], xb = xs[1];
    for (int r = 0; r < RB; r += 2) {
      const float4 na = xs[r + 2], nb = xs[r + 3];
      float2 pa[TP], pb[TP];
#pragma unroll
      for (int q = 0; q < TP; ++q) {
        if (VAR == 4) {
          float ax = xa.x * y0[q].x, ay = xa.x * y0[q].y, bx = xb.x * y0[q].x, by = xb.x * y0[q].y;
          ax = __fmaf_rn(xa.y, y1[q].x, ax); ay = __fmaf_rn(xa.y, y1[q].y, ay);
          bx = __fmaf_rn(xb.y, y1[q].x, bx); by = __fmaf_rn(xb.y, y1[q].y, by);
          ax = __fmaf_rn(xa.z, y2[q].x, ax); ay = __fmaf_rn(xa.z, y2[q].y, ay);
          bx = __fmaf_rn(xb.z, y2[q].x, bx); by = __fmaf_rn(xb.z, y2[q].y, by);
          pa[q] = make_float2(__fadd_rn(__fadd_rn(xa.w, ry[q].x), ax), __fadd_rn(__fadd_rn(xa.w, ry[q].y), ay));
          pb[q] = make_float2(__fadd_rn(__fadd_rn(xb.w, ry[q].x), bx), __fadd_rn(__fadd_rn(xb.w, ry[q].y), by));
        } else if (VAR == 5 || VAR == 6) {  // hybrid: packed 2-operand ops, scalar 3-operand FFMA
          float2 ta = __fmul2_rn(make_float2(xa.x, xa.x), y0[q]);
          float2 tb = __fmul2_rn(make_float2(xb.x, xb.x), y0[q]);
          ta.x = __fmaf_rn(xa.y, y1[q].x, ta.x); ta.y = __fmaf_rn(xa.y, y1[q].y, ta.y);
          tb.x = __fmaf_rn(xb.y, y1[q].x, tb.x); tb.y = __fmaf_rn(xb.y, y1[q].y, tb.y);
          if (VAR == 5) {
            ta.x = __fmaf_rn(xa.z, y2[q].x, ta.x); ta.y = __fmaf_rn(xa.z, y2[q].y, ta.y);
            tb.x = __fmaf_rn(xb.z, y2[q].x, tb.x); tb.y = __fmaf_rn(xb.z, y2[q].y, tb.y);
          } else {
            ta = __ffma2_rn(make_float2(xa.z, xa.z), y2[q], ta);
            tb = __ffma2_rn(make_float2(xb.z, xb.z), y2[q], tb);
          }
          pa[q] = __fadd2_rn(__fadd2_rn(make_float2(xa.w, xa.w), ry[q]), ta);
          pb[q] = __fadd2_rn(__fadd2_rn(make_float2(xb.w, xb.w), ry[q]), tb);
        } else {
          float2 ta = __fmul2_rn(make_float2(xa.x, xa.x), y0[q]);
          float2 tb = __fmul2_rn(make_float2(xb.x, xb.x), y0[q]);
          ta = __ffma2_rn(make_float2(xa.y, xa.y), y1[q], ta);
          tb = __ffma2_rn(make_float2(xb.y, xb.y), y1[q], tb);
          ta = __ffma2_rn(make_float2(xa.z, xa.z), y2[q], ta);
          tb = __ffma2_rn(make_float2(xb.z, xb.z), y2[q], tb);
          pa[q] = __fadd2_rn(__fadd2_rn(make_float2(xa.w, xa.w), ry[q]), ta);
          pb[q] = __fadd2_rn(__fadd2_rn(make_float2(xb.w, xb.w), ry[q]), tb);
        }
        if (VAR != 3) {
          cm[2 * q] = fminf(fminf(cm[2 * q], pa[q].x), pb[q].x);
          cm[2 * q + 1] = fminf(fminf(cm[2 * q + 1], pa[q].y), pb[q].y);
        } else {
          cm[2 * q] = fminf(fminf(cm[2 * q], pa[q].x), pb[q].y);
          cm[2 * q + 1] = fminf(fminf(cm[2 * q + 1], pa[q].y), pb[q].x);
        }
      }
      if (VAR == 0 || VAR == 1 || VAR == 4 || VAR == 5 || VAR == 6) {
        float ma = fminf(pa[0].x, pa[0].y), mb = fminf(pb[0].x, pb[0].y);
#pragma unroll
        for (int q = 1; q < TP; ++q) {
          ma = fminf(fminf(ma, pa[q].x), pa[q].y);
          mb = fminf(fminf(mb, pb[q].x), pb[q].y);
        }
        if (VAR == 1) {
          accm = fminf(fminf(accm, ma), mb);
        } else {
          const float wa = wmin(ma), wb = wmin(mb);
          const unsigned ka = __ballot_sync(0xffffffffu, ma == wa), kb = __ballot_sync(0xffffffffu, mb == wb);
          if (lane == 0) rowout[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 64 + ((r >> 1) & 63)] = make_uint4(__float_as_uint(wa), ka, __float_as_uint(wb), kb);
        }
      }
      xa = na;
      xb = nb;
    }
  }
  float s = accm;
#pragma unroll
  for (int t = 0; t < T; ++t) s = fminf(s, cm[t]);
  out[blockIdx.x * 128 + threadIdx.x] = s;
}

// VAR 7: rows packed as pairs (a,b), columns held as scalars; full bookkeeping
template <int T, int WPS>
__global__ void __launch_bounds__(128, WPS) kx(const float4 *__restrict__ xs_g, const float *__restrict__ yg, float *out,
                                               uint4 *rowout, int RB, int reps) {
  extern __shared__ float4 xs[];  // per row pair: (xa0,xb0,xa1,xb1), (xa2,xb2,rxa,rxb)
  for (int r = threadIdx.x; r < RB + 2; r += 128) xs[r] = xs_g[(blockIdx.x * 7 + r) % 4096];
  float y0[T], y1[T], y2[T], ry[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float *p = yg + ((blockIdx.x * 128 + threadIdx.x) % 4096) * 64 + t * 4;
    y0[t] = p[0]; y1[t] = p[1]; y2[t] = p[2]; ry[t] = p[3];
  }
  __syncthreads();
  float cm[T];
#pragma unroll
  for (int t = 0; t < T; ++t) cm[t] = CUDART_INF_F;
  const int lane = threadIdx.x & 31;
  for (int rep = 0; rep < reps; ++rep) {
    float4 xA = xs[0], xB = xs[1];
    for (int r = 0; r < RB; r += 2) {
      const float4 nA = xs[r + 2], nB = xs[r + 3];
      const float2 X0 = make_float2(xA.x, xA.y), X1 = make_float2(xA.z, xA.w), X2 = make_float2(xB.x, xB.y), RX = make_float2(xB.z, xB.w);
      float2 P[T];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        float2 tt = __fmul2_rn(make_float2(y0[t], y0[t]), X0);
        tt = __ffma2_rn(make_float2(y1[t], y1[t]), X1, tt);
        tt = __ffma2_rn(make_float2(y2[t], y2[t]), X2, tt);
        P[t] = __fadd2_rn(__fadd2_rn(make_float2(ry[t], ry[t]), RX), tt);
        cm[t] = fminf(fminf(cm[t], P[t].x), P[t].y);
      }
      float ma = fminf(P[0].x, P[1].x), mb = fminf(P[0].y, P[1].y);
#pragma unroll
      for (int t = 2; t < T; t += 2) {
        ma = fminf(fminf(ma, P[t].x), P[t + 1].x);
        mb = fminf(fminf(mb, P[t].y), P[t + 1].y);
      }
      const float wa = wmin(ma), wb = wmin(mb);
      const unsigned ka = __ballot_sync(0xffffffffu, ma == wa), kb = __ballot_sync(0xffffffffu, mb == wb);
      if (lane == 0) rowout[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 64 + ((r >> 1) & 63)] = make_uint4(__float_as_uint(wa), ka, __float_as_uint(wb), kb);
      xA = nA;
      xB = nB;
    }
  }
  float s = CUDART_INF_F;
#pragma unroll
  for (int t = 0; t < T; ++t) s = fminf(s, cm[t]);
  out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int T, int WPS>
__global__ void __launch_bounds__(128, WPS) kx4(const float4 *__restrict__ xs_g, const float *__restrict__ yg, float *out,
                                                uint4 *rowout, int RB, int reps) {
  extern __shared__ float4 xs[];
  for (int r = threadIdx.x; r < RB + 4; r += 128) xs[r] = xs_g[(blockIdx.x * 7 + r) % 4096];
  float y0[T], y1[T], y2[T], ry[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float *p = yg + ((blockIdx.x * 128 + threadIdx.x) % 4096) * 64 + t * 4;
    y0[t] = p[0]; y1[t] = p[1]; y2[t] = p[2]; ry[t] = p[3];
  }
  __syncthreads();
  float cm[T];
#pragma unroll
  for (int t = 0; t < T; ++t) cm[t] = CUDART_INF_F;
  const int lane = threadIdx.x & 31;
  for (int rep = 0; rep < reps; ++rep) {
    float4 xA = xs[0], xB = xs[1], xC = xs[2], xD = xs[3];
    for (int r = 0; r < RB; r += 4) {
      const float4 nA = xs[r + 4], nB = xs[r + 5], nC = xs[r + 6], nD = xs[r + 7];
      const float2 X0 = make_float2(xA.x, xA.y), X1 = make_float2(xA.z, xA.w), X2 = make_float2(xB.x, xB.y), RX = make_float2(xB.z, xB.w);
      const float2 Z0 = make_float2(xC.x, xC.y), Z1 = make_float2(xC.z, xC.w), Z2 = make_float2(xD.x, xD.y), RZ = make_float2(xD.z, xD.w);
      float2 P[T], Q[T];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        float2 tt = __fmul2_rn(make_float2(y0[t], y0[t]), X0);
        float2 uu = __fmul2_rn(make_float2(y0[t], y0[t]), Z0);
        tt = __ffma2_rn(make_float2(y1[t], y1[t]), X1, tt);
        uu = __ffma2_rn(make_float2(y1[t], y1[t]), Z1, uu);
        tt = __ffma2_rn(make_float2(y2[t], y2[t]), X2, tt);
        uu = __ffma2_rn(make_float2(y2[t], y2[t]), Z2, uu);
        P[t] = __fadd2_rn(__fadd2_rn(make_float2(ry[t], ry[t]), RX), tt);
        Q[t] = __fadd2_rn(__fadd2_rn(make_float2(ry[t], ry[t]), RZ), uu);
        cm[t] = fminf(fminf(fminf(cm[t], P[t].x), P[t].y), fminf(Q[t].x, Q[t].y));
      }
      float ma = fminf(P[0].x, P[1].x), mb = fminf(P[0].y, P[1].y), mc = fminf(Q[0].x, Q[1].x), md = fminf(Q[0].y, Q[1].y);
#pragma unroll
      for (int t = 2; t < T; t += 2) {
        ma = fminf(fminf(ma, P[t].x), P[t + 1].x);
        mb = fminf(fminf(mb, P[t].y), P[t + 1].y);
        mc = fminf(fminf(mc, Q[t].x), Q[t + 1].x);
        md = fminf(fminf(md, Q[t].y), Q[t + 1].y);
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
  }
  float s = CUDART_INF_F;
#pragma unroll
  for (int t = 0; t < T; ++t) s = fminf(s, cm[t]);
  out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int T, int WPS>
void runx4(const char *name, const float4 *xs, const float *y, float *out, uint4 *rowout) {
  const int RB = 256, reps = 8, grid = 148 * 16;
  const size_t smem = (RB + 4) * sizeof(float4);
  kx4<T, WPS><<<grid, 128, smem>>>(xs, y, out, rowout, RB, reps);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) kx4<T, WPS><<<grid, 128, smem>>>(xs, y, out, rowout, RB, reps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5;
  const double pairs = (double)grid * 128 * T * RB * reps;
  printf("%-44s T=%2d  %8.3f ms  %6.2fe12 pairs/s  (8 FLOP/pair: %5.1f%% of 74.45 TF)  err=%s\n", name, T, ms,
         pairs / ms / 1e9, pairs * 8 / ms / 1e9 / 74.45 * 100, cudaGetErrorString(cudaGetLastError()));
}

// rows packed, STAGE-MAJOR source order (all FMUL2, then all FFMA2 #1, ...): consecutive packed instructions share
// the row-pair operand, which the operand-reuse cache can hold -- if ptxas keeps the order
template <int T, int WPS, int ROWS4>
__global__ void __launch_bounds__(128, WPS) ks(const float4 *__restrict__ xs_g, const float *__restrict__ yg, float *out,
                                               uint4 *rowout, int RB, int reps) {
  extern __shared__ float4 xs[];
  for (int r = threadIdx.x; r < RB + 4; r += 128) xs[r] = xs_g[(blockIdx.x * 7 + r) % 4096];
  float y0[T], y1[T], y2[T], ry[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float *p = yg + ((blockIdx.x * 128 + threadIdx.x) % 4096) * 64 + t * 4;
    y0[t] = p[0]; y1[t] = p[1]; y2[t] = p[2]; ry[t] = p[3];
  }
  __syncthreads();
  float cm[T];
#pragma unroll
  for (int t = 0; t < T; ++t) cm[t] = CUDART_INF_F;
  const int lane = threadIdx.x & 31;
  for (int rep = 0; rep < reps; ++rep) {
    float4 xA = xs[0], xB = xs[1];
    for (int r = 0; r < RB; r += 2) {
      const float4 nA = xs[r + 2], nB = xs[r + 3];
      const float2 X0 = make_float2(xA.x, xA.y), X1 = make_float2(xA.z, xA.w), X2 = make_float2(xB.x, xB.y), RX = make_float2(xB.z, xB.w);
      float2 P[T], S[T];
#pragma unroll
      for (int t = 0; t < T; ++t) P[t] = __fmul2_rn(make_float2(y0[t], y0[t]), X0);
#pragma unroll
      for (int t = 0; t < T; ++t) P[t] = __ffma2_rn(make_float2(y1[t], y1[t]), X1, P[t]);
#pragma unroll
      for (int t = 0; t < T; ++t) P[t] = __ffma2_rn(make_float2(y2[t], y2[t]), X2, P[t]);
#pragma unroll
      for (int t = 0; t < T; ++t) S[t] = __fadd2_rn(RX, make_float2(ry[t], ry[t]));
#pragma unroll
      for (int t = 0; t < T; ++t) P[t] = __fadd2_rn(S[t], P[t]);
#pragma unroll
      for (int t = 0; t < T; ++t) cm[t] = fminf(fminf(cm[t], P[t].x), P[t].y);
      float ma = fminf(P[0].x, P[1].x), mb = fminf(P[0].y, P[1].y);
#pragma unroll
      for (int t = 2; t < T; t += 2) {
        ma = fminf(fminf(ma, P[t].x), P[t + 1].x);
        mb = fminf(fminf(mb, P[t].y), P[t + 1].y);
      }
      const float wa = wmin(ma), wb = wmin(mb);
      const unsigned ka = __ballot_sync(0xffffffffu, ma == wa), kb = __ballot_sync(0xffffffffu, mb == wb);
      if (lane == 0) rowout[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 64 + ((r >> 1) & 63)] = make_uint4(__float_as_uint(wa), ka, __float_as_uint(wb), kb);
      xA = nA;
      xB = nB;
    }
  }
  float s = CUDART_INF_F;
#pragma unroll
  for (int t = 0; t < T; ++t) s = fminf(s, cm[t]);
  out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int T, int WPS>
void runs(const char *name, const float4 *xs, const float *y, float *out, uint4 *rowout) {
  const int RB = 256, reps = 8, grid = 148 * 16;
  const size_t smem = (RB + 4) * sizeof(float4);
  ks<T, WPS, 0><<<grid, 128, smem>>>(xs, y, out, rowout, RB, reps);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) ks<T, WPS, 0><<<grid, 128, smem>>>(xs, y, out, rowout, RB, reps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5;
  const double pairs = (double)grid * 128 * T * RB * reps;
  printf("%-44s T=%2d  %8.3f ms  %6.2fe12 pairs/s  (8 FLOP/pair: %5.1f%% of 74.45 TF)  err=%s\n", name, T, ms,
         pairs / ms / 1e9, pairs * 8 / ms / 1e9 / 74.45 * 100, cudaGetErrorString(cudaGetLastError()));
}

template <int T, int WPS>
void runx(const char *name, const float4 *xs, const float *y, float *out, uint4 *rowout) {
  const int RB = 256, reps = 8, grid = 148 * 16;
  const size_t smem = (RB + 2) * sizeof(float4);
  kx<T, WPS><<<grid, 128, smem>>>(xs, y, out, rowout, RB, reps);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) kx<T, WPS><<<grid, 128, smem>>>(xs, y, out, rowout, RB, reps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5;
  const double pairs = (double)grid * 128 * T * RB * reps;
  printf("%-44s T=%2d  %8.3f ms  %6.2fe12 pairs/s  (8 FLOP/pair: %5.1f%% of 74.45 TF)  err=%s\n", name, T, ms,
         pairs / ms / 1e9, pairs * 8 / ms / 1e9 / 74.45 * 100, cudaGetErrorString(cudaGetLastError()));
}

template <int T, int VAR, int WPS>
void run(const char *name, const float4 *xs, const float *y, float *out, uint4 *rowout) {
  const int RB = 256, reps = 8, grid = 148 * 16;
  const size_t smem = (RB + 2) * sizeof(float4);
  k<T, VAR, WPS><<<grid, 128, smem>>>(xs, y, out, rowout, RB, reps);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) k<T, VAR, WPS><<<grid, 128, smem>>>(xs, y, out, rowout, RB, reps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5;
  const double pairs = (double)grid * 128 * T * RB * reps;
  printf("%-44s T=%2d  %8.3f ms  %6.2fe12 pairs/s  (8 FLOP/pair: %5.1f%% of 74.45 TF)  err=%s\n", name, T, ms,
         pairs / ms / 1e9, pairs * 8 / ms / 1e9 / 74.45 * 100, cudaGetErrorString(cudaGetLastError()));
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
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) % 1000) / 1000.0f - 0.5f;
  cudaMemcpy(xs, h.data(), 4096 * sizeof(float4), cudaMemcpyHostToDevice);
  cudaMemcpy(y, h.data(), 4096 * 64 * sizeof(float), cudaMemcpyHostToDevice);
  run<16, 0, 3>("full (CREDUX+ballot+STS) 3 warps/sched", xs, y, out, rowout);
  run<16, 1, 3>("no cross-lane step", xs, y, out, rowout);
  run<16, 2, 3>("col-min only", xs, y, out, rowout);
  run<16, 3, 3>("math + col-min (same op count as 2)", xs, y, out, rowout);
  run<16, 4, 3>("full, scalar FFMA/FADD (no packing)", xs, y, out, rowout);
  run<16, 5, 3>("full, hybrid: FMUL2/FADD2 packed, FFMA scalar", xs, y, out, rowout);
  run<16, 6, 3>("full, hybrid: one FFMA pair scalar, one FFMA2", xs, y, out, rowout);
  run<8, 5, 5>("full hybrid, T=8", xs, y, out, rowout);
  run<8, 6, 5>("full hybrid-6, T=8", xs, y, out, rowout);
  runx<16, 3>("full, rows packed (a,b), y scalar, T=16", xs, y, out, rowout);
  runx<16, 4>("full, rows packed, T=16, 4 warps/sched", xs, y, out, rowout);
  runs<16, 3>("full, rows packed, STAGE-MAJOR, T=16 3 w/s", xs, y, out, rowout);
  runs<16, 4>("full, rows packed, STAGE-MAJOR, T=16 4 w/s", xs, y, out, rowout);
  runs<8, 5>("full, rows packed, STAGE-MAJOR, T=8 5 w/s", xs, y, out, rowout);
  runx<8, 5>("full, rows packed, T=8", xs, y, out, rowout);
  runx4<16, 2>("full, rows packed x4 rows, T=16, 2 w/s", xs, y, out, rowout);
  runx4<16, 3>("full, rows packed x4 rows, T=16, 3 w/s", xs, y, out, rowout);
  runx4<8, 4>("full, rows packed x4 rows, T=8, 4 w/s", xs, y, out, rowout);
  runx4<12, 3>("full, rows packed x4 rows, T=12, 3 w/s", xs, y, out, rowout);
  runx<24, 2>("full, rows packed, T=24, 2 warps/sched", xs, y, out, rowout);
  run<8, 0, 5>("full, T=8, 5 warps/sched", xs, y, out, rowout);
  run<8, 1, 5>("no cross-lane, T=8", xs, y, out, rowout);
  run<8, 2, 5>("col-min only, T=8", xs, y, out, rowout);
  run<8, 4, 5>("full scalar, T=8", xs, y, out, rowout);
  run<8, 0, 4>("full, T=8, 4 warps/sched", xs, y, out, rowout);
  run<16, 0, 2>("full, T=16, 2 warps/sched (255 regs)", xs, y, out, rowout);
  return 0;
}
