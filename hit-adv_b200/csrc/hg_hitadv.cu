// hg_hitadv.cu -- fused HiT-ADV deformation (SURVEY.md section 8f "next" #1).
//
// Replaces, per attack iteration, ShapeAttack/HiT_ADV.py:168-175 + kernel_density :298-304:
//     w[b,j,n]   = exp(-||x_n - c_j|| / (2 delta_j^2))                      (un-squared norm, as the reference)
//     out[b,:,n] = sum_j (x_n + p_j) w_jn / sum_j w_jn
// which the reference evaluates with two `repeat`ed [B,3,K,J] tensors (2 x 604 MB at B=256, K=1024, J=192), a
// J-step Python loop of elementwise kernels (~1k launches forward, as many backward) and autograd-saved copies.
// Here: one forward and one backward kernel, nothing of size B*K*J is stored (weights are recomputed).
//
// Forward keeps the reference's FP32 accumulation order over j (acc += (x + p_j) * w_j; den += w_j), so it differs
// from the reference only through exp() rounding.  Backward (x and c are constants in the attack):
//     dL/dp_jc     = sum_n g_cn w_jn / W_n
//     dL/dw_jn     = sum_c g_cn ((x_cn + p_jc) - out_cn) / W_n
//     dL/ddelta_j  = sum_n dL/dw_jn * w_jn * r_jn / delta_j^3
// reduced over n with a fixed-order tree (deterministic).
#include "hg_common.cuh"

namespace {

constexpr int kDefThreads = 128;

__device__ __forceinline__ float deform_weight(float x0, float x1, float x2, float c0, float c1, float c2,
                                               float inv2d2, float *r_out) {
  const float d0 = x0 - c0, d1 = x1 - c1, d2 = x2 - c2;
  const float r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
  if (r_out) *r_out = r;
  return expf(-r * inv2d2);
}

// one thread per (b, n); the J centres / offsets / bandwidths of cloud b staged in shared memory
__global__ void __launch_bounds__(kDefThreads) deform_fwd_kernel(const float *__restrict__ ori,
                                                                 const float *__restrict__ centers,
                                                                 const float *__restrict__ perturb,
                                                                 const float *__restrict__ delta, int K, int J,
                                                                 float *__restrict__ out, float *__restrict__ deno) {
  extern __shared__ float sh[];  // [J][8]: c0 c1 c2 p0 p1 p2 1/(2 d^2) pad
  const int b = blockIdx.y;
  for (int j = threadIdx.x; j < J; j += kDefThreads) {
    const float d = delta[(size_t)b * J + j];
    sh[j * 8 + 0] = centers[((size_t)b * 3 + 0) * J + j];
    sh[j * 8 + 1] = centers[((size_t)b * 3 + 1) * J + j];
    sh[j * 8 + 2] = centers[((size_t)b * 3 + 2) * J + j];
    sh[j * 8 + 3] = perturb[((size_t)b * J + j) * 3 + 0];
    sh[j * 8 + 4] = perturb[((size_t)b * J + j) * 3 + 1];
    sh[j * 8 + 5] = perturb[((size_t)b * J + j) * 3 + 2];
    sh[j * 8 + 6] = 1.0f / (2.0f * d * d);
    sh[j * 8 + 7] = 0.f;
  }
  __syncthreads();
  const int n = blockIdx.x * kDefThreads + threadIdx.x;
  if (n >= K) return;
  const float x0 = ori[((size_t)b * 3 + 0) * K + n], x1 = ori[((size_t)b * 3 + 1) * K + n],
              x2 = ori[((size_t)b * 3 + 2) * K + n];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, den = 0.f;
  for (int j = 0; j < J; ++j) {
    const float4 c = *reinterpret_cast<const float4 *>(sh + j * 8);
    const float4 p = *reinterpret_cast<const float4 *>(sh + j * 8 + 4);  // p.x = p2 .. see below
    // layout: c = (c0,c1,c2,p0), p = (p1,p2,inv2d2,pad)
    const float w = deform_weight(x0, x1, x2, c.x, c.y, c.z, p.z, nullptr);
    a0 = __fadd_rn(a0, __fmul_rn(__fadd_rn(x0, c.w), w));
    a1 = __fadd_rn(a1, __fmul_rn(__fadd_rn(x1, p.x), w));
    a2 = __fadd_rn(a2, __fmul_rn(__fadd_rn(x2, p.y), w));
    den = __fadd_rn(den, w);
  }
  out[((size_t)b * 3 + 0) * K + n] = __fdiv_rn(a0, den);
  out[((size_t)b * 3 + 1) * K + n] = __fdiv_rn(a1, den);
  out[((size_t)b * 3 + 2) * K + n] = __fdiv_rn(a2, den);
  deno[(size_t)b * K + n] = den;
}

// one CTA per (b, group of kJT centres); threads stride over n and share the per-point loads across the group;
// 4*kJT block reductions (dp0, dp1, dp2, ddelta per centre) with a fixed-order tree
constexpr int kJT = 8;
__global__ void __launch_bounds__(kDefThreads) deform_bwd_kernel(
    const float *__restrict__ ori, const float *__restrict__ centers, const float *__restrict__ perturb,
    const float *__restrict__ delta, const float *__restrict__ out, const float *__restrict__ deno,
    const float *__restrict__ grad_out, int K, int J, float *__restrict__ grad_perturb,
    float *__restrict__ grad_delta) {
  const int b = blockIdx.y, j0 = blockIdx.x * kJT, tid = threadIdx.x;
  float c0[kJT], c1[kJT], c2[kJT], p0[kJT], p1[kJT], p2[kJT], inv2d2[kJT], invd3[kJT];
  float s0[kJT], s1[kJT], s2[kJT], sd[kJT];
#pragma unroll
  for (int u = 0; u < kJT; ++u) {
    const int j = min(j0 + u, J - 1);
    c0[u] = centers[((size_t)b * 3 + 0) * J + j];
    c1[u] = centers[((size_t)b * 3 + 1) * J + j];
    c2[u] = centers[((size_t)b * 3 + 2) * J + j];
    p0[u] = perturb[((size_t)b * J + j) * 3 + 0];
    p1[u] = perturb[((size_t)b * J + j) * 3 + 1];
    p2[u] = perturb[((size_t)b * J + j) * 3 + 2];
    const float d = delta[(size_t)b * J + j];
    inv2d2[u] = 1.0f / (2.0f * d * d);
    invd3[u] = 1.0f / (d * d * d);
    s0[u] = s1[u] = s2[u] = sd[u] = 0.f;
  }
  for (int n = tid; n < K; n += kDefThreads) {
    const float x0 = ori[((size_t)b * 3 + 0) * K + n], x1 = ori[((size_t)b * 3 + 1) * K + n],
                x2 = ori[((size_t)b * 3 + 2) * K + n];
    const float iw = 1.0f / deno[(size_t)b * K + n];
    const float g0 = grad_out[((size_t)b * 3 + 0) * K + n], g1 = grad_out[((size_t)b * 3 + 1) * K + n],
                g2 = grad_out[((size_t)b * 3 + 2) * K + n];
    const float o0 = out[((size_t)b * 3 + 0) * K + n], o1 = out[((size_t)b * 3 + 1) * K + n],
                o2 = out[((size_t)b * 3 + 2) * K + n];
    const float gx = g0 * (x0 - o0) + g1 * (x1 - o1) + g2 * (x2 - o2);
#pragma unroll
    for (int u = 0; u < kJT; ++u) {
      float r;
      const float w = deform_weight(x0, x1, x2, c0[u], c1[u], c2[u], inv2d2[u], &r);
      const float wn = w * iw;
      s0[u] += g0 * wn;
      s1[u] += g1 * wn;
      s2[u] += g2 * wn;
      const float dw = (gx + g0 * p0[u] + g1 * p1[u] + g2 * p2[u]) * iw;  // sum_c g_c ((x_c + p_c) - out_c) / W
      sd[u] += dw * w * r * invd3[u];
    }
  }
  __shared__ float red[4 * kJT][kDefThreads + 1];
#pragma unroll
  for (int u = 0; u < kJT; ++u) {
    red[4 * u + 0][tid] = s0[u];
    red[4 * u + 1][tid] = s1[u];
    red[4 * u + 2][tid] = s2[u];
    red[4 * u + 3][tid] = sd[u];
  }
  __syncthreads();
  // 4*kJT rows of kDefThreads partials: warp w reduces rows w, w+4, ... in a fixed order
  const int lane = tid & 31, warp = tid >> 5;
  for (int row = warp; row < 4 * kJT; row += kDefThreads / 32) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < kDefThreads / 32; ++q) v += red[row][lane + 32 * q];
#pragma unroll
    for (int st = 16; st > 0; st >>= 1) v += __shfl_down_sync(0xffffffffu, v, st);
    if (lane == 0) {
      const int u = row >> 2, comp = row & 3, j = j0 + u;
      if (j < J) {
        if (comp < 3)
          grad_perturb[((size_t)b * J + j) * 3 + comp] = v;
        else
          grad_delta[(size_t)b * J + j] = v;
      }
    }
  }
}

}  // namespace

HG_API int hg_hitadv_deform_fwd_f32(const float *ori, const float *centers, const float *perturb, const float *delta,
                                    int B, int K, int J, float *out, float *deno, hgStream stream_) {
  HG_NVTX_RANGE("hg_hitadv_deform_fwd_f32");
  HG_REQUIRE(ori && centers && perturb && delta && out && deno, HG_E_BADARG, "hitadv_deform_fwd: null pointer");
  HG_REQUIRE(B > 0 && K > 0 && J > 0, HG_E_BADARG, "hitadv_deform_fwd: sizes must be positive");
  HG_REQUIRE(B <= 65535 && (size_t)J * 32 <= 160 * 1024, HG_E_UNSUPPORTED, "hitadv_deform_fwd: B or J too large");
  const size_t smem = (size_t)J * 8 * sizeof(float);
  if (smem > 48 * 1024)
    HG_CUDA(cudaFuncSetAttribute(deform_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  deform_fwd_kernel<<<dim3((K + kDefThreads - 1) / kDefThreads, B), kDefThreads, smem, hg_stream(stream_)>>>(
      ori, centers, perturb, delta, K, J, out, deno);
  HG_CHECK_LAUNCH("deform_fwd_kernel");
  return HG_OK;
}

HG_API int hg_hitadv_deform_bwd_f32(const float *ori, const float *centers, const float *perturb, const float *delta,
                                    const float *out, const float *deno, const float *grad_out, int B, int K, int J,
                                    float *grad_perturb, float *grad_delta, hgStream stream_) {
  HG_NVTX_RANGE("hg_hitadv_deform_bwd_f32");
  HG_REQUIRE(ori && centers && perturb && delta && out && deno && grad_out && grad_perturb && grad_delta, HG_E_BADARG,
             "hitadv_deform_bwd: null pointer");
  HG_REQUIRE(B > 0 && K > 0 && J > 0, HG_E_BADARG, "hitadv_deform_bwd: sizes must be positive");
  HG_REQUIRE(B <= 65535, HG_E_UNSUPPORTED, "hitadv_deform_bwd: B too large");
  deform_bwd_kernel<<<dim3((J + kJT - 1) / kJT, B), kDefThreads, 0, hg_stream(stream_)>>>(
      ori, centers, perturb, delta, out, deno, grad_out, K, J, grad_perturb, grad_delta);
  HG_CHECK_LAUNCH("deform_bwd_kernel");
  return HG_OK;
}
