// hg_common.cuh -- shared helpers of libhitgeom (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hitgeom.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libhitgeom is written for sm_100a (B200) only"
#endif

#define HG_API extern "C" __attribute__((visibility("default")))

// ---- error reporting (thread-local message, never exit()) -------------------------------------------------
void hg_set_error(const char *fmt, ...);

#define HG_REQUIRE(cond, code, ...)   \
  do {                                \
    if (!(cond)) {                    \
      hg_set_error(__VA_ARGS__);      \
      return (code);                  \
    }                                 \
  } while (0)

// NVTX range around every compute entry point of the C ABI (header-only NVTX v3: a no-op unless a tool is attached),
// so that nsys / ncu timelines of a reference run show which reference call a kernel belongs to
struct HgNvtxRange {
  explicit HgNvtxRange(const char *name) { nvtxRangePushA(name); }
  ~HgNvtxRange() { nvtxRangePop(); }
  HgNvtxRange(const HgNvtxRange &) = delete;
  HgNvtxRange &operator=(const HgNvtxRange &) = delete;
};
#define HG_NVTX_RANGE(name) HgNvtxRange hg_nvtx_range_(name)

extern int g_hg_tune_knn_win;     // hg_tune("knn_win", n)
extern int g_hg_tune_knn_tc_off;  // hg_tune("knn_tc", 1) switches the tensor-core kNN prefilter off
extern int g_hg_tune_nn_exact;  // hg_tune("nn_exact", v), see hg_nn_bidir.cu
extern int g_hg_tune_small_fused_off;  // hg_tune("small_fused", 1): general paths for small clouds too (A/B, tests)
extern int g_hg_tune_fps_threads;  // hg_tune("fps_threads", 128|256): development knob
extern int g_hg_tune_scatter;  // hg_tune("scatter", v): development knob, see hg_abi.cu
extern unsigned long long g_hg_launches;  // kernels launched by this library (bench.py's gpu_launches)

#define HG_CHECK_LAUNCH(name)                                            \
  do {                                                                   \
    ++g_hg_launches;                                                     \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) {                                            \
      hg_set_error("%s: %s", (name), cudaGetErrorString(e__));           \
      return (int)e__;                                                   \
    }                                                                    \
  } while (0)

#define HG_CUDA(call)                                                    \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) {                                            \
      hg_set_error("%s: %s", #call, cudaGetErrorString(e__));            \
      return (int)e__;                                                   \
    }                                                                    \
  } while (0)

static inline cudaStream_t hg_stream(hgStream s) { return reinterpret_cast<cudaStream_t>(s); }

static inline size_t hg_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

int hg_sm_count();
// true the first time it is called on the current device with this flag array (function attributes such as the
// dynamic shared-memory limit are per device: one `static` flag per process would miss the second GPU)
struct HgPerDeviceOnce {
  unsigned char done[64] = {0};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = 1;
    return true;
  }
};

// per-kernel device timing hooks (hg_abi.cu); tags are part of the ABI (include/hitgeom.h HG_PROF_*)
bool hg_prof_begin(int tag, cudaStream_t s);
void hg_prof_end(int tag, cudaStream_t s, bool began);

// ---- exact FP32 building blocks (SURVEY.md section 8 a-bis) ------------------------------------------------
// Everything that decides an index goes through these intrinsics, which nvcc never contracts or reorders.

// torch.bmm with inner dim 3: fma(a2,b2, fma(a1,b1, a0*b0))
__device__ __forceinline__ float hg_dot3_fma(float a0, float a1, float a2, float b0, float b1, float b2) {
  return __fmaf_rn(a2, b2, __fmaf_rn(a1, b1, __fmul_rn(a0, b0)));
}
// torch.sum(x**2, dim) for three channels: (a0*a0 + a1*a1) + a2*a2, no FMA
__device__ __forceinline__ float hg_sumsq3_seq(float a0, float a1, float a2) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2));
}
// nvcc -O3 contraction of (ax-bx)*(ax-bx) + (ay-by)*(ay-by) + (az-bz)*(az-bz) in the pointnet2_ops kernels:
// fma(dz,dz, fma(dx,dx, round(dy*dy))) -- in `a*a + b*b` the LEFT product is fused, the right one rounded (read
// off the SASS of query_ball_point_kernel / three_nn_kernel / furthest_point_sampling_kernel built for sm_100).
__device__ __forceinline__ float hg_dist3_fma(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
// pytorch3d knn_points: `dist += diff*diff` per coordinate -> fma(dz,dz, fma(dy,dy, dx*dx))
__device__ __forceinline__ float hg_dist3_seq(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// order-preserving float <-> uint32 (so that integer min/max atomics select the float min/max exactly)
__device__ __forceinline__ unsigned hg_ord(float f) {
  const unsigned u = __float_as_uint(f);
  return u ^ (((unsigned)((int)u >> 31)) | 0x80000000u);
}
__device__ __forceinline__ float hg_unord(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}

__device__ __forceinline__ float hg_warp_min_f32(float v) {
  float m;
  asm("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));  // CREDUX.MIN.F32 (sm_100a)
  return m;
}
__device__ __forceinline__ float hg_warp_max_f32(float v) {
  float m;
  asm("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}

// ---- bulk asynchronous copies (TMA unit, 1-D): global -> shared, completion on an mbarrier ------------------------
// cp.async.bulk (SASS: UBLKCP) moves a whole contiguous block with ONE instruction issued by one thread; the copy
// engine keeps it in flight while the CTA computes on the previous block.  Sizes and both addresses must be multiples
// of 16 bytes.  hg_mbar_wait spins on the phase parity (try_wait suspends the thread in hardware between polls).
__device__ __forceinline__ unsigned hg_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hg_mbar_init(uint64_t *bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hg_smem_addr(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void hg_mbar_init_fence() {  // make the initialised barriers visible to the async proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void hg_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hg_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hg_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   hg_smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(hg_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void hg_mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "HG_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra HG_DONE_%=;\n\t"
      "bra HG_WAIT_%=;\n\t"
      "HG_DONE_%=:\n\t"
      "}" ::"r"(hg_smem_addr(bar)),
      "r"(parity)
      : "memory");
}

// __match_any_sync on keys of `nbits` bits, built from ballots: MATCH.ANY resolves one distinct value at a time (about
// 600 cycles for 32 distinct keys, measured through the CSR builder: 48 us for 16k edges per cloud), a ballot per key
// bit costs a few cycles each
__device__ __forceinline__ unsigned hg_match_any_bits(int key, int nbits) {
  unsigned same = 0xffffffffu;
  for (int b = 0; b < nbits; ++b) {
    const bool bit = (key >> b) & 1;
    const unsigned v = __ballot_sync(0xffffffffu, bit);
    same &= bit ? v : ~v;
  }
  return same;
}

// ---- deterministic reverse map (CSR) used by every scatter-style backward ----------------------------------
// keys [B,E] (destination of edge e, or <0 to skip) -> off [B,N+1], list [B,E] with, for each destination,
// its edges in ascending e.  Summing in list order is what replaces the reference's float atomicAdd.
size_t hg_csr_workspace_bytes(int B, int N, int E);         // for hg_csr_build_unordered
size_t hg_csr_stable_workspace_bytes(int B, int N, int E);  // for hg_csr_build
struct HgCsr {
  int *off;   // [B, N+1]
  int *list;  // [B, E]
};
int hg_csr_build(const int *keys, int B, int E, int N, void *workspace, size_t workspace_bytes, HgCsr *out,
                 cudaStream_t stream);
// Same layout, built with parallel integer atomics: entries of a segment are in arbitrary order, walk them with
// hg_csr_next() to get ascending edge order (deterministic sums without a stable sort).
int hg_csr_build_unordered(const int *keys, int B, int E, int N, void *workspace, size_t workspace_bytes, HgCsr *out,
                           cudaStream_t stream);
// smallest entry of list[p0..p1) that is greater than `last` (INT_MAX if none)
__device__ __forceinline__ int hg_csr_next(const int *__restrict__ list, int p0, int p1, int last) {
  int best = 0x7fffffff;
  for (int q = p0; q < p1; ++q) {
    const int e = list[q];
    if (e > last && e < best) best = e;
  }
  return best;
}

// ---- tensor-core kNN prefilter for feature clouds (hg_knn_tc.cu) ---------------------------------------------------
bool hg_knn_tc_supported(int K, int C, int k1);
int hg_knn_tc_run(const float *pc, const float *xx, int B, int K, int C, int k1, float *vals, int *idx, cudaStream_t stream);

// ---- 3-D streaming kNN (hg_knn3.cu) ---------------------------------------------------------------------------
#define HG_KNN_FORM_EXPANDED 0  // dist = (xx_j + (-2 zz)) + xx_i   (KNNDist / DGCNN)
#define HG_KNN_FORM_DIRECT 1    // dist = fma(dz,dz, fma(dy,dy, dx*dx))   (pytorch3d knn_points)
int hg_knn3_launch_i32(int form, const float *q, const float *r, int B, int Nq, int Nr, int k1, float *vals, int *idx,
                       cudaStream_t stream);
int hg_knn3_launch_i64(int form, const float *q, const float *r, int B, int Nq, int Nr, int k1, float *vals,
                       long long *idx, cudaStream_t stream);
size_t hg_knn3_seed_workspace_bytes(int B, int N);
int hg_knn3_self_seeded_i32(const float *pc, int B, int N, int k1, float *vals, int *idx, void *workspace,
                            size_t workspace_bytes, cudaStream_t stream, int *idx_state = nullptr, int state_valid = 0);
