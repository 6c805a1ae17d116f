// hg_knn_tc.cu -- tensor-core prefilter for the k-nearest-neighbour search on FEATURE clouds (DGCNN edge-conv layers
// 2-4: C = 64, 64, 128 channels, model/dgcnn_cls.py:7-13), followed by an exact FP32 re-evaluation.
//
// The distance matrix of a C-channel cloud is a real contraction (2*C FLOP per pair), the one place on this path where
// the 5th-generation tensor cores are admissible -- provided every index stays the reference's.
// Tensor cores multiply in TF32, so their distances only SELECT candidates; every value and index that is returned
// comes from the reference's FP32 arithmetic:
//
//   1. knn_tc_filter_kernel (tcgen05 + TMEM + TMA).  One CTA per (cloud, 128 query rows).  The 128 x C query block and
//      the 256 x C candidate tiles (both K-major slices of the same [K, C] matrix) are brought in by TMA
//      (cp.async.bulk.tensor, 128-byte swizzle) and multiplied with tcgen05.mma.kind::tf32 into a 128 x 256 FP32
//      accumulator in tensor memory; the four warps read it back with tcgen05.ld (thread = row).
//      Pass 1 keeps, per row, the minima of 32 interleaved column classes; the k-th smallest of them is an upper bound
//      tau on the row's k-th smallest APPROXIMATE distance.  Pass 2 recomputes the tiles (the MMAs are cheap) and lists
//      the columns with approximate distance <= tau + 2 eps, where eps bounds |approximate - reference| for that row:
//      TF32 keeps 11 significant bits of each operand, so |zz_tf32 - zz| <= |x_i||x_j| 2^-9 and
//      eps_i = 1.1 * 2^-8 |x_i| max|x| + (FP32 rounding of both sides).  Every column whose REFERENCE distance is among
//      the row's k smallest is on the list (its approximate distance is <= (its reference distance) + eps <= tau + 2 eps).
//   2. knn_tc_exact_kernel.  One warp per row: the listed candidates (typically 1.5-2.5 k of them) are re-evaluated
//      with the reference's sequential FMA chain and formula, and the k smallest (value, index) pairs are taken, lowest
//      index first among equal values -- the same selection as knn_select_rows_kernel.  A row whose list overflowed is
//      re-evaluated over all K columns.
// Nothing of size K x K is stored.
#include <cuda.h>

#include "hg_common.cuh"

namespace {

constexpr int kTcRows = 128;    // query rows per CTA (UMMA M)
constexpr int kTcTile = 256;    // candidate columns per accumulator tile (UMMA N)
constexpr int kTcCap = 96;      // candidate slots per row
constexpr int kTcPanelK = 32;   // TF32 elements per 128-byte swizzled panel row

// ---- small PTX wrappers (tcgen05 / TMA / mbarrier) ----------------------------------------------------------------
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, unsigned parity) {
  // bounded spin: a protocol bug must surface as a trapped kernel, never as a hung GPU
  unsigned ok = 0;
  for (unsigned spin = 0; spin < (1u << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(hg_smem_addr(bar)), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tc_tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          hg_smem_addr(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(hg_smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(hg_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor of a K-major, 128-byte-swizzled panel ([rows] x 128 bytes, 8-row groups 1024 bytes
// apart): start address, LBO = 1 (unused for swizzled K-major), SBO = 1024 B, version 1 (Blackwell), SWIZZLE_128B
__device__ __forceinline__ uint64_t tc_smem_desc(const void *panel, int k_byte_offset) {
  const uint32_t addr = hg_smem_addr(panel) + (uint32_t)k_byte_offset;
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);        // bits [0,14)
  d |= (uint64_t)1 << 16;                         // leading byte offset (>>4) = 1
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset (>>4)
  d |= (uint64_t)1 << 46;                         // descriptor version
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = 256
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- 1. tensor-core filter ------------------------------------------------------------------------------------------
// dynamic shared memory (1024-byte aligned): A panels [C/32][128 rows][128 B], B panels [C/32][256 rows][128 B]
template <int C>
__global__ void __launch_bounds__(128) knn_tc_filter_kernel(const __grid_constant__ CUtensorMap map, int K, int k1,
                                                            int b0, const float *__restrict__ xx /*[B,K]*/,
                                                            const float *__restrict__ xxmax /*[B]*/,
                                                            int *__restrict__ cand /*[nb,K,kTcCap]*/,
                                                            int *__restrict__ cnt /*[nb,K]*/) {
  constexpr int NP = C / kTcPanelK;                 // panels along the contraction
  constexpr int kABytes = kTcRows * 128, kBBytes = kTcTile * 128;
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char *sA = smraw, *sB = smraw + NP * kABytes;
  float *sxx = reinterpret_cast<float *>(sB + NP * kBBytes);  // [ntiles*256] squared norms of the cloud (reference
                                                              // rounding), +inf past K: padding columns never qualify
  __shared__ __align__(8) uint64_t bar_a, bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const int bl = blockIdx.y, b = b0 + bl, row0 = blockIdx.x * kTcRows;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ntiles = (K + kTcTile - 1) / kTcTile;

  if (tid == 0) {
    hg_mbar_init(&bar_a, 1);
    hg_mbar_init(&bar_b, 1);
    hg_mbar_init(&bar_mma, 1);
    hg_mbar_init_fence();
  }
  if (warp == 0) {  // 256 columns of tensor memory for the 128 x 256 FP32 accumulator
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(hg_smem_addr(&tmem_base_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int j = tid; j < ntiles * kTcTile; j += 128) sxx[j] = j < K ? xx[(size_t)b * K + j] : CUDART_INF_F;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {  // the query block: NP boxes of [128 rows x 32 elements]
    hg_mbar_expect_tx(&bar_a, NP * kABytes);
    for (int p = 0; p < NP; ++p) tc_tma_load_3d(sA + p * kABytes, &map, &bar_a, p * kTcPanelK, row0, b);
  }

  const int i = row0 + tid;  // this thread's query row (TMEM lane tid)
  const float xi = i < K ? sxx[i] : 0.f;
  // |approximate - reference| <= eps for every column of this row (see the header)
  const float eps = 1.1f * 0.00390625f * sqrtf(xi * xxmax[b]) + 1e-5f * (xi + xxmax[b]);
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  float m[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) m[c] = CUDART_INF_F;
  float thr = 0.f;
  int n = 0;
  unsigned phase_b = 0, phase_m = 0;

  for (int pass = 0; pass < 2; ++pass) {
    for (int t = 0; t < ntiles; ++t) {
      if (tid == 0) {  // candidate tile t: NP panels, each two boxes of 128 rows; then the MMAs, then the commit
        hg_mbar_expect_tx(&bar_b, NP * kBBytes);
        for (int p = 0; p < NP; ++p) {
          tc_tma_load_3d(sB + p * kBBytes, &map, &bar_b, p * kTcPanelK, t * kTcTile, b);
          tc_tma_load_3d(sB + p * kBBytes + kABytes, &map, &bar_b, p * kTcPanelK, t * kTcTile + 128, b);
        }
        if (pass == 0 && t == 0) tc_mbar_wait(&bar_a, 0);
        tc_mbar_wait(&bar_b, phase_b);
        tc_fence_after();
        constexpr uint32_t idesc = tc_idesc(kTcRows, kTcTile);
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)  // four K = 8 steps (32 bytes each) inside a 128-byte panel row
            tc_mma_tf32(tmem, tc_smem_desc(sA + p * kABytes, ks * 32), tc_smem_desc(sB + p * kBBytes, ks * 32), idesc,
                        (p | ks) ? 1u : 0u);
        tc_commit(&bar_mma);  // arrives when the MMAs above have completed (implies fence::before_thread_sync)
      }
      phase_b ^= 1;
      tc_mbar_wait(&bar_mma, phase_m);
      phase_m ^= 1;
      __syncwarp();  // lane 0 of warp 0 took the producer branch above: reconverge before the warp-aligned loads
      tc_fence_after();
#pragma unroll 1
      for (int q = 0; q < kTcTile / 32; ++q) {
        float acc[32];
        tc_ld32(trow + (uint32_t)(q * 32), acc);
        const int j0 = t * kTcTile + q * 32;
        const float4 *sx4 = reinterpret_cast<const float4 *>(sxx + j0);  // broadcast loads: every thread reads the same
        if (pass == 0) {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 xj = sx4[c4];
            m[4 * c4] = fminf(m[4 * c4], fmaf(-2.0f, acc[4 * c4], xi + xj.x));
            m[4 * c4 + 1] = fminf(m[4 * c4 + 1], fmaf(-2.0f, acc[4 * c4 + 1], xi + xj.y));
            m[4 * c4 + 2] = fminf(m[4 * c4 + 2], fmaf(-2.0f, acc[4 * c4 + 2], xi + xj.z));
            m[4 * c4 + 3] = fminf(m[4 * c4 + 3], fmaf(-2.0f, acc[4 * c4 + 3], xi + xj.w));
          }
        } else {
          unsigned hits = 0u;  // bit c: column j0 + c is within the band
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 xj = sx4[c4];
            hits |= (fmaf(-2.0f, acc[4 * c4], xi + xj.x) <= thr ? 1u : 0u) << (4 * c4);
            hits |= (fmaf(-2.0f, acc[4 * c4 + 1], xi + xj.y) <= thr ? 1u : 0u) << (4 * c4 + 1);
            hits |= (fmaf(-2.0f, acc[4 * c4 + 2], xi + xj.z) <= thr ? 1u : 0u) << (4 * c4 + 2);
            hits |= (fmaf(-2.0f, acc[4 * c4 + 3], xi + xj.w) <= thr ? 1u : 0u) << (4 * c4 + 3);
          }
          if (i >= K) hits = 0u;
          while (hits) {  // ascending column order
            const int c = __ffs(hits) - 1;
            hits &= hits - 1u;
            if (n < kTcCap) cand[((size_t)bl * K + i) * kTcCap + n] = j0 + c;
            ++n;
          }
        }
      }
      tc_fence_before();
      __syncthreads();  // the accumulator and the B panels are free again
      tc_fence_after();
    }
    if (pass == 0) {
      // k1-th smallest of the 32 class minima: bitonic sort in registers (static indices), ascending
#pragma unroll
      for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1)
#pragma unroll
          for (int a = 0; a < 32; ++a) {
            const int p2 = a ^ stride;
            if (p2 > a) {
              const bool up = (a & size) == 0;
              const float lo = fminf(m[a], m[p2]), hi = fmaxf(m[a], m[p2]);
              m[a] = up ? lo : hi;
              m[p2] = up ? hi : lo;
            }
          }
      float tau = m[31];
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c == k1 - 1) tau = m[c];
      thr = tau + 2.0f * eps;  // +inf (fewer than k1 populated classes) lists every column: the row overflows
    }
  }
  if (i < K) cnt[(size_t)bl * K + i] = n;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---- 1b. fused single pass: tensor-core filter + exact FP32 evaluation out of the SAME shared-memory tiles ----------
// The candidate tile that TMA brought in for the MMA holds, row by row, exactly the FP32 features the reference
// arithmetic needs.  So a column the TF32 distances cannot rule out is evaluated on the spot -- the thread (= query row)
// keeps its own row in registers (C <= 64) or re-reads it from the A panels, walks the candidate's 128-byte-swizzled row
// with conflict-free 16-byte loads, runs the reference's sequential FMA chain and inserts (value, index) into a sorted
// register list of the row's KM smallest.  Columns are visited in ascending order, so a strict '<' insertion keeps the
// lowest index among equal values (torch.topk's order on this path, hg_knn.cu).  The admission threshold starts from
// the k-th smallest of 32 column-class minima of tile 0 (an upper bound on the k-th smallest approximate distance) and
// follows the list's k-th EXACT value from then on:  approx_j <= (k-th exact so far) + 2 eps  is necessary for column j
// to enter the final list, eps bounding |approximate - reference| for the row (header).  One pass over the tiles:
// half the TMA traffic and MMAs of the two-pass filter, no candidate lists in global memory, no L2 gathers.
template <int KM>
struct TcList {
  float v[KM];
  int id[KM];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < KM; ++s) {
      v[s] = CUDART_INF_F;
      id[s] = 0;
    }
  }
  // sorted insertion, strict '<': an equal value stays behind the entries already there
  __device__ __forceinline__ void push(float d, int j) {
#pragma unroll
    for (int s = KM - 1; s >= 1; --s) {
      const bool above = d < v[s - 1], here = d < v[s];
      id[s] = above ? id[s - 1] : (here ? j : id[s]);
      v[s] = above ? v[s - 1] : (here ? d : v[s]);
    }
    if (d < v[0]) {
      v[0] = d;
      id[0] = j;
    }
  }
  // k1-th smallest = maximum of the first k1 entries of the ascending list (written as a masked maximum: a
  // "select entry k1-1" chain is turned into a dynamically indexed load by the compiler, which sends the list to local memory)
  __device__ __forceinline__ float kth(int k1) const {
    float r = -CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < KM; ++s)
      if (s < k1) r = fmaxf(r, v[s]);
    return r;
  }
};

template <int C, int KM>
__global__ void __launch_bounds__(128, C <= 64 ? 2 : 1)
    knn_tc_fused_kernel(const __grid_constant__ CUtensorMap map, int K, int k1, int b0, const float *__restrict__ xx,
                        const float *__restrict__ xxmax, float *__restrict__ vals, int *__restrict__ idx) {
  constexpr int NP = C / kTcPanelK;
  constexpr int kABytes = kTcRows * 128, kBBytes = kTcTile * 128;
  constexpr bool kOwnInRegs = C <= 64;
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char *sA = smraw, *sB = smraw + NP * kABytes;
  float *sxx = reinterpret_cast<float *>(sB + NP * kBBytes);
  __shared__ __align__(8) uint64_t bar_a, bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const int bl = blockIdx.y, b = b0 + bl, row0 = blockIdx.x * kTcRows;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ntiles = (K + kTcTile - 1) / kTcTile;

  if (tid == 0) {
    hg_mbar_init(&bar_a, 1);
    hg_mbar_init(&bar_b, 1);
    hg_mbar_init(&bar_mma, 1);
    hg_mbar_init_fence();
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(hg_smem_addr(&tmem_base_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int j = tid; j < ntiles * kTcTile; j += 128) sxx[j] = j < K ? xx[(size_t)b * K + j] : CUDART_INF_F;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    hg_mbar_expect_tx(&bar_a, NP * kABytes);
    for (int p = 0; p < NP; ++p) tc_tma_load_3d(sA + p * kABytes, &map, &bar_a, p * kTcPanelK, row0, b);
  }

  const int i = row0 + tid;
  const bool live = i < K;
  const float xi = live ? sxx[i] : 0.f;
  const float eps = 1.1f * 0.00390625f * sqrtf(xi * xxmax[b]) + 1e-5f * (xi + xxmax[b]);
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const unsigned sw = (unsigned)(tid & 7);  // swizzle phase of this thread's own row in the A panels

  float4 own[kOwnInRegs ? C / 4 : 1];
  TcList<KM> top;
  top.init();
  float thr = CUDART_INF_F;
  unsigned phase_b = 0, phase_m = 0;

  // reference distances of columns jr0, jr1 of the current tile (two independent chains in flight): zz = sequential
  // FMA chain over the channels in ascending order, dist = (xx_j + (-2 zz)) + xx_i; + 0.0f turns a -0.0 into +0.0 like
  // the selection kernels do
  auto exact2 = [&](int jr0, int jr1, float xxj0, float xxj1, float &d0, float &d1) {
    float acc0 = 0.f, acc1 = 0.f;
    const unsigned sw0 = (unsigned)(jr0 & 7), sw1 = (unsigned)(jr1 & 7);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const unsigned char *r0 = sB + p * kBBytes + jr0 * 128, *r1 = sB + p * kBBytes + jr1 * 128;
      const unsigned char *rowi = sA + p * kABytes + tid * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v0 = *reinterpret_cast<const float4 *>(r0 + ((c ^ sw0) << 4));
        const float4 v1 = *reinterpret_cast<const float4 *>(r1 + ((c ^ sw1) << 4));
        const float4 a = kOwnInRegs ? own[kOwnInRegs ? p * 8 + c : 0]
                                    : *reinterpret_cast<const float4 *>(rowi + ((c ^ sw) << 4));
        acc0 = __fmaf_rn(a.x, v0.x, acc0);
        acc1 = __fmaf_rn(a.x, v1.x, acc1);
        acc0 = __fmaf_rn(a.y, v0.y, acc0);
        acc1 = __fmaf_rn(a.y, v1.y, acc1);
        acc0 = __fmaf_rn(a.z, v0.z, acc0);
        acc1 = __fmaf_rn(a.z, v1.z, acc1);
        acc0 = __fmaf_rn(a.w, v0.w, acc0);
        acc1 = __fmaf_rn(a.w, v1.w, acc1);
      }
    }
    d0 = __fadd_rn(__fadd_rn(__fadd_rn(xxj0, __fmul_rn(-2.0f, acc0)), xi), 0.0f);
    d1 = __fadd_rn(__fadd_rn(__fadd_rn(xxj1, __fmul_rn(-2.0f, acc1)), xi), 0.0f);
  };

  for (int t = 0; t < ntiles; ++t) {
    if (tid == 0) {
      hg_mbar_expect_tx(&bar_b, NP * kBBytes);
      for (int p = 0; p < NP; ++p) {
        tc_tma_load_3d(sB + p * kBBytes, &map, &bar_b, p * kTcPanelK, t * kTcTile, b);
        tc_tma_load_3d(sB + p * kBBytes + kABytes, &map, &bar_b, p * kTcPanelK, t * kTcTile + 128, b);
      }
      if (t == 0) tc_mbar_wait(&bar_a, 0);
      tc_mbar_wait(&bar_b, phase_b);
      tc_fence_after();
      constexpr uint32_t idesc = tc_idesc(kTcRows, kTcTile);
#pragma unroll
      for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc_mma_tf32(tmem, tc_smem_desc(sA + p * kABytes, ks * 32), tc_smem_desc(sB + p * kBBytes, ks * 32), idesc,
                      (p | ks) ? 1u : 0u);
      tc_commit(&bar_mma);
    }
    // every thread reads the tiles with ordinary loads below: each one observes the TMA completions itself
    if (t == 0) {
      tc_mbar_wait(&bar_a, 0);
      if (kOwnInRegs) {
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
          for (int c = 0; c < 8; ++c)
            own[kOwnInRegs ? p * 8 + c : 0] =
                *reinterpret_cast<const float4 *>(sA + p * kABytes + tid * 128 + ((c ^ sw) << 4));
      }
    }
    tc_mbar_wait(&bar_b, phase_b);
    phase_b ^= 1;
    tc_mbar_wait(&bar_mma, phase_m);
    phase_m ^= 1;
    __syncwarp();
    tc_fence_after();

    if (t == 0) {
      // threshold for the first tile: k1-th smallest of the 32 column-class minima of its approximate distances
      float m[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) m[c] = CUDART_INF_F;
#pragma unroll 1
      for (int q = 0; q < kTcTile / 32; ++q) {
        float acc[32];
        tc_ld32(trow + (uint32_t)(q * 32), acc);
        const float4 *sx4 = reinterpret_cast<const float4 *>(sxx + q * 32);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 xj = sx4[c4];
          m[4 * c4] = fminf(m[4 * c4], fmaf(-2.0f, acc[4 * c4], xi + xj.x));
          m[4 * c4 + 1] = fminf(m[4 * c4 + 1], fmaf(-2.0f, acc[4 * c4 + 1], xi + xj.y));
          m[4 * c4 + 2] = fminf(m[4 * c4 + 2], fmaf(-2.0f, acc[4 * c4 + 2], xi + xj.z));
          m[4 * c4 + 3] = fminf(m[4 * c4 + 3], fmaf(-2.0f, acc[4 * c4 + 3], xi + xj.w));
        }
      }
#pragma unroll
      for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1)
#pragma unroll
          for (int a = 0; a < 32; ++a) {
            const int p2 = a ^ stride;
            if (p2 > a) {
              const bool up = (a & size) == 0;
              const float lo = fminf(m[a], m[p2]), hi = fmaxf(m[a], m[p2]);
              m[a] = up ? lo : hi;
              m[p2] = up ? hi : lo;
            }
          }
      float tau = -CUDART_INF_F;  // m[k1 - 1] of the ascending array, as a masked maximum (see TcList::kth)
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < k1) tau = fmaxf(tau, m[c]);
      thr = tau + 2.0f * eps;
    }

    // the tile's 256-bit hit mask first, then every lane walks ITS hits (two per step) at its own pace: a warp takes
    // max-over-lanes(hits in the tile) / 2 steps, not the sum over the eight 32-column chunks of the per-chunk maxima
    unsigned hm[kTcTile / 32];
#pragma unroll
    for (int q = 0; q < kTcTile / 32; ++q) {
      float acc[32];
      tc_ld32(trow + (uint32_t)(q * 32), acc);
      const float4 *sx4 = reinterpret_cast<const float4 *>(sxx + t * kTcTile + q * 32);
      unsigned hits = 0u;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 xj = sx4[c4];
        hits |= (fmaf(-2.0f, acc[4 * c4], xi + xj.x) <= thr ? 1u : 0u) << (4 * c4);
        hits |= (fmaf(-2.0f, acc[4 * c4 + 1], xi + xj.y) <= thr ? 1u : 0u) << (4 * c4 + 1);
        hits |= (fmaf(-2.0f, acc[4 * c4 + 2], xi + xj.z) <= thr ? 1u : 0u) << (4 * c4 + 2);
        hits |= (fmaf(-2.0f, acc[4 * c4 + 3], xi + xj.w) <= thr ? 1u : 0u) << (4 * c4 + 3);
      }
      hm[q] = live ? hits : 0u;
    }
    auto next_hit = [&]() -> int {  // lowest set bit of the 256-bit mask, cleared; -1 when none is left
      unsigned w = 0u;
      int base = -1;
#pragma unroll
      for (int q = kTcTile / 32 - 1; q >= 0; --q)
        if (hm[q]) {
          w = hm[q];
          base = q;
        }
#pragma unroll
      for (int q = 0; q < kTcTile / 32; ++q)
        if (q == base) hm[q] = w & (w - 1u);
      return base < 0 ? -1 : base * 32 + __ffs(w) - 1;
    };
    while (true) {
      const int jr0 = next_hit();
      if (!__any_sync(0xffffffffu, jr0 >= 0)) break;
      if (jr0 >= 0) {
        const int jr1 = next_hit();
        const int jrb = jr1 >= 0 ? jr1 : jr0;
        float d0, d1;
        exact2(jr0, jrb, sxx[t * kTcTile + jr0], sxx[t * kTcTile + jrb], d0, d1);
        if (d0 < top.v[KM - 1]) top.push(d0, t * kTcTile + jr0);
        if (jr1 >= 0 && d1 < top.v[KM - 1]) top.push(d1, t * kTcTile + jr1);
      }
    }
    thr = fminf(thr, (k1 == KM ? top.v[KM - 1] : top.kth(k1)) + 2.0f * eps);
    tc_fence_before();
    __syncthreads();  // the accumulator and the B panels are free again
    tc_fence_after();
  }
  if (live) {
    const size_t o = ((size_t)b * K + i) * k1;
#pragma unroll
    for (int s = 0; s < KM; ++s)
      if (s < k1) {
        if (vals) vals[o + s] = top.v[s];
        idx[o + s] = top.id[s];
      }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---- 2. exact re-evaluation + selection ---------------------------------------------------------------------------
// reference distance (hg_knn.cu): zz = sequential FMA chain over the channels, dist = (xx_j + (-2 zz)) + xx_i
template <int C>
__device__ __forceinline__ float tc_exact_dist(const float *__restrict__ xi_s, const float *__restrict__ xj, float xxi,
                                               float xxj) {
  float acc = 0.f;
  const float4 *p = reinterpret_cast<const float4 *>(xj);
  const float4 *q = reinterpret_cast<const float4 *>(xi_s);
  // 64 channels at a time: all sixteen 16-byte loads of the candidate row in flight before the (sequential) FMA chain
  // starts -- the loop is bound by the latency of these L2 gathers, not by the chain
#pragma unroll
  for (int h = 0; h < C / 64 + (C % 64 ? 1 : 0); ++h) {
    constexpr int kChunk = 16;
    float4 v[kChunk];
#pragma unroll
    for (int c = 0; c < kChunk; ++c)
      if (h * kChunk + c < C / 4) v[c] = __ldg(p + h * kChunk + c);
#pragma unroll
    for (int c = 0; c < kChunk; ++c)
      if (h * kChunk + c < C / 4) {
        const float4 a = q[h * kChunk + c];
        acc = __fmaf_rn(a.x, v[c].x, acc);
        acc = __fmaf_rn(a.y, v[c].y, acc);
        acc = __fmaf_rn(a.z, v[c].z, acc);
        acc = __fmaf_rn(a.w, v[c].w, acc);
      }
  }
  return __fadd_rn(__fadd_rn(xxj, __fmul_rn(-2.0f, acc)), xxi);
}

// Exact -2*zz chains of the 32 candidates of one round (lane = candidate), with COALESCED gathers: read by their owner
// lanes, 32 candidate rows are 32 different 128-byte lines per load instruction (the first version: 33.5 M L1
// wavefronts per call, which is what bounded the kernel, not its FMA chains).  Here the warp loads whole rows together
// -- sixteen lanes per 256-byte half-row pair, full lines -- stages them in shared memory (row stride padded by one
// 16-byte chunk: conflict-free LDS.128 by the owners) and every lane then runs its chain out of shared memory,
// 64 channels per stage.  Returns the lane's accumulator (the reference's sequential FMA chain over all C channels).
constexpr int kStageChunks = 16;                   // 16-byte chunks of a candidate row per stage (64 channels)
constexpr int kStageStride = kStageChunks + 1;     // in float4
template <int C>
__device__ __forceinline__ float tc_chain_staged(const float *__restrict__ xi_s, const float *__restrict__ cloud, int myj,
                                                 float4 *__restrict__ stage /*[32][kStageStride]*/, int lane) {
  float acc = 0.f;
  const float4 *q4 = reinterpret_cast<const float4 *>(xi_s);
  constexpr int R = C / 4;  // chunks per row
#pragma unroll 1
  for (int h0 = 0; h0 < R; h0 += kStageChunks) {
    const int rh = (R - h0 < kStageChunks) ? (R - h0) : kStageChunks;  // chunks of this stage (8 or 16)
    __syncwarp();
    for (int g = lane; g < 32 * rh; g += 32) {  // chunk g of the stage: candidate g / rh, chunk g % rh
      const int cq = g / rh, ch = g - cq * rh;
      const int j = __shfl_sync(0xffffffffu, myj, cq);
      stage[cq * kStageStride + ch] = __ldg(reinterpret_cast<const float4 *>(cloud + (size_t)j * C) + h0 + ch);
    }
    __syncwarp();
    const float4 *mine = stage + lane * kStageStride;
#pragma unroll 4
    for (int c = 0; c < rh; ++c) {
      const float4 a = q4[h0 + c], v = mine[c];
      acc = __fmaf_rn(a.x, v.x, acc);
      acc = __fmaf_rn(a.y, v.y, acc);
      acc = __fmaf_rn(a.z, v.z, acc);
      acc = __fmaf_rn(a.w, v.w, acc);
    }
  }
  return acc;
}

// rank[u] = number of the row's candidate keys below mk[u] (broadcast 16-byte reads, two keys each; slots beyond NS idle)
template <int NS>
__device__ __forceinline__ void tc_rank(const unsigned long long *__restrict__ keys, int n,
                                        const unsigned long long (&mk)[kTcCap / 32], int (&rank)[kTcCap / 32]) {
  const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(keys);
  int t = 0;
  for (; t + 1 < n; t += 2) {
    const ulonglong2 kk = k2[t >> 1];
#pragma unroll
    for (int u = 0; u < NS; ++u) rank[u] += (kk.x < mk[u] ? 1 : 0) + (kk.y < mk[u] ? 1 : 0);
  }
  if (t < n) {
    const unsigned long long kt = keys[t];
#pragma unroll
    for (int u = 0; u < NS; ++u) rank[u] += kt < mk[u] ? 1 : 0;
  }
}

template <int C>
__global__ void __launch_bounds__(128, 6) knn_tc_exact_kernel(const float *__restrict__ pc /*[nb,K,C] of this batch*/,
                                                           const float *__restrict__ xx /*[nb,K]*/, int K, int k1,
                                                           int nrows, const int *__restrict__ cand,
                                                           const int *__restrict__ cnt, float *__restrict__ vals,
                                                           int *__restrict__ idx) {
  extern __shared__ __align__(16) float esm[];  // per warp: x_i [C], then the staging area [32][kStageStride] float4
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *xi_s = esm + (size_t)warp * C;
  float4 *stage = reinterpret_cast<float4 *>(esm + 4 * C) + (size_t)warp * 32 * kStageStride;
  __shared__ __align__(16) unsigned long long ck[4][kTcCap];
  for (int r = blockIdx.x * 4 + warp; r < nrows; r += gridDim.x * 4) {
    const int bl = r / K;
    const float *cloud = pc + (size_t)bl * K * C;
    const float *xr = pc + (size_t)r * C;
    __syncwarp();
    for (int c = lane; c < C; c += 32) xi_s[c] = xr[c];
    __syncwarp();
    const float xxi = xx[r];
    const float *xxb = xx + (size_t)bl * K;
    const int n = cnt[r];
    if (n <= kTcCap) {
      // exact distance of every listed candidate, as one sortable 64-bit key (order-preserving value bits, index)
      const int *crow = cand + (size_t)r * kTcCap;
      const int nslots = (n + 31) >> 5;  // warp-uniform
      unsigned long long mk[kTcCap / 32];
#pragma unroll
      for (int u = 0; u < kTcCap / 32; ++u) {
        mk[u] = ~0ull;
        if (u < nslots) {  // (warp-uniform)
          const int q = lane + 32 * u;
          const int j = crow[q < n ? q : 0];  // idle lanes shadow candidate 0 (valid memory), result discarded
          const float acc = tc_chain_staged<C>(xi_s, cloud, j, stage, lane);
          if (q < n) {
            // reference formula; + 0.0f: a -0.0 would sort before +0.0 as a key although the two compare equal
            const float d = __fadd_rn(__fadd_rn(__fadd_rn(xxb[j], __fmul_rn(-2.0f, acc)), xxi), 0.0f);
            mk[u] = ((unsigned long long)hg_ord(d) << 32) | (unsigned)j;
            ck[warp][q] = mk[u];
          }
        }
      }
      __syncwarp();
      // rank of every candidate among the row's candidates by (value, index): the keys are distinct, so the ranks are
      // a permutation; the candidates ranked below k1 are the answer, already in their output slots
      int rank[kTcCap / 32];
#pragma unroll
      for (int u = 0; u < kTcCap / 32; ++u) rank[u] = 0;
      if (nslots == 1) tc_rank<1>(ck[warp], n, mk, rank);
      else if (nslots == 2) tc_rank<2>(ck[warp], n, mk, rank);
      else tc_rank<3>(ck[warp], n, mk, rank);
#pragma unroll
      for (int u = 0; u < kTcCap / 32; ++u)
        if (u < nslots && lane + 32 * u < n && rank[u] < k1) {
          if (vals) vals[(size_t)r * k1 + rank[u]] = hg_unord((unsigned)(mk[u] >> 32));
          idx[(size_t)r * k1 + rank[u]] = (int)(unsigned)(mk[u] & 0xffffffffu);
        }
    } else {
      // list overflow (heavy ties, degenerate features; rare): k1 rounds of "smallest (value, index) after the previous
      // pick", each re-evaluating the whole row exactly -- slow, but needs no K-sized buffer (shared memory is what
      // bounds the number of resident warps here)
      float pv = -CUDART_INF_F;
      int pj = -1;
      for (int t = 0; t < k1; ++t) {
        float bv = CUDART_INF_F;
        int bj = 0x7fffffff;
        for (int j = lane; j < K; j += 32) {
          const float x = tc_exact_dist<C>(xi_s, cloud + (size_t)j * C, xxi, xxb[j]);
          const bool after = (x > pv) || (x == pv && j > pj);
          if (after && (x < bv || (x == bv && j < bj))) {
            bv = x;
            bj = j;
          }
        }
        const float wv = hg_warp_min_f32(bv);
        const int wj = __reduce_min_sync(0xffffffffu, (bv == wv) ? bj : 0x7fffffff);
        pv = wv;
        pj = wj;
        if (lane == 0) {
          if (vals) vals[(size_t)r * k1 + t] = wv;
          idx[(size_t)r * k1 + t] = wj;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) knn_tc_max_kernel(const float *__restrict__ xx, int K, float *__restrict__ out) {
  const int b = blockIdx.x;
  float m = 0.f;
  for (int j = threadIdx.x; j < K; j += 256) m = fmaxf(m, xx[(size_t)b * K + j]);
  m = hg_warp_max_f32(m);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    out[b] = m;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int C>
int run_tc(const float *pc, const float *xx, float *xxmax, int B, int K, int k1, float *vals, int *idx, int *cand,
           int *cnt, int nb_max, cudaStream_t stream) {
  EncodeTiledFn enc = encode_tiled();
  HG_REQUIRE(enc != nullptr, HG_E_UNSUPPORTED, "knn (tensor-core path): cuTensorMapEncodeTiled is not available");
  constexpr int NP = C / kTcPanelK;
  const size_t fsmem = (size_t)NP * (kTcRows + kTcTile) * 128 + (size_t)((K + kTcTile - 1) / kTcTile) * kTcTile * sizeof(float);
  const size_t esmem = (size_t)4 * C * sizeof(float) + (size_t)4 * 32 * kStageStride * sizeof(float4);
  static HgPerDeviceOnce once;
  if (once.first()) {
    HG_CUDA(cudaFuncSetAttribute(knn_tc_filter_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    HG_CUDA((cudaFuncSetAttribute(knn_tc_fused_kernel<C, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)));
    HG_CUDA((cudaFuncSetAttribute(knn_tc_fused_kernel<C, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)));
    HG_CUDA(cudaFuncSetAttribute(knn_tc_exact_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  }
  knn_tc_max_kernel<<<B, 256, 0, stream>>>(xx, K, xxmax);
  HG_CHECK_LAUNCH("knn_tc_max_kernel");
  for (int b0 = 0; b0 < B; b0 += nb_max) {
    const int nb = B - b0 < nb_max ? B - b0 : nb_max;
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)C, (cuuint64_t)K, (cuuint64_t)B};
    const cuuint64_t gstr[2] = {(cuuint64_t)C * sizeof(float), (cuuint64_t)K * C * sizeof(float)};
    const cuuint32_t box[3] = {kTcPanelK, kTcRows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(pc), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HG_REQUIRE(cr == CUDA_SUCCESS, HG_E_UNSUPPORTED, "knn (tensor-core path): cuTensorMapEncodeTiled failed (%d)", (int)cr);
    dim3 grid((K + kTcRows - 1) / kTcRows, nb);
    if (g_hg_tune_knn_tc_off != 3) {  // (3: the two-kernel filter + gather path, kept for A/B)
      // vals / idx are indexed with the global cloud number inside the fused kernel
      if (k1 <= 20)
        knn_tc_fused_kernel<C, 20><<<grid, 128, fsmem, stream>>>(map, K, k1, b0, xx, xxmax, vals, idx);
      else
        knn_tc_fused_kernel<C, 32><<<grid, 128, fsmem, stream>>>(map, K, k1, b0, xx, xxmax, vals, idx);
      HG_CHECK_LAUNCH("knn_tc_fused_kernel");
      continue;
    }
    knn_tc_filter_kernel<C><<<grid, 128, fsmem, stream>>>(map, K, k1, b0, xx, xxmax, cand, cnt);
    HG_CHECK_LAUNCH("knn_tc_filter_kernel");
    const int nrows = nb * K;
    int eg = (nrows + 3) / 4;
    const int cap = hg_sm_count() * 16;
    if (eg > cap) eg = cap;
    knn_tc_exact_kernel<C><<<eg, 128, esmem, stream>>>(pc + (size_t)b0 * K * C, xx + (size_t)b0 * K, K, k1, nrows, cand, cnt,
                                                      vals ? vals + (size_t)b0 * K * k1 : nullptr,
                                                      idx + (size_t)b0 * K * k1);
    HG_CHECK_LAUNCH("knn_tc_exact_kernel");
  }
  return HG_OK;
}

}  // namespace

// Shapes the tensor-core path takes: C a multiple of 32 up to 128 (DGCNN: 64, 64, 128), clouds of 256..4096 points.
bool hg_knn_tc_supported(int K, int C, int k1) {
  return (C == 32 || C == 64 || C == 96 || C == 128) && K >= 256 && K <= 4096 && k1 <= 32;
}

size_t hg_knn_tc_scratch_per_cloud(int K) { return (size_t)K * (kTcCap + 1) * sizeof(int); }

// xx [B,K] are the reference squared norms (knn_sumsq_kernel); scratch holds nb_max clouds' candidate lists + one
// float per cloud.
int hg_knn_tc_run(const float *pc, const float *xx, int B, int K, int C, int k1, float *vals, int *idx, void *scratch,
                  size_t scratch_bytes, cudaStream_t stream) {
  float *xxmax = (float *)scratch;
  const size_t head = hg_align((size_t)B * sizeof(float));
  HG_REQUIRE(scratch_bytes > head + hg_knn_tc_scratch_per_cloud(K), HG_E_WORKSPACE, "knn (tensor-core path): scratch too small");
  int nb_max = (int)((scratch_bytes - head) / hg_knn_tc_scratch_per_cloud(K));
  if (nb_max > B) nb_max = B;
  if (nb_max > 65535) nb_max = 65535;
  int *cand = (int *)((char *)scratch + head);
  int *cnt = cand + (size_t)nb_max * K * kTcCap;
  switch (C) {
    case 32: return run_tc<32>(pc, xx, xxmax, B, K, k1, vals, idx, cand, cnt, nb_max, stream);
    case 64: return run_tc<64>(pc, xx, xxmax, B, K, k1, vals, idx, cand, cnt, nb_max, stream);
    case 96: return run_tc<96>(pc, xx, xxmax, B, K, k1, vals, idx, cand, cnt, nb_max, stream);
    case 128: return run_tc<128>(pc, xx, xxmax, B, K, k1, vals, idx, cand, cnt, nb_max, stream);
  }
  hg_set_error("knn (tensor-core path): unsupported channel count %d", C);
  return HG_E_UNSUPPORTED;
}
