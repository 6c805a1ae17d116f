// hg_knn3.cu -- streaming k-nearest-neighbour selection for 3-D points (KNNDist, DGCNN layer 1, knn_points).
//
// Per (query, candidate) pair the FMA pipe does the same 5 operations as the Chamfer kernel (FMUL, 2 FFMA,
// 2 FADD, issued packed as FMUL2/FFMA2/FADD2 over two CANDIDATES at a time); what differs is the selection.
// A sorted-insertion test per pair would make every warp take the slow path on almost every candidate
// (32 lanes x QT queries, each inserting ~k ln(N/k) times), so selection is split in two:
//   * main loop (uniform, no divergence): for each query and candidate pair, gm = min(d, d') and ONE
//     comparison against that query's threshold (its current k-th value); a hit only sets a bit in a
//     per-query 32-bit mask;
//   * drain (divergent, rare): after each sub-tile every lane walks its hit bits, re-evaluates the two
//     candidates with the same arithmetic and inserts into its sorted list, then refreshes the thresholds.
// The first sub-tiles are short (8, 8, 16, 32, 64 ... candidates) so that the thresholds tighten
// geometrically and the number of stale hits stays ~k per doubling.
// Lists live in registers (statically indexed; KM = 6 / 20 / 32 entries, the first k are written out).
//
// Order of visits per query is ascending candidate index (sub-tiles ascending, bits ascending), insertion is
// strict '<' => equal values keep the lowest index first: the canonical tie order of the oracle.
#include "hg_common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kTileC = 256;  // candidates per shared-memory tile (128 pairs)

// exact scalar distance, identical operation sequence to the packed main loop
template <int FORM>
__device__ __forceinline__ float knn_dist_exact(float a0, float a1, float a2, float a3, float c0, float c1, float c2,
                                                float cw) {
  if (FORM == HG_KNN_FORM_EXPANDED) {  // a = (-2q0,-2q1,-2q2, xx_i), cw = xx_j
    const float nzz = __fmaf_rn(a2, c2, __fmaf_rn(a1, c1, __fmul_rn(a0, c0)));  // == -2*zz exactly
    return __fadd_rn(__fadd_rn(cw, nzz), a3);
  } else {  // a = (q0,q1,q2,-), c = NEGATED candidate: d = fma(dz,dz, fma(dy,dy, dx*dx)), dx = q0 + (-c0)
    const float dx = __fadd_rn(a0, c0), dy = __fadd_rn(a1, c1), dz = __fadd_rn(a2, c2);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
  }
}

// Sorted list held in registers (static indexing only).  Candidates reach a query in ascending index order; a
// candidate enters only if strictly smaller than the current last element and bubbles up while strictly
// smaller than its predecessor => lowest index first on ties.
template <int KM>
struct TopK {
  float v[KM];
  int id[KM];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int t = 0; t < KM; ++t) {
      v[t] = CUDART_INF_F;
      id[t] = 0;
    }
  }
  __device__ __forceinline__ void push(float d, int j) {
    if (d < v[KM - 1]) {
      v[KM - 1] = d;
      id[KM - 1] = j;
#pragma unroll
      for (int t = KM - 1; t > 0; --t) {
        if (v[t] < v[t - 1]) {
          const float tv = v[t];
          v[t] = v[t - 1];
          v[t - 1] = tv;
          const int ti = id[t];
          id[t] = id[t - 1];
          id[t - 1] = ti;
        }
      }
    }
  }
};

// KM >= k1 list entries are kept; the first k1 are written out.
template <int FORM, int QT, int KM, typename IdxT>
__global__ void __launch_bounds__(kThreads) knn3_kernel(const float *__restrict__ queries,
                                                        const float *__restrict__ refs, int Nq, int Nr, int k1,
                                                        float *__restrict__ vals, IdxT *__restrict__ idx) {
  __shared__ float4 cand[kTileC];  // two float4 per candidate pair
  const int b = blockIdx.y, tid = threadIdx.x;
  const float *q = queries + (size_t)b * Nq * 3;
  const float *r = refs + (size_t)b * Nr * 3;

  constexpr int MW = kTileC / 64;  // hit-mask words per query: one bit per candidate pair of a tile
  float a0[QT], a1[QT], a2[QT], a3[QT], thr[QT];
  unsigned mask[QT][MW];
  TopK<KM> top[QT];
#pragma unroll
  for (int t = 0; t < QT; ++t) {
    const int i = (blockIdx.x * QT + t) * kThreads + tid;
    float q0 = 0.f, q1 = 0.f, q2 = 0.f;
    if (i < Nq) {
      q0 = __ldg(q + (size_t)i * 3);
      q1 = __ldg(q + (size_t)i * 3 + 1);
      q2 = __ldg(q + (size_t)i * 3 + 2);
    }
    if (FORM == HG_KNN_FORM_EXPANDED) {
      a0[t] = -2.0f * q0;
      a1[t] = -2.0f * q1;
      a2[t] = -2.0f * q2;
      a3[t] = hg_sumsq3_seq(q0, q1, q2);
    } else {
      a0[t] = q0;
      a1[t] = q1;
      a2[t] = q2;
      a3[t] = 0.f;
    }
    thr[t] = CUDART_INF_F;
#pragma unroll
    for (int w = 0; w < MW; ++w) mask[t][w] = 0u;
    top[t].init();
  }

  for (int base = 0; base < Nr; base += kTileC) {
    __syncthreads();
    // stage the tile as candidate PAIRS: A = (c0_j, c0_j1, c1_j, c1_j1), B = (c2_j, c2_j1, w_j, w_j1)
    for (int p = tid; p < kTileC / 2; p += kThreads) {
      float c[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = base + 2 * p + h;
        c[h][0] = c[h][1] = c[h][2] = 0.f;
        c[h][3] = CUDART_INF_F;  // padding can never beat a threshold (EXPANDED: inf; DIRECT: handled below)
        if (j < Nr) {
          const float x = __ldg(r + (size_t)j * 3), y = __ldg(r + (size_t)j * 3 + 1), z = __ldg(r + (size_t)j * 3 + 2);
          if (FORM == HG_KNN_FORM_EXPANDED) {
            c[h][0] = x; c[h][1] = y; c[h][2] = z;
            c[h][3] = hg_sumsq3_seq(x, y, z);
          } else {
            c[h][0] = -x; c[h][1] = -y; c[h][2] = -z;
            c[h][3] = 0.f;
          }
        } else if (FORM != HG_KNN_FORM_EXPANDED) {
          c[h][0] = CUDART_INF_F;  // dx = q + inf = inf -> d = inf
        }
      }
      cand[2 * p] = make_float4(c[0][0], c[1][0], c[0][1], c[1][1]);
      cand[2 * p + 1] = make_float4(c[0][2], c[1][2], c[0][3], c[1][3]);
    }
    __syncthreads();
    const int npairs = (min(kTileC, Nr - base) + 1) >> 1;

    // Windows between two drains.  First tile: 4,4,8,16,32,32,32 pairs so that the thresholds tighten
    // geometrically.  Later tiles: the whole tile in one window -- hits are rare per lane but frequent per warp
    // there, and a longer window lets several lanes insert in the same (divergent) drain iteration.
    int p0 = 0;
    int step = (base == 0) ? 4 : kTileC / 2;
    while (p0 < npairs) {
      const int cnt = min(step, npairs - p0);
      // ---- main loop: packed distances, one threshold test per (query, candidate pair) ----------------------
#pragma unroll
      for (int w = 0; w < MW; ++w) {
        const int gcnt = min(32, cnt - 32 * w);
#pragma unroll 2
        for (int g = 0; g < gcnt; ++g) {
          const float4 cA = cand[2 * (p0 + 32 * w + g)], cB = cand[2 * (p0 + 32 * w + g) + 1];
          const unsigned bit = 1u << g;
#pragma unroll
          for (int t = 0; t < QT; ++t) {
            float2 d;
            if (FORM == HG_KNN_FORM_EXPANDED) {
              float2 tt = __fmul2_rn(make_float2(a0[t], a0[t]), make_float2(cA.x, cA.y));
              tt = __ffma2_rn(make_float2(a1[t], a1[t]), make_float2(cA.z, cA.w), tt);
              tt = __ffma2_rn(make_float2(a2[t], a2[t]), make_float2(cB.x, cB.y), tt);
              const float2 s = __fadd2_rn(make_float2(cB.z, cB.w), tt);
              d = __fadd2_rn(s, make_float2(a3[t], a3[t]));
            } else {
              const float2 dx = __fadd2_rn(make_float2(a0[t], a0[t]), make_float2(cA.x, cA.y));
              const float2 dy = __fadd2_rn(make_float2(a1[t], a1[t]), make_float2(cA.z, cA.w));
              const float2 dz = __fadd2_rn(make_float2(a2[t], a2[t]), make_float2(cB.x, cB.y));
              d = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            }
            if (fminf(d.x, d.y) < thr[t]) mask[t][w] |= bit;
          }
        }
      }
      // ---- drain: each lane inserts its own hits (ascending candidate order per query) ----------------------
#pragma unroll
      for (int t = 0; t < QT; ++t) {
        while (true) {
          int w = -1;
          unsigned m = 0u;
#pragma unroll
          for (int u = MW - 1; u >= 0; --u)
            if (mask[t][u]) {
              w = u;
              m = mask[t][u];
            }
          if (w < 0) break;
          const int g = __ffs(m) - 1;
#pragma unroll
          for (int u = 0; u < MW; ++u)
            if (u == w) mask[t][u] = m & (m - 1u);
          const int pp = p0 + 32 * w + g;
          const float4 cA = cand[2 * pp], cB = cand[2 * pp + 1];
          const int j0 = base + 2 * pp;
          top[t].push(knn_dist_exact<FORM>(a0[t], a1[t], a2[t], a3[t], cA.x, cA.z, cB.x, cB.z), j0);
          top[t].push(knn_dist_exact<FORM>(a0[t], a1[t], a2[t], a3[t], cA.y, cA.w, cB.y, cB.w), j0 + 1);
        }
        thr[t] = top[t].v[KM - 1];
      }
      p0 += cnt;
      if (base == 0 && step < 32 && p0 >= 2 * step) step *= 2;  // 4,4,8,16,32,32,...
    }
  }

#pragma unroll
  for (int t = 0; t < QT; ++t) {
    const int i = (blockIdx.x * QT + t) * kThreads + tid;
    if (i < Nq) {
#pragma unroll
      for (int s = 0; s < KM; ++s) {
        if (s < k1) {
          if (vals) vals[((size_t)b * Nq + i) * k1 + s] = top[t].v[s];
          idx[((size_t)b * Nq + i) * k1 + s] = (IdxT)top[t].id[s];
        }
      }
    }
  }
}

template <int FORM, int QT, int KM, typename IdxT>
int launch_qt(const float *q, const float *r, int B, int Nq, int Nr, int k1, float *vals, IdxT *idx,
              cudaStream_t stream) {
  dim3 grid((Nq + QT * kThreads - 1) / (QT * kThreads), B);
  const bool prof = hg_prof_begin(HG_PROF_KNN, stream);
  knn3_kernel<FORM, QT, KM, IdxT><<<grid, kThreads, 0, stream>>>(q, r, Nq, Nr, k1, vals, idx);
  hg_prof_end(HG_PROF_KNN, stream, prof);
  HG_CHECK_LAUNCH("knn3_kernel");
  return HG_OK;
}

template <int FORM, typename IdxT>
int launch_form(const float *q, const float *r, int B, int Nq, int Nr, int k1, float *vals, IdxT *idx,
                cudaStream_t stream) {
  if (k1 < 1 || k1 > 32) {
    hg_set_error("knn: k=%d outside [1,32]", k1);
    return HG_E_UNSUPPORTED;
  }
  // queries per lane: amortise the candidate loads, but keep small query sets spread over the machine and the
  // register-resident lists (2*KM registers per query) within budget
  if (k1 <= 6) {
    if (Nq >= 3 * kThreads) return launch_qt<FORM, 4, 6, IdxT>(q, r, B, Nq, Nr, k1, vals, idx, stream);
    if (Nq > kThreads) return launch_qt<FORM, 2, 6, IdxT>(q, r, B, Nq, Nr, k1, vals, idx, stream);
    return launch_qt<FORM, 1, 6, IdxT>(q, r, B, Nq, Nr, k1, vals, idx, stream);
  }
  if (k1 <= 20) {
    if (Nq > kThreads) return launch_qt<FORM, 2, 20, IdxT>(q, r, B, Nq, Nr, k1, vals, idx, stream);
    return launch_qt<FORM, 1, 20, IdxT>(q, r, B, Nq, Nr, k1, vals, idx, stream);
  }
  return launch_qt<FORM, 1, 32, IdxT>(q, r, B, Nq, Nr, k1, vals, idx, stream);
}

}  // namespace

int hg_knn3_launch_i32(int form, const float *q, const float *r, int B, int Nq, int Nr, int k1, float *vals, int *idx,
                       cudaStream_t stream) {
  return form == HG_KNN_FORM_EXPANDED
             ? launch_form<HG_KNN_FORM_EXPANDED, int>(q, r, B, Nq, Nr, k1, vals, idx, stream)
             : launch_form<HG_KNN_FORM_DIRECT, int>(q, r, B, Nq, Nr, k1, vals, idx, stream);
}

int hg_knn3_launch_i64(int form, const float *q, const float *r, int B, int Nq, int Nr, int k1, float *vals,
                       long long *idx, cudaStream_t stream) {
  return form == HG_KNN_FORM_EXPANDED
             ? launch_form<HG_KNN_FORM_EXPANDED, long long>(q, r, B, Nq, Nr, k1, vals, idx, stream)
             : launch_form<HG_KNN_FORM_DIRECT, long long>(q, r, B, Nq, Nr, k1, vals, idx, stream);
}
