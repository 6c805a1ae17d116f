/*
 * hitgeom_oracle.c -- CPU restatement of the HiT-ADV point-set-geometry hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (hit-adv_b200/) may import, link or call
 * this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and there only as the checker / the timed CPU baseline.
 *
 * Parity status: PINNED for every function whose reference is Python/torch -- tests/test_oracle_golden.py
 * checks this file against the tests/golden npz fixtures, which were produced by importing the unmodified reference
 * from /root/reference (tests/golden/make_golden.py).  The pointnet2_ops functions (orc_p2_*) restate
 * CUDA kernels that cannot run in the build container; they are pinned against fixtures generated on a
 * B200 by the reference's own kernels compiled into oracle/_ref (tests/golden/make_golden_gpu.py).
 * orc_knn_points restates pytorch3d.ops.knn_points (pytorch3d==0.7.2, requirements.txt:8), whose source
 * is not part of the reference tree: PARITY UNPINNED for that one function (documented semantics only).
 *
 * Threading: every function is single-threaded and loops over the clouds [0,B); oracle.py splits B across
 * host threads (ctypes releases the GIL), which is how the multi-core CPU baseline is timed.
 *
 * Arithmetic conventions (SURVEY.md section 8 a-bis).  Built with -ffp-contract=off, so every + - * below
 * is an individually rounded FP32 operation and every fmaf() is a single-rounded fused multiply-add.
 *   dot_fma(a,b)  = fma(a[D-1],b[D-1], ... fma(a[1],b[1], a[0]*b[0]))   <- torch.bmm/matmul, small inner dim
 *   sumsq_seq(a)  = ((a0*a0 + a1*a1) + a2*a2) [cascade of 16s for C>=16] <- torch.sum(x**2, dim)
 *   dist_fma(a,b) = fma(dz,dz, fma(dx,dx, dy*dy)), dx=a0-b0 ...          <- nvcc -O3 contraction of the
 *                                                                          pointnet2_ops .cu expressions
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_version(void) { return 1; }

/* torch.bmm with a small inner dimension: sequential FMA chain, first product rounded on its own. */
static inline float dot_fma(const float *a, const float *b, int D) {
  float acc = a[0] * b[0];
  for (int c = 1; c < D; ++c) acc = fmaf(a[c], b[c], acc);
  return acc;
}

/* torch.sum(x**2, dim) on CPU: ATen's cascade sum (aten/src/ATen/native/cpu/SumKernel.cpp) -- 16-element
 * sequential partial sums, partials accumulated sequentially in a second level (third level from 256 on).
 * For fewer than 16 channels this is the plain left-to-right sum ((a0^2 + a1^2) + a2^2).  Probed bit-exact
 * against torch 2.11 for C = 3, 64, 128 (tests/golden/dgcnn_knn.npz). */
static inline float sumsq_seq(const float *a, int D) {
  float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;
  for (int c = 0; c < D; ++c) {
    acc0 = acc0 + a[c] * a[c];
    if (((c + 1) & 15) == 0) {
      acc1 = acc1 + acc0;
      acc0 = 0.0f;
      if (((c + 1) & 255) == 0) {
        acc2 = acc2 + acc1;
        acc1 = 0.0f;
      }
    }
  }
  return (acc0 + acc1) + acc2;
}

/* nvcc -O3 contracts  dx*dx + dy*dy + dz*dz  as fma(dz,dz, fma(dx,dx, round(dy*dy))): in `a*a + b*b` the LEFT
 * product is fused and the right one rounded.  Read off the SASS of the reference built for sm_100
 * (query_ball_point_kernel, three_nn_kernel, furthest_point_sampling_kernel) and pinned by
 * tests/golden/pointnet2_ref.npz, which those kernels produced on a B200. */
static inline float dist_fma3(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* pytorch3d's kernel accumulates `dist += diff*diff` coordinate by coordinate: fma(dz,dz, fma(dy,dy, dx*dx)) */
static inline float dist_seq3(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* ------------------------------------------------------------------------------------------------
 * util/set_distance.py:15-32  _Distance.batch_pairwise_dist(x=gts, y=preds)
 *   P[b,i,j] = (rx_i + ry_j) - 2*zz_ij, rx = diag(bmm(x,x^T)) (FMA chain), zz = bmm(x,y^T) (FMA chain)
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_pairwise_dist(const float *x, const float *y, int B, int Nx, int Ny, int D, float *P) {
  for (int b = 0; b < B; ++b) {
    const float *xb = x + (size_t)b * Nx * D, *yb = y + (size_t)b * Ny * D;
    float *ry = (float *)malloc(sizeof(float) * (size_t)(Ny > 0 ? Ny : 1));
    for (int j = 0; j < Ny; ++j) ry[j] = dot_fma(yb + (size_t)j * D, yb + (size_t)j * D, D);
    for (int i = 0; i < Nx; ++i) {
      const float *xi = xb + (size_t)i * D;
      float rx = dot_fma(xi, xi, D);
      float *Pi = P + ((size_t)b * Nx + i) * Ny;
      for (int j = 0; j < Ny; ++j) {
        float zz = dot_fma(xi, yb + (size_t)j * D, D);
        Pi[j] = (rx + ry[j]) - 2.0f * zz;
      }
    }
    free(ry);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * util/set_distance.py:45-49 / 63-69: the two torch.min reductions over P (first index wins ties).
 *   min1[b,j] = min_i P[b,i,j]  (torch.min(P,1): each pred/adv point -> nearest gt/ori)   arg1 = that i
 *   min2[b,i] = min_j P[b,i,j]  (torch.min(P,2): each gt/ori point  -> nearest pred/adv)  arg2 = that j
 * Matrix-free (P is never stored) so that the 16384-point configs fit; arithmetic identical to above.
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_nn_bidir(const float *gts, const float *preds, int B, int N2, int N1, int D, float *min1,
                         int *arg1, float *min2, int *arg2) {
  for (int b = 0; b < B; ++b) {
    const float *xb = gts + (size_t)b * N2 * D, *yb = preds + (size_t)b * N1 * D;
    float *m1 = min1 + (size_t)b * N1, *m2 = min2 + (size_t)b * N2;
    int *a1 = arg1 + (size_t)b * N1, *a2 = arg2 + (size_t)b * N2;
    float *ry = (float *)malloc(sizeof(float) * (size_t)(N1 > 0 ? N1 : 1));
    for (int j = 0; j < N1; ++j) {
      ry[j] = dot_fma(yb + (size_t)j * D, yb + (size_t)j * D, D);
      m1[j] = INFINITY;
      a1[j] = 0;
    }
    for (int i = 0; i < N2; ++i) {
      const float *xi = xb + (size_t)i * D;
      float rx = dot_fma(xi, xi, D);
      float best = INFINITY;
      int besti = 0;
      if (D == 3) {
        const float x0 = xi[0], x1 = xi[1], x2 = xi[2];
        for (int j = 0; j < N1; ++j) {
          const float *yj = yb + (size_t)j * 3;
          float zz = fmaf(x2, yj[2], fmaf(x1, yj[1], x0 * yj[0]));
          float p = (rx + ry[j]) - 2.0f * zz;
          if (p < best) { best = p; besti = j; }
          if (p < m1[j]) { m1[j] = p; a1[j] = i; }
        }
      } else {
        for (int j = 0; j < N1; ++j) {
          float zz = dot_fma(xi, yb + (size_t)j * D, D);
          float p = (rx + ry[j]) - 2.0f * zz;
          if (p < best) { best = p; besti = j; }
          if (p < m1[j]) { m1[j] = p; a1[j] = i; }
        }
      }
      m2[i] = best;
      a2[i] = besti;
    }
    free(ry);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * util/set_distance.py:46-49 (Chamfer: torch.mean of the mins) and :66-69 (Hausdorff: torch.max).
 * mode 0 = chamfer, 1 = hausdorff.  hd_arg*: first index attaining the max (torch.max(dim) semantics).
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_set_loss(const float *min1, const float *min2, int B, int N1, int N2, int mode, float *loss1,
                         float *loss2, int *hd_arg1, int *hd_arg2) {
  for (int b = 0; b < B; ++b) {
    const float *m1 = min1 + (size_t)b * N1, *m2 = min2 + (size_t)b * N2;
    if (mode == 0) {
      double s1 = 0, s2 = 0;
      for (int j = 0; j < N1; ++j) s1 += m1[j];
      for (int i = 0; i < N2; ++i) s2 += m2[i];
      loss1[b] = (float)(s1 / N1);
      loss2[b] = (float)(s2 / N2);
      if (hd_arg1) hd_arg1[b] = -1;
      if (hd_arg2) hd_arg2[b] = -1;
    } else {
      float v1 = -INFINITY, v2 = -INFINITY;
      int i1 = 0, i2 = 0;
      for (int j = 0; j < N1; ++j)
        if (m1[j] > v1) { v1 = m1[j]; i1 = j; }
      for (int i = 0; i < N2; ++i)
        if (m2[i] > v2) { v2 = m2[i]; i2 = i; }
      loss1[b] = v1;
      loss2[b] = v2;
      if (hd_arg1) hd_arg1[b] = i1;
      if (hd_arg2) hd_arg2[b] = i2;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Backward of the above (what autograd derives for set_distance.py:31,46-49,66-69), analytic, in
 * double: dP[i,j]/dy_j = 2(y_j - x_i), dP[i,j]/dx_i = 2(x_i - y_j); torch.min/max route the gradient
 * to the single saved index.  g1/g2 = dL/dloss1[b], dL/dloss2[b].  grad_gts may be NULL.
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_set_loss_bwd(const float *gts, const float *preds, const int *arg1, const int *arg2,
                             const int *hd_arg1, const int *hd_arg2, const float *g1, const float *g2, int B,
                             int N2, int N1, int D, int mode, float *grad_preds, float *grad_gts) {
  for (int b = 0; b < B; ++b) {
    const float *xb = gts + (size_t)b * N2 * D, *yb = preds + (size_t)b * N1 * D;
    const int *a1 = arg1 + (size_t)b * N1, *a2 = arg2 + (size_t)b * N2;
    double *gy = (double *)calloc((size_t)N1 * D + 1, sizeof(double));
    double *gx = (double *)calloc((size_t)N2 * D + 1, sizeof(double));
    if (mode == 0) {
      double c1 = (double)g1[b] / N1, c2 = (double)g2[b] / N2;
      for (int j = 0; j < N1; ++j) {
        int i = a1[j];
        for (int c = 0; c < D; ++c) {
          double d = 2.0 * ((double)yb[(size_t)j * D + c] - (double)xb[(size_t)i * D + c]);
          gy[(size_t)j * D + c] += c1 * d;
          gx[(size_t)i * D + c] -= c1 * d;
        }
      }
      for (int i = 0; i < N2; ++i) {
        int j = a2[i];
        for (int c = 0; c < D; ++c) {
          double d = 2.0 * ((double)yb[(size_t)j * D + c] - (double)xb[(size_t)i * D + c]);
          gy[(size_t)j * D + c] += c2 * d;
          gx[(size_t)i * D + c] -= c2 * d;
        }
      }
    } else {
      int j = hd_arg1[b], i = a1[j];
      for (int c = 0; c < D; ++c) {
        double d = 2.0 * ((double)yb[(size_t)j * D + c] - (double)xb[(size_t)i * D + c]);
        gy[(size_t)j * D + c] += (double)g1[b] * d;
        gx[(size_t)i * D + c] -= (double)g1[b] * d;
      }
      i = hd_arg2[b];
      j = a2[i];
      for (int c = 0; c < D; ++c) {
        double d = 2.0 * ((double)yb[(size_t)j * D + c] - (double)xb[(size_t)i * D + c]);
        gy[(size_t)j * D + c] += (double)g2[b] * d;
        gx[(size_t)i * D + c] -= (double)g2[b] * d;
      }
    }
    for (size_t t = 0; t < (size_t)N1 * D; ++t) grad_preds[(size_t)b * N1 * D + t] = (float)gy[t];
    if (grad_gts)
      for (size_t t = 0; t < (size_t)N2 * D; ++t) grad_gts[(size_t)b * N2 * D + t] = (float)gx[t];
    free(gy);
    free(gx);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Self k-NN over the asymmetric FP32 matrix used by both
 *   util/dist_utils.py:148-156 (KNNDist):  dist[i,j] = (xx_j + (-2*zz_ij)) + xx_i, topk(-dist, k+1)
 *   model/dgcnn_cls.py:8-12  (DGCNN knn):  pw[i,j] = ((-xx_j) - (-2*zz_ij)) - xx_i = -dist[i,j], topk(pw,k)
 * pc is point-major [B,K,C] (the host mirror transposes channel-major input).  xx = sequential no-FMA sum
 * of squares (torch.sum(x**2,dim)); zz = FMA chain (torch.matmul).  Output: the k1 smallest dist per row,
 * ascending, lowest index first among equal values (torch.topk leaves tie order unspecified; this is the
 * canonical order the parity tests use, SURVEY.md section 7 hard part 2).
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_knn_self(const float *pc, int B, int K, int C, int k1, float *vals, int *idx) {
  if (k1 > K) return 1;
  for (int b = 0; b < B; ++b) {
    const float *p = pc + (size_t)b * K * C;
    float *xx = (float *)malloc(sizeof(float) * (size_t)K);
    for (int j = 0; j < K; ++j) xx[j] = sumsq_seq(p + (size_t)j * C, C);
    for (int i = 0; i < K; ++i) {
      float *v = vals + ((size_t)b * K + i) * k1;
      int *id = idx + ((size_t)b * K + i) * k1;
      int cnt = 0;
      const float *pi = p + (size_t)i * C;
      for (int j = 0; j < K; ++j) {
        float zz = dot_fma(pi, p + (size_t)j * C, C);
        float d = (xx[j] + (-2.0f * zz)) + xx[i];
        if (cnt < k1 || d < v[cnt - 1]) {
          int t = cnt < k1 ? cnt : k1 - 1;
          while (t > 0 && d < v[t - 1]) {
            v[t] = v[t - 1];
            id[t] = id[t - 1];
            --t;
          }
          v[t] = d;
          id[t] = j;
          if (cnt < k1) ++cnt;
        }
      }
    }
    free(xx);
  }
  return 0;
}

/* util/dist_utils.py:157-167: value = mean of the k non-first neighbours, outlier mask, masked mean. */
ORC_API int orc_knn_outlier_fwd(const float *vals, int B, int K, int k1, float alpha, float *value, float *mask,
                                float *loss) {
  int k = k1 - 1;
  for (int b = 0; b < B; ++b) {
    double s = 0, ss = 0;
    for (int i = 0; i < K; ++i) {
      const float *v = vals + ((size_t)b * K + i) * k1;
      float acc = v[1];
      for (int t = 2; t < k1; ++t) acc = acc + v[t];
      float val = acc / (float)k;
      value[(size_t)b * K + i] = val;
      s += val;
    }
    double mean = s / K;
    for (int i = 0; i < K; ++i) {
      double d = (double)value[(size_t)b * K + i] - mean;
      ss += d * d;
    }
    float meanf = (float)mean, stdf = (float)sqrt(ss / (K - 1));
    float thr = meanf + alpha * stdf;
    double l = 0;
    for (int i = 0; i < K; ++i) {
      float m = value[(size_t)b * K + i] > thr ? 1.0f : 0.0f;
      mask[(size_t)b * K + i] = m;
      l += (double)(value[(size_t)b * K + i] * m);
    }
    loss[b] = (float)(l / K);
  }
  return 0;
}

/* backward of KNNDist through the saved top-k indices (SURVEY.md section 8a "Backward"), double. */
ORC_API int orc_knn_outlier_bwd(const float *pc, const int *idx, const float *mask, const float *g, int B, int K,
                                int C, int k1, float *grad_pc) {
  int k = k1 - 1;
  for (int b = 0; b < B; ++b) {
    const float *p = pc + (size_t)b * K * C;
    double *gp = (double *)calloc((size_t)K * C + 1, sizeof(double));
    double coef = (double)g[b] / ((double)K * k);
    for (int i = 0; i < K; ++i) {
      if (mask[(size_t)b * K + i] == 0.0f) continue;
      for (int t = 1; t < k1; ++t) {
        int n = idx[((size_t)b * K + i) * k1 + t];
        for (int c = 0; c < C; ++c) {
          double d = 2.0 * ((double)p[(size_t)i * C + c] - (double)p[(size_t)n * C + c]) * coef;
          gp[(size_t)i * C + c] += d;
          gp[(size_t)n * C + c] -= d;
        }
      }
    }
    for (size_t t = 0; t < (size_t)K * C; ++t) grad_pc[(size_t)b * K * C + t] = (float)gp[t];
    free(gp);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * pytorch3d.ops.knn_points(p1,p2,K) -- PARITY UNPINNED (third party, not in /root/reference; call sites
 * ShapeAttack/HiT_ADV.py:78-80,320-321; util/dist_utils.py:482-489; FGM/GeoA3_args.py:284).
 * Documented semantics: squared L2 from direct differences, K smallest ascending, int64 indices.
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_knn_points(const float *p1, const float *p2, int B, int N, int M, int K, float *dists,
                           int64_t *idx) {
  if (K > M) return 1;
  for (int b = 0; b < B; ++b) {
    const float *q = p1 + (size_t)b * N * 3, *r = p2 + (size_t)b * M * 3;
    for (int i = 0; i < N; ++i) {
      float *v = dists + ((size_t)b * N + i) * K;
      int64_t *id = idx + ((size_t)b * N + i) * K;
      int cnt = 0;
      for (int j = 0; j < M; ++j) {
        float d = dist_seq3(q[i * 3], q[i * 3 + 1], q[i * 3 + 2], r[j * 3], r[j * 3 + 1], r[j * 3 + 2]);
        if (cnt < K || d < v[cnt - 1]) {
          int t = cnt < K ? cnt : K - 1;
          while (t > 0 && d < v[t - 1]) {
            v[t] = v[t - 1];
            id[t] = id[t - 1];
            --t;
          }
          v[t] = d;
          id[t] = j;
          if (cnt < K) ++cnt;
        }
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * model/pointnet2_utils.py:19-40 square_distance(src,dst): ((-2*zz) + rs_n) + rd_m, rs/rd no-FMA sums.
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_square_distance(const float *src, const float *dst, int B, int N, int M, int C, float *out) {
  for (int b = 0; b < B; ++b) {
    const float *s = src + (size_t)b * N * C, *d = dst + (size_t)b * M * C;
    for (int n = 0; n < N; ++n) {
      float rs = sumsq_seq(s + (size_t)n * C, C);
      for (int m = 0; m < M; ++m) {
        float zz = dot_fma(s + (size_t)n * C, d + (size_t)m * C, C);
        float rd = sumsq_seq(d + (size_t)m * C, C);
        out[((size_t)b * N + n) * M + m] = ((-2.0f * zz) + rs) + rd;
      }
    }
  }
  return 0;
}

/* model/pointnet2_utils.py:63-84 torch farthest_point_sample; `start` = the torch.randint draw (:75). */
ORC_API int orc_fps_torch(const float *xyz, int B, int N, int npoint, const int64_t *start, int64_t *centroids) {
  for (int b = 0; b < B; ++b) {
    const float *p = xyz + (size_t)b * N * 3;
    float *distance = (float *)malloc(sizeof(float) * (size_t)N);
    for (int k = 0; k < N; ++k) distance[k] = 1e10f;
    int64_t farthest = start[b];
    for (int i = 0; i < npoint; ++i) {
      centroids[(size_t)b * npoint + i] = farthest;
      float cx = p[farthest * 3], cy = p[farthest * 3 + 1], cz = p[farthest * 3 + 2];
      float best = -INFINITY;
      int64_t besti = 0;
      for (int k = 0; k < N; ++k) {
        float dx = p[k * 3] - cx, dy = p[k * 3 + 1] - cy, dz = p[k * 3 + 2] - cz;
        float d = (dx * dx + dy * dy) + dz * dz;
        if (d < distance[k]) distance[k] = d;
        if (distance[k] > best) { best = distance[k]; besti = k; }
      }
      farthest = besti;
    }
    free(distance);
  }
  return 0;
}

/* model/pointnet2_utils.py:87-107 torch query_ball_point: keep d <= f32(r**2), ascending, pad with first. */
ORC_API int orc_query_ball_torch(float radius2_f32, int nsample, const float *xyz, const float *new_xyz, int B,
                                 int N, int S, int64_t *group_idx) {
  for (int b = 0; b < B; ++b) {
    const float *p = xyz + (size_t)b * N * 3, *q = new_xyz + (size_t)b * S * 3;
    for (int s = 0; s < S; ++s) {
      int64_t *g = group_idx + ((size_t)b * S + s) * nsample;
      float rs = sumsq_seq(q + (size_t)s * 3, 3);
      int cnt = 0;
      for (int n = 0; n < N && cnt < nsample; ++n) {
        float zz = dot_fma(q + (size_t)s * 3, p + (size_t)n * 3, 3);
        float d = ((-2.0f * zz) + rs) + sumsq_seq(p + (size_t)n * 3, 3);
        if (!(d > radius2_f32)) g[cnt++] = n;
      }
      int64_t first = cnt > 0 ? g[0] : (int64_t)N;
      for (int t = cnt; t < nsample; ++t) g[t] = first;
    }
  }
  return 0;
}

/* ================================================================================================
 * pointnet2_ops: restatements of the nine CUDA kernels, thread structure included where it decides
 * the result (FPS tie order).  Launch-shape helper: cuda_utils.h:13-17 opt_n_threads.
 * ============================================================================================== */
ORC_API int orc_p2_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

/* sampling_gpu.cu:69-173 furthest_point_sampling_kernel (+ sampling.cpp:66-87: temp filled with 1e10). */
ORC_API int orc_p2_fps(const float *dataset_all, int B, int n, int m, int *idxs_all) {
  if (m <= 0) return 0;
  const int bs = orc_p2_opt_n_threads(n);
  for (int b = 0; b < B; ++b) {
    const float *dataset = dataset_all + (size_t)b * n * 3;
    int *idxs = idxs_all + (size_t)b * m;
    float *temp = (float *)malloc(sizeof(float) * (size_t)n);
    float *dists = (float *)malloc(sizeof(float) * (size_t)bs);
    int *dists_i = (int *)malloc(sizeof(int) * (size_t)bs);
    for (int k = 0; k < n; ++k) temp[k] = 1e10f;
    int old = 0;
    idxs[0] = old;
    for (int j = 1; j < m; ++j) {
      float x1 = dataset[old * 3], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
      for (int tid = 0; tid < bs; ++tid) {
        int besti = 0;
        float best = -1;
        for (int k = tid; k < n; k += bs) {
          float x2 = dataset[k * 3], y2 = dataset[k * 3 + 1], z2 = dataset[k * 3 + 2];
          float mag = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
          if ((double)mag <= 1e-3) continue;
          float d = dist_fma3(x2, y2, z2, x1, y1, z1);
          float d2 = d < temp[k] ? d : temp[k];
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      for (int stride = bs / 2; stride >= 1; stride /= 2) {
        for (int tid = 0; tid < stride; ++tid) { /* __update(dists, dists_i, tid, tid+stride) :59-65 */
          float v1 = dists[tid], v2 = dists[tid + stride];
          int i1 = dists_i[tid], i2 = dists_i[tid + stride];
          dists[tid] = v1 > v2 ? v1 : v2;
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      idxs[j] = old;
    }
    free(temp);
    free(dists);
    free(dists_i);
  }
  return 0;
}

/* sampling_gpu.cu:8-20 */
ORC_API int orc_p2_gather(const float *points, const int *idx, int b, int c, int n, int m, float *out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j)
        out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]];
  return 0;
}

/* sampling_gpu.cu:34-47 (atomicAdd there; here a fixed j-ascending order, accumulated in double) */
ORC_API int orc_p2_gather_grad(const float *grad_out, const int *idx, int b, int c, int n, int m,
                               float *grad_points) {
  double *acc = (double *)calloc((size_t)b * c * n + 1, sizeof(double));
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j)
        acc[((size_t)i * c + l) * n + idx[(size_t)i * m + j]] += grad_out[((size_t)i * c + l) * m + j];
  for (size_t t = 0; t < (size_t)b * c * n; ++t) grad_points[t] = (float)acc[t];
  free(acc);
  return 0;
}

/* ball_query_gpu.cu:9-44 (+ ball_query.cpp:19-21: idx zero-initialised) */
ORC_API int orc_p2_ball_query(const float *new_xyz, const float *xyz, int b, int n, int m, float radius,
                              int nsample, int *idx) {
  const float radius2 = radius * radius;
  memset(idx, 0, sizeof(int) * (size_t)b * m * nsample);
  for (int bi = 0; bi < b; ++bi) {
    const float *p = xyz + (size_t)bi * n * 3, *q = new_xyz + (size_t)bi * m * 3;
    for (int j = 0; j < m; ++j) {
      int *o = idx + ((size_t)bi * m + j) * nsample;
      int cnt = 0;
      for (int k = 0; k < n && cnt < nsample; ++k) {
        float d2 = dist_fma3(q[j * 3], q[j * 3 + 1], q[j * 3 + 2], p[k * 3], p[k * 3 + 1], p[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[l] = k;
          o[cnt] = k;
          ++cnt;
        }
      }
    }
  }
  return 0;
}

/* group_points_gpu.cu:8-28 */
ORC_API int orc_p2_group(const float *points, const int *idx, int b, int c, int n, int npoints, int nsample,
                         float *out) {
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k)
          out[(((size_t)bi * c + l) * npoints + j) * nsample + k] =
              points[((size_t)bi * c + l) * n + idx[((size_t)bi * npoints + j) * nsample + k]];
  return 0;
}

/* group_points_gpu.cu:43-64 (atomicAdd there; fixed (j,k)-ascending order in double here) */
ORC_API int orc_p2_group_grad(const float *grad_out, const int *idx, int b, int c, int n, int npoints,
                              int nsample, float *grad_points) {
  double *acc = (double *)calloc((size_t)b * c * n + 1, sizeof(double));
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k)
          acc[((size_t)bi * c + l) * n + idx[((size_t)bi * npoints + j) * nsample + k]] +=
              grad_out[(((size_t)bi * c + l) * npoints + j) * nsample + k];
  for (size_t t = 0; t < (size_t)b * c * n; ++t) grad_points[t] = (float)acc[t];
  free(acc);
  return 0;
}

/* interpolate_gpu.cu:9-59 three_nn_kernel: strict '<' insertion, double accumulators, float d. */
ORC_API int orc_p2_three_nn(const float *unknown, const float *known, int b, int n, int m, float *dist2,
                            int *idx) {
  for (int bi = 0; bi < b; ++bi) {
    const float *u = unknown + (size_t)bi * n * 3, *kn = known + (size_t)bi * m * 3;
    for (int j = 0; j < n; ++j) {
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        float d = dist_fma3(u[j * 3], u[j * 3 + 1], u[j * 3 + 2], kn[k * 3], kn[k * 3 + 1], kn[k * 3 + 2]);
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      float *o = dist2 + ((size_t)bi * n + j) * 3;
      int *oi = idx + ((size_t)bi * n + j) * 3;
      o[0] = (float)best1; o[1] = (float)best2; o[2] = (float)best3;
      oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
    }
  }
  return 0;
}

/* interpolate_gpu.cu:72-101: out = p1*w1 + p2*w2 + p3*w3, nvcc contraction fma(p3,w3, fma(p1,w1, round(p2*w2))) */
ORC_API int orc_p2_three_interpolate(const float *points, const int *idx, const float *weight, int b, int c,
                                     int m, int n, float *out) {
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        const int *id = idx + ((size_t)bi * n + j) * 3;
        const float *w = weight + ((size_t)bi * n + j) * 3;
        const float *p = points + ((size_t)bi * c + l) * m;
        out[((size_t)bi * c + l) * n + j] = fmaf(p[id[2]], w[2], fmaf(p[id[0]], w[0], p[id[1]] * w[1]));
      }
  return 0;
}

/* interpolate_gpu.cu:116-143 (atomicAdd there; fixed (j,t)-ascending order in double here) */
ORC_API int orc_p2_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int b,
                                          int c, int n, int m, float *grad_points) {
  double *acc = (double *)calloc((size_t)b * c * m + 1, sizeof(double));
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        const int *id = idx + ((size_t)bi * n + j) * 3;
        const float *w = weight + ((size_t)bi * n + j) * 3;
        float g = grad_out[((size_t)bi * c + l) * n + j];
        for (int t = 0; t < 3; ++t) acc[((size_t)bi * c + l) * m + id[t]] += (double)(g * w[t]);
      }
  for (size_t t = 0; t < (size_t)b * c * m; ++t) grad_points[t] = (float)acc[t];
  free(acc);
  return 0;
}
