// hg_csr.cu -- deterministic reverse map for scatter-style backward passes.
//
// The reference accumulates every scatter gradient with float atomicAdd (sampling_gpu.cu:43,
// group_points_gpu.cu:60, interpolate_gpu.cu:139-141) or with ATen index_put(accumulate=True); the sum
// order, and therefore the low bits of the result, change from run to run.  Here each destination gets the
// list of its source edges in ascending edge order (a stable counting sort per cloud); consumers add in
// list order, so results are bit-reproducible.  Integer atomics (counts) are order-independent.
#include "hg_common.cuh"

namespace {

constexpr int kCsrThreads = 256;
constexpr int kCsrWarps = kCsrThreads / 32;

// One CTA per cloud.  The edge range is cut into kCsrWarps contiguous segments, one per warp: per-warp key
// counts -> exclusive prefix over (key, warp) -> every warp places its own segment stably (chunks of 32 edges in
// ascending order, rank inside a chunk from __match_any_sync), all warps in parallel.
__global__ void __launch_bounds__(kCsrThreads) csr_build_kernel(const int *__restrict__ keys, int E, int N,
                                                                int *__restrict__ off_all,
                                                                int *__restrict__ cursor_all,
                                                                int *__restrict__ list_all) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int *k = keys + (size_t)b * E;
  int *off = off_all + (size_t)b * (N + 1);
  int *cursor = cursor_all + (size_t)b * kCsrWarps * N;  // [kCsrWarps][N]
  int *list = list_all + (size_t)b * E;
  __shared__ int warp_tot[kCsrWarps];
  __shared__ int carry_s;

  const int seg = ((E + kCsrWarps - 1) / kCsrWarps + 31) / 32 * 32;  // edges per warp segment (multiple of 32)
  const int e_lo = min(E, warp * seg), e_hi = min(E, (warp + 1) * seg);
  int *mine = cursor + (size_t)warp * N;

  for (int i = tid; i <= N; i += kCsrThreads) off[i] = 0;
  for (int i = tid; i < kCsrWarps * N; i += kCsrThreads) cursor[i] = 0;
  __syncthreads();
  for (int e = e_lo + lane; e < e_hi; e += 32) {
    const int key = k[e];
    if (key >= 0 && key < N) {
      atomicAdd(&off[key + 1], 1);
      atomicAdd(&mine[key], 1);
    }
  }
  if (tid == 0) carry_s = 0;
  __syncthreads();
  // inclusive scan of off[1..N] (off[0] stays 0) -> off[i] = #edges with key < i
  for (int base = 1; base <= N; base += kCsrThreads) {
    const int i = base + tid;
    int v = (i <= N) ? off[i] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += t;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    int add = carry_s;
    for (int w = 0; w < warp; ++w) add += warp_tot[w];
    if (i <= N) off[i] = v + add;
    __syncthreads();
    if (tid == kCsrThreads - 1) carry_s = v + add;
    __syncthreads();
  }
  // per-warp start of every key: segment start + edges of that key in the earlier warps' segments
  for (int key = tid; key < N; key += kCsrThreads) {
    int run = off[key];
    for (int w = 0; w < kCsrWarps; ++w) {
      const int c = cursor[(size_t)w * N + key];
      cursor[(size_t)w * N + key] = run;
      run += c;
    }
  }
  __syncthreads();
  for (int base = e_lo; base < e_hi; base += 32) {
    const int e = base + lane;
    int key = (e < e_hi) ? k[e] : -1;
    if (key >= N) key = -1;
    const unsigned same = __match_any_sync(0xffffffffu, key);
    const int rank = __popc(same & ((1u << lane) - 1u));
    int pos = 0;
    if (key >= 0) pos = mine[key] + rank;
    __syncwarp();
    if (key >= 0) {
      list[pos] = e;
      if (rank == 0) mine[key] = pos + __popc(same);
    }
    __syncwarp();
  }
}

}  // namespace

// Tiny problems (a couple of thousand edges per cloud) are launch-bound: one stable single-kernel build beats memset +
// count + scan + fill.  Beyond that the single CTA per cloud of the stable builder is the bottleneck (48 us against
// ~20 us for 388 clouds x 6144 edges).  hg_csr_build_unordered switches on this and the workspace formula follows it.
static inline bool csr_small(int E) { return E <= 2048; }

size_t hg_csr_workspace_bytes(int B, int N, int E) {
  const size_t cursors = csr_small(E) ? (size_t)kCsrWarps * N : (size_t)N;
  return hg_align((size_t)B * (N + 1) * sizeof(int)) + hg_align((size_t)B * cursors * sizeof(int)) +
         hg_align((size_t)B * E * sizeof(int));
}

size_t hg_csr_stable_workspace_bytes(int B, int N, int E) {
  return hg_align((size_t)B * (N + 1) * sizeof(int)) + hg_align((size_t)B * kCsrWarps * N * sizeof(int)) +
         hg_align((size_t)B * E * sizeof(int));
}

int hg_csr_build(const int *keys, int B, int E, int N, void *workspace, size_t workspace_bytes, HgCsr *out,
                 cudaStream_t stream) {
  HG_REQUIRE(workspace && workspace_bytes >= hg_csr_stable_workspace_bytes(B, N, E), HG_E_WORKSPACE,
             "csr: workspace too small (%zu < %zu)", workspace_bytes, hg_csr_stable_workspace_bytes(B, N, E));
  char *p = (char *)workspace;
  int *off = (int *)p;
  p += hg_align((size_t)B * (N + 1) * sizeof(int));
  int *cursor = (int *)p;
  p += hg_align((size_t)B * kCsrWarps * N * sizeof(int));
  int *list = (int *)p;
  csr_build_kernel<<<B, kCsrThreads, 0, stream>>>(keys, E, N, off, cursor, list);
  HG_CHECK_LAUNCH("csr_build_kernel");
  out->off = off;
  out->list = list;
  return HG_OK;
}

// ---- unordered variant for the hot-loop backward passes ---------------------------------------------------------
// Same off/list layout, but the position of an edge inside its destination's segment is taken from an integer
// atomic (fully parallel over the edges, no serial placement).  Consumers then walk a segment in ascending edge
// order with hg_csr_next() (selection over the few entries), so the summation order -- and the result -- is still
// independent of how the atomics resolved.  Segments are tiny on this path (Chamfer: multiplicity <= ~3; kNN:
// only edges leaving outlier points), which is what makes the O(c^2) walk cheaper than a stable sort.
namespace {

__global__ void __launch_bounds__(256) csr_count_kernel(const int *__restrict__ keys, long long total, int E, int N,
                                                        int *__restrict__ off_all) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int key = keys[g];
    if (key >= 0 && key < N) atomicAdd(off_all + (g / E) * (N + 1) + key + 1, 1);
  }
}

__global__ void __launch_bounds__(256) csr_scan_kernel(int N, int *__restrict__ off_all, int *__restrict__ cursor_all) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int *off = off_all + (size_t)b * (N + 1);
  int *cursor = cursor_all + (size_t)b * N;
  __shared__ int warp_tot[8];
  __shared__ int carry_s;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 1; base <= N; base += 256) {
    const int i = base + tid;
    int v = (i <= N) ? off[i] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += t;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    int add = carry_s;
    for (int w = 0; w < warp; ++w) add += warp_tot[w];
    if (i <= N) {
      off[i] = v + add;
      if (i < N) cursor[i] = v + add;  // start of segment i
    }
    __syncthreads();
    if (tid == 255) carry_s = v + add;
    __syncthreads();
  }
  if (tid == 0) cursor[0] = 0;
}

__global__ void __launch_bounds__(256) csr_fill_kernel(const int *__restrict__ keys, long long total, int E, int N,
                                                       int *__restrict__ cursor_all, int *__restrict__ list_all) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int key = keys[g];
    if (key >= 0 && key < N) {
      const long long b = g / E;
      const int pos = atomicAdd(cursor_all + b * N + key, 1);
      list_all[b * E + pos] = (int)(g - b * E);
    }
  }
}

}  // namespace

int hg_csr_build_unordered(const int *keys, int B, int E, int N, void *workspace, size_t workspace_bytes, HgCsr *out,
                           cudaStream_t stream) {
  HG_REQUIRE(workspace && workspace_bytes >= hg_csr_workspace_bytes(B, N, E), HG_E_WORKSPACE,
             "csr: workspace too small (%zu < %zu)", workspace_bytes, hg_csr_workspace_bytes(B, N, E));
  if (csr_small(E)) return hg_csr_build(keys, B, E, N, workspace, workspace_bytes, out, stream);  // sorted is fine too
  char *p = (char *)workspace;
  int *off = (int *)p;
  p += hg_align((size_t)B * (N + 1) * sizeof(int));
  int *cursor = (int *)p;
  p += hg_align((size_t)B * N * sizeof(int));
  int *list = (int *)p;
  HG_CUDA(cudaMemsetAsync(off, 0, (size_t)B * (N + 1) * sizeof(int), stream));
  const long long total = (long long)B * E;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)hg_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  csr_count_kernel<<<(int)blocks, 256, 0, stream>>>(keys, total, E, N, off);
  HG_CHECK_LAUNCH("csr_count_kernel");
  csr_scan_kernel<<<B, 256, 0, stream>>>(N, off, cursor);
  HG_CHECK_LAUNCH("csr_scan_kernel");
  csr_fill_kernel<<<(int)blocks, 256, 0, stream>>>(keys, total, E, N, cursor, list);
  HG_CHECK_LAUNCH("csr_fill_kernel");
  out->off = off;
  out->list = list;
  return HG_OK;
}
