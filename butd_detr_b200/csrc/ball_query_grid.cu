// Ball query through a uniform grid (cell list) — same results, bit for bit, as the brute-force
// scan of the reference (`query_ball_point_kernel`, ball_query_gpu.cu:14-49): the first `nsample`
// in-ball points in ASCENDING INDEX ORDER, remaining slots = first hit, empty ball = zeros.
//
// The reference (and bd_ball_query) tests every centre against every point: 2048 x 50 000 =
// 102 M distance tests per scene at SA1, although a ball only holds ~60 points.  Here points are
// binned once per scene into cells of edge >= radius (counting sort with atomics; order inside a
// cell is irrelevant), a warp visits the 27 cells around its centre, collects the hits in shared
// memory and selects the `nsample` smallest indices by rank — which restores the reference's
// index order exactly, independent of the binning order.  Distances use the reference's
// FMUL/FFMA/FFMA contraction (common.cuh), so the hit set is identical.
#include "common.cuh"

namespace {

constexpr int G_MAX = 1 << 16;   // cells per scene
constexpr int Q_WARPS = 8;
constexpr int Q_CAP = 768;       // hits kept per centre before falling back to the ordered scan
constexpr int Q_SORT = 1024;     // Q_CAP rounded up to a power of two (bitonic sort padding)

struct GridMeta {
  float minx, miny, minz, inv_cell;
  int gx, gy, gz, pad;
};

// order-preserving float <-> uint mapping, so min / max can use integer atomics
__device__ __forceinline__ unsigned f2o(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// partial bounding boxes: grid (BB_PARTS, B); bbox[b] = {min x,y,z, max x,y,z} as ordered uints
constexpr int BB_PARTS = 32;
__global__ void __launch_bounds__(256) bq_bbox_kernel(const float *__restrict__ xyz, int ld, int n,
                                                      unsigned *__restrict__ bbox) {
  const int b = blockIdx.y;
  xyz += static_cast<long long>(b) * n * ld;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += BB_PARTS * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __ldg(xyz + static_cast<long long>(i) * ld + c);
      mn[c] = fminf(mn[c], v), mx[c] = fmaxf(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xFFFFFFFFu, mn[c], off));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xFFFFFFFFu, mx[c], off));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(bbox + b * 8 + c, f2o(mn[c]));
      atomicMax(bbox + b * 8 + 4 + c, f2o(mx[c]));
    }
  }
}

__global__ void bq_bbox_init_kernel(unsigned *__restrict__ bbox, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * 8) bbox[i] = (i & 4) ? 0u : 0xFFFFFFFFu;  // max slots start at the smallest key, min slots at the largest
}

__global__ void bq_meta_kernel(const unsigned *__restrict__ bbox, float radius, GridMeta *__restrict__ meta) {
  const int b = threadIdx.x;  // one thread per scene (launched <<<1, B>>>, B <= 1024)
  float lo[3], hi[3];
  for (int c = 0; c < 3; ++c) lo[c] = o2f(bbox[b * 8 + c]), hi[c] = o2f(bbox[b * 8 + 4 + c]);
  // cell edge: slightly larger than the radius (rounding margin), doubled until the grid fits G_MAX
  float cell = radius * 1.001f;
  int g[3];
  for (;;) {
    long long total = 1;
    for (int c = 0; c < 3; ++c) {
      const float ext = fmaxf(hi[c] - lo[c], 0.f);
      g[c] = static_cast<int>(fminf(ext / cell, 1.0e6f)) + 1;
      total *= g[c];
    }
    if (total <= G_MAX) break;
    cell *= 2.f;
  }
  GridMeta m;
  m.minx = lo[0], m.miny = lo[1], m.minz = lo[2], m.inv_cell = 1.0f / cell;
  m.gx = g[0], m.gy = g[1], m.gz = g[2], m.pad = 0;
  meta[b] = m;
}

__device__ __forceinline__ int cell_coord(float v, float mn, float inv) { return static_cast<int>(floorf((v - mn) * inv)); }

__global__ void bq_count_kernel(const float *__restrict__ xyz, int ld, int n, const GridMeta *__restrict__ meta,
                                int *__restrict__ cell_of, int *__restrict__ count) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const GridMeta m = meta[b];
  const float *p = xyz + (static_cast<long long>(b) * n + i) * ld;
  const int cx = min(max(cell_coord(__ldg(p), m.minx, m.inv_cell), 0), m.gx - 1);
  const int cy = min(max(cell_coord(__ldg(p + 1), m.miny, m.inv_cell), 0), m.gy - 1);
  const int cz = min(max(cell_coord(__ldg(p + 2), m.minz, m.inv_cell), 0), m.gz - 1);
  const int cell = (cz * m.gy + cy) * m.gx + cx;
  cell_of[static_cast<long long>(b) * n + i] = cell;
  atomicAdd(count + static_cast<long long>(b) * (G_MAX + 1) + cell, 1);
}

// exclusive scan of the G_MAX cell counts of one scene (in place: count -> start), cursor = start
__global__ void __launch_bounds__(1024) bq_scan_kernel(int *__restrict__ count, int *__restrict__ cursor,
                                                       const GridMeta *__restrict__ meta) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int *c = count + static_cast<long long>(b) * (G_MAX + 1);
  int *cur = cursor + static_cast<long long>(b) * G_MAX;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  const GridMeta gm = meta[b];
  const int cells = gm.gx * gm.gy * gm.gz;  // cells beyond the grid keep start = total (set below)
  const int span = (cells + 1023) / 1024 * 1024;
  for (int base = 0; base < span; base += 1024) {
    const int v = c[base + tid];
    int x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int y = __shfl_up_sync(0xFFFFFFFFu, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, w, off);
        if (lane >= off) w += y;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + (warp ? warp_sum[warp - 1] : 0) + x - v;
    c[base + tid] = excl;
    cur[base + tid] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = excl + v;
    __syncthreads();
  }
  const int total = carry_s;
  for (int i = span + tid; i <= G_MAX; i += 1024) c[i] = total;
}

// cell-ordered copies: the index (bd_grid_order) and (x, y, z, index) — the query then streams
// 16-byte records of contiguous cells instead of chasing indices into the 24-byte point rows
__global__ void bq_fill_kernel(const float *__restrict__ xyz, int ld, int n, const int *__restrict__ cell_of,
                               int *__restrict__ cursor, int *__restrict__ sorted, float4 *__restrict__ sorted_pts) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cell = cell_of[static_cast<long long>(b) * n + i];
  const int pos = atomicAdd(cursor + static_cast<long long>(b) * G_MAX + cell, 1);
  sorted[static_cast<long long>(b) * n + pos] = i;
  const float *p = xyz + (static_cast<long long>(b) * n + i) * ld;
  sorted_pts[static_cast<long long>(b) * n + pos] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __int_as_float(i));
}

__global__ void __launch_bounds__(Q_WARPS * 32)
bq_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int ld, int n, int m, float radius2,
                int nsample, const GridMeta *__restrict__ meta, const int *__restrict__ start,
                const float4 *__restrict__ sorted, int *__restrict__ idx) {
  __shared__ int hits[Q_WARPS][Q_SORT];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * Q_WARPS + warp;
  if (j >= m) return;
  const GridMeta g = meta[b];
  xyz += static_cast<long long>(b) * n * ld;
  start += static_cast<long long>(b) * (G_MAX + 1);
  sorted += static_cast<long long>(b) * n;
  const float *c = new_xyz + (static_cast<long long>(b) * m + j) * 3;
  const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
  int *row = idx + (static_cast<long long>(b) * m + j) * nsample;
  int *h = hits[warp];
  int cnt = 0;
  bool overflow = false;
  // centre cell, NOT clamped (a centre outside the cloud's box simply sees fewer cells)
  const int ix = cell_coord(cx, g.minx, g.inv_cell), iy = cell_coord(cy, g.miny, g.inv_cell),
            iz = cell_coord(cz, g.minz, g.inv_cell);
  for (int dz = -1; dz <= 1 && !overflow; ++dz) {
    const int z = iz + dz;
    if (z < 0 || z >= g.gz) continue;
    for (int dy = -1; dy <= 1 && !overflow; ++dy) {
      const int y = iy + dy;
      if (y < 0 || y >= g.gy) continue;
      // the three x-neighbours are contiguous cells: one contiguous range of `sorted`
      const int x0 = max(ix - 1, 0), x1 = min(ix + 1, g.gx - 1);
      if (x0 > x1) continue;
      const int cell0 = (z * g.gy + y) * g.gx + x0;
      const int s0 = __ldg(start + cell0), s1 = __ldg(start + cell0 + (x1 - x0) + 1);
      for (int t0 = s0; t0 < s1; t0 += 32) {
        const int t = t0 + lane;
        int k = -1;
        bool hit = false;
        if (t < s1) {
          const float4 q = __ldg(sorted + t);
          k = __float_as_int(q.w);
          hit = bd::sqdist_ref(cx, cy, cz, q.x, q.y, q.z) < radius2;
        }
        const unsigned ballot = __ballot_sync(0xFFFFFFFFu, hit);
        if (ballot) {
          const int pos = cnt + __popc(ballot & ((1u << lane) - 1u));
          if (cnt + __popc(ballot) > Q_CAP) { overflow = true; break; }
          if (hit) h[pos] = k;
          cnt += __popc(ballot);
        }
      }
    }
  }
  __syncwarp();
  if (overflow) {
    // very dense ball: ordered brute-force scan for this centre (reference algorithm)
    int c2 = 0, first = 0;
    for (int i0 = 0; i0 < n && c2 < nsample; i0 += 32) {
      const int i = i0 + lane;
      bool hit = false;
      if (i < n) {
        const float *p = xyz + static_cast<long long>(i) * ld;
        hit = bd::sqdist_ref(cx, cy, cz, __ldg(p), __ldg(p + 1), __ldg(p + 2)) < radius2;
      }
      const unsigned ballot = __ballot_sync(0xFFFFFFFFu, hit);
      if (ballot) {
        if (c2 == 0) first = i0 + __ffs(ballot) - 1;
        const int pos = c2 + __popc(ballot & ((1u << lane) - 1u));
        if (hit && pos < nsample) row[pos] = i;
        c2 += __popc(ballot);
      }
    }
    for (int s = min(c2, nsample) + lane; s < nsample; s += 32) row[s] = first;
    return;
  }
  int first = 0x7FFFFFFF;
  if (cnt <= 64) {
    // rank selection: hit i goes to slot #{hits with a smaller index}; slots >= nsample are dropped
    for (int i = lane; i < cnt; i += 32) {
      const int ki = h[i];
      int rank = 0;
      for (int q = 0; q < cnt; ++q) rank += (h[q] < ki);
      if (rank < nsample) row[rank] = ki;
      first = min(first, ki);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) first = min(first, __shfl_xor_sync(0xFFFFFFFFu, first, off));
    if (cnt == 0) first = 0;
  } else {
    // many hits: warp-wide bitonic sort of the (padded) list in shared memory, O(n log^2 n / 32)
    // instead of the O(n^2 / 32) rank count; the first nsample entries are the answer
    int n2 = 128;
    while (n2 < cnt) n2 <<= 1;
    for (int i = cnt + lane; i < n2; i += 32) h[i] = 0x7FFFFFFF;
    __syncwarp();
    for (int k2 = 2; k2 <= n2; k2 <<= 1) {
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < (n2 >> 1); i += 32) {
          const int l = 2 * j * (i / j) + (i % j), r = l + j;  // j is a power of two: shifts after unswitching
          const int a = h[l], c = h[r];
          const bool up = (l & k2) == 0;
          if ((a > c) == up) { h[l] = c; h[r] = a; }
        }
        __syncwarp();
      }
    }
    for (int s = lane; s < nsample && s < cnt; s += 32) row[s] = h[s];
    first = h[0];
  }
  for (int s = min(cnt, nsample) + lane; s < nsample; s += 32) row[s] = first;
}

}  // namespace

extern "C" long long bd_ball_query_grid_workspace_bytes(int B, int n) {
  // meta | count/start (G_MAX+1) | cursor (G_MAX) | cell_of (n) | sorted (n)   per scene, ints
  // + (x, y, z, index) records in cell order (16-byte aligned)
  return static_cast<long long>(B) * (sizeof(GridMeta) + sizeof(int) * (2LL * G_MAX + 1 + 2LL * n) + 32 + 16LL * n) + 96;
}

namespace {
struct GridWs {
  GridMeta *meta;
  int *count, *cursor, *cell_of, *sorted;
  unsigned *bbox;
  float4 *sorted_pts;
};
GridWs grid_ws(void *workspace, int B, int n) {
  unsigned char *ws = static_cast<unsigned char *>(workspace);
  GridWs g;
  g.meta = reinterpret_cast<GridMeta *>(ws);
  g.count = reinterpret_cast<int *>(ws + static_cast<size_t>(B) * sizeof(GridMeta));
  g.cursor = g.count + static_cast<size_t>(B) * (G_MAX + 1);
  g.cell_of = g.cursor + static_cast<size_t>(B) * G_MAX;
  g.sorted = g.cell_of + static_cast<size_t>(B) * n;
  g.bbox = reinterpret_cast<unsigned *>(g.sorted + static_cast<size_t>(B) * n);
  uintptr_t a = reinterpret_cast<uintptr_t>(g.bbox + static_cast<size_t>(B) * 8);
  g.sorted_pts = reinterpret_cast<float4 *>((a + 15) & ~static_cast<uintptr_t>(15));
  return g;
}
}  // namespace

// Cell list of B clouds of n points (cells of edge >= radius): the build half of
// bd_ball_query_grid.  Afterwards bd_grid_order() is the points' indices grouped by cell (x-fastest
// cell order) — a spatially coherent permutation, also used by bd_fps_ordered.
extern "C" int bd_grid_build(const float *xyz, int ld_xyz, int B, int n, float radius, void *workspace,
                             bd_stream_t stream) {
  BD_REQUIRE(xyz && workspace, "bd_grid_build: null pointer");
  BD_REQUIRE(B > 0 && B <= 1024 && n > 0 && ld_xyz >= 3 && radius > 0.f, "bd_grid_build: bad sizes");
  cudaStream_t s = bd::as_stream(stream);
  const GridWs g = grid_ws(workspace, B, n);
  BD_CUDA(cudaMemsetAsync(g.count, 0, sizeof(int) * static_cast<size_t>(B) * (G_MAX + 1), s), "bd_grid_build");
  bq_bbox_init_kernel<<<bd::ceil_div(B * 8, 256), 256, 0, s>>>(g.bbox, B);
  bq_bbox_kernel<<<dim3(BB_PARTS, B), 256, 0, s>>>(xyz, ld_xyz, n, g.bbox);
  bq_meta_kernel<<<1, B, 0, s>>>(g.bbox, radius, g.meta);
  dim3 pgrid(bd::ceil_div(n, 256), B);
  bq_count_kernel<<<pgrid, 256, 0, s>>>(xyz, ld_xyz, n, g.meta, g.cell_of, g.count);
  bq_scan_kernel<<<B, 1024, 0, s>>>(g.count, g.cursor, g.meta);
  bq_fill_kernel<<<pgrid, 256, 0, s>>>(xyz, ld_xyz, n, g.cell_of, g.cursor, g.sorted, g.sorted_pts);
  BD_CHECK_LAUNCH("bd_grid_build");
  return BD_OK;
}

extern "C" const int *bd_grid_order(void *workspace, int B, int n) {
  return workspace ? grid_ws(workspace, B, n).sorted : nullptr;
}

// Query half: `workspace` holds the cell list built by bd_grid_build with the SAME xyz / radius.
extern "C" int bd_ball_query_grid_query(const float *new_xyz, const float *xyz, int ld_xyz, int B, int n, int m,
                                        float radius, int nsample, int *idx, void *workspace, bd_stream_t stream) {
  BD_REQUIRE(new_xyz && xyz && idx && workspace, "bd_ball_query_grid_query: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && m > 0 && nsample > 0 && ld_xyz >= 3 && radius > 0.f, "bd_ball_query_grid_query: bad sizes");
  const GridWs g = grid_ws(workspace, B, n);
  dim3 qgrid(bd::ceil_div(m, Q_WARPS), B);
  bq_query_kernel<<<qgrid, Q_WARPS * 32, 0, bd::as_stream(stream)>>>(new_xyz, xyz, ld_xyz, n, m, radius * radius, nsample,
                                                                      g.meta, g.count, g.sorted_pts, idx);
  BD_CHECK_LAUNCH("bd_ball_query_grid_query");
  return BD_OK;
}

extern "C" int bd_ball_query_grid(const float *new_xyz, const float *xyz, int ld_xyz, int B, int n, int m, float radius,
                                  int nsample, int *idx, void *workspace, bd_stream_t stream) {
  BD_REQUIRE(new_xyz && xyz && idx && workspace, "bd_ball_query_grid: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && m > 0 && nsample > 0 && ld_xyz >= 3 && radius > 0.f, "bd_ball_query_grid: bad sizes");
  BD_REQUIRE(B <= 1024, "bd_ball_query_grid: B too large");
  const int rc = bd_grid_build(xyz, ld_xyz, B, n, radius, workspace, stream);
  if (rc != BD_OK) return rc;
  return bd_ball_query_grid_query(new_xyz, xyz, ld_xyz, B, n, m, radius, nsample, idx, workspace, stream);
}
