// Ball query through a uniform grid (cell list) — same results, bit for bit, as the brute-force
// scan of the reference (`query_ball_point_kernel`, ball_query_gpu.cu:14-49): the first `nsample`
// in-ball points in ASCENDING INDEX ORDER, remaining slots = first hit, empty ball = zeros.
//
// The reference (and bd_ball_query) tests every centre against every point: 2048 x 50 000 =
// 102 M distance tests per scene at SA1, although a ball only holds ~60 points.  Here points are
// binned once per scene into cells of edge >= radius (counting sort with atomics; order inside a
// cell is irrelevant), a warp visits the cells around its centre and keeps the `nsample` SMALLEST
// in-ball indices — which restores the reference's index order exactly, independent of the binning
// order.  Distances use the reference's FMUL/FFMA/FFMA contraction (common.cuh), so the hit set is
// identical.
//
// Selection (one warp per centre, no shared-memory sort): hits are appended to a small per-warp
// buffer; the warp keeps a sorted register file of 128 candidates (4 per lane) and merges the buffer
// into its upper half with a shuffle-based bitonic network whenever 64 new hits have arrived; once
// the file is full, the current nsample-th smallest index becomes a THRESHOLD and later hits above
// it are dropped at the distance test — dense balls cost no more than sparse ones, there is no hit
// cap and no fallback path.  The same cell list serves the bucketed furthest point sampling
// (bd_fps_grid, fps.cu), which reads the cell-ordered records.
#include "common.cuh"

namespace {

constexpr int G_MAX = 1 << 16;   // cells per scene
constexpr int Q_WARPS = 8;
constexpr int Q_BUF = 160;       // per-warp hit buffer: a merge is due at 128 (first) / 64 (later) entries, +31 slack
constexpr int Q_NS_MAX = 64;     // nsample handled by the register selection (larger: bd_ball_query)
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int I_MAX = 0x7FFFFFFF;

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
// (non-finite coordinates are ignored: they end up clamped into a border cell)
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
      if (isfinite(v)) mn[c] = fminf(mn[c], v), mx[c] = fmaxf(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(FULL, mn[c], off));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(FULL, mx[c], off));
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

// one thread per scene.  radius > 0: cell edge slightly larger than the radius (rounding margin),
// doubled until the grid fits G_MAX cells (at most 64 times, then a single cell); radius <= 0
// (callers that only want a spatial order): largest extent / 32.
__global__ void bq_meta_kernel(const unsigned *__restrict__ bbox, float radius, int B, GridMeta *__restrict__ meta) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float lo[3], ext[3];
  for (int c = 0; c < 3; ++c) {
    lo[c] = o2f(bbox[b * 8 + c]);
    const float hi = o2f(bbox[b * 8 + 4 + c]);
    ext[c] = hi - lo[c];
    if (!(ext[c] >= 0.f) || !isfinite(ext[c])) lo[c] = 0.f, ext[c] = 0.f;  // empty / non-finite axis: one cell
  }
  float cell = radius > 0.f ? radius * 1.001f : fmaxf(fmaxf(ext[0], ext[1]), ext[2]) * (1.0f / 32.0f);
  if (!(cell > 0.f) || !isfinite(cell)) cell = 1.0f;
  int g[3] = {1, 1, 1};
  bool fits = false;
  for (int it = 0; it < 64 && !fits; ++it) {
    long long total = 1;
    for (int c = 0; c < 3; ++c) {
      g[c] = static_cast<int>(fminf(ext[c] / cell, 1.0e6f)) + 1;
      total *= g[c];
    }
    fits = total <= G_MAX;
    if (!fits) cell *= 2.f;
  }
  if (!fits) g[0] = g[1] = g[2] = 1;
  GridMeta m;
  m.minx = lo[0], m.miny = lo[1], m.minz = lo[2], m.inv_cell = 1.0f / cell;
  m.gx = g[0], m.gy = g[1], m.gz = g[2], m.pad = 0;
  meta[b] = m;
}

// cell coordinate; the float -> int conversion saturates and maps NaN to 0, callers clamp
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

// exclusive scan of the cell counts of one scene (in place: count -> start), cursor = start
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
  const int cells = gm.gx * gm.gy * gm.gz;
  const int span = (cells + 1023) / 1024 * 1024;  // <= G_MAX; entries in [cells, span) hold 0 and scan to `total`
  for (int base = 0; base < span; base += 1024) {
    const int v = c[base + tid];
    int x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int y = __shfl_up_sync(FULL, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(FULL, w, off);
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
  if (tid == 0) c[span] = carry_s;  // one-past-the-last cell when cells == span
}

// cell-ordered copies: the index (bd_grid_order) and (x, y, z, index) — queries and the bucketed FPS
// stream 16-byte records of contiguous cells instead of chasing indices into the 24-byte point rows
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

// Warp-wide bitonic sort of 32 * KPL keys held KPL per lane; element e = lane * KPL + r, ascending.
template <int KPL>
__device__ __forceinline__ void warp_sort(int (&v)[KPL], int lane) {
#pragma unroll
  for (int k2 = 2; k2 <= 32 * KPL; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      if (j >= KPL) {
        const int lj = j / KPL;
        const bool lower = (lane & lj) == 0;
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
          const bool up = ((lane * KPL + r) & k2) == 0;
          const int other = __shfl_xor_sync(FULL, v[r], lj);
          v[r] = (up == lower) ? min(v[r], other) : max(v[r], other);
        }
      } else {
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
          if (r & j) continue;
          const bool up = ((lane * KPL + r) & k2) == 0;
          const int a = v[r], b = v[r | j];
          v[r] = up ? min(a, b) : max(a, b);
          v[r | j] = up ? max(a, b) : min(a, b);
        }
      }
    }
  }
}

template <int KPL>
__device__ __forceinline__ void write_row(const int (&v)[KPL], int lane, int cnt, int nsample, int *__restrict__ row) {
  const int keep = min(cnt, nsample);
  const int first = cnt ? __shfl_sync(FULL, v[0], 0) : 0;
#pragma unroll
  for (int r = 0; r < KPL; ++r) {
    const int e = lane * KPL + r;
    if (e < keep) row[e] = v[r];
  }
  for (int s = keep + lane; s < nsample; s += 32) row[s] = first;
}

__global__ void __launch_bounds__(Q_WARPS * 32)
bq_query_kernel(const float *__restrict__ new_xyz, int n, int m, float radius, float radius2, int nsample,
                const GridMeta *__restrict__ meta, const int *__restrict__ start, const float4 *__restrict__ sorted,
                int *__restrict__ idx) {
  __shared__ int hits[Q_WARPS][Q_BUF];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * Q_WARPS + warp;
  if (j >= m) return;
  const GridMeta g = meta[b];
  start += static_cast<long long>(b) * (G_MAX + 1);
  sorted += static_cast<long long>(b) * n;
  const float *c = new_xyz + (static_cast<long long>(b) * m + j) * 3;
  const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
  int *row = idx + (static_cast<long long>(b) * m + j) * nsample;
  int *h = hits[warp];
  int v[4] = {I_MAX, I_MAX, I_MAX, I_MAX};  // sorted file of the 128 smallest hits so far (valid once `full`)
  bool full = false;
  int cnt = 0;           // entries in the buffer
  int thresh = I_MAX;    // once the file is full: its nsample-th smallest index; larger hits cannot matter
  // centre cell, NOT clamped (a centre outside the cloud's box simply sees fewer cells); cells are
  // at least `radius` wide for grids built with this radius, `reach` covers any other cell size
  const int ix = cell_coord(cx, g.minx, g.inv_cell), iy = cell_coord(cy, g.miny, g.inv_cell),
            iz = cell_coord(cz, g.minz, g.inv_cell);
  const int reach = max(1, static_cast<int>(ceilf(radius * g.inv_cell * 1.0001f)));
  const int z0 = max(iz - reach, 0), z1 = min(iz + reach, g.gz - 1);
  const int y0 = max(iy - reach, 0), y1 = min(iy + reach, g.gy - 1);
  const int x0 = max(ix - reach, 0), x1 = min(ix + reach, g.gx - 1);
  // The (z, y) rows of neighbouring cells are contiguous ranges of `sorted` (x-neighbours are adjacent cells).
  // All range bounds are fetched at once (one lane per row: one memory latency instead of one per row) and
  // handed out by shuffles.  (Measured and dropped: concatenating the ranges with a warp scan and walking them
  // 64 candidates at a time — the 36 locating shuffles per step cost more than the round trips they save:
  // 607 vs 479 us at 148 scenes; the kernel is issue-bound.)
  auto consume = [&](bool valid, const float4 &q) {
    int k = -1;
    bool hit = false;
    if (valid) {
      k = __float_as_int(q.w);
      hit = bd::sqdist_ref(cx, cy, cz, q.x, q.y, q.z) < radius2 && k < thresh;
    }
    const unsigned ballot = __ballot_sync(FULL, hit);
    if (!ballot) return;
    if (hit) h[cnt + __popc(ballot & ((1u << lane) - 1u))] = k;
    cnt += __popc(ballot);
    const int due = full ? 64 : 128;
    if (cnt >= due) {  // merge `due` buffered hits into the (upper part of the) sorted file
      __syncwarp();
      const int keep = 128 - due;  // file entries that stay: the 64 smallest, or none the first time
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int e = lane * 4 + r;
        if (e >= keep) v[r] = h[e - keep];
      }
      warp_sort<4>(v, lane);
      full = true;
      const int r_sel = (nsample - 1) & 3;
      const int mine = r_sel == 0 ? v[0] : (r_sel == 1 ? v[1] : (r_sel == 2 ? v[2] : v[3]));
      thresh = __shfl_sync(FULL, mine, (nsample - 1) >> 2);
      const int left = cnt - due;  // < 32
      const int moved = lane < left ? h[due + lane] : 0;
      __syncwarp();
      if (lane < left) h[lane] = moved;
      cnt = left;
    }
  };
  const int nyr = y1 - y0 + 1, nrows = (x0 <= x1 && y0 <= y1 && z0 <= z1) ? nyr * (z1 - z0 + 1) : 0;
  for (int rb = 0; rb < nrows; rb += 32) {  // 32 rows per pass (cells >= radius: a single pass of <= 9 rows)
    const int rr = rb + lane;
    int rs = 0, re = 0;
    if (rr < nrows) {
      const int z = z0 + rr / nyr, y = y0 + rr % nyr;
      const int cell0 = (z * g.gy + y) * g.gx + x0;
      rs = __ldg(start + cell0);
      re = __ldg(start + cell0 + (x1 - x0) + 1);
    }
    const int nr = min(32, nrows - rb);
    for (int r = 0; r < nr; ++r) {
      const int s0 = __shfl_sync(FULL, rs, r), s1 = __shfl_sync(FULL, re, r);
      for (int t0 = s0; t0 < s1; t0 += 32) {
        const int t = t0 + lane;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < s1) q = __ldg(sorted + t);
        consume(t < s1, q);
      }
    }
  }
  __syncwarp();
  if (full) {
    if (cnt) {  // < 64 leftovers replace the upper half
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int e = lane * 4 + r;
        if (e >= 64) v[r] = (e - 64 < cnt) ? h[e - 64] : I_MAX;
      }
      warp_sort<4>(v, lane);
    }
    write_row<4>(v, lane, 128, nsample, row);
  } else if (cnt <= 32) {
    int v1[1] = {lane < cnt ? h[lane] : I_MAX};
    warp_sort<1>(v1, lane);
    write_row<1>(v1, lane, cnt, nsample, row);
  } else if (cnt <= 64) {
    int v2[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) v2[r] = (lane * 2 + r < cnt) ? h[lane * 2 + r] : I_MAX;
    warp_sort<2>(v2, lane);
    write_row<2>(v2, lane, cnt, nsample, row);
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = (lane * 4 + r < cnt) ? h[lane * 4 + r] : I_MAX;
    warp_sort<4>(v, lane);
    write_row<4>(v, lane, cnt, nsample, row);
  }
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

const float4 *bd::grid_sorted_points(void *grid_workspace, int B, int n) { return grid_ws(grid_workspace, B, n).sorted_pts; }

// Cell list of B clouds of n points: the build half of bd_ball_query_grid.  radius > 0: cells of edge
// >= radius (what bd_ball_query_grid_query with that radius wants); radius <= 0: cells of 1/32 of the
// cloud's largest extent (a spatial order for bd_fps_grid; queries still work, they visit more cells).
// Afterwards bd_grid_order() is the points' indices grouped by cell (x-fastest cell order).
extern "C" int bd_grid_build(const float *xyz, int ld_xyz, int B, int n, float radius, void *workspace,
                             bd_stream_t stream) {
  BD_REQUIRE(xyz && workspace, "bd_grid_build: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && ld_xyz >= 3, "bd_grid_build: bad sizes");
  cudaStream_t s = bd::as_stream(stream);
  const GridWs g = grid_ws(workspace, B, n);
  BD_CUDA(cudaMemsetAsync(g.count, 0, sizeof(int) * static_cast<size_t>(B) * (G_MAX + 1), s), "bd_grid_build");
  bq_bbox_init_kernel<<<bd::ceil_div(B * 8, 256), 256, 0, s>>>(g.bbox, B);
  bq_bbox_kernel<<<dim3(BB_PARTS, B), 256, 0, s>>>(xyz, ld_xyz, n, g.bbox);
  bq_meta_kernel<<<bd::ceil_div(B, 128), 128, 0, s>>>(g.bbox, radius, B, g.meta);
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

// Query half: `workspace` holds the cell list built by bd_grid_build over the SAME xyz (any cell size;
// cells of edge >= radius are the efficient case).  nsample <= 64.
extern "C" int bd_ball_query_grid_query(const float *new_xyz, const float *xyz, int ld_xyz, int B, int n, int m,
                                        float radius, int nsample, int *idx, void *workspace, bd_stream_t stream) {
  BD_REQUIRE(new_xyz && xyz && idx && workspace, "bd_ball_query_grid_query: null pointer");
  BD_REQUIRE(B > 0 && B <= 65535 && n > 0 && m > 0 && nsample > 0 && ld_xyz >= 3 && radius > 0.f,
             "bd_ball_query_grid_query: bad sizes");
  if (nsample > Q_NS_MAX)  // the register selection keeps 64 candidates: larger groups take the ordered scan
    return bd_ball_query(new_xyz, xyz, ld_xyz, B, n, m, radius, nsample, idx, stream);
  const GridWs g = grid_ws(workspace, B, n);
  dim3 qgrid(bd::ceil_div(m, Q_WARPS), B);
  bq_query_kernel<<<qgrid, Q_WARPS * 32, 0, bd::as_stream(stream)>>>(new_xyz, n, m, radius, radius * radius, nsample,
                                                                      g.meta, g.count, g.sorted_pts, idx);
  BD_CHECK_LAUNCH("bd_ball_query_grid_query");
  return BD_OK;
}

extern "C" int bd_ball_query_grid(const float *new_xyz, const float *xyz, int ld_xyz, int B, int n, int m, float radius,
                                  int nsample, int *idx, void *workspace, bd_stream_t stream) {
  BD_REQUIRE(new_xyz && xyz && idx && workspace, "bd_ball_query_grid: null pointer");
  BD_REQUIRE(B > 0 && B <= 65535 && n > 0 && m > 0 && nsample > 0 && ld_xyz >= 3 && radius > 0.f,
             "bd_ball_query_grid: bad sizes");
  if (nsample > Q_NS_MAX) return bd_ball_query(new_xyz, xyz, ld_xyz, B, n, m, radius, nsample, idx, stream);
  const int rc = bd_grid_build(xyz, ld_xyz, B, n, radius, workspace, stream);
  if (rc != BD_OK) return rc;
  return bd_ball_query_grid_query(new_xyz, xyz, ld_xyz, B, n, m, radius, nsample, idx, workspace, stream);
}
