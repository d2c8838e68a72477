// Device-side Hungarian matcher, replacing
// /root/reference/models/losses.py:256-331 (`HungarianMatcher.forward`): the reference builds the cost matrix with a
// dozen torch launches, copies it to the host (a device synchronisation per prediction head, 7 per training
// step) and solves every scene with scipy.optimize.linear_sum_assignment on the CPU.
//
//   matcher_cost_kernel   one warp per (scene, query): softmax over the class logits (registers), then per target of
//                         the scene  w_class * -(prob . positive_map[t])  (or -prob[label[t]])  +  w_bbox * L1  +
//                         w_giou * -GIoU3D,  written TARGET-major (cost[t][q], q fastest) — the order the solver reads.
//                         Same fp32 formulas, in the same order, as models/losses.py:28-91.
//   hungarian_kernel      one CTA per scene, one thread per query column: shortest-augmenting-path assignment
//                         (the algorithm of scipy's rectangular_lsap: Dijkstra over the columns from each new row,
//                         dual update, augmentation), duals and path costs in fp64 like scipy.  The column scan and
//                         the arg-min of every Dijkstra step are parallel over the columns; a step costs two CTA
//                         barriers.  Output: the matched (query, target) pairs of every scene, sorted by query index —
//                         what linear_sum_assignment returns for the (Q x T) matrix.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int MC_WARPS = 8;
constexpr int MC_MAX_PER_LANE = 16;  // classes per lane: C <= 512

__device__ __forceinline__ void box_corners(const float *b, float (&c)[6]) {  // losses.py:28-39
  const float w = fmaxf(b[3], 1e-6f), h = fmaxf(b[4], 1e-6f), d = fmaxf(b[5], 1e-6f);
  c[0] = b[0] - 0.5f * w, c[1] = b[1] - 0.5f * h, c[2] = b[2] - 0.5f * d;
  c[3] = b[0] + 0.5f * w, c[4] = b[1] + 0.5f * h, c[5] = b[2] + 0.5f * d;
}

__device__ __forceinline__ float giou3d(const float (&a)[6], const float (&b)[6]) {  // losses.py:42-91
  const float xa = fmaxf(a[0], b[0]), ya = fmaxf(a[1], b[1]), za = fmaxf(a[2], b[2]);
  const float xb = fminf(a[3], b[3]), yb = fminf(a[4], b[4]), zb = fminf(a[5], b[5]);
  const float inter = __fmul_rn(__fmul_rn(fmaxf(__fsub_rn(xb, xa), 0.f), fmaxf(__fsub_rn(yb, ya), 0.f)), fmaxf(__fsub_rn(zb, za), 0.f));
  const float va = __fmul_rn(__fmul_rn(__fsub_rn(a[3], a[0]), __fsub_rn(a[4], a[1])), __fsub_rn(a[5], a[2]));
  const float vb = __fmul_rn(__fmul_rn(__fsub_rn(b[3], b[0]), __fsub_rn(b[4], b[1])), __fsub_rn(b[5], b[2]));
  const float uni = __fsub_rn(__fadd_rn(va, vb), inter);
  const float iou = inter / uni;
  const float w0 = fmaxf(__fsub_rn(fmaxf(a[3], b[3]), fminf(a[0], b[0])), 0.f);
  const float w1 = fmaxf(__fsub_rn(fmaxf(a[4], b[4]), fminf(a[1], b[1])), 0.f);
  const float w2 = fmaxf(__fsub_rn(fmaxf(a[5], b[5]), fminf(a[2], b[2])), 0.f);
  const float vol = __fmul_rn(__fmul_rn(w0, w1), w2);
  return __fsub_rn(iou, __fsub_rn(vol, uni) / vol);
}

__global__ void __launch_bounds__(MC_WARPS * 32)
matcher_cost_kernel(const float *__restrict__ logits, const float *__restrict__ boxes, const float *__restrict__ tgt_boxes,
                    const float *__restrict__ pmap, int ld_pm, const long long *__restrict__ labels,
                    const int *__restrict__ tgt_off, int Q, int C, float w_class, float w_bbox, float w_giou,
                    float *__restrict__ cost) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int q = blockIdx.x * MC_WARPS + (threadIdx.x >> 5);
  if (q >= Q) return;
  const int t0 = tgt_off[b], t1 = tgt_off[b + 1];
  if (t1 <= t0) return;
  const float *lg = logits + (static_cast<long long>(b) * Q + q) * C;
  float p[MC_MAX_PER_LANE];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < MC_MAX_PER_LANE; ++i) {
    const int c = lane + i * 32;
    p[i] = c < C ? __ldg(lg + c) : -INFINITY;
    mx = fmaxf(mx, p[i]);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MC_MAX_PER_LANE; ++i) {
    p[i] = (lane + i * 32 < C) ? expf(p[i] - mx) : 0.f;
    sum += p[i];
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
#pragma unroll
  for (int i = 0; i < MC_MAX_PER_LANE; ++i) p[i] = p[i] / sum;
  const float *pb = boxes + (static_cast<long long>(b) * Q + q) * 6;
  float pbox[6], pc[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) pbox[i] = __ldg(pb + i);
  box_corners(pbox, pc);
  float *out = cost + static_cast<long long>(t0) * Q + q;
  for (int t = t0; t < t1; ++t) {
    float dot = 0.f;
    if (labels) {  // hard labels: -prob[label]
      const long long lab = labels[t];
      const int i = static_cast<int>(lab >> 5);
#pragma unroll
      for (int k = 0; k < MC_MAX_PER_LANE; ++k)
        if (k == i && (lab & 31) == lane) dot = p[k];
    } else {
      const float *pm = pmap + static_cast<long long>(t) * ld_pm;
#pragma unroll
      for (int i = 0; i < MC_MAX_PER_LANE; ++i) {
        const int c = lane + i * 32;
        if (c < C) dot = fmaf(p[i], __ldg(pm + c), dot);
      }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xFFFFFFFFu, dot, o);
    if (lane == 0) {
      const float *tb = tgt_boxes + static_cast<long long>(t) * 6;
      float tbox[6], tcn[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) tbox[i] = __ldg(tb + i);
      box_corners(tbox, tcn);
      float l1 = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) l1 = __fadd_rn(l1, fabsf(__fsub_rn(pbox[i], tbox[i])));
      const float g = giou3d(pc, tcn);
      // C = cost_bbox * L1 + cost_class * (-prob.pm) + cost_giou * (-giou)   (losses.py:309-313, same order)
      out[static_cast<long long>(t - t0) * Q] =
          __fadd_rn(__fadd_rn(__fmul_rn(w_bbox, l1), __fmul_rn(w_class, -dot)), __fmul_rn(w_giou, -g));
    }
  }
}

struct MinKey {
  double v;
  int j;  // column; bit 30 set = column already assigned (loses ties against a free column)
};
__device__ __forceinline__ bool key_less(const MinKey &a, const MinKey &b) {
  return a.v < b.v || (a.v == b.v && a.j < b.j);
}

__global__ void __launch_bounds__(1024)
hungarian_kernel(const float *__restrict__ cost, const int *__restrict__ tgt_off, int Q, long long *__restrict__ match_q,
                 long long *__restrict__ match_t, int *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char hs_raw[];
  const int b = blockIdx.x, j = threadIdx.x, lane = j & 31, warp = j >> 5;
  const int t0 = tgt_off[b], T = tgt_off[b + 1] - t0;
  if (T <= 0) return;
  double *u = reinterpret_cast<double *>(hs_raw);       // [T] row duals
  int *col4row = reinterpret_cast<int *>(u + T);         // [T]
  int *row4col = col4row + T;                            // [Q]
  int *path = row4col + Q;                               // [Q]
  __shared__ MinKey wmin[32];
  __shared__ int s_next, s_jmin;
  __shared__ double s_low;
  const float *Cm = cost + static_cast<long long>(t0) * Q;  // [t][q]
  for (int i = j; i < T; i += blockDim.x) u[i] = 0.0, col4row[i] = -1;
  if (j < Q) row4col[j] = -1;
  double v = 0.0;  // dual of this thread's column
  __syncthreads();
  bool failed = false;
  for (int cur = 0; cur < T && !failed; ++cur) {
    double shortest = DBL_MAX;
    bool sc = false;
    if (j < Q) path[j] = -1;
    int i = cur, sink = -1;
    double minval = 0.0;
    while (sink < 0) {
      const double ui = u[i];
      MinKey k = {DBL_MAX, 0x7FFFFFFF};
      if (j < Q && !sc) {
        const double r = minval + static_cast<double>(__ldg(Cm + static_cast<long long>(i) * Q + j)) - ui - v;
        if (r < shortest) shortest = r, path[j] = i;
        k.v = shortest;
        k.j = j | (row4col[j] >= 0 ? (1 << 30) : 0);
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        MinKey other;
        other.v = __shfl_xor_sync(0xFFFFFFFFu, k.v, o);
        other.j = __shfl_xor_sync(0xFFFFFFFFu, k.j, o);
        if (key_less(other, k)) k = other;
      }
      if (lane == 0) wmin[warp] = k;
      __syncthreads();
      if (warp == 0) {
        MinKey m = lane < (blockDim.x >> 5) ? wmin[lane] : MinKey{DBL_MAX, 0x7FFFFFFF};
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          MinKey other;
          other.v = __shfl_xor_sync(0xFFFFFFFFu, m.v, o);
          other.j = __shfl_xor_sync(0xFFFFFFFFu, m.j, o);
          if (key_less(other, m)) m = other;
        }
        if (lane == 0) {
          s_low = m.v;
          s_jmin = m.j & ~(1 << 30);
          s_next = m.v == DBL_MAX ? -2 : row4col[m.j & ~(1 << 30)];
        }
      }
      __syncthreads();
      if (s_next == -2) {  // no finite entry left (inf / NaN costs): scipy raises; here the scene is flagged
        failed = true;
        break;
      }
      minval = s_low;
      const int jmin = s_jmin;
      if (j == jmin) sc = true;
      if (s_next < 0) sink = jmin;
      else i = s_next;
      __syncthreads();  // s_* are rewritten by the next step
    }
    if (failed) break;
    // dual update (rectangular_lsap: u[cur] += minval; u[i] += minval - shortest[col4row[i]] for the other
    // visited rows; v[j] -= minval - shortest[j] for the visited columns): a visited, assigned column owns the
    // update of its row
    if (j == 0) u[cur] += minval;
    if (j < Q && sc) {
      const double d = minval - shortest;
      const int r = row4col[j];
      if (r >= 0) u[r] += d;
      v -= d;
    }
    __syncthreads();
    if (j == 0) {  // augment along the path
      int c = sink;
      while (true) {
        const int r = path[c];
        row4col[c] = r;
        const int prev = col4row[r];
        col4row[r] = c;
        c = prev;
        if (r == cur) break;
      }
    }
    __syncthreads();
  }
  if (failed) {
    if (j == 0 && status) atomicExch(status, 1);
    for (int t = j; t < T; t += blockDim.x) match_q[t0 + t] = -1, match_t[t0 + t] = -1;
    return;
  }
  // pairs sorted by query index (linear_sum_assignment's row_ind is ascending)
  for (int t = j; t < T; t += blockDim.x) {
    const int mine = col4row[t];
    int rank = 0;
    for (int k = 0; k < T; ++k) rank += col4row[k] < mine;
    match_q[t0 + rank] = mine;
    match_t[t0 + rank] = t;
  }
}

}  // namespace

extern "C" int bd_matcher_cost(const float *logits, const float *boxes, const float *tgt_boxes, const float *positive_map,
                               int ld_pm, const long long *labels, const int *tgt_offset, int B, int Q, int C,
                               float w_class, float w_bbox, float w_giou, float *cost, bd_stream_t stream) {
  BD_REQUIRE(logits && boxes && tgt_boxes && tgt_offset && cost && (positive_map || labels), "bd_matcher_cost: null pointer");
  BD_REQUIRE(B > 0 && B <= 65535 && Q > 0 && C > 0 && C <= 32 * MC_MAX_PER_LANE && (labels || ld_pm >= C),
             "bd_matcher_cost: bad sizes (C <= 512, ld_pm >= C)");
  matcher_cost_kernel<<<dim3(bd::ceil_div(Q, MC_WARPS), B), MC_WARPS * 32, 0, bd::as_stream(stream)>>>(
      logits, boxes, tgt_boxes, positive_map, ld_pm, labels, tgt_offset, Q, C, w_class, w_bbox, w_giou, cost);
  BD_CHECK_LAUNCH("bd_matcher_cost");
  return BD_OK;
}

extern "C" int bd_hungarian(const float *cost, const int *tgt_offset, int B, int Q, int max_targets, long long *match_q,
                            long long *match_t, int *status, bd_stream_t stream) {
  BD_REQUIRE(cost && tgt_offset && match_q && match_t, "bd_hungarian: null pointer");
  BD_REQUIRE(B > 0 && Q > 0 && Q <= 1024 && max_targets >= 0 && max_targets <= Q,
             "bd_hungarian: needs targets per scene <= queries <= 1024");
  if (max_targets == 0) return BD_OK;
  const int threads = ((Q + 31) / 32) * 32;
  const size_t smem = static_cast<size_t>(max_targets) * (sizeof(double) + sizeof(int)) + 2 * static_cast<size_t>(Q) * sizeof(int) + 16;
  hungarian_kernel<<<B, threads, smem, bd::as_stream(stream)>>>(cost, tgt_offset, Q, match_q, match_t, status);
  BD_CHECK_LAUNCH("bd_hungarian");
  return BD_OK;
}
