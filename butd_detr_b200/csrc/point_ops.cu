// Irregular point operators for sm_100a: ball query, gathers/grouping, 3-NN, interpolation
// (forward + the three scatter-add backward ops).  They replace the kernels of
// /root/reference/pointnet2/_ext_src/src/{ball_query,group_points,interpolate,sampling}_gpu.cu.
//
// The reference launches ONE CTA per batch element for every one of these (grid = B), i.e. a
// single SM of 148 at B = 1.  Here every op is decomposed over (scene, centre/row) so the
// grid covers the machine, global reads are coalesced (token-major rows) and the candidate
// points of a ball query are staged through shared memory once per CTA.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

// --------------------------------------------------------------------------- ball query
// One warp per query centre; the CTA's 8 warps share point tiles staged in shared memory
// (SoA, conflict-free).  32 candidates are tested per step; __ballot_sync + popc keep the
// reference's "first nsample hits in ascending index order" semantics exactly
// (ball_query_gpu.cu:32-45).  d2 uses the reference's FMUL/FFMA/FFMA contraction.
constexpr int BQ_WARPS = 8;
constexpr int BQ_TILE = 2048;

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int ld, int n, int m,
                  float radius2, int nsample, int *__restrict__ idx) {
  __shared__ float sx[BQ_TILE], sy[BQ_TILE], sz[BQ_TILE];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * BQ_WARPS + warp;
  const bool active = j < m;
  xyz += static_cast<long long>(b) * n * ld;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  int *row = nullptr;
  if (active) {
    const float *c = new_xyz + (static_cast<long long>(b) * m + j) * 3;
    cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    row = idx + (static_cast<long long>(b) * m + j) * nsample;
  }
  int cnt = 0, first = 0;
  bool done = !active;
  for (int base = 0; base < n; base += BQ_TILE) {
    const int tile = min(BQ_TILE, n - base);
    __syncthreads();
    for (int i = threadIdx.x; i < tile; i += BQ_WARPS * 32) {
      const float *p = xyz + static_cast<long long>(base + i) * ld;
      sx[i] = __ldg(p), sy[i] = __ldg(p + 1), sz[i] = __ldg(p + 2);
    }
    __syncthreads();
    if (!done) {
      for (int i0 = 0; i0 < tile; i0 += 32) {
        const int i = i0 + lane;
        bool hit = false;
        if (i < tile) hit = bd::sqdist_ref(cx, cy, cz, sx[i], sy[i], sz[i]) < radius2;
        const unsigned ballot = __ballot_sync(0xFFFFFFFFu, hit);
        if (ballot) {
          if (cnt == 0) first = base + i0 + __ffs(ballot) - 1;
          const int pos = cnt + __popc(ballot & ((1u << lane) - 1u));
          if (hit && pos < nsample) row[pos] = base + i;
          cnt += __popc(ballot);
          if (cnt >= nsample) { done = true; break; }
        }
      }
    }
    if (__syncthreads_and(done)) break;
  }
  if (active) {
    // slots past the hit count repeat the first hit; an empty ball stays all-zero
    // (torch::zeros in ball_query.cpp:24-26)
    for (int s = min(cnt, nsample) + lane; s < nsample; s += 32) row[s] = first;
  }
}

// --------------------------------------------------------------------------- channel-major ops (Part A)
__global__ void gather_points_kernel(const float *__restrict__ points, const int *__restrict__ idx, int C, int N,
                                     int m, float *__restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int a = __ldg(idx + static_cast<long long>(b) * m + j);
  out[(static_cast<long long>(b) * C + c) * m + j] = __ldg(points + (static_cast<long long>(b) * C + c) * N + a);
}

__global__ void gather_points_grad_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, int C,
                                          int N, int m, float *__restrict__ grad_points) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int a = __ldg(idx + static_cast<long long>(b) * m + j);
  atomicAdd(grad_points + (static_cast<long long>(b) * C + c) * N + a,
            __ldg(grad_out + (static_cast<long long>(b) * C + c) * m + j));
}

__global__ void group_points_kernel(const float *__restrict__ points, const int *__restrict__ idx, int C, int n,
                                    int mns, float *__restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mns) return;
  const int a = __ldg(idx + static_cast<long long>(b) * mns + i);
  out[(static_cast<long long>(b) * C + c) * mns + i] = __ldg(points + (static_cast<long long>(b) * C + c) * n + a);
}

__global__ void group_points_grad_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, int C,
                                         int n, int mns, float *__restrict__ grad_points) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mns) return;
  const int a = __ldg(idx + static_cast<long long>(b) * mns + i);
  atomicAdd(grad_points + (static_cast<long long>(b) * C + c) * n + a,
            __ldg(grad_out + (static_cast<long long>(b) * C + c) * mns + i));
}

// three_nn: one thread per unknown point, known points staged in shared memory.  The
// reference keeps its running bests in double initialised to 1e40 and compares float d
// against them (interpolate_gpu.cu:32-55); with float bests initialised to +inf every
// comparison has the same outcome and (float)1e40 == +inf, so results are identical.
constexpr int NN_TILE = 1024;
__global__ void __launch_bounds__(128)
three_nn_kernel(const float *__restrict__ unknown, const float *__restrict__ known, int n, int m,
                float *__restrict__ dist2, int *__restrict__ idx) {
  __shared__ float kx[NN_TILE], ky[NN_TILE], kz[NN_TILE];
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  known += static_cast<long long>(b) * m * 3;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (j < n) {
    const float *u = unknown + (static_cast<long long>(b) * n + j) * 3;
    ux = u[0], uy = u[1], uz = u[2];
  }
  const float inf = __int_as_float(0x7f800000);
  float b1 = inf, b2 = inf, b3 = inf;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int base = 0; base < m; base += NN_TILE) {
    const int tile = min(NN_TILE, m - base);
    __syncthreads();
    for (int i = threadIdx.x; i < tile; i += blockDim.x) {
      kx[i] = known[(base + i) * 3], ky[i] = known[(base + i) * 3 + 1], kz[i] = known[(base + i) * 3 + 2];
    }
    __syncthreads();
    for (int i = 0; i < tile; ++i) {
      const float d = bd::sqdist_ref(ux, uy, uz, kx[i], ky[i], kz[i]);
      const int k = base + i;
      if (d < b1) {
        b3 = b2, i3 = i2, b2 = b1, i2 = i1, b1 = d, i1 = k;
      } else if (d < b2) {
        b3 = b2, i3 = i2, b2 = d, i2 = k;
      } else if (d < b3) {
        b3 = d, i3 = k;
      }
    }
  }
  if (j < n) {
    float *od = dist2 + (static_cast<long long>(b) * n + j) * 3;
    int *oi = idx + (static_cast<long long>(b) * n + j) * 3;
    od[0] = b1, od[1] = b2, od[2] = b3;
    oi[0] = i1, oi[1] = i2, oi[2] = i3;
  }
}

__global__ void three_interpolate_kernel(const float *__restrict__ points, const int *__restrict__ idx,
                                         const float *__restrict__ weight, int C, int m, int n,
                                         float *__restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int *ii = idx + (static_cast<long long>(b) * n + j) * 3;
  const float *w = weight + (static_cast<long long>(b) * n + j) * 3;
  const float *src = points + (static_cast<long long>(b) * C + c) * m;
  // FMUL(p2,w2), FFMA(p1,w1), FFMA(p3,w3) as in the reference's compiled kernel (interpolate_gpu.cu:103-105)
  out[(static_cast<long long>(b) * C + c) * n + j] =
      __fmaf_rn(__ldg(src + ii[2]), w[2], __fmaf_rn(__ldg(src + ii[0]), w[0], __fmul_rn(__ldg(src + ii[1]), w[1])));
}

__global__ void three_interpolate_grad_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx,
                                              const float *__restrict__ weight, int C, int n, int m,
                                              float *__restrict__ grad_points) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int *ii = idx + (static_cast<long long>(b) * n + j) * 3;
  const float *w = weight + (static_cast<long long>(b) * n + j) * 3;
  const float g = __ldg(grad_out + (static_cast<long long>(b) * C + c) * n + j);
  float *dst = grad_points + (static_cast<long long>(b) * C + c) * m;
  atomicAdd(dst + ii[0], g * w[0]);
  atomicAdd(dst + ii[1], g * w[1]);
  atomicAdd(dst + ii[2], g * w[2]);
}

// --------------------------------------------------------------------------- token-major ops (Part B)
__global__ void gather_rows_kernel(const float *__restrict__ src, int ld_src, const int *__restrict__ idx, int n_src,
                                   int m, int w, float *__restrict__ out, int ld_out, long long total) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c = static_cast<int>(e % w);
  const long long r = e / w;  // b * m + j
  const long long b = r / m;
  const int a = __ldg(idx + r);
  out[r * ld_out + c] = __ldg(src + (b * n_src + a) * ld_src + c);
}

__global__ void group_rows_kernel(const float *__restrict__ xyz, int ld_xyz, const float *__restrict__ feats,
                                  int ld_feats, int C, const float *__restrict__ new_xyz,
                                  const int *__restrict__ idx, int n, int m, int ns, float inv_radius,
                                  float *__restrict__ out, int ld_out, long long total) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c = static_cast<int>(e % ld_out);
  const long long row = e / ld_out;  // (b*m + j)*ns + s
  if (c >= 3 + C) {  // zero padding up to the leading dimension (keeps K a multiple of 4)
    out[e] = 0.f;
    return;
  }
  const long long bj = row / ns;
  const long long b = bj / m;
  const int a = __ldg(idx + row);
  float v;
  if (c < 3) {
    // (x - centre) * (1/r): ATen's CUDA in-place `/= radius` with a host scalar multiplies by
    // the reciprocal (pointnet2_utils.py:350-352)
    v = __fmul_rn(__fsub_rn(__ldg(xyz + (b * n + a) * ld_xyz + c), __ldg(new_xyz + bj * 3 + c)), inv_radius);
  } else {
    v = __ldg(feats + (b * n + a) * ld_feats + (c - 3));
  }
  out[e] = v;
}

__global__ void maxpool_rows_kernel(const float *__restrict__ in, int ns, int C, float *__restrict__ out,
                                    long long total) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c = static_cast<int>(e % C);
  const long long r = e / C;
  const float *p = in + r * ns * C + c;
  float v = __ldg(p);
  for (int s = 1; s < ns; ++s) v = fmaxf(v, __ldg(p + static_cast<long long>(s) * C));
  out[e] = v;
}

__global__ void fp_interp_concat_kernel(const float *__restrict__ dist2, const int *__restrict__ idx,
                                        const float *__restrict__ known_feats, int C2,
                                        const float *__restrict__ unknown_feats, int C1, int n, int m,
                                        float *__restrict__ out, long long total) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int W = C1 + C2;
  const int c = static_cast<int>(e % W);
  const long long r = e / W;  // b*n + j
  const long long b = r / n;
  if (c >= C2) {
    out[e] = __ldg(unknown_feats + r * C1 + (c - C2));
    return;
  }
  // pointnet2_modules.py:394-397 (torch op order) ; sqrt/div are IEEE (no fast-math here)
  const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + r * 3 + 0)), 1e-8f));
  const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + r * 3 + 1)), 1e-8f));
  const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + r * 3 + 2)), 1e-8f));
  const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
  const float w0 = __fdiv_rn(r0, norm), w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm);
  const float *kf = known_feats + b * m * C2 + c;
  const float p0 = __ldg(kf + static_cast<long long>(__ldg(idx + r * 3 + 0)) * C2);
  const float p1 = __ldg(kf + static_cast<long long>(__ldg(idx + r * 3 + 1)) * C2);
  const float p2 = __ldg(kf + static_cast<long long>(__ldg(idx + r * 3 + 2)) * C2);
  out[e] = __fmaf_rn(p2, w2, __fmaf_rn(p0, w0, __fmul_rn(p1, w1)));
}

// Same operator, one WARP per output row (C1 % 4 == 0, C2 % 4 == 0): the three interpolation weights are computed
// once per row (the element-per-thread kernel above recomputes 3 square roots and 4 divisions for each of the
// C1 + C2 elements), features move as 16-byte vectors, and the row can be written as fp16 (HALF: the operand
// format of the linear layer that reads it next; same fp32 value, rounded once).
template <bool HALF>
__global__ void __launch_bounds__(256)
fp_interp_concat_rows_kernel(const float *__restrict__ dist2, const int *__restrict__ idx,
                             const float *__restrict__ known_feats, int C2, const float *__restrict__ unknown_feats, int C1,
                             int n, int m, void *__restrict__ out, long long rows) {
  const long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const long long b = r / n;
  const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + r * 3 + 0)), 1e-8f));
  const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + r * 3 + 1)), 1e-8f));
  const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + r * 3 + 2)), 1e-8f));
  const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
  const float w0 = __fdiv_rn(r0, norm), w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm);
  const float *kf = known_feats + b * m * C2;
  const float4 *k0 = reinterpret_cast<const float4 *>(kf + static_cast<long long>(__ldg(idx + r * 3 + 0)) * C2);
  const float4 *k1 = reinterpret_cast<const float4 *>(kf + static_cast<long long>(__ldg(idx + r * 3 + 1)) * C2);
  const float4 *k2 = reinterpret_cast<const float4 *>(kf + static_cast<long long>(__ldg(idx + r * 3 + 2)) * C2);
  const float4 *uf = reinterpret_cast<const float4 *>(unknown_feats + r * C1);
  const int W = C1 + C2;
  for (int v = lane; v < W / 4; v += 32) {
    float4 o;
    if (v < C2 / 4) {
      const float4 p0 = __ldg(k0 + v), p1 = __ldg(k1 + v), p2 = __ldg(k2 + v);
      o.x = __fmaf_rn(p2.x, w2, __fmaf_rn(p0.x, w0, __fmul_rn(p1.x, w1)));
      o.y = __fmaf_rn(p2.y, w2, __fmaf_rn(p0.y, w0, __fmul_rn(p1.y, w1)));
      o.z = __fmaf_rn(p2.z, w2, __fmaf_rn(p0.z, w0, __fmul_rn(p1.z, w1)));
      o.w = __fmaf_rn(p2.w, w2, __fmaf_rn(p0.w, w0, __fmul_rn(p1.w, w1)));
    } else {
      o = __ldg(uf + (v - C2 / 4));
    }
    if (HALF) {
      const __half2 h0 = __floats2half2_rn(fminf(fmaxf(o.x, -65504.f), 65504.f), fminf(fmaxf(o.y, -65504.f), 65504.f));
      const __half2 h1 = __floats2half2_rn(fminf(fmaxf(o.z, -65504.f), 65504.f), fminf(fmaxf(o.w, -65504.f), 65504.f));
      reinterpret_cast<uint2 *>(static_cast<__half *>(out) + r * W)[v] =
          make_uint2(*reinterpret_cast<const unsigned *>(&h0), *reinterpret_cast<const unsigned *>(&h1));
    } else {
      reinterpret_cast<float4 *>(static_cast<float *>(out) + r * W)[v] = o;
    }
  }
}

__global__ void transpose_rows_kernel(const float *__restrict__ in, int n, int C, float *__restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  in += static_cast<long long>(b) * n * C;
  out += static_cast<long long>(b) * n * C;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = j0 + r, c = c0 + threadIdx.x;
    if (j < n && c < C) tile[r][threadIdx.x] = in[static_cast<long long>(j) * C + c];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, j = j0 + threadIdx.x;
    if (j < n && c < C) out[static_cast<long long>(c) * n + j] = tile[threadIdx.x][r];
  }
}

inline unsigned grid1d(long long total, int block) { return static_cast<unsigned>((total + block - 1) / block); }

}  // namespace

extern "C" {

int bd_ball_query(const float *new_xyz, const float *xyz, int ld_xyz, int B, int n, int m, float radius, int nsample,
                  int *idx, bd_stream_t stream) {
  BD_REQUIRE(new_xyz && xyz && idx, "bd_ball_query: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && m > 0 && nsample > 0 && ld_xyz >= 3, "bd_ball_query: bad sizes");
  BD_REQUIRE(B <= 65535, "bd_ball_query: B too large");
  const float radius2 = radius * radius;  // ball_query_gpu.cu:25 (float product)
  dim3 grid(bd::ceil_div(m, BQ_WARPS), B);
  ball_query_kernel<<<grid, BQ_WARPS * 32, 0, bd::as_stream(stream)>>>(new_xyz, xyz, ld_xyz, n, m, radius2, nsample,
                                                                      idx);
  BD_CHECK_LAUNCH("bd_ball_query");
  return BD_OK;
}

int bd_gather_points(const float *points, const int *idx, int B, int C, int N, int m, float *out,
                     bd_stream_t stream) {
  BD_REQUIRE(points && idx && out, "bd_gather_points: null pointer");
  BD_REQUIRE(B > 0 && C > 0 && N > 0 && m > 0 && C <= 65535 && B <= 65535, "bd_gather_points: bad sizes");
  dim3 grid(bd::ceil_div(m, 128), C, B);
  gather_points_kernel<<<grid, 128, 0, bd::as_stream(stream)>>>(points, idx, C, N, m, out);
  BD_CHECK_LAUNCH("bd_gather_points");
  return BD_OK;
}

int bd_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int m, float *grad_points,
                          bd_stream_t stream) {
  BD_REQUIRE(grad_out && idx && grad_points, "bd_gather_points_grad: null pointer");
  BD_REQUIRE(B > 0 && C > 0 && N > 0 && m > 0 && C <= 65535 && B <= 65535, "bd_gather_points_grad: bad sizes");
  BD_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * static_cast<size_t>(B) * C * N, bd::as_stream(stream)),
          "bd_gather_points_grad");
  dim3 grid(bd::ceil_div(m, 128), C, B);
  gather_points_grad_kernel<<<grid, 128, 0, bd::as_stream(stream)>>>(grad_out, idx, C, N, m, grad_points);
  BD_CHECK_LAUNCH("bd_gather_points_grad");
  return BD_OK;
}

int bd_group_points(const float *points, const int *idx, int B, int C, int n, int m, int ns, float *out,
                    bd_stream_t stream) {
  BD_REQUIRE(points && idx && out, "bd_group_points: null pointer");
  BD_REQUIRE(B > 0 && C > 0 && n > 0 && m > 0 && ns > 0 && C <= 65535 && B <= 65535, "bd_group_points: bad sizes");
  dim3 grid(bd::ceil_div(m * ns, 256), C, B);
  group_points_kernel<<<grid, 256, 0, bd::as_stream(stream)>>>(points, idx, C, n, m * ns, out);
  BD_CHECK_LAUNCH("bd_group_points");
  return BD_OK;
}

int bd_group_points_grad(const float *grad_out, const int *idx, int B, int C, int n, int m, int ns,
                         float *grad_points, bd_stream_t stream) {
  BD_REQUIRE(grad_out && idx && grad_points, "bd_group_points_grad: null pointer");
  BD_REQUIRE(B > 0 && C > 0 && n > 0 && m > 0 && ns > 0 && C <= 65535 && B <= 65535,
             "bd_group_points_grad: bad sizes");
  BD_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * static_cast<size_t>(B) * C * n, bd::as_stream(stream)),
          "bd_group_points_grad");
  dim3 grid(bd::ceil_div(m * ns, 256), C, B);
  group_points_grad_kernel<<<grid, 256, 0, bd::as_stream(stream)>>>(grad_out, idx, C, n, m * ns, grad_points);
  BD_CHECK_LAUNCH("bd_group_points_grad");
  return BD_OK;
}

int bd_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2, int *idx,
                bd_stream_t stream) {
  BD_REQUIRE(unknown && known && dist2 && idx, "bd_three_nn: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && m > 0 && B <= 65535, "bd_three_nn: bad sizes");
  dim3 grid(bd::ceil_div(n, 128), B);
  three_nn_kernel<<<grid, 128, 0, bd::as_stream(stream)>>>(unknown, known, n, m, dist2, idx);
  BD_CHECK_LAUNCH("bd_three_nn");
  return BD_OK;
}

int bd_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C, int m, int n,
                         float *out, bd_stream_t stream) {
  BD_REQUIRE(points && idx && weight && out, "bd_three_interpolate: null pointer");
  BD_REQUIRE(B > 0 && C > 0 && n > 0 && m > 0 && C <= 65535 && B <= 65535, "bd_three_interpolate: bad sizes");
  dim3 grid(bd::ceil_div(n, 128), C, B);
  three_interpolate_kernel<<<grid, 128, 0, bd::as_stream(stream)>>>(points, idx, weight, C, m, n, out);
  BD_CHECK_LAUNCH("bd_three_interpolate");
  return BD_OK;
}

int bd_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B, int C, int n, int m,
                              float *grad_points, bd_stream_t stream) {
  BD_REQUIRE(grad_out && idx && weight && grad_points, "bd_three_interpolate_grad: null pointer");
  BD_REQUIRE(B > 0 && C > 0 && n > 0 && m > 0 && C <= 65535 && B <= 65535, "bd_three_interpolate_grad: bad sizes");
  BD_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * static_cast<size_t>(B) * C * m, bd::as_stream(stream)),
          "bd_three_interpolate_grad");
  dim3 grid(bd::ceil_div(n, 128), C, B);
  three_interpolate_grad_kernel<<<grid, 128, 0, bd::as_stream(stream)>>>(grad_out, idx, weight, C, n, m, grad_points);
  BD_CHECK_LAUNCH("bd_three_interpolate_grad");
  return BD_OK;
}

int bd_gather_rows(const float *src, int ld_src, const int *idx, int B, int n_src, int m, int w, float *out,
                   int ld_out, bd_stream_t stream) {
  BD_REQUIRE(src && idx && out, "bd_gather_rows: null pointer");
  BD_REQUIRE(B > 0 && n_src > 0 && m > 0 && w > 0 && ld_src >= w && ld_out >= w, "bd_gather_rows: bad sizes");
  const long long total = static_cast<long long>(B) * m * w;
  gather_rows_kernel<<<grid1d(total, 256), 256, 0, bd::as_stream(stream)>>>(src, ld_src, idx, n_src, m, w, out, ld_out,
                                                                           total);
  BD_CHECK_LAUNCH("bd_gather_rows");
  return BD_OK;
}

int bd_group_rows(const float *xyz, int ld_xyz, const float *feats, int ld_feats, int C, const float *new_xyz,
                  const int *idx, int B, int n, int m, int ns, float radius, float *out, int ld_out,
                  bd_stream_t stream) {
  BD_REQUIRE(xyz && new_xyz && idx && out && (feats || C == 0), "bd_group_rows: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && m > 0 && ns > 0 && C >= 0 && ld_xyz >= 3 && ld_feats >= C && ld_out >= 3 + C,
             "bd_group_rows: bad sizes");
  const long long total = static_cast<long long>(B) * m * ns * ld_out;
  group_rows_kernel<<<grid1d(total, 256), 256, 0, bd::as_stream(stream)>>>(xyz, ld_xyz, feats, ld_feats, C, new_xyz, idx,
                                                                          n, m, ns, 1.0f / radius, out, ld_out, total);
  BD_CHECK_LAUNCH("bd_group_rows");
  return BD_OK;
}

int bd_maxpool_rows(const float *in, int rows_out, int ns, int C, float *out, bd_stream_t stream) {
  BD_REQUIRE(in && out, "bd_maxpool_rows: null pointer");
  BD_REQUIRE(rows_out > 0 && ns > 0 && C > 0, "bd_maxpool_rows: bad sizes");
  const long long total = static_cast<long long>(rows_out) * C;
  maxpool_rows_kernel<<<grid1d(total, 256), 256, 0, bd::as_stream(stream)>>>(in, ns, C, out, total);
  BD_CHECK_LAUNCH("bd_maxpool_rows");
  return BD_OK;
}

static int fp_interp_concat_impl(const float *dist2, const int *idx, const float *known_feats, int C2,
                                 const float *unknown_feats, int C1, int B, int n, int m, void *out, int out_half,
                                 bd_stream_t stream) {
  BD_REQUIRE(dist2 && idx && known_feats && out && (unknown_feats || C1 == 0), "bd_fp_interp_concat: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && m > 0 && C2 > 0 && C1 >= 0, "bd_fp_interp_concat: bad sizes");
  const bool vec = C1 % 4 == 0 && C2 % 4 == 0 && (reinterpret_cast<uintptr_t>(known_feats) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(unknown_feats) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  BD_REQUIRE(!out_half || vec, "bd_fp_interp_concat_h: fp16 rows need C1 % 4 == 0, C2 % 4 == 0 and 16-byte aligned tensors");
  if (vec) {
    const long long rows = static_cast<long long>(B) * n;
    if (out_half)
      fp_interp_concat_rows_kernel<true><<<grid1d(rows, 8), 256, 0, bd::as_stream(stream)>>>(dist2, idx, known_feats, C2,
                                                                                           unknown_feats, C1, n, m, out, rows);
    else
      fp_interp_concat_rows_kernel<false><<<grid1d(rows, 8), 256, 0, bd::as_stream(stream)>>>(dist2, idx, known_feats, C2,
                                                                                            unknown_feats, C1, n, m, out, rows);
  } else {
    const long long total = static_cast<long long>(B) * n * (C1 + C2);
    fp_interp_concat_kernel<<<grid1d(total, 256), 256, 0, bd::as_stream(stream)>>>(dist2, idx, known_feats, C2, unknown_feats,
                                                                                  C1, n, m, static_cast<float *>(out), total);
  }
  BD_CHECK_LAUNCH("bd_fp_interp_concat");
  return BD_OK;
}

int bd_fp_interp_concat(const float *dist2, const int *idx, const float *known_feats, int C2,
                        const float *unknown_feats, int C1, int B, int n, int m, float *out, bd_stream_t stream) {
  return fp_interp_concat_impl(dist2, idx, known_feats, C2, unknown_feats, C1, B, n, m, out, 0, stream);
}

int bd_fp_interp_concat_h(const float *dist2, const int *idx, const float *known_feats, int C2,
                          const float *unknown_feats, int C1, int B, int n, int m, void *out, int out_half,
                          bd_stream_t stream) {
  return fp_interp_concat_impl(dist2, idx, known_feats, C2, unknown_feats, C1, B, n, m, out, out_half, stream);
}

int bd_transpose_rows(const float *in, int B, int n, int C, float *out, bd_stream_t stream) {
  BD_REQUIRE(in && out, "bd_transpose_rows: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && C > 0 && B <= 65535, "bd_transpose_rows: bad sizes");
  dim3 grid(bd::ceil_div(n, 32), bd::ceil_div(C, 32), B);
  transpose_rows_kernel<<<grid, dim3(32, 8), 0, bd::as_stream(stream)>>>(in, n, C, out);
  BD_CHECK_LAUNCH("bd_transpose_rows");
  return BD_OK;
}

}  // extern "C"
