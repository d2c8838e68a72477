/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's nine PointNet++ native ops, whose only
 * implementation in nickgkan/butd_detr is CUDA (`pointnet2/_ext_src/src/*.cu`; every host
 * wrapper rejects CPU tensors, e.g. ball_query.cpp:32-34).  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load this library.
 *
 * Floating-point contract: the reference is compiled by nvcc with the default -fmad=true,
 * and the sm_100 SASS of the unmodified sources (oracle/build_ref_ext.py + cuobjdump) shows
 *     d  = FFMA(dz,dz, FFMA(dx,dx, FMUL(dy,dy)))          (FPS, ball query, three_nn)
 *     mag= FFMA(z,z,  FFMA(x,x,  FMUL(y,y)))  then a DOUBLE compare against 1e-3   (FPS)
 *     out= FFMA(p3,w3, FFMA(p1,w1, FMUL(p2,w2)))          (three_interpolate)
 * (for a*a + b*b the compiler multiplies the SECOND product and fuses the first into it)
 * so those contractions are written here with explicit fmaf(); compile with
 * -ffp-contract=off so the C compiler adds none of its own.
 *
 * Parity pin: tests/golden/pointops_refcuda.npz holds outputs of the reference's own CUDA
 * kernels (built unmodified, run on a B200) for seeded inputs; tests/test_oracle_pointops.py
 * checks this file against them bit-for-bit (indices) / exactly (distances).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_TOTAL_THREADS 512

/* cuda_utils.h:20-24 — largest power of two <= work_size, clamped to [1, 512]; the same
 * double-precision log ratio as the reference so boundary behaviour is identical. */
int orc_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int v = 1 << pow_2;
  if (v > ORC_TOTAL_THREADS) v = ORC_TOTAL_THREADS;
  if (v < 1) v = 1;
  return v;
}

static inline float sqdist_fma(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* sampling_gpu.cu:74-178 (kernel), sampling.cpp:70-91 (host: temp = 1e10, idx zeros).
 * One CUDA block of bs threads per batch element; thread t scans k = t, t+bs, ... keeping the
 * first strict maximum (:113-114); the shared-memory tree (:120-173) keeps the LOWER slot on
 * ties (`v2 > v1 ? i2 : i1`, :64-70).  Emulated literally: per-"thread" running best, then
 * the same tree. */
int orc_fps(const float *xyz, int B, int N, int m, int *idx) {
  if (m <= 0) return 0;
  const int bs = orc_opt_n_threads(N);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const float *p = xyz + (size_t)b * N * 3;
    int *out = idx + (size_t)b * m;
    float *temp = (float *)malloc(sizeof(float) * (size_t)N);
    unsigned char *skip = (unsigned char *)malloc((size_t)N);
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    for (int k = 0; k < N; ++k) {
      temp[k] = 1e10f;
      const float x = p[3 * k], y = p[3 * k + 1], z = p[3 * k + 2];
      const float mag = fmaf(z, z, fmaf(x, x, y * y));
      skip[k] = ((double)mag <= 1e-3) ? 1 : 0; /* :105-106, double compare */
    }
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < m; ++j) {
      const float x1 = p[3 * old], y1 = p[3 * old + 1], z1 = p[3 * old + 2];
      for (int t = 0; t < bs; ++t) { dists[t] = -1.0f; dists_i[t] = 0; }
      for (int k = 0; k < N; ++k) {
        if (skip[k]) continue;
        const float d = sqdist_fma(p[3 * k], p[3 * k + 1], p[3 * k + 2], x1, y1, z1);
        const float d2 = d < temp[k] ? d : temp[k];
        temp[k] = d2;
        const int t = k % bs;
        if (d2 > dists[t]) { dists[t] = d2; dists_i[t] = k; }
      }
      for (int s = bs / 2; s >= 1; s >>= 1) {
        for (int t = 0; t < s; ++t) {
          const float v1 = dists[t], v2 = dists[t + s];
          const int i1 = dists_i[t], i2 = dists_i[t + s];
          dists[t] = v1 > v2 ? v1 : v2;
          dists_i[t] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      out[j] = old;
    }
    free(temp); free(skip); free(dists); free(dists_i);
  }
  return 0;
}

/* sampling_gpu.cu:13-25 — out[b,c,j] = points[b,c,idx[b,j]] */
int orc_gather(const float *points, const int *idx, int B, int C, int N, int m, float *out) {
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < m; ++j)
        out[((size_t)b * C + c) * m + j] = points[((size_t)b * C + c) * N + idx[(size_t)b * m + j]];
  return 0;
}

/* sampling_gpu.cu:39-52 — scatter-add (atomicAdd in the reference; index order here). */
int orc_gather_grad(const float *grad_out, const int *idx, int B, int C, int N, int m,
                    float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < m; ++j)
        grad_points[((size_t)b * C + c) * N + idx[(size_t)b * m + j]] +=
            grad_out[((size_t)b * C + c) * m + j];
  return 0;
}

/* ball_query_gpu.cu:14-49; host zero-fills idx (ball_query.cpp:24-26).  Arguments in the
 * extension's order (new_xyz, xyz, radius, nsample), pointnet2_utils.py:282. */
int orc_ball_query(const float *new_xyz, const float *xyz, int B, int n, int m, float radius,
                   int nsample, int *idx) {
  const float radius2 = radius * radius;
  memset(idx, 0, sizeof(int) * (size_t)B * m * nsample);
#pragma omp parallel for collapse(2) schedule(dynamic, 16)
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < m; ++j) {
      const float *p = xyz + (size_t)b * n * 3;
      const float *c = new_xyz + ((size_t)b * m + j) * 3;
      int *o = idx + ((size_t)b * m + j) * nsample;
      const float cx = c[0], cy = c[1], cz = c[2];
      int cnt = 0;
      for (int k = 0; k < n && cnt < nsample; ++k) {
        const float d2 = sqdist_fma(cx, cy, cz, p[3 * k], p[3 * k + 1], p[3 * k + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[l] = k;
          o[cnt] = k;
          ++cnt;
        }
      }
    }
  return 0;
}

/* group_points_gpu.cu:13-33 — out[b,c,j,s] = points[b,c,idx[b,j,s]] */
int orc_group(const float *points, const int *idx, int B, int C, int n, int m, int ns, float *out) {
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float *src = points + ((size_t)b * C + c) * n;
      const int *ib = idx + (size_t)b * m * ns;
      float *dst = out + ((size_t)b * C + c) * m * ns;
      for (int i = 0; i < m * ns; ++i) dst[i] = src[ib[i]];
    }
  return 0;
}

/* group_points_gpu.cu:48-69 */
int orc_group_grad(const float *grad_out, const int *idx, int B, int C, int n, int m, int ns,
                   float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * n);
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      float *dst = grad_points + ((size_t)b * C + c) * n;
      const int *ib = idx + (size_t)b * m * ns;
      const float *src = grad_out + ((size_t)b * C + c) * m * ns;
      for (int i = 0; i < m * ns; ++i) dst[ib[i]] += src[i];
    }
  return 0;
}

/* interpolate_gpu.cu:14-64 — three nearest `known` for each `unknown`; double running bests
 * initialised to 1e40 (:32), strict '<' insertion, SQUARED distances returned as float. */
int orc_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                 int *idx) {
#pragma omp parallel for collapse(2) schedule(static, 64)
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < n; ++j) {
      const float *u = unknown + ((size_t)b * n + j) * 3;
      const float *kn = known + (size_t)b * m * 3;
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int i1 = 0, i2 = 0, i3 = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sqdist_fma(u[0], u[1], u[2], kn[3 * k], kn[3 * k + 1], kn[3 * k + 2]);
        if (d < best1) {
          best3 = best2; i3 = i2; best2 = best1; i2 = i1; best1 = d; i1 = k;
        } else if (d < best2) {
          best3 = best2; i3 = i2; best2 = d; i2 = k;
        } else if (d < best3) {
          best3 = d; i3 = k;
        }
      }
      float *od = dist2 + ((size_t)b * n + j) * 3;
      int *oi = idx + ((size_t)b * n + j) * 3;
      od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
      oi[0] = i1; oi[1] = i2; oi[2] = i3;
    }
  return 0;
}

/* interpolate_gpu.cu:77-106 — out[b,c,j] = sum_t points[b,c,idx[b,j,t]] * weight[b,j,t] */
int orc_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C,
                          int m, int n, float *out) {
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float *src = points + ((size_t)b * C + c) * m;
      for (int j = 0; j < n; ++j) {
        const int *ii = idx + ((size_t)b * n + j) * 3;
        const float *w = weight + ((size_t)b * n + j) * 3;
        out[((size_t)b * C + c) * n + j] =
            fmaf(src[ii[2]], w[2], fmaf(src[ii[0]], w[0], src[ii[1]] * w[1]));
      }
    }
  return 0;
}

/* interpolate_gpu.cu:121-148 */
int orc_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B,
                               int C, int n, int m, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * m);
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      float *dst = grad_points + ((size_t)b * C + c) * m;
      for (int j = 0; j < n; ++j) {
        const int *ii = idx + ((size_t)b * n + j) * 3;
        const float *w = weight + ((size_t)b * n + j) * 3;
        const float g = grad_out[((size_t)b * C + c) * n + j];
        dst[ii[0]] += g * w[0];
        dst[ii[1]] += g * w[1];
        dst[ii[2]] += g * w[2];
      }
    }
  return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
