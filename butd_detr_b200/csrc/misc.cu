// Small row-wise operators of the forward (LayerNorm, top-k query selection, L2 normalise,
// embedding lookup, adds) + library-level plumbing (error string, version).
#include "common.cuh"

#include <stdarg.h>

namespace bd {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int g_pdl = 1;

int sm_count() {
  static int cached[64] = {0};  // per device (benign race: every writer stores the same value)
  int dev = 0;
  cudaGetDevice(&dev);
  int n = cached[dev & 63];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cached[dev & 63] = n;
  }
  return n;
}

}  // namespace bd

namespace {

// One warp per row.  Y = LN(X + R) * gamma + beta, biased variance (torch.nn.LayerNorm).
constexpr int LN_MAX_PER_LANE = 32;
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const float *__restrict__ X, const float *__restrict__ R, const float *__restrict__ gamma,
                     const float *__restrict__ beta, float *__restrict__ Y, int M, int D, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float *x = X + static_cast<long long>(row) * D;
  const float *r = R ? R + static_cast<long long>(row) * D : nullptr;
  float v[LN_MAX_PER_LANE];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + i * 32;
    v[i] = 0.f;
    if (c < D) {
      v[i] = x[c] + (r ? r[c] : 0.f);
      sum += v[i];
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, off);
  const float mean = sum / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < D) {
      const float d = v[i] - mean;
      sq = fmaf(d, d, sq);
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, off);
  const float rstd = rsqrtf(sq / static_cast<float>(D) + eps);
  float *y = Y + static_cast<long long>(row) * D;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < D) y[c] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
  }
}

// RoBERTa input embeddings (the text side; transformers' RobertaEmbeddings.forward as
// called from /root/reference/models/bdetr.py:168): word[id] + position[pid] + token_type[0] -> LayerNorm, with
// pid = pad_idx + (number of non-pad tokens up to and including this one) for a non-pad token, pad_idx for a pad
// token.  One warp per token; ids outside the tables are clamped (no fault on a bad id).
__global__ void __launch_bounds__(256)
roberta_embed_kernel(const long long *__restrict__ ids, const float *__restrict__ word, int vocab,
                     const float *__restrict__ pos, int n_pos, const float *__restrict__ type,
                     const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ Y, int B, int L,
                     int D, int pad_idx, float eps) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<long long>(B) * L) return;
  const int l = static_cast<int>(row % L);
  const long long *rid = ids + (row - l);
  int cnt = 0;
  for (int i = lane; i <= l; i += 32) cnt += rid[i] != pad_idx;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, off);
  long long id = rid[l];
  int pid = id != pad_idx ? pad_idx + cnt : pad_idx;
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  pid = pid >= n_pos ? n_pos - 1 : pid;
  const float *w = word + id * D, *pe = pos + static_cast<long long>(pid) * D;
  float v[LN_MAX_PER_LANE];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + i * 32;
    v[i] = 0.f;
    if (c < D) {
      v[i] = __ldg(w + c) + __ldg(pe + c) + __ldg(type + c);
      sum += v[i];
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, off);
  const float mean = sum / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < D) {
      const float d = v[i] - mean;
      sq = fmaf(d, d, sq);
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, off);
  const float rstd = rsqrtf(sq / static_cast<float>(D) + eps);
  float *y = Y + row * D;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < D) y[c] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
  }
}

// Query selection: sigmoid, then a full bitonic sort of (value, ~index) keys in shared memory.
constexpr int TOPK_MAX = 4096;
__global__ void __launch_bounds__(1024)
topk_sigmoid_kernel(const float *__restrict__ logits, int n, int n_pad, int k, int *__restrict__ idx) {
  __shared__ unsigned long long keys[TOPK_MAX];
  const float *row = logits + static_cast<long long>(blockIdx.x) * n;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
    unsigned long long key = 0ull;  // padding sorts last
    if (i < n) {
      const float s = 1.0f / (1.0f + expf(-row[i]));  // torch.sigmoid, fp32
      // sigmoid > 0 so its bit pattern orders like the value; NaN logits sort first like torch
      key = (static_cast<unsigned long long>(__float_as_uint(s)) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(i));
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < n_pad / 2; t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x)
    idx[static_cast<long long>(blockIdx.x) * k + i] = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(keys[i]));
}

__global__ void __launch_bounds__(256)
l2_normalize_rows_kernel(const float *__restrict__ X, float *__restrict__ Y, int M, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float *x = X + static_cast<long long>(row) * D;
  float sq = 0.f;
  for (int c = lane; c < D; c += 32) sq = fmaf(x[c], x[c], sq);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, off);
  const float denom = fmaxf(sqrtf(sq), 1e-12f);  // F.normalize eps
  float *y = Y + static_cast<long long>(row) * D;
  for (int c = lane; c < D; c += 32) y[c] = x[c] / denom;
}

__global__ void embedding_rows_kernel(const float *__restrict__ table, int w, const long long *__restrict__ ids,
                                      float *__restrict__ out, int ld_out, long long total) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c = static_cast<int>(e % w);
  const long long r = e / w;
  out[r * ld_out + c] = __ldg(table + ids[r] * w + c);
}

__global__ void add_rows_kernel(const float *__restrict__ X1, int ld1, const float *__restrict__ X2, int ld2,
                                float *__restrict__ Y, int ldy, int w, long long total) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c = static_cast<int>(e % w);
  const long long r = e / w;
  Y[r * ldy + c] = X1[r * ld1 + c] + X2[r * ld2 + c];
}

__global__ void concat_rows_kernel(const float *__restrict__ X1, int ld1, int w1, const float *__restrict__ X2,
                                   int ld2, int w2, float *__restrict__ Y, int ldy, long long total) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int w = w1 + w2;
  const int c = static_cast<int>(e % w);
  const long long r = e / w;
  Y[r * ldy + c] = c < w1 ? X1[r * ld1 + c] : X2[r * ld2 + (c - w1)];
}

}  // namespace

extern "C" {

int bd_concat_rows(const float *X1, int ld1, int w1, const float *X2, int ld2, int w2, float *Y, int ldy, int M,
                   bd_stream_t stream) {
  BD_REQUIRE(X1 && X2 && Y, "bd_concat_rows: null pointer");
  BD_REQUIRE(M > 0 && w1 > 0 && w2 > 0 && ld1 >= w1 && ld2 >= w2 && ldy >= w1 + w2, "bd_concat_rows: bad sizes");
  const long long total = static_cast<long long>(M) * (w1 + w2);
  concat_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, bd::as_stream(stream)>>>(
      X1, ld1, w1, X2, ld2, w2, Y, ldy, total);
  BD_CHECK_LAUNCH("bd_concat_rows");
  return BD_OK;
}

int bd_version(void) { return 100; }
const char *bd_last_error(void) { return bd::g_err; }
const char *bd_arch(void) { return "sm_100a"; }

int bd_add_layernorm_f32(const float *X, const float *R, const float *gamma, const float *beta, float *Y, int M,
                         int D, float eps, bd_stream_t stream) {
  BD_REQUIRE(X && gamma && beta && Y, "bd_add_layernorm_f32: null pointer");
  BD_REQUIRE(M > 0 && D > 0 && D <= 32 * LN_MAX_PER_LANE, "bd_add_layernorm_f32: bad sizes (D <= 1024)");
  add_layernorm_kernel<<<bd::ceil_div(M, 8), 256, 0, bd::as_stream(stream)>>>(X, R, gamma, beta, Y, M, D, eps);
  BD_CHECK_LAUNCH("bd_add_layernorm_f32");
  return BD_OK;
}

int bd_roberta_embed(const long long *ids, const float *word, int vocab, const float *pos, int n_pos, const float *type,
                     const float *gamma, const float *beta, float *Y, int B, int L, int D, int pad_idx, float eps,
                     bd_stream_t stream) {
  BD_REQUIRE(ids && word && pos && type && gamma && beta && Y, "bd_roberta_embed: null pointer");
  BD_REQUIRE(B > 0 && L > 0 && D > 0 && D <= 32 * LN_MAX_PER_LANE && vocab > 0 && n_pos > pad_idx && pad_idx >= 0,
             "bd_roberta_embed: bad sizes (D <= 1024, pad_idx < n_pos)");
  const long long rows = static_cast<long long>(B) * L;
  roberta_embed_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, bd::as_stream(stream)>>>(
      ids, word, vocab, pos, n_pos, type, gamma, beta, Y, B, L, D, pad_idx, eps);
  BD_CHECK_LAUNCH("bd_roberta_embed");
  return BD_OK;
}

int bd_topk_sigmoid(const float *logits, int B, int n, int k, int *idx, bd_stream_t stream) {
  BD_REQUIRE(logits && idx, "bd_topk_sigmoid: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && k > 0 && k <= n && n <= TOPK_MAX, "bd_topk_sigmoid: need 0 < k <= n <= 4096");
  int n_pad = 2;
  while (n_pad < n) n_pad <<= 1;
  topk_sigmoid_kernel<<<B, 1024, 0, bd::as_stream(stream)>>>(logits, n, n_pad, k, idx);
  BD_CHECK_LAUNCH("bd_topk_sigmoid");
  return BD_OK;
}

int bd_l2_normalize_rows(const float *X, float *Y, int M, int D, bd_stream_t stream) {
  BD_REQUIRE(X && Y, "bd_l2_normalize_rows: null pointer");
  BD_REQUIRE(M > 0 && D > 0, "bd_l2_normalize_rows: bad sizes");
  l2_normalize_rows_kernel<<<bd::ceil_div(M, 8), 256, 0, bd::as_stream(stream)>>>(X, Y, M, D);
  BD_CHECK_LAUNCH("bd_l2_normalize_rows");
  return BD_OK;
}

int bd_embedding_rows(const float *table, int w, const long long *ids, int M, float *out, int ld_out,
                      bd_stream_t stream) {
  BD_REQUIRE(table && ids && out, "bd_embedding_rows: null pointer");
  BD_REQUIRE(M > 0 && w > 0 && ld_out >= w, "bd_embedding_rows: bad sizes");
  const long long total = static_cast<long long>(M) * w;
  embedding_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, bd::as_stream(stream)>>>(
      table, w, ids, out, ld_out, total);
  BD_CHECK_LAUNCH("bd_embedding_rows");
  return BD_OK;
}

int bd_add_rows(const float *X1, int ld1, const float *X2, int ld2, float *Y, int ldy, int M, int w,
                bd_stream_t stream) {
  BD_REQUIRE(X1 && X2 && Y, "bd_add_rows: null pointer");
  BD_REQUIRE(M > 0 && w > 0 && ld1 >= w && ld2 >= w && ldy >= w, "bd_add_rows: bad sizes");
  const long long total = static_cast<long long>(M) * w;
  add_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, bd::as_stream(stream)>>>(X1, ld1, X2, ld2, Y,
                                                                                                 ldy, w, total);
  BD_CHECK_LAUNCH("bd_add_rows");
  return BD_OK;
}

}  // extern "C"

// 1 (default): the tensor-core kernels are launched with programmatic stream serialization (PDL),
// 0: plain stream order.  Process-wide; set it before capturing a CUDA graph.
extern "C" int bd_set_pdl(int enabled) {
  bd::g_pdl = enabled ? 1 : 0;
  return BD_OK;
}
