// fp32 SIMT linear layer  Y = act((A [+ A2]) · Wᵀ + bias)  — the exact-arithmetic path behind
// the 1e-3 fp32 parity gate (BASELINE.json configs[0]).  It replaces cuDNN 1x1 conv + BN +
// ReLU (pointnet2/pytorch_utils.py:11-36), nn.Linear and nn.Conv1d(k=1) call sites of
// models/bdetr.py / models/modules.py / models/encoder_decoder_layers.py; BatchNorm (eval) is
// folded into W / bias by the host.  The bf16 tcgen05 path lives in gemm_tc.cu.
//
// Both operands are K-major (activations are token-major rows, W is torch's (N,K) layout), so
// global loads are float4 along K and the tile is transposed into shared memory once.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

template <int BM, int BN, int BK, int TM, int TN, bool VEC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
linear_f32_kernel(const float *__restrict__ A, int lda, const float *__restrict__ A2, int lda2,
                  const float *__restrict__ W, const float *__restrict__ bias, float *__restrict__ Y, int ldy, int M,
                  int N, int K, int relu) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TX = BN / TN;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int row0 = blockIdx.x * BM, col0 = blockIdx.y * BN;

  constexpr int A_ELEMS = BM * BK, W_ELEMS = BN * BK;
  constexpr int VW = VEC ? 4 : 1;
  constexpr int A_IT = (A_ELEMS / VW + NT - 1) / NT, W_IT = (W_ELEMS / VW + NT - 1) / NT;
  float ra[A_IT][VW], rw[W_IT][VW];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {
      const int e = tid + it * NT;
#pragma unroll
      for (int v = 0; v < VW; ++v) ra[it][v] = 0.f;
      if (e < A_ELEMS / VW) {
        const int r = e / (BK / VW), kq = (e % (BK / VW)) * VW;
        const int gr = row0 + r, gk = k0 + kq;
        if (gr < M) {
          if constexpr (VEC) {
            if (gk < K) {  // K % 4 == 0 on this path: the whole float4 is in range
              float4 v = *reinterpret_cast<const float4 *>(A + static_cast<long long>(gr) * lda + gk);
              if (A2) {
                const float4 u = *reinterpret_cast<const float4 *>(A2 + static_cast<long long>(gr) * lda2 + gk);
                v.x += u.x, v.y += u.y, v.z += u.z, v.w += u.w;
              }
              ra[it][0] = v.x, ra[it][1] = v.y, ra[it][2] = v.z, ra[it][3] = v.w;
            }
          } else {
            if (gk < K) {
              float v = A[static_cast<long long>(gr) * lda + gk];
              if (A2) v += A2[static_cast<long long>(gr) * lda2 + gk];
              ra[it][0] = v;
            }
          }
        }
      }
    }
#pragma unroll
    for (int it = 0; it < W_IT; ++it) {
      const int e = tid + it * NT;
#pragma unroll
      for (int v = 0; v < VW; ++v) rw[it][v] = 0.f;
      if (e < W_ELEMS / VW) {
        const int r = e / (BK / VW), kq = (e % (BK / VW)) * VW;
        const int gc = col0 + r, gk = k0 + kq;
        if (gc < N && gk < K) {
          if constexpr (VEC) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(W + static_cast<long long>(gc) * K + gk));
            rw[it][0] = v.x, rw[it][1] = v.y, rw[it][2] = v.z, rw[it][3] = v.w;
          } else {
            rw[it][0] = __ldg(W + static_cast<long long>(gc) * K + gk);
          }
        }
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {
      const int e = tid + it * NT;
      if (e < A_ELEMS / VW) {
        const int r = e / (BK / VW), kq = (e % (BK / VW)) * VW;
#pragma unroll
        for (int v = 0; v < VW; ++v) As[kq + v][r] = ra[it][v];
      }
    }
#pragma unroll
    for (int it = 0; it < W_IT; ++it) {
      const int e = tid + it * NT;
      if (e < W_ELEMS / VW) {
        const int r = e / (BK / VW), kq = (e % (BK / VW)) * VW;
#pragma unroll
        for (int v = 0; v < VW; ++v) Ws[kq + v][r] = rw[it][v];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int n_tiles = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles();
  __syncthreads();
  for (int t = 0; t < n_tiles; ++t) {
    if (t + 1 < n_tiles) load_tiles((t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], w[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] = Ws[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
    if (t + 1 < n_tiles) {
      store_tiles();
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gr = row0 + ty * TM + i;
    if (gr >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gc = col0 + tx * TN + j;
      if (gc >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + gc) : 0.f);
      if (relu == 1) v = fmaxf(v, 0.f);
      if (relu == 2) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));  // exact (erf) GELU
      Y[static_cast<long long>(gr) * ldy + gc] = v;
    }
  }
}


// Narrow-input layers (K <= 8: the first layer of the learned position embeddings — xyz, or centre + size, or a
// detected box — models/modules.py PositionEmbeddingLearned, models/bdetr.py:217-225): 3 to 8 multiply-adds per
// output, i.e. a pure write stream.  W^T and the bias sit in shared memory, one warp walks over rows, a lane
// owns column pairs (4- or 8-byte stores, a full line per warp); the row can be written as fp16 (HALF: the operand
// format of the layer that follows, same fp32 value rounded once).  Same operation order as linear_f32_kernel
// (fma over ascending k, bias added last).
template <bool HALF>
__global__ void __launch_bounds__(256)
linear_smallk_kernel(const float *__restrict__ A, int lda, const float *__restrict__ W, const float *__restrict__ bias,
                     void *__restrict__ Y, int ldy, int M, int N, int K, int relu) {
  extern __shared__ float sk_smem[];  // Wt[K][Np] then bias[Np], Np = N rounded up to 2
  const int Np = (N + 1) & ~1;
  float *Wt = sk_smem, *bs = sk_smem + K * Np;
  for (int i = threadIdx.x; i < K * Np; i += blockDim.x) {
    const int k = i / Np, n = i - k * Np;
    Wt[i] = n < N ? __ldg(W + static_cast<long long>(n) * K + k) : 0.f;
  }
  for (int i = threadIdx.x; i < Np; i += blockDim.x) bs[i] = (bias && i < N) ? __ldg(bias + i) : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  constexpr int RW = 4;  // rows per warp and pass: their input loads and output stores are in flight together
  for (long long r0 = (static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5)) * RW; r0 < M;
       r0 += static_cast<long long>(gridDim.x) * wpb * RW) {
    float x[RW][8];
#pragma unroll
    for (int u = 0; u < RW; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) x[u][k] = (k < K && r0 + u < M) ? __ldg(A + (r0 + u) * lda + k) : 0.f;
    for (int n = 2 * lane; n < N; n += 64) {
      float a0[RW], a1[RW];
#pragma unroll
      for (int u = 0; u < RW; ++u) a0[u] = a1[u] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < K) {
          const float2 w = *reinterpret_cast<const float2 *>(Wt + k * Np + n);
#pragma unroll
          for (int u = 0; u < RW; ++u) a0[u] = fmaf(x[u][k], w.x, a0[u]), a1[u] = fmaf(x[u][k], w.y, a1[u]);
        }
      }
      const float b0 = bs[n], b1 = bs[n + 1];
#pragma unroll
      for (int u = 0; u < RW; ++u) {
        const long long r = r0 + u;
        if (r >= M) break;
        float v0 = a0[u] + b0, v1 = a1[u] + b1;
        if (relu) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f);
        if (HALF) {
          __half *y = static_cast<__half *>(Y) + r * ldy + n;
          if (n + 1 < N) *reinterpret_cast<__half2 *>(y) = __floats2half2_rn(fminf(fmaxf(v0, -65504.f), 65504.f), fminf(fmaxf(v1, -65504.f), 65504.f));
          else y[0] = __float2half_rn(fminf(fmaxf(v0, -65504.f), 65504.f));
        } else {
          float *y = static_cast<float *>(Y) + r * ldy + n;
          if (n + 1 < N) *reinterpret_cast<float2 *>(y) = make_float2(v0, v1);
          else y[0] = v0;
        }
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
void launch(const float *A, int lda, const float *A2, int lda2, const float *W, const float *bias, float *Y, int ldy,
            int M, int N, int K, int relu, cudaStream_t s) {
  constexpr int BK = 16;
  dim3 grid(bd::ceil_div(M, BM), bd::ceil_div(N, BN));
  const bool vec = (K % 4 == 0) && (lda % 4 == 0) && (!A2 || lda2 % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (!A2 || reinterpret_cast<uintptr_t>(A2) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(W) % 16 == 0);
  if (vec)
    linear_f32_kernel<BM, BN, BK, TM, TN, true>
        <<<grid, (BM / TM) * (BN / TN), 0, s>>>(A, lda, A2, lda2, W, bias, Y, ldy, M, N, K, relu);
  else
    linear_f32_kernel<BM, BN, BK, TM, TN, false>
        <<<grid, (BM / TM) * (BN / TN), 0, s>>>(A, lda, A2, lda2, W, bias, Y, ldy, M, N, K, relu);
}

}  // namespace

extern "C" int bd_linear_f32(const float *A, int lda, const float *A2, int lda2, const float *W, const float *bias,
                             float *Y, int ldy, int M, int N, int K, int relu, bd_stream_t stream) {
  BD_REQUIRE(A && W && Y, "bd_linear_f32: null pointer");
  BD_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldy >= N && (!A2 || lda2 >= K), "bd_linear_f32: bad sizes");
  BD_REQUIRE(bd::ceil_div(N, 16) <= 65535, "bd_linear_f32: N too large");
  cudaStream_t s = bd::as_stream(stream);
  if (M >= 4096)
    launch<128, 64, 8, 4>(A, lda, A2, lda2, W, bias, Y, ldy, M, N, K, relu, s);
  else if (N <= 16)
    launch<64, 16, 4, 1>(A, lda, A2, lda2, W, bias, Y, ldy, M, N, K, relu, s);
  else
    launch<32, 64, 2, 4>(A, lda, A2, lda2, W, bias, Y, ldy, M, N, K, relu, s);
  BD_CHECK_LAUNCH("bd_linear_f32");
  return BD_OK;
}

// Y (M, N) = act(A (M, K <= 8) . W^T + bias), fp32 rows or (y_half) fp16 rows; ldy % 2 == 0 and Y 8-byte aligned
// (fp32) / 4-byte aligned (fp16).  relu = 0 / 1.
extern "C" int bd_linear_smallk(const float *A, int lda, const float *W, const float *bias, void *Y, int ldy, int y_half,
                                int M, int N, int K, int relu, bd_stream_t stream) {
  BD_REQUIRE(A && W && Y, "bd_linear_smallk: null pointer");
  BD_REQUIRE(M > 0 && N > 0 && K > 0 && K <= 8 && lda >= K && ldy >= N && ldy % 2 == 0 && (relu == 0 || relu == 1) &&
                 (reinterpret_cast<uintptr_t>(Y) & 7) == 0 && N <= 4096,
             "bd_linear_smallk: needs 0 < K <= 8, N <= 4096, even ldy, 8-byte aligned Y, relu in {0, 1}");
  const int Np = (N + 1) & ~1;
  const size_t smem = sizeof(float) * static_cast<size_t>(K + 1) * Np;
  const int blocks = bd::ceil_div(M, 8 * 4 * 4) < 4 * bd::sm_count() ? bd::ceil_div(M, 8 * 4 * 4) : 4 * bd::sm_count();
  if (y_half)
    linear_smallk_kernel<true><<<blocks > 0 ? blocks : 1, 256, smem, bd::as_stream(stream)>>>(A, lda, W, bias, Y, ldy, M, N, K, relu);
  else
    linear_smallk_kernel<false><<<blocks > 0 ? blocks : 1, 256, smem, bd::as_stream(stream)>>>(A, lda, W, bias, Y, ldy, M, N, K, relu);
  BD_CHECK_LAUNCH("bd_linear_smallk");
  return BD_OK;
}
