// bf16 tensor-core linear layer on tcgen05:  Y = act((A [+ A2]) · Wᵀ + bias)
//
//   A  : (M, K) fp32 row-major activations (token-major), converted to bf16 while being staged
//        into shared memory in the canonical K-major core-matrix layout (tc_common.cuh);
//        the optional second operand A2 (positional embedding) is added during staging.
//   Wp : weights PRE-PACKED by the host into that same layout, one contiguous block per
//        (n-tile, k-chunk), so each block arrives with a single cp.async.bulk (TMA engine)
//        signalled on an mbarrier.  Packing: Wp[nt][kc][part][BN/8][KC/8][8 rows][8 k] bf16,
//        zero padded to BN x KC; part = {hi} (split 1) or {hi, lo} (split 3, bf16x3).
//   D  : 128 x BN fp32 accumulator in TMEM (tcgen05.mma cta_group::1, M = 128, K = 16 per
//        instruction, issued by one thread); epilogue tcgen05.ld -> bias / ReLU -> fp32 store.
//
// One CTA = one 128-row x BN-column output tile; K is consumed in chunks of KC <= 288 that are
// staged whole (A chunk <= 72 KB, W chunk <= 90 KB).  Replaces cuBLAS / cuDNN-1x1-conv call
// sites of the reference in the bf16 mode (BASELINE.json configs[1]).
#include "tc_common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 256;

__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tc_kernel(const float *__restrict__ A, int lda, const float *__restrict__ A2, int lda2,
                 const __nv_bfloat16 *__restrict__ Wp, const float *__restrict__ bias, float *__restrict__ Y, int ldy,
                 int M, int N, int K, int KC, int n_chunks, int BN, int relu, int split) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_w, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TC_BM;
  const int nt = blockIdx.y;
  // split == 3: every fp32 operand x is carried as bf16 hi + bf16 lo (x ~ hi + lo to 16 mantissa
  // bits) and D += Ahi*Whi + Alo*Whi + Ahi*Wlo, i.e. fp32-grade products at 3 MMAs per k-step.
  const uint32_t parts = split == 3 ? 2u : 1u;
  const uint32_t a_part = TC_BM * KC * 2, w_part = static_cast<uint32_t>(BN) * KC * 2;
  const uint32_t a_bytes = a_part * parts, w_bytes = w_part * parts;
  unsigned char *sA = smem;
  unsigned char *sW = smem + a_bytes;
  const uint32_t sbo = (KC / 8) * 128;
  const uint32_t ncols = tc::tmem_cols_pow2(BN);

  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), ncols);
  if (tid == 32) {
    tc::mbar_init(tc::smem_u32(&bar_w), 1);
    tc::mbar_init(tc::smem_u32(&bar_mma), 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = tc::idesc_bf16(TC_BM, BN);
  const int chunks_per_row = KC / 8;

  for (int c = 0; c < n_chunks; ++c) {
    if (c > 0) tc::mbar_wait(tc::smem_u32(&bar_mma), (c - 1) & 1);  // previous chunk's MMAs have read smem
    if (tid == 0) {
      tc::mbar_arrive_expect_tx(tc::smem_u32(&bar_w), w_bytes);
      tc::bulk_g2s(tc::smem_u32(sW), Wp + (static_cast<size_t>(nt) * n_chunks + c) * BN * KC * parts, w_bytes,
                   tc::smem_u32(&bar_w));
    }
    // stage A: each thread converts 8 consecutive k of one row (32 B in, 16 B out).  Lanes are
    // mapped (row % 8 fastest, then 4 adjacent k-chunks) so a warp writes 512 contiguous bytes.
    const int k_base = c * KC;
    const int n_quads = (chunks_per_row + 3) / 4;
    for (int e = tid; e < TC_BM / 8 * n_quads * 32; e += TC_THREADS) {
      const int l = e & 31, blk = e >> 5;
      const int rg = blk / n_quads, q = blk % n_quads;
      const int r = rg * 8 + (l & 7), ch = q * 4 + (l >> 3);
      if (ch >= chunks_per_row) continue;
      const int gr = row0 + r, gk = k_base + ch * 8;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      if (gr < M && gk < K) {
        const float *src = A + static_cast<long long>(gr) * lda + gk;
        const float *src2 = A2 ? A2 + static_cast<long long>(gr) * lda2 + gk : nullptr;
        if (gk + 8 <= K && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
            (!src2 || (reinterpret_cast<uintptr_t>(src2) & 15) == 0)) {
          const float4 a = *reinterpret_cast<const float4 *>(src), b = *reinterpret_cast<const float4 *>(src + 4);
          v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
          if (src2) {
            const float4 p = *reinterpret_cast<const float4 *>(src2), q2 = *reinterpret_cast<const float4 *>(src2 + 4);
            v[0] += p.x, v[1] += p.y, v[2] += p.z, v[3] += p.w, v[4] += q2.x, v[5] += q2.y, v[6] += q2.z, v[7] += q2.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (gk + i < K) v[i] = src[i] + (src2 ? src2[i] : 0.f);
        }
      }
      uint4 pk;
      pk.x = tc::pack_bf16x2(v[0], v[1]), pk.y = tc::pack_bf16x2(v[2], v[3]);
      pk.z = tc::pack_bf16x2(v[4], v[5]), pk.w = tc::pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4 *>(sA + tc::canon_off(r, ch, sbo)) = pk;
      if (parts == 2) {
        float lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) lo[i] = v[i] - __bfloat162float(__float2bfloat16_rn(v[i]));
        pk.x = tc::pack_bf16x2(lo[0], lo[1]), pk.y = tc::pack_bf16x2(lo[2], lo[3]);
        pk.z = tc::pack_bf16x2(lo[4], lo[5]), pk.w = tc::pack_bf16x2(lo[6], lo[7]);
        *reinterpret_cast<uint4 *>(sA + a_part + tc::canon_off(r, ch, sbo)) = pk;
      }
    }
    tc::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::mbar_wait(tc::smem_u32(&bar_w), c & 1);
      tc::fence_after_sync();
      const uint32_t a0 = tc::smem_u32(sA), w0 = tc::smem_u32(sW);
      for (int s = 0; s < KC / 16; ++s) {
        const uint64_t da = tc::smem_desc(a0 + s * 256, 128, sbo);
        const uint64_t db = tc::smem_desc(w0 + s * 256, 128, sbo);
        tc::mma_bf16(tmem, da, db, idesc, (c > 0 || s > 0) ? 1u : 0u);
        if (parts == 2) {
          tc::mma_bf16(tmem, tc::smem_desc(a0 + a_part + s * 256, 128, sbo), db, idesc, 1u);
          tc::mma_bf16(tmem, da, tc::smem_desc(w0 + w_part + s * 256, 128, sbo), idesc, 1u);
        }
      }
      tc::mma_commit(tc::smem_u32(&bar_mma));
    }
  }
  tc::mbar_wait(tc::smem_u32(&bar_mma), (n_chunks - 1) & 1);
  tc::fence_after_sync();

  // epilogue: warp w reads TMEM lanes 32*(w%4)..+31 (its rows); the two warpgroups split the columns
  const int r = (warp & 3) * 32 + lane;
  const int gr = row0 + r;
  const int half = (BN / 16 + 1) / 2;  // 16-column groups per warpgroup
  const int g0 = (warp >> 2) * half, g1 = min(BN / 16, g0 + half);
  for (int g = g0; g < g1; ++g) {
    uint32_t acc[16];
    tc::tmem_ld16(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g * 16, acc);
    tc::tmem_ld_wait();
    if (gr < M) {
      const int col0 = nt * BN + g * 16;
      float *dst = Y + static_cast<long long>(gr) * ldy + col0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = col0 + j;
        if (col < N) {
          float v = __uint_as_float(acc[j]) + (bias ? __ldg(bias + col) : 0.f);
          if (relu) v = fmaxf(v, 0.f);
          dst[j] = v;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

}  // namespace

extern "C" int bd_linear_tc(const float *A, int lda, const float *A2, int lda2, const void *Wp, const float *bias,
                            float *Y, int ldy, int M, int N, int K, int KC, int n_chunks, int BN, int relu,
                            int split, bd_stream_t stream) {
  BD_REQUIRE(A && Wp && Y, "bd_linear_tc: null pointer");
  BD_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldy >= N && (!A2 || lda2 >= K), "bd_linear_tc: bad sizes");
  BD_REQUIRE(KC % 16 == 0 && KC >= 16 && KC <= 288 && n_chunks >= 1 && n_chunks * KC >= K,
             "bd_linear_tc: KC must be a multiple of 16 in [16,288] covering K");
  BD_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= 256, "bd_linear_tc: BN must be a multiple of 16 in [16,256]");
  const int n_tiles = bd::ceil_div(N, BN);
  BD_REQUIRE(n_tiles <= 65535, "bd_linear_tc: N too large");
  BD_REQUIRE(split == 1 || split == 3, "bd_linear_tc: split must be 1 (bf16) or 3 (bf16x3)");
  const size_t smem = static_cast<size_t>(TC_BM + BN) * KC * 2 * (split == 3 ? 2 : 1);
  BD_REQUIRE(smem <= 226 * 1024, "bd_linear_tc: tile does not fit shared memory");
  static thread_local bool configured = false;
  if (!configured) {
    BD_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024),
            "bd_linear_tc");
    configured = true;
  }
  dim3 grid(bd::ceil_div(M, TC_BM), n_tiles);
  linear_tc_kernel<<<grid, TC_THREADS, smem, bd::as_stream(stream)>>>(
      A, lda, A2, lda2, static_cast<const __nv_bfloat16 *>(Wp), bias, Y, ldy, M, N, K, KC, n_chunks, BN, relu, split);
  BD_CHECK_LAUNCH("bd_linear_tc");
  return BD_OK;
}
