// Persistent tensor-core linear layer for 16-bit rows in, 16-bit rows out (fp16 mode):
//     Y16 = act(A16 · Wᵀ + bias)
// the hot shape of the transformer (projections, FFN hidden layers, memory K / V projections, prediction-head
// stems: 1x1 conv / nn.Linear call sites of /root/reference/models/encoder_decoder_layers.py and modules.py).
//
// linear_tc_kernel (gemm_tc.cu) runs ONE CTA per SM for the 288-wide outputs (two 144-column accumulators need
// all 512 TMEM columns) and a CTA is a serial chain: operand copies -> MMAs -> TMEM drain -> store -> exit, about
// 10 us per 128-row tile of which the copies need 3.3 us.  Here a CTA is PERSISTENT over the row tiles of its
// column group and the three roles run decoupled across tiles:
//   loader warp    streams (A chunk by ONE 2-D tensor copy in the 128-byte-swizzle operand layout, weight chunk by
//                  one bulk copy) through an S-stage ring WITHOUT stopping at tile boundaries: the copies of tile
//                  i + 1 run under the MMAs, the drain and the store of tile i
//   MMA warp       tcgen05.mma into the accumulators (TMEM); waits for the epilogue's "accumulators drained"
//   epilogue warps TMEM -> +bias, ReLU inside the fp16 conversion -> their OWN shared-memory tile (not the ring) ->
//                  one bulk store (TMA engine) per row; the tile is reused once those stores have read it
// Same operand formats, same MMA order, same epilogue arithmetic as linear_tc_kernel: results are bit-identical.
#include "tc_common.cuh"

namespace {

constexpr int ST_BM = 128;
constexpr int ST_WARPS = 8;                      // epilogue warps
constexpr int ST_THREADS = (ST_WARPS + 2) * 32;  // + MMA issuer + loader
constexpr int ST_MAX_STAGES = 4;
constexpr int KC = tc::KB;
constexpr uint32_t A_PART = ST_BM * KC * 2;  // 16 KB

struct StreamParams {
  CUtensorMap tmA;  // A as a 2-D fp16 tensor (K inner, M rows), box = 64 k x 128 rows, 128-byte swizzle
  const __nv_bfloat16 *Wp;
  const float *bias;
  __half *Y;
  int ldy, M, N, K, n_chunks, BN, n_sub, relu, n_tiles, n_stages;
};

// GELU: the exact (erf) GELU of RoBERTa's feed-forward block instead of ReLU — its own instantiation (erff() in the
// shared epilogue loop slows the plain kernel; see gemm_tc.cu)
template <bool GELU>
__global__ void __launch_bounds__(ST_THREADS, 1) linear_stream_kernel(const __grid_constant__ StreamParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) unsigned long long bar_full[ST_MAX_STAGES], bar_empty[ST_MAX_STAGES], bar_acc, bar_free;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[512];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  bd::pdl_launch_dependents();
  const int ng = blockIdx.y;
  const int BN = p.BN, n_sub = p.n_sub, NC = n_sub * BN, S = p.n_stages;
  const uint32_t w_blk = static_cast<uint32_t>(BN) * KC * 2, w_bytes = w_blk * n_sub;
  const uint32_t stage_bytes = A_PART + w_bytes;
  const uint32_t ncols = tc::tmem_cols_pow2(NC);
  __half *tile = reinterpret_cast<__half *>(smem + S * stage_bytes);  // 128 rows x (NC + 8) halfs
  const int ldt = NC + 8;

  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), ncols);
  if (tid == 32) {
    for (int i = 0; i < ST_MAX_STAGES; ++i) {
      tc::mbar_init(tc::smem_u32(&bar_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&bar_empty[i]), 1);
    }
    tc::mbar_init(tc::smem_u32(&bar_acc), 1);
    tc::mbar_init(tc::smem_u32(&bar_free), ST_WARPS);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < NC; i += ST_THREADS) {
    const int col = ng * NC + i;
    bias_s[i] = (p.bias && col < p.N) ? __ldg(p.bias + col) : 0.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, tmem_base_s, 0);

  if (warp == ST_WARPS + 1) {
    // ---------------------------------------------------------------------------------- loader
    if (tc::elect_one()) {
      bool waited = false;
      uint32_t g = 0;  // chunks issued so far (over all tiles of this CTA)
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        for (int c = 0; c < p.n_chunks; ++c, ++g) {
          const uint32_t st = g % S;
          if (g >= static_cast<uint32_t>(S)) tc::mbar_wait(tc::smem_u32(&bar_empty[st]), ((g / S) - 1) & 1);
          const uint32_t bar = tc::smem_u32(&bar_full[st]);
          tc::mbar_arrive_expect_tx(bar, stage_bytes);
          tc::bulk_g2s(tc::smem_u32(smem + st * stage_bytes + A_PART),
                       p.Wp + (static_cast<size_t>(ng) * p.n_chunks + c) * (w_bytes / 2), w_bytes, bar);
          if (!waited) {  // the weights do not depend on the preceding kernel, the activations do
            bd::pdl_wait();
            waited = true;
          }
          // rows >= M and k >= K are zero-filled by the TMA unit and count towards the 16 KB
          tc::tma_load_2d(tc::smem_u32(smem + st * stage_bytes), &p.tmA, c * KC, t * ST_BM, bar);
        }
      }
    }
    __syncwarp();
  } else if (warp == ST_WARPS) {
    // ------------------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = tc::idesc_ab(1, ST_BM, BN);
    uint32_t g = 0, it = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      if (it > 0) {  // the epilogue has read the previous tile's accumulators out of TMEM
        tc::mbar_wait(tc::smem_u32(&bar_free), (it - 1) & 1);
        tc::fence_after_sync();
      }
      for (int c = 0; c < p.n_chunks; ++c, ++g) {
        const uint32_t st = g % S;
        tc::mbar_wait(tc::smem_u32(&bar_full[st]), (g / S) & 1);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t a0 = tc::smem_u32(smem + st * stage_bytes), w0 = a0 + A_PART;
#pragma unroll
          for (int s = 0; s < KC / 16; ++s) {
            const uint64_t da = tc::smem_desc_sw128(a0 + s * 32);
            const uint32_t acc = (c > 0 || s > 0) ? 1u : 0u;
            for (int sub = 0; sub < n_sub; ++sub)
              tc::mma_bf16(tmem + sub * BN, da, tc::smem_desc_sw128(w0 + sub * w_blk + s * 32), idesc, acc);
          }
          tc::mma_commit(tc::smem_u32(&bar_empty[st]));
          if (c == p.n_chunks - 1) tc::mma_commit(tc::smem_u32(&bar_acc));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------------------- epilogue
    bd::pdl_wait();  // Y may still be read by the preceding kernels of the stream
    const int r = (warp & 3) * 32 + lane;  // accumulator row of this thread = TMEM lane
    const uint32_t tbase = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int n_groups = NC / 16;
    const int per = (n_groups + 1) / 2;
    const int g0 = (warp >> 2) * per, g1 = min(n_groups, g0 + per);
    const int col_base = ng * NC;
    const int n_valid = min(NC, p.N - col_base);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      tc::mbar_wait(tc::smem_u32(&bar_acc), it & 1);
      tc::fence_after_sync();
      if (it > 0) {  // the previous tile's bulk stores have read the shared tile
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      for (int g = g0; g < g1; g += 2) {
        uint32_t acc[2][16];
        tc::tmem_ld16(tbase + g * 16, acc[0]);
        if (g + 1 < g1) tc::tmem_ld16(tbase + (g + 1) * 16, acc[1]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (g + u >= g1) break;
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = __uint_as_float(acc[u][j]) + bias_s[(g + u) * 16 + j];
            if (GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
            o[j] = v;
          }
          uint4 *dst = reinterpret_cast<uint4 *>(tile + r * ldt + (g + u) * 16);
          if (!GELU && p.relu) {
            dst[0] = make_uint4(tc::pack_f16x2_relu(o[0], o[1]), tc::pack_f16x2_relu(o[2], o[3]), tc::pack_f16x2_relu(o[4], o[5]), tc::pack_f16x2_relu(o[6], o[7]));
            dst[1] = make_uint4(tc::pack_f16x2_relu(o[8], o[9]), tc::pack_f16x2_relu(o[10], o[11]), tc::pack_f16x2_relu(o[12], o[13]), tc::pack_f16x2_relu(o[14], o[15]));
          } else {
            dst[0] = make_uint4(tc::pack_f16x2(o[0], o[1]), tc::pack_f16x2(o[2], o[3]), tc::pack_f16x2(o[4], o[5]), tc::pack_f16x2(o[6], o[7]));
            dst[1] = make_uint4(tc::pack_f16x2(o[8], o[9]), tc::pack_f16x2(o[10], o[11]), tc::pack_f16x2(o[12], o[13]), tc::pack_f16x2(o[14], o[15]));
          }
        }
      }
      // accumulators drained: the MMAs of the next tile may overwrite them
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_free));
      tc::fence_proxy_async_smem();  // the tile is read by bulk stores (async proxy)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int row = t * ST_BM + tid;
      if (tid < ST_BM && row < p.M) {
        __half *y = p.Y + static_cast<long long>(row) * p.ldy + col_base;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y), "r"(tc::smem_u32(tile + tid * ldt)),
                     "r"(n_valid * 2)
                     : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

// ---------------------------------------------------------------------------------------------------
// Persistent  Y = LayerNorm(R + A16 · Wᵀ + bias) * gamma + beta  (+ fp16 copy of Y): the out-projection / FFN
// second layer + residual + LayerNorm block every attention and FFN site ends with.  Same roles as above; a CTA
// owns complete rows (one column group, N <= 320).  The fp32 tile of 128 x 288 values (150 KB) would not fit
// beside the ring, so the epilogue works on the two 64-row halves of the accumulators in turn: the four warps
// that own the half's TMEM lanes drain it (+bias) into a 64-row shared tile, then all eight warps normalise it,
// one warp per row, residual rows prefetched one round ahead, results written straight to global memory (fp32
// rows + the fp16 copy the next projection reads).  TMEM is released after the second drain, so the next tile's
// MMAs run under the second half's LayerNorm and the operand copies never stop.  Arithmetic identical to
// linear_tc_kernel<1, 0>: bit-identical results.
struct StreamLnParams {
  CUtensorMap tmA;
  const __nv_bfloat16 *Wp;
  const float *bias, *R, *gamma, *beta;
  float *Y;
  __half *Y16;
  int ldr, ldy, ldy16, M, N, K, n_chunks, BN, n_sub, n_tiles, n_stages;
  float eps;
};

__global__ void __launch_bounds__(ST_THREADS, 1) linear_ln_stream_kernel(const __grid_constant__ StreamLnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) unsigned long long bar_full[ST_MAX_STAGES], bar_empty[ST_MAX_STAGES], bar_acc, bar_free;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[512];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  bd::pdl_launch_dependents();
  const int BN = p.BN, n_sub = p.n_sub, NC = n_sub * BN, S = p.n_stages;
  const uint32_t w_blk = static_cast<uint32_t>(BN) * KC * 2, w_bytes = w_blk * n_sub;
  const uint32_t stage_bytes = A_PART + w_bytes;
  const uint32_t ncols = tc::tmem_cols_pow2(NC);
  float *tile = reinterpret_cast<float *>(smem + S * stage_bytes);  // 64 rows x (NC + 4) floats
  const int ldt = NC + 4;

  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), ncols);
  if (tid == 32) {
    for (int i = 0; i < ST_MAX_STAGES; ++i) {
      tc::mbar_init(tc::smem_u32(&bar_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&bar_empty[i]), 1);
    }
    tc::mbar_init(tc::smem_u32(&bar_acc), 1);
    tc::mbar_init(tc::smem_u32(&bar_free), ST_WARPS);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < NC; i += ST_THREADS) bias_s[i] = (p.bias && i < p.N) ? __ldg(p.bias + i) : 0.f;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, tmem_base_s, 0);

  if (warp == ST_WARPS + 1) {
    // ---------------------------------------------------------------------------------- loader
    if (tc::elect_one()) {
      bool waited = false;
      uint32_t g = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        for (int c = 0; c < p.n_chunks; ++c, ++g) {
          const uint32_t st = g % S;
          if (g >= static_cast<uint32_t>(S)) tc::mbar_wait(tc::smem_u32(&bar_empty[st]), ((g / S) - 1) & 1);
          const uint32_t bar = tc::smem_u32(&bar_full[st]);
          tc::mbar_arrive_expect_tx(bar, stage_bytes);
          tc::bulk_g2s(tc::smem_u32(smem + st * stage_bytes + A_PART), p.Wp + static_cast<size_t>(c) * (w_bytes / 2), w_bytes, bar);
          if (!waited) {
            bd::pdl_wait();
            waited = true;
          }
          tc::tma_load_2d(tc::smem_u32(smem + st * stage_bytes), &p.tmA, c * KC, t * ST_BM, bar);
        }
      }
    }
    __syncwarp();
  } else if (warp == ST_WARPS) {
    // ------------------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = tc::idesc_ab(1, ST_BM, BN);
    uint32_t g = 0, it = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      if (it > 0) {
        tc::mbar_wait(tc::smem_u32(&bar_free), (it - 1) & 1);
        tc::fence_after_sync();
      }
      for (int c = 0; c < p.n_chunks; ++c, ++g) {
        const uint32_t st = g % S;
        tc::mbar_wait(tc::smem_u32(&bar_full[st]), (g / S) & 1);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t a0 = tc::smem_u32(smem + st * stage_bytes), w0 = a0 + A_PART;
#pragma unroll
          for (int s = 0; s < KC / 16; ++s) {
            const uint64_t da = tc::smem_desc_sw128(a0 + s * 32);
            const uint32_t acc = (c > 0 || s > 0) ? 1u : 0u;
            for (int sub = 0; sub < n_sub; ++sub)
              tc::mma_bf16(tmem + sub * BN, da, tc::smem_desc_sw128(w0 + sub * w_blk + s * 32), idesc, acc);
          }
          tc::mma_commit(tc::smem_u32(&bar_empty[st]));
          if (c == p.n_chunks - 1) tc::mma_commit(tc::smem_u32(&bar_acc));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------------------- epilogue
    bd::pdl_wait();  // the residual rows come from the preceding kernels; Y / Y16 may still be read by them
    const int r = (warp & 3) * 32 + lane;  // accumulator row of this thread = TMEM lane
    const int my_half = (warp & 3) >> 1;   // rows 0..63 / 64..127
    const uint32_t tbase = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int n_groups = NC / 16;
    const int per = (n_groups + 1) / 2;
    const int g0 = (warp >> 2) * per, g1 = min(n_groups, g0 + per);
    constexpr int LN_ROWS = 4, LN_V = 3;  // rows per warp and round; float4 groups per lane (N <= 320 -> at most 3)
    const int nv = p.N >> 2;
    const float inv_n = 1.0f / static_cast<float>(p.N);
    float4 gam[LN_V], bet[LN_V], xa[LN_ROWS][LN_V], xb[LN_ROWS][LN_V];
#pragma unroll
    for (int i = 0; i < LN_V; ++i) {
      const int v = lane + i * 32;
      gam[i] = v < nv ? __ldg(reinterpret_cast<const float4 *>(p.gamma) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      bet[i] = v < nv ? __ldg(reinterpret_cast<const float4 *>(p.beta) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // round q of a tile (q = 0..3): rows 64 (q >> 1) + warp + ((q & 1) * LN_ROWS + u) * 8 of the tile
    auto ln_load = [&](int row0, int q, float4 (&x)[LN_ROWS][LN_V]) {
#pragma unroll
      for (int u = 0; u < LN_ROWS; ++u) {
        const int gr = row0 + 64 * (q >> 1) + warp + ((q & 1) * LN_ROWS + u) * ST_WARPS;
        const bool rok = gr < p.M;
        const float4 *res = reinterpret_cast<const float4 *>(p.R + (rok ? static_cast<long long>(gr) * p.ldr : 0));
#pragma unroll
        for (int i = 0; i < LN_V; ++i) {
          const int v = lane + i * 32;
          x[u][i] = (rok && v < nv) ? __ldg(res + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    auto ln_rows = [&](int row0, int q, float4 (&x)[LN_ROWS][LN_V]) {
      float sum[LN_ROWS], sq[LN_ROWS];
#pragma unroll
      for (int u = 0; u < LN_ROWS; ++u) {
        const int rl = warp + ((q & 1) * LN_ROWS + u) * ST_WARPS;  // row inside the 64-row tile
        const float4 *t4 = reinterpret_cast<const float4 *>(tile + rl * ldt);
        sum[u] = sq[u] = 0.f;
#pragma unroll
        for (int i = 0; i < LN_V; ++i) {
          const int v = lane + i * 32;
          if (v < nv) {
            const float4 tv = t4[v];
            x[u][i].x += tv.x, x[u][i].y += tv.y, x[u][i].z += tv.z, x[u][i].w += tv.w;
          }
          sum[u] += (x[u][i].x + x[u][i].y) + (x[u][i].z + x[u][i].w);
          sq[u] = fmaf(x[u][i].x, x[u][i].x, fmaf(x[u][i].y, x[u][i].y, fmaf(x[u][i].z, x[u][i].z, fmaf(x[u][i].w, x[u][i].w, sq[u]))));
        }
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int u = 0; u < LN_ROWS; ++u) {
          sum[u] += __shfl_xor_sync(0xFFFFFFFFu, sum[u], off);
          sq[u] += __shfl_xor_sync(0xFFFFFFFFu, sq[u], off);
        }
      }
#pragma unroll
      for (int u = 0; u < LN_ROWS; ++u) {
        const int gr = row0 + 64 * (q >> 1) + warp + ((q & 1) * LN_ROWS + u) * ST_WARPS;
        if (gr >= p.M) continue;
        const float mean = sum[u] * inv_n;
        const float rstd = rsqrtf(fmaxf(sq[u] * inv_n - mean * mean, 0.f) + p.eps);
        float4 *y = reinterpret_cast<float4 *>(p.Y + static_cast<long long>(gr) * p.ldy);
#pragma unroll
        for (int i = 0; i < LN_V; ++i) {
          const int v = lane + i * 32;
          if (v < nv) {
            const float4 o4 = make_float4((x[u][i].x - mean) * rstd * gam[i].x + bet[i].x, (x[u][i].y - mean) * rstd * gam[i].y + bet[i].y,
                                          (x[u][i].z - mean) * rstd * gam[i].z + bet[i].z, (x[u][i].w - mean) * rstd * gam[i].w + bet[i].w);
            y[v] = o4;
            if (p.Y16)
              reinterpret_cast<uint2 *>(p.Y16 + static_cast<long long>(gr) * p.ldy16)[v] =
                  make_uint2(tc::pack_f16x2(o4.x, o4.y), tc::pack_f16x2(o4.z, o4.w));
          }
        }
      }
    };
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      const int row0 = t * ST_BM;
      ln_load(row0, 0, xa);  // requested before the accumulators are ready: the latency is off the path
      tc::mbar_wait(tc::smem_u32(&bar_acc), it & 1);
      tc::fence_after_sync();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (my_half == h) {  // drain this half's TMEM lanes: +bias -> shared tile (row stride NC + 4 floats)
          for (int g = g0; g < g1; g += 2) {
            uint32_t acc[2][16];
            tc::tmem_ld16(tbase + g * 16, acc[0]);
            if (g + 1 < g1) tc::tmem_ld16(tbase + (g + 1) * 16, acc[1]);
            tc::tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (g + u >= g1) break;
              float o[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(acc[u][j]) + bias_s[(g + u) * 16 + j];
              float4 *dst = reinterpret_cast<float4 *>(tile + (r - 64 * h) * ldt + (g + u) * 16);
              dst[0] = make_float4(o[0], o[1], o[2], o[3]);
              dst[1] = make_float4(o[4], o[5], o[6], o[7]);
              dst[2] = make_float4(o[8], o[9], o[10], o[11]);
              dst[3] = make_float4(o[12], o[13], o[14], o[15]);
            }
          }
          // this warp's accumulator rows are out of TMEM (the eighth arrival releases the accumulators)
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_free));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the half's tile is complete
        ln_load(row0, 2 * h + 1, xb);
        ln_rows(row0, 2 * h, xa);
        if (h == 0) ln_load(row0, 2, xa);
        ln_rows(row0, 2 * h + 1, xb);
        asm volatile("bar.sync 1, 256;" ::: "memory");  // every warp has read the tile: the next drain may overwrite it
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

int g_stream_on = 1;

}  // namespace

extern "C" int bd_linear_stream_set(int on) {
  g_stream_on = on != 0;
  return BD_OK;
}

// Used by bd_linear_tc_h when the call qualifies (returns 1 = launched, 0 = not applicable, < 0 = error code
// negated): fp16 A rows (lda % 8 == 0, 16-byte aligned), fp16 Y rows (ldy % 8 == 0, 16-byte aligned), no second
// operand, every column group a multiple of 8 columns wide, and more row tiles than one wave of SMs would hold
// (below that a CTA has one tile and nothing to overlap).
int bd_linear_stream_try(const void *A, int lda, const void *Wp, const float *bias, void *Y, int ldy, int M, int N, int K,
                         int n_chunks, int BN, int n_sub, int relu, cudaStream_t stream) {
  if (!g_stream_on || relu > 2 || relu < 0) return 0;
  const int NC = n_sub * BN, n_groups = bd::ceil_div(N, NC), n_tiles = bd::ceil_div(M, ST_BM);
  const int n_sm = bd::sm_count();
  if (NC > 512 || N % 8 != 0 || NC % 8 != 0 || n_groups > 65535) return 0;
  if (static_cast<long long>(n_tiles) * n_groups <= n_sm) return 0;
  const uint32_t stage = A_PART + static_cast<uint32_t>(NC) * KC * 2;
  const size_t tile = static_cast<size_t>(ST_BM) * (NC + 8) * 2;
  int stages = static_cast<int>((217 * 1024 - 1024 - tile) / stage);
  if (stages > ST_MAX_STAGES) stages = ST_MAX_STAGES;
  if (stages < 2) return 0;
  StreamParams p = {};
  tc::EncodeTiledFn enc = tc::encode_tiled();
  if (!enc) return 0;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(M)};
  const cuuint32_t box[2] = {KC, ST_BM}, estr[2] = {1, 1};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(lda) * sizeof(__half)};
  if (enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 0;
  p.Wp = static_cast<const __nv_bfloat16 *>(Wp), p.bias = bias, p.Y = static_cast<__half *>(Y), p.ldy = ldy;
  p.M = M, p.N = N, p.K = K, p.n_chunks = n_chunks, p.BN = BN, p.n_sub = n_sub, p.relu = relu, p.n_tiles = n_tiles;
  p.n_stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage + tile + 1024;
  static bd::PerDeviceOnce configured;
  if (configured.run([&]() {
        cudaError_t e = cudaFuncSetAttribute(linear_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(linear_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024);
        return e;
      }) != cudaSuccess)
    return 0;
  int ctas = n_sm / n_groups;
  if (ctas < 1) ctas = 1;
  if (ctas > n_tiles) ctas = n_tiles;
  const cudaError_t e = relu == 2 ? bd::launch_pdl(linear_stream_kernel<true>, dim3(ctas, n_groups), dim3(ST_THREADS), smem, stream, p)
                                  : bd::launch_pdl(linear_stream_kernel<false>, dim3(ctas, n_groups), dim3(ST_THREADS), smem, stream, p);
  if (e != cudaSuccess) {
    bd::set_error("bd_linear_tc_h (persistent): %s", cudaGetErrorString(e));
    return -BD_ERR_CUDA;
  }
  return 1;
}

// Used by bd_linear_ln_tc_h (same return convention as bd_linear_stream_try): fp16 A rows, one column group
// (complete rows), more row tiles than SMs.
int bd_linear_ln_stream_try(const void *A, int lda, const void *Wp, const float *bias, const float *R, int ldr,
                            const float *gamma, const float *beta, float eps, float *Y, int ldy, void *Y16, int ldy16, int M,
                            int N, int K, int n_chunks, int BN, int n_sub, cudaStream_t stream) {
  if (!g_stream_on) return 0;
  const int NC = n_sub * BN, n_tiles = bd::ceil_div(M, ST_BM);
  const int n_sm = bd::sm_count();
  if (NC > 512 || NC < N || N > 320 || N % 4 != 0 || n_tiles <= n_sm) return 0;
  const uint32_t stage = A_PART + static_cast<uint32_t>(NC) * KC * 2;
  const size_t tile = static_cast<size_t>(64) * (NC + 4) * 4;
  int stages = static_cast<int>((217 * 1024 - 1024 - tile) / stage);
  if (stages > ST_MAX_STAGES) stages = ST_MAX_STAGES;
  if (stages < 2) return 0;
  StreamLnParams p = {};
  tc::EncodeTiledFn enc = tc::encode_tiled();
  if (!enc) return 0;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(M)};
  const cuuint32_t box[2] = {KC, ST_BM}, estr[2] = {1, 1};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(lda) * sizeof(__half)};
  if (enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 0;
  p.Wp = static_cast<const __nv_bfloat16 *>(Wp), p.bias = bias, p.R = R, p.gamma = gamma, p.beta = beta, p.eps = eps;
  p.Y = Y, p.Y16 = static_cast<__half *>(Y16), p.ldr = ldr, p.ldy = ldy, p.ldy16 = ldy16;
  p.M = M, p.N = N, p.K = K, p.n_chunks = n_chunks, p.BN = BN, p.n_sub = n_sub, p.n_tiles = n_tiles, p.n_stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage + tile + 1024;
  static bd::PerDeviceOnce configured;
  if (configured.run([&]() {
        return cudaFuncSetAttribute(linear_ln_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024);
      }) != cudaSuccess)
    return 0;
  const int ctas = n_tiles < n_sm ? n_tiles : n_sm;
  const cudaError_t e = bd::launch_pdl(linear_ln_stream_kernel, dim3(ctas), dim3(ST_THREADS), smem, stream, p);
  if (e != cudaSuccess) {
    bd::set_error("bd_linear_ln_tc_h (persistent): %s", cudaGetErrorString(e));
    return -BD_ERR_CUDA;
  }
  return 1;
}
