// One set-abstraction level in ONE kernel on tcgen05:
//     QueryAndGroup (gather + centre + 1/radius)  ->  SharedMLP (3 x [1x1 conv + folded BN + ReLU])
//     ->  max-pool over nsample
// replacing pointnet2_utils.py:334-359 + pytorch_utils.py:25-36 + pointnet2_modules.py:243-257 of
// the reference (QueryAndGroup -> SharedMLP -> F.max_pool2d), which materialise the grouped tensor
// and both hidden activations in HBM (1.07 GB per hidden layer of SA1 at 32 scenes).
//
// A CTA owns 128 grouped rows (= 128 / nsample centres) and carries them through all three layers:
//   layer 1   A = gathered rows, converted to bf16 hi / lo while staged into a 2-stage ring
//             (128-byte-swizzle K-major tiles, tc_common.cuh); accumulator in TMEM
//   epilogue  TMEM -> +bias, ReLU -> bf16 hi / lo -> written straight into shared memory AS THE
//             A OPERAND of the next layer (same swizzled layout): hidden activations never leave
//             the SM
//   layer 2, epilogue, layer 3 likewise; last epilogue = +bias, ReLU, max over the nsample rows of
//             each centre, one pooled row per centre to HBM.
// Weights of the three layers stream through one 2-stage ring of cp.async.bulk copies (pre-packed
// by the host exactly like bd_linear_tc's, full_rows tiling); the blocks of layer L+1 are requested
// when the accumulator of layer L completes, so they arrive under its epilogue.
// Roles: warps 0-7 stage / run the epilogues, warp 8 issues the MMAs (issue blocks on the tensor
// queue, so it must not be a staging warp).  bf16x3: D += Ahi*Whi + Alo*Whi + Ahi*Wlo.
// One row tile per CTA on purpose.  A persistent variant (barriers / TMEM / biases set up once, the next tile's
// neighbour indices and first gathers in flight under the current tile's epilogues, chunk counters running on
// across tiles) was built and measured at 148 scenes: SA2 2.34 -> 2.94 ms, SA3 0.76 -> 0.99, SA4 0.38 -> 0.48 —
// the prefetched plan + gather registers live across the epilogues push the 96-register budget of two CTAs per
// SM into 272 bytes of spills, and the block scheduler staggers independent CTAs better than two lock-stepped
// persistent ones.  Dropped.
#include "tc_common.cuh"

namespace {

constexpr int SA_BM = 128;
constexpr int SA_WARPS = 8;
constexpr int SA_THREADS = (SA_WARPS + 1) * 32;
constexpr int SA_ITEMS = 4;
constexpr int KC = tc::KB;
constexpr uint32_t A_PART = SA_BM * KC * 2;  // 16 KB: one 64-wide block of 128 rows

struct SaLayer {
  const __nv_bfloat16 *Wp;
  const float *bias;
  int N, BN, n_sub, n_chunks;  // n_chunks k-chunks of 64; chunk = parts * n_sub * BN * 128 bytes
};

struct SaMlpParams {
  const int *idx;
  const float *feat, *xyz, *cen;
  float *Y;
  int ldf, ldx, C, ns, n, m, ldy;
  int M, K1, split;
  float inv_r;
  SaLayer L[3];
  // 16-bit feature rows between the levels (fp16 mode): feat16 = `feat` points at fp16 rows (ldf in halfs, C % 8 == 0:
  // a 16-byte chunk of 8 features is copied into the operand tile as it is — no conversion, half the gather
  // bytes); Y16 = the pooled rows are ALSO written as fp16 (the next level's gather source)
  int feat16, ldy16;
  __half *Y16;
  uint32_t w_stage;  // bytes of one weight-ring stage (largest chunk)
  uint32_t x_bytes;  // bytes of the A-ring / hidden-activation region
};

// Last layer with the operands SWAPPED (D^T = W2 . H^T: accumulator lane = output channel, accumulator column =
// grouped row): the max over the nsample rows of a centre is then a thread-local loop over TMEM columns — no
// cross-lane reduction, no shared tile, no barrier — and a warp writes 32 consecutive channels of one pooled row.
// Columns [c_begin, c_begin + c_count) of the accumulator at `tacc` (this thread's lane), nsample in {16,32,64,128}.
__device__ __forceinline__ void pool_swapped(uint32_t tacc, int c_begin, int c_count, int ns, float bias, long long row0,
                                             int M, float *__restrict__ Y, int ldy, int ch, __half *__restrict__ Y16 = nullptr,
                                             int ldy16 = 0) {
  float m = -INFINITY;
  for (int g = 0; g < c_count / 16; g += 2) {
    uint32_t a0[16], a1[16];
    tc::tmem_ld16(tacc + c_begin + g * 16, a0);
    tc::tmem_ld16(tacc + c_begin + (g + 1) * 16, a1);
    tc::tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t(&a)[16] = u ? a1 : a0;
      float m0 = fmaxf(__uint_as_float(a[0]), __uint_as_float(a[1])), m1 = fmaxf(__uint_as_float(a[2]), __uint_as_float(a[3]));
      float m2 = fmaxf(__uint_as_float(a[4]), __uint_as_float(a[5])), m3 = fmaxf(__uint_as_float(a[6]), __uint_as_float(a[7]));
      m0 = fmaxf(m0, fmaxf(__uint_as_float(a[8]), __uint_as_float(a[9]))), m1 = fmaxf(m1, fmaxf(__uint_as_float(a[10]), __uint_as_float(a[11])));
      m2 = fmaxf(m2, fmaxf(__uint_as_float(a[12]), __uint_as_float(a[13]))), m3 = fmaxf(m3, fmaxf(__uint_as_float(a[14]), __uint_as_float(a[15])));
      m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
      const int col_end = c_begin + (g + u + 1) * 16;  // columns consumed so far (exclusive)
      if (col_end % ns == 0) {  // a centre's rows are complete: max_r relu(x_r + b) = relu(max_r x_r + b)
        const long long orow = (row0 + col_end) / ns - 1;
        if (orow * ns < M) {
          const float o = fmaxf(m + bias, 0.f);
          Y[orow * ldy + ch] = o;
          if (Y16) Y16[orow * ldy16 + ch] = __float2half_rn(fminf(o, 65504.f));
        }
        m = -INFINITY;
      }
    }
  }
}

template <int PARTS>
__global__ void __launch_bounds__(SA_THREADS, PARTS == 1 ? 2 : 1) sa_mlp_tc_kernel(const SaMlpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sX = smem;              // layer 1: A ring (2 stages); layers 2, 3: hidden activations
  unsigned char *sW = smem + p.x_bytes;  // 2 weight stages
  __shared__ __align__(8) unsigned long long bar_w[2], bar_a[2], bar_mma[2], bar_acc, bar_h;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[3][256];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const int row0 = blockIdx.x * SA_BM;
  const int n1 = p.L[0].n_chunks;
  const int nmax = max(p.L[0].N, max(p.L[1].N, p.L[2].N));
  const uint32_t ncols = tc::tmem_cols_pow2(nmax);
  // last layer with swapped operands (pool_swapped): output widths 128 / 256, nsample 16 .. 128 (64 when one
  // accumulator's columns are split between the two warpgroups)
  const bool swap2 = p.L[2].BN == SA_BM && p.ns >= 16 && (p.L[2].n_sub == 2 || p.ns <= 64);

  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), ncols);
  if (tid == 32) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(tc::smem_u32(&bar_w[i]), 1);
      tc::mbar_init(tc::smem_u32(&bar_a[i]), SA_WARPS);
      tc::mbar_init(tc::smem_u32(&bar_mma[i]), 1);
    }
    tc::mbar_init(tc::smem_u32(&bar_acc), 1);
    tc::mbar_init(tc::smem_u32(&bar_h), SA_WARPS);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < 3 * 256; i += SA_THREADS) {
    const int l = i >> 8, c = i & 255;
    bias_s[l][c] = (p.L[l].bias && c < p.L[l].N) ? __ldg(p.L[l].bias + c) : 0.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, tmem_base_s, 0);
  bd::pdl_launch_dependents();

  // weight chunk t (global numbering over the three layers) -> ring stage t & 1
  auto issue_w = [&](int l, int c, int t) {  // one thread
    const SaLayer &L = p.L[l];
    const uint32_t bytes = static_cast<uint32_t>(PARTS) * L.n_sub * L.BN * 128u;
    const uint32_t bar = tc::smem_u32(&bar_w[t & 1]);
    tc::mbar_arrive_expect_tx(bar, bytes);
    tc::bulk_g2s(tc::smem_u32(sW + (t & 1) * p.w_stage), L.Wp + static_cast<size_t>(c) * (bytes / 2), bytes, bar);
  };

  if (warp == SA_WARPS) {
    // ------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      issue_w(0, 0, 0);
      if (n1 > 1) issue_w(0, 1, 1);
    }
    int t = 0;
    for (int l = 0; l < 3; ++l) {
      const SaLayer &L = p.L[l];
      const uint32_t idesc = tc::idesc_ab(PARTS, SA_BM, L.BN);
      const uint32_t w_blk = static_cast<uint32_t>(L.BN) * 128u;
      const int k_total = l == 0 ? p.K1 : p.L[l - 1].N;  // valid K of this layer
      const bool swapped = l == 2 && swap2;  // D^T = W . H^T (BN = 128 = SA_BM: same instruction shape)
      if (l > 0) {  // hidden activations of the previous layer are in shared memory (and TMEM was drained)
        tc::mbar_wait(tc::smem_u32(&bar_h), (l - 1) & 1);
      }
      for (int c = 0; c < L.n_chunks; ++c, ++t) {
        const int st = t & 1;
        const uint32_t par = (t >> 1) & 1;
        tc::mbar_wait(tc::smem_u32(&bar_w[st]), par);
        if (l == 0) tc::mbar_wait(tc::smem_u32(&bar_a[st]), par);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          // A operand: layer 1 -> ring stage; layers 2, 3 -> block c of the hidden activations
          const uint32_t a0 = tc::smem_u32(sX) + (l == 0 ? st * (PARTS * A_PART) : c * A_PART);
          const uint32_t a_lo = l == 0 ? A_PART : static_cast<uint32_t>(L.n_chunks) * A_PART;
          const uint32_t w0 = tc::smem_u32(sW + st * p.w_stage);
          const int ksteps = min(KC / 16, (k_total - c * KC + 15) / 16);
          for (int s = 0; s < ksteps; ++s) {
            const uint64_t da_hi = tc::smem_desc_sw128(a0 + s * 32);
            const uint64_t da_lo = tc::smem_desc_sw128(a0 + a_lo + s * 32);
            const uint32_t acc = (c > 0 || s > 0) ? 1u : 0u;
            for (int sub = 0; sub < L.n_sub; ++sub) {
              const uint32_t d = tmem + sub * L.BN;
              const uint64_t dw_hi = tc::smem_desc_sw128(w0 + sub * w_blk + s * 32);
              if (swapped) tc::mma_bf16(d, dw_hi, da_hi, idesc, acc);
              else tc::mma_bf16(d, da_hi, dw_hi, idesc, acc);
              if (PARTS == 2) {
                const uint64_t dw_lo = tc::smem_desc_sw128(w0 + (L.n_sub + sub) * w_blk + s * 32);
                if (swapped) {
                  tc::mma_bf16(d, dw_hi, da_lo, idesc, 1u);
                  tc::mma_bf16(d, dw_lo, da_hi, idesc, 1u);
                } else {
                  tc::mma_bf16(d, da_lo, dw_hi, idesc, 1u);
                  tc::mma_bf16(d, da_hi, dw_lo, idesc, 1u);
                }
              }
            }
          }
          tc::mma_commit(tc::smem_u32(&bar_mma[st]));
          if (c == L.n_chunks - 1) tc::mma_commit(tc::smem_u32(&bar_acc));
        }
        __syncwarp();
      }
    }
  } else {
    // -------------------------------------------------------------- staging / epilogue warps
    bd::pdl_wait();  // neighbour indices, features and centres come from the preceding kernels
    // staging plan (same for every k-chunk): a warp-item covers 8 rows x 4 chunks of 8 k
    long long f_off[SA_ITEMS], x_off[SA_ITEMS], c_off[SA_ITEMS];
    uint32_t s_off[SA_ITEMS];
    int k_off[SA_ITEMS];
    bool row_ok[SA_ITEMS];
#pragma unroll
    for (int it = 0; it < SA_ITEMS; ++it) {
      const int blk = warp + it * SA_WARPS;  // 0..31
      const int r = (blk >> 1) * 8 + (lane & 7), ch = (blk & 1) * 4 + (lane >> 3);
      row_ok[it] = row0 + r < p.M;
      k_off[it] = ch * 8;
      s_off[it] = tc::sw128_off(r, ch);
      const long long gr = row_ok[it] ? row0 + r : 0;
      const int a = __ldg(p.idx + gr);
      const long long bj = gr / p.ns, b = bj / p.m;
      f_off[it] = (b * p.n + a) * p.ldf;
      x_off[it] = (b * p.n + a) * p.ldx;
      c_off[it] = bj * 3;
    }
    const bool feat_vec = ((p.ldf | p.C) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.feat) & 15) == 0;
    float4 ra0[SA_ITEMS][2], ra1[SA_ITEMS][2];
    auto issue_loads = [&](int c, float4 (&ra)[SA_ITEMS][2]) {
#pragma unroll
      for (int it = 0; it < SA_ITEMS; ++it) {
        const int k0 = c * KC + k_off[it];
        const bool ok = row_ok[it] && k0 < p.K1;
        const float *f = p.feat + f_off[it];
        const __half *fh = reinterpret_cast<const __half *>(p.feat) + f_off[it];
        if (p.feat16 && k0 + 8 <= p.C) {  // 8 fp16 features: one 16-byte load, kept as bits (zeros for a dead row)
          ra[it][0] = ok ? __ldg(reinterpret_cast<const float4 *>(fh + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (ok && k0 + 8 <= p.C && feat_vec) {  // chunk entirely inside the feature row
          ra[it][0] = __ldg(reinterpret_cast<const float4 *>(f + k0));
          ra[it][1] = __ldg(reinterpret_cast<const float4 *>(f + k0) + 1);
        } else {  // chunk straddles features / relative xyz / padding
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = k0 + i;
            float x = 0.f;
            if (ok && k < p.C) x = p.feat16 ? __half2float(fh[k]) : __ldg(f + k);
            else if (ok && k < p.C + 3)
              x = __fmul_rn(__fsub_rn(__ldg(p.xyz + x_off[it] + (k - p.C)), __ldg(p.cen + c_off[it] + (k - p.C))), p.inv_r);
            v[i] = x;
          }
          ra[it][0] = make_float4(v[0], v[1], v[2], v[3]);
          ra[it][1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    };
    issue_loads(0, ra0);
    if (n1 > 1) issue_loads(1, ra1);
    auto step = [&](int c, float4 (&ra)[SA_ITEMS][2]) {
      const int st = c & 1;
      unsigned char *sA = sX + st * (PARTS * A_PART);
      if (c >= 2) {  // stage (A and W halves) consumed by chunk c - 2
        tc::mbar_wait(tc::smem_u32(&bar_mma[st]), ((c >> 1) - 1) & 1);
        if (tid == 0) issue_w(0, c, c);
      }
      const int kend = min(KC, ((p.K1 - c * KC + 15) / 16) * 16);  // k-steps the MMAs will read
#pragma unroll
      for (int it = 0; it < SA_ITEMS; ++it) {
        if (k_off[it] >= kend) continue;
        float v[8] = {ra[it][0].x, ra[it][0].y, ra[it][0].z, ra[it][0].w, ra[it][1].x, ra[it][1].y, ra[it][1].z, ra[it][1].w};
        if (p.feat16 && c * KC + k_off[it] + 8 <= p.C) {  // the registers hold 8 fp16 features as loaded
          if (PARTS == 1) {  // already the operand format
            *reinterpret_cast<float4 *>(sA + s_off[it]) = ra[it][0];
            continue;
          }
          const __half2 *h2 = reinterpret_cast<const __half2 *>(&ra[it][0]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f2 = __half22float2(h2[i]);
            v[2 * i] = f2.x, v[2 * i + 1] = f2.y;
          }
        }
        uint4 hi, lo;
        tc::cvt8(PARTS, v, hi, lo);
        *reinterpret_cast<uint4 *>(sA + s_off[it]) = hi;
        if (PARTS == 2) *reinterpret_cast<uint4 *>(sA + A_PART + s_off[it]) = lo;
      }
      if (c + 2 < n1) issue_loads(c + 2, ra);
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_a[st]));
    };
    for (int c = 0; c < n1; c += 2) {
      step(c, ra0);
      if (c + 1 < n1) step(c + 1, ra1);
    }

    const int r = (warp & 3) * 32 + lane;  // accumulator row of this thread = TMEM lane
    const uint32_t tbase = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    int t_next = n1;  // global index of the next layer's first weight chunk
    for (int l = 0; l < 3; ++l) {
      const SaLayer &L = p.L[l];
      tc::mbar_wait(tc::smem_u32(&bar_acc), l & 1);  // accumulator of layer l complete; every stage is free
      tc::fence_after_sync();
      if (l < 2 && tid == 0) {  // next layer's weights arrive under this epilogue (at most 2 chunks: K <= 128)
        for (int c = 0; c < p.L[l + 1].n_chunks; ++c) issue_w(l + 1, c, t_next + c);
      }
      if (l < 2) t_next += p.L[l + 1].n_chunks;
      const int n_groups = L.N / 16;
      const int per = (n_groups + 1) / 2;
      const int g0 = (warp >> 2) * per, g1 = min(n_groups, g0 + per);
      if (l < 2) {
        // ---- hidden layer: +bias, ReLU, split, store as the next layer's A operand (K = L.N)
        const uint32_t h_lo = static_cast<uint32_t>(L.N / KC) * A_PART;
        for (int g = g0; g < g1; g += 2) {
          uint32_t acc[2][16];
          tc::tmem_ld16(tbase + g * 16, acc[0]);
          if (g + 1 < g1) tc::tmem_ld16(tbase + (g + 1) * 16, acc[1]);
          tc::tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (g + u >= g1) break;
            const int k0 = (g + u) * 16;
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[u][h8 * 8 + j]) + bias_s[l][k0 + h8 * 8 + j];
              uint4 hi, lo;
              tc::cvt8_relu(PARTS, v, hi, lo);  // ReLU inside the conversion
              const int k = k0 + h8 * 8;
              const uint32_t off = static_cast<uint32_t>(k / KC) * A_PART + tc::sw128_off(r, (k % KC) / 8);
              *reinterpret_cast<uint4 *>(sX + off) = hi;
              if (PARTS == 2) *reinterpret_cast<uint4 *>(sX + h_lo + off) = lo;
            }
          }
        }
        tc::fence_before_sync();  // TMEM reads done before the next layer's MMAs overwrite the accumulator
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_h));
      } else if (swap2) {
        // ---- last layer, swapped operands: this thread's accumulator lane is output channel `ch`
        const int sub = L.n_sub == 2 ? (warp >> 2) : 0;
        const int ch = sub * SA_BM + r;
        const int c_begin = L.n_sub == 2 ? 0 : (warp >> 2) * 64, c_count = L.n_sub == 2 ? SA_BM : 64;
        pool_swapped(tbase + sub * SA_BM, c_begin, c_count, p.ns, bias_s[l][ch], row0, p.M, p.Y, p.ldy, ch, p.Y16, p.ldy16);
      } else {
        // ---- last layer: +bias, ReLU -> shared tile -> max over the nsample rows of each centre,
        //      in column passes of at most 128 (tile = 128 x 132 floats: small enough for two CTAs
        //      per SM in the single-part mode).  The tile reuses the A / W regions: every MMA and
        //      weight copy has completed.  (Measured alternative: warp reductions of the accumulators —
        //      CREDUX.MAX.F32, one per column and warp — are SLOWER here: SA2 +6 %, SA3 / SA4 (half-warp
        //      masks) +26 %; the reduction instruction is the scarce resource, the tile pass is not.)
        const int PW = min(L.N, 128);  // columns per pass
        const int ldt = PW + 4;
        float *tile = reinterpret_cast<float *>(smem);
        const int pg = PW / 16 / 2;  // 16-column groups per warpgroup and pass
        for (int c0 = 0; c0 < L.N; c0 += PW) {
          for (int g = 0; g < pg; g += 2) {
            const int col = (warp >> 2) * (PW / 2) + g * 16;  // column inside the pass
            uint32_t acc[2][16];
            tc::tmem_ld16(tbase + c0 + col, acc[0]);
            if (g + 1 < pg) tc::tmem_ld16(tbase + c0 + col + 16, acc[1]);
            tc::tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (g + u >= pg) break;
              float o[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = fmaxf(__uint_as_float(acc[u][j]) + bias_s[l][c0 + col + u * 16 + j], 0.f);
              float4 *dst = reinterpret_cast<float4 *>(tile + r * ldt + col + u * 16);
              dst[0] = make_float4(o[0], o[1], o[2], o[3]);
              dst[1] = make_float4(o[4], o[5], o[6], o[7]);
              dst[2] = make_float4(o[8], o[9], o[10], o[11]);
              dst[3] = make_float4(o[12], o[13], o[14], o[15]);
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int groups = SA_BM / p.ns;
          for (int e = tid; e < groups * PW; e += SA_WARPS * 32) {
            const int g = e / PW, col = e - g * PW;
            const long long orow = static_cast<long long>(row0) / p.ns + g;
            if (orow * p.ns >= p.M) continue;
            const float *tt = tile + (g * p.ns) * ldt + col;
            float mx = tt[0];
            for (int q = 1; q < p.ns; ++q) mx = fmaxf(mx, tt[q * ldt]);
            p.Y[orow * p.ldy + c0 + col] = mx;
            if (p.Y16) p.Y16[orow * p.ldy16 + c0 + col] = __float2half_rn(fminf(mx, 65504.f));
          }
          if (c0 + PW < L.N) asm volatile("bar.sync 1, 256;" ::: "memory");  // tile free for the next pass
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}


// ------------------------------------------------------------------ resident-weights variant
// SA1 (6 input channels -> 64 -> 64 -> 128, nsample 64): all three weight matrices fit in shared
// memory (64 KB as bf16 hi + lo), so a PERSISTENT CTA loads them once and walks over its row
// tiles: per tile only 128 x 24 bytes of gathered points come in and 2 x 128 pooled floats go out.
// Two CTAs per SM overlap each other's epilogues and MMAs.  The max over nsample rows is a
// warp REDUX on the (non-negative, hence integer-ordered) ReLU outputs — no shared tile.
//   K1 <= 16 (one MMA k-step), N0 = N1 = 64, N2 in {64, 128}, nsample in {32, 64}.
template <int PARTS>
__global__ void __launch_bounds__(SA_THREADS, PARTS == 1 ? 3 : 2) sa_mlp_resident_kernel(const SaMlpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int N2 = p.L[2].N;
  const uint32_t w0_bytes = PARTS * 64u * 128u, w2_bytes = PARTS * static_cast<uint32_t>(N2) * 128u;
  unsigned char *sW0 = smem, *sW1 = sW0 + w0_bytes, *sW2 = sW1 + w0_bytes;
  unsigned char *sX = sW2 + w2_bytes;  // A of layer 1, then hidden activations (PARTS x 16 KB)
  __shared__ __align__(8) unsigned long long bar_w, bar_a, bar_h, bar_acc;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[3][128];
  __shared__ float pool_s[4][128];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const int n_tiles = (p.M + SA_BM - 1) / SA_BM;
  const bool swap2 = N2 == SA_BM;  // last layer with swapped operands: thread-local pooling (pool_swapped); nsample 32 / 64

  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), 128);
  if (tid == 32) {
    tc::mbar_init(tc::smem_u32(&bar_w), 1);
    tc::mbar_init(tc::smem_u32(&bar_a), SA_WARPS);
    tc::mbar_init(tc::smem_u32(&bar_h), SA_WARPS);
    tc::mbar_init(tc::smem_u32(&bar_acc), 1);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < 3 * 128; i += SA_THREADS) {
    const int l = i >> 7, c = i & 127;
    bias_s[l][c] = (p.L[l].bias && c < p.L[l].N) ? __ldg(p.L[l].bias + c) : 0.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, tmem_base_s, 0);
  bd::pdl_launch_dependents();

  if (warp == SA_WARPS) {
    // ------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t bar = tc::smem_u32(&bar_w);
      tc::mbar_arrive_expect_tx(bar, 2 * w0_bytes + w2_bytes);
      tc::bulk_g2s(tc::smem_u32(sW0), p.L[0].Wp, w0_bytes, bar);
      tc::bulk_g2s(tc::smem_u32(sW1), p.L[1].Wp, w0_bytes, bar);
      tc::bulk_g2s(tc::smem_u32(sW2), p.L[2].Wp, w2_bytes, bar);
    }
    tc::mbar_wait(tc::smem_u32(&bar_w), 0);
    const uint32_t idesc64 = tc::idesc_ab(PARTS, SA_BM, 64), idesc2 = tc::idesc_ab(PARTS, SA_BM, N2);
    const uint32_t a0 = tc::smem_u32(sX);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      for (int l = 0; l < 3; ++l) {
        if (l == 0) tc::mbar_wait(tc::smem_u32(&bar_a), it & 1);
        else tc::mbar_wait(tc::smem_u32(&bar_h), (2 * it + l - 1) & 1);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t w = tc::smem_u32(l == 0 ? sW0 : l == 1 ? sW1 : sW2);
          const uint32_t w_lo = l == 2 ? static_cast<uint32_t>(N2) * 128u : 64u * 128u;
          const uint32_t idesc = l == 2 ? idesc2 : idesc64;
          const int ksteps = l == 0 ? 1 : 4;
          for (int s = 0; s < ksteps; ++s) {
            const uint64_t da_hi = tc::smem_desc_sw128(a0 + s * 32), da_lo = tc::smem_desc_sw128(a0 + A_PART + s * 32);
            const uint64_t dw_hi = tc::smem_desc_sw128(w + s * 32), dw_lo = tc::smem_desc_sw128(w + w_lo + s * 32);
            if (l == 2 && swap2) {  // D^T = W2 . H^T (N2 = 128: same instruction shape)
              tc::mma_bf16(tmem, dw_hi, da_hi, idesc, s > 0 ? 1u : 0u);
              if (PARTS == 2) {
                tc::mma_bf16(tmem, dw_hi, da_lo, idesc, 1u);
                tc::mma_bf16(tmem, dw_lo, da_hi, idesc, 1u);
              }
              continue;
            }
            tc::mma_bf16(tmem, da_hi, dw_hi, idesc, s > 0 ? 1u : 0u);
            if (PARTS == 2) {
              tc::mma_bf16(tmem, da_lo, dw_hi, idesc, 1u);
              tc::mma_bf16(tmem, da_hi, dw_lo, idesc, 1u);
            }
          }
          tc::mma_commit(tc::smem_u32(&bar_acc));
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------------- gather / epilogues
    // thread t < 128 gathers row t (k 0..7 = [features | relative xyz | 0]); t >= 128 writes the
    // zero chunk k 8..15 of row t - 128 (the k-step is 16 wide)
    bd::pdl_wait();  // neighbour indices and centres come from the preceding kernels
    const int r = tid & 127;
    const bool gatherer = tid < 128;
    const uint32_t a_off = tc::sw128_off(r, gatherer ? 0 : 1);
    float v[8];
    auto prefetch = [&](int tile) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      const long long gr = static_cast<long long>(tile) * SA_BM + r;
      if (!gatherer || tile >= n_tiles || gr >= p.M) return;
      const int a = __ldg(p.idx + gr);
      const long long bj = gr / p.ns, b = bj / p.m;
      const float *f = p.feat + (b * p.n + a) * p.ldf, *x = p.xyz + (b * p.n + a) * p.ldx, *c = p.cen + bj * 3;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < p.C) v[i] = __ldg(f + i);
        else if (i < p.C + 3) v[i] = __fmul_rn(__fsub_rn(__ldg(x + (i - p.C)), __ldg(c + (i - p.C))), p.inv_r);
      }
    };
    const int row = (warp & 3) * 32 + lane;  // accumulator row of this thread = TMEM lane
    const int half = warp >> 2;              // the two warpgroups split the columns
    const uint32_t tbase = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int per_tile = SA_BM / p.ns;       // centres per tile
    prefetch(blockIdx.x);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      {  // layer-1 operand (the previous tile's layer-3 MMAs, which read this region, were waited for)
        uint4 hi, lo;
        tc::cvt8(PARTS, v, hi, lo);
        *reinterpret_cast<uint4 *>(sX + a_off) = hi;
        if (PARTS == 2) *reinterpret_cast<uint4 *>(sX + A_PART + a_off) = lo;
        tc::fence_before_sync();  // this thread's TMEM reads of the previous tile are complete
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_a));
      }
      prefetch(tile + gridDim.x);  // in flight during this tile's three layers
      for (int l = 0; l < 2; ++l) {
        tc::mbar_wait(tc::smem_u32(&bar_acc), (3 * it + l) & 1);
        tc::fence_after_sync();
        uint32_t acc[2][16];
        tc::tmem_ld16(tbase + half * 32, acc[0]);
        tc::tmem_ld16(tbase + half * 32 + 16, acc[1]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const int k = half * 32 + u * 16 + h8 * 8;
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(acc[u][h8 * 8 + j]) + bias_s[l][k + j];
            uint4 hi, lo;
            tc::cvt8_relu(PARTS, o, hi, lo);  // ReLU inside the conversion
            const uint32_t off = tc::sw128_off(row, k / 8);
            *reinterpret_cast<uint4 *>(sX + off) = hi;
            if (PARTS == 2) *reinterpret_cast<uint4 *>(sX + A_PART + off) = lo;
          }
        }
        tc::fence_before_sync();
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_h));
      }
      // ---- last layer: +bias, ReLU, max over the rows of each centre
      tc::mbar_wait(tc::smem_u32(&bar_acc), (3 * it + 2) & 1);
      tc::fence_after_sync();
      if (swap2) {  // lane = output channel; this warpgroup's 64 accumulator columns = its centre(s)
        pool_swapped(tbase, half * 64, 64, p.ns, bias_s[2][row], static_cast<long long>(tile) * SA_BM, p.M, p.Y, p.ldy, row,
                     p.Y16, p.ldy16);
        continue;  // (the next tile's layer-1 operand store is ordered after these TMEM reads by its fence)
      }
      const int ncol = N2 / 2;  // columns of this warpgroup
      for (int g = 0; g < ncol / 16; ++g) {
        uint32_t acc[16];
        tc::tmem_ld16(tbase + half * ncol + g * 16, acc);
        tc::tmem_ld_wait();
        // max over the warp's 32 rows of the RAW accumulators (CREDUX.MAX.F32), bias and ReLU on the one survivor:
        // max_r relu(x_r + b) = relu(max_r x_r + b) exactly (fp32 addition is monotonic)
        float mine = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#ifdef SA1_NO_POOL
          const float mx = __uint_as_float(acc[j]);
#else
          const float mx = tc::redux_max_f32(__uint_as_float(acc[j]));
#endif
          if (lane == j) mine = mx;
        }
        if (lane < 16) pool_s[warp & 3][half * ncol + g * 16 + lane] = fmaxf(mine + bias_s[2][half * ncol + g * 16 + lane], 0.f);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int e = tid; e < per_tile * N2; e += SA_WARPS * 32) {
        const int g = e / N2, col = e - g * N2;
        const long long orow = static_cast<long long>(tile) * per_tile + g;
        if (orow * p.ns < p.M) {
          const int w0 = g * (p.ns / 32);
          float mx = pool_s[w0][col];
          if (p.ns == 64) mx = fmaxf(mx, pool_s[w0 + 1][col]);
          p.Y[orow * p.ldy + col] = mx;
          if (p.Y16) p.Y16[orow * p.ldy16 + col] = __float2half_rn(fminf(mx, 65504.f));
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // pool_s free again
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

}  // namespace

// Fused set-abstraction level (see the header of this file).  idx (B,m,ns) neighbour indices,
// feats (B,n,C) token-major rows (ld_feats), xyz (B,n,3) rows (ld_xyz), new_xyz (B,m,3) centres.
// Layer l: packed weights Wp[l] (pack_weight_tc, full_rows tiling: BN[l] x n_sub[l] = N[l]),
// bias[l]; layer 1's K columns ordered [features | xyz | 0].  Y (B*m, N[2]) pooled rows.
static int sa_mlp_tc_impl(const int *idx, const float *feats, int ld_feats, int C, const float *xyz, int ld_xyz,
                          const float *new_xyz, int B, int n, int m, int ns, float radius, const void *Wp0,
                          const float *b0, int N0, const void *Wp1, const float *b1, int N1, const void *Wp2,
                          const float *b2, int N2, float *Y, int ldy, int split, int feats_half, void *Y16, int ldy16,
                          bd_stream_t stream);
extern "C" int bd_sa_mlp_tc(const int *idx, const float *feats, int ld_feats, int C, const float *xyz, int ld_xyz,
                            const float *new_xyz, int B, int n, int m, int ns, float radius, const void *Wp0,
                            const float *b0, int N0, const void *Wp1, const float *b1, int N1, const void *Wp2,
                            const float *b2, int N2, float *Y, int ldy, int split, bd_stream_t stream) {
  return sa_mlp_tc_impl(idx, feats, ld_feats, C, xyz, ld_xyz, new_xyz, B, n, m, ns, radius, Wp0, b0, N0, Wp1, b1, N1, Wp2, b2,
                        N2, Y, ldy, split, 0, nullptr, 0, stream);
}
// bd_sa_mlp_tc with 16-bit feature rows between the levels: feats_half != 0 -> `feats` are fp16 rows (ld_feats in
// halfs, C % 8 == 0, ld_feats % 8 == 0, 16-byte aligned); Y16 != NULL -> the pooled rows are also written as fp16
// (ldy16 halfs per row).  Same values: the fp32 path rounds the gathered features to the operand format anyway.
extern "C" int bd_sa_mlp_tc_h(const int *idx, const void *feats, int ld_feats, int C, int feats_half, const float *xyz,
                              int ld_xyz, const float *new_xyz, int B, int n, int m, int ns, float radius, const void *Wp0,
                              const float *b0, int N0, const void *Wp1, const float *b1, int N1, const void *Wp2,
                              const float *b2, int N2, float *Y, int ldy, void *Y16, int ldy16, int split,
                              bd_stream_t stream) {
  BD_REQUIRE(!feats_half || (C > 0 && C % 8 == 0 && ld_feats % 8 == 0 && (reinterpret_cast<uintptr_t>(feats) & 15) == 0),
             "bd_sa_mlp_tc_h: fp16 feature rows need C % 8 == 0, ld_feats % 8 == 0 and 16-byte alignment");
  BD_REQUIRE(!Y16 || ldy16 >= N2, "bd_sa_mlp_tc_h: ldy16 < N2");
  return sa_mlp_tc_impl(idx, static_cast<const float *>(feats), ld_feats, C, xyz, ld_xyz, new_xyz, B, n, m, ns, radius, Wp0, b0,
                        N0, Wp1, b1, N1, Wp2, b2, N2, Y, ldy, split, feats_half, Y16, ldy16, stream);
}
static int sa_mlp_tc_impl(const int *idx, const float *feats, int ld_feats, int C, const float *xyz, int ld_xyz,
                          const float *new_xyz, int B, int n, int m, int ns, float radius, const void *Wp0,
                          const float *b0, int N0, const void *Wp1, const float *b1, int N1, const void *Wp2,
                          const float *b2, int N2, float *Y, int ldy, int split, int feats_half, void *Y16, int ldy16,
                          bd_stream_t stream) {
  BD_REQUIRE(idx && xyz && new_xyz && Wp0 && Wp1 && Wp2 && Y && (feats || C == 0), "bd_sa_mlp_tc: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && m > 0 && ns > 0 && C >= 0 && ld_xyz >= 3 && ld_feats >= C && ldy >= N2,
             "bd_sa_mlp_tc: bad sizes");
  BD_REQUIRE(SA_BM % ns == 0, "bd_sa_mlp_tc: nsample must divide 128");
  BD_REQUIRE((N0 == 64 || N0 == 128) && (N1 == 64 || N1 == 128) && (N2 == 64 || N2 == 128 || N2 == 256),
             "bd_sa_mlp_tc: layer widths must be 64 / 128 (hidden) and 64 / 128 / 256 (output)");
  BD_REQUIRE(split == 1 || split == 3, "bd_sa_mlp_tc: split must be 1 (bf16) or 3 (bf16x3)");
  BD_REQUIRE(static_cast<long long>(B) * m * ns < (1LL << 31), "bd_sa_mlp_tc: too many rows");
  const int parts = split == 3 ? 2 : 1;
  SaMlpParams p = {};
  p.idx = idx, p.feat = feats ? feats : xyz, p.xyz = xyz, p.cen = new_xyz, p.Y = Y;
  p.ldf = ld_feats, p.ldx = ld_xyz, p.C = C, p.ns = ns, p.n = n, p.m = m, p.ldy = ldy;
  p.M = B * m * ns, p.K1 = (C + 3 + 7) / 8 * 8, p.split = split, p.inv_r = 1.0f / radius;
  p.feat16 = feats_half ? 1 : 0, p.Y16 = static_cast<__half *>(Y16), p.ldy16 = ldy16;
  const void *W[3] = {Wp0, Wp1, Wp2};
  const float *bs[3] = {b0, b1, b2};
  const int N[3] = {N0, N1, N2};
  const int K[3] = {p.K1, N0, N1};
  uint32_t w_stage = 0;
  for (int l = 0; l < 3; ++l) {
    SaLayer &L = p.L[l];
    L.Wp = static_cast<const __nv_bfloat16 *>(W[l]), L.bias = bs[l], L.N = N[l];
    L.n_sub = N[l] > 160 ? 2 : 1, L.BN = N[l] / L.n_sub;  // = tc_tiling(N, K, full_rows=True)
    L.n_chunks = (K[l] + KC - 1) / KC;
    const uint32_t chunk = static_cast<uint32_t>(parts) * L.n_sub * L.BN * 128u;
    if (chunk > w_stage) w_stage = chunk;
  }
  p.w_stage = w_stage;
  if (p.K1 <= 16 && N0 == 64 && N1 == 64 && N2 <= 128 && (ns == 32 || ns == 64) && !feats_half) {
    // resident-weights persistent kernel (SA1)
    const size_t smem_r = static_cast<size_t>(parts) * (2 * 64 + N2 + SA_BM) * 128 + 1024;
    static bd::PerDeviceOnce configured_r;  // function attributes are per device
    BD_CUDA(configured_r.run([&]() {
      cudaError_t e = cudaFuncSetAttribute(sa_mlp_resident_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(sa_mlp_resident_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
      return e;
    }), "bd_sa_mlp_tc");
    const int n_sm = bd::sm_count();
    const int tiles = bd::ceil_div(p.M, SA_BM);
    const int per_sm = parts == 1 ? 3 : 2;  // 49 KB of shared memory per CTA in the single-part mode, 97 KB otherwise
    const int ctas = tiles < per_sm * n_sm ? tiles : per_sm * n_sm;
    if (parts == 2)
      BD_CUDA(bd::launch_pdl(sa_mlp_resident_kernel<2>, dim3(ctas), dim3(SA_THREADS), smem_r, bd::as_stream(stream), p), "bd_sa_mlp_tc");
    else
      BD_CUDA(bd::launch_pdl(sa_mlp_resident_kernel<1>, dim3(ctas), dim3(SA_THREADS), smem_r, bd::as_stream(stream), p), "bd_sa_mlp_tc");
    BD_CHECK_LAUNCH("bd_sa_mlp_tc");
    return BD_OK;
  }
  const uint32_t a_ring = (p.L[0].n_chunks > 1 ? 2u : 1u) * parts * A_PART;
  const uint32_t hidden = static_cast<uint32_t>(parts) * (static_cast<uint32_t>(N0 > N1 ? N0 : N1) / KC) * A_PART;
  p.x_bytes = a_ring > hidden ? a_ring : hidden;
  const size_t pipe = static_cast<size_t>(p.x_bytes) + 2u * w_stage;
  const size_t tile = static_cast<size_t>(SA_BM) * ((N2 < 128 ? N2 : 128) + 4) * 4;  // pooled in passes of <= 128 columns
  const size_t smem = (pipe > tile ? pipe : tile) + 1024;
  BD_REQUIRE(smem <= 218 * 1024, "bd_sa_mlp_tc: needs %zu bytes of shared memory (> 218 KB)", smem);
  static bd::PerDeviceOnce configured;  // function attributes are per device
  BD_CUDA(configured.run([&]() {
    cudaError_t e = cudaFuncSetAttribute(sa_mlp_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sa_mlp_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024);
    return e;
  }), "bd_sa_mlp_tc");
  const dim3 grid(bd::ceil_div(p.M, SA_BM));
  if (parts == 2)
    BD_CUDA(bd::launch_pdl(sa_mlp_tc_kernel<2>, grid, dim3(SA_THREADS), smem, bd::as_stream(stream), p), "bd_sa_mlp_tc");
  else
    BD_CUDA(bd::launch_pdl(sa_mlp_tc_kernel<1>, grid, dim3(SA_THREADS), smem, bd::as_stream(stream), p), "bd_sa_mlp_tc");
  BD_CHECK_LAUNCH("bd_sa_mlp_tc");
  return BD_OK;
}
