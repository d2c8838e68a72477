// Tensor-core linear layer on tcgen05:
//     Y = act((A [+ A2]) · Wᵀ + bias)                       (plain epilogue)
//     Y = LayerNorm(R + (A [+ A2]) · Wᵀ + bias) * g + b      (fused residual + LayerNorm epilogue)
//
//   A  : (M, K) fp32 row-major activations (token-major), K % 8 == 0, 16-byte aligned rows.  They
//        are converted to bf16 — a hi part, or hi + lo parts in the "bf16x3" mode — while being
//        staged into shared memory in the 128-byte-swizzle K-major layout (tc_common.cuh).  A2
//        (positional embedding) is added during staging.  Global loads are software-pipelined
//        two k-chunks ahead through two register sets.
//   Wp : weights PRE-PACKED by the host into that same layout, one contiguous block per
//        (CTA column group, k-chunk of 64): Wp[ng][kc][part][sub][BN rows][64 k] bf16 (swizzled,
//        zero padded), so a block arrives with ONE cp.async.bulk (TMA engine) on an mbarrier.
//   D  : n_sub accumulators of 128 x BN fp32 in TMEM (tcgen05.mma cta_group::1, M = 128, K = 16
//        per instruction, single issuing thread).  bf16x3: D += Ahi*Whi + Alo*Whi + Ahi*Wlo.
//   Pipeline: ring of up to 4 shared-memory stages; the MMAs of chunk c run asynchronously
//        (tcgen05.commit -> mbarrier) while the CTA stages the following chunks.
//   Epilogue: tcgen05.ld -> (+bias, ReLU) -> shared-memory tile -> warp-per-row coalesced stores,
//        or, for the fused variant (CTA owns complete rows, N <= 320), + residual -> LayerNorm.
//
// Replaces the cuBLAS / cuDNN-1x1-conv + BN + ReLU / LayerNorm call sites of the reference.
#include "tc_common.cuh"

namespace {

int g_tc_two_per_sm = 1;  // bd_linear_tc_set_occupancy()
constexpr int TC_BM = 128;
constexpr int TC_WARPS = 8;                       // producer / epilogue warps
constexpr int TC_THREADS = (TC_WARPS + 2) * 32;  // + one MMA-issuing warp + one loader warp
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_ITEMS = 4;  // staged items (8 consecutive k of one row) per thread and k-chunk
constexpr int KC = tc::KB;   // k-chunk = one 64-element swizzle block
constexpr uint32_t A_PART = TC_BM * KC * 2;  // 16 KB

struct LinearTcParams {
  CUtensorMap tmA, tmA2;  // A (and A2) as 2-D tensors (K inner, M rows), box = 64 k x 128 rows, no swizzle
  const float *A, *A2, *bias, *R, *gamma, *beta;
  const __nv_bfloat16 *Wp;
  float *Y;
  int lda, lda2, ldy, ldr;
  int M, N, K, n_chunks, BN, n_sub, relu, split, n_stages;
  float eps;
  // gather mode (QueryAndGroup fused into the A staging): row r = (scene b, centre j, sample s),
  // A[r, :] = [ feats[b, idx[r], 0:C] | (xyz[b, idx[r]] - centre[b, j]) * inv_radius | 0 ... ]
  const int *g_idx;
  const float *g_feat, *g_xyz, *g_cen;
  int g_ldf, g_ldx, g_C, g_ns, g_n, g_m;
  float g_inv_r;
  int pool;  // > 0: max over groups of `pool` consecutive rows in the epilogue (F.max_pool2d over nsample)
  // 16-bit activations in HBM (fp16 mode, plain A): a16 = A is an fp16 (M, K) matrix and arrives by ONE tensor
  // copy per chunk directly in the swizzled operand layout (no conversion pass, half the bytes); y16 = the
  // plain epilogue writes fp16 rows to Y16 INSTEAD of fp32 rows to Y; the LayerNorm epilogue writes its fp32
  // rows AND, when Y16 is set, an fp16 copy (the operand of the next projection)
  int a16, y16, ldy16;
  __half *Y16;
  long long *dbg;  // optional clock64() stamps of CTA (0,0), thread 0 (tuning aid)
};

#define TC_STAMP_T(i, t)                                                                          \
  do {                                                                                            \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == (t)) p.dbg[i] = clock64();  \
  } while (0)
#define TC_STAMP(i)                                                                               \
  do {                                                                                            \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.dbg[i] = clock64();    \
  } while (0)

// MODE: 0 = A, 1 = A + A2, 2 = gathered rows (QueryAndGroup).  EPI: 0 = bias/ReLU, 1 = residual + LayerNorm,
// 2 = bias/ReLU + max-pool over groups of rows.
// GELU: the plain epilogue applies the exact (erf) GELU instead of ReLU — its own instantiation, because erff() inside
// the shared epilogue loop costs every other linear of the forward (measured: 7.8 % of the whole step).
template <int EPI, int MODE, bool GELU = false>
__global__ void __launch_bounds__(TC_THREADS, (MODE == 0 && EPI != 1) ? 2 : 1) linear_tc_kernel(const __grid_constant__ LinearTcParams p) {
  constexpr bool LN_EPI = EPI == 1;
  constexpr bool HAS_A2 = MODE == 1;
  constexpr bool GATHER = MODE == 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // swizzle-128B tiles need 1024-byte aligned bases (in the shared address space)
  unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) unsigned long long bar_w[TC_MAX_STAGES], bar_a[TC_MAX_STAGES], bar_mma[TC_MAX_STAGES], bar_raw[TC_MAX_STAGES], bar_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);  // provably warp-uniform
  bd::pdl_launch_dependents();  // the next kernel may start its prologue under this one
  TC_STAMP(0);
  const int row0 = blockIdx.x * TC_BM;
  const int ng = blockIdx.y;  // column group: n_sub consecutive BN-wide tiles
  const int BN = p.BN, n_sub = p.n_sub;
  const uint32_t parts = p.split == 3 ? 2u : 1u;
  const uint32_t w_blk = static_cast<uint32_t>(BN) * KC * 2;  // multiple of 1024 (BN % 8 == 0)
  // A (and A2) arrive as raw fp32 chunks by tensor copy and are converted in place; the gathered
  // variant stages through registers (LSU)
  constexpr bool TMA_A = MODE == 0 || (MODE == 1 && EPI != 1);  // (LayerNorm + A2: full-row weights leave no room)
  constexpr uint32_t RAW = 2 * A_PART;  // one raw fp32 chunk: 128 rows x 256 B
  const bool A16 = MODE == 0 && p.a16;  // fp16 activations: the tensor copy delivers the operand tile itself
  const uint32_t a_bytes = A16 ? A_PART : (TMA_A ? (HAS_A2 ? 2 * RAW : RAW) : A_PART * parts), w_bytes = w_blk * parts * n_sub;
  const uint32_t stage_bytes = a_bytes + w_bytes;
  const uint32_t ncols = tc::tmem_cols_pow2(n_sub * BN);

  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), ncols);
  if (tid == 32) {
    for (int i = 0; i < TC_MAX_STAGES; ++i) {
      tc::mbar_init(tc::smem_u32(&bar_w[i]), 1);
      tc::mbar_init(tc::smem_u32(&bar_a[i]), TC_WARPS);
      tc::mbar_init(tc::smem_u32(&bar_mma[i]), 1);
      tc::mbar_init(tc::smem_u32(&bar_raw[i]), 1);
    }
    tc::mbar_init(tc::smem_u32(&bar_done), 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  TC_STAMP(1);
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, tmem_base_s, 0);
  const uint32_t idesc = tc::idesc_ab(static_cast<int>(parts), TC_BM, BN);
  __shared__ float bias_s[512];

  // Per-thread staging plan, identical for every k-chunk.  A warp-item covers 8 rows x 4 chunks
  // (lane = chunk_local * 8 + row_local): 128 contiguous bytes per row from global memory and
  // conflict-free 16-byte st.shared into the swizzled tile.
  if (GATHER) bd::pdl_wait();  // the neighbour indices read below come from the preceding kernel
  long long g_off[TC_ITEMS], g2_off[TC_ITEMS], s_cen[TC_ITEMS];
  uint32_t s_off[TC_ITEMS];
  int k_off[TC_ITEMS];
  bool row_ok[TC_ITEMS];
#pragma unroll
  for (int it = 0; it < TC_ITEMS; ++it) {
    const int blk = warp + it * TC_WARPS;  // 0..31
    const int r = (blk >> 1) * 8 + (lane & 7), ch = (blk & 1) * 4 + (lane >> 3);
    row_ok[it] = row0 + r < p.M;
    k_off[it] = ch * 8;
    s_off[it] = tc::sw128_off(r, ch);
    g_off[it] = static_cast<long long>(row0 + r) * p.lda + ch * 8;
    g2_off[it] = static_cast<long long>(row0 + r) * p.lda2 + ch * 8;
    if (GATHER) {  // g_off -> gathered feature row, g2_off -> gathered xyz row; centre kept in cen[]
      const long long gr = row_ok[it] ? row0 + r : 0;
      const int a = __ldg(p.g_idx + gr);
      const long long bj = gr / p.g_ns, b = bj / p.g_m;
      g_off[it] = (b * p.g_n + a) * p.g_ldf;
      g2_off[it] = (b * p.g_n + a) * p.g_ldx;
      s_cen[it] = bj * 3;
    }
  }
  float4 ra0[TC_ITEMS][2], rb0[TC_ITEMS][2], ra1[TC_ITEMS][2], rb1[TC_ITEMS][2];
  auto issue_loads = [&](int c, float4 (&ra)[TC_ITEMS][2], float4 (&rb)[TC_ITEMS][2]) {
#pragma unroll
    for (int it = 0; it < TC_ITEMS; ++it) {
      const bool ok = row_ok[it] && (c * KC + k_off[it] < p.K);  // K % 8 == 0: whole item in range
      if (GATHER) {
        const int k0 = c * KC + k_off[it];
        const float *f = p.g_feat + g_off[it];
        if (ok && k0 + 8 <= p.g_C && ((p.g_ldf | p.g_C) & 3) == 0) {  // chunk entirely inside the feature row
          ra[it][0] = __ldg(reinterpret_cast<const float4 *>(f + k0));
          ra[it][1] = __ldg(reinterpret_cast<const float4 *>(f + k0) + 1);
        } else {  // chunk straddles features / relative xyz / padding
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = k0 + i;
            float x = 0.f;
            if (ok && k < p.g_C) x = __ldg(f + k);
            else if (ok && k < p.g_C + 3)
              x = __fmul_rn(__fsub_rn(__ldg(p.g_xyz + g2_off[it] + (k - p.g_C)), __ldg(p.g_cen + s_cen[it] + (k - p.g_C))),
                            p.g_inv_r);
            v[i] = x;
          }
          ra[it][0] = make_float4(v[0], v[1], v[2], v[3]);
          ra[it][1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        continue;
      }
      const float4 *src = reinterpret_cast<const float4 *>(p.A + (ok ? g_off[it] + c * KC : 0));
      ra[it][0] = __ldg(src), ra[it][1] = __ldg(src + 1);
      if (HAS_A2) {
        const float4 *src2 = reinterpret_cast<const float4 *>(p.A2 + (ok ? g2_off[it] + c * KC : 0));
        rb[it][0] = __ldg(src2), rb[it][1] = __ldg(src2 + 1);
      }
    }
  };

  // S-stage ring, three roles decoupled by mbarriers (no CTA-wide barrier in the main loop):
  //   producers (warps 0-7): wait stage free -> convert + store A(c) -> one arrival per warp on
  //                          bar_a[stage]
  //   issuer (warp 8):       wait bar_w / bar_a of chunk c -> MMAs -> tcgen05.commit -> bar_mma[stage]
  //                          (= stage free).  Issue blocks while the tensor pipe drains its queue, so
  //                          it must not be a producer warp: staging of chunk c+1 overlaps MMA(c).
  const int S = p.n_stages;
  auto issue_w = [&](int c) {  // one thread
    const int st = c % S;
    tc::mbar_arrive_expect_tx(tc::smem_u32(&bar_w[st]), w_bytes);
    tc::bulk_g2s(tc::smem_u32(smem + st * stage_bytes + a_bytes),
                 p.Wp + (static_cast<size_t>(ng) * p.n_chunks + c) * (w_bytes / 2), w_bytes, tc::smem_u32(&bar_w[st]));
  };

  if (warp == TC_WARPS) {
    // ------------------------------------------------------------------------------ MMA issuer
    for (int c = 0; c < p.n_chunks; ++c) {
      const int st = c % S;
      const uint32_t par = (c / S) & 1;
      tc::mbar_wait(tc::smem_u32(&bar_w[st]), par);
      TC_STAMP_T(10 + 3 * c, TC_WARPS * 32);
      tc::mbar_wait(tc::smem_u32(A16 ? &bar_raw[st] : &bar_a[st]), par);  // fp16 A: straight from the tensor copy
      TC_STAMP_T(11 + 3 * c, TC_WARPS * 32);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t a0 = tc::smem_u32(smem + st * stage_bytes), w0 = a0 + a_bytes;
#pragma unroll
        for (int s = 0; s < KC / 16; ++s) {
          const uint64_t da_hi = tc::smem_desc_sw128(a0 + s * 32);
          const uint64_t da_lo = tc::smem_desc_sw128(a0 + A_PART + s * 32);
          const uint32_t acc = (c > 0 || s > 0) ? 1u : 0u;
          for (int sub = 0; sub < n_sub; ++sub) {
            const uint32_t d = tmem + sub * BN;
            const uint64_t dw_hi = tc::smem_desc_sw128(w0 + sub * w_blk + s * 32);
            tc::mma_bf16(d, da_hi, dw_hi, idesc, acc);
            if (parts == 2) {
              const uint64_t dw_lo = tc::smem_desc_sw128(w0 + (n_sub + sub) * w_blk + s * 32);
              tc::mma_bf16(d, da_lo, dw_hi, idesc, 1u);
              tc::mma_bf16(d, da_hi, dw_lo, idesc, 1u);
            }
          }
        }
        tc::mma_commit(tc::smem_u32(&bar_mma[st]));
        if (c == p.n_chunks - 1) tc::mma_commit(tc::smem_u32(&bar_done));
      }
      __syncwarp();

      TC_STAMP_T(12 + 3 * c, TC_WARPS * 32);
    }
  } else if (warp == TC_WARPS + 1) {
    // ---------------------------------------------------------------------------------- loader
    // weight block of every chunk, and (plain-A variants) the chunk's raw fp32 activation rows:
    // one bulk copy of <= 256 bytes per row straight into the stage's A region [128 rows x 256 B].
    // The TMA path ingests ~2x what the LSU path (global loads into registers) does per SM.
    if (lane == 0)
      for (int c = 0; c < S && c < p.n_chunks; ++c) issue_w(c);
    if (TMA_A) {
      bd::pdl_wait();  // activations come from the preceding kernels
      for (int c = 0; c < p.n_chunks; ++c) {
        const int st = c % S;
        if (c >= S) tc::mbar_wait(tc::smem_u32(&bar_mma[st]), ((c / S) - 1) & 1);
        if (lane == 0) {
          if (c >= S) issue_w(c);
          // one tensor copy per chunk: box (64 k x 128 rows) at (c * 64, row0); rows >= M and k >= K
          // are zero-filled by the TMA unit and still count towards the full 32 KB
          const uint32_t bar = tc::smem_u32(&bar_raw[st]);
          tc::mbar_arrive_expect_tx(bar, A16 ? A_PART : (HAS_A2 ? 2 * RAW : RAW));
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                  tc::smem_u32(smem + st * stage_bytes)),
              "l"(reinterpret_cast<uint64_t>(&p.tmA)), "r"(c * KC), "r"(row0), "r"(bar)
              : "memory");
          if (HAS_A2)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                    tc::smem_u32(smem + st * stage_bytes + RAW)),
                "l"(reinterpret_cast<uint64_t>(&p.tmA2)), "r"(c * KC), "r"(row0), "r"(bar)
                : "memory");
        }
        __syncwarp();
      }
    } else {
      for (int c = S; c < p.n_chunks; ++c) {  // refills (the producers only wait for the stage)
        tc::mbar_wait(tc::smem_u32(&bar_mma[c % S]), ((c / S) - 1) & 1);
        if (lane == 0) issue_w(c);
      }
    }
  } else {
    // ------------------------------------------------------------------------------- producers
    bd::pdl_wait();  // activations (A, A2, gather sources, residual) come from the preceding kernels
    if (!TMA_A) {
      issue_loads(0, ra0, rb0);
      if (p.n_chunks > 1) issue_loads(1, ra1, rb1);
    }
    TC_STAMP(2);
    // plain-A variant: read the raw chunk (lane = 16-byte piece of a row: 2 rows per warp
    // instruction, conflict-free), barrier among the producers, write the converted operand over it
    auto step_tma = [&](int c) {
      const int st = c % S;
      unsigned char *sA = smem + st * stage_bytes;
      tc::mbar_wait(tc::smem_u32(&bar_raw[st]), (c / S) & 1);
      float4 raw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 16 + 2 * i + (lane >> 4);
        raw[i] = *reinterpret_cast<const float4 *>(sA + r * (KC * 4) + (lane & 15) * 16);
        if (HAS_A2) {  // + positional embedding (second raw chunk of the stage; never overwritten)
          const float4 r2 = *reinterpret_cast<const float4 *>(sA + RAW + r * (KC * 4) + (lane & 15) * 16);
          raw[i].x += r2.x, raw[i].y += r2.y, raw[i].z += r2.z, raw[i].w += r2.w;
        }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");  // every raw value is in registers
      const int piece = lane & 15;
      const bool k_ok = c * KC + piece * 4 < p.K;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 16 + 2 * i + (lane >> 4);
        const bool ok = k_ok && row0 + r < p.M;
        const float v0 = ok ? raw[i].x : 0.f, v1 = ok ? raw[i].y : 0.f, v2 = ok ? raw[i].z : 0.f, v3 = ok ? raw[i].w : 0.f;
        const uint32_t off = tc::sw128_off(r, piece >> 1) + (piece & 1) * 8;
        if (parts == 2) {
          uint32_t h0, l0, h1, l1;
          tc::split_bf16x2(v0, v1, h0, l0);
          tc::split_bf16x2(v2, v3, h1, l1);
          *reinterpret_cast<uint2 *>(sA + off) = make_uint2(h0, h1);
          *reinterpret_cast<uint2 *>(sA + A_PART + off) = make_uint2(l0, l1);
        } else {
          *reinterpret_cast<uint2 *>(sA + off) = make_uint2(tc::pack_f16x2(v0, v1), tc::pack_f16x2(v2, v3));
        }
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_a[st]));
      TC_STAMP(4 + c);
    };
    auto step = [&](int c, float4 (&ra)[TC_ITEMS][2], float4 (&rb)[TC_ITEMS][2]) {
      const int st = c % S;
      unsigned char *sA = smem + st * stage_bytes;
      if (c >= S) {  // the stage's previous tenant, chunk c - S, must have been consumed
        tc::mbar_wait(tc::smem_u32(&bar_mma[st]), ((c / S) - 1) & 1);
        TC_STAMP(30 + c);
      }
#pragma unroll
      for (int it = 0; it < TC_ITEMS; ++it) {
        const bool ok = row_ok[it] && (c * KC + k_off[it] < p.K);
        float v[8] = {ra[it][0].x, ra[it][0].y, ra[it][0].z, ra[it][0].w, ra[it][1].x, ra[it][1].y, ra[it][1].z, ra[it][1].w};
        if (HAS_A2) {
          v[0] += rb[it][0].x, v[1] += rb[it][0].y, v[2] += rb[it][0].z, v[3] += rb[it][0].w;
          v[4] += rb[it][1].x, v[5] += rb[it][1].y, v[6] += rb[it][1].z, v[7] += rb[it][1].w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ok ? v[i] : 0.f;
        uint4 hi, lo;
        tc::cvt8(static_cast<int>(parts), v, hi, lo);
        *reinterpret_cast<uint4 *>(sA + s_off[it]) = hi;
        if (parts == 2) *reinterpret_cast<uint4 *>(sA + A_PART + s_off[it]) = lo;
      }
      if (c + 2 < p.n_chunks) issue_loads(c + 2, ra, rb);  // refill this register set
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_a[st]));
      TC_STAMP(4 + c);
    };
    if (A16) {
      // nothing to stage: the MMA warp consumes the tensor copies directly
    } else if (TMA_A) {
      for (int c = 0; c < p.n_chunks; ++c) step_tma(c);
    } else {
      for (int c = 0; c < p.n_chunks; c += 2) {
        step(c, ra0, rb0);
        if (c + 1 < p.n_chunks) step(c + 1, ra1, rb1);
      }
    }
    // bias -> shared memory while the last MMAs run
    for (int i = tid; i < n_sub * BN; i += TC_WARPS * 32) {
      const int col = ng * n_sub * BN + i;
      bias_s[i] = (p.bias && col < p.N) ? __ldg(p.bias + col) : 0.f;
    }
    const int NC = n_sub * BN;
    const int col_base = ng * NC;
    const int n_valid = min(NC, p.N - col_base);
    const int ldt = NC + 4;
    float *tile = reinterpret_cast<float *>(smem);

    // LayerNorm epilogue: one warp per row, lane owns the float4 column groups lane + 32 i
    // (N <= 320, N % 4 == 0: at most 3 per lane), LN_ROWS rows per round.  The residual rows of round
    // 0 are requested now, before the accumulator is ready, and those of round k + 1 while round
    // k is reduced: their latency is off the path.
    constexpr int LN_ROWS = 4, LN_ROUNDS = TC_BM / (TC_WARPS * LN_ROWS), LN_V = 3;
    const int nv = p.N >> 2;  // float4 groups per row
    float4 gam[LN_EPI ? LN_V : 1], bet[LN_EPI ? LN_V : 1], xa[LN_EPI ? LN_ROWS : 1][LN_V], xb[LN_EPI ? LN_ROWS : 1][LN_V];
    auto ln_load = [&](int round, float4 (&x)[LN_EPI ? LN_ROWS : 1][LN_V]) {
#pragma unroll
      for (int u = 0; u < LN_ROWS; ++u) {
        const int r = warp + (round * LN_ROWS + u) * TC_WARPS, gr = row0 + r;
        const bool rok = gr < p.M;
        const float4 *res = reinterpret_cast<const float4 *>(p.R + (rok ? static_cast<long long>(gr) * p.ldr : 0));
#pragma unroll
        for (int i = 0; i < LN_V; ++i) {
          const int v = lane + i * 32;
          x[u][i] = (rok && v < nv) ? __ldg(res + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    if (LN_EPI) {
#pragma unroll
      for (int i = 0; i < LN_V; ++i) {
        const int v = lane + i * 32;
        gam[i] = v < nv ? __ldg(reinterpret_cast<const float4 *>(p.gamma) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        bet[i] = v < nv ? __ldg(reinterpret_cast<const float4 *>(p.beta) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      ln_load(0, xa);
    }

    asm volatile("bar.sync 1, 256;" ::: "memory");  // producers only: bias_s visible
    tc::mbar_wait(tc::smem_u32(&bar_done), 0);       // every MMA (and weight copy) has retired:
    tc::fence_after_sync();                          // the pipeline stages become the output tile
    TC_STAMP(40);

    // ---- epilogue 1: TMEM -> (+bias, ReLU) -> shared tile (row stride NC + 4 floats).  Thread =
    //      accumulator row = TMEM lane; the two warpgroups split the 16-column groups.
    {
      const int r = (warp & 3) * 32 + lane;
      const uint32_t tbase = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      const int n_groups = NC / 16;
      const int per = (n_groups + 1) / 2;
      const int g0 = (warp >> 2) * per, g1 = min(n_groups, g0 + per);
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
            if (GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));  // exact (erf) GELU: RoBERTa's FFN
            else if (!LN_EPI && p.relu && !(EPI == 0 && p.y16)) v = fmaxf(v, 0.f);
            o[j] = v;
          }
          if (EPI == 0 && p.y16) {  // fp16 rows: row stride NC + 8 halfs; ReLU inside the conversion (F2FP.RELU)
            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(tile) + r * (NC + 8) + (g + u) * 16);
            if (!GELU && p.relu) {
              dst[0] = make_uint4(tc::pack_f16x2_relu(o[0], o[1]), tc::pack_f16x2_relu(o[2], o[3]), tc::pack_f16x2_relu(o[4], o[5]), tc::pack_f16x2_relu(o[6], o[7]));
              dst[1] = make_uint4(tc::pack_f16x2_relu(o[8], o[9]), tc::pack_f16x2_relu(o[10], o[11]), tc::pack_f16x2_relu(o[12], o[13]), tc::pack_f16x2_relu(o[14], o[15]));
            } else {
              dst[0] = make_uint4(tc::pack_f16x2(o[0], o[1]), tc::pack_f16x2(o[2], o[3]), tc::pack_f16x2(o[4], o[5]), tc::pack_f16x2(o[6], o[7]));
              dst[1] = make_uint4(tc::pack_f16x2(o[8], o[9]), tc::pack_f16x2(o[10], o[11]), tc::pack_f16x2(o[12], o[13]), tc::pack_f16x2(o[14], o[15]));
            }
            continue;
          }
          float4 *dst = reinterpret_cast<float4 *>(tile + r * ldt + (g + u) * 16);
          dst[0] = make_float4(o[0], o[1], o[2], o[3]);
          dst[1] = make_float4(o[4], o[5], o[6], o[7]);
          dst[2] = make_float4(o[8], o[9], o[10], o[11]);
          dst[3] = make_float4(o[12], o[13], o[14], o[15]);
        }
      }
    }
    if (EPI == 0) tc::fence_proxy_async_smem();  // the tile is read by bulk stores (async proxy)
    asm volatile("bar.sync 1, 256;" ::: "memory");
    TC_STAMP(41);

    // ---- epilogue 2: write-out
    if (EPI == 2) {
      // max over groups of `pool` consecutive rows (the nsample neighbours of one centre); M % pool == 0
      const int groups = TC_BM / p.pool;
      for (int e = tid; e < groups * n_valid; e += TC_WARPS * 32) {
        const int g = e / n_valid, col = e - g * n_valid;
        const long long orow = static_cast<long long>(row0) / p.pool + g;
        if (orow * p.pool >= p.M) continue;
        const float *t = tile + (g * p.pool) * ldt + col;
        float mx = t[0];
        for (int q = 1; q < p.pool; ++q) mx = fmaxf(mx, t[q * ldt]);
        p.Y[orow * p.ldy + col_base + col] = mx;
      }
    } else if (EPI == 0 && p.y16) {
      // fp16 rows (ldy16 % 8 == 0, 16-byte aligned Y16: checked by the host wrapper)
      const __half *t16 = reinterpret_cast<const __half *>(tile);
      if (n_valid % 8 == 0 && col_base % 8 == 0) {
        if (tid < TC_BM && row0 + tid < p.M) {
          __half *y = p.Y16 + static_cast<long long>(row0 + tid) * p.ldy16 + col_base;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y),
                       "r"(tc::smem_u32(t16 + tid * (NC + 8))), "r"(n_valid * 2)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      } else {
        for (int r = warp; r < TC_BM; r += TC_WARPS) {
          const int gr = row0 + r;
          if (gr >= p.M) break;
          __half *y = p.Y16 + static_cast<long long>(gr) * p.ldy16 + col_base;
          for (int q = lane; q < n_valid; q += 32) y[q] = t16[r * (NC + 8) + q];
        }
      }
    } else if (EPI == 0) {
      const bool vec = (p.ldy % 4 == 0) && (n_valid % 4 == 0) && (col_base % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0);
      if (vec) {
        // one bulk store (TMA engine) per row, issued by the row's thread: the warps are done
        // after the issue, the copies drain asynchronously
        if (tid < TC_BM && row0 + tid < p.M) {
          float *y = p.Y + static_cast<long long>(row0 + tid) * p.ldy + col_base;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y),
                       "r"(tc::smem_u32(tile + tid * ldt)), "r"(n_valid * 4)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      } else {
        for (int r = warp; r < TC_BM; r += TC_WARPS) {
          const int gr = row0 + r;
          if (gr >= p.M) break;
          float *y = p.Y + static_cast<long long>(gr) * p.ldy + col_base;
          const float *t = tile + r * ldt;
          for (int q = lane; q < n_valid; q += 32) y[q] = t[q];
        }
      }
    } else {
      const float inv_n = 1.0f / static_cast<float>(p.N);
      auto ln_rows = [&](int round, float4 (&x)[LN_EPI ? LN_ROWS : 1][LN_V]) {
        // sum and sum of squares of x = residual + (acc + bias) in one sweep, reduced together by
        // one round of shuffles; variance = E[x^2] - mean^2 (fp32: the rows are O(1), |mean| << 1e3)
        float sum[LN_ROWS], sq[LN_ROWS];
#pragma unroll
        for (int u = 0; u < LN_ROWS; ++u) {
          const int r = warp + (round * LN_ROWS + u) * TC_WARPS;
          const float4 *t4 = reinterpret_cast<const float4 *>(tile + r * ldt);
          sum[u] = sq[u] = 0.f;
#pragma unroll
          for (int i = 0; i < LN_V; ++i) {
            const int v = lane + i * 32;
            if (v < nv) {
              const float4 t = t4[v];
              x[u][i].x += t.x, x[u][i].y += t.y, x[u][i].z += t.z, x[u][i].w += t.w;
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
          const int r = warp + (round * LN_ROWS + u) * TC_WARPS, gr = row0 + r;
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
              if (p.Y16)  // fp16 copy: the A operand of the projections that read this row next
                reinterpret_cast<uint2 *>(p.Y16 + static_cast<long long>(gr) * p.ldy16)[v] =
                    make_uint2(tc::pack_f16x2(o4.x, o4.y), tc::pack_f16x2(o4.z, o4.w));
            }
          }
        }
      };
#pragma unroll
      for (int round = 0; round < LN_ROUNDS; round += 2) {
        if (round + 1 < LN_ROUNDS) ln_load(round + 1, xb);
        ln_rows(round, xa);
        if (round + 2 < LN_ROUNDS) ln_load(round + 2, xa);
        if (round + 1 < LN_ROUNDS) ln_rows(round + 1, xb);
      }
    }
    TC_STAMP(42);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

using tc::EncodeTiledFn;
using tc::encode_tiled;

int launch_linear_tc(LinearTcParams &p, bool ln, cudaStream_t stream) {
  if (!p.g_idx && !(ln && p.A2)) {  // A (and A2) are read through the TMA unit
    EncodeTiledFn enc = encode_tiled();
    BD_REQUIRE(enc != nullptr, "bd_linear_tc: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(p.K), static_cast<cuuint64_t>(p.M)};
    const cuuint32_t box[2] = {KC, TC_BM}, estr[2] = {1, 1};
    if (p.a16) {  // fp16 rows, 128-byte swizzle: the copy lands in the MMA operand layout
      const cuuint64_t strides[1] = {static_cast<cuuint64_t>(p.lda) * sizeof(__half)};
      const CUresult r = enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<float *>(p.A), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      BD_REQUIRE(r == CUDA_SUCCESS, "bd_linear_tc: cuTensorMapEncodeTiled (fp16) failed (%d) for M=%d K=%d ld=%d",
                 static_cast<int>(r), p.M, p.K, p.lda);
    } else
    for (int which = 0; which < (p.A2 ? 2 : 1); ++which) {
      const cuuint64_t strides[1] = {static_cast<cuuint64_t>(which ? p.lda2 : p.lda) * sizeof(float)};
      const CUresult r = enc(which ? &p.tmA2 : &p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                             const_cast<float *>(which ? p.A2 : p.A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      BD_REQUIRE(r == CUDA_SUCCESS, "bd_linear_tc: cuTensorMapEncodeTiled failed (%d) for M=%d K=%d ld=%d", static_cast<int>(r),
                 p.M, p.K, which ? p.lda2 : p.lda);
    }
  }
  const int NC = p.n_sub * p.BN;
  const uint32_t parts = p.split == 3 ? 2 : 1;
  // the A region of a stage holds the raw fp32 chunk(s) first: 32 KB, 64 KB with A2; gather: the operand only
  const bool lsu_a = p.g_idx || (ln && p.A2);
  const size_t a_region = p.a16 ? A_PART : (lsu_a ? parts * A_PART : (p.A2 ? 4 * A_PART : 2 * A_PART));
  const size_t stage = a_region + static_cast<size_t>(parts) * NC * KC * 2;
  int stages = static_cast<int>((217 * 1024) / stage);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > p.n_chunks) stages = p.n_chunks;
  BD_REQUIRE(stages >= 1 && (stages >= 2 || p.n_chunks == 1),
             "bd_linear_tc: a pipeline stage needs %zu bytes of shared memory; two must fit 217 KB", stage);
  // Two CTAs per SM when that is possible at all (accumulators <= 256 TMEM columns, kernel variant
  // compiled for two, at least two stages in 109 KB) and the grid is more than one wave: a CTA's
  // timeline is serial (operand latency -> main loop -> write-out), the second CTA fills its gaps.
  const int n_row_tiles = bd::ceil_div(p.M, TC_BM), n_col_groups = bd::ceil_div(p.N, NC);
  if (g_tc_two_per_sm && !ln && !p.A2 && !p.g_idx && tc::tmem_cols_pow2(NC) <= 256 &&
      static_cast<long long>(n_row_tiles) * n_col_groups > bd::sm_count()) {
    const int s2 = static_cast<int>((109 * 1024 - 1024) / stage);
    const size_t tile2 = static_cast<size_t>(TC_BM) * (NC + 4) * 4;
    if (s2 >= 2 && tile2 + 1024 <= 109 * 1024 && s2 < stages) stages = s2;
  }
  p.n_stages = stages;
  const size_t pipe = stage * stages;
  const size_t tile = static_cast<size_t>(TC_BM) * (NC + 4) * 4;
  const size_t smem = (pipe > tile ? pipe : tile) + 1024;  // slack: the dynamic base is 1024-aligned by hand
  BD_REQUIRE(smem <= 218 * 1024, "bd_linear_tc: tiling needs %zu bytes of shared memory (> 218 KB)", smem);
  static bd::PerDeviceOnce configured;  // function attributes are per device
  BD_CUDA(configured.run([&]() {
    cudaError_t e = cudaSuccess;
    const void *kernels[] = {(const void *)linear_tc_kernel<0, 0>, (const void *)linear_tc_kernel<0, 1>,
                             (const void *)linear_tc_kernel<1, 0>, (const void *)linear_tc_kernel<1, 1>,
                             (const void *)linear_tc_kernel<0, 2>, (const void *)linear_tc_kernel<2, 0>,
                             (const void *)linear_tc_kernel<0, 0, true>};
    for (const void *k : kernels)
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024);
    return e;
  }), "bd_linear_tc");
  const int n_groups = bd::ceil_div(p.N, NC);
  BD_REQUIRE(n_groups <= 65535, "bd_linear_tc: N too large");
  dim3 grid(bd::ceil_div(p.M, TC_BM), n_groups);
  BD_REQUIRE(p.relu != 2 || (!p.g_idx && p.pool <= 0 && !ln && !p.A2),
             "bd_linear_tc: the GELU epilogue (relu = 2) exists for the plain A, plain epilogue variant only");
  if (p.g_idx)
    BD_CUDA(bd::launch_pdl(linear_tc_kernel<0, 2>, grid, dim3(TC_THREADS), smem, stream, p), "bd_linear_tc");
  else if (p.pool > 0)
    BD_CUDA(bd::launch_pdl(linear_tc_kernel<2, 0>, grid, dim3(TC_THREADS), smem, stream, p), "bd_linear_tc");
  else if (ln && p.A2)
    BD_CUDA(bd::launch_pdl(linear_tc_kernel<1, 1>, grid, dim3(TC_THREADS), smem, stream, p), "bd_linear_tc");
  else if (ln)
    BD_CUDA(bd::launch_pdl(linear_tc_kernel<1, 0>, grid, dim3(TC_THREADS), smem, stream, p), "bd_linear_tc");
  else if (p.A2)
    BD_CUDA(bd::launch_pdl(linear_tc_kernel<0, 1>, grid, dim3(TC_THREADS), smem, stream, p), "bd_linear_tc");
  else if (p.relu == 2)
    BD_CUDA(bd::launch_pdl(linear_tc_kernel<0, 0, true>, grid, dim3(TC_THREADS), smem, stream, p), "bd_linear_tc");
  else
    BD_CUDA(bd::launch_pdl(linear_tc_kernel<0, 0>, grid, dim3(TC_THREADS), smem, stream, p), "bd_linear_tc");
  return BD_OK;
}

int check_common(const float *A, const void *Wp, float *Y, int lda, int A2_ok, int ldy, int M, int N, int K,
                 int kc, int n_chunks, int BN, int n_sub, int split) {
  BD_REQUIRE(A && Wp && Y, "bd_linear_tc: null pointer");
  BD_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldy >= N && A2_ok, "bd_linear_tc: bad sizes");
  BD_REQUIRE(kc == KC && n_chunks >= 1 && n_chunks * KC >= K, "bd_linear_tc: KC must be 64 and n_chunks * 64 >= K");
  BD_REQUIRE(K % 8 == 0 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0,
             "bd_linear_tc: needs K %% 8 == 0, lda %% 4 == 0 and a 16-byte aligned A (use bd_linear_f32 otherwise)");
  BD_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= 256, "bd_linear_tc: BN must be a multiple of 16 in [16,256]");
  BD_REQUIRE(n_sub >= 1 && n_sub * BN <= 512, "bd_linear_tc: n_sub * BN must fit 512 TMEM columns");
  BD_REQUIRE(split == 1 || split == 3, "bd_linear_tc: split must be 1 (bf16) or 3 (bf16x3)");
  return BD_OK;
}

long long *g_tc_dbg = nullptr;

}  // namespace

// 1 (default): plain linears whose accumulators fit 256 TMEM columns limit their pipeline to 109 KB
// of shared memory so that two CTAs share an SM; 0: always the deepest pipeline, one CTA per SM.
extern "C" int bd_linear_tc_set_occupancy(int two_per_sm) {
  g_tc_two_per_sm = two_per_sm ? 1 : 0;
  return BD_OK;
}

// Tuning aid: device buffer of >= 64 long longs receiving clock64() stamps of CTA (0,0); NULL disables.
extern "C" int bd_linear_tc_set_debug(long long *buf) {
  g_tc_dbg = buf;
  return BD_OK;
}

extern "C" int bd_linear_tc(const float *A, int lda, const float *A2, int lda2, const void *Wp, const float *bias,
                            float *Y, int ldy, int M, int N, int K, int kc, int n_chunks, int BN, int n_sub,
                            int relu, int split, bd_stream_t stream) {
  const int rc = check_common(A, Wp, Y, lda,
                              !A2 || (lda2 >= K && lda2 % 4 == 0 && (reinterpret_cast<uintptr_t>(A2) & 15) == 0), ldy,
                              M, N, K, kc, n_chunks, BN, n_sub, split);
  if (rc != BD_OK) return rc;
  LinearTcParams p = {};
  p.A = A, p.A2 = A2, p.bias = bias, p.Wp = static_cast<const __nv_bfloat16 *>(Wp), p.Y = Y;
  p.lda = lda, p.lda2 = lda2, p.ldy = ldy;
  p.M = M, p.N = N, p.K = K, p.n_chunks = n_chunks, p.BN = BN, p.n_sub = n_sub, p.relu = relu, p.split = split;
  p.dbg = g_tc_dbg;
  const int r2 = launch_linear_tc(p, false, bd::as_stream(stream));
  if (r2 != BD_OK) return r2;
  BD_CHECK_LAUNCH("bd_linear_tc");
  return BD_OK;
}

extern "C" int bd_linear_ln_tc(const float *A, int lda, const float *A2, int lda2, const void *Wp, const float *bias,
                               const float *R, int ldr, const float *gamma, const float *beta, float eps, float *Y,
                               int ldy, int M, int N, int K, int kc, int n_chunks, int BN, int n_sub, int split,
                               bd_stream_t stream) {
  const int rc = check_common(A, Wp, Y, lda,
                              !A2 || (lda2 >= K && lda2 % 4 == 0 && (reinterpret_cast<uintptr_t>(A2) & 15) == 0), ldy,
                              M, N, K, kc, n_chunks, BN, n_sub, split);
  if (rc != BD_OK) return rc;
  BD_REQUIRE(R && gamma && beta && ldr >= N, "bd_linear_ln_tc: null pointer / bad ldr");
  BD_REQUIRE(n_sub * BN >= N && N <= 320, "bd_linear_ln_tc: one CTA must own complete rows (N <= n_sub*BN, N <= 320)");
  BD_REQUIRE(N % 4 == 0 && ldr % 4 == 0 && ldy % 4 == 0 &&
                 ((reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(gamma) |
                   reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
             "bd_linear_ln_tc: N, ldr, ldy must be multiples of 4 and R, Y, gamma, beta 16-byte aligned");
  LinearTcParams p = {};
  p.A = A, p.A2 = A2, p.bias = bias, p.Wp = static_cast<const __nv_bfloat16 *>(Wp), p.Y = Y;
  p.R = R, p.gamma = gamma, p.beta = beta, p.eps = eps, p.ldr = ldr;
  p.lda = lda, p.lda2 = lda2, p.ldy = ldy;
  p.M = M, p.N = N, p.K = K, p.n_chunks = n_chunks, p.BN = BN, p.n_sub = n_sub, p.relu = 0, p.split = split;
  p.dbg = g_tc_dbg;
  const int r2 = launch_linear_tc(p, true, bd::as_stream(stream));
  if (r2 != BD_OK) return r2;
  BD_CHECK_LAUNCH("bd_linear_ln_tc");
  return BD_OK;
}

// 16-bit activation variants (fp16 operand mode only, split = 1, no A2).  a_half: A is an fp16 (M, K) matrix
// (lda in halfs, lda % 8 == 0); y_half: Y is written as fp16 rows (ldy in halfs, ldy % 8 == 0).  Values are the
// ones the fp32 entry points produce followed by the fp16 rounding their consumers apply anyway.
int bd_linear_stream_try(const void *A, int lda, const void *Wp, const float *bias, void *Y, int ldy, int M, int N, int K,
                         int n_chunks, int BN, int n_sub, int relu, cudaStream_t stream);  // gemm_stream.cu

int bd_linear_ln_stream_try(const void *A, int lda, const void *Wp, const float *bias, const float *R, int ldr,
                            const float *gamma, const float *beta, float eps, float *Y, int ldy, void *Y16, int ldy16, int M,
                            int N, int K, int n_chunks, int BN, int n_sub, cudaStream_t stream);  // gemm_stream.cu

extern "C" int bd_linear_tc_h(const void *A, int lda, int a_half, const float *A2, int lda2, const void *Wp,
                              const float *bias, void *Y, int ldy, int y_half, int M, int N, int K, int kc, int n_chunks,
                              int BN, int n_sub, int relu, bd_stream_t stream) {
  BD_REQUIRE(A && Wp && Y, "bd_linear_tc_h: null pointer");
  BD_REQUIRE(!A2 || (!a_half && lda2 >= K && lda2 % 4 == 0 && (reinterpret_cast<uintptr_t>(A2) & 15) == 0),
             "bd_linear_tc_h: A2 (fp32, 16-byte aligned rows) only with an fp32 A");
  BD_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldy >= N, "bd_linear_tc_h: bad sizes");
  BD_REQUIRE(kc == KC && n_chunks >= 1 && n_chunks * KC >= K && K % 8 == 0, "bd_linear_tc_h: KC must be 64, K %% 8 == 0");
  BD_REQUIRE(lda % (a_half ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0, "bd_linear_tc_h: A rows must be 16-byte aligned");
  BD_REQUIRE(!y_half || (ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0), "bd_linear_tc_h: fp16 Y rows must be 16-byte aligned");
  BD_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= 256 && n_sub >= 1 && n_sub * BN <= 512, "bd_linear_tc_h: bad tiling");
  if (a_half && y_half && !A2) {  // 16-bit rows in and out, more than one wave of tiles: the persistent kernel (gemm_stream.cu)
    const int rs = bd_linear_stream_try(A, lda, Wp, bias, Y, ldy, M, N, K, n_chunks, BN, n_sub, relu, bd::as_stream(stream));
    if (rs < 0) return -rs;
    if (rs == 1) {
      BD_CHECK_LAUNCH("bd_linear_tc_h");
      return BD_OK;
    }
  }
  LinearTcParams p = {};
  p.A = static_cast<const float *>(A), p.A2 = A2, p.lda2 = lda2, p.bias = bias, p.Wp = static_cast<const __nv_bfloat16 *>(Wp);
  p.Y = static_cast<float *>(Y), p.Y16 = static_cast<__half *>(Y), p.lda = lda, p.ldy = ldy, p.ldy16 = ldy;
  p.a16 = a_half ? 1 : 0, p.y16 = y_half ? 1 : 0;
  p.M = M, p.N = N, p.K = K, p.n_chunks = n_chunks, p.BN = BN, p.n_sub = n_sub, p.relu = relu, p.split = 1;
  p.dbg = g_tc_dbg;
  const int r2 = launch_linear_tc(p, false, bd::as_stream(stream));
  if (r2 != BD_OK) return r2;
  BD_CHECK_LAUNCH("bd_linear_tc_h");
  return BD_OK;
}

extern "C" int bd_linear_ln_tc_h(const void *A, int lda, int a_half, const void *Wp, const float *bias, const float *R,
                                 int ldr, const float *gamma, const float *beta, float eps, float *Y, int ldy, void *Y16,
                                 int ldy16, int M, int N, int K, int kc, int n_chunks, int BN, int n_sub,
                                 bd_stream_t stream) {
  BD_REQUIRE(A && Wp && Y && R && gamma && beta, "bd_linear_ln_tc_h: null pointer");
  BD_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldy >= N && ldr >= N, "bd_linear_ln_tc_h: bad sizes");
  BD_REQUIRE(kc == KC && n_chunks >= 1 && n_chunks * KC >= K && K % 8 == 0, "bd_linear_ln_tc_h: KC must be 64, K %% 8 == 0");
  BD_REQUIRE(lda % (a_half ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0, "bd_linear_ln_tc_h: A rows must be 16-byte aligned");
  BD_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= 256 && n_sub >= 1 && n_sub * BN <= 512, "bd_linear_ln_tc_h: bad tiling");
  BD_REQUIRE(n_sub * BN >= N && N <= 320, "bd_linear_ln_tc_h: one CTA must own complete rows (N <= n_sub*BN, N <= 320)");
  BD_REQUIRE(N % 4 == 0 && ldr % 4 == 0 && ldy % 4 == 0 &&
                 ((reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(gamma) |
                   reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
             "bd_linear_ln_tc_h: N, ldr, ldy must be multiples of 4 and R, Y, gamma, beta 16-byte aligned");
  BD_REQUIRE(!Y16 || (ldy16 % 4 == 0 && ldy16 >= N && (reinterpret_cast<uintptr_t>(Y16) & 7) == 0), "bd_linear_ln_tc_h: bad Y16");
  if (a_half) {  // fp16 rows in, more row tiles than SMs: the persistent kernel (gemm_stream.cu)
    const int rs = bd_linear_ln_stream_try(A, lda, Wp, bias, R, ldr, gamma, beta, eps, Y, ldy, Y16, ldy16, M, N, K, n_chunks, BN,
                                           n_sub, bd::as_stream(stream));
    if (rs < 0) return -rs;
    if (rs == 1) {
      BD_CHECK_LAUNCH("bd_linear_ln_tc_h");
      return BD_OK;
    }
  }
  LinearTcParams p = {};
  p.A = static_cast<const float *>(A), p.bias = bias, p.Wp = static_cast<const __nv_bfloat16 *>(Wp), p.Y = Y;
  p.R = R, p.gamma = gamma, p.beta = beta, p.eps = eps, p.ldr = ldr;
  p.lda = lda, p.ldy = ldy, p.a16 = a_half ? 1 : 0, p.Y16 = static_cast<__half *>(Y16), p.ldy16 = ldy16;
  p.M = M, p.N = N, p.K = K, p.n_chunks = n_chunks, p.BN = BN, p.n_sub = n_sub, p.relu = 0, p.split = 1;
  p.dbg = g_tc_dbg;
  const int r2 = launch_linear_tc(p, true, bd::as_stream(stream));
  if (r2 != BD_OK) return r2;
  BD_CHECK_LAUNCH("bd_linear_ln_tc_h");
  return BD_OK;
}

// First SharedMLP layer of a set-abstraction level with QueryAndGroup fused into the operand
// staging (pointnet2_utils.py:334-359 + pytorch_utils.py:25-36): no grouped tensor in HBM.
// K = C + 3 rounded up to 8; the weight's K columns must be ordered [features | xyz | 0].
extern "C" int bd_sa_group_linear_tc(const int *idx, const float *feats, int ld_feats, int C, const float *xyz,
                                     int ld_xyz, const float *new_xyz, int B, int n, int m, int ns, float radius,
                                     const void *Wp, const float *bias, float *Y, int ldy, int N, int kc, int n_chunks,
                                     int BN, int n_sub, int split, bd_stream_t stream) {
  BD_REQUIRE(idx && xyz && new_xyz && Wp && Y && (feats || C == 0), "bd_sa_group_linear_tc: null pointer");
  BD_REQUIRE(B > 0 && n > 0 && m > 0 && ns > 0 && C >= 0 && N > 0 && ld_xyz >= 3 && ld_feats >= C && ldy >= N,
             "bd_sa_group_linear_tc: bad sizes");
  const int K = (C + 3 + 7) / 8 * 8;
  BD_REQUIRE(kc == KC && n_chunks * KC >= K, "bd_sa_group_linear_tc: KC must be 64 and cover C + 3");
  BD_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= 256 && n_sub >= 1 && n_sub * BN <= 512 && (split == 1 || split == 3),
             "bd_sa_group_linear_tc: bad tiling");
  BD_REQUIRE(static_cast<long long>(B) * m * ns < (1LL << 31), "bd_sa_group_linear_tc: too many rows");
  LinearTcParams p = {};
  p.A = xyz;  // unused in gather mode (kept non-null)
  p.lda = K, p.bias = bias, p.Wp = static_cast<const __nv_bfloat16 *>(Wp), p.Y = Y, p.ldy = ldy;
  p.M = B * m * ns, p.N = N, p.K = K, p.n_chunks = n_chunks, p.BN = BN, p.n_sub = n_sub, p.relu = 1, p.split = split;
  p.g_idx = idx, p.g_feat = feats ? feats : xyz, p.g_xyz = xyz, p.g_cen = new_xyz;
  p.g_ldf = ld_feats, p.g_ldx = ld_xyz, p.g_C = C, p.g_ns = ns, p.g_n = n, p.g_m = m, p.g_inv_r = 1.0f / radius;
  p.dbg = g_tc_dbg;
  const int r2 = launch_linear_tc(p, false, bd::as_stream(stream));
  if (r2 != BD_OK) return r2;
  BD_CHECK_LAUNCH("bd_sa_group_linear_tc");
  return BD_OK;
}

// Last SharedMLP layer of a set-abstraction level with the max-pool over nsample fused into the
// epilogue (pointnet2_modules.py:251-257): Y (M / pool, N) = max over groups of `pool` rows of
// relu(A Wᵀ + bias).  pool must divide 128.
extern "C" int bd_linear_pool_tc(const float *A, int lda, const void *Wp, const float *bias, float *Y, int ldy, int M,
                                 int N, int K, int kc, int n_chunks, int BN, int n_sub, int pool, int split,
                                 bd_stream_t stream) {
  const int rc = check_common(A, Wp, Y, lda, 1, ldy, M, N, K, kc, n_chunks, BN, n_sub, split);
  if (rc != BD_OK) return rc;
  BD_REQUIRE(pool > 0 && TC_BM % pool == 0 && M % pool == 0, "bd_linear_pool_tc: pool must divide 128 and M");
  LinearTcParams p = {};
  p.A = A, p.bias = bias, p.Wp = static_cast<const __nv_bfloat16 *>(Wp), p.Y = Y, p.lda = lda, p.ldy = ldy;
  p.M = M, p.N = N, p.K = K, p.n_chunks = n_chunks, p.BN = BN, p.n_sub = n_sub, p.relu = 1, p.split = split;
  p.pool = pool;
  p.dbg = g_tc_dbg;
  const int r2 = launch_linear_tc(p, false, bd::as_stream(stream));
  if (r2 != BD_OK) return r2;
  BD_CHECK_LAUNCH("bd_linear_pool_tc");
  return BD_OK;
}
