// Multi-head attention core on tcgen05 — softmax(Q·Kᵀ·scale + mask)·V for nn.MultiheadAttention
// (call sites /root/reference/models/encoder_decoder_layers.py:87,99,111,149,179,365,373,384,394).
//
// Two kernels:
//  1. attention_pack_kernel — converts the fp32 Q / K / V rows of every (scene, head) ONCE into
//     bf16 operand tiles (hi, + lo in the "bf16x3" mode) laid out exactly as the tensor core
//     wants them in shared memory (128-byte swizzle, K-major; Vᵀ via a transposing pass; the
//     softmax scale·log2e folded into Q; head_dim 36 zero-padded to 64).  A K/V tile is reused by
//     every 128-query tile of the head, so the conversion is amortised Lq/128 times.
//  2. attention_tc_kernel — one CTA = 128 queries x 1 head x 1 scene, thread t owns query row t
//     (= TMEM lane t).  Tiles arrive by cp.async.bulk (TMA engine) on mbarriers, K/V
//     double-buffered one tile ahead.  Per tile of 128 keys:
//        S  = Q·Kᵀ        tcgen05.mma  M=128, N=128, K=64
//        softmax           tcgen05.ld of the thread's score row, exact online softmax in fp32,
//                          probabilities written to shared memory as the A operand of MMA 2
//        Ot = P·V          tcgen05.mma  M=128, N=48, K=128 keys into a scratch TMEM accumulator
//        O  = O·corr + Ot  in registers (36 fp32 per thread): no TMEM rescaling pass
//     bf16x3: every product is Ahi*Bhi + Alo*Bhi + Ahi*Blo (fp32-grade) on the bf16 tensor pipe.
#include "tc_common.cuh"

namespace {

constexpr int AT_BM = 128, AT_NV = 48, AT_THREADS = 128, AT_HD = 36;
constexpr uint32_t QK_PART = 128 * 128;  // Q tile: 128 rows x 128 B
constexpr uint32_t V_BLK = AT_NV * 128;  // 48 rows x 128 B (one block of 64 keys)
constexpr uint32_t P_BLK = 128 * 128;    // 128 rows x 64 keys
// key-tile width BK (64 or 128 keys): K tile = BK rows x 128 B, V^T / P tiles = BK / 64 blocks
__host__ __device__ constexpr uint32_t k_part(int BK) { return BK * 128u; }
__host__ __device__ constexpr uint32_t v_part(int BK) { return (BK / 64) * V_BLK; }
__host__ __device__ constexpr uint32_t p_part(int BK) { return (BK / 64) * P_BLK; }

struct AttnParams {
  // DIRECT variant: K and V as 2-D fp16 tensors (inner = the H * 36 columns of the projection output, rows =
  // B * Lk tokens), box = 64 columns x 128 rows, 128-byte swizzle — a K / V tile of one head arrives in the MMA
  // operand layout straight from the rows the projection kernel wrote (no pack kernel, no workspace)
  CUtensorMap tmK, tmV;
  const float *Q, *K, *V;
  const unsigned char *mask;
  float *O;
  unsigned char *Qp, *Kp, *Vp;  // packed operand tiles (workspace)
  int ldq, ldk, ldv, ldo;
  long long sq_b, sk_b, sv_b, so_b;
  int Lq, Lk, H, nq, nk;  // nk = key tiles of BK keys
  float scale_log2;
  int skip_q;  // pack kernel: Q tiles are converted inside the attention kernel, start at the K tiles
  int q16, k16, v16, o16;  // the tensor is fp16 in HBM (leading dimensions / batch strides then count halfs)
  int dbg;  // timing experiments only (bit 0: no softmax arithmetic, bit 1: no P.V MMAs, bit 2: no Q.K^T MMAs)
};

// 8 consecutive head dims of an fp32 or fp16 row (element offset `off`; `second`: dims 4..7 are inside the head)
__device__ __forceinline__ void load_chunk8(const float *base, long long off, bool half, bool second, float (&v)[8], bool first = true) {
  if (half) {
    const __half *s = reinterpret_cast<const __half *>(base) + off;  // 8-byte aligned (ld % 4 == 0, head offset 72 h bytes)
    const uint2 a = first ? __ldg(reinterpret_cast<const uint2 *>(s)) : make_uint2(0u, 0u);
    const uint2 b = second ? __ldg(reinterpret_cast<const uint2 *>(s + 4)) : make_uint2(0u, 0u);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&a.x)), f1 = __half22float2(*reinterpret_cast<const __half2 *>(&a.y));
    const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&b.x)), f3 = __half22float2(*reinterpret_cast<const __half2 *>(&b.y));
    v[0] = f0.x, v[1] = f0.y, v[2] = f1.x, v[3] = f1.y, v[4] = f2.x, v[5] = f2.y, v[6] = f3.x, v[7] = f3.y;
  } else {
    const float4 *s = reinterpret_cast<const float4 *>(base + off);
    const float4 a = first ? __ldg(s) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b = second ? __ldg(s + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  }
}

// ------------------------------------------------------------------------------------------ pack
// grid.x = nq + 2 * nk : [0,nq) Q tiles, [nq,nq+nk) K tiles, rest V tiles ; grid.y = H ; grid.z = B
template <int PARTS, int BK>
__global__ void __launch_bounds__(256) attention_pack_kernel(const AttnParams p) {
  constexpr uint32_t K_PART = k_part(BK), V_PART = v_part(BK);
  const int tid = threadIdx.x;
  const int b = blockIdx.z, h = blockIdx.y;
  bd::pdl_launch_dependents();
  bd::pdl_wait();  // Q / K / V come from the preceding projection kernels
  int t = blockIdx.x + (p.skip_q ? p.nq : 0);
  if (t < p.nq + p.nk) {
    // ---- row tiles (Q or K): item = (row, 16-byte chunk of 8 head dims); chunks 5..7 are padding
    const bool is_q = t < p.nq;
    if (!is_q) t -= p.nq;
    const float *src = is_q ? p.Q : p.K;
    const bool src16 = is_q ? p.q16 : p.k16;
    const long long src_off = (is_q ? b * p.sq_b : b * p.sk_b) + h * AT_HD;
    const int tile_rows = is_q ? 128 : BK;
    const uint32_t part = is_q ? QK_PART : K_PART;
    const int ld = is_q ? p.ldq : p.ldk, n_rows = is_q ? p.Lq : p.Lk, row0 = t * tile_rows;
    const float mul = is_q ? p.scale_log2 : 1.0f;
    unsigned char *dst = (is_q ? p.Qp + (static_cast<size_t>(b) * p.H + h) * p.nq * (PARTS * QK_PART)
                               : p.Kp + (static_cast<size_t>(b) * p.H + h) * p.nk * (PARTS * K_PART)) +
                         static_cast<size_t>(t) * (PARTS * part);
    // 128 rows x 6 chunks (the three K = 16 steps read chunks 0..5; 36 dims -> chunks 0..4, chunk 5
    // is zero) = 768 items, 3 per thread, chunk-fastest: the lanes of a warp read ~5 whole rows
    // (144 contiguous bytes each) instead of 32 different ones; loads batched
    float a[3][8];
    int rr[3], cc[3];
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      const int e = tid + it * 256;
      const int r = e / 6, ch = e - r * 6;
      rr[it] = r < tile_rows ? r : -1, cc[it] = ch;
#pragma unroll
      for (int i = 0; i < 8; ++i) a[it][i] = 0.f;
      if (r < tile_rows && row0 + r < n_rows && ch * 8 < AT_HD)
        load_chunk8(src, src_off + static_cast<long long>(row0 + r) * ld + ch * 8, src16, ch * 8 + 4 < AT_HD, a[it]);
    }
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      const float v[8] = {a[it][0] * mul, a[it][1] * mul, a[it][2] * mul, a[it][3] * mul,
                          a[it][4] * mul, a[it][5] * mul, a[it][6] * mul, a[it][7] * mul};
      uint4 hi, lo;
      tc::cvt8(PARTS, v, hi, lo);
      if (rr[it] < 0) continue;
      const uint32_t off = tc::sw128_off(rr[it], cc[it]);
      *reinterpret_cast<uint4 *>(dst + off) = hi;
      if (PARTS == 2) *reinterpret_cast<uint4 *>(dst + part + off) = lo;
    }
  } else {
    // ---- V^T tile: rows = head dims (48, 36 valid), K = 128 keys in two blocks of 64
    t -= p.nq + p.nk;
    const float *src = p.V + (p.v16 ? 0 : b * p.sv_b + h * AT_HD);
    const __half *src_h = reinterpret_cast<const __half *>(p.V) + b * p.sv_b + h * AT_HD;
    const int k0 = t * BK;
    unsigned char *dst = p.Vp + ((static_cast<size_t>(b) * p.H + h) * p.nk + t) * (PARTS * V_PART);
    // 48 rows x BK/8 key-chunks items (768 for BK = 128), 3 per thread; item = 8 keys of one head dim
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      const int e = tid + it * 256;
      if (e >= AT_NV * (BK / 8)) break;
      const int d = e % AT_NV, kc = e / AT_NV;  // consecutive threads -> consecutive head dims (coalesced)
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int key = k0 + kc * 8 + i;
        // row AT_HD of V^T is all ones: column AT_HD of P.V is then the row sum of P (the softmax
        // denominator comes out of the tensor core with the numerator)
        v[i] = (d < AT_HD && key < p.Lk)
                   ? (p.v16 ? __half2float(__ldg(src_h + static_cast<long long>(key) * p.ldv + d)) : __ldg(src + static_cast<long long>(key) * p.ldv + d))
                   : (d == AT_HD ? 1.f : 0.f);
      }
      uint4 hi, lo;
      tc::cvt8(PARTS, v, hi, lo);
      const uint32_t off = (kc >> 3) * V_BLK + tc::sw128_off(d, kc & 7);
      *reinterpret_cast<uint4 *>(dst + off) = hi;
      if (PARTS == 2) *reinterpret_cast<uint4 *>(dst + V_PART + off) = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------ main
template <int PARTS, int BK, int STAGES>
__global__ void __launch_bounds__(AT_THREADS, STAGES == 1 ? 2 : 1) attention_tc_kernel(const AttnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr uint32_t K_PART = k_part(BK), V_PART = v_part(BK), P_PART = p_part(BK);
  constexpr uint32_t KV_STAGE = PARTS * (K_PART + V_PART);
  constexpr int AT_BK = BK;
  unsigned char *sQ = smem;                       // PARTS * 16 KB
  unsigned char *sKV = sQ + PARTS * QK_PART;      // STAGES x (K tile | V^T tile)
  unsigned char *sP = sKV + STAGES * KV_STAGE;
  __shared__ __align__(8) unsigned long long bar_mma, bar_q, bar_kv[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ unsigned char kvalid[AT_BK];

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x, q0 = qt * AT_BM;
  const unsigned char *mask = p.mask ? p.mask + static_cast<long long>(b) * p.Lk : nullptr;
  const size_t bh = static_cast<size_t>(b) * p.H + h;
  const unsigned char *Qp = p.Qp + (bh * p.nq + qt) * (PARTS * QK_PART);
  const unsigned char *Kp = p.Kp + bh * p.nk * (PARTS * K_PART);
  const unsigned char *Vp = p.Vp + bh * p.nk * (PARTS * V_PART);

  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), AT_BK == 64 ? 128 : 256);
  if (tid == 32) {
    tc::mbar_init(tc::smem_u32(&bar_mma), 1);
    tc::mbar_init(tc::smem_u32(&bar_q), 1);
    tc::mbar_init(tc::smem_u32(&bar_kv[0]), 1);
    tc::mbar_init(tc::smem_u32(&bar_kv[1]), 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, tmem_base_s, 0);
  const uint32_t tmem_s = tmem, tmem_o = tmem + AT_BK;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;

  auto issue_kv = [&](int j) {  // thread 0: K and V^T tiles of key tile j -> stage j % STAGES
    const uint32_t st = j % STAGES;
    const uint32_t bar = tc::smem_u32(&bar_kv[st]);
    tc::mbar_arrive_expect_tx(bar, KV_STAGE);
    tc::bulk_g2s(tc::smem_u32(sKV + st * KV_STAGE), Kp + static_cast<size_t>(j) * (PARTS * K_PART), PARTS * K_PART, bar);
    tc::bulk_g2s(tc::smem_u32(sKV + st * KV_STAGE + PARTS * K_PART), Vp + static_cast<size_t>(j) * (PARTS * V_PART),
                 PARTS * V_PART, bar);
  };
  if (tid == 0) {
    tc::mbar_arrive_expect_tx(tc::smem_u32(&bar_q), PARTS * QK_PART);
    tc::bulk_g2s(tc::smem_u32(sQ), Qp, PARTS * QK_PART, tc::smem_u32(&bar_q));
    issue_kv(0);
  }

  float m_run = -INFINITY, l_run = 0.f;
  float o_acc[AT_HD];
#pragma unroll
  for (int i = 0; i < AT_HD; ++i) o_acc[i] = 0.f;
  uint32_t phase = 0;
  const uint32_t idesc_s = tc::idesc_ab(PARTS, AT_BM, AT_BK), idesc_o = tc::idesc_ab(PARTS, AT_BM, AT_NV);
  const int n_tiles = p.nk;

  for (int j = 0; j < n_tiles; ++j) {
    const uint32_t st = j % STAGES;
    const int k0 = j * AT_BK;
    // prefetch the next K/V tile into the other stage (its previous MMAs were waited for by every
    // thread at the end of the last iteration); single-stage: the tile itself is requested here and
    // a co-resident CTA covers the latency
    if (tid == 0) {
      if (STAGES == 2 && j + 1 < n_tiles) issue_kv(j + 1);
      if (STAGES == 1 && j > 0) issue_kv(j);
    }
    const bool my_valid = tid >= AT_BK || ((k0 + tid < p.Lk) && !(mask && mask[k0 + tid]));
    if (tid < AT_BK) kvalid[tid] = my_valid;
    const bool all_valid = __syncthreads_and(my_valid);  // common case: full, unmasked tile

    // ---- S = Q K^T
    if (warp == 0) {
      if (j == 0) tc::mbar_wait(tc::smem_u32(&bar_q), 0);
      tc::mbar_wait(tc::smem_u32(&bar_kv[st]), (j / STAGES) & 1);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t q = tc::smem_u32(sQ), k = tc::smem_u32(sKV + st * KV_STAGE);
#pragma unroll
        for (int s = 0; s < 3; ++s) {  // head_dim 36 -> 48: chunks 6, 7 of the tiles are not written
          const uint64_t dq = tc::smem_desc_sw128(q + s * 32), dk = tc::smem_desc_sw128(k + s * 32);
          tc::mma_bf16(tmem_s, dq, dk, idesc_s, s > 0 ? 1u : 0u);
          if (PARTS == 2) {
            tc::mma_bf16(tmem_s, tc::smem_desc_sw128(q + QK_PART + s * 32), dk, idesc_s, 1u);
            tc::mma_bf16(tmem_s, dq, tc::smem_desc_sw128(k + K_PART + s * 32), idesc_s, 1u);
          }
        }
        tc::mma_commit(tc::smem_u32(&bar_mma));
      }
      __syncwarp();
    }
    tc::mbar_wait(tc::smem_u32(&bar_mma), phase);
    phase ^= 1u;
    tc::fence_after_sync();

    // ---- online softmax on the thread's row (two passes over TMEM: max, then exp / sum / pack)
    float mx = -INFINITY;
#pragma unroll
    for (int g = 0; g < AT_BK / 16; g += 2) {
      uint32_t acc0[16], acc1[16];
      tc::tmem_ld16(tmem_s + lane_base + g * 16, acc0);
      tc::tmem_ld16(tmem_s + lane_base + (g + 1) * 16, acc1);
      tc::tmem_ld_wait();
      if (all_valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(acc0[i]), __uint_as_float(acc1[i])));
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (kvalid[g * 16 + i]) mx = fmaxf(mx, __uint_as_float(acc0[i]));
          if (kvalid[(g + 1) * 16 + i]) mx = fmaxf(mx, __uint_as_float(acc1[i]));
        }
      }
    }
    const float m_new = fmaxf(m_run, mx);
    const float corr = (m_new == -INFINITY) ? 1.f : exp2f(m_run - m_new);
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < AT_BK / 16; ++g) {
      uint32_t acc[16];
      tc::tmem_ld16(tmem_s + lane_base + g * 16, acc);
      tc::tmem_ld_wait();
      float pv[16];
      if (all_valid) {  // m_new is finite here
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pv[i] = exp2f(__uint_as_float(acc[i]) - m_new);
          sum += pv[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pv[i] = (kvalid[g * 16 + i] && m_new != -INFINITY) ? exp2f(__uint_as_float(acc[i]) - m_new) : 0.f;
          sum += pv[i];
        }
      }
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const float v8[8] = {pv[c2 * 8 + 0], pv[c2 * 8 + 1], pv[c2 * 8 + 2], pv[c2 * 8 + 3],
                             pv[c2 * 8 + 4], pv[c2 * 8 + 5], pv[c2 * 8 + 6], pv[c2 * 8 + 7]};
        uint4 hi, lo;
        tc::cvt8(PARTS, v8, hi, lo);
        const int kc = g * 2 + c2;  // 16-byte chunk of keys
        const uint32_t off = (kc >> 3) * P_BLK + tc::sw128_off(tid, kc & 7);
        *reinterpret_cast<uint4 *>(sP + off) = hi;
        if (PARTS == 2) *reinterpret_cast<uint4 *>(sP + P_PART + off) = lo;
      }
    }
    l_run = l_run * corr + sum;
    m_run = m_new;
    tc::fence_before_sync();  // TMEM reads of S done before the next tile's MMA overwrites it
    tc::fence_proxy_async_smem();
    __syncthreads();

    // ---- Ot = P V
    if (warp == 0) {
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t pa = tc::smem_u32(sP), va = tc::smem_u32(sKV + st * KV_STAGE + PARTS * K_PART);
#pragma unroll
        for (int blk = 0; blk < AT_BK / 64; ++blk) {
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const uint64_t dp = tc::smem_desc_sw128(pa + blk * P_BLK + s * 32);
            const uint64_t dv = tc::smem_desc_sw128(va + blk * V_BLK + s * 32);
            tc::mma_bf16(tmem_o, dp, dv, idesc_o, (blk | s) ? 1u : 0u);
            if (PARTS == 2) {
              tc::mma_bf16(tmem_o, tc::smem_desc_sw128(pa + P_PART + blk * P_BLK + s * 32), dv, idesc_o, 1u);
              tc::mma_bf16(tmem_o, dp, tc::smem_desc_sw128(va + V_PART + blk * V_BLK + s * 32), idesc_o, 1u);
            }
          }
        }
        tc::mma_commit(tc::smem_u32(&bar_mma));
      }
      __syncwarp();
    }
    tc::mbar_wait(tc::smem_u32(&bar_mma), phase);
    phase ^= 1u;
    tc::fence_after_sync();
    {
      uint32_t a0[16], a1[16], a2[16];
      tc::tmem_ld16(tmem_o + lane_base, a0);
      tc::tmem_ld16(tmem_o + lane_base + 16, a1);
      tc::tmem_ld16(tmem_o + lane_base + 32, a2);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) o_acc[i] = fmaf(o_acc[i], corr, __uint_as_float(a0[i]));
#pragma unroll
      for (int i = 0; i < 16; ++i) o_acc[16 + i] = fmaf(o_acc[16 + i], corr, __uint_as_float(a1[i]));
#pragma unroll
      for (int i = 0; i < 4; ++i) o_acc[32 + i] = fmaf(o_acc[32 + i], corr, __uint_as_float(a2[i]));
    }
    tc::fence_before_sync();
    __syncthreads();  // stage st, the P tile, kvalid and both accumulators are free again
  }

  if (warp == 0) tc::tmem_dealloc(tmem, AT_BK == 64 ? 128 : 256);
  if (q0 + tid < p.Lq) {
    float *dst = p.O + b * p.so_b + static_cast<long long>(q0 + tid) * p.ldo + h * AT_HD;
    const float inv = 1.0f / l_run;  // l == 0 (every key masked) -> NaN like the reference softmax
#pragma unroll
    for (int d = 0; d < AT_HD; d += 4) {
      float4 o4 = make_float4(o_acc[d] * inv, o_acc[d + 1] * inv, o_acc[d + 2] * inv, o_acc[d + 3] * inv);
      if (l_run == 0.f) o4 = make_float4(NAN, NAN, NAN, NAN);
      dst[d] = o4.x, dst[d + 1] = o4.y, dst[d + 2] = o4.z, dst[d + 3] = o4.w;
    }
  }
}


// ------------------------------------------------------------------ warp-specialised main kernel
// One CTA = 256 queries (two 128-row tiles) x 1 head x 1 scene, 12 warps:
//   warp 0      loader: cp.async.bulk of the Q tiles, then the K and V^T tiles of every 128-key
//               tile through two 2-stage rings (full / empty mbarriers)
//   warp 1      MMA issuer (one elected lane): S_t = Q_t.K^T into TMEM, Ot_t = P_t.V with the A
//               operand P_t read FROM TMEM (tcgen05.mma .ts form) — the probabilities never
//               touch shared memory
//   warps 4-7   softmax of query tile 0, warps 8-11 of query tile 1 (thread = query row = TMEM
//               lane).  S_t is read from TMEM, exponentiated, split into bf16 hi / lo and stored
//               back IN PLACE over the scores it came from (32 fp32 columns -> 16 hi + 16 lo
//               columns); the tile's P.V result is folded into the register accumulator one
//               iteration later, when the tensor pipe has long finished it.
// While one tile's warps do their softmax the tensor pipe runs the other tile's two MMAs
// (ping-pong), so MUFU / FMA work and tensor work overlap inside one CTA.
// TMEM columns: [0,128) S_0 / P_0, [128,256) S_1 / P_1, [256,304) Ot_0, [320,368) Ot_1.
// TILES = 1 (short key sequences): one query tile per CTA, 6 warps (loader, MMA issuer, four softmax
// warps), 256 TMEM columns ([0,128) S / P, [128,176) Ot) and ~61 KB of shared memory, so TWO CTAs share an
// SM: with one or two key tiles a CTA is a serial chain (load -> S -> softmax -> P.V -> write-out) and
// the second CTA is what fills its gaps.
constexpr int WS_THREADS = 384, WS_THREADS1 = 192, WS_BK = 128;

// second softmax pass over one 128-key score row held in TMEM (see attention_ws_kernel)
template <int PARTS, bool MASKED>
__device__ __forceinline__ void softmax_pass2(uint32_t tS, unsigned long long neg_m2, const uint32_t (&vw)[4], bool dead) {
  // the load of chunk c + 1 is in flight while chunk c is exponentiated (distinct columns: the
  // in-place store of chunk c cannot touch them)
  uint32_t ab[2][32];
  tc::tmem_ld32(tS, ab[0]);
  tc::tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t o[32];
    uint32_t(&a)[32] = ab[c & 1];
    if (c + 1 < 4) tc::tmem_ld32(tS + (c + 1) * 32, ab[(c + 1) & 1]);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float x0 = __uint_as_float(a[i]), x1 = __uint_as_float(a[i + 1]);
      tc::add_f32x2(x0, x1, neg_m2);
      float p0 = tc::ex2_approx(x0), p1 = tc::ex2_approx(x1);
      if (MASKED) {
        if (dead || !((vw[c] >> i) & 1u)) p0 = 0.f;
        if (dead || !((vw[c] >> (i + 1)) & 1u)) p1 = 0.f;
      }
      if (PARTS == 2) {
        tc::split_bf16x2(p0, p1, o[i >> 1], o[16 + (i >> 1)]);
      } else {
        o[i >> 1] = tc::pack_f16x2(p0, p1);
      }
    }
    if (PARTS == 2) {
      tc::tmem_st32(tS + c * 32, o);
    } else {
      uint32_t o16[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o16[i] = o[i];
      tc::tmem_st16(tS + c * 32, o16);
    }
    if (c + 1 < 4) tc::tmem_ld_wait();
  }
}

// A short last key tile (fewer than 97 keys): both softmax passes over the `nch` 32-key chunks that hold keys only
// (plain loops, one chunk in flight: these tiles are a small share of the work; the rest of the S / P columns
// stays stale and the P.V MMAs stop at the same chunk).  Returns the row maximum of the tile through `mx`.
template <int PARTS>
__device__ __forceinline__ float softmax_short_max(uint32_t tS, const uint32_t (&vw)[4], int nch) {
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (c < nch) {
      uint32_t a[32];
      tc::tmem_ld32(tS + c * 32, a);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if ((vw[c] >> i) & 1u) mx = fmaxf(mx, __uint_as_float(a[i]));
    }
  }
  return mx;
}
template <int PARTS>
__device__ __forceinline__ void softmax_short_exp(uint32_t tS, float neg_m, const uint32_t (&vw)[4], bool dead, int nch) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (c < nch) {
      uint32_t a[32], o[32];
      tc::tmem_ld32(tS + c * 32, a);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0 = tc::ex2_approx(__uint_as_float(a[i]) + neg_m), p1 = tc::ex2_approx(__uint_as_float(a[i + 1]) + neg_m);
        if (dead || !((vw[c] >> i) & 1u)) p0 = 0.f;
        if (dead || !((vw[c] >> (i + 1)) & 1u)) p1 = 0.f;
        if (PARTS == 2) tc::split_bf16x2(p0, p1, o[i >> 1], o[16 + (i >> 1)]);
        else o[i >> 1] = tc::pack_f16x2(p0, p1);
      }
      if (PARTS == 2) {
        tc::tmem_st32(tS + c * 32, o);
      } else {
        uint32_t o16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o16[i] = o[i];
        tc::tmem_st16(tS + c * 32, o16);
      }
    }
  }
}

// DIRECT (fp16 K / V rows in HBM, PARTS == 1): the loader fetches each head's K tile [128 keys x 64 columns]
// and V tile [128 keys x 64 columns] with ONE tensor copy each from the projection output.  A tensor copy must
// start on a 16-byte boundary and a head starts at byte 72 h, so odd heads start 4 columns early (SHIFT = 4):
// the tile holds columns 36 h - SHIFT .. + 63 — the head's 36 dims at SHIFT .. SHIFT + 35, around them the
// neighbouring heads' dims / zero fill.  Q's tile is built with zeros everywhere but at its 36 dims (same
// SHIFT), so the foreign columns of K contribute exactly nothing, and the foreign columns of P.V are not read.  V stays in its natural [key][dim] layout: the B operand of P.V is MN-major.
// The softmax denominator comes from a second, N = 16 MMA of P against a constant tile whose first row is ones.
constexpr uint32_t ONES_BLK = 16 * 128;  // 16 rows x 64 keys
// HD = 64 (DIRECT only; RoBERTa-base heads of the text encoder): four K = 16 steps, P.V with N = 64, no shift.
template <int PARTS, int TILES, bool DIRECT = false, int HD = AT_HD>
__global__ void __launch_bounds__(TILES == 2 ? WS_THREADS : WS_THREADS1, (TILES == 2 || HD != AT_HD) ? 1 : 2) attention_ws_kernel(const __grid_constant__ AttnParams p) {
  static_assert(!DIRECT || (PARTS == 1 && TILES == 1), "DIRECT: fp16 operands, one query tile per CTA");
  static_assert(HD == AT_HD || (HD == 64 && DIRECT), "head_dim 64 only with fp16 K / V rows (DIRECT)");
  constexpr int NV = HD == AT_HD ? AT_NV : 64;  // accumulator columns of P.V (head dims padded to a multiple of 16)
  constexpr int NO = HD == AT_HD ? 40 : 64;     // accumulator columns kept in registers
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr uint32_t K_PART = k_part(WS_BK), V_PART = DIRECT ? WS_BK * 128u : v_part(WS_BK);
  constexpr uint32_t Q_TILE = PARTS * QK_PART, K_TILE = PARTS * K_PART, V_TILE = PARTS * V_PART;
  unsigned char *sQ = smem;                   // TILES query tiles
  unsigned char *sK = sQ + TILES * Q_TILE;    // 2 stages
  constexpr int SM0 = TILES == 2 ? 4 : 2;     // first softmax warp
  constexpr uint32_t TM_COLS = TILES == 2 ? 512 : 256, TM_O = TILES == 2 ? 256 : 128;
  unsigned char *sV = sK + 2 * K_TILE;    // 2 stages
  unsigned char *sOnes = sV + 2 * V_TILE;  // DIRECT: [2 blocks of 64 keys][16 rows][128 B], row 0 = ones
  __shared__ __align__(8) unsigned long long bar_q, bar_kf[2], bar_ke[2], bar_vf[2], bar_ve[2], bar_s[2], bar_p[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const int lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, qt0 = blockIdx.x * TILES;
  const int nt = min(TILES, p.nq - qt0);  // query tiles of this CTA
  const int nk = p.nk;
  const size_t bh = static_cast<size_t>(b) * p.H + h;

  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), TM_COLS);
  if (tid == 32) {
    tc::mbar_init(tc::smem_u32(&bar_q), 4 * TILES);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(tc::smem_u32(&bar_kf[i]), 1);
      tc::mbar_init(tc::smem_u32(&bar_ke[i]), 1);
      tc::mbar_init(tc::smem_u32(&bar_vf[i]), 1);
      tc::mbar_init(tc::smem_u32(&bar_ve[i]), 1);
      tc::mbar_init(tc::smem_u32(&bar_s[i]), 1);
      tc::mbar_init(tc::smem_u32(&bar_p[i]), 128);
    }
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, tmem_base_s, 0);
  bd::pdl_launch_dependents();
  bd::pdl_wait();  // packed K / V tiles (pack kernel) and Q rows (projection kernel) are complete

  if (warp == 0) {
    // ------------------------------------------------------------------------------ loader
    if (tc::elect_one()) {
      const unsigned char *Kp = p.Kp + bh * nk * K_TILE;
      const unsigned char *Vp = p.Vp + bh * nk * V_TILE;
      for (int j = 0; j < nk; ++j) {
        const int st = j & 1;
        const uint32_t par = ((j >> 1) - 1) & 1;
        if (j >= 2) tc::mbar_wait(tc::smem_u32(&bar_ke[st]), par);
        tc::mbar_arrive_expect_tx(tc::smem_u32(&bar_kf[st]), K_TILE);
        if (DIRECT)  // rows past this scene's keys (next scene / zero fill) are masked keys; the fill counts as bytes
          tc::tma_load_2d(tc::smem_u32(sK + st * K_TILE), &p.tmK, HD == AT_HD ? h * AT_HD - (h & 1) * 4 : h * HD, b * p.Lk + j * WS_BK, tc::smem_u32(&bar_kf[st]));
        else
          tc::bulk_g2s(tc::smem_u32(sK + st * K_TILE), Kp + static_cast<size_t>(j) * K_TILE, K_TILE, tc::smem_u32(&bar_kf[st]));
        if (j >= 2) tc::mbar_wait(tc::smem_u32(&bar_ve[st]), par);
        tc::mbar_arrive_expect_tx(tc::smem_u32(&bar_vf[st]), V_TILE);
        if (DIRECT)
          tc::tma_load_2d(tc::smem_u32(sV + st * V_TILE), &p.tmV, HD == AT_HD ? h * AT_HD - (h & 1) * 4 : h * HD, b * p.Lk + j * WS_BK, tc::smem_u32(&bar_vf[st]));
        else
          tc::bulk_g2s(tc::smem_u32(sV + st * V_TILE), Vp + static_cast<size_t>(j) * V_TILE, V_TILE, tc::smem_u32(&bar_vf[st]));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // -------------------------------------------------------------------------- MMA issuer
    const uint32_t idesc_o = tc::idesc_ab(PARTS, AT_BM, NV) | (DIRECT ? tc::idesc_b_mn : 0u),
                   idesc_1 = tc::idesc_ab(PARTS, AT_BM, 16);
    tc::mbar_wait(tc::smem_u32(&bar_q), 0);
    for (int j = 0; j <= nk; ++j) {  // iteration j: P.V of key tile j-1, then Q.K^T of key tile j
      for (int t = 0; t < nt; ++t) {
        const uint32_t tS = tmem + t * 128, tO = tmem + TM_O + t * 64;
        if (j > 0) {
          const int jj = j - 1, st = jj & 1;
          tc::mbar_wait(tc::smem_u32(&bar_p[t]), jj & 1);
          if (t == 0) tc::mbar_wait(tc::smem_u32(&bar_vf[st]), (jj >> 1) & 1);
          tc::fence_after_sync();
          if (tc::elect_one()) {
            const uint32_t va = tc::smem_u32(sV + st * V_TILE);
            // a short last key tile: only the 32-key chunks that hold keys were exponentiated (the rest of the
            // S / P columns are stale), so P.V stops there
            const int ksteps = 2 * ((min(WS_BK, p.Lk - jj * WS_BK) + 31) >> 5);
            if (!(p.dbg & 2))
#pragma unroll
            for (int s = 0; s < WS_BK / 16; ++s) {  // 16 keys per step: 8 TMEM columns of packed bf16 pairs
              if (s >= ksteps) break;
              const uint32_t a_hi = tS + (s >> 1) * 32 + (s & 1) * 8;
              const uint64_t dv = DIRECT ? tc::smem_desc_sw128_mn(va + s * 2048)  // 16 keys = two 8-row swizzle atoms
                                         : tc::smem_desc_sw128(va + (s >> 2) * V_BLK + (s & 3) * 32);
              tc::mma_bf16_ts(tO, a_hi, dv, idesc_o, s > 0 ? 1u : 0u);
              if (DIRECT)  // row sums of P: column 48 of the accumulator
                tc::mma_bf16_ts(tO + NV, a_hi, tc::smem_desc_sw128(tc::smem_u32(sOnes) + (s >> 2) * ONES_BLK + (s & 3) * 32),
                                idesc_1, s > 0 ? 1u : 0u);
              if (PARTS == 2) {
                tc::mma_bf16_ts(tO, a_hi + 16, dv, idesc_o, 1u);
                tc::mma_bf16_ts(tO, a_hi, tc::smem_desc_sw128(va + V_PART + (s >> 2) * V_BLK + (s & 3) * 32), idesc_o, 1u);
              }
            }
            if (t == nt - 1) tc::mma_commit(tc::smem_u32(&bar_ve[st]));
            if (j == nk) tc::mma_commit(tc::smem_u32(&bar_s[t]));
          }
          __syncwarp();
        }
        if (j < nk) {
          const int st = j & 1;
          if (t == 0) tc::mbar_wait(tc::smem_u32(&bar_kf[st]), (j >> 1) & 1);
          tc::fence_after_sync();
          if (tc::elect_one()) {
            const uint32_t q = tc::smem_u32(sQ + t * Q_TILE), k = tc::smem_u32(sK + st * K_TILE);
            // scores only for the 32-key chunks of this tile that hold keys (N = 32, 64, 96 or 128)
            const uint32_t idesc_s = tc::idesc_ab(PARTS, AT_BM, 32 * ((min(WS_BK, p.Lk - j * WS_BK) + 31) >> 5));
            if (!(p.dbg & 4))
#pragma unroll
            for (int s = 0; s < NV / 16; ++s) {  // head_dim 36 -> 48: three K=16 steps (the tile is padded to 64)
              const uint64_t dq = tc::smem_desc_sw128(q + s * 32), dk = tc::smem_desc_sw128(k + s * 32);
              tc::mma_bf16(tS, dq, dk, idesc_s, s > 0 ? 1u : 0u);
              if (PARTS == 2) {
                tc::mma_bf16(tS, tc::smem_desc_sw128(q + QK_PART + s * 32), dk, idesc_s, 1u);
                tc::mma_bf16(tS, dq, tc::smem_desc_sw128(k + K_PART + s * 32), idesc_s, 1u);
              }
            }
            tc::mma_commit(tc::smem_u32(&bar_s[t]));
            if (t == nt - 1) tc::mma_commit(tc::smem_u32(&bar_ke[st]));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= SM0) {
    // ----------------------------------------------------------------------- softmax warps
    // First, while the K / V tiles are in flight: the CTA's own Q rows, fp32 -> bf16 hi / lo
    // (softmax scale * log2 e folded in) straight into the operand tiles — a Q tile has exactly
    // one reader, so it never goes through the pack kernel and HBM.  Item = (row, 16-byte chunk
    // of 8 head dims); chunks 0..5 cover the three K = 16 steps (36 dims, rest zero).
    {
      const long long q_off = b * p.sq_b + h * HD;
      constexpr int NCH = NV / 8;  // 16-byte chunks the MMAs read per row
      for (int e = tid - SM0 * 32; e < nt * AT_BM * NCH; e += TILES * 128) {
        const int rr = e / NCH, ch = e - rr * NCH;  // rr = row within the CTA's nt * 128 queries
        const int q = qt0 * AT_BM + rr;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int d0 = ch * 8 - ((DIRECT && HD == AT_HD) ? (h & 1) * 4 : 0);  // first head dim of the chunk (DIRECT, odd heads: shifted by 4)
        const bool first = d0 >= 0 && d0 < HD, second = d0 + 4 < HD;
        if (q < p.Lq && (first || second)) {
          load_chunk8(p.Q, q_off + static_cast<long long>(q) * p.ldq + d0, p.q16, second, v, first);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] *= p.scale_log2;
        }
        uint4 hi, lo;
        tc::cvt8(PARTS, v, hi, lo);
        unsigned char *dst = sQ + (rr >> 7) * Q_TILE + tc::sw128_off(rr & 127, ch);
        *reinterpret_cast<uint4 *>(dst) = hi;
        if (PARTS == 2) *reinterpret_cast<uint4 *>(dst + QK_PART) = lo;
      }
      if (DIRECT) {  // constant tile of the row-sum MMA: row 0 (the first 128 bytes of each block) = 1.0 (fp16), rest 0
        for (int e = tid - SM0 * 32; e < 2 * static_cast<int>(ONES_BLK) / 16; e += TILES * 128) {
          const uint32_t one2 = (e & 127) < 8 ? 0x3C003C00u : 0u;
          *reinterpret_cast<uint4 *>(sOnes + e * 16) = make_uint4(one2, one2, one2, one2);
        }
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bar_q));
    }
    if (((warp - SM0) >> 2) < nt) {
    const int t = (warp - SM0) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + t * 128, tO = tmem + lane_base + TM_O + t * 64;
    const unsigned char *mask = p.mask ? p.mask + static_cast<long long>(b) * p.Lk : nullptr;
    float m_run = -INFINITY, corr_prev = 1.f;
    float o_acc[NO];  // [0,36) output dims, [36] running softmax denominator, rest padding (DIRECT: accumulator
    // columns 0..39 — the head's dims at SHIFT .. SHIFT + 35 — and the denominator in l_dir)
    float l_dir = 0.f;
#pragma unroll
    for (int i = 0; i < NO; ++i) o_acc[i] = 0.f;

    auto fold_o = [&]() {  // o_acc = o_acc * corr + Ot (result of the previous key tile)
      uint32_t a[32], c[NO - 32];
      tc::tmem_ld32(tO, a);
      if constexpr (NO == 40) tc::tmem_ld8(tO + 32, c);
      else tc::tmem_ld32(tO + 32, c);
      if (DIRECT) {  // the row sum of P is the first column after the head dims
        uint32_t l;
        tc::tmem_ld1(tO + NV, l);
        tc::tmem_ld_wait();
        l_dir = fmaf(l_dir, corr_prev, __uint_as_float(l));
      } else
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], corr_prev, __uint_as_float(a[i]));
#pragma unroll
      for (int i = 0; i < NO - 32; ++i) o_acc[32 + i] = fmaf(o_acc[32 + i], corr_prev, __uint_as_float(c[i]));
    };

    for (int j = 0; j < nk; ++j) {
      const int k0 = j * WS_BK;
      // key validity of this tile as four ballot words (bit i of word c = key k0 + 32 c + i)
      uint32_t vw[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
      bool all_valid = !mask && k0 + WS_BK <= p.Lk;
      if (!all_valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int key = k0 + c * 32 + lane;
          vw[c] = __ballot_sync(0xFFFFFFFFu, key < p.Lk && !(mask && mask[key]));
        }
        all_valid = (vw[0] & vw[1] & vw[2] & vw[3]) == 0xFFFFFFFFu;
      }
      tc::mbar_wait(tc::smem_u32(&bar_s[t]), j & 1);
      tc::fence_after_sync();
      if (j > 0) fold_o();
      if (p.dbg & 1) {
        tc::fence_before_sync();
        tc::mbar_arrive(tc::smem_u32(&bar_p[t]));
        continue;
      }

      // ---- pass 1: row maximum (a short last tile: only the 32-key chunks that hold keys)
      const int nch = (min(WS_BK, p.Lk - k0) + 31) >> 5;
      float mx = -INFINITY;
      if (nch < 4) mx = softmax_short_max<PARTS>(tS, vw, nch);
      else
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
        uint32_t a0[32], a1[32];
        tc::tmem_ld32(tS + c * 32, a0);
        tc::tmem_ld32(tS + (c + 1) * 32, a1);
        tc::tmem_ld_wait();
        if (all_valid) {
          float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(a0[i]), __uint_as_float(a1[i])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(a0[i + 1]), __uint_as_float(a1[i + 1])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(a0[i + 2]), __uint_as_float(a1[i + 2])));
            m3 = fmaxf(m3, fmaxf(__uint_as_float(a0[i + 3]), __uint_as_float(a1[i + 3])));
          }
          mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if ((vw[c] >> i) & 1u) mx = fmaxf(mx, __uint_as_float(a0[i]));
            if ((vw[c + 1] >> i) & 1u) mx = fmaxf(mx, __uint_as_float(a1[i]));
          }
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const bool dead = m_new == -INFINITY;  // every key so far is masked
      const float corr = dead ? 1.f : tc::ex2_approx(m_run - m_new);
      const float neg_m = dead ? 0.f : -m_new;
      const unsigned long long neg_m2 = tc::pack_f32x2(neg_m, neg_m);

      // ---- pass 2: p = 2^(s - m), split into bf16 hi / lo, stored over the scores in place
      if (all_valid) {
        softmax_pass2<PARTS, false>(tS, neg_m2, vw, dead);
      } else if (nch < 4) {
        softmax_short_exp<PARTS>(tS, neg_m, vw, dead, nch);
      } else {
        softmax_pass2<PARTS, true>(tS, neg_m2, vw, dead);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc::mbar_arrive(tc::smem_u32(&bar_p[t]));
      corr_prev = corr;
      m_run = m_new;
    }
    tc::mbar_wait(tc::smem_u32(&bar_s[t]), nk & 1);
    tc::fence_after_sync();
    fold_o();

    const int q = (qt0 + t) * AT_BM + row;
    if (q < p.Lq) {
      const long long o_off = b * p.so_b + static_cast<long long>(q) * p.ldo + h * HD;
      const float l_run = DIRECT ? l_dir : o_acc[AT_HD];
      const float inv = 1.0f / l_run;  // l == 0 (every key masked) -> NaN like the reference softmax
      if (DIRECT && HD == AT_HD && (h & 1)) {  // odd heads: the dims sit 4 columns up (static register indices in both branches)
#pragma unroll
        for (int d = 0; d < AT_HD; ++d) o_acc[d] = o_acc[d + 4];
      }
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 o4 = make_float4(o_acc[d] * inv, o_acc[d + 1] * inv, o_acc[d + 2] * inv, o_acc[d + 3] * inv);
        if (l_run == 0.f) o4 = make_float4(NAN, NAN, NAN, NAN);
        if (p.o16) {  // fp16 rows: the out-projection reads them as its operand (8-byte aligned: ld % 4 == 0)
          __half2 h0 = __floats2half2_rn(o4.x, o4.y), h1 = __floats2half2_rn(o4.z, o4.w);
          *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(p.O) + o_off + d) =
              make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
        } else {
          *reinterpret_cast<float4 *>(p.O + o_off + d) = o4;
        }
      }
    }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, TM_COLS);
}


// ------------------------------------------------------------------ short key sequences (Lk <= 128)
// One key tile: no online softmax, no running accumulators — and no reason to spend 256 TMEM columns and 160
// registers per thread on a CTA.  The launches with 80 text tokens as keys (1024 x 80, 256 x 80, 80 x 80) do almost
// no arithmetic and are bound by the fixed latency chain of a CTA (copies -> S -> softmax -> P.V -> write-out) at two
// CTAs per SM.  This variant keeps P in SHARED memory (over the Q and K tiles, dead once S is complete) so that O can
// be accumulated over the score columns: 128 TMEM columns, 53 KB of shared memory and <= 85 registers per thread,
// i.e. FOUR CTAs per SM.  fp16 K / V rows by tensor copy as in the DIRECT variant (same column shift for odd
// heads), same arithmetic: p = 2^(s - max), fp16 P, denominators from the ones tile.
// Up to SH_EXTRA keys beyond the tile (Lk = 132 detected boxes: 4) are handled by the row's own thread on the
// FMA pipe: their scores from the row's Q values (read back from the operand tile) and the fp16 K rows in global
// memory (fp16 x fp16 products are exact in fp32, as on the tensor core), their probabilities rounded to fp16
// like P, their share of P.V and of the denominator added to the accumulator columns read back from TMEM.
constexpr int SH_EXTRA = 8;
constexpr uint32_t SH_SMEM = 3 * QK_PART + 2 * ONES_BLK + 1024;  // Q | K | V | ones
__global__ void __launch_bounds__(WS_THREADS1, 4) attention_short_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sQ = smem, *sK = sQ + QK_PART, *sV = sK + QK_PART, *sOnes = sV + QK_PART;
  unsigned char *sP = smem;  // two blocks of 64 keys x 128 rows x 128 B over Q | K
  __shared__ __align__(8) unsigned long long bar_q, bar_kf, bar_vf, bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const int lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  const int shift = (h & 1) * 4;
  const int nch = (min(WS_BK, p.Lk) + 31) >> 5;  // 32-key chunks that hold keys

  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), 128);
  if (tid == 32) {
    tc::mbar_init(tc::smem_u32(&bar_q), 1);
    tc::mbar_init(tc::smem_u32(&bar_kf), 1);
    tc::mbar_init(tc::smem_u32(&bar_vf), 1);
    tc::mbar_init(tc::smem_u32(&bar_s), 1);
    tc::mbar_init(tc::smem_u32(&bar_p), 128);
    tc::mbar_init(tc::smem_u32(&bar_o), 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, tmem_base_s, 0);
  bd::pdl_launch_dependents();
  bd::pdl_wait();

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(tc::smem_u32(&bar_kf), QK_PART);
      tc::tma_load_2d(tc::smem_u32(sK), &p.tmK, h * AT_HD - shift, b * p.Lk, tc::smem_u32(&bar_kf));
      tc::mbar_arrive_expect_tx(tc::smem_u32(&bar_vf), QK_PART);
      tc::tma_load_2d(tc::smem_u32(sV), &p.tmV, h * AT_HD - shift, b * p.Lk, tc::smem_u32(&bar_vf));
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t idesc_s = tc::idesc_ab(1, AT_BM, 32 * nch), idesc_o = tc::idesc_ab(1, AT_BM, AT_NV) | tc::idesc_b_mn,
                   idesc_1 = tc::idesc_ab(1, AT_BM, 16);
    tc::mbar_wait(tc::smem_u32(&bar_q), 0);
    tc::mbar_wait(tc::smem_u32(&bar_kf), 0);
    tc::fence_after_sync();
    if (tc::elect_one()) {
#pragma unroll
      for (int s = 0; s < AT_NV / 16; ++s)
        tc::mma_bf16(tmem, tc::smem_desc_sw128(tc::smem_u32(sQ) + s * 32), tc::smem_desc_sw128(tc::smem_u32(sK) + s * 32), idesc_s,
                     s > 0 ? 1u : 0u);
      tc::mma_commit(tc::smem_u32(&bar_s));
    }
    __syncwarp();
    tc::mbar_wait(tc::smem_u32(&bar_p), 0);
    tc::mbar_wait(tc::smem_u32(&bar_vf), 0);
    tc::fence_after_sync();
    if (tc::elect_one()) {
      for (int s = 0; s < 2 * nch; ++s) {  // 16 keys per step
        const uint64_t dp = tc::smem_desc_sw128(tc::smem_u32(sP) + (s >> 2) * QK_PART + (s & 3) * 32);
        tc::mma_bf16(tmem, dp, tc::smem_desc_sw128_mn(tc::smem_u32(sV) + s * 2048), idesc_o, s > 0 ? 1u : 0u);
        tc::mma_bf16(tmem + AT_NV, dp, tc::smem_desc_sw128(tc::smem_u32(sOnes) + (s >> 2) * ONES_BLK + (s & 3) * 32), idesc_1,
                     s > 0 ? 1u : 0u);
      }
      tc::mma_commit(tc::smem_u32(&bar_o));
    }
    __syncwarp();
  } else {
    // ---- Q rows -> operand tile (scale * log2 e folded in; zero outside the head's 36 dims), ones tile
    {
      const long long q_off = b * p.sq_b + h * AT_HD;
      for (int e = tid - 64; e < AT_BM * 6; e += 128) {
        const int rr = e / 6, ch = e - rr * 6;
        const int q = qt * AT_BM + rr;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int d0 = ch * 8 - shift;
        const bool first = d0 >= 0 && d0 < AT_HD, second = d0 + 4 < AT_HD;
        if (q < p.Lq && (first || second)) {
          load_chunk8(p.Q, q_off + static_cast<long long>(q) * p.ldq + d0, p.q16, second, v, first);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] *= p.scale_log2;
        }
        uint4 hi, lo;
        tc::cvt8(1, v, hi, lo);
        *reinterpret_cast<uint4 *>(sQ + tc::sw128_off(rr, ch)) = hi;
      }
      for (int e = tid - 64; e < 2 * static_cast<int>(ONES_BLK) / 16; e += 128) {
        const uint32_t one2 = (e & 127) < 8 ? 0x3C003C00u : 0u;
        *reinterpret_cast<uint4 *>(sOnes + e * 16) = make_uint4(one2, one2, one2, one2);
      }
      tc::fence_proxy_async_smem();
      // a barrier of the four staging warps rather than four mbarrier arrivals: each thread later reads and
      // overwrites the Q row other threads staged, and this orders those accesses in a way racecheck follows too
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid == 64) tc::mbar_arrive(tc::smem_u32(&bar_q));
    }
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tS = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const unsigned char *mask = p.mask ? p.mask + static_cast<long long>(b) * p.Lk : nullptr;
    uint32_t vw[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int key = c * 32 + lane;
      vw[c] = __ballot_sync(0xFFFFFFFFu, key < p.Lk && !(mask && mask[key]));
    }
    tc::mbar_wait(tc::smem_u32(&bar_s), 0);
    tc::fence_after_sync();
    float mx = -INFINITY;
    // ---- keys beyond the tile: scores on the FMA pipe (before this row's Q values are overwritten by P)
    const int n_extra = max(0, p.Lk - WS_BK);
    const int col0 = h * AT_HD - shift;  // global column of tile column 0 (a multiple of 8)
    float s_x[SH_EXTRA], p_x[SH_EXTRA];
#pragma unroll
    for (int e = 0; e < SH_EXTRA; ++e) s_x[e] = -INFINITY, p_x[e] = 0.f;
    if (n_extra > 0) {
      uint4 qh[6];
#pragma unroll
      for (int ch = 0; ch < 6; ++ch) qh[ch] = *reinterpret_cast<const uint4 *>(sQ + tc::sw128_off(row, ch));
      const __half *Kg = reinterpret_cast<const __half *>(p.K) + b * p.sk_b + col0;
#pragma unroll
      for (int e = 0; e < SH_EXTRA; ++e) {
        if (e < n_extra && !(mask && mask[WS_BK + e])) {
          const __half *kr = Kg + static_cast<long long>(WS_BK + e) * p.ldk;
          float acc = 0.f;
#pragma unroll
          for (int ch = 0; ch < 6; ++ch) {
            if (col0 + ch * 8 + 8 <= p.H * AT_HD) {  // inside the K rows (the Q values beyond are zero anyway)
              const uint4 kv = __ldg(reinterpret_cast<const uint4 *>(kr + ch * 8));
              const __half2 *q2 = reinterpret_cast<const __half2 *>(&qh[ch]), *k2 = reinterpret_cast<const __half2 *>(&kv);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 qf = __half22float2(q2[i]), kf = __half22float2(k2[i]);
                acc = fmaf(qf.x, kf.x, acc), acc = fmaf(qf.y, kf.y, acc);
              }
            }
          }
          s_x[e] = acc;
          mx = fmaxf(mx, acc);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < nch) {
        uint32_t a[32];
        tc::tmem_ld32(tS + c * 32, a);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if ((vw[c] >> i) & 1u) mx = fmaxf(mx, __uint_as_float(a[i]));
      }
    }
    const bool dead = mx == -INFINITY;  // every key masked
    const float neg_m = dead ? 0.f : -mx;
    if (n_extra > 0) {
#pragma unroll
      for (int e = 0; e < SH_EXTRA; ++e)
        if (s_x[e] != -INFINITY) p_x[e] = __half2float(__float2half_rn(tc::ex2_approx(s_x[e] + neg_m)));  // rounded like P
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < nch) {
        uint32_t a[32], o[16];
        tc::tmem_ld32(tS + c * 32, a);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = tc::ex2_approx(__uint_as_float(a[i]) + neg_m), p1 = tc::ex2_approx(__uint_as_float(a[i + 1]) + neg_m);
          if (dead || !((vw[c] >> i) & 1u)) p0 = 0.f;
          if (dead || !((vw[c] >> (i + 1)) & 1u)) p1 = 0.f;
          o[i >> 1] = tc::pack_f16x2(p0, p1);
        }
        unsigned char *blk = sP + (c >> 1) * QK_PART;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<uint4 *>(blk + tc::sw128_off(row, (c & 1) * 4 + q4)) = make_uint4(o[4 * q4], o[4 * q4 + 1], o[4 * q4 + 2], o[4 * q4 + 3]);
      }
    }
    tc::fence_before_sync();        // the scores have been read: O may overwrite their columns
    tc::fence_proxy_async_smem();   // P is read by the tensor core (async proxy)
    tc::mbar_arrive(tc::smem_u32(&bar_p));
    tc::mbar_wait(tc::smem_u32(&bar_o), 0);
    tc::fence_after_sync();
    {
      uint32_t a[32], c8[8], l;
      tc::tmem_ld32(tS, a);
      tc::tmem_ld8(tS + 32, c8);
      tc::tmem_ld1(tS + AT_NV, l);
      tc::tmem_ld_wait();
      const int q = qt * AT_BM + row;
      if (q < p.Lq) {
        float o[40];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(a[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[32 + i] = __uint_as_float(c8[i]);
        float l_run = __uint_as_float(l);
        if (n_extra > 0) {  // the extra keys' share of P.V and of the denominator (tile column j = global column col0 + j)
          const __half *Vg = reinterpret_cast<const __half *>(p.V) + b * p.sv_b + col0;
#pragma unroll
          for (int e = 0; e < SH_EXTRA; ++e) {
            if (p_x[e] != 0.f) {
              const __half *vr = Vg + static_cast<long long>(WS_BK + e) * p.ldv;
              l_run += p_x[e];
#pragma unroll
              for (int ch = 0; ch < 5; ++ch) {
                if (col0 + ch * 8 + 8 <= p.H * AT_HD) {
                  const uint4 vv = __ldg(reinterpret_cast<const uint4 *>(vr + ch * 8));
                  const __half2 *v2 = reinterpret_cast<const __half2 *>(&vv);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 vf = __half22float2(v2[i]);
                    o[ch * 8 + 2 * i] = fmaf(p_x[e], vf.x, o[ch * 8 + 2 * i]), o[ch * 8 + 2 * i + 1] = fmaf(p_x[e], vf.y, o[ch * 8 + 2 * i + 1]);
                  }
                }
              }
            }
          }
        }
        if (shift) {
#pragma unroll
          for (int d = 0; d < AT_HD; ++d) o[d] = o[d + 4];
        }
        const float inv = 1.0f / l_run;  // l == 0 (every key masked) -> NaN like the reference softmax
        const long long o_off = b * p.so_b + static_cast<long long>(q) * p.ldo + h * AT_HD;
#pragma unroll
        for (int d = 0; d < AT_HD; d += 4) {
          float4 o4 = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
          if (l_run == 0.f) o4 = make_float4(NAN, NAN, NAN, NAN);
          if (p.o16) {
            __half2 h0 = __floats2half2_rn(o4.x, o4.y), h1 = __floats2half2_rn(o4.z, o4.w);
            *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(p.O) + o_off + d) =
                make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
          } else {
            *reinterpret_cast<float4 *>(p.O + o_off + d) = o4;
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

}  // namespace

// Implementation switch (A/B measurements and the parity tests of both kernels): 1 = the
// warp-specialised kernel above (default), 0 = the first-generation single-role kernel.
static int g_attn_impl = 1, g_attn_dbg = 0;
// key tiles (of 128) up to which the one-tile-per-CTA, two-CTAs-per-SM variant runs in the fp16 mode.  Measured at
// 128 scenes (tools/microbench_attention_tiles.py): it wins at EVERY shape of the model — 1024x1024 0.633 vs 0.698 ms,
// 80x1024 0.159 vs 0.187, 80x80 0.037 vs 0.046 — so the default is "always".  The bf16x3 mode keeps two ping-ponged
// tiles per CTA: its operand tiles (hi + lo) leave room for one CTA per SM either way.
static int g_attn_small_nk = 1 << 30;
static int g_attn_short = 1;   // one key tile (Lk <= 128), fp16 K / V, head dim 36: the four-CTAs-per-SM kernel
extern "C" int bd_attention_tc_set_short(int on) {
  g_attn_short = on != 0;
  return BD_OK;
}
static int g_attn_direct = 1;  // fp16 K / V: tensor copies from the projection output instead of the pack kernel
extern "C" int bd_attention_tc_set_direct(int on) {
  g_attn_direct = on != 0;
  return BD_OK;
}
extern "C" int bd_attention_tc_set_small_nk(int nk) {
  g_attn_small_nk = nk;
  return BD_OK;
}
extern "C" int bd_attention_tc_select(int impl) {
  BD_REQUIRE(impl >= 0 && ((impl & 15) == 0 || (impl & 15) == 1), "bd_attention_tc_select: impl must be 0 or 1");
  g_attn_impl = impl & 15;
#ifdef BD_ATTN_EXPERIMENTS  // tools/attn_experiments.py: bits 4.. switch the softmax / the two MMAs off (timing only, garbage results)
  g_attn_dbg = impl >> 4;
#else
  BD_REQUIRE((impl >> 4) == 0, "bd_attention_tc_select: the timing-experiment bits need a -DBD_ATTN_EXPERIMENTS build");
#endif
  return BD_OK;
}

// key-tile width of the legacy kernel: bf16x3 -> 64-key tiles, single K/V stage (two CTAs per SM);
// bf16 -> 128-key tiles, double-buffered.  The warp-specialised kernel always uses 128-key tiles.
inline int attn_bk(int split) { return (g_attn_impl == 0 && split == 3) ? 64 : 128; }

extern "C" long long bd_attention_tc_workspace_bytes(int B, int H, int Lq, int Lk, int split) {
  const long long parts = split == 3 ? 2 : 1;
  const int BK = 128;  // upper bound of both tilings
  const long long nq = (Lq + 127) / 128, nk = (Lk + BK - 1) / BK;
  return static_cast<long long>(B) * H * parts * (nq * QK_PART + nk * (k_part(BK) + v_part(BK)));
}

// phase: 1 = pack K / V into the workspace, 2 = attention over a packed workspace, 3 = both
static int attention_tc_phases(const float *Q, int ldq, long long sq_b, const float *K, int ldk, long long sk_b,
                               const float *V, int ldv, long long sv_b, const unsigned char *key_padding_mask,
                               float *O, int ldo, long long so_b, int B, int H, int Lq, int Lk, int hd, float scale,
                               int split, void *workspace, bd_stream_t stream, int phase, int io16 = 0);

extern "C" int bd_attention_tc(const float *Q, int ldq, long long sq_b, const float *K, int ldk, long long sk_b,
                               const float *V, int ldv, long long sv_b, const unsigned char *key_padding_mask,
                               float *O, int ldo, long long so_b, int B, int H, int Lq, int Lk, int hd, float scale,
                               int split, void *workspace, bd_stream_t stream) {
  return attention_tc_phases(Q, ldq, sq_b, K, ldk, sk_b, V, ldv, sv_b, key_padding_mask, O, ldo, so_b, B, H, Lq, Lk, hd,
                             scale, split, workspace, stream, 3);
}

// bd_attention_tc with 16-bit tensors in HBM: io_half bit 0 = Q, bit 1 = K, bit 2 = V, bit 3 = O are fp16 (their
// leading dimensions and batch strides count halfs; rows 8-byte aligned).  Values: Q / K / V are rounded to fp16
// inside the kernel anyway; an fp16 O is what the out-projection would round it to.
extern "C" int bd_attention_tc_h(const void *Q, int ldq, long long sq_b, const void *K, int ldk, long long sk_b,
                                 const void *V, int ldv, long long sv_b, const unsigned char *key_padding_mask, void *O,
                                 int ldo, long long so_b, int io_half, int B, int H, int Lq, int Lk, int hd, float scale,
                                 int split, void *workspace, bd_stream_t stream) {
  BD_REQUIRE(g_attn_impl == 1, "bd_attention_tc_h: only with the warp-specialised kernel");
  return attention_tc_phases(static_cast<const float *>(Q), ldq, sq_b, static_cast<const float *>(K), ldk, sk_b,
                             static_cast<const float *>(V), ldv, sv_b, key_padding_mask, static_cast<float *>(O), ldo, so_b, B,
                             H, Lq, Lk, hd, scale, split, workspace, stream, 3, io_half & 15);
}

// The two halves of bd_attention_tc for keys / values that are known long before their queries (the
// decoder's memory K / V): pack once, on another stream, off the critical path; attend later.  Same Lq,
// Lk, split and workspace in both calls (the workspace layout depends on them).
extern "C" int bd_attention_tc_pack_kv(const void *K, int ldk, long long sk_b, const void *V, int ldv, long long sv_b,
                                       int kv_half, int B, int H, int Lq, int Lk, int hd, int split, void *workspace,
                                       bd_stream_t stream) {
  BD_REQUIRE(g_attn_impl == 1, "bd_attention_tc_pack_kv: only with the warp-specialised kernel");
  return attention_tc_phases(static_cast<const float *>(K), 4, 0, static_cast<const float *>(K), ldk, sk_b,
                             static_cast<const float *>(V), ldv, sv_b, nullptr, reinterpret_cast<float *>(workspace), 4, 0, B, H,
                             Lq, Lk, hd, 1.0f, split, workspace, stream, 1, kv_half ? 6 : 0);
}
extern "C" int bd_attention_tc_packed(const void *Q, int ldq, long long sq_b, const unsigned char *key_padding_mask,
                                      void *O, int ldo, long long so_b, int io_half, int B, int H, int Lq, int Lk, int hd,
                                      float scale, int split, void *workspace, bd_stream_t stream) {
  BD_REQUIRE(g_attn_impl == 1, "bd_attention_tc_packed: only with the warp-specialised kernel");
  return attention_tc_phases(static_cast<const float *>(Q), ldq, sq_b, static_cast<const float *>(Q), 4, 0,
                             static_cast<const float *>(Q), 4, 0, key_padding_mask, static_cast<float *>(O), ldo, so_b, B, H,
                             Lq, Lk, hd, scale, split, workspace, stream, 2, io_half & 9);
}

static int attention_tc_phases(const float *Q, int ldq, long long sq_b, const float *K, int ldk, long long sk_b,
                               const float *V, int ldv, long long sv_b, const unsigned char *key_padding_mask,
                               float *O, int ldo, long long so_b, int B, int H, int Lq, int Lk, int hd, float scale,
                               int split, void *workspace, bd_stream_t stream, int phase, int io16) {
  BD_REQUIRE(Q && K && V && O, "bd_attention_tc: null pointer");
  BD_REQUIRE(io16 == 0 || (g_attn_impl == 1 && ldv % 4 == 0 && sv_b % 4 == 0 && (reinterpret_cast<uintptr_t>(V) & 7) == 0),
             "bd_attention_tc: fp16 tensors need the warp-specialised kernel and 8-byte aligned rows");
  BD_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0 && B <= 65535 && H <= 65535, "bd_attention_tc: bad sizes");
  BD_REQUIRE(hd == AT_HD || hd == 64, "bd_attention_tc: built for head_dim 36 (d_model 288 / 8 heads) and 64 (fp16 K / V only)");
  BD_REQUIRE(split == 1 || split == 3, "bd_attention_tc: split must be 1 (bf16) or 3 (bf16x3)");
  BD_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && sq_b % 4 == 0 && sk_b % 4 == 0 &&
                 (reinterpret_cast<uintptr_t>(Q) & 15) == 0 && (reinterpret_cast<uintptr_t>(K) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
             "bd_attention_tc: Q / K rows and the workspace must be 16-byte aligned");
  const int impl = g_attn_impl;
  BD_REQUIRE(hd == AT_HD || impl == 1, "bd_attention_tc: head_dim 64 only with the warp-specialised kernel");
  BD_REQUIRE(impl == 0 || (ldo % 4 == 0 && so_b % 4 == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0),
             "bd_attention_tc: output rows must be 16-byte aligned");
  const int BK = attn_bk(split);
  AttnParams p = {};
  p.Q = Q, p.K = K, p.V = V, p.mask = key_padding_mask, p.O = O;
  p.ldq = ldq, p.ldk = ldk, p.ldv = ldv, p.ldo = ldo;
  p.sq_b = sq_b, p.sk_b = sk_b, p.sv_b = sv_b, p.so_b = so_b;
  p.Lq = Lq, p.Lk = Lk, p.H = H;
  p.nq = bd::ceil_div(Lq, AT_BM), p.nk = bd::ceil_div(Lk, BK);
  p.scale_log2 = scale * 1.4426950408889634f;
  p.q16 = io16 & 1, p.k16 = (io16 >> 1) & 1, p.v16 = (io16 >> 2) & 1, p.o16 = (io16 >> 3) & 1;
  p.dbg = g_attn_dbg;
  const size_t parts = split == 3 ? 2 : 1;
  unsigned char *ws = static_cast<unsigned char *>(workspace);
  p.Qp = ws;
  p.Kp = p.Qp + static_cast<size_t>(B) * H * p.nq * parts * QK_PART;
  p.Vp = p.Kp + static_cast<size_t>(B) * H * p.nk * parts * k_part(BK);
  constexpr size_t WS_SMEM1 = 2 * (QK_PART + k_part(128) + v_part(128)) + 1024, WS_SMEM2 = 2 * WS_SMEM1 - 1024;
  // one query tile per CTA: Q tile + two stages of K and V^T
  constexpr size_t WS_SMEM1S = QK_PART + 2 * (k_part(128) + v_part(128)) + 1024, WS_SMEM2S = 2 * WS_SMEM1S - 1024;
  // DIRECT: Q tile + two stages of (K tile, V tile of 128 keys x 128 B) + the ones tile
  constexpr size_t WS_SMEM_DIRECT = QK_PART + 2 * (k_part(128) + 128 * 128) + 2 * ONES_BLK + 1024;
  static bd::PerDeviceOnce configured;  // function attributes are per device
  BD_CUDA(configured.run([&]() {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<1, 128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<2, 64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_ws_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_ws_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM2);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_ws_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM1S);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_ws_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM2S);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_ws_kernel<1, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_DIRECT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_ws_kernel<1, 1, true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_DIRECT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_short_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_SMEM);
    return e;
  }), "bd_attention_tc");
  cudaStream_t s = bd::as_stream(stream);
  dim3 pgrid(p.nq + 2 * p.nk, H, B), grid(p.nq, H, B), wgrid(bd::ceil_div(p.nq, 2), H, B);
  if (impl == 1) {
    p.skip_q = 1;
    pgrid.x = 2 * p.nk;
    const bool small = (parts == 1 || g_attn_small_nk < (1 << 30)) && p.nk <= g_attn_small_nk;  // one query tile per CTA, two CTAs per SM
    // fp16 K and V rows in HBM, dense batches: no pack kernel — the attention kernel's loader fetches the
    // operand tiles from the projection output by tensor copies (phase 1 then has nothing to do)
    const bool direct = (g_attn_direct || hd == 64) && parts == 1 && small && phase == 3 && p.k16 && p.v16 && ldk % 8 == 0 && ldv % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(K) & 15) == 0 && (reinterpret_cast<uintptr_t>(V) & 15) == 0 &&
                        sk_b == static_cast<long long>(Lk) * ldk && sv_b == static_cast<long long>(Lk) * ldv;
    BD_REQUIRE(direct || workspace, "bd_attention_tc: null workspace");
    BD_REQUIRE(direct || hd == AT_HD, "bd_attention_tc: head_dim 64 needs fp16 K and V rows (io_half bits 1, 2), split 1, dense "
                                      "batches, ldk / ldv multiples of 8 and 16-byte aligned K / V");
    if (direct) {
      tc::EncodeTiledFn enc = tc::encode_tiled();
      BD_REQUIRE(enc != nullptr, "bd_attention_tc: cuTensorMapEncodeTiled is not available from this driver");
      const cuuint64_t dims[2] = {static_cast<cuuint64_t>(H) * hd, static_cast<cuuint64_t>(B) * Lk};
      const cuuint32_t box[2] = {64, WS_BK}, estr[2] = {1, 1};
      for (int which = 0; which < 2; ++which) {
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(which ? ldv : ldk) * sizeof(__half)};
        const CUresult r = enc(which ? &p.tmV : &p.tmK, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                               const_cast<float *>(which ? V : K), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        BD_REQUIRE(r == CUDA_SUCCESS, "bd_attention_tc: cuTensorMapEncodeTiled failed (%d) for B=%d Lk=%d ld=%d",
                   static_cast<int>(r), B, Lk, which ? ldv : ldk);
      }
      if (hd == AT_HD && Lk <= WS_BK + SH_EXTRA && g_attn_short)
        BD_CUDA(bd::launch_pdl(attention_short_kernel, grid, dim3(WS_THREADS1), SH_SMEM, s, p), "bd_attention_tc");
      else if (hd == 64)
        BD_CUDA(bd::launch_pdl(attention_ws_kernel<1, 1, true, 64>, grid, dim3(WS_THREADS1), WS_SMEM_DIRECT, s, p), "bd_attention_tc");
      else
        BD_CUDA(bd::launch_pdl(attention_ws_kernel<1, 1, true>, grid, dim3(WS_THREADS1), WS_SMEM_DIRECT, s, p), "bd_attention_tc");
    } else if (parts == 2) {
      if (phase & 1) BD_CUDA(bd::launch_pdl(attention_pack_kernel<2, 128>, pgrid, dim3(256), 0, s, p), "bd_attention_tc");
      if (!(phase & 2)) {
      } else if (small)
        BD_CUDA(bd::launch_pdl(attention_ws_kernel<2, 1>, grid, dim3(WS_THREADS1), WS_SMEM2S, s, p), "bd_attention_tc");
      else
        BD_CUDA(bd::launch_pdl(attention_ws_kernel<2, 2>, wgrid, dim3(WS_THREADS), WS_SMEM2, s, p), "bd_attention_tc");
    } else {
      if (phase & 1) BD_CUDA(bd::launch_pdl(attention_pack_kernel<1, 128>, pgrid, dim3(256), 0, s, p), "bd_attention_tc");
      if (!(phase & 2)) {
      } else if (small)
        BD_CUDA(bd::launch_pdl(attention_ws_kernel<1, 1>, grid, dim3(WS_THREADS1), WS_SMEM1S, s, p), "bd_attention_tc");
      else
        BD_CUDA(bd::launch_pdl(attention_ws_kernel<1, 2>, wgrid, dim3(WS_THREADS), WS_SMEM1, s, p), "bd_attention_tc");
    }
  } else if (!workspace) {
    BD_REQUIRE(false, "bd_attention_tc: null workspace");
  } else if (parts == 2) {
    const size_t smem = 2 * (QK_PART + (k_part(64) + v_part(64)) + p_part(64)) + 1024;
    attention_pack_kernel<2, 64><<<pgrid, 256, 0, s>>>(p);
    attention_tc_kernel<2, 64, 1><<<grid, AT_THREADS, smem, s>>>(p);
  } else {
    const size_t smem = QK_PART + 2 * (k_part(128) + v_part(128)) + p_part(128) + 1024;
    attention_pack_kernel<1, 128><<<pgrid, 256, 0, s>>>(p);
    attention_tc_kernel<1, 128, 2><<<grid, AT_THREADS, smem, s>>>(p);
  }
  BD_CHECK_LAUNCH("bd_attention_tc");
  return BD_OK;
}
