// Multi-head attention core on tcgen05 — softmax(Q·Kᵀ·scale + mask)·V for nn.MultiheadAttention
// (call sites /root/reference/models/encoder_decoder_layers.py:87,99,111,149,179,365,373,384,394).
//
// One CTA = 128 queries x 1 head x 1 scene; thread t owns query row t (= TMEM lane t).
// Per tile of 128 keys:
//   S  = Q·Kᵀ        tcgen05.mma  M=128, N=128, K=64 (head_dim 36 zero-padded to one swizzle block)
//   softmax           tcgen05.ld of the thread's score row, exact online softmax in fp32
//                     (running max / sum in registers), probabilities written back to shared
//                     memory as the bf16 A operand of the second MMA
//   Ot = P·V          tcgen05.mma  M=128, N=48, K=128 keys into a scratch TMEM accumulator
//   O  = O·corr + Ot  in registers (36 fp32 per thread), so no TMEM rescaling pass is needed
// Operands are bf16 hi (+ lo in the "bf16x3" mode: 3 MMAs per product, fp32-grade) in the
// 128-byte-swizzle K-major layout of tc_common.cuh; Vᵀ is produced by a transposing stage.
// The softmax scale (and log2 e) is folded into Q, like torch scales q before QKᵀ.
#include "tc_common.cuh"

namespace {

constexpr int AT_BM = 128, AT_BK = 128, AT_NV = 48, AT_THREADS = 128;
constexpr uint32_t QK_PART = 128 * 128;        // 128 rows x 128 B
constexpr uint32_t V_BLK = AT_NV * 128;        // 48 rows x 128 B (one block of 64 keys)
constexpr uint32_t V_PART = 2 * V_BLK;
constexpr uint32_t P_BLK = 128 * 128;
constexpr uint32_t P_PART = 2 * P_BLK;

struct AttnParams {
  const float *Q, *K, *V;
  const unsigned char *mask;
  float *O;
  int ldq, ldk, ldv, ldo;
  long long sq_b, sk_b, sv_b, so_b;
  int Lq, Lk, hd;
  float scale_log2;
};

template <int PARTS>
__global__ void __launch_bounds__(AT_THREADS, 1) attention_tc_kernel(const AttnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sQ = smem;
  unsigned char *sK = sQ + PARTS * QK_PART;
  unsigned char *sV = sK + PARTS * QK_PART;
  unsigned char *sP = sV + PARTS * V_PART;  // 1024-aligned: V_PART = 12 KB
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ unsigned char kvalid[AT_BK];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_BM;
  constexpr int hd = 36;  // enforced by the host wrapper (d_model 288 / 8 heads)
  const float *Qg = p.Q + b * p.sq_b + h * hd;
  const float *Kg = p.K + b * p.sk_b + h * hd;
  const float *Vg = p.V + b * p.sv_b + h * hd;
  const unsigned char *mask = p.mask ? p.mask + static_cast<long long>(b) * p.Lk : nullptr;

  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), 256);
  if (tid == 32) {
    tc::mbar_init(tc::smem_u32(&bar), 1);
    tc::fence_mbar_init();
  }
  // zero the operand tiles once: padded head dims / padded V rows stay zero for the whole kernel
  for (uint32_t i = tid; i < (PARTS * (2 * QK_PART + V_PART)) / 16; i += AT_THREADS)
    reinterpret_cast<uint4 *>(sQ)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, tmem_base_s, 0);
  const uint32_t tmem_s = tmem, tmem_o = tmem + 128;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;

  // rows of Q / K: (row, 16-byte chunk) items, chunk = 8 head dims; hd = 36 -> chunks 0..4
  const int n_ch = (hd + 7) / 8;
  auto stage_rows = [&](unsigned char *dst, const float *src, int ld, int row0, int n_rows, float mul) {
    for (int e = tid; e < AT_BM * n_ch; e += AT_THREADS) {
      const int r = e % AT_BM, ch = e / AT_BM;  // consecutive threads -> consecutive rows (conflict-free stores)
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      if (row0 + r < n_rows) {
        const float *s = src + static_cast<long long>(row0 + r) * ld + ch * 8;
        const int nv = min(8, hd - ch * 8);
        if (nv == 8) {
          const float4 a = __ldg(reinterpret_cast<const float4 *>(s)), c = __ldg(reinterpret_cast<const float4 *>(s) + 1);
          v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = c.x, v[5] = c.y, v[6] = c.z, v[7] = c.w;
        } else {
          for (int i = 0; i < nv; ++i) v[i] = __ldg(s + i);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= mul;
      }
      uint4 hi, lo;
      tc::split_bf16x8(v, hi, lo);
      const uint32_t off = tc::sw128_off(r, ch);
      *reinterpret_cast<uint4 *>(dst + off) = hi;
      if (PARTS == 2) *reinterpret_cast<uint4 *>(dst + QK_PART + off) = lo;
    }
  };
  stage_rows(sQ, Qg, p.ldq, q0, p.Lq, p.scale_log2);

  float m_run = -INFINITY, l_run = 0.f;
  float o_acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) o_acc[i] = 0.f;
  uint32_t phase = 0;
  const uint32_t idesc_s = tc::idesc_bf16(AT_BM, AT_BK), idesc_o = tc::idesc_bf16(AT_BM, AT_NV);

  for (int k0 = 0; k0 < p.Lk; k0 += AT_BK) {
    // ---- stage K tile (rows = keys) and V^T tile (rows = head dims, K = keys)
    stage_rows(sK, Kg, p.ldk, k0, p.Lk, 1.0f);
    for (int e = tid; e < hd * (AT_BK / 8); e += AT_THREADS) {
      const int d = e % hd, kc = e / hd;  // consecutive threads -> consecutive head dims (coalesced)
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int key = k0 + kc * 8 + i;
        v[i] = key < p.Lk ? __ldg(Vg + static_cast<long long>(key) * p.ldv + d) : 0.f;
      }
      uint4 hi, lo;
      tc::split_bf16x8(v, hi, lo);
      const uint32_t off = (kc >> 3) * V_BLK + tc::sw128_off(d, kc & 7);
      *reinterpret_cast<uint4 *>(sV + off) = hi;
      if (PARTS == 2) *reinterpret_cast<uint4 *>(sV + V_PART + off) = lo;
    }
    kvalid[tid] = (k0 + tid < p.Lk) && !(mask && mask[k0 + tid]);
    tc::fence_proxy_async_smem();
    __syncthreads();

    // ---- S = Q K^T
    if (warp == 0) {
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t q = tc::smem_u32(sQ), k = tc::smem_u32(sK);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint64_t dq = tc::smem_desc_sw128(q + s * 32), dk = tc::smem_desc_sw128(k + s * 32);
          tc::mma_bf16(tmem_s, dq, dk, idesc_s, s > 0 ? 1u : 0u);
          if (PARTS == 2) {
            tc::mma_bf16(tmem_s, tc::smem_desc_sw128(q + QK_PART + s * 32), dk, idesc_s, 1u);
            tc::mma_bf16(tmem_s, dq, tc::smem_desc_sw128(k + QK_PART + s * 32), idesc_s, 1u);
          }
        }
        tc::mma_commit(tc::smem_u32(&bar));
      }
      __syncwarp();
    }
    tc::mbar_wait(tc::smem_u32(&bar), phase);
    phase ^= 1u;
    tc::fence_after_sync();

    // ---- online softmax on the thread's row (two passes over TMEM: max, then exp / sum / pack)
    float mx = -INFINITY;
#pragma unroll
    for (int g = 0; g < AT_BK / 16; ++g) {
      uint32_t acc[16];
      tc::tmem_ld16(tmem_s + lane_base + g * 16, acc);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (kvalid[g * 16 + j]) mx = fmaxf(mx, __uint_as_float(acc[j]));
    }
    const float m_new = fmaxf(m_run, mx);
    const float corr = (m_new == -INFINITY) ? 1.f : exp2f(m_run - m_new);
    float sum = 0.f;
    const int r = tid;
#pragma unroll
    for (int g = 0; g < AT_BK / 16; ++g) {
      uint32_t acc[16];
      tc::tmem_ld16(tmem_s + lane_base + g * 16, acc);
      tc::tmem_ld_wait();
      float pv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        pv[j] = (kvalid[g * 16 + j] && m_new != -INFINITY) ? exp2f(__uint_as_float(acc[j]) - m_new) : 0.f;
        sum += pv[j];
      }
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const float v8[8] = {pv[c2 * 8 + 0], pv[c2 * 8 + 1], pv[c2 * 8 + 2], pv[c2 * 8 + 3],
                             pv[c2 * 8 + 4], pv[c2 * 8 + 5], pv[c2 * 8 + 6], pv[c2 * 8 + 7]};
        uint4 hi, lo;
        tc::split_bf16x8(v8, hi, lo);
        const int kc = g * 2 + c2;  // 16-byte chunk of keys
        const uint32_t off = (kc >> 3) * P_BLK + tc::sw128_off(r, kc & 7);
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
        const uint32_t pa = tc::smem_u32(sP), va = tc::smem_u32(sV);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
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
        tc::mma_commit(tc::smem_u32(&bar));
      }
      __syncwarp();
    }
    tc::mbar_wait(tc::smem_u32(&bar), phase);
    phase ^= 1u;
    tc::fence_after_sync();
    {
      uint32_t a0[16], a1[16], a2[16];
      tc::tmem_ld16(tmem_o + lane_base, a0);
      tc::tmem_ld16(tmem_o + lane_base + 16, a1);
      tc::tmem_ld16(tmem_o + lane_base + 32, a2);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) o_acc[j] = fmaf(o_acc[j], corr, __uint_as_float(a0[j]));
#pragma unroll
      for (int j = 0; j < 16; ++j) o_acc[16 + j] = fmaf(o_acc[16 + j], corr, __uint_as_float(a1[j]));
#pragma unroll
      for (int j = 0; j < 4; ++j) o_acc[32 + j] = fmaf(o_acc[32 + j], corr, __uint_as_float(a2[j]));
    }
    tc::fence_before_sync();
    __syncthreads();  // K / V / P tiles and both accumulators are free for the next key tile
  }

  if (warp == 0) tc::tmem_dealloc(tmem, 256);
  if (q0 + tid < p.Lq) {
    float *dst = p.O + b * p.so_b + static_cast<long long>(q0 + tid) * p.ldo + h * hd;
    const float inv = 1.0f / l_run;  // l == 0 (every key masked) -> NaN like the reference softmax
#pragma unroll
    for (int d = 0; d < hd; d += 4) {
      float4 o4 = make_float4(o_acc[d] * inv, o_acc[d + 1] * inv, o_acc[d + 2] * inv, o_acc[d + 3] * inv);
      if (l_run == 0.f) o4 = make_float4(NAN, NAN, NAN, NAN);
      dst[d] = o4.x, dst[d + 1] = o4.y, dst[d + 2] = o4.z, dst[d + 3] = o4.w;
    }
  }
}

}  // namespace

extern "C" int bd_attention_tc(const float *Q, int ldq, long long sq_b, const float *K, int ldk, long long sk_b,
                               const float *V, int ldv, long long sv_b, const unsigned char *key_padding_mask,
                               float *O, int ldo, long long so_b, int B, int H, int Lq, int Lk, int hd, float scale,
                               int split, bd_stream_t stream) {
  BD_REQUIRE(Q && K && V && O, "bd_attention_tc: null pointer");
  BD_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0 && B <= 65535 && H <= 65535, "bd_attention_tc: bad sizes");
  BD_REQUIRE(hd == 36, "bd_attention_tc: built for head_dim 36 (d_model 288 / 8 heads)");
  BD_REQUIRE(split == 1 || split == 3, "bd_attention_tc: split must be 1 (bf16) or 3 (bf16x3)");
  BD_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && sq_b % 4 == 0 && sk_b % 4 == 0 &&
                 (reinterpret_cast<uintptr_t>(Q) & 15) == 0 && (reinterpret_cast<uintptr_t>(K) & 15) == 0,
             "bd_attention_tc: Q / K rows must be 16-byte aligned");
  AttnParams p = {};
  p.Q = Q, p.K = K, p.V = V, p.mask = key_padding_mask, p.O = O;
  p.ldq = ldq, p.ldk = ldk, p.ldv = ldv, p.ldo = ldo;
  p.sq_b = sq_b, p.sk_b = sk_b, p.sv_b = sv_b, p.so_b = so_b;
  p.Lq = Lq, p.Lk = Lk, p.hd = hd;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int parts = split == 3 ? 2 : 1;
  const size_t smem = static_cast<size_t>(parts) * (2 * QK_PART + V_PART + P_PART) + 1024;
  static thread_local bool configured = false;
  if (!configured) {
    BD_CUDA(cudaFuncSetAttribute(attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
            "bd_attention_tc");
    BD_CUDA(cudaFuncSetAttribute(attention_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
            "bd_attention_tc");
    configured = true;
  }
  dim3 grid(bd::ceil_div(Lq, AT_BM), H, B);
  if (parts == 2)
    attention_tc_kernel<2><<<grid, AT_THREADS, smem, bd::as_stream(stream)>>>(p);
  else
    attention_tc_kernel<1><<<grid, AT_THREADS, smem, bd::as_stream(stream)>>>(p);
  BD_CHECK_LAUNCH("bd_attention_tc");
  return BD_OK;
}
