// fp32 multi-head attention core (flash-style, single pass over the keys, online softmax) —
// the exact-arithmetic path for nn.MultiheadAttention's softmax(QKᵀ/√hd + mask)·V
// (call sites /root/reference/models/encoder_decoder_layers.py:87,99,111,149,179,365,373,384,394).
// The reference materialises (B·8, Lq, Lk) score tensors and head-averaged weights it then
// discards; here scores never leave shared memory / registers.
//
// CTA = 64 queries x 1 head x 1 scene, 256 threads.  QKᵀ: 4x4 register tile per thread from
// transposed shared tiles (float4, conflict-free).  P·V: 4 rows x 4 cols per thread.
#include "common.cuh"

namespace {

constexpr int AT_BQ = 64, AT_BK = 64, AT_THREADS = 256, AT_LD = AT_BQ + 4;

template <int HD>
__global__ void __launch_bounds__(AT_THREADS)
attention_f32_kernel(const float *__restrict__ Q, int ldq, long long sq_b, const float *__restrict__ K, int ldk,
                     long long sk_b, const float *__restrict__ V, int ldv, long long sv_b,
                     const unsigned char *__restrict__ mask, float *__restrict__ O, int ldo, long long so_b, int Lq,
                     int Lk, float scale) {
  static_assert(HD % 4 == 0 && HD <= 64, "head_dim must be a multiple of 4, at most 64 (16 x HD / 4 <= 256 P.V threads)");
  extern __shared__ __align__(16) float at_dyn[];  // attn_f32_smem(HD) bytes
  float (*Qt)[AT_LD] = reinterpret_cast<float (*)[AT_LD]>(at_dyn);
  float (*Kt)[AT_LD] = Qt + HD;
  float (*Pt)[AT_LD] = Kt + HD;
  float (*Vs)[HD] = reinterpret_cast<float (*)[HD]>(Pt + AT_BK);
  __shared__ float corr_s[AT_BQ], l_s[AT_BQ];
  __shared__ unsigned char kvalid[AT_BK];

  const int tid = threadIdx.x;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_BQ;
  Q += b * sq_b + h * HD;
  K += b * sk_b + h * HD;
  V += b * sv_b + h * HD;
  O += b * so_b + h * HD;
  if (mask) mask += static_cast<long long>(b) * Lk;

  for (int e = tid; e < AT_BQ * HD; e += AT_THREADS) {
    const int r = e / HD, d = e % HD;
    Qt[d][r] = (q0 + r < Lq) ? Q[static_cast<long long>(q0 + r) * ldq + d] : 0.f;
  }

  const int tx = tid & 15, ty = tid >> 4;  // S tile: rows ty*4.., cols tx*4..
  constexpr int CG = HD / 4;
  const bool pv_active = tid < 16 * CG;
  const int rg = tid / CG, cg = tid % CG;  // PV tile: rows rg*4.., cols cg*4..
  float m_run[4], l_run[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY, l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  for (int k0 = 0; k0 < Lk; k0 += AT_BK) {
    __syncthreads();  // previous tile fully consumed (also orders the Qt fill on the first pass)
    for (int e = tid; e < AT_BK * HD; e += AT_THREADS) {
      const int r = e / HD, d = e % HD;
      const bool in = k0 + r < Lk;
      Kt[d][r] = in ? K[static_cast<long long>(k0 + r) * ldk + d] : 0.f;
      Vs[r][d] = in ? V[static_cast<long long>(k0 + r) * ldv + d] : 0.f;
    }
    if (tid < AT_BK) kvalid[tid] = (k0 + tid < Lk) && !(mask && mask[k0 + tid]);
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < HD; ++d) {
      const float4 qv = *reinterpret_cast<const float4 *>(&Qt[d][ty * 4]);
      const float4 kv = *reinterpret_cast<const float4 *>(&Kt[d][tx * 4]);
      const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
    }
    float p[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = kvalid[tx * 4 + j] ? s[i][j] * scale : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);
      const float corr = (m_new == -INFINITY) ? 1.f : __expf(m_run[i] - m_new);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[i][j] = (m_new == -INFINITY) ? 0.f : __expf(s[i][j] - m_new);
        sum += p[i][j];
      }
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, off);
      l_run[i] = l_run[i] * corr + sum;
      m_run[i] = m_new;
      if (tx == 0) corr_s[ty * 4 + i] = corr;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4 *>(&Pt[tx * 4 + j][ty * 4]) = make_float4(p[0][j], p[1][j], p[2][j], p[3][j]);
    __syncthreads();
    if (pv_active) {
      const float4 cv = *reinterpret_cast<const float4 *>(&corr_s[rg * 4]);
      const float ca[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] *= ca[i];
#pragma unroll 8
      for (int k = 0; k < AT_BK; ++k) {
        const float4 pv = *reinterpret_cast<const float4 *>(&Pt[k][rg * 4]);
        const float4 vv = *reinterpret_cast<const float4 *>(&Vs[k][cg * 4]);
        const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pa[i], va[j], o[i][j]);
      }
    }
  }
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) l_s[ty * 4 + i] = l_run[i];
  }
  __syncthreads();
  if (pv_active) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = q0 + rg * 4 + i;
      if (r < Lq) {
        const float inv = 1.0f / l_s[rg * 4 + i];  // l == 0 (all keys masked) -> NaN like the reference
        float4 v = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
        if (l_s[rg * 4 + i] == 0.f) v = make_float4(NAN, NAN, NAN, NAN);
        float *dst = O + static_cast<long long>(r) * ldo + cg * 4;
        dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;
      }
    }
  }
}

constexpr size_t attn_f32_smem(int hd) { return sizeof(float) * ((2 * hd + AT_BK) * AT_LD + AT_BK * hd); }

}  // namespace

extern "C" int bd_attention_f32(const float *Q, int ldq, long long sq_b, const float *K, int ldk, long long sk_b,
                                const float *V, int ldv, long long sv_b, const unsigned char *key_padding_mask,
                                float *O, int ldo, long long so_b, int B, int H, int Lq, int Lk, int hd, float scale,
                                bd_stream_t stream) {
  BD_REQUIRE(Q && K && V && O, "bd_attention_f32: null pointer");
  BD_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0 && B <= 65535 && H <= 65535, "bd_attention_f32: bad sizes");
  dim3 grid(bd::ceil_div(Lq, AT_BQ), H, B);
  cudaStream_t s = bd::as_stream(stream);
  if (hd == 36)
    attention_f32_kernel<36><<<grid, AT_THREADS, attn_f32_smem(36), s>>>(Q, ldq, sq_b, K, ldk, sk_b, V, ldv, sv_b,
                                                                         key_padding_mask, O, ldo, so_b, Lq, Lk, scale);
  else if (hd == 32)
    attention_f32_kernel<32><<<grid, AT_THREADS, attn_f32_smem(32), s>>>(Q, ldq, sq_b, K, ldk, sk_b, V, ldv, sv_b,
                                                                         key_padding_mask, O, ldo, so_b, Lq, Lk, scale);
  else if (hd == 64) {  // RoBERTa-base heads (the text encoder): 69 KB of shared memory
    static bd::PerDeviceOnce configured;
    BD_CUDA(configured.run([&]() {
      return cudaFuncSetAttribute(attention_f32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(attn_f32_smem(64)));
    }), "bd_attention_f32");
    attention_f32_kernel<64><<<grid, AT_THREADS, attn_f32_smem(64), s>>>(Q, ldq, sq_b, K, ldk, sk_b, V, ldv, sv_b,
                                                                         key_padding_mask, O, ldo, so_b, Lq, Lk, scale);
  } else {
    bd::set_error("bd_attention_f32: head_dim %d not built (36 = d_model 288 / 8 heads, 32, 64)", hd);
    return BD_ERR_UNSUPPORTED;
  }
  BD_CHECK_LAUNCH("bd_attention_f32");
  return BD_OK;
}
