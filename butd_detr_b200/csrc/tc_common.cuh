// tcgen05 / TMEM / mbarrier / bulk-copy primitives (inline PTX) shared by the tensor-core kernels.
// sm_100a only.  Layout conventions used by every tensor-core kernel in this library:
//
//   * operands are K-MAJOR bf16, staged in shared memory in the canonical 128-BYTE-SWIZZLE layout
//     (UMMA LayoutType::SWIZZLE_128B, the layout TMA's CU_TENSOR_MAP_SWIZZLE_128B produces): a tile
//     of R rows is cut along K into blocks of 64 elements (128 bytes per row); inside a block
//         byte(row r, k) = (r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2
//     i.e. 8-row groups of 1024 bytes whose 16-byte chunks are XOR-swizzled with the row number —
//     conflict-free both for 16-byte st.shared from threads (gathered rows, softmax probabilities,
//     MLP activations) and for the tensor core's operand fetch.  Block b of a tile starts at
//     b * R * 128 bytes; every block base is 1024-byte aligned.  (A first version used the
//     un-swizzled core-matrix layout: correct, but each MMA then took ~4x its nominal cycles.)
//   * weights are pre-packed by the host in exactly this layout, so a block arrives with one
//     cp.async.bulk (TMA engine) signalled on an mbarrier.
//   * accumulators live in TMEM: D[row i][col j] = lane i, column base + j (cta_group::1, M=128).
#pragma once
#include <cuda.h>  // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no libcuda link)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace tc {

// cuTensorMapEncodeTiled through the runtime's driver entry point (host)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

constexpr int KB = 64;  // K elements per swizzle block (128 bytes of bf16)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// one elected lane of a fully converged warp (the CUTLASS pattern that lets the compiler keep
// MMA descriptors in uniform registers instead of emitting per-lane R2UR loops)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// ---- bulk async copy (TMA engine, 1-D): global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// generic-proxy shared-memory writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__host__ __device__ inline uint32_t tmem_cols_pow2(uint32_t n) {
  uint32_t c = 32;
  while (c < n) c <<= 1;
  return c;
}

// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// wider / narrower TMEM loads of the same 32x32b shape, and the matching stores (registers -> TMEM)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t &r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};" ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};" ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors
// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major, canonical value 1),
// [32,46) SBO>>4 = 1024 B between 8-row groups, [46,48) version = 1, [61,64) layout type = 2.
// Advancing along K inside a 64-element block = adding the byte offset (32 B per 16 elements)
// to the start address; the hardware applies the XOR swizzle on the final address bits.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// shared-memory matrix descriptor, MN-MAJOR, SWIZZLE_128B: the operand lies in shared memory as rows of
// 128 bytes = 64 consecutive elements along M / N for ONE k, 8 consecutive k per 1024-byte swizzle atom
// (exactly what a 128-byte-swizzle tensor copy of a row-major [k][n] matrix produces); SBO = 1024 B between
// groups of 8 k, LBO (between 64-element groups along M / N) unused for N <= 64.  A K = 16 step advances the
// start address by 2048 B.  The instruction descriptor must carry the matching major bit (idesc_b_mn).
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t smem_addr) { return smem_desc_sw128(smem_addr); }
constexpr uint32_t idesc_b_mn = 1u << 16;  // InstrDescriptor b_major = MN
// 2-D tensor copy (TMA) global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
// instruction descriptor, kind::f16, A/B = bf16 K-major, D = f32 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : A operand read from tensor memory (lane = row, one 32-bit
// column = two consecutive bf16 along K; a K=16 step is 8 columns); issued by ONE thread
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all previously issued MMAs of this thread -> one arrival on an mbarrier
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// The single-pass mode (split = 1) uses FP16 operands: 11 significant bits against bf16's 8 — plain
// bf16 operands miss the 1e-2 output gate after ~30 stacked layers (1.2-2.2e-2), fp16 operands meet
// it; every operand of this model (post-LayerNorm activations, weights, probabilities) is far
// inside fp16's range, and accumulation is fp32 in TMEM either way.  The two-part mode (split = 3)
// stays bf16 hi + lo.
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_ab(int parts, uint32_t M, uint32_t N) {
  return parts == 2 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24)) : idesc_f16(M, N);
}

// byte offset of (row r, 16-byte chunk c in 0..7) inside one 64-wide swizzle block
__device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t c) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((c ^ (r & 7u)) << 4);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}

// 8 fp32 -> 8 bf16 (hi) and, optionally, the 8 bf16 residuals (lo) of the bf16x3 split
__device__ __forceinline__ void split_bf16x8(const float (&v)[8], uint4 &hi, uint4 &lo) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = v[i] - __bfloat162float(__float2bfloat16_rn(v[i]));
  hi.x = pack_bf16x2(v[0], v[1]), hi.y = pack_bf16x2(v[2], v[3]);
  hi.z = pack_bf16x2(v[4], v[5]), hi.w = pack_bf16x2(v[6], v[7]);
  lo.x = pack_bf16x2(r[0], r[1]), lo.y = pack_bf16x2(r[2], r[3]);
  lo.z = pack_bf16x2(r[4], r[5]), lo.w = pack_bf16x2(r[6], r[7]);
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  // .satfinite: magnitudes beyond fp16's 65504 saturate instead of becoming inf (an inf operand would
  // turn the whole accumulator row into inf / NaN); NaN stays NaN.  Same instruction count (F2FP.SATFINITE).
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// same with max(., 0) applied by the conversion (F2FP.RELU): ReLU costs no instruction of its own
__device__ __forceinline__ uint32_t pack_f16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// warp-wide maximum of an fp32 value (CREDUX.MAX.F32, sm_100a): whole warp / the caller's 16-lane half
__device__ __forceinline__ float redux_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float redux_max_f32_half(float v, int lane) {
  float r;
  if (lane < 16) asm volatile("redux.sync.max.f32 %0, %1, 0x0000ffff;" : "=f"(r) : "f"(v));
  else asm volatile("redux.sync.max.f32 %0, %1, 0xffff0000;" : "=f"(r) : "f"(v));
  return r;
}
// 8 fp32 -> operand chunk(s): parts == 2: bf16 hi + bf16 lo; parts == 1: fp16 (lo untouched)
__device__ __forceinline__ void cvt8(int parts, const float (&v)[8], uint4 &hi, uint4 &lo) {
  if (parts == 2) {
    split_bf16x8(v, hi, lo);
  } else {
    hi.x = pack_f16x2(v[0], v[1]), hi.y = pack_f16x2(v[2], v[3]);
    hi.z = pack_f16x2(v[4], v[5]), hi.w = pack_f16x2(v[6], v[7]);
  }
}

// relu(v) -> operand chunk(s)
__device__ __forceinline__ void cvt8_relu(int parts, const float (&v)[8], uint4 &hi, uint4 &lo) {
  if (parts == 2) {
    const float r[8] = {fmaxf(v[0], 0.f), fmaxf(v[1], 0.f), fmaxf(v[2], 0.f), fmaxf(v[3], 0.f),
                        fmaxf(v[4], 0.f), fmaxf(v[5], 0.f), fmaxf(v[6], 0.f), fmaxf(v[7], 0.f)};
    split_bf16x8(r, hi, lo);
  } else {
    hi.x = pack_f16x2_relu(v[0], v[1]), hi.y = pack_f16x2_relu(v[2], v[3]);
    hi.z = pack_f16x2_relu(v[4], v[5]), hi.w = pack_f16x2_relu(v[6], v[7]);
  }
}

// ---- packed fp32x2 arithmetic (FADD2) and the raw MUFU exponential used by the softmax warps
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned long long pack_f32x2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void add_f32x2(float &a, float &b, unsigned long long y) {  // (a, b) += y
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack_f32x2(a, b)), "l"(y));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}
// (p0, p1) -> packed bf16 pair hi (round to nearest) and the packed bf16 pair of the residuals
__device__ __forceinline__ void split_bf16x2(float p0, float p1, uint32_t &hi, uint32_t &lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(p1), "f"(p0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack_f32x2(p0, p1)), "l"(pack_f32x2(h0, h1)));
  float l0, l1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(l0), "=f"(l1) : "l"(d));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l1), "f"(l0));
}

}  // namespace tc
