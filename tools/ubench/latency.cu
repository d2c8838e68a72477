// Micro-latencies that bound the FPS round (B200): REDUX, shuffle butterfly, shared atomics, barriers,
// L2 / DRAM loads.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu && ./latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 256;

template <int MODE>
__global__ void chain(unsigned *out, long long *cyc, const unsigned *src, int stride) {
  __shared__ unsigned long long s_key;
  __shared__ unsigned s_w[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned v = threadIdx.x * 2654435761u + 12345u;
  unsigned long long idx = (threadIdx.x * 128 + blockIdx.x * 4096) % (1u << 20);
  if (threadIdx.x == 0) s_key = 0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    if (MODE == 0) {  // one REDUX per iteration, dependent
      v = __reduce_max_sync(0xFFFFFFFFu, v) + lane + i;
    } else if (MODE == 1) {  // REDUX pair (hi then lo of matching lanes)
      const unsigned hi = __reduce_max_sync(0xFFFFFFFFu, v);
      const unsigned lo = __reduce_max_sync(0xFFFFFFFFu, v == hi ? (v ^ 0x5bd1e995u) : 0u);
      v = hi + lo + lane + i;
    } else if (MODE == 2) {  // 5-step shuffle butterfly (max)
      unsigned m = v;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
      v = m + lane + i;
    } else if (MODE == 3) {  // REDUX + ballot + shfl (no-tie fast path)
      const unsigned hi = __reduce_max_sync(0xFFFFFFFFu, v);
      const unsigned b = __ballot_sync(0xFFFFFFFFu, v == hi);
      const unsigned lo = __shfl_sync(0xFFFFFFFFu, v ^ 0x5bd1e995u, __ffs(b) - 1);
      v = hi + lo + lane + i;
    } else if (MODE == 4) {  // shared 64-bit atomicMax by lane 0 + __syncthreads + broadcast read
      if (lane == 0) atomicMax(&s_key, (static_cast<unsigned long long>(v) << 32) | i);
      __syncthreads();
      v = static_cast<unsigned>(s_key >> 32) + lane + i;
    } else if (MODE == 5) {  // STS + __syncthreads + LDS + REDUX pair (the CTA stage of the old kernel)
      if (lane == 0) s_w[warp] = v;
      __syncthreads();
      const unsigned w = lane < (blockDim.x >> 5) ? s_w[lane] : 0u;
      const unsigned hi = __reduce_max_sync(0xFFFFFFFFu, w);
      const unsigned lo = __reduce_max_sync(0xFFFFFFFFu, w == hi ? w ^ 3u : 0u);
      v = hi + lo + lane + i;
      __syncthreads();
    } else if (MODE == 6) {  // dependent global load (pointer chase over `stride` bytes), one lane
      idx = src[(idx * stride / 4) % (1u << 24)] % (1u << 20);
      v += idx;
    } else if (MODE == 7) {  // coalesced 16-byte warp load, address depends on the previous value
      const uint4 q = reinterpret_cast<const uint4 *>(src)[((v & 0xFFFFu) * 32 + lane) % (1u << 22)];
      v = __shfl_sync(0xFFFFFFFFu, q.x ^ q.w, 0) + i;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / ITERS;
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

template <int MODE>
void run(const char *name, int blocks, int threads, const unsigned *src, int stride = 0) {
  unsigned *out;
  long long *cyc;
  cudaMalloc(&out, sizeof(unsigned) * blocks * threads);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  chain<MODE><<<blocks, threads>>>(out, cyc, src, stride);
  chain<MODE><<<blocks, threads>>>(out, cyc, src, stride);
  cudaDeviceSynchronize();
  long long h[4] = {0, 0, 0, 0};
  cudaMemcpy(h, cyc, sizeof(long long) * (blocks < 4 ? blocks : 4), cudaMemcpyDeviceToHost);
  printf("%-64s blocks=%3d threads=%4d : %5lld cycles/iter\n", name, blocks, threads, h[0]);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  unsigned *src;
  const size_t n = 1u << 26;  // 256 MB of indices: larger than L2
  cudaMalloc(&src, n * 4);
  unsigned *h = new unsigned[n];
  unsigned x = 1;
  for (size_t i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; h[i] = x >> 4; }
  cudaMemcpy(src, h, n * 4, cudaMemcpyHostToDevice);
  for (int threads : {32, 512, 1024}) {
    run<0>("REDUX.MAX dependent", 1, threads, src);
    run<1>("REDUX pair (hi, lo of ties)", 1, threads, src);
    run<2>("shuffle butterfly max (5 steps)", 1, threads, src);
    run<3>("REDUX + ballot + shfl", 1, threads, src);
    run<4>("ATOMS.MAX.64 by lane 0 + bar.sync + LDS", 1, threads, src);
    run<5>("STS + bar.sync + LDS + REDUX pair + bar.sync", 1, threads, src);
  }
  run<1>("REDUX pair, 148 CTAs", 148, 512, src);
  run<6>("dependent LDG, 4 MB window (L2 hit)", 1, 32, src, 64);
  run<7>("coalesced 512 B warp load, 64 MB window", 1, 32, src);
  run<7>("coalesced 512 B warp load, 64 MB window, 148 x 512 thr", 148, 512, src);
  return 0;
}
