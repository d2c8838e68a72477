// Shared helpers for libbutd_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/butd_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libbutd_b200 is written for sm_100a (B200) only"
#endif

namespace bd {

void set_error(const char *fmt, ...);

inline cudaStream_t as_stream(bd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (cached per device).
int sm_count();

// One-time per-DEVICE set-up of a kernel (cudaFuncSetAttribute is a per-device property: a process
// that moves to a second GPU must opt that device in as well).  run(f) calls f() the first time it
// is reached with a given current device and remembers success.
struct PerDeviceOnce {
  std::mutex mu;
  unsigned long long done = 0ull;
  template <class F>
  cudaError_t run(F &&f) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    std::lock_guard<std::mutex> guard(mu);
    if (done & bit) return cudaSuccess;
    const cudaError_t e = f();
    if (e == cudaSuccess) done |= bit;
    return e;
  }
};

// (x, y, z, bits(index)) records of the cell list built by bd_grid_build, in cell order (B, n)
const float4 *grid_sorted_points(void *grid_workspace, int B, int n);

#define BD_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      bd::set_error(__VA_ARGS__);        \
      return BD_ERR_INVALID_ARG;         \
    }                                    \
  } while (0)

#define BD_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaPeekAtLastError();                                         \
    if (e__ != cudaSuccess) {                                                        \
      cudaGetLastError();                                                            \
      bd::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__));    \
      return BD_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

#define BD_CUDA(call, name)                                                          \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      bd::set_error("%s: %s", name, cudaGetErrorString(e__));                        \
      return BD_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

// ---- programmatic dependent launch (PDL).  The forward is a chain of ~320 short kernels; with PDL
// a kernel is scheduled as soon as its predecessor in the stream lets it (pdl_launch_dependents, or
// the predecessor's CTAs leaving their SMs) and runs its prologue — barrier init, TMEM allocation,
// the bulk copies of its WEIGHTS — under the predecessor's tail.  pdl_wait() then blocks until the
// predecessor has completed and its writes are visible; every read of activations and every
// global write of a PDL-launched kernel comes after it.  Only kernels containing pdl_wait() may be
// launched through launch_pdl().
extern int g_pdl;  // bd_set_pdl(): 1 = use PDL launches (default), 0 = plain stream order
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// Squared distance with the contraction pattern of the reference's compiled kernels
// (sm_100 SASS of ball_query_gpu.cu / sampling_gpu.cu / interpolate_gpu.cu:
//  FMUL dy,dy ; FFMA dx,dx ; FFMA dz,dz — for a*a + b*b the compiler multiplies the second
//  product and fuses the first).  Written with explicit intrinsics so the result
//  does not depend on this file's compile flags.
__device__ __forceinline__ float sqdist_ref(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// |p|^2 as compiled in the reference FPS kernel (sampling_gpu.cu:104).
__device__ __forceinline__ float sqnorm_ref(float x, float y, float z) {
  return __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
}

}  // namespace bd
