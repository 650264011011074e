// Common helpers for the molclr_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

#ifndef __CUDA_ARCH__
#define MOLCLR_HOST 1
#endif

namespace molclr {

// ---------------------------------------------------------------- error reporting (host)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

extern unsigned long long g_launches;   // kernels launched by this library (diagnostics: molclr_launch_count)

#define MOLCLR_CHECK_LAUNCH(what)                                  \
  do {                                                             \
    ++::molclr::g_launches;                                        \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return ::molclr::cuda_fail(_e, what);   \
  } while (0)

#define MOLCLR_REQUIRE(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) { ::molclr::set_error(__VA_ARGS__); return -2; }  \
  } while (0)

int sm_count();   // cached multiprocessor count of the current device

// ---------------------------------------------------------------- kernel launches (programmatic dependent launch)
// A training step is ~230 dependent launches of 3 - 170 us each on one stream.  Every kernel of this library is launched with
// programmatic stream serialization: its launch (grid set-up, CTA scheduling, the prologue before pdl_sync()) overlaps the tail of
// the kernel before it, and pdl_sync() -- the first thing every kernel does before it touches global memory -- waits until that
// kernel has COMPLETED and its writes are visible.  Results are exactly those of plain stream order; molclr_set_pdl(0) switches
// the attribute off (the device-side instructions are then no-ops).
extern int g_pdl;

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);     // (errors: cudaGetLastError in MOLCLR_CHECK_LAUNCH)
}
#define MOLCLR_LAUNCH(kernel, grid, block, smem, stream, ...) ::molclr::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)

// Tuning / debugging switches (MOLCLR_GEMM_*, MOLCLR_AGG_*, MOLCLR_NTX_*) are read from the environment ONLY in a library built
// with -DMOLCLR_DEBUG_SWITCHES (python -m molclr_b200.build --debug-switches -> libmolclr_b200_dbg.so, used by tools/);
// the product library never looks at the environment, so a stray variable cannot change its results or its speed.
inline const char* debug_env(const char* name) {
#ifdef MOLCLR_DEBUG_SWITCHES
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

// ---------------------------------------------------------------- device helpers
// Executed by every thread at the top of every kernel, before its first global-memory access: wait for the preceding kernels of the
// stream (complete, writes visible), then allow the NEXT kernel's launch to begin (its CTAs become resident as this kernel's CTAs
// exit and wait at their own pdl_sync()).  No-ops in a kernel launched without the programmatic attribute.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ float round_tf32(float x) {
  // round-to-nearest (ties away) to 10 mantissa bits; low 13 bits of the result are zero, so the
  // tensor core's operand truncation is exact on it.
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming (read-once) 128-bit load: do not pollute L1
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// L2 prefetch of the 128-byte line holding p (no register cost, no L1 allocation)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// Packed fp32 pairs (FADD2 / FFMA2 on sm_100): the same IEEE round-to-nearest results as four scalar instructions in half
// the issue slots -- for kernels that are bound by instruction issue rather than by the FP32 pipe.
__device__ __forceinline__ float4 f4_add2(float4 a, float4 b) {
  float4 r;
  asm("{\n\t.reg .b64 a0, a1, b0, b1, r0, r1;\n\t"
      "mov.b64 a0, {%4, %5}; mov.b64 a1, {%6, %7}; mov.b64 b0, {%8, %9}; mov.b64 b1, {%10, %11};\n\t"
      "add.rn.f32x2 r0, a0, b0; add.rn.f32x2 r1, a1, b1;\n\t"
      "mov.b64 {%0, %1}, r0; mov.b64 {%2, %3}, r1;\n\t}"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
      : "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w));
  return r;
}
__device__ __forceinline__ float4 f4_fma2(float4 a, float4 b, float4 c) {
  float4 r;
  asm("{\n\t.reg .b64 a0, a1, b0, b1, c0, c1, r0, r1;\n\t"
      "mov.b64 a0, {%4, %5}; mov.b64 a1, {%6, %7}; mov.b64 b0, {%8, %9}; mov.b64 b1, {%10, %11}; mov.b64 c0, {%12, %13}; mov.b64 c1, {%14, %15};\n\t"
      "fma.rn.f32x2 r0, a0, b0, c0; fma.rn.f32x2 r1, a1, b1, c1;\n\t"
      "mov.b64 {%0, %1}, r0; mov.b64 {%2, %3}, r1;\n\t}"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
      : "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w), "f"(c.x), "f"(c.y), "f"(c.z), "f"(c.w));
  return r;
}

__device__ __forceinline__ float4 f4_tf32(float4 v) {
  return make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
}
// tf32( v - tf32(v) ): the "lo" half of the 2-term TF32 split used by the error-compensated GEMM
__device__ __forceinline__ float4 f4_tf32_residual(float4 v) {
  return make_float4(round_tf32(v.x - round_tf32(v.x)), round_tf32(v.y - round_tf32(v.y)), round_tf32(v.z - round_tf32(v.z)),
                     round_tf32(v.w - round_tf32(v.w)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kNumAtomType = 119;   // ginet_molclr.py:9
constexpr int kNumChirality = 3;    // ginet_molclr.py:10
constexpr int kNumBondType = 5;     // ginet_molclr.py:12 (4 = self loop)
constexpr int kNumBondDir = 3;      // ginet_molclr.py:13
constexpr int kSelfLoopAttr = 4 * 3 + 0;   // packed (type 4, dir 0)
constexpr int kNumEdgeClass = 15;

}  // namespace molclr
