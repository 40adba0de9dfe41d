// Shared device/host helpers for the vsn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ---- error plumbing (C-ABI: int return codes + thread-local message) ----------
void vsn_set_error(const char* fmt, ...);
#define VSN_CHECK(cond, ...)                 \
  do {                                       \
    if (!(cond)) {                           \
      vsn_set_error(__VA_ARGS__);            \
      return 1;                              \
    }                                        \
  } while (0)
#define VSN_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      vsn_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)
void vsn_count_launch();
#define VSN_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    vsn_count_launch();                                                                  \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      vsn_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 3;                                                                          \
    }                                                                                    \
  } while (0)

int vsn_num_sms();

// ---- small device utilities ----------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  bf162 t = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(t);
}
// Exact-erf GELU (nn.GELU default, models/swin_transformer_3d.py:60, models/vit_3d.py:72) and its derivative.
// erf(z) = 1 - 2^(-z Q(z)) on z in [0,4] with a degree-5 fit of Q (max abs error 3.5e-6 on erf, 2.1e-6 on GELU;
// clamped at z = 4 where 1 - erf < 2e-8): one MUFU and ~12 FP32 instructions instead of libdevice erff's ~30.
// 2^x for x <= 0 on the FMA / ALU pipes (Cody-Waite split + degree-4 polynomial, relative error < 5e-5):
// MUFU.EX2 issues at 16 clk per warp on B200, so kernels that need one exponential per element alternate.
__device__ __forceinline__ float exp2_neg_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;                 // 1.5 * 2^23: integer part lands in the low mantissa bits
  const float r = x - (t - 12582912.f);           // r in [-0.5, 0.5]
  float p = fmaf(r, 0.0096181291f, 0.0555041087f);
  p = fmaf(p, r, 0.2402265070f);
  p = fmaf(p, r, 0.6931471806f);
  p = fmaf(p, r, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
template <bool POLY>
__device__ __forceinline__ float exp2_neg(float x) {
  if constexpr (POLY) {
    return exp2_neg_poly(x);
  } else {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
}
template <bool POLY>
__device__ __forceinline__ float erfc_half_pos(float ax) {   // 0.5 * erfc(ax / sqrt(2)) for ax >= 0
  const float z = fminf(ax * 0.70710678118654752440f, 4.0f);
  float q = fmaf(z, -0.000233418324f, 0.00402740239f);
  q = fmaf(q, z, -0.031229802f);
  q = fmaf(q, z, 0.149565667f);
  q = fmaf(q, z, 0.918361976f);
  q = fmaf(q, z, 1.62790073f);
  return 0.5f * exp2_neg<POLY>(-z * q);
}
template <bool POLY = false>
__device__ __forceinline__ float gelu_erf(float x) {
  const float t = x * erfc_half_pos<POLY>(fabsf(x));     // x * (1 - Phi(|x|))
  return x < 0.f ? t : x - t;
}
// The cdf term uses the MUFU or the polynomial exponential (POLY); the pdf term always uses the other one,
// so every element costs exactly one MUFU.
template <bool POLY = false>
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float h = erfc_half_pos<POLY>(fabsf(x));
  const float cdf = x < 0.f ? h : 1.0f - h;
  const float g = exp2_neg<!POLY>(x * x * -0.72134752044448170368f);   // exp(-x^2/2)
  return fmaf(x * 0.39894228040143267794f, g, cdf);
}
// bf16 values of a packed pair as fp32 (exact)
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a thread that owns a contiguous run of a row reads or
// writes whole 32-byte sectors with one instruction.  The address must be 32-byte aligned.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t* v) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
