// Shared device/host helpers for the vsn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

// The 16-bit activation / operand type of the whole path.  Default build: bfloat16 (libvsn_b200.so, 2e-2 tolerance).
// -DVSN_F16 builds the SAME kernels with IEEE half operands (libvsn_b200_f16.so): 11 significant bits, the
// mantissa of TF32 -- the precision mode north_star's 1e-3 tolerance asks for, and the type the reference itself trains
// in (torch.autocast(float16) + GradScaler, train/train_transformer.py:1141-1160).  Every kernel keeps fp32
// accumulation and fp32 statistics in both builds; only the 16-bit encoding differs, so layouts, strides, TMA maps and
// shared-memory plans are identical.  The name `bf16` is kept for the type in both builds.
#ifdef VSN_F16
typedef __half bf16;
typedef __half2 bf162;
#define VSN_ONE_PAIR 0x3C003C00u   // 1.0 | 1.0
#define VSN_T16 "f16"             // PTX type name of the pair (mma.sync, red.global.add)
#define VSN_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#else
typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;
#define VSN_ONE_PAIR 0x3F803F80u
#define VSN_T16 "bf16"
#define VSN_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#endif

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
bool vsn_pdl_enabled();

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
#ifdef VSN_F16
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf162 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  bf162 t = *reinterpret_cast<bf162*>(&u);
  return __half22float2(t);
}
__host__ __device__ __forceinline__ bf16 f2b(float v) { return __float2half_rn(v); }
__host__ __device__ __forceinline__ float b2f(bf16 v) { return __half2float(v); }
#else
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  bf162 t = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(t);
}
__host__ __device__ __forceinline__ bf16 f2b(float v) { return __float2bfloat16(v); }
__host__ __device__ __forceinline__ float b2f(bf16 v) { return __bfloat162float(v); }
#endif
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
  return exp2_neg<POLY>(fmaf(-z, q, -1.0f));               // the factor 0.5 rides in the exponent
}
// gelu(x) = max(x, 0) - |x| * (1 - Phi(|x|)): no select, no cancellation in either tail
template <bool POLY = false>
__device__ __forceinline__ float gelu_erf(float x) {
  return fmaf(-fabsf(x), erfc_half_pos<POLY>(fabsf(x)), fmaxf(x, 0.f));
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
#ifdef VSN_F16
__device__ __forceinline__ float bf16lo(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u & 0xFFFFu))); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u >> 16))); }
#else
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
#endif


// ---- packed (half2) GELU for GEMM epilogues --------------------------------------------------------------------
// The epilogue of the MLP GEMMs is bound by instruction issue, and its output is rounded to bf16 (2^-9) anyway:
// the erf polynomial, the exponential (ex2.approx.f16x2: two per MUFU) and the final products run two elements
// per instruction in fp16 (2^-11).  Inputs are bf16-representable values, exact in fp16 for |x| < 65504.
__device__ __forceinline__ uint32_t h2_ex2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ __half2 h2_from_u32(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2_to_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
// 0.5 * erfc(|x| / sqrt(2)) = 1 - Phi(|x|) for two values (no cancellation in the tails)
__device__ __forceinline__ __half2 h2_half_erfc(__half2 ax) {
  const __half2 z = __hmin2(__hmul2(ax, __float2half2_rn(0.70710678f)), __float2half2_rn(4.0f));
  __half2 q = __hfma2(z, __float2half2_rn(0.000233418324f), __float2half2_rn(-0.00402740239f));   // -q(z)
  q = __hfma2(q, z, __float2half2_rn(0.031229802f));
  q = __hfma2(q, z, __float2half2_rn(-0.149565667f));
  q = __hfma2(q, z, __float2half2_rn(-0.918361976f));
  q = __hfma2(q, z, __float2half2_rn(-1.62790073f));
  return h2_from_u32(h2_ex2(h2_to_u32(__hfma2(z, q, __float2half2_rn(-1.0f)))));                   // 2^(-z q - 1)
}
// exponent argument -(z q(z)) - 1 of 0.5 * erfc(|x| / sqrt(2)) for two values (fp16 polynomial)
__device__ __forceinline__ __half2 h2_half_erfc_arg(__half2 ax) {
  const __half2 z = __hmin2(__hmul2(ax, __float2half2_rn(0.70710678f)), __float2half2_rn(4.0f));
  __half2 q = __hfma2(z, __float2half2_rn(0.000233418324f), __float2half2_rn(-0.00402740239f));   // -q(z)
  q = __hfma2(q, z, __float2half2_rn(0.031229802f));
  q = __hfma2(q, z, __float2half2_rn(-0.149565667f));
  q = __hfma2(q, z, __float2half2_rn(-0.918361976f));
  q = __hfma2(q, z, __float2half2_rn(-1.62790073f));
  return __hfma2(z, q, __float2half2_rn(-1.0f));
}
// GELU of the two bf16 values packed in `pk` (low = element 0), results as fp32:
// gelu(x) = max(x, 0) - |x| * (1 - Phi(|x|)).  The polynomial runs packed in fp16; the exponential and the final
// combination stay in fp32 (ex2.approx.f16x2 is only 2^-9.9 accurate and its error does not average out over the
// token sums of the weight gradients).
__device__ __forceinline__ float2 gelu_erf_bf16x2(uint32_t pk) {
  const float x0 = bf16lo(pk), x1 = bf16hi(pk);
  const __half2 x = __floats2half2_rn(x0, x1);
  const float2 arg = __half22float2(h2_half_erfc_arg(h2_from_u32(h2_to_u32(x) & 0x7FFF7FFFu)));
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(arg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(arg.y));
  return make_float2(fmaf(-fabsf(x0), e0, fmaxf(x0, 0.f)), fmaf(-fabsf(x1), e1, fmaxf(x1, 0.f)));
}
// GELU'(x) = Phi(x) + x * phi(x) of the two bf16 values packed in `pk`, results as fp32.
// Phi(x) = x < 0 ? (1 - Phi(|x|)) : Phi(|x|), selected per half with a sign mask (no cancellation for x < 0).
__device__ __forceinline__ float2 gelu_erf_grad_bf16x2(uint32_t pk) {
  const __half2 x = __floats2half2_rn(bf16lo(pk), bf16hi(pk));
  const uint32_t xu = h2_to_u32(x);
  const __half2 ax = h2_from_u32(xu & 0x7FFF7FFFu);
  const __half2 he = h2_half_erfc(ax);
  const uint32_t m = ((xu >> 15) & 0x00010001u) * 0xFFFFu;                         // 0xFFFF where x < 0
  const uint32_t hi = h2_to_u32(__hsub2(__float2half2_rn(1.0f), he));
  const __half2 cdf = h2_from_u32((h2_to_u32(he) & m) | (hi & ~m));
  const __half2 axc = __hmin2(ax, __float2half2_rn(8.0f));                          // phi(8) ~ 5e-15
  const __half2 g = h2_from_u32(h2_ex2(h2_to_u32(__hmul2(__hmul2(axc, axc), __float2half2_rn(-0.72134752f)))));
  return __half22float2(__hfma2(__hmul2(x, __float2half2_rn(0.39894228f)), g, cdf));
}

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

// ---- programmatic dependent launch (sm_90+) -------------------------------------------------------------------
// pdl_trigger(): this grid no longer needs to finish before the next kernel of the stream may be SCHEDULED (its CTAs
// take SM resources as ours retire and run their prologue); pdl_wait(): block until every kernel the stream ordered
// before this one has completed and its memory is visible.  A kernel launched with the attribute must call
// pdl_wait() before it touches global memory; one launched without it is unaffected by either call.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
