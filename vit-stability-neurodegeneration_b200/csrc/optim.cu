// HBM-bound multi-tensor kernels for the optimiser side of the path:
//   SAM  (regularization/sam.py:38-155): per-tensor gradient norms, global norm, perturb (+save), restore
//   EMA  (utils/ema.py:72-108): on-device ring of the last 3 snapshots + weighted average
// One launch walks every parameter tensor through a chunk table (tensor id, element offset), so the
// reference's per-parameter Python loops and host syncs (isnan/isinf -> bool) disappear.
#include "common.cuh"

namespace {

constexpr int MT_CHUNK = 16384;   // elements per block
constexpr int MT_THREADS = 256;

struct MTTable {
  const long long* sizes;         // [n_tensors]
  const int* chunk_tensor;        // [n_chunks]
  const long long* chunk_off;     // [n_chunks]
};

__device__ __forceinline__ bool chunk_span(const MTTable& t, int& tensor, long long& off, long long& n) {
  tensor = t.chunk_tensor[blockIdx.x];
  off = t.chunk_off[blockIdx.x];
  const long long left = t.sizes[tensor] - off;
  n = left < MT_CHUNK ? left : MT_CHUNK;
  return n > 0;
}

// sq[t] += sum (w * g)^2 over the chunk, w = |p| if adaptive else 1
__global__ void __launch_bounds__(MT_THREADS) mt_sqnorm_kernel(const long long* g_ptrs, const long long* p_ptrs,
                                                               MTTable tab, float* sq, int adaptive) {
  int tensor; long long off, n;
  if (!chunk_span(tab, tensor, off, n)) return;
  const float* g = reinterpret_cast<const float*>(g_ptrs[tensor]) + off;
  const float* p = adaptive ? reinterpret_cast<const float*>(p_ptrs[tensor]) + off : nullptr;
  float acc = 0.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(g) | (p ? reinterpret_cast<uintptr_t>(p) : 0)) & 15) == 0;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = threadIdx.x; i < n4; i += MT_THREADS) {
      float4 v = reinterpret_cast<const float4*>(g)[i];
      if (adaptive) {
        const float4 w = reinterpret_cast<const float4*>(p)[i];
        v.x *= fabsf(w.x); v.y *= fabsf(w.y); v.z *= fabsf(w.z); v.w *= fabsf(w.w);
      }
      acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += MT_THREADS) {
      const float v = adaptive ? g[i] * fabsf(p[i]) : g[i];
      acc += v * v;
    }
  } else {
    for (long long i = threadIdx.x; i < n; i += MT_THREADS) {
      const float v = adaptive ? g[i] * fabsf(p[i]) : g[i];
      acc += v * v;
    }
  }
  __shared__ float red[MT_THREADS / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < MT_THREADS / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(sq + tensor, v);
  }
}

// Global SAM norm from the per-tensor sums of squares, with the reference's finite-ness rules:
// tensors whose norm is NaN/Inf are left out (and flagged so the perturbation skips them); an empty or
// non-finite total becomes 1e-12; a zero total disables the perturbation.
// out[0] = scale = rho / (norm + 1e-12) (0 when skipped), out[1] = norm.
__global__ void sam_scale_kernel(const float* sq, int n, float rho, float* out, int* skip_flags) {
  __shared__ double red[32];
  __shared__ int cnt[32];
  double acc = 0.0;
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float nrm = sqrtf(sq[i]);
    const bool ok = isfinite(nrm);
    skip_flags[i] = ok ? 0 : 1;
    if (ok) { acc += static_cast<double>(nrm) * nrm; ++c; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; cnt[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0; int k = 0;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) { tot += red[i]; k += cnt[i]; }
    float norm = k == 0 ? 1e-12f : static_cast<float>(sqrt(tot));
    if (!isfinite(norm)) norm = 1e-12f;
    out[1] = norm;
    out[0] = (norm == 0.f) ? 0.f : rho / (norm + 1e-12f);
  }
}

// The element-wise kernels below move 16 bytes per thread and access (float4; 8-byte bf16x4 stores in the cast), four
// independent accesses in flight per array: one chunk is 64 KB per array, a block streams it in 16 float4 per thread.
// Tensors (or tails) that are not 16-byte aligned / a multiple of 4 take the scalar loop.
__device__ __forceinline__ bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr,
                                          const void* d = nullptr, const void* e = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(e)) & 15) == 0;
}

// old = p;  p += (p^2 if adaptive else 1) * g * scale   (skipped per tensor flag / when scale == 0)
__global__ void __launch_bounds__(MT_THREADS) mt_sam_perturb_kernel(const long long* p_ptrs, const long long* g_ptrs,
                                                                    const long long* old_ptrs, MTTable tab,
                                                                    const float* scale_dev, const int* skip_flags,
                                                                    int adaptive) {
  int tensor; long long off, n;
  if (!chunk_span(tab, tensor, off, n)) return;
  float* p = reinterpret_cast<float*>(p_ptrs[tensor]) + off;
  const float* g = reinterpret_cast<const float*>(g_ptrs[tensor]) + off;
  float* old = reinterpret_cast<float*>(old_ptrs[tensor]) + off;
  const float scale = skip_flags[tensor] ? 0.f : scale_dev[0];
  long long done = 0;
  if (aligned16(p, g, old)) {
    const long long n4 = n >> 2;
#pragma unroll 4
    for (long long i = threadIdx.x; i < n4; i += MT_THREADS) {
      float4 w = reinterpret_cast<float4*>(p)[i];
      reinterpret_cast<float4*>(old)[i] = w;
      if (scale != 0.f) {
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        w.x = w.x + (adaptive ? w.x * w.x : 1.0f) * gv.x * scale;
        w.y = w.y + (adaptive ? w.y * w.y : 1.0f) * gv.y * scale;
        w.z = w.z + (adaptive ? w.z * w.z : 1.0f) * gv.z * scale;
        w.w = w.w + (adaptive ? w.w * w.w : 1.0f) * gv.w * scale;
        reinterpret_cast<float4*>(p)[i] = w;
      }
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < n; i += MT_THREADS) {
    const float w = p[i];
    old[i] = w;
    if (scale != 0.f) p[i] = w + (adaptive ? w * w : 1.0f) * g[i] * scale;
  }
}

__global__ void __launch_bounds__(MT_THREADS) mt_copy_kernel(const long long* dst_ptrs, const long long* src_ptrs,
                                                             MTTable tab) {
  int tensor; long long off, n;
  if (!chunk_span(tab, tensor, off, n)) return;
  float* d = reinterpret_cast<float*>(dst_ptrs[tensor]) + off;
  const float* s = reinterpret_cast<const float*>(src_ptrs[tensor]) + off;
  long long done = 0;
  if (aligned16(d, s)) {
    const long long n4 = n >> 2;
#pragma unroll 4
    for (long long i = threadIdx.x; i < n4; i += MT_THREADS)
      reinterpret_cast<float4*>(d)[i] = reinterpret_cast<const float4*>(s)[i];
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < n; i += MT_THREADS) d[i] = s[i];
}

// slot_new = p;  ema = w0*s0 + w1*s1 + w2*p accumulated oldest -> newest with fma (k = 1..3 live snapshots;
// the newest is always the current parameters).  Unused older slots are passed as null tables.
__global__ void __launch_bounds__(MT_THREADS) mt_ema_kernel(const long long* p_ptrs, const long long* new_ptrs,
                                                            const long long* s0_ptrs, const long long* s1_ptrs,
                                                            const long long* ema_ptrs, MTTable tab, float w0, float w1,
                                                            float w2) {
  int tensor; long long off, n;
  if (!chunk_span(tab, tensor, off, n)) return;
  const float* p = reinterpret_cast<const float*>(p_ptrs[tensor]) + off;
  float* snew = reinterpret_cast<float*>(new_ptrs[tensor]) + off;
  const float* s0 = s0_ptrs ? reinterpret_cast<const float*>(s0_ptrs[tensor]) + off : nullptr;
  const float* s1 = s1_ptrs ? reinterpret_cast<const float*>(s1_ptrs[tensor]) + off : nullptr;
  float* ema = reinterpret_cast<float*>(ema_ptrs[tensor]) + off;
  long long done = 0;
  if (aligned16(p, snew, s0, s1, ema)) {
    const long long n4 = n >> 2;
#pragma unroll 4
    for (long long i = threadIdx.x; i < n4; i += MT_THREADS) {
      const float4 v = reinterpret_cast<const float4*>(p)[i];
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s0) {
        const float4 a = reinterpret_cast<const float4*>(s0)[i];
        acc.x = fmaf(w0, a.x, acc.x); acc.y = fmaf(w0, a.y, acc.y); acc.z = fmaf(w0, a.z, acc.z); acc.w = fmaf(w0, a.w, acc.w);
      }
      if (s1) {
        const float4 a = reinterpret_cast<const float4*>(s1)[i];
        acc.x = fmaf(w1, a.x, acc.x); acc.y = fmaf(w1, a.y, acc.y); acc.z = fmaf(w1, a.z, acc.z); acc.w = fmaf(w1, a.w, acc.w);
      }
      acc.x = fmaf(w2, v.x, acc.x); acc.y = fmaf(w2, v.y, acc.y); acc.z = fmaf(w2, v.z, acc.z); acc.w = fmaf(w2, v.w, acc.w);
      reinterpret_cast<float4*>(snew)[i] = v;
      reinterpret_cast<float4*>(ema)[i] = acc;
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < n; i += MT_THREADS) {
    const float v = p[i];
    float acc = 0.f;
    if (s0) acc = fmaf(w0, s0[i], acc);
    if (s1) acc = fmaf(w1, s1[i], acc);
    acc = fmaf(w2, v, acc);
    snew[i] = v;
    ema[i] = acc;
  }
}

// AdamW over every parameter tensor in ONE pass (torch.optim.AdamW, train/train_transformer.py:2125-2147):
//   g' = g * inv_scale;  p *= 1 - lr*wd[t];  m += (g' - m)(1 - b1);  v = b2 v + (1 - b2) g'^2;
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps);   optionally g = 0 (the next pass accumulates into it).
// ctl = { bc1, sqrt(bc2), skip, inv_scale } comes from adamw_prepare_kernel: the step count, the GradScaler's
// found_inf / inv_scale live on the device, so a skipped step costs no host synchronisation.
// 28 B read+written per parameter (p, g, m, v in; p, m, v out) + 4 B for the fused gradient clear.
__global__ void __launch_bounds__(MT_THREADS) mt_adamw_kernel(const long long* p_ptrs, const long long* g_ptrs,
                                                              const long long* m_ptrs, const long long* v_ptrs,
                                                              MTTable tab, const float* __restrict__ wd,
                                                              const float* __restrict__ ctl, float lr, float b2, float omb1,
                                                              float omb2, float eps, int zero_grad) {
  int tensor; long long off, n;
  if (!chunk_span(tab, tensor, off, n)) return;
  const float bc1 = ctl[0], bc2s = ctl[1], inv_scale = ctl[3];
  float* g = reinterpret_cast<float*>(g_ptrs[tensor]) + off;
  if (ctl[2] != 0.f) {                                    // found_inf: the whole step is skipped; the fused clear still
    if (zero_grad)                                        // happens (the next pass accumulates into these buffers)
      for (long long i = threadIdx.x; i < n; i += MT_THREADS) g[i] = 0.f;
    return;
  }
  float* p = reinterpret_cast<float*>(p_ptrs[tensor]) + off;
  float* m = reinterpret_cast<float*>(m_ptrs[tensor]) + off;
  float* v = reinterpret_cast<float*>(v_ptrs[tensor]) + off;
  const float decay = 1.f - lr * wd[tensor], step_size = lr / bc1;   // omb1/omb2 = 1 - beta, rounded from double like torch's
  auto upd = [&](float& pw, float gw, float& mw, float& vw) {
    gw *= inv_scale;
    pw *= decay;
    mw = fmaf(gw - mw, omb1, mw);
    vw = fmaf(omb2 * gw, gw, b2 * vw);
    pw = pw - step_size * (mw / (sqrtf(vw) / bc2s + eps));
  };
  long long done = 0;
  if (aligned16(p, g, m, v)) {
    const long long n4 = n >> 2;
#pragma unroll 2
    for (long long i = threadIdx.x; i < n4; i += MT_THREADS) {
      float4 pw = reinterpret_cast<float4*>(p)[i], mw = reinterpret_cast<float4*>(m)[i], vw = reinterpret_cast<float4*>(v)[i];
      const float4 gw = reinterpret_cast<const float4*>(g)[i];
      upd(pw.x, gw.x, mw.x, vw.x); upd(pw.y, gw.y, mw.y, vw.y); upd(pw.z, gw.z, mw.z, vw.z); upd(pw.w, gw.w, mw.w, vw.w);
      reinterpret_cast<float4*>(p)[i] = pw;
      reinterpret_cast<float4*>(m)[i] = mw;
      reinterpret_cast<float4*>(v)[i] = vw;
      if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < n; i += MT_THREADS) {
    float pw = p[i], mw = m[i], vw = v[i];
    upd(pw, g[i], mw, vw);
    p[i] = pw; m[i] = mw; v[i] = vw;
    if (zero_grad) g[i] = 0.f;
  }
}

// step += 1 unless found_inf; ctl = { 1 - b1^step, sqrt(1 - b2^step), skip, inv_scale }
__global__ void adamw_prepare_kernel(float* step, const float* found_inf, const float* inv_scale, float b1, float b2,
                                     float* ctl) {
  const bool skip = found_inf != nullptr && found_inf[0] != 0.f;
  float s = step[0];
  if (!skip) { s += 1.f; step[0] = s; }
  ctl[0] = 1.f - powf(b1, s);
  ctl[1] = sqrtf(1.f - powf(b2, s));
  ctl[2] = skip ? 1.f : 0.f;
  ctl[3] = inv_scale != nullptr ? inv_scale[0] : 1.f;
}

// dst_bf16 = bf16(src_fp32) per tensor: the bf16 shadow of the GEMM weights
__global__ void __launch_bounds__(MT_THREADS) mt_cast_bf16_kernel(const long long* src_ptrs, const long long* dst_ptrs,
                                                                  MTTable tab) {
  int tensor; long long off, n;
  if (!chunk_span(tab, tensor, off, n)) return;
  const float* s = reinterpret_cast<const float*>(src_ptrs[tensor]) + off;
  bf16* d = reinterpret_cast<bf16*>(dst_ptrs[tensor]) + off;
  long long done = 0;
  if (aligned16(s) && (reinterpret_cast<uintptr_t>(d) & 7) == 0) {
    const long long n4 = n >> 2;
#pragma unroll 4
    for (long long i = threadIdx.x; i < n4; i += MT_THREADS) {
      const float4 v = reinterpret_cast<const float4*>(s)[i];
      reinterpret_cast<uint2*>(d)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < n; i += MT_THREADS) d[i] = f2b(s[i]);
}

}  // namespace

extern "C" int vsn_mt_cast_bf16(const long long* src_ptrs, const long long* dst_ptrs, const long long* sizes,
                                const int* chunk_tensor, const long long* chunk_off, int n_chunks, void* stream) {
  if (n_chunks == 0) return 0;
  MTTable t{sizes, chunk_tensor, chunk_off};
  mt_cast_bf16_kernel<<<n_chunks, MT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src_ptrs, dst_ptrs, t);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_mt_chunk_elems() { return MT_CHUNK; }

extern "C" int vsn_mt_sqnorm(const long long* g_ptrs, const long long* p_ptrs, const long long* sizes,
                             const int* chunk_tensor, const long long* chunk_off, int n_chunks, float* sq,
                             int adaptive, void* stream) {
  if (n_chunks == 0) return 0;
  VSN_CHECK(!adaptive || p_ptrs != nullptr, "vsn_mt_sqnorm: adaptive mode needs the parameter pointers");
  MTTable t{sizes, chunk_tensor, chunk_off};
  mt_sqnorm_kernel<<<n_chunks, MT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g_ptrs, p_ptrs, t, sq, adaptive);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_sam_scale(const float* sq, int n_tensors, float rho, float* out2, int* skip_flags, void* stream) {
  sam_scale_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sq, n_tensors, rho, out2, skip_flags);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_mt_sam_perturb(const long long* p_ptrs, const long long* g_ptrs, const long long* old_ptrs,
                                  const long long* sizes, const int* chunk_tensor, const long long* chunk_off,
                                  int n_chunks, const float* scale_dev, const int* skip_flags, int adaptive,
                                  void* stream) {
  if (n_chunks == 0) return 0;
  MTTable t{sizes, chunk_tensor, chunk_off};
  mt_sam_perturb_kernel<<<n_chunks, MT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p_ptrs, g_ptrs, old_ptrs, t,
                                                                                             scale_dev, skip_flags, adaptive);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_mt_copy(const long long* dst_ptrs, const long long* src_ptrs, const long long* sizes,
                           const int* chunk_tensor, const long long* chunk_off, int n_chunks, void* stream) {
  if (n_chunks == 0) return 0;
  MTTable t{sizes, chunk_tensor, chunk_off};
  mt_copy_kernel<<<n_chunks, MT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dst_ptrs, src_ptrs, t);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_mt_ema(const long long* p_ptrs, const long long* new_ptrs, const long long* s0_ptrs,
                          const long long* s1_ptrs, const long long* ema_ptrs, const long long* sizes,
                          const int* chunk_tensor, const long long* chunk_off, int n_chunks, float w0, float w1,
                          float w2, void* stream) {
  if (n_chunks == 0) return 0;
  MTTable t{sizes, chunk_tensor, chunk_off};
  mt_ema_kernel<<<n_chunks, MT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p_ptrs, new_ptrs, s0_ptrs, s1_ptrs,
                                                                                     ema_ptrs, t, w0, w1, w2);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_adamw_prepare(float* step, const float* found_inf, const float* inv_scale, float beta1, float beta2,
                                 float* ctl4, void* stream) {
  adamw_prepare_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(step, found_inf, inv_scale, beta1, beta2, ctl4);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_mt_adamw(const long long* p_ptrs, const long long* g_ptrs, const long long* m_ptrs,
                            const long long* v_ptrs, const long long* sizes, const int* chunk_tensor,
                            const long long* chunk_off, int n_chunks, const float* weight_decay, const float* ctl4,
                            float lr, float beta2, float one_minus_beta1, float one_minus_beta2, float eps,
                            int zero_grad, void* stream) {
  if (n_chunks == 0) return 0;
  MTTable t{sizes, chunk_tensor, chunk_off};
  mt_adamw_kernel<<<n_chunks, MT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p_ptrs, g_ptrs, m_ptrs, v_ptrs, t, weight_decay, ctl4, lr, beta2, one_minus_beta1, one_minus_beta2, eps, zero_grad);
  VSN_LAUNCH_CHECK();
  return 0;
}
