// Data-movement kernels around the GEMMs: patch gather (im2col of non-overlapping patches),
// stage pad/crop, patch-merging gather/scatter, row casts, token mean and the tiny classifier head.
// All are index math + coalesced 16-byte accesses; none of them re-reads data.
#include "common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(bf16* p, float v) { *p = f2b(v); }

// out[(b,d,h,w), (i,j,k)] = vol[b, d*pd+i, h*ph+j, w*pw+k]  (zero beyond D,H,W).
// PatchEmbed3D's pad + Conv3d(k=s=patch) operand (models/swin_transformer_3d.py:532-539) and the ViT
// Rearrange 'b c (d p1)(h p2)(w p3) -> b (d h w)(p1 p2 p3 c)' with c=1 (models/vit_3d.py:365-370).
// One thread per (token, i, j): it copies the pw contiguous voxels of one patch row.
template <typename InT, typename OutT>
__global__ void patch_gather_kernel(const InT* __restrict__ vol, OutT* __restrict__ out, int B, int D, int H, int W,
                                    int gd, int gh, int gw, int pd, int ph, int pw) {
  const long long total = static_cast<long long>(B) * gd * gh * gw * pd * ph;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    // order (b, d, h, i, j, w): consecutive threads walk w, so global reads of one (i,j) row are contiguous
    long long t = idx;
    const int w = static_cast<int>(t % gw); t /= gw;
    const int j = static_cast<int>(t % ph); t /= ph;
    const int i = static_cast<int>(t % pd); t /= pd;
    const int h = static_cast<int>(t % gh); t /= gh;
    const int d = static_cast<int>(t % gd); t /= gd;
    const int b = static_cast<int>(t);
    const int zd = d * pd + i, zh = h * ph + j;
    const long long token = ((static_cast<long long>(b) * gd + d) * gh + h) * gw + w;
    OutT* o = out + token * (pd * ph * pw) + (i * ph + j) * pw;
    const bool ok = zd < D && zh < H;
    const InT* src = vol + ((static_cast<long long>(b) * D + zd) * H + zh) * W + static_cast<long long>(w) * pw;
    for (int k = 0; k < pw; ++k) {
      const float v = (ok && w * pw + k < W) ? ld_as_float<InT>(src + k) : 0.f;
      st_from_float(o + k, v);
    }
  }
}


// Fast path for 4x4x4 patches (the Swin patch embedding): one thread per (token, i) reads the four 4-voxel rows
// j = 0..3 (consecutive threads walk w, so each row load of a warp is contiguous) and writes the 16 outputs
// as one full 32-byte sector.  W % 4 == 0 and a 4-voxel aligned volume row are required.
template <typename InT>
__device__ __forceinline__ void load4(const InT* p, float* v);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* v) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__half>(const __half* p, float* v) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xFFFF0000u);
}

template <typename InT>
__global__ void __launch_bounds__(256) patch_gather444_kernel(const InT* __restrict__ vol, bf16* __restrict__ out, int B,
                                                              int D, int H, int W, int gd, int gh, int gw) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  const long long total = static_cast<long long>(B) * gd * gh * 4 * gw;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    long long t = idx;
    const int w = static_cast<int>(t % gw); t /= gw;
    const int i = static_cast<int>(t & 3); t >>= 2;
    const int h = static_cast<int>(t % gh); t /= gh;
    const int d = static_cast<int>(t % gd); t /= gd;
    const int b = static_cast<int>(t);
    const int zd = d * 4 + i;
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int zh = h * 4 + j;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (zd < D && zh < H) load4<InT>(vol + ((static_cast<long long>(b) * D + zd) * H + zh) * W + w * 4, v);
      o[2 * j] = pack_bf16(v[0], v[1]);
      o[2 * j + 1] = pack_bf16(v[2], v[3]);
    }
    const long long token = ((static_cast<long long>(b) * gd + d) * gh + h) * gw + w;
    st_global_v8(out + token * 64 + i * 16, o);
  }
}

// Copy the overlapping box of two channels-last grids and zero-fill the rest of dst.
// pad (models/swin_transformer_3d.py:457-461) when dst is larger, crop (:508) when smaller.
template <typename IdxT>
__global__ void grid_copy_kernel(const float* __restrict__ src, int sD, int sH, int sW, float* __restrict__ dst,
                                 int dD, int dH, int dW, int B, int C4) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  const IdxT total = static_cast<IdxT>(B) * dD * dH * dW * C4;
  const IdxT stride = static_cast<IdxT>(gridDim.x) * blockDim.x;
  for (IdxT idx = static_cast<IdxT>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    IdxT t = idx / static_cast<IdxT>(C4);
    const int c = static_cast<int>(idx - t * static_cast<IdxT>(C4));
    IdxT u = t / static_cast<IdxT>(dW);
    const int w = static_cast<int>(t - u * static_cast<IdxT>(dW)); t = u;
    u = t / static_cast<IdxT>(dH);
    const int h = static_cast<int>(t - u * static_cast<IdxT>(dH)); t = u;
    u = t / static_cast<IdxT>(dD);
    const int d = static_cast<int>(t - u * static_cast<IdxT>(dD));
    const int b = static_cast<int>(u);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d < sD && h < sH && w < sW)
      v = reinterpret_cast<const float4*>(src)[(((static_cast<long long>(b) * sD + d) * sH + h) * sW + w) * C4 + c];
    reinterpret_cast<float4*>(dst)[idx] = v;
  }
}

// PatchMerging gather (models/swin_transformer_3d.py:553-568): out[(b,d,h,w), q*C + c] = x[b, 2d+od, 2h+oh, 2w+ow, c]
// with q enumerating (od,oh,ow) = (0,0,0),(1,0,0),(0,1,0),(0,0,1),(1,1,0),(1,0,1),(0,1,1),(1,1,1); positions
// outside the REAL grid (rD,rH,rW) read as zero (crop + odd pad); x lives on the padded stage grid (pD,pH,pW).
// scatter=true runs the same index map backwards (gradient): x[...] = out[...], untouched positions keep
// their (pre-zeroed) value.
// IdxT = unsigned (problems below 2^31 float4 elements, i.e. all of them in practice): the index decomposition is
// four 32-bit divisions instead of six 64-bit ones, which were costing more than the two memory accesses they address.
template <bool SCATTER, typename IdxT>
__global__ void merge_gather_kernel(float* __restrict__ x, int pD, int pH, int pW, int rD, int rH, int rW,
                                    float* __restrict__ out, int oD, int oH, int oW, int B, int C4) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  const IdxT total = static_cast<IdxT>(B) * oD * oH * oW * 8 * C4;
  const IdxT stride = static_cast<IdxT>(gridDim.x) * blockDim.x;
  for (IdxT idx = static_cast<IdxT>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    IdxT t = idx / static_cast<IdxT>(C4);
    const int c = static_cast<int>(idx - t * static_cast<IdxT>(C4));
    const int q = static_cast<int>(t & 7); t >>= 3;
    IdxT u = t / static_cast<IdxT>(oW);
    const int w = static_cast<int>(t - u * static_cast<IdxT>(oW)); t = u;
    u = t / static_cast<IdxT>(oH);
    const int h = static_cast<int>(t - u * static_cast<IdxT>(oH)); t = u;
    u = t / static_cast<IdxT>(oD);
    const int d = static_cast<int>(t - u * static_cast<IdxT>(oD));
    const int b = static_cast<int>(u);
    // q -> (od, oh, ow) in the reference's concatenation order
    const int od = (0xB2 >> q) & 1;  // q: 0 1 2 3 4 5 6 7 -> 0 1 0 0 1 1 0 1
    const int oh = (0xD4 >> q) & 1;  //                    -> 0 0 1 0 1 0 1 1
    const int ow = (0xE8 >> q) & 1;  //                    -> 0 0 0 1 0 1 1 1
    const int sd = 2 * d + od, sh = 2 * h + oh, sw = 2 * w + ow;
    const bool ok = sd < rD && sh < rH && sw < rW;
    float4* xp = reinterpret_cast<float4*>(x) + (((static_cast<long long>(b) * pD + sd) * pH + sh) * pW + sw) * C4 + c;
    float4* op = reinterpret_cast<float4*>(out) + idx;
    if (SCATTER) {
      if (ok) *xp = *op;
    } else {
      *op = ok ? *xp : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// dst_bf16[r, c] = src_f32[r, c] * row_scale[r / rows_per_group]
__global__ void cast_rows_kernel(const float* __restrict__ src, bf16* __restrict__ dst, const float* __restrict__ scale,
                                 int rows_per_group, long long rows, int C4) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  const long long total = rows * C4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const float4 v = reinterpret_cast<const float4*>(src)[idx];
    const float s = scale ? scale[(idx / C4) / rows_per_group] : 1.f;
    uint2 u;
    u.x = pack_bf16(v.x * s, v.y * s);
    u.y = pack_bf16(v.z * s, v.w * s);
    reinterpret_cast<uint2*>(dst)[idx] = u;
  }
}

// fp32 -> bf16 over a flat range (weights)
__global__ void cast_flat_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = f2b(src[i]);
}

// out[b, c] = mean_t x[b, t, c]  (AdaptiveAvgPool3d(1), models/swin_transformer_3d.py:696) and its gradient
__global__ void token_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int T, int C) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* p = x + static_cast<long long>(b) * T * C + c;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += p[static_cast<long long>(t) * C];
  out[static_cast<long long>(b) * C + c] = s / T;
}
__global__ void token_mean_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int T, int C) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float v = dout[static_cast<long long>(b) * C + c] / T;
  float* p = dx + static_cast<long long>(b) * T * C + c;
  for (int t = 0; t < T; ++t) p[static_cast<long long>(t) * C] = v;
}

// logits[b,k] = sum_f feat[b,f] * W[k,f] + bias[k]   (head Linear, fp32; K is a handful of classes)
__global__ void head_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ W,
                                const float* __restrict__ bias, float* __restrict__ logits, int B, int K, int F) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * K) return;
  const int b = warp / K, k = warp % K;
  float s = 0.f;
  for (int f = lane; f < F; f += 32) s += feat[static_cast<long long>(b) * F + f] * W[static_cast<long long>(k) * F + f];
  s = warp_sum(s);
  if (lane == 0) logits[warp] = s + (bias ? bias[k] : 0.f);
}
// dfeat[b,f] = sum_k dl[b,k] W[k,f];  dW[k,f] += sum_b dl[b,k] feat[b,f];  db[k] += sum_b dl[b,k]
__global__ void head_bwd_kernel(const float* __restrict__ dl, const float* __restrict__ feat,
                                const float* __restrict__ W, float* __restrict__ dfeat, float* __restrict__ dW,
                                float* __restrict__ db, int B, int K, int F) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < F) {
    for (int b = 0; b < B; ++b) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += dl[b * K + k] * W[static_cast<long long>(k) * F + f];
      dfeat[static_cast<long long>(b) * F + f] = s;
    }
    for (int k = 0; k < K; ++k) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += dl[b * K + k] * feat[static_cast<long long>(b) * F + f];
      dW[static_cast<long long>(k) * F + f] += s;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < K && db != nullptr) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dl[b * K + threadIdx.x];
    db[threadIdx.x] += s;
  }
}

// ViT token assembly (models/vit_3d.py:447-449): x[b,0,:] = cls + pos[0]; x[b,1+t,:] = emb[b,t,:] + pos[1+t]
__global__ void vit_assemble_kernel(const float* __restrict__ emb, const float* __restrict__ cls,
                                    const float* __restrict__ pos, float* __restrict__ x, int B, int T, int C4) {
  const long long total = static_cast<long long>(B) * (T + 1) * C4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    long long t = idx;
    const int c = static_cast<int>(t % C4); t /= C4;
    const int tok = static_cast<int>(t % (T + 1)); t /= (T + 1);
    const int b = static_cast<int>(t);
    float4 v = tok == 0 ? reinterpret_cast<const float4*>(cls)[c]
                        : reinterpret_cast<const float4*>(emb)[(static_cast<long long>(b) * T + tok - 1) * C4 + c];
    const float4 p = reinterpret_cast<const float4*>(pos)[static_cast<long long>(tok) * C4 + c];
    reinterpret_cast<float4*>(x)[idx] = make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
  }
}
// gradient of the above: dpos[tok] += sum_b dx[b,tok]; dcls += sum_b dx[b,0]  (demb is dx[:,1:] read in place)
__global__ void vit_assemble_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dcls,
                                        float* __restrict__ dpos, int B, int T, int C) {
  const long long total = static_cast<long long>(T + 1) * C;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dx[static_cast<long long>(b) * total + idx];
    dpos[idx] += s;
    if (idx < C) dcls[idx] += s;
  }
}

// MixUp of fp16 volumes (dataset/dataset.py:276-281): the reference works in place on the fp16 tensors,
// `sample1.mul_(alpha)` (rounds to fp16) then `.add_(sample2, alpha=1 - alpha)` (fp32 arithmetic, one more rounding):
// out[b] = fp16(fp16(lam[b] * x[b]) + fp16(1 - lam[b]) * x[perm[b]]), 16-byte accesses.  lam[b] == 1 copies x[b].
// (The reference mixes on the CPU: `mul_` takes its Python scalar as fp32, `add_` rounds its alpha to the TENSOR's dtype
// first -- checked against torch's CPU kernels element by element -- and both compute in fp32.)
__device__ __forceinline__ float mix_f16(float l, float a, float c) {
  return fmaf(__half2float(__float2half_rn(1.0f - l)), c, __half2float(__float2half_rn(l * a)));
}
__global__ void mixup_f16_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, const float* __restrict__ lam,
                                 const int* __restrict__ perm, int B, long long vec_per_sample) {
  const long long total = static_cast<long long>(B) * vec_per_sample;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / vec_per_sample);
    const long long r = i - static_cast<long long>(b) * vec_per_sample;
    const float l = lam[b];
    uint4 a = x[i];
    if (l != 1.0f) {
      const uint4 c = x[static_cast<long long>(perm[b]) * vec_per_sample + r];
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
      const __half2* ch = reinterpret_cast<const __half2*>(&c);
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fa = __half22float2(ah[k]), fc = __half22float2(ch[k]);
        oh[k] = __floats2half2_rn(mix_f16(l, fa.x, fc.x), mix_f16(l, fa.y, fc.y));
      }
      a = o;
    }
    out[i] = a;
  }
}

inline unsigned grid_for(long long total, int block) {
  long long g = ceil_div_ll(total, block);
  const long long cap = static_cast<long long>(vsn_num_sms()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}


// ---- the input step before the path: fp16 cache -> MixUp -> z-score, all on the device ----
// Per-volume sum and sum of squares of the (optionally mixed) fp16 volume, fp32 per thread, double across threads:
// monai NormalizeIntensity() = (x - mean) / std over the whole image, population std, std == 0 -> no division
// (train/train_transformer.py:1729-1752 puts it last in every transform chain, i.e. after dataset/dataset.py's MixUp).
// grid = (chunks, B); scratch[b] = {sum, sumsq} must be zero.
__global__ void __launch_bounds__(256) volume_stats_f16_kernel(const uint4* __restrict__ x, const float* __restrict__ lam,
                                                               const int* __restrict__ perm, long long vec_per_sample,
                                                               double* __restrict__ scratch) {
  const int b = blockIdx.y;
  const float l = lam != nullptr ? lam[b] : 1.0f;
  const uint4* xa = x + static_cast<long long>(b) * vec_per_sample;
  const uint4* xc = l != 1.0f ? x + static_cast<long long>(perm[b]) * vec_per_sample : nullptr;
  double s1 = 0.0, s2 = 0.0;
  for (long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < vec_per_sample;
       i0 += static_cast<long long>(gridDim.x) * blockDim.x * 8) {
    float f1 = 0.f, f2 = 0.f;                   // at most 64 voxels in fp32 before they go to the double sums
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long long i = i0 + static_cast<long long>(u) * gridDim.x * blockDim.x;
      if (i < vec_per_sample) {
        const uint4 a = xa[i];
        const __half2* ah = reinterpret_cast<const __half2*>(&a);
        uint4 c = make_uint4(0, 0, 0, 0);
        if (xc != nullptr) c = xc[i];
        const __half2* ch = reinterpret_cast<const __half2*>(&c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 v = __half22float2(ah[k]);
          if (xc != nullptr) {
            const float2 w = __half22float2(ch[k]);
            v = __half22float2(__floats2half2_rn(mix_f16(l, v.x, w.x), mix_f16(l, v.y, w.y)));
          }
          f1 += v.x + v.y;
          f2 = fmaf(v.x, v.x, fmaf(v.y, v.y, f2));
        }
      }
    }
    s1 += f1; s2 += f2;
  }
  __shared__ double r1[8], r2[8];
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < 8; ++i) { a += r1[i]; c += r2[i]; }
    atomicAdd(scratch + 2 * b, a);
    atomicAdd(scratch + 2 * b + 1, c);
  }
}

// stats[b] = {mean, 1/std} (1 when std == 0); clears the scratch sums for the next batch
__global__ void volume_stats_finish_kernel(double* __restrict__ scratch, float* __restrict__ stats, int B, double n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double mean = scratch[2 * b] / n;
  double var = scratch[2 * b + 1] / n - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const double sd = sqrt(var);
  stats[2 * b] = static_cast<float>(mean);
  stats[2 * b + 1] = sd > 0.0 ? static_cast<float>(1.0 / sd) : 1.0f;
  scratch[2 * b] = 0.0;
  scratch[2 * b + 1] = 0.0;
}

// out[b] = fp16((mix(x[b], x[perm[b]]) - mean[b]) * rstd[b]): MixUp and NormalizeIntensity in one pass over the volume
__global__ void __launch_bounds__(256) mixup_zscore_f16_kernel(const uint4* __restrict__ x, uint4* __restrict__ out,
                                                               const float* __restrict__ lam, const int* __restrict__ perm,
                                                               const float* __restrict__ stats, int B,
                                                               long long vec_per_sample) {
  pdl_trigger();
  const long long total = static_cast<long long>(B) * vec_per_sample;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / vec_per_sample);
    const long long r = i - static_cast<long long>(b) * vec_per_sample;
    const float l = lam != nullptr ? lam[b] : 1.0f;
    const float mean = stats[2 * b], rstd = stats[2 * b + 1];
    const uint4 a = x[i];
    uint4 c = make_uint4(0, 0, 0, 0);
    if (l != 1.0f) c = x[static_cast<long long>(perm[b]) * vec_per_sample + r];
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* ch = reinterpret_cast<const __half2*>(&c);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 v = __half22float2(ah[k]);
      if (l != 1.0f) {
        const float2 w = __half22float2(ch[k]);
        v = __half22float2(__floats2half2_rn(mix_f16(l, v.x, w.x), mix_f16(l, v.y, w.y)));
      }
      oh[k] = __floats2half2_rn((v.x - mean) * rstd, (v.y - mean) * rstd);
    }
    out[i] = o;
  }
}


// ---- test-time augmentation views (eval/test_time_augmentation.py:221-354) -------------------------------------
// out[b, v, z] = trilinear sample of vol[b] at M_v * (z, 1): every view of the reference's TTA -- identity, flip,
// small rotations + translations (RandAffine, bilinear, border padding), centre crop + trilinear resize -- is an affine
// map from the output voxel index to source voxel coordinates, so all views of all volumes come out of one launch and
// the model runs once on a [B*V] batch instead of V batch-1 forwards with host-side resampling in between.
// views [V][18]: a row-major 3x4 matrix (d, h, w), then the box lo[3], hi[3] (voxel indices) the coordinates and the
// upper neighbours are clamped to: the whole volume for padding_mode = "border", the crop box for crop + resize.
__global__ void __launch_bounds__(256) tta_views_f16_kernel(const __half* __restrict__ vol, __half* __restrict__ out,
                                                            const float* __restrict__ mats, int B, int V, int D, int H,
                                                            int W) {
  const long long per = static_cast<long long>(D) * H * W;
  const long long total = static_cast<long long>(B) * V * per;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long t = idx;
    const int w = static_cast<int>(t % W); t /= W;
    const int h = static_cast<int>(t % H); t /= H;
    const int d = static_cast<int>(t % D); t /= D;
    const int v = static_cast<int>(t % V);
    const int b = static_cast<int>(t / V);
    const float* m = mats + v * 18;
    float sd = fmaf(m[0], d, fmaf(m[1], h, fmaf(m[2], w, m[3])));
    float sh = fmaf(m[4], d, fmaf(m[5], h, fmaf(m[6], w, m[7])));
    float sw = fmaf(m[8], d, fmaf(m[9], h, fmaf(m[10], w, m[11])));
    sd = fminf(fmaxf(sd, m[12]), m[15]);
    sh = fminf(fmaxf(sh, m[13]), m[16]);
    sw = fminf(fmaxf(sw, m[14]), m[17]);
    const int d0 = static_cast<int>(sd), h0 = static_cast<int>(sh), w0 = static_cast<int>(sw);
    const int d1 = min(d0 + 1, static_cast<int>(m[15])), h1 = min(h0 + 1, static_cast<int>(m[16])),
              w1 = min(w0 + 1, static_cast<int>(m[17]));
    const float fd = sd - d0, fh = sh - h0, fw = sw - w0;
    const __half* src = vol + static_cast<long long>(b) * per;
    auto at = [&](int a, int c, int e) { return __half2float(src[(static_cast<long long>(a) * H + c) * W + e]); };
    float r;
    if (fd == 0.f && fh == 0.f && fw == 0.f) {
      r = at(d0, h0, w0);                                    // identity / flip: a bit-exact copy
    } else {
      const float c00 = at(d0, h0, w0) * (1.f - fw) + at(d0, h0, w1) * fw;
      const float c01 = at(d0, h1, w0) * (1.f - fw) + at(d0, h1, w1) * fw;
      const float c10 = at(d1, h0, w0) * (1.f - fw) + at(d1, h0, w1) * fw;
      const float c11 = at(d1, h1, w0) * (1.f - fw) + at(d1, h1, w1) * fw;
      const float c0 = c00 * (1.f - fh) + c01 * fh, c1 = c10 * (1.f - fh) + c11 * fh;
      r = c0 * (1.f - fd) + c1 * fd;
    }
    out[idx] = __float2half_rn(r);
  }
}

// ---- ViT to_patch_embedding head: Rearrange + LayerNorm(P) in one pass (models/vit_3d.py:364-371) ---------------
// The [T, P] fp32 rows of the Rearrange never exist in HBM (at vit-3c: 637 MB written and read twice per step).
// A block owns the SLAB of one (b, d, h): the pd*ph volume rows of its gw patches, W voxels each.  It stages the slab
// in shared memory with 16-byte loads -- for fixed i the ph rows are one contiguous run of the volume (4.6 KB at
// vit-3c), where per-patch reads would be 32-byte pieces at a 288-byte stride, half of every DRAM burst wasted (the
// first version of this kernel, one block per patch, ran at 1.5 TB/s) -- and then works on the patches out of it.
constexpr int PLN_MAXW = 16;      // voxels of one patch row (pgrad: held by a thread)
constexpr int PLN_GROUPS = 32;    // float4 groups per lane in the forward (P <= 4096)

struct PatchGeom { int B, D, H, W, gd, gh, gw, pd, ph, pw, pw_shift, ph_shift, Wp; };   // shifts: log2 or -1; Wp = gw*pw

// slab[row = i*ph + j][x < Wp] <- vol[b, d*pd + i, h*ph + j, x] (zero beyond D, H, W); all threads of the block
template <typename InT>
__device__ __forceinline__ void stage_slab(const InT* __restrict__ vol, const PatchGeom& g, unsigned slab_id, InT* slab) {
  unsigned t = slab_id;
  const int h = t % g.gh; t /= g.gh;
  const int d = t % g.gd;
  const long long b = t / g.gd;
  constexpr int VE = 16 / sizeof(InT);                 // elements per 16-byte chunk
  const int rows = g.pd * g.ph;
  const bool vec = (g.W % VE == 0) && (g.Wp % VE == 0) && ((reinterpret_cast<uintptr_t>(vol) & 15) == 0);
  if (vec) {
    // batches of eight chunks per thread: all eight loads are in flight before the first shared-memory store needs one
    // (one load - one store per trip made every trip a full memory round trip: 16 of them per thread at vit-3c)
    const int cpr = g.Wp / VE, total = rows * cpr;
    constexpr int NB = 8;
    for (int base = threadIdx.x; base < total; base += NB * blockDim.x) {
      uint4 u[NB];
      int dst[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const int idx = base + k * blockDim.x;
        u[k] = make_uint4(0, 0, 0, 0);
        dst[k] = -1;
        if (idx < total) {
          const int row = idx / cpr, c = idx - row * cpr;
          const int pi = g.ph_shift >= 0 ? row >> g.ph_shift : row / g.ph, pj = row - pi * g.ph;
          const int zd = d * g.pd + pi, zh = h * g.ph + pj;
          dst[k] = row * g.Wp + c * VE;
          if (zd < g.D && zh < g.H && (c + 1) * VE <= g.W)
            u[k] = *reinterpret_cast<const uint4*>(vol + ((b * g.D + zd) * g.H + zh) * g.W + c * VE);
        }
      }
#pragma unroll
      for (int k = 0; k < NB; ++k)
        if (dst[k] >= 0) *reinterpret_cast<uint4*>(slab + dst[k]) = u[k];
    }
  } else {
    const int total = rows * g.Wp;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int row = idx / g.Wp, x = idx - row * g.Wp;
      const int pi = row / g.ph, pj = row - pi * g.ph;
      const int zd = d * g.pd + pi, zh = h * g.ph + pj;
      InT v;
      memset(&v, 0, sizeof(InT));
      if (zd < g.D && zh < g.H && x < g.W) v = vol[((b * g.D + zd) * g.H + zh) * g.W + x];
      slab[idx] = v;
    }
  }
}

// Forward: block = one slab, warp w = patch w of the slab.  Lane l walks the float4 groups j = l + 32 i of the patch's
// P values exactly as the generic LayerNorm kernel walks a row (same partial sums, same butterfly, same expressions),
// so statistics and normalised values are bit-identical to patch_gather + vsn_layernorm_fwd; group j is the 4 voxels
// k0 = 4 j % pw of patch row 4 j / pw.  pw % 4 == 0, P % 128 == 0.
template <typename InT>
__global__ void patch_ln_fwd_kernel(const InT* __restrict__ vol, PatchGeom g, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, bf16* __restrict__ y, float* __restrict__ mean_out,
                                    float* __restrict__ rstd_out, float eps) {
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t pln_smem[];
  InT* slab = reinterpret_cast<InT*>(pln_smem);
  stage_slab<InT>(vol, g, blockIdx.x, slab);
  __syncthreads();
  const int P = g.pd * g.ph * g.pw, nvec = P >> 2;
  const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int w = threadIdx.x >> 5; w < g.gw; w += nwarps) {
    const unsigned token = blockIdx.x * g.gw + w;
    const InT* base = slab + w * g.pw;
    auto group = [&](int j, float* f) {
      const int e = 4 * j, row = g.pw_shift >= 0 ? e >> g.pw_shift : e / g.pw, k0 = e - row * g.pw;
      load4<InT>(base + row * g.Wp + k0, f);
    };
    float s = 0.f;
    for (int j = lane; j < nvec; j += 32) {
      float f[4];
      group(j, f);
      s += (f[0] + f[1]) + (f[2] + f[3]);
    }
    const float mu = warp_sum(s) / P;
    float q = 0.f;
    for (int j = lane; j < nvec; j += 32) {
      float f[4];
      group(j, f);
      const float a = f[0] - mu, bb = f[1] - mu, c = f[2] - mu, dd = f[3] - mu;
      q += (a * a + bb * bb) + (c * c + dd * dd);
    }
    const float rstd = rsqrtf(warp_sum(q) / P + eps);
    bf16* yr = y + static_cast<long long>(token) * P;
    for (int j = lane; j < nvec; j += 32) {
      float f[4];
      group(j, f);
      const float4 ga = *reinterpret_cast<const float4*>(gamma + 4 * j);
      const float4 be = *reinterpret_cast<const float4*>(beta + 4 * j);
      const float o0 = (f[0] - mu) * rstd * ga.x + be.x, o1 = (f[1] - mu) * rstd * ga.y + be.y;
      const float o2 = (f[2] - mu) * rstd * ga.z + be.z, o3 = (f[3] - mu) * rstd * ga.w + be.w;
      uint2 u;
      u.x = pack_bf16(o0, o1);
      u.y = pack_bf16(o2, o3);
      *reinterpret_cast<uint2*>(yr + 4 * j) = u;
    }
    if (lane == 0) {
      mean_out[token] = mu;
      rstd_out[token] = rstd;
    }
  }
}

// dgamma[c] += sum_t dy[t,c] * xhat[t,c], dbeta[c] += sum_t dy[t,c] with xhat re-gathered from the volume (the input
// of this LayerNorm is data: there is no dx).  Persistent blocks over slabs, thread (i, j) owns the pw columns of
// patch row (i, j): its partial sums stay in registers over all patches, one atomic per column and block at the end.
template <typename InT>
__global__ void patch_ln_pgrad_kernel(const bf16* __restrict__ dy, const InT* __restrict__ vol, PatchGeom g,
                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta, unsigned slabs) {
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t pln_smem[];
  InT* slab = reinterpret_cast<InT*>(pln_smem);
  const int P = g.pd * g.ph * g.pw;
  const int c0 = threadIdx.x * g.pw;                    // blockDim.x == pd * ph: thread = patch row
  float ag[PLN_MAXW], ab[PLN_MAXW];
#pragma unroll
  for (int k = 0; k < PLN_MAXW; ++k) ag[k] = ab[k] = 0.f;
  for (unsigned sid = blockIdx.x; sid < slabs; sid += gridDim.x) {
    __syncthreads();                                    // the previous slab is no longer read
    stage_slab<InT>(vol, g, sid, slab);
    __syncthreads();
    const InT* myrow = slab + threadIdx.x * g.Wp;
    constexpr int TB = 3;                               // patches per trip: their dy loads are issued together
    for (int w0 = 0; w0 < g.gw; w0 += TB) {
      uint2 raw[TB][PLN_MAXW / 4];
      float mu[TB], rs[TB];
#pragma unroll
      for (int u = 0; u < TB; ++u) {
        const bool live = w0 + u < g.gw;
        const unsigned token = sid * g.gw + (live ? w0 + u : w0);
        const bf16* dyr = dy + static_cast<long long>(token) * P + c0;
        mu[u] = mean[token];
        rs[u] = live ? rstd[token] : 0.f;
#pragma unroll
        for (int k4 = 0; k4 < PLN_MAXW / 4; ++k4)
          raw[u][k4] = (live && 4 * k4 < g.pw) ? *reinterpret_cast<const uint2*>(dyr + 4 * k4) : make_uint2(0, 0);
      }
#pragma unroll
      for (int u = 0; u < TB; ++u) {
        const int w = w0 + u < g.gw ? w0 + u : w0;
#pragma unroll
        for (int k4 = 0; k4 < PLN_MAXW / 4; ++k4) {
          if (4 * k4 < g.pw) {
            float v[4];
            load4<InT>(myrow + w * g.pw + 4 * k4, v);
            const float2 d01 = unpack_bf16(raw[u][k4].x), d23 = unpack_bf16(raw[u][k4].y);
            const float dv[4] = {d01.x, d01.y, d23.x, d23.y};     // a dead slot holds zeros: adds nothing
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ab[4 * k4 + k] += dv[k];
              ag[4 * k4 + k] = fmaf(dv[k], (v[k] - mu[u]) * rs[u], ag[4 * k4 + k]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < PLN_MAXW; ++k)
    if (k < g.pw) {
      atomicAdd(dgamma + c0 + k, ag[k]);
      atomicAdd(dbeta + c0 + k, ab[k]);
    }
}

}  // namespace

// in_dtype: 0 fp32, 1 fp16, 2 bf16.  out_bf16: 1 -> bf16 rows (GEMM operand), 0 -> fp32 rows (feeds LayerNorm).
extern "C" int vsn_patch_gather(const void* vol, int in_dtype, void* out, int out_bf16, int B, int D, int H, int W,
                                int pd, int ph, int pw, void* stream) {
  const int gd = ceil_div(D, pd), gh = ceil_div(H, ph), gw = ceil_div(W, pw);
  const long long total = static_cast<long long>(B) * gd * gh * gw * pd * ph;
  if (total == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (pd == 4 && ph == 4 && pw == 4 && out_bf16 && W % 4 == 0 && in_dtype >= 0 && in_dtype <= 2 &&
      (reinterpret_cast<uintptr_t>(vol) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 32 == 0)) {
    const long long items = static_cast<long long>(B) * gd * gh * 4 * gw;
    const unsigned g4 = grid_for(items, 256);
    bf16* o = reinterpret_cast<bf16*>(out);
    if (in_dtype == 0) patch_gather444_kernel<float><<<g4, 256, 0, s>>>(reinterpret_cast<const float*>(vol), o, B, D, H, W, gd, gh, gw);
    else if (in_dtype == 1) patch_gather444_kernel<__half><<<g4, 256, 0, s>>>(reinterpret_cast<const __half*>(vol), o, B, D, H, W, gd, gh, gw);
    else patch_gather444_kernel<__nv_bfloat16><<<g4, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(vol), o, B, D, H, W, gd, gh, gw);
    VSN_LAUNCH_CHECK();
    return 0;
  }
  const unsigned grid = grid_for(total, 256);
#define VSN_PG(InT, OutT)                                                                                          \
  patch_gather_kernel<InT, OutT><<<grid, 256, 0, s>>>(reinterpret_cast<const InT*>(vol), reinterpret_cast<OutT*>(out), \
                                                     B, D, H, W, gd, gh, gw, pd, ph, pw)
  if (in_dtype == 0 && out_bf16) VSN_PG(float, bf16);
  else if (in_dtype == 0) VSN_PG(float, float);
  else if (in_dtype == 1 && out_bf16) VSN_PG(__half, bf16);
  else if (in_dtype == 1) VSN_PG(__half, float);
  else if (in_dtype == 2 && out_bf16) VSN_PG(__nv_bfloat16, bf16);
  else if (in_dtype == 2) VSN_PG(__nv_bfloat16, float);
  else { vsn_set_error("vsn_patch_gather: bad in_dtype %d", in_dtype); return 1; }
#undef VSN_PG
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_grid_copy(const float* src, int sD, int sH, int sW, float* dst, int dD, int dH, int dW, int B,
                             int C, void* stream) {
  VSN_CHECK(C % 4 == 0, "vsn_grid_copy: C must be a multiple of 4");
  const long long total = static_cast<long long>(B) * dD * dH * dW * (C / 4);
  if (total == 0) return 0;
  if (total + 256LL * grid_for(total, 256) < (1LL << 31))
    grid_copy_kernel<unsigned><<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, sD, sH, sW, dst, dD,
                                                                                            dH, dW, B, C / 4);
  else
    grid_copy_kernel<long long><<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, sD, sH, sW, dst, dD,
                                                                                            dH, dW, B, C / 4);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_merge_gather(float* x, int pD, int pH, int pW, int rD, int rH, int rW, float* out, int B, int C,
                                int scatter, void* stream) {
  VSN_CHECK(C % 4 == 0, "vsn_merge_gather: C must be a multiple of 4");
  const int oD = (rD + 1) / 2, oH = (rH + 1) / 2, oW = (rW + 1) / 2;
  const long long total = static_cast<long long>(B) * oD * oH * oW * 8 * (C / 4);
  if (total == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool small = total + 256LL * grid_for(total, 256) < (1LL << 31);     // index + grid stride fit 32 bits
  if (scatter && small)
    merge_gather_kernel<true, unsigned><<<grid_for(total, 256), 256, 0, s>>>(x, pD, pH, pW, rD, rH, rW, out, oD, oH, oW, B, C / 4);
  else if (scatter)
    merge_gather_kernel<true, long long><<<grid_for(total, 256), 256, 0, s>>>(x, pD, pH, pW, rD, rH, rW, out, oD, oH, oW, B, C / 4);
  else if (small)
    merge_gather_kernel<false, unsigned><<<grid_for(total, 256), 256, 0, s>>>(x, pD, pH, pW, rD, rH, rW, out, oD, oH, oW, B, C / 4);
  else
    merge_gather_kernel<false, long long><<<grid_for(total, 256), 256, 0, s>>>(x, pD, pH, pW, rD, rH, rW, out, oD, oH, oW, B, C / 4);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_cast_rows_bf16(const float* src, void* dst, const float* row_scale, int rows_per_group,
                                  long long rows, int C, void* stream) {
  VSN_CHECK(C % 4 == 0, "vsn_cast_rows_bf16: C must be a multiple of 4");
  const long long total = rows * (C / 4);
  if (total == 0) return 0;
  cast_rows_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<bf16*>(dst), row_scale, rows_per_group > 0 ? rows_per_group : 1, rows, C / 4);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_cast_bf16(const float* src, void* dst, long long n, void* stream) {
  if (n == 0) return 0;
  cast_flat_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, reinterpret_cast<bf16*>(dst), n);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_token_mean(const float* x, float* out, int B, int T, int C, int backward, void* stream) {
  if (B == 0 || T == 0 || C == 0) return 0;
  dim3 grid(ceil_div(C, 128), B);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (backward) token_mean_bwd_kernel<<<grid, 128, 0, s>>>(x, out, T, C);   // x = d(out) [B,C], out = dx [B,T,C]
  else token_mean_kernel<<<grid, 128, 0, s>>>(x, out, T, C);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_head_fwd(const float* feat, const float* W, const float* bias, float* logits, int B, int K, int F,
                            void* stream) {
  if (B * K == 0) return 0;
  head_fwd_kernel<<<ceil_div(B * K * 32, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(feat, W, bias, logits, B, K, F);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_head_bwd(const float* dlogits, const float* feat, const float* W, float* dfeat, float* dW,
                            float* db, int B, int K, int F, void* stream) {
  VSN_CHECK(K <= 128, "vsn_head_bwd: at most 128 classes (got %d)", K);
  if (B * K == 0) return 0;
  head_bwd_kernel<<<ceil_div(F, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dlogits, feat, W, dfeat, dW, db, B, K, F);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_vit_assemble(const float* emb, const float* cls, const float* pos, float* x, int B, int T, int C,
                                void* stream) {
  VSN_CHECK(C % 4 == 0, "vsn_vit_assemble: C must be a multiple of 4");
  const long long total = static_cast<long long>(B) * (T + 1) * (C / 4);
  vit_assemble_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(emb, cls, pos, x, B, T, C / 4);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_vit_assemble_bwd(const float* dx, float* dcls, float* dpos, int B, int T, int C, void* stream) {
  const long long total = static_cast<long long>(T + 1) * C;
  vit_assemble_bwd_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dx, dcls, dpos, B, T, C);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_mixup_f16(const void* x, void* out, const float* lam, const int* perm, int B,
                             long long elems_per_sample, void* stream) {
  VSN_CHECK(elems_per_sample % 8 == 0, "vsn_mixup_f16: elements per sample must be a multiple of 8");
  VSN_CHECK(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
            "vsn_mixup_f16: 16-byte aligned volumes expected");
  VSN_CHECK(x != out, "vsn_mixup_f16: not in place (a sample is read as its own partner's input)");
  const long long total = static_cast<long long>(B) * (elems_per_sample / 8);
  if (total == 0) return 0;
  mixup_f16_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), lam, perm, B, elems_per_sample / 8);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_volume_stats_f16(const void* x, const float* lam, const int* perm, int B, long long elems_per_sample,
                                    double* scratch, float* stats, void* stream) {
  VSN_CHECK(elems_per_sample % 8 == 0, "vsn_volume_stats_f16: elements per sample must be a multiple of 8");
  VSN_CHECK(reinterpret_cast<uintptr_t>(x) % 16 == 0, "vsn_volume_stats_f16: 16-byte aligned volumes expected");
  VSN_CHECK((lam == nullptr) == (perm == nullptr), "vsn_volume_stats_f16: lam and perm go together");
  if (B == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long vec = elems_per_sample / 8;
  int chunks = ceil_div(2 * vsn_num_sms(), B);
  const long long maxc = (vec + 255) / 256;
  if (chunks > maxc) chunks = static_cast<int>(maxc);
  if (chunks < 1) chunks = 1;
  volume_stats_f16_kernel<<<dim3(chunks, B), 256, 0, s>>>(reinterpret_cast<const uint4*>(x), lam, perm, vec, scratch);
  VSN_LAUNCH_CHECK();
  volume_stats_finish_kernel<<<ceil_div(B, 128), 128, 0, s>>>(scratch, stats, B, static_cast<double>(elems_per_sample));
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_mixup_zscore_f16(const void* x, void* out, const float* lam, const int* perm, const float* stats, int B,
                                    long long elems_per_sample, void* stream) {
  VSN_CHECK(elems_per_sample % 8 == 0, "vsn_mixup_zscore_f16: elements per sample must be a multiple of 8");
  VSN_CHECK(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
            "vsn_mixup_zscore_f16: 16-byte aligned volumes expected");
  VSN_CHECK(x != out || lam == nullptr, "vsn_mixup_zscore_f16: MixUp is not in place (a sample is read as its partner's input)");
  VSN_CHECK((lam == nullptr) == (perm == nullptr), "vsn_mixup_zscore_f16: lam and perm go together");
  const long long total = static_cast<long long>(B) * (elems_per_sample / 8);
  if (total == 0) return 0;
  mixup_zscore_f16_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), lam, perm, stats, B, elems_per_sample / 8);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_tta_views_f16(const void* vol, void* out, const float* mats, int B, int V, int D, int H, int W,
                                 void* stream) {
  const long long total = static_cast<long long>(B) * V * D * H * W;
  if (total == 0) return 0;
  VSN_CHECK(vol != out, "vsn_tta_views_f16: not in place");
  tta_views_f16_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(vol), reinterpret_cast<__half*>(out), mats, B, V, D, H, W);
  VSN_LAUNCH_CHECK();
  return 0;
}

static int patch_ln_geom(int B, int D, int H, int W, int pd, int ph, int pw, int elem, PatchGeom& g, size_t& smem) {
  VSN_CHECK(pd > 0 && ph > 0 && pw > 0 && pd * ph <= 1024 && pw <= PLN_MAXW && pw % 4 == 0 &&
                (pd * ph * pw) % 128 == 0 && pd * ph * pw <= 128 * PLN_GROUPS,
            "vsn_patch_ln: patch (%d,%d,%d) needs pd*ph <= 1024, pw a multiple of 4 up to %d, P a multiple of 128 up to %d",
            pd, ph, pw, PLN_MAXW, 128 * PLN_GROUPS);
  auto lg = [](int v) { int k = 0; while ((1 << k) < v) ++k; return (1 << k) == v ? k : -1; };
  const int gw = ceil_div(W, pw);
  g = {B, D, H, W, ceil_div(D, pd), ceil_div(H, ph), gw, pd, ph, pw, lg(pw), lg(ph), gw * pw};
  VSN_CHECK(static_cast<long long>(B) * g.gd * g.gh * g.gw < (1LL << 31), "vsn_patch_ln: too many patches");
  smem = static_cast<size_t>(pd) * ph * g.Wp * elem;
  VSN_CHECK(smem <= 200 * 1024, "vsn_patch_ln: the slab of one patch row (%zu bytes) does not fit shared memory", smem);
  return 0;
}

template <typename K>
static int pln_smem_attr(K kernel, size_t smem) {
  if (smem > 48 * 1024) VSN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  return 0;
}

// ViT patch embedding head (models/vit_3d.py:364-371): y [B*gd*gh*gw, P] = LayerNorm(P)(Rearrange(vol)) as the 16-bit
// operand of the embedding GEMM, mean / rstd per patch; the patch rows are read straight from the volume.
extern "C" int vsn_patch_ln_fwd(const void* vol, int in_dtype, int B, int D, int H, int W, int pd, int ph, int pw,
                                const float* gamma, const float* beta, void* y, float* mean, float* rstd, float eps,
                                void* stream) {
  VSN_CHECK(in_dtype >= 0 && in_dtype <= 2, "vsn_patch_ln_fwd: bad in_dtype %d", in_dtype);
  PatchGeom g;
  size_t smem;
  if (int rc = patch_ln_geom(B, D, H, W, pd, ph, pw, in_dtype == 0 ? 4 : 2, g, smem)) return rc;
  const unsigned slabs = static_cast<unsigned>(B) * g.gd * g.gh;
  if (slabs == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  bf16* yo = reinterpret_cast<bf16*>(y);
  const int threads = 32 * (g.gw < 16 ? g.gw : 16);
#define VSN_PLN_FWD(T)                                                                                        \
  {                                                                                                           \
    if (int rc = pln_smem_attr(patch_ln_fwd_kernel<T>, smem)) return rc;                                      \
    patch_ln_fwd_kernel<T><<<slabs, threads, smem, s>>>(reinterpret_cast<const T*>(vol), g, gamma, beta, yo, mean, rstd, eps); \
  }
  if (in_dtype == 0) VSN_PLN_FWD(float)
  else if (in_dtype == 1) VSN_PLN_FWD(__half)
  else VSN_PLN_FWD(__nv_bfloat16)
#undef VSN_PLN_FWD
  VSN_LAUNCH_CHECK();
  return 0;
}

// Parameter gradients of that LayerNorm: dgamma / dbeta [P] += over patches, xhat re-gathered from the volume.
extern "C" int vsn_patch_ln_param_grad(const void* dy, const void* vol, int in_dtype, int B, int D, int H, int W, int pd,
                                       int ph, int pw, const float* mean, const float* rstd, float* dgamma, float* dbeta,
                                       void* stream) {
  VSN_CHECK(in_dtype >= 0 && in_dtype <= 2, "vsn_patch_ln_param_grad: bad in_dtype %d", in_dtype);
  PatchGeom g;
  size_t smem;
  if (int rc = patch_ln_geom(B, D, H, W, pd, ph, pw, in_dtype == 0 ? 4 : 2, g, smem)) return rc;
  const unsigned slabs = static_cast<unsigned>(B) * g.gd * g.gh;
  if (slabs == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bf16* dyp = reinterpret_cast<const bf16*>(dy);
  unsigned grid = 0;
  // persistent blocks: exactly as many as are resident at once (shared memory decides: two or three per SM)
#define VSN_PLN_PG(T)                                                                                         \
  {                                                                                                           \
    if (int rc = pln_smem_attr(patch_ln_pgrad_kernel<T>, smem)) return rc;                                    \
    int per_sm = 0;                                                                                           \
    VSN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, patch_ln_pgrad_kernel<T>, pd * ph, smem)); \
    const unsigned cap = static_cast<unsigned>(per_sm > 0 ? per_sm : 1) * static_cast<unsigned>(vsn_num_sms()); \
    grid = slabs < cap ? slabs : cap;                                                                         \
    patch_ln_pgrad_kernel<T><<<grid, pd * ph, smem, s>>>(dyp, reinterpret_cast<const T*>(vol), g, mean, rstd, dgamma, dbeta, slabs); \
  }
  if (in_dtype == 0) VSN_PLN_PG(float)
  else if (in_dtype == 1) VSN_PLN_PG(__half)
  else VSN_PLN_PG(__nv_bfloat16)
#undef VSN_PLN_PG
  VSN_LAUNCH_CHECK();
  return 0;
}
