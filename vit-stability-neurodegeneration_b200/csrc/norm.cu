// LayerNorm forward / backward and column reductions (HBM-bound kernels).
// Replaces nn.LayerNorm on the path: norm1/norm2 (models/swin_transformer_3d.py:236,255,330,373),
// PatchEmbed3D.norm (:541), PatchMerging.norm (:570), backbone.norm (:693), and the ViT norms
// (models/vit_3d.py:69,98,371,373,401).  eps = 1e-5, biased variance, statistics in fp32.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int LN_WARPS = 8;

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<bf16>(bf16 v) { return b2f(v); }

__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(bf16* p, float v) { *p = f2b(v); }

// One warp per row.  Rows up to 1024 wide are held in registers (one HBM read); wider rows are
// re-read from L1/L2 for the second and third sweep.
template <typename OutT>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const float* __restrict__ x, long long ldx,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, OutT* __restrict__ y,
                                                               long long ldy, float* __restrict__ mean_out,
                                                               float* __restrict__ rstd_out, long long rows, int C,
                                                               float eps) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  OutT* yr = y + row * ldy;
  const int nvec = C >> 2;  // C is a multiple of 4
  if (nvec <= 256) {
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = lane + i * 32;
      if (j < nvec) {
        v[i] = *reinterpret_cast<const float4*>(xr + 4 * j);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mu = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = lane + i * 32;
      if (j < nvec) {
        const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = lane + i * 32;
      if (j < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * j);
        const float4 b = *reinterpret_cast<const float4*>(beta + 4 * j);
        const float o0 = (v[i].x - mu) * rstd * g.x + b.x, o1 = (v[i].y - mu) * rstd * g.y + b.y;
        const float o2 = (v[i].z - mu) * rstd * g.z + b.z, o3 = (v[i].w - mu) * rstd * g.w + b.w;
        if constexpr (sizeof(OutT) == 2) {
          uint2 u;
          u.x = pack_bf16(o0, o1);
          u.y = pack_bf16(o2, o3);
          *reinterpret_cast<uint2*>(yr + 4 * j) = u;
        } else {
          *reinterpret_cast<float4*>(yr + 4 * j) = make_float4(o0, o1, o2, o3);
        }
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = rstd;
    }
  } else {
    float s = 0.f;
    for (int j = lane; j < nvec; j += 32) {
      const float4 t = *reinterpret_cast<const float4*>(xr + 4 * j);
      s += (t.x + t.y) + (t.z + t.w);
    }
    const float mu = warp_sum(s) / C;
    float q = 0.f;
    for (int j = lane; j < nvec; j += 32) {
      const float4 t = *reinterpret_cast<const float4*>(xr + 4 * j);
      const float a = t.x - mu, b = t.y - mu, c = t.z - mu, d = t.w - mu;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    for (int j = lane; j < nvec; j += 32) {
      const float4 t = *reinterpret_cast<const float4*>(xr + 4 * j);
      const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * j);
      const float4 b = *reinterpret_cast<const float4*>(beta + 4 * j);
      const float o0 = (t.x - mu) * rstd * g.x + b.x, o1 = (t.y - mu) * rstd * g.y + b.y;
      const float o2 = (t.z - mu) * rstd * g.z + b.z, o3 = (t.w - mu) * rstd * g.w + b.w;
      if constexpr (sizeof(OutT) == 2) {
        uint2 u;
        u.x = pack_bf16(o0, o1);
        u.y = pack_bf16(o2, o3);
        *reinterpret_cast<uint2*>(yr + 4 * j) = u;
      } else {
        *reinterpret_cast<float4*>(yr + 4 * j) = make_float4(o0, o1, o2, o3);
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
}

// dx = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)) [+ resid_grad]; optional bf16 copy of
// dx scaled per row group (the DropPath factor of the branch that consumes it next).
template <typename DyT>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const DyT* __restrict__ dy, long long lddy,
                                                               const float* __restrict__ x, long long ldx,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ resid_grad, long long ldr,
                                                               float* __restrict__ dx, long long lddx,
                                                               bf16* __restrict__ dx_bf16, long long ldb,
                                                               const float* __restrict__ row_scale,
                                                               int rows_per_group, long long rows, int C) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const DyT* dyr = dy + row * lddy;
  const float* xr = x + row * ldx;
  const float mu = mean[row], rs = rstd[row];
  float s1 = 0.f, s2 = 0.f;
  for (int j = lane * 4; j < C; j += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(xr + j);
    const float4 g = *reinterpret_cast<const float4*>(gamma + j);
    float d0, d1, d2, d3;
    if constexpr (sizeof(DyT) == 2) {
      const uint2 u = *reinterpret_cast<const uint2*>(dyr + j);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
      d0 = a.x; d1 = a.y; d2 = b.x; d3 = b.y;
    } else {
      const float4 t = *reinterpret_cast<const float4*>(dyr + j);
      d0 = t.x; d1 = t.y; d2 = t.z; d3 = t.w;
    }
    d0 *= g.x; d1 *= g.y; d2 *= g.z; d3 *= g.w;
    s1 += (d0 + d1) + (d2 + d3);
    s2 += (d0 * (xv.x - mu) + d1 * (xv.y - mu)) + (d2 * (xv.z - mu) + d3 * (xv.w - mu));
  }
  s1 = warp_sum(s1) / C;
  s2 = warp_sum(s2) * rs / C;  // mean(dy*g*xhat)
  float scale = 1.f;
  if (dx_bf16 != nullptr && row_scale != nullptr) scale = row_scale[row / rows_per_group];
  for (int j = lane * 4; j < C; j += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(xr + j);
    const float4 g = *reinterpret_cast<const float4*>(gamma + j);
    float d0, d1, d2, d3;
    if constexpr (sizeof(DyT) == 2) {
      const uint2 u = *reinterpret_cast<const uint2*>(dyr + j);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
      d0 = a.x; d1 = a.y; d2 = b.x; d3 = b.y;
    } else {
      const float4 t = *reinterpret_cast<const float4*>(dyr + j);
      d0 = t.x; d1 = t.y; d2 = t.z; d3 = t.w;
    }
    float o0 = rs * (d0 * g.x - s1 - (xv.x - mu) * rs * s2);
    float o1 = rs * (d1 * g.y - s1 - (xv.y - mu) * rs * s2);
    float o2 = rs * (d2 * g.z - s1 - (xv.z - mu) * rs * s2);
    float o3 = rs * (d3 * g.w - s1 - (xv.w - mu) * rs * s2);
    if (resid_grad != nullptr) {
      const float4 r = *reinterpret_cast<const float4*>(resid_grad + row * ldr + j);
      o0 += r.x; o1 += r.y; o2 += r.z; o3 += r.w;
    }
    if (dx != nullptr) *reinterpret_cast<float4*>(dx + row * lddx + j) = make_float4(o0, o1, o2, o3);
    if (dx_bf16 != nullptr) {
      uint2 u;
      u.x = pack_bf16(o0 * scale, o1 * scale);
      u.y = pack_bf16(o2 * scale, o3 * scale);
      *reinterpret_cast<uint2*>(dx_bf16 + row * ldb + j) = u;
    }
  }
}


// ---- fast paths: C == LPR * V * 4 (LPR lanes per row, V float4 per lane) -------------------------------------
// A row is held entirely in registers by a group of LPR lanes (32/LPR rows per warp), every lane issues V
// independent 16-byte loads, and the row statistics are butterfly reductions inside the group.
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename OutT, int LPR, int V>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_fast_kernel(const float* __restrict__ x, long long ldx,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, OutT* __restrict__ y,
                                                                    long long ldy, float* __restrict__ mean_out,
                                                                    float* __restrict__ rstd_out, long long rows, float eps) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  constexpr int C = LPR * V * 4;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, l = lane % LPR;
  const long long row = (static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool ok = row < rows;
  const float* xr = x + (ok ? row : 0) * ldx;
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = ok ? *reinterpret_cast<const float4*>(xr + 4 * (l + LPR * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = group_sum<LPR>(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(group_sum<LPR>(q) * (1.0f / C) + eps);
  if (!ok) return;
  OutT* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int j = 4 * (l + LPR * i);
    const float4 g = *reinterpret_cast<const float4*>(gamma + j);
    const float4 b = *reinterpret_cast<const float4*>(beta + j);
    const float o0 = (v[i].x - mu) * rstd * g.x + b.x, o1 = (v[i].y - mu) * rstd * g.y + b.y;
    const float o2 = (v[i].z - mu) * rstd * g.z + b.z, o3 = (v[i].w - mu) * rstd * g.w + b.w;
    if constexpr (sizeof(OutT) == 2) {
      uint2 u;
      u.x = pack_bf16(o0, o1);
      u.y = pack_bf16(o2, o3);
      *reinterpret_cast<uint2*>(yr + j) = u;
    } else {
      *reinterpret_cast<float4*>(yr + j) = make_float4(o0, o1, o2, o3);
    }
  }
  if (l == 0) {
    if (mean_out) mean_out[row] = mu;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// Backward with the parameter gradients fused: every lane owns fixed columns for all the rows it visits
// (persistent row loop), accumulates dgamma / dbeta for them in registers, and the block adds its partial
// sums to global memory once (C atomics per block) -- dy and x are read exactly once.
template <typename DyT, int LPR, int V>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_fast_kernel(const DyT* __restrict__ dy, long long lddy,
                                                                    const float* __restrict__ x, long long ldx,
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ resid_grad, long long ldr,
                                                                    float* __restrict__ dx, long long lddx,
                                                                    bf16* __restrict__ dx_bf16, long long ldb,
                                                                    const float* __restrict__ row_scale,
                                                                    int rows_per_group, long long rows,
                                                                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  constexpr int C = LPR * V * 4;
  constexpr int RPW = 32 / LPR;
  __shared__ float red[LN_WARPS][C + 4];
  const int lane = threadIdx.x & 31, l = lane % LPR, warp = threadIdx.x >> 5;
  float4 g[V];
#pragma unroll
  for (int i = 0; i < V; ++i) g[i] = *reinterpret_cast<const float4*>(gamma + 4 * (l + LPR * i));
  float4 ag[V], ab[V];
#pragma unroll
  for (int i = 0; i < V; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool want_pg = dgamma != nullptr;
  const long long stride = static_cast<long long>(gridDim.x) * LN_WARPS * RPW;
  for (long long row = (static_cast<long long>(blockIdx.x) * LN_WARPS + warp) * RPW + lane / LPR;
       row - lane / LPR < rows; row += stride) {
    const bool ok = row < rows;
    const long long r = ok ? row : 0;
    const float mu = mean[r], rs = rstd[r];
    float4 xv[V], d[V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int j = 4 * (l + LPR * i);
      xv[i] = *reinterpret_cast<const float4*>(x + r * ldx + j);
      if constexpr (sizeof(DyT) == 2) {
        const uint2 u = *reinterpret_cast<const uint2*>(dy + r * lddy + j);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        d[i] = make_float4(a.x, a.y, b.x, b.y);
      } else {
        d[i] = *reinterpret_cast<const float4*>(dy + r * lddy + j);
      }
      if (!ok) d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      xv[i].x = (xv[i].x - mu) * rs; xv[i].y = (xv[i].y - mu) * rs;      // xhat
      xv[i].z = (xv[i].z - mu) * rs; xv[i].w = (xv[i].w - mu) * rs;
      if (want_pg) {
        ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
        ag[i].x = fmaf(d[i].x, xv[i].x, ag[i].x); ag[i].y = fmaf(d[i].y, xv[i].y, ag[i].y);
        ag[i].z = fmaf(d[i].z, xv[i].z, ag[i].z); ag[i].w = fmaf(d[i].w, xv[i].w, ag[i].w);
      }
      d[i].x *= g[i].x; d[i].y *= g[i].y; d[i].z *= g[i].z; d[i].w *= g[i].w;
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * xv[i].x + d[i].y * xv[i].y) + (d[i].z * xv[i].z + d[i].w * xv[i].w);
    }
    s1 = group_sum<LPR>(s1) * (1.0f / C);
    s2 = group_sum<LPR>(s2) * (1.0f / C);    // mean(dy*g*xhat)
    if (!ok) continue;
    float scale = 1.f;
    if (dx_bf16 != nullptr && row_scale != nullptr) scale = row_scale[row / rows_per_group];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int j = 4 * (l + LPR * i);
      float o0 = rs * (d[i].x - s1 - xv[i].x * s2), o1 = rs * (d[i].y - s1 - xv[i].y * s2);
      float o2 = rs * (d[i].z - s1 - xv[i].z * s2), o3 = rs * (d[i].w - s1 - xv[i].w * s2);
      if (resid_grad != nullptr) {
        const float4 rg = *reinterpret_cast<const float4*>(resid_grad + row * ldr + j);
        o0 += rg.x; o1 += rg.y; o2 += rg.z; o3 += rg.w;
      }
      if (dx != nullptr) *reinterpret_cast<float4*>(dx + row * lddx + j) = make_float4(o0, o1, o2, o3);
      if (dx_bf16 != nullptr) {
        uint2 u;
        u.x = pack_bf16(o0 * scale, o1 * scale);
        u.y = pack_bf16(o2 * scale, o3 * scale);
        *reinterpret_cast<uint2*>(dx_bf16 + row * ldb + j) = u;
      }
    }
  }
  if (!want_pg) return;
  // fold the row groups of the warp, then the warps of the block (smem, dgamma then dbeta), then one atomic
  // per column and block
#pragma unroll
  for (int i = 0; i < V; ++i) {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      ag[i].x += __shfl_xor_sync(0xffffffffu, ag[i].x, o); ag[i].y += __shfl_xor_sync(0xffffffffu, ag[i].y, o);
      ag[i].z += __shfl_xor_sync(0xffffffffu, ag[i].z, o); ag[i].w += __shfl_xor_sync(0xffffffffu, ag[i].w, o);
      ab[i].x += __shfl_xor_sync(0xffffffffu, ab[i].x, o); ab[i].y += __shfl_xor_sync(0xffffffffu, ab[i].y, o);
      ab[i].z += __shfl_xor_sync(0xffffffffu, ab[i].z, o); ab[i].w += __shfl_xor_sync(0xffffffffu, ab[i].w, o);
    }
  }
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    if (lane < LPR) {
#pragma unroll
      for (int i = 0; i < V; ++i) *reinterpret_cast<float4*>(&red[warp][4 * (l + LPR * i)]) = which ? ab[i] : ag[i];
    }
    __syncthreads();
    for (int col = threadIdx.x; col < C; col += LN_WARPS * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) t += red[w][col];
      atomicAdd((which ? dbeta : dgamma) + col, t);
    }
    __syncthreads();
  }
}

// ---- PatchMerging: gather of the 2x2x2 neighbours fused into LayerNorm(8C) ------------------------------------
// (models/swin_transformer_3d.py:553-572: x0..x7 concatenated in the order (od,oh,ow) = (0,0,0),(1,0,0),(0,1,0),
// (0,0,1),(1,1,0),(1,0,1),(0,1,1),(1,1,1), positions beyond the real grid read as zero, then norm(8C), then the
// reduction Linear.)  The merged row never exists in HBM: a warp owns one merged row, lane l holds the float4 at
// columns 128 i + 4 l (i < V = C/16) -- the same element-to-lane map and the same summation order as
// ln_fwd_fast_kernel<., 32, V>, so the statistics are bit-identical to gather + LayerNorm -- and reads each of them
// from its source token.  The backward writes dx straight to the source tokens (every real token of the stage grid
// belongs to exactly one merged row) and accumulates dgamma / dbeta in registers over a persistent row loop.
struct MergeGeom {
  int pD, pH, pW;     // padded stage grid the source rows live on
  int rD, rH, rW;     // real grid (crop): sources beyond it read as zero
  int oD, oH, oW;     // merged grid
};
// source row of segment q of merged row `row`, or -1 when it lies beyond the real grid
__device__ __forceinline__ long long merge_src_row(const MergeGeom& g, long long row, int q) {
  unsigned t = static_cast<unsigned>(row);            // rows < 2^31 (checked by the host): 32-bit divisions
  unsigned u = t / static_cast<unsigned>(g.oW);
  const int w = static_cast<int>(t - u * g.oW); t = u;
  u = t / static_cast<unsigned>(g.oH);
  const int h = static_cast<int>(t - u * g.oH); t = u;
  u = t / static_cast<unsigned>(g.oD);
  const int d = static_cast<int>(t - u * g.oD);
  const long long b = u;
  const int sd = 2 * d + ((0xB2 >> q) & 1), sh = 2 * h + ((0xD4 >> q) & 1), sw = 2 * w + ((0xE8 >> q) & 1);
  if (sd >= g.rD || sh >= g.rH || sw >= g.rW) return -1;
  return ((b * g.pD + sd) * g.pH + sh) * g.pW + sw;
}

template <int C>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_merge_fwd_kernel(const float* __restrict__ x, MergeGeom geo,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta,
                                                                     bf16* __restrict__ y, float* __restrict__ mean_out,
                                                                     float* __restrict__ rstd_out, long long rows,
                                                                     float eps) {
  pdl_trigger();
  constexpr int C8 = 8 * C, V = C8 / 128;
  static_assert(C % 32 == 0 && C8 % 128 == 0, "a float4 never straddles two segments and lanes cover the row evenly");
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  // the eight source rows of this merged row: lane q < 8 computes its own, then broadcasts
  long long mine = lane < 8 ? merge_src_row(geo, row, lane) : -1;
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int j = 128 * i + 4 * lane, q = j / C, cc = j - q * C;
    const long long src = __shfl_sync(0xffffffffu, mine, q);
    v[i] = src >= 0 ? *reinterpret_cast<const float4*>(x + src * C + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = warp_sum(s) * (1.0f / C8);
  float qs = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
    qs += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(qs) * (1.0f / C8) + eps);
  bf16* yr = y + row * C8;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int j = 128 * i + 4 * lane;
    const float4 g = *reinterpret_cast<const float4*>(gamma + j);
    const float4 b = *reinterpret_cast<const float4*>(beta + j);
    uint2 u;
    u.x = pack_bf16((v[i].x - mu) * rstd * g.x + b.x, (v[i].y - mu) * rstd * g.y + b.y);
    u.y = pack_bf16((v[i].z - mu) * rstd * g.z + b.z, (v[i].w - mu) * rstd * g.w + b.w);
    *reinterpret_cast<uint2*>(yr + j) = u;
  }
  if (lane == 0) {
    mean_out[row] = mu;
    rstd_out[row] = rstd;
  }
}

// CH = float4 per lane handled at a time.  V == CH: the row stays in registers between the statistics and the dx
// pass; V > CH: the second pass re-reads x / dy (the warp's own 4.5 KB per row: L1 hits) so that the registers go
// to the dgamma / dbeta accumulators of ALL the lane's columns.  The kernel is latency-bound on its register budget
// (accumulators + one row): blocks of MB_WARPS = 4 warps so that three or four of them share an SM.
constexpr int MB_WARPS = 4;
template <int C>
__global__ void __launch_bounds__(MB_WARPS * 32, (C <= 96 ? 4 : 2)) ln_merge_bwd_kernel(
    const bf16* __restrict__ dy, const float* __restrict__ x, MergeGeom geo, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, float* __restrict__ dx, long long rows,
    float* __restrict__ dgamma, float* __restrict__ dbeta, bf16* __restrict__ dx_bf16,
    const float* __restrict__ row_scale) {
  pdl_trigger();
  constexpr int C8 = 8 * C, V = C8 / 128, CH = 6, NCH = V / CH;
  static_assert(V % CH == 0, "row = whole chunks");
  __shared__ float red[MB_WARPS][CH * 128 + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 ag[V], ab[V];
#pragma unroll
  for (int i = 0; i < V; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long stride = static_cast<long long>(gridDim.x) * MB_WARPS;
  for (long long row = static_cast<long long>(blockIdx.x) * MB_WARPS + warp; row < rows; row += stride) {
    const long long mine = lane < 8 ? merge_src_row(geo, row, lane) : -1;
    const float mu = mean[row], rs = rstd[row];
    float bscale = 1.f;
    if (dx_bf16 != nullptr && row_scale != nullptr)
      bscale = row_scale[static_cast<unsigned>(row) / static_cast<unsigned>(geo.oD * geo.oH * geo.oW)];
    const bf16* dyr = dy + row * C8;
    float4 xv[CH], d[CH];
    int off[CH];                  // float4 index of this lane's element in x / dx, or -1 beyond the real grid
    float s1 = 0.f, s2 = 0.f;
    auto load = [&](int ch) {
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int j = 128 * (ch * CH + k) + 4 * lane, q = j / C, cc = j - q * C;
        const long long src = __shfl_sync(0xffffffffu, mine, q);
        off[k] = src >= 0 ? static_cast<int>(src * (C / 4)) + (cc >> 2) : -1;
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int j = 128 * (ch * CH + k) + 4 * lane;
        xv[k] = off[k] >= 0 ? reinterpret_cast<const float4*>(x)[off[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
        const uint2 u = *reinterpret_cast<const uint2*>(dyr + j);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        d[k] = make_float4(a.x, a.y, b.x, b.y);
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        xv[k].x = (xv[k].x - mu) * rs; xv[k].y = (xv[k].y - mu) * rs;      // xhat
        xv[k].z = (xv[k].z - mu) * rs; xv[k].w = (xv[k].w - mu) * rs;
      }
    };
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      load(ch);
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int i = ch * CH + k;
        ab[i].x += d[k].x; ab[i].y += d[k].y; ab[i].z += d[k].z; ab[i].w += d[k].w;
        ag[i].x = fmaf(d[k].x, xv[k].x, ag[i].x); ag[i].y = fmaf(d[k].y, xv[k].y, ag[i].y);
        ag[i].z = fmaf(d[k].z, xv[k].z, ag[i].z); ag[i].w = fmaf(d[k].w, xv[k].w, ag[i].w);
        const float4 g = *reinterpret_cast<const float4*>(gamma + 128 * i + 4 * lane);
        d[k].x *= g.x; d[k].y *= g.y; d[k].z *= g.z; d[k].w *= g.w;
        s1 += (d[k].x + d[k].y) + (d[k].z + d[k].w);
        s2 += (d[k].x * xv[k].x + d[k].y * xv[k].y) + (d[k].z * xv[k].z + d[k].w * xv[k].w);
      }
    }
    s1 = warp_sum(s1) * (1.0f / C8);
    s2 = warp_sum(s2) * (1.0f / C8);    // mean(dy*g*xhat)
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      if (NCH > 1) {
        load(ch);
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          const float4 g = *reinterpret_cast<const float4*>(gamma + 128 * (ch * CH + k) + 4 * lane);
          d[k].x *= g.x; d[k].y *= g.y; d[k].z *= g.z; d[k].w *= g.w;
        }
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        if (off[k] >= 0) {
          const float4 o = make_float4(rs * (d[k].x - s1 - xv[k].x * s2), rs * (d[k].y - s1 - xv[k].y * s2),
                                       rs * (d[k].z - s1 - xv[k].z * s2), rs * (d[k].w - s1 - xv[k].w * s2));
          reinterpret_cast<float4*>(dx)[off[k]] = o;
          if (dx_bf16 != nullptr) {       // the copy the previous block's backward starts from (its DropPath factor)
            uint2 u;
            u.x = pack_bf16(o.x * bscale, o.y * bscale);
            u.y = pack_bf16(o.z * bscale, o.w * bscale);
            reinterpret_cast<uint2*>(dx_bf16)[off[k]] = u;
          }
        }
      }
    }
  }
  // fold the warps of the block chunk by chunk (smem), then one atomic per column and block
#pragma unroll
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
      for (int k = 0; k < CH; ++k)
        *reinterpret_cast<float4*>(&red[warp][128 * k + 4 * lane]) = which ? ab[ch * CH + k] : ag[ch * CH + k];
      __syncthreads();
      for (int col = threadIdx.x; col < CH * 128; col += MB_WARPS * 32) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < MB_WARPS; ++w) t += red[w][col];
        atomicAdd((which ? dbeta : dgamma) + ch * CH * 128 + col, t);
      }
      __syncthreads();
    }
  }
}

struct LnShape { int lpr, v; };
__host__ inline LnShape ln_fast_shape(int C) {
  static const int table[][2] = {{8, 2}, {8, 3}, {8, 4}, {16, 3}, {16, 4}, {32, 3}, {32, 4}, {32, 6}};
  for (const auto& t : table)
    if (t[0] * t[1] * 4 == C) return {t[0], t[1]};
  return {0, 0};
}

// Column reductions over rows: dbeta[c] += sum_r dy[r,c]; dgamma[c] += sum_r dy[r,c]*xhat[r,c].
// With x == nullptr it is a plain column sum (bias gradients of the Linear layers).
constexpr int CR_TX = 32, CR_TY = 8, CR_ROWS = 256;
template <typename DyT>
__global__ void __launch_bounds__(CR_TX * CR_TY) colreduce_kernel(const DyT* __restrict__ dy, long long lddy,
                                                                  const float* __restrict__ x, long long ldx,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd,
                                                                  float* __restrict__ dgamma,
                                                                  float* __restrict__ dbeta, long long rows, int C) {
  __shared__ float sg[CR_TY][CR_TX + 1], sb[CR_TY][CR_TX + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * CR_TX + tx;
  const long long r0 = static_cast<long long>(blockIdx.y) * CR_ROWS;
  const long long r1 = r0 + CR_ROWS < rows ? r0 + CR_ROWS : rows;
  float ag = 0.f, ab = 0.f;
  if (c < C) {
    for (long long r = r0 + ty; r < r1; r += CR_TY) {
      const float d = to_f<DyT>(dy[r * lddy + c]);
      ab += d;
      if (x != nullptr) ag += d * (x[r * ldx + c] - mean[r]) * rstd[r];
    }
  }
  sg[ty][tx] = ag;
  sb[ty][tx] = ab;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < CR_TY; ++i) {
      ag += sg[i][tx];
      ab += sb[i][tx];
    }
    if (dbeta != nullptr) atomicAdd(dbeta + c, ab);
    if (dgamma != nullptr && x != nullptr) atomicAdd(dgamma + c, ag);
  }
}

}  // namespace

extern "C" int vsn_colreduce(const void* dy, long long lddy, int dy_bf16, const float* x, long long ldx,
                             const float* mean, const float* rstd, float* dgamma, float* dbeta, long long rows, int C,
                             void* stream);

extern "C" int vsn_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, void* y,
                                 long long ldy, int y_bf16, float* mean, float* rstd, long long rows, int C,
                                 float eps, void* stream) {
  VSN_CHECK(C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "vsn_layernorm_fwd: C/ld must be multiples of 4 (C=%d)", C);
  if (rows == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const LnShape fs = ln_fast_shape(C);
  if (fs.lpr != 0) {
    const unsigned g2 = static_cast<unsigned>(ceil_div_ll(rows, LN_WARPS * (32 / fs.lpr)));
#define VSN_LN_FWD(L, VV)                                                                                         \
    if (fs.lpr == L && fs.v == VV) {                                                                                \
      if (y_bf16) ln_fwd_fast_kernel<bf16, L, VV><<<g2, LN_WARPS * 32, 0, s>>>(x, ldx, gamma, beta,                 \
                      reinterpret_cast<bf16*>(y), ldy, mean, rstd, rows, eps);                                      \
      else ln_fwd_fast_kernel<float, L, VV><<<g2, LN_WARPS * 32, 0, s>>>(x, ldx, gamma, beta,                       \
                      reinterpret_cast<float*>(y), ldy, mean, rstd, rows, eps);                                     \
    }
    VSN_LN_FWD(8, 2) VSN_LN_FWD(8, 3) VSN_LN_FWD(8, 4) VSN_LN_FWD(16, 3) VSN_LN_FWD(16, 4)
    VSN_LN_FWD(32, 3) VSN_LN_FWD(32, 4) VSN_LN_FWD(32, 6)
#undef VSN_LN_FWD
    VSN_LAUNCH_CHECK();
    return 0;
  }
  const unsigned grid = static_cast<unsigned>(ceil_div_ll(rows, LN_WARPS));
  if (y_bf16)
    ln_fwd_kernel<bf16><<<grid, LN_WARPS * 32, 0, s>>>(x, ldx, gamma, beta, reinterpret_cast<bf16*>(y), ldy, mean,
                                                         rstd, rows, C, eps);
  else
    ln_fwd_kernel<float><<<grid, LN_WARPS * 32, 0, s>>>(x, ldx, gamma, beta, reinterpret_cast<float*>(y), ldy, mean,
                                                          rstd, rows, C, eps);
  VSN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vsn_layernorm_bwd(const void* dy, long long lddy, int dy_bf16, const float* x, long long ldx,
                                 const float* mean, const float* rstd, const float* gamma, const float* resid_grad,
                                 long long ldr, float* dx, long long lddx, void* dx_bf16, long long ldb,
                                 const float* row_scale, int rows_per_group, float* dgamma, float* dbeta,
                                 long long rows, int C, void* stream) {
  VSN_CHECK(C % 4 == 0, "vsn_layernorm_bwd: C must be a multiple of 4 (C=%d)", C);
  VSN_CHECK((dgamma == nullptr) == (dbeta == nullptr), "vsn_layernorm_bwd: dgamma and dbeta go together");
  if (rows == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int rpg = rows_per_group > 0 ? rows_per_group : 1;
  const LnShape fs = ln_fast_shape(C);
  if (fs.lpr != 0 && lddy % 4 == 0 && ldx % 4 == 0) {
    const long long need = ceil_div_ll(rows, LN_WARPS * (32 / fs.lpr));
    // every block ends with one atomic per column for dgamma / dbeta: short problems (the late stages) take fewer,
    // longer blocks so that those atomics do not outweigh the row work
    static int cap_small = -1;
    if (cap_small < 0) { const char* e = getenv("VSN_LN_BWD_CAP"); cap_small = e ? atoi(e) : 2; }
    const long long cap = (rows * C < (1LL << 25) ? cap_small : 4) * static_cast<long long>(vsn_num_sms());
    const unsigned g2 = static_cast<unsigned>(need < cap ? need : cap);
#define VSN_LN_BWD(L, VV)                                                                                          \
    if (fs.lpr == L && fs.v == VV) {                                                                                 \
      if (dy_bf16) ln_bwd_fast_kernel<bf16, L, VV><<<g2, LN_WARPS * 32, 0, s>>>(reinterpret_cast<const bf16*>(dy), lddy, \
                      x, ldx, mean, rstd, gamma, resid_grad, ldr, dx, lddx, reinterpret_cast<bf16*>(dx_bf16), ldb,   \
                      row_scale, rpg, rows, dgamma, dbeta);                                                          \
      else ln_bwd_fast_kernel<float, L, VV><<<g2, LN_WARPS * 32, 0, s>>>(reinterpret_cast<const float*>(dy), lddy,  \
                      x, ldx, mean, rstd, gamma, resid_grad, ldr, dx, lddx, reinterpret_cast<bf16*>(dx_bf16), ldb,   \
                      row_scale, rpg, rows, dgamma, dbeta);                                                          \
    }
    VSN_LN_BWD(8, 2) VSN_LN_BWD(8, 3) VSN_LN_BWD(8, 4) VSN_LN_BWD(16, 3) VSN_LN_BWD(16, 4)
    VSN_LN_BWD(32, 3) VSN_LN_BWD(32, 4) VSN_LN_BWD(32, 6)
#undef VSN_LN_BWD
    VSN_LAUNCH_CHECK();
    return 0;
  }
  if (dgamma != nullptr) {   // generic shapes: separate column reduction, then the row kernel
    if (int rc = vsn_colreduce(dy, lddy, dy_bf16, x, ldx, mean, rstd, dgamma, dbeta, rows, C, stream)) return rc;
  }
  const unsigned grid = static_cast<unsigned>(ceil_div_ll(rows, LN_WARPS));
  if (dy_bf16)
    ln_bwd_kernel<bf16><<<grid, LN_WARPS * 32, 0, s>>>(reinterpret_cast<const bf16*>(dy), lddy, x, ldx, mean, rstd,
                                                         gamma, resid_grad, ldr, dx, lddx,
                                                         reinterpret_cast<bf16*>(dx_bf16), ldb, row_scale, rpg, rows, C);
  else
    ln_bwd_kernel<float><<<grid, LN_WARPS * 32, 0, s>>>(reinterpret_cast<const float*>(dy), lddy, x, ldx, mean,
                                                          rstd, gamma, resid_grad, ldr, dx, lddx,
                                                          reinterpret_cast<bf16*>(dx_bf16), ldb, row_scale, rpg, rows,
                                                          C);
  VSN_LAUNCH_CHECK();
  return 0;
}

// dgamma/dbeta (x != null) or a plain column sum into dbeta (x == null).  Accumulates (+=).
extern "C" int vsn_colreduce(const void* dy, long long lddy, int dy_bf16, const float* x, long long ldx,
                             const float* mean, const float* rstd, float* dgamma, float* dbeta, long long rows, int C,
                             void* stream) {
  if (rows == 0 || C == 0) return 0;
  dim3 grid(ceil_div(C, CR_TX), static_cast<unsigned>(ceil_div_ll(rows, CR_ROWS)));
  VSN_CHECK(grid.y <= 65535, "vsn_colreduce: too many rows (%lld)", rows);
  dim3 block(CR_TX, CR_TY);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dy_bf16)
    colreduce_kernel<bf16><<<grid, block, 0, s>>>(reinterpret_cast<const bf16*>(dy), lddy, x, ldx, mean, rstd, dgamma,
                                                    dbeta, rows, C);
  else
    colreduce_kernel<float><<<grid, block, 0, s>>>(reinterpret_cast<const float*>(dy), lddy, x, ldx, mean, rstd,
                                                     dgamma, dbeta, rows, C);
  VSN_LAUNCH_CHECK();
  return 0;
}

static bool merge_ln_geom(int pD, int pH, int pW, int rD, int rH, int rW, MergeGeom& g) {
  g = {pD, pH, pW, rD, rH, rW, (rD + 1) / 2, (rH + 1) / 2, (rW + 1) / 2};
  return rD <= pD && rH <= pH && rW <= pW && rD > 0 && rH > 0 && rW > 0;
}

// PatchMerging gather + LayerNorm(8C) in one pass: y[(b,d,h,w), :] = LN(concat of the 2x2x2 neighbours of x) as the
// bf16 operand of the reduction GEMM; mean / rstd [rows] for the backward.  C = 96 or 192 (the stages whose merged row
// a warp holds in registers); other widths use vsn_merge_gather + vsn_layernorm_fwd.
extern "C" int vsn_merge_ln_fwd(const float* x, int pD, int pH, int pW, int rD, int rH, int rW, int B, int C,
                                const float* gamma, const float* beta, void* y, float* mean, float* rstd, float eps,
                                void* stream) {
  MergeGeom g;
  VSN_CHECK(merge_ln_geom(pD, pH, pW, rD, rH, rW, g), "vsn_merge_ln_fwd: bad grid (%d,%d,%d) / (%d,%d,%d)", pD, pH, pW,
            rD, rH, rW);
  VSN_CHECK(C == 96 || C == 192, "vsn_merge_ln_fwd: C must be 96 or 192 (got %d)", C);
  const long long rows = static_cast<long long>(B) * g.oD * g.oH * g.oW;
  if (rows == 0) return 0;
  VSN_CHECK(rows < (1LL << 31), "vsn_merge_ln_fwd: too many merged rows (%lld)", rows);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>(ceil_div_ll(rows, LN_WARPS));
  if (C == 96) ln_merge_fwd_kernel<96><<<grid, LN_WARPS * 32, 0, s>>>(x, g, gamma, beta, reinterpret_cast<bf16*>(y), mean, rstd, rows, eps);
  else ln_merge_fwd_kernel<192><<<grid, LN_WARPS * 32, 0, s>>>(x, g, gamma, beta, reinterpret_cast<bf16*>(y), mean, rstd, rows, eps);
  VSN_LAUNCH_CHECK();
  return 0;
}

// Backward of the above: dx[source token, :] = LayerNorm backward of dy [rows, 8C] (bf16), written to the source tokens
// on the padded stage grid (tokens outside the real grid are NOT written: the caller zeroes dx when pD,pH,pW differ
// from rD,rH,rW); dgamma / dbeta [8C] accumulate (+=).  dx_bf16 (optional, same shape): dx * row_scale[b] as the 16-bit
// operand the previous block's backward starts from (row_scale = its MLP DropPath factor per sample, or null).
extern "C" int vsn_merge_ln_bwd(const void* dy, const float* x, int pD, int pH, int pW, int rD, int rH, int rW, int B,
                                int C, const float* mean, const float* rstd, const float* gamma, float* dx,
                                float* dgamma, float* dbeta, void* dx_bf16, const float* row_scale, void* stream) {
  MergeGeom g;
  VSN_CHECK(merge_ln_geom(pD, pH, pW, rD, rH, rW, g), "vsn_merge_ln_bwd: bad grid (%d,%d,%d) / (%d,%d,%d)", pD, pH, pW,
            rD, rH, rW);
  VSN_CHECK(C == 96 || C == 192, "vsn_merge_ln_bwd: C must be 96 or 192 (got %d)", C);
  VSN_CHECK(dgamma != nullptr && dbeta != nullptr, "vsn_merge_ln_bwd: dgamma / dbeta are required");
  const long long rows = static_cast<long long>(B) * g.oD * g.oH * g.oW;
  if (rows == 0) return 0;
  VSN_CHECK(rows < (1LL << 31) && static_cast<long long>(B) * pD * pH * pW * (C / 4) < (1LL << 31),
            "vsn_merge_ln_bwd: grid too large for 32-bit float4 indices (%lld merged rows)", rows);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long need = ceil_div_ll(rows, MB_WARPS);
  // persistent blocks (each ends with one atomic per column): a few resident waves
  const long long cap = (rows * 8LL * C < (1LL << 25) ? 4 : 8) * static_cast<long long>(vsn_num_sms());
  const unsigned grid = static_cast<unsigned>(need < cap ? need : cap);
  if (C == 96) ln_merge_bwd_kernel<96><<<grid, MB_WARPS * 32, 0, s>>>(reinterpret_cast<const bf16*>(dy), x, g, mean, rstd, gamma, dx, rows, dgamma, dbeta, reinterpret_cast<bf16*>(dx_bf16), row_scale);
  else ln_merge_bwd_kernel<192><<<grid, MB_WARPS * 32, 0, s>>>(reinterpret_cast<const bf16*>(dy), x, g, mean, rstd, gamma, dx, rows, dgamma, dbeta, reinterpret_cast<bf16*>(dx_bf16), row_scale);
  VSN_LAUNCH_CHECK();
  return 0;
}
