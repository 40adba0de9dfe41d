// Fused attention core, forward and backward, for both block types of the path:
//   * Swin-3D shifted-window attention (models/swin_transformer_3d.py:162-199 + the roll / window_partition /
//     window_reverse / mask plumbing at :330-358,463-492).  The cyclic shift, the window gather/scatter, the
//     relative-position-bias lookup and the {0,-100} shift mask are all index math inside the kernel: no rolled
//     copy, no [nW,N,N] mask tensor and no [h,N,N] bias tensor ever exists in HBM, and the N x N logits never
//     leave the SM.
//   * ViT-3D global attention (models/vit_3d.py:129-142): same kernels in "dense" mode.
// Flash-style streaming softmax; bf16 operands, fp32 accumulation and fp32 softmax statistics.
// Round-1 implementation uses warp-level mma.sync tiles (m16n8k16); see DESIGN.md for the tcgen05 plan.
//
// Layouts: qkv is [T, 3C] bf16 with q|k|v column blocks and head-major [heads][hd] inside each (the reshape at
// :166-170 / chunk(3) + 'b n (h d)' in the ViT); out is [T, C] bf16; rows are tokens of the (padded) stage grid
// in natural (b,d,h,w) order, so neither input nor output is ever re-ordered.
#include <stdlib.h>
#include "common.cuh"
#include "wattn_tc.cuh"
#include "dattn_tc.cuh"

namespace {

constexpr int TQ = 64;        // rows owned by one CTA (4 warps x 16)
constexpr int TK = 64;        // streamed tile
constexpr int MAX_WIN_TOK = 384;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float NEG_BIG = -1.0e30f;

struct AttnParams {
  const bf16* qkv;      // [T, 3C]
  bf16* out;            // fwd: [T, C]
  float* lse;           // [S, heads, Npad]   natural-log logsumexp of each row
  // backward
  const bf16* dout;     // [T, C]
  const float* delta;   // [S, heads, Npad]   rowsum(dO * O)
  bf16* dqkv;           // [T, 3C]
  float* dbias_dense;   // [heads, Npad, Npad] fp32 accumulator (window mode) or null
  const float* table;   // [table_len, heads] fp32 relative_position_bias_table or null
  int table_len;
  int S, N, Npad, heads, C;
  float scale;
  // window geometry (window mode only)
  int B, Dp, Hp, Wp, wd, wh, ww, sd, sh, sw, nWd, nWh, nWw, use_mask;
  int groups;           // backward-dq: sequences are strided over blockIdx.z = group
};

// ---- tiny PTX helpers -------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." VSN_T16 "." VSN_T16 ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Swizzled [64][HD] bf16 tile: 16-byte chunk c of row r lives at chunk (c ^ f(r)); conflict-free for ldmatrix.
template <int HD>
__device__ __forceinline__ uint32_t tile_off(int r, int chunk) {
  if constexpr (HD == 32) return r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4);
  else return r * 128 + ((chunk ^ (r & 7)) << 4);
}

// ---- token geometry -----------------------------------------------------------------
template <bool WIN>
struct SeqMap {
  int* row;            // smem [MAX_WIN_TOK] global row of local token (or -1)      (WIN only)
  short* lin;          // smem relative-position linear code d*(2wh-1)(2ww-1)+h*(2ww-1)+w  (WIN only)
  unsigned char* reg;  // smem region id 0..26 of the token on the rolled grid          (WIN only)
  long long base;      // dense: first row of the sequence
  int N;
  __device__ __forceinline__ long long grow(int i) const {
    if constexpr (WIN) return row[i];
    else return i < N ? base + i : -1;
  }
};

// Fill the per-sequence tables (window mode).  All threads of the CTA participate; caller syncs.
__device__ void fill_window_tables(const AttnParams& p, int s, int* row, short* lin, unsigned char* reg) {
  const int nW = p.nWd * p.nWh * p.nWw;
  const int b = s / nW, wi = s % nW;
  const int wa = wi / (p.nWh * p.nWw), wb = (wi / p.nWw) % p.nWh, wc = wi % p.nWw;
  for (int i = threadIdx.x; i < p.Npad; i += blockDim.x) {
    if (i < p.N) {
      const int ld = i / (p.wh * p.ww), lh = (i / p.ww) % p.wh, lw = i % p.ww;
      const int dd = wa * p.wd + ld, hh = wb * p.wh + lh, wv = wc * p.ww + lw;   // position on the rolled grid
      int d0 = dd + p.sd; if (d0 >= p.Dp) d0 -= p.Dp;                            // torch.roll(-shift) source
      int h0 = hh + p.sh; if (h0 >= p.Hp) h0 -= p.Hp;
      int w0 = wv + p.sw; if (w0 >= p.Wp) w0 -= p.Wp;
      row[i] = ((b * p.Dp + d0) * p.Hp + h0) * p.Wp + w0;
      lin[i] = static_cast<short>(ld * ((2 * p.wh - 1) * (2 * p.ww - 1)) + lh * (2 * p.ww - 1) + lw);
      const int rd = dd < p.Dp - p.wd ? 0 : (dd < p.Dp - p.sd ? 1 : 2);
      const int rh = hh < p.Hp - p.wh ? 0 : (hh < p.Hp - p.sh ? 1 : 2);
      const int rw = wv < p.Wp - p.ww ? 0 : (wv < p.Wp - p.sw ? 1 : 2);
      reg[i] = static_cast<unsigned char>(9 * rd + 3 * rh + rw);
    } else {
      row[i] = -1; lin[i] = 0; reg[i] = 0;
    }
  }
}

// Async copy of a [64][HD] tile (rows t0.. of the sequence, column block `col`) into swizzled smem.
template <int HD, bool WIN>
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, long long ld, int col, const SeqMap<WIN>& m,
                                          int t0) {
  constexpr int CH = HD / 8;
  for (int idx = threadIdx.x; idx < TK * CH; idx += blockDim.x) {
    const int r = idx / CH, c = idx % CH;
    const long long gr = m.grow(t0 + r);
    const bf16* g = src + (gr < 0 ? 0 : gr) * ld + col + c * 8;
    cp_async16(reinterpret_cast<uint8_t*>(dst) + tile_off<HD>(r, c), g, gr >= 0);
  }
}

// logit (log2 domain) of query i vs key j given raw dot product
template <bool WIN>
__device__ __forceinline__ float logit2(float dot, int i, int j, const AttnParams& p, const SeqMap<WIN>& m,
                                        const float* tbl, int off) {
  if (j >= p.N) return NEG_BIG;
  float s = dot * p.scale;
  if constexpr (WIN) {
    if (tbl != nullptr && i < p.N) s += tbl[m.lin[i] - m.lin[j] + off];
    if (p.use_mask && m.reg[i] != m.reg[j]) s += -100.0f;
  }
  return s * LOG2E;
}

template <int HD>
struct Smem {
  static constexpr int TILE = TQ * HD * 2;   // bytes
};

// ===================================== forward =========================================
template <int HD, bool WIN>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int TILE = Smem<HD>::TILE;
  bf16* sQ = reinterpret_cast<bf16*>(smem);
  bf16* sK = reinterpret_cast<bf16*>(smem + TILE);          // 2 buffers
  bf16* sV = reinterpret_cast<bf16*>(smem + 3 * TILE);      // 2 buffers
  uint8_t* extra = smem + 5 * TILE;
  SeqMap<WIN> m;
  float* tbl = nullptr;
  const int a = blockIdx.y, s = blockIdx.z, q0 = blockIdx.x * TQ;
  m.N = p.N;
  m.base = static_cast<long long>(s) * p.N;
  if constexpr (WIN) {
    m.row = reinterpret_cast<int*>(extra);
    m.lin = reinterpret_cast<short*>(extra + MAX_WIN_TOK * 4);
    m.reg = reinterpret_cast<unsigned char*>(extra + MAX_WIN_TOK * 6);
    fill_window_tables(p, s, m.row, m.lin, m.reg);
    if (p.table != nullptr) {
      tbl = reinterpret_cast<float*>(extra + MAX_WIN_TOK * 8);
      for (int i = threadIdx.x; i < p.table_len; i += blockDim.x) tbl[i] = p.table[static_cast<long long>(i) * p.heads + a];
    }
    __syncthreads();
  }
  const int off = WIN ? ((p.wd - 1) * (2 * p.wh - 1) * (2 * p.ww - 1) + (p.wh - 1) * (2 * p.ww - 1) + (p.ww - 1)) : 0;
  const long long ld = 3LL * p.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = p.Npad / TK;

  load_tile<HD, WIN>(sQ, p.qkv, ld, a * HD, m, q0);
  load_tile<HD, WIN>(sK, p.qkv, ld, p.C + a * HD, m, 0);
  load_tile<HD, WIN>(sV, p.qkv, ld, 2 * p.C + a * HD, m, 0);
  cp_async_commit();

  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mrow[2] = {NEG_BIG, NEG_BIG}, lrow[2] = {0.f, 0.f};
  uint32_t qa[HD / 16][4];
  const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;   // sequence-local query rows of this thread

  for (int kt = 0; kt < ntiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ntiles) {
      load_tile<HD, WIN>(sK + (buf ^ 1) * TQ * HD, p.qkv, ld, p.C + a * HD, m, (kt + 1) * TK);
      load_tile<HD, WIN>(sV + (buf ^ 1) * TQ * HD, p.qkv, ld, 2 * p.C + a * HD, m, (kt + 1) * TK);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4(s_u32(sQ) + tile_off<HD>(r, kk * 2 + (lane >> 4)), qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3]);
      }
    }
    const uint32_t kb = s_u32(sK + buf * TQ * HD), vb = s_u32(sV + buf * TQ * HD);
    float sc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int n = np * 16 + (lane & 7) + (lane >> 4) * 8;
        ldsm_x4(kb + tile_off<HD>(n, kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
        mma16816(sc[np * 2], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b0, b1);
        mma16816(sc[np * 2 + 1], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b2, b3);
      }
    }
    // logits + streaming softmax
    float mx[2] = {NEG_BIG, NEG_BIG};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int j = kt * TK + nt * 8 + 2 * t;
      sc[nt][0] = logit2<WIN>(sc[nt][0], r_lo, j, p, m, tbl, off);
      sc[nt][1] = logit2<WIN>(sc[nt][1], r_lo, j + 1, p, m, tbl, off);
      sc[nt][2] = logit2<WIN>(sc[nt][2], r_hi, j, p, m, tbl, off);
      sc[nt][3] = logit2<WIN>(sc[nt][3], r_hi, j + 1, p, m, tbl, off);
      mx[0] = fmaxf(mx[0], fmaxf(sc[nt][0], sc[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(sc[nt][2], sc[nt][3]));
    }
    float corr[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      const float mn = fmaxf(mrow[h], mx[h]);
      corr[h] = exp2f(mrow[h] - mn);
      mrow[h] = mn;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(sc[nt][0] - mrow[0]), p1 = exp2f(sc[nt][1] - mrow[0]);
      const float p2 = exp2f(sc[nt][2] - mrow[1]), p3 = exp2f(sc[nt][3] - mrow[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) lrow[h] = lrow[h] * corr[h] + rs[h];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {           // 16 keys per step
#pragma unroll
      for (int np = 0; np < HD / 16; ++np) {   // 16 head-dim columns per ldmatrix
        uint32_t b0, b1, b2, b3;
        const int kr = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(vb + tile_off<HD>(kr, np * 2 + (lane >> 4)), b0, b1, b2, b3);
        mma16816(o[np * 2], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b0, b1);
        mma16816(o[np * 2 + 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b2, b3);
      }
    }
    __syncthreads();   // everyone is done with this buffer before it is refilled
  }
  // finalize
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    lrow[h] += __shfl_xor_sync(0xffffffffu, lrow[h], 1);
    lrow[h] += __shfl_xor_sync(0xffffffffu, lrow[h], 2);
  }
  const int rr[2] = {r_lo, r_hi};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long gr = rr[h] < p.Npad ? m.grow(rr[h]) : -1;
    const float inv = 1.f / lrow[h];
    if (gr >= 0) {
      bf16* op = p.out + gr * p.C + a * HD + 2 * t;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i)
        *reinterpret_cast<uint32_t*>(op + i * 8) = pack_bf16(o[i][h * 2] * inv, o[i][h * 2 + 1] * inv);
    }
    if (t == 0 && p.lse != nullptr && rr[h] < p.Npad)
      p.lse[(static_cast<long long>(s) * p.heads + a) * p.Npad + rr[h]] = (mrow[h] + log2f(lrow[h])) / LOG2E;
  }
}

// ============================ backward: delta = rowsum(dO * O) ==========================
// One warp per (token row, head).  Stored in the same [S, heads, Npad] layout as lse.
template <bool WIN>
__global__ void __launch_bounds__(128) attn_delta_kernel(const AttnParams p, const bf16* __restrict__ o, float* __restrict__ delta) {
  __shared__ int s_row[WIN ? MAX_WIN_TOK : 1];
  __shared__ short s_lin[WIN ? MAX_WIN_TOK : 1];
  __shared__ unsigned char s_reg[WIN ? MAX_WIN_TOK : 1];
  const int s = blockIdx.x;
  SeqMap<WIN> m;
  m.N = p.N;
  m.base = static_cast<long long>(s) * p.N;
  if constexpr (WIN) {
    m.row = s_row; m.lin = s_lin; m.reg = s_reg;
    fill_window_tables(p, s, s_row, s_lin, s_reg);
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd = p.C / p.heads;
  for (int item = warp; item < p.Npad * p.heads; item += 4) {
    const int i = item % p.Npad, a = item / p.Npad;
    const long long gr = m.grow(i);
    float acc = 0.f;
    if (gr >= 0) {
      for (int e = lane * 2; e < hd; e += 64) {
        const float2 x = unpack_bf16(*reinterpret_cast<const uint32_t*>(o + gr * p.C + a * hd + e));
        const float2 y = unpack_bf16(*reinterpret_cast<const uint32_t*>(p.dout + gr * p.C + a * hd + e));
        acc += x.x * y.x + x.y * y.y;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) delta[(static_cast<long long>(s) * p.heads + a) * p.Npad + i] = acc;
  }
}

// ================================ backward: dQ (+ dBias) =================================
// CTA owns 64 query rows of one head and walks the key tiles; with a bias table it also accumulates
// dS into a CTA-private [64, Npad] fp32 strip over all the sequences of its group, flushed once.
template <int HD, bool WIN>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int TILE = Smem<HD>::TILE;
  bf16* sQ = reinterpret_cast<bf16*>(smem);
  bf16* sdO = reinterpret_cast<bf16*>(smem + TILE);
  bf16* sK = reinterpret_cast<bf16*>(smem + 2 * TILE);   // 2 buffers
  bf16* sV = reinterpret_cast<bf16*>(smem + 4 * TILE);   // 2 buffers
  uint8_t* extra = smem + 6 * TILE;
  SeqMap<WIN> m;
  float* tbl = nullptr;
  float* acc = nullptr;   // [TQ][Npad] dS accumulator
  const int a = blockIdx.y, q0 = blockIdx.x * TQ;
  m.N = p.N;
  if constexpr (WIN) {
    m.row = reinterpret_cast<int*>(extra);
    m.lin = reinterpret_cast<short*>(extra + MAX_WIN_TOK * 4);
    m.reg = reinterpret_cast<unsigned char*>(extra + MAX_WIN_TOK * 6);
    if (p.table != nullptr) {
      tbl = reinterpret_cast<float*>(extra + MAX_WIN_TOK * 8);
      acc = tbl + ((p.table_len + 3) & ~3);
      for (int i = threadIdx.x; i < p.table_len; i += blockDim.x) tbl[i] = p.table[static_cast<long long>(i) * p.heads + a];
      for (int i = threadIdx.x; i < TQ * (p.Npad + 4); i += blockDim.x) acc[i] = 0.f;
    }
  }
  const int off = WIN ? ((p.wd - 1) * (2 * p.wh - 1) * (2 * p.ww - 1) + (p.wh - 1) * (2 * p.ww - 1) + (p.ww - 1)) : 0;
  const long long ld = 3LL * p.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = p.Npad / TK;
  const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;

  for (int s = blockIdx.z; s < p.S; s += p.groups) {
    __syncthreads();
    m.base = static_cast<long long>(s) * p.N;
    if constexpr (WIN) {
      fill_window_tables(p, s, m.row, m.lin, m.reg);
      __syncthreads();
    }
    load_tile<HD, WIN>(sQ, p.qkv, ld, a * HD, m, q0);
    load_tile<HD, WIN>(sdO, p.dout, p.C, a * HD, m, q0);
    load_tile<HD, WIN>(sK, p.qkv, ld, p.C + a * HD, m, 0);
    load_tile<HD, WIN>(sV, p.qkv, ld, 2 * p.C + a * HD, m, 0);
    cp_async_commit();
    const long long stat = (static_cast<long long>(s) * p.heads + a) * p.Npad;
    const float lse_lo = p.lse[stat + r_lo] * LOG2E, lse_hi = p.lse[stat + r_hi] * LOG2E;
    const float dl_lo = p.delta[stat + r_lo], dl_hi = p.delta[stat + r_hi];
    float dq[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    uint32_t qa[HD / 16][4], da[HD / 16][4];

    for (int kt = 0; kt < ntiles; ++kt) {
      const int buf = kt & 1;
      if (kt + 1 < ntiles) {
        load_tile<HD, WIN>(sK + (buf ^ 1) * TQ * HD, p.qkv, ld, p.C + a * HD, m, (kt + 1) * TK);
        load_tile<HD, WIN>(sV + (buf ^ 1) * TQ * HD, p.qkv, ld, 2 * p.C + a * HD, m, (kt + 1) * TK);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      if (kt == 0) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          ldsm_x4(s_u32(sQ) + tile_off<HD>(r, kk * 2 + (lane >> 4)), qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3]);
          ldsm_x4(s_u32(sdO) + tile_off<HD>(r, kk * 2 + (lane >> 4)), da[kk][0], da[kk][1], da[kk][2], da[kk][3]);
        }
      }
      const uint32_t kb = s_u32(sK + buf * TQ * HD), vb = s_u32(sV + buf * TQ * HD);
      float sc[8][4], dp[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;
          const int n = np * 16 + (lane & 7) + (lane >> 4) * 8;
          const uint32_t o2 = tile_off<HD>(n, kk * 2 + ((lane >> 3) & 1));
          ldsm_x4(kb + o2, b0, b1, b2, b3);
          mma16816(sc[np * 2], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b0, b1);
          mma16816(sc[np * 2 + 1], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b2, b3);
          ldsm_x4(vb + o2, b0, b1, b2, b3);
          mma16816(dp[np * 2], da[kk][0], da[kk][1], da[kk][2], da[kk][3], b0, b1);
          mma16816(dp[np * 2 + 1], da[kk][0], da[kk][1], da[kk][2], da[kk][3], b2, b3);
        }
      }
      uint32_t sa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int j = kt * TK + nt * 8 + 2 * t;
        const float p0 = exp2f(logit2<WIN>(sc[nt][0], r_lo, j, p, m, tbl, off) - lse_lo);
        const float p1 = exp2f(logit2<WIN>(sc[nt][1], r_lo, j + 1, p, m, tbl, off) - lse_lo);
        const float p2 = exp2f(logit2<WIN>(sc[nt][2], r_hi, j, p, m, tbl, off) - lse_hi);
        const float p3 = exp2f(logit2<WIN>(sc[nt][3], r_hi, j + 1, p, m, tbl, off) - lse_hi);
        const float d0 = p0 * (dp[nt][0] - dl_lo), d1 = p1 * (dp[nt][1] - dl_lo);
        const float d2 = p2 * (dp[nt][2] - dl_hi), d3 = p3 * (dp[nt][3] - dl_hi);
        if (acc != nullptr) {
          // row stride Npad+4 floats: the 8 rows x 4 column pairs of a warp hit 32 distinct banks
          float2* alo = reinterpret_cast<float2*>(acc + (warp * 16 + g) * (p.Npad + 4) + j);
          float2* ahi = reinterpret_cast<float2*>(acc + (warp * 16 + g + 8) * (p.Npad + 4) + j);
          float2 v = *alo; v.x += d0; v.y += d1; *alo = v;
          v = *ahi; v.x += d2; v.y += d3; *ahi = v;
        }
        sa[nt >> 1][(nt & 1) * 2] = pack_bf16(d0, d1);
        sa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(d2, d3);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int np = 0; np < HD / 16; ++np) {
          uint32_t b0, b1, b2, b3;
          const int kr = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          ldsm_x4_t(kb + tile_off<HD>(kr, np * 2 + (lane >> 4)), b0, b1, b2, b3);
          mma16816(dq[np * 2], sa[kk][0], sa[kk][1], sa[kk][2], sa[kk][3], b0, b1);
          mma16816(dq[np * 2 + 1], sa[kk][0], sa[kk][1], sa[kk][2], sa[kk][3], b2, b3);
        }
      }
      __syncthreads();
    }
    const int rr[2] = {r_lo, r_hi};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long gr = m.grow(rr[h]);
      if (gr >= 0) {
        bf16* op = p.dqkv + gr * ld + a * HD + 2 * t;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i)
          *reinterpret_cast<uint32_t*>(op + i * 8) = pack_bf16(dq[i][h * 2] * p.scale, dq[i][h * 2 + 1] * p.scale);
      }
    }
  }
  if (acc != nullptr) {
    __syncthreads();
    float* dst = p.dbias_dense + (static_cast<long long>(a) * p.Npad + q0) * p.Npad;
    for (int i = threadIdx.x; i < TQ * p.Npad; i += blockDim.x) {
      const float v = acc[(i / p.Npad) * (p.Npad + 4) + (i % p.Npad)];
      if (v != 0.f) atomicAdd(dst + i, v);
    }
  }
}

// ================================ backward: dK, dV =======================================
// CTA owns 64 key rows of one head and walks the query tiles; everything is computed transposed
// (keys are the MMA M dimension) so the per-key accumulators stay in registers.
template <int HD, bool WIN>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int TILE = Smem<HD>::TILE;
  bf16* sK = reinterpret_cast<bf16*>(smem);
  bf16* sV = reinterpret_cast<bf16*>(smem + TILE);
  bf16* sQ = reinterpret_cast<bf16*>(smem + 2 * TILE);    // 2 buffers
  bf16* sdO = reinterpret_cast<bf16*>(smem + 4 * TILE);   // 2 buffers
  float* sLse = reinterpret_cast<float*>(smem + 6 * TILE);   // [2][64]
  float* sDl = sLse + 2 * TQ;                                // [2][64]
  uint8_t* extra = reinterpret_cast<uint8_t*>(sDl + 2 * TQ);
  SeqMap<WIN> m;
  float* tbl = nullptr;
  const int a = blockIdx.y, s = blockIdx.z, k0 = blockIdx.x * TK;
  m.N = p.N;
  m.base = static_cast<long long>(s) * p.N;
  if constexpr (WIN) {
    m.row = reinterpret_cast<int*>(extra);
    m.lin = reinterpret_cast<short*>(extra + MAX_WIN_TOK * 4);
    m.reg = reinterpret_cast<unsigned char*>(extra + MAX_WIN_TOK * 6);
    fill_window_tables(p, s, m.row, m.lin, m.reg);
    if (p.table != nullptr) {
      tbl = reinterpret_cast<float*>(extra + MAX_WIN_TOK * 8);
      for (int i = threadIdx.x; i < p.table_len; i += blockDim.x) tbl[i] = p.table[static_cast<long long>(i) * p.heads + a];
    }
    __syncthreads();
  }
  const int off = WIN ? ((p.wd - 1) * (2 * p.wh - 1) * (2 * p.ww - 1) + (p.wh - 1) * (2 * p.ww - 1) + (p.ww - 1)) : 0;
  const long long ld = 3LL * p.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = p.Npad / TQ;
  const long long stat = (static_cast<long long>(s) * p.heads + a) * p.Npad;
  const int c_lo = k0 + warp * 16 + g, c_hi = c_lo + 8;   // sequence-local key rows of this thread

  auto load_stats = [&](int buf, int qt) {
    if (threadIdx.x < TQ) {
      sLse[buf * TQ + threadIdx.x] = p.lse[stat + qt * TQ + threadIdx.x] * LOG2E;
      sDl[buf * TQ + threadIdx.x] = p.delta[stat + qt * TQ + threadIdx.x];
    }
  };
  load_tile<HD, WIN>(sK, p.qkv, ld, p.C + a * HD, m, k0);
  load_tile<HD, WIN>(sV, p.qkv, ld, 2 * p.C + a * HD, m, k0);
  load_tile<HD, WIN>(sQ, p.qkv, ld, a * HD, m, 0);
  load_tile<HD, WIN>(sdO, p.dout, p.C, a * HD, m, 0);
  cp_async_commit();
  load_stats(0, 0);

  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  uint32_t ka[HD / 16][4], va[HD / 16][4];

  for (int qt = 0; qt < ntiles; ++qt) {
    const int buf = qt & 1;
    if (qt + 1 < ntiles) {
      load_tile<HD, WIN>(sQ + (buf ^ 1) * TQ * HD, p.qkv, ld, a * HD, m, (qt + 1) * TQ);
      load_tile<HD, WIN>(sdO + (buf ^ 1) * TQ * HD, p.dout, p.C, a * HD, m, (qt + 1) * TQ);
      cp_async_commit();
      load_stats(buf ^ 1, qt + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (qt == 0) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4(s_u32(sK) + tile_off<HD>(r, kk * 2 + (lane >> 4)), ka[kk][0], ka[kk][1], ka[kk][2], ka[kk][3]);
        ldsm_x4(s_u32(sV) + tile_off<HD>(r, kk * 2 + (lane >> 4)), va[kk][0], va[kk][1], va[kk][2], va[kk][3]);
      }
    }
    const uint32_t qb = s_u32(sQ + buf * TQ * HD), ob = s_u32(sdO + buf * TQ * HD);
    float st[8][4], dpt[8][4];   // S^T and dP^T: rows = keys, columns = queries of this tile
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
      dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int n = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const uint32_t o2 = tile_off<HD>(n, kk * 2 + ((lane >> 3) & 1));
        ldsm_x4(qb + o2, b0, b1, b2, b3);
        mma16816(st[np * 2], ka[kk][0], ka[kk][1], ka[kk][2], ka[kk][3], b0, b1);
        mma16816(st[np * 2 + 1], ka[kk][0], ka[kk][1], ka[kk][2], ka[kk][3], b2, b3);
        ldsm_x4(ob + o2, b0, b1, b2, b3);
        mma16816(dpt[np * 2], va[kk][0], va[kk][1], va[kk][2], va[kk][3], b0, b1);
        mma16816(dpt[np * 2 + 1], va[kk][0], va[kk][1], va[kk][2], va[kk][3], b2, b3);
      }
    }
    uint32_t pa[4][4], sa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int jl = nt * 8 + 2 * t;            // query column inside the tile
      const int i0 = qt * TQ + jl;              // sequence-local query index
      const float l0 = sLse[buf * TQ + jl], l1 = sLse[buf * TQ + jl + 1];
      const float e0 = sDl[buf * TQ + jl], e1 = sDl[buf * TQ + jl + 1];
      // logit2(query, key): padded QUERIES contribute nothing; padded KEYS are never written out
      float p0 = (i0 < p.N) ? exp2f(logit2<WIN>(st[nt][0], i0, c_lo, p, m, tbl, off) - l0) : 0.f;
      float p1 = (i0 + 1 < p.N) ? exp2f(logit2<WIN>(st[nt][1], i0 + 1, c_lo, p, m, tbl, off) - l1) : 0.f;
      float p2 = (i0 < p.N) ? exp2f(logit2<WIN>(st[nt][2], i0, c_hi, p, m, tbl, off) - l0) : 0.f;
      float p3 = (i0 + 1 < p.N) ? exp2f(logit2<WIN>(st[nt][3], i0 + 1, c_hi, p, m, tbl, off) - l1) : 0.f;
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
      sa[nt >> 1][(nt & 1) * 2] = pack_bf16(p0 * (dpt[nt][0] - e0), p1 * (dpt[nt][1] - e1));
      sa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2 * (dpt[nt][2] - e0), p3 * (dpt[nt][3] - e1));
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {           // 16 queries per step
#pragma unroll
      for (int np = 0; np < HD / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        const int kr = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const uint32_t o2 = tile_off<HD>(kr, np * 2 + (lane >> 4));
        ldsm_x4_t(ob + o2, b0, b1, b2, b3);
        mma16816(dv[np * 2], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b0, b1);
        mma16816(dv[np * 2 + 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b2, b3);
        ldsm_x4_t(qb + o2, b0, b1, b2, b3);
        mma16816(dk[np * 2], sa[kk][0], sa[kk][1], sa[kk][2], sa[kk][3], b0, b1);
        mma16816(dk[np * 2 + 1], sa[kk][0], sa[kk][1], sa[kk][2], sa[kk][3], b2, b3);
      }
    }
    __syncthreads();
  }
  const int cc[2] = {c_lo, c_hi};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long gr = m.grow(cc[h]);
    if (gr >= 0) {
      bf16* kp = p.dqkv + gr * ld + p.C + a * HD + 2 * t;
      bf16* vp = p.dqkv + gr * ld + 2 * p.C + a * HD + 2 * t;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        *reinterpret_cast<uint32_t*>(kp + i * 8) = pack_bf16(dk[i][h * 2] * p.scale, dk[i][h * 2 + 1] * p.scale);
        *reinterpret_cast<uint32_t*>(vp + i * 8) = pack_bf16(dv[i][h * 2], dv[i][h * 2 + 1]);
      }
    }
  }
}

// dtable[r, a] += sum over (i,j) with relative_position_index[i,j] == r of dense[a, i, j]
// transposed = 1: dense is [heads][key j][query i] (tcgen05 path), else [heads][query i][key j].
__global__ void bias_table_grad_kernel(const float* __restrict__ dense, float* __restrict__ dtable, int heads, int N,
                                       int Npad, int wd, int wh, int ww, int table_len, int transposed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= table_len * heads) return;
  const int a = idx % heads, r = idx / heads;
  const int sw = 2 * ww - 1, shw = (2 * wh - 1) * sw;
  const int dd = r / shw - (wd - 1), dh = (r / sw) % (2 * wh - 1) - (wh - 1), dw = r % sw - (ww - 1);
  float acc = 0.f;
  for (int i = 0; i < N; ++i) {
    const int id = i / (wh * ww), ih = (i / ww) % wh, iw = i % ww;
    const int jd = id - dd, jh = ih - dh, jw = iw - dw;
    if (jd < 0 || jd >= wd || jh < 0 || jh >= wh || jw < 0 || jw >= ww) continue;
    const int j = (jd * wh + jh) * ww + jw;
    acc += transposed ? dense[(static_cast<long long>(a) * Npad + j) * Npad + i]
                      : dense[(static_cast<long long>(a) * Npad + i) * Npad + j];
  }
  dtable[static_cast<long long>(r) * heads + a] += acc;
}

template <int HD>
size_t extra_bytes(const AttnParams& p, bool win, bool with_acc) {
  if (!win) return 0;
  size_t b = MAX_WIN_TOK * 8;
  if (p.table != nullptr) {
    b += static_cast<size_t>((p.table_len + 3) & ~3) * 4;
    if (with_acc) b += static_cast<size_t>(TQ) * (p.Npad + 4) * 4;
  }
  return b;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  VSN_CHECK(bytes <= 227 * 1024, "attention kernel needs %zu bytes of shared memory", bytes);
  VSN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  return 0;
}

int fill_params(AttnParams& p, int S, int N, int heads, int hd, int win, const int* geom, const float* table,
                int table_len, float scale) {
  VSN_CHECK(hd == 32 || hd == 64, "attention head_dim must be 32 or 64 (got %d)", hd);
  p.S = S; p.N = N; p.Npad = ceil_div(N, TQ) * TQ; p.heads = heads; p.C = heads * hd; p.scale = scale;
  p.table = table; p.table_len = table_len; p.groups = 1;
  if (win) {
    VSN_CHECK(geom != nullptr, "window attention needs the geometry array");
    p.B = geom[0]; p.Dp = geom[1]; p.Hp = geom[2]; p.Wp = geom[3];
    p.wd = geom[4]; p.wh = geom[5]; p.ww = geom[6];
    p.sd = geom[7]; p.sh = geom[8]; p.sw = geom[9]; p.use_mask = geom[10];
    VSN_CHECK(p.Dp % p.wd == 0 && p.Hp % p.wh == 0 && p.Wp % p.ww == 0, "stage grid must be a multiple of the window");
    p.nWd = p.Dp / p.wd; p.nWh = p.Hp / p.wh; p.nWw = p.Wp / p.ww;
    VSN_CHECK(N == p.wd * p.wh * p.ww, "window token count mismatch");
    VSN_CHECK(N <= MAX_WIN_TOK, "window of %d tokens exceeds the supported %d", N, MAX_WIN_TOK);
    VSN_CHECK(S == p.B * p.nWd * p.nWh * p.nWw, "window count mismatch");
    VSN_CHECK(table == nullptr || table_len == (2 * p.wd - 1) * (2 * p.wh - 1) * (2 * p.ww - 1), "bias table length mismatch");
    VSN_CHECK(static_cast<long long>(p.B) * p.Dp * p.Hp * p.Wp < (1LL << 31), "token count overflows int32");
    VSN_CHECK(p.Dp < 1024 && p.Hp < 1024 && p.Wp < 1024, "stage grid axis exceeds 1023 tokens");
  } else {
    p.B = S; p.Dp = p.Hp = p.Wp = p.wd = p.wh = p.ww = 1; p.sd = p.sh = p.sw = 0; p.nWd = p.nWh = p.nWw = 1; p.use_mask = 0;
  }
  VSN_CHECK(S <= 65535, "too many sequences for one launch (%d)", S);
  return 0;
}


WinAttnArgs to_tc_args(const AttnParams& p) {
  WinAttnArgs a = {};
  a.qkv = p.qkv; a.out = p.out; a.lse = p.lse; a.table = p.table; a.table_len = p.table_len;
  a.dout = p.dout; a.dqkv = p.dqkv; a.dbias_dense = p.dbias_dense;
  a.S = p.S; a.N = p.N; a.heads = p.heads; a.C = p.C; a.scale = p.scale;
  a.B = p.B; a.Dp = p.Dp; a.Hp = p.Hp; a.Wp = p.Wp; a.wd = p.wd; a.wh = p.wh; a.ww = p.ww;
  a.sd = p.sd; a.sh = p.sh; a.sw = p.sw; a.nWd = p.nWd; a.nWh = p.nWh; a.nWw = p.nWw; a.use_mask = p.use_mask;
  return a;
}

}  // namespace

// geom (window mode, 11 ints): B, Dp, Hp, Wp, wd, wh, ww, shift_d, shift_h, shift_w, use_mask.
extern "C" int vsn_attn_fwd(const void* qkv, void* out, float* lse, int S, int N, int heads, int hd, int win,
                            const int* geom, const float* table, int table_len, float scale, void* stream) {
  AttnParams p = {};
  if (int rc = fill_params(p, S, N, heads, hd, win, geom, table, table_len, scale)) return rc;
  p.qkv = reinterpret_cast<const bf16*>(qkv); p.out = reinterpret_cast<bf16*>(out); p.lse = lse;
  if (S == 0) return 0;
  dim3 grid(p.Npad / TQ, heads, S);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (win && wattn_tc_supported(p.wd, p.wh, p.ww, hd)) return wattn_tc_fwd(to_tc_args(p), st);
  if (!win && dattn_tc_supported(hd)) {
    DenseAttnArgs d = {};
    d.qkv = p.qkv; d.out = p.out; d.lse = p.lse; d.S = S; d.N = N; d.Npad = p.Npad; d.heads = heads; d.C = p.C; d.scale = scale;
    return dattn_tc_fwd(d, st);
  }
#define VSN_FWD(HD, WIN)                                                            \
  {                                                                                 \
    const size_t sm = 5 * Smem<HD>::TILE + extra_bytes<HD>(p, WIN, false);          \
    if (int rc = set_smem(attn_fwd_kernel<HD, WIN>, sm)) return rc;                 \
    attn_fwd_kernel<HD, WIN><<<grid, 128, sm, st>>>(p);                             \
  }
  if (hd == 32 && win) VSN_FWD(32, true)
  else if (hd == 32) VSN_FWD(32, false)
  else if (win) VSN_FWD(64, true)
  else VSN_FWD(64, false)
#undef VSN_FWD
  VSN_LAUNCH_CHECK();
  return 0;
}

// delta and dbias_dense are caller-provided scratch: delta [S, heads, Npad] fp32; dbias_dense [heads, Npad, Npad]
// fp32 (window mode with a table; zeroed by the caller); dtable [table_len, heads] is accumulated into (+=).
extern "C" int vsn_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                            void* dqkv, float* dbias_dense, float* dtable, int S, int N, int heads, int hd, int win,
                            const int* geom, const float* table, int table_len, float scale, void* stream) {
  AttnParams p = {};
  if (int rc = fill_params(p, S, N, heads, hd, win, geom, table, table_len, scale)) return rc;
  p.qkv = reinterpret_cast<const bf16*>(qkv); p.dout = reinterpret_cast<const bf16*>(dout);
  p.lse = const_cast<float*>(lse); p.delta = delta; p.dqkv = reinterpret_cast<bf16*>(dqkv);
  p.dbias_dense = dbias_dense;
  VSN_CHECK(table == nullptr || dtable != nullptr, "bias table given without its gradient buffer");
  const bool tc_path = win && wattn_tc_supported(p.wd, p.wh, p.ww, hd);
  VSN_CHECK(table == nullptr || tc_path || dbias_dense != nullptr, "the mma.sync path needs the dense bias-gradient scratch");
  if (S == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bf16* o = reinterpret_cast<const bf16*>(out);
  if (win && wattn_tc_supported(p.wd, p.wh, p.ww, hd)) {
    p.out = const_cast<bf16*>(o);
    WinAttnArgs ta = to_tc_args(p);
    ta.dtable = table != nullptr ? dtable : nullptr;
    return wattn_tc_bwd(ta, delta, st);
  }
  if (!win && dattn_tc_supported(hd)) {
    // dense head_dim 64 (ViT-3D): tcgen05 kernels; `dbias_dense` is the fp32 dQ scratch [S*N, C] here
    DenseAttnArgs d = {};
    d.qkv = p.qkv; d.out = const_cast<bf16*>(o); d.lse = p.lse; d.dout = p.dout; d.dqkv = p.dqkv; d.delta = delta;
    d.dq_acc = dbias_dense; d.S = S; d.N = N; d.Npad = p.Npad; d.heads = heads; d.C = p.C; d.scale = scale;
    return dattn_tc_bwd(d, st);
  }
  if (win) attn_delta_kernel<true><<<S, 128, 0, st>>>(p, o, delta);
  else attn_delta_kernel<false><<<S, 128, 0, st>>>(p, o, delta);
  VSN_LAUNCH_CHECK();
  const int qtiles = p.Npad / TQ;
  // dQ: with a bias table, keep ~2 waves of persistent CTAs so the dS strip is flushed rarely
  int groups = S;
  if (table != nullptr) {
    groups = ceil_div(2 * vsn_num_sms(), qtiles * heads);
    if (groups > S) groups = S;
    if (groups < 1) groups = 1;
  }
  p.groups = groups;
  dim3 gq(qtiles, heads, groups), gk(qtiles, heads, S);
#define VSN_BWD(HD, WIN)                                                            \
  {                                                                                 \
    const size_t sq = 6 * Smem<HD>::TILE + extra_bytes<HD>(p, WIN, true);           \
    if (int rc = set_smem(attn_bwd_dq_kernel<HD, WIN>, sq)) return rc;              \
    attn_bwd_dq_kernel<HD, WIN><<<gq, 128, sq, st>>>(p);                            \
    VSN_LAUNCH_CHECK();                                                             \
    const size_t sk = 6 * Smem<HD>::TILE + 4 * TQ * 4 + extra_bytes<HD>(p, WIN, false); \
    if (int rc = set_smem(attn_bwd_dkv_kernel<HD, WIN>, sk)) return rc;             \
    attn_bwd_dkv_kernel<HD, WIN><<<gk, 128, sk, st>>>(p);                           \
  }
  if (hd == 32 && win) VSN_BWD(32, true)
  else if (hd == 32) VSN_BWD(32, false)
  else if (win) VSN_BWD(64, true)
  else VSN_BWD(64, false)
#undef VSN_BWD
  VSN_LAUNCH_CHECK();
  if (table != nullptr) {
    const int n = table_len * heads;
    bias_table_grad_kernel<<<ceil_div(n, 128), 128, 0, st>>>(dbias_dense, dtable, heads, N, p.Npad, p.wd, p.wh, p.ww, table_len, 0);
    VSN_LAUNCH_CHECK();
  }
  return 0;
}
