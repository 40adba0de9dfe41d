// Swin-3D shifted-window attention core on the 5th-gen tensor cores (tcgen05 + TMEM), forward and backward.
// Covers WindowAttention3D.forward between the qkv and proj Linears together with the torch.roll /
// window_partition / window_reverse / shift-mask plumbing around it
// (models/swin_transformer_3d.py:72-89,132-152,162-196,330-358,463-492) for head_dim 32 and windows of at
// most 252 tokens (padded to 256 inside the SM) -- every stage of the shipped Swin configs.
//
// Forward, one persistent CTA per (head, group of windows):
//   smem   compact relative-position bias table [(da,db,w_i)][w_j] * log2(e) as bf16 (13.4 KB, built once per CTA),
//          a 2-deep ring of {Q,K,V}[256][32] bf16 tiles in the 64-byte-swizzled UMMA layout (96 KB),
//          per-token region codes and source rows of the window (written by the loader with the tiles).
//   TMEM   two 256-column regions; region h holds S = Q_h K^T for the half h of the window's queries
//          (128 lanes x 256 fp32 columns = four 64-key parts).  A softmax thread writes P of a part (bf16
//          pairs) back over the first 32 of the part's own 64 columns, from where it feeds the second MMA; the
//          part's partial output O_part = P_part V (exponentiated against the part's local max) accumulates
//          in the other 32 columns.
//   warps  0-15 softmax in two independent groups (group = query half: its own TMEM region and barriers, so one
//          group computes while the other waits for its MMAs), 16 = cp.async gather of the next window (roll and
//          window_partition are index math on the source rows), 17 = MMA issuer (converged warp, elect.sync).
// Nothing of size N x N ever goes to HBM: no rolled copy, no mask tensor, no bias tensor, no logits.
#include <stdlib.h>
#include "tc.cuh"
#include "wattn_tc.cuh"

namespace {

constexpr int NP = 256;                    // padded tokens per window
constexpr int HD = 32;
constexpr int TILE_BYTES = NP * HD * 2;    // 16 KB
constexpr int STAGE_BYTES = 3 * TILE_BYTES;   // Q | K | V (forward kernel; the one-hot region tile follows the ring)
// One-hot region tiles of the forward's tensor-core shift mask: queries carry REGION_Q, keys REGION_K (both exact in
// bf16).  REGION_Q * REGION_K * scale = 565.734375 * 32^-1/2 = 100.0087 nats: the reference's 100 to 9e-5 relative
// (a single value e for both sides cannot do better than e^2 = 564.06 -> 99.7 or 576 -> 101.8 nats).
constexpr float REGION_Q = 9.3125f;            // bf16 0x4115
constexpr float REGION_K = 60.75f;             // bf16 0x4273
#ifdef VSN_F16
constexpr uint32_t REGION_Q_BITS = 0x48A8u, REGION_K_BITS = 0x5398u;   // the same two values as IEEE half
#else
constexpr uint32_t REGION_Q_BITS = 0x4115u, REGION_K_BITS = 0x4273u;
#endif
constexpr float REGION_SQ = REGION_Q * REGION_K;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float MASK_L2E = -100.0f * LOG2E;
constexpr float PAD_BIAS = -30000.0f;      // keys >= N

__device__ __forceinline__ uint32_t swz64(int row, int chunk) {   // byte offset inside a [rows][32] bf16 tile
  return static_cast<uint32_t>(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

struct TokenGeom {
  int row;      // global row of the token in the [T, *] matrices
  int code;     // region id 0..26 on the rolled grid (models/swin_transformer_3d.py:463-492)
  int lin;      // relative-position linear code of the token inside its window
};

// Relative-position bias in smem, window dims static: row ((a*(2WH-1) + b)*WW + wi) holds the 16-byte vector
// { table[(a, b, wi - wj + WW-1), head] * log2(e) : wj = 0..WW-1 } as bf16, with a = di-dj+WD-1, b = hi-hj+WH-1.
// For query token i=(di,hi,wi) and the key row kr=(dj,hj) of the window the vector sits at row
//   rowbase(i) - (dj*(2WH-1) + hj)*WW,   rowbase(i) = ((di+WD-1)*(2WH-1) + hi+WH-1)*WW + wi,
// so a thread fetches the bias of WW consecutive keys with one LDS.128 and every element position is a
// compile-time constant.  (2WD-1)(2WH-1)WW rows = 13.4 KB for the (6,7,6) window instead of a dense 126 KB tile.
template <int WD, int WH, int WW>
struct BiasTab {
  static_assert(WW <= 8, "one 16-byte row holds the WW keys of a key row");
  static constexpr int ROWS = (2 * WD - 1) * (2 * WH - 1) * WW;
  static constexpr int BYTES = ROWS * 16;
  __device__ static void build(const WinAttnArgs& p, int head, uint8_t* tab) {
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
      const int wi = r % WW, ab = r / WW;
      float v[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        v[t] = 0.f;
        if (t < WW && p.table != nullptr)
          v[t] = p.table[static_cast<long long>(ab * (2 * WW - 1) + (wi - t + WW - 1)) * p.heads + head] * LOG2E;
      }
      uint4 u;
      u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(tab + r * 16) = u;
    }
  }
  __device__ __forceinline__ static int rowbase(int i) {
    const int di = i / (WH * WW), lr = i - di * (WH * WW), hi = lr / WW, wi = lr - hi * WW;
    return ((di + WD - 1) * (2 * WH - 1) + hi + WH - 1) * WW + wi;
  }
};

// Per-window part of the geometry (three runtime integer divisions): computed once per window, not per tile.
struct WinCoord {
  int bm;     // batch index << 1 | masked (the window straddles a region boundary: last along a shifted axis)
  int dhw;    // origin of the window on the rolled (shifted) grid, 10 bits per axis
  __device__ __forceinline__ bool masked() const { return bm & 1; }
};
__device__ __forceinline__ WinCoord win_coord(const WinAttnArgs& p, int s, int wd, int wh, int ww) {
  const int nW = p.nWd * p.nWh * p.nWw;
  const int b = s / nW, wi = s - b * nW;
  const int wa = wi / (p.nWh * p.nWw), wr = wi - wa * (p.nWh * p.nWw), wb = wr / p.nWw, wc = wr - wb * p.nWw;
  const bool masked = p.use_mask && ((p.sd > 0 && wa == p.nWd - 1) || (p.sh > 0 && wb == p.nWh - 1) || (p.sw > 0 && wc == p.nWw - 1));
  WinCoord c;
  c.bm = (b << 1) | (masked ? 1 : 0);
  c.dhw = ((wa * wd) << 20) | ((wb * wh) << 10) | (wc * ww);
  return c;
}
// token i of the window (window_partition order) -> source row after the cyclic shift and region code
// (models/swin_transformer_3d.py:333-341,463-492); only divisions by compile-time constants
template <int WD, int WH, int WW>
__device__ __forceinline__ TokenGeom token_geom_w(const WinAttnArgs& p, const WinCoord& c, int i) {
  const int ld = i / (WH * WW), lr = i - ld * (WH * WW), lh = lr / WW, lw = lr - lh * WW;
  const int dd = (c.dhw >> 20) + ld, hh = ((c.dhw >> 10) & 1023) + lh, wv = (c.dhw & 1023) + lw;
  int d0 = dd + p.sd; if (d0 >= p.Dp) d0 -= p.Dp;
  int h0 = hh + p.sh; if (h0 >= p.Hp) h0 -= p.Hp;
  int w0 = wv + p.sw; if (w0 >= p.Wp) w0 -= p.Wp;
  TokenGeom g;
  g.row = (((c.bm >> 1) * p.Dp + d0) * p.Hp + h0) * p.Wp + w0;
  const int rd = dd < p.Dp - WD ? 0 : (dd < p.Dp - p.sd ? 1 : 2);
  const int rh = hh < p.Hp - WH ? 0 : (hh < p.Hp - p.sh ? 1 : 2);
  const int rw = wv < p.Wp - WW ? 0 : (wv < p.Wp - p.sw ? 1 : 2);
  g.code = 9 * rd + 3 * rh + rw;
  g.lin = 0;
  return g;
}
template <int WD, int WH, int WW>
__device__ __forceinline__ TokenGeom token_geom_t(const WinAttnArgs& p, int s, int i) {
  return token_geom_w<WD, WH, WW>(p, win_coord(p, s, WD, WH, WW), i);
}

// =========================================== forward ==============================================
constexpr int SM_WARPS = 16;                      // softmax warps: 4 TMEM lane quadrants x 4 column parts
constexpr int FWD_THREADS = (SM_WARPS + 4) * 32;  // + one warpgroup of helpers: loader warp, MMA warp, two idle
constexpr int FWD_STAGES = 2;

template <int WD, int WH, int WW>
struct FwdSmem {
  static constexpr int ONEHOT = FWD_STAGES * STAGE_BYTES;      // query-side and key-side one-hot tiles, shared by the two stages (see the loader)
  static constexpr int BIAS = ONEHOT + 2 * TILE_BYTES;
  static constexpr int KEYCODE = BIAS + BiasTab<WD, WH, WW>::BYTES;
  static constexpr int ROWIDX = KEYCODE + FWD_STAGES * NP;     // int [2][256] source row of every token + [2] masked flag
  static constexpr int STATS = ROWIDX + FWD_STAGES * NP * 4 + 16;   // float2 [2][4 parts][128]
  static constexpr int BARS = STATS + 2 * 4 * 128 * 8;
  static constexpr int TOTAL = BARS + 24 * 8;
};

// 2^x for x <= 0 on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial, relative error
// ~1e-4, far below the bf16 rounding of P): relieves the MUFU pipe, which is the bound of this kernel.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;                 // 1.5 * 2^23: integer part lands in the low mantissa bits
  const float r = x - (t - 12582912.f);           // r in [-0.5, 0.5]
  float p = fmaf(r, 0.0555041087f, 0.2402264923f);
  p = fmaf(p, r, 0.6931471806f);
  p = fmaf(p, r, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// x[k] = s[k] * cscale + bias(i, key 64*PART + k) for the 64 keys of a column part (padded keys -> PAD_BIAS)
template <int WD, int WH, int WW, int PART>
__device__ __forceinline__ void add_bias(uint32_t* x, const uint8_t* tab, int rowbase, float cscale) {
  constexpr int N = WD * WH * WW;
  constexpr int KR0 = (PART * 64) / WW, KR1 = (PART * 64 + 63) / WW;
#pragma unroll
  for (int kr = KR0; kr <= KR1; ++kr) {
    if (kr * WW >= N) continue;
    const int dj = kr / WH, hj = kr - dj * WH;
    const uint4 bb = *reinterpret_cast<const uint4*>(tab + (rowbase - (dj * (2 * WH - 1) + hj) * WW) * 16);
    const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int wj = 0; wj < WW; ++wj) {
      const int k = kr * WW + wj - PART * 64;
      if (k < 0 || k >= 64) continue;
      const float b = (wj & 1) ? bf16hi(bw[wj >> 1]) : bf16lo(bw[wj >> 1]);
      x[k] = __float_as_uint(fmaf(__uint_as_float(x[k]), cscale, b));
    }
  }
#pragma unroll
  for (int k = 0; k < 64; ++k)
    if (PART * 64 + k >= N) x[k] = __float_as_uint(PAD_BIAS);
}

template <int WD, int WH, int WW, int VAR>
__global__ void __launch_bounds__(FWD_THREADS, 1) wattn_fwd_kernel(const WinAttnArgs p) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  using SM = FwdSmem<WD, WH, WW>;
  using BT = BiasTab<WD, WH, WW>;
  constexpr int N = WD * WH * WW;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stages = smem;
  uint8_t* bias_s = smem + SM::BIAS;
  uint8_t* keycode = smem + SM::KEYCODE;                            // loader scratch: region code of every token
  int* rowidx = reinterpret_cast<int*>(smem + SM::ROWIDX);        // written by the loader with the tiles
  int* winmask = rowidx + FWD_STAGES * NP;
  float2* stats = reinterpret_cast<float2*>(smem + SM::STATS);    // (local max, local sum)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BARS);
  uint64_t* qkv_full = bars;        // [2] loader lanes -> MMA
  uint64_t* qkv_empty = bars + 2;   // [2] MMA commit -> loader
  uint64_t* s_full = bars + 4;      // [2] MMA commit -> softmax
  uint64_t* p_ready = bars + 6;     // [2] softmax threads -> MMA
  uint64_t* o_full = bars + 8;      // [2] MMA commit -> softmax (epilogue)
  uint64_t* o_read = bars + 10;     // [2] softmax threads -> MMA (region may be overwritten)
  uint64_t* st_full = bars + 12;    // [4 quadrants][2] row statistics of a unit are in smem
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int n_it = (p.S - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int U = 2 * n_it;

  if (threadIdx.x == 0) {
    if ((tc::smem_u32(smem) & 1023u) != 0) {
      printf("vsn_b200: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&qkv_full[i], 32);
      tc::mbar_init(&qkv_empty[i], 1);
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&p_ready[i], SM_WARPS * 16);   // one group of 8 warps per query half
      tc::mbar_init(&o_full[i], 1);
      tc::mbar_init(&o_read[i], SM_WARPS * 16);
    }
    for (int i = 0; i < 8; ++i) tc::mbar_init(&st_full[i], 128);
    tc::fence_barrier_init();
  }
  if (warp == SM_WARPS + 1) tc::tmem_alloc(tmem_slot, 512);
  BT::build(p, head, bias_s);
  // rows N..255 of every tile stay zero for the whole kernel (the loader only writes rows < N)
  for (int idx = threadIdx.x; idx < (FWD_STAGES * 3 + 2) * (NP - N) * 4; idx += blockDim.x) {
    const int tile = idx / ((NP - N) * 4), rem = idx - tile * ((NP - N) * 4);
    const int row = N + rem / 4, c = rem & 3;
    *reinterpret_cast<uint4*>(stages + tile * TILE_BYTES + swz64(row, c)) = make_uint4(0, 0, 0, 0);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // Register budget: the launch gives every thread 96 registers; the helper warpgroup hands most of its share back
  // and the softmax warpgroups grow to 104 (they hold 64 logits each and were spilling).
  // (each role's code sits inside the branch of its own setmaxnreg: after a join ptxas allocates for the lower limit)
  if (warp >= SM_WARPS) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == SM_WARPS) {
    // ------------------------------------------------------------------ loader
    const long long ld = 3LL * p.C;
    for (int it = 0; it < n_it; ++it) {
      const int s = blockIdx.x + it * gridDim.x;
      const int st = it & 1;
      tc::mbar_wait_relaxed(&qkv_empty[st], ((it >> 1) & 1) ^ 1);
      uint8_t* sq = stages + st * STAGE_BYTES;
      const WinCoord wc = win_coord(p, s, WD, WH, WW);
      for (int i = lane; i < N; i += 32) {
        const TokenGeom g = token_geom_w<WD, WH, WW>(p, wc, i);
        rowidx[st * NP + i] = g.row;
        keycode[i] = static_cast<uint8_t>(g.code);      // (this lane reads it back below)
        const bf16* src = p.qkv + g.row * ld + head * HD;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t o = swz64(i, c);
          tc::cp_async16(sq + o, src + c * 8);
          tc::cp_async16(sq + TILE_BYTES + o, src + p.C + c * 8);
          tc::cp_async16(sq + 2 * TILE_BYTES + o, src + 2 * p.C + c * 8);
        }
      }
      if (wc.masked()) {
        // shift mask on the tensor cores: Eq[i][c] = REGION_Q and Ek[i][c] = REGION_K where c is the region of token i,
        // so that (Eq Ek^T)[i][j] = REGION_SQ for tokens of the same region, 0 otherwise -- added to Q K^T by two more
        // K=16 steps.  The reference adds -100 to pairs of DIFFERENT regions (models/swin_transformer_3d.py:463-492);
        // softmax is invariant under the per-row constant that separates the two forms.
        // The tiles are not double-buffered: the S MMAs of the previous window (its second query half last) must have
        // read them -- they retire while this window's Q / K / V are still in flight.
        if (it > 0) tc::mbar_wait(&s_full[1], (it - 1) & 1);
        uint8_t* onehot = stages + FWD_STAGES * STAGE_BYTES;
        for (int i = lane; i < N; i += 32) {
          const int code = keycode[i];
          uint32_t wq[4] = {0u, 0u, 0u, 0u}, wk[4] = {0u, 0u, 0u, 0u};
          wq[(code & 7) >> 1] = (code & 1) ? (REGION_Q_BITS << 16) : REGION_Q_BITS;   // the pair's high / low half
          wk[(code & 7) >> 1] = (code & 1) ? (REGION_K_BITS << 16) : REGION_K_BITS;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const bool hit = c == (code >> 3);
            *reinterpret_cast<uint4*>(onehot + swz64(i, c)) =
                hit ? make_uint4(wq[0], wq[1], wq[2], wq[3]) : make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(onehot + TILE_BYTES + swz64(i, c)) =
                hit ? make_uint4(wk[0], wk[1], wk[2], wk[3]) : make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
      if (lane == 0) winmask[st] = wc.masked() ? 1 : 0;
      tc::cp_async_wait_all();
      tc::fence_proxy_async();
      tc::mbar_arrive(&qkv_full[st]);
    }
  } else if (warp == SM_WARPS + 1) {
    // ------------------------------------------------------------------ MMA issuer (warp runs converged)
    {
      const uint32_t idesc_s = tc::make_idesc_bf16(128, NP, 0, 0);
      const uint32_t idesc_o = tc::make_idesc_bf16(128, HD, 0, 1);
      const uint64_t desc_st0 = tc::make_smem_desc_sw64(tc::smem_u32(stages), 16, 512);
      for (int u = 0; u <= U; ++u) {
        if (u < U) {
          const int it = u >> 1, h = u & 1, st = it & 1;
          if (h == 0) tc::mbar_wait(&qkv_full[st], (it >> 1) & 1);
          tc::mbar_wait(&o_read[h], ((u >> 1) & 1) ^ 1);
          tc::fence_after_sync();
          const uint64_t dq = tc::desc_advance(desc_st0, st * STAGE_BYTES + h * (128 * 64));
          const uint64_t dk = tc::desc_advance(desc_st0, st * STAGE_BYTES + TILE_BYTES);
          const bool win_masked = winmask[st] != 0;
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
              tc::mma_bf16_ss(tmem_base + h * NP, tc::desc_advance(dq, k * 32), tc::desc_advance(dk, k * 32), idesc_s,
                              k > 0 ? 1u : 0u);
            if (win_masked) {        // + Eq_h Ek^T: the shift mask
              const uint64_t de = tc::desc_advance(desc_st0, FWD_STAGES * STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < 2; ++k)
                tc::mma_bf16_ss(tmem_base + h * NP, tc::desc_advance(de, h * (128 * 64) + k * 32),
                                tc::desc_advance(de, TILE_BYTES + k * 32), idesc_s, 1u);
            }
            tc::mma_commit(&s_full[h]);
          }
          __syncwarp();
        }
        if (u >= 1) {
          const int v = u - 1, it = v >> 1, h = v & 1, st = it & 1;
          tc::mbar_wait(&p_ready[h], (v >> 1) & 1);
          tc::fence_after_sync();
          const uint64_t dv = tc::desc_advance(desc_st0, st * STAGE_BYTES + 2 * TILE_BYTES);
          if (tc::elect_one()) {
            // keys 64*part .. 64*part+63 were exponentiated against their own local max: one accumulator each
#pragma unroll
            for (int k = 0; k < NP / 16; ++k)
              tc::mma_bf16_ts(tmem_base + h * NP + (k >> 2) * 64 + 32, tmem_base + h * NP + (k >> 2) * 64 + (k & 3) * 8,
                              tc::desc_advance(dv, k * 1024), idesc_o, (k & 3) ? 1u : 0u);
            tc::mma_commit(&o_full[h]);
            if (h == 1) tc::mma_commit(&qkv_empty[st]);
          }
          __syncwarp();
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ------------------------------------------------------------------ softmax + epilogue
    // Two independent groups of 8 warps: group g owns the query half h = g of every window (its own TMEM region,
    // its own barriers), so while one group waits for its MMAs (P V, then the next S) the other one keeps the
    // MUFU / FMA pipes busy -- the groups run half a period apart instead of all 16 warps in lock step.
    // Thread = (TMEM lane rl, column half ch): two sequential passes over 64 of the row's 256 logits each
    // (part = 2*pass + ch), exponentiated against the LOCAL max of those 64 (no cross-thread exchange on the
    // critical path).  P V is accumulated per part; the epilogue rescales the four partial outputs to the row max.
    const int h = warp >> 3, q4 = warp & 3, ch = (warp >> 2) & 1;
    const int rl = q4 * 32 + lane;
    const int i = h * 128 + rl;
    const int ib = i < N ? i : N - 1;
    const int rowbase = BT::rowbase(ib);
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const uint32_t region = tmem_base + lane_addr + h * NP;
    const float cscale = p.scale * LOG2E;

    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      const int s = blockIdx.x + it * gridDim.x;

      tc::mbar_wait(&s_full[h], it & 1);
      tc::fence_after_sync();
      // geometry of the window as the loader worked it out (stage st is not refilled before this unit's P V ran)
      const int out_row = rowidx[st * NP + ib];
      // masked windows carry REGION_SQ * cscale on every same-region logit (the row maximum is one of them): taken
      // out of the stored log-sum-exp, which the backward kernel combines with the reference's -100 form
      const float lse_off = winmask[st] != 0 ? REGION_SQ * cscale : 0.f;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int part = 2 * pass + ch;
        uint32_t x[64];
        tc::tmem_ld_32x32b_x32(region + part * 64, x);
        tc::tmem_ld_32x32b_x32(region + part * 64 + 32, x + 32);
        tc::tmem_ld_wait();
        // (P of a part goes back over the first 32 of its OWN 64 logit columns and its partial output accumulates
        // in the other 32: no thread writes columns that another one still has to read)
        if (pass == 0) {
          if (ch == 0) add_bias<WD, WH, WW, 0>(x, bias_s, rowbase, cscale);
          else add_bias<WD, WH, WW, 1>(x, bias_s, rowbase, cscale);
        } else {
          if (ch == 0) add_bias<WD, WH, WW, 2>(x, bias_s, rowbase, cscale);
          else add_bias<WD, WH, WW, 3>(x, bias_s, rowbase, cscale);
        }
        float mx0 = -3.0e38f, mx1 = -3.0e38f;
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(x[k]), __uint_as_float(x[k + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(x[k + 2]), __uint_as_float(x[k + 3])));
        }
        // a part made only of padded keys / masked-out keys keeps a finite reference
        const float m = fmaxf(fmaxf(mx0, mx1), -20000.f);

        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          // VAR&1: 1 of 4 exponentials on the FMA pipe instead of the MUFU pipe
          const float p0 = tc::ex2_approx(__uint_as_float(x[2 * k]) - m);
          const float p1 = (VAR & 1) ? exp2_poly(__uint_as_float(x[2 * k + 1]) - m) : tc::ex2_approx(__uint_as_float(x[2 * k + 1]) - m);
          const float p2 = tc::ex2_approx(__uint_as_float(x[2 * k + 2]) - m);
          const float p3 = tc::ex2_approx(__uint_as_float(x[2 * k + 3]) - m);
          l0 += p0 + p1;
          l1 += p2 + p3;
          x[k] = pack_bf16(p0, p1);
          x[k + 1] = pack_bf16(p2, p3);
        }
        tc::tmem_st_32x32b_x32(region + part * 64, x);
        stats[h * 512 + part * 128 + rl] = make_float2(m, l0 + l1);
        tc::mbar_arrive(&st_full[q4 * 2 + h]);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc::mbar_arrive(&p_ready[h]);

      // ---- epilogue of this unit: combine the four partial outputs (16 of the 32 output columns per thread)
      tc::mbar_wait(&st_full[q4 * 2 + h], it & 1);
      const float2* sp = stats + h * 512 + rl;
      const float2 s0 = sp[0], s1 = sp[128], s2 = sp[256], s3 = sp[384];
      const float m = fmaxf(fmaxf(s0.x, s1.x), fmaxf(s2.x, s3.x));
      const float w0 = tc::ex2_approx(s0.x - m), w1 = tc::ex2_approx(s1.x - m);
      const float w2 = tc::ex2_approx(s2.x - m), w3 = tc::ex2_approx(s3.x - m);
      const float l = fmaf(w0, s0.y, fmaf(w1, s1.y, fmaf(w2, s2.y, w3 * s3.y)));
      const float inv = 1.f / l;
      tc::mbar_wait(&o_full[h], it & 1);
      tc::fence_after_sync();
      const uint32_t ob = region + 32 + ch * 16;            // partial output of part p: columns 64 p + 32 ..
      float r[16];
      {
        uint32_t o[16];
        tc::tmem_ld_32x32b_x16(ob, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = w0 * __uint_as_float(o[k]);
        tc::tmem_ld_32x32b_x16(ob + 64, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = fmaf(w1, __uint_as_float(o[k]), r[k]);
        tc::tmem_ld_32x32b_x16(ob + 128, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = fmaf(w2, __uint_as_float(o[k]), r[k]);
        tc::tmem_ld_32x32b_x16(ob + 192, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = fmaf(w3, __uint_as_float(o[k]), r[k]) * inv;
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&o_read[h]);
      if (i < N) {
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = pack_bf16(r[2 * k], r[2 * k + 1]);
        st_global_v8(p.out + static_cast<long long>(out_row) * p.C + head * HD + ch * 16, w);
      }
      if (ch == 0 && p.lse != nullptr)
        p.lse[(static_cast<long long>(s) * p.heads + head) * NP + i] = m + log2f(l) - lse_off;   // log2 domain (consumed by wattn_bwd_kernel)
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == SM_WARPS + 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

int groups_for(int S, int heads) {
  int g = vsn_num_sms() / heads;
  if (g < 1) g = 1;
  if (g > S) g = S;
  // even out the tail: the smallest group count with the same number of rounds
  const int rounds = ceil_div(S, g);
  return ceil_div(S, rounds);
}

}  // namespace

// The tcgen05 kernels are instantiated for the (6,7,6) window of the shipped Swin-3D configs
// (config-defaults.yaml WINDOW_SIZE); other windows / head dims use the mma.sync kernels in attn.cu.
bool wattn_tc_supported(int wd, int wh, int ww, int hd) { return hd == HD && wd == 6 && wh == 7 && ww == 6; }

int wattn_tc_fwd(const WinAttnArgs& a, cudaStream_t stream) {
  using SM = FwdSmem<6, 7, 6>;
  static int var = -1;
  if (var < 0) {
    // VSN_WATTN_VARIANT (measurements only): 0 = every exponential on the MUFU pipe, 1 = one of four on the FMA
    // pipe (polynomial)
    const char* e = getenv("VSN_WATTN_VARIANT");
    var = e ? atoi(e) & 1 : 0;
    VSN_CUDA(cudaFuncSetAttribute(wattn_fwd_kernel<6, 7, 6, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    VSN_CUDA(cudaFuncSetAttribute(wattn_fwd_kernel<6, 7, 6, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
  }
  dim3 grid(groups_for(a.S, a.heads), a.heads);
  if (var == 1) wattn_fwd_kernel<6, 7, 6, 1><<<grid, FWD_THREADS, SM::TOTAL, stream>>>(a);
  else wattn_fwd_kernel<6, 7, 6, 0><<<grid, FWD_THREADS, SM::TOTAL, stream>>>(a);
  VSN_LAUNCH_CHECK();
  return 0;
}

// =========================================== backward =============================================
// CTA = (head, half kh of the window's keys), persistent over a group of windows.  TMEM lanes are KEYS:
//   S^T = K_half Q^T and dP^T = V_half dO^T are computed per sub-tile of 32 queries (two 64-column buffers),
//   the compute threads turn them into P^T = exp2(S^T*c + B^T - lse) and dS^T = P^T (dP^T - delta) in place
//   (bf16 pairs over their own columns), which feed straight from TMEM
//     dV += P^T dO_sub,   dK += dS^T Q_sub,   dBias[:, sub] += dS^T I16   (identity B operand: the dense
//     bias gradient of this (head, key half) accumulates over ALL windows of the CTA in 256 TMEM columns),
//   while a copy of dS^T in an MN-major smem tile gives dQ_half = dS K_half (two 128-query blocks).
//   dQ is the sum of the two key halves: bf16x2 red.add into the zeroed Q block of dqkv.
// Warps: 0-7 compute group 0 (even sub-tiles), 8-15 group 1 (odd sub-tiles; also the per-window read-out of
// dK / dV / dQ), 16 loader, 17 MMA issuer.  Thread = (key row, 16 of the sub-tile's 32 queries).
constexpr int BWD_THREADS = 20 * 32;           // 16 compute warps + one warpgroup of helpers (loader, MMA issuer, two idle)
constexpr int QS = 32;                         // queries per sub-tile
constexpr int NSUB = NP / QS;                  // 8
// stage: Q 16K | dO 16K | K_half 8K | V_half 8K | lse2 1K | delta 1K | region code 256 | source row 1K | masked flag
constexpr int BWD_OFF_CODE = 51200, BWD_OFF_ROW = 51456, BWD_OFF_FLAG = 52480;
constexpr int BWD_STAGE_BYTES = 52736;         // 2 stages = 103 KB (1024-aligned for the dS tile behind them)
constexpr int DS_TILE_BYTES = 128 * NP * 2;    // 64 KB: dS^T as MN-major A operand, 128B swizzle

template <int WD, int WH, int WW>
struct BwdSmem {
  static constexpr int DS = 2 * BWD_STAGE_BYTES;                    // 105472 (1024-aligned)
  static constexpr int BIAS = DS + DS_TILE_BYTES;
  static constexpr int IDENT = BIAS + ((BiasTab<WD, WH, WW>::BYTES + 1023) / 1024) * 1024;
  static constexpr int BARS = IDENT + 1024;
  static constexpr int TOTAL = BARS + 32 * 8;
};

// Transposed bias table: row ((a*(2WH-1)+b)*WW + wj) holds { table[(a, b, wi - wj + WW-1)] * log2e : wi = 0..WW-1 },
// so a KEY thread fetches the bias against the WW queries of one query row with one LDS.128:
//   row = rowbaseT(j) + (di*(2WH-1) + hi)*WW,  rowbaseT(j) = ((WD-1-dj)*(2WH-1) + WH-1-hj)*WW + wj.
template <int WD, int WH, int WW>
__device__ void build_bias_t(const WinAttnArgs& p, int head, uint8_t* tab) {
  constexpr int ROWS = BiasTab<WD, WH, WW>::ROWS;
  for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
    const int wj = r % WW, ab = r / WW;
    float v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      v[t] = 0.f;
      if (t < WW && p.table != nullptr)
        v[t] = p.table[static_cast<long long>(ab * (2 * WW - 1) + (t - wj + WW - 1)) * p.heads + head] * LOG2E;
    }
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(tab + r * 16) = u;
  }
}

// x[k] = s[k]*cscale + bias(query q0+k, my key) for 16 consecutive queries starting at q0 (q0 % 6 == W0)
template <int WD, int WH, int WW, int W0>
__device__ __forceinline__ void add_bias_t16(float* x, const uint8_t* tab, int rowbaseT, int qr0, float cscale) {
  constexpr int NR = (W0 + 15) / WW + 1;
  constexpr int QR_MAX = WD * WH - 1;
  uint32_t bw[NR][4];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    int qr = qr0 + r;
    qr = qr > QR_MAX ? QR_MAX : qr;                       // padded queries read a valid row (result unused)
    const int di = qr / WH, hi = qr - di * WH;
    const uint4 bb = *reinterpret_cast<const uint4*>(tab + (rowbaseT + (di * (2 * WH - 1) + hi) * WW) * 16);
    bw[r][0] = bb.x; bw[r][1] = bb.y; bw[r][2] = bb.z; bw[r][3] = bb.w;
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int r = (W0 + k) / WW, wi = (W0 + k) % WW;
    const float b = (wi & 1) ? bf16hi(bw[r][wi >> 1]) : bf16lo(bw[r][wi >> 1]);
    x[k] = fmaf(x[k], cscale, b);
  }
}

// 8 bf16 (16 bytes) added to global memory with one REDG.BF16x8
__device__ __forceinline__ void red_add_bf16x8(bf16* addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.noftz.v4." VSN_T16 "x2 [%0], {%1, %2, %3, %4};" ::"l"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// Per-phase cycle counters of the backward kernel: compiled in with -DVSN_WATTN_TIMING (they cost a dozen registers
// in a kernel that runs at the 96-register cap: 0.590 -> 0.565 ms at stage 0 without them), printed when the
// environment has VSN_WATTN_TIMING=1.
__device__ unsigned long long* g_bwd_timing = nullptr;   // [gridDim][16]
#ifdef VSN_WATTN_TIMING
#define TCLK() clock64()
#else
#define TCLK() 0ll
#endif

template <int WD, int WH, int WW>
__global__ void __launch_bounds__(BWD_THREADS, 1) wattn_bwd_kernel(const WinAttnArgs p, const float* __restrict__ delta_g) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  using SM = BwdSmem<WD, WH, WW>;
  constexpr int N = WD * WH * WW;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stages = smem;
  uint8_t* ds_tile = smem + SM::DS;
  uint8_t* bias_s = smem + SM::BIAS;
  uint8_t* ident = smem + SM::IDENT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BARS);
  uint64_t* ld_full = bars;          // [2] loader lanes -> MMA
  uint64_t* ld_empty = bars + 2;     // [2] MMA commit -> loader
  uint64_t* s_full = bars + 4;       // [2 buffers] MMA commit -> compute group
  uint64_t* p_ready = bars + 6;      // [2 buffers] compute group -> MMA
  uint64_t* ds_free = bars + 8;      // [2 query blocks] MMA commit (dQ block done) -> compute threads (smem tile half)
  uint64_t* unit_done = bars + 10;   // MMA commit: every MMA of the window (incl. dQ) complete -> group 1
  uint64_t* acc_read = bars + 11;    // group 1: dV / dK accumulators read out -> MMA (next window's first dV/dK MMA)
  uint64_t* kv_done = bars + 12;     // MMA commit: dV, dK of the window complete -> group 1
  uint64_t* dq_read = bars + 13;     // group 1: dQ accumulator read out -> MMA (next window's first dQ MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, kh = blockIdx.z;
  const int n_units = (p.S - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int TT = n_units * NSUB;     // sub-tiles of this CTA

  // TMEM map (columns): [0,256) dBias accumulator | buffer b at 256+64b: S^T [0,32) dP^T [32,64) |
  //                     384.. dQ (2 x 32) | 448.. dV | 480.. dK
  constexpr uint32_t COL_BUF = 256, COL_DQ = 384, COL_DV = 448, COL_DK = 480;

  if (threadIdx.x == 0) {
    if ((tc::smem_u32(smem) & 1023u) != 0) {
      printf("vsn_b200: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&ld_full[i], 32);
      tc::mbar_init(&ld_empty[i], 1);
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&p_ready[i], 256);
      tc::mbar_init(&ds_free[i], 1);
    }
    tc::mbar_init(unit_done, 1);
    tc::mbar_init(acc_read, 256);
    tc::mbar_init(kv_done, 1);
    tc::mbar_init(dq_read, 256);
    tc::fence_barrier_init();
  }
  if (warp == 17) tc::tmem_alloc(tmem_slot, 512);
  build_bias_t<WD, WH, WW>(p, head, bias_s);
  // zero everything that the loader never writes: pad rows of the tiles, the identity tile, lse/delta pads
  for (int i = threadIdx.x; i < 2 * BWD_STAGE_BYTES / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(stages)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(ident)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (threadIdx.x < 16)   // I16 as a K-major [16 rows][32] tile (64-byte swizzle), element (n, n) = 1
    *reinterpret_cast<bf16*>(ident + swz64(threadIdx.x, threadIdx.x >> 3) + (threadIdx.x & 7) * 2) = f2b(1.0f);
  for (int i = threadIdx.x; i < 2 * (NP - N); i += blockDim.x) {
    const int st = i / (NP - N), q = N + i % (NP - N);
    reinterpret_cast<float*>(stages + st * BWD_STAGE_BYTES + 49152)[q] = 30000.f;   // lse2 pad -> P = 0
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // Register budget as in the forward kernel: 96 per thread at launch, the helper warpgroup drops to 56 and the
  // compute warpgroups grow to 104 (each role inside the branch of its own setmaxnreg).
  if (warp >= 16) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 16) {
    // ------------------------------------------------------------------ loader
    const long long ld = 3LL * p.C;
    for (int n = 0; n < n_units; ++n) {
      const int s = blockIdx.x + n * gridDim.x;
      const int st = n & 1;
      tc::mbar_wait_relaxed(&ld_empty[st], ((n >> 1) & 1) ^ 1);
      uint8_t* sb = stages + st * BWD_STAGE_BYTES;
      const WinCoord wc = win_coord(p, s, WD, WH, WW);
      // the window's geometry is published with the tiles: the compute threads do no index math (their per-window
      // win_coord + token_geom divisions were ~550 clk per sub-tile and group, profiles/r2_wattn_bwd_notes.md)
      if (lane == 0) *reinterpret_cast<int*>(sb + BWD_OFF_FLAG) = wc.masked() ? 1 : 0;
      for (int i = lane; i < N; i += 32) {
        const TokenGeom g = token_geom_w<WD, WH, WW>(p, wc, i);
        sb[BWD_OFF_CODE + i] = static_cast<uint8_t>(g.code);
        reinterpret_cast<int*>(sb + BWD_OFF_ROW)[i] = g.row;
        const bf16* src = p.qkv + g.row * ld + head * HD;
        const bf16* dsrc = p.dout + static_cast<long long>(g.row) * p.C + head * HD;
        const int kr = i - kh * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t o = swz64(i, c);
          tc::cp_async16(sb + o, src + c * 8);
          tc::cp_async16(sb + 16384 + o, dsrc + c * 8);
          if (kr >= 0 && kr < 128) {
            const uint32_t ok = swz64(kr, c);
            tc::cp_async16(sb + 32768 + ok, src + p.C + c * 8);
            tc::cp_async16(sb + 40960 + ok, src + 2 * p.C + c * 8);
          }
        }
      }
      const float* lse_g = p.lse + (static_cast<long long>(s) * p.heads + head) * NP;
      const float* dl_g = delta_g + (static_cast<long long>(s) * p.heads + head) * NP;
      for (int i = lane; i < N / 4; i += 32) {
        tc::cp_async16(sb + 49152 + i * 16, lse_g + i * 4);
        tc::cp_async16(sb + 50176 + i * 16, dl_g + i * 4);
      }
      tc::cp_async_wait_all();
      tc::fence_proxy_async();
      tc::mbar_arrive(&ld_full[st]);
    }
  } else if (warp == 17) {
    // ------------------------------------------------------------------ MMA issuer (warp runs converged)
    {
      const uint32_t idesc_s = tc::make_idesc_bf16(128, QS, 0, 0);     // S^T, dP^T: A K-major, B K-major
      const uint32_t idesc_kv = tc::make_idesc_bf16(128, HD, 0, 1);    // dV, dK: A TMEM, B MN-major
      const uint32_t idesc_b = tc::make_idesc_bf16(128, 16, 0, 0);     // dBias: A TMEM, B = I16
      const uint32_t idesc_q = tc::make_idesc_bf16(128, HD, 1, 1);     // dQ: A MN-major smem, B MN-major
      const uint64_t desc_ident = tc::make_smem_desc_sw64(tc::smem_u32(ident), 16, 512);
      const uint64_t desc_ds = tc::make_smem_desc_sw128(tc::smem_u32(ds_tile), 16384, 1024);
      const uint64_t desc_st0 = tc::make_smem_desc_sw64(tc::smem_u32(stages), 16, 512);   // stage 0, offset 0
      long long tm_ld = 0, tm_pr = 0, tm_acc = 0, tm_t0 = TCLK(), tq;
      int pend_mb = -1, pend_st = 0;
      uint64_t pend_dst = 0;
      auto issue_dq = [&]() {
        if (pend_mb < 0) return;
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tc::mma_bf16_ss(tmem_base + COL_DQ + pend_mb * 32, tc::desc_advance(desc_ds, pend_mb * 32768 + k * 2048),
                            tc::desc_advance(pend_dst, 32768 + k * 1024), idesc_q, k);
          tc::mma_commit(&ds_free[pend_mb]);
          if (pend_mb == 1) {
            tc::mma_commit(unit_done);
            tc::mma_commit(&ld_empty[pend_st]);
          }
        }
        __syncwarp();
        pend_mb = -1;
      };
      for (int T = 0; T <= TT; ++T) {
        if (T < TT) {
          const int n = T / NSUB, t = T % NSUB, st = n & 1, b = T & 1;
          tq = TCLK();
          if (t == 0) tc::mbar_wait(&ld_full[st], (n >> 1) & 1);
          tm_ld += TCLK() - tq;
          tc::fence_after_sync();
          const uint64_t dst = tc::desc_advance(desc_st0, st * BWD_STAGE_BYTES);
          const uint32_t d = tmem_base + COL_BUF + b * 64;
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              tc::mma_bf16_ss(d, tc::desc_advance(dst, 32768 + k * 32), tc::desc_advance(dst, t * (QS * 64) + k * 32),
                              idesc_s, k);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              tc::mma_bf16_ss(d + 32, tc::desc_advance(dst, 40960 + k * 32),
                              tc::desc_advance(dst, 16384 + t * (QS * 64) + k * 32), idesc_s, k);
            tc::mma_commit(&s_full[b]);
          }
          __syncwarp();
        }
        issue_dq();
        if (T >= 1) {
          const int V = T - 1, n = V / NSUB, t = V % NSUB, st = n & 1, b = V & 1;
          tq = TCLK();
          tc::mbar_wait(&p_ready[b], (V >> 1) & 1);
          tm_pr += TCLK() - tq;
          tq = TCLK();
          if (t == 0) tc::mbar_wait(acc_read, (n & 1) ^ 1);       // previous window's dK / dV were read out
          if (t == 3) tc::mbar_wait(dq_read, (n & 1) ^ 1);        // previous window's dQ was read out
          tm_acc += TCLK() - tq;
          tc::fence_after_sync();
          const uint64_t dst = tc::desc_advance(desc_st0, st * BWD_STAGE_BYTES);
          const uint32_t a_p = tmem_base + COL_BUF + b * 64;        // P^T  (bf16 pairs: 8 columns per 16 queries, at 16*half)
          const uint32_t a_ds = a_p + 32;                           // dS^T
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint32_t rows = (t * QS + k * 16) * 64;
              tc::mma_bf16_ts(tmem_base + COL_DV, a_p + k * 16, tc::desc_advance(dst, 16384 + rows), idesc_kv,
                              (t | k) ? 1u : 0u);
              tc::mma_bf16_ts(tmem_base + COL_DK, a_ds + k * 16, tc::desc_advance(dst, rows), idesc_kv, (t | k) ? 1u : 0u);
              tc::mma_bf16_ts(tmem_base + t * QS + k * 16, a_ds + k * 16, desc_ident, idesc_b, n ? 1u : 0u);
            }
            if (t == NSUB - 1) tc::mma_commit(kv_done);
          }
          __syncwarp();
          // dQ for the 128-query block that is now complete in the smem tile: eight K=16 steps that nobody waits for
          // soon -- issued in the next iteration, behind that sub-tile's S^T / dP^T (the tensor pipe runs in order)
          if ((t & 3) == 3) { pend_mb = t >> 2; pend_dst = dst; pend_st = st; }
        }
      }
      issue_dq();
      if (g_bwd_timing != nullptr && lane == 0) {
        unsigned long long* o = g_bwd_timing + ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16;
        o[0] = TCLK() - tm_t0; o[1] = tm_ld; o[2] = tm_pr; o[3] = tm_acc; o[4] = TT;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ------------------------------------------------------------------ compute groups
    const int g = warp >> 3, q4 = warp & 3, half = (warp >> 2) & 1;
    const int r = q4 * 32 + lane;                  // key row inside the half = TMEM lane
    const int j = kh * 128 + r;                    // key token of the window
    const int jb = j < N ? j : N - 1;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const float cscale = p.scale * LOG2E;
    int rowbaseT;
    {
      const int dj = jb / (WH * WW), lr = jb - dj * (WH * WW), hj = lr / WW, wj = lr - hj * WW;
      rowbaseT = ((WD - 1 - dj) * (2 * WH - 1) + (WH - 1 - hj)) * WW + wj;
    }
    // smem dS tile: element (query q, key r) at (q/64)*16384 + (r/8)*1024 + (r%8)*128 + (((q%64)/8) ^ (r%8))*16
    const uint32_t ds_row = (r >> 3) * 1024 + (r & 7) * 128;
    bool masked = false;
    uint32_t ck4 = 0;
    long long tc_wait = 0, tc_ld = 0, tc_math = 0, tc_st = 0, tc_ro = 0, tc_q;
    // source rows of this thread's key and of the query row it reads out (group 1), current and previous window
    int krow = 0, qrow = 0, krow_prev = 0, qrow_prev = 0;

    // Read-out of window rn (group 1 only), run one sub-tile into the NEXT window so that kv_done / unit_done have
    // long fired and nothing here waits.
    auto readout = [&](int rn, int rk, int rq) {
        // ---- window read-out.  dV / dK rows of this key half: direct stores, released as soon as they are in
        // registers (the next window's first MMA overwrites them); dQ partial: bf16 red.add (REDG.BF16x8) into the
        // zeroed Q block, needed back only at the next window's 4th sub-tile.
        uint32_t acc[32];
        tc::mbar_wait(kv_done, rn & 1);
        tc::fence_after_sync();
        tc::tmem_ld_32x32b_x32(tmem_base + lane_addr + (half ? COL_DK : COL_DV), acc);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        tc::mbar_arrive(acc_read);
        if (j < N) {
          const float sc = half ? p.scale : 1.f;
          bf16* dst = p.dqkv + static_cast<long long>(rk) * (3LL * p.C) + (half ? p.C : 2 * p.C) + head * HD;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(acc[8 * c]) * sc, __uint_as_float(acc[8 * c + 1]) * sc);
            w.y = pack_bf16(__uint_as_float(acc[8 * c + 2]) * sc, __uint_as_float(acc[8 * c + 3]) * sc);
            w.z = pack_bf16(__uint_as_float(acc[8 * c + 4]) * sc, __uint_as_float(acc[8 * c + 5]) * sc);
            w.w = pack_bf16(__uint_as_float(acc[8 * c + 6]) * sc, __uint_as_float(acc[8 * c + 7]) * sc);
            reinterpret_cast<uint4*>(dst)[c] = w;
          }
        }
        const int qi = half * 128 + r;      // query row of the dQ block this thread reads
        tc::mbar_wait(unit_done, rn & 1);
        tc::fence_after_sync();
        tc::tmem_ld_32x32b_x32(tmem_base + lane_addr + COL_DQ + half * 32, acc);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        tc::mbar_arrive(dq_read);
        if (qi < N) {
          bf16* dst = p.dqkv + static_cast<long long>(rq) * (3LL * p.C) + head * HD;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            red_add_bf16x8(dst + 8 * c,
                           pack_bf16(__uint_as_float(acc[8 * c]) * p.scale, __uint_as_float(acc[8 * c + 1]) * p.scale),
                           pack_bf16(__uint_as_float(acc[8 * c + 2]) * p.scale, __uint_as_float(acc[8 * c + 3]) * p.scale),
                           pack_bf16(__uint_as_float(acc[8 * c + 4]) * p.scale, __uint_as_float(acc[8 * c + 5]) * p.scale),
                           pack_bf16(__uint_as_float(acc[8 * c + 6]) * p.scale, __uint_as_float(acc[8 * c + 7]) * p.scale));
        }
    };

    for (int T = g; T < TT; T += 2) {
      const int n = T / NSUB, t = T % NSUB, st = n & 1;
      const uint8_t* sb = stages + st * BWD_STAGE_BYTES;
      const int q0 = t * QS + half * 16;

      tc_q = TCLK();
      tc::mbar_wait(&s_full[g], (T >> 1) & 1);
      tc::fence_after_sync();
      tc_wait += TCLK() - tc_q; tc_q = TCLK();
      if (t < 2) {                                   // first sub-tile of this group in the window: the loader's geometry
        masked = *reinterpret_cast<const int*>(sb + BWD_OFF_FLAG) != 0;      // (s_full implies the stage's ld_full)
        ck4 = static_cast<uint32_t>(sb[BWD_OFF_CODE + jb]) * 0x01010101u;
        if (g == 1) {
          krow_prev = krow; qrow_prev = qrow;
          krow = reinterpret_cast<const int*>(sb + BWD_OFF_ROW)[jb];
          qrow = reinterpret_cast<const int*>(sb + BWD_OFF_ROW)[min(half * 128 + r, N - 1)];
        }
      }
      float x[16], dp[16];
      {
        uint32_t u0[16], u1[16];
        const uint32_t base = tmem_base + lane_addr + COL_BUF + g * 64 + half * 16;
        tc::tmem_ld_32x32b_x16(base, u0);
        tc::tmem_ld_32x32b_x16(base + 32, u1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) { x[k] = __uint_as_float(u0[k]); dp[k] = __uint_as_float(u1[k]); }
      }
      tc_ld += TCLK() - tc_q; tc_q = TCLK();
      const int qr0 = q0 / WW;
      switch (q0 % WW) {
        case 0: add_bias_t16<WD, WH, WW, 0>(x, bias_s, rowbaseT, qr0, cscale); break;
        case 2: add_bias_t16<WD, WH, WW, 2>(x, bias_s, rowbaseT, qr0, cscale); break;
        default: add_bias_t16<WD, WH, WW, 4>(x, bias_s, rowbaseT, qr0, cscale); break;
      }
      if (masked) {
        const uint4 qc = *reinterpret_cast<const uint4*>(sb + BWD_OFF_CODE + q0);
        const uint32_t qw[4] = {qc.x, qc.y, qc.z, qc.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint32_t ne = __vcmpne4(qw[w], ck4);
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const uint32_t m32 = __byte_perm(ne, 0, 0x8888u | (0x1111u * b));
            x[w * 4 + b] += __uint_as_float(m32 & __float_as_uint(MASK_L2E));
          }
        }
      }
      uint32_t pk[8], dk[8];
      {
        const float4* l4 = reinterpret_cast<const float4*>(sb + 49152 + q0 * 4);
        const float4* d4 = reinterpret_cast<const float4*>(sb + 50176 + q0 * 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 l = l4[c], dl = d4[c];
          const float p0 = tc::ex2_approx(x[4 * c] - l.x), p1 = tc::ex2_approx(x[4 * c + 1] - l.y);
          const float p2 = tc::ex2_approx(x[4 * c + 2] - l.z), p3 = tc::ex2_approx(x[4 * c + 3] - l.w);
          pk[2 * c] = pack_bf16(p0, p1);
          pk[2 * c + 1] = pack_bf16(p2, p3);
          dk[2 * c] = pack_bf16(p0 * (dp[4 * c] - dl.x), p1 * (dp[4 * c + 1] - dl.y));
          dk[2 * c + 1] = pack_bf16(p2 * (dp[4 * c + 2] - dl.z), p3 * (dp[4 * c + 3] - dl.w));
        }
      }
      tc_math += TCLK() - tc_q; tc_q = TCLK();
      // P^T / dS^T over the first 8 of this thread's own 16 columns (bf16 pairs): TMEM A operands
      const uint32_t pb = tmem_base + lane_addr + COL_BUF + g * 64 + half * 16;
      tc::tmem_st_32x32b_x8(pb, pk);
      tc::tmem_st_32x32b_x8(pb + 32, dk);
      // dS^T copy for dQ; the previous window's dQ MMAs of this query block must be done with the tile
      if ((t & 3) < 2) tc::mbar_wait(&ds_free[t >> 2], (n & 1) ^ 1);
      {
        uint8_t* dst = ds_tile + (q0 >> 6) * 16384 + ds_row;
        const int c0 = (q0 & 63) >> 3;
        *reinterpret_cast<uint4*>(dst + ((c0 ^ (r & 7)) << 4)) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
        *reinterpret_cast<uint4*>(dst + (((c0 + 1) ^ (r & 7)) << 4)) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
      }
      tc::fence_proxy_async();
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc::mbar_arrive(&p_ready[g]);
      tc_st += TCLK() - tc_q; tc_q = TCLK();

      if (g == 1 && t == 1 && n > 0) {
        readout(n - 1, krow_prev, qrow_prev);
        tc_ro += TCLK() - tc_q;
      }
    }
    if (g == 1 && n_units > 0) readout(n_units - 1, krow, qrow);
    if (g_bwd_timing != nullptr && lane == 0 && q4 == 0 && half == 0) {
      unsigned long long* o = g_bwd_timing + ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + 5 + g * 5;
      o[0] = tc_wait; o[1] = tc_ld; o[2] = tc_math; o[3] = tc_st; o[4] = tc_ro;
    }
    // ---- CTA end: fold this CTA's dense bias gradient [128 keys x 256 queries] (TMEM) onto the
    // relative_position_bias_table: dT[rpi(i,j)] += dB[j][i], rpi(i,j) = lin(i) - lin(j) + off
    // (models/swin_transformer_3d.py:132-152,186-190).  Shared-memory fp32 atomics into a 1573-entry table, then
    // one global atomic per table entry and CTA -- instead of 32768 global atomics plus a separate reduction.
    if (p.dtable != nullptr && TT > 0) {
      constexpr int TL = (2 * WD - 1) * (2 * WH - 1) * (2 * WW - 1);
      constexpr int OFF = (WD - 1) * (2 * WH - 1) * (2 * WW - 1) + (WH - 1) * (2 * WW - 1) + (WW - 1);
      float* tab_s = reinterpret_cast<float*>(ds_tile);            // the dS tile is free now
      short* lin_s = reinterpret_cast<short*>(ds_tile + 8192);
      const int ct = threadIdx.x;                                   // 0..511 (compute threads)
      // all MMAs of the CTA are complete once the last unit_done has fired
      tc::mbar_wait(unit_done, (n_units - 1) & 1);
      tc::fence_after_sync();
      for (int e = ct; e < TL; e += 512) tab_s[e] = 0.f;
      if (ct < NP) {
        const int ii = ct < N ? ct : N - 1;
        const int di = ii / (WH * WW), lr = ii - di * (WH * WW), hi = lr / WW, wi = lr - hi * WW;
        lin_s[ct] = static_cast<short>(di * ((2 * WH - 1) * (2 * WW - 1)) + hi * (2 * WW - 1) + wi);
      }
      tc::named_bar_sync(2, 512);
      const int lin_j = lin_s[jb];
      const int slice = (g * 2 + half) * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t acc[32];      // the TMEM load is warp-collective: never under a divergent branch
        tc::tmem_ld_32x32b_x32(tmem_base + lane_addr + slice + c * 32, acc);
        tc::tmem_ld_wait();
        if (j < N) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int i = slice + c * 32 + k;
            if (i < N) atomicAdd(&tab_s[lin_s[i] - lin_j + OFF], __uint_as_float(acc[k]));
          }
        }
      }
      tc::named_bar_sync(2, 512);
      for (int e = ct; e < TL; e += 512) atomicAdd(p.dtable + static_cast<long long>(e) * p.heads + head, tab_s[e]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 17) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// delta[s, head, i] = sum_e O[row(i), head*32+e] * dO[row(i), head*32+e]  (rowsum(dO * O) of the softmax backward);
// also zeroes the Q block of dqkv
template <int WD, int WH, int WW>
__global__ void __launch_bounds__(256) wattn_delta_kernel(const WinAttnArgs p, float* __restrict__ delta) {
  pdl_trigger();   // the next kernel of the stream may be scheduled (it waits for this grid before reading)
  constexpr int N = WD * WH * WW;
  const int s = blockIdx.x;
  const WinCoord wc = win_coord(p, s, WD, WH, WW);
  // blockIdx.y takes a slice of the heads so that launches with few windows (late stages) still fill the chip
  const int hp = (p.heads + gridDim.y - 1) / gridDim.y, h0 = blockIdx.y * hp;
  const int nh = min(hp, p.heads - h0);
  for (int item = threadIdx.x; item < N * nh; item += blockDim.x) {
    const int head = h0 + item % nh, i = item / nh;
    const TokenGeom g = token_geom_w<WD, WH, WW>(p, wc, i);
    const uint4* o = reinterpret_cast<const uint4*>(p.out + static_cast<long long>(g.row) * p.C + head * HD);
    const uint4* d = reinterpret_cast<const uint4*>(p.dout + static_cast<long long>(g.row) * p.C + head * HD);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 a = o[c], b = d[c];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 x = unpack_bf16(aw[t]), y = unpack_bf16(bw[t]);
        acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
      }
    }
    delta[(static_cast<long long>(s) * p.heads + head) * NP + i] = acc;
    // the Q block of dqkv is accumulated with red.add by the two key-half CTAs of the backward kernel: zero it here
    // (every (token, head) is visited exactly once), which replaces a strided 2-D memset
    uint4* zq = reinterpret_cast<uint4*>(p.dqkv + static_cast<long long>(g.row) * (3LL * p.C) + head * HD);
#pragma unroll
    for (int c = 0; c < 4; ++c) zq[c] = make_uint4(0, 0, 0, 0);
  }
}

int wattn_tc_bwd(const WinAttnArgs& a, float* delta, cudaStream_t stream) {
  using SM = BwdSmem<6, 7, 6>;
  static bool attr_set = false;
  if (!attr_set) {
    VSN_CUDA(cudaFuncSetAttribute(wattn_bwd_kernel<6, 7, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    attr_set = true;
  }
  {
    int ys = ceil_div(2 * vsn_num_sms(), a.S);
    if (ys > a.heads) ys = a.heads;
    if (ys < 1) ys = 1;
    ys = ceil_div(a.heads, ceil_div(a.heads, ys));      // no empty slices
    wattn_delta_kernel<6, 7, 6><<<dim3(a.S, ys), 256, 0, stream>>>(a, delta);
  }
  VSN_LAUNCH_CHECK();
  int groups = vsn_num_sms() / (2 * a.heads);
  if (groups < 1) groups = 1;
  if (groups > a.S) groups = a.S;
  groups = ceil_div(a.S, ceil_div(a.S, groups));
  dim3 grid(groups, a.heads, 2);
  static int timing = -1;
  static unsigned long long* tbuf = nullptr;
  if (timing < 0) {
    const char* e = getenv("VSN_WATTN_TIMING");
    timing = (e != nullptr && e[0] == '1') ? 1 : 0;
    if (timing) {
      VSN_CUDA(cudaMalloc(&tbuf, 1024 * 16 * 8));
      VSN_CUDA(cudaMemcpyToSymbol(g_bwd_timing, &tbuf, sizeof(tbuf)));
    }
  }
  wattn_bwd_kernel<6, 7, 6><<<grid, BWD_THREADS, SM::TOTAL, stream>>>(a, delta);
  VSN_LAUNCH_CHECK();
  if (timing) {
    const int n = grid.x * grid.y * grid.z;
    static unsigned long long host[1024 * 16];
    VSN_CUDA(cudaStreamSynchronize(stream));
    VSN_CUDA(cudaMemcpy(host, tbuf, n * 16 * 8, cudaMemcpyDeviceToHost));
    double acc[16] = {0};
    for (int i = 0; i < n; ++i) for (int k = 0; k < 16; ++k) acc[k] += static_cast<double>(host[i * 16 + k]) / n;
    fprintf(stderr, "wattn_bwd timing (mean cycles per CTA over %d CTAs, %.0f sub-tiles): MMA warp total %.0f | wait ld_full %.0f "
            "| wait p_ready %.0f | wait acc_read %.0f || group0: wait s_full %.0f ldtm %.0f math %.0f store+arrive %.0f || "
            "group1: wait s_full %.0f ldtm %.0f math %.0f store+arrive %.0f readout %.0f\n", n, acc[4], acc[0], acc[1], acc[2], acc[3],
            acc[5], acc[6], acc[7], acc[8], acc[10], acc[11], acc[12], acc[13], acc[14]);
  }
  return 0;
}
