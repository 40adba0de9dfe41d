// bf16 GEMM on the 5th-gen tensor cores: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators in
// TMEM, two buffers) -> tcgen05.ld -> fused epilogue.  Persistent CTAs (one per SM, or one CTA pair per two SMs
// for cta_group::2 tiles of 256 rows), 128 x BN tiles, warp-specialised:
//   warp 0    TMA producer (one lane), runs ahead across tile boundaries
//   warp 1    TMEM allocator + MMA issuer (converged warp, elect.sync)
//   warps 2.. 8 or 16 epilogue warps (each owns the 32 TMEM lanes its warp id maps to and a share of the columns)
// Launched with programmatic dependent launch: everything up to the first operand load overlaps the previous kernel.
//
// Operand majors cover all three GEMMs of a Linear layer without materialising transposes:
//   forward  Y[M,N]  = X[M,K]  * W[N,K]^T     A K-major,  B K-major
//   dgrad    dX[M,K] = dY[M,N] * W[N,K]       A K-major,  B MN-major (W read as stored)
//   wgrad    dW[N,K] = dY[M,N]^T * X[M,K]     A MN-major, B MN-major, split over tokens, fp32 red.add
// Replaces the cuBLAS calls behind nn.Linear / F.linear in the reference
// (models/swin_transformer_3d.py:52-69,154-156,550; models/vit_3d.py:59-75,102-105,372).
#include <cuda.h>
#include <stdlib.h>
#include "tc.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int A_BYTES = BM * BK * 2;

struct GemmArgs {
  int M, N, K;          // output is M x N, reduction length K
  int a_mn, b_mn;       // operand majors (0 = K-major, 1 = MN-major)
  int kb_per_split;     // k-blocks handled by one split
  int splits;           // K splits per output tile (fp32 atomic accumulation when > 1)
  int total_tiles;      // tiles_m * tiles_n * splits
  int stationary;       // 1: CTA b keeps column (n-tile, split) = b % cols for all its m-tiles (cols divides the grid)
  int m_units;          // row units of the schedule: m-tiles, or pairs of m-tiles for CTA pairs
  void* out;
  long long ldo;
  int out_kind;         // 0 bf16 store, 1 fp32 store, 2 fp32 atomic add
  const float* bias;    // [N] or null
  int act;              // 0 none, 1 GELU (pre-activation saved to aux), 2 multiply by GELU'(aux)
  bf16* aux;
  long long ldaux;
  const float* resid;   // fp32 [M, ldr] or null: out = resid + scale * (acc + bias)
  long long ldr;
  const float* row_scale;  // per row-group scale (DropPath keep/(1-p)); null = 1
  int rows_per_group;
  float alpha;
  int wide;             // 1: out / aux / resid rows are 32-byte aligned -> 256-bit accesses
  int gelu_fp32;        // bit 0: GELU in fp32 (default: the forward activation feeds every later layer and the
                        // model-level gradient parity needs it), bit 1: GELU' in fp32 (default: packed half2)
  float* rowsum;        // [M] fp32 or null: rowsum[m] += sum_k A(m,k) (bias gradient of a Linear in its wgrad GEMM)
};

// TWO = CTA pair: a CTA holds its 128 rows of A and half of the B tile, so the same smem buys a deeper ring
// AUX = number of [128 x BN] tiles of the epilogue's second operand staged in shared memory by TMA (MODE 5: the
// saved pre-activation of the GELU' epilogue, double-buffered); they come out of the operand ring's budget.
template <int BN, bool TWO = false, int AUX = 0>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;                       // whole B tile
  static constexpr int B_CTA_BYTES = TWO ? B_BYTES / 2 : B_BYTES;    // what one CTA stores
  static constexpr int STAGE_BYTES = A_BYTES + B_CTA_BYTES;
  static constexpr int AUX_TILE_BYTES = BM * BN * 2;
  static constexpr int AUX_BYTES = AUX * AUX_TILE_BYTES;
  static constexpr int RING_BUDGET = 192 * 1024 - AUX_BYTES;
  static constexpr int STAGES = RING_BUDGET / STAGE_BYTES > 8 ? 8 : RING_BUDGET / STAGE_BYTES;
  static constexpr int MISC = STAGES * STAGE_BYTES + AUX_BYTES;     // barriers, bias slice and ones tile start here
  static constexpr int RS_COL = 2 * BN;              // two 16-column row-sum accumulators after the two tiles
  static constexpr int TMEM_COLS = 2 * BN + 32 <= 128 ? 128 : 2 * BN + 32 <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = MISC + 1024 /*align slack*/ + 256 /*barriers*/ + 2048 /*bias*/ +
                                    2048 /*ones tile*/;
  static_assert(AUX == 0 || (STAGES >= 2 && BN % 64 == 0), "staged aux tiles: 64-column boxes, at least two ring stages");
  static_assert(STAGE_BYTES % 1024 == 0, "stage bases must keep the 1024-byte alignment of the 128B swizzle");
};


struct TileCoord { int m0, n0, kb0, nkb; };

// Walks the tiles of this CTA.  Stationary schedule: the CTA keeps its (n-tile, split) column and walks the row
// units with stride grid / columns, so the weight tile and the bias slice stay the same for the whole kernel;
// otherwise tiles (split fastest, then n, then rows) are dealt round-robin over the grid.
// TWO (CTA pair, cta_group::2): the schedule runs over 256-row units, rows [0,128) of a unit belong to the leader
// CTA (rank 0), rows [128,256) to its peer; the lower half of the last unit may lie below the matrix (TMA fills
// zeros, the epilogue skips the rows).
// All integer divisions happen once, in the constructor: the small-K problems of the early stages run tens of tiles
// per CTA with little work each, and the per-tile div/mod chains were 40 % of their executed instructions.
template <bool TWO>
struct TileIter {
  int m, n, split;          // current row unit, n-tile, split
  int dm, dn, ds;           // grid stride decomposed over (rows, n-tiles, splits)
  int tiles_n, splits, m_units, bn, rank, kbps, nkb_total;
  bool first;
  __device__ __forceinline__ TileIter(const GemmArgs& p, int tiles_n_, int bn_, int rank_) {
    const int vb = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int vg = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    tiles_n = tiles_n_; splits = p.splits; m_units = p.m_units; bn = bn_; rank = rank_;
    kbps = p.kb_per_split; nkb_total = (p.K + BK - 1) / BK;
    const int cols = tiles_n * splits;
    int start, stride;
    if (p.stationary) {              // the CTA's column never changes: only the row unit advances
      start = (vb / cols) * cols + vb % cols;
      stride = (vg / cols) * cols;
    } else {
      start = vb;
      stride = vg;
    }
    split = start % splits; n = (start / splits) % tiles_n; m = start / cols;
    ds = stride % splits; dn = (stride / splits) % tiles_n; dm = stride / cols;
    first = true;
  }
  __device__ __forceinline__ bool next(TileCoord& t) {
    if (!first) {
      split += ds;
      int c = split >= splits ? 1 : 0;
      split -= c * splits;
      n += dn + c;
      c = n >= tiles_n ? 1 : 0;
      n -= c * tiles_n;
      m += dm + c;
    }
    first = false;
    if (m >= m_units) return false;
    t.n0 = n * bn;
    t.m0 = (m * (TWO ? 2 : 1) + rank) * BM;
    t.kb0 = split * kbps;
    t.nkb = min(nkb_total, t.kb0 + kbps) - t.kb0;
    return true;
  }
};

// MODE specialises the hot epilogues at compile time (straight-line code, no per-chunk parameter loads and uniform
// branches); MODE 0 reads everything from the arguments:
//   1 = bias + GELU, bf16 output + saved pre-activation, 256-bit stores, full tiles (the fc1 forward)
//   2 = GELU' from the saved pre-activation, bf16 output, 256-bit accesses, full tiles (the fc2 dgrad)
//   3 = bias + residual + row scale, fp32 output, 256-bit accesses, full tiles (proj / fc2 forward)
//   4 = optional bias, bf16 output, 256-bit stores, full tiles (qkv forward, plain dgrads)
//   5 = 2 with the saved pre-activation tile staged in shared memory by TMA (`aux_row` = this thread's row of the
//       128B-swizzled [128][64] sub-tile the chunk lies in, `aux_c16` = the chunk's first 16-byte column): the epilogue
//       warps no longer wait for a dependent global load per chunk (the load and the GELU' arithmetic were additive:
//       0.080 ms stores only, +0.098 the load, +0.083 the arithmetic, 0.239 together at M435456 N384 K96)
template <int MODE_>
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& pa, int row, int n, float rs, const uint32_t* v,
                                               const float* bias_s, const uint8_t* aux_row = nullptr, int aux_c16 = 0) {
  constexpr int MODE = MODE_ == 5 ? 2 : MODE_;
  struct View {       // the fields the epilogue reads, constants where the mode fixes them
    const GemmArgs& a;
    __device__ __forceinline__ int act() const { return MODE == 1 ? 1 : MODE == 2 ? 2 : MODE >= 3 ? 0 : a.act; }
    __device__ __forceinline__ int out_kind() const { return MODE == 3 ? 1 : MODE >= 1 ? 0 : a.out_kind; }
    __device__ __forceinline__ bool wide() const { return MODE >= 1 ? true : a.wide != 0; }
    __device__ __forceinline__ bool has_bias() const { return (MODE == 1 || MODE == 3) ? true : MODE == 2 ? false : a.bias != nullptr; }
    __device__ __forceinline__ bool has_resid() const { return MODE == 3 ? true : MODE >= 1 ? false : a.resid != nullptr; }
  };
  const View m{pa};
  const GemmArgs& p = pa;
  const int nvalid = MODE >= 1 ? 32 : min(32, p.N - n);
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  if (m.has_bias()) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j);   // broadcast read
      f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
    }
  }
  if (m.act() == 1) {
    // GELU is evaluated on the bf16-rounded pre-activation so backward (which only has the saved bf16
    // value) differentiates exactly the function forward applied; the rounding reuses the packed pairs.
    bf16* ap = p.aux + static_cast<long long>(row) * p.ldaux + n;
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
    if (nvalid == 32 && m.wide()) {
      st_global_v8(ap, pk);
      st_global_v8(ap + 16, pk + 8);
    } else if (nvalid == 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(ap + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { ap[j] = f2b(f[j]); }
    }
    if (p.gelu_fp32 & 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        f[2 * j] = gelu_erf<false>(bf16lo(pk[j]));
        f[2 * j + 1] = gelu_erf<false>(bf16hi(pk[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 y = gelu_erf_bf16x2(pk[j]);
        f[2 * j] = y.x;
        f[2 * j + 1] = y.y;
      }
    }
  } else if (m.act() == 2) {
    const bf16* ap = p.aux + static_cast<long long>(row) * p.ldaux + n;
    if (nvalid == 32) {
      uint32_t w[16];
      if constexpr (MODE_ == 5) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = *reinterpret_cast<const uint4*>(aux_row + (((aux_c16 + j) ^ (row & 7)) << 4));
          w[4 * j] = u.x; w[4 * j + 1] = u.y; w[4 * j + 2] = u.z; w[4 * j + 3] = u.w;
        }
      } else if (p.gelu_fp32 & 4) {     // measurement only (VSN_GELU_FP32 bit 2): no aux read, wrong results
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = 0x3C003C00u + row;
      } else if (m.wide()) {
        ld_global_v8(ap, w);
        ld_global_v8(ap + 16, w + 8);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = *reinterpret_cast<const uint4*>(ap + 8 * j);
          w[4 * j] = u.x; w[4 * j + 1] = u.y; w[4 * j + 2] = u.z; w[4 * j + 3] = u.w;
        }
      }
      if (p.gelu_fp32 & 8) {            // measurement only (bit 3): aux read but no GELU' arithmetic, wrong results
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          f[2 * t] *= bf16lo(w[t]);
          f[2 * t + 1] *= bf16hi(w[t]);
        }
      } else if (p.gelu_fp32 & 2) {
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          f[2 * t] *= gelu_erf_grad<false>(bf16lo(w[t]));
          f[2 * t + 1] *= gelu_erf_grad<true>(bf16hi(w[t]));
        }
      } else {
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float2 gg = gelu_erf_grad_bf16x2(w[t]);
          f[2 * t] *= gg.x;
          f[2 * t + 1] *= gg.y;
        }
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { f[j] *= gelu_erf_grad<false>(b2f(ap[j])); }
    }
  }
  if (rs != 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= rs;
  }
  if (m.has_resid()) {
    const float* rp = p.resid + static_cast<long long>(row) * p.ldr + n;
    if (nvalid == 32 && m.wide()) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t r8[8];
        ld_global_v8(rp + j, r8);
#pragma unroll
        for (int t = 0; t < 8; ++t) f[j + t] += __uint_as_float(r8[t]);
      }
    } else if (nvalid == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
        f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { f[j] += rp[j]; }
    }
  }
  if (m.out_kind() == 0) {
    bf16* op = reinterpret_cast<bf16*>(p.out) + static_cast<long long>(row) * p.ldo + n;
    if (nvalid == 32 && m.wide()) {
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
      st_global_v8(op, o);
      st_global_v8(op + 16, o + 8);
    } else if (nvalid == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        u.x = pack_bf16(f[j], f[j + 1]); u.y = pack_bf16(f[j + 2], f[j + 3]);
        u.z = pack_bf16(f[j + 4], f[j + 5]); u.w = pack_bf16(f[j + 6], f[j + 7]);
        *reinterpret_cast<uint4*>(op + j) = u;
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { op[j] = f2b(f[j]); }
    }
  } else if (m.out_kind() == 1) {
    float* op = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + n;
    if (nvalid == 32 && m.wide()) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = __float_as_uint(f[j + t]);
        st_global_v8(op + j, o);
      }
    } else if (nvalid == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { op[j] = f[j]; }
    }
  } else {
    // split-K partial sums: vectorised fp32 reductions (red.global.add.v4.f32)
    float* op = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + n;
    if (nvalid == 32 && (reinterpret_cast<uintptr_t>(op) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + j), "f"(f[j]), "f"(f[j + 1]),
                     "f"(f[j + 2]), "f"(f[j + 3]) : "memory");
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { atomicAdd(op + j, f[j]); }
    }
  }
}

// Persistent: CTA b works on tiles b, b + gridDim.x, ...  The TMA ring runs ahead across tile boundaries and the
// accumulator is double-buffered in TMEM, so the MMAs of tile i+1 overlap the epilogue of tile i.
// EW = number of epilogue warps (8 or 16).  Small-K problems are bound by the epilogue (one MUFU + ~25 FP32
// instructions per element for the GELU variants, the stores for the others), so they run 16 epilogue warps
// (4 per scheduler) to hide latency; large-K problems keep 8 with software-pipelined TMEM loads.
template <int BN, int EW, int MODE, bool TWO>
__global__ void __launch_bounds__(64 + EW * 32, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmAux,
                                                                  const GemmArgs p) {
  constexpr bool AUXS = MODE == 5;                 // the epilogue's aux tile is staged in shared memory by TMA
  static_assert(!AUXS || !TWO, "staged aux tiles are built for single-CTA tiles");
  using C = Cfg<BN, TWO, AUXS ? 2 : 0>;
  constexpr int EPI_THREADS = EW * 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* aux_s = smem + C::STAGES * C::STAGE_BYTES;                                   // [2][BN/64][128][64] bf16
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::MISC);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* acc_full = empty_bar + C::STAGES;     // [2] MMA commit -> epilogue
  uint64_t* acc_empty = acc_full + 2;             // [2] epilogue threads -> MMA
  uint64_t* aux_full = acc_empty + 2;             // [2] TMA -> epilogue
  uint64_t* aux_empty = aux_full + 2;             // [2] epilogue threads -> TMA producer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_empty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + C::MISC + 256);   // [2][256]
  uint8_t* ones_s = smem + C::MISC + 256 + 2048;                     // [16][64] bf16, all 1.0

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN;
  // TWO: a pair of CTAs (cluster of 2) works on 256 x BN tiles with cta_group::2 MMAs issued by the leader: every
  // CTA streams its own 128 rows of A and only HALF of the B tile, which is what the deep-K problems of the late
  // stages are bound by (L2 -> SM operand traffic).  full[] and acc_empty[] live in the leader, empty[] and
  // acc_full[] are signalled in both CTAs by multicast commits.
  const int rank = TWO ? static_cast<int>(tc::cluster_ctarank()) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], TWO ? 2 * EPI_THREADS : EPI_THREADS);
      tc::mbar_init(&aux_full[b], 1);
      tc::mbar_init(&aux_empty[b], EPI_THREADS);
    }
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    if constexpr (AUXS) tc::prefetch_tmap(&tmAux);
  }
  if (warp == 1) {
    if constexpr (TWO) tc::tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    else tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  if (p.rowsum != nullptr) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) reinterpret_cast<uint32_t*>(ones_s)[i] = VSN_ONE_PAIR;   // 1.0 pairs
    tc::fence_proxy_async();
  }
  tc::fence_before_sync();
  __syncthreads();
  if constexpr (TWO) tc::cluster_sync();    // the peer's barriers are initialised before anything targets them
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above (barriers, TMEM, descriptor prefetch) ran while the previous kernel of the stream was still
  // finishing; from here on its results are read.
  pdl_trigger();
  pdl_wait();

  // B stays resident: with the n-stationary schedule (the CTA's weight tile never changes) the ring is cut to a
  // multiple of the tile's k-blocks, so stage s always holds the same k-block of the same weight tile and B is
  // fetched on the first pass over the ring only (the weight tile re-read per output tile was a third of the
  // L2 -> SM traffic of the K = 96 problems).
  const int nkb_tile = min((p.K + BK - 1) / BK, p.kb_per_split);
  const bool b_resident = !TWO && p.stationary && p.splits == 1 && nkb_tile <= C::STAGES;
  const int ring = b_resident ? (C::STAGES / nkb_tile) * nkb_tile : C::STAGES;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;          // ring position
      uint32_t ph = 0;    // and its phase
      TileCoord t;
      TileIter<TWO> tiles(p, tiles_n, BN, rank);
      int issued = 0;
      for (int lt = 0; tiles.next(t); ++lt) {
        if constexpr (AUXS) {
          // the tile's saved pre-activation, as BN/64 swizzled [128][64] boxes, into the buffer the epilogue of two
          // tiles ago has finished reading
          const int ab = lt & 1;
          tc::mbar_wait(&aux_empty[ab], ((lt >> 1) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&aux_full[ab], C::AUX_TILE_BYTES);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tc::tma_load_2d(aux_s + ab * C::AUX_TILE_BYTES + j * 16384, &tmAux, &aux_full[ab], t.n0 + j * 64, t.m0);
        }
        for (int i = 0; i < t.nkb; ++i, ++issued, s = (s + 1 == ring ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * C::STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = (t.kb0 + i) * BK;
          if constexpr (TWO) {
            // both CTAs' bytes (2 x A + the two halves of B) land on the leader's barrier
            if (rank == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2 * A_BYTES + C::B_BYTES);
            if (!p.a_mn) {
              tc::tma_load_2d_pair(sa, &tmA, &full_bar[s], k0, t.m0);
            } else {
              tc::tma_load_2d_pair(sa, &tmA, &full_bar[s], t.m0, k0);
              tc::tma_load_2d_pair(sa + 8192, &tmA, &full_bar[s], t.m0 + 64, k0);
            }
            if (!p.b_mn) {
              tc::tma_load_2d_pair(sb, &tmB, &full_bar[s], k0, t.n0 + rank * (BN / 2));     // box {64 k, BN/2 rows}
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                tc::tma_load_2d_pair(sb + j * 8192, &tmB, &full_bar[s], t.n0 + rank * (BN / 2) + j * 64, k0);
            }
            continue;
          }
          const bool load_b = !b_resident || issued < ring;
          tc::mbar_arrive_expect_tx(&full_bar[s], load_b ? A_BYTES + C::B_BYTES : A_BYTES);
          if (!p.a_mn) {
            tc::tma_load_2d(sa, &tmA, &full_bar[s], k0, t.m0);           // box {64 k, 128 rows}
          } else {
            tc::tma_load_2d(sa, &tmA, &full_bar[s], t.m0, k0);           // box {64 m, 64 k-rows}
            tc::tma_load_2d(sa + 8192, &tmA, &full_bar[s], t.m0 + 64, k0);
          }
          if (!load_b) {
          } else if (!p.b_mn) {
            tc::tma_load_2d(sb, &tmB, &full_bar[s], k0, t.n0);           // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tc::tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], t.n0 + j * 64, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the warp runs converged (all lanes wait on the barriers), one elected lane issues
    // (CTA pairs: the leader's warp issues for both CTAs, the peer's warp only owns its TMEM allocation)
    if (!TWO || rank == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(TWO ? 2 * BM : BM, BN, p.a_mn, p.b_mn);
      const uint32_t idesc_rs = tc::make_idesc_bf16(BM, 16, p.a_mn, 0);
      const uint64_t ones_desc = tc::make_smem_desc_sw128(tc::smem_u32(ones_s), 16, 1024);
      const uint32_t a_step = p.a_mn ? 2048u : 32u;   // bytes per K=16 step
      const uint32_t b_step = p.b_mn ? 2048u : 32u;
      const uint64_t adesc0 = tc::make_smem_desc_sw128(tc::smem_u32(smem), p.a_mn ? 8192u : 16u, 1024);
      const uint64_t bdesc0 = tc::make_smem_desc_sw128(tc::smem_u32(smem) + A_BYTES, p.b_mn ? 8192u : 16u, 1024);
      int s = 0;
      uint32_t ph = 0;
      TileCoord t;
      TileIter<TWO> tiles(p, tiles_n, BN, rank);
      for (int lt = 0; tiles.next(t); ++lt) {
        const int buf = lt & 1;
        tc::mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t d = tmem_base + buf * BN;
        const bool do_rs = p.rowsum != nullptr && t.n0 == 0 && C::RS_COL + 32 <= C::TMEM_COLS;
        for (int i = 0; i < t.nkb; ++i, s = (s + 1 == ring ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
          tc::mbar_wait(&full_bar[s], ph);
          tc::fence_after_sync();
          const uint64_t ad = tc::desc_advance(adesc0, s * C::STAGE_BYTES);
          const uint64_t bd = tc::desc_advance(bdesc0, s * C::STAGE_BYTES);
          const int krem = p.K - (t.kb0 + i) * BK;
          const int ksteps = krem >= BK ? BK / 16 : (krem + 15) / 16;
          if constexpr (TWO) {
            if (tc::elect_one()) {
              for (int k = 0; k < ksteps; ++k)
                tc::mma_bf16_ss_pair(d, tc::desc_advance(ad, k * a_step), tc::desc_advance(bd, k * b_step), idesc,
                                     (i > 0 || k > 0) ? 1u : 0u);
              tc::mma_commit_pair(&empty_bar[s], 3);                       // the stage is free in both CTAs
              if (i == t.nkb - 1) tc::mma_commit_pair(&acc_full[buf], 3);  // both halves of the accumulator
            }
            __syncwarp();
            continue;
          }
          if (tc::elect_one()) {
            if (ksteps == BK / 16) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                tc::mma_bf16_ss(d, tc::desc_advance(ad, k * a_step), tc::desc_advance(bd, k * b_step), idesc,
                                (i > 0 || k > 0) ? 1u : 0u);
                // row sums of A on the tensor cores: one N=16 MMA against the all-ones tile (first n-tile only)
                if (do_rs) tc::mma_bf16_ss(tmem_base + C::RS_COL + buf * 16, tc::desc_advance(ad, k * a_step), ones_desc,
                                           idesc_rs, (i > 0 || k > 0) ? 1u : 0u);
              }
            } else {
              for (int k = 0; k < ksteps; ++k) {
                tc::mma_bf16_ss(d, tc::desc_advance(ad, k * a_step), tc::desc_advance(bd, k * b_step), idesc,
                                (i > 0 || k > 0) ? 1u : 0u);
                if (do_rs) tc::mma_bf16_ss(tmem_base + C::RS_COL + buf * 16, tc::desc_advance(ad, k * a_step), ones_desc,
                                           idesc_rs, (i > 0 || k > 0) ? 1u : 0u);
              }
            }
            tc::mma_commit(&empty_bar[s]);                       // frees the smem stage once these MMAs retire
            if (i == t.nkb - 1) tc::mma_commit(&acc_full[buf]);  // accumulator complete
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue: thread <-> output row; epilogue warp e handles the 32-column chunks c = e/4 (mod EW/4) ----
    const int lg = warp & 3;                     // TMEM lane group this warp may touch
    constexpr int NPAR = EW / 4;                 // warps per lane group
    const int par = (warp - 2) >> 2;             // which chunks of the tile
    const int et = threadIdx.x - 64;
    constexpr int NCH = (BN / 32 + NPAR - 1) / NPAR;   // chunks per warp (upper bound)
    TileCoord t;
    TileIter<TWO> tiles(p, tiles_n, BN, rank);
    for (int lt = 0; tiles.next(t); ++lt) {
      const int buf = lt & 1;
      const int row = t.m0 + lg * 32 + lane;
      // The tile's bias slice goes through smem (broadcast reads instead of dependent global loads per chunk).
      // Stationary schedule: staged once, the column never changes; otherwise once per tile into the buffer
      // whose previous readers finished two tiles ago.
      float* bs = bias_s + (p.stationary ? 0 : buf * 256);
      if ((MODE == 1 || MODE == 3 || (MODE != 2 && p.bias != nullptr)) && (lt == 0 || !p.stationary)) {
        if (et < BN) bs[et] = (t.n0 + et < p.N) ? p.bias[t.n0 + et] : 0.f;
        tc::named_bar_sync(1, EPI_THREADS);
      }
      tc::mbar_wait(&acc_full[buf], (lt >> 1) & 1);
      if constexpr (AUXS) tc::mbar_wait(&aux_full[buf], (lt >> 1) & 1);
      tc::fence_after_sync();
      const bool row_ok = row < p.M;
      const uint8_t* aux_rowp = aux_s + buf * C::AUX_TILE_BYTES + (lg * 32 + lane) * 128;   // + sub-tile * 16384
      float rs = p.alpha;
      if (p.row_scale != nullptr && row_ok) rs *= p.row_scale[row / p.rows_per_group];
      const uint32_t tb = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + buf * BN;
      if constexpr (EW == 8) {
        // software pipeline over this warp's chunks: the TMEM load of chunk i+1 is in flight while chunk i
        // goes through the epilogue math and stores
        uint32_t v[2][32];
        if (par * 32 < BN) tc::tmem_ld_32x32b_x32(tb + par * 32, v[0]);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c0 = (par + i * NPAR) * 32;
          if (c0 < BN) {
            tc::tmem_ld_wait();
            if (c0 + NPAR * 32 < BN) tc::tmem_ld_32x32b_x32(tb + c0 + NPAR * 32, v[(i + 1) & 1]);
            const int n = t.n0 + c0;
            if (row_ok && n < p.N) epilogue_chunk<(MODE == 5 ? 2 : MODE)>(p, row, n, rs, v[i & 1], bs + c0);   // (MODE 5 runs EW = 16)
          }
        }
      } else {
#pragma unroll 1
        for (int c0 = par * 32; c0 < BN; c0 += NPAR * 32) {
          uint32_t v[32];
          tc::tmem_ld_32x32b_x32(tb + c0, v);
          tc::tmem_ld_wait();
          const int n = t.n0 + c0;
          if (row_ok && n < p.N) {
            if constexpr (AUXS) epilogue_chunk<MODE>(p, row, n, rs, v, bs + c0, aux_rowp + (c0 >> 6) * 16384, (c0 & 63) >> 3);
            else epilogue_chunk<MODE>(p, row, n, rs, v, bs + c0);
          }
        }
      }
      if constexpr (AUXS) tc::mbar_arrive(&aux_empty[buf]);      // this thread has read its part of the aux tile
      if (p.rowsum != nullptr && t.n0 == 0 && par == 0 && C::RS_COL + 32 <= C::TMEM_COLS) {
        const uint32_t rsv = tc::tmem_ld_32x32b_x1(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + C::RS_COL + buf * 16);
        tc::tmem_ld_wait();
        if (row_ok) atomicAdd(p.rowsum + row, __uint_as_float(rsv));
      }
      tc::fence_before_sync();
      if constexpr (TWO) tc::mbar_arrive_cluster(tc::leader_addr(&acc_empty[buf]));
      else tc::mbar_arrive(&acc_empty[buf]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if constexpr (TWO) tc::cluster_sync();    // no CTA leaves while its peer may still signal its barriers / read its smem
  if (warp == 1) {
    tc::fence_after_sync();
    if constexpr (TWO) tc::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---- host side ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor map: dim0 = contiguous, dim1 = rows (stride ld elements), 128B swizzle, zero OOB fill.
int make_tmap_2d(CUtensorMap* map, const void* base, long long dim0, long long dim1, long long ld, int box0, int box1) {
  EncodeTiledFn fn = get_encode_fn();
  VSN_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  VSN_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "GEMM operand base must be 16-byte aligned");
  VSN_CHECK((ld * 2) % 16 == 0, "GEMM operand leading dimension must be a multiple of 8 elements (got %lld)", ld);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim0), static_cast<cuuint64_t>(dim1)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, VSN_TMAP_DTYPE, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VSN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (dims %lld x %lld ld %lld box %d x %d)",
            static_cast<int>(r), dim0, dim1, ld, box0, box1);
  return 0;
}

template <int BN, int EW, int MODE, bool TWO>
int launch_mode(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int grid, cudaStream_t stream,
                const CUtensorMap* tmAuxp = nullptr) {
  using CfgT = Cfg<BN, TWO, MODE == 5 ? 2 : 0>;
  const CUtensorMap& tmAux = tmAuxp != nullptr ? *tmAuxp : tmA;      // unused unless MODE == 5
  static bool attr_set = false;
  if (!attr_set) {
    VSN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EW, MODE, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  CfgT::SMEM_BYTES));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((TWO ? 2 : 1) * grid);
  cfg.blockDim = dim3(64 + EW * 32);
  cfg.dynamicSmemBytes = CfgT::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (vsn_pdl_enabled()) {     // programmatic dependent launch: the prologue overlaps the previous kernel's tail
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if constexpr (TWO) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  VSN_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EW, MODE, TWO>, tmA, tmB, tmAux, a));
  VSN_LAUNCH_CHECK();
  return 0;
}

// which compile-time epilogue fits this call (0 = the generic one)
int epilogue_mode(const GemmArgs& a, int bn) {
  static int off = -1;     // VSN_GEMM_GENERIC=1: always the generic epilogue (measurements only)
  if (off < 0) { const char* e = getenv("VSN_GEMM_GENERIC"); off = (e != nullptr && e[0] == '1') ? 1 : 0; }
  if (off || !a.wide || a.N % bn != 0 || a.rowsum != nullptr || a.out_kind == 2) return 0;
  if (a.act == 1 && a.out_kind == 0 && a.bias != nullptr && a.resid == nullptr) return 1;
  if (a.act == 2 && a.out_kind == 0 && a.bias == nullptr && a.resid == nullptr) return 2;
  if (a.act == 0 && a.out_kind == 1 && a.bias != nullptr && a.resid != nullptr) return 3;
  if (a.act == 0 && a.out_kind == 0 && a.resid == nullptr) return 4;
  return 0;
}

template <int BN, int EW>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmArgs a, int splits, bool pair, cudaStream_t stream) {
  a.splits = splits;
  const int tiles_m = ceil_div(a.M, BM), cols = ceil_div(a.N, BN) * splits;
  const int mul = pair ? 2 : 1;
  a.m_units = ceil_div(tiles_m, mul);
  a.total_tiles = a.m_units * cols;
  const int slots = vsn_num_sms() / mul;               // CTAs (or CTA pairs) the chip runs at once
  int grid = a.total_tiles < slots ? a.total_tiles : slots;
  a.stationary = 0;
  // (measured: pays for few wide column tiles; round-robin is better for split-K and for many / narrow columns)
  static int stat_min_bn = -1;   // VSN_GEMM_STAT_BN: smallest tile width that takes the n-stationary schedule (measurements)
  if (stat_min_bn < 0) { const char* e = getenv("VSN_GEMM_STAT_BN"); stat_min_bn = e ? atoi(e) : 96; }
  if (splits == 1 && (cols == 1 || (BN >= stat_min_bn && cols <= 4)) && a.m_units >= 2 * (slots / cols)) {
    // n-stationary persistent schedule: as many rows of `cols` CTAs as fit on the chip
    const int rows = slots / cols < a.m_units ? slots / cols : a.m_units;
    grid = rows * cols;
    a.stationary = 1;
  }
  if constexpr (EW == 16) {
    const int mode = epilogue_mode(a, BN);
    if (pair) {
      switch (mode) {
        case 1: return launch_mode<BN, EW, 1, true>(tmA, tmB, a, grid, stream);
        case 2: return launch_mode<BN, EW, 2, true>(tmA, tmB, a, grid, stream);
        case 3: return launch_mode<BN, EW, 3, true>(tmA, tmB, a, grid, stream);
        case 4: return launch_mode<BN, EW, 4, true>(tmA, tmB, a, grid, stream);
        default: return launch_mode<BN, EW, 0, true>(tmA, tmB, a, grid, stream);
      }
    }
    if constexpr (BN % 64 == 0 && BN <= 192) {
      // GELU' epilogue with the saved pre-activation staged by TMA: short reductions only (the aux tiles take half of the
      // operand ring's shared memory); VSN_GEMM_AUX_TMA=0 keeps the global loads (measurements)
      static int aux_tma = -1;
      if (aux_tma < 0) { const char* e = getenv("VSN_GEMM_AUX_TMA"); aux_tma = (e != nullptr && e[0] == '0') ? 0 : 1; }
      if (mode == 2 && aux_tma && ceil_div(a.K, BK) <= Cfg<BN, false, 2>::STAGES && a.ldaux % 8 == 0) {
        CUtensorMap tmAux;
        if (int rc = make_tmap_2d(&tmAux, a.aux, a.N, a.M, a.ldaux, 64, BM)) return rc;
        return launch_mode<BN, EW, 5, false>(tmA, tmB, a, grid, stream, &tmAux);
      }
    }
    switch (mode) {
      case 1: return launch_mode<BN, EW, 1, false>(tmA, tmB, a, grid, stream);
      case 2: return launch_mode<BN, EW, 2, false>(tmA, tmB, a, grid, stream);
      case 3: return launch_mode<BN, EW, 3, false>(tmA, tmB, a, grid, stream);
      case 4: return launch_mode<BN, EW, 4, false>(tmA, tmB, a, grid, stream);
      default: break;
    }
  }
  return launch_mode<BN, EW, 0, false>(tmA, tmB, a, grid, stream);
}

// VSN_GEMM_EW=8|16 overrides the epilogue-warp heuristic (measurements only)
int forced_ew() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VSN_GEMM_EW"); v = e ? atoi(e) : 0; }
  return v;
}

}  // namespace

// shared with the dense-attention kernels (dattn_tc.cu)
int vsn_make_tmap_2d_bf16(CUtensorMap* map, const void* base, long long dim0, long long dim1, long long ld, int box0,
                          int box1) {
  return make_tmap_2d(map, base, dim0, dim1, ld, box0, box1);
}

extern "C" int vsn_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, int M,
                             int N, int K, void* out, long long ldo, int out_kind, const float* bias, int act,
                             void* aux, long long ldaux, const float* resid, long long ldr, const float* row_scale,
                             int rows_per_group, float alpha, int split_k, float* rowsum_out, void* stream) {
  VSN_CHECK(M > 0 && N > 0 && K > 0, "vsn_gemm_bf16: empty problem %d x %d x %d", M, N, K);
  VSN_CHECK(out_kind >= 0 && out_kind <= 2, "vsn_gemm_bf16: bad out_kind %d", out_kind);
  VSN_CHECK(act == 0 || aux != nullptr, "vsn_gemm_bf16: activation modes need the aux buffer");
  VSN_CHECK(split_k <= 1 || out_kind == 2, "vsn_gemm_bf16: split_k needs atomic fp32 output");
  VSN_CHECK(out_kind != 2 || (bias == nullptr && act == 0 && resid == nullptr),
            "vsn_gemm_bf16: atomic output takes no bias/activation/residual");
  // Tile width: the widest of {256, 192, 128, 96, 64} that divides N (an MN-major B operand is loaded in
  // 64-column blocks), falling back to smaller tiles while the problem has fewer tiles than SMs.
  int BN;
  if (b_mn) BN = (N % 256 == 0) ? 256 : (N % 192 == 0) ? 192 : (N <= 64) ? 64 : 128;
  else if (N % 256 == 0) BN = 256;
  else if (N % 192 == 0 && !(K <= 256 && N % 128 == 0)) BN = 192;   // 16 epilogue warps want 4 | BN/32
  else if (N % 128 == 0) BN = 128;
  else if (N % 96 == 0) BN = 96;
  else if (N <= 64) BN = 64;
  else if (N <= 96) BN = 96;
  else BN = 128;
  if (rowsum_out != nullptr && BN == 256) BN = 128;   // the row-sum accumulators need 32 spare TMEM columns
  // Tile-width fallback: smaller tiles while the problem has fewer tiles than the chip has SMs.  Atomic outputs (the
  // wgrad GEMMs) run whole waves of the persistent grid, and 85 % of a wave is a wave: 144 tiles of 128 x 192 on 148 SMs
  // must not fall back to three times as many 64-wide ones (measured at 32256 tokens, N 1536, K 384: 93 us after the
  // fallback against 50 us for the best split without it).  VSN_WGRAD_WANT_PCT / VSN_WGRAD_MODEL=0: measurements only.
  static int want_pct = -1, wgrad_model = -1;
  if (want_pct < 0) { const char* e = getenv("VSN_WGRAD_WANT_PCT"); want_pct = e ? atoi(e) : 85; }
  if (wgrad_model < 0) { const char* e = getenv("VSN_WGRAD_MODEL"); wgrad_model = (e != nullptr && e[0] == '0') ? 0 : 1; }
  const int want_tiles = out_kind == 2 ? (vsn_num_sms() * want_pct) / 100 : vsn_num_sms();
  auto narrowed = [&](int bn, int sk) {
    while (bn > 64 && ceil_div(M, BM) * ceil_div(N, bn) * sk < want_tiles) {
      const int next = bn == 256 ? 128 : bn == 192 ? (b_mn ? 64 : 96) : bn == 128 ? 64 : bn == 96 ? (N % 64 == 0 || N <= 64 ? 64 : 96) : 64;
      if (next == bn) break;
      if (N % next != 0 && next < N) break;   // keep tiles exact when they were exact
      bn = next;
    }
    return bn;
  };
  if (split_k == 0 && out_kind == 2) {
    // Automatic split of the reduction for atomic outputs (the wgrad GEMMs: few output tiles, a very long K).
    const int base = ceil_div(M, BM) * ceil_div(N, BN), nkb_all = ceil_div(K, BK), sms = vsn_num_sms();
    if (base <= 8 && nkb_all >= 512) {
      // A handful of output tiles over tens or hundreds of thousands of tokens streams its operands from HBM: what counts
      // is whole waves of the persistent grid.  Cost = rounds x (k-blocks per tile + the tile's epilogue, worth
      // about eight k-blocks of fp32 reductions).  (Measured: 0.092 -> 0.069 ms for 435456 tokens, N 384, K 96.)
      long long best_cost = -1;
      int best = 1;
      for (int s = 1; s <= nkb_all && base * s <= 4 * sms; ++s) {
        const int kb = ceil_div(nkb_all, s);
        const long long cost = static_cast<long long>(ceil_div(base * ceil_div(nkb_all, kb), sms)) * (kb + 8);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
      }
      split_k = best;
    } else if (wgrad_model) {
      // Many output tiles (late stages, operands resident in L2): the kernel is bound by the operand bytes its CTAs pull
      // per k-block, (128 + BN) x 64 x 2, and runs in rounds of one tile per SM.  Cost of a split = rounds x (k-blocks
      // per tile x (128 + BN) + the tile's epilogue, ~6 x BN in the same unit), BN being the width the fallback above
      // leaves for that split.  The model reproduces the measured sweep (scripts/exp_wgrad_split.py: 1.6-1.9 ns per unit
      // over splits 4..48 at 32256 tokens) and replaces "two 128 x 128 tiles per SM", which left a 30 %-filled second
      // round at N 1536 / 1152, K 384 (67.6 -> 50.5 us, 55.4 -> 41.2 us at the best fixed split).
      long long best_cost = -1;
      int best = 1;
      const int tm = ceil_div(M, BM);
      for (int s = 1; s <= nkb_all && s <= 64; ++s) {
        const int kb = ceil_div(nkb_all, s), se = ceil_div(nkb_all, kb);
        if (se != s) continue;                           // the same effective split as a smaller s
        const int bn = narrowed(BN, se);
        const long long tiles = static_cast<long long>(tm) * ceil_div(N, bn) * se;
        const long long cost = ceil_div_ll(tiles, sms) * (static_cast<long long>(kb) * (128 + bn) + 6LL * bn);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
      }
      split_k = best;
    } else {
      // (the rule the model replaced: about two 128 x 128 tiles' worth of work per SM)
      const int t128 = ceil_div(M, 128) * ceil_div(N, N <= 64 ? 64 : 128);
      int s = (2 * sms) / (t128 > 0 ? t128 : 1);
      split_k = s < 1 ? 1 : (s > nkb_all ? nkb_all : s);
    }
  }
  BN = narrowed(BN, split_k < 1 ? 1 : split_k);
  // CTA pairs (cta_group::2, 256 x BN tiles): for store epilogues with at least two m-tiles; an MN-major B operand is
  // loaded in 64-column boxes, so its half tile must be a whole number of them.
  bool pair;
  {
    static int mode = -1;   // VSN_GEMM_PAIR=0|1 forces CTA pairs off / on wherever they are possible (measurements)
    if (mode < 0) { const char* e = getenv("VSN_GEMM_PAIR"); mode = e ? atoi(e) : 2; }
    const bool ew16 = out_kind != 2 && forced_ew() != 8;     // the pair kernels are built with 16 epilogue warps
    const bool possible = ew16 && rowsum_out == nullptr && ceil_div(M, BM) >= 2 && (!b_mn || BN % 128 == 0);
    // (measured: pays when the problem is not streaming one operand from HBM, i.e. both N and K reasonably large)
    pair = possible && (mode == 1 || (mode == 2 && N >= 192 && K >= 192));
  }
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap_2d(&tmA, A, K, M, lda, BK, BM);
  else rc = make_tmap_2d(&tmA, A, M, K, lda, 64, BK);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_2d(&tmB, B, K, N, ldb, BK, pair ? BN / 2 : BN);
  else rc = make_tmap_2d(&tmB, B, N, K, ldb, 64, BK);
  if (rc) return rc;

  const int nkb = ceil_div(K, BK);
  int splits = split_k < 1 ? 1 : split_k;
  if (splits > nkb) splits = nkb;
  int kbps = ceil_div(nkb, splits);
  splits = ceil_div(nkb, kbps);

  GemmArgs a;
  a.M = M; a.N = N; a.K = K; a.a_mn = a_mn; a.b_mn = b_mn; a.kb_per_split = kbps;
  a.out = out; a.ldo = ldo; a.out_kind = out_kind; a.bias = bias; a.act = act;
  a.aux = reinterpret_cast<bf16*>(aux); a.ldaux = ldaux; a.resid = resid; a.ldr = ldr;
  a.row_scale = row_scale; a.rows_per_group = rows_per_group > 0 ? rows_per_group : 1; a.alpha = alpha;
  {
    const long long osz = out_kind == 0 ? 2 : 4;
    bool wide = (reinterpret_cast<uintptr_t>(out) % 32 == 0) && ((ldo * osz) % 32 == 0);
    if (aux != nullptr) wide = wide && (reinterpret_cast<uintptr_t>(aux) % 32 == 0) && ((ldaux * 2) % 32 == 0);
    if (resid != nullptr) wide = wide && (reinterpret_cast<uintptr_t>(resid) % 32 == 0) && ((ldr * 4) % 32 == 0);
    a.wide = wide ? 1 : 0;
  }
  a.rowsum = rowsum_out;
  {
    static int gm = -1;
    if (gm < 0) { const char* e = getenv("VSN_GELU_FP32"); gm = e ? atoi(e) & 15 : 1; }   // default: fp32 forward, packed backward
    a.gelu_fp32 = gm;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int ew = out_kind != 2 ? 16 : 8;   // measured: 16 epilogue warps never lose on the store/activation epilogues
  if (forced_ew() == 8 || forced_ew() == 16) ew = forced_ew();
  if (ew == 16) {
    switch (BN) {
      case 64: return launch<64, 16>(tmA, tmB, a, splits, pair, s);
      case 96: return launch<96, 16>(tmA, tmB, a, splits, pair, s);
      case 192: return launch<192, 16>(tmA, tmB, a, splits, pair, s);
      case 256: return launch<256, 16>(tmA, tmB, a, splits, pair, s);
      default: return launch<128, 16>(tmA, tmB, a, splits, pair, s);
    }
  }
  switch (BN) {
    case 64: return launch<64, 8>(tmA, tmB, a, splits, pair, s);
    case 96: return launch<96, 8>(tmA, tmB, a, splits, pair, s);
    case 192: return launch<192, 8>(tmA, tmB, a, splits, pair, s);
    case 256: return launch<256, 8>(tmA, tmB, a, splits, pair, s);
    default: return launch<128, 8>(tmA, tmB, a, splits, pair, s);
  }
}
