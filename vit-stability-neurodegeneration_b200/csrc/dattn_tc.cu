// ViT-3D global multi-head attention core on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), forward and backward,
// for head_dim 64 and any sequence length (the shipped ViT-S/B/L/H all use dim_head 64; vit-3c has 811 tokens):
//     out = softmax((q k^T) * scale) v          models/vit_3d.py:131-141 between `to_qkv` and `to_out`
// qkv is [S*N, 3C] bf16 (q | k | v column blocks, head-major inside: 'b n (h d)'), out [S*N, C] bf16; sequence s
// owns rows s*N .. s*N+N-1.  Nothing of size N x N ever reaches HBM.
//
// Forward: CTA = (128-query tile, head, sequence); 256 TMEM columns and 82 KB of smem, so TWO CTAs share an SM and
//   one's softmax runs under the other's MMAs.  Per 128-key tile: S = Q K^T (4 SS MMAs, N = 128) -> 8 softmax warps,
//   thread = (query row, half of the keys): exp2 against the thread's OWN running reference, P (bf16) written back
//   over its own logit columns -> O_half += P_half V_half (TS MMAs, N = 64).  The two key halves of every tile keep
//   separate accumulators and references, so the two threads of a row never exchange anything until the final
//   combine; a reference is only raised (and its accumulator rescaled) when the tile maximum exceeds it by 2^8
//   (P <= 256 stays exact enough in bf16 and fp32), which after the first tile is rare.
// Backward: CTA = (128-key tile, head, sequence), TMEM lanes are keys, loop over the 128-query tiles:
//   S^T = K Q^T, dP^T = V dO^T (8 SS MMAs) -> P^T = exp2(S^T c - lse), dS^T = P^T (dP^T - delta) in place (bf16 pairs)
//   -> dV += P^T dO, dK += dS^T Q (TS MMAs from TMEM), dQ_tile = dS K through an MN-major smem copy of dS^T; dQ is
//   summed over the key-tile CTAs with fp32 red.add into a scratch buffer and cast to bf16 by a small kernel.
#include <cuda.h>
#include <math.h>
#include "tc.cuh"
#include "dattn_tc.cuh"

int vsn_make_tmap_2d_bf16(CUtensorMap* map, const void* base, long long dim0, long long dim1, long long ld, int box0,
                          int box1);

namespace {

constexpr int TQ = 128;                 // tile rows (queries in forward, keys in backward)
constexpr int HD = 64;
constexpr int TILE = TQ * HD * 2;       // 16 KB: [128 rows][64 bf16], 128-byte swizzle
constexpr float LOG2E = 1.4426950408889634f;
constexpr float NEG_INIT = -1.0e30f;

__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// =========================================== forward ==============================================
constexpr int FWD_THREADS = 10 * 32;    // 8 softmax warps + TMA warp + MMA warp
struct FwdSmem {
  static constexpr int Q = 0;
  static constexpr int KV = TILE;                 // 2 stages x (K | V)
  static constexpr int STATS = KV + 4 * TILE;     // float2 [128 rows][2 halves]
  static constexpr int BARS = STATS + 128 * 2 * 8;
  static constexpr int TOTAL = BARS + 16 * 8;
};

__global__ void __launch_bounds__(FWD_THREADS, 2)
dattn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const DenseAttnArgs p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + FwdSmem::Q;
  uint8_t* sKV = smem + FwdSmem::KV;
  float2* stats = reinterpret_cast<float2*>(smem + FwdSmem::STATS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BARS);
  uint64_t* q_full = bars;            // TMA -> MMA
  uint64_t* kv_full = bars + 1;       // [2] TMA -> MMA
  uint64_t* kv_empty = bars + 3;      // [2] MMA commit -> TMA
  uint64_t* s_full = bars + 5;        // MMA commit -> softmax
  uint64_t* p_ready = bars + 6;       // softmax -> MMA
  uint64_t* o_full = bars + 7;        // MMA commit (P V of the tile done) -> softmax (rescale / epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, s = blockIdx.z;
  const int nkt = (p.N + TQ - 1) / TQ;
  const long long row0 = static_cast<long long>(s) * p.N;

  if (threadIdx.x == 0) {
    if ((tc::smem_u32(smem) & 1023u) != 0) {
      printf("vsn_b200: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&kv_full[i], 1); tc::mbar_init(&kv_empty[i], 1); }
    tc::mbar_init(s_full, 1);
    tc::mbar_init(p_ready, 256);
    tc::mbar_init(o_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, 256);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t COL_S = 0, COL_O = 128;     // O of key half h at COL_O + 64 h

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tc::prefetch_tmap(&tm_qkv);
      pdl_wait();
      tc::mbar_arrive_expect_tx(q_full, TILE);
      tc::tma_load_2d(sQ, &tm_qkv, q_full, head * HD, static_cast<int>(row0 + qt * TQ));
      for (int j = 0; j < nkt; ++j) {
        const int st = j & 1;
        tc::mbar_wait_relaxed(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&kv_full[st], 2 * TILE);
        tc::tma_load_2d(sKV + st * 2 * TILE, &tm_qkv, &kv_full[st], p.C + head * HD, static_cast<int>(row0 + j * TQ));
        tc::tma_load_2d(sKV + st * 2 * TILE + TILE, &tm_qkv, &kv_full[st], 2 * p.C + head * HD,
                        static_cast<int>(row0 + j * TQ));
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (converged warp)
    const uint32_t idesc_s = tc::make_idesc_bf16(128, 128, 0, 0);      // S = Q K^T: both K-major
    const uint32_t idesc_o = tc::make_idesc_bf16(128, HD, 0, 1);       // O += P V: A in TMEM, V MN-major
    const uint64_t dq = tc::make_smem_desc_sw128(tc::smem_u32(sQ), 16, 1024);
    tc::mbar_wait(q_full, 0);
    for (int j = 0; j < nkt; ++j) {
      const int st = j & 1;
      tc::mbar_wait(&kv_full[st], (j >> 1) & 1);
      tc::fence_after_sync();
      const uint64_t dk = tc::make_smem_desc_sw128(tc::smem_u32(sKV + st * 2 * TILE), 16, 1024);
      const uint64_t dv = tc::make_smem_desc_sw128(tc::smem_u32(sKV + st * 2 * TILE + TILE), 8192, 1024);
      // (the S columns still hold P of the previous tile: its P V MMAs were issued before and the pipe is in order)
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc::mma_bf16_ss(tmem_base + COL_S, tc::desc_advance(dq, k * 32), tc::desc_advance(dk, k * 32), idesc_s, k);
        tc::mma_commit(s_full);
      }
      __syncwarp();
      tc::mbar_wait(p_ready, j & 1);
      tc::fence_after_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::mma_bf16_ts(tmem_base + COL_O + h * 64, tmem_base + COL_S + h * 64 + k * 8,
                            tc::desc_advance(dv, (h * 4 + k) * 2048), idesc_o, (j | k) ? 1u : 0u);
        tc::mma_commit(&kv_empty[st]);
        tc::mma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax (thread = query row, key half)
    const int q4 = warp & 3, h = warp >> 2;
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + COL_S + h * 64;
    const uint32_t o_addr = tmem_base + lane_addr + COL_O + h * 64;
    const float cscale = p.scale * LOG2E;
    float m_run = NEG_INIT, l_run = 0.f;
    for (int j = 0; j < nkt; ++j) {
      tc::mbar_wait(s_full, j & 1);
      tc::fence_after_sync();
      uint32_t x[64];
      tc::tmem_ld_32x32b_x32(s_addr, x);
      tc::tmem_ld_32x32b_x32(s_addr + 32, x + 32);
      tc::tmem_ld_wait();
      const int key0 = j * TQ + h * 64;
      float mx = -INFINITY;
      if (key0 + 64 <= p.N) {
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          const float v = __uint_as_float(x[k]) * cscale;
          x[k] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          const float v = key0 + k < p.N ? __uint_as_float(x[k]) * cscale : -INFINITY;
          x[k] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      }
      // Raise the reference only when the tile exceeds it by more than 2^8; the TMEM accesses are warp-collective, so
      // the decision is taken per warp and every lane then moves to its own exact maximum.
      if (__any_sync(0xffffffffu, mx > m_run + 8.0f)) {
        const float m_new = fmaxf(m_run, mx);
        if (j > 0) {
          const float f = tc::ex2_approx(m_run - m_new);
          tc::mbar_wait(o_full, (j - 1) & 1);           // the accumulator is complete up to the previous tile
          tc::fence_after_sync();
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {          // 8 columns at a time: the 64 logits stay in registers meanwhile
            uint32_t o[8];
            tc::tmem_ld_32x32b_x8(o_addr + c * 8, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * f);
            tc::tmem_st_32x32b_x8(o_addr + c * 8, o);
          }
          l_run *= f;
        }
        m_run = m_new;
      }
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float p0 = tc::ex2_approx(__uint_as_float(x[2 * k]) - m_run);
        const float p1 = tc::ex2_approx(__uint_as_float(x[2 * k + 1]) - m_run);
        l0 += p0;
        l1 += p1;
        x[k] = pack_bf16(p0, p1);
      }
      l_run += l0 + l1;
      tc::tmem_st_32x32b_x32(s_addr, x);                 // P (64 bf16) over the first 32 of the thread's own columns
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc::mbar_arrive(p_ready);
    }
    // ---- epilogue: combine the two key halves of the row; this thread writes output columns [32 h, 32 h + 32)
    stats[r * 2 + h] = make_float2(m_run, l_run);
    tc::named_bar_sync(1, 256);
    const float2 sa = stats[r * 2], sb = stats[r * 2 + 1];
    const float m = fmaxf(sa.x, sb.x);
    const float wa = tc::ex2_approx(sa.x - m), wb = tc::ex2_approx(sb.x - m);
    const float l = wa * sa.y + wb * sb.y;
    const float inv = 1.f / l;
    tc::mbar_wait(o_full, (nkt - 1) & 1);
    tc::fence_after_sync();
    uint32_t oa[32], ob[32];
    tc::tmem_ld_32x32b_x32(tmem_base + lane_addr + COL_O + h * 32, oa);
    tc::tmem_ld_32x32b_x32(tmem_base + lane_addr + COL_O + 64 + h * 32, ob);
    tc::tmem_ld_wait();
    const int i = qt * TQ + r;
    if (i < p.N) {
      bf16* dst = p.out + (row0 + i) * p.C + head * HD + h * 32;
      const float fa = wa * inv, fb = wb * inv;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = c * 16 + 2 * k;
          w[k] = pack_bf16(fmaf(fa, __uint_as_float(oa[e]), fb * __uint_as_float(ob[e])),
                           fmaf(fa, __uint_as_float(oa[e + 1]), fb * __uint_as_float(ob[e + 1])));
        }
        st_global_v8(dst + c * 16, w);
      }
      if (h == 0 && p.lse != nullptr)
        p.lse[(static_cast<long long>(s) * p.heads + head) * p.Npad + i] = m + log2f(l);     // log2 domain
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 256);
  }
}

// =========================================== backward =============================================
constexpr int BWD_THREADS = 10 * 32;
struct BwdSmem {
  static constexpr int K = 0;                       // K tile, V tile (loaded once)
  static constexpr int V = TILE;
  static constexpr int QDO = 2 * TILE;              // 2 stages x (Q | dO)
  static constexpr int DS = QDO + 4 * TILE;         // dS^T as MN-major A operand: 2 query blocks x [128 keys][64 queries]
  static constexpr int LSE = DS + 2 * TILE;         // 2 stages x (lse2[128] | delta[128]) fp32
  static constexpr int BARS = LSE + 2 * 1024;
  static constexpr int TOTAL = BARS + 16 * 8;
};

__global__ void __launch_bounds__(BWD_THREADS, 1)
dattn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                 const DenseAttnArgs p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + BwdSmem::K;
  uint8_t* sV = smem + BwdSmem::V;
  uint8_t* sQDO = smem + BwdSmem::QDO;
  uint8_t* sDS = smem + BwdSmem::DS;
  float* sLSE = reinterpret_cast<float*>(smem + BwdSmem::LSE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::BARS);
  uint64_t* kv_full = bars;           // TMA -> MMA (K, V tiles)
  uint64_t* qdo_full = bars + 1;      // [2] TMA bytes + 32 loader lanes (lse / delta) -> MMA, compute
  uint64_t* qdo_empty = bars + 3;     // [2] MMA commit -> TMA warp
  uint64_t* s_full = bars + 5;        // MMA commit (S^T, dP^T) -> compute
  uint64_t* p_ready = bars + 6;       // compute -> MMA
  uint64_t* dq_full = bars + 7;       // MMA commit (dV, dK, dQ of the query tile) -> compute (dQ read-out)
  uint64_t* dq_read = bars + 8;       // compute -> MMA (dQ accumulator may be overwritten)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, head = blockIdx.y, s = blockIdx.z;
  const int nqt = (p.N + TQ - 1) / TQ;
  const long long row0 = static_cast<long long>(s) * p.N;

  if (threadIdx.x == 0) {
    if ((tc::smem_u32(smem) & 1023u) != 0) {
      printf("vsn_b200: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tc::mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&qdo_full[i], 33); tc::mbar_init(&qdo_empty[i], 1); }
    tc::mbar_init(s_full, 1);
    tc::mbar_init(p_ready, 256);
    tc::mbar_init(dq_full, 1);
    tc::mbar_init(dq_read, 256);
    tc::fence_barrier_init();
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer + lse / delta loader
    pdl_wait();
    if (lane == 0) {
      tc::prefetch_tmap(&tm_qkv);
      tc::prefetch_tmap(&tm_do);
      tc::mbar_arrive_expect_tx(kv_full, 2 * TILE);
      tc::tma_load_2d(sK, &tm_qkv, kv_full, p.C + head * HD, static_cast<int>(row0 + kt * TQ));
      tc::tma_load_2d(sV, &tm_qkv, kv_full, 2 * p.C + head * HD, static_cast<int>(row0 + kt * TQ));
    }
    const float* lse_g = p.lse + (static_cast<long long>(s) * p.heads + head) * p.Npad;
    const float* dl_g = p.delta + (static_cast<long long>(s) * p.heads + head) * p.Npad;
    for (int i = 0; i < nqt; ++i) {
      const int st = i & 1;
      tc::mbar_wait_relaxed(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
      if (lane == 0) {
        tc::mbar_arrive_expect_tx(&qdo_full[st], 2 * TILE);
        tc::tma_load_2d(sQDO + st * 2 * TILE, &tm_qkv, &qdo_full[st], head * HD, static_cast<int>(row0 + i * TQ));
        tc::tma_load_2d(sQDO + st * 2 * TILE + TILE, &tm_do, &qdo_full[st], head * HD, static_cast<int>(row0 + i * TQ));
      }
      // queries past the end of the sequence get lse = +big: P = exp2(x - big) = 0, so they contribute nothing
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int q = i * TQ + c * 32 + lane;
        sLSE[st * 256 + c * 32 + lane] = q < p.N ? lse_g[q] : 1.0e30f;
        sLSE[st * 256 + 128 + c * 32 + lane] = q < p.N ? dl_g[q] : 0.f;
      }
      tc::mbar_arrive(&qdo_full[st]);
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (converged warp)
    const uint32_t idesc_s = tc::make_idesc_bf16(128, 128, 0, 0);      // S^T = K Q^T, dP^T = V dO^T: K-major both
    const uint32_t idesc_kv = tc::make_idesc_bf16(128, HD, 0, 1);      // dV, dK: A in TMEM, B MN-major
    const uint32_t idesc_q = tc::make_idesc_bf16(128, HD, 1, 1);       // dQ = dS K: A MN-major smem, B MN-major
    const uint64_t dk = tc::make_smem_desc_sw128(tc::smem_u32(sK), 16, 1024);
    const uint64_t dv = tc::make_smem_desc_sw128(tc::smem_u32(sV), 16, 1024);
    const uint64_t dk_mn = tc::make_smem_desc_sw128(tc::smem_u32(sK), 8192, 1024);
    const uint64_t dds = tc::make_smem_desc_sw128(tc::smem_u32(sDS), TILE, 1024);
    tc::mbar_wait(kv_full, 0);
    for (int i = 0; i < nqt; ++i) {
      const int st = i & 1;
      tc::mbar_wait(&qdo_full[st], (i >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t aq = tc::smem_u32(sQDO + st * 2 * TILE), ado = aq + TILE;
      const uint64_t dq = tc::make_smem_desc_sw128(aq, 16, 1024), ddo = tc::make_smem_desc_sw128(ado, 16, 1024);
      const uint64_t dq_mn = tc::make_smem_desc_sw128(aq, 8192, 1024), ddo_mn = tc::make_smem_desc_sw128(ado, 8192, 1024);
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc::mma_bf16_ss(tmem_base + COL_S, tc::desc_advance(dk, k * 32), tc::desc_advance(dq, k * 32), idesc_s, k);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc::mma_bf16_ss(tmem_base + COL_DP, tc::desc_advance(dv, k * 32), tc::desc_advance(ddo, k * 32), idesc_s, k);
        tc::mma_commit(s_full);
      }
      __syncwarp();
      tc::mbar_wait(p_ready, i & 1);
      if (i > 0) tc::mbar_wait(dq_read, (i - 1) & 1);       // the previous query tile's dQ has been read out
      tc::fence_after_sync();
      if (tc::elect_one()) {
        // P^T / dS^T pairs of the 16 queries kk: columns 64 (kk / 4) + 8 (kk % 4) of their region
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t a = (kk >> 2) * 64 + (kk & 3) * 8;
          tc::mma_bf16_ts(tmem_base + COL_DV, tmem_base + COL_S + a, tc::desc_advance(ddo_mn, kk * 2048), idesc_kv,
                          (i | kk) ? 1u : 0u);
          tc::mma_bf16_ts(tmem_base + COL_DK, tmem_base + COL_DP + a, tc::desc_advance(dq_mn, kk * 2048), idesc_kv,
                          (i | kk) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)       // dQ[128 queries, 64] = sum over 16-key steps of dS[:, keys] K[keys, :]
          tc::mma_bf16_ss(tmem_base + COL_DQ, tc::desc_advance(dds, kk * 2048), tc::desc_advance(dk_mn, kk * 2048),
                          idesc_q, kk);
        tc::mma_commit(dq_full);
        tc::mma_commit(&qdo_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ compute (thread = key row, 64-query half)
    const int q4 = warp & 3, h = warp >> 2;
    const int r = q4 * 32 + lane;
    const int j = kt * TQ + r;                       // key token of the sequence
    const bool key_ok = j < p.N;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const float cscale = p.scale * LOG2E;
    // smem dS tile: element (query q, key r) of block q / 64 at (r / 8) * 1024 + (r % 8) * 128 + (((q % 64) / 8) ^ (r % 8)) * 16
    uint8_t* ds_row = sDS + h * TILE + (r >> 3) * 1024 + (r & 7) * 128;
    for (int i = 0; i < nqt; ++i) {
      const int st = i & 1;
      tc::mbar_wait(s_full, i & 1);
      tc::fence_after_sync();
      const float* lse_s = sLSE + st * 256 + h * 64;
      const float* dl_s = lse_s + 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t xs[32], xd[32];
        tc::tmem_ld_32x32b_x32(tmem_base + lane_addr + COL_S + h * 64 + c * 32, xs);
        tc::tmem_ld_32x32b_x32(tmem_base + lane_addr + COL_DP + h * 64 + c * 32, xd);
        tc::tmem_ld_wait();
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 l = *reinterpret_cast<const float4*>(lse_s + c * 32 + q * 4);
          const float4 d = *reinterpret_cast<const float4*>(dl_s + c * 32 + q * 4);
          float p0 = tc::ex2_approx(fmaf(__uint_as_float(xs[4 * q]), cscale, -l.x));
          float p1 = tc::ex2_approx(fmaf(__uint_as_float(xs[4 * q + 1]), cscale, -l.y));
          float p2 = tc::ex2_approx(fmaf(__uint_as_float(xs[4 * q + 2]), cscale, -l.z));
          float p3 = tc::ex2_approx(fmaf(__uint_as_float(xs[4 * q + 3]), cscale, -l.w));
          if (!key_ok) { p0 = 0.f; p1 = 0.f; p2 = 0.f; p3 = 0.f; }
          pk[2 * q] = pack_bf16(p0, p1);
          pk[2 * q + 1] = pack_bf16(p2, p3);
          dk[2 * q] = pack_bf16(p0 * (__uint_as_float(xd[4 * q]) - d.x), p1 * (__uint_as_float(xd[4 * q + 1]) - d.y));
          dk[2 * q + 1] = pack_bf16(p2 * (__uint_as_float(xd[4 * q + 2]) - d.z), p3 * (__uint_as_float(xd[4 * q + 3]) - d.w));
        }
        // bf16 pairs of queries 64 h + 32 c .. + 31 over columns 64 h + 16 c .. + 15 of their own regions
        tc::tmem_st_32x32b_x8(tmem_base + lane_addr + COL_S + h * 64 + c * 16, pk);
        tc::tmem_st_32x32b_x8(tmem_base + lane_addr + COL_S + h * 64 + c * 16 + 8, pk + 8);
        tc::tmem_st_32x32b_x8(tmem_base + lane_addr + COL_DP + h * 64 + c * 16, dk);
        tc::tmem_st_32x32b_x8(tmem_base + lane_addr + COL_DP + h * 64 + c * 16 + 8, dk + 8);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(ds_row + (((4 * c + u) ^ (r & 7)) << 4)) =
              make_uint4(dk[4 * u], dk[4 * u + 1], dk[4 * u + 2], dk[4 * u + 3]);
      }
      tc::fence_proxy_async();
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc::mbar_arrive(p_ready);

      // ---- dQ of this query tile: lanes are QUERIES now; this thread adds columns [32 h, 32 h + 32) of row r
      tc::mbar_wait(dq_full, i & 1);
      tc::fence_after_sync();
      uint32_t acc[32];
      tc::tmem_ld_32x32b_x32(tmem_base + lane_addr + COL_DQ + h * 32, acc);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      tc::mbar_arrive(dq_read);
      const int qi = i * TQ + r;
      if (qi < p.N) {
        float* dst = p.dq_acc + (row0 + qi) * p.C + head * HD + h * 32;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          red_add_f32x4(dst + 4 * c, __uint_as_float(acc[4 * c]), __uint_as_float(acc[4 * c + 1]),
                        __uint_as_float(acc[4 * c + 2]), __uint_as_float(acc[4 * c + 3]));
      }
    }
    // ---- dV / dK rows of this key tile (the last dq_full commit covers their MMAs too)
    uint32_t av[32], ak[32];
    tc::tmem_ld_32x32b_x32(tmem_base + lane_addr + COL_DV + h * 32, av);
    tc::tmem_ld_32x32b_x32(tmem_base + lane_addr + COL_DK + h * 32, ak);
    tc::tmem_ld_wait();
    if (key_ok) {
      bf16* dkp = p.dqkv + (row0 + j) * (3LL * p.C) + p.C + head * HD + h * 32;
      bf16* dvp = dkp + p.C;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t wk[8], wv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = c * 16 + 2 * k;
          wk[k] = pack_bf16(__uint_as_float(ak[e]) * p.scale, __uint_as_float(ak[e + 1]) * p.scale);
          wv[k] = pack_bf16(__uint_as_float(av[e]), __uint_as_float(av[e + 1]));
        }
        st_global_v8(dkp + c * 16, wk);
        st_global_v8(dvp + c * 16, wv);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// delta[s, head, i] = rowsum(dO * O) (softmax backward) and dq_acc = 0; one warp per token row, a lane pair per head.
__global__ void __launch_bounds__(256) dattn_delta_kernel(const DenseAttnArgs p) {
  pdl_trigger();
  const long long T = static_cast<long long>(p.S) * p.N;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < T; row += warps) {
    const int s = static_cast<int>(row / p.N), i = static_cast<int>(row - static_cast<long long>(s) * p.N);
    const bf16* o = p.out + row * p.C;
    const bf16* d = p.dout + row * p.C;
    float* z = p.dq_acc + row * p.C;
    for (int c0 = 0; c0 < p.C; c0 += 256) {          // 8 columns per lane and pass; a head = 8 lanes
      const int c = c0 + lane * 8;
      float acc = 0.f;
      if (c < p.C) {
        const uint4 a = *reinterpret_cast<const uint4*>(o + c), b = *reinterpret_cast<const uint4*>(d + c);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 x = unpack_bf16(aw[t]), y = unpack_bf16(bw[t]);
          acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
        }
        *reinterpret_cast<float4*>(z + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(z + c + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if ((lane & 7) == 0 && c < p.C) p.delta[(static_cast<long long>(s) * p.heads + c / HD) * p.Npad + i] = acc;
    }
  }
}

// dqkv[:, 0:C] = bf16(scale * dq_acc)
__global__ void __launch_bounds__(256) dattn_dq_cast_kernel(const DenseAttnArgs p) {
  pdl_trigger();
  const long long T = static_cast<long long>(p.S) * p.N;
  const int c8 = p.C / 8;
  const long long total = T * c8;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = idx / c8;
    const int c = static_cast<int>(idx - row * c8) * 8;
    const float4 a = *reinterpret_cast<const float4*>(p.dq_acc + row * p.C + c);
    const float4 b = *reinterpret_cast<const float4*>(p.dq_acc + row * p.C + c + 4);
    uint4 w;
    w.x = pack_bf16(a.x * p.scale, a.y * p.scale);
    w.y = pack_bf16(a.z * p.scale, a.w * p.scale);
    w.z = pack_bf16(b.x * p.scale, b.y * p.scale);
    w.w = pack_bf16(b.z * p.scale, b.w * p.scale);
    *reinterpret_cast<uint4*>(p.dqkv + row * (3LL * p.C) + c) = w;
  }
}

int check_args(const DenseAttnArgs& a) {
  VSN_CHECK(a.C == a.heads * HD, "dense tcgen05 attention: head_dim must be 64");
  VSN_CHECK(a.C % 8 == 0, "dense tcgen05 attention: C must be a multiple of 8");
  VSN_CHECK(static_cast<long long>(a.S) * a.N < (1LL << 31), "dense tcgen05 attention: too many rows for 32-bit TMA coordinates");
  return 0;
}

}  // namespace

bool dattn_tc_supported(int hd) { return hd == HD; }

int dattn_tc_fwd(const DenseAttnArgs& a, cudaStream_t stream) {
  if (int rc = check_args(a)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    VSN_CUDA(cudaFuncSetAttribute(dattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
    attr_set = true;
  }
  const long long T = static_cast<long long>(a.S) * a.N;
  CUtensorMap tm;
  if (int rc = vsn_make_tmap_2d_bf16(&tm, a.qkv, 3LL * a.C, T, 3LL * a.C, HD, TQ)) return rc;
  dim3 grid((a.N + TQ - 1) / TQ, a.heads, a.S);
  dattn_fwd_kernel<<<grid, FWD_THREADS, FwdSmem::TOTAL, stream>>>(tm, a);
  VSN_LAUNCH_CHECK();
  return 0;
}

int dattn_tc_bwd(const DenseAttnArgs& a, cudaStream_t stream) {
  if (int rc = check_args(a)) return rc;
  VSN_CHECK(a.dq_acc != nullptr, "dense tcgen05 attention backward needs the fp32 dQ scratch [S*N, C]");
  static bool attr_set = false;
  if (!attr_set) {
    VSN_CUDA(cudaFuncSetAttribute(dattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL));
    attr_set = true;
  }
  const long long T = static_cast<long long>(a.S) * a.N;
  CUtensorMap tq, td;
  if (int rc = vsn_make_tmap_2d_bf16(&tq, a.qkv, 3LL * a.C, T, 3LL * a.C, HD, TQ)) return rc;
  if (int rc = vsn_make_tmap_2d_bf16(&td, a.dout, a.C, T, a.C, HD, TQ)) return rc;
  {
    long long blocks = (T * 32 + 255) / 256;
    const long long cap = static_cast<long long>(vsn_num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    dattn_delta_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a);
    VSN_LAUNCH_CHECK();
  }
  dim3 grid((a.N + TQ - 1) / TQ, a.heads, a.S);
  dattn_bwd_kernel<<<grid, BWD_THREADS, BwdSmem::TOTAL, stream>>>(tq, td, a);
  VSN_LAUNCH_CHECK();
  {
    long long blocks = (T * (a.C / 8) + 255) / 256;
    const long long cap = static_cast<long long>(vsn_num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    dattn_dq_cast_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a);
    VSN_LAUNCH_CHECK();
  }
  return 0;
}
