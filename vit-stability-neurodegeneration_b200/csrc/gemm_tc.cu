// bf16 GEMM on the 5th-gen tensor cores: TMA -> 128B-swizzled smem -> tcgen05.mma (fp32 accumulators in
// TMEM) -> tcgen05.ld -> fused epilogue.  One 128 x BN output tile per CTA, warp-specialised:
//   warp 0   TMA producer (one elected lane)
//   warp 1   TMEM allocator + MMA issuer (one elected lane)
//   warps 2-5 epilogue (each owns the 32 TMEM lanes its warp id maps to)
//
// Operand majors cover all three GEMMs of a Linear layer without materialising transposes:
//   forward  Y[M,N]  = X[M,K]  * W[N,K]^T     A K-major,  B K-major
//   dgrad    dX[M,K] = dY[M,N] * W[N,K]       A K-major,  B MN-major (W read as stored)
//   wgrad    dW[N,K] = dY[M,N]^T * X[M,K]     A MN-major, B MN-major, split over tokens, fp32 red.add
// Replaces the cuBLAS calls behind nn.Linear / F.linear in the reference
// (models/swin_transformer_3d.py:52-69,154-156,550; models/vit_3d.py:59-75,102-105,372).
#include <cuda.h>
#include "tc.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int A_BYTES = BM * BK * 2;

struct GemmArgs {
  int M, N, K;          // output is M x N, reduction length K
  int a_mn, b_mn;       // operand majors (0 = K-major, 1 = MN-major)
  int kb_per_split;     // k-blocks handled by one blockIdx.z
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
};

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN <= 96) ? 4 : (BN <= 128 ? 3 : 2);
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* accum_bar = empty_bar + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int nkb_total = (p.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(nkb_total, kb0 + p.kb_per_split);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(accum_bar, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (nkb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int i = 0; i < nkb; ++i) {
          const int s = i % C::STAGES;
          const uint32_t ph = (i / C::STAGES) & 1;
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * C::STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          tc::mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
          const int k0 = (kb0 + i) * BK;
          if (!p.a_mn) {
            tc::tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);           // box {64 k, 128 rows}
          } else {
            tc::tma_load_2d(sa, &tmA, &full_bar[s], m0, k0);           // box {64 m, 64 k-rows}
            tc::tma_load_2d(sa + 8192, &tmA, &full_bar[s], m0 + 64, k0);
          }
          if (!p.b_mn) {
            tc::tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);           // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tc::tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + j * 64, k0);
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
        const uint32_t a_step = p.a_mn ? 2048u : 32u;   // bytes per K=16 step
        const uint32_t b_step = p.b_mn ? 2048u : 32u;
        const uint32_t a_lbo = p.a_mn ? 8192u : 16u;
        const uint32_t b_lbo = p.b_mn ? 8192u : 16u;
        for (int i = 0; i < nkb; ++i) {
          const int s = i % C::STAGES;
          const uint32_t ph = (i / C::STAGES) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + s * C::STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const int krem = p.K - (kb0 + i) * BK;
          const int ksteps = krem >= BK ? BK / 16 : (krem + 15) / 16;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t ad = tc::make_smem_desc_sw128(sa + k * a_step, a_lbo, 1024);
            const uint64_t bd = tc::make_smem_desc_sw128(sb + k * b_step, b_lbo, 1024);
            tc::mma_bf16_ss(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          tc::mma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs retire
        }
        tc::mma_commit(accum_bar);         // accumulator complete
      }
    } else {
      // ---- epilogue: thread <-> output row --------------------------------------------
      const int lg = warp & 3;                     // TMEM lane group this warp may touch
      const int row = m0 + lg * 32 + lane;
      tc::mbar_wait(accum_bar, 0);
      tc::fence_after_sync();
      const bool row_ok = row < p.M;
      float rs = p.alpha;
      if (p.row_scale != nullptr && row_ok) rs *= p.row_scale[row / p.rows_per_group];
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
        const int n = n0 + c0;
        if (!row_ok || n >= p.N) continue;
        const int nvalid = min(32, p.N - n);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr) {
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(p.bias + n + j);
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          } else {
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { f[j] += p.bias[n + j]; }
          }
        }
        if (p.act == 1) {
          bf16* ap = p.aux + static_cast<long long>(row) * p.ldaux + n;
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_bf16(f[j], f[j + 1]); u.y = pack_bf16(f[j + 2], f[j + 3]);
              u.z = pack_bf16(f[j + 4], f[j + 5]); u.w = pack_bf16(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(ap + j) = u;
            }
          } else {
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { ap[j] = __float2bfloat16(f[j]); }
          }
          // GELU is evaluated on the bf16-rounded pre-activation so backward (which only has the
          // saved bf16 value) differentiates exactly the function forward applied.
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_erf(__bfloat162float(__float2bfloat16(f[j])));
        } else if (p.act == 2) {
          const bf16* ap = p.aux + static_cast<long long>(row) * p.ldaux + n;
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 u = *reinterpret_cast<const uint4*>(ap + j);
              float2 t;
              t = unpack_bf16(u.x); f[j] *= gelu_erf_grad(t.x); f[j + 1] *= gelu_erf_grad(t.y);
              t = unpack_bf16(u.y); f[j + 2] *= gelu_erf_grad(t.x); f[j + 3] *= gelu_erf_grad(t.y);
              t = unpack_bf16(u.z); f[j + 4] *= gelu_erf_grad(t.x); f[j + 5] *= gelu_erf_grad(t.y);
              t = unpack_bf16(u.w); f[j + 6] *= gelu_erf_grad(t.x); f[j + 7] *= gelu_erf_grad(t.y);
            }
          } else {
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { f[j] *= gelu_erf_grad(__bfloat162float(ap[j])); }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] *= rs;
        if (p.resid != nullptr) {
          const float* rp = p.resid + static_cast<long long>(row) * p.ldr + n;
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
              f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
            }
          } else {
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { f[j] += rp[j]; }
          }
        }
        if (p.out_kind == 0) {
          bf16* op = reinterpret_cast<bf16*>(p.out) + static_cast<long long>(row) * p.ldo + n;
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_bf16(f[j], f[j + 1]); u.y = pack_bf16(f[j + 2], f[j + 3]);
              u.z = pack_bf16(f[j + 4], f[j + 5]); u.w = pack_bf16(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(op + j) = u;
            }
          } else {
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { op[j] = __float2bfloat16(f[j]); }
          }
        } else if (p.out_kind == 1) {
          float* op = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + n;
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { op[j] = f[j]; }
          }
        } else {
          float* op = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + n;
          _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) { atomicAdd(op + j, f[j]); }
        }
      }
      tc::fence_before_sync();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
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
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VSN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (dims %lld x %lld ld %lld box %d x %d)",
            static_cast<int>(r), dim0, dim1, ld, box0, box1);
  return 0;
}

template <int BN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int splits, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VSN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid(ceil_div(a.M, BM), ceil_div(a.N, BN), splits);
  gemm_tc_kernel<BN><<<grid, 192, Cfg<BN>::SMEM_BYTES, stream>>>(tmA, tmB, a);
  VSN_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int vsn_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, int M,
                             int N, int K, void* out, long long ldo, int out_kind, const float* bias, int act,
                             void* aux, long long ldaux, const float* resid, long long ldr, const float* row_scale,
                             int rows_per_group, float alpha, int split_k, void* stream) {
  VSN_CHECK(M > 0 && N > 0 && K > 0, "vsn_gemm_bf16: empty problem %d x %d x %d", M, N, K);
  VSN_CHECK(out_kind >= 0 && out_kind <= 2, "vsn_gemm_bf16: bad out_kind %d", out_kind);
  VSN_CHECK(act == 0 || aux != nullptr, "vsn_gemm_bf16: activation modes need the aux buffer");
  VSN_CHECK(split_k <= 1 || out_kind == 2, "vsn_gemm_bf16: split_k needs atomic fp32 output");
  VSN_CHECK(out_kind != 2 || (bias == nullptr && act == 0 && resid == nullptr),
            "vsn_gemm_bf16: atomic output takes no bias/activation/residual");
  int BN;
  if (b_mn) BN = (N <= 64) ? 64 : 128;
  else if (N % 128 == 0) BN = 128;
  else if (N % 96 == 0) BN = 96;
  else if (N <= 64) BN = 64;
  else if (N <= 96) BN = 96;
  else BN = 128;

  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap_2d(&tmA, A, K, M, lda, BK, BM);
  else rc = make_tmap_2d(&tmA, A, M, K, lda, 64, BK);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_2d(&tmB, B, K, N, ldb, BK, BN);
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
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (BN) {
    case 64: return launch<64>(tmA, tmB, a, splits, s);
    case 96: return launch<96>(tmA, tmB, a, splits, s);
    default: return launch<128>(tmA, tmB, a, splits, s);
  }
}
