// Micro-benchmark (measurement only, not part of the library): tensor-pipe time of the exact tcgen05.mma mixes the
// window-attention backward issues per sub-tile, with nothing else running on the SM -- the floor the kernel's
// MMA-issuer loop can reach -- and of each operand flavour on its own (K-major vs MN-major smem operands, A from
// TMEM, identity-matrix bias-gradient MMAs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vit-stability-neurodegeneration_b200/csrc \
//        scripts/ub/mma_seq.cu -o gpurun_out/mma_seq && gpurun_out/mma_seq
#include <cstdio>
#include <vector>
#include "tc.cuh"

void vsn_set_error(const char*, ...) {}
void vsn_count_launch() {}

constexpr int REPS = 64;

// MODE 0: current kernel, 32-query sub-tile: S^T, dP^T = 2+2 SS N=32 (K-major); dV, dK = 2+2 TS N=32 (B MN-major);
//         dBias = 2 TS N=16; every 4th sub-tile 8 SS N=32 with MN-major A (128B swizzle) and MN-major B (dQ).
// MODE 1: quarter design, 64-query sub-tile: 2+2 SS N=64; 4+4 TS N=32; 4 TS N=16; every 2nd sub-tile 8 dQ MMAs.
// MODE 2..6: 16 MMAs of one flavour: 2 = SS K-major N=32, 3 = TS B MN-major N=32, 4 = dQ flavour, 5 = TS N=16 ident,
//         6 = SS K-major N=64, 7 = SS K-major N=128, 8 = dQ flavour with N = 64 (two heads' worth of columns)
template <int MODE>
__global__ void __launch_bounds__(128, 1) bench(unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    const uint32_t id_s32 = tc::make_idesc_bf16(128, 32, 0, 0), id_s64 = tc::make_idesc_bf16(128, 64, 0, 0);
    const uint32_t id_s128 = tc::make_idesc_bf16(128, 128, 0, 0);
    const uint32_t id_kv = tc::make_idesc_bf16(128, 32, 0, 1), id_b = tc::make_idesc_bf16(128, 16, 0, 0);
    const uint32_t id_q = tc::make_idesc_bf16(128, 32, 1, 1), id_q64 = tc::make_idesc_bf16(128, 64, 1, 1);
    const uint64_t d_k = tc::make_smem_desc_sw64(tc::smem_u32(smem), 16, 512);              // [256][32] K-major tiles
    const uint64_t d_q = tc::make_smem_desc_sw64(tc::smem_u32(smem + 16384), 16, 512);
    const uint64_t d_id = tc::make_smem_desc_sw64(tc::smem_u32(smem + 32768), 16, 512);
    const uint64_t d_ds = tc::make_smem_desc_sw128(tc::smem_u32(smem + 65536), 16384, 1024);   // dS tile, MN-major A
    const uint64_t d_kmn = tc::make_smem_desc_sw128(tc::smem_u32(smem + 49152), 16384, 1024);  // [k][64] MN-major B (N = 64)
    unsigned long long best = ~0ull;
    for (int trial = 0; trial < 5; ++trial) {
      __syncwarp();
      const unsigned long long t0 = clock64();
      if (tc::elect_one()) {
#pragma unroll 1
        for (int r = 0; r < REPS; ++r) {
          const uint32_t buf = tm + 256 + (r & 1) * 128;
          if (MODE == 0) {
            const uint32_t b2 = tm + 256 + (r & 1) * 64;
            for (int k = 0; k < 2; ++k) tc::mma_bf16_ss(b2, tc::desc_advance(d_k, k * 32), tc::desc_advance(d_q, k * 32), id_s32, k);
            for (int k = 0; k < 2; ++k) tc::mma_bf16_ss(b2 + 32, tc::desc_advance(d_k, 8192 + k * 32), tc::desc_advance(d_q, k * 32), id_s32, k);
            const uint32_t a = tm + 256 + ((r + 1) & 1) * 64;
            for (int k = 0; k < 2; ++k) {
              tc::mma_bf16_ts(tm + 448, a + k * 16, tc::desc_advance(d_q, k * 1024), id_kv, 1);
              tc::mma_bf16_ts(tm + 480, a + 32 + k * 16, tc::desc_advance(d_k, k * 1024), id_kv, 1);
              tc::mma_bf16_ts(tm + (r & 7) * 32 + k * 16, a + 32 + k * 16, d_id, id_b, 1);
            }
            if ((r & 3) == 3)
              for (int k = 0; k < 8; ++k)
                tc::mma_bf16_ss(tm + 384, tc::desc_advance(d_ds, k * 2048), tc::desc_advance(d_k, k * 1024), id_q, k);
          } else if (MODE == 1) {
            for (int k = 0; k < 2; ++k) tc::mma_bf16_ss(buf, tc::desc_advance(d_k, k * 32), tc::desc_advance(d_q, k * 32), id_s64, k);
            for (int k = 0; k < 2; ++k) tc::mma_bf16_ss(buf + 64, tc::desc_advance(d_k, 8192 + k * 32), tc::desc_advance(d_q, k * 32), id_s64, k);
            const uint32_t a = tm + 256 + ((r + 1) & 1) * 128;
            for (int k = 0; k < 4; ++k) {
              tc::mma_bf16_ts(tm + 128, a + k * 8, tc::desc_advance(d_q, k * 1024), id_kv, 1);
              tc::mma_bf16_ts(tm + 160, a + 64 + k * 8, tc::desc_advance(d_k, k * 1024), id_kv, 1);
              tc::mma_bf16_ts(tm + (r & 1) * 64 + k * 16, a + 64 + k * 8, d_id, id_b, 1);
            }
            if (r & 1)
              for (int k = 0; k < 8; ++k)
                tc::mma_bf16_ss(tm + 192, tc::desc_advance(d_ds, k * 2048), tc::desc_advance(d_k, k * 1024), id_q, k);
          } else {
            for (int k = 0; k < 16; ++k) {
              if (MODE == 2) tc::mma_bf16_ss(buf, tc::desc_advance(d_k, (k & 1) * 32), tc::desc_advance(d_q, (k & 1) * 32), id_s32, 1);
              if (MODE == 3) tc::mma_bf16_ts(tm + 128, buf + (k & 3) * 8, tc::desc_advance(d_q, (k & 7) * 1024), id_kv, 1);
              if (MODE == 4) tc::mma_bf16_ss(tm + 192, tc::desc_advance(d_ds, (k & 7) * 2048), tc::desc_advance(d_k, (k & 7) * 1024), id_q, 1);
              if (MODE == 5) tc::mma_bf16_ts(tm + (k & 7) * 16, buf + (k & 3) * 8, d_id, id_b, 1);
              if (MODE == 6) tc::mma_bf16_ss(buf, tc::desc_advance(d_k, (k & 1) * 32), tc::desc_advance(d_q, (k & 1) * 32), id_s64, 1);
              if (MODE == 7) tc::mma_bf16_ss(buf, tc::desc_advance(d_k, (k & 1) * 32), tc::desc_advance(d_q, (k & 1) * 32), id_s128, 1);
              if (MODE == 8) tc::mma_bf16_ss(tm + 192, tc::desc_advance(d_ds, (k & 7) * 2048), tc::desc_advance(d_kmn, (k & 7) * 2048), id_q64, 1);
            }
          }
        }
        tc::mma_commit(&bar);
      }
      __syncwarp();
      tc::mbar_wait(&bar, trial & 1);
      const unsigned long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = best;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc::fence_after_sync(); tc::tmem_dealloc(tm, 512); }
}

template <int MODE>
void run(const char* name, unsigned long long* dout, int grid, double per) {
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  bench<MODE><<<grid, 128, 160 * 1024>>>(dout);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned long long> h(grid);
  cudaMemcpy(h.data(), dout, grid * 8, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (auto v : h) mean += double(v) / grid;
  printf("%-72s grid %3d: %8.1f clk per %s %s\n", name, grid, mean / REPS / per, per == 1 ? "sub-tile" : "MMA",
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  unsigned long long* dout;
  cudaMalloc(&dout, 1024 * 8);
  for (int grid : {1, 148}) {
    run<0>("current mix, 32-query sub-tile (4096 elements)", dout, grid, 1);
    run<1>("quarter design mix, 64-query sub-tile (8192 elements)", dout, grid, 1);
    run<2>("SS K-major N=32", dout, grid, 16);
    run<6>("SS K-major N=64", dout, grid, 16);
    run<7>("SS K-major N=128", dout, grid, 16);
    run<3>("TS, B MN-major (64B swizzle) N=32  [dV, dK]", dout, grid, 16);
    run<5>("TS, B identity N=16  [dBias]", dout, grid, 16);
    run<4>("SS, A MN-major (128B swizzle), B MN-major N=32  [dQ]", dout, grid, 16);
    run<8>("SS, A MN-major (128B swizzle), B MN-major (128B swizzle) N=64", dout, grid, 16);
  }
  return 0;
}
