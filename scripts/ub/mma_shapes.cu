// Micro-benchmark (measurement only, not part of the library): cycles per tcgen05.mma (kind::f16, bf16 operands,
// M = 128, K = 16 per instruction) as a function of N, of the A operand's home (smem descriptor vs TMEM) and of
// whether consecutive MMAs accumulate into the same TMEM columns or rotate over four accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vit-stability-neurodegeneration_b200/csrc \
//        scripts/ub/mma_shapes.cu -o gpurun_out/mma_shapes && gpurun_out/mma_shapes
#include <cstdio>
#include <vector>
#include "tc.cuh"

void vsn_set_error(const char*, ...) {}
void vsn_count_launch() {}

constexpr int REPS = 256;

template <int N, bool A_TMEM, int NACC>
__global__ void __launch_bounds__(128, 1) bench(unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
    const uint64_t da = tc::make_smem_desc_sw64(tc::smem_u32(smem), 16, 512);            // A: [128][32] K-major
    const uint64_t db = tc::make_smem_desc_sw64(tc::smem_u32(smem + 16384), 16, 512);    // B: [N<=256][32] K-major
    unsigned long long best = ~0ull;
    for (int trial = 0; trial < 5; ++trial) {
      __syncwarp();
      const unsigned long long t0 = clock64();
      if (tc::elect_one()) {
#pragma unroll 8
        for (int r = 0; r < REPS; ++r) {
          const uint32_t d = tm + 256 + (r % NACC) * 64;   // NACC > 1 only used with N <= 64
          if (A_TMEM) tc::mma_bf16_ts(d, tm + (r & 1) * 8, tc::desc_advance(db, (r & 1) * 32), idesc, r >= NACC);
          else tc::mma_bf16_ss(d, tc::desc_advance(da, (r & 1) * 32), tc::desc_advance(db, (r & 1) * 32), idesc, r >= NACC);
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

template <int N, bool A_TMEM, int NACC>
void run(const char* name, unsigned long long* dout, int grid) {
  cudaFuncSetAttribute(bench<N, A_TMEM, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  bench<N, A_TMEM, NACC><<<grid, 128, 48 * 1024>>>(dout);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned long long> h(grid);
  cudaMemcpy(h.data(), dout, grid * 8, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (auto v : h) mean += double(v) / grid;
  printf("%-34s grid %3d: %7.1f clk per MMA (%.0f clk total, floor %d) %s\n", name, grid, mean / REPS, mean, 128 * N / 256,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  unsigned long long* dout;
  cudaMalloc(&dout, 1024 * 8);
  for (int grid : {1, 148}) {
    run<16, false, 1>("SS N=16  same accumulator", dout, grid);
    run<32, false, 1>("SS N=32  same accumulator", dout, grid);
    run<32, false, 4>("SS N=32  4 accumulators", dout, grid);
    run<64, false, 1>("SS N=64  same accumulator", dout, grid);
    run<64, false, 4>("SS N=64  4 accumulators", dout, grid);
    run<128, false, 1>("SS N=128 same accumulator", dout, grid);
    run<256, false, 1>("SS N=256 same accumulator", dout, grid);
    run<16, true, 1>("TS N=16  same accumulator", dout, grid);
    run<32, true, 1>("TS N=32  same accumulator", dout, grid);
    run<32, true, 4>("TS N=32  4 accumulators", dout, grid);
    run<64, true, 1>("TS N=64  same accumulator", dout, grid);
    run<128, true, 1>("TS N=128 same accumulator", dout, grid);
    run<256, true, 1>("TS N=256 same accumulator", dout, grid);
  }
  return 0;
}
