// Internal interface between attn.cu (C-ABI entry points, mma.sync kernels for dense / large windows) and
// wattn_tc.cu (tcgen05/TMEM window-attention kernels for head_dim 32 and windows of at most 252 tokens).
#pragma once
#include "common.cuh"

struct WinAttnArgs {
  const bf16* qkv;      // [T, 3C]
  bf16* out;            // [T, C]        forward output (backward: forward output, read for delta)
  float* lse;           // [S, heads, 256] natural-log logsumexp per row
  const float* table;   // [table_len, heads] relative_position_bias_table (may be null)
  int table_len;
  // backward only
  const bf16* dout;     // [T, C]
  bf16* dqkv;           // [T, 3C]
  float* dbias_dense;   // unused by the tcgen05 path (scratch of the mma.sync path)
  float* dtable;        // [table_len, heads] fp32 gradient of the bias table, accumulated (+=); may be null
  int S, N, heads, C;
  float scale;
  int B, Dp, Hp, Wp, wd, wh, ww, sd, sh, sw, nWd, nWh, nWw, use_mask;
};

// true when the tcgen05 kernels cover this problem (head_dim 32, window (6,7,6))
bool wattn_tc_supported(int wd, int wh, int ww, int hd);
int wattn_tc_fwd(const WinAttnArgs& a, cudaStream_t stream);
// delta [S, heads, 256] (scratch) = rowsum(dO * O); lse is the log2-domain logsumexp written by wattn_tc_fwd;
// the bias-table gradient is folded inside the kernel (no dense [heads,N,N] buffer).
int wattn_tc_bwd(const WinAttnArgs& a, float* delta, cudaStream_t stream);
