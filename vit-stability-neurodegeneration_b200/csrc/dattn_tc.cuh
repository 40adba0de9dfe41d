// Internal interface between attn.cu (C-ABI entry points) and dattn_tc.cu (tcgen05/TMEM/TMA kernels for dense
// multi-head attention with head_dim 64: the ViT-3D blocks, models/vit_3d.py:129-141).
#pragma once
#include "common.cuh"

struct DenseAttnArgs {
  const bf16* qkv;      // [S*N, 3C]
  bf16* out;            // [S*N, C]      forward output (backward: forward output, read for delta)
  float* lse;           // [S, heads, Npad] log2-domain logsumexp per row (written by forward, read by backward)
  // backward only
  const bf16* dout;     // [S*N, C]
  bf16* dqkv;           // [S*N, 3C]
  float* delta;         // [S, heads, Npad] scratch: rowsum(dO * O)
  float* dq_acc;        // [S*N, C] fp32 scratch: dQ summed over the key-tile CTAs (zeroed by the library)
  int S, N, Npad, heads, C;
  float scale;
};

bool dattn_tc_supported(int hd);
int dattn_tc_fwd(const DenseAttnArgs& a, cudaStream_t stream);
int dattn_tc_bwd(const DenseAttnArgs& a, cudaStream_t stream);
