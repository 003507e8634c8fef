// Shared declarations of the mrvi_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/mrvi_b200.h"

namespace mrvi {

constexpr int kResnetHidden = 128;   // ResnetBlock.n_hidden default, reference _components.py:31
constexpr float kLnEps = 1e-6f;      // flax.linen.LayerNorm default epsilon
constexpr int kMaxKeys = 8;          // groupby keys per call
constexpr int kMaxE = 32;            // n_latent_sample (outerprod_dim) supported by the kernels
constexpr int kMaxHeadDim = 8;
constexpr int kMaxHeads = 8;
// cudaFuncAttributeMaxDynamicSharedMemorySize is a process-wide, per-device property of a kernel, not of a handle:
// it is always raised to the device limit, never to a handle-specific size (a second handle with smaller dims
// must not lower it under the first one's launches)
constexpr int kMaxSmemOptin = 227 * 1024;

// Row-major packed weight: [Kpad][ld] with Kpad = roundup(K,4), ld = roundup(N,4); zero padding.
struct DenseW {
  const float* w = nullptr;  // device
  const float* b = nullptr;  // device, [ld] zero padded
  int K = 0, N = 0, Kpad = 0, ld = 0;
};

struct LnW {
  const float* scale = nullptr;  // device [N] or null
  const float* bias = nullptr;
  int N = 0;
};

// ResnetBlock (reference _components.py:27-51)
struct ResnetW {
  DenseW d0;      // n_in -> 128
  LnW ln0;        // 128
  DenseW dskip;   // n_in -> 128, w == nullptr when n_in == 128 (identity skip)
  DenseW dout;    // 128 -> n_out
  LnW ln1;        // n_out
};

constexpr int kMaxBlocks = 4;  // ResnetBlocks per MLP supported

// Everything the fp32 kernels read, in device memory.  Passed by value to kernels.
struct NetW {
  int G, S, L, Lq, H, E, C, nh, M, out_dim, has_zbase, n_enc_blocks, n_qz_blocks;
  // ---- qu (_EncoderXU, reference _module.py:173-202)
  DenseW enc0, enc1;
  const float* gamma[2];  // [S][H]
  const float* beta[2];   // [S][H]
  const float* sample_effect;  // [S][H]
  ResnetW enc_blocks[kMaxBlocks];
  DenseW enc_mean;  // H -> Lq
  DenseW enc_scale; // H -> Lq, softplus head of q(u|x) (reference _components.py:96); w == nullptr when not supplied
  // ---- qz (EncoderUZ + AttentionBlock)
  LnW u_ln;
  DenseW q_tok;   // Lq -> E, no bias   (DenseGeneral_0)
  DenseW zbase;   // Lu -> L            (qz/Dense_0), only if has_zbase
  const float* wq;  // [nh][hd] query kernel (input dim 1)
  const float* bq;  // [nh][hd]
  const float* wo;  // [nh][hd][C]
  const float* bo;  // [C]
  const float* ktab;  // [S][E][nh][hd]  per-sample key rows
  const float* vtab;  // [S][E][nh][hd]  per-sample value rows
  ResnetW mlp0_block;       // E*C -> 128 -> M
  DenseW mlp0_out;          // M -> E
  ResnetW mlp1_blocks[kMaxBlocks];  // (Lq+E | M) -> 128 -> M
  DenseW mlp1_out;          // M -> out_dim
};

// Per-cell intermediates written by the encoder and read by the q(z|u,s) chain.
struct CellBuf {
  float* u;      // [n][Lq]   qu.mean
  float* u_ln;   // [n][Lq]   LayerNorm(u)                (reference _module.py:142)
  float* q_tok;  // [n][E]    DenseGeneral((E,1))(u_ln)   (reference _components.py:195-197)
  float* zbase;  // [n][L]    Dense(L)(u) or u            (reference _module.py:166-170)
  // ---- sampled paths only (use_mean=False / normalize_distances=True, reference _module.py:315-318)
  float* u_scale;              // [n][Lq] scale of q(u|x), written by the encoder when non-null
  int per_row;                 // 1: u_ln / q_tok / zbase are indexed by ROW (cell * rows_per_cell + s): every
                               //    (cell, s) row carries its own draw of u
  const int32_t* cell_sample;  // non-null: all rows of cell c use the per-sample tables of sample cell_sample[c]
                               //    (the local baseline of reference _model.py:508-540 evaluates the cell's own sample)
  float* raw_out;              // non-null (use_map=False, sampled paths, k_qz_fp32 only): [rows][L] second half of the last
                               //    Dense, the pre-softplus scale of q(eps) (reference _module.py:326-329)
};

// Group-reduction plan for one chunk of cells (device pointers).
struct GroupPlan {
  int n_keys;
  const int32_t* codes[kMaxKeys];  // [n] original order, -1 = none
  double* sums[kMaxKeys];          // [n_groups][S][S]
  const int32_t* order;            // [n] cell index of sorted position i (null when n_keys == 0)
  const int32_t* comp_sorted;      // [n] composite code at sorted position i
  const int32_t* n_groups;         // [n_keys] device copy of the group counts: codes outside [0, n_groups[k]) never reach the sums
};

__device__ __forceinline__ float gelu_tanh_f(float v) {
  // jax.nn.gelu(approximate=True)
  const float c = 0.7978845608028654f;
  return 0.5f * v * (1.0f + tanhf(c * (v + 0.044715f * v * v * v)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline int64_t round_up64(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

}  // namespace mrvi
