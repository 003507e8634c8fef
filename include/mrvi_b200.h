/*
 * mrvi_b200.h -- C ABI of the B200-native counterfactual local-sample-distance path.
 *
 * The reference (YosefLab/mrvi) has no FFI: the path is the Python method
 * MrVI.get_local_sample_distances (src/mrvi/_model.py:627-714), whose device work is
 *   mapped_inference_fn(...)  -> z[B,S,L]          (src/mrvi/_model.py:326-365)
 *   _compute_distances_from_representations(...)   (src/mrvi/_model.py:542-584)
 *   the per-category running sums                  (src/mrvi/_model.py:469-504)
 * executed by JAX/XLA.  This header is what a binding for that device work would bind: plain
 * pointers and sizes, no torch / jax types.  The Python host side (mrvi_b200/_model.py) mirrors
 * the reference method and calls these entry points through ctypes; INTEGRATION.md shows the
 * stub a maintainer would add to the reference.
 *
 * Conventions: every function returns 0 on success or a negative MrviStatus; nothing throws or
 * aborts; the message of the last failure on the calling thread is mrvi_last_error().  The caller
 * owns all I/O buffers.  One handle per GPU; a handle is not thread-safe; handles are independent.
 */
#ifndef MRVI_B200_H_
#define MRVI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRVI_B200_ABI_VERSION 5

typedef struct MrviHandle MrviHandle;

typedef enum MrviStatus {
  MRVI_OK = 0,
  MRVI_ERR_INVALID_ARG = -1,   /* bad dims / null pointer / unsupported value            */
  MRVI_ERR_PARAM = -2,         /* missing parameter, wrong element count                 */
  MRVI_ERR_CUDA = -3,          /* CUDA runtime error (message has the cudaError string)  */
  MRVI_ERR_NO_DEVICE = -4,     /* no sm_100 device / device index out of range           */
  MRVI_ERR_UNSUPPORTED = -5    /* hyper-parameters outside what the kernels are built for */
} MrviStatus;

/* Hyper-parameters that shape the path: MrVAE fields (src/mrvi/_module.py:209-230) and the
 * EncoderUZ / AttentionBlock fields (src/mrvi/_module.py:114-128, _components.py:165-179). */
typedef struct MrviDims {
  int32_t n_input;          /* G  genes                                                   */
  int32_t n_sample;         /* S  samples                                                 */
  int32_t n_latent;         /* L                                                          */
  int32_t n_latent_u;       /* Lu, or 0 when u and z are isomorphic (Lu == L)             */
  int32_t encoder_n_hidden; /* H  (default 128)                                           */
  int32_t encoder_n_layers; /* ResnetBlocks in NormalDistOutputNN (default 2)             */
  int32_t n_latent_sample;  /* E  outerprod_dim (default 16)                              */
  int32_t n_channels;       /* C  (default 4) == attention head_dim                       */
  int32_t n_heads;          /* nh (default 2)                                             */
  int32_t qz_n_hidden;      /* M  ResnetBlock output width inside the qz MLPs (default 32) */
  int32_t qz_n_layers;      /* ResnetBlocks in the second qz MLP (default 1)              */
  int32_t use_map;          /* 1: qz emits L values; 0: 2L, first L kept (_module.py:326-329) */
} MrviDims;

/* Arithmetic of the dense contractions (fixed at create time). */
typedef enum MrviPrecision {
  MRVI_PRECISION_FP32 = 0, /* fp32 FFMA everywhere; parity bar 1e-5                       */
  MRVI_PRECISION_TC = 1    /* tcgen05 tensor cores, 16-bit/TF32 operands, fp32 accumulate; parity bar 1e-3 */
} MrviPrecision;

/* norm of _compute_distances_from_representations (src/mrvi/_model.py:548-556) */
typedef enum MrviNorm { MRVI_NORM_L2 = 0, MRVI_NORM_L1 = 1, MRVI_NORM_LINF = 2 } MrviNorm;

/* ABI version of the loaded library (compare with MRVI_B200_ABI_VERSION). */
int mrvi_abi_version(void);

/* Message of the last error raised on this thread ("" if none). Never NULL. */
const char* mrvi_last_error(void);

/* Number of CUDA devices visible to the library (0 if none). */
int mrvi_device_count(void);

/* Number of kernels this library has launched in this process (for bench accounting). */
int64_t mrvi_kernel_launch_count(void);

/*
 * Build a handle from the trained parameters {"params": module.params} (src/mrvi/_model.py:324),
 * flattened to n_params named float32 arrays in HOST memory: names[i] is the flax path
 * ("qu/Dense_0/kernel", ...; schema in mrvi_b200/_params.py and DESIGN.md), numels[i] its
 * element count.  Weights are copied and re-packed on `device`; per-sample tables (LayerNorm'd
 * sample embeddings, attention K/V rows) are precomputed there.
 */
int mrvi_create(const MrviDims* dims, int32_t n_params, const char* const* names,
                const float* const* values, const int64_t* numels, int32_t device,
                int32_t precision, MrviHandle** out);

int mrvi_destroy(MrviHandle* h);

/*
 * One pass of the path over n_cells cells with every buffer in DEVICE memory on the handle's
 * device; all work is enqueued on `stream` (a cudaStream_t, NULL = default stream) and the call
 * returns without synchronising.
 *
 *   X            [n_cells, ldx] float32 counts, row-major, ldx >= n_input  (array_dict["X"], _model.py:378)
 *   sample_idx   [n_cells] int32 in [0, S): each cell's own sample code    (array_dict["sample"])
 *   n_keys       number of groupby keys (may be 0)
 *   group_codes  n_keys device pointers, each [n_cells] int32 in [0, n_groups[k]) or -1 (= in no group)
 *   n_groups     HOST array [n_keys]
 *   norm         MrviNorm
 *   cell_out     [n_cells, S, S] float32 or NULL          (the "cell" reduction, _model.py:485-491)
 *   group_sums   n_keys device pointers, each [n_groups[k], S, S] float64; the per-category sums
 *                of _model.py:469-484 are ADDED into them (caller zeroes; lets shards accumulate)
 *   z_out        [n_cells, S, L] float32 or NULL          (mean representations, _model.py:396-404)
 *   u_out        [n_cells, Lq]   float32 or NULL          (qu.mean, _module.py:312-314)
 */
int mrvi_run_device(MrviHandle* h, int64_t n_cells, const float* X, int64_t ldx,
                    const int32_t* sample_idx, int32_t n_keys, const int32_t* const* group_codes,
                    const int32_t* n_groups, int32_t norm, float* cell_out,
                    double* const* group_sums, float* z_out, float* u_out, void* stream);

/*
 * Same pass with every buffer in HOST memory (pinned or pageable).  Cells are streamed through
 * the GPU in chunks of `chunk_cells` (0 = library default) with the host->device copy of chunk
 * i+1 and the device->host copy of chunk i-1 overlapping the kernels of chunk i.  Returns after
 * all results are in the host buffers.  group_sums[k] is OVERWRITTEN with the float64 sums over
 * these n_cells; group_counts[k] ([n_groups[k]] int64, may be NULL) with the number of cells per code.
 */
int mrvi_run_host(MrviHandle* h, int64_t n_cells, const float* X, int64_t ldx,
                  const int32_t* sample_idx, int32_t n_keys, const int32_t* const* group_codes,
                  const int32_t* n_groups, int32_t norm, float* cell_out,
                  double* const* group_sums, int64_t* const* group_counts, float* z_out,
                  float* u_out, int64_t chunk_cells);

/*
 * mrvi_run_host with X given as a host CSR matrix (scipy.sparse.csr_matrix layout: indptr [n_cells+1]
 * int64, indices [nnz] int32 column ids in [0, n_input), data [nnz] float32).  Replaces the host-side
 * densify of scvi's AnnDataLoader (src/mrvi/_model.py:317-319): only the non-zeros cross PCIe, each
 * chunk is expanded to dense rows on the GPU.  Other arguments as mrvi_run_host.
 */
int mrvi_run_host_csr(MrviHandle* h, int64_t n_cells, const int64_t* indptr, const int32_t* indices,
                      const float* data, const int32_t* sample_idx, int32_t n_keys,
                      const int32_t* const* group_codes, const int32_t* n_groups, int32_t norm,
                      float* cell_out, double* const* group_sums, int64_t* const* group_counts,
                      float* z_out, float* u_out, int64_t chunk_cells);

/*
 * Stage (3) alone on device buffers: pairwise distances of given representations
 * z [n_cells, S, L] (the seam of _compute_distances_from_representations, _model.py:542-584),
 * with the same outputs as mrvi_run_device.  Used by the parity tests to isolate the stage.
 */
int mrvi_distances_device(MrviHandle* h, int64_t n_cells, const float* z, int32_t n_keys,
                          const int32_t* const* group_codes, const int32_t* n_groups,
                          int32_t norm, float* cell_out, double* const* group_sums, void* stream);

/*
 * The sampled paths of the same method: get_local_sample_distances(use_mean=False) and
 * (normalize_distances=True), i.e. the "sampled_distances" / "normalized_distances" / "sampled_representations"
 * inputs of compute_local_statistics (src/mrvi/_model.py:406-450, 508-540), with every buffer in HOST memory.
 *
 * For every cell, mc_samples draws of u ~ q(u|x) are taken PER COUNTERFACTUAL SAMPLE (each vmapped call owns a PRNG
 * stream, _model.py:136-159, _module.py:315-318), z_s is evaluated on each draw and the S x S distances are formed
 * per (cell, draw).  With normalize != 0 they are z-scored per cell against the local baseline -- mean and variance
 * of the distance between 2*mc_samples paired draws of z for the cell's own sample (_model.py:508-540) -- and
 * averaged over the draws (_model.py:447-450; norm must be MRVI_NORM_L2 then, _model.py:436-439).
 *
 *   u_noise     [S][mc][n_cells][Lq] float32 standard-normal draws, or NULL: generated on the device by
 *               Philox4x32-10 from (seed, cell index, draw, sample) -- independent of chunking and sharding;
 *               cell_offset is the global index of cell 0 (a shard passes its offset so that shards reproduce
 *               the single-GPU result)
 *   base_noise  [2*mc][n_cells][Lq] draws of the baseline (normalize != 0 only), or NULL as above
 *   cell_out    normalize ? [n_cells, S, S] : [n_cells, mc, S, S]   float32 or NULL
 *   group_sums  normalize ? [n_groups[k], S, S] : [n_groups[k], mc, S, S]   float64, OVERWRITTEN
 *   z_out       [n_cells, mc, S, L] float32 or NULL ("sampled_representations")
 *   u_out       [n_cells, mc, S, Lq] float32 or NULL: the draws u = mean + scale * noise themselves (row (cell, 0, own sample)
 *               with mc = 1 is get_latent_representation(use_mean=False), _model.py:221-271)
 *   eps_noise   use_map=0 models only (qz_kwargs={"use_map": False}): eps ~ Normal(loc, softplus(raw) + 1e-3) is then DRAWN
 *               too (_module.py:326-329).  [S][mc][n_cells][n_latent] draws, or NULL = Philox (stream 2);
 *               base_eps_noise [2*mc][n_cells][n_latent] for the baseline rows (stream 3).  Ignored when use_map=1.
 *               Sampled eps is implemented in MRVI_PRECISION_FP32 only (MRVI_ERR_UNSUPPORTED in tensor-core mode).
 * JAX's threefry streams are not reproduced: with explicit draws the result is exact (oracle parity), against the
 * reference itself it is statistical.
 */
int mrvi_run_sampled_host(MrviHandle* h, int64_t n_cells, const float* X, int64_t ldx,
                          const int32_t* sample_idx, int32_t mc_samples, uint64_t seed, int64_t cell_offset,
                          const float* u_noise, const float* base_noise, const float* eps_noise,
                          const float* base_eps_noise, int32_t n_keys,
                          const int32_t* const* group_codes, const int32_t* n_groups, int32_t norm,
                          int32_t normalize, float* cell_out, double* const* group_sums,
                          int64_t* const* group_counts, float* z_out, float* u_out, int64_t chunk_cells);

/*
 * The latent-space stage of MrVI.differential_expression (src/mrvi/_model.py:1160-1218, 1283), host
 * buffers.  For every cell, mc_samples draws of u per counterfactual sample (as in mrvi_run_sampled_host) give the
 * residuals eps_ [mc, S_kept, L] = inference(..., use_mean=False)["eps"] (_model.py:1175-1194); per (draw, latent dim)
 * they are standardised over the kept samples (population std, 1e-6 added to it), and
 *   betas[a, k, d]      = sum_s Amat[cell][k, s] eps[a, s, d]                    (_model.py:1203)
 *   betas_norm[a, l, d] = sum_k betas[a, k, d] prefactor[cell][k, l]             (_model.py:1206)
 *   effect_size[cell, l] = mean_a sum_d betas_norm[a, l, d]^2                    (_model.py:1207)
 *   beta[cell, k, d]     = mean_a betas[a, k, d] * std[a, d]                     (_model.py:1210, 1283)
 * The design-matrix algebra (Amat = pinv(X^T D X + lambd I) X^T D, prefactor = sqrtm(.), _model.py:1152-1158), the chi2
 * p-values (_model.py:1208) and the Benjamini-Hochberg adjustment (_model.py:1363) stay on the host (mrvi_b200._model).
 *
 *   kept_samples [n_kept_samples] sample codes (sample_subset, _model.py:1112-1119), or NULL = all n_sample samples
 *   Amat         per_cell_design ? [n_cells][n_covariates][n_kept] : [n_covariates][n_kept]      float32
 *   prefactor    per_cell_design ? [n_cells][n_covariates][n_covariates] : [n_covariates][n_covariates]
 *   beta_out     [n_cells][n_covariates][n_latent] float32;  effect_size_out [n_cells][n_covariates] float32
 *   eps_out      [n_cells][mc][n_sample][n_latent] float32 or NULL (all samples, unstandardised: for the parity tests)
 *   u_noise / eps_noise / seed / cell_offset: as in mrvi_run_sampled_host.
 * The decoder-side outputs of the reference method (lfc, pde, baseline_expression) are outside this library.
 */
int mrvi_run_de_host(MrviHandle* h, int64_t n_cells, const float* X, int64_t ldx, const int32_t* sample_idx,
                     int32_t mc_samples, uint64_t seed, int64_t cell_offset, const float* u_noise,
                     const float* eps_noise, int32_t n_kept_samples, const int32_t* kept_samples, int32_t n_covariates, const float* Amat,
                     const float* prefactor, int32_t per_cell_design, float* beta_out, float* effect_size_out,
                     float* eps_out, int64_t chunk_cells);

/*
 * Diagnostics of the tensor-core mode's attention table: number of intervals (0 = no table, the
 * softmax is evaluated directly), the tabulated range of t = alpha_h q_i + beta_h (the whole range reachable through
 * LayerNorm(u)), and the verified maximum interpolation error relative to max_j |kv_s[j]|.  Any pointer may be NULL.
 */
int mrvi_attention_table_info(MrviHandle* h, int32_t* n_intervals, float* t_min, float* t_max, float* max_rel_err);

/*
 * mrvi_run_host for a dense count matrix stored as integers (raw counts in adata.X or a layer): X [n_cells, ldx]
 * elements of x_dtype, ldx >= n_input in ELEMENTS.  uint8 / uint16 matrices cross PCIe as they are (1 or 2 bytes per
 * value, no host pass); int32 / int64 matrices are narrowed chunk by chunk on the host to uint8 / uint16 (to float32
 * when a chunk holds a negative value or one >= 65536) -- no float32 copy of the whole matrix is ever made.  The
 * values are widened to float32 on the GPU, i.e. the result is identical to mrvi_run_host on X.astype(float32) -- what
 * scvi's loader would hand the reference (src/mrvi/_model.py:317-319, 377-384).  Other arguments as mrvi_run_host.
 */
typedef enum MrviXDtype { MRVI_X_UINT8 = 1, MRVI_X_UINT16 = 2, MRVI_X_INT32 = 3, MRVI_X_INT64 = 4 } MrviXDtype;
int mrvi_run_host_counts(MrviHandle* h, int64_t n_cells, const void* X, int32_t x_dtype, int64_t ldx,
                         const int32_t* sample_idx, int32_t n_keys, const int32_t* const* group_codes,
                         const int32_t* n_groups, int32_t norm, float* cell_out, double* const* group_sums,
                         int64_t* const* group_counts, float* z_out, float* u_out, int64_t chunk_cells);

/*
 * mrvi_run_host narrows dense count chunks losslessly before the host->device copy: a chunk whose float32 values are
 * all integers in [0, 255] ([0, 65535]) crosses PCIe as uint8 (uint16) and is widened back to the identical float32
 * values on the GPU; any other chunk is copied as it is (env MRVI_B200_NO_HOST_PACK=1 disables the narrowing).
 * When X is page-locked, the copy engine pulls the trailing rows of every chunk raw while the host threads narrow the
 * leading rows, in the proportion of the measured host rate (env MRVI_B200_NO_SPLIT_H2D=1: all rows narrowed;
 * MRVI_B200_PCIE_GBPS: assumed raw rate of the link, default 47).  Results are bit-identical on every route.  With four
 * or more GPUs fed from one host, float32 input travels raw.
 * mrvi_last_h2d_bytes: bytes of X the most recent mrvi_run_host call actually copied host->device.
 * mrvi_narrow_counts: the host-side narrowing itself (no GPU needed; exposed for tests): writes dense
 * [n_rows][n_cols] uint8 / uint16 to `out` (capacity 2 * n_rows * n_cols bytes) and returns 1 / 2, or 0 when the
 * rows are not narrowable.
 */
int64_t mrvi_last_h2d_bytes(MrviHandle* h);
int mrvi_narrow_counts(const float* X, int64_t ldx, int64_t n_rows, int32_t n_cols, void* out, int32_t n_threads);

/*
 * Multi-GPU: cells shard over ranks in contiguous ranges, every rank runs mrvi_run_device on its
 * shard with its own handle, and the ONLY exchange step of the path is the sum of the per-category accumulators of
 * src/mrvi/_model.py:469-484 over ranks: one ncclAllReduce(sum, float64) per key, grouped into a single NCCL launch,
 * over NVLink / NVSwitch.  NCCL is resolved at run time (dlopen of libnccl.so.2: no link-time dependency).
 *
 *   mrvi_comm_unique_id   fills 128 bytes (an ncclUniqueId) on one rank; the caller distributes them to every rank
 *   mrvi_comm_create      ncclCommInitRank on the handle's device; the communicator belongs to the handle
 *   mrvi_allreduce_groups in-place all-reduce of group_sums[k] ([n_groups[k], S, S] float64, DEVICE memory) for every
 *                         key, enqueued on `stream`; nccl_comm: an ncclComm_t of the caller, or NULL = the handle's own
 */
int mrvi_comm_unique_id(void* out128);
int mrvi_comm_create(MrviHandle* h, const void* id128, int32_t n_ranks, int32_t rank);
int mrvi_allreduce_groups(MrviHandle* h, void* nccl_comm, int32_t n_keys, double* const* group_sums, const int32_t* n_groups,
                          void* stream);

/* Page-locked host memory for the big outputs (keep_cell blocks, representations): device->host copies into it run at
 * PCIe speed and overlap the kernels; the Python host side allocates `cell_out` / `z_out` through these. */
int mrvi_alloc_pinned(int64_t bytes, void** out);
int mrvi_free_pinned(void* p);

/* Timing taps: device milliseconds spent in each stage of the most recent mrvi_run_* call
 * (CUDA events on the launching stream; valid after that stream is synchronised).
 * out_ms[0]=encoder (x->u), [1]=q(z|u,s) chain, [2]=distances+reductions, [3]=whole call. */
int mrvi_last_stage_ms(MrviHandle* h, float out_ms[4]);

#ifdef __cplusplus
}
#endif
#endif /* MRVI_B200_H_ */
