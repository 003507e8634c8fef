// C-ABI implementation of include/mrvi_b200.h: handle, weight packing, launch orchestration.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cub/device/device_radix_sort.cuh>

#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <thread>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_pack.h"
#include "fp32_kernels.cuh"
#include "tc_kernels.cuh"
#include "tc_encoder.cuh"
#include "tc_fused.cuh"
#include "tc_dist.cuh"
#include "sampled_kernels.cuh"

namespace mrvi {

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_dev_handles[64];    // live handles per device of this process: GPUs fed by one process share the host's cores and memory bus
static int live_devices() {
  int n = 0;
  for (int d = 0; d < 64; ++d) n += g_dev_handles[d].load() > 0;
  return n;
}

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(MRVI_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,           \
                  cudaGetErrorString(_e));                                                      \
  } while (0)

#define LAUNCH_CHECK()                                                                          \
  do {                                                                                          \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                         \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess)                                                                      \
      return fail(MRVI_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__,       \
                  cudaGetErrorString(_e));                                                      \
  } while (0)

// ------------------------------------------------------------------------------------------
struct DeviceArena {  // owns every device allocation of a handle
  std::vector<void*> ptrs;
  ~DeviceArena() { for (void* p : ptrs) cudaFree(p); }
  cudaError_t alloc(void** out, size_t bytes) {
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
    if (e == cudaSuccess) ptrs.push_back(*out);
    return e;
  }
  void release(void* p) {
    for (size_t i = 0; i < ptrs.size(); ++i)
      if (ptrs[i] == p) { cudaFree(p); ptrs.erase(ptrs.begin() + i); return; }
  }
};

struct HostStage {  // device-side staging of mrvi_run_host, cached across calls
  int64_t chunk = 0;
  int n_keys = 0;
  bool has_cell = false, has_z = false, has_u = false;
  float* X[2] = {nullptr, nullptr};
  uint8_t* dpack[2] = {nullptr, nullptr};     // narrowed counts on the device (host_pack.cpp)
  uint8_t* hpack[2] = {nullptr, nullptr};     // pinned host staging of the narrowed chunk
  size_t pack_bytes = 0, dpack_bytes = 0;
  std::unique_ptr<PackPool> pool;
  bool pack_unavailable = false; // pinned staging could not be allocated
  int pack_skipped = 0;          // chunks sent raw since the rate fell below the threshold
  double pack_rate_gbps = 0.0;   // measured host narrowing rate (float32 bytes consumed per second), smoothed
  int64_t* csr_ptr[2] = {nullptr, nullptr};   // CSR staging (mrvi_run_host_csr)
  int32_t* csr_idx[2] = {nullptr, nullptr};
  float* csr_val[2] = {nullptr, nullptr};
  int64_t csr_nnz_cap = 0;
  int32_t* sid[2] = {nullptr, nullptr};
  int32_t* codes[2][kMaxKeys] = {};
  float* cell[2] = {nullptr, nullptr};
  float* z[2] = {nullptr, nullptr};
  float* u[2] = {nullptr, nullptr};
  double* sums[kMaxKeys] = {};
  size_t sums_elems[kMaxKeys] = {};
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[2] = {}, ev_comp[2] = {}, ev_d2h[2] = {};
  std::vector<void*> owned;
  void free_buffers() {
    for (void* p : owned) cudaFree(p);
    owned.clear();
    for (int i = 0; i < 2; ++i) { if (hpack[i]) cudaFreeHost(hpack[i]); hpack[i] = nullptr; dpack[i] = nullptr; }
    pack_bytes = 0;
    dpack_bytes = 0;
    // nothing below may survive a failed re-allocation: a later call must see "no buffers", never stale pointers
    chunk = 0; n_keys = 0; has_cell = has_z = has_u = false; csr_nnz_cap = 0;
    for (int i = 0; i < 2; ++i) {
      X[i] = nullptr; sid[i] = nullptr; cell[i] = nullptr; z[i] = nullptr; u[i] = nullptr;
      csr_ptr[i] = nullptr; csr_idx[i] = nullptr; csr_val[i] = nullptr;
      for (int k = 0; k < kMaxKeys; ++k) codes[i][k] = nullptr;
    }
    for (int k = 0; k < kMaxKeys; ++k) { sums[k] = nullptr; sums_elems[k] = 0; }
  }
  ~HostStage() {
    free_buffers();
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_comp) cudaStreamDestroy(s_comp);
    if (s_d2h) cudaStreamDestroy(s_d2h);
    for (int i = 0; i < 2; ++i) {
      if (ev_h2d[i]) cudaEventDestroy(ev_h2d[i]);
      if (ev_comp[i]) cudaEventDestroy(ev_comp[i]);
      if (ev_d2h[i]) cudaEventDestroy(ev_d2h[i]);
    }
  }
};

}  // namespace mrvi

using namespace mrvi;

struct MrviHandle {
  MrviDims dims{};
  int device = 0;
  int precision = 0;
  int sm_count = 148;
  DeviceArena arena;
  NetW net{};
  TcNet tc{};
  EncTc enc_tc{};
  EncTail enc_tail{};
  int64_t last_h2d_bytes = 0;  // bytes of X the last mrvi_run_host call actually copied host->device
  size_t dist_tc_smem = 0;  // k_dist_gram_tc (128 < S <= 256), 0 = unavailable
  float* h1buf = nullptr;  // [chunk][128] Dense_0 output of the tensor-core encoder GEMM
  // workspace for one chunk of cells
  int64_t chunk_cells = 0;
  CellBuf cb{};
  float* zbuf = nullptr;
  int32_t *comp_in = nullptr, *idx_in = nullptr, *comp_sorted = nullptr, *order = nullptr;
  int32_t* n_groups_dev = nullptr;
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  // launch geometry of the fp32 kernels
  int enc_ldh = 0, enc_ldhid = 0;
  size_t enc_smem = 0;
  int qz_lda = 0, qz_ldhid = 0, qz_ldm = 0, qz_ldc = 0;
  size_t qz_smem = 0;
  // stage timing
  std::vector<cudaEvent_t> ev;  // 4 per chunk: start, after enc, after qz, after dist
  int ev_used = 0;
  cudaEvent_t ev_call[2] = {nullptr, nullptr};
  HostStage hs;
  void* nccl_comm = nullptr;   // created by mrvi_comm_create (NCCL through dlopen: no link-time dependency)
  ~MrviHandle() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) if (ev_call[i]) cudaEventDestroy(ev_call[i]);
  }
};

namespace mrvi {

// ------------------------------------------------------------------------------------------
// parameter lookup + packing
struct ParamTable {
  std::map<std::string, std::pair<const float*, int64_t>> m;
  const float* get(const std::string& name, int64_t numel, int* err) const {
    auto it = m.find(name);
    if (it == m.end()) {
      *err = fail(MRVI_ERR_PARAM, "missing parameter '%s' (expected %lld elements)", name.c_str(), (long long)numel);
      return nullptr;
    }
    if (it->second.second != numel) {
      *err = fail(MRVI_ERR_PARAM, "parameter '%s' has %lld elements, expected %lld", name.c_str(),
                  (long long)it->second.second, (long long)numel);
      return nullptr;
    }
    return it->second.first;
  }
  bool has(const std::string& name) const { return m.count(name) != 0; }
};

int tc_upload(DeviceArena& arena, const void* host, size_t bytes, const void** dev) {
  void* d = nullptr;
  CUDA_TRY(arena.alloc(&d, bytes));
  if (bytes) CUDA_TRY(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  *dev = d;
  return 0;
}

static int upload(MrviHandle* h, const std::vector<float>& host, const float** out) {
  float* d = nullptr;
  CUDA_TRY(h->arena.alloc((void**)&d, host.size() * sizeof(float)));
  if (!host.empty()) CUDA_TRY(cudaMemcpy(d, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

// kernel [K][N] row-major (+ bias [N]) -> padded [Kpad][ld]
static int pack_dense(MrviHandle* h, const ParamTable& pt, const std::string& name, int K, int N,
                      bool has_bias, DenseW* out) {
  int err = 0;
  const float* w = pt.get(name + "/kernel", (int64_t)K * N, &err);
  if (!w) return err;
  const float* b = nullptr;
  if (has_bias) { b = pt.get(name + "/bias", N, &err); if (!b) return err; }
  out->K = K; out->N = N; out->Kpad = round_up(K, 4); out->ld = round_up(N, 4);
  std::vector<float> wp((size_t)out->Kpad * out->ld, 0.f), bp(out->ld, 0.f);
  for (int k = 0; k < K; ++k) memcpy(&wp[(size_t)k * out->ld], w + (size_t)k * N, N * sizeof(float));
  if (b) memcpy(bp.data(), b, N * sizeof(float));
  if ((err = upload(h, wp, &out->w))) return err;
  if ((err = upload(h, bp, &out->b))) return err;
  return 0;
}

static int pack_vec(MrviHandle* h, const ParamTable& pt, const std::string& name, int64_t n, const float** out) {
  int err = 0;
  const float* v = pt.get(name, n, &err);
  if (!v) return err;
  std::vector<float> tmp(v, v + n);
  return upload(h, tmp, out);
}

static int pack_ln(MrviHandle* h, const ParamTable& pt, const std::string& name, int N, LnW* out) {
  int err;
  out->N = N;
  if ((err = pack_vec(h, pt, name + "/scale", N, &out->scale))) return err;
  if ((err = pack_vec(h, pt, name + "/bias", N, &out->bias))) return err;
  return 0;
}

static int pack_resnet(MrviHandle* h, const ParamTable& pt, const std::string& prefix, int n_in, int n_out, ResnetW* rb) {
  int err;
  const int R = kResnetHidden;
  if ((err = pack_dense(h, pt, prefix + "/Dense_0", n_in, R, true, &rb->d0))) return err;
  if ((err = pack_ln(h, pt, prefix + "/LayerNorm_0", R, &rb->ln0))) return err;
  int k = 1;
  if (n_in != R) {
    if ((err = pack_dense(h, pt, prefix + "/Dense_1", n_in, R, true, &rb->dskip))) return err;
    k = 2;
  } else {
    rb->dskip = DenseW{};
  }
  if ((err = pack_dense(h, pt, prefix + "/Dense_" + std::to_string(k), R, n_out, true, &rb->dout))) return err;
  if ((err = pack_ln(h, pt, prefix + "/LayerNorm_1", n_out, &rb->ln1))) return err;
  return 0;
}

static int build_net(MrviHandle* h, const ParamTable& pt) {
  const MrviDims& d = h->dims;
  NetW& n = h->net;
  int err;
  const int G = d.n_input, S = d.n_sample, L = d.n_latent, H = d.encoder_n_hidden, E = d.n_latent_sample;
  const int C = d.n_channels, nh = d.n_heads, M = d.qz_n_hidden;
  const int Lq = d.n_latent_u > 0 ? d.n_latent_u : L;
  n.G = G; n.S = S; n.L = L; n.Lq = Lq; n.H = H; n.E = E; n.C = C; n.nh = nh; n.M = M;
  n.out_dim = d.use_map ? L : 2 * L;
  n.has_zbase = d.n_latent_u > 0;
  n.n_enc_blocks = d.encoder_n_layers;
  n.n_qz_blocks = d.qz_n_layers;
  if ((err = pack_dense(h, pt, "qu/Dense_0", G, H, true, &n.enc0))) return err;
  if ((err = pack_dense(h, pt, "qu/Dense_1", H, H, true, &n.enc1))) return err;
  for (int i = 0; i < 2; ++i) {
    std::string cn = "qu/ConditionalNormalization_" + std::to_string(i);
    if ((err = pack_vec(h, pt, cn + "/gamma_conditional/embedding", (int64_t)S * H, &n.gamma[i]))) return err;
    if ((err = pack_vec(h, pt, cn + "/beta_conditional/embedding", (int64_t)S * H, &n.beta[i]))) return err;
  }
  if ((err = pack_vec(h, pt, "qu/Embed_0/embedding", (int64_t)S * H, &n.sample_effect))) return err;
  for (int b = 0; b < d.encoder_n_layers; ++b)
    if ((err = pack_resnet(h, pt, "qu/NormalDistOutputNN_0/ResnetBlock_" + std::to_string(b), H, H, &n.enc_blocks[b]))) return err;
  if ((err = pack_dense(h, pt, "qu/NormalDistOutputNN_0/Dense_0", H, Lq, true, &n.enc_mean))) return err;
  n.enc_scale = DenseW{};
  if (pt.has("qu/NormalDistOutputNN_0/Dense_1/kernel")) {  // optional: only the sampled paths read the scale head
    if ((err = pack_dense(h, pt, "qu/NormalDistOutputNN_0/Dense_1", H, Lq, true, &n.enc_scale))) return err;
  }
  if ((err = pack_ln(h, pt, "qz/u_ln", Lq, &n.u_ln))) return err;
  const std::string ab = "qz/AttentionBlock_0";
  if ((err = pack_dense(h, pt, ab + "/DenseGeneral_0", Lq, E, false, &n.q_tok))) return err;  // kernel (Lq,E,1)
  if (n.has_zbase) { if ((err = pack_dense(h, pt, "qz/Dense_0", Lq, L, true, &n.zbase))) return err; }
  const std::string mha = ab + "/MultiHeadDotProductAttention_0";
  if ((err = pack_vec(h, pt, mha + "/query/kernel", nh * C, &n.wq))) return err;
  if ((err = pack_vec(h, pt, mha + "/query/bias", nh * C, &n.bq))) return err;
  if ((err = pack_vec(h, pt, mha + "/out/kernel", (int64_t)nh * C * C, &n.wo))) return err;
  if ((err = pack_vec(h, pt, mha + "/out/bias", C, &n.bo))) return err;
  if ((err = pack_resnet(h, pt, ab + "/MLP_0/ResnetBlock_0", E * C, M, &n.mlp0_block))) return err;
  if ((err = pack_dense(h, pt, ab + "/MLP_0/Dense_0", M, E, true, &n.mlp0_out))) return err;
  for (int b = 0; b < d.qz_n_layers; ++b)
    if ((err = pack_resnet(h, pt, ab + "/MLP_1/ResnetBlock_" + std::to_string(b), b == 0 ? Lq + E : M, M, &n.mlp1_blocks[b]))) return err;
  if ((err = pack_dense(h, pt, ab + "/MLP_1/Dense_0", M, n.out_dim, true, &n.mlp1_out))) return err;

  // per-sample tables on the device
  const float *embed, *ln_s, *ln_b, *wkv, *wk, *bk, *wv, *bv;
  if ((err = pack_vec(h, pt, "qz/Embed_0/embedding", (int64_t)S * E, &embed))) return err;
  if ((err = pack_vec(h, pt, "qz/sample_embed_ln/scale", E, &ln_s))) return err;
  if ((err = pack_vec(h, pt, "qz/sample_embed_ln/bias", E, &ln_b))) return err;
  if ((err = pack_vec(h, pt, ab + "/DenseGeneral_1/kernel", (int64_t)E * E, &wkv))) return err;
  if ((err = pack_vec(h, pt, mha + "/key/kernel", nh * C, &wk))) return err;
  if ((err = pack_vec(h, pt, mha + "/key/bias", nh * C, &bk))) return err;
  if ((err = pack_vec(h, pt, mha + "/value/kernel", nh * C, &wv))) return err;
  if ((err = pack_vec(h, pt, mha + "/value/bias", nh * C, &bv))) return err;
  float *ktab, *vtab, *kvtok;
  CUDA_TRY(h->arena.alloc((void**)&ktab, (size_t)S * E * nh * C * sizeof(float)));
  CUDA_TRY(h->arena.alloc((void**)&vtab, (size_t)S * E * nh * C * sizeof(float)));
  CUDA_TRY(h->arena.alloc((void**)&kvtok, (size_t)S * E * sizeof(float)));
  k_sample_tables<<<(S + 127) / 128, 128>>>(embed, ln_s, ln_b, wkv, wk, bk, wv, bv, S, E, nh, C, ktab, vtab, kvtok);
  LAUNCH_CHECK();
  CUDA_TRY(cudaDeviceSynchronize());
  n.ktab = ktab; n.vtab = vtab;
  h->tc.kvtok = kvtok;
  return 0;
}

static int setup_workspace(MrviHandle* h) {
  const NetW& n = h->net;
  // chunk: z buffer <= 2 GiB (of 180 GB), <= 262144 cells, multiple of 128 cells.  Long chunks matter for the
  // grouped reductions (every CTA flushes its run accumulators at least once per launch) and for the persistent
  // kernels' tails
  int64_t per_cell = (int64_t)n.S * n.L * sizeof(float);
  int64_t ch = (2048ll << 20) / per_cell;
  if (ch > 262144) ch = 262144;
  ch = ch / 128 * 128;
  if (ch < 128) ch = 128;
  h->chunk_cells = ch;
  CUDA_TRY(h->arena.alloc((void**)&h->cb.u, ch * n.Lq * sizeof(float)));
  CUDA_TRY(h->arena.alloc((void**)&h->cb.u_ln, ch * n.Lq * sizeof(float)));
  CUDA_TRY(h->arena.alloc((void**)&h->cb.q_tok, ch * n.E * sizeof(float)));
  CUDA_TRY(h->arena.alloc((void**)&h->cb.zbase, ch * n.L * sizeof(float)));
  CUDA_TRY(h->arena.alloc((void**)&h->zbuf, ch * per_cell));
  CUDA_TRY(h->arena.alloc((void**)&h->comp_in, ch * sizeof(int32_t)));
  CUDA_TRY(h->arena.alloc((void**)&h->idx_in, ch * sizeof(int32_t)));
  CUDA_TRY(h->arena.alloc((void**)&h->comp_sorted, ch * sizeof(int32_t)));
  CUDA_TRY(h->arena.alloc((void**)&h->order, ch * sizeof(int32_t)));
  CUDA_TRY(h->arena.alloc((void**)&h->n_groups_dev, kMaxKeys * sizeof(int32_t)));
  size_t tmp = 0;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp, h->comp_in, h->comp_sorted, h->idx_in, h->order, (int)ch));
  h->cub_tmp_bytes = tmp;
  CUDA_TRY(h->arena.alloc(&h->cub_tmp, tmp));

  // fp32 kernel geometry
  const int Hld = round_up(n.H, 4);
  bool enc_skip = false;
  for (int b = 0; b < n.n_enc_blocks; ++b) enc_skip |= (n.enc_blocks[b].dskip.w != nullptr);
  h->enc_ldh = std::max(Hld, kResnetHidden) + 4;
  h->enc_ldhid = std::max((enc_skip ? 2 * kResnetHidden : kResnetHidden), 64 + round_up(n.L, 4)) + 4;
  h->enc_smem = ((size_t)kTM * 36 + 2 * (size_t)kTM * h->enc_ldh + (size_t)kTM * h->enc_ldhid) * sizeof(float) +
                3 * kTM * sizeof(float*) + kTM * sizeof(int);
  h->qz_lda = std::max(round_up(n.E * n.C, 4), round_up(n.M, 4)) + 4;
  h->qz_ldhid = std::max(2 * kResnetHidden, round_up(n.out_dim, 4)) + 4;
  h->qz_ldm = round_up(n.M, 4) + 4;
  h->qz_ldc = round_up(n.Lq + n.E, 4) + 4;
  h->qz_smem = (size_t)kTM * (h->qz_lda + h->qz_ldhid + h->qz_ldm + h->qz_ldc) * sizeof(float);
  if (h->enc_smem > 227 * 1024 || h->qz_smem > 227 * 1024)
    return fail(MRVI_ERR_UNSUPPORTED, "hyper-parameters need %zu / %zu bytes of shared memory (> 227 KiB)",
                h->enc_smem, h->qz_smem);
  // k_dist_small keeps a warp's S x L tile in shared memory: above 48 KiB it needs the opt-in (see kMaxSmemOptin)
  if ((size_t)(kThreads / 32) * n.S * (n.L | 1) * sizeof(float) > (size_t)kMaxSmemOptin && n.S * n.S <= 1024)
    return fail(MRVI_ERR_UNSUPPORTED, "n_sample x n_latent tile needs more than 227 KiB of shared memory");
  CUDA_TRY(cudaFuncSetAttribute(k_dist_small<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  CUDA_TRY(cudaFuncSetAttribute(k_dist_small<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  CUDA_TRY(cudaFuncSetAttribute(k_dist_small<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  CUDA_TRY(cudaFuncSetAttribute(k_encoder_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  CUDA_TRY(cudaFuncSetAttribute(k_qz_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  return 0;
}

// ------------------------------------------------------------------------------------------
// stage launches (all on `st`)
static int launch_encoder(MrviHandle* h, const float* X, int64_t ldx, const int32_t* sid, int64_t n, cudaStream_t st) {
  const int grid = (int)((n + kTM - 1) / kTM);
  const int vec_ok = ((ldx % 4) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  if (h->precision == MRVI_PRECISION_TC && h->enc_tc.ready && h->enc_tail.ready) {
    enc_full_tc_launch(h->enc_tc, h->enc_tail, X, ldx, sid, n, h->net.G, h->cb, st);
    LAUNCH_CHECK();
    return 0;
  }
  const float* h1 = nullptr;
  if (h->precision == MRVI_PRECISION_TC && h->enc_tc.ready) {
    enc_tc_launch(h->enc_tc, X, ldx, n, h->net.G, h->h1buf, st);
    LAUNCH_CHECK();
    h1 = h->h1buf;
  }
  k_encoder_fp32<<<grid, kThreads, h->enc_smem, st>>>(h->net, X, ldx, sid, n, h->cb, h->enc_ldh, h->enc_ldhid, vec_ok, h1);
  LAUNCH_CHECK();
  return 0;
}

static int launch_qz(MrviHandle* h, const CellBuf& cb, int64_t n, int rows_per_cell, float* z, cudaStream_t st) {
  const int64_t rows = n * rows_per_cell;
  if (rows == 0) return 0;
  const int grid = (int)((rows + kTM - 1) / kTM);
  k_qz_fp32<<<grid, kThreads, h->qz_smem, st>>>(h->net, cb, n, rows_per_cell, z, h->qz_lda, h->qz_ldhid, h->qz_ldm, h->qz_ldc);
  LAUNCH_CHECK();
  return 0;
}

template <int NORM>
static int launch_dist_t(MrviHandle* h, const float* z, int64_t n, float* cell_out, const GroupPlan& gp, cudaStream_t st) {
  const int S = h->net.S, L = h->net.L;
  if (S * S <= 1024) {
    const int warps = kThreads / 32;
    // enough warps to fill the GPU, but keep runs long when grouping
    int64_t target_warps = (int64_t)h->sm_count * 4 * warps;
    int cpw = (int)std::max<int64_t>(1, (n + target_warps - 1) / target_warps);
    const int64_t n_warps = (n + cpw - 1) / cpw;
    const int grid = (int)((n_warps + warps - 1) / warps);
    const size_t smem = (size_t)warps * S * (L | 1) * sizeof(float);
    k_dist_small<NORM><<<grid, kThreads, smem, st>>>(z, n, S, L, cell_out, gp, cpw);
  } else {
    const int nb = (S + kDB - 1) / kDB;
    int64_t target = std::max<int64_t>(1, (int64_t)h->sm_count * 4 / (nb * nb));
    int cpc = (int)std::max<int64_t>(1, (n + target - 1) / target);
    dim3 grid((unsigned)((n + cpc - 1) / cpc), (unsigned)(nb * nb));
    const int Lp = round_up(L, 4);
    const size_t smem = (size_t)2 * kDB * Lp * sizeof(float);
    k_dist_tiled<NORM><<<grid, kThreads, smem, st>>>(z, n, S, L, cell_out, gp, cpc);
  }
  LAUNCH_CHECK();
  return 0;
}

static int launch_dist(MrviHandle* h, const float* z, int64_t n, int norm, float* cell_out, const GroupPlan& gp, cudaStream_t st) {
  if (cell_out == nullptr && gp.n_keys == 0) return 0;
  if (h->precision == MRVI_PRECISION_TC && h->dist_tc_smem && dist_tc_supported(h->net.S, h->net.L, norm)) {
    dist_tc_launch(z, n, h->net.S, h->net.L, cell_out, gp, h->sm_count, h->dist_tc_smem, st);
    LAUNCH_CHECK();
    return 0;
  }
  switch (norm) {
    case MRVI_NORM_L2: return launch_dist_t<0>(h, z, n, cell_out, gp, st);
    case MRVI_NORM_L1: return launch_dist_t<1>(h, z, n, cell_out, gp, st);
    case MRVI_NORM_LINF: return launch_dist_t<2>(h, z, n, cell_out, gp, st);
  }
  return fail(MRVI_ERR_INVALID_ARG, "norm %d not supported", norm);
}

// builds the sorted order of one chunk for the group reductions
static int plan_groups(MrviHandle* h, int64_t n, int n_keys, const int32_t* const* codes, int64_t offset,
                       const int32_t* n_groups, double* const* sums, GroupPlan* gp, cudaStream_t st) {
  memset(gp, 0, sizeof(*gp));
  gp->n_keys = n_keys;
  if (n_keys == 0) return 0;
  int64_t prod = 1;
  for (int k = 0; k < n_keys; ++k) {
    gp->codes[k] = codes[k] + offset;
    gp->sums[k] = sums[k];
    prod *= (int64_t)n_groups[k] + 1;
    if (prod > (1ll << 30)) return fail(MRVI_ERR_UNSUPPORTED, "too many group combinations across keys");
  }
  int bits = 1;
  while ((1ll << bits) < prod) ++bits;
  k_composite_codes<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(*gp, h->n_groups_dev, n, h->comp_in, h->idx_in);
  LAUNCH_CHECK();
  size_t tmp = h->cub_tmp_bytes;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(h->cub_tmp, tmp, h->comp_in, h->comp_sorted, h->idx_in, h->order, (int)n, 0, bits, st));
  g_launches.fetch_add(2, std::memory_order_relaxed);  // cub: histogram + onesweep (at least)
  gp->order = h->order;
  gp->comp_sorted = h->comp_sorted;
  gp->n_groups = h->n_groups_dev;
  return 0;
}

static cudaEvent_t next_event(MrviHandle* h) {
  if (h->ev_used == (int)h->ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev.push_back(e);
  }
  return h->ev[h->ev_used++];
}

static int validate_run(MrviHandle* h, int64_t n, const void* X, int64_t ldx, const void* sid, int n_keys,
                        const int32_t* const* codes, const int32_t* n_groups, int norm, double* const* sums) {
  if (!h) return fail(MRVI_ERR_INVALID_ARG, "null handle");
  if (n < 0) return fail(MRVI_ERR_INVALID_ARG, "n_cells < 0");
  if (n > 0 && (!X || !sid)) return fail(MRVI_ERR_INVALID_ARG, "X / sample_idx is null");
  if (ldx < h->dims.n_input) return fail(MRVI_ERR_INVALID_ARG, "ldx (%lld) < n_input (%d)", (long long)ldx, h->dims.n_input);
  if (n_keys < 0 || n_keys > kMaxKeys) return fail(MRVI_ERR_INVALID_ARG, "n_keys must be in [0, %d]", kMaxKeys);
  if (norm < 0 || norm > 2) return fail(MRVI_ERR_INVALID_ARG, "norm %d not supported", norm);
  for (int k = 0; k < n_keys; ++k) {
    if (!codes || !codes[k] || !n_groups || !sums || !sums[k]) return fail(MRVI_ERR_INVALID_ARG, "group key %d: null pointer", k);
    if (n_groups[k] <= 0) return fail(MRVI_ERR_INVALID_ARG, "group key %d: n_groups must be > 0", k);
  }
  return 0;
}

// host buffers only: every sample code must index the per-sample tables (the kernels clamp, the ABI rejects)
static int validate_sample_idx(const MrviHandle* h, const int32_t* sid, int64_t n) {
  const int S = h->dims.n_sample;
  for (int64_t i = 0; i < n; ++i)
    if (sid[i] < 0 || sid[i] >= S)
      return fail(MRVI_ERR_INVALID_ARG, "sample_idx[%lld] = %d outside [0, %d)", (long long)i, sid[i], S);
  return 0;
}

// core: device buffers, one stream, chunked over cells
static int run_device_impl(MrviHandle* h, int64_t n, const float* X, int64_t ldx, const int32_t* sid, int n_keys,
                           const int32_t* const* codes, const int32_t* n_groups, int norm, float* cell_out,
                           double* const* sums, float* z_out, float* u_out, cudaStream_t st, bool time_stages) {
  const NetW& net = h->net;
  if (n_keys > 0)
    CUDA_TRY(cudaMemcpyAsync(h->n_groups_dev, n_groups, n_keys * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  const bool need_dist = cell_out != nullptr || n_keys > 0;
  const bool need_z = need_dist || z_out != nullptr;
  for (int64_t lo = 0; lo < n; lo += h->chunk_cells) {
    const int64_t m = std::min(h->chunk_cells, n - lo);
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
    if (time_stages) { e0 = next_event(h); e1 = next_event(h); e2 = next_event(h); e3 = next_event(h); cudaEventRecord(e0, st); }
    int err;
    float* z = z_out ? z_out + lo * net.S * net.L : h->zbuf;
    if ((err = launch_encoder(h, X + lo * ldx, ldx, sid + lo, m, st))) return err;
    if (time_stages) cudaEventRecord(e1, st);
    if (h->precision == MRVI_PRECISION_TC && need_dist && h->tc.fused_smem_bytes && tc_fused_supported(h->tc, net.S, norm)) {
      // fused: chain + distances + group sums in one kernel; z never leaves the SM unless asked for
      GroupPlan gp;
      if ((err = plan_groups(h, m, n_keys, codes, lo, n_groups, sums, &gp, st))) return err;
      tc_launch_fused(h->tc, h->net, m, h->cb, z_out ? z : nullptr, cell_out ? cell_out + lo * net.S * net.S : nullptr, gp, st);
      LAUNCH_CHECK();
      if (time_stages) { cudaEventRecord(e2, st); cudaEventRecord(e3, st); }
      if (u_out) CUDA_TRY(cudaMemcpyAsync(u_out + lo * net.Lq, h->cb.u, m * net.Lq * sizeof(float), cudaMemcpyDeviceToDevice, st));
      continue;
    }
    if (need_z) {
      if (h->precision == MRVI_PRECISION_TC) {
        if (h->tc.fused_smem_bytes) tc_launch_fused_chain(h->tc, h->net, m, h->cb, z, st);  // two threads per row, 1 CTA / SM
        else if ((err = tc_launch_qz(h->tc, h->net, m, h->cb, z, st))) return err;
        LAUNCH_CHECK();
      } else {
        if ((err = launch_qz(h, h->cb, m, net.S, z, st))) return err;
      }
    }
    if (time_stages) cudaEventRecord(e2, st);
    if (u_out) CUDA_TRY(cudaMemcpyAsync(u_out + lo * net.Lq, h->cb.u, m * net.Lq * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (need_dist) {
      GroupPlan gp;
      if ((err = plan_groups(h, m, n_keys, codes, lo, n_groups, sums, &gp, st))) return err;
      if ((err = launch_dist(h, z, m, norm, cell_out ? cell_out + lo * net.S * net.S : nullptr, gp, st))) return err;
    }
    if (time_stages) cudaEventRecord(e3, st);
  }
  return 0;
}


// ------------------------------------------------------------------------------------------
// sampled paths: see sampled_kernels.cuh
// q(z|u,s) chain (+ distances / group sums) over n (virtual) cells whose rows carry their own features
struct EpsNoise {   // use_map=False on the sampled paths: eps is drawn, not its mean (reference _module.py:326-329)
  const float* noise = nullptr;   // explicit draws [(b * A + a)][noise_n][L] on the device, or null = Philox
  int64_t noise_cell0 = 0, noise_n = 0, gcell0 = 0, n_cells = 0;
  int A = 0, B = 0;
  uint64_t seed = 0;
  uint32_t stream = 2;
};

static int add_eps_noise(MrviHandle* h, const CellBuf& cb, float* z, const EpsNoise& en, cudaStream_t st) {
  const int64_t rows = en.n_cells * en.A * en.B;
  if (rows == 0) return 0;
  k_eps_noise<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(z, cb.raw_out, en.n_cells, en.A, en.B, h->net.L, en.noise, en.noise_cell0,
                                                             en.noise_n, en.gcell0, en.seed, en.stream);
  LAUNCH_CHECK();
  return 0;
}

static int chain_and_dist(MrviHandle* h, const CellBuf& cb, int64_t n, int norm, float* z_out, float* z_scratch,
                          float* cell_out, const GroupPlan& gp, cudaStream_t st, const EpsNoise* en = nullptr) {
  const NetW& net = h->net;
  const bool need_dist = cell_out != nullptr || gp.n_keys > 0;
  int err;
  if (h->precision == MRVI_PRECISION_TC && need_dist && h->tc.fused_smem_bytes && tc_fused_supported(h->tc, net.S, norm)) {
    tc_launch_fused(h->tc, net, n, cb, z_out, cell_out, gp, st);
    LAUNCH_CHECK();
    return 0;
  }
  float* z = z_out ? z_out : z_scratch;
  if (h->precision == MRVI_PRECISION_TC) {
    if (h->tc.fused_smem_bytes) tc_launch_fused_chain(h->tc, net, n, cb, z, st);
    else if ((err = tc_launch_qz(h->tc, net, n, cb, z, st))) return err;
    LAUNCH_CHECK();
  } else {
    if ((err = launch_qz(h, cb, n, net.S, z, st))) return err;
    if (en && (err = add_eps_noise(h, cb, z, *en, st))) return err;
  }
  if (need_dist && (err = launch_dist(h, z, n, norm, cell_out, gp, st))) return err;
  return 0;
}

static int run_sampled_host_impl(MrviHandle* h, int64_t n_cells, const float* X, int64_t ldx, const int32_t* sample_idx,
                                 int mc, uint64_t seed, int64_t cell_offset, const float* u_noise, const float* base_noise,
                                 const float* eps_noise, const float* base_eps_noise,
                                 int n_keys, const int32_t* const* group_codes, const int32_t* n_groups, int norm,
                                 int normalize, float* cell_out, double* const* group_sums, int64_t* const* group_counts,
                                 float* z_out, float* u_out, int64_t chunk_cells) {
  int err;
  if ((err = validate_run(h, n_cells, X, ldx, sample_idx, n_keys, group_codes, n_groups, norm, group_sums))) return err;
  if (mc <= 0 || mc > 4096) return fail(MRVI_ERR_INVALID_ARG, "mc_samples must be in [1, 4096]");
  if ((err = validate_sample_idx(h, sample_idx, n_cells))) return err;
  if (normalize && norm != MRVI_NORM_L2) return fail(MRVI_ERR_INVALID_ARG, "Norm must be 'l2' when using normalized distances");
  const bool sample_eps = !h->dims.use_map;   // eps ~ Normal(loc, softplus(raw) + 1e-3) is drawn as well (reference _module.py:326-329)
  if (sample_eps && h->precision != MRVI_PRECISION_FP32)
    return fail(MRVI_ERR_UNSUPPORTED, "sampled paths with use_map=False (sampled eps) need MRVI_PRECISION_FP32");
  if (h->net.enc_scale.w == nullptr)
    return fail(MRVI_ERR_PARAM, "missing parameter 'qu/NormalDistOutputNN_0/Dense_1/kernel' (scale head of q(u|x), needed by the sampled paths)");
  if (h->net.Lq > kMaxLq) return fail(MRVI_ERR_UNSUPPORTED, "n_latent_u > %d", kMaxLq);
  CUDA_TRY(cudaSetDevice(h->device));
  const NetW& net = h->net;
  const int S = net.S, L = net.L, G = net.G, Lq = net.Lq, E = net.E;
  const int64_t SS = (int64_t)S * S;
  for (int k = 0; k < n_keys; ++k) {
    for (int64_t i = 0; i < n_cells; ++i) {
      const int g = group_codes[k][i];
      if (g < -1 || g >= n_groups[k])
        return fail(MRVI_ERR_INVALID_ARG, "group key %d: code %d of cell %lld outside [-1, %d)", k, g, (long long)i, n_groups[k]);
    }
    if ((int64_t)n_groups[k] * (normalize ? 1 : mc) > (1 << 24)) return fail(MRVI_ERR_UNSUPPORTED, "n_groups * mc_samples too large");
    if (group_counts && group_counts[k]) {
      for (int g = 0; g < n_groups[k]; ++g) group_counts[k][g] = 0;
      for (int64_t i = 0; i < n_cells; ++i) if (group_codes[k][i] >= 0) group_counts[k][group_codes[k][i]]++;
    }
  }
  const int gmul = normalize ? 1 : mc;  // group sums keep the mc axis unless normalised
  // ---- chunk: virtual cells (cell, draw) must fit the handle's sort / workspace buffers; bound staging to ~1.5 GiB
  const int64_t per_cell = (int64_t)mc * S * (Lq + E + L + L + S) * 4 + (int64_t)G * 4 + (int64_t)2 * mc * (Lq + E + 2 * L) * 4;
  int64_t ch = std::max<int64_t>(1, std::min<int64_t>(h->chunk_cells / mc, (1536ll << 20) / per_cell));
  if (chunk_cells > 0) ch = std::min(ch, chunk_cells);
  ch = std::min<int64_t>(ch, std::max<int64_t>(n_cells, 1));
  if (h->chunk_cells / mc < 1) return fail(MRVI_ERR_UNSUPPORTED, "mc_samples=%d too large for the workspace", mc);
  const int64_t nvmax = ch * mc;

  DeviceArena ws;  // everything of this call; released on return
  cudaStream_t st = nullptr;
  float *dX, *d_scale, *d_mu, *d_var, *d_zb, *d_z, *d_dv = nullptr, *d_cell = nullptr, *d_unoise = nullptr, *d_bnoise = nullptr;
  float *d_enoise = nullptr, *d_benoise = nullptr, *d_udraw = nullptr;
  int32_t *d_sid, *d_codes[kMaxKeys] = {}, *d_vcodes[kMaxKeys] = {};
  double* d_sums[kMaxKeys] = {};
  CellBuf rcb{}, bcb{};
  CUDA_TRY(ws.alloc((void**)&dX, ch * G * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&d_sid, ch * sizeof(int32_t)));
  CUDA_TRY(ws.alloc((void**)&d_scale, ch * Lq * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&rcb.u_ln, nvmax * S * Lq * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&rcb.q_tok, nvmax * S * E * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&rcb.zbase, nvmax * S * L * sizeof(float)));
  rcb.per_row = 1;
  CUDA_TRY(ws.alloc((void**)&d_z, nvmax * S * L * sizeof(float)));
  if (u_out) CUDA_TRY(ws.alloc((void**)&d_udraw, nvmax * S * Lq * sizeof(float)));
  if (sample_eps) {
    CUDA_TRY(ws.alloc((void**)&rcb.raw_out, nvmax * S * L * sizeof(float)));
    if (normalize) CUDA_TRY(ws.alloc((void**)&bcb.raw_out, ch * 2 * mc * L * sizeof(float)));
    if (eps_noise) {
      CUDA_TRY(ws.alloc((void**)&d_enoise, (size_t)S * mc * n_cells * L * sizeof(float)));
      CUDA_TRY(cudaMemcpyAsync(d_enoise, eps_noise, (size_t)S * mc * n_cells * L * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    if (normalize && base_eps_noise) {
      CUDA_TRY(ws.alloc((void**)&d_benoise, (size_t)2 * mc * n_cells * L * sizeof(float)));
      CUDA_TRY(cudaMemcpyAsync(d_benoise, base_eps_noise, (size_t)2 * mc * n_cells * L * sizeof(float), cudaMemcpyHostToDevice, st));
    }
  }
  if (normalize) {
    CUDA_TRY(ws.alloc((void**)&bcb.u_ln, ch * 2 * mc * Lq * sizeof(float)));
    CUDA_TRY(ws.alloc((void**)&bcb.q_tok, ch * 2 * mc * E * sizeof(float)));
    CUDA_TRY(ws.alloc((void**)&bcb.zbase, ch * 2 * mc * L * sizeof(float)));
    bcb.per_row = 1;
    bcb.cell_sample = d_sid;
    CUDA_TRY(ws.alloc((void**)&d_zb, ch * 2 * mc * L * sizeof(float)));
    CUDA_TRY(ws.alloc((void**)&d_mu, ch * sizeof(float)));
    CUDA_TRY(ws.alloc((void**)&d_var, ch * sizeof(float)));
    CUDA_TRY(ws.alloc((void**)&d_dv, nvmax * SS * sizeof(float)));
    if (cell_out) CUDA_TRY(ws.alloc((void**)&d_cell, ch * SS * sizeof(float)));
  } else if (cell_out) {
    CUDA_TRY(ws.alloc((void**)&d_cell, nvmax * SS * sizeof(float)));
  }
  for (int k = 0; k < n_keys; ++k) {
    CUDA_TRY(ws.alloc((void**)&d_codes[k], ch * sizeof(int32_t)));
    if (!normalize) CUDA_TRY(ws.alloc((void**)&d_vcodes[k], nvmax * sizeof(int32_t)));
    CUDA_TRY(ws.alloc((void**)&d_sums[k], (size_t)n_groups[k] * gmul * SS * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(d_sums[k], 0, (size_t)n_groups[k] * gmul * SS * sizeof(double), st));
  }
  if (u_noise) {
    CUDA_TRY(ws.alloc((void**)&d_unoise, (size_t)S * mc * n_cells * Lq * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(d_unoise, u_noise, (size_t)S * mc * n_cells * Lq * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  if (normalize && base_noise) {
    CUDA_TRY(ws.alloc((void**)&d_bnoise, (size_t)2 * mc * n_cells * Lq * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(d_bnoise, base_noise, (size_t)2 * mc * n_cells * Lq * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  int32_t ng_eff[kMaxKeys];
  for (int k = 0; k < n_keys; ++k) ng_eff[k] = n_groups[k] * gmul;
  if (n_keys > 0) CUDA_TRY(cudaMemcpyAsync(h->n_groups_dev, ng_eff, n_keys * sizeof(int32_t), cudaMemcpyHostToDevice, st));

  h->ev_used = 0;
  CUDA_TRY(cudaEventRecord(h->ev_call[0], st));
  for (int64_t lo = 0; lo < n_cells; lo += ch) {
    const int64_t m = std::min(ch, n_cells - lo), nv = m * mc;
    CUDA_TRY(cudaMemcpy2DAsync(dX, (size_t)G * sizeof(float), X + lo * ldx, (size_t)ldx * sizeof(float), (size_t)G * sizeof(float),
                               (size_t)m, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_sid, sample_idx + lo, m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    for (int k = 0; k < n_keys; ++k)
      CUDA_TRY(cudaMemcpyAsync(d_codes[k], group_codes[k] + lo, m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    // ---- q(u|x): mean and scale (fp32 tail; the G -> H contraction on the tensor cores in TC mode)
    CellBuf ecb = h->cb;
    ecb.u_scale = d_scale;
    {
      const float* h1 = nullptr;
      if (h->precision == MRVI_PRECISION_TC && h->enc_tc.ready) {
        enc_tc_launch(h->enc_tc, dX, G, m, G, h->h1buf, st);
        LAUNCH_CHECK();
        h1 = h->h1buf;
      }
      k_encoder_fp32<<<(unsigned)((m + kTM - 1) / kTM), kThreads, h->enc_smem, st>>>(h->net, dX, G, d_sid, m, ecb, h->enc_ldh,
                                                                                    h->enc_ldhid, (G % 4) == 0, h1);
      LAUNCH_CHECK();
    }
    // ---- draws of u and the per-row chain inputs for the (cell, draw, sample) rows
    {
      const int64_t rows = nv * S;
      k_sample_rows<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(h->net, ecb.u, d_scale, m, mc, S, d_unoise, lo, n_cells,
                                                                   cell_offset + lo, seed, 0u, rcb.u_ln, rcb.q_tok, rcb.zbase, d_udraw);
      LAUNCH_CHECK();
      if (u_out) CUDA_TRY(cudaMemcpyAsync(u_out + lo * mc * S * Lq, d_udraw, rows * Lq * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    float* dz_out = z_out ? d_z : nullptr;
    EpsNoise en, ben;
    en.noise = d_enoise; en.noise_cell0 = lo; en.noise_n = n_cells; en.gcell0 = cell_offset + lo; en.n_cells = m;
    en.A = mc; en.B = S; en.seed = seed; en.stream = 2u;
    ben = en; ben.noise = d_benoise; ben.A = 2 * mc; ben.B = 1; ben.stream = 3u;
    const EpsNoise* enp = sample_eps ? &en : nullptr;
    if (!normalize) {
      GroupPlan gp;
      const int32_t* vptr[kMaxKeys];
      for (int k = 0; k < n_keys; ++k) {
        k_virtual_codes<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(d_codes[k], m, mc, n_groups[k], d_vcodes[k]);
        LAUNCH_CHECK();
        vptr[k] = d_vcodes[k];
      }
      if ((err = plan_groups(h, nv, n_keys, vptr, 0, ng_eff, d_sums, &gp, st))) return err;
      if ((err = chain_and_dist(h, rcb, nv, norm, dz_out, d_z, d_cell, gp, st, enp))) return err;
      if (cell_out)
        CUDA_TRY(cudaMemcpyAsync(cell_out + lo * mc * SS, d_cell, nv * SS * sizeof(float), cudaMemcpyDeviceToHost, st));
    } else {
      // local baseline: 2 mc draws for the cell's own sample (reference _model.py:508-540)
      const int64_t brow = m * 2 * mc;
      k_sample_rows<<<(unsigned)((brow + 127) / 128), 128, 0, st>>>(h->net, ecb.u, d_scale, m, 2 * mc, 1, d_bnoise, lo, n_cells,
                                                                   cell_offset + lo, seed, 1u, bcb.u_ln, bcb.q_tok, bcb.zbase);
      LAUNCH_CHECK();
      if (h->precision == MRVI_PRECISION_TC) {
        if ((err = tc_launch_qz(h->tc, net, m, bcb, d_zb, st, 2 * mc))) return err;
        LAUNCH_CHECK();
      } else {
        if ((err = launch_qz(h, bcb, m, 2 * mc, d_zb, st))) return err;
        if (sample_eps && (err = add_eps_noise(h, bcb, d_zb, ben, st))) return err;
      }
      k_baseline_stats<<<(unsigned)((m + 7) / 8), 256, 0, st>>>(d_zb, m, mc, L, d_mu, d_var);
      LAUNCH_CHECK();
      GroupPlan none;
      memset(&none, 0, sizeof(none));
      if ((err = chain_and_dist(h, rcb, nv, norm, dz_out, d_z, d_dv, none, st, enp))) return err;
      GroupPlan gp;
      const int32_t* cptr[kMaxKeys];
      for (int k = 0; k < n_keys; ++k) cptr[k] = d_codes[k];
      if ((err = plan_groups(h, m, n_keys, cptr, 0, ng_eff, d_sums, &gp, st))) return err;
      const int eb = (int)((SS + 255) / 256);
      int64_t target = std::max<int64_t>(1, (int64_t)h->sm_count * 8 / eb);
      const int cpw = (int)std::max<int64_t>(1, (m + target - 1) / target);
      dim3 grid((unsigned)((m + cpw - 1) / cpw), (unsigned)eb);
      k_normalize_mc<<<grid, 256, 0, st>>>(d_dv, d_mu, d_var, m, mc, (int)SS, d_cell, gp, cpw);
      LAUNCH_CHECK();
      if (cell_out) CUDA_TRY(cudaMemcpyAsync(cell_out + lo * SS, d_cell, m * SS * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (z_out) CUDA_TRY(cudaMemcpyAsync(z_out + lo * mc * S * L, d_z, nv * S * L * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));  // staging buffers are single: finish the chunk before the next upload
  }
  CUDA_TRY(cudaEventRecord(h->ev_call[1], st));
  for (int k = 0; k < n_keys; ++k)
    CUDA_TRY(cudaMemcpyAsync(group_sums[k], d_sums[k], (size_t)n_groups[k] * gmul * SS * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// ------------------------------------------------------------------------------------------
// Differential-expression eps stage (reference _model.py:1160-1218): for every cell, mc draws of u per
// kept counterfactual sample -> the chain's residual eps (z_base zeroed: the chain then returns eps itself, no
// cancellation against z_base) -> per-cell standardisation over the samples, betas = Amat . eps, whitened effect sizes.
static int run_de_host_impl(MrviHandle* h, int64_t n_cells, const float* X, int64_t ldx, const int32_t* sample_idx, int mc,
                            uint64_t seed, int64_t cell_offset, const float* u_noise, const float* eps_noise, int n_kept,
                            const int32_t* kept, int n_cov, const float* Amat, const float* prefactor, int per_cell_design, float* beta_out,
                            float* effect_out, float* eps_out, int64_t chunk_cells) {
  if (!h) return fail(MRVI_ERR_INVALID_ARG, "null handle");
  if (n_cells < 0 || (n_cells > 0 && (!X || !sample_idx))) return fail(MRVI_ERR_INVALID_ARG, "X / sample_idx is null");
  if (ldx < h->dims.n_input) return fail(MRVI_ERR_INVALID_ARG, "ldx (%lld) < n_input (%d)", (long long)ldx, h->dims.n_input);
  if (mc <= 0 || mc > 4096) return fail(MRVI_ERR_INVALID_ARG, "mc_samples must be in [1, 4096]");
  if (n_cov <= 0 || n_cov > kDeMaxCov) return fail(MRVI_ERR_UNSUPPORTED, "n_covariates must be in [1, %d]", kDeMaxCov);
  if (!Amat || !prefactor || !beta_out || !effect_out) return fail(MRVI_ERR_INVALID_ARG, "Amat / prefactor / beta_out / effect_out is null");
  const NetW& net = h->net;
  const int S = net.S, L = net.L, G = net.G, Lq = net.Lq, E = net.E;
  const int Sk = kept ? n_kept : S;
  if (Sk <= 0 || Sk > S) return fail(MRVI_ERR_INVALID_ARG, "n_kept_samples must be in [1, n_sample]");
  for (int j = 0; kept && j < Sk; ++j)
    if (kept[j] < 0 || kept[j] >= S) return fail(MRVI_ERR_INVALID_ARG, "kept_samples[%d] = %d outside [0, %d)", j, kept[j], S);
  int err;
  if ((err = validate_sample_idx(h, sample_idx, n_cells))) return err;
  const bool sample_eps = !h->dims.use_map;
  if (sample_eps && h->precision != MRVI_PRECISION_FP32)
    return fail(MRVI_ERR_UNSUPPORTED, "differential expression with use_map=False (sampled eps) needs MRVI_PRECISION_FP32");
  if (net.enc_scale.w == nullptr)
    return fail(MRVI_ERR_PARAM, "missing parameter 'qu/NormalDistOutputNN_0/Dense_1/kernel' (scale head of q(u|x), needed by the sampled paths)");
  if (Lq > kMaxLq) return fail(MRVI_ERR_UNSUPPORTED, "n_latent_u > %d", kMaxLq);
  const size_t de_smem = de_stats_smem(n_cov, Sk, L);
  if (de_smem > (size_t)kMaxSmemOptin) return fail(MRVI_ERR_UNSUPPORTED, "n_covariates x n_samples too large for the statistics kernel");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaFuncSetAttribute(k_de_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
  // ---- chunk: (cell, draw) virtual cells of the chain; staging bounded to ~6 GiB (of 180)
  const int64_t per_cell = (int64_t)mc * S * (Lq + E + 2 * L) * 4 + (int64_t)G * 4;
  if (h->chunk_cells / mc < 1) return fail(MRVI_ERR_UNSUPPORTED, "mc_samples=%d too large for the workspace", mc);
  int64_t ch = std::max<int64_t>(1, std::min<int64_t>(h->chunk_cells / mc, (6144ll << 20) / per_cell));
  if (chunk_cells > 0) ch = std::min(ch, chunk_cells);
  ch = std::min<int64_t>(ch, std::max<int64_t>(n_cells, 1));
  const int64_t nvmax = ch * mc;

  DeviceArena ws;
  cudaStream_t st = nullptr;
  float *dX, *d_scale, *d_eps, *d_A, *d_P, *d_beta, *d_eff, *d_unoise = nullptr, *d_enoise = nullptr;
  int32_t *d_sid, *d_kept = nullptr;
  CellBuf rcb{};
  const int64_t design_cells = per_cell_design ? ch : 1;
  CUDA_TRY(ws.alloc((void**)&dX, ch * G * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&d_sid, ch * sizeof(int32_t)));
  CUDA_TRY(ws.alloc((void**)&d_scale, ch * Lq * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&rcb.u_ln, nvmax * S * Lq * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&rcb.q_tok, nvmax * S * E * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&rcb.zbase, nvmax * S * L * sizeof(float)));
  rcb.per_row = 1;
  CUDA_TRY(ws.alloc((void**)&d_eps, nvmax * S * L * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&d_A, design_cells * n_cov * Sk * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&d_P, design_cells * n_cov * n_cov * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&d_beta, ch * n_cov * L * sizeof(float)));
  CUDA_TRY(ws.alloc((void**)&d_eff, ch * n_cov * sizeof(float)));
  if (kept) {
    CUDA_TRY(ws.alloc((void**)&d_kept, Sk * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpyAsync(d_kept, kept, Sk * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  if (u_noise) {
    CUDA_TRY(ws.alloc((void**)&d_unoise, (size_t)S * mc * n_cells * Lq * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(d_unoise, u_noise, (size_t)S * mc * n_cells * Lq * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  if (sample_eps) {
    CUDA_TRY(ws.alloc((void**)&rcb.raw_out, nvmax * S * L * sizeof(float)));
    if (eps_noise) {
      CUDA_TRY(ws.alloc((void**)&d_enoise, (size_t)S * mc * n_cells * L * sizeof(float)));
      CUDA_TRY(cudaMemcpyAsync(d_enoise, eps_noise, (size_t)S * mc * n_cells * L * sizeof(float), cudaMemcpyHostToDevice, st));
    }
  }
  if (!per_cell_design) {
    CUDA_TRY(cudaMemcpyAsync(d_A, Amat, (size_t)n_cov * Sk * sizeof(float), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_P, prefactor, (size_t)n_cov * n_cov * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  h->ev_used = 0;
  CUDA_TRY(cudaEventRecord(h->ev_call[0], st));
  for (int64_t lo = 0; lo < n_cells; lo += ch) {
    const int64_t m = std::min(ch, n_cells - lo), nv = m * mc;
    CUDA_TRY(cudaMemcpy2DAsync(dX, (size_t)G * sizeof(float), X + lo * ldx, (size_t)ldx * sizeof(float), (size_t)G * sizeof(float),
                               (size_t)m, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_sid, sample_idx + lo, m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (per_cell_design) {
      CUDA_TRY(cudaMemcpyAsync(d_A, Amat + (size_t)lo * n_cov * Sk, (size_t)m * n_cov * Sk * sizeof(float), cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_P, prefactor + (size_t)lo * n_cov * n_cov, (size_t)m * n_cov * n_cov * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    CellBuf ecb = h->cb;
    ecb.u_scale = d_scale;
    {
      const float* h1 = nullptr;
      if (h->precision == MRVI_PRECISION_TC && h->enc_tc.ready) {
        enc_tc_launch(h->enc_tc, dX, G, m, G, h->h1buf, st);
        LAUNCH_CHECK();
        h1 = h->h1buf;
      }
      k_encoder_fp32<<<(unsigned)((m + kTM - 1) / kTM), kThreads, h->enc_smem, st>>>(h->net, dX, G, d_sid, m, ecb, h->enc_ldh,
                                                                                    h->enc_ldhid, (G % 4) == 0, h1);
      LAUNCH_CHECK();
    }
    {
      const int64_t rows = nv * S;
      k_sample_rows<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(h->net, ecb.u, d_scale, m, mc, S, d_unoise, lo, n_cells,
                                                                   cell_offset + lo, seed, 0u, rcb.u_ln, rcb.q_tok, rcb.zbase);
      LAUNCH_CHECK();
      CUDA_TRY(cudaMemsetAsync(rcb.zbase, 0, (size_t)rows * L * sizeof(float), st));   // chain output = eps
    }
    GroupPlan none;
    memset(&none, 0, sizeof(none));
    EpsNoise en;
    en.noise = d_enoise; en.noise_cell0 = lo; en.noise_n = n_cells; en.gcell0 = cell_offset + lo; en.n_cells = m;
    en.A = mc; en.B = S; en.seed = seed; en.stream = 2u;
    if ((err = chain_and_dist(h, rcb, nv, MRVI_NORM_L2, d_eps, d_eps, nullptr, none, st, sample_eps ? &en : nullptr))) return err;
    k_de_stats<<<(unsigned)m, kDeWarps * 32, de_smem, st>>>(d_eps, m, mc, S, L, d_kept, Sk, n_cov, d_A, d_P, per_cell_design ? 1 : 0,
                                                            d_beta, d_eff);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(beta_out + (size_t)lo * n_cov * L, d_beta, (size_t)m * n_cov * L * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(effect_out + (size_t)lo * n_cov, d_eff, (size_t)m * n_cov * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (eps_out) CUDA_TRY(cudaMemcpyAsync(eps_out + (size_t)lo * mc * S * L, d_eps, (size_t)nv * S * L * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  CUDA_TRY(cudaEventRecord(h->ev_call[1], st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// ------------------------------------------------------------------------------------------
// NCCL, resolved at run time (dlopen): the library has no link-time dependency on it, and inside a torch process the
// already loaded libnccl.so.2 is picked up.
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
static NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return api;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(lib, "ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GroupStart && api.GroupEnd;
  return api;
}
#define NCCL_TRY(expr)                                                                              \
  do {                                                                                              \
    int _r = (expr);                                                                                \
    if (_r != 0)                                                                                    \
      return fail(MRVI_ERR_CUDA, "%s failed: %s", #expr, nccl_api().GetErrorString ? nccl_api().GetErrorString(_r) : "NCCL error"); \
  } while (0)

}  // namespace mrvi

// ==========================================================================================
extern "C" {

int mrvi_abi_version(void) { return MRVI_B200_ABI_VERSION; }

const char* mrvi_last_error(void) { return g_last_error.c_str(); }

int64_t mrvi_kernel_launch_count(void) { return g_launches.load(); }

int mrvi_create(const MrviDims* dims, int32_t n_params, const char* const* names, const float* const* values,
                const int64_t* numels, int32_t device, int32_t precision, MrviHandle** out) {
  if (!dims || !out || (n_params > 0 && (!names || !values || !numels)))
    return fail(MRVI_ERR_INVALID_ARG, "mrvi_create: null argument");
  *out = nullptr;
  const MrviDims& d = *dims;
  if (d.n_input <= 0 || d.n_sample <= 0 || d.n_latent <= 0 || d.n_latent_u < 0 || d.encoder_n_hidden <= 0)
    return fail(MRVI_ERR_INVALID_ARG, "mrvi_create: non-positive dimension");
  if (d.n_latent_sample <= 0 || d.n_latent_sample > kMaxE)
    return fail(MRVI_ERR_UNSUPPORTED, "n_latent_sample=%d outside [1, %d]", d.n_latent_sample, kMaxE);
  if (d.n_channels <= 0 || d.n_channels > kMaxHeadDim || d.n_heads <= 0 || d.n_heads > kMaxHeads)
    return fail(MRVI_ERR_UNSUPPORTED, "n_channels=%d / n_heads=%d outside supported range (<= %d / <= %d)",
                d.n_channels, d.n_heads, kMaxHeadDim, kMaxHeads);
  if (d.encoder_n_layers < 0 || d.encoder_n_layers > kMaxBlocks || d.qz_n_layers < 0 || d.qz_n_layers > kMaxBlocks)
    return fail(MRVI_ERR_UNSUPPORTED, "more than %d ResnetBlocks per MLP", kMaxBlocks);
  if (d.encoder_n_hidden > 512 || d.qz_n_hidden > 128 || d.n_latent > 96)
    return fail(MRVI_ERR_UNSUPPORTED, "encoder_n_hidden<=512, qz_n_hidden<=128, n_latent<=96 required");
  if (d.n_latent_u == d.n_latent)
    return fail(MRVI_ERR_INVALID_ARG, "n_latent_u must be 0 when equal to n_latent (isomorphic)");
  if (precision != MRVI_PRECISION_FP32 && precision != MRVI_PRECISION_TC)
    return fail(MRVI_ERR_INVALID_ARG, "unknown precision %d", precision);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(MRVI_ERR_NO_DEVICE, "no CUDA device visible");
  }
  if (device < 0 || device >= n_dev) return fail(MRVI_ERR_NO_DEVICE, "device %d out of range (%d visible)", device, n_dev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(MRVI_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  CUDA_TRY(cudaSetDevice(device));

  std::unique_ptr<MrviHandle> h(new MrviHandle());
  h->dims = d;
  h->device = device;
  h->precision = precision;
  h->sm_count = prop.multiProcessorCount;
  ParamTable pt;
  for (int i = 0; i < n_params; ++i) {
    if (!names[i] || !values[i]) return fail(MRVI_ERR_INVALID_ARG, "parameter %d: null name or value", i);
    pt.m[names[i]] = {values[i], numels[i]};
  }
  int err;
  if ((err = build_net(h.get(), pt))) return err;
  if ((err = setup_workspace(h.get()))) return err;
  if (precision == MRVI_PRECISION_TC) {
    if ((err = tc_build(h->tc, h->net, h->dims, pt.m, h->arena, h->sm_count))) return fail(err, "%s", tc_error());
    if ((err = enc_tc_build(h->enc_tc, h->net, pt.m, h->arena))) return fail(err, "%s", tc_error());
    if (h->enc_tc.ready) CUDA_TRY(h->arena.alloc((void**)&h->h1buf, (size_t)h->chunk_cells * 128 * sizeof(float)));
    if ((err = tc_fused_prepare(h->tc, h->dims.n_latent))) return fail(err, "%s", tc_error());
    if (dist_tc_supported(h->net.S, h->net.L, MRVI_NORM_L2) && (err = dist_tc_prepare(h->net.L, h->net.S, &h->dist_tc_smem)))
      return fail(err, "%s", tc_error());
    if (h->enc_tc.ready && !getenv("MRVI_B200_FP32_ENCODER_TAIL")) {
      if ((err = enc_tail_build(h->enc_tail, h->net, pt.m, h->arena))) return fail(err, "%s", tc_error());
    }
  }
  CUDA_TRY(cudaEventCreate(&h->ev_call[0]));
  CUDA_TRY(cudaEventCreate(&h->ev_call[1]));
  if (device < 64) g_dev_handles[device].fetch_add(1);
  *out = h.release();
  return 0;
}

int mrvi_destroy(MrviHandle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (h->nccl_comm && nccl_api().ok) nccl_api().CommDestroy(h->nccl_comm);
  if (h->device < 64) g_dev_handles[h->device].fetch_sub(1);
  delete h;
  return 0;
}

int mrvi_run_device(MrviHandle* h, int64_t n_cells, const float* X, int64_t ldx, const int32_t* sample_idx,
                    int32_t n_keys, const int32_t* const* group_codes, const int32_t* n_groups, int32_t norm,
                    float* cell_out, double* const* group_sums, float* z_out, float* u_out, void* stream) {
  int err;
  if ((err = validate_run(h, n_cells, X, ldx, sample_idx, n_keys, group_codes, n_groups, norm, group_sums))) return err;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  h->ev_used = 0;
  CUDA_TRY(cudaEventRecord(h->ev_call[0], st));
  if ((err = run_device_impl(h, n_cells, X, ldx, sample_idx, n_keys, group_codes, n_groups, norm, cell_out,
                             group_sums, z_out, u_out, st, true))) return err;
  CUDA_TRY(cudaEventRecord(h->ev_call[1], st));
  return 0;
}

int mrvi_distances_device(MrviHandle* h, int64_t n_cells, const float* z, int32_t n_keys,
                          const int32_t* const* group_codes, const int32_t* n_groups, int32_t norm,
                          float* cell_out, double* const* group_sums, void* stream) {
  if (!h) return fail(MRVI_ERR_INVALID_ARG, "null handle");
  if (n_cells > 0 && !z) return fail(MRVI_ERR_INVALID_ARG, "z is null");
  int err;
  if ((err = validate_run(h, n_cells, z, h->dims.n_input, z, n_keys, group_codes, n_groups, norm, group_sums))) return err;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const NetW& net = h->net;
  if (n_keys > 0)
    CUDA_TRY(cudaMemcpyAsync(h->n_groups_dev, n_groups, n_keys * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  for (int64_t lo = 0; lo < n_cells; lo += h->chunk_cells) {
    const int64_t m = std::min(h->chunk_cells, n_cells - lo);
    GroupPlan gp;
    if ((err = plan_groups(h, m, n_keys, group_codes, lo, n_groups, group_sums, &gp, st))) return err;
    if ((err = launch_dist(h, z + lo * net.S * net.L, m, norm, cell_out ? cell_out + lo * net.S * net.S : nullptr, gp, st))) return err;
  }
  return 0;
}

// x_kind: element type of a dense X: 0 = float32, 1 / 2 = uint8 / uint16, 3 / 4 = int32 / int64 counts (mrvi_run_host_counts)
static int run_host_impl(MrviHandle* h, int64_t n_cells, const float* X, int x_kind, int64_t ldx, const int64_t* indptr,
                         const int32_t* indices, const float* data, const int32_t* sample_idx,
                         int32_t n_keys, const int32_t* const* group_codes, const int32_t* n_groups, int32_t norm,
                         float* cell_out, double* const* group_sums, int64_t* const* group_counts, float* z_out,
                         float* u_out, int64_t chunk_cells) {
  int err;
  const bool csr = indptr != nullptr;
  if ((err = validate_run(h, n_cells, csr ? (const void*)indptr : (const void*)X, csr ? h->dims.n_input : ldx, sample_idx, n_keys,
                          group_codes, n_groups, norm, group_sums))) return err;
  if (csr && n_cells > 0 && (!indices || !data) && indptr[n_cells] > indptr[0])
    return fail(MRVI_ERR_INVALID_ARG, "CSR indices / data is null");
  if ((err = validate_sample_idx(h, sample_idx, n_cells))) return err;
  CUDA_TRY(cudaSetDevice(h->device));
  const NetW& net = h->net;
  const int S = net.S, L = net.L, G = net.G, Lq = net.Lq;
  // ---- group counts: host integer bookkeeping
  for (int k = 0; k < n_keys; ++k) {
    for (int64_t i = 0; i < n_cells; ++i) {
      const int g = group_codes[k][i];
      if (g < -1 || g >= n_groups[k])
        return fail(MRVI_ERR_INVALID_ARG, "group key %d: code %d of cell %lld outside [-1, %d)", k, g, (long long)i, n_groups[k]);
    }
    if (group_counts && group_counts[k]) {
      for (int g = 0; g < n_groups[k]; ++g) group_counts[k][g] = 0;
      for (int64_t i = 0; i < n_cells; ++i) if (group_codes[k][i] >= 0) group_counts[k][group_codes[k][i]]++;
    }
  }
  // ---- chunk size: bound device staging (2 buffers) to ~1 GiB per buffer set
  int64_t per_cell = (int64_t)G * 4 + 4 + 4 * n_keys + (cell_out ? (int64_t)S * S * 4 : 0) + (z_out ? (int64_t)S * L * 4 : 0) + (u_out ? Lq * 4 : 0);
  int64_t ch = chunk_cells > 0 ? chunk_cells : std::max<int64_t>(256, std::min<int64_t>(16384, (768ll << 20) / per_cell));
  ch = std::min<int64_t>(ch, std::max<int64_t>(n_cells, 1));
  HostStage& hs = h->hs;
  if (!hs.s_h2d) {
    CUDA_TRY(cudaStreamCreateWithFlags(&hs.s_h2d, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&hs.s_comp, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&hs.s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(cudaEventCreateWithFlags(&hs.ev_h2d[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&hs.ev_comp[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&hs.ev_d2h[i], cudaEventDisableTiming));
    }
  }
  // lossless narrowing of count chunks before the copy (host_pack.cpp): dense input, big enough to matter
  static const bool pack_disabled = getenv("MRVI_B200_NO_HOST_PACK") != nullptr;
  static const double pack_min_gbps = getenv("MRVI_B200_PACK_MIN_GBPS") ? atof(getenv("MRVI_B200_PACK_MIN_GBPS")) : 45.0;
  const int x_elem = x_kind == 1 ? 1 : x_kind == 2 ? 2 : x_kind == 4 ? 8 : 4;  // bytes per element of the caller's matrix
  const bool narrow_in = !csr && (x_kind == 1 || x_kind == 2);  // already uint8 / uint16: copied as it is, widened on the GPU
  const bool int_in = !csr && (x_kind == 3 || x_kind == 4);     // int32 / int64 counts: always narrowed (or converted) on the host
  // GPUs fed from this host: ranks of a torchrun job (one process per GPU) or handles of one process (MrVI(devices=...))
  int sharers = std::max(1, live_devices());
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) sharers = std::max(sharers, atoi(e));
  // float32 input is narrowed with one or two GPUs per host; with four or more the host cores per GPU are too few to matter
  // and the probing costs more than it finds (measured at 4 GPUs, C4 shards: 22.3 M cells/s all raw, 19.0 M with narrowing
  // attempts that fall back to raw)
  const bool use_pack = !csr && !narrow_in && (int_in || (!pack_disabled && sharers <= 2 && n_cells * (int64_t)G >= (1ll << 22)));
  const size_t pack_elem = int_in ? 4 : 2;   // staging bytes per value: the int path may have to fall back to float32
  if (use_pack && !hs.pool) {
    int nt = (int)std::thread::hardware_concurrency() / sharers;   // they share the host cores
    if (const char* e = getenv("MRVI_B200_PACK_THREADS")) nt = atoi(e);
    hs.pool.reset(new PackPool(std::max(1, std::min(nt, 32))));
  }
  // ... and the host memory bus / PCIe switches: with N GPUs pulling raw float32 at once each gets ~45 GB/s * 2 / N
  // (measured 180 GB/s aggregate at N = 8, SCALE_r01), so narrowing keeps paying at a proportionally lower packing rate
  const double pack_floor = pack_min_gbps * std::min(1.0, 2.0 / sharers);
  // (pipeline: the narrowing of chunk c+1 overlaps the copy and the kernels of chunk c)
  bool sums_ok = true;
  for (int k = 0; k < n_keys; ++k) sums_ok &= hs.sums_elems[k] >= (size_t)n_groups[k] * S * S;
  int64_t max_nnz = 0;
  if (csr)
    for (int64_t lo = 0; lo < n_cells; lo += ch) max_nnz = std::max(max_nnz, indptr[std::min(n_cells, lo + ch)] - indptr[lo]);
  if (hs.chunk < ch || hs.n_keys < n_keys || (cell_out && !hs.has_cell) || (z_out && !hs.has_z) || (u_out && !hs.has_u) || !sums_ok ||
      (csr && hs.csr_nnz_cap < max_nnz) || (use_pack && !hs.pack_unavailable && hs.pack_bytes < (size_t)ch * G * pack_elem) ||
      (narrow_in && hs.dpack_bytes < (size_t)ch * G * 2)) {
    CUDA_TRY(cudaDeviceSynchronize());
    hs.free_buffers();
    auto A = [&](void** p, size_t bytes) -> cudaError_t {
      cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
      if (e == cudaSuccess) hs.owned.push_back(*p);
      return e;
    };
    for (int b = 0; b < 2; ++b) {
      CUDA_TRY(A((void**)&hs.X[b], ch * G * sizeof(float)));
      if ((use_pack && !hs.pack_unavailable) || narrow_in) {
        CUDA_TRY(A((void**)&hs.dpack[b], (size_t)ch * G * 2));
        hs.dpack_bytes = (size_t)ch * G * 2;
      }
      if (use_pack && !hs.pack_unavailable) {
        if (cudaHostAlloc((void**)&hs.hpack[b], (size_t)ch * G * pack_elem, cudaHostAllocDefault) != cudaSuccess) {
          cudaGetLastError();          // no pinned memory to spare: the narrowing is an optimisation, not a requirement
          hs.hpack[b] = nullptr;
          hs.pack_unavailable = true;
        } else {
          hs.pack_bytes = (size_t)ch * G * pack_elem;
        }
      }
      hs.csr_ptr[b] = nullptr; hs.csr_idx[b] = nullptr; hs.csr_val[b] = nullptr;
      if (csr) {
        CUDA_TRY(A((void**)&hs.csr_ptr[b], (ch + 1) * sizeof(int64_t)));
        CUDA_TRY(A((void**)&hs.csr_idx[b], std::max<int64_t>(max_nnz, 1) * sizeof(int32_t)));
        CUDA_TRY(A((void**)&hs.csr_val[b], std::max<int64_t>(max_nnz, 1) * sizeof(float)));
      }
      CUDA_TRY(A((void**)&hs.sid[b], ch * sizeof(int32_t)));
      for (int k = 0; k < n_keys; ++k) CUDA_TRY(A((void**)&hs.codes[b][k], ch * sizeof(int32_t)));
      hs.cell[b] = nullptr; hs.z[b] = nullptr; hs.u[b] = nullptr;
      if (cell_out) CUDA_TRY(A((void**)&hs.cell[b], ch * S * S * sizeof(float)));
      if (z_out) CUDA_TRY(A((void**)&hs.z[b], ch * S * L * sizeof(float)));
      if (u_out) CUDA_TRY(A((void**)&hs.u[b], ch * Lq * sizeof(float)));
    }
    for (int k = 0; k < kMaxKeys; ++k) { hs.sums[k] = nullptr; hs.sums_elems[k] = 0; }
    for (int k = 0; k < n_keys; ++k) {
      hs.sums_elems[k] = (size_t)n_groups[k] * S * S;
      CUDA_TRY(A((void**)&hs.sums[k], hs.sums_elems[k] * sizeof(double)));
    }
    hs.csr_nnz_cap = csr ? max_nnz : 0;
    hs.chunk = ch; hs.n_keys = n_keys; hs.has_cell = cell_out != nullptr; hs.has_z = z_out != nullptr; hs.has_u = u_out != nullptr;
  }
  ch = std::min(ch, hs.chunk);
  if (int_in && (hs.pack_unavailable || !hs.hpack[0] || !hs.hpack[1]))
    return fail(MRVI_ERR_CUDA, "no pinned host memory for the staging of an int32 / int64 count matrix");
  for (int k = 0; k < n_keys; ++k)
    CUDA_TRY(cudaMemsetAsync(hs.sums[k], 0, (size_t)n_groups[k] * S * S * sizeof(double), hs.s_comp));
  h->ev_used = 0;
  h->last_h2d_bytes = 0;
  int pack_mode[2] = {0, 0};
  int64_t pack_rows[2] = {0, 0};   // leading rows of the chunk that travel narrowed (the rest, if any, is copied raw)
  bool x_pinned = false;           // the caller's matrix is page-locked: raw rows can be pulled by the copy engine concurrently
  if (!csr && x_kind == 0 && X != nullptr && !getenv("MRVI_B200_NO_SPLIT_H2D")) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, X) == cudaSuccess) x_pinned = pa.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  const double pcie_gbps = getenv("MRVI_B200_PCIE_GBPS") ? std::max(1.0, atof(getenv("MRVI_B200_PCIE_GBPS"))) : 47.0;  // raw float32 rate of the link
  CUDA_TRY(cudaEventRecord(h->ev_call[0], hs.s_comp));
  int64_t c = 0;
  // (chunk size, measured on the C4 shard with the narrowing at 120 GB/s: 8192-cell chunks 12.5 ms, 16384 11.1 ms, 32768 12.1 ms:
  // a chunk's launches and the encoder's partial wave against the un-overlapped first narrowing / last kernels)
  for (int64_t lo = 0; lo < n_cells; lo += ch, ++c) {
    const int b = (int)(c & 1);
    const int64_t m = std::min(ch, n_cells - lo);
    // H2D (wait until the kernels of chunk c-2 released input buffers b)
    if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(hs.s_h2d, hs.ev_comp[b], 0));
    pack_rows[b] = m;
    if (csr) {
      const int64_t nz0 = indptr[lo], nnz = indptr[lo + m] - nz0;
      CUDA_TRY(cudaMemcpyAsync(hs.csr_ptr[b], indptr + lo, (m + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, hs.s_h2d));
      h->last_h2d_bytes += (m + 1) * 8 + nnz * 8;
      if (nnz > 0) {
        CUDA_TRY(cudaMemcpyAsync(hs.csr_idx[b], indices + nz0, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, hs.s_h2d));
        CUDA_TRY(cudaMemcpyAsync(hs.csr_val[b], data + nz0, nnz * sizeof(float), cudaMemcpyHostToDevice, hs.s_h2d));
      }
    } else if (narrow_in) {
      const uint8_t* Xb = reinterpret_cast<const uint8_t*>(X);
      if (ldx == G)  // contiguous rows: one linear copy (a pitched copy of 2 KB rows is descriptor-bound)
        CUDA_TRY(cudaMemcpyAsync(hs.dpack[b], Xb + (size_t)lo * G * x_elem, (size_t)m * G * x_elem, cudaMemcpyHostToDevice, hs.s_h2d));
      else
        CUDA_TRY(cudaMemcpy2DAsync(hs.dpack[b], (size_t)G * x_elem, Xb + (size_t)lo * ldx * x_elem, (size_t)ldx * x_elem,
                                   (size_t)G * x_elem, (size_t)m, cudaMemcpyHostToDevice, hs.s_h2d));
      pack_mode[b] = x_elem;
      h->last_h2d_bytes += (int64_t)m * G * x_elem;
    } else if (int_in) {
      if (c >= 2) CUDA_TRY(cudaEventSynchronize(hs.ev_h2d[b]));  // the copy of chunk c-2 has left the pinned staging buffer
      const int mode = narrow_counts_int(*hs.pool, reinterpret_cast<const uint8_t*>(X) + (size_t)lo * ldx * x_elem, x_elem, ldx, m, G, hs.hpack[b]);
      if (mode == 4) {  // negative or >= 65536 somewhere: the chunk travels as the float32 values the loader would hand
        CUDA_TRY(cudaMemcpyAsync(hs.X[b], hs.hpack[b], (size_t)m * G * sizeof(float), cudaMemcpyHostToDevice, hs.s_h2d));
        pack_mode[b] = 0;
        h->last_h2d_bytes += (int64_t)m * G * 4;
      } else {
        CUDA_TRY(cudaMemcpyAsync(hs.dpack[b], hs.hpack[b], (size_t)m * G * mode, cudaMemcpyHostToDevice, hs.s_h2d));
        pack_mode[b] = mode;
        h->last_h2d_bytes += (int64_t)m * G * mode;
      }
    } else {
      int narrow = 0;
      // Pageable input: the narrowing pays only if the host reads X faster than PCIe would carry it raw (~47 GB/s measured);
      // it is kept while the measured rate stays above that floor and re-probed now and then (the host may have been busy).
      // Page-locked input needs no floor: the split below gives the host threads the share of rows they can keep up with --
      // as long as the host memory bus has room for the extra pass (one or two GPUs per host; with more, the raw copies alone
      // saturate it -- 180 GB/s aggregate at 8 GPUs -- and narrowing would add traffic, so the floor stays).
      const double floor_gbps = (x_pinned && sharers <= 2) ? 0.08 * pack_floor : pack_floor;
      if (use_pack && hs.pack_rate_gbps != 0.0 && hs.pack_rate_gbps < floor_gbps && ++hs.pack_skipped >= 64) {
        hs.pack_rate_gbps = 0.0;
        hs.pack_skipped = 0;
      }
      const bool want_pack = use_pack && !hs.pack_unavailable && hs.hpack[b] != nullptr &&
                             (hs.pack_rate_gbps == 0.0 || hs.pack_rate_gbps >= floor_gbps);
      auto copy_raw = [&](int64_t r0, int64_t rows) -> cudaError_t {   // rows [r0, r0 + rows) of the chunk as float32
        if (rows <= 0) return cudaSuccess;
        h->last_h2d_bytes += rows * G * 4;
        if (ldx == G)
          return cudaMemcpyAsync(hs.X[b] + r0 * G, X + (lo + r0) * ldx, (size_t)rows * G * sizeof(float), cudaMemcpyHostToDevice, hs.s_h2d);
        return cudaMemcpy2DAsync(hs.X[b] + r0 * G, (size_t)G * sizeof(float), X + (lo + r0) * ldx, (size_t)ldx * sizeof(float),
                                 (size_t)G * sizeof(float), (size_t)rows, cudaMemcpyHostToDevice, hs.s_h2d);
      };
      // Split of the chunk between the two ways across: the host threads narrow the leading m_pack rows (they read float32 at
      // C GB/s, the link then carries 1/4 of it) while the copy engine pulls the remaining rows raw from the caller's
      // page-locked matrix (P GB/s of float32) -- both finish together for m_pack / m = C / (P + 3/4 C).  Neither the host
      // cores nor PCIe alone bounds the transfer any more (C = 61, P = 47: 0.66 of the rows are narrowed and the chunk
      // crosses 1.5x faster than all-narrowed, 2x faster than all-raw).  Pageable input is narrowed as a whole: a raw copy
      // from pageable memory is staged by the driver on the calling thread.
      int64_t m_pack = m;
      if (want_pack && x_pinned && hs.pack_rate_gbps > 0.0) {
        const double C = hs.pack_rate_gbps, P = pcie_gbps * std::min(1.0, 4.0 / sharers);   // (8 GPUs pull 180 GB/s raw in aggregate)
        const double f = C / (P + 0.75 * C);
        if (f < 0.92) m_pack = std::min<int64_t>(m, ((int64_t)(f * (double)m) + 63) / 64 * 64);
      }
      if (want_pack) {
        CUDA_TRY(copy_raw(m_pack, m - m_pack));   // travels while the host narrows the leading rows
        if (c >= 2) CUDA_TRY(cudaEventSynchronize(hs.ev_h2d[b]));  // the copy of chunk c-2 has left the pinned staging buffer
        const auto t0 = std::chrono::steady_clock::now();
        narrow = narrow_counts(*hs.pool, X + lo * ldx, ldx, m_pack, G, hs.hpack[b]);
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double gbps = (double)m_pack * G * 4 / std::max(sec, 1e-9) * 1e-9;
        if (m_pack * (int64_t)G >= (1ll << 21)) hs.pack_rate_gbps = hs.pack_rate_gbps == 0.0 ? gbps : 0.7 * hs.pack_rate_gbps + 0.3 * gbps;
      }
      pack_mode[b] = narrow;
      pack_rows[b] = m_pack;
      if (narrow) {
        CUDA_TRY(cudaMemcpyAsync(hs.dpack[b], hs.hpack[b], (size_t)m_pack * G * narrow, cudaMemcpyHostToDevice, hs.s_h2d));
        h->last_h2d_bytes += (int64_t)m_pack * G * narrow;
      } else {
        CUDA_TRY(copy_raw(0, m_pack));
      }
    }
    CUDA_TRY(cudaMemcpyAsync(hs.sid[b], sample_idx + lo, m * sizeof(int32_t), cudaMemcpyHostToDevice, hs.s_h2d));
    for (int k = 0; k < n_keys; ++k)
      CUDA_TRY(cudaMemcpyAsync(hs.codes[b][k], group_codes[k] + lo, m * sizeof(int32_t), cudaMemcpyHostToDevice, hs.s_h2d));
    CUDA_TRY(cudaEventRecord(hs.ev_h2d[b], hs.s_h2d));
    // compute (wait for inputs, and for the D2H of chunk c-2 to release output buffers b)
    CUDA_TRY(cudaStreamWaitEvent(hs.s_comp, hs.ev_h2d[b], 0));
    if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(hs.s_comp, hs.ev_d2h[b], 0));
    if (!csr && pack_mode[b]) {  // widen the narrowed counts back to float32
      const int64_t ne = pack_rows[b] * (int64_t)G;
      const unsigned blocks = (unsigned)((ne + 16 * 256 - 1) / (16 * 256));
      if (pack_mode[b] == 1) k_unpack_counts<uint8_t><<<blocks, 256, 0, hs.s_comp>>>(hs.dpack[b], hs.X[b], ne);
      else k_unpack_counts<uint16_t><<<blocks, 256, 0, hs.s_comp>>>(reinterpret_cast<const uint16_t*>(hs.dpack[b]), hs.X[b], ne);
      LAUNCH_CHECK();
    }
    if (csr) {  // expand the chunk's non-zeros to dense rows on the GPU
      CUDA_TRY(cudaMemsetAsync(hs.X[b], 0, (size_t)m * G * sizeof(float), hs.s_comp));
      k_csr_densify<<<(unsigned)((m + 7) / 8), 256, 0, hs.s_comp>>>(hs.csr_ptr[b], indptr[lo], hs.csr_idx[b], hs.csr_val[b], m, G, hs.X[b]);
      LAUNCH_CHECK();
    }
    if ((err = run_device_impl(h, m, hs.X[b], G, hs.sid[b], n_keys, hs.codes[b], n_groups, norm,
                               cell_out ? hs.cell[b] : nullptr, hs.sums, z_out ? hs.z[b] : nullptr,
                               u_out ? hs.u[b] : nullptr, hs.s_comp, false))) return err;
    CUDA_TRY(cudaEventRecord(hs.ev_comp[b], hs.s_comp));
    // D2H
    if (cell_out || z_out || u_out) {
      CUDA_TRY(cudaStreamWaitEvent(hs.s_d2h, hs.ev_comp[b], 0));
      if (cell_out) CUDA_TRY(cudaMemcpyAsync(cell_out + lo * S * S, hs.cell[b], m * S * S * sizeof(float), cudaMemcpyDeviceToHost, hs.s_d2h));
      if (z_out) CUDA_TRY(cudaMemcpyAsync(z_out + lo * S * L, hs.z[b], m * S * L * sizeof(float), cudaMemcpyDeviceToHost, hs.s_d2h));
      if (u_out) CUDA_TRY(cudaMemcpyAsync(u_out + lo * Lq, hs.u[b], m * Lq * sizeof(float), cudaMemcpyDeviceToHost, hs.s_d2h));
    }
    CUDA_TRY(cudaEventRecord(hs.ev_d2h[b], hs.s_d2h));
  }
  CUDA_TRY(cudaEventRecord(h->ev_call[1], hs.s_comp));
  for (int k = 0; k < n_keys; ++k)
    CUDA_TRY(cudaMemcpyAsync(group_sums[k], hs.sums[k], (size_t)n_groups[k] * S * S * sizeof(double), cudaMemcpyDeviceToHost, hs.s_comp));
  CUDA_TRY(cudaStreamSynchronize(hs.s_h2d));
  CUDA_TRY(cudaStreamSynchronize(hs.s_comp));
  CUDA_TRY(cudaStreamSynchronize(hs.s_d2h));
  return 0;
}

int mrvi_run_host(MrviHandle* h, int64_t n_cells, const float* X, int64_t ldx, const int32_t* sample_idx,
                  int32_t n_keys, const int32_t* const* group_codes, const int32_t* n_groups, int32_t norm,
                  float* cell_out, double* const* group_sums, int64_t* const* group_counts, float* z_out,
                  float* u_out, int64_t chunk_cells) {
  return run_host_impl(h, n_cells, X, 0, ldx, nullptr, nullptr, nullptr, sample_idx, n_keys, group_codes, n_groups, norm, cell_out,
                       group_sums, group_counts, z_out, u_out, chunk_cells);
}

int mrvi_run_host_counts(MrviHandle* h, int64_t n_cells, const void* X, int32_t x_dtype, int64_t ldx, const int32_t* sample_idx,
                         int32_t n_keys, const int32_t* const* group_codes, const int32_t* n_groups, int32_t norm,
                         float* cell_out, double* const* group_sums, int64_t* const* group_counts, float* z_out,
                         float* u_out, int64_t chunk_cells) {
  if (x_dtype < MRVI_X_UINT8 || x_dtype > MRVI_X_INT64) return fail(MRVI_ERR_INVALID_ARG, "x_dtype must be one of MrviXDtype");
  return run_host_impl(h, n_cells, reinterpret_cast<const float*>(X), x_dtype, ldx, nullptr, nullptr, nullptr,
                       sample_idx, n_keys, group_codes, n_groups, norm, cell_out, group_sums, group_counts, z_out, u_out, chunk_cells);
}

int mrvi_run_host_csr(MrviHandle* h, int64_t n_cells, const int64_t* indptr, const int32_t* indices, const float* data,
                      const int32_t* sample_idx, int32_t n_keys, const int32_t* const* group_codes, const int32_t* n_groups,
                      int32_t norm, float* cell_out, double* const* group_sums, int64_t* const* group_counts, float* z_out,
                      float* u_out, int64_t chunk_cells) {
  if (!indptr) return fail(MRVI_ERR_INVALID_ARG, "indptr is null");
  return run_host_impl(h, n_cells, nullptr, 0, h ? h->dims.n_input : 0, indptr, indices, data, sample_idx, n_keys, group_codes, n_groups,
                       norm, cell_out, group_sums, group_counts, z_out, u_out, chunk_cells);
}

int mrvi_run_sampled_host(MrviHandle* h, int64_t n_cells, const float* X, int64_t ldx, const int32_t* sample_idx,
                          int32_t mc_samples, uint64_t seed, int64_t cell_offset, const float* u_noise, const float* base_noise,
                          const float* eps_noise, const float* base_eps_noise,
                          int32_t n_keys, const int32_t* const* group_codes, const int32_t* n_groups, int32_t norm,
                          int32_t normalize, float* cell_out, double* const* group_sums, int64_t* const* group_counts,
                          float* z_out, float* u_out, int64_t chunk_cells) {
  return run_sampled_host_impl(h, n_cells, X, ldx, sample_idx, mc_samples, seed, cell_offset, u_noise, base_noise, eps_noise,
                               base_eps_noise, n_keys,
                               group_codes, n_groups, norm, normalize, cell_out, group_sums, group_counts, z_out, u_out, chunk_cells);
}

int mrvi_run_de_host(MrviHandle* h, int64_t n_cells, const float* X, int64_t ldx, const int32_t* sample_idx, int32_t mc_samples,
                     uint64_t seed, int64_t cell_offset, const float* u_noise, const float* eps_noise, int32_t n_kept_samples,
                     const int32_t* kept_samples, int32_t n_covariates, const float* Amat, const float* prefactor, int32_t per_cell_design, float* beta_out,
                     float* effect_size_out, float* eps_out, int64_t chunk_cells) {
  return run_de_host_impl(h, n_cells, X, ldx, sample_idx, mc_samples, seed, cell_offset, u_noise, eps_noise, n_kept_samples, kept_samples,
                          n_covariates, Amat, prefactor, per_cell_design, beta_out, effect_size_out, eps_out, chunk_cells);
}

#ifdef MRVI_DBG_FUSED
int mrvi_dbg_fused_read(long long* out, int n) { return (int)cudaMemcpyFromSymbol(out, mrvi::g_dbg_fused, sizeof(long long) * n); }
#endif
#ifdef MRVI_DBG_ENC
int mrvi_dbg_enc_read(long long* out, int n) { return (int)cudaMemcpyFromSymbol(out, mrvi::g_dbg_enc, sizeof(long long) * n); }
#endif
#ifdef MRVI_DBG_DIST
int mrvi_dbg_dist_read(long long* out, int n) {  // debug builds only (tools/dbg)
  return (int)cudaMemcpyFromSymbol(out, mrvi::g_dbg_dist, sizeof(long long) * n);
}
#endif

int mrvi_attention_table_info(MrviHandle* h, int32_t* n_intervals, float* t_min, float* t_max, float* max_rel_err) {
  if (!h) return fail(MRVI_ERR_INVALID_ARG, "null handle");
  const TcNet& tc = h->tc;
  if (n_intervals) *n_intervals = tc.att_tab ? tc.att_n : 0;
  if (t_min) *t_min = tc.att_t0;
  if (t_max) *t_max = tc.att_tab ? tc.att_t0 + (float)tc.att_n / tc.att_scale : tc.att_t0;
  if (max_rel_err) *max_rel_err = tc.att_err;
  return 0;
}

int64_t mrvi_last_h2d_bytes(MrviHandle* h) { return h ? h->last_h2d_bytes : 0; }

int mrvi_narrow_counts(const float* X, int64_t ldx, int64_t n_rows, int32_t n_cols, void* out, int32_t n_threads) {
  if (!X || !out || n_rows < 0 || n_cols <= 0 || ldx < n_cols) return fail(MRVI_ERR_INVALID_ARG, "mrvi_narrow_counts: bad argument");
  PackPool pool(std::max(1, (int)n_threads));
  return narrow_counts(pool, X, ldx, n_rows, n_cols, out);
}

int mrvi_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int mrvi_alloc_pinned(int64_t bytes, void** out) {
  if (!out || bytes < 0) return fail(MRVI_ERR_INVALID_ARG, "mrvi_alloc_pinned: bad argument");
  *out = nullptr;
  CUDA_TRY(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 16), cudaHostAllocPortable));
  return 0;
}

int mrvi_free_pinned(void* p) {
  if (p) CUDA_TRY(cudaFreeHost(p));
  return 0;
}

int mrvi_comm_unique_id(void* out128) {
  if (!out128) return fail(MRVI_ERR_INVALID_ARG, "mrvi_comm_unique_id: null argument");
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(MRVI_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  NcclApi::UniqueId id;
  NCCL_TRY(api.GetUniqueId(&id));
  memcpy(out128, id.internal, 128);
  return 0;
}

int mrvi_comm_create(MrviHandle* h, const void* id128, int32_t n_ranks, int32_t rank) {
  if (!h || !id128 || n_ranks <= 0 || rank < 0 || rank >= n_ranks) return fail(MRVI_ERR_INVALID_ARG, "mrvi_comm_create: bad argument");
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(MRVI_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->nccl_comm) { api.CommDestroy(h->nccl_comm); h->nccl_comm = nullptr; }
  NcclApi::UniqueId id;
  memcpy(id.internal, id128, 128);
  NCCL_TRY(api.CommInitRank(&h->nccl_comm, n_ranks, id, rank));
  return 0;
}

int mrvi_allreduce_groups(MrviHandle* h, void* nccl_comm, int32_t n_keys, double* const* group_sums, const int32_t* n_groups,
                          void* stream) {
  if (!h || n_keys < 0 || n_keys > kMaxKeys || (n_keys > 0 && (!group_sums || !n_groups)))
    return fail(MRVI_ERR_INVALID_ARG, "mrvi_allreduce_groups: bad argument");
  void* comm = nccl_comm ? nccl_comm : h->nccl_comm;
  if (!comm) return fail(MRVI_ERR_INVALID_ARG, "mrvi_allreduce_groups: no communicator (pass one or call mrvi_comm_create)");
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(MRVI_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t SS = (size_t)h->dims.n_sample * h->dims.n_sample;
  NCCL_TRY(api.GroupStart());
  for (int k = 0; k < n_keys; ++k) {
    if (!group_sums[k] || n_groups[k] <= 0) { api.GroupEnd(); return fail(MRVI_ERR_INVALID_ARG, "group key %d: null pointer or n_groups <= 0", k); }
    NCCL_TRY(api.AllReduce(group_sums[k], group_sums[k], (size_t)n_groups[k] * SS, /*ncclDouble*/ 8, /*ncclSum*/ 0, comm,
                           static_cast<cudaStream_t>(stream)));
  }
  NCCL_TRY(api.GroupEnd());
  return 0;
}

int mrvi_last_stage_ms(MrviHandle* h, float out_ms[4]) {
  if (!h || !out_ms) return fail(MRVI_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  out_ms[0] = out_ms[1] = out_ms[2] = out_ms[3] = 0.f;
  for (int i = 0; i + 3 < h->ev_used; i += 4) {
    float a = 0, b = 0, c = 0;
    CUDA_TRY(cudaEventElapsedTime(&a, h->ev[i], h->ev[i + 1]));
    CUDA_TRY(cudaEventElapsedTime(&b, h->ev[i + 1], h->ev[i + 2]));
    CUDA_TRY(cudaEventElapsedTime(&c, h->ev[i + 2], h->ev[i + 3]));
    out_ms[0] += a; out_ms[1] += b; out_ms[2] += c;
  }
  CUDA_TRY(cudaEventElapsedTime(&out_ms[3], h->ev_call[0], h->ev_call[1]));
  return 0;
}

}  // extern "C"
