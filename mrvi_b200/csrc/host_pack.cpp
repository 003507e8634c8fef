// Host-side lossless narrowing of a count matrix before it crosses PCIe (mrvi_run_host, dense float32 X).
//
// The end-to-end rate of the path is bound by the host->device copy of X (8 KB per cell at G = 2000, ~50 GB/s),
// not by the kernels.  scRNA-seq counts are small non-negative integers stored as float32 (the reference's
// AnnDataLoader hands `adata.X` as float32, src/mrvi/_model.py:317-319, 377-384): a chunk whose values are all integers
// in [0, 255] (or [0, 65535]) is sent as uint8 (uint16) and widened back to the identical float32 values on the
// GPU (k_unpack_counts).  Any other chunk (non-integers, negatives, larger values) is sent as it is.  The narrowing
// runs on a small persistent thread pool and overlaps the copy / kernels of the previous chunk.
#include "host_pack.h"

#if (defined(__x86_64__) || defined(_M_X64)) && !defined(MRVI_PACK_NO_AVX2)
#define MRVI_PACK_AVX2 1
#include <immintrin.h>
#endif

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mrvi {

// ------------------------------------------------------------------------------------------ thread pool
struct PackPool::Impl {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  std::function<void(int, int)> fn;
  uint64_t generation = 0;
  int pending = 0;
  bool stop = false;
  int n = 1;
};

PackPool::PackPool(int n_threads) : impl_(new Impl) {
  impl_->n = std::max(1, n_threads);
  for (int t = 1; t < impl_->n; ++t) {
    impl_->workers.emplace_back([this, t] {
      uint64_t seen = 0;
      for (;;) {
        std::function<void(int, int)> fn;
        {
          std::unique_lock<std::mutex> lk(impl_->mu);
          impl_->cv_go.wait(lk, [&] { return impl_->stop || impl_->generation != seen; });
          if (impl_->stop) return;
          seen = impl_->generation;
          fn = impl_->fn;
        }
        fn(t, impl_->n);
        {
          std::lock_guard<std::mutex> lk(impl_->mu);
          if (--impl_->pending == 0) impl_->cv_done.notify_one();
        }
      }
    });
  }
}

PackPool::~PackPool() {
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    impl_->stop = true;
  }
  impl_->cv_go.notify_all();
  for (auto& w : impl_->workers) w.join();
  delete impl_;
}

int PackPool::size() const { return impl_->n; }

void PackPool::run(const std::function<void(int, int)>& fn) {
  if (impl_->n == 1) { fn(0, 1); return; }
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    impl_->fn = fn;
    impl_->pending = impl_->n - 1;
    ++impl_->generation;
  }
  impl_->cv_go.notify_all();
  fn(0, impl_->n);
  std::unique_lock<std::mutex> lk(impl_->mu);
  impl_->cv_done.wait(lk, [&] { return impl_->pending == 0; });
}

constexpr int kPrefetchBytes = 2048;   // (a prefetch past the end of the matrix is harmless: the instruction never faults)

// ------------------------------------------------------------------------------------------ row kernels
// flags: bit 0 = a value is not a non-negative integer < 2^31 (cannot narrow), bit 1 = a value is >= 256
namespace {

template <typename T>
uint32_t narrow_rows_scalar(const float* X, int64_t ldx, int64_t r0, int64_t r1, int G, T* out) {
  uint32_t bad = 0, big = 0;
  for (int64_t r = r0; r < r1; ++r) {
    const float* x = X + r * ldx;
    T* o = out + r * G;
    for (int k = 0; k < G; ++k) {
      const float v = x[k];
      const int32_t iv = (v >= 0.f && v < 2147483520.f) ? (int32_t)v : -1;
      bad |= (uint32_t)(iv < 0) | (uint32_t)((float)iv != v);
      big |= (uint32_t)iv;
      o[k] = (T)iv;
    }
  }
  const uint32_t limit = sizeof(T) == 1 ? 0xFFFFFF00u : 0xFFFF0000u;
  return (bad ? 1u : 0u) | ((big & limit) ? 2u : 0u);
}

#ifdef MRVI_PACK_AVX2
__attribute__((target("avx2"))) uint32_t narrow_rows_u8_avx2(const float* X, int64_t ldx, int64_t r0, int64_t r1, int G,
                                                             uint8_t* out) {
  __m256i vor = _mm256_setzero_si256();
  __m256 vbad = _mm256_setzero_ps();
  const __m256i fix = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  uint32_t tail_flags = 0;
  for (int64_t r = r0; r < r1; ++r) {
    const float* x = X + r * ldx;
    uint8_t* o = out + r * G;
    const bool stream16 = (reinterpret_cast<uintptr_t>(o) & 15) == 0;
    int k = 0;
    for (; k + 32 <= G; k += 32) {
      // software prefetch 2 KB ahead (two lines per iteration = what the iteration consumes): on the pool's virtualised Xeon
      // hosts the hardware prefetchers leave a streaming core at ~4 GB/s; with it 8 threads read 54-63 GB/s instead of 35
      _mm_prefetch(reinterpret_cast<const char*>(x + k) + kPrefetchBytes, _MM_HINT_T0);
      _mm_prefetch(reinterpret_cast<const char*>(x + k) + kPrefetchBytes + 64, _MM_HINT_T0);
      const __m256 a0 = _mm256_loadu_ps(x + k), a1 = _mm256_loadu_ps(x + k + 8), a2 = _mm256_loadu_ps(x + k + 16),
                   a3 = _mm256_loadu_ps(x + k + 24);
      const __m256i i0 = _mm256_cvttps_epi32(a0), i1 = _mm256_cvttps_epi32(a1), i2 = _mm256_cvttps_epi32(a2),
                    i3 = _mm256_cvttps_epi32(a3);
      vbad = _mm256_or_ps(vbad, _mm256_or_ps(_mm256_or_ps(_mm256_cmp_ps(_mm256_cvtepi32_ps(i0), a0, _CMP_NEQ_UQ),
                                                          _mm256_cmp_ps(_mm256_cvtepi32_ps(i1), a1, _CMP_NEQ_UQ)),
                                             _mm256_or_ps(_mm256_cmp_ps(_mm256_cvtepi32_ps(i2), a2, _CMP_NEQ_UQ),
                                                          _mm256_cmp_ps(_mm256_cvtepi32_ps(i3), a3, _CMP_NEQ_UQ))));
      vor = _mm256_or_si256(vor, _mm256_or_si256(_mm256_or_si256(i0, i1), _mm256_or_si256(i2, i3)));
      const __m256i p01 = _mm256_packus_epi32(i0, i1), p23 = _mm256_packus_epi32(i2, i3);   // 16-bit, lane interleaved
      const __m256i p = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(p01, p23), fix);      // 8-bit, in order
      // the narrowed chunk is only read back by the DMA engine: non-temporal stores skip the read-for-ownership of
      // the staging buffer (a fifth of the packer's memory traffic)
      if (stream16) {
        _mm_stream_si128(reinterpret_cast<__m128i*>(o + k), _mm256_castsi256_si128(p));
        _mm_stream_si128(reinterpret_cast<__m128i*>(o + k + 16), _mm256_extracti128_si256(p, 1));
      } else {
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(o + k), p);
      }
    }
    if (k < G) {
      for (; k < G; ++k) {
        const float v = x[k];
        const int32_t iv = (v >= 0.f && v < 2147483520.f) ? (int32_t)v : -1;
        tail_flags |= ((iv < 0) || ((float)iv != v)) ? 1u : 0u;
        tail_flags |= ((uint32_t)iv & 0xFFFFFF00u) ? 2u : 0u;
        o[k] = (uint8_t)iv;
      }
    }
  }
  _mm_sfence();
  alignas(32) uint32_t ors[8];
  _mm256_store_si256(reinterpret_cast<__m256i*>(ors), vor);
  uint32_t o = 0;
  for (int j = 0; j < 8; ++j) o |= ors[j];
  uint32_t flags = tail_flags;
  if (_mm256_movemask_ps(vbad) != 0 || (o & 0x80000000u)) flags |= 1u;  // sign bit: negative (or NaN -> 0x80000000)
  if (o & 0x7FFFFF00u) flags |= 2u;
  return flags;
}

bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}
#else   // aarch64 hosts (Grace): the scalar path, which the compiler vectorises for the target
bool have_avx2() { return false; }
uint32_t narrow_rows_u8_avx2(const float* X, int64_t ldx, int64_t r0, int64_t r1, int G, uint8_t* out) {
  return narrow_rows_scalar<uint8_t>(X, ldx, r0, r1, G, out);
}
#endif

}  // namespace

int narrow_counts(PackPool& pool, const float* X, int64_t ldx, int64_t n_rows, int G, void* out) {
  if (n_rows == 0) return 1;
  std::vector<uint32_t> flags(pool.size(), 0u);
  auto span = [&](int t, int nt, int64_t* r0, int64_t* r1) {
    const int64_t per = (n_rows + nt - 1) / nt;
    *r0 = std::min<int64_t>(n_rows, (int64_t)t * per);
    *r1 = std::min<int64_t>(n_rows, *r0 + per);
  };
  pool.run([&](int t, int nt) {
    int64_t r0, r1;
    span(t, nt, &r0, &r1);
    flags[t] = have_avx2() ? narrow_rows_u8_avx2(X, ldx, r0, r1, G, static_cast<uint8_t*>(out))
                           : narrow_rows_scalar<uint8_t>(X, ldx, r0, r1, G, static_cast<uint8_t*>(out));
  });
  uint32_t f = 0;
  for (uint32_t v : flags) f |= v;
  if (f & 1u) return 0;        // not narrowable
  if (!(f & 2u)) return 1;     // uint8
  // some value >= 256: second pass as uint16 (rare for count data; exactness re-checked)
  std::fill(flags.begin(), flags.end(), 0u);
  pool.run([&](int t, int nt) {
    int64_t r0, r1;
    span(t, nt, &r0, &r1);
    flags[t] = narrow_rows_scalar<uint16_t>(X, ldx, r0, r1, G, static_cast<uint16_t*>(out));
  });
  f = 0;
  for (uint32_t v : flags) f |= v;
  return (f & 3u) ? 0 : 2;
}


namespace {
template <typename SRC, typename T>
uint32_t narrow_rows_int(const SRC* X, int64_t ldx, int64_t r0, int64_t r1, int G, T* out) {
  uint64_t ored = 0;
  for (int64_t r = r0; r < r1; ++r) {
    const SRC* x = X + r * ldx;
    T* o = out + r * G;
    for (int k0 = 0; k0 < G; k0 += 64) {
      __builtin_prefetch(reinterpret_cast<const char*>(x + k0) + kPrefetchBytes);
      __builtin_prefetch(reinterpret_cast<const char*>(x + k0) + kPrefetchBytes + 64);
      if (sizeof(SRC) == 8) {
        __builtin_prefetch(reinterpret_cast<const char*>(x + k0) + kPrefetchBytes + 128);
        __builtin_prefetch(reinterpret_cast<const char*>(x + k0) + kPrefetchBytes + 192);
      }
      const int k1 = std::min(G, k0 + 64);
      for (int k = k0; k < k1; ++k) {
        const SRC v = x[k];
        ored |= (uint64_t)(int64_t)v;   // a negative value sets the high bits
        o[k] = (T)v;
      }
    }
  }
  const uint64_t limit = sizeof(T) == 1 ? ~uint64_t(0xFF) : ~uint64_t(0xFFFF);
  return (ored & limit) ? 2u : 0u;
}
template <typename SRC>
void rows_to_f32(const SRC* X, int64_t ldx, int64_t r0, int64_t r1, int G, float* out) {
  for (int64_t r = r0; r < r1; ++r) {
    const SRC* x = X + r * ldx;
    float* o = out + r * G;
    for (int k = 0; k < G; ++k) o[k] = (float)x[k];
  }
}
template <typename SRC>
int narrow_counts_int_t(PackPool& pool, const SRC* X, int64_t ldx, int64_t n_rows, int G, void* out) {
  if (n_rows == 0) return 1;
  std::vector<uint32_t> flags(pool.size(), 0u);
  auto span = [&](int t, int nt, int64_t* r0, int64_t* r1) {
    const int64_t per = (n_rows + nt - 1) / nt;
    *r0 = std::min<int64_t>(n_rows, (int64_t)t * per);
    *r1 = std::min<int64_t>(n_rows, *r0 + per);
  };
  auto any = [&]() { uint32_t f = 0; for (uint32_t v : flags) f |= v; return f != 0; };
  pool.run([&](int t, int nt) { int64_t r0, r1; span(t, nt, &r0, &r1); flags[t] = narrow_rows_int<SRC, uint8_t>(X, ldx, r0, r1, G, static_cast<uint8_t*>(out)); });
  if (!any()) return 1;
  pool.run([&](int t, int nt) { int64_t r0, r1; span(t, nt, &r0, &r1); flags[t] = narrow_rows_int<SRC, uint16_t>(X, ldx, r0, r1, G, static_cast<uint16_t*>(out)); });
  if (!any()) return 2;
  pool.run([&](int t, int nt) { int64_t r0, r1; span(t, nt, &r0, &r1); rows_to_f32<SRC>(X, ldx, r0, r1, G, static_cast<float*>(out)); });
  return 4;
}
}  // namespace

int narrow_counts_int(PackPool& pool, const void* X, int elem_bytes, int64_t ldx, int64_t n_rows, int G, void* out) {
  return elem_bytes == 8 ? narrow_counts_int_t<int64_t>(pool, static_cast<const int64_t*>(X), ldx, n_rows, G, out)
                         : narrow_counts_int_t<int32_t>(pool, static_cast<const int32_t*>(X), ldx, n_rows, G, out);
}

}  // namespace mrvi
