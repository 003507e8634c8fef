// Host-side lossless narrowing of count matrices before the host->device copy (see host_pack.cpp).
#pragma once
#include <stdint.h>

#include <functional>

namespace mrvi {

class PackPool {  // small persistent thread pool; run() executes fn(thread, n_threads) on every thread and joins
 public:
  explicit PackPool(int n_threads);
  ~PackPool();
  PackPool(const PackPool&) = delete;
  PackPool& operator=(const PackPool&) = delete;
  int size() const;
  void run(const std::function<void(int, int)>& fn);

 private:
  struct Impl;
  Impl* impl_;
};

// Narrow rows [0, n_rows) of X (row stride ldx floats, G columns) into `out` (dense [n_rows][G]).
// Returns 1: written as uint8, 2: written as uint16, 0: not narrowable (out is garbage; send the float32 rows).
int narrow_counts(PackPool& pool, const float* X, int64_t ldx, int64_t n_rows, int G, void* out);

// The same for a count matrix stored as int32 (elem_bytes = 4) or int64 (elem_bytes = 8) -- raw counts are often kept
// that way in adata.X / a layer.  Returns 1 / 2 as above; when a value is negative or >= 65536 the rows are written to
// `out` as float32 instead (the exact conversion scvi's loader would do) and 4 is returned.  `out`: 4 bytes per value.
int narrow_counts_int(PackPool& pool, const void* X, int elem_bytes, int64_t ldx, int64_t n_rows, int G, void* out);

}  // namespace mrvi
