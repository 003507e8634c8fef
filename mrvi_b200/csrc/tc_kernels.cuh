// tcgen05 tensor-core path (MRVI_PRECISION_TC).  Placeholder until the kernels land: creating a
// handle with MRVI_PRECISION_TC fails loudly with MRVI_ERR_UNSUPPORTED (no silent fallback).
#pragma once
#include <map>
#include <string>

#include "common.cuh"

namespace mrvi {

struct DeviceArena;

struct TcNet {
  const float* kvtok = nullptr;  // [S][E] per-sample kv tokens (filled by build_net)
};

inline const char* tc_error() { return "MRVI_PRECISION_TC is not built into this library yet"; }

inline int tc_build(TcNet&, const NetW&, const MrviDims&,
                    const std::map<std::string, std::pair<const float*, int64_t>>&, DeviceArena&, int) {
  return MRVI_ERR_UNSUPPORTED;
}

inline int tc_launch_chain(const TcNet&, const NetW&, const float*, int64_t, const int32_t*, int64_t,
                           const CellBuf&, float*, cudaStream_t, cudaEvent_t) {
  return MRVI_ERR_UNSUPPORTED;
}

}  // namespace mrvi
