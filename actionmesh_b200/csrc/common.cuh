// Host-side helpers shared by all translation units: error reporting across the C ABI and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace amb {

// error codes returned across the C ABI (0 = ok)
enum : int { AMB_OK = 0, AMB_ERR_ARG = -1, AMB_ERR_CUDA = -2, AMB_ERR_DRIVER = -3, AMB_ERR_UNSUPPORTED = -4 };

void set_last_error(const char* fmt, ...);  // defined in abi.cu (thread-local buffer)

#define AMB_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      amb::set_last_error(__VA_ARGS__); \
      return amb::AMB_ERR_ARG;          \
    }                                   \
  } while (0)

#define AMB_CHECK_CUDA(expr)                                                                     \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      amb::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return amb::AMB_ERR_CUDA;                                                                  \
    }                                                                                            \
  } while (0)

// Encode a bf16 tiled tensor map with 128-byte swizzle.  dims/strides innermost first; strides in BYTES for dims>=1.
// Resolved through cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda (loads on CPU boxes).
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box);

int num_sms();
int ensure_smem_optin_impl(const void* kern, int bytes);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: remembered per (kernel, device).
template <typename Kern>
inline int ensure_smem_optin(Kern kern, int bytes) {
  return ensure_smem_optin_impl(reinterpret_cast<const void*>(kern), bytes);
}

}  // namespace amb
