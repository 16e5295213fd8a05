// api.cu -- ABI version and error strings of libpointdae_b200.so.
#include "common.cuh"

extern "C" int pdae_abi_version(void) { return PDAE_ABI_VERSION; }

extern "C" const char *pdae_strerror(int code) {
  if (code == 0) return "success";
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  switch (code) {
    case PDAE_E_INVALID: return "invalid argument (negative size, null pointer, or k out of range)";
    case PDAE_E_UNSUPPORTED: return "shape not supported by the sm_100a kernels (grid or dimension limit)";
    case PDAE_E_WORKSPACE: return "workspace missing or too small (see pdae_*_workspace_bytes)";
    default: return "unknown pointdae_b200 error";
  }
}
