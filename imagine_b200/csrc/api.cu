// Library life cycle, error reporting and the scratch workspace of libimagine_b200.
#include <cstdio>
#include <cstring>
#include "common.cuh"

namespace imb {

State& state() {
  static State s;
  return s;
}

int set_cuda_error(cudaError_t e, const char* where) {
  State& s = state();
  snprintf(s.cuda_err, sizeof(s.cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e),
           cudaGetErrorString(e));
  cudaGetLastError();  // clear the sticky-less error state
  return e == cudaErrorMemoryAllocation ? IMB_ENOMEM : IMB_ECUDA;
}

int check_init() { return state().initialised ? IMB_OK : IMB_ENOTINIT; }

int ensure_workspace(size_t bytes) {
  State& s = state();
  if (bytes <= s.ws_bytes) return IMB_OK;
  // growing: wait for users of the old block, then replace it
  IMB_CUDA(cudaDeviceSynchronize());
  if (s.ws) {
    IMB_CUDA(cudaFree(s.ws));
    s.ws = nullptr;
    s.ws_bytes = 0;
  }
  size_t want = align_up(bytes + bytes / 4, (size_t)1 << 20);
  cudaError_t e = cudaMalloc(&s.ws, want);
  if (e != cudaSuccess) {
    want = align_up(bytes, (size_t)1 << 20);
    e = cudaMalloc(&s.ws, want);
  }
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc(workspace)");
  s.ws_bytes = want;
  return IMB_OK;
}

int ensure_buffer(void** buf, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return IMB_OK;
  IMB_CUDA(cudaDeviceSynchronize());
  if (*buf) {
    IMB_CUDA(cudaFree(*buf));
    *buf = nullptr;
    *cap = 0;
  }
  const size_t want = align_up(bytes, (size_t)1 << 16);
  cudaError_t e = cudaMalloc(buf, want);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc(buffer)");
  *cap = want;
  return IMB_OK;
}

}  // namespace imb

using imb::state;

extern "C" {

int imb_init(int device) {
  imb::State& s = state();
  if (s.initialised && s.device == device) return IMB_OK;
  if (s.initialised) imb_finalize();
  IMB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  IMB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    snprintf(s.cuda_err, sizeof(s.cuda_err),
             "device %d is sm_%d%d; libimagine_b200 is built for sm_100a only", device, prop.major,
             prop.minor);
    return IMB_EUNSUPPORTED;
  }
  s.device = device;
  s.sm_count = prop.multiProcessorCount;
  s.pinned_bytes = 4096;
  IMB_CUDA(cudaMallocHost(&s.pinned, s.pinned_bytes));
  s.launches = 0;
  s.initialised = true;
  return IMB_OK;
}

int imb_finalize(void) {
  imb::State& s = state();
  if (!s.initialised) return IMB_OK;
  cudaDeviceSynchronize();
  if (s.ws) cudaFree(s.ws);
  if (s.seeds) cudaFree(s.seeds);
  if (s.aux) cudaFree(s.aux);
  s.seeds = s.aux = nullptr;
  s.seeds_bytes = s.aux_bytes = 0;
  if (s.pinned) cudaFreeHost(s.pinned);
  s.ws = nullptr;
  s.ws_bytes = 0;
  s.pinned = nullptr;
  s.initialised = false;
  s.device = -1;
  return IMB_OK;
}

int imb_version(void) { return IMB_VERSION; }

const char* imb_strerror(int code) {
  switch (code) {
    case IMB_OK: return "ok";
    case IMB_EINVAL: return "invalid argument";
    case IMB_ECUDA: return "CUDA error (see imb_last_cuda_error)";
    case IMB_ENOMEM: return "out of device memory";
    case IMB_ENOTINIT: return "imb_init has not been called";
    case IMB_EBOUNDS: return "one of the requested points is out of the grid bounds";
    case IMB_ESINGULAR: return "singular matrix";
    case IMB_EUNSUPPORTED: return "unsupported configuration";
    case IMB_ESTATE:
      return "the scratch workspace was used by another call between the phases of a sharded evaluation";
    default: return "unknown error code";
  }
}

const char* imb_last_cuda_error(void) { return state().cuda_err; }

int64_t imb_launch_count(void) { return state().launches; }

}  // extern "C"
