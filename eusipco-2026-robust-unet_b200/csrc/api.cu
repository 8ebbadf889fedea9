// Library-level entry points: version, error reporting, device checks, tensor-map encoding.
#include <stdarg.h>
#include <stdlib.h>

#include "rbu_common.cuh"
#include "tma_host.cuh"

namespace {
thread_local char g_err[1024] = "";
constexpr int MAX_DEVICES = 64;
std::atomic<int> g_sms[MAX_DEVICES];          // per device ordinal; 0 = not queried yet
}  // namespace

std::atomic<unsigned long long> g_rbu_launches{0};

extern "C" unsigned long long rbu_launch_count(void) { return g_rbu_launches.load(std::memory_order_relaxed); }

bool rbu_first_use_on_device(std::atomic<unsigned long long>* done_mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return true;   // unknown: set the attribute again
  const unsigned long long bit = 1ull << dev;
  return (done_mask->fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

void rbu_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int rbu_num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 148;
  int n = g_sms[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

int rbu_stream_blocks(long items_per_image, int per_iter, int images) {
  // Measured on B200 (training step, batch 64 at 256^2, same-call A/B of the minimum iterations per thread):
  // 1 -> 16.3-16.8 ms in the bandwidth-bound kernels, 4 -> 16.0-16.2, 8 -> 15.8-16.1, 16 / 32 -> 16.2-16.4 (too few blocks).
  constexpr int MIN_ITERS = 8, FLOOR_BLOCKS_PER_SM = 3;
  const long sms = rbu_num_sms();
  if (images < 1) images = 1;
  const long full = (items_per_image + per_iter - 1) / per_iter;                 // one iteration per thread
  long b = (items_per_image + (long)per_iter * MIN_ITERS - 1) / ((long)per_iter * MIN_ITERS);
  const long need = (FLOOR_BLOCKS_PER_SM * sms + images - 1) / images;            // keep ~3 blocks per SM in flight
  if (b < need) b = full < need ? full : need;
  long cap = (sms * 32 + images - 1) / images;
  if (cap < 1) cap = 1;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

extern "C" int rbu_version(void) { return 100; }

extern "C" const char* rbu_last_error(void) { return g_err; }

extern "C" int rbu_sm_count(void) { return rbu_num_sms(); }

extern "C" int rbu_device_check(void) {
  int dev = 0;
  RBU_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0;
  RBU_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    rbu_set_error("rbunet needs a compute-capability 10.x (Blackwell, sm_100a) device, found %d.x", major);
    return RBU_ERR_UNSUPPORTED;
  }
  return RBU_OK;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int rbu_encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box) {
  return rbu_encode_tmap_bf16_sw(out, base, rank, dims, strides_bytes, box, 128);
}

int rbu_encode_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) {
      rbu_set_error("cuTensorMapEncodeTiled is not available from the driver (%s)", cudaGetErrorString(e));
      return RBU_ERR_CUDA;
    }
    fn = reinterpret_cast<encode_tiled_fn>(sym);
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    rbu_set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u] "
                  "stride0 %llu base %p",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                  (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
                  rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), base);
    return RBU_ERR_CUDA;
  }
  return RBU_OK;
}
