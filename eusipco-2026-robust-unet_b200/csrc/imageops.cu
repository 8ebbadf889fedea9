// the reference's host-side image operations either side of the network, on the GPU.
//
// rbu_enhance_image  -- tif_to_image.py:139-171 `enhance_image` (duplicated at train_water_segmentation.py:103-174 and
//                       predict_coastline.py:545-581): per band np.percentile(band, [2, 98]) ("linear" method), stretch
//                       clip((x - p2) / (p98 - p2) * 255, 0, 255), band 0 multiplied by 0.7 below 100, truncation to the
//                       integer input type and then uint8.  Integer / byte work: bit-exact.  The percentiles come from
//                       an exact histogram (order statistics of integer data) and numpy's _lerp in float64; the stretch
//                       is three explicitly rounded float64 operations per sample, as numpy evaluates it.
// rbu_coastline_mask -- predict_coastline.py:595-602: cv2.dilate(mask, getStructuringElement(MORPH_ELLIPSE, (k, k)))
//                       minus mask.  The element is restated from OpenCV (row spans [j1, j2) per kernel row); pixels
//                       outside the image do not take part (cv2's default border for dilation).
// Both are bound by HBM (one read + one write of a byte image; the histogram pass is one more read).
#include "rbu_common.cuh"

#include <math.h>

namespace {

constexpr int MAX_K = 64;

// ------------------------------------------------------------------------------------------------ histogram
// grid (blocks, B): interleaved [H*W][C] samples of image b -> hist[b][c][nbins].  8-bit data: per-block shared-memory
// histograms (C <= 8 bands); 16-bit data: global atomics (65536 bins per band do not fit in shared memory).
template <typename T>
__global__ void __launch_bounds__(256)
hist_kernel(const T* __restrict__ img, long HW, int C, int nbins, unsigned* __restrict__ hist) {
  extern __shared__ unsigned sh[];
  const bool use_sh = sizeof(T) == 1;
  if (use_sh) {
    for (int i = threadIdx.x; i < C * 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
  }
  const T* src = img + (long)blockIdx.y * HW * C;
  unsigned* h = hist + (long)blockIdx.y * C * nbins;
  const long total = HW * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const unsigned v = src[i];
    if (use_sh)
      atomicAdd(&sh[c * 256 + v], 1u);
    else
      atomicAdd(&h[(long)c * nbins + v], 1u);
  }
  if (use_sh) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * 256; i += blockDim.x)
      if (sh[i]) atomicAdd(&h[i], sh[i]);
  }
}

// ------------------------------------------------------------------------------------------------ percentiles
// One block per (image, band): cumulative histogram -> the four order statistics (previous / next index of the 2nd and
// 98th percentile) -> numpy's _lerp.  pct[b][c] = {p2, p98}.
__global__ void __launch_bounds__(1024)
percentile_kernel(const unsigned* __restrict__ hist, int nbins, long k0, long k1, double t_lo, long k2, long k3,
                  double t_hi, double* __restrict__ pct) {
  __shared__ unsigned long long part[1024];
  __shared__ int found[4];
  const unsigned* h = hist + (long)blockIdx.x * nbins;
  const int per = nbins / 1024 > 0 ? nbins / 1024 : 1;       // 64 bins per thread (16-bit), 1 for 8-bit (256 threads busy)
  const int first = threadIdx.x * per;
  unsigned long long s = 0;
  if (first < nbins)
    for (int i = 0; i < per; ++i) s += h[first + i];
  part[threadIdx.x] = s;
  __syncthreads();
  // inclusive scan of the 1024 partial sums (Hillis-Steele; one block, negligible)
  for (int off = 1; off < 1024; off <<= 1) {
    unsigned long long v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  const unsigned long long before = part[threadIdx.x] - s;   // samples in bins below this thread's range
  const long ks[4] = {k0, k1, k2, k3};
  if (first < nbins) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned long long need = (unsigned long long)ks[j] + 1;   // smallest value v with cum(v) >= k + 1
      if (before < need && need <= before + s) {
        unsigned long long run = before;
        for (int i = 0; i < per; ++i) {
          run += h[first + i];
          if (run >= need) { found[j] = first + i; break; }
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double r[2];
    const double ts[2] = {t_lo, t_hi};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double a = (double)found[2 * j], b = (double)found[2 * j + 1];
      const double diff = __dsub_rn(b, a);
      double v = __dadd_rn(a, __dmul_rn(diff, ts[j]));
      if (ts[j] >= 0.5) v = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, ts[j])));
      r[j] = v;
    }
    pct[blockIdx.x * 2] = r[0];
    pct[blockIdx.x * 2 + 1] = r[1];
  }
}

// ------------------------------------------------------------------------------------------------ stretch
template <typename T>
__global__ void __launch_bounds__(256)
stretch_kernel(const T* __restrict__ img, long HW, int C, const double* __restrict__ pct, int enhance_water,
               uint8_t* __restrict__ out) {
  const T* src = img + (long)blockIdx.y * HW * C;
  uint8_t* dst = out + (long)blockIdx.y * HW * C;
  const double* pc = pct + (long)blockIdx.y * C * 2;
  const long total = HW * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const double p2 = pc[2 * c], p98 = pc[2 * c + 1];
    double v = __dmul_rn(__ddiv_rn(__dsub_rn((double)src[i], p2), __dsub_rn(p98, p2)), 255.0);
    // np.clip keeps NaN (constant band: 0/0); the integer conversion of NaN is defined as 0 here
    if (v != v) v = 0.0;
    v = fmin(fmax(v, 0.0), 255.0);
    if (enhance_water && c == 0 && v < 100.0) v = __dmul_rn(v, 0.7);
    dst[i] = (uint8_t)(int)v;   // truncation toward zero, as the reference's assignment into an integer array
  }
}

// ------------------------------------------------------------------------------------------------ dilation - mask
struct Spans {
  int k, ay, ax;
  signed char j1[MAX_K], j2[MAX_K];
};

constexpr int DT = 32;   // output tile edge

__global__ void __launch_bounds__(DT * 8)
coastline_kernel(const uint8_t* __restrict__ mask, int H, int W, const __grid_constant__ Spans sp,
                 uint8_t* __restrict__ out) {
  extern __shared__ uint8_t tile[];   // (DT + k - 1)^2 bytes: the tile and its footprint margin, 0 outside the image
  const int k = sp.k, tw = DT + k - 1;
  const long img = (long)blockIdx.z * H * W;
  const int y0 = blockIdx.y * DT, x0 = blockIdx.x * DT;
  for (int i = threadIdx.x; i < tw * tw; i += blockDim.x) {
    const int ty = i / tw, tx = i - ty * tw;
    const int y = y0 + ty - sp.ay, x = x0 + tx - sp.ax;
    tile[i] = (y >= 0 && y < H && x >= 0 && x < W) ? mask[img + (long)y * W + x] : (uint8_t)0;
  }
  __syncthreads();
  const int lx = threadIdx.x & (DT - 1);
  for (int ly = threadIdx.x / DT; ly < DT; ly += blockDim.x / DT) {
    const int y = y0 + ly, x = x0 + lx;
    if (y >= H || x >= W) continue;
    unsigned m = 0;
    for (int i = 0; i < k; ++i) {
      const uint8_t* rowp = tile + (ly + i) * tw + lx;
      for (int j = sp.j1[i]; j < sp.j2[i]; ++j) m = max(m, (unsigned)rowp[j]);
    }
    const uint8_t center = tile[(ly + sp.ay) * tw + lx + sp.ax];
    out[img + (long)y * W + x] = (uint8_t)(m - center);   // uint8 arithmetic of the reference (m >= center)
  }
}

int nearbyint_even(double x) { return (int)nearbyint(x); }   // cvRound under the default rounding mode

}  // namespace

extern "C" size_t rbu_enhance_workspace_bytes(int B, int C, int bits) {
  if (B <= 0 || C <= 0 || (bits != 8 && bits != 16)) return 0;
  const size_t nbins = bits == 8 ? 256 : 65536;
  return (size_t)B * C * nbins * sizeof(unsigned) + (size_t)B * C * 2 * sizeof(double);
}

extern "C" int rbu_enhance_image(const void* img, int bits, int B, int H, int W, int C, int enhance_water, uint8_t* out,
                                 double* percentiles_out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBU_CHECK_ARG(img && out && workspace && B > 0 && H > 0 && W > 0, "rbu_enhance_image: bad arguments");
  RBU_CHECK_ARG(bits == 8 || bits == 16, "rbu_enhance_image: bits must be 8 or 16 (integer digital numbers)");
  RBU_CHECK_ARG(C >= 1 && C <= 8, "rbu_enhance_image: 1..8 bands");
  RBU_CHECK_ARG(workspace_bytes >= rbu_enhance_workspace_bytes(B, C, bits), "rbu_enhance_image: workspace too small");
  RBU_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "rbu_enhance_image: workspace must be 16-byte aligned");
  const int nbins = bits == 8 ? 256 : 65536;
  const long HW = (long)H * W;
  unsigned* hist = reinterpret_cast<unsigned*>(workspace);
  double* pct = reinterpret_cast<double*>(hist + (size_t)B * C * nbins);
  RBU_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)B * C * nbins * sizeof(unsigned), stream));

  // numpy's "linear" percentile: virtual index (n - 1) * q, previous = floor, gamma = virtual - previous
  long ks[4];
  double ts[2];
  const double qs[2] = {2.0 / 100.0, 98.0 / 100.0};
  for (int j = 0; j < 2; ++j) {
    const double virt = (double)(HW - 1) * qs[j];
    long prev = (long)floor(virt);
    long next = prev + 1;
    ts[j] = virt - (double)prev;
    if (virt >= (double)(HW - 1)) prev = next = HW - 1;
    ks[2 * j] = prev;
    ks[2 * j + 1] = next;
  }

  long blocks = (HW * C + 256L * 16 - 1) / (256L * 16);
  const long cap = (long)rbu_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const dim3 grid((unsigned)blocks, (unsigned)B);
  if (bits == 8)
    hist_kernel<uint8_t><<<grid, 256, C * 256 * sizeof(unsigned), stream>>>(static_cast<const uint8_t*>(img), HW, C, nbins, hist);
  else
    hist_kernel<uint16_t><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(img), HW, C, nbins, hist);
  RBU_CHECK_LAUNCH();
  percentile_kernel<<<B * C, 1024, 0, stream>>>(hist, nbins, ks[0], ks[1], ts[0], ks[2], ks[3], ts[1], pct);
  RBU_CHECK_LAUNCH();
  if (bits == 8)
    stretch_kernel<uint8_t><<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(img), HW, C, pct, enhance_water, out);
  else
    stretch_kernel<uint16_t><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(img), HW, C, pct, enhance_water, out);
  RBU_CHECK_LAUNCH();
  if (percentiles_out)
    RBU_CHECK_CUDA(cudaMemcpyAsync(percentiles_out, pct, (size_t)B * C * 2 * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  return RBU_OK;
}

extern "C" int rbu_coastline_mask(const uint8_t* mask, int B, int H, int W, int ksize, uint8_t* out, void* stream_) {
  RBU_CHECK_ARG(mask && out && B > 0 && H > 0 && W > 0, "rbu_coastline_mask: bad arguments");
  RBU_CHECK_ARG(ksize >= 1 && ksize <= MAX_K, "rbu_coastline_mask: kernel size must be 1..%d", MAX_K);
  RBU_CHECK_ARG(mask != out, "rbu_coastline_mask: in-place operation is not supported");
  // OpenCV getStructuringElement(MORPH_ELLIPSE, (k, k)): row i is ones on [j1, j2)
  Spans sp;
  memset(&sp, 0, sizeof(sp));
  sp.k = ksize;
  sp.ay = sp.ax = ksize / 2;
  const int r = ksize / 2, c = ksize / 2;
  const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
  for (int i = 0; i < ksize; ++i) {
    int j1 = 0, j2 = 0;
    if (ksize == 1) {
      j2 = 1;
    } else {
      const int dy = i - r;
      if (abs(dy) <= r) {
        const int dx = nearbyint_even(c * sqrt((r * r - dy * dy) * inv_r2));
        j1 = c - dx > 0 ? c - dx : 0;
        j2 = c + dx + 1 < ksize ? c + dx + 1 : ksize;
      }
    }
    sp.j1[i] = (signed char)j1;
    sp.j2[i] = (signed char)j2;
  }
  const int tw = DT + ksize - 1;
  const dim3 grid((unsigned)rbu_cdiv(W, DT), (unsigned)rbu_cdiv(H, DT), (unsigned)B);
  coastline_kernel<<<grid, DT * 8, tw * tw, (cudaStream_t)stream_>>>(mask, H, W, sp, out);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}
