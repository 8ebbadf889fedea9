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

// ------------------------------------------------------------------------------------------------ order statistics
// The 2nd / 98th percentile need four order statistics per (image, band).  They are found exactly with two radix levels
// of 256 bins, both counted in shared memory (a flat 65536-bin histogram per band would have to live in global memory:
// 780 us of global atomics for 100 M samples, measured):
//   level 1: histogram of the high byte (the value itself for 8-bit data)  -> bucket h_j and rank-within-bucket r_j
//   level 2: histogram of the low byte of the samples whose high byte is h_j -> low byte l_j;  value = h_j << 8 | l_j
// Workspace layout per (image, band): hi[256] u32 | lo[4][256] u32 | sel[4] {bucket, rank} | pct[2] f64 | lut[nvals] u8.
constexpr int WARPS = 8;

struct Sel {
  int bucket;
  unsigned rank;
};

// grid (blocks, B); LEVEL 1: bins[b][c][256] += high byte counts.  LEVEL 2: bins[b][c][4][256] += low byte counts of the
// samples in the four selected buckets.  Per-warp shared-memory histograms, flushed once per block.
template <typename T, int LEVEL>
__global__ void __launch_bounds__(WARPS * 32)
radix_hist_kernel(const T* __restrict__ img, long HW, int C, int copies, const Sel* __restrict__ sel,
                  unsigned* __restrict__ bins) {
  extern __shared__ unsigned sh[];    // `copies` private histograms (warps share one round-robin): fewer same-bin conflicts
  constexpr int PER_BAND = LEVEL == 1 ? 256 : 1024;
  const int n_sh = C * PER_BAND;
  unsigned* mine = sh + ((threadIdx.x >> 5) % copies) * n_sh;
  for (int i = threadIdx.x; i < copies * n_sh; i += blockDim.x) sh[i] = 0;
  __shared__ int tgt[8 * 4];
  if (LEVEL == 2 && threadIdx.x < C * 4) tgt[threadIdx.x] = sel[(long)blockIdx.y * C * 4 + threadIdx.x].bucket;
  __syncthreads();
  const T* src = img + (long)blockIdx.y * HW * C;
  const long total = HW * C;
  constexpr int SHIFT = sizeof(T) == 2 ? 8 : 0;
  auto count = [&](unsigned v, int c) {
    if (LEVEL == 1) {
      atomicAdd(&mine[c * 256 + (v >> SHIFT)], 1u);
    } else {
      const int hi = (int)(v >> 8);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (hi == tgt[c * 4 + j]) atomicAdd(&mine[(c * 4 + j) * 256 + (v & 255u)], 1u);
    }
  };
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (total % VEC == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec_ok) {
    const long nvec = total / VEC;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + i);
      const unsigned w[4] = {q.x, q.y, q.z, q.w};
      int c = (int)((i * VEC) % C);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const unsigned word = w[e * (int)sizeof(T) / 4];
        const unsigned v = sizeof(T) == 2 ? ((word >> (16 * (e & 1))) & 0xffffu) : ((word >> (8 * (e & 3))) & 0xffu);
        count(v, c);
        if (++c == C) c = 0;
      }
    }
  } else {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
      count((unsigned)src[i], (int)(i % C));
  }
  __syncthreads();
  unsigned* dst = bins + (long)blockIdx.y * n_sh;
  for (int i = threadIdx.x; i < n_sh; i += blockDim.x) {
    unsigned s = 0;
    for (int w = 0; w < copies; ++w) s += sh[w * n_sh + i];
    if (s) atomicAdd(&dst[i], s);
  }
}

// One block (256 threads) per (image, band).  LEVEL 1: ranks ks[j] -> (bucket, rank inside the bucket).  LEVEL 2 (or
// LEVEL 1 of 8-bit data with `final`): the bucket index completes the value; numpy's _lerp gives the percentiles.
__device__ __forceinline__ int find_rank(const unsigned* __restrict__ h, unsigned long long need, unsigned long long* scan,
                                         unsigned long long* before_out) {
  // inclusive scan of 256 bins in shared memory (Hillis-Steele), then the first bin whose cumulative count reaches need
  __shared__ int found;
  __shared__ unsigned long long found_before;
  const unsigned v = h[threadIdx.x];
  scan[threadIdx.x] = v;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    const unsigned long long add = threadIdx.x >= off ? scan[threadIdx.x - off] : 0;
    __syncthreads();
    scan[threadIdx.x] += add;
    __syncthreads();
  }
  const unsigned long long incl = scan[threadIdx.x], excl = incl - v;
  if (excl < need && need <= incl) {
    found = threadIdx.x;
    found_before = excl;
  }
  __syncthreads();
  const int f = found;
  *before_out = found_before;
  __syncthreads();
  return f;
}

template <int LEVEL>
__global__ void __launch_bounds__(256)
select_kernel(const unsigned* __restrict__ bins, long k0, long k1, long k2, long k3, double t_lo, double t_hi, int final,
              Sel* __restrict__ sel, double* __restrict__ pct) {
  __shared__ unsigned long long scan[256];
  __shared__ int value[4];
  const long ks[4] = {k0, k1, k2, k3};
  Sel* my = sel + (long)blockIdx.x * 4;
  for (int j = 0; j < 4; ++j) {
    unsigned long long before;
    int f;
    if (LEVEL == 1) {
      f = find_rank(bins + (long)blockIdx.x * 256, (unsigned long long)ks[j] + 1, scan, &before);
      if (threadIdx.x == 0) {
        my[j].bucket = f;
        my[j].rank = (unsigned)((unsigned long long)ks[j] - before);
        value[j] = f;
      }
    } else {
      f = find_rank(bins + ((long)blockIdx.x * 4 + j) * 256, (unsigned long long)my[j].rank + 1, scan, &before);
      if (threadIdx.x == 0) value[j] = (my[j].bucket << 8) | f;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && (LEVEL == 2 || final)) {
    const double ts[2] = {t_lo, t_hi};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double a = (double)value[2 * j], b = (double)value[2 * j + 1];
      const double diff = __dsub_rn(b, a);
      double v = __dadd_rn(a, __dmul_rn(diff, ts[j]));
      if (ts[j] >= 0.5) v = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, ts[j])));
      pct[blockIdx.x * 2 + j] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ stretch
// The result depends only on (image, band, value): evaluate the reference's float64 expression once per value into a
// byte table, then the image pass is a gather (one read, one byte written per sample).
__global__ void __launch_bounds__(256)
lut_kernel(const double* __restrict__ pct, int C, int nvals, int enhance_water, uint8_t* __restrict__ lut) {
  const int bc = blockIdx.y;                   // image * C + band
  const int v0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (v0 >= nvals) return;
  const double p2 = pct[2 * bc], p98 = pct[2 * bc + 1];
  double v = __dmul_rn(__ddiv_rn(__dsub_rn((double)v0, p2), __dsub_rn(p98, p2)), 255.0);
  // np.clip keeps NaN (constant band: 0/0); the integer conversion of NaN is defined as 0 here
  if (v != v) v = 0.0;
  v = fmin(fmax(v, 0.0), 255.0);
  if (enhance_water && (bc % C) == 0 && v < 100.0) v = __dmul_rn(v, 0.7);
  lut[(long)bc * nvals + v0] = (uint8_t)(int)v;   // truncation toward zero, as the reference's integer-array assignment
}

template <typename T>
__global__ void __launch_bounds__(256)
apply_lut_kernel(const T* __restrict__ img, long HW, int C, int nvals, const uint8_t* __restrict__ lut,
                 uint8_t* __restrict__ out) {
  const T* src = img + (long)blockIdx.y * HW * C;
  uint8_t* dst = out + (long)blockIdx.y * HW * C;
  const uint8_t* tab = lut + (long)blockIdx.y * C * nvals;
  const long total = HW * C;
  constexpr int VEC = 8;      // samples per thread and iteration: one 8-byte store
  const bool vec_ok = (total % VEC == 0) && ((reinterpret_cast<uintptr_t>(src) & (VEC * sizeof(T) - 1)) == 0) &&
                      ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
  if (vec_ok) {
    const long nvec = total / VEC;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
      unsigned v[VEC];
      if (sizeof(T) == 2) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + i);
        v[0] = q.x & 0xffffu; v[1] = q.x >> 16; v[2] = q.y & 0xffffu; v[3] = q.y >> 16;
        v[4] = q.z & 0xffffu; v[5] = q.z >> 16; v[6] = q.w & 0xffffu; v[7] = q.w >> 16;
      } else {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(src) + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[e] = (q.x >> (8 * e)) & 0xffu; v[4 + e] = (q.y >> (8 * e)) & 0xffu; }
      }
      int c = (int)((i * VEC) % C);
      unsigned lo = 0, hi = 0;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const unsigned r = __ldg(tab + (long)c * nvals + v[e]);
        if (e < 4) lo |= r << (8 * e); else hi |= r << (8 * (e - 4));
        if (++c == C) c = 0;
      }
      reinterpret_cast<uint2*>(dst)[i] = make_uint2(lo, hi);
    }
  } else {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
      dst[i] = __ldg(tab + (long)(i % C) * nvals + (unsigned)src[i]);
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

// Fast path (W % 4 == 0, 4-byte aligned images): four horizontally adjacent pixels per thread as one packed word.  The
// tile (128 x 8 outputs + margins) sits in shared memory as aligned 32-bit words; the window byte at offset o of a row is
// the funnel shift of words o/4 and o/4 + 1, and the running maximum is a per-byte __vmaxu4.
constexpr int CT_W = 128, CT_H = 8;
__global__ void __launch_bounds__(256)
coastline_vec_kernel(const uint8_t* __restrict__ mask, int H, int W, const __grid_constant__ Spans sp,
                     uint8_t* __restrict__ out) {
  extern __shared__ uint32_t wt[];
  const int k = sp.k;
  const int L4 = (sp.ax + 3) & ~3;                       // left margin rounded up to whole words
  const int R4 = ((k - 1 - sp.ax + 3) & ~3) + 4;         // right margin + one word for the funnel shift
  const int roww = (L4 + CT_W + R4) >> 2;                // words per tile row
  const int rows = CT_H + k - 1;
  const long img = (long)blockIdx.z * H * W;
  const int y0 = blockIdx.y * CT_H, x0 = blockIdx.x * CT_W;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(mask + img);
  const int Ww = W >> 2;
  for (int i = threadIdx.x; i < roww * rows; i += 256) {
    const int ty = i / roww, tx = i - ty * roww;
    const int y = y0 + ty - sp.ay;
    const int xw = ((x0 - L4) >> 2) + tx;                // word column in the image (x0, L4 multiples of 4)
    wt[i] = (y >= 0 && y < H && xw >= 0 && xw < Ww) ? __ldg(src + (long)y * Ww + xw) : 0u;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + 4 * tx, y = y0 + ty;
  if (x >= W || y >= H) return;
  uint32_t m = 0;
  for (int i = 0; i < k; ++i) {
    const int j1 = sp.j1[i], j2 = sp.j2[i];
    if (j1 >= j2) continue;
    const uint32_t* rowp = wt + (ty + i) * roww;
    int o = 4 * tx + L4 - sp.ax + j1;                    // byte offset of the first window position
    int wi = o >> 2;
    uint32_t lo = rowp[wi], hi = rowp[wi + 1];
    for (int j = j1; j < j2; ++j, ++o) {
      const int sh = (o & 3) * 8;
      if (sh == 0 && j != j1) { ++wi; lo = hi; hi = rowp[wi + 1]; }
      m = __vmaxu4(m, __funnelshift_r(lo, hi, sh));
    }
  }
  const uint32_t center = wt[(ty + sp.ay) * roww + tx + (L4 >> 2)];
  reinterpret_cast<uint32_t*>(out + img)[(long)y * Ww + (x >> 2)] = __vsub4(m, center);   // per-byte wrapping subtraction
}

// Larger structuring elements: horizontal window maxima by doubling.  Level l of a tile row holds, at every byte, the
// maximum over the 2^l bytes starting there (level 0 = the tile); a window [a, a + len) is then the maximum of two level
// floor(log2 len) entries at a and a + len - 2^l -- two funnel-shifted words per kernel row instead of len bytes.
__device__ __forceinline__ uint32_t word_at(const uint32_t* rowp, int o) {      // 4 bytes starting at byte offset o
  const int wi = o >> 2;
  return __funnelshift_r(rowp[wi], rowp[wi + 1], (o & 3) * 8);
}

__global__ void __launch_bounds__(256)
coastline_lvl_kernel(const uint8_t* __restrict__ mask, int H, int W, const __grid_constant__ Spans sp, int levels,
                     uint8_t* __restrict__ out) {
  extern __shared__ uint32_t wt[];      // [levels][rows][roww]
  const int k = sp.k;
  const int L4 = (sp.ax + 3) & ~3;
  const int R4 = ((k - 1 - sp.ax + 3) & ~3) + 4;
  const int roww = (L4 + CT_W + R4) >> 2;
  const int rows = CT_H + k - 1;
  const int lsz = roww * rows;
  const long img = (long)blockIdx.z * H * W;
  const int y0 = blockIdx.y * CT_H, x0 = blockIdx.x * CT_W;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(mask + img);
  const int Ww = W >> 2;
  for (int i = threadIdx.x; i < lsz; i += 256) {
    const int ty = i / roww, tx = i - ty * roww;
    const int y = y0 + ty - sp.ay;
    const int xw = ((x0 - L4) >> 2) + tx;
    wt[i] = (y >= 0 && y < H && xw >= 0 && xw < Ww) ? __ldg(src + (long)y * Ww + xw) : 0u;
  }
  __syncthreads();
  for (int l = 1; l < levels; ++l) {
    const uint32_t* prev = wt + (l - 1) * lsz;
    uint32_t* cur = wt + l * lsz;
    const int sb = 1 << (l - 1);                         // byte shift between the two halves of the window
    for (int i = threadIdx.x; i < lsz; i += 256) {
      const int tx = i % roww;
      const uint32_t a = prev[i];
      uint32_t b;
      if (sb < 4) {
        const uint32_t nx = tx + 1 < roww ? prev[i + 1] : 0u;
        b = __funnelshift_r(a, nx, sb * 8);
      } else {
        const int dw = sb >> 2;
        b = tx + dw < roww ? prev[i + dw] : 0u;
      }
      cur[i] = __vmaxu4(a, b);
    }
    __syncthreads();
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + 4 * tx, y = y0 + ty;
  if (x >= W || y >= H) return;
  uint32_t m = 0;
  const int o0 = 4 * tx + L4 - sp.ax;
  for (int i = 0; i < k; ++i) {
    const int j1 = sp.j1[i], len = sp.j2[i] - j1;
    if (len <= 0) continue;
    const int l = 31 - __clz(len);
    const uint32_t* rowp = wt + l * lsz + (ty + i) * roww;
    m = __vmaxu4(m, word_at(rowp, o0 + j1));
    m = __vmaxu4(m, word_at(rowp, o0 + j1 + len - (1 << l)));
  }
  const uint32_t center = wt[(ty + sp.ay) * roww + tx + (L4 >> 2)];
  reinterpret_cast<uint32_t*>(out + img)[(long)y * Ww + (x >> 2)] = __vsub4(m, center);
}

int nearbyint_even(double x) { return (int)nearbyint(x); }   // cvRound under the default rounding mode

}  // namespace

namespace {
struct EnhanceLayout {
  size_t hi, lo, sel, pct, lut, total;
};
EnhanceLayout enhance_layout(int B, int C, int bits) {
  const size_t bc = (size_t)B * C, nvals = bits == 8 ? 256 : 65536;
  EnhanceLayout L;
  L.hi = 0;
  L.lo = L.hi + bc * 256 * sizeof(unsigned);
  L.sel = L.lo + bc * 1024 * sizeof(unsigned);
  L.pct = L.sel + bc * 4 * sizeof(Sel);
  L.lut = L.pct + bc * 2 * sizeof(double);
  L.total = L.lut + bc * nvals;
  L.total = (L.total + 255) & ~(size_t)255;
  return L;
}
}  // namespace

extern "C" size_t rbu_enhance_workspace_bytes(int B, int C, int bits) {
  if (B <= 0 || C <= 0 || C > 8 || (bits != 8 && bits != 16)) return 0;
  return enhance_layout(B, C, bits).total;
}

extern "C" int rbu_enhance_image(const void* img, int bits, int B, int H, int W, int C, int enhance_water, uint8_t* out,
                                 double* percentiles_out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBU_CHECK_ARG(img && out && workspace && B > 0 && H > 0 && W > 0, "rbu_enhance_image: bad arguments");
  RBU_CHECK_ARG(bits == 8 || bits == 16, "rbu_enhance_image: bits must be 8 or 16 (integer digital numbers)");
  RBU_CHECK_ARG(C >= 1 && C <= 8, "rbu_enhance_image: 1..8 bands");
  RBU_CHECK_ARG(B <= 65535, "rbu_enhance_image: at most 65535 images per call");
  RBU_CHECK_ARG(workspace_bytes >= rbu_enhance_workspace_bytes(B, C, bits), "rbu_enhance_image: workspace too small");
  RBU_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "rbu_enhance_image: workspace must be 16-byte aligned");
  const EnhanceLayout L = enhance_layout(B, C, bits);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  unsigned* hi = reinterpret_cast<unsigned*>(ws + L.hi);
  unsigned* lo = reinterpret_cast<unsigned*>(ws + L.lo);
  Sel* sel = reinterpret_cast<Sel*>(ws + L.sel);
  double* pct = reinterpret_cast<double*>(ws + L.pct);
  uint8_t* lut = ws + L.lut;
  const int nvals = bits == 8 ? 256 : 65536;
  const long HW = (long)H * W;
  RBU_CHECK_CUDA(cudaMemsetAsync(ws, 0, L.sel, stream));   // both histogram levels

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

  long blocks = (HW * C + 256L * 64 - 1) / (256L * 64);
  const long cap = rbu_cdiv((long)rbu_num_sms() * 4, B);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const dim3 grid((unsigned)blocks, (unsigned)B);
  // level 1: up to 8 private copies within 48 KB of shared memory; level 2 touches ~1% of the samples: one copy
  int copies = 48 / C;
  if (copies > WARPS) copies = WARPS;
  const size_t sh1 = (size_t)copies * C * 256 * sizeof(unsigned), sh2 = (size_t)C * 1024 * sizeof(unsigned);
  if (bits == 8) {
    radix_hist_kernel<uint8_t, 1><<<grid, WARPS * 32, sh1, stream>>>(static_cast<const uint8_t*>(img), HW, C, copies, nullptr, hi);
    RBU_CHECK_LAUNCH();
    select_kernel<1><<<B * C, 256, 0, stream>>>(hi, ks[0], ks[1], ks[2], ks[3], ts[0], ts[1], 1, sel, pct);
    RBU_CHECK_LAUNCH();
  } else {
    const uint16_t* src = static_cast<const uint16_t*>(img);
    radix_hist_kernel<uint16_t, 1><<<grid, WARPS * 32, sh1, stream>>>(src, HW, C, copies, nullptr, hi);
    RBU_CHECK_LAUNCH();
    select_kernel<1><<<B * C, 256, 0, stream>>>(hi, ks[0], ks[1], ks[2], ks[3], ts[0], ts[1], 0, sel, pct);
    RBU_CHECK_LAUNCH();
    radix_hist_kernel<uint16_t, 2><<<grid, WARPS * 32, sh2, stream>>>(src, HW, C, 1, sel, lo);
    RBU_CHECK_LAUNCH();
    select_kernel<2><<<B * C, 256, 0, stream>>>(lo, ks[0], ks[1], ks[2], ks[3], ts[0], ts[1], 1, sel, pct);
    RBU_CHECK_LAUNCH();
  }
  lut_kernel<<<dim3(nvals / 256, B * C), 256, 0, stream>>>(pct, C, nvals, enhance_water, lut);
  RBU_CHECK_LAUNCH();
  long ablocks = (HW * C + 256L * 8 * 4 - 1) / (256L * 8 * 4);
  const long acap = rbu_cdiv((long)rbu_num_sms() * 16, B);
  if (ablocks > acap) ablocks = acap;
  if (ablocks < 1) ablocks = 1;
  const dim3 agrid((unsigned)ablocks, (unsigned)B);
  if (bits == 8)
    apply_lut_kernel<uint8_t><<<agrid, 256, 0, stream>>>(static_cast<const uint8_t*>(img), HW, C, nvals, lut, out);
  else
    apply_lut_kernel<uint16_t><<<agrid, 256, 0, stream>>>(static_cast<const uint16_t*>(img), HW, C, nvals, lut, out);
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
  if (ksize <= 8 && W % 4 == 0 && (((uintptr_t)mask | (uintptr_t)out) & 3) == 0) {
    const int L4 = (sp.ax + 3) & ~3, R4 = ((ksize - 1 - sp.ax + 3) & ~3) + 4;
    const int smem = ((L4 + CT_W + R4) >> 2) * (CT_H + ksize - 1) * 4;
    const dim3 grid((unsigned)rbu_cdiv(W, CT_W), (unsigned)rbu_cdiv(H, CT_H), (unsigned)B);
    coastline_vec_kernel<<<grid, 256, smem, (cudaStream_t)stream_>>>(mask, H, W, sp, out);
    RBU_CHECK_LAUNCH();
    return RBU_OK;
  }
  if (W % 4 == 0 && (((uintptr_t)mask | (uintptr_t)out) & 3) == 0) {
    int maxlen = 1;
    for (int i = 0; i < ksize; ++i) maxlen = sp.j2[i] - sp.j1[i] > maxlen ? sp.j2[i] - sp.j1[i] : maxlen;
    int levels = 1;
    while ((2 << (levels - 1)) <= maxlen) ++levels;           // levels = floor(log2(maxlen)) + 1
    const int L4 = (sp.ax + 3) & ~3, R4 = ((ksize - 1 - sp.ax + 3) & ~3) + 4;
    const size_t smem = (size_t)levels * ((L4 + CT_W + R4) >> 2) * (CT_H + ksize - 1) * 4;
    if (smem <= 200 * 1024) {
      static size_t attr = 0;
      if (smem > 48 * 1024 && smem > attr) {
        RBU_CHECK_CUDA(cudaFuncSetAttribute(coastline_lvl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = 200 * 1024;
      }
      const dim3 grid((unsigned)rbu_cdiv(W, CT_W), (unsigned)rbu_cdiv(H, CT_H), (unsigned)B);
      coastline_lvl_kernel<<<grid, 256, smem, (cudaStream_t)stream_>>>(mask, H, W, sp, levels, out);
      RBU_CHECK_LAUNCH();
      return RBU_OK;
    }
  }
  const int tw = DT + ksize - 1;
  const dim3 grid((unsigned)rbu_cdiv(W, DT), (unsigned)rbu_cdiv(H, DT), (unsigned)B);
  coastline_kernel<<<grid, DT * 8, tw * tw, (cudaStream_t)stream_>>>(mask, H, W, sp, out);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}
