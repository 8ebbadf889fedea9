// Kernels specific to the plain 2-class U-Net path (train_water_segmentation.py:209-288,304,
// 384-388; predict_coastline.py:390-392): the `final` 1x1 convolution to two logits, its backward, and the fused
// CrossEntropyLoss + argmax confusion counts.  Everything else of that network (3x3 conv + bias, BatchNorm + ReLU,
// max-pool, ConvTranspose, weight gradients) reuses the Robust U-Net kernels.
#include "rbu_common.cuh"

namespace {

constexpr int NT = 256;

// logits[n][k][hw] = w[k] . x[n,hw,:] + b[k], k = 0,1 (fp32, NCHW like the reference's output)
template <int TPP>
__global__ void __launch_bounds__(NT)
head2_fwd_kernel(const bf16* __restrict__ x, long ld, long P, int HW, int C, const float* __restrict__ w,
                 const float* __restrict__ b, float* __restrict__ logits) {
  const int G = C >> 3;
  const int li = threadIdx.x % TPP;
  const int slot = threadIdx.x / TPP;
  constexpr int SLOTS = NT / TPP;
  for (long pb = (long)blockIdx.x * SLOTS; pb < P; pb += (long)gridDim.x * SLOTS) {
    const long p = pb + slot;
    float a0 = 0.f, a1 = 0.f;
    if (p < P)
      for (int cg = li; cg < G; cg += TPP) {
        float v[8];
        unpack8(ld_bf16x8_stream(x + p * ld + cg * 8), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          a0 += __ldg(w + cg * 8 + e) * v[e];
          a1 += __ldg(w + C + cg * 8 + e) * v[e];
        }
      }
#pragma unroll
    for (int o = TPP / 2; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (li == 0 && p < P) {
      const long n = p / HW, pl = p - n * HW;
      logits[(n * 2 + 0) * HW + pl] = a0 + b[0];
      logits[(n * 2 + 1) * HW + pl] = a1 + b[1];
    }
  }
}

// dx = dz0*w0 + dz1*w1;  partials per block: [2][C] (dW) then [2] (db)
__global__ void __launch_bounds__(NT)
head2_bwd_kernel(const float* __restrict__ dlogits, const bf16* __restrict__ x, long x_ld, bf16* __restrict__ dx,
                 long dx_ld, long P, int HW, int C, const float* __restrict__ w, int px_per_block,
                 float* __restrict__ partials) {
  const int G = C >> 3, rows = NT / G;
  const int cg = threadIdx.x % G, row = threadIdx.x / G;
  const long p0 = (long)blockIdx.x * px_per_block;
  const long p1 = min(p0 + px_per_block, P);
  float w0[8], w1[8], a0[8], a1[8];
  float b0 = 0.f, b1 = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) { w0[e] = w[cg * 8 + e]; w1[e] = w[C + cg * 8 + e]; a0[e] = 0.f; a1[e] = 0.f; }
  if (row < rows)
    for (long p = p0 + row; p < p1; p += rows) {
      const long n = p / HW, pl = p - n * HW;
      const float d0 = dlogits[(n * 2 + 0) * HW + pl], d1 = dlogits[(n * 2 + 1) * HW + pl];
      float v[8], o[8];
      unpack8(ld_bf16x8_stream(x + p * x_ld + cg * 8), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a0[e] += d0 * v[e];
        a1[e] += d1 * v[e];
        o[e] = d0 * w0[e] + d1 * w1[e];
      }
      if (cg == 0) { b0 += d0; b1 += d1; }
      st_bf16x8(dx + p * dx_ld + cg * 8, pack8(o));
    }
  __shared__ float sv[NT][8];
  float* dst = partials + (long)blockIdx.x * (2 * C + 8);
  for (int q = 0; q < 2; ++q) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) sv[threadIdx.x][e] = q ? a1[e] : a0[e];
    __syncthreads();
    if (row == 0) {
      float t[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = sv[cg][e];
      for (int r = 1; r < rows; ++r)
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] += sv[r * G + cg][e];
#pragma unroll
      for (int e = 0; e < 8; ++e) dst[q * C + cg * 8 + e] = t[e];
    }
  }
  __syncthreads();
  sv[threadIdx.x][0] = b0;
  sv[threadIdx.x][1] = b1;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int r = 0; r < rows; ++r) { s0 += sv[r * G][0]; s1 += sv[r * G][1]; }
    dst[2 * C] = s0;
    dst[2 * C + 1] = s1;
  }
}

__global__ void __launch_bounds__(256)
colsum2_kernel(const float* __restrict__ part, int nblk, int K, long stride, float* __restrict__ out) {
  __shared__ double sh[8][32];
  const int kx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + kx;
  double s = 0.0;
  if (k < K)
    for (int b = ly; b < nblk; b += 8) s += (double)part[(long)b * stride + k];
  sh[ly][kx] = s;
  __syncthreads();
  if (ly == 0 && k < K) {
    double t = 0.0;
    for (int j = 0; j < 8; ++j) t += sh[j][kx];
    out[k] = (float)t;
  }
}

// nn.CrossEntropyLoss() on two logits (train_water_segmentation.py:304) + argmax counts (:384-388):
// loss_i = logsumexp(z0, z1) - z_target;  pred = argmax (ties -> class 0, like torch.argmax);
// counts per image: TP, FP, FN (class 1 = water), TN by difference.  grid (nblk, B)
__global__ void __launch_bounds__(256)
ce2_partial_kernel(const float* __restrict__ logits, const long long* __restrict__ target, long HW,
                   float* __restrict__ partials, unsigned long long* __restrict__ counts) {
  const int b = blockIdx.y;
  const float* z0 = logits + ((long)b * 2) * HW;
  const float* z1 = z0 + HW;
  const long long* t = target + (long)b * HW;
  float loss = 0.f;
  unsigned tp = 0, fp = 0, fn = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < HW; i += (long)gridDim.x * blockDim.x) {
    const float a = z0[i], c = z1[i];
    const bool y = t[i] != 0;
    const float m = fmaxf(a, c);
    const float lse = m + logf(expf(a - m) + expf(c - m));
    loss += lse - (y ? c : a);
    const bool pb = c > a;
    tp += (pb && y);
    fp += (pb && !y);
    fn += (!pb && y);
  }
  __shared__ float sf[8];
  __shared__ unsigned su[3][8];
  loss = warp_sum(loss);
  tp = __reduce_add_sync(0xffffffffu, tp);
  fp = __reduce_add_sync(0xffffffffu, fp);
  fn = __reduce_add_sync(0xffffffffu, fn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sf[warp] = loss; su[0][warp] = tp; su[1][warp] = fp; su[2][warp] = fn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    unsigned long long c0 = 0, c1 = 0, c2 = 0;
    for (int w = 0; w < 8; ++w) { a += sf[w]; c0 += su[0][w]; c1 += su[1][w]; c2 += su[2][w]; }
    partials[(long)b * gridDim.x + blockIdx.x] = a;
    atomicAdd(&counts[b * 4 + 0], c0);
    atomicAdd(&counts[b * 4 + 1], c1);
    atomicAdd(&counts[b * 4 + 2], c2);
  }
}

__global__ void __launch_bounds__(256)
ce2_finalize_kernel(const float* __restrict__ partials, int n, double total, long HW, int B, float* __restrict__ loss_out,
                    unsigned long long* __restrict__ counts) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)partials[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int j = 0; j < 256; ++j) t += sh[j];
    loss_out[0] = (float)(t / total);
  }
  for (int b = threadIdx.x; b < B; b += 256)
    counts[b * 4 + 3] = (unsigned long long)HW - counts[b * 4] - counts[b * 4 + 1] - counts[b * 4 + 2];
}

// dlogits = (softmax - onehot) * g / (B*HW)
__global__ void __launch_bounds__(256)
ce2_backward_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int B, long HW,
                    const float* __restrict__ gout, float* __restrict__ dlogits) {
  const float g = (gout ? gout[0] : 1.f) / ((float)B * (float)HW);
  const long total = (long)B * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / HW, pl = i - b * HW;
    const float a = logits[(b * 2) * HW + pl], c = logits[(b * 2 + 1) * HW + pl];
    const float p1 = 1.f / (1.f + expf(a - c));      // softmax probability of class 1
    const float y = target[i] != 0 ? 1.f : 0.f;
    dlogits[(b * 2) * HW + pl] = ((1.f - p1) - (1.f - y)) * g;
    dlogits[(b * 2 + 1) * HW + pl] = (p1 - y) * g;
  }
}

int pick_tpp2(int C) {
  const int G = C >> 3;
  int t = 4;
  while (t * 2 <= G && t < 32) t <<= 1;
  return t;
}

}  // namespace

#define VIEW_OK(ptr, ld) ((ptr) != nullptr && ((uintptr_t)(ptr) & 15) == 0 && (ld) % 8 == 0)

extern "C" int rbu_head2_forward(const void* x, int64_t ld, int64_t P, int HW, int C, const float* w, const float* b,
                                 float* logits, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  RBU_CHECK_ARG(VIEW_OK(x, ld) && w && b && logits && C >= 8 && C % 8 == 0 && P > 0 && HW > 0 && P % HW == 0,
                "rbu_head2_forward: bad arguments");
  const int tpp = pick_tpp2(C);
  long blocks = (P + NT / tpp - 1) / (NT / tpp);
  const long cap = (long)rbu_num_sms() * 16;
  if (blocks > cap) blocks = cap;
#define LAUNCH(T) head2_fwd_kernel<T><<<(unsigned)blocks, NT, 0, st>>>((const bf16*)x, ld, P, HW, C, w, b, logits)
  if (tpp == 4) LAUNCH(4); else if (tpp == 8) LAUNCH(8); else if (tpp == 16) LAUNCH(16); else LAUNCH(32);
#undef LAUNCH
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" size_t rbu_head2_backward_workspace_bytes(int64_t P, int C) {
  const int rows = NT / (C >> 3);
  long blocks = (long)rbu_num_sms() * 8;
  const long maxb = (P + rows * 4 - 1) / (rows * 4);
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  return (size_t)blocks * (2 * C + 8) * sizeof(float);
}

extern "C" int rbu_head2_backward(const float* dlogits, const void* x, int64_t x_ld, void* dx, int64_t dx_ld, int64_t P,
                                  int HW, int C, const float* w, float* dw, float* db, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  RBU_CHECK_ARG(dlogits && VIEW_OK(x, x_ld) && VIEW_OK(dx, dx_ld) && w && dw && db && C >= 8 && C <= 256 &&
                    (C & (C - 1)) == 0 && P > 0 && HW > 0, "rbu_head2_backward: bad arguments");
  RBU_CHECK_ARG(workspace && workspace_bytes >= rbu_head2_backward_workspace_bytes(P, C), "rbu_head2_backward: workspace too small");
  const int blocks = (int)(rbu_head2_backward_workspace_bytes(P, C) / ((2 * C + 8) * sizeof(float)));
  const int ppb = (int)((P + blocks - 1) / blocks);
  const int nb = (int)((P + ppb - 1) / ppb);
  head2_bwd_kernel<<<nb, NT, 0, st>>>(dlogits, (const bf16*)x, x_ld, (bf16*)dx, dx_ld, P, HW, C, w, ppb, (float*)workspace);
  RBU_CHECK_LAUNCH();
  colsum2_kernel<<<rbu_cdiv(2 * C, 32), 256, 0, st>>>((const float*)workspace, nb, 2 * C, 2 * C + 8, dw);
  RBU_CHECK_LAUNCH();
  colsum2_kernel<<<1, 256, 0, st>>>((const float*)workspace + 2 * C, nb, 2, 2 * C + 8, db);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" size_t rbu_ce2_workspace_bytes(int B, int64_t HW) {
  long nblk = (HW + 256 * 8 - 1) / (256 * 8);
  if (nblk > 64) nblk = 64;
  if (nblk < 1) nblk = 1;
  return (size_t)B * nblk * sizeof(float);
}

extern "C" int rbu_ce2_forward(const float* logits, const int64_t* target, int B, int64_t HW, void* workspace,
                               size_t workspace_bytes, float* loss_out, int64_t* counts, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  RBU_CHECK_ARG(logits && target && loss_out && counts && B > 0 && B <= 65535 && HW > 0, "rbu_ce2_forward: bad arguments");
  RBU_CHECK_ARG(workspace && workspace_bytes >= rbu_ce2_workspace_bytes(B, HW), "rbu_ce2_forward: workspace too small");
  const int nblk = (int)(rbu_ce2_workspace_bytes(B, HW) / sizeof(float) / B);
  RBU_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * 4 * sizeof(int64_t), st));
  ce2_partial_kernel<<<dim3(nblk, B), 256, 0, st>>>(logits, (const long long*)target, HW, (float*)workspace,
                                                    (unsigned long long*)counts);
  RBU_CHECK_LAUNCH();
  ce2_finalize_kernel<<<1, 256, 0, st>>>((const float*)workspace, B * nblk, (double)B * (double)HW, HW, B, loss_out,
                                         (unsigned long long*)counts);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_ce2_backward(const float* logits, const int64_t* target, int B, int64_t HW, const float* grad_out,
                                float* dlogits, void* stream_) {
  RBU_CHECK_ARG(logits && target && dlogits && B > 0 && HW > 0, "rbu_ce2_backward: bad arguments");
  const long total = (long)B * HW;
  long blocks = (total + 255) / 256;
  const long cap = (long)rbu_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  ce2_backward_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(logits, (const long long*)target, B, HW, grad_out,
                                                                          dlogits);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}
