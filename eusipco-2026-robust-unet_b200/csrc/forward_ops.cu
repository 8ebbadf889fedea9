// Bandwidth-bound forward kernels of the Robust U-Net blocks (K2-K8 of SURVEY.md §2.1): BatchNorm
// statistics / finalize / apply(+ReLU+Dropout2d scale), ChannelAttention gate, SpatialAttention
// reduce + 7x7 gate, residual output, 2x2 max-pool, AttentionGate psi / apply, stem im2col, `outc` head.
// All activations are NHWC bf16 views (16-byte vector accesses, 8 channels per thread); statistics and
// gates are fp32; every reduction is two-stage with a fixed order (deterministic, no float atomics).
#include "rbu_common.cuh"

namespace {

constexpr int NT = 256;

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics (nn.BatchNorm2d train mode, Main_Final.py:158,160,173,127,132,210) and, with
// pool=1, the per-(n,c) mean / max / min (+ first-index arg) that ChannelAttention's AdaptiveAvg/MaxPool
// (Main_Final.py:87-88,98-99) need — computed on the raw conv output; BN is a per-channel affine map.
// grid (chunks, N); thread = (channel group of 8, pixel row)
// ------------------------------------------------------------------------------------------------
template <int POOL>
__global__ void __launch_bounds__(NT)
bn_stats_kernel(const bf16* __restrict__ x, long ld, int HW, int C, int chunk_px, float* __restrict__ part_f) {
  const int G = C >> 3;
  const int rows = NT / G;
  const int cg = threadIdx.x % G;
  const int row = threadIdx.x / G;
  const int n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int p0 = chunk * chunk_px;
  const int p1 = min(p0 + chunk_px, HW);
  float sum[8], sq[8], mx[8], mn[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sum[e] = 0.f; sq[e] = 0.f; mx[e] = -INFINITY; mn[e] = INFINITY; }
  if (row < rows) {
    const bf16* base = x + (long)n * HW * ld + cg * 8;
    constexpr int U = 4;   // independent 16-byte loads in flight per thread
    for (int p = p0 + row; p < p1; p += rows * U) {
      bf16x8 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (p + u * rows < p1) raw[u] = ld_bf16x8_stream(base + (long)(p + u * rows) * ld);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (p + u * rows < p1) {
          float v[8];
          unpack8(raw[u], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            sum[e] += v[e];
            sq[e] += v[e] * v[e];
            if (POOL) { mx[e] = fmaxf(mx[e], v[e]); mn[e] = fminf(mn[e], v[e]); }
          }
        }
      }
    }
  }
  __shared__ float sv[NT][8];
  const long obase = ((long)n * chunks + chunk);
  // q = 0: sum, 1: sum of squares, 2: max, 3: min; rows combined in row order
  for (int q = 0; q < (POOL ? 4 : 2); ++q) {
#pragma unroll
    for (int e = 0; e < 8; ++e) sv[threadIdx.x][e] = q == 0 ? sum[e] : q == 1 ? sq[e] : q == 2 ? mx[e] : mn[e];
    __syncthreads();
    if (row == 0) {
      float a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = sv[cg][e];
      for (int r = 1; r < rows; ++r)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float t = sv[r * G + cg][e];
          a[e] = q < 2 ? a[e] + t : q == 2 ? fmaxf(a[e], t) : fminf(a[e], t);
        }
#pragma unroll
      for (int e = 0; e < 8; ++e) part_f[(obase * 4 + q) * C + cg * 8 + e] = a[e];
    }
    __syncthreads();
  }
}

// Stage 2 of the statistics: thread per (n, c) combines the chunk partials of image n in chunk order
// -> per-(n,c) sum / sumsq (double, workspace) and the pooled mean / max / min.
__global__ void bn_reduce_nc_kernel(const float* __restrict__ part_f, int chunks, int HW, int C, int pool,
                                    double* __restrict__ nsum /* [N][2][C] */, float* __restrict__ nc_mean,
                                    float* __restrict__ nc_max, float* __restrict__ nc_min) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  float bmx = -INFINITY, bmn = INFINITY;
  for (int k = 0; k < chunks; ++k) {
    const long o = (long)n * chunks + k;
    s += (double)part_f[(o * 4 + 0) * C + c];
    q += (double)part_f[(o * 4 + 1) * C + c];
    if (pool) {
      bmx = fmaxf(bmx, part_f[(o * 4 + 2) * C + c]);
      bmn = fminf(bmn, part_f[(o * 4 + 3) * C + c]);
    }
  }
  nsum[((long)n * 2 + 0) * C + c] = s;
  nsum[((long)n * 2 + 1) * C + c] = q;
  if (pool) {
    nc_mean[(long)n * C + c] = (float)(s / (double)HW);
    nc_max[(long)n * C + c] = bmx;
    nc_min[(long)n * C + c] = bmn;
  }
}

// Stage 2 for many chunks (the per-half-tile partials written by the 3x3 convolution's epilogue: up to 512 chunks per
// image): block = 32 channels x 8 parts of one image; part j combines chunks j, j+8, ... in order, the 8 parts are then
// combined in part order (fixed order -> deterministic).
__global__ void __launch_bounds__(256)
bn_reduce_nc_wide_kernel(const float* __restrict__ part_f, int chunks, int HW, int C, int pool,
                         double* __restrict__ nsum /* [N][2][C] */, float* __restrict__ nc_mean,
                         float* __restrict__ nc_max, float* __restrict__ nc_min) {
  __shared__ double shs[2][8][32];
  __shared__ float shm[2][8][32];
  const int cx = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int n = blockIdx.y;
  double s = 0.0, q = 0.0;
  float bmx = -INFINITY, bmn = INFINITY;
  if (c < C)
    for (int k = j; k < chunks; k += 8) {
      const long o = ((long)n * chunks + k) * 4;
      s += (double)part_f[(o + 0) * C + c];
      q += (double)part_f[(o + 1) * C + c];
      if (pool) {
        bmx = fmaxf(bmx, part_f[(o + 2) * C + c]);
        bmn = fminf(bmn, part_f[(o + 3) * C + c]);
      }
    }
  shs[0][j][cx] = s; shs[1][j][cx] = q; shm[0][j][cx] = bmx; shm[1][j][cx] = bmn;
  __syncthreads();
  if (j != 0 || c >= C) return;
  for (int r = 1; r < 8; ++r) {
    s += shs[0][r][cx]; q += shs[1][r][cx];
    bmx = fmaxf(bmx, shm[0][r][cx]); bmn = fminf(bmn, shm[1][r][cx]);
  }
  nsum[((long)n * 2 + 0) * C + c] = s;
  nsum[((long)n * 2 + 1) * C + c] = q;
  if (pool) {
    nc_mean[(long)n * C + c] = (float)(s / (double)HW);
    nc_max[(long)n * C + c] = bmx;
    nc_min[(long)n * C + c] = bmn;
  }
}

// Stage 3: block = 32 channels x 8 lanes; lane j adds images j, j+8, ... and the 8 lane sums are combined in lane
// order (fixed order -> deterministic).  BN affine (train: batch statistics + running update; eval: running stats).
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const double* __restrict__ nsum, int N, int HW, int C, int training, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                   float momentum, float eps, float* __restrict__ scale, float* __restrict__ shift,
                   float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  __shared__ double sh[2][8][32];
  const int cx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double s = 0.0, q = 0.0;
  if (training && c < C)
    for (int n = ly; n < N; n += 8) {
      s += nsum[((long)n * 2 + 0) * C + c];
      q += nsum[((long)n * 2 + 1) * C + c];
    }
  sh[0][ly][cx] = s;
  sh[1][ly][cx] = q;
  __syncthreads();
  if (ly != 0 || c >= C) return;
  float mean, var;
  if (training) {
    double tsum = 0.0, tsq = 0.0;
    for (int j = 0; j < 8; ++j) { tsum += sh[0][j][cx]; tsq += sh[1][j][cx]; }
    const double M = (double)N * (double)HW;
    const double m = tsum / M;
    double v = tsq / M - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      const float unbiased = M > 1.0 ? (float)(v * M / (M - 1.0)) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = 1.0f / sqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  if (mean_out) mean_out[c] = mean;
  if (rstd_out) rstd_out[c] = rstd;
}

// Streaming skeleton shared by the elementwise kernels: grid = (blocks, N); a block walks the HW*G 16-byte items of
// image n = blockIdx.y in steps of NT*U, every thread issuing its U independent 16-byte loads before the first use
// (memory-level parallelism), all index math in 32 bits with G = C/8 a power of two.  NT % G == 0, so the channel
// group of a thread never changes and its per-channel parameters live in registers.
constexpr int EW_U = 4;

// y = [relu](scale[c]*x + shift[c]) * drop[n,c]     (BN apply + ReLU + Dropout2d, Main_Final.py:182-184,220-221)
__global__ void __launch_bounds__(NT)
affine_act_kernel(const bf16* __restrict__ x, long x_ld, bf16* __restrict__ y, long y_ld, int HW, int C, int lg,
                  const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ drop,
                  int relu) {
  const int n = blockIdx.y, G = 1 << lg, items = HW << lg;
  const int cg = threadIdx.x & (G - 1);
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float d = drop ? drop[(long)n * C + cg * 8 + e] : 1.f;   // d >= 0: relu(t) * d == relu(t * d)
    sc[e] = scale[cg * 8 + e] * d;
    sh[e] = shift[cg * 8 + e] * d;
  }
  const bf16* xb = x + (long)n * HW * x_ld + cg * 8;
  bf16* yb = y + (long)n * HW * y_ld + cg * 8;
  for (int base = blockIdx.x * (NT * EW_U) + threadIdx.x; base < items; base += gridDim.x * (NT * EW_U)) {
    bf16x8 raw[EW_U];
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const int i = base + u * NT;
      if (i < items) raw[u] = ld_bf16x8_stream(xb + (long)(i >> lg) * x_ld);
    }
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const int i = base + u * NT;
      if (i < items) {
        float v[8];
        unpack8(raw[u], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float t = sc[e] * v[e] + sh[e];
          v[e] = relu ? fmaxf(t, 0.f) : t;
        }
        st_bf16x8(yb + (long)(i >> lg) * y_ld, pack8(v));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// ChannelAttention gate (Main_Final.py:97-101): g = sigmoid(V2 relu(V1 u_avg) + V2 relu(V1 u_max)),
// u = BN2 affine of the pooled raw conv2 output.  One block per image.  Also emits the fused per-(n,c)
// affine  A2g = scale*g, B2g = shift*g  so that  CA(BN2(y2)) = A2g*y2 + B2g.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
ca_gate_kernel(const float* __restrict__ nc_mean, const float* __restrict__ nc_max, const float* __restrict__ nc_min,
               const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ V1,
               const float* __restrict__ V2, int C, int Ch, float* __restrict__ g_out, float* __restrict__ A2g,
               float* __restrict__ B2g, float* __restrict__ u_avg_out, float* __restrict__ u_max_out,
               float* __restrict__ h_avg_out, float* __restrict__ h_max_out, float* __restrict__ tv_out,
               int* __restrict__ arg_init) {
  extern __shared__ float sm[];
  float* u_avg = sm;            // [C]
  float* u_max = sm + C;        // [C]
  float* h_avg = sm + 2 * C;    // [Ch]
  float* h_max = h_avg + Ch;    // [Ch]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += NT) {
    const float sc = scale[c], sh = shift[c];
    const float ua = sc * nc_mean[(long)n * C + c] + sh;
    // AdaptiveMaxPool of b = sc*y2 + sh picks max(y2) when sc >= 0, min(y2) otherwise; tv is that raw value and
    // arg_init resets the slot in which rbu_sa_reduce records the FIRST pixel attaining it (gradient routing)
    const float tv = sc >= 0.f ? nc_max[(long)n * C + c] : nc_min[(long)n * C + c];
    const float um = sc * tv + sh;
    tv_out[(long)n * C + c] = tv;
    arg_init[(long)n * C + c] = 0x7fffffff;
    u_avg[c] = ua;
    u_max[c] = um;
    u_avg_out[(long)n * C + c] = ua;
    u_max_out[(long)n * C + c] = um;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < Ch; j += NT / 32) {
    float a = 0.f, m = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float w = V1[(long)j * C + c];
      a += w * u_avg[c];
      m += w * u_max[c];
    }
    a = warp_sum(a);
    m = warp_sum(m);
    if (lane == 0) {
      h_avg[j] = a;
      h_max[j] = m;
      h_avg_out[(long)n * Ch + j] = a;
      h_max_out[(long)n * Ch + j] = m;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += NT) {
    float o = 0.f;
    for (int j = 0; j < Ch; ++j) o += V2[(long)c * Ch + j] * (fmaxf(h_avg[j], 0.f) + fmaxf(h_max[j], 0.f));
    const float g = sigmoidf_acc(o);
    g_out[(long)n * C + c] = g;
    A2g[(long)n * C + c] = scale[c] * g;
    B2g[(long)n * C + c] = shift[c] * g;
  }
}

// ------------------------------------------------------------------------------------------------
// SpatialAttention reduce (Main_Final.py:113-115): per pixel mean_c / max_c (+ first argmax) of
// c = A2g[n,c]*y2 + B2g[n,c].  TPP lanes cooperate on one pixel.
// ------------------------------------------------------------------------------------------------
// Also records, per (n,c), the FIRST pixel at which the raw conv output equals the pooled extreme `tv` chosen by
// ChannelAttention's max-pool (atomicMin on the pixel index; one hit per (n,c) in general) -- the arg-max that
// adaptive_max_pool2d_backward routes the gradient to.  grid = (blocks, N); lane li of a TPP-lane group owns the
// channel groups li, li+TPP, ... (K of them), whose per-(n,c) coefficients stay in registers.
template <int TPP, int K, bool ARG>   // ARG: also the per-pixel arg-max channel and the ChannelAttention arg-max pixel (training)
__global__ void __launch_bounds__(NT)
sa_reduce_kernel(const bf16* __restrict__ y2, long ld, int HW, int C, const float* __restrict__ A2g,
                 const float* __restrict__ B2g, const float* __restrict__ tv, int* __restrict__ nc_arg,
                 float2* __restrict__ s_out, int* __restrict__ amax_out) {
  constexpr int SLOTS = NT / TPP;
  constexpr int U = K == 1 ? 4 : (K == 2 ? 2 : 1);
  const int G = C >> 3;
  const int n = blockIdx.y;
  const int li = threadIdx.x % TPP;
  const int slot = threadIdx.x / TPP;
  float a[K][8], b[K][8];
  // The ChannelAttention targets are kept packed as bf16 (they are maxima / minima of stored bf16 values): one pair-wise
  // bf16 comparison per channel pair (HSETP2: float semantics, +0 == -0) says "no hit in this vector" in four
  // instructions; the exact per-element comparison against the fp32 targets runs only behind it.
  bf16x8 tp[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int cg = li + k * TPP;
    float t[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const long o = (long)n * C + cg * 8 + e;
      a[k][e] = cg < G ? A2g[o] : 0.f;
      b[k][e] = cg < G ? B2g[o] : 0.f;
      t[e] = (ARG && cg < G) ? tv[o] : 0.f;
    }
    tp[k] = pack8(t);
  }
  const float invC = 1.f / (float)C;
  // this thread's pixel pointer walks by a constant stride (no 64-bit multiply per load)
  const long step_u = (long)SLOTS * ld;
  const long step_it = (long)gridDim.x * (SLOTS * U) * ld;
  const bf16* ptr = y2 + ((long)n * HW + (long)blockIdx.x * (SLOTS * U) + slot) * ld + li * 8;
  const bool lane_live = li < G;
  float2* __restrict__ const s_n = s_out + (long)n * HW;             // this image's rows of the outputs
  int* __restrict__ const amax_n = ARG ? amax_out + (long)n * HW : nullptr;
  int* __restrict__ const arg_n = ARG ? nc_arg + (long)n * C : nullptr;
  const float* __restrict__ const tv_n = ARG ? tv + (long)n * C : nullptr;
  for (int pb = blockIdx.x * (SLOTS * U); pb < HW; pb += gridDim.x * (SLOTS * U), ptr += step_it) {
    bf16x8 raw[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = pb + u * SLOTS + slot;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (pp < HW && (K > 1 || lane_live)) raw[u][k] = ld_bf16x8(ptr + u * step_u + k * (TPP * 8));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = pb + u * SLOTS + slot;
      float sum = 0.f, best = -INFINITY;
      int bi = 0x7fffffff;
      if (pp < HW) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int cg = li + k * TPP;
          if (K > 1 || lane_live) {
            float v[8];
            unpack8(raw[u][k], v);
            if (ARG) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float c = a[k][e] * v[e] + b[k][e];
                sum += c;
                if (c > best) { best = c; bi = cg * 8 + e; }
              }
              const bool miss = (__hne2_mask(raw[u][k].v[0], tp[k].v[0]) & __hne2_mask(raw[u][k].v[1], tp[k].v[1]) &
                                 __hne2_mask(raw[u][k].v[2], tp[k].v[2]) & __hne2_mask(raw[u][k].v[3], tp[k].v[3])) == 0xffffffffu;
              if (!miss) {                  // rare (one pixel per (n,c) in general): the fp32 targets are re-read here
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (v[e] == tv_n[cg * 8 + e]) atomicMin(arg_n + cg * 8 + e, pp);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float c = a[k][e] * v[e] + b[k][e];
                sum += c;
                best = fmaxf(best, c);
              }
            }
          }
        }
      }
#pragma unroll
      for (int o = TPP / 2; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        if (ARG) {
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        } else {
          best = fmaxf(best, ov);
        }
      }
      if (li == 0 && pp < HW) {
        s_n[pp] = make_float2(sum * invC, best);
        if (ARG) amax_n[pp] = bi;
      }
    }
  }
}

// Training form for C >= 64: the per-(image, channel) coefficients and the packed arg-max targets live in shared memory (a
// block works on one image), so a lane can own FOUR channel groups without 100 registers of coefficients: TPP = C/32 lanes
// per pixel instead of C/8 -- two fewer combination rounds per pixel and the rounds, the guards and the store are paid once
// per four vectors.  Same arithmetic and visiting order as sa_reduce_kernel (channels ascending within a lane, strict >;
// across lanes the lower channel wins a tie).
template <int TPP>
__global__ void __launch_bounds__(NT, 3)
sa_reduce_arg_kernel(const bf16* __restrict__ y2, long ld, int HW, int C, const float* __restrict__ A2g,
                     const float* __restrict__ B2g, const float* __restrict__ tv, int* __restrict__ nc_arg,
                     float2* __restrict__ s_out, int* __restrict__ amax_out) {
  constexpr int K = 4, U = 2, SLOTS = NT / TPP;
  extern __shared__ float4 sa_smem[];
  float* const sa = reinterpret_cast<float*>(sa_smem);       // [C] A2g[n]
  float* const sb = sa + C;                                  // [C] B2g[n]
  bf16* const st = reinterpret_cast<bf16*>(sb + C);          // [C] tv[n], exact in bf16 (extremes of stored bf16 values)
  const int n = blockIdx.y;
  const float* __restrict__ const tv_n = tv + (long)n * C;
  for (int c = threadIdx.x; c < C; c += NT) {
    sa[c] = A2g[(long)n * C + c];
    sb[c] = B2g[(long)n * C + c];
    st[c] = __float2bfloat16_rn(tv_n[c]);
  }
  __syncthreads();
  const int li = threadIdx.x % TPP;
  const int slot = threadIdx.x / TPP;
  const float invC = 1.f / (float)C;
  const long step_u = (long)SLOTS * ld;
  const long step_it = (long)gridDim.x * (SLOTS * U) * ld;
  const bf16* ptr = y2 + ((long)n * HW + (long)blockIdx.x * (SLOTS * U) + slot) * ld + li * 8;
  float2* __restrict__ const s_n = s_out + (long)n * HW;
  int* __restrict__ const amax_n = amax_out + (long)n * HW;
  int* __restrict__ const arg_n = nc_arg + (long)n * C;
  for (int pb = blockIdx.x * (SLOTS * U); pb < HW; pb += gridDim.x * (SLOTS * U), ptr += step_it) {
    bf16x8 raw[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = pb + u * SLOTS + slot;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (pp < HW) raw[u][k] = ld_bf16x8(ptr + u * step_u + k * (TPP * 8));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = pb + u * SLOTS + slot;
      float sum = 0.f, best = -INFINITY;
      int bi = 0x7fffffff;
      if (pp < HW) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int cg = li + k * TPP;
          const float4 a0 = *reinterpret_cast<const float4*>(sa + cg * 8), a1 = *reinterpret_cast<const float4*>(sa + cg * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(sb + cg * 8), b1 = *reinterpret_cast<const float4*>(sb + cg * 8 + 4);
          const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          const bf16x8 tp = *reinterpret_cast<const bf16x8*>(st + cg * 8);
          float v[8];
          unpack8(raw[u][k], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float c = a[e] * v[e] + b[e];
            sum += c;
            if (c > best) { best = c; bi = cg * 8 + e; }
          }
          const bool miss = (__hne2_mask(raw[u][k].v[0], tp.v[0]) & __hne2_mask(raw[u][k].v[1], tp.v[1]) &
                             __hne2_mask(raw[u][k].v[2], tp.v[2]) & __hne2_mask(raw[u][k].v[3], tp.v[3])) == 0xffffffffu;
          if (!miss) {                      // rare: the exact comparison against the fp32 targets
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (v[e] == tv_n[cg * 8 + e]) atomicMin(arg_n + cg * 8 + e, pp);
          }
        }
      }
#pragma unroll
      for (int o = TPP / 2; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (li == 0 && pp < HW) {
        s_n[pp] = make_float2(sum * invC, best);
        amax_n[pp] = bi;
      }
    }
  }
}

// g_s = sigmoid(conv7x7([s_avg, s_max]))  (Main_Final.py:109,116-117); weight layout [1][2][7][7].
// Shared-memory tiled stencil: a block computes a 32 x 8 tile of one image from its 38 x 14 halo (zero padded).
// A block computes a 32 x 16 tile from its 38 x 22 halo; a thread owns four horizontally adjacent pixels, so one 10-wide
// window of a halo row serves 4 x 7 taps (17 shared-memory loads per pixel instead of 49 + 98 weight broadcasts).
constexpr int SA_TX = 32, SA_TY = 16, SA_HX = SA_TX + 6, SA_HY = SA_TY + 6, SA_NT = (SA_TX / 4) * SA_TY;
__global__ void __launch_bounds__(SA_NT)
sa_gate_kernel(const float2* __restrict__ s, int H, int W, const float* __restrict__ k7, float* __restrict__ gs) {
  __shared__ float wk[98];
  __shared__ float2 tile[SA_HY][SA_HX];
  const int tid = threadIdx.x;
  if (tid < 98) wk[tid] = k7[tid];
  const int x0 = blockIdx.x * SA_TX, y0 = blockIdx.y * SA_TY;
  const long nb = (long)blockIdx.z * H * W;
  for (int i = tid; i < SA_HX * SA_HY; i += SA_NT) {
    const int hy = i / SA_HX, hx = i - hy * SA_HX;
    const int yy = y0 + hy - 3, xx = x0 + hx - 3;
    tile[hy][hx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(&s[nb + (long)yy * W + xx]) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int tx = (tid % (SA_TX / 4)) * 4, ty = tid / (SA_TX / 4);
  const int x = x0 + tx, y = y0 + ty;
  if (x >= W || y >= H) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    float2 win[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) win[j] = tile[ty + r][tx + j];
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const float w0 = wk[r * 7 + q], w1 = wk[49 + r * 7 + q];
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] += w0 * win[q + e].x + w1 * win[q + e].y;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (x + e < W) gs[nb + (long)y * W + x + e] = sigmoidf_acc(acc[e]);
}

// out = relu((A2g*y2 + B2g) * g_s + r),  r = As*ys + Bs (projection shortcut) or r = x (identity)
// (Main_Final.py:179,190-194)
__global__ void __launch_bounds__(NT)
rb_out_kernel(const bf16* __restrict__ y2, long y2_ld, const bf16* __restrict__ rsrc, long r_ld, bf16* __restrict__ out,
              long out_ld, int HW, int C, int lg, const float* __restrict__ A2g, const float* __restrict__ B2g,
              const float* __restrict__ gs, const float* __restrict__ As, const float* __restrict__ Bs) {
  const int n = blockIdx.y, G = 1 << lg, items = HW << lg;
  const int cg = threadIdx.x & (G - 1);
  float a2[8], b2[8], as[8], bs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = cg * 8 + e;
    a2[e] = A2g[(long)n * C + c];
    b2[e] = B2g[(long)n * C + c];
    as[e] = As ? As[c] : 1.f;
    bs[e] = As ? Bs[c] : 0.f;
  }
  const bf16* yb = y2 + (long)n * HW * y2_ld + cg * 8;
  const bf16* rb = rsrc + (long)n * HW * r_ld + cg * 8;
  bf16* ob = out + (long)n * HW * out_ld + cg * 8;
  const float* gb = gs + (long)n * HW;
  for (int base = blockIdx.x * (NT * EW_U) + threadIdx.x; base < items; base += gridDim.x * (NT * EW_U)) {
    bf16x8 ry[EW_U], rr[EW_U];
    float g[EW_U];
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const int i = base + u * NT;
      if (i < items) {
        const int pl = i >> lg;
        ry[u] = ld_bf16x8_stream(yb + (long)pl * y2_ld);
        rr[u] = ld_bf16x8_stream(rb + (long)pl * r_ld);
        g[u] = __ldg(gb + pl);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const int i = base + u * NT;
      if (i < items) {
        float v[8], r[8];
        unpack8(ry[u], v);
        unpack8(rr[u], r);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaxf((a2[e] * v[e] + b2[e]) * g[u] + (as[e] * r[e] + bs[e]), 0.f);
        st_bf16x8(ob + (long)(i >> lg) * out_ld, pack8(v));
      }
    }
  }
}

// 2x2 stride-2 max pool (nn.MaxPool2d(2), Main_Final.py:235,239,243,249)
__global__ void __launch_bounds__(NT)
maxpool_kernel(const bf16* __restrict__ x, long x_ld, bf16* __restrict__ y, long y_ld, int N, int Ho, int Wo, int C) {
  const int G = C >> 3;
  const long total = (long)N * Ho * Wo * G;
  const int W = 2 * Wo;
  for (long i = blockIdx.x * (long)NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
    const int cg = (int)(i % G);
    const long po = i / G;
    const int wo = (int)(po % Wo);
    const int ho = (int)((po / Wo) % Ho);
    const int n = (int)(po / ((long)Wo * Ho));
    const long pi = ((long)n * 2 * Ho + 2 * ho) * W + 2 * wo;
    float a[8], b[8], c[8], d[8];
    unpack8(ld_bf16x8(x + pi * x_ld + cg * 8), a);
    unpack8(ld_bf16x8(x + (pi + 1) * x_ld + cg * 8), b);
    unpack8(ld_bf16x8(x + (pi + W) * x_ld + cg * 8), c);
    unpack8(ld_bf16x8(x + (pi + W + 1) * x_ld + cg * 8), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = fmaxf(fmaxf(a[e], b[e]), fmaxf(c[e], d[e]));
    st_bf16x8(y + po * y_ld + cg * 8, pack8(a));
  }
}

// ------------------------------------------------------------------------------------------------
// AttentionGate (Main_Final.py:143-148)
//   psi stage : t = relu(Ag*yg+Bg + Ax*yx+Bx); q0 = w_psi . t + b_psi  -> q0[P], block partials (sum, sumsq)
//   finalize  : BatchNorm2d(1) over all pixels (scalar statistics)
//   apply     : psi = sigmoid(a*q0+b); out = skip * psi
// ------------------------------------------------------------------------------------------------
// TPP lanes cooperate on a pixel; lane li owns channel groups li, li+TPP, ... (K of them) whose per-channel
// coefficients (Ag, Ax, Bg+Bx, w_psi) stay in registers; U pixels per slot are loaded before the first use.
template <int TPP, int K>
__global__ void __launch_bounds__(NT)
ag_psi_kernel(const bf16* __restrict__ yg, long yg_ld, const bf16* __restrict__ yx, long yx_ld, long P, int F,
              const float* __restrict__ Ag, const float* __restrict__ Bg, const float* __restrict__ Ax,
              const float* __restrict__ Bx, const float* __restrict__ wpsi, const float* __restrict__ bpsi,
              float* __restrict__ q0, float* __restrict__ partials) {
  constexpr int SLOTS = NT / TPP;
  constexpr int U = K == 1 ? 4 : (K == 2 ? 2 : 1);
  const int G = F >> 3;
  const int li = threadIdx.x % TPP;
  const int slot = threadIdx.x / TPP;
  float ag[K][8], ax[K][8], bs[K][8], wp[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int cg = li + k * TPP;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cg * 8 + e;
      const bool ok = cg < G;
      ag[k][e] = ok ? Ag[c] : 0.f;
      ax[k][e] = ok ? Ax[c] : 0.f;
      bs[k][e] = ok ? Bg[c] + Bx[c] : 0.f;
      wp[k][e] = ok ? wpsi[c] : 0.f;
    }
  }
  const float bias = bpsi[0];
  float ls = 0.f, lq = 0.f;
  for (long pb = (long)blockIdx.x * (SLOTS * U); pb < P; pb += (long)gridDim.x * (SLOTS * U)) {
    bf16x8 ra[U][K], rb[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long pp = pb + u * SLOTS + slot;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (pp < P && li + k * TPP < G) {
          ra[u][k] = ld_bf16x8_stream(yg + pp * yg_ld + (li + k * TPP) * 8);
          rb[u][k] = ld_bf16x8_stream(yx + pp * yx_ld + (li + k * TPP) * 8);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long pp = pb + u * SLOTS + slot;
      float acc = 0.f;
      if (pp < P) {
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (li + k * TPP < G) {
            float a[8], b[8];
            unpack8(ra[u][k], a);
            unpack8(rb[u][k], b);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc += wp[k][e] * fmaxf(ag[k][e] * a[e] + ax[k][e] * b[e] + bs[k][e], 0.f);
          }
      }
#pragma unroll
      for (int o = TPP / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (li == 0 && pp < P) {
        const float q = acc + bias;
        q0[pp] = q;
        ls += q;
        lq += q * q;
      }
    }
  }
  __shared__ float r0[NT / 32], r1[NT / 32];
  ls = warp_sum(ls);
  lq = warp_sum(lq);
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = ls; r1[threadIdx.x >> 5] = lq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < NT / 32; ++w) { a += r0[w]; b += r1[w]; }
    partials[blockIdx.x * 2] = a;
    partials[blockIdx.x * 2 + 1] = b;
  }
}

// scalar BN finalize: stats[0..3] = scale, shift, mean, rstd.  One block of 256 threads: thread j adds partials
// j, j+256, ... in double, then the 256 lane sums are combined in lane order (fixed order -> deterministic).
__global__ void __launch_bounds__(256)
scalar_bn_finalize_kernel(const float* __restrict__ partials, int nblk, long M, int training,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                          float* __restrict__ stats) {
  __shared__ double sh[2][256];
  double s = 0.0, q = 0.0;
  if (training)
    for (int i = threadIdx.x; i < nblk; i += 256) { s += (double)partials[2 * i]; q += (double)partials[2 * i + 1]; }
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x != 0) return;
  float mean, var;
  if (training) {
    s = 0.0; q = 0.0;
    for (int j = 0; j < 256; ++j) { s += sh[0][j]; q += sh[1][j]; }
    const double m = s / (double)M;
    double v = q / (double)M - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) {
      running_mean[0] = (1.f - momentum) * running_mean[0] + momentum * mean;
      running_var[0] = (1.f - momentum) * running_var[0] + momentum * (M > 1 ? (float)(v * (double)M / (double)(M - 1)) : var);
    }
  } else {
    mean = running_mean[0];
    var = running_var[0];
  }
  const float rstd = 1.0f / sqrtf(var + eps);
  stats[0] = gamma[0] * rstd;
  stats[1] = beta[0] - mean * gamma[0] * rstd;
  stats[2] = mean;
  stats[3] = rstd;
}

__global__ void __launch_bounds__(NT)
ag_apply_kernel(const bf16* __restrict__ skip, long s_ld, bf16* __restrict__ out, long o_ld, long P, int lg,
                const float* __restrict__ q0, const float* __restrict__ stats, float* __restrict__ psi_out) {
  const int G = 1 << lg;
  const long items = P << lg;
  const int cg = threadIdx.x & (G - 1);
  const float a = stats[0], b = stats[1];
  const bf16* sb = skip + cg * 8;
  bf16* ob = out + cg * 8;
  for (long base = (long)blockIdx.x * (NT * EW_U) + threadIdx.x; base < items; base += (long)gridDim.x * (NT * EW_U)) {
    bf16x8 raw[EW_U];
    float q[EW_U];
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const long i = base + u * NT;
      if (i < items) {
        raw[u] = ld_bf16x8(sb + (i >> lg) * s_ld);
        q[u] = __ldg(q0 + (i >> lg));
      }
    }
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const long i = base + u * NT;
      if (i < items) {
        const float psi = sigmoidf_acc(a * q[u] + b);
        if (cg == 0) psi_out[i >> lg] = psi;
        float v[8];
        unpack8(raw[u], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= psi;
        st_bf16x8(ob + (i >> lg) * o_ld, pack8(v));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stem: fp32 NCHW image -> bf16 3x3 patches [P][Kp], k = tap*nc + c (zero padded), so that inc.conv1
// (3x3, Cin=3|4) and inc.shortcut (1x1) run as ONE tensor-core GEMM with K = Kp (Main_Final.py:157,172,233).
// ------------------------------------------------------------------------------------------------
// Generic channel count: one thread per 16-byte piece of a patch row, reads through L1.
__global__ void __launch_bounds__(NT)
stem_im2col_any_kernel(const float* __restrict__ x, int N, int nc, int H, int W, int Kp, bf16* __restrict__ out) {
  const int KG = Kp >> 3;
  const long total = (long)N * H * W * KG;
  const int HW = H * W;
  for (long i = blockIdx.x * (long)NT + threadIdx.x; i < total; i += (long)gridDim.x * NT) {
    const int kg = (int)(i % KG);
    const long p = i / KG;
    const int n = (int)(p / HW);
    const int pl = (int)(p - (long)n * HW);
    const int h = pl / W, w = pl - h * W;
    const float* xn = x + (long)n * nc * HW;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kg * 8 + e;
      float t = 0.f;
      if (k < 9 * nc) {
        const int tap = k / nc, c = k - tap * nc;
        const int dy = tap / 3;
        const int hh = h + dy - 1, ww = w + (tap - dy * 3) - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) t = __ldg(xn + (long)c * HW + hh * W + ww);
      }
      v[e] = t;
    }
    st_bf16x8(out + i * 8, pack8(v));
  }
}

// NC = 3 / 4 input channels (the reference's models): a block stages the three input rows of a run of pixels in shared
// memory (coalesced reads, zero padding resolved once) and every thread assembles one 16-byte piece of a patch row with
// compile-time tap / channel offsets; a warp writes 512 contiguous bytes.
template <int NC, int KG>
__global__ void __launch_bounds__(NT)
stem_im2col_kernel(const float* __restrict__ x, int N, int H, int W, bf16* __restrict__ out) {
  constexpr int TP = NT / KG;            // pixels per tile row
  constexpr int TR = 4;                  // output rows per tile (TR + 2 input rows staged: 1.5 reads per output row)
  constexpr int TWD = TP + 2;
  __shared__ float t[NC][TR + 2][TWD];
  const int tiles_w = (W + TP - 1) / TP, tiles_h = (H + TR - 1) / TR;
  const long tiles = (long)N * tiles_h * tiles_w;
  const int px = threadIdx.x / KG, kg = threadIdx.x - px * KG;
  for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w);
    const long nh = tile / tiles_w;
    const int h0 = (int)(nh % tiles_h) * TR;
    const int n = (int)(nh / tiles_h);
    const int w0 = tw * TP;
    __syncthreads();
    for (int i = threadIdx.x; i < NC * (TR + 2) * TWD; i += NT) {
      const int c = i / ((TR + 2) * TWD), rem = i - c * (TR + 2) * TWD;
      const int r = rem / TWD, j = rem - r * TWD;
      const int hh = h0 + r - 1, ww = w0 + j - 1;
      (&t[0][0][0])[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + (((long)n * NC + c) * H + hh) * W + ww) : 0.f;
    }
    __syncthreads();
    const int w = w0 + px;
    if (px < TP && w < W) {
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        if (h0 + r < H) {
          float v[8];
#pragma unroll
          for (int q = 0; q < KG; ++q) {
            if (kg == q) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int k = q * 8 + e;               // compile-time after unrolling
                v[e] = k < 9 * NC ? t[k % NC][r + (k / NC) / 3][px + (k / NC) % 3] : 0.f;
              }
            }
          }
          st_bf16x8(out + (((long)n * H + h0 + r) * W + w) * (KG * 8) + kg * 8, pack8(v));
        }
      }
    }
  }
}

// outc: probs = sigmoid(w . x + b)   (1x1 conv C->1 + Sigmoid, fp32; Main_Final.py:274-277,321)
template <int TPP>
__global__ void __launch_bounds__(NT)
head_fwd_kernel(const bf16* __restrict__ x, long ld, long P, int C, const float* __restrict__ w,
                const float* __restrict__ b, float* __restrict__ probs, float* __restrict__ logits) {
  const int G = C >> 3;
  const int li = threadIdx.x % TPP;
  const int slot = threadIdx.x / TPP;
  constexpr int SLOTS = NT / TPP;
  if (G <= TPP) {
    // one channel group per lane: weights in registers, four pixels (independent 16-byte loads) per iteration
    constexpr int U = 4;
    float wr[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) wr[e] = li < G ? w[li * 8 + e] : 0.f;
    const float b0 = b[0];
    for (long p0 = blockIdx.x * (long)(SLOTS * U) + slot; p0 < P; p0 += (long)gridDim.x * (SLOTS * U)) {
      bf16x8 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long pp = p0 + u * SLOTS;
        if (pp < P && li < G) raw[u] = ld_bf16x8_stream(x + pp * ld + li * 8);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long pp = p0 + u * SLOTS;
        float acc = 0.f;
        if (pp < P && li < G) {
          float v[8];
          unpack8(raw[u], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc += wr[e] * v[e];
        }
#pragma unroll
        for (int o = TPP / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (li == 0 && pp < P) {
          const float z = acc + b0;
          probs[pp] = sigmoidf_acc(z);
          if (logits) logits[pp] = z;
        }
      }
    }
    return;
  }
  for (long p = blockIdx.x * (long)SLOTS + slot; p < P; p += (long)gridDim.x * SLOTS) {
    float acc = 0.f;
    for (int cg = li; cg < G; cg += TPP) {
      float v[8];
      unpack8(ld_bf16x8_stream(x + p * ld + cg * 8), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += w[cg * 8 + e] * v[e];
    }
#pragma unroll
    for (int o = TPP / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (li == 0) {
      const float z = acc + b[0];
      probs[p] = sigmoidf_acc(z);
      if (logits) logits[p] = z;
    }
  }
}

int grid_for(long items, int per_block) {
  long b = (items + per_block - 1) / per_block;
  const long cap = (long)rbu_num_sms() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int ilog2(int v) {
  int l = 0;
  while ((1 << (l + 1)) <= v) ++l;
  return l;
}

// grid.x for the streaming skeleton: enough blocks for ~32 resident warps per SM over >= 4 waves, at most one
// block per NT*EW_U items
int ew_blocks(long items_per_image, int images) { return rbu_stream_blocks(items_per_image, NT * EW_U, images); }

int pick_tpp(int C) {
  const int G = C >> 3;
  int t = 4;
  while (t * 2 <= G && t < 32) t <<= 1;
  return t;
}

int stats_chunks(int N, int HW, int C) {
  const int rows = NT / (C >> 3);
  long want = ((long)rbu_num_sms() * 16 + N - 1) / N;         // ~16 blocks per SM in total: small tail wave
  long maxc = (HW + (long)rows * 8 - 1) / ((long)rows * 8);   // at least ~8 pixels per thread
  if (maxc < 1) maxc = 1;
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace

#define VIEW_OK(ptr, ld) ((ptr) != nullptr && ((uintptr_t)(ptr) & 15) == 0 && (ld) % 8 == 0)
#define EW_CH_OK(C) ((C) >= 8 && (C) <= 2048 && ((C) & ((C) - 1)) == 0)

extern "C" size_t rbu_bn_stats_workspace_bytes(int N, int HW, int C) {
  if (C < 8 || C % 8 || C > 2048) return 0;
  return (size_t)N * stats_chunks(N, HW, C) * 4 * C * sizeof(float) + (size_t)N * 2 * C * sizeof(double) + 16;
}

extern "C" int rbu_bn_stats(const void* x, int64_t ld, int N, int HW, int C, int pool, int training,
                            const float* gamma, const float* beta, float* running_mean, float* running_var,
                            float momentum, float eps, float* scale, float* shift, float* mean_out, float* rstd_out,
                            float* nc_mean, float* nc_max, float* nc_min, void* workspace, size_t workspace_bytes,
                            void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBU_CHECK_ARG(VIEW_OK(x, ld), "rbu_bn_stats: misaligned view");
  RBU_CHECK_ARG(N > 0 && HW > 0 && C >= 8 && C % 8 == 0 && C <= 2048, "rbu_bn_stats: unsupported shape N=%d HW=%d C=%d", N, HW, C);
  RBU_CHECK_ARG(gamma && beta && scale && shift, "rbu_bn_stats: null parameter pointer");
  RBU_CHECK_ARG(training || (running_mean && running_var), "rbu_bn_stats: eval mode needs running statistics");
  RBU_CHECK_ARG(!pool || (nc_mean && nc_max && nc_min), "rbu_bn_stats: pool outputs missing");
  if (!training && !pool) {  // pure eval affine: no pass over the data
    bn_finalize_kernel<<<rbu_cdiv(C, 32), 256, 0, stream>>>(nullptr, 0, HW, C, 0, gamma, beta, running_mean, running_var,
                                                             momentum, eps, scale, shift, mean_out, rstd_out);
    RBU_CHECK_LAUNCH();
    return RBU_OK;
  }
  RBU_CHECK_ARG(workspace && workspace_bytes >= rbu_bn_stats_workspace_bytes(N, HW, C), "rbu_bn_stats: workspace too small");
  RBU_CHECK_ARG(N <= 65535, "rbu_bn_stats: batch too large");
  const int chunks = stats_chunks(N, HW, C);
  const int chunk_px = rbu_cdiv(HW, chunks);
  float* part_f = (float*)workspace;
  if (pool)
    bn_stats_kernel<1><<<dim3(chunks, N), NT, 0, stream>>>((const bf16*)x, ld, HW, C, chunk_px, part_f);
  else
    bn_stats_kernel<0><<<dim3(chunks, N), NT, 0, stream>>>((const bf16*)x, ld, HW, C, chunk_px, part_f);
  RBU_CHECK_LAUNCH();
  double* nsum = (double*)(((uintptr_t)(part_f + (size_t)N * chunks * 4 * C) + 15) & ~(uintptr_t)15);
  bn_reduce_nc_kernel<<<dim3(rbu_cdiv(C, 128), N), 128, 0, stream>>>(part_f, chunks, HW, C, pool, nsum, nc_mean, nc_max,
                                                                      nc_min);
  RBU_CHECK_LAUNCH();
  bn_finalize_kernel<<<rbu_cdiv(C, 32), 256, 0, stream>>>(nsum, N, HW, C, training, gamma, beta, running_mean, running_var,
                                                           momentum, eps, scale, shift, mean_out, rstd_out);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

// BatchNorm statistics (+ pooled per-image statistics) from partials that another kernel already produced in the
// [N][chunks][4][C] layout of rbu_bn_stats' first stage -- the 3x3 convolution epilogue (rbu_conv_gemm tile_stats):
// no pass over the activation tensor.  workspace: N*2*C doubles.
extern "C" int rbu_bn_stats_from_partials(const float* partials, int chunks, int N, int HW, int C, int pool, int training,
                                          const float* gamma, const float* beta, float* running_mean, float* running_var,
                                          float momentum, float eps, float* scale, float* shift, float* mean_out,
                                          float* rstd_out, float* nc_mean, float* nc_max, float* nc_min, void* workspace,
                                          size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBU_CHECK_ARG(partials && chunks > 0 && N > 0 && N <= 65535 && HW > 0 && C >= 8 && C <= 2048,
                "rbu_bn_stats_from_partials: bad shape");
  RBU_CHECK_ARG(gamma && beta && scale && shift, "rbu_bn_stats_from_partials: null parameter pointer");
  RBU_CHECK_ARG(training || (running_mean && running_var), "rbu_bn_stats_from_partials: eval mode needs running statistics");
  RBU_CHECK_ARG(!pool || (nc_mean && nc_max && nc_min), "rbu_bn_stats_from_partials: pool outputs missing");
  RBU_CHECK_ARG(workspace && ((uintptr_t)workspace & 15) == 0 && workspace_bytes >= (size_t)N * 2 * C * sizeof(double),
                "rbu_bn_stats_from_partials: workspace too small");
  double* nsum = (double*)workspace;
  bn_reduce_nc_wide_kernel<<<dim3(rbu_cdiv(C, 32), N), 256, 0, stream>>>(partials, chunks, HW, C, pool, nsum, nc_mean, nc_max,
                                                                        nc_min);
  RBU_CHECK_LAUNCH();
  bn_finalize_kernel<<<rbu_cdiv(C, 32), 256, 0, stream>>>(nsum, N, HW, C, training, gamma, beta, running_mean, running_var,
                                                           momentum, eps, scale, shift, mean_out, rstd_out);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_affine_act(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t P, int HW, int C,
                              const float* scale, const float* shift, const float* drop, int relu, void* stream_) {
  RBU_CHECK_ARG(VIEW_OK(x, x_ld) && VIEW_OK(y, y_ld) && scale && shift && EW_CH_OK(C) && P > 0 && HW > 0 && P % HW == 0 &&
                    P / HW <= 65535, "rbu_affine_act: bad arguments (C must be a power of two in [8, 2048])");
  const int lg = ilog2(C >> 3);
  affine_act_kernel<<<dim3(ew_blocks((long)HW << lg, (int)(P / HW)), (unsigned)(P / HW)), NT, 0, (cudaStream_t)stream_>>>(
      (const bf16*)x, x_ld, (bf16*)y, y_ld, HW, C, lg, scale, shift, drop, relu);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_ca_gate(const float* nc_mean, const float* nc_max, const float* nc_min, const float* scale,
                           const float* shift, const float* V1, const float* V2, int N, int C, int Ch, float* g,
                           float* A2g, float* B2g, float* u_avg, float* u_max, float* h_avg, float* h_max, float* tv,
                           int* nc_arg, void* stream_) {
  RBU_CHECK_ARG(nc_mean && nc_max && nc_min && scale && shift && V1 && V2 && g && A2g && B2g && u_avg && u_max &&
                    h_avg && h_max && tv && nc_arg, "rbu_ca_gate: null pointer");
  RBU_CHECK_ARG(N > 0 && C > 0 && Ch > 0 && (2 * C + 2 * Ch) * 4 <= 48 * 1024, "rbu_ca_gate: unsupported shape");
  ca_gate_kernel<<<N, NT, (2 * C + 2 * Ch) * sizeof(float), (cudaStream_t)stream_>>>(
      nc_mean, nc_max, nc_min, scale, shift, V1, V2, C, Ch, g, A2g, B2g, u_avg, u_max, h_avg, h_max, tv, nc_arg);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_sa_reduce(const void* y2, int64_t ld, int64_t P, int HW, int C, const float* A2g, const float* B2g,
                             const float* tv, int* nc_arg, float* s_out, int* amax_out, void* stream_) {
  RBU_CHECK_ARG(VIEW_OK(y2, ld) && A2g && B2g && s_out && EW_CH_OK(C) && HW > 0 && P > 0 && P % HW == 0 && P / HW <= 65535,
                "rbu_sa_reduce: bad arguments");
  const bool arg = amax_out != nullptr;     // inference passes NULL: no arg-max bookkeeping
  RBU_CHECK_ARG(!arg || (tv && nc_arg), "rbu_sa_reduce: the arg-max outputs need tv and nc_arg");
  cudaStream_t st = (cudaStream_t)stream_;
  const int G = C >> 3, N = (int)(P / HW);
  int tpp = G >= 32 ? 32 : (G < 4 ? 4 : G);
  int K = G > 32 ? G / 32 : 1;
  // Inference form: two channel groups per lane from 64 channels up, so that the per-pixel combination across lanes
  // (shuffles, stores) is paid once per two vectors (3.41 -> 3.11 ms at 32 x 1024^2, same-box A/B).  With the arg-max
  // bookkeeping the same layout needs 94 registers instead of 74, costs a resident block per SM and loses (0.83 / 0.79 ms).
  if (!arg && K == 1 && G >= 8) { tpp = G >= 32 ? 16 : G / 2; K = 2; }
  // Training form from 64 channels up: coefficients in shared memory, four channel groups per lane (same-box A/B against
  // the register form below, two alternations: 0.796 / 0.790 -> 0.712 / 0.700 ms per step; profiles/r02_l_ab_sa_smem.txt)
  if (arg && G >= 8 && G <= 128) {
    const int t4 = G / 4;                                               // lanes per pixel, four channel groups each
    const dim3 grid4((unsigned)rbu_stream_blocks(HW, NT / t4 * 2, N), (unsigned)N);
    const size_t smem = (size_t)C * 10;
#define LAUNCH4(T)                                                                                                           \
  sa_reduce_arg_kernel<T><<<grid4, NT, smem, st>>>((const bf16*)y2, ld, HW, C, A2g, B2g, tv, nc_arg, (float2*)s_out, amax_out)
    if (t4 == 2) LAUNCH4(2); else if (t4 == 4) LAUNCH4(4); else if (t4 == 8) LAUNCH4(8); else if (t4 == 16) LAUNCH4(16); else LAUNCH4(32);
#undef LAUNCH4
    RBU_CHECK_LAUNCH();
    return RBU_OK;
  }
  const int U = K == 1 ? 4 : (K == 2 ? 2 : 1);
  const int per_block = NT / tpp * U;
  const long blocks = rbu_stream_blocks(HW, per_block, N);
  const dim3 grid((unsigned)blocks, (unsigned)N);
#define LAUNCH(T, KK)                                                                                                        \
  do {                                                                                                                    \
    if (arg)                                                                                                              \
      sa_reduce_kernel<T, KK, true><<<grid, NT, 0, st>>>((const bf16*)y2, ld, HW, C, A2g, B2g, tv, nc_arg, (float2*)s_out, \
                                                         amax_out);                                                       \
    else                                                                                                                  \
      sa_reduce_kernel<T, KK, false><<<grid, NT, 0, st>>>((const bf16*)y2, ld, HW, C, A2g, B2g, tv, nc_arg, (float2*)s_out,\
                                                          amax_out);                                                      \
  } while (0)
  if (K == 1) { if (tpp == 4) LAUNCH(4, 1); else if (tpp == 8) LAUNCH(8, 1); else if (tpp == 16) LAUNCH(16, 1); else LAUNCH(32, 1); }
  else if (K == 2) { if (tpp == 4) LAUNCH(4, 2); else if (tpp == 8) LAUNCH(8, 2); else if (tpp == 16) LAUNCH(16, 2); else LAUNCH(32, 2); }
  else if (K == 4) LAUNCH(32, 4);
  else { RBU_CHECK_ARG(K == 8, "rbu_sa_reduce: unsupported channel count %d", C); LAUNCH(32, 8); }
#undef LAUNCH
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_sa_gate(const float* s, int N, int H, int W, const float* k7, float* gs, void* stream_) {
  RBU_CHECK_ARG(s && k7 && gs && N > 0 && H > 0 && W > 0, "rbu_sa_gate: bad arguments");
  RBU_CHECK_ARG(N <= 65535, "rbu_sa_gate: batch too large");
  sa_gate_kernel<<<dim3(rbu_cdiv(W, SA_TX), rbu_cdiv(H, SA_TY), N), SA_NT, 0, (cudaStream_t)stream_>>>(
      (const float2*)s, H, W, k7, gs);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_rb_out(const void* y2, int64_t y2_ld, const void* rsrc, int64_t r_ld, void* out, int64_t out_ld,
                          int64_t P, int HW, int C, const float* A2g, const float* B2g, const float* gs,
                          const float* As, const float* Bs, void* stream_) {
  RBU_CHECK_ARG(VIEW_OK(y2, y2_ld) && VIEW_OK(rsrc, r_ld) && VIEW_OK(out, out_ld) && A2g && B2g && gs && EW_CH_OK(C) &&
                    HW > 0 && P > 0 && P % HW == 0 && P / HW <= 65535, "rbu_rb_out: bad arguments");
  RBU_CHECK_ARG((As == nullptr) == (Bs == nullptr), "rbu_rb_out: As/Bs must both be set or both NULL");
  const int lg = ilog2(C >> 3);
  rb_out_kernel<<<dim3(ew_blocks((long)HW << lg, (int)(P / HW)), (unsigned)(P / HW)), NT, 0, (cudaStream_t)stream_>>>(
      (const bf16*)y2, y2_ld, (const bf16*)rsrc, r_ld, (bf16*)out, out_ld, HW, C, lg, A2g, B2g, gs, As, Bs);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_maxpool2x2(const void* x, int64_t x_ld, void* y, int64_t y_ld, int N, int Ho, int Wo, int C,
                              void* stream_) {
  RBU_CHECK_ARG(VIEW_OK(x, x_ld) && VIEW_OK(y, y_ld) && N > 0 && Ho > 0 && Wo > 0 && C % 8 == 0, "rbu_maxpool2x2: bad arguments");
  maxpool_kernel<<<grid_for((long)N * Ho * Wo * (C >> 3), NT * 2), NT, 0, (cudaStream_t)stream_>>>(
      (const bf16*)x, x_ld, (bf16*)y, y_ld, N, Ho, Wo, C);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

namespace {
void ag_psi_cfg(int F, int& tpp, int& K, int& per_block) {
  const int G = F >> 3;
  tpp = G >= 32 ? 32 : (G < 4 ? 4 : G);
  K = G > 32 ? G / 32 : 1;
  const int U = K == 1 ? 4 : (K == 2 ? 2 : 1);
  per_block = NT / tpp * U;
}
}  // namespace

extern "C" int rbu_ag_psi_blocks(int64_t P, int F) {
  int tpp, K, per_block;
  ag_psi_cfg(F, tpp, K, per_block);
  long b = (P + per_block - 1) / per_block;
  const long cap = (long)rbu_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int rbu_ag_psi(const void* yg, int64_t yg_ld, const void* yx, int64_t yx_ld, int64_t P, int F,
                          const float* Ag, const float* Bg, const float* Ax, const float* Bx, const float* wpsi,
                          const float* bpsi, int training, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, float momentum, float eps, float* q0, float* stats, float* partials,
                          void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  RBU_CHECK_ARG(VIEW_OK(yg, yg_ld) && VIEW_OK(yx, yx_ld) && Ag && Bg && Ax && Bx && wpsi && bpsi && q0 && stats &&
                    partials && gamma && beta && EW_CH_OK(F), "rbu_ag_psi: bad arguments");
  int tpp, K, per_block;
  ag_psi_cfg(F, tpp, K, per_block);
  const int grid = rbu_ag_psi_blocks(P, F);
#define LAUNCH(T, KK) ag_psi_kernel<T, KK><<<grid, NT, 0, st>>>((const bf16*)yg, yg_ld, (const bf16*)yx, yx_ld, P, F, Ag, Bg, Ax, Bx, wpsi, bpsi, q0, partials)
  if (K == 1) { if (tpp == 4) LAUNCH(4, 1); else if (tpp == 8) LAUNCH(8, 1); else if (tpp == 16) LAUNCH(16, 1); else LAUNCH(32, 1); }
  else if (K == 2) LAUNCH(32, 2);
  else { RBU_CHECK_ARG(K == 4, "rbu_ag_psi: unsupported channel count %d", F); LAUNCH(32, 4); }
#undef LAUNCH
  RBU_CHECK_LAUNCH();
  scalar_bn_finalize_kernel<<<1, 256, 0, st>>>(partials, grid, P, training, gamma, beta, running_mean, running_var,
                                              momentum, eps, stats);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_ag_apply(const void* skip, int64_t s_ld, void* out, int64_t o_ld, int64_t P, int C, const float* q0,
                            const float* stats, float* psi, void* stream_) {
  RBU_CHECK_ARG(VIEW_OK(skip, s_ld) && VIEW_OK(out, o_ld) && q0 && stats && psi && EW_CH_OK(C), "rbu_ag_apply: bad arguments");
  const int lg = ilog2(C >> 3);
  ag_apply_kernel<<<ew_blocks(P << lg, 1), NT, 0, (cudaStream_t)stream_>>>((const bf16*)skip, s_ld, (bf16*)out, o_ld, P, lg,
                                                                          q0, stats, psi);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_stem_im2col(const float* x, int N, int nc, int H, int W, int Kp, void* out, void* stream_) {
  RBU_CHECK_ARG(x && out && N > 0 && nc > 0 && H > 0 && W > 0 && Kp % 8 == 0 && Kp >= 9 * nc && ((uintptr_t)out & 15) == 0,
                "rbu_stem_im2col: bad arguments");
  cudaStream_t st = (cudaStream_t)stream_;
  if (nc == 3 && Kp == 32) {
    const long tiles = (long)N * rbu_cdiv(H, 4) * rbu_cdiv(W, NT / 4);
    stem_im2col_kernel<3, 4><<<(unsigned)(tiles < (long)rbu_num_sms() * 16 ? tiles : (long)rbu_num_sms() * 16), NT, 0, st>>>(
        x, N, H, W, (bf16*)out);
  } else if (nc == 4 && Kp == 40) {
    const long tiles = (long)N * rbu_cdiv(H, 4) * rbu_cdiv(W, NT / 5);
    stem_im2col_kernel<4, 5><<<(unsigned)(tiles < (long)rbu_num_sms() * 16 ? tiles : (long)rbu_num_sms() * 16), NT, 0, st>>>(
        x, N, H, W, (bf16*)out);
  } else {
    stem_im2col_any_kernel<<<grid_for((long)N * H * W * (Kp / 8), NT), NT, 0, st>>>(x, N, nc, H, W, Kp, (bf16*)out);
  }
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_head_forward(const void* x, int64_t ld, int64_t P, int C, const float* w, const float* b,
                                float* probs, float* logits, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  RBU_CHECK_ARG(VIEW_OK(x, ld) && w && b && probs && C >= 8 && C % 8 == 0, "rbu_head_forward: bad arguments");
  const int tpp = pick_tpp(C);
  const int grid = grid_for(P, NT / tpp * 4);
#define LAUNCH(T) head_fwd_kernel<T><<<grid, NT, 0, st>>>((const bf16*)x, ld, P, C, w, b, probs, logits)
  if (tpp == 4) LAUNCH(4); else if (tpp == 8) LAUNCH(8); else if (tpp == 16) LAUNCH(16); else LAUNCH(32);
#undef LAUNCH
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}
