// Weight re-packing: fp32 torch-layout parameters -> bf16 GEMM operands [Ncols][taps][K] (K contiguous),
// in the forward and data-gradient forms the implicit-GEMM kernel consumes, plus the SIMT reference
// convolution used by the parity tests (test-only device reference, never on the product path).
#include "rbu_common.cuh"

namespace {

// mode 0: conv fwd   src [Nn][K][T]      -> dst[n][t][k] = src[n][k][t]
// mode 1: conv dgrad src [K][Nn][T]      -> dst[n][t][k] = src[k][n][T-1-t]        (180-degree rotated taps)
// mode 2: convT fwd  src [K][Cout][4]    -> dst[q*Cout+co][0][k] = src[k][co][q]   (Nn = 4*Cout, T = 1)
// mode 3: convT dgrad src [Nn][K][4]     -> dst[n][q][k] = src[n][k][q]            (T = 4)
__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int Nn, int T, int K,
                                   int mode, int Cout) {
  const long total = (long)Nn * T * K;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int t = (int)((i / K) % T);
    const int n = (int)(i / ((long)K * T));
    long s;
    if (mode == 0) s = ((long)n * K + k) * T + t;
    else if (mode == 1) s = ((long)k * Nn + n) * T + (T - 1 - t);
    else if (mode == 2) { const int q = n / Cout, co = n % Cout; s = ((long)k * Cout + co) * 4 + q; }
    else s = ((long)n * K + k) * 4 + t;
    dst[i] = __float2bfloat16_rn(src[s]);
  }
}

__device__ __forceinline__ long pack_src_index(int mode, int Nn, int T, int K, int Cout, int n, int t, int k) {
  if (mode == 0) return ((long)n * K + k) * T + t;
  if (mode == 1) return ((long)k * Nn + n) * T + (T - 1 - t);
  if (mode == 2) { const int q = n / Cout, co = n % Cout; return ((long)k * Cout + co) * 4 + q; }
  return ((long)n * K + k) * 4 + t;
}

// All weight operands of the network in ONE launch: a block looks its job up in the table by binary search over the
// jobs' first-block prefix.  Work units (pack_job_blocks() tells the host how many a job has):
//   mode 0 (conv fwd):   one unit = (n, 128 input channels): 128*T consecutive source floats are staged in shared memory
//                        and written as T rows of 128 bf16 -- both sides coalesced (the element-wise version read one
//                        float per 32-byte sector: the kernel was bound by L2 sector traffic, 0.5 ms per step)
//   mode 1 (conv dgrad): one unit = (32 output channels, 32 input channels): per output channel 32*T consecutive source
//                        floats; written as 64-byte runs of the transposed, tap-reversed operand
//   mode 2, 3, 4, 5:     1024 consecutive destination elements (small tensors)
//   mode 4: stem operand [2C][Kp] -- rows 0..C-1 the 3x3 conv1 weights as (tap-major, channel) columns, rows C..2C-1 the
//   1x1 shortcut weights in the centre-tap columns, zero elsewhere (Nn = 2C, K = Kp, T = number of input channels, src =
//   conv1 weight, src2 = shortcut weight); mode 5: the same without the shortcut rows.
constexpr int PK0 = 128;          // mode 0: channels per unit
constexpr int PK1 = 32;           // mode 1: output channels (k) per unit
constexpr int PK1N = 16;          // mode 1: input channels (n) per unit (18 KB of shared memory: 8 blocks per SM)
constexpr int PACK_MAX_T = 9;

__host__ __device__ inline long long pack_job_blocks(int Nn, int T, int K, int mode) {
  if (mode == 0 && T <= PACK_MAX_T) return (long long)Nn * ((K + PK0 - 1) / PK0);
  if (mode == 1 && T <= PACK_MAX_T) return (long long)((K + PK1 - 1) / PK1) * ((Nn + PK1N - 1) / PK1N);
  const long long total = (mode == 4 || mode == 5) ? (long long)Nn * K : (long long)Nn * T * K;
  return (total + 1023) / 1024;
}

__global__ void __launch_bounds__(256)
pack_multi_kernel(const rbu_pack_job* __restrict__ jobs, int njobs) {
  __shared__ float tile[PK1 * PK1N * PACK_MAX_T];     // 18 KB (mode 1); mode 0 uses the first 128*T floats
  // job lookup: every thread tests one table entry (one global-load latency instead of a 7-deep dependent binary search)
  __shared__ int job_idx;
  const long long b = blockIdx.x;
  for (int i = threadIdx.x; i < njobs; i += 256) {
    const long long fb = jobs[i].first_block;
    const long long nb = i + 1 < njobs ? jobs[i + 1].first_block : 0x7fffffffffffffffLL;
    if (fb <= b && b < nb) job_idx = i;
  }
  __syncthreads();
  const rbu_pack_job j = jobs[job_idx];
  bf16* dst = reinterpret_cast<bf16*>(j.dst);
  const long long lb = b - j.first_block;
  const int T = j.T, K = j.K, Nn = j.Nn;
  if (j.mode == 0 && T <= PACK_MAX_T) {
    const int kchunks = (K + PK0 - 1) / PK0;
    const int n = (int)(lb / kchunks), k0 = (int)(lb - (long long)n * kchunks) * PK0;
    const int kw = min(PK0, K - k0);
    const float* sp = j.src + ((long long)n * K + k0) * T;
    {
      // [k][t], contiguous in the source; all loads of a thread are issued before the first shared-memory store
      float v[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int i = threadIdx.x + u * 256;
        v[u] = i < kw * T ? __ldg(sp + i) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int i = threadIdx.x + u * 256;
        if (i < kw * T) tile[i] = v[u];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kw * T; i += 256) {
      const int t = i / kw, k = i - t * kw;
      dst[((long long)n * T + t) * K + k0 + k] = __float2bfloat16_rn(tile[k * T + t]);
    }
    return;
  }
  if (j.mode == 1 && T <= PACK_MAX_T) {
    // src [K][Nn][T] -> dst[n][T-1-t][k]
    const int nchunks = (Nn + PK1N - 1) / PK1N;
    const int kc = (int)(lb / nchunks), k0 = kc * PK1, n0 = (int)(lb - (long long)kc * nchunks) * PK1N;
    const int kw = min(PK1, K - k0), nw = min(PK1N, Nn - n0);
    const int row = nw * T;                                                   // contiguous source floats per k
    for (int base = 0; base < kw * row; base += 256 * 6) {
      float v[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = base + threadIdx.x + u * 256;
        float x = 0.f;
        if (i < kw * row) {
          const int kk = i / row, r = i - kk * row;
          x = __ldg(j.src + ((long long)(k0 + kk) * Nn + n0) * T + r);
        }
        v[u] = x;
      }
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = base + threadIdx.x + u * 256;
        if (i < kw * row) {
          const int kk = i / row, r = i - kk * row;
          tile[kk * (PK1N * PACK_MAX_T) + r] = v[u];
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kw * row; i += 256) {
      const int kk = i % kw, r = i / kw;                                       // r = nn * T + t'
      const int nn = r / T, tp = r - nn * T;
      dst[((long long)(n0 + nn) * T + tp) * K + k0 + kk] =
          __float2bfloat16_rn(tile[kk * (PK1N * PACK_MAX_T) + nn * T + (T - 1 - tp)]);
    }
    return;
  }
  const long long base = lb * 1024;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long i = base + u * 256 + threadIdx.x;
    if (i >= j.total) return;
    float v;
    if (j.mode == 5) {          // stem operand of a plain conv (no shortcut rows): [C][Kp], Nn = C, T = input channels
      const int nc = j.T;
      const int k = (int)(i % j.K), n = (int)(i / j.K);
      v = 0.f;
      if (k < 9 * nc) { const int tap = k / nc, c = k - tap * nc; v = j.src[((long)n * nc + c) * 9 + tap]; }
    } else if (j.mode == 4) {
      const int C = j.Nn / 2, nc = j.T;
      const int k = (int)(i % j.K), n = (int)(i / j.K);
      v = 0.f;
      if (n < C) {
        if (k < 9 * nc) { const int tap = k / nc, c = k - tap * nc; v = j.src[((long)n * nc + c) * 9 + tap]; }
      } else if (k >= 4 * nc && k < 5 * nc) {
        v = j.src2[(long)(n - C) * nc + (k - 4 * nc)];
      }
    } else {
      const int k = (int)(i % j.K);
      const int t = (int)((i / j.K) % j.T);
      const int n = (int)(i / ((long long)j.K * j.T));
      v = j.src[pack_src_index(j.mode, j.Nn, j.T, j.K, j.Cout, n, t, k)];
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}

__global__ void conv_direct_ref_kernel(const bf16* __restrict__ x, long x_ld, int N, int H, int W, int Cin,
                                       const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                                       int ksz, int dil, float* __restrict__ out) {
  const long total = (long)N * H * W * Cout;
  const int pad = dil * (ksz / 2);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long pix = i / Cout;
    const int wq = (int)(pix % W);
    const int hq = (int)((pix / W) % H);
    const int n = (int)(pix / ((long)W * H));
    float acc = bias ? bias[co] : 0.f;
    for (int r = 0; r < ksz; ++r) {
      const int hh = hq + r * dil - pad;
      if (hh < 0 || hh >= H) continue;
      for (int s = 0; s < ksz; ++s) {
        const int ww = wq + s * dil - pad;
        if (ww < 0 || ww >= W) continue;
        const bf16* xp = x + (((long)n * H + hh) * W + ww) * x_ld;
        const float* wp = w + ((long)co * Cin * ksz + r) * ksz + s;
        for (int ci = 0; ci < Cin; ++ci) acc += __bfloat162float(xp[ci]) * bf16_round(wp[(long)ci * ksz * ksz]);
      }
    }
    out[i] = acc;
  }
}

}  // namespace

extern "C" int rbu_pack_weight(const float* src, void* dst, int Nn, int T, int K, int mode, int Cout, void* stream) {
  RBU_CHECK_ARG(src && dst && Nn > 0 && T > 0 && K > 0 && mode >= 0 && mode <= 3, "rbu_pack_weight: bad arguments");
  RBU_CHECK_ARG(mode != 2 || (Cout > 0 && Nn == 4 * Cout && T == 1), "rbu_pack_weight: mode 2 needs Nn == 4*Cout, T == 1");
  RBU_CHECK_ARG(mode != 3 || T == 4, "rbu_pack_weight: mode 3 needs T == 4");
  const long total = (long)Nn * T * K;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, Nn, T, K, mode, Cout);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" long long rbu_pack_job_blocks(int Nn, int T, int K, int mode) { return pack_job_blocks(Nn, T, K, mode); }

extern "C" int rbu_pack_weights_multi(const rbu_pack_job* jobs_device, int njobs, long long total_blocks, void* stream) {
  RBU_CHECK_ARG(jobs_device && njobs > 0 && total_blocks > 0 && total_blocks < (1LL << 31), "rbu_pack_weights_multi: bad arguments");
  pack_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs_device, njobs);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_conv_direct_ref(const void* x, int64_t x_ld, int N, int H, int W, int Cin, const float* w,
                                   const float* bias, int Cout, int ksz, int dil, float* out, void* stream) {
  RBU_CHECK_ARG(x && w && out && (ksz == 1 || ksz == 3 || ksz == 7) && dil >= 1, "rbu_conv_direct_ref: bad arguments");
  const long total = (long)N * H * W * Cout;
  const int blocks = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
  conv_direct_ref_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, x_ld, N, H, W, Cin, w, bias, Cout,
                                                                    ksz, dil, out);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}
