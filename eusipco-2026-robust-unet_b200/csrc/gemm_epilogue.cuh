// Shared epilogue of the implicit-GEMM convolution kernels: one warp drains 32 fp32 accumulator columns of its
// 32 TMEM lanes (= 32 output pixels), adds bias / addend, rounds to bf16 and stores 16 bytes per 8 channels.
// Eight epilogue warps per CTA (two per TMEM lane group, interleaved over the 32-column chunks) keep enough global
// loads / stores in flight for the small-K GEMMs whose epilogue, not the MMA, bounds the tile time; the addend and
// bias loads of a chunk are issued BEFORE waiting on tcgen05.ld so their latency overlaps the TMEM read.
#pragma once
#include "rbu_common.cuh"
#include "rbu_ptx.cuh"

constexpr int EPI_WARPS = 8;
constexpr int EPI_STAT_CHUNKS = 4;                                   // 32-column chunks per epilogue warp (block_n <= 256)
constexpr int EPI_STAT_FLOATS = EPI_WARPS * EPI_STAT_CHUNKS * 64;    // per-CTA BatchNorm accumulators: 8 KB of shared memory

// Column sums over the 32 lanes of a warp of 32 per-lane values (recursive halving: 31 shuffles): lane l returns
// the sum over all lanes of v[l].  Destroys v.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#define RBU_HALVE(OFF, HALF)                                                        \
  {                                                                                 \
    const bool up = (lane & (OFF)) != 0;                                            \
    _Pragma("unroll") for (int i = 0; i < (HALF); ++i) {                            \
      const float send = up ? v[i] : v[i + (HALF)];                                 \
      const float keep = up ? v[i + (HALF)] : v[i];                                 \
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, (OFF));                      \
    }                                                                               \
  }
  RBU_HALVE(16, 16) RBU_HALVE(8, 8) RBU_HALVE(4, 4) RBU_HALVE(2, 2) RBU_HALVE(1, 1)
#undef RBU_HALVE
  return v[0];
}

// Division by a launch-constant divisor without the ~25-instruction integer-division sequence (tile -> coordinate
// decoding sits on every epilogue warp's per-tile path).  Exact for 0 <= x < 2^31.
struct FastDiv {
  uint32_t mul, shr;
  int d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  f.mul = 0;
  f.shr = 0;
  if (d > 1) {
    int l = 0;
    while ((1ll << l) < d) ++l;                                   // ceil(log2 d) >= 1
    f.mul = (uint32_t)(((1ull << (31 + l)) + (uint64_t)d - 1) / (uint64_t)d);
    f.shr = (uint32_t)(l - 1);
  }
  return f;
}
__device__ __forceinline__ int fast_div(int x, const FastDiv& f) {
  return f.d == 1 ? x : (int)(__umulhi((uint32_t)x, f.mul) >> f.shr);
}

// Same recursive halving for max / min (per-image pooled statistics of ChannelAttention).
template <bool MAX>
__device__ __forceinline__ float warp_colext32(float (&v)[32], int lane) {
#define RBU_HALVE(OFF, HALF)                                                        \
  {                                                                                 \
    const bool up = (lane & (OFF)) != 0;                                            \
    _Pragma("unroll") for (int i = 0; i < (HALF); ++i) {                            \
      const float send = up ? v[i] : v[i + (HALF)];                                 \
      const float keep = up ? v[i + (HALF)] : v[i];                                 \
      const float got = __shfl_xor_sync(0xffffffffu, send, (OFF));                  \
      v[i] = MAX ? fmaxf(keep, got) : fminf(keep, got);                             \
    }                                                                               \
  }
  RBU_HALVE(16, 16) RBU_HALVE(8, 8) RBU_HALVE(4, 4) RBU_HALVE(2, 2) RBU_HALVE(1, 1)
#undef RBU_HALVE
  return v[0];
}

struct EpiOut {
  float* stat_acc;   // this warp's [EPI_STAT_CHUNKS][32][2] shared-memory accumulators, or nullptr (statistics off)
  bf16* y;
  long long y_ld;
  const float* bias;
  const float* scale;   // optional per-column scale applied to the accumulator before the bias (eval-mode BatchNorm fold)
  int relu;             // ReLU after scale / bias / addend on the first `relu` output channels (0: none)
  const bf16* addend;
  long long addend_ld;
  int Ncols;
  int scatter, Cout, H, W;   // scatter: ConvTranspose2d pixel shuffle, column (2i+j)*Cout+co of (n,h,w) -> y[n,2h+i,2w+j,co]
};

// Issue the addend loads of one 32-column chunk (no-op without an addend).  They are consumed by epi_finish(); the
// tile loops issue them for the NEXT tile's first chunk before waiting on the current accumulator, so the global-load
// latency of the small-K GEMMs (attention-gate data gradients: K = 32..256, read-modify-write of a 2x wider tensor)
// overlaps the previous tile instead of sitting on every tile's critical path.
__device__ __forceinline__ void epi_prefetch(const EpiOut& o, int col, bool valid, long long pix, uint4 (&ad)[4]) {
  if (o.addend && valid && col < o.Ncols) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (col + g * 8 < o.Ncols) ad[g] = __ldg(reinterpret_cast<const uint4*>(o.addend + pix * o.addend_ld + col + g * 8));
  }
}

// taddr: TMEM address of (lane group base, first column of the chunk); col: first GEMM column of the chunk
// `chunk` = index of this chunk among the warp's chunks of the tile (selects the statistics accumulator).
__device__ __forceinline__ void epi_finish(const EpiOut& o, uint32_t taddr, int col, bool valid, long long pix, int n, int h,
                                           int w, const uint4 (&ad)[4], int chunk = 0) {
  const bool live = valid && col < o.Ncols;
  uint32_t r[32];
  ptx::tmem_ld_32x32(taddr, r);
  ptx::tmem_ld_wait();
  float rv[32];      // the values as stored (bf16-rounded), zero for rows / columns outside the tensor
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c8 = col + g * 8;
    if (live && c8 < o.Ncols) {
      bf16* dst;
      int cb = c8;  // bias channel
      if (o.scatter) {
        const int q = c8 / o.Cout;
        cb = c8 - q * o.Cout;
        const long long opix = ((long long)n * (2 * o.H) + 2 * h + (q >> 1)) * (2 * o.W) + 2 * w + (q & 1);
        dst = o.y + opix * o.y_ld + cb;
      } else {
        dst = o.y + pix * o.y_ld + c8;
      }
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[g * 8 + e]);
      if (o.scale) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(o.scale + cb));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(o.scale + cb + 4));
        f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
        f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
      }
      if (o.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(o.bias + cb));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(o.bias + cb + 4));
        f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
        f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
      }
      if (o.addend) {
        float a8[8];
        bf16x8 t;
        *reinterpret_cast<uint4*>(&t) = ad[g];
        unpack8(t, a8);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += a8[e];
      }
      if (cb < o.relu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
      }
      const bf16x8 packed = pack8(f);
      st_bf16x8(dst, packed);
      if (o.stat_acc) unpack8(packed, &rv[g * 8]);
    } else if (o.stat_acc) {
#pragma unroll
      for (int e = 0; e < 8; ++e) rv[g * 8 + e] = 0.f;
    }
  }
  if (o.stat_acc) {
    // BatchNorm batch statistics of the stored tensor: per-column sum and sum of squares over this warp's 32 pixels,
    // accumulated in the warp's private shared-memory slots (no atomics; fixed order -> deterministic)
    const int lane = threadIdx.x & 31;
    float sq[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) sq[i] = rv[i] * rv[i];
    const float cs = warp_colsum32(rv, lane);
    const float cq = warp_colsum32(sq, lane);
    float2* slot = reinterpret_cast<float2*>(o.stat_acc) + chunk * 32 + lane;
    float2 cur = *slot;
    cur.x += cs;
    cur.y += cq;
    *slot = cur;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Staged variant: the warp writes its 32 pixels x 32 channels (2 KB, SWIZZLE_64B: the 16-byte piece g of pixel row r
// sits at r*64 + ((g ^ ((r >> 1) & 3)) << 4), conflict-free for 16-byte accesses) into shared memory and one lane
// hands the box to the TMA store engine.  No per-thread global addresses, no predicates (TMA clips the box at the
// tensor edge), and no L1 tag traffic (32 rows per st.global.v4 request otherwise).
// epi_math: accumulator chunk -> scale / bias / addend / ReLU -> four packed 16-byte pieces.
template <bool ADD, bool BIAS, bool EXTRA>   // EXTRA: per-column scale and / or ReLU (eval-mode folds)
__device__ __forceinline__ void epi_math(const uint32_t (&r)[32], const float* __restrict__ bias,
                                         const float* __restrict__ scale, int relu, int cb, const uint4 (&ad)[4],
                                         uint4 (&out)[4]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[g * 8 + e]);
    if (EXTRA) {
      if (scale) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + cb + g * 8));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + cb + g * 8 + 4));
        f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
        f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
      }
    }
    if (BIAS) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cb + g * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cb + g * 8 + 4));
      f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
      f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
    }
    if (ADD) {
      float a8[8];
      bf16x8 t;
      *reinterpret_cast<uint4*>(&t) = ad[g];
      unpack8(t, a8);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += a8[e];
    }
    if (EXTRA) {
      if (cb + g * 8 < relu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
      }
    }
    const bf16x8 packed = pack8(f);
    out[g] = *reinterpret_cast<const uint4*>(&packed);
  }
}
