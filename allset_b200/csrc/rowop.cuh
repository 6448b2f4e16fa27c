// rowop.cuh -- row-wise dense glue with mixed-precision I/O, dropout and a fused backward (included by
// allset_kernels.cu inside its anonymous namespace).
//
//   z = residual + act(x + bias);   y = LayerNorm_{gamma,beta}(z);   out = dropout_p(act2(y))    (each stage optional)
//
// This is what sits between the Linears of the reference's MLP (reference src/layers.py:571-579: Linear -> ReLU ->
// norm -> dropout) and around PMA's rFF (src/layers.py:153-157), as ONE pass over the rows forward and ONE pass
// backward, with bf16 or fp32 rows on either side (fp32 statistics and parameters).  The training path of the bf16 mode
// is "bf16 tensor-core GEMM -> rowop -> GEMM -> rowop" forward and the mirrored chain backward; ATen spends a bias kernel,
// a ReLU kernel, a LayerNorm kernel (1.67 ms per [1M,128] rows measured) and a dropout kernel per stage, each with an
// autograd-saved copy.
//
// Layout: one warp per row, the row in registers.  d = 32 * NPL, NPL in {2,4,8,16,32}; a lane owns NCH chunks of
// CE = min(NPL, 4) consecutive elements, chunk c of lane l = columns (c*32 + l)*CE ..+CE, so every warp access is one
// contiguous 32*CE*sizeof(T)-byte piece.
//
// Dropout is counter-based: element (row, col) is kept iff a 16-bit slice of a hash of (seed, row, col / 2) >= p*65536,
// so the backward pass regenerates the mask instead of reading one (torch's dropout saves a bool mask: +1 byte per
// element each way).  Statistically equivalent to F.dropout, not the same stream of random numbers.

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace rowop {

using bf16 = __nv_bfloat16;

// 32-bit finaliser (two multiply / xor-shift rounds, "lowbias32"): ~8 instructions for the 32 random bits that decide
// two elements.  A 64-bit splitmix per four elements cost ~11 instructions per element -- a third of the forward kernel.
__device__ __forceinline__ unsigned mix32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU;
  x ^= x >> 15; x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

template <typename T, int CE>
struct Vec;                                     // CE consecutive elements of type T <-> CE floats

template <int CE>
struct Vec<float, CE> {
  __device__ static __forceinline__ void load(const float* p, float (&f)[CE]) {
    if (CE == 4) { const float4 v = __ldcs(reinterpret_cast<const float4*>(p)); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
    else { const float2 v = __ldcs(reinterpret_cast<const float2*>(p)); f[0] = v.x; f[1] = v.y; }
  }
  __device__ static __forceinline__ void store(float* p, const float (&f)[CE]) {
    if (CE == 4) *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    else *reinterpret_cast<float2*>(p) = make_float2(f[0], f[1]);
  }
};

template <int CE>
struct Vec<bf16, CE> {
  __device__ static __forceinline__ void load(const bf16* p, float (&f)[CE]) {
    if (CE == 4) {
      const uint2 v = __ldcs(reinterpret_cast<const uint2*>(p));
      f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
      f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    } else {
      const unsigned v = __ldcs(reinterpret_cast<const unsigned*>(p));
      f[0] = __uint_as_float(v << 16); f[1] = __uint_as_float(v & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void store(bf16* p, const float (&f)[CE]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
    if (CE == 4) {
      const __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
      *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const unsigned*>(&a), *reinterpret_cast<const unsigned*>(&b));
    } else {
      *reinterpret_cast<unsigned*>(p) = *reinterpret_cast<const unsigned*>(&a);
    }
  }
};

template <int CE>
__device__ __forceinline__ void load_param(const float* p, float (&f)[CE]) {
  if (CE == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  else { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); f[0] = v.x; f[1] = v.y; }
}

// keep-mask bits of the CE elements of one chunk (bit k set = keep element k): one 32-bit hash per PAIR of consecutive
// columns, 16 bits per element, keyed by (seed, row, column pair)
template <int CE>
__device__ __forceinline__ unsigned keep_bits(uint64_t seed, long long row, int d, int col, unsigned thr16) {
  const unsigned s_lo = (unsigned)seed, s_hi = (unsigned)(seed >> 32);
  const unsigned base = (unsigned)row * (unsigned)(d >> 1) + (unsigned)(col >> 1);
  unsigned bits = 0;
#pragma unroll
  for (int p = 0; p < CE / 2; ++p) {
    const unsigned h = mix32((base + (unsigned)p) * 0x9E3779B1U + s_lo) ^ s_hi;
    bits |= ((h & 0xffffu) >= thr16 ? 1u : 0u) << (2 * p);
    bits |= ((h >> 16) >= thr16 ? 1u : 0u) << (2 * p + 1);
  }
  return bits;
}

// Sum each of N per-lane values over the 32 lanes and leave ALL N totals in every lane.  Instead of N butterflies
// (5 N shuffles) the first log2(N) steps are TRANSPOSED -- a lane keeps half of its values and hands the other half to
// its partner -- so one value per lane is left for the remaining steps, and the totals are fetched back with one indexed
// shuffle each: N-1 + (5 - log2 N) + N shuffles (N = 4: 10 instead of 20; N = 8: 17 instead of 40).
template <int N>
__device__ __forceinline__ void warp_allsum(float (&v)[N], int lane) {
  constexpr unsigned kFull = 0xffffffffu;
  int bit = 16;
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1, bit >>= 1) {
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = upper ? v[half + i] : v[i];
      const float send = upper ? v[i] : v[half + i];
      v[i] = keep + __shfl_xor_sync(kFull, send, bit);
    }
  }
#pragma unroll
  for (; bit >= 1; bit >>= 1) v[0] += __shfl_xor_sync(kFull, v[0], bit);
  const float mine = v[0];
#pragma unroll
  for (int r = 0; r < N; ++r) {
    int src = 0, b = 16;
#pragma unroll
    for (int half = N / 2; half >= 1; half >>= 1, b >>= 1)
      if (r & half) src |= b;
    v[r] = __shfl_sync(kFull, mine, src);
  }
}

struct FwdArgs {
  const void* x; const float* bias; int relu; const void* residual; const float* gamma; const float* beta; float eps;
  float keep_scale; unsigned thr16; uint64_t seed; long long rows; void* out; float* stats; int relu_out;
};

struct BwdArgs {
  const void* dy; const void* x; const float* bias; int relu; const void* residual; const float* gamma;
  const float* stats; float keep_scale; unsigned thr16; uint64_t seed; long long rows; void* dx; void* dres;
  float* partial; const float* beta; int relu_out;
};

// Which stages a launch has.  The generic kernels (F < 0) read the flags from the arguments; at ~110 SASS instructions
// per element (every optional stage present as predicated code, the dropout hash computed to be discarded) they were
// ISSUE-bound at 2.9 TB/s, so the stage combinations the MLP / PMA chains actually launch are compiled with the flags
// as constants and the dead stages removed.
enum : int { kBias = 1, kRelu = 2, kRes = 4, kLn = 8, kReluOut = 16, kDrop = 32 };
template <int F>
struct Flag {
  template <typename A> __device__ static __forceinline__ bool bias(const A& a) { return F < 0 ? a.bias != nullptr : (F & kBias) != 0; }
  template <typename A> __device__ static __forceinline__ bool relu(const A& a) { return F < 0 ? a.relu != 0 : (F & kRelu) != 0; }
  template <typename A> __device__ static __forceinline__ bool res(const A& a) { return F < 0 ? a.residual != nullptr : (F & kRes) != 0; }
  template <typename A> __device__ static __forceinline__ bool ln(const A& a) { return F < 0 ? a.gamma != nullptr : (F & kLn) != 0; }
  template <typename A> __device__ static __forceinline__ bool relu_out(const A& a) { return F < 0 ? a.relu_out != 0 : (F & kReluOut) != 0; }
  template <typename A> __device__ static __forceinline__ bool drop(const A& a) { return F < 0 ? a.thr16 != 0u : (F & kDrop) != 0; }
};
__host__ inline int flags_of(bool bias, bool relu, bool res, bool ln, bool relu_out, bool drop) {
  return (bias ? kBias : 0) | (relu ? kRelu : 0) | (res ? kRes : 0) | (ln ? kLn : 0) | (relu_out ? kReluOut : 0) | (drop ? kDrop : 0);
}

// rows a warp holds at once: the loads of all of them are issued before the first use, so a warp has R row-sized
// requests in flight (a 256-byte bf16 row per warp is far too little to cover HBM latency: measured 2.3 TB/s with R = 1)
template <int NPL>
struct RowsPerWarp { static constexpr int R = NPL <= 4 ? 4 : (NPL == 8 ? 2 : 1); };

template <typename TIn, typename TOut, int NPL, int F>
__global__ void __launch_bounds__(256) fwd_kernel(FwdArgs a) {
  constexpr int D = 32 * NPL;
  constexpr int CE = NPL < 4 ? NPL : 4;
  constexpr int NCH = NPL / CE;
  constexpr int R = RowsPerWarp<NPL>::R;
  using FL = Flag<F>;
  const int lane = threadIdx.x & 31;
  const long long row0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * R;
  if (row0 >= a.rows) return;
  const int nr = (int)((a.rows - row0) < R ? (a.rows - row0) : R);
  float v[R][NCH][CE];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const TIn* xr = static_cast<const TIn*>(a.x) + (row0 + (r < nr ? r : 0)) * D;
#pragma unroll
    for (int c = 0; c < NCH; ++c) Vec<TIn, CE>::load(xr + (c * 32 + lane) * CE, v[r][c]);
  }
  if (FL::bias(a)) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float b[CE];
      load_param<CE>(a.bias + (c * 32 + lane) * CE, b);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < CE; ++k) v[r][c][k] += b[k];
    }
  }
  if (FL::relu(a)) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int k = 0; k < CE; ++k) v[r][c][k] = fmaxf(v[r][c][k], 0.f);
  }
  if (FL::res(a)) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const TIn* rr = static_cast<const TIn*>(a.residual) + (row0 + (r < nr ? r : 0)) * D;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float rs[CE];
        Vec<TIn, CE>::load(rr + (c * 32 + lane) * CE, rs);
#pragma unroll
        for (int k = 0; k < CE; ++k) v[r][c][k] += rs[k];
      }
    }
  }
  if (FL::ln(a)) {
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int k = 0; k < CE; ++k) sum += v[r][c][k];
      mean[r] = sum;
    }
    warp_allsum<R>(mean, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mean[r] *= (1.f / D);
      float sq = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int k = 0; k < CE; ++k) { const float t = v[r][c][k] - mean[r]; sq += t * t; }
      rstd[r] = sq;
    }
    warp_allsum<R>(rstd, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rstd[r] = rsqrtf(rstd[r] * (1.f / D) + a.eps);
      if (a.stats != nullptr && lane == 0 && r < nr)
        *reinterpret_cast<float2*>(a.stats + (row0 + r) * 2) = make_float2(mean[r], rstd[r]);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float g[CE], b[CE];
      load_param<CE>(a.gamma + (c * 32 + lane) * CE, g);
#pragma unroll
      for (int k = 0; k < CE; ++k) b[k] = 0.f;
      if (a.beta != nullptr) load_param<CE>(a.beta + (c * 32 + lane) * CE, b);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < CE; ++k) v[r][c][k] = fmaf((v[r][c][k] - mean[r]) * rstd[r], g[k], b[k]);
    }
  }
  if (FL::relu_out(a)) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int k = 0; k < CE; ++k) v[r][c][k] = fmaxf(v[r][c][k], 0.f);
  }
  if (FL::drop(a)) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const unsigned keep = keep_bits<CE>(a.seed, row0 + r, D, (c * 32 + lane) * CE, a.thr16);
#pragma unroll
        for (int k = 0; k < CE; ++k) v[r][c][k] = ((keep >> k) & 1u) ? v[r][c][k] * a.keep_scale : 0.f;
      }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (r >= nr) break;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      Vec<TOut, CE>::store(static_cast<TOut*>(a.out) + (row0 + r) * D + (c * 32 + lane) * CE, v[r][c]);
  }
}

// Backward.  A warp walks rows with a grid stride so that every lane owns fixed columns: d(gamma), d(beta), d(bias)
// accumulate in registers over all rows of the warp, are combined per CTA through shared memory and written as ONE
// partial row per CTA (partial[cta][3][D]; the host sums over CTAs: deterministic, no atomics).
//   g0 = dy * keep / (1-p) * [y > 0 if act2];  zh = (z - mean) * rstd;  y = zh * gamma + beta;  g = g0 * gamma
//   dz = rstd * (g - mean_d(g) - zh * mean_d(g * zh))        (dz = g0 without LayerNorm)
//   d(residual) = dz;   d(x) = dz * [x + bias > 0]  (relu)  else dz
template <typename TG, typename TX, int NPL, int F>
__global__ void __launch_bounds__(256) bwd_kernel(BwdArgs a) {
  constexpr int D = 32 * NPL;
  constexpr int CE = NPL < 4 ? NPL : 4;
  constexpr int NCH = NPL / CE;
  constexpr int R = RowsPerWarp<NPL>::R;
  using FL = Flag<F>;
  __shared__ float red[8][3][32 * CE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * 8;
  float bsv[NCH][CE], gmv[NCH][CE], btv[NCH][CE], dgam[NCH][CE], dbet[NCH][CE], dbia[NCH][CE];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int k = 0; k < CE; ++k) { bsv[c][k] = 0.f; gmv[c][k] = 1.f; btv[c][k] = 0.f; dgam[c][k] = dbet[c][k] = dbia[c][k] = 0.f; }
    if (FL::bias(a)) load_param<CE>(a.bias + (c * 32 + lane) * CE, bsv[c]);
    if (FL::ln(a)) load_param<CE>(a.gamma + (c * 32 + lane) * CE, gmv[c]);
    if (FL::relu_out(a) && a.beta != nullptr) load_param<CE>(a.beta + (c * 32 + lane) * CE, btv[c]);
  }
  for (long long row0 = ((long long)blockIdx.x * 8 + warp) * R; row0 < a.rows; row0 += nwarps * R) {
    const int nr = (int)((a.rows - row0) < R ? (a.rows - row0) : R);
    float pre[R][NCH][CE], g[R][NCH][CE], zh[R][NCH][CE];
    float mean[R], rstd[R], s1[R], s2[R];
    // all loads of the R rows first (independent requests in flight), then the arithmetic
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + (r < nr ? r : 0);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        Vec<TX, CE>::load(static_cast<const TX*>(a.x) + row * D + (c * 32 + lane) * CE, pre[r][c]);
        Vec<TG, CE>::load(static_cast<const TG*>(a.dy) + row * D + (c * 32 + lane) * CE, g[r][c]);
        if (FL::res(a)) Vec<TX, CE>::load(static_cast<const TX*>(a.residual) + row * D + (c * 32 + lane) * CE, zh[r][c]);
      }
      mean[r] = 0.f;
      rstd[r] = 1.f;
      if (FL::ln(a)) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(a.stats + row * 2));
        mean[r] = st.x;
        rstd[r] = st.y;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      s1[r] = 0.f;
      s2[r] = 0.f;
      if (r >= nr) {                                   // rows beyond the end contribute nothing to the column sums
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int k = 0; k < CE; ++k) g[r][c][k] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        unsigned keep = 0xfu;
        if (FL::drop(a)) keep = keep_bits<CE>(a.seed, row0 + r, D, (c * 32 + lane) * CE, a.thr16);
#pragma unroll
        for (int k = 0; k < CE; ++k) {
          if (FL::bias(a)) pre[r][c][k] += bsv[c][k];
          float z = FL::relu(a) ? fmaxf(pre[r][c][k], 0.f) : pre[r][c][k];
          if (FL::res(a)) z += zh[r][c][k];
          float g0 = g[r][c][k];
          if (FL::drop(a)) g0 = ((keep >> k) & 1u) ? g0 * a.keep_scale : 0.f;
          if (FL::ln(a)) {
            const float h = (z - mean[r]) * rstd[r];
            zh[r][c][k] = h;
            if (FL::relu_out(a) && !(fmaf(h, gmv[c][k], btv[c][k]) > 0.f)) g0 = 0.f;
            const float gg = g0 * gmv[c][k];
            g[r][c][k] = gg;
            dgam[c][k] = fmaf(g0, h, dgam[c][k]);
            dbet[c][k] += g0;
            s1[r] += gg;
            s2[r] = fmaf(gg, h, s2[r]);
          } else {
            if (FL::relu_out(a) && !(z > 0.f)) g0 = 0.f;
            g[r][c][k] = g0;
          }
        }
      }
    }
    if (FL::ln(a)) {
      float both[2 * R];
#pragma unroll
      for (int r = 0; r < R; ++r) { both[r] = s1[r]; both[R + r] = s2[r]; }
      warp_allsum<2 * R>(both, lane);
#pragma unroll
      for (int r = 0; r < R; ++r) { s1[r] = both[r]; s2[r] = both[R + r]; }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r >= nr) break;
      const long long row = row0 + r;
      const float m1 = s1[r] * (1.f / D), m2 = s2[r] * (1.f / D);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col = (c * 32 + lane) * CE;
        float dz[CE], dp[CE];
#pragma unroll
        for (int k = 0; k < CE; ++k) {
          dz[k] = FL::ln(a) ? rstd[r] * (g[r][c][k] - m1 - zh[r][c][k] * m2) : g[r][c][k];
          dp[k] = (FL::relu(a) && !(pre[r][c][k] > 0.f)) ? 0.f : dz[k];
          if (FL::bias(a)) dbia[c][k] += dp[k];
        }
        if (FL::res(a) && a.dres != nullptr) Vec<TG, CE>::store(static_cast<TG*>(a.dres) + row * D + col, dz);
        Vec<TG, CE>::store(static_cast<TG*>(a.dx) + row * D + col, dp);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int k = 0; k < CE; ++k) {
      red[warp][0][lane * CE + k] = dgam[c][k];
      red[warp][1][lane * CE + k] = dbet[c][k];
      red[warp][2][lane * CE + k] = dbia[c][k];
    }
    __syncthreads();
    if (warp < 3) {                                 // warp q sums quantity q over the 8 warps
#pragma unroll
      for (int k = 0; k < CE; ++k) {
        float s = red[0][warp][lane * CE + k];
#pragma unroll
        for (int w2 = 1; w2 < 8; ++w2) s += red[w2][warp][lane * CE + k];
        a.partial[((size_t)blockIdx.x * 3 + warp) * D + (c * 32 + lane) * CE + k] = s;
      }
    }
    __syncthreads();
  }
}

// ---- dispatch: the stage combinations the MLP / PMA chains launch get constant flags (same dtype on both sides, widths
//      64 / 128 / 256); everything else takes the generic kernels (F = -1) -------------------------------------------------
#define ROWOP_COMBOS(X)                                                                                   \
  X(kLn) X(kBias | kRelu | kLn | kDrop) X(kBias | kRelu | kLn) X(kBias | kRelu | kDrop) X(kBias | kRelu) \
  X(kBias) X(kBias | kRelu | kRes | kLn | kReluOut | kDrop) X(kBias | kRelu | kRes | kLn | kReluOut) X(kDrop)

template <typename TIn, typename TOut, int NPL>
void launch_fwd_n(const FwdArgs& a, int flags, bool specialise, cudaStream_t st) {
  constexpr int R = RowsPerWarp<NPL>::R;
  const long long warps = (a.rows + R - 1) / R;
  const unsigned blocks = (unsigned)((warps + 7) / 8);
  if (specialise) {
#define ROWOP_CASE(FV) if (flags == (FV)) { fwd_kernel<TIn, TOut, NPL, (FV)><<<blocks, 256, 0, st>>>(a); return; }
    ROWOP_COMBOS(ROWOP_CASE)
#undef ROWOP_CASE
  }
  fwd_kernel<TIn, TOut, NPL, -1><<<blocks, 256, 0, st>>>(a);
}

template <typename TIn, typename TOut>
bool launch_fwd(const FwdArgs& a, int d, cudaStream_t st) {
  const int flags = flags_of(a.bias != nullptr, a.relu != 0, a.residual != nullptr, a.gamma != nullptr, a.relu_out != 0,
                             a.thr16 != 0u);
  switch (d) {
    case 64: launch_fwd_n<TIn, TOut, 2>(a, flags, true, st); return true;
    case 128: launch_fwd_n<TIn, TOut, 4>(a, flags, true, st); return true;
    case 256: launch_fwd_n<TIn, TOut, 8>(a, flags, true, st); return true;
    case 512: launch_fwd_n<TIn, TOut, 16>(a, flags, false, st); return true;
    case 1024: launch_fwd_n<TIn, TOut, 32>(a, flags, false, st); return true;
    default: return false;
  }
}

template <typename TG, typename TX, int NPL>
void launch_bwd_n(const BwdArgs& a, int flags, bool specialise, unsigned blocks, cudaStream_t st) {
  if (specialise) {
#define ROWOP_CASE(FV) if (flags == (FV)) { bwd_kernel<TG, TX, NPL, (FV)><<<blocks, 256, 0, st>>>(a); return; }
    ROWOP_COMBOS(ROWOP_CASE)
#undef ROWOP_CASE
  }
  bwd_kernel<TG, TX, NPL, -1><<<blocks, 256, 0, st>>>(a);
}

template <typename TG, typename TX>
bool launch_bwd(const BwdArgs& a, int d, unsigned blocks, cudaStream_t st) {
  const int flags = flags_of(a.bias != nullptr, a.relu != 0, a.residual != nullptr, a.gamma != nullptr, a.relu_out != 0,
                             a.thr16 != 0u);
  constexpr bool same = sizeof(TG) == sizeof(TX);       // mixed-dtype gradients only occur at the model's edges
  switch (d) {
    case 64: launch_bwd_n<TG, TX, 2>(a, flags, same, blocks, st); return true;
    case 128: launch_bwd_n<TG, TX, 4>(a, flags, same, blocks, st); return true;
    case 256: launch_bwd_n<TG, TX, 8>(a, flags, same, blocks, st); return true;
    case 512: launch_bwd_n<TG, TX, 16>(a, flags, false, blocks, st); return true;
    case 1024: launch_bwd_n<TG, TX, 32>(a, flags, false, blocks, st); return true;
    default: return false;
  }
}

}  // namespace rowop
