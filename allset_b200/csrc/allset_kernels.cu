// allset_kernels.cu -- sm_100a kernels + C ABI for AllSet's V->E / E->V multiset aggregation.
//
// Replaces, per direction, the reference's index_select -> norm*x_j -> torch_scatter.scatter chain
// (reference src/layers.py:633,638-639,656) and PMA's gather -> leaky_relu -> segment softmax ->
// weight -> scatter-add chain (src/layers.py:145-153,168-194) with ONE launch over a CSR-by-target
// incidence list.  See include/allset_b200.h for the contract of each entry point and DESIGN.md
// for the data layout and the roofline each kernel is bounded by (HBM bandwidth for all of them).
//
// Work decomposition (common to all gather kernels)
//   * a feature row is cut into 16-byte chunks (8 bf16 / 4 fp32); lane `gl` of a lane GROUP of
//     G = 2^k <= 32 lanes owns chunk gl of the row, so one row load is a single coalesced
//     128-bit-per-lane request (d=128 bf16: G=16, two segments per warp; d=128 fp32: G=32);
//   * one group owns one target segment: it reads G column indices with one coalesced load,
//     broadcasts them with group-masked shuffles, and issues U=8 independent row loads
//     (ld.global.nc.L1::no_allocate.v4) before the first use -- the memory-level parallelism
//     that hides HBM latency on 256-byte random rows;
//   * accumulation is fp32 in registers in CSR (= caller's COO) order, so fp32 sums are
//     bit-identical to the reference's sequential CPU scatter_add_; no atomics anywhere;
//   * segments longer than `long_threshold` are skipped by the group kernel and handled by a
//     CTA-per-segment kernel (8 warps stride over the segment, shared-memory tree at the end);
//   * rows wider than 32 chunks are processed as independent 32-chunk slabs (blockIdx.y).
// Rows whose byte width or base address is not 16-byte aligned take the same kernels with
// 1-element chunks (scalar path) -- still CUDA, never a host fallback.

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cub/cub.cuh>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "allset_b200.h"
#include "rowop.cuh"

namespace {

thread_local char g_err[512] = {0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ALLSET_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return ALLSET_OK;
}

#ifndef ALLSET_THREADS
#define ALLSET_THREADS 256
#endif
#ifndef ALLSET_UNROLL_SUM
#define ALLSET_UNROLL_SUM 6
#endif
#ifndef ALLSET_UNROLL_PMA
#define ALLSET_UNROLL_PMA 4
#endif
constexpr int kThreads = ALLSET_THREADS;        // 8 warps per CTA
// Independent row loads in flight per lane.  Measured on B200 (10M/2M/deg-30 graph, d=128 bf16): the sum kernel
// peaks at 6 (U=4: 3.49 ms, U=6: 2.52 ms, U=8: 2.85 ms for V->E) -- more loads per lane cost registers and hence
// resident warps; the PMA kernels carry more state per lane and peak at 4.
constexpr int kUnrollSum = ALLSET_UNROLL_SUM;
constexpr int kUnrollPma = ALLSET_UNROLL_PMA;

// ---------------------------------------------------------------------------------------------
// chunk access: VECTOR = one 16-byte chunk per lane; otherwise one element per lane
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <typename T, bool VECTOR>
struct Chunk;

template <>
struct Chunk<float, true> {
  static constexpr int N = 4;
  using Raw = uint4;
  __device__ static __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  __device__ static __forceinline__ Raw load(const float* p) { return ld_nc_16(p); }
  __device__ static __forceinline__ void unpack(const Raw& r, float (&f)[N]) {
    f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y);
    f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
  }
  __device__ static __forceinline__ void store(float* p, const float (&f)[N]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
  __device__ static __forceinline__ void add(const Raw& r, float (&acc)[N]) {
    acc[0] = __fadd_rn(acc[0], __uint_as_float(r.x)); acc[1] = __fadd_rn(acc[1], __uint_as_float(r.y));
    acc[2] = __fadd_rn(acc[2], __uint_as_float(r.z)); acc[3] = __fadd_rn(acc[3], __uint_as_float(r.w));
  }
};

template <>
struct Chunk<__nv_bfloat16, true> {
  static constexpr int N = 8;
  using Raw = uint4;
  __device__ static __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  __device__ static __forceinline__ Raw load(const __nv_bfloat16* p) { return ld_nc_16(p); }
  __device__ static __forceinline__ void unpack(const Raw& r, float (&f)[N]) {
    // bf16 -> fp32 is a 16-bit shift: exact
    f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
    f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
    f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
    f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&f)[N]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]);
    __nv_bfloat162 e = __floats2bfloat162_rn(f[6], f[7]);
    uint4 o;
    o.x = *reinterpret_cast<unsigned*>(&a); o.y = *reinterpret_cast<unsigned*>(&b);
    o.z = *reinterpret_cast<unsigned*>(&c); o.w = *reinterpret_cast<unsigned*>(&e);
    *reinterpret_cast<uint4*>(p) = o;
  }
  // acc += bf16 halves, one FHADD.BF16 each (sm_100 mixed-precision add: bf16 -> fp32 is exact, the add rounds
  // to nearest in fp32), so a 16-byte chunk costs 8 instructions with no unpack.
  __device__ static __forceinline__ void add(const Raw& r, float (&acc)[N]) {
    add2(r.x, acc[0], acc[1]); add2(r.y, acc[2], acc[3]); add2(r.z, acc[4], acc[5]); add2(r.w, acc[6], acc[7]);
  }
  __device__ static __forceinline__ void add2(unsigned w, float& a0, float& a1) {
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\t"
        "add.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t}"
        : "+f"(a0), "+f"(a1) : "r"(w));
  }
};

template <>
struct Chunk<float, false> {
  static constexpr int N = 1;
  using Raw = float;
  __device__ static __forceinline__ Raw zero() { return 0.f; }
  __device__ static __forceinline__ Raw load(const float* p) { return __ldg(p); }
  __device__ static __forceinline__ void unpack(const Raw& r, float (&f)[N]) { f[0] = r; }
  __device__ static __forceinline__ void store(float* p, const float (&f)[N]) { *p = f[0]; }
  __device__ static __forceinline__ void add(const Raw& r, float (&acc)[N]) { acc[0] = __fadd_rn(acc[0], r); }
};

template <>
struct Chunk<__nv_bfloat16, false> {
  static constexpr int N = 1;
  using Raw = unsigned short;
  __device__ static __forceinline__ Raw zero() { return 0; }
  __device__ static __forceinline__ Raw load(const __nv_bfloat16* p) {
    return __ldg(reinterpret_cast<const unsigned short*>(p));
  }
  __device__ static __forceinline__ void unpack(const Raw& r, float (&f)[N]) {
    f[0] = __uint_as_float(static_cast<unsigned>(r) << 16);
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&f)[N]) {
    *p = __float2bfloat16_rn(f[0]);
  }
  __device__ static __forceinline__ void add(const Raw& r, float (&acc)[N]) {
    acc[0] = __fadd_rn(acc[0], __uint_as_float(static_cast<unsigned>(r) << 16));
  }
};

template <int G>
__device__ __forceinline__ unsigned group_mask(int lane) {
  if constexpr (G == 32) {
    return 0xffffffffu;
  } else {
    return ((1u << G) - 1u) << (lane & ~(G - 1));
  }
}

__device__ __forceinline__ float leaky(float s, float slope) { return s > 0.f ? s : s * slope; }

// Sum `val` over the lanes that own the same head.  lph = lanes per head inside the group
// (>= G means the whole group is one head).  All lanes of the group call this together.
template <int G>
__device__ __forceinline__ float head_reduce(float val, unsigned gmask, int lph, int gl) {
  if (lph >= G) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) val += __shfl_xor_sync(gmask, val, o, G);
    return val;
  }
  if ((lph & (lph - 1)) == 0) {
    for (int o = lph >> 1; o > 0; o >>= 1) val += __shfl_xor_sync(gmask, val, o, G);
    return val;
  }
  const int myhead = gl / lph;
  float s = 0.f;
  for (int k = 0; k < G; ++k) {
    const float o = __shfl_sync(gmask, val, k, G);
    if (k / lph == myhead) s += o;
  }
  return s;
}

// ---------------------------------------------------------------------------------------------
// AllDeepSets: segmented sum / mean
// ---------------------------------------------------------------------------------------------
// Accumulate rows [first, end) of one segment, `stride` slots between successive batches of G.
template <typename T, bool VECTOR, int G, bool WEIGHTED>
__device__ __forceinline__ void accumulate_rows(const T* __restrict__ x, const int* __restrict__ col,
                                                const float* __restrict__ w,
                                                const float* __restrict__ sscale, int first, int end,
                                                int stride, int d, int feat, bool active, int gl,
                                                unsigned gmask,
                                                float (&acc)[Chunk<T, VECTOR>::N]) {
  using CH = Chunk<T, VECTOR>;
  constexpr int N = CH::N;
  constexpr int U = G < kUnrollSum ? G : kUnrollSum;
  for (int base = first; base < end; base += stride) {
    const int n = min(G, end - base);
    int myidx = 0;
    float myw = 1.f;
    if (gl < n) {
      myidx = __ldg(col + base + gl);
      if (WEIGHTED) {
        if (w != nullptr) myw = __ldg(w + base + gl);
        if (sscale != nullptr) myw *= __ldg(sscale + myidx);
      }
    }
#pragma unroll 1
    for (int u = 0; u < n; u += U) {
      typename CH::Raw raw[U];
      float wk[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int idx = __shfl_sync(gmask, myidx, u + k, G);
        if (WEIGHTED) wk[k] = __shfl_sync(gmask, myw, u + k, G);
        raw[k] = CH::zero();
        if (u + k < n && active) raw[k] = CH::load(x + (size_t)idx * (size_t)d + feat);
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (WEIGHTED) {
          float f[N];
          CH::unpack(raw[k], f);
          // mul then add, each rounded: the reference materialises norm*x_j before the scatter
#pragma unroll
          for (int i = 0; i < N; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(wk[k], f[i]));
        } else {
          CH::add(raw[k], acc);
        }
      }
    }
  }
}

template <typename T, bool VECTOR, int G, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads)
segreduce_group_kernel(const T* __restrict__ x, const int* __restrict__ rowptr,
                       const int* __restrict__ col, const float* __restrict__ w,
                       const float* __restrict__ sscale, long long n_tgt, int d, int mean,
                       int skip_over, T* __restrict__ out) {
  using CH = Chunk<T, VECTOR>;
  constexpr int N = CH::N;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const unsigned gmask = group_mask<G>(lane);
  const long long seg = ((long long)blockIdx.x * kThreads + threadIdx.x) / G;
  if (seg >= n_tgt) return;
  const int feat = (blockIdx.y * 32 + gl) * N;
  const bool active = feat < d;
  const int beg = __ldg(rowptr + seg), end = __ldg(rowptr + seg + 1);
  if (end - beg > skip_over) return;
  float acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = 0.f;
  accumulate_rows<T, VECTOR, G, WEIGHTED>(x, col, w, sscale, beg, end, G, d, feat, active, gl, gmask, acc);
  if (mean) {
    const float cnt = (float)max(end - beg, 1);
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = __fdiv_rn(acc[i], cnt);
  }
  if (active) CH::store(out + (size_t)seg * (size_t)d + feat, acc);
}

template <typename T, bool VECTOR, int G, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads)
segreduce_cta_kernel(const T* __restrict__ x, const int* __restrict__ rowptr,
                     const int* __restrict__ col, const float* __restrict__ w,
                     const float* __restrict__ sscale, const int* __restrict__ long_ids, int d,
                     int mean, T* __restrict__ out) {
  using CH = Chunk<T, VECTOR>;
  constexpr int N = CH::N;
  constexpr int NG = kThreads / G;
  __shared__ float red[kThreads * N];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const int grp = threadIdx.x / G;
  const unsigned gmask = group_mask<G>(lane);
  const long long seg = long_ids[blockIdx.x];
  const int feat = (blockIdx.y * 32 + gl) * N;
  const bool active = feat < d;
  const int beg = __ldg(rowptr + seg), end = __ldg(rowptr + seg + 1);
  float acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = 0.f;
  accumulate_rows<T, VECTOR, G, WEIGHTED>(x, col, w, sscale, beg + grp * G, end, NG * G, d, feat, active,
                                          gl, gmask, acc);
#pragma unroll
  for (int i = 0; i < N; ++i) red[threadIdx.x * N + i] = acc[i];
  __syncthreads();
  if (grp == 0) {
    for (int g = 1; g < NG; ++g) {
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] += red[(g * G + gl) * N + i];
    }
    if (mean) {
      const float cnt = (float)max(end - beg, 1);
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = __fdiv_rn(acc[i], cnt);
    }
    if (active) CH::store(out + (size_t)seg * (size_t)d + feat, acc);
  }
}

// a / c for a small positive integer count c, rc = __frcp_rn(c): one Newton correction of a * rc gives the
// correctly rounded quotient (Markstein) without the slow-path call of the generic fp32 division.
__device__ __forceinline__ float div_count(float a, float c, float rc) {
  const float q = a * rc;
  const float r = fmaf(-q, c, a);
  const float res = fmaf(r, rc, q);
  return fabsf(q) == INFINITY ? q : res;
}

// ---------------------------------------------------------------------------------------------
// STREAM kernels: rows staged through shared memory by the TMA (cp.async.bulk), one row per copy
// ---------------------------------------------------------------------------------------------
// Each warp owns a contiguous block of target segments, i.e. ONE contiguous range [kb, ke) of the CSR, and runs its
// own producer/consumer ring in shared memory:
//   produce : 32 lanes read 32 column ids with one coalesced load and each lane issues ONE bulk copy
//             (cp.async.bulk global -> shared, `row_bytes` each) that completes on the stage's mbarrier.  Bytes in
//             flight live in shared memory, not in registers: (stages-1) * 8 KB per warp with ~40 registers/thread.
//   consume : after the mbarrier flips, the whole warp reads one staged row per LDS (lane = 1/32 of the row,
//             conflict-free), adds it into fp32 accumulators (FHADD.BF16 for bf16 rows) and flushes the accumulator
//             to `out` whenever the flat stream crosses a segment boundary.  Control flow is warp-uniform.
// Per row this costs ~1/32 copy instruction + 1 LDS + (row_bytes/64) adds per warp instead of 16 LDG + shuffles +
// unpack per lane group, and it removes the per-segment rowptr -> col -> row dependency chain that makes short
// segments (E->V: ~6 incidences per vertex; real data: size-1 self-loop hyperedges) latency-bound.
#ifndef ALLSET_STREAM_WARPS
#define ALLSET_STREAM_WARPS 24
#endif
#ifndef ALLSET_STREAM_STAGE_BYTES
#define ALLSET_STREAM_STAGE_BYTES 4096
#endif
#ifndef ALLSET_STREAM_SMEM
#define ALLSET_STREAM_SMEM (192 * 1024)
#endif
constexpr int kStreamWarps = ALLSET_STREAM_WARPS;
constexpr int kStreamStageBytes = ALLSET_STREAM_STAGE_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#if defined(ALLSET_STREAM_COPY) && ALLSET_STREAM_COPY == 1
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
#endif
#ifndef ALLSET_L2_HINTS
#define ALLSET_L2_HINTS 1
#endif
// L2 eviction policies (PMA kernel only; measured neutral-to-negative for the plain sum kernel): gathered feature
// rows are streamed (evict_first) so that small, heavily re-read side records
// (the per-row attention scores: |E| * H * 4 B = 64 MB in E->V, which fits the 126 MB L2) stay resident (evict_last).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <bool HINT>
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint64_t policy) {
  if (HINT && ALLSET_L2_HINTS)
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(policy) : "memory");
  else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// arrive on `bar` once every cp.async this thread has issued so far has landed (does not add to the pending count)
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// Stage `n` rows of `rowb` bytes: lane r holds the source row id of row r.  ALLSET_STREAM_COPY == 0: 16-byte
// cp.async (LDGSTS) per lane, 512 contiguous shared bytes per warp instruction -- ~3 instructions per 256-byte row;
// == 1: one cp.async.bulk (UBLKCP) per row.  UBLKCP takes its operands from UNIFORM registers, so the compiler
// serialises per-lane bulk copies into a 9-instruction loop per row (measured: 36 % of all instructions of the sum
// kernel, profiles/r01_stream_bulk_*.md), which is why LDGSTS is the default.
#ifndef ALLSET_STREAM_COPY
#define ALLSET_STREAM_COPY 0
#endif
constexpr int kStreamBarCount = ALLSET_STREAM_COPY == 0 ? 32 : 1;
template <int ROWB, bool HINT>
__device__ __forceinline__ void stage_rows(uint32_t dst, const unsigned char* __restrict__ src, int my_row_id, int n,
                                           uint32_t bar, int lane, uint64_t policy, size_t pitch = ROWB) {
#if ALLSET_STREAM_COPY == 0
  constexpr int CPR = ROWB / 16;                      // 16-byte chunks per row
  const int total = n * CPR;
#pragma unroll 4
  for (int q0 = 0; q0 < total; q0 += 32) {
    const int q = q0 + lane;
    const int row = q / CPR, ch = q % CPR;
    const int idx = __shfl_sync(0xffffffffu, my_row_id, row & 31);
    if (q < total) cp_async16<HINT>(dst + (uint32_t)q * 16u, src + (size_t)idx * pitch + (size_t)ch * 16, policy);
  }
#else
  if (lane < n) bulk_g2s(dst + (uint32_t)lane * ROWB, src + (size_t)my_row_id * pitch, ROWB, bar);
#endif
}
// same for a small per-row side record of `sb` bytes (sb % 16 == 0)
__device__ __forceinline__ void stage_side(uint32_t dst, const unsigned char* __restrict__ src, int my_row_id, int n,
                                           uint32_t sb, uint32_t bar, int lane, uint64_t policy, size_t pitch) {
#if ALLSET_STREAM_COPY == 0
  const int cpr = (int)(sb / 16u);
  const int total = n * cpr;
  for (int q0 = 0; q0 < total; q0 += 32) {
    const int q = q0 + lane;
    const int row = q / cpr, ch = q % cpr;
    const int idx = __shfl_sync(0xffffffffu, my_row_id, row & 31);
    if (q < total) cp_async16<true>(dst + (uint32_t)q * 16u, src + (size_t)idx * pitch + (size_t)ch * 16, policy);
  }
#else
  if (lane < n) bulk_g2s(dst + (uint32_t)lane * sb, src + (size_t)my_row_id * pitch, sb, bar);
#endif
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// Work split of the stream kernels: warp `chunk` of `n_chunks` owns the target segments [first, first + nseg) such that
// every chunk carries (nearly) the same COST = incidences gathered + segments flushed, cost(s) = rowptr[s] + s.  The
// boundary is found on the fly by a warp-cooperative 32-ary search over rowptr (~5 dependent loads), so skewed
// (power-law) segment sizes do not unbalance the warps and no partition array has to be built or passed.
__device__ __forceinline__ long long chunk_boundary(const int* __restrict__ rowptr, long long n_tgt, long long total,
                                                    long long chunk, long long n_chunks, int lane) {
  if (chunk <= 0) return 0;
  if (chunk >= n_chunks) return n_tgt;
  const long long target = total * chunk / n_chunks;      // total <= 2^33, chunk < 2^20: no overflow
  if (target <= 0) return 0;
  // cost is strictly increasing; invariant cost(lo) < target <= cost(hi); the answer is the smallest s with
  // cost(s) >= target, i.e. hi once hi - lo == 1
  long long lo = 0, hi = n_tgt;
  while (hi - lo > 1) {
    const long long stride = (hi - lo + 31) / 32;
    long long sj = lo + (long long)(lane + 1) * stride;   // lane 31 (and possibly more) probes hi itself
    if (sj > hi) sj = hi;
    const bool ge = (long long)__ldg(rowptr + sj) + sj >= target;
    const int first = __ffs(__ballot_sync(0xffffffffu, ge)) - 1;
    long long nhi = lo + (long long)(first + 1) * stride;
    if (nhi > hi) nhi = hi;
    if (first > 0) lo = lo + (long long)first * stride;
    hi = nhi;
  }
  return hi;
}

// Long segments.  A chunk boundary that falls inside a segment of at least kSplitMinLen incidences CUTS it (when both
// pieces keep kSplitMinPiece incidences), so that a 4096-row hyperedge of a power-law graph is reduced by as many warps
// as its share of the work instead of serialising one warp (measured before: 5.0-5.8 ms with 4096-row segments inside
// one warp's stream vs 4.3 ms for bucketed CTA kernels).  Pieces are combined without atomics, in stream order:
//   * a chunk has at most ONE non-final piece, at its end: its partial state goes to workspace slot [chunk] and a flag
//     is released (1 = first piece of the segment, 2 = continuing piece);
//   * a chunk has at most one FINAL piece, at its start: the warp parks that state in its own `head` slot, finishes the
//     rest of its stream, then walks back over the flags of the preceding chunks (decoupled look-back: they run on the
//     same or earlier CTAs), adds their partials in ascending order, writes the row and clears the flags it consumed.
// The workspace (allset_stream_workspace_bytes, zero-initialised once by the caller) is therefore left zeroed.
constexpr int kSplitMinLen = 256;
constexpr int kSplitMinPiece = 32;
constexpr long long kStreamMaxChunks = 148LL * kStreamWarps * 8;

struct Cut {
  long long seg;   // first segment that STARTS at or after the cut (inside == 0) / the segment being cut (inside == 1)
  int k;           // stream position of the cut
  int inside;
};

__device__ __forceinline__ Cut chunk_cut(const int* __restrict__ rowptr, long long n_tgt, long long total,
                                         long long chunk, long long n_chunks, int lane, bool split) {
  Cut c;
  c.seg = chunk_boundary(rowptr, n_tgt, total, chunk, n_chunks, lane);
  c.k = __ldg(rowptr + c.seg);
  c.inside = 0;
  if (split && chunk > 0 && chunk < n_chunks && c.seg > 0) {
    const long long target = total * chunk / n_chunks;
    const long long s = c.seg - 1;
    const int beg = __ldg(rowptr + s);
    const long long o = target - ((long long)beg + s);      // 0 < o <= len + 1: where the target falls inside segment s
    const long long len = (long long)c.k - beg;
    // Every target maps to a position inside ITS OWN segment's closed range [beg, end], monotonically in the target, so
    // consecutive cuts can never cross: short segments round up to their end (as without a workspace); a long segment
    // rounds DOWN to its start when the target is within kSplitMinPiece of it, is cut in the middle, and rounds up near
    // its end.  (Rounding a long segment's early target UP would hand the whole segment to the previous chunk while the
    // next target cuts it in the middle -- the next chunk would then wait for a first piece nobody publishes.)
    if (len >= kSplitMinLen) {
      if (o < kSplitMinPiece) {
        c.seg = s;
        c.k = beg;
      } else if (len - o >= kSplitMinPiece) {
        c.seg = s;
        c.k = beg + (int)o;
        c.inside = 1;
      }
    }
  }
  return c;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct StreamWs {           // views into the caller's workspace
  int* status;              // [4]: word 0 != 0 if a look-back wait ever timed out (diagnostic; never in a correct run)
  int* flags;               // [kStreamMaxChunks]
  float* tail;              // [kStreamMaxChunks][stride]: published non-final pieces
  float* head;              // [kStreamMaxChunks][stride]: a warp's own parked final piece
};
__device__ __forceinline__ StreamWs stream_ws(void* ws, int stride) {
  StreamWs v;
  v.status = reinterpret_cast<int*>(ws);
  v.flags = v.status + 4;
  v.tail = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ws) + 16 + kStreamMaxChunks * 4);
  v.head = v.tail + kStreamMaxChunks * (long long)stride;
  return v;
}
// Wait for the flag of a preceding chunk.  Bounded: a logic error must flag the status word, not hang the GPU.
__device__ __forceinline__ int wait_piece_flag(const StreamWs& wsv, long long q, int lane) {
  int f = 0;
  for (int spins = 0; spins < (1 << 24); ++spins) {
    f = ld_acquire_gpu(wsv.flags + q);
    if (f != 0) return f;
  }
  if (lane == 0) atomicExch(wsv.status, 1);
  return 1;                 // give up: treat it as the first piece
}

// One lane's share of a staged row: LB bytes (LB = row_bytes / 32).  LB >= 16 is read as LB/16 16-byte chunks,
// chunk c of lane l at byte (c*32 + l)*16, so every LDS.128 of the warp covers 512 contiguous bytes.
template <typename T, int LB>
struct LaneRow {
  static constexpr int ES = sizeof(T);
  static constexpr int NA = LB / ES;                 // fp32 accumulators per lane
  static constexpr int CH = LB >= 16 ? LB / 16 : 1;  // 16-byte chunks per lane
  static constexpr int CB = LB >= 16 ? 16 : LB;      // bytes per chunk
  static constexpr int EPC = CB / ES;                // elements per chunk
  __device__ static __forceinline__ int offset(int lane, int c) { return (c * 32 + lane) * CB; }
  // acc += row (fp32 accumulate); `p` = shared-space byte address of the row
  __device__ static __forceinline__ void add(uint32_t p, int lane, float (&acc)[NA]) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      uint32_t r[4] = {0u, 0u, 0u, 0u};
      const uint32_t a = p + offset(lane, c);
      if (CB == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
      else if (CB == 8) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
      else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r[0]) : "r"(a));
#pragma unroll
      for (int q = 0; q < CB / 4; ++q) {
        if (ES == 2) Chunk<__nv_bfloat16, true>::add2(r[q], acc[c * EPC + 2 * q], acc[c * EPC + 2 * q + 1]);
        else acc[c * EPC + q] = __fadd_rn(acc[c * EPC + q], __uint_as_float(r[q]));
      }
    }
  }
  // acc += wk * row, product and sum rounded separately (the reference materialises norm * x_j first)
  __device__ static __forceinline__ void add_weighted(uint32_t p, int lane, float wk, float (&acc)[NA]) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      uint32_t r[4] = {0u, 0u, 0u, 0u};
      const uint32_t a = p + offset(lane, c);
      if (CB == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
      else if (CB == 8) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
      else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r[0]) : "r"(a));
#pragma unroll
      for (int q = 0; q < CB / 4; ++q) {
        if (ES == 2) {
          const float lo = __uint_as_float(r[q] << 16), hi = __uint_as_float(r[q] & 0xffff0000u);
          acc[c * EPC + 2 * q] = __fadd_rn(acc[c * EPC + 2 * q], __fmul_rn(wk, lo));
          acc[c * EPC + 2 * q + 1] = __fadd_rn(acc[c * EPC + 2 * q + 1], __fmul_rn(wk, hi));
        } else {
          acc[c * EPC + q] = __fadd_rn(acc[c * EPC + q], __fmul_rn(wk, __uint_as_float(r[q])));
        }
      }
    }
  }
  __device__ static __forceinline__ void store(T* row, int lane, const float (&acc)[NA]) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      unsigned char* dst = reinterpret_cast<unsigned char*>(row) + offset(lane, c);
      if (ES == 2) {
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < CB / 4; ++q) {
          __nv_bfloat162 v = __floats2bfloat162_rn(acc[c * EPC + 2 * q], acc[c * EPC + 2 * q + 1]);
          o[q] = *reinterpret_cast<uint32_t*>(&v);
        }
        // streaming stores (st.global.cs): output rows are written once and not re-read by this kernel
        if (CB == 16) __stcs(reinterpret_cast<uint4*>(dst), make_uint4(o[0], o[1], o[2], o[3]));
        else if (CB == 8) __stcs(reinterpret_cast<uint2*>(dst), make_uint2(o[0], o[1]));
        else __stcs(reinterpret_cast<unsigned int*>(dst), o[0]);
      } else {
        if (CB == 16) __stcs(reinterpret_cast<float4*>(dst), make_float4(acc[c * EPC], acc[c * EPC + 1], acc[c * EPC + 2], acc[c * EPC + 3]));
        else if (CB == 8) __stcs(reinterpret_cast<float2*>(dst), make_float2(acc[c * EPC], acc[c * EPC + 1]));
        else __stcs(reinterpret_cast<float*>(dst), acc[c * EPC]);
      }
    }
  }
  // the same row image into shared memory (`p` = shared-space byte address): staging for bulk stores
  __device__ static __forceinline__ void store_shared(uint32_t p, int lane, const float (&acc)[NA]) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const uint32_t a = p + offset(lane, c);
      uint32_t o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int q = 0; q < CB / 4; ++q) {
        if (ES == 2) {
          __nv_bfloat162 v = __floats2bfloat162_rn(acc[c * EPC + 2 * q], acc[c * EPC + 2 * q + 1]);
          o[q] = *reinterpret_cast<uint32_t*>(&v);
        } else {
          o[q] = __float_as_uint(acc[c * EPC + q]);
        }
      }
      if (CB == 16) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
      else if (CB == 8) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(o[0]), "r"(o[1]) : "memory");
      else asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(o[0]) : "memory");
    }
  }
};

// Fused exchange: besides `out`, every reduced row is also stored into the same row of up to 7 peer replicas
// (peer-mapped device pointers over NVLink, or ONE multicast address the NVSwitch replicates), so the all-gather that
// would follow the kernel rides on its epilogue.  delta[j] = byte distance from `out` row 0 to the same row in peer j's
// buffer.  `mask` (optional, one byte per target row) selects the peers that need the row: bit j = peer j -- a vertex
// row is needed only by the ranks whose hyperedge range touches the vertex.
//   PUSH == 1: the reducing warp stores the row to every selected peer itself (st.global, fire-and-forget until the
//              NVLink queue back-pressures the warp);
//   PUSH == 2: the row is parked in a 2-slot shared-memory staging buffer and ONE lane hands it to the TMA
//              (cp.async.bulk shared -> global, one bulk store per selected peer), so the warp goes back to gathering
//              while the copies drain; a slot is reused after cp.async.bulk.wait_group.read.
struct PeerOuts {
  long long delta[7];
  int n;
};

// static indexing only (a dynamically indexed by-value struct would be copied to local memory)
template <typename T, typename LR, int NA>
__device__ __forceinline__ void store_to_peers(T* row, const PeerOuts& peers, unsigned mask, int lane,
                                               const float (&acc)[NA]) {
#pragma unroll
  for (int j = 0; j < 7; ++j)
    if (j < peers.n && ((mask >> j) & 1u))
      LR::store(reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(row) + peers.delta[j]), lane, acc);
}

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <typename T, typename LR, int NA, int ROWB>
__device__ __forceinline__ void bulk_to_peers(T* row, const PeerOuts& peers, unsigned mask, int lane, uint32_t stage,
                                              int& nslot, const float (&acc)[NA]) {
  const uint32_t slot = stage + (uint32_t)(nslot & 1) * ROWB;
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the copies that read this slot are done
  __syncwarp();
  LR::store_shared(slot, lane, acc);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j < peers.n && ((mask >> j) & 1u)) bulk_s2g(reinterpret_cast<unsigned char*>(row) + peers.delta[j], slot, ROWB);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  ++nslot;
}

template <typename T, int LB, bool WEIGHTED, int PUSH>
__global__ void __launch_bounds__(kStreamWarps * 32, 1)
segreduce_stream_kernel(const T* __restrict__ x, const int* __restrict__ rowptr, const int* __restrict__ col,
                        const float* __restrict__ w, const float* __restrict__ sscale, long long n_tgt, int d,
                        int mean, int seg_per_warp, int stages, T* __restrict__ out, PeerOuts peers,
                        const unsigned char* __restrict__ peer_mask, void* ws) {
  using LR = LaneRow<T, LB>;
  constexpr int NA = LR::NA;
  constexpr int ROWB = LB * 32;
  constexpr int RPS = (kStreamStageBytes / ROWB) < 32 ? (kStreamStageBytes / ROWB) : 32;   // rows per stage
  constexpr int PS = NA * 32;                                                            // floats per parked piece
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // seg_per_warp only sizes the grid (host); the actual block of segments is cost-balanced
  const long long n_chunks = (n_tgt + seg_per_warp - 1) / seg_per_warp;
  const long long chunk = (long long)blockIdx.x * kStreamWarps + warp;
  if (chunk >= n_chunks) return;                      // warp-uniform; nothing below synchronises across warps
  const long long total_cost = (long long)__ldg(rowptr + n_tgt) + n_tgt;
  const bool split = ws != nullptr && n_chunks <= kStreamMaxChunks;
  const Cut c0 = chunk_cut(rowptr, n_tgt, total_cost, chunk, n_chunks, lane, split);
  const Cut c1 = chunk_cut(rowptr, n_tgt, total_cost, chunk + 1, n_chunks, lane, split);
  const long long s_first = c0.seg;
  const int nseg = (int)(c1.seg - s_first);            // segments that END inside this chunk
  const int kb = c0.k, ke = c1.k;
  if (nseg <= 0 && ke <= kb) return;
  const uint32_t ring = smem_u32(smem) + (uint32_t)warp * (uint32_t)stages * (RPS * ROWB);
  const uint32_t bars = smem_u32(smem) + (uint32_t)kStreamWarps * (uint32_t)stages * (RPS * ROWB) +
                        (uint32_t)warp * (uint32_t)stages * 8u;
  float* wbuf = reinterpret_cast<float*>(smem + (size_t)kStreamWarps * stages * (RPS * ROWB) +
                                         (size_t)kStreamWarps * stages * 8) + (size_t)warp * stages * RPS;
  // staging rows of the bulk-store exchange: behind the ring, the barriers and the (optional) weight buffer
  const uint32_t stage_out = smem_u32(smem) + (uint32_t)kStreamWarps * (uint32_t)stages * (RPS * ROWB + 8 + (WEIGHTED ? RPS * 4 : 0)) +
                             16u + (uint32_t)warp * 2u * ROWB;
  int nslot = 0;
  if (lane < stages) mbar_init(bars + lane * 8, kStreamBarCount);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  const int* __restrict__ rp = rowptr + s_first;
  const StreamWs wsv = stream_ws(ws, PS);
  float acc[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.f;
  // segment cursor: end of the current segment and (prefetched, warp-uniform load) of the next one; the tail piece of a
  // cut segment (seg == nseg) never ends inside the stream
  int seg = 0, cur_beg = kb, pos = kb;
  int cur_end = nseg > 0 ? __ldg(rp + 1) : INT32_MAX, nxt_end = __ldg(rp + min(2, max(nseg, 0)));

  // a finished row: scale (mean), store, exchange
  auto emit = [&](T* row, long long s_idx, int cnt_i, float (&val)[NA]) {
    if (mean) {
      const float cnt = (float)max(cnt_i, 1);
      const float rc = __frcp_rn(cnt);
#pragma unroll
      for (int i = 0; i < NA; ++i) val[i] = div_count(val[i], cnt, rc);
    }
    LR::store(row, lane, val);
    if (PUSH != 0) {
      const unsigned m = peer_mask != nullptr ? (unsigned)__ldg(peer_mask + s_idx) : 0xffu;
      if (PUSH == 1) store_to_peers<T, LR, NA>(row, peers, m, lane, val);
      else bulk_to_peers<T, LR, NA, ROWB>(row, peers, m, lane, stage_out, nslot, val);
    }
  };
  T* __restrict__ ob = out + (size_t)s_first * (size_t)d;           // output row of the current segment
  bool head_pending = c0.inside != 0;                                // the first flush parks a cut segment's final piece
  auto flush = [&]() {
    if (head_pending) {
      // final piece of a segment that began in an earlier chunk: park it, combine after the stream (see below)
#pragma unroll
      for (int i = 0; i < NA; ++i) wsv.head[chunk * PS + i * 32 + lane] = acc[i];
      head_pending = false;
    } else {
      emit(ob, s_first + seg, cur_end - cur_beg, acc);
    }
    ob += d;
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.f;
    ++seg;
    cur_beg = cur_end;
    cur_end = seg < nseg ? nxt_end : INT32_MAX;
    nxt_end = __ldg(rp + min(seg + 2, max(nseg, 0)));
  };

  // ---- producer side ---------------------------------------------------------------------------------------
  int ipos = kb;                       // next stream position to issue
  int pf_idx = 0;                      // column id of stream position ipos + lane, loaded one step ahead
  float pf_w = 1.f;
  auto prefetch = [&]() {
    pf_idx = 0;
    pf_w = 1.f;
    if (lane < RPS && ipos + lane < ke) {
      pf_idx = __ldg(col + ipos + lane);
      if (WEIGHTED) {
        if (w != nullptr) pf_w = __ldg(w + ipos + lane);
        if (sscale != nullptr) pf_w *= __ldg(sscale + pf_idx);
      }
    }
  };
  auto issue = [&](int st) {
    const int n = min(RPS, ke - ipos);
    const uint32_t bar = bars + st * 8;
#if ALLSET_STREAM_COPY == 1
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)n * ROWB);
    __syncwarp();
#endif
    stage_rows<ROWB, false>(ring + (uint32_t)st * (RPS * ROWB), reinterpret_cast<const unsigned char*>(x), pf_idx, n, bar, lane, 0ull);
    if (WEIGHTED && lane < n) wbuf[st * RPS + lane] = pf_w;
#if ALLSET_STREAM_COPY == 0
    cp_async_arrive(bar);
#endif
    ipos += RPS;
    prefetch();
  };
  prefetch();
  for (int st = 0; st < stages && ipos < ke; ++st) issue(st);

  // ---- consumer side ---------------------------------------------------------------------------------------
  // invariant: pos < cur_end whenever a row is consumed (empty segments are flushed as soon as they are reached)
  while (pos >= cur_end) flush();
  int st = 0;
  uint32_t phase = 0;
  for (int cbase = kb; cbase < ke; cbase += RPS) {
    const int n = min(RPS, ke - cbase);
    mbar_wait(bars + st * 8, phase);
    if (WEIGHTED) __syncwarp();
    const uint32_t rows = ring + (uint32_t)st * (RPS * ROWB);
    const float* wrow = wbuf + st * RPS;
    if (PUSH == 0 && n == RPS) {
      // full stage: fully unrolled, one compare per row against the (stage-relative) end of the current segment
      // (the fused-exchange variants keep a single flush site: their epilogue stores to up to 8 replicas)
      int rel = cur_end - cbase;
#pragma unroll
      for (int r = 0; r < RPS; ++r) {
        if (WEIGHTED) LR::add_weighted(rows + r * ROWB, lane, wrow[r], acc);
        else LR::add(rows + r * ROWB, lane, acc);
        if (rel == r + 1) {
          do {                                   // the segment ends with this row (then skip empty segments)
            flush();
            rel = cur_end - cbase;
          } while (rel == r + 1);
        }
      }
      pos += RPS;
    } else {
      for (int r = 0; r < n; ++r) {
        if (WEIGHTED) LR::add_weighted(rows + r * ROWB, lane, wrow[r], acc);
        else LR::add(rows + r * ROWB, lane, acc);
        ++pos;
        while (pos >= cur_end) flush();
      }
    }
    __syncwarp();
    if (ipos < ke) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of this stage before the async refill
      issue(st);
    }
    if (++st == stages) {
      st = 0;
      phase ^= 1u;
    }
  }
  while (seg < nseg) flush();

  // ---- pieces of cut segments ------------------------------------------------------------------------------
  if (c1.inside) {
    // non-final piece at the end of the stream: publish it for the warp that owns the end of the segment
#pragma unroll
    for (int i = 0; i < NA; ++i) wsv.tail[chunk * PS + i * 32 + lane] = acc[i];
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_gpu(wsv.flags + chunk, (c0.inside && nseg == 0) ? 2 : 1);
  }
  if (c0.inside && nseg > 0) {
    long long q0 = chunk - 1;
    while (wait_piece_flag(wsv, q0, lane) != 1 && q0 > 0) --q0;      // walk back to the chunk holding the first piece
    float tot[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) tot[i] = 0.f;
    for (long long q = q0; q < chunk; ++q) {
#pragma unroll
      for (int i = 0; i < NA; ++i) tot[i] = __fadd_rn(tot[i], __ldcg(wsv.tail + q * PS + i * 32 + lane));
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) tot[i] = __fadd_rn(tot[i], wsv.head[chunk * PS + i * 32 + lane]);
    __syncwarp();
    if (lane == 0)
      for (long long q = q0; q < chunk; ++q) wsv.flags[q] = 0;      // leave the workspace zeroed for the next launch
    emit(out + (size_t)s_first * (size_t)d, s_first, __ldg(rp + 1) - __ldg(rp), tot);
  }
  if (PUSH == 2) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
}

// The exchange by itself: rows [0, n_rows) of this rank's range (already in its own replica) -> the same rows of the
// selected peers.  Used where the producer of the rows is not one of the fused-exchange kernels (a cuBLAS GEMM epilogue,
// a rowop pass): one read of the rows, up to 7 NVLink stores per 16-byte chunk, 4 chunks in flight per thread.
__global__ void __launch_bounds__(256)
push_rows_kernel(const uint4* __restrict__ src, long long n_chunks, int chunks_per_row, PeerOuts peers,
                 const unsigned char* __restrict__ mask) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; g0 < n_chunks; g0 += 4 * stride) {
    uint4 v[4];
    unsigned m[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long g = g0 + u * stride;
      m[u] = 0u;
      if (g < n_chunks) {
        v[u] = __ldcs(src + g);
        m[u] = mask != nullptr ? (unsigned)__ldg(mask + g / chunks_per_row) : 0xffu;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long g = g0 + u * stride;
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (j < peers.n && ((m[u] >> j) & 1u))
          *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(const_cast<uint4*>(src + g)) + peers.delta[j]) = v[u];
    }
  }
}

// d{0,1} = a * b{0,1} + d{0,1}: one FFMA2 (sm_100 packed fp32, both halves rounded like scalar FFMA)
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float b0, float b1) {
  asm("{\n\t.reg .b64 pa, pb, pc;\n\tmov.b64 pa, {%2, %2};\n\tmov.b64 pb, {%3, %4};\n\tmov.b64 pc, {%0, %1};\n\t"
      "fma.rn.f32x2 pc, pa, pb, pc;\n\tmov.b64 {%0, %1}, pc;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a) {
  asm("{\n\t.reg .b64 pa, pc;\n\tmov.b64 pa, {%2, %2};\n\tmov.b64 pc, {%0, %1};\n\t"
      "mul.rn.f32x2 pc, pc, pa;\n\tmov.b64 {%0, %1}, pc;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kPmaOnlinePiece = 8;   // pieces up to this many rows take the single-pass online-softmax path

// PMA over the same TMA ring: every staged item is a value row (row_bytes) plus the H fp32 scores of that source row
// (a second bulk copy onto the same mbarrier).  The online softmax runs per lane-chunk in the log2 domain
// (a2 = leaky_relu(score) * log2(e); p = 2^(a2 - m2)), state (m2, l, acc) is flushed at segment boundaries:
// out = acc / (l + 1e-16) + seed  (PyG softmax then scatter-add then += att_r, reference layers.py:168-194,153).
template <typename T, int LB, int PUSH>
__global__ void __launch_bounds__(kStreamWarps * 32, 1)
pma_stream_kernel(const T* __restrict__ v, const float* __restrict__ score, const float* __restrict__ seed,
                  const int* __restrict__ rowptr, const int* __restrict__ col, long long n_tgt, int H, int C,
                  float slope, int seg_per_warp, int stages, T* __restrict__ out, float* __restrict__ stats,
                  PeerOuts peers, long long v_pitch, long long s_pitch, const unsigned char* __restrict__ peer_mask,
                  void* ws) {
  using LR = LaneRow<T, LB>;
  constexpr int ES = LR::ES, NA = LR::NA, CH = LR::CH, CB = LR::CB, EPC = LR::EPC;
  constexpr int ROWB = LB * 32;
  constexpr int RPS = (kStreamStageBytes / ROWB) < 32 ? (kStreamStageBytes / ROWB) : 32;
  constexpr int PS = (NA + 2 * CH) * 32;                      // floats per parked piece: acc, then m2, then l
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_chunks = (n_tgt + seg_per_warp - 1) / seg_per_warp;
  const long long chunk = (long long)blockIdx.x * kStreamWarps + warp;
  if (chunk >= n_chunks) return;
  const long long total_cost = (long long)__ldg(rowptr + n_tgt) + n_tgt;
  const bool split = ws != nullptr && n_chunks <= kStreamMaxChunks;
  const Cut c0 = chunk_cut(rowptr, n_tgt, total_cost, chunk, n_chunks, lane, split);
  const Cut c1 = chunk_cut(rowptr, n_tgt, total_cost, chunk + 1, n_chunks, lane, split);
  const long long s_first = c0.seg;
  const int nseg = (int)(c1.seg - s_first);                   // segments that END inside this chunk
  const int kb = c0.k, ke = c1.k;
  if (nseg <= 0 && ke <= kb) return;
  const int d = H * C;
  const uint32_t SB = (uint32_t)H * 4u;                                       // score bytes per row
  const uint32_t ring = smem_u32(smem) + (uint32_t)warp * (uint32_t)stages * (RPS * ROWB);
  const uint32_t sring = smem_u32(smem) + (uint32_t)kStreamWarps * (uint32_t)stages * (RPS * ROWB) +
                         (uint32_t)warp * (uint32_t)stages * RPS * SB;
  const uint32_t bars = smem_u32(smem) + (uint32_t)kStreamWarps * (uint32_t)stages * RPS * (ROWB + SB) +
                        (uint32_t)warp * (uint32_t)stages * 8u;
  const uint32_t stage_out = smem_u32(smem) + (uint32_t)kStreamWarps * (uint32_t)stages * (RPS * (ROWB + SB) + 8u) + 16u +
                             (uint32_t)warp * 2u * ROWB;
  int nslot = 0;
  if (lane < stages) mbar_init(bars + lane * 8, kStreamBarCount);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  const int* __restrict__ rp = rowptr + s_first;
  const StreamWs wsv = stream_ws(ws, PS);

  int hc[CH];                                   // head of each of this lane's chunks
  float sd[NA];                                 // seed slice of this lane
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int f0 = LR::offset(lane, c) / ES;
    hc[c] = f0 / C;
#pragma unroll
    for (int i = 0; i < EPC; ++i) sd[c * EPC + i] = __ldg(seed + f0 + i);
  }
  float m2[CH], l[CH], acc[NA];
#pragma unroll
  for (int c = 0; c < CH; ++c) { m2[c] = -INFINITY; l[c] = 0.f; }
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.f;
  bool fresh = true;                            // no piece of the current segment accumulated yet (warp-uniform)
  int seg = 0, pos = kb;
  int cur_end = nseg > 0 ? __ldg(rp + 1) : INT32_MAX, nxt_end = __ldg(rp + min(2, max(nseg, 0)));

  // a finished segment: out = acc / (l + 1e-16) + seed, statistics, exchange
  auto emit = [&](T* row, float* sb_out, long long s_idx, const float (&mm)[CH], const float (&ll)[CH],
                  const float (&aa)[NA]) {
    float o[NA];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const float denom = ll[c] + 1e-16f;
      const float inv = 1.0f / denom;
#pragma unroll
      for (int i = 0; i < EPC; ++i) o[c * EPC + i] = fmaf(aa[c * EPC + i], inv, sd[c * EPC + i]);
      if (sb_out != nullptr && (LR::offset(lane, c) / ES) % C == 0) {
        sb_out[hc[c] * 2 + 0] = mm[c] * kLn2;
        sb_out[hc[c] * 2 + 1] = denom;
      }
    }
    LR::store(row, lane, o);
    if (PUSH != 0) {
      const unsigned m = peer_mask != nullptr ? (unsigned)__ldg(peer_mask + s_idx) : 0xffu;
      if (PUSH == 1) store_to_peers<T, LR, NA>(row, peers, m, lane, o);
      else bulk_to_peers<T, LR, NA, ROWB>(row, peers, m, lane, stage_out, nslot, o);
    }
  };
  T* __restrict__ ob = out + (size_t)s_first * (size_t)d;
  float* __restrict__ sbo = stats != nullptr ? stats + (size_t)s_first * H * 2 : nullptr;
  bool head_pending = c0.inside != 0;
  auto park = [&](float* slot) {                 // (acc, m2, l) of a piece of a cut segment
#pragma unroll
    for (int i = 0; i < NA; ++i) slot[i * 32 + lane] = acc[i];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      slot[(NA + c) * 32 + lane] = m2[c];
      slot[(NA + CH + c) * 32 + lane] = l[c];
    }
  };
  auto flush = [&]() {
    if (head_pending) {                       // final piece of a cut segment: combined after the stream
      park(wsv.head + chunk * PS);
      head_pending = false;
    } else {
      emit(ob, sbo, s_first + seg, m2, l, acc);
    }
    ob += d;
    if (sbo != nullptr) sbo += 2 * H;
#pragma unroll
    for (int c = 0; c < CH; ++c) { m2[c] = -INFINITY; l[c] = 0.f; }
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.f;
    fresh = true;
    ++seg;
    cur_end = seg < nseg ? nxt_end : INT32_MAX;
    nxt_end = __ldg(rp + min(seg + 2, max(nseg, 0)));
  };

  // acc += p * row chunk c (unpack to fp32, packed FFMA2)
  auto fma_chunk = [&](int c, const uint32_t (&r)[4], float p) {
#pragma unroll
    for (int q = 0; q < CB / 4; ++q) {
      if (ES == 2) {
        ffma2(acc[c * EPC + 2 * q], acc[c * EPC + 2 * q + 1], p, __uint_as_float(r[q] << 16),
              __uint_as_float(r[q] & 0xffff0000u));
      } else {
        acc[c * EPC + q] = fmaf(p, __uint_as_float(r[q]), acc[c * EPC + q]);
      }
    }
  };
  auto load_chunk = [&](uint32_t rowaddr, int c, uint32_t (&r)[4]) {
    const uint32_t a = rowaddr + LR::offset(lane, c);
    if (CB == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
    else if (CB == 8) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
    else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r[0]) : "r"(a));
  };
  // a2 (already leaky_relu'ed and scaled by log2 e in place, see the consumer loop) of row `scoreaddr`, head of chunk c
  auto load_a2 = [&](uint32_t scoreaddr, int c) {
    float a2;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a2) : "r"(scoreaddr + (uint32_t)hc[c] * 4u));
    return a2;
  };

  auto short_piece = [&](auto nc, int c, uint32_t prow, uint32_t psc) {
    constexpr int N = decltype(nc)::value;
    float a[N];
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] = load_a2(psc + (uint32_t)k * SB, c);
    float pm = a[0];
#pragma unroll
    for (int k = 1; k < N; ++k) pm = fmaxf(pm, a[k]);
    float mn = pm, lc = 0.f;
    if (!fresh) {                               // one rescale of the carried state
      mn = fmaxf(m2[c], pm);
      const float cf = ex2_approx(m2[c] - mn);
      lc = l[c] * cf;
#pragma unroll
      for (int i = 0; i < EPC; i += 2) {
        if (EPC >= 2) fmul2(acc[c * EPC + i], acc[c * EPC + i + 1], cf);
        else acc[c * EPC + i] *= cf;
      }
    }
    m2[c] = mn;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      uint32_t r0[4];
      load_chunk(prow + (uint32_t)k * ROWB, c, r0);
      const float p0 = ex2_approx(a[k] - mn);
      lc += p0;
      fma_chunk(c, r0, p0);
    }
    l[c] = lc;
  };

  const uint64_t pol_rows = l2_policy_evict_first(), pol_side = l2_policy_evict_last();
  const bool dense = (v_pitch == (long long)ROWB) && (s_pitch == (long long)SB);
  int ipos = kb, pf_idx = 0;
  auto prefetch = [&]() {
    pf_idx = 0;
    if (lane < RPS && ipos + lane < ke) pf_idx = __ldg(col + ipos + lane);
  };
  auto issue = [&](int st) {
    const int n = min(RPS, ke - ipos);
    const uint32_t bar = bars + st * 8;
#if ALLSET_STREAM_COPY == 1
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)n * (ROWB + SB));
    __syncwarp();
#endif
    if (dense) {      // the usual case keeps the compile-time row pitch (a shift instead of a 64-bit multiply per chunk)
      stage_rows<ROWB, true>(ring + (uint32_t)st * (RPS * ROWB), reinterpret_cast<const unsigned char*>(v), pf_idx, n, bar, lane, pol_rows);
      stage_side(sring + (uint32_t)st * RPS * SB, reinterpret_cast<const unsigned char*>(score), pf_idx, n, SB, bar, lane, pol_side, SB);
    } else {
      stage_rows<ROWB, true>(ring + (uint32_t)st * (RPS * ROWB), reinterpret_cast<const unsigned char*>(v), pf_idx, n, bar, lane, pol_rows,
                             (size_t)v_pitch);
      stage_side(sring + (uint32_t)st * RPS * SB, reinterpret_cast<const unsigned char*>(score), pf_idx, n, SB, bar, lane,
                 s_pitch == v_pitch ? pol_rows : pol_side, (size_t)s_pitch);
    }
#if ALLSET_STREAM_COPY == 0
    cp_async_arrive(bar);
#endif
    ipos += RPS;
    prefetch();
  };
  prefetch();
  for (int st = 0; st < stages && ipos < ke; ++st) issue(st);

  int st = 0;
  uint32_t phase = 0;
  for (int cbase = kb; cbase < ke; cbase += RPS) {
    const int n = min(RPS, ke - cbase);
    mbar_wait(bars + st * 8, phase);
    const uint32_t rows = ring + (uint32_t)st * (RPS * ROWB);
    const uint32_t scs = sring + (uint32_t)st * RPS * SB;
    // scores of the stage -> a2 = leaky_relu(score) * log2(e), in place, each (row, head) once (not once per lane)
    for (int i = lane; i < (n * H) / 4; i += 32) {
      float4 sc;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sc.x), "=f"(sc.y), "=f"(sc.z), "=f"(sc.w) : "r"(scs + (uint32_t)i * 16u));
      sc.x = leaky(sc.x, slope) * kLog2e; sc.y = leaky(sc.y, slope) * kLog2e;
      sc.z = leaky(sc.z, slope) * kLog2e; sc.w = leaky(sc.w, slope) * kLog2e;
      asm volatile("st.shared.v4.f32 [%4], {%0,%1,%2,%3};" ::"f"(sc.x), "f"(sc.y), "f"(sc.z), "f"(sc.w), "r"(scs + (uint32_t)i * 16u) : "memory");
    }
    __syncwarp();
    int r = 0;
    while (r < n) {
      while (pos >= cur_end) flush();
      const int plen = min(n - r, cur_end - pos);          // rows of the current segment inside this stage
      const uint32_t prow = rows + (uint32_t)r * ROWB, psc = scs + (uint32_t)r * SB;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (plen <= kPmaOnlinePiece) {
          // short piece (the E->V direction: ~5 rows per piece): straight-line code per piece length -- the scores of the
          // piece stay in registers between the max pass and the weight pass, one rescale of the carried state per piece
          // (none at all for the first piece of a segment), no loop or remainder control.  History: two-pass loops unrolled
          // by 4 / 2 were issue-bound on their control overhead (70 warp instructions per incidence); a single-pass online
          // max (two ex2 and a rescale per ROW) brought that to 61; this form needs ~11 per row.
          switch (plen) {
            case 1: short_piece(std::integral_constant<int, 1>{}, c, prow, psc); break;
            case 2: short_piece(std::integral_constant<int, 2>{}, c, prow, psc); break;
            case 3: short_piece(std::integral_constant<int, 3>{}, c, prow, psc); break;
            case 4: short_piece(std::integral_constant<int, 4>{}, c, prow, psc); break;
            case 5: short_piece(std::integral_constant<int, 5>{}, c, prow, psc); break;
            case 6: short_piece(std::integral_constant<int, 6>{}, c, prow, psc); break;
            case 7: short_piece(std::integral_constant<int, 7>{}, c, prow, psc); break;
            default: short_piece(std::integral_constant<int, 8>{}, c, prow, psc); break;
          }
          continue;
        }
        // pass 1: piece max
        float pm = -INFINITY;
        int k = 0;
        for (; k + 4 <= plen; k += 4) {
          const float a0 = load_a2(psc + (k + 0) * SB, c), a1 = load_a2(psc + (k + 1) * SB, c);
          const float a2v = load_a2(psc + (k + 2) * SB, c), a3 = load_a2(psc + (k + 3) * SB, c);
          pm = fmaxf(fmaxf(pm, fmaxf(a0, a1)), fmaxf(a2v, a3));
        }
        for (; k < plen; ++k) pm = fmaxf(pm, load_a2(psc + k * SB, c));
        // one rescale of the carried state per piece
        const float mn = fmaxf(m2[c], pm);
        const float cf = ex2_approx(m2[c] - mn);            // 2^-inf = 0 for the first piece of a segment
        m2[c] = mn;
        float lc = l[c] * cf;
#pragma unroll
        for (int i = 0; i < EPC; i += 2) {
          if (EPC >= 2) fmul2(acc[c * EPC + i], acc[c * EPC + i + 1], cf);
          else acc[c * EPC + i] *= cf;
        }
        // pass 2: weights and weighted sum
        k = 0;
        for (; k + 2 <= plen; k += 2) {
          uint32_t r0[4], r1[4];
          load_chunk(prow + (k + 0) * ROWB, c, r0);
          load_chunk(prow + (k + 1) * ROWB, c, r1);
          const float p0 = ex2_approx(load_a2(psc + (k + 0) * SB, c) - mn);
          const float p1 = ex2_approx(load_a2(psc + (k + 1) * SB, c) - mn);
          lc += p0;
          fma_chunk(c, r0, p0);
          lc += p1;
          fma_chunk(c, r1, p1);
        }
        if (k < plen) {
          uint32_t r0[4];
          load_chunk(prow + k * ROWB, c, r0);
          const float p0 = ex2_approx(load_a2(psc + k * SB, c) - mn);
          lc += p0;
          fma_chunk(c, r0, p0);
        }
        l[c] = lc;
      }
      r += plen;
      pos += plen;
      fresh = false;
    }
    __syncwarp();
    if (ipos < ke) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(st);
    }
    if (++st == stages) {
      st = 0;
      phase ^= 1u;
    }
  }
  while (seg < nseg) flush();

  // ---- pieces of cut segments (see chunk_cut) -------------------------------------------------------------
  if (c1.inside) {
    park(wsv.tail + chunk * PS);
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_gpu(wsv.flags + chunk, (c0.inside && nseg == 0) ? 2 : 1);
  }
  if (c0.inside && nseg > 0) {
    long long q0 = chunk - 1;
    while (wait_piece_flag(wsv, q0, lane) != 1 && q0 > 0) --q0;
    float tm[CH], tl[CH], ta[NA];
#pragma unroll
    for (int c = 0; c < CH; ++c) { tm[c] = -INFINITY; tl[c] = 0.f; }
#pragma unroll
    for (int i = 0; i < NA; ++i) ta[i] = 0.f;
    // online-softmax merge of the pieces in stream order; the warp's own parked piece comes last
    for (long long q = q0; q <= chunk; ++q) {
      const float* slot = q < chunk ? wsv.tail + q * PS : wsv.head + chunk * PS;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float pm = __ldcg(slot + (NA + c) * 32 + lane), pl = __ldcg(slot + (NA + CH + c) * 32 + lane);
        const float mn = fmaxf(tm[c], pm);
        const float ca = ex2_approx(tm[c] - mn), cb = ex2_approx(pm - mn);    // 2^-inf = 0; an empty piece has pm = -inf
        const float fa = (tm[c] == -INFINITY) ? 0.f : ca, fb = (pm == -INFINITY) ? 0.f : cb;
        tm[c] = mn;
        tl[c] = tl[c] * fa + pl * fb;
#pragma unroll
        for (int i = 0; i < EPC; ++i)
          ta[c * EPC + i] = ta[c * EPC + i] * fa + __ldcg(slot + (c * EPC + i) * 32 + lane) * fb;
      }
    }
    __syncwarp();
    if (lane == 0)
      for (long long q = q0; q < chunk; ++q) wsv.flags[q] = 0;
    emit(out + (size_t)s_first * (size_t)d, stats != nullptr ? stats + (size_t)s_first * H * 2 : nullptr, s_first, tm, tl, ta);
  }
  if (PUSH == 2) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
}

// grad_w[k] = tgt_scale[t] * <x[col[k]], grad_out[t]>; one warp per segment (LearnMask only).
template <typename T>
__global__ void __launch_bounds__(kThreads)
segreduce_bwd_w_kernel(const T* __restrict__ x, const T* __restrict__ go,
                       const int* __restrict__ rowptr, const int* __restrict__ col,
                       const float* __restrict__ tscale, long long n_tgt, int d,
                       float* __restrict__ gw) {
  const int lane = threadIdx.x & 31;
  const long long seg = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (seg >= n_tgt) return;
  const int beg = __ldg(rowptr + seg), end = __ldg(rowptr + seg + 1);
  const float sc = tscale != nullptr ? __ldg(tscale + seg) : 1.f;
  const T* grow = go + (size_t)seg * (size_t)d;
  for (int k = beg; k < end; ++k) {
    const T* xrow = x + (size_t)__ldg(col + k) * (size_t)d;
    float s = 0.f;
    for (int f = lane; f < d; f += 32) s += (float)xrow[f] * (float)grow[f];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) gw[k] = s * sc;
  }
}

// ---------------------------------------------------------------------------------------------
// AllSetTransformer: PMA forward (online segment softmax fused with the weighted row sum)
// ---------------------------------------------------------------------------------------------
template <typename T, bool VECTOR, int G>
__device__ __forceinline__ void pma_accumulate(const T* __restrict__ v, const float* __restrict__ score,
                                               const int* __restrict__ col, int first, int end,
                                               int stride, int d, int H, int h, float slope, int feat,
                                               bool active, int gl, unsigned gmask, float& m, float& l,
                                               float (&acc)[Chunk<T, VECTOR>::N]) {
  using CH = Chunk<T, VECTOR>;
  constexpr int N = CH::N;
  constexpr int U = G < kUnrollPma ? G : kUnrollPma;
  for (int base = first; base < end; base += stride) {
    const int n = min(G, end - base);
    int myidx = 0;
    if (gl < n) myidx = __ldg(col + base + gl);
#pragma unroll 1
    for (int u = 0; u < n; u += U) {
      typename CH::Raw raw[U];
      float a[U];
      float bm = -INFINITY;
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int idx = __shfl_sync(gmask, myidx, u + k, G);
        raw[k] = CH::zero();
        a[k] = -INFINITY;
        if (u + k < n && active) {
          a[k] = leaky(__ldg(score + (size_t)idx * (size_t)H + h), slope);   // padding slots stay -inf for any slope
          raw[k] = CH::load(v + (size_t)idx * (size_t)d + feat);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) bm = fmaxf(bm, a[k]);
      const float m_new = fmaxf(m, bm);
      const float c = (m == -INFINITY) ? 0.f : expf(m - m_new);
      l *= c;
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] *= c;
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const float p = (a[k] == -INFINITY) ? 0.f : expf(a[k] - m_new);
        l += p;
        float f[N];
        CH::unpack(raw[k], f);
#pragma unroll
        for (int i = 0; i < N; ++i) acc[i] = fmaf(p, f[i], acc[i]);
      }
      m = m_new;
    }
  }
}

template <typename T, bool VECTOR, int N>
__device__ __forceinline__ void pma_finish(float m, float l, float (&acc)[N], const float* __restrict__ seed,
                                           int feat, int C, int h, bool active, long long seg, int d,
                                           int H, T* __restrict__ out, float* __restrict__ stats) {
  using CH = Chunk<T, VECTOR>;
  if (!active) return;
  const float denom = l + 1e-16f;
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = acc[i] / denom + __ldg(seed + feat + i);
  CH::store(out + (size_t)seg * (size_t)d + feat, acc);
  if (stats != nullptr && feat == h * C) {
    stats[((size_t)seg * H + h) * 2 + 0] = m;
    stats[((size_t)seg * H + h) * 2 + 1] = denom;
  }
}

template <typename T, bool VECTOR, int G>
__global__ void __launch_bounds__(kThreads)
pma_fwd_group_kernel(const T* __restrict__ v, const float* __restrict__ score,
                     const float* __restrict__ seed, const int* __restrict__ rowptr,
                     const int* __restrict__ col, long long n_tgt, int H, int C, float slope,
                     int skip_over, T* __restrict__ out, float* __restrict__ stats) {
  using CH = Chunk<T, VECTOR>;
  constexpr int N = CH::N;
  const int d = H * C;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const unsigned gmask = group_mask<G>(lane);
  const long long seg = ((long long)blockIdx.x * kThreads + threadIdx.x) / G;
  if (seg >= n_tgt) return;
  const int feat = (blockIdx.y * 32 + gl) * N;
  const bool active = feat < d;
  const int h = active ? feat / C : 0;
  const int beg = __ldg(rowptr + seg), end = __ldg(rowptr + seg + 1);
  if (end - beg > skip_over) return;
  float m = -INFINITY, l = 0.f;
  float acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = 0.f;
  pma_accumulate<T, VECTOR, G>(v, score, col, beg, end, G, d, H, h, slope, feat, active, gl, gmask, m, l, acc);
  pma_finish<T, VECTOR, N>(m, l, acc, seed, feat, C, h, active, seg, d, H, out, stats);
}

template <typename T, bool VECTOR, int G>
__global__ void __launch_bounds__(kThreads)
pma_fwd_cta_kernel(const T* __restrict__ v, const float* __restrict__ score,
                   const float* __restrict__ seed, const int* __restrict__ rowptr,
                   const int* __restrict__ col, const int* __restrict__ long_ids, int H, int C,
                   float slope, T* __restrict__ out, float* __restrict__ stats) {
  using CH = Chunk<T, VECTOR>;
  constexpr int N = CH::N;
  constexpr int NG = kThreads / G;
  __shared__ float red[kThreads * N];
  __shared__ float red_m[kThreads];
  __shared__ float red_l[kThreads];
  const int d = H * C;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const int grp = threadIdx.x / G;
  const unsigned gmask = group_mask<G>(lane);
  const long long seg = long_ids[blockIdx.x];
  const int feat = (blockIdx.y * 32 + gl) * N;
  const bool active = feat < d;
  const int h = active ? feat / C : 0;
  const int beg = __ldg(rowptr + seg), end = __ldg(rowptr + seg + 1);
  float m = -INFINITY, l = 0.f;
  float acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = 0.f;
  pma_accumulate<T, VECTOR, G>(v, score, col, beg + grp * G, end, NG * G, d, H, h, slope, feat, active, gl,
                               gmask, m, l, acc);
  red_m[threadIdx.x] = m;
  red_l[threadIdx.x] = l;
#pragma unroll
  for (int i = 0; i < N; ++i) red[threadIdx.x * N + i] = acc[i];
  __syncthreads();
  if (grp == 0) {
    float mt = m;
    for (int g = 1; g < NG; ++g) mt = fmaxf(mt, red_m[g * G + gl]);
    float lt = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = 0.f;
    for (int g = 0; g < NG; ++g) {
      const float mg = red_m[g * G + gl];
      const float c = (mg == -INFINITY) ? 0.f : expf(mg - mt);
      lt += red_l[g * G + gl] * c;
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = fmaf(red[(g * G + gl) * N + i], c, acc[i]);
    }
    pma_finish<T, VECTOR, N>(mt, lt, acc, seed, feat, C, h, active, seg, d, H, out, stats);
  }
}

// alpha[k,h] in CSR order; one warp per segment (only when attention weights are requested)
__global__ void __launch_bounds__(kThreads)
pma_alpha_kernel(const float* __restrict__ score, const float* __restrict__ stats,
                 const int* __restrict__ rowptr, const int* __restrict__ col, long long n_tgt, int H,
                 float slope, float* __restrict__ alpha) {
  const int lane = threadIdx.x & 31;
  const long long seg = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (seg >= n_tgt) return;
  const int beg = __ldg(rowptr + seg), end = __ldg(rowptr + seg + 1);
  const long long total = (long long)(end - beg) * H;
  for (long long i = lane; i < total; i += 32) {
    const int k = beg + (int)(i / H), h = (int)(i % H);
    const float a = leaky(__ldg(score + (size_t)__ldg(col + k) * H + h), slope);
    const float m = stats[((size_t)seg * H + h) * 2], den = stats[((size_t)seg * H + h) * 2 + 1];
    alpha[(size_t)k * H + h] = expf(a - m) / den;
  }
}

// ---------------------------------------------------------------------------------------------
// per-row per-head dot: out[r,h] = sum_c a[r,h,c] * (b[r,h,c] - sub[h,c])
// ---------------------------------------------------------------------------------------------
template <typename T, bool VECTOR, int G>
__global__ void __launch_bounds__(kThreads)
rowdot_group_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ sub,
                    long long n_rows, int H, int C, float* __restrict__ out) {
  using CH = Chunk<T, VECTOR>;
  constexpr int N = CH::N;
  const int d = H * C;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const unsigned gmask = group_mask<G>(lane);
  const long long row = ((long long)blockIdx.x * kThreads + threadIdx.x) / G;
  if (row >= n_rows) return;
  const int feat = gl * N;
  const bool active = feat < d;
  float s = 0.f;
  if (active) {
    float fa[N], fb[N];
    CH::unpack(CH::load(a + (size_t)row * d + feat), fa);
    CH::unpack(CH::load(b + (size_t)row * d + feat), fb);
#pragma unroll
    for (int i = 0; i < N; ++i) s = fmaf(fa[i], sub != nullptr ? fb[i] - __ldg(sub + feat + i) : fb[i], s);
  }
  const int lph = C / N;
  s = head_reduce<G>(s, gmask, lph, gl);
  if (active && feat % C == 0) out[(size_t)row * H + feat / C] = s;
}

// generic fallback: one thread per (row, head)
template <typename T>
__global__ void __launch_bounds__(kThreads)
rowdot_thread_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ sub,
                     long long n_rows, int H, int C, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_rows * H) return;
  const int h = (int)(i % H);
  const size_t off = (size_t)(i / H) * H * C + (size_t)h * C;
  float s = 0.f;
  for (int c = 0; c < C; ++c) {
    const float bv = (float)b[off + c] - (sub != nullptr ? sub[h * C + c] : 0.f);
    s = fmaf((float)a[off + c], bv, s);
  }
  out[i] = s;
}

// ---------------------------------------------------------------------------------------------
// PMA backward over the transposed CSR (segments = source rows)
// ---------------------------------------------------------------------------------------------
template <typename T, bool VECTOR, int G>
__device__ __forceinline__ void pma_bwd_accumulate(const T* __restrict__ go, const float* __restrict__ stats,
                                                   const float* __restrict__ D, const int* __restrict__ col,
                                                   int first, int end, int stride, int d, int H, int h,
                                                   float a_s, int feat, bool active, int gl, unsigned gmask,
                                                   float& sD, float (&gv)[Chunk<T, VECTOR>::N]) {
  using CH = Chunk<T, VECTOR>;
  constexpr int N = CH::N;
  constexpr int U = G < kUnrollPma ? G : kUnrollPma;
  for (int base = first; base < end; base += stride) {
    const int n = min(G, end - base);
    int myidx = 0;
    if (gl < n) myidx = __ldg(col + base + gl);
#pragma unroll 1
    for (int u = 0; u < n; u += U) {
      typename CH::Raw raw[U];
      float2 st[U];
      float dd[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int idx = __shfl_sync(gmask, myidx, u + k, G);
        raw[k] = CH::zero();
        st[k] = make_float2(INFINITY, 1.f);   // alpha = exp(-inf) = 0 for empty slots
        dd[k] = 0.f;
        if (u + k < n && active) {
          st[k] = __ldg(reinterpret_cast<const float2*>(stats) + (size_t)idx * H + h);
          dd[k] = __ldg(D + (size_t)idx * H + h);
          raw[k] = CH::load(go + (size_t)idx * (size_t)d + feat);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const float alpha = expf(a_s - st[k].x) / st[k].y;
        sD = fmaf(alpha, dd[k], sD);
        float f[N];
        CH::unpack(raw[k], f);
#pragma unroll
        for (int i = 0; i < N; ++i) gv[i] = fmaf(alpha, f[i], gv[i]);
      }
    }
  }
}

// FUSED: the whole row fits one slab, so <grad_v, v> is reduced in-register and grad_score is final.
// Otherwise grad_score receives S = sum_k alpha_k D[t_k,h] and pma_score_finish_kernel completes it.
template <typename T, bool VECTOR, int G, bool FUSED>
__global__ void __launch_bounds__(kThreads)
pma_bwd_group_kernel(const T* __restrict__ go, const T* __restrict__ v, const float* __restrict__ score,
                     const float* __restrict__ stats, const float* __restrict__ D,
                     const int* __restrict__ rowptr, const int* __restrict__ col, long long n_src, int H,
                     int C, float slope, int skip_over, T* __restrict__ gvout, float* __restrict__ gscore) {
  using CH = Chunk<T, VECTOR>;
  constexpr int N = CH::N;
  const int d = H * C;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const unsigned gmask = group_mask<G>(lane);
  const long long seg = ((long long)blockIdx.x * kThreads + threadIdx.x) / G;
  if (seg >= n_src) return;
  const int feat = (blockIdx.y * 32 + gl) * N;
  const bool active = feat < d;
  const int h = active ? feat / C : 0;
  const int beg = __ldg(rowptr + seg), end = __ldg(rowptr + seg + 1);
  if (end - beg > skip_over) return;
  const float s_raw = active ? __ldg(score + (size_t)seg * H + h) : 0.f;
  const float a_s = leaky(s_raw, slope);
  float sD = 0.f;
  float gv[N];
#pragma unroll
  for (int i = 0; i < N; ++i) gv[i] = 0.f;
  pma_bwd_accumulate<T, VECTOR, G>(go, stats, D, col, beg, end, G, d, H, h, a_s, feat, active, gl, gmask, sD, gv);
  if (active) CH::store(gvout + (size_t)seg * (size_t)d + feat, gv);
  if (FUSED) {
    float dot = 0.f;
    if (active) {
      float fv[N];
      CH::unpack(CH::load(v + (size_t)seg * (size_t)d + feat), fv);
#pragma unroll
      for (int i = 0; i < N; ++i) dot = fmaf(gv[i], fv[i], dot);
    }
    dot = head_reduce<G>(dot, gmask, C / N, gl);
    if (active && feat % C == 0) gscore[(size_t)seg * H + h] = (s_raw > 0.f ? 1.f : slope) * (dot - sD);
  } else {
    if (active && feat % C == 0) gscore[(size_t)seg * H + h] = sD;
  }
}

template <typename T, bool VECTOR, int G, bool FUSED>
__global__ void __launch_bounds__(kThreads)
pma_bwd_cta_kernel(const T* __restrict__ go, const T* __restrict__ v, const float* __restrict__ score,
                   const float* __restrict__ stats, const float* __restrict__ D,
                   const int* __restrict__ rowptr, const int* __restrict__ col,
                   const int* __restrict__ long_ids, int H, int C, float slope, T* __restrict__ gvout,
                   float* __restrict__ gscore) {
  using CH = Chunk<T, VECTOR>;
  constexpr int N = CH::N;
  constexpr int NG = kThreads / G;
  __shared__ float red[kThreads * N];
  __shared__ float red_s[kThreads];
  const int d = H * C;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const int grp = threadIdx.x / G;
  const unsigned gmask = group_mask<G>(lane);
  const long long seg = long_ids[blockIdx.x];
  const int feat = (blockIdx.y * 32 + gl) * N;
  const bool active = feat < d;
  const int h = active ? feat / C : 0;
  const int beg = __ldg(rowptr + seg), end = __ldg(rowptr + seg + 1);
  const float s_raw = active ? __ldg(score + (size_t)seg * H + h) : 0.f;
  const float a_s = leaky(s_raw, slope);
  float sD = 0.f;
  float gv[N];
#pragma unroll
  for (int i = 0; i < N; ++i) gv[i] = 0.f;
  pma_bwd_accumulate<T, VECTOR, G>(go, stats, D, col, beg + grp * G, end, NG * G, d, H, h, a_s, feat, active,
                                   gl, gmask, sD, gv);
  red_s[threadIdx.x] = sD;
#pragma unroll
  for (int i = 0; i < N; ++i) red[threadIdx.x * N + i] = gv[i];
  __syncthreads();
  if (grp == 0) {   // G >= 32 or a whole number of groups per warp: warp 0 (or part of it) is uniform here
    for (int g = 1; g < NG; ++g) {
      sD += red_s[g * G + gl];
#pragma unroll
      for (int i = 0; i < N; ++i) gv[i] += red[(g * G + gl) * N + i];
    }
    if (active) CH::store(gvout + (size_t)seg * (size_t)d + feat, gv);
    if (FUSED) {
      float dot = 0.f;
      if (active) {
        float fv[N];
        CH::unpack(CH::load(v + (size_t)seg * (size_t)d + feat), fv);
#pragma unroll
        for (int i = 0; i < N; ++i) dot = fmaf(gv[i], fv[i], dot);
      }
      dot = head_reduce<G>(dot, gmask, C / N, gl);
      if (active && feat % C == 0) gscore[(size_t)seg * H + h] = (s_raw > 0.f ? 1.f : slope) * (dot - sD);
    } else {
      if (active && feat % C == 0) gscore[(size_t)seg * H + h] = sD;
    }
  }
}

// multi-slab rows: grad_score[s,h] = leaky'(score) * (<grad_v[s,h,:], v[s,h,:]> - S[s,h])
template <typename T>
__global__ void __launch_bounds__(kThreads)
pma_score_finish_kernel(const T* __restrict__ gv, const T* __restrict__ v, const float* __restrict__ score,
                        long long n_src, int H, int C, float slope, float* __restrict__ gscore) {
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_src * H) return;
  const size_t off = (size_t)(i / H) * H * C + (size_t)(i % H) * C;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s = fmaf((float)gv[off + c], (float)v[off + c], s);
  gscore[i] = (score[i] > 0.f ? 1.f : slope) * (s - gscore[i]);
}

// ---------------------------------------------------------------------------------------------
// Dense glue of MLP / PMA (reference layers.py:571-579, 153-157): out = LayerNorm(residual + relu(x + bias)), every
// stage optional.  One warp per row, the row lives in registers (d = 128 * NV: NV float4 per lane), two-pass mean /
// variance with warp shuffles -> one HBM read and one write of the row, where ATen spends a bias kernel, a ReLU
// kernel and a LayerNorm kernel (measured 1.67 ms per [1M,128] LayerNorm alone vs 0.16 ms at HBM speed).
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
bias_act_norm_kernel(const float* __restrict__ x, const float* __restrict__ bias, int relu,
                     const float* __restrict__ residual, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, long long rows, float* __restrict__ out,
                     float* __restrict__ stats) {
  constexpr int D = 128 * NV;
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __ldcs(xr + i * 32 + lane);
  if (bias != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + i * 32 + lane);
      v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
    }
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x = fmaxf(v[i].x, 0.f); v[i].y = fmaxf(v[i].y, 0.f); v[i].z = fmaxf(v[i].z, 0.f); v[i].w = fmaxf(v[i].w, 0.f);
    }
  }
  if (residual != nullptr) {
    const float4* rr = reinterpret_cast<const float4*>(residual + row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 r = __ldcs(rr + i * 32 + lane);
      v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
    }
  }
  if (gamma != nullptr) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + e * e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.f / D) + eps);
    if (stats != nullptr && lane == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(mean, rstd);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (beta != nullptr) b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
      v[i].x = (v[i].x - mean) * rstd * g.x + b.x; v[i].y = (v[i].y - mean) * rstd * g.y + b.y;
      v[i].z = (v[i].z - mean) * rstd * g.z + b.z; v[i].w = (v[i].w - mean) * rstd * g.w + b.w;
    }
  }
  float4* orow = reinterpret_cast<float4*>(out + row * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) orow[i * 32 + lane] = v[i];
}

// Backward of bias_act_norm for d = 128 * NV.  A warp walks rows with a grid stride, so every lane owns fixed columns:
// d(gamma), d(beta), d(bias) accumulate in registers over all rows of the warp, are combined per CTA in shared memory
// and written as ONE partial row per CTA (`partial[cta][3][D]`, summed by the host: deterministic, no atomics).
//   z = residual + act(x + bias),  zh = (z - mean) * rstd,  y = zh * gamma + beta
//   g = dy * gamma;  dz = rstd * (g - mean_d(g) - zh * mean_d(g * zh))   (dz = dy without LayerNorm)
//   d(residual) = dz;  d(x) = dz * [x + bias > 0]  (relu)  else dz
template <int NV>
__global__ void __launch_bounds__(256)
bias_act_norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ bias,
                         int relu, const float* __restrict__ residual, const float* __restrict__ gamma,
                         const float* __restrict__ stats, long long rows, float* __restrict__ dx,
                         float* __restrict__ dres, float* __restrict__ partial) {
  constexpr int D = 128 * NV;
  __shared__ float red[8][3][32 * 4];                 // per warp, per quantity, one float4 slot per lane (re-used per i)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * 8;
  float4 bsv[NV], gmv[NV];
  float4 dgam[NV], dbet[NV], dbia[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    bsv[i] = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias) + i * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    gmv[i] = gamma != nullptr ? __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane) : make_float4(1.f, 1.f, 1.f, 1.f);
    dgam[i] = dbet[i] = dbia[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += nwarps) {
    float4 pre[NV], g[NV], zh[NV];
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    const float4* dyr = reinterpret_cast<const float4*>(dy + row * D);
    float mean = 0.f, rstd = 1.f;
    if (gamma != nullptr) {
      const float2 st = __ldg(reinterpret_cast<const float2*>(stats + row * 2));
      mean = st.x; rstd = st.y;
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 p = __ldcs(xr + i * 32 + lane);
      p.x += bsv[i].x; p.y += bsv[i].y; p.z += bsv[i].z; p.w += bsv[i].w;
      pre[i] = p;
      float4 z = p;
      if (relu) { z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f); z.z = fmaxf(z.z, 0.f); z.w = fmaxf(z.w, 0.f); }
      if (residual != nullptr) {
        const float4 r = __ldcs(reinterpret_cast<const float4*>(residual + row * D) + i * 32 + lane);
        z.x += r.x; z.y += r.y; z.z += r.z; z.w += r.w;
      }
      const float4 d_ = __ldcs(dyr + i * 32 + lane);
      float4 h = make_float4((z.x - mean) * rstd, (z.y - mean) * rstd, (z.z - mean) * rstd, (z.w - mean) * rstd);
      zh[i] = h;
      float4 gg = make_float4(d_.x * gmv[i].x, d_.y * gmv[i].y, d_.z * gmv[i].z, d_.w * gmv[i].w);
      g[i] = gg;
      if (gamma != nullptr) {
        dgam[i].x += d_.x * h.x; dgam[i].y += d_.y * h.y; dgam[i].z += d_.z * h.z; dgam[i].w += d_.w * h.w;
        dbet[i].x += d_.x; dbet[i].y += d_.y; dbet[i].z += d_.z; dbet[i].w += d_.w;
        s1 += (gg.x + gg.y) + (gg.z + gg.w);
        s2 += (gg.x * h.x + gg.y * h.y) + (gg.z * h.z + gg.w * h.w);
      }
    }
    if (gamma != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      s1 *= (1.f / D);
      s2 *= (1.f / D);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 dz = g[i];
      if (gamma != nullptr) {
        dz.x = rstd * (g[i].x - s1 - zh[i].x * s2); dz.y = rstd * (g[i].y - s1 - zh[i].y * s2);
        dz.z = rstd * (g[i].z - s1 - zh[i].z * s2); dz.w = rstd * (g[i].w - s1 - zh[i].w * s2);
      }
      if (dres != nullptr) reinterpret_cast<float4*>(dres + row * D)[i * 32 + lane] = dz;
      float4 dp = dz;
      if (relu) {
        dp.x = pre[i].x > 0.f ? dz.x : 0.f; dp.y = pre[i].y > 0.f ? dz.y : 0.f;
        dp.z = pre[i].z > 0.f ? dz.z : 0.f; dp.w = pre[i].w > 0.f ? dz.w : 0.f;
      }
      dbia[i].x += dp.x; dbia[i].y += dp.y; dbia[i].z += dp.z; dbia[i].w += dp.w;
      reinterpret_cast<float4*>(dx + row * D)[i * 32 + lane] = dp;
    }
  }
  // CTA-level combine of the column sums, one NV slice at a time through shared memory
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    reinterpret_cast<float4*>(red[warp][0])[lane] = dgam[i];
    reinterpret_cast<float4*>(red[warp][1])[lane] = dbet[i];
    reinterpret_cast<float4*>(red[warp][2])[lane] = dbia[i];
    __syncthreads();
    if (warp < 3) {                                 // warp q sums quantity q over the 8 warps
      float4 a = reinterpret_cast<float4*>(red[0][warp])[lane];
#pragma unroll
      for (int w2 = 1; w2 < 8; ++w2) {
        const float4 b = reinterpret_cast<float4*>(red[w2][warp])[lane];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      reinterpret_cast<float4*>(partial + ((size_t)blockIdx.x * 3 + warp) * D)[i * 32 + lane] = a;
    }
    __syncthreads();
  }
}

// any width: one warp per row, three passes over the (L1/L2-resident) row
__global__ void __launch_bounds__(256)
bias_act_norm_generic_kernel(const float* __restrict__ x, const float* __restrict__ bias, int relu,
                             const float* __restrict__ residual, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float eps, long long rows, int d,
                             float* __restrict__ out, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* xr = x + row * d;
  const float* rr = residual != nullptr ? residual + row * d : nullptr;
  auto value = [&](int j) {
    float t = xr[j];
    if (bias != nullptr) t += __ldg(bias + j);
    if (relu) t = fmaxf(t, 0.f);
    if (rr != nullptr) t += rr[j];
    return t;
  };
  float mean = 0.f, rstd = 1.f;
  if (gamma != nullptr) {
    float sum = 0.f;
    for (int j = lane; j < d; j += 32) sum += value(j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mean = sum / (float)d;
    float sq = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float t = value(j) - mean;
      sq += t * t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    rstd = rsqrtf(sq / (float)d + eps);
    if (stats != nullptr && lane == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(mean, rstd);
  }
  for (int j = lane; j < d; j += 32) {
    float t = value(j);
    if (gamma != nullptr) t = (t - mean) * rstd * __ldg(gamma + j) + (beta != nullptr ? __ldg(beta + j) : 0.f);
    out[row * d + j] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// CSR construction
// ---------------------------------------------------------------------------------------------
__global__ void csr_keys_kernel(const long long* __restrict__ tgt, long long nnz, int* __restrict__ keys,
                                int* __restrict__ pos) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) {
    keys[i] = (int)tgt[i];
    pos[i] = (int)i;
  }
}

// rowptr[k] = first sorted slot whose key >= k; col[i] = src[perm[i]]
__global__ void csr_fill_kernel(const int* __restrict__ keys_sorted, const int* __restrict__ perm,
                                const long long* __restrict__ src, long long nnz, long long n_tgt,
                                int* __restrict__ rowptr, int* __restrict__ col) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nnz) return;
  if (i < nnz) col[i] = (int)src[perm[i]];
  const long long lo = (i == 0) ? 0 : (long long)keys_sorted[i - 1] + 1;
  const long long hi = (i == nnz) ? n_tgt : (long long)keys_sorted[i];
  for (long long k = lo; k <= hi && k <= n_tgt; ++k) rowptr[k] = (int)i;
}

struct LongSegment {
  const int* rowptr;
  int threshold;
  __host__ __device__ bool operator()(int r) const { return rowptr[r + 1] - rowptr[r] > threshold; }
};

int bits_for(long long n) {
  int b = 1;
  while (b < 31 && (1LL << b) < n) ++b;
  return b;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
struct Shape {
  bool vector;   // 16-byte chunk path
  int G;         // lanes per group
  int slabs;     // gridDim.y
};

Shape plan(int d, int elem_bytes, const void* p0, const void* p1, const void* p2 = nullptr) {
  Shape s;
  const int per = 16 / elem_bytes;
  const uintptr_t bits = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2;
  s.vector = (d % per == 0) && (bits % 16 == 0);
  const int nch = s.vector ? d / per : d;
  int g = 1;
  while (g < 32 && g < nch) g <<= 1;
  s.G = g;
  s.slabs = (nch + 31) / 32;
  return s;
}

unsigned blocks_for(long long units, int G) {
  const long long per = kThreads / G;
  return (unsigned)((units + per - 1) / per);
}

#define ALLSET_DISPATCH_G(G_, ...)                 \
  switch (G_) {                                    \
    case 1: { constexpr int G = 1; __VA_ARGS__; break; }   \
    case 2: { constexpr int G = 2; __VA_ARGS__; break; }   \
    case 4: { constexpr int G = 4; __VA_ARGS__; break; }   \
    case 8: { constexpr int G = 8; __VA_ARGS__; break; }   \
    case 16: { constexpr int G = 16; __VA_ARGS__; break; } \
    default: { constexpr int G = 32; __VA_ARGS__; break; } \
  }

template <typename T, bool VECTOR, bool WEIGHTED>
void launch_segreduce(const Shape& sh, const T* x, const int* rowptr, const int* col, const float* w,
                      const float* sscale, long long n_tgt, int d, int mean, const int* long_ids,
                      int n_long, int long_threshold, T* out, cudaStream_t st) {
  const int skip = n_long > 0 ? long_threshold : INT32_MAX;
  ALLSET_DISPATCH_G(sh.G, {
    dim3 grid(blocks_for(n_tgt, G), sh.slabs);
    segreduce_group_kernel<T, VECTOR, G, WEIGHTED><<<grid, kThreads, 0, st>>>(x, rowptr, col, w, sscale, n_tgt,
                                                                           d, mean, skip, out);
    if (n_long > 0) {
      dim3 lgrid(n_long, sh.slabs);
      segreduce_cta_kernel<T, VECTOR, G, WEIGHTED><<<lgrid, kThreads, 0, st>>>(x, rowptr, col, w, sscale,
                                                                            long_ids, d, mean, out);
    }
  });
}

template <typename T>
void segreduce_typed(const Shape& sh, const void* x, const int* rowptr, const int* col, const float* w,
                     const float* sscale, long long n_tgt, int d, int mean, const int* long_ids, int n_long,
                     int long_threshold, void* out, cudaStream_t st) {
  const bool weighted = (w != nullptr) || (sscale != nullptr);
  const T* xi = static_cast<const T*>(x);
  T* oi = static_cast<T*>(out);
  if (sh.vector) {
    if (weighted) launch_segreduce<T, true, true>(sh, xi, rowptr, col, w, sscale, n_tgt, d, mean, long_ids, n_long, long_threshold, oi, st);
    else launch_segreduce<T, true, false>(sh, xi, rowptr, col, w, sscale, n_tgt, d, mean, long_ids, n_long, long_threshold, oi, st);
  } else {
    if (weighted) launch_segreduce<T, false, true>(sh, xi, rowptr, col, w, sscale, n_tgt, d, mean, long_ids, n_long, long_threshold, oi, st);
    else launch_segreduce<T, false, false>(sh, xi, rowptr, col, w, sscale, n_tgt, d, mean, long_ids, n_long, long_threshold, oi, st);
  }
}

template <typename T, bool VECTOR>
void launch_pma_fwd(const Shape& sh, const T* v, const float* score, const float* seed, const int* rowptr,
                    const int* col, long long n_tgt, int H, int C, float slope, const int* long_ids, int n_long,
                    int long_threshold, T* out, float* stats, cudaStream_t st) {
  const int skip = n_long > 0 ? long_threshold : INT32_MAX;
  ALLSET_DISPATCH_G(sh.G, {
    dim3 grid(blocks_for(n_tgt, G), sh.slabs);
    pma_fwd_group_kernel<T, VECTOR, G><<<grid, kThreads, 0, st>>>(v, score, seed, rowptr, col, n_tgt, H, C, slope,
                                                                 skip, out, stats);
    if (n_long > 0) {
      dim3 lgrid(n_long, sh.slabs);
      pma_fwd_cta_kernel<T, VECTOR, G><<<lgrid, kThreads, 0, st>>>(v, score, seed, rowptr, col, long_ids, H, C,
                                                                  slope, out, stats);
    }
  });
}

template <typename T, bool VECTOR, bool FUSED>
void launch_pma_bwd(const Shape& sh, const T* go, const T* v, const float* score, const float* stats,
                    const float* D, const int* rowptr, const int* col, long long n_src, int H, int C, float slope,
                    const int* long_ids, int n_long, int long_threshold, T* gv, float* gscore, cudaStream_t st) {
  const int skip = n_long > 0 ? long_threshold : INT32_MAX;
  ALLSET_DISPATCH_G(sh.G, {
    dim3 grid(blocks_for(n_src, G), sh.slabs);
    pma_bwd_group_kernel<T, VECTOR, G, FUSED><<<grid, kThreads, 0, st>>>(go, v, score, stats, D, rowptr, col, n_src,
                                                                        H, C, slope, skip, gv, gscore);
    if (n_long > 0) {
      dim3 lgrid(n_long, sh.slabs);
      pma_bwd_cta_kernel<T, VECTOR, G, FUSED><<<lgrid, kThreads, 0, st>>>(go, v, score, stats, D, rowptr, col,
                                                                         long_ids, H, C, slope, gv, gscore);
    }
  });
  if (!FUSED) {
    const long long total = n_src * H;
    pma_score_finish_kernel<T><<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, st>>>(
        gv, v, score, n_src, H, C, slope, gscore);
  }
}

template <typename T>
void pma_bwd_typed(const Shape& sh, bool fused, const void* go, const void* v, const float* score,
                   const float* stats, const float* D, const int* rowptr, const int* col, long long n_src, int H,
                   int C, float slope, const int* long_ids, int n_long, int long_threshold, void* gv,
                   float* gscore, cudaStream_t st) {
  const T* g = static_cast<const T*>(go);
  const T* vv = static_cast<const T*>(v);
  T* o = static_cast<T*>(gv);
  if (sh.vector) {
    if (fused) launch_pma_bwd<T, true, true>(sh, g, vv, score, stats, D, rowptr, col, n_src, H, C, slope, long_ids, n_long, long_threshold, o, gscore, st);
    else launch_pma_bwd<T, true, false>(sh, g, vv, score, stats, D, rowptr, col, n_src, H, C, slope, long_ids, n_long, long_threshold, o, gscore, st);
  } else {
    if (fused) launch_pma_bwd<T, false, true>(sh, g, vv, score, stats, D, rowptr, col, n_src, H, C, slope, long_ids, n_long, long_threshold, o, gscore, st);
    else launch_pma_bwd<T, false, false>(sh, g, vv, score, stats, D, rowptr, col, n_src, H, C, slope, long_ids, n_long, long_threshold, o, gscore, st);
  }
}


// ---- stream-kernel dispatch ----------------------------------------------------------------------------------
int stream_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ALLSET_STREAM");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v;
}

struct StreamPlan {
  bool ok;
  int lane_bytes, seg_per_warp, stages, rows_per_stage;
  unsigned blocks;
  size_t smem;
};

// Usable when a row is 128/256/512/1024 bytes (32 lanes x 4/8/16/32 B), rows are 16-byte aligned (bulk copies) and
// there are enough segments that every warp gets a stream of several stages.
StreamPlan plan_stream(int d, int elem_bytes, long long n_tgt, const void* x, const void* out, bool weighted) {
  StreamPlan p{};
  p.ok = false;
  if (!stream_enabled()) return p;
  const long long row_bytes = (long long)d * elem_bytes;
  if (row_bytes != 128 && row_bytes != 256 && row_bytes != 512 && row_bytes != 1024) return p;
  if (((uintptr_t)x % 16) != 0 || ((uintptr_t)out % 16) != 0) return p;
  // >= 16 segments per warp (a stream of several stages), ~8 waves of warps (one CTA per SM) when the graph is large
  // enough, never less than one full wave -- below that the group kernel has more parallelism
  const long long chunks = 148LL * kStreamWarps * 8;
  long long spw = (n_tgt + chunks - 1) / chunks;
  if (spw < 16) spw = 16;
  if (n_tgt < 148LL * kStreamWarps * spw) return p;
  p.lane_bytes = (int)(row_bytes / 32);
  p.seg_per_warp = (int)spw;
  p.rows_per_stage = (int)(kStreamStageBytes / row_bytes < 32 ? kStreamStageBytes / row_bytes : 32);
  const size_t stage_bytes = (size_t)p.rows_per_stage * row_bytes;
  int stages = (int)(ALLSET_STREAM_SMEM / (kStreamWarps * stage_bytes));
  if (stages > 8) stages = 8;
  if (stages < 2) return p;
  p.stages = stages;
  p.smem = (size_t)kStreamWarps * stages * stage_bytes + (size_t)kStreamWarps * stages * 8 +
           (weighted ? (size_t)kStreamWarps * stages * p.rows_per_stage * 4 : 0) + 16;
  const long long warps = (n_tgt + spw - 1) / spw;
  p.blocks = (unsigned)((warps + kStreamWarps - 1) / kStreamWarps);
  p.ok = true;
  return p;
}

// how the fused exchange leaves the SM: 1 = stores by the reducing warp, 2 = TMA bulk stores from a staging slot
int push_mode() {
  // read per launch (a host-side lookup): A/B runs and tests flip it.  Measured on 2 x B200 (profiles/r02_scaling.md): the
  // stores by the reducing warp are faster while the kernel is compute-bound (V->E 1.35 vs 1.44 ms), so they are the default;
  // ALLSET_PUSH=bulk selects the TMA path.
  const char* e = getenv("ALLSET_PUSH");
  return (e != nullptr && (e[0] == 'b' || e[0] == '2')) ? 2 : 1;
}

struct StreamExtra {            // optional arguments of the stream kernels
  const unsigned char* peer_mask;
  void* ws;
};

template <typename T, int LB, bool WEIGHTED>
int launch_stream(const StreamPlan& p, const T* x, const int* rowptr, const int* col, const float* w,
                  const float* sscale, long long n_tgt, int d, int mean, T* out, const PeerOuts& peers,
                  const StreamExtra& ex, cudaStream_t st) {
  // the shared-memory opt-in is per DEVICE and this process may drive several: set it on every launch
  if (peers.n > 0) {
    if (WEIGHTED) return fail(ALLSET_EUNSUPPORTED, "segreduce_fwd_bcast: per-incidence weights are not supported");
    const size_t smem_bulk = p.smem + (size_t)kStreamWarps * 2 * LB * 32;
    if (push_mode() == 2 && smem_bulk <= 227 * 1024) {
      auto kern = segreduce_stream_kernel<T, LB, false, 2>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bulk);
      if (e != cudaSuccess) return fail(ALLSET_ECUDA, "segreduce_stream: smem opt-in: %s", cudaGetErrorString(e));
      kern<<<p.blocks, kStreamWarps * 32, smem_bulk, st>>>(x, rowptr, col, w, sscale, n_tgt, d, mean, p.seg_per_warp,
                                                           p.stages, out, peers, ex.peer_mask, ex.ws);
      return ALLSET_OK;
    }
    auto kern = segreduce_stream_kernel<T, LB, false, 1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return fail(ALLSET_ECUDA, "segreduce_stream: smem opt-in: %s", cudaGetErrorString(e));
    kern<<<p.blocks, kStreamWarps * 32, p.smem, st>>>(x, rowptr, col, w, sscale, n_tgt, d, mean, p.seg_per_warp,
                                                      p.stages, out, peers, ex.peer_mask, ex.ws);
    return ALLSET_OK;
  }
  auto kern = segreduce_stream_kernel<T, LB, WEIGHTED, 0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) return fail(ALLSET_ECUDA, "segreduce_stream: smem opt-in: %s", cudaGetErrorString(e));
  kern<<<p.blocks, kStreamWarps * 32, p.smem, st>>>(x, rowptr, col, w, sscale, n_tgt, d, mean, p.seg_per_warp,
                                                    p.stages, out, peers, nullptr, ex.ws);
  return ALLSET_OK;
}

template <typename T, bool WEIGHTED>
int stream_by_width(const StreamPlan& p, const T* x, const int* rowptr, const int* col, const float* w,
                    const float* sscale, long long n_tgt, int d, int mean, T* out, const PeerOuts& peers,
                    const StreamExtra& ex, cudaStream_t st) {
  switch (p.lane_bytes) {
    case 4: return launch_stream<T, 4, WEIGHTED>(p, x, rowptr, col, w, sscale, n_tgt, d, mean, out, peers, ex, st);
    case 8: return launch_stream<T, 8, WEIGHTED>(p, x, rowptr, col, w, sscale, n_tgt, d, mean, out, peers, ex, st);
    case 16: return launch_stream<T, 16, WEIGHTED>(p, x, rowptr, col, w, sscale, n_tgt, d, mean, out, peers, ex, st);
    default: return launch_stream<T, 32, WEIGHTED>(p, x, rowptr, col, w, sscale, n_tgt, d, mean, out, peers, ex, st);
  }
}

template <typename T>
int segreduce_stream_typed(const StreamPlan& p, const void* x, const int* rowptr, const int* col, const float* w,
                           const float* sscale, long long n_tgt, int d, int mean, void* out, const PeerOuts& peers,
                           const StreamExtra& ex, cudaStream_t st) {
  const bool weighted = (w != nullptr) || (sscale != nullptr);
  if (weighted)
    return stream_by_width<T, true>(p, static_cast<const T*>(x), rowptr, col, w, sscale, n_tgt, d, mean,
                                    static_cast<T*>(out), peers, ex, st);
  return stream_by_width<T, false>(p, static_cast<const T*>(x), rowptr, col, w, sscale, n_tgt, d, mean,
                                   static_cast<T*>(out), peers, ex, st);
}

// peer_outs[j] = address, in peer j's replica, of the row that `out` row 0 is -> byte deltas relative to `out`
int make_peers(void* out, void* const* peer_outs, int n_peers, PeerOuts* po) {
  po->n = 0;
  if (n_peers < 0 || n_peers > 7) return fail(ALLSET_EINVAL, "at most 7 peer replicas (got %d)", n_peers);
  if (n_peers > 0 && peer_outs == nullptr) return fail(ALLSET_EINVAL, "peer_outs is null");
  for (int j = 0; j < n_peers; ++j) {
    if (peer_outs[j] == nullptr || ((uintptr_t)peer_outs[j] % 16) != 0)
      return fail(ALLSET_EINVAL, "peer_outs[%d] is null or not 16-byte aligned", j);
    po->delta[j] = (long long)((intptr_t)peer_outs[j] - (intptr_t)out);
  }
  po->n = n_peers;
  return ALLSET_OK;
}

template <typename T, int LB>
int launch_pma_stream(const StreamPlan& p, const T* v, const float* score, const float* seed, const int* rowptr,
                      const int* col, long long n_tgt, int H, int C, float slope, T* out, float* stats,
                      const PeerOuts& peers, const StreamExtra& ex, cudaStream_t st, long long v_pitch, long long s_pitch) {
  const size_t smem_bulk = p.smem + (size_t)kStreamWarps * 2 * LB * 32;
  const int push = peers.n == 0 ? 0 : ((push_mode() == 2 && smem_bulk <= 227 * 1024) ? 2 : 1);
  auto kern = push == 0 ? pma_stream_kernel<T, LB, 0> : (push == 1 ? pma_stream_kernel<T, LB, 1> : pma_stream_kernel<T, LB, 2>);
  const size_t smem = push == 2 ? smem_bulk : p.smem;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
  if (e != cudaSuccess) return fail(ALLSET_ECUDA, "pma_stream: smem opt-in: %s", cudaGetErrorString(e));
  kern<<<p.blocks, kStreamWarps * 32, smem, st>>>(v, score, seed, rowptr, col, n_tgt, H, C, slope, p.seg_per_warp,
                                                  p.stages, out, stats, peers,
                                                  v_pitch > 0 ? v_pitch : (long long)LB * 32,
                                                  s_pitch > 0 ? s_pitch : (long long)H * 4, ex.peer_mask, ex.ws);
  return ALLSET_OK;
}

template <typename T>
int pma_stream_typed(const StreamPlan& p, const void* v, const float* score, const float* seed, const int* rowptr,
                     const int* col, long long n_tgt, int H, int C, float slope, void* out, float* stats,
                     const PeerOuts& peers, const StreamExtra& ex, cudaStream_t st, long long v_pitch, long long s_pitch) {
  const T* vi = static_cast<const T*>(v);
  T* oi = static_cast<T*>(out);
  switch (p.lane_bytes) {
    case 4: return launch_pma_stream<T, 4>(p, vi, score, seed, rowptr, col, n_tgt, H, C, slope, oi, stats, peers, ex, st, v_pitch, s_pitch);
    case 8: return launch_pma_stream<T, 8>(p, vi, score, seed, rowptr, col, n_tgt, H, C, slope, oi, stats, peers, ex, st, v_pitch, s_pitch);
    case 16: return launch_pma_stream<T, 16>(p, vi, score, seed, rowptr, col, n_tgt, H, C, slope, oi, stats, peers, ex, st, v_pitch, s_pitch);
    default: return launch_pma_stream<T, 32>(p, vi, score, seed, rowptr, col, n_tgt, H, C, slope, oi, stats, peers, ex, st, v_pitch, s_pitch);
  }
}

bool bad_dtype(int dtype) { return dtype != ALLSET_F32 && dtype != ALLSET_BF16; }
int elem_bytes(int dtype) { return dtype == ALLSET_F32 ? 4 : 2; }

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
namespace {
#include "mlp_tcgen05.cuh"
#include "linear_wgrad.cuh"
}  // namespace

namespace {
// Epilogue 2 can store 32-byte pieces of a thread's own row straight from registers (STG.256) instead of staging the
// tile in shared memory for coalesced stores.  Measured a wash (+-5 % either way depending on the dtype pair), so the
// staged path stays the default; ALLSET_MLP2_DIRECT=1 selects the direct one for A/B runs.
int mlp2_direct(const void* out, int64_t pitch) {
  static const bool direct = getenv("ALLSET_MLP2_DIRECT") != nullptr;
  return (direct && (uintptr_t)out % 32 == 0 && pitch % 32 == 0) ? 1 : 0;
}

template <int MODE>
int mlp2_dispatch(const mlp5::Params& p, int x_dtype, int out_dtype, int32_t d, cudaStream_t st) {
  using bf16 = __nv_bfloat16;
#define ALLSET_MLP2_CASE(D_)                                                                                    \
  if (d == D_) {                                                                                                \
    if (x_dtype == ALLSET_F32 && out_dtype == ALLSET_F32) return mlp5::launch<float, float, D_, MODE>(p, st);   \
    if (x_dtype == ALLSET_F32 && out_dtype == ALLSET_BF16) return mlp5::launch<float, bf16, D_, MODE>(p, st);   \
    if (x_dtype == ALLSET_BF16 && out_dtype == ALLSET_F32) return mlp5::launch<bf16, float, D_, MODE>(p, st);   \
    return mlp5::launch<bf16, bf16, D_, MODE>(p, st);                                                           \
  }
  ALLSET_MLP2_CASE(64)
  ALLSET_MLP2_CASE(128)
#undef ALLSET_MLP2_CASE
  return fail(ALLSET_EUNSUPPORTED, "mlp2: width %d not supported", (int)d);
}
}  // namespace

extern "C" {

int allset_version(void) { return ALLSET_ABI_VERSION; }

int allset_stream_eligible(int dtype, int32_t d, int64_t n_tgt) {
  if (bad_dtype(dtype) || d <= 0 || n_tgt <= 0) return 0;
  // alignment is checked again at launch; a 256-byte aligned dummy stands in for the row pointers here
  return plan_stream(d, elem_bytes(dtype), n_tgt, reinterpret_cast<const void*>(256), reinterpret_cast<const void*>(256),
                     false).ok ? 1 : 0;
}

const char* allset_last_error(void) { return g_err; }

size_t allset_csr_workspace_bytes(int64_t nnz, int64_t n_tgt) {
  if (nnz < 0 || n_tgt < 0 || nnz >= INT32_MAX || n_tgt >= INT32_MAX) return 0;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr,
                                  (int*)nullptr, (int)nnz, 0, bits_for(n_tgt));
  const size_t arr = align_up((size_t)nnz * sizeof(int), 256);
  return 3 * arr + align_up(cub_bytes, 256) + 256;
}

int allset_csr_from_coo(const int64_t* tgt, const int64_t* src, int64_t nnz, int64_t n_tgt, int32_t* rowptr,
                        int32_t* col, int32_t* perm, void* workspace, size_t workspace_bytes, void* stream) {
  if (nnz < 0 || n_tgt < 0) return fail(ALLSET_EINVAL, "csr_from_coo: negative size");
  if (nnz >= INT32_MAX || n_tgt >= INT32_MAX)
    return fail(ALLSET_ERANGE, "csr_from_coo: nnz=%lld / n_tgt=%lld do not fit int32", (long long)nnz,
                (long long)n_tgt);
  if (rowptr == nullptr || (nnz > 0 && (tgt == nullptr || src == nullptr || col == nullptr || perm == nullptr)))
    return fail(ALLSET_EINVAL, "csr_from_coo: null pointer");
  const size_t need = allset_csr_workspace_bytes(nnz, n_tgt);
  if (workspace_bytes < need || (workspace == nullptr && need > 0))
    return fail(ALLSET_EWORKSPACE, "csr_from_coo: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t arr = align_up((size_t)nnz * sizeof(int), 256);
  char* ws = static_cast<char*>(workspace);
  int* keys = reinterpret_cast<int*>(ws);
  int* keys_sorted = reinterpret_cast<int*>(ws + arr);
  int* pos = reinterpret_cast<int*>(ws + 2 * arr);
  void* cub_ws = ws + 3 * arr;
  size_t cub_bytes = workspace_bytes - 3 * arr;
  const unsigned blocks = (unsigned)((nnz + 1 + 255) / 256);
  if (nnz > 0) {
    csr_keys_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(tgt), nnz, keys, pos);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys, keys_sorted, pos, perm, (int)nnz, 0,
                                                    bits_for(n_tgt), st);
    if (e != cudaSuccess) return fail(ALLSET_ECUDA, "csr_from_coo: radix sort: %s", cudaGetErrorString(e));
  }
  csr_fill_kernel<<<blocks, 256, 0, st>>>(keys_sorted, perm, reinterpret_cast<const long long*>(src), nnz, n_tgt,
                                          rowptr, col);
  return check_launch("csr_from_coo");
}

size_t allset_long_segments_workspace_bytes(int64_t n_tgt) {
  if (n_tgt < 0 || n_tgt >= INT32_MAX) return 0;
  size_t bytes = 0;
  cub::CountingInputIterator<int> it(0);
  LongSegment pred{nullptr, 0};
  cub::DeviceSelect::If(nullptr, bytes, it, (int*)nullptr, (int*)nullptr, (int)n_tgt, pred);
  return align_up(bytes, 256) + 256;
}

int allset_long_segments(const int32_t* rowptr, int64_t n_tgt, int32_t threshold, int32_t* long_ids,
                         int32_t* n_long, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_tgt < 0 || n_tgt >= INT32_MAX) return fail(ALLSET_ERANGE, "long_segments: n_tgt out of range");
  if (rowptr == nullptr || n_long == nullptr || (n_tgt > 0 && long_ids == nullptr))
    return fail(ALLSET_EINVAL, "long_segments: null pointer");
  const size_t need = allset_long_segments_workspace_bytes(n_tgt);
  if (workspace_bytes < need || workspace == nullptr)
    return fail(ALLSET_EWORKSPACE, "long_segments: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cub::CountingInputIterator<int> it(0);
  LongSegment pred{rowptr, threshold};
  size_t bytes = workspace_bytes;
  cudaError_t e = cub::DeviceSelect::If(workspace, bytes, it, long_ids, n_long, (int)n_tgt, pred, st);
  if (e != cudaSuccess) return fail(ALLSET_ECUDA, "long_segments: %s", cudaGetErrorString(e));
  return check_launch("long_segments");
}

namespace {
size_t stream_ws_bytes(int32_t d) {
  return 16 + (size_t)kStreamMaxChunks * 4 + 2 * (size_t)kStreamMaxChunks * (size_t)(d + 128) * 4;
}
// the workspace that lets the stream kernels cut long segments; NULL = never cut (long segments then need the buckets)
int check_ws(const char* what, void* ws, size_t ws_bytes, int32_t d) {
  if (ws == nullptr) return ALLSET_OK;
  if (((uintptr_t)ws % 16) != 0 || ws_bytes < stream_ws_bytes(d))
    return fail(ALLSET_EWORKSPACE, "%s: workspace must be 16-byte aligned and >= allset_stream_workspace_bytes(%d) = %zu bytes",
                what, (int)d, stream_ws_bytes(d));
  return ALLSET_OK;
}
}  // namespace

size_t allset_stream_workspace_bytes(int32_t d) { return d > 0 ? stream_ws_bytes(d) : 0; }

static int segreduce_fwd_impl(const void* x, int dtype, int64_t n_src, int32_t d, const int32_t* rowptr,
                              const int32_t* col, const float* w, const float* src_scale, int64_t n_tgt, int op,
                              const int32_t* long_ids, int32_t n_long, int32_t long_threshold, void* out,
                              void* const* peer_outs, int32_t n_peers, const uint8_t* peer_mask, void* ws,
                              size_t ws_bytes, void* stream) {
  if (bad_dtype(dtype)) return fail(ALLSET_EINVAL, "segreduce_fwd: unknown dtype %d", dtype);
  if (op != ALLSET_SUM && op != ALLSET_MEAN) return fail(ALLSET_EINVAL, "segreduce_fwd: unknown op %d", op);
  if (d <= 0 || n_tgt < 0 || n_src < 0 || n_long < 0) return fail(ALLSET_EINVAL, "segreduce_fwd: bad size");
  if (n_tgt >= INT32_MAX || n_src >= INT32_MAX) return fail(ALLSET_ERANGE, "segreduce_fwd: rows do not fit int32");
  if (n_tgt == 0) return ALLSET_OK;
  if (rowptr == nullptr || out == nullptr || (n_long > 0 && long_ids == nullptr))
    return fail(ALLSET_EINVAL, "segreduce_fwd: null pointer");
  if (x == nullptr || col == nullptr) {
    // legal only for an incidence list without entries; rowptr is then all zero and nothing is read
    if (n_src != 0) return fail(ALLSET_EINVAL, "segreduce_fwd: null x/col with n_src > 0");
  }
  if (int rc = check_ws("segreduce_fwd", ws, ws_bytes, d)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // The stream kernel handles every segment length itself: with a workspace it cuts long segments at chunk boundaries
  // (so the long-segment buckets are not needed), without one it takes graphs whose segments are all moderate.
  const StreamPlan sp = plan_stream(d, elem_bytes(dtype), n_tgt, x, out, w != nullptr || src_scale != nullptr);
  PeerOuts peers;
  if (int rc = make_peers(out, peer_outs, n_peers, &peers)) return rc;
  if (sp.ok && (n_long == 0 || ws != nullptr) && x != nullptr && col != nullptr) {
    const StreamExtra ex{peer_mask, ws};
    const int rc = dtype == ALLSET_F32
        ? segreduce_stream_typed<float>(sp, x, rowptr, col, w, src_scale, n_tgt, d, op == ALLSET_MEAN, out, peers, ex, st)
        : segreduce_stream_typed<__nv_bfloat16>(sp, x, rowptr, col, w, src_scale, n_tgt, d, op == ALLSET_MEAN, out, peers, ex, st);
    if (rc != ALLSET_OK) return rc;
    return check_launch("segreduce_fwd(stream)");
  }
  if (n_peers > 0)
    return fail(ALLSET_EUNSUPPORTED, "segreduce_fwd_bcast: shape not eligible for the fused-exchange stream kernel "
                                     "(row bytes %lld, %lld segments, %d long); use segreduce_fwd + an all-gather",
                (long long)d * elem_bytes(dtype), (long long)n_tgt, (int)n_long);
  const Shape sh = plan(d, elem_bytes(dtype), x, out);
  if (dtype == ALLSET_F32)
    segreduce_typed<float>(sh, x, rowptr, col, w, src_scale, n_tgt, d, op == ALLSET_MEAN, long_ids, n_long,
                           long_threshold, out, st);
  else
    segreduce_typed<__nv_bfloat16>(sh, x, rowptr, col, w, src_scale, n_tgt, d, op == ALLSET_MEAN, long_ids, n_long,
                                   long_threshold, out, st);
  return check_launch("segreduce_fwd");
}

int allset_segreduce_fwd(const void* x, int dtype, int64_t n_src, int32_t d, const int32_t* rowptr,
                         const int32_t* col, const float* w, const float* src_scale, int64_t n_tgt, int op,
                         const int32_t* long_ids, int32_t n_long, int32_t long_threshold, void* out,
                         void* ws, size_t ws_bytes, void* stream) {
  return segreduce_fwd_impl(x, dtype, n_src, d, rowptr, col, w, src_scale, n_tgt, op, long_ids, n_long,
                            long_threshold, out, nullptr, 0, nullptr, ws, ws_bytes, stream);
}

int allset_segreduce_fwd_bcast(const void* x, int dtype, int64_t n_src, int32_t d, const int32_t* rowptr,
                               const int32_t* col, const float* w, const float* src_scale, int64_t n_tgt, int op,
                               void* out, void* const* peer_outs, int32_t n_peers, const uint8_t* peer_mask,
                               void* ws, size_t ws_bytes, void* stream) {
  return segreduce_fwd_impl(x, dtype, n_src, d, rowptr, col, w, src_scale, n_tgt, op, nullptr, 0, 0, out, peer_outs,
                            n_peers, peer_mask, ws, ws_bytes, stream);
}


int allset_push_rows(const void* rows, int64_t n_rows, int64_t row_bytes, void* const* peer_rows, int32_t n_peers,
                     const uint8_t* peer_mask, void* stream) {
  if (n_rows < 0 || row_bytes <= 0) return fail(ALLSET_EINVAL, "push_rows: bad size");
  if (n_rows == 0 || n_peers == 0) return ALLSET_OK;
  if (rows == nullptr) return fail(ALLSET_EINVAL, "push_rows: null pointer");
  if (row_bytes % 16 != 0 || ((uintptr_t)rows % 16) != 0)
    return fail(ALLSET_EUNSUPPORTED, "push_rows: rows must be 16-byte aligned multiples of 16 bytes");
  PeerOuts peers;
  if (int rc = make_peers(const_cast<void*>(rows), peer_rows, n_peers, &peers)) return rc;
  const long long n_chunks = (long long)n_rows * (row_bytes / 16);
  long long blocks = (n_chunks + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  push_rows_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(rows), n_chunks, (int)(row_bytes / 16), peers, peer_mask);
  return check_launch("push_rows");
}

int allset_bias_act_norm(const float* x, const float* bias, int relu, const float* residual, const float* gamma,
                         const float* beta, float eps, int64_t rows, int32_t d, float* out, float* stats,
                         void* stream) {
  if (rows < 0 || d <= 0) return fail(ALLSET_EINVAL, "bias_act_norm: bad size");
  if (rows == 0) return ALLSET_OK;
  if (x == nullptr || out == nullptr) return fail(ALLSET_EINVAL, "bias_act_norm: null pointer");
  if (beta != nullptr && gamma == nullptr) return fail(ALLSET_EINVAL, "bias_act_norm: beta without gamma");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((rows + 7) / 8);          // 8 warps = 8 rows per CTA
  const uintptr_t bits = (uintptr_t)x | (uintptr_t)out | (uintptr_t)bias | (uintptr_t)residual | (uintptr_t)gamma |
                         (uintptr_t)beta;
  const bool vec = (d % 128 == 0) && (bits % 16 == 0) && ((uintptr_t)stats % 8 == 0);
  if (vec && d == 128)
    bias_act_norm_kernel<1><<<blocks, 256, 0, st>>>(x, bias, relu, residual, gamma, beta, eps, rows, out, stats);
  else if (vec && d == 256)
    bias_act_norm_kernel<2><<<blocks, 256, 0, st>>>(x, bias, relu, residual, gamma, beta, eps, rows, out, stats);
  else if (vec && d == 512)
    bias_act_norm_kernel<4><<<blocks, 256, 0, st>>>(x, bias, relu, residual, gamma, beta, eps, rows, out, stats);
  else if (vec && d == 1024)
    bias_act_norm_kernel<8><<<blocks, 256, 0, st>>>(x, bias, relu, residual, gamma, beta, eps, rows, out, stats);
  else
    bias_act_norm_generic_kernel<<<blocks, 256, 0, st>>>(x, bias, relu, residual, gamma, beta, eps, rows, d, out, stats);
  return check_launch("bias_act_norm");
}

int32_t allset_bias_act_norm_bwd_blocks(int64_t rows) {
  // persistent-style grid: each warp walks rows with a grid stride so the column sums amortise
  long long b = (rows + 7) / 8;
  const long long cap = 148LL * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int32_t)b;
}

int allset_bias_act_norm_bwd(const float* dy, const float* x, const float* bias, int relu, const float* residual,
                             const float* gamma, const float* stats, int64_t rows, int32_t d, float* dx,
                             float* dres, float* partial, void* stream) {
  if (rows < 0 || d <= 0) return fail(ALLSET_EINVAL, "bias_act_norm_bwd: bad size");
  if (rows == 0) return ALLSET_OK;
  if (dy == nullptr || x == nullptr || dx == nullptr || partial == nullptr)
    return fail(ALLSET_EINVAL, "bias_act_norm_bwd: null pointer");
  if (gamma != nullptr && stats == nullptr) return fail(ALLSET_EINVAL, "bias_act_norm_bwd: stats required with gamma");
  const uintptr_t bits = (uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx | (uintptr_t)bias | (uintptr_t)residual |
                         (uintptr_t)gamma | (uintptr_t)dres | (uintptr_t)partial;
  if (!(d == 128 || d == 256 || d == 512 || d == 1024) || bits % 16 != 0 || (uintptr_t)stats % 8 != 0)
    return fail(ALLSET_EUNSUPPORTED, "bias_act_norm_bwd: needs d in {128,256,512,1024} and 16-byte aligned rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)allset_bias_act_norm_bwd_blocks(rows);
  if (d == 128)
    bias_act_norm_bwd_kernel<1><<<blocks, 256, 0, st>>>(dy, x, bias, relu, residual, gamma, stats, rows, dx, dres, partial);
  else if (d == 256)
    bias_act_norm_bwd_kernel<2><<<blocks, 256, 0, st>>>(dy, x, bias, relu, residual, gamma, stats, rows, dx, dres, partial);
  else if (d == 512)
    bias_act_norm_bwd_kernel<4><<<blocks, 256, 0, st>>>(dy, x, bias, relu, residual, gamma, stats, rows, dx, dres, partial);
  else
    bias_act_norm_bwd_kernel<8><<<blocks, 256, 0, st>>>(dy, x, bias, relu, residual, gamma, stats, rows, dx, dres, partial);
  return check_launch("bias_act_norm_bwd");
}

// ---- rowop: mixed-precision row glue with dropout, forward + backward (rowop.cuh) --------------------------------
namespace {
bool rowop_width_ok(int32_t d) { return d == 64 || d == 128 || d == 256 || d == 512 || d == 1024; }
void rowop_dropout(float p, unsigned* thr16, float* keep_scale) {
  long t = lroundf(p * 65536.f);
  if (t < 0) t = 0;
  if (t > 65535) t = 65535;
  *thr16 = (unsigned)t;
  *keep_scale = 65536.f / (float)(65536 - t);        // 1 / (1 - p) for the p that is actually applied
}
}  // namespace

int allset_rowop_supported(int32_t d) { return rowop_width_ok(d) ? 1 : 0; }

int allset_rowop_fwd(const void* x, int x_dtype, const float* bias, int relu, const void* residual, const float* gamma,
                     const float* beta, float eps, int relu_out, float drop_p, uint64_t seed, int64_t rows, int32_t d,
                     void* out, int out_dtype, float* stats, void* stream) {
  if (rows < 0 || d <= 0) return fail(ALLSET_EINVAL, "rowop_fwd: bad size");
  if (bad_dtype(x_dtype) || bad_dtype(out_dtype)) return fail(ALLSET_EINVAL, "rowop_fwd: dtype must be 0 (f32) or 1 (bf16)");
  if (!(drop_p >= 0.f) || drop_p >= 1.f) return fail(ALLSET_EINVAL, "rowop_fwd: dropout p must be in [0, 1)");
  if (rows == 0) return ALLSET_OK;
  if (x == nullptr || out == nullptr) return fail(ALLSET_EINVAL, "rowop_fwd: null pointer");
  if (beta != nullptr && gamma == nullptr) return fail(ALLSET_EINVAL, "rowop_fwd: beta without gamma");
  const uintptr_t bits = (uintptr_t)x | (uintptr_t)out | (uintptr_t)bias | (uintptr_t)residual | (uintptr_t)gamma |
                         (uintptr_t)beta;
  if (!rowop_width_ok(d) || bits % 16 != 0 || (uintptr_t)stats % 8 != 0)
    return fail(ALLSET_EUNSUPPORTED, "rowop_fwd: needs d in {64,128,256,512,1024} and 16-byte aligned rows");
  rowop::FwdArgs a{x, bias, relu, residual, gamma, beta, eps, 1.f, 0u, seed, rows, out, stats, relu_out};
  if (drop_p > 0.f) rowop_dropout(drop_p, &a.thr16, &a.keep_scale);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  using bf16 = __nv_bfloat16;
  if (x_dtype == ALLSET_F32 && out_dtype == ALLSET_F32) rowop::launch_fwd<float, float>(a, d, st);
  else if (x_dtype == ALLSET_F32) rowop::launch_fwd<float, bf16>(a, d, st);
  else if (out_dtype == ALLSET_F32) rowop::launch_fwd<bf16, float>(a, d, st);
  else rowop::launch_fwd<bf16, bf16>(a, d, st);
  return check_launch("rowop_fwd");
}

int allset_rowop_bwd(const void* dy, int g_dtype, const void* x, int x_dtype, const float* bias, int relu,
                     const void* residual, const float* gamma, const float* beta, const float* stats, int relu_out,
                     float drop_p, uint64_t seed, int64_t rows, int32_t d, void* dx, void* dres, float* partial,
                     void* stream) {
  if (rows < 0 || d <= 0) return fail(ALLSET_EINVAL, "rowop_bwd: bad size");
  if (bad_dtype(x_dtype) || bad_dtype(g_dtype)) return fail(ALLSET_EINVAL, "rowop_bwd: dtype must be 0 (f32) or 1 (bf16)");
  if (!(drop_p >= 0.f) || drop_p >= 1.f) return fail(ALLSET_EINVAL, "rowop_bwd: dropout p must be in [0, 1)");
  if (rows == 0) return ALLSET_OK;
  if (dy == nullptr || x == nullptr || dx == nullptr || partial == nullptr)
    return fail(ALLSET_EINVAL, "rowop_bwd: null pointer");
  if (gamma != nullptr && stats == nullptr) return fail(ALLSET_EINVAL, "rowop_bwd: stats required with gamma");
  const uintptr_t bits = (uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx | (uintptr_t)bias | (uintptr_t)residual |
                         (uintptr_t)gamma | (uintptr_t)dres | (uintptr_t)partial;
  if (!rowop_width_ok(d) || bits % 16 != 0 || (uintptr_t)stats % 8 != 0)
    return fail(ALLSET_EUNSUPPORTED, "rowop_bwd: needs d in {64,128,256,512,1024} and 16-byte aligned rows");
  rowop::BwdArgs a{dy, x, bias, relu, residual, gamma, stats, 1.f, 0u, seed, rows, dx, dres, partial, beta, relu_out};
  if (drop_p > 0.f) rowop_dropout(drop_p, &a.thr16, &a.keep_scale);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)allset_bias_act_norm_bwd_blocks(rows);
  using bf16 = __nv_bfloat16;
  if (g_dtype == ALLSET_F32 && x_dtype == ALLSET_F32) rowop::launch_bwd<float, float>(a, d, blocks, st);
  else if (g_dtype == ALLSET_F32) rowop::launch_bwd<float, bf16>(a, d, blocks, st);
  else if (x_dtype == ALLSET_F32) rowop::launch_bwd<bf16, float>(a, d, blocks, st);
  else rowop::launch_bwd<bf16, bf16>(a, d, blocks, st);
  return check_launch("rowop_bwd");
}

int allset_mlp2_fwd(const void* x, int x_dtype, const float* ln0_gamma, const float* ln0_beta, float ln0_eps,
                    const float* w1, const float* b1, const float* ln1_gamma, const float* ln1_beta, float ln1_eps,
                    const float* w2, const float* b2, int relu_out, int64_t rows, int32_t d, void* out, int out_dtype,
                    int64_t out_pitch, int32_t* status, void* stream) {
  if (rows < 0 || d <= 0) return fail(ALLSET_EINVAL, "mlp2_fwd: bad size");
  if (bad_dtype(x_dtype) || bad_dtype(out_dtype)) return fail(ALLSET_EINVAL, "mlp2_fwd: dtype must be 0 (f32) or 1 (bf16)");
  if (d != 64 && d != 128) return fail(ALLSET_EUNSUPPORTED, "mlp2_fwd: width %d not supported (64 or 128)", (int)d);
  if (rows == 0) return ALLSET_OK;
  if (x == nullptr || out == nullptr || w1 == nullptr) return fail(ALLSET_EINVAL, "mlp2_fwd: null pointer");
  const int single = (w2 == nullptr) ? 1 : 0;        // one Linear: out = [relu](LN0?(x) W1^T + b1)
  if (single && (b2 != nullptr || ln1_gamma != nullptr)) return fail(ALLSET_EINVAL, "mlp2_fwd: w2 == NULL (single Linear) excludes b2 / ln1");
  if ((ln0_beta != nullptr && ln0_gamma == nullptr) || (ln1_beta != nullptr && ln1_gamma == nullptr))
    return fail(ALLSET_EINVAL, "mlp2_fwd: LayerNorm beta without gamma");
  const uintptr_t bits = (uintptr_t)x | (uintptr_t)out | (uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)ln0_gamma | (uintptr_t)ln1_gamma;
  if (bits % 16 != 0) return fail(ALLSET_EUNSUPPORTED, "mlp2_fwd: x, out, w1, w2 and the LayerNorm gammas must be 16-byte aligned");
  const int64_t dense_pitch = (int64_t)d * elem_bytes(out_dtype);
  if (out_pitch == 0) out_pitch = dense_pitch;
  if (out_pitch < dense_pitch || out_pitch % 16 != 0)
    return fail(ALLSET_EINVAL, "mlp2_fwd: out_pitch must be 0 or a multiple of 16 >= the row size");
  mlp5::Params p{x, out, ln0_gamma, ln0_beta, w1, b1, ln1_gamma, ln1_beta, single ? w1 : w2, b2, ln0_eps, ln1_eps, relu_out,
                 single, (long long)rows, status, 0, 0, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, 0,
                 mlp2_direct(out, out_pitch), (long long)out_pitch};
  return mlp2_dispatch<0>(p, x_dtype, out_dtype, d, static_cast<cudaStream_t>(stream));
}

int allset_pma_tail_fwd(const void* x, int x_dtype, const float* ln0_gamma, const float* ln0_beta, float ln0_eps,
                        const float* w1, const float* b1, const float* w2, const float* b2, const float* ln1_gamma,
                        const float* ln1_beta, float ln1_eps, int relu_final, int64_t rows, int32_t d, void* out,
                        int out_dtype, int32_t* status, void* stream) {
  if (rows < 0 || d <= 0) return fail(ALLSET_EINVAL, "pma_tail_fwd: bad size");
  if (bad_dtype(x_dtype) || bad_dtype(out_dtype)) return fail(ALLSET_EINVAL, "pma_tail_fwd: dtype must be 0 (f32) or 1 (bf16)");
  if (d != 64 && d != 128) return fail(ALLSET_EUNSUPPORTED, "pma_tail_fwd: width %d not supported (64 or 128)", (int)d);
  if (rows == 0) return ALLSET_OK;
  if (x == nullptr || out == nullptr || w1 == nullptr || w2 == nullptr || ln0_gamma == nullptr || ln1_gamma == nullptr)
    return fail(ALLSET_EINVAL, "pma_tail_fwd: null pointer");
  const uintptr_t bits = (uintptr_t)x | (uintptr_t)out | (uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)ln0_gamma;
  if (bits % 16 != 0 || (uintptr_t)x % 32 != 0)
    return fail(ALLSET_EUNSUPPORTED, "pma_tail_fwd: out, w1, w2, ln0_gamma must be 16-byte and x 32-byte aligned");
  mlp5::Params p{x, out, ln0_gamma, ln0_beta, w1, b1, nullptr, nullptr, w2, b2, ln0_eps, 1e-5f, 1,
                 0, (long long)rows, status, 1, relu_final, ln1_gamma, ln1_beta, ln1_eps, nullptr, nullptr, nullptr, 0,
                 mlp2_direct(out, (int64_t)d * elem_bytes(out_dtype)), (long long)d * elem_bytes(out_dtype)};
  return mlp2_dispatch<1>(p, x_dtype, out_dtype, d, static_cast<cudaStream_t>(stream));
}


int allset_linear_score_fwd(const void* x, int x_dtype, const float* w, const float* b, const float* w_eff,
                            const float* b_eff, int32_t heads, int64_t rows, int32_t d, void* out, int out_dtype,
                            int64_t out_pitch, float* score, int32_t* status, void* stream) {
  if (rows < 0 || d <= 0 || heads <= 0) return fail(ALLSET_EINVAL, "linear_score_fwd: bad size");
  if (bad_dtype(x_dtype) || bad_dtype(out_dtype)) return fail(ALLSET_EINVAL, "linear_score_fwd: dtype must be 0 (f32) or 1 (bf16)");
  if (d != 64 && d != 128) return fail(ALLSET_EUNSUPPORTED, "linear_score_fwd: width %d not supported (64 or 128)", (int)d);
  if ((int64_t)heads * d * 4 > 4096) return fail(ALLSET_EUNSUPPORTED, "linear_score_fwd: heads * d must be <= 1024");
  if (rows == 0) return ALLSET_OK;
  if (x == nullptr || out == nullptr || w == nullptr || w_eff == nullptr || score == nullptr)
    return fail(ALLSET_EINVAL, "linear_score_fwd: null pointer");
  const uintptr_t bits = (uintptr_t)x | (uintptr_t)out | (uintptr_t)w;
  if (bits % 16 != 0) return fail(ALLSET_EUNSUPPORTED, "linear_score_fwd: x, out, w must be 16-byte aligned");
  const int64_t dense_pitch = (int64_t)d * elem_bytes(out_dtype);
  if (out_pitch == 0) out_pitch = dense_pitch;
  if (out_pitch < dense_pitch || out_pitch % 16 != 0)
    return fail(ALLSET_EINVAL, "linear_score_fwd: out_pitch must be 0 or a multiple of 16 >= the row size");
  mlp5::Params p{x, out, nullptr, nullptr, w, b, nullptr, nullptr, w, nullptr, 1e-5f, 1e-5f, 0,
                 1, (long long)rows, status, 0, 0, nullptr, nullptr, 0.f, w_eff, b_eff, score, (int)heads,
                 mlp2_direct(out, out_pitch), (long long)out_pitch};
  return mlp2_dispatch<2>(p, x_dtype, out_dtype, d, static_cast<cudaStream_t>(stream));
}

int allset_linear_fwd(const void* x, int x_dtype, const float* ln_gamma, const float* ln_beta, float ln_eps,
                      const float* w, int w_transposed, const float* b, int relu, int precision, int64_t rows, int32_t d,
                      void* out, int out_dtype, int32_t* status, void* stream) {
  if (rows < 0 || d <= 0) return fail(ALLSET_EINVAL, "linear_fwd: bad size");
  if (bad_dtype(x_dtype) || bad_dtype(out_dtype)) return fail(ALLSET_EINVAL, "linear_fwd: dtype must be 0 (f32) or 1 (bf16)");
  if (precision != ALLSET_PREC_BF16 && precision != ALLSET_PREC_SPLIT) return fail(ALLSET_EINVAL, "linear_fwd: unknown precision %d", precision);
  if (d != 64 && d != 128) return fail(ALLSET_EUNSUPPORTED, "linear_fwd: width %d not supported (64 or 128)", (int)d);
  if (rows == 0) return ALLSET_OK;
  if (x == nullptr || out == nullptr || w == nullptr) return fail(ALLSET_EINVAL, "linear_fwd: null pointer");
  if (ln_beta != nullptr && ln_gamma == nullptr) return fail(ALLSET_EINVAL, "linear_fwd: LayerNorm beta without gamma");
  if (w_transposed && ln_gamma != nullptr) return fail(ALLSET_EINVAL, "linear_fwd: w_transposed excludes the LayerNorm prologue");
  const uintptr_t bits = (uintptr_t)x | (uintptr_t)out | (uintptr_t)w | (uintptr_t)ln_gamma;
  if (bits % 16 != 0) return fail(ALLSET_EUNSUPPORTED, "linear_fwd: x, out, w and gamma must be 16-byte aligned");
  const int64_t pitch = (int64_t)d * elem_bytes(out_dtype);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == ALLSET_PREC_SPLIT) {
    if (x_dtype != ALLSET_F32 || out_dtype != ALLSET_F32)
      return fail(ALLSET_EUNSUPPORTED, "linear_fwd: split precision takes and returns f32 rows");
    if ((uintptr_t)out % 32 != 0) return fail(ALLSET_EUNSUPPORTED, "linear_fwd: split precision needs a 32-byte aligned out");
    mlp5::Params p{x, out, ln_gamma, ln_beta, w, b, nullptr, nullptr, w, nullptr, ln_eps, 1e-5f, relu, 1, (long long)rows,
                   status, 0, 0, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, 0, 1, (long long)pitch, w_transposed ? 1 : 0};
    if (d == 64) return mlp5::launch<float, float, 64, 3>(p, st);
    return mlp5::launch<float, float, 128, 3>(p, st);
  }
  mlp5::Params p{x, out, ln_gamma, ln_beta, w, b, nullptr, nullptr, w, nullptr, ln_eps, 1e-5f, relu, 1, (long long)rows,
                 status, 0, 0, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, 0, mlp2_direct(out, pitch), (long long)pitch,
                 w_transposed ? 1 : 0};
  return mlp2_dispatch<0>(p, x_dtype, out_dtype, d, st);
}

int allset_linear_wgrad_partials(int64_t rows) { return rows <= 0 ? 0 : wgrad5::n_partials((long long)rows); }

int allset_linear_wgrad(const void* dy, const void* x, int dtype, int precision, int64_t rows, int32_t d, float* dw,
                        float* workspace, int64_t workspace_floats, int32_t* status, void* stream) {
  if (rows < 0 || d <= 0) return fail(ALLSET_EINVAL, "linear_wgrad: bad size");
  if (bad_dtype(dtype)) return fail(ALLSET_EINVAL, "linear_wgrad: dtype must be 0 (f32) or 1 (bf16)");
  if (precision != ALLSET_PREC_BF16 && precision != ALLSET_PREC_SPLIT) return fail(ALLSET_EINVAL, "linear_wgrad: unknown precision %d", precision);
  if (d != 64 && d != 128) return fail(ALLSET_EUNSUPPORTED, "linear_wgrad: width %d not supported (64 or 128)", (int)d);
  if (dw == nullptr) return fail(ALLSET_EINVAL, "linear_wgrad: null pointer");
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    cudaMemsetAsync(dw, 0, (size_t)d * d * sizeof(float), st);
    return check_launch("linear_wgrad");
  }
  if (dy == nullptr || x == nullptr || workspace == nullptr) return fail(ALLSET_EINVAL, "linear_wgrad: null pointer");
  if ((dtype == ALLSET_F32) != (precision == ALLSET_PREC_SPLIT))
    return fail(ALLSET_EUNSUPPORTED, "linear_wgrad: f32 rows take split precision, bf16 rows bf16 precision");
  const uintptr_t bits = (uintptr_t)dy | (uintptr_t)x | (uintptr_t)workspace;
  if (bits % 32 != 0) return fail(ALLSET_EUNSUPPORTED, "linear_wgrad: dy, x and the workspace must be 32-byte aligned");
  if (workspace_floats < (int64_t)wgrad5::n_partials((long long)rows) * d * d)
    return fail(ALLSET_EINVAL, "linear_wgrad: workspace too small (allset_linear_wgrad_partials(rows) * d * d floats)");
  static const int swap = getenv("ALLSET_WGRAD_SWAP") != nullptr ? 1 : 0;
  static const int ahead = getenv("ALLSET_WGRAD_L2_PREFETCH") != nullptr ? atoi(getenv("ALLSET_WGRAD_L2_PREFETCH")) : 2;
  wgrad5::Params p{dy, x, (long long)rows, workspace, status, swap, ahead};
  if (dtype == ALLSET_F32) {
    if (d == 64) return wgrad5::launch<float, 64, true>(p, dw, st);
    return wgrad5::launch<float, 128, true>(p, dw, st);
  }
  if (d == 64) return wgrad5::launch<__nv_bfloat16, 64, false>(p, dw, st);
  return wgrad5::launch<__nv_bfloat16, 128, false>(p, dw, st);
}

int allset_segreduce_bwd_w(const void* x, const void* grad_out, int dtype, int32_t d, const int32_t* rowptr,
                           const int32_t* col, const float* tgt_scale, int64_t n_tgt, float* grad_w,
                           void* stream) {
  if (bad_dtype(dtype)) return fail(ALLSET_EINVAL, "segreduce_bwd_w: unknown dtype %d", dtype);
  if (d <= 0 || n_tgt < 0) return fail(ALLSET_EINVAL, "segreduce_bwd_w: bad size");
  if (n_tgt == 0) return ALLSET_OK;
  if (x == nullptr || grad_out == nullptr || rowptr == nullptr || col == nullptr || grad_w == nullptr)
    return fail(ALLSET_EINVAL, "segreduce_bwd_w: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = blocks_for(n_tgt, 32);
  if (dtype == ALLSET_F32)
    segreduce_bwd_w_kernel<float><<<blocks, kThreads, 0, st>>>(static_cast<const float*>(x),
                                                              static_cast<const float*>(grad_out), rowptr, col,
                                                              tgt_scale, n_tgt, d, grad_w);
  else
    segreduce_bwd_w_kernel<__nv_bfloat16><<<blocks, kThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(grad_out), rowptr, col, tgt_scale,
        n_tgt, d, grad_w);
  return check_launch("segreduce_bwd_w");
}

namespace {
// whether (dtype, H, C, n_tgt) takes the PMA stream kernel: the sum kernel's conditions plus H % 4 == 0 (scores travel
// as 16-byte pieces), no lane chunk straddling two heads, and the score ring fitting shared memory
bool pma_stream_shape_ok(int dtype, int32_t H, int32_t C, int64_t n_tgt, StreamPlan* out_plan) {
  alignas(256) static char dummy[256];
  StreamPlan sp = plan_stream(H * C, elem_bytes(dtype), n_tgt, dummy, dummy, false);
  if (!sp.ok) return false;
  const int chunk_elems = (sp.lane_bytes >= 16 ? 16 : sp.lane_bytes) / elem_bytes(dtype);
  const size_t stage_rows = (size_t)kStreamWarps * sp.stages * sp.rows_per_stage;
  sp.smem += stage_rows * (size_t)H * 4;
  if (H % 4 != 0 || C % chunk_elems != 0 || sp.smem > 227 * 1024) return false;
  if (out_plan != nullptr) *out_plan = sp;
  return true;
}
}  // namespace

int allset_pma_stream_eligible(int dtype, int32_t H, int32_t C, int64_t n_tgt) {
  if (bad_dtype(dtype) || H <= 0 || C <= 0 || n_tgt <= 0) return 0;
  return pma_stream_shape_ok(dtype, H, C, n_tgt, nullptr) ? 1 : 0;
}

static int pma_fwd_impl(const void* v, const float* score, const float* seed, int dtype, int32_t H, int32_t C,
                        float slope, const int32_t* rowptr, const int32_t* col, int64_t n_tgt,
                        const int32_t* long_ids, int32_t n_long, int32_t long_threshold, void* out, float* stats,
                        void* const* peer_outs, int32_t n_peers, const uint8_t* peer_mask, void* ws, size_t ws_bytes,
                        void* stream, int64_t v_pitch = 0, int64_t s_pitch = 0) {
  if (bad_dtype(dtype)) return fail(ALLSET_EINVAL, "pma_fwd: unknown dtype %d", dtype);
  const bool strided = v_pitch != 0 || s_pitch != 0;       // rows / scores with a byte pitch: stream kernel only
  if (strided && (v_pitch < (int64_t)H * C * elem_bytes(dtype) || s_pitch < (int64_t)H * 4 || v_pitch % 16 != 0 ||
                  s_pitch % 16 != 0))
    return fail(ALLSET_EINVAL, "pma_fwd_strided: pitches must cover a row / a score record and be multiples of 16");
  if (H <= 0 || C <= 0 || n_tgt < 0 || n_long < 0) return fail(ALLSET_EINVAL, "pma_fwd: bad size");
  if (!(slope == slope) || slope > 3.0e38f || slope < -3.0e38f)   // the reference's PMA takes any finite slope, 0 included
    return fail(ALLSET_EINVAL, "pma_fwd: negative_slope must be finite");
  if (n_tgt >= INT32_MAX) return fail(ALLSET_ERANGE, "pma_fwd: rows do not fit int32");
  if (n_tgt == 0) return ALLSET_OK;
  if (seed == nullptr || rowptr == nullptr || out == nullptr || (n_long > 0 && long_ids == nullptr))
    return fail(ALLSET_EINVAL, "pma_fwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int d = H * C;
  if (int rc = check_ws("pma_fwd", ws, ws_bytes, d)) return rc;
  {
    StreamPlan sp{};
    sp.ok = pma_stream_shape_ok(dtype, H, C, n_tgt, &sp);
    if (sp.ok && (((uintptr_t)v % 16) != 0 || ((uintptr_t)out % 16) != 0 || ((uintptr_t)score % 16) != 0 ||
                  (n_long != 0 && ws == nullptr) || v == nullptr || col == nullptr || score == nullptr))
      sp.ok = false;
    PeerOuts peers;
    if (int rc = make_peers(out, peer_outs, n_peers, &peers)) return rc;
    if (sp.ok) {
      const StreamExtra ex{peer_mask, ws};
      const int rc = dtype == ALLSET_F32
          ? pma_stream_typed<float>(sp, v, score, seed, rowptr, col, n_tgt, H, C, slope, out, stats, peers, ex, st, v_pitch, s_pitch)
          : pma_stream_typed<__nv_bfloat16>(sp, v, score, seed, rowptr, col, n_tgt, H, C, slope, out, stats, peers, ex, st, v_pitch, s_pitch);
      if (rc != ALLSET_OK) return rc;
      return check_launch("pma_fwd(stream)");
    }
    if (n_peers > 0)
      return fail(ALLSET_EUNSUPPORTED, "pma_fwd_bcast: shape not eligible for the fused-exchange stream kernel; use "
                                       "pma_fwd + an all-gather");
    if (strided)
      return fail(ALLSET_EUNSUPPORTED, "pma_fwd_strided: shape not eligible for the stream kernel (see "
                                       "allset_stream_eligible); use dense arrays with allset_pma_fwd");
  }
  Shape sh = plan(d, elem_bytes(dtype), v, out);
  if (sh.vector && C % (16 / elem_bytes(dtype)) != 0) {   // a 16-byte chunk would straddle two heads
    sh.vector = false;
    sh.G = 1;
    while (sh.G < 32 && sh.G < d) sh.G <<= 1;
    sh.slabs = (d + 31) / 32;
  }
  if (dtype == ALLSET_F32) {
    const float* vi = static_cast<const float*>(v);
    float* oi = static_cast<float*>(out);
    if (sh.vector) launch_pma_fwd<float, true>(sh, vi, score, seed, rowptr, col, n_tgt, H, C, slope, long_ids, n_long, long_threshold, oi, stats, st);
    else launch_pma_fwd<float, false>(sh, vi, score, seed, rowptr, col, n_tgt, H, C, slope, long_ids, n_long, long_threshold, oi, stats, st);
  } else {
    const __nv_bfloat16* vi = static_cast<const __nv_bfloat16*>(v);
    __nv_bfloat16* oi = static_cast<__nv_bfloat16*>(out);
    if (sh.vector) launch_pma_fwd<__nv_bfloat16, true>(sh, vi, score, seed, rowptr, col, n_tgt, H, C, slope, long_ids, n_long, long_threshold, oi, stats, st);
    else launch_pma_fwd<__nv_bfloat16, false>(sh, vi, score, seed, rowptr, col, n_tgt, H, C, slope, long_ids, n_long, long_threshold, oi, stats, st);
  }
  return check_launch("pma_fwd");
}

int allset_pma_fwd(const void* v, const float* score, const float* seed, int dtype, int32_t H, int32_t C,
                   float slope, const int32_t* rowptr, const int32_t* col, int64_t n_tgt,
                   const int32_t* long_ids, int32_t n_long, int32_t long_threshold, void* out, float* stats,
                   void* ws, size_t ws_bytes, void* stream) {
  return pma_fwd_impl(v, score, seed, dtype, H, C, slope, rowptr, col, n_tgt, long_ids, n_long, long_threshold, out,
                      stats, nullptr, 0, nullptr, ws, ws_bytes, stream);
}

int allset_pma_fwd_strided(const void* v, int64_t v_pitch, const float* score, int64_t s_pitch, const float* seed,
                           int dtype, int32_t H, int32_t C, float slope, const int32_t* rowptr, const int32_t* col,
                           int64_t n_tgt, void* out, float* stats, void* ws, size_t ws_bytes, void* stream) {
  if (v_pitch <= 0 || s_pitch <= 0) return fail(ALLSET_EINVAL, "pma_fwd_strided: pitches must be positive");
  return pma_fwd_impl(v, score, seed, dtype, H, C, slope, rowptr, col, n_tgt, nullptr, 0, 0, out, stats, nullptr, 0,
                      nullptr, ws, ws_bytes, stream, v_pitch, s_pitch);
}

int allset_pma_fwd_bcast(const void* v, const float* score, const float* seed, int dtype, int32_t H, int32_t C,
                         float slope, const int32_t* rowptr, const int32_t* col, int64_t n_tgt, void* out,
                         float* stats, void* const* peer_outs, int32_t n_peers, const uint8_t* peer_mask, void* ws,
                         size_t ws_bytes, void* stream) {
  return pma_fwd_impl(v, score, seed, dtype, H, C, slope, rowptr, col, n_tgt, nullptr, 0, 0, out, stats, peer_outs,
                      n_peers, peer_mask, ws, ws_bytes, stream);
}

int allset_pma_alpha(const float* score, const float* stats, int32_t H, float slope, const int32_t* rowptr,
                     const int32_t* col, int64_t n_tgt, float* alpha, void* stream) {
  if (H <= 0 || n_tgt < 0) return fail(ALLSET_EINVAL, "pma_alpha: bad size");
  if (n_tgt == 0) return ALLSET_OK;
  if (stats == nullptr || rowptr == nullptr) return fail(ALLSET_EINVAL, "pma_alpha: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pma_alpha_kernel<<<blocks_for(n_tgt, 32), kThreads, 0, st>>>(score, stats, rowptr, col, n_tgt, H, slope, alpha);
  return check_launch("pma_alpha");
}

int allset_rowdot_heads(const void* a, const void* b, const float* sub, int dtype, int64_t n_rows, int32_t H,
                        int32_t C, float* out, void* stream) {
  if (bad_dtype(dtype)) return fail(ALLSET_EINVAL, "rowdot_heads: unknown dtype %d", dtype);
  if (H <= 0 || C <= 0 || n_rows < 0) return fail(ALLSET_EINVAL, "rowdot_heads: bad size");
  if (n_rows == 0) return ALLSET_OK;
  if (a == nullptr || b == nullptr || out == nullptr) return fail(ALLSET_EINVAL, "rowdot_heads: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int d = H * C;
  const Shape sh = plan(d, elem_bytes(dtype), a, b);
  const bool grouped = sh.vector && sh.slabs == 1 && C % (16 / elem_bytes(dtype)) == 0;
  if (grouped) {
    ALLSET_DISPATCH_G(sh.G, {
      if (dtype == ALLSET_F32)
        rowdot_group_kernel<float, true, G><<<blocks_for(n_rows, G), kThreads, 0, st>>>(
            static_cast<const float*>(a), static_cast<const float*>(b), sub, n_rows, H, C, out);
      else
        rowdot_group_kernel<__nv_bfloat16, true, G><<<blocks_for(n_rows, G), kThreads, 0, st>>>(
            static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), sub, n_rows, H, C, out);
    });
  } else {
    const long long total = n_rows * H;
    const unsigned blocks = (unsigned)((total + kThreads - 1) / kThreads);
    if (dtype == ALLSET_F32)
      rowdot_thread_kernel<float><<<blocks, kThreads, 0, st>>>(static_cast<const float*>(a),
                                                              static_cast<const float*>(b), sub, n_rows, H, C, out);
    else
      rowdot_thread_kernel<__nv_bfloat16><<<blocks, kThreads, 0, st>>>(
          static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), sub, n_rows, H, C, out);
  }
  return check_launch("rowdot_heads");
}

int allset_pma_bwd(const void* grad_out, const void* v, const float* score, const float* stats, const float* D,
                   int dtype, int32_t H, int32_t C, float slope, const int32_t* rowptrT, const int32_t* colT,
                   int64_t n_src, const int32_t* long_ids, int32_t n_long, int32_t long_threshold, void* grad_v,
                   float* grad_score, void* stream) {
  if (bad_dtype(dtype)) return fail(ALLSET_EINVAL, "pma_bwd: unknown dtype %d", dtype);
  if (H <= 0 || C <= 0 || n_src < 0 || n_long < 0) return fail(ALLSET_EINVAL, "pma_bwd: bad size");
  if (n_src >= INT32_MAX) return fail(ALLSET_ERANGE, "pma_bwd: rows do not fit int32");
  if (n_src == 0) return ALLSET_OK;
  if (v == nullptr || score == nullptr || rowptrT == nullptr || grad_v == nullptr || grad_score == nullptr ||
      (n_long > 0 && long_ids == nullptr))
    return fail(ALLSET_EINVAL, "pma_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int d = H * C;
  Shape sh = plan(d, elem_bytes(dtype), grad_out, v, grad_v);
  if (sh.vector && C % (16 / elem_bytes(dtype)) != 0) {
    sh.vector = false;
    sh.G = 1;
    while (sh.G < 32 && sh.G < d) sh.G <<= 1;
    sh.slabs = (d + 31) / 32;
  }
  const bool fused = sh.slabs == 1;
  if (dtype == ALLSET_F32)
    pma_bwd_typed<float>(sh, fused, grad_out, v, score, stats, D, rowptrT, colT, n_src, H, C, slope, long_ids, n_long,
                         long_threshold, grad_v, grad_score, st);
  else
    pma_bwd_typed<__nv_bfloat16>(sh, fused, grad_out, v, score, stats, D, rowptrT, colT, n_src, H, C, slope, long_ids,
                                 n_long, long_threshold, grad_v, grad_score, st);
  return check_launch("pma_bwd");
}

}  // extern "C"
