// linear_wgrad.cuh -- weight gradient of a Linear on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Included by allset_kernels.cu inside its anonymous namespace, after mlp_tcgen05.cuh (PTX wrappers from mlp5::).
//
//     dW[n, k] = sum_r dY[r, n] * X[r, k]          dY, X [rows, D] row-major (fp32 | bf16), dW [D, D] fp32
//
// = the `grad_weight` of every nn.Linear of the reference's MLP / PMA (src/layers.py:571-579, :128-130, :76-80) in the
// training loop (src/train.py:478-482).  The reduction runs over the ROWS (10^6 .. 10^7 of them), the output is one
// D x D tile: the kernel is a stream over both inputs, HBM-bound (2 * D * s bytes per row against 2 * D * D flops).
//
// Operands without a transpose: a tile of 64 rows of dY is, as it lies in memory, the "MN-major" A operand of
// D = A B^T (M index = column of dY = contiguous, K index = row), and the same tile of X is the MN-major B operand.
// The producers copy rows into the canonical MN-major SWIZZLE_128B layout (64-element x 8-row atoms, 16-byte chunk index
// XOR row & 7; atoms ordered [8-row group][64-column block], i.e. LBO = 1024 B between column blocks, SBO = 2048 B
// between row groups -- cute's tile_to_shape(Layout_MN_SW128_Atom, (128, 64))) and one thread issues
// tcgen05.mma M=128, N=D, K=16 with both major bits set.  Every CTA accumulates its share of the row tiles in ONE TMEM
// accumulator and writes a D x D partial; a second small kernel adds the partials in a fixed order (deterministic,
// no atomics).
//
// D = 64: tcgen05 M=64 uses a different TMEM data path, so the A operand is the 128-column concatenation [dY | X] of the
// two 64-column tiles (they sit in adjacent column blocks of the stage anyway) and B is the X block: rows 0..63 of the
// accumulator are dY^T X, rows 64..127 (X^T X) are ignored.
//
// SPLIT (fp32 mode): both operands as two bf16 terms, dW ~= dYl Xh + dYh Xl + dYh Xh (see mlp2_ws_kernel MODE 3).

namespace wgrad5 {

using mlp5::ld_nc_32;
using mlp5::mbar_arrive;
using mlp5::mbar_wait_bounded;
using mlp5::pack_bf16;
using mlp5::proxy_fence_async;
using mlp5::st_global32;
using mlp5::st_shared16;
using mlp5::tc_fence_after;
using mlp5::tc_fence_before;
using mlp5::tmem_alloc;
using mlp5::tmem_dealloc;
using mlp5::tmem_ld32;
using mlp5::umma_bf16;
using mlp5::umma_commit;

constexpr int kKT = 64;                 // rows (reduction index) per stage
constexpr int kStages = 3;
constexpr int kEpiWarps = 4;            // TMEM lane quadrants
constexpr int kProdWarps = 8;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;

struct Params {
  const void* dy;
  const void* x;
  long long rows;
  float* partial;     // [gridDim.x, D, D]
  int* status;
  int swap_offsets;   // debug: exchange LBO and SBO in the descriptors
  int l2_prefetch_ahead;   // tiles ahead the producers prefetch into L2 (0 = off)
};

// MN-major SWIZZLE_128B descriptor: start address >> 4, LBO (between 64-element column blocks) >> 4 in [16,30),
// SBO (between 8-row groups) >> 4 in [32,46), version 1 in [46,48), layout type 2 in [61,64)
__device__ __forceinline__ uint64_t smem_desc_mn128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor, D fp32, A/B bf16, both MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t instr_desc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int D, bool SPLIT>
struct Geo {
  static constexpr int PART = (D == 128) ? 2 * kKT * 256 : kKT * 256;   // one precision term of a stage: dY tile + X tile
  static constexpr int STAGE = PART * (SPLIT ? 2 : 1);
  static constexpr int SMEM = 1024 + kStages * STAGE + 128;
  static constexpr int TMEM_COLS = (D <= 32) ? 32 : (D <= 64) ? 64 : 128;
  // byte offset of the 16-byte chunk (row kk, columns [8 c8, 8 c8 + 8)) of tensor t (0 = dY, 1 = X) inside a part
  __device__ static __forceinline__ uint32_t chunk(int t, int kk, int c8) {
    if (D == 128)
      return (uint32_t)(t * (kKT * 256) + (kk >> 3) * 2048 + (c8 >> 3) * 1024 + (kk & 7) * 128 + (((c8 & 7) ^ (kk & 7)) << 4));
    return (uint32_t)((kk >> 3) * 2048 + t * 1024 + (kk & 7) * 128 + ((c8 ^ (kk & 7)) << 4));
  }
};

template <typename T>
struct Row8;   // 8 consecutive elements of a row
template <>
struct Row8<float> {
  uint4 a, b;
  __device__ __forceinline__ void load(const unsigned char* p) { ld_nc_32(p, a, b); }
  __device__ __forceinline__ void zero() { a = b = make_uint4(0, 0, 0, 0); }
  template <bool SPLIT>
  __device__ __forceinline__ void store(uint32_t hi_addr, uint32_t lo_addr) const {
    const float v[8] = {__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w),
                        __uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w)};
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      h[e / 2] = pack_bf16(v[e], v[e + 1]);
      if (SPLIT) l[e / 2] = pack_bf16(v[e] - __uint_as_float(h[e / 2] << 16), v[e + 1] - __uint_as_float(h[e / 2] & 0xFFFF0000u));
    }
    st_shared16(hi_addr, h[0], h[1], h[2], h[3]);
    if (SPLIT) st_shared16(lo_addr, l[0], l[1], l[2], l[3]);
  }
};
template <>
struct Row8<__nv_bfloat16> {
  uint4 a;
  __device__ __forceinline__ void load(const unsigned char* p) { a = ld_nc_16(p); }
  __device__ __forceinline__ void zero() { a = make_uint4(0, 0, 0, 0); }
  template <bool SPLIT>
  __device__ __forceinline__ void store(uint32_t hi_addr, uint32_t lo_addr) const {
    st_shared16(hi_addr, a.x, a.y, a.z, a.w);
    if (SPLIT) st_shared16(lo_addr, 0u, 0u, 0u, 0u);          // a bf16 row has no low-order term
  }
};

template <typename T, int D, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const Params p) {
  using G = Geo<D, SPLIT>;
  static_assert(D == 64 || D == 128, "wgrad: widths 64 and 128");
  constexpr int LPR = D / 8;                  // lanes per row (8 elements each)
  constexpr int RPI = 32 / LPR;               // rows per warp instruction
  constexpr int NI = (kKT / kProdWarps) / RPI;   // loads per tensor, lane and stage

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t sStage = raw_addr + pad;
  const uint32_t sBar = sStage + kStages * G::STAGE;                 // full[kStages], empty[kStages], acc_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * G::STAGE + 96);
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * kStages, bar_acc = sBar + 16 * kStages;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, kProdWarps);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), G::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long n_tiles = (p.rows + kKT - 1) / kKT;

  if (warp > kEpiWarps) {
    // ======================= producers: rows -> MN-major swizzled tiles ===========================================
    const int pw = warp - kEpiWarps - 1;
    const int sub = lane / LPR, c8 = lane % LPR;
    const unsigned char* yb = static_cast<const unsigned char*>(p.dy) + (size_t)c8 * 8 * sizeof(T);
    const unsigned char* xb = static_cast<const unsigned char*>(p.x) + (size_t)c8 * 8 * sizeof(T);
    Row8<T> bufY[NI], bufX[NI];
    auto load = [&](Row8<T>(&buf)[NI], const unsigned char* base, long long tile) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const long long gr = tile * kKT + pw * (kKT / kProdWarps) + i * RPI + sub;
        if (gr < p.rows) buf[i].load(base + (size_t)gr * (D * sizeof(T)));
        else buf[i].zero();
      }
    };
    auto store = [&](const Row8<T>(&buf)[NI], int t, uint32_t stage) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int kk = pw * (kKT / kProdWarps) + i * RPI + sub;
        const uint32_t off = G::chunk(t, kk, c8);
        buf[i].template store<SPLIT>(stage + off, stage + G::PART + off);
      }
    };
    long long tile = blockIdx.x;
    if (tile < n_tiles) {
      load(bufY, yb, tile);
      load(bufX, xb, tile);
    }
    for (uint32_t it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t s = it % kStages;
      const uint32_t stage = sStage + s * G::STAGE;
      const long long next = tile + gridDim.x;
      {   // L2 prefetch a few tiles ahead (the register prefetch reaches one tile = 32-64 KB per SM, short of HBM latency x rate)
        const long long r0 = (tile + (long long)p.l2_prefetch_ahead * gridDim.x) * kKT;
        if (p.l2_prefetch_ahead > 0 && r0 < p.rows) {
          const long long nrows = (p.rows - r0) < kKT ? (p.rows - r0) : kKT;
          const int lines = (int)((nrows * (long long)(D * sizeof(T)) + 127) >> 7);
          const size_t off = (size_t)r0 * (D * sizeof(T));
          for (int i = pw * 32 + lane; i < lines; i += kProdWarps * 32) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const unsigned char*>(p.dy) + off + ((size_t)i << 7)));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const unsigned char*>(p.x) + off + ((size_t)i << 7)));
          }
        }
      }
      if (it >= kStages) mbar_wait_bounded<1000>(bar_empty + 8 * s, ((it / kStages) - 1) & 1u, p.status);
      store(bufY, 0, stage);
      if (next < n_tiles) load(bufY, yb, next);
      store(bufX, 1, stage);
      if (next < n_tiles) load(bufX, xb, next);
      proxy_fence_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
    }
  } else if (warp == kEpiWarps) {
    // ======================= MMA issuer (one thread) ==============================================================
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16_mn(128, D);
      const uint32_t lbo = p.swap_offsets ? 2048u : 1024u, sbo = p.swap_offsets ? 1024u : 2048u;
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % kStages;
        const uint32_t stage = sStage + s * G::STAGE;
        mbar_wait_bounded(bar_full + 8 * s, (it / kStages) & 1u, p.status);
        tc_fence_after();
        constexpr int NT = SPLIT ? 3 : 1;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          // SPLIT: dY_lo X_hi, dY_hi X_lo, dY_hi X_hi
          const uint32_t pa = stage + ((SPLIT && t == 0) ? G::PART : 0);
          const uint32_t pb = stage + ((SPLIT && t == 1) ? G::PART : 0) + (D == 128 ? kKT * 256 : 1024);
#pragma unroll
          for (int ks = 0; ks < kKT / 16; ++ks)
            umma_bf16(tmem_base, smem_desc_mn128(pa + ks * 4096, lbo, sbo), smem_desc_mn128(pb + ks * 4096, lbo, sbo),
                      idesc, (it > 0 || t > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * s);
      }
      umma_commit(bar_acc);
    }
    __syncwarp();
  } else {
    // ======================= epilogue: TMEM -> this CTA's partial =================================================
    mbar_wait_bounded<2000>(bar_acc, 0u, p.status);
    tc_fence_after();
    const int r = tid;                                     // TMEM lane = output row n
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float* dst = p.partial + ((size_t)blockIdx.x * D + r) * D;
#pragma unroll 1
    for (int c = 0; c < D; c += 32) {
      float v[32];
      tmem_ld32(tmem_base + lane_off + c, v);
      if (r < D) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          st_global32(dst + c + 8 * q, __float_as_uint(v[8 * q]), __float_as_uint(v[8 * q + 1]), __float_as_uint(v[8 * q + 2]),
                      __float_as_uint(v[8 * q + 3]), __float_as_uint(v[8 * q + 4]), __float_as_uint(v[8 * q + 5]),
                      __float_as_uint(v[8 * q + 6]), __float_as_uint(v[8 * q + 7]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, G::TMEM_COLS);
}

// dw[i] = sum_c partial[c][i] in a fixed order
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int n_part, int n, float* __restrict__ dw) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = 0;
  for (; c + 4 <= n_part; c += 4) {
    s0 += partial[(size_t)c * n + i];
    s1 += partial[(size_t)(c + 1) * n + i];
    s2 += partial[(size_t)(c + 2) * n + i];
    s3 += partial[(size_t)(c + 3) * n + i];
  }
  for (; c < n_part; ++c) s0 += partial[(size_t)c * n + i];
  dw[i] = (s0 + s1) + (s2 + s3);
}

inline int n_partials(long long rows) {
  const long long n_tiles = (rows + kKT - 1) / kKT;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)(n_tiles < sms ? n_tiles : sms);
}

template <typename T, int D, bool SPLIT>
int launch(const Params& p, float* dw, cudaStream_t st) {
  constexpr int smem = Geo<D, SPLIT>::SMEM;
  cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<T, D, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return fail(ALLSET_ECUDA, "linear_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int grid = n_partials(p.rows);
  wgrad_kernel<T, D, SPLIT><<<(unsigned)grid, kThreads, smem, st>>>(p);
  int rc = check_launch("linear_wgrad");
  if (rc != ALLSET_OK) return rc;
  wgrad_reduce_kernel<<<(D * D + 255) / 256, 256, 0, st>>>(p.partial, grid, D * D, dw);
  return check_launch("linear_wgrad_reduce");
}

}  // namespace wgrad5
