// mlp_tcgen05.cuh -- fused two-layer MLP on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Included by allset_kernels.cu inside its anonymous namespace (uses smem_u32 / mbar_init / ld_nc_16 from there).
//
// The dense glue around the aggregation is, per half layer, the reference's
//     MLP.forward  (src/layers.py:571-579):  norm0 -> Linear -> ReLU -> norm -> dropout -> Linear
// wrapped in F.relu by HalfNLHconv.forward (src/layers.py:631,634).  In eval mode with equal widths
// (in = hid = out = D, the shape of 3 of the 4 MLPs of every AllDeepSets layer) this kernel does the whole
// chain in ONE pass over the rows:
//
//     out[r,:] = [relu]( LN1?( relu( LN0?(x[r,:]) W1^T + b1 ) ) W2^T + b2 )
//
// HBM traffic = read x once + write out once (the separate LayerNorm / bias / ReLU passes and the hidden
// activation never touch HBM), so the kernel is HBM-bound: 768 B per row at D=128 fp32-in / bf16-out against
// 2 x 65.5 kFLOP per row -- 170 flop/B, below the bf16 ridge (~250 flop/B), which is why bf16 operands on
// tcgen05 are the right tool: fp32 SIMT FMA (75 TFLOP/s) would be 3x over the HBM time.
//
// Operand layout: W1, W2 (fp32 [D, D], nn.Linear layout = K-major "B" operand) are converted once per CTA to bf16 in
// shared memory in the canonical K-major SWIZZLE_128B UMMA layout (8-row x 128-byte atoms, 16-byte chunk index XOR
// row & 7); the activation tiles (128 rows) use the same layout.  One tcgen05.mma is M=128, N=D, K=16 (kind::f16,
// bf16 x bf16 -> fp32 in TMEM); a GEMM of the tile is D/16 of them + one tcgen05.commit -> mbarrier, issued by ONE
// thread.  Epilogues read the accumulator with tcgen05.ld.32x32b.x32 (thread = row = TMEM lane).  The pipeline
// (warp roles, barriers) is described above mlp2_ws_kernel.

namespace mlp5 {

constexpr int kTileM = 128;

struct Params {
  const void* x;
  void* out;
  const float* ln0_g;
  const float* ln0_b;
  const float* w1;
  const float* b1;
  const float* ln1_g;
  const float* ln1_b;
  const float* w2;
  const float* b2;
  float eps0, eps1;
  int relu_out;
  int single;       // 1: ONE Linear only (out = [relu](LN0?(x) W1^T + b1)); w2 / b2 / ln1 unused (warp-specialised kernel)
  long long rows;
  int* status;      // device word, set to 1 if an mbarrier wait timed out (debug aid; never in a correct run)
  // tail != 0 (PMA tail, src/layers.py:155-157):  out = [relu_final]( LNf( y + relu(MLP(y)) ) ), y = LN0(x) WITH its
  // affine part; ln1 must be absent, ln0 present.  lnf_* = the LayerNorm after the residual (PMA.ln1).
  int tail;
  int relu_final;
  const float* lnf_g;
  const float* lnf_b;
  float epsf;
  // MODE 2 (PMA.lin_V + the folded lin_K score, src/layers.py:128-130): single Linear AND, in fp32 on the CUDA cores
  // of the producer warps, score[r, h] = <x[r, :], w_eff[h, :]> + b_eff[h] for H heads
  const float* w_eff;    // [H, D] f32
  const float* b_eff;    // [H] f32 or NULL
  float* score;          // [rows, H] f32
  int heads;
  int direct;            // 1: epilogue 2 stores 32-byte pieces of a thread's own row (STG.256) instead of staging
  long long out_pitch;   // bytes between output rows (>= D * sizeof(TOut), multiple of 16); lets lin_V write packed records
  // MODE 3 (split precision) / single Linear: 1 = `w1` is stored [in, out] (the kernel computes x W instead of x W^T:
  // the input gradient of a Linear, dx = dy W); excludes LayerNorm 0
  int w_transposed;
  int l2_prefetch_ahead; // tiles ahead the producers prefetch into L2 (0 = off; default 2, ALLSET_MLP2_L2_PREFETCH overrides)
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands, fp32 accumulate; issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 TMEM lanes (one per thread of the warp) x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// the reverse: each thread writes 32 consecutive fp32 columns of its TMEM lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in
// bits [0,14), leading byte offset (unused for swizzled K-major, canonical value 1) in [16,30), stride byte
// offset = 1024 B between 8-row atoms in [32,46), version 1 (sm_100) in [46,48), layout type 2 = SWIZZLE_128B
// in [61,64).  Advancing by UMMA_K = 16 bf16 inside the 128-byte swizzle row = +32 B on the start address.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D fp32, A/B bf16, both K-major
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// bounded wait: a wrong descriptor must not hang the box -- after ~2 s flag the status word and carry on.
// SLEEP_NS > 0: back off between polls (producers run far ahead; their polling must not steal issue slots from
// the epilogue warps on the same scheduler).
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity, int* status) {
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    if (SLEEP_NS > 0) {
      // try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or the hint expires
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar), "r"(parity), "r"((uint32_t)SLEEP_NS)
          : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar), "r"(parity)
          : "memory");
    }
    if (ok) return;
    if ((spins & 1023u) == 1023u) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) {
        if (status != nullptr) atomicExch(status, 1);
        return;
      }
    }
  }
}

// byte offset of the 16-byte chunk holding columns [8*j, 8*j+8) of row r in a [rows x K] bf16 operand tile
// stored as K/64 blocks of [rows x 64] in the K-major SWIZZLE_128B layout
template <int ROWS>
__device__ __forceinline__ uint32_t sw128_chunk(int r, int j) {
  return (uint32_t)((j >> 3) * (ROWS * 128) + (r >> 3) * 1024 + (r & 7) * 128 + (((j & 7) ^ (r & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared8(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// 32 bytes per thread in one request (LDG.256, sm_100): a full sector even when the lanes of a warp are rows apart
__device__ __forceinline__ void ld_nc_32(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void st_global32(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                            uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
               "r"(f), "r"(g), "r"(h)
               : "memory");
}
__device__ __forceinline__ uint4 ld_shared16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// packed fp32 pairs (sm_100 FADD2 / FFMA2 / FMUL2: two IEEE fp32 results per instruction)
__device__ __forceinline__ void fadd2(float& d0, float& d1, float b0, float b1) {            // d += b
  asm("{\n\t.reg .b64 pd, pb;\n\tmov.b64 pd, {%0, %1};\n\tmov.b64 pb, {%2, %3};\n\t"
      "add.rn.f32x2 pd, pd, pb;\n\tmov.b64 {%0, %1}, pd;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void ffma2p(float& d0, float& d1, float a0, float a1, float b0, float b1) {   // d += a * b
  asm("{\n\t.reg .b64 pa, pb, pc;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %5};\n\tmov.b64 pc, {%0, %1};\n\t"
      "fma.rn.f32x2 pc, pa, pb, pc;\n\tmov.b64 {%0, %1}, pc;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

template <typename T>
struct RowChunk;   // one 16-byte chunk of a feature row as floats
template <>
struct RowChunk<float> {
  static constexpr int N = 4;
  __device__ static void unpack(const uint4& q, float (&v)[4]) {
    v[0] = __uint_as_float(q.x); v[1] = __uint_as_float(q.y); v[2] = __uint_as_float(q.z); v[3] = __uint_as_float(q.w);
  }
};
template <>
struct RowChunk<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void unpack(const uint4& q, float (&v)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
};

template <int D>
struct Layout {
  static constexpr int KB = D / 64;                 // 128-byte swizzle blocks along K
  static constexpr int A_BYTES = kTileM * D * 2;    // bf16 activation tile
  static constexpr int W_BYTES = D * D * 2;         // one bf16 weight matrix
  static constexpr int TMEM_COLS = (2 * D <= 32) ? 32 : (2 * D <= 64) ? 64 : (2 * D <= 128) ? 128 : (2 * D <= 256) ? 256 : 512;
  template <typename TOut>
  __host__ __device__ static constexpr int pass_bytes() { return (D * (int)sizeof(TOut) < 256) ? D * (int)sizeof(TOut) : 256; }
  template <typename TOut>
  __host__ __device__ static constexpr int buf_bytes() {
    return (A_BYTES > kTileM * pass_bytes<TOut>()) ? A_BYTES : kTileM * pass_bytes<TOut>();
  }
  template <typename TOut>
  __host__ __device__ static constexpr int smem_bytes() {   // + 1024 slack to align the base, + parameters, barriers, TMEM pointer
    return 1024 + 2 * W_BYTES + buf_bytes<TOut>() + 4 * D * 4 + 64;
  }
};

// one GEMM of the tile: acc[128 x D] = A[128 x D] * W[D x D]^T
template <int D>
__device__ __forceinline__ void issue_gemm(uint32_t a_base, uint32_t w_base, uint32_t tmem_acc, uint32_t bar) {
  constexpr uint32_t idesc = instr_desc_bf16(kTileM, D);
#pragma unroll
  for (int k = 0; k < D / 16; ++k) {
    const uint32_t koff = (uint32_t)((k >> 2) * (kTileM * 128) + (k & 3) * 32);
    const uint32_t woff = (uint32_t)((k >> 2) * (D * 128) + (k & 3) * 32);
    umma_bf16(tmem_acc, smem_desc_sw128(a_base + koff), smem_desc_sw128(w_base + woff), idesc, k > 0 ? 1u : 0u);
  }
  umma_commit(bar);
}

// =================================================================================================================
// Warp-specialised pipeline, ONE persistent CTA of 16 warps (512 threads x 128 registers = the whole file) per SM
//   warps 0-7  : two epilogue groups of 4 warps (TMEM lane quadrant = warp & 3).  Group g owns the local tiles
//                g, g+2, ... and a private set of buffers: A stage g, hidden tile g, accumulators acc1[g] / acc2[g].
//                Per tile: wait acc1_full -> epi 1 (+b1, ReLU, LayerNorm, bf16) -> hidden tile -> group barrier ->
//                thread 0 of the group issues GEMM 2 of this tile AND GEMM 1 of the group's next tile -> wait
//                acc2_full -> epi 2 (+b2, ReLU, convert) -> staged in the hidden buffer -> coalesced stores.
//                While one group waits for its GEMM 2 the other one is inside its epilogue.
//   warps 8-15 : producers: rolling register prefetch of the next rows, LayerNorm 0, bf16 A tile (stage = tile & 1)
// mbarriers per buffer set: a_full (8 producer warps -> issuing thread), a_empty (tcgen05.commit -> producers),
// acc1_full / acc2_full (tcgen05.commit -> epilogue group).  TMEM: 4*D columns (512 at D = 128).
// =================================================================================================================
constexpr int kEpiGroups = 2;
constexpr int kEpiWarps = 4 * kEpiGroups;
constexpr int kProdWarps = 8;
constexpr int kWsThreads = (kEpiWarps + kProdWarps) * 32;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void epi_bar_sync(int group) {
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

template <int D>
struct WsLayout {
  using L = Layout<D>;
  static constexpr int TMEM_COLS = (4 * D <= 256) ? 256 : 512;
  template <typename TOut>
  __host__ __device__ static constexpr int smem_bytes() {
    return 1024 + 2 * L::W_BYTES + 2 * L::A_BYTES + 2 * L::template buf_bytes<TOut>() + 6 * D * 4 + 256 + 4 * kTileM * 8;
  }
};

// Producer geometry: 8 lanes per row, 4 rows per warp instruction.  Lane cl of a row owns the 16-byte chunks
// cl, cl+8, cl+16, ... (every load instruction still covers 128 contiguous bytes of each of its 4 rows), so a row
// statistic costs 3 shuffle steps for 4 rows instead of 5 for one, and the per-row scalars are shared by 4 rows.
// A warp owns 16 rows of the tile = 4 row groups; a "half" = 2 groups (8 rows) = one prefetch buffer.
template <typename TIn, int D>
struct Producer {
  static constexpr int EPC = RowChunk<TIn>::N;          // elements per 16-byte chunk
  static constexpr int CHUNKS = D / EPC;                // chunks per row: 32 / 16 / 16 / 8
  static constexpr int LPR = 8;
  static constexpr int CPL = CHUNKS / LPR;              // chunks per lane: 4 / 2 / 2 / 1
  static constexpr int EPL = CPL * EPC;                 // elements per lane and row
  static constexpr int RPI = 32 / LPR;
  static constexpr int ROWS_PER_WARP = kTileM / kProdWarps;
  static constexpr int G = ROWS_PER_WARP / RPI / 2;     // row groups per half
  static_assert(CPL >= 1 && G == 2, "producer geometry");
  // Row of the tile handled by (producer warp pw, half, row group gi, sub-row of the instruction): four consecutive rows
  // per warp instruction.  (Tried: rows {0, 1, 4, 5} + 2 gi, which halves the bank conflicts of the 8-byte stores of fp32
  // rows -- rows r and r ^ 4 land in opposite 64-byte halves of the 128-byte swizzle; measured neutral, +-2 %, on one box.)
  __device__ static __forceinline__ int row_of(int pw, int half, int gi, int sub) {
    return pw * ROWS_PER_WARP + (half * G + gi) * RPI + sub;
  }

  using Buf = uint4[G][CPL];

  __device__ static __forceinline__ void load(Buf& buf, const unsigned char* xb, long long row0, long long rows,
                                              int pw, int half, int sub, int cl) {
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const int r = row_of(pw, half, gi, sub);
      const long long gr = row0 + r;
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        buf[gi][j] = (gr < rows) ? ld_nc_16(xb + (size_t)gr * (D * sizeof(TIn)) + (cl + LPR * j) * 16)
                                 : make_uint4(0, 0, 0, 0);
    }
  }
  // L2 prefetch of a whole tile (128 dense rows = one contiguous block) by the 256 producer threads: the register
  // prefetch reaches one tile ahead (32-64 KB in flight per SM, short of what HBM latency asks for), this one two tiles
  // ahead at no register cost, so that the register loads of the next iteration hit L2.
  __device__ static __forceinline__ void prefetch_l2(const unsigned char* xb, long long row0, long long rows, int ptid) {
    if (row0 >= rows) return;
    const long long nrows = (rows - row0) < kTileM ? (rows - row0) : kTileM;
    const int lines = (int)((nrows * (long long)(D * sizeof(TIn)) + 127) >> 7);
    const unsigned char* base = xb + (size_t)row0 * (D * sizeof(TIn));
    for (int i = ptid; i < lines; i += kProdWarps * 32)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ((size_t)i << 7)));
  }
  // LayerNorm 0 WITHOUT its affine part (gamma / beta are folded into W1 / b1 at setup), the G row groups in lockstep.
  // PMA score of the rows of one half, fp32 on the CUDA cores: lane cl of a row holds EPL of its D elements; 8 heads at
  // a time are reduced over the 8 lanes of the row by a TRANSPOSED butterfly (4 + 2 + 1 shuffles: each step a lane keeps
  // half of its partial sums and adds the partner's), after which lane cl owns head hb + cl and stores it.
  __device__ static __forceinline__ void score(const Buf& buf, const float* sWe, const float* sBe, int H, float* out,
                                               long long row0, long long rows, int pw, int half, int sub, int cl) {
    float x[G][EPL];
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        float t[EPC];
        RowChunk<TIn>::unpack(buf[gi][j], t);
#pragma unroll
        for (int e = 0; e < EPC; ++e) x[gi][j * EPC + e] = t[e];
      }
    }
    for (int hb = 0; hb < H; hb += 8) {
      float a[G][8];
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        float a0[G], a1[G];
#pragma unroll
        for (int gi = 0; gi < G; ++gi) a0[gi] = a1[gi] = 0.f;
        if (hb + h < H) {
          const float* w = sWe + (hb + h) * D;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            float wc[EPC];                                    // one shared-memory read of w_eff serves both row groups
#pragma unroll
            for (int e = 0; e < EPC; ++e) wc[e] = w[(cl + LPR * j) * EPC + e];
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
#pragma unroll
              for (int e = 0; e < EPC; e += 2)
                ffma2p(a0[gi], a1[gi], x[gi][j * EPC + e], x[gi][j * EPC + e + 1], wc[e], wc[e + 1]);
            }
          }
        }
#pragma unroll
        for (int gi = 0; gi < G; ++gi) a[gi][h] = a0[gi] + a1[gi];
      }
      // transposed butterfly over the 8 lanes of the row, both row groups in lockstep
      float b4[G][4], b2[G][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
          const float send = (cl & 4) ? a[gi][i] : a[gi][i + 4];
          const float keep = (cl & 4) ? a[gi][i + 4] : a[gi][i];
          b4[gi][i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
          const float send = (cl & 2) ? b4[gi][i] : b4[gi][i + 2];
          const float keep = (cl & 2) ? b4[gi][i + 2] : b4[gi][i];
          b2[gi][i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
      }
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        const float send = (cl & 1) ? b2[gi][0] : b2[gi][1];
        const float keep = (cl & 1) ? b2[gi][1] : b2[gi][0];
        const float total = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        const long long gr = row0 + row_of(pw, half, gi, sub);
        if (gr < rows && hb + cl < H) out[gr * H + hb + cl] = total + sBe[hb + cl];
      }
    }
  }

  using Packed = uint32_t[G][EPL / 2];          // the normalised rows of one half as bf16 pairs
  // the rows of one half as floats, LayerNorm 0 (without its affine part) applied
  __device__ static __forceinline__ void rows_f32(const Buf& buf, float (&v)[G][EPL], float2 (&stats)[G], bool has_ln0,
                                                  float eps0) {
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        float t[EPC];
        RowChunk<TIn>::unpack(buf[gi][j], t);
#pragma unroll
        for (int e = 0; e < EPC; ++e) v[gi][j * EPC + e] = t[e];
      }
    }
    if (has_ln0) {
      float s[G];
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        float a0 = v[gi][0], a1 = v[gi][1];
#pragma unroll
        for (int e = 2; e < EPL; e += 2) fadd2(a0, a1, v[gi][e], v[gi][e + 1]);
        s[gi] = a0 + a1;
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi) s[gi] += __shfl_xor_sync(0xffffffffu, s[gi], o);
      }
      float mean_[G];
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        const float nm = -s[gi] * (1.f / D);
        mean_[gi] = -nm;
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; e += 2) {
          fadd2(v[gi][e], v[gi][e + 1], nm, nm);
          ffma2p(q0, q1, v[gi][e], v[gi][e + 1], v[gi][e], v[gi][e + 1]);
        }
        s[gi] = q0 + q1;
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi) s[gi] += __shfl_xor_sync(0xffffffffu, s[gi], o);
      }
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        const float rstd = rsqrtf(s[gi] * (1.f / D) + eps0);
#pragma unroll
        for (int e = 0; e < EPL; e += 2) fmul2(v[gi][e], v[gi][e + 1], rstd);
        stats[gi] = make_float2(mean_[gi], rstd);   // PMA tail: the epilogue re-derives the residual from x with these
      }
    }
  }
  __device__ static __forceinline__ void compute(const Buf& buf, Packed& pk, float2 (&stats)[G], bool has_ln0,
                                                 float eps0) {
    float v[G][EPL];
    rows_f32(buf, v, stats, has_ln0, eps0);
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
#pragma unroll
      for (int e = 0; e < EPL; e += 2) pk[gi][e / 2] = pack_bf16(v[gi][e], v[gi][e + 1]);
    }
  }
  // split precision: the three bf16 terms (x = t0 + t1 + t2 to 2^-25) of the columns of this lane that fall into K block
  // KB_ (64 columns) of its rows, each term stored as soon as it is formed (v is consumed: the residual is left in it)
  template <int KBLK, int KB_>
  __device__ static __forceinline__ void split_store_block(float (&v)[G][EPL], uint32_t s0, uint32_t s1, uint32_t s2,
                                                           int pw, int half, int sub, int cl) {
    constexpr int JB = CPL / KBLK;                       // chunks of this lane per K block
    static_assert(CPL % KBLK == 0, "split_store_block: chunks per lane must split evenly over the K blocks");
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const uint32_t base = t == 0 ? s0 : (t == 1 ? s1 : s2);
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        const int r = row_of(pw, half, gi, sub);
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          const int j = KB_ * JB + jj;
          uint32_t w[EPC / 2];
#pragma unroll
          for (int e = 0; e < EPC; e += 2) {
            float& a = v[gi][j * EPC + e];
            float& b = v[gi][j * EPC + e + 1];
            const uint32_t h = pack_bf16(a, b);
            w[e / 2] = h;
            if (t < 2) {
              a -= __uint_as_float(h << 16);
              b -= __uint_as_float(h & 0xFFFF0000u);
            }
          }
          const int col = (cl + LPR * j) * EPC;
          const uint32_t dst = base + sw128_chunk<kTileM>(r, col >> 3) + (uint32_t)((col & 7) * 2);
          if constexpr (EPC == 4) {
            st_shared8(dst, w[0], w[1]);
          } else {
            st_shared16(dst, w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
  }
  // the only part that needs the A stage to be free: EPL/2 registers per row -> swizzled shared memory
  __device__ static __forceinline__ void store(const Packed& pk, const float2 (&stats)[G], float2* stat, uint32_t sAst,
                                               int pw, int half, int sub, int cl) {
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const int r = row_of(pw, half, gi, sub);
      if (stat != nullptr && cl == 0) stat[r] = stats[gi];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int col = (cl + LPR * j) * EPC;                  // first column of this chunk
        const uint32_t dst = sAst + sw128_chunk<kTileM>(r, col >> 3) + (uint32_t)((col & 7) * 2);
        const uint32_t* w = &pk[gi][j * EPC / 2];
        if constexpr (EPC == 4) {
          st_shared8(dst, w[0], w[1]);
        } else {
          st_shared16(dst, w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
};

template <typename TIn, typename TOut, int D, int MODE>
__global__ void __launch_bounds__(kWsThreads, 1) mlp2_ws_kernel(const Params p) {
  constexpr bool TAIL = (MODE == 1);
  constexpr bool SCORE = (MODE == 2);
  // MODE 3: ONE Linear with fp32 accuracy out of bf16 tensor-core products.  Both operands are split into THREE bf16 terms
  // (x = x1 + x2 + x3 exactly to 2^-25: 3 x 8 significant bits) and the six products down to 2^-18 relative are
  // accumulated in fp32 in TMEM, smallest first:  x3 W1 + x1 W3 + x2 W2 + x2 W1 + x1 W2 + x1 W1  (the dropped terms are
  // <= 2^-26).  Two terms / three products (the usual "3x" split) leave 2^-17 per product: measured on the citeseer
  // golden model that is enough for the logits (1.4e-5) but flips ~10 ReLUs that sit within 1e-5 of zero, and a flipped
  // unit is an O(1) change of that row's gradient -- so the reference-precision mode pays for the third term.
  // Shared memory: the three weight terms take the W1 / W2 slots and the second hidden buffer, the three A terms the two
  // A stages and the first hidden buffer, i.e. ONE tile's worth of A that both epilogue groups' tiles pass through in
  // turn, refilled per 64-column K block while the tensor core works on the other block; the output leaves through the
  // direct (thread-per-row) stores because no staging buffer is left.
  constexpr bool SPLIT = (MODE == 3);
  using L = Layout<D>;
  static_assert(D == 64 || D == 128, "mlp2_ws: widths 64 and 128");
  constexpr int PASS_BYTES = L::template pass_bytes<TOut>();
  constexpr int OUT_ROW_BYTES = D * (int)sizeof(TOut);
  constexpr int NPASS = OUT_ROW_BYTES / PASS_BYTES;
  constexpr int CPP = PASS_BYTES / (int)sizeof(TOut);
  constexpr int CHUNKS_PER_ROW = PASS_BYTES / 16;
  constexpr int BUF = L::template buf_bytes<TOut>();

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t sW1 = raw_addr + pad;
  const uint32_t sW2 = sW1 + L::W_BYTES;
  const uint32_t sA0 = sW2 + L::W_BYTES;                 // 2 stages of L::A_BYTES
  const uint32_t sA1 = sA0 + 2 * L::A_BYTES;             // 2 hidden tiles / output staging buffers of BUF bytes
  constexpr int PAR_OFF = 2 * L::W_BYTES + 2 * L::A_BYTES + 2 * BUF;
  float* sPar = reinterpret_cast<float*>(smem + PAR_OFF);     // b1', b2' (folded) | tail: gamma0, beta0, gamma_f, beta_f
  const uint32_t sBar = sW1 + PAR_OFF + 6 * D * 4;            // 8 mbarriers, [kind][buffer set]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + PAR_OFF + 6 * D * 4 + 128);
  float2* sStat = reinterpret_cast<float2*>(smem + PAR_OFF + 6 * D * 4 + 256);   // tail: (mean0, rstd0) [group][k & 1][row]
  const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 16, bar_acc1_full = sBar + 32, bar_acc2_full = sBar + 48;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool has_ln0 = p.ln0_g != nullptr, has_ln1 = p.ln1_g != nullptr;

  // ---- setup ------------------------------------------------------------------------------------------------
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_a_full + 8 * b, kProdWarps);
      mbar_init(bar_a_empty + 8 * b, 1);
      mbar_init(bar_acc1_full + 8 * b, 1);
      mbar_init(bar_acc2_full + 8 * b, SPLIT ? kProdWarps : 1);   // SPLIT: "K block 1 filled" (see gemm1)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), WsLayout<D>::TMEM_COLS);
  // weights -> bf16 K-major SWIZZLE_128B, with the LayerNorm in front of each Linear folded in:
  //   LN(x) W^T + b = n(x) (W diag(gamma))^T + (b + W beta),  n(x) = (x - mean) * rstd
  constexpr int NW = SPLIT ? 3 : 2;                     // weight tiles to fill (SPLIT: the three terms of w1)
#pragma unroll 2
  for (int idx = tid; idx < NW * D * (D / 8); idx += kWsThreads) {
    const int which = idx / (D * (D / 8));
    const int rem = idx - which * (D * (D / 8));
    const bool first = SPLIT || !which;
    int n, j;
    if (p.w_transposed) { n = rem % D; j = rem / D; }   // consecutive threads read consecutive output columns
    else { n = rem / (D / 8); j = rem % (D / 8); }
    const float* wsrc = first ? p.w1 : p.w2;
    float wv[8];
    if (p.w_transposed) {
#pragma unroll
      for (int e = 0; e < 8; ++e) wv[e] = wsrc[(size_t)(j * 8 + e) * D + n];
    } else {
      const float4 lo = *reinterpret_cast<const float4*>(wsrc + (size_t)n * D + j * 8);
      const float4 hi = *reinterpret_cast<const float4*>(wsrc + (size_t)n * D + j * 8 + 4);
      wv[0] = lo.x; wv[1] = lo.y; wv[2] = lo.z; wv[3] = lo.w; wv[4] = hi.x; wv[5] = hi.y; wv[6] = hi.z; wv[7] = hi.w;
    }
    const float* gam = first ? p.ln0_g : p.ln1_g;
    if (gam != nullptr) {
#pragma unroll
      for (int e = 0; e < 8; ++e) wv[e] *= gam[j * 8 + e];
    }
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      uint32_t h = pack_bf16(wv[e], wv[e + 1]);
      if (SPLIT) {
        for (int t = 0; t < which; ++t) {               // peel `which` leading terms
          wv[e] -= __uint_as_float(h << 16);
          wv[e + 1] -= __uint_as_float(h & 0xFFFF0000u);
          h = pack_bf16(wv[e], wv[e + 1]);
        }
      }
      pk[e / 2] = h;
    }
    const uint32_t wdst = (which == 0) ? sW1 : (which == 1) ? sW2 : sA1 + BUF;
    st_shared16(wdst + sw128_chunk<D>(n, j), pk[0], pk[1], pk[2], pk[3]);
  }
  // biases: sPar[0..D) = b1 + W1 beta0, sPar[D..2D) = b2 + W2 beta1 (fp32; 4 threads per output row, independent
  // 16-byte loads so the whole fold is one round trip to L2)
  for (int n = tid >> 2; n < 2 * D; n += kWsThreads / 4) {
    const int which = n / D, row = n - which * D, part = tid & 3;
    const float* bias = which ? p.b2 : p.b1;
    const float* beta = which ? (has_ln1 ? p.ln1_b : nullptr) : (has_ln0 ? p.ln0_b : nullptr);
    float acc = 0.f;
    if (beta != nullptr && !(which && p.single)) {
      const float* w = (which ? p.w2 : p.w1) + (size_t)row * D + part * (D / 4);
      const float* bt = beta + part * (D / 4);
#pragma unroll
      for (int kk = 0; kk < D / 4; kk += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(w + kk);
        acc = fmaf(wv.x, bt[kk], fmaf(wv.y, bt[kk + 1], fmaf(wv.z, bt[kk + 2], fmaf(wv.w, bt[kk + 3], acc))));
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0) sPar[n] = (bias ? bias[row] : 0.f) + acc;
  }
  if (SCORE) {        // w_eff [H, D] fp32 lives where the tail's row statistics would (H * D * 4 <= 4 KB)
    float* sWe = reinterpret_cast<float*>(sStat);
    for (int i = tid; i < p.heads * D; i += kWsThreads) sWe[i] = p.w_eff[i];
    for (int i = tid; i < p.heads; i += kWsThreads) sPar[2 * D + i] = p.b_eff ? p.b_eff[i] : 0.f;
  }
  if (TAIL) {
    for (int i = tid; i < D; i += kWsThreads) {
      sPar[2 * D + i] = p.ln0_g[i];
      sPar[3 * D + i] = p.ln0_b ? p.ln0_b[i] : 0.f;
      sPar[4 * D + i] = p.lnf_g ? p.lnf_g[i] : 1.f;
      sPar[5 * D + i] = p.lnf_b ? p.lnf_b[i] : 0.f;
    }
  }
  proxy_fence_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long n_tiles = (p.rows + kTileM - 1) / kTileM;

  if (warp >= kEpiWarps) {
    // ======================= producers =====================================================================
    using P = Producer<TIn, D>;
    const int pw = warp - kEpiWarps;
    const int sub = lane / P::LPR, cl = lane % P::LPR;
    const unsigned char* xb = static_cast<const unsigned char*>(p.x);
    typename P::Buf bufA, bufB;
    long long tile = blockIdx.x;
    if (tile < n_tiles) {
      P::load(bufA, xb, tile * kTileM, p.rows, pw, 0, sub, cl);
      P::load(bufB, xb, tile * kTileM, p.rows, pw, 1, sub, cl);
    }
    for (uint32_t it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t st = it & 1u;
      const long long next = tile + gridDim.x;
      if (p.l2_prefetch_ahead > 0)
        P::prefetch_l2(xb, (tile + (long long)p.l2_prefetch_ahead * gridDim.x) * kTileM, p.rows, pw * 32 + lane);
      const uint32_t sAst = sA0 + st * L::A_BYTES;
      float2* stat = TAIL ? sStat + (st * 2 + ((it >> 1) & 1u)) * kTileM : nullptr;
      // all the arithmetic happens BEFORE the stage is known to be free; only the shared-memory stores wait for it
      typename P::Packed pkA, pkB;
      float2 stA[P::G] = {}, stB[P::G] = {};
      if constexpr (SPLIT) {
        // Three term tiles, ONE tile's worth of shared memory, pipelined per K BLOCK (64 columns): block kb of all three
        // tiles is refilled as soon as the MMAs of the previous tile that read it have retired (a_empty[kb]), while the
        // tensor core still works on the other block.  a_empty is indexed by K block here (one phase per tile, waited on
        // by the producers only); "filled" is signalled per (block, owning group): a_full[g] / acc2_full[g].
        constexpr int KBLK = D / 64;
        const uint32_t t0 = sA0, t1 = sA0 + L::A_BYTES, t2 = sA1;
        float vA[P::G][P::EPL], vB[P::G][P::EPL];
        float2 stx[P::G];
        if (it >= 1) mbar_wait_bounded<200>(bar_a_empty, (it - 1) & 1u, p.status);
        P::rows_f32(bufA, vA, stx, has_ln0, p.eps0);
        P::template split_store_block<KBLK, 0>(vA, t0, t1, t2, pw, 0, sub, cl);
        P::rows_f32(bufB, vB, stx, has_ln0, p.eps0);
        P::template split_store_block<KBLK, 0>(vB, t0, t1, t2, pw, 1, sub, cl);
        proxy_fence_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a_full + 8 * st);   // st = it & 1 = the group that owns this tile
        if (next < n_tiles) {                              // prefetch AFTER the first block: fewer live registers before it
          P::load(bufA, xb, next * kTileM, p.rows, pw, 0, sub, cl);
          P::load(bufB, xb, next * kTileM, p.rows, pw, 1, sub, cl);
        }
        if constexpr (KBLK == 2) {
          if (it >= 1) mbar_wait_bounded<200>(bar_a_empty + 8, (it - 1) & 1u, p.status);
          P::template split_store_block<KBLK, KBLK - 1>(vA, t0, t1, t2, pw, 0, sub, cl);
          P::template split_store_block<KBLK, KBLK - 1>(vB, t0, t1, t2, pw, 1, sub, cl);
          proxy_fence_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc2_full + 8 * st);
        }
        continue;
      }
      if (SCORE) P::score(bufA, reinterpret_cast<const float*>(sStat), sPar + 2 * D, p.heads, p.score, tile * kTileM,
                          p.rows, pw, 0, sub, cl);
      P::compute(bufA, pkA, stA, has_ln0, p.eps0);
      if (next < n_tiles) P::load(bufA, xb, next * kTileM, p.rows, pw, 0, sub, cl);
      if (SCORE) P::score(bufB, reinterpret_cast<const float*>(sStat), sPar + 2 * D, p.heads, p.score, tile * kTileM,
                          p.rows, pw, 1, sub, cl);
      P::compute(bufB, pkB, stB, has_ln0, p.eps0);
      if (next < n_tiles) P::load(bufB, xb, next * kTileM, p.rows, pw, 1, sub, cl);
      if (it >= 2) mbar_wait_bounded<2000>(bar_a_empty + 8 * st, ((it >> 1) - 1) & 1u, p.status);
      P::store(pkA, stA, stat, sAst, pw, 0, sub, cl);     // (this sStat slot was read at the top of tile k-2's epilogue)
      P::store(pkB, stB, stat, sAst, pw, 1, sub, cl);
      proxy_fence_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a_full + 8 * st);
    }
  } else {
    // ======================= epilogue groups: thread = row = TMEM lane ==========================================
    const int g = warp >> 2;                           // group = buffer index
    const int r = tid & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tmem_acc1 = tmem_base + g * D, tmem_acc2 = tmem_base + 2 * D + g * D;
    const uint32_t sH = sA1 + g * BUF;
    unsigned char* ob = static_cast<unsigned char*>(p.out);
    const bool issuer = (r == 0);                      // one thread per group issues its tcgen05.mma / commit
    const bool direct = SPLIT || p.direct != 0;        // SPLIT: the staging buffers hold operand terms
    const uint32_t sAg = sA0 + g * L::A_BYTES;
    auto gemm1 = [&](uint32_t kk) {                    // GEMM 1 of this group's kk-th tile: A stage g x W1 -> acc1[g]
      if constexpr (SPLIT) {
        // K block by K block (see the producers); per block the six products, smallest first:
        // x3 W1, x1 W3, x2 W2, x2 W1, x1 W2, x1 W1, all into the same fp32 accumulator
        constexpr int KBLK = D / 64;
        const uint32_t xa[3] = {sA0, sA0 + (uint32_t)L::A_BYTES, sA1};
        const uint32_t wa[3] = {sW1, sW2, sA1 + (uint32_t)BUF};
        constexpr int XI[6] = {2, 0, 1, 1, 0, 0}, WI[6] = {0, 2, 1, 0, 1, 0};
#pragma unroll
        for (int kb = 0; kb < KBLK; ++kb) {
          // "block kb of this group's kk-th tile is filled": one barrier per (block, group) -- the two issuers alternate
          // tiles and may be early or late relative to each other, which a parity wait on a shared barrier cannot tell
          // apart.  Block 1 borrows the acc2_full slots (unused by a single Linear).
          mbar_wait_bounded((kb == 0 ? bar_a_full : bar_acc2_full) + 8 * g, kk & 1u, p.status);
          tc_fence_after();
#pragma unroll
          for (int t = 0; t < 6; ++t) {
#pragma unroll
            for (int ks = 4 * kb; ks < 4 * kb + 4; ++ks) {
              const uint32_t koff = (uint32_t)((ks >> 2) * (kTileM * 128) + (ks & 3) * 32);
              const uint32_t woff = (uint32_t)((ks >> 2) * (D * 128) + (ks & 3) * 32);
              umma_bf16(tmem_acc1, smem_desc_sw128(xa[XI[t]] + koff), smem_desc_sw128(wa[WI[t]] + woff),
                        instr_desc_bf16(kTileM, D), (kb > 0 || t > 0 || ks > 4 * kb) ? 1u : 0u);
            }
          }
          umma_commit(bar_a_empty + 8 * kb);
        }
        umma_commit(bar_acc1_full + 8 * g);
        return;
      }
      mbar_wait_bounded(bar_a_full + 8 * g, kk & 1u, p.status);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        const uint32_t koff = (uint32_t)((ks >> 2) * (kTileM * 128) + (ks & 3) * 32);
        const uint32_t woff = (uint32_t)((ks >> 2) * (D * 128) + (ks & 3) * 32);
        umma_bf16(tmem_acc1, smem_desc_sw128(sAg + koff), smem_desc_sw128(sW1 + woff), instr_desc_bf16(kTileM, D),
                  ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_a_empty + 8 * g);
      umma_commit(bar_acc1_full + 8 * g);
    };
    long long tile = blockIdx.x + (long long)g * gridDim.x;
    if (issuer && tile < n_tiles) gemm1(0);
    __syncwarp();
    if (TAIL) epi_bar_sync(g);                         // every thread of the group is ordered after the issuer's a_full acquire
    uint32_t k = 0;                                    // this group's tile counter
    for (; tile < n_tiles; tile += 2LL * gridDim.x, ++k) {
      const long long row0 = tile * kTileM;
      // ---- epilogue 1 ----------------------------------------------------------------------------------------
      mbar_wait_bounded(bar_acc1_full + 8 * g, k & 1u, p.status);
      tc_fence_after();
      const bool has_next = tile + 2LL * gridDim.x < n_tiles;
      // PMA tail: this row's (mean0, rstd0).  Read NOW: GEMM 1 of the group's next tile is issued after epilogue 1, its
      // commit frees the A stage, and from then on the producers may write the statistics of tile k+2 into this slot.
      float2 st0 = make_float2(0.f, 1.f);
      if (TAIL) st0 = sStat[(g * 2 + (k & 1u)) * kTileM + r];
      if (!p.single) {
      const uint32_t t1 = tmem_acc1 + lane_off;
      float mean = 0.f, rstd = 1.f;
      if (has_ln1) {
        float sa[4] = {0.f, 0.f, 0.f, 0.f}, qa[4] = {0.f, 0.f, 0.f, 0.f};      // 2 independent packed chains each
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
          float v[32];
          tmem_ld32(t1 + c, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            fadd2(v[i], v[i + 1], sPar[c + i], sPar[c + i + 1]);
            v[i] = fmaxf(v[i], 0.f);
            v[i + 1] = fmaxf(v[i + 1], 0.f);
            fadd2(sa[i & 2], sa[(i & 2) + 1], v[i], v[i + 1]);
            ffma2p(qa[i & 2], qa[(i & 2) + 1], v[i], v[i + 1], v[i], v[i + 1]);
          }
        }
        const float ssum = (sa[0] + sa[1]) + (sa[2] + sa[3]);
        const float qsum = (qa[0] + qa[1]) + (qa[2] + qa[3]);
        mean = ssum * (1.f / D);
        rstd = rsqrtf(fmaxf(qsum * (1.f / D) - mean * mean, 0.f) + p.eps1);
      }
      const float nmr = -mean * rstd;              // LayerNorm 1 without its affine part (folded into W2 / b2)
#pragma unroll 1
      for (int c = 0; c < D; c += 32) {
        float v[32];
        tmem_ld32(t1 + c, v);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          fadd2(v[i], v[i + 1], sPar[c + i], sPar[c + i + 1]);
          float h0 = fmaxf(v[i], 0.f), h1 = fmaxf(v[i + 1], 0.f);
          if (has_ln1) {
            v[i] = nmr;
            v[i + 1] = nmr;
            ffma2(v[i], v[i + 1], rstd, h0, h1);
          } else {
            v[i] = h0;
            v[i + 1] = h1;
          }
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          st_shared16(sH + sw128_chunk<kTileM>(r, (c >> 3) + q4), pack_bf16(v[8 * q4], v[8 * q4 + 1]),
                      pack_bf16(v[8 * q4 + 2], v[8 * q4 + 3]), pack_bf16(v[8 * q4 + 4], v[8 * q4 + 5]),
                      pack_bf16(v[8 * q4 + 6], v[8 * q4 + 7]));
      }
      proxy_fence_async();
      tc_fence_before();
      epi_bar_sync(g);                           // hidden tile complete, acc1[g] drained by all 128 threads
      if (issuer) {
        tc_fence_after();
        issue_gemm<D>(sH, sW2, tmem_acc2, bar_acc2_full + 8 * g);       // acc2[g] was drained before the last barrier
        if (has_next) gemm1(k + 1);                                     // runs under epilogue 2 of this tile
      }
      __syncwarp();
      // ---- epilogue 2 ----------------------------------------------------------------------------------------
      mbar_wait_bounded(bar_acc2_full + 8 * g, k & 1u, p.status);
      tc_fence_after();
      }  // !single
      const uint32_t tfin = p.single ? tmem_acc1 : tmem_acc2;           // single Linear: acc1 is the result
      const int bias_off = p.single ? 0 : D;
      float rstdf = 1.f, nmrf = 0.f;
      if constexpr (TAIL) {
        // ---- PMA tail, pass 1: z = LN0(x) + relu(acc2 + b2') -> back into acc2[g] (TMEM), row statistics of z ---------
        // (mean0, rstd0) were written by the producers before their a_full arrive; the issuer acquired that barrier
        // before GEMM 1 of this tile and has been through a group barrier with every thread here since (waiting on
        // a_full again would be wrong: the producers may already have completed the NEXT phase of that barrier).
        const float nmr0 = -st0.x * st0.y;
        const long long gr = row0 + r;
        const unsigned char* xrow = static_cast<const unsigned char*>(p.x) + (size_t)(gr < p.rows ? gr : 0) * (D * sizeof(TIn));
        float sa[4] = {0.f, 0.f, 0.f, 0.f}, qa[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
          float v[32], xf[32];
          // a thread re-reads ITS row: 32-byte loads (LDG.256), so every request is one whole sector although the 32 rows
          // of a warp instruction are D * sizeof(TIn) bytes apart
          constexpr int XC = RowChunk<TIn>::N;
#pragma unroll
          for (int j = 0; j < 32 / (2 * XC); ++j) {
            uint4 qa = make_uint4(0, 0, 0, 0), qb = make_uint4(0, 0, 0, 0);
            if (gr < p.rows) ld_nc_32(xrow + (size_t)(c + j * 2 * XC) * sizeof(TIn), qa, qb);
            float t[XC];
            RowChunk<TIn>::unpack(qa, t);
#pragma unroll
            for (int e = 0; e < XC; ++e) xf[j * 2 * XC + e] = t[e];
            RowChunk<TIn>::unpack(qb, t);
#pragma unroll
            for (int e = 0; e < XC; ++e) xf[j * 2 * XC + XC + e] = t[e];
          }
          tmem_ld32(tmem_acc2 + lane_off + c, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            fadd2(v[i], v[i + 1], sPar[D + c + i], sPar[D + c + i + 1]);
            v[i] = fmaxf(v[i], 0.f);
            v[i + 1] = fmaxf(v[i + 1], 0.f);
            float t0 = nmr0, t1 = nmr0;
            ffma2(t0, t1, st0.y, xf[i], xf[i + 1]);                                  // (x - mean0) * rstd0
            float y0 = sPar[3 * D + c + i], y1 = sPar[3 * D + c + i + 1];
            ffma2p(y0, y1, t0, t1, sPar[2 * D + c + i], sPar[2 * D + c + i + 1]);    // * gamma0 + beta0
            fadd2(v[i], v[i + 1], y0, y1);                                           // z = y + relu(h)
            fadd2(sa[i & 2], sa[(i & 2) + 1], v[i], v[i + 1]);
            ffma2p(qa[i & 2], qa[(i & 2) + 1], v[i], v[i + 1], v[i], v[i + 1]);
          }
          tmem_st32(tmem_acc2 + lane_off + c, v);
        }
        tmem_wait_st();
        const float mf = ((sa[0] + sa[1]) + (sa[2] + sa[3])) * (1.f / D);
        rstdf = rsqrtf(fmaxf(((qa[0] + qa[1]) + (qa[2] + qa[3])) * (1.f / D) - mf * mf, 0.f) + p.epsf);
        nmrf = -mf * rstdf;
      }
#pragma unroll 1
      for (int pass = 0; pass < NPASS; ++pass) {
        const uint32_t t2 = tfin + lane_off + pass * CPP;
        const uint32_t rowb = sH + (uint32_t)r * PASS_BYTES;
#pragma unroll 1
        for (int c = 0; c < CPP; c += 32) {
          float v[32];
          tmem_ld32(t2 + c, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int cc = pass * CPP + c + i;
            if constexpr (TAIL) {                  // LNf(z): (z - mean) * rstd * gamma_f + beta_f, then the outer ReLU
              float t0 = nmrf, t1 = nmrf;
              ffma2(t0, t1, rstdf, v[i], v[i + 1]);
              v[i] = sPar[5 * D + cc];
              v[i + 1] = sPar[5 * D + cc + 1];
              ffma2p(v[i], v[i + 1], t0, t1, sPar[4 * D + cc], sPar[4 * D + cc + 1]);
              if (p.relu_final) {
                v[i] = fmaxf(v[i], 0.f);
                v[i + 1] = fmaxf(v[i + 1], 0.f);
              }
            } else {
              fadd2(v[i], v[i + 1], sPar[bias_off + cc], sPar[bias_off + cc + 1]);
              if (p.relu_out) {
                v[i] = fmaxf(v[i], 0.f);
                v[i + 1] = fmaxf(v[i + 1], 0.f);
              }
            }
          }
          if (direct) {
            // thread-per-row stores in whole 32-byte sectors: no staging round trip through shared memory
            const long long gr = row0 + r;
            if (gr < p.rows) {
              unsigned char* dst = ob + (size_t)gr * (size_t)p.out_pitch + (size_t)(pass * CPP + c) * sizeof(TOut);
              if constexpr (sizeof(TOut) == 2) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
                  st_global32(dst + 32 * q, pack_bf16(v[16 * q], v[16 * q + 1]), pack_bf16(v[16 * q + 2], v[16 * q + 3]),
                              pack_bf16(v[16 * q + 4], v[16 * q + 5]), pack_bf16(v[16 * q + 6], v[16 * q + 7]),
                              pack_bf16(v[16 * q + 8], v[16 * q + 9]), pack_bf16(v[16 * q + 10], v[16 * q + 11]),
                              pack_bf16(v[16 * q + 12], v[16 * q + 13]), pack_bf16(v[16 * q + 14], v[16 * q + 15]));
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  st_global32(dst + 32 * q, __float_as_uint(v[8 * q]), __float_as_uint(v[8 * q + 1]),
                              __float_as_uint(v[8 * q + 2]), __float_as_uint(v[8 * q + 3]), __float_as_uint(v[8 * q + 4]),
                              __float_as_uint(v[8 * q + 5]), __float_as_uint(v[8 * q + 6]), __float_as_uint(v[8 * q + 7]));
              }
            }
          } else if constexpr (sizeof(TOut) == 2) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int c16 = (c >> 3) + q4;
              st_shared16(rowb + (uint32_t)((c16 ^ (r & 7)) << 4), pack_bf16(v[8 * q4], v[8 * q4 + 1]),
                          pack_bf16(v[8 * q4 + 2], v[8 * q4 + 3]), pack_bf16(v[8 * q4 + 4], v[8 * q4 + 5]),
                          pack_bf16(v[8 * q4 + 6], v[8 * q4 + 7]));
            }
          } else {
#pragma unroll
            for (int q8 = 0; q8 < 8; ++q8) {
              const int c16 = (c >> 2) + q8;
              st_shared16(rowb + (uint32_t)((c16 ^ (r & 7)) << 4), __float_as_uint(v[4 * q8]),
                          __float_as_uint(v[4 * q8 + 1]), __float_as_uint(v[4 * q8 + 2]), __float_as_uint(v[4 * q8 + 3]));
            }
          }
        }
        tc_fence_before();                       // (last pass: the accumulator is drained before the barriers below)
        // A warp stages and stores only ITS OWN 32 rows (its TMEM lane quadrant), so __syncwarp orders the staging
        // writes before the coalesced reads; group barriers remain only where other warps' data is involved.
        if (p.single && pass == NPASS - 1) {     // acc1[g] must be drained by all 128 threads before GEMM 1 of the next tile
          epi_bar_sync(g);
          if (issuer && has_next) gemm1(k + 1);
        }
        __syncwarp();
        if (direct) continue;                  // nothing staged: the hidden tile is untouched, no barrier needed
        const int wrow = (warp & 3) * 32;
#pragma unroll
        for (int idx = lane; idx < 32 * CHUNKS_PER_ROW; idx += 32) {
          const int rr = wrow + idx / CHUNKS_PER_ROW, c16 = idx % CHUNKS_PER_ROW;
          const uint4 qv = ld_shared16(sH + (uint32_t)rr * PASS_BYTES + (uint32_t)((c16 ^ (rr & 7)) << 4));
          const long long gr = row0 + rr;
          if (gr < p.rows)
            *reinterpret_cast<uint4*>(ob + (size_t)gr * (size_t)p.out_pitch + pass * PASS_BYTES + c16 * 16) = qv;
        }
        if (pass == NPASS - 1) epi_bar_sync(g);  // the staging buffer is the hidden tile: every warp done before epilogue 1
        else __syncwarp();                       // of the next tile overwrites it; between passes only this warp's rows
      }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, WsLayout<D>::TMEM_COLS);
}

template <typename TIn, typename TOut, int D, int MODE>
int launch(const Params& p_in, cudaStream_t st) {
  static const int ahead = getenv("ALLSET_MLP2_L2_PREFETCH") != nullptr ? atoi(getenv("ALLSET_MLP2_L2_PREFETCH")) : 2;
  Params p = p_in;
  p.l2_prefetch_ahead = ahead;
  const long long n_tiles = (p.rows + kTileM - 1) / kTileM;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  constexpr int smem = WsLayout<D>::template smem_bytes<TOut>();
  cudaError_t e = cudaFuncSetAttribute(mlp2_ws_kernel<TIn, TOut, D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return fail(ALLSET_ECUDA, "mlp2_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  int fits = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fits, mlp2_ws_kernel<TIn, TOut, D, MODE>, kWsThreads, smem);
  if (e != cudaSuccess || fits < 1)
    return fail(ALLSET_ECUDA, "mlp2_fwd: the 16-warp kernel does not fit one SM (%s)",
                e != cudaSuccess ? cudaGetErrorString(e) : "0 resident CTAs");
  long long grid = sms;
  if (grid > n_tiles) grid = n_tiles;
  mlp2_ws_kernel<TIn, TOut, D, MODE><<<(unsigned)grid, kWsThreads, smem, st>>>(p);
  return check_launch("mlp2_fwd");
}

}  // namespace mlp5
