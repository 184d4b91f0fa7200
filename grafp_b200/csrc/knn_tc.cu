// k-NN graph, tensor-core path (sm_100a): TMA-staged tiles -> tcgen05.mma (kind::tf32, 3xTF32
// split for fp32-grade distances) -> TMEM accumulators -> fused per-row top-K epilogue.
//
// One CTA owns 128 query rows of one segment (UMMA_M = 128) and streams all key tiles of that
// segment past them.  The N x M distance matrix exists only as 128 x BN fp32 tiles in tensor
// memory; HBM sees the normalised features (hi/lo split, written by knn_normalize_kernel) and
// the k neighbour ids per row.
//
// Warp roles (192 threads):  warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (warp w reads TMEM lanes 32*(w%4)..+31, one accumulator row per thread,
// so the running top-K list of a row lives in one thread and needs no cross-thread traffic).
//
// Pipelines: smem ring full/empty mbarriers (TMA <-> MMA), double-buffered TMEM accumulator
// full/empty mbarriers (MMA <-> epilogue): selection of key tile t overlaps the MMAs of tile t+1.
#include <mutex>

#include "tc_ptx.cuh"

namespace grafp {
namespace tc {
using namespace tcptx;

constexpr int BM = 128;        // query rows per CTA = UMMA_M
constexpr int BK = 32;         // fp32 per k-chunk = one 128-byte swizzle-atom row
constexpr int UMMA_K = 8;      // kind::tf32
constexpr int kThreads = 192;
constexpr int kEpilogueThreads = 128;
constexpr int kMaxListK = 64;  // this first-generation kernel keeps whole K-entry lists in shared memory

template <int BN, int KREG>
struct Cfg {
  static constexpr int kStages = (BN == 256) ? 2 : (KREG == 0 ? 2 : 3);
  static constexpr uint32_t kABytes = BM * BK * 4;  // one of hi / lo
  static constexpr uint32_t kBBytes = BN * BK * 4;
  static constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BN;  // two accumulator buffers
  static constexpr uint32_t kYsqBytes = 2 * BN * 4;
  static constexpr uint32_t kListBytes = (KREG == 0) ? kMaxListK * kEpilogueThreads * 8 : 0;
  static constexpr uint32_t kBarBytes = 128;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kYsqBytes + kListBytes + kBarBytes + 1024;
};

// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = BN
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// per-row top-K lists
// ---------------------------------------------------------------------------------------------
template <int KREG>
struct RegList {
  float d[KREG];
  int id[KREG];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int p = 0; p < KREG; ++p) { d[p] = INFINITY; id[p] = 0; }
  }
  // sorted ascending; a candidate equal to an entry goes behind it (keys arrive in ascending id order)
  __device__ __forceinline__ void offer(float v, int key) {
    if (v < d[KREG - 1]) {
#pragma unroll
      for (int p = KREG - 1; p > 0; --p) {
        if (v < d[p - 1]) { d[p] = d[p - 1]; id[p] = id[p - 1]; }
        else if (v < d[p]) { d[p] = v; id[p] = key; }
      }
      if (v < d[0]) { d[0] = v; id[0] = key; }
    }
  }
};

struct SmemList {  // column-per-thread layout [rank][thread]: conflict-free
  float* d;
  int* id;
  int K;
  float worst;
  __device__ __forceinline__ void init(float* dbase, int* ibase, int r, int K_) {
    d = dbase + r; id = ibase + r; K = K_;
    for (int p = 0; p < K; ++p) { d[p * kEpilogueThreads] = INFINITY; id[p * kEpilogueThreads] = 0; }
    worst = INFINITY;
  }
  __device__ __forceinline__ void offer(float v, int key) {
    if (v < worst) {
      int p = K - 1;
      while (p > 0 && d[(p - 1) * kEpilogueThreads] > v) {
        d[p * kEpilogueThreads] = d[(p - 1) * kEpilogueThreads];
        id[p * kEpilogueThreads] = id[(p - 1) * kEpilogueThreads];
        --p;
      }
      d[p * kEpilogueThreads] = v;
      id[p * kEpilogueThreads] = key;
      worst = d[(K - 1) * kEpilogueThreads];
    }
  }
};

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int KREG>
__global__ void __launch_bounds__(kThreads, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
              const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
              const float* __restrict__ xsq, const float* __restrict__ ysq, long long* __restrict__ nn_idx,
              int* __restrict__ nn_idx32, int N, int M, int C, int K, int k_out, int stride) {
  using cfg = Cfg<BN, KREG>;
  extern __shared__ unsigned char smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_base = smem;
  float* ysq_s = reinterpret_cast<float*>(smem + cfg::kStages * cfg::kStageBytes);  // [2][BN]
  unsigned char* list_base = reinterpret_cast<unsigned char*>(ysq_s) + cfg::kYsqBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(list_base + cfg::kListBytes);
  // bars: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then the TMEM base address
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * cfg::kStages;
  const uint32_t bar_tfull = bar_empty + 8 * cfg::kStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int m0 = blockIdx.x * BM;
  const int num_tiles = (M + BN - 1) / BN;
  const int num_kc = (C + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < cfg::kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
    fence_barrier_init();
  }
  if (warp == 1) {  // one warp allocates tensor memory and later frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int t = 0; t < num_tiles; ++t) {
        for (int c = 0; c < num_kc; ++c, ++it) {
          const int s = it % cfg::kStages;
          const uint32_t ph = (it / cfg::kStages) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t full = bar_full + 8 * s;
          mbar_arrive_expect_tx(full, cfg::kStageBytes);
          const uint32_t sa = smem_u32(stage_base + s * cfg::kStageBytes);
          tma_load_3d(sa, &tm_a_hi, full, c * BK, m0, b);
          tma_load_3d(sa + cfg::kABytes, &tm_a_lo, full, c * BK, m0, b);
          tma_load_3d(sa + 2 * cfg::kABytes, &tm_b_hi, full, c * BK, t * BN, b);
          tma_load_3d(sa + 2 * cfg::kABytes + cfg::kBBytes, &tm_b_lo, full, c * BK, t * BN, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      int it = 0;
      for (int t = 0; t < num_tiles; ++t) {
        const int as = t & 1;
        const uint32_t aph = (t >> 1) & 1;
        mbar_wait(bar_tempty + 8 * as, aph ^ 1);  // epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        const uint32_t acc = tmem_base + as * BN;
        for (int c = 0; c < num_kc; ++c, ++it) {
          const int s = it % cfg::kStages;
          const uint32_t ph = (it / cfg::kStages) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(stage_base + s * cfg::kStageBytes);
          const uint32_t a_hi = sa, a_lo = sa + cfg::kABytes;
          const uint32_t b_hi = sa + 2 * cfg::kABytes, b_lo = b_hi + cfg::kBBytes;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint32_t off = kk * UMMA_K * 4;
            const uint64_t d_ahi = make_smem_desc(a_hi + off), d_alo = make_smem_desc(a_lo + off);
            const uint64_t d_bhi = make_smem_desc(b_hi + off), d_blo = make_smem_desc(b_lo + off);
            // 3xTF32: x.y ~= lo.hi + hi.lo + hi.hi (lo.lo is below fp32 resolution)
            umma_tf32(acc, d_alo, d_bhi, idesc, (c | kk) != 0 ? 1u : 0u);
            umma_tf32(acc, d_ahi, d_blo, idesc, 1u);
            umma_tf32(acc, d_ahi, d_bhi, idesc, 1u);
          }
          tcgen05_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs retire
        }
        tcgen05_commit(bar_tfull + 8 * as);   // accumulator tile complete
      }
    }
  } else {
    // ===== epilogue: distances + running per-row top-K =====
    const int sub = warp & 3;                 // TMEM lane group this warp may read
    const int r = sub * 32 + lane;            // accumulator row == thread
    const int et = (warp - 2) * 32 + lane;    // 0..127 index among the epilogue threads
    const int q = m0 + r;
    const float sq_i = (q < N) ? xsq[(long long)b * N + q] : 0.f;
    const float* ysq_b = ysq + (long long)b * M;

    RegList<(KREG > 0 ? KREG : 1)> rl;
    SmemList sl;
    if constexpr (KREG > 0) rl.init();
    else sl.init(reinterpret_cast<float*>(list_base), reinterpret_cast<int*>(list_base + kMaxListK * kEpilogueThreads * 4), r, K);

    for (int t = 0; t < num_tiles; ++t) {
      const int as = t & 1;
      const uint32_t aph = (t >> 1) & 1;
      float* ys = ysq_s + as * BN;
      for (int i = et; i < BN; i += kEpilogueThreads) {
        const int key = t * BN + i;
        ys[i] = (key < M) ? ysq_b[key] : INFINITY;  // keys past the end can never be selected
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpilogueThreads) : "memory");
      mbar_wait(bar_tfull + 8 * as, aph);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + as * BN;
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(trow + cc * 32, v);
        tmem_ld_wait();
        const int key0 = t * BN + cc * 32;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 y4 = *reinterpret_cast<const float4*>(ys + cc * 32 + j4 * 4);
          const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // D = (|x|^2 + (-2 s)) + |y|^2, in the reference's association order (torch_edge.py:16-18)
            const float dist = __fadd_rn(fmaf(-2.f, __uint_as_float(v[j4 * 4 + e]), sq_i), yv[e]);
            if constexpr (KREG > 0) rl.offer(dist, key0 + j4 * 4 + e);
            else sl.offer(dist, key0 + j4 * 4 + e);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
    }

    if (q < N) {
      const long long o = ((long long)b * N + q) * k_out;
      if constexpr (KREG > 0) {
#pragma unroll
        for (int p = 0; p < KREG; ++p) {
          if (p % stride == 0 && p / stride < k_out) {
            nn_idx[o + p / stride] = rl.id[p];
            if (nn_idx32 != nullptr) nn_idx32[o + p / stride] = rl.id[p];
          }
        }
      } else {
        for (int j = 0; j < k_out; ++j) {
          const int id = sl.id[j * stride * kEpilogueThreads];
          nn_idx[o + j] = id;
          if (nn_idx32 != nullptr) nn_idx32[o + j] = id;
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cfg::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static EncodeTiledFn encode_fn() { return tcptx::encode_tiled_fn(); }

// rows x C fp32 matrix per batch item, box = box_rows x 32 elements, 128B swizzle, zero fill out of bounds
static bool make_map(CUtensorMap* map, const void* base, int B, int rows, int C, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)rows * C * 4};
  const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int KREG>
static int launch_variant(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi,
                          const CUtensorMap& tb_lo, const float* xsq, const float* ysq, long long* nn_idx, int* nn_idx32,
                          int B, int N, int M, int C, int K, int k_out, int stride, cudaStream_t s) {
  using cfg = Cfg<BN, KREG>;
  static DeviceOnce once;
  if (once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(knn_tc_kernel<BN, KREG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(knn_tc): %s", cudaGetErrorString(e)); return (int)e; }
    once.mark();
  }
  dim3 grid((N + BM - 1) / BM, B);
  knn_tc_kernel<BN, KREG><<<grid, kThreads, cfg::kSmemBytes, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, xsq, ysq, nn_idx, nn_idx32,
                                                                 N, M, C, K, k_out, stride);
  return check_launch("knn_tc");
}

}  // namespace tc

namespace tcptx {
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}
}  // namespace tcptx

bool knn_tc_supported(int N, int M, int C, int K, int dtype) {
  // (the tf32 planes live in fp32 containers whatever the input dtype: the normalise kernel reads fp32 or bf16 rows)
  return (dtype == GRAFP_F32 || dtype == GRAFP_BF16) && C % 4 == 0 && C >= 32 && N >= 128 && M >= 128 && K >= 1 && K <= tc::kMaxListK;
}

int launch_knn_tc(const void* xhi, const void* xlo, const float* xsq, const void* yhi, const void* ylo,
                  const float* ysq, const float* relpos, long long* nn_idx, int* nn_idx32, int B, int N, int M, int C,
                  int K, int k_out, int stride, int dtype, cudaStream_t s) {
  using namespace tc;
  if (relpos != nullptr || !knn_tc_supported(N, M, C, K, dtype)) {
    set_error("knn_tc: unsupported configuration");
    return GRAFP_EUNSUPPORTED;
  }
  const int bn = (M > 128) ? 256 : 128;
  const bool use_list = K > 8;
  const int box_b = use_list ? 128 : bn;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  if (!make_map(&ta_hi, xhi, B, N, C, BM) || !make_map(&ta_lo, xlo, B, N, C, BM) ||
      !make_map(&tb_hi, yhi, B, M, C, box_b) || !make_map(&tb_lo, ylo, B, M, C, box_b)) {
    set_error("knn_tc: cuTensorMapEncodeTiled failed (driver entry point unavailable or bad shape)");
    return GRAFP_EUNSUPPORTED;
  }
#define GRAFP_TC_LAUNCH(BN_, KREG_) \
  launch_variant<BN_, KREG_>(ta_hi, ta_lo, tb_hi, tb_lo, xsq, ysq, nn_idx, nn_idx32, B, N, M, C, K, k_out, stride, s)
  if (use_list) return GRAFP_TC_LAUNCH(128, 0);
  if (bn == 256) return (K <= 3) ? GRAFP_TC_LAUNCH(256, 3) : GRAFP_TC_LAUNCH(256, 8);
  return (K <= 3) ? GRAFP_TC_LAUNCH(128, 3) : GRAFP_TC_LAUNCH(128, 8);
#undef GRAFP_TC_LAUNCH
}

}  // namespace grafp
