// 1x1 convolution over node rows as a tensor-core GEMM with the BatchNorm statistics in its epilogue (sm_100a).
//
//   Y[R x Cout] = X[R x Cin] . W[Cout x Cin]^T          (reference: Conv2d(Cin, Cout, 1) of torch_nn.py:52-64,
//   sums[c]        = sum_r Y[r][c]                        torch_vertex.py:152-162,183-194, graph_encoder.py:45-67 over
//   sums[Cout + c] = sum_r Y[r][c]^2                      (B, C, N, 1) activations, rows = (b, n), K-major operands)
//
// Every such convolution of the encoder is followed by a train-mode BatchNorm whose first pass re-reads Y only to
// take its per-channel mean and variance.  Here the accumulator tile is already in registers on its way out, so
// the moments are taken there and the BatchNorm is left with its apply pass: one read of Y instead of two.
//
// Grouped convolutions (BasicConv's groups = 4) keep the dense 128 x BN output tile and make the MMA schedule
// block-diagonal: a group's columns of the tile are one MMA of N = Cout / groups over that group's k-range.
//
// fp32 rows run as TF32 (tcgen05.mma.kind::tf32, fp32 accumulate - the arithmetic cuDNN uses for the same layer when
// torch.backends.cudnn.allow_tf32 is set, which is PyTorch's default; the host takes this path only then), bf16 rows as
// kind::f16 with bf16 operands.  The layer is HBM-bound at the first two stages of the encoder; beyond, a 128 x 256
// tile is bound by the operand traffic from L2 (the host keeps k-ranges above 2 KB per row on cuDNN: ops.py).
//
// Persistent CTAs (one per SM), tile = 128 rows x BN output channels, column tile fastest so the CTAs that share an X
// row block run at the same time and it is read from HBM once.  Warp roles (320 threads): warp 0 = TMA producer
// (X and W k-chunks of 128 bytes per row, SWIZZLE_128B, 3-6 stage ring), warp 1 = TMEM allocator + MMA issuer (one
// elected thread), warps 2-9 = epilogue.  Two accumulator buffers in tensor memory: the epilogue of tile t overlaps
// the loads and MMAs of tile t+1.
//
// Epilogue (two teams of four warps, alternating 64-column units, one staging buffer each): tcgen05.ld (one accumulator
// row per thread) -> round to the output type -> swizzled staging tile in shared memory -> TMA store (coalesced,
// asynchronous, clipped at the matrix edge); while the store drains, thread (c, h) sums column c over row half h of the
// staging tile.  Moments are taken of the values as stored (the bf16 roundings included), about the first row of the
// half (no cancellation when |mean| >> std), widened to double and shifted back to raw moments there; a thread keeps
// its double accumulators across tiles of the same column block and issues its red.global.add.f64 only when the
// block changes.  (With the weights (Cout, Cin / groups) the first line reads W[Cout x Cin/groups] per group.)
#include <cuda_bf16.h>

#include "tc_ptx.cuh"

namespace grafp {
namespace cg {
using namespace tcptx;

constexpr int BM = 128;
constexpr int kThreads = 320;      // producer warp, MMA warp, two epilogue teams of four warps
constexpr int kTeamThreads = 128;
constexpr uint32_t kBoxBytes = BM * 128;  // one staging / operand box: 128 rows of 128 bytes

template <typename T, int BN>
struct Cfg {
  static constexpr int BK = 128 / sizeof(T);        // elements per k-chunk row (one swizzle line)
  static constexpr int UMMA_K = 32 / sizeof(T);     // 8 (tf32) / 16 (bf16)
  static constexpr uint32_t kABytes = BM * 128;
  static constexpr uint32_t kBBytes = BN * 128;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 256 ? 3 : (BN == 128 ? 4 : 6);  // 144 / 128 / 144 KB of operands in flight
  static constexpr int kUnits = BN / 64;             // the epilogue stages 64 accumulator columns at a time
  static constexpr int kBoxCols = 128 / sizeof(T);   // output columns per staging box row
  static constexpr int kUnitBoxes = 64 / kBoxCols;   // 2 (fp32) / 1 (bf16) boxes of 128 rows x 128 bytes
  static constexpr uint32_t kUnitBytes = kUnitBoxes * kBoxBytes;
  static constexpr uint32_t kStagingBytes = 2 * kUnitBytes;  // one buffer per epilogue team
  static constexpr uint32_t kTmemCols = 2 * BN;
  static constexpr uint32_t kBarBytes = 256;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;
};

// instruction descriptor: D = f32, A = B = tf32 (2) or bf16 (1), both K-major, M = 128, N = n (a multiple of 16, <= BN)
template <typename T, int BN>
__device__ __forceinline__ uint32_t make_idesc(int n) {
  constexpr uint32_t fmt = sizeof(T) == 4 ? 2u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc_of(uint32_t lo) {
  return (static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | lo;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void team_barrier(int team) { asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(kTeamThreads) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_bf16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return __uint_as_float(static_cast<uint32_t>(v) << 16);
}
__device__ __forceinline__ uint32_t pack_bf16(uint32_t lo_bits, uint32_t hi_bits) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(__uint_as_float(lo_bits), __uint_as_float(hi_bits));
  return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

template <typename T, int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv1x1_stats_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                     const __grid_constant__ CUtensorMap tm_y, double* __restrict__ sums, long long R, int Cin, int Cout,
                     int Cg, int Og, int wN) {
  using cfg = Cfg<T, BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t ring = smem_u32(smem);
  const uint32_t staging = ring + cfg::kStages * cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + cfg::kStages * cfg::kStageBytes + cfg::kStagingBytes);
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * cfg::kStages;
  const uint32_t bar_tfull = bar_empty + 8 * cfg::kStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int col_tiles = (Cout + BN - 1) / BN;
  const long long row_tiles = (R + BM - 1) / BM;
  const long long num_tiles = row_tiles * col_tiles;
  // Grouped convolution (Cg = Cin / groups input, Og = Cout / groups output channels per group; dense: Cg = Cin,
  // Og = Cout): the output tile stays 128 x BN, the MMA schedule becomes block-diagonal - the tile's columns are filled
  // in sub-blocks of wN = min(BN, Og) columns, each from the k-range [g Cg, (g + 1) Cg) of its own group g.
  const int num_kc = (Cg + cfg::BK - 1) / cfg::BK;
  const int sub_blocks = BN / wN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < cfg::kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    // an accumulator buffer is handed back by every epilogue warp that reads it: both teams, or (BN = 64) one
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, cfg::kUnits > 1 ? 8 : 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
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
    if (elect_one()) {
      int it = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int r0 = static_cast<int>(tile / col_tiles) * BM;
        const int n0 = static_cast<int>(tile % col_tiles) * BN;
        for (int sb = 0; sb < sub_blocks; ++sb) {
          const int nc0 = n0 + sb * wN;
          if (nc0 >= Cout) break;
          const int kbase = (nc0 / Og) * Cg;
          for (int c = 0; c < num_kc; ++c, ++it) {
            const int s = it % cfg::kStages;
            const uint32_t ph = (it / cfg::kStages) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            const uint32_t full = bar_full + 8 * s;
            mbar_arrive_expect_tx(full, cfg::kABytes + static_cast<uint32_t>(wN) * 128u);
            const uint32_t sa = ring + s * cfg::kStageBytes;
            // (a chunk may run past the group's k-range into the next group's columns of x: the matching columns of the
            //  W box lie beyond its Cg-wide rows and arrive as zeros)
            tma_load_2d(sa, &tm_x, full, kbase + c * cfg::BK, r0);
            tma_load_2d(sa + cfg::kABytes, &tm_w, full, c * cfg::BK, nc0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = make_idesc<T, BN>(wN);
      int it = 0, lt = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int as = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        const int n0 = static_cast<int>(tile % col_tiles) * BN;
        mbar_wait(bar_tempty + 8 * as, aph ^ 1);  // the epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        for (int sb = 0; sb < sub_blocks; ++sb) {
          if (n0 + sb * wN >= Cout) break;
          const uint32_t acc = tmem_base + as * BN + sb * wN;
          for (int c = 0; c < num_kc; ++c, ++it) {
            const int s = it % cfg::kStages;
            const uint32_t ph = (it / cfg::kStages) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            tcgen05_fence_after();
            const uint32_t sa = ring + s * cfg::kStageBytes;
            const uint32_t a_lo0 = desc_lo(sa), b_lo0 = desc_lo(sa + cfg::kABytes);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {  // four 32-byte k-steps per 128-byte chunk row
              if constexpr (sizeof(T) == 4) umma_tf32(acc, desc_of(a_lo0 + kk * 2), desc_of(b_lo0 + kk * 2), idesc, (c | kk) != 0 ? 1u : 0u);
              else umma_f16(acc, desc_of(a_lo0 + kk * 2), desc_of(b_lo0 + kk * 2), idesc, (c | kk) != 0 ? 1u : 0u);
            }
            tcgen05_commit(bar_empty + 8 * s);  // frees the stage when these MMAs retire
          }
        }
        tcgen05_commit(bar_tfull + 8 * as);
      }
    }
  } else {
    // ===== epilogue: two teams of four warps (one warp per TMEM lane group each) =====
    // A tile leaves in units of 64 columns; team t takes the units u = t (mod 2) (BN = 64: the tiles lt = t (mod 2)),
    // each through its own staging buffer, so the dependent chain of one unit (tcgen05.ld -> st.shared -> barrier ->
    // TMA store -> column sums) overlaps the other team's.
    const int team = (warp - 2) >> 2;
    const int sub = warp & 3;               // TMEM lane group of this warp
    const int r = sub * 32 + lane;          // accumulator row of this thread
    const int et = ((warp - 2) & 3) * 32 + lane;
    const int scol = et & 63;               // column of the staged unit this thread sums ...
    const int shalf = et >> 6;              // ... over rows [64 shalf, 64 shalf + 64)
    const uint32_t swz = static_cast<uint32_t>(r & 7);
    const uint32_t buf = staging + team * cfg::kUnitBytes;
    constexpr int kOwn = cfg::kUnits > 1 ? cfg::kUnits / 2 : 1;
    double a1[kOwn], a2[kOwn];
#pragma unroll
    for (int uu = 0; uu < kOwn; ++uu) { a1[uu] = 0.0; a2[uu] = 0.0; }
    int cur_ct = -1;
    auto unit_of = [&](int uu) { return cfg::kUnits > 1 ? 2 * uu + team : 0; };
    auto flush = [&]() {
      if (cur_ct < 0) return;
#pragma unroll
      for (int uu = 0; uu < kOwn; ++uu) {
        const int col = cur_ct * BN + unit_of(uu) * 64 + scol;
        if (col < Cout) {
          red_add_f64(sums + col, a1[uu]);
          red_add_f64(sums + Cout + col, a2[uu]);
        }
        a1[uu] = 0.0; a2[uu] = 0.0;
      }
    };
    uint32_t cbase, cchunk;  // this thread's column inside the staging buffer
    if constexpr (sizeof(T) == 4) { cbase = buf + (scol >> 5) * kBoxBytes + (scol & 3) * 4; cchunk = static_cast<uint32_t>((scol & 31) >> 2); }
    else { cbase = buf + (scol & 7) * 2; cchunk = static_cast<uint32_t>(scol >> 3); }
    auto at = [&](int row) -> float {
      const uint32_t addr = cbase + row * 128 + ((cchunk ^ static_cast<uint32_t>(row & 7)) << 4);
      if constexpr (sizeof(T) == 4) return lds_f32(addr);
      else return lds_bf16(addr);
    };
    int lt = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      if (cfg::kUnits == 1 && (lt & 1) != team) continue;
      const int rt = static_cast<int>(tile / col_tiles);
      const int ct = static_cast<int>(tile % col_tiles);
      if (ct != cur_ct) { flush(); cur_ct = ct; }
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      const long long left = R - (long long)rt * BM;
      const int rows_valid = left < BM ? static_cast<int>(left) : BM;
      const int row_begin = shalf * 64;
      const int nrows = min(64, max(0, rows_valid - row_begin));
      mbar_wait(bar_tfull + 8 * as, aph);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + as * BN;
#pragma unroll
      for (int uu = 0; uu < kOwn; ++uu) {
        const int u = unit_of(uu);
        uint32_t v0[32], v1[32];
        tmem_ld32(trow + u * 64, v0);
        tmem_ld32(trow + u * 64 + 32, v1);
        if (et == 0) bulk_wait_read_0();  // the team's previous store has read the staging buffer ...
        team_barrier(team);               // ... and every thread of the team is done summing it
        tmem_ld_wait();
        if (uu == kOwn - 1) {  // this team is done with the accumulator buffer
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
        }
        if constexpr (sizeof(T) == 4) {
          const uint32_t base = buf + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            sts128(base + ((static_cast<uint32_t>(j) ^ swz) << 4), v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
            sts128(base + kBoxBytes + ((static_cast<uint32_t>(j) ^ swz) << 4), v1[4 * j], v1[4 * j + 1], v1[4 * j + 2], v1[4 * j + 3]);
          }
        } else {
          const uint32_t base = buf + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            sts128(base + ((static_cast<uint32_t>(j) ^ swz) << 4), pack_bf16(v0[8 * j], v0[8 * j + 1]), pack_bf16(v0[8 * j + 2], v0[8 * j + 3]),
                   pack_bf16(v0[8 * j + 4], v0[8 * j + 5]), pack_bf16(v0[8 * j + 6], v0[8 * j + 7]));
            sts128(base + ((static_cast<uint32_t>(4 + j) ^ swz) << 4), pack_bf16(v1[8 * j], v1[8 * j + 1]), pack_bf16(v1[8 * j + 2], v1[8 * j + 3]),
                   pack_bf16(v1[8 * j + 4], v1[8 * j + 5]), pack_bf16(v1[8 * j + 6], v1[8 * j + 7]));
          }
        }
        fence_proxy_async();  // the staging writes become visible to the TMA store
        team_barrier(team);
        if (et == 0) {
#pragma unroll
          for (int bx = 0; bx < cfg::kUnitBoxes; ++bx)
            tma_store_2d(&tm_y, buf + bx * kBoxBytes, ct * BN + u * 64 + bx * cfg::kBoxCols, rt * BM);
          bulk_commit();
        }
        // column moments of the staged unit, about the first row of this thread's row half
        const int col = ct * BN + u * 64 + scol;
        if (col < Cout && nrows > 0) {
          const float sh = at(row_begin);
          float p1a = 0.f, p1b = 0.f, p2a = 0.f, p2b = 0.f;
          if (nrows == 64) {
#pragma unroll 8
            for (int i = 0; i < 64; i += 2) {
              const float d0 = at(row_begin + i) - sh, d1 = at(row_begin + i + 1) - sh;
              p1a += d0; p1b += d1;
              p2a = fmaf(d0, d0, p2a); p2b = fmaf(d1, d1, p2b);
            }
          } else {
            for (int i = 0; i < nrows; ++i) {
              const float d = at(row_begin + i) - sh;
              p1a += d;
              p2a = fmaf(d, d, p2a);
            }
          }
          const double n = (double)nrows, sd = (double)sh, q1 = (double)p1a + (double)p1b;
          a1[uu] += q1 + n * sd;
          a2[uu] += ((double)p2a + (double)p2b) + 2.0 * sd * q1 + n * sd * sd;
        }
      }
    }
    flush();
    if (et == 0) bulk_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cfg::kTmemCols) : "memory");
  }
}

// 2-D row-major matrix [rows][cols] of `esize`-byte elements, box = box_rows x 128 bytes, 128-byte swizzle
bool make_map(CUtensorMap* map, const void* base, long long rows, int cols, int box_rows, int dtype) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return false;
  const int es = dtype == GRAFP_F32 ? 4 : 2;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * es};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, dtype == GRAFP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <typename T, int BN>
int launch_variant(const void* x, const void* w, void* y, double* sums, long long R, int Cin, int Cout, int groups, int dtype,
                   cudaStream_t s) {
  using cfg = Cfg<T, BN>;
  static DeviceOnce once;
  if (once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(conv1x1_stats_kernel<T, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv1x1_stats): %s", cudaGetErrorString(e)); return (int)e; }
    once.mark();
  }
  const int Cg = Cin / groups, Og = Cout / groups;
  const int wN = (groups > 1 && Og < BN) ? Og : BN;   // columns one MMA fills: a whole tile, or one group's share of it
  CUtensorMap tx, tw, ty;
  if (!make_map(&tx, x, R, Cin, BM, dtype) || !make_map(&tw, w, Cout, Cg, wN, dtype) || !make_map(&ty, y, R, Cout, BM, dtype)) {
    set_error("conv1x1_bn_stats: cuTensorMapEncodeTiled failed (driver entry point unavailable or bad shape)");
    return GRAFP_EUNSUPPORTED;
  }
  const long long tiles = ((R + BM - 1) / BM) * ((Cout + BN - 1) / BN);
  const int sms = num_sms();
  const int grid = (int)(tiles < sms ? tiles : sms);
  conv1x1_stats_kernel<T, BN><<<grid, kThreads, cfg::kSmemBytes, s>>>(tx, tw, ty, sums, R, Cin, Cout, Cg, Og, wN);
  return check_launch("conv1x1_bn_stats");
}

}  // namespace cg

// tile width for Cout output channels in `groups` groups, or 0 when the shape cannot be tiled: a tile must not straddle
// a group boundary unless it holds whole groups, and a group's share of a tile is one MMA (N a multiple of 16)
static int tile_width(int Cout, int groups) {
  const int bn = Cout > 128 ? 256 : (Cout > 64 ? 128 : 64);
  if (groups == 1) return bn;
  const int Og = Cout / groups;
  for (int cand : {bn, 128, 64}) {
    if (cand > bn) continue;
    if (Og >= cand ? (Og % cand == 0) : (cand % Og == 0 && Og % 16 == 0)) return cand;
  }
  return 0;
}

bool conv1x1_stats_supported(long long R, int Cin, int Cout, int groups, int dtype) {
  if (dtype != GRAFP_F32 && dtype != GRAFP_BF16) return false;
  const int es = dtype == GRAFP_F32 ? 4 : 2;
  if (!(R >= 1 && R < (1LL << 31) - 256 && Cin >= 1 && Cout >= 1 && groups >= 1)) return false;
  if (Cin % groups != 0 || Cout % groups != 0) return false;
  if ((Cin * es) % 16 != 0 || (Cout * es) % 16 != 0 || ((Cin / groups) * es) % 16 != 0) return false;
  return tile_width(Cout, groups) != 0;
}

int launch_conv1x1_stats(const void* x, const void* w, void* y, double* sums, long long R, int Cin, int Cout, int groups,
                         int dtype, cudaStream_t s) {
  if (!conv1x1_stats_supported(R, Cin, Cout, groups, dtype)) {
    set_error("conv1x1_bn_stats: needs fp32 / bf16 rows, Cin, Cout and Cin / groups multiples of 16 bytes, and groups whose "
              "output share is a multiple of 16 channels that tiles 64 / 128 / 256 columns");
    return GRAFP_EUNSUPPORTED;
  }
  if (!aligned16(x) || !aligned16(w) || !aligned16(y)) {
    set_error("conv1x1_bn_stats: x, w and y must be 16-byte aligned");
    return GRAFP_EINVAL;
  }
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)2 * Cout * sizeof(double), s);
  if (e != cudaSuccess) { set_error("conv1x1_bn_stats: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
  const int bn = tile_width(Cout, groups);
#define GRAFP_CG_LAUNCH(T_, BN_) cg::launch_variant<T_, BN_>(x, w, y, sums, R, Cin, Cout, groups, dtype, s)
  if (dtype == GRAFP_F32) {
    if (bn == 256) return GRAFP_CG_LAUNCH(float, 256);
    if (bn == 128) return GRAFP_CG_LAUNCH(float, 128);
    return GRAFP_CG_LAUNCH(float, 64);
  }
  if (bn == 256) return GRAFP_CG_LAUNCH(__nv_bfloat16, 256);
  if (bn == 128) return GRAFP_CG_LAUNCH(__nv_bfloat16, 128);
  return GRAFP_CG_LAUNCH(__nv_bfloat16, 64);
#undef GRAFP_CG_LAUNCH
}

}  // namespace grafp
