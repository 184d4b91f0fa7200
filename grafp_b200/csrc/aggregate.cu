// Memory-bound graph-aggregation kernels: fused neighbour gather + max-relative
// reduction (K2), its argmax-routed scatter backward (K3), the plain gather, the
// EdgeConv feature builder and the max-over-k reduction, forward and backward.
//
// Layout: node features are rows (B, N, C); one thread handles VEC consecutive
// channels of one node, so a warp reads/writes whole 128-byte row segments and the
// neighbour gathers are row-contiguous 16-byte loads served by L1/L2 (a segment's
// features are 256 KiB, i.e. cache resident while its rows are being gathered).
#include "common.cuh"

namespace grafp {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------
// K2: max-relative aggregation forward (torch_vertex.py:21-32)
// ------------------------------------------------------------------------------------
template <typename T, int VEC, bool I64, bool HAS_CTR>
__global__ void __launch_bounds__(kThreads)
mr_aggregate_fwd_kernel(const T* __restrict__ x, const T* __restrict__ src, const void* __restrict__ nbr,
                        const void* __restrict__ ctr, T* __restrict__ out, uint8_t* __restrict__ argmax,
                        long long rows, int N, int M, int C, int k) {
  const int cv = C / VEC;
  const long long items = rows * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long row = it / cv;
    const int c = static_cast<int>(it - row * cv) * VEC;
    const long long b = row / N;
    float self[VEC];
    Pack<T, VEC>::load(x + row * C + c, self);
    float best[VEC];
    int arg[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { best[e] = -INFINITY; arg[e] = 0; }
    const long long ibase = row * k;
#pragma unroll 4
    for (int j = 0; j < k; ++j) {
      const int nb = load_index<I64>(nbr, ibase + j);
      float xj[VEC];
      Pack<T, VEC>::load(src + (b * M + nb) * (long long)C + c, xj);
      float xc[VEC];
      if constexpr (HAS_CTR) {
        const int ci = load_index<I64>(ctr, ibase + j);
        Pack<T, VEC>::load(x + (b * N + ci) * (long long)C + c, xc);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) xc[e] = self[e];
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float d = xj[e] - xc[e];
        if (d > best[e] || d != d) {  // strict '>' keeps the first maximiser; NaN propagates like torch.max
          if (!(best[e] != best[e])) { best[e] = d; arg[e] = j; }
        }
      }
    }
    // interleave [x_c, m_c] (torch_vertex.py:32)
    T* o = out + row * 2 * C + 2 * c;
    if constexpr (VEC == 4) {
      float lo[4] = {self[0], best[0], self[1], best[1]};
      float hi[4] = {self[2], best[2], self[3], best[3]};
      Pack<T, 4>::store(o, lo);
      Pack<T, 4>::store(o + 4, hi);
      if (argmax != nullptr) {
        *reinterpret_cast<uchar4*>(argmax + row * C + c) =
            make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
      }
    } else {
      float a[1] = {self[0]}, m[1] = {best[0]};
      Pack<T, 1>::store(o, a);
      Pack<T, 1>::store(o + 1, m);
      if (argmax != nullptr) argmax[row * C + c] = (uint8_t)arg[0];
    }
  }
}

// Fast path (4-channel packs, C/4 a power of two <= 256, centre == row): grid.y = segment, so all index
// arithmetic is 32-bit shifts/masks (the generic kernel spends most of its issue slots on 64-bit
// divisions).  Every thread keeps U independent (row, 4-channel) items in flight so the dependent chain
// id -> neighbour row -> store is overlapped U-deep.
template <typename T, bool I64, int U>
__global__ void __launch_bounds__(kThreads)
mr_aggregate_fwd_fast_kernel(const T* __restrict__ x, const T* __restrict__ src, const void* __restrict__ nbr,
                             T* __restrict__ out, uint8_t* __restrict__ argmax, int N, int M, int C, int k,
                             int cv_shift) {
  const int b = blockIdx.y;
  const int rpb = kThreads >> cv_shift;                       // rows per block pass
  const int r_local = threadIdx.x >> cv_shift;
  const int c = (threadIdx.x & ((1 << cv_shift) - 1)) * 4;
  const T* xb = x + (long long)b * N * C + c;
  const T* sb = src + (long long)b * M * C + c;
  T* ob = out + (long long)b * N * 2 * C + 2 * c;
  uint8_t* ab = argmax ? argmax + (long long)b * N * C + c : nullptr;
  const long long ib = (long long)b * N * k;
  for (int n0 = blockIdx.x * rpb * U + r_local; n0 < N; n0 += gridDim.x * rpb * U) {
    int n[U];
    float self[U][4], best[U][4];
    int arg[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      n[u] = min(n0 + u * rpb, N - 1);  // clamp: dead lanes recompute the last row, stores are guarded
      Pack<T, 4>::load(xb + (long long)n[u] * C, self[u]);
#pragma unroll
      for (int e = 0; e < 4; ++e) { best[u][e] = -INFINITY; arg[u][e] = 0; }
    }
    // The compare chain is the issue-slot bottleneck (FSETP/FSEL/SEL all share the ALU pipe), so it is
    // kept to one compare + two selects per element; NaN / inf inputs are detected on the FMA pipe
    // (d * 0 accumulates to NaN) and such rows are redone by the exact slow path below.
    float poison[U];
#pragma unroll
    for (int u = 0; u < U; ++u) poison[u] = 0.f;
    for (int j = 0; j < k; ++j) {
      int nb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) nb[u] = load_index<I64>(nbr, ib + n[u] * k + j);
      float xj[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) Pack<T, 4>::load(sb + (long long)nb[u] * C, xj[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d = xj[u][e] - self[u][e];
          poison[u] = fmaf(d, 0.f, poison[u]);
          const bool gt = d > best[u][e];  // strict: the first maximiser wins
          best[u][e] = gt ? d : best[u][e];
          arg[u][e] = gt ? j : arg[u][e];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (poison[u] != 0.f) {  // NaN or inf seen: redo with torch.max semantics (NaN propagates, first NaN wins)
#pragma unroll
        for (int e = 0; e < 4; ++e) { best[u][e] = -INFINITY; arg[u][e] = 0; }
        for (int j = 0; j < k; ++j) {
          const int nbj = load_index<I64>(nbr, ib + n[u] * k + j);
          float xj[4];
          Pack<T, 4>::load(sb + (long long)nbj * C, xj);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float d = xj[e] - self[u][e];
            if (d > best[u][e] || d != d) {
              if (!(best[u][e] != best[u][e])) { best[u][e] = d; arg[u][e] = j; }
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (n0 + u * rpb >= N) continue;
      T* o = ob + (long long)n[u] * 2 * C;
      const float il[8] = {self[u][0], best[u][0], self[u][1], best[u][1], self[u][2], best[u][2], self[u][3], best[u][3]};
      Pack8<T>::store(o, il);  // one full-sector store per thread
      if (ab != nullptr) {
        const unsigned int packed = (unsigned)arg[u][0] | ((unsigned)arg[u][1] << 8) | ((unsigned)arg[u][2] << 16) |
                                    ((unsigned)arg[u][3] << 24);
        *reinterpret_cast<unsigned int*>(ab + (long long)n[u] * C) = packed;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// K3 (generic form): dense pass + atomic scatter pass, ordered by the stream.
//   dense:   grad_x[row][c] = g[row][2c] (- g[row][2c+1] when the centre is the row itself)
//            (+ g[row][2c+1] again when the winning neighbour is the row itself, i.e. the
//             centre and neighbour contributions cancel exactly and no atomic is needed)
//   scatter: grad_src[nbr[row][argmax]][c] += g[row][2c+1];  grad_x[ctr[row][argmax]][c] -= g[row][2c+1]
// ------------------------------------------------------------------------------------
template <typename T, int VEC>
__device__ __forceinline__ void load_grad_pair(const T* g, float (&g0)[VEC], float (&g1)[VEC]) {
  if constexpr (VEC == 4) {
    float a[4], b[4];
    Pack<T, 4>::load(g, a);
    Pack<T, 4>::load(g + 4, b);
    g0[0] = a[0]; g1[0] = a[1]; g0[1] = a[2]; g1[1] = a[3];
    g0[2] = b[0]; g1[2] = b[1]; g0[3] = b[2]; g1[3] = b[3];
  } else {
    float a[1], b[1];
    Pack<T, 1>::load(g, a);
    Pack<T, 1>::load(g + 1, b);
    g0[0] = a[0]; g1[0] = b[0];
  }
}

template <int VEC>
__device__ __forceinline__ void load_argmax(const uint8_t* p, int (&a)[VEC]) {
  if constexpr (VEC == 4) {
    const uchar4 t = *reinterpret_cast<const uchar4*>(p);
    a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
  } else {
    a[0] = *p;
  }
}

// SELF_SKIP: keys are x itself and the centre is the row (graph built by the k-NN op).
template <typename T, int VEC, bool I64, bool SELF_SKIP>
__global__ void __launch_bounds__(kThreads)
mr_aggregate_bwd_dense_kernel(const T* __restrict__ g, const uint8_t* __restrict__ argmax,
                              const void* __restrict__ nbr, T* __restrict__ grad_x, long long rows, int N, int C,
                              int k, bool centre_is_row) {
  const int cv = C / VEC;
  const long long items = rows * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long row = it / cv;
    const int c = static_cast<int>(it - row * cv) * VEC;
    float g0[VEC], g1[VEC];
    load_grad_pair<T, VEC>(g + row * 2 * C + 2 * c, g0, g1);
    float r[VEC];
    if constexpr (SELF_SKIP) {
      int a[VEC];
      load_argmax<VEC>(argmax + row * C + c, a);
      const int n = static_cast<int>(row % N);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int nb = load_index<I64>(nbr, row * k + a[e]);
        r[e] = (nb == n) ? g0[e] : g0[e] - g1[e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) r[e] = centre_is_row ? g0[e] - g1[e] : g0[e];
    }
    Pack<T, VEC>::store(grad_x + row * C + c, r);
  }
}

template <typename T, int VEC, bool I64, bool HAS_CTR, bool SELF_SKIP>
__global__ void __launch_bounds__(kThreads)
mr_aggregate_bwd_scatter_kernel(const T* __restrict__ g, const uint8_t* __restrict__ argmax,
                                const void* __restrict__ nbr, const void* __restrict__ ctr, T* __restrict__ grad_x,
                                T* __restrict__ grad_src, long long rows, int N, int M, int C, int k) {
  const int cv = C / VEC;
  const long long items = rows * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long row = it / cv;
    const int c = static_cast<int>(it - row * cv) * VEC;
    const long long b = row / N;
    const int n = static_cast<int>(row - b * N);
    float g0[VEC], g1[VEC];
    load_grad_pair<T, VEC>(g + row * 2 * C + 2 * c, g0, g1);
    int a[VEC];
    load_argmax<VEC>(argmax + row * C + c, a);
    int nb[VEC];
    bool same = true;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      nb[e] = load_index<I64>(nbr, row * k + a[e]);
      same = same && (nb[e] == nb[0]) && (a[e] == a[0]);
    }
    if (same) {  // the common case: one 16-byte reduction per role
      if (!(SELF_SKIP && nb[0] == n)) {
        Pack<T, VEC>::red_add(grad_src + (b * M + nb[0]) * (long long)C + c, g1);
      }
      if constexpr (HAS_CTR) {
        const int ci = load_index<I64>(ctr, row * k + a[0]);
        float neg[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) neg[e] = -g1[e];
        Pack<T, VEC>::red_add(grad_x + (b * N + ci) * (long long)C + c, neg);
      }
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float one[1] = {g1[e]};
        if (!(SELF_SKIP && nb[e] == n)) {
          Pack<T, 1>::red_add(grad_src + (b * M + nb[e]) * (long long)C + c + e, one);
        }
        if constexpr (HAS_CTR) {
          const int ci = load_index<I64>(ctr, row * k + a[e]);
          float neg[1] = {-g1[e]};
          Pack<T, 1>::red_add(grad_x + (b * N + ci) * (long long)C + c + e, neg);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// K3 (fused form, the one the k-NN graphs of the encoder use): one thread-block cluster per
// segment.  Phase 1: every CTA streams its share of grad_out rows once, writes the dense part
// of grad_x with plain stores and parks g[.., 2c+1] + argmax in shared memory.  A cluster
// barrier orders all dense stores of the segment before phase 2, which routes the parked
// values to their winning neighbour rows with vector reductions (red.global.add.v4.f32) that
// hit the just-written, L2-resident grad_x rows.  HBM traffic = algorithmic bytes: grad_out,
// argmax and the ids are read once, grad_x is written once.
// ------------------------------------------------------------------------------------
constexpr int kFusedThreads = 256;

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <typename T, bool I64, int U>
__global__ void __launch_bounds__(kFusedThreads, 4)
mr_aggregate_bwd_fused_kernel(const T* __restrict__ g, const uint8_t* __restrict__ argmax, const void* __restrict__ nbr,
                              T* __restrict__ grad_x, int N, int C, int k, int rows_per_cta, int cv_shift) {
  extern __shared__ __align__(16) unsigned char fused_smem[];
  const int cv = 1 << cv_shift;
  float4* g1s = reinterpret_cast<float4*>(fused_smem);                                     // [rows_per_cta * cv]
  unsigned int* ams = reinterpret_cast<unsigned int*>(fused_smem + (size_t)rows_per_cta * cv * 16);
  const unsigned csize = cluster_nctarank();
  const long long b = blockIdx.x / csize;
  const int row0 = static_cast<int>(cluster_ctarank()) * rows_per_cta;
  const int nrows = max(0, min(rows_per_cta, N - row0));
  const int items = nrows << cv_shift;
  const T* gb = g + b * (long long)N * 2 * C;
  const uint8_t* ab = argmax + b * (long long)N * C;
  T* gxb = grad_x + b * (long long)N * C;
  const long long ib = b * (long long)N * k;

  // phase 1: U independent items per thread in flight (loads first, then the dependent id look-ups)
  for (int it0 = threadIdx.x; it0 < items; it0 += kFusedThreads * U) {
    float g0[U][4], g1[U][4];
    unsigned int packed[U];
    int n[U], c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = min(it0 + u * kFusedThreads, items - 1);
      n[u] = row0 + (it >> cv_shift);
      c[u] = (it & (cv - 1)) * 4;
      {
        float gp[8];
        Pack8<T>::load(gb + (long long)n[u] * 2 * C + 2 * c[u], gp);
#pragma unroll
        for (int e = 0; e < 4; ++e) { g0[u][e] = gp[2 * e]; g1[u][e] = gp[2 * e + 1]; }
      }
      packed[u] = __ldg(reinterpret_cast<const unsigned int*>(ab + (long long)n[u] * C + c[u]));
    }
    int nb[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e) nb[u][e] = load_index<I64>(nbr, ib + n[u] * k + ((packed[u] >> (8 * e)) & 0xff));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = it0 + u * kFusedThreads;
      if (it >= items) continue;
      float r[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) r[e] = (nb[u][e] == n[u]) ? g0[u][e] : g0[u][e] - g1[u][e];
      Pack<T, 4>::store(gxb + (long long)n[u] * C + c[u], r);
      g1s[it] = make_float4(g1[u][0], g1[u][1], g1[u][2], g1[u][3]);
      ams[it] = packed[u];
    }
  }
  __threadfence();
  cluster_sync_all();

  // phase 2: route the parked values to the winning neighbour rows of this segment
  for (int it = threadIdx.x; it < items; it += kFusedThreads) {
    const int n = row0 + (it >> cv_shift);
    const int c = (it & (cv - 1)) * 4;
    const float4 gv = g1s[it];
    const unsigned int packed = ams[it];
    const float g1[4] = {gv.x, gv.y, gv.z, gv.w};
    // One 16-byte reduction per distinct winning neighbour (channels that picked another neighbour add 0):
    // the reduction issue rate, not the bytes, bounds this phase.
    int a[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) a[e] = (packed >> (8 * e)) & 0xff;
    unsigned todo = 0xf;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (todo & (1u << e)) {
        const int j = a[e];
        float v[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          const bool hit = (a[f] == j);
          v[f] = hit ? g1[f] : 0.f;
          if (hit) todo &= ~(1u << f);
        }
        const int nb = load_index<I64>(nbr, ib + n * k + j);
        if (nb != n) Pack<T, 4>::red_add(gxb + (long long)nb * C + c, v);
      }
    }
  }
}

template <typename T, bool I64, int U>
int launch_mr_bwd_fused(const T* g, const uint8_t* argmax, const void* nbr, T* grad_x, int B, int N, int C, int k,
                        cudaStream_t s, bool* launched) {
  *launched = false;
  const int cv = C / 4;
  if ((cv & (cv - 1)) != 0 || !aligned32(g)) return GRAFP_OK;  // C/4 a power of two (shift/mask indexing), 256-bit loads
  int cv_shift = 0;
  while ((1 << cv_shift) < cv) ++cv_shift;
  // cluster of 8 CTAs per segment when the per-CTA share (g1 stash 4 B + argmax 1 B per element) fits
  // 48 KB (4 CTAs resident per SM in different phases), else the smallest cluster that fits 96 KB
  int cl = 0;
  if (N >= 64 && (long long)((N + 7) / 8) * C * 5 <= 48 * 1024) cl = 8;
  else {
    for (int cand : {2, 4, 8}) {
      const long long rows = (N + cand - 1) / cand;
      if (rows * C * 5 <= 96 * 1024) { cl = cand; break; }
    }
  }
  if (cl == 0) return GRAFP_OK;
  if ((long long)B * cl > 0x7fffffffLL) return GRAFP_OK;
  const int rows_per_cta = (N + cl - 1) / cl;
  const size_t smem = (size_t)rows_per_cta * C * 5;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mr_aggregate_bwd_fused_kernel<T, I64, U>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(mr_bwd_fused): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * cl));
  cfg.blockDim = dim3(kFusedThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, mr_aggregate_bwd_fused_kernel<T, I64, U>, g, argmax, nbr, grad_x, N, C, k,
                                     rows_per_cta, cv_shift);
  if (e != cudaSuccess) { set_error("mr_aggregate_bwd_fused launch: %s", cudaGetErrorString(e)); return (int)e; }
  *launched = true;
  return check_launch("mr_aggregate_bwd_fused");
}

// ------------------------------------------------------------------------------------
// plain gather (torch_nn.py:79-98) and its scatter-add backward
// ------------------------------------------------------------------------------------
template <typename T, int VEC, bool I64>
__global__ void __launch_bounds__(kThreads)
gather_fwd_kernel(const T* __restrict__ src, const void* __restrict__ idx, T* __restrict__ out, long long edges,
                  int N, int M, int C, int k) {
  const int cv = C / VEC;
  const long long items = edges * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long edge = it / cv;  // (b, n, j) flattened
    const int c = static_cast<int>(it - edge * cv) * VEC;
    const long long b = edge / ((long long)N * k);
    const int nb = load_index<I64>(idx, edge);
    float v[VEC];
    Pack<T, VEC>::load(src + (b * M + nb) * (long long)C + c, v);
    Pack<T, VEC>::store(out + edge * C + c, v);
  }
}

template <typename T, int VEC, bool I64>
__global__ void __launch_bounds__(kThreads)
gather_bwd_kernel(const T* __restrict__ g, const void* __restrict__ idx, T* __restrict__ grad_src, long long edges,
                  int N, int M, int C, int k) {
  const int cv = C / VEC;
  const long long items = edges * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long edge = it / cv;
    const int c = static_cast<int>(it - edge * cv) * VEC;
    const long long b = edge / ((long long)N * k);
    const int nb = load_index<I64>(idx, edge);
    float v[VEC];
    Pack<T, VEC>::load(g + edge * C + c, v);
    Pack<T, VEC>::red_add(grad_src + (b * M + nb) * (long long)C + c, v);
  }
}

// ------------------------------------------------------------------------------------
// EdgeConv features [x_i | x_j - x_i] (torch_vertex.py:46-51) and backward
// ------------------------------------------------------------------------------------
template <typename T, int VEC, bool I64, bool HAS_CTR>
__global__ void __launch_bounds__(kThreads)
edge_gather_fwd_kernel(const T* __restrict__ x, const T* __restrict__ src, const void* __restrict__ nbr,
                       const void* __restrict__ ctr, T* __restrict__ out, long long edges, int N, int M, int C,
                       int k) {
  const int cv = C / VEC;
  const long long items = edges * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long edge = it / cv;
    const int c = static_cast<int>(it - edge * cv) * VEC;
    const long long row = edge / k;
    const long long b = row / N;
    const int nb = load_index<I64>(nbr, edge);
    long long crow = row;
    if constexpr (HAS_CTR) crow = b * N + load_index<I64>(ctr, edge);
    float xi[VEC], xj[VEC], d[VEC];
    Pack<T, VEC>::load(x + crow * C + c, xi);
    Pack<T, VEC>::load(src + (b * M + nb) * (long long)C + c, xj);
#pragma unroll
    for (int e = 0; e < VEC; ++e) d[e] = xj[e] - xi[e];
    Pack<T, VEC>::store(out + edge * 2 * C + c, xi);
    Pack<T, VEC>::store(out + edge * 2 * C + C + c, d);
  }
}

// dense part (centre == row): grad_x[row][c] = sum_j (ga[row][j][c] - gb[row][j][c]); plain store
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
edge_gather_bwd_dense_kernel(const T* __restrict__ g, T* __restrict__ grad_x, long long rows, int C, int k) {
  const int cv = C / VEC;
  const long long items = rows * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long row = it / cv;
    const int c = static_cast<int>(it - row * cv) * VEC;
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    for (int j = 0; j < k; ++j) {
      float ga[VEC], gb[VEC];
      Pack<T, VEC>::load(g + (row * k + j) * 2 * C + c, ga);
      Pack<T, VEC>::load(g + (row * k + j) * 2 * C + C + c, gb);
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] += ga[e] - gb[e];
    }
    Pack<T, VEC>::store(grad_x + row * C + c, acc);
  }
}

template <typename T, int VEC, bool I64, bool HAS_CTR>
__global__ void __launch_bounds__(kThreads)
edge_gather_bwd_scatter_kernel(const T* __restrict__ g, const void* __restrict__ nbr, const void* __restrict__ ctr,
                               T* __restrict__ grad_x, T* __restrict__ grad_src, long long edges, int N, int M,
                               int C, int k) {
  const int cv = C / VEC;
  const long long items = edges * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long edge = it / cv;
    const int c = static_cast<int>(it - edge * cv) * VEC;
    const long long b = edge / ((long long)N * k);
    const int nb = load_index<I64>(nbr, edge);
    float gb[VEC];
    Pack<T, VEC>::load(g + edge * 2 * C + C + c, gb);
    Pack<T, VEC>::red_add(grad_src + (b * M + nb) * (long long)C + c, gb);
    if constexpr (HAS_CTR) {
      const int ci = load_index<I64>(ctr, edge);
      float ga[VEC];
      Pack<T, VEC>::load(g + edge * 2 * C + c, ga);
#pragma unroll
      for (int e = 0; e < VEC; ++e) ga[e] -= gb[e];
      Pack<T, VEC>::red_add(grad_x + (b * N + ci) * (long long)C + c, ga);
    }
  }
}

// ------------------------------------------------------------------------------------
// max over the neighbour axis (torch_vertex.py:51,69) and backward
// ------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
max_over_k_fwd_kernel(const T* __restrict__ h, T* __restrict__ out, uint8_t* __restrict__ argmax, long long rows,
                      int C, int k) {
  const int cv = C / VEC;
  const long long items = rows * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long row = it / cv;
    const int c = static_cast<int>(it - row * cv) * VEC;
    float best[VEC];
    int arg[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { best[e] = -INFINITY; arg[e] = 0; }
#pragma unroll 4
    for (int j = 0; j < k; ++j) {
      float v[VEC];
      Pack<T, VEC>::load(h + (row * k + j) * C + c, v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if (v[e] > best[e] || v[e] != v[e]) {
          if (!(best[e] != best[e])) { best[e] = v[e]; arg[e] = j; }
        }
      }
    }
    Pack<T, VEC>::store(out + row * C + c, best);
    if (argmax != nullptr) {
      if constexpr (VEC == 4) {
        *reinterpret_cast<uchar4*>(argmax + row * C + c) =
            make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
      } else {
        argmax[row * C + c] = (uint8_t)arg[0];
      }
    }
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
max_over_k_bwd_kernel(const T* __restrict__ g, const uint8_t* __restrict__ argmax, T* __restrict__ grad_h,
                      long long rows, int C, int k) {
  const int cv = C / VEC;
  const long long items = rows * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long row = it / cv;
    const int c = static_cast<int>(it - row * cv) * VEC;
    float gv[VEC];
    int a[VEC];
    Pack<T, VEC>::load(g + row * C + c, gv);
    load_argmax<VEC>(argmax + row * C + c, a);
    for (int j = 0; j < k; ++j) {
      float v[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] = (a[e] == j) ? gv[e] : 0.f;
      Pack<T, VEC>::store(grad_h + (row * k + j) * C + c, v);
    }
  }
}

// ------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------
namespace {

// development switches: GRAFP_MR_FWD_VARIANT = 0 generic kernel, 1 / 2 / 4 fast kernel with that many items in
// flight per thread (default 4); GRAFP_MR_BWD_VARIANT = 0 two-kernel form, 1 / 2 / 4 fused cluster form (default 2)
int fwd_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GRAFP_MR_FWD_VARIANT");
    v = e ? atoi(e) : 4;
  }
  return v;
}
int bwd_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GRAFP_MR_BWD_VARIANT");
    v = e ? atoi(e) : 2;
  }
  return v;
}

template <typename F>
int dispatch_vec_idx(int C, bool ptrs_aligned, int idx_is_i64, F&& f) {
  const bool vec4 = (C % 4 == 0) && ptrs_aligned;
  if (vec4) {
    return idx_is_i64 ? f(std::integral_constant<int, 4>{}, std::true_type{})
                      : f(std::integral_constant<int, 4>{}, std::false_type{});
  }
  return idx_is_i64 ? f(std::integral_constant<int, 1>{}, std::true_type{})
                    : f(std::integral_constant<int, 1>{}, std::false_type{});
}

}  // namespace

template <typename T>
int launch_mr_aggregate_fwd(const void* x, const void* y, const void* nbr, const void* ctr, int idx_is_i64, void* out,
                            uint8_t* argmax, int B, int N, int M, int C, int k, cudaStream_t s) {
  const T* xs = static_cast<const T*>(x);
  const T* src = y ? static_cast<const T*>(y) : xs;
  const long long rows = (long long)B * N;
  const bool al = aligned16(x) && aligned16(src) && aligned16(out) && (argmax == nullptr || ((uintptr_t)argmax & 3) == 0);
  return dispatch_vec_idx(C, al, idx_is_i64, [&](auto vec, auto i64) {
    constexpr int VEC = decltype(vec)::value;
    constexpr bool I64 = decltype(i64)::value;
    const int grid = grid_for(rows * (C / VEC), kThreads, 8);
    if constexpr (VEC == 4) {
      const int cv = C / 4;
      const int variant = fwd_variant();
      if (!ctr && variant > 0 && (cv & (cv - 1)) == 0 && cv <= kThreads && B <= 65535 && aligned32(out)) {
        int cv_shift = 0;
        while ((1 << cv_shift) < cv) ++cv_shift;
        const int rpb = kThreads >> cv_shift;
        const int u = variant;  // items in flight per thread: 1, 2 or 4
        const int passes = (N + rpb * u - 1) / (rpb * u);
        int gx = (num_sms() * 8 + B - 1) / B;  // about 8 resident CTAs per SM across the whole batch
        if (gx > passes) gx = passes;
        if (gx < 1) gx = 1;
        dim3 grid2(gx, B);
        T* o = static_cast<T*>(out);
        if (u == 1) mr_aggregate_fwd_fast_kernel<T, I64, 1><<<grid2, kThreads, 0, s>>>(xs, src, nbr, o, argmax, N, M, C, k, cv_shift);
        else if (u == 2) mr_aggregate_fwd_fast_kernel<T, I64, 2><<<grid2, kThreads, 0, s>>>(xs, src, nbr, o, argmax, N, M, C, k, cv_shift);
        else mr_aggregate_fwd_fast_kernel<T, I64, 4><<<grid2, kThreads, 0, s>>>(xs, src, nbr, o, argmax, N, M, C, k, cv_shift);
        return check_launch("mr_aggregate_fwd");
      }
    }
    if (ctr) {
      mr_aggregate_fwd_kernel<T, VEC, I64, true><<<grid, kThreads, 0, s>>>(xs, src, nbr, ctr, static_cast<T*>(out),
                                                                            argmax, rows, N, M, C, k);
    } else {
      mr_aggregate_fwd_kernel<T, VEC, I64, false><<<grid, kThreads, 0, s>>>(xs, src, nbr, ctr, static_cast<T*>(out),
                                                                             argmax, rows, N, M, C, k);
    }
    return check_launch("mr_aggregate_fwd");
  });
}

template <typename T>
int launch_mr_aggregate_bwd(const void* g, const uint8_t* argmax, const void* nbr, const void* ctr, int idx_is_i64,
                            void* grad_x, void* grad_y, int B, int N, int M, int C, int k, cudaStream_t s) {
  const long long rows = (long long)B * N;
  T* gx = static_cast<T*>(grad_x);
  T* gsrc = grad_y ? static_cast<T*>(grad_y) : gx;
  const bool al = aligned16(g) && aligned16(grad_x) && aligned16(gsrc) && (((uintptr_t)argmax & 3) == 0);
  if (grad_y) {
    cudaError_t e = cudaMemsetAsync(grad_y, 0, (size_t)B * M * C * sizeof(T), s);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync(grad_y): %s", cudaGetErrorString(e)); return (int)e; }
  }
  return dispatch_vec_idx(C, al, idx_is_i64, [&](auto vec, auto i64) {
    constexpr int VEC = decltype(vec)::value;
    constexpr bool I64 = decltype(i64)::value;
    const int grid = grid_for(rows * (C / VEC), kThreads, 8);
    const T* gs = static_cast<const T*>(g);
    const bool self_skip = (ctr == nullptr) && (grad_y == nullptr);
    if constexpr (VEC == 4) {
      const int bv = bwd_variant();  // 0: two-kernel form; 1 / 2 / 4: fused form with that many items in flight
      if (self_skip && bv > 0) {
        bool launched = false;
        const int rc = bv == 1 ? launch_mr_bwd_fused<T, I64, 1>(gs, argmax, nbr, gx, B, N, C, k, s, &launched)
                     : bv == 2 ? launch_mr_bwd_fused<T, I64, 2>(gs, argmax, nbr, gx, B, N, C, k, s, &launched)
                               : launch_mr_bwd_fused<T, I64, 4>(gs, argmax, nbr, gx, B, N, C, k, s, &launched);
        if (rc != GRAFP_OK || launched) return rc;
      }
    }
    if (self_skip) {
      mr_aggregate_bwd_dense_kernel<T, VEC, I64, true><<<grid, kThreads, 0, s>>>(gs, argmax, nbr, gx, rows, N, C, k, true);
      mr_aggregate_bwd_scatter_kernel<T, VEC, I64, false, true><<<grid, kThreads, 0, s>>>(gs, argmax, nbr, ctr, gx, gsrc,
                                                                                         rows, N, M, C, k);
    } else if (ctr == nullptr) {
      mr_aggregate_bwd_dense_kernel<T, VEC, I64, false><<<grid, kThreads, 0, s>>>(gs, argmax, nbr, gx, rows, N, C, k, true);
      mr_aggregate_bwd_scatter_kernel<T, VEC, I64, false, false><<<grid, kThreads, 0, s>>>(gs, argmax, nbr, ctr, gx, gsrc,
                                                                                          rows, N, M, C, k);
    } else {
      mr_aggregate_bwd_dense_kernel<T, VEC, I64, false><<<grid, kThreads, 0, s>>>(gs, argmax, nbr, gx, rows, N, C, k, false);
      mr_aggregate_bwd_scatter_kernel<T, VEC, I64, true, false><<<grid, kThreads, 0, s>>>(gs, argmax, nbr, ctr, gx, gsrc,
                                                                                         rows, N, M, C, k);
    }
    return check_launch("mr_aggregate_bwd");
  });
}

template <typename T>
int launch_gather_fwd(const void* src, const void* idx, int idx_is_i64, void* out, int B, int N, int M, int C, int k,
                      cudaStream_t s) {
  const long long edges = (long long)B * N * k;
  return dispatch_vec_idx(C, aligned16(src) && aligned16(out), idx_is_i64, [&](auto vec, auto i64) {
    constexpr int VEC = decltype(vec)::value;
    constexpr bool I64 = decltype(i64)::value;
    const int grid = grid_for(edges * (C / VEC), kThreads, 8);
    gather_fwd_kernel<T, VEC, I64><<<grid, kThreads, 0, s>>>(static_cast<const T*>(src), idx, static_cast<T*>(out),
                                                             edges, N, M, C, k);
    return check_launch("gather_fwd");
  });
}

template <typename T>
int launch_gather_bwd(const void* g, const void* idx, int idx_is_i64, void* grad_src, int B, int N, int M, int C, int k,
                      cudaStream_t s) {
  const long long edges = (long long)B * N * k;
  cudaError_t e = cudaMemsetAsync(grad_src, 0, (size_t)B * M * C * sizeof(T), s);
  if (e != cudaSuccess) { set_error("cudaMemsetAsync(grad_src): %s", cudaGetErrorString(e)); return (int)e; }
  return dispatch_vec_idx(C, aligned16(g) && aligned16(grad_src), idx_is_i64, [&](auto vec, auto i64) {
    constexpr int VEC = decltype(vec)::value;
    constexpr bool I64 = decltype(i64)::value;
    const int grid = grid_for(edges * (C / VEC), kThreads, 8);
    gather_bwd_kernel<T, VEC, I64><<<grid, kThreads, 0, s>>>(static_cast<const T*>(g), idx, static_cast<T*>(grad_src),
                                                             edges, N, M, C, k);
    return check_launch("gather_bwd");
  });
}

template <typename T>
int launch_edge_gather_fwd(const void* x, const void* y, const void* nbr, const void* ctr, int idx_is_i64, void* out,
                           int B, int N, int M, int C, int k, cudaStream_t s) {
  const long long edges = (long long)B * N * k;
  const T* xs = static_cast<const T*>(x);
  const T* src = y ? static_cast<const T*>(y) : xs;
  return dispatch_vec_idx(C, aligned16(x) && aligned16(src) && aligned16(out), idx_is_i64, [&](auto vec, auto i64) {
    constexpr int VEC = decltype(vec)::value;
    constexpr bool I64 = decltype(i64)::value;
    const int grid = grid_for(edges * (C / VEC), kThreads, 8);
    if (ctr) {
      edge_gather_fwd_kernel<T, VEC, I64, true><<<grid, kThreads, 0, s>>>(xs, src, nbr, ctr, static_cast<T*>(out), edges,
                                                                           N, M, C, k);
    } else {
      edge_gather_fwd_kernel<T, VEC, I64, false><<<grid, kThreads, 0, s>>>(xs, src, nbr, ctr, static_cast<T*>(out), edges,
                                                                            N, M, C, k);
    }
    return check_launch("edge_gather_fwd");
  });
}

template <typename T>
int launch_edge_gather_bwd(const void* g, const void* nbr, const void* ctr, int idx_is_i64, void* grad_x, void* grad_y,
                           int B, int N, int M, int C, int k, cudaStream_t s) {
  const long long rows = (long long)B * N;
  const long long edges = rows * k;
  T* gx = static_cast<T*>(grad_x);
  T* gsrc = grad_y ? static_cast<T*>(grad_y) : gx;
  cudaError_t e = cudaSuccess;
  if (grad_y) e = cudaMemsetAsync(grad_y, 0, (size_t)B * M * C * sizeof(T), s);
  if (e == cudaSuccess && ctr) e = cudaMemsetAsync(grad_x, 0, (size_t)B * N * C * sizeof(T), s);
  if (e != cudaSuccess) { set_error("cudaMemsetAsync(edge grads): %s", cudaGetErrorString(e)); return (int)e; }
  return dispatch_vec_idx(C, aligned16(g) && aligned16(grad_x) && aligned16(gsrc), idx_is_i64, [&](auto vec, auto i64) {
    constexpr int VEC = decltype(vec)::value;
    constexpr bool I64 = decltype(i64)::value;
    const T* gs = static_cast<const T*>(g);
    if (ctr) {
      const int grid = grid_for(edges * (C / VEC), kThreads, 8);
      edge_gather_bwd_scatter_kernel<T, VEC, I64, true><<<grid, kThreads, 0, s>>>(gs, nbr, ctr, gx, gsrc, edges, N, M, C, k);
    } else {
      const int grid_d = grid_for(rows * (C / VEC), kThreads, 8);
      edge_gather_bwd_dense_kernel<T, VEC><<<grid_d, kThreads, 0, s>>>(gs, gx, rows, C, k);
      const int grid = grid_for(edges * (C / VEC), kThreads, 8);
      edge_gather_bwd_scatter_kernel<T, VEC, I64, false><<<grid, kThreads, 0, s>>>(gs, nbr, ctr, gx, gsrc, edges, N, M, C, k);
    }
    return check_launch("edge_gather_bwd");
  });
}

template <typename T>
int launch_max_over_k_fwd(const void* h, void* out, uint8_t* argmax, int B, int N, int C, int k, cudaStream_t s) {
  const long long rows = (long long)B * N;
  const bool vec4 = (C % 4 == 0) && aligned16(h) && aligned16(out) && (argmax == nullptr || ((uintptr_t)argmax & 3) == 0);
  if (vec4) {
    max_over_k_fwd_kernel<T, 4><<<grid_for(rows * (C / 4), kThreads, 8), kThreads, 0, s>>>(
        static_cast<const T*>(h), static_cast<T*>(out), argmax, rows, C, k);
  } else {
    max_over_k_fwd_kernel<T, 1><<<grid_for(rows * C, kThreads, 8), kThreads, 0, s>>>(
        static_cast<const T*>(h), static_cast<T*>(out), argmax, rows, C, k);
  }
  return check_launch("max_over_k_fwd");
}

template <typename T>
int launch_max_over_k_bwd(const void* g, const uint8_t* argmax, void* grad_h, int B, int N, int C, int k,
                          cudaStream_t s) {
  const long long rows = (long long)B * N;
  const bool vec4 = (C % 4 == 0) && aligned16(g) && aligned16(grad_h) && (((uintptr_t)argmax & 3) == 0);
  if (vec4) {
    max_over_k_bwd_kernel<T, 4><<<grid_for(rows * (C / 4), kThreads, 8), kThreads, 0, s>>>(
        static_cast<const T*>(g), argmax, static_cast<T*>(grad_h), rows, C, k);
  } else {
    max_over_k_bwd_kernel<T, 1><<<grid_for(rows * C, kThreads, 8), kThreads, 0, s>>>(
        static_cast<const T*>(g), argmax, static_cast<T*>(grad_h), rows, C, k);
  }
  return check_launch("max_over_k_bwd");
}

#define GRAFP_INSTANTIATE(T)                                                                                          \
  template int launch_mr_aggregate_fwd<T>(const void*, const void*, const void*, const void*, int, void*, uint8_t*,  \
                                          int, int, int, int, int, cudaStream_t);                                    \
  template int launch_mr_aggregate_bwd<T>(const void*, const uint8_t*, const void*, const void*, int, void*, void*,  \
                                          int, int, int, int, int, cudaStream_t);                                    \
  template int launch_gather_fwd<T>(const void*, const void*, int, void*, int, int, int, int, int, cudaStream_t);    \
  template int launch_gather_bwd<T>(const void*, const void*, int, void*, int, int, int, int, int, cudaStream_t);    \
  template int launch_edge_gather_fwd<T>(const void*, const void*, const void*, const void*, int, void*, int, int,   \
                                         int, int, int, cudaStream_t);                                               \
  template int launch_edge_gather_bwd<T>(const void*, const void*, const void*, int, void*, void*, int, int, int,    \
                                         int, int, cudaStream_t);                                                    \
  template int launch_max_over_k_fwd<T>(const void*, void*, uint8_t*, int, int, int, int, cudaStream_t);             \
  template int launch_max_over_k_bwd<T>(const void*, const uint8_t*, void*, int, int, int, int, cudaStream_t);

GRAFP_INSTANTIATE(float)
GRAFP_INSTANTIATE(__nv_bfloat16)

}  // namespace grafp
