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
// K2, pipelined form (k = KN neighbours, centre == row): persistent CTAs, one item = (row, 16 bytes of
// channels: 4 fp32 / 8 bf16) per thread and iteration.  The item's own 16 bytes and its KN
// neighbour slices are fetched with cp.async into a per-thread shared-memory slot D iterations
// ahead, so D * (KN + 1) * 16 bytes per thread are in flight without holding registers, and the
// neighbour ids (the dependent load in front of the gathers) are read one iteration earlier
// still.  The gathers are L2 hits (a segment's features are 256 KiB); HBM sees the algorithmic
// bytes only: x and the ids once, the interleaved output and the argmax once.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// 16 bytes of node features as fp32 registers, and the matching output / argmax / gradient forms
template <typename T>
struct Item16;

template <>
struct Item16<float> {
  static constexpr int V = 4;
  static __device__ __forceinline__ void unpack(const uint4& q, float (&v)[4]) {
    v[0] = __uint_as_float(q.x); v[1] = __uint_as_float(q.y); v[2] = __uint_as_float(q.z); v[3] = __uint_as_float(q.w);
  }
  static __device__ __forceinline__ float round_like_t(float d) { return d; }
  // [x_0, m_0, x_1, m_1, ...]: 2 V values = 32 bytes, one 256-bit store
  static __device__ __forceinline__ void store_interleaved(float* p, const float (&x)[4], const float (&m)[4]) {
    const float il[8] = {x[0], m[0], x[1], m[1], x[2], m[2], x[3], m[3]};
    Pack8<float>::store(p, il);
  }
  static __device__ __forceinline__ void store_argmax(uint8_t* p, const int (&a)[4]) {
    *reinterpret_cast<unsigned int*>(p) = (unsigned)a[0] | ((unsigned)a[1] << 8) | ((unsigned)a[2] << 16) | ((unsigned)a[3] << 24);
  }
  // 32 bytes of interleaved gradient [g0_0, g1_0, g0_1, g1_1, ...] -> the two planes
  static __device__ __forceinline__ void unpack_pairs(const uint4& a, const uint4& b, float (&g0)[4], float (&g1)[4]) {
    g0[0] = __uint_as_float(a.x); g1[0] = __uint_as_float(a.y); g0[1] = __uint_as_float(a.z); g1[1] = __uint_as_float(a.w);
    g0[2] = __uint_as_float(b.x); g1[2] = __uint_as_float(b.y); g0[3] = __uint_as_float(b.z); g1[3] = __uint_as_float(b.w);
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ void red_add(float* p, const float (&v)[4]) { Pack<float, 4>::red_add(p, v); }
};

template <>
struct Item16<__nv_bfloat16> {
  static constexpr int V = 8;
  static __device__ __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
  static __device__ __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ void unpack(const uint4& q, float (&v)[8]) {
    v[0] = lo(q.x); v[1] = hi(q.x); v[2] = lo(q.y); v[3] = hi(q.y);
    v[4] = lo(q.z); v[5] = hi(q.z); v[6] = lo(q.w); v[7] = hi(q.w);
  }
  // the reference subtracts in bf16: round the difference before it is compared
  static __device__ __forceinline__ float round_like_t(float d) { return __bfloat162float(__float2bfloat16_rn(d)); }
  static __device__ __forceinline__ void store_interleaved(__nv_bfloat16* p, const float (&x)[8], const float (&m)[8]) {
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = pack2(x[e], m[e]);
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
  }
  static __device__ __forceinline__ void store_argmax(uint8_t* p, const int (&a)[8]) {
    uint2 t;
    t.x = (unsigned)a[0] | ((unsigned)a[1] << 8) | ((unsigned)a[2] << 16) | ((unsigned)a[3] << 24);
    t.y = (unsigned)a[4] | ((unsigned)a[5] << 8) | ((unsigned)a[6] << 16) | ((unsigned)a[7] << 24);
    *reinterpret_cast<uint2*>(p) = t;
  }
  static __device__ __forceinline__ void unpack_pairs(const uint4& a, const uint4& b, float (&g0)[8], float (&g1)[8]) {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) { g0[e] = lo(w[e]); g1[e] = hi(w[e]); }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
  }
  static __device__ __forceinline__ void red_add(__nv_bfloat16* p, const float (&v)[8]) {
    asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(pack2(v[0], v[1])),
                 "r"(pack2(v[2], v[3])), "r"(pack2(v[4], v[5])), "r"(pack2(v[6], v[7]))
                 : "memory");
  }
};

template <typename T, bool I64, int KN, int D>
__global__ void __launch_bounds__(kThreads)
mr_aggregate_fwd_pipe_kernel(const T* __restrict__ x, const T* __restrict__ src, const void* __restrict__ nbr,
                             T* __restrict__ out, uint8_t* __restrict__ argmax, int N, int M, int C, int cv_shift,
                             int ips_shift, int total_iters) {
  using It = Item16<T>;
  constexpr int V = It::V;
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  // slot (stage s, operand o, thread t) -> 16 bytes; consecutive threads are consecutive: conflict-free
  const uint32_t slot0 = static_cast<uint32_t>(__cvta_generic_to_shared(pipe_smem)) + threadIdx.x * 16;
  auto slot = [&](int s, int o) { return slot0 + (uint32_t)((s * (KN + 1) + o) * kThreads * 16); };
  const int cmask = (1 << cv_shift) - 1;
  const int c = (threadIdx.x & cmask) * V;  // 256 % cv == 0: the channel pack of a thread never changes

  // work item w = (segment b, chunk of 256 items): b = w >> ips_shift
  auto row_of = [&](int w, long long& rowg, int& b) {
    b = w >> ips_shift;
    const int chunk = w & ((1 << ips_shift) - 1);
    const int n = (chunk * kThreads + (int)threadIdx.x) >> cv_shift;
    rowg = (long long)b * N + n;
  };
  auto load_ids = [&](int w, int (&ids)[KN]) {
    long long rowg; int b;
    row_of(w, rowg, b);
#pragma unroll
    for (int j = 0; j < KN; ++j) ids[j] = load_index<I64>(nbr, rowg * KN + j);
  };
  auto issue = [&](int w, int s, const int (&ids)[KN]) {
    long long rowg; int b;
    row_of(w, rowg, b);
    cp_async16(slot(s, 0), x + rowg * C + c);
    const T* sb = src + (long long)b * M * C + c;
#pragma unroll
    for (int j = 0; j < KN; ++j) cp_async16(slot(s, 1 + j), sb + (long long)ids[j] * C);
  };

  const int w0 = blockIdx.x, stride = gridDim.x;
  int ids[KN];
#pragma unroll
  for (int s = 0; s < D; ++s) {
    const int w = w0 + s * stride;
    if (w < total_iters) { load_ids(w, ids); issue(w, s, ids); }
    cp_async_commit();
  }
  if (w0 + D * stride < total_iters) load_ids(w0 + D * stride, ids);

  int s = 0;
  for (int w = w0; w < total_iters; w += stride) {
    cp_async_wait<D - 1>();
    float self[V], best[V];
    int arg[V];
    It::unpack(*reinterpret_cast<const uint4*>(pipe_smem + (slot(s, 0) - (slot0 - threadIdx.x * 16))), self);
#pragma unroll
    for (int e = 0; e < V; ++e) { best[e] = -INFINITY; arg[e] = 0; }
    float dj[KN][V];
    float poison = 0.f;
#pragma unroll
    for (int j = 0; j < KN; ++j) {
      float xj[V];
      It::unpack(*reinterpret_cast<const uint4*>(pipe_smem + (slot(s, 1 + j) - (slot0 - threadIdx.x * 16))), xj);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float d = It::round_like_t(xj[e] - self[e]);
        dj[j][e] = d;
        poison = fmaf(d, 0.f, poison);      // NaN / inf anywhere -> NaN
        const bool gt = d > best[e];        // strict: the first maximiser wins (torch.max)
        best[e] = gt ? d : best[e];
        arg[e] = gt ? j : arg[e];
      }
    }
    if (poison != 0.f) {  // redo with exact torch.max semantics (NaN propagates, first NaN wins)
#pragma unroll
      for (int e = 0; e < V; ++e) { best[e] = -INFINITY; arg[e] = 0; }
#pragma unroll
      for (int j = 0; j < KN; ++j) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float d = dj[j][e];
          if (d > best[e] || d != d) {
            if (!(best[e] != best[e])) { best[e] = d; arg[e] = j; }
          }
        }
      }
    }
    long long rowg; int b;
    row_of(w, rowg, b);
    It::store_interleaved(out + rowg * 2 * C + 2 * c, self, best);
    if (argmax != nullptr) It::store_argmax(argmax + rowg * C + c, arg);
    // refill this slot D iterations ahead (the slot's values are in registers / stored by now)
    const int wn = w + D * stride;
    if (wn < total_iters) issue(wn, s, ids);
    cp_async_commit();
    if (wn + stride < total_iters) load_ids(wn + stride, ids);
    if (++s == D) s = 0;
  }
  cp_async_wait<0>();
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
// K3 (cluster form, the one the k-NN graphs of the encoder use): one thread-block cluster per
// segment.  A CTA's whole share of the segment - its rows of grad_out (contiguous) and their
// neighbour ids - is pulled into shared memory by two cp.async.bulk copies issued by one thread and
// awaited on an mbarrier, while the argmax bytes stream into registers.  Phase 1 writes the dense
// part of grad_x with plain stores; a cluster barrier (release / acquire at cluster scope) orders all
// dense stores of the segment before phase 2, which routes g[.., 2c+1] to the winning neighbour rows
// with vector reductions (red.global.add.v4.f32 / .v4.bf16x2) that hit the just-written, L2-resident
// grad_x rows.  HBM traffic = algorithmic bytes: grad_out, argmax and the ids are read once, grad_x
// is written once.  Phase 2 walks the neighbour slots j = 0..k-1 in a fixed order, so the lanes that
// share a source row issue their reduction to the SAME target row in the same instruction
// (contiguous runs that the LSU / L2 merge per sector) and an item issues at most k-1 of them.
// An item is (row, 16 bytes of channels): 4 fp32 or 8 bf16 channels.
// ------------------------------------------------------------------------------------
constexpr int kFusedThreads = 256;

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void k3_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void k3_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void k3_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void k3_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "K3_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra K3_WAIT_DONE;\n\t"
      "bra K3_WAIT_LOOP;\n\t"
      "K3_WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}

// FENCE: a device-scope fence in front of the cluster barrier (the round-1 form; the barrier's own release / acquire
// at cluster scope already orders the dense stores before the reductions of the other CTAs of the cluster).
template <typename T, bool I64, bool FENCE>
__global__ void __launch_bounds__(kFusedThreads, 3)
mr_aggregate_bwd_cluster_kernel(const T* __restrict__ g, const uint8_t* __restrict__ argmax,
                                const void* __restrict__ nbr, T* __restrict__ grad_x, int N, int C, int k,
                                int rows_per_cta, int cv_shift) {
  using It = Item16<T>;
  constexpr int V = It::V;
  using idx_t = typename std::conditional<I64, long long, int>::type;
  constexpr int kItems = 2048 / kFusedThreads;  // items per thread: the host caps a CTA's share at 2048 items
  extern __shared__ __align__(128) unsigned char tma_smem[];
  const int cv = 1 << cv_shift;
  const size_t row_bytes = (size_t)2 * C * sizeof(T);
  const unsigned char* gt = tma_smem;                                                       // [rows][2C] grad_out
  const idx_t* ids = reinterpret_cast<const idx_t*>(tma_smem + (size_t)rows_per_cta * row_bytes);  // [rows][k] neighbour ids
  const uint32_t bar = static_cast<uint32_t>(
      __cvta_generic_to_shared(tma_smem + (size_t)rows_per_cta * (row_bytes + k * sizeof(idx_t))));
  const unsigned csize = cluster_nctarank();
  const long long b = blockIdx.x / csize;
  const int row0 = static_cast<int>(cluster_ctarank()) * rows_per_cta;
  const int nrows = max(0, min(rows_per_cta, N - row0));
  const int items = nrows << cv_shift;
  T* gxb = grad_x + b * (long long)N * C;

  if (threadIdx.x == 0) {
    k3_mbar_init(bar, 1);
    if (nrows > 0) {
      const uint32_t gbytes = (uint32_t)(nrows * row_bytes), ibytes = (uint32_t)(nrows * k * sizeof(idx_t));
      k3_mbar_expect_tx(bar, gbytes + ibytes);
      k3_bulk_g2s(static_cast<uint32_t>(__cvta_generic_to_shared(tma_smem)), g + (b * N + row0) * 2LL * C, gbytes, bar);
      k3_bulk_g2s(static_cast<uint32_t>(__cvta_generic_to_shared(ids)),
                  static_cast<const idx_t*>(nbr) + (b * N + row0) * (long long)k, ibytes, bar);
    }
  }
  // the argmax bytes stream straight into registers while the bulk copies are in flight
  unsigned int am[kItems][V / 4];
#pragma unroll
  for (int u = 0; u < kItems; ++u) {
    const int it = threadIdx.x + u * kFusedThreads;
    const uint8_t* ap = argmax + (b * N + row0 + (it >> cv_shift)) * (long long)C + (it & (cv - 1)) * V;
#pragma unroll
    for (int q = 0; q < V / 4; ++q) am[u][q] = (it < items) ? __ldg(reinterpret_cast<const unsigned int*>(ap) + q) : 0u;
  }
  __syncthreads();  // the barrier is initialised before anyone waits on it
  if (nrows > 0) k3_mbar_wait(bar, 0);

  // phase 1: dense part of grad_x for this CTA's rows (everything it needs is in shared memory or registers)
  // (Measured and dropped, round 2: a per-row bit mask of the self-edge slots instead of one id look-up per channel,
  //  and skipping self slots before the select chain in phase 2 - fewer instructions, 120 -> 130 us: the extra
  //  dependent shared-memory loads at the head of each item cost more than the selects they save.)
#pragma unroll
  for (int u = 0; u < kItems; ++u) {
    const int it = threadIdx.x + u * kFusedThreads;
    if (it < items) {
      const int rl = it >> cv_shift;
      const int n = row0 + rl;
      const int c = (it & (cv - 1)) * V;
      const uint4* gp = reinterpret_cast<const uint4*>(gt + (size_t)rl * row_bytes + (size_t)2 * c * sizeof(T));
      float g0[V], g1[V];
      It::unpack_pairs(gp[0], gp[1], g0, g1);
      float r[V];
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int nb = static_cast<int>(ids[rl * k + ((am[u][e >> 2] >> (8 * (e & 3))) & 0xff)]);
        r[e] = (nb == n) ? g0[e] : g0[e] - g1[e];
      }
      It::store(gxb + (long long)n * C + c, r);
    }
  }
  if constexpr (FENCE) __threadfence();
  cluster_sync_all();

  // phase 2: route g[.., 2c+1] to the winning neighbour rows of this segment (L2-resident, just written)
#pragma unroll
  for (int u = 0; u < kItems; ++u) {
    const int it = threadIdx.x + u * kFusedThreads;
    if (it < items) {
      const int rl = it >> cv_shift;
      const int n = row0 + rl;
      const int c = (it & (cv - 1)) * V;
      const uint4* gp = reinterpret_cast<const uint4*>(gt + (size_t)rl * row_bytes + (size_t)2 * c * sizeof(T));
      float g0[V], g1[V];
      It::unpack_pairs(gp[0], gp[1], g0, g1);
      for (int j = 0; j < k; ++j) {
        float v[V];
        bool any = false;
#pragma unroll
        for (int f = 0; f < V; ++f) {
          const bool hit = ((am[u][f >> 2] >> (8 * (f & 3))) & 0xff) == (unsigned)j;
          v[f] = hit ? g1[f] : 0.f;
          any |= hit;
        }
        const int nb = static_cast<int>(ids[rl * k + j]);
        if (any && nb != n) It::red_add(gxb + (long long)nb * C + c, v);
      }
    }
  }
}

// (Measured and dropped, round 2: a variant that loads the CTA's share of grad_out straight into registers - no bulk
//  copy, no mbarrier, no shared-memory round trip - runs in the same 120.8 us as this one (121.9 us): the kernel is bound
//  by the reductions at L2, not by how its inputs arrive.  profiles/r03c_bench_k23.log.)
template <typename T, bool I64>
int launch_mr_bwd_cluster(const T* g, const uint8_t* argmax, const void* nbr, T* grad_x, int B, int N, int C, int k,
                          bool fence, cudaStream_t s, bool* launched) {
  *launched = false;
  constexpr int V = Item16<T>::V;
  if (C % V != 0) return GRAFP_OK;
  const int cv = C / V;
  const size_t idsz = I64 ? 8 : 4;
  if ((cv & (cv - 1)) != 0 || !aligned16(g) || !aligned16(grad_x) || !aligned16(nbr) || (((uintptr_t)argmax) & 3) != 0 || N < 64)
    return GRAFP_OK;
  int cv_shift = 0;
  while ((1 << cv_shift) < cv) ++cv_shift;
  const size_t row_bytes = (size_t)2 * C * sizeof(T);
  // smallest cluster whose per-CTA share fits 8 items per thread and ~68 KB of shared memory (three CTAs per SM);
  // bulk copies need 16-byte multiples and 16-byte aligned sources for every CTA of the cluster
  int cl = 0;
  for (int cand : {1, 2, 4, 8}) {
    const long long rows = (N + cand - 1) / cand;
    if (rows * cv <= 8 * kFusedThreads && rows * (row_bytes + k * idsz) + 16 <= 68 * 1024 &&
        (rows * k * idsz) % 16 == 0 && ((long long)N * k * idsz) % 16 == 0 && (rows * row_bytes) % 16 == 0) {
      cl = cand;
      break;
    }
  }
  if (cl == 0 || (long long)B * cl > 0x7fffffffLL) return GRAFP_OK;
  const int rows_per_cta = (N + cl - 1) / cl;
  if (N % rows_per_cta != 0 && ((N % rows_per_cta) * k * idsz) % 16 != 0) return GRAFP_OK;  // ragged last share
  const size_t smem = (size_t)rows_per_cta * (row_bytes + k * idsz) + 16;
  auto launch = [&](auto kernel, DeviceOnce& once) -> int {
    if (once.pending()) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 68 * 1024);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(mr_bwd_cluster): %s", cudaGetErrorString(e)); return (int)e; }
      once.mark();
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, g, argmax, nbr, grad_x, N, C, k, rows_per_cta, cv_shift);
    if (e != cudaSuccess) { set_error("mr_aggregate_bwd_cluster launch: %s", cudaGetErrorString(e)); return (int)e; }
    *launched = true;
    return check_launch("mr_aggregate_bwd_cluster");
  };
  static DeviceOnce once_fence, once_nofence;
  if (fence) return launch(mr_aggregate_bwd_cluster_kernel<T, I64, true>, once_fence);
  return launch(mr_aggregate_bwd_cluster_kernel<T, I64, false>, once_nofence);
}

// ------------------------------------------------------------------------------------
// K3, gather form (fp32, k = KN, graphs built by the k-NN op): no atomics, no zero-fill, no
// cluster barrier, deterministic.  A small kernel first transposes each segment's graph into
// reverse CSR (who points at row n, through which neighbour slot), sorted so the summation
// order is fixed; the main kernel then owns one (row, 4-channel) item per thread and iteration:
//   grad_x[n][c] = g[n][2c] - [winner(n,c) != n] g[n][2c+1] + sum over in-edges (m, j) of [argmax[m][c] == j] g[m][2c+1]
// Its own g slice and argmax word arrive through a cp.async pipeline D iterations deep (HBM
// stream); the in-edge rows are L2 hits (a segment's grad_out is 512 KiB).  HBM traffic = the
// algorithmic bytes plus the 16 bytes per row of reverse graph.
// ------------------------------------------------------------------------------------
template <bool I64>
__global__ void __launch_bounds__(kThreads)
mr_bwd_build_reverse_kernel(const void* __restrict__ nbr, int* rev_off, unsigned int* rev_src, int N, int k) {
  extern __shared__ int rev_smem[];
  int* off = rev_smem;           // [N + 1] in-degree counts, then exclusive offsets
  int* cur = rev_smem + N + 1;   // [N] fill cursors
  __shared__ int wsum[kThreads / 32];
  const long long b = blockIdx.x;
  const long long ebase = b * N * k;
  const int E = N * k;
  for (int i = threadIdx.x; i <= N; i += kThreads) off[i] = 0;
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += kThreads) {
    const int m = e / k;
    const int t = load_index<I64>(nbr, ebase + e);
    if (t != m) atomicAdd(&off[t], 1);  // self edges cancel against the centre term and are not listed
  }
  __syncthreads();
  // exclusive scan: contiguous span per thread, warp scan of the span sums, then the warp totals
  const int span = (N + kThreads - 1) / kThreads;
  const int lo = min((int)threadIdx.x * span, N), hi = min(lo + span, N);
  int local = 0;
  for (int i = lo; i < hi; ++i) local += off[i];
  int incl = local;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += wsum[w];
  int run = wbase + incl - local;
  for (int i = lo; i < hi; ++i) { const int c = off[i]; off[i] = run; cur[i] = run; run += c; }
  if (threadIdx.x == kThreads - 1) off[N] = run;
  __syncthreads();
  int* ro = rev_off + b * (N + 1);
  for (int i = threadIdx.x; i <= N; i += kThreads) ro[i] = off[i];
  unsigned int* rs = rev_src + ebase;
  for (int e = threadIdx.x; e < E; e += kThreads) {
    const int m = e / k, j = e - m * k;
    const int t = load_index<I64>(nbr, ebase + e);
    if (t != m) rs[atomicAdd(&cur[t], 1)] = (unsigned)m | ((unsigned)j << 24);
  }
  __syncthreads();
  // fixed summation order: sort every (short) list by source row
  for (int n = threadIdx.x; n < N; n += kThreads) {
    const int s0 = off[n], s1 = off[n + 1];
    for (int i = s0 + 1; i < s1; ++i) {
      const unsigned v = rs[i];
      int p = i - 1;
      while (p >= s0 && rs[p] > v) { rs[p + 1] = rs[p]; --p; }
      rs[p + 1] = v;
    }
  }
}

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

template <bool I64, int KN, int D>
__global__ void __launch_bounds__(kThreads)
mr_aggregate_bwd_gather_kernel(const float* __restrict__ g, const uint8_t* __restrict__ argmax,
                               const void* __restrict__ nbr, const int* __restrict__ rev_off,
                               const unsigned int* __restrict__ rev_src, float* __restrict__ grad_x, int N, int C,
                               int cv_shift, int ips_shift, int total_iters) {
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  // per stage: two 16-byte planes (the thread's 8 interleaved grad_out floats) and one 4-byte plane (its argmax word)
  constexpr int kStageBytes = kThreads * 36;
  const uint32_t smem0 = static_cast<uint32_t>(__cvta_generic_to_shared(pipe_smem));
  const int cmask = (1 << cv_shift) - 1;
  const int c = (threadIdx.x & cmask) * 4;

  auto row_of = [&](int w, int& b, int& n) {
    b = w >> ips_shift;
    const int chunk = w & ((1 << ips_shift) - 1);
    n = (chunk * kThreads + (int)threadIdx.x) >> cv_shift;
  };
  auto issue = [&](int w, int s) {
    int b, n;
    row_of(w, b, n);
    const long long rowg = (long long)b * N + n;
    const uint32_t st = smem0 + s * kStageBytes;
    const float* gp = g + rowg * 2 * C + 2 * c;
    cp_async16(st + threadIdx.x * 16, gp);
    cp_async16(st + kThreads * 16 + threadIdx.x * 16, gp + 4);
    cp_async4(st + kThreads * 32 + threadIdx.x * 4, argmax + rowg * C + c);
  };
  struct Meta { int ids[KN]; int o0, o1; };
  auto load_meta = [&](int w, Meta& mt) {
    int b, n;
    row_of(w, b, n);
    const long long rowg = (long long)b * N + n;
#pragma unroll
    for (int j = 0; j < KN; ++j) mt.ids[j] = load_index<I64>(nbr, rowg * KN + j);
    const int* ro = rev_off + (long long)b * (N + 1) + n;
    mt.o0 = __ldg(ro);
    mt.o1 = __ldg(ro + 1);
  };

  const int w0 = blockIdx.x, stride = gridDim.x;
#pragma unroll
  for (int s = 0; s < D; ++s) {
    if (w0 + s * stride < total_iters) issue(w0 + s * stride, s);
    cp_async_commit();
  }
  Meta cur, nxt;
  if (w0 < total_iters) load_meta(w0, cur);
  nxt = cur;

  int s = 0;
  for (int w = w0; w < total_iters; w += stride) {
    if (w + stride < total_iters) load_meta(w + stride, nxt);  // consumed next iteration
    cp_async_wait<D - 1>();
    const unsigned char* st = pipe_smem + s * kStageBytes;
    const float4 ga = *reinterpret_cast<const float4*>(st + threadIdx.x * 16);
    const float4 gb = *reinterpret_cast<const float4*>(st + kThreads * 16 + threadIdx.x * 16);
    const unsigned int am = *reinterpret_cast<const unsigned int*>(st + kThreads * 32 + threadIdx.x * 4);
    int b, n;
    row_of(w, b, n);
    const float g0[4] = {ga.x, ga.z, gb.x, gb.z};
    const float g1[4] = {ga.y, ga.w, gb.y, gb.w};
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int a = (am >> (8 * e)) & 0xff;
      int win = cur.ids[KN - 1];
#pragma unroll
      for (int j = KN - 2; j >= 0; --j) win = (a == j) ? cur.ids[j] : win;
      r[e] = (win == n) ? g0[e] : g0[e] - g1[e];
    }
    // in-edges: (m, j) pairs whose slot j of row m is this row; two in flight per trip
    const long long segrow = (long long)b * N;
    const unsigned int* rs = rev_src + segrow * KN;
    for (int e0 = cur.o0; e0 < cur.o1; e0 += 2) {
      const bool two = e0 + 1 < cur.o1;
      const unsigned ent0 = __ldg(rs + e0);
      const unsigned ent1 = two ? __ldg(rs + e0 + 1) : ent0;
      const long long m0 = segrow + (ent0 & 0xffffffu), m1 = segrow + (ent1 & 0xffffffu);
      const unsigned am0 = __ldg(reinterpret_cast<const unsigned int*>(argmax + m0 * C + c));
      const unsigned am1 = __ldg(reinterpret_cast<const unsigned int*>(argmax + m1 * C + c));
      float gm0[8], gm1[8];
      Pack8<float>::load(g + m0 * 2 * C + 2 * c, gm0);
      Pack8<float>::load(g + m1 * 2 * C + 2 * c, gm1);
      const unsigned j0 = ent0 >> 24, j1 = ent1 >> 24;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        r[e] += (((am0 >> (8 * e)) & 0xff) == j0) ? gm0[2 * e + 1] : 0.f;
        r[e] += (two && ((am1 >> (8 * e)) & 0xff) == j1) ? gm1[2 * e + 1] : 0.f;
      }
    }
    *reinterpret_cast<float4*>(grad_x + (segrow + n) * C + c) = make_float4(r[0], r[1], r[2], r[3]);
    const int wn = w + D * stride;
    if (wn < total_iters) issue(wn, s);
    cp_async_commit();
    cur = nxt;
    if (++s == D) s = 0;
  }
  cp_async_wait<0>();
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

// row form with compile-time k: U (row, 4-channel) items per thread and iteration, all ids then all gathers in flight
template <typename T, int VEC, bool I64, int KN, int U>
__global__ void __launch_bounds__(kThreads)
gather_fwd_row_kernel(const T* __restrict__ src, const void* __restrict__ idx, T* __restrict__ out, unsigned rows,
                      unsigned cv, int N, int M, int C) {
  const unsigned items = rows * cv;
  const unsigned step = gridDim.x * kThreads;
  for (unsigned it0 = blockIdx.x * kThreads + threadIdx.x; it0 < items; it0 += step * U) {
    unsigned row[U], c[U];
    int nb[U][KN];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned it = it0 + u * step;
      ok[u] = it < items;
      const unsigned itc = ok[u] ? it : items - 1;
      row[u] = itc / cv;
      c[u] = (itc - row[u] * cv) * VEC;
#pragma unroll
      for (int j = 0; j < KN; ++j) nb[u][j] = load_index<I64>(idx, (long long)row[u] * KN + j);
    }
    float v[U][KN][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long seg = (long long)(row[u] / (unsigned)N) * M;
#pragma unroll
      for (int j = 0; j < KN; ++j) Pack<T, VEC>::load(src + (seg + nb[u][j]) * (long long)C + c[u], v[u][j]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int j = 0; j < KN; ++j) Pack<T, VEC>::store(out + ((long long)row[u] * KN + j) * C + c[u], v[u][j]);
    }
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
// neighbour sum: out[b][n][c] = sum_j src[b][idx[b][n][j]][c]  (GINConv2d: batched_index_select + torch.sum over
// the neighbour axis, torch_vertex.py:84-88) without the (B, C, N, k) intermediate; backward routes grad_out[b][n]
// to its k neighbour rows with red.v4 into a zero-filled grad_src.  Summation order j = 0..k-1.
// ------------------------------------------------------------------------------------
template <typename T, int VEC, bool I64, int U>
__global__ void __launch_bounds__(kThreads)
neighbor_sum_fwd_kernel(const T* __restrict__ src, const void* __restrict__ idx, T* __restrict__ out, long long rows,
                        int N, int M, int C, int k) {
  const long long cv = C / VEC;
  const long long items = rows * cv;
  const long long step = (long long)gridDim.x * kThreads;
  for (long long it0 = blockIdx.x * (long long)kThreads + threadIdx.x; it0 < items; it0 += step * U) {
    long long row[U];
    int c[U];
    bool ok[U];
    float acc[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long it = it0 + u * step;
      ok[u] = it < items;
      const long long itc = ok[u] ? it : items - 1;
      row[u] = itc / cv;
      c[u] = static_cast<int>(itc - row[u] * cv) * VEC;
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[u][e] = 0.f;
    }
    for (int j0 = 0; j0 < k; j0 += 4) {  // four neighbours x U items in flight
      float v[U][4][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long seg = (row[u] / N) * M;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = min(j0 + jj, k - 1);
          const int nb = load_index<I64>(idx, row[u] * k + j);
          Pack<T, VEC>::load(src + (seg + nb) * (long long)C + c[u], v[u][jj]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          if (j0 + jj < k) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[u][e] += v[u][jj][e];
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ok[u]) Pack<T, VEC>::store(out + row[u] * C + c[u], acc[u]);
    }
  }
}

template <typename T, int VEC, bool I64>
__global__ void __launch_bounds__(kThreads)
neighbor_sum_bwd_kernel(const T* __restrict__ g, const void* __restrict__ idx, T* __restrict__ grad_src, long long rows,
                        int N, int M, int C, int k) {
  const long long cv = C / VEC;
  const long long items = rows * cv;
  for (long long it = blockIdx.x * (long long)kThreads + threadIdx.x; it < items;
       it += (long long)gridDim.x * kThreads) {
    const long long row = it / cv;
    const int c = static_cast<int>(it - row * cv) * VEC;
    const long long seg = (row / N) * M;
    float gv[VEC];
    Pack<T, VEC>::load(g + row * C + c, gv);
    for (int j = 0; j < k; ++j) {
      const int nb = load_index<I64>(idx, row * k + j);
      Pack<T, VEC>::red_add(grad_src + (seg + nb) * (long long)C + c, gv);
    }
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

// Row form of the EdgeConv features for graphs whose centre is the row itself (the DenseDilatedKnnGraph case) and a
// compile-time k: a thread owns U (row, 4-channel) items per iteration, reads each centre slice once (the edge form
// above re-reads it for every neighbour), and has the U * KN neighbour ids and then the U * (KN + 1) 16-byte loads in
// flight before the first store - the same latency-hiding recipe that took K2 from 61 % to 80 % of the HBM rate.
// 32-bit index arithmetic (the host checks the item count).
template <typename T, int VEC, bool I64, int KN, int U>
__global__ void __launch_bounds__(kThreads)
edge_gather_fwd_row_kernel(const T* __restrict__ x, const T* __restrict__ src, const void* __restrict__ nbr,
                           T* __restrict__ out, unsigned rows, unsigned cv, int N, int M, int C) {
  const unsigned items = rows * cv;
  const unsigned step = gridDim.x * kThreads;
  for (unsigned it0 = blockIdx.x * kThreads + threadIdx.x; it0 < items; it0 += step * U) {
    unsigned row[U], c[U];
    int nb[U][KN];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned it = it0 + u * step;
      ok[u] = it < items;
      const unsigned itc = ok[u] ? it : items - 1;
      row[u] = itc / cv;
      c[u] = (itc - row[u] * cv) * VEC;
#pragma unroll
      for (int j = 0; j < KN; ++j) nb[u][j] = load_index<I64>(nbr, (long long)row[u] * KN + j);
    }
    float xi[U][VEC], xj[U][KN][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long seg = (long long)(row[u] / (unsigned)N) * M;
      Pack<T, VEC>::load(x + (long long)row[u] * C + c[u], xi[u]);
#pragma unroll
      for (int j = 0; j < KN; ++j) Pack<T, VEC>::load(src + (seg + nb[u][j]) * (long long)C + c[u], xj[u][j]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int j = 0; j < KN; ++j) {
        float d[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) d[e] = xj[u][j][e] - xi[u][e];
        T* o = out + ((long long)row[u] * KN + j) * 2 * C + c[u];
        Pack<T, VEC>::store(o, xi[u]);
        Pack<T, VEC>::store(o + C, d);
      }
    }
  }
}

// One-pass backward of the same case: a thread owns a (row, 4-channel) item, reads its 2 * KN slices of grad_out once,
// keeps the dense part sum_j (ga - gb) in registers and routes every gb slice to its neighbour row with red.v4 -
// including the dense part itself, into a grad_x that the host zero-fills first (one extra write pass of N * C,
// against the second read pass of k * N * C that the dense + scatter pair of kernels needs).
template <typename T, int VEC, bool I64, int KN>
__global__ void __launch_bounds__(kThreads)
edge_gather_bwd_row_kernel(const T* __restrict__ g, const void* __restrict__ nbr, T* __restrict__ grad_x, unsigned rows,
                           unsigned cv, int N, int C) {
  const unsigned items = rows * cv;
  for (unsigned it = blockIdx.x * kThreads + threadIdx.x; it < items; it += gridDim.x * kThreads) {
    const unsigned row = it / cv;
    const unsigned c = (it - row * cv) * VEC;
    const long long seg = (long long)(row / (unsigned)N) * N;
    int nb[KN];
#pragma unroll
    for (int j = 0; j < KN; ++j) nb[j] = load_index<I64>(nbr, (long long)row * KN + j);
    float ga[KN][VEC], gb[KN][VEC];
#pragma unroll
    for (int j = 0; j < KN; ++j) {
      const T* p = g + ((long long)row * KN + j) * 2 * C + c;
      Pack<T, VEC>::load(p, ga[j]);
      Pack<T, VEC>::load(p + C, gb[j]);
    }
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
    for (int j = 0; j < KN; ++j) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] += ga[j][e] - gb[j][e];
      Pack<T, VEC>::red_add(grad_x + (seg + nb[j]) * (long long)C + c, gb[j]);
    }
    Pack<T, VEC>::red_add(grad_x + (long long)row * C + c, acc);
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

// compile-time k, U items per thread and iteration: all U * KN loads are issued before the first compare
template <typename T, int KN, int U>
__global__ void __launch_bounds__(kThreads)
max_over_k_fwd_row_kernel(const T* __restrict__ h, T* __restrict__ out, uint8_t* __restrict__ argmax, unsigned rows,
                          unsigned cv, int C) {
  constexpr int VEC = 4;
  const unsigned items = rows * cv;
  const unsigned step = gridDim.x * kThreads;
  for (unsigned it0 = blockIdx.x * kThreads + threadIdx.x; it0 < items; it0 += step * U) {
    float v[U][KN][VEC];
    unsigned row[U], c[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned it = it0 + u * step;
      ok[u] = it < items;
      const unsigned itc = ok[u] ? it : items - 1;
      row[u] = itc / cv;
      c[u] = (itc - row[u] * cv) * VEC;
#pragma unroll
      for (int j = 0; j < KN; ++j) Pack<T, VEC>::load(h + ((long long)row[u] * KN + j) * C + c[u], v[u][j]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      float best[VEC];
      int arg[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) { best[e] = -INFINITY; arg[e] = 0; }
#pragma unroll
      for (int j = 0; j < KN; ++j) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (v[u][j][e] > best[e] || v[u][j][e] != v[u][j][e]) {
            if (!(best[e] != best[e])) { best[e] = v[u][j][e]; arg[e] = j; }
          }
        }
      }
      Pack<T, VEC>::store(out + (long long)row[u] * C + c[u], best);
      if (argmax != nullptr) {
        *reinterpret_cast<uchar4*>(argmax + (long long)row[u] * C + c[u]) =
            make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
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

// K2 dispatch (option OPT_MR_FWD_FORM): pipelined persistent kernel where it applies (k == 3, centre == row, whole
// 256-item chunks), else the register-prefetch kernel (4-channel packs, C/4 a power of two), else the generic one.
template <typename T>
int launch_mr_aggregate_fwd(const void* x, const void* y, const void* nbr, const void* ctr, int idx_is_i64, void* out,
                            uint8_t* argmax, int B, int N, int M, int C, int k, cudaStream_t s) {
  const T* xs = static_cast<const T*>(x);
  const T* src = y ? static_cast<const T*>(y) : xs;
  const long long rows = (long long)B * N;
  const bool al = aligned16(x) && aligned16(src) && aligned16(out) && (argmax == nullptr || ((uintptr_t)argmax & 3) == 0);
  const int form = option(OPT_MR_FWD_FORM);
  return dispatch_vec_idx(C, al, idx_is_i64, [&](auto vec, auto i64) {
    constexpr int VEC = decltype(vec)::value;
    constexpr bool I64 = decltype(i64)::value;
    const int grid = grid_for(rows * (C / VEC), kThreads, 8);
    if constexpr (VEC == 4) {
      // pipelined form: one item = 16 bytes of channels
      constexpr int V = Item16<T>::V;
      if (!ctr && form >= 2 && k == 3 && C % V == 0 && aligned32(out) && (argmax == nullptr || ((uintptr_t)argmax & 7) == 0)) {
        const int cvp = C / V;
        const long long items_per_seg = (long long)N * cvp;
        const long long ips = items_per_seg / kThreads;
        if ((cvp & (cvp - 1)) == 0 && cvp <= kThreads && items_per_seg % kThreads == 0 && (ips & (ips - 1)) == 0 && ips >= 1 &&
            (long long)B * ips < 0x7fffffffLL) {
          int cv_shift = 0, ips_shift = 0;
          while ((1 << cv_shift) < cvp) ++cv_shift;
          while ((1LL << ips_shift) < ips) ++ips_shift;
          const int total = (int)(B * ips);
          constexpr int D = 4;
          const size_t smem = (size_t)D * 4 * kThreads * 16;
          static DeviceOnce once;
          if (once.pending()) {
            cudaError_t e = cudaFuncSetAttribute(mr_aggregate_fwd_pipe_kernel<T, I64, 3, D>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(mr_fwd_pipe): %s", cudaGetErrorString(e)); return (int)e; }
            once.mark();
          }
          int ctas = num_sms() * 3;  // 64 KB of slots per CTA: three CTAs per SM
          if (ctas > total) ctas = total;
          mr_aggregate_fwd_pipe_kernel<T, I64, 3, D><<<ctas, kThreads, smem, s>>>(xs, src, nbr, static_cast<T*>(out), argmax, N, M,
                                                                                C, cv_shift, ips_shift, total);
          return check_launch("mr_aggregate_fwd_pipe");
        }
      }
      const int cv = C / 4;
      if (!ctr && form >= 1 && (cv & (cv - 1)) == 0 && cv <= kThreads && B <= 65535 && aligned32(out)) {
        int cv_shift = 0;
        while ((1 << cv_shift) < cv) ++cv_shift;
        const int rpb = kThreads >> cv_shift;
        constexpr int U = 4;  // items in flight per thread
        const int passes = (N + rpb * U - 1) / (rpb * U);
        int gx = (num_sms() * 8 + B - 1) / B;  // about 8 resident CTAs per SM across the whole batch
        if (gx > passes) gx = passes;
        if (gx < 1) gx = 1;
        mr_aggregate_fwd_fast_kernel<T, I64, U><<<dim3(gx, B), kThreads, 0, s>>>(xs, src, nbr, static_cast<T*>(out), argmax, N, M, C,
                                                                             k, cv_shift);
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

size_t mr_bwd_workspace_bytes(int B, int N, int k) {
  // reverse CSR of every segment's graph: offsets (B, N + 1) int32 and sources (B, N * k) uint32
  const size_t off = ((size_t)B * (N + 1) * 4 + 255) / 256 * 256;
  return off + (size_t)B * N * k * 4 + 256;
}

// K3 dispatch (option OPT_MR_BWD_FORM): 2 (default) / 1 cluster kernel without / with the device-scope fence,
// 3 deterministic gather over the reverse graph (fp32, k == 3, needs the workspace), 0 dense + scatter pair.
// Graphs with explicit centre ids or a separate key set always take the pair.
template <typename T>
int launch_mr_aggregate_bwd(const void* g, const uint8_t* argmax, const void* nbr, const void* ctr, int idx_is_i64,
                            void* grad_x, void* grad_y, int B, int N, int M, int C, int k, void* workspace,
                            size_t workspace_bytes, cudaStream_t s) {
  const long long rows = (long long)B * N;
  T* gx = static_cast<T*>(grad_x);
  T* gsrc = grad_y ? static_cast<T*>(grad_y) : gx;
  const bool al = aligned16(g) && aligned16(grad_x) && aligned16(gsrc) && (((uintptr_t)argmax & 3) == 0);
  if (grad_y) {
    cudaError_t e = cudaMemsetAsync(grad_y, 0, (size_t)B * M * C * sizeof(T), s);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync(grad_y): %s", cudaGetErrorString(e)); return (int)e; }
  }
  const int form = option(OPT_MR_BWD_FORM);
  return dispatch_vec_idx(C, al, idx_is_i64, [&](auto vec, auto i64) {
    constexpr int VEC = decltype(vec)::value;
    constexpr bool I64 = decltype(i64)::value;
    const int grid = grid_for(rows * (C / VEC), kThreads, 8);
    const T* gs = static_cast<const T*>(g);
    const bool self_skip = (ctr == nullptr) && (grad_y == nullptr);
    if constexpr (VEC == 4 && std::is_same<T, float>::value) {
      // gather form over the reverse graph in a workspace
      const int cv = C / 4;
      const long long items_per_seg = (long long)N * cv;
      const long long ips = items_per_seg / kThreads;
      if (self_skip && form == 3 && k == 3 && workspace != nullptr &&
          workspace_bytes >= mr_bwd_workspace_bytes(B, N, k) && (cv & (cv - 1)) == 0 && cv <= kThreads &&
          items_per_seg % kThreads == 0 && (ips & (ips - 1)) == 0 && ips >= 1 && (long long)B * ips < 0x7fffffffLL &&
          N <= 8192 && aligned32(g)) {
        int cv_shift = 0, ips_shift = 0;
        while ((1 << cv_shift) < cv) ++cv_shift;
        while ((1LL << ips_shift) < ips) ++ips_shift;
        char* wbase = reinterpret_cast<char*>(((uintptr_t)workspace + 255) / 256 * 256);
        int* rev_off = reinterpret_cast<int*>(wbase);
        unsigned int* rev_src = reinterpret_cast<unsigned int*>(wbase + ((size_t)B * (N + 1) * 4 + 255) / 256 * 256);
        const size_t smem_rev = (size_t)(2 * N + 2) * sizeof(int);
        constexpr int D = 4;
        const size_t smem_pipe = (size_t)D * kThreads * 36;
        static DeviceOnce once;
        if (once.pending()) {
          cudaError_t e = cudaFuncSetAttribute(mr_bwd_build_reverse_kernel<I64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
          if (e == cudaSuccess) e = cudaFuncSetAttribute(mr_aggregate_bwd_gather_kernel<I64, 3, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pipe);
          if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(mr_bwd_gather): %s", cudaGetErrorString(e)); return (int)e; }
          once.mark();
        }
        mr_bwd_build_reverse_kernel<I64><<<B, kThreads, smem_rev, s>>>(nbr, rev_off, rev_src, N, k);
        const int total = (int)(B * ips);
        int ctas = num_sms() * 6;  // 36 KB of slots per CTA
        if (ctas > total) ctas = total;
        mr_aggregate_bwd_gather_kernel<I64, 3, D><<<ctas, kThreads, smem_pipe, s>>>(
            reinterpret_cast<const float*>(gs), argmax, nbr, rev_off, rev_src, reinterpret_cast<float*>(gx), N, C, cv_shift,
            ips_shift, total);
        return check_launch("mr_aggregate_bwd_gather");
      }
    }
    if constexpr (VEC == 4) {
      if (self_skip && form >= 1) {  // default: cluster form with bulk staging
        bool launched = false;
        const int rc = launch_mr_bwd_cluster<T, I64>(gs, argmax, nbr, gx, B, N, C, k, form == 1, s, &launched);
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
    if constexpr (VEC == 4) {
      const long long rows = (long long)B * N;
      const bool row_form = option(OPT_GATHER_ROW) != 0;
      if (row_form && k >= 2 && k <= 4 && rows * (C / VEC) < 0x7fffffffLL) {
        const unsigned cv = C / VEC;
        const int g2 = grid_for((rows * cv + 1) / 2, kThreads, 8);
#define GRAFP_GATHER_ROW(KN_) gather_fwd_row_kernel<T, VEC, I64, KN_, 2><<<g2, kThreads, 0, s>>>(static_cast<const T*>(src), idx, static_cast<T*>(out), (unsigned)rows, cv, N, M, C)
        if (k == 2) GRAFP_GATHER_ROW(2); else if (k == 3) GRAFP_GATHER_ROW(3); else GRAFP_GATHER_ROW(4);
#undef GRAFP_GATHER_ROW
        return check_launch("gather_fwd_row");
      }
    }
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
int launch_neighbor_sum_fwd(const void* src, const void* idx, int idx_is_i64, void* out, int B, int N, int M, int C, int k,
                            cudaStream_t s) {
  const long long rows = (long long)B * N;
  return dispatch_vec_idx(C, aligned16(src) && aligned16(out), idx_is_i64, [&](auto vec, auto i64) {
    constexpr int VEC = decltype(vec)::value;
    constexpr bool I64 = decltype(i64)::value;
    const int grid = grid_for((rows * (C / VEC) + 1) / 2, kThreads, 8);
    neighbor_sum_fwd_kernel<T, VEC, I64, 2><<<grid, kThreads, 0, s>>>(static_cast<const T*>(src), idx, static_cast<T*>(out),
                                                                      rows, N, M, C, k);
    return check_launch("neighbor_sum_fwd");
  });
}

template <typename T>
int launch_neighbor_sum_bwd(const void* g, const void* idx, int idx_is_i64, void* grad_src, int B, int N, int M, int C,
                            int k, cudaStream_t s) {
  const long long rows = (long long)B * N;
  cudaError_t e = cudaMemsetAsync(grad_src, 0, (size_t)B * M * C * sizeof(T), s);
  if (e != cudaSuccess) { set_error("cudaMemsetAsync(grad_src): %s", cudaGetErrorString(e)); return (int)e; }
  return dispatch_vec_idx(C, aligned16(g) && aligned16(grad_src), idx_is_i64, [&](auto vec, auto i64) {
    constexpr int VEC = decltype(vec)::value;
    constexpr bool I64 = decltype(i64)::value;
    const int grid = grid_for(rows * (C / VEC), kThreads, 8);
    neighbor_sum_bwd_kernel<T, VEC, I64><<<grid, kThreads, 0, s>>>(static_cast<const T*>(g), idx, static_cast<T*>(grad_src),
                                                                   rows, N, M, C, k);
    return check_launch("neighbor_sum_bwd");
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
    if constexpr (VEC == 4) {
      const long long rows = (long long)B * N;
      const bool row_form = option(OPT_EDGE_ROW) != 0;
      if (!ctr && row_form && k >= 2 && k <= 4 && rows * (C / VEC) < 0x7fffffffLL) {
        const unsigned cv = C / VEC;
        const int g2 = grid_for((rows * cv + 1) / 2, kThreads, 8);
#define GRAFP_EDGE_ROW(KN_) edge_gather_fwd_row_kernel<T, VEC, I64, KN_, 2><<<g2, kThreads, 0, s>>>(xs, src, nbr, static_cast<T*>(out), (unsigned)rows, cv, N, M, C)
        if (k == 2) GRAFP_EDGE_ROW(2); else if (k == 3) GRAFP_EDGE_ROW(3); else GRAFP_EDGE_ROW(4);
#undef GRAFP_EDGE_ROW
        return check_launch("edge_gather_fwd_row");
      }
    }
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
    if constexpr (VEC == 4) {
      // one-pass form (default; GRAFP_EDGE_BWD_ROW=0 selects the dense + scatter pair): 229-241 us against 280-290 us
      const bool row_form = option(OPT_EDGE_BWD_ROW) != 0;
      if (!ctr && !grad_y && row_form && k >= 2 && k <= 4 && rows * (C / VEC) < 0x7fffffffLL) {
        cudaError_t e2 = cudaMemsetAsync(grad_x, 0, (size_t)B * N * C * sizeof(T), s);
        if (e2 != cudaSuccess) { set_error("cudaMemsetAsync(edge grad_x): %s", cudaGetErrorString(e2)); return (int)e2; }
        const unsigned cv = C / VEC;
        const int g1 = grid_for(rows * cv, kThreads, 8);
#define GRAFP_EDGE_BWD_ROW(KN_) edge_gather_bwd_row_kernel<T, VEC, I64, KN_><<<g1, kThreads, 0, s>>>(gs, nbr, gx, (unsigned)rows, cv, N, C)
        if (k == 2) GRAFP_EDGE_BWD_ROW(2); else if (k == 3) GRAFP_EDGE_BWD_ROW(3); else GRAFP_EDGE_BWD_ROW(4);
#undef GRAFP_EDGE_BWD_ROW
        return check_launch("edge_gather_bwd_row");
      }
    }
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
  const bool row_form = option(OPT_MAXK_ROW) != 0;
  if (vec4 && row_form && k >= 2 && k <= 4 && rows * (C / 4) < 0x7fffffffLL) {
    const unsigned cv = C / 4;
    const int g2 = grid_for((rows * cv + 1) / 2, kThreads, 8);
#define GRAFP_MAXK_ROW(KN_) max_over_k_fwd_row_kernel<T, KN_, 2><<<g2, kThreads, 0, s>>>(static_cast<const T*>(h), static_cast<T*>(out), argmax, (unsigned)rows, cv, C)
    if (k == 2) GRAFP_MAXK_ROW(2); else if (k == 3) GRAFP_MAXK_ROW(3); else GRAFP_MAXK_ROW(4);
#undef GRAFP_MAXK_ROW
    return check_launch("max_over_k_fwd_row");
  }
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
                                          int, int, int, int, int, void*, size_t, cudaStream_t);                     \
  template int launch_gather_fwd<T>(const void*, const void*, int, void*, int, int, int, int, int, cudaStream_t);    \
  template int launch_gather_bwd<T>(const void*, const void*, int, void*, int, int, int, int, int, cudaStream_t);    \
  template int launch_neighbor_sum_fwd<T>(const void*, const void*, int, void*, int, int, int, int, int, cudaStream_t); \
  template int launch_neighbor_sum_bwd<T>(const void*, const void*, int, void*, int, int, int, int, int, cudaStream_t); \
  template int launch_edge_gather_fwd<T>(const void*, const void*, const void*, const void*, int, void*, int, int,   \
                                         int, int, int, cudaStream_t);                                               \
  template int launch_edge_gather_bwd<T>(const void*, const void*, const void*, int, void*, void*, int, int, int,    \
                                         int, int, cudaStream_t);                                                    \
  template int launch_max_over_k_fwd<T>(const void*, void*, uint8_t*, int, int, int, int, cudaStream_t);             \
  template int launch_max_over_k_bwd<T>(const void*, const uint8_t*, void*, int, int, int, int, cudaStream_t);

GRAFP_INSTANTIATE(float)
GRAFP_INSTANTIATE(__nv_bfloat16)

// ------------------------------------------------------------------------------------
// index range check (the reference's advanced indexing raises for ids outside [0, M); see grafp_check_index)
// ------------------------------------------------------------------------------------
template <bool I64>
__global__ void __launch_bounds__(kThreads)
check_index_kernel(const void* __restrict__ idx, long long count, int limit, int* __restrict__ bad) {
  int local = 0;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < count; i += (long long)gridDim.x * kThreads) {
    long long v;
    if constexpr (I64) v = __ldg(reinterpret_cast<const long long*>(idx) + i);
    else v = __ldg(reinterpret_cast<const int*>(idx) + i);
    local += (v < 0 || v >= limit) ? 1 : 0;
  }
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && local != 0) atomicAdd(bad, local);
}

int launch_check_index(const void* idx, int idx_is_i64, long long count, int limit, int* bad_count, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(bad_count, 0, sizeof(int), s);
  if (e != cudaSuccess) { set_error("grafp_check_index: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
  const int grid = grid_for(count, kThreads, 4);
  if (idx_is_i64) check_index_kernel<true><<<grid, kThreads, 0, s>>>(idx, count, limit, bad_count);
  else check_index_kernel<false><<<grid, kThreads, 0, s>>>(idx, count, limit, bad_count);
  return check_launch("check_index");
}

}  // namespace grafp
