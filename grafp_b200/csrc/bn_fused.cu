// Train-mode BatchNorm over node rows, fused with the ReLU and / or the residual add that follow it
// in the Grapher / FFN blocks (reference: torch_vertex.py:152-162,183-194 fc1 / fc2 = Conv2d + BatchNorm2d,
// graph_encoder.py:45-67 FFN, torch_nn.py:52-64 BasicConv = Conv2d + BatchNorm2d + ReLU).  SURVEY 8(f) row 2.
//
// Activations are rows (R = B*N, C), fp32 or bf16 (statistics and parameters always fp32); a thread owns 16 bytes
// of channels (V = 4 fp32 / 8 bf16), C % V == 0 and C / V a power of two.  HBM-bound streaming, TWO launches each way:
//   forward : statistics pass (read x) + apply pass (read x [, residual], write y)            -> 3-4 tensor passes
//             (PyTorch: BN 3 + ReLU 2 or add 3)
//   backward: reduction pass (read dy, x) + apply pass (read dy, x, write dx)                  -> 5 tensor passes
//             (PyTorch: ReLU backward 3 + BN backward 5); the ReLU mask is recomputed from x, so neither the
//             BN output nor the ReLU output is kept for backward.
// There are no finalize kernels (round 1 had three of them, 9-17 us of pure latency each, 126 calls per step): every
// block of the first pass folds its per-channel partial sums into 2C doubles with red.f64 (a few hundred
// reductions per address), and the second pass turns those into mean / invstd (or the two gradient sums) in its
// prologue; block-0 threads also write the saved statistics, the running statistics and dweight / dbias.
// Statistics are accumulated around a common shift (row 0 of the tensor) so E[x^2] - E[x]^2 never cancels
// catastrophically, in fp32 per thread and block, in double across blocks.
// With OPT_BN_REVERSE the second pass walks the rows back to front: the first pass leaves the tail of the tensor in
// the 126 MB L2, so the second pass starts on L2 hits instead of evicting them before it gets there.
#include <type_traits>

#include "common.cuh"

namespace grafp {
namespace {

constexpr int kBnThreads = 256;
constexpr int kBnMaxTpr = 32;  // threads per row chunk: a block covers <= 32 * V channels, which bounds the number of
                               // red.f64 per block (blocks * channels-per-block * 2 per pass)

inline double* bn_sums_of(void* workspace) { return reinterpret_cast<double*>(((uintptr_t)workspace + 255) / 256 * 256); }

// L2 eviction policies for the two-pass kernels: what pass 2 re-reads first is loaded "evict last" in pass 1, everything
// that will not be touched again "evict first".
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ldg16_hint(const void* p, unsigned long long policy) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
  return v;
}

template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int V = 4;
  static __device__ __forceinline__ float scalar(const float* p) { return __ldg(p); }
  // raw 16-byte loads: the unrolled loops keep the packed registers in flight and unpack at use (for bf16 the unpacked
  // form is twice the registers: 174 per thread and one block per SM before this)
  static __device__ __forceinline__ uint4 load_raw(const float* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ uint4 load_raw_hint(const float* p, unsigned long long policy) { return ldg16_hint(p, policy); }
  static __device__ __forceinline__ void unpack(const uint4& t, float (&v)[4]) {
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
  static __device__ __forceinline__ void load_hint(const float* p, float (&v)[4], unsigned long long policy) {
    const uint4 t = ldg16_hint(p, policy);
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int V = 8;
  static __device__ __forceinline__ float scalar(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ uint4 load_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ uint4 load_raw_hint(const __nv_bfloat16* p, unsigned long long policy) { return ldg16_hint(p, policy); }
  static __device__ __forceinline__ void unpack(const uint4& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  static __device__ __forceinline__ void load_hint(const __nv_bfloat16* p, float (&v)[8], unsigned long long policy) {
    const uint4 t = ldg16_hint(p, policy);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) { Pack8<__nv_bfloat16>::load(p, v); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) { Pack8<__nv_bfloat16>::store(p, v); }
};

struct BnGeom {
  int cv;      // 16-byte packs per row
  int tpr;     // threads per row inside a block (power of two <= kBnMaxTpr)
  int rpp;     // rows per block pass = 256 / tpr
  int ctiles;  // channel tiles (grid.y) = cv / tpr
  int gx;      // row blocks (grid.x)
};

inline bool bn_geometry(long long R, int C, int V, BnGeom* g) {
  if (C < V || C % V != 0 || R < 2) return false;
  const int cv = C / V;
  if ((cv & (cv - 1)) != 0) return false;
  g->cv = cv;
  g->tpr = cv < kBnMaxTpr ? cv : kBnMaxTpr;
  g->rpp = kBnThreads / g->tpr;
  g->ctiles = cv / g->tpr;
  long long need = (R + g->rpp - 1) / g->rpp;
  long long cap = (long long)num_sms() * 4 / g->ctiles;
  if (cap < 1) cap = 1;
  g->gx = (int)(need < cap ? need : cap);
  return true;
}

__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_f64_cg(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// block reduction over the row lanes that share a channel pack: thread (rl, tc) holds NV values; after the call the
// threads with rl == 0 hold the sums.  `sm` has NV * 256 floats.
template <int NV>
__device__ __forceinline__ void reduce_row_lanes(float (&v)[NV], float* sm, int tpr_shift) {
  const int rpp = kBnThreads >> tpr_shift;
  const int rl = threadIdx.x >> tpr_shift;
#pragma unroll
  for (int e = 0; e < NV; ++e) sm[e * kBnThreads + threadIdx.x] = v[e];
  __syncthreads();
  for (int half = rpp >> 1; half >= 1; half >>= 1) {
    if (rl < half) {
      const int o = threadIdx.x + (half << tpr_shift);
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        v[e] += sm[e * kBnThreads + o];
        sm[e * kBnThreads + threadIdx.x] = v[e];
      }
    }
    __syncthreads();
  }
}

// ---- forward statistics: sums[c] += sum (x - sh_c), sums[C + c] += sum (x - sh_c)^2, sh = row 0 ----
template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bn_stats_kernel(const T* __restrict__ x, double* __restrict__ sums, long long R, int C, int tpr_shift) {
  constexpr int V = Vec16<T>::V;
  __shared__ float sm[2 * V * kBnThreads];
  const int tpr = 1 << tpr_shift;
  const int rpp = kBnThreads >> tpr_shift;
  const int tc = threadIdx.x & (tpr - 1);
  const int rl = threadIdx.x >> tpr_shift;
  const int c = (blockIdx.y * tpr + tc) * V;
  const T* xp = x + c;
  const long long step = (long long)gridDim.x * rpp;
  long long r = (long long)blockIdx.x * rpp + rl;
  float sh[V], acc[2 * V];
  Vec16<T>::load(xp, sh);
#pragma unroll
  for (int e = 0; e < 2 * V; ++e) acc[e] = 0.f;
  auto add = [&](const float (&v)[V]) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float d = v[e] - sh[e];
      acc[e] += d;
      acc[V + e] = fmaf(d, d, acc[V + e]);
    }
  };
  for (; r + 3 * step < R; r += 4 * step) {  // four independent rows in flight
    float v[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) Vec16<T>::load(xp + (r + u * step) * C, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) add(v[u]);
  }
  for (; r < R; r += step) {
    float v[V];
    Vec16<T>::load(xp + r * C, v);
    add(v);
  }
  reduce_row_lanes<2 * V>(acc, sm, tpr_shift);
  if (rl == 0) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      red_add_f64(sums + c + e, (double)acc[e]);
      red_add_f64(sums + C + c + e, (double)acc[V + e]);
    }
  }
}

// ---- forward apply: y = (x - mean) * a + bias (+ residual) (ReLU), a = weight * invstd ----
// The per-channel coefficients are finished once per block into shared memory (one channel per thread and round: the
// double-precision variance, square root and division are ~100 instructions per channel - done per thread for its own
// V channels they cost more than the ~14 items a thread then streams), the rows stream through packed registers.
template <typename T, bool RELU, bool RES>
__global__ void __launch_bounds__(kBnThreads, 4)
bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ weight,
                const float* __restrict__ bias, const double* __restrict__ sums, T* __restrict__ out,
                float* __restrict__ save_mean, float* __restrict__ save_invstd, float* running_mean, float* running_var,
                const float* __restrict__ conv_bias, long long* num_batches_tracked,
                long long R, int C, int cv, float eps, float momentum, int reverse, int raw_moments) {
  constexpr int V = Vec16<T>::V;
  extern __shared__ float coef[];  // [3][C]: mean, a, bias
  const double inv_rows = 1.0 / (double)R;
  for (int ch = threadIdx.x; ch < C; ch += kBnThreads) {
    // raw_moments: the sums are sum x and sum x^2 (the convolution's epilogue, conv_gemm.cu); else about row 0 of x
    const double sh = raw_moments ? 0.0 : (double)Vec16<T>::scalar(x + ch);
    const double s1 = ld_f64_cg(sums + ch) * inv_rows;
    const double q = ld_f64_cg(sums + C + ch) * inv_rows;
    const double var = fmax(q - s1 * s1, 0.0);
    const float mu = (float)(sh + s1);
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    coef[ch] = mu;
    coef[C + ch] = __ldg(weight + ch) * is;
    coef[2 * C + ch] = __ldg(bias + ch);
    if (blockIdx.x == 0) {  // the saved and the running statistics
      save_mean[ch] = mu;
      save_invstd[ch] = is;
      // conv_bias: the convolution in front ran without its bias (it cancels in x - mean); the running mean sees it
      const float shift = conv_bias != nullptr ? __ldg(conv_bias + ch) : 0.f;
      if (running_mean != nullptr) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (mu + shift);
      if (running_var != nullptr) {
        const double unbiased = var * ((double)R / (double)(R - 1));
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  __syncthreads();
  const long long total = R * cv;                          // 16-byte items
  const long long T_ = (long long)gridDim.x * kBnThreads;  // a multiple of cv: a thread's channel pack never changes
  const long long i0 = (long long)blockIdx.x * kBnThreads + threadIdx.x;
  if (i0 >= total) return;
  const int c = (int)(i0 % cv) * V;
  float mu[V], a[V], bb[V];
#pragma unroll
  for (int e = 0; e < V; ++e) { mu[e] = coef[c + e]; a[e] = coef[C + c + e]; bb[e] = coef[2 * C + c + e]; }
  // (x - mean) first: exact-ish difference, no cancellation against a pre-multiplied shift when |mean| >> std
  auto one = [&](const float (&v)[V], const float (&rv)[V], T* o) {
    float y[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      y[e] = fmaf(v[e] - mu[e], a[e], bb[e]);
      if (RES) y[e] += rv[e];
      if (RELU) y[e] = fmaxf(y[e], 0.f);
    }
    Vec16<T>::store(o, y);
  };
  const long long n = (total - i0 + T_ - 1) / T_;  // items of this thread: i0 + j * T_, j < n
  auto at = [&](long long j) { return (i0 + (reverse ? (n - 1 - j) : j) * T_) * V; };
  long long j = 0;
  for (; j + 3 < n; j += 4) {
    uint4 raw[4], rraw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      raw[u] = Vec16<T>::load_raw(x + at(j + u));
      if (RES) rraw[u] = Vec16<T>::load_raw(res + at(j + u));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float v[V], rv[V];
      Vec16<T>::unpack(raw[u], v);
      if (RES) Vec16<T>::unpack(rraw[u], rv);
      one(v, rv, out + at(j + u));
    }
  }
  for (; j < n; ++j) {
    float v[V], rv[V];
    Vec16<T>::load(x + at(j), v);
    if (RES) Vec16<T>::load(res + at(j), rv);
    one(v, rv, out + at(j));
  }
}

// ---- backward reduction: s1 = sum dz, s2 = sum dz * xhat, s3 = sum (x - mean), dz = dy masked by the recomputed ReLU ----
template <typename T, bool RELU>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ weight,
                     const float* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ invstd,
                     double* __restrict__ sums, long long R, int C, int tpr_shift) {
  constexpr int V = Vec16<T>::V;
  __shared__ float sm[3 * V * kBnThreads];
  const int tpr = 1 << tpr_shift;
  const int rpp = kBnThreads >> tpr_shift;
  const int tc = threadIdx.x & (tpr - 1);
  const int rl = threadIdx.x >> tpr_shift;
  const int c = (blockIdx.y * tpr + tc) * V;
  float mu[V], is[V], a[V], b[V], acc[3 * V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    mu[e] = __ldg(mean + c + e);
    is[e] = __ldg(invstd + c + e);
    a[e] = RELU ? __ldg(weight + c + e) * is[e] : 0.f;  // ReLU mask: (x - mean) * a + b > 0, the forward's expression
    b[e] = RELU ? __ldg(bias + c + e) : 0.f;
    acc[e] = 0.f; acc[V + e] = 0.f; acc[2 * V + e] = 0.f;
  }
  const T* xp = x + c;
  const T* gp = dy + c;
  const long long step = (long long)gridDim.x * rpp;
  auto add = [&](const float (&v)[V], const float (&g)[V]) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float d = v[e] - mu[e];
      float gz = g[e];
      if (RELU) gz = fmaf(d, a[e], b[e]) > 0.f ? gz : 0.f;
      acc[e] += gz;
      acc[V + e] = fmaf(gz, d * is[e], acc[V + e]);
      acc[2 * V + e] += d;
    }
  };
  long long r = (long long)blockIdx.x * rpp + rl;
  for (; r + 3 * step < R; r += 4 * step) {
    float v[4][V], g[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) { Vec16<T>::load(xp + (r + u * step) * C, v[u]); Vec16<T>::load(gp + (r + u * step) * C, g[u]); }
#pragma unroll
    for (int u = 0; u < 4; ++u) add(v[u], g[u]);
  }
  for (; r < R; r += step) {
    float v[V], g[V];
    Vec16<T>::load(xp + r * C, v);
    Vec16<T>::load(gp + r * C, g);
    add(v, g);
  }
  reduce_row_lanes<3 * V>(acc, sm, tpr_shift);
  if (rl == 0) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      red_add_f64(sums + c + e, (double)acc[e]);
      red_add_f64(sums + C + c + e, (double)acc[V + e]);
      red_add_f64(sums + 2 * C + c + e, (double)acc[2 * V + e]);
    }
  }
}

// ---- backward apply: dx = weight * invstd * (dz - s1 / R - xhat * s2 / R) ----
// `colsum` (optional): the per-channel column sums of dx - the bias gradient of the convolution in front of the
// BatchNorm.  Analytically sum_r dx = a (s1 - R c1 - invstd c2 s3) with s3 = sum_r (x - mean): exactly zero in real
// arithmetic, and in floating point the rounding residue of c1 = fl(s1 / R) and of the mean.  That residue is what
// the reference's separate reduction over dx measures too; here it costs one extra sum in the reduction pass.
template <typename T, bool RELU>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ weight,
                    const float* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const double* __restrict__ sums, T* __restrict__ dx, float* __restrict__ dweight,
                    float* __restrict__ dbias, float* __restrict__ colsum, long long R, int C, int cv, int reverse) {
  constexpr int V = Vec16<T>::V;
  const long long total = R * cv;
  const long long T_ = (long long)gridDim.x * kBnThreads;
  const long long i0 = (long long)blockIdx.x * kBnThreads + threadIdx.x;
  const int c = (int)(i0 % cv) * V;
  const double inv_rows = 1.0 / (double)R;
  float mu[V], is[V], a[V], b[V], c1[V], c2[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    mu[e] = __ldg(mean + c + e);
    is[e] = __ldg(invstd + c + e);
    a[e] = __ldg(weight + c + e) * is[e];
    b[e] = RELU ? __ldg(bias + c + e) : 0.f;
    const double s1 = ld_f64_cg(sums + c + e), s2 = ld_f64_cg(sums + C + c + e);
    c1[e] = (float)(s1 * inv_rows);
    c2[e] = (float)(s2 * inv_rows);
    if (i0 < cv) {
      dbias[c + e] = (float)s1;
      dweight[c + e] = (float)s2;
      if (colsum != nullptr) {
        const double s3 = ld_f64_cg(sums + 2 * C + c + e);
        colsum[c + e] = (float)((double)a[e] * ((s1 - (double)R * (double)c1[e]) - (double)is[e] * (double)c2[e] * s3));
      }
    }
  }
  if (i0 >= total) return;
  auto one = [&](const float (&v)[V], const float (&g)[V], T* o) {
    float r[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float d = v[e] - mu[e];
      float gz = g[e];
      if (RELU) gz = fmaf(d, a[e], b[e]) > 0.f ? gz : 0.f;
      r[e] = a[e] * (gz - c1[e] - d * is[e] * c2[e]);
    }
    Vec16<T>::store(o, r);
  };
  const long long n = (total - i0 + T_ - 1) / T_;
  auto at = [&](long long j) { return (i0 + (reverse ? (n - 1 - j) : j) * T_) * V; };
  long long j = 0;
  for (; j + 3 < n; j += 4) {
    float v[4][V], g[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) { Vec16<T>::load(x + at(j + u), v[u]); Vec16<T>::load(dy + at(j + u), g[u]); }
#pragma unroll
    for (int u = 0; u < 4; ++u) one(v[u], g[u], dx + at(j + u));
  }
  for (; j < n; ++j) {
    float v[V], g[V];
    Vec16<T>::load(x + at(j), v);
    Vec16<T>::load(dy + at(j), g);
    one(v, g, dx + at(j));
  }
}

// ------------------------------------------------------------------------------------------------------------
// Single-launch forms: both passes in one cooperative kernel (every block resident), separated by a grid barrier.
// A block keeps its (row block, channel tile) assignment in both passes, so the second pass re-reads exactly the rows
// the block streamed in the first - in reverse order, starting with what is still in L2 - and the launch gap and the
// ramp-down / ramp-up between two kernels disappear (the 134 MB tensors, where cuDNN's single persistent BatchNorm
// kernel beat the two-kernel form).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // cumulative: the block's reductions (ordered before this by the barrier above) become visible
    atomicAdd(counter, 1u);
    while (ld_acquire_u32(counter) < expected) __nanosleep(40);
  }
  __syncthreads();
}

template <typename T, bool RELU, bool RES>
__global__ void __launch_bounds__(kBnThreads, sizeof(T) == 4 ? 4 : 3)
bn_fwd_persistent_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ weight,
                         const float* __restrict__ bias, double* __restrict__ sums, unsigned int* __restrict__ counter,
                         T* __restrict__ out, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                         float* running_mean, float* running_var, const float* __restrict__ conv_bias,
                         long long* num_batches_tracked, long long R, int C, int tpr_shift, float eps,
                         float momentum, int first_reverse, int keep) {
  constexpr int V = Vec16<T>::V;
  __shared__ float sm[2 * V * kBnThreads];
  const int tpr = 1 << tpr_shift;
  const int rpp = kBnThreads >> tpr_shift;
  const int tc = threadIdx.x & (tpr - 1);
  const int rl = threadIdx.x >> tpr_shift;
  const int c = (blockIdx.y * tpr + tc) * V;
  const T* xp = x + c;
  const long long step = (long long)gridDim.x * rpp;
  const long long r0 = (long long)blockIdx.x * rpp + rl;
  float sh[V], acc[2 * V];
  Vec16<T>::load(xp, sh);
#pragma unroll
  for (int e = 0; e < 2 * V; ++e) acc[e] = 0.f;
  auto add = [&](const float (&v)[V]) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float d = v[e] - sh[e];
      acc[e] += d;
      acc[V + e] = fmaf(d, d, acc[V + e]);
    }
  };
  // this thread's rows: r0 + j * step, j < n.  first_reverse: the first pass walks them back to front (the producer of x
  // wrote front to back, so the tail of x is what L2 still holds) and the second pass front to back; else the other
  // way round.  Either way the second pass starts on the rows the first pass read last.
  const long long n = r0 < R ? (R - r0 + step - 1) / step : 0;
  auto row1 = [&](long long j) { return (r0 + (first_reverse ? (n - 1 - j) : j) * step) * C; };
  auto row2 = [&](long long j) { return (r0 + (first_reverse ? j : (n - 1 - j)) * step) * C; };
  // L2: the last `keep` rows of pass 1 are what pass 2 reads first - keep them; the rest streams through
  const unsigned long long pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  const long long keep_from = keep >= 0 ? (n > keep ? n - keep : 0) : n;  // keep < 0: no hints at all
  // (two copies of each unrolled loop, with and without cache hints: a per-load select makes ptxas interleave both)
  auto pass1 = [&](auto hinted) {
    long long j = 0;
    for (; j + 3 < n; j += 4) {
      uint4 raw[4];
      const unsigned long long pol = (j >= keep_from) ? pol_keep : pol_stream;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        raw[u] = decltype(hinted)::value ? Vec16<T>::load_raw_hint(xp + row1(j + u), pol) : Vec16<T>::load_raw(xp + row1(j + u));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[V];
        Vec16<T>::unpack(raw[u], v);
        add(v);
      }
    }
    for (; j < n; ++j) {
      float v[V];
      Vec16<T>::load(xp + row1(j), v);
      add(v);
    }
  };
  if (keep >= 0) pass1(std::true_type{}); else pass1(std::false_type{});
  reduce_row_lanes<2 * V>(acc, sm, tpr_shift);
  if (rl == 0) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      red_add_f64(sums + c + e, (double)acc[e]);
      red_add_f64(sums + C + c + e, (double)acc[V + e]);
    }
  }
  grid_barrier(counter, gridDim.x * gridDim.y);

  // The block's tpr * V channels are finished once, one channel per thread, into shared memory (the double-precision
  // variance / square root / division are ~100 instructions per channel; every thread doing its own V channels made
  // this prologue a visible share of the smaller launches).
  float mu[V], a[V], bb[V];
  {
    const int nch = tpr * V;                      // <= 256 channels per block
    const int ch0 = blockIdx.y * tpr * V;
    if (threadIdx.x < nch) {
      const int ch = ch0 + threadIdx.x;
      const double inv_rows = 1.0 / (double)R;
      const double s1 = ld_f64_cg(sums + ch) * inv_rows;
      const double q = ld_f64_cg(sums + C + ch) * inv_rows;
      const double var = fmax(q - s1 * s1, 0.0);
      const float m = (float)((double)Vec16<T>::scalar(x + ch) + s1);
      const float is = (float)(1.0 / sqrt(var + (double)eps));
      sm[threadIdx.x] = m;
      sm[kBnThreads + threadIdx.x] = __ldg(weight + ch) * is;
      if (blockIdx.x == 0) {  // the saved and the running statistics
        save_mean[ch] = m;
        save_invstd[ch] = is;
        const float shift = conv_bias != nullptr ? __ldg(conv_bias + ch) : 0.f;
        if (running_mean != nullptr) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (m + shift);
        if (running_var != nullptr) {
          const double unbiased = var * ((double)R / (double)(R - 1));
          running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
        }
      }
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < V; ++e) {
      mu[e] = sm[tc * V + e];
      a[e] = sm[kBnThreads + tc * V + e];
      bb[e] = __ldg(bias + c + e);
    }
  }
  auto one = [&](const float (&v)[V], const float (&rv)[V], T* o) {
    float y[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      y[e] = fmaf(v[e] - mu[e], a[e], bb[e]);
      if (RES) y[e] += rv[e];
      if (RELU) y[e] = fmaxf(y[e], 0.f);
    }
    Vec16<T>::store(o, y);
  };
  auto pass2 = [&](auto hinted) {
    long long j = 0;
    for (; j + 3 < n; j += 4) {
      uint4 raw[4], rraw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long off = row2(j + u) + c;
        // second and last use of x, only use of the residual: do not displace what is still to be re-read
        raw[u] = decltype(hinted)::value ? Vec16<T>::load_raw_hint(x + off, pol_stream) : Vec16<T>::load_raw(x + off);
        if (RES) rraw[u] = decltype(hinted)::value ? Vec16<T>::load_raw_hint(res + off, pol_stream) : Vec16<T>::load_raw(res + off);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[V], rv[V];
        Vec16<T>::unpack(raw[u], v);
        if (RES) Vec16<T>::unpack(rraw[u], rv);
        one(v, rv, out + row2(j + u) + c);
      }
    }
    for (; j < n; ++j) {
      float v[V], rv[V];
      const long long off = row2(j) + c;
      Vec16<T>::load(x + off, v);
      if (RES) Vec16<T>::load(res + off, rv);
      one(v, rv, out + off);
    }
  };
  if (keep >= 0) pass2(std::true_type{}); else pass2(std::false_type{});
}

template <typename T, bool RELU>
__global__ void __launch_bounds__(kBnThreads, 2)
bn_bwd_persistent_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ weight,
                         const float* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ invstd,
                         double* __restrict__ sums, unsigned int* __restrict__ counter, T* __restrict__ dx,
                         float* __restrict__ dweight, float* __restrict__ dbias, float* __restrict__ colsum, long long R,
                         int C, int tpr_shift, int first_reverse, int keep) {
  constexpr int V = Vec16<T>::V;
  __shared__ float sm[3 * V * kBnThreads];
  const int tpr = 1 << tpr_shift;
  const int rpp = kBnThreads >> tpr_shift;
  const int tc = threadIdx.x & (tpr - 1);
  const int rl = threadIdx.x >> tpr_shift;
  const int c = (blockIdx.y * tpr + tc) * V;
  float mu[V], is[V], a[V], b[V], acc[3 * V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    mu[e] = __ldg(mean + c + e);
    is[e] = __ldg(invstd + c + e);
    a[e] = __ldg(weight + c + e) * is[e];
    b[e] = RELU ? __ldg(bias + c + e) : 0.f;
    acc[e] = 0.f; acc[V + e] = 0.f; acc[2 * V + e] = 0.f;
  }
  const T* xp = x + c;
  const T* gp = dy + c;
  const long long step = (long long)gridDim.x * rpp;
  const long long r0 = (long long)blockIdx.x * rpp + rl;
  auto add = [&](const float (&v)[V], const float (&g)[V]) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float d = v[e] - mu[e];
      float gz = g[e];
      if (RELU) gz = fmaf(d, a[e], b[e]) > 0.f ? gz : 0.f;
      acc[e] += gz;
      acc[V + e] = fmaf(gz, d * is[e], acc[V + e]);
      acc[2 * V + e] += d;
    }
  };
  const long long n = r0 < R ? (R - r0 + step - 1) / step : 0;   // (directions: see bn_fwd_persistent_kernel)
  auto row1 = [&](long long j) { return (r0 + (first_reverse ? (n - 1 - j) : j) * step) * C; };
  auto row2 = [&](long long j) { return (r0 + (first_reverse ? j : (n - 1 - j)) * step) * C; };
  const unsigned long long pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  const long long keep_from = keep >= 0 ? (n > keep ? n - keep : 0) : n;
  auto pass1 = [&](auto hinted) {
    long long j = 0;
    for (; j + 3 < n; j += 4) {
      uint4 raw[4], graw[4];
      const unsigned long long pol = (j >= keep_from) ? pol_keep : pol_stream;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        raw[u] = decltype(hinted)::value ? Vec16<T>::load_raw_hint(xp + row1(j + u), pol) : Vec16<T>::load_raw(xp + row1(j + u));
        graw[u] = decltype(hinted)::value ? Vec16<T>::load_raw_hint(gp + row1(j + u), pol) : Vec16<T>::load_raw(gp + row1(j + u));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[V], g[V];
        Vec16<T>::unpack(raw[u], v);
        Vec16<T>::unpack(graw[u], g);
        add(v, g);
      }
    }
    for (; j < n; ++j) {
      float v[V], g[V];
      Vec16<T>::load(xp + row1(j), v);
      Vec16<T>::load(gp + row1(j), g);
      add(v, g);
    }
  };
  if (keep >= 0) pass1(std::true_type{}); else pass1(std::false_type{});
  reduce_row_lanes<3 * V>(acc, sm, tpr_shift);
  if (rl == 0) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      red_add_f64(sums + c + e, (double)acc[e]);
      red_add_f64(sums + C + c + e, (double)acc[V + e]);
      red_add_f64(sums + 2 * C + c + e, (double)acc[2 * V + e]);
    }
  }
  grid_barrier(counter, gridDim.x * gridDim.y);

  const double inv_rows = 1.0 / (double)R;
  float c1[V], c2[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    const double s1 = ld_f64_cg(sums + c + e), s2 = ld_f64_cg(sums + C + c + e);
    c1[e] = (float)(s1 * inv_rows);
    c2[e] = (float)(s2 * inv_rows);
    if (blockIdx.x == 0 && rl == 0) {
      dbias[c + e] = (float)s1;
      dweight[c + e] = (float)s2;
      if (colsum != nullptr) {
        const double s3 = ld_f64_cg(sums + 2 * C + c + e);
        colsum[c + e] = (float)((double)a[e] * ((s1 - (double)R * (double)c1[e]) - (double)is[e] * (double)c2[e] * s3));
      }
    }
  }
  auto one = [&](const float (&v)[V], const float (&g)[V], T* o) {
    float rr[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float d = v[e] - mu[e];
      float gz = g[e];
      if (RELU) gz = fmaf(d, a[e], b[e]) > 0.f ? gz : 0.f;
      rr[e] = a[e] * (gz - c1[e] - d * is[e] * c2[e]);
    }
    Vec16<T>::store(o, rr);
  };
  auto pass2 = [&](auto hinted) {
    long long j = 0;
    for (; j + 3 < n; j += 4) {
      uint4 raw[4], graw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long off = row2(j + u) + c;
        raw[u] = decltype(hinted)::value ? Vec16<T>::load_raw_hint(x + off, pol_stream) : Vec16<T>::load_raw(x + off);
        graw[u] = decltype(hinted)::value ? Vec16<T>::load_raw_hint(dy + off, pol_stream) : Vec16<T>::load_raw(dy + off);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[V], g[V];
        Vec16<T>::unpack(raw[u], v);
        Vec16<T>::unpack(graw[u], g);
        one(v, g, dx + row2(j + u) + c);
      }
    }
    for (; j < n; ++j) {
      float v[V], g[V];
      const long long off = row2(j) + c;
      Vec16<T>::load(x + off, v);
      Vec16<T>::load(dy + off, g);
      one(v, g, dx + off);
    }
  };
  if (keep >= 0) pass2(std::true_type{}); else pass2(std::false_type{});
}

// Cooperative launch of one of the single-launch kernels on grid (gx, ctiles); returns false when all blocks cannot be
// resident (the caller then takes the two-kernel form).  Blocks per SM from the occupancy calculator, cached per
// kernel and device.
struct CoopInfo {
  int blocks_per_sm[64] = {};
};
template <typename Kernel>
bool coop_capacity(Kernel kernel, CoopInfo& info, int* capacity) {
  const int dev = current_device();
  if (dev < 0 || dev >= 64) return false;
  if (info.blocks_per_sm[dev] == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kBnThreads, 0) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = -1;
    }
    info.blocks_per_sm[dev] = n;
  }
  if (info.blocks_per_sm[dev] < 0) return false;
  *capacity = info.blocks_per_sm[dev] * num_sms();
  return true;
}

// Rows per thread of pass 1 to load with the "evict last" policy: `mb` megabytes of L2 shared by `tensors` input tensors,
// one row slice per (row, 16-byte pack) thread item, i.e. C * elem bytes per row.  mb <= 0: no cache hints (-1).
int keep_rows_per_thread(int mb, int tensors, int C, size_t elem) {
  if (mb <= 0) return -1;
  const double rows_kept = (double)mb * 1048576.0 / ((double)tensors * C * elem);  // rows of the tensor that fit the budget
  // a thread owns every (gx * rpp)-th row; gx * rpp is ~ (4 blocks per SM) * 256 / (C / V) rows per sweep of the grid
  const int V = (int)(16 / elem);
  const double rows_per_sweep = (double)num_sms() * 4 * kBnThreads / ((double)C / V);
  const double k = rows_kept / rows_per_sweep;
  return k < 1.0 ? 0 : (int)k;
}

int apply_grid(long long total, int cv, int per_sm = 8) {
  // whole waves of `per_sm` CTAs per SM, rounded so that grid * 256 is a multiple of cv
  long long need = (total + kBnThreads - 1) / kBnThreads;
  long long g = (long long)num_sms() * per_sm;
  if (g > need) g = need;
  const int q = cv > kBnThreads ? cv / kBnThreads : 1;
  g = (g + q - 1) / q * q;
  return (int)(g < 1 ? q : g);
}

int shift_of(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

template <typename T>
int bn_fwd_t(const void* x_, const void* res_, const float* weight, const float* bias, float* running_mean, float* running_var,
             const float* conv_bias, long long* nbt, void* out_, float* save_mean, float* save_invstd, long long R, int C,
             float eps, float momentum, int relu, void* workspace, bool from_moments, cudaStream_t s) {
  constexpr int V = Vec16<T>::V;
  BnGeom g;
  if (!bn_geometry(R, C, V, &g)) { set_error("bn_train_fwd: needs C %% %d == 0, C/%d a power of two and at least 2 rows", V, V); return GRAFP_EUNSUPPORTED; }
  if (relu && res_ != nullptr) { set_error("bn_train_fwd: ReLU together with a residual is not implemented"); return GRAFP_EUNSUPPORTED; }
  const T* x = static_cast<const T*>(x_);
  const T* res = static_cast<const T*>(res_);
  T* out = static_cast<T*>(out_);
  double* sums = bn_sums_of(workspace);
  unsigned int* counter = reinterpret_cast<unsigned int*>(sums + (size_t)3 * C);
  if (!from_moments) {
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)3 * C * sizeof(double) + 16, s);
    if (e != cudaSuccess) { set_error("bn_train_fwd: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
  }
  if (!from_moments && option(OPT_BN_PERSISTENT) != 0) {
    int tsh = shift_of(g.tpr);
    int first_reverse = option(OPT_BN_REVERSE) != 0;
    int keep = keep_rows_per_thread(option(OPT_BN_L2_KEEP_MB), 1, C, sizeof(T));
    void* args[] = {(void*)&x, (void*)&res, (void*)&weight, (void*)&bias, (void*)&sums, (void*)&counter, (void*)&out,
                    (void*)&save_mean, (void*)&save_invstd, (void*)&running_mean, (void*)&running_var, (void*)&conv_bias,
                    (void*)&nbt, (void*)&R, (void*)&C, (void*)&tsh, (void*)&eps, (void*)&momentum, (void*)&first_reverse,
                    (void*)&keep};
    auto try_launch = [&](auto kernel, CoopInfo& info) -> int {
      int cap = 0;
      if (!coop_capacity(kernel, info, &cap)) return -1;
      int gx = g.gx;
      if ((long long)gx * g.ctiles > cap) gx = cap / g.ctiles;
      if (gx < 1) return -1;
      const cudaError_t le = cudaLaunchCooperativeKernel((const void*)kernel, dim3(gx, g.ctiles), dim3(kBnThreads), args, 0, s);
      if (le != cudaSuccess) { cudaGetLastError(); return -1; }
      return 0;
    };
    static CoopInfo i_relu, i_res, i_plain;
    int rc;
    if (relu) rc = try_launch(bn_fwd_persistent_kernel<T, true, false>, i_relu);
    else if (res) rc = try_launch(bn_fwd_persistent_kernel<T, false, true>, i_res);
    else rc = try_launch(bn_fwd_persistent_kernel<T, false, false>, i_plain);
    if (rc == 0) return check_launch("bn_train_fwd (single launch)");
  }
  // from_moments: the workspace already holds sum x and sum x^2 per channel (grafp_conv1x1_bn_stats_fwd) - apply only
  if (!from_moments) bn_stats_kernel<T><<<dim3(g.gx, g.ctiles), kBnThreads, 0, s>>>(x, sums, R, C, shift_of(g.tpr));
  const int raw = from_moments ? 1 : 0;
  const long long total = R * g.cv;
  const int grid = apply_grid(total, g.cv, 4);
  const size_t coef_bytes = (size_t)3 * C * sizeof(float);
  if (coef_bytes > 48 * 1024) { set_error("bn_train_fwd: more than 4096 channels"); return GRAFP_EUNSUPPORTED; }
  const int rev = option(OPT_BN_REVERSE) != 0;
#define GRAFP_BN_APPLY(RELU_, RES_)                                                                                        \
  bn_apply_kernel<T, RELU_, RES_><<<grid, kBnThreads, coef_bytes, s>>>(x, res, weight, bias, sums, out, save_mean, save_invstd, \
                                                            running_mean, running_var, conv_bias, nbt, R, C, g.cv, eps,  \
                                                            momentum, rev, raw)
  if (relu) GRAFP_BN_APPLY(true, false);
  else if (res) GRAFP_BN_APPLY(false, true);
  else GRAFP_BN_APPLY(false, false);
#undef GRAFP_BN_APPLY
  return check_launch("bn_train_fwd");
}

template <typename T>
int bn_bwd_t(const void* dy_, const void* x_, const float* weight, const float* bias, const float* save_mean,
             const float* save_invstd, void* dx_, float* dweight, float* dbias, float* dx_colsum, long long R, int C, int relu,
             void* workspace, cudaStream_t s) {
  constexpr int V = Vec16<T>::V;
  BnGeom g;
  if (!bn_geometry(R, C, V, &g)) { set_error("bn_train_bwd: needs C %% %d == 0, C/%d a power of two and at least 2 rows", V, V); return GRAFP_EUNSUPPORTED; }
  const T* dy = static_cast<const T*>(dy_);
  const T* x = static_cast<const T*>(x_);
  T* dx = static_cast<T*>(dx_);
  double* sums = reinterpret_cast<double*>(((uintptr_t)workspace + 255) / 256 * 256);
  unsigned int* counter = reinterpret_cast<unsigned int*>(sums + (size_t)3 * C);
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)3 * C * sizeof(double) + 16, s);
  if (e != cudaSuccess) { set_error("bn_train_bwd: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
  int tsh = shift_of(g.tpr);
  if (option(OPT_BN_PERSISTENT) != 0) {
    int first_reverse = option(OPT_BN_REVERSE) != 0;
    int keep = keep_rows_per_thread(option(OPT_BN_L2_KEEP_MB), 2, C, sizeof(T));
    void* args[] = {(void*)&dy, (void*)&x, (void*)&weight, (void*)&bias, (void*)&save_mean, (void*)&save_invstd, (void*)&sums,
                    (void*)&counter, (void*)&dx, (void*)&dweight, (void*)&dbias, (void*)&dx_colsum, (void*)&R, (void*)&C,
                    (void*)&tsh, (void*)&first_reverse, (void*)&keep};
    auto try_launch = [&](auto kernel, CoopInfo& info) -> int {
      int cap = 0;
      if (!coop_capacity(kernel, info, &cap)) return -1;
      int gx = g.gx;
      if ((long long)gx * g.ctiles > cap) gx = cap / g.ctiles;
      if (gx < 1) return -1;
      const cudaError_t le = cudaLaunchCooperativeKernel((const void*)kernel, dim3(gx, g.ctiles), dim3(kBnThreads), args, 0, s);
      if (le != cudaSuccess) { cudaGetLastError(); return -1; }
      return 0;
    };
    static CoopInfo i_relu, i_plain;
    const int rc = relu ? try_launch(bn_bwd_persistent_kernel<T, true>, i_relu) : try_launch(bn_bwd_persistent_kernel<T, false>, i_plain);
    if (rc == 0) return check_launch("bn_train_bwd (single launch)");
  }
  if (relu) bn_bwd_reduce_kernel<T, true><<<dim3(g.gx, g.ctiles), kBnThreads, 0, s>>>(dy, x, weight, bias, save_mean, save_invstd, sums, R, C, tsh);
  else bn_bwd_reduce_kernel<T, false><<<dim3(g.gx, g.ctiles), kBnThreads, 0, s>>>(dy, x, weight, bias, save_mean, save_invstd, sums, R, C, tsh);
  const long long total = R * g.cv;
  const int grid = apply_grid(total, g.cv);
  const int rev = option(OPT_BN_REVERSE) != 0;
  if (relu) bn_bwd_apply_kernel<T, true><<<grid, kBnThreads, 0, s>>>(dy, x, weight, bias, save_mean, save_invstd, sums, dx, dweight, dbias, dx_colsum, R, C, g.cv, rev);
  else bn_bwd_apply_kernel<T, false><<<grid, kBnThreads, 0, s>>>(dy, x, weight, bias, save_mean, save_invstd, sums, dx, dweight, dbias, dx_colsum, R, C, g.cv, rev);
  return check_launch("bn_train_bwd");
}

}  // namespace

size_t bn_workspace_bytes(int C) { return (size_t)3 * C * sizeof(double) + 16 + 256; }

bool bn_supported(long long R, int C, int dtype) {
  BnGeom g;
  const int V = dtype == GRAFP_F32 ? 4 : 8;
  return bn_geometry(R, C, V, &g) && R * (long long)(C / V) < (1LL << 40);
}

int launch_bn_train_fwd(const void* x, const void* res, const float* weight, const float* bias, float* running_mean,
                        float* running_var, const float* conv_bias, long long* num_batches_tracked, void* out, float* save_mean,
                        float* save_invstd, long long R, int C, float eps, float momentum, int relu, int dtype, void* workspace,
                        bool from_moments, cudaStream_t s) {
  if (dtype == GRAFP_F32)
    return bn_fwd_t<float>(x, res, weight, bias, running_mean, running_var, conv_bias, num_batches_tracked, out, save_mean,
                           save_invstd, R, C, eps, momentum, relu, workspace, from_moments, s);
  return bn_fwd_t<__nv_bfloat16>(x, res, weight, bias, running_mean, running_var, conv_bias, num_batches_tracked, out, save_mean,
                                 save_invstd, R, C, eps, momentum, relu, workspace, from_moments, s);
}

double* bn_workspace_sums(void* workspace) { return bn_sums_of(workspace); }

int launch_bn_train_bwd(const void* dy, const void* x, const float* weight, const float* bias, const float* save_mean,
                        const float* save_invstd, void* dx, float* dweight, float* dbias, float* dx_colsum, long long R, int C,
                        int relu, int dtype, void* workspace, cudaStream_t s) {
  if (dtype == GRAFP_F32)
    return bn_bwd_t<float>(dy, x, weight, bias, save_mean, save_invstd, dx, dweight, dbias, dx_colsum, R, C, relu, workspace, s);
  return bn_bwd_t<__nv_bfloat16>(dy, x, weight, bias, save_mean, save_invstd, dx, dweight, dbias, dx_colsum, R, C, relu, workspace, s);
}

}  // namespace grafp
