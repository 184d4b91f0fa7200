// Train-mode BatchNorm over node rows, fused with the ReLU and / or the residual add that follow it
// in the Grapher / FFN blocks (reference: torch_vertex.py:152-162,183-194 fc1 / fc2 = Conv2d + BatchNorm2d,
// graph_encoder.py:45-67 FFN, torch_nn.py:52-64 BasicConv = Conv2d + BatchNorm2d + ReLU).  SURVEY 8(f) row 2.
//
// Activations are rows (R = B*N, C) fp32, C % 4 == 0 and C/4 a power of two.  HBM-bound streaming:
//   forward : statistics pass (read x) + apply pass (read x [, residual], write y)            -> 3-4 tensor passes
//             (PyTorch: BN 3 + ReLU 2 or add 3)
//   backward: reduction pass (read dy, x) + apply pass (read dy, x, write dx)                  -> 5 tensor passes
//             (PyTorch: ReLU backward 3 + BN backward 5); the ReLU mask is recomputed from x, so neither the
//             BN output nor the ReLU output is kept for backward.
// Statistics are accumulated per thread around a shift (the first element it sees), merged with Chan's
// formula and finalised in double precision, so they do not suffer E[x^2] - E[x]^2 cancellation.
#include "common.cuh"

namespace grafp {
namespace {

constexpr int kBnThreads = 256;
constexpr int kBnMaxPartials = 1024;

struct BnGeom {
  int cv;      // float4 packs per row
  int tpr;     // threads per row inside a block (power of two <= 256)
  int rpp;     // rows per block pass = 256 / tpr
  int ctiles;  // channel tiles (grid.y) = cv / tpr
  int gx;      // row blocks (grid.x)
};

inline bool bn_geometry(long long R, int C, BnGeom* g) {
  if (C < 4 || C % 4 != 0 || R < 2) return false;
  const int cv = C / 4;
  if ((cv & (cv - 1)) != 0) return false;
  g->cv = cv;
  g->tpr = cv < kBnThreads ? cv : kBnThreads;
  g->rpp = kBnThreads / g->tpr;
  g->ctiles = cv / g->tpr;
  long long need = (R + g->rpp - 1) / g->rpp;
  long long cap = (long long)num_sms() * 4 / g->ctiles;
  if (cap < 1) cap = 1;
  if (cap > kBnMaxPartials) cap = kBnMaxPartials;
  g->gx = (int)(need < cap ? need : cap);
  return true;
}

__device__ __forceinline__ void chan_merge(float& na, float& ma, float& Ma, float nb, float mb, float Mb) {
  if (nb == 0.f) return;
  const float n = na + nb;
  const float d = mb - ma;
  ma = ma + d * (nb / n);
  Ma = Ma + Mb + d * d * (na * nb / n);
  na = n;
}

// ---- forward statistics: per block and channel (mean, M2) over the rows the block owns ----
__global__ void __launch_bounds__(kBnThreads)
bn_stats_kernel(const float* __restrict__ x, float2* __restrict__ partial, int* __restrict__ pcount, long long R, int C,
                int tpr_shift) {
  __shared__ float4 s_mean[kBnThreads];
  __shared__ float4 s_m2[kBnThreads];
  __shared__ float s_n[kBnThreads];
  const int tpr = 1 << tpr_shift;
  const int rpp = kBnThreads >> tpr_shift;
  const int tc = threadIdx.x & (tpr - 1);
  const int rl = threadIdx.x >> tpr_shift;
  const int c4 = blockIdx.y * tpr + tc;
  const float4* xp = reinterpret_cast<const float4*>(x) + c4;
  const long long cv = C >> 2;
  const long long step = (long long)gridDim.x * rpp;
  long long r = (long long)blockIdx.x * rpp + rl;
  float4 sh = make_float4(0.f, 0.f, 0.f, 0.f), s = sh, q = sh;
  float n = 0.f;
  if (r < R) sh = __ldg(xp + r * cv);
  for (; r + 3 * step < R; r += 4 * step) {  // four independent rows in flight
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(xp + (r + u * step) * cv);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float dx = v[u].x - sh.x, dy = v[u].y - sh.y, dz = v[u].z - sh.z, dw = v[u].w - sh.w;
      s.x += dx; s.y += dy; s.z += dz; s.w += dw;
      q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
    }
    n += 4.f;
  }
  for (; r < R; r += step) {
    const float4 v = __ldg(xp + r * cv);
    const float dx = v.x - sh.x, dy = v.y - sh.y, dz = v.z - sh.z, dw = v.w - sh.w;
    s.x += dx; s.y += dy; s.z += dz; s.w += dw;
    q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
    n += 1.f;
  }
  const float inv = n > 0.f ? 1.f / n : 0.f;
  float4 mean = make_float4(sh.x + s.x * inv, sh.y + s.y * inv, sh.z + s.z * inv, sh.w + s.w * inv);
  float4 m2 = make_float4(q.x - s.x * s.x * inv, q.y - s.y * s.y * inv, q.z - s.z * s.z * inv, q.w - s.w * s.w * inv);
  s_mean[threadIdx.x] = mean; s_m2[threadIdx.x] = m2; s_n[threadIdx.x] = n;
  __syncthreads();
  // tree merge over the row lanes that share a channel pack
  for (int half = rpp >> 1; half >= 1; half >>= 1) {
    if (rl < half) {
      const int o = threadIdx.x + (half << tpr_shift);
      float na = s_n[threadIdx.x];
      const float nb = s_n[o];
      float4 ma = s_mean[threadIdx.x], Ma = s_m2[threadIdx.x];
      const float4 mb = s_mean[o], Mb = s_m2[o];
      float t;
      t = na; chan_merge(t, ma.x, Ma.x, nb, mb.x, Mb.x);
      t = na; chan_merge(t, ma.y, Ma.y, nb, mb.y, Mb.y);
      t = na; chan_merge(t, ma.z, Ma.z, nb, mb.z, Mb.z);
      chan_merge(na, ma.w, Ma.w, nb, mb.w, Mb.w);
      s_mean[threadIdx.x] = ma; s_m2[threadIdx.x] = Ma; s_n[threadIdx.x] = na;
    }
    __syncthreads();
  }
  if (rl == 0) {
    const float4 m = s_mean[threadIdx.x], M = s_m2[threadIdx.x];
    float2* p = partial + (long long)blockIdx.x * C + c4 * 4;
    p[0] = make_float2(m.x, M.x); p[1] = make_float2(m.y, M.y); p[2] = make_float2(m.z, M.z); p[3] = make_float2(m.w, M.w);
    if (tc == 0 && blockIdx.y == 0) pcount[blockIdx.x] = (int)s_n[threadIdx.x];
  }
}

// 32 channels x 32 partial-groups per block (coalesced 256-byte reads of the partials, <= 19 partials per thread
// with four loads in flight: these kernels are pure latency, 20 - 34 us each with 8 groups and a serial loop):
// merge the block partials in double (two division-free passes: the global mean, then
// M2 = sum [M2_p + n_p (mean_p - mean)^2]), emit mean / invstd, update the running statistics
constexpr int kFinGroups = 32;
constexpr int kFinThreads = 32 * kFinGroups;

__device__ __forceinline__ double group_sum(double v, double (*sm)[33], int cl, int pg) {
  __syncthreads();
  sm[pg][cl] = v;
  __syncthreads();
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
  for (int i = 0; i < kFinGroups; i += 4) { t0 += sm[i][cl]; t1 += sm[i + 1][cl]; t2 += sm[i + 2][cl]; t3 += sm[i + 3][cl]; }
  return (t0 + t1) + (t2 + t3);
}

__global__ void __launch_bounds__(kFinThreads)
bn_stats_finalize_kernel(const float2* __restrict__ partial, const int* __restrict__ pcount, int parts, int C, long long R,
                         float eps, float momentum, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                         float* running_mean, float* running_var) {
  __shared__ double sm[kFinGroups][33];
  const int cl = threadIdx.x & 31, pg = threadIdx.x >> 5;
  const int c = min(blockIdx.x * 32 + cl, C - 1);
  const float2* pc = partial + c;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int p = pg;
  for (; p + 3 * kFinGroups < parts; p += 4 * kFinGroups) {
    const float m0 = pc[(long long)p * C].x, m1 = pc[(long long)(p + kFinGroups) * C].x;
    const float m2_ = pc[(long long)(p + 2 * kFinGroups) * C].x, m3 = pc[(long long)(p + 3 * kFinGroups) * C].x;
    a0 += (double)pcount[p] * (double)m0;
    a1 += (double)pcount[p + kFinGroups] * (double)m1;
    a2 += (double)pcount[p + 2 * kFinGroups] * (double)m2_;
    a3 += (double)pcount[p + 3 * kFinGroups] * (double)m3;
  }
  for (; p < parts; p += kFinGroups) a0 += (double)pcount[p] * (double)pc[(long long)p * C].x;
  const double mean = group_sum((a0 + a1) + (a2 + a3), sm, cl, pg) / (double)R;
  a0 = a1 = a2 = a3 = 0.0;
  p = pg;
  for (; p + 3 * kFinGroups < parts; p += 4 * kFinGroups) {
    const float2 v0 = pc[(long long)p * C], v1 = pc[(long long)(p + kFinGroups) * C];
    const float2 v2 = pc[(long long)(p + 2 * kFinGroups) * C], v3 = pc[(long long)(p + 3 * kFinGroups) * C];
    const double d0 = (double)v0.x - mean, d1 = (double)v1.x - mean, d2 = (double)v2.x - mean, d3 = (double)v3.x - mean;
    a0 += (double)v0.y + (double)pcount[p] * d0 * d0;
    a1 += (double)v1.y + (double)pcount[p + kFinGroups] * d1 * d1;
    a2 += (double)v2.y + (double)pcount[p + 2 * kFinGroups] * d2 * d2;
    a3 += (double)v3.y + (double)pcount[p + 3 * kFinGroups] * d3 * d3;
  }
  for (; p < parts; p += kFinGroups) {
    const float2 v = pc[(long long)p * C];
    const double d = (double)v.x - mean;
    a0 += (double)v.y + (double)pcount[p] * d * d;
  }
  const double m2 = group_sum((a0 + a1) + (a2 + a3), sm, cl, pg);
  if (pg != 0 || blockIdx.x * 32 + cl >= C) return;
  const double var = m2 / (double)R;
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var != nullptr) {
    const double unbiased = m2 / (double)(R - 1);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ---- forward apply: y = (x - mean) * a + bias (+ residual) (ReLU), a = weight * invstd ----
template <bool RELU, bool RES>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ weight,
                const float* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ invstd,
                float* __restrict__ out, long long total4, int cv) {
  const long long T = (long long)gridDim.x * kBnThreads;  // a multiple of cv: a thread's channel pack never changes
  const long long i0 = (long long)blockIdx.x * kBnThreads + threadIdx.x;
  const int c = (int)(i0 % cv) * 4;
  const float4 w = *reinterpret_cast<const float4*>(weight + c), bb = *reinterpret_cast<const float4*>(bias + c);
  const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
  const float4 a = make_float4(w.x * is.x, w.y * is.y, w.z * is.z, w.w * is.w);
  const float4* xp = reinterpret_cast<const float4*>(x);
  const float4* rp = reinterpret_cast<const float4*>(res);
  float4* op = reinterpret_cast<float4*>(out);
  // (x - mean) first: exact-ish difference, no cancellation against a pre-multiplied shift when |mean| >> std
  auto one = [&](float4 v, float4 rv) {
    float4 y = make_float4(fmaf(v.x - mu.x, a.x, bb.x), fmaf(v.y - mu.y, a.y, bb.y), fmaf(v.z - mu.z, a.z, bb.z),
                           fmaf(v.w - mu.w, a.w, bb.w));
    if (RES) { y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w; }
    if (RELU) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
    return y;
  };
  long long i = i0;
  for (; i + 3 * T < total4; i += 4 * T) {
    float4 v[4], rv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = __ldg(xp + i + u * T);
      rv[u] = RES ? __ldg(rp + i + u * T) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) op[i + u * T] = one(v[u], rv[u]);
  }
  for (; i < total4; i += T) op[i] = one(__ldg(xp + i), RES ? __ldg(rp + i) : make_float4(0.f, 0.f, 0.f, 0.f));
}

// ---- backward reduction: s1 = sum dz, s2 = sum dz * xhat, dz = dy masked by the recomputed ReLU ----
template <bool RELU>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ weight,
                     const float* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ invstd,
                     float2* __restrict__ partial, long long R, int C, int tpr_shift) {
  __shared__ float4 s_a[kBnThreads];
  __shared__ float4 s_b[kBnThreads];
  const int tpr = 1 << tpr_shift;
  const int rpp = kBnThreads >> tpr_shift;
  const int tc = threadIdx.x & (tpr - 1);
  const int rl = threadIdx.x >> tpr_shift;
  const int c4 = blockIdx.y * tpr + tc;
  const int c = c4 * 4;
  const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;  // ReLU mask: (x - mean) * a + b > 0, the forward's expression
  if (RELU) {
    const float4 w = *reinterpret_cast<const float4*>(weight + c);
    b = *reinterpret_cast<const float4*>(bias + c);
    a = make_float4(w.x * is.x, w.y * is.y, w.z * is.z, w.w * is.w);
  }
  const float4* xp = reinterpret_cast<const float4*>(x) + c4;
  const float4* gp = reinterpret_cast<const float4*>(dy) + c4;
  const long long cv = C >> 2;
  const long long step = (long long)gridDim.x * rpp;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  auto acc = [&](float4 v, float4 g) {
    v.x -= mu.x; v.y -= mu.y; v.z -= mu.z; v.w -= mu.w;
    if (RELU) {
      g.x = fmaf(v.x, a.x, b.x) > 0.f ? g.x : 0.f; g.y = fmaf(v.y, a.y, b.y) > 0.f ? g.y : 0.f;
      g.z = fmaf(v.z, a.z, b.z) > 0.f ? g.z : 0.f; g.w = fmaf(v.w, a.w, b.w) > 0.f ? g.w : 0.f;
    }
    s1.x += g.x; s1.y += g.y; s1.z += g.z; s1.w += g.w;
    s2.x = fmaf(g.x, v.x * is.x, s2.x); s2.y = fmaf(g.y, v.y * is.y, s2.y);
    s2.z = fmaf(g.z, v.z * is.z, s2.z); s2.w = fmaf(g.w, v.w * is.w, s2.w);
  };
  long long r = (long long)blockIdx.x * rpp + rl;
  for (; r + 3 * step < R; r += 4 * step) {
    float4 v[4], g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { v[u] = __ldg(xp + (r + u * step) * cv); g[u] = __ldg(gp + (r + u * step) * cv); }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc(v[u], g[u]);
  }
  for (; r < R; r += step) acc(__ldg(xp + r * cv), __ldg(gp + r * cv));
  s_a[threadIdx.x] = s1; s_b[threadIdx.x] = s2;
  __syncthreads();
  for (int half = rpp >> 1; half >= 1; half >>= 1) {
    if (rl < half) {
      const int o = threadIdx.x + (half << tpr_shift);
      float4 p = s_a[threadIdx.x], q = s_b[threadIdx.x];
      const float4 po = s_a[o], qo = s_b[o];
      p.x += po.x; p.y += po.y; p.z += po.z; p.w += po.w;
      q.x += qo.x; q.y += qo.y; q.z += qo.z; q.w += qo.w;
      s_a[threadIdx.x] = p; s_b[threadIdx.x] = q;
    }
    __syncthreads();
  }
  if (rl == 0) {
    const float4 p = s_a[threadIdx.x], q = s_b[threadIdx.x];
    float2* o = partial + (long long)blockIdx.x * C + c;
    o[0] = make_float2(p.x, q.x); o[1] = make_float2(p.y, q.y); o[2] = make_float2(p.z, q.z); o[3] = make_float2(p.w, q.w);
  }
}

__global__ void __launch_bounds__(kFinThreads)
bn_bwd_finalize_kernel(const float2* __restrict__ partial, int parts, int C, float* __restrict__ dweight,
                       float* __restrict__ dbias) {
  __shared__ double sm[kFinGroups][33];
  const int cl = threadIdx.x & 31, pg = threadIdx.x >> 5;
  const int c = min(blockIdx.x * 32 + cl, C - 1);
  const float2* pc = partial + c;
  double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
  int p = pg;
  for (; p + 3 * kFinGroups < parts; p += 4 * kFinGroups) {
    float2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = pc[(long long)(p + u * kFinGroups) * C];
#pragma unroll
    for (int u = 0; u < 4; ++u) { s1[u] += (double)v[u].x; s2[u] += (double)v[u].y; }
  }
  for (; p < parts; p += kFinGroups) {
    const float2 v = pc[(long long)p * C];
    s1[0] += (double)v.x;
    s2[0] += (double)v.y;
  }
  const double t1 = group_sum((s1[0] + s1[1]) + (s1[2] + s1[3]), sm, cl, pg);
  const double t2 = group_sum((s2[0] + s2[1]) + (s2[2] + s2[3]), sm, cl, pg);
  if (pg == 0 && blockIdx.x * 32 + cl < C) { dbias[c] = (float)t1; dweight[c] = (float)t2; }
}

// ---- backward apply: dx = weight * invstd * (dz - s1 / R - xhat * s2 / R) ----
template <bool RELU, bool COLSUM>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ weight,
                    const float* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ dweight, const float* __restrict__ dbias, float* __restrict__ dx,
                    float* __restrict__ colsum_partial, long long total4, int cv, float inv_rows) {
  const long long T = (long long)gridDim.x * kBnThreads;
  const long long i0 = (long long)blockIdx.x * kBnThreads + threadIdx.x;
  const int c = (int)(i0 % cv) * 4;
  const float4 w = *reinterpret_cast<const float4*>(weight + c);
  const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
  const float4 a = make_float4(w.x * is.x, w.y * is.y, w.z * is.z, w.w * is.w);
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (RELU) b = *reinterpret_cast<const float4*>(bias + c);
  const float4 d1 = *reinterpret_cast<const float4*>(dbias + c), d2 = *reinterpret_cast<const float4*>(dweight + c);
  const float4 c1 = make_float4(d1.x * inv_rows, d1.y * inv_rows, d1.z * inv_rows, d1.w * inv_rows);
  const float4 c2 = make_float4(d2.x * inv_rows, d2.y * inv_rows, d2.z * inv_rows, d2.w * inv_rows);
  const float4* xp = reinterpret_cast<const float4*>(x);
  const float4* gp = reinterpret_cast<const float4*>(dy);
  float4* op = reinterpret_cast<float4*>(dx);
  auto one = [&](float4 v, float4 g) {
    v.x -= mu.x; v.y -= mu.y; v.z -= mu.z; v.w -= mu.w;
    if (RELU) {
      g.x = fmaf(v.x, a.x, b.x) > 0.f ? g.x : 0.f; g.y = fmaf(v.y, a.y, b.y) > 0.f ? g.y : 0.f;
      g.z = fmaf(v.z, a.z, b.z) > 0.f ? g.z : 0.f; g.w = fmaf(v.w, a.w, b.w) > 0.f ? g.w : 0.f;
    }
    float4 r;
    r.x = a.x * (g.x - c1.x - v.x * is.x * c2.x);
    r.y = a.y * (g.y - c1.y - v.y * is.y * c2.y);
    r.z = a.z * (g.z - c1.z - v.z * is.z * c2.z);
    r.w = a.w * (g.w - c1.w - v.w * is.w * c2.w);
    return r;
  };
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);  // column sums of dx over this thread's rows (COLSUM)
  long long i = i0;
  for (; i + 3 * T < total4; i += 4 * T) {
    float4 v[4], g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { v[u] = __ldg(xp + i + u * T); g[u] = __ldg(gp + i + u * T); }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 r = one(v[u], g[u]);
      op[i + u * T] = r;
      if (COLSUM) { cs.x += r.x; cs.y += r.y; cs.z += r.z; cs.w += r.w; }
    }
  }
  for (; i < total4; i += T) {
    const float4 r = one(__ldg(xp + i), __ldg(gp + i));
    op[i] = r;
    if (COLSUM) { cs.x += r.x; cs.y += r.y; cs.z += r.z; cs.w += r.w; }
  }
  if (COLSUM) {
    // block partial per channel: threads tid, tid + cv, ... of a block share a channel pack (cv <= 256), or own it (cv > 256)
    __shared__ float4 s_cs[kBnThreads];
    s_cs[threadIdx.x] = cs;
    __syncthreads();
    const int span = cv < kBnThreads ? cv : kBnThreads;
    if ((int)threadIdx.x < span) {
      float4 t = cs;
      for (int o = threadIdx.x + span; o < kBnThreads; o += span) {
        const float4 q = s_cs[o];
        t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
      }
      // channel pack of thread tid in this block: (blockIdx.x * 256 + tid) % cv
      *reinterpret_cast<float4*>(colsum_partial + (long long)blockIdx.x * (4LL * span) + 4 * threadIdx.x) = t;
    }
  }
}

// column sums of dx: sum the block partials of bn_bwd_apply_kernel (block b, slot t holds channel pack (b * 256 + t) % cv)
__global__ void __launch_bounds__(kFinThreads)
bn_colsum_finalize_kernel(const float* __restrict__ colsum_partial, int blocks, int cv, float* __restrict__ colsum) {
  __shared__ double sm[kFinGroups][33];
  const int cl = threadIdx.x & 31, pg = threadIdx.x >> 5;
  const int C = cv * 4;
  const int c = min(blockIdx.x * 32 + cl, C - 1);
  const int c4 = c >> 2, e = c & 3;
  const int span = cv < kBnThreads ? cv : kBnThreads;
  // cv <= 256: every block holds every channel pack once, at slot c4; else block b holds packs (b * 256 + t) % cv,
  // i.e. pack c4 lives in the blocks first, first + per, ... at slot t
  const int per = cv <= kBnThreads ? 1 : cv / kBnThreads;
  const int first = cv <= kBnThreads ? 0 : c4 / kBnThreads;
  const int slot = cv <= kBnThreads ? c4 : c4 % kBnThreads;
  const float* src = colsum_partial + 4 * slot + e;
  const long long bstride = 4LL * span;
  const int step = kFinGroups * per;
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  int b = first + pg * per;
  for (; b + 3 * step < blocks; b += 4 * step) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = src[(long long)(b + u * step) * bstride];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] += (double)v[u];
  }
  for (; b < blocks; b += step) a[0] += (double)src[(long long)b * bstride];
  const double tot = group_sum((a[0] + a[1]) + (a[2] + a[3]), sm, cl, pg);
  if (pg == 0 && blockIdx.x * 32 + cl < C) colsum[c] = (float)tot;
}

int apply_grid(long long total4, int cv) {
  // whole waves of 8 CTAs per SM, rounded so that grid * 256 is a multiple of cv
  long long need = (total4 + kBnThreads - 1) / kBnThreads;
  long long g = (long long)num_sms() * 8;
  if (g > need) g = need;
  const int q = cv > kBnThreads ? cv / kBnThreads : 1;
  g = (g + q - 1) / q * q;
  return (int)(g < 1 ? q : g);
}

int shift_of(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

}  // namespace

size_t bn_workspace_bytes(int C) {
  const size_t partials = (size_t)kBnMaxPartials * (size_t)C * sizeof(float2) + kBnMaxPartials * sizeof(int);
  const size_t colsums = (size_t)(num_sms() * 8 + 8) * 4 * kBnThreads * sizeof(float);  // one float4 per thread of the apply grid
  return (partials > colsums ? partials : colsums) + 256;
}

bool bn_supported(long long R, int C) {
  BnGeom g;
  return bn_geometry(R, C, &g) && R * (long long)(C / 4) < (1LL << 40);
}

int launch_bn_train_fwd(const float* x, const float* res, const float* weight, const float* bias, float* running_mean,
                        float* running_var, float* out, float* save_mean, float* save_invstd, long long R, int C, float eps,
                        float momentum, int relu, void* workspace, cudaStream_t s) {
  BnGeom g;
  if (!bn_geometry(R, C, &g)) { set_error("bn_train_fwd: needs C %% 4 == 0, C/4 a power of two and at least 2 rows"); return GRAFP_EUNSUPPORTED; }
  if (relu && res != nullptr) { set_error("bn_train_fwd: ReLU together with a residual is not implemented"); return GRAFP_EUNSUPPORTED; }
  char* wb = reinterpret_cast<char*>(((uintptr_t)workspace + 255) / 256 * 256);
  float2* partial = reinterpret_cast<float2*>(wb);
  int* pcount = reinterpret_cast<int*>(wb + (size_t)kBnMaxPartials * C * sizeof(float2));
  bn_stats_kernel<<<dim3(g.gx, g.ctiles), kBnThreads, 0, s>>>(x, partial, pcount, R, C, shift_of(g.tpr));
  bn_stats_finalize_kernel<<<(C + 31) / 32, kFinThreads, 0, s>>>(partial, pcount, g.gx, C, R, eps, momentum, save_mean, save_invstd,
                                                           running_mean, running_var);
  const long long total4 = R * g.cv;
  const int grid = apply_grid(total4, g.cv);
  if (relu) bn_apply_kernel<true, false><<<grid, kBnThreads, 0, s>>>(x, res, weight, bias, save_mean, save_invstd, out, total4, g.cv);
  else if (res) bn_apply_kernel<false, true><<<grid, kBnThreads, 0, s>>>(x, res, weight, bias, save_mean, save_invstd, out, total4, g.cv);
  else bn_apply_kernel<false, false><<<grid, kBnThreads, 0, s>>>(x, res, weight, bias, save_mean, save_invstd, out, total4, g.cv);
  return check_launch("bn_train_fwd");
}

int launch_bn_train_bwd(const float* dy, const float* x, const float* weight, const float* bias, const float* save_mean,
                        const float* save_invstd, float* dx, float* dweight, float* dbias, float* dx_colsum, long long R, int C,
                        int relu, void* workspace, cudaStream_t s) {
  BnGeom g;
  if (!bn_geometry(R, C, &g)) { set_error("bn_train_bwd: needs C %% 4 == 0, C/4 a power of two and at least 2 rows"); return GRAFP_EUNSUPPORTED; }
  char* wb = reinterpret_cast<char*>(((uintptr_t)workspace + 255) / 256 * 256);
  float2* partial = reinterpret_cast<float2*>(wb);
  const int tsh = shift_of(g.tpr);
  if (relu) bn_bwd_reduce_kernel<true><<<dim3(g.gx, g.ctiles), kBnThreads, 0, s>>>(dy, x, weight, bias, save_mean, save_invstd, partial, R, C, tsh);
  else bn_bwd_reduce_kernel<false><<<dim3(g.gx, g.ctiles), kBnThreads, 0, s>>>(dy, x, weight, bias, save_mean, save_invstd, partial, R, C, tsh);
  bn_bwd_finalize_kernel<<<(C + 31) / 32, kFinThreads, 0, s>>>(partial, g.gx, C, dweight, dbias);
  const long long total4 = R * g.cv;
  const int grid = apply_grid(total4, g.cv);
  const float inv_rows = (float)(1.0 / (double)R);
  float* csp = reinterpret_cast<float*>(wb);  // the reduction partials are consumed by now (stream order)
  if (dx_colsum != nullptr) {
    if (relu) bn_bwd_apply_kernel<true, true><<<grid, kBnThreads, 0, s>>>(dy, x, weight, bias, save_mean, save_invstd, dweight, dbias, dx, csp, total4, g.cv, inv_rows);
    else bn_bwd_apply_kernel<false, true><<<grid, kBnThreads, 0, s>>>(dy, x, weight, bias, save_mean, save_invstd, dweight, dbias, dx, csp, total4, g.cv, inv_rows);
    bn_colsum_finalize_kernel<<<(C + 31) / 32, kFinThreads, 0, s>>>(csp, grid, g.cv, dx_colsum);
  } else {
    if (relu) bn_bwd_apply_kernel<true, false><<<grid, kBnThreads, 0, s>>>(dy, x, weight, bias, save_mean, save_invstd, dweight, dbias, dx, nullptr, total4, g.cv, inv_rows);
    else bn_bwd_apply_kernel<false, false><<<grid, kBnThreads, 0, s>>>(dy, x, weight, bias, save_mean, save_invstd, dweight, dbias, dx, nullptr, total4, g.cv, inv_rows);
  }
  return check_launch("bn_train_bwd");
}

}  // namespace grafp
