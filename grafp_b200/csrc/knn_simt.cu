// k-NN graph, CUDA-core path: node normalisation (shared with the tensor-core path)
// and an exact-fp32 tiled Gram kernel with the top-k selection fused behind every
// 64x64 distance tile, so the N x M distance matrix only ever exists tile by tile in
// shared memory.  This path handles every shape (any C, N, M, K <= 64) and is the
// on-device cross-check for the tcgen05 path.
#include "knn.cuh"

#include <cuda_fp16.h>

namespace grafp {

constexpr float kF16PlaneScale = 4096.f;  // mode 3: planes hold x_hat * 2^12 (exact scaling), see knn_tc2.cu

// ------------------------------------------------------------------------------------
// normalisation: x_hat = x / max(||x||_2, 1e-12) per node (F.normalize, torch_edge.py:281)
// one warp per node row; optionally also emits the TF32 hi/lo split used by the
// tensor-core path and the squared norm of x_hat (torch_edge.py:17).
// ------------------------------------------------------------------------------------
template <typename T, int MODE>  // MODE 0: x_hat fp32; 1: hi/lo tf32 split; 2: x_hat bf16; 3: hi/lo fp16 split of x_hat * 2^12
__global__ void __launch_bounds__(256)
knn_normalize_kernel(const T* __restrict__ x, float* __restrict__ xhat, float* __restrict__ lo,
                     float* __restrict__ sq, long long rows, int C, bool normalize) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp0; row < rows; row += nwarps) {
    const T* xr = x + row * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = to_float(xr[c]);
      ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float denom = normalize ? fmaxf(sqrtf(ss), 1e-12f) : 1.f;
    float s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      float v = __fdiv_rn(to_float(xr[c]), denom);
      if constexpr (MODE == 1) {
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const float l = v - hi;  // exact
        xhat[row * C + c] = hi;
        lo[row * C + c] = __uint_as_float(__float_as_uint(l) & 0xffffe000u);
      } else if constexpr (MODE == 2) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        reinterpret_cast<__nv_bfloat16*>(xhat)[row * C + c] = h;
        v = __bfloat162float(h);  // the Gram runs on the rounded values, so must |x|^2
      } else if constexpr (MODE == 3) {
        const float sv = v * kF16PlaneScale;
        const __half h = __float2half_rn(sv);
        reinterpret_cast<__half*>(xhat)[row * C + c] = h;
        reinterpret_cast<__half*>(lo)[row * C + c] = __float2half_rn(sv - __half2float(h));
      } else {
        xhat[row * C + c] = v;
      }
      s2 = fmaf(v, v, s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (lane == 0) sq[row] = s2;
  }
}

// Vectorised form for C % 4 == 0 and C <= 1024: a group of LPR lanes owns one node row, every lane
// keeps its 4-channel packs in registers (one 16-byte load per pack), the two reductions are
// xor-shuffles inside the group.  LPR = min(32, pow2 >= C/4), so a warp covers 32/LPR rows.
template <typename T, int MODE, int LPR, int PACKS>
__global__ void __launch_bounds__(256)
knn_normalize_vec_kernel(const T* __restrict__ x, float* __restrict__ xhat, float* __restrict__ lo,
                         float* __restrict__ sq, long long rows, int C, bool normalize) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  constexpr int RPW = 32 / LPR;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int cv = C / 4;
  for (long long rb = warp0 * RPW; rb < rows; rb += nwarps * RPW) {
    const long long row = rb + lane / LPR;
    const bool live = row < rows;
    float v[PACKS][4];
    float ss = 0.f;
#pragma unroll
    for (int p = 0; p < PACKS; ++p) {
      const int c4 = sub + p * LPR;
      if (live && c4 < cv) Pack<T, 4>::load(x + row * C + c4 * 4, v[p]);
      else { v[p][0] = v[p][1] = v[p][2] = v[p][3] = 0.f; }
#pragma unroll
      for (int e = 0; e < 4; ++e) ss = fmaf(v[p][e], v[p][e], ss);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float denom = normalize ? fmaxf(sqrtf(ss), 1e-12f) : 1.f;
    float s2 = 0.f;
#pragma unroll
    for (int p = 0; p < PACKS; ++p) {
      const int c4 = sub + p * LPR;
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float q = __fdiv_rn(v[p][e], denom);
        if constexpr (MODE == 1) {
          h[e] = __uint_as_float(__float_as_uint(q) & 0xffffe000u);
          l[e] = __uint_as_float(__float_as_uint(q - h[e]) & 0xffffe000u);
        } else if constexpr (MODE == 2) {
          h[e] = q;
          q = __bfloat162float(__float2bfloat16_rn(q));
        } else if constexpr (MODE == 3) {
          const float sv = q * kF16PlaneScale;
          h[e] = __half2float(__float2half_rn(sv));
          l[e] = sv - h[e];  // exact; rounded to fp16 by the store
        } else {
          h[e] = q;
        }
        s2 = fmaf(q, q, s2);
      }
      if (live && c4 < cv) {
        if constexpr (MODE == 2) {
          Pack<__nv_bfloat16, 4>::store(reinterpret_cast<__nv_bfloat16*>(xhat) + row * C + c4 * 4, h);
        } else if constexpr (MODE == 3) {
          const __half2 h01 = __floats2half2_rn(h[0], h[1]), h23 = __floats2half2_rn(h[2], h[3]);
          const __half2 l01 = __floats2half2_rn(l[0], l[1]), l23 = __floats2half2_rn(l[2], l[3]);
          uint2 hv, lv;
          hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
          lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(xhat) + row * C + c4 * 4) = hv;
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(lo) + row * C + c4 * 4) = lv;
        } else {
          Pack<float, 4>::store(xhat + row * C + c4 * 4, h);
          if constexpr (MODE == 1) Pack<float, 4>::store(lo + row * C + c4 * 4, l);
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (live && sub == 0) sq[row] = s2;
  }
}

// Lean fp32 / bf16 -> fp16 hi/lo plane producer for the f16x3 tensor-core kernels (mode 3, C % 4 == 0, C <= 1024).
// Same row ownership as the vectorised kernel above, but the warp never diverges (rows past the end are
// clamped and only their stores are guarded), the quotient x / denom is formed as one reciprocal per row
// plus a Newton correction per element (q0 = x * r; q = fma(fma(-q0, denom, x), r, q0), which is the
// correctly rounded quotient except in rare double-rounding cases), and the planes leave as packed
// 8-byte stores.  ~4x fewer instructions than the generic kernel, which was issue-bound.
template <typename T, int LPR, int PACKS>
__global__ void __launch_bounds__(256)
knn_normalize_f16_kernel(const T* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo,
                         float* __restrict__ sq, long long rows, int C, bool normalize) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  constexpr int RPW = 32 / LPR;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int cv = C / 4;
  for (long long rb = warp0 * RPW; rb < rows; rb += nwarps * RPW) {
    const long long row_raw = rb + lane / LPR;
    const bool live = row_raw < rows;
    const long long row = live ? row_raw : rows - 1;
    float4 v[PACKS];
    float ss = 0.f;
#pragma unroll
    for (int p = 0; p < PACKS; ++p) {
      const int c4 = sub + p * LPR;
      v[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 < cv) {
        if constexpr (std::is_same<T, float>::value) {
          v[p] = __ldg(reinterpret_cast<const float4*>(x + row * C) + c4);
        } else {  // bf16 rows (the generic kernel these took before was issue-bound: 95 us against 45)
          float t[4];
          Pack<T, 4>::load(x + row * C + c4 * 4, t);
          v[p] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      ss = fmaf(v[p].x, v[p].x, ss); ss = fmaf(v[p].y, v[p].y, ss);
      ss = fmaf(v[p].z, v[p].z, ss); ss = fmaf(v[p].w, v[p].w, ss);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float denom = normalize ? fmaxf(sqrtf(ss), 1e-12f) : 1.f;
    const float r = __frcp_rn(denom);
    float s2 = 0.f;
#pragma unroll
    for (int p = 0; p < PACKS; ++p) {
      const int c4 = sub + p * LPR;
      const float xin[4] = {v[p].x, v[p].y, v[p].z, v[p].w};
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float q0 = xin[e] * r;
        const float q = fmaf(fmaf(-q0, denom, xin[e]), r, q0);
        s2 = fmaf(q, q, s2);
        const float sv = q * kF16PlaneScale;
        h[e] = __half2float(__float2half_rn(sv));
        l[e] = sv - h[e];  // exact; rounded to fp16 by the store
      }
      if (live && c4 < cv) {
        const __half2 h01 = __floats2half2_rn(h[0], h[1]), h23 = __floats2half2_rn(h[2], h[3]);
        const __half2 l01 = __floats2half2_rn(l[0], l[1]), l23 = __floats2half2_rn(l[2], l[3]);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(hi + row * C + c4 * 4) = hv;
        *reinterpret_cast<uint2*>(lo + row * C + c4 * 4) = lv;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (live && sub == 0) sq[row] = s2;
  }
}

template <typename T>
static bool launch_normalize_f16(const T* xs, float* hi, float* lo, float* sq, long long rows, int C, bool normalize,
                                 cudaStream_t s) {
  const int cv = C / 4;
  __half* h = reinterpret_cast<__half*>(hi);
  __half* l = reinterpret_cast<__half*>(lo);
#define GRAFP_NORM16_CASE(LPR_, PACKS_)                                                              \
  {                                                                                                  \
    const long long need = (rows + (256 / LPR_) - 1) / (256 / LPR_);                                 \
    const long long cap = (long long)num_sms() * 8;                                                  \
    const int blocks = (int)(need < cap ? (need < 1 ? 1 : need) : cap);                              \
    knn_normalize_f16_kernel<T, LPR_, PACKS_><<<blocks, 256, 0, s>>>(xs, h, l, sq, rows, C, normalize); \
    return true;                                                                                     \
  }
  if (cv <= 4) GRAFP_NORM16_CASE(4, 1)
  if (cv <= 8) GRAFP_NORM16_CASE(8, 1)
  if (cv <= 16) GRAFP_NORM16_CASE(16, 1)
  if (cv <= 32) GRAFP_NORM16_CASE(32, 1)
  if (cv <= 64) GRAFP_NORM16_CASE(32, 2)
  if (cv <= 128) GRAFP_NORM16_CASE(32, 4)
  if (cv <= 256) GRAFP_NORM16_CASE(32, 8)
#undef GRAFP_NORM16_CASE
  return false;
}

template <typename T, int MODE>
static bool launch_normalize_vec(const T* xs, float* xhat, float* lo, float* sq, long long rows, int C, bool normalize,
                                 int blocks, cudaStream_t s) {
  const int cv = C / 4;
#define GRAFP_NORM_CASE(LPR_, PACKS_)                                                                        \
  knn_normalize_vec_kernel<T, MODE, LPR_, PACKS_><<<blocks, 256, 0, s>>>(xs, xhat, lo, sq, rows, C, normalize); \
  return true;
  if (cv <= 4) { GRAFP_NORM_CASE(4, 1) }
  if (cv <= 8) { GRAFP_NORM_CASE(8, 1) }
  if (cv <= 16) { GRAFP_NORM_CASE(16, 1) }
  if (cv <= 32) { GRAFP_NORM_CASE(32, 1) }
  if (cv <= 64) { GRAFP_NORM_CASE(32, 2) }
  if (cv <= 128) { GRAFP_NORM_CASE(32, 4) }
  if (cv <= 256) { GRAFP_NORM_CASE(32, 8) }
#undef GRAFP_NORM_CASE
  return false;
}

template <typename T>
int launch_knn_normalize(const void* x, float* xhat, float* lo, float* sq, long long rows, int C, int mode,
                         bool normalize, cudaStream_t s) {
  const int threads = 256;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const T* xs = static_cast<const T*>(x);
  const bool vec = (C % 4 == 0) && aligned16(x) && aligned16(xhat) && aligned16(lo);
  if (vec && mode == 3) {
    if (launch_normalize_f16<T>(xs, xhat, lo, sq, rows, C, normalize, s)) return check_launch("knn_normalize_f16");
  }
  if (vec) {
    bool done = false;
    if (mode == 0) done = launch_normalize_vec<T, 0>(xs, xhat, lo, sq, rows, C, normalize, (int)blocks, s);
    else if (mode == 1) done = launch_normalize_vec<T, 1>(xs, xhat, lo, sq, rows, C, normalize, (int)blocks, s);
    else if (mode == 3) done = launch_normalize_vec<T, 3>(xs, xhat, lo, sq, rows, C, normalize, (int)blocks, s);
    else done = launch_normalize_vec<T, 2>(xs, xhat, lo, sq, rows, C, normalize, (int)blocks, s);
    if (done) return check_launch("knn_normalize");
  }
  if (mode == 0) knn_normalize_kernel<T, 0><<<(int)blocks, threads, 0, s>>>(xs, xhat, lo, sq, rows, C, normalize);
  else if (mode == 1) knn_normalize_kernel<T, 1><<<(int)blocks, threads, 0, s>>>(xs, xhat, lo, sq, rows, C, normalize);
  else if (mode == 3) knn_normalize_kernel<T, 3><<<(int)blocks, threads, 0, s>>>(xs, xhat, lo, sq, rows, C, normalize);
  else knn_normalize_kernel<T, 2><<<(int)blocks, threads, 0, s>>>(xs, xhat, lo, sq, rows, C, normalize);
  return check_launch("knn_normalize");
}
__global__ void __launch_bounds__(256) fill_f32_kernel(float* __restrict__ p, long long n, float v) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) p[i] = v;
}
int launch_fill_f32(float* p, long long n, float v, cudaStream_t s) {
  fill_f32_kernel<<<grid_for(n, 256, 4), 256, 0, s>>>(p, n, v);
  return check_launch("fill_f32");
}

template int launch_knn_normalize<float>(const void*, float*, float*, float*, long long, int, int, bool, cudaStream_t);
template int launch_knn_normalize<__nv_bfloat16>(const void*, float*, float*, float*, long long, int, int, bool, cudaStream_t);

// ------------------------------------------------------------------------------------
// exact fp32 Gram + fused top-K
//   CTA = 64 query rows of one segment, 256 threads, 4x4 register micro-tiles.
//   For every 64-key tile: accumulate x_hat . y_hat over C in chunks of 16 through shared
//   memory, form D = (|x|^2 + (-2 s)) + |y|^2 (+ relpos) in that order (torch_edge.py:16-18),
//   park the tile in shared memory and let one thread per query row merge it into that
//   row's sorted top-K list (ties: lower key id first).
// ------------------------------------------------------------------------------------
constexpr int TQ = 64, TK = 64, TC = 16;

__global__ void __launch_bounds__(256)
knn_simt_kernel(const float* __restrict__ xh, const float* __restrict__ xsq, const float* __restrict__ yh,
                const float* __restrict__ ysq, const float* __restrict__ relpos, long long* __restrict__ nn_idx,
                int* __restrict__ nn_idx32, int N, int M, int C, int K, int k_out, int stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);         // [TC][TQ]
  float* Bs = As + TC * TQ;                               // [TC][TK]
  float* Ds = Bs + TC * TK;                               // [TQ][TK + 1]
  float* Ld = Ds + TQ * (TK + 1);                         // [TQ][K] distances, ascending
  int* Li = reinterpret_cast<int*>(Ld + TQ * K);          // [TQ][K] key ids

  const int b = blockIdx.y;
  const int q0 = blockIdx.x * TQ;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const float* xb = xh + (long long)b * N * C;
  const float* yb = yh + (long long)b * M * C;

  for (int i = tid; i < TQ * K; i += 256) { Ld[i] = INFINITY; Li[i] = 0; }

  const int lrow = tid >> 2;        // 0..63: tile row loaded by this thread
  const int lc = (tid & 3) * 4;     // 0,4,8,12: first of its 4 channels
  const bool vec_ok = (C % 4 == 0);

  for (int k0 = 0; k0 < M; k0 += TK) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int c0 = 0; c0 < C; c0 += TC) {
      float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
      const int qa = q0 + lrow, kb = k0 + lrow, cc = c0 + lc;
      if (vec_ok && cc + 3 < C) {
        if (qa < N) { const float4 t = *reinterpret_cast<const float4*>(xb + (long long)qa * C + cc); av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w; }
        if (kb < M) { const float4 t = *reinterpret_cast<const float4*>(yb + (long long)kb * C + cc); bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w; }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (qa < N && cc + e < C) av[e] = xb[(long long)qa * C + cc + e];
          if (kb < M && cc + e < C) bv[e] = yb[(long long)kb * C + cc + e];
        }
      }
      __syncthreads();  // previous chunk fully consumed
#pragma unroll
      for (int e = 0; e < 4; ++e) { As[(lc + e) * TQ + lrow] = av[e]; Bs[(lc + e) * TK + lrow] = bv[e]; }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        const float4 a = *reinterpret_cast<const float4*>(As + c * TQ + ty * 4);
        const float4 bq = *reinterpret_cast<const float4*>(Bs + c * TK + tx * 4);
        const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
    }

    // distances of this tile -> shared
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = q0 + ty * 4 + i;
      const float sqq = (q < N) ? xsq[(long long)b * N + q] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = k0 + tx * 4 + j;
        float d = INFINITY;
        if (q < N && key < M) {
          d = __fadd_rn(fmaf(-2.f, acc[i][j], sqq), ysq[(long long)b * M + key]);
          if (relpos != nullptr) d = __fadd_rn(d, relpos[(long long)q * M + key]);
        }
        Ds[(ty * 4 + i) * (TK + 1) + tx * 4 + j] = d;
      }
    }
    __syncthreads();

    if (tid < TQ) {
      float* ld = Ld + tid * K;
      int* li = Li + tid * K;
      float worst = ld[K - 1];
      const int lim = min(TK, M - k0);
      for (int j = 0; j < lim; ++j) {
        const float d = Ds[tid * (TK + 1) + j];
        if (d < worst) {
          int p = K - 1;
          while (p > 0 && ld[p - 1] > d) { ld[p] = ld[p - 1]; li[p] = li[p - 1]; --p; }
          ld[p] = d; li[p] = k0 + j;
          worst = ld[K - 1];
        }
      }
    }
    // the next tile's first __syncthreads() orders the Ds reads before its rewrite
  }
  __syncthreads();

  // emit ranks 0, stride, 2*stride, ...
  for (int i = tid; i < TQ * k_out; i += 256) {
    const int r = i / k_out, j = i - r * k_out;
    const int q = q0 + r;
    if (q < N) {
      const int id = Li[r * K + j * stride];
      const long long o = ((long long)b * N + q) * k_out + j;
      nn_idx[o] = id;
      if (nn_idx32 != nullptr) nn_idx32[o] = id;
    }
  }
}

int launch_knn_simt(const float* xh, const float* xsq, const float* yh, const float* ysq, const float* relpos,
                    long long* nn_idx, int* nn_idx32, int B, int N, int M, int C, int K, int k_out, int stride,
                    cudaStream_t s) {
  const size_t smem = sizeof(float) * (TC * TQ + TC * TK + TQ * (TK + 1)) + (sizeof(float) + sizeof(int)) * TQ * K;
  static DeviceOnce once;
  if (once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(knn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(knn_simt): %s", cudaGetErrorString(e)); return (int)e; }
    once.mark();
  }
  dim3 grid((N + TQ - 1) / TQ, B);
  knn_simt_kernel<<<grid, 256, smem, s>>>(xh, xsq, yh, ysq, relpos, nn_idx, nn_idx32, N, M, C, K, k_out, stride);
  return check_launch("knn_simt");
}

}  // namespace grafp
