// NT-Xent contrastive loss, forward and backward, without ever materialising the (2B x 2B) similarity matrix
// (reference: simclr/ntxent.py:17-29 - a Python loop over the 2B rows of z z^T / tau, each taking the log-softmax of
// the row without its diagonal entry and picking the partner's entry).
//
//   z      (n2, d) fp32: the two views interleaved, rows 2m and 2m + 1 are partners (ntxent.py:18: stack(dim=1).view)
//   a_ij   = z_i . z_j / tau                    (i != j)
//   loss   = 1/n2 sum_i [ lse_i - a_{i, i^1} ],  lse_i = log sum_{j != i} exp(a_ij)
//   dz_i   = g / (n2 tau) [ sum_{j != i} (P_ij + P_ji) z_j - 2 z_{i^1} ],  P_ij = exp(a_ij - lse_i)
//
// A CTA owns 32 rows of z and streams all n2 rows through shared memory in tiles of 32: a 32 x 32 block of logits per
// tile (2 x 2 per thread, fp32 FMA - at tau = 0.05 the logits span +-20 and feed an exp, so no reduced precision), an
// online log-sum-exp per row in the forward, and in the backward the weights w_ij = P_ij + P_ji parked in shared
// memory and applied to the same tile of z (the similarity is symmetric, so one logit block serves both terms).
// Compute-bound on the fp32 pipe: 2 n2^2 d flop forward, 4 n2^2 d backward (17 / 34 GFLOP at the global batch of 4096).
#include "common.cuh"

namespace grafp {
namespace {

constexpr int kNtTile = 32;       // rows of z per CTA and per streamed tile
constexpr int kNtThreads = 256;   // 16 x 16 threads, 2 x 2 logits each
constexpr int kNtMaxD = 256;

// rows [row0, row0 + 32) of z into shared memory as [32][d + 1] (padded: conflict-free column walks); rows >= n2 are 0
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ z, int row0, int n2, int d) {
  const int dv = d >> 2;
  for (int i = threadIdx.x; i < kNtTile * dv; i += kNtThreads) {
    const int r = i / dv, c4 = i - r * dv;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n2) v = __ldg(reinterpret_cast<const float4*>(z + (size_t)(row0 + r) * d) + c4);
    float* p = dst + r * (d + 1) + 4 * c4;
    p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
  }
}

// this thread's 2 x 2 logits of the (zi tile) x (zj tile)^T block, scaled by 1 / tau
__device__ __forceinline__ void logits_2x2(const float* zi, const float* zj, int ty, int tx, int d, float inv_tau, float (&a)[2][2]) {
  const float* i0 = zi + (2 * ty) * (d + 1);
  const float* i1 = i0 + (d + 1);
  const float* j0 = zj + (2 * tx) * (d + 1);
  const float* j1 = j0 + (d + 1);
  float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll 8
  for (int k = 0; k < d; ++k) {
    const float x0 = i0[k], x1 = i1[k], y0 = j0[k], y1 = j1[k];
    s00 = fmaf(x0, y0, s00); s01 = fmaf(x0, y1, s01);
    s10 = fmaf(x1, y0, s10); s11 = fmaf(x1, y1, s11);
  }
  a[0][0] = s00 * inv_tau; a[0][1] = s01 * inv_tau; a[1][0] = s10 * inv_tau; a[1][1] = s11 * inv_tau;
}

__global__ void __launch_bounds__(kNtThreads)
ntxent_fwd_kernel(const float* __restrict__ z, float* __restrict__ lse, float* __restrict__ row_loss, int n2, int d,
                  float inv_tau, int row_lo, int row_hi) {
  extern __shared__ float nt_smem[];
  float* zi = nt_smem;
  float* zj = nt_smem + kNtTile * (d + 1);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row0 = row_lo + blockIdx.x * kNtTile;   // rows [row_lo, row_hi) of the (global) batch: a rank's own anchors
  load_tile(zi, z, row0, n2, d);
  float m[2] = {-INFINITY, -INFINITY}, s[2] = {0.f, 0.f}, pos[2] = {0.f, 0.f};
  for (int col0 = 0; col0 < n2; col0 += kNtTile) {
    __syncthreads();  // the previous tile is consumed (and zi is loaded, first trip)
    load_tile(zj, z, col0, n2, d);
    __syncthreads();
    float a[2][2];
    logits_2x2(zi, zj, ty, tx, d, inv_tau, a);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int i = row0 + 2 * ty + p;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int j = col0 + 2 * tx + q;
        if (j >= n2 || j == i) continue;        // the diagonal is removed before the softmax (ntxent.py:24)
        if (j == (i ^ 1)) pos[p] = a[p][q];
        const float v = a[p][q];
        if (v > m[p]) { s[p] = s[p] * __expf(m[p] - v) + 1.f; m[p] = v; }
        else s[p] += __expf(v - m[p]);
      }
    }
  }
  // combine the 16 threads (tx) that share a row pair: they are 16 consecutive lanes
#pragma unroll
  for (int p = 0; p < 2; ++p) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m[p], o);
      const float so = __shfl_xor_sync(0xffffffffu, s[p], o);
      const float po = __shfl_xor_sync(0xffffffffu, pos[p], o);
      const float mn = fmaxf(m[p], mo);
      const float sa = (m[p] == -INFINITY) ? 0.f : s[p] * __expf(m[p] - mn);
      const float sb = (mo == -INFINITY) ? 0.f : so * __expf(mo - mn);
      s[p] = sa + sb; m[p] = mn; pos[p] += po;  // exactly one lane holds the partner's logit, the others 0
    }
    const int i = row0 + 2 * ty + p;
    if (tx == 0 && i < row_hi) {
      const float l = m[p] + __logf(s[p]);
      lse[i] = l;
      row_loss[i] = l - pos[p];
    }
  }
}

// loss = sum(row_loss[row_lo : row_hi]) / n2 in a fixed order (one CTA)
__global__ void __launch_bounds__(256)
ntxent_loss_reduce_kernel(const float* __restrict__ row_loss, float* __restrict__ loss, int n2, int row_lo, int row_hi) {
  __shared__ double sm[256];
  double acc = 0.0;
  for (int i = row_lo + threadIdx.x; i < row_hi; i += 256) acc += (double)row_loss[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if ((int)threadIdx.x < h) sm[threadIdx.x] += sm[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(sm[0] / (double)n2);
}

__global__ void __launch_bounds__(kNtThreads)
ntxent_bwd_kernel(const float* __restrict__ z, const float* __restrict__ lse, const float* __restrict__ grad_loss,
                  float* __restrict__ dz, int n2, int d, float inv_tau, int row_lo, int row_hi, float grad_scale) {
  extern __shared__ float nt_smem[];
  float* zi = nt_smem;
  float* zj = zi + kNtTile * (d + 1);
  float* w = zj + kNtTile * (d + 1);   // [32][33] weights of the current tile
  float* lse_j = w + kNtTile * (kNtTile + 1);  // [32]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row0 = row_lo + blockIdx.x * kNtTile;
  load_tile(zi, z, row0, n2, d);
  float lse_i[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) lse_i[p] = (row0 + 2 * ty + p < n2) ? __ldg(lse + row0 + 2 * ty + p) : 0.f;
  // accumulators: thread t owns row r = t / 8 (0..31) and the 4-channel packs c4 = t % 8 + 8 q of that row
  const int ar = threadIdx.x >> 3, ac = threadIdx.x & 7;
  const int npk = (d >> 2);
  float4 acc[kNtMaxD / 32];
#pragma unroll
  for (int q = 0; q < kNtMaxD / 32; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int col0 = 0; col0 < n2; col0 += kNtTile) {
    __syncthreads();
    load_tile(zj, z, col0, n2, d);
    if (threadIdx.x < kNtTile) lse_j[threadIdx.x] = (col0 + (int)threadIdx.x < n2) ? __ldg(lse + col0 + threadIdx.x) : 0.f;
    __syncthreads();
    float a[2][2];
    logits_2x2(zi, zj, ty, tx, d, inv_tau, a);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int i = row0 + 2 * ty + p;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int jl = 2 * tx + q, j = col0 + jl;
        float wij = 0.f;
        if (i < n2 && j < n2 && j != i) {
          wij = __expf(a[p][q] - lse_i[p]) + __expf(a[p][q] - lse_j[jl]);  // P_ij + P_ji (a is symmetric)
          if (j == (i ^ 1)) wij -= 2.f;                                     // the partner's -1 from either row
        }
        w[(2 * ty + p) * (kNtTile + 1) + jl] = wij;
      }
    }
    __syncthreads();
    // dz_i += sum_j w_ij z_j over this tile
    const float* wr = w + ar * (kNtTile + 1);
#pragma unroll 4
    for (int jl = 0; jl < kNtTile; ++jl) {
      const float wv = wr[jl];
      const float* zr = zj + jl * (d + 1);
#pragma unroll
      for (int q = 0; q < kNtMaxD / 32; ++q) {
        const int c4 = ac + 8 * q;
        if (c4 < npk) {
          acc[q].x = fmaf(wv, zr[4 * c4], acc[q].x); acc[q].y = fmaf(wv, zr[4 * c4 + 1], acc[q].y);
          acc[q].z = fmaf(wv, zr[4 * c4 + 2], acc[q].z); acc[q].w = fmaf(wv, zr[4 * c4 + 3], acc[q].w);
        }
      }
    }
  }
  const int i = row0 + ar;
  if (i < row_hi) {
    const float scale = __ldg(grad_loss) * grad_scale * inv_tau / (float)n2;
#pragma unroll
    for (int q = 0; q < kNtMaxD / 32; ++q) {
      const int c4 = ac + 8 * q;
      if (c4 < npk) {
        reinterpret_cast<float4*>(dz + (size_t)(i - row_lo) * d)[c4] =   // dz holds the rows [row_lo, row_hi)
            make_float4(acc[q].x * scale, acc[q].y * scale, acc[q].z * scale, acc[q].w * scale);
      }
    }
  }
}

}  // namespace

bool ntxent_supported(int n2, int d) { return n2 >= 2 && n2 % 2 == 0 && d >= 4 && d % 4 == 0 && d <= kNtMaxD; }

int launch_ntxent_fwd(const float* z, float* lse, float* row_loss, float* loss, int n2, int d, float inv_tau, int row_lo,
                      int row_hi, cudaStream_t s) {
  const size_t smem = (size_t)2 * kNtTile * (d + 1) * sizeof(float);
  static DeviceOnce once;
  if (once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(ntxent_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ntxent_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(ntxent): %s", cudaGetErrorString(e)); return (int)e; }
    once.mark();
  }
  ntxent_fwd_kernel<<<(row_hi - row_lo + kNtTile - 1) / kNtTile, kNtThreads, smem, s>>>(z, lse, row_loss, n2, d, inv_tau, row_lo, row_hi);
  ntxent_loss_reduce_kernel<<<1, 256, 0, s>>>(row_loss, loss, n2, row_lo, row_hi);
  return check_launch("ntxent_fwd");
}

int launch_ntxent_bwd(const float* z, const float* lse, const float* grad_loss, float* dz, int n2, int d, float inv_tau,
                      int row_lo, int row_hi, float grad_scale, cudaStream_t s) {
  const size_t smem = ((size_t)2 * kNtTile * (d + 1) + kNtTile * (kNtTile + 1) + kNtTile) * sizeof(float);
  static DeviceOnce once;
  if (once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(ntxent_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(ntxent_bwd): %s", cudaGetErrorString(e)); return (int)e; }
    once.mark();
  }
  ntxent_bwd_kernel<<<(row_hi - row_lo + kNtTile - 1) / kNtTile, kNtThreads, smem, s>>>(z, lse, grad_loss, dz, n2, d, inv_tau, row_lo,
                                                                                       row_hi, grad_scale);
  return check_launch("ntxent_bwd");
}

}  // namespace grafp
