// Downsample block (graph_encoder.py:16-28: Conv2d(3x3, stride 2, padding 1) + BatchNorm over a (B, C, N, 1) node list).
// The image is one pixel wide, so the layer is a 3-tap, stride-2 convolution along the node axis:
//   out[n'] = W[:, :, 0, 1] x[2n' - 1] + W[:, :, 1, 1] x[2n'] + W[:, :, 2, 1] x[2n' + 1]       (x[-1] = 0)
// i.e. a 1x1 convolution with 3C input channels over "tap rows" taps[b][n'] = (x[2n'-1], x[2n'], x[2n'+1]).  In row
// layout those three rows are 3C CONSECUTIVE elements of x, so building the tap rows is one shifted copy and its
// backward one pass that folds the overlapping first / last thirds back:
//   dx[2n']     = dtaps[n'][C : 2C]
//   dx[2n' + 1] = dtaps[n'][2C : 3C] + dtaps[n' + 1][0 : C]                                   (second term 0 at the end)
// Both kernels move 16 bytes per thread and iteration, coalesced, 2.5 tensor passes each - in place of the ten PyTorch
// ops (pad, slice, cat, clone and their backwards: ~4 ms of a 105 ms training step) that built the same rows.
#include <cuda_bf16.h>

#include "common.cuh"

namespace grafp {
namespace {

constexpr int kThreads = 256;

// items are 16-byte packs; a tap row has 3 * cv of them, a node row cv
__global__ void __launch_bounds__(kThreads)
taps_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ taps, long long total, int half_n, int cv) {
  const long long stride = (long long)gridDim.x * kThreads;
  const int row_items = 3 * cv;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    const long long row = i / row_items;            // (b, n')
    const int q = (int)(i - row * row_items);
    const int np = (int)(row % half_n);
    // tap row (b, n') starts one node row before x[b][2n']
    const long long src = (2 * row - 1) * cv + q;   // (b * N + 2n' - 1) * cv + q  with  N = 2 * half_n
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (np != 0 || q >= cv) v = __ldg(x + src);
    taps[i] = v;
  }
}

template <typename T>
__device__ __forceinline__ uint4 add16(const uint4& a, const uint4& b);
template <>
__device__ __forceinline__ uint4 add16<float>(const uint4& a, const uint4& b) {
  return make_uint4(__float_as_uint(__uint_as_float(a.x) + __uint_as_float(b.x)), __float_as_uint(__uint_as_float(a.y) + __uint_as_float(b.y)),
                    __float_as_uint(__uint_as_float(a.z) + __uint_as_float(b.z)), __float_as_uint(__uint_as_float(a.w) + __uint_as_float(b.w)));
}
template <>
__device__ __forceinline__ uint4 add16<__nv_bfloat16>(const uint4& a, const uint4& b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int e = 0; e < 4; ++e) pr[e] = __hadd2(pa[e], pb[e]);
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
taps_bwd_kernel(const uint4* __restrict__ dtaps, uint4* __restrict__ dx, long long total, int half_n, int cv) {
  const long long stride = (long long)gridDim.x * kThreads;
  const int pair_items = 2 * cv;  // node rows 2n', 2n' + 1
  const int row_items = 3 * cv;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    const long long row = i / pair_items;           // (b, n')
    const int q = (int)(i - row * pair_items);      // [0, cv): node 2n', [cv, 2cv): node 2n' + 1
    const int np = (int)(row % half_n);
    uint4 v = __ldg(dtaps + row * row_items + cv + q);
    if (q >= cv && np + 1 < half_n) v = add16<T>(v, __ldg(dtaps + (row + 1) * row_items + (q - cv)));
    dx[i] = v;
  }
}

int grid_for(long long total) {
  long long g = (total + kThreads - 1) / kThreads;
  const long long cap = (long long)num_sms() * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

bool downsample_taps_supported(int N, int C, int dtype) {
  const int es = dtype == GRAFP_F32 ? 4 : 2;
  return (dtype == GRAFP_F32 || dtype == GRAFP_BF16) && N >= 2 && N % 2 == 0 && C >= 1 && (C * es) % 16 == 0;
}

int launch_downsample_taps_fwd(const void* x, void* taps, int B, int N, int C, int dtype, cudaStream_t s) {
  const int cv = C * (dtype == GRAFP_F32 ? 4 : 2) / 16;
  const long long total = (long long)B * (N / 2) * 3 * cv;
  taps_fwd_kernel<<<grid_for(total), kThreads, 0, s>>>(static_cast<const uint4*>(x), static_cast<uint4*>(taps), total, N / 2, cv);
  return check_launch("downsample_taps_fwd");
}

int launch_downsample_taps_bwd(const void* dtaps, void* dx, int B, int N, int C, int dtype, cudaStream_t s) {
  const int cv = C * (dtype == GRAFP_F32 ? 4 : 2) / 16;
  const long long total = (long long)B * N * cv;
  if (dtype == GRAFP_F32)
    taps_bwd_kernel<float><<<grid_for(total), kThreads, 0, s>>>(static_cast<const uint4*>(dtaps), static_cast<uint4*>(dx), total, N / 2, cv);
  else
    taps_bwd_kernel<__nv_bfloat16><<<grid_for(total), kThreads, 0, s>>>(static_cast<const uint4*>(dtaps), static_cast<uint4*>(dx), total, N / 2, cv);
  return check_launch("downsample_taps_bwd");
}

}  // namespace grafp
