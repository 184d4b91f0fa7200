// Peak point-cloud front end as ONE kernel (reference: peak_extractor.py:56-82): min-max normalise the
// log-mel segment, append the time / frequency position ramps, Conv2d(3 -> F, kh x kw, stride (s, 1), "same" padding) +
// ReLU, and write the result directly as node rows (B, N = Ho * W, F) - the channels-last layout the encoder's stem
// consumes - instead of amin / amax / sub / div / cat / cuDNN conv / ReLU / reshape + the layout copy.
// One CTA per segment: the (3, H + 2 ph, W + 2 pw) zero-padded input planes live in shared memory (the two ramp
// planes are generated, never read), a thread computes all F channels of its nodes.
// Backward: only the convolution's weight and bias have gradients (the spectrogram does not); per-segment partial
// sums (recomputing the normalised planes, ReLU mask from the saved output) + a fixed-order reduction over segments.
#include "common.cuh"

namespace grafp {
namespace {

constexpr int kPeThreads = 256;
constexpr int kPeF = 8;  // filters (config n_filters)

struct PeGeom {
  int H, W, kh, kw, sh, ph, pw, Ho, HP, WP;
};

__host__ __device__ inline PeGeom pe_geom(int H, int W, int kh, int kw, int sh) {
  PeGeom g;
  g.H = H; g.W = W; g.kh = kh; g.kw = kw; g.sh = sh;
  g.ph = kh / 2; g.pw = kw / 2;
  g.Ho = (H + 2 * g.ph - kh) / sh + 1;
  g.HP = H + 2 * g.ph; g.WP = W + 2 * g.pw;
  return g;
}

// normalised spectrogram + ramps into shared memory planes [3][HP][WP] (channel order of the reference's cat: T, F, peaks)
__device__ __forceinline__ void pe_fill_planes(float* planes, float* red, const float* __restrict__ spec, const PeGeom& g) {
  const int n = g.H * g.W;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += kPeThreads) { const float v = __ldg(spec + i); mn = fminf(mn, v); mx = fmaxf(mx, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = mn; red[8 + (threadIdx.x >> 5)] = mx; }
  __syncthreads();
  mn = red[0]; mx = red[8];
#pragma unroll
  for (int w = 1; w < kPeThreads / 32; ++w) { mn = fminf(mn, red[w]); mx = fmaxf(mx, red[8 + w]); }
  const float range = mx - mn;
  const int plane = g.HP * g.WP;
  for (int i = threadIdx.x; i < 3 * plane; i += kPeThreads) planes[i] = 0.f;
  __syncthreads();
  // torch.linspace(0, 1, steps): start + i * step for the first half, end - (steps - 1 - i) * step for the second
  const float tstep = g.W > 1 ? 1.f / (float)(g.W - 1) : 0.f, fstep = g.H > 1 ? 1.f / (float)(g.H - 1) : 0.f;
  for (int i = threadIdx.x; i < n; i += kPeThreads) {
    const int h = i / g.W, w = i - h * g.W;
    const int at = (h + g.ph) * g.WP + (w + g.pw);
    planes[at] = (w < g.W / 2) ? (float)w * tstep : 1.f - (float)(g.W - 1 - w) * tstep;
    planes[plane + at] = (h < g.H / 2) ? (float)h * fstep : 1.f - (float)(g.H - 1 - h) * fstep;
    planes[2 * plane + at] = (__ldg(spec + i) - mn) / range;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kPeThreads)
peak_extract_fwd_kernel(const float* __restrict__ spec, const float* __restrict__ weight, const float* __restrict__ bias,
                        float* __restrict__ out, int H, int W, int kh, int kw, int sh) {
  extern __shared__ float pe_smem[];
  const PeGeom g = pe_geom(H, W, kh, kw, sh);
  const int plane = g.HP * g.WP, taps = 3 * kh * kw;
  float* planes = pe_smem;                 // [3][HP][WP]
  float* wt = planes + 3 * plane;          // [3 * kh * kw][F] (filters contiguous: two 128-bit broadcast loads per tap)
  float* red = wt + taps * kPeF;           // [16]
  const long long b = blockIdx.x;
  for (int i = threadIdx.x; i < taps * kPeF; i += kPeThreads) {
    const int t = i / kPeF, f = i - t * kPeF;
    wt[i] = __ldg(weight + (size_t)f * taps + t);
  }
  pe_fill_planes(planes, red, spec + b * (long long)H * W, g);
  float bs[kPeF];
#pragma unroll
  for (int f = 0; f < kPeF; ++f) bs[f] = __ldg(bias + f);
  const int nodes = g.Ho * W;
  float* ob = out + b * (long long)nodes * kPeF;
  for (int node = threadIdx.x; node < nodes; node += kPeThreads) {
    const int oh = node / W, ow = node - oh * W;
    float acc[kPeF];
#pragma unroll
    for (int f = 0; f < kPeF; ++f) acc[f] = bs[f];
    for (int ci = 0; ci < 3; ++ci) {
      for (int r = 0; r < kh; ++r) {
        const float* row = planes + ci * plane + (oh * sh + r) * g.WP + ow;
        const float* wr = wt + ((ci * kh + r) * kw) * kPeF;
        for (int c = 0; c < kw; ++c) {
          const float v = row[c];
          const float4 w0 = *reinterpret_cast<const float4*>(wr + c * kPeF);
          const float4 w1 = *reinterpret_cast<const float4*>(wr + c * kPeF + 4);
          acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
      }
    }
    float4* o = reinterpret_cast<float4*>(ob + (long long)node * kPeF);
    o[0] = make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
    o[1] = make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
  }
}

// per-segment partial gradients: partial[b][t * F + f] = sum_nodes gz[node][f] * in[tap t of node], partial[b][taps * F + f] = sum gz
__global__ void __launch_bounds__(kPeThreads)
peak_extract_bwd_kernel(const float* __restrict__ spec, const float* __restrict__ out, const float* __restrict__ grad_out,
                        float* __restrict__ partial, int H, int W, int kh, int kw, int sh) {
  extern __shared__ float pe_smem[];
  const PeGeom g = pe_geom(H, W, kh, kw, sh);
  const int plane = g.HP * g.WP, taps = 3 * kh * kw, nodes = g.Ho * W;
  float* planes = pe_smem;                 // [3][HP][WP]
  float* gz = planes + 3 * plane;          // [nodes][F] upstream gradient through the ReLU
  float* red = gz + nodes * kPeF;          // [16]
  const long long b = blockIdx.x;
  pe_fill_planes(planes, red, spec + b * (long long)H * W, g);
  const float* ob = out + b * (long long)nodes * kPeF;
  const float* gb = grad_out + b * (long long)nodes * kPeF;
  for (int i = threadIdx.x; i < nodes * kPeF; i += kPeThreads) gz[i] = __ldg(ob + i) > 0.f ? __ldg(gb + i) : 0.f;
  __syncthreads();
  float* pb = partial + b * (long long)(taps + 1) * kPeF;
  // one (tap, 4-filter half) item per thread and trip: taps * 2 items
  for (int item = threadIdx.x; item < (taps + 1) * 2; item += kPeThreads) {
    const int t = item >> 1, half = (item & 1) * 4;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (t < taps) {
      const int ci = t / (kh * kw), r = (t / kw) % kh, c = t % kw;
      const float* base = planes + ci * plane + r * g.WP + c;
      for (int oh = 0; oh < g.Ho; ++oh) {
        const float* row = base + oh * sh * g.WP;
        const float* gr = gz + (oh * W) * kPeF + half;
        for (int ow = 0; ow < W; ++ow) {
          const float v = row[ow];
          const float4 q = *reinterpret_cast<const float4*>(gr + ow * kPeF);
          a0 = fmaf(v, q.x, a0); a1 = fmaf(v, q.y, a1); a2 = fmaf(v, q.z, a2); a3 = fmaf(v, q.w, a3);
        }
      }
    } else {  // bias: plain sum of gz
      for (int node = 0; node < nodes; ++node) {
        const float4 q = *reinterpret_cast<const float4*>(gz + node * kPeF + half);
        a0 += q.x; a1 += q.y; a2 += q.z; a3 += q.w;
      }
    }
    *reinterpret_cast<float4*>(pb + t * kPeF + half) = make_float4(a0, a1, a2, a3);
  }
}

// dweight[f][t] = sum_b partial[b][t][f], dbias[f] = sum_b partial[b][taps][f]; fixed order, double accumulation
__global__ void __launch_bounds__(256)
peak_extract_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dweight, float* __restrict__ dbias, int B, int taps) {
  const int i = blockIdx.x * 256 + threadIdx.x;  // index into [taps + 1][F]
  if (i >= (taps + 1) * kPeF) return;
  double acc = 0.0;
  const long long stride = (long long)(taps + 1) * kPeF;
  int bb = 0;
  for (; bb + 3 < B; bb += 4) {
    const float v0 = partial[bb * stride + i], v1 = partial[(bb + 1) * stride + i];
    const float v2 = partial[(bb + 2) * stride + i], v3 = partial[(bb + 3) * stride + i];
    acc += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
  }
  for (; bb < B; ++bb) acc += (double)partial[bb * stride + i];
  const int t = i / kPeF, f = i - t * kPeF;
  if (t < taps) dweight[(size_t)f * taps + t] = (float)acc;
  else dbias[f] = (float)acc;
}

}  // namespace

bool peak_extract_supported(int H, int W, int F, int kh, int kw, int sh) {
  if (F != kPeF || H < 1 || W < 1 || kh < 1 || kw < 1 || (kh & 1) == 0 || (kw & 1) == 0 || sh < 1) return false;
  const PeGeom g = pe_geom(H, W, kh, kw, sh);
  const size_t fwd = ((size_t)3 * g.HP * g.WP + (size_t)3 * kh * kw * kPeF + 16) * sizeof(float);
  const size_t bwd = ((size_t)3 * g.HP * g.WP + (size_t)g.Ho * W * kPeF + 16) * sizeof(float);
  return fwd <= 200 * 1024 && bwd <= 200 * 1024;
}

size_t peak_extract_workspace_bytes(int B, int kh, int kw) { return (size_t)B * (3 * kh * kw + 1) * kPeF * sizeof(float); }

int launch_peak_extract_fwd(const float* spec, const float* weight, const float* bias, float* out, int B, int H, int W, int kh,
                            int kw, int sh, cudaStream_t s) {
  const PeGeom g = pe_geom(H, W, kh, kw, sh);
  const size_t smem = ((size_t)3 * g.HP * g.WP + (size_t)3 * kh * kw * kPeF + 16) * sizeof(float);
  static DeviceOnce once;
  if (once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(peak_extract_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(peak_extract_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(peak_extract): %s", cudaGetErrorString(e)); return (int)e; }
    once.mark();
  }
  peak_extract_fwd_kernel<<<B, kPeThreads, smem, s>>>(spec, weight, bias, out, H, W, kh, kw, sh);
  return check_launch("peak_extract_fwd");
}

int launch_peak_extract_bwd(const float* spec, const float* out, const float* grad_out, float* partial, float* dweight,
                            float* dbias, int B, int H, int W, int kh, int kw, int sh, cudaStream_t s) {
  const PeGeom g = pe_geom(H, W, kh, kw, sh);
  const size_t smem = ((size_t)3 * g.HP * g.WP + (size_t)g.Ho * W * kPeF + 16) * sizeof(float);
  static DeviceOnce once;
  if (once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(peak_extract_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(peak_extract_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(peak_extract): %s", cudaGetErrorString(e)); return (int)e; }
    once.mark();
  }
  peak_extract_bwd_kernel<<<B, kPeThreads, smem, s>>>(spec, out, grad_out, partial, H, W, kh, kw, sh);
  const int taps = 3 * kh * kw;
  peak_extract_reduce_kernel<<<((taps + 1) * kPeF + 255) / 256, 256, 0, s>>>(partial, dweight, dbias, B, taps);
  return check_launch("peak_extract_bwd");
}

}  // namespace grafp
