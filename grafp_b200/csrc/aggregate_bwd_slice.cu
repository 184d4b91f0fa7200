// K3, slice form (fp32, graphs whose centre of row n is n): the argmax-routed scatter backward of
// the max-relative aggregation (autograd of torch_vertex.py:21-32 / torch_nn.py:79-98) as a
// deterministic gather that runs entirely out of shared memory.
//
//   grad_x[n][c] = g[n][2c] - [winner(n,c) != n] g[n][2c+1] + sum over in-edges (m, j), m != n, of [argmax[m][c] == j] g[m][2c+1]
//
// The scatter is independent per channel, so the work unit is (segment, slice of Cs channels):
// the slice's grad_out columns of ALL N rows (N x 2Cs floats, <= 64 KB) and its argmax bytes are
// staged in shared memory, and every in-edge look-up is a shared-memory read instead of an L2
// round trip (the gather-over-L2 form moved ~4x the algorithmic bytes across L2 and sat at the
// L2 bandwidth; the atomic forms sat at the L2 reduction rate).  One persistent CTA per SM walks
// a contiguous range of units; every thread owns one (row, 8-channel) item of a unit and loads
// it straight from HBM into registers TWO units ahead (grad_out as 4 x 16 bytes, the 8 argmax
// bytes), so ~144 KB per SM are always in flight and HBM sees one continuous stream of exactly
// the algorithmic bytes: grad_out, argmax and the ids read once, grad_x written once.  Only the
// values other rows gather - g[.., 2c+1] and the argmax bytes - and the dense term go through
// shared memory (compact, double-buffered, one barrier per unit); the gather pass takes its rows
// in in-degree order so the lanes of a warp walk in-edge lists of the same length.  The reverse
// graph (who points at row n, through which neighbour slot) is rebuilt in shared memory whenever
// the CTA moves to a new segment (every C/Cs units), with the lists sorted so the summation
// order - and therefore the result - is bit-reproducible.  No atomics on global memory, no
// zero-fill, no workspace.
#include "common.cuh"

namespace grafp {
namespace {

constexpr int kSliceThreads = 1024;
constexpr size_t kSliceSmemLimit = 232448;  // 227 KB opt-in maximum per CTA

struct SliceLayout {
  size_t g1_bytes;    // one buffer of g[.., 2c+1] for the slice: N * Cs * 4 (the dense-term buffer has the same size)
  size_t am_bytes;    // one buffer of argmax bytes: N * Cs (padded to 16)
  size_t dd_at, am_at, off_at, cur_at, mask_at, rs_at, perm_at, wsum_at, total;
};

__host__ __device__ inline SliceLayout slice_layout(int N, int Cs, int k) {
  SliceLayout L;
  L.g1_bytes = (size_t)N * Cs * 4;
  L.am_bytes = ((size_t)N * Cs + 15) / 16 * 16;
  size_t at = 2 * L.g1_bytes;
  L.dd_at = at;   at += 2 * L.g1_bytes;
  L.am_at = at;   at += 2 * L.am_bytes;
  L.off_at = at;  at += ((size_t)(N + 1) * 4 + 15) / 16 * 16;
  L.cur_at = at;  at += ((size_t)N * 4 + 15) / 16 * 16;
  L.mask_at = at; at += ((size_t)N * 4 + 15) / 16 * 16;
  L.rs_at = at;   at += ((size_t)N * k * 4 + 15) / 16 * 16;
  L.perm_at = at; at += ((size_t)N * 2 + 15) / 16 * 16;
  L.wsum_at = at; at += 2 * 32 * 4;
  L.total = at;
  return L;
}

// one unit's share of a thread: one (row, 8-channel) item = 16 interleaved grad_out floats + 8 argmax bytes
struct SliceRegs {
  float4 g[4];
  uint2 am;
};

template <bool I64>
__global__ void __launch_bounds__(kSliceThreads, 1)
mr_bwd_slice_kernel(const float* __restrict__ g, const uint8_t* __restrict__ argmax, const void* __restrict__ nbr,
                    float* __restrict__ grad_x, int N, int C, int k, int cs_shift, int slices, long long total_units) {
  extern __shared__ __align__(128) unsigned char sl_smem[];
  const int Cs = 1 << cs_shift;
  const SliceLayout L = slice_layout(N, Cs, k);
  int* off = reinterpret_cast<int*>(sl_smem + L.off_at);             // [N + 1] in-degree counts, then exclusive offsets
  int* cur = reinterpret_cast<int*>(sl_smem + L.cur_at);             // [N] fill cursors
  unsigned int* smask = reinterpret_cast<unsigned int*>(sl_smem + L.mask_at);  // [N] bit j: slot j of row n is n itself
  unsigned int* rs = reinterpret_cast<unsigned int*>(sl_smem + L.rs_at);       // [N * k] in-edges: item base of the source row | slot << 24
  unsigned short* perm = reinterpret_cast<unsigned short*>(sl_smem + L.perm_at);  // rows ordered by in-degree
  int* wsum = reinterpret_cast<int*>(sl_smem + L.wsum_at);  // [32] warp totals of the scan, then [32] degree-bin cursors
  int* bins = wsum + 32;
  const int tid = threadIdx.x;
  const int ishift = cs_shift - 3;  // (row, 8-channel) items per row = Cs / 8
  const int items = N << ishift;    // <= kSliceThreads by the choice of Cs: one item per thread and unit
  const bool live = tid < items;

  const long long u0 = (long long)blockIdx.x * total_units / gridDim.x;
  const long long u1 = (long long)(blockIdx.x + 1) * total_units / gridDim.x;

  // Every address of an item derives from one 32-bit element offset into grad_out (the host checks that
  // B * N * 2C fits): goff = (b * N + n) * 2C + slice * 2Cs + c8 * 16; argmax byte / grad_x float offset = goff / 2.
  const unsigned c8 = (unsigned)tid & ((1u << ishift) - 1);
  const unsigned rowoff = (unsigned)(tid >> ishift) * 2u * C + c8 * 16u;
  auto unit_base = [&](long long u) {
    const long long b = u / slices;
    const int sl = static_cast<int>(u - b * slices);
    return (unsigned)(b * N) * 2u * C + (unsigned)sl * 2u * Cs;
  };
  // this thread's item of unit u, straight from HBM into registers (consumed two units later)
  auto fetch = [&](long long u, SliceRegs& R) {
    if (u >= u1 || !live) return;
    const unsigned goff = unit_base(u) + rowoff;
    const float4* gp = reinterpret_cast<const float4*>(g + goff);
#pragma unroll
    for (int q = 0; q < 4; ++q) R.g[q] = __ldg(gp + q);
    R.am = __ldg(reinterpret_cast<const uint2*>(argmax + (goff >> 1)));
  };

  // reverse graph of segment b in shared memory (self edges are not listed: their centre and
  // neighbour contributions cancel exactly, which the dense term accounts for through smask)
  auto build_reverse = [&](long long b) {
    const long long ebase = b * (long long)N * k;
    const int E = N * k;
    __syncthreads();  // the previous segment's lists are no longer being read
    for (int i = tid; i <= N; i += kSliceThreads) off[i] = 0;
    for (int i = tid; i < N; i += kSliceThreads) smask[i] = 0u;
    if (tid < 32) bins[tid] = 0;
    __syncthreads();
    for (int e = tid; e < E; e += kSliceThreads) {
      const int m = e / k, j = e - m * k;
      const int t = load_index<I64>(nbr, ebase + e);
      if (t == m) atomicOr(&smask[m], 1u << j);
      else if ((unsigned)t < (unsigned)N) atomicAdd(&off[t], 1);
    }
    __syncthreads();
    // exclusive scan: contiguous span per thread, warp scan of the span sums, then the warp totals
    const int span = (N + kSliceThreads - 1) / kSliceThreads;
    const int lo = min(tid * span, N), hi = min(lo + span, N);
    int local = 0;
    for (int i = lo; i < hi; ++i) local += off[i];
    int incl = local;
    const int lane = tid & 31, warp = tid >> 5;
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
    if (tid == kSliceThreads - 1) off[N] = run;
    __syncthreads();
    for (int e = tid; e < E; e += kSliceThreads) {
      const int m = e / k, j = e - m * k;
      const int t = load_index<I64>(nbr, ebase + e);
      if (t != m && (unsigned)t < (unsigned)N) rs[atomicAdd(&cur[t], 1)] = ((unsigned)m << ishift) | ((unsigned)j << 24);
    }
    __syncthreads();
    // fixed summation order: sort every (short) list; count the rows of every in-degree
    for (int n = tid; n < N; n += kSliceThreads) {
      const int s0 = off[n], s1 = off[n + 1];
      for (int i = s0 + 1; i < s1; ++i) {
        const unsigned v = rs[i];
        int p = i - 1;
        while (p >= s0 && rs[p] > v) { rs[p + 1] = rs[p]; --p; }
        rs[p + 1] = v;
      }
      atomicAdd(&bins[min(s1 - s0, 31)], 1);
    }
    __syncthreads();
    // rows ordered by in-degree (counting sort), so the lanes of a warp walk lists of the same length in
    // pass B; the order inside a bin only decides which thread takes which row, not the arithmetic
    if (tid < 32) {
      const int c = bins[tid];
      int incl2 = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl2, o);
        if (tid >= o) incl2 += v;
      }
      bins[tid] = incl2 - c;
    }
    __syncthreads();
    for (int n = tid; n < N; n += kSliceThreads)
      perm[atomicAdd(&bins[min(off[n + 1] - off[n], 31)], 1)] = (unsigned short)n;
    __syncthreads();
  };

  long long cur_b = -1;
  // one unit: pass A parks this thread's item where the gathers can reach it, pass B gathers for a row picked
  // by in-degree order; `P` is consumed by pass A and refilled with the item of unit u + 2 right after
  auto process = [&](long long u, int buf, SliceRegs& P) {
    const long long b = u / slices;
    if (b != cur_b) { build_reverse(b); cur_b = b; }
    const unsigned ub = unit_base(u);
    float4* g1c = reinterpret_cast<float4*>(sl_smem + buf * L.g1_bytes);
    float4* dd = reinterpret_cast<float4*>(sl_smem + L.dd_at + buf * L.g1_bytes);
    uint2* ams = reinterpret_cast<uint2*>(sl_smem + L.am_at + buf * L.am_bytes);
    if (live) {
      const unsigned sm = smask[tid >> ishift];
      const float4 a = P.g[0], bq = P.g[1], c = P.g[2], e = P.g[3];
      const uint2 am = P.am;
      g1c[2 * tid] = make_float4(a.y, a.w, bq.y, bq.w);
      g1c[2 * tid + 1] = make_float4(c.y, c.w, e.y, e.w);
      ams[tid] = am;
      float4 d0, d1;
      d0.x = ((sm >> (am.x & 0xff)) & 1u) ? a.x : a.x - a.y;
      d0.y = ((sm >> ((am.x >> 8) & 0xff)) & 1u) ? a.z : a.z - a.w;
      d0.z = ((sm >> ((am.x >> 16) & 0xff)) & 1u) ? bq.x : bq.x - bq.y;
      d0.w = ((sm >> (am.x >> 24)) & 1u) ? bq.z : bq.z - bq.w;
      d1.x = ((sm >> (am.y & 0xff)) & 1u) ? c.x : c.x - c.y;
      d1.y = ((sm >> ((am.y >> 8) & 0xff)) & 1u) ? c.z : c.z - c.w;
      d1.z = ((sm >> ((am.y >> 16) & 0xff)) & 1u) ? e.x : e.x - e.y;
      d1.w = ((sm >> (am.y >> 24)) & 1u) ? e.z : e.z - e.w;
      dd[2 * tid] = d0;
      dd[2 * tid + 1] = d1;
    }
    fetch(u + 2, P);  // in flight across the barrier, pass B and the whole next unit
    __syncthreads();
    if (live) {
      const int n = perm[tid >> ishift];
      const unsigned it = ((unsigned)n << ishift) + c8;
      float4 r0 = dd[2 * it], r1 = dd[2 * it + 1];
      const int e1 = off[n + 1];
#pragma unroll 1
      for (int e = off[n]; e < e1; ++e) {
        const unsigned ent = rs[e];
        const unsigned j = ent >> 24;
        const unsigned mi = (ent & 0xffffffu) + c8;  // item index of (source row, this 8-channel pack)
        const uint2 amm = ams[mi];
        const float4 m0 = g1c[2 * mi], m1 = g1c[2 * mi + 1];
        r0.x += ((amm.x & 0xff) == j) ? m0.x : 0.f;
        r0.y += (((amm.x >> 8) & 0xff) == j) ? m0.y : 0.f;
        r0.z += (((amm.x >> 16) & 0xff) == j) ? m0.z : 0.f;
        r0.w += ((amm.x >> 24) == j) ? m0.w : 0.f;
        r1.x += ((amm.y & 0xff) == j) ? m1.x : 0.f;
        r1.y += (((amm.y >> 8) & 0xff) == j) ? m1.y : 0.f;
        r1.z += (((amm.y >> 16) & 0xff) == j) ? m1.z : 0.f;
        r1.w += ((amm.y >> 24) == j) ? m1.w : 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(grad_x + ((ub + (unsigned)n * 2u * C + c8 * 16u) >> 1));
      dst[0] = r0;
      dst[1] = r1;
    }
  };

  SliceRegs P0, P1;
  fetch(u0, P0);
  fetch(u0 + 1, P1);
  for (long long u = u0; u < u1; u += 2) {
    process(u, 0, P0);
    if (u + 1 < u1) process(u + 1, 1, P1);
  }
}

}  // namespace

// Picks the slice width and launches; *launched = false (and GRAFP_OK) when the shape is outside the envelope.
template <bool I64>
int launch_mr_bwd_slice(const float* g, const uint8_t* argmax, const void* nbr, float* grad_x, int B, int N, int C, int k,
                        cudaStream_t s, bool* launched) {
  *launched = false;
  if (k < 1 || k > 32 || N < 1 || N > 65535 || C % 8 != 0 || (long long)B * N * 2 * C >= (1LL << 31)) return GRAFP_OK;
  if (!aligned16(g) || !aligned16(grad_x)) return GRAFP_OK;
  // widest power-of-two slice whose grad_out columns fit 64 KB per buffer; at least 8 channels (64-byte row pieces)
  int cs_shift = 3;
  while ((2 << cs_shift) <= 128 && C % (2 << cs_shift) == 0 && (size_t)N * (2 << cs_shift) <= 8 * kSliceThreads) ++cs_shift;
  const int Cs = 1 << cs_shift;
  if (C % Cs != 0 || (size_t)N * Cs > 8 * (size_t)kSliceThreads) return GRAFP_OK;  // one (row, 8-channel) item per thread
  if ((reinterpret_cast<uintptr_t>(argmax) & 7u) != 0) return GRAFP_OK;  // argmax is read as 8-byte words
  const SliceLayout L = slice_layout(N, Cs, k);
  if (L.total > kSliceSmemLimit) return GRAFP_OK;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mr_bwd_slice_kernel<I64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSliceSmemLimit);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(mr_bwd_slice): %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  const int slices = C / Cs;
  const long long total = (long long)B * slices;
  int ctas = num_sms();
  if (ctas > total) ctas = (int)total;
  mr_bwd_slice_kernel<I64><<<ctas, kSliceThreads, L.total, s>>>(g, argmax, nbr, grad_x, N, C, k, cs_shift, slices, total);
  *launched = true;
  return check_launch("mr_aggregate_bwd_slice");
}

template int launch_mr_bwd_slice<true>(const float*, const uint8_t*, const void*, float*, int, int, int, int, cudaStream_t, bool*);
template int launch_mr_bwd_slice<false>(const float*, const uint8_t*, const void*, float*, int, int, int, int, cudaStream_t, bool*);

}  // namespace grafp
