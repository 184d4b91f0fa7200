// Shared device/host helpers for libgrafp_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <type_traits>

#include "../../include/grafp_b200.h"

namespace grafp {

// ---- host-side error plumbing (thread local; no exceptions cross the ABI) ----
void set_error(const char* fmt, ...);
void clear_error();
int check_launch(const char* what);  // cudaGetLastError -> return code + message
int num_sms();

#define GRAFP_REQUIRE(cond, code, ...)  \
  do {                                  \
    if (!(cond)) {                      \
      ::grafp::set_error(__VA_ARGS__);  \
      return (code);                    \
    }                                   \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- run-time options (grafp_set_option / grafp_get_option; defaults from GRAFP_<NAME> read once at load) ----
enum Option {
  OPT_MR_FWD_FORM = 0,   // K2: 0 generic, 1 register-prefetch, 2 cp.async-pipelined persistent kernel (default)
  OPT_MR_BWD_FORM,       // K3: 0 dense + scatter pair, 1 cluster kernel with a register stash, 2 cluster kernel with bulk
                         //     staging (default), 3 deterministic gather over the reverse graph (needs the workspace)
  OPT_KNN_EPILOGUE,      // K1 selection: 0 auto, 1 vote-gated scan, 2 candidate queues, 3 group maxima (K <= 4)
  OPT_EDGE_BWD_ROW,      // EdgeConv backward: 1 one-pass row form (default), 0 dense + scatter pair
  OPT_GATHER_ROW,        // plain gather: 1 row form (default)
  OPT_EDGE_ROW,          // EdgeConv features: 1 row form (default)
  OPT_MAXK_ROW,          // max over k: 1 row form (default)
  OPT_BN_REVERSE,        // K5: 1 = first pass back to front, second pass front to back (default); 0 = the other way round
  OPT_BN_PERSISTENT,     // K5: 1 = both passes in one cooperative launch (default), 0 = two launches
  OPT_BN_L2_KEEP_MB,     // K5: megabytes of pass-1 input kept in L2 ("evict last") for pass 2 (default 80); 0 = no cache hints
  OPT_CHECK_INDEX,       // 1 = validate neighbour / centre ids against [0, M) before the aggregation kernels run
  OPT_CONV_GEMM,         // 1 = dense 1x1 convolutions in front of a train-mode BatchNorm run as the tcgen05 GEMM with the
                         //     statistics in its epilogue (host-side switch, read by ops.py); 0 = cuDNN + full BatchNorm
  OPT_COUNT
};
int option(Option o);

// One-time per-DEVICE kernel configuration (cudaFuncSetAttribute is per device, not per process).
int current_device();
struct DeviceOnce {
  bool done[64] = {};
  bool pending() const { const int d = current_device(); return d < 0 || d >= 64 || !done[d]; }
  void mark() { const int d = current_device(); if (d >= 0 && d < 64) done[d] = true; }
};

// ---- element access: VEC consecutive channels as fp32 registers ----
template <typename T, int VEC>
struct Pack;

template <>
struct Pack<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ void store_streaming(float* p, const float (&v)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
  static __device__ __forceinline__ void red_add(float* p, const float (&v)[4]) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3])
                 : "memory");
  }
};

template <>
struct Pack<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
  static __device__ __forceinline__ void red_add(float* p, const float (&v)[1]) { atomicAdd(p, v[0]); }
};

template <>
struct Pack<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
  static __device__ __forceinline__ void store_streaming(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    __stcs(reinterpret_cast<uint2*>(p), t);
  }
  static __device__ __forceinline__ void red_add(__nv_bfloat16* p, const float (&v)[4]) {
    atomicAdd(reinterpret_cast<__nv_bfloat162*>(p), __floats2bfloat162_rn(v[0], v[1]));
    atomicAdd(reinterpret_cast<__nv_bfloat162*>(p) + 1, __floats2bfloat162_rn(v[2], v[3]));
  }
};

template <>
struct Pack<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[1]) { v[0] = __bfloat162float(*p); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[1]) { *p = __float2bfloat16_rn(v[0]); }
  static __device__ __forceinline__ void red_add(__nv_bfloat16* p, const float (&v)[1]) {
    atomicAdd(p, __float2bfloat16_rn(v[0]));
  }
};

// 8 consecutive elements (32 bytes of fp32 -> one 256-bit LDG/STG on sm_100; 16 bytes of bf16)
template <typename T>
struct Pack8;

template <>
struct Pack8<float> {
  static constexpr int kAlign = 32;
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
  }
};

template <>
struct Pack8<__nv_bfloat16> {
  static constexpr int kAlign = 16;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

template <bool I64>
__device__ __forceinline__ int load_index(const void* idx, long long pos) {
  if constexpr (I64) {
    return static_cast<int>(__ldg(reinterpret_cast<const long long*>(idx) + pos));
  } else {
    return __ldg(reinterpret_cast<const int*>(idx) + pos);
  }
}

// Launch geometry for grid-stride element-wise kernels: a whole number of waves of
// `ctas_per_sm` resident CTAs on every SM.
inline int grid_for(long long items, int threads, int ctas_per_sm) {
  long long need = (items + threads - 1) / threads;
  long long cap = static_cast<long long>(num_sms()) * ctas_per_sm;
  if (need < 1) need = 1;
  return static_cast<int>(need < cap ? need : cap);
}

// ---- launchers implemented in aggregate.cu ----
template <typename T>
int launch_mr_aggregate_fwd(const void* x, const void* y, const void* nbr, const void* ctr, int idx_is_i64, void* out,
                            uint8_t* argmax, int B, int N, int M, int C, int k, cudaStream_t s);
template <typename T>
int launch_mr_aggregate_bwd(const void* g, const uint8_t* argmax, const void* nbr, const void* ctr, int idx_is_i64,
                            void* grad_x, void* grad_y, int B, int N, int M, int C, int k, void* workspace,
                            size_t workspace_bytes, cudaStream_t s);
size_t mr_bwd_workspace_bytes(int B, int N, int k);
template <typename T>
int launch_gather_fwd(const void* src, const void* idx, int idx_is_i64, void* out, int B, int N, int M, int C, int k,
                      cudaStream_t s);
template <typename T>
int launch_gather_bwd(const void* g, const void* idx, int idx_is_i64, void* grad_src, int B, int N, int M, int C, int k,
                      cudaStream_t s);
template <typename T>
int launch_neighbor_sum_fwd(const void* src, const void* idx, int idx_is_i64, void* out, int B, int N, int M, int C, int k,
                            cudaStream_t s);
template <typename T>
int launch_neighbor_sum_bwd(const void* g, const void* idx, int idx_is_i64, void* grad_src, int B, int N, int M, int C,
                            int k, cudaStream_t s);
template <typename T>
int launch_edge_gather_fwd(const void* x, const void* y, const void* nbr, const void* ctr, int idx_is_i64, void* out,
                           int B, int N, int M, int C, int k, cudaStream_t s);
template <typename T>
int launch_edge_gather_bwd(const void* g, const void* nbr, const void* ctr, int idx_is_i64, void* grad_x, void* grad_y,
                           int B, int N, int M, int C, int k, cudaStream_t s);
template <typename T>
int launch_max_over_k_fwd(const void* h, void* out, uint8_t* argmax, int B, int N, int C, int k, cudaStream_t s);
template <typename T>
int launch_max_over_k_bwd(const void* g, const uint8_t* argmax, void* grad_h, int B, int N, int C, int k,
                          cudaStream_t s);

int launch_check_index(const void* idx, int idx_is_i64, long long count, int limit, int* bad_count, cudaStream_t s);

// ---- bn_fused.cu ----
size_t bn_workspace_bytes(int C);
bool bn_supported(long long R, int C, int dtype);
int launch_bn_train_fwd(const void* x, const void* res, const float* weight, const float* bias, float* running_mean,
                        float* running_var, const float* conv_bias, long long* num_batches_tracked, void* out, float* save_mean,
                        float* save_invstd, long long R, int C, float eps, float momentum, int relu, int dtype, void* workspace,
                        bool from_moments, cudaStream_t s);
double* bn_workspace_sums(void* workspace);  // where the per-channel sums live inside a BatchNorm workspace
// conv_gemm.cu: Y = X W^T on tcgen05 with sum y / sum y^2 per output channel (2 * Cout doubles, zeroed by the call)
bool conv1x1_stats_supported(long long R, int Cin, int Cout, int groups, int dtype);
int launch_conv1x1_stats(const void* x, const void* w, void* y, double* sums, long long R, int Cin, int Cout, int groups,
                         int dtype, cudaStream_t s);
int launch_bn_train_bwd(const void* dy, const void* x, const float* weight, const float* bias, const float* save_mean,
                        const float* save_invstd, void* dx, float* dweight, float* dbias, float* dx_colsum, long long R, int C,
                        int relu, int dtype, void* workspace, cudaStream_t s);

// ---- downsample.cu ----
bool downsample_taps_supported(int N, int C, int dtype);
int launch_downsample_taps_fwd(const void* x, void* taps, int B, int N, int C, int dtype, cudaStream_t s);
int launch_downsample_taps_bwd(const void* dtaps, void* dx, int B, int N, int C, int dtype, cudaStream_t s);

// ---- peak_extract.cu ----
bool peak_extract_supported(int H, int W, int F, int kh, int kw, int sh);
size_t peak_extract_workspace_bytes(int B, int kh, int kw);
int launch_peak_extract_fwd(const float* spec, const float* weight, const float* bias, float* out, int B, int H, int W, int kh,
                            int kw, int sh, cudaStream_t s);
int launch_peak_extract_bwd(const float* spec, const float* out, const float* grad_out, float* partial, float* dweight,
                            float* dbias, int B, int H, int W, int kh, int kw, int sh, cudaStream_t s);

// ---- ntxent.cu ----
bool ntxent_supported(int n2, int d);
int launch_ntxent_fwd(const float* z, float* lse, float* row_loss, float* loss, int n2, int d, float inv_tau, int row_lo,
                      int row_hi, cudaStream_t s);
int launch_ntxent_bwd(const float* z, const float* lse, const float* grad_loss, float* dz, int n2, int d, float inv_tau,
                      int row_lo, int row_hi, float grad_scale, cudaStream_t s);

}  // namespace grafp
