// Request-state kernels of the denoising step whose arguments arrive from the HOST by value:
// latent gather (+ CFG duplication + Euler input scaling), fused CFG combine + scheduler update,
// scalar-table upload and a batched row gather. The per-request descriptors (device pointers,
// element counts, sigma / sigma_next) travel inside the kernel parameter block, so a step issues
// no host->device copy and never synchronises the stream: the reference builds torch.tensor(...)
// on the host and `.to(device)`s it 2-3 times per resolution and step
// (scheduling_euler_discrete.py:176,213,254; scheduling_flow_match_euler_discrete.py:183-189),
// each of which drains the stream when the source is pageable memory.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/sduss_b200.h"
#include "host_util.h"

namespace b200 {

constexpr int STEP_MAX_REQ = 32;  // requests per launch (the launchers split longer lists)

struct LatentRefDev {
  const void* src;
  void* dst;
  long long elems, off_a, off_b;
  float s, sn;
  int pad;
};
struct StepBatch {
  LatentRefDev r[STEP_MAX_REQ];
};

// ---- rounding to the arithmetic type of a torch tensor op chain
template <int DT> __device__ __forceinline__ float rnd(float v);
template <> __device__ __forceinline__ float rnd<B200_DT_BF16>(float v) {
  return __bfloat162float(__float2bfloat16(v));
}
template <> __device__ __forceinline__ float rnd<B200_DT_F16>(float v) {
  return __half2float(__float2half_rn(v));
}
template <> __device__ __forceinline__ float rnd<B200_DT_F32>(float v) { return v; }

template <int DT> __device__ __forceinline__ float ld(const void* p, long long i);
template <> __device__ __forceinline__ float ld<B200_DT_BF16>(const void* p, long long i) {
  return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
}
template <> __device__ __forceinline__ float ld<B200_DT_F16>(const void* p, long long i) {
  return __half2float(static_cast<const __half*>(p)[i]);
}
template <> __device__ __forceinline__ float ld<B200_DT_F32>(const void* p, long long i) {
  return static_cast<const float*>(p)[i];
}
template <int DT> __device__ __forceinline__ void st(void* p, long long i, float v);
template <> __device__ __forceinline__ void st<B200_DT_BF16>(void* p, long long i, float v) {
  static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
}
template <> __device__ __forceinline__ void st<B200_DT_F16>(void* p, long long i, float v) {
  static_cast<__half*>(p)[i] = __float2half_rn(v);
}
template <> __device__ __forceinline__ void st<B200_DT_F32>(void* p, long long i, float v) {
  static_cast<float*>(p)[i] = v;
}

// ------------------------------------------------------------------ latent gather
// staging[off_a + i] (and staging[off_b + i] when off_b >= 0: the CFG duplicate) =
//   bf16( scale ? x_i / sqrt(sigma^2 + 1) : x_i )
// The scaling is EulerDiscreteScheduler.batch_scale_model_input
// (scheduling_euler_discrete.py:161-184): sigma is built in the sample dtype and every tensor op
// (**2, +1, **0.5, /) rounds to that dtype.
template <int DT>
__global__ void gather_latents_kernel(const __grid_constant__ StepBatch b, int scale,
                                      __nv_bfloat16* __restrict__ staging) {
  const LatentRefDev& r = b.r[blockIdx.y];
  float den = 1.f;
  if (scale) {
    const float s = rnd<DT>(r.s);
    const float s2 = rnd<DT>(s * s);
    const float s2p = rnd<DT>(s2 + 1.f);
    den = rnd<DT>(sqrtf(s2p));
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < r.elems;
       i += (long long)gridDim.x * blockDim.x) {
    float v = ld<DT>(r.src, i);
    if (scale) v = rnd<DT>(__fdiv_rn(v, den));
    const __nv_bfloat16 o = __float2bfloat16(v);
    staging[r.off_a + i] = o;
    if (r.off_b >= 0) staging[r.off_b + i] = o;
  }
}

// ------------------------------------------------------------------ CFG + scheduler update
// Per request r: eps = u + g (c - u) when cfg (pipeline_stable_diffusion_3_esymred.py:326-329,
// pipeline_stable_diffusion_xl_esymred.py:382-385; tensor ops in the model-output dtype ET, each
// rounding to ET), then in fp32 on the fp32-upcast sample:
//   mode 0 (flow match, scheduling_flow_match_euler_discrete.py:159-203): x' = x + (s' - s) eps
//   mode 1 (Euler epsilon,  scheduling_euler_discrete.py:187-274): x0 = x - s eps; d = (x - x0)/s;
//          x' = x + d (s' - s)
//   mode 2 (Euler v_prediction): x0 = eps * (-s / sqrt(s^2+1)) + x / (s^2+1); rest as mode 1
// Explicit round-to-nearest mul/add (no FMA contraction): the reference's op order. x' is stored
// in the latent's own dtype XT. (The reference stores it in the model-output dtype; identical when
// XT == ET, and an fp32 / fp16 latent is not degraded to bf16 between steps.)
template <int XT, int ET>
__global__ void cfg_step_kernel(const __grid_constant__ StepBatch b, const void* __restrict__ eps,
                                float guidance, int cfg, int mode) {
  const LatentRefDev& r = b.r[blockIdx.y];
  const float s = r.s, sn = r.sn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < r.elems;
       i += (long long)gridDim.x * blockDim.x) {
    float e;
    if (cfg) {
      const float u = ld<ET>(eps, r.off_a + i);
      const float c = ld<ET>(eps, r.off_b + i);
      const float diff = rnd<ET>(__fsub_rn(c, u));
      const float sc = rnd<ET>(__fmul_rn(guidance, diff));
      e = rnd<ET>(__fadd_rn(u, sc));
    } else {
      e = ld<ET>(eps, r.off_b + i);
    }
    const float x = ld<XT>(r.src, i);
    float xn;
    if (mode == 0) {
      xn = __fadd_rn(x, __fmul_rn(__fsub_rn(sn, s), e));
    } else {
      float x0;
      if (mode == 1) {
        x0 = __fsub_rn(x, __fmul_rn(s, e));
      } else {
        const float s2p = __fadd_rn(__fmul_rn(s, s), 1.f);
        x0 = __fadd_rn(__fmul_rn(e, __fdiv_rn(-s, __fsqrt_rn(s2p))), __fdiv_rn(x, s2p));
      }
      const float d = __fdiv_rn(__fsub_rn(x, x0), s);
      xn = __fadd_rn(x, __fmul_rn(d, __fsub_rn(sn, s)));
    }
    st<XT>(r.dst, i, xn);
  }
}

// ------------------------------------------------------------------ small uploads
constexpr int F32_MAX = 256;
struct F32Batch {
  float v[F32_MAX];
};
__global__ void write_f32_kernel(const __grid_constant__ F32Batch b, float* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = b.v[i];
}

constexpr int ROWS_MAX = 64;
struct RowBatch {
  const void* src[ROWS_MAX];
};
// dst + i * stride <- src[i][0:bytes): 16-byte vectors when everything is 16-byte aligned
template <typename V>
__global__ void gather_rows_kernel(const __grid_constant__ RowBatch b, char* dst,
                                   long long stride, long long n_vec) {
  const V* s = static_cast<const V*>(b.src[blockIdx.y]);
  V* d = reinterpret_cast<V*>(dst + blockIdx.y * stride);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_vec;
       i += (long long)gridDim.x * blockDim.x)
    d[i] = s[i];
}

template <typename F>
int for_batches(const B200LatentRef* reqs, int n, F&& launch) {
  for (int base = 0; base < n; base += STEP_MAX_REQ) {
    const int m = n - base < STEP_MAX_REQ ? n - base : STEP_MAX_REQ;
    StepBatch b;
    long long max_elems = 0;
    for (int i = 0; i < m; ++i) {
      const B200LatentRef& q = reqs[base + i];
      if (!q.src || q.elems <= 0) return B200_ERR_INVALID;
      b.r[i] = LatentRefDev{q.src, q.dst, q.elems, q.off_a, q.off_b, q.sigma, q.sigma_next, 0};
      if (q.elems > max_elems) max_elems = q.elems;
    }
    unsigned gx = unsigned((max_elems + 1023) / 1024);
    if (gx > 256) gx = 256;
    const int rc = launch(b, dim3(gx, m));
    if (rc != B200_OK) return rc;
  }
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_gather_latents(const B200LatentRef* reqs_host, int n_requests, int dtype,
                                   int scale_input, void* staging, void* stream) {
  if (!reqs_host || n_requests <= 0 || !staging || dtype < 0 || dtype > 2) return B200_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(staging);
  for (int i = 0; i < n_requests; ++i)
    if (reqs_host[i].off_a < 0) return B200_ERR_INVALID;
  return for_batches(reqs_host, n_requests, [&](const StepBatch& b, dim3 grid) {
    switch (dtype) {
      case B200_DT_BF16: gather_latents_kernel<B200_DT_BF16><<<grid, 256, 0, st>>>(b, scale_input, out); break;
      case B200_DT_F16: gather_latents_kernel<B200_DT_F16><<<grid, 256, 0, st>>>(b, scale_input, out); break;
      default: gather_latents_kernel<B200_DT_F32><<<grid, 256, 0, st>>>(b, scale_input, out); break;
    }
    return launch_status();
  });
}

extern "C" int b200_cfg_scheduler_step(const void* eps, int eps_dtype,
                                       const B200LatentRef* reqs_host, int n_requests,
                                       int latent_dtype, float guidance, int cfg, int mode,
                                       void* stream) {
  if (!eps || !reqs_host || n_requests <= 0 || mode < 0 || mode > 2 || latent_dtype < 0 ||
      latent_dtype > 2 || (eps_dtype != B200_DT_BF16 && eps_dtype != B200_DT_F32))
    return B200_ERR_INVALID;
  for (int i = 0; i < n_requests; ++i)
    if (!reqs_host[i].dst || reqs_host[i].off_b < 0 || (cfg && reqs_host[i].off_a < 0))
      return B200_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return for_batches(reqs_host, n_requests, [&](const StepBatch& b, dim3 grid) {
#define B200_STEP_CASE(XT, ET)                                                      \
  if (latent_dtype == XT && eps_dtype == ET) {                                      \
    cfg_step_kernel<XT, ET><<<grid, 256, 0, st>>>(b, eps, guidance, cfg, mode);     \
    return launch_status();                                                         \
  }
    B200_STEP_CASE(B200_DT_BF16, B200_DT_BF16)
    B200_STEP_CASE(B200_DT_F16, B200_DT_BF16)
    B200_STEP_CASE(B200_DT_F32, B200_DT_BF16)
    B200_STEP_CASE(B200_DT_BF16, B200_DT_F32)
    B200_STEP_CASE(B200_DT_F16, B200_DT_F32)
    B200_STEP_CASE(B200_DT_F32, B200_DT_F32)
#undef B200_STEP_CASE
    return int(B200_ERR_UNSUPPORTED);
  });
}

extern "C" int b200_write_f32(void* dst, const float* vals_host, int n, void* stream) {
  if (!dst || !vals_host || n <= 0) return B200_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* out = static_cast<float*>(dst);
  for (int base = 0; base < n; base += F32_MAX) {
    const int m = n - base < F32_MAX ? n - base : F32_MAX;
    F32Batch b;
    for (int i = 0; i < m; ++i) b.v[i] = vals_host[base + i];
    write_f32_kernel<<<(m + 255) / 256, 256, 0, st>>>(b, out + base, m);
    const int rc = launch_status();
    if (rc != B200_OK) return rc;
  }
  return B200_OK;
}

extern "C" int b200_gather_rows(void* dst, long long dst_stride_bytes,
                                const void* const* src_host, int n, long long bytes_each,
                                void* stream) {
  if (!dst || !src_host || n <= 0 || bytes_each <= 0 || (bytes_each & 3) || (dst_stride_bytes & 3))
    return B200_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* out = static_cast<char*>(dst);
  for (int base = 0; base < n; base += ROWS_MAX) {
    const int m = n - base < ROWS_MAX ? n - base : ROWS_MAX;
    RowBatch b;
    bool v16 = (bytes_each % 16 == 0) && (dst_stride_bytes % 16 == 0) &&
               (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    for (int i = 0; i < m; ++i) {
      if (!src_host[base + i]) return B200_ERR_INVALID;
      b.src[i] = src_host[base + i];
      v16 = v16 && (reinterpret_cast<uintptr_t>(b.src[i]) % 16 == 0);
    }
    const long long n_vec = bytes_each / (v16 ? 16 : 4);
    long long gx = (n_vec + 255) / 256;
    if (gx > 1184) gx = 1184;  // 8 waves of CTAs over 148 SMs; grid-stride beyond
    char* d = out + base * dst_stride_bytes;
    if (v16)
      gather_rows_kernel<uint4><<<dim3(unsigned(gx), m), 256, 0, st>>>(b, d, dst_stride_bytes, n_vec);
    else
      gather_rows_kernel<uint32_t><<<dim3(unsigned(gx), m), 256, 0, st>>>(b, d, dst_stride_bytes, n_vec);
    const int rc = launch_status();
    if (rc != B200_OK) return rc;
  }
  return B200_OK;
}
