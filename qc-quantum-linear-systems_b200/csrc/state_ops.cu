// Error reporting, device info, state initialisation, host<->device access and reductions.
#include <stdarg.h>
#include <string.h>

#include "qls_common.cuh"

static thread_local char g_err[512] = "";

void qls_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int qls_abi_version(void) { return QLS_ABI_VERSION; }
extern "C" const char* qls_last_error(void) { return g_err; }

extern "C" int qls_device_info(int device, int* sm_count, uint64_t* smem_optin, uint64_t* mem_total, uint64_t* mem_free) {
  cudaDeviceProp prop;
  QLS_CUDA(cudaGetDeviceProperties(&prop, device));
  QLS_CUDA(cudaSetDevice(device));
  size_t fr = 0, tot = 0;
  QLS_CUDA(cudaMemGetInfo(&fr, &tot));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (smem_optin) *smem_optin = prop.sharedMemPerBlockOptin;
  if (mem_total) *mem_total = tot;
  if (mem_free) *mem_free = fr;
  return QLS_OK;
}

// ---- init / scale -----------------------------------------------------------------------------------

__global__ void qls_fill_zero_kernel(uint4* __restrict__ p, uint64_t n16) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = z;
}

template <typename V>
__global__ void qls_set_one_kernel(V* p, uint64_t idx) {
  V v;
  v.x = 1;
  v.y = 0;
  p[idx] = v;
}

template <typename V>
__global__ void qls_scale_kernel(V* __restrict__ p, uint64_t n, V f) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = cmul(p[i], f);
}

static inline unsigned grid_for(uint64_t n, int threads, int per_sm_blocks = 8) {
  uint64_t want = (n + threads - 1) / threads;
  uint64_t cap = 148ull * per_sm_blocks;
  return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

extern "C" int qls_sv_init_basis(void* state, int n_local, int dtype, uint64_t basis_index, void* stream) {
  QLS_ARG(state && n_local >= 0 && n_local < 40, "bad state / n_local");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t n = 1ull << n_local;
  const uint64_t bytes = n * (dtype == QLS_C128 ? 16 : 8);
  if (bytes >= 16) {
    const uint64_t n16 = bytes / 16;
    qls_fill_zero_kernel<<<grid_for(n16, 256), 256, 0, s>>>(reinterpret_cast<uint4*>(state), n16);
  } else {
    QLS_CUDA(cudaMemsetAsync(state, 0, bytes, s));
  }
  if (basis_index != UINT64_MAX) {
    QLS_ARG(basis_index < n, "basis index outside the state");
    if (dtype == QLS_C128) qls_set_one_kernel<<<1, 1, 0, s>>>(reinterpret_cast<double2*>(state), basis_index);
    else qls_set_one_kernel<<<1, 1, 0, s>>>(reinterpret_cast<float2*>(state), basis_index);
  }
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}

extern "C" int qls_sv_scale(void* state, int n_local, int dtype, double re, double im, void* stream) {
  QLS_ARG(state, "null state");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t n = 1ull << n_local;
  if (dtype == QLS_C128)
    qls_scale_kernel<<<grid_for(n, 256), 256, 0, s>>>(reinterpret_cast<double2*>(state), n, make_double2(re, im));
  else
    qls_scale_kernel<<<grid_for(n, 256), 256, 0, s>>>(reinterpret_cast<float2*>(state), n,
                                                      make_float2((float)re, (float)im));
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}

extern "C" int qls_sv_write(void* state, int dtype, uint64_t offset, const void* amps_host, uint64_t count, void* stream) {
  QLS_ARG(state && (amps_host || count == 0), "null pointer");
  const size_t esz = dtype == QLS_C128 ? 16 : 8;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  QLS_CUDA(cudaMemcpyAsync(static_cast<char*>(state) + offset * esz, amps_host, count * esz, cudaMemcpyHostToDevice, s));
  QLS_CUDA(cudaStreamSynchronize(s));
  return QLS_OK;
}

extern "C" int qls_sv_read(const void* state, int dtype, uint64_t offset, void* amps_host, uint64_t count, void* stream) {
  QLS_ARG(state && (amps_host || count == 0), "null pointer");
  const size_t esz = dtype == QLS_C128 ? 16 : 8;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  QLS_CUDA(cudaMemcpyAsync(amps_host, static_cast<const char*>(state) + offset * esz, count * esz, cudaMemcpyDeviceToHost, s));
  QLS_CUDA(cudaStreamSynchronize(s));
  return QLS_OK;
}

// ---- reductions -----------------------------------------------------------------------------------------
// Grid-stride accumulation in double, warp-shuffle tree, one partial per block, then a single block
// folds the partials in a fixed order (deterministic: no floating-point atomics).

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int THREADS>
__device__ __forceinline__ void block_store_partial(double a, double b, double* partials) {
  __shared__ double sa[THREADS / 32], sb[THREADS / 32];
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    sa[w] = a;
    sb[w] = b;
  }
  __syncthreads();
  if (w == 0) {
    a = lane < THREADS / 32 ? sa[lane] : 0.0;
    b = lane < THREADS / 32 ? sb[lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      partials[2 * blockIdx.x] = a;
      partials[2 * blockIdx.x + 1] = b;
    }
  }
}

template <typename V, int THREADS>
__global__ void __launch_bounds__(THREADS)
    qls_reduce_kernel(const V* __restrict__ psi, const V* __restrict__ other, int kind, uint64_t a, uint64_t b,
                      uint64_t begin, uint64_t count, uint64_t index_offset, double* __restrict__ partials) {
  double s0 = 0.0, s1 = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * THREADS;
  for (uint64_t j = (uint64_t)blockIdx.x * THREADS + threadIdx.x; j < count; j += stride) {
    const uint64_t i = begin + j;
    const V v = psi[i];
    const double p = (double)v.x * v.x + (double)v.y * v.y;
    if (kind == QLS_RED_NORM_RANGE) {
      s0 += p;
    } else if (kind == QLS_RED_PROB_MASK) {
      if (((index_offset | i) & a) == b) s0 += p;
    } else if (kind == QLS_RED_EXPECT_Z) {
      s0 += (((index_offset | i) >> a) & 1ull) ? -p : p;
    } else {  // <other|psi> = sum conj(other_i) psi_i
      const V o = other[i];
      s0 += (double)o.x * v.x + (double)o.y * v.y;
      s1 += (double)o.x * v.y - (double)o.y * v.x;
    }
  }
  block_store_partial<THREADS>(s0, s1, partials);
}

__global__ void qls_reduce_final_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) {
    a += partials[2 * i];
    b += partials[2 * i + 1];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (threadIdx.x == 0) {
    out[0] = a;
    out[1] = b;
  }
}

extern "C" int qls_reduce(const void* state, int n_local, int dtype, int kind, uint64_t a, uint64_t b,
                          uint64_t index_offset, const void* other, void* workspace, double* out_host, void* stream) {
  QLS_ARG(state && workspace && out_host, "null pointer");
  QLS_ARG(kind >= QLS_RED_NORM_RANGE && kind <= QLS_RED_INNER, "bad reduction kind %d", kind);
  QLS_ARG(kind != QLS_RED_INNER || other, "inner product needs `other`");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t n = 1ull << n_local;
  uint64_t begin = 0, count = n;
  if (kind == QLS_RED_NORM_RANGE) {
    QLS_ARG(a <= n && b <= n - a, "range outside the state");
    begin = a;
    count = b;
  }
  constexpr int THREADS = 256;
  const int max_blocks = (int)((QLS_REDUCE_WORKSPACE_BYTES - 16) / 16);
  uint64_t want = (count + THREADS * 4 - 1) / (THREADS * 4);
  int blocks = (int)(want < 1 ? 1 : (want > 148ull * 8 ? 148 * 8 : want));
  if (blocks > max_blocks) blocks = max_blocks;
  double* partials = reinterpret_cast<double*>(workspace) + 2;
  double* result = reinterpret_cast<double*>(workspace);
  if (dtype == QLS_C128)
    qls_reduce_kernel<double2, THREADS><<<blocks, THREADS, 0, s>>>(
        reinterpret_cast<const double2*>(state), reinterpret_cast<const double2*>(other), kind, a, b, begin, count,
        index_offset, partials);
  else
    qls_reduce_kernel<float2, THREADS><<<blocks, THREADS, 0, s>>>(
        reinterpret_cast<const float2*>(state), reinterpret_cast<const float2*>(other), kind, a, b, begin, count,
        index_offset, partials);
  qls_reduce_final_kernel<<<1, 32, 0, s>>>(partials, blocks, result);
  QLS_CUDA(cudaGetLastError());
  QLS_CUDA(cudaMemcpyAsync(out_host, result, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  QLS_CUDA(cudaStreamSynchronize(s));
  return QLS_OK;
}
