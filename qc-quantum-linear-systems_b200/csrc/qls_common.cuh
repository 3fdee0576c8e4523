// Shared device/host helpers for libqls_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "qls_abi.h"

static_assert(sizeof(QlsOp) == 128, "QlsOp must be 128 bytes (mirrored in _abi.py)");
static_assert(sizeof(QlsPass) == 128, "QlsPass must be 128 bytes (mirrored in _abi.py)");
static_assert(sizeof(QlsRsOp) == 192, "QlsRsOp must be 192 bytes (mirrored in _abi.py)");

void qls_set_error(const char* fmt, ...);

#define QLS_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      qls_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));     \
      return QLS_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define QLS_ARG(cond, ...)            \
  do {                                \
    if (!(cond)) {                    \
      qls_set_error(__VA_ARGS__);     \
      return QLS_ERR_ARG;             \
    }                                 \
  } while (0)

struct QlsProgram {
  int device;
  int dtype;
  uint32_t n_passes, n_ops;
  QlsPass* passes_host;  // host copy (launch geometry is computed on the host)
  QlsOp* ops_dev;
  QlsOp* ops_host;  // host copy of the sweep-form ops (the batched estimator splits a pass at its MATVEC ops)
  QlsRsOp* rs_ops_dev;
  uint32_t n_rs_ops;
  unsigned char* pool_dev;
  uint64_t pool_bytes;
  // register-stage form: per-pass constant-memory images (header + micro-ops + matrices), device resident
  unsigned char* rs_img_dev;
  uint64_t* rs_img_off;    // [n_passes] byte offset into rs_img_dev
  uint32_t* rs_img_bytes;  // [n_passes] 0: the pass has no register-stage form
  uint8_t* rs_img_kr;      // [n_passes] register qubits per thread of the pass's stages (4 or 5)
  // specialised (NVRTC-compiled) kernels, one per pass that has one; nullptr: none attached (jit.cu)
  struct QlsJitKernel** jit;
  // launch-wide state array of the batched estimator's GEMM form (batch_sv.cu): allocated on first use, grown on demand,
  // freed with the program (a stream-ordered allocation per call cost up to 4 ms per launch when the pool had been trimmed)
  void* batch_scratch;
  uint64_t batch_scratch_bytes;
};

// register-stage fast path (rs_pass.cu)
int qls_rs_build_images(QlsProgram* prog, const QlsRsOp* rs_ops_host, const unsigned char* pool_host);
bool qls_rs_eligible(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, int n_local);
int qls_launch_rs_pass(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, void* state, int n_local,
                       uint64_t index_offset, cudaStream_t stream);

// specialised pass kernels (jit.cu)
void qls_jit_release(QlsProgram* prog);
bool qls_jit_eligible(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, int n_local);
int qls_launch_jit_pass(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, void* state, int n_local,
                        uint64_t index_offset, cudaStream_t stream, int sel_pos = -1, int sel_val = 0, int sel_pos2 = -1,
                        int sel_val2 = 0);

// ---- complex arithmetic on double2 / float2 -----------------------------------------------------
template <typename R>
struct Vec2;
template <>
struct Vec2<double> {
  using type = double2;
};
template <>
struct Vec2<float> {
  using type = float2;
};

template <typename V>
__device__ __forceinline__ V cmul(V a, V b) {
  V r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
// acc += a * b
template <typename V>
__device__ __forceinline__ void cfma(V& acc, V a, V b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

__device__ __forceinline__ uint32_t insert_zero_bit(uint32_t x, uint32_t pos) {
  uint32_t low = x & ((1u << pos) - 1u);
  return ((x >> pos) << (pos + 1)) | low;
}

// (not unrolled: n is 1-3 in practice and a fully unrolled, predicated loop over the 6/16 possible
// segments costs ~150 issue slots per call -- more than the table lookups it serves)
__device__ __forceinline__ uint64_t gather_segs64(uint64_t x, const QlsSeg* s, int n) {
  uint64_t r = 0;
#pragma unroll 1
  for (int i = 0; i < n; ++i) r |= ((x >> s[i].src) & ((1ull << s[i].width) - 1ull)) << s[i].dst;
  return r;
}
__device__ __forceinline__ uint32_t gather_segs32(uint32_t x, const QlsSeg* s, int n) {
  uint32_t r = 0;
#pragma unroll 1
  for (int i = 0; i < n; ++i) r |= ((x >> s[i].src) & ((1u << s[i].width) - 1u)) << s[i].dst;
  return r;
}

// 16-byte async copy global -> shared (LDGSTS), bypassing L1
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// rows r < total_rows of state.reshape(-1, 2^nb) (complex128) with (r & row_mask) == row_val are replaced by row @ U^T,
// in place on the FP64 tensor pipe (ctrl_gemm.cu; 6 <= nb <= 10)
int qls_gemm_rows_inplace(void* state, uint64_t total_rows, int nb, uint64_t row_mask, uint64_t row_val, const void* u_dev,
                          cudaStream_t stream);

// A __constant__ image that is refreshed before every launch (register-stage pass description, wide dense matrices of the sweep
// kernel) exists once per device: two launches that use it must not overlap.  Launches on ONE stream are ordered anyway; a
// launch on another stream first waits for the previous user (event recorded after it), so the constraint is enforced here
// instead of being left to the caller.  Call `before` ahead of the cudaMemcpyToSymbolAsync and `after` behind the kernel launch.
struct QlsConstGuard {
  cudaEvent_t ev[64] = {};
  cudaStream_t last[64] = {};
  bool used[64] = {};
  int before(int device, cudaStream_t s) {
    const int d = device & 63;
    if (used[d] && last[d] != s) {
      cudaError_t e = cudaStreamWaitEvent(s, ev[d], 0);
      if (e != cudaSuccess) return (int)e;
    }
    return 0;
  }
  int after(int device, cudaStream_t s) {
    const int d = device & 63;
    if (!ev[d]) {
      cudaError_t e = cudaEventCreateWithFlags(&ev[d], cudaEventDisableTiming);
      if (e != cudaSuccess) return (int)e;
    }
    cudaError_t e = cudaEventRecord(ev[d], s);
    if (e != cudaSuccess) return (int)e;
    used[d] = true;
    last[d] = s;
    return 0;
  }
};
