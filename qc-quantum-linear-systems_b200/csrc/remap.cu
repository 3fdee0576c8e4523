// Staging kernels for global-qubit remaps of a sharded statevector: the amplitudes whose local bit
// `qubit` equals `bit_value` are gathered into / scattered from a contiguous buffer that is then
// exchanged with the partner rank (NCCL send/recv or a peer-memory copy over NVLink).
#include <string.h>

#include "qls_common.cuh"

// element j of the slab (j in [0, 2^(n-1))) sits at insert_bit(j, qubit, bit_value) in the state
template <typename V, bool PACK>
__global__ void __launch_bounds__(256) qls_pack_bit_kernel(V* __restrict__ state, V* __restrict__ buffer, int qubit,
                                                            uint64_t bit, uint64_t begin, uint64_t count) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t lowmask = (1ull << qubit) - 1ull;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
    const uint64_t j = begin + t;
    const uint64_t i = ((j >> qubit) << (qubit + 1)) | (bit << qubit) | (j & lowmask);
    if (PACK) buffer[t] = state[i];
    else state[i] = buffer[t];
  }
}

template <bool PACK>
static int pack_impl(void* state, void* buffer, int n_local, int dtype, int qubit, int bit_value, uint64_t begin,
                     uint64_t count, void* stream) {
  QLS_ARG(state && buffer, "null pointer");
  QLS_ARG(qubit >= 0 && qubit < n_local, "qubit %d outside the local register of %d", qubit, n_local);
  QLS_ARG(begin + count <= (1ull << (n_local - 1)), "chunk outside the slab");
  if (count == 0) return QLS_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  uint64_t want = (count + 255) / 256;
  unsigned grid = (unsigned)(want > 148ull * 16 ? 148 * 16 : want);
  if (dtype == QLS_C128)
    qls_pack_bit_kernel<double2, PACK><<<grid, 256, 0, s>>>(reinterpret_cast<double2*>(state),
                                                            reinterpret_cast<double2*>(buffer), qubit,
                                                            (uint64_t)(bit_value & 1), begin, count);
  else
    qls_pack_bit_kernel<float2, PACK><<<grid, 256, 0, s>>>(reinterpret_cast<float2*>(state),
                                                           reinterpret_cast<float2*>(buffer), qubit,
                                                           (uint64_t)(bit_value & 1), begin, count);
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}

extern "C" int qls_pack_bit(const void* state, void* buffer, int n_local, int dtype, int qubit, int bit_value,
                            uint64_t chunk_begin, uint64_t chunk_count, void* stream) {
  return pack_impl<true>(const_cast<void*>(state), buffer, n_local, dtype, qubit, bit_value, chunk_begin, chunk_count, stream);
}
extern "C" int qls_unpack_bit(void* state, const void* buffer, int n_local, int dtype, int qubit, int bit_value,
                              uint64_t chunk_begin, uint64_t chunk_count, void* stream) {
  return pack_impl<false>(state, const_cast<void*>(buffer), n_local, dtype, qubit, bit_value, chunk_begin, chunk_count, stream);
}

// ---- several bits at once (k-qubit remap = one all-to-all instead of k pairwise half-shard swaps) -----

struct QlsBitSet {
  int k;
  int pos[QLS_MAX_REMAP_BITS];  // ascending local bit positions
  uint64_t value;               // OR of (bit value << position)
};

// element j of the sub-block (j in [0, 2^(n-k))) sits at deposit(j) | value: j with a zero bit
// inserted at every position of the set
template <typename V, bool PACK>
__global__ void __launch_bounds__(256) qls_pack_bits_kernel(V* __restrict__ state, V* __restrict__ buffer, const QlsBitSet bs,
                                                             uint64_t begin, uint64_t count) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
    uint64_t i = begin + t;
#pragma unroll
    for (int m = 0; m < QLS_MAX_REMAP_BITS; ++m)
      if (m < bs.k) i = ((i >> bs.pos[m]) << (bs.pos[m] + 1)) | (i & ((1ull << bs.pos[m]) - 1ull));
    i |= bs.value;
    if (PACK) buffer[t] = state[i];
    else state[i] = buffer[t];
  }
}

template <bool PACK>
static int pack_bits_impl(void* state, void* buffer, int n_local, int dtype, const int* qubits, int k, uint64_t value_bits,
                          uint64_t begin, uint64_t count, void* stream) {
  QLS_ARG(state && buffer && qubits, "null pointer");
  QLS_ARG(k >= 1 && k <= QLS_MAX_REMAP_BITS && k <= n_local, "bad bit count %d", k);
  QlsBitSet bs;
  bs.k = k;
  bs.value = 0;
  for (int m = 0; m < k; ++m) {
    QLS_ARG(qubits[m] >= 0 && qubits[m] < n_local && (m == 0 || qubits[m] > qubits[m - 1]), "bit positions must be ascending and local");
    bs.pos[m] = qubits[m];
    bs.value |= ((value_bits >> m) & 1ull) << qubits[m];
  }
  for (int m = k; m < QLS_MAX_REMAP_BITS; ++m) bs.pos[m] = 0;
  QLS_ARG(begin + count <= (1ull << (n_local - k)), "chunk outside the sub-block");
  if (count == 0) return QLS_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  uint64_t want = (count + 255) / 256;
  unsigned grid = (unsigned)(want > 148ull * 16 ? 148 * 16 : want);
  if (dtype == QLS_C128)
    qls_pack_bits_kernel<double2, PACK><<<grid, 256, 0, s>>>(reinterpret_cast<double2*>(state), reinterpret_cast<double2*>(buffer), bs,
                                                             begin, count);
  else
    qls_pack_bits_kernel<float2, PACK><<<grid, 256, 0, s>>>(reinterpret_cast<float2*>(state), reinterpret_cast<float2*>(buffer), bs,
                                                            begin, count);
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}

extern "C" int qls_pack_bits(const void* state, void* buffer, int n_local, int dtype, const int* qubits, int k,
                             uint64_t value_bits, uint64_t chunk_begin, uint64_t chunk_count, void* stream) {
  return pack_bits_impl<true>(const_cast<void*>(state), buffer, n_local, dtype, qubits, k, value_bits, chunk_begin, chunk_count, stream);
}
extern "C" int qls_unpack_bits(void* state, const void* buffer, int n_local, int dtype, const int* qubits, int k,
                               uint64_t value_bits, uint64_t chunk_begin, uint64_t chunk_count, void* stream) {
  return pack_bits_impl<false>(state, const_cast<void*>(buffer), n_local, dtype, qubits, k, value_bits, chunk_begin, chunk_count, stream);
}

// ---- peer-memory exchange: no staging buffers, no NCCL --------------------------------------------
// One pairwise round of a k-qubit exchange as ONE kernel per GPU: element t of my sub-block (local bits =
// my_value) is swapped with element t of the partner's sub-block (local bits = peer_value) through the
// partner's state mapped into this process (CUDA IPC, peer access over NVLink).  Each GPU of the pair swaps
// its own half of the element range ([begin, begin + count)), so every element is touched by exactly one
// GPU: no write-after-read hazard between the two kernels and NVLink is used in both directions by both.
template <typename V>
__global__ void __launch_bounds__(256) qls_swap_peer_kernel(V* __restrict__ mine, V* __restrict__ peer, const QlsBitSet bs,
                                                             uint64_t my_value, uint64_t peer_value, uint64_t begin,
                                                             uint64_t count) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
    uint64_t i = begin + t;
#pragma unroll
    for (int m = 0; m < QLS_MAX_REMAP_BITS; ++m)
      if (m < bs.k) i = ((i >> bs.pos[m]) << (bs.pos[m] + 1)) | (i & ((1ull << bs.pos[m]) - 1ull));
    const V a = mine[i | my_value];
    const V b = peer[i | peer_value];
    mine[i | my_value] = b;
    peer[i | peer_value] = a;
  }
}

extern "C" int qls_swap_peer(void* state, void* peer_state, int n_local, int dtype, const int* qubits, int k,
                             uint64_t my_value_bits, uint64_t peer_value_bits, uint64_t begin, uint64_t count, void* stream) {
  QLS_ARG(state && peer_state && qubits, "null pointer");
  QLS_ARG(k >= 1 && k <= QLS_MAX_REMAP_BITS && k <= n_local, "bad bit count %d", k);
  QlsBitSet bs;
  bs.k = k;
  bs.value = 0;
  uint64_t mv = 0, pv = 0;
  for (int m = 0; m < k; ++m) {
    QLS_ARG(qubits[m] >= 0 && qubits[m] < n_local && (m == 0 || qubits[m] > qubits[m - 1]), "bit positions must be ascending and local");
    bs.pos[m] = qubits[m];
    mv |= ((my_value_bits >> m) & 1ull) << qubits[m];
    pv |= ((peer_value_bits >> m) & 1ull) << qubits[m];
  }
  for (int m = k; m < QLS_MAX_REMAP_BITS; ++m) bs.pos[m] = 0;
  QLS_ARG(begin + count <= (1ull << (n_local - k)), "range outside the sub-block");
  if (count == 0) return QLS_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  uint64_t want = (count + 255) / 256;
  unsigned grid = (unsigned)(want > 148ull * 16 ? 148 * 16 : want);
  if (dtype == QLS_C128)
    qls_swap_peer_kernel<double2><<<grid, 256, 0, s>>>(reinterpret_cast<double2*>(state), reinterpret_cast<double2*>(peer_state), bs, mv,
                                                       pv, begin, count);
  else
    qls_swap_peer_kernel<float2><<<grid, 256, 0, s>>>(reinterpret_cast<float2*>(state), reinterpret_cast<float2*>(peer_state), bs, mv, pv,
                                                      begin, count);
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}

extern "C" int qls_copy_async(void* dst, const void* src, uint64_t bytes, void* stream) {
  QLS_ARG(dst && src, "null pointer");
  if (bytes == 0) return QLS_OK;
  QLS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)));
  return QLS_OK;
}

extern "C" int qls_copy2d_async(void* dst, uint64_t dst_pitch, const void* src, uint64_t src_pitch, uint64_t width_bytes,
                                uint64_t height, void* stream) {
  QLS_ARG(dst && src, "null pointer");
  if (width_bytes == 0 || height == 0) return QLS_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (height == 1 || (dst_pitch == width_bytes && src_pitch == width_bytes)) {
    QLS_CUDA(cudaMemcpyAsync(dst, src, width_bytes * height, cudaMemcpyDefault, s));
    return QLS_OK;
  }
  const uint64_t max_pitch = 0x7fffffffull;  // cudaMemcpy2D's pitch limit; wider strides: one copy per row (rows are huge then)
  if (dst_pitch > max_pitch || src_pitch > max_pitch) {
    for (uint64_t r = 0; r < height; ++r)
      QLS_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + r * dst_pitch, static_cast<const char*>(src) + r * src_pitch, width_bytes,
                               cudaMemcpyDefault, s));
    return QLS_OK;
  }
  QLS_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height, cudaMemcpyDefault, s));
  return QLS_OK;
}

// kernels of `device` may dereference memory of `peer` (mapped through CUDA IPC) once this has succeeded
extern "C" int qls_enable_peer_access(int device, int peer) {
  QLS_CUDA(cudaSetDevice(device));
  int can = 0;
  QLS_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
  QLS_ARG(can, "device %d cannot access device %d directly", device, peer);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();  // clear the sticky-free error state
    return QLS_OK;
  }
  QLS_CUDA(e);
  return QLS_OK;
}

// Open a CUDA IPC memory handle (64 bytes, from cudaIpcGetMemHandle in the exporting process) in the
// context of `device`: the runtime maps the allocation for that device and enables peer access to its
// owner (cudaIpcMemLazyEnablePeerAccess).  *out is the base of the exporting process's allocation.
extern "C" int qls_ipc_open(const void* handle64, int device, void** out) {
  QLS_ARG(handle64 && out, "null pointer");
  QLS_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  QLS_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return QLS_OK;
}
extern "C" int qls_ipc_close(void* ptr) {
  if (ptr) QLS_CUDA(cudaIpcCloseMemHandle(ptr));
  return QLS_OK;
}
