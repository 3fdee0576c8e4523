// Dense controlled-e^{iAt} block: selected rows of psi.reshape(-1, 2^nb) are replaced by row @ U^T.
// v0: shared-memory tiled complex GEMM on the FP64/FP32 vector pipes, out of place into `scratch`,
// then a copy-back.  (The FP64 tensor-pipe version replaces the inner product; same interface.)
#include "qls_common.cuh"

struct RowSel {
  int n_seg;
  QlsSeg seg[QLS_MAX_FSEG];
  uint64_t fix_val;  // row-index bits fixed by the controls
};

constexpr int GM = 64, GN = 64, GK = 16;  // C tile GM x GN, K step GK; 256 threads, 4x4 outputs each

template <typename V>
__global__ void __launch_bounds__(256)
    qls_ctrl_gemm_kernel(const V* __restrict__ psi, V* __restrict__ out, const V* __restrict__ U, int nb, RowSel rs,
                         uint64_t n_rows) {
  __shared__ V As[GK][GM + 1];
  __shared__ V Bs[GK][GN + 1];
  const int N = 1 << nb;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx -> output column group (n), ty -> row group (m)
  const uint64_t m0 = (uint64_t)blockIdx.y * GM;
  const int n0 = blockIdx.x * GN;
  V acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j].x = acc[i][j].y = 0;

  for (int k0 = 0; k0 < N; k0 += GK) {
    // A tile: rows m0..m0+GM (selected psi rows), cols k0..k0+GK  (coalesced along k)
    for (int e = tid; e < GM * GK; e += 256) {
      const int mm = e / GK, kk = e % GK;
      const uint64_t m = m0 + mm;
      V v;
      v.x = v.y = 0;
      if (m < n_rows) {
        const uint64_t row = gather_segs64(m, rs.seg, rs.n_seg) | rs.fix_val;
        v = psi[(row << nb) + k0 + kk];
      }
      As[kk][mm] = v;
    }
    for (int e = tid; e < GN * GK; e += 256) {
      const int nn = e / GK, kk = e % GK;
      Bs[kk][nn] = U[(size_t)(n0 + nn) * N + k0 + kk];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      V a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cfma(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t m = m0 + ty * 4 + i;
    if (m >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(m << nb) + n0 + tx * 4 + j] = acc[i][j];
  }
}

template <typename V>
__global__ void __launch_bounds__(256)
    qls_ctrl_gemm_copyback_kernel(V* __restrict__ psi, const V* __restrict__ out, int nb, RowSel rs, uint64_t n_elems) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t cmask = (1ull << nb) - 1ull;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += stride) {
    const uint64_t m = e >> nb;
    const uint64_t row = gather_segs64(m, rs.seg, rs.n_seg) | rs.fix_val;
    psi[(row << nb) | (e & cmask)] = out[e];
  }
}

extern "C" int qls_ctrl_gemm(void* state, void* scratch, int n_local, int dtype, int nb, uint64_t ctrl_mask,
                             uint64_t ctrl_val, uint64_t index_offset, const void* u_dev, void* stream) {
  QLS_ARG(state && scratch && u_dev, "null pointer");
  QLS_ARG(nb >= 6 && nb <= n_local && nb <= 14, "nb=%d unsupported (need 6 <= nb <= min(14, n_local))", nb);
  QLS_ARG((ctrl_val & ~ctrl_mask) == 0, "ctrl_val has bits outside ctrl_mask");
  QLS_ARG((ctrl_mask & ((1ull << nb) - 1ull)) == 0, "controls overlap the target register");
  const uint64_t local_all = (n_local >= 64) ? ~0ull : ((1ull << n_local) - 1ull);
  // controls on rank (global) bits: a predicate on this shard
  if ((index_offset & ctrl_mask & ~local_all) != (ctrl_val & ~local_all)) return QLS_OK;
  const uint64_t lmask = (ctrl_mask & local_all) >> nb, lval = (ctrl_val & local_all) >> nb;
  const int row_bits = n_local - nb;
  RowSel rs;
  memset(&rs, 0, sizeof(rs));
  rs.fix_val = lval;
  int dst = 0, src = 0;
  while (dst < row_bits) {  // contiguous runs of free row bits
    if ((lmask >> dst) & 1ull) {
      ++dst;
      continue;
    }
    int w = 0;
    while (dst + w < row_bits && !((lmask >> (dst + w)) & 1ull)) ++w;
    QLS_ARG(rs.n_seg < QLS_MAX_FSEG, "too many control fragments");
    rs.seg[rs.n_seg].src = (uint8_t)src;
    rs.seg[rs.n_seg].width = (uint8_t)w;
    rs.seg[rs.n_seg].dst = (uint8_t)dst;
    ++rs.n_seg;
    src += w;
    dst += w;
  }
  const uint64_t n_rows = 1ull << src;
  const int N = 1 << nb;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(N / GN, (unsigned)((n_rows + GM - 1) / GM));
  const uint64_t n_elems = n_rows << nb;
  uint64_t cb = (n_elems + 255) / 256;
  unsigned cgrid = (unsigned)(cb > 148ull * 16 ? 148 * 16 : cb);
  if (dtype == QLS_C128) {
    qls_ctrl_gemm_kernel<double2><<<grid, 256, 0, s>>>(reinterpret_cast<const double2*>(state),
                                                       reinterpret_cast<double2*>(scratch),
                                                       reinterpret_cast<const double2*>(u_dev), nb, rs, n_rows);
    qls_ctrl_gemm_copyback_kernel<double2><<<cgrid, 256, 0, s>>>(reinterpret_cast<double2*>(state),
                                                                 reinterpret_cast<const double2*>(scratch), nb, rs, n_elems);
  } else {
    qls_ctrl_gemm_kernel<float2><<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(state),
                                                      reinterpret_cast<float2*>(scratch),
                                                      reinterpret_cast<const float2*>(u_dev), nb, rs, n_rows);
    qls_ctrl_gemm_copyback_kernel<float2><<<cgrid, 256, 0, s>>>(reinterpret_cast<float2*>(state),
                                                                reinterpret_cast<const float2*>(scratch), nb, rs, n_elems);
  }
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}
