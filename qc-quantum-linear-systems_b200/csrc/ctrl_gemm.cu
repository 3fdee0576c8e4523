// Dense controlled-e^{iAt} block: selected rows of psi.reshape(-1, 2^nb) are replaced by row @ U^T.
//   complex128: FP64 tensor pipe.  The complex product is one REAL GEMM
//       C[M x 2N] = A[M x 2K] * B'[2K x 2N],   B'[2k+c][2n+d] = (Re U[n][k], -Im; Im, Re)
//   (A = the interleaved (re, im) rows of psi as they lie in HBM, C = interleaved output), issued
//   as mma.sync.aligned.m8n8k4.f64 (DMMA).  (tcgen05.mma has no f64 kind; DMMA is the FP64 tensor
//   path on sm_100a.)
//   * nb <= 10: IN PLACE.  A thread-block cluster owns 64 whole rows: CTA c of the cluster computes
//     columns [64 c, 64 c + 64) of the product for all of them (64 x 128 real block tile, 8 warps of
//     32 x 32, A tile k-major and both tiles padded by 4 doubles so that every fragment load is
//     bank-conflict free, register double buffering of the global loads); a cluster barrier after the
//     K loop -- every CTA of the cluster has then read the rows for the last time -- and the tile is
//     written back over the rows it came from.  No scratch state, no copy-back pass.  N = 1024 needs a
//     cluster of 16 (non-portable size, opted in).
//   * nb >= 11: the rows do not fit one cluster: out of place into `scratch` (128 x 64 tile), then a
//     copy-back of the selected rows.
//   complex64: shared-memory tiled complex GEMM on the FP32 vector pipe (not a benchmarked path).
#include <cooperative_groups.h>

#include "qls_common.cuh"

struct RowSel {
  int n_seg;
  QlsSeg seg[QLS_MAX_FSEG];
  uint64_t fix_val;  // row-index bits fixed by the controls
};

namespace cg = cooperative_groups;

constexpr int GM = 64, GN = 64, GK = 16;  // C tile GM x GN, K step GK; 256 threads, 4x4 outputs each

template <typename V>
__global__ void __launch_bounds__(256)
    qls_ctrl_gemm_kernel(const V* __restrict__ psi, V* __restrict__ out, const V* __restrict__ U, int nb, RowSel rs,
                         uint64_t m_base, uint64_t n_rows) {
  __shared__ V As[GK][GM + 1];
  __shared__ V Bs[GK][GN + 1];
  const int N = 1 << nb;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx -> output column group (n), ty -> row group (m)
  const uint64_t m0 = m_base + (uint64_t)blockIdx.y * GM;  // rows [m_base, n_rows) of the selection; out holds them from 0
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
    for (int j = 0; j < 4; ++j) out[((m - m_base) << nb) + n0 + tx * 4 + j] = acc[i][j];
  }
}

// ---- FP64 tensor-pipe version -----------------------------------------------------------------------
constexpr int DM = 128, DN = 64, DK = 16;  // real-GEMM block tile (M x N' x K'), N' and K' count doubles
constexpr int LDAT = DM + 4, LDB = DN + 4;

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
    qls_ctrl_gemm_dmma_kernel(const double2* __restrict__ psi, double2* __restrict__ out, const double2* __restrict__ U,
                              int nb, RowSel rs, uint64_t m_base, uint64_t n_rows) {
  extern __shared__ __align__(16) double dsm[];
  double (*As)[DK][LDAT] = reinterpret_cast<double (*)[DK][LDAT]>(dsm);                  // [buf][k'][m]
  double (*Bs)[DK][LDB] = reinterpret_cast<double (*)[DK][LDB]>(dsm + 2 * DK * LDAT);   // [buf][k'][n']
  const int N = 1 << nb;              // complex columns of psi rows = complex rows/cols of U
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;  // warp tile origin inside the block tile
  const uint64_t m0 = m_base + (uint64_t)blockIdx.y * DM;  // rows [m_base, n_rows) of the selection; out holds them from 0
  const int n0c = blockIdx.x * (DN / 2);  // complex column origin

  // global -> register staging: A: 128 rows x 8 complex = 4 per thread; B: 32 complex rows of U x 8 complex = 1 per thread
  const int a_k = tid & 7;   // complex k within the tile
  const int a_m = tid >> 3;  // 0..31, rows a_m + 32*j
  const double2* a_ptr[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint64_t m = m0 + a_m + 32 * j;
    const uint64_t row = gather_segs64(m < n_rows ? m : 0, rs.seg, rs.n_seg) | rs.fix_val;
    a_ptr[j] = psi + (row << nb) + a_k;
  }
  const int b_k = tid & 7, b_n = tid >> 3;  // U[n0c + b_n][k0c + b_k]
  const double2* b_ptr = U + (size_t)(n0c + b_n) * N + b_k;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  double2 ra[4], rb;
  auto gload = [&](int kc) {  // kc: complex k origin
#pragma unroll
    for (int j = 0; j < 4; ++j) ra[j] = a_ptr[j][kc];
    rb = b_ptr[kc];
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[buf][2 * a_k][a_m + 32 * j] = ra[j].x;
      As[buf][2 * a_k + 1][a_m + 32 * j] = ra[j].y;
    }
    Bs[buf][2 * b_k][2 * b_n] = rb.x;
    Bs[buf][2 * b_k + 1][2 * b_n] = -rb.y;
    Bs[buf][2 * b_k][2 * b_n + 1] = rb.y;
    Bs[buf][2 * b_k + 1][2 * b_n + 1] = rb.x;
  };

  const int ktiles = N / (DK / 2);
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) gload((kt + 1) * (DK / 2));
#pragma unroll
    for (int k4 = 0; k4 < DK; k4 += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[buf][k4 + t][wm + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[buf][k4 + t][wn + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (kt + 1 < ktiles) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  // C fragment: (row g, real columns 2t, 2t+1) = one complex number
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t m = m0 + wm + 8 * i + g;
    if (m >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nc = n0c + (wn + 8 * j) / 2 + t;
      out[((m - m_base) << nb) + nc] = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

// ---- in-place version: a cluster of gridDim-x-consecutive CTAs owns CM whole rows ---------------------------
// Block tile 64 rows x 128 real columns (= 64 complex columns), 8 warps of 32 x 32: 32 accumulator doubles per thread,
// 128 registers, two CTAs per SM -- a cluster of 16 (N = 1024) then fits a GPC twice.  (A 128 x 128 tile with one CTA per
// SM measured the same tensor-pipe activity while resident, 73 %, but only 7 clusters of 16 were resident at a time:
// 1.83 ms per 2^12 x 2^10 x 2^10 product against 1.31 ms for the out-of-place kernel.)
template <int CM, int CN>
__global__ void __launch_bounds__(256, 2)
    qls_ctrl_gemm_inplace_kernel(double2* __restrict__ psi, const double2* __restrict__ U, int nb, RowSel rs, uint64_t n_rows,
                                 uint64_t row_limit, int col_tiles) {
  constexpr int LDCA = CM + 4, LDCB = CN + 4;
  constexpr int WM = CM / 32;            // warps along m; 8 / WM along n, every warp owns 32 x 32
  constexpr int AJ = CM / 32, BJ = CN / 64;  // double2 loads per thread and k tile: A rows, U rows
  static_assert(WM * 32 == CM && (8 / WM) * 32 == CN, "8 warps of 32 x 32");
  extern __shared__ __align__(16) double dsm[];
  double (*As)[DK][LDCA] = reinterpret_cast<double (*)[DK][LDCA]>(dsm);                   // [buf][k'][m]
  double (*Bs)[DK][LDCB] = reinterpret_cast<double (*)[DK][LDCB]>(dsm + 2 * DK * LDCA);  // [buf][k'][n']
  const int N = 1 << nb;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp % WM) * 32, wn = (warp / WM) * 32;  // warp tile origin inside the block tile (32 x 32)
  const uint64_t m0 = (uint64_t)(blockIdx.x / col_tiles) * CM;
  const int n0c = (int)(blockIdx.x % col_tiles) * (CN / 2);  // complex column origin

  // global -> register staging.  A: CM rows x 8 complex = AJ per thread; B: CN / 2 complex rows of U x 8 complex = BJ per thread
  const int a_k = tid & 7, a_m = tid >> 3;
  const double2* a_ptr[AJ];
  bool a_ok[AJ];
#pragma unroll
  for (int j = 0; j < AJ; ++j) {
    const uint64_t m = m0 + a_m + 32 * j;
    const uint64_t row = gather_segs64(m < n_rows ? m : 0, rs.seg, rs.n_seg) | rs.fix_val;
    a_ok[j] = m < n_rows && row < row_limit;
    a_ptr[j] = psi + ((a_ok[j] ? row : 0) << nb) + a_k;
  }
  const double2* b_ptr[BJ];
#pragma unroll
  for (int j = 0; j < BJ; ++j) b_ptr[j] = U + (size_t)(n0c + a_m + 32 * j) * N + a_k;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  double2 ra[AJ], rb[BJ];
  auto gload = [&](int kc) {
#pragma unroll
    for (int j = 0; j < AJ; ++j) ra[j] = a_ok[j] ? a_ptr[j][kc] : make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < BJ; ++j) rb[j] = __ldg(b_ptr[j] + kc);
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < AJ; ++j) {
      As[buf][2 * a_k][a_m + 32 * j] = ra[j].x;
      As[buf][2 * a_k + 1][a_m + 32 * j] = ra[j].y;
    }
#pragma unroll
    for (int j = 0; j < BJ; ++j) {
      const int n = 2 * (a_m + 32 * j);
      Bs[buf][2 * a_k][n] = rb[j].x;
      Bs[buf][2 * a_k + 1][n] = -rb[j].y;
      Bs[buf][2 * a_k][n + 1] = rb[j].y;
      Bs[buf][2 * a_k + 1][n + 1] = rb[j].x;
    }
  };

  const int ktiles = N / (DK / 2);
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) gload((kt + 1) * (DK / 2));
#pragma unroll
    for (int k4 = 0; k4 < DK; k4 += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[buf][k4 + t][wm + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[buf][k4 + t][wn + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (kt + 1 < ktiles) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  // every CTA of the cluster has read these rows for the last time: they may be overwritten
  cg::this_cluster().sync();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t m = m0 + wm + 8 * i + g;
    if (m >= n_rows) continue;
    const uint64_t row = gather_segs64(m, rs.seg, rs.n_seg) | rs.fix_val;
    if (row >= row_limit) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nc = n0c + (wn + 8 * j) / 2 + t;
      psi[(row << nb) + nc] = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

template <int CM, int CN>
static int launch_inplace(void* state, const void* u_dev, int nb, const RowSel& rs, uint64_t n_rows, uint64_t total_rows,
                          cudaStream_t s) {
  auto kern = qls_ctrl_gemm_inplace_kernel<CM, CN>;
  int col_tiles = (2 << nb) / CN;  // = cluster size
  if (col_tiles < 1) col_tiles = 1;
  QLS_ARG((2 << nb) % CN == 0, "nb=%d too small for a %d-column tile", nb, CN);
  const uint64_t row_tiles = (n_rows + CM - 1) / CM;
  QLS_ARG(row_tiles * col_tiles < (1ull << 31), "too many tiles");
  const size_t dsmem = sizeof(double) * 2 * DK * (CM + 4 + CN + 4);
  static bool configured = false;
  if (!configured) {
    QLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
    QLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(row_tiles * col_tiles), 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = dsmem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)col_tiles;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (getenv("QLS_GEMM_DEBUG")) {
    int nclusters = 0;
    cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    fprintf(stderr, "qls gemm in place: tile %d x %d, cluster %d, grid %u, max active clusters %d\n", CM, CN, col_tiles,
            cfg.gridDim.x, nclusters);
  }
  QLS_CUDA(cudaLaunchKernelEx(&cfg, kern, reinterpret_cast<double2*>(state), reinterpret_cast<const double2*>(u_dev), nb, rs,
                              n_rows, total_rows, col_tiles));
  return QLS_OK;
}

// rows r < total_rows of psi.reshape(-1, 2^nb) with (r & row_mask) == row_val are replaced by row @ U^T, in place
// (complex128, 6 <= nb <= 10).  Used by qls_ctrl_gemm and by the batched estimator (batch_sv.cu), where the "state" is the
// concatenation of the circuits' states.
int qls_gemm_rows_inplace(void* state, uint64_t total_rows, int nb, uint64_t row_mask, uint64_t row_val, const void* u_dev,
                          cudaStream_t s) {
  QLS_ARG(nb >= 6 && nb <= 10, "in-place GEMM needs 6 <= nb <= 10, got %d", nb);
  int row_bits = 0;
  while ((1ull << row_bits) < total_rows) ++row_bits;
  RowSel rs;
  memset(&rs, 0, sizeof(rs));
  rs.fix_val = row_val;
  int dst = 0, src = 0;
  while (dst < row_bits) {  // contiguous runs of free row bits
    if ((row_mask >> dst) & 1ull) {
      ++dst;
      continue;
    }
    int w = 0;
    while (dst + w < row_bits && !((row_mask >> (dst + w)) & 1ull)) ++w;
    QLS_ARG(rs.n_seg < QLS_MAX_FSEG, "too many control fragments");
    rs.seg[rs.n_seg].src = (uint8_t)src;
    rs.seg[rs.n_seg].width = (uint8_t)w;
    rs.seg[rs.n_seg].dst = (uint8_t)dst;
    ++rs.n_seg;
    src += w;
    dst += w;
  }
  const uint64_t n_rows = 1ull << src;
  // tile shape: 64 rows x 64 complex columns per CTA unless that needs a cluster of 16 (nb = 10), where only about half as
  // many CTAs become resident (measured); then 32 rows x 128 complex columns, a cluster of 8.  QLS_GEMM_TILE=64|32 forces one.
  static int force = -1;
  if (force < 0) {
    const char* e = getenv("QLS_GEMM_TILE");
    force = e ? atoi(e) : 0;
  }
  const bool wide = force ? force == 32 : nb >= 10;
  return wide ? launch_inplace<32, 256>(state, u_dev, nb, rs, n_rows, total_rows, s)
              : launch_inplace<64, 128>(state, u_dev, nb, rs, n_rows, total_rows, s);
}

template <typename V>
__global__ void __launch_bounds__(256)
    qls_ctrl_gemm_copyback_kernel(V* __restrict__ psi, const V* __restrict__ out, int nb, RowSel rs, uint64_t m_base,
                                  uint64_t n_elems) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t cmask = (1ull << nb) - 1ull;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += stride) {
    const uint64_t m = m_base + (e >> nb);
    const uint64_t row = gather_segs64(m, rs.seg, rs.n_seg) | rs.fix_val;
    psi[(row << nb) | (e & cmask)] = out[e];
  }
}

extern "C" int qls_ctrl_gemm(void* state, void* scratch, int n_local, int dtype, int nb, uint64_t ctrl_mask,
                             uint64_t ctrl_val, uint64_t index_offset, const void* u_dev, void* stream) {
  QLS_ARG(state && u_dev, "null pointer");
  QLS_ARG(nb >= 6 && nb <= n_local && nb <= 14, "nb=%d unsupported (need 6 <= nb <= min(14, n_local))", nb);
  // scratch == NULL: in place (complex128, nb <= 10: a thread-block cluster owns whole rows).  Otherwise out of place in chunks
  // of QLS_GEMM_CHUNK_ROWS selected rows + copy-back -- the faster form wherever the bounded scratch can be afforded (n = 23,
  // nb = 10: 1.31 + 0.05 ms per product against 1.66 ms in place: clusters of 8-16 CTAs leave SMs of every GPC unused).
  const bool in_place = scratch == nullptr;
  QLS_ARG(!in_place || (dtype == QLS_C128 && nb <= 10), "in place (scratch == NULL) needs complex128 and nb <= 10");
  QLS_ARG((ctrl_val & ~ctrl_mask) == 0, "ctrl_val has bits outside ctrl_mask");
  QLS_ARG((ctrl_mask & ((1ull << nb) - 1ull)) == 0, "controls overlap the target register");
  const uint64_t local_all = (n_local >= 64) ? ~0ull : ((1ull << n_local) - 1ull);
  // controls on rank (global) bits: a predicate on this shard
  if ((index_offset & ctrl_mask & ~local_all) != (ctrl_val & ~local_all)) return QLS_OK;
  const uint64_t lmask = (ctrl_mask & local_all) >> nb, lval = (ctrl_val & local_all) >> nb;
  const int row_bits = n_local - nb;
  if (in_place)
    return qls_gemm_rows_inplace(state, 1ull << row_bits, nb, lmask, lval, u_dev, reinterpret_cast<cudaStream_t>(stream));
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
  const size_t dsmem = sizeof(double) * 2 * DK * (LDAT + LDB);
  if (dtype == QLS_C128)
    QLS_CUDA(cudaFuncSetAttribute(qls_ctrl_gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
  for (uint64_t m_base = 0; m_base < n_rows; m_base += QLS_GEMM_CHUNK_ROWS) {
    const uint64_t rows = n_rows - m_base < QLS_GEMM_CHUNK_ROWS ? n_rows - m_base : (uint64_t)QLS_GEMM_CHUNK_ROWS;
    const uint64_t n_elems = rows << nb;
    const uint64_t cb = (n_elems + 255) / 256;
    const unsigned cgrid = (unsigned)(cb > 148ull * 16 ? 148 * 16 : cb);
    if (dtype == QLS_C128) {
      dim3 dgrid((2 * N) / DN, (unsigned)((rows + DM - 1) / DM));
      qls_ctrl_gemm_dmma_kernel<<<dgrid, 256, dsmem, s>>>(reinterpret_cast<const double2*>(state), reinterpret_cast<double2*>(scratch),
                                                      reinterpret_cast<const double2*>(u_dev), nb, rs, m_base, m_base + rows);
      qls_ctrl_gemm_copyback_kernel<double2><<<cgrid, 256, 0, s>>>(reinterpret_cast<double2*>(state),
                                                                   reinterpret_cast<const double2*>(scratch), nb, rs, m_base, n_elems);
    } else {
      dim3 grid(N / GN, (unsigned)((rows + GM - 1) / GM));
      qls_ctrl_gemm_kernel<float2><<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(state), reinterpret_cast<float2*>(scratch),
                                                        reinterpret_cast<const float2*>(u_dev), nb, rs, m_base, m_base + rows);
      qls_ctrl_gemm_copyback_kernel<float2><<<cgrid, 256, 0, s>>>(reinterpret_cast<float2*>(state),
                                                                  reinterpret_cast<const float2*>(scratch), nb, rs, m_base, n_elems);
    }
  }
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}
