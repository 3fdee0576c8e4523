// Streaming tile-pass kernel: one HBM read + one HBM write per touched amplitude, any number of
// fused blocks applied in shared memory in between.
//
// A tile is 2^T amplitudes: the low L qubits (contiguous runs of 2^L amplitudes -> coalesced,
// 128-bit accesses) plus n_high arbitrary higher qubits (2^n_high runs gathered from strided
// addresses).  One CTA stages one tile with 16-byte cp.async (LDGSTS, L1 bypass), applies the
// pass's ops (tile_ops.cuh), and writes the runs back with 128-bit stores.  Several CTAs are
// resident per SM so that one tile's loads overlap another tile's math and stores.
#include "tile_ops.cuh"

struct TileLaunch {
  int T, L, n_high, n_fseg;
  uint8_t high[QLS_MAX_HIGH];
  QlsSeg fseg[QLS_MAX_FSEG];
  uint64_t fix_val;
  uint64_t index_offset;
  uint32_t n_ops;
  uint32_t pad;
  uint16_t heavy_off[QLS_MAX_HEAVY];  // element offsets into c_heavy of the pass's k >= 4 dense matrices (0xFFFF: in the pool)
};

template <typename R, int THREADS, bool HEAVY>
__global__ void __launch_bounds__(THREADS, HEAVY ? 2 : 3)
    qls_tile_pass_kernel(typename Vec2<R>::type* __restrict__ state, const TileLaunch tl,
                         const QlsOp* __restrict__ ops, const unsigned char* __restrict__ pool) {
  using V = typename Vec2<R>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* sm = reinterpret_cast<V*>(smem_raw);
  const int tid = threadIdx.x;
  const int T = tl.T, L = tl.L;

  // Per-CTA index tables (the descriptors are byte arrays in the kernel parameters: walking
  // them per amplitude costs more L1 traffic than the tile): hoff[r] = spread of the high-qubit
  // combination r, s_base = this tile's base index (tile bits zero, pass-level controls fixed).
  uint64_t* hoff = reinterpret_cast<uint64_t*>(smem_raw + (sizeof(V) << T));
  uint64_t* s_base = hoff + (1 << tl.n_high);
  __shared__ QlsSeg s_fseg[QLS_MAX_FSEG];  // copied with compile-time indices: keeps `tl` in the constant bank
  __shared__ uint8_t s_high[QLS_MAX_HIGH];
  if (tid == 0) {
#pragma unroll
    for (int j = 0; j < QLS_MAX_FSEG; ++j) s_fseg[j] = tl.fseg[j];
#pragma unroll
    for (int j = 0; j < QLS_MAX_HIGH; ++j) s_high[j] = tl.high[j];
  }
  __syncthreads();
  for (int r = tid; r < (1 << tl.n_high); r += THREADS) {
    uint64_t g = 0;
    for (int j = 0; j < tl.n_high; ++j) g |= (uint64_t)((r >> j) & 1) << s_high[j];
    hoff[r] = g;
  }
  if (tid == 0) s_base[0] = gather_segs64((uint64_t)blockIdx.x, s_fseg, tl.n_fseg) | tl.fix_val;
  __syncthreads();
  const uint64_t base = s_base[0];

  // 16-byte units: one complex128 or two complex64
  constexpr int APU = 16 / (int)sizeof(V);      // amplitudes per unit
  const int units = (1 << T) / APU;
  const int lmask = (1 << L) - 1;
  uint4* sm16 = reinterpret_cast<uint4*>(smem_raw);

  for (int u = tid; u < units; u += THREADS) {
    const int i = u * APU;
    const uint64_t g = base | (uint64_t)(i & lmask) | hoff[i >> L];
    cp_async16(&sm16[u], &state[g]);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  apply_ops<V, THREADS, HEAVY>(sm, T, ops, (int)tl.n_ops, pool, base | tl.index_offset, tid, HEAVY ? tl.heavy_off : nullptr);

  for (int u = tid; u < units; u += THREADS) {
    const int i = u * APU;
    const uint64_t g = base | (uint64_t)(i & lmask) | hoff[i >> L];
    *reinterpret_cast<uint4*>(&state[g]) = sm16[u];
  }
}

// ---- host side --------------------------------------------------------------------------------------

static QlsConstGuard g_heavy_guard;  // one c_heavy table per device: launches that use it never overlap

template <typename R, bool HEAVY>
static int launch_pass(const QlsProgram* prog, const QlsPass& p, void* state, int n_local, uint64_t index_offset,
                       cudaStream_t stream) {
  constexpr int THREADS = 256;
  using V = typename Vec2<R>::type;
  TileLaunch tl;
  memset(&tl, 0, sizeof(tl));
  tl.T = p.tile_bits;
  tl.L = p.low_bits;
  tl.n_high = p.n_high;
  tl.n_fseg = p.n_fseg;
  memcpy(tl.high, p.high, sizeof(tl.high));
  memcpy(tl.fseg, p.fseg, sizeof(tl.fseg));
  tl.fix_val = p.fix_val;
  tl.index_offset = index_offset;
  tl.n_ops = p.op_count;
  for (int j = 0; j < QLS_MAX_HEAVY; ++j) tl.heavy_off[j] = 0xFFFFu;
  bool uses_const = false;
  if (HEAVY && sizeof(V) == 16) {
    QLS_ARG(g_heavy_guard.before(prog->device, stream) == 0, "stream wait on the previous user of the matrix table failed");
    uses_const = true;
    // matrices of the wide dense blocks -> __constant__ table (device to device, stream ordered)
    uint32_t used = 0;
    int seen = 0;
    for (uint32_t o = p.op_begin; o < p.op_begin + p.op_count && seen < QLS_MAX_HEAVY; ++o) {
      const QlsOp& op = prog->ops_host[o];
      if (op.kind != QLS_OP_DENSE || (op.flags & QLS_OP_FLAG_PARAM) || op.k < 4) continue;
      const uint32_t elems = 1u << (2 * op.k);
      if (used + elems <= QLS_HEAVY_CONST) {
        QLS_CUDA(cudaMemcpyToSymbolAsync(c_heavy, prog->pool_dev + op.pool_off, (size_t)elems * 16, (size_t)used * 16,
                                         cudaMemcpyDeviceToDevice, stream));
        tl.heavy_off[seen] = (uint16_t)used;
        used += elems;
      }
      ++seen;
    }
  }
  const int fixed = __builtin_popcountll(p.fix_mask);
  const int free_bits = n_local - p.tile_bits - fixed;
  QLS_ARG(free_bits >= 0 && free_bits < 31, "pass geometry: n_local=%d tile_bits=%d fixed=%d", n_local, p.tile_bits, fixed);
  const size_t smem = ((size_t)sizeof(V) << p.tile_bits) + (sizeof(uint64_t) << p.n_high) + 2 * sizeof(uint64_t);
  auto kern = qls_tile_pass_kernel<R, THREADS, HEAVY>;
  static bool attr_set[64] = {false};  // per template instantiation and device
  if (!attr_set[prog->device & 63]) {
    QLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[prog->device & 63] = true;
  }
  const unsigned grid = 1u << free_bits;
  kern<<<grid, THREADS, smem, stream>>>(reinterpret_cast<V*>(state), tl, prog->ops_dev + p.op_begin, prog->pool_dev);
  QLS_CUDA(cudaGetLastError());
  if (uses_const) QLS_ARG(g_heavy_guard.after(prog->device, stream) == 0, "event record behind the pass kernel failed");
  return QLS_OK;
}

static int validate_pass(const QlsPass& p, int n_local, int dtype) {
  QLS_ARG(p.tile_bits >= 1 && p.tile_bits <= 13, "tile_bits %d out of range", p.tile_bits);
  QLS_ARG(p.tile_bits <= n_local, "tile_bits %d > n_local %d", p.tile_bits, n_local);
  QLS_ARG(p.n_high >= 0 && p.n_high <= QLS_MAX_HIGH && p.low_bits + p.n_high == p.tile_bits, "bad tile split");
  QLS_ARG(p.low_bits >= (dtype == QLS_C64 ? 1 : 0), "low_bits too small for 16-byte units");
  QLS_ARG(p.n_fseg >= 0 && p.n_fseg <= QLS_MAX_FSEG, "bad n_fseg");
  for (int j = 0; j < p.n_high; ++j)
    QLS_ARG(p.high[j] >= p.low_bits && p.high[j] < n_local, "high qubit %d out of range", (int)p.high[j]);
  return QLS_OK;
}

static int g_sweep_only = 0;
extern "C" int qls_set_sweep_only(int on) {
  g_sweep_only = on ? 1 : 0;
  return QLS_OK;
}

static int g_use_jit = 1;
extern "C" int qls_set_use_jit(int on) {
  g_use_jit = on ? 1 : 0;
  return QLS_OK;
}

// One pass restricted to the half of the state whose index bit `sel_pos` equals `sel_val` (a position the pass neither
// transforms nor fixes).  Specialised kernels only: QLS_ERR_UNSUPPORTED otherwise, and the caller runs the pass whole.
extern "C" int qls_program_run_part(const QlsProgram* prog, void* state, int n_local, uint64_t index_offset, uint32_t pass_index,
                                    int sel_pos, int sel_val, int sel_pos2, int sel_val2, void* stream) {
  QLS_ARG(prog && state, "null program or state");
  QLS_ARG(pass_index < prog->n_passes, "pass %u outside program of %u", pass_index, prog->n_passes);
  QLS_ARG(sel_pos >= 0 && sel_pos < n_local && (sel_val == 0 || sel_val == 1), "bad selection bit %d = %d", sel_pos, sel_val);
  QLS_ARG(sel_pos2 < n_local && sel_pos2 != sel_pos && (sel_val2 == 0 || sel_val2 == 1), "bad second selection bit %d = %d", sel_pos2,
          sel_val2);
  QLS_CUDA(cudaSetDevice(prog->device));
  const QlsPass& p = prog->passes_host[pass_index];
  int rc = validate_pass(p, n_local, prog->dtype);
  if (rc) return rc;
  if (g_sweep_only || !g_use_jit || !qls_jit_eligible(prog, p, pass_index, n_local)) {
    qls_set_error("pass %u has no specialised kernel: it cannot run on half of the state", pass_index);
    return QLS_ERR_UNSUPPORTED;
  }
  return qls_launch_jit_pass(prog, p, pass_index, state, n_local, index_offset, reinterpret_cast<cudaStream_t>(stream), sel_pos,
                             sel_val, sel_pos2 < 0 ? -1 : sel_pos2, sel_val2);
}

extern "C" int qls_program_run(const QlsProgram* prog, void* state, int n_local, uint64_t index_offset,
                               uint32_t pass_begin, uint32_t pass_end, void* stream) {
  QLS_ARG(prog && state, "null program or state");
  QLS_ARG(pass_begin <= pass_end && pass_end <= prog->n_passes, "pass range [%u,%u) outside program of %u", pass_begin,
          pass_end, prog->n_passes);
  QLS_CUDA(cudaSetDevice(prog->device));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  for (uint32_t i = pass_begin; i < pass_end; ++i) {
    const QlsPass& p = prog->passes_host[i];
    int rc = validate_pass(p, n_local, prog->dtype);
    if (rc) return rc;
    QLS_ARG(!(p.pad & 4u), "pass %u holds batch-only ops (PARAM / MATVEC): use qls_batch_expect_z", i);
    if (!g_sweep_only && g_use_jit && qls_jit_eligible(prog, p, i, n_local)) {
      rc = qls_launch_jit_pass(prog, p, i, state, n_local, index_offset, s);
      if (rc) return rc;
      continue;
    }
    if (!g_sweep_only && qls_rs_eligible(prog, p, i, n_local)) {
      rc = qls_launch_rs_pass(prog, p, i, state, n_local, index_offset, s);
      if (rc) return rc;
      continue;
    }
    bool heavy = (p.pad & 1u) != 0;  // set by qls_program_create when the pass holds a k >= 4 dense op
    if (prog->dtype == QLS_C128)
      rc = heavy ? launch_pass<double, true>(prog, p, state, n_local, index_offset, s)
                 : launch_pass<double, false>(prog, p, state, n_local, index_offset, s);
    else
      rc = heavy ? launch_pass<float, true>(prog, p, state, n_local, index_offset, s)
                 : launch_pass<float, false>(prog, p, state, n_local, index_offset, s);
    if (rc) return rc;
  }
  return QLS_OK;
}

extern "C" int qls_program_create(const QlsPass* passes_host, uint32_t n_passes, const QlsOp* ops_host, uint32_t n_ops,
                                  const QlsRsOp* rs_ops_host, uint32_t n_rs_ops, const void* pool_host,
                                  uint64_t pool_bytes, int dtype, int device, QlsProgram** out) {
  QLS_ARG(out, "null out");
  QLS_ARG(dtype == QLS_C128 || dtype == QLS_C64, "bad dtype %d", dtype);
  QLS_ARG(n_passes == 0 || passes_host, "null passes");
  QLS_ARG(n_ops == 0 || ops_host, "null ops");
  QLS_ARG(n_rs_ops == 0 || rs_ops_host, "null register-stage ops");
  QLS_CUDA(cudaSetDevice(device));
  const size_t esz = dtype == QLS_C128 ? 16 : 8;
  for (uint32_t i = 0; i < n_ops; ++i) {
    const QlsOp& op = ops_host[i];
    QLS_ARG(op.kind >= QLS_OP_DENSE && op.kind <= QLS_OP_MATVEC, "op %u: bad kind %d", i, op.kind);
    QLS_ARG(op.n_lseg <= QLS_MAX_SEG && op.n_xseg <= QLS_MAX_SEG, "op %u: too many segments", i);
    QLS_ARG(op.pool_off % esz == 0 && op.pool_off < pool_bytes + (pool_bytes == 0), "op %u: bad pool offset", i);
    if (op.kind == QLS_OP_MATVEC) {
      QLS_ARG(op.k >= 1 && op.k <= 13, "op %u: matvec k=%d unsupported", i, op.k);
      QLS_ARG(op.pool_off + (esz << (2 * op.k)) <= pool_bytes, "op %u: matrix outside pool", i);
    } else if (op.kind == QLS_OP_DENSE && (op.flags & QLS_OP_FLAG_PARAM)) {
      QLS_ARG(op.k == 1 && op.pad0 <= 3, "op %u: bad parameterised rotation", i);
    } else if (op.kind == QLS_OP_DENSE) {
      QLS_ARG(op.k >= 1 && op.k <= 5, "op %u: dense k=%d unsupported", i, op.k);
      QLS_ARG(op.pool_off + (esz << (2 * op.k)) <= pool_bytes, "op %u: matrix outside pool", i);
      for (int j = 1; j < op.k; ++j) QLS_ARG(op.tq[j] > op.tq[j - 1], "op %u: targets not ascending", i);
    }
  }
  for (uint32_t i = 0; i < n_rs_ops; ++i) {
    const QlsRsOp& op = rs_ops_host[i];
    QLS_ARG(op.kind >= QLS_RS_STAGE && op.kind <= QLS_RS_MUXRY, "rs op %u: bad kind %d", i, op.kind);
    QLS_ARG(op.n_tseg <= QLS_MAX_SEG && op.n_xseg <= QLS_MAX_SEG, "rs op %u: too many segments", i);
    if (op.kind != QLS_RS_STAGE) QLS_ARG(op.pool_off % esz == 0 && op.pool_off < pool_bytes, "rs op %u: bad pool offset", i);
    if (op.kind == QLS_RS_G2) QLS_ARG(op.r[0] < op.r[1] && op.r[1] < 5, "rs op %u: bad register qubits", i);
    if (op.kind == QLS_RS_G1 || op.kind == QLS_RS_MUXRY) QLS_ARG(op.r[0] < 5, "rs op %u: bad register qubit", i);
  }
  QlsProgram* p = new QlsProgram();
  p->device = device;
  p->rs_ops_dev = nullptr;
  p->n_rs_ops = n_rs_ops;
  p->dtype = dtype;
  p->n_passes = n_passes;
  p->n_ops = n_ops;
  p->pool_bytes = pool_bytes;
  p->passes_host = new QlsPass[n_passes ? n_passes : 1];
  p->ops_dev = nullptr;
  p->ops_host = new QlsOp[n_ops ? n_ops : 1];
  for (uint32_t i = 0; i < n_ops; ++i) p->ops_host[i] = ops_host[i];
  p->pool_dev = nullptr;
  p->rs_img_dev = nullptr;
  p->rs_img_off = nullptr;
  p->rs_img_bytes = nullptr;
  p->rs_img_kr = nullptr;
  p->jit = nullptr;
  p->batch_scratch = nullptr;
  p->batch_scratch_bytes = 0;
  for (uint32_t i = 0; i < n_passes; ++i) {
    p->passes_host[i] = passes_host[i];
    QlsPass& ps = p->passes_host[i];
    if ((uint64_t)ps.op_begin + ps.op_count > n_ops) {
      qls_set_error("pass %u: op range outside program", i);
      delete[] p->passes_host;
      delete[] p->ops_host;
      delete p;
      return QLS_ERR_ARG;
    }
    bool heavy = false, matvec = false, batch_only = false;
    for (uint32_t o = ps.op_begin; o < ps.op_begin + ps.op_count; ++o) {
      const QlsOp& op = ops_host[o];
      if (op.kind == QLS_OP_DENSE && op.k >= 4) heavy = true;
      if (op.kind == QLS_OP_MATVEC) matvec = batch_only = true;
      if (op.kind == QLS_OP_DENSE && (op.flags & QLS_OP_FLAG_PARAM)) batch_only = true;
      for (int j = 0; j < (op.kind == QLS_OP_DIAG ? 0 : op.k); ++j)
        if (op.tq[j] >= ps.tile_bits) {
          qls_set_error("pass %u op %u: target bit %d outside tile of %d bits", i, o, (int)op.tq[j], ps.tile_bits);
          delete[] p->passes_host;
          delete[] p->ops_host;
          delete p;
          return QLS_ERR_ARG;
        }
    }
    uint32_t n_stages = 0;
    if ((uint64_t)ps.rs_begin + ps.rs_count > n_rs_ops) {
      qls_set_error("pass %u: register-stage range outside program", i);
      delete[] p->passes_host;
      delete[] p->ops_host;
      delete p;
      return QLS_ERR_ARG;
    }
    for (uint32_t o = ps.rs_begin; o < ps.rs_begin + ps.rs_count; ++o)
      if (rs_ops_host[o].kind == QLS_RS_STAGE) ++n_stages;
    if (ps.rs_count && rs_ops_host[ps.rs_begin].kind != QLS_RS_STAGE) n_stages = 0;  // malformed: disables the form
    ps.pad = (heavy ? 1u : 0u) | (matvec ? 2u : 0u) | (batch_only ? 4u : 0u) | (n_stages << 8);
  }
  cudaError_t e = cudaSuccess;
  if (n_ops) {
    e = cudaMalloc(&p->ops_dev, sizeof(QlsOp) * (size_t)n_ops);
    if (e == cudaSuccess) e = cudaMemcpy(p->ops_dev, ops_host, sizeof(QlsOp) * (size_t)n_ops, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && n_rs_ops) {
    e = cudaMalloc(&p->rs_ops_dev, sizeof(QlsRsOp) * (size_t)n_rs_ops);
    if (e == cudaSuccess)
      e = cudaMemcpy(p->rs_ops_dev, rs_ops_host, sizeof(QlsRsOp) * (size_t)n_rs_ops, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && pool_bytes) {
    e = cudaMalloc(&p->pool_dev, pool_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(p->pool_dev, pool_host, pool_bytes, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && qls_rs_build_images(p, rs_ops_host, static_cast<const unsigned char*>(pool_host)) != QLS_OK)
    e = cudaErrorUnknown;
  if (e != cudaSuccess) {
    if (e != cudaErrorUnknown) qls_set_error("qls_program_create: %s", cudaGetErrorString(e));
    if (p->rs_img_dev) cudaFree(p->rs_img_dev);
    delete[] p->rs_img_off;
    delete[] p->rs_img_bytes;
    delete[] p->rs_img_kr;
    if (p->ops_dev) cudaFree(p->ops_dev);
    if (p->rs_ops_dev) cudaFree(p->rs_ops_dev);
    if (p->pool_dev) cudaFree(p->pool_dev);
    delete[] p->passes_host;
    delete[] p->ops_host;
    delete p;
    return e == cudaErrorMemoryAllocation ? QLS_ERR_NOMEM : QLS_ERR_CUDA;
  }
  *out = p;
  return QLS_OK;
}

extern "C" int qls_program_destroy(QlsProgram* prog) {
  if (!prog) return QLS_OK;
  cudaSetDevice(prog->device);
  if (prog->ops_dev) cudaFree(prog->ops_dev);
  if (prog->rs_ops_dev) cudaFree(prog->rs_ops_dev);
  if (prog->pool_dev) cudaFree(prog->pool_dev);
  if (prog->rs_img_dev) cudaFree(prog->rs_img_dev);
  if (prog->batch_scratch) cudaFree(prog->batch_scratch);
  qls_jit_release(prog);
  delete[] prog->rs_img_off;
  delete[] prog->rs_img_bytes;
  delete[] prog->rs_img_kr;
  delete[] prog->passes_host;
  delete[] prog->ops_host;
  delete prog;
  return QLS_OK;
}
