// Batched small-circuit kernel: one CTA evolves one whole statevector (n <= 12 qubits in fp64,
// 13 in fp32) in shared memory and reduces <Z_q> -- the Hadamard tests that VQLS sends through
// Estimator.run (reference call site vqls_qiskit_implementation.py:55-62; SURVEY.md 8a rows
// a10/a11).  All circuits of a launch share one op program; what differs per circuit is the
// parameter vector theta[circuit][*] that the QLS_OP_FLAG_PARAM rotations read.
//
// The in-tile op implementations are the ones of the streaming kernel (tile_ops.cuh); added here:
//   * PARAM   : ry / rx / rz / p with angle = coeff * theta[param_index] + offset (optionally controlled)
//   * MATVEC  : a dense 2^k x 2^k unitary on k contiguous qubits (k up to n), out of place into a
//               second shared-memory buffer -- the controlled A_i / U_b^dagger blocks.
//
// GEMM form (complex128, MATVEC on 6..10 qubits): the 4096 mat-vecs of a launch are ONE product.  The pass is cut at its
// MATVEC ops; a segment kernel (one CTA per circuit, state in shared memory) runs the ops before the cut and writes the
// state into a launch-wide scratch array with the MATVEC's target qubits moved to the low index bits; the scratch array is
// then a statevector of n + log2(circuits) qubits whose rows are multiplied by U^T on the FP64 tensor pipe
// (qls_gemm_rows_inplace, ctrl_gemm.cu: U is read once per 128 rows instead of once per circuit); the next segment kernel
// reads its state back and goes on (or reduces <Z_q>).  The one-kernel form below remains for everything else.
#include "tile_ops.cuh"

enum { PARAM_RY = 0, PARAM_RX = 1, PARAM_RZ = 2, PARAM_P = 3 };

// (c, s) = cos / sin of HALF the angle for ry / rx / rz, of the angle itself for p
template <typename V, typename R, int THREADS>
__device__ __forceinline__ void op_param_cs(V* __restrict__ sm, int T, const QlsOp& op, R c, R s, int tid) {
  V m00, m01, m10, m11;
  switch (op.pad0) {
    case PARAM_RY:
      m00 = {c, 0}; m01 = {-s, 0}; m10 = {s, 0}; m11 = {c, 0};
      break;
    case PARAM_RX:
      m00 = {c, 0}; m01 = {0, -s}; m10 = {0, -s}; m11 = {c, 0};
      break;
    case PARAM_RZ:
      m00 = {c, -s}; m01 = {0, 0}; m10 = {0, 0}; m11 = {c, s};
      break;
    default:  // P(angle) = diag(1, e^{i angle})
      m00 = {1, 0}; m01 = {0, 0}; m10 = {0, 0}; m11 = {c, s};
  }
  const uint32_t tq = op.tq[0];
  const uint32_t lc_mask = op.lc_mask, lc_val = op.lc_val;
  const int n = 1 << (T - 1);
  for (int g = tid; g < n; g += THREADS) {
    const uint32_t i0 = insert_zero_bit(g, tq);
    if ((i0 & lc_mask) != lc_val) continue;
    const uint32_t i1 = i0 | (1u << tq);
    const V a0 = sm[i0], a1 = sm[i1];
    V b0 = cmul(m00, a0), b1 = cmul(m10, a0);
    cfma(b0, m01, a1);
    cfma(b1, m11, a1);
    sm[i0] = b0;
    sm[i1] = b1;
  }
}

template <typename V, typename R, int THREADS>
__device__ __forceinline__ void op_param(V* __restrict__ sm, int T, const QlsOp& op, R angle, int tid) {
  R s, c;
  sincos(op.pad0 == PARAM_P ? angle : angle * R(0.5), &s, &c);
  op_param_cs<V, R, THREADS>(sm, T, op, c, s, tid);
}

// out[o | r << lo] = sum_c U[r][c] in[o | c << lo] for every outer index o whose controls hold;
// other amplitudes are copied.  One warp per output row r, lanes stride the columns (coalesced
// 512-byte reads of the U row), shuffle reduction; up to 4 outer indices share each U read.
template <typename V, int THREADS>
__device__ __forceinline__ void op_matvec(const V* __restrict__ in, V* __restrict__ out, int T, const QlsOp& op,
                                          const V* __restrict__ U, int tid) {
  const int k = op.k, lo = op.tq[0];
  const int D = 1 << k;
  const int n_outer = 1 << (T - k);
  const uint32_t lc_mask = op.lc_mask, lc_val = op.lc_val;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  for (int i = tid; i < (1 << T); i += THREADS) out[i] = in[i];
  __syncthreads();
  for (int ob = 0; ob < n_outer; ob += 4) {
    uint32_t base[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t o = ob + j;
      // deposit the outer counter around the k target bits
      base[j] = ((o >> lo) << (lo + k)) | (o & ((1u << lo) - 1u));
      ok[j] = (o < (uint32_t)n_outer) && ((base[j] & lc_mask) == lc_val);
    }
    if (!(ok[0] || ok[1] || ok[2] || ok[3])) continue;
    for (int r = warp; r < D; r += NW) {
      V acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j].x = acc[j].y = 0;
      for (int c = lane; c < D; c += 32) {
        const V u = __ldg(&U[(size_t)r * D + c]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ok[j]) cfma(acc[j], u, in[base[j] | ((uint32_t)c << lo)]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
          acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, s);
          acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, s);
        }
        if (lane == 0 && ok[j]) out[base[j] | ((uint32_t)r << lo)] = acc[j];
      }
    }
  }
  __syncthreads();
}

template <typename R, int THREADS>
__global__ void __launch_bounds__(THREADS)
    qls_batch_expect_z_kernel(const QlsOp* __restrict__ ops, int n_ops, const unsigned char* __restrict__ pool, int n,
                              int z_qubit, const double* __restrict__ params, int n_params, int two_buffers,
                              double* __restrict__ out) {
  using V = typename Vec2<R>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* cur = reinterpret_cast<V*>(smem_raw);
  V* alt = cur + ((size_t)1 << n);
  const int tid = threadIdx.x;
  const int dim = 1 << n;
  const double* theta = params + (size_t)blockIdx.x * n_params;

  for (int i = tid; i < dim; i += THREADS) {
    V z;
    z.x = (i == 0) ? R(1) : R(0);
    z.y = 0;
    cur[i] = z;
  }
  __syncthreads();

  int o = 0;
  while (o < n_ops) {
    const QlsOp& op = ops[o];
    if (op.kind == QLS_OP_DENSE && (op.flags & QLS_OP_FLAG_PARAM)) {
      const R angle = (R)((double)op.param_coeff * theta[op.param_index] + (double)op.param_offset);
      op_param<V, R, THREADS>(cur, n, op, angle, tid);
      __syncthreads();
      ++o;
    } else if (op.kind == QLS_OP_MATVEC) {
      if (two_buffers) {
        op_matvec<V, THREADS>(cur, alt, n, op, reinterpret_cast<const V*>(pool + op.pool_off), tid);
        V* t = cur;
        cur = alt;
        alt = t;
      }
      ++o;
    } else {  // maximal run of streaming-kernel ops
      int e = o + 1;
      while (e < n_ops && !(ops[e].kind == QLS_OP_MATVEC) &&
             !(ops[e].kind == QLS_OP_DENSE && (ops[e].flags & QLS_OP_FLAG_PARAM)))
        ++e;
      apply_ops<V, THREADS, true>(cur, n, ops + o, e - o, pool, 0ull, tid);
      o = e;
    }
  }

  // <Z_q> = sum_i (-1)^{bit_q(i)} |psi_i|^2
  double s = 0.0;
  for (int i = tid; i < dim; i += THREADS) {
    const V v = cur[i];
    const double p = (double)v.x * v.x + (double)v.y * v.y;
    s += ((i >> z_qubit) & 1) ? -p : p;
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  __shared__ double wsum[THREADS / 32];
  if ((tid & 31) == 0) wsum[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) t += wsum[w];
    out[blockIdx.x] = t;
  }
}

// Scratch layout of a state stored for a MATVEC on bits [lo, lo + k): those bits become bits [0, k), bits [0, lo) follow,
// higher bits stay.
// ops [o_begin, o_end) of the pass (no MATVEC among them) on one circuit per CTA.  load_k > 0: the state comes from the
// scratch array (stored for the MATVEC on [load_lo, load_lo + load_k)), else |0...0>; store_k > 0: it goes back to the
// scratch array for the next MATVEC, else <Z_q> is reduced into out[circuit].
// (HEAVY = the pass holds a dense block on 4 or 5 qubits: their register-blocked code costs 252 registers per thread, i.e.
// ONE resident CTA per SM -- an 8 x slower ansatz segment when it is compiled in without being needed)
template <int THREADS, bool HEAVY>
__global__ void __launch_bounds__(THREADS, HEAVY ? 1 : 768 / THREADS)
    qls_batch_segment_kernel(const QlsOp* __restrict__ ops, int o_begin, int o_end, const unsigned char* __restrict__ pool, int n,
                             int z_qubit, const double* __restrict__ params, int n_params, double2* __restrict__ scratch,
                             int load_lo, int load_k, int store_lo, int store_k, double* __restrict__ out) {
  using V = double2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* cur = reinterpret_cast<V*>(smem_raw);
  const int tid = threadIdx.x;
  const int dim = 1 << n;
  const double* theta = params + (size_t)blockIdx.x * n_params;
  V* mine = scratch + ((size_t)blockIdx.x << n);
  if (load_k > 0) {
    // (consecutive threads read consecutive scratch entries; the permuted shared-memory writes are conflict-light)
    for (int j = tid; j < dim; j += THREADS) {
      const uint32_t tgt = (uint32_t)j & ((1u << load_k) - 1u);
      const uint32_t low = ((uint32_t)j >> load_k) & ((1u << load_lo) - 1u);
      const uint32_t i = (tgt << load_lo) | low | (((uint32_t)j >> (load_lo + load_k)) << (load_lo + load_k));
      cur[i] = mine[j];
    }
  } else {
    for (int i = tid; i < dim; i += THREADS) cur[i] = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
  }
  // cos / sin of every parameterised rotation of the segment, once per circuit (one thread per rotation) instead of once
  // per thread and rotation: double-precision sincos was a third of the instructions of an ansatz segment
  double2* cs = reinterpret_cast<double2*>(cur + dim);
  for (int j = o_begin + tid; j < o_end; j += THREADS) {
    const QlsOp& op = ops[j];
    if (op.kind == QLS_OP_DENSE && (op.flags & QLS_OP_FLAG_PARAM)) {
      const double angle = (double)op.param_coeff * theta[op.param_index] + (double)op.param_offset;
      double sn, cn;
      sincos(op.pad0 == PARAM_P ? angle : angle * 0.5, &sn, &cn);
      cs[j - o_begin] = make_double2(cn, sn);
    }
  }
  __syncthreads();
  int o = o_begin;
  while (o < o_end) {
    const QlsOp& op = ops[o];
    if (op.kind == QLS_OP_DENSE && (op.flags & QLS_OP_FLAG_PARAM)) {
      const double2 v = cs[o - o_begin];
      op_param_cs<V, double, THREADS>(cur, n, op, v.x, v.y, tid);
      __syncthreads();
      ++o;
    } else {
      int e = o + 1;
      while (e < o_end && !(ops[e].kind == QLS_OP_DENSE && (ops[e].flags & QLS_OP_FLAG_PARAM))) ++e;
      apply_ops<V, THREADS, HEAVY>(cur, n, ops + o, e - o, pool, 0ull, tid);
      o = e;
    }
  }
  if (store_k > 0) {
    for (int j = tid; j < dim; j += THREADS) {
      const uint32_t tgt = (uint32_t)j & ((1u << store_k) - 1u);
      const uint32_t low = ((uint32_t)j >> store_k) & ((1u << store_lo) - 1u);
      const uint32_t i = (tgt << store_lo) | low | (((uint32_t)j >> (store_lo + store_k)) << (store_lo + store_k));
      mine[j] = cur[i];
    }
    return;
  }
  double s = 0.0;
  for (int i = tid; i < dim; i += THREADS) {
    const V v = cur[i];
    const double p = v.x * v.x + v.y * v.y;
    s += ((i >> z_qubit) & 1) ? -p : p;
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  __shared__ double wsum[THREADS / 32];
  if ((tid & 31) == 0) wsum[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) t += wsum[w];
    out[blockIdx.x] = t;
  }
}

// true if every MATVEC of the pass can go through the tensor-pipe GEMM
static bool batch_gemm_form(const QlsProgram* prog, const QlsPass& p, int n_qubits) {
  if (prog->dtype != QLS_C128) return false;
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("QLS_BATCH_GEMM");
    off = (e && e[0] == '0') ? 1 : 0;
  }
  if (off) return false;
  for (uint32_t o = p.op_begin; o < p.op_begin + p.op_count; ++o) {
    const QlsOp& op = prog->ops_host[o];
    if (op.kind != QLS_OP_MATVEC) continue;
    if (op.k < 6 || op.k > 10 || op.tq[0] + op.k > n_qubits || op.xc_mask) return false;
    if (op.lc_mask & (((1u << op.k) - 1u) << op.tq[0])) return false;
  }
  return true;
}

static int batch_expect_z_gemm(const QlsProgram* prog, const QlsPass& p, int n, int z_qubit, const double* params_dev,
                               uint32_t n_circuits, uint32_t n_params, double* out_dev, cudaStream_t s) {
  // 128 threads per circuit: the per-op overhead (record decode, dispatch, block barrier) is paid per warp, the arithmetic
  // of a 10-qubit state is a few instructions per thread either way
  constexpr int THREADS = 128;
  auto kern = (p.pad & 1u) ? qls_batch_segment_kernel<THREADS, true> : qls_batch_segment_kernel<THREADS, false>;
  const size_t smem = ((size_t)16 << n) + (size_t)16 * p.op_count;  // state + (cos, sin) of the parameterised rotations
  QLS_ARG(smem <= 200 * 1024, "state does not fit shared memory (%zu bytes)", smem);
  QLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  // circuits per chunk: the scratch array stays within 512 MiB
  uint32_t chunk = (uint32_t)(((size_t)512 << 20) / ((size_t)16 << n));
  if (chunk > n_circuits) chunk = n_circuits;
  double2* scratch = nullptr;
  if (p.pad & 2u) {
    const uint64_t need = ((uint64_t)16 << n) * chunk;
    QlsProgram* mp = const_cast<QlsProgram*>(prog);  // the scratch array is a cache of the program object
    if (mp->batch_scratch_bytes < need) {
      if (mp->batch_scratch) QLS_CUDA(cudaFree(mp->batch_scratch));  // (synchronises: earlier launches are done with it)
      mp->batch_scratch = nullptr;
      mp->batch_scratch_bytes = 0;
      QLS_CUDA(cudaMalloc(&mp->batch_scratch, need));
      mp->batch_scratch_bytes = need;
    }
    scratch = reinterpret_cast<double2*>(mp->batch_scratch);
  }
  int rc = QLS_OK;
  for (uint32_t c0 = 0; c0 < n_circuits && rc == QLS_OK; c0 += chunk) {
    const uint32_t nc = n_circuits - c0 < chunk ? n_circuits - c0 : chunk;
    const double* th = params_dev ? params_dev + (size_t)c0 * n_params : nullptr;
    int seg_begin = (int)p.op_begin, load_lo = 0, load_k = 0;
    const int end = (int)(p.op_begin + p.op_count);
    for (int o = seg_begin; o <= end && rc == QLS_OK; ++o) {
      const bool cut = o < end && prog->ops_host[o].kind == QLS_OP_MATVEC;
      if (!cut && o < end) continue;
      const int store_lo = cut ? (int)prog->ops_host[o].tq[0] : 0, store_k = cut ? (int)prog->ops_host[o].k : 0;
      kern<<<nc, THREADS, smem, s>>>(prog->ops_dev, seg_begin, o, prog->pool_dev, n, z_qubit, th, (int)n_params, scratch, load_lo,
                                     load_k, store_lo, store_k, out_dev + c0);
      if (cut) {
        const QlsOp& op = prog->ops_host[o];
        // controls of the MATVEC in the permuted row index: bits below the targets keep their place, bits above move down by k
        uint64_t rmask = 0, rval = 0;
        for (int b = 0; b < n; ++b)
          if ((op.lc_mask >> b) & 1u) {
            const int rb = b < (int)op.tq[0] ? b : b - (int)op.k;
            rmask |= 1ull << rb;
            rval |= (uint64_t)((op.lc_val >> b) & 1u) << rb;
          }
        rc = qls_gemm_rows_inplace(scratch, (uint64_t)nc << (n - op.k), op.k, rmask, rval, prog->pool_dev + op.pool_off, s);
        load_lo = store_lo;
        load_k = store_k;
        seg_begin = o + 1;
      }
    }
  }
  if (rc != QLS_OK) return rc;
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}

extern "C" int qls_batch_expect_z(const QlsProgram* prog, uint32_t pass_index, int n_qubits, int z_qubit,
                                  const double* params_dev, uint32_t n_circuits, uint32_t n_params, double* out_dev,
                                  void* stream) {
  QLS_ARG(prog && out_dev, "null program / output");
  QLS_ARG(pass_index < prog->n_passes, "pass index %u outside program", pass_index);
  QLS_ARG(n_params == 0 || params_dev, "null parameter array");
  const QlsPass& p = prog->passes_host[pass_index];
  QLS_ARG(p.tile_bits == n_qubits && p.n_high == 0 && p.fix_mask == 0, "batch programs need one whole-state pass");
  QLS_ARG(z_qubit >= 0 && z_qubit < n_qubits, "observable qubit outside the circuit");
  const int max_n = prog->dtype == QLS_C128 ? 12 : 13;
  QLS_ARG(n_qubits >= 1 && n_qubits <= max_n, "batch kernel supports 1..%d qubits, got %d", max_n, n_qubits);
  if (n_circuits == 0) return QLS_OK;
  QLS_CUDA(cudaSetDevice(prog->device));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (batch_gemm_form(prog, p, n_qubits))
    return batch_expect_z_gemm(prog, p, n_qubits, z_qubit, params_dev, n_circuits, n_params, out_dev, s);
  const int two = (p.pad & 2u) ? 1 : 0;  // set by qls_program_create when the pass holds a MATVEC
  const size_t esz = prog->dtype == QLS_C128 ? 16 : 8;
  const size_t smem = (esz << n_qubits) * (two ? 2 : 1);
  QLS_ARG(smem <= 200 * 1024, "state does not fit shared memory (%zu bytes)", smem);
  constexpr int THREADS = 256;
  if (prog->dtype == QLS_C128) {
    auto kern = qls_batch_expect_z_kernel<double, THREADS>;
    QLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kern<<<n_circuits, THREADS, smem, s>>>(prog->ops_dev + p.op_begin, (int)p.op_count, prog->pool_dev, n_qubits, z_qubit,
                                           params_dev, (int)n_params, two, out_dev);
  } else {
    auto kern = qls_batch_expect_z_kernel<float, THREADS>;
    QLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kern<<<n_circuits, THREADS, smem, s>>>(prog->ops_dev + p.op_begin, (int)p.op_count, prog->pool_dev, n_qubits, z_qubit,
                                           params_dev, (int)n_params, two, out_dev);
  }
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}
