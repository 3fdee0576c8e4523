// Register-stage tile-pass kernel (the fast path for full-size tiles).
//
// Persistent CTAs (2 per SM, 256 threads).  A tile of 2^T amplitudes (T = 12 complex128 /
// 13 complex64 -> 64 KB) is staged into shared memory with 16-byte cp.async, then lives in the
// registers of the CTA: every thread holds 2^KR amplitudes (KR = 4 / 5: the stage's "register
// qubits"), the remaining 8 tile bits are the thread index.  (512 threads x 2^(KR-1) amplitudes was
// measured slower: more stages per pass, see profiles/.)  All micro-ops of a stage -- 1- and
// 2-qubit dense blocks on register qubits, diagonal tables, multiplexed RY -- run on registers with
// compile-time register indices; shared memory is touched only to load, to re-distribute the tile
// between stages (one write + one read), and to write back.  The shared-memory image is XOR
// swizzled so that both the coalesced global<->shared copies and the strided per-stage accesses
// spread over the banks.  Address arithmetic that does not depend on the tile (thread -> index
// deposits, high-qubit spreads, op records) is computed once per CTA, not once per amplitude.
#include "qls_common.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_MAX_OPS = 128;
constexpr int RS_MAX_STAGES = 24;

struct RsLaunch {
  int L, n_high, n_fseg, n_rs;
  uint8_t high[QLS_MAX_HIGH];
  QlsSeg fseg[QLS_MAX_FSEG];
  uint64_t fix_val;
  uint64_t index_offset;
  uint64_t n_tiles;
};

// swizzle of a tile-local amplitude index; SH = log2(amplitudes per 16-byte unit)
template <int SH>
__device__ __forceinline__ uint32_t swz(uint32_t i) {
  const uint32_t u = i >> SH;
  return i ^ ((((u >> 3) ^ (u >> 6) ^ (u >> 9)) & 7u) << SH);
}

// address of register combination r: base ^ (xor of the per-qubit offsets of r's set bits); with a
// compile-time r this folds into at most KR XORs on registers (no table lookups)
template <int KR>
__device__ __forceinline__ uint32_t rs_addr(uint32_t base, const uint32_t (&c)[KR], int r) {
  uint32_t a = base;
#pragma unroll
  for (int j = 0; j < KR; ++j)
    if ((r >> j) & 1) a ^= c[j];
  return a;
}

constexpr uint32_t QLS_RS_FLAG_DIRECT = 0x10u;  // STAGE: every register qubit sits above the lane bits

// CTL = false: the block has no control inside the tile -> straight-line code over all groups, so
// the scheduler can interleave their independent FMA chains (a per-group predicate turns every
// group into its own basic block and serialises them).
template <typename V, int A, int RQ, bool REAL, bool CTL>
__device__ __forceinline__ void rs_g1(V (&a)[A], const V* __restrict__ M, bool tok, uint32_t rmask, uint32_t rval) {
  const V m00 = __ldg(M), m01 = __ldg(M + 1), m10 = __ldg(M + 2), m11 = __ldg(M + 3);
#pragma unroll
  for (int p = 0; p < A / 2; ++p) {
    const int i0 = ((p >> RQ) << (RQ + 1)) | (p & ((1 << RQ) - 1));
    const int i1 = i0 | (1 << RQ);
    if (!CTL || (tok && (((uint32_t)i0 & rmask) == rval))) {
      const V x = a[i0], y = a[i1];
      V b0, b1;
      if (REAL) {
        b0.x = fma(m01.x, y.x, m00.x * x.x);
        b0.y = fma(m01.x, y.y, m00.x * x.y);
        b1.x = fma(m11.x, y.x, m10.x * x.x);
        b1.y = fma(m11.x, y.y, m10.x * x.y);
      } else {
        b0 = cmul(m00, x);
        cfma(b0, m01, y);
        b1 = cmul(m10, x);
        cfma(b1, m11, y);
      }
      a[i0] = b0;
      a[i1] = b1;
    }
  }
}

template <typename V, int A, int R1, int R2, bool REAL, bool CTL>
__device__ __forceinline__ void rs_g2(V (&a)[A], const V* __restrict__ M, bool tok, uint32_t rmask, uint32_t rval) {
  static_assert(R1 < R2, "register qubits must be ascending");
#pragma unroll
  for (int g = 0; g < A / 4; ++g) {
    const int t0 = ((g >> R1) << (R1 + 1)) | (g & ((1 << R1) - 1));               // zero at R1
    const int i00 = ((t0 >> R2) << (R2 + 1)) | (t0 & ((1 << R2) - 1));            // zero at R2
    const int i01 = i00 | (1 << R1), i10 = i00 | (1 << R2), i11 = i01 | (1 << R2);
    if (!CTL || (tok && (((uint32_t)i00 & rmask) == rval))) {
      const V x0 = a[i00], x1 = a[i01], x2 = a[i10], x3 = a[i11];
      V o[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const V m0 = __ldg(M + 4 * r), m1 = __ldg(M + 4 * r + 1), m2 = __ldg(M + 4 * r + 2), m3 = __ldg(M + 4 * r + 3);
        if (REAL) {  // H (x) H, CX, RY (x) RY ...: half the multiplies
          o[r].x = fma(m3.x, x3.x, fma(m2.x, x2.x, fma(m1.x, x1.x, m0.x * x0.x)));
          o[r].y = fma(m3.x, x3.y, fma(m2.x, x2.y, fma(m1.x, x1.y, m0.x * x0.y)));
        } else {
          o[r] = cmul(m0, x0);
          cfma(o[r], m1, x1);
          cfma(o[r], m2, x2);
          cfma(o[r], m3, x3);
        }
      }
      a[i00] = o[0];
      a[i01] = o[1];
      a[i10] = o[2];
      a[i11] = o[3];
    }
  }
}

// table-index contribution of register combination r = OR of the per-register-qubit contributions
// (op.rt[1 << j]); with a compile-time r this folds into at most KR ORs on registers
template <int KR>
__device__ __forceinline__ uint32_t rs_tidx(uint32_t base, const uint32_t (&d)[KR], int r) {
  uint32_t t = base;
#pragma unroll
  for (int j = 0; j < KR; ++j)
    if ((r >> j) & 1) t |= d[j];
  return t;
}

template <typename V, int A, int KR>
__device__ __forceinline__ void rs_diag(V (&a)[A], const QlsRsOp& op, const V* __restrict__ table, uint32_t ti0, bool tok) {
  uint32_t d[KR];
#pragma unroll
  for (int j = 0; j < KR; ++j) d[j] = op.rt[1 << j];
  const uint32_t rm = op.ctl_rmask, rv = op.ctl_rval;
#pragma unroll
  for (int r = 0; r < A; ++r) {
    if (tok && (((uint32_t)r & rm) == rv)) a[r] = cmul(a[r], __ldg(&table[rs_tidx<KR>(ti0, d, r)]));
  }
}

template <typename V, int A, int KR, int RQ>
__device__ __forceinline__ void rs_muxry(V (&a)[A], const QlsRsOp& op, const V* __restrict__ table, uint32_t ti0, bool tok) {
  uint32_t d[KR];
#pragma unroll
  for (int j = 0; j < KR; ++j) d[j] = op.rt[1 << j];
  const uint32_t rm = op.ctl_rmask, rv = op.ctl_rval;
#pragma unroll
  for (int p = 0; p < A / 2; ++p) {
    const int i0 = ((p >> RQ) << (RQ + 1)) | (p & ((1 << RQ) - 1));
    const int i1 = i0 | (1 << RQ);
    if (tok && (((uint32_t)i0 & rm) == rv)) {
      const V cs = __ldg(&table[rs_tidx<KR>(ti0, d, i0)]);
      const V x = a[i0], y = a[i1];
      V b0, b1;
      b0.x = cs.x * x.x - cs.y * y.x;
      b0.y = cs.x * x.y - cs.y * y.y;
      b1.x = cs.y * x.x + cs.x * y.x;
      b1.y = cs.y * x.y + cs.x * y.y;
      a[i0] = b0;
      a[i1] = b1;
    }
  }
}

template <typename V, int A, int KR>
__device__ __forceinline__ void rs_dispatch_g1(V (&a)[A], const QlsRsOp& op, const V* M, bool tok) {
  const bool real = (op.flags & QLS_OP_FLAG_REAL) != 0;
  const uint32_t rm = op.ctl_rmask, rv = op.ctl_rval;
  const bool ctl = (op.ctl_tmask | op.ctl_rmask) != 0;
#define QLS_G1_CASE(q)                                                   \
  case q:                                                                \
    if (ctl) {                                                           \
      if (real) rs_g1<V, A, q, true, true>(a, M, tok, rm, rv);           \
      else rs_g1<V, A, q, false, true>(a, M, tok, rm, rv);               \
    } else {                                                             \
      if (real) rs_g1<V, A, q, true, false>(a, M, tok, rm, rv);          \
      else rs_g1<V, A, q, false, false>(a, M, tok, rm, rv);              \
    }                                                                    \
    break;
  switch (op.r[0]) {
    QLS_G1_CASE(0)
    QLS_G1_CASE(1)
    QLS_G1_CASE(2)
    QLS_G1_CASE(3)
    case 4:
      if constexpr (KR >= 5) {
        if (real) rs_g1<V, A, 4, true, true>(a, M, tok, rm, rv);
        else rs_g1<V, A, 4, false, true>(a, M, tok, rm, rv);
      }
      break;
  }
#undef QLS_G1_CASE
}

template <typename V, int A, int KR>
__device__ __forceinline__ void rs_dispatch_g2(V (&a)[A], const QlsRsOp& op, const V* M, bool tok) {
  const uint32_t rm = op.ctl_rmask, rv = op.ctl_rval;
  const bool real = (op.flags & QLS_OP_FLAG_REAL) != 0;
  // complex 4x4: keep one basic block per group (the straight-line form interleaves all 64 FMA chains,
  // runs out of registers and spills -- measured slower); real 4x4 and all 2x2 blocks go straight-line
  const bool ctl = (op.ctl_tmask | op.ctl_rmask) != 0 || !real;
#define QLS_G2_CASE(p, q)                                                  \
  case (p * 8 + q):                                                        \
    if (ctl) {                                                             \
      if (real) rs_g2<V, A, p, q, true, true>(a, M, tok, rm, rv);          \
      else rs_g2<V, A, p, q, false, true>(a, M, tok, rm, rv);              \
    } else {                                                               \
      rs_g2<V, A, p, q, true, false>(a, M, tok, rm, rv);                   \
    }                                                                      \
    break;
  switch (op.r[0] * 8 + op.r[1]) {
    QLS_G2_CASE(0, 1)
    QLS_G2_CASE(0, 2)
    QLS_G2_CASE(0, 3)
    QLS_G2_CASE(1, 2)
    QLS_G2_CASE(1, 3)
    QLS_G2_CASE(2, 3)
    default:
      if constexpr (KR >= 5) {
        switch (op.r[0]) {
          case 0: if (real) rs_g2<V, A, 0, 4, true, true>(a, M, tok, rm, rv); else rs_g2<V, A, 0, 4, false, true>(a, M, tok, rm, rv); break;
          case 1: if (real) rs_g2<V, A, 1, 4, true, true>(a, M, tok, rm, rv); else rs_g2<V, A, 1, 4, false, true>(a, M, tok, rm, rv); break;
          case 2: if (real) rs_g2<V, A, 2, 4, true, true>(a, M, tok, rm, rv); else rs_g2<V, A, 2, 4, false, true>(a, M, tok, rm, rv); break;
          case 3: if (real) rs_g2<V, A, 3, 4, true, true>(a, M, tok, rm, rv); else rs_g2<V, A, 3, 4, false, true>(a, M, tok, rm, rv); break;
        }
      }
      break;
  }
#undef QLS_G2_CASE
}

template <typename V, int A, int KR>
__device__ __forceinline__ void rs_dispatch_muxry(V (&a)[A], const QlsRsOp& op, const V* table, uint32_t ti0, bool tok) {
  switch (op.r[0]) {
    case 0: rs_muxry<V, A, KR, 0>(a, op, table, ti0, tok); break;
    case 1: rs_muxry<V, A, KR, 1>(a, op, table, ti0, tok); break;
    case 2: rs_muxry<V, A, KR, 2>(a, op, table, ti0, tok); break;
    case 3: rs_muxry<V, A, KR, 3>(a, op, table, ti0, tok); break;
    case 4:
      if constexpr (KR >= 5) rs_muxry<V, A, KR, 4>(a, op, table, ti0, tok);
      break;
  }
}

template <typename R, int T, int KR>
__global__ void __launch_bounds__(RS_THREADS, 2)
    qls_rs_pass_kernel(typename Vec2<R>::type* __restrict__ state, const RsLaunch rl, const QlsRsOp* __restrict__ rs_ops,
                       const unsigned char* __restrict__ pool) {
  using V = typename Vec2<R>::type;
  static_assert((1 << (T - KR)) == RS_THREADS, "the threads index the non-register tile bits");
  constexpr int A = 1 << KR;
  constexpr int SH = sizeof(V) == 16 ? 0 : 1;
  constexpr int APU = 1 << SH;           // amplitudes per 16-byte unit
  constexpr int UNITS = (1 << T) / APU;  // 4096
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* sm = reinterpret_cast<V*>(smem_raw);
  uint4* sm16 = reinterpret_cast<uint4*>(smem_raw);
  QlsRsOp* sops = reinterpret_cast<QlsRsOp*>(smem_raw + (sizeof(V) << T));
  uint64_t* hoff = reinterpret_cast<uint64_t*>(sops + rl.n_rs);            // [1 << n_high] (+1: tile base slot)
  uint64_t* s_base = hoff + (1 << rl.n_high);
  uint16_t* ptb = reinterpret_cast<uint16_t*>(s_base + 2);                 // [stages][RS_THREADS]

  const int tid = threadIdx.x;
  const int L = rl.L;
  const int n_rs = rl.n_rs;

  // ---- once per CTA ----------------------------------------------------------------------------
  // The descriptor arrays of the kernel parameters are copied to shared memory with COMPILE-TIME
  // indices: a dynamically indexed parameter array makes the compiler spill the whole parameter
  // struct to local memory, and then even the op-loop bound is a local load on the critical path.
  __shared__ QlsSeg s_fseg[QLS_MAX_FSEG];
  __shared__ uint8_t s_high[QLS_MAX_HIGH];
  if (tid == 0) {
#pragma unroll
    for (int j = 0; j < QLS_MAX_FSEG; ++j) s_fseg[j] = rl.fseg[j];
#pragma unroll
    for (int j = 0; j < QLS_MAX_HIGH; ++j) s_high[j] = rl.high[j];
  }
  for (int i = tid; i < n_rs * (int)(sizeof(QlsRsOp) / 16); i += RS_THREADS)
    reinterpret_cast<uint4*>(sops)[i] = reinterpret_cast<const uint4*>(rs_ops)[i];
  __syncthreads();
  for (int r = tid; r < (1 << rl.n_high); r += RS_THREADS) {
    uint64_t g = 0;
    for (int j = 0; j < rl.n_high; ++j) g |= (uint64_t)((r >> j) & 1) << s_high[j];
    hoff[r] = g;
  }
  __syncthreads();
  {
    int st = 0;
    for (int o = 0; o < n_rs; ++o) {
      if (sops[o].kind == QLS_RS_STAGE) {
        const uint32_t tb = gather_segs32((uint32_t)tid, sops[o].tseg, sops[o].n_tseg);
        ptb[st * RS_THREADS + tid] = (uint16_t)swz<SH>(tb);
        ++st;
      }
    }
  }
  __syncthreads();

  bool first_direct = false, last_direct = false;
  if (n_rs > 0) first_direct = (sops[0].flags & QLS_RS_FLAG_DIRECT) != 0;
  for (int o = 0; o < n_rs; ++o)
    if (sops[o].kind == QLS_RS_STAGE) last_direct = (sops[o].flags & QLS_RS_FLAG_DIRECT) != 0;

  // ---- persistent loop over tiles ------------------------------------------------------------------
  // The tile counter -> base index deposit walks byte-sized segment descriptors in the kernel
  // parameters; one thread does it per tile and publishes the result (every thread doing it cost
  // more L1 traffic than the tile itself).
  if (tid == 0) s_base[0] = gather_segs64((uint64_t)blockIdx.x, s_fseg, rl.n_fseg) | rl.fix_val;
  __syncthreads();
  int parity = 0;
  for (uint64_t tile = blockIdx.x; tile < rl.n_tiles; tile += gridDim.x) {
    const uint64_t base = s_base[parity];
    const uint64_t gbase = base | rl.index_offset;
    if (tid == 0 && tile + gridDim.x < rl.n_tiles)
      s_base[parity ^ 1] = gather_segs64(tile + gridDim.x, s_fseg, rl.n_fseg) | rl.fix_val;  // visible after the syncs below
    parity ^= 1;

    if (!first_direct) {
#pragma unroll 4
      for (int u = tid; u < UNITS; u += RS_THREADS) {
        const int i = u * APU;
        const uint64_t g = base | (uint64_t)(i & ((1 << L) - 1)) | hoff[i >> L];
        cp_async16(&sm16[swz<SH>((uint32_t)i) >> SH], &state[g]);
      }
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
    }

    V a[A];
    int stage = -1;
    uint32_t mytb = 0;
    uint32_t cq[KR];  // swizzled shared-memory offset of each register qubit of the current stage
    uint32_t lq[KR];  // its tile-local bit
    for (int o = 0; o < n_rs; ++o) {
      const QlsRsOp& op = sops[o];
      if (op.kind == QLS_RS_STAGE) {
        if (stage >= 0) {
#pragma unroll
          for (int r = 0; r < A; ++r) sm[rs_addr<KR>(mytb, cq, r)] = a[r];
          __syncthreads();
        }
        ++stage;
        mytb = ptb[stage * RS_THREADS + tid];
#pragma unroll
        for (int j = 0; j < KR; ++j) {
          lq[j] = 1u << (j < 4 ? op.r[j] : op.pad0);
          cq[j] = swz<SH>(lq[j]);
        }
        if (stage == 0 && first_direct) {
          // all register qubits sit above the 5 lane bits: every warp load is 512 contiguous bytes
          const uint32_t ltb = swz<SH>(mytb);  // the swizzle is an involution
#pragma unroll
          for (int r = 0; r < A; ++r) {
            const uint32_t idx = rs_addr<KR>(ltb, lq, r);
            a[r] = state[base | (uint64_t)(idx & ((1u << L) - 1u)) | hoff[idx >> L]];
          }
        } else {
#pragma unroll
          for (int r = 0; r < A; ++r) a[r] = sm[rs_addr<KR>(mytb, cq, r)];
          __syncthreads();
        }
        continue;
      }
      if ((gbase & op.xc_mask) != op.xc_val) continue;  // block-uniform
      const bool tok = ((uint32_t)tid & op.ctl_tmask) == op.ctl_tval;
      const V* data = reinterpret_cast<const V*>(pool + op.pool_off);
      switch (op.kind) {
        case QLS_RS_G1:
          rs_dispatch_g1<V, A, KR>(a, op, data, tok);
          break;
        case QLS_RS_G2:
          rs_dispatch_g2<V, A, KR>(a, op, data, tok);
          break;
        case QLS_RS_DIAG: {
          const uint32_t ti0 = (uint32_t)gather_segs64(gbase, op.xseg, op.n_xseg) | gather_segs32((uint32_t)tid, op.tseg, op.n_tseg);
          rs_diag<V, A, KR>(a, op, data, ti0, tok);
          break;
        }
        case QLS_RS_MUXRY: {
          const uint32_t ti0 = (uint32_t)gather_segs64(gbase, op.xseg, op.n_xseg) | gather_segs32((uint32_t)tid, op.tseg, op.n_tseg);
          rs_dispatch_muxry<V, A, KR>(a, op, data, ti0, tok);
          break;
        }
      }
    }
    if (last_direct) {
      const uint32_t ltb = swz<SH>(mytb);
#pragma unroll
      for (int r = 0; r < A; ++r) {
        const uint32_t idx = rs_addr<KR>(ltb, lq, r);
        state[base | (uint64_t)(idx & ((1u << L) - 1u)) | hoff[idx >> L]] = a[r];
      }
    } else {
#pragma unroll
      for (int r = 0; r < A; ++r) sm[rs_addr<KR>(mytb, cq, r)] = a[r];
      __syncthreads();
#pragma unroll 4
      for (int u = tid; u < UNITS; u += RS_THREADS) {
        const int i = u * APU;
        const uint64_t g = base | (uint64_t)(i & ((1 << L) - 1)) | hoff[i >> L];
        *reinterpret_cast<uint4*>(&state[g]) = sm16[swz<SH>((uint32_t)i) >> SH];
      }
    }
    __syncthreads();  // smem image and s_base slot are free for the next tile
  }
}

template <typename R, int T, int KR>
int launch_rs(const QlsProgram* prog, const QlsPass& p, void* state, int n_local, uint64_t index_offset, cudaStream_t stream) {
  using V = typename Vec2<R>::type;
  RsLaunch rl;
  memset(&rl, 0, sizeof(rl));
  rl.L = p.low_bits;
  rl.n_high = p.n_high;
  rl.n_fseg = p.n_fseg;
  rl.n_rs = (int)p.rs_count;
  memcpy(rl.high, p.high, sizeof(rl.high));
  memcpy(rl.fseg, p.fseg, sizeof(rl.fseg));
  rl.fix_val = p.fix_val;
  rl.index_offset = index_offset;
  const int fixed = __builtin_popcountll(p.fix_mask);
  const int free_bits = n_local - p.tile_bits - fixed;
  QLS_ARG(free_bits >= 0 && free_bits < 40, "pass geometry: n_local=%d tile_bits=%d fixed=%d", n_local, p.tile_bits, fixed);
  rl.n_tiles = 1ull << free_bits;
  int n_stages = p.pad >> 8;  // counted by qls_program_create
  const size_t smem = ((size_t)sizeof(V) << T) + (size_t)p.rs_count * sizeof(QlsRsOp) + (sizeof(uint64_t) << p.n_high) +
                      2 * sizeof(uint64_t) + (size_t)n_stages * RS_THREADS * sizeof(uint16_t);
  QLS_ARG(smem <= 112 * 1024, "register-stage pass needs %zu bytes of shared memory", smem);
  auto kern = qls_rs_pass_kernel<R, T, KR>;
  static bool attr_set[64] = {false};
  static int sm_count[64] = {0};
  if (!attr_set[prog->device & 63]) {
    QLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    QLS_CUDA(cudaDeviceGetAttribute(&sm_count[prog->device & 63], cudaDevAttrMultiProcessorCount, prog->device));
    attr_set[prog->device & 63] = true;
  }
  uint64_t grid = 2ull * (uint64_t)sm_count[prog->device & 63];
  if (grid > rl.n_tiles) grid = rl.n_tiles;
  kern<<<(unsigned)grid, RS_THREADS, smem, stream>>>(reinterpret_cast<V*>(state), rl, prog->rs_ops_dev + p.rs_begin,
                                                     prog->pool_dev);
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}

}  // namespace

// true if pass `p` of `prog` can run in register-stage form on a state of n_local qubits
bool qls_rs_eligible(const QlsProgram* prog, const QlsPass& p, int n_local) {
  const int want_T = prog->dtype == QLS_C128 ? 12 : 13;
  return p.rs_count > 0 && p.rs_count <= (uint32_t)RS_MAX_OPS && p.tile_bits == want_T && n_local >= want_T &&
         (int)(p.pad >> 8) >= 1 && (int)(p.pad >> 8) <= RS_MAX_STAGES;
}

int qls_launch_rs_pass(const QlsProgram* prog, const QlsPass& p, void* state, int n_local, uint64_t index_offset,
                       cudaStream_t stream) {
  if (prog->dtype == QLS_C128) return launch_rs<double, 12, 4>(prog, p, state, n_local, index_offset, stream);
  return launch_rs<float, 13, 5>(prog, p, state, n_local, index_offset, stream);
}
