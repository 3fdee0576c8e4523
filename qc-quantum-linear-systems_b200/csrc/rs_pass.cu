// Register-stage tile-pass kernel (the fast path for full-size tiles).
//
// Persistent CTAs (2 per SM, 256 threads).  A tile of 2^T amplitudes (T = 12 complex128 /
// 13 complex64 -> 64 KB) lives in the REGISTERS of the CTA: every thread holds 2^KR amplitudes
// (KR = 4 / 5: the stage's "register qubits"), the remaining 8 tile bits are the thread index.  All
// micro-ops of a stage -- 1- and 2-qubit dense blocks on register qubits, diagonal tables,
// multiplexed RY -- run on registers with compile-time register indices; shared memory is touched
// only to re-distribute the tile between stages (one XOR-swizzled write + one read) and to make
// the global accesses coalesced when the stage's lanes are not the low tile bits.
//
// What keeps the FP64 pipe fed (ncu of the previous version: 27 % FP64-pipe activity, 60 % of the
// issue slots spent on address/decode work, long-scoreboard stalls on matrix loads):
//   * the whole pass description -- header, micro-op records AND gate matrices -- sits in
//     __constant__ memory (uploaded per pass, stream ordered).  Every thread reads it at a
//     warp-uniform address, so records decode on the uniform datapath and the DFMAs take their
//     matrix element straight from a uniform register: no per-thread load in the inner loops;
//   * stages are entered lazily: a stage whose ops are all switched off for this tile by controls
//     outside the tile costs nothing, and a tile on which no op fires is neither read nor written;
//   * the next tile of the CTA is prefetched into L2 (cp.async.bulk.prefetch) while the current
//     one is in registers, so the HBM read latency is off the critical path;
//   * diagonal tables / multiplexer angles are fetched in batches before they are used.
#include <string.h>

#include <vector>

#include "qls_common.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_MAX_OPS = 160;
constexpr int RS_MAX_STAGES = 24;
constexpr int RS_MAX_XC = 8;
constexpr int RS_IMG_BYTES = 60 * 1024;
constexpr uint32_t QLS_RS_FLAG_DIRECT = 0x10u;  // STAGE: every register qubit sits above the lane bits

struct RsHeader {  // 256 bytes
  int32_t L, n_high, n_fseg, n_ops;
  uint64_t fix_val;
  int32_t n_xc;  // distinct (mask, value) pairs of controls outside the tile; 0: some op always fires
  int32_t n_stages;
  uint64_t xc_mask[RS_MAX_XC], xc_val[RS_MAX_XC];
  uint8_t high[QLS_MAX_HIGH];
  QlsSeg fseg[QLS_MAX_FSEG];
  uint8_t pad[24];
};
static_assert(sizeof(RsHeader) == 256, "RsHeader layout");

struct RsImage {
  RsHeader h;
  unsigned char body[RS_IMG_BYTES - sizeof(RsHeader)];  // QlsRsOp[n_ops], then the gate matrices
};

__constant__ RsImage c_img;

// swizzle of a tile-local amplitude index; SH = log2(amplitudes per 16-byte unit)
template <int SH>
__device__ __forceinline__ uint32_t swz(uint32_t i) {
  const uint32_t u = i >> SH;
  return i ^ ((((u >> 3) ^ (u >> 6) ^ (u >> 9)) & 7u) << SH);
}

// address of register combination r: base ^ (xor of the per-qubit offsets of r's set bits); with a
// compile-time r this folds into at most KR XORs on registers
template <int KR>
__device__ __forceinline__ uint32_t rs_addr(uint32_t base, const uint32_t (&c)[KR], int r) {
  uint32_t a = base;
#pragma unroll
  for (int j = 0; j < KR; ++j)
    if ((r >> j) & 1) a ^= c[j];
  return a;
}

__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gptr), "r"(bytes) : "memory");
}

// ---- micro-ops on the register tile ------------------------------------------------------------------
// CTL = false: no control inside the tile -> straight-line code over all groups.  CTL = true: `tok`
// is the per-thread (thread-bit controls) and (i & rmask) == rval the per-group (register-bit
// controls, warp-uniform) predicate.  M points into __constant__ memory: its elements reach the
// DFMAs as uniform-register operands.

template <typename V, int A, int RQ, bool REAL, bool CTL>
__device__ __forceinline__ void rs_g1(V (&a)[A], const V* __restrict__ M, bool tok, uint32_t rmask, uint32_t rval) {
  const V m00 = M[0], m01 = M[1], m10 = M[2], m11 = M[3];
  if (CTL && !tok) return;
#pragma unroll
  for (int p = 0; p < A / 2; ++p) {
    const int i0 = ((p >> RQ) << (RQ + 1)) | (p & ((1 << RQ) - 1));
    const int i1 = i0 | (1 << RQ);
    if (!CTL || (((uint32_t)i0 & rmask) == rval)) {
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
  if (CTL && !tok) return;
#pragma unroll
  for (int g = 0; g < A / 4; ++g) {
    const int t0 = ((g >> R1) << (R1 + 1)) | (g & ((1 << R1) - 1));               // zero at R1
    const int i00 = ((t0 >> R2) << (R2 + 1)) | (t0 & ((1 << R2) - 1));            // zero at R2
    const int i01 = i00 | (1 << R1), i10 = i00 | (1 << R2), i11 = i01 | (1 << R2);
    if (!CTL || (((uint32_t)i00 & rmask) == rval)) {
      const V x0 = a[i00], x1 = a[i01], x2 = a[i10], x3 = a[i11];
      V o[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const V m0 = M[4 * r], m1 = M[4 * r + 1], m2 = M[4 * r + 2], m3 = M[4 * r + 3];
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
  uint32_t any = 0;
#pragma unroll
  for (int j = 0; j < KR; ++j) {
    d[j] = op.rt[1 << j];
    any |= d[j];
  }
  const uint32_t rm = op.ctl_rmask, rv = op.ctl_rval;
  if ((any | rm) == 0) {  // warp-uniform: the table does not look at a register qubit -> one phase per thread
    const V ph = __ldg(&table[ti0]);
    if (tok) {
#pragma unroll
      for (int r = 0; r < A; ++r) a[r] = cmul(a[r], ph);
    }
    return;
  }
  constexpr int CH = 4;  // phases fetched per batch (register budget)
#pragma unroll
  for (int c = 0; c < A; c += CH) {
    V ph[CH];
#pragma unroll
    for (int r = 0; r < CH; ++r) ph[r] = __ldg(&table[rs_tidx<KR>(ti0, d, c + r)]);  // index is valid whatever the controls say
#pragma unroll
    for (int r = 0; r < CH; ++r)
      if (tok && (((uint32_t)(c + r) & rm) == rv)) a[c + r] = cmul(a[c + r], ph[r]);
  }
}

template <typename V, int A, int KR, int RQ>
__device__ __forceinline__ void rs_muxry(V (&a)[A], const QlsRsOp& op, const V* __restrict__ table, uint32_t ti0, bool tok) {
  uint32_t d[KR];
#pragma unroll
  for (int j = 0; j < KR; ++j) d[j] = op.rt[1 << j];
  const uint32_t rm = op.ctl_rmask, rv = op.ctl_rval;
  constexpr int CH = 4;
#pragma unroll
  for (int c = 0; c < A / 2; c += CH) {
    V cs[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int p = c + q;
      const int i0 = ((p >> RQ) << (RQ + 1)) | (p & ((1 << RQ) - 1));
      cs[q] = __ldg(&table[rs_tidx<KR>(ti0, d, i0)]);
    }
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int p = c + q;
      const int i0 = ((p >> RQ) << (RQ + 1)) | (p & ((1 << RQ) - 1));
      const int i1 = i0 | (1 << RQ);
      if (tok && (((uint32_t)i0 & rm) == rv)) {
        const V x = a[i0], y = a[i1];
        V b0, b1;
        b0.x = cs[q].x * x.x - cs[q].y * y.x;
        b0.y = cs[q].x * x.y - cs[q].y * y.y;
        b1.x = cs[q].y * x.x + cs[q].x * y.x;
        b1.y = cs[q].y * x.y + cs[q].x * y.y;
        a[i0] = b0;
        a[i1] = b1;
      }
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
  const bool ctl = (op.ctl_tmask | op.ctl_rmask) != 0;
#define QLS_G2_CASE(p, q)                                                  \
  case (p * 8 + q):                                                        \
    if (ctl) {                                                             \
      if (real) rs_g2<V, A, p, q, true, true>(a, M, tok, rm, rv);          \
      else rs_g2<V, A, p, q, false, true>(a, M, tok, rm, rv);              \
    } else {                                                               \
      if (real) rs_g2<V, A, p, q, true, false>(a, M, tok, rm, rv);         \
      else rs_g2<V, A, p, q, false, false>(a, M, tok, rm, rv);             \
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

template <int KR, int SH>
__device__ __forceinline__ void rs_stage_bits(const QlsRsOp& sp, uint32_t (&cq)[KR], uint32_t (&lq)[KR]) {
#pragma unroll
  for (int j = 0; j < KR; ++j) {
    lq[j] = 1u << (j < 4 ? sp.r[j] : sp.pad0);
    cq[j] = swz<SH>(lq[j]);
  }
}

// A warp vote makes a condition PROVABLY warp-uniform for ptxas.  Every branch that guards a barrier
// goes through it: with a condition ptxas cannot prove uniform (loop-carried state such as the current
// stage) it gives up on the uniform datapath for the whole tile loop -- the op records and gate
// matrices then come through per-thread LDC + vector registers instead of LDCU + uniform-register
// DFMA operands, and the kernel spills.
__device__ __forceinline__ bool uni(bool c) { return __any_sync(0xffffffffu, c) != 0; }

// true if no op of the pass fires on the tile whose global base index is gbase (warp-uniform)
__device__ __forceinline__ bool rs_tile_idle(uint64_t gbase) {
  const int n = c_img.h.n_xc;
  if (n == 0) return false;
  bool hit = false;
  for (int j = 0; j < n; ++j) hit |= (gbase & c_img.h.xc_mask[j]) == c_img.h.xc_val[j];
  return !hit;
}

template <typename R, int T, int KR>
__global__ void __launch_bounds__(RS_THREADS, 2)
    qls_rs_pass_kernel(typename Vec2<R>::type* __restrict__ state, const uint64_t index_offset, const uint64_t n_tiles,
                       const unsigned char* __restrict__ pool) {
  using V = typename Vec2<R>::type;
  static_assert((1 << (T - KR)) == RS_THREADS, "the threads index the non-register tile bits");
  constexpr int A = 1 << KR;
  constexpr int SH = sizeof(V) == 16 ? 0 : 1;
  constexpr int APU = 1 << SH;           // amplitudes per 16-byte unit
  constexpr int UNITS = (1 << T) / APU;  // 4096
  constexpr uint32_t TILE_BYTES = (uint32_t)sizeof(V) << T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* sm = reinterpret_cast<V*>(smem_raw);
  uint4* sm16 = reinterpret_cast<uint4*>(smem_raw);
  uint64_t* hoff = reinterpret_cast<uint64_t*>(smem_raw + TILE_BYTES);  // [1 << n_high]
  uint16_t* ptb = reinterpret_cast<uint16_t*>(hoff + (1 << c_img.h.n_high));  // [stages][RS_THREADS]

  const QlsRsOp* __restrict__ ops = reinterpret_cast<const QlsRsOp*>(c_img.body);
  const int tid = threadIdx.x;
  const int L = c_img.h.L;
  const int n_ops = c_img.h.n_ops;
  const int n_high = c_img.h.n_high;
  const int n_fseg = c_img.h.n_fseg;

  // ---- once per CTA: high-qubit spread table, per-stage thread -> swizzled tile offset ----------------
  for (int r = tid; r < (1 << n_high); r += RS_THREADS) {
    uint64_t g = 0;
    for (int j = 0; j < n_high; ++j) g |= (uint64_t)((r >> j) & 1) << c_img.h.high[j];
    hoff[r] = g;
  }
  {
    int st = 0;
    for (int o = 0; o < n_ops; ++o) {
      if (ops[o].kind == QLS_RS_STAGE) {
        const uint32_t tb = gather_segs32((uint32_t)tid, ops[o].tseg, ops[o].n_tseg);
        ptb[st * RS_THREADS + tid] = (uint16_t)swz<SH>(tb);
        ++st;
      }
    }
  }
  // L2 prefetch geometry: the tile is 2^n_high runs of 2^L amplitudes; units of <= 4 KB
  const uint32_t run_bytes = (uint32_t)sizeof(V) << L;
  const uint32_t pf_bytes = run_bytes < 4096u ? run_bytes : 4096u;
  const uint32_t pf_units = TILE_BYTES / pf_bytes;
  const uint32_t pf_per_run_log = 31 - __clz(run_bytes / pf_bytes);
  __syncthreads();

  // ---- persistent loop over tiles ------------------------------------------------------------------
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t base = gather_segs64(tile, c_img.h.fseg, n_fseg) | c_img.h.fix_val;
    const uint64_t gbase = base | index_offset;
    {  // next tile of this CTA -> L2
      const uint64_t nt = tile + gridDim.x;
      if (nt < n_tiles) {
        const uint64_t nbase = gather_segs64(nt, c_img.h.fseg, n_fseg) | c_img.h.fix_val;
        if (!rs_tile_idle(nbase | index_offset)) {
          for (uint32_t u = tid; u < pf_units; u += RS_THREADS) {
            const uint64_t g = nbase | hoff[u >> pf_per_run_log];
            prefetch_l2_bulk(reinterpret_cast<const unsigned char*>(&state[g]) + (size_t)(u & ((1u << pf_per_run_log) - 1u)) * pf_bytes,
                             pf_bytes);
          }
        }
      }
    }
    if (uni(rs_tile_idle(gbase))) continue;  // warp- and block-uniform

    V a[A];
    int cur = -1;        // op index of the STAGE whose layout the registers are in (-1: tile not loaded yet)
    int pend = -1;       // latest STAGE record seen
    int pend_ord = -1;   // its ordinal (row of ptb)
    bool cur_direct = false;
    uint32_t mytb = 0;
    for (int o = 0; o < n_ops; ++o) {
      const QlsRsOp& op = ops[o];
      if (uni(op.kind == QLS_RS_STAGE)) {
        pend = o;
        ++pend_ord;
        continue;
      }
      if (uni((gbase & op.xc_mask) != op.xc_val)) continue;  // block-uniform
      if (uni(pend != cur)) {
        // ---- enter the pending stage (barriers are unconditional inside this block) ----
        const QlsRsOp& sp = ops[pend];
        const bool was_loaded = cur >= 0;
        const bool new_direct = (sp.flags & QLS_RS_FLAG_DIRECT) != 0;
        __syncthreads();  // every thread is done with the previous shared-memory image
        if (was_loaded) {
          uint32_t cq[KR], lq[KR];
          rs_stage_bits<KR, SH>(ops[cur], cq, lq);
#pragma unroll
          for (int r = 0; r < A; ++r) sm[rs_addr<KR>(mytb, cq, r)] = a[r];
        } else if (!new_direct) {
#pragma unroll 4
          for (int u = tid; u < UNITS; u += RS_THREADS) {
            const int i = u * APU;
            const uint64_t g = base | (uint64_t)(i & ((1 << L) - 1)) | hoff[i >> L];
            cp_async16(&sm16[swz<SH>((uint32_t)i) >> SH], &state[g]);
          }
          cp_async_commit();
          cp_async_wait<0>();
        }
        mytb = ptb[pend_ord * RS_THREADS + tid];
        uint32_t cq[KR], lq[KR];  // swizzled shared-memory offset / tile-local bit of each register qubit
        rs_stage_bits<KR, SH>(sp, cq, lq);
        __syncthreads();
        if (!was_loaded && new_direct) {
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
        }
        cur_direct = new_direct;
        cur = pend;
      }
      const bool tok = ((uint32_t)tid & op.ctl_tmask) == op.ctl_tval;
      switch (op.kind) {
        case QLS_RS_G1:
          rs_dispatch_g1<V, A, KR>(a, op, reinterpret_cast<const V*>(c_img.body + op.pool_off), tok);
          break;
        case QLS_RS_G2:
          rs_dispatch_g2<V, A, KR>(a, op, reinterpret_cast<const V*>(c_img.body + op.pool_off), tok);
          break;
        case QLS_RS_DIAG: {
          const uint32_t ti0 = (uint32_t)gather_segs64(gbase, op.xseg, op.n_xseg) | gather_segs32((uint32_t)tid, op.tseg, op.n_tseg);
          rs_diag<V, A, KR>(a, op, reinterpret_cast<const V*>(pool + op.pool_off), ti0, tok);
          break;
        }
        case QLS_RS_MUXRY: {
          const uint32_t ti0 = (uint32_t)gather_segs64(gbase, op.xseg, op.n_xseg) | gather_segs32((uint32_t)tid, op.tseg, op.n_tseg);
          rs_dispatch_muxry<V, A, KR>(a, op, reinterpret_cast<const V*>(pool + op.pool_off), ti0, tok);
          break;
        }
      }
    }
    if (uni(cur < 0)) continue;  // nothing fired (possible only when tile skipping is off): tile untouched
    uint32_t cq[KR], lq[KR];
    rs_stage_bits<KR, SH>(ops[cur], cq, lq);
    __syncthreads();
    if (!cur_direct) {
#pragma unroll
      for (int r = 0; r < A; ++r) sm[rs_addr<KR>(mytb, cq, r)] = a[r];
    }
    __syncthreads();
    if (cur_direct) {
      const uint32_t ltb = swz<SH>(mytb);
#pragma unroll
      for (int r = 0; r < A; ++r) {
        const uint32_t idx = rs_addr<KR>(ltb, lq, r);
        state[base | (uint64_t)(idx & ((1u << L) - 1u)) | hoff[idx >> L]] = a[r];
      }
    } else {
#pragma unroll 4
      for (int u = tid; u < UNITS; u += RS_THREADS) {
        const int i = u * APU;
        const uint64_t g = base | (uint64_t)(i & ((1 << L) - 1)) | hoff[i >> L];
        *reinterpret_cast<uint4*>(&state[g]) = sm16[swz<SH>((uint32_t)i) >> SH];
      }
    }
  }
}

template <typename R, int T, int KR>
int launch_rs(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, void* state, int n_local, uint64_t index_offset,
              cudaStream_t stream) {
  using V = typename Vec2<R>::type;
  const int fixed = __builtin_popcountll(p.fix_mask);
  const int free_bits = n_local - p.tile_bits - fixed;
  QLS_ARG(free_bits >= 0 && free_bits < 40, "pass geometry: n_local=%d tile_bits=%d fixed=%d", n_local, p.tile_bits, fixed);
  const uint64_t n_tiles = 1ull << free_bits;
  const int n_stages = p.pad >> 8;  // counted by qls_program_create
  const size_t smem = ((size_t)sizeof(V) << T) + (sizeof(uint64_t) << p.n_high) + (size_t)n_stages * RS_THREADS * sizeof(uint16_t);
  QLS_ARG(smem <= 112 * 1024, "register-stage pass needs %zu bytes of shared memory", smem);
  auto kern = qls_rs_pass_kernel<R, T, KR>;
  static bool attr_set[64] = {false};
  static int sm_count[64] = {0};
  if (!attr_set[prog->device & 63]) {
    QLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    QLS_CUDA(cudaDeviceGetAttribute(&sm_count[prog->device & 63], cudaDevAttrMultiProcessorCount, prog->device));
    attr_set[prog->device & 63] = true;
  }
  // the pass description goes to constant memory; stream order keeps it alive for exactly this launch
  QLS_CUDA(cudaMemcpyToSymbolAsync(c_img, prog->rs_img_dev + prog->rs_img_off[pass_index], prog->rs_img_bytes[pass_index], 0,
                                   cudaMemcpyDeviceToDevice, stream));
  uint64_t grid = 2ull * (uint64_t)sm_count[prog->device & 63];
  if (grid > n_tiles) grid = n_tiles;
  kern<<<(unsigned)grid, RS_THREADS, smem, stream>>>(reinterpret_cast<V*>(state), index_offset, n_tiles, prog->pool_dev);
  QLS_CUDA(cudaGetLastError());
  return QLS_OK;
}

}  // namespace

// Build the constant-memory image of every pass that has a register-stage form: header, micro-op
// records, and the G1/G2 matrices (their pool_off re-based to the image body).  A pass whose image
// would not fit gets size 0 and keeps the sweep form.  Host arrays; called by qls_program_create.
int qls_rs_build_images(QlsProgram* prog, const QlsRsOp* rs_ops_host, const unsigned char* pool_host) {
  const size_t esz = prog->dtype == QLS_C128 ? 16 : 8;
  std::vector<unsigned char> all;
  prog->rs_img_off = new uint64_t[prog->n_passes ? prog->n_passes : 1];
  prog->rs_img_bytes = new uint32_t[prog->n_passes ? prog->n_passes : 1];
  for (uint32_t i = 0; i < prog->n_passes; ++i) {
    const QlsPass& p = prog->passes_host[i];
    prog->rs_img_off[i] = all.size();
    prog->rs_img_bytes[i] = 0;
    const int n_stages = (int)(p.pad >> 8);
    if (p.rs_count == 0 || p.rs_count > (uint32_t)RS_MAX_OPS || n_stages < 1 || n_stages > RS_MAX_STAGES) continue;
    std::vector<unsigned char> img(sizeof(RsHeader) + (size_t)p.rs_count * sizeof(QlsRsOp), 0);
    RsHeader h;
    memset(&h, 0, sizeof(h));
    h.L = p.low_bits;
    h.n_high = p.n_high;
    h.n_fseg = p.n_fseg;
    h.n_ops = (int32_t)p.rs_count;
    h.fix_val = p.fix_val;
    h.n_stages = n_stages;
    memcpy(h.high, p.high, sizeof(h.high));
    memcpy(h.fseg, p.fseg, sizeof(h.fseg));
    bool skippable = true;
    int n_xc = 0;
    for (uint32_t o = 0; o < p.rs_count; ++o) {
      QlsRsOp op = rs_ops_host[p.rs_begin + o];
      if (op.kind == QLS_RS_G1 || op.kind == QLS_RS_G2) {
        const size_t bytes = esz * (op.kind == QLS_RS_G1 ? 4 : 16);
        const size_t off = img.size();
        img.resize(off + bytes);
        memcpy(img.data() + off, pool_host + op.pool_off, bytes);
        op.pool_off = off - sizeof(RsHeader);
      }
      if (op.kind != QLS_RS_STAGE) {
        if (op.xc_mask == 0) {
          skippable = false;
        } else if (skippable) {
          int j = 0;
          while (j < n_xc && !(h.xc_mask[j] == op.xc_mask && h.xc_val[j] == op.xc_val)) ++j;
          if (j == n_xc) {
            if (n_xc == RS_MAX_XC) {
              skippable = false;
            } else {
              h.xc_mask[n_xc] = op.xc_mask;
              h.xc_val[n_xc] = op.xc_val;
              ++n_xc;
            }
          }
        }
      }
      memcpy(img.data() + sizeof(RsHeader) + (size_t)o * sizeof(QlsRsOp), &op, sizeof(op));
    }
    h.n_xc = skippable ? n_xc : 0;
    memcpy(img.data(), &h, sizeof(h));
    while (img.size() % 16) img.push_back(0);
    if (img.size() > (size_t)RS_IMG_BYTES) continue;
    prog->rs_img_bytes[i] = (uint32_t)img.size();
    all.insert(all.end(), img.begin(), img.end());
  }
  prog->rs_img_dev = nullptr;
  if (!all.empty()) {
    QLS_CUDA(cudaMalloc(&prog->rs_img_dev, all.size()));
    QLS_CUDA(cudaMemcpy(prog->rs_img_dev, all.data(), all.size(), cudaMemcpyHostToDevice));
  }
  return QLS_OK;
}

// true if pass `pass_index` of `prog` can run in register-stage form on a state of n_local qubits
bool qls_rs_eligible(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, int n_local) {
  const int want_T = prog->dtype == QLS_C128 ? 12 : 13;
  return prog->rs_img_bytes && prog->rs_img_bytes[pass_index] > 0 && p.tile_bits == want_T && n_local >= want_T;
}

int qls_launch_rs_pass(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, void* state, int n_local,
                       uint64_t index_offset, cudaStream_t stream) {
  if (prog->dtype == QLS_C128) return launch_rs<double, 12, 4>(prog, p, pass_index, state, n_local, index_offset, stream);
  return launch_rs<float, 13, 5>(prog, p, pass_index, state, n_local, index_offset, stream);
}
