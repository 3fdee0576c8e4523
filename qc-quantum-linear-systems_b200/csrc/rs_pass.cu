// Register-stage tile-pass kernel (the fast path for full-size tiles; 98 % of the GPU time of a solve).
//
// Persistent CTAs (2 per SM, 256 threads, 128 registers).  A tile of 2^T amplitudes (T = 12 complex128 /
// 13 complex64 -> 64 KB) lives in the REGISTERS of the CTA: every thread holds 2^KR amplitudes
// (KR = 4 / 5: the stage's "register qubits"), the remaining 8 tile bits are the thread index.  All
// micro-ops of a stage -- 1- and 2-qubit dense blocks on register qubits, diagonal tables,
// multiplexed RY -- run on registers with compile-time register indices; shared memory is touched
// only to re-distribute the tile between stages (one XOR-swizzled write + one read) and to make
// the global accesses coalesced when a stage's two lowest tile bits are not lane bits.
//
// What keeps the FP64 pipe fed (profiles/README.md has the ncu history: 27 % -> 47 % FP64-pipe activity
// on the data-register passes, 90 % for the gate core alone):
//   * the whole pass description -- header, micro-op records, gate matrices and one straight-line
//     STEP LIST per assignment of the control qubits outside the tile -- sits in __constant__ memory
//     (built by qls_rs_build_images, uploaded per pass, stream ordered).  Every thread reads it at a
//     warp-uniform address, so steps decode on the uniform datapath and the DFMAs take their matrix
//     element straight from a uniform register: no per-thread load in the gate loops, no control test,
//     no stage bookkeeping in the kernel;
//   * a step list enters a stage only if one of its ops fires for that control assignment, and a tile on
//     which no op fires is neither read nor written;
//   * tile loads/stores bypass L1 (ld/st.global.cg, cp.async.cg) so that L1 keeps the phase tables;
//     tables whose index does not depend on the lane are fetched once per warp (shared-memory scratch line).
// Variants that were measured and dropped (two token-passing warp-groups per CTA, 2^11-amplitude tiles with
// four CTAs per SM, 32 amplitudes per thread, L2 prefetch of the next tile) are described in profiles/README.md;
// the first two are still selectable (QLS_RS_PINGPONG=1, LoweringOptions.tile_bits=11).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "qls_common.cuh"

namespace {

constexpr int RS_MAX_OPS = 160;
constexpr int RS_MAX_STAGES = 24;
constexpr int RS_MAX_CBITS = 6;
constexpr int RS_IMG_BYTES = 60 * 1024;
constexpr uint32_t QLS_RS_FLAG_DIRECT = 0x10u;  // STAGE: every register qubit sits above the lane bits
constexpr uint32_t QLS_RS_FLAG_KR5 = 0x20u;     // STAGE: five register qubits (r[0..3], pad0), 2^(T-5) threads

struct RsHeader {  // 256 bytes
  int32_t L, n_high, n_fseg, n_recs;
  uint64_t fix_val;
  int32_t n_cbits;     // control qubits outside the tile whose values select the step list (variant)
  int32_t n_stages;
  uint32_t steps_off;  // byte offset of RsStep[] in the image body
  uint32_t rtw_off;    // byte offset of uint32_t[n_recs][32]: rt[] widened to 32-bit words (single precision only)
  uint8_t cbit[8];     // their global index bit positions
  uint8_t high[QLS_MAX_HIGH];
  QlsSeg fseg[QLS_MAX_FSEG];
  uint16_t vstart[1 << RS_MAX_CBITS];  // first step of each variant; RS_IDLE: no op fires, tile untouched
  uint8_t pad[8];
};
static_assert(sizeof(RsHeader) == 256, "RsHeader layout");

// One step of a variant's straight-line program (8 bytes, constant memory).
struct RsStep {
  uint16_t code;  // RS_C_*
  uint16_t rec;   // index of the QlsRsOp record (op fields / the stage entered)
  uint16_t aux;   // TRANS: record index of the stage being left
  uint16_t moff;  // G1/G2: matrix offset in the image body, in 16-byte units
};

enum {  // dense numbering: the step switch compiles to one jump table
  RS_C_LOAD_D = 0,   // global -> registers (lanes are the low tile bits)
  RS_C_LOAD_S = 1,   // global -> shared (cp.async) -> registers
  RS_C_TRANS = 2,    // registers -> shared -> registers in the next stage's layout
  RS_C_STORE_D = 3,  // registers -> global; ends the tile
  RS_C_STORE_S = 4,  // registers -> shared -> global; ends the tile
  RS_C_ACQ = 5,      // take the SM's FP64 token (two-group kernels): a run of heavy ops follows
  RS_C_DIAG = 6,     // table indexed by register qubits, no control inside the tile
  RS_C_DIAGC = 7,    // ... with a control inside the tile
  RS_C_DIAG1 = 8,    // one phase per thread
  RS_C_DIAGW = 9,    // table index is the same for all lanes of a warp: the entries are fetched once per warp
  RS_C_DIAGWC = 10,  // ... controlled by register qubits
  RS_C_MUX = 11,     // + register qubit (0..4)
  RS_C_G1 = 16,      // + 4 * register qubit + 2 * real + ctl
  RS_C_G2 = 36       // + 4 * pair + 2 * real + ctl
};
constexpr uint16_t RS_F_RELEASE = 1;  // RsStep.moff of TRANS / STORE_*: hand the FP64 token back first
constexpr uint16_t RS_IDLE = 0xFFFFu;

// pair index of register qubits p < q: the six pairs below 4 first (KR = 4), then (p, 4)
__host__ __device__ constexpr int rs_pair(int p, int q) { return q == 4 ? 6 + p : (p == 0 ? q - 1 : p + q); }
static_assert(rs_pair(0, 1) == 0 && rs_pair(0, 2) == 1 && rs_pair(0, 3) == 2 && rs_pair(1, 2) == 3 && rs_pair(1, 3) == 4 &&
                  rs_pair(2, 3) == 5 && rs_pair(0, 4) == 6 && rs_pair(3, 4) == 9, "pair numbering");

struct RsImage {
  RsHeader h;
  unsigned char body[RS_IMG_BYTES - sizeof(RsHeader)];  // QlsRsOp[n_recs], gate matrices, RsStep[]
};

__constant__ RsImage c_img;
static QlsConstGuard g_img_guard;  // every instantiation of the kernel reads the one image of its device

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

// The same contribution read from the image's widened copy of rt[] (32-bit words, RsHeader.rtw_off).  The
// single-precision kernel must use this form: with rs_tidx over five 16-bit rt[] loads ptxas leaves the uniform
// datapath for the WHOLE step loop (FFMA..UR count 0, gate matrices fetched per thread with LDC).
template <typename V, int KR>
__device__ __forceinline__ uint32_t rs_tcombo(const uint32_t (&d)[KR], const uint32_t* __restrict__ rw, int r) {
  if constexpr (sizeof(V) == 8) return rw[r];
  else return rs_tidx<KR>(0u, d, r);
}

// one phase per thread: the table looks at no register qubit and there is no register-bit control
template <typename V, int A>
__device__ __forceinline__ void rs_diag1(V (&a)[A], const V* __restrict__ table, uint32_t ti0, bool tok) {
  V ph = __ldg(&table[ti0]);
  if (!tok) {
    ph.x = 1;
    ph.y = 0;
  }
#pragma unroll
  for (int r = 0; r < A; ++r) a[r] = cmul(a[r], ph);
}

template <typename V, int A, int KR, bool CTL>
__device__ __forceinline__ void rs_diag(V (&a)[A], const QlsRsOp& op, const uint32_t* __restrict__ rw, const V* __restrict__ table,
                                        uint32_t ti0, bool tok) {
  uint32_t d[KR];
#pragma unroll
  for (int j = 0; j < KR; ++j) d[j] = sizeof(V) == 8 ? 0u : (uint32_t)op.rt[1 << j];
  const uint32_t rm = op.ctl_rmask, rv = op.ctl_rval;
  constexpr int CH = 4;  // phases fetched per batch (register budget)
  const V* tb = table + ti0;  // the register-qubit fields are disjoint from ti0: | is +, and they are warp-uniform
#pragma unroll
  for (int c = 0; c < A; c += CH) {
    V ph[CH];
#pragma unroll
    for (int r = 0; r < CH; ++r) ph[r] = __ldg(tb + rs_tcombo<V, KR>(d, rw, c + r));  // index is valid whatever the controls say
#pragma unroll
    for (int r = 0; r < CH; ++r)
      if (!CTL || (tok && (((uint32_t)(c + r) & rm) == rv))) a[c + r] = cmul(a[c + r], ph[r]);
  }
}

// Table whose index does not depend on the lane (QFT phase ladders on the clock register while the
// lanes are data qubits): the warp needs only the 2^KR entries of its register combinations.  Lane r
// fetches entry r, the warp shares them through a 2^KR-entry scratch line in shared memory
// (broadcast reads), instead of every lane loading all of them.
// CTL: control on register qubits only (warp-uniform predicate; a per-thread predicate between the two
// warp barriers makes ptxas give up the uniform datapath for the whole step loop).
template <typename V, int A, int KR, bool CTL>
__device__ __forceinline__ void rs_diagw(V (&a)[A], const QlsRsOp& op, const uint32_t* __restrict__ rw, const V* __restrict__ table,
                                         uint32_t ti0, V* __restrict__ scr, int lane) {
  {  // every lane takes part (lanes >= A duplicate entry lane % A): no divergent branch in the step loop
    const int e = lane & (A - 1);
    uint32_t t = ti0;
    if constexpr (sizeof(V) == 8) {
      t |= rw[e];
    } else {
#pragma unroll
      for (int j = 0; j < KR; ++j) t |= ((e >> j) & 1) ? (uint32_t)op.rt[1 << j] : 0u;
    }
    scr[e] = __ldg(&table[t]);
  }
  __syncwarp();
  const uint32_t rm = op.ctl_rmask, rv = op.ctl_rval;
#pragma unroll
  for (int r = 0; r < A; ++r) {
    const V ph = scr[r];
    if (!CTL || (((uint32_t)r & rm) == rv)) a[r] = cmul(a[r], ph);
  }
  __syncwarp();  // the scratch line is free for the next table
}

template <typename V, int A, int KR, int RQ>
__device__ __forceinline__ void rs_muxry(V (&a)[A], const QlsRsOp& op, const uint32_t* __restrict__ rw, const V* __restrict__ table,
                                         uint32_t ti0, bool tok) {
  uint32_t d[KR];
#pragma unroll
  for (int j = 0; j < KR; ++j) d[j] = sizeof(V) == 8 ? 0u : (uint32_t)op.rt[1 << j];
  const uint32_t rm = op.ctl_rmask, rv = op.ctl_rval;
  constexpr int CH = 4;
  const V* tb = table + ti0;
#pragma unroll
  for (int c = 0; c < A / 2; c += CH) {
    V cs[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int p = c + q;
      const int i0 = ((p >> RQ) << (RQ + 1)) | (p & ((1 << RQ) - 1));
      cs[q] = __ldg(tb + rs_tcombo<V, KR>(d, rw, i0));
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

// tile-local bit (lq) and swizzled shared-memory offset (cq) of each register qubit of a stage.
// (Reading them precomputed from the record's rt[] makes ptxas drop the uniform datapath for the whole
// step loop -- measured with the DFMA.*UR count -- so they are recomputed: a dozen uniform ALU ops.)
template <int KR, int SH>
__device__ __forceinline__ void rs_stage_bits(const QlsRsOp& sp, uint32_t (&cq)[KR], uint32_t (&lq)[KR]) {
#pragma unroll
  for (int j = 0; j < KR; ++j) {
    lq[j] = 1u << (j < 4 ? sp.r[j] : sp.pad0);
    cq[j] = swz<SH>(lq[j]);
  }
}

// variant of the tile whose global base index is gbase: the values of the pass's control qubits
__device__ __forceinline__ uint32_t rs_variant(uint64_t gbase) {
  uint32_t v = 0;
  const int n = c_img.h.n_cbits;
  for (int j = 0; j < n; ++j) v |= (uint32_t)((gbase >> c_img.h.cbit[j]) & 1ull) << j;
  return v;
}

// The per-tile program is a straight list of RsStep records chosen by the tile's control-qubit
// values (built on the host, qls_rs_build_images): no control test, no stage bookkeeping and no
// loop-carried state besides the step counter is left in the kernel.  ptxas keeps the whole step
// loop on the uniform datapath only if it can PROVE every branch around a barrier warp-uniform; a
// warp shuffle of the step counter makes that proof trivial (loop-carried state it cannot see
// through makes it fall back to per-thread LDC + vector-register operands, and the kernel spills).
// GROUPS = 2: one CTA per SM holds TWO warp-groups, each working on its own tile in its own 64 KB of
// shared memory with its own named barrier.  Identical CTAs that merely share an SM fall into
// lockstep (whoever gets ahead leaves the FP64 pipe to the other one, which then catches up), so
// HBM, shared-memory and FP64 phases add up instead of overlapping (measured: profiles/).  The two
// groups therefore pass an FP64 token (RS_C_ACQ / RS_F_RELEASE steps placed by the host around runs
// of heavy ops): while one group computes, the other one loads, re-distributes or stores.
template <int ID_BASE>
__device__ __forceinline__ void rs_group_sync(uint32_t group, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(ID_BASE + group), "r"(threads) : "memory");
}

template <typename R, int T, int KR, int CTAS, int GROUPS>
__global__ void __launch_bounds__(GROUPS << (T - KR), CTAS)
    qls_rs_pass_kernel(typename Vec2<R>::type* __restrict__ state, const uint64_t index_offset, const uint64_t n_tiles,
                       const unsigned char* __restrict__ pool, const uint32_t dbg) {
  using V = typename Vec2<R>::type;
  constexpr int RS_THREADS = 1 << (T - KR);  // the threads index the non-register tile bits
  constexpr int A = 1 << KR;
  constexpr int SH = sizeof(V) == 16 ? 0 : 1;
  constexpr int APU = 1 << SH;           // amplitudes per 16-byte unit
  constexpr int UNITS = (1 << T) / APU;  // 4096
  constexpr uint32_t TILE_BYTES = (uint32_t)sizeof(V) << T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int fp64_token;  // 0: free
  // warp-group of this thread (provably warp-uniform for ptxas: shuffle) and its thread index inside it
  const uint32_t grp = GROUPS == 1 ? 0u : __shfl_sync(0xffffffffu, threadIdx.x >> (T - KR), 0);
  const int tid = threadIdx.x & (RS_THREADS - 1);
  V* sm = reinterpret_cast<V*>(smem_raw + (size_t)grp * TILE_BYTES);
  uint4* sm16 = reinterpret_cast<uint4*>(smem_raw + (size_t)grp * TILE_BYTES);
  uint64_t* hoff = reinterpret_cast<uint64_t*>(smem_raw + (size_t)GROUPS * TILE_BYTES);  // [1 << n_high]
  uint16_t* ptb = reinterpret_cast<uint16_t*>(hoff + ((1 << c_img.h.n_high) | 1) + 1);  // [stages][RS_THREADS]; even entry count: 16-byte aligned
  // per-warp scratch line of A table entries (RS_C_DIAGW), behind ptb (16-byte aligned: ptb rows are 2 * RS_THREADS bytes)
  V* wscr = reinterpret_cast<V*>(ptb + (size_t)c_img.h.n_stages * RS_THREADS) + (threadIdx.x >> 5) * A;
#define RS_SYNC() do { if (GROUPS == 1) __syncthreads(); else rs_group_sync<1>(grp, RS_THREADS); } while (0)

  const QlsRsOp* recs = reinterpret_cast<const QlsRsOp*>(c_img.body);
  const int L = c_img.h.L;
  const int n_recs = c_img.h.n_recs;
  const int n_high = c_img.h.n_high;
  const int n_fseg = c_img.h.n_fseg;

  // ---- once per CTA: high-qubit spread table, per-stage thread -> swizzled tile offset ----------------
  for (int r = threadIdx.x; r < (1 << n_high); r += GROUPS * RS_THREADS) {
    uint64_t g = 0;
    for (int j = 0; j < n_high; ++j) g |= (uint64_t)((r >> j) & 1) << c_img.h.high[j];
    hoff[r] = g;
  }
  if (threadIdx.x == 0) fp64_token = 0;
  for (int o = 0; o < n_recs && threadIdx.x < RS_THREADS; ++o) {
    if (recs[o].kind == QLS_RS_STAGE) {
      const uint32_t tb = gather_segs32((uint32_t)tid, recs[o].tseg, recs[o].n_tseg);
      ptb[(uint32_t)recs[o].pool_off * RS_THREADS + tid] = (uint16_t)swz<SH>(tb);  // pool_off of a STAGE = its ordinal
    }
  }
  __syncthreads();

  // ---- persistent loop over tiles ------------------------------------------------------------------
  const uint64_t tile_stride = (uint64_t)GROUPS * gridDim.x;
  for (uint64_t tile = (uint64_t)GROUPS * blockIdx.x + grp; tile < n_tiles; tile += tile_stride) {
    const uint64_t base = gather_segs64(tile, c_img.h.fseg, n_fseg) | c_img.h.fix_val;
    const uint64_t gbase = base | index_offset;
    uint32_t s = __shfl_sync(0xffffffffu, (uint32_t)c_img.h.vstart[rs_variant(gbase)], 0);
    if (s == RS_IDLE) continue;  // no op fires on this tile: neither read nor written
    const RsStep* steps = reinterpret_cast<const RsStep*>(c_img.body + c_img.h.steps_off);

    V a[A];
    for (;;) {
      s = __shfl_sync(0xffffffffu, s, 0);  // provably warp-uniform step counter (see above)
      const RsStep st = steps[s];
      ++s;
      const QlsRsOp& op = recs[st.rec];
      const V* M = reinterpret_cast<const V*>(c_img.body + 16u * st.moff);
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(c_img.body + c_img.h.rtw_off) + 32u * st.rec;  // single precision only
#ifdef QLS_RS_TIMING  // timing experiments only (profiles/README.md): switch parts of the kernel off
      const bool skip = KR == 4 && sizeof(V) == 16 && dbg != 0 && (((dbg & 1u) && st.code >= RS_C_G1) || ((dbg & 2u) && st.code == RS_C_TRANS) ||
                                    ((dbg & 4u) && st.code >= RS_C_DIAG && st.code < RS_C_G1));
      if (!__any_sync(0xffffffffu, skip))
#endif
      switch (st.code) {
        case RS_C_LOAD_D: {
          // all register qubits sit above the 5 lane bits: every warp load is 512 contiguous bytes
          uint32_t cq[KR], lq[KR];
          rs_stage_bits<KR, SH>(op, cq, lq);
          const uint32_t ltb = swz<SH>(ptb[(uint32_t)op.pool_off * RS_THREADS + tid]);  // the swizzle is an involution
#pragma unroll
          for (int r = 0; r < A; ++r) {
            const uint32_t idx = rs_addr<KR>(ltb, lq, r);
            a[r] = __ldcg(&state[base | (uint64_t)(idx & ((1u << L) - 1u)) | hoff[idx >> L]]);  // L2 only: L1 is kept for the phase tables
          }
          break;
        }
#ifndef QLS_RS_NO_LOADS
        case RS_C_LOAD_S: {
          uint32_t cq[KR], lq[KR];
          rs_stage_bits<KR, SH>(op, cq, lq);
          const uint32_t mytb = ptb[(uint32_t)op.pool_off * RS_THREADS + tid];
          RS_SYNC();  // the previous tile's write-back has left shared memory
#pragma unroll 4
          for (int u = tid; u < UNITS; u += RS_THREADS) {
            const int i = u * APU;
            const uint64_t g = base | (uint64_t)(i & ((1 << L) - 1)) | hoff[i >> L];
            cp_async16(&sm16[swz<SH>((uint32_t)i) >> SH], &state[g]);
          }
          cp_async_commit();
          cp_async_wait<0>();
          RS_SYNC();
#pragma unroll
          for (int r = 0; r < A; ++r) a[r] = sm[rs_addr<KR>(mytb, cq, r)];
          break;
        }
#endif
#ifndef QLS_RS_NO_TRANS
        case RS_C_TRANS: {
          const QlsRsOp& from = recs[st.aux];
          uint32_t cq[KR], lq[KR];
          rs_stage_bits<KR, SH>(from, cq, lq);
          uint32_t mytb = ptb[(uint32_t)from.pool_off * RS_THREADS + tid];
          RS_SYNC();  // every thread is done with the previous shared-memory image (and with its FP64 run)
          if (GROUPS > 1 && st.moff == RS_F_RELEASE && tid == 0) atomicExch(&fp64_token, 0);
#pragma unroll
          for (int r = 0; r < A; ++r) sm[rs_addr<KR>(mytb, cq, r)] = a[r];
          rs_stage_bits<KR, SH>(op, cq, lq);
          mytb = ptb[(uint32_t)op.pool_off * RS_THREADS + tid];
          RS_SYNC();
#pragma unroll
          for (int r = 0; r < A; ++r) a[r] = sm[rs_addr<KR>(mytb, cq, r)];
          break;
        }
#endif
        case RS_C_STORE_D: {
          if (GROUPS > 1 && st.moff == RS_F_RELEASE) {
            RS_SYNC();
            if (tid == 0) atomicExch(&fp64_token, 0);
          }
          uint32_t cq[KR], lq[KR];
          rs_stage_bits<KR, SH>(op, cq, lq);
          const uint32_t ltb = swz<SH>(ptb[(uint32_t)op.pool_off * RS_THREADS + tid]);
#pragma unroll
          for (int r = 0; r < A; ++r) {
            const uint32_t idx = rs_addr<KR>(ltb, lq, r);
            __stcg(&state[base | (uint64_t)(idx & ((1u << L) - 1u)) | hoff[idx >> L]], a[r]);
          }
          break;
        }
#ifndef QLS_RS_NO_STORES
        case RS_C_STORE_S: {
          uint32_t cq[KR], lq[KR];
          rs_stage_bits<KR, SH>(op, cq, lq);
          const uint32_t mytb = ptb[(uint32_t)op.pool_off * RS_THREADS + tid];
          RS_SYNC();
          if (GROUPS > 1 && st.moff == RS_F_RELEASE && tid == 0) atomicExch(&fp64_token, 0);
#pragma unroll
          for (int r = 0; r < A; ++r) sm[rs_addr<KR>(mytb, cq, r)] = a[r];
          RS_SYNC();
#pragma unroll 4
          for (int u = tid; u < UNITS; u += RS_THREADS) {
            const int i = u * APU;
            const uint64_t g = base | (uint64_t)(i & ((1 << L) - 1)) | hoff[i >> L];
            __stcg(reinterpret_cast<uint4*>(&state[g]), sm16[swz<SH>((uint32_t)i) >> SH]);
          }
          break;
        }
#endif
#ifndef QLS_RS_NO_ACQ
        case RS_C_ACQ: {
          if (GROUPS > 1) {
            if (tid == 0) {
              while (atomicCAS(&fp64_token, 0, 1) != 0) __nanosleep(32);
            }
            RS_SYNC();
          }
          break;
        }
#endif
#ifndef QLS_RS_NO_DIAGW
        case RS_C_DIAGW:
        case RS_C_DIAGWC: {
          const uint32_t ti0 = (uint32_t)gather_segs64(gbase, op.xseg, op.n_xseg) | gather_segs32((uint32_t)tid, op.tseg, op.n_tseg);
          const V* table = reinterpret_cast<const V*>(pool + op.pool_off);
          rs_diagw<V, A, KR, true>(a, op, rw, table, ti0, wscr, tid & 31);  // one instantiation: (r & 0) == 0 without a control
          break;
        }
#endif
#ifndef QLS_RS_NO_DIAG
        case RS_C_DIAG:
        case RS_C_DIAGC:
        case RS_C_DIAG1: {
          const uint32_t ti0 = (uint32_t)gather_segs64(gbase, op.xseg, op.n_xseg) | gather_segs32((uint32_t)tid, op.tseg, op.n_tseg);
          const V* table = reinterpret_cast<const V*>(pool + op.pool_off);
          if (st.code == RS_C_DIAG) rs_diag<V, A, KR, false>(a, op, rw, table, ti0, true);
          else if (st.code == RS_C_DIAGC) rs_diag<V, A, KR, true>(a, op, rw, table, ti0, ((uint32_t)tid & op.ctl_tmask) == op.ctl_tval);
          else rs_diag1<V, A>(a, table, ti0, ((uint32_t)tid & op.ctl_tmask) == op.ctl_tval);
          break;
        }
#endif
#define QLS_MUX_CASE(q)                                                                                              \
  case RS_C_MUX + q: {                                                                                               \
    const bool tok = ((uint32_t)tid & op.ctl_tmask) == op.ctl_tval;                                                  \
    const uint32_t ti0 = (uint32_t)gather_segs64(gbase, op.xseg, op.n_xseg) | gather_segs32((uint32_t)tid, op.tseg, op.n_tseg); \
    rs_muxry<V, A, KR, q>(a, op, rw, reinterpret_cast<const V*>(pool + op.pool_off), ti0, tok);                          \
    break;                                                                                                           \
  }
#ifndef QLS_RS_NO_MUX
        QLS_MUX_CASE(0)
        QLS_MUX_CASE(1)
        QLS_MUX_CASE(2)
        QLS_MUX_CASE(3)
#endif
#define QLS_G1_CASE(q)                                                                                               \
  case RS_C_G1 + 4 * q + 0: rs_g1<V, A, q, false, false>(a, M, true, 0, 0); break;                                   \
  case RS_C_G1 + 4 * q + 2: rs_g1<V, A, q, true, false>(a, M, true, 0, 0); break;                                    \
  case RS_C_G1 + 4 * q + 1:                                                                                          \
    rs_g1<V, A, q, false, true>(a, M, ((uint32_t)tid & op.ctl_tmask) == op.ctl_tval, op.ctl_rmask, op.ctl_rval);      \
    break;                                                                                                           \
  case RS_C_G1 + 4 * q + 3:                                                                                          \
    rs_g1<V, A, q, true, true>(a, M, ((uint32_t)tid & op.ctl_tmask) == op.ctl_tval, op.ctl_rmask, op.ctl_rval);       \
    break;
#ifndef QLS_RS_NO_G1
        QLS_G1_CASE(0)
        QLS_G1_CASE(1)
        QLS_G1_CASE(2)
        QLS_G1_CASE(3)
#endif
#define QLS_G2_CASE(p, q)                                                                                            \
  case RS_C_G2 + 4 * rs_pair(p, q) + 0: rs_g2<V, A, p, q, false, false>(a, M, true, 0, 0); break;                    \
  case RS_C_G2 + 4 * rs_pair(p, q) + 2: rs_g2<V, A, p, q, true, false>(a, M, true, 0, 0); break;                     \
  case RS_C_G2 + 4 * rs_pair(p, q) + 1:                                                                              \
    rs_g2<V, A, p, q, false, true>(a, M, ((uint32_t)tid & op.ctl_tmask) == op.ctl_tval, op.ctl_rmask, op.ctl_rval);   \
    break;                                                                                                           \
  case RS_C_G2 + 4 * rs_pair(p, q) + 3:                                                                              \
    rs_g2<V, A, p, q, true, true>(a, M, ((uint32_t)tid & op.ctl_tmask) == op.ctl_tval, op.ctl_rmask, op.ctl_rval);    \
    break;
#ifndef QLS_RS_NO_G2
        QLS_G2_CASE(0, 1)
        QLS_G2_CASE(0, 2)
        QLS_G2_CASE(0, 3)
        QLS_G2_CASE(1, 2)
        QLS_G2_CASE(1, 3)
        QLS_G2_CASE(2, 3)
#endif
        default:
#ifndef QLS_RS_NO_KR5
          if constexpr (KR >= 5) {
            switch (st.code) {
              QLS_MUX_CASE(4)
              QLS_G1_CASE(4)
              QLS_G2_CASE(0, 4)
              QLS_G2_CASE(1, 4)
              QLS_G2_CASE(2, 4)
              QLS_G2_CASE(3, 4)
            }
          }
#endif
          break;
#undef QLS_MUX_CASE
#undef QLS_G1_CASE
#undef QLS_G2_CASE
      }
      if (st.code == RS_C_STORE_D || st.code == RS_C_STORE_S) break;  // the tile is back in HBM
    }
  }
}

#undef RS_SYNC

template <typename R, int T, int KR, int CTAS, int GROUPS>
int launch_rs(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, void* state, int n_local, uint64_t index_offset,
              cudaStream_t stream) {
  using V = typename Vec2<R>::type;
  constexpr int RS_THREADS = 1 << (T - KR);
  const int fixed = __builtin_popcountll(p.fix_mask);
  const int free_bits = n_local - p.tile_bits - fixed;
  QLS_ARG(free_bits >= 0 && free_bits < 40, "pass geometry: n_local=%d tile_bits=%d fixed=%d", n_local, p.tile_bits, fixed);
  const uint64_t n_tiles = 1ull << free_bits;
  const int n_stages = p.pad >> 8;  // counted by qls_program_create
  const size_t smem = ((size_t)GROUPS * sizeof(V) << T) + sizeof(uint64_t) * (((1u << p.n_high) | 1u) + 1u) + (size_t)n_stages * RS_THREADS * sizeof(uint16_t) +
                      (size_t)(GROUPS * RS_THREADS / 32) * (sizeof(V) << KR);
  QLS_ARG(smem <= (size_t)(224 / CTAS) * 1024, "register-stage pass needs %zu bytes of shared memory", smem);
  auto kern = qls_rs_pass_kernel<R, T, KR, CTAS, GROUPS>;
  static bool attr_set[64] = {false};
  static int sm_count[64] = {0};
  if (!attr_set[prog->device & 63]) {
    QLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (224 / CTAS) * 1024));
    QLS_CUDA(cudaDeviceGetAttribute(&sm_count[prog->device & 63], cudaDevAttrMultiProcessorCount, prog->device));
    attr_set[prog->device & 63] = true;
  }
  // the pass description goes to constant memory; stream order keeps it alive for exactly this launch (and a launch on another
  // stream of the same device waits for this one: QlsConstGuard)
  QLS_ARG(g_img_guard.before(prog->device, stream) == 0, "stream wait on the previous user of the pass image failed");
  QLS_CUDA(cudaMemcpyToSymbolAsync(c_img, prog->rs_img_dev + prog->rs_img_off[pass_index], prog->rs_img_bytes[pass_index], 0,
                                   cudaMemcpyDeviceToDevice, stream));
  uint64_t grid = (uint64_t)CTAS * (uint64_t)sm_count[prog->device & 63];
  if (grid * GROUPS > n_tiles) grid = (n_tiles + GROUPS - 1) / GROUPS;
  const char* dbg_env = getenv("QLS_RS_DEBUG");
  kern<<<(unsigned)grid, GROUPS * RS_THREADS, smem, stream>>>(reinterpret_cast<V*>(state), index_offset, n_tiles, prog->pool_dev,
                                                     dbg_env ? (uint32_t)atoi(dbg_env) : 0u);
  QLS_CUDA(cudaGetLastError());
  QLS_ARG(g_img_guard.after(prog->device, stream) == 0, "event record behind the pass kernel failed");
  return QLS_OK;
}

}  // namespace

// Build the constant-memory image of every pass that has a register-stage form: header, micro-op
// records (STAGE records get their ordinal in pool_off), the G1/G2 matrices, and one straight-line
// step list per VARIANT = assignment of the pass's control qubits outside the tile.  A variant's list
// holds only the ops that fire for it, enters a stage only if one of its ops fires (LOAD before the
// first, TRANS between, STORE after the last) and is marked RS_IDLE when nothing fires.  A pass with
// more than RS_MAX_CBITS such control qubits, or whose image would not fit, gets size 0 and keeps
// the sweep form.  Host arrays; called by qls_program_create.
int qls_rs_build_images(QlsProgram* prog, const QlsRsOp* rs_ops_host, const unsigned char* pool_host) {
  const size_t esz = prog->dtype == QLS_C128 ? 16 : 8;
  const char* lock_env = getenv("QLS_RS_LOCK_MIN");
  const int lock_min = lock_env ? atoi(lock_env) : 4;  // lightest run of ops (in real-2x2 units) worth the FP64 token
  std::vector<unsigned char> all;
  prog->rs_img_off = new uint64_t[prog->n_passes ? prog->n_passes : 1];
  prog->rs_img_bytes = new uint32_t[prog->n_passes ? prog->n_passes : 1];
  prog->rs_img_kr = new uint8_t[prog->n_passes ? prog->n_passes : 1];
  for (uint32_t i = 0; i < prog->n_passes; ++i) {
    const QlsPass& p = prog->passes_host[i];
    prog->rs_img_off[i] = all.size();
    prog->rs_img_bytes[i] = 0;
    prog->rs_img_kr[i] = 0;
    const int n_stages = (int)(p.pad >> 8);
    if (p.rs_count == 0 || p.rs_count > (uint32_t)RS_MAX_OPS || n_stages < 1 || n_stages > RS_MAX_STAGES) continue;
    const int KR = (rs_ops_host[p.rs_begin].flags & QLS_RS_FLAG_KR5) ? 5 : 4;  // first record is a STAGE (checked by create)
    if (KR != (prog->dtype == QLS_C64 ? 5 : 4)) continue;  // complex64: 32 amplitudes per thread, complex128: 16
    RsHeader h;
    memset(&h, 0, sizeof(h));
    h.L = p.low_bits;
    h.n_high = p.n_high;
    h.n_fseg = p.n_fseg;
    h.n_recs = (int32_t)p.rs_count;
    h.fix_val = p.fix_val;
    h.n_stages = n_stages;
    memcpy(h.high, p.high, sizeof(h.high));
    memcpy(h.fseg, p.fseg, sizeof(h.fseg));

    std::vector<QlsRsOp> recs(rs_ops_host + p.rs_begin, rs_ops_host + p.rs_begin + p.rs_count);
    std::vector<unsigned char> mats;  // 16-byte aligned matrix store
    std::vector<uint16_t> moff(recs.size(), 0), code(recs.size(), 0);
    uint64_t cmask = 0;
    bool ok = true;
    int ordinal = 0;
    for (size_t o = 0; o < recs.size() && ok; ++o) {
      QlsRsOp& op = recs[o];
      if (op.kind == QLS_RS_STAGE) {
        op.pool_off = (uint64_t)ordinal++;
        ok = (((op.flags & QLS_RS_FLAG_KR5) ? 5 : 4) == KR);  // one register width per pass
        continue;
      }
      // controls on pass-level fixed bits hold for every tile of the pass (or never: the op is dead)
      if ((op.xc_val & p.fix_mask & op.xc_mask) != (p.fix_val & op.xc_mask & p.fix_mask)) {
        op.xc_mask = ~0ull;  // never fires: no variant has all 64 bits set
        op.xc_val = ~0ull;
        op.kind = QLS_RS_DIAG;
        code[o] = RS_C_DIAG1;
        continue;
      }
      op.xc_mask &= ~p.fix_mask;
      op.xc_val &= ~p.fix_mask;
      cmask |= op.xc_mask;
      const bool real = (op.flags & QLS_OP_FLAG_REAL) != 0;
      const bool ctl = (op.ctl_tmask | op.ctl_rmask) != 0;
      if (op.kind == QLS_RS_G1 || op.kind == QLS_RS_G2) {
        const size_t bytes = esz * (op.kind == QLS_RS_G1 ? 4 : 16);
        while (mats.size() % 16) mats.push_back(0);
        moff[o] = 0;  // patched below once the matrix base is known
        const size_t off = mats.size();
        mats.resize(off + bytes);
        memcpy(mats.data() + off, pool_host + op.pool_off, bytes);
        op.pool_off = off;  // relative to the matrix store for now
        if (op.kind == QLS_RS_G1) {
          ok = op.r[0] < KR;
          code[o] = (uint16_t)(RS_C_G1 + 4 * op.r[0] + (real ? 2 : 0) + (ctl ? 1 : 0));
        } else {
          ok = op.r[0] < op.r[1] && op.r[1] < KR;
          code[o] = (uint16_t)(RS_C_G2 + 4 * rs_pair(op.r[0], op.r[1]) + (real ? 2 : 0) + (ctl ? 1 : 0));
        }
      } else if (op.kind == QLS_RS_MUXRY) {
        ok = op.r[0] < KR;
        code[o] = (uint16_t)(RS_C_MUX + op.r[0]);
      } else {
        bool reg_bits = op.ctl_rmask != 0;
        for (int j = 0; j < KR; ++j) reg_bits |= op.rt[1 << j] != 0;
        bool lane_free = true;  // no lane bit (thread bits 0..4) in the table index
        for (int j = 0; j < op.n_tseg; ++j) lane_free &= op.tseg[j].src >= 5;
        if (op.ctl_tmask) lane_free = false;  // the warp-shared form takes register-qubit controls only
        code[o] = !reg_bits ? RS_C_DIAG1 : lane_free ? (ctl ? RS_C_DIAGWC : RS_C_DIAGW) : (ctl ? RS_C_DIAGC : RS_C_DIAG);
      }
    }
    if (!ok || __builtin_popcountll(cmask) > RS_MAX_CBITS) continue;
    for (int b = 0; b < 64; ++b)
      if ((cmask >> b) & 1) h.cbit[h.n_cbits++] = (uint8_t)b;

    const size_t recs_bytes = recs.size() * sizeof(QlsRsOp);
    const size_t mats_off = recs_bytes;  // body offset; sizeof(QlsRsOp) is a multiple of 16
    for (size_t o = 0; o < recs.size(); ++o)
      if (recs[o].kind == QLS_RS_G1 || recs[o].kind == QLS_RS_G2) moff[o] = (uint16_t)((mats_off + recs[o].pool_off) / 16);
    while (mats.size() % 16) mats.push_back(0);

    std::vector<RsStep> steps;
    for (uint32_t v = 0; v < (1u << h.n_cbits); ++v) {
      uint64_t assign = 0;
      for (int j = 0; j < h.n_cbits; ++j) assign |= (uint64_t)((v >> j) & 1u) << h.cbit[j];
      const size_t first = steps.size();
      int cur = -1, pend = -1;
      for (size_t o = 0; o < recs.size(); ++o) {
        const QlsRsOp& op = recs[o];
        if (op.kind == QLS_RS_STAGE) {
          pend = (int)o;
          continue;
        }
        if ((assign & op.xc_mask) != op.xc_val || pend < 0) continue;
        if (pend != cur) {
          const bool direct = (recs[pend].flags & QLS_RS_FLAG_DIRECT) != 0;
          RsStep st = {0, (uint16_t)pend, (uint16_t)(cur < 0 ? 0 : cur), 0};
          st.code = cur < 0 ? (direct ? RS_C_LOAD_D : RS_C_LOAD_S) : RS_C_TRANS;
          steps.push_back(st);
          cur = pend;
        }
        steps.push_back(RsStep{code[o], (uint16_t)o, 0, moff[o]});
      }
      if (cur < 0) {
        h.vstart[v] = RS_IDLE;
        continue;
      }
      const bool direct = (recs[cur].flags & QLS_RS_FLAG_DIRECT) != 0;
      steps.push_back(RsStep{(uint16_t)(direct ? RS_C_STORE_D : RS_C_STORE_S), (uint16_t)cur, 0, 0});
      // FP64 token: a run of ops between two data-movement steps that is heavy enough takes the token
      // at its start (ACQ) and the movement step that ends it gives the token back (RS_F_RELEASE)
      std::vector<RsStep> out;
      for (size_t a = first; a < steps.size();) {
        out.push_back(steps[a]);  // LOAD / TRANS (or the final STORE)
        size_t b = a + 1;
        int weight = 0;
        while (b < steps.size() && steps[b].code >= RS_C_DIAG) {
          const uint16_t c = steps[b].code;
          if (c >= RS_C_G2) weight += (c & 2) ? 2 : 4;       // real : complex 4x4
          else if (c >= RS_C_G1) weight += (c & 2) ? 1 : 2;  // real : complex 2x2
          else weight += 1;                                  // diagonal / multiplexed RY
          ++b;
        }
        if (b > a + 1 && weight >= lock_min && b < steps.size()) {
          out.push_back(RsStep{RS_C_ACQ, 0, 0, 0});
          out.insert(out.end(), steps.begin() + a + 1, steps.begin() + b);
          steps[b].moff = RS_F_RELEASE;
        } else {
          out.insert(out.end(), steps.begin() + a + 1, steps.begin() + b);
        }
        a = b;
      }
      steps.resize(first);
      steps.insert(steps.end(), out.begin(), out.end());
      h.vstart[v] = (uint16_t)first;
    }
    h.steps_off = (uint32_t)(recs_bytes + mats.size());
    // complex64 kernel: table-index contribution of every register combination as 32-bit words (see rs_tcombo)
    std::vector<uint32_t> rtw;
    if (KR == 5) {
      rtw.assign(recs.size() * 32, 0u);
      for (size_t o = 0; o < recs.size(); ++o)
        if (recs[o].kind == QLS_RS_DIAG || recs[o].kind == QLS_RS_MUXRY)
          for (int r = 0; r < 32; ++r)
            for (int j = 0; j < KR; ++j)
              if ((r >> j) & 1) rtw[o * 32 + r] |= recs[o].rt[1 << j];
      h.rtw_off = (uint32_t)(h.steps_off + steps.size() * sizeof(RsStep));
    }
    const size_t total = sizeof(RsHeader) + recs_bytes + mats.size() + steps.size() * sizeof(RsStep) + rtw.size() * sizeof(uint32_t);
    if (total > (size_t)RS_IMG_BYTES || steps.size() >= RS_IDLE || (mats_off + mats.size()) / 16 > 0xFFFFu) continue;
    std::vector<unsigned char> img(total, 0);
    memcpy(img.data(), &h, sizeof(h));
    memcpy(img.data() + sizeof(RsHeader), recs.data(), recs_bytes);
    memcpy(img.data() + sizeof(RsHeader) + recs_bytes, mats.data(), mats.size());
    memcpy(img.data() + sizeof(RsHeader) + h.steps_off, steps.data(), steps.size() * sizeof(RsStep));
    if (!rtw.empty()) memcpy(img.data() + sizeof(RsHeader) + h.rtw_off, rtw.data(), rtw.size() * sizeof(uint32_t));
    while (img.size() % 16) img.push_back(0);
    prog->rs_img_bytes[i] = (uint32_t)img.size();
    prog->rs_img_kr[i] = (uint8_t)KR;
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
  const bool t_ok = prog->dtype == QLS_C128 ? (p.tile_bits == 12 || (p.tile_bits == 11 && prog->rs_img_kr[pass_index] == 4))
                                            : p.tile_bits == 13;
  return prog->rs_img_bytes && prog->rs_img_bytes[pass_index] > 0 && t_ok && n_local >= p.tile_bits;
}

int qls_launch_rs_pass(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, void* state, int n_local,
                       uint64_t index_offset, cudaStream_t stream) {
  static int pingpong = -1;  // QLS_RS_PINGPONG=0: two independent CTAs per SM instead of two token-passing groups
  if (pingpong < 0) {
    const char* e = getenv("QLS_RS_PINGPONG");
    pingpong = e ? atoi(e) : 0;
  }
  // QLS_RS_ONLY_F32 / QLS_RS_ONLY_F64 and the QLS_RS_NO_<step> guards in the kernel exist for tools/uniformity_bisect.py
  // (compile-only experiments on ptxas' uniform-datapath decision); a product build defines none of them
#ifndef QLS_RS_ONLY_F32
  if (prog->dtype == QLS_C128) {
    if (p.tile_bits == 11) return launch_rs<double, 11, 4, 4, 1>(prog, p, pass_index, state, n_local, index_offset, stream);
    if (pingpong) return launch_rs<double, 12, 4, 1, 2>(prog, p, pass_index, state, n_local, index_offset, stream);
    return launch_rs<double, 12, 4, 2, 1>(prog, p, pass_index, state, n_local, index_offset, stream);
  }
#endif
#ifndef QLS_RS_ONLY_F64
  return launch_rs<float, 13, 5, 2, 1>(prog, p, pass_index, state, n_local, index_offset, stream);
#else
  return QLS_ERR_ARG;
#endif
}
