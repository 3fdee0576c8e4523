// In-shared-memory application of fused blocks to one tile of 2^T amplitudes.
// Shared by the streaming tile-pass kernel (tile_pass.cu) and the batched small-circuit kernel
// (batch_sv.cu), where the "tile" is the whole statevector of one circuit.
#pragma once
#include "qls_common.cuh"

// ---- attached diagonal tables ------------------------------------------------------------------------
// Diagonal runs next to an unconditional dense block are applied to that block's inputs / outputs
// while they sit in registers (QLS_OP_FLAG_ATTACH_*), so the tile is swept once instead of twice or
// three times.  `Att` caches what is uniform for the tile: table pointers and the part of each table
// index that comes from index bits outside the tile.
template <typename V>
struct Att {
  int n;
  uint32_t pre_mask;  // bit a set: attachment a applies to inputs, else to outputs
  const V* table[QLS_MAX_ATTACH];
  uint32_t xidx[QLS_MAX_ATTACH];
  const QlsOp* rec;
};

template <typename V>
__device__ __forceinline__ Att<V> make_att(const QlsOp* __restrict__ rec, int n, const unsigned char* __restrict__ pool,
                                           uint64_t gbase) {
  Att<V> a;
  a.n = n;
  a.rec = rec;
  a.pre_mask = 0;
#pragma unroll
  for (int i = 0; i < QLS_MAX_ATTACH; ++i) {
    if (i < n) {
      a.table[i] = reinterpret_cast<const V*>(pool + rec[i].pool_off);
      a.xidx[i] = (uint32_t)gather_segs64(gbase, rec[i].xseg, rec[i].n_xseg);
      if (rec[i].flags & QLS_OP_FLAG_ATTACH_PRE) a.pre_mask |= 1u << i;
    }
  }
  return a;
}

// product of the attached tables of one side (PRE / POST) at tile-local index i
template <typename V, bool PRE>
__device__ __forceinline__ V att_apply(const Att<V>& a, uint32_t i, V v) {
#pragma unroll
  for (int t = 0; t < QLS_MAX_ATTACH; ++t) {
    if (t < a.n && (((a.pre_mask >> t) & 1u) != 0) == PRE) {
      const QlsOp& r = a.rec[t];
      uint32_t ti = a.xidx[t];
      const int nl = r.n_lseg;
#pragma unroll
      for (int s = 0; s < QLS_MAX_SEG; ++s)
        if (s < nl) ti |= ((i >> r.lseg[s].src) & ((1u << r.lseg[s].width) - 1u)) << r.lseg[s].dst;
      v = cmul(v, __ldg(&a.table[t][ti]));
    }
  }
  return v;
}

// ---- DENSE: 2^K x 2^K matrix on K tile-local qubits ----------------------------------------------
// One work item = (group of 2^K amplitudes, slice of R = 2^K / S output rows).  Inputs are streamed
// from shared memory, the R accumulators live in registers, outputs are written once all reads of
// the group are done (S > 1 needs a block barrier for that; S == 1 does not).
template <typename V, int K, int S, int THREADS, bool REAL>
__device__ __forceinline__ void op_dense(V* __restrict__ sm, int T, const QlsOp& op, const V* __restrict__ M, int tid,
                                         const Att<V>& att) {
  constexpr int D = 1 << K;
  constexpr int R = D / S;
  const int gbits = T - K;
  const int ngroups = 1 << gbits;
  constexpr int GPI = THREADS / S;  // groups per iteration: all S row slices of a group run together
  uint32_t bit[K];
#pragma unroll
  for (int j = 0; j < K; ++j) bit[j] = 1u << op.tq[j];
  const uint32_t lc_mask = op.lc_mask, lc_val = op.lc_val;
  const int h = tid / GPI;  // row slice (warp-uniform)
  const bool has_pre = att.pre_mask != 0;
  const bool has_post = att.n > 0 && att.pre_mask != ((1u << att.n) - 1u);

  for (int g0 = 0; g0 < ngroups; g0 += GPI) {
    const int g = g0 + (tid % GPI);
    bool active = g < ngroups;
    uint32_t idx = 0;
    V acc[R];
    if (active) {
      idx = g;
#pragma unroll
      for (int j = 0; j < K; ++j) idx = insert_zero_bit(idx, op.tq[j]);
      active = (idx & lc_mask) == lc_val;
    }
    if (active) {
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r].x = acc[r].y = 0;
      const V* Mrow = M + ((h * R) << K);
#pragma unroll(K <= 3 ? D : 2)
      for (int c = 0; c < D; ++c) {
        uint32_t off = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) off |= ((c >> j) & 1) ? bit[j] : 0u;
        V v = sm[idx | off];
        if (has_pre) v = att_apply<V, true>(att, idx | off, v);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const V m = __ldg(&Mrow[(r << K) | c]);
          if (REAL) {
            acc[r].x = fma(m.x, v.x, acc[r].x);
            acc[r].y = fma(m.x, v.y, acc[r].y);
          } else {
            cfma(acc[r], m, v);
          }
        }
      }
    }
    if (S > 1) __syncthreads();
    if (active) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = h * R + r;
        uint32_t off = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) off |= ((row >> j) & 1) ? bit[j] : 0u;
        V o = acc[r];
        if (has_post) o = att_apply<V, false>(att, idx | off, o);
        sm[idx | off] = o;
      }
    }
    if (S > 1 && g0 + GPI < ngroups) __syncthreads();
  }
}

// ---- DENSE on 4 or 5 qubits with the matrix in __constant__ memory (complex128) ------------------------------
// The generic form above fetches every matrix element with a per-thread load (one LDG per four DFMAs, 16-deep
// accumulator blocks that spill at the 128-register bound).  Here the host copies the matrices of a pass's wide dense
// blocks into a __constant__ table before the launch (stream ordered, like the register-stage kernel's pass image): the
// element index is warp-uniform (row slice from the warp index, row and column from unrolled loops), so the FMAs take
// their matrix operand through the uniform datapath, and R = 8 accumulators (S = 4 row slices for k = 5, 2 for k = 4)
// fit the register budget.  One image per device at a time: the passes of a device are issued on ONE stream.
#define QLS_HEAVY_CONST 3072  // complex128 elements: three 32 x 32 or twelve 16 x 16 matrices (48 KB of the 64 KB bank)
#define QLS_MAX_HEAVY 12
static __constant__ double2 c_heavy[QLS_HEAVY_CONST];

template <int K, int S, int THREADS>
__device__ __forceinline__ void op_dense_cm(double2* __restrict__ sm, int T, const QlsOp& op, uint32_t cbase, int tid) {
  constexpr int D = 1 << K;
  constexpr int R = D / S;
  constexpr int GPI = THREADS / S;  // groups per iteration; GPI >= 32: the row slice is warp-uniform
  static_assert(GPI % 32 == 0, "row slice must be warp-uniform");
  const int ngroups = 1 << (T - K);
  uint32_t bit[K];
#pragma unroll
  for (int j = 0; j < K; ++j) bit[j] = 1u << op.tq[j];
  const uint32_t lc_mask = op.lc_mask, lc_val = op.lc_val;
  const int h = __shfl_sync(0xffffffffu, tid / GPI, 0);  // (provably uniform: keeps the matrix index on the uniform datapath)
  const double2* __restrict__ Mrow = c_heavy + cbase + ((h * R) << K);
  for (int g0 = 0; g0 < ngroups; g0 += GPI) {
    const int g = g0 + (tid % GPI);
    bool active = g < ngroups;
    uint32_t idx = 0;
    if (active) {
      idx = g;
#pragma unroll
      for (int j = 0; j < K; ++j) idx = insert_zero_bit(idx, op.tq[j]);
      active = (idx & lc_mask) == lc_val;
    }
    double2 acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r].x = acc[r].y = 0;
    if (active) {
#pragma unroll 4
      for (int c = 0; c < D; ++c) {
        uint32_t off = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) off |= ((c >> j) & 1) ? bit[j] : 0u;
        const double2 v = sm[idx | off];
#pragma unroll
        for (int r = 0; r < R; ++r) cfma(acc[r], Mrow[(r << K) | c], v);
      }
    }
    __syncthreads();  // every row slice has read the group's inputs
    if (active) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = h * R + r;
        uint32_t off = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) off |= ((row >> j) & 1) ? bit[j] : 0u;
        sm[idx | off] = acc[r];
      }
    }
    if (g0 + GPI < ngroups) __syncthreads();
  }
}

// ---- DIAG: amp *= table[bits of the (global) index] (times any attached tables) --------------------
template <typename V, int THREADS>
__device__ __forceinline__ void op_diag(V* __restrict__ sm, int T, const QlsOp& op, const V* __restrict__ table,
                                        uint64_t gbase, int tid, const Att<V>& att) {
  const uint32_t xidx = (uint32_t)gather_segs64(gbase, op.xseg, op.n_xseg);
  const int nl = op.n_lseg;
  const uint32_t lc_mask = op.lc_mask, lc_val = op.lc_val;
  const int n = 1 << T;
  for (int i = tid; i < n; i += THREADS) {
    if ((i & lc_mask) != lc_val) continue;
    uint32_t ti = xidx;
#pragma unroll
    for (int s = 0; s < QLS_MAX_SEG; ++s)
      if (s < nl) ti |= ((i >> op.lseg[s].src) & ((1u << op.lseg[s].width) - 1u)) << op.lseg[s].dst;
    V v = cmul(sm[i], __ldg(&table[ti]));
    if (att.n) v = att_apply<V, false>(att, i, v);
    sm[i] = v;
  }
}

// ---- MUXRY: RY(angle[index bits]) on one tile-local qubit; table entries are (cos, sin) of angle/2 -
template <typename V, int THREADS>
__device__ __forceinline__ void op_muxry(V* __restrict__ sm, int T, const QlsOp& op, const V* __restrict__ table,
                                         uint64_t gbase, int tid) {
  const uint32_t xidx = (uint32_t)gather_segs64(gbase, op.xseg, op.n_xseg);
  const int nl = op.n_lseg;
  const uint32_t lc_mask = op.lc_mask, lc_val = op.lc_val;
  const uint32_t tq = op.tq[0];
  const int n = 1 << (T - 1);
  for (int g = tid; g < n; g += THREADS) {
    const uint32_t i0 = insert_zero_bit(g, tq);
    if ((i0 & lc_mask) != lc_val) continue;
    uint32_t ti = xidx;
#pragma unroll
    for (int s = 0; s < QLS_MAX_SEG; ++s)
      if (s < nl) ti |= ((i0 >> op.lseg[s].src) & ((1u << op.lseg[s].width) - 1u)) << op.lseg[s].dst;
    const V cs = __ldg(&table[ti]);
    const uint32_t i1 = i0 | (1u << tq);
    const V a0 = sm[i0], a1 = sm[i1];
    V b0, b1;
    b0.x = cs.x * a0.x - cs.y * a1.x;
    b0.y = cs.x * a0.y - cs.y * a1.y;
    b1.x = cs.y * a0.x + cs.x * a1.x;
    b1.y = cs.y * a0.y + cs.x * a1.y;
    sm[i0] = b0;
    sm[i1] = b1;
  }
}

// ---- dispatcher ------------------------------------------------------------------------------------
// Applies ops[0..n_ops) to the tile; every thread of the block must call it.  gbase = global index
// of tile-local index 0 (external controls and table fields outside the tile are read from it).
// heavy_off (optional): for the k >= 4 dense ops of the pass in order, the element offset of their matrix in c_heavy
// (0xFFFF: not there, use the pool).
template <typename V, int THREADS, bool HEAVY>
__device__ __forceinline__ void apply_ops(V* __restrict__ sm, int T, const QlsOp* __restrict__ ops, int n_ops,
                                          const unsigned char* __restrict__ pool, uint64_t gbase, int tid,
                                          const uint16_t* __restrict__ heavy_off = nullptr) {
  int o = 0;
  int heavy_seen = 0;
  while (o < n_ops) {
    const QlsOp& op = ops[o];
    int n_att = 0;
    while (n_att < QLS_MAX_ATTACH && o + 1 + n_att < n_ops &&
           (ops[o + 1 + n_att].flags & (QLS_OP_FLAG_ATTACH_PRE | QLS_OP_FLAG_ATTACH_POST)))
      ++n_att;
    const int next = o + 1 + n_att;
    // ordinal among the pass's wide dense blocks (counted whether or not the block fires on this tile: the host numbers them all)
    const bool wide = HEAVY && op.kind == QLS_OP_DENSE && !(op.flags & QLS_OP_FLAG_PARAM) && op.k >= 4;
    const int my_heavy = heavy_seen;
    if (wide) ++heavy_seen;
    if ((gbase & op.xc_mask) != op.xc_val) {  // block-uniform
      o = next;
      continue;
    }
    const V* data = reinterpret_cast<const V*>(pool + op.pool_off);
    const bool real = (op.flags & QLS_OP_FLAG_REAL) != 0;
    const Att<V> att = make_att<V>(ops + o + 1, n_att, pool, gbase);
    switch (op.kind) {
      case QLS_OP_DENSE:
        switch (op.k) {
          case 1:
            if (real) op_dense<V, 1, 1, THREADS, true>(sm, T, op, data, tid, att);
            else op_dense<V, 1, 1, THREADS, false>(sm, T, op, data, tid, att);
            break;
          case 2:
            if (real) op_dense<V, 2, 1, THREADS, true>(sm, T, op, data, tid, att);
            else op_dense<V, 2, 1, THREADS, false>(sm, T, op, data, tid, att);
            break;
          case 3:
            op_dense<V, 3, 1, THREADS, false>(sm, T, op, data, tid, att);
            break;
          case 4:
          case 5:
            if (HEAVY) {
              const uint32_t co = (heavy_off && my_heavy < QLS_MAX_HEAVY) ? heavy_off[my_heavy] : 0xFFFFu;
              if constexpr (sizeof(V) == 16) {
                if (co != 0xFFFFu && n_att == 0) {
                  if (op.k == 4) op_dense_cm<4, 2, THREADS>(reinterpret_cast<double2*>(sm), T, op, co, tid);
                  else op_dense_cm<5, 4, THREADS>(reinterpret_cast<double2*>(sm), T, op, co, tid);
                  break;
                }
              }
              if (op.k == 4) op_dense<V, 4, 1, THREADS, false>(sm, T, op, data, tid, att);
              else op_dense<V, 5, 2, THREADS, false>(sm, T, op, data, tid, att);
            }
            break;
        }
        break;
      case QLS_OP_DIAG:
        op_diag<V, THREADS>(sm, T, op, data, gbase, tid, att);
        break;
      case QLS_OP_MUXRY:
        op_muxry<V, THREADS>(sm, T, op, data, gbase, tid);
        break;
    }
    __syncthreads();
    o = next;
  }
}
