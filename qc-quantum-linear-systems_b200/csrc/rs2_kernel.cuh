// Kernel half of a specialised pass unit; included at the END of the generated source (after
// qls_tile_base / qls_run_offset / qls_tile_idle / qls_tile_body).  See rs2_runtime.cuh.
#ifndef QLS_RS2_KERNEL_CUH
#define QLS_RS2_KERNEL_CUH

#define QLS_RUN_BYTES (16u << QLS_L)
#define QLS_TILE_BYTES (16u << QLS_TILE_BITS)

#ifndef QLS_HOST_EMU

// landing offset (bytes) of run r inside a tile buffer
QLS_DEV u32 qls_land_bytes(u32 r) { return (QLS_L < QLS_TILE_BITS ? ((r << QLS_L) + QLS_LAND_PAD(r)) : 0u) * 16u; }

// Block: QLS_GROUPS * 256 consumer threads + one warpgroup whose first warp is the producer (its other three
// warps only give their registers back).  The kernel is compiled for 640 threads (96 registers); the consumer
// warpgroups then raise their budget to QLS_CONSUMER_REGS with setmaxnreg, the producer warpgroup drops to 24.
// setmaxnreg only redistributes the registers the CTA was launched with (640 * 96 = 61440): the increase is
// granted once the producer warpgroup has released enough, so 512 * R + 128 * 24 <= 61440, i.e. R <= 112
// (a larger request spins forever in USETMAXREG.TRY_ALLOC).
#define QLS_BLOCK_THREADS (QLS_GROUPS * QLS_GROUP_THREADS + 128)
#ifndef QLS_CONSUMER_REGS
#define QLS_CONSUMER_REGS 112
#endif
static_assert(QLS_GROUPS * QLS_GROUP_THREADS * QLS_CONSUMER_REGS + 128 * 24 <= QLS_BLOCK_THREADS * 96, "setmaxnreg budget");
#define QLS_STR2(x) #x
#define QLS_STR(x) QLS_STR2(x)

extern "C" __global__ void __launch_bounds__(QLS_BLOCK_THREADS, 1)
    qls_rs2_pass(V* __restrict__ state, const u64 index_offset, const u64 n_tiles, const unsigned char* __restrict__ pool,
                 const __grid_constant__ QlsTensorMap tmap, u64* __restrict__ qls_trace, const u32 sel_pos, const u32 sel_val,
                 const u32 sel_pos2, const u32 sel_val2) {
  // sel_pos < 64: the launch covers a HALF (with sel_pos2 < 64 as well: a QUARTER) of the pass -- the tiles whose ordinal has
  // bit sel_pos equal to sel_val (and bit sel_pos2 > sel_pos equal to sel_val2); n_tiles is already reduced: ordinal t of the
  // launch is ordinal qls_tile_sel(t) of the pass.  A sharded state runs the passes next to an exchange on one part of the
  // shard while another part is on the wire (qls_b200/shard.py).
#define qls_ins_bit(t, pos, val) ((((t) >> (pos)) << ((pos) + 1u)) | ((u64)(val) << (pos)) | ((t) & ((1ull << (pos)) - 1ull)))
  auto qls_tile_sel = [=](u64 t) -> u64 {
    if (sel_pos < 64u) t = qls_ins_bit(t, sel_pos, sel_val);
    if (sel_pos2 < 64u) t = qls_ins_bit(t, sel_pos2, sel_val2);
    return t;
  };
  extern __shared__ unsigned char qls_smem_raw[];
  // 1024-byte alignment: the tensor-map swizzle pattern is a function of the shared-memory address
  unsigned char* const qls_smem = qls_smem_raw + ((1024u - (qls_smem_u32(qls_smem_raw) & 1023u)) & 1023u);
  u64* full = reinterpret_cast<u64*>(qls_smem + QLS_NBUF * QLS_BUF_BYTES);  // [QLS_NBUF] tile has landed (producer -> groups)
  u64* done = full + QLS_NBUF;                                              // [QLS_NBUF] tile is finished (group -> producer)
  volatile u32* ctl = reinterpret_cast<volatile u32*>(done + QLS_NBUF);     // [0] ticket counter, [1 + 2 g + parity] ticket of group g,
                                                                            // [5 + b] ordinal + 1 of the tile buffer b is armed for,
                                                                            // [8 + who] trace cursor (debug builds)
  const u32 warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) {
    for (int b = 0; b < QLS_NBUF; ++b) {
      qls_mbar_init(full + b, 1);
      qls_mbar_init(done + b, 1);
    }
    for (int j = 0; j < 12; ++j) ctl[j] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // The q-th tile of this CTA is tile blockIdx.x + q * gridDim.x; it uses buffer q % QLS_NBUF in round q / QLS_NBUF.
  if (warp >= QLS_GROUPS * (QLS_GROUP_THREADS / 32)) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp != QLS_GROUPS * (QLS_GROUP_THREADS / 32)) return;
    // ===== producer warp: bulk copies in, bulk copies out =====
    const u32 lane = threadIdx.x & 31u;
#ifdef QLS_TRACE
#define QLS_PTR(ev, qq) do { if (lane == 0) qls_trace_put(qls_trace, ctl + 8, 2u, (ev), (u32)(qq)); } while (0)
#else
#define QLS_PTR(ev, qq) do { } while (0)
#endif
    auto retire = [&](u64 qq) {  // tile qq is finished: write it back and free its buffer
      const u32 b = (u32)(qq % QLS_NBUF);
      QLS_PTR(20, qq);
      qls_mbar_wait(done + b, (u32)(qq / QLS_NBUF) & 1u);
      QLS_PTR(21, qq);
      const u64 base = qls_tile_base(qls_tile_sel((u64)blockIdx.x + qq * gridDim.x));
      if (!qls_tile_idle(base | index_offset)) {
        unsigned char* buf = qls_smem + b * QLS_BUF_BYTES;
#ifdef QLS_EXP_NOLOAD  // timing experiment: no global traffic at all
#elif defined(QLS_TMA)
        if (lane < QLS_TMA_OPS)
          qls_tma_store(&tmap, (u32)((base | qls_tma_split_offset(lane)) >> 3), buf + lane * (QLS_TILE_BYTES / QLS_TMA_OPS));
#else
        for (u32 r = lane; r < QLS_NRUNS; r += 32u)
          qls_bulk_store(state + (base | qls_run_offset(r)), buf + qls_land_bytes(r), QLS_RUN_BYTES);
#endif
        qls_bulk_commit();
        qls_bulk_wait_read();  // the buffer may be overwritten once the copy engine has read it
      }
      __syncwarp();
      QLS_PTR(22, qq);
    };
    u64 q = 0;
    for (;; ++q) {
      const u64 tile = (u64)blockIdx.x + q * gridDim.x;
      if (tile >= n_tiles) break;
      if (q >= QLS_NBUF) retire(q - QLS_NBUF);
      const u32 b = (u32)(q % QLS_NBUF);
      const u64 base = qls_tile_base(qls_tile_sel(tile));
      const bool idle = qls_tile_idle(base | index_offset);
      if (lane == 0) {
        // Publish which tile this round of the buffer belongs to BEFORE arming it.  A parity wait cannot tell round
        // R + 2 from round R: a group that runs two tiles ahead of a load still in flight (tiles land out of order)
        // would otherwise take the buffer of a tile that has not landed.
        ctl[5 + b] = (u32)q + 1u;
        __threadfence_block();
#ifdef QLS_EXP_NOLOAD
        qls_mbar_arrive(full + b);
#else
        if (idle) qls_mbar_arrive(full + b);  // nothing to move: the group only passes the tile through
        else qls_mbar_arrive_expect_tx(full + b, QLS_TILE_BYTES);
#endif
      }
      __syncwarp();
      if (!idle) {
        unsigned char* buf = qls_smem + b * QLS_BUF_BYTES;
#ifdef QLS_EXP_NOLOAD
#elif defined(QLS_TMA)
        if (lane < QLS_TMA_OPS)
          qls_tma_load(buf + lane * (QLS_TILE_BYTES / QLS_TMA_OPS), &tmap, (u32)((base | qls_tma_split_offset(lane)) >> 3), full + b);
#else
        for (u32 r = lane; r < QLS_NRUNS; r += 32u)
          qls_bulk_load(buf + qls_land_bytes(r), state + (base | qls_run_offset(r)), QLS_RUN_BYTES, full + b);
#endif
      }
      QLS_PTR(23, q);
    }
    for (u64 qq = q >= QLS_NBUF ? q - QLS_NBUF : 0; qq < q; ++qq) retire(qq);
    qls_bulk_wait_all();
  } else {
    // ===== consumer groups =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 " QLS_STR(QLS_CONSUMER_REGS) ";");
    const u32 qls_group = warp / (QLS_GROUP_THREADS / 32);
    const u32 tid = threadIdx.x & (QLS_GROUP_THREADS - 1);
#ifdef QLS_EXP_ONEGROUP  // timing experiment: the second group does nothing
    if (qls_group == 1) return;
#endif
    // Tickets: thread 0 of a group draws the ordinal of the group's NEXT tile before the barrier that ends the current
    // one, so that barrier publishes it (no barrier of its own).  Two slots per group, alternating: a warp that is
    // slow to read the current ordinal is at most one barrier behind the writer.
    if (tid == 0) ctl[1 + 2 * qls_group] = atomicAdd(const_cast<u32*>(&ctl[0]), 1u);
    QLS_BAR();
    for (u32 it = 0;; ++it) {
      // (the shuffle makes the ordinal provably warp-uniform: tile base and control predicates then live in uniform
      // registers / uniform predicates instead of being spilled as per-thread booleans)
      const u32 q = __shfl_sync(0xffffffffu, ctl[1 + 2 * qls_group + (it & 1u)], 0);
      const u32 qls_q = q;
      volatile u32* const qls_cursor = ctl + 8;
      (void)qls_cursor;
      const u64 tile = (u64)blockIdx.x + (u64)q * gridDim.x;
      if (tile >= n_tiles) break;
      const u32 b = q % QLS_NBUF;
      QLS_TR(1);
      {  // the buffer must be armed for THIS tile (see the producer) before the parity of its barrier means anything
        u32 spins = 0;
        while (ctl[5 + b] != q + 1u) {
          if (++spins > (1u << 26)) __trap();
        }
      }
      qls_mbar_wait(full + b, (q / QLS_NBUF) & 1u);
      QLS_TR(2);
      const u64 gbase = qls_tile_base(qls_tile_sel(tile)) | index_offset;
      if (!qls_tile_idle(gbase)) {
        qls_tile_body(reinterpret_cast<V*>(qls_smem + b * QLS_BUF_BYTES), gbase, tid, qls_group, pool, qls_trace, qls_q, qls_cursor);
        qls_fence_proxy_async();  // the bulk store reads this group's shared-memory writes through the async proxy
      }
      QLS_TR(3);
      if (tid == 0) ctl[1 + 2 * qls_group + ((it + 1u) & 1u)] = atomicAdd(const_cast<u32*>(&ctl[0]), 1u);
      QLS_BAR();
      if (tid == 0) qls_mbar_arrive(done + b);
      QLS_TR(4);
    }
  }
}

#else  // QLS_HOST_EMU: one group, real host threads, the copies done by thread 0

extern "C" void qls_emu_run_pass(V* state, u64 index_offset, u64 n_tiles, const unsigned char* pool) {
  static V buf[QLS_BUF_UNITS];
  qls_emu_run(
      n_tiles,
      [&](u64 tile, bool store) {
        const u64 base = qls_tile_base(tile);
        if (qls_tile_idle(base | index_offset)) return;
        for (u32 i = 0; i < QLS_TILE_AMPS; ++i) {
          V* g = state + (base | qls_run_offset(i >> QLS_L) | (u64)(i & ((1u << QLS_L) - 1u)));
          if (store) *g = buf[qls_land_unit(i)];
          else buf[qls_land_unit(i)] = *g;
        }
      },
      [&](u64 tile, u32 tid) {
        const u64 gbase = qls_tile_base(tile) | index_offset;
        if (!qls_tile_idle(gbase)) qls_tile_body(buf, gbase, tid, 0u, pool, nullptr, 0u, nullptr);
      });
}

#endif
#endif  // QLS_RS2_KERNEL_CUH
