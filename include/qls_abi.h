/*
 * qls_abi.h -- C ABI of the B200 statevector engine (libqls_b200.so).
 *
 * This is the drop-in boundary for the one hot path of SURFQuantum/qc-quantum-linear-systems:
 * applying a gate stream to a complex statevector and reducing the result.  In the reference
 * that path is reached through Python objects of third-party packages, so each entry point
 * cites the reference call site (path relative to the reference tree) whose work it performs:
 *
 *   Statevector(hhl_circuit).data      quantum_linear_systems/implementations/hhl_qiskit_implementation.py:49
 *   Statevector(res.state).data        quantum_linear_systems/implementations/vqls_qiskit_implementation.py:65
 *   HHL.solve -> _calculate_norm       quantum_linear_systems/implementations/hhl_qiskit_implementation.py:35-38,54
 *   estimator.run(...).result().values quantum_linear_systems/implementations/vqls_qiskit_implementation.py:55-62
 *   extract_hhl_solution_vector_...    quantum_linear_systems/utils.py:77-94
 *
 * Conventions: plain pointers and sizes only.  `state`, `scratch`, `workspace`, `*_dev` are
 * DEVICE pointers owned by the caller (the Python host allocates them as torch tensors);
 * `*_host` are host pointers.  `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  Amplitude index bit q = qubit q (little endian).  dtype: QLS_C128 = interleaved
 * (re, im) doubles, QLS_C64 = interleaved floats.  Every function returns QLS_OK (0) or an error
 * code; qls_last_error() gives the text for the calling thread.  Calls are asynchronous on
 * `stream` unless stated otherwise.
 */
#ifndef QLS_ABI_H
#define QLS_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QLS_ABI_VERSION 1

enum { QLS_C128 = 0, QLS_C64 = 1 };
enum { QLS_OK = 0, QLS_ERR_ARG = 1, QLS_ERR_CUDA = 2, QLS_ERR_UNSUPPORTED = 3, QLS_ERR_NOMEM = 4 };

/* ---- fused-block program ------------------------------------------------------------------ */

enum {
  QLS_OP_DENSE = 1, /* 2^k x 2^k complex matrix on k tile qubits (k <= 5), optional controls   */
  QLS_OP_DIAG = 2,  /* merged diagonal run: amp *= table[index bits]                          */
  QLS_OP_MUXRY = 3, /* uniformly controlled RY: (cos, sin) table selected by index bits       */
  QLS_OP_MATVEC = 4 /* batch kernel only: dense 2^k x 2^k on k contiguous tile qubits tq[0]..   */
};

#define QLS_MAX_SEG 6
#define QLS_MAX_HIGH 8
#define QLS_MAX_FSEG 16
#define QLS_OP_FLAG_REAL 1u  /* DENSE matrix is real (imaginary parts are zero) */
#define QLS_OP_FLAG_PARAM 2u /* batch kernel only: a 1-qubit rotation whose angle comes from the parameter
                              * array; pad0 selects the gate: 0 ry, 1 rx, 2 rz, 3 p */
/* A DIAG record flagged ATTACH_* is not an operation of its own: it rides on the closest preceding
 * un-flagged record (its host: an unconditional DENSE, or a DIAG) and is applied to the host's
 * inputs (PRE) or outputs (POST) while they are in registers -- no extra shared-memory sweep. */
#define QLS_OP_FLAG_ATTACH_PRE 4u
#define QLS_OP_FLAG_ATTACH_POST 8u
#define QLS_MAX_ATTACH 4

/* bit-field gather: ((x >> src) & ((1 << width) - 1)) << dst */
typedef struct QlsSeg {
  uint8_t src, width, dst, pad;
} QlsSeg;

/* One operation applied to a shared-memory tile of 2^tile_bits amplitudes. 128 bytes. */
typedef struct QlsOp {
  int32_t kind;       /* QLS_OP_*                                                             */
  int32_t k;          /* DENSE: number of target qubits; MUXRY: 1; DIAG: unused               */
  uint8_t tq[8];      /* tile-local bit positions of the targets, ascending; matrix bit j <-> tq[j] */
  uint32_t lc_mask;   /* controls that are tile qubits: (local_index & lc_mask) == lc_val     */
  uint32_t lc_val;
  uint64_t xc_mask;   /* controls outside the tile, tested on the tile's global base index    */
  uint64_t xc_val;
  uint64_t pool_off;  /* byte offset of the matrix / table in the constant pool               */
  uint8_t n_lseg;     /* DIAG/MUXRY: table index fields taken from the tile-local index       */
  uint8_t n_xseg;     /*             ... and from the tile's global base index                */
  uint8_t flags;
  uint8_t pad0;       /* QLS_OP_FLAG_PARAM: rotation kind                                     */
  uint32_t pad1;
  QlsSeg lseg[QLS_MAX_SEG];
  QlsSeg xseg[QLS_MAX_SEG];
  uint32_t param_index; /* batch kernel: angle = params[circuit][param_index] when QLS_OP_FLAG_PARAM */
  float param_coeff;    /*               angle = coeff * param + offset                              */
  float param_offset;
  uint8_t reserved[12];
} QlsOp;

/* One pass over the state: every tile (2^tile_bits amplitudes: the low `low_bits` qubits plus
 * `n_high` further qubits) is staged in shared memory, ops [op_begin, op_begin+op_count) are
 * applied, and the tile is written back: one HBM read + one write per touched amplitude. */
typedef struct QlsPass {
  int32_t tile_bits;
  int32_t low_bits;
  int32_t n_high;
  uint8_t high[QLS_MAX_HIGH]; /* qubit positions of tile-local bits low_bits.., ascending     */
  uint32_t op_begin;
  uint32_t op_count;
  uint64_t fix_mask;          /* pass-level controls (LOCAL index bits outside the tile):     */
  uint64_t fix_val;           /* tiles with (index & fix_mask) != fix_val are not touched     */
  int32_t n_fseg;             /* deposit of the tile counter into the free index bits         */
  QlsSeg fseg[QLS_MAX_FSEG];
  uint32_t pad;
  uint32_t rs_begin;          /* register-stage form: records [rs_begin, rs_begin + rs_count) of   */
  uint32_t rs_count;          /* the program's QlsRsOp array; rs_count == 0 -> sweep form only     */
} QlsPass;

/* ---- register-stage form of a pass (the fast path for full-size tiles) -----------------------
 * The 2^tile_bits amplitudes of a tile are held in the registers of the CTA's 256 threads
 * (2^reg_bits each).  A pass is a sequence of STAGES; a stage fixes which reg_bits tile qubits are
 * "register qubits" (the other 8 index the thread) and applies its micro-ops entirely in
 * registers.  Between stages the tile is exchanged through shared memory (one write + one read).
 * Only 1- and 2-qubit dense blocks, diagonal tables and multiplexed RY are expressible; passes with
 * wider dense blocks keep the shared-memory sweep form above. */
enum {
  QLS_RS_STAGE = 1, /* tseg: thread index -> tile-local index deposit; rt[r]: swizzled offset of register combo r */
  QLS_RS_G1 = 2,    /* 2x2 on register qubit r[0]                                                  */
  QLS_RS_G2 = 3,    /* 4x4 on register qubits r[0] < r[1] (matrix bit 0 <-> r[0])                  */
  QLS_RS_DIAG = 4,  /* amp *= table[xseg(global base) | tseg(thread index) | rt[register combo]]    */
  QLS_RS_MUXRY = 5  /* RY on register qubit r[0], (cos, sin) from the same kind of table index     */
};

typedef struct QlsRsOp { /* 192 bytes */
  int32_t kind;
  uint8_t r[4];        /* target register-qubit indexes                                           */
  uint32_t ctl_tmask;  /* controls among the thread-index bits ...                                */
  uint32_t ctl_tval;
  uint32_t ctl_rmask;  /* ... and among the register-combo bits                                   */
  uint32_t ctl_rval;
  uint64_t xc_mask;    /* controls outside the tile (tested on the tile's global base index)      */
  uint64_t xc_val;
  uint64_t pool_off;
  uint8_t n_tseg, n_xseg, flags, pad0;
  uint32_t pad1;
  QlsSeg tseg[QLS_MAX_SEG];
  QlsSeg xseg[QLS_MAX_SEG];
  uint16_t rt[32];
  uint8_t reserved[24];
} QlsRsOp;

typedef struct QlsProgram QlsProgram; /* opaque: device copies of ops + constant pool */

int qls_abi_version(void);
const char* qls_last_error(void);
/* sm count, opt-in shared memory per block, total/free device memory of `device` */
int qls_device_info(int device, int* sm_count, uint64_t* smem_optin, uint64_t* mem_total, uint64_t* mem_free);

/* Upload a lowered program (passes/ops/pool are HOST arrays; pool entries are in `dtype`).
 * Synchronous.  Replaces the per-gate loop of Statevector._evolve_instruction [hhl..:49]. */
int qls_program_create(const QlsPass* passes_host, uint32_t n_passes, const QlsOp* ops_host, uint32_t n_ops,
                       const QlsRsOp* rs_ops_host, uint32_t n_rs_ops, const void* pool_host, uint64_t pool_bytes,
                       int dtype, int device, QlsProgram** out);
/* 0: use the register-stage form where a pass has one (default); 1: always use the sweep form. */
int qls_set_sweep_only(int on);
int qls_program_destroy(QlsProgram* prog);
/* Run passes [pass_begin, pass_end) on a state of 2^n_local amplitudes.  `index_offset` is the
 * global index of local amplitude 0 (rank << n_local for a sharded state, else 0). */
int qls_program_run(const QlsProgram* prog, void* state, int n_local, uint64_t index_offset, uint32_t pass_begin,
                    uint32_t pass_end, void* stream);

/* ---- specialised pass kernels ----------------------------------------------------------------
 * A pass with a register-stage form can be given a kernel compiled for exactly that pass (stage layouts,
 * register indices, table-index fields and gate matrices as compile-time constants; qls_b200/jit.py
 * writes the source, csrc/rs2_runtime.cuh + rs2_kernel.cuh are the skeleton: a producer warp moving
 * tiles with bulk asynchronous copies, two consumer groups, three shared-memory tile buffers).
 * qls_program_run launches it instead of the interpreter kernel.  [same call site, hhl..:49] */

/* CUDA C++ -> sm_100a cubin with NVRTC (no GPU needed).  `header_names[i]` is what the source #includes,
 * `header_sources[i]` its text.  *cubin_out is malloc'ed: release it with qls_jit_free.  The tail of the
 * compiler log is copied to `log` (may be NULL). */
int qls_jit_compile(const char* source, const char* source_name /* file name in the line table; may be NULL */,
                    const char* const* header_names, const char* const* header_sources, int n_headers,
                    const char* const* options, int n_options, void** cubin_out, uint64_t* cubin_bytes, char* log,
                    uint64_t log_cap);
int qls_jit_free(void* cubin);
/* Attach `kernel_name` of the cubin to pass `pass_index` (block of `threads` threads, `smem_bytes` of dynamic
 * shared memory; signature (state, index_offset, n_tiles, pool)). */
int qls_program_set_pass_kernel(QlsProgram* prog, uint32_t pass_index, const void* cubin, uint64_t cubin_bytes,
                                const char* kernel_name, uint32_t threads, uint32_t smem_bytes);
/* The attached kernel lands its tiles with tensor-map copies (cp.async.bulk.tensor): describe the tile of the pass.
 * Tile qubits [0, low_bits) are rows of 128 bytes (box dimensions 0 and 1); each of the n_groups <= 3 groups is
 * group_bits[g] tile qubits at consecutive index bits from group_pos[g] (one box dimension each).  The map itself
 * is encoded per launch (it holds the state pointer).  Without this call the kernel is expected to use per-run
 * bulk copies. */
int qls_program_set_pass_tma(QlsProgram* prog, uint32_t pass_index, int low_bits, int n_groups, const int* group_pos,
                             const int* group_bits, int n_local);
/* Debug: device buffer (>= 16 + 16 * 8192 bytes, zeroed) that kernels compiled with -DQLS_TRACE fill with the
 * timeline of CTA 0 (tools/jit_trace.py); NULL (default) switches it off. */
int qls_set_trace_buffer(void* trace_dev);
/* 1 (default): use attached specialised kernels; 0: always the interpreter kernels (A/B tests). */
int qls_set_use_jit(int on);

/* ---- state initialisation and host <-> device access ------------------------------------- */

/* psi = 0; psi[basis_index] = 1 (pass UINT64_MAX to only zero).  [Statevector.from_instruction] */
int qls_sv_init_basis(void* state, int n_local, int dtype, uint64_t basis_index, void* stream);
/* copy `count` amplitudes host -> state[offset..) / state[offset..) -> host (same dtype);
 * synchronous with respect to the host buffer.  [state preparation a7 / `.data` hhl..:49] */
int qls_sv_write(void* state, int dtype, uint64_t offset, const void* amps_host, uint64_t count, void* stream);
int qls_sv_read(const void* state, int dtype, uint64_t offset, void* amps_host, uint64_t count, void* stream);
/* psi *= (re + i*im)  -- a definition's global phase */
int qls_sv_scale(void* state, int n_local, int dtype, double re, double im, void* stream);

/* ---- reductions (synchronous: result lands in out_host) ----------------------------------- */

enum {
  QLS_RED_NORM_RANGE = 0, /* a = offset, b = count : sum |psi_i|^2 over the range  [HHL._calculate_norm, utils.py:94] */
  QLS_RED_PROB_MASK = 1,  /* a = mask,  b = value : sum |psi_i|^2 over ((index_offset|i) & mask) == value          */
  QLS_RED_EXPECT_Z = 2,   /* a = qubit            : sum (-1)^{bit_a} |psi_i|^2   [Statevector.expectation_value]   */
  QLS_RED_INNER = 3       /* <other|psi> -> out[0] + i*out[1]  (fidelity)                                            */
};
#define QLS_REDUCE_WORKSPACE_BYTES (64u * 1024u)
int qls_reduce(const void* state, int n_local, int dtype, int kind, uint64_t a, uint64_t b, uint64_t index_offset,
               const void* other, void* workspace, double* out_host /* [2] */, void* stream);

/* ---- dense controlled-e^{iAt} block as a complex GEMM on the FP64 tensor pipe -------------- */

/* rows of psi.reshape(-1, 2^nb) whose global row index satisfies (row_index & ctrl_mask) ==
 * ctrl_val are replaced by row @ U^T (U row-major 2^nb x 2^nb, device, in `dtype`).
 *   scratch != NULL: out of place in chunks of QLS_GEMM_CHUNK_ROWS selected rows + copy-back; `scratch` holds
 *                    min(selected rows, QLS_GEMM_CHUNK_ROWS) * 2^nb amplitudes (256 MiB at nb = 10, whatever the state size);
 *   scratch == NULL: in place (complex128, nb <= 10): no second buffer at all, ~25 % slower (see csrc/ctrl_gemm.cu).
 * [PhaseEstimation's controlled-U^(2^j) blocks inside HHL.construct_circuit, hhl..:35-38] */
#define QLS_GEMM_CHUNK_ROWS (1u << 14)
int qls_ctrl_gemm(void* state, void* scratch, int n_local, int dtype, int nb, uint64_t ctrl_mask, uint64_t ctrl_val,
                  uint64_t index_offset, const void* u_dev, void* stream);

/* ---- batch of small circuits (VQLS Hadamard tests) ----------------------------------------- */

/* K circuits of n_qubits (<= 12) sharing one op program; ops flagged as parameterised take their
 * angle from params_dev[circuit*P + index].  out_dev[circuit] = <Z_qubit>.  Programs with a dense operator on 6..10
 * qubits keep a launch-wide state array (<= 512 MiB) inside the program object: launches of ONE program must be stream
 * ordered with one another (different programs are independent).
 * [Estimator.run(circuits, observables, parameter_values), vqls..:55-62] */
int qls_batch_expect_z(const QlsProgram* prog, uint32_t pass_index, int n_qubits, int z_qubit, const double* params_dev,
                       uint32_t n_circuits, uint32_t n_params, double* out_dev, void* stream);

/* ---- global-qubit remap helpers (sharded states) -------------------------------------------- */

/* Gather / scatter the half-open slab of amplitudes whose local bit `qubit` equals `bit_value`
 * into / from a contiguous buffer (count = 2^(n_local-1)); used to stage NCCL send/recv. */
int qls_pack_bit(const void* state, void* buffer, int n_local, int dtype, int qubit, int bit_value, uint64_t chunk_begin,
                 uint64_t chunk_count, void* stream);
int qls_unpack_bit(void* state, const void* buffer, int n_local, int dtype, int qubit, int bit_value, uint64_t chunk_begin,
                   uint64_t chunk_count, void* stream);
/* The same for k local bits at once (qubits_host: ascending positions; bit m of value_bits is the value
 * of qubits_host[m]; count = 2^(n_local-k) per sub-block): a k-qubit remap exchanges 2^k - 1 such
 * sub-blocks, one with every rank that differs in the k global bits, instead of k half-shard swaps. */
#define QLS_MAX_REMAP_BITS 6
int qls_pack_bits(const void* state, void* buffer, int n_local, int dtype, const int* qubits_host, int k, uint64_t value_bits,
                  uint64_t chunk_begin, uint64_t chunk_count, void* stream);
int qls_unpack_bits(void* state, const void* buffer, int n_local, int dtype, const int* qubits_host, int k, uint64_t value_bits,
                    uint64_t chunk_begin, uint64_t chunk_count, void* stream);
/* One pairwise round without staging: swap elements [begin, begin+count) of my sub-block (local bits =
 * my_value_bits) with the same elements of the partner's sub-block (local bits = peer_value_bits) in
 * `peer_state`, the partner's state buffer mapped into this process (CUDA IPC; peer access over NVLink).
 * The two GPUs of a pair must be given disjoint element ranges and a device-side barrier before and after. */
/* Map another process's allocation (64-byte cudaIpcMemHandle_t) for kernels of `device`; *out = its base. */
int qls_ipc_open(const void* handle64, int device, void** out);
int qls_ipc_close(void* ptr);
/* cudaDeviceEnablePeerAccess(peer) in the context of `device` (idempotent). */
int qls_enable_peer_access(int device, int peer);
int qls_swap_peer(void* state, void* peer_state, int n_local, int dtype, const int* qubits_host, int k,
                  uint64_t my_value_bits, uint64_t peer_value_bits, uint64_t begin, uint64_t count, void* stream);
/* Stream-ordered copy between any two device buffers visible to this process, peer-mapped ones included (cudaMemcpyAsync,
 * i.e. the copy engines: a chunk of a remap travels to the partner's receive slot without occupying SMs).  Remap transport of
 * qls_b200/shard.py; no reference counterpart (the reference is single-process: hhl_qiskit_implementation.py:49). */
int qls_copy_async(void* dst, const void* src, uint64_t bytes, void* stream);
/* Pitched variant (cudaMemcpy2DAsync; one copy per row where a pitch exceeds its 2 GiB limit): the sub-block of an exchange whose
 * local positions are consecutive is rows of 2^p amplitudes at one pitch -- the copy engines gather it into the partner's receive
 * slot and scatter the received one, no kernel involved. */
int qls_copy2d_async(void* dst, uint64_t dst_pitch, const void* src, uint64_t src_pitch, uint64_t width_bytes, uint64_t height,
                     void* stream);
/* One pass of a program on the half of the state whose index bit `sel_pos` is `sel_val` -- with `sel_pos2` >= 0 on the quarter
 * whose bit `sel_pos2` is `sel_val2` as well -- (specialised kernels only, else QLS_ERR_UNSUPPORTED), and the number of SMs the specialised pass kernels leave free: the pieces with which qls_b200/shard.py
 * overlaps an exchange with the local passes on both sides of it.  Same reference call as qls_program_run
 * (hhl_qiskit_implementation.py:49). */
int qls_program_run_part(const QlsProgram* prog, void* state, int n_local, uint64_t index_offset, uint32_t pass_index, int sel_pos,
                         int sel_val, int sel_pos2, int sel_val2, void* stream);
int qls_set_pass_grid_reserve(int n_sms);

#ifdef __cplusplus
}
#endif
#endif /* QLS_ABI_H */
