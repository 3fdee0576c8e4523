// Specialised pass kernels: NVRTC compilation (qls_jit_compile) and attachment to a program
// (qls_program_set_pass_kernel).  The source of a pass comes from qls_b200/jit.py; the kernel skeleton
// it includes is csrc/rs2_runtime.cuh + csrc/rs2_kernel.cuh.
//
// libnvrtc and the driver library are opened with dlopen at first use, so that libqls_b200.so itself
// only depends on the CUDA runtime (and loads in the GPU-less build container, where NVRTC still
// compiles for sm_100a).
#include <cuda.h>  // types and enums only: the driver entry points are resolved with dlsym
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "qls_common.cuh"

namespace {

// ---- NVRTC (a handful of entry points, resolved at run time) ------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* lib = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

Nvrtc* nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("QLS_NVRTC_LIB"), "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (n.lib) break;
    }
    if (!n.lib) return;
#define QLS_SYM(field, sym) n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.lib, sym))
    QLS_SYM(CreateProgram, "nvrtcCreateProgram");
    QLS_SYM(DestroyProgram, "nvrtcDestroyProgram");
    QLS_SYM(CompileProgram, "nvrtcCompileProgram");
    QLS_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    QLS_SYM(GetCUBIN, "nvrtcGetCUBIN");
    QLS_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    QLS_SYM(GetProgramLog, "nvrtcGetProgramLog");
    QLS_SYM(GetErrorString, "nvrtcGetErrorString");
#undef QLS_SYM
  });
  const bool ok = n.lib && n.CreateProgram && n.DestroyProgram && n.CompileProgram && n.GetCUBINSize && n.GetCUBIN &&
                  n.GetProgramLogSize && n.GetProgramLog;
  return ok ? &n : nullptr;
}

// ---- driver API (module load + launch of a cubin) -----------------------------------------------------------
struct Driver {
  void* lib = nullptr;
  int (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  int (*ModuleUnload)(CUmodule) = nullptr;
  int (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  int (*FuncSetAttribute)(CUfunction, int, int) = nullptr;
  int (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
  int (*GetErrorString)(int, const char**) = nullptr;
  int (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
};

Driver* driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!d.lib) return;
#define QLS_SYM(field, sym) d.field = reinterpret_cast<decltype(d.field)>(dlsym(d.lib, sym))
    QLS_SYM(ModuleLoadData, "cuModuleLoadData");
    QLS_SYM(ModuleUnload, "cuModuleUnload");
    QLS_SYM(ModuleGetFunction, "cuModuleGetFunction");
    QLS_SYM(FuncSetAttribute, "cuFuncSetAttribute");
    QLS_SYM(LaunchKernel, "cuLaunchKernel");
    QLS_SYM(GetErrorString, "cuGetErrorString");
    QLS_SYM(TensorMapEncodeTiled, "cuTensorMapEncodeTiled");
#undef QLS_SYM
  });
  const bool ok = d.lib && d.ModuleLoadData && d.ModuleUnload && d.ModuleGetFunction && d.FuncSetAttribute && d.LaunchKernel;
  return ok ? &d : nullptr;
}

const char* cu_err(Driver* d, int rc) {
  const char* s = nullptr;
  if (d->GetErrorString) d->GetErrorString(rc, &s);
  return s ? s : "?";
}

}  // namespace

struct QlsJitKernel {
  CUmodule mod;
  CUfunction fn;
  uint32_t threads, smem;
  // tensor map of the pass's tile (0 groups and low_bits 0: the kernel lands tiles with per-run bulk copies)
  int tma_low_bits;
  int tma_groups;
  int tma_pos[3], tma_bits[3];
};

// Tensor map of one pass over a state of 2^n_local complex128 amplitudes, viewed as float64:
//   dim 0   16 doubles  = 8 amplitudes = one 128-byte row (the swizzle span)
//   dim 1   rows of the whole state (stride 128 bytes); the box takes 2^(low_bits - 3) of them and coordinate 1
//           carries the tile's base index, so one map serves every tile of the pass
//   dim 2.. one per group of tile qubits above the low run: 2^bits entries, stride 16 << (index bit) bytes
static int qls_encode_tile_map(Driver* d, const QlsJitKernel* k, void* state, int n_local, CUtensorMap* out) {
  if (!d->TensorMapEncodeTiled) {
    qls_set_error("cuTensorMapEncodeTiled is not available in this driver");
    return QLS_ERR_UNSUPPORTED;
  }
  cuuint64_t dims[5] = {16, 1ull << (n_local - 3), 1, 1, 1};
  cuuint64_t strides[4] = {128, 128, 128, 128};
  cuuint32_t box[5] = {16, 1u << (k->tma_low_bits - 3), 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int g = 0; g < k->tma_groups; ++g) {
    dims[2 + g] = 1ull << k->tma_bits[g];
    box[2 + g] = 1u << k->tma_bits[g];
    strides[1 + g] = 16ull << k->tma_pos[g];
  }
  const int rc = d->TensorMapEncodeTiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, state, dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc) {
    qls_set_error("cuTensorMapEncodeTiled: %s (n_local %d, low bits %d, %d groups)", cu_err(d, rc), n_local, k->tma_low_bits,
                  k->tma_groups);
    return QLS_ERR_CUDA;
  }
  return QLS_OK;
}

extern "C" int qls_jit_compile(const char* source, const char* source_name, const char* const* header_names,
                               const char* const* header_sources, int n_headers, const char* const* options, int n_options,
                               void** cubin_out, uint64_t* cubin_bytes, char* log, uint64_t log_cap) {
  QLS_ARG(source && cubin_out && cubin_bytes, "null pointer");
  if (log && log_cap) log[0] = 0;
  Nvrtc* n = nvrtc();
  if (!n) {
    qls_set_error("NVRTC (libnvrtc.so.12) could not be loaded: %s", dlerror());
    return QLS_ERR_UNSUPPORTED;
  }
  nvrtcProgram prog = nullptr;
  int rc = n->CreateProgram(&prog, source, source_name && *source_name ? source_name : "qls_pass.cu", n_headers, header_sources, header_names);
  if (rc) {
    qls_set_error("nvrtcCreateProgram: %s", n->GetErrorString ? n->GetErrorString(rc) : "?");
    return QLS_ERR_CUDA;
  }
  rc = n->CompileProgram(prog, n_options, options);
  size_t ls = 0;
  if (log && log_cap && n->GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
    std::vector<char> buf(ls + 1, 0);
    n->GetProgramLog(prog, buf.data());
    const size_t take = ls < log_cap - 1 ? ls : log_cap - 1;
    memcpy(log, buf.data() + (ls - take), take);  // keep the tail: the errors are at the end
    log[take] = 0;
  }
  if (rc) {
    qls_set_error("nvrtcCompileProgram: %s", n->GetErrorString ? n->GetErrorString(rc) : "?");
    n->DestroyProgram(&prog);
    return QLS_ERR_CUDA;
  }
  size_t sz = 0;
  rc = n->GetCUBINSize(prog, &sz);
  if (rc || sz == 0) {
    qls_set_error("nvrtcGetCUBINSize: %s (size %zu) -- a real architecture (sm_100a) must be requested", n->GetErrorString ? n->GetErrorString(rc) : "?", sz);
    n->DestroyProgram(&prog);
    return QLS_ERR_CUDA;
  }
  char* out = static_cast<char*>(malloc(sz));
  if (!out) {
    n->DestroyProgram(&prog);
    qls_set_error("out of host memory");
    return QLS_ERR_NOMEM;
  }
  rc = n->GetCUBIN(prog, out);
  n->DestroyProgram(&prog);
  if (rc) {
    free(out);
    qls_set_error("nvrtcGetCUBIN failed");
    return QLS_ERR_CUDA;
  }
  *cubin_out = out;
  *cubin_bytes = sz;
  return QLS_OK;
}

extern "C" int qls_jit_free(void* cubin) {
  free(cubin);
  return QLS_OK;
}

extern "C" int qls_program_set_pass_kernel(QlsProgram* prog, uint32_t pass_index, const void* cubin, uint64_t cubin_bytes,
                                           const char* kernel_name, uint32_t threads, uint32_t smem_bytes) {
  QLS_ARG(prog && cubin && cubin_bytes && kernel_name, "null pointer");
  QLS_ARG(pass_index < prog->n_passes, "pass %u outside program of %u", pass_index, prog->n_passes);
  QLS_ARG(threads >= 32 && threads <= 1024 && smem_bytes <= 227u * 1024u, "bad launch shape (%u threads, %u bytes)", threads, smem_bytes);
  Driver* d = driver();
  if (!d) {
    qls_set_error("the CUDA driver library (libcuda.so.1) could not be loaded");
    return QLS_ERR_UNSUPPORTED;
  }
  QLS_CUDA(cudaSetDevice(prog->device));
  QLS_CUDA(cudaFree(0));  // make sure the primary context exists and is current
  if (!prog->jit) {
    prog->jit = new QlsJitKernel*[prog->n_passes];
    for (uint32_t i = 0; i < prog->n_passes; ++i) prog->jit[i] = nullptr;
  }
  QlsJitKernel* k = new QlsJitKernel();
  memset(k, 0, sizeof(*k));
  k->threads = threads;
  k->smem = smem_bytes;
  int rc = d->ModuleLoadData(&k->mod, cubin);
  if (rc) {
    qls_set_error("cuModuleLoadData: %s", cu_err(d, rc));
    delete k;
    return QLS_ERR_CUDA;
  }
  rc = d->ModuleGetFunction(&k->fn, k->mod, kernel_name);
  if (!rc) rc = d->FuncSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_bytes);
  if (rc) {
    qls_set_error("kernel %s: %s", kernel_name, cu_err(d, rc));
    d->ModuleUnload(k->mod);
    delete k;
    return QLS_ERR_CUDA;
  }
  if (prog->jit[pass_index]) {
    d->ModuleUnload(prog->jit[pass_index]->mod);
    delete prog->jit[pass_index];
  }
  prog->jit[pass_index] = k;
  return QLS_OK;
}

extern "C" int qls_program_set_pass_tma(QlsProgram* prog, uint32_t pass_index, int low_bits, int n_groups, const int* group_pos,
                                        const int* group_bits, int n_local) {
  QLS_ARG(prog && prog->jit && pass_index < prog->n_passes && prog->jit[pass_index], "no specialised kernel attached to pass %u", pass_index);
  QLS_ARG(low_bits >= 4 && low_bits <= 11 && n_groups >= 0 && n_groups <= 3 && (n_groups == 0 || (group_pos && group_bits)),
          "bad tensor-map description (low bits %d, %d groups)", low_bits, n_groups);
  QLS_ARG(n_local >= 12 && n_local <= 35, "tensor map: n_local %d out of range", n_local);
  QlsJitKernel* k = prog->jit[pass_index];
  k->tma_low_bits = low_bits;
  k->tma_groups = n_groups;
  for (int g = 0; g < n_groups; ++g) {
    QLS_ARG(group_bits[g] >= 1 && group_bits[g] <= 8 && group_pos[g] >= low_bits && group_pos[g] + group_bits[g] <= n_local,
            "tensor-map group %d: bits %d at index bit %d", g, group_bits[g], group_pos[g]);
    k->tma_pos[g] = group_pos[g];
    k->tma_bits[g] = group_bits[g];
  }
  // trial encoding (any 16-byte aligned address will do): a description the driver rejects is reported now
  Driver* d = driver();
  CUtensorMap probe;
  const int rc = qls_encode_tile_map(d, k, prog->pool_dev ? (void*)prog->pool_dev : (void*)(uintptr_t)0x1000, n_local, &probe);
  if (rc) k->tma_low_bits = k->tma_groups = 0;
  return rc;
}

static void* g_trace_dev = nullptr;
extern "C" int qls_set_trace_buffer(void* trace_dev) {
  g_trace_dev = trace_dev;
  return QLS_OK;
}

void qls_jit_release(QlsProgram* prog) {
  if (!prog->jit) return;
  Driver* d = driver();
  for (uint32_t i = 0; i < prog->n_passes; ++i)
    if (prog->jit[i]) {
      if (d) d->ModuleUnload(prog->jit[i]->mod);
      delete prog->jit[i];
    }
  delete[] prog->jit;
  prog->jit = nullptr;
}

bool qls_jit_eligible(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, int n_local) {
  return prog->jit && prog->jit[pass_index] && p.tile_bits == 12 && n_local >= p.tile_bits;
}

static int g_grid_reserve = 0;
// SMs the specialised pass kernels leave free (a persistent CTA per SM otherwise holds every SM for the whole pass, and the
// barrier kernel of an exchange running beside it would wait for the pass to end)
extern "C" int qls_set_pass_grid_reserve(int n_sms) {
  QLS_ARG(n_sms >= 0 && n_sms < 64, "bad SM reserve %d", n_sms);
  g_grid_reserve = n_sms;
  return QLS_OK;
}

int qls_launch_jit_pass(const QlsProgram* prog, const QlsPass& p, uint32_t pass_index, void* state, int n_local,
                        uint64_t index_offset, cudaStream_t stream, int sel_pos, int sel_val, int sel_pos2, int sel_val2) {
  Driver* d = driver();
  const QlsJitKernel* k = prog->jit[pass_index];
  const int fixed = __builtin_popcountll(p.fix_mask);
  const int free_bits = n_local - p.tile_bits - fixed;
  QLS_ARG(free_bits >= 0 && free_bits < 40, "pass geometry: n_local=%d tile_bits=%d fixed=%d", n_local, p.tile_bits, fixed);
  uint64_t n_tiles = 1ull << free_bits;
  // restrict to the tiles whose state-index bits sel_pos (, sel_pos2) have the given values: find their bits in the tile ordinal
  uint32_t sel_ord[2] = {64u, 64u}, sel_v[2] = {0u, 0u};
  {
    const int pos[2] = {sel_pos, sel_pos2}, val[2] = {sel_val, sel_val2};
    int n_sel = 0;
    for (int s = 0; s < 2; ++s) {
      if (pos[s] < 0) continue;
      int ord = -1;
      for (int j = 0; j < p.n_fseg; ++j)
        if (pos[s] >= p.fseg[j].dst && pos[s] < p.fseg[j].dst + p.fseg[j].width) ord = p.fseg[j].src + (pos[s] - p.fseg[j].dst);
      QLS_ARG(ord >= 0, "pass %u: index bit %d is not a free position of the pass", pass_index, pos[s]);
      sel_ord[n_sel] = (uint32_t)ord;
      sel_v[n_sel] = (uint32_t)(val[s] & 1);
      ++n_sel;
    }
    QLS_ARG(free_bits >= n_sel, "pass %u: too few free positions", pass_index);
    if (n_sel == 2) {
      QLS_ARG(sel_ord[0] != sel_ord[1], "pass %u: the two selection bits coincide", pass_index);
      if (sel_ord[0] > sel_ord[1]) {  // the kernel inserts the lower ordinal bit first
        const uint32_t o = sel_ord[0], v = sel_v[0];
        sel_ord[0] = sel_ord[1];
        sel_v[0] = sel_v[1];
        sel_ord[1] = o;
        sel_v[1] = v;
      }
    }
    n_tiles >>= n_sel;
  }
  static int sm_count[64] = {0};
  if (!sm_count[prog->device & 63])
    QLS_CUDA(cudaDeviceGetAttribute(&sm_count[prog->device & 63], cudaDevAttrMultiProcessorCount, prog->device));
  uint64_t grid = (uint64_t)(sm_count[prog->device & 63] - g_grid_reserve);
  if (grid < 1) grid = 1;
  if (grid > n_tiles) grid = n_tiles;
  const void* pool = prog->pool_dev;
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (k->tma_low_bits) {
    const int erc = qls_encode_tile_map(d, k, state, n_local, &tmap);
    if (erc) return erc;
  }
  void* trace = g_trace_dev;
  void* args[] = {&state, &index_offset, &n_tiles, &pool, &tmap, &trace, &sel_ord[0], &sel_v[0], &sel_ord[1], &sel_v[1]};
  const int rc = d->LaunchKernel(k->fn, (unsigned)grid, 1, 1, k->threads, 1, 1, k->smem, stream, args, nullptr);
  if (rc) {
    qls_set_error("cuLaunchKernel (pass %u): %s", pass_index, cu_err(d, rc));
    return QLS_ERR_CUDA;
  }
  return QLS_OK;
}
