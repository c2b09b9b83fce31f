// bo_api.cpp -- C ABI of libb200optas.so (see include/b200optas.h).
//
// Host side only: validates the tapes, generates the CUDA source (bo_codegen.cpp), compiles it
// for sm_100a with NVRTC (cubins cached on disk, keyed by a hash of source + headers + compiler
// version), loads it through the CUDA driver API and launches it.  libcuda.so.1 is opened lazily
// with dlopen so that the library itself loads (and can compile) on a GPU-less machine; every
// entry point that needs a device fails with BO_ERR_NO_DEVICE there -- there is no CPU fallback.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>
#include <string>
#include <vector>

#include "b200optas.h"
#include "bo_codegen.h"
#include "bo_coop.h"
#include "bo_opcodes.h"
#include "bo_sparse.h"
#include "bo_team.h"

namespace {

thread_local std::string g_err;

int set_err(int code, const char* fmt, ...) {
  char buf[4096];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

// ------------------------------------------------------------------------------------------
// CUDA driver API, resolved at run time
// ------------------------------------------------------------------------------------------
struct Driver {
  void* lib = nullptr;
  bool tried = false;
  std::string why;
#define BO_DRV(name) decltype(&::name) name = nullptr;
  BO_DRV(cuInit)
  BO_DRV(cuDeviceGetCount)
  BO_DRV(cuDeviceGet)
  BO_DRV(cuDeviceGetAttribute)
  BO_DRV(cuCtxGetCurrent)
  BO_DRV(cuCtxSetCurrent)
  BO_DRV(cuCtxGetDevice)
  BO_DRV(cuCtxPushCurrent_v2)
  BO_DRV(cuCtxPopCurrent_v2)
  BO_DRV(cuDevicePrimaryCtxRetain)
  BO_DRV(cuModuleLoadData)
  BO_DRV(cuModuleUnload)
  BO_DRV(cuModuleGetFunction)
  BO_DRV(cuFuncGetAttribute)
  BO_DRV(cuFuncSetAttribute)
  BO_DRV(cuOccupancyMaxActiveBlocksPerMultiprocessor)
  BO_DRV(cuLaunchKernel)
  BO_DRV(cuMemAlloc_v2)
  BO_DRV(cuMemFree_v2)
  BO_DRV(cuMemcpyHtoDAsync_v2)
  BO_DRV(cuMemcpyDtoHAsync_v2)
  BO_DRV(cuMemsetD8Async)
  BO_DRV(cuStreamSynchronize)
  BO_DRV(cuStreamCreate)
  BO_DRV(cuStreamDestroy_v2)
  BO_DRV(cuStreamWaitEvent)
  BO_DRV(cuPointerGetAttribute)
  BO_DRV(cuEventCreate)
  BO_DRV(cuEventDestroy_v2)
  BO_DRV(cuEventRecord)
  BO_DRV(cuEventSynchronize)
  BO_DRV(cuEventElapsedTime)
  BO_DRV(cuGetErrorString)
#undef BO_DRV
};

Driver g_drv;

bool load_driver() {
  if (g_drv.tried) return g_drv.lib != nullptr;
  g_drv.tried = true;
  void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    g_drv.why = "libcuda.so.1 not found (no NVIDIA driver on this machine)";
    return false;
  }
  bool ok = true;
#define BO_SYM(field, sym)                                          \
  g_drv.field = reinterpret_cast<decltype(g_drv.field)>(dlsym(lib, sym)); \
  if (!g_drv.field) { ok = false; g_drv.why = std::string("missing driver symbol ") + sym; }
  BO_SYM(cuInit, "cuInit")
  BO_SYM(cuDeviceGetCount, "cuDeviceGetCount")
  BO_SYM(cuDeviceGet, "cuDeviceGet")
  BO_SYM(cuDeviceGetAttribute, "cuDeviceGetAttribute")
  BO_SYM(cuCtxGetCurrent, "cuCtxGetCurrent")
  BO_SYM(cuCtxSetCurrent, "cuCtxSetCurrent")
  BO_SYM(cuCtxGetDevice, "cuCtxGetDevice")
  BO_SYM(cuCtxPushCurrent_v2, "cuCtxPushCurrent_v2")
  BO_SYM(cuCtxPopCurrent_v2, "cuCtxPopCurrent_v2")
  BO_SYM(cuDevicePrimaryCtxRetain, "cuDevicePrimaryCtxRetain")
  BO_SYM(cuModuleLoadData, "cuModuleLoadData")
  BO_SYM(cuModuleUnload, "cuModuleUnload")
  BO_SYM(cuModuleGetFunction, "cuModuleGetFunction")
  BO_SYM(cuFuncGetAttribute, "cuFuncGetAttribute")
  BO_SYM(cuFuncSetAttribute, "cuFuncSetAttribute")
  BO_SYM(cuOccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
  BO_SYM(cuLaunchKernel, "cuLaunchKernel")
  BO_SYM(cuMemAlloc_v2, "cuMemAlloc_v2")
  BO_SYM(cuMemFree_v2, "cuMemFree_v2")
  BO_SYM(cuMemcpyHtoDAsync_v2, "cuMemcpyHtoDAsync_v2")
  BO_SYM(cuMemcpyDtoHAsync_v2, "cuMemcpyDtoHAsync_v2")
  BO_SYM(cuMemsetD8Async, "cuMemsetD8Async")
  BO_SYM(cuStreamSynchronize, "cuStreamSynchronize")
  BO_SYM(cuStreamCreate, "cuStreamCreate")
  BO_SYM(cuStreamDestroy_v2, "cuStreamDestroy_v2")
  BO_SYM(cuStreamWaitEvent, "cuStreamWaitEvent")
  BO_SYM(cuPointerGetAttribute, "cuPointerGetAttribute")
  BO_SYM(cuEventCreate, "cuEventCreate")
  BO_SYM(cuEventDestroy_v2, "cuEventDestroy_v2")
  BO_SYM(cuEventRecord, "cuEventRecord")
  BO_SYM(cuEventSynchronize, "cuEventSynchronize")
  BO_SYM(cuEventElapsedTime, "cuEventElapsedTime")
  BO_SYM(cuGetErrorString, "cuGetErrorString")
#undef BO_SYM
  if (!ok) {
    dlclose(lib);
    return false;
  }
  if (g_drv.cuInit(0) != CUDA_SUCCESS) {
    g_drv.why = "cuInit failed (no usable CUDA device)";
    dlclose(lib);
    return false;
  }
  g_drv.lib = lib;
  return true;
}

const char* cu_err(CUresult r) {
  const char* s = nullptr;
  if (g_drv.cuGetErrorString && g_drv.cuGetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown CUDA error";
}

#define BO_CU(call)                                                                       \
  do {                                                                                    \
    CUresult _r = (call);                                                                 \
    if (_r != CUDA_SUCCESS) return set_err(BO_ERR_CUDA, "%s failed: %s", #call, cu_err(_r)); \
  } while (0)

// Make sure a context is current on this thread.  `device` >= 0 names the ordinal explicitly (bo_options.device =
// ordinal + 1); otherwise the context already current on the thread is used (torch's primary context of the device the
// caller selected and touched), and only when there is none the primary context of device 0.
int ensure_context(CUdevice* dev_out, int device = -1, CUcontext* ctx_out = nullptr) {
  if (!load_driver()) return set_err(BO_ERR_NO_DEVICE, "%s", g_drv.why.c_str());
  int n = 0;
  if (g_drv.cuDeviceGetCount(&n) != CUDA_SUCCESS || n == 0) return set_err(BO_ERR_NO_DEVICE, "no CUDA device");
  CUcontext ctx = nullptr;
  g_drv.cuCtxGetCurrent(&ctx);
  if (device >= 0) {
    if (device >= n) return set_err(BO_ERR_INVALID, "device ordinal %d out of range (%d devices)", device, n);
    CUdevice cur = -1;
    if (!ctx || g_drv.cuCtxGetDevice(&cur) != CUDA_SUCCESS || cur != device) {
      CUdevice dev;
      BO_CU(g_drv.cuDeviceGet(&dev, device));
      BO_CU(g_drv.cuDevicePrimaryCtxRetain(&ctx, dev));
      BO_CU(g_drv.cuCtxSetCurrent(ctx));
    }
  } else if (!ctx) {
    CUdevice dev;
    BO_CU(g_drv.cuDeviceGet(&dev, 0));
    BO_CU(g_drv.cuDevicePrimaryCtxRetain(&ctx, dev));
    BO_CU(g_drv.cuCtxSetCurrent(ctx));
  }
  if (dev_out) BO_CU(g_drv.cuCtxGetDevice(dev_out));
  if (ctx_out) *ctx_out = ctx;
  return BO_OK;
}

// A handle's module, buffers and events belong to the context it was created in: make that context current for the
// duration of a call that arrives on a thread whose current context is another one (or none).
struct CtxScope {
  bool pushed = false;
  int enter(CUcontext want) {
    if (!want) return BO_OK;
    CUcontext cur = nullptr;
    g_drv.cuCtxGetCurrent(&cur);
    if (cur == want) return BO_OK;
    BO_CU(g_drv.cuCtxPushCurrent_v2(want));
    pushed = true;
    return BO_OK;
  }
  ~CtxScope() {
    if (pushed) {
      CUcontext old;
      g_drv.cuCtxPopCurrent_v2(&old);
    }
  }
};

// ------------------------------------------------------------------------------------------
// paths, hashing, files
// ------------------------------------------------------------------------------------------
std::string lib_dir() {
  Dl_info info;
  if (dladdr(reinterpret_cast<void*>(&bo_abi_version), &info) && info.dli_fname) {
    std::string p(info.dli_fname);
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? "." : p.substr(0, k);
  }
  return ".";
}

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ULL) {
  for (unsigned char c : s) {
    h ^= c;
    h *= 1099511628211ULL;
  }
  return h;
}

bool read_file(const std::string& path, std::string* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  *out = ss.str();
  return true;
}

bool write_file(const std::string& path, const std::string& data) {
  const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  {
    std::ofstream f(tmp, std::ios::binary);
    if (!f) return false;
    f.write(data.data(), (std::streamsize)data.size());
    if (!f) return false;
  }
  return std::rename(tmp.c_str(), path.c_str()) == 0;
}

// ------------------------------------------------------------------------------------------
// JIT: source -> cubin (NVRTC, sm_100a), with an on-disk cache
// ------------------------------------------------------------------------------------------
struct Compiled {
  std::string cubin;
  std::string log;
  int regs = -1, local_bytes = -1, smem_bytes = -1;  // parsed from the ptxas -v log when available
  bool from_cache = false;
};

void parse_ptxas(const std::string& log, const char* kernel, Compiled* c) {
  // "ptxas info    : Compiling entry function 'bo_solve_kernel' for 'sm_100a'" followed (after the
  // lines of its non-inlined device functions) by "... Used N registers, ..." for the entry itself:
  // take the LAST "Used" line after the entry marker, and the largest stack frame reported.
  const size_t at = log.find(std::string("entry function '") + kernel + "'");
  if (at == std::string::npos) return;
  size_t pos = at, used = std::string::npos;
  while ((pos = log.find("Used ", pos)) != std::string::npos) {
    used = pos;
    pos += 5;
  }
  if (used != std::string::npos) c->regs = atoi(log.c_str() + used + 5);
  pos = at;
  int frame = 0;
  while ((pos = log.find(" bytes stack frame", pos)) != std::string::npos) {
    size_t b = pos;
    while (b > 0 && isdigit((unsigned char)log[b - 1])) --b;
    frame = std::max(frame, atoi(log.c_str() + b));
    pos += 5;
  }
  c->local_bytes = frame;
  const size_t sm = log.find(" bytes smem", at);
  if (sm != std::string::npos) {
    size_t b = sm;
    while (b > 0 && isdigit((unsigned char)log[b - 1])) --b;
    c->smem_bytes = atoi(log.c_str() + b);
  } else {
    c->smem_bytes = 0;
  }
}

int jit_compile(const std::string& source, const char* unit_name, const char* kernel, const bo_options& opts,
                Compiled* out) {
  const std::string inc = opts.include_dir ? opts.include_dir : lib_dir() + "/csrc/jit";
  const std::string cache = opts.cache_dir ? opts.cache_dir : lib_dir() + "/_jitcache";
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);

  // hash = source + every header it may include + compiler version
  uint64_t h = fnv1a(source);
  for (const char* hdr : {"bo_common.cuh", "bo_ipm_reg.cuh", "bo_qp_reg.cuh", "bo_ipm_cta.cuh", "bo_stream_eval.cuh", "bo_ipm_team.cuh", "bo_team_layout.cuh"}) {
    std::string text;
    if (!read_file(inc + "/" + hdr, &text)) return set_err(BO_ERR_INVALID, "JIT header %s/%s not found", inc.c_str(), hdr);
    h = fnv1a(text, h);
  }
  // experiment hook: extra -D options for the kernel templates (part of the cache key)
  const char* extra_env = getenv("B200OPTAS_JIT_DEFINES");
  const std::string extra = extra_env ? extra_env : "";
  h = fnv1a("nvrtc" + std::to_string(major) + "." + std::to_string(minor) + "sm_100a-lineinfo" + extra, h);
  char hex[32];
  snprintf(hex, sizeof hex, "%016llx", (unsigned long long)h);
  const std::string stem = cache + "/" + unit_name + "_" + hex;
  const bool use_cache = !(opts.flags & BO_FLAG_NO_CACHE);

  if (use_cache && read_file(stem + ".cubin", &out->cubin) && !out->cubin.empty()) {
    read_file(stem + ".log", &out->log);
    out->from_cache = true;
    parse_ptxas(out->log, kernel, out);
    if (opts.flags & BO_FLAG_VERBOSE) fprintf(stderr, "[b200optas] cubin cache hit %s.cubin\n", stem.c_str());
    return BO_OK;
  }

  nvrtcProgram prog;
  const std::string file = std::string(unit_name) + ".cu";
  if (nvrtcCreateProgram(&prog, source.c_str(), file.c_str(), 0, nullptr, nullptr) != NVRTC_SUCCESS)
    return set_err(BO_ERR_COMPILE, "nvrtcCreateProgram failed");
  const std::string iflag = "-I" + inc;
  const std::string iflag2 = "-I" + (opts.include_dir ? std::string(opts.include_dir) : lib_dir() + "/../include");
  std::vector<std::string> extra_flags;
  {
    std::istringstream ss(extra);
    std::string tok;
    while (ss >> tok) extra_flags.push_back(tok);
  }
  std::vector<const char*> flags = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--ptxas-options=-v",
                                    iflag.c_str(), iflag2.c_str()};
  for (const auto& f : extra_flags) flags.push_back(f.c_str());
  const nvrtcResult res = nvrtcCompileProgram(prog, (int)flags.size(), flags.data());
  size_t log_size = 0;
  nvrtcGetProgramLogSize(prog, &log_size);
  out->log.assign(log_size, '\0');
  if (log_size) nvrtcGetProgramLog(prog, &out->log[0]);
  if (res != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    return set_err(BO_ERR_COMPILE, "NVRTC: %s\n%s", nvrtcGetErrorString(res), out->log.c_str());
  }
  size_t sz = 0;
  if (nvrtcGetCUBINSize(prog, &sz) != NVRTC_SUCCESS || sz == 0) {
    nvrtcDestroyProgram(&prog);
    return set_err(BO_ERR_COMPILE, "NVRTC produced no cubin");
  }
  out->cubin.assign(sz, '\0');
  nvrtcGetCUBIN(prog, &out->cubin[0]);
  nvrtcDestroyProgram(&prog);
  parse_ptxas(out->log, kernel, out);
  if (opts.flags & BO_FLAG_VERBOSE) fprintf(stderr, "[b200optas] compiled %s: %s\n", unit_name, out->log.c_str());
  if (use_cache) {
    mkdir(cache.c_str(), 0755);
    if (write_file(stem + ".cubin", out->cubin)) {
      write_file(stem + ".log", out->log);
      write_file(stem + ".cu", source);
    }
  }
  return BO_OK;
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
struct LoadedKernel {
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  int regs = -1, local_bytes = -1, smem_bytes = -1;
};

int load_kernel(const Compiled& c, const char* name, LoadedKernel* k) {
  BO_CU(g_drv.cuModuleLoadData(&k->mod, c.cubin.data()));
  BO_CU(g_drv.cuModuleGetFunction(&k->fn, k->mod, name));
  g_drv.cuFuncGetAttribute(&k->regs, CU_FUNC_ATTRIBUTE_NUM_REGS, k->fn);
  g_drv.cuFuncGetAttribute(&k->local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, k->fn);
  g_drv.cuFuncGetAttribute(&k->smem_bytes, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, k->fn);
  return BO_OK;
}

bool is_device_ptr(const void* p) {
  if (!p) return false;
  unsigned int type = 0;
  const CUresult r = g_drv.cuPointerGetAttribute(&type, CU_POINTER_ATTRIBUTE_MEMORY_TYPE, (CUdeviceptr)(uintptr_t)p);
  if (r != CUDA_SUCCESS) return false;  // plain (unregistered) host memory
  return type == CU_MEMORYTYPE_DEVICE || type == CU_MEMORYTYPE_UNIFIED;
}

// Page-locked host memory is mapped into the device address space (unified addressing): a kernel can read and write it in
// place.  Returns the device-side address of such a buffer, 0 for anything else (pageable host memory, device memory).
CUdeviceptr mapped_host_ptr(const void* p) {
  if (!p) return 0;
  unsigned int type = 0;
  if (g_drv.cuPointerGetAttribute(&type, CU_POINTER_ATTRIBUTE_MEMORY_TYPE, (CUdeviceptr)(uintptr_t)p) != CUDA_SUCCESS) return 0;
  if (type != CU_MEMORYTYPE_HOST) return 0;
  CUdeviceptr d = 0;
  if (g_drv.cuPointerGetAttribute(&d, CU_POINTER_ATTRIBUTE_DEVICE_POINTER, (CUdeviceptr)(uintptr_t)p) != CUDA_SUCCESS) return 0;
  return d;
}

// grow-only device scratch buffer
struct DevBuf {
  CUdeviceptr ptr = 0;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return BO_OK;
    if (ptr) g_drv.cuMemFree_v2(ptr);
    ptr = 0;
    cap = 0;
    BO_CU(g_drv.cuMemAlloc_v2(&ptr, bytes));
    cap = bytes;
    return BO_OK;
  }
  void release() {
    if (ptr && g_drv.cuMemFree_v2) g_drv.cuMemFree_v2(ptr);
    ptr = 0;
    cap = 0;
  }
};

struct Timer {
  bool enabled = false;
  std::vector<std::pair<CUevent, CUevent>> pool;
  size_t used = 0;
  double carried_ms = 0.0;
  int64_t carried_n = 0;
  int begin(CUstream s, size_t* slot) {
    if (!enabled) return BO_OK;
    if (used == pool.size()) {
      if (pool.size() >= 4096) {
        double ms = 0.0;
        int64_t n = 0;
        collect(&ms, &n);
        carried_ms += ms;
        carried_n += n;
      } else {
        CUevent a, b;
        BO_CU(g_drv.cuEventCreate(&a, CU_EVENT_DEFAULT));
        BO_CU(g_drv.cuEventCreate(&b, CU_EVENT_DEFAULT));
        pool.emplace_back(a, b);
      }
    }
    *slot = used++;
    BO_CU(g_drv.cuEventRecord(pool[*slot].first, s));
    return BO_OK;
  }
  int end(CUstream s, size_t slot) {
    if (!enabled) return BO_OK;
    BO_CU(g_drv.cuEventRecord(pool[slot].second, s));
    return BO_OK;
  }
  int collect(double* ms_total, int64_t* n) {
    double tot = carried_ms;
    int64_t cnt = carried_n;
    for (size_t i = 0; i < used; ++i) {
      float ms = 0.f;
      BO_CU(g_drv.cuEventSynchronize(pool[i].second));
      BO_CU(g_drv.cuEventElapsedTime(&ms, pool[i].first, pool[i].second));
      tot += ms;
      ++cnt;
    }
    used = 0;
    carried_ms = 0.0;
    carried_n = 0;
    *ms_total = tot;
    *n = cnt;
    return BO_OK;
  }
  void release() {
    for (auto& p : pool) {
      g_drv.cuEventDestroy_v2(p.first);
      g_drv.cuEventDestroy_v2(p.second);
    }
    pool.clear();
  }
};

bo_options normalise(const bo_options* in) {
  bo_options o;
  memset(&o, 0, sizeof o);
  if (in) o = *in;
  const bool user_max_iter = o.max_iter > 0;
  if (o.max_iter <= 0) o.max_iter = 100;
  // every iteration costs at least one trip: a caller who raises max_iter without naming a trip budget must not be
  // cut off by the default one
  if (o.max_trips <= 0) o.max_trips = user_max_iter ? std::max(250, (5 * o.max_iter) / 2) : 250;
  if (!(o.tol > 0)) o.tol = 1e-8;
  if (!(o.acceptable_tol > 0)) o.acceptable_tol = 1e-6;
  if (!(o.mu_init > 0)) o.mu_init = 0.1;
  // max_step == 0 ("default") is resolved per problem in bo_problem_create: 0.5 when the tapes contain trigonometric
  // kinematics, unlimited otherwise
  return o;
}

// The step cap exists for Newton steps of several radians through trigonometric kinematics; on problems without any
// (QPs, linear models, unscaled task-space variables) it would only limit how far from the seed an optimum may lie.
bool tape_has_trig(const bo::Tape& t) {
  for (int64_t i = 0; i < t.n_instr(); ++i) {
    const int op = t.instr[4 * i] & 0xFF;
    if (op == BO_OP_SIN || op == BO_OP_COS || op == BO_OP_TAN) return true;
  }
  return false;
}

struct SolverParams {  // must match bo_solver_params in csrc/jit/bo_common.cuh
  int32_t max_iter;
  double tol;
  double acceptable_tol;
  double mu_init;
  double max_step;
  int32_t max_trips;
  CUdeviceptr ldl_tab;
  CUdeviceptr dtab;
  CUdeviceptr scratch;
  long long scratch_stride;
};

}  // namespace

struct bo_problem {
  bo::ProblemSource ps;
  bo_options opts;
  std::string cache_dir, include_dir;
  std::string source;
  Compiled compiled;
  LoadedKernel kernel;
  bool loaded = false;
  int tpb = 64;
  DevBuf d_p, d_x0, d_x, d_lam, d_f, d_status, d_iters, d_kkt, d_counter, d_ldl_tab;
  bo::SparsePlan plan;
  bo::CoopPlan coop_plan;
  bool sparse = false, large = false, coop = false, team = false, qp = false;
  bo::TeamPlan team_plan;
  int smem_dynamic = 0;
  DevBuf d_dtab, d_scratch;
  int blocks_per_sm = 1, n_sm = 1;
  CUcontext ctx = nullptr;
  CUstream pipe[2] = {nullptr, nullptr};  // host-buffer calls: chunks alternate between two streams (copy / compute overlap)
  Timer timer;
};

struct bo_function {
  bo::Tape tape;
  bo_options opts;
  std::string cache_dir, include_dir;
  std::string source;
  Compiled compiled;
  LoadedKernel kernel;
  bool loaded = false;
  int tpb = 128;
  int smem_dynamic = 0;
  int blocks_per_sm = 1, n_sm = 1;
  std::vector<DevBuf> d_in, d_out;
  CUcontext ctx = nullptr;
  Timer timer;
};

extern "C" {

int bo_abi_version(void) { return BO_ABI_VERSION; }

const char* bo_last_error(void) { return g_err.c_str(); }

int bo_device_count(void) {
  if (!load_driver()) return 0;
  int n = 0;
  if (g_drv.cuDeviceGetCount(&n) != CUDA_SUCCESS) return 0;
  return n;
}

// ------------------------------------------------------------------------------------------
// solver
// ------------------------------------------------------------------------------------------
int bo_problem_create(const bo_problem_desc* desc, const bo_options* opts_in, bo_problem** out) {
  if (!desc || !out) return set_err(BO_ERR_INVALID, "bo_problem_create: null argument");
  *out = nullptr;
  if (desc->nx <= 0 || desc->np < 0 || desc->n_eq < 0 || desc->n_ineq < 0)
    return set_err(BO_ERR_INVALID, "bo_problem_create: bad dimensions");
  std::unique_ptr<bo_problem> pr(new bo_problem);
  pr->opts = normalise(opts_in);
  if (pr->opts.cache_dir) { pr->cache_dir = pr->opts.cache_dir; pr->opts.cache_dir = pr->cache_dir.c_str(); }
  if (pr->opts.include_dir) { pr->include_dir = pr->opts.include_dir; pr->opts.include_dir = pr->include_dir.c_str(); }
  bo::ProblemSource& ps = pr->ps;
  ps.nx = desc->nx;
  ps.np = desc->np;
  ps.n_eq = desc->n_eq;
  ps.n_ineq = desc->n_ineq;
  std::string err;
  if (!bo::copy_tape(desc->fc, &ps.fc, &err) || !bo::copy_tape(desc->kkt, &ps.kkt, &err) ||
      !bo::copy_sparsity(desc->jac_eq, ps.n_eq, ps.nx, false, &ps.jac_eq, &err) ||
      !bo::copy_sparsity(desc->jac_ineq, ps.n_ineq, ps.nx, false, &ps.jac_ineq, &err) ||
      !bo::copy_sparsity(desc->hess, ps.nx, ps.nx, true, &ps.hess, &err))
    return set_err(BO_ERR_INVALID, "bo_problem_create: %s", err.c_str());
  auto expect = [&](const bo::Tape& t, std::vector<int32_t> in, std::vector<int32_t> outs, const char* nm) {
    if (t.in_sizes != in || t.out_sizes != outs) {
      err = std::string(nm) + " tape has the wrong input/output segment sizes";
      return false;
    }
    return true;
  };
  if (!expect(ps.fc, {ps.nx, ps.np}, {1, ps.n_eq, ps.n_ineq}, "fc") ||
      !expect(ps.kkt, {ps.nx, ps.np, ps.n_eq, ps.n_ineq},
              {1, ps.nx, ps.n_eq, ps.n_ineq, ps.jac_eq.nnz(), ps.jac_ineq.nnz(), ps.hess.nnz()}, "kkt"))
    return set_err(BO_ERR_INVALID, "bo_problem_create: %s", err.c_str());
  // Tiers (all one instance per thread, see csrc/jit/bo_ipm_reg.cuh):
  //   dense   nx+n_eq <= 14            generated straight-line code, unrolled LDL' in registers
  //   sparse  otherwise                generated tapes, table-driven sparse LDL' in thread-local memory
  //   large   tapes > 30k instructions everything table-driven (interpreted tapes), factor in global scratch
  if (ps.nx + ps.n_eq > 8192 || ps.n_ineq > 16384)
    return set_err(BO_ERR_UNSUPPORTED, "bo_problem_create: nx+n_eq=%d, n_ineq=%d exceeds the built tiers", ps.nx + ps.n_eq,
                   ps.n_ineq);
  // whole warps only: the kernels use full-mask warp votes
  pr->tpb = pr->opts.threads_per_block > 0 ? ((pr->opts.threads_per_block + 31) / 32) * 32 : 64;
  const bool auto_max_step = pr->opts.max_step == 0.0;
  if (pr->opts.max_step == 0.0) pr->opts.max_step = (tape_has_trig(ps.kkt) || tape_has_trig(ps.fc)) ? 0.5 : -1.0;
  if (getenv("BO_DEBUG")) fprintf(stderr, "[bo] emitting source\n");
  // KKT systems beyond a dozen rows are factored sparsely: symbolic analysis here, once
  const bool pivoted = (pr->opts.flags & BO_FLAG_PIVOTED_LDL) != 0;
  pr->sparse = !pivoted && (ps.nx + ps.n_eq > 14);
  pr->large = pr->sparse && (ps.kkt.n_instr() > 30000 || ps.nx + ps.n_eq > 400);
  if (pivoted && ps.nx + ps.n_eq > 160)
    return set_err(BO_ERR_UNSUPPORTED, "bo_problem_create: BO_FLAG_PIVOTED_LDL is limited to nx+n_eq <= 160");
  // Cooperative tier (one instance per CTA, factor in shared memory; csrc/jit/bo_ipm_cta.cuh): the default for
  // everything the large tier used to take, on request (BO_FLAG_COOP) for any sparse-tier problem.
  if (pr->sparse && !(pr->opts.flags & BO_FLAG_NO_COOP)) {
    const int tpb = pr->opts.threads_per_block > 0 ? ((pr->opts.threads_per_block + 31) / 32) * 32 : (pr->large ? 256 : 64);
    bo::CoopPlan cp = bo::make_coop_plan(ps, tpb, pr->opts.cache_dir ? std::string(pr->opts.cache_dir) : lib_dir() + "/_jitcache");
    size_t smem = (size_t)cp.smem_doubles * sizeof(double);
    {
      // Experiment knob B200OPTAS_COOP_W_SMEM=1: the vectors of the instance in shared memory as well (when they fit).
      // Measured on B200 and left OFF: C3 789 k -> 582 k inst/s (35 KB per CTA: 6 resident CTAs per SM instead of 16; the
      // tier is bound by dependent-instruction latency and lives on resident CTAs, not on the L2 round trips of its
      // workspace), joint-space planner 51.0 k -> 52.2 k.
      const size_t with_w = smem + bo::coop_scratch_doubles(ps, cp) * sizeof(double);
      const char* env = getenv("B200OPTAS_COOP_W_SMEM");
      const size_t limit = (env && atoi(env)) ? 200 * 1024 : 0;
      if (with_w <= limit) {
        cp.w_in_smem = true;
        smem = with_w;
      }
    }
    // worth it when the tapes split into enough independent pieces to occupy the CTA (horizon problems do: one
    // piece per stage), or when the problem is too large for a thread anyway
    // ... and the per-instance state is big enough that a CTA beats a thread (the thread-per-instance tiers keep it in
    // thread-local memory and fall off a cliff once that stops fitting L1).  Measured on B200 (tools/tier_choice.py,
    // tools/tier_break_even.py; profiles/r01_tier_choice.txt, r01_tier_break_even.txt), thread vs CTA in inst/s:
    //   position + axis IK  nx 21, 20 eq, 14 ineq (5.4 KB per lane)   15.8 M  vs 1.16 M
    //   MPC tick T = 6      nx 24, 14 eq, 54 ineq (10 KB per lane)     2.23 M vs 2.52 M
    //   MPC tick T = 10     nx 40, 22 eq, 90 ineq (17 KB per lane)     0.74 M vs 1.64 M
    //   MPC tick T = 20     nx 80, 42 eq, 180 ineq                     0.12 M vs 0.81 M
    const bool big_enough = ps.nx + ps.n_eq + ps.n_ineq > 64;
    const bool wanted = pr->large || (pr->opts.flags & BO_FLAG_COOP) || (cp.kkt.n_components >= 16 && big_enough);
    if (wanted && cp.vals_size() + ps.nx + ps.n_eq + 2 < 32767 && smem <= 227 * 1024) {  // 15-bit target indices reach into bp
      pr->coop = true;
      pr->large = false;
      pr->tpb = tpb;
      pr->smem_dynamic = (int)smem;
      pr->coop_plan = std::move(cp);
    }
  }
  // QP path (csrc/jit/bo_qp_reg.cuh): quadratic cost with linear constraints, decided from the tape.  The thread-per-instance
  // tiers run it (it needs no trial points and no per-iteration tape, so its state is small); horizon-sized QPs that
  // qualify for the cooperative tier stay there.
  pr->qp = !pr->coop && !pr->large && !pivoted && !(pr->opts.flags & (BO_FLAG_NO_QP | BO_FLAG_TEAM)) && bo::problem_is_qp(ps);
  if (pr->qp && auto_max_step) pr->opts.max_step = -1.0;  // sin / cos of the parameters only: the step cap has no role in a QP
  // Team tier (csrc/jit/bo_ipm_team.cuh): small dense problems, G threads in G warps per instance, state in shared memory
  // Measured on B200 (tools/team_check.py, profiles/r02_team_vs_thread.txt): the team tier wins once the per-instance state
  // no longer fits a thread (C2: 24 rows of variables + constraints, 3 KB of thread-local state: 6.0 vs 10.2 ms per 65536)
  // and loses below that (planar differential-IK QP, 13 rows, 1.7 KB: 0.51 vs 0.26 ms; Booth 14.5 vs 9.9 us per 4096).
  const bool team_worth_it = ps.nx + ps.n_eq + ps.n_ineq >= 20 || (pr->opts.flags & BO_FLAG_TEAM);
  if (!pr->sparse && !pivoted && !pr->qp && team_worth_it && !(pr->opts.flags & BO_FLAG_NO_TEAM)) {
    int G = pr->opts.threads_per_block > 0 ? pr->opts.threads_per_block / 32 : 4;
    G = std::max(1, std::min(G, 8));
    std::string why;
    const size_t smem = ((size_t)bo::team_smem_elems(ps, G) * 32 + 2 * 32 * (size_t)(ps.np + ps.nx)) * sizeof(double) + 32 * sizeof(int) + 128 + 32 * sizeof(long long);
    if (smem <= 227 * 1024 && bo::make_team_plan(ps, G, &pr->team_plan, &why)) {
      pr->team = true;
      pr->tpb = 32 * G;
      pr->smem_dynamic = (int)smem;
    } else if (pr->opts.flags & BO_FLAG_VERBOSE) {
      fprintf(stderr, "[b200optas] team tier not used: %s\n", why.empty() ? "shared memory" : why.c_str());
    }
  }
  if (pr->team) {
    pr->source = bo::emit_team_source(ps, pr->team_plan);
  } else if (pr->coop) {
    pr->source = bo::emit_coop_source(ps, pr->coop_plan, pr->tpb);
  } else {
    if (pr->sparse) pr->plan = bo::make_sparse_plan(ps, pr->large);
    pr->source = bo::emit_problem_source(ps, pr->tpb, pivoted, pr->sparse ? &pr->plan : nullptr, pr->large, pr->qp);
  }
  if (getenv("BO_DEBUG")) fprintf(stderr, "[bo] source %zu bytes\n", pr->source.size());
  int rc = jit_compile(pr->source, "bo_solve", "bo_solve_kernel", pr->opts, &pr->compiled);
  if (rc != BO_OK) return rc;
  pr->kernel.regs = pr->compiled.regs;
  pr->kernel.local_bytes = pr->compiled.local_bytes;
  pr->kernel.smem_bytes = pr->compiled.smem_bytes;
  if (!(pr->opts.flags & BO_FLAG_COMPILE_ONLY)) {
    CUdevice dev;
    rc = ensure_context(&dev, pr->opts.device - 1, &pr->ctx);
    if (rc != BO_OK) return rc;
    rc = load_kernel(pr->compiled, "bo_solve_kernel", &pr->kernel);
    if (rc != BO_OK) return rc;
    BO_CU(g_drv.cuDeviceGetAttribute(&pr->n_sm, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
    if (pr->smem_dynamic > 0)
      BO_CU(g_drv.cuFuncSetAttribute(pr->kernel.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, pr->smem_dynamic));
    // thread-per-instance tiers keep their state in thread-local memory: ask for the largest L1 split
    // (B200OPTAS_L1_CARVEOUT = shared-memory percentage 0..100 overrides; -1 leaves the driver's heuristic)
    if (!pr->coop && pr->smem_dynamic == 0) {
      const char* cv = getenv("B200OPTAS_L1_CARVEOUT");
      const int carve = cv ? atoi(cv) : -1;
      if (carve >= 0) BO_CU(g_drv.cuFuncSetAttribute(pr->kernel.fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, carve));
    }
    BO_CU(g_drv.cuOccupancyMaxActiveBlocksPerMultiprocessor(&pr->blocks_per_sm, pr->kernel.fn, pr->tpb, (size_t)pr->smem_dynamic));
    if (pr->blocks_per_sm < 1) pr->blocks_per_sm = 1;
    if (pr->opts.blocks_per_sm > 0 && pr->opts.blocks_per_sm < pr->blocks_per_sm) pr->blocks_per_sm = pr->opts.blocks_per_sm;
    if ((rc = pr->d_counter.reserve(16 * sizeof(unsigned long long))) != BO_OK) return rc;  // one work counter per chunk
    if (pr->coop) {
      const bo::CoopPlan& cp = pr->coop_plan;
      const size_t bytes = cp.itab.size() * sizeof(int32_t);
      if ((rc = pr->d_ldl_tab.reserve(bytes)) != BO_OK) return rc;
      BO_CU(g_drv.cuMemcpyHtoDAsync_v2(pr->d_ldl_tab.ptr, cp.itab.data(), bytes, nullptr));
      const size_t dbytes = std::max<size_t>(cp.dtab.size(), 1) * sizeof(double);
      if ((rc = pr->d_dtab.reserve(dbytes)) != BO_OK) return rc;
      if (!cp.dtab.empty()) BO_CU(g_drv.cuMemcpyHtoDAsync_v2(pr->d_dtab.ptr, cp.dtab.data(), cp.dtab.size() * sizeof(double), nullptr));
      const size_t ctas = (size_t)pr->n_sm * pr->blocks_per_sm;
      if ((rc = pr->d_scratch.reserve(ctas * bo::coop_scratch_doubles(ps, cp) * sizeof(double))) != BO_OK) return rc;
      BO_CU(g_drv.cuStreamSynchronize(nullptr));
    } else if (pr->sparse) {
      const size_t bytes = pr->plan.table.size() * sizeof(int32_t);
      if ((rc = pr->d_ldl_tab.reserve(bytes)) != BO_OK) return rc;
      BO_CU(g_drv.cuMemcpyHtoDAsync_v2(pr->d_ldl_tab.ptr, pr->plan.table.data(), bytes, nullptr));
      if (pr->large) {
        const size_t dbytes = std::max<size_t>(pr->plan.dtable.size(), 1) * sizeof(double);
        if ((rc = pr->d_dtab.reserve(dbytes)) != BO_OK) return rc;
        if (!pr->plan.dtable.empty())
          BO_CU(g_drv.cuMemcpyHtoDAsync_v2(pr->d_dtab.ptr, pr->plan.dtable.data(), pr->plan.dtable.size() * sizeof(double), nullptr));
        // large state per lane (hundreds of KB): cap the resident lanes so the scratch stays within a few GB
        {
          const int cap = pr->opts.blocks_per_sm > 0 ? pr->opts.blocks_per_sm : 4;  // latency-bound: lanes = throughput
          if (pr->blocks_per_sm > cap) pr->blocks_per_sm = cap;
        }
        const size_t lanes = (size_t)pr->n_sm * pr->blocks_per_sm * pr->tpb;
        if ((rc = pr->d_scratch.reserve(lanes * (size_t)pr->plan.vals_size() * sizeof(double))) != BO_OK) return rc;
      }
      BO_CU(g_drv.cuStreamSynchronize(nullptr));
    }
    pr->loaded = true;
    pr->timer.enabled = (pr->opts.flags & BO_FLAG_TIMING) != 0;
  }
  *out = pr.release();
  return BO_OK;
}

int bo_problem_destroy(bo_problem* pr) {
  if (!pr) return BO_OK;
  if (pr->loaded) {
    CtxScope scope;
    scope.enter(pr->ctx);
    for (DevBuf* b : {&pr->d_p, &pr->d_x0, &pr->d_x, &pr->d_lam, &pr->d_f, &pr->d_status, &pr->d_iters, &pr->d_kkt, &pr->d_counter, &pr->d_ldl_tab, &pr->d_dtab, &pr->d_scratch})
      b->release();
    pr->timer.release();
    for (CUstream& ps : pr->pipe)
      if (ps) g_drv.cuStreamDestroy_v2(ps);
    if (pr->kernel.mod) g_drv.cuModuleUnload(pr->kernel.mod);
  }
  delete pr;
  return BO_OK;
}

static int64_t copy_out(const std::string& s, char* buf, int64_t cap) {
  if (buf && cap > 0) {
    const int64_t n = (int64_t)s.size() < cap - 1 ? (int64_t)s.size() : cap - 1;
    memcpy(buf, s.data(), (size_t)n);
    buf[n] = '\0';
  }
  return (int64_t)s.size();
}

int64_t bo_problem_source(const bo_problem* pr, char* buf, int64_t cap) {
  if (!pr) return set_err(BO_ERR_INVALID, "null problem");
  return copy_out(pr->source, buf, cap);
}

int64_t bo_problem_ldl_table(const bo_problem* pr, int32_t* buf, int64_t cap) {
  if (!pr) return set_err(BO_ERR_INVALID, "null problem");
  const std::vector<int32_t>& tab = pr->coop ? pr->coop_plan.itab : pr->plan.table;
  const int64_t n = (pr->sparse || pr->coop) ? (int64_t)tab.size() : 0;
  if (buf && cap >= n && n > 0) memcpy(buf, tab.data(), (size_t)n * sizeof(int32_t));
  return n;
}

int64_t bo_problem_dtable(const bo_problem* pr, double* buf, int64_t cap) {
  if (!pr) return set_err(BO_ERR_INVALID, "null problem");
  const std::vector<double>& tab = pr->coop ? pr->coop_plan.dtab : pr->plan.dtable;
  const int64_t n = (pr->large || pr->coop) ? (int64_t)tab.size() : 0;
  if (buf && cap >= n && n > 0) memcpy(buf, tab.data(), (size_t)n * sizeof(double));
  return n;
}

int bo_problem_kernel_info(const bo_problem* pr, int32_t* regs, int32_t* local_bytes, int32_t* smem_bytes) {
  if (!pr) return set_err(BO_ERR_INVALID, "null problem");
  if (regs) *regs = pr->kernel.regs;
  if (local_bytes) *local_bytes = pr->kernel.local_bytes;
  if (smem_bytes) *smem_bytes = pr->kernel.smem_bytes;
  return BO_OK;
}

int bo_problem_options(const bo_problem* pr, bo_options* out) {
  if (!pr || !out) return set_err(BO_ERR_INVALID, "null argument");
  *out = pr->opts;
  out->cache_dir = nullptr;  // borrowed strings are not handed back
  out->include_dir = nullptr;
  return BO_OK;
}

int bo_problem_tier_info(const bo_problem* pr, int64_t* info, int32_t cap) {
  if (!pr) return set_err(BO_ERR_INVALID, "null problem");
  int64_t v[BO_TIER_INFO_LEN] = {0};
  v[0] = pr->team ? 4 : (pr->coop ? 3 : (pr->large ? 2 : (pr->sparse ? 1 : 0)));
  if (pr->qp) v[0] = pr->sparse ? 6 : 5;
  v[1] = pr->tpb;
  v[2] = pr->smem_dynamic;
  if (pr->coop) {
    const bo::CoopPlan& cp = pr->coop_plan;
    v[3] = cp.n_levels;
    v[4] = cp.fc.nsub;
    v[5] = cp.kkt.nsub;
    v[6] = cp.kkt.n_pe;
    v[7] = cp.kkt.n_part;
    v[8] = cp.kkt.max_len;
    v[9] = cp.kkt.total_instr;
    v[10] = cp.ldl_g;
    v[11] = cp.solve_g;
    v[12] = cp.n_contrib;
    v[13] = cp.n_work_kkt;
    v[22] = cp.n_work_fc;
    v[23] = cp.ldl_w;
    v[24] = cp.fac_steps;
    v[25] = cp.solve_steps;
    v[26] = cp.kkt_wstride;
    v[27] = cp.fc_wstride;
    v[28] = cp.n_segments;
    v[29] = cp.gen_tapes ? 1 : 0;
    v[30] = cp.kkt.n_classes;
    v[31] = cp.kkt.code_rows;
    v[14] = cp.kkt.n_components;
    v[15] = cp.fc.max_len;
    v[16] = cp.fc.total_instr;
    v[17] = cp.kkt.pre_len;
    v[18] = (int64_t)bo::coop_scratch_doubles(pr->ps, cp);
    v[19] = cp.vals_size();
  } else if (pr->sparse) {
    v[12] = pr->plan.flops;
    v[19] = pr->plan.vals_size();
  } else if (pr->team) {
    const bo::TeamPlan& tp = pr->team_plan;
    v[10] = tp.G;
    v[8] = *std::max_element(tp.kkt.cost.begin(), tp.kkt.cost.end());
    v[9] = tp.kkt.total_cost;
    v[15] = *std::max_element(tp.fc.cost.begin(), tp.fc.cost.end());
    v[16] = tp.fc.total_cost;
    v[18] = bo::team_smem_elems(pr->ps, tp.G);
    const int nk = pr->ps.nx + pr->ps.n_eq;
    v[12] = (int64_t)nk * (nk + 1) * (nk + 2) / 6;  // multiply-adds of one dense LDL'
  }
  v[20] = pr->blocks_per_sm;
  v[21] = pr->n_sm;
  if (info)
    for (int i = 0; i < cap && i < BO_TIER_INFO_LEN; ++i) info[i] = v[i];
  return BO_OK;
}

int bo_solve(bo_problem* pr, int64_t B, const double* p, const double* x0, double* x, double* lam, double* f,
             int32_t* status, int32_t* iters, double* kkt_res, void* cuda_stream) {
  if (!pr || !x || B < 0) return set_err(BO_ERR_INVALID, "bo_solve: bad argument");
  if (!pr->loaded) return set_err(BO_ERR_NO_DEVICE, "bo_solve: problem was created compile-only or without a device");
  if (pr->ps.np > 0 && !p) return set_err(BO_ERR_INVALID, "bo_solve: p is NULL but np > 0");
  if (B == 0) return BO_OK;
  CtxScope scope;
  int rc = scope.enter(pr->ctx);
  if (rc != BO_OK) return rc;
  CUstream st = (CUstream)cuda_stream;
  const size_t nx = pr->ps.nx, np = pr->ps.np, nl = pr->ps.n_eq + pr->ps.n_ineq;
  // Host buffers: the batch is cut into chunks that alternate between two internal streams, so that the upload of chunk
  // k+1 and the download of chunk k-1 run under the kernel of chunk k (instances are independent; every chunk is its own
  // persistent launch with its own work counter, and a later chunk's CTAs move in as an earlier one's run out of work).
  const bool host_call = (p && np > 0 && !is_device_ptr(p)) || (x0 && !is_device_ptr(x0)) || !is_device_ptr(x);
  int n_chunks = 1;
  // (the large / cooperative tiers index a global scratch by CTA: their launches must not overlap, so they stay one launch)
  // Measured on B200 (C2, 65536 instances, 4 chunks): 9.1 ms per call against 8.1 ms for one launch -- every chunk pays the
  // straggler tail of a persistent launch and the kernels of the two streams overlap little -- so it is opt-in.
  if (host_call && (pr->opts.flags & BO_FLAG_PIPELINE) && !(pr->large || pr->coop)) {
    const int64_t min_chunk = 16384;
    n_chunks = (int)std::min<int64_t>(8, std::max<int64_t>(1, B / min_chunk));
    if (n_chunks > 1 && !pr->pipe[0]) {
      BO_CU(g_drv.cuStreamCreate(&pr->pipe[0], CU_STREAM_NON_BLOCKING));
      BO_CU(g_drv.cuStreamCreate(&pr->pipe[1], CU_STREAM_NON_BLOCKING));
    }
  }

  CUdeviceptr dp, dx0, dx, dlam, df, dstat, dit, dkkt;
  // Page-locked host buffers are handed to the kernel as they are (zero copy): every instance reads its p / x0 row once
  // when it starts (the team tier by bulk copies a tile ahead) and writes its results once when it ends, so the PCIe
  // traffic hides under the iterations of the other instances instead of two serial copies around the launch.
  // B200OPTAS_ZERO_COPY = 0 none, 1 results only, 2 (default) inputs and results.  Pageable buffers are staged.
  static const int zero_copy_mode = [] {
    const char* e = getenv("B200OPTAS_ZERO_COPY");
    return e ? atoi(e) : 2;
  }();
  bool any_mapped = false;
  // device staging buffers for the whole batch (chunks address them by offset)
  auto stage = [&](const void* host, size_t bytes_per, DevBuf& buf, CUdeviceptr* dptr, bool* is_host, bool is_input = false) -> int {
    *dptr = 0;
    *is_host = false;
    if (!host || bytes_per == 0) return BO_OK;
    if (is_device_ptr(host)) {
      *dptr = (CUdeviceptr)(uintptr_t)host;
      return BO_OK;
    }
    if (n_chunks == 1 && zero_copy_mode >= (is_input ? 2 : 1)) {
      const CUdeviceptr m = mapped_host_ptr(host);
      if (m) {
        *dptr = m;
        any_mapped = true;
        return BO_OK;
      }
    }
    *is_host = true;
    int r = buf.reserve((size_t)B * bytes_per);
    if (r != BO_OK) return r;
    *dptr = buf.ptr;
    return BO_OK;
  };
  bool hp, hx0, hx, hlam, hf, hstat, hit, hkkt;
  if ((rc = stage(p, np * sizeof(double), pr->d_p, &dp, &hp, true)) != BO_OK) return rc;
  if ((rc = stage(x0, nx * sizeof(double), pr->d_x0, &dx0, &hx0, true)) != BO_OK) return rc;
  if ((rc = stage(x, nx * sizeof(double), pr->d_x, &dx, &hx)) != BO_OK) return rc;
  if ((rc = stage(lam, nl * sizeof(double), pr->d_lam, &dlam, &hlam)) != BO_OK) return rc;
  if ((rc = stage(f, sizeof(double), pr->d_f, &df, &hf)) != BO_OK) return rc;
  if ((rc = stage(status, sizeof(int32_t), pr->d_status, &dstat, &hstat)) != BO_OK) return rc;
  if ((rc = stage(iters, sizeof(int32_t), pr->d_iters, &dit, &hit)) != BO_OK) return rc;
  if ((rc = stage(kkt_res, sizeof(double), pr->d_kkt, &dkkt, &hkkt)) != BO_OK) return rc;
  const bool any_host = hp || hx0 || hx || hlam || hf || hstat || hit || hkkt || any_mapped;
  if (!(hp || hx0 || hx || hlam || hf || hstat || hit || hkkt)) n_chunks = 1;

  SolverParams prm{pr->opts.max_iter, pr->opts.tol, pr->opts.acceptable_tol, pr->opts.mu_init, pr->opts.max_step,
                   pr->opts.max_trips, pr->sparse ? pr->d_ldl_tab.ptr : 0, (pr->large || pr->coop) ? pr->d_dtab.ptr : 0,
                   (pr->large || pr->coop) ? pr->d_scratch.ptr : 0, (long long)pr->n_sm * pr->blocks_per_sm * pr->tpb};
  const int64_t chunk = (B + n_chunks - 1) / n_chunks;
  for (int c = 0; c < n_chunks; ++c) {
    const int64_t b0 = (int64_t)c * chunk, bc = std::min<int64_t>(B, b0 + chunk) - b0;
    if (bc <= 0) break;
    CUstream cs = n_chunks > 1 ? pr->pipe[c & 1] : st;
    auto off = [&](CUdeviceptr base, size_t bytes_per) { return base ? base + (CUdeviceptr)((size_t)b0 * bytes_per) : 0; };
    if (hp) BO_CU(g_drv.cuMemcpyHtoDAsync_v2(off(dp, np * sizeof(double)), p + (size_t)b0 * np, (size_t)bc * np * sizeof(double), cs));
    if (hx0) BO_CU(g_drv.cuMemcpyHtoDAsync_v2(off(dx0, nx * sizeof(double)), x0 + (size_t)b0 * nx, (size_t)bc * nx * sizeof(double), cs));
    long long Bll = bc;
    CUdeviceptr dcounter = pr->d_counter.ptr + (CUdeviceptr)((c % 16) * sizeof(unsigned long long));
    BO_CU(g_drv.cuMemsetD8Async(dcounter, 0, sizeof(unsigned long long), cs));
    long long grid_ll = (long long)pr->n_sm * pr->blocks_per_sm;
    // coop: one instance per CTA; team: 32 instances per CTA; else one per thread
    const long long need = pr->coop ? (long long)bc : (pr->team ? (bc + 31) / 32 : (bc + pr->tpb - 1) / pr->tpb);
    if (grid_ll > need) grid_ll = need;
    const unsigned grid = (unsigned)grid_ll;
    CUdeviceptr a_p = off(dp, np * sizeof(double)), a_x0 = off(dx0, nx * sizeof(double)), a_x = off(dx, nx * sizeof(double)),
                a_lam = off(dlam, nl * sizeof(double)), a_f = off(df, sizeof(double)), a_st = off(dstat, sizeof(int32_t)),
                a_it = off(dit, sizeof(int32_t)), a_kkt = off(dkkt, sizeof(double));
    void* args[] = {&Bll, &a_p, &a_x0, &a_x, &a_lam, &a_f, &a_st, &a_it, &a_kkt, &dcounter, &prm};
    size_t slot = 0;
    if ((rc = pr->timer.begin(cs, &slot)) != BO_OK) return rc;
    BO_CU(g_drv.cuLaunchKernel(pr->kernel.fn, grid, 1, 1, (unsigned)pr->tpb, 1, 1, (unsigned)pr->smem_dynamic, cs, args, nullptr));
    if ((rc = pr->timer.end(cs, slot)) != BO_OK) return rc;
    if (hx) BO_CU(g_drv.cuMemcpyDtoHAsync_v2(x + (size_t)b0 * nx, a_x, (size_t)bc * nx * sizeof(double), cs));
    if (hlam) BO_CU(g_drv.cuMemcpyDtoHAsync_v2(lam + (size_t)b0 * nl, a_lam, (size_t)bc * nl * sizeof(double), cs));
    if (hf) BO_CU(g_drv.cuMemcpyDtoHAsync_v2(f + b0, a_f, (size_t)bc * sizeof(double), cs));
    if (hstat) BO_CU(g_drv.cuMemcpyDtoHAsync_v2(status + b0, a_st, (size_t)bc * sizeof(int32_t), cs));
    if (hit) BO_CU(g_drv.cuMemcpyDtoHAsync_v2(iters + b0, a_it, (size_t)bc * sizeof(int32_t), cs));
    if (hkkt) BO_CU(g_drv.cuMemcpyDtoHAsync_v2(kkt_res + b0, a_kkt, (size_t)bc * sizeof(double), cs));
  }
  if (any_host) {
    if (n_chunks > 1) {
      BO_CU(g_drv.cuStreamSynchronize(pr->pipe[0]));
      BO_CU(g_drv.cuStreamSynchronize(pr->pipe[1]));
    } else {
      BO_CU(g_drv.cuStreamSynchronize(st));
    }
  }
  return BO_OK;
}

int bo_problem_kernel_time(bo_problem* pr, double* ms_total, int64_t* n_launches) {
  if (!pr || !ms_total) return set_err(BO_ERR_INVALID, "null argument");
  if (!pr->loaded) return set_err(BO_ERR_NO_DEVICE, "problem not loaded on a device");
  CtxScope scope;
  scope.enter(pr->ctx);
  int64_t n = 0;
  int rc = pr->timer.collect(ms_total, &n);
  if (n_launches) *n_launches = n;
  return rc;
}

// ------------------------------------------------------------------------------------------
// host-side marshalling: label arrays -> [B][total] rows (batched dict2vec)
// ------------------------------------------------------------------------------------------
// Worker threads of bo_pack_rows, started on first use and kept: spawning 16 threads per call cost more than the copy
// (0.3 of 0.47 ms for 65536 x 10 doubles).  The caller takes part in the work.  A forked child starts its own pool.
namespace {
struct PackPool {
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  std::vector<std::thread>* workers = new std::vector<std::thread>;
  const std::function<void(int)>* job = nullptr;
  int n_jobs = 0, next = 0, pending = 0;
  uint64_t generation = 0;
  pid_t owner = 0;
  void worker() {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(m);
    for (;;) {
      cv_work.wait(lk, [&] { return generation != seen; });
      seen = generation;
      drain(lk);
    }
  }
  void drain(std::unique_lock<std::mutex>& lk) {  // called with the lock held
    while (next < n_jobs) {
      const int t = next++;
      const std::function<void(int)>* f = job;
      lk.unlock();
      (*f)(t);
      lk.lock();
      if (--pending == 0) cv_done.notify_all();
    }
  }
  void run(int n, const std::function<void(int)>& f) {
    std::unique_lock<std::mutex> lk(m);
    if (owner != getpid()) {  // first use, or first use after fork(): the parent's threads do not exist here
      if (owner != 0) workers = new std::vector<std::thread>;  // the old handles name threads of the parent: leave them alone
      owner = getpid();
    }
    const int want = std::min(n - 1, 15);
    while ((int)workers->size() < want) workers->emplace_back([this] { worker(); });
    job = &f;
    n_jobs = n;
    next = 0;
    pending = n;
    ++generation;
    cv_work.notify_all();
    drain(lk);
    cv_done.wait(lk, [&] { return pending == 0; });
    job = nullptr;
    n_jobs = 0;
  }
};
PackPool& pack_pool() {
  static PackPool* pool = new PackPool;  // never destroyed: its threads may outlive static destruction
  return *pool;
}
}  // namespace

int bo_pack_rows(double* dst, int64_t B, int64_t total, int32_t n_seg, const double* const* src, const int64_t* stride,
                 const int32_t* m, const int32_t* n, const int64_t* off, int32_t n_threads) {
  if (!dst || B < 0 || total < 0 || n_seg < 0 || (n_seg > 0 && (!src || !stride || !m || !n || !off)))
    return set_err(BO_ERR_INVALID, "bo_pack_rows: bad argument");
  for (int k = 0; k < n_seg; ++k)
    if (m[k] < 0 || n[k] < 0 || off[k] < 0 || off[k] + (int64_t)m[k] * n[k] > total)
      return set_err(BO_ERR_INVALID, "bo_pack_rows: segment %d does not fit a row of %lld", k, (long long)total);
  if (B == 0 || total == 0) return BO_OK;
  auto work = [&](int64_t b0, int64_t b1) {
    for (int64_t b = b0; b < b1; ++b) {
      double* row = dst + b * total;
      for (int k = 0; k < n_seg; ++k) {
        const int mk = m[k], nk = n[k];
        double* out = row + off[k];
        if (!src[k]) {
          for (int e = 0; e < mk * nk; ++e) out[e] = 0.0;
          continue;
        }
        const int64_t sb = stride[3 * k], sr = stride[3 * k + 1], sc = stride[3 * k + 2];
        const double* in = src[k] + b * sb;
        if (sr == 1 && (nk == 1 || sc == mk)) {  // the instance's block already lies in vec() order: straight copy
          const int len = mk * nk;
          for (int e = 0; e < len; ++e) out[e] = in[e];
          continue;
        }
        for (int c = 0; c < nk; ++c)  // column-major flattening of the instance's matrix
          for (int r = 0; r < mk; ++r) out[c * mk + r] = in[r * sr + c * sc];
      }
    }
  };
  int nt = n_threads > 0 ? n_threads : (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
  if ((int64_t)nt > B / 4096) nt = (int)std::max<int64_t>(1, B / 4096);  // a thread is not worth less than 4096 rows
  if (nt <= 1) {
    work(0, B);
    return BO_OK;
  }
  const int64_t chunk = (B + nt - 1) / nt;
  pack_pool().run(nt, [&](int t) {
    const int64_t b0 = t * chunk, b1 = std::min<int64_t>(B, b0 + chunk);
    if (b0 < b1) work(b0, b1);
  });
  return BO_OK;
}

// ------------------------------------------------------------------------------------------
// streaming function evaluation
// ------------------------------------------------------------------------------------------
int bo_function_create(const bo_tape* tape, const bo_options* opts_in, bo_function** out) {
  if (!tape || !out) return set_err(BO_ERR_INVALID, "bo_function_create: null argument");
  *out = nullptr;
  std::unique_ptr<bo_function> fn(new bo_function);
  fn->opts = normalise(opts_in);
  if (fn->opts.cache_dir) { fn->cache_dir = fn->opts.cache_dir; fn->opts.cache_dir = fn->cache_dir.c_str(); }
  if (fn->opts.include_dir) { fn->include_dir = fn->opts.include_dir; fn->opts.include_dir = fn->include_dir.c_str(); }
  std::string err;
  if (!bo::copy_tape(*tape, &fn->tape, &err)) return set_err(BO_ERR_INVALID, "bo_function_create: %s", err.c_str());
  if (fn->tape.in_sizes.empty() || fn->tape.out_sizes.empty())
    return set_err(BO_ERR_INVALID, "bo_function_create: need at least one input and one output segment");
  if (fn->tape.in_sizes.size() > 16 || fn->tape.out_sizes.size() > 16)
    return set_err(BO_ERR_UNSUPPORTED, "bo_function_create: at most 16 input and 16 output segments");
  size_t per_instance = 0;
  for (int s : fn->tape.in_sizes) per_instance += (size_t)s;
  for (int s : fn->tape.out_sizes) per_instance += (size_t)s;
  // the pipeline stages of BO_TPB instances must fit in shared memory (227 KB per CTA on sm_100): inputs are always
  // double-buffered; outputs double-buffered (default) or single-buffered (B200OPTAS_K1_OUT_STAGES=1: less shared
  // memory per CTA, more resident CTAs per SM; the stores of a tile then drain before the next tile computes)
  size_t in_doubles = 0, out_doubles = 0;
  for (int s : fn->tape.in_sizes) in_doubles += (size_t)s;
  for (int s : fn->tape.out_sizes) out_doubles += (size_t)s;
  const char* os_env = getenv("B200OPTAS_K1_OUT_STAGES");
  const int out_stages = (os_env && atoi(os_env) == 1) ? 1 : 2;
  auto smem_for = [&](int t) { return (size_t)t * (2 * in_doubles + (size_t)out_stages * out_doubles) * sizeof(double) + 64; };
  int tpb = fn->opts.threads_per_block > 0 ? fn->opts.threads_per_block : 128;
  while (tpb > 32 && smem_for(tpb) > 200 * 1024) tpb /= 2;
  if (smem_for(tpb) > 220 * 1024)
    return set_err(BO_ERR_UNSUPPORTED, "bo_function_create: %zu doubles per instance do not fit the shared-memory pipeline",
                   per_instance);
  fn->tpb = tpb;
  fn->smem_dynamic = (int)smem_for(tpb);
  fn->source = bo::emit_function_source(fn->tape, tpb, out_stages);
  int rc = jit_compile(fn->source, "bo_eval", "bo_eval_kernel", fn->opts, &fn->compiled);
  if (rc != BO_OK) return rc;
  fn->kernel.regs = fn->compiled.regs;
  fn->kernel.local_bytes = fn->compiled.local_bytes;
  fn->kernel.smem_bytes = fn->compiled.smem_bytes;
  if (!(fn->opts.flags & BO_FLAG_COMPILE_ONLY)) {
    CUdevice dev;
    rc = ensure_context(&dev, fn->opts.device - 1, &fn->ctx);
    if (rc != BO_OK) return rc;
    rc = load_kernel(fn->compiled, "bo_eval_kernel", &fn->kernel);
    if (rc != BO_OK) return rc;
    BO_CU(g_drv.cuFuncSetAttribute(fn->kernel.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, fn->smem_dynamic));
    BO_CU(g_drv.cuDeviceGetAttribute(&fn->n_sm, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
    BO_CU(g_drv.cuOccupancyMaxActiveBlocksPerMultiprocessor(&fn->blocks_per_sm, fn->kernel.fn, tpb, fn->smem_dynamic));
    if (fn->blocks_per_sm < 1) fn->blocks_per_sm = 1;
    fn->d_in.resize(fn->tape.in_sizes.size());
    fn->d_out.resize(fn->tape.out_sizes.size());
    fn->loaded = true;
    fn->timer.enabled = (fn->opts.flags & BO_FLAG_TIMING) != 0;
  }
  *out = fn.release();
  return BO_OK;
}

int bo_function_destroy(bo_function* fn) {
  if (!fn) return BO_OK;
  if (fn->loaded) {
    CtxScope scope;
    scope.enter(fn->ctx);
    for (auto& b : fn->d_in) b.release();
    for (auto& b : fn->d_out) b.release();
    fn->timer.release();
    if (fn->kernel.mod) g_drv.cuModuleUnload(fn->kernel.mod);
  }
  delete fn;
  return BO_OK;
}

int64_t bo_function_source(const bo_function* fn, char* buf, int64_t cap) {
  if (!fn) return set_err(BO_ERR_INVALID, "null function");
  return copy_out(fn->source, buf, cap);
}

int bo_function_kernel_info(const bo_function* fn, int32_t* regs, int32_t* local_bytes, int32_t* smem_bytes) {
  if (!fn) return set_err(BO_ERR_INVALID, "null function");
  if (regs) *regs = fn->kernel.regs;
  if (local_bytes) *local_bytes = fn->kernel.local_bytes;
  if (smem_bytes) *smem_bytes = fn->loaded ? fn->kernel.smem_bytes + fn->smem_dynamic : fn->smem_dynamic;
  return BO_OK;
}

int bo_function_eval(bo_function* fn, int64_t B, const double* const* in, double* const* out, void* cuda_stream) {
  if (!fn || !in || !out || B < 0) return set_err(BO_ERR_INVALID, "bo_function_eval: bad argument");
  if (!fn->loaded) return set_err(BO_ERR_NO_DEVICE, "bo_function_eval: function was created compile-only or without a device");
  if (B == 0) return BO_OK;
  CtxScope scope;
  int rc = scope.enter(fn->ctx);
  if (rc != BO_OK) return rc;
  CUstream st = (CUstream)cuda_stream;
  const size_t n_in = fn->tape.in_sizes.size(), n_out = fn->tape.out_sizes.size();
  // kernel argument block: must match bo_eval_args in csrc/jit/bo_stream_eval.cuh
  std::vector<unsigned char> argbuf((n_in + n_out) * sizeof(void*) + sizeof(int) + 8, 0);
  CUdeviceptr* ptrs = reinterpret_cast<CUdeviceptr*>(argbuf.data());
  bool any_host = false, aligned = true;
  struct Out {
    void* host;
    CUdeviceptr dev;
    size_t bytes;
  };
  std::vector<Out> outs;
  for (size_t k = 0; k < n_in; ++k) {
    const size_t bytes = (size_t)B * fn->tape.in_sizes[k] * sizeof(double);
    if (bytes == 0) continue;
    if (!in[k]) return set_err(BO_ERR_INVALID, "bo_function_eval: input %zu is NULL", k);
    if (is_device_ptr(in[k])) {
      ptrs[k] = (CUdeviceptr)(uintptr_t)in[k];
    } else {
      any_host = true;
      if ((rc = fn->d_in[k].reserve(bytes)) != BO_OK) return rc;
      BO_CU(g_drv.cuMemcpyHtoDAsync_v2(fn->d_in[k].ptr, in[k], bytes, st));
      ptrs[k] = fn->d_in[k].ptr;
    }
    if (ptrs[k] & 15) aligned = false;
  }
  for (size_t k = 0; k < n_out; ++k) {
    const size_t bytes = (size_t)B * fn->tape.out_sizes[k] * sizeof(double);
    if (bytes == 0) continue;
    if (!out[k]) return set_err(BO_ERR_INVALID, "bo_function_eval: output %zu is NULL", k);
    if (is_device_ptr(out[k])) {
      ptrs[n_in + k] = (CUdeviceptr)(uintptr_t)out[k];
    } else {
      any_host = true;
      if ((rc = fn->d_out[k].reserve(bytes)) != BO_OK) return rc;
      ptrs[n_in + k] = fn->d_out[k].ptr;
      outs.push_back({out[k], fn->d_out[k].ptr, bytes});
    }
    if (ptrs[n_in + k] & 15) aligned = false;
  }
  *reinterpret_cast<int*>(argbuf.data() + (n_in + n_out) * sizeof(void*)) = aligned ? 1 : 0;

  long long Bll = B;
  void* args[] = {&Bll, argbuf.data()};
  const long long n_tiles = (B + fn->tpb - 1) / fn->tpb;
  long long grid = (long long)fn->n_sm * fn->blocks_per_sm;
  if (grid > n_tiles) grid = n_tiles;
  size_t slot = 0;
  if ((rc = fn->timer.begin(st, &slot)) != BO_OK) return rc;
  BO_CU(g_drv.cuLaunchKernel(fn->kernel.fn, (unsigned)grid, 1, 1, (unsigned)fn->tpb, 1, 1, (unsigned)fn->smem_dynamic, st, args,
                             nullptr));
  if ((rc = fn->timer.end(st, slot)) != BO_OK) return rc;
  for (const Out& o : outs) BO_CU(g_drv.cuMemcpyDtoHAsync_v2(o.host, o.dev, o.bytes, st));
  if (any_host) BO_CU(g_drv.cuStreamSynchronize(st));
  return BO_OK;
}

int bo_function_kernel_time(bo_function* fn, double* ms_total, int64_t* n_launches) {
  if (!fn || !ms_total) return set_err(BO_ERR_INVALID, "null argument");
  if (!fn->loaded) return set_err(BO_ERR_NO_DEVICE, "function not loaded on a device");
  CtxScope scope;
  scope.enter(fn->ctx);
  int64_t n = 0;
  int rc = fn->timer.collect(ms_total, &n);
  if (n_launches) *n_launches = n;
  return rc;
}

}  // extern "C"
