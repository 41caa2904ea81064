// See b200q_jit.h.
#include "b200q_jit.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>

#include "b200q_codegen.h"

namespace b200q {
namespace {

// ---- NVRTC through dlopen ----------------------------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* h = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*Version)(int*, int*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string why;
  int major = 0, minor = 0;
};

Nvrtc& nvrtc() {
  static Nvrtc N;
  static std::once_flag once;
  std::call_once(once, [] {
    if (const char* e = getenv("B200Q_JIT")) {
      if (atoi(e) == 0) { N.why = "disabled by B200Q_JIT=0"; return; }
    }
    const char* names[] = {getenv("B200Q_NVRTC"), "libnvrtc.so.12", "libnvrtc.so",
                           "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      N.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (N.h) break;
    }
    if (!N.h) { N.why = "libnvrtc not found (dlopen)"; return; }
#define B200Q_SYM(field, name)                                        \
  N.field = reinterpret_cast<decltype(N.field)>(dlsym(N.h, name));    \
  if (!N.field) { N.why = std::string("libnvrtc lacks ") + name; N.h = nullptr; return; }
    B200Q_SYM(CreateProgram, "nvrtcCreateProgram")
    B200Q_SYM(DestroyProgram, "nvrtcDestroyProgram")
    B200Q_SYM(CompileProgram, "nvrtcCompileProgram")
    B200Q_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    B200Q_SYM(GetCUBIN, "nvrtcGetCUBIN")
    B200Q_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    B200Q_SYM(GetProgramLog, "nvrtcGetProgramLog")
    B200Q_SYM(Version, "nvrtcVersion")
    B200Q_SYM(GetErrorString, "nvrtcGetErrorString")
#undef B200Q_SYM
    N.Version(&N.major, &N.minor);
  });
  return N;
}

const char* kArch = "--gpu-architecture=sm_100a";

std::string hash_text(const std::string& s) {   // 128-bit FNV-style digest, hex
  uint64_t a = 0xcbf29ce484222325ull, b = 0x9ae16a3b2f90404full;
  for (unsigned char c : s) {
    a = (a ^ c) * 0x100000001b3ull;
    b = (b + c) * 0xff51afd7ed558ccdull;
    b ^= b >> 29;
  }
  char buf[40];
  snprintf(buf, sizeof buf, "%016llx%016llx", (unsigned long long)a, (unsigned long long)b);
  return buf;
}

std::mutex g_mem_mu;
std::map<std::string, std::shared_ptr<JitKernel>> g_mem_cache;   // by hash of (source, options)

bool read_file(const std::string& path, std::vector<char>* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize(n > 0 ? size_t(n) : 0);
  const bool ok = n > 0 && fread(out->data(), 1, size_t(n), f) == size_t(n);
  fclose(f);
  return ok;
}

void write_file_atomic(const std::string& path, const std::vector<char>& data) {
  const std::string tmp = path + ".tmp" + std::to_string((long)getpid()) + "_" +
                          std::to_string((unsigned long long)(uintptr_t)&data);
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  fclose(f);
  if (ok) rename(tmp.c_str(), path.c_str());
  else unlink(tmp.c_str());
}

}  // namespace

std::string jit_cache_dir() {
  static std::string dir = [] {
    std::string d;
    if (const char* e = getenv("B200Q_JIT_CACHE")) d = e;
    if (d.empty()) {
      Dl_info info;
      if (dladdr((void*)&jit_cache_dir, &info) && info.dli_fname) {
        d = info.dli_fname;
        const size_t p = d.find_last_of('/');
        d = (p == std::string::npos ? std::string(".") : d.substr(0, p)) + "/jit_cache";
      } else d = "/tmp/b200q_jit_cache";
    }
    mkdir(d.c_str(), 0755);
    return d;
  }();
  return dir;
}

int jit_available(std::string* why) {
  Nvrtc& N = nvrtc();
  if (!N.h) {
    if (why) *why = N.why;
    return -1;
  }
  return 0;
}

bool jit_compile_source(const std::string& source, std::vector<char>* cubin, std::string* log) {
  Nvrtc& N = nvrtc();
  if (!N.h) {
    if (log) *log = N.why;
    return false;
  }
  nvrtcProgram prog = nullptr;
  int rc = N.CreateProgram(&prog, source.c_str(), "b200qj_pass.cu", 0, nullptr, nullptr);
  if (rc) {
    if (log) *log = std::string("nvrtcCreateProgram: ") + N.GetErrorString(rc);
    return false;
  }
  const char* opts[] = {kArch, "--std=c++17", "-lineinfo", "-default-device", "-w"};
  rc = N.CompileProgram(prog, 5, opts);
  size_t ls = 0;
  N.GetProgramLogSize(prog, &ls);
  if (log && ls > 1) {
    log->resize(ls);
    N.GetProgramLog(prog, &(*log)[0]);
  }
  bool ok = rc == 0;
  if (ok) {
    size_t cs = 0;
    ok = N.GetCUBINSize(prog, &cs) == 0 && cs > 0;
    if (ok) {
      cubin->resize(cs);
      ok = N.GetCUBIN(prog, cubin->data()) == 0;
    }
  } else if (log) {
    *log = std::string("nvrtcCompileProgram: ") + N.GetErrorString(rc) + "\n" + *log;
  }
  N.DestroyProgram(&prog);
  return ok;
}

namespace {

// compile (or fetch from the caches) the kernel for one source text
std::shared_ptr<JitKernel> get_kernel(const std::string& source, size_t smem, int threads, int min_blocks) {
  Nvrtc& N = nvrtc();
  const std::string key = hash_text(source + "|" + kArch + "|" + std::to_string(N.major) + "." + std::to_string(N.minor));
  {
    std::lock_guard<std::mutex> lk(g_mem_mu);
    auto it = g_mem_cache.find(key);
    if (it != g_mem_cache.end()) return it->second;
  }
  auto k = std::make_shared<JitKernel>();
  k->hash = key;
  k->smem = smem;
  k->threads = threads;
  k->min_blocks = min_blocks;
  const std::string path = jit_cache_dir() + "/" + key + ".cubin";
  if (read_file(path, &k->cubin)) {
    k->compiled = true;
  } else {
    k->compiled = jit_compile_source(source, &k->cubin, &k->log);
    if (k->compiled) write_file_atomic(path, k->cubin);
    else {
      k->failed = true;
      if (getenv("B200Q_JIT_VERBOSE")) fprintf(stderr, "[b200q jit] compile failed:\n%s\n", k->log.c_str());
    }
  }
  if (getenv("B200Q_JIT_KEEP_SOURCE")) k->source = source;
  std::lock_guard<std::mutex> lk(g_mem_mu);
  auto it = g_mem_cache.find(key);
  if (it != g_mem_cache.end()) return it->second;
  g_mem_cache[key] = k;
  return k;
}

std::mutex g_load_mu;

int ensure_loaded(JitKernel& k) {
  if (k.loaded) return 0;
  std::lock_guard<std::mutex> lk(g_load_mu);
  if (k.loaded) return 0;
  if (!k.compiled) return -1000;
  cudaError_t e = cudaLibraryLoadData(&k.lib, k.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&k.fn, k.lib, "b200qj_pass");
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute((const void*)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem);
  if (e != cudaSuccess) {
    cudaGetLastError();
    k.failed = true;
    k.compiled = false;
    k.log = std::string("load: ") + cudaGetErrorString(e);
    if (getenv("B200Q_JIT_VERBOSE")) fprintf(stderr, "[b200q jit] %s\n", k.log.c_str());
    return -1000;
  }
  k.loaded = true;
  return 0;
}

int sm_count_jit() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev] = v;
  }
  return cache[dev];
}

}  // namespace

int jit_prepare(Plan& plan, int threads, bool with_remote_last) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!plan.jit) plan.jit = std::make_shared<PlanJit>();
  PlanJit& J = *plan.jit;
  std::string why;
  const int np = (int)plan.passes.size();
  J.local.resize(np);
  J.remote.resize(np);
  if (jit_available(&why)) {
    J.prepared = true;
    return 0;
  }
  struct Job { int pass; bool remote; };
  std::vector<Job> jobs;
  for (int i = 0; i < np; ++i) {
    if (!codegen_supported(plan, plan.passes[i])) continue;
    if (!J.local[i]) jobs.push_back({i, false});
    if (with_remote_last && i == np - 1 && !J.remote[i]) jobs.push_back({i, true});
  }
  if (!jobs.empty()) {
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("B200Q_JIT_THREADS")) nt = atoi(e);
    nt = std::max(1, std::min<int>(nt, (int)jobs.size()));
    std::atomic<int> next(0);
    auto worker = [&] {
      for (;;) {
        const int j = next.fetch_add(1);
        if (j >= (int)jobs.size()) break;
        GenOptions go;
        go.remote = jobs[j].remote ? 1 : 0;
        go.min_blocks = plan.opt.chunk_bits >= 13 ? 1 : (plan.opt.chunk_bits == 12 ? 2 : 4);
        if (const char* e = getenv("B200Q_JIT_IPT")) go.items_per_thread = atoi(e);
        if (go.items_per_thread == 2 && plan.opt.chunk_bits == 12) go.min_blocks = 3;
        if (go.items_per_thread == 2 && plan.opt.chunk_bits == 13) go.min_blocks = 1;
        if (const char* e = getenv("B200Q_JIT_MIN_BLOCKS")) go.min_blocks = atoi(e);   // A/B: CTAs per SM the compiler must fit
        if (const char* e = getenv("B200Q_JIT_PREFETCH")) go.prefetch = atoi(e);
        if (const char* e = getenv("B200Q_JIT_DEBUG_SKIP_OPS")) go.debug_skip_ops = atoi(e);
        if (const char* e = getenv("B200Q_JIT_DEBUG_ONE_TILE")) go.debug_one_tile = atoi(e);
        size_t smem = 0;
        const std::string src = codegen_pass(plan, plan.passes[jobs[j].pass], go, &smem);
        auto k = get_kernel(src, smem, (1 << (plan.opt.chunk_bits - B200Q_REG_CHUNK_BITS)) / std::max(1, go.items_per_thread),
                            go.min_blocks);
        (jobs[j].remote ? J.remote : J.local)[jobs[j].pass] = k;
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
  }
  J.n_ok = J.n_failed = 0;
  for (int i = 0; i < np; ++i) {
    if (J.local[i] && J.local[i]->compiled) ++J.n_ok;
    else if (J.local[i]) ++J.n_failed;
  }
  J.prepared = true;
  return J.n_ok;
}

int jit_launch(Plan& plan, int i, void* state, const void* mats, int64_t batch, int64_t mbs, cudaStream_t stream,
               const b200q_remote_t* remote_in) {
  if (!plan.jit || !plan.jit->prepared) return -1000;
  PlanJit& J = *plan.jit;
  const bool rem = remote_in && remote_in->enabled;
  if (i < 0 || i >= (int)J.local.size()) return -1000;
  JitKernel* k = (rem ? J.remote[i] : J.local[i]).get();
  if (!k || !k->compiled) return -1000;
  if (ensure_loaded(*k)) return -1000;
  const b200q_pass_t& P = plan.passes[i];
  const int vs = plan.dtype == B200Q_C64 ? 1 : 0;
  b200q_remote_t remote;
  std::memset(&remote, 0, sizeof remote);
  if (rem) remote = *remote_in;
  uint64_t chunks_per_state = (1ull << plan.n_qubits) >> vs;
  uint32_t tile_shift = uint32_t(int(P.n_bits) - int(P.tile_bits));
  const uint64_t ntiles = 1ull << tile_shift;
  static const int ctas_env = [] { const char* e = getenv("B200Q_JIT_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  static const int oversub_env = [] { const char* e = getenv("B200Q_JIT_OVERSUB"); return e ? atoi(e) : 0; }();
  // Grid: a multiple of the resident CTA count, oversubscribed (8x: the hardware hands out CTAs as SMs free up --
  // measured 42.9 ms against 44.8 ms with exactly the resident count on the 28-qubit circuit), but never so far
  // that a CTA walks fewer than 16 tiles: its prologue (coefficient table from the matrix buffer) must stay a few
  // per cent of its life (small shards of the sharded path: 2^27 amplitudes are 16 384 tiles)
  int oversub = oversub_env > 0 ? oversub_env : (P.n_ops <= 4 ? 16 : 8);
  if (oversub_env <= 0) {
    const uint64_t base = uint64_t(sm_count_jit()) * k->min_blocks;
    while (oversub > 1 && ntiles * uint64_t(batch) < base * uint64_t(oversub) * 16ull) oversub >>= 1;
  }
  const uint64_t resident = uint64_t(sm_count_jit()) * (ctas_env > 0 ? ctas_env : k->min_blocks * oversub);
  const void* state_p = state;
  const void* mats_p = mats;
  int64_t mbs_v = mbs;
  uint64_t n_work;
  dim3 grid;
  if (mbs == 0) {
    n_work = ntiles * uint64_t(batch);
    grid = dim3((unsigned)std::min<uint64_t>(n_work, resident), 1, 1);
    void* args[] = {&state_p, &mats_p, &chunks_per_state, &mbs_v, &tile_shift, &n_work, &remote};
    cudaError_t e = cudaLaunchKernel((const void*)k->fn, grid, dim3(k->threads), args, k->smem, stream);
    if (e != cudaSuccess) return (int)e;
  } else {
    n_work = ntiles;
    const uint64_t gx = std::min<uint64_t>(ntiles, std::max<uint64_t>(1, resident / uint64_t(std::min<int64_t>(batch, (int64_t)resident))));
    const size_t csz = 16;
    for (int64_t b0 = 0; b0 < batch; b0 += 32768) {
      const int64_t nb = std::min<int64_t>(32768, batch - b0);
      grid = dim3((unsigned)gx, (unsigned)nb, 1);
      state_p = (const char*)state + uint64_t(b0) * chunks_per_state * csz;
      mats_p = (const char*)mats + size_t(b0) * size_t(mbs) * (vs ? 8 : 16);
      void* args[] = {&state_p, &mats_p, &chunks_per_state, &mbs_v, &tile_shift, &n_work, &remote};
      cudaError_t e = cudaLaunchKernel((const void*)k->fn, grid, dim3(k->threads), args, k->smem, stream);
      if (e != cudaSuccess) return (int)e;
    }
  }
  return 0;
}

}  // namespace b200q
