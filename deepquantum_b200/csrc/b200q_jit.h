// Run-time compilation of the per-pass kernels (b200q_codegen) with NVRTC for sm_100a, and their launch.
// NVRTC is loaded with dlopen (the library must load on a machine without it, e.g. the CPU-only build container);
// cubins are cached in memory per source text and on disk next to the library (lib/jit_cache/<hash>.cubin), so a
// circuit structure is compiled once per installation.  The kernels are loaded through the CUDA runtime's
// cudaLibrary API (context independent, no driver-API linkage).
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

#include "b200q_planner.h"

namespace b200q {

struct JitKernel {
  std::string source, log, hash;
  std::vector<char> cubin;
  size_t smem = 0;
  int threads = 0, min_blocks = 2;
  cudaLibrary_t lib = nullptr;
  cudaKernel_t fn = nullptr;
  bool compiled = false, loaded = false, failed = false;
};

// per plan: one kernel per pass, plus the fused-exchange variant of the last pass
struct PlanJit {
  std::vector<std::shared_ptr<JitKernel>> local, remote;
  bool prepared = false;
  int n_ok = 0, n_failed = 0;
};

// 0 = usable.  Otherwise the reason is in *why (NVRTC missing, disabled by B200Q_JIT=0, ...).
int jit_available(std::string* why);
// Generates and compiles (in parallel, `threads` <= 0: all host cores) the kernels of every pass the generator
// supports.  Compilation needs no GPU.  Returns the number of passes that have a kernel.
int jit_prepare(Plan& plan, int threads, bool with_remote_last);
// Launches the specialised kernel of pass i; returns -1000 if there is none (the caller falls back to the generic
// tile kernel), 0 on success, a CUDA error code otherwise.
int jit_launch(Plan& plan, int pass_index, void* state, const void* mats, int64_t batch, int64_t mbs,
               cudaStream_t stream, const b200q_remote_t* remote);
// Compile one source text to a cubin (used by jit_prepare and the tests); returns false and fills `log` on error.
bool jit_compile_source(const std::string& source, std::vector<char>* cubin, std::string* log);
std::string jit_cache_dir();

}  // namespace b200q
