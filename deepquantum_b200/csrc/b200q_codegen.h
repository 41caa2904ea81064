// Per-pass kernel specialisation (host C++, no CUDA): turns ONE pass descriptor of a plan into the CUDA source of a
// kernel in which everything structural -- op sequence, register slots, control masks, tile geometry, address bit
// deposits -- is a compile-time constant, so the interpretive op loop of b200q_tile_kernel (op-word load, control
// tests, branch tree, coefficient-record select per op per round per tile) disappears.  The source is compiled at run
// time with NVRTC for sm_100a (b200q_jit.cpp) and, TEST-ONLY, with g++ (the same text has host definitions of the
// packed-FP32 primitives) so that tests/ can step it thread by thread against the oracle without a GPU.
//
// Replaces, like the tile kernel: qmath.evolve_state (qmath.py:485-506) and Gate.op_state_control
// (operation.py:203-219) for every gate of a fused group.
#pragma once
#include <string>

#include "b200q_planner.h"

namespace b200q {

struct GenOptions {
  int remote = 0;      // 1: the write-back round is the fused exchange scatter (b200q_remote_t)
  int min_blocks = 2;  // __launch_bounds__ second argument
  int debug_skip_ops = 0, debug_one_tile = 0;   // measurement switches (B200Q_JIT_DEBUG_*): never set in production
  int items_per_thread = 1;   // 2: half the threads per tile, each walks two items per round
  int prefetch = 0;    // L2 prefetch of the CTA's next tile while the current one is computed
};

// True if the generator covers this pass (full-size tile, un-padded state).  Small states keep the generic kernel:
// they are launch-latency bound and not worth a compilation.
bool codegen_supported(const Plan& plan, const b200q_pass_t& P);

// CUDA/C++ source of the kernel `b200qj_pass` for pass P.  `smem_bytes` receives the dynamic shared memory size.
std::string codegen_pass(const Plan& plan, const b200q_pass_t& P, const GenOptions& opt, size_t* smem_bytes,
                         std::string* stats = nullptr);

}  // namespace b200q
