// Host-side fusion planner: turns an ordered gate list into a list of passes (b200q_program.h).
// Pure host C++ (no CUDA), so it is unit-tested on CPU.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200q.h"
#include "b200q_program.h"

namespace b200q {

struct PlanOptions {
  int chunk_bits = 12;    // log2(16-byte chunks per tile): 12 -> 64 KiB tiles, 256 threads per CTA
  int low_bits = -1;      // contiguous low amplitude bits forced into every tile (-1: default per dtype)
  int max_rounds = B200Q_MAX_ROUNDS;
  int max_ops = B200Q_MAX_OPS;
  int structured = 1;     // Hadamard / rotation hints select the in-place add-sub and three-shear butterflies
  int coalesce_bits = 1;  // chunk-index bits that stay lane bits in the rounds that touch global memory (1: every
                          // warp-level access uses whole 32-byte sectors; measured equal to 3 = 128-byte runs, with
                          // fewer rounds: a single gate on a low qubit is then ONE round at the copy bandwidth)
  int defer_diag = 2;     // 1: leave register-slot diagonals pending when nothing in the round depends on them;
                          // 2: and never spend a round on diagonal gates alone (measured +4.6 % on C2)
  int min_round_gates = 4;  // a later round with fewer executable gates ends the pass (1: off; measured +2 % on C2)
  int xc1_penalty = 0;    // slot choice: penalty (in quarter gates) per CNOT whose control becomes a register slot
                          // (measured neutral on C2: fewer register swaps, more rounds; off)
  int free_phase = 0;     // plans run by the specialised kernels: a hinted diag(1, i^q) gate costs nothing when its
                          // qubit is a register slot (renaming / frame bit) but a thread phase plus its application
                          // otherwise, so such gates wait for a round that holds their qubit in registers
  int fuse = 1;           // 0: one gate per pass (the un-fused baseline used for A/B measurements)
};

struct PlanJit;   // run-time specialised kernels of the passes (b200q_jit.h); never touched by the planner

struct PlanStats {
  int n_gates = 0, n_passes = 0, n_rounds = 0, n_ops = 0, n_direct = 0;
};

class Plan {
 public:
  int n_qubits = 0;       // physical local qubits
  int n_bits = 0;         // padded index bits (>= register slots)
  int dtype = 0;          // B200Q_C64 / B200Q_C128
  PlanOptions opt;
  std::vector<b200q_pass_t> passes;
  std::vector<int> pass_gate_count;
  PlanStats stats;
  std::string error;
  std::shared_ptr<PlanJit> jit;
};

// Returns nullptr and fills `err` on invalid input.
Plan* make_plan(int n_qubits, int dtype, const b200q_gate_t* gates, int n_gates, const PlanOptions& opt,
                std::string* err);

}  // namespace b200q
