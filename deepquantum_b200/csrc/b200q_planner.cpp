// Fusion planner (host).  See b200q_program.h for the pass / round / op model.
//
// The reference walks its nn.Sequential one gate at a time (circuit.py:261) and each gate moves the
// whole state through memory at least twice (qmath.py:503-504).  Here consecutive gates are grouped
// so that the state is read and written once per *pass*:
//   1. pick the tile bits of the next pass greedily: the set of index bits that lets the longest
//      run of pending gates execute (only NON-diagonal targets have to be tile bits -- controls and
//      diagonal gates act on any bit, they become per-thread / per-CTA predicates and phases);
//   2. inside the tile, pick register slots round by round the same way;
//   3. gates are reordered only across gates they commute with (two gates commute if on every
//      shared qubit both act diagonally -- as control or diagonal target).
#include "b200q_planner.h"

#include <algorithm>
#include <cstring>

namespace b200q {
namespace {

struct IGate {
  int kind = 0, k = 0;
  int t[B200Q_MAX_TARGETS] = {0};
  uint64_t ctrl = 0;
  uint64_t tmask = 0;  // bits acted on non-diagonally
  uint64_t dmask = 0;  // bits acted on diagonally (controls, diagonal targets)
  int64_t mat = 0;
  int flags = 0;
  bool reg_kind = false;  // executable inside a register round
  bool big = false;       // dense on 5..6 targets: a pass of its own (b200q_program.h, "dense pass")
  int op_kind = 0;
  int pool = 0;
};

inline int popc(uint64_t x) { return __builtin_popcountll(x); }
constexpr int kScanWindow = 2048;  // pending gates inspected per scheduling decision

struct Builder {
  int n_qubits, n_bits, vs, rb, rc, na, t_full, dtype;
  PlanOptions opt;
  std::vector<IGate> g;
  std::vector<char> done;
  int first_undone = 0;
  mutable int last_xc1 = 0;   // X gates of the last scan() whose control is a register slot
  uint64_t all_bits;

  // Gates executable (in order) when non-diagonal targets must lie in `allowed`.
  //   mode 0: tile choice  (every kind counts)
  //   mode 1: register round (register kinds execute; MATK gates block)
  //   mode 2: direct round (only MATK gates execute; register kinds block)
  // strict_x: an X gate whose control is a register slot needs a physical amplitude swap (~100 MOVs) instead of
  // the free relabelling; in strict mode such gates wait for a round where the control is thread-level.
  // avoid_rr: a diagonal gate with two (or more) selectors on register slots has no in-place code path (it runs
  // through the generic diagonal of the non-lean kernel); diagonal gates can run in ANY round, so it waits for a
  // round where at most one of its qubits is a register slot.
  int scan(uint64_t allowed, int mode, int limit, int pool_left, std::vector<int>* out, bool strict_x = false,
           bool avoid_rr = false) const {
    uint64_t bfull = 0, bdiag = 0;
    int cnt = 0, visited = 0;
    last_xc1 = 0;
    for (int i = first_undone; i < (int)g.size(); ++i) {
      if (done[i]) continue;
      if ((bfull & all_bits) == all_bits) break;
      if (++visited > kScanWindow) break;
      const IGate& a = g[i];
      bool ok = !a.big && !(a.tmask & (bfull | bdiag)) && !(a.dmask & bfull) && !(a.tmask & ~allowed);
      if (ok && mode == 1 && !a.reg_kind) ok = false;
      if (ok && strict_x && a.op_kind == B200Q_OP_X && (a.ctrl & allowed)) ok = false;
      if (ok && mode == 2 && a.reg_kind) ok = false;
      if (ok && avoid_rr && mode == 1 && a.op_kind == B200Q_OP_DIAG && popc((a.dmask & ~a.ctrl) & allowed) >= 2) ok = false;
      if (ok && a.pool > pool_left) ok = false;
      if (ok) {
        ++cnt;
        if (mode == 1 && a.op_kind == B200Q_OP_X && (a.ctrl & allowed)) ++last_xc1;   // needs a register swap
        pool_left -= a.pool;
        if (out) out->push_back(i);
        if (cnt >= limit) break;
      } else {
        bfull |= a.tmask;
        bdiag |= a.dmask;
      }
    }
    return cnt;
  }

  int first_missing_bit(uint64_t s) const {
    for (int i = first_undone; i < (int)g.size(); ++i) {
      if (done[i]) continue;
      const uint64_t miss = g[i].tmask & ~s;
      if (miss) return __builtin_ctzll(miss);
    }
    return -1;
  }
};

}  // namespace

Plan* make_plan(int n_qubits, int dtype, const b200q_gate_t* gates, int n_gates, const PlanOptions& opt_in,
                std::string* err) {
  auto fail = [&](const std::string& m) -> Plan* {
    if (err) *err = m;
    return nullptr;
  };
  if (n_qubits < 1 || n_qubits > B200Q_MAX_QUBITS - 2) return fail("n_qubits out of range");
  if (dtype != B200Q_C64 && dtype != B200Q_C128) return fail("dtype must be B200Q_C64 or B200Q_C128");
  if (n_gates < 0 || (n_gates > 0 && !gates)) return fail("bad gate list");

  Builder B;
  B.opt = opt_in;
  if (B.opt.chunk_bits < 8 || B.opt.chunk_bits > 13) return fail("chunk_bits must be in 8..13");
  B.dtype = dtype;
  B.vs = dtype == B200Q_C64 ? 1 : 0;
  B.rc = B200Q_REG_CHUNK_BITS;
  B.rb = B.rc + B.vs;
  B.na = 1 << B.rb;
  B.t_full = B.opt.chunk_bits + B.vs;
  B.n_qubits = n_qubits;
  B.n_bits = std::max(n_qubits, B.rb);
  B.all_bits = (1ull << n_qubits) - 1ull;
  if (B.opt.low_bits <= 0) B.opt.low_bits = 5 + B.vs;  // 512-byte contiguous runs
  if (B.opt.max_rounds < 3 || B.opt.max_rounds > B200Q_MAX_ROUNDS) B.opt.max_rounds = B200Q_MAX_ROUNDS;
  if (B.opt.max_ops <= 0 || B.opt.max_ops > B200Q_MAX_OPS) B.opt.max_ops = B200Q_MAX_OPS;
  const int t_eff = std::min(B.t_full, B.n_bits);

  // ---- validate and classify ------------------------------------------------------------------
  B.g.resize(n_gates);
  for (int i = 0; i < n_gates; ++i) {
    const b200q_gate_t& s = gates[i];
    IGate& a = B.g[i];
    a.kind = s.kind;
    a.k = s.n_targets;
    a.ctrl = s.controls;
    a.mat = s.mat_offset;
    a.flags = s.flags;
    if (a.k < 1 || a.k > B200Q_MAX_TARGETS) return fail("gate " + std::to_string(i) + ": bad n_targets");
    if (s.mat_offset < 0 || s.mat_offset > 0xffffffffll) return fail("gate matrix offset out of range");
    uint64_t tm = 0;
    for (int j = 0; j < a.k; ++j) {
      a.t[j] = s.targets[j];
      if (a.t[j] < 0 || a.t[j] >= n_qubits) return fail("gate " + std::to_string(i) + ": target out of range");
      if (tm >> a.t[j] & 1) return fail("gate " + std::to_string(i) + ": repeated target");
      tm |= 1ull << a.t[j];
    }
    if (a.ctrl & tm) return fail("gate " + std::to_string(i) + ": control equals target");
    if (a.ctrl & ~B.all_bits) return fail("gate " + std::to_string(i) + ": control out of range");
    switch (a.kind) {
      case B200Q_GATE_X:
        if (a.k != 1) return fail("X gate needs exactly one target");
        a.reg_kind = true; a.op_kind = B200Q_OP_X; a.tmask = tm; a.dmask = a.ctrl; a.pool = 0;
        break;
      case B200Q_GATE_MAT:
        if (a.k == 1) { a.reg_kind = true; a.op_kind = B200Q_OP_MAT1; a.pool = 4; }
        else if (a.k >= 3 && a.k <= B200Q_DENSE_MAX && (a.flags & B200Q_GATE_GRAD)) {
          a.reg_kind = false; a.op_kind = B200Q_OP_MATK; a.pool = 0; a.big = true;   // cotangent wanted: own pass
        }
        else if (a.k <= B200Q_MATK_MAX) { a.reg_kind = false; a.op_kind = B200Q_OP_MATK; a.pool = 1 << (2 * a.k); }
        else if (a.k <= B200Q_DENSE_MAX) { a.reg_kind = false; a.op_kind = B200Q_OP_MATK; a.pool = 0; a.big = true; }
        else return fail("dense gates on more than 6 targets are not supported");
        a.tmask = tm; a.dmask = a.ctrl;
        break;
      case B200Q_GATE_DIAG:
        if (a.k <= 2) { a.reg_kind = true; a.op_kind = B200Q_OP_DIAG; a.pool = 4; a.tmask = 0; a.dmask = tm | a.ctrl; }
        else if (a.k <= B200Q_DENSE_MAX && (a.flags & B200Q_GATE_GRAD)) {
          a.reg_kind = false; a.op_kind = B200Q_OP_MATK; a.pool = 0; a.big = true; a.tmask = tm; a.dmask = a.ctrl;
        }
        else if (a.k <= B200Q_MATK_MAX) {
          a.reg_kind = false; a.op_kind = B200Q_OP_MATK; a.pool = 1 << (2 * a.k); a.tmask = tm; a.dmask = a.ctrl;
        } else if (a.k <= B200Q_DENSE_MAX) {
          a.reg_kind = false; a.op_kind = B200Q_OP_MATK; a.pool = 0; a.big = true; a.tmask = tm; a.dmask = a.ctrl;
        } else return fail("diagonal gates on more than 6 targets are not supported");
        break;
      default:
        return fail("unknown gate kind");
    }
  }
  B.done.assign(n_gates, 0);

  Plan* plan = new Plan();
  plan->n_qubits = n_qubits;
  plan->n_bits = B.n_bits;
  plan->dtype = dtype;
  plan->opt = B.opt;
  plan->stats.n_gates = n_gates;

  // coalescing of the global rounds: lanes own the lowest `coalesce_bits` chunk bits (3: 128-byte runs)
  const int min_chunk_loc = std::max(0, std::min(B.opt.coalesce_bits, (t_eff - B.vs) - B.rc));
  const int min_loc = min_chunk_loc + B.vs;

  while (true) {
    while (B.first_undone < n_gates && B.done[B.first_undone]) ++B.first_undone;
    if (B.first_undone >= n_gates) break;
    const IGate& g0 = B.g[B.first_undone];
    const int op_cap = B.opt.fuse ? B.opt.max_ops : 1;
    if (g0.big) {   // dense pass: the gate alone, every earlier gate is done (scan() never lets a later gate overtake it
                    // on a shared qubit, and never puts it into a tile pass)
      b200q_pass_t P;
      std::memset(&P, 0, sizeof(P));
      P.n_bits = (uint8_t)B.n_bits;
      P.n_qubits = (uint8_t)n_qubits;
      P.tile_bits = (uint8_t)g0.k;
      P.n_rounds = 0;
      P.n_ops = 1;
      b200q_op_t& op = P.ops[0];
      op.kind = B200Q_OP_MATK;
      op.k = (uint8_t)g0.k;
      op.flags = (uint8_t)((g0.flags & B200Q_GATE_ADJOINT) ? B200Q_FLAG_ADJOINT : 0);
      op.mat_src = (uint32_t)g0.mat;
      op.gate_id = (uint32_t)B.first_undone;
      op.ctrl_glob = g0.ctrl;
      op.dsel_slot[0] = op.dsel_slot[1] = 0xff;
      op.code = B200Q_CODE_NONE;
      for (int j = 0; j < g0.k; ++j) op.dsel_glob[0] |= uint64_t(g0.t[j] & 0xff) << (8 * j);
      B.done[B.first_undone] = 1;
      plan->passes.push_back(P);
      plan->pass_gate_count.push_back(1);
      plan->stats.n_ops += 1;
      plan->stats.n_direct += 1;
      continue;
    }

    // ---- 1. tile bits ---------------------------------------------------------------------------
    uint64_t S = g0.tmask;
    if (B.vs) S |= 1ull;
    for (int b = 0; b < B.opt.low_bits && b < B.n_bits && popc(S) < t_eff; ++b) S |= 1ull << b;
    int cur = B.scan(S, 0, op_cap, B200Q_POOL_MAX, nullptr);
    while (popc(S) < t_eff) {
      int best = -1, bestc = cur;
      if (cur < op_cap) {
        for (int b = 0; b < n_qubits; ++b) {
          if (S >> b & 1) continue;
          const int c = B.scan(S | (1ull << b), 0, op_cap, B200Q_POOL_MAX, nullptr);
          if (c > bestc) { bestc = c; best = b; }
        }
      }
      if (best < 0) {
        best = cur < op_cap ? B.first_missing_bit(S) : -1;
        if (best < 0)
          for (int b = 0; b < B.n_bits; ++b)
            if (!(S >> b & 1)) { best = b; break; }
        bestc = B.scan(S | (1ull << best), 0, op_cap, B200Q_POOL_MAX, nullptr);
      }
      S |= 1ull << best;
      cur = bestc;
    }

    b200q_pass_t P;
    std::memset(&P, 0, sizeof(P));
    P.n_bits = (uint8_t)B.n_bits;
    P.n_qubits = (uint8_t)n_qubits;
    P.tile_bits = (uint8_t)t_eff;
    int loc_of[64];
    for (int b = 0, j = 0, q = 0; b < B.n_bits; ++b) {
      loc_of[b] = -1;
      if (S >> b & 1) { loc_of[b] = j; P.tile_phys[j++] = (uint8_t)b; }
      else P.nontile_phys[q++] = (uint8_t)b;
    }
    P.n_nontile = (uint8_t)(B.n_bits - t_eff);

    int n_ops = 0, pool = 0, n_rounds = 0, gates_in_pass = 0;

    // choose the register slots of one round; returns physical-bit mask R and the executable gates
    auto choose_round_impl = [&](bool restricted, bool strict, bool avoid_rr, uint64_t* R_out, std::vector<int>* list) {
      const int limit = op_cap - gates_in_pass;
      const int pool_left = B200Q_POOL_MAX - pool;
      uint64_t R = B.vs ? 1ull : 0ull;
      uint64_t cand = 0;
      for (int j = B.vs; j < t_eff; ++j)
        if (!restricted || j >= min_loc) cand |= 1ull << P.tile_phys[j];
      // score of a slot set: executable gates, minus a penalty per CNOT whose control is a slot too (a register
      // swap, ~3x the cost of the relabelling it is when the control stays a thread-level bit)
      const int W = 4, PEN = B.opt.xc1_penalty;
      auto score = [&](uint64_t Rs) {
        const int c = B.scan(Rs, 1, limit, pool_left, nullptr, strict, avoid_rr);
        return W * c - PEN * B.last_xc1;
      };
      int cnt = limit > 0 ? score(R) : 0;
      for (int s = 0; s < B.rc && limit > 0; ++s) {
        int best = -1, bestc = cnt;
        for (int j = B.vs; j < t_eff; ++j) {
          const uint64_t bit = 1ull << P.tile_phys[j];
          if (!(cand & bit) || (R & bit)) continue;
          const int c = score(R | bit);
          if (c > bestc) { bestc = c; best = j; }
        }
        if (best < 0) break;
        R |= 1ull << P.tile_phys[best];
        cnt = bestc;
      }
      // fill the remaining slots with free candidate bits, highest first, preferring bits that do not turn a
      // relabelling CNOT into a register swap
      for (int pass2 = 0; pass2 < 2; ++pass2)
        for (int j = t_eff - 1; j >= B.vs && popc(R) < B.rb; --j) {
          const uint64_t bit = 1ull << P.tile_phys[j];
          if (!(cand & bit) || (R & bit)) continue;
          if (pass2 == 0 && PEN > 0 && limit > 0 && score(R | bit) < cnt) continue;
          R |= bit;
        }
      for (int j = t_eff - 1; j >= B.vs && popc(R) < B.rb; --j) {  // tiny tiles: take anything
        const uint64_t bit = 1ull << P.tile_phys[j];
        if (!(R & bit)) R |= bit;
      }
      list->clear();
      if (limit > 0) B.scan(R, 1, limit, pool_left, list, strict, avoid_rr);
      *R_out = R;
    };
    auto choose_round = [&](bool restricted, uint64_t* R_out, std::vector<int>* list) {
      // strict X placement fragments rounds (153 -> 285): off
      choose_round_impl(restricted, false, B.opt.structured != 0, R_out, list);
      if (list->empty() && B.opt.structured) choose_round_impl(restricted, false, false, R_out, list);
    };

    auto is_restricted_ok = [&](uint64_t R) {
      for (int j = B.vs; j < min_loc; ++j)
        if (R >> P.tile_phys[j] & 1) return false;
      return true;
    };

    auto emit_round = [&](uint64_t R, const std::vector<int>& list, bool src_global) -> b200q_round_t* {
      b200q_round_t& Rd = P.rounds[n_rounds++];
      std::memset(&Rd, 0, sizeof(Rd));
      Rd.src_global = src_global;
      Rd.dst_global = 0;
      int slot_of[64];
      for (int b = 0; b < 64; ++b) slot_of[b] = -1;
      int ns = 0;
      for (int j = 0; j < t_eff; ++j)
        if (R >> P.tile_phys[j] & 1) { slot_of[P.tile_phys[j]] = ns; Rd.slot_bit[ns++] = (uint8_t)j; }
      // item-index bits: ascending; for tile<->tile rounds move three bits with distinct chunk-bit
      // residues mod 3 to the front so that quarter-warps are bank-conflict free under swz()
      std::vector<int> nr;
      for (int j = 0; j < t_eff; ++j)
        if (!(R >> P.tile_phys[j] & 1)) nr.push_back(j);
      Rd.op_begin = (uint16_t)n_ops;
      // A 1-selector diagonal whose qubit is a register slot of THIS round costs a sweep over half the registers;
      // outside the registers it is one scalar multiply of the thread's phase.  Diagonal gates can run in any
      // round: leave such a gate pending (a later round / pass usually finds its qubit outside the registers)
      // unless a later gate of this round needs it done (acts non-diagonally on its qubit).
      std::vector<char> skip(list.size(), 0);
      if (B.opt.structured && B.opt.defer_diag) {
        int kept = 0;
        for (size_t li = 0; li < list.size(); ++li) {
          const IGate& a = B.g[list[li]];
          bool defer = false;
          const bool free_phase = B.opt.free_phase && a.op_kind == B200Q_OP_DIAG && a.ctrl == 0 && a.k == 1 &&
                                  (a.flags & B200Q_GATE_PHASE_MASK) != 0;
          if (free_phase) {
            // free when its qubit is a register slot; otherwise wait (no gate of this round can need it done: a gate
            // acting non-diagonally on the qubit would have it as a register slot)
            defer = slot_of[a.t[0]] < 0;
          } else if (a.op_kind == B200Q_OP_DIAG && a.ctrl == 0) {
            int nreg = 0;
            for (int j = 0; j < a.k; ++j) nreg += slot_of[a.t[j]] >= 0;
            if (nreg >= 1) {
              defer = true;
              for (size_t lj = li + 1; lj < list.size() && defer; ++lj)
                if (B.g[list[lj]].tmask & a.dmask) defer = false;
            }
          }
          skip[li] = defer;
          kept += !defer;
        }
        if (kept == 0) std::fill(skip.begin(), skip.end(), 0);
      }
      size_t list_pos = 0;
      for (int gi : list) {
        if (skip[list_pos++]) continue;
        const IGate& a = B.g[gi];
        b200q_op_t op;
        std::memset(&op, 0, sizeof(op));
        op.kind = (uint8_t)a.op_kind;
        op.flags = (uint8_t)(((a.flags & B200Q_GATE_ADJOINT) ? B200Q_FLAG_ADJOINT : 0) |
                             ((a.flags & B200Q_GATE_REAL) ? B200Q_FLAG_REAL : 0) |
                             ((a.flags & B200Q_GATE_RXLIKE) ? B200Q_FLAG_RXLIKE : 0) |
                             ((a.flags & B200Q_GATE_HADAMARD) ? B200Q_FLAG_HAD : 0) |
                             ((a.flags & B200Q_GATE_ROTATION) ? B200Q_FLAG_ROT : 0) |
                             ((a.op_kind == B200Q_OP_DIAG && a.k == 1) ? (a.flags & B200Q_GATE_PHASE_MASK) : 0));
        op.mat_src = (uint32_t)a.mat;
        op.gate_id = (uint32_t)gi;
        op.k = (uint8_t)a.k;
        op.dsel_slot[0] = op.dsel_slot[1] = 0xff;
        for (uint64_t c = a.ctrl; c; c &= c - 1) {
          const int b = __builtin_ctzll(c);
          if (slot_of[b] >= 0) op.ctrl_reg |= 1u << slot_of[b];
          else if (loc_of[b] >= 0) op.ctrl_loc |= 1u << loc_of[b];
          else op.ctrl_glob |= 1ull << b;
        }
        if (a.op_kind == B200Q_OP_MAT1 || a.op_kind == B200Q_OP_X) {
          op.slot = (uint8_t)slot_of[a.t[0]];
        } else if (a.op_kind == B200Q_OP_DIAG) {
          for (int j = 0; j < a.k; ++j) {
            const int b = a.t[j];
            if (slot_of[b] >= 0) op.dsel_slot[j] = (uint8_t)slot_of[b];
            else if (loc_of[b] >= 0) op.dsel_loc[j] = 1u << loc_of[b];
            else op.dsel_glob[j] = 1ull << b;
          }
        }
        // complex64: an op that touches the lane slot (index bit 0: the two amplitudes of a 16-byte chunk) is
        // wrapped in two LSWAP ops that exchange the lane with a chunk-level slot the op does not use, so that
        // it runs through the ordinary chunk-slot code (the scalar lane paths were 3x more expensive).
        int lswap = -1;
        if (B.vs && B.opt.structured) {
          uint32_t used = op.ctrl_reg;   // amplitude-level slot mask
          if (a.op_kind == B200Q_OP_MAT1 || a.op_kind == B200Q_OP_X) used |= 1u << op.slot;
          for (int j = 0; j < 2; ++j)
            if (op.dsel_slot[j] != 0xff) used |= 1u << op.dsel_slot[j];
          // X on the lane with chunk-level controls only is a physical lane swap (X_LANE), not wrapped: an X
          // relabelling between two LSWAPs would leave the lane flipped
          const bool x_on_lane = a.op_kind == B200Q_OP_X && op.slot == 0 && !(op.ctrl_reg & 1u);
          if ((used & 1u) && !x_on_lane) {
            for (int cs = B.rc - 1; cs >= 0 && lswap < 0; --cs)
              if (!((used >> (cs + 1)) & 1u)) lswap = cs;
          }
          if (lswap >= 0) {
            const int as = lswap + 1;   // amplitude-level slot index of the partner chunk slot
            if (op.ctrl_reg & 1u) op.ctrl_reg = (op.ctrl_reg & ~1u) | (1u << as);
            if ((a.op_kind == B200Q_OP_MAT1 || a.op_kind == B200Q_OP_X) && op.slot == 0) op.slot = (uint8_t)as;
            for (int j = 0; j < 2; ++j)
              if (op.dsel_slot[j] == 0) op.dsel_slot[j] = (uint8_t)as;
          }
        }
        const int need = lswap >= 0 ? 3 : 1;
        if (n_ops + need > B200Q_MAX_OPS) break;   // the rest of the list stays pending
        auto push_lswap = [&]() {
          b200q_op_t& w = P.ops[n_ops++];
          std::memset(&w, 0, sizeof(w));
          w.kind = B200Q_OP_LSWAP;
          w.slot = (uint8_t)(lswap + 1);
          w.code = (uint8_t)(B200Q_CODE_LSWAP + lswap);
          w.dsel_slot[0] = w.dsel_slot[1] = 0xff;
          w.gate_id = (uint32_t)gi;
        };
        if (lswap >= 0) push_lswap();
        if (a.pool) {
          op.pool_off = (uint16_t)pool;
          op.pool_n = (uint16_t)a.pool;
          pool += a.pool;
        }
        // pre-decoded dispatch code
        op.tctrl = (op.ctrl_loc != 0 || op.ctrl_glob != 0) ? 1 : 0;
        const bool lane_slot = B.vs && op.slot == 0;
        const bool lane_ctrl = B.vs && (op.ctrl_reg & 1u);
        const int nsel = (op.dsel_slot[0] != 0xff) + (op.dsel_slot[1] != 0xff);
        if (a.op_kind == B200Q_OP_MAT1) {
          if (op.ctrl_reg == 0 && !lane_slot) {
            const int var = (op.flags & B200Q_FLAG_REAL) ? 0 : ((op.flags & B200Q_FLAG_RXLIKE) ? 1 : 2);
            op.code = (uint8_t)(B200Q_CODE_MAT1_FAST + 4 * var + (op.slot - B.vs));
            if ((op.flags & B200Q_FLAG_HAD) && !op.tctrl && B.opt.structured) {
              op.code = (uint8_t)(B200Q_CODE_MAT1_HAD + (op.slot - B.vs));
              P.has_scale = 1;
            } else if ((op.flags & B200Q_FLAG_ROT) && var < 2 && B.opt.structured) {
              op.code = (uint8_t)((var == 1 ? B200Q_CODE_MAT1_ROTX : B200Q_CODE_MAT1_ROTY) + (op.slot - B.vs));
              if (!op.tctrl) P.has_scale = 1;
            }
          } else op.code = B200Q_CODE_MAT1_SLOW;
        } else if (a.op_kind == B200Q_OP_X) {
          const uint32_t cm = op.ctrl_reg >> B.vs;
          if (op.ctrl_reg == 0 && !lane_slot) { op.code = B200Q_CODE_X_RELABEL; op.arg = (uint8_t)(op.slot - B.vs); }
          else if (lane_slot && !lane_ctrl && B.opt.structured) { op.code = B200Q_CODE_X_LANE; op.arg = (uint8_t)cm; }
          else if (!lane_slot && !lane_ctrl && cm != 0 && (cm & (cm - 1)) == 0 && B.opt.structured) {
            op.code = B200Q_CODE_X_C1;
            op.arg = (uint8_t)(4 * (op.slot - B.vs) + __builtin_ctz(cm));
          } else op.code = B200Q_CODE_X_SLOW;
        } else {
          op.code = B200Q_CODE_DIAG;
          const bool lane_sel = B.vs && (op.dsel_slot[0] == 0 || op.dsel_slot[1] == 0);
          if (B.opt.structured && op.ctrl_reg == 0 && !lane_sel) {
            if (nsel == 0) op.code = B200Q_CODE_DIAG_T;
            else if (nsel == 1) {
              const int j = op.dsel_slot[0] != 0xff ? 0 : 1;
              op.code = (uint8_t)(B200Q_CODE_DIAG_R + (op.dsel_slot[j] - B.vs));
              op.arg = (uint8_t)j;
            }
          }
        }
        P.ops[n_ops++] = op;
        if (op.ctrl_glob) P.gctrl_ops[P.n_gctrl++] = (uint8_t)(n_ops - 1);
        if (lswap >= 0) push_lswap();
        B.done[gi] = 1;
        ++gates_in_pass;
      }
      Rd.op_end = (uint16_t)n_ops;
      for (size_t k = 0; k < nr.size() && k < sizeof(Rd.nonreg_bit); ++k) Rd.nonreg_bit[k] = (uint8_t)nr[k];
      return &Rd;
    };

    auto reorder_for_banks = [&](b200q_round_t* Rd) {
      const int nn = t_eff - B.rb;
      if (nn < 3) return;
      std::vector<int> nr(Rd->nonreg_bit, Rd->nonreg_bit + nn), front, rest;
      bool used[3] = {false, false, false};
      for (int j : nr) {
        const int res = (j - B.vs) % 3;
        if (front.size() < 3 && !used[res]) { used[res] = true; front.push_back(j); }
        else rest.push_back(j);
      }
      front.insert(front.end(), rest.begin(), rest.end());
      for (int k = 0; k < nn && k < (int)sizeof(Rd->nonreg_bit); ++k) Rd->nonreg_bit[k] = (uint8_t)front[k];
    };

    std::vector<int> list, list2;
    uint64_t R = 0, R2 = 0;
    bool first = true;
    b200q_round_t* last = nullptr;
    bool last_restricted_ok = false;
    while (true) {
      if (n_rounds >= B.opt.max_rounds - 1) break;
      if (gates_in_pass >= op_cap) break;
      if (n_ops + 3 > B200Q_MAX_OPS) break;
      if (first) {
        choose_round(true, &R, &list);
        if (list.empty()) {
          choose_round(false, &R2, &list2);
          std::vector<int> mk;
          const bool any_mid = !list2.empty() || B.scan(S, 2, 1, B200Q_POOL_MAX - pool, &mk) > 0;
          if (!any_mid) break;
          last = emit_round(R, list, true);  // plain load round
          last_restricted_ok = true;
          first = false;
          continue;
        }
        last = emit_round(R, list, true);
        last_restricted_ok = true;
        first = false;
        continue;
      }
      choose_round(false, &R, &list);
      if (!list.empty() && B.opt.structured && B.opt.defer_diag >= 2) {
        // a round of diagonal gates only is not worth a transpose of the tile: they run (for free) in the
        // first round of the next pass, or in this pass's write-back round if one has to be added anyway
        bool only_diag = true;
        for (int gi : list) only_diag = only_diag && B.g[gi].op_kind == B200Q_OP_DIAG;
        if (only_diag) break;
      }
      // a late round with very few gates costs a whole tile transpose: end the pass instead, the next pass picks
      // tile bits and slots afresh
      if (!list.empty() && n_rounds >= 2 && (int)list.size() < B.opt.min_round_gates &&
          gates_in_pass >= 4 * B.opt.min_round_gates)
        break;
      if (!list.empty()) {
        last = emit_round(R, list, false);
        last_restricted_ok = is_restricted_ok(R);
        continue;
      }
      // a dense multi-target gate, applied in place in shared memory
      std::vector<int> mk;
      if (B.scan(S, 2, 1, B200Q_POOL_MAX - pool, &mk) == 0 || n_ops >= B200Q_MAX_OPS) break;
      {
        const IGate& a = B.g[mk[0]];
        b200q_round_t& Rd = P.rounds[n_rounds++];
        std::memset(&Rd, 0, sizeof(Rd));
        Rd.direct = 1;
        Rd.op_begin = (uint16_t)n_ops;
        b200q_op_t& op = P.ops[n_ops++];
        std::memset(&op, 0, sizeof(op));
        op.kind = B200Q_OP_MATK;
        op.k = (uint8_t)a.k;
        op.flags = (uint8_t)((a.flags & B200Q_GATE_ADJOINT) ? B200Q_FLAG_ADJOINT : 0);
        op.mat_src = (uint32_t)a.mat;
        op.gate_id = (uint32_t)mk[0];
        op.dsel_slot[0] = op.dsel_slot[1] = 0xff;
        for (int j = 0; j < a.k; ++j) op.tk[j] = (uint8_t)loc_of[a.t[j]];
        for (uint64_t c = a.ctrl; c; c &= c - 1) {
          const int b = __builtin_ctzll(c);
          if (loc_of[b] >= 0) op.ctrl_loc |= 1u << loc_of[b];
          else op.ctrl_glob |= 1ull << b;
        }
        op.pool_off = (uint16_t)pool;
        op.pool_n = (uint16_t)a.pool;
        pool += a.pool;
        op.code = B200Q_CODE_NONE;
        op.tctrl = 1;
        Rd.op_end = (uint16_t)n_ops;
        B.done[mk[0]] = 1;
        ++gates_in_pass;
        ++plan->stats.n_direct;
        last = &Rd;
        last_restricted_ok = false;
      }
    }
    if (n_rounds == 0 || gates_in_pass == 0) {
      delete plan;
      return fail("planner made no progress (internal error)");
    }
    if (last && !last->direct && last_restricted_ok) {
      last->dst_global = 1;
    } else {
      choose_round(true, &R, &list);
      last = emit_round(R, list, false);
      last->dst_global = 1;
    }
    for (int r = 0; r < n_rounds; ++r)
      if (!P.rounds[r].direct && !P.rounds[r].src_global && !P.rounds[r].dst_global) reorder_for_banks(&P.rounds[r]);
    P.n_rounds = (uint8_t)n_rounds;
    P.n_ops = (uint8_t)n_ops;
    P.lean = 1;
    for (int o = 0; o < n_ops; ++o) {
      if (P.ops[o].code >= B200Q_CODE_LEAN_END) P.lean = 0;
      if (P.ops[o].kind == B200Q_OP_MATK) P.needs_pool = 1;
    }
    if (!P.lean) P.needs_pool = 1;
    P.pool_elems = (uint16_t)pool;
    plan->passes.push_back(P);
    plan->pass_gate_count.push_back(gates_in_pass);
    plan->stats.n_rounds += n_rounds;
    plan->stats.n_ops += n_ops;
  }
  plan->stats.n_passes = (int)plan->passes.size();
  // complex64: the state BETWEEN passes is kept in the SoA chunk format (see b200q_program.h)
  if (dtype == B200Q_C64) {
    const int np = (int)plan->passes.size();
    // (a dense pass reads and writes the caller layout: its neighbours use AoS on that side)
    auto tile_pass = [&](int i) { return i >= 0 && i < np && plan->passes[i].n_rounds > 0; };
    for (int i = 0; i < np; ++i) {
      if (!tile_pass(i)) { plan->passes[i].layout = 0; continue; }
      plan->passes[i].layout = (uint8_t)((tile_pass(i - 1) ? B200Q_LAYOUT_SRC_SOA : 0) | (tile_pass(i + 1) ? B200Q_LAYOUT_DST_SOA : 0));
    }
  }
  return plan;
}

}  // namespace b200q
