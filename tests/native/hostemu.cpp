// TEST-ONLY: steps the tile-kernel body (csrc/b200q_tile_body.h) thread by thread on the CPU so the
// planner output and the kernel's index logic can be checked against the oracle without a GPU.
// Never linked into libb200q.so; built by tests/native/build.py into tests/native/_build/.
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../deepquantum_b200/csrc/b200q_planner.h"
#include "../../deepquantum_b200/csrc/b200q_tile_body.h"

using namespace b200q;

namespace {
template <typename Real>
void run_pass(const Plan& pl, const b200q_pass_t& P, void* state_v, const void* mats_v, int64_t batch, int64_t mbs,
              const b200q_remote_t* remote = nullptr) {
  using chunk = typename Traits<Real>::chunk;
  constexpr int VS = Traits<Real>::VS;
  const int cb = pl.opt.chunk_bits;
  const int nthreads = 1 << (cb - B200Q_REG_CHUNK_BITS);
  if (P.n_rounds == 0) {   // dense pass (5..6 targets): plain loops, the semantics of b200q_dense_kernel
    const b200q_op_t& op = P.ops[0];
    const int K = op.k, D = 1 << K;
    const bool adj = (op.flags & B200Q_FLAG_ADJOINT) != 0;
    const uint64_t n_amps = 1ull << pl.n_qubits;
    for (int64_t b = 0; b < batch; ++b) {
      cx<Real>* st = reinterpret_cast<cx<Real>*>(state_v) + uint64_t(b) * n_amps;
      const cx<Real>* m = reinterpret_cast<const cx<Real>*>(mats_v) + b * mbs + op.mat_src;
      uint64_t tm = 0, off[64];
      for (int j = 0; j < K; ++j) tm |= 1ull << ((op.dsel_glob[0] >> (8 * j)) & 0xff);
      for (int r = 0; r < D; ++r) {
        off[r] = 0;
        for (int j = 0; j < K; ++j)
          if ((r >> j) & 1) off[r] |= 1ull << ((op.dsel_glob[0] >> (8 * j)) & 0xff);
      }
      std::vector<cx<Real>> x(D);
      for (uint64_t base = 0; base < n_amps; ++base) {
        if ((base & tm) || (base & op.ctrl_glob) != op.ctrl_glob) continue;
        for (int r = 0; r < D; ++r) x[r] = st[base | off[r]];
        for (int r = 0; r < D; ++r) {
          Real yr = 0, yi = 0;
          for (int c = 0; c < D; ++c) {
            cx<Real> w = adj ? m[c * D + r] : m[r * D + c];
            if (adj) w.y = -w.y;
            yr += w.x * x[c].x - w.y * x[c].y;
            yi += w.x * x[c].y + w.y * x[c].x;
          }
          st[base | off[r]].x = yr; st[base | off[r]].y = yi;
        }
      }
    }
    return;
  }
  std::vector<chunk> tile(size_t(1) << cb);
  std::vector<cx<Real>> pool(B200Q_POOL_MAX);
  std::vector<Real> coef(size_t(B200Q_MAX_OPS) * B200Q_COEF_PER_OP);
  const uint64_t chunks_per_state = (1ull << pl.n_qubits) >> VS;
  const uint64_t ntiles = 1ull << (int(P.n_bits) - int(P.tile_bits));
  for (int64_t b = 0; b < batch; ++b) {
    chunk* gstate = reinterpret_cast<chunk*>(state_v) + uint64_t(b) * chunks_per_state;
    const cx<Real>* m = reinterpret_cast<const cx<Real>*>(mats_v) + b * mbs;
    for (int tid = 0; tid < nthreads; ++tid) fill_coefs<Real>(P, tid, nthreads, coef.data(), m);
    std::vector<OpWord> words(B200Q_MAX_OPS + 1);
    for (int tid = 0; tid < nthreads; ++tid) fill_opwords(P, tid, nthreads, words.data());
    std::vector<uint64_t> dest_tab(B200Q_DEST_TAB_ENTRIES);
    if (remote && remote->enabled)
      for (int tid = 0; tid < nthreads; ++tid) fill_dest_tab(*remote, tid, nthreads, dest_tab.data());
    const Real gscale = P.has_scale ? Real(pass_scale<Real>(P, m)) : Real(1);
    for (uint64_t t = 0; t < ntiles; ++t) {
      const uint64_t cta_base = tile_base(P, t);
      const uint64_t enabled = tile_enabled(P, cta_base);
      // poison the tile so that a read of an unwritten slot is caught
      std::memset(tile.data(), 0xff, tile.size() * sizeof(chunk));
      std::vector<RoundTab> tabs(B200Q_MAX_ROUNDS);
      for (int tid = 0; tid < nthreads; ++tid) fill_round_tabs<Real>(P, tid, nthreads, tabs.data());
      if (P.pool_elems)
        for (int tid = 0; tid < nthreads; ++tid) fill_pool<Real>(P, tid, nthreads, pool.data(), m, false);
      for (int r = 0; r < P.n_rounds; ++r) {
        const b200q_round_t& Rd = P.rounds[r];
        if (Rd.direct) {
          for (int o = Rd.op_begin; o < Rd.op_end; ++o)
            for (int tid = 0; tid < nthreads; ++tid)
              run_direct_op<Real>(P, P.ops[o], tid, nthreads, cta_base, tile.data(), pool.data());
        } else {
          for (int tid = 0; tid < nthreads; ++tid)
            if (P.lean)
              run_round<Real, true>(P, Rd, tabs[r], tid, cta_base, enabled, tile.data(), pool.data(), coef.data(),
                                    words.data(), gscale, gstate, chunks_per_state, remote, dest_tab.data());
            else
              run_round<Real, false>(P, Rd, tabs[r], tid, cta_base, enabled, tile.data(), pool.data(), coef.data(),
                                     words.data(), gscale, gstate, chunks_per_state, remote, dest_tab.data());
        }
      }
    }
  }
}
}  // namespace

namespace {
template <typename Real>
void run_pass_adjoint(const Plan& pl, const b200q_pass_t& P, void* psi_v, void* lam_v, const void* mats_v,
                      double* grad, const unsigned char* need) {
  using chunk = typename Traits<Real>::chunk;
  constexpr int VS = Traits<Real>::VS;
  const int cb = pl.opt.chunk_bits;
  const int nthreads = 1 << (cb - B200Q_REG_CHUNK_BITS);
  if (P.n_rounds == 0) {   // dense pass: the semantics of b200q_adjoint_run's dense branch, plain loops
    const b200q_op_t& op = P.ops[0];
    const int K = op.k, D = 1 << K;
    const bool adj = (op.flags & B200Q_FLAG_ADJOINT) != 0;
    const uint64_t n_amps = 1ull << pl.n_qubits;
    b200q_pass_t Q = P;
    Q.ops[0].flags ^= B200Q_FLAG_ADJOINT;
    run_pass<Real>(pl, Q, psi_v, mats_v, 1, 0);
    if (!need || need[op.gate_id]) {
      const cx<Real>* psi = reinterpret_cast<const cx<Real>*>(psi_v);
      const cx<Real>* lam = reinterpret_cast<const cx<Real>*>(lam_v);
      uint64_t tm = 0, off[64];
      for (int r = 0; r < D; ++r) {
        off[r] = 0;
        for (int j = 0; j < K; ++j)
          if ((r >> j) & 1) off[r] |= 1ull << ((op.dsel_glob[0] >> (8 * j)) & 0xff);
      }
      tm = off[D - 1];
      for (uint64_t base = 0; base < n_amps; ++base) {
        if ((base & tm) || (base & op.ctrl_glob) != op.ctrl_glob) continue;
        for (int r = 0; r < D; ++r)
          for (int c = 0; c < D; ++c) {
            const cx<Real> l = lam[base | off[r]], q = psi[base | off[c]];
            const double re = double(l.x) * q.x + double(l.y) * q.y, im = double(l.y) * q.x - double(l.x) * q.y;
            const int dst = adj ? c * D + r : r * D + c;
            grad[2 * (uint64_t(op.mat_src) + dst)] += re;
            grad[2 * (uint64_t(op.mat_src) + dst) + 1] += adj ? -im : im;
          }
      }
    }
    run_pass<Real>(pl, Q, lam_v, mats_v, 1, 0);
    return;
  }
  std::vector<chunk> tp(size_t(1) << cb), tl(size_t(1) << cb);
  std::vector<cx<Real>> pool(B200Q_POOL_MAX);
  std::vector<double> acc(size_t(B200Q_MAX_OPS) * B200Q_ACC_PER_OP), gfac(B200Q_MAX_OPS);
  std::vector<Real> coef(size_t(B200Q_MAX_OPS) * B200Q_COEF_PER_OP);
  std::vector<OpWord> words(B200Q_MAX_OPS + 1);
  const uint64_t chunks_per_state = (1ull << pl.n_qubits) >> VS;
  const uint64_t ntiles = 1ull << (int(P.n_bits) - int(P.tile_bits));
  uint64_t want = 0;
  for (int o = 0; o < P.n_ops; ++o) {
    const b200q_op_t& op = P.ops[o];
    if (op.kind == B200Q_OP_X) continue;
    if (need && !need[op.gate_id]) continue;
    if (op.kind == B200Q_OP_MATK && op.k > 2) continue;
    want |= 1ull << o;
  }
  chunk* gpsi = reinterpret_cast<chunk*>(psi_v);
  chunk* glam = reinterpret_cast<chunk*>(lam_v);
  const cx<Real>* m = reinterpret_cast<const cx<Real>*>(mats_v);
  for (int tid = 0; tid < nthreads; ++tid) fill_coefs<Real>(P, tid, nthreads, coef.data(), m, true);
  for (int tid = 0; tid < nthreads; ++tid) fill_opwords(P, tid, nthreads, words.data());
  const Real gscale = Real(adjoint_scales<Real>(P, m, want, gfac.data()));
  for (uint64_t t = 0; t < ntiles; ++t) {
    const uint64_t cta_base = tile_base(P, t);
    std::memset(tp.data(), 0xff, tp.size() * sizeof(chunk));
    std::memset(tl.data(), 0xff, tl.size() * sizeof(chunk));
    std::fill(acc.begin(), acc.end(), 0.0);
    const int nwarps = (nthreads + 31) / 32;
    std::vector<double> wacc(size_t(nwarps) * B200Q_MAX_OPS * B200Q_WACC_PER_OP, 0.0);
    std::vector<RoundTab> tabs(B200Q_MAX_ROUNDS);
    for (int tid = 0; tid < nthreads; ++tid) fill_round_tabs<Real>(P, tid, nthreads, tabs.data());
    if (P.pool_elems)
      for (int tid = 0; tid < nthreads; ++tid) fill_pool<Real>(P, tid, nthreads, pool.data(), m, true);
    for (int r = int(P.n_rounds) - 1; r >= 0; --r) {
      const b200q_round_t& Rd = P.rounds[r];
      if (Rd.direct) {
        for (int o = int(Rd.op_end) - 1; o >= int(Rd.op_begin); --o)
          for (int tid = 0; tid < nthreads; ++tid)
            run_direct_op_adjoint<Real>(P, P.ops[o], tid, nthreads, cta_base, tp.data(), tl.data(), pool.data(),
                                        (want >> o) & 1ull, acc.data() + o * B200Q_ACC_PER_OP);
      } else {
        for (int tid = 0; tid < nthreads; ++tid)
          run_round_adjoint<Real>(P, Rd, tabs[r], tid, cta_base, tp.data(), tl.data(), pool.data(), coef.data(),
                                  words.data(), gscale, gpsi, glam, chunks_per_state, want, wacc.data());
      }
    }
    for (int tid = 0; tid < nthreads; ++tid) merge_warp_acc(P, tid, nthreads, nwarps, wacc.data(), acc.data());
    for (int tid = 0; tid < nthreads; ++tid)
      flush_grad(P, tid, nthreads, want, acc.data(), gfac.data(), grad, [](double* p, double v) { *p += v; });
  }
}
}  // namespace

// psi: final state (in/out), lam: cotangent of the final state (in/out), grad: zeroed complex128 buffer
extern "C" int hostemu_adjoint(int n_qubits, int dtype, const b200q_gate_t* gates, int n_gates, int chunk_bits,
                               void* psi, void* lam, const void* mats, double* grad, const unsigned char* need,
                               char* err_out, int err_len) {
  PlanOptions opt;
  if (chunk_bits) opt.chunk_bits = chunk_bits;
  std::string err;
  Plan* pl = make_plan(n_qubits, dtype, gates, n_gates, opt, &err);
  if (!pl) {
    if (err_out && err_len > 0) { std::strncpy(err_out, err.c_str(), err_len - 1); err_out[err_len - 1] = 0; }
    return -1;
  }
  for (int i = (int)pl->passes.size() - 1; i >= 0; --i) {
    if (dtype == B200Q_C64) run_pass_adjoint<float>(*pl, pl->passes[i], psi, lam, mats, grad, need);
    else run_pass_adjoint<double>(*pl, pl->passes[i], psi, lam, mats, grad, need);
  }
  delete pl;
  return 0;
}

extern "C" int hostemu_run(int n_qubits, int dtype, const b200q_gate_t* gates, int n_gates, int chunk_bits,
                           int low_bits, int max_rounds, int fuse, void* state, const void* mats, int64_t batch,
                           int64_t mbs, int* stats_out, char* err_out, int err_len) {
  PlanOptions opt;
  if (chunk_bits) opt.chunk_bits = chunk_bits;
  if (low_bits) opt.low_bits = low_bits;
  if (max_rounds) opt.max_rounds = max_rounds;
  opt.fuse = fuse & 1;
  opt.structured = (fuse & 2) ? 0 : 1;   // test hook: bit 1 selects the general (un-structured) op codes
  std::string err;
  Plan* pl = make_plan(n_qubits, dtype, gates, n_gates, opt, &err);
  if (!pl) {
    if (err_out && err_len > 0) {
      std::strncpy(err_out, err.c_str(), err_len - 1);
      err_out[err_len - 1] = 0;
    }
    return -1;
  }
  for (const auto& P : pl->passes) {
    if (dtype == B200Q_C64) run_pass<float>(*pl, P, state, mats, batch, mbs);
    else run_pass<double>(*pl, P, state, mats, batch, mbs);
  }
  if (stats_out) {
    stats_out[0] = pl->stats.n_passes;
    stats_out[1] = pl->stats.n_rounds;
    stats_out[2] = pl->stats.n_ops;
    stats_out[3] = pl->stats.n_direct;
  }
  delete pl;
  return 0;
}

// Passes [first, last) of the plan only: lets tests compare the generic kernel body with the generated pass kernels
// pass by pass.
extern "C" int hostemu_run_range(int n_qubits, int dtype, const b200q_gate_t* gates, int n_gates, int chunk_bits,
                                 int fuse, int first, int last, void* state, const void* mats, int64_t batch) {
  PlanOptions opt;
  if (chunk_bits) opt.chunk_bits = chunk_bits;
  opt.fuse = fuse & 1;
  std::string err;
  Plan* pl = make_plan(n_qubits, dtype, gates, n_gates, opt, &err);
  if (!pl) return -1;
  for (int i = first; i < last && i < (int)pl->passes.size(); ++i) {
    if (dtype == B200Q_C64) run_pass<float>(*pl, pl->passes[i], state, mats, batch, 0);
    else run_pass<double>(*pl, pl->passes[i], state, mats, batch, 0);
  }
  delete pl;
  return 0;
}

// Fused pass + exchange for ONE rank with all ranks' buffers in one address space: the plan runs on `state`; its
// last pass scatters into buffers[0..W-1] exactly like the kernel does through the NVLink peer mappings.
// perm: bit permutation of the distributed index (NULL: block transpose), same convention as the C ABI.
extern "C" int hostemu_run_exchange_rank(int n_local, int dtype, const b200q_gate_t* gates, int n_gates,
                                         int chunk_bits, int coalesce_bits, void* state, void** buffers,
                                         const void* mats, int n_ranks, int rank, const uint8_t* perm, char* err_out,
                                         int err_len) {
  PlanOptions opt;
  if (chunk_bits) opt.chunk_bits = chunk_bits;
  if (coalesce_bits >= 0) opt.coalesce_bits = coalesce_bits;
  std::string err;
  Plan* pl = make_plan(n_local, dtype, gates, n_gates, opt, &err);
  if (!pl) {
    if (err_out && err_len > 0) { std::strncpy(err_out, err.c_str(), err_len - 1); err_out[err_len - 1] = 0; }
    return -1;
  }
  int g = 0;
  while ((1 << g) < n_ranks) ++g;
  const int vs = dtype == B200Q_C64 ? 1 : 0, nl = n_local, nt = nl + g;
  b200q_remote_t R;
  std::memset(&R, 0, sizeof R);
  for (int q = 0; q < n_ranks; ++q) R.peer[q] = buffers[q];
  uint8_t pm[48];
  for (int j = 0; j < nt; ++j) pm[j] = perm ? perm[j] : (uint8_t)j;
  if (!perm)
    for (int k = 0; k < g; ++k) { pm[nl - g + k] = (uint8_t)(nl + k); pm[nl + k] = (uint8_t)(nl - g + k); }
  R.n_chunk_bits = nl - vs;
  for (int j = vs; j < nl; ++j) R.perm[j - vs] = (uint8_t)(pm[j] - vs);
  for (int k = 0; k < g; ++k)
    if ((rank >> k) & 1) {
      const int pos = pm[nl + k] - vs;
      R.base |= pos < R.n_chunk_bits ? (1ull << pos) : (1ull << (B200Q_DEST_RANK_SHIFT + pos - R.n_chunk_bits));
    }
  R.enabled = 1;
  for (size_t i = 0; i < pl->passes.size(); ++i) {
    const b200q_remote_t* rp = i + 1 == pl->passes.size() ? &R : nullptr;
    if (dtype == B200Q_C64) run_pass<float>(*pl, pl->passes[i], state, mats, 1, 0, rp);
    else run_pass<double>(*pl, pl->passes[i], state, mats, 1, 0, rp);
  }
  delete pl;
  return 0;
}

// ---- qudit (Fock tensor) kernel: same geometry helpers, dense contraction -----------------------------
#include "../../deepquantum_b200/csrc/b200q_qudit_geom.h"
#include <complex>

template <typename Real>
static void emu_qudit(std::complex<Real>* st, const QuditGeom& g, const std::complex<Real>* m, int64_t batch) {
  const int D = g.D, G = g.G;
  std::vector<std::complex<Real>> xs(size_t(D) * G), ys(size_t(D) * G);
  const long long nblocks = (g.n_rest + G - 1) / G;
  for (int64_t b = 0; b < batch; ++b) {
    std::complex<Real>* s = st + b * g.state_size;
    for (long long blk = 0; blk < nblocks; ++blk) {
      const long long r0 = blk * G;
      for (int e = 0; e < G * D; ++e) {
        int gi, t;
        qudit_elem(g, e, &gi, &t);
        xs[size_t(t) * G + gi] = (r0 + gi < g.n_rest) ? s[qudit_offset(g, r0 + gi, t)] : std::complex<Real>(0);
      }
      for (int r = 0; r < D; ++r)
        for (int gi = 0; gi < G; ++gi) {
          std::complex<Real> acc(0);
          for (int c = 0; c < D; ++c) acc += m[r * D + c] * xs[size_t(c) * G + gi];
          ys[size_t(r) * G + gi] = acc;
        }
      for (int e = 0; e < G * D; ++e) {
        int gi, t;
        qudit_elem(g, e, &gi, &t);
        if (r0 + gi < g.n_rest) s[qudit_offset(g, r0 + gi, t)] = ys[size_t(t) * G + gi];
      }
    }
  }
}

extern "C" int hostemu_qudit(void* state, int n_modes, int d, int dtype, const void* matrix, const int32_t* modes,
                             int n_targets, int64_t batch) {
  QuditGeom g;
  const char* err = "";
  if (qudit_make_geom(n_modes, d, modes, n_targets, dtype == B200Q_C64 ? 8 : 16, &g, &err)) return -1;
  if (dtype == B200Q_C64) emu_qudit<float>((std::complex<float>*)state, g, (const std::complex<float>*)matrix, batch);
  else emu_qudit<double>((std::complex<double>*)state, g, (const std::complex<double>*)matrix, batch);
  return 0;
}
