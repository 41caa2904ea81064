// sm_100a kernels and the extern "C" entry points declared in include/b200q.h.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200q.h"
#include "b200q_codegen.h"
#include "b200q_jit.h"
#include "b200q_planner.h"
#include "b200q_tile_body.h"

using namespace b200q;

struct b200q_plan {
  Plan* p;
};

namespace b200q {
thread_local std::string g_err;
int set_err(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_err(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return (int)e;
}
}  // namespace b200q

namespace {

// ------------------------------------------------------------------------------------------------
// fused tile kernel
// ------------------------------------------------------------------------------------------------
template <int CB> struct TileCfg {
  static constexpr int kThreads = 1 << (CB - B200Q_REG_CHUNK_BITS);
  static constexpr int kMinBlocks = CB >= 13 ? 1 : (CB == 12 ? 2 : 4);
};
// Resident CTAs per SM of the forward kernel.  The lean complex64 instantiation is latency-bound on warps
// (measured: 1 -> 2 CTAs/SM is 1.58x) and its shared memory (74 KiB) would allow THREE 64 KiB tiles per SM, but
// the 80-register budget that requires spills the 64 data registers in the op loop: measured 4.2 ms per pass
// instead of 2.3 ms.  Kept as a compile-time switch.
#ifndef B200Q_LEAN_BLOCKS
#define B200Q_LEAN_BLOCKS 2
#endif
template <typename Real, int CB, bool LEAN> struct FwdCfg {
  static constexpr int kMinBlocks = (LEAN && CB == 12 && sizeof(Real) == 4) ? B200Q_LEAN_BLOCKS : TileCfg<CB>::kMinBlocks;
};

// Shared memory: tile | cx pool (DIAG / MATK / slow MAT1 matrices) | MAT1 coefficient records | round tables.
template <typename Real, int CB> struct TileSmem {
  static constexpr size_t kTile = size_t(16) << CB;
  static constexpr size_t kPool = size_t(B200Q_POOL_MAX) * sizeof(cx<Real>);
  static constexpr size_t kCoef = size_t(B200Q_MAX_OPS) * B200Q_COEF_PER_OP * sizeof(Real);
  static constexpr size_t kTabs = sizeof(RoundTab) * B200Q_MAX_ROUNDS;
  static constexpr size_t kWords = sizeof(OpWord) * (B200Q_MAX_OPS + 1);
  static constexpr size_t kBaseTab = sizeof(uint64_t) * 6 * 32;     // tile index -> physical base, 5 bits at a time
  static constexpr size_t kBase = kTile + kCoef + kTabs + kWords + kBaseTab;   // the cx pool comes last: only passes with
  static constexpr size_t kTotal = kBase + kPool;                   // dense / general ops allocate it
};

// Persistent CTAs: the prologue (matrix staging, coefficient records, round address tables) runs once per
// CTA; the CTA then walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  A work item is (tile, state of the
// batch); with per-state matrices (mat_batch_stride != 0) the batch is on blockIdx.y instead.
template <typename Real, int CB, bool LEAN>
__global__ void __launch_bounds__(TileCfg<CB>::kThreads, FwdCfg<Real, CB, LEAN>::kMinBlocks)
b200q_tile_kernel(const __grid_constant__ b200q_pass_t P, typename Traits<Real>::chunk* __restrict__ state,
                  const cx<Real>* __restrict__ mats, uint64_t chunks_per_state, int64_t mat_batch_stride,
                  uint32_t tile_shift, uint64_t n_work, const __grid_constant__ b200q_remote_t remote) {
  using chunk = typename Traits<Real>::chunk;
  using SM = TileSmem<Real, CB>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  chunk* tile = reinterpret_cast<chunk*>(smem_raw);
  Real* coef = reinterpret_cast<Real*>(smem_raw + SM::kTile);
  RoundTab* tabs = reinterpret_cast<RoundTab*>(smem_raw + SM::kTile + SM::kCoef);
  OpWord* words = reinterpret_cast<OpWord*>(smem_raw + SM::kTile + SM::kCoef + SM::kTabs);
  uint64_t* base_tab = reinterpret_cast<uint64_t*>(smem_raw + SM::kTile + SM::kCoef + SM::kTabs + SM::kWords);
  cx<Real>* pool = reinterpret_cast<cx<Real>*>(smem_raw + SM::kBase);
  uint64_t* dest_tab = reinterpret_cast<uint64_t*>(smem_raw + SM::kBase + (P.needs_pool ? SM::kPool : 0));
  const int tid = threadIdx.x;
  const int nthreads = TileCfg<CB>::kThreads;
  const cx<Real>* m = mats + int64_t(blockIdx.y) * mat_batch_stride;

  fill_round_tabs<Real>(P, tid, nthreads, tabs);
  if (P.needs_pool) fill_pool<Real>(P, tid, nthreads, pool, m, false);
  fill_coefs<Real>(P, tid, nthreads, coef, m);
  fill_opwords(P, tid, nthreads, words);
  if (remote.enabled) fill_dest_tab(remote, tid, nthreads, dest_tab);
  // Per-tile setup without loops over index bits: the physical base of a tile is the OR of six table entries
  // (five tile-index bits each), and lane l of every warp keeps, in tile-index space, the outside-the-tile
  // control masks of ops l and l + 32, so the per-tile set of enabled ops is two ballots.
  for (int e = tid; e < 6 * 32; e += nthreads) {
    const int k = e >> 5, v = e & 31;
    uint64_t b = 0;
    for (int j = 0; j < 5; ++j) {
      const int q = 5 * k + j;
      if (q < int(P.n_nontile) && ((v >> j) & 1)) b |= 1ull << P.nontile_phys[q];
    }
    base_tab[e] = b;
  }
  uint32_t gmask[2] = {0u, 0u};
  for (int h = 0; h < 2; ++h) {
    const int o = (tid & 31) + 32 * h;
    if (o < int(P.n_ops) && P.ops[o].ctrl_glob) {
      for (int q = 0; q < int(P.n_nontile); ++q)
        if ((P.ops[o].ctrl_glob >> P.nontile_phys[q]) & 1ull) gmask[h] |= 1u << q;
    }
  }
  const int n_groups = (int(P.n_nontile) + 4) / 5;
  const Real gscale = P.has_scale ? Real(pass_scale<Real>(P, m)) : Real(1);
  __syncthreads();
  const int nr = P.n_rounds;
  const uint64_t tile_mask = (1ull << tile_shift) - 1ull;
  for (uint64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
    const uint32_t tile_id = uint32_t(w & tile_mask);
    uint64_t cta_base = 0;
    for (int k = 0; k < n_groups; ++k) cta_base |= base_tab[32 * k + ((tile_id >> (5 * k)) & 31u)];
    const uint64_t enabled = uint64_t(__ballot_sync(0xffffffffu, (tile_id & gmask[0]) == gmask[0])) |
                             (uint64_t(__ballot_sync(0xffffffffu, (tile_id & gmask[1]) == gmask[1])) << 32);
    chunk* gstate = state + (uint64_t(blockIdx.y) + (w >> tile_shift)) * chunks_per_state;
    for (int r = 0; r < nr; ++r) {
      const b200q_round_t& Rd = P.rounds[r];
      if (Rd.direct) {
        for (int o = Rd.op_begin; o < Rd.op_end; ++o) {
          run_direct_op<Real>(P, P.ops[o], tid, nthreads, cta_base, tile, pool);
          __syncthreads();
        }
      } else {
        run_round<Real, LEAN>(P, Rd, tabs[r], tid, cta_base, enabled, tile, pool, coef, words, gscale, gstate,
                              chunks_per_state, &remote, dest_tab);
        if (r + 1 < nr) __syncthreads();
      }
    }
    if (nr > 1) __syncthreads();   // the next tile's first round overwrites the shared-memory tile
  }
}

inline int sm_count(int dev) {
  static int cache[64] = {0};
  if (dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev] = v;
  }
  return cache[dev];
}

static_assert(sizeof(b200q_pass_t) + sizeof(b200q_remote_t) + 64 <= 4096, "kernel parameters must stay below 4 KB");

template <typename Real, int CB>
int launch_pass(const b200q_pass_t& P, void* state, const void* mats, int n_qubits, int64_t batch,
                int64_t mat_batch_stride, cudaStream_t stream, const b200q_remote_t* remote_in = nullptr) {
  b200q_remote_t remote;
  std::memset(&remote, 0, sizeof remote);
  if (remote_in) remote = *remote_in;
  using chunk = typename Traits<Real>::chunk;
  constexpr int VS = Traits<Real>::VS;
  const size_t kDest = size_t(B200Q_DEST_TAB_ENTRIES) * sizeof(uint64_t);
  const size_t smem_max = TileSmem<Real, CB>::kTotal + kDest;
  const size_t smem = (P.needs_pool ? TileSmem<Real, CB>::kTotal : TileSmem<Real, CB>::kBase) + (remote.enabled ? kDest : 0);
  auto kern = P.lean ? b200q_tile_kernel<Real, CB, true> : b200q_tile_kernel<Real, CB, false>;
  const int blocks_per_sm = P.lean ? FwdCfg<Real, CB, true>::kMinBlocks : FwdCfg<Real, CB, false>::kMinBlocks;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    int rc = cuda_err(cudaFuncSetAttribute(b200q_tile_kernel<Real, CB, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max),
                      "cudaFuncSetAttribute");
    if (!rc)
      rc = cuda_err(cudaFuncSetAttribute(b200q_tile_kernel<Real, CB, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max),
                    "cudaFuncSetAttribute");
    if (rc) return rc;
    attr_set[dev] = true;
  }
  const uint64_t chunks_per_state = (1ull << n_qubits) >> VS;
  const int tile_shift = int(P.n_bits) - int(P.tile_bits);
  const uint64_t ntiles = 1ull << tile_shift;
  static const int ctas_env = [] { const char* e = getenv("B200Q_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  // The grid is a multiple of the resident CTA count, oversubscribed: the hardware then hands out CTAs as SMs
  // free up (dynamic balance; measured on a single-gate pass: 91.6 % of the copy bandwidth with exactly the
  // resident count, 98.5 % with 16x), while every CTA still amortises its prologue over several tiles.  HBM-bound
  // passes (few ops) take the larger factor; an odd factor leaves a partial last wave (measured -12 %).
  const int oversub = P.n_ops <= 4 ? 16 : 4;
  const uint64_t resident = uint64_t(sm_count(dev)) * (ctas_env > 0 ? ctas_env : blocks_per_sm * oversub);
  if (mat_batch_stride == 0) {
    const uint64_t n_work = ntiles * uint64_t(batch);
    dim3 grid((unsigned)std::min<uint64_t>(n_work, resident), 1, 1);
    kern<<<grid, TileCfg<CB>::kThreads, smem, stream>>>(P, reinterpret_cast<chunk*>(state),
                                                        reinterpret_cast<const cx<Real>*>(mats), chunks_per_state, 0,
                                                        (uint32_t)tile_shift, n_work, remote);
  } else {
    const uint64_t gx = std::min<uint64_t>(ntiles, std::max<uint64_t>(1, resident / uint64_t(std::min<int64_t>(batch, (int64_t)resident))));
    for (int64_t b0 = 0; b0 < batch; b0 += 32768) {
      const int64_t nb = std::min<int64_t>(32768, batch - b0);
      dim3 grid((unsigned)gx, (unsigned)nb, 1);
      kern<<<grid, TileCfg<CB>::kThreads, smem, stream>>>(
          P, reinterpret_cast<chunk*>(state) + uint64_t(b0) * chunks_per_state,
          reinterpret_cast<const cx<Real>*>(mats) + b0 * mat_batch_stride, chunks_per_state, mat_batch_stride,
          (uint32_t)tile_shift, ntiles, remote);
    }
  }
  return cuda_err(cudaGetLastError(), "tile kernel launch");
}

// JIT policy: B200Q_JIT=0 disables the specialised kernels; B200Q_JIT_MIN_QUBITS (default 20) is the smallest state
// worth a compilation (below it a pass is launch-latency bound); plans are compiled lazily at their first run.
int jit_min_qubits() {
  static const int v = [] { const char* e = getenv("B200Q_JIT_MIN_QUBITS"); return e ? atoi(e) : 20; }();
  return v;
}

// ------------------------------------------------------------------------------------------------
// dense gate on 5..6 targets (UAnyGate / get_unitary-style blocks, reference gate.py:2745-2790): a pass of its own.
// The 2^k x 2^k matrix is staged once per CTA in shared memory (adjoint folded in); a CTA then walks groups of 2^k
// amplitudes, GP groups at a time: gather into shared memory, one output row per thread, scatter.  CUDA cores:
// 8 * 4^k flops per group against 2 * 2^k * B bytes -- k = 6 complex64 is 32 flop / byte, still under the FP32
// ridge of a B200 (~ 11 flop / byte) by 3x only: the tensor-core version is the next step.
// ------------------------------------------------------------------------------------------------
struct DenseArgs {
  uint64_t ctrl;        // controls, physical bits
  uint32_t mat_src;
  int32_t n_qubits, k, adjoint;
  uint8_t tbit[8];      // physical bit of matrix-index bit j
  uint8_t sorted[8];    // the same, ascending
};

template <typename Real>
__global__ void __launch_bounds__(256)
b200q_dense_kernel(cx<Real>* __restrict__ state, const cx<Real>* __restrict__ mats, const DenseArgs A,
                   uint64_t amps_per_state, int64_t mat_batch_stride) {
  extern __shared__ __align__(16) unsigned char dsm[];
  const int D = 1 << A.k;
  cx<Real>* M = reinterpret_cast<cx<Real>*>(dsm);            // [D][D], row-major, adjoint folded
  cx<Real>* xs = M + D * D;                                  // [GP][D]
  const int GP = blockDim.x / D;
  const cx<Real>* m = mats + int64_t(blockIdx.y) * mat_batch_stride + A.mat_src;
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
    const int r = e / D, c = e % D;
    cx<Real> v = A.adjoint ? m[c * D + r] : m[e];
    if (A.adjoint) v.y = -v.y;
    M[e] = v;
  }
  __syncthreads();
  cx<Real>* st = state + uint64_t(blockIdx.y) * amps_per_state;
  const int gl = threadIdx.x / D, r = threadIdx.x % D;
  uint32_t off = 0;   // offset of matrix index r inside a group
  for (int j = 0; j < A.k; ++j) off |= 0;   // (computed below as 64-bit)
  uint64_t roff = 0;
  for (int j = 0; j < A.k; ++j)
    if ((r >> j) & 1) roff |= 1ull << A.tbit[j];
  const uint64_t n_groups = amps_per_state >> A.k;
  for (uint64_t g0 = uint64_t(blockIdx.x) * GP; g0 < n_groups; g0 += uint64_t(gridDim.x) * GP) {
    const uint64_t g = g0 + gl;
    uint64_t base = g;
    for (int j = 0; j < A.k; ++j) base = ((base >> A.sorted[j]) << (A.sorted[j] + 1)) | (base & ((1ull << A.sorted[j]) - 1ull));
    const bool active = gl < GP && g < n_groups && (base & A.ctrl) == A.ctrl;
    if (active) xs[gl * D + r] = st[base | roff];
    __syncthreads();
    if (active) {
      Real yr = Real(0), yi = Real(0);
      const cx<Real>* row = M + r * D;
      const cx<Real>* x = xs + gl * D;
      for (int c = 0; c < D; ++c) {
        const cx<Real> w = row[c], v = x[c];
        yr += w.x * v.x - w.y * v.y;
        yi += w.x * v.y + w.y * v.x;
      }
      cx<Real> y; y.x = yr; y.y = yi;
      st[base | roff] = y;
    }
    __syncthreads();
  }
}

bool dense_tc_enabled() {
  const char* e = getenv("B200Q_DENSE_TC");   // read per call: tools/dense_tc_bench.py switches it for its A/B
  return !(e && atoi(e) == 0);
}

template <typename Real>
int launch_dense_pass(const b200q_pass_t& P, void* state, const void* mats, int n_qubits, int64_t batch, int64_t mbs,
                      cudaStream_t stream, bool flip_adjoint) {
  const b200q_op_t& op = P.ops[0];
  if (sizeof(Real) == 4 && n_qubits >= 12 && op.k >= 4 && dense_tc_enabled()) {   // complex64: tensor cores (b200q_dense_tc.cu)
    int32_t tg[8];
    for (int j = 0; j < op.k; ++j) tg[j] = int32_t((op.dsel_glob[0] >> (8 * j)) & 0xff);
    const int adj = (((op.flags & B200Q_FLAG_ADJOINT) != 0) != flip_adjoint) ? 1 : 0;
    for (int64_t b = 0; b < batch; ++b) {
      const int rc = b200q_dense_tc_apply(reinterpret_cast<cx<float>*>(state) + (uint64_t(b) << n_qubits), n_qubits,
                                          reinterpret_cast<const cx<float>*>(mats) + b * mbs + op.mat_src, tg, op.k,
                                          op.ctrl_glob, adj, stream);
      if (rc) return rc;
    }
    return 0;
  }
  DenseArgs A;
  std::memset(&A, 0, sizeof A);
  A.ctrl = op.ctrl_glob;
  A.mat_src = op.mat_src;
  A.n_qubits = n_qubits;
  A.k = op.k;
  A.adjoint = (((op.flags & B200Q_FLAG_ADJOINT) != 0) != flip_adjoint) ? 1 : 0;
  for (int j = 0; j < A.k; ++j) A.tbit[j] = A.sorted[j] = uint8_t((op.dsel_glob[0] >> (8 * j)) & 0xff);
  std::sort(A.sorted, A.sorted + A.k);
  const int D = 1 << A.k;
  const int threads = 256, GP = threads / D;
  const size_t smem = (size_t(D) * D + size_t(GP) * D) * sizeof(cx<Real>);
  auto kern = b200q_dense_kernel<Real>;
  int rc = cuda_err(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "cudaFuncSetAttribute(dense)");
  if (rc) return rc;
  const uint64_t n_amps = 1ull << n_qubits;
  const uint64_t n_groups = n_amps >> A.k;
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t want = (n_groups + GP - 1) / GP;
  const unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, uint64_t(sm_count(dev)) * 8));
  for (int64_t b0 = 0; b0 < batch; b0 += 32768) {
    const int64_t nb = std::min<int64_t>(32768, batch - b0);
    kern<<<dim3(gx, (unsigned)nb), threads, smem, stream>>>(
        reinterpret_cast<cx<Real>*>(state) + uint64_t(b0) * n_amps,
        reinterpret_cast<const cx<Real>*>(mats) + b0 * mbs, A, n_amps, mbs);
  }
  return cuda_err(cudaGetLastError(), "dense pass launch");
}

// Cotangent of the gate of a dense pass (reverse sweep, include/b200q.h "adjoint differentiation"):
//   G[r][c] = sum over the groups whose controls are set of lambda[r, group] * conj(psi[c, group]),
// psi already un-applied, lambda not yet.  A CTA stages kCotAmps amplitudes of both states per step, every thread
// owns D * D / 256 entries (r, c): partial sums over the staged groups in the state's precision, carried in double
// across steps, one atomicAdd per entry and CTA at the end (complex128 buffer laid out like the matrix buffer).
constexpr int kCotAmps = 2048;
template <typename Real>
__global__ void __launch_bounds__(256)
b200q_dense_cotangent_kernel(const cx<Real>* __restrict__ psi, const cx<Real>* __restrict__ lam, const DenseArgs A,
                             uint64_t n_amps, double* __restrict__ grad) {
  extern __shared__ __align__(16) unsigned char dsm[];
  const int D = 1 << A.k, GP = kCotAmps >> A.k;
  cx<Real>* xp = reinterpret_cast<cx<Real>*>(dsm);   // [GP][D + 1] psi   (+1: rows of one group on different banks)
  cx<Real>* xl = xp + GP * (D + 1);                  // [GP][D + 1] lambda
  const int n_ent = D * D, per = (n_ent + 255) / 256;   // <= 16 entries per thread
  double acc_r[16], acc_i[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc_r[j] = acc_i[j] = 0.0;
  const uint64_t n_groups = n_amps >> A.k;
  for (uint64_t g0 = uint64_t(blockIdx.x) * GP; g0 < n_groups; g0 += uint64_t(gridDim.x) * GP) {
    for (int i = threadIdx.x; i < GP * D; i += 256) {
      const int gl = i >> A.k, r = i & (D - 1);
      const uint64_t g = g0 + gl;
      uint64_t base = g;
      for (int j = 0; j < A.k; ++j)
        base = ((base >> A.sorted[j]) << (A.sorted[j] + 1)) | (base & ((1ull << A.sorted[j]) - 1ull));
      uint64_t roff = 0;
      for (int j = 0; j < A.k; ++j)
        if ((r >> j) & 1) roff |= 1ull << A.tbit[j];
      cx<Real> vp, vl;
      vp.x = vp.y = vl.x = vl.y = Real(0);
      if (g < n_groups && (base & A.ctrl) == A.ctrl) { vp = psi[base | roff]; vl = lam[base | roff]; }
      xp[gl * (D + 1) + r] = vp;
      xl[gl * (D + 1) + r] = vl;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j >= per) break;
      const int e = threadIdx.x + j * 256;
      if (e >= n_ent) break;
      const int r = e >> A.k, c = e & (D - 1);
      Real sr = Real(0), si = Real(0);
      for (int gl = 0; gl < GP; ++gl) {
        const cx<Real> l = xl[gl * (D + 1) + r], p = xp[gl * (D + 1) + c];
        sr += l.x * p.x + l.y * p.y;     // l * conj(p)
        si += l.y * p.x - l.x * p.y;
      }
      acc_r[j] += double(sr);
      acc_i[j] += double(si);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j >= per) break;
    const int e = threadIdx.x + j * 256;
    if (e >= n_ent) break;
    const int r = e >> A.k, c = e & (D - 1);
    // stored matrix M with U = M^dagger (B200Q_FLAG_ADJOINT): cotangent of M = conjugate transpose of that of U
    const int dst = A.adjoint ? c * D + r : r * D + c;
    double* out = grad + 2 * (uint64_t(A.mat_src) + dst);
    if (acc_r[j] != 0.0) atomicAdd(out, acc_r[j]);
    if (acc_i[j] != 0.0) atomicAdd(out + 1, A.adjoint ? -acc_i[j] : acc_i[j]);
  }
}

template <typename Real>
int launch_dense_cotangent(const b200q_pass_t& P, const void* psi, const void* lam, void* grad, int n_qubits,
                           cudaStream_t stream) {
  const b200q_op_t& op = P.ops[0];
  DenseArgs A;
  std::memset(&A, 0, sizeof A);
  A.ctrl = op.ctrl_glob;
  A.mat_src = op.mat_src;
  A.n_qubits = n_qubits;
  A.k = op.k;
  A.adjoint = (op.flags & B200Q_FLAG_ADJOINT) ? 1 : 0;
  for (int j = 0; j < A.k; ++j) A.tbit[j] = A.sorted[j] = uint8_t((op.dsel_glob[0] >> (8 * j)) & 0xff);
  std::sort(A.sorted, A.sorted + A.k);
  const int D = 1 << A.k, GP = kCotAmps >> A.k;
  const size_t smem = size_t(2) * GP * (D + 1) * sizeof(cx<Real>);
  auto kern = b200q_dense_cotangent_kernel<Real>;
  int rc = cuda_err(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "cudaFuncSetAttribute(dense cotangent)");
  if (rc) return rc;
  const uint64_t n_amps = 1ull << n_qubits, n_groups = n_amps >> A.k;
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t want = (n_groups + GP - 1) / GP;
  const unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, uint64_t(sm_count(dev)) * 2));
  kern<<<gx, 256, smem, stream>>>(reinterpret_cast<const cx<Real>*>(psi), reinterpret_cast<const cx<Real>*>(lam), A,
                                  n_amps, reinterpret_cast<double*>(grad));
  return cuda_err(cudaGetLastError(), "dense cotangent launch");
}

int launch_pass_any(const Plan& pl, const b200q_pass_t& P, void* state, const void* mats, int64_t batch,
                    int64_t mbs, cudaStream_t stream, const b200q_remote_t* remote = nullptr, int pass_index = -1) {
  if (P.n_rounds == 0) {   // dense pass (5..6 targets)
    if (remote && remote->enabled) return set_err(B200Q_EUNSUPPORTED, "a dense pass cannot carry the fused exchange");
    return pl.dtype == B200Q_C64 ? launch_dense_pass<float>(P, state, mats, pl.n_qubits, batch, mbs, stream, false)
                                 : launch_dense_pass<double>(P, state, mats, pl.n_qubits, batch, mbs, stream, false);
  }
  if (pass_index >= 0 && pl.jit && pl.jit->prepared) {
    const int rc = jit_launch(const_cast<Plan&>(pl), pass_index, state, mats, batch, mbs, stream, remote);
    if (rc != -1000) return rc > 0 ? cuda_err((cudaError_t)rc, "specialised pass kernel launch") : rc;
  }
  const int cb = pl.opt.chunk_bits;
  if (pl.dtype == B200Q_C64) {
    switch (cb) {
      case 11: return launch_pass<float, 11>(P, state, mats, pl.n_qubits, batch, mbs, stream, remote);
      case 12: return launch_pass<float, 12>(P, state, mats, pl.n_qubits, batch, mbs, stream, remote);
      case 13: return launch_pass<float, 13>(P, state, mats, pl.n_qubits, batch, mbs, stream, remote);
    }
  } else {
    switch (cb) {
      case 11: return launch_pass<double, 11>(P, state, mats, pl.n_qubits, batch, mbs, stream, remote);
      case 12: return launch_pass<double, 12>(P, state, mats, pl.n_qubits, batch, mbs, stream, remote);
      case 13: return launch_pass<double, 13>(P, state, mats, pl.n_qubits, batch, mbs, stream, remote);
    }
  }
  return set_err(B200Q_EUNSUPPORTED, "chunk_bits must be 11, 12 or 13");
}

// ------------------------------------------------------------------------------------------------
// adjoint (reverse) sweep kernel: two tiles (psi, lambda) per CTA
// ------------------------------------------------------------------------------------------------
template <typename Real, int CB> struct AdjSmem {
  static constexpr size_t kTiles = size_t(32) << CB;
  static constexpr size_t kPool = size_t(B200Q_POOL_MAX) * sizeof(cx<Real>);
  static constexpr size_t kAcc = size_t(B200Q_MAX_OPS) * B200Q_ACC_PER_OP * sizeof(double);
  static constexpr size_t kGfac = size_t(B200Q_MAX_OPS) * sizeof(double);
  static constexpr size_t kWacc = size_t(TileCfg<CB>::kThreads / 32) * B200Q_MAX_OPS * B200Q_WACC_PER_OP * sizeof(double);
  static constexpr size_t kTabs = sizeof(RoundTab) * B200Q_MAX_ROUNDS;
  static constexpr size_t kCoef = size_t(B200Q_MAX_OPS) * B200Q_COEF_PER_OP * sizeof(Real);
  static constexpr size_t kWords = sizeof(OpWord) * (B200Q_MAX_OPS + 1);
  static constexpr size_t kTotal = kTiles + kPool + kAcc + kGfac + kWacc + kTabs + kCoef + kWords;
};

template <typename Real, int CB>
__global__ void __launch_bounds__(TileCfg<CB>::kThreads, 1)
b200q_adjoint_kernel(const __grid_constant__ b200q_pass_t P, typename Traits<Real>::chunk* __restrict__ psi,
                     typename Traits<Real>::chunk* __restrict__ lam, const cx<Real>* __restrict__ mats,
                     double* __restrict__ grad, uint64_t want_mask, uint64_t chunks_per_state) {
  using chunk = typename Traits<Real>::chunk;
  using SM = AdjSmem<Real, CB>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  chunk* tile_psi = reinterpret_cast<chunk*>(smem_raw);
  chunk* tile_lam = reinterpret_cast<chunk*>(smem_raw + (size_t(16) << CB));
  unsigned char* q = smem_raw + SM::kTiles;
  cx<Real>* pool = reinterpret_cast<cx<Real>*>(q); q += SM::kPool;
  double* acc = reinterpret_cast<double*>(q); q += SM::kAcc;
  double* gfac = reinterpret_cast<double*>(q); q += SM::kGfac;
  double* wacc = reinterpret_cast<double*>(q); q += SM::kWacc;
  RoundTab* tabs = reinterpret_cast<RoundTab*>(q); q += SM::kTabs;
  Real* coef = reinterpret_cast<Real*>(q); q += SM::kCoef;
  OpWord* words = reinterpret_cast<OpWord*>(q);
  const int tid = threadIdx.x;
  const int nthreads = TileCfg<CB>::kThreads;
  const uint64_t cta_base = tile_base(P, blockIdx.x);
  for (int e = tid; e < int(P.n_ops) * B200Q_ACC_PER_OP; e += nthreads) acc[e] = 0.0;
  for (int e = tid; e < int(SM::kWacc / sizeof(double)); e += nthreads) wacc[e] = 0.0;
  fill_round_tabs<Real>(P, tid, nthreads, tabs);
  if (P.pool_elems) fill_pool<Real>(P, tid, nthreads, pool, mats, true);
  fill_coefs<Real>(P, tid, nthreads, coef, mats, true);
  fill_opwords(P, tid, nthreads, words);
  __shared__ double gscale_sh;
  if (tid == 0) gscale_sh = adjoint_scales<Real>(P, mats, want_mask, gfac);
  __syncthreads();
  const Real gscale = Real(gscale_sh);
  for (int r = int(P.n_rounds) - 1; r >= 0; --r) {
    const b200q_round_t& Rd = P.rounds[r];
    if (Rd.direct) {
      for (int o = int(Rd.op_end) - 1; o >= int(Rd.op_begin); --o) {
        run_direct_op_adjoint<Real>(P, P.ops[o], tid, nthreads, cta_base, tile_psi, tile_lam, pool,
                                    (want_mask >> o) & 1ull, acc + o * B200Q_ACC_PER_OP);
        __syncthreads();
      }
    } else {
      run_round_adjoint<Real>(P, Rd, tabs[r], tid, cta_base, tile_psi, tile_lam, pool, coef, words, gscale, psi, lam,
                              chunks_per_state, want_mask, wacc);
      __syncthreads();
    }
  }
  if (want_mask) {
    merge_warp_acc(P, tid, nthreads, nthreads / 32, wacc, acc);
    __syncthreads();
  }
  if (want_mask)
    flush_grad(P, tid, nthreads, want_mask, acc, gfac, grad, [](double* p, double v) { atomicAdd(p, v); });
}

template <typename Real, int CB>
int launch_adjoint_pass(const b200q_pass_t& P, void* psi, void* lam, const void* mats, void* grad, uint64_t want_mask,
                        int n_qubits, cudaStream_t stream) {
  using chunk = typename Traits<Real>::chunk;
  constexpr int VS = Traits<Real>::VS;
  const size_t smem = AdjSmem<Real, CB>::kTotal;
  auto kern = b200q_adjoint_kernel<Real, CB>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    int rc = cuda_err(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "cudaFuncSetAttribute(adjoint)");
    if (rc) return rc;
    attr_set[dev] = true;
  }
  const uint64_t chunks_per_state = (1ull << n_qubits) >> VS;
  const uint64_t ntiles = 1ull << (int(P.n_bits) - int(P.tile_bits));
  if (ntiles > 0x7fffffffull) return set_err(B200Q_EUNSUPPORTED, "too many tiles for one launch");
  kern<<<(unsigned)ntiles, TileCfg<CB>::kThreads, smem, stream>>>(
      P, reinterpret_cast<chunk*>(psi), reinterpret_cast<chunk*>(lam), reinterpret_cast<const cx<Real>*>(mats),
      reinterpret_cast<double*>(grad), want_mask, chunks_per_state);
  return cuda_err(cudaGetLastError(), "adjoint kernel launch");
}

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__device__ __forceinline__ void block_sum_atomic(double (&v)[NV], double* out) {
  __shared__ double sh[NV][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) sh[k][w] = s;
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = lane < nw ? sh[k][lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) atomicAdd(out + k, s);
    }
  }
}

// grid: (blocks, batch).  Real2 = float2 / double2 amplitudes.
template <typename Real>
__global__ void __launch_bounds__(256) norm2_kernel(const cx<Real>* __restrict__ st, uint64_t n_amps, double* out) {
  const cx<Real>* s = st + uint64_t(blockIdx.y) * n_amps;
  double acc[1] = {0.0};
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_amps; i += uint64_t(gridDim.x) * blockDim.x) {
    const cx<Real> a = s[i];
    acc[0] += double(a.x) * double(a.x) + double(a.y) * double(a.y);
  }
  block_sum_atomic<1>(acc, out + blockIdx.y);
}

template <typename Real>
__global__ void __launch_bounds__(256)
inner_kernel(const cx<Real>* __restrict__ bra, const cx<Real>* __restrict__ ket, uint64_t n_amps, double* out) {
  const cx<Real>* a = bra + uint64_t(blockIdx.y) * n_amps;
  const cx<Real>* b = ket + uint64_t(blockIdx.y) * n_amps;
  double acc[2] = {0.0, 0.0};
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_amps; i += uint64_t(gridDim.x) * blockDim.x) {
    const cx<Real> x = a[i], y = b[i];
    acc[0] += double(x.x) * double(y.x) + double(x.y) * double(y.y);  // conj(x) * y
    acc[1] += double(x.x) * double(y.y) - double(x.y) * double(y.x);
  }
  block_sum_atomic<2>(acc, out + 2 * blockIdx.y);
}

// Z-string expectations.  Each thread strides over amplitudes; masks are processed in groups of 8
// accumulators so the state is re-read once per 8 masks only when n_masks > 8 (L2 absorbs nothing at
// 2^30 amplitudes, hence the grouping is done INSIDE the amplitude loop: |a|^2 is computed once).
constexpr int kZGroup = 16;
// NK: masks handled by this instantiation (1, 2, 4, 8 or 16): the per-amplitude work is NK parity tests, so a
// circuit with two observables does an eighth of the work of the 16-wide sweep.  complex64: |a|^2 is formed in
// float (relative 6e-8, below the state's own rounding) and accumulated in double.
template <typename Real, int NK>
__global__ void __launch_bounds__(256)
expz_kernel(const cx<Real>* __restrict__ st, uint64_t n_amps, const uint64_t* __restrict__ masks, int n_masks,
            int mask0, uint64_t index_offset, double* out) {
  const cx<Real>* s = st + uint64_t(blockIdx.y) * n_amps;
  uint64_t mk[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) mk[k] = (mask0 + k < n_masks) ? masks[mask0 + k] : 0ull;
  double acc[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) acc[k] = 0.0;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_amps; i += uint64_t(gridDim.x) * blockDim.x) {
    const cx<Real> a = s[i];
    const double p = sizeof(Real) == 4 ? double(float(a.x) * float(a.x) + float(a.y) * float(a.y))
                                       : double(a.x) * double(a.x) + double(a.y) * double(a.y);
    const uint64_t idx = i | index_offset;
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[k] += (__popcll(idx & mk[k]) & 1) ? -p : p;
  }
  double* o = out + uint64_t(blockIdx.y) * n_masks + mask0;
  __shared__ double sh[NK][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) sh[k][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < NK && mask0 + (int)threadIdx.x < n_masks) {
    double v = 0.0;
    for (int j = 0; j < 8; ++j) v += sh[threadIdx.x][j];
    atomicAdd(o + threadIdx.x, v);
  }
}

template <typename Real>
void launch_expz(dim3 grid, cudaStream_t s, const void* state, uint64_t n, const uint64_t* masks, int n_masks, int m0,
                 uint64_t index_offset, double* out) {
  const int left = n_masks - m0;
  const cx<Real>* st = (const cx<Real>*)state;
  if (left >= 9) expz_kernel<Real, 16><<<grid, 256, 0, s>>>(st, n, masks, n_masks, m0, index_offset, out);
  else if (left >= 5) expz_kernel<Real, 8><<<grid, 256, 0, s>>>(st, n, masks, n_masks, m0, index_offset, out);
  else if (left >= 3) expz_kernel<Real, 4><<<grid, 256, 0, s>>>(st, n, masks, n_masks, m0, index_offset, out);
  else if (left == 2) expz_kernel<Real, 2><<<grid, 256, 0, s>>>(st, n, masks, n_masks, m0, index_offset, out);
  else expz_kernel<Real, 1><<<grid, 256, 0, s>>>(st, n, masks, n_masks, m0, index_offset, out);
}

template <typename Real>
__global__ void __launch_bounds__(256)
zweights_kernel(const cx<Real>* __restrict__ st, cx<Real>* __restrict__ lam, uint64_t n_amps,
                const uint64_t* __restrict__ masks, const double* __restrict__ weights, int n_masks,
                uint64_t index_offset) {
  extern __shared__ unsigned char zw_raw[];
  uint64_t* smk = reinterpret_cast<uint64_t*>(zw_raw);
  double* sw = reinterpret_cast<double*>(smk + n_masks);
  for (int k = threadIdx.x; k < n_masks; k += blockDim.x) {
    smk[k] = masks[k];
    sw[k] = weights[uint64_t(blockIdx.y) * n_masks + k];
  }
  __syncthreads();
  const cx<Real>* s = st + uint64_t(blockIdx.y) * n_amps;
  cx<Real>* l = lam + uint64_t(blockIdx.y) * n_amps;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_amps; i += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t idx = i | index_offset;
    double f = 0.0;
    for (int k = 0; k < n_masks; ++k) f += (__popcll(idx & smk[k]) & 1) ? -sw[k] : sw[k];
    const cx<Real> a = s[i];
    cx<Real> r;
    r.x = Real(double(a.x) * f);
    r.y = Real(double(a.y) * f);
    l[i] = r;
  }
}

template <typename Real>
__global__ void __launch_bounds__(256) init_basis_kernel(cx<Real>* st, uint64_t n_amps, uint64_t basis) {
  cx<Real>* s = st + uint64_t(blockIdx.y) * n_amps;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_amps; i += uint64_t(gridDim.x) * blockDim.x) {
    cx<Real> v;
    v.x = (i == basis) ? Real(1) : Real(0);
    v.y = Real(0);
    s[i] = v;
  }
}

inline unsigned reduce_blocks(uint64_t n_amps) {
  const uint64_t want = (n_amps + 256 * 8 - 1) / (256 * 8);
  return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, 148ull * 8));
}

int check_state_args(const void* state, int n_qubits, int dtype, int64_t batch) {
  if (!state) return set_err(B200Q_EINVAL, "null state pointer");
  if (n_qubits < 1 || n_qubits > B200Q_MAX_QUBITS - 2) return set_err(B200Q_EINVAL, "n_qubits out of range");
  if (dtype != B200Q_C64 && dtype != B200Q_C128) return set_err(B200Q_EINVAL, "bad dtype");
  if (batch < 1 || batch > 65535) return set_err(B200Q_EINVAL, "batch must be in 1..65535");
  return 0;
}

}  // namespace

// ================================================================================================
// extern "C"
// ================================================================================================
extern "C" {

const char* b200q_version(void) { return "b200q 0.1.0 (sm_100a)"; }
const char* b200q_last_error(void) { return g_err.c_str(); }

int b200q_device_check(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return set_err(B200Q_ENODEVICE, "no CUDA device visible; b200q has no CPU fallback");
  }
  if (device < 0 || device >= count) return set_err(B200Q_ENODEVICE, "device index out of range");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (major != 10) {
    char buf[128];
    snprintf(buf, sizeof buf, "device %d is sm_%d%d; b200q is built for sm_100a only", device, major, minor);
    return set_err(B200Q_ENODEVICE, buf);
  }
  return 0;
}

int b200q_plan_create(int n_qubits, int dtype, const b200q_gate_t* gates, int n_gates,
                      const b200q_plan_options_t* options, b200q_plan_t** plan_out) {
  if (!plan_out) return set_err(B200Q_EINVAL, "plan_out is null");
  *plan_out = nullptr;
  PlanOptions opt;
  if (options) {
    if (options->chunk_bits) opt.chunk_bits = options->chunk_bits;
    if (options->low_bits) opt.low_bits = options->low_bits;
    if (options->max_rounds) opt.max_rounds = options->max_rounds;
    opt.fuse = options->fuse;
    if (options->reserved[0]) opt.structured = 0;   // A/B switch: general op codes only
    if (options->reserved[1]) opt.coalesce_bits = options->reserved[1] - 1;   // experiment: 1 + lane-owned chunk bits
  }
  {   // plans that the specialised kernels will run (see jit_min_qubits) schedule the hinted phase gates for them
    std::string why;
    opt.free_phase = (opt.structured && n_qubits >= jit_min_qubits() && jit_available(&why) == 0) ? 1 : 0;
    if (const char* e = getenv("B200Q_FREE_PHASE")) opt.free_phase = atoi(e);
  }
  if (const char* e = getenv("B200Q_DEFER_DIAG")) opt.defer_diag = atoi(e);
  if (const char* e = getenv("B200Q_MIN_ROUND_GATES")) opt.min_round_gates = atoi(e);
  if (const char* e = getenv("B200Q_XC1_PENALTY")) opt.xc1_penalty = atoi(e);
  if (const char* e = getenv("B200Q_COALESCE_BITS")) {
    opt.coalesce_bits = atoi(e);
  }
  if (opt.chunk_bits < 11 || opt.chunk_bits > 13) return set_err(B200Q_EINVAL, "chunk_bits must be 11, 12 or 13");
  std::string err;
  Plan* p = make_plan(n_qubits, dtype, gates, n_gates, opt, &err);
  if (!p) return set_err(B200Q_EINVAL, err);
  *plan_out = new b200q_plan{p};
  return 0;
}

void b200q_plan_destroy(b200q_plan_t* plan) {
  if (!plan) return;
  delete plan->p;
  delete plan;
}

int b200q_plan_get_stats(const b200q_plan_t* plan, b200q_plan_stats_t* s) {
  if (!plan || !s) return set_err(B200Q_EINVAL, "null argument");
  const Plan& p = *plan->p;
  s->n_gates = p.stats.n_gates;
  s->n_passes = p.stats.n_passes;
  s->n_rounds = p.stats.n_rounds;
  s->n_ops = p.stats.n_ops;
  s->n_direct_ops = p.stats.n_direct;
  s->tile_bits = std::min(p.opt.chunk_bits + (p.dtype == B200Q_C64 ? 1 : 0), p.n_bits);
  s->threads_per_cta = 1 << (p.opt.chunk_bits - B200Q_REG_CHUNK_BITS);
  s->smem_bytes = (16 << p.opt.chunk_bits) + B200Q_POOL_MAX * (p.dtype == B200Q_C64 ? 8 : 16) +
                  B200Q_MAX_OPS * B200Q_COEF_PER_OP * (p.dtype == B200Q_C64 ? 4 : 8) + 16 * (B200Q_MAX_OPS + 1) + (int)sizeof(RoundTab) * B200Q_MAX_ROUNDS;
  return 0;
}

int b200q_plan_pass_gates(const b200q_plan_t* plan, int i) {
  if (!plan || i < 0 || i >= (int)plan->p->passes.size()) return set_err(B200Q_EINVAL, "bad pass index");
  return plan->p->pass_gate_count[i];
}

int b200q_plan_pass_gate_ids(const b200q_plan_t* plan, int i, int32_t* out, int cap) {
  if (!plan || i < 0 || i >= (int)plan->p->passes.size()) return set_err(B200Q_EINVAL, "bad pass index");
  const b200q_pass_t& P = plan->p->passes[i];
  int n = 0;
  for (int o = 0; o < P.n_ops; ++o) {
    const int32_t id = (int32_t)P.ops[o].gate_id;
    bool seen = false;
    for (int q = 0; q < n && q < cap && !seen; ++q) seen = out && out[q] == id;
    if (seen) continue;
    if (out && n < cap) out[n] = id;
    ++n;
  }
  return n;
}

int b200q_plan_export(const b200q_plan_t* plan, void* buf, size_t buf_size, size_t* needed) {
  if (!plan) return set_err(B200Q_EINVAL, "null plan");
  const size_t need = plan->p->passes.size() * sizeof(b200q_pass_t);
  if (needed) *needed = need;
  if (buf && buf_size >= need && need) std::memcpy(buf, plan->p->passes.data(), need);
  return 0;
}

int b200q_plan_run_range(const b200q_plan_t* plan, int first, int last, void* state, const void* matrices,
                         int64_t batch, int64_t mbs, void* stream) {
  if (!plan) return set_err(B200Q_EINVAL, "null plan");
  const Plan& p = *plan->p;
  int rc = check_state_args(state, p.n_qubits, p.dtype, batch);
  if (rc) return rc;
  if (first < 0 || last > (int)p.passes.size() || first > last) return set_err(B200Q_EINVAL, "bad pass range");
  if (!matrices && p.stats.n_ops) {
    bool need = false;
    for (const auto& ps : p.passes) need |= ps.pool_elems != 0;
    if (need) return set_err(B200Q_EINVAL, "null matrix buffer");
  }
  if ((!p.jit || !p.jit->prepared) && p.n_qubits >= jit_min_qubits()) jit_prepare(*plan->p, 0, false);
  for (int i = first; i < last; ++i) {
    rc = launch_pass_any(p, p.passes[i], state, matrices, batch, mbs, (cudaStream_t)stream, nullptr, i);
    if (rc) return rc;
  }
  return 0;
}

int b200q_plan_run_exchange(const b200q_plan_t* plan, void* state, const void* matrices, void* const* peer_buffers,
                            int n_ranks, int rank, const uint8_t* perm, void* stream) {
  if (!plan) return set_err(B200Q_EINVAL, "null plan");
  const Plan& p = *plan->p;
  int rc = check_state_args(state, p.n_qubits, p.dtype, 1);
  if (rc) return rc;
  if (!peer_buffers || n_ranks < 1 || n_ranks > B200Q_MAX_RANKS || (n_ranks & (n_ranks - 1)) || rank < 0 || rank >= n_ranks)
    return set_err(B200Q_EINVAL, "bad rank arguments (1, 2, 4 or 8 ranks)");
  int g = 0;
  while ((1 << g) < n_ranks) ++g;
  const int vs = p.dtype == B200Q_C64 ? 1 : 0;
  const int nl = p.n_qubits, nt = nl + g;
  if (nl - g - vs < 0 || p.n_bits != p.n_qubits || nl - vs > 40) return set_err(B200Q_EUNSUPPORTED, "shard size not supported by the fused exchange");
  if (p.passes.empty()) return set_err(B200Q_EINVAL, "empty plan");
  b200q_remote_t R;
  std::memset(&R, 0, sizeof R);
  for (int r = 0; r < n_ranks; ++r) {
    if (!peer_buffers[r]) return set_err(B200Q_EINVAL, "null peer buffer");
    R.peer[r] = peer_buffers[r];
  }
  // bit permutation of the distributed index (amplitude bits; bits >= nl are rank bits); default: block transpose
  uint8_t pm[48];
  for (int j = 0; j < nt; ++j) pm[j] = perm ? perm[j] : (uint8_t)j;
  if (!perm)
    for (int k = 0; k < g; ++k) { pm[nl - g + k] = (uint8_t)(nl + k); pm[nl + k] = (uint8_t)(nl - g + k); }
  uint64_t seen = 0;
  for (int j = 0; j < nt; ++j) {
    if (pm[j] >= nt || (seen >> pm[j] & 1)) return set_err(B200Q_EINVAL, "perm is not a permutation of the index bits");
    seen |= 1ull << pm[j];
  }
  if (vs && pm[0] != 0) return set_err(B200Q_EUNSUPPORTED, "complex64: index bit 0 (inside a 16-byte chunk) must stay in place");
  R.n_chunk_bits = nl - vs;
  for (int j = vs; j < nl; ++j) R.perm[j - vs] = (uint8_t)(pm[j] - vs);
  for (int k = 0; k < g; ++k)
    if ((rank >> k) & 1) {
      const int pos = pm[nl + k] - vs;
      R.base |= pos < R.n_chunk_bits ? (1ull << pos) : (1ull << (B200Q_DEST_RANK_SHIFT + pos - R.n_chunk_bits));
    }
  R.enabled = 1;
  const int last = (int)p.passes.size() - 1;
  if (p.dtype == B200Q_C64 && (p.passes[last].layout & B200Q_LAYOUT_DST_SOA))
    return set_err(B200Q_EUNSUPPORTED, "last pass does not write the caller layout");
  if (p.n_qubits >= jit_min_qubits() && (!p.jit || !p.jit->prepared || !p.jit->remote[last])) jit_prepare(*plan->p, 0, true);
  for (int i = 0; i <= last; ++i) {
    rc = launch_pass_any(p, p.passes[i], state, matrices, 1, 0, (cudaStream_t)stream, i == last ? &R : nullptr, i);
    if (rc) return rc;
  }
  return 0;
}

int b200q_plan_codegen(const b200q_plan_t* plan, int pass_index, int remote, char* buf, size_t buf_size, size_t* needed,
                       size_t* smem_bytes) {
  if (!plan || pass_index < 0 || pass_index >= (int)plan->p->passes.size()) return set_err(B200Q_EINVAL, "bad pass index");
  const Plan& p = *plan->p;
  if (!codegen_supported(p, p.passes[pass_index]))
    return set_err(B200Q_EUNSUPPORTED, "pass not covered by the kernel generator (small or padded state)");
  GenOptions go;
  go.remote = remote ? 1 : 0;
  go.min_blocks = p.opt.chunk_bits >= 13 ? 1 : (p.opt.chunk_bits == 12 ? 2 : 4);
  if (const char* e = getenv("B200Q_JIT_PREFETCH")) go.prefetch = atoi(e);
  if (const char* e = getenv("B200Q_JIT_IPT")) go.items_per_thread = atoi(e);
  if (go.items_per_thread == 2 && p.opt.chunk_bits == 12) go.min_blocks = 3;
  if (const char* e = getenv("B200Q_JIT_DEBUG_SKIP_OPS")) go.debug_skip_ops = atoi(e);
  if (const char* e = getenv("B200Q_JIT_DEBUG_ONE_TILE")) go.debug_one_tile = atoi(e);
  size_t smem = 0;
  std::string stats;
  const std::string src = codegen_pass(p, p.passes[pass_index], go, &smem, &stats);
  if (needed) *needed = src.size() + 1;
  if (smem_bytes) *smem_bytes = smem;
  if (buf && buf_size >= src.size() + 1) std::memcpy(buf, src.c_str(), src.size() + 1);
  g_err = stats;   // generator statistics of the pass, readable through b200q_last_error()
  return 0;
}

int b200q_plan_compile(b200q_plan_t* plan, int threads, int with_exchange_variant) {
  if (!plan) return set_err(B200Q_EINVAL, "null plan");
  std::string why;
  if (jit_available(&why)) return set_err(B200Q_EUNSUPPORTED, "run-time compilation unavailable: " + why);
  return jit_prepare(*plan->p, threads, with_exchange_variant != 0);
}

int b200q_plan_jit_status(const b200q_plan_t* plan, int32_t* n_specialised, int32_t* n_failed) {
  if (!plan) return set_err(B200Q_EINVAL, "null plan");
  const Plan& p = *plan->p;
  int ok = 0, bad = 0;
  if (p.jit) {
    for (const auto& k : p.jit->local) {
      if (k && k->compiled) ++ok;
      else if (k) ++bad;
    }
  }
  if (n_specialised) *n_specialised = ok;
  if (n_failed) *n_failed = bad;
  if (bad && p.jit) {
    for (const auto& k : p.jit->local)
      if (k && !k->compiled) { g_err = k->log; break; }
  }
  return 0;
}

int b200q_plan_run(const b200q_plan_t* plan, void* state, const void* matrices, int64_t batch, int64_t mbs,
                   void* stream) {
  if (!plan) return set_err(B200Q_EINVAL, "null plan");
  return b200q_plan_run_range(plan, 0, (int)plan->p->passes.size(), state, matrices, batch, mbs, stream);
}

int b200q_apply_gate(void* state, int n_qubits, int dtype, int kind, const void* matrix, const int32_t* targets,
                     int n_targets, const int32_t* controls, int n_controls, int adjoint, int64_t batch, int64_t mbs,
                     void* stream) {
  int rc = check_state_args(state, n_qubits, dtype, batch);
  if (rc) return rc;
  if (!targets || n_targets < 1 || n_targets > B200Q_MAX_TARGETS) return set_err(B200Q_EINVAL, "bad targets");
  if (n_controls < 0 || (n_controls > 0 && !controls)) return set_err(B200Q_EINVAL, "bad controls");
  b200q_gate_t g;
  std::memset(&g, 0, sizeof g);
  g.kind = kind;
  g.n_targets = n_targets;
  for (int j = 0; j < n_targets; ++j) g.targets[j] = targets[j];
  for (int j = 0; j < n_controls; ++j) {
    if (controls[j] < 0 || controls[j] >= n_qubits) return set_err(B200Q_EINVAL, "control out of range");
    g.controls |= 1ull << controls[j];
  }
  g.flags = adjoint ? B200Q_GATE_ADJOINT : 0;
  PlanOptions opt;
  std::string err;
  Plan* p = make_plan(n_qubits, dtype, &g, 1, opt, &err);
  if (!p) return set_err(B200Q_EINVAL, err);
  for (const auto& ps : p->passes) {
    rc = launch_pass_any(*p, ps, state, matrix, batch, mbs, (cudaStream_t)stream);
    if (rc) break;
  }
  delete p;
  return rc;
}

#define B200Q_DISPATCH_REAL(dtype, CALL)          \
  do {                                            \
    if ((dtype) == B200Q_C64) { using Real = float; CALL; } \
    else { using Real = double; CALL; }           \
  } while (0)

int b200q_norm2(const void* state, int n_qubits, int dtype, int64_t batch, double* out_dev, void* stream) {
  int rc = check_state_args(state, n_qubits, dtype, batch);
  if (rc) return rc;
  if (!out_dev) return set_err(B200Q_EINVAL, "null output");
  cudaStream_t s = (cudaStream_t)stream;
  const uint64_t n = 1ull << n_qubits;
  rc = cuda_err(cudaMemsetAsync(out_dev, 0, sizeof(double) * batch, s), "memset");
  if (rc) return rc;
  dim3 grid(reduce_blocks(n), (unsigned)batch);
  B200Q_DISPATCH_REAL(dtype, (norm2_kernel<Real><<<grid, 256, 0, s>>>((const cx<Real>*)state, n, out_dev)));
  return cuda_err(cudaGetLastError(), "norm2 launch");
}

int b200q_inner_product(const void* bra, const void* ket, int n_qubits, int dtype, int64_t batch, double* out_dev,
                        void* stream) {
  int rc = check_state_args(bra, n_qubits, dtype, batch);
  if (rc) return rc;
  if (!ket || !out_dev) return set_err(B200Q_EINVAL, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const uint64_t n = 1ull << n_qubits;
  rc = cuda_err(cudaMemsetAsync(out_dev, 0, 2 * sizeof(double) * batch, s), "memset");
  if (rc) return rc;
  dim3 grid(reduce_blocks(n), (unsigned)batch);
  B200Q_DISPATCH_REAL(dtype, (inner_kernel<Real><<<grid, 256, 0, s>>>((const cx<Real>*)bra, (const cx<Real>*)ket, n,
                                                                      out_dev)));
  return cuda_err(cudaGetLastError(), "inner product launch");
}

int b200q_expectation_z(const void* state, int n_qubits, int dtype, int64_t batch, const uint64_t* masks_dev,
                        int n_masks, uint64_t index_offset, double* out_dev, void* stream) {
  int rc = check_state_args(state, n_qubits, dtype, batch);
  if (rc) return rc;
  if (!masks_dev || !out_dev || n_masks < 1) return set_err(B200Q_EINVAL, "bad mask arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const uint64_t n = 1ull << n_qubits;
  rc = cuda_err(cudaMemsetAsync(out_dev, 0, sizeof(double) * batch * n_masks, s), "memset");
  if (rc) return rc;
  dim3 grid(reduce_blocks(n), (unsigned)batch);
  for (int m0 = 0; m0 < n_masks; m0 += kZGroup) {
    B200Q_DISPATCH_REAL(dtype, (launch_expz<Real>(grid, s, state, n, masks_dev, n_masks, m0, index_offset, out_dev)));
  }
  return cuda_err(cudaGetLastError(), "expectation_z launch");
}

int b200q_apply_z_weights(const void* state, void* lambda_out, int n_qubits, int dtype, int64_t batch,
                          const uint64_t* masks_dev, const double* weights_dev, int n_masks, uint64_t index_offset,
                          void* stream) {
  int rc = check_state_args(state, n_qubits, dtype, batch);
  if (rc) return rc;
  if (!lambda_out || !masks_dev || !weights_dev || n_masks < 1 || n_masks > 2048)
    return set_err(B200Q_EINVAL, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const uint64_t n = 1ull << n_qubits;
  dim3 grid(reduce_blocks(n), (unsigned)batch);
  const size_t sm = size_t(n_masks) * 16;
  B200Q_DISPATCH_REAL(dtype, (zweights_kernel<Real><<<grid, 256, sm, s>>>((const cx<Real>*)state, (cx<Real>*)lambda_out,
                                                                          n, masks_dev, weights_dev, n_masks,
                                                                          index_offset)));
  return cuda_err(cudaGetLastError(), "apply_z_weights launch");
}

int b200q_init_basis(void* state, int n_qubits, int dtype, int64_t batch, uint64_t basis_index, void* stream) {
  int rc = check_state_args(state, n_qubits, dtype, batch);
  if (rc) return rc;
  const uint64_t n = 1ull << n_qubits;
  if (basis_index >= n) return set_err(B200Q_EINVAL, "basis index out of range");
  dim3 grid(reduce_blocks(n), (unsigned)batch);
  B200Q_DISPATCH_REAL(dtype,
                      (init_basis_kernel<Real><<<grid, 256, 0, (cudaStream_t)stream>>>((cx<Real>*)state, n, basis_index)));
  return cuda_err(cudaGetLastError(), "init_basis launch");
}

int b200q_adjoint_run(const b200q_plan_t* plan, void* psi, void* lambda, const void* matrices, void* grad_out,
                      const uint8_t* need_grad_host, void* stream) {
  if (!plan) return set_err(B200Q_EINVAL, "null plan");
  const Plan& p = *plan->p;
  int rc = check_state_args(psi, p.n_qubits, p.dtype, 1);
  if (rc) return rc;
  if (!lambda || !matrices || !grad_out) return set_err(B200Q_EINVAL, "null argument");
  const int cb = p.opt.chunk_bits;
  if (cb > 12) return set_err(B200Q_EUNSUPPORTED, "the adjoint sweep holds two tiles per CTA: plan with chunk_bits <= 12");
  for (int i = (int)p.passes.size() - 1; i >= 0; --i) {
    const b200q_pass_t& P = p.passes[i];
    if (P.n_rounds == 0) {   // dense pass: psi <- U^dagger psi, cotangent from (lambda, psi), lambda <- U^dagger lambda
      const bool need = need_grad_host ? need_grad_host[P.ops[0].gate_id] != 0 : true;
      for (int which = 0; which < 2; ++which) {
        void* s = which == 0 ? psi : lambda;
        if (which == 1 && need) {
          rc = p.dtype == B200Q_C64
                   ? launch_dense_cotangent<float>(P, psi, lambda, grad_out, p.n_qubits, (cudaStream_t)stream)
                   : launch_dense_cotangent<double>(P, psi, lambda, grad_out, p.n_qubits, (cudaStream_t)stream);
          if (rc) return rc;
        }
        rc = p.dtype == B200Q_C64 ? launch_dense_pass<float>(P, s, matrices, p.n_qubits, 1, 0, (cudaStream_t)stream, true)
                                  : launch_dense_pass<double>(P, s, matrices, p.n_qubits, 1, 0, (cudaStream_t)stream, true);
        if (rc) return rc;
      }
      continue;
    }
    uint64_t want = 0;
    for (int o = 0; o < P.n_ops; ++o) {
      const b200q_op_t& op = P.ops[o];
      if (op.kind == B200Q_OP_X) continue;
      const bool need = need_grad_host ? need_grad_host[op.gate_id] != 0 : true;
      if (!need) continue;
      if (op.kind == B200Q_OP_MATK && op.k > 2) {
        if (need_grad_host)
          return set_err(B200Q_EUNSUPPORTED, "gradient of a dense gate on 3-4 targets inside a fused pass: plan the gate "
                                             "with B200Q_GATE_GRAD (a pass of its own)");
        continue;
      }
      want |= 1ull << o;
    }
    if (p.dtype == B200Q_C64) {
      rc = cb == 11 ? launch_adjoint_pass<float, 11>(P, psi, lambda, matrices, grad_out, want, p.n_qubits, (cudaStream_t)stream)
                    : launch_adjoint_pass<float, 12>(P, psi, lambda, matrices, grad_out, want, p.n_qubits, (cudaStream_t)stream);
    } else {
      rc = cb == 11 ? launch_adjoint_pass<double, 11>(P, psi, lambda, matrices, grad_out, want, p.n_qubits, (cudaStream_t)stream)
                    : launch_adjoint_pass<double, 12>(P, psi, lambda, matrices, grad_out, want, p.n_qubits, (cudaStream_t)stream);
    }
    if (rc) return rc;
  }
  return 0;
}

}  // extern "C"
