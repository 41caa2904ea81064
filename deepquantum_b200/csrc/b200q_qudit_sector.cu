// Structured qudit (Fock tensor) gates: evolve_state(state, matrix, nmode, wires, qudit = cutoff) as called by the
// photonic back-end (reference photonic/operation.py:142-146 -> qmath.py:485-506) for gate CLASSES whose Fock matrix
// has a known block structure (never decided from the values):
//
//   NUMBER      two-mode gates that conserve the photon number i + j of their modes -- the whole beamsplitter family
//               (photonic/gate.py:202-1012) and the cross-Kerr gate: the d^2 x d^2 matrix is block-diagonal over the
//               2d - 1 sectors i + j = s, block sizes 1, 2, ..., d, ..., 2, 1 (6.7 % of the entries at cutoff 10);
//   DIFFERENCE  two-mode squeezing (photonic/gate.py:1157-1333) conserves i - j: same sector sizes along the other diagonal;
//   DENSE1      one-mode gates (squeezer, displacement): one sector, the d x d matrix;
//   DIAG        phase shifter, Kerr, cross-Kerr: one multiplication per amplitude.
//
// The members of a sector of one fibre (fixed digits of all other modes) lie on a line in memory,
// base + t * step, t = 0 .. m-1.  A THREAD owns one sector of one fibre at a time: it loads the m amplitudes into
// registers, multiplies by the m x m block (broadcast reads of the packed blocks from shared memory, packed f32x2
// FMAs) and stores the m results in place -- no shared-memory staging of amplitudes, no barrier, no index table,
// ~30 instructions per amplitude against ~260 of the generic ELL kernel (b200q_qudit.cu), which stays the path for
// unstructured matrices.  The 32 lanes of a warp hold 32 consecutive fibres, so that every load / store of a warp is
// one contiguous run of 32 amplitudes whenever the lowest mode is not a target; a warp walks all sectors of its fibres,
// so lines are re-touched by the same warp (L1) when it is.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/b200q.h"
#include "b200q_qudit_geom.h"

namespace b200q {
int set_err(int code, const std::string& msg);
int cuda_err(cudaError_t e, const char* what);
}  // namespace b200q
using b200q::cuda_err;
using b200q::set_err;

namespace {

constexpr int kMaxSecD = 16;
constexpr int kMaxSectors = 2 * kMaxSecD - 1;
constexpr int kSecThreads = 256;

template <typename Real> struct cxs { Real x, y; };

struct SectorTab {
  int32_t n_sectors, total_w, d, k, dj;
  int64_t S0, S1, step;       // strides of the digits (i, j) of the matrix index r = i * d + j (k = 1: only i), member step
  struct Sec { uint8_t i0, j0, m, mp; uint32_t woff; } s[kMaxSectors];
};

// sectors of a structure; returns 0 if the structure / size is not handled here
int make_sectors(int structure, int d, int k, int64_t S0, int64_t S1, SectorTab* T) {
  std::memset(T, 0, sizeof *T);
  T->d = d; T->k = k; T->S0 = S0; T->S1 = S1;
  if (d > kMaxSecD) return 0;
  auto add = [&](int i0, int j0, int m) {
    SectorTab::Sec& e = T->s[T->n_sectors++];
    e.i0 = (uint8_t)i0; e.j0 = (uint8_t)j0; e.m = (uint8_t)m; e.mp = (uint8_t)((m + 1) & ~1);
    e.woff = (uint32_t)T->total_w;
    T->total_w += m * e.mp;
  };
  if (structure == B200Q_QUDIT_DENSE1 && k == 1) {
    T->dj = 0; T->step = S0;
    add(0, 0, d);
  } else if (structure == B200Q_QUDIT_NUMBER && k == 2) {
    T->dj = -1; T->step = S0 - S1;
    for (int s = 0; s <= 2 * d - 2; ++s) {
      const int i0 = s < d ? 0 : s - d + 1, i1 = s < d ? s : d - 1;
      add(i0, s - i0, i1 - i0 + 1);
    }
  } else if (structure == B200Q_QUDIT_DIFFERENCE && k == 2) {
    T->dj = 1; T->step = S0 + S1;
    for (int q = -(d - 1); q <= d - 1; ++q) {
      if (q >= 0) add(q, 0, d - q); else add(0, -q, d + q);
    }
  } else {
    return 0;
  }
  return 1;
}

// packed blocks: W[woff + r * mp + c] = post[row(r)] * M[row(r)][row(c)] * pre[row(c)], row(t) = matrix index of member t
// of the sector; `pre` / `post` (or null): diagonal gates applied right before / after the gate, folded into its blocks
template <typename Real>
__global__ void pack_blocks_kernel(const cxs<Real>* __restrict__ m, const SectorTab T, cxs<Real>* __restrict__ out,
                                   const cxs<Real>* __restrict__ pre = nullptr,
                                   const cxs<Real>* __restrict__ post = nullptr) {
  const int D = T.k == 2 ? T.d * T.d : T.d;
  for (int e = threadIdx.x + blockIdx.x * blockDim.x; e < T.total_w; e += blockDim.x * gridDim.x) {
    int s = 0;
    while (s + 1 < T.n_sectors && (uint32_t)e >= T.s[s + 1].woff) ++s;
    const int local = e - (int)T.s[s].woff, r = local / T.s[s].mp, c = local % T.s[s].mp;
    cxs<Real> v; v.x = v.y = Real(0);
    if (c < T.s[s].m) {
      const int ir = T.s[s].i0 + r, jr = T.s[s].j0 + T.dj * r, ic = T.s[s].i0 + c, jc = T.s[s].j0 + T.dj * c;
      const int row = T.k == 2 ? ir * T.d + jr : ir, col = T.k == 2 ? ic * T.d + jc : ic;
      v = m[row * D + col];
      if (pre) { const cxs<Real> w = pre[col]; const Real x = v.x * w.x - v.y * w.y; v.y = v.x * w.y + v.y * w.x; v.x = x; }
      if (post) { const cxs<Real> w = post[row]; const Real x = v.x * w.x - v.y * w.y; v.y = v.x * w.y + v.y * w.x; v.x = x; }
    }
    out[e] = v;
  }
}

// ---- one sector of one fibre in registers ---------------------------------------------------------------------
template <typename Real, int M> struct SectorOp;

// acc += (w, w) * x on a packed pair (ptxas folds the broadcast into the FFMA2 operand form)
__device__ __forceinline__ void fma_bc(unsigned long long& acc, float w, unsigned long long x) {
  unsigned long long ww;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ww), "l"(x));
}

template <int M> struct SectorOp<float, M> {
  static __device__ __forceinline__ void run(cxs<float>* p, long long step, const cxs<float>* W) {
    constexpr int MP = (M + 1) & ~1;
    unsigned long long x[M];
#pragma unroll
    for (int t = 0; t < M; ++t) x[t] = *reinterpret_cast<const unsigned long long*>(p + t * step);
#pragma unroll
    for (int r = 0; r < M; ++r) {
      unsigned long long a1 = 0ull, a2 = 0ull;   // sum w.re * x, sum w.im * x (packed pairs)
#pragma unroll
      for (int c = 0; c < M; c += 2) {
        const float4 w2 = *reinterpret_cast<const float4*>(W + r * MP + c);   // two weights, 16-byte aligned (MP even)
        fma_bc(a1, w2.x, x[c]);
        fma_bc(a2, w2.y, x[c]);
        if (c + 1 < M) {
          fma_bc(a1, w2.z, x[c + 1]);
          fma_bc(a2, w2.w, x[c + 1]);
        }
      }
      float r1, i1, r2, i2;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(r1), "=f"(i1) : "l"(a1));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(r2), "=f"(i2) : "l"(a2));
      cxs<float> y; y.x = r1 - i2; y.y = i1 + r2;
      p[r * step] = y;
    }
  }
};

template <int M> struct SectorOp<double, M> {
  static __device__ __forceinline__ void run(cxs<double>* p, long long step, const cxs<double>* W) {
    constexpr int MP = (M + 1) & ~1;
    double2 x[M];
#pragma unroll
    for (int t = 0; t < M; ++t) x[t] = *reinterpret_cast<const double2*>(p + t * step);
#pragma unroll
    for (int r = 0; r < M; ++r) {
      double yr = 0.0, yi = 0.0;
#pragma unroll
      for (int c = 0; c < M; ++c) {
        const double2 w = *reinterpret_cast<const double2*>(W + r * MP + c);
        yr = fma(w.x, x[c].x, yr); yr = fma(-w.y, x[c].y, yr);
        yi = fma(w.x, x[c].y, yi); yi = fma(w.y, x[c].x, yi);
      }
      *reinterpret_cast<double2*>(p + r * step) = make_double2(yr, yi);
    }
  }
};

template <typename Real, int MAXM, int M>
__device__ __forceinline__ void run_if(cxs<Real>* p, long long step, const cxs<Real>* W) {
  if constexpr (M <= MAXM) SectorOp<Real, M>::run(p, step, W);
}

template <typename Real, int MAXM>
__device__ __forceinline__ void run_sector(int m, cxs<Real>* p, long long step, const cxs<Real>* w) {
  switch (m) {
    case 1: run_if<Real, MAXM, 1>(p, step, w); break;
    case 2: run_if<Real, MAXM, 2>(p, step, w); break;
    case 3: run_if<Real, MAXM, 3>(p, step, w); break;
    case 4: run_if<Real, MAXM, 4>(p, step, w); break;
    case 5: run_if<Real, MAXM, 5>(p, step, w); break;
    case 6: run_if<Real, MAXM, 6>(p, step, w); break;
    case 7: run_if<Real, MAXM, 7>(p, step, w); break;
    case 8: run_if<Real, MAXM, 8>(p, step, w); break;
    case 9: run_if<Real, MAXM, 9>(p, step, w); break;
    case 10: run_if<Real, MAXM, 10>(p, step, w); break;
    case 11: run_if<Real, MAXM, 11>(p, step, w); break;
    case 12: run_if<Real, MAXM, 12>(p, step, w); break;
    case 13: run_if<Real, MAXM, 13>(p, step, w); break;
    case 14: run_if<Real, MAXM, 14>(p, step, w); break;
    case 15: run_if<Real, MAXM, 15>(p, step, w); break;
    case 16: run_if<Real, MAXM, 16>(p, step, w); break;
    default: break;
  }
}

template <typename Real, int MAXM>
__global__ void __launch_bounds__(kSecThreads)
qudit_sector_kernel(cxs<Real>* __restrict__ state, const __grid_constant__ SectorTab T, const QuditGeom g,
                    const cxs<Real>* __restrict__ wpacked) {
  extern __shared__ __align__(16) unsigned char sec_smem[];
  cxs<Real>* W = reinterpret_cast<cxs<Real>*>(sec_smem);
  for (int e = threadIdx.x; e < T.total_w; e += kSecThreads) W[e] = wpacked[e];
  __syncthreads();
  const long long f = (long long)blockIdx.x * kSecThreads + threadIdx.x;   // fibre: lanes = consecutive fibres
  if (f >= g.n_rest) return;
  cxs<Real>* base = state + (long long)blockIdx.y * g.state_size + expand_rest(g, f);
  for (int s = 0; s < T.n_sectors; ++s) {
    cxs<Real>* p = base + (long long)T.s[s].i0 * T.S0 + (long long)T.s[s].j0 * T.S1;
    const cxs<Real>* w = W + T.s[s].woff;
    run_sector<Real, MAXM>(T.s[s].m, p, T.step, w);   // uniform over the grid: every thread is in the same sector
  }
}

// Staged variant for gates that touch the LOWEST mode (stride 1): there consecutive fibres are d (or d^2) amplitudes
// apart and a warp's direct accesses would spread over 32 lines each.  A CTA copies F consecutive fibres (runs of d
// contiguous amplitudes; one contiguous block when both targets are the last two modes) into shared memory, its threads
// -- (fibre, sector class q of Q) -- run the same register blocks on the shared copy, and the block is written back.
struct StagedCfg { int32_t F, Q, pitch; };

template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(sa), "l"(gmem_src), "n"(BYTES) : "memory");
}

template <typename Real, int MAXM>
__global__ void __launch_bounds__(kSecThreads)
qudit_sector_staged_kernel(cxs<Real>* __restrict__ state, const __grid_constant__ SectorTab T, const QuditGeom g,
                           const cxs<Real>* __restrict__ wpacked, const StagedCfg cfg) {
  extern __shared__ __align__(16) unsigned char sec_smem[];
  const int D = g.D, F = cfg.F, pitch = cfg.pitch;
  cxs<Real>* W = reinterpret_cast<cxs<Real>*>(sec_smem);
  cxs<Real>* tile = W + ((T.total_w + 1) & ~1);
  long long* fb = reinterpret_cast<long long*>(tile + size_t(F) * pitch);
  long long* toff = fb + F;
  for (int e = threadIdx.x; e < T.total_w; e += kSecThreads) W[e] = wpacked[e];
  const long long f0 = (long long)blockIdx.x * F;
  for (int i = threadIdx.x; i < F; i += kSecThreads) fb[i] = (f0 + i < g.n_rest) ? expand_rest(g, f0 + i) : -1;
  for (int r = threadIdx.x; r < D; r += kSecThreads)
    toff[r] = T.k == 2 ? (long long)(r / T.d) * T.S0 + (long long)(r % T.d) * T.S1 : (long long)r * T.S0;
  __syncthreads();
  cxs<Real>* st = state + (long long)blockIdx.y * g.state_size;
  // copy in: a warp per fibre, lanes over its D members (runs of d contiguous amplitudes), global -> shared with
  // cp.async: all ~F * D / 256 copies of a thread are in flight at once and no register is staged (with register
  // staging the phase was latency-bound: 51 % of the stall samples sat on the shared store waiting for its load)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int f = warp; f < F; f += kSecThreads / 32) {
    const long long b = fb[f];
    if (b < 0) continue;
    for (int r = lane; r < D; r += 32) cp_async_elem<sizeof(cxs<Real>)>(tile + f * pitch + r, st + b + toff[r]);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  {
    const int fibre = threadIdx.x % F, q = threadIdx.x / F;
    if (q < cfg.Q && fb[fibre] >= 0) {
      const int sstep = T.k == 2 ? T.d + T.dj : 1;   // member step inside a staged fibre (row-major i, j)
      for (int s = q; s < T.n_sectors; s += cfg.Q) {
        cxs<Real>* p = tile + fibre * pitch + (T.k == 2 ? T.s[s].i0 * T.d + T.s[s].j0 : T.s[s].i0);
        run_sector<Real, MAXM>(T.s[s].m, p, sstep, W + T.s[s].woff);
      }
    }
  }
  __syncthreads();
  for (int f = warp; f < F; f += kSecThreads / 32) {
    const long long b = fb[f];
    if (b < 0) continue;
#pragma unroll 4
    for (int r = lane; r < D; r += 32) st[b + toff[r]] = tile[f * pitch + r];
  }
}

// DIAG: state[idx] *= M[r][r], r from the digits of idx
template <typename Real>
__global__ void __launch_bounds__(256)
qudit_diag_kernel(cxs<Real>* __restrict__ state, const cxs<Real>* __restrict__ m, int d, int k, unsigned long long S0,
                  unsigned long long S1, unsigned long long state_size) {
  __shared__ cxs<Real> dg[256];
  const int D = k == 2 ? d * d : d;
  for (int r = threadIdx.x; r < D; r += blockDim.x) dg[r] = m[r * D + r];
  __syncthreads();
  cxs<Real>* st = state + (unsigned long long)blockIdx.y * state_size;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const bool small = state_size < (1ull << 32);
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < state_size; idx += stride) {
    int r;
    if (small) {   // 32-bit divisions
      const unsigned ii = (unsigned)idx;
      const unsigned i = (ii / (unsigned)S0) % (unsigned)d;
      r = k == 2 ? int(i * d + (ii / (unsigned)S1) % (unsigned)d) : int(i);
    } else {
      const unsigned long long i = (idx / S0) % (unsigned long long)d;
      r = k == 2 ? int(i * d + (idx / S1) % (unsigned long long)d) : int(i);
    }
    const cxs<Real> w = dg[r], v = st[idx];
    cxs<Real> y; y.x = w.x * v.x - w.y * v.y; y.y = w.x * v.y + w.y * v.x;
    st[idx] = y;
  }
}

void* g_wpacked[64] = {nullptr};

bool group_fold_enabled() {      // B200Q_FOCK_FOLD=0: diagonal neighbours go through the staged group kernel (A/B)
  const char* e = getenv("B200Q_FOCK_FOLD");
  return !(e && atoi(e) == 0);
}

bool sector_staged_enabled() {   // B200Q_FOCK_STAGED=0: direct register kernel on every mode (A/B measurements)
  const char* e = getenv("B200Q_FOCK_STAGED");
  return !(e && atoi(e) == 0);
}

template <typename Real>
int launch_sectors(void* state, const QuditGeom& g, const SectorTab& T, int64_t batch, cudaStream_t s, int dev);

template <typename Real>
int run_sectors(void* state, const QuditGeom& g, const SectorTab& T, const void* matrix, int64_t batch, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return set_err(B200Q_EINVAL, "bad device");
  if (!g_wpacked[dev]) {
    const int rc = cuda_err(cudaMalloc(&g_wpacked[dev], size_t(kMaxSectors) * kMaxSecD * kMaxSecD * 16), "cudaMalloc(sector blocks)");
    if (rc) return rc;
  }
  pack_blocks_kernel<Real><<<4, 256, 0, s>>>((const cxs<Real>*)matrix, T, (cxs<Real>*)g_wpacked[dev]);
  return launch_sectors<Real>(state, g, T, batch, s, dev);
}

// the packed blocks of T are in g_wpacked[dev]: direct register kernel, or the staged variant on the lowest mode
template <typename Real>
int launch_sectors(void* state, const QuditGeom& g, const SectorTab& T, int64_t batch, cudaStream_t s, int dev) {
  if (g.low_stride == 1 && g.k == 2 && sector_staged_enabled()) {   // (one-mode gates: the direct kernel is faster)
    StagedCfg cfg;
    cfg.pitch = g.D | 1;
    int F = 256;
    while (F > 8 && size_t(F) * cfg.pitch * sizeof(cxs<Real>) > 56 * 1024) F >>= 1;
    if (const char* e = getenv("B200Q_FOCK_STAGED_F")) {   // A/B: fibres per CTA
      const int v = atoi(e);
      if (v >= 8 && v <= 256 && (v & (v - 1)) == 0 && size_t(v) * cfg.pitch * sizeof(cxs<Real>) <= 160 * 1024) F = v;
    }
    cfg.F = F;
    cfg.Q = kSecThreads / F < T.n_sectors ? kSecThreads / F : T.n_sectors;
    if (cfg.Q < 1) cfg.Q = 1;
    const size_t smem = size_t((T.total_w + 1) & ~1) * sizeof(cxs<Real>) + size_t(F) * cfg.pitch * sizeof(cxs<Real>) +
                        size_t(F + g.D) * sizeof(long long);
    const long long nblocks = (g.n_rest + F - 1) / F;
    if (nblocks > 0x7fffffffLL) return set_err(B200Q_EUNSUPPORTED, "state too large for one launch");
    dim3 grid((unsigned)nblocks, (unsigned)batch);
#define B200Q_STAGED_LAUNCH(MAXM)                                                                                    \
  do {                                                                                                               \
    auto kern = qudit_sector_staged_kernel<Real, MAXM>;                                                              \
    const int rc = cuda_err(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),      \
                            "cudaFuncSetAttribute(sector staged)");                                                  \
    if (rc) return rc;                                                                                               \
    kern<<<grid, kSecThreads, smem, s>>>((cxs<Real>*)state, T, g, (const cxs<Real>*)g_wpacked[dev], cfg);            \
  } while (0)
    if (T.d <= 4) B200Q_STAGED_LAUNCH(4);
    else if (T.d <= 8) B200Q_STAGED_LAUNCH(8);
    else if (T.d <= 10) B200Q_STAGED_LAUNCH(10);
    else B200Q_STAGED_LAUNCH(16);
#undef B200Q_STAGED_LAUNCH
    return cuda_err(cudaGetLastError(), "qudit staged sector kernel launch");
  }
  const size_t smem = size_t(T.total_w) * sizeof(cxs<Real>);
  const long long nblocks = (g.n_rest + kSecThreads - 1) / kSecThreads;
  if (nblocks > 0x7fffffffLL) return set_err(B200Q_EUNSUPPORTED, "state too large for one launch");
  dim3 grid((unsigned)nblocks, (unsigned)batch);
#define B200Q_SECTOR_LAUNCH(MAXM)                                                                                    \
  do {                                                                                                               \
    auto kern = qudit_sector_kernel<Real, MAXM>;                                                                     \
    if (smem > 48 * 1024) {                                                                                          \
      const int rc = cuda_err(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),    \
                              "cudaFuncSetAttribute(sector)");                                                       \
      if (rc) return rc;                                                                                             \
    }                                                                                                                \
    kern<<<grid, kSecThreads, smem, s>>>((cxs<Real>*)state, T, g, (const cxs<Real>*)g_wpacked[dev]);                 \
  } while (0)
  if (T.d <= 4) B200Q_SECTOR_LAUNCH(4);
  else if (T.d <= 8) B200Q_SECTOR_LAUNCH(8);
  else if (T.d <= 10) B200Q_SECTOR_LAUNCH(10);
  else B200Q_SECTOR_LAUNCH(16);
#undef B200Q_SECTOR_LAUNCH
  return cuda_err(cudaGetLastError(), "qudit sector kernel launch");
}

}  // namespace

extern "C" int b200q_qudit_apply_structured(void* state, int n_modes, int d, int dtype, const void* matrix,
                                            const int32_t* modes, int n_targets, int structure, int64_t batch,
                                            void* stream) {
  if (structure == B200Q_QUDIT_GENERAL)
    return b200q_qudit_apply(state, n_modes, d, dtype, matrix, modes, n_targets, batch, stream);
  if (!state || !matrix || !modes) return set_err(B200Q_EINVAL, "null argument");
  if (dtype != B200Q_C64 && dtype != B200Q_C128) return set_err(B200Q_EINVAL, "bad dtype");
  if (batch < 1 || batch > 65535) return set_err(B200Q_EINVAL, "bad batch");
  if (structure < B200Q_QUDIT_GENERAL || structure > B200Q_QUDIT_DIFFERENCE) return set_err(B200Q_EINVAL, "bad structure");
  QuditGeom g;
  const char* err = "";
  const int rc = qudit_make_geom(n_modes, d, modes, n_targets, dtype == B200Q_C64 ? 8 : 16, &g, &err);
  if (rc) return set_err(rc == -2 ? B200Q_EUNSUPPORTED : B200Q_EINVAL, err);
  // strides of the matrix digits: i (most significant) acts on modes[0], j on modes[1] (reference order)
  const int64_t S0 = n_targets == 2 ? g.stride[1] : g.stride[0], S1 = n_targets == 2 ? g.stride[0] : 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (structure == B200Q_QUDIT_DIAG) {
    int dev = 0;
    cudaGetDevice(&dev);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long want = ((unsigned long long)g.state_size + 255) / 256;
    const unsigned gx = (unsigned)(want < (unsigned long long)sms * 16 ? (want ? want : 1) : (unsigned long long)sms * 16);
    dim3 grid(gx, (unsigned)batch);
    if (dtype == B200Q_C64)
      qudit_diag_kernel<float><<<grid, 256, 0, s>>>((cxs<float>*)state, (const cxs<float>*)matrix, d, n_targets,
                                                    (unsigned long long)S0, (unsigned long long)(S1 ? S1 : 1),
                                                    (unsigned long long)g.state_size);
    else
      qudit_diag_kernel<double><<<grid, 256, 0, s>>>((cxs<double>*)state, (const cxs<double>*)matrix, d, n_targets,
                                                     (unsigned long long)S0, (unsigned long long)(S1 ? S1 : 1),
                                                     (unsigned long long)g.state_size);
    return cuda_err(cudaGetLastError(), "qudit diag kernel launch");
  }
  SectorTab T;
  if (!make_sectors(structure, d, n_targets, S0, S1, &T))   // cutoff > 16 or structure / arity mismatch: generic kernel
    return b200q_qudit_apply(state, n_modes, d, dtype, matrix, modes, n_targets, batch, stream);
  if (dtype == B200Q_C64) return run_sectors<float>(state, g, T, matrix, batch, s);
  return run_sectors<double>(state, g, T, matrix, batch, s);
}

// ---- Fock transformation matrices on the device --------------------------------------------------------------------
// The reference evaluates the Fock matrix of every gate with Python-level recurrences over the cutoff
// (photonic/gate.py:347-374 beamsplitter, 1091-1114 squeezer; arXiv:2004.11002 Eq. 74-75 and 51-52).  Batched over the
// gates of a class that is still ~400 small launches per forward of config 5 (3.3 ms of launch latency against 11 ms of
// gate passes); here ONE launch per gate class: a CTA per gate, the recurrence in shared memory, complex128 arithmetic
// in the reference's order of operations, result cast to the state's dtype.
namespace {

struct zc { double x, y; };
__device__ __forceinline__ zc zmul(zc a, zc b) { zc r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
__device__ __forceinline__ zc zscale(double s, zc a) { zc r; r.x = s * a.x; r.y = s * a.y; return r; }
__device__ __forceinline__ zc zadd(zc a, zc b) { zc r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }

template <typename Real>
__device__ __forceinline__ void zstore(void* out, long long i, zc v) {
  cxs<Real> o; o.x = Real(v.x); o.y = Real(v.y);
  reinterpret_cast<cxs<Real>*>(out)[i] = o;
}

// T[m, n, p, q] of the mode-mixing matrix u (row-major 2 x 2): q = 0 closed form, then the q-recurrence
template <typename Real>
__global__ void __launch_bounds__(256) fock_bs_matrix_kernel(const zc* __restrict__ u, int d, void* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char bsm[];
  const int d3 = d * d * d;
  zc* prev = reinterpret_cast<zc*>(bsm);
  zc* cur = prev + d3;
  zc* pw0 = cur + d3;          // u00^k
  zc* pw1 = pw0 + d;           // u10^k
  const zc u00 = u[blockIdx.x * 4 + 0], u01 = u[blockIdx.x * 4 + 1], u10 = u[blockIdx.x * 4 + 2], u11 = u[blockIdx.x * 4 + 3];
  if (threadIdx.x == 0) {      // cumulative products, like torch.cumprod (photonic._int_powers)
    zc a; a.x = 1.0; a.y = 0.0;
    zc b = a;
    for (int k = 0; k < d; ++k) {
      pw0[k] = a; pw1[k] = b;
      a = zmul(a, u00); b = zmul(b, u10);
    }
  }
  __syncthreads();
  const long long base = (long long)blockIdx.x * d3 * d;
  for (int e = threadIdx.x; e < d3; e += blockDim.x) {
    const int m = e / (d * d), n = (e / d) % d, p = e % d;
    zc v; v.x = v.y = 0.0;
    if (p == m + n) {
      const double coef = exp(0.5 * (lgamma(double(p) + 1.0) - lgamma(double(m) + 1.0) - lgamma(double(n) + 1.0)));
      v = zmul(zscale(coef, pw0[m]), pw1[n]);
    }
    prev[e] = v;
    zstore<Real>(out, base + (long long)e * d, v);
  }
  __syncthreads();
  for (int q = 1; q < d; ++q) {
    const double sq = sqrt(double(q));
    for (int e = threadIdx.x; e < d3; e += blockDim.x) {
      const int m = e / (d * d), n = (e / d) % d, p = e % d;
      zc v; v.x = v.y = 0.0;
      if (m + n - p == q) {
        zc a; a.x = a.y = 0.0;
        zc b = a;
        if (m > 0) a = zmul(zscale(sqrt(double(m)) / sq, u01), prev[e - d * d]);
        if (n > 0) b = zmul(zscale(sqrt(double(n)) / sq, u11), prev[e - d]);
        v = zadd(a, b);
      }
      cur[e] = v;
      zstore<Real>(out, base + (long long)e * d + q, v);
    }
    __syncthreads();
    zc* t = prev; prev = cur; cur = t;
  }
}

// S(r, theta): column 0 by the rank-1 recurrence over even rows, column n + 1 from columns n and n - 1
template <typename Real>
__global__ void __launch_bounds__(64) fock_squeezing_matrix_kernel(const double* __restrict__ prm, int d,
                                                                    void* __restrict__ out) {
  __shared__ zc cols[3][64];
  const double r = prm[blockIdx.x * 2], theta = prm[blockIdx.x * 2 + 1];
  const double sech = 1.0 / cosh(r), th = tanh(r);
  zc ep; ep.x = cos(theta) * th; ep.y = sin(theta) * th;        // e^{i theta} tanh r
  zc em; em.x = cos(theta) * th; em.y = -sin(theta) * th;       // e^{-i theta} tanh r
  const int m = threadIdx.x;
  const long long base = (long long)blockIdx.x * d * d;
  if (m == 0) {
    zc c; c.x = sqrt(sech); c.y = 0.0;
    cols[0][0] = c;
    for (int k = 1; k < d; ++k) {
      zc v; v.x = v.y = 0.0;
      if (k % 2 == 0) v = zmul(zscale(-sqrt(double(k - 1)) / sqrt(double(k)), ep), cols[0][k - 2]);
      cols[0][k] = v;
    }
  }
  __syncthreads();
  if (m < d) zstore<Real>(out, base + (long long)m * d, cols[0][m]);
  for (int n = 0; n + 1 < d; ++n) {
    const zc* pn = cols[n % 3];
    const zc* pn1 = cols[(n + 2) % 3];     // column n - 1
    zc* nx = cols[(n + 1) % 3];
    if (m < d) {
      zc v; v.x = v.y = 0.0;
      if ((m + n) % 2 == 1) {
        zc sh; sh.x = sh.y = 0.0;
        if (m > 0) sh = pn[m - 1];
        v = zscale(sqrt(double(m)) / sqrt(double(n + 1)) * sech, sh);
        if (n >= 1) v = zadd(v, zmul(zscale(sqrt(double(n)) / sqrt(double(n + 1)), em), pn1[m]));
      }
      nx[m] = v;
      zstore<Real>(out, base + (long long)m * d + n + 1, v);
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int b200q_fock_bs_matrix(const void* mixing, int n_gates, int d, int dtype, void* out, void* stream) {
  if (!mixing || !out) return set_err(B200Q_EINVAL, "null argument");
  if (dtype != B200Q_C64 && dtype != B200Q_C128) return set_err(B200Q_EINVAL, "bad dtype");
  if (n_gates < 1 || d < 1 || d > kMaxSecD) return set_err(B200Q_EUNSUPPORTED, "cutoff above 16: build the matrix in torch");
  const size_t smem = (size_t(2) * d * d * d + 2 * d) * sizeof(zc);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == B200Q_C64) {
    auto kern = fock_bs_matrix_kernel<float>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<n_gates, 256, smem, s>>>((const zc*)mixing, d, out);
  } else {
    auto kern = fock_bs_matrix_kernel<double>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<n_gates, 256, smem, s>>>((const zc*)mixing, d, out);
  }
  return cuda_err(cudaGetLastError(), "fock_bs_matrix launch");
}

extern "C" int b200q_fock_squeezing_matrix(const double* r_theta, int n_gates, int d, int dtype, void* out, void* stream) {
  if (!r_theta || !out) return set_err(B200Q_EINVAL, "null argument");
  if (dtype != B200Q_C64 && dtype != B200Q_C128) return set_err(B200Q_EINVAL, "bad dtype");
  if (n_gates < 1 || d < 1 || d > 64) return set_err(B200Q_EUNSUPPORTED, "cutoff above 64: build the matrix in torch");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == B200Q_C64) fock_squeezing_matrix_kernel<float><<<n_gates, 64, 0, s>>>(r_theta, d, out);
  else fock_squeezing_matrix_kernel<double><<<n_gates, 64, 0, s>>>(r_theta, d, out);
  return cuda_err(cudaGetLastError(), "fock_squeezing_matrix launch");
}

// ---- groups of structured gates on one pair of modes: ONE pass over the state ------------------------------------------
// A two-mode gate together with the one-mode gates that directly precede / follow it on its two modes (the squeezers in
// front of the first beamsplitter layer of config 5, the phase shifter + beamsplitter pairs of an MZI mesh) runs on
// the staged copy of the fibres of its mode pair: copy F fibres into shared memory, apply the gates one after the other
// -- block-structured gates by the register blocks of run_sector (a one-mode gate is d blocks of d members along a row
// or a column of the d x d fibre), diagonal gates elementwise --, write the fibres back.
namespace {

constexpr int kGroupMaxOps = 4;
struct GroupOp {
  int32_t kind;        // 0: sector blocks, 1: diagonal (a D-entry table)
  int32_t n_sectors;
  uint32_t wbase;      // offset of the op's packed weights in the shared table
  struct Sec { uint8_t soff, m; int8_t sstep; uint8_t pad; uint16_t woff; } s[kMaxSectors];
};
struct GroupProg {
  int32_t n_ops, total_w, d, D, F, Q, pitch, lanes_over_fibres;
  int64_t S0, S1;      // strides of the tile digits (row, column)
  GroupOp op[kGroupMaxOps];
};

// D-entry table of a diagonal gate over the (row, column) digits of the fibre: where = 0 row mode, 1 column mode,
// 2 both (matrix index row * d + col), 3 both, reversed (col * d + row)
template <typename Real>
__global__ void pack_diag_kernel(const cxs<Real>* __restrict__ m, int d, int where, cxs<Real>* __restrict__ out,
                                 int accumulate = 0) {
  const int D = d * d, Dm = where >= 2 ? D : d;
  for (int r = threadIdx.x; r < D; r += blockDim.x) {
    const int ti = r / d, tj = r % d;
    const int idx = where == 0 ? ti : where == 1 ? tj : where == 2 ? ti * d + tj : tj * d + ti;
    cxs<Real> v = m[idx * Dm + idx];
    if (accumulate) {      // product with the table already there (several diagonal gates in a row)
      const cxs<Real> w = out[r];
      const Real x = v.x * w.x - v.y * w.y;
      v.y = v.x * w.y + v.y * w.x;
      v.x = x;
    }
    out[r] = v;
  }
}

template <typename Real, int MAXM>
__global__ void __launch_bounds__(kSecThreads)
qudit_group_kernel(cxs<Real>* __restrict__ state, const __grid_constant__ GroupProg P, const QuditGeom g,
                   const cxs<Real>* __restrict__ wpacked) {
  extern __shared__ __align__(16) unsigned char sec_smem[];
  const int D = P.D, F = P.F, pitch = P.pitch;
  cxs<Real>* W = reinterpret_cast<cxs<Real>*>(sec_smem);
  cxs<Real>* tile = W + ((P.total_w + 1) & ~1);
  long long* fb = reinterpret_cast<long long*>(tile + size_t(F) * pitch);
  long long* toff = fb + F;
  for (int e = threadIdx.x; e < P.total_w; e += kSecThreads) W[e] = wpacked[e];
  const long long f0 = (long long)blockIdx.x * F;
  for (int i = threadIdx.x; i < F; i += kSecThreads) fb[i] = (f0 + i < g.n_rest) ? expand_rest(g, f0 + i) : -1;
  for (int r = threadIdx.x; r < D; r += kSecThreads) toff[r] = (long long)(r / P.d) * P.S0 + (long long)(r % P.d) * P.S1;
  __syncthreads();
  cxs<Real>* st = state + (long long)blockIdx.y * g.state_size;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (P.lanes_over_fibres) {   // neither mode is the lowest: consecutive FIBRES are contiguous in memory (runs >= d)
    for (int r = warp; r < D; r += kSecThreads / 32) {
      const long long o = toff[r];
      for (int f = lane; f < F; f += 32)
        if (fb[f] >= 0) cp_async_elem<sizeof(cxs<Real>)>(tile + f * pitch + r, st + fb[f] + o);
    }
  } else {                     // a mode is the lowest: the members of a fibre are runs of d contiguous amplitudes
    for (int f = warp; f < F; f += kSecThreads / 32) {
      const long long b = fb[f];
      if (b < 0) continue;
      for (int r = lane; r < D; r += 32) cp_async_elem<sizeof(cxs<Real>)>(tile + f * pitch + r, st + b + toff[r]);
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const int fibre = threadIdx.x % F, q = threadIdx.x / F;
  for (int o = 0; o < P.n_ops; ++o) {
    const GroupOp& op = P.op[o];
    if (op.kind == 0) {
      if (q < P.Q && fb[fibre] >= 0)
        for (int s = q; s < op.n_sectors; s += P.Q)
          run_sector<Real, MAXM>(op.s[s].m, tile + fibre * pitch + op.s[s].soff, (long long)op.s[s].sstep,
                                 W + op.wbase + op.s[s].woff);
    } else {
      const cxs<Real>* dg = W + op.wbase;
      for (int f = warp; f < F; f += kSecThreads / 32)
        for (int r = lane; r < D; r += 32) {
          const cxs<Real> w = dg[r], v = tile[f * pitch + r];
          cxs<Real> y; y.x = w.x * v.x - w.y * v.y; y.y = w.x * v.y + w.y * v.x;
          tile[f * pitch + r] = y;
        }
    }
    __syncthreads();
  }
  if (P.lanes_over_fibres) {
    for (int r = warp; r < D; r += kSecThreads / 32) {
      const long long o = toff[r];
      for (int f = lane; f < F; f += 32)
        if (fb[f] >= 0) st[fb[f] + o] = tile[f * pitch + r];
    }
  } else {
    for (int f = warp; f < F; f += kSecThreads / 32) {
      const long long b = fb[f];
      if (b < 0) continue;
#pragma unroll 4
      for (int r = lane; r < D; r += 32) st[b + toff[r]] = tile[f * pitch + r];
    }
  }
}

template <typename Real>
int run_group(void* state, const QuditGeom& g, GroupProg& P, const b200q_qudit_op_t* ops, int n_ops,
              const int32_t* tile_modes, int64_t batch, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return set_err(B200Q_EINVAL, "bad device");
  if (!g_wpacked[dev]) {
    const int rc = cuda_err(cudaMalloc(&g_wpacked[dev], size_t(kMaxSectors) * kMaxSecD * kMaxSecD * 16), "cudaMalloc(sector blocks)");
    if (rc) return rc;
  }
  const int d = P.d;
  cxs<Real>* wp = (cxs<Real>*)g_wpacked[dev];
  {
    // one block-structured two-mode gate whose neighbours are all DIAGONAL (phase shifter / Kerr + beamsplitter, the pairs
    // of an MZI mesh): the diagonals are folded into the gate's packed blocks, W' = post . W . pre, and the gate runs
    // as if it were alone -- direct register kernel, no staging
    int main_op = -1, n_main = 0;
    for (int o = 0; o < n_ops; ++o)
      if (ops[o].structure != B200Q_QUDIT_DIAG) { main_op = o; ++n_main; }
    if (n_main == 1 && ops[main_op].n_targets == 2 && ops[main_op].modes[0] == tile_modes[0] && group_fold_enabled()) {
      SectorTab T;
      if (!make_sectors(ops[main_op].structure, d, 2, P.S0, P.S1, &T)) return set_err(B200Q_EUNSUPPORTED, "gate structure not supported in a group");
      cxs<Real>* pre = wp + 4096;          // D <= 256 entries each, behind the packed blocks (<= 3 000 elements)
      cxs<Real>* post = pre + 256;
      int n_pre = 0, n_post = 0;
      for (int o = 0; o < n_ops; ++o) {
        if (o == main_op) continue;
        const b200q_qudit_op_t& in = ops[o];
        const bool on_row = in.modes[0] == tile_modes[0];
        const int where = in.n_targets == 1 ? (on_row ? 0 : 1) : (on_row ? 2 : 3);
        int& cnt = o < main_op ? n_pre : n_post;
        pack_diag_kernel<Real><<<1, 256, 0, s>>>((const cxs<Real>*)(uintptr_t)in.matrix, d, where, o < main_op ? pre : post, cnt > 0);
        ++cnt;
      }
      pack_blocks_kernel<Real><<<4, 256, 0, s>>>((const cxs<Real>*)(uintptr_t)ops[main_op].matrix, T, wp,
                                                 n_pre ? pre : nullptr, n_post ? post : nullptr);
      return launch_sectors<Real>(state, g, T, batch, s, dev);
    }
  }
  P.total_w = 0;
  for (int o = 0; o < n_ops; ++o) {
    const b200q_qudit_op_t& in = ops[o];
    GroupOp& op = P.op[o];
    std::memset(&op, 0, sizeof op);
    const bool on_row = in.modes[0] == tile_modes[0];          // first target of the op is the tile's row mode
    const cxs<Real>* mat = (const cxs<Real>*)(uintptr_t)in.matrix;
    op.wbase = (uint32_t)P.total_w;
    if (in.structure == B200Q_QUDIT_DIAG) {
      op.kind = 1;
      const int where = in.n_targets == 1 ? (on_row ? 0 : 1) : (on_row ? 2 : 3);
      pack_diag_kernel<Real><<<1, 256, 0, s>>>(mat, d, where, wp + op.wbase);
      P.total_w += (P.D + 1) & ~1;
      continue;
    }
    SectorTab T;
    if (!make_sectors(in.structure, d, in.n_targets, 1, 1, &T)) return set_err(B200Q_EUNSUPPORTED, "gate structure not supported in a group");
    pack_blocks_kernel<Real><<<4, 256, 0, s>>>(mat, T, wp + op.wbase);
    op.kind = 0;
    if (in.n_targets == 1) {        // d blocks of d members along the row (column) digit, all with the same d x d weights
      op.n_sectors = d;
      for (int t = 0; t < d; ++t) {
        op.s[t].m = (uint8_t)d;
        op.s[t].woff = 0;
        if (on_row) { op.s[t].soff = (uint8_t)t; op.s[t].sstep = (int8_t)d; }
        else { op.s[t].soff = (uint8_t)(t * d); op.s[t].sstep = 1; }
      }
    } else {
      op.n_sectors = T.n_sectors;
      for (int t = 0; t < T.n_sectors; ++t) {
        const int i0 = T.s[t].i0, j0 = T.s[t].j0;   // digits of the op's first / second target
        op.s[t].m = T.s[t].m;
        op.s[t].woff = (uint16_t)T.s[t].woff;
        if (on_row) { op.s[t].soff = (uint8_t)(i0 * d + j0); op.s[t].sstep = (int8_t)(d + T.dj); }
        else { op.s[t].soff = (uint8_t)(j0 * d + i0); op.s[t].sstep = (int8_t)(T.dj * d + 1); }
      }
    }
    P.total_w += (T.total_w + 1) & ~1;
  }
  if (size_t(P.total_w) > size_t(kMaxSectors) * kMaxSecD * kMaxSecD) return set_err(B200Q_EUNSUPPORTED, "group weights too large");
  P.pitch = P.D | 1;
  int F = 256;
  while (F > 8 && size_t(F) * P.pitch * sizeof(cxs<Real>) > 56 * 1024) F >>= 1;
  P.F = F;
  P.Q = kSecThreads / F < 1 ? 1 : kSecThreads / F;
  const size_t smem = size_t((P.total_w + 1) & ~1) * sizeof(cxs<Real>) + size_t(F) * P.pitch * sizeof(cxs<Real>) +
                      size_t(F + P.D) * sizeof(long long);
  if (smem > 200 * 1024) return set_err(B200Q_EUNSUPPORTED, "group does not fit shared memory");
  const long long nblocks = (g.n_rest + F - 1) / F;
  if (nblocks > 0x7fffffffLL) return set_err(B200Q_EUNSUPPORTED, "state too large for one launch");
  dim3 grid((unsigned)nblocks, (unsigned)batch);
#define B200Q_GROUP_LAUNCH(MAXM)                                                                                     \
  do {                                                                                                               \
    auto kern = qudit_group_kernel<Real, MAXM>;                                                                      \
    const int rc = cuda_err(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),      \
                            "cudaFuncSetAttribute(group)");                                                          \
    if (rc) return rc;                                                                                               \
    kern<<<grid, kSecThreads, smem, s>>>((cxs<Real>*)state, P, g, (const cxs<Real>*)wp);                             \
  } while (0)
  if (d <= 4) B200Q_GROUP_LAUNCH(4);
  else if (d <= 8) B200Q_GROUP_LAUNCH(8);
  else if (d <= 10) B200Q_GROUP_LAUNCH(10);
  else B200Q_GROUP_LAUNCH(16);
#undef B200Q_GROUP_LAUNCH
  return cuda_err(cudaGetLastError(), "qudit group kernel launch");
}

}  // namespace

extern "C" int b200q_qudit_apply_group(void* state, int n_modes, int d, int dtype, const int32_t* tile_modes,
                                       const b200q_qudit_op_t* ops, int n_ops, int64_t batch, void* stream) {
  if (!state || !tile_modes || !ops) return set_err(B200Q_EINVAL, "null argument");
  if (dtype != B200Q_C64 && dtype != B200Q_C128) return set_err(B200Q_EINVAL, "bad dtype");
  if (batch < 1 || batch > 65535) return set_err(B200Q_EINVAL, "bad batch");
  if (n_ops < 1 || n_ops > kGroupMaxOps) return set_err(B200Q_EINVAL, "a group holds 1..4 gates");
  if (d > kMaxSecD) return set_err(B200Q_EUNSUPPORTED, "cutoff above 16");
  for (int o = 0; o < n_ops; ++o) {
    const b200q_qudit_op_t& in = ops[o];
    if (!in.matrix || in.n_targets < 1 || in.n_targets > 2) return set_err(B200Q_EINVAL, "bad gate in group");
    if (in.structure < B200Q_QUDIT_DIAG || in.structure > B200Q_QUDIT_DIFFERENCE) return set_err(B200Q_EINVAL, "unstructured gate in group");
    for (int j = 0; j < in.n_targets; ++j)
      if (in.modes[j] != tile_modes[0] && in.modes[j] != tile_modes[1]) return set_err(B200Q_EINVAL, "gate leaves the mode pair of its group");
    if (in.n_targets == 2 && in.modes[0] == in.modes[1]) return set_err(B200Q_EINVAL, "repeated mode");
  }
  QuditGeom g;
  const char* err = "";
  const int rc = qudit_make_geom(n_modes, d, tile_modes, 2, dtype == B200Q_C64 ? 8 : 16, &g, &err);
  if (rc) return set_err(rc == -2 ? B200Q_EUNSUPPORTED : B200Q_EINVAL, err);
  GroupProg P;
  std::memset(&P, 0, sizeof P);
  P.n_ops = n_ops; P.d = d; P.D = d * d;
  P.S0 = g.stride[1]; P.S1 = g.stride[0];
  P.lanes_over_fibres = g.low_stride == 1 ? 0 : 1;
  if (dtype == B200Q_C64) return run_group<float>(state, g, P, ops, n_ops, tile_modes, batch, (cudaStream_t)stream);
  return run_group<double>(state, g, P, ops, n_ops, tile_modes, batch, (cudaStream_t)stream);
}
