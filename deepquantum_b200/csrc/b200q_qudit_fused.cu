// Fused Fock (qudit) pass: several gates of the photonic tensor path applied to the state in ONE read + write.
//
// Replaces a run of `evolve_state(state, matrix, nmode, wires, qudit = cutoff)` calls (reference
// photonic/operation.py:142-146, driven gate by gate from photonic/circuit.py:405-431) whose modes all lie in a
// set of T *tile modes*: a CTA stages the d^T amplitudes that differ only in the tile modes (80 KB for four modes
// at cutoff 10, complex64) in shared memory, applies every gate of the pass to the tile, and writes it back -- the
// qudit counterpart of the fused qubit tile pass.  The host planner (photonic.plan_fock_passes) always keeps the
// fastest-varying mode in the tile, so every global access is a run of `cutoff` consecutive amplitudes.
//
// Gate matrices stay dense d^k x d^k in device memory (autograd outputs, never read on the host); a tiny kernel
// compacts them to ELL rows whose column entries are already tile-local offsets: a two-mode beamsplitter at cutoff 10
// is 6.7 % dense (photon-number conservation, photonic/gate.py:356-373), a squeezer parity-sparse.
// Work per tile and gate: every thread owns a fixed set of outputs (row, group), accumulates them in registers from
// the shared tile, then -- after a barrier -- overwrites its outputs in place.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/b200q.h"

namespace b200q {
int set_err(int code, const std::string& msg);
int cuda_err(cudaError_t e, const char* what);
}  // namespace b200q
using b200q::cuda_err;
using b200q::set_err;

namespace {

constexpr int kFThreads = 512;
constexpr int kMaxFGates = B200Q_QUDIT_FUSED_MAX_GATES;
constexpr int kMaxModes = 16;
constexpr int kMaxOut = 24;        // outputs per thread and gate: d^T <= kMaxOut * kFThreads
constexpr int kMaxD2 = 256;        // d^k of a gate

template <typename Real> struct cxf { Real x, y; };

struct FGate {
  int32_t k;            // 1 or 2 targets
  int32_t D;            // d^k
  int32_t NG;           // groups = d^(T-k)
  int32_t lstride[2];   // tile-local stride of matrix digit j (j = 0: most significant matrix digit, reference order)
  int32_t gseg[3];      // group index -> tile-local base: radices of the three segments between / around the targets
  int32_t gmul[3];      //                                 and their tile-local multipliers
  int64_t mat_off;      // element offset of the dense matrix
  int32_t ell_off;      // first ELL row of this gate in the workspace (rows of kMaxD2 entries)
  int32_t pad;
};

struct FPass {
  int32_t n_modes, d, T, n_gates, n_rest, tile_size, tile_hi;   // tile_hi = tile_size / d
  int32_t rest_radix[kMaxModes];
  int64_t rest_stride[kMaxModes];
  int64_t tile_stride[8];     // global stride of tile digit j (j = 0 most significant in the tile)
  int64_t state_size, n_tiles;
  FGate gates[kMaxFGates];
};

// ELL workspace: per gate and row up to kMaxD2 entries (value, tile-local offset), plus the row widths
template <typename Real>
__global__ void __launch_bounds__(kMaxD2)
fused_build_ell(const cxf<Real>* __restrict__ mats, const FPass P, cxf<Real>* __restrict__ vals, int32_t* __restrict__ offs,
                int32_t* __restrict__ width, int d) {
  const FGate& G = P.gates[blockIdx.x];
  const int r = threadIdx.x;
  __shared__ int wmax;
  if (threadIdx.x == 0) wmax = 0;
  __syncthreads();
  if (r < G.D) {
    const cxf<Real>* m = mats + G.mat_off + (int64_t)r * G.D;
    cxf<Real>* v = vals + ((int64_t)G.ell_off + r) * kMaxD2;
    int32_t* o = offs + ((int64_t)G.ell_off + r) * kMaxD2;
    int cnt = 0;
    for (int c = 0; c < G.D; ++c) {
      const cxf<Real> w = m[c];
      if (w.x != Real(0) || w.y != Real(0)) {
        v[cnt] = w;
        o[cnt] = G.k == 1 ? c * G.lstride[0] : (c / d) * G.lstride[0] + (c % d) * G.lstride[1];
        ++cnt;
      }
    }
    atomicMax(&wmax, cnt);
  }
  __syncthreads();
  if (threadIdx.x == 0) width[blockIdx.x] = wmax;
  __syncthreads();
  if (r < G.D) {   // pad the row to the common width with zero entries on offset 0
    const int w = wmax;
    const cxf<Real>* m = mats + G.mat_off + (int64_t)r * G.D;
    int cnt = 0;
    for (int c = 0; c < G.D; ++c) cnt += (m[c].x != Real(0) || m[c].y != Real(0)) ? 1 : 0;
    cxf<Real>* v = vals + ((int64_t)G.ell_off + r) * kMaxD2;
    int32_t* o = offs + ((int64_t)G.ell_off + r) * kMaxD2;
    for (int c = cnt; c < w; ++c) { v[c].x = v[c].y = Real(0); o[c] = 0; }
  }
}

__device__ __forceinline__ int group_base(const FGate& G, int grp) {
  // grp enumerates the non-target tile digits, least significant segment first
  const int a = grp % G.gseg[0];
  const int q = grp / G.gseg[0];
  const int b = q % G.gseg[1];
  const int c = q / G.gseg[1];
  return a * G.gmul[0] + b * G.gmul[1] + c * G.gmul[2];
}

template <typename Real>
__global__ void __launch_bounds__(kFThreads)
qudit_fused_kernel(cxf<Real>* __restrict__ state, const __grid_constant__ FPass P, const cxf<Real>* __restrict__ vals,
                   const int32_t* __restrict__ offs, const int32_t* __restrict__ width) {
  extern __shared__ __align__(16) unsigned char fsm[];
  cxf<Real>* tile = reinterpret_cast<cxf<Real>*>(fsm);                       // [tile_size]
  int64_t* goff = reinterpret_cast<int64_t*>(tile + P.tile_size);            // [tile_hi]: global offset, last digit 0
  uint16_t* gbase = reinterpret_cast<uint16_t*>(goff + P.tile_hi);            // per gate: [NG] tile-local group bases
  const int tid = threadIdx.x, d = P.d, T = P.T;
  // tables, once per CTA
  for (int h = tid; h < P.tile_hi; h += kFThreads) {
    int x = h;
    int64_t o = 0;
    for (int j = T - 2; j >= 0; --j) { o += int64_t(x % d) * P.tile_stride[j]; x /= d; }
    goff[h] = o;
  }
  {
    int gb = 0;
    for (int gi = 0; gi < P.n_gates; ++gi) {
      const FGate& G = P.gates[gi];
      for (int g = tid; g < G.NG; g += kFThreads) gbase[gb + g] = (uint16_t)group_base(G, g);
      gb += G.NG;
    }
  }
  __syncthreads();
  const int64_t last_stride = P.tile_stride[T - 1];
  cxf<Real>* st = state + (int64_t)blockIdx.y * P.state_size;
  for (int64_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
    int64_t base = 0;
    {
      int64_t x = t;
      for (int j = 0; j < P.n_rest; ++j) { base += (x % P.rest_radix[j]) * P.rest_stride[j]; x /= P.rest_radix[j]; }
    }
    // ---- load
    for (int e = tid; e < P.tile_size; e += kFThreads) {
      const int h = e / d, l = e - h * d;
      tile[e] = st[base + goff[h] + l * last_stride];
    }
    __syncthreads();
    // ---- gates
    int gb = 0;
    for (int gi = 0; gi < P.n_gates; ++gi) {
      const FGate& G = P.gates[gi];
      const int w = width[gi];
      const cxf<Real>* gv = vals + (int64_t)G.ell_off * kMaxD2;
      const int32_t* go = offs + (int64_t)G.ell_off * kMaxD2;
      Real yr[kMaxOut], yi[kMaxOut];
#pragma unroll
      for (int i = 0; i < kMaxOut; ++i) {
        const int o = tid + i * kFThreads;
        yr[i] = yi[i] = Real(0);
        if (o < P.tile_size) {
          const int row = o / G.NG, grp = o - row * G.NG;
          const cxf<Real>* x = tile + gbase[gb + grp];
          const cxf<Real>* rv = gv + (int64_t)row * kMaxD2;
          const int32_t* ro = go + (int64_t)row * kMaxD2;
          Real ar = Real(0), ai = Real(0);
          for (int j = 0; j < w; ++j) {
            const cxf<Real> m = rv[j], v = x[ro[j]];
            ar = fma(m.x, v.x, ar); ar = fma(-m.y, v.y, ar);
            ai = fma(m.x, v.y, ai); ai = fma(m.y, v.x, ai);
          }
          yr[i] = ar; yi[i] = ai;
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kMaxOut; ++i) {
        const int o = tid + i * kFThreads;
        if (o < P.tile_size) {
          const int row = o / G.NG, grp = o - row * G.NG;
          const int roff = G.k == 1 ? row * G.lstride[0] : (row / d) * G.lstride[0] + (row % d) * G.lstride[1];
          cxf<Real> y; y.x = yr[i]; y.y = yi[i];
          tile[gbase[gb + grp] + roff] = y;
        }
      }
      __syncthreads();
      gb += G.NG;
    }
    // ---- store
    for (int e = tid; e < P.tile_size; e += kFThreads) {
      const int h = e / d, l = e - h * d;
      st[base + goff[h] + l * last_stride] = tile[e];
    }
    __syncthreads();
  }
}

struct FWorkspace { void* vals = nullptr; int32_t* offs = nullptr; int32_t* width = nullptr; };
FWorkspace g_fws[64];

int get_fws(int dev, FWorkspace** out) {
  if (dev < 0 || dev >= 64) return set_err(B200Q_EINVAL, "bad device");
  FWorkspace& w = g_fws[dev];
  if (!w.vals) {
    const size_t rows = size_t(kMaxFGates) * kMaxD2;
    int rc = cuda_err(cudaMalloc(&w.vals, rows * kMaxD2 * 16), "cudaMalloc(fused qudit workspace)");
    if (!rc) rc = cuda_err(cudaMalloc((void**)&w.offs, rows * kMaxD2 * sizeof(int32_t)), "cudaMalloc(fused qudit workspace)");
    if (!rc) rc = cuda_err(cudaMalloc((void**)&w.width, kMaxFGates * sizeof(int32_t)), "cudaMalloc(fused qudit workspace)");
    if (rc) return rc;
  }
  *out = &w;
  return 0;
}

template <typename Real>
int run_fused(void* state, const FPass& P, const void* mats, int64_t batch, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  FWorkspace* w = nullptr;
  int rc = get_fws(dev, &w);
  if (rc) return rc;
  fused_build_ell<Real><<<P.n_gates, kMaxD2, 0, s>>>((const cxf<Real>*)mats, P, (cxf<Real>*)w->vals, w->offs, w->width, P.d);
  int ng_total = 0;
  for (int i = 0; i < P.n_gates; ++i) ng_total += P.gates[i].NG;
  const size_t smem = size_t(P.tile_size) * sizeof(cxf<Real>) + size_t(P.tile_hi) * sizeof(int64_t) +
                      ((size_t(ng_total) * sizeof(uint16_t) + 15) & ~size_t(15));
  auto kern = qudit_fused_kernel<Real>;
  static size_t smem_set[64][2] = {{0}};
  const int ti = sizeof(Real) == 4 ? 0 : 1;
  if (smem > smem_set[dev][ti]) {
    rc = cuda_err(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(fused qudit)");
    if (rc) return rc;
    smem_set[dev][ti] = smem;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = smem <= 110 * 1024 ? 2 : 1;
  const int64_t gx = std::min<int64_t>(P.n_tiles, int64_t(sms) * per_sm * 4);
  dim3 grid((unsigned)gx, (unsigned)batch);
  kern<<<grid, kFThreads, smem, s>>>((cxf<Real>*)state, P, (const cxf<Real>*)w->vals, w->offs, w->width);
  return cuda_err(cudaGetLastError(), "fused qudit kernel launch");
}

}  // namespace

extern "C" int b200q_qudit_fused(void* state, int n_modes, int d, int dtype, const int32_t* tile_modes, int n_tile,
                                 const b200q_qudit_gate_t* gates, int n_gates, const void* matrices, int64_t batch,
                                 void* stream) {
  if (!state || !tile_modes || !gates || !matrices) return set_err(B200Q_EINVAL, "null argument");
  if (dtype != B200Q_C64 && dtype != B200Q_C128) return set_err(B200Q_EINVAL, "bad dtype");
  if (batch < 1 || batch > 65535) return set_err(B200Q_EINVAL, "bad batch");
  if (n_modes < 1 || n_modes > kMaxModes || d < 2) return set_err(B200Q_EINVAL, "bad geometry");
  if (n_tile < 1 || n_tile > 8 || n_tile > n_modes) return set_err(B200Q_EINVAL, "bad tile");
  if (n_gates < 1 || n_gates > kMaxFGates) return set_err(B200Q_EINVAL, "too many gates for one fused pass");
  FPass P;
  std::memset(&P, 0, sizeof P);
  P.n_modes = n_modes; P.d = d; P.T = n_tile; P.n_gates = n_gates;
  int64_t stride[kMaxModes];
  int64_t acc = 1;
  for (int m = n_modes - 1; m >= 0; --m) { stride[m] = acc; acc *= d; if (acc > (int64_t(1) << 40)) return set_err(B200Q_EUNSUPPORTED, "state too large"); }
  P.state_size = acc;
  bool in_tile[kMaxModes] = {false};
  int prev = -1;
  int64_t tsz = 1;
  for (int j = 0; j < n_tile; ++j) {
    const int m = tile_modes[j];
    if (m <= prev || m >= n_modes) return set_err(B200Q_EINVAL, "tile modes must be ascending and in range");
    prev = m;
    in_tile[m] = true;
    P.tile_stride[j] = stride[m];
    tsz *= d;
  }
  if (tsz > int64_t(kMaxOut) * kFThreads || tsz > 65535) return set_err(B200Q_EUNSUPPORTED, "tile too large (cutoff^tile modes)");
  P.tile_size = (int32_t)tsz;
  P.tile_hi = (int32_t)(tsz / d);
  P.n_tiles = 1;
  for (int m = n_modes - 1; m >= 0; --m)
    if (!in_tile[m]) { P.rest_radix[P.n_rest] = d; P.rest_stride[P.n_rest] = stride[m]; ++P.n_rest; P.n_tiles *= d; }
  // tile-local stride of tile digit j
  int32_t lst[8];
  { int32_t a = 1; for (int j = n_tile - 1; j >= 0; --j) { lst[j] = a; a *= d; } }
  int ell_rows = 0;
  for (int gi = 0; gi < n_gates; ++gi) {
    const b200q_qudit_gate_t& g = gates[gi];
    FGate& G = P.gates[gi];
    if (g.n_targets < 1 || g.n_targets > 2) return set_err(B200Q_EUNSUPPORTED, "fused Fock passes take 1- and 2-mode gates");
    int pos[2] = {-1, -1};
    for (int j = 0; j < g.n_targets; ++j) {
      for (int q = 0; q < n_tile; ++q)
        if (tile_modes[q] == g.modes[j]) pos[j] = q;
      if (pos[j] < 0) return set_err(B200Q_EINVAL, "gate mode outside the tile");
    }
    if (g.n_targets == 2 && pos[0] == pos[1]) return set_err(B200Q_EINVAL, "repeated gate mode");
    G.k = g.n_targets;
    G.D = g.n_targets == 1 ? d : d * d;
    if (G.D > kMaxD2) return set_err(B200Q_EUNSUPPORTED, "cutoff^targets above 256");
    G.NG = P.tile_size / G.D;
    for (int j = 0; j < g.n_targets; ++j) G.lstride[j] = lst[pos[j]];
    // segments of non-target tile digits (least significant first): below the lower target, between, above
    int lo = pos[0], hi = pos[0];
    if (g.n_targets == 2) { lo = std::max(pos[0], pos[1]); hi = std::min(pos[0], pos[1]); }   // lo = less significant (larger index)
    auto pw = [&](int e) { int32_t r = 1; for (int i = 0; i < e; ++i) r *= d; return r; };
    G.gseg[0] = pw(n_tile - 1 - lo); G.gmul[0] = 1;
    if (g.n_targets == 2) {
      G.gseg[1] = pw(lo - hi - 1); G.gmul[1] = lst[lo] * d;
      G.gseg[2] = pw(hi); G.gmul[2] = lst[hi] * d;
    } else {
      G.gseg[1] = pw(lo); G.gmul[1] = lst[lo] * d;
      G.gseg[2] = 1; G.gmul[2] = 0;
    }
    G.mat_off = g.mat_offset;
    G.ell_off = ell_rows;
    ell_rows += G.D;
  }
  if (dtype == B200Q_C64) return run_fused<float>(state, P, matrices, batch, (cudaStream_t)stream);
  return run_fused<double>(state, P, matrices, batch, (cudaStream_t)stream);
}
