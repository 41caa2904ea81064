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
constexpr int kGPT = 4;            // groups per thread and matrix row: every ELL entry is loaded once for kGPT outputs
constexpr int kMaxUnits = 6;       // (row, group quad) units per thread and gate: d^T <= kGPT * kMaxUnits * kFThreads
constexpr int kMaxOut = kGPT * kMaxUnits;
constexpr int kMaxD2 = 256;        // d^k of a gate

template <typename Real> struct cxf { Real x, y; };

struct FGate {
  int32_t k;            // 1 or 2 targets
  int32_t D;            // d^k
  int32_t NG;           // groups = d^(T-k)
  int32_t lstride[2];   // tile-local stride of matrix digit j (j = 0: most significant matrix digit, reference order)
  int32_t n_ntd;        // non-target tile digits, least significant first, and their (padded) tile-local strides:
  int32_t ntd_stride[8];   //   the group index enumerates them
  int64_t mat_off;      // element offset of the dense matrix
  int32_t ell_off;      // first ELL row of this gate in the workspace (rows of kMaxD2 entries)
  int32_t pad;
};

struct FPass {
  int32_t n_modes, d, T, n_gates, n_rest, tile_size, tile_hi;   // tile_hi = tile_size / d (logical sizes)
  int32_t tile_alloc;         // elements of the PADDED shared-memory tile (see lpad)
  int32_t lstride_tile[8];    // padded tile-local stride of tile digit j: odd in 8-byte words, so that threads walking
                              // any digit hit distinct banks (powers of an even cutoff map to 4-8 banks only)
  int32_t rest_radix[kMaxModes];
  int64_t rest_stride[kMaxModes];
  int64_t tile_stride[8];     // global stride of tile digit j (j = 0 most significant in the tile)
  int64_t state_size, n_tiles;
  uint32_t magic_d;           // ceil(2^32 / d): e / d = __umulhi(e, magic_d) for e < 2^16
  int32_t ell_cap;            // ELL entries of shared memory reserved for the gates of the pass (0: read from global)
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

__device__ __forceinline__ int group_base(const FGate& G, int grp, int d) {
  // grp enumerates the non-target tile digits, least significant first
  int b = 0;
  for (int j = 0; j < G.n_ntd; ++j) { b += (grp % d) * G.ntd_stride[j]; grp /= d; }
  return b;
}

template <typename Real>
__global__ void __launch_bounds__(kFThreads)
qudit_fused_kernel(cxf<Real>* __restrict__ state, const __grid_constant__ FPass P, const cxf<Real>* __restrict__ vals,
                   const int32_t* __restrict__ offs, const int32_t* __restrict__ width) {
  extern __shared__ __align__(16) unsigned char fsm[];
  cxf<Real>* tile = reinterpret_cast<cxf<Real>*>(fsm);                       // [tile_alloc], padded
  int64_t* goff = reinterpret_cast<int64_t*>(tile + P.tile_alloc);           // [tile_hi]: global offset, last digit 0
  uint16_t* lpad = reinterpret_cast<uint16_t*>(goff + P.tile_hi);             // [tile_hi]: padded tile-local offset
  uint16_t* gbase = lpad + ((P.tile_hi + 7) & ~7);                            // per gate: [NG] tile-local group bases
  const int tid = threadIdx.x, d = P.d, T = P.T;
  int ng_total = 0;
  for (int gi = 0; gi < P.n_gates; ++gi) ng_total += P.gates[gi].NG;
  // ELL rows of every gate of the pass, compacted to their common width, staged ONCE per CTA (the inner loop then
  // reads values and offsets with broadcast shared loads instead of dependent global loads)
  cxf<Real>* svals = reinterpret_cast<cxf<Real>*>(gbase + ((ng_total + 7) & ~7));
  uint16_t* soffs = reinterpret_cast<uint16_t*>(svals + P.ell_cap);
  __shared__ int s_ell_base[kMaxFGates + 1];
  __shared__ int s_staged;
  if (tid == 0) {
    int acc = 0;
    for (int gi = 0; gi < P.n_gates; ++gi) { s_ell_base[gi] = acc; acc += P.gates[gi].D * width[gi]; }
    s_ell_base[P.n_gates] = acc;
    s_staged = (P.ell_cap > 0 && acc <= P.ell_cap) ? 1 : 0;
  }
  __syncthreads();
  const bool staged = s_staged != 0;
  if (staged) {
    for (int gi = 0; gi < P.n_gates; ++gi) {
      const FGate& G = P.gates[gi];
      const int w = width[gi];
      for (int i = tid; i < G.D * w; i += kFThreads) {
        const int r = i / w, j = i - r * w;
        svals[s_ell_base[gi] + i] = vals[((int64_t)G.ell_off + r) * kMaxD2 + j];
        soffs[s_ell_base[gi] + i] = (uint16_t)offs[((int64_t)G.ell_off + r) * kMaxD2 + j];
      }
    }
  }
  // tables, once per CTA
  for (int h = tid; h < P.tile_hi; h += kFThreads) {
    int x = h;
    int64_t o = 0;
    int lp = 0;
    for (int j = T - 2; j >= 0; --j) { o += int64_t(x % d) * P.tile_stride[j]; lp += (x % d) * P.lstride_tile[j]; x /= d; }
    goff[h] = o;
    lpad[h] = (uint16_t)lp;
  }
  {
    int gb = 0;
    for (int gi = 0; gi < P.n_gates; ++gi) {
      const FGate& G = P.gates[gi];
      for (int g = tid; g < G.NG; g += kFThreads) gbase[gb + g] = (uint16_t)group_base(G, g, d);
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
      const int h = (int)__umulhi((unsigned)e, P.magic_d), l = e - h * d;
      tile[lpad[h] + l] = st[base + goff[h] + l * last_stride];
    }
    __syncthreads();
    // ---- gates
    int gb = 0;
    for (int gi = 0; gi < P.n_gates; ++gi) {
      const FGate& G = P.gates[gi];
      const int w = width[gi];
      const cxf<Real>* gv = vals + (int64_t)G.ell_off * kMaxD2;
      const int32_t* go = offs + (int64_t)G.ell_off * kMaxD2;
      // unit u = (row, group quad gq): the thread accumulates groups gq, gq + NU, gq + 2 NU, gq + 3 NU of matrix row `row`
      // (consecutive threads -> consecutive groups; every ELL entry feeds kGPT independent accumulator chains)
      const int NU = (G.NG + kGPT - 1) / kGPT;
      const int n_units = G.D * NU;
      Real yr[kMaxUnits][kGPT], yi[kMaxUnits][kGPT];
#pragma unroll
      for (int i = 0; i < kMaxUnits; ++i) {
        const int u = tid + i * kFThreads;
#pragma unroll
        for (int m = 0; m < kGPT; ++m) yr[i][m] = yi[i][m] = Real(0);
        if (u < n_units) {
          const int row = u / NU, gq = u - row * NU;
          const cxf<Real>* x[kGPT];
#pragma unroll
          for (int m = 0; m < kGPT; ++m) x[m] = tile + gbase[gb + min(gq + m * NU, G.NG - 1)];
          if (staged) {
            const cxf<Real>* rv = svals + s_ell_base[gi] + row * w;
            const uint16_t* ro = soffs + s_ell_base[gi] + row * w;
            for (int j = 0; j < w; ++j) {
              const cxf<Real> mm = rv[j];
              const int off = ro[j];
#pragma unroll
              for (int m = 0; m < kGPT; ++m) {
                const cxf<Real> v = x[m][off];
                yr[i][m] = fma(mm.x, v.x, yr[i][m]); yr[i][m] = fma(-mm.y, v.y, yr[i][m]);
                yi[i][m] = fma(mm.x, v.y, yi[i][m]); yi[i][m] = fma(mm.y, v.x, yi[i][m]);
              }
            }
          } else {
            const cxf<Real>* rv = gv + (int64_t)row * kMaxD2;
            const int32_t* ro = go + (int64_t)row * kMaxD2;
            for (int j = 0; j < w; ++j) {
              const cxf<Real> mm = rv[j];
              const int off = ro[j];
#pragma unroll
              for (int m = 0; m < kGPT; ++m) {
                const cxf<Real> v = x[m][off];
                yr[i][m] = fma(mm.x, v.x, yr[i][m]); yr[i][m] = fma(-mm.y, v.y, yr[i][m]);
                yi[i][m] = fma(mm.x, v.y, yi[i][m]); yi[i][m] = fma(mm.y, v.x, yi[i][m]);
              }
            }
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kMaxUnits; ++i) {
        const int u = tid + i * kFThreads;
        if (u < n_units) {
          const int row = u / NU, gq = u - row * NU;
          const int roff = G.k == 1 ? row * G.lstride[0] : (row / d) * G.lstride[0] + (row % d) * G.lstride[1];
#pragma unroll
          for (int m = 0; m < kGPT; ++m) {
            const int grp = gq + m * NU;
            if (grp < G.NG) {
              cxf<Real> y; y.x = yr[i][m]; y.y = yi[i][m];
              tile[gbase[gb + grp] + roff] = y;
            }
          }
        }
      }
      __syncthreads();
      gb += G.NG;
    }
    // ---- store
    for (int e = tid; e < P.tile_size; e += kFThreads) {
      const int h = (int)__umulhi((unsigned)e, P.magic_d), l = e - h * d;
      st[base + goff[h] + l * last_stride] = tile[lpad[h] + l];
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
  const size_t base_smem = size_t(P.tile_alloc) * sizeof(cxf<Real>) + size_t(P.tile_hi) * sizeof(int64_t) +
                           size_t((P.tile_hi + 7) & ~7) * sizeof(uint16_t) + size_t((ng_total + 7) & ~7) * sizeof(uint16_t);
  // whatever is left of ~200 KB holds the ELL rows of the pass (value + 16-bit offset per entry)
  const size_t budget = 200 * 1024;
  FPass Q = P;
  Q.ell_cap = base_smem < budget ? (int32_t)std::min<size_t>((budget - base_smem) / (sizeof(cxf<Real>) + 2), 60000) & ~7 : 0;
  const size_t smem = base_smem + size_t(Q.ell_cap) * (sizeof(cxf<Real>) + 2) + 16;
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
  const int per_sm = 1;
  const int64_t gx = std::min<int64_t>(P.n_tiles, int64_t(sms) * per_sm * 4);
  dim3 grid((unsigned)gx, (unsigned)batch);
  kern<<<grid, kFThreads, smem, s>>>((cxf<Real>*)state, Q, (const cxf<Real>*)w->vals, w->offs, w->width);
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
  P.magic_d = (uint32_t)((0x100000000ull + uint64_t(d) - 1) / uint64_t(d));
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
  // padded tile-local stride of tile digit j: s_(j) = s_(j+1) * d (+ 1 for an even cutoff), odd in elements
  int32_t lst[8];
  {
    const int pad = (d % 2 == 0) ? 1 : 0;
    lst[n_tile - 1] = 1;
    for (int j = n_tile - 2; j >= 0; --j) lst[j] = lst[j + 1] * d + pad;
    P.tile_alloc = ((lst[0] * d + 1) + 1) & ~1;   // even count: the int64 table behind it stays 16-byte aligned
    if (P.tile_alloc > 65535) return set_err(B200Q_EUNSUPPORTED, "tile too large (cutoff^tile modes)");
    for (int j = 0; j < n_tile; ++j) P.lstride_tile[j] = lst[j];
  }
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
    // (with padded strides a segment of several digits is no longer a plain multiple: group_base() below expands
    //  every non-target digit separately from `ntd` / `ntd_stride`)
    G.n_ntd = 0;
    for (int q = n_tile - 1; q >= 0; --q)
      if (q != pos[0] && (g.n_targets == 1 || q != pos[1])) G.ntd_stride[G.n_ntd++] = lst[q];
    (void)lo; (void)hi; (void)pw;
    G.mat_off = g.mat_offset;
    G.ell_off = ell_rows;
    ell_rows += G.D;
  }
  if (dtype == B200Q_C64) return run_fused<float>(state, P, matrices, batch, (cudaStream_t)stream);
  return run_fused<double>(state, P, matrices, batch, (cudaStream_t)stream);
}
