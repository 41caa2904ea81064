// Qudit (Fock tensor) gate application: evolve_state(state, matrix, nmode, wires, qudit = cutoff) as called by
// the photonic back-end (reference photonic/operation.py:142-146 -> qmath.py:485-506), in place.
//
// Same family as the qubit tile kernel -- stage the strided partner amplitudes of a block of groups in
// shared memory with coalesced global accesses, contract, write back once -- with digit (base-d) index
// arithmetic instead of bit tricks, and with the matrix held in ELL (row-compressed) form: a two-mode
// beamsplitter on cutoff 10 is a 100 x 100 matrix with 6.7 % non-zeros (photon-number conservation,
// photonic/gate.py:356-373) and the squeezer is parity-sparse, so skipping exact zeros turns a
// compute-leaning dense GEMM (800 flop/amplitude) back into a bandwidth-bound sweep.  The ELL table is
// built on the device from the dense matrix (no host read of matrix values).
#include <cuda_runtime.h>

#include <cstdint>
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

constexpr int kMaxD = 256;      // d^k of the gate
constexpr int kThreads = 256;
constexpr int kLoadUnroll = 4;

template <typename Real> struct cxq { Real x, y; };

struct EllHeader { int width; };

// one thread per matrix row: compact the non-zeros
template <typename Real>
__global__ void build_ell_kernel(const cxq<Real>* __restrict__ m, int D, cxq<Real>* __restrict__ vals,
                                 short* __restrict__ cols, EllHeader* hdr) {
  __shared__ int wmax;
  if (threadIdx.x == 0) wmax = 0;
  __syncthreads();
  const int r = threadIdx.x;
  if (r < D) {
    int cnt = 0;
    for (int c = 0; c < D; ++c) {
      const cxq<Real> v = m[r * D + c];
      if (v.x != Real(0) || v.y != Real(0)) {
        vals[r * D + cnt] = v;
        cols[r * D + cnt] = (short)c;
        ++cnt;
      }
    }
    for (int c = cnt; c < D; ++c) { cxq<Real> z; z.x = z.y = Real(0); vals[r * D + c] = z; cols[r * D + c] = 0; }
    atomicMax(&wmax, cnt);
  }
  __syncthreads();
  if (threadIdx.x == 0) hdr->width = wmax;
}

// Mixed-radix walker over the elements e = x0 + R0 * (x1 + R1 * x2) of a CTA's tile, advanced by a fixed step:
// the element -> (group, digit combo) mapping of qudit_elem() without its divisions (two 32-bit divisions by
// run-time values per element and per phase were most of the instructions of the load / store loops).
struct Walker {
  int x0, x1, x2, s0, s1, s2, R0, R1;
  __device__ __forceinline__ void init(int e, int step, int r0, int r1) {
    R0 = r0; R1 = r1;
    x0 = e % r0; int q = e / r0; x1 = q % r1; x2 = q / r1;
    s0 = step % r0; q = step / r0; s1 = q % r1; s2 = q / r1;
  }
  __device__ __forceinline__ void next() {
    x0 += s0;
    const int c0 = x0 >= R0;
    x0 -= c0 ? R0 : 0;
    x1 += s1 + c0;
    const int c1 = x1 >= R1;
    x1 -= c1 ? R1 : 0;
    x2 += s2 + c1;
  }
};
constexpr int kNoCarry = 1 << 28;   // radix of a digit that never carries

// Tile element walker in memory order (same order as qudit_elem): gives (gi, t) of the current element.
struct ElemWalker {
  Walker w;
  int d, swap_digits, lane_over_target;
  __device__ __forceinline__ void init(const QuditGeom& g, int e, int step) {
    d = g.d;
    lane_over_target = g.lane_over_target;
    swap_digits = !(g.k == 1 || g.stride[0] < g.stride[1]);
    if (lane_over_target) w.init(e, step, g.d, g.G);   // lowest target digit fastest, then group, then the other digit
    else w.init(e, step, g.G, kNoCarry);               // group fastest, then digit combo
  }
  __device__ __forceinline__ void get(int* gi, int* t) const {
    if (lane_over_target) { *gi = w.x1; *t = swap_digits ? (w.x2 + d * w.x0) : (w.x0 + d * w.x2); }
    else { *gi = w.x0; *t = w.x1; }
  }
  __device__ __forceinline__ void next() { w.next(); }
};

constexpr int kOut = 2;   // outputs (groups) per thread and matrix row in the contraction (4: twice the bank conflicts, no faster)

// Complex accumulator y += w * x.  float: two packed f32x2 FMAs per product on the pair (x.re, x.im) with the
// broadcast operands (w.re, w.re), (w.im, w.im): acc1 = sum w.re * x, acc2 = sum w.im * x,
// y = (acc1.re - acc2.im, acc1.im + acc2.re) -- half the issue slots of four scalar FMAs, no operand swaps.
template <typename Real> struct CAcc;
template <> struct CAcc<float> {
  unsigned long long a1 = 0ull, a2 = 0ull;
  struct W { unsigned long long re2, im2; };
  static __device__ __forceinline__ W make_w(cxq<float> w) {
    W o;
    asm("mov.b64 %0, {%1, %1};" : "=l"(o.re2) : "f"(w.x));
    asm("mov.b64 %0, {%1, %1};" : "=l"(o.im2) : "f"(w.y));
    return o;
  }
  __device__ __forceinline__ void mac(const W& w, const cxq<float>* x) {
    const unsigned long long xv = *reinterpret_cast<const unsigned long long*>(x);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(w.re2), "l"(xv));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2) : "l"(w.im2), "l"(xv));
  }
  __device__ __forceinline__ cxq<float> result() const {
    float r1, i1, r2, i2;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r1), "=f"(i1) : "l"(a1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r2), "=f"(i2) : "l"(a2));
    cxq<float> y; y.x = r1 - i2; y.y = i1 + r2;
    return y;
  }
};
template <> struct CAcc<double> {
  double yr = 0.0, yi = 0.0;
  typedef cxq<double> W;
  static __device__ __forceinline__ W make_w(cxq<double> w) { return w; }
  __device__ __forceinline__ void mac(const W& w, const cxq<double>* x) {
    const cxq<double> v = *x;
    yr = fma(w.x, v.x, yr); yr = fma(-w.y, v.y, yr);
    yi = fma(w.x, v.y, yi); yi = fma(w.y, v.x, yi);
  }
  __device__ __forceinline__ cxq<double> result() const { cxq<double> y; y.x = yr; y.y = yi; return y; }
};

// Contraction of a CTA's tile.  STAGED: `vals` / `cols` are the shared-memory copies (row stride = width).
template <typename Real, bool STAGED>
__device__ __forceinline__ void contract_rows(const cxq<Real>* xs, cxq<Real>* ys, const cxq<Real>* vals,
                                              const short* cols, int row_stride, int width, int D, int G, int GP) {
  const int Gq = (G + kOut - 1) / kOut;
  Walker cw;
  cw.init(threadIdx.x, kThreads, Gq, kNoCarry);
  for (int o = threadIdx.x; o < Gq * D; o += kThreads, cw.next()) {
    const int gi = cw.x0, r = cw.x1;
    int gm[kOut];
#pragma unroll
    for (int m = 0; m < kOut; ++m) gm[m] = (gi + m * Gq < G) ? gi + m * Gq : gi;   // out of range: recompute gi
    CAcc<Real> acc[kOut];
    const cxq<Real>* vrow = vals + r * row_stride;
    const short* crow = cols + r * row_stride;
#pragma unroll 2
    for (int j = 0; j < width; ++j) {
      const typename CAcc<Real>::W w = CAcc<Real>::make_w(vrow[j]);
      const cxq<Real>* xrow = xs + int(crow[j]) * GP;
#pragma unroll
      for (int m = 0; m < kOut; ++m) acc[m].mac(w, xrow + gm[m]);
    }
#pragma unroll
    for (int m = 0; m < kOut; ++m)
      if (m == 0 || gi + m * Gq < G) ys[r * GP + gm[m]] = acc[m].result();
  }
}

template <typename Real>
__global__ void __launch_bounds__(kThreads)
qudit_apply_kernel(cxq<Real>* __restrict__ state, const QuditGeom g, const cxq<Real>* __restrict__ ell_vals,
                   const short* __restrict__ ell_cols, const EllHeader* __restrict__ hdr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D = g.D, G = g.G, GP = G + 1;
  cxq<Real>* xs = reinterpret_cast<cxq<Real>*>(smem_raw);   // [D][GP]
  cxq<Real>* ys = xs + size_t(D) * GP;                       // [D][GP]
  long long* gbase = reinterpret_cast<long long*>(ys + size_t(D) * GP);   // [G]  flat offset of each group
  long long* toff = gbase + G;                                            // [D]  flat offset of each digit combo
  cxq<Real>* st = state + (long long)blockIdx.y * g.state_size;
  const long long r0 = (long long)blockIdx.x * G;
  const int width = hdr->width;
  const int total = G * D;
  // the 64-bit divisions of the index expansion are done once per group / digit combo, not per element
  for (int i = threadIdx.x; i < G; i += kThreads) gbase[i] = (r0 + i < g.n_rest) ? expand_rest(g, r0 + i) : -1;
  for (int t = threadIdx.x; t < D; t += kThreads) toff[t] = qudit_offset(g, 0, t) - expand_rest(g, 0);
  __syncthreads();
  // ---- load: element e -> (group gi, matrix digit combo t) ordered for coalescing --------------------
  // (kLoadUnroll independent global loads in flight per thread before the first shared store: with one load per
  // iteration the loop was latency-bound -- 26 % of the samples sat on the store waiting for its load -- and a
  // B200 needs ~40 KB in flight per SM to reach its HBM bandwidth)
  {
    ElemWalker ew;
    ew.init(g, threadIdx.x, kThreads);
    for (int e0 = threadIdx.x; e0 < total; e0 += kLoadUnroll * kThreads) {
      cxq<Real> v[kLoadUnroll];
      int dst[kLoadUnroll];
#pragma unroll
      for (int u = 0; u < kLoadUnroll; ++u) {
        v[u].x = v[u].y = Real(0);
        dst[u] = -1;
        if (e0 + u * kThreads < total) {
          int gi, t;
          ew.get(&gi, &t);
          const long long b = gbase[gi];
          if (b >= 0) v[u] = st[b + toff[t]];
          dst[u] = t * GP + gi;
        }
        ew.next();
      }
#pragma unroll
      for (int u = 0; u < kLoadUnroll; ++u)
        if (dst[u] >= 0) xs[dst[u]] = v[u];
    }
  }
  __syncthreads();
  // ---- contract: y[r][gi] = sum_j vals[r][j] * x[cols[r][j]][gi]; kOut groups (gi + m * Gq) per thread and
  // matrix row, so that every ELL entry is loaded once for kOut outputs; ELL rows staged in shared memory when
  // they fit (width <= ell_cap) ---------------------------------------------------------------------------
  {
    const bool staged = g.ell_cap > 0 && width <= g.ell_cap;
    cxq<Real>* svals = reinterpret_cast<cxq<Real>*>(toff + D);
    short* scols = reinterpret_cast<short*>(svals + size_t(D) * g.ell_cap);
    if (staged) {
      for (int i = threadIdx.x; i < D * width; i += kThreads) {
        const int r = i / width, j = i - r * width;
        svals[i] = ell_vals[r * D + j];
        scols[i] = ell_cols[r * D + j];
      }
      __syncthreads();
      contract_rows<Real, true>(xs, ys, svals, scols, width, width, D, G, GP);
    } else {
      contract_rows<Real, false>(xs, ys, ell_vals, ell_cols, D, width, D, G, GP);
    }
  }
  __syncthreads();
  // ---- store (same order as the load) -------------------------------------------------------------------
  {
    ElemWalker ew;
    ew.init(g, threadIdx.x, kThreads);
    for (int e = threadIdx.x; e < total; e += kThreads, ew.next()) {
      int gi, t;
      ew.get(&gi, &t);
      const long long b = gbase[gi];
      if (b >= 0) st[b + toff[t]] = ys[t * GP + gi];
    }
  }
}

struct Workspace { void* vals = nullptr; short* cols = nullptr; EllHeader* hdr = nullptr; };
Workspace g_ws[64];

int get_workspace(int dev, Workspace** out) {
  if (dev < 0 || dev >= 64) return set_err(B200Q_EINVAL, "bad device");
  Workspace& w = g_ws[dev];
  if (!w.vals) {
    int rc = cuda_err(cudaMalloc(&w.vals, size_t(kMaxD) * kMaxD * 16), "cudaMalloc(qudit workspace)");
    if (rc) return rc;
    rc = cuda_err(cudaMalloc((void**)&w.cols, size_t(kMaxD) * kMaxD * sizeof(short)), "cudaMalloc(qudit workspace)");
    if (rc) return rc;
    rc = cuda_err(cudaMalloc((void**)&w.hdr, sizeof(EllHeader)), "cudaMalloc(qudit workspace)");
    if (rc) return rc;
  }
  *out = &w;
  return 0;
}

template <typename Real>
int run_qudit(void* state, const QuditGeom& g, const void* matrix, int64_t batch, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  Workspace* w = nullptr;
  int rc = get_workspace(dev, &w);
  if (rc) return rc;
  build_ell_kernel<Real><<<1, kMaxD, 0, s>>>((const cxq<Real>*)matrix, g.D, (cxq<Real>*)w->vals, w->cols, w->hdr);
  const size_t smem = size_t(2) * g.D * (g.G + 1) * sizeof(cxq<Real>) + size_t(g.G + g.D) * sizeof(long long) +
                      size_t(g.ell_bytes);
  auto kern = qudit_apply_kernel<Real>;
  static size_t smem_set[64] = {0};
  if (smem > smem_set[dev]) {
    rc = cuda_err(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute");
    if (rc) return rc;
    smem_set[dev] = smem;
  }
  const long long nblocks = (g.n_rest + g.G - 1) / g.G;
  if (nblocks > 0x7fffffffLL) return set_err(B200Q_EUNSUPPORTED, "state too large for one launch");
  dim3 grid((unsigned)nblocks, (unsigned)batch);
  kern<<<grid, kThreads, smem, s>>>((cxq<Real>*)state, g, (const cxq<Real>*)w->vals, w->cols, w->hdr);
  return cuda_err(cudaGetLastError(), "qudit kernel launch");
}

}  // namespace

extern "C" int b200q_qudit_apply(void* state, int n_modes, int d, int dtype, const void* matrix, const int32_t* modes,
                                 int n_targets, int64_t batch, void* stream) {
  if (!state || !matrix || !modes) return set_err(B200Q_EINVAL, "null argument");
  if (dtype != B200Q_C64 && dtype != B200Q_C128) return set_err(B200Q_EINVAL, "bad dtype");
  if (batch < 1 || batch > 65535) return set_err(B200Q_EINVAL, "bad batch");
  QuditGeom g;
  const char* err = "";
  const int rc = qudit_make_geom(n_modes, d, modes, n_targets, dtype == B200Q_C64 ? 8 : 16, &g, &err);
  if (rc) return set_err(rc == -2 ? B200Q_EUNSUPPORTED : B200Q_EINVAL, err);
  if (dtype == B200Q_C64) return run_qudit<float>(state, g, matrix, batch, (cudaStream_t)stream);
  return run_qudit<double>(state, g, matrix, batch, (cudaStream_t)stream);
}
