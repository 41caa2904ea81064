// Per-thread body of the fused tile kernel, written as __host__ __device__ templates so that the
// exact index logic that runs on sm_100a can also be stepped thread-by-thread on the CPU by the
// test-only emulator (tests/native/hostemu.cpp).  No CUDA-only intrinsic appears in this file.
//
// Replaces: qmath.evolve_state (qmath.py:485-506) and Gate.op_state_control (operation.py:203-219)
// for every gate of a fused group, applied in place.
#pragma once
#include "b200q_program.h"

#if defined(__CUDACC__)
#define B200Q_HD __host__ __device__ __forceinline__
#else
#define B200Q_HD inline
#endif

namespace b200q {

template <typename Real> struct cx { Real x, y; };
struct alignas(16) chunk_f { float x, y, z, w; };  // two complex64 amplitudes (2m, 2m+1)
struct alignas(16) chunk_d { double x, y; };        // one complex128 amplitude

template <typename Real> struct Traits;
template <> struct Traits<float> {
  static constexpr int VS = 1;   // log2(amplitudes per 16-byte chunk)
  static constexpr int RB = 5;   // register slots (amplitude bits held per thread)
  static constexpr int NA = 32;  // amplitudes per thread
  using chunk = chunk_f;
};
template <> struct Traits<double> {
  static constexpr int VS = 0;
  static constexpr int RB = 4;
  static constexpr int NA = 16;
  using chunk = chunk_d;
};

// XOR swizzle of the chunk index inside the shared-memory tile: the low 3 bits (the 8 x 16-byte bank
// groups of a 128-byte wavefront) are xor-ed with every higher 3-bit group, so a quarter-warp whose
// lanes differ in ANY three chunk-index bits with distinct positions mod 3 is conflict free.
B200Q_HD uint32_t swz(uint32_t c) { return c ^ (((c >> 3) ^ (c >> 6) ^ (c >> 9) ^ (c >> 12)) & 7u); }

B200Q_HD void unpack(const chunk_f& v, float* ar, float* ai, int c) {
  ar[2 * c] = v.x; ai[2 * c] = v.y; ar[2 * c + 1] = v.z; ai[2 * c + 1] = v.w;
}
B200Q_HD void unpack(const chunk_d& v, double* ar, double* ai, int c) { ar[c] = v.x; ai[c] = v.y; }
B200Q_HD chunk_f pack(const float* ar, const float* ai, int c, chunk_f*) {
  chunk_f v; v.x = ar[2 * c]; v.y = ai[2 * c]; v.z = ar[2 * c + 1]; v.w = ai[2 * c + 1]; return v;
}
B200Q_HD chunk_d pack(const double* ar, const double* ai, int c, chunk_d*) {
  chunk_d v; v.x = ar[c]; v.y = ai[c]; return v;
}
B200Q_HD chunk_f zero_chunk(chunk_f*) { chunk_f v; v.x = v.y = v.z = v.w = 0.f; return v; }
B200Q_HD chunk_d zero_chunk(chunk_d*) { chunk_d v; v.x = v.y = 0.0; return v; }

// ------------------------------------------------------------------------------------------------
// register ops
// ------------------------------------------------------------------------------------------------
template <typename Real, int NA, int S>
B200Q_HD void op_mat1(Real* ar, Real* ai, const cx<Real>* m, uint32_t creg) {
  const Real m00r = m[0].x, m00i = m[0].y, m01r = m[1].x, m01i = m[1].y;
  const Real m10r = m[2].x, m10i = m[2].y, m11r = m[3].x, m11i = m[3].y;
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    if (i & (1 << S)) continue;
    if ((uint32_t(i) & creg) != creg) continue;
    const int j = i | (1 << S);
    const Real xr = ar[i], xi = ai[i], yr = ar[j], yi = ai[j];
    ar[i] = m00r * xr - m00i * xi + m01r * yr - m01i * yi;
    ai[i] = m00r * xi + m00i * xr + m01r * yi + m01i * yr;
    ar[j] = m10r * xr - m10i * xi + m11r * yr - m11i * yi;
    ai[j] = m10r * xi + m10i * xr + m11r * yi + m11i * yr;
  }
}

template <typename Real, int NA, int S>
B200Q_HD void op_x(Real* ar, Real* ai, uint32_t creg) {
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    if (i & (1 << S)) continue;
    if ((uint32_t(i) & creg) != creg) continue;
    const int j = i | (1 << S);
    const Real tr = ar[i], ti = ai[i];
    ar[i] = ar[j]; ai[i] = ai[j];
    ar[j] = tr; ai[j] = ti;
  }
}

template <typename Real, int NA>
B200Q_HD void op_diag(Real* ar, Real* ai, const cx<Real>* d, uint32_t sel_base, uint32_t r0, uint32_t r1,
                      uint32_t creg) {
  const cx<Real> d0 = d[0], d1 = d[1], d2 = d[2], d3 = d[3];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    if ((uint32_t(i) & creg) != creg) continue;
    const uint32_t idx = sel_base | ((r0 >> i) & 1u) | (((r1 >> i) & 1u) << 1);
    const cx<Real> lo = (idx & 1u) ? d1 : d0;
    const cx<Real> hi = (idx & 1u) ? d3 : d2;
    const cx<Real> p = (idx & 2u) ? hi : lo;
    const Real xr = ar[i], xi = ai[i];
    ar[i] = p.x * xr - p.y * xi;
    ai[i] = p.x * xi + p.y * xr;
  }
}

template <typename Real, int NA, int RB>
B200Q_HD void dispatch_mat1(int slot, Real* ar, Real* ai, const cx<Real>* m, uint32_t creg) {
  switch (slot) {
    case 0: op_mat1<Real, NA, 0>(ar, ai, m, creg); break;
    case 1: op_mat1<Real, NA, 1>(ar, ai, m, creg); break;
    case 2: op_mat1<Real, NA, 2>(ar, ai, m, creg); break;
    case 3: op_mat1<Real, NA, 3>(ar, ai, m, creg); break;
    default:
      if (RB > 4) op_mat1<Real, NA, (RB > 4 ? 4 : 0)>(ar, ai, m, creg);
      break;
  }
}
template <typename Real, int NA, int RB>
B200Q_HD void dispatch_x(int slot, Real* ar, Real* ai, uint32_t creg) {
  switch (slot) {
    case 0: op_x<Real, NA, 0>(ar, ai, creg); break;
    case 1: op_x<Real, NA, 1>(ar, ai, creg); break;
    case 2: op_x<Real, NA, 2>(ar, ai, creg); break;
    case 3: op_x<Real, NA, 3>(ar, ai, creg); break;
    default:
      if (RB > 4) op_x<Real, NA, (RB > 4 ? 4 : 0)>(ar, ai, creg);
      break;
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-level helpers
// ------------------------------------------------------------------------------------------------
// Physical index (in amplitudes) of the tile's first amplitude: the tile number's bits are
// deposited into the non-tile bit positions.
B200Q_HD uint64_t tile_base(const b200q_pass_t& P, uint64_t tile_id) {
  uint64_t base = 0;
  for (int j = 0; j < P.n_nontile; ++j) base |= ((tile_id >> j) & 1ull) << P.nontile_phys[j];
  return base;
}

// Stage the pass's gate matrices into the shared-memory pool.  Thread `tid` handles quads
// tid, tid + nthreads, ...; a quad is 4 consecutive pool elements of one op.
template <typename Real>
B200Q_HD void fill_pool(const b200q_pass_t& P, int tid, int nthreads, cx<Real>* pool, const cx<Real>* mats) {
  const int nquads = P.pool_elems >> 2;
  for (int q = tid; q < nquads; q += nthreads) {
    const int e0 = q << 2;
    int o = 0;
    for (int t = 0; t < P.n_ops; ++t) {
      const int off = P.ops[t].pool_off;
      if (P.ops[t].pool_n != 0 && e0 >= off && e0 < off + P.ops[t].pool_n) o = t;
    }
    const b200q_op_t& op = P.ops[o];
    const bool adj = (op.flags & B200Q_FLAG_ADJOINT) != 0;
    const cx<Real>* src = mats + op.mat_src;
    const int dim = 1 << (op.kind == B200Q_OP_MAT1 ? 1 : int(op.k));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 - op.pool_off + u;
      cx<Real> v;
      if (op.kind == B200Q_OP_DIAG) {
        if (e < dim) { v = src[e * (dim + 1)]; } else { v.x = Real(1); v.y = Real(0); }
      } else {
        const int r = e / dim, c = e % dim;
        v = adj ? src[c * dim + r] : src[r * dim + c];
      }
      if (adj) v.y = -v.y;
      pool[e0 + u] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// one register round of one thread
// ------------------------------------------------------------------------------------------------
template <typename Real>
B200Q_HD void run_round(const b200q_pass_t& P, const b200q_round_t& Rd, int tid, uint64_t cta_base,
                        typename Traits<Real>::chunk* tile, const cx<Real>* pool,
                        typename Traits<Real>::chunk* gstate, uint64_t total_chunks) {
  using Tr = Traits<Real>;
  using chunk = typename Tr::chunk;
  constexpr int VS = Tr::VS, RB = Tr::RB, NA = Tr::NA;
  const int item_bits = int(P.tile_bits) - RB;
  if (tid >= (1 << item_bits)) return;

  uint32_t lb = 0;   // tile-local amplitude index of this thread's item (register bits zero)
  uint64_t pb = cta_base;  // the same as a physical index
  for (int k = 0; k < item_bits; ++k) {
    const uint32_t bit = (uint32_t(tid) >> k) & 1u;
    const int loc = Rd.nonreg_bit[k];
    lb |= bit << loc;
    pb |= uint64_t(bit) << P.tile_phys[loc];
  }

  Real ar[NA], ai[NA];
  uint64_t gst[4];   // chunk stride of each chunk-level register slot in global memory
  uint32_t sst[4];   // swizzled stride of each chunk-level register slot in the tile
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int loc = Rd.slot_bit[s + VS];
    gst[s] = 1ull << (int(P.tile_phys[loc]) - VS);
    sst[s] = swz(1u << (loc - VS));
  }
  const uint64_t gbase = pb >> VS;
  const uint32_t sbase = swz(lb >> VS);

  if (Rd.src_global) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint64_t idx = gbase + ((c & 1) ? gst[0] : 0) + ((c & 2) ? gst[1] : 0) + ((c & 4) ? gst[2] : 0) +
                           ((c & 8) ? gst[3] : 0);
      chunk v = zero_chunk((chunk*)nullptr);
      if (idx < total_chunks) v = gstate[idx];
      unpack(v, ar, ai, c);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint32_t idx = sbase ^ ((c & 1) ? sst[0] : 0) ^ ((c & 2) ? sst[1] : 0) ^ ((c & 4) ? sst[2] : 0) ^
                           ((c & 8) ? sst[3] : 0);
      unpack(tile[idx], ar, ai, c);
    }
  }

  for (int o = Rd.op_begin; o < Rd.op_end; ++o) {
    const b200q_op_t& op = P.ops[o];
    if ((cta_base & op.ctrl_glob) != op.ctrl_glob) continue;
    if ((lb & op.ctrl_loc) != op.ctrl_loc) continue;
    const cx<Real>* m = pool + op.pool_off;
    switch (op.kind) {
      case B200Q_OP_MAT1: dispatch_mat1<Real, NA, RB>(op.slot, ar, ai, m, op.ctrl_reg); break;
      case B200Q_OP_X: dispatch_x<Real, NA, RB>(op.slot, ar, ai, op.ctrl_reg); break;
      case B200Q_OP_DIAG: {
        uint32_t sel = 0;
        if ((cta_base & op.dsel_glob[0]) | uint64_t(lb & op.dsel_loc[0])) sel |= 1u;
        if ((cta_base & op.dsel_glob[1]) | uint64_t(lb & op.dsel_loc[1])) sel |= 2u;
        op_diag<Real, NA>(ar, ai, m, sel, op.dsel_reg[0], op.dsel_reg[1], op.ctrl_reg);
        break;
      }
      default: break;
    }
  }

  if (Rd.dst_global) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint64_t idx = gbase + ((c & 1) ? gst[0] : 0) + ((c & 2) ? gst[1] : 0) + ((c & 4) ? gst[2] : 0) +
                           ((c & 8) ? gst[3] : 0);
      if (idx < total_chunks) gstate[idx] = pack(ar, ai, c, (chunk*)nullptr);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint32_t idx = sbase ^ ((c & 1) ? sst[0] : 0) ^ ((c & 2) ? sst[1] : 0) ^ ((c & 4) ? sst[2] : 0) ^
                           ((c & 8) ? sst[3] : 0);
      tile[idx] = pack(ar, ai, c, (chunk*)nullptr);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dense k-target op applied in place in the shared-memory tile (k = 2..4)
// ------------------------------------------------------------------------------------------------
template <typename Real>
B200Q_HD cx<Real>* tile_amp(typename Traits<Real>::chunk* tile, uint32_t loc) {
  constexpr int VS = Traits<Real>::VS;
  cx<Real>* base = reinterpret_cast<cx<Real>*>(tile);
  return base + ((swz(loc >> VS) << VS) | (loc & ((1u << VS) - 1u)));
}

template <typename Real, int K>
B200Q_HD void run_matk(const b200q_pass_t& P, const b200q_op_t& op, int tid, int nthreads, uint64_t cta_base,
                       typename Traits<Real>::chunk* tile, const cx<Real>* pool) {
  constexpr int D = 1 << K;
  if ((cta_base & op.ctrl_glob) != op.ctrl_glob) return;
  int srt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) srt[j] = op.tk[j];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = a + 1; b < K; ++b)
      if (srt[b] < srt[a]) { const int t = srt[a]; srt[a] = srt[b]; srt[b] = t; }
  const cx<Real>* m = pool + op.pool_off;
  const int ngroups = 1 << (int(P.tile_bits) - K);
  for (int g = tid; g < ngroups; g += nthreads) {
    uint32_t base = uint32_t(g);
#pragma unroll
    for (int j = 0; j < K; ++j) base = ((base >> srt[j]) << (srt[j] + 1)) | (base & ((1u << srt[j]) - 1u));
    if ((base & op.ctrl_loc) != op.ctrl_loc) continue;
    cx<Real> x[D];
    cx<Real>* ptr[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      uint32_t off = base;
#pragma unroll
      for (int j = 0; j < K; ++j) off |= ((uint32_t(i) >> j) & 1u) << op.tk[j];
      ptr[i] = tile_amp<Real>(tile, off);
      x[i] = *ptr[i];
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      Real yr = Real(0), yi = Real(0);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const cx<Real> w = m[r * D + c];
        yr += w.x * x[c].x - w.y * x[c].y;
        yi += w.x * x[c].y + w.y * x[c].x;
      }
      cx<Real> y; y.x = yr; y.y = yi;
      *ptr[r] = y;
    }
  }
}

template <typename Real>
B200Q_HD void run_direct_op(const b200q_pass_t& P, const b200q_op_t& op, int tid, int nthreads, uint64_t cta_base,
                            typename Traits<Real>::chunk* tile, const cx<Real>* pool) {
  switch (op.k) {
    case 2: run_matk<Real, 2>(P, op, tid, nthreads, cta_base, tile, pool); break;
    case 3: run_matk<Real, 3>(P, op, tid, nthreads, cta_base, tile, pool); break;
    case 4: run_matk<Real, 4>(P, op, tid, nthreads, cta_base, tile, pool); break;
    default: break;
  }
}

}  // namespace b200q
