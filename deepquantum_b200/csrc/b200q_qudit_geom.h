// Index geometry of the qudit (Fock tensor) kernel, shared with the test-only CPU emulator.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define B200Q_QHD __host__ __device__ __forceinline__
#else
#define B200Q_QHD inline
#endif

struct QuditGeom {
  int k;                 // targets (1 or 2)
  int d;                 // qudit dimension
  int D;                 // d^k
  int G;                 // groups per CTA
  long long n_rest;      // number of groups = d^(n-k)
  long long stride[2];   // flat stride of target j (j = 0: least significant matrix digit)
  long long low_stride;  // the smaller target stride
  long long seg_inner;   // rest-index block below the lowest target   (= low_stride)
  long long seg_mid;     // rest-index block between the targets       (k = 2)
  long long state_size;  // d^n
  int lane_over_target;  // 1: lanes run over the lowest target digit (its stride is 1)
  int ell_cap;           // ELL rows of at most this many entries are staged in shared memory (0: never)
  int ell_bytes;         // shared memory reserved for them
};

// flat offset of the group `rest` (target digits zero)
B200Q_QHD long long expand_rest(const QuditGeom& g, long long rest) {
  const long long inner = rest % g.seg_inner;
  long long up = rest / g.seg_inner;
  long long off = inner;
  if (g.k == 1) return off + up * g.seg_inner * g.d;
  const long long mid = up % g.seg_mid;
  up /= g.seg_mid;
  return off + mid * g.seg_inner * g.d + up * g.seg_inner * g.d * g.seg_mid * g.d;
}


// element e of a CTA's tile -> (group gi, matrix digit combo t), ordered so that consecutive e are consecutive
// in memory as far as the layout allows
B200Q_QHD void qudit_elem(const QuditGeom& g, int e, int* gi, int* t) {
  if (g.lane_over_target) {      // lowest target digit fastest, then group, then the other digit
    const int t_low = e % g.d;
    const int rest = e / g.d;
    *gi = rest % g.G;
    const int t_hi = rest / g.G;
    *t = (g.k == 1 || g.stride[0] < g.stride[1]) ? (t_low + g.d * t_hi) : (t_hi + g.d * t_low);
  } else {
    *gi = e % g.G;
    *t = e / g.G;
  }
}
B200Q_QHD long long qudit_offset(const QuditGeom& g, long long rest, int t) {
  long long off = expand_rest(g, rest) + (long long)(t % g.d) * g.stride[0];
  if (g.k == 2) off += (long long)(t / g.d) * g.stride[1];
  return off;
}

// Fills the geometry; returns 0 or a negative error code (message in *err).
inline int qudit_make_geom(int n_modes, int d, const int32_t* modes, int n_targets, int elt_bytes, QuditGeom* g,
                           const char** err) {
  if (n_modes < 1 || n_modes > 40 || d < 2 || d > 64) { *err = "bad n_modes / d"; return -1; }
  if (n_targets < 1 || n_targets > 2) { *err = "qudit gates on 1 or 2 modes only"; return -2; }
  g->k = n_targets;
  g->d = d;
  g->D = n_targets == 1 ? d : d * d;
  if (g->D > 256) { *err = "d^k > 256"; return -2; }
  long long size = 1;
  for (int i = 0; i < n_modes; ++i) {
    size *= d;
    if (size > (1LL << 40)) { *err = "state too large"; return -1; }
  }
  g->state_size = size;
  for (int j = 0; j < n_targets; ++j)
    if (modes[j] < 0 || modes[j] >= n_modes) { *err = "mode out of range"; return -1; }
  if (n_targets == 2 && modes[0] == modes[1]) { *err = "repeated mode"; return -1; }
  long long st[2] = {1, 1};
  for (int j = 0; j < n_targets; ++j) {
    long long s = 1;
    for (int i = modes[n_targets - 1 - j] + 1; i < n_modes; ++i) s *= d;
    st[j] = s;   // matrix digit j (0 = least significant) acts on modes[k-1-j] (reference order, qmath.py:497-504)
  }
  g->stride[0] = st[0];
  g->stride[1] = n_targets == 2 ? st[1] : 0;
  const long long lo = n_targets == 2 ? (st[0] < st[1] ? st[0] : st[1]) : st[0];
  const long long hi = n_targets == 2 ? (st[0] < st[1] ? st[1] : st[0]) : 0;
  g->low_stride = lo;
  g->seg_inner = lo;
  g->seg_mid = n_targets == 2 ? hi / (lo * d) : 1;
  g->n_rest = size / g->D;
  g->lane_over_target = lo == 1 ? 1 : 0;
  // photon-number conserving two-mode gates and all one-mode gates have at most d entries per row
  g->ell_cap = d;
  g->ell_bytes = ((g->D * g->ell_cap * (elt_bytes + 2)) + 15) & ~15;
  if (g->ell_bytes > 12 * 1024) { g->ell_cap = 0; g->ell_bytes = 0; }
  // 52 KiB per CTA: four CTAs (plus 1 KiB each of system shared memory) fit the 228 KiB of an SM
  int G = (int)((52 * 1024 - g->ell_bytes) / (2 * g->D * elt_bytes));
  if (G > 256) G = 256;
  if (G < 1) G = 1;
  g->G = G;
  return 0;
}
