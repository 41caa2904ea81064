// Per-thread body of the fused tile kernel, written as __host__ __device__ templates so that the
// exact index logic that runs on sm_100a can also be stepped thread-by-thread on the CPU by the
// test-only emulator (tests/native/hostemu.cpp).
//
// Replaces: qmath.evolve_state (qmath.py:485-506) and Gate.op_state_control (operation.py:203-219)
// for every gate of a fused group, applied in place.
//
// Register model.  A thread holds 16 *elements* (one per 16-byte chunk), indexed by 4 chunk-level
// register slots.  For complex128 an element is one amplitude (re, im doubles).  For complex64 an
// element is the PAIR of amplitudes (2m, 2m+1) that share a chunk, kept as two packed float2
// registers re = (re0, re1), im = (im0, im1): every op on a chunk-level slot is then a 2-wide SIMD
// op issued as Blackwell packed-FP32 instructions (FFMA2 / FMUL2: `__ffma2_rn`), half the
// instruction count of scalar FFMA.  The extra complex64 slot 0 (index bit 0) lives across the two
// lanes of the pair ("lane slot").
#pragma once
#include "b200q_program.h"

#if defined(__CUDACC__)
#define B200Q_HD __host__ __device__ __forceinline__
#else
#define B200Q_HD inline
#endif

namespace b200q {

template <typename Real> struct cx { Real x, y; };

// Two float lanes in ONE 64-bit register (an aligned even/odd pair): the operand format of the
// Blackwell packed-FP32 instructions.  On the device the carrier is a 64-bit integer so that the
// register allocator keeps the pair together and FFMA2 needs no packing moves.
#if defined(__CUDA_ARCH__)
struct alignas(8) pk { unsigned long long u; };
__device__ __forceinline__ pk pk_make(float x, float y) {
  pk r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.u) : "f"(x), "f"(y)); return r;
}
__device__ __forceinline__ float pk_x(pk a) { return __uint_as_float((unsigned)(a.u & 0xffffffffull)); }
__device__ __forceinline__ float pk_y(pk a) { return __uint_as_float((unsigned)(a.u >> 32)); }
__device__ __forceinline__ pk vmul(pk a, pk b) {
  pk r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u)); return r;
}
__device__ __forceinline__ pk vfma(pk a, pk b, pk c) {
  pk r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.u) : "l"(a.u), "l"(b.u), "l"(c.u)); return r;
}
__device__ __forceinline__ pk vneg(pk a) { pk r; r.u = a.u ^ 0x8000000080000000ull; return r; }
#else
struct alignas(8) pk { float x, y; };
inline pk pk_make(float x, float y) { pk r; r.x = x; r.y = y; return r; }
inline float pk_x(pk a) { return a.x; }
inline float pk_y(pk a) { return a.y; }
inline pk vmul(pk a, pk b) { return pk_make(a.x * b.x, a.y * b.y); }
inline pk vfma(pk a, pk b, pk c) { return pk_make(a.x * b.x + c.x, a.y * b.y + c.y); }
inline pk vneg(pk a) { return pk_make(-a.x, -a.y); }
#endif
struct alignas(16) chunk_f { pk lo, hi; };   // two complex64 amplitudes (AoS or SoA, see program.h)
struct alignas(16) chunk_d { double x, y; };  // one complex128 amplitude

// ---- 2-wide / scalar vector primitives -----------------------------------------------------------
B200Q_HD pk vset(float a, pk*) { return pk_make(a, a); }
B200Q_HD double vset(double a, double*) { return a; }
B200Q_HD double vneg(double a) { return -a; }
B200Q_HD double vmul(double a, double b) { return a * b; }
B200Q_HD double vfma(double a, double b, double c) { return a * b + c; }
// keep lane 0 of `old`, take lane 1 of `nw` (control on index bit 0)
B200Q_HD pk vblend1(pk old, pk nw) { return pk_make(pk_x(old), pk_y(nw)); }
B200Q_HD double vblend1(double, double nw) { return nw; }
B200Q_HD double vhsum(pk a) { return double(pk_x(a)) + double(pk_y(a)); }
B200Q_HD double vhsum(double a) { return a; }
B200Q_HD double vlane1(pk a) { return double(pk_y(a)); }
B200Q_HD double vlane1(double a) { return a; }
B200Q_HD pk vzero(pk*) { return pk_make(0.f, 0.f); }
B200Q_HD double vzero(double*) { return 0.0; }
B200Q_HD pk vmake(float a0, float a1, pk*) { return pk_make(a0, a1); }
B200Q_HD double vmake(double a0, double, double*) { return a0; }
B200Q_HD float vget(pk a, int l) { return l ? pk_y(a) : pk_x(a); }
B200Q_HD double vget(double a, int) { return a; }

template <typename Real> struct Traits;
template <> struct Traits<float> {
  static constexpr int VS = 1;   // log2(amplitudes per 16-byte chunk)
  static constexpr int RB = 5;   // register slots (amplitude bits held per thread)
  using chunk = chunk_f;
  using V = pk;
};
template <> struct Traits<double> {
  static constexpr int VS = 0;
  static constexpr int RB = 4;
  using chunk = chunk_d;
  using V = double;
};
constexpr int NE = 16;  // register elements (chunks) per thread

// XOR swizzle of the chunk index inside the shared-memory tile: the low 3 bits (the 8 x 16-byte bank
// groups of a 128-byte wavefront) are xor-ed with every higher 3-bit group, so a quarter-warp whose
// lanes differ in ANY three chunk-index bits with distinct positions mod 3 is conflict free.
B200Q_HD uint32_t swz(uint32_t c) { return c ^ (((c >> 3) ^ (c >> 6) ^ (c >> 9) ^ (c >> 12)) & 7u); }

// Opaque re-definition of a register value: keeps the register allocator from splitting the live ranges
// of the 16-element arrays across the op loop's back edge (it otherwise rotates them through a second
// set of registers with ~60 MOVs per op, measured in SASS).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void vpin(pk& a) { asm volatile("" : "+l"(a.u)); }
__device__ __forceinline__ void vpin(double& a) { asm volatile("" : "+d"(a)); }
#else
inline void vpin(pk&) {}
inline void vpin(double&) {}
#endif

// In-place swaps with tied operands: written in PTX so that the compiler sees two values updated in
// place instead of two values crossing over (the latter makes it rotate the whole register arrays
// through copies at the head of the op loop -- ~46 MOVs per op, measured in SASS).
#if defined(__CUDA_ARCH__)
// 64-bit swap as two in-place 32-bit xor swaps: with moves through a temporary the register allocator
// renames instead of swapping and pays the copies back at the op-loop back edge (2x the instructions).
__device__ __forceinline__ void vswap(pk& a, pk& b) {
  asm("{\n .reg .b32 a0, a1, b0, b1;\n mov.b64 {a0, a1}, %0;\n mov.b64 {b0, b1}, %1;\n"
      " xor.b32 a0, a0, b0;\n xor.b32 b0, b0, a0;\n xor.b32 a0, a0, b0;\n"
      " xor.b32 a1, a1, b1;\n xor.b32 b1, b1, a1;\n xor.b32 a1, a1, b1;\n"
      " mov.b64 %0, {a0, a1};\n mov.b64 %1, {b0, b1};\n}" : "+l"(a.u), "+l"(b.u));
}
__device__ __forceinline__ void vswap_lane1(pk& a, pk& b) {   // swap the high lanes only
  asm("{\n .reg .b32 al, ah, bl, bh;\n mov.b64 {al, ah}, %0;\n mov.b64 {bl, bh}, %1;\n"
      " mov.b64 %0, {al, bh};\n mov.b64 %1, {bl, ah};\n}" : "+l"(a.u), "+l"(b.u));
}
__device__ __forceinline__ void vswap_lanes(pk& a) {           // swap the two lanes of one register pair
  asm("{\n .reg .b32 al, ah;\n mov.b64 {al, ah}, %0;\n mov.b64 %0, {ah, al};\n}" : "+l"(a.u));
}
__device__ __forceinline__ void vswap(double& a, double& b) {
  asm("{\n .reg .f64 t;\n mov.f64 t, %0;\n mov.f64 %0, %1;\n mov.f64 %1, t;\n}" : "+d"(a), "+d"(b));
}
#else
inline void vswap(pk& a, pk& b) { const pk t = a; a = b; b = t; }
inline void vswap_lane1(pk& a, pk& b) { const float t = a.y; a.y = b.y; b.y = t; }
inline void vswap_lanes(pk& a) { const float t = a.x; a.x = a.y; a.y = t; }
inline void vswap(double& a, double& b) { const double t = a; a = b; b = t; }
#endif
B200Q_HD void vswap_lane1(double& a, double& b) { vswap(a, b); }

// chunk <-> registers.  `soa` selects the complex64 chunk format; shared memory is always SoA.
B200Q_HD void unpack(const chunk_f& v, pk& re, pk& im, bool soa) {
  if (soa) { re = v.lo; im = v.hi; }
  else { re = pk_make(pk_x(v.lo), pk_x(v.hi)); im = pk_make(pk_y(v.lo), pk_y(v.hi)); }
}
B200Q_HD void unpack(const chunk_d& v, double& re, double& im, bool) { re = v.x; im = v.y; }
B200Q_HD chunk_f pack(const pk& re, const pk& im, bool soa, chunk_f*) {
  chunk_f v;
  if (soa) { v.lo = re; v.hi = im; }
  else { v.lo = pk_make(pk_x(re), pk_x(im)); v.hi = pk_make(pk_y(re), pk_y(im)); }
  return v;
}
B200Q_HD chunk_d pack(const double& re, const double& im, bool, chunk_d*) { chunk_d v; v.x = re; v.y = im; return v; }
B200Q_HD chunk_f zero_chunk(chunk_f*) { chunk_f v; v.lo = pk_make(0.f, 0.f); v.hi = v.lo; return v; }
B200Q_HD chunk_d zero_chunk(chunk_d*) { chunk_d v; v.x = v.y = 0.0; return v; }

// ------------------------------------------------------------------------------------------------
// dense 2x2 on a chunk-level slot
// ------------------------------------------------------------------------------------------------
enum { VAR_GENERAL = 0, VAR_REAL = 1, VAR_RXLIKE = 2 };

template <typename V> struct Coef {  // broadcast coefficients (and negated imaginary parts)
  V r00, i00, r01, i01, r10, i10, r11, i11, n00, n01, n10, n11;
};
// `flip`: the slot's two register halves hold the logical values 1 and 0 (X relabelling, see run_round):
// use X M X, i.e. entry (r, c) -> (r^1, c^1).
template <typename Real, typename V>
B200Q_HD Coef<V> make_coef(const cx<Real>* m, uint32_t flip) {
  Coef<V> c;
  const uint32_t f = flip ? 3u : 0u;
  c.r00 = vset(m[0 ^ f].x, (V*)nullptr); c.i00 = vset(m[0 ^ f].y, (V*)nullptr);
  c.r01 = vset(m[1 ^ f].x, (V*)nullptr); c.i01 = vset(m[1 ^ f].y, (V*)nullptr);
  c.r10 = vset(m[2 ^ f].x, (V*)nullptr); c.i10 = vset(m[2 ^ f].y, (V*)nullptr);
  c.r11 = vset(m[3 ^ f].x, (V*)nullptr); c.i11 = vset(m[3 ^ f].y, (V*)nullptr);
  c.n00 = vneg(c.i00); c.n01 = vneg(c.i01); c.n10 = vneg(c.i10); c.n11 = vneg(c.i11);
  return c;
}

// In-place 2x2 butterflies.  Every output is finished by ONE fma whose destination is the register
// of its own input (x0 = r00*ar + t0 overwrites ar, ...): all partial sums t0..t3 are formed first,
// so no output ever needs a move to reach its canonical register.  On the device the complex64
// version is written in PTX on 64-bit operands (fma.rn.f32x2) with tied in/out registers -- with
// plain C++ the register allocator cannot keep the 16-element arrays in place across the slot
// `switch` and spends more instructions on MOVs than on arithmetic (measured with ncu).
template <int VAR, typename V>
B200Q_HD void bfly(V& ar, V& ai, V& br, V& bi, const Coef<V>& m) {
  if (VAR == VAR_REAL) {
    const V t0 = vmul(m.r01, br), t1 = vmul(m.r01, bi), t2 = vmul(m.r10, ar), t3 = vmul(m.r10, ai);
    ar = vfma(m.r00, ar, t0); ai = vfma(m.r00, ai, t1); br = vfma(m.r11, br, t2); bi = vfma(m.r11, bi, t3);
  } else if (VAR == VAR_RXLIKE) {
    const V t0 = vmul(m.n01, bi), t1 = vmul(m.i01, br), t2 = vmul(m.n10, ai), t3 = vmul(m.i10, ar);
    ar = vfma(m.r00, ar, t0); ai = vfma(m.r00, ai, t1); br = vfma(m.r11, br, t2); bi = vfma(m.r11, bi, t3);
  } else {
    const V t0 = vfma(m.n01, bi, vfma(m.r01, br, vmul(m.n00, ai)));
    const V t1 = vfma(m.i01, br, vfma(m.r01, bi, vmul(m.i00, ar)));
    const V t2 = vfma(m.n11, bi, vfma(m.n10, ai, vmul(m.r10, ar)));
    const V t3 = vfma(m.i11, br, vfma(m.i10, ar, vmul(m.r10, ai)));
    ar = vfma(m.r00, ar, t0); ai = vfma(m.r00, ai, t1); br = vfma(m.r11, br, t2); bi = vfma(m.r11, bi, t3);
  }
}
#if defined(__CUDA_ARCH__)
template <int VAR>
__device__ __forceinline__ void bfly_pk(pk& ar, pk& ai, pk& br, pk& bi, const Coef<pk>& m) {
  if (VAR == VAR_REAL) {
    asm("{\n .reg .b64 t0, t1, t2, t3;\n"
        " mul.rn.f32x2 t0, %5, %2;\n mul.rn.f32x2 t1, %5, %3;\n"
        " mul.rn.f32x2 t2, %6, %0;\n mul.rn.f32x2 t3, %6, %1;\n"
        " fma.rn.f32x2 %0, %4, %0, t0;\n fma.rn.f32x2 %1, %4, %1, t1;\n"
        " fma.rn.f32x2 %2, %7, %2, t2;\n fma.rn.f32x2 %3, %7, %3, t3;\n}"
        : "+l"(ar.u), "+l"(ai.u), "+l"(br.u), "+l"(bi.u)
        : "l"(m.r00.u), "l"(m.r01.u), "l"(m.r10.u), "l"(m.r11.u));
  } else if (VAR == VAR_RXLIKE) {
    asm("{\n .reg .b64 t0, t1, t2, t3;\n"
        " mul.rn.f32x2 t0, %6, %3;\n mul.rn.f32x2 t1, %5, %2;\n"
        " mul.rn.f32x2 t2, %8, %1;\n mul.rn.f32x2 t3, %7, %0;\n"
        " fma.rn.f32x2 %0, %4, %0, t0;\n fma.rn.f32x2 %1, %4, %1, t1;\n"
        " fma.rn.f32x2 %2, %9, %2, t2;\n fma.rn.f32x2 %3, %9, %3, t3;\n}"
        : "+l"(ar.u), "+l"(ai.u), "+l"(br.u), "+l"(bi.u)
        : "l"(m.r00.u), "l"(m.i01.u), "l"(m.n01.u), "l"(m.i10.u), "l"(m.n10.u), "l"(m.r11.u));
  } else {
    asm("{\n .reg .b64 t0, t1, t2, t3;\n"
        " mul.rn.f32x2 t0, %6, %1;\n mul.rn.f32x2 t1, %5, %0;\n"
        " mul.rn.f32x2 t2, %10, %0;\n mul.rn.f32x2 t3, %10, %1;\n"
        " fma.rn.f32x2 t0, %7, %2, t0;\n fma.rn.f32x2 t1, %7, %3, t1;\n"
        " fma.rn.f32x2 t2, %12, %1, t2;\n fma.rn.f32x2 t3, %11, %0, t3;\n"
        " fma.rn.f32x2 t0, %9, %3, t0;\n fma.rn.f32x2 t1, %8, %2, t1;\n"
        " fma.rn.f32x2 t2, %15, %3, t2;\n fma.rn.f32x2 t3, %14, %2, t3;\n"
        " fma.rn.f32x2 %0, %4, %0, t0;\n fma.rn.f32x2 %1, %4, %1, t1;\n"
        " fma.rn.f32x2 %2, %13, %2, t2;\n fma.rn.f32x2 %3, %13, %3, t3;\n}"
        : "+l"(ar.u), "+l"(ai.u), "+l"(br.u), "+l"(bi.u)
        : "l"(m.r00.u), "l"(m.i00.u), "l"(m.n00.u), "l"(m.r01.u), "l"(m.i01.u), "l"(m.n01.u), "l"(m.r10.u),
          "l"(m.i10.u), "l"(m.n10.u), "l"(m.r11.u), "l"(m.i11.u), "l"(m.n11.u));
  }
}
template <int VAR>
__device__ __forceinline__ void bfly_dispatch(pk& ar, pk& ai, pk& br, pk& bi, const Coef<pk>& m) {
  bfly_pk<VAR>(ar, ai, br, bi, m);
}
template <int VAR>
__device__ __forceinline__ void bfly_dispatch(double& ar, double& ai, double& br, double& bi, const Coef<double>& m) {
  bfly<VAR, double>(ar, ai, br, bi, m);
}
#else
template <int VAR, typename V>
inline void bfly_dispatch(V& ar, V& ai, V& br, V& bi, const Coef<V>& m) { bfly<VAR, V>(ar, ai, br, bi, m); }
#endif

template <typename V, int S, int VAR, bool CTRL>
B200Q_HD void mat1_chunk(V* re, V* im, const Coef<V>& m, uint32_t cm, uint32_t cv, bool lane_ctrl) {
#pragma unroll
  for (int c = 0; c < NE; ++c) {
    if (c & (1 << S)) continue;
    if (CTRL && (uint32_t(c) & cm) != cv) continue;
    const int d = c | (1 << S);
    if (CTRL && lane_ctrl) {
      V ar = re[c], ai = im[c], br = re[d], bi = im[d];
      bfly_dispatch<VAR>(ar, ai, br, bi, m);
      re[c] = vblend1(re[c], ar); im[c] = vblend1(im[c], ai); re[d] = vblend1(re[d], br); im[d] = vblend1(im[d], bi);
    } else {
      bfly_dispatch<VAR>(re[c], im[c], re[d], im[d], m);
    }
  }
}

template <typename V, int VAR, bool CTRL>
B200Q_HD void mat1_chunk_slot(int s, V* re, V* im, const Coef<V>& m, uint32_t cm, uint32_t cv, bool lane_ctrl) {
  switch (s) {
    case 0: mat1_chunk<V, 0, VAR, CTRL>(re, im, m, cm, cv, lane_ctrl); break;
    case 1: mat1_chunk<V, 1, VAR, CTRL>(re, im, m, cm, cv, lane_ctrl); break;
    case 2: mat1_chunk<V, 2, VAR, CTRL>(re, im, m, cm, cv, lane_ctrl); break;
    default: mat1_chunk<V, 3, VAR, CTRL>(re, im, m, cm, cv, lane_ctrl); break;
  }
}

// dense 2x2 across the two lanes of every element (complex64 index bit 0)
template <typename Real>
B200Q_HD void mat1_lane(pk* re, pk* im, const cx<Real>* m, uint32_t cm, uint32_t cv) {
  const float m00r = m[0].x, m00i = m[0].y, m01r = m[1].x, m01i = m[1].y;
  const float m10r = m[2].x, m10i = m[2].y, m11r = m[3].x, m11i = m[3].y;
#pragma unroll
  for (int c = 0; c < NE; ++c) {
    if ((uint32_t(c) & cm) != cv) continue;
    const float ar = pk_x(re[c]), ai = pk_x(im[c]), br = pk_y(re[c]), bi = pk_y(im[c]);
    re[c] = pk_make(m00r * ar - m00i * ai + m01r * br - m01i * bi, m10r * ar - m10i * ai + m11r * br - m11i * bi);
    im[c] = pk_make(m00r * ai + m00i * ar + m01r * bi + m01i * br, m10r * ai + m10i * ar + m11r * bi + m11i * br);
  }
}
template <typename Real>
B200Q_HD void mat1_lane(double*, double*, const cx<Real>*, uint32_t, uint32_t) {}

// MAT1 dispatch.  `slot` is the amplitude-level register slot; `xm` is the thread's X relabelling mask
// over the chunk-level slots (register element c holds the logical element c ^ xm).
template <typename Real>
B200Q_HD void apply_mat1(const b200q_op_t& op, typename Traits<Real>::V* re, typename Traits<Real>::V* im,
                         const cx<Real>* m, uint32_t xm) {
  using V = typename Traits<Real>::V;
  constexpr int VS = Traits<Real>::VS;
  const uint32_t cm = op.ctrl_reg >> VS, cv = cm & ~xm;
  const bool lane_ctrl = VS && (op.ctrl_reg & 1u);
#ifndef EXP_NO_LANE
  if (VS && op.slot == 0) { mat1_lane<Real>(re, im, m, cm, cv); return; }
#endif
  const int s = int(op.slot) - VS;
  const Coef<V> k = make_coef<Real, V>(m, (xm >> s) & 1u);
  const bool ctrl = cm != 0 || lane_ctrl;
  const int var = (op.flags & B200Q_FLAG_REAL) ? VAR_REAL : ((op.flags & B200Q_FLAG_RXLIKE) ? VAR_RXLIKE : VAR_GENERAL);
  if (!ctrl) {
    if (var == VAR_REAL) mat1_chunk_slot<V, VAR_REAL, false>(s, re, im, k, 0, 0, false);
    else if (var == VAR_RXLIKE) mat1_chunk_slot<V, VAR_RXLIKE, false>(s, re, im, k, 0, 0, false);
    else mat1_chunk_slot<V, VAR_GENERAL, false>(s, re, im, k, 0, 0, false);
  } else {
#ifndef EXP_NO_CTRLMAT
    mat1_chunk_slot<V, VAR_GENERAL, true>(s, re, im, k, cm, cv, lane_ctrl);
#endif
  }
}

// ------------------------------------------------------------------------------------------------
// X (amplitude swap)
// ------------------------------------------------------------------------------------------------
template <typename V, int S>
B200Q_HD void x_chunk(V* re, V* im, uint32_t cm, uint32_t cv, bool lane_ctrl) {
#pragma unroll
  for (int c = 0; c < NE; ++c) {
    if (c & (1 << S)) continue;
    if ((uint32_t(c) & cm) != cv) continue;
    const int d = c | (1 << S);
    if (lane_ctrl) { vswap_lane1(re[c], re[d]); vswap_lane1(im[c], im[d]); }
    else { vswap(re[c], re[d]); vswap(im[c], im[d]); }
  }
}
// CX with exactly one chunk-level control slot CS (and no lane control): the elements to swap are those
// whose REGISTER bit CS equals `cval` -- one thread-level branch, compile-time indices inside.
template <typename V, int S, int CS>
B200Q_HD void x_chunk_c1(V* re, V* im, bool cval) {
  if (cval) {
#pragma unroll
    for (int c = 0; c < NE; ++c)
      if (!(c & (1 << S)) && (c & (1 << CS))) { vswap(re[c], re[c | (1 << S)]); vswap(im[c], im[c | (1 << S)]); }
  } else {
#pragma unroll
    for (int c = 0; c < NE; ++c)
      if (!(c & (1 << S)) && !(c & (1 << CS))) { vswap(re[c], re[c | (1 << S)]); vswap(im[c], im[c | (1 << S)]); }
  }
}
template <typename V, int S>
B200Q_HD void x_chunk_c1_cs(int cs, V* re, V* im, bool cval) {
  switch (cs) {
    case 0: if (S != 0) x_chunk_c1<V, S, (S != 0 ? 0 : 1)>(re, im, cval); break;
    case 1: if (S != 1) x_chunk_c1<V, S, (S != 1 ? 1 : 0)>(re, im, cval); break;
    case 2: if (S != 2) x_chunk_c1<V, S, (S != 2 ? 2 : 0)>(re, im, cval); break;
    default: if (S != 3) x_chunk_c1<V, S, (S != 3 ? 3 : 0)>(re, im, cval); break;
  }
}

B200Q_HD void x_lane(pk* re, pk* im, uint32_t cm, uint32_t cv) {
#pragma unroll
  for (int c = 0; c < NE; ++c) {
    if ((uint32_t(c) & cm) != cv) continue;
    vswap_lanes(re[c]);
    vswap_lanes(im[c]);
  }
}
B200Q_HD void x_lane(double*, double*, uint32_t, uint32_t) {}

// X on a chunk-level slot whose controls are not register slots is a pure relabelling: it only toggles the
// thread's mask `xm` (register element c then holds logical element c ^ xm) -- no data moves at all; the
// mask is folded into the scatter address at the end of the round.
template <typename Real>
B200Q_HD void apply_x(const b200q_op_t& op, typename Traits<Real>::V* re, typename Traits<Real>::V* im, uint32_t& xm) {
  using V = typename Traits<Real>::V;
  constexpr int VS = Traits<Real>::VS;
  const uint32_t cm = op.ctrl_reg >> VS, cv = cm & ~xm;
  const bool lane_ctrl = VS && (op.ctrl_reg & 1u);
#ifndef EXP_NO_LANE
  if (VS && op.slot == 0) { x_lane(re, im, cm, cv); return; }
#endif
  const int s = int(op.slot) - VS;
  if (op.ctrl_reg == 0) { xm ^= 1u << s; return; }
  if (!lane_ctrl && (cm & (cm - 1)) == 0) {   // exactly one chunk-level control
    const int cs = (cm & 1) ? 0 : ((cm & 2) ? 1 : ((cm & 4) ? 2 : 3));
    const bool cval = ((xm >> cs) & 1u) == 0;
    switch (s) {
      case 0: x_chunk_c1_cs<V, 0>(cs, re, im, cval); break;
      case 1: x_chunk_c1_cs<V, 1>(cs, re, im, cval); break;
      case 2: x_chunk_c1_cs<V, 2>(cs, re, im, cval); break;
      default: x_chunk_c1_cs<V, 3>(cs, re, im, cval); break;
    }
    return;
  }
  switch (s) {
    case 0: x_chunk<V, 0>(re, im, cm, cv, lane_ctrl); break;
    case 1: x_chunk<V, 1>(re, im, cm, cv, lane_ctrl); break;
    case 2: x_chunk<V, 2>(re, im, cm, cv, lane_ctrl); break;
    default: x_chunk<V, 3>(re, im, cm, cv, lane_ctrl); break;
  }
}

// ------------------------------------------------------------------------------------------------
// diagonal ops
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void cmul_inplace(pk& re, pk& im, pk pr, pk pi, pk npi) {
  asm("{\n .reg .b64 t0, t1;\n mul.rn.f32x2 t0, %4, %1;\n mul.rn.f32x2 t1, %3, %0;\n"
      " fma.rn.f32x2 %0, %2, %0, t0;\n fma.rn.f32x2 %1, %2, %1, t1;\n}"
      : "+l"(re.u), "+l"(im.u) : "l"(pr.u), "l"(pi.u), "l"(npi.u));
}
#else
inline void cmul_inplace(pk& re, pk& im, pk pr, pk pi, pk npi) {
  const pk t0 = vmul(npi, im), t1 = vmul(pi, re);
  re = vfma(pr, re, t0); im = vfma(pr, im, t1);
}
#endif
B200Q_HD void cmul_inplace(double& re, double& im, double pr, double pi, double npi) {
  const double t0 = npi * im, t1 = pi * re;
  re = pr * re + t0; im = pr * im + t1;
}

// elements with slot bit 0 are multiplied by p0, with bit 1 by p1 (p == 1 is skipped)
template <typename Real, typename V, int S>
B200Q_HD void diag_chunk(V* re, V* im, cx<Real> p0, cx<Real> p1) {
  const bool do0 = !(p0.x == Real(1) && p0.y == Real(0));
  const bool do1 = !(p1.x == Real(1) && p1.y == Real(0));
  const V p0r = vset(p0.x, (V*)nullptr), p0i = vset(p0.y, (V*)nullptr), n0i = vset(-p0.y, (V*)nullptr);
  const V p1r = vset(p1.x, (V*)nullptr), p1i = vset(p1.y, (V*)nullptr), n1i = vset(-p1.y, (V*)nullptr);
#pragma unroll
  for (int c = 0; c < NE; ++c) {
    if (c & (1 << S)) { if (do1) cmul_inplace(re[c], im[c], p1r, p1i, n1i); }
    else { if (do0) cmul_inplace(re[c], im[c], p0r, p0i, n0i); }
  }
}

// per-lane phase vectors (complex64): lane l of every element is multiplied by q[l]
B200Q_HD void diag_lanes(pk* re, pk* im, cx<float> q0, cx<float> q1) {
  const pk pr = pk_make(q0.x, q1.x), pi = pk_make(q0.y, q1.y), npi = pk_make(-q0.y, -q1.y);
#pragma unroll
  for (int c = 0; c < NE; ++c) cmul_inplace(re[c], im[c], pr, pi, npi);
}
B200Q_HD void diag_lanes(double*, double*, cx<double>, cx<double>) {}

// fully general (slow) path: any mix of register / lane selectors and register controls
template <typename Real>
B200Q_HD void diag_generic(const b200q_op_t& op, typename Traits<Real>::V* re, typename Traits<Real>::V* im,
                           const cx<Real>* d, uint32_t tsel, uint32_t xm) {
  using V = typename Traits<Real>::V;
  constexpr int VS = Traits<Real>::VS;
  constexpr int NL = 1 << VS;
#pragma unroll
  for (int c = 0; c < NE; ++c) {
    Real lr[2] = {Real(1), Real(1)}, li[2] = {Real(0), Real(0)};
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const uint32_t i = ((uint32_t(c) ^ xm) << VS) | uint32_t(l);  // LOGICAL amplitude-level register index
      if ((i & op.ctrl_reg) != op.ctrl_reg) continue;
      uint32_t idx = tsel;
      if (op.dsel_slot[0] != 0xff) idx |= (i >> op.dsel_slot[0]) & 1u;
      if (op.dsel_slot[1] != 0xff) idx |= ((i >> op.dsel_slot[1]) & 1u) << 1;
      lr[l] = d[idx].x; li[l] = d[idx].y;
    }
    const V pr = vmake(lr[0], lr[1], (V*)nullptr), pi = vmake(li[0], li[1], (V*)nullptr);
    const V npi = vmake(-li[0], -li[1], (V*)nullptr);
    cmul_inplace(re[c], im[c], pr, pi, npi);
  }
}

// DIAG dispatch.  FOLD: phases that are uniform over the thread's registers are accumulated into the
// scalar (rho_r, rho_i) and applied once per round.
template <typename Real, bool FOLD>
B200Q_HD void apply_diag(const b200q_op_t& op, typename Traits<Real>::V* re, typename Traits<Real>::V* im,
                         const cx<Real>* d, uint32_t tsel, uint32_t xm, Real& rho_r, Real& rho_i, bool& rho_dirty) {
  using V = typename Traits<Real>::V;
  constexpr int VS = Traits<Real>::VS;
  const int s0 = op.dsel_slot[0], s1 = op.dsel_slot[1];
  const int nreg = (s0 != 0xff) + (s1 != 0xff);
  const uint32_t cm = op.ctrl_reg >> VS;
  const bool lane_ctrl = VS && (op.ctrl_reg & 1u);
  if (nreg == 2 || cm != 0) {
#ifndef EXP_NO_DIAGGEN
    diag_generic<Real>(op, re, im, d, tsel, xm);
#endif
    return;
  }
  if (nreg == 0) {
    const cx<Real> p = d[tsel];
    if (lane_ctrl) {  // control on index bit 0: only lane 1 is multiplied
      cx<Real> one; one.x = Real(1); one.y = Real(0);
      diag_lanes(re, im, one, p);
      return;
    }
    if (FOLD) {
      const Real r = rho_r * p.x - rho_i * p.y, i = rho_r * p.y + rho_i * p.x;
      rho_r = r; rho_i = i; rho_dirty = true;
    } else {
      const V pr = vset(p.x, (V*)nullptr), pi = vset(p.y, (V*)nullptr), npi = vset(-p.y, (V*)nullptr);
#pragma unroll
      for (int c = 0; c < NE; ++c) cmul_inplace(re[c], im[c], pr, pi, npi);
    }
    return;
  }
  // exactly one register selector
  const int j = (s0 != 0xff) ? 0 : 1;
  const int slot = j == 0 ? s0 : s1;
  if (VS && slot == 0) { diag_lanes(re, im, d[tsel], d[tsel | (1u << j)]); return; }
  if (lane_ctrl) { diag_generic<Real>(op, re, im, d, tsel, xm); return; }
  const uint32_t fl = (xm >> (slot - VS)) & 1u;   // X relabelling: the register halves hold logical 1 / 0
  const cx<Real> p0 = d[fl ? (tsel | (1u << j)) : tsel], p1 = d[fl ? tsel : (tsel | (1u << j))];
  switch (slot - VS) {
    case 0: diag_chunk<Real, V, 0>(re, im, p0, p1); break;
    case 1: diag_chunk<Real, V, 1>(re, im, p0, p1); break;
    case 2: diag_chunk<Real, V, 2>(re, im, p0, p1); break;
    default: diag_chunk<Real, V, 3>(re, im, p0, p1); break;
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
// tid, tid + nthreads, ...; a quad is 4 consecutive pool elements of one op.  `flip_adjoint` makes
// the pool hold U^dagger of every op (the reverse sweep).
template <typename Real>
B200Q_HD void fill_pool(const b200q_pass_t& P, int tid, int nthreads, cx<Real>* pool, const cx<Real>* mats,
                        bool flip_adjoint) {
  const int nquads = P.pool_elems >> 2;
  for (int q = tid; q < nquads; q += nthreads) {
    const int e0 = q << 2;
    int o = 0;
    for (int t = 0; t < P.n_ops; ++t) {
      const int off = P.ops[t].pool_off;
      if (P.ops[t].pool_n != 0 && e0 >= off && e0 < off + P.ops[t].pool_n) o = t;
    }
    const b200q_op_t& op = P.ops[o];
    const bool adj = ((op.flags & B200Q_FLAG_ADJOINT) != 0) != flip_adjoint;
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

// Addressing of one thread's 16 elements in a round.
template <typename Real> struct RoundAddr {
  uint32_t lb;       // tile-local amplitude index of the item (register bits zero)
  uint64_t gbase;    // chunk index in global memory
  uint32_t sbase;    // swizzled chunk index in the tile
  bool active;
};   // the slot strides stay in the shared-memory RoundTab: keeping them in registers across the op loop spilled

// Per-round address tables, built once per CTA in shared memory (the per-thread bit-deposit loops they replace
// were 13 % of all executed instructions): the item index tid is split into its low 4 and high 5 bits, each
// looked up in a small table of deposited local / physical offsets.
struct RoundTab {
  uint32_t lb_lo[16], lb_hi[32];   // tile-local amplitude offset
  uint64_t pb_lo[16], pb_hi[32];   // physical amplitude offset
  uint64_t gst[4];                 // global chunk stride of each chunk-level slot
  uint32_t sst[4];                 // swizzled tile stride of each chunk-level slot
};

template <typename Real>
B200Q_HD void fill_round_tabs(const b200q_pass_t& P, int tid, int nthreads, RoundTab* tabs) {
  constexpr int VS = Traits<Real>::VS, RB = Traits<Real>::RB;
  const int item_bits = int(P.tile_bits) - RB;
  const int per_round = 16 + 32 + 4;
  for (int e = tid; e < int(P.n_rounds) * per_round; e += nthreads) {
    const int r = e / per_round, j = e % per_round;
    const b200q_round_t& Rd = P.rounds[r];
    RoundTab& T = tabs[r];
    if (Rd.direct) continue;
    if (j < 48) {
      const int lo = j < 16;
      const int v = lo ? j : j - 16;
      const int k0 = lo ? 0 : 4, k1 = lo ? 4 : 9;
      uint32_t lb = 0;
      uint64_t pb = 0;
      for (int k = k0; k < k1 && k < item_bits; ++k) {
        const uint32_t bit = (uint32_t(v) >> (k - k0)) & 1u;
        const int loc = Rd.nonreg_bit[k];
        lb |= bit << loc;
        pb |= uint64_t(bit) << P.tile_phys[loc];
      }
      if (lo) { T.lb_lo[v] = lb; T.pb_lo[v] = pb; } else { T.lb_hi[v] = lb; T.pb_hi[v] = pb; }
    } else {
      const int s = j - 48;
      const int loc = Rd.slot_bit[s + VS];
      T.gst[s] = 1ull << (int(P.tile_phys[loc]) - VS);
      T.sst[s] = swz(1u << (loc - VS));
    }
  }
}

template <typename Real>
B200Q_HD RoundAddr<Real> round_addr(const b200q_pass_t& P, const RoundTab& T, int tid, uint64_t cta_base) {
  constexpr int VS = Traits<Real>::VS, RB = Traits<Real>::RB;
  RoundAddr<Real> A;
  const int item_bits = int(P.tile_bits) - RB;
  A.active = tid < (1 << item_bits);
  const int lo = tid & 15, hi = (tid >> 4) & 31;
  const uint32_t lb = T.lb_lo[lo] | T.lb_hi[hi];
  const uint64_t pb = cta_base | T.pb_lo[lo] | T.pb_hi[hi];
  A.lb = lb;
  A.gbase = pb >> VS;
  A.sbase = swz(lb >> VS);
  return A;
}

B200Q_HD uint64_t gidx(uint64_t gbase, const uint64_t* gst, int c) {
  return gbase ^ ((c & 1) ? gst[0] : 0) ^ ((c & 2) ? gst[1] : 0) ^ ((c & 4) ? gst[2] : 0) ^ ((c & 8) ? gst[3] : 0);
}
B200Q_HD uint32_t sidx(uint32_t sbase, const uint32_t* sst, int c) {
  return sbase ^ ((c & 1) ? sst[0] : 0) ^ ((c & 2) ? sst[1] : 0) ^ ((c & 4) ? sst[2] : 0) ^ ((c & 8) ? sst[3] : 0);
}

// Destination tables of the fused exchange: entry [b][v] = OR over the set bits of v of the image of chunk-index
// bit 8 b + (bit); entry [0][0] also carries the images of the source rank bits (`base`).  5 x 256 x 8 bytes.
#define B200Q_DEST_TAB_ENTRIES 1280
B200Q_HD void fill_dest_tab(const b200q_remote_t& R, int tid, int nthreads, uint64_t* tab) {
  for (int e = tid; e < B200Q_DEST_TAB_ENTRIES; e += nthreads) {
    const int b = e >> 8, v = e & 255;
    uint64_t d = b == 0 ? R.base : 0ull;
    for (int j = 0; j < 8; ++j) {
      const int bit = 8 * b + j;
      if (!((v >> j) & 1) || bit >= R.n_chunk_bits) continue;
      const int pos = R.perm[bit];
      d |= pos < R.n_chunk_bits ? (1ull << pos) : (1ull << (B200Q_DEST_RANK_SHIFT + pos - R.n_chunk_bits));
    }
    tab[e] = d;
  }
}

// Global addressing without bounds checks (the state is not padded): the register-slot bits of gbase are zero,
// so element c lives at base + sum of the strides of its set bits -- 64-bit pointer adds shared between the 16
// elements instead of a 64-bit XOR chain, compare and select per element.  `sgn` makes stride s negative
// (scatter with the slot relabelled: the base then already contains the stride).
template <typename Real, typename Ptr>
B200Q_HD void element_ptrs(Ptr base, const uint64_t* gst, uint32_t sgn, Ptr* p) {
  p[0] = base;
#pragma unroll
  for (int c = 1; c < NE; ++c) {
    const int s = (c & 1) ? 0 : ((c & 2) ? 1 : ((c & 4) ? 2 : 3));   // lowest set bit
    const int64_t st = ((sgn >> s) & 1u) ? -int64_t(gst[s]) : int64_t(gst[s]);
    p[c] = p[c & (c - 1)] + st;
  }
}

template <typename Real>
B200Q_HD void gather(const RoundAddr<Real>& A, const RoundTab& T, bool from_global, bool soa_global,
                     const typename Traits<Real>::chunk* tile, const typename Traits<Real>::chunk* gstate,
                     uint64_t total_chunks, bool padded, typename Traits<Real>::V* re, typename Traits<Real>::V* im) {
  using chunk = typename Traits<Real>::chunk;
  if (from_global) {
    uint64_t gst[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) gst[s] = T.gst[s];
    if (!padded) {
      const chunk* p[NE];
      element_ptrs<Real, const chunk*>(gstate + A.gbase, gst, 0u, p);
#pragma unroll
      for (int c = 0; c < NE; ++c) unpack(*p[c], re[c], im[c], soa_global);
    } else {
#pragma unroll
      for (int c = 0; c < NE; ++c) {
        const uint64_t idx = gidx(A.gbase, gst, c);
        chunk v = zero_chunk((chunk*)nullptr);
        if (idx < total_chunks) v = gstate[idx];
        unpack(v, re[c], im[c], soa_global);
      }
    }
  } else {
    uint32_t sst[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) sst[s] = T.sst[s];
#pragma unroll
    for (int c = 0; c < NE; ++c) unpack(tile[sidx(A.sbase, sst, c)], re[c], im[c], true);
  }
}

// `xm`: X relabelling mask -- register element c is written to logical element c ^ xm.
template <typename Real>
B200Q_HD void scatter(const RoundAddr<Real>& A, const RoundTab& T, bool to_global, bool soa_global,
                      typename Traits<Real>::chunk* tile,
                      typename Traits<Real>::chunk* gstate, uint64_t total_chunks, bool padded, uint32_t xm,
                      const typename Traits<Real>::V* re, const typename Traits<Real>::V* im,
                      const b200q_remote_t* remote = nullptr, const uint64_t* dest_tab = nullptr) {
  using chunk = typename Traits<Real>::chunk;
  if (to_global) {
    uint64_t gst[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) gst[s] = T.gst[s];
    uint64_t gbase = A.gbase;
#pragma unroll
    for (int s = 0; s < 4; ++s)
      if ((xm >> s) & 1u) gbase ^= gst[s];
    if (remote != nullptr && remote->enabled) {
      // fused exchange: destination = bit permutation of the chunk index, four byte-indexed table lookups
      // (dest_tab, built once per CTA by fill_dest_tab); rank in the high bits selects the peer buffer
#pragma unroll
      for (int c = 0; c < NE; ++c) {
        const uint64_t idx = gidx(gbase, gst, c);
        if (idx < total_chunks) {
          const uint64_t d = dest_tab[idx & 255u] | dest_tab[256 + ((idx >> 8) & 255u)] |
                             dest_tab[512 + ((idx >> 16) & 255u)] | dest_tab[768 + ((idx >> 24) & 255u)] |
                             dest_tab[1024 + ((idx >> 32) & 255u)];
          chunk* dst = reinterpret_cast<chunk*>(remote->peer[d >> B200Q_DEST_RANK_SHIFT]);
          dst[d & ((1ull << B200Q_DEST_RANK_SHIFT) - 1ull)] = pack(re[c], im[c], soa_global, (chunk*)nullptr);
        }
      }
      return;
    }
    if (!padded) {
      chunk* p[NE];
      element_ptrs<Real, chunk*>(gstate + gbase, gst, xm, p);
#pragma unroll
      for (int c = 0; c < NE; ++c) *p[c] = pack(re[c], im[c], soa_global, (chunk*)nullptr);
    } else {
#pragma unroll
      for (int c = 0; c < NE; ++c) {
        const uint64_t idx = gidx(gbase, gst, c);
        if (idx < total_chunks) gstate[idx] = pack(re[c], im[c], soa_global, (chunk*)nullptr);
      }
    }
  } else {
    uint32_t sst[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) sst[s] = T.sst[s];
    uint32_t sbase = A.sbase;
#pragma unroll
    for (int s = 0; s < 4; ++s)
      if ((xm >> s) & 1u) sbase ^= sst[s];
#pragma unroll
    for (int c = 0; c < NE; ++c) tile[sidx(sbase, sst, c)] = pack(re[c], im[c], true, (chunk*)nullptr);
  }
}

// thread-level selector value of a DIAG op
B200Q_HD uint32_t diag_tsel(const b200q_op_t& op, uint64_t cta_base, uint32_t lb) {
  uint32_t sel = 0;
  if ((cta_base & op.dsel_glob[0]) | uint64_t(lb & op.dsel_loc[0])) sel |= 1u;
  if ((cta_base & op.dsel_glob[1]) | uint64_t(lb & op.dsel_loc[1])) sel |= 2u;
  return sel;
}

// ------------------------------------------------------------------------------------------------
// forward path: coefficient records, per-tile setup, one register round of one thread
// ------------------------------------------------------------------------------------------------
// Coefficient record of a MAT1 op in shared memory, built ONCE per CTA (fill_coefs): 2 flip states x 16
// scalars, indexed by the op's position in the pass.  The op loop only issues one to three 128-bit
// shared loads per op (packing 12 broadcast coefficients per op per thread, with flip-dependent indexing
// and sign flips, was 25 % of all executed instructions).  Layouts (per flip state):
//   dense    [r00 r01 r10 r11 | i01 n01 i10 n10 | i00 n00 i11 n11]      (n = negated imaginary part;
//            REAL ops read the first quad, RXLIKE ops the first two, GENERAL ops all three)
//   rotation [u, -u, v, -v | neg, ...]  three in-place shears  a += U b; b += V a; a += U b  with
//            U = e01 / (1 + c), V = e10 (U = i u, V = i v for the Rx family; u, v real for the Ry family),
//            after the matrix has been negated if c = e00 < 0 (`neg` = 1: the sign goes to the pass scalar
//            or, for controlled ops, to the thread's phase rho).  No temporaries: the register allocator
//            has nothing to shuffle back at the op-loop back edge.
#define B200Q_COEF_PER_FLIP 12
#define B200Q_COEF_PER_OP 24

B200Q_HD bool code_is_rot(int code) { return code >= B200Q_CODE_MAT1_ROTX && code < B200Q_CODE_MAT1_ROTY + 4; }
B200Q_HD bool code_is_had(int code) { return code >= B200Q_CODE_MAT1_HAD && code < B200Q_CODE_MAT1_HAD + 4; }

// effective 2x2 entry (row-major idx) of a MAT1 op for flip state f, adjoint folded in
template <typename Real>
B200Q_HD cx<Real> mat1_entry(const b200q_op_t& op, const cx<Real>* mats, int idx, int f, bool flip_adjoint = false) {
  if (f) idx ^= 3;   // X M X: entry (r, c) -> (r^1, c^1)
  const bool adj = ((op.flags & B200Q_FLAG_ADJOINT) != 0) != flip_adjoint;
  if (adj && (idx == 1 || idx == 2)) idx ^= 3;   // transpose
  cx<Real> v = mats[op.mat_src + idx];
  if (adj) v.y = -v.y;
  return v;
}

// Shear form of a multiplication by the unit-modulus phase d: (x, y) <- (x + t y, ...) three times, after d has
// been negated if Re d < 0.  mode: 0 identity, 1 shears, 2 shears + negate, 3 negate only, 4 not unit modulus
// (general complex multiply from the raw value).
template <typename Real>
B200Q_HD void phase_shears(cx<Real> d, Real* out) {
  const Real tol = sizeof(Real) == 4 ? Real(1e-6) : Real(1e-13);
  const Real dev = d.x * d.x + d.y * d.y - Real(1);
  Real t = Real(0), s = Real(0), mode;
  if (dev > tol || dev < -tol) mode = Real(4);
  else if (d.y == Real(0)) mode = d.x > Real(0) ? Real(0) : Real(3);
  else {
    const Real sg = d.x < Real(0) ? Real(-1) : Real(1);
    const Real c = sg * d.x, si = sg * d.y;
    t = -si / (Real(1) + c);
    s = si;
    mode = sg < Real(0) ? Real(2) : Real(1);
  }
  out[0] = t; out[1] = s; out[2] = mode; out[3] = Real(0);
}

// `flip_adjoint`: build the records of U^dagger for every op (the reverse sweep).
template <typename Real>
B200Q_HD void fill_coefs(const b200q_pass_t& P, int tid, int nthreads, Real* coef, const cx<Real>* mats,
                         bool flip_adjoint = false) {
  for (int e = tid; e < int(P.n_ops) * 2; e += nthreads) {
    const int o = e >> 1, f = e & 1;
    const b200q_op_t& op = P.ops[o];
    Real* k = coef + o * B200Q_COEF_PER_OP + f * B200Q_COEF_PER_FLIP;
    if (op.kind == B200Q_OP_DIAG) {
      // [0..15]: shear entries (t, s, mode, 0) of the 4 diagonal values; [16..23]: the raw values
      const bool adj = ((op.flags & B200Q_FLAG_ADJOINT) != 0) != flip_adjoint;
      const int dim = 1 << int(op.k);
      for (int i = 2 * f; i < 2 * f + 2; ++i) {
        cx<Real> d;
        d.x = Real(1); d.y = Real(0);
        if (i < dim) { d = mats[op.mat_src + i * (dim + 1)]; if (adj) d.y = -d.y; }
        phase_shears<Real>(d, coef + o * B200Q_COEF_PER_OP + 4 * i);
        coef[o * B200Q_COEF_PER_OP + 16 + 2 * i] = d.x;
        coef[o * B200Q_COEF_PER_OP + 17 + 2 * i] = d.y;
      }
      continue;
    }
    if (op.kind != B200Q_OP_MAT1) continue;
    const cx<Real> m00 = mat1_entry(op, mats, 0, f, flip_adjoint), m01 = mat1_entry(op, mats, 1, f, flip_adjoint);
    const cx<Real> m10 = mat1_entry(op, mats, 2, f, flip_adjoint), m11 = mat1_entry(op, mats, 3, f, flip_adjoint);
    if (code_is_rot(op.code)) {
      const bool isx = op.code < B200Q_CODE_MAT1_ROTY;
      const Real sg = m00.x < Real(0) ? Real(-1) : Real(1);
      const Real c = sg * m00.x;
      const Real e01 = sg * (isx ? m01.y : m01.x), e10 = sg * (isx ? m10.y : m10.x);
      const Real u = e01 / (Real(1) + c), v = e10;
      k[0] = u; k[1] = -u; k[2] = v; k[3] = -v;
      k[4] = sg < Real(0) ? Real(1) : Real(0);
      for (int j = 5; j < B200Q_COEF_PER_FLIP; ++j) k[j] = Real(0);
    } else {
      k[0] = m00.x; k[1] = m01.x; k[2] = m10.x; k[3] = m11.x;
      k[4] = m01.y; k[5] = -m01.y; k[6] = m10.y; k[7] = -m10.y;
      k[8] = m00.y; k[9] = -m00.y; k[10] = m11.y; k[11] = -m11.y;
      for (int j = 12; j < B200Q_COEF_PER_FLIP; ++j) k[j] = Real(0);
    }
  }
}

// Op words: the per-op fields the round loop needs, 16 bytes per op in shared memory (one 128-bit load,
// prefetched one op ahead) instead of a handful of indexed constant-bank loads per op.
struct alignas(16) OpWord {
  uint32_t x;   // code | arg << 8 | flags << 16 | op index << 24
  uint32_t y;   // thread-level control mask (tile-local bits)
  uint32_t z;   // DIAG: tile-local selector masks, dsel_loc[0] | dsel_loc[1] << 16
  uint32_t w;
};
#define B200Q_OW_CTRL_LOC 0x10000u
#define B200Q_OW_CTRL_GLOB 0x20000u
#define B200Q_OW_DSEL_GLOB0 0x40000u
#define B200Q_OW_DSEL_GLOB1 0x80000u

B200Q_HD void fill_opwords(const b200q_pass_t& P, int tid, int nthreads, OpWord* words) {
  for (int o = tid; o <= int(P.n_ops); o += nthreads) {
    OpWord w;
    w.x = B200Q_CODE_NONE; w.y = 0; w.z = 0; w.w = 0;
    if (o < int(P.n_ops)) {
      const b200q_op_t& op = P.ops[o];
      uint32_t fl = 0;
      if (op.ctrl_loc) fl |= B200Q_OW_CTRL_LOC;
      if (op.ctrl_glob) fl |= B200Q_OW_CTRL_GLOB;
      if (op.dsel_glob[0]) fl |= B200Q_OW_DSEL_GLOB0;
      if (op.dsel_glob[1]) fl |= B200Q_OW_DSEL_GLOB1;
      w.x = uint32_t(op.code) | (uint32_t(op.arg) << 8) | fl | (uint32_t(o) << 24);
      w.y = op.ctrl_loc;
      w.z = (op.dsel_loc[0] & 0xffffu) | (op.dsel_loc[1] << 16);
    }
    words[o] = w;
  }
}

// Product of the deferred common scalars of the pass: entry (0,0) of every Hadamard-structured op and the
// sign of every UN-controlled rotation op whose matrix was negated.  Applied by the last round of the pass.
template <typename Real>
B200Q_HD double pass_scale(const b200q_pass_t& P, const cx<Real>* mats) {
  double g = 1.0;
  for (int o = 0; o < int(P.n_ops); ++o) {
    const b200q_op_t& op = P.ops[o];
    if (code_is_had(op.code)) g *= double(mats[op.mat_src].x);
    else if (code_is_rot(op.code) && !op.tctrl && mats[op.mat_src].x < Real(0)) g = -g;
  }
  return g;
}

// Per-tile, CTA-uniform: which ops pass their global (outside-the-tile) controls.
B200Q_HD uint64_t tile_enabled(const b200q_pass_t& P, uint64_t cta_base) {
  uint64_t en = ~0ull;
  for (int i = 0; i < int(P.n_gctrl); ++i) {
    const int o = P.gctrl_ops[i];
    if ((cta_base & P.ops[o].ctrl_glob) != P.ops[o].ctrl_glob) en &= ~(1ull << o);
  }
  return en;
}

template <typename Real> struct alignas(16) CoefQuad;
template <> struct alignas(16) CoefQuad<float> { float a, b, c, d; };
template <> struct alignas(16) CoefQuad<double> { double a, b; };
B200Q_HD void load_quad(const float* k, int q, float* out) {
  const CoefQuad<float> v = reinterpret_cast<const CoefQuad<float>*>(k)[q];
  out[0] = v.a; out[1] = v.b; out[2] = v.c; out[3] = v.d;
}
B200Q_HD void load_quad(const double* k, int q, double* out) {
  const CoefQuad<double> v0 = reinterpret_cast<const CoefQuad<double>*>(k)[2 * q];
  const CoefQuad<double> v1 = reinterpret_cast<const CoefQuad<double>*>(k)[2 * q + 1];
  out[0] = v0.a; out[1] = v0.b; out[2] = v1.a; out[3] = v1.b;
}

template <typename Real, int S, int VAR>
B200Q_HD void mat1_fast(typename Traits<Real>::V* re, typename Traits<Real>::V* im, const Real* k) {
  using V = typename Traits<Real>::V;
  Coef<V> c;
  Real q[4];
  load_quad(k, 0, q);
  c.r00 = vset(q[0], (V*)nullptr); c.r01 = vset(q[1], (V*)nullptr);
  c.r10 = vset(q[2], (V*)nullptr); c.r11 = vset(q[3], (V*)nullptr);
  if (VAR != VAR_REAL) {
    load_quad(k, 1, q);
    c.i01 = vset(q[0], (V*)nullptr); c.n01 = vset(q[1], (V*)nullptr);
    c.i10 = vset(q[2], (V*)nullptr); c.n10 = vset(q[3], (V*)nullptr);
  }
  if (VAR == VAR_GENERAL) {
    load_quad(k, 2, q);
    c.i00 = vset(q[0], (V*)nullptr); c.n00 = vset(q[1], (V*)nullptr);
    c.i11 = vset(q[2], (V*)nullptr); c.n11 = vset(q[3], (V*)nullptr);
  }
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    if (e & (1 << S)) continue;
    bfly_dispatch<VAR>(re[e], im[e], re[e | (1 << S)], im[e | (1 << S)], c);
  }
}

// x += k * y, in place
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void axpy_inplace(pk& x, pk k, pk y) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x.u) : "l"(k.u), "l"(y.u));
}
#else
inline void axpy_inplace(pk& x, pk k, pk y) { x = vfma(k, y, x); }
#endif
B200Q_HD void axpy_inplace(double& x, double k, double y) { x = k * y + x; }

// Rotation-structured op as three in-place shears (see the record layout above).  ISX: the Rx family
// (imaginary shears), else the Ry family (real shears).
template <typename Real, int S, bool ISX>
B200Q_HD void rot_fast(typename Traits<Real>::V* re, typename Traits<Real>::V* im, const Real* k, bool& neg) {
  using V = typename Traits<Real>::V;
  Real q[4], q1[4];
  load_quad(k, 0, q);
  load_quad(k, 1, q1);
  neg = q1[0] != Real(0);
  const V u = vset(q[0], (V*)nullptr), nu = vset(q[1], (V*)nullptr);
  const V v = vset(q[2], (V*)nullptr), nv = vset(q[3], (V*)nullptr);
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    if (e & (1 << S)) continue;
    V& ar = re[e]; V& ai = im[e]; V& br = re[e | (1 << S)]; V& bi = im[e | (1 << S)];
    if (ISX) {
      axpy_inplace(ar, nu, bi); axpy_inplace(ai, u, br);
      axpy_inplace(br, nv, ai); axpy_inplace(bi, v, ar);
      axpy_inplace(ar, nu, bi); axpy_inplace(ai, u, br);
    } else {
      axpy_inplace(ar, u, br); axpy_inplace(ai, u, bi);
      axpy_inplace(br, v, ar); axpy_inplace(bi, v, ai);
      axpy_inplace(ar, u, br); axpy_inplace(ai, u, bi);
    }
  }
}

// Hadamard-structured op x * [[1, 1], [1, -1]]: two in-place packed instructions per component
// (a += b; b = a - 2b); the common scalar x is applied once per pass (pass_scale).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void had_inplace(pk& ar, pk& ai, pk& br, pk& bi, pk m2) {
  asm("add.rn.f32x2 %0, %0, %2;\n add.rn.f32x2 %1, %1, %3;\n"
      " fma.rn.f32x2 %2, %4, %2, %0;\n fma.rn.f32x2 %3, %4, %3, %1;"
      : "+l"(ar.u), "+l"(ai.u), "+l"(br.u), "+l"(bi.u) : "l"(m2.u));
}
#else
inline void had_inplace(pk& ar, pk& ai, pk& br, pk& bi, pk m2) {
  ar = pk_make(ar.x + br.x, ar.y + br.y); ai = pk_make(ai.x + bi.x, ai.y + bi.y);
  br = vfma(m2, br, ar); bi = vfma(m2, bi, ai);
}
#endif
B200Q_HD void had_inplace(double& ar, double& ai, double& br, double& bi, double m2) {
  ar += br; ai += bi; br = m2 * br + ar; bi = m2 * bi + ai;
}
template <typename Real, int S>
B200Q_HD void had_fast(typename Traits<Real>::V* re, typename Traits<Real>::V* im, bool flip) {
  using V = typename Traits<Real>::V;
  const V m2 = vset(Real(-2), (V*)nullptr);
  if (flip) {
#pragma unroll
    for (int e = 0; e < NE; ++e)
      if (!(e & (1 << S))) had_inplace(re[e | (1 << S)], im[e | (1 << S)], re[e], im[e], m2);
  } else {
#pragma unroll
    for (int e = 0; e < NE; ++e)
      if (!(e & (1 << S))) had_inplace(re[e], im[e], re[e | (1 << S)], im[e | (1 << S)], m2);
  }
}

// complex64: exchange the lane bit with chunk slot S, in registers: A = element with slot bit 0, B = with slot
// bit 1; A' = (A.x, B.x), B' = (A.y, B.y), i.e. A.y <-> B.x, written as an in-place xor swap (with moves the
// register allocator copies the whole 64-register state to fresh registers and back).  The relabelling
// (flip) state of the slot travels with the data: bit S of xm is exchanged with the lane-flip bit 4; the
// planner brackets every lane-touching op between two LSWAPs and never puts an X relabelling in between,
// so the lane-flip bit is zero outside the brackets.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void lane_xpose(pk& a, pk& b) {
  asm("{\n .reg .b32 a0, a1, b0, b1;\n mov.b64 {a0, a1}, %0;\n mov.b64 {b0, b1}, %1;\n"
      " xor.b32 a1, a1, b0;\n xor.b32 b0, b0, a1;\n xor.b32 a1, a1, b0;\n"
      " mov.b64 %0, {a0, a1};\n mov.b64 %1, {b0, b1};\n}" : "+l"(a.u), "+l"(b.u));
}
#else
inline void lane_xpose(pk& a, pk& b) { const float t = a.y; a.y = b.x; b.x = t; }
#endif
B200Q_HD void lane_xpose(double&, double&) {}

template <typename V, int S>
B200Q_HD void lane_swap(V* re, V* im, uint32_t& xm) {
#pragma unroll
  for (int e = 0; e < NE; ++e)
    if (!(e & (1 << S))) { lane_xpose(re[e], re[e | (1 << S)]); lane_xpose(im[e], im[e | (1 << S)]); }
  const uint32_t d = ((xm >> S) ^ (xm >> 4)) & 1u;
  xm ^= (d << S) | (d << 4);
}

// One half (slot bit == H) of a register-slot diagonal: phase in shear form (see phase_shears).
template <typename Real, int S, int H>
B200Q_HD void diag_half(typename Traits<Real>::V* re, typename Traits<Real>::V* im, const Real* ent, const Real* raw) {
  using V = typename Traits<Real>::V;
  Real q[4];
  load_quad(ent, 0, q);
  const int mode = int(q[2]);
  if (mode == 0) return;
  if (mode == 4) {
    const V pr = vset(raw[0], (V*)nullptr), pi = vset(raw[1], (V*)nullptr), npi = vset(-raw[1], (V*)nullptr);
#pragma unroll
    for (int e = 0; e < NE; ++e)
      if (((e >> S) & 1) == H) cmul_inplace(re[e], im[e], pr, pi, npi);
    return;
  }
  if (mode <= 2) {
    const V t = vset(q[0], (V*)nullptr), s = vset(q[1], (V*)nullptr);
#pragma unroll
    for (int e = 0; e < NE; ++e)
      if (((e >> S) & 1) == H) { axpy_inplace(re[e], t, im[e]); axpy_inplace(im[e], s, re[e]); axpy_inplace(re[e], t, im[e]); }
  }
  if (mode >= 2) {
#pragma unroll
    for (int e = 0; e < NE; ++e)
      if (((e >> S) & 1) == H) { re[e] = vneg(re[e]); im[e] = vneg(im[e]); }
  }
}

// thread-level value of DIAG selector j (tile-local bit of the thread's item, or a bit outside the tile)
B200Q_HD uint32_t dsel_value(const b200q_pass_t& P, const OpWord& w, int j, uint32_t lb, uint64_t cta_base) {
  const uint32_t loc = j ? (w.z >> 16) : (w.z & 0xffffu);
  uint32_t v = (lb & loc) ? 1u : 0u;
  if (w.x & (j ? B200Q_OW_DSEL_GLOB1 : B200Q_OW_DSEL_GLOB0))
    v |= (cta_base & P.ops[w.x >> 24].dsel_glob[j]) ? 1u : 0u;
  return v;
}

template <typename Real, int S>
B200Q_HD void diag_reg(const b200q_pass_t& P, const OpWord& w, typename Traits<Real>::V* re,
                       typename Traits<Real>::V* im, const Real* rec, uint32_t lb, uint64_t cta_base, uint32_t xm) {
  const int j = (w.x >> 8) & 1, other = j ^ 1;
  const uint32_t base = dsel_value(P, w, other, lb, cta_base) << other;
  const uint32_t f = (xm >> S) & 1u;
  const uint32_t i0 = base | (f << j), i1 = base | ((f ^ 1u) << j);
  diag_half<Real, S, 0>(re, im, rec + 4 * i0, rec + 16 + 2 * i0);
  diag_half<Real, S, 1>(re, im, rec + 4 * i1, rec + 16 + 2 * i1);
}

// X on chunk slot S controlled by chunk slot CS (arg = 4 S + CS)
template <typename V>
B200Q_HD void x_c1(int arg, V* re, V* im, uint32_t xm) {
  const int cs = arg & 3;
  const bool cval = ((xm >> cs) & 1u) == 0;
  switch (arg >> 2) {
    case 0: x_chunk_c1_cs<V, 0>(cs, re, im, cval); break;
    case 1: x_chunk_c1_cs<V, 1>(cs, re, im, cval); break;
    case 2: x_chunk_c1_cs<V, 2>(cs, re, im, cval); break;
    default: x_chunk_c1_cs<V, 3>(cs, re, im, cval); break;
  }
}

// One register round of one thread.  LEAN: only the temp-free op codes are compiled in (b200q_program.h).
template <typename Real, bool LEAN>
B200Q_HD void run_round(const b200q_pass_t& P, const b200q_round_t& Rd, const RoundTab& T, int tid, uint64_t cta_base,
                        uint64_t enabled, typename Traits<Real>::chunk* tile, const cx<Real>* pool,
                        const Real* coef, const OpWord* words, Real gscale,
                        typename Traits<Real>::chunk* gstate, uint64_t total_chunks,
                        const b200q_remote_t* remote = nullptr, const uint64_t* dest_tab = nullptr) {
  using V = typename Traits<Real>::V;
  RoundAddr<Real> A = round_addr<Real>(P, T, tid, cta_base);
  if (!A.active) return;
  V re[NE], im[NE];
  const bool padded = int(P.n_bits) != int(P.n_qubits);   // tiny states: index space padded to the register bits
  gather<Real>(A, T, Rd.src_global, (P.layout & B200Q_LAYOUT_SRC_SOA) != 0, tile, gstate, total_chunks, padded, re,
               im);

  Real rho_r = Real(1), rho_i = Real(0);
  bool rho_dirty = false;
  uint32_t xm = 0;   // X relabelling mask over the chunk-level register slots
  const OpWord* W = words + Rd.op_begin;
  const int n = int(Rd.op_end) - int(Rd.op_begin);
  OpWord nx = W[0];
  for (int i = 0; i < n; ++i) {
    const OpWord cur = nx;
    nx = W[i + 1];   // prefetch (the table has a sentinel entry)
    const uint32_t o = cur.x >> 24;
    if (cur.x & (B200Q_OW_CTRL_LOC | B200Q_OW_CTRL_GLOB)) {
      if ((A.lb & cur.y) != cur.y) continue;
      if ((cur.x & B200Q_OW_CTRL_GLOB) && !((enabled >> o) & 1ull)) continue;
    }
    const Real* rec = coef + o * B200Q_COEF_PER_OP;
#define B200Q_KREC(S) (rec + ((xm >> S) & 1u) * B200Q_COEF_PER_FLIP)
#define B200Q_FAST(CV, VAR, S) \
  case B200Q_CODE_MAT1_FAST + 4 * CV + S: mat1_fast<Real, S, VAR>(re, im, B200Q_KREC(S)); break;
#define B200Q_PERSLOT(S)                                                                         \
  case B200Q_CODE_MAT1_HAD + S: had_fast<Real, S>(re, im, ((xm >> S) & 1u) != 0); break;        \
  case B200Q_CODE_MAT1_ROTX + S: {                                                               \
    bool neg;                                                                                    \
    rot_fast<Real, S, true>(re, im, B200Q_KREC(S), neg);                                         \
    if (neg && (cur.x & (B200Q_OW_CTRL_LOC | B200Q_OW_CTRL_GLOB))) { rho_r = -rho_r; rho_i = -rho_i; rho_dirty = true; } \
    break;                                                                                       \
  }                                                                                              \
  case B200Q_CODE_MAT1_ROTY + S: {                                                               \
    bool neg;                                                                                    \
    rot_fast<Real, S, false>(re, im, B200Q_KREC(S), neg);                                        \
    if (neg && (cur.x & (B200Q_OW_CTRL_LOC | B200Q_OW_CTRL_GLOB))) { rho_r = -rho_r; rho_i = -rho_i; rho_dirty = true; } \
    break;                                                                                       \
  }                                                                                              \
  case B200Q_CODE_DIAG_R + S: diag_reg<Real, S>(P, cur, re, im, rec, A.lb, cta_base, xm); break; \
  case B200Q_CODE_LSWAP + S: lane_swap<V, S>(re, im, xm); break;
    const uint32_t code = cur.x & 0xffu;
    if (code >= B200Q_CODE_X_RELABEL) {   // light ops first: their cost is the dispatch itself
      if (code == B200Q_CODE_X_RELABEL) {
        xm ^= 1u << ((cur.x >> 8) & 3u);
      } else if (code == B200Q_CODE_DIAG_T) {
        const uint32_t idx = dsel_value(P, cur, 0, A.lb, cta_base) | (dsel_value(P, cur, 1, A.lb, cta_base) << 1);
        const Real dr = rec[16 + 2 * idx], di = rec[17 + 2 * idx];
        const Real r = rho_r * dr - rho_i * di, im2 = rho_r * di + rho_i * dr;
        rho_r = r; rho_i = im2; rho_dirty = true;
      } else if (code == B200Q_CODE_X_C1) {
        x_c1<V>(int((cur.x >> 8) & 15u), re, im, xm);
      } else if (code == B200Q_CODE_X_LANE) {   // X on the lane bit, chunk-slot controls in arg: swaps the two lanes
        const uint32_t cm = (cur.x >> 8) & 15u;
        x_lane(re, im, cm, cm & ~xm);
      } else if (!LEAN) {
        const b200q_op_t& op = P.ops[o];
        switch (code) {
          B200Q_FAST(0, VAR_REAL, 0) B200Q_FAST(0, VAR_REAL, 1) B200Q_FAST(0, VAR_REAL, 2) B200Q_FAST(0, VAR_REAL, 3)
          B200Q_FAST(1, VAR_RXLIKE, 0) B200Q_FAST(1, VAR_RXLIKE, 1) B200Q_FAST(1, VAR_RXLIKE, 2)
          B200Q_FAST(1, VAR_RXLIKE, 3)
          B200Q_FAST(2, VAR_GENERAL, 0) B200Q_FAST(2, VAR_GENERAL, 1) B200Q_FAST(2, VAR_GENERAL, 2)
          B200Q_FAST(2, VAR_GENERAL, 3)
          case B200Q_CODE_MAT1_SLOW: apply_mat1<Real>(op, re, im, pool + op.pool_off, xm); break;
          case B200Q_CODE_X_SLOW: apply_x<Real>(op, re, im, xm); break;
          case B200Q_CODE_DIAG:
            apply_diag<Real, true>(op, re, im, pool + op.pool_off, diag_tsel(op, cta_base, A.lb), xm, rho_r, rho_i,
                                   rho_dirty);
            break;
          default: break;
        }
      }
    } else {
      switch (code) {
        B200Q_PERSLOT(0) B200Q_PERSLOT(1) B200Q_PERSLOT(2) B200Q_PERSLOT(3)
        default: break;
      }
    }
#undef B200Q_FAST
#undef B200Q_PERSLOT
#undef B200Q_KREC
  }
  if (Rd.dst_global && P.has_scale && gscale != Real(1)) { rho_r *= gscale; rho_i *= gscale; rho_dirty = true; }
  if (rho_dirty) {
    const V pr = vset(rho_r, (V*)nullptr), pi = vset(rho_i, (V*)nullptr), npi = vset(-rho_i, (V*)nullptr);
#pragma unroll
    for (int c = 0; c < NE; ++c) cmul_inplace(re[c], im[c], pr, pi, npi);
  }
  scatter<Real>(A, T, Rd.dst_global, (P.layout & B200Q_LAYOUT_DST_SOA) != 0, tile, gstate, total_chunks, padded, xm,
                re, im, remote, dest_tab);
}

// ------------------------------------------------------------------------------------------------
// dense k-target op applied in place in the shared-memory tile (k = 2..4)
// ------------------------------------------------------------------------------------------------
// Amplitude `loc` of the (SoA) tile as separate re / im scalars.
template <typename Real> struct TileAmp { Real* re; Real* im; };
B200Q_HD TileAmp<float> tile_amp(chunk_f* tile, uint32_t loc) {
  float* base = reinterpret_cast<float*>(tile + swz(loc >> 1));
  TileAmp<float> a; a.re = base + (loc & 1u); a.im = base + 2 + (loc & 1u);
  return a;
}
B200Q_HD TileAmp<double> tile_amp(chunk_d* tile, uint32_t loc) {
  double* base = reinterpret_cast<double*>(tile + swz(loc));
  TileAmp<double> a; a.re = base; a.im = base + 1;
  return a;
}

template <int K>
B200Q_HD void sort_targets(const b200q_op_t& op, int* srt) {
#pragma unroll
  for (int j = 0; j < K; ++j) srt[j] = op.tk[j];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = a + 1; b < K; ++b)
      if (srt[b] < srt[a]) { const int t = srt[a]; srt[a] = srt[b]; srt[b] = t; }
}

template <typename Real, int K>
B200Q_HD void run_matk(const b200q_pass_t& P, const b200q_op_t& op, int tid, int nthreads, uint64_t cta_base,
                       typename Traits<Real>::chunk* tile, const cx<Real>* pool) {
  constexpr int D = 1 << K;
  if ((cta_base & op.ctrl_glob) != op.ctrl_glob) return;
  int srt[K];
  sort_targets<K>(op, srt);
  const cx<Real>* m = pool + op.pool_off;
  const int ngroups = 1 << (int(P.tile_bits) - K);
  for (int g = tid; g < ngroups; g += nthreads) {
    uint32_t base = uint32_t(g);
#pragma unroll
    for (int j = 0; j < K; ++j) base = ((base >> srt[j]) << (srt[j] + 1)) | (base & ((1u << srt[j]) - 1u));
    if ((base & op.ctrl_loc) != op.ctrl_loc) continue;
    Real xr[D], xi[D];
    TileAmp<Real> ptr[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      uint32_t off = base;
#pragma unroll
      for (int j = 0; j < K; ++j) off |= ((uint32_t(i) >> j) & 1u) << op.tk[j];
      ptr[i] = tile_amp(tile, off);
      xr[i] = *ptr[i].re; xi[i] = *ptr[i].im;
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      Real yr = m[r * D].x * xr[0] - m[r * D].y * xi[0];
      Real yi = m[r * D].x * xi[0] + m[r * D].y * xr[0];
#pragma unroll
      for (int c = 1; c < D; ++c) {
        const cx<Real> w = m[r * D + c];
        yr += w.x * xr[c] - w.y * xi[c];
        yi += w.x * xi[c] + w.y * xr[c];
      }
      *ptr[r].re = yr; *ptr[r].im = yi;
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

// ================================================================================================
// Adjoint (reverse) sweep: psi <- U^dagger psi, G += lambda (x) conj(psi), lambda <- U^dagger lambda
// for the ops of a pass taken in reverse (reference adjoint.py:47-83, generalised from one scalar
// parameter to the full cotangent of every gate matrix so that PyTorch autograd chains it to any
// parametrisation).  PyTorch convention: grad_M[r][c] = sum_rest lambda_out[r,rest] conj(psi_in[c,rest]).
// ================================================================================================
#define B200Q_ACC_PER_OP 32 /* doubles of gradient accumulator per op (2x2: 8 used, 4x4: 32 used) */

// acc[2*(2r+c)], acc[2*(2r+c)+1] += Re, Im of sum lambda[r] conj(psi[c]) over the controlled pairs
template <typename V, int S>
B200Q_HD void accum_mat1_chunk(const V* pr, const V* pi, const V* lr, const V* li, uint32_t cm, uint32_t cv,
                               bool lane_ctrl, double* acc) {
  V a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = vzero((V*)nullptr);
#pragma unroll
  for (int c = 0; c < NE; ++c) {
    if (c & (1 << S)) continue;
    if ((uint32_t(c) & cm) != cv) continue;
    const int d = c | (1 << S);
    // G[r][q] += l_r * conj(p_q):  re = lr*pr + li*pi ; im = li*pr - lr*pi
    a[0] = vfma(li[c], pi[c], vfma(lr[c], pr[c], a[0])); a[1] = vfma(vneg(lr[c]), pi[c], vfma(li[c], pr[c], a[1]));
    a[2] = vfma(li[c], pi[d], vfma(lr[c], pr[d], a[2])); a[3] = vfma(vneg(lr[c]), pi[d], vfma(li[c], pr[d], a[3]));
    a[4] = vfma(li[d], pi[c], vfma(lr[d], pr[c], a[4])); a[5] = vfma(vneg(lr[d]), pi[c], vfma(li[d], pr[c], a[5]));
    a[6] = vfma(li[d], pi[d], vfma(lr[d], pr[d], a[6])); a[7] = vfma(vneg(lr[d]), pi[d], vfma(li[d], pr[d], a[7]));
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = lane_ctrl ? vlane1(a[k]) : vhsum(a[k]);
}

B200Q_HD void accum_mat1_lane(const pk* pr, const pk* pi, const pk* lr, const pk* li, uint32_t cm, uint32_t cv,
                              double* acc) {
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < NE; ++c) {
    if ((uint32_t(c) & cm) != cv) continue;
    const float p0r = pk_x(pr[c]), p0i = pk_x(pi[c]), p1r = pk_y(pr[c]), p1i = pk_y(pi[c]);
    const float l0r = pk_x(lr[c]), l0i = pk_x(li[c]), l1r = pk_y(lr[c]), l1i = pk_y(li[c]);
    a[0] += l0r * p0r + l0i * p0i; a[1] += l0i * p0r - l0r * p0i;
    a[2] += l0r * p1r + l0i * p1i; a[3] += l0i * p1r - l0r * p1i;
    a[4] += l1r * p0r + l1i * p0i; a[5] += l1i * p0r - l1r * p0i;
    a[6] += l1r * p1r + l1i * p1i; a[7] += l1i * p1r - l1r * p1i;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = double(a[k]);
}
B200Q_HD void accum_mat1_lane(const double*, const double*, const double*, const double*, uint32_t, uint32_t,
                              double*) {}

template <typename Real>
B200Q_HD void accum_mat1(const b200q_op_t& op, uint32_t xm, const typename Traits<Real>::V* pr,
                         const typename Traits<Real>::V* pi, const typename Traits<Real>::V* lr,
                         const typename Traits<Real>::V* li, double* acc) {
  using V = typename Traits<Real>::V;
  constexpr int VS = Traits<Real>::VS;
  const uint32_t cm = op.ctrl_reg >> VS, cv = cm & ~xm;
  const bool lane_ctrl = VS && (op.ctrl_reg & 1u);
  if (VS && op.slot == 0) { accum_mat1_lane(pr, pi, lr, li, cm, cv, acc); return; }
  const int s = int(op.slot) - VS;
  switch (s) {
    case 0: accum_mat1_chunk<V, 0>(pr, pi, lr, li, cm, cv, lane_ctrl, acc); break;
    case 1: accum_mat1_chunk<V, 1>(pr, pi, lr, li, cm, cv, lane_ctrl, acc); break;
    case 2: accum_mat1_chunk<V, 2>(pr, pi, lr, li, cm, cv, lane_ctrl, acc); break;
    default: accum_mat1_chunk<V, 3>(pr, pi, lr, li, cm, cv, lane_ctrl, acc); break;
  }
  if ((xm >> s) & 1u) {   // the register halves hold logical 1 / 0: entry (r, c) <-> (r^1, c^1)
    double t;
    t = acc[0]; acc[0] = acc[6]; acc[6] = t;
    t = acc[1]; acc[1] = acc[7]; acc[7] = t;
    t = acc[2]; acc[2] = acc[4]; acc[4] = t;
    t = acc[3]; acc[3] = acc[5]; acc[5] = t;
  }
}

// acc[2*idx], acc[2*idx+1] += Re, Im of sum lambda conj(psi) over the amplitudes whose selector value is idx
template <typename Real>
B200Q_HD void accum_diag(const b200q_op_t& op, uint32_t tsel, uint32_t xm, const typename Traits<Real>::V* pr,
                         const typename Traits<Real>::V* pi, const typename Traits<Real>::V* lr,
                         const typename Traits<Real>::V* li, double* acc) {
  constexpr int VS = Traits<Real>::VS;
  constexpr int NL = 1 << VS;
  Real a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < NE; ++c) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const uint32_t i = ((uint32_t(c) ^ xm) << VS) | uint32_t(l);   // logical register index
      if ((i & op.ctrl_reg) != op.ctrl_reg) continue;
      uint32_t idx = tsel;
      if (op.dsel_slot[0] != 0xff) idx |= (i >> op.dsel_slot[0]) & 1u;
      if (op.dsel_slot[1] != 0xff) idx |= ((i >> op.dsel_slot[1]) & 1u) << 1;
      const Real plr = Real(vget(pr[c], l)), pli = Real(vget(pi[c], l));
      const Real llr = Real(vget(lr[c], l)), lli = Real(vget(li[c], l));
      const Real gr = llr * plr + lli * pli;
      const Real gi = lli * plr - llr * pli;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a[2 * k] += idx == uint32_t(k) ? gr : Real(0);
        a[2 * k + 1] += idx == uint32_t(k) ? gi : Real(0);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = double(a[k]);
}

// Accumulator hook: on the device the per-thread partials are warp-reduced and added to the CTA's
// shared accumulator; the host emulator adds directly.
#if defined(__CUDA_ARCH__)
template <int NV>
__device__ __forceinline__ void cta_accumulate(double* cta_acc, const double* v) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(cta_acc + k, s);
  }
}
#else
template <int NV>
inline void cta_accumulate(double* cta_acc, const double* v) {
  for (int k = 0; k < NV; ++k) cta_acc[k] += v[k];
}
#endif

// Per-warp accumulators (one slice of shared memory per warp): warp-reduce, lane 0 adds WITHOUT an atomic
// (shared-memory double atomics are compare-and-swap loops: they were 11 % of the reverse sweep's samples).
// `warp_accumulate_idx`: only entry pair `idx` of the thread is non-zero (thread-level diagonal); when the
// whole warp agrees on idx (selectors outside the lane bits, 2/3 of the cases) two reductions replace eight.
#if defined(__CUDA_ARCH__)
template <int NV>
__device__ __forceinline__ void warp_accumulate(double* warp_acc, const double* v) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_acc[k] += s;
  }
}
// Eight sums at once: a transposing reduction -- at xor distances 16, 8, 4 every lane hands half of its values to
// its partner and keeps the other half, then the one remaining value is reduced over distances 2 and 1: 9
// 64-bit shuffles and adds instead of 40; lane 4*k afterwards holds the total of value k.
template <>
__device__ __forceinline__ void warp_accumulate<8>(double* warp_acc, const double* v) {
  const int lane = threadIdx.x & 31;
  double b[4], c[2], d;
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double send = hi ? v[j] : v[j + 4], keep = hi ? v[j + 4] : v[j];
      b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = (lane & 8) != 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double send = hi ? b[j] : b[j + 2], keep = hi ? b[j + 2] : b[j];
      c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = (lane & 4) != 0;
    const double send = hi ? c[0] : c[1], keep = hi ? c[1] : c[0];
    d = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  if ((lane & 3) == 0) warp_acc[((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] += d;
}
__device__ __forceinline__ void warp_accumulate_idx(double* warp_acc, uint32_t idx, bool on, double sr, double si) {
  int uniform;
  __match_all_sync(0xffffffffu, on ? idx : 0xffu, &uniform);
  if (uniform) {
    if (!on) return;
    double v[2] = {sr, si};
    warp_accumulate<2>(warp_acc + 2 * idx, v);
  } else {
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (on) { v[2 * idx] = sr; v[2 * idx + 1] = si; }
    warp_accumulate<8>(warp_acc, v);
  }
}
#else
template <int NV>
inline void warp_accumulate(double* warp_acc, const double* v) {
  for (int k = 0; k < NV; ++k) warp_acc[k] += v[k];
}
inline void warp_accumulate_idx(double* warp_acc, uint32_t idx, bool on, double sr, double si) {
  if (on) { warp_acc[2 * idx] += sr; warp_acc[2 * idx + 1] += si; }
}
#endif
#define B200Q_WACC_PER_OP 8

// ---- reverse sweep on the lean op set ---------------------------------------------------------------------
// U^dagger is applied with the same in-place ops as the forward kernel (records built with flip_adjoint).
// Two things may stay PENDING on the pair (psi, lambda), because the cotangent G = lambda (x) conj(psi) of the
// other ops does not care or can be fixed afterwards:
//   * a common unit-modulus phase `rho` per thread (thread-level diagonals, signs of controlled rotations):
//     both states carry it, it cancels in lambda * conj(psi); applied once per round;
//   * the real scalar x of every Hadamard-structured op (applied as the bare add/sub butterfly): op o then sees
//     psi and lambda both short of  prod x_j  over the Hadamard ops j > o of the pass -- its G is multiplied by
//     gfac[o] = prod x_j^2 at flush time, the states by the full product in the last (global) round.
// Hadamard ops whose own gradient is requested take the general (scaled) path and do not count.

// gfac[o] for every op of the pass; returns the scalar the states are short of at the end of the pass
template <typename Real>
B200Q_HD double adjoint_scales(const b200q_pass_t& P, const cx<Real>* mats, uint64_t want_mask, double* gfac) {
  double g = 1.0, sign = 1.0;
  for (int o = int(P.n_ops) - 1; o >= 0; --o) {
    gfac[o] = g * g;
    const b200q_op_t& op = P.ops[o];
    if (code_is_had(op.code) && !((want_mask >> o) & 1ull)) g *= double(mats[op.mat_src].x);
    else if (code_is_rot(op.code) && !op.tctrl && mats[op.mat_src].x < Real(0)) sign = -sign;
  }
  return g * sign;
}

// (sr, si) = sum over the elements e with ((e >> S) & 1) == H (S < 0: all) of lambda_e * conj(psi_e)
template <typename V, int S, int H>
B200Q_HD void dot_lam_conj_psi(const V* pr, const V* pi, const V* lr, const V* li, double& sr, double& si) {
  V ar = vzero((V*)nullptr), ai = vzero((V*)nullptr);
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    if (S >= 0 && ((e >> (S < 0 ? 0 : S)) & 1) != H) continue;
    ar = vfma(li[e], pi[e], vfma(lr[e], pr[e], ar));
    ai = vfma(vneg(lr[e]), pi[e], vfma(li[e], pr[e], ai));
  }
  sr = vhsum(ar); si = vhsum(ai);
}

template <typename Real, int S>
B200Q_HD void diag_reg_adjoint(const b200q_pass_t& P, const OpWord& w, typename Traits<Real>::V* pr,
                               typename Traits<Real>::V* pi, typename Traits<Real>::V* lr,
                               typename Traits<Real>::V* li, const Real* rec, uint32_t lb, uint64_t cta_base,
                               uint32_t xm, bool want, double* acc) {
  using V = typename Traits<Real>::V;
  const int j = (w.x >> 8) & 1, other = j ^ 1;
  const uint32_t base = dsel_value(P, w, other, lb, cta_base) << other;
  const uint32_t f = (xm >> S) & 1u;
  const uint32_t i0 = base | (f << j), i1 = base | ((f ^ 1u) << j);
  diag_half<Real, S, 0>(pr, pi, rec + 4 * i0, rec + 16 + 2 * i0);
  diag_half<Real, S, 1>(pr, pi, rec + 4 * i1, rec + 16 + 2 * i1);
  if (want) {   // G[idx] = sum lambda_out conj(psi_in) over the half with selector value idx
    double sr, si;
    dot_lam_conj_psi<V, S, 0>(pr, pi, lr, li, sr, si);
    acc[2 * i0] += sr; acc[2 * i0 + 1] += si;
    dot_lam_conj_psi<V, S, 1>(pr, pi, lr, li, sr, si);
    acc[2 * i1] += sr; acc[2 * i1 + 1] += si;
  }
  diag_half<Real, S, 0>(lr, li, rec + 4 * i0, rec + 16 + 2 * i0);
  diag_half<Real, S, 1>(lr, li, rec + 4 * i1, rec + 16 + 2 * i1);
}

// One register round of the reverse sweep.  The roles of src/dst (and of the pass layouts) are
// swapped with respect to the forward round.  Every thread of the CTA must call this (the gradient
// reduction inside is warp-collective).  `coef`: records of U^dagger (fill_coefs with flip_adjoint);
// `gscale`: adjoint_scales(), applied by the round that writes back to global memory.
template <typename Real>
B200Q_HD void run_round_adjoint(const b200q_pass_t& P, const b200q_round_t& Rd, const RoundTab& T, int tid,
                                uint64_t cta_base,
                                typename Traits<Real>::chunk* tile_psi, typename Traits<Real>::chunk* tile_lam,
                                const cx<Real>* pool, const Real* coef, const OpWord* words, Real gscale,
                                typename Traits<Real>::chunk* gpsi,
                                typename Traits<Real>::chunk* glam, uint64_t total_chunks, uint64_t want_mask,
                                double* wacc_base) {
  using V = typename Traits<Real>::V;
  RoundAddr<Real> A = round_addr<Real>(P, T, tid, cta_base);
  double* wacc = wacc_base + size_t(tid >> 5) * (B200Q_MAX_OPS * B200Q_WACC_PER_OP);   // this warp's slice
  V pr[NE], pi[NE], lr[NE], li[NE];
  uint32_t xm = 0;
  constexpr int VS_ = Traits<Real>::VS;
  const bool soa_in = (P.layout & B200Q_LAYOUT_DST_SOA) != 0, soa_out = (P.layout & B200Q_LAYOUT_SRC_SOA) != 0;
  if (A.active) {
    gather<Real>(A, T, Rd.dst_global, soa_in, tile_psi, gpsi, total_chunks, true, pr, pi);
    gather<Real>(A, T, Rd.dst_global, soa_in, tile_lam, glam, total_chunks, true, lr, li);
  } else {
#pragma unroll
    for (int c = 0; c < NE; ++c) {
      pr[c] = pi[c] = lr[c] = li[c] = vzero((V*)nullptr);
    }
  }
  Real dummy_r = Real(1), dummy_i = Real(0);
  bool dummy_d = false;
  Real rho_r = Real(1), rho_i = Real(0);   // pending common phase of (psi, lambda)
  bool rho_dirty = false;
  for (int o = int(Rd.op_end) - 1; o >= int(Rd.op_begin); --o) {
    const b200q_op_t& op = P.ops[o];
    if ((cta_base & op.ctrl_glob) != op.ctrl_glob) continue;   // CTA-uniform
    const bool on = A.active && ((A.lb & op.ctrl_loc) == op.ctrl_loc);
    const bool want = (want_mask >> o) & 1ull;
    const cx<Real>* w = pool + op.pool_off;
    const OpWord cur = words[o];
    const Real* rec = coef + o * B200Q_COEF_PER_OP;
    const uint32_t code = op.code;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define B200Q_KREC(S) (rec + ((xm >> S) & 1u) * B200Q_COEF_PER_FLIP)
#define B200Q_ADJ_ROT(S, ISX)                                                                  \
  {                                                                                            \
    if (on) {                                                                                  \
      bool neg;                                                                                \
      rot_fast<Real, S, ISX>(pr, pi, B200Q_KREC(S), neg);                                      \
      if (want) accum_mat1<Real>(op, xm, pr, pi, lr, li, acc);                                 \
      if (want && neg) {   /* psi already carries the sign of the negated matrix, lambda not yet */ \
        for (int k = 0; k < 8; ++k) acc[k] = -acc[k];                                          \
      }                                                                                        \
      rot_fast<Real, S, ISX>(lr, li, B200Q_KREC(S), neg);                                      \
      if (neg && op.tctrl) { rho_r = -rho_r; rho_i = -rho_i; rho_dirty = true; }               \
    }                                                                                          \
    if (want) warp_accumulate<8>(wacc + o * B200Q_WACC_PER_OP, acc);                          \
  }
#define B200Q_ADJ_PERSLOT(S)                                                                   \
  case B200Q_CODE_MAT1_HAD + S:                                                                \
    if (on) { had_fast<Real, S>(pr, pi, ((xm >> S) & 1u) != 0); had_fast<Real, S>(lr, li, ((xm >> S) & 1u) != 0); } \
    break;                                                                                     \
  case B200Q_CODE_MAT1_ROTX + S: B200Q_ADJ_ROT(S, true) break;                                 \
  case B200Q_CODE_MAT1_ROTY + S: B200Q_ADJ_ROT(S, false) break;                                \
  case B200Q_CODE_DIAG_R + S:                                                                  \
    if (on) diag_reg_adjoint<Real, S>(P, cur, pr, pi, lr, li, rec, A.lb, cta_base, xm, want, acc); \
    if (want) warp_accumulate<8>(wacc + o * B200Q_WACC_PER_OP, acc);                          \
    break;                                                                                     \
  case B200Q_CODE_LSWAP + S:                                                                   \
    if (A.active) { uint32_t xm2 = xm; lane_swap<V, S>(pr, pi, xm); lane_swap<V, S>(lr, li, xm2); } \
    break;
    const bool lean_had = code_is_had(int(code)) && !want;
    if (code < B200Q_CODE_X_RELABEL && (lean_had || !code_is_had(int(code)))) {
      switch (code) {
        B200Q_ADJ_PERSLOT(0) B200Q_ADJ_PERSLOT(1) B200Q_ADJ_PERSLOT(2) B200Q_ADJ_PERSLOT(3)
        default: break;
      }
      continue;
    }
#undef B200Q_ADJ_PERSLOT
#undef B200Q_ADJ_ROT
#undef B200Q_KREC
    if (code == B200Q_CODE_X_RELABEL) {
      if (on) xm ^= 1u << ((cur.x >> 8) & 3u);
      continue;
    }
    if (code == B200Q_CODE_X_C1) {
      if (on) { x_c1<V>(int((cur.x >> 8) & 15u), pr, pi, xm); x_c1<V>(int((cur.x >> 8) & 15u), lr, li, xm); }
      continue;
    }
    if (code == B200Q_CODE_X_LANE) {
      if (on) { const uint32_t cm = (cur.x >> 8) & 15u; x_lane(pr, pi, cm, cm & ~xm); x_lane(lr, li, cm, cm & ~xm); }
      continue;
    }
    if (code == B200Q_CODE_DIAG_T) {
      const uint32_t idx = dsel_value(P, cur, 0, A.lb, cta_base) | (dsel_value(P, cur, 1, A.lb, cta_base) << 1);
      double gr = 0.0, gi = 0.0;
      if (on) {
        const Real dr = rec[16 + 2 * idx], di = rec[17 + 2 * idx];   // entry of U^dagger = conj(d)
        if (want) {   // G[idx] = d * sum lambda conj(psi)  (the pending phases cancel)
          double sr, si;
          dot_lam_conj_psi<V, -1, 0>(pr, pi, lr, li, sr, si);
          gr = double(dr) * sr + double(di) * si;
          gi = double(dr) * si - double(di) * sr;
        }
        const Real r = rho_r * dr - rho_i * di, im2 = rho_r * di + rho_i * dr;
        rho_r = r; rho_i = im2; rho_dirty = true;
      }
      if (want) warp_accumulate_idx(wacc + o * B200Q_WACC_PER_OP, idx, on, gr, gi);
      continue;
    }
    switch (op.kind) {
      case B200Q_OP_MAT1:
        if (on) {
          apply_mat1<Real>(op, pr, pi, w, xm);
          if (want) accum_mat1<Real>(op, xm, pr, pi, lr, li, acc);
          apply_mat1<Real>(op, lr, li, w, xm);
        }
        if (want) warp_accumulate<8>(wacc + o * B200Q_WACC_PER_OP, acc);
        break;
      case B200Q_OP_X:
        if (on) { uint32_t xm2 = xm; apply_x<Real>(op, pr, pi, xm); apply_x<Real>(op, lr, li, xm2); }
        break;
      case B200Q_OP_LSWAP:
        if (A.active) {
          uint32_t xm2 = xm;
          switch (int(op.slot) - VS_) {
            case 0: lane_swap<V, 0>(pr, pi, xm); lane_swap<V, 0>(lr, li, xm2); break;
            case 1: lane_swap<V, 1>(pr, pi, xm); lane_swap<V, 1>(lr, li, xm2); break;
            case 2: lane_swap<V, 2>(pr, pi, xm); lane_swap<V, 2>(lr, li, xm2); break;
            default: lane_swap<V, 3>(pr, pi, xm); lane_swap<V, 3>(lr, li, xm2); break;
          }
        }
        break;
      case B200Q_OP_DIAG: {
        const uint32_t tsel = diag_tsel(op, cta_base, A.lb);
        if (on) {
          apply_diag<Real, false>(op, pr, pi, w, tsel, xm, dummy_r, dummy_i, dummy_d);
          if (want) accum_diag<Real>(op, tsel, xm, pr, pi, lr, li, acc);
          apply_diag<Real, false>(op, lr, li, w, tsel, xm, dummy_r, dummy_i, dummy_d);
        }
        if (want) warp_accumulate<8>(wacc + o * B200Q_WACC_PER_OP, acc);
        break;
      }
      default: break;
    }
  }
  if (!A.active) return;
  if (Rd.src_global && gscale != Real(1)) { rho_r *= gscale; rho_i *= gscale; rho_dirty = true; }
  if (rho_dirty) {
    const V qr = vset(rho_r, (V*)nullptr), qi = vset(rho_i, (V*)nullptr), nqi = vset(-rho_i, (V*)nullptr);
#pragma unroll
    for (int c = 0; c < NE; ++c) { cmul_inplace(pr[c], pi[c], qr, qi, nqi); cmul_inplace(lr[c], li[c], qr, qi, nqi); }
  }
  scatter<Real>(A, T, Rd.src_global, soa_out, tile_psi, gpsi, total_chunks, true, xm, pr, pi);
  scatter<Real>(A, T, Rd.src_global, soa_out, tile_lam, glam, total_chunks, true, xm, lr, li);
}

// Dense k-target op of the reverse sweep, in place in both shared-memory tiles.  Gradient
// accumulation is provided for k = 2 (the parametric two-qubit gates); k >= 3 gates are applied but
// must not require a gradient (checked on the host).
template <typename Real, int K>
B200Q_HD void run_matk_adjoint(const b200q_pass_t& P, const b200q_op_t& op, int tid, int nthreads, uint64_t cta_base,
                               typename Traits<Real>::chunk* tile_psi, typename Traits<Real>::chunk* tile_lam,
                               const cx<Real>* pool, bool want, double* cta_acc_op) {
  constexpr int D = 1 << K;
  constexpr int NACC = (K == 2) ? 32 : 1;
  if ((cta_base & op.ctrl_glob) != op.ctrl_glob) return;
  int srt[K];
  sort_targets<K>(op, srt);
  const cx<Real>* w = pool + op.pool_off;
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  const int ngroups = 1 << (int(P.tile_bits) - K);
  for (int g = tid; g < ngroups; g += nthreads) {
    uint32_t base = uint32_t(g);
#pragma unroll
    for (int j = 0; j < K; ++j) base = ((base >> srt[j]) << (srt[j] + 1)) | (base & ((1u << srt[j]) - 1u));
    if ((base & op.ctrl_loc) != op.ctrl_loc) continue;
    Real xr[D], xi[D], lr[D], li[D], yr[D], yi[D];
    TileAmp<Real> pp[D], lp[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      uint32_t off = base;
#pragma unroll
      for (int j = 0; j < K; ++j) off |= ((uint32_t(i) >> j) & 1u) << op.tk[j];
      pp[i] = tile_amp(tile_psi, off);
      lp[i] = tile_amp(tile_lam, off);
      xr[i] = *pp[i].re; xi[i] = *pp[i].im;
      lr[i] = *lp[i].re; li[i] = *lp[i].im;
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      Real sr = Real(0), si = Real(0);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const cx<Real> m = w[r * D + c];
        sr += m.x * xr[c] - m.y * xi[c];
        si += m.x * xi[c] + m.y * xr[c];
      }
      yr[r] = sr; yi[r] = si;
      *pp[r].re = sr; *pp[r].im = si;
    }
    if (K == 2 && want) {
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
          acc[(2 * (r * D + c)) % NACC] += double(lr[r] * yr[c] + li[r] * yi[c]);
          acc[(2 * (r * D + c) + 1) % NACC] += double(li[r] * yr[c] - lr[r] * yi[c]);
        }
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      Real sr = Real(0), si = Real(0);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const cx<Real> m = w[r * D + c];
        sr += m.x * lr[c] - m.y * li[c];
        si += m.x * li[c] + m.y * lr[c];
      }
      *lp[r].re = sr; *lp[r].im = si;
    }
  }
  if (K == 2 && want) cta_accumulate<NACC>(cta_acc_op, acc);
}

template <typename Real>
B200Q_HD void run_direct_op_adjoint(const b200q_pass_t& P, const b200q_op_t& op, int tid, int nthreads,
                                    uint64_t cta_base, typename Traits<Real>::chunk* tile_psi,
                                    typename Traits<Real>::chunk* tile_lam, const cx<Real>* pool, bool want,
                                    double* cta_acc_op) {
  switch (op.k) {
    case 2: run_matk_adjoint<Real, 2>(P, op, tid, nthreads, cta_base, tile_psi, tile_lam, pool, want, cta_acc_op); break;
    case 3: run_matk_adjoint<Real, 3>(P, op, tid, nthreads, cta_base, tile_psi, tile_lam, pool, false, cta_acc_op); break;
    case 4: run_matk_adjoint<Real, 4>(P, op, tid, nthreads, cta_base, tile_psi, tile_lam, pool, false, cta_acc_op); break;
    default: break;
  }
}

// Add the per-warp accumulators of the register rounds into the CTA accumulator (entries 0..7 of every op).
B200Q_HD void merge_warp_acc(const b200q_pass_t& P, int tid, int nthreads, int nwarps, const double* wacc,
                             double* cta_acc) {
  for (int e = tid; e < int(P.n_ops) * B200Q_WACC_PER_OP; e += nthreads) {
    const int o = e / B200Q_WACC_PER_OP, k = e % B200Q_WACC_PER_OP;
    double v = 0.0;
    for (int w = 0; w < nwarps; ++w) v += wacc[size_t(w) * (B200Q_MAX_OPS * B200Q_WACC_PER_OP) + e];
    cta_acc[o * B200Q_ACC_PER_OP + k] += v;
  }
}

// Flush one CTA's accumulators into the global gradient buffer (complex128, laid out like the
// matrix buffer).  `add(ptr, v)` is atomicAdd on the device.
template <typename AddFn>
B200Q_HD void flush_grad(const b200q_pass_t& P, int tid, int nthreads, uint64_t want_mask, const double* cta_acc,
                         const double* gfac, double* grad, AddFn add) {
  for (int e = tid; e < int(P.n_ops) * B200Q_ACC_PER_OP; e += nthreads) {
    const int o = e / B200Q_ACC_PER_OP, q = e % B200Q_ACC_PER_OP;
    if (!((want_mask >> o) & 1ull)) continue;
    const b200q_op_t& op = P.ops[o];
    const int comp = q & 1, ent = q >> 1;
    const bool adj = (op.flags & B200Q_FLAG_ADJOINT) != 0;
    int dst;
    if (op.kind == B200Q_OP_MAT1) {
      if (ent >= 4) continue;
      const int r = ent >> 1, c = ent & 1;
      dst = adj ? c * 2 + r : r * 2 + c;
    } else if (op.kind == B200Q_OP_DIAG) {
      const int dim = 1 << op.k;
      if (ent >= dim) continue;
      dst = ent * (dim + 1);
    } else if (op.kind == B200Q_OP_MATK && op.k == 2) {
      const int r = ent >> 2, c = ent & 3;
      dst = adj ? c * 4 + r : r * 4 + c;
    } else {
      continue;
    }
    double v = cta_acc[e] * gfac[o];
    if (adj && comp == 1) v = -v;   // grad wrt M where U = M^dagger: conjugate (and transpose above)
    if (v != 0.0) add(grad + 2 * (uint64_t(op.mat_src) + dst) + comp, v);
  }
}

}  // namespace b200q
