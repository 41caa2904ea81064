// Per-pass kernel specialisation: see b200q_codegen.h.
//
// What the generated kernel does differently from the generic tile kernel (b200q_tile_body.h), op by op:
//   * the 32 data registers of a thread (16 elements x re/im) are NAMED variables w0..w31 in straight-line code, and
//     the generator tracks which variable holds which logical element.  A CNOT whose control and target are both
//     register slots is then a renaming of variables at generation time: zero instructions.  S / S^dagger on a
//     register slot is a renaming (re <-> im) plus one sign flip per element (folded by ptxas into the operand
//     modifier of the consuming FFMA2 / FADD2).
//   * ops whose control is a thread-level or tile-level index bit do not move data either: every register slot
//     carries a per-thread *Pauli frame* (fx, fz: the registers hold  X^fx Z^fz |psi>  up to the thread's phase).
//     X gates and CNOTs with thread-level controls toggle fx; Hadamard-structured ops exchange fx and fz
//     (H X = Z H, H Z = X H, sign (-1)^(fx fz)); Rx / Ry rotations only change the sign of their shear coefficients
//     (Rx(t) Z = Z Rx(-t), Ry(t) X = X Ry(-t), ...); CNOTs between register slots propagate the frame by the
//     Clifford rules (fx_t ^= fx_c, fz_c ^= fz_t); diagonal ops select their entries by fx.  At the end of a round
//     fx becomes part of the scatter address (free) and fz is applied as a sign to half of the registers.
//   * thread-level diagonal ops and the deferred scalars (Hadamard 1/sqrt2 factors, rotation signs) fold into one
//     complex scalar per thread (rr, ri), applied once per round.
//   * coefficients are plain scalars in shared memory (prepared once per CTA): ptxas uses the scalar-broadcast and
//     negation operand forms of FFMA2 (`R.F32`, `-R.F32x2`), so no packing or sign instructions are issued.
//   * all index arithmetic (tile -> physical base, thread -> tile-local offset, element strides, swizzle) is made of
//     generation-time constants.
#include "b200q_codegen.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <sstream>
#include <vector>

namespace b200q {
namespace {

uint32_t swz_host(uint32_t c) { return c ^ (((c >> 3) ^ (c >> 6) ^ (c >> 9) ^ (c >> 12)) & 7u); }

// ---- the fixed part of every generated source --------------------------------------------------------------------
const char* kPreamble = R"SRC(
typedef unsigned int u32;
typedef unsigned long long u64;
typedef long long i64;
#if defined(__CUDACC__)
#define DEV __device__ __forceinline__
#define B200QJ_CONST __constant__ const
#else
#include <cstdlib>
#include <cstring>
#define DEV static inline
#define B200QJ_CONST static const
#endif

#if B200QJ_F32
typedef float Real;
#if defined(__CUDACC__)
// two float lanes in ONE 64-bit register pair: the operand format of the Blackwell packed-FP32 instructions
struct alignas(8) V { u64 u; };
DEV V vmk(float x, float y) { V r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.u) : "f"(x), "f"(y)); return r; }
DEV float vx(V a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.u)); return x; }
DEV float vy(V a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.u)); return y; }
DEV V vfma(V a, V b, V c) { V r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.u) : "l"(a.u), "l"(b.u), "l"(c.u)); return r; }
DEV V vmul(V a, V b) { V r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u)); return r; }
DEV V vadd(V a, V b) { V r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u)); return r; }
DEV V vsub(V a, V b) { V r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u)); return r; }
// In-place updates with tied operands: the value stays in its register, so the (re, im) register quad an element was
// loaded into is still a quad when the element is stored with one 128-bit access (no packing moves).
DEV void vfa(V& x, V k, V y) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x.u) : "l"(k.u), "l"(y.u)); }   // x += k y
DEV void vsc(V& x, V k) { asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(x.u) : "l"(k.u)); }                        // x *= k
DEV void vcm(V& re, V& im, V pr, V pi, V npi) {                                                            // (re, im) *= p
  asm("{\n .reg .b64 t0, t1;\n mul.rn.f32x2 t0, %4, %1;\n mul.rn.f32x2 t1, %3, %0;\n"
      " fma.rn.f32x2 %0, %2, %0, t0;\n fma.rn.f32x2 %1, %2, %1, t1;\n}"
      : "+l"(re.u), "+l"(im.u) : "l"(pr.u), "l"(pi.u), "l"(npi.u));
}
DEV void vng(V& x) {
  asm("{\n .reg .f32 a, b;\n mov.b64 {a, b}, %0;\n neg.f32 a, a;\n neg.f32 b, b;\n mov.b64 %0, {a, b};\n}" : "+l"(x.u));
}
// in-place butterfly (x, y) <- (x + y, x - y) with tied operands: the values stay in their registers, so the
// (re, im) register quads of the elements survive until the 128-bit stores (no packing moves)
DEV void vhad(V& x, V& y) {
  asm("{\n .reg .b64 t;\n mov.b64 t, %0;\n add.rn.f32x2 %0, t, %1;\n sub.rn.f32x2 %1, t, %1;\n}" : "+l"(x.u), "+l"(y.u));
}
#else
struct V { float x, y; };
DEV V vmk(float x, float y) { V r; r.x = x; r.y = y; return r; }
DEV float vx(V a) { return a.x; }
DEV float vy(V a) { return a.y; }
DEV V vfma(V a, V b, V c) { return vmk(a.x * b.x + c.x, a.y * b.y + c.y); }
DEV V vmul(V a, V b) { return vmk(a.x * b.x, a.y * b.y); }
DEV V vadd(V a, V b) { return vmk(a.x + b.x, a.y + b.y); }
DEV V vsub(V a, V b) { return vmk(a.x - b.x, a.y - b.y); }
DEV void vhad(V& x, V& y) { const V t = x; x = vadd(t, y); y = vsub(t, y); }
DEV void vfa(V& x, V k, V y) { x = vfma(k, y, x); }
DEV void vsc(V& x, V k) { x = vmul(x, k); }
DEV void vcm(V& re, V& im, V pr, V pi, V npi) {
  const V t0 = vmul(npi, im), t1 = vmul(pi, re);
  re = vfma(pr, re, t0); im = vfma(pr, im, t1);
}
DEV void vng(V& x) { x = vmk(-x.x, -x.y); }
#endif
DEV V vbc(float s) { return vmk(s, s); }
DEV V vneg(V a) { return vmk(-vx(a), -vy(a)); }   // ptxas folds it into the operand modifier of the consumer
struct alignas(16) chunk { V lo, hi; };
#else
typedef double Real;
typedef double V;
DEV V vfma(V a, V b, V c) { return a * b + c; }
DEV V vmul(V a, V b) { return a * b; }
DEV V vadd(V a, V b) { return a + b; }
DEV V vsub(V a, V b) { return a - b; }
DEV V vbc(double s) { return s; }
DEV V vneg(V a) { return -a; }
DEV void vhad(V& x, V& y) { const V t = x; x = t + y; y = t - y; }
DEV void vfa(V& x, V k, V y) { x = k * y + x; }
DEV void vsc(V& x, V k) { x *= k; }
DEV void vcm(V& re, V& im, V pr, V pi, V npi) {
  const V t0 = npi * im, t1 = pi * re;
  re = pr * re + t0; im = pr * im + t1;
}
DEV void vng(V& x) { x = -x; }
struct alignas(16) chunk { double lo, hi; };
#endif
struct cplx { Real x, y; };
struct alignas(8) Remote {   // mirrors b200q_remote_t (b200q_program.h)
  void* peer[8];
  u64 base;
  unsigned char perm[40];
  int n_chunk_bits;
  int enabled;
};
#if defined(__CUDACC__)
DEV void b200qj_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
DEV u32 b200qj_swz(u32 c) { return c ^ (((c >> 3) ^ (c >> 6) ^ (c >> 9) ^ (c >> 12)) & 7u); }
// amplitude `loc` of the (SoA) shared-memory tile: pointer to its real part; imaginary part at +B200QJ_IMOFF
#if B200QJ_F32
#define B200QJ_IMOFF 2
DEV Real* b200qj_amp(chunk* tile, u32 loc) { return reinterpret_cast<Real*>(tile + b200qj_swz(loc >> 1)) + (loc & 1u); }
#else
#define B200QJ_IMOFF 1
DEV Real* b200qj_amp(chunk* tile, u32 loc) { return reinterpret_cast<Real*>(tile + b200qj_swz(loc)); }
#endif
DEV cplx b200qj_m1(const cplx* m, int idx, int adj) {
  if (adj && (idx == 1 || idx == 2)) idx ^= 3;
  cplx v = m[idx];
  if (adj) v.y = -v.y;
  return v;
}
)SRC";

// Coefficient preparation, once per CTA: DESC holds (type, matrix offset, flags, coefficient offset) per op.
//   type 1 / 2: Rx- / Ry-structured rotation as three shears -> [u, v, u_x, v_x, neg, 0]; (u_x, v_x): the same for
//               X M X (frame bit fx set); neg = 1 if the matrix was negated to make its diagonal positive
//   type 3: dense 2x2 -> m00r m00i m01r m01i m10r m10i m11r m11i          (adjoint folded in)
//   type 4: diagonal, k selectors -> d0r d0i ... d3r d3i                      (missing entries 1)
//   type 5: dense 2^k x 2^k for the in-tile contraction -> row-major (re, im) pairs
//   type 6: Hadamard-structured op: only its scalar m00 goes into the pass scale
// flags: bit 0 adjoint, bit 1 "sign goes to the pass scale" (un-controlled rotation), bits 8.. k
const char* kPrep = R"SRC(
DEV void b200qj_prep(int tid, Real* coef, const cplx* mats) {
  for (int o = tid; o < B200QJ_NDESC; o += B200QJ_NT) {
    const u32 ty = DESC[4 * o], src = DESC[4 * o + 1], fl = DESC[4 * o + 2], off = DESC[4 * o + 3];
    const int adj = int(fl & 1u), k = int((fl >> 8) & 7u);
    const cplx* m = mats + src;
    Real* c = coef + off;
    if (ty == 1u || ty == 2u) {
      const cplx m00 = b200qj_m1(m, 0, adj), m01 = b200qj_m1(m, 1, adj), m10 = b200qj_m1(m, 2, adj);
      const Real sg = m00.x < Real(0) ? Real(-1) : Real(1);
      const Real cc = sg * m00.x;
      const Real e01 = sg * (ty == 1u ? m01.y : m01.x), e10 = sg * (ty == 1u ? m10.y : m10.x);
      c[0] = e01 / (Real(1) + cc); c[1] = e10;
      c[2] = e10 / (Real(1) + cc); c[3] = e01;
      c[4] = sg < Real(0) ? Real(1) : Real(0); c[5] = Real(0);
    } else if (ty == 3u) {
      for (int i = 0; i < 4; ++i) { const cplx v = b200qj_m1(m, i, adj); c[2 * i] = v.x; c[2 * i + 1] = v.y; }
    } else if (ty == 4u) {
      const int dim = 1 << k;
      for (int i = 0; i < 4; ++i) {
        cplx v; v.x = Real(1); v.y = Real(0);
        if (i < dim) { v = m[i * (dim + 1)]; if (adj) v.y = -v.y; }
        c[2 * i] = v.x; c[2 * i + 1] = v.y;
      }
    } else if (ty == 5u) {
      const int dim = 1 << k;
      for (int e = 0; e < dim * dim; ++e) {
        const int r = e / dim, q = e % dim;
        cplx v = adj ? m[q * dim + r] : m[r * dim + q];
        if (adj) v.y = -v.y;
        c[2 * e] = v.x; c[2 * e + 1] = v.y;
      }
    }
  }
  if (tid == 0) {
    double g = 1.0;
    for (int o = 0; o < B200QJ_NDESC; ++o) {
      const u32 ty = DESC[4 * o], src = DESC[4 * o + 1], fl = DESC[4 * o + 2];
      if (ty == 6u) g *= double(mats[src].x);
      else if ((ty == 1u || ty == 2u) && (fl & 2u) && mats[src].x < Real(0)) g = -g;
    }
    coef[B200QJ_SCALE_OFF] = Real(g);
  }
}
)SRC";

const char* kDestTab = R"SRC(
// destination tables of the fused exchange (see fill_dest_tab in b200q_tile_body.h): 5 x 256 entries
DEV void b200qj_fill_dest_tab(const Remote& R, int tid, u64* tab) {
  for (int e = tid; e < 1280; e += B200QJ_NT) {
    const int b = e >> 8, v = e & 255;
    u64 d = b == 0 ? R.base : 0ull;
    for (int j = 0; j < 8; ++j) {
      const int bit = 8 * b + j;
      if (!((v >> j) & 1) || bit >= R.n_chunk_bits) continue;
      const int pos = R.perm[bit];
      d |= pos < R.n_chunk_bits ? (1ull << pos) : (1ull << (40 + pos - R.n_chunk_bits));
    }
    tab[e] = d;
  }
}
DEV chunk* b200qj_dest(const Remote& R, const u64* tab, u64 idx) {
  const u64 d = tab[idx & 255u] | tab[256 + ((idx >> 8) & 255u)] | tab[512 + ((idx >> 16) & 255u)] |
                tab[768 + ((idx >> 24) & 255u)] | tab[1024 + ((idx >> 32) & 255u)];
  return reinterpret_cast<chunk*>(R.peer[d >> 40]) + (d & ((1ull << 40) - 1ull));
}
)SRC";

enum { D_ROTX = 1, D_ROTY = 2, D_MAT1 = 3, D_DIAG = 4, D_MATK = 5, D_HAD = 6 };
enum OpClass { C_NONE, C_HAD, C_ROT, C_MAT1, C_DIAG, C_X, C_LSWAP, C_MATK };

struct Gen {
  const Plan& pl;
  const b200q_pass_t& P;
  GenOptions opt;
  bool f32;
  int VS, RB, CB, T, NT, IPT, item_bits;
  std::ostringstream o;
  // per-op classification
  std::vector<int> cls, coef_off;
  std::vector<uint32_t> desc;
  int ncoef = 0, scale_off = 0;
  // per-round state
  int vr[16], vi[16];
  bool mx[5], mz[5], msg, mrho;
  int alias;
  int tmp_id = 0;
  // statistics
  int n_free_phase = 0;
  int n_rename = 0, n_frame_x = 0, n_mat_x = 0, n_mat_z = 0, n_rho = 0, n_generic_diag = 0, n_phys_x = 0;

  Gen(const Plan& pl_, const b200q_pass_t& P_, const GenOptions& opt_) : pl(pl_), P(P_), opt(opt_) {
    f32 = pl.dtype == B200Q_C64;
    VS = f32 ? 1 : 0;
    RB = B200Q_REG_CHUNK_BITS + VS;
    CB = pl.opt.chunk_bits;
    T = P.tile_bits;
    IPT = std::max(1, opt.items_per_thread);
    NT = (1 << (CB - B200Q_REG_CHUNK_BITS)) / IPT;
    item_bits = T - RB;
  }

  static std::string vfmt(const char* fmt, va_list ap) {
    va_list ap2;
    va_copy(ap2, ap);
    const int n = vsnprintf(nullptr, 0, fmt, ap2);
    va_end(ap2);
    std::string s(size_t(n > 0 ? n : 0), '\0');
    if (n > 0) vsnprintf(&s[0], size_t(n) + 1, fmt, ap);
    return s;
  }
  void pf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    o << vfmt(fmt, ap);
    va_end(ap);
  }
  static std::string sf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::string s = vfmt(fmt, ap);
    va_end(ap);
    return s;
  }
  std::string W(int v) const { return "w" + std::to_string(v); }
  std::string RL(double v) const {   // Real literal (re-entrant: passes are generated on several threads)
    char buf[64];
    snprintf(buf, sizeof buf, f32 ? "%.9g" : "%.17g", v);
    std::string s(buf);
    if (s.find('.') == std::string::npos && s.find('e') == std::string::npos) s += ".0";
    if (f32) s += "f";
    return s;
  }

  // expression that deposits bit k of `x` at position pos[k] (k = 0..n-1), consecutive runs merged
  std::string deposit(const std::string& x, const std::vector<int>& pos, bool wide) const {
    std::string out;
    const int n = (int)pos.size();
    for (int k = 0; k < n;) {
      int len = 1;
      while (k + len < n && pos[k + len] == pos[k] + len) ++len;
      const unsigned long long mask = (1ull << len) - 1ull;
      std::string t = wide ? "(u64)(" + x + ")" : "(" + x + ")";
      std::string term;
      if (wide) term = sf("((((%s) >> %d) & 0x%llxull) << %d)", t.c_str(), k, mask, pos[k]);
      else term = sf("((((%s) >> %d) & 0x%llxu) << %d)", t.c_str(), k, mask, pos[k]);
      out += out.empty() ? term : " | " + term;
      k += len;
    }
    if (out.empty()) out = wide ? "0ull" : "0u";
    return out;
  }

  // ---- classification + coefficient table ---------------------------------------------------------------------
  void classify() {
    const int n = P.n_ops;
    cls.assign(n, C_NONE);
    coef_off.assign(n, 0);
    auto add_desc = [&](int type, const b200q_op_t& op, int extra_flags, int k, int nreal) {
      const int off = ncoef;
      desc.push_back((uint32_t)type);
      desc.push_back(op.mat_src);
      desc.push_back((uint32_t)(((op.flags & B200Q_FLAG_ADJOINT) ? 1u : 0u) | (uint32_t)extra_flags | ((uint32_t)k << 8)));
      desc.push_back((uint32_t)off);
      ncoef += (nreal + 3) & ~3;
      return off;
    };
    for (int i = 0; i < n; ++i) {
      const b200q_op_t& op = P.ops[i];
      const bool tctrl = op.ctrl_loc != 0 || op.ctrl_glob != 0;
      switch (op.kind) {
        case B200Q_OP_X: cls[i] = C_X; break;
        case B200Q_OP_LSWAP: cls[i] = C_LSWAP; break;
        case B200Q_OP_MAT1:
          if ((op.flags & B200Q_FLAG_HAD) && (op.flags & B200Q_FLAG_REAL) && op.ctrl_reg == 0 && !tctrl) {
            cls[i] = C_HAD;
            coef_off[i] = add_desc(D_HAD, op, 0, 0, 0);
          } else if ((op.flags & B200Q_FLAG_ROT) && (op.flags & (B200Q_FLAG_REAL | B200Q_FLAG_RXLIKE)) && op.ctrl_reg == 0) {
            cls[i] = C_ROT;
            coef_off[i] = add_desc((op.flags & B200Q_FLAG_RXLIKE) ? D_ROTX : D_ROTY, op, tctrl ? 0 : 2, 0, 6);
          } else {
            cls[i] = C_MAT1;
            coef_off[i] = add_desc(D_MAT1, op, 0, 0, 8);
          }
          break;
        case B200Q_OP_DIAG:
          cls[i] = C_DIAG;
          coef_off[i] = add_desc(D_DIAG, op, 0, op.k, 8);
          break;
        case B200Q_OP_MATK:
          cls[i] = C_MATK;
          coef_off[i] = add_desc(D_MATK, op, 0, op.k, 2 << (2 * op.k));
          break;
        default: break;
      }
    }
    scale_off = ncoef;
    ncoef += 4;
  }

  // ---- frame / slot helpers --------------------------------------------------------------------------------------
  bool is_lane(int a) const { return VS && a == 0; }
  int cbit(int a) const { return 1 << (a - VS); }   // element-index bit of chunk slot `a` (amplitude-level index)
  int resolve(int a) const {
    if (alias < 0) return a;
    if (a == alias) return 0;
    if (a == 0) return alias;
    return a;
  }
  uint32_t resolve_mask(uint32_t m) const {
    uint32_t r = 0;
    for (int a = 0; a < RB; ++a)
      if (m >> a & 1u) r |= 1u << resolve(a);
    return r;
  }
  std::string FX(int a) const { return "fx" + std::to_string(a); }
  std::string FZ(int a) const { return "fz" + std::to_string(a); }

  std::string pred(const b200q_op_t& op) const {
    std::string s;
    if (op.ctrl_loc) s = sf("((lb & 0x%xu) == 0x%xu)", op.ctrl_loc, op.ctrl_loc);
    if (op.ctrl_glob) {
      const std::string g = sf("((cb & 0x%llxull) == 0x%llxull)", (unsigned long long)op.ctrl_glob,
                               (unsigned long long)op.ctrl_glob);
      s = s.empty() ? g : s + " && " + g;
    }
    return s;
  }

  void swap_vars(int a, int b) {
    pf("    { const V t_ = %s; %s = %s; %s = t_; }\n", W(a).c_str(), W(a).c_str(), W(b).c_str(), W(b).c_str());
  }

  // apply a pending X frame bit physically
  void mat_x(int a) {
    if (!mx[a]) return;
    ++n_mat_x;
    // X^x Z^z = (-1)^(x z) Z^z X^x: applying the X bit while a Z bit stays pending costs a sign
    if (mz[a]) { pf("    sg ^= %s & %s;\n", FX(a).c_str(), FZ(a).c_str()); msg = true; }
    pf("    if (%s) {\n", FX(a).c_str());
    if (is_lane(a)) {
      for (int e = 0; e < 16; ++e) {
        pf("      %s = vmk(vy(%s), vx(%s)); %s = vmk(vy(%s), vx(%s));\n", W(vr[e]).c_str(), W(vr[e]).c_str(),
           W(vr[e]).c_str(), W(vi[e]).c_str(), W(vi[e]).c_str(), W(vi[e]).c_str());
      }
    } else {
      const int b = cbit(a);
      for (int e = 0; e < 16; ++e)
        if (!(e & b)) { pf("  "); swap_vars(vr[e], vr[e | b]); pf("  "); swap_vars(vi[e], vi[e | b]); }
    }
    pf("    }\n    %s = 0u;\n", FX(a).c_str());
    mx[a] = false;
  }
  // apply a pending Z frame bit physically
  void mat_z(int a) {
    if (!mz[a]) return;
    ++n_mat_z;
    if (is_lane(a)) {
      pf("    { const V s_ = vmk(%s, %s ? %s : %s);\n", RL(1).c_str(), FZ(a).c_str(), RL(-1).c_str(), RL(1).c_str());
      for (int e = 0; e < 16; ++e)
        pf("      vsc(%s, s_); vsc(%s, s_);\n", W(vr[e]).c_str(), W(vi[e]).c_str());
    } else {
      const int b = cbit(a);
      pf("    { const V s_ = vbc(%s ? %s : %s);\n", FZ(a).c_str(), RL(-1).c_str(), RL(1).c_str());
      for (int e = 0; e < 16; ++e)
        if (e & b)
          pf("      vsc(%s, s_); vsc(%s, s_);\n", W(vr[e]).c_str(), W(vi[e]).c_str());
    }
    pf("    }\n    %s = 0u;\n", FZ(a).c_str());
    mz[a] = false;
  }

  // rho *= (pr + i pi)
  void rho_mul(const std::string& pr, const std::string& pi) {
    pf("      { const Real t_ = rr * %s - ri * %s; ri = rr * %s + ri * %s; rr = t_; }\n", pr.c_str(), pi.c_str(), pi.c_str(),
       pr.c_str());
    mrho = true;
  }
  // element (re, im) *= (pr + i pi), coefficients given as V expressions (pr, pi, -pi)
  void cmul_elem(int e, const char* pr, const char* pi, const char* npi) {
    const std::string r = W(vr[e]), i = W(vi[e]);
    pf("      vcm(%s, %s, %s, %s, %s);\n", r.c_str(), i.c_str(), pr, pi, npi);
  }

  // ---- ops ------------------------------------------------------------------------------------------------------
  void op_had(int a) {
    pf("    // H-structured on slot %d\n", a);
    if (is_lane(a)) {
      for (int e = 0; e < 16; ++e)
        for (int c = 0; c < 2; ++c) {
          const std::string w = W(c ? vi[e] : vr[e]);
          pf("    %s = vmk(vx(%s) + vy(%s), vx(%s) - vy(%s));\n", w.c_str(), w.c_str(), w.c_str(), w.c_str(), w.c_str());
        }
    } else {
      const int b = cbit(a);
      for (int e = 0; e < 16; ++e) {
        if (e & b) continue;
        for (int c = 0; c < 2; ++c) {
          const std::string x = W(c ? vi[e] : vr[e]), y = W(c ? vi[e | b] : vr[e | b]);
          pf("    vhad(%s, %s);\n", x.c_str(), y.c_str());
        }
      }
    }
    // H X^x Z^z = (-1)^(x z) X^z Z^x H
    if (mx[a] && mz[a]) { pf("    sg ^= %s & %s;\n", FX(a).c_str(), FZ(a).c_str()); msg = true; }
    if (mx[a] || mz[a]) {
      pf("    { const u32 t_ = %s; %s = %s; %s = t_; }\n", FX(a).c_str(), FX(a).c_str(), FZ(a).c_str(), FZ(a).c_str());
      std::swap(mx[a], mz[a]);
    }
  }

  void op_rot(const b200q_op_t& op, int oi, int a) {
    const bool isx = (op.flags & B200Q_FLAG_RXLIKE) != 0;
    const std::string p = pred(op);
    const int off = coef_off[oi];
    pf("    { // %s-structured rotation on slot %d%s\n", isx ? "Rx" : "Ry", a, p.empty() ? "" : " (thread-level control)");
    // frame: X swaps the off-diagonal entries (record [2], [3]); Z negates them
    if (mx[a]) pf("      Real u_ = %s ? coef[%d] : coef[%d], v_ = %s ? coef[%d] : coef[%d];\n", FX(a).c_str(), off + 2, off,
                  FX(a).c_str(), off + 3, off + 1);
    else pf("      Real u_ = coef[%d], v_ = coef[%d];\n", off, off + 1);
    if (mz[a]) pf("      if (%s) { u_ = -u_; v_ = -v_; }\n", FZ(a).c_str());
    if (!p.empty()) pf("      if (%s) {\n", p.c_str());
    if (is_lane(a)) {
      for (int e = 0; e < 16; ++e) {
        const std::string r = W(vr[e]), i = W(vi[e]);
        pf("      { Real ar = vx(%s), ai = vx(%s), br = vy(%s), bi = vy(%s);\n", r.c_str(), i.c_str(), r.c_str(), i.c_str());
        if (isx)
          pf("        ar -= u_ * bi; ai += u_ * br; br -= v_ * ai; bi += v_ * ar; ar -= u_ * bi; ai += u_ * br;\n");
        else
          pf("        ar += u_ * br; ai += u_ * bi; br += v_ * ar; bi += v_ * ai; ar += u_ * br; ai += u_ * bi;\n");
        pf("        %s = vmk(ar, br); %s = vmk(ai, bi); }\n", r.c_str(), i.c_str());
      }
    } else {
      pf("      const V U_ = vbc(u_), NU_ = vbc(-u_), V_ = vbc(v_), NV_ = vbc(-v_);\n");
      const int b = cbit(a);
      for (int e = 0; e < 16; ++e) {
        if (e & b) continue;
        const std::string ar = W(vr[e]), ai = W(vi[e]), br = W(vr[e | b]), bi = W(vi[e | b]);
        if (isx) {
          pf("      vfa(%s, NU_, %s); vfa(%s, U_, %s);\n", ar.c_str(), bi.c_str(), ai.c_str(), br.c_str());
          pf("      vfa(%s, NV_, %s); vfa(%s, V_, %s);\n", br.c_str(), ai.c_str(), bi.c_str(), ar.c_str());
          pf("      vfa(%s, NU_, %s); vfa(%s, U_, %s);\n", ar.c_str(), bi.c_str(), ai.c_str(), br.c_str());
        } else {
          pf("      vfa(%s, U_, %s); vfa(%s, U_, %s);\n", ar.c_str(), br.c_str(), ai.c_str(), bi.c_str());
          pf("      vfa(%s, V_, %s); vfa(%s, V_, %s);\n", br.c_str(), ar.c_str(), bi.c_str(), ai.c_str());
          pf("      vfa(%s, U_, %s); vfa(%s, U_, %s);\n", ar.c_str(), br.c_str(), ai.c_str(), bi.c_str());
        }
      }
    }
    if (!p.empty()) {
      // the sign of a negated controlled rotation cannot go to the pass scale
      pf("        if (coef[%d] != %s) sg ^= 1u;\n      }\n", off + 4, RL(0).c_str());
      msg = true;
    }
    pf("    }\n");
  }

  // dense 2x2 on slot `a`, register controls `cr` (resolved mask), thread predicate
  void op_mat1(const b200q_op_t& op, int oi, int a, uint32_t cr) {
    mat_z(a);   // the registers hold X^fx Z^fz |psi>: Z first, then X
    mat_x(a);
    for (int c = 0; c < RB; ++c)
      if (cr >> c & 1u) mat_x(c);
    const std::string p = pred(op);
    const int off = coef_off[oi];
    const bool real = (op.flags & B200Q_FLAG_REAL) != 0, rxl = (op.flags & B200Q_FLAG_RXLIKE) != 0 && !real;
    const bool lane_ctrl = VS && (cr & 1u);
    uint32_t cm = cr >> VS;   // element-index control mask
    pf("    { // dense 2x2 on slot %d, register controls 0x%x%s\n", a, cr, p.empty() ? "" : ", thread-level control");
    pf("      const Real m00r = coef[%d], m00i = coef[%d], m01r = coef[%d], m01i = coef[%d];\n", off, off + 1, off + 2, off + 3);
    pf("      const Real m10r = coef[%d], m10i = coef[%d], m11r = coef[%d], m11i = coef[%d];\n", off + 4, off + 5, off + 6,
       off + 7);
    if (!p.empty()) pf("      if (%s) {\n", p.c_str());
    if (is_lane(a)) {
      for (int e = 0; e < 16; ++e) {
        if ((uint32_t(e) & cm) != cm) continue;
        const std::string r = W(vr[e]), i = W(vi[e]);
        pf("      { const Real ar = vx(%s), ai = vx(%s), br = vy(%s), bi = vy(%s);\n", r.c_str(), i.c_str(), r.c_str(),
           i.c_str());
        if (real)
          pf("        %s = vmk(m00r * ar + m01r * br, m10r * ar + m11r * br); %s = vmk(m00r * ai + m01r * bi, m10r * ai + m11r * bi); }\n",
             r.c_str(), i.c_str());
        else if (rxl)
          pf("        %s = vmk(m00r * ar - m01i * bi, m11r * br - m10i * ai); %s = vmk(m00r * ai + m01i * br, m11r * bi + m10i * ar); }\n",
             r.c_str(), i.c_str());
        else
          pf("        %s = vmk(m00r * ar - m00i * ai + m01r * br - m01i * bi, m10r * ar - m10i * ai + m11r * br - m11i * bi);\n"
             "        %s = vmk(m00r * ai + m00i * ar + m01r * bi + m01i * br, m10r * ai + m10i * ar + m11r * bi + m11i * br); }\n",
             r.c_str(), i.c_str());
      }
    } else {
      pf("      const V M00R = vbc(m00r), M01R = vbc(m01r), M10R = vbc(m10r), M11R = vbc(m11r);\n");
      pf("      const V M00I = vbc(m00i), M01I = vbc(m01i), M10I = vbc(m10i), M11I = vbc(m11i);\n");
      pf("      const V N00I = vbc(-m00i), N01I = vbc(-m01i), N10I = vbc(-m10i), N11I = vbc(-m11i);\n");
      const int b = cbit(a);
      for (int e = 0; e < 16; ++e) {
        if (e & b) continue;
        if ((uint32_t(e) & cm) != cm) continue;
        const std::string ar = W(vr[e]), ai = W(vi[e]), br = W(vr[e | b]), bi = W(vi[e | b]);
        pf("      { const V ar = %s, ai = %s, br = %s, bi = %s;\n", ar.c_str(), ai.c_str(), br.c_str(), bi.c_str());
        if (real) {
          pf("        const V xr = vfma(M00R, ar, vmul(M01R, br)), xi = vfma(M00R, ai, vmul(M01R, bi));\n");
          pf("        const V yr = vfma(M11R, br, vmul(M10R, ar)), yi = vfma(M11R, bi, vmul(M10R, ai));\n");
        } else if (rxl) {
          pf("        const V xr = vfma(M00R, ar, vmul(N01I, bi)), xi = vfma(M00R, ai, vmul(M01I, br));\n");
          pf("        const V yr = vfma(M11R, br, vmul(N10I, ai)), yi = vfma(M11R, bi, vmul(M10I, ar));\n");
        } else {
          pf("        const V xr = vfma(M00R, ar, vfma(N00I, ai, vfma(M01R, br, vmul(N01I, bi))));\n");
          pf("        const V xi = vfma(M00R, ai, vfma(M00I, ar, vfma(M01R, bi, vmul(M01I, br))));\n");
          pf("        const V yr = vfma(M10R, ar, vfma(N10I, ai, vfma(M11R, br, vmul(N11I, bi))));\n");
          pf("        const V yi = vfma(M10R, ai, vfma(M10I, ar, vfma(M11R, bi, vmul(M11I, br))));\n");
        }
        if (lane_ctrl)
          pf("        %s = vmk(vx(ar), vy(xr)); %s = vmk(vx(ai), vy(xi)); %s = vmk(vx(br), vy(yr)); %s = vmk(vx(bi), vy(yi)); }\n",
             ar.c_str(), ai.c_str(), br.c_str(), bi.c_str());
        else
          pf("        %s = xr; %s = xi; %s = yr; %s = yi; }\n", ar.c_str(), ai.c_str(), br.c_str(), bi.c_str());
      }
    }
    if (!p.empty()) pf("      }\n");
    pf("    }\n");
  }

  // X (amplitude swap) on slot `a`, register controls `cr` (resolved), thread predicate
  void op_x(const b200q_op_t& op, int a, uint32_t cr) {
    const std::string p = pred(op);
    const int ncr = __builtin_popcount(cr);
    if (ncr == 0) {   // pure frame update
      ++n_frame_x;
      if (p.empty()) pf("    %s ^= 1u;   // X on slot %d\n", FX(a).c_str(), a);
      else pf("    %s ^= (u32)(%s);   // X on slot %d, thread-level control\n", FX(a).c_str(), p.c_str(), a);
      mx[a] = true;
      return;
    }
    if (ncr == 1 && p.empty()) {   // CNOT between register slots: Clifford frame rule + data movement
      const int c = __builtin_ctz(cr);
      pf("    // CNOT slot %d -> slot %d\n", c, a);
      if (!is_lane(a) && !is_lane(c)) {
        ++n_rename;
        const int bt = cbit(a), bc = cbit(c);
        for (int e = 0; e < 16; ++e)
          if ((e & bc) && !(e & bt)) { std::swap(vr[e], vr[e | bt]); std::swap(vi[e], vi[e | bt]); }
      } else if (is_lane(a)) {   // swap the two lanes of the elements whose control bit is set
        ++n_phys_x;
        const int bc = cbit(c);
        for (int e = 0; e < 16; ++e)
          if (e & bc)
            pf("    %s = vmk(vy(%s), vx(%s)); %s = vmk(vy(%s), vx(%s));\n", W(vr[e]).c_str(), W(vr[e]).c_str(),
               W(vr[e]).c_str(), W(vi[e]).c_str(), W(vi[e]).c_str(), W(vi[e]).c_str());
      } else {   // control = lane: exchange the high lanes of every pair
        ++n_phys_x;
        const int bt = cbit(a);
        for (int e = 0; e < 16; ++e) {
          if (e & bt) continue;
          for (int k = 0; k < 2; ++k) {
            const std::string x = W(k ? vi[e] : vr[e]), y = W(k ? vi[e | bt] : vr[e | bt]);
            pf("    { const Real t_ = vy(%s); %s = vmk(vx(%s), vy(%s)); %s = vmk(vx(%s), t_); }\n", x.c_str(), x.c_str(),
               x.c_str(), y.c_str(), y.c_str(), y.c_str());
          }
        }
      }
      // CNOT X_c = X_c X_t CNOT ; CNOT Z_t = Z_c Z_t CNOT
      if (mx[c]) { pf("    %s ^= %s;\n", FX(a).c_str(), FX(c).c_str()); mx[a] = true; }
      if (mz[a]) { pf("    %s ^= %s;\n", FZ(c).c_str(), FZ(a).c_str()); mz[c] = true; }
      return;
    }
    // general: bring the frame of every involved slot to zero, then swap physically
    ++n_phys_x;
    mat_z(a);   // the registers hold X^fx Z^fz |psi>: Z first, then X
    mat_x(a);
    for (int c = 0; c < RB; ++c)
      if (cr >> c & 1u) mat_x(c);
    const bool lane_ctrl = VS && (cr & 1u);
    const uint32_t cm = cr >> VS;
    pf("    { // X on slot %d, register controls 0x%x\n", a, cr);
    if (!p.empty()) pf("      if (%s) {\n", p.c_str());
    if (is_lane(a)) {
      for (int e = 0; e < 16; ++e) {
        if ((uint32_t(e) & cm) != cm) continue;
        pf("      %s = vmk(vy(%s), vx(%s)); %s = vmk(vy(%s), vx(%s));\n", W(vr[e]).c_str(), W(vr[e]).c_str(),
           W(vr[e]).c_str(), W(vi[e]).c_str(), W(vi[e]).c_str(), W(vi[e]).c_str());
      }
    } else {
      const int b = cbit(a);
      for (int e = 0; e < 16; ++e) {
        if (e & b) continue;
        if ((uint32_t(e) & cm) != cm) continue;
        for (int k = 0; k < 2; ++k) {
          const int x = k ? vi[e] : vr[e], y = k ? vi[e | b] : vr[e | b];
          if (lane_ctrl)
            pf("      { const Real t_ = vy(%s); %s = vmk(vx(%s), vy(%s)); %s = vmk(vx(%s), t_); }\n", W(x).c_str(),
               W(x).c_str(), W(x).c_str(), W(y).c_str(), W(y).c_str(), W(y).c_str());
          else { pf("  "); swap_vars(x, y); }
        }
      }
    }
    if (!p.empty()) pf("      }\n");
    pf("    }\n");
  }

  void op_diag(const b200q_op_t& op, int oi) {
    const int off = coef_off[oi];
    const std::string p = pred(op);
    const uint32_t cr = resolve_mask(op.ctrl_reg);
    int sel_a[2] = {-1, -1};   // resolved register slot of selector j, or -1
    std::string tsel;          // thread-level part of the diagonal index
    int nreg = 0;
    for (int j = 0; j < int(op.k) && j < 2; ++j) {
      if (op.dsel_slot[j] != 0xff) { sel_a[j] = resolve(op.dsel_slot[j]); ++nreg; continue; }
      std::string t;
      if (op.dsel_loc[j]) t = sf("(((lb >> %d) & 1u) << %d)", __builtin_ctz(op.dsel_loc[j]), j);
      else if (op.dsel_glob[j]) t = sf("((u32)((cb >> %d) & 1ull) << %d)", __builtin_ctzll(op.dsel_glob[j]), j);
      if (!t.empty()) tsel = tsel.empty() ? t : tsel + " | " + t;
    }
    if (tsel.empty()) tsel = "0u";
    // hinted diag(1, i^q): no arithmetic when the qubit is a register slot (or, for Z, anywhere)
    int q = int(op.k) == 1 ? int((op.flags & B200Q_FLAG_PHASE_MASK) >> B200Q_FLAG_PHASE_SHIFT) : 0;
    if (q && (op.flags & B200Q_FLAG_ADJOINT)) q = (4 - q) & 3;
    if (q == 2 && cr == 0) {
      if (nreg == 1) {
        const int a = sel_a[0];
        // Z X^x Z^z = (-1)^x X^x Z^(z+1)
        if (mx[a]) {
          if (p.empty()) pf("    sg ^= %s;\n", FX(a).c_str());
          else pf("    sg ^= %s & (u32)(%s);\n", FX(a).c_str(), p.c_str());
          msg = true;
        }
        if (p.empty()) pf("    %s ^= 1u;   // Z on slot %d\n", FZ(a).c_str(), a);
        else pf("    %s ^= (u32)(%s);   // Z on slot %d, thread-level control\n", FZ(a).c_str(), p.c_str(), a);
        mz[a] = true;
      } else {
        if (p.empty()) pf("    sg ^= %s;   // Z on a thread-level bit\n", tsel.c_str());
        else pf("    sg ^= (%s) & (u32)(%s);   // controlled Z on thread-level bits\n", tsel.c_str(), p.c_str());
        msg = true;
      }
      ++n_free_phase;
      return;
    }
    if ((q == 1 || q == 3) && cr == 0 && nreg == 1 && p.empty()) {
      // S X^x Z^z = i^x X^x S Z^(x+z);  S^dagger = S Z: (-i)^x, one more Z
      const int a = sel_a[0];
      pf("    // %s on slot %d: re <-> im renaming\n", q == 1 ? "S" : "S^dagger", a);
      if (is_lane(a)) {
        for (int e = 0; e < 16; ++e) {
          const std::string r = W(vr[e]), i = W(vi[e]);
          pf("    { const Real t_ = vy(%s); %s = vmk(vx(%s), -vy(%s)); %s = vmk(vx(%s), t_); }\n", r.c_str(), r.c_str(),
             r.c_str(), i.c_str(), i.c_str(), i.c_str());
        }
      } else {
        const int b = cbit(a);
        for (int e = 0; e < 16; ++e) {
          if (!(e & b)) continue;
          std::swap(vr[e], vi[e]);
          pf("    vng(%s);\n", W(vr[e]).c_str());
        }
      }
      if (mx[a]) {
        pf("    if (%s) { const Real t_ = rr; rr = %sri; ri = %st_; }\n", FX(a).c_str(), q == 1 ? "-" : "", q == 1 ? "" : "-");
        mrho = true;
        pf("    %s ^= %s;\n", FZ(a).c_str(), FX(a).c_str());
        mz[a] = true;
      }
      if (q == 3) { pf("    %s ^= 1u;\n", FZ(a).c_str()); mz[a] = true; }
      ++n_free_phase;
      return;
    }
    pf("    { // diagonal, %d selector(s), %d on register slots\n", int(op.k), nreg);
    if (nreg == 0 && cr == 0) {
      if (!p.empty()) pf("      if (%s)\n", p.c_str());
      pf("      { const u32 ix_ = %s; const Real pr_ = coef[%d + 2 * ix_], pi_ = coef[%d + 2 * ix_];\n", tsel.c_str(), off,
         off + 1);
      rho_mul("pr_", "pi_");
      pf("      }\n    }\n");
      return;
    }
    if (nreg == 1 && cr == 0) {
      const int j = sel_a[0] >= 0 ? 0 : 1;
      const int a = sel_a[j];
      // entries for register value 0 / 1 of the selector; the X frame bit exchanges them
      if (mx[a])
        pf("      const u32 i0_ = (%s) | (%s << %d), i1_ = (%s) | ((%s ^ 1u) << %d);\n", tsel.c_str(), FX(a).c_str(), j,
           tsel.c_str(), FX(a).c_str(), j);
      else pf("      const u32 i0_ = (%s), i1_ = (%s) | %uu;\n", tsel.c_str(), tsel.c_str(), 1u << j);
      pf("      const Real p0r = coef[%d + 2 * i0_], p0i = coef[%d + 2 * i0_], p1r = coef[%d + 2 * i1_], p1i = coef[%d + 2 * i1_];\n",
         off, off + 1, off, off + 1);
      if (!p.empty()) pf("      if (%s) {\n", p.c_str());
      if (is_lane(a)) {
        pf("      const V PR_ = vmk(p0r, p1r), PI_ = vmk(p0i, p1i), NPI_ = vmk(-p0i, -p1i);\n");
        for (int e = 0; e < 16; ++e) cmul_elem(e, "PR_", "PI_", "NPI_");
      } else {
        pf("      const V P0R = vbc(p0r), P0I = vbc(p0i), N0I = vbc(-p0i), P1R = vbc(p1r), P1I = vbc(p1i), N1I = vbc(-p1i);\n");
        const int b = cbit(a);
        for (int e = 0; e < 16; ++e) {
          if (e & b) cmul_elem(e, "P1R", "P1I", "N1I");
          else cmul_elem(e, "P0R", "P0I", "N0I");
        }
      }
      if (!p.empty()) pf("      }\n");
      pf("    }\n");
      return;
    }
    // general: two register selectors and / or register controls
    ++n_generic_diag;
    for (int j = 0; j < 2; ++j)
      if (sel_a[j] >= 0) mat_x(sel_a[j]);
    for (int c = 0; c < RB; ++c)
      if (cr >> c & 1u) mat_x(c);
    if (!p.empty()) pf("      if (%s) {\n", p.c_str());
    pf("      const u32 ts_ = %s;\n", tsel.c_str());
    const int NL = 1 << VS;
    std::map<std::pair<int, int>, int> groups;
    for (int e = 0; e < 16; ++e) {
      int spec[2] = {-1, -1};
      for (int l = 0; l < NL; ++l) {
        const uint32_t i = (uint32_t(e) << VS) | uint32_t(l);   // amplitude-level register index
        if ((i & cr) != cr) continue;
        int rp = 0;
        for (int j = 0; j < 2; ++j)
          if (sel_a[j] >= 0 && (i >> sel_a[j] & 1u)) rp |= 1 << j;
        spec[l] = rp;
      }
      if (spec[0] < 0 && spec[1] < 0) continue;
      const std::pair<int, int> key(spec[0], NL == 2 ? spec[1] : spec[0]);
      auto it = groups.find(key);
      int gid;
      if (it == groups.end()) {
        gid = tmp_id++;
        groups[key] = gid;
        auto ent = [&](int s, const char* part) {
          if (s < 0) return part[0] == 'r' ? RL(1) : RL(0);
          return sf("coef[%d + 2 * (ts_ | %du)]", off + (part[0] == 'r' ? 0 : 1), s);
        };
        if (NL == 2) {
          pf("      const V GR%d = vmk(%s, %s), GI%d = vmk(%s, %s), GN%d = vneg(GI%d);\n", gid, ent(spec[0], "r").c_str(),
             ent(spec[1], "r").c_str(), gid, ent(spec[0], "i").c_str(), ent(spec[1], "i").c_str(), gid, gid);
        } else {
          pf("      const V GR%d = vbc(%s), GI%d = vbc(%s), GN%d = vneg(GI%d);\n", gid, ent(spec[0], "r").c_str(), gid,
             ent(spec[0], "i").c_str(), gid, gid);
        }
      } else gid = it->second;
      const std::string gr = "GR" + std::to_string(gid), gi = "GI" + std::to_string(gid), gn = "GN" + std::to_string(gid);
      cmul_elem(e, gr.c_str(), gi.c_str(), gn.c_str());
    }
    if (!p.empty()) pf("      }\n");
    pf("    }\n");
  }

  // ---- one register round ----------------------------------------------------------------------------------------
  void emit_round(int r) {
    const b200q_round_t& Rd = P.rounds[r];
    const bool src_soa = (P.layout & B200Q_LAYOUT_SRC_SOA) != 0, dst_soa = (P.layout & B200Q_LAYOUT_DST_SOA) != 0;
    pf("\n// ---- round %d: ops %d..%d, %s -> %s\n", r, int(Rd.op_begin), int(Rd.op_end) - 1,
       Rd.src_global ? "global" : "tile", Rd.dst_global ? "global" : "tile");
    pf("DEV void b200qj_round%d(const int tid, const u64 cb, chunk* __restrict__ tile, chunk* __restrict__ g, "
       "const Real* __restrict__ coef, const Remote& rm, const u64* __restrict__ dest_tab) {\n", r);
    std::vector<int> lbpos, gpos;
    for (int k = 0; k < item_bits; ++k) {
      lbpos.push_back(Rd.nonreg_bit[k]);
      gpos.push_back(int(P.tile_phys[Rd.nonreg_bit[k]]) - VS);
    }
    // IPT > 1: a thread walks IPT items of the round one after the other (fewer threads per tile: more CTAs, i.e.
    // more tiles in different phases, fit the register file of an SM)
    if (IPT > 1) pf("#if defined(__CUDACC__)\n#pragma unroll 1\n#endif\n  for (int it_ = 0; it_ < %d; ++it_) {\n  const int vt = tid + it_ * %d;\n", IPT, NT);
    else pf("  {\n  const int vt = tid;\n");
    pf("  const u32 lb = %s;\n", deposit("(u32)vt", lbpos, false).c_str());
    pf("  (void)lb; (void)cb; (void)coef; (void)rm; (void)dest_tab;\n");
    uint64_t gst[4];
    uint32_t sst[4];
    for (int s = 0; s < 4; ++s) {
      const int loc = Rd.slot_bit[s + VS];
      gst[s] = 1ull << (int(P.tile_phys[loc]) - VS);
      sst[s] = swz_host(1u << (loc - VS));
    }
    auto goff = [&](int e) { uint64_t v = 0; for (int s = 0; s < 4; ++s) if (e >> s & 1) v += gst[s]; return v; };
    auto soff = [&](int e) { uint32_t v = 0; for (int s = 0; s < 4; ++s) if (e >> s & 1) v ^= sst[s]; return v; };
    if (Rd.src_global || Rd.dst_global)
      pf("  const u64 gb = (cb >> %d) | (%s);\n", VS, deposit("(u32)vt", gpos, true).c_str());
    if (!Rd.src_global || !Rd.dst_global) pf("  const u32 sb = b200qj_swz(lb >> %d);\n", VS);
    pf("  V ");
    for (int v = 0; v < 32; ++v) pf("w%d%s", v, v == 31 ? ";\n" : ", ");
    // gather
    for (int e = 0; e < 16; ++e) {
      vr[e] = 2 * e;
      vi[e] = 2 * e + 1;
      if (Rd.src_global) pf("  { const chunk c_ = g[gb + 0x%llxull];", (unsigned long long)goff(e));
      else pf("  { const chunk c_ = tile[sb ^ 0x%xu];", soff(e));
      if (f32 && Rd.src_global && !src_soa)
        pf(" w%d = vmk(vx(c_.lo), vx(c_.hi)); w%d = vmk(vy(c_.lo), vy(c_.hi)); }\n", 2 * e, 2 * e + 1);
      else pf(" w%d = c_.lo; w%d = c_.hi; }\n", 2 * e, 2 * e + 1);
    }
    // frame
    for (int a = 0; a < RB; ++a) { mx[a] = mz[a] = false; pf("  u32 fx%d = 0u, fz%d = 0u;\n", a, a); }
    pf("  u32 sg = 0u; Real rr = %s, ri = %s;\n", RL(1).c_str(), RL(0).c_str());
    msg = mrho = false;
    alias = -1;
    for (int oi = Rd.op_begin; oi < Rd.op_end; ++oi) {
      const b200q_op_t& op = P.ops[oi];
      if (opt.debug_skip_ops == 1) continue;   // measurement only: the memory traffic of the pass without its arithmetic
      if (opt.debug_skip_ops == 2 && cls[oi] == C_DIAG) continue;   // (compiler experiments: leave out one op class)
      if (opt.debug_skip_ops == 3 && cls[oi] == C_HAD) continue;
      if (opt.debug_skip_ops == 4 && cls[oi] == C_ROT) continue;
      if (opt.debug_skip_ops == 5 && cls[oi] == C_X) continue;
      switch (cls[oi]) {
        case C_LSWAP:
          alias = alias < 0 ? int(op.slot) : -1;
          break;
        case C_HAD: op_had(resolve(op.slot)); break;
        case C_ROT: op_rot(op, oi, resolve(op.slot)); break;
        case C_MAT1: op_mat1(op, oi, resolve(op.slot), resolve_mask(op.ctrl_reg)); break;
        case C_X: op_x(op, resolve(op.slot), resolve_mask(op.ctrl_reg)); break;
        case C_DIAG: op_diag(op, oi); break;
        default: break;
      }
    }
    // end of round: Z frame bits and the lane X bit are applied; chunk-slot X bits go into the scatter address
    for (int a = 0; a < RB; ++a) mat_z(a);
    if (VS) mat_x(0);
    const bool scale_here = Rd.dst_global && scale_used;
    if (mrho || msg || scale_here) {
      ++n_rho;
      pf("  {\n");
      if (scale_here) pf("    const Real gs_ = coef[%d]; rr *= gs_; ri *= gs_;\n", scale_off);
      if (msg) pf("    if (sg) { rr = -rr; ri = -ri; }\n");
      if (mrho) {
        pf("    const V PR_ = vbc(rr), PI_ = vbc(ri), NPI_ = vbc(-ri);\n");
        for (int e = 0; e < 16; ++e) cmul_elem(e, "PR_", "PI_", "NPI_");
      } else {
        pf("    const V PR_ = vbc(rr);\n");
        for (int e = 0; e < 16; ++e)
          pf("      vsc(%s, PR_); vsc(%s, PR_);\n", W(vr[e]).c_str(), W(vi[e]).c_str());
      }
      pf("  }\n");
    }
    // scatter
    auto packed = [&](int e, bool aos) {
      const std::string re = W(vr[e]), im = W(vi[e]);
      if (aos) return sf("{ chunk c_; c_.lo = vmk(vx(%s), vx(%s)); c_.hi = vmk(vy(%s), vy(%s));", re.c_str(), im.c_str(),
                         re.c_str(), im.c_str());
      return sf("{ chunk c_; c_.lo = %s; c_.hi = %s;", re.c_str(), im.c_str());
    };
    if (Rd.dst_global) {
      const bool aos = f32 && !dst_soa;
      // element e of the registers is logical element e ^ fx: base gets the strides of the set frame bits, and the
      // stride of such a slot counts backwards
      pf("  u64 gw = gb;\n");
      bool any = false;
      for (int s = 0; s < 4; ++s)
        if (mx[s + VS]) {
          pf("  const i64 st%d = %s ? -(i64)0x%llxull : (i64)0x%llxull; gw += %s ? 0x%llxull : 0ull;\n", s,
             FX(s + VS).c_str(), (unsigned long long)gst[s], (unsigned long long)gst[s], FX(s + VS).c_str(),
             (unsigned long long)gst[s]);
          any = true;
        }
      (void)any;
      for (int e = 0; e < 16; ++e) {
        std::string idx = "gw";
        uint64_t c = 0;
        for (int s = 0; s < 4; ++s)
          if (e >> s & 1) {
            if (mx[s + VS]) idx += sf(" + st%d", s);
            else c += gst[s];
          }
        if (c) idx += sf(" + 0x%llxull", (unsigned long long)c);
        if (opt.remote)
          pf("  %s *b200qj_dest(rm, dest_tab, (u64)(%s)) = c_; }\n", packed(e, aos).c_str(), idx.c_str());
        else pf("  %s g[(u64)(%s)] = c_; }\n", packed(e, aos).c_str(), idx.c_str());
      }
    } else {
      pf("  u32 sw = sb;\n");
      for (int s = 0; s < 4; ++s)
        if (mx[s + VS]) pf("  sw ^= %s ? 0x%xu : 0u;\n", FX(s + VS).c_str(), sst[s]);
      for (int e = 0; e < 16; ++e) pf("  %s tile[sw ^ 0x%xu] = c_; }\n", packed(e, false).c_str(), soff(e));
    }
    pf("  }\n}\n");
  }

  // dense k-target op applied in place in the shared-memory tile (one op per direct round)
  void emit_direct(int oi) {
    const b200q_op_t& op = P.ops[oi];
    const int K = op.k, D = 1 << K, off = coef_off[oi];
    pf("\n// ---- dense %d-target op %d in the tile\n", K, oi);
    pf("DEV void b200qj_direct%d(const int tid, const u64 cb, chunk* __restrict__ tile, const Real* __restrict__ coef) {\n", oi);
    if (op.ctrl_glob)
      pf("  if ((cb & 0x%llxull) != 0x%llxull) return;\n", (unsigned long long)op.ctrl_glob, (unsigned long long)op.ctrl_glob);
    else pf("  (void)cb;\n");
    int srt[B200Q_MATK_MAX];
    for (int j = 0; j < K; ++j) srt[j] = op.tk[j];
    std::sort(srt, srt + K);
    pf("  for (int g_ = tid; g_ < %d; g_ += B200QJ_NT) {\n    u32 base = (u32)g_;\n", 1 << (T - K));
    for (int j = 0; j < K; ++j)
      pf("    base = ((base >> %d) << %d) | (base & 0x%xu);\n", srt[j], srt[j] + 1, (1u << srt[j]) - 1u);
    if (op.ctrl_loc) pf("    if ((base & 0x%xu) != 0x%xu) continue;\n", op.ctrl_loc, op.ctrl_loc);
    pf("    Real xr[%d], xi[%d]; Real* p_[%d];\n", D, D, D);
    for (int i = 0; i < D; ++i) {
      uint32_t o_ = 0;
      for (int j = 0; j < K; ++j)
        if (i >> j & 1) o_ |= 1u << op.tk[j];
      pf("    p_[%d] = b200qj_amp(tile, base | 0x%xu); xr[%d] = p_[%d][0]; xi[%d] = p_[%d][B200QJ_IMOFF];\n", i, o_, i, i, i, i);
    }
    pf("    for (int r_ = 0; r_ < %d; ++r_) {\n      Real yr = %s, yi = %s;\n", D, RL(0).c_str(), RL(0).c_str());
    pf("      for (int c_ = 0; c_ < %d; ++c_) {\n", D);
    pf("        const Real wr = coef[%d + 2 * (r_ * %d + c_)], wi = coef[%d + 2 * (r_ * %d + c_)];\n", off, D, off + 1, D);
    pf("        yr += wr * xr[c_] - wi * xi[c_]; yi += wr * xi[c_] + wi * xr[c_];\n      }\n");
    pf("      p_[r_][0] = yr; p_[r_][B200QJ_IMOFF] = yi;\n    }\n  }\n}\n");
  }

  bool scale_used = false;

  std::string run(size_t* smem_bytes, std::string* stats) {
    classify();
    for (int i = 0; i < P.n_ops; ++i)
      if (cls[i] == C_HAD || (cls[i] == C_ROT && !(P.ops[i].ctrl_loc || P.ops[i].ctrl_glob))) scale_used = true;
    const int ndesc = (int)desc.size() / 4;
    const size_t tile_bytes = size_t(16) << CB;
    const size_t coef_bytes = ((size_t(ncoef) * (f32 ? 4 : 8)) + 15) & ~size_t(15);
    const size_t total = tile_bytes + coef_bytes + (opt.remote ? 1280 * 8 : 0);
    if (smem_bytes) *smem_bytes = total;
    pf("// generated by b200q_codegen: %s, %d qubits, tile bits %d, %d rounds, %d ops\n", f32 ? "complex64" : "complex128",
       int(P.n_qubits), T, int(P.n_rounds), int(P.n_ops));
    pf("#define B200QJ_F32 %d\n#define B200QJ_NT %d\n#define B200QJ_NDESC %d\n#define B200QJ_SCALE_OFF %d\n", f32 ? 1 : 0, NT,
       std::max(ndesc, 1), scale_off);
    o << kPreamble;
    pf("B200QJ_CONST u32 DESC[%d] = {", std::max(ndesc, 1) * 4);
    if (ndesc == 0) pf("0u, 0u, 0u, 0u");
    for (size_t i = 0; i < desc.size(); ++i) pf("%s%uu", i ? ", " : "", desc[i]);
    pf("};\n");
    o << kPrep;
    if (opt.remote) o << kDestTab;
    else pf("DEV chunk* b200qj_dest(const Remote&, const u64*, u64) { return 0; }\n");
    struct Step { int round, direct_op; };
    std::vector<Step> steps;
    for (int r = 0; r < P.n_rounds; ++r) {
      const b200q_round_t& Rd = P.rounds[r];
      if (Rd.direct) {
        for (int oi = Rd.op_begin; oi < Rd.op_end; ++oi) { emit_direct(oi); steps.push_back({-1, oi}); }
      } else {
        emit_round(r);
        steps.push_back({r, -1});
      }
    }
    // tile index -> physical base
    std::vector<int> ntp;
    for (int j = 0; j < P.n_nontile; ++j) ntp.push_back(P.nontile_phys[j]);
    const std::string cbexpr = deposit("tile_id", ntp, true);
    auto call_step = [&](const Step& s, const char* ind) {
      if (s.round >= 0) pf("%sb200qj_round%d(tid, cb, tile, g, coef, rm, dest_tab);\n", ind, s.round);
      else pf("%sb200qj_direct%d(tid, cb, tile, coef);\n", ind, s.direct_op);
    };
    pf("\n#if defined(__CUDACC__)\n");
    pf("extern \"C\" __global__ void __launch_bounds__(%d, %d)\n", NT, opt.min_blocks);
    pf("b200qj_pass(chunk* __restrict__ state, const cplx* __restrict__ mats, const u64 chunks_per_state, const i64 mbs,\n"
       "            const u32 tile_shift, const u64 n_work, const __grid_constant__ Remote rm) {\n");
    pf("  extern __shared__ __align__(16) unsigned char smem_[];\n");
    pf("  chunk* tile = reinterpret_cast<chunk*>(smem_);\n");
    pf("  Real* coef = reinterpret_cast<Real*>(smem_ + %zu);\n", tile_bytes);
    pf("  u64* dest_tab = reinterpret_cast<u64*>(smem_ + %zu);\n", tile_bytes + coef_bytes);
    pf("  const int tid = threadIdx.x;\n");
    pf("  b200qj_prep(tid, coef, mats + (i64)blockIdx.y * mbs);\n");
    if (opt.remote) pf("  b200qj_fill_dest_tab(rm, tid, dest_tab);\n");
    pf("  __syncthreads();\n");
    pf("  for (u64 w_ = blockIdx.x; w_ < n_work; w_ += gridDim.x) {\n");
    pf("    const u32 tile_id = (u32)(w_ & ((1ull << tile_shift) - 1ull));\n");
    if (opt.debug_one_tile) {   // measurement only: the work items cycle over 512 tiles that stay in L2 (no DRAM traffic)
      pf("    const u32 tile_id_dbg_ = tile_id; (void)tile_id_dbg_;\n");
      pf("    const u64 cb = %s;\n", deposit("(tile_id & 511u)", ntp, true).c_str());
    } else pf("    const u64 cb = %s;\n", cbexpr.c_str());
    pf("    chunk* g = state + ((u64)blockIdx.y + (w_ >> tile_shift)) * chunks_per_state;\n");
    // L2 prefetch of the NEXT tile of this CTA: its DRAM reads are in flight while this tile is computed, so the
    // first round's loads hit L2 and the memory system never idles during the compute phases (registers and shared
    // memory are full: there is nowhere else to prefetch to)
    bool lines_ok = opt.prefetch != 0;
    for (int j = 0; j < 3 && lines_ok; ++j) lines_ok = int(P.tile_phys[j + VS]) == j + VS;
    if (lines_ok) {
      std::vector<int> lpos;   // chunk bit j + 3 of the tile -> physical chunk bit
      for (int j = 3; j < CB; ++j) lpos.push_back(int(P.tile_phys[j + VS]) - VS);
      const int nlines = 1 << (CB - 3);
      pf("    {\n      const u64 wn_ = w_ + gridDim.x;\n      if (wn_ < n_work) {\n");
      pf("        const u32 tile_id = (u32)(wn_ & ((1ull << tile_shift) - 1ull));\n");
      pf("        const chunk* gn_ = state + ((u64)blockIdx.y + (wn_ >> tile_shift)) * chunks_per_state + ((%s) >> %d);\n",
         cbexpr.c_str(), VS);
      for (int l = 0; l < nlines; l += NT)
        pf("        b200qj_prefetch_l2(gn_ + (%s));\n", deposit(sf("(u32)(tid + %d)", l), lpos, true).c_str());
      pf("      }\n    }\n");
    }
    for (size_t s = 0; s < steps.size(); ++s) {
      call_step(steps[s], "    ");
      if (s + 1 < steps.size() || steps.size() > 1) pf("    __syncthreads();\n");
    }
    pf("  }\n}\n");
    pf("#else\n");
    // TEST-ONLY host entry: steps the same round functions thread by thread
    pf("extern \"C\" void b200qj_emulate(void* state_, const void* mats_, u64 chunks_per_state, i64 mbs, u32 tile_shift,\n"
       "                               u64 n_work, int grid_y, const void* remote_) {\n");
    pf("  chunk* state = (chunk*)state_;\n  const cplx* mats = (const cplx*)mats_;\n");
    pf("  Remote rm; memset(&rm, 0, sizeof rm); if (remote_) memcpy(&rm, remote_, sizeof rm);\n");
    pf("  chunk* tile = (chunk*)aligned_alloc(16, %zu);\n", tile_bytes);
    pf("  Real* coef = (Real*)aligned_alloc(16, %zu);\n", std::max(coef_bytes, size_t(16)));
    pf("  u64* dest_tab = (u64*)aligned_alloc(16, 1280 * 8);\n");
    pf("  for (int by = 0; by < grid_y; ++by) {\n");
    pf("    for (int tid = 0; tid < B200QJ_NT; ++tid) b200qj_prep(tid, coef, mats + (i64)by * mbs);\n");
    if (opt.remote) pf("    for (int tid = 0; tid < B200QJ_NT; ++tid) b200qj_fill_dest_tab(rm, tid, dest_tab);\n");
    pf("    for (u64 w_ = 0; w_ < n_work; ++w_) {\n");
    pf("      const u32 tile_id = (u32)(w_ & ((1ull << tile_shift) - 1ull));\n");
    pf("      const u64 cb = %s;\n", cbexpr.c_str());
    pf("      chunk* g = state + ((u64)by + (w_ >> tile_shift)) * chunks_per_state;\n");
    pf("      memset(tile, 0xff, %zu);\n", tile_bytes);
    for (size_t s = 0; s < steps.size(); ++s) {
      pf("      for (int tid = 0; tid < B200QJ_NT; ++tid) ");
      call_step(steps[s], "");
    }
    pf("    }\n  }\n  free(tile); free(coef); free(dest_tab);\n}\n#endif\n");
    if (stats) {
      *stats = sf("ops %d rounds %d rename %d frame_x %d phys_x %d mat_x %d mat_z %d rho %d generic_diag %d free_phase %d "
                  "coef_reals %d", int(P.n_ops), int(P.n_rounds), n_rename, n_frame_x, n_phys_x, n_mat_x, n_mat_z, n_rho,
                  n_generic_diag, n_free_phase, ncoef);
    }
    return o.str();
  }
};

}  // namespace

bool codegen_supported(const Plan& plan, const b200q_pass_t& P) {
  const int vs = plan.dtype == B200Q_C64 ? 1 : 0;
  if (P.n_rounds == 0) return false;   // dense pass (5..6 targets): b200q_dense_kernel
  if (P.n_bits != P.n_qubits) return false;
  if (int(P.tile_bits) != plan.opt.chunk_bits + vs) return false;
  if (plan.opt.chunk_bits < 11 || plan.opt.chunk_bits > 13) return false;
  return true;
}

std::string codegen_pass(const Plan& plan, const b200q_pass_t& P, const GenOptions& opt, size_t* smem_bytes,
                         std::string* stats) {
  Gen g(plan, P, opt);
  return g.run(smem_bytes, stats);
}

}  // namespace b200q
