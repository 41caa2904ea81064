// Dense gate on 5 or 6 targets (complex64) on the 5th-generation tensor cores: tcgen05.mma with the accumulator
// in tensor memory.
//
// Replaces, for UAnyGate / get_unitary-style blocks (reference gate.py:2745-2790 -> qmath.evolve_state,
// qmath.py:485-506), the CUDA-core contraction of b200q_dense_kernel.  The block is always treated as a 6-target
// block (a 5-target gate is padded with one untouched bit, U' = I (x) U), so the complex 64 x 64 contraction is the
// real 128 x 128 x N GEMM
//        [ Yr ]   [ Ur  -Ui ] [ Xr ]
//        [ Yi ] = [ Ui   Ur ] [ Xi ]          (N = groups of 64 amplitudes)
// with M = 128 = one tcgen05.mma tile.  FP32 accuracy from TF32 tensor cores by the 3-product split
// A = A_hi + A_lo, B = B_hi + B_lo (hi = upper 19 bits, lo = exact remainder): D = A_lo B_hi + A_hi B_lo + A_hi B_hi
// (measured error against the FP32 CUDA-core kernel: see tools/dense_tc_bench.py / DESIGN.md).
//
// One CTA of 256 threads per SM, persistent over tiles of kN = 64 groups:
//   all threads   gather the tile's amplitudes from global memory (strided partner amplitudes of the 6 block bits),
//                 split them and store B_hi / B_lo in the canonical K-major no-swizzle operand layout
//   thread 0      issues 3 x 16 tcgen05.mma (M 128, N kN, K 8, kind::tf32; A and B from shared-memory descriptors,
//                 D in TMEM), then tcgen05.commit -> mbarrier
//   all threads   wait on the mbarrier, tcgen05.ld their accumulator row (warp w: TMEM lanes 32 (w & 3) .., columns 32 (w >> 2) ..),
//                 transpose through shared memory and scatter the results back with 8-byte stores
// The gate matrix (A operand, hi and lo: 2 x 64 KiB) is staged once per CTA.
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

constexpr int kTcThreads = 256;        // 8 warps: warp w reads TMEM lanes 32 (w & 3) .., columns 32 (w >> 2) ..
constexpr int kN = 64;                 // groups per tile = N of the MMA = TMEM columns
constexpr int kK = 128;                // real-ified block dimension
constexpr int kGS = kTcThreads / 64;     // groups gathered per step by the CTA
constexpr int kPT = kN / kGS;          // amplitudes per thread and tile
constexpr int kStage = kK + 4;          // row stride of the epilogue staging: lanes (n & 7, j & 3) -> 32 distinct banks
constexpr uint32_t kLBO = 128;         // bytes between the two 16-byte K chunks of one MMA step (next core matrix)
constexpr uint32_t kSBO = (kK / 4) * 128;   // bytes between groups of 8 rows

struct TcArgs {
  uint64_t ctrl;          // controls (physical bits)
  int32_t n_qubits, k, adjoint;
  uint8_t bbit[8];        // physical bit of block-index bit j (j < k: the gate's targets; j >= k: padding bits)
  uint8_t sorted[8];
};

// canonical K-major, no swizzle: element (row r, k) of an operand with kK columns, in floats
__device__ __forceinline__ uint32_t canon(uint32_t r, uint32_t k) {
  return ((r & 7u) * 16u + (k >> 2) * kLBO + (r >> 3) * kSBO + (k & 3u) * 4u) >> 2;
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  // start address [0,14) (>>4), leading byte offset [16,30) (>>4), stride byte offset [32,46) (>>4),
  // descriptor version 1 for sm_100 at [46,48), layout type SWIZZLE_NONE = 0 at [61,64)
  return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(kLBO >> 4) << 16) | (uint64_t(kSBO >> 4) << 32) | (uint64_t(1) << 46);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  // hi = x rounded to TF32 (10 mantissa bits, low 13 bits zero), lo = exact remainder (the tensor core then drops
  // the low 13 bits of lo: 2^-22 relative to x)
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 1)
b200q_dense_tc_kernel(float2* __restrict__ state, const float2* __restrict__ mat, const TcArgs A, uint64_t n_tiles) {
  extern __shared__ __align__(1024) unsigned char tsm[];
  float* a_hi = reinterpret_cast<float*>(tsm);                  // [128][128] canonical
  float* a_lo = a_hi + kK * kK;
  float* b_hi = a_lo + kK * kK;                                 // [kN][128] canonical
  float* b_lo = b_hi + kN * kK;
  float* stage = b_hi;                                          // epilogue staging [kN][128] (after the MMAs)
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base_sh;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int k = A.k, D = 1 << k;

  // ---- A operand: real-ified U' = I (x) U (adjoint folded), hi / lo, once per CTA
  for (int e = tid; e < kK * kK; e += kTcThreads) {
    const int m = e >> 7, kk = e & 127;
    const int r = m & 63, c = kk & 63;                    // complex row / column of U'
    float ur = 0.f, ui = 0.f;
    if ((r >> k) == (c >> k)) {
      const int rl = r & (D - 1), cl = c & (D - 1);
      float2 v = A.adjoint ? mat[cl * D + rl] : mat[rl * D + cl];
      if (A.adjoint) v.y = -v.y;
      ur = v.x; ui = v.y;
    }
    // [ Ur -Ui ; Ui Ur ]
    const float val = (m < 64) ? ((kk < 64) ? ur : -ui) : ((kk < 64) ? ui : ur);
    float hi, lo;
    split_tf32(val, hi, lo);
    a_hi[canon(m, kk)] = hi;
    a_lo[canon(m, kk)] = lo;
  }
  const uint32_t mbar_addr = (uint32_t)__cvta_generic_to_shared(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_addr), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_sh)),
                 "n"(kN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem_d = tmem_base_sh;

  // instruction descriptor: D F32, A / B TF32, both K-major, N = kN, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kN >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  const uint32_t a_hi_s = (uint32_t)__cvta_generic_to_shared(a_hi), a_lo_s = (uint32_t)__cvta_generic_to_shared(a_lo);
  const uint32_t b_hi_s = (uint32_t)__cvta_generic_to_shared(b_hi), b_lo_s = (uint32_t)__cvta_generic_to_shared(b_lo);

  // Thread -> (group n, amplitude j) of a tile: lane = (n & 7) + 8 (j & 3), so that a warp's stores into the canonical
  // operand layout (8 rows x 16 bytes per core matrix) hit 32 distinct banks; warp w and iteration i supply the rest:
  // n >> 3 = i & 7, j >> 2 = 2 w + (i >> 3).  Index expansion (zero bits inserted at the 6 block positions) is a bit
  // deposit, hence additive over disjoint bit groups: per-thread constants for n and j, one expansion per tile.
  const int lane = tid & 31;
  const int nl = lane & 7, jl = lane >> 3;
  auto expand = [&](uint64_t g) {
#pragma unroll
    for (int q = 0; q < 6; ++q) g = ((g >> A.sorted[q]) << (A.sorted[q] + 1)) | (g & ((1ull << A.sorted[q]) - 1ull));
    return g;
  };
  auto jdep = [&](int jj) {
    uint64_t o = 0;
    for (int q = 0; q < 6; ++q)
      if ((jj >> q) & 1) o |= 1ull << A.bbit[q];
    return o;
  };
  uint64_t e_nh[8];
#pragma unroll
  for (int h = 0; h < 8; ++h) e_nh[h] = expand(uint64_t(h * 8 + nl));
  const uint64_t d_j[2] = {jdep((2 * warp) * 4 + jl), jdep((2 * warp + 1) * 4 + jl)};
  uint32_t phase = 0;
  for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    // ---- gather kN groups x 64 amplitudes -> B_hi / B_lo
    const uint64_t tbase = expand(t * kN);
    uint64_t gaddr[kPT];
    bool act[kPT];
#pragma unroll
    for (int i = 0; i < kPT; ++i) {
      const uint64_t base = tbase | e_nh[i & 7];
      gaddr[i] = base | d_j[i >> 3];
      act[i] = (base & A.ctrl) == A.ctrl;
    }
    float2 v[kPT];
#pragma unroll
    for (int i = 0; i < kPT; ++i) v[i] = state[gaddr[i]];
#pragma unroll
    for (int i = 0; i < kPT; ++i) {
      const int n = (i & 7) * 8 + nl, jj = (2 * warp + (i >> 3)) * 4 + jl;
      float hr, lr, hi_, li;
      split_tf32(v[i].x, hr, lr);
      split_tf32(v[i].y, hi_, li);
      b_hi[canon(n, jj)] = hr; b_lo[canon(n, jj)] = lr;
      b_hi[canon(n, 64 + jj)] = hi_; b_lo[canon(n, 64 + jj)] = li;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- MMA: D = A_lo B_hi + A_hi B_lo + A_hi B_hi, 16 K-steps of 8 each
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      uint32_t acc = 0;
      const uint32_t pa[3] = {a_lo_s, a_hi_s, a_hi_s}, pb[3] = {b_hi_s, b_lo_s, b_hi_s};
#pragma unroll
      for (int p = 0; p < 3; ++p) {
#pragma unroll
        for (int ks = 0; ks < kK / 8; ++ks) {
          mma_tf32(tmem_d, smem_desc(pa[p] + ks * 2 * kLBO), smem_desc(pb[p] + ks * 2 * kLBO), idesc, acc);
          acc = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar_addr) : "memory");
    }
    // ---- wait for the accumulator
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(mbar_addr), "r"(phase)
            : "memory");
      }
      phase ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    // ---- epilogue: row m = tid of D (lane 32 warp + lane), columns 0 .. kN-1
    uint32_t r[32];
    const int lane_q = warp & 3, col0 = (warp >> 2) * 32, row = lane_q * 32 + (tid & 31);
    const uint32_t taddr = tmem_d + (uint32_t(lane_q * 32) << 16) + uint32_t(col0);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // stage[n][m]: (the MMAs have completed, B is free)
#pragma unroll
    for (int n = 0; n < 32; ++n) stage[(col0 + n) * kStage + row] = __uint_as_float(r[n]);
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPT; ++i) {
      const int n = (i & 7) * 8 + nl, jj = (2 * warp + (i >> 3)) * 4 + jl;
      if (act[i]) {
        float2 y;
        y.x = stage[n * kStage + jj];
        y.y = stage[n * kStage + 64 + jj];
        state[gaddr[i]] = y;
      }
    }
    __syncthreads();   // staging (= B) is rewritten by the next tile
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(kN));
}

static_assert(kN * kStage <= 2 * kN * kK, "the staging reuses the B operands");
static_assert(kN == 64 && kTcThreads == 256 && kPT == 16, "epilogue mapping: 4 lane quarters x 2 column halves of 32");

}  // namespace

extern "C" int b200q_dense_tc_apply(void* state, int n_qubits, const void* matrix, const int32_t* targets, int n_targets,
                                    uint64_t controls, int adjoint, void* stream) {
  if (!state || !matrix || !targets) return set_err(B200Q_EINVAL, "null argument");
  if (n_targets < 4 || n_targets > 6) return set_err(B200Q_EUNSUPPORTED, "the tensor-core block takes 4 to 6 targets");
  if (n_qubits < 12 || n_qubits > 38) return set_err(B200Q_EUNSUPPORTED, "the tensor-core block needs at least 12 qubits");
  TcArgs A;
  std::memset(&A, 0, sizeof A);
  A.ctrl = controls;
  A.n_qubits = n_qubits;
  A.k = n_targets;
  A.adjoint = adjoint ? 1 : 0;
  uint64_t used = controls;
  for (int j = 0; j < n_targets; ++j) {
    if (targets[j] < 0 || targets[j] >= n_qubits || (used >> targets[j] & 1)) return set_err(B200Q_EINVAL, "bad targets");
    used |= 1ull << targets[j];
    A.bbit[j] = (uint8_t)targets[j];
  }
  int nb = n_targets;
  for (int b = 0; b < n_qubits && nb < 6; ++b)       // padding bits: the lowest bits the gate does not touch
    if (!(used >> b & 1)) { A.bbit[nb++] = (uint8_t)b; used |= 1ull << b; }
  if (nb < 6) return set_err(B200Q_EUNSUPPORTED, "not enough free bits to pad the block to 6");
  for (int j = 0; j < 6; ++j) A.sorted[j] = A.bbit[j];
  std::sort(A.sorted, A.sorted + 6);
  const size_t smem = size_t(2 * kK * kK + 2 * kN * kK) * sizeof(float) + 1024;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    const int rc = cuda_err(cudaFuncSetAttribute(b200q_dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(dense tc)");
    if (rc) return rc;
    attr_set[dev] = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t n_tiles = (1ull << (n_qubits - 6)) / kN;
  const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)sms);
  b200q_dense_tc_kernel<<<grid, kTcThreads, smem, (cudaStream_t)stream>>>((float2*)state, (const float2*)matrix, A, n_tiles);
  return cuda_err(cudaGetLastError(), "dense tensor-core launch");
}
