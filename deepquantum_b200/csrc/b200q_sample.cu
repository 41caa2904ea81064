// Sampling kernels: the step right after the gate-application path in every reference example
// (SURVEY.md section 8f rank 1).
//
// Replaces qmath.measure + block_sample (reference qmath.py:543-638): the reference materialises |a|^2 for
// the whole state (one extra 2^n float tensor), runs torch.multinomial over blocks of 2^24 probabilities and
// builds a Python Counter.  Here the state is read ONCE (block masses, HBM-bound), the inverse CDF is searched
// over the 2^(n-12) block masses, and each shot then scans only its own 32 KiB block with one warp.
//   b200q_block_mass      sum |a|^2 over blocks of 2^block_bits amplitudes (double accumulation, no atomics)
//   b200q_sample_blocks   per shot: first index i of the block with  sum_{j<=i} |a_j|^2 > residual
//   b200q_marginal_probs  exact marginal probability of a sorted list of outcomes on a wire subset
#include <cuda_runtime.h>

#include <string>

#include "../../include/b200q.h"

namespace b200q {
int set_err(int code, const std::string& msg);
int cuda_err(cudaError_t e, const char* what);
}  // namespace b200q
using b200q::cuda_err;
using b200q::set_err;

namespace {

template <typename Real> struct amp2 { Real x, y; };

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid.x = blocks of the state, grid.y = batch; 256 threads; each thread strides over the block with 16-byte loads
template <typename Real>
__global__ void __launch_bounds__(256)
block_mass_kernel(const amp2<Real>* __restrict__ st, uint64_t n_amps, int block_bits, double* __restrict__ mass) {
  const uint64_t blk = blockIdx.x;
  const uint64_t n_blocks = gridDim.x;
  const amp2<Real>* s = st + uint64_t(blockIdx.y) * n_amps + (blk << block_bits);
  const uint32_t len = 1u << block_bits;
  double acc = 0.0;
  if (sizeof(Real) == 4) {
    const float4* p = reinterpret_cast<const float4*>(s);   // two complex64 amplitudes
    for (uint32_t i = threadIdx.x; i < (len >> 1); i += blockDim.x) {
      const float4 v = p[i];
      acc += double(v.x) * double(v.x) + double(v.y) * double(v.y) + double(v.z) * double(v.z) + double(v.w) * double(v.w);
    }
    if (len == 1 && threadIdx.x == 0) acc = double(s[0].x) * double(s[0].x) + double(s[0].y) * double(s[0].y);
  } else {
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
      const amp2<Real> v = s[i];
      acc += double(v.x) * double(v.x) + double(v.y) * double(v.y);
    }
  }
  __shared__ double sh[8];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) mass[uint64_t(blockIdx.y) * n_blocks + blk] = v;
  }
}

// one warp per shot
template <typename Real>
__global__ void __launch_bounds__(256)
sample_blocks_kernel(const amp2<Real>* __restrict__ st, int block_bits, const int64_t* __restrict__ block_idx,
                     const double* __restrict__ residual, int64_t shots, int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t shot = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (shot >= shots) return;
  const uint64_t blk = uint64_t(block_idx[shot]);
  const double r = residual[shot];
  const amp2<Real>* s = st + (blk << block_bits);
  const uint32_t len = 1u << block_bits;
  double run = 0.0;
  int64_t found = -1, last_nz = -1;
  for (uint32_t i0 = 0; i0 < len && found < 0; i0 += 32) {
    const uint32_t i = i0 + lane;
    double p = 0.0;
    if (i < len) {
      const amp2<Real> v = s[i];
      p = double(v.x) * double(v.x) + double(v.y) * double(v.y);
    }
    double incl = p;   // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, p > 0.0 && run + incl > r);
    const unsigned nz = __ballot_sync(0xffffffffu, p > 0.0);
    if (hit) found = int64_t(i0) + (__ffs(hit) - 1);
    if (nz) last_nz = int64_t(i0) + (31 - __clz(nz));
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (found < 0) found = last_nz >= 0 ? last_nz : int64_t(len) - 1;   // residual beyond the scan total (rounding)
  if (lane == 0) out[shot] = int64_t(blk << block_bits) + found;
}

// keys: sorted values of (index & mask); out[j] += |a_i|^2 for every i with (i & mask) == keys[j]
template <typename Real>
__global__ void __launch_bounds__(256)
marginal_probs_kernel(const amp2<Real>* __restrict__ st, uint64_t n_amps, uint64_t mask,
                      const uint64_t* __restrict__ keys, int n_keys, double* __restrict__ out) {
  extern __shared__ unsigned char mp_raw[];
  uint64_t* sk = reinterpret_cast<uint64_t*>(mp_raw);
  double* acc = reinterpret_cast<double*>(sk + n_keys);
  for (int k = threadIdx.x; k < n_keys; k += blockDim.x) { sk[k] = keys[k]; acc[k] = 0.0; }
  __syncthreads();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_amps; i += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t v = i & mask;
    int lo = 0, hi = n_keys - 1, j = -1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const uint64_t km = sk[mid];
      if (km == v) { j = mid; break; }
      if (km < v) lo = mid + 1; else hi = mid - 1;
    }
    if (j >= 0) {
      const amp2<Real> a = st[i];
      atomicAdd(acc + j, double(a.x) * double(a.x) + double(a.y) * double(a.y));
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_keys; k += blockDim.x)
    if (acc[k] != 0.0) atomicAdd(out + k, acc[k]);
}

int check_args(const void* state, int n_qubits, int dtype) {
  if (!state) return set_err(B200Q_EINVAL, "null state pointer");
  if (n_qubits < 1 || n_qubits > 38) return set_err(B200Q_EINVAL, "n_qubits out of range");
  if (dtype != B200Q_C64 && dtype != B200Q_C128) return set_err(B200Q_EINVAL, "bad dtype");
  return 0;
}

}  // namespace

extern "C" {

int b200q_block_mass(const void* state, int n_qubits, int dtype, int64_t batch, int block_bits, double* mass_dev,
                     void* stream) {
  int rc = check_args(state, n_qubits, dtype);
  if (rc) return rc;
  if (!mass_dev || batch < 1 || batch > 65535) return set_err(B200Q_EINVAL, "bad arguments");
  if (block_bits < 0 || block_bits > n_qubits || block_bits > 20) return set_err(B200Q_EINVAL, "bad block_bits");
  const uint64_t n = 1ull << n_qubits;
  const uint64_t nb = n >> block_bits;
  if (nb > 0x7fffffffull) return set_err(B200Q_EUNSUPPORTED, "too many blocks");
  dim3 grid((unsigned)nb, (unsigned)batch);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == B200Q_C64) block_mass_kernel<float><<<grid, 256, 0, s>>>((const amp2<float>*)state, n, block_bits, mass_dev);
  else block_mass_kernel<double><<<grid, 256, 0, s>>>((const amp2<double>*)state, n, block_bits, mass_dev);
  return cuda_err(cudaGetLastError(), "block_mass launch");
}

int b200q_sample_blocks(const void* state, int n_qubits, int dtype, int block_bits, const int64_t* block_idx_dev,
                        const double* residual_dev, int64_t shots, int64_t* out_index_dev, void* stream) {
  int rc = check_args(state, n_qubits, dtype);
  if (rc) return rc;
  if (!block_idx_dev || !residual_dev || !out_index_dev || shots < 1) return set_err(B200Q_EINVAL, "bad arguments");
  if (block_bits < 0 || block_bits > n_qubits || block_bits > 20) return set_err(B200Q_EINVAL, "bad block_bits");
  const unsigned grid = (unsigned)((shots + 7) / 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == B200Q_C64)
    sample_blocks_kernel<float><<<grid, 256, 0, s>>>((const amp2<float>*)state, block_bits, block_idx_dev, residual_dev,
                                                     shots, out_index_dev);
  else
    sample_blocks_kernel<double><<<grid, 256, 0, s>>>((const amp2<double>*)state, block_bits, block_idx_dev,
                                                      residual_dev, shots, out_index_dev);
  return cuda_err(cudaGetLastError(), "sample_blocks launch");
}

int b200q_marginal_probs(const void* state, int n_qubits, int dtype, uint64_t mask, const uint64_t* keys_sorted_dev,
                         int n_keys, double* out_dev, void* stream) {
  int rc = check_args(state, n_qubits, dtype);
  if (rc) return rc;
  if (!keys_sorted_dev || !out_dev || n_keys < 1 || n_keys > 2048) return set_err(B200Q_EINVAL, "n_keys must be in 1..2048");
  const uint64_t n = 1ull << n_qubits;
  cudaStream_t s = (cudaStream_t)stream;
  rc = cuda_err(cudaMemsetAsync(out_dev, 0, sizeof(double) * n_keys, s), "memset");
  if (rc) return rc;
  const uint64_t want = (n + 256 * 16 - 1) / (256 * 16);
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 148ull * 8 ? 148ull * 8 : want));
  const size_t sm = size_t(n_keys) * 16;
  if (dtype == B200Q_C64)
    marginal_probs_kernel<float><<<grid, 256, sm, s>>>((const amp2<float>*)state, n, mask, keys_sorted_dev, n_keys, out_dev);
  else
    marginal_probs_kernel<double><<<grid, 256, sm, s>>>((const amp2<double>*)state, n, mask, keys_sorted_dev, n_keys,
                                                        out_dev);
  return cuda_err(cudaGetLastError(), "marginal_probs launch");
}

}  // extern "C"
