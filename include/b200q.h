/* b200q -- C ABI of the B200-native statevector gate-application engine.
 *
 * This is the drop-in boundary for the ONE hot path of TuringQ/deepquantum (v4.5.0 @ 727c44d) that
 * this repository accelerates: the dense statevector gate-application loop driven by
 * QubitCircuit.forward.  The reference has no FFI of its own (it is pure Python on PyTorch); the
 * functions below are what a binding for that path replaces, each cited as file:line relative to
 * the reference tree.  INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - All `state`, `matrices`, `*_dev` pointers are DEVICE pointers owned by the caller (PyTorch
 *     tensors); the library never allocates or frees state memory.  `stream` is a cudaStream_t.
 *   - A state of n qubits is 2^n interleaved complex numbers (complex64: dtype 0, complex128: 1);
 *     `batch` states are contiguous.  Qubit *bit* b is bit b of the flat amplitude index; reference
 *     wire w is bit n-1-w (operation.py:45-55, distributed.py:18).
 *   - Gate matrices are dense row-major 2^k x 2^k complex arrays in device memory (they may be
 *     autograd outputs: the library never reads them on the host).  Matrix index bit j (j = 0 the
 *     least significant) acts on `targets[j]`; the reference enumerates wires with wires[0] as the
 *     MOST significant matrix bit (qmath.py:497-504), i.e. targets[j] = n-1-wires[k-1-j].
 *   - Every function returns 0 on success, a negative B200Q_E* code on invalid arguments, or a
 *     positive cudaError_t.  b200q_last_error() returns the message (thread-local).
 *   - Only sm_100 devices are accepted (b200q_device_check); there is no CPU fallback.
 */
#ifndef B200Q_H
#define B200Q_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200Q_C64 0
#define B200Q_C128 1

#define B200Q_EINVAL (-1)
#define B200Q_EUNSUPPORTED (-2)
#define B200Q_ENODEVICE (-3)

/* gate kinds (what the planner may assume about the matrix *structure*; values stay on device) */
#define B200Q_GATE_MAT 0  /* dense 2^k x 2^k                                 (gate.py get_matrix)   */
#define B200Q_GATE_DIAG 1 /* diagonal 2^k x 2^k: Z,S,T,P,Rz,Rzz,...   (gate.py:737,1634,2295 ...)   */
#define B200Q_GATE_X 2    /* Pauli-X on one target (+controls): X, CNOT, Toffoli (gate.py:841,1934)  */

#define B200Q_GATE_ADJOINT 1 /* flags bit 0: apply the conjugate transpose (Gate.inverse, gate.py:417) */
/* Optional structure hints for 1-target dense gates (by gate CLASS, never by value); the kernel then
 * skips the multiplications by the structural zeros. */
#define B200Q_GATE_REAL 2    /* every entry real: Hadamard, Ry (gate.py:1069, 1538)                     */
#define B200Q_GATE_RXLIKE 4  /* diagonal real, off-diagonal purely imaginary: Rx, Pauli-Y (gate.py:1443) */
#define B200Q_GATE_HADAMARD 8 /* x * [[1, 1], [1, -1]], x real (set together with REAL): Hadamard (gate.py:1069) */
#define B200Q_GATE_ROTATION 16 /* unit-determinant rotation [[c, x], [y, c]], c^2 - x y = 1: with RXLIKE Rx
                                  (gate.py:1443), with REAL Ry (gate.py:1538)                              */

/* Optional hints for 1-target DIAGONAL gates (by gate CLASS): the matrix is exactly diag(1, i^q), q = 1 (S), 2 (Z),
 * 3 (S^dagger), stored in flag bits 5-6.  The specialised pass kernels then apply the gate without arithmetic (a
 * renaming of re / im registers, or a Pauli-frame bit) when its qubit is held in registers (gate.py:1143-1367). */
#define B200Q_GATE_PHASE_SHIFT 5
#define B200Q_GATE_PHASE_MASK (3 << B200Q_GATE_PHASE_SHIFT)
#define B200Q_GATE_PHASE_S (1 << B200Q_GATE_PHASE_SHIFT)
#define B200Q_GATE_PHASE_Z (2 << B200Q_GATE_PHASE_SHIFT)
#define B200Q_GATE_PHASE_SDG (3 << B200Q_GATE_PHASE_SHIFT)

/* flags bit 7: b200q_adjoint_run will be asked for the cotangent of this gate (a parametrised UAnyGate / LatentGate /
 * HamiltonianGate, gate.py:2745-2931).  Only matters for dense gates on 3 or more targets: they then get a pass of
 * their own (like the gates on 5-6 targets), whose reverse step accumulates the full 2^k x 2^k cotangent. */
#define B200Q_GATE_GRAD 128

#define B200Q_MAX_TARGETS 6

typedef struct b200q_gate {
  int32_t kind;                       /* B200Q_GATE_*                                              */
  int32_t n_targets;                  /* k                                                         */
  int32_t targets[B200Q_MAX_TARGETS]; /* bit position of matrix-index bit j                        */
  uint64_t controls;                  /* mask of control bit positions (operation.py:203-219)      */
  int64_t mat_offset;                 /* element offset of this gate's matrix in the matrix buffer */
  int32_t flags;
  int32_t reserved;
} b200q_gate_t;

typedef struct b200q_plan_options {
  int32_t chunk_bits; /* log2 of 16-byte chunks per tile (11..13); 0 = default (12, 64 KiB tiles)  */
  int32_t low_bits;   /* contiguous low index bits kept in every tile; 0 = default                */
  int32_t max_rounds; /* register rounds per pass; 0 = default                                     */
  int32_t fuse;       /* 1 = fuse gates into passes (default), 0 = one gate per pass              */
  int32_t reserved[4];
} b200q_plan_options_t;

typedef struct b200q_plan_stats {
  int32_t n_gates, n_passes, n_rounds, n_ops, n_direct_ops, tile_bits, threads_per_cta, smem_bytes;
} b200q_plan_stats_t;

typedef struct b200q_plan b200q_plan_t;

/* ---- library ---------------------------------------------------------------------------------- */
const char* b200q_version(void);
const char* b200q_last_error(void);
/* 0 if `device` is an sm_100 (B200) GPU, B200Q_ENODEVICE otherwise. */
int b200q_device_check(int device);

/* ---- planner (host only): replaces the per-gate nn.Sequential walk of
 *      QubitCircuit._forward_helper (circuit.py:244-263) by a fused pass schedule --------------- */
int b200q_plan_create(int n_qubits, int dtype, const b200q_gate_t* gates, int n_gates,
                      const b200q_plan_options_t* options, b200q_plan_t** plan_out);
void b200q_plan_destroy(b200q_plan_t* plan);
int b200q_plan_get_stats(const b200q_plan_t* plan, b200q_plan_stats_t* stats_out);
/* Indices (into the `gates` array of b200q_plan_create) of the gates applied by pass `i`, in order of first use; returns
 * their number (writes at most `cap`).  Lets a caller that schedules several plans (the sharded path) see where a
 * plan's passes get sparse. */
int b200q_plan_pass_gate_ids(const b200q_plan_t* plan, int i, int32_t* out, int cap);
/* Number of gates executed by pass `pass_index` (for per-pass reporting). */
int b200q_plan_pass_gates(const b200q_plan_t* plan, int pass_index);
/* Copies the raw pass descriptors (b200q_pass_t, csrc/b200q_program.h) for inspection by tests.
 * `*needed_out` receives the byte size; copies only if buf_size is large enough. */
int b200q_plan_export(const b200q_plan_t* plan, void* buf, size_t buf_size, size_t* needed_out);

/* ---- execution: replaces qmath.evolve_state (qmath.py:485-506) and Gate.op_state_control
 *      (operation.py:203-219) for every gate of the plan, in place ----------------------------- */
int b200q_plan_run(const b200q_plan_t* plan, void* state, const void* matrices, int64_t batch,
                   int64_t matrix_batch_stride, void* stream);
/* Run passes [first, last) only (per-pass timing in bench.py, overlap with exchanges). */
int b200q_plan_run_range(const b200q_plan_t* plan, int first_pass, int last_pass, void* state,
                         const void* matrices, int64_t batch, int64_t matrix_batch_stride, void* stream);
/* Sharded path: the plan of a local segment FUSED with the exchange that follows it (replaces the exchanges of
 * dist_one_targ_gate / dist_many_targ_gate / dist_swap_gate, distributed.py:57-202, and comm_exchange_arrays,
 * communication.py:58-91).  The last pass of the plan does not write the local shard: every 16-byte chunk goes
 * straight into the receive buffer of the rank that owns it after the exchange (peer_buffers[r], device pointers
 * valid on THIS device: peer mappings over NVLink for r != rank, the local receive buffer for r == rank).
 * The exchange is a permutation of the bits of the DISTRIBUTED index (n_local + log2(n_ranks) bits, the top ones
 * being the rank): perm[j] = destination position of bit j; NULL = the block transpose that swaps the rank bits
 * with the top local bits.  complex64: perm[0] must be 0.  The caller synchronises the ranks afterwards (all
 * stores are complete when the kernels have finished on every rank) and swaps the roles of shard and buffer. */
int b200q_plan_run_exchange(const b200q_plan_t* plan, void* state, const void* matrices, void* const* peer_buffers,
                            int n_ranks, int rank, const uint8_t* perm, void* stream);
/* ---- per-pass kernel specialisation (same path as above, compiled at run time) ---------------------------------
 * The generic tile kernel interprets the op list of a pass; for states of >= 20 qubits (B200Q_JIT_MIN_QUBITS) every
 * pass of a plan is instead turned into CUDA source with the op sequence, register slots, control masks and index
 * arithmetic as constants, compiled with NVRTC for sm_100a (cubins cached next to the library) and launched by
 * b200q_plan_run* in place of the generic kernel.  B200Q_JIT=0 disables it.
 *   b200q_plan_codegen:    the generated source of one pass (`remote` != 0: the fused-exchange variant); the needed
 *                          buffer size (with the terminating 0) goes to *needed_out, the dynamic shared memory of the
 *                          kernel to *smem_bytes_out.  Host only.
 *   b200q_plan_compile:    compile every pass now (needs no GPU; `threads` <= 0: all host cores); returns the number
 *                          of passes that have a specialised kernel, or a negative error.
 *   b200q_plan_jit_status: how many passes of the plan run specialised kernels / failed to compile. */
int b200q_plan_codegen(const b200q_plan_t* plan, int pass_index, int remote, char* buf, size_t buf_size,
                       size_t* needed_out, size_t* smem_bytes_out);
int b200q_plan_compile(b200q_plan_t* plan, int threads, int with_exchange_variant);
int b200q_plan_jit_status(const b200q_plan_t* plan, int32_t* n_specialised_out, int32_t* n_failed_out);

/* One gate, no plan object: the direct counterpart of
 *   evolve_state(state, matrix, nqudit, wires)            (qmath.py:485)      controls == NULL
 *   Gate.op_state_control(x, matrix)                      (operation.py:203)  controls != NULL  */
int b200q_apply_gate(void* state, int n_qubits, int dtype, int kind, const void* matrix, const int32_t* targets,
                     int n_targets, const int32_t* controls, int n_controls, int adjoint, int64_t batch,
                     int64_t matrix_batch_stride, void* stream);

/* Dense gate on 4..6 targets, complex64, on the tensor cores (tcgen05.mma, accumulator in tensor memory, FP32
 * accuracy by the 3-product TF32 split): evolve_state / op_state_control for UAnyGate-style blocks (gate.py:2745-2790).
 * `matrix`: dense 2^k x 2^k row-major complex64 on the device; `controls`: mask of control bits; needs >= 12 qubits.
 * b200q_plan_run uses it for the dense passes (5 and 6 targets) of complex64 plans; B200Q_DENSE_TC=0 selects the
 * CUDA-core contraction instead. */
int b200q_dense_tc_apply(void* state, int n_qubits, const void* matrix, const int32_t* targets, int n_targets,
                         uint64_t controls, int adjoint, void* stream);

/* ---- reductions: qmath.expectation (qmath.py:830-860), inner_product_dist (distributed.py:288) */
/* out_dev[b] = sum_i |state[b][i]|^2 */
int b200q_norm2(const void* state, int n_qubits, int dtype, int64_t batch, double* out_dev, void* stream);
/* out_dev[2b], out_dev[2b+1] = Re, Im of <bra[b]|ket[b]> */
int b200q_inner_product(const void* bra, const void* ket, int n_qubits, int dtype, int64_t batch, double* out_dev,
                        void* stream);
/* out_dev[b*n_masks + m] = sum_i |state[b][i]|^2 * (-1)^popcount(i & masks[m])   (Z-string observables,
 * layer.py:127-165 with basis 'z').  `masks_dev` is a device array of n_masks uint64.
 * `index_offset` is OR-ed into i (rank offset of a shard, distributed.py:18). */
int b200q_expectation_z(const void* state, int n_qubits, int dtype, int64_t batch, const uint64_t* masks_dev,
                        int n_masks, uint64_t index_offset, double* out_dev, void* stream);
/* lambda[i] = state[i] * sum_m w[m] * (-1)^popcount(i & masks[m]): the seed of the adjoint backward pass
 * (adjoint.py:47-56) for a weighted sum of Z-string observables.  weights_dev: batch x n_masks doubles. */
int b200q_apply_z_weights(const void* state, void* lambda_out, int n_qubits, int dtype, int64_t batch,
                          const uint64_t* masks_dev, const double* weights_dev, int n_masks,
                          uint64_t index_offset, void* stream);
/* state[b][i] = (i == basis_index) ? 1 : 0    (QubitState 'zeros', state.py:31-44) */
int b200q_init_basis(void* state, int n_qubits, int dtype, int64_t batch, uint64_t basis_index, void* stream);

/* ---- adjoint differentiation (adjoint.py:47-83) ----------------------------------------------
 * Reverse sweep over the gates of `plan`: for g = last .. first
 *   psi <- U_g^dagger psi                      (psi enters as the final state, leaves as the initial one)
 *   grad_out[g][r][c] += sum_rest lambda[r,rest] * conj(psi[c,rest])        (PyTorch cotangent of U_g)
 *   lambda <- U_g^dagger lambda                (lambda enters as dL/d(final state), leaves as dL/d(initial))
 * `grad_out` is a ZEROED device buffer of complex128 laid out like the matrix buffer (same element
 * offsets, always double precision).  need_grad_host[g] == 0 skips gate g (NULL: accumulate all gates
 * the sweep supports: 1- and 2-target dense, diagonal, and every dense gate that has a pass of its own -- 5-6
 * targets, or 3-4 targets planned with B200Q_GATE_GRAD).  batch = 1. */
int b200q_adjoint_run(const b200q_plan_t* plan, void* psi, void* lambda, const void* matrices, void* grad_out,
                      const uint8_t* need_grad_host, void* stream);

/* ---- qudit (Fock tensor) path: evolve_state(..., qudit = cutoff) as called by
 *      photonic/operation.py:142-146.  `state` is [batch, d^n]; digit 0 is the MOST significant
 *      (mode 0), `modes[0]` is the most significant digit of the matrix index (reference order). */
int b200q_qudit_apply(void* state, int n_modes, int d, int dtype, const void* matrix, const int32_t* modes,
                      int n_targets, int64_t batch, void* stream);

/* The same call for gate CLASSES whose Fock matrix has a known block structure (decided by the class, never from the
 * values; entries outside the structure are not read): the kernel then keeps one block of one fibre in registers and
 * needs no shared-memory staging of amplitudes.  B200Q_QUDIT_GENERAL, cutoffs above 16 and arity mismatches fall back
 * to b200q_qudit_apply.
 *   DIAG        diagonal: phase shifter, Kerr, cross-Kerr                       (photonic/gate.py:135-199, 2628-2680)
 *   DENSE1      any one-mode gate: squeezer, displacement                       (photonic/gate.py:1015-1154, 1336-1489)
 *   NUMBER      two-mode, conserves the photon number of its modes: the beamsplitter family, MZI
 *                                                                                (photonic/gate.py:202-1012)
 *   DIFFERENCE  two-mode, conserves the photon-number difference: two-mode squeezing  (photonic/gate.py:1157-1333) */
#define B200Q_QUDIT_GENERAL 0
#define B200Q_QUDIT_DIAG 1
#define B200Q_QUDIT_DENSE1 2
#define B200Q_QUDIT_NUMBER 3
#define B200Q_QUDIT_DIFFERENCE 4
int b200q_qudit_apply_structured(void* state, int n_modes, int d, int dtype, const void* matrix, const int32_t* modes,
                                 int n_targets, int structure, int64_t batch, void* stream);

/* A two-mode structured gate together with the structured one-mode (or diagonal) gates that directly precede / follow it
 * on its two modes -- squeezers in front of a beamsplitter, the phase shifter + beamsplitter pairs of an MZI mesh
 * (photonic/circuit.py applies them one by one, each a pass over the state) -- in ONE pass: `ops` (1..4, applied in
 * order) all act inside `tile_modes` (two modes); `matrix` is the device pointer of the gate's dense d^k x d^k matrix. */
typedef struct b200q_qudit_op {
  int32_t n_targets;
  int32_t modes[2];     /* modes[0] = most significant digit of the matrix index */
  int32_t structure;    /* B200Q_QUDIT_DIAG .. B200Q_QUDIT_DIFFERENCE */
  uint64_t matrix;      /* device pointer */
} b200q_qudit_op_t;
int b200q_qudit_apply_group(void* state, int n_modes, int d, int dtype, const int32_t* tile_modes,
                            const b200q_qudit_op_t* ops, int n_ops, int64_t batch, void* stream);

/* Fock transformation matrices of a whole gate class in one launch (the reference runs Python-level recurrences over the
 * cutoff per gate: photonic/gate.py:347-374 beamsplitter family, 1091-1114 squeezer; arXiv:2004.11002 Eq. 74-75, 51-52).
 * `mixing`: n_gates x 2 x 2 complex128 mode-mixing matrices on the device; `r_theta`: n_gates x 2 doubles.
 * `out`: n_gates x d^4 (index m, n, p, q) resp. n_gates x d^2 complex of `dtype`.  Cutoff <= 16 (beamsplitter), <= 64. */
int b200q_fock_bs_matrix(const void* mixing, int n_gates, int d, int dtype, void* out, void* stream);
int b200q_fock_squeezing_matrix(const double* r_theta, int n_gates, int d, int dtype, void* out, void* stream);

/* Fused Fock pass: `n_gates` consecutive evolve_state(..., qudit = cutoff) calls of the photonic tensor path
 * (photonic/circuit.py:405-431 applies them one by one) whose modes all lie in `tile_modes` (ascending, at most
 * cutoff^n_tile = 12 288 amplitudes; the planner keeps the last mode in the tile so that global accesses are runs of
 * `cutoff` amplitudes), applied with ONE read and ONE write of the state.  `matrices`: device buffer holding the
 * dense cutoff^k x cutoff^k matrices (k = 1, 2) at `mat_offset` (elements); `modes[0]` is the most significant digit
 * of the matrix index, as in b200q_qudit_apply. */
#define B200Q_QUDIT_FUSED_MAX_GATES 16
typedef struct b200q_qudit_gate {
  int32_t n_targets;
  int32_t modes[2];
  int32_t reserved;
  int64_t mat_offset;
} b200q_qudit_gate_t;
int b200q_qudit_fused(void* state, int n_modes, int d, int dtype, const int32_t* tile_modes, int n_tile,
                      const b200q_qudit_gate_t* gates, int n_gates, const void* matrices, int64_t batch, void* stream);

/* ---- sampling: qmath.measure + block_sample (qmath.py:543-638), the step after the path -----------
 * Inverse-CDF sampling in two levels.  The caller (Python) draws `shots` uniforms, searches them in the
 * prefix sum of the block masses and passes, per shot, the block and the residual mass inside it.
 *   b200q_block_mass:     mass_dev[b * n_blocks + k] = sum of |a|^2 over block k (2^block_bits amplitudes) of
 *                         state b; double accumulation, deterministic (no atomics); one read of the state.
 *   b200q_sample_blocks:  out_index_dev[s] = first amplitude index i of block block_idx_dev[s] whose running
 *                         mass (in index order, double) exceeds residual_dev[s]; one warp per shot.
 *   b200q_marginal_probs: out_dev[j] = sum of |a_i|^2 over all i with (i & mask) == keys_sorted_dev[j]
 *                         (exact marginal probability of the sampled outcomes on a wire subset, the
 *                         `with_prob=True` branch of qmath.py:629-632); n_keys <= 2048. */
int b200q_block_mass(const void* state, int n_qubits, int dtype, int64_t batch, int block_bits, double* mass_dev,
                     void* stream);
int b200q_sample_blocks(const void* state, int n_qubits, int dtype, int block_bits, const int64_t* block_idx_dev,
                        const double* residual_dev, int64_t shots, int64_t* out_index_dev, void* stream);
int b200q_marginal_probs(const void* state, int n_qubits, int dtype, uint64_t mask, const uint64_t* keys_sorted_dev,
                         int n_keys, double* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200Q_H */
