// Pass / round / op descriptors shared by the host planner and the sm_100a tile kernel.
//
// One *pass* = one kernel launch that reads every amplitude of the (local) statevector once and
// writes it once.  A pass owns a set of `tile_bits` physical index bits (the *tile bits*): every CTA
// stages the 2^tile_bits amplitudes that differ only in those bits (64 KiB by default) and applies a
// whole group of gates to them before writing back -- the fused replacement for one
// `permute -> reshape(copy) -> mm` per gate in the reference (qmath.py:497-506).
//
// Inside a pass the tile is processed in *rounds*.  In a round every thread holds 16 chunks of 16
// bytes (32 complex64 or 16 complex128 amplitudes) in registers: the amplitudes that differ only in
// the round's *register slots* (5 index bits for complex64 -- bit 0 is always slot 0 because a
// 16-byte chunk holds amplitudes 2m and 2m+1 -- 4 bits for complex128).  All ops of the round whose
// targets are register slots are applied in registers; between rounds the tile is transposed through
// shared memory.  The first round reads straight from global memory, the last one writes straight
// back, so a pass whose gates fit one round never touches shared memory.
//
// complex64 chunks have two formats: AoS (re0, im0, re1, im1) -- the caller-visible layout of the
// state -- and SoA (re0, re1, im0, im1), which a multi-pass plan uses for the state BETWEEN its
// passes and always inside shared memory, so that a 128-bit load lands directly in the register
// pairs that the packed FFMA2 arithmetic (two amplitudes per instruction) consumes.
#pragma once
#include <stdint.h>

#define B200Q_REG_CHUNK_BITS 4  /* 16 chunks of 16 B per thread */
#define B200Q_MAX_TILE_BITS 14  /* amplitude bits per tile (complex64, 13 chunk bits) */
#define B200Q_MAX_ROUNDS 8
#define B200Q_MAX_OPS 48
#define B200Q_POOL_MAX 512      /* complex elements of gate matrices staged in shared memory */
#define B200Q_MAX_QUBITS 40
#define B200Q_MATK_MAX 4        /* dense k-target ops applied inside a tile */

enum {
  B200Q_OP_MAT1 = 0,  // dense 2x2 on a register slot (+controls)
  B200Q_OP_X = 1,     // amplitude swap on a register slot (+controls): X, CNOT, Toffoli, ...
  B200Q_OP_DIAG = 2,  // diagonal over <= 2 selector bits anywhere in the index (+controls)
  B200Q_OP_MATK = 3,  // dense 2^k x 2^k, k = 2..4, on arbitrary tile bits, applied from shared memory
  B200Q_OP_LSWAP = 4  // complex64: exchange the lane bit (index bit 0) with register slot `slot` (planner-made)
};

/* Pre-decoded dispatch codes, filled by the planner.
 * LEAN codes (< B200Q_CODE_LEAN_END) are temp-free in-place register updates; a pass made only of them runs in
 * the lean instantiation of the tile kernel (no spills, small instruction footprint).  The other codes are the
 * general fallbacks (dense 2x2 with temporaries, register-slot-controlled ops, two-register-selector diagonals). */
#define B200Q_CODE_MAT1_HAD 0    /*  0..3 : un-controlled x*[[1,1],[1,-1]] on chunk slot 0..3 (scalar deferred)      */
#define B200Q_CODE_MAT1_ROTX 4   /*  4..7 : [[c, i q], [i q, c]], c^2 + q^2 = 1 (Rx) as three in-place shears         */
#define B200Q_CODE_MAT1_ROTY 8   /*  8..11: [[c, -y], [y, c]],   c^2 + y^2 = 1 (Ry) as three in-place shears         */
#define B200Q_CODE_DIAG_R 12     /* 12..15: diagonal with ONE selector on chunk slot 0..3 (other selector, if any,
                                            thread-level): unit-modulus phases as three in-place shears            */
#define B200Q_CODE_LSWAP 16      /* 16..19: exchange the complex64 lane bit with chunk slot 0..3 (in registers)    */
#define B200Q_CODE_X_RELABEL 20  /* chunk-level slot, no register-slot control: toggles the relabelling mask       */
#define B200Q_CODE_X_C1 21       /* chunk-level target, exactly one chunk-level control (arg = 4*target + control) */
#define B200Q_CODE_DIAG_T 22     /* diagonal with thread-level / global selectors only: folds into the phase rho   */
#define B200Q_CODE_X_LANE 23     /* complex64: X on the lane bit (arg = mask of chunk-level control slots)         */
#define B200Q_CODE_NONE 24
#define B200Q_CODE_LEAN_END 25
#define B200Q_CODE_MAT1_FAST 25  /* 25..36: variant (0 real, 1 rx-like, 2 general) * 4 + chunk slot, temporaries    */
#define B200Q_CODE_MAT1_SLOW 37  /* register-slot controls or the lane slot */
#define B200Q_CODE_X_SLOW 38
#define B200Q_CODE_DIAG 39       /* any diagonal (generic path) */

#define B200Q_FLAG_ADJOINT 1u  /* use the conjugate transpose of the stored matrix */
#define B200Q_FLAG_REAL 2u     /* MAT1: every entry is real (H, Ry, ...)                      */
#define B200Q_FLAG_RXLIKE 4u   /* MAT1: diagonal real, off-diagonal purely imaginary (Rx)    */
#define B200Q_FLAG_HAD 8u      /* MAT1: x * [[1, 1], [1, -1]] with x real (Hadamard)          */
#define B200Q_FLAG_ROT 16u     /* MAT1: unit-determinant rotation (with RXLIKE: Rx, with REAL: Ry) */
#define B200Q_FLAG_PHASE_SHIFT 5  /* DIAG, k = 1: exactly diag(1, i^q), q in bits 5-6 (1 S, 2 Z, 3 S^dagger)     */
#define B200Q_FLAG_PHASE_MASK (3u << B200Q_FLAG_PHASE_SHIFT)

#define B200Q_LAYOUT_SRC_SOA 1u /* pass reads complex64 chunks as (re0,re1,im0,im1) */
#define B200Q_LAYOUT_DST_SOA 2u /* pass writes them so */

typedef struct {
  uint8_t kind;
  uint8_t slot;      // MAT1 / X: register slot of the target
  uint8_t k;         // DIAG: number of selector bits (1..2); MATK: number of targets (2..4)
  uint8_t flags;
  uint16_t pool_off; // first element of this op's matrix in the shared-memory pool (multiple of 4)
  uint16_t pool_n;   // elements in the pool (multiple of 4)
  uint32_t mat_src;  // element offset of the dense 2^k x 2^k row-major matrix in the device buffer
  uint32_t ctrl_reg; // controls that are register slots: mask over the register amplitude index
  uint32_t ctrl_loc; // controls that are tile bits but not register slots: mask over tile-local bits
  uint32_t dsel_loc[2]; // DIAG selector j as a single tile-local bit mask (0 if not thread-level)
  uint64_t ctrl_glob;   // controls outside the tile: mask over physical index bits
  uint64_t dsel_glob[2];// DIAG selector j as a single physical bit mask outside the tile (or 0)
  uint8_t tk[4];     // MATK: tile-local bit of matrix-index bit j (j = 0 is the LSB)
  uint8_t dsel_slot[2]; // DIAG selector j: register slot index, or 0xff if not a register slot
  uint8_t code;      // pre-decoded dispatch code (B200Q_CODE_*), filled by the planner
  uint8_t tctrl;     // 1 if the op has thread-level or global controls (ctrl_loc / ctrl_glob non-zero)
  uint8_t arg;       // code-specific: X_C1 4*target + control chunk slots; DIAG_R index j of the register selector
  uint8_t pad[3];
  uint32_t gate_id;  // index of the source gate (diagnostics / adjoint gradient slot)
} b200q_op_t;

/* Fused pass + exchange (sharded path): the round that writes back to global memory stores every 16-byte chunk
 * straight into the receive buffer of the rank that owns it after the exchange -- over NVLink peer mappings for the
 * other ranks -- instead of the local shard.  The exchange is an arbitrary PERMUTATION OF INDEX BITS of the
 * distributed state: chunk-index bit j of this rank's shard goes to position perm[j]; positions >= n_chunk_bits
 * are rank bits (perm[j] = n_chunk_bits + k: the bit selects bit k of the destination rank).  The rank bits of
 * the source are constants of the launch: their images are pre-folded into `base` (same encoding).  Encoding of a
 * destination: chunk index in bits 0..39, rank in bits 40.. .  The block transpose of the reference scheme
 * (top rank_bits local bits <-> rank bits) is the special case the simple constructor builds. */
#define B200Q_MAX_RANKS 8
#define B200Q_DEST_RANK_SHIFT 40
typedef struct {
  void* peer[B200Q_MAX_RANKS];  // base of every rank's receive buffer (peer-mapped device pointers)
  uint64_t base;                // images of the source rank bits (destination encoding)
  uint8_t perm[40];             // destination position of chunk-index bit j
  int32_t n_chunk_bits;         // chunk-index bits of a shard = n_local - VS
  int32_t enabled;              // 0: ordinary in-place scatter
} b200q_remote_t;

typedef struct {
  uint8_t src_global;  // 1: gather from global memory, 0: from the shared-memory tile
  uint8_t dst_global;  // 1: scatter to global memory, 0: to the shared-memory tile
  uint8_t direct;      // 1: MATK ops applied in place in shared memory (no register gather)
  uint8_t pad;
  uint16_t op_begin, op_end;
  uint8_t slot_bit[5];     // tile-local amplitude bit of each register slot
  uint8_t nonreg_bit[11];  // tile-local amplitude bits enumerated by the thread's item index
} b200q_round_t;

/* A dense gate on 5 or 6 targets does not fit a tile round: it is a pass of its own ("dense pass"), marked by
 * n_rounds == 0 and n_ops == 1.  ops[0]: kind MATK, k = number of targets, mat_src / flags as usual, ctrl_glob = mask of
 * ALL controls over physical bits, dsel_glob[0] = the physical bit of matrix-index bit j in byte j.  Run by
 * b200q_dense_kernel (one CTA group per 2^k amplitudes, matrix staged in shared memory). */
#define B200Q_DENSE_MAX 6
typedef struct {
  uint8_t n_bits;     // physical index bits of the (padded) local state
  uint8_t n_qubits;   // index bits of the state itself (< n_bits only for states smaller than the register bits)
  uint8_t tile_bits;  // amplitude bits per tile
  uint8_t n_rounds;
  uint8_t n_ops;
  uint16_t pool_elems;
  uint8_t n_nontile;
  uint8_t layout;     // B200Q_LAYOUT_* (complex64 only)
  uint8_t lean;       // every op has a LEAN code: run the lean kernel instantiation
  uint8_t needs_pool; // some op reads the dense matrix pool in shared memory (MATK, general MAT1 / X / DIAG paths)
  uint8_t has_scale;  // the pass has ops with a deferred common scalar (pass_scale), applied by the last round
  uint8_t n_gctrl;    // ops with controls outside the tile (evaluated once per tile, see tile_enabled)
  uint8_t gctrl_ops[B200Q_MAX_OPS];
  uint8_t tile_phys[B200Q_MAX_TILE_BITS];   // physical bit of tile-local bit j (ascending)
  uint8_t nontile_phys[B200Q_MAX_QUBITS];   // physical bits enumerated by the tile (CTA) index
  b200q_round_t rounds[B200Q_MAX_ROUNDS];
  b200q_op_t ops[B200Q_MAX_OPS];
} b200q_pass_t;
