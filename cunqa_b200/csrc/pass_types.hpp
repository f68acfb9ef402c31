// pass_types.hpp -- POD descriptors shared by the host planner and the sm_100a tile-pass kernel.
//
// A *tile pass* is one in-place sweep over the whole state (2*S algorithmic bytes): the state is
// cut into tiles of 2^t amplitudes spanned by a chosen set of t "tile qubits" (the low L of them
// are always physical qubits 0..L-1, so that every tile is made of contiguous runs of 2^L
// amplitudes = 128-bit coalesced segments); each CTA stages a tile in shared memory, applies the
// pass's whole op list to it, and writes it back.
#pragma once
#include <cstdint>

namespace cb200 {

constexpr int kMaxTileBits = 13;  // 2^13 complex128 = 128 KiB of shared memory
constexpr int kMaxOps = 48;       // ops per pass
constexpr int kMaxDenseK = 5;     // dense block width (qubits)
constexpr int kMaxDiagK = 12;     // diagonal table width (qubits)
constexpr int kMaxFix = 12;       // targets + in-tile controls of one op
constexpr int kMatScalars = 2560; // scalars (re/im) of dense-matrix storage per pass (kernel params)
constexpr int kDiagLut = 192;
constexpr int kFanMaxCenters = 8;       // centers of a phase fan
constexpr int kFanMaxLeafBits = 6;      // in-tile leaves of a phase fan (tables of 2^6 entries per center)
constexpr int kFanMaxChunks = 4;        // out-of-tile leaves, 8 per chunk (256-entry table per chunk and center)
constexpr int kMaxSweeps = 24;          // register sweeps per pass
constexpr int kStabEntries = 1024;      // per-tile store of small diagonal sub-tables (entries; 16 KB of complex128)
constexpr int kStabMaxBits = 6;        // a diagonal op is staged when at most this many of its qubits are tile qubits
constexpr uint16_t kNotStaged = 0xFFFF;     // per diagonal op: 64 entries for tile bits 0..5 + 128 for tile bits 6..12

enum OpKind : uint8_t {
    OP_DENSE = 0,  // dense 2^k x 2^k on in-tile targets, optional controls
    OP_DIAG = 1,   // diagonal table over k qubits anywhere (in-tile or not)
    OP_PERMX = 2,  // X on tbit[0] (controls optional): pure pair swap, no flops
    OP_SWAP = 3,   // swap tbit[0] <-> tbit[1] (controls optional)
    OP_PAIR = 4,   // register block: two dense gates A (k qubits, local bits [0,k)) and B (k2 qubits, local
                   // bits [s,s+k2); k2 = 0: A alone) applied back to back to the 2^nfix amplitudes a thread holds in
                   // registers; nfix = 4 local bits (tbit[0..3], the gates' bits first, then spectator bits no gate
                   // touches) whenever the tile has four bits
    OP_FAN = 5,    // phase fan at the end of a pass: every "center" (a tile bit, k <= 8 of them) carries a phase that is a
                   // product of one factor per other qubit ("leaf": in-tile non-center bits and out-of-tile qubits); no term
                   // joins two centers or two leaves.  amplitude *= g * prod_{centers b set} [c_b(tile) * T_b[leaf bits]].
                   // This is what the controlled-phase ladders of a QFT leave behind once the terms that commute with the
                   // rest of the pass are pulled out of the per-amplitude tables (planner.cpp: encode_pass).
    OP_LINPERM = 6, // a run of unconditioned X / CX / SWAP gates as ONE permutation of the tile: X, CX and SWAP map index
                    // bits linearly over GF(2), so a run of them moves the amplitude at tile index s(d) = c ^ M d to d.  The
                    // shared-memory swizzle is linear too and is folded in: lin[i] (uint16 view of bytes 4..35 of the
                    // descriptor, DevOp::lin) = swz(column i of M) for tile bit i, lin[15] = swz(c): the slot of the source
                    // of d is the XOR of lin[15] and the lin[i] of the bits set in d.  One load + one store per amplitude
                    // and two barriers whatever the length of the run (a CX ladder of an ansatz layer: 6-9 gates).
};

struct DevOp {
    uint8_t kind;
    uint8_t k;       // DENSE: #targets; DIAG: #qubits of the table; PAIR: #targets of gate A; LINPERM: #gates of the run
    uint8_t nfix;    // DENSE/PERM/PAIR: number of entries in fixpos (targets + in-tile controls)
    uint8_t has_mask;
    union {
        struct {
            uint8_t tbit[8];        // DENSE/PERM: tile-bit position of target m
            union {
                uint8_t fixpos[kMaxFix]; // ascending tile-bit positions where a bit is inserted into the group index
                uint16_t sboff[4];       // PAIR: swizzled tile index of local bit m, swz(1 << tbit[m])
            };
            union {
                uint8_t dsrc[kMaxDiagK]; // DIAG: source of table-index bit m: <64 tile bit, >=64 physical qubit (v-64) outside the tile
                uint16_t seg[6];         // PAIR: group index -> tile index deposit, idx = sum_k (g & seg[k]) << k  (k <= nfix)
            };
        };
        uint16_t lin[16];           // LINPERM: swizzled columns of the index map, [15] = swizzled constant
    };
    uint32_t ctrl_in;   // in-tile control bits (OR-ed into the tile index; those bits are in fixpos)
    uint32_t mat_off;   // DENSE: scalar offset into PassParams::mats; DIAG: element offset into the diag-table buffer
    uint64_t ctrl_out;  // physical-qubit mask of controls outside the tile (tested on the tile base)
    int32_t mask_row;   // row of the per-state flag matrix (batched feed-forward) when has_mask
    uint8_t k2, s;      // DENSE: k2 = 1: plain (re, im) elements even in a pass with mat_form = 1 (unused by encode_pass today: such
                        // passes have no sweeps); s = the gate's first register bit in its sweep
                        // PAIR: #targets of gate B and its first local bit; B's matrix follows A's in mats
                        // DIAG: k2 = slot of this op among the pass's diagonal ops (index of its pext tables); s = 1 when the
                        // table depends on a register bit of its run (looked up per amplitude)
    uint16_t stage_off;  // DIAG: entry offset of this op's per-tile sub-table in the shared-memory table store; kNotStaged = read from global memory
};
static_assert(sizeof(DevOp) == 64, "DevOp layout");

template <typename T>
struct PassParams {
    int32_t n;        // qubits per state
    int32_t t;        // tile bits
    int32_t L;        // low tile bits that are physical qubits 0..L-1 (contiguous run = 2^L amps)
    int32_t n_ops;
    int32_t batch;    // independent states
    int32_t mat_form; // 0: <= 2-qubit matrices as plain (re, im) pairs; 1: three-multiplication form (tile_kernel.cuh), read from
                      // the kernel parameters by the CM kernel variant
    uint64_t tiles_per_state;
    uint64_t cond_mask; // ops whose activity depends on the tile (controls outside the tile / per-state mask)
    uint64_t chunk_mask, chunk_val;  // only tiles whose base satisfies (base & chunk_mask) == chunk_val are processed
                                     // (a sharded pass pipelined behind a chunked qubit exchange); 0 / 0 = every tile
    int32_t n_mat;      // scalars used in mats (staged into shared memory once per CTA)
    int32_t reg_stride; // mat_form 1 with several registers in one launch: state b of the launch reads its matrices at
                        // mats + b * reg_stride (0 = one set for all states)
    int32_t needs_full; // kernel variant: 0 = lean (1-2 qubit dense gates, common register blocks, permutations), 1 = lean +
                        // diagonal runs, 2 = full (dense k>=3, unusual register blocks, generic diagonal path)
    int32_t n_diag;     // number of OP_DIAG ops; diag_op[slot] = op index of the slot-th one
    int32_t n_stab;     // entries of the shared-memory sub-table store used by this pass (multiple of 8)
    uint8_t diag_op[kMaxOps];
    // phase fan (at most one per pass, always the last op): DevOp kind OP_FAN with k = #centers, tbit[c] = tile bit of
    // center c, nfix = #in-tile leaf bits, fixpos[] = their tile bits (ascending), ctrl_in = their tile-bit mask, dsrc[0..3]
    // = the four centers that are register bits of the sweep, mat_off = first entry of its block in the diagonal-table
    // buffer: [0] scalar g, then T[c][2^nfix], then X[chunk][c][256] (the out-of-tile leaves, 8 physical qubits per chunk)
    // register sweeps: consecutive device ops (dense gates of <= 2 qubits, diagonal runs, the fan) that are applied to
    // the 16 amplitudes a thread holds in registers between ONE load and ONE store of the tile: sweep_at[o] = sweep that
    // starts at op o (0xFF none); a sweep covers ops [first, first + len) with register bits rb[0..3] (asc: ascending)
    struct Sweep { uint8_t first, len, rb[4], asc[4]; };
    int32_t n_sweeps;
    uint8_t sweep_at[kMaxOps];
    Sweep sweeps[kMaxSweeps];
    int32_t fan_op;      // op index, -1 = none
    int32_t fan_chunks;  // number of out-of-tile leaf chunks
    uint8_t fan_q[kFanMaxChunks][8];  // physical qubit of every chunk bit (0xFF = unused)
    uint8_t stab_slot[kStabEntries / 8];  // diagonal slot that owns each 8-entry chunk of the store
    uint8_t tile_q[16]; // physical qubit of tile bit j (ascending)
    uint32_t tile_seg[kMaxTileBits + 1];  // tile number -> (tile base >> L): sum_k (x & tile_seg[k]) << k, k <= t - L (the bits of the
                                          // tile number between two high tile qubits move up together)
    DevOp ops[kMaxOps];
    alignas(16) T mats[kMatScalars];  // dense matrices: <= 2 qubits in the three-multiplication form of tile_kernel.cuh (3 scalars per
                                      // element), wider blocks and the gates of register sweeps as (re, im) pairs
};

// dim3 / smem sizing helpers shared with the host
constexpr int kTileThreads = 256;

}  // namespace cb200
