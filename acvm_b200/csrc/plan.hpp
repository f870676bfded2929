// Plan compiler: ACIR opcodes -> a statically scheduled stream of fixed-size device records.
//
// The reference solves opcodes strictly in order, one instance at a time, discovering which
// witnesses are known as it goes (acvm/src/pwg/mod.rs:236-303, arithmetic.rs:27-239).  For a batch
// of instances of the SAME circuit the known-set after opcode i is a static property of the
// circuit, so it is resolved ONCE here on the host:
//   * every Arithmetic opcode becomes ASSIGN / CHECK micro-gates with the unknown's coefficient
//     inverted at plan time (the reference inverts per gate per instance),
//   * micro-ops are list-scheduled into "steps" of S independent slots (dependencies from the
//     witness dataflow), so one CTA = one tile of T instances executes S*T lanes per step,
//   * the record stream is what the kernel stages through shared memory with TMA bulk copies.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "acir.hpp"

namespace acvmb {

// ---- device record (192 B, 16 B aligned) ------------------------------------------------------
struct OpRec {
    uint32_t w[8];      // [0]=kind|flags<<8  [1]=acir opcode index  [2]=out slot  [3]=x  [4]=y  [5]=w1  [6]=w2  [7]=aux
                        // gate / logic / range micro-ops: an operand field with bit 31 set is an index into the shared-memory ring of
                        // recent values instead of a column; the output of a gate also goes to ring entry w[7], that of AND / XOR to
                        // ring entry w[5] (RING_NONE = not kept)
    // gate constants (8 x u32 little-endian limbs each), layout by flag:
    //   GF_MUL   : c0 = cM*R^2, c1 = alpha, c2 = beta, c3 = c1*R (w1), c4 = gamma     out = cM*(x+alpha)*(y+beta) + c1*w1 + gamma
    //   GF_MUL|GF_ONE_RED : no c0                                                     out = (x+alpha)*(y+beta)/R + c1*w1 + gamma
    //   otherwise: c1 = cY*R (y), c2 = c1*R (w1), c3 = c2*R (w2), c4 = cC            out = cY*y + c1*w1 + c2*w2 + cC
    // all in terms of STORED column values (plan.cpp "scaled columns"); linear operands beyond GF_NPROD are +- additions;
    // c4 is stored as const*R when the gate has a reduction (it is its initial accumulator), as is otherwise
    uint32_t c[5][8];
};
static_assert(sizeof(OpRec) == 192, "OpRec layout");

enum MicroKind : uint32_t {
    MK_NOP = 0,
    MK_GATE_ASSIGN = 1,   // out := gate expression (see OpRec::c)
    MK_GATE_CHECK = 2,    // same expression must be 0 else UnsatisfiedConstrain@opcode
    MK_AND = 3,           // out := (x & y) masked to aux bits, mod p       (logic.rs:11-56)
    MK_XOR = 4,
    MK_RANGE = 5,         // num_bits(x) > aux  => UnsatisfiedConstrain     (range.rs:7-18)
    MK_SHA256 = 6,        // payload[aux..]: n_in, out_check_mask, var_size_witness|NONE, 0, (witness,num_bits)*n_in, 32 outputs
    MK_KECCAK256 = 7,
    MK_FIXED_BASE = 8,    // x=low y=high out=x-coord w1=y-coord slot
    MK_PEDERSEN = 9,      // payload[aux..]: n_in, domain_separator, witness*, out_x, out_y
    MK_GATE_GENERAL = 10, // value-dependent arithmetic gate (unknown is a mul operand)
    MK_COPY_CHECK = 11,   // (reserved)
    MK_COPY = 13,         // out := x                                   (MemoryInit, memory_op.rs:47-60)
    MK_TO_LE_RADIX = 14,  // x = value; payload[aux..]: n_b, radix, check-mask words, b witnesses   (directives/mod.rs:60-87)
    MK_QUOTIENT = 15,     // x = a, y = b, w1 = predicate|NONE, out = q, w2 = r                     (directives/mod.rs:28-59)
    MK_MEM_READ = 16,     // x = index, w1 = predicate|NONE, out = witness; payload[aux..]: base, len  (memory_op.rs:62-110)
    MK_MEM_WRITE = 17,    // x = index, y = value, w1 = predicate|NONE;     payload[aux..]: base, len  (memory_op.rs:111-123)
    MK_BLAKE2S = 18,      // same payload as MK_SHA256                                  (hash.rs:28-48 -> blake2 0.10.6)
    MK_HASH_TO_FIELD = 19,// payload: n_in, check, NONE, 0, (witness,num_bits)*, 1 output: blake2s digest reduced mod p (hash.rs:13-24)
    MK_ECDSA = 20,        // out := verify; payload[aux..]: curve (0 k1, 1 r1), 32 pkx, 32 pky, 64 sig, 32 hashed-message witnesses (signature/ecdsa.rs)
    MK_CURVE_PART = 21,   // out (3 slots from w[2]) := sum of table points picked by windows of one scalar; parameters in c[0] (heavy_ops.cuh)
    MK_JAC_ADD = 22,      // out (3 slots from w[2]) := point at w[3] + point at w[4]   (Jacobian, Montgomery form, Z = 0 is infinity)
    MK_JAC_FINAL = 23,    // (x, y) at w[2], w[5] := affine(point at w[3] [+ point at w[4]]); c[0][0]: validate fixed-base scalar w[6], w[7]
    MK_REQUIRE = 12,      // payload[aux..]: n, (witness, mu_index)*n : first one not assigned in this lane => MissingAssignment
    // Hash calls over byte-valued inputs, split so that only the compression is on the dependency chain (plan.cpp hash_packed):
    MK_HASH_PACK = 25,    // out column := up to 32 message bytes packed (byte i at bit 8i); payload[aux..]: n, witness * n
    MK_HASH_CORE = 26,    // out column := digest (32 bytes packed); payload[aux..]: func (0 SHA256, 1 Keccak256, 2 Blake2s), n_bytes,
                          // n_chunks, chunk column * n_chunks (32 message bytes each: a MK_HASH_PACK column or an earlier digest column);
                          // w[6] == 1: the same descriptor sits in the record's c[][] words instead (up to 37 chunks);
                          // w[5] != NONE (SHA256 over k * 64 bytes): payload offset of the padding block's 64 K[i] + W[i] words
    MK_HASH_UNPACK = 27,  // x = digest column; payload[aux..]: check_mask, 32 output witnesses (insert_value each byte)
    MK_INT_OP = 24,       // out := BinaryIntOp(x, y) of a lowered Brillig opcode; w[7] = op | bit_size << 8, 1 <= bit_size <= 128
                          // (brillig_vm/src/arithmetic.rs:23-81; a condition on which the reference panics => EK_REFERENCE_PANIC)
};

// flags (w[0] >> 8)
enum : uint32_t {
    GF_MUL = 1u << 0,        // cM term present (x and y loaded)
    GF_Y = 1u << 1,          // y term present (always set when GF_MUL)
    GF_NLIN_SHIFT = 2,       // bits 2..3: number of extra linear terms (0..2) -> w1, w2
    GF_W1_IS_X = 1u << 4,    // (unused since the x/y linear terms are folded into alpha/beta)
    GF_OUT_CHECK = 1u << 5,  // output slot already holds a value: compare instead of store (insert_value, mod.rs:338-357)
    GF_HEAVY = 1u << 6,      // needs the FULL kernel variant
    GF_OUT2_CHECK = 1u << 7, // second output (point y coordinate) already holds a value
    GF_ADDSUB = 1u << 8,     // linear gate whose coefficients are all +-1: out = +-y +-w1 +-w2 + cC, no multiplication
    GF_NEG_Y = 1u << 9, GF_NEG_W1 = 1u << 10, GF_NEG_W2 = 1u << 11,
    GF_ONE_RED = 1u << 12,   // GF_MUL gate whose product (x+c1)*(y+c2) enters the SAME reduction as the linear products (scaled columns)
    GF_NPROD_SHIFT = 13,     // bits 13..14: how many of the linear operands (y, w1, w2 in this order; w1 for GF_MUL) are products
                             // with a plan constant; the others are plain +- additions (signs in GF_NEG_*)
};

// error kinds mirrored from OpcodeResolutionError (acvm/src/pwg/mod.rs:100-114) + reference panics
enum ErrKind : uint32_t {
    EK_NONE = 0,
    EK_MISSING_ASSIGNMENT = 1,        // OpcodeNotSolvable(MissingAssignment(w)); aux = witness
    EK_TOO_MANY_UNKNOWNS = 2,         // OpcodeNotSolvable(ExpressionHasTooManyUnknowns)
    EK_UNSUPPORTED_BLACKBOX = 3,      // UnsupportedBlackBoxFunc; aux = func
    EK_UNSATISFIED_CONSTRAIN = 4,     // opcode_location = Resolved(Acir(opcode_index))
    EK_INDEX_OUT_OF_BOUNDS = 5,       // aux = index (array_size via acvmb_status.aux2)
    EK_BLACKBOX_FAILED = 6,           // aux = func
    EK_BRILLIG_FAILED = 7,
    EK_REFERENCE_PANIC = 8,           // the reference would panic!() here (malformed circuit / API misuse)
};

constexpr uint32_t RING_FLAG = 0x80000000u, RING_NONE = 0xFFFFFFFFu;

enum StatusCode : uint32_t { ST_SOLVED = 0, ST_IN_PROGRESS = 1, ST_FAILURE = 2, ST_REQUIRES_FOREIGN_CALL = 3 };

struct StaticFail {
    bool present = false;
    uint32_t opcode = 0, kind = 0, aux = 0;
    std::string detail;
};

struct PlanStats {
    uint64_t n_opcodes = 0, n_micro = 0, n_steps = 0, n_slots_filled = 0;
    uint64_t n_gate_assign = 0, n_gate_check = 0, n_logic = 0, n_range = 0, n_hash = 0, n_curve = 0;
    uint64_t ref_fr_mul = 0;       // Fr multiplications the reference performs per instance (SURVEY 8d accounting)
    uint64_t ref_fr_inv = 0;       // field inversions the reference performs per instance
    uint64_t dev_imad = 0;         // 32x32 multiply-accumulates the device executes per instance (gate ops)
    uint64_t alg_bytes = 0;        // algorithmic HBM bytes per instance: 32 B per operand read + 32 B per witness written
    uint64_t n_temps = 0;
    uint64_t n_gate_general = 0;   // value-dependent gates resolved per lane
    uint64_t n_directive = 0, n_memory = 0, n_brillig = 0;
    uint64_t n_brillig_device = 0;   // Brillig opcodes lowered to device gates (no host segment)
    uint64_t n_gate_one_reduction = 0;   // multiplicative gates that need ONE Montgomery reduction (scaled columns)
    uint64_t n_operand_reads = 0, n_ring_reads = 0;   // operand loads of gate / logic / range micro-ops, and how many come from the ring
};

// The opcode list is cut into segments: device segments are step ranges of the record stream; a host segment is one
// Brillig opcode executed by the host VM on columns copied out of / back into HBM (north star: "Brillig opcodes
// execute on the host brillig_vm with results DMA'd back into the device WitnessMap").
struct Segment {
    uint32_t kind;   // 0 = device, 1 = host Brillig, 2 = host PermutationSort (descriptor: n, tuple, n_sort_by, sort_by*, slots, n_bits, {witness, known}*)
    uint32_t a;      // device: first step        host: ACIR opcode index
    uint32_t b;      // device: number of steps   host: offset into Plan::host_desc
    uint32_t c;
};

struct Plan {
    uint32_t S = 16;                   // slots per step
    uint32_t num_witnesses = 0;        // current_witness_index + 1 (dense output width)
    uint32_t n_slots = 0;              // witnesses + temporaries
    uint32_t n_opcodes = 0;
    uint32_t chunk_steps = 2;          // steps per TMA stage
    uint32_t ring_slots = 0;           // entries of the shared-memory ring of recent values the stream was compiled for (0 = none)
    bool needs_full_kernel = false;
    std::vector<uint32_t> input_witnesses;  // order of the per-instance input columns
    std::vector<uint32_t> input_scaled;     // per input: 1 = the column holds value*R (Montgomery), the scatter converts
    // scaled columns (plan.cpp): per witness, 8 limbs of mu_w * R -- the output gather multiplies the stored value by it
    // (Montgomery product) to get the canonical value.  Empty: every column is canonical.
    std::vector<uint32_t> unscale;
    std::vector<OpRec> stream;         // n_steps_padded * S records
    uint32_t n_steps = 0;              // padded to a multiple of chunk_steps
    std::vector<uint32_t> payload;     // variable-length operand lists (hash inputs ...)
    std::vector<uint32_t> assign_opcode;  // per witness: opcode index that assigns it, 0xFFFFFFFF = never, 0xFFFFFFFE = input,
                                          // 0xFFFFFFFD = value-dependent: the per-lane table mu_assign[mu_index_of[w]] decides
    std::vector<uint32_t> mu_index_of;    // per witness: index into the per-lane "maybe assigned" table, or 0xFFFFFFFF
    uint32_t n_mu = 0;
    std::vector<Segment> segments;
    // host segment descriptors: pred_slot, n_inputs, {is_array, count, slots...}*, n_outputs, {is_array, count, {witness, known}...}*
    std::vector<uint32_t> host_desc;
    std::vector<uint8_t> acir_gz;      // original circuit bytes, kept only when host segments need the Brillig bytecode
    StaticFail static_fail;            // the whole batch fails here (unless an instance failed earlier)
    PlanStats stats;
};

struct PlanOptions {
    uint32_t S = 16;
    uint32_t chunk_steps = 2;
    uint32_t temp_pool = 2048;
    bool split_curve = true;   // lower FixedBaseScalarMul / Pedersen into parallel partial-sum micro-ops when S >= 8
    // BlackBoxFuncCall::Pedersen is REFUSED unless this is set: barretenberg's generator tables cannot be reproduced from
    // the reference tree (both reference KATs fail), so the kernel computes the plookup-structured
    // commitment over this project's own generators -- benchmarks and structure tests opt in, a drop-in caller must not
    // receive a different hash behind the reference's opcode tag.
    bool allow_unpinned_pedersen = false;
    // Lower Brillig opcodes whose bytecode is straight-line field arithmetic (Const / Mov / BinaryFieldOp Add, Sub, Mul /
    // Stop -- the stdlib's `bytecode: vec![Stop]` constant loads, stdlib/src/blackbox_fallbacks/uint.rs:51-63,85) to device
    // gates at plan time instead of a host segment each.
    bool device_brillig = true;
    // Columns written and read only by arithmetic gates hold lambda_w * value for a per-column plan constant lambda_w
    // (plan.cpp "scaled columns"): one Montgomery reduction per multiplicative gate instead of two.
    bool scaled_columns = true;
    // SHA256 / Keccak256 / Blake2s calls over byte inputs become pack / core / unpack micro-ops (MK_HASH_*): the gather of the
    // message and the scatter of the digest run on other slot threads, a digest that is the message of a later call is
    // handed over as one packed column.
    bool packed_hashes = true;
    bool sha_pad_table = true;   // SHA256 over k * 64 bytes: the padding block's K + W table is a plan constant (MK_HASH_CORE w[5])
    // Tiles narrower than a warp: the heavy micro-ops of one step are spread over different warps (Scheduler::emit).
    // tile_lanes = the T the runtime will use (0: the runtime's own rule, 32 with curve calls, else 128 / S).
    bool spread_heavy = true;
    uint32_t tile_lanes = 0;
    // Curve micro-ops with two or more levels of slack (the H1 sums of a Pedersen chain) run in the idle slots of the level just
    // before their first successor instead of in steps of their own (Scheduler::flush).
    bool slack_scheduling = true;
    // Recent-value ring in shared memory (vm_kernel_impl.cuh): the last `ring_slots` values written by gate / logic micro-ops
    // of a tile are kept on chip, and an operand whose producer is that recent is read from there instead of from L2
    // (operand fields get bit 31 set and carry the ring index).  0 disables.
    uint32_t ring_slots = 0;
};

// Throws std::runtime_error for opcodes outside the device scope.
Plan compile_plan(const Circuit& c, const std::vector<uint32_t>& input_witnesses, const PlanOptions& opt);

// flat byte blob (host metadata + stream) used for the one-time multi-GPU broadcast
std::vector<uint8_t> serialize_plan(const Plan& p);
Plan deserialize_plan(const uint8_t* data, size_t len);

}  // namespace acvmb
