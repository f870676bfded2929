/* acvm_b200 -- C ABI of the B200-native batched ACIR witness solver.
 *
 * This is the drop-in boundary for ONE hot path of noir-lang/acvm 0.27.0: solving the same
 * compiled Circuit for many independent initial witnesses.  Every entry point cites the reference
 * interface it replaces (paths relative to the acvm repository root).  Plain pointers and sizes
 * only; the caller owns every buffer; nothing unwinds across the boundary (all functions return
 * 0 on success or a negative acvmb_rc, with a message available from acvmb_last_error()).
 *
 * A context is bound to one CUDA device (or to several, acvmb_ctx_create_multi) and is single-threaded, like the reference's
 * `Barretenberg` (RefCell<Store>, barretenberg_blackbox_solver/src/wasm/mod.rs:59-63).
 * There is NO CPU fallback: creating a context without a usable sm_100 device fails.
 *
 * Byte conventions: a field element is 32 bytes big-endian, canonical value of BN254 Fr
 * (FieldElement::to_be_bytes, acir_field/src/generic_ark.rs:269-277); inputs are reduced mod p
 * like FieldElement::from_be_bytes_reduce (:281-283).
 */
#ifndef ACVM_B200_H
#define ACVM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct acvmb_ctx acvmb_ctx;
typedef struct acvmb_circuit acvmb_circuit;
typedef struct acvmb_batch acvmb_batch;
typedef struct acvmb_vm acvmb_vm;

typedef enum {
    ACVMB_OK = 0,
    ACVMB_ERR_INVALID_ARG = -1,
    ACVMB_ERR_NO_DEVICE = -2,   /* no CUDA device / not sm_100: the product never falls back to the CPU */
    ACVMB_ERR_CUDA = -3,
    ACVMB_ERR_DECODE = -4,      /* bytes are not gzip(bincode(Circuit)) */
    ACVMB_ERR_UNSUPPORTED = -5, /* opcode outside the device scope */
    ACVMB_ERR_OOM = -6,
    ACVMB_ERR_STATE = -7        /* API misuse where the reference would panic (e.g. finalize before Solved) */
} acvmb_rc;

/* ACVMStatus (acvm/src/pwg/mod.rs:33-51) */
typedef enum { ACVMB_SOLVED = 0, ACVMB_IN_PROGRESS = 1, ACVMB_FAILURE = 2, ACVMB_REQUIRES_FOREIGN_CALL = 3 } acvmb_status_code;

/* OpcodeResolutionError variants (acvm/src/pwg/mod.rs:100-114) + "the reference panics here" */
typedef enum {
    ACVMB_E_NONE = 0,
    ACVMB_E_MISSING_ASSIGNMENT = 1,   /* OpcodeNotSolvable(MissingAssignment(aux)) */
    ACVMB_E_TOO_MANY_UNKNOWNS = 2,    /* OpcodeNotSolvable(ExpressionHasTooManyUnknowns) */
    ACVMB_E_UNSUPPORTED_BLACKBOX = 3, /* UnsupportedBlackBoxFunc(aux) */
    ACVMB_E_UNSATISFIED_CONSTRAIN = 4,/* opcode_location = Resolved(Acir(opcode_index)) (mod.rs:286-296) */
    ACVMB_E_INDEX_OUT_OF_BOUNDS = 5,
    ACVMB_E_BLACKBOX_FAILED = 6,      /* BlackBoxFunctionFailed(aux = func, ...) */
    ACVMB_E_BRILLIG_FAILED = 7,
    ACVMB_E_REFERENCE_PANIC = 8
} acvmb_err_kind;

/* one per instance */
typedef struct {
    uint32_t code;          /* acvmb_status_code */
    uint32_t err_kind;      /* acvmb_err_kind */
    uint32_t opcode_index;  /* instruction pointer at failure (ACVM::instruction_pointer, mod.rs:171-173) */
    uint32_t aux;           /* witness index / blackbox func, per err_kind */
} acvmb_status;

/* run record of the last device solve of a batch (for bench.py's roofline) */
typedef struct {
    double kernel_ms;            /* CUDA-event time of the step-VM kernel(s) on the library's stream */
    double scatter_ms, gather_ms;/* input scatter / output gather kernels */
    uint64_t kernel_launches;    /* number of OUR kernels launched */
    uint32_t T, S;               /* tile width (instances per CTA), slots per step */
    uint32_t n_tiles, threads_per_cta;
    uint32_t resident_instances; /* instances per sub-batch */
    uint32_t n_subbatches;
} acvmb_run_info;

typedef struct {
    uint64_t n_opcodes, n_micro_ops, n_steps, n_slots_filled;
    uint64_t n_gate_assign, n_gate_check, n_logic, n_range, n_hash, n_curve;
    uint64_t ref_fr_mul;     /* Fr multiplications the reference performs per instance */
    uint64_t ref_fr_inv;     /* field inversions the reference performs per instance */
    uint64_t dev_imad;       /* 32x32 multiply-accumulates executed per instance on the device */
    uint64_t alg_bytes;      /* algorithmic HBM bytes per instance */
    uint64_t n_temps;
    uint32_t num_witnesses, n_slots, S, needs_full_kernel;
    uint32_t static_fail_present, static_fail_opcode, static_fail_kind, static_fail_aux;
    uint32_t n_segments;        /* device + host segments of the plan */
    uint32_t n_host_segments;   /* host segments: Brillig opcodes that need the host VM, PermutationSort directives */
    uint32_t n_brillig;         /* Brillig opcodes in the circuit */
    uint32_t n_brillig_device;  /* ... of which lowered to device gates at plan time (straight-line field bytecode) */
    uint64_t n_gate_one_reduction; /* multiplicative gates that run ONE Montgomery reduction (scaled columns) */
    uint32_t scaled_columns;    /* 1: some columns hold lambda_w * value; canonical values are produced by the output gather */
    uint32_t ring_slots;        /* entries of the shared-memory ring of recent values the stream was compiled for */
    uint64_t n_operand_reads;   /* operand loads of gate / logic / range micro-ops per instance ... */
    uint64_t n_ring_reads;      /* ... of which served from the shared-memory ring */
} acvmb_plan_info;

/* ---- context ---------------------------------------------------------------------------- */
int acvmb_ctx_create(int device, acvmb_ctx** out);
/* Multi-device context (one node): circuits created on it are compiled once and replicated to every device with ONE broadcast
 * (ncclBroadcast when libnccl.so.2 can be loaded, peer copies otherwise); acvmb_solve_batch / _ex shard the batch contiguously
 * over the devices (one host thread each) and fill the caller's buffers; nothing moves between devices during a solve.  The
 * device-resident acvmb_batch_* calls and the single-instance acvmb_vm_* mirror use devices[0]. */
int acvmb_ctx_create_multi(const int* devices, int n, acvmb_ctx** out);
int acvmb_ctx_n_devices(const acvmb_ctx* ctx);
const char* acvmb_ctx_broadcast_backend(const acvmb_ctx* ctx);   /* "nccl" | "peer-copy" | "none" (nothing replicated yet) */
/* destroy every circuit / batch / vm created on a context BEFORE the context itself (they keep a pointer to it) */
void acvmb_ctx_destroy(acvmb_ctx* ctx);
const char* acvmb_last_error(void);           /* thread-local message of the last failing call */
int acvmb_device_name(acvmb_ctx* ctx, char* buf, size_t len);
/* pinned host memory for the I/O buffers of acvmb_solve_batch (pageable buffers work, slower) */
void* acvmb_host_alloc(size_t bytes);
void acvmb_host_free(void* p);

/* ---- circuit: replaces Circuit::read (acir/src/circuit/mod.rs:155-161) + ACVM::new's opcode
 * ownership (acvm/src/pwg/mod.rs:146-156).  `input_witnesses` are the keys of the initial
 * WitnessMap -- identical for every instance of a batch.  The plan is compiled here, once. ---- */
int acvmb_circuit_from_acir(acvmb_ctx* ctx, const uint8_t* gz_bincode, size_t len, const uint32_t* input_witnesses,
                            uint32_t n_inputs, acvmb_circuit** out);
void acvmb_circuit_destroy(acvmb_circuit* c);
int acvmb_circuit_info(const acvmb_circuit* c, acvmb_plan_info* out);
/* per witness: opcode index that assigns it, 0xFFFFFFFE = initial witness, 0xFFFFFFFF = never */
int acvmb_circuit_assign_opcodes(const acvmb_circuit* c, uint32_t* out, uint32_t n);
/* one-time multi-GPU distribution: rank 0 serialises the compiled plan, every other rank loads it */
int acvmb_circuit_serialize(const acvmb_circuit* c, uint8_t* buf, size_t cap, size_t* needed);
int acvmb_circuit_deserialize(acvmb_ctx* ctx, const uint8_t* blob, size_t len, acvmb_circuit** out);

/* ---- batched ACVM::new + solve + finalize (acvm/src/pwg/mod.rs:146-156,236-241,176-181) ----
 * inputs_be32  [batch][n_inputs][32]   initial witness values, order of `input_witnesses`
 * out_witness  [batch][n_out][32]      solved witnesses; n_out = num_witnesses when out_ids == NULL
 *                                      (dense WitnessMap, index 0..current_witness_index), else the
 *                                      listed witness indices only.  Unassigned entries are zero.
 *                                      May be NULL (statuses only).
 * out_status   [batch] */
int acvmb_solve_batch(acvmb_circuit* c, uint32_t batch, const uint8_t* inputs_be32, const uint32_t* out_ids,
                      uint32_t n_out_ids, uint8_t* out_witness_be32, acvmb_status* out_status);
/* same, plus out_present [batch][n_out]: 1 where the instance's WitnessMap holds the witness (needed when a circuit
 * has value-dependent gates, arithmetic.rs:217-221: which witnesses get assigned then differs per instance) */
int acvmb_solve_batch_ex(acvmb_circuit* c, uint32_t batch, const uint8_t* inputs_be32, const uint32_t* out_ids,
                         uint32_t n_out_ids, uint8_t* out_witness_be32, uint8_t* out_present, acvmb_status* out_status);
int acvmb_last_run_info(const acvmb_circuit* c, acvmb_run_info* out);

/* ---- device-resident batch: the same solve split into its phases, so callers (and bench.py) can
 * keep the witness columns in HBM, time the kernel alone, or overlap their own I/O. ---- */
int acvmb_batch_create(acvmb_circuit* c, uint32_t n_instances, acvmb_batch** out);
int acvmb_batch_resize(acvmb_batch* b, uint32_t n_instances);              /* active instances <= created capacity */
void acvmb_batch_destroy(acvmb_batch* b);
int acvmb_batch_upload(acvmb_batch* b, const uint8_t* inputs_be32);         /* H2D + scatter */
int acvmb_batch_run(acvmb_batch* b, float* kernel_ms);                      /* step-VM kernel(s), synchronous */
/* resident inputs: copy input sets to HBM once, then re-run (status reset + scatter + kernel) from HBM;
 * total_ms / vm_ms are CUDA-event times on the library's stream */
int acvmb_batch_stage_inputs(acvmb_batch* b, uint32_t slot, const uint8_t* inputs_be32);
int acvmb_batch_run_staged(acvmb_batch* b, uint32_t slot, float* total_ms, float* vm_ms);
int acvmb_batch_status(acvmb_batch* b, acvmb_status* out_status);           /* D2H of the fail words */
int acvmb_batch_download(acvmb_batch* b, uint32_t first_instance, uint32_t n_instances, const uint32_t* out_ids,
                         uint32_t n_out_ids, uint8_t* out_witness_be32);    /* gather + D2H */
int acvmb_batch_download_ex(acvmb_batch* b, uint32_t first_instance, uint32_t n_instances, const uint32_t* out_ids,
                            uint32_t n_out_ids, uint8_t* out_witness_be32, uint8_t* out_present);
/* per-instance checksum of the solved witness map, computed on the device: out[i] = sum over the witnesses w instance i holds
 * of mix(w, value) mod 2^64, mix = FNV-style fold of the 8 little-endian 32-bit limbs seeded with (w+1)*0x9E3779B97F4A7C15.
 * For size-independent property tests at full batch sizes (no D2H of the witness itself). out: [n_instances] */
int acvmb_batch_checksum(acvmb_batch* b, uint64_t* out);

/* ---- single-instance mirror of the ACVM struct (acvm/src/pwg/mod.rs:129-304), batch of 1 ---- */
int acvmb_vm_new(acvmb_ctx* ctx, const uint8_t* gz_bincode, size_t len, const uint32_t* witness_idx,
                 const uint8_t* witness_be32, uint32_t n_initial, acvmb_vm** out);
void acvmb_vm_destroy(acvmb_vm* vm);
int acvmb_vm_solve(acvmb_vm* vm, acvmb_status* out);                         /* ACVM::solve */
/* ACVM::solve_opcode (acvm/src/pwg/mod.rs:243-303): one opcode per call.  The device runs a re-ordered schedule of the whole
 * circuit, so the first call solves everything and each call then advances the instruction pointer over the solved state:
 * status, instruction_pointer and the witnesses visible through acvmb_vm_witness are those the reference has after the same
 * number of solve_opcode calls.  ACVMB_ERR_STATE once Solved (the reference indexes past `opcodes` and panics). */
int acvmb_vm_solve_opcode(acvmb_vm* vm, acvmb_status* out);
int acvmb_vm_status(const acvmb_vm* vm, acvmb_status* out);                  /* ACVM::get_status */
int acvmb_vm_instruction_pointer(const acvmb_vm* vm, uint32_t* out);         /* ACVM::instruction_pointer */
int acvmb_vm_num_witnesses(const acvmb_vm* vm, uint32_t* out);
/* ACVM::witness_map: copies value of `witness`; *present = 0 when unassigned */
int acvmb_vm_witness(const acvmb_vm* vm, uint32_t witness, uint8_t out_be32[32], int* present);
/* ACVM::finalize: ACVMB_ERR_STATE unless Solved (the reference panics); dense map + presence flags */
int acvmb_vm_finalize(acvmb_vm* vm, uint8_t* out_be32, uint8_t* present, uint32_t n);

/* ACVM::get_pending_foreign_call / resolve_pending_foreign_call (acvm/src/pwg/mod.rs:197-228).  out_lens[i] == 0xFFFFFFFF
 * marks a ForeignCallOutput::Single, any other value an Array of that many values; values are flattened in order. */
int acvmb_vm_pending_foreign_call(const acvmb_vm* vm, char* function, size_t function_cap, uint32_t* n_inputs,
                                  uint32_t* input_lens, uint32_t max_inputs, uint8_t* values_be32, uint32_t max_values,
                                  uint32_t* n_values);
int acvmb_vm_resolve_foreign_call(acvmb_vm* vm, uint32_t n_outputs, const uint32_t* out_lens, const uint8_t* values_be32);

/* WitnessMap on disk: gzip(bincode(BTreeMap<Witness, FieldElement>)) (acir/src/native_types/witness_map.rs:108-146) */
int acvmb_witness_map_compress(const uint32_t* witness_idx, const uint8_t* values_be32, uint32_t n, uint8_t* out, size_t cap,
                               size_t* needed);
int acvmb_witness_map_decompress(const uint8_t* gz, size_t len, uint32_t* witness_idx, uint8_t* values_be32, uint32_t cap,
                                 uint32_t* n);

/* ---- BlackBoxFunctionSolver trait, batched (blackbox_solver/src/lib.rs:27-45) and the free hash
 * functions (:47-60).  Thin wrappers: each builds a one-opcode circuit and runs the same kernels. ---- */
int acvmb_fixed_base_scalar_mul(acvmb_ctx* ctx, const uint8_t* low_be32, const uint8_t* high_be32, uint32_t batch,
                                uint8_t* out_xy_be32 /*[batch][2][32]*/, acvmb_status* out_status);
int acvmb_pedersen(acvmb_ctx* ctx, const uint8_t* inputs_be32 /*[batch][n_inputs][32]*/, uint32_t n_inputs, uint32_t batch,
                   uint32_t domain_separator, uint8_t* out_xy_be32, acvmb_status* out_status);
int acvmb_sha256(acvmb_ctx* ctx, const uint8_t* msgs /*[batch][msg_len]*/, uint32_t msg_len, uint32_t batch,
                 uint8_t* digests /*[batch][32]*/);
int acvmb_keccak256(acvmb_ctx* ctx, const uint8_t* msgs, uint32_t msg_len, uint32_t batch, uint8_t* digests);
/* acvm_blackbox_solver::ecdsa_secp256k1_verify / ecdsa_secp256r1_verify (blackbox_solver/src/lib.rs:67-83,101-210), batched.
 * out_valid[i] = 1 / 0.  Inputs on which the reference panics (r or s outside [1, n-1], x not on the curve, hash >= n,
 * R at infinity) give out_status[i] = {ACVMB_FAILURE, ACVMB_E_REFERENCE_PANIC} and out_valid[i] = 0.  As in the reference,
 * only the parity of public_key_y is used (the key is rebuilt from its compressed encoding). */
int acvmb_ecdsa_secp256k1_verify(acvmb_ctx* ctx, const uint8_t* hashed_msg /*[batch][32]*/, const uint8_t* public_key_x /*[batch][32]*/,
                                 const uint8_t* public_key_y /*[batch][32]*/, const uint8_t* signature /*[batch][64]*/, uint32_t batch,
                                 uint8_t* out_valid /*[batch]*/, acvmb_status* out_status);
int acvmb_ecdsa_secp256r1_verify(acvmb_ctx* ctx, const uint8_t* hashed_msg, const uint8_t* public_key_x, const uint8_t* public_key_y,
                                 const uint8_t* signature, uint32_t batch, uint8_t* out_valid, acvmb_status* out_status);

/* ---- host-only: decode + compile without a device (CPU tests of the decoder / plan compiler) ---- */
int acvmb_plan_compile_host(const uint8_t* gz_bincode, size_t len, const uint32_t* input_witnesses, uint32_t n_inputs, uint32_t S,
                            acvmb_plan_info* info, uint8_t* blob, size_t cap, size_t* needed);
/* same with plan options: temp_pool (0 = default), flags bit 0 = accept Pedersen (parity unpinned, see "pedersen_unpinned"),
 * bit 1 = keep every Brillig opcode on the host VM (no plan-time lowering to device gates), bit 2 = canonical columns only, bit 3 = one micro-op per
 * hash call (no pack / core / unpack split), bit 4 = no spreading of heavy micro-ops over warps ("spread_heavy" = 0),
 * bit 5 = curve micro-ops with slack keep steps of their own ("slack_scheduling" = 0), bits 8..23 = entries of the shared-memory
 * ring of recent values (0 = default, 0xFFFF = none) */
int acvmb_plan_compile_host_ex(const uint8_t* gz_bincode, size_t len, const uint32_t* input_witnesses, uint32_t n_inputs, uint32_t S,
                               uint32_t temp_pool, uint32_t flags, acvmb_plan_info* info, uint8_t* blob, size_t cap, size_t* needed);

int acvmb_pedersen_generator_host(uint32_t index, uint8_t out_xy_be32[64]);

/* host-only: switch settings of the permutation network mapping 0..n-1 onto `outputs`, i.e. sorting::route
 * (acvm/src/pwg/directives/sorting.rs:164-235) as Directive::PermutationSort calls it (directives/mod.rs:113-114).
 * bits (one byte per switch) may be NULL to query *n_bits. */
int acvmb_permutation_route_host(const uint32_t* outputs, uint32_t n, uint8_t* bits, uint32_t cap, uint32_t* n_bits);
/* run one Brillig opcode of a circuit on the host VM (status: 0 finished, 1 failure, 2 foreign-call wait, 3 reference panic) */
int acvmb_brillig_run_host(const uint8_t* gz_bincode, size_t len, uint32_t opcode_index, const uint8_t* in_values_be32,
                           uint32_t n_in_values, uint8_t* out_values_be32, uint32_t n_out_values, uint32_t* status,
                           uint32_t* fail_pc);

/* ---- measurement helpers ---- */
int acvmb_imad_microbench(acvmb_ctx* ctx, double* imad32_per_s, double* imad_wide_per_s, double* imad_wide_carry_per_s,
                          double* sm_clock_mhz);
/* bare device->host copy rate (GB/s) of the context's device into `host` (pinned memory from acvmb_host_alloc): the ceiling of
 * every end-to-end number that returns witness maps */
int acvmb_d2h_microbench(acvmb_ctx* ctx, void* host, size_t bytes, uint32_t reps, double* gb_per_s);
/* register-resident Montgomery multiplications per second for the 5 FMA/ALU pipe-split levels of the K0 field
 * library (fr_mul_per_s[0..4]): the practical Fr-mul ceiling */
int acvmb_frmul_microbench(acvmb_ctx* ctx, double* fr_mul_per_s);
/* wide IMADs per second of three carry-OUT-only forms: [0] + addc capture per product, [1] carry dropped, [2] chains of two */
int acvmb_imad_cc_microbench(acvmb_ctx* ctx, double* out3);
/* tuning knobs: "T" (instances per CTA), "S" (slots per step), "chunk_steps", "n_stage", "split", "max_resident_bytes",
 * "staging_bytes", "split_curve" (0: one micro-op per curve call), "temp_pool" (temporary columns), "pedersen_unpinned" (1: accept
 * BlackBoxFuncCall::Pedersen / acvmb_pedersen although the values are NOT barretenberg's -- refused by default),
 * "device_brillig" (0: every Brillig opcode runs on the host VM), "scaled_columns" (0: every witness column holds the
 * canonical value -- two Montgomery reductions per multiplicative gate instead of one), "packed_hashes" (0: a hash call is ONE micro-op that gathers its
 * message byte by byte and scatters its digest itself), "sha_pad_table" (0: the padding-only last block of a SHA256 call over k * 64 bytes
 * recomputes its message schedule instead of reading the plan's constant table), "ring_bytes" (shared memory per CTA
 * for the ring of recent values that serves operand reads on chip; 0: every operand comes from L2 / HBM), "spread_heavy" (0: the micro-ops of a
 * step fill its slots in sorted order; default 1 puts the hash / curve / general micro-ops of one step into different warps when
 * the tile is narrower than a warp, so that they do not serialise inside one), "slack_scheduling" (0: curve micro-ops that nothing
 * needs for two or more dependency levels -- the H1 sums of a Pedersen chain -- run up front in steps of their own; default 1 moves
 * them into the idle slots of the level before their first use), "cache_batch" (0: free
 * the column buffers at the end of every acvmb_solve_batch; default 1 keeps those of the last call, per context, for an
 * identical next call); plan options apply to circuits created afterwards */
int acvmb_ctx_set_option(acvmb_ctx* ctx, const char* key, uint64_t value);

#ifdef __cplusplus
}
#endif
#endif /* ACVM_B200_H */
