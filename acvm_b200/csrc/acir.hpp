// ACIR data model + decoder for the reference wire format.
//
// Replaces (reference): acir::circuit::Circuit::read  (acir/src/circuit/mod.rs:155-161) =
// GzDecoder -> bincode 1.3.3 defaults -> Circuit.  Struct / enum declaration order follows
//   Circuit               acir/src/circuit/mod.rs:18-41
//   Opcode                acir/src/circuit/opcodes.rs:15-34
//   Expression            acir/src/native_types/expression/mod.rs:17-28
//   BlackBoxFuncCall      acir/src/circuit/opcodes/black_box_function_call.rs:20-115
//   Directive             acir/src/circuit/directives.rs:5-36
//   Brillig               acir/src/circuit/brillig.rs:9-33, brillig/src/opcodes.rs:60-134
// Only the data model the solver needs is kept (no Expression algebra, no Display).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <memory>
#include <vector>

#include "fr_host.hpp"

namespace acvmb {

struct DecodeError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct MulTerm {
    U256 c;
    uint32_t a, b;
};
struct LinTerm {
    U256 c;
    uint32_t w;
};
struct Expression {
    std::vector<MulTerm> mul_terms;
    std::vector<LinTerm> linear_combinations;
    U256 q_c;
    bool present = true;  // false == Option::None when used as Option<Expression>
};

struct FunctionInput {
    uint32_t witness, num_bits;
};

// declaration order == bincode tag
enum BlackBoxFunc : uint32_t {
    BB_AND = 0, BB_XOR, BB_RANGE, BB_SHA256, BB_Blake2s, BB_SchnorrVerify, BB_Pedersen, BB_HashToField128Security,
    BB_EcdsaSecp256k1, BB_EcdsaSecp256r1, BB_FixedBaseScalarMul, BB_Keccak256, BB_Keccak256VariableLength,
    BB_RecursiveAggregation, BB_COUNT
};
const char* blackbox_name(uint32_t f);

struct BlackBoxCall {
    uint32_t func = 0;
    // all FunctionInputs in get_inputs_vec() order (black_box_function_call.rs:205-292)
    std::vector<FunctionInput> inputs;
    // all output witnesses in get_outputs_vec() order
    std::vector<uint32_t> outputs;
    uint32_t domain_separator = 0;    // Pedersen
    uint32_t n_message_inputs = 0;    // hash inputs that form the message (excludes var_message_size)
    // SchnorrVerify / ECDSA keep their segment sizes for completeness
    uint32_t seg[4] = {0, 0, 0, 0};
};

enum DirectiveKind : uint32_t { DIR_Quotient = 0, DIR_ToLeRadix = 1, DIR_PermutationSort = 2 };
struct Directive {
    uint32_t kind = 0;
    Expression a, b;              // Quotient: a, b ; ToLeRadix: a
    uint32_t q = 0, r = 0;        // Quotient
    Expression predicate;         // Quotient (Option)
    std::vector<uint32_t> out;    // ToLeRadix: b ; PermutationSort: bits
    uint32_t radix = 0;
    std::vector<std::vector<Expression>> sort_inputs;
    uint32_t tuple = 0;
    std::vector<uint32_t> sort_by;
};

struct RegOrMem {
    uint32_t kind;  // 0 register, 1 heap array, 2 heap vector
    uint64_t a, b;
};
struct BrilligOp {
    uint32_t tag = 0;
    uint64_t r0 = 0, r1 = 0, r2 = 0;   // register / label operands, meaning per tag
    uint32_t bop = 0, bit_size = 0;
    U256 value;                        // Const
    std::string function;              // ForeignCall
    std::vector<RegOrMem> destinations, inputs;
    // BlackBox
    uint32_t bb_tag = 0;
    uint64_t bb[10] = {0};
};
struct BrilligInput {
    bool is_array;
    std::vector<Expression> exprs;  // 1 for Single
};
struct BrilligOutput {
    bool is_array;
    std::vector<uint32_t> witnesses;  // 1 for Simple
};
struct ForeignCallOutput {
    bool is_array;
    std::vector<U256> values;
};
struct Brillig {
    std::vector<BrilligInput> inputs;
    std::vector<BrilligOutput> outputs;
    std::vector<std::vector<ForeignCallOutput>> foreign_call_results;
    std::vector<BrilligOp> bytecode;
    Expression predicate;
};

struct MemOp {
    uint32_t block_id = 0;
    Expression operation, index, value, predicate;
};

enum OpcodeKind : uint32_t { OP_Arithmetic = 0, OP_BlackBox = 1, OP_Directive = 2, OP_Brillig = 3, OP_MemoryOp = 4, OP_MemoryInit = 5 };

// Everything but an Arithmetic opcode's expression lives out of line: Arithmetic opcodes are the bulk of every circuit, and
// with the blackbox / directive / Brillig / memory payloads inline one opcode was 1.1 KB (1.2 GB for a 2^20-gate circuit).
struct OpcodeExtra {
    BlackBoxCall bb;              // BlackBoxFuncCall
    Directive dir;                // Directive
    Brillig brillig;              // Brillig
    MemOp mem;                    // MemoryOp
    uint32_t block_id = 0;        // MemoryInit
    std::vector<uint32_t> init;   // MemoryInit
};

struct Opcode {
    uint32_t kind = 0;
    Expression expr;              // Arithmetic

    BlackBoxCall& bb() { return ext().bb; }
    const BlackBoxCall& bb() const { return ext().bb; }
    Directive& dir() { return ext().dir; }
    const Directive& dir() const { return ext().dir; }
    Brillig& brillig() { return ext().brillig; }
    const Brillig& brillig() const { return ext().brillig; }
    MemOp& mem() { return ext().mem; }
    const MemOp& mem() const { return ext().mem; }
    uint32_t& block_id() { return ext().block_id; }
    uint32_t block_id() const { return ext().block_id; }
    std::vector<uint32_t>& init() { return ext().init; }
    const std::vector<uint32_t>& init() const { return ext().init; }

    Opcode() = default;
    Opcode(Opcode&&) = default;
    Opcode& operator=(Opcode&&) = default;
    Opcode(const Opcode& o) : kind(o.kind), expr(o.expr), x_(o.x_ ? new OpcodeExtra(*o.x_) : nullptr) {}
    Opcode& operator=(const Opcode& o) {
        if (this != &o) {
            kind = o.kind;
            expr = o.expr;
            x_.reset(o.x_ ? new OpcodeExtra(*o.x_) : nullptr);
        }
        return *this;
    }

   private:
    OpcodeExtra& ext() {
        if (!x_) x_.reset(new OpcodeExtra);
        return *x_;
    }
    const OpcodeExtra& ext() const {
        static const OpcodeExtra empty;
        return x_ ? *x_ : empty;
    }
    std::unique_ptr<OpcodeExtra> x_;
};

struct AssertMessage {
    uint32_t loc_kind;
    uint64_t acir_index, brillig_index;
    std::string message;
};

struct Circuit {
    uint32_t current_witness_index = 0;
    std::vector<Opcode> opcodes;
    std::vector<uint32_t> private_parameters, public_parameters, return_values;
    std::vector<AssertMessage> assert_messages;
};

// gunzip + bincode decode; throws DecodeError
Circuit decode_circuit(const uint8_t* data, size_t len);
std::vector<uint8_t> gunzip(const uint8_t* data, size_t len);
std::vector<uint8_t> gzip_bytes(const std::vector<uint8_t>& raw);

// WitnessMap on disk (acir/src/native_types/witness_map.rs:108-146): gzip(bincode(BTreeMap<u32, hex string>))
std::vector<std::pair<uint32_t, U256>> decode_witness_map(const uint8_t* data, size_t len);
std::vector<uint8_t> encode_witness_map(const std::vector<std::pair<uint32_t, U256>>& wm);

}  // namespace acvmb
