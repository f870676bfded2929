// ACIR wire-format decoder (see acir.hpp for the reference citations).
#include "acir.hpp"

#include <zlib.h>

#include <algorithm>

#include <cstring>

namespace acvmb {

static const char* kBlackBoxNames[BB_COUNT] = {
    "and", "xor", "range", "sha256", "blake2s", "schnorr_verify", "pedersen", "hash_to_field_128_security",
    "ecdsa_secp256k1", "ecdsa_secp256r1", "fixed_base_scalar_mul", "keccak256", "keccak256_variable_length",
    "recursive_aggregation"};

const char* blackbox_name(uint32_t f) { return f < BB_COUNT ? kBlackBoxNames[f] : "?"; }

std::vector<uint8_t> gunzip(const uint8_t* data, size_t len) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 15 + 16) != Z_OK) throw DecodeError("inflateInit2 failed");
    zs.next_in = const_cast<Bytef*>(data);
    size_t left = len;   // fed in pieces: avail_in is 32 bits wide
    std::vector<uint8_t> out;
    out.reserve(len * 2);
    std::vector<uint8_t> buf(1 << 20);
    int rc;
    do {
        if (zs.avail_in == 0 && left) {
            const size_t piece = std::min<size_t>(left, (size_t)1 << 30);
            zs.avail_in = (uInt)piece;
            left -= piece;
        }
        zs.next_out = buf.data();
        zs.avail_out = (uInt)buf.size();
        rc = inflate(&zs, Z_NO_FLUSH);
        if ((rc != Z_OK && rc != Z_STREAM_END) || (rc == Z_OK && zs.avail_out == buf.size() && zs.avail_in == 0 && left == 0)) {
            inflateEnd(&zs);
            throw DecodeError("gzip stream is corrupt");
        }
        out.insert(out.end(), buf.data(), buf.data() + (buf.size() - zs.avail_out));
    } while (rc != Z_STREAM_END);
    inflateEnd(&zs);
    return out;
}

std::vector<uint8_t> gzip_bytes(const std::vector<uint8_t>& raw) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, Z_BEST_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK)
        throw DecodeError("deflateInit2 failed");
    std::vector<uint8_t> out(deflateBound(&zs, raw.size()) + 32);
    zs.next_in = const_cast<Bytef*>(raw.data());
    zs.avail_in = (uInt)raw.size();
    zs.next_out = out.data();
    zs.avail_out = (uInt)out.size();
    int rc = deflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END) {
        deflateEnd(&zs);
        throw DecodeError("deflate failed");
    }
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

namespace {

struct Reader {
    const uint8_t* d;
    size_t n, o = 0;
    void need(size_t k) const {
        if (o + k > n) throw DecodeError("unexpected end of bincode stream");
    }
    uint8_t u8() {
        need(1);
        return d[o++];
    }
    uint32_t u32() {
        need(4);
        uint32_t v;
        memcpy(&v, d + o, 4);
        o += 4;
        return v;
    }
    uint64_t u64() {
        need(8);
        uint64_t v;
        memcpy(&v, d + o, 8);
        o += 8;
        return v;
    }
    uint64_t len() {
        uint64_t v = u64();
        if (v > n) throw DecodeError("implausible length prefix");
        return v;
    }
    std::string str() {
        uint64_t k = len();
        need(k);
        std::string s((const char*)d + o, k);
        o += k;
        return s;
    }
    U256 fe() {  // FieldElement serialises as its hex string (acir_field/src/generic_ark.rs:114-134)
        U256 v;
        uint64_t k = len();
        need(k);
        if (!hf::from_hex((const char*)d + o, k, v)) throw DecodeError("bad field element hex");
        o += k;
        return v;
    }
    std::vector<uint32_t> vec_u32() {
        uint64_t k = len();
        std::vector<uint32_t> v(k);
        for (auto& x : v) x = u32();
        return v;
    }
};

Expression r_expr(Reader& r) {
    Expression e;
    uint64_t k = r.len();
    e.mul_terms.resize(k);
    for (auto& t : e.mul_terms) {
        t.c = r.fe();
        t.a = r.u32();
        t.b = r.u32();
    }
    k = r.len();
    e.linear_combinations.resize(k);
    for (auto& t : e.linear_combinations) {
        t.c = r.fe();
        t.w = r.u32();
    }
    e.q_c = r.fe();
    return e;
}

Expression r_opt_expr(Reader& r) {
    uint8_t t = r.u8();
    if (t == 0) {
        Expression e;
        e.present = false;
        return e;
    }
    if (t != 1) throw DecodeError("bad Option tag");
    return r_expr(r);
}

FunctionInput r_fi(Reader& r) {
    FunctionInput f;
    f.witness = r.u32();
    f.num_bits = r.u32();
    return f;
}
void r_vfi(Reader& r, std::vector<FunctionInput>& out, uint32_t* count = nullptr) {
    uint64_t k = r.len();
    for (uint64_t i = 0; i < k; ++i) out.push_back(r_fi(r));
    if (count) *count = (uint32_t)k;
}

BlackBoxCall r_blackbox(Reader& r) {
    BlackBoxCall b;
    b.func = r.u32();
    switch (b.func) {
        case BB_AND:
        case BB_XOR:
            b.inputs.push_back(r_fi(r));
            b.inputs.push_back(r_fi(r));
            b.outputs.push_back(r.u32());
            break;
        case BB_RANGE:
            b.inputs.push_back(r_fi(r));
            break;
        case BB_SHA256:
        case BB_Blake2s:
        case BB_Keccak256:
            r_vfi(r, b.inputs, &b.n_message_inputs);
            b.outputs = r.vec_u32();
            break;
        case BB_SchnorrVerify:
            b.inputs.push_back(r_fi(r));
            b.inputs.push_back(r_fi(r));
            r_vfi(r, b.inputs, &b.seg[0]);
            r_vfi(r, b.inputs, &b.seg[1]);
            b.outputs.push_back(r.u32());
            break;
        case BB_Pedersen:
            r_vfi(r, b.inputs, &b.n_message_inputs);
            b.domain_separator = r.u32();
            b.outputs.push_back(r.u32());
            b.outputs.push_back(r.u32());
            break;
        case BB_HashToField128Security:
            r_vfi(r, b.inputs, &b.n_message_inputs);
            b.outputs.push_back(r.u32());
            break;
        case BB_EcdsaSecp256k1:
        case BB_EcdsaSecp256r1:
            for (int i = 0; i < 4; ++i) r_vfi(r, b.inputs, &b.seg[i]);
            b.outputs.push_back(r.u32());
            break;
        case BB_FixedBaseScalarMul:
            b.inputs.push_back(r_fi(r));
            b.inputs.push_back(r_fi(r));
            b.outputs.push_back(r.u32());
            b.outputs.push_back(r.u32());
            break;
        case BB_Keccak256VariableLength:
            r_vfi(r, b.inputs, &b.n_message_inputs);
            b.inputs.push_back(r_fi(r));  // var_message_size is the last input (get_inputs_vec)
            b.outputs = r.vec_u32();
            break;
        case BB_RecursiveAggregation: {
            r_vfi(r, b.inputs, &b.seg[0]);
            r_vfi(r, b.inputs, &b.seg[1]);
            r_vfi(r, b.inputs, &b.seg[2]);
            b.inputs.push_back(r_fi(r));
            uint8_t t = r.u8();
            if (t == 1) {
                std::vector<FunctionInput> ignored;  // input aggregation object: not an input (:276-279)
                r_vfi(r, ignored);
            } else if (t != 0) {
                throw DecodeError("bad Option tag");
            }
            b.outputs = r.vec_u32();
            break;
        }
        default:
            throw DecodeError("unknown BlackBoxFuncCall tag");
    }
    return b;
}

Directive r_directive(Reader& r) {
    Directive d;
    d.kind = r.u32();
    switch (d.kind) {
        case DIR_Quotient:
            d.a = r_expr(r);
            d.b = r_expr(r);
            d.q = r.u32();
            d.r = r.u32();
            d.predicate = r_opt_expr(r);
            break;
        case DIR_ToLeRadix:
            d.a = r_expr(r);
            d.out = r.vec_u32();
            d.radix = r.u32();
            break;
        case DIR_PermutationSort: {
            uint64_t k = r.len();
            d.sort_inputs.resize(k);
            for (auto& v : d.sort_inputs) {
                uint64_t m = r.len();
                for (uint64_t i = 0; i < m; ++i) v.push_back(r_expr(r));
            }
            d.tuple = r.u32();
            d.out = r.vec_u32();
            d.sort_by = r.vec_u32();
            break;
        }
        default:
            throw DecodeError("unknown Directive tag");
    }
    return d;
}

RegOrMem r_rom(Reader& r) {
    RegOrMem m;
    m.kind = r.u32();
    if (m.kind > 2) throw DecodeError("bad RegisterOrMemory tag");
    m.a = r.u64();
    m.b = m.kind == 0 ? 0 : r.u64();
    return m;
}

BrilligOp r_brillig_op(Reader& r) {
    BrilligOp o;
    o.tag = r.u32();
    switch (o.tag) {
        case 0:  // BinaryFieldOp{destination, op, lhs, rhs}
            o.r0 = r.u64();
            o.bop = r.u32();
            o.r1 = r.u64();
            o.r2 = r.u64();
            break;
        case 1:  // BinaryIntOp{destination, op, bit_size, lhs, rhs}
            o.r0 = r.u64();
            o.bop = r.u32();
            o.bit_size = r.u32();
            o.r1 = r.u64();
            o.r2 = r.u64();
            break;
        case 2:
        case 3:  // JumpIfNot / JumpIf {condition, location}
            o.r0 = r.u64();
            o.r1 = r.u64();
            break;
        case 4:
        case 5:  // Jump / Call {location}
            o.r0 = r.u64();
            break;
        case 6:  // Const{destination, value}
            o.r0 = r.u64();
            o.value = r.fe();
            break;
        case 7:
            break;  // Return
        case 8: {   // ForeignCall{function, destinations, inputs}
            o.function = r.str();
            uint64_t k = r.len();
            for (uint64_t i = 0; i < k; ++i) o.destinations.push_back(r_rom(r));
            k = r.len();
            for (uint64_t i = 0; i < k; ++i) o.inputs.push_back(r_rom(r));
            break;
        }
        case 9:
        case 10:
        case 11:  // Mov / Load / Store : two register operands
            o.r0 = r.u64();
            o.r1 = r.u64();
            break;
        case 12: {  // BlackBox(BlackBoxOp)  brillig/src/black_box.rs:7-48
            o.bb_tag = r.u32();
            static const int nwords[9] = {4, 4, 4, 3, 9, 9, 7, 5, 4};
            if (o.bb_tag > 8) throw DecodeError("bad brillig BlackBoxOp tag");
            for (int i = 0; i < nwords[o.bb_tag]; ++i) o.bb[i] = r.u64();
            break;
        }
        case 13:
        case 14:
            break;  // Trap / Stop
        default:
            throw DecodeError("unknown brillig opcode tag");
    }
    return o;
}

Brillig r_brillig(Reader& r) {
    Brillig b;
    uint64_t k = r.len();
    for (uint64_t i = 0; i < k; ++i) {
        BrilligInput in;
        uint32_t t = r.u32();
        if (t == 0) {
            in.is_array = false;
            in.exprs.push_back(r_expr(r));
        } else if (t == 1) {
            in.is_array = true;
            uint64_t m = r.len();
            for (uint64_t j = 0; j < m; ++j) in.exprs.push_back(r_expr(r));
        } else {
            throw DecodeError("bad BrilligInputs tag");
        }
        b.inputs.push_back(std::move(in));
    }
    k = r.len();
    for (uint64_t i = 0; i < k; ++i) {
        BrilligOutput out;
        uint32_t t = r.u32();
        if (t == 0) {
            out.is_array = false;
            out.witnesses.push_back(r.u32());
        } else if (t == 1) {
            out.is_array = true;
            out.witnesses = r.vec_u32();
        } else {
            throw DecodeError("bad BrilligOutputs tag");
        }
        b.outputs.push_back(std::move(out));
    }
    k = r.len();
    for (uint64_t i = 0; i < k; ++i) {
        std::vector<ForeignCallOutput> res;
        uint64_t m = r.len();
        for (uint64_t j = 0; j < m; ++j) {
            ForeignCallOutput f;
            uint32_t t = r.u32();
            if (t == 0) {
                f.is_array = false;
                f.values.push_back(r.fe());
            } else if (t == 1) {
                f.is_array = true;
                uint64_t q = r.len();
                for (uint64_t z = 0; z < q; ++z) f.values.push_back(r.fe());
            } else {
                throw DecodeError("bad ForeignCallOutput tag");
            }
            res.push_back(std::move(f));
        }
        b.foreign_call_results.push_back(std::move(res));
    }
    k = r.len();
    for (uint64_t i = 0; i < k; ++i) b.bytecode.push_back(r_brillig_op(r));
    b.predicate = r_opt_expr(r);
    return b;
}

void r_opcode(Reader& r, Opcode& op) {
    op.kind = r.u32();
    switch (op.kind) {
        case OP_Arithmetic:
            op.expr = r_expr(r);
            break;
        case OP_BlackBox:
            op.bb() = r_blackbox(r);
            break;
        case OP_Directive:
            op.dir() = r_directive(r);
            break;
        case OP_Brillig:
            op.brillig() = r_brillig(r);
            break;
        case OP_MemoryOp:
            op.mem().block_id = r.u32();
            op.mem().operation = r_expr(r);
            op.mem().index = r_expr(r);
            op.mem().value = r_expr(r);
            op.mem().predicate = r_opt_expr(r);
            break;
        case OP_MemoryInit:
            op.block_id() = r.u32();
            op.init() = r.vec_u32();
            break;
        default:
            throw DecodeError("unknown Opcode tag");
    }
}

}  // namespace

Circuit decode_circuit(const uint8_t* data, size_t len) {
    std::vector<uint8_t> raw = gunzip(data, len);
    Reader r{raw.data(), raw.size()};
    Circuit c;
    c.current_witness_index = r.u32();
    uint64_t k = r.len();
    c.opcodes.resize(k);
    for (auto& op : c.opcodes) r_opcode(r, op);
    c.private_parameters = r.vec_u32();
    c.public_parameters = r.vec_u32();
    c.return_values = r.vec_u32();
    k = r.len();
    for (uint64_t i = 0; i < k; ++i) {
        AssertMessage m;
        m.loc_kind = r.u32();
        if (m.loc_kind > 1) throw DecodeError("bad OpcodeLocation tag");
        m.acir_index = r.u64();
        m.brillig_index = m.loc_kind == 1 ? r.u64() : 0;
        m.message = r.str();
        c.assert_messages.push_back(std::move(m));
    }
    if (r.o != r.n) throw DecodeError("trailing bytes after Circuit");
    return c;
}

std::vector<std::pair<uint32_t, U256>> decode_witness_map(const uint8_t* data, size_t len) {
    std::vector<uint8_t> raw = gunzip(data, len);
    Reader r{raw.data(), raw.size()};
    uint64_t k = r.len();
    std::vector<std::pair<uint32_t, U256>> out;
    out.reserve(k);
    for (uint64_t i = 0; i < k; ++i) {
        uint32_t w = r.u32();
        out.emplace_back(w, r.fe());
    }
    if (r.o != r.n) throw DecodeError("trailing bytes after WitnessMap");
    return out;
}

std::vector<uint8_t> encode_witness_map(const std::vector<std::pair<uint32_t, U256>>& wm) {
    std::vector<uint8_t> raw;
    auto put64 = [&](uint64_t v) {
        for (int i = 0; i < 8; ++i) raw.push_back((uint8_t)(v >> (8 * i)));
    };
    put64(wm.size());
    static const char* hexd = "0123456789abcdef";
    for (auto& kv : wm) {
        for (int i = 0; i < 4; ++i) raw.push_back((uint8_t)(kv.first >> (8 * i)));
        put64(64);
        uint8_t be[32];
        hf::to_be_bytes(kv.second, be);
        for (int i = 0; i < 32; ++i) {
            raw.push_back((uint8_t)hexd[be[i] >> 4]);
            raw.push_back((uint8_t)hexd[be[i] & 15]);
        }
    }
    return gzip_bytes(raw);
}

}  // namespace acvmb
