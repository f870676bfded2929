// CPU "reference-algorithm restatement" of acvm::pwg for timing and cross-checks.
//
// TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  Never linked into libacvm_b200.so.
// The real reference (Rust, acvm 0.27.0) cannot be built in this environment (no cargo/rustc), so
// this file restates its algorithm AS WRITTEN, including the costs that dominate it on a CPU:
//   * WitnessMap = BTreeMap<Witness, FieldElement>   (acir/src/native_types/witness_map.rs:42)   -> std::map
//   * ArithmeticSolver::evaluate allocates a fresh Expression (two Vecs) per opcode (arithmetic.rs:212-239)
//   * the solved witness is -(sum / coeff): one field inversion PER GATE (arithmetic.rs:103-125,
//     generic_ark.rs:375-380); ark-ff 0.4.2 Fp::inverse = binary extended Euclid (Guajardo et al. Alg. 16)
//   * FieldElement::is_zero compares serialised bytes (generic_ark.rs:88-92,164-166)
//   * Fr = 4 x 64-bit Montgomery limbs, portable u64/u128 path (ark-ff without the `asm` feature)
// One solver instance per thread; `threads` instances run concurrently (the reference itself is
// single-threaded per ACVM).
//
// Opcode stream format (built by oracle/cref.py from the oracle's own decoder), u64 words:
//   kind(0=Arithmetic,1=AND,2=XOR,3=RANGE) ...
//   Arithmetic: n_mul, n_lin, then n_mul*(c[4],a,b), n_lin*(c[4],w), qc[4]
//   AND/XOR:    lhs, rhs, num_bits, out          RANGE: w, num_bits
#include <cstdint>
#include <cstring>
#include <map>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;

namespace {

struct Fr { uint64_t l[4]; };  // Montgomery form

const uint64_t Pm[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
const uint64_t INV = 0xc2e1f593efffffffULL;
const Fr R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
const Fr ONE_RAW = {{1, 0, 0, 0}};

inline bool geq_p(const uint64_t* a) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > Pm[i]) return true;
        if (a[i] < Pm[i]) return false;
    }
    return true;
}
inline void sub_p(uint64_t* a) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a[i] - Pm[i] - br;
        a[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
}
inline Fr add(const Fr& a, const Fr& b) {
    Fr r;
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_p(r.l)) sub_p(r.l);
    return r;
}
inline Fr neg(const Fr& a) {
    if ((a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0) return a;
    Fr r;
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)Pm[i] - a.l[i] - br;
        r.l[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
    return r;
}
inline Fr mul(const Fr& a, const Fr& b) {  // CIOS Montgomery
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * INV;
        c = (u128)m * Pm[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * Pm[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_p(r.l)) sub_p(r.l);
    return r;
}
inline Fr from_canonical(const uint64_t* v) {
    Fr a = {{v[0], v[1], v[2], v[3]}};
    return mul(a, R2);
}
inline Fr to_canonical(const Fr& a) { return mul(a, ONE_RAW); }

// FieldElement::is_zero / PartialEq go through to_be_bytes(): Montgomery -> canonical -> Vec<u8>
inline bool is_zero(const Fr& a) {
    Fr c = to_canonical(a);
    std::vector<uint8_t> bytes(32);
    memcpy(bytes.data(), c.l, 32);
    for (uint8_t b : bytes)
        if (b) return false;
    return true;
}
inline bool eq(const Fr& a, const Fr& b) {
    Fr x = to_canonical(a), y = to_canonical(b);
    return memcmp(x.l, y.l, 32) == 0;
}

// ---- binary extended Euclid inversion on 256-bit ints (ark-ff Fp::inverse, Guajardo et al. Alg. 16) ----
inline bool is_even(const uint64_t* a) { return (a[0] & 1) == 0; }
inline void shr1(uint64_t* a) {
    for (int i = 0; i < 3; ++i) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
    a[3] >>= 1;
}
inline bool is_one(const uint64_t* a) { return a[0] == 1 && !a[1] && !a[2] && !a[3]; }
inline int cmp(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return -1;
    }
    return 0;
}
inline uint64_t add_n(uint64_t* a, const uint64_t* b) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a[i] + b[i];
        a[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
inline void sub_n(uint64_t* a, const uint64_t* b) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a[i] - b[i] - br;
        a[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
}
inline Fr inverse(const Fr& a) {  // returns 0 for 0 (generic_ark.rs:242-245)
    if ((a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0) return a;
    // operates on the Montgomery representation: result = a^{-1} * R^2 * ... handled like ark-ff (b = R2, c = 0)
    uint64_t u[4], v[4];
    memcpy(u, a.l, 32);
    memcpy(v, Pm, 32);
    Fr b = R2, c = {{0, 0, 0, 0}};
    while (!is_one(u) && !is_one(v)) {
        while (is_even(u)) {
            shr1(u);
            if (is_even(b.l)) shr1(b.l);
            else {
                uint64_t carry = add_n(b.l, Pm);
                shr1(b.l);
                if (carry) b.l[3] |= 1ULL << 63;
            }
        }
        while (is_even(v)) {
            shr1(v);
            if (is_even(c.l)) shr1(c.l);
            else {
                uint64_t carry = add_n(c.l, Pm);
                shr1(c.l);
                if (carry) c.l[3] |= 1ULL << 63;
            }
        }
        if (cmp(v, u) < 0) {
            sub_n(u, v);
            // b -= c (mod p)
            if (cmp(b.l, c.l) < 0) add_n(b.l, Pm);
            sub_n(b.l, c.l);
        } else {
            sub_n(v, u);
            if (cmp(c.l, b.l) < 0) add_n(c.l, Pm);
            sub_n(c.l, b.l);
        }
    }
    return is_one(u) ? b : c;
}
inline Fr divf(const Fr& a, const Fr& b) { return mul(a, inverse(b)); }  // generic_ark.rs:375-380

struct MulT { Fr c; uint32_t a, b; };
struct LinT { Fr c; uint32_t w; };
struct Expr {
    std::vector<MulT> mul;
    std::vector<LinT> lin;
    Fr qc;
};
struct Op {
    uint32_t kind;
    Expr e;
    uint32_t a, b, nb, out;
};
typedef std::map<uint32_t, Fr> WMap;

enum { ST_SOLVED = 0, ST_FAILURE = 2 };
enum { E_NONE = 0, E_MISSING = 1, E_TOO_MANY = 2, E_UNSAT = 4, E_PANIC = 8 };

struct Result { uint32_t code, err, opcode, aux; };

// arithmetic.rs:212-239
Expr evaluate(const Expr& e, const WMap& wm) {
    Expr r;
    r.qc = Fr{{0, 0, 0, 0}};
    for (const MulT& t : e.mul) {
        auto l = wm.find(t.a), rr = wm.find(t.b);
        bool kl = l != wm.end(), kr = rr != wm.end();
        if (kl && kr) r.qc = add(r.qc, mul(mul(t.c, l->second), rr->second));
        else if (!kl && !kr) { if (!is_zero(t.c)) r.mul.push_back(t); }
        else {
            Fr v = mul(t.c, kl ? l->second : rr->second);
            if (!is_zero(v)) r.lin.push_back(LinT{v, kl ? t.b : t.a});
        }
    }
    for (const LinT& t : e.lin) {
        auto it = wm.find(t.w);
        if (it != wm.end()) r.qc = add(r.qc, mul(t.c, it->second));
        else if (!is_zero(t.c)) r.lin.push_back(t);
    }
    r.qc = add(r.qc, e.qc);
    return r;
}

// mod.rs:338-357
int insert_value(uint32_t w, const Fr& v, WMap& wm) {
    auto it = wm.find(w);
    if (it == wm.end()) { wm.emplace(w, v); return E_NONE; }
    Fr old = it->second;
    it->second = v;
    return eq(old, v) ? E_NONE : E_UNSAT;
}

// arithmetic.rs:27-127 (the MulTerm::OneUnknown arms are unreachable after evaluate())
int solve_arith(WMap& wm, const Expr& expr) {
    Expr op = evaluate(expr, wm);
    if (op.mul.size() > 1) return E_PANIC;
    bool mul_too_many = op.mul.size() == 1;
    // solve_fan_in_term :176-209
    int unknowns = 0;
    LinT unk{};
    Fr sum = {{0, 0, 0, 0}};
    for (const LinT& t : op.lin) {
        auto it = wm.find(t.w);
        if (it != wm.end()) sum = add(sum, mul(t.c, it->second));
        else { unk = t; ++unknowns; }
        if (unknowns > 1) break;
    }
    if (mul_too_many || unknowns > 1) return E_TOO_MANY;
    Fr total = add(sum, op.qc);
    if (unknowns == 0) return is_zero(total) ? E_NONE : E_UNSAT;
    if (is_zero(unk.c)) return is_zero(total) ? E_NONE : E_UNSAT;
    return insert_value(unk.w, neg(divf(total, unk.c)), wm);
}

inline int nbits(const uint64_t* v) {
    for (int i = 3; i >= 0; --i)
        if (v[i]) return 64 * i + 64 - __builtin_clzll(v[i]);
    return 0;
}

Result solve_one(const std::vector<Op>& ops, WMap& wm) {
    for (uint32_t ip = 0; ip < ops.size(); ++ip) {
        const Op& o = ops[ip];
        int err = E_NONE;
        uint32_t aux = 0;
        if (o.kind == 0) err = solve_arith(wm, o.e);
        else if (o.kind == 1 || o.kind == 2) {
            auto l = wm.find(o.a), r = wm.find(o.b);
            if (l == wm.end()) { err = E_MISSING; aux = o.a; }
            else if (r == wm.end()) { err = E_MISSING; aux = o.b; }
            else {
                Fr x = to_canonical(l->second), y = to_canonical(r->second);
                uint64_t z[4];
                for (int i = 0; i < 4; ++i) {
                    uint64_t m;
                    int lo = 64 * i;
                    if ((int)o.nb >= lo + 64) m = ~0ULL; else if ((int)o.nb <= lo) m = 0; else m = (1ULL << (o.nb - lo)) - 1;
                    uint64_t p = x.l[i] & m, q = y.l[i] & m;
                    z[i] = o.kind == 1 ? (p & q) : (p ^ q);
                }
                while (geq_p(z)) sub_p(z);
                err = insert_value(o.out, from_canonical(z), wm);
            }
        } else if (o.kind == 3) {
            auto l = wm.find(o.a);
            if (l == wm.end()) { err = E_MISSING; aux = o.a; }
            else {
                Fr x = to_canonical(l->second);
                if ((uint32_t)nbits(x.l) > o.nb) err = E_UNSAT;
            }
        }
        if (err != E_NONE) return Result{ST_FAILURE, (uint32_t)err, ip, aux};
    }
    return Result{ST_SOLVED, 0, (uint32_t)ops.size(), 0};
}

std::vector<Op> parse_ops(const uint64_t* s, uint64_t n_words, uint64_t n_ops) {
    std::vector<Op> ops(n_ops);
    uint64_t o = 0;
    for (auto& op : ops) {
        op.kind = (uint32_t)s[o++];
        if (op.kind == 0) {
            uint64_t nm = s[o++], nl = s[o++];
            op.e.mul.resize(nm);
            for (auto& t : op.e.mul) { t.c = from_canonical(s + o); o += 4; t.a = (uint32_t)s[o++]; t.b = (uint32_t)s[o++]; }
            op.e.lin.resize(nl);
            for (auto& t : op.e.lin) { t.c = from_canonical(s + o); o += 4; t.w = (uint32_t)s[o++]; }
            op.e.qc = from_canonical(s + o); o += 4;
        } else if (op.kind == 1 || op.kind == 2) {
            op.a = (uint32_t)s[o++]; op.b = (uint32_t)s[o++]; op.nb = (uint32_t)s[o++]; op.out = (uint32_t)s[o++];
        } else {
            op.a = (uint32_t)s[o++]; op.nb = (uint32_t)s[o++];
        }
    }
    (void)n_words;
    return ops;
}

// ---- "optimised CPU" variant (so the GPU speed-up is not inflated by map / allocation / inversion overhead) ----
// Same circuit, same results, but the way a careful CPU implementation of the BATCH problem would do it: the known-set is
// resolved once per circuit, the inverse of the solved witness' coefficient is hoisted out of the per-instance loop, the
// witness map is a dense vector in Montgomery form, no allocation in the gate loop.  Arithmetic / AND / XOR / RANGE only.
struct OptGate {
    std::vector<MulT> mul;
    std::vector<LinT> lin;
    Fr qc;
    Fr k;            // -1/coeff of the unknown (Montgomery), valid when assign
    uint32_t target;
    bool assign;
};

}  // namespace

extern "C" int ref_solve_batch_optimized(const uint64_t* stream, uint64_t n_words, uint64_t n_ops, const uint32_t* input_ids,
                                         uint32_t n_inputs, const uint64_t* inputs, uint32_t n_inst, uint32_t n_witnesses,
                                         uint64_t* out_witness, uint32_t* results, uint32_t threads) {
    std::vector<Op> ops = parse_ops(stream, n_words, n_ops);
    // plan: static known-set (valid for every instance of circuits without value-dependent gates)
    std::vector<uint8_t> known(n_witnesses, 0);
    for (uint32_t k = 0; k < n_inputs; ++k) known[input_ids[k]] = 1;
    std::vector<OptGate> plan(ops.size());
    for (size_t i = 0; i < ops.size(); ++i) {
        const Op& o = ops[i];
        OptGate& g = plan[i];
        g.assign = false;
        g.target = 0;
        if (o.kind != 0) {
            if (o.kind == 1 || o.kind == 2) known[o.out] = 1;
            continue;
        }
        int unknowns = 0;
        LinT unk{};
        for (const MulT& t : o.e.mul) {
            if (!known[t.a] || !known[t.b]) return -1;   // value-dependent / unsolvable gates: not covered by this variant
            g.mul.push_back(t);
        }
        for (const LinT& t : o.e.lin) {
            if (known[t.w]) g.lin.push_back(t);
            else { unk = t; ++unknowns; }
        }
        if (unknowns > 1) return -1;
        g.qc = o.e.qc;
        if (unknowns == 1) {
            g.assign = true;
            g.target = unk.w;
            g.k = unk.c;                 // inverted below, all gates at once (hoisted out of the per-instance loop)
            known[unk.w] = 1;
        }
    }
    {   // Montgomery's trick: one inversion for the whole circuit
        std::vector<Fr> prefix(plan.size());
        Fr acc = from_canonical(ONE_RAW.l);
        for (size_t i = 0; i < plan.size(); ++i) {
            prefix[i] = acc;
            if (plan[i].assign) acc = mul(acc, plan[i].k);
        }
        Fr inv_acc = inverse(acc);
        for (size_t i = plan.size(); i-- > 0;) {
            if (!plan[i].assign) continue;
            Fr c = plan[i].k;
            plan[i].k = neg(mul(inv_acc, prefix[i]));
            inv_acc = mul(inv_acc, c);
        }
    }
    if (threads == 0) threads = 1;
    auto work = [&](uint32_t t) {
        std::vector<Fr> w(n_witnesses);
        for (uint32_t i = t; i < n_inst; i += threads) {
            for (uint32_t k = 0; k < n_inputs; ++k) w[input_ids[k]] = from_canonical(inputs + ((size_t)i * n_inputs + k) * 4);
            Result r{ST_SOLVED, 0, (uint32_t)ops.size(), 0};
            for (size_t ip = 0; ip < ops.size(); ++ip) {
                const Op& o = ops[ip];
                if (o.kind == 0) {
                    const OptGate& g = plan[ip];
                    Fr acc = g.qc;
                    for (const MulT& m : g.mul) acc = add(acc, mul(mul(m.c, w[m.a]), w[m.b]));
                    for (const LinT& l : g.lin) acc = add(acc, mul(l.c, w[l.w]));
                    if (g.assign) w[g.target] = mul(acc, g.k);
                    else if (acc.l[0] | acc.l[1] | acc.l[2] | acc.l[3]) { r = Result{ST_FAILURE, E_UNSAT, (uint32_t)ip, 0}; break; }
                } else if (o.kind == 1 || o.kind == 2) {
                    Fr x = to_canonical(w[o.a]), y = to_canonical(w[o.b]);
                    uint64_t z[4];
                    for (int q = 0; q < 4; ++q) {
                        uint64_t m;
                        int lo = 64 * q;
                        if ((int)o.nb >= lo + 64) m = ~0ULL; else if ((int)o.nb <= lo) m = 0; else m = (1ULL << (o.nb - lo)) - 1;
                        uint64_t a = x.l[q] & m, b = y.l[q] & m;
                        z[q] = o.kind == 1 ? (a & b) : (a ^ b);
                    }
                    while (geq_p(z)) sub_p(z);
                    w[o.out] = from_canonical(z);
                } else {
                    Fr x = to_canonical(w[o.a]);
                    if ((uint32_t)nbits(x.l) > o.nb) { r = Result{ST_FAILURE, E_UNSAT, (uint32_t)ip, 0}; break; }
                }
            }
            memcpy(results + (size_t)i * 4, &r, 16);
            if (out_witness)
                for (uint32_t k = 0; k < n_witnesses; ++k)
                    if (known[k]) {
                        Fr c = to_canonical(w[k]);
                        memcpy(out_witness + ((size_t)i * n_witnesses + k) * 4, c.l, 32);
                    }
        }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    return 0;
}

namespace {
}  // namespace

extern "C" {
// inputs: [n_inst][n_inputs][4] canonical u64 limbs.  out_witness (optional): [n_inst][n_witnesses][4] canonical,
// out_present (optional): [n_inst][n_witnesses] bytes.  results: [n_inst][4] u32.
int ref_solve_batch(const uint64_t* stream, uint64_t n_words, uint64_t n_ops, const uint32_t* input_ids, uint32_t n_inputs,
                    const uint64_t* inputs, uint32_t n_inst, uint32_t n_witnesses, uint64_t* out_witness, uint8_t* out_present,
                    uint32_t* results, uint32_t threads) {
    std::vector<Op> ops = parse_ops(stream, n_words, n_ops);
    if (threads == 0) threads = 1;
    auto work = [&](uint32_t t) {
        for (uint32_t i = t; i < n_inst; i += threads) {
            WMap wm;
            for (uint32_t k = 0; k < n_inputs; ++k) wm[input_ids[k]] = from_canonical(inputs + ((size_t)i * n_inputs + k) * 4);
            Result r = solve_one(ops, wm);
            memcpy(results + (size_t)i * 4, &r, 16);
            if (out_witness) {
                for (auto& kv : wm) {
                    if (kv.first >= n_witnesses) continue;
                    Fr c = to_canonical(kv.second);
                    memcpy(out_witness + ((size_t)i * n_witnesses + kv.first) * 4, c.l, 32);
                    if (out_present) out_present[(size_t)i * n_witnesses + kv.first] = 1;
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    return 0;
}
}
