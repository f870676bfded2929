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
//   4 SHA256 / 5 Keccak256: n_in, n_in*(w, num_bits), 32 outputs      (blackbox/hash.rs:28-103)
//   6 FixedBaseScalarMul: low, high, out_x, out_y                     (blackbox/fixed_base_scalar_mul.rs, wasm/scalar_mul.rs:17-65)
//   7 Pedersen: n_in, n_in*w, domain_separator, out_x, out_y          (blackbox/pedersen.rs; PARITY UNPINNED, see oracle/pedersen.py:
//               the lookup-table structure over the generators handed in by ref_set_pedersen_generators)
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
    std::vector<uint32_t> ins, bits, outs;   // hash / Pedersen operands
};
typedef std::map<uint32_t, Fr> WMap;

enum { ST_SOLVED = 0, ST_FAILURE = 2 };
enum { E_NONE = 0, E_MISSING = 1, E_TOO_MANY = 2, E_UNSAT = 4, E_BB_FAILED = 6, E_PANIC = 8 };

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

// ---- SHA-256 (FIPS 180-4) and Keccak-256 (pad 0x01..0x80, rate 136): blackbox_solver/src/lib.rs:47-60 -> sha2 / sha3 crates ----
void sha256(const std::vector<uint8_t>& msg, uint8_t out[32]) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    std::vector<uint8_t> m(msg);
    m.push_back(0x80);
    while (m.size() % 64 != 56) m.push_back(0);
    uint64_t nbit = (uint64_t)msg.size() * 8;
    for (int i = 7; i >= 0; --i) m.push_back((uint8_t)(nbit >> (8 * i)));
    auto rotr = [](uint32_t x, int r) { return (x >> r) | (x << (32 - r)); };
    for (size_t off = 0; off < m.size(); off += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; ++i)
            w[i] = ((uint32_t)m[off + 4 * i] << 24) | ((uint32_t)m[off + 4 * i + 1] << 16) | ((uint32_t)m[off + 4 * i + 2] << 8) | m[off + 4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(h[i >> 2] >> (24 - 8 * (i & 3)));
}

void keccak256(const std::vector<uint8_t>& msg, uint8_t out[32]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
        0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800AULL, 0x800000008000000AULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
    auto rol = [](uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; };
    std::vector<uint8_t> m(msg);
    m.push_back(0x01);
    while (m.size() % 136) m.push_back(0);
    m.back() |= 0x80;
    uint64_t A[5][5] = {{0}};
    for (size_t off = 0; off < m.size(); off += 136) {
        for (int i = 0; i < 17; ++i) {
            uint64_t lane = 0;
            for (int k = 7; k >= 0; --k) lane = (lane << 8) | m[off + 8 * i + k];
            A[i % 5][i / 5] ^= lane;
        }
        for (int rnd = 0; rnd < 24; ++rnd) {
            uint64_t Cc[5], D[5], B[5][5];
            for (int x = 0; x < 5; ++x) Cc[x] = A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4];
            for (int x = 0; x < 5; ++x) D[x] = Cc[(x + 4) % 5] ^ rol(Cc[(x + 1) % 5], 1);
            for (int x = 0; x < 5; ++x)
                for (int y = 0; y < 5; ++y) B[y][(2 * x + 3 * y) % 5] = rol(A[x][y] ^ D[x], ROT[x][y]);
            for (int x = 0; x < 5; ++x)
                for (int y = 0; y < 5; ++y) A[x][y] = B[x][y] ^ (~B[(x + 1) % 5][y] & B[(x + 2) % 5][y]);
            A[0][0] ^= RC[rnd];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 8; ++k) out[8 * i + k] = (uint8_t)(A[i % 5][i / 5] >> (8 * k));
}

// ---- Grumpkin y^2 = x^3 - 17 over Fr (Montgomery form), Jacobian coordinates ----
struct Jac { Fr X, Y, Z; bool inf; };
const Fr FR_ZERO = {{0, 0, 0, 0}};
inline Fr subf(const Fr& a, const Fr& b) { return add(a, neg(b)); }
inline bool raw_zero(const Fr& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
inline bool raw_eq(const Fr& a, const Fr& b) { return memcmp(a.l, b.l, 32) == 0; }
Jac jac_dbl(const Jac& p) {
    if (p.inf || raw_zero(p.Y)) return Jac{FR_ZERO, FR_ZERO, FR_ZERO, true};
    Fr A = mul(p.X, p.X), B = mul(p.Y, p.Y), Cc = mul(B, B);
    Fr t = add(p.X, B);
    Fr D = subf(subf(mul(t, t), A), Cc);
    D = add(D, D);
    Fr E = add(add(A, A), A), F = mul(E, E);
    Jac r;
    r.X = subf(F, add(D, D));
    Fr c8 = add(Cc, Cc); c8 = add(c8, c8); c8 = add(c8, c8);
    r.Y = subf(mul(E, subf(D, r.X)), c8);
    Fr yz = mul(p.Y, p.Z);
    r.Z = add(yz, yz);
    r.inf = false;
    return r;
}
Jac jac_add_affine(const Jac& p, const Fr& qx, const Fr& qy) {   // mixed addition, q finite
    if (p.inf) return Jac{qx, qy, from_canonical(ONE_RAW.l), false};
    Fr Z2 = mul(p.Z, p.Z), U2 = mul(qx, Z2), S2 = mul(qy, mul(Z2, p.Z));
    Fr H = subf(U2, p.X), Rr = subf(S2, p.Y);
    if (raw_zero(H)) {
        if (raw_zero(Rr)) return jac_dbl(p);
        return Jac{FR_ZERO, FR_ZERO, FR_ZERO, true};
    }
    Fr H2 = mul(H, H), H3 = mul(H2, H), V = mul(p.X, H2);
    Jac r;
    r.X = subf(subf(mul(Rr, Rr), H3), add(V, V));
    r.Y = subf(mul(Rr, subf(V, r.X)), mul(p.Y, H3));
    r.Z = mul(p.Z, H);
    r.inf = false;
    return r;
}
void jac_to_affine(const Jac& p, Fr& x, Fr& y) {   // (0, 0) for the point at infinity (unpinned in the reference)
    if (p.inf) { x = y = FR_ZERO; return; }
    Fr zi = inverse(p.Z), zi2 = mul(zi, zi);
    x = mul(p.X, zi2);
    y = mul(p.Y, mul(zi2, zi));
}

// Pedersen lookup tables: 58 tables of (i+1)*G_k (29 per parity; 512 entries, the last of each half 4) + 1024 IV points.
// Built once from the generators the Python oracle derives (oracle/pedersen.py): affine, Montgomery form.
struct Aff { Fr x, y; };
std::vector<std::vector<Aff>> g_ped_tables;
std::vector<Aff> g_ped_iv;
std::vector<Aff> build_table(const Aff& g, size_t n) {
    std::vector<Aff> t(n);
    Jac acc{FR_ZERO, FR_ZERO, FR_ZERO, true};
    for (size_t i = 0; i < n; ++i) {
        acc = jac_add_affine(acc, g.x, g.y);
        jac_to_affine(acc, t[i].x, t[i].y);
    }
    return t;
}
Jac ped_hash_single(const Fr& v_mont, int parity) {
    Fr v = to_canonical(v_mont);
    Jac acc{FR_ZERO, FR_ZERO, FR_ZERO, true};
    for (int i = 0; i < 29; ++i) {
        int bit = 9 * i;
        uint64_t s = v.l[bit / 64] >> (bit % 64);
        if (bit % 64 > 55 && bit / 64 < 3) s |= v.l[bit / 64 + 1] << (64 - bit % 64);
        s &= 511;
        const Aff& pt = g_ped_tables[29 * parity + i][s];
        acc = jac_add_affine(acc, pt.x, pt.y);
    }
    return acc;
}
Jac jac_add(const Jac& a, const Jac& b) {   // general addition via the affine form of b (one extra inversion; CPU baseline only)
    if (b.inf) return a;
    Fr bx, by;
    jac_to_affine(b, bx, by);
    return jac_add_affine(a, bx, by);
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
        } else {
            // blackbox/mod.rs:55-62: every input assigned, else MissingAssignment(first missing)
            for (uint32_t w : o.ins)
                if (err == E_NONE && wm.find(w) == wm.end()) { err = E_MISSING; aux = w; }
            if (err == E_NONE && (o.kind == 4 || o.kind == 5)) {
                std::vector<uint8_t> msg;   // low ceil(num_bits/8) bytes of every input, little-endian (generic_ark.rs:305-317)
                for (size_t k = 0; k < o.ins.size(); ++k) {
                    Fr v = to_canonical(wm.find(o.ins[k])->second);
                    uint32_t nbytes = (o.bits[k] + 7) / 8;
                    for (uint32_t j = 0; j < nbytes && j < 32; ++j) msg.push_back((uint8_t)(v.l[j / 8] >> (8 * (j % 8))));
                }
                uint8_t dg[32];
                if (o.kind == 4) sha256(msg, dg); else keccak256(msg, dg);
                for (int i = 0; i < 32 && err == E_NONE; ++i) {
                    uint64_t b[4] = {dg[i], 0, 0, 0};
                    err = insert_value(o.outs[i], from_canonical(b), wm);
                }
            } else if (err == E_NONE && o.kind == 6) {
                Fr lo = to_canonical(wm.find(o.ins[0])->second), hi = to_canonical(wm.find(o.ins[1])->second);
                static const uint64_t ORDER[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
                uint64_t sc[4] = {lo.l[0], lo.l[1], hi.l[0], hi.l[1]};
                if (lo.l[2] | lo.l[3] | hi.l[2] | hi.l[3]) { err = E_BB_FAILED; aux = 10; }
                else if (cmp(sc, ORDER) >= 0) { err = E_BB_FAILED; aux = 10; }
                else {
                    // generic double-and-add s*G (what the wasm's compute_public_key does internally is not in the reference)
                    static const uint64_t GY[4] = {0x833fc48d823f272cULL, 0x2d270d45f1181294ULL, 0xcf135e7506a45d63ULL, 0x0000000000000002ULL};
                    Fr gx = from_canonical(ONE_RAW.l), gy = from_canonical(GY);
                    Jac acc{FR_ZERO, FR_ZERO, FR_ZERO, true};
                    for (int bit = 255; bit >= 0; --bit) {
                        acc = jac_dbl(acc);
                        if ((sc[bit / 64] >> (bit % 64)) & 1) acc = jac_add_affine(acc, gx, gy);
                    }
                    Fr x, y;
                    jac_to_affine(acc, x, y);
                    err = insert_value(o.outs[0], x, wm);
                    if (err == E_NONE) err = insert_value(o.outs[1], y, wm);
                }
            } else if (err == E_NONE && o.kind == 7) {
                Fr x = FR_ZERO, y = FR_ZERO;
                if (!o.ins.empty()) {
                    if (g_ped_tables.empty()) { err = E_PANIC; }
                    else {
                        Fr r = g_ped_iv[o.nb % g_ped_iv.size()].x;
                        for (uint32_t w : o.ins) {   // hash_pair(r, v) = (H0(r) + H1(v)).x
                            Jac s = jac_add(ped_hash_single(r, 0), ped_hash_single(wm.find(w)->second, 1));
                            Fr sy;
                            jac_to_affine(s, r, sy);
                        }
                        uint64_t n[4] = {o.ins.size(), 0, 0, 0};
                        Jac s = jac_add(ped_hash_single(r, 0), ped_hash_single(from_canonical(n), 1));
                        jac_to_affine(s, x, y);
                    }
                }
                if (err == E_NONE) err = insert_value(o.outs[0], x, wm);
                if (err == E_NONE) err = insert_value(o.outs[1], y, wm);
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
        } else if (op.kind == 3) {
            op.a = (uint32_t)s[o++]; op.nb = (uint32_t)s[o++];
        } else if (op.kind == 4 || op.kind == 5) {
            uint64_t n = s[o++];
            for (uint64_t k = 0; k < n; ++k) { op.ins.push_back((uint32_t)s[o++]); op.bits.push_back((uint32_t)s[o++]); }
            for (int k = 0; k < 32; ++k) op.outs.push_back((uint32_t)s[o++]);
        } else if (op.kind == 6) {
            op.ins = {(uint32_t)s[o], (uint32_t)s[o + 1]};
            op.outs = {(uint32_t)s[o + 2], (uint32_t)s[o + 3]};
            o += 4;
        } else {
            uint64_t n = s[o++];
            for (uint64_t k = 0; k < n; ++k) op.ins.push_back((uint32_t)s[o++]);
            op.nb = (uint32_t)s[o++];
            op.outs = {(uint32_t)s[o], (uint32_t)s[o + 1]};
            o += 2;
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
// gens: 58 generators then the IV generator, each (x[4], y[4]) canonical limbs (oracle/pedersen.py derives them)
int ref_set_pedersen_generators(const uint64_t* gens) {
    g_ped_tables.clear();
    for (int k = 0; k < 58; ++k) {
        Aff g{from_canonical(gens + 8 * k), from_canonical(gens + 8 * k + 4)};
        g_ped_tables.push_back(build_table(g, (k % 29) == 28 ? 4 : 512));
    }
    Aff iv{from_canonical(gens + 8 * 58), from_canonical(gens + 8 * 58 + 4)};
    g_ped_iv = build_table(iv, 1024);
    return 0;
}

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
