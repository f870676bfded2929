"""Writer for the reference ACIR wire format + the synthetic benchmark circuits.

`Circuit::write` = gzip(bincode(circuit)) (acir/src/circuit/mod.rs:145-152); field elements are
serialised as 64-char lowercase hex strings (acir_field/src/generic_ark.rs:114-121).  The solver
itself only READS this format (C++ decoder in csrc/acir.cpp); the writer exists so tests, the
bench and users can hand the library the same bytes the reference would consume.

The synthetic generator follows SURVEY.md 8(d) / BASELINE.md 4: splitmix64 stream, width-3 PLONK
gates  q_M*w_a*w_b + q_l*w_a + q_r*w_b + q_o*w_new + q_c = 0, operands from the previous 64
witnesses ("local") or from all earlier witnesses ("global"), every 16th opcode re-emits an earlier
gate as an all-known check.
"""
import gzip
import struct

P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
MASK64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def field(self):
        v = self.next() | (self.next() << 64) | (self.next() << 128) | (self.next() << 192)
        return v % P

    def nonzero_field(self):
        while True:
            v = self.field()
            if v:
                return v

    def below(self, n):
        return self.next() % n


class Writer:
    def __init__(self):
        self.parts = []

    def u8(self, v):
        self.parts.append(struct.pack("<B", v))

    def u32(self, v):
        self.parts.append(struct.pack("<I", v))

    def u64(self, v):
        self.parts.append(struct.pack("<Q", v))

    def fe(self, v):
        self.parts.append(_FE_PREFIX)
        self.parts.append(b"%064x" % (v % P))

    def string(self, s):
        b = s.encode()
        self.u64(len(b))
        self.parts.append(b)

    def bytes(self):
        return b"".join(self.parts)


_FE_PREFIX = struct.pack("<Q", 64)


def w_expression(w, mul_terms, lin, q_c):
    w.u64(len(mul_terms))
    for (c, a, b) in mul_terms:
        w.fe(c)
        w.u32(a)
        w.u32(b)
    w.u64(len(lin))
    for (c, x) in lin:
        w.fe(c)
        w.u32(x)
    w.fe(q_c)


BLACKBOX_TAGS = {"AND": 0, "XOR": 1, "RANGE": 2, "SHA256": 3, "Blake2s": 4, "SchnorrVerify": 5, "Pedersen": 6,
                 "HashToField128Security": 7, "EcdsaSecp256k1": 8, "EcdsaSecp256r1": 9, "FixedBaseScalarMul": 10,
                 "Keccak256": 11, "Keccak256VariableLength": 12, "RecursiveAggregation": 13}


class CircuitBuilder:
    """Append opcodes, then `to_bytes()` gives gzip(bincode(Circuit))."""

    def __init__(self):
        self.w = Writer()
        self.n_opcodes = 0
        self.max_witness = 0
        self.private_parameters = []
        self.public_parameters = []
        self.return_values = []

    def _see(self, *ws):
        for x in ws:
            if x > self.max_witness:
                self.max_witness = x

    def arithmetic(self, mul_terms, lin, q_c=0):
        self.w.u32(0)
        w_expression(self.w, mul_terms, lin, q_c)
        self.n_opcodes += 1
        for (_, a, b) in mul_terms:
            self._see(a, b)
        for (_, x) in lin:
            self._see(x)

    def _fi(self, fi):
        self.w.u32(fi[0])
        self.w.u32(fi[1])
        self._see(fi[0])

    def _vfi(self, v):
        self.w.u64(len(v))
        for fi in v:
            self._fi(fi)

    def _vw(self, v):
        self.w.u64(len(v))
        for x in v:
            self.w.u32(x)
            self._see(x)

    def logic(self, name, lhs, rhs, output):
        self.w.u32(1)
        self.w.u32(BLACKBOX_TAGS[name])
        self._fi(lhs)
        self._fi(rhs)
        self.w.u32(output)
        self._see(output)
        self.n_opcodes += 1

    def range(self, inp):
        self.w.u32(1)
        self.w.u32(BLACKBOX_TAGS["RANGE"])
        self._fi(inp)
        self.n_opcodes += 1

    def hash256(self, name, inputs, outputs):
        self.w.u32(1)
        self.w.u32(BLACKBOX_TAGS[name])
        self._vfi(inputs)
        self._vw(outputs)
        self.n_opcodes += 1

    def hash_to_field(self, inputs, output):
        self.w.u32(1)
        self.w.u32(BLACKBOX_TAGS["HashToField128Security"])
        self._vfi(inputs)
        self.w.u32(output)
        self._see(output)
        self.n_opcodes += 1

    def keccak_var(self, inputs, var_message_size, outputs):
        self.w.u32(1)
        self.w.u32(BLACKBOX_TAGS["Keccak256VariableLength"])
        self._vfi(inputs)
        self._fi(var_message_size)
        self._vw(outputs)
        self.n_opcodes += 1

    def pedersen(self, inputs, domain_separator, outputs):
        self.w.u32(1)
        self.w.u32(BLACKBOX_TAGS["Pedersen"])
        self._vfi(inputs)
        self.w.u32(domain_separator)
        self.w.u32(outputs[0])
        self.w.u32(outputs[1])
        self._see(*outputs)
        self.n_opcodes += 1

    def fixed_base_scalar_mul(self, low, high, outputs):
        self.w.u32(1)
        self.w.u32(BLACKBOX_TAGS["FixedBaseScalarMul"])
        self._fi(low)
        self._fi(high)
        self.w.u32(outputs[0])
        self.w.u32(outputs[1])
        self._see(*outputs)
        self.n_opcodes += 1

    def ecdsa(self, name, public_key_x, public_key_y, signature, hashed_message, output):
        """name: EcdsaSecp256k1 | EcdsaSecp256r1; field order of black_box_function_call.rs:60-75"""
        self.w.u32(1)
        self.w.u32(BLACKBOX_TAGS[name])
        self._vfi(public_key_x)
        self._vfi(public_key_y)
        self._vfi(signature)
        self._vfi(hashed_message)
        self.w.u32(output)
        self._see(output)
        self.n_opcodes += 1

    def recursive_aggregation(self, verification_key, proof, public_inputs, key_hash, input_aggregation_object,
                              output_aggregation_object):
        """black_box_function_call.rs:84-112; input_aggregation_object is an Option<Vec<FunctionInput>>"""
        self.w.u32(1)
        self.w.u32(BLACKBOX_TAGS["RecursiveAggregation"])
        self._vfi(verification_key)
        self._vfi(proof)
        self._vfi(public_inputs)
        self._fi(key_hash)
        if input_aggregation_object is None:
            self.w.u8(0)
        else:
            self.w.u8(1)
            self._vfi(input_aggregation_object)
        self._vw(output_aggregation_object)
        self.n_opcodes += 1

    def _expr(self, e):
        """e = (mul_terms, lin, q_c)"""
        w_expression(self.w, e[0], e[1], e[2])
        for (_, a, b) in e[0]:
            self._see(a, b)
        for (_, x) in e[1]:
            self._see(x)

    def _opt_expr(self, e):
        if e is None:
            self.w.u8(0)
        else:
            self.w.u8(1)
            self._expr(e)

    def directive_quotient(self, a, b, q, r, predicate=None):
        """Directive::Quotient (acir/src/circuit/directives.rs:5-11); a, b, predicate are (mul, lin, q_c) expressions."""
        self.w.u32(2)
        self.w.u32(0)
        self._expr(a)
        self._expr(b)
        self.w.u32(q)
        self.w.u32(r)
        self._see(q, r)
        self._opt_expr(predicate)
        self.n_opcodes += 1

    def directive_to_le_radix(self, a, b, radix):
        self.w.u32(2)
        self.w.u32(1)
        self._expr(a)
        self._vw(b)
        self.w.u32(radix)
        self.n_opcodes += 1

    def directive_permutation_sort(self, inputs, tuple_, bits, sort_by):
        """Directive::PermutationSort (acir/src/circuit/directives.rs:24-35); inputs: list of lists of expressions."""
        self.w.u32(2)
        self.w.u32(2)
        self.w.u64(len(inputs))
        for element in inputs:
            self.w.u64(len(element))
            for e in element:
                self._expr(e)
        self.w.u32(tuple_)
        self._vw(bits)
        self.w.u64(len(sort_by))
        for i in sort_by:
            self.w.u32(i)
        self.n_opcodes += 1

    def memory_init(self, block_id, init):
        self.w.u32(5)
        self.w.u32(block_id)
        self._vw(init)
        self.n_opcodes += 1

    def memory_op(self, block_id, operation, index, value, predicate=None):
        """Opcode::MemoryOp (acir/src/circuit/opcodes.rs:24-31); operation/index/value/predicate are expressions."""
        self.w.u32(4)
        self.w.u32(block_id)
        self._expr(operation)
        self._expr(index)
        self._expr(value)
        self._opt_expr(predicate)
        self.n_opcodes += 1

    # ---- Brillig (acir/src/circuit/brillig.rs:9-33, brillig/src/opcodes.rs:60-134) ----
    _BRILLIG_TAGS = {"BinaryFieldOp": 0, "BinaryIntOp": 1, "JumpIfNot": 2, "JumpIf": 3, "Jump": 4, "Call": 5, "Const": 6, "Return": 7,
                     "ForeignCall": 8, "Mov": 9, "Load": 10, "Store": 11, "BlackBox": 12, "Trap": 13, "Stop": 14}
    _BRILLIG_BB_TAGS = {"Sha256": 0, "Blake2s": 1, "Keccak256": 2, "HashToField128Security": 3, "EcdsaSecp256k1": 4,
                        "EcdsaSecp256r1": 5, "SchnorrVerify": 6, "Pedersen": 7, "FixedBaseScalarMul": 8}

    def _reg_or_mem(self, d):
        kind = {"Register": 0, "HeapArray": 1, "HeapVector": 2}[d[0]]
        self.w.u32(kind)
        self.w.u64(d[1])
        if kind:
            self.w.u64(d[2])

    def _brillig_op(self, o):
        w = self.w
        op = o["op"]
        w.u32(self._BRILLIG_TAGS[op])
        if op == "BinaryFieldOp":
            w.u64(o["destination"]); w.u32(o["bop"]); w.u64(o["lhs"]); w.u64(o["rhs"])
        elif op == "BinaryIntOp":
            w.u64(o["destination"]); w.u32(o["bop"]); w.u32(o["bit_size"]); w.u64(o["lhs"]); w.u64(o["rhs"])
        elif op in ("JumpIfNot", "JumpIf"):
            w.u64(o["condition"]); w.u64(o["location"])
        elif op in ("Jump", "Call"):
            w.u64(o["location"])
        elif op == "Const":
            w.u64(o["destination"]); w.fe(o["value"])
        elif op == "ForeignCall":
            w.string(o["function"])
            w.u64(len(o["destinations"]))
            for d in o["destinations"]:
                self._reg_or_mem(d)
            w.u64(len(o["inputs"]))
            for d in o["inputs"]:
                self._reg_or_mem(d)
        elif op == "Mov":
            w.u64(o["destination"]); w.u64(o["source"])
        elif op == "Load":
            w.u64(o["destination"]); w.u64(o["source_pointer"])
        elif op == "Store":
            w.u64(o["destination_pointer"]); w.u64(o["source"])
        elif op == "BlackBox":
            bb = o["bb"]
            w.u32(self._BRILLIG_BB_TAGS[bb["name"]])
            if bb["name"] in ("Sha256", "Blake2s", "Keccak256"):
                for v in (*bb["message"], *bb["output"]):
                    w.u64(v)
            elif bb["name"] == "HashToField128Security":
                for v in (*bb["message"], bb["output"]):
                    w.u64(v)
            elif bb["name"] in ("EcdsaSecp256k1", "EcdsaSecp256r1"):
                for v in (*bb["hashed_msg"], *bb["public_key_x"], *bb["public_key_y"], *bb["signature"], bb["result"]):
                    w.u64(v)
            elif bb["name"] == "FixedBaseScalarMul":
                for v in (bb["low"], bb["high"], *bb["result"]):
                    w.u64(v)
            else:
                raise NotImplementedError(bb["name"])

    def brillig(self, inputs, outputs, bytecode, foreign_call_results=(), predicate=None):
        """inputs: [("Single", expr) | ("Array", [expr..])], outputs: [("Simple", w) | ("Array", [w..])],
        bytecode: list of dicts shaped like oracle/acir.py decodes them, foreign_call_results: [[("Single", v)|("Array",[v..])..]..]"""
        w = self.w
        w.u32(3)
        w.u64(len(inputs))
        for kind, e in inputs:
            if kind == "Single":
                w.u32(0)
                self._expr(e)
            else:
                w.u32(1)
                w.u64(len(e))
                for x in e:
                    self._expr(x)
        w.u64(len(outputs))
        for kind, o in outputs:
            if kind == "Simple":
                w.u32(0)
                w.u32(o)
                self._see(o)
            else:
                w.u32(1)
                self._vw(o)
        w.u64(len(foreign_call_results))
        for res in foreign_call_results:
            w.u64(len(res))
            for kind, v in res:
                if kind == "Single":
                    w.u32(0)
                    w.fe(v)
                else:
                    w.u32(1)
                    w.u64(len(v))
                    for x in v:
                        w.fe(x)
        w.u64(len(bytecode))
        for o in bytecode:
            self._brillig_op(o)
        self._opt_expr(predicate)
        self.n_opcodes += 1

    def to_raw(self, current_witness_index=None):
        head = Writer()
        head.u32(self.max_witness + 1 if current_witness_index is None else current_witness_index)
        head.u64(self.n_opcodes)
        tail = Writer()
        for s in (self.private_parameters, self.public_parameters, self.return_values):
            tail.u64(len(s))
            for x in sorted(set(s)):
                tail.u32(x)
        tail.u64(0)  # assert_messages
        return head.bytes() + self.w.bytes() + tail.bytes()

    def to_bytes(self, current_witness_index=None, level=1):
        return gzip.compress(self.to_raw(current_witness_index), compresslevel=level, mtime=0)


def wexpr(w):
    """the expression `1*w`"""
    return ([], [(1, w)], 0)


def cexpr(c):
    """a constant expression"""
    return ([], [], c % P)


SEED_BASE = 0xAC1DB20000000000
N_INPUTS = 8


def synthetic_arith_circuit(n_gates, seed_id=1, mode="local", coeffs="dense", window=64, check_every=16):
    """SURVEY 8(d) generator.  Returns (acir_bytes, input_witnesses, n_witnesses).

    Witnesses 0..7 are the per-instance inputs; opcode i either defines a new witness from two
    earlier ones or (every `check_every`-th opcode) re-emits an earlier gate verbatim."""
    rng = SplitMix64(SEED_BASE + seed_id)
    b = CircuitBuilder()
    gates = []
    n_defined = N_INPUTS
    pow2 = [pow(2, j, P) for j in range(0, 64)]
    for i in range(n_gates):
        if check_every and i % check_every == check_every - 1 and gates:
            g = gates[rng.below(len(gates))]
            b.arithmetic(*g)
            continue
        lo = max(0, n_defined - window) if mode == "local" else 0
        a = lo + rng.below(n_defined - lo)
        c = lo + rng.below(n_defined - lo)
        out = n_defined
        if coeffs == "dense":
            qm, ql, qr, qo, qc = (rng.nonzero_field() for _ in range(5))
        else:  # "noir-like": q_o = -1, q_M in {0,1}, q_l,q_r in {0,+-1,2^j}
            def small():
                k = rng.below(4)
                if k == 0:
                    return 0
                if k == 1:
                    return 1
                if k == 2:
                    return P - 1
                return pow2[rng.below(64)]
            qm, ql, qr, qo, qc = rng.below(2), small(), small(), P - 1, (rng.below(2) and small())
        mul = [(qm, a, c)] if qm else []
        lin = [(q, w) for (q, w) in ((ql, a), (qr, c), (qo, out)) if q]
        g = (mul, lin, qc)
        b.arithmetic(*g)
        gates.append(g)
        n_defined += 1
    b.private_parameters = list(range(N_INPUTS))
    return b.to_bytes(current_witness_index=n_defined - 1 if n_defined else 0), list(range(N_INPUTS)), n_defined


def synthetic_inputs(batch, n_inputs=N_INPUTS, seed_id=1, first_instance=0):
    """Per-instance inputs, uniform Fr, as [batch][n_inputs][32] big-endian bytes (instance-indexed stream)."""
    out = bytearray(batch * n_inputs * 32)
    for i in range(batch):
        rng = SplitMix64((SEED_BASE + seed_id) ^ (0x9E3779B97F4A7C15 * (first_instance + i + 1) & MASK64))
        for k in range(n_inputs):
            out[(i * n_inputs + k) * 32:(i * n_inputs + k + 1) * 32] = rng.field().to_bytes(32, "big")
    return bytes(out)


# ---- BASELINE.json configs 2-4 as concrete synthetic circuits ------------------------------------
def pedersen_chain_circuit(n_calls, n_fresh=64):
    """Config 2: `n_calls` chained Pedersen{[prev.x, fresh_i], domain_separator 0}; fresh_i cycles over `n_fresh` inputs."""
    b = CircuitBuilder()
    inputs = list(range(1, n_fresh + 2))        # w1 = chain seed, w2.. = fresh values
    prev, nxt = 1, n_fresh + 2
    for i in range(n_calls):
        b.pedersen([(prev, 254), (2 + (i % n_fresh), 254)], 0, (nxt, nxt + 1))
        prev, nxt = nxt, nxt + 2
    b.private_parameters = inputs
    return b.to_bytes(), inputs, nxt


def hash_chain_circuit(n_calls, n_fresh=32):
    """Config 3: `n_calls` hash calls alternating SHA256 / Keccak256 over 64 byte-witnesses = previous 32-byte digest ||
    32 fresh bytes (the fresh bytes cycle over `n_fresh` byte-valued inputs instead of 32 new inputs per call)."""
    b = CircuitBuilder()
    inputs = list(range(1, 32 + n_fresh + 1))   # w1..w32 = initial "digest", then the fresh byte pool
    prev = list(range(1, 33))
    pool = list(range(33, 33 + n_fresh))
    nxt = 33 + n_fresh
    for i in range(n_calls):
        fresh = [pool[(i * 7 + k) % n_fresh] for k in range(32)]
        outs = list(range(nxt, nxt + 32))
        b.hash256("SHA256" if i % 2 == 0 else "Keccak256", [(w, 8) for w in prev + fresh], outs)
        prev, nxt = outs, nxt + 32
    b.private_parameters = inputs
    return b.to_bytes(), inputs, nxt


def mixed_circuit(n_ops, seed_id=4, window=64):
    """Config 4: 93% dense arithmetic gates, 4% RANGE/AND/XOR (32-bit), 2% SHA256/Keccak256 over 64 byte-witnesses,
    1% Pedersen(2)/FixedBaseScalarMul.  Returns (bytes, inputs, n_witnesses, counts)."""
    rng = SplitMix64(SEED_BASE + seed_id)
    b = CircuitBuilder()
    n_in = N_INPUTS
    fields = list(range(n_in))          # witnesses holding arbitrary field values (most recent last)
    words = []                          # witnesses known to be < 2^32
    counts = dict(arith=0, logic=0, range=0, hash=0, pedersen=0, fixed_base=0)
    nxt = n_in

    def pick(pool):
        lo = max(0, len(pool) - window)
        return pool[lo + rng.below(len(pool) - lo)]

    for _ in range(n_ops):
        k = rng.below(100)
        if k < 93 or (k < 97 and len(fields) < 2):
            a, c = pick(fields), pick(fields)
            qm, ql, qr, qo, qc = (rng.nonzero_field() for _ in range(5))
            b.arithmetic([(qm, a, c)], [(ql, a), (qr, c), (qo, nxt)], qc)
            fields.append(nxt)
            nxt += 1
            counts["arith"] += 1
        elif k < 97:
            a, c = pick(fields), pick(fields)
            sel = rng.below(3)
            if sel == 2 and words:
                b.range((pick(words), 32))
                counts["range"] += 1
            else:
                b.logic("AND" if sel == 0 else "XOR", (a, 32), (c, 32), nxt)
                words.append(nxt)
                fields.append(nxt)
                nxt += 1
                counts["logic"] += 1
        elif k < 99:
            ins = [(pick(fields), 8) for _ in range(64)]
            outs = list(range(nxt, nxt + 32))
            b.hash256("SHA256" if rng.below(2) else "Keccak256", ins, outs)
            words += outs
            fields += outs
            nxt += 32
            counts["hash"] += 1
        else:
            if rng.below(2) or len(words) < 2:
                b.pedersen([(pick(fields), 254), (pick(fields), 254)], 0, (nxt, nxt + 1))
                counts["pedersen"] += 1
            else:
                b.fixed_base_scalar_mul((pick(words), 128), (pick(words), 128), (nxt, nxt + 1))
                counts["fixed_base"] += 1
            fields += [nxt, nxt + 1]
            nxt += 2
    b.private_parameters = list(range(n_in))
    return b.to_bytes(), list(range(n_in)), nxt, counts
