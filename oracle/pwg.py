"""Restatement of acvm::pwg (the partial witness generator) on Python ints.

TEST INFRASTRUCTURE ONLY -- the checker for the CUDA path, never the thing measured or shipped.
Follows, statement for statement:
  ACVM struct / solve / solve_opcode      acvm/src/pwg/mod.rs:129-304
  witness_to_value / get_value / insert   acvm/src/pwg/mod.rs:309-357
  ArithmeticSolver                        acvm/src/pwg/arithmetic.rs:27-239
  blackbox dispatch + glue                acvm/src/pwg/blackbox/{mod,logic,range,hash,pedersen,fixed_base_scalar_mul}.rs
  directives                              acvm/src/pwg/directives/mod.rs:23-121 (Quotient, ToLeRadix)
  memory ops                              acvm/src/pwg/memory_op.rs:16-123
"""
from dataclasses import dataclass
from typing import Dict, List, Optional

from . import field as F
from . import grumpkin, hashes
from .acir import Expression, Opcode

# ---- status / error model (acvm/src/pwg/mod.rs:33-127) ---------------------------------------
SOLVED, IN_PROGRESS, FAILURE, REQUIRES_FOREIGN_CALL = "Solved", "InProgress", "Failure", "RequiresForeignCall"


class ReferencePanic(Exception):
    """The reference would panic!() here (API misuse / malformed circuit)."""


@dataclass
class ResolutionError(Exception):
    kind: str  # OpcodeNotSolvable.MissingAssignment | OpcodeNotSolvable.ExpressionHasTooManyUnknowns |
    #            UnsupportedBlackBoxFunc | UnsatisfiedConstrain | IndexOutOfBounds | BlackBoxFunctionFailed |
    #            BrilligFunctionFailed
    opcode_location: Optional[int] = None  # ErrorLocation::Resolved(Acir(ip)) once patched
    witness: Optional[int] = None
    index: Optional[int] = None
    array_size: Optional[int] = None
    func: Optional[str] = None
    message: Optional[str] = None

    def __str__(self):
        return f"{self.kind}@{self.opcode_location} w={self.witness} {self.message or ''}"


# ---- ArithmeticSolver (acvm/src/pwg/arithmetic.rs) -------------------------------------------

def _mul_term_helper(term, wm):  # arithmetic.rs:146-161
    q_m, w_l, w_r = term
    l, r = wm.get(w_l), wm.get(w_r)
    if l is None and r is None:
        return ("TooManyUnknowns",)
    if l is not None and r is not None:
        return ("Solved", F.mul(F.mul(q_m, l), r))
    if l is None:
        return ("OneUnknown", F.mul(q_m, r), w_l)
    return ("OneUnknown", F.mul(q_m, l), w_r)


def evaluate(expr: Expression, wm) -> Expression:  # arithmetic.rs:212-239
    res = Expression()
    for (c, w1, w2) in expr.mul_terms:
        m = _mul_term_helper((c, w1, w2), wm)
        if m[0] == "OneUnknown":
            if m[1] != 0:
                res.linear_combinations.append((m[1], m[2]))
        elif m[0] == "TooManyUnknowns":
            if c != 0:
                res.mul_terms.append((c, w1, w2))
        else:
            res.q_c = F.add(res.q_c, m[1])
    for (c, w) in expr.linear_combinations:
        v = wm.get(w)
        if v is not None:
            res.q_c = F.add(res.q_c, F.mul(c, v))
        elif c != 0:
            res.linear_combinations.append((c, w))
    res.q_c = F.add(res.q_c, expr.q_c)
    return res


def _solve_mul_term(op: Expression, wm):  # arithmetic.rs:133-144
    n = len(op.mul_terms)
    if n == 0:
        return ("Solved", 0)
    if n == 1:
        return _mul_term_helper(op.mul_terms[0], wm)
    raise ReferencePanic("Mul term in the arithmetic opcode must contain either zero or one term")


def _solve_fan_in_term(op: Expression, wm):  # arithmetic.rs:176-209
    unknown = (0, 0)
    num_unknowns = 0
    result = 0
    for term in op.linear_combinations:
        q_l, w_l = term
        v = wm.get(w_l)
        if v is not None:
            result = F.add(result, F.mul(q_l, v))
        else:
            unknown = term
            num_unknowns += 1
        if num_unknowns > 1:
            return ("Unsolvable",)
    if num_unknowns == 0:
        return ("Satisfied", result)
    return ("Solvable", result, unknown)


def insert_value(w, value, wm):  # mod.rs:338-357
    old = wm.get(w)
    wm[w] = value
    if old is not None and old != value:
        raise ResolutionError("UnsatisfiedConstrain")


def witness_to_value(wm, w):  # mod.rs:309-317
    v = wm.get(w)
    if v is None:
        raise ResolutionError("OpcodeNotSolvable.MissingAssignment", witness=w)
    return v


def get_value(expr, wm):  # mod.rs:321-332 + any_witness_from_expression :362-372
    e = evaluate(expr, wm)
    c = e.to_const()
    if c is not None:
        return c
    w = e.linear_combinations[0][1] if e.linear_combinations else e.mul_terms[0][1]
    raise ResolutionError("OpcodeNotSolvable.MissingAssignment", witness=w)


def solve_arithmetic(wm, expr: Expression):  # arithmetic.rs:27-127
    op = evaluate(expr, wm)
    mul_result = _solve_mul_term(op, wm)
    status = _solve_fan_in_term(op, wm)
    too_many = ResolutionError("OpcodeNotSolvable.ExpressionHasTooManyUnknowns")
    unsat = ResolutionError("UnsatisfiedConstrain")
    if mul_result[0] == "TooManyUnknowns" or status[0] == "Unsolvable":
        raise too_many
    if mul_result[0] == "OneUnknown" and status[0] == "Solvable":
        # unreachable after evaluate(); kept for fidelity (arithmetic.rs:43-69)
        q, w1 = mul_result[1], mul_result[2]
        a, (b, w2) = status[1], status[2]
        if w1 != w2:
            raise too_many
        total = F.add(a, op.q_c)
        if F.add(q, b) == 0:
            if total != 0:
                raise unsat
            return
        insert_value(w1, F.div(F.neg(total), F.add(q, b)), wm)
        return
    if mul_result[0] == "OneUnknown" and status[0] == "Satisfied":  # arithmetic.rs:70-91 (unreachable too)
        partial, unk = mul_result[1], mul_result[2]
        total = F.add(status[1], op.q_c)
        if partial == 0:
            if total != 0:
                raise unsat
            return
        insert_value(unk, F.neg(F.div(total, partial)), wm)
        return
    if mul_result[0] == "Solved" and status[0] == "Satisfied":  # arithmetic.rs:92-102
        if F.add(F.add(mul_result[1], status[1]), op.q_c) != 0:
            raise unsat
        return
    # Solved + Solvable  (arithmetic.rs:103-125)
    total = F.add(F.add(mul_result[1], status[1]), op.q_c)
    coeff, unk = status[2]
    if coeff == 0:
        if total != 0:
            raise unsat
        return
    insert_value(unk, F.neg(F.div(total, coeff)), wm)


# ---- blackbox (acvm/src/pwg/blackbox/) -------------------------------------------------------

class StubbedBackend:
    """acvm/tests/solver.rs:20-46 -- every trait method panics."""

    def pedersen(self, inputs, domain_separator):
        raise ReferencePanic("Path not trodden by this test")

    def fixed_base_scalar_mul(self, low, high):
        raise ReferencePanic("Path not trodden by this test")

    def schnorr_verify(self, *a):
        raise ReferencePanic("Path not trodden by this test")


class OracleBackend(StubbedBackend):
    """BlackBoxFunctionSolver (blackbox_solver/src/lib.rs:27-45) backed by oracle/grumpkin + oracle/pedersen."""

    def fixed_base_scalar_mul(self, low, high):
        return grumpkin.fixed_base_scalar_mul(low, high)

    def pedersen(self, inputs, domain_separator):
        from . import pedersen
        return pedersen.commit_native(list(inputs), domain_separator)


def _bb_inputs(bb):  # black_box_function_call.rs:205-292 get_inputs_vec
    n = bb["name"]
    if n in ("AND", "XOR"):
        return [bb["lhs"], bb["rhs"]]
    if n == "RANGE":
        return [bb["input"]]
    if n in ("SHA256", "Blake2s", "Keccak256", "Pedersen", "HashToField128Security"):
        return list(bb["inputs"])
    if n == "Keccak256VariableLength":
        return list(bb["inputs"]) + [bb["var_message_size"]]
    if n == "FixedBaseScalarMul":
        return [bb["low"], bb["high"]]
    if n == "SchnorrVerify":
        return [bb["public_key_x"], bb["public_key_y"]] + list(bb["signature"]) + list(bb["message"])
    if n in ("EcdsaSecp256k1", "EcdsaSecp256r1"):
        return list(bb["public_key_x"]) + list(bb["public_key_y"]) + list(bb["signature"]) + list(bb["hashed_message"])
    if n == "RecursiveAggregation":
        # the input aggregation object is deliberately NOT an input (black_box_function_call.rs:276-279)
        return list(bb["verification_key"]) + list(bb["proof"]) + list(bb["public_inputs"]) + [bb["key_hash"]]
    raise ValueError(n)


def _hash_input(wm, inputs, message_size):  # hash.rs:51-87
    msg = bytearray()
    for (w, nbits) in inputs:
        msg += F.fetch_nearest_bytes(witness_to_value(wm, w), nbits)
    if message_size is not None:
        take = F.to_u128(witness_to_value(wm, message_size[0]))
        if take > len(msg):
            raise ResolutionError("BlackBoxFunctionFailed", func="Keccak256",
                                  message=f"the number of bytes to take from the message is more than the number of bytes in the message. {take} > {len(msg)}")
        msg = msg[:take]
    return bytes(msg)


def _hash256(wm, bb, fn, message_size=None):  # hash.rs:28-48, 89-103
    digest = fn(_hash_input(wm, bb["inputs"], message_size))
    outs = bb["outputs"]
    if len(outs) != 32:
        raise ResolutionError("BlackBoxFunctionFailed", func=bb["name"], message=f"Expected 32 outputs but encountered {len(outs)}")
    for w, byte in zip(outs, digest):
        insert_value(w, byte, wm)


def solve_blackbox(backend, wm, bb):  # blackbox/mod.rs:50-163
    for (w, _) in _bb_inputs(bb):
        if w not in wm:
            raise ResolutionError("OpcodeNotSolvable.MissingAssignment", witness=w)
    n = bb["name"]
    if n in ("AND", "XOR"):  # logic.rs:11-56
        if bb["lhs"][1] != bb["rhs"][1]:
            raise ReferencePanic("number of bits specified for each input must be the same")
        a, b = witness_to_value(wm, bb["lhs"][0]), witness_to_value(wm, bb["rhs"][0])
        fn = F.and_ if n == "AND" else F.xor
        insert_value(bb["output"], fn(a, b, bb["lhs"][1]), wm)
    elif n == "RANGE":  # range.rs:7-18
        if F.num_bits(witness_to_value(wm, bb["input"][0])) > bb["input"][1]:
            raise ResolutionError("UnsatisfiedConstrain")
    elif n == "SHA256":
        _hash256(wm, bb, hashes.sha256)
    elif n == "Blake2s":
        _hash256(wm, bb, hashes.blake2s)
    elif n == "Keccak256":
        _hash256(wm, bb, hashes.keccak256)
    elif n == "Keccak256VariableLength":
        _hash256(wm, bb, hashes.keccak256, bb["var_message_size"])
    elif n == "HashToField128Security":  # hash.rs:13-24, blackbox_solver/src/lib.rs:62-65,94-99
        d = hashes.blake2s(_hash_input(wm, bb["inputs"], None))
        insert_value(bb["output"], F.from_be_bytes_reduce(d), wm)
    elif n == "Pedersen":  # pedersen.rs:11-28
        scalars = [witness_to_value(wm, w) for (w, _) in bb["inputs"]]
        x, y = backend.pedersen(scalars, bb["domain_separator"])
        insert_value(bb["outputs"][0], x, wm)
        insert_value(bb["outputs"][1], y, wm)
    elif n == "FixedBaseScalarMul":  # fixed_base_scalar_mul.rs:11-27
        low, high = witness_to_value(wm, bb["low"][0]), witness_to_value(wm, bb["high"][0])
        try:
            x, y = backend.fixed_base_scalar_mul(low, high)
        except grumpkin.BlackBoxFailed as e:
            raise ResolutionError("BlackBoxFunctionFailed", func=e.func, message=e.reason)
        insert_value(bb["outputs"][0], x, wm)
        insert_value(bb["outputs"][1], y, wm)
    elif n in ("EcdsaSecp256k1", "EcdsaSecp256r1"):  # signature/ecdsa.rs:12-97, signature/mod.rs:5-18
        from . import ecdsa
        low_bytes = lambda ins: bytes(witness_to_value(wm, w) & 0xFF for (w, _) in ins)   # to_be_bytes().last()
        hashed = low_bytes(bb["hashed_message"])
        for key, size, label in (("public_key_x", 32, "pubkey_x"), ("public_key_y", 32, "pubkey_y"), ("signature", 64, "signature")):
            if len(bb[key]) != size:
                raise ResolutionError("BlackBoxFunctionFailed", func=n,
                                      message=f"expected {label} size {size} but received {len(bb[key])}")
        try:
            valid = ecdsa.verify(n, hashed, low_bytes(bb["public_key_x"]), low_bytes(bb["public_key_y"]), low_bytes(bb["signature"]))
        except ecdsa.ReferencePanic as e:
            raise ReferencePanic(str(e))
        insert_value(bb["output"], 1 if valid else 0, wm)
    elif n == "RecursiveAggregation":  # mod.rs:154-161
        for w in bb["output_aggregation_object"]:
            insert_value(w, 0, wm)
    else:
        raise NotImplementedError(f"oracle: blackbox {n} is outside the hot-path scope (SURVEY 8f)")


# ---- directives (acvm/src/pwg/directives/mod.rs) ---------------------------------------------

def _to_radix_le(v, radix):  # num-bigint to_radix_le: zero -> [0]
    if v == 0:
        return [0]
    out = []
    while v:
        out.append(v % radix)
        v //= radix
    return out


def solve_directive(wm, d):
    if d["name"] == "Quotient":  # directives/mod.rs:28-59
        a, b = get_value(d["a"], wm), get_value(d["b"], wm)
        pred = get_value(d["predicate"], wm) if d["predicate"] is not None else 1
        if pred == 0 or b == 0:
            q = r = 0
        else:
            q, r = a // b, a % b
        insert_value(d["q"], q % F.P, wm)
        insert_value(d["r"], r % F.P, wm)
    elif d["name"] == "ToLeRadix":  # directives/mod.rs:60-87
        digits = _to_radix_le(get_value(d["a"], wm), d["radix"])
        if len(d["b"]) < len(digits):
            raise ResolutionError("UnsatisfiedConstrain")
        for i, w in enumerate(d["b"]):
            insert_value(w, (digits[i] & 0xFF) if i < len(digits) else 0, wm)  # from_be_bytes_reduce(&[digit as u8])
    elif d["name"] == "PermutationSort":  # directives/mod.rs:88-121
        from . import sorting
        rows = []
        for element in d["inputs"]:
            if len(element) != d["tuple"]:
                raise ReferencePanic("assert_eq!(element.len(), *tuple as usize)")
            rows.append([get_value(e, wm) for e in element])
        if len(rows) >= 2 and any(i > d["tuple"] for i in d["sort_by"]):
            # the reference panics only if a comparison reaches the index, which depends on slice::sort_by's internals
            raise NotImplementedError("oracle: sort_by index outside the tuple")
        try:
            control = sorting.permutation_sort_bits(rows, d["sort_by"])
        except sorting.ReferencePanic as e:
            raise ReferencePanic(str(e))
        for w, bit in zip(d["bits"], control):
            insert_value(w, 1 if bit else 0, wm)
    else:
        raise ValueError(d["name"])


# ---- memory (acvm/src/pwg/memory_op.rs) ------------------------------------------------------

class MemoryOpSolver:
    def __init__(self):
        self.block_value: Dict[int, int] = {}
        self.block_len = 0

    def _write(self, index, value):  # memory_op.rs:22-36
        if index >= self.block_len:
            raise ResolutionError("IndexOutOfBounds", index=index, array_size=self.block_len)
        self.block_value[index] = value

    def _read(self, index):  # memory_op.rs:38-44
        if index not in self.block_value:
            raise ResolutionError("IndexOutOfBounds", index=index, array_size=self.block_len)
        return self.block_value[index]

    def init(self, init, wm):  # memory_op.rs:47-60
        self.block_len = len(init)
        for i, w in enumerate(init):
            self._write(i, witness_to_value(wm, w))

    def solve_memory_op(self, op, wm):  # memory_op.rs:62-123
        operation = get_value(op["operation"], wm)
        index = get_value(op["index"], wm)
        if F.num_bits(index) > 64:
            raise ReferencePanic("memory index does not fit u64")
        memory_index = index & 0xFFFFFFFF  # `as MemoryIndex` (u32) truncation
        value = evaluate(op["value"], wm)
        is_read = operation == 0
        pred = get_value(op["predicate"], wm) if op["predicate"] is not None else 1
        if is_read:
            w = value.to_witness()
            if w is None:
                raise ReferencePanic("Memory must be read into a specified witness index, encountered an Expression")
            insert_value(w, 0 if pred == 0 else self._read(memory_index), wm)
        elif pred != 0:
            self._write(memory_index, get_value(value, wm))


# ---- the VM (acvm/src/pwg/mod.rs:129-304) ----------------------------------------------------

class ACVM:
    def __init__(self, backend, opcodes: List[Opcode], initial_witness: Dict[int, int]):  # mod.rs:146-156
        self.backend = backend
        self.opcodes = opcodes
        self.witness_map = dict(initial_witness)
        self.block_solvers: Dict[int, MemoryOpSolver] = {}
        self.instruction_pointer = 0
        self.status = SOLVED if not opcodes else IN_PROGRESS
        self.error: Optional[ResolutionError] = None

    def solve(self):  # mod.rs:236-241
        while self.status == IN_PROGRESS:
            self.solve_opcode()
        return self.status

    def solve_opcode(self):  # mod.rs:243-303
        op = self.opcodes[self.instruction_pointer]
        try:
            if op.kind == "Arithmetic":
                solve_arithmetic(self.witness_map, op.body)
            elif op.kind == "BlackBoxFuncCall":
                solve_blackbox(self.backend, self.witness_map, op.body)
            elif op.kind == "Directive":
                solve_directive(self.witness_map, op.body)
            elif op.kind == "MemoryInit":
                self.block_solvers.setdefault(op.body["block_id"], MemoryOpSolver()).init(op.body["init"], self.witness_map)
            elif op.kind == "MemoryOp":
                self.block_solvers.setdefault(op.body["block_id"], MemoryOpSolver()).solve_memory_op(op.body, self.witness_map)
            elif op.kind == "Brillig":
                from . import brillig_vm
                wait = brillig_vm.solve_brillig(self.witness_map, op.body, self.backend, self.instruction_pointer)
                if wait is not None:
                    self.status = REQUIRES_FOREIGN_CALL
                    self.pending_foreign_call = wait
                    return self.status
            else:
                raise ValueError(op.kind)
        except ReferencePanic as e:
            # the reference process would abort here; the oracle reports it as a distinguished failure so that
            # batch tests can compare it with the device's ACVMB_E_REFERENCE_PANIC status
            self.error = ResolutionError("ReferencePanic", message=str(e))
            self.status = FAILURE
            return self.status
        except ResolutionError as e:
            if e.kind in ("UnsatisfiedConstrain", "IndexOutOfBounds"):  # mod.rs:286-296
                e.opcode_location = self.instruction_pointer
            self.error = e
            self.status = FAILURE
            return self.status
        self.instruction_pointer += 1
        self.status = SOLVED if self.instruction_pointer == len(self.opcodes) else IN_PROGRESS
        return self.status

    def resolve_pending_foreign_call(self, result):  # mod.rs:206-228; result = [("Single", v) | ("Array", [v..]), ...]
        if self.status != REQUIRES_FOREIGN_CALL:
            raise ReferencePanic("ACVM is not expecting a foreign call response as no call was made")
        self.opcodes[self.instruction_pointer].body["foreign_call_results"].append(list(result))
        self.status = IN_PROGRESS

    def finalize(self):  # mod.rs:176-181
        if self.status != SOLVED:
            raise ReferencePanic("ACVM is not ready to be finalized")
        return self.witness_map


def solve_circuit(circuit, initial_witness, backend=None):
    """Convenience: returns (status, witness_map, error)."""
    vm = ACVM(backend or OracleBackend(), circuit.opcodes, initial_witness)
    st = vm.solve()
    return st, vm.witness_map, vm.error
