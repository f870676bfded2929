"""Brillig VM + the pwg glue around it, restated on Python ints.  TEST INFRASTRUCTURE ONLY.

Follows  brillig_vm/src/lib.rs:86-389 (VM), registers.rs:13-42, memory.rs:17-48, arithmetic.rs:7-96, black_box.rs:42-165
and      acvm/src/pwg/brillig.rs:20-150 (BrilligSolver::solve, zero_out_brillig_outputs).
"""
from . import field as F
from . import grumpkin, hashes
from .pwg import ReferencePanic, ResolutionError, get_value, insert_value

MAX_REGISTERS = 1 << 16


class VM:
    def __init__(self, registers, memory, bytecode, foreign_call_results, backend):
        self.registers = list(registers)
        self.memory = list(memory)
        self.bytecode = bytecode
        self.foreign_call_results = foreign_call_results
        self.backend = backend
        self.pc = 0
        self.fc_counter = 0
        self.call_stack = []
        self.status = ("InProgress",)

    # registers.rs
    def get(self, r):
        if r >= MAX_REGISTERS:
            raise ReferencePanic("Reading register past maximum!")
        return self.registers[r] if r < len(self.registers) else 0

    def set(self, r, v):
        if r >= MAX_REGISTERS:
            raise ReferencePanic("Writing register past maximum!")
        if len(self.registers) <= r:
            self.registers += [0] * (r + 1 - len(self.registers))
        self.registers[r] = v

    @staticmethod
    def to_usize(v):  # value.rs:39-42
        if v.bit_length() > 64:
            raise ReferencePanic("register does not fit into u64")
        return v

    def mread(self, p):
        if p >= len(self.memory):
            raise ReferencePanic("memory index out of range")
        return self.memory[p]

    def mwrite(self, p, vals):
        if len(self.memory) < p + len(vals):
            self.memory += [0] * (p + len(vals) - len(self.memory))
        self.memory[p:p + len(vals)] = vals

    def _fail(self, msg):
        self.status = ("Failure", msg, list(self.call_stack) + [self.pc])
        return self.status

    def _set_pc(self, v):
        assert self.pc < len(self.bytecode)
        self.pc = v
        if self.pc >= len(self.bytecode):
            self.status = ("Finished",)
        return self.status

    def process_opcodes(self):
        while True:
            st = self.process_opcode()
            if st[0] in ("Finished", "Failure", "ForeignCallWait"):
                return self.status

    def process_opcode(self):
        if self.pc >= len(self.bytecode):
            raise ReferencePanic("index out of bounds: bytecode")
        o = self.bytecode[self.pc]
        op = o["op"]
        if op == "BinaryFieldOp":
            a, b = self.get(o["lhs"]), self.get(o["rhs"])
            r = [F.add, F.sub, F.mul, F.div, lambda x, y: int(x == y)][o["bop"]](a, b)
            self.set(o["destination"], r)
            return self._set_pc(self.pc + 1)
        if op == "BinaryIntOp":
            r = bigint_op(o["bop"], self.get(o["lhs"]), self.get(o["rhs"]), o["bit_size"])
            self.set(o["destination"], r % F.P)
            return self._set_pc(self.pc + 1)
        if op == "Jump":
            return self._set_pc(o["location"])
        if op == "JumpIf":
            return self._set_pc(o["location"] if self.get(o["condition"]) != 0 else self.pc + 1)
        if op == "JumpIfNot":
            return self._set_pc(o["location"] if self.get(o["condition"]) == 0 else self.pc + 1)
        if op == "Return":
            if self.call_stack:
                return self._set_pc(self.call_stack.pop() + 1)
            return self._fail("return opcode hit, but callstack already empty")
        if op == "ForeignCall":
            if self.fc_counter >= len(self.foreign_call_results):
                self.status = ("ForeignCallWait", o["function"])
                return self.status
            values = self.foreign_call_results[self.fc_counter]
            invalid = False
            for dest, out in zip(o["destinations"], values):
                if dest[0] == "Register":
                    if out[0] != "Single":
                        raise ReferencePanic("Function result size does not match brillig bytecode (expected 1 result)")
                    self.set(dest[1], out[1])
                elif dest[0] == "HeapArray":
                    if out[0] != "Array":
                        raise ReferencePanic("Function result size does not match brillig bytecode size")
                    if len(out[1]) != dest[2]:
                        invalid = True
                        break
                    self.mwrite(self.to_usize(self.get(dest[1])), list(out[1]))
                else:
                    if out[0] != "Array":
                        raise ReferencePanic("Function result size does not match brillig bytecode size")
                    self.set(dest[2], len(out[1]))
                    self.mwrite(self.to_usize(self.get(dest[1])), list(out[1]))
            if len(o["destinations"]) != len(values):
                self._fail(f"{len(values)} output values were provided as a foreign call result for {len(o['destinations'])} destination slots")
            if invalid:
                self._fail("Function result size does not match brillig bytecode")
            self.fc_counter += 1
            return self._set_pc(self.pc + 1)
        if op == "Mov":
            self.set(o["destination"], self.get(o["source"]))
            return self._set_pc(self.pc + 1)
        if op == "Trap":
            return self._fail("explicit trap hit in brillig")
        if op == "Stop":
            self.status = ("Finished",)
            return self.status
        if op == "Load":
            self.set(o["destination"], self.mread(self.to_usize(self.get(o["source_pointer"]))))
            return self._set_pc(self.pc + 1)
        if op == "Store":
            self.mwrite(self.to_usize(self.get(o["destination_pointer"])), [self.get(o["source"])])
            return self._set_pc(self.pc + 1)
        if op == "Call":
            self.call_stack.append(self.pc)
            return self._set_pc(o["location"])
        if op == "Const":
            self.set(o["destination"], o["value"])
            return self._set_pc(self.pc + 1)
        if op == "BlackBox":
            bb = o["bb"]
            n = bb["name"]
            if n in ("Sha256", "Keccak256", "Blake2s"):
                p, ln = self.to_usize(self.get(bb["message"][0])), self.to_usize(self.get(bb["message"][1]))
                if p + ln > len(self.memory):
                    raise ReferencePanic("memory slice out of range")
                msg = bytes(v & 0xFF for v in self.memory[p:p + ln])
                d = {"Sha256": hashes.sha256, "Keccak256": hashes.keccak256, "Blake2s": hashes.blake2s}[n](msg)
                self.mwrite(self.to_usize(self.get(bb["output"][0])), list(d))
            elif n == "HashToField128Security":  # black_box.rs:66-72
                p, ln = self.to_usize(self.get(bb["message"][0])), self.to_usize(self.get(bb["message"][1]))
                if p + ln > len(self.memory):
                    raise ReferencePanic("memory slice out of range")
                d = hashes.blake2s(bytes(v & 0xFF for v in self.memory[p:p + ln]))
                self.set(bb["output"], int.from_bytes(d, "big") % F.P)
            elif n in ("EcdsaSecp256k1", "EcdsaSecp256r1"):  # black_box.rs:73-127
                from . import ecdsa
                parts = []
                for key, size, label in (("public_key_x", 32, "public key x"), ("public_key_y", 32, "public key y"), ("signature", 64, "signature")):
                    p, ln = self.to_usize(self.get(bb[key][0])), bb[key][1]
                    if p + ln > len(self.memory):
                        raise ReferencePanic("memory slice out of range")
                    if ln != size:
                        return self._fail(f"Invalid {label} length")
                    parts.append(bytes(v & 0xFF for v in self.memory[p:p + ln]))
                p, ln = self.to_usize(self.get(bb["hashed_msg"][0])), self.to_usize(self.get(bb["hashed_msg"][1]))
                if p + ln > len(self.memory):
                    raise ReferencePanic("memory slice out of range")
                hashed = bytes(v & 0xFF for v in self.memory[p:p + ln])
                try:
                    ok = ecdsa.verify(n, hashed, *parts)
                except ecdsa.ReferencePanic as e:
                    raise ReferencePanic(str(e))
                self.set(bb["result"], 1 if ok else 0)
            elif n == "FixedBaseScalarMul":
                try:
                    x, y = self.backend.fixed_base_scalar_mul(self.get(bb["low"]), self.get(bb["high"]))
                except grumpkin.BlackBoxFailed as e:
                    return self._fail(str(e))
                self.mwrite(self.to_usize(self.get(bb["result"][0])), [x, y])
            else:
                raise NotImplementedError(f"oracle brillig blackbox {n}")
            return self._set_pc(self.pc + 1)
        raise ValueError(op)


def bigint_op(op, a, b, bs):  # arithmetic.rs:23-81
    m = 1 << bs
    if op == 0:
        return (a + b) % m
    if op == 1:
        if m + a < b:
            raise ReferencePanic("attempt to subtract with overflow")
        return (m + a - b) % m
    if op == 2:
        return (a * b) % m
    if op == 3:  # SignedDiv
        def signed(v):
            return v if v < (1 << (bs - 1)) else v - (1 << bs)
        sa, sb = signed(a), signed(b)
        if sb == 0:
            raise ReferencePanic("attempt to divide by zero")
        q = abs(sa) // abs(sb)
        if (sa < 0) != (sb < 0):
            q = -q
        if q >= 0:
            return q
        if (1 << bs) < -q:  # to_big_unsigned: BigUint subtraction underflow
            raise ReferencePanic("attempt to subtract with overflow")
        return (1 << bs) - (-q)
    if op == 4:
        if b % m == 0:
            raise ReferencePanic("attempt to divide by zero")
        return (a % m) // (b % m)
    if op == 5:
        return int(a % m == b % m)
    if op == 6:
        return int(a % m < b % m)
    if op == 7:
        return int(a % m <= b % m)
    if op == 8:
        return (a & b) % m
    if op == 9:
        return (a | b) % m
    if op == 10:
        return (a ^ b) % m
    if op in (11, 12):
        if bs > 128:
            raise ReferencePanic("unsupported bit size for right shift")
        if b.bit_length() > 128:
            raise ReferencePanic("shift does not fit u128")
        if op == 11:
            return 0 if b >= 512 else (a << b) % m
        return 0 if b >= 512 else (a >> b) % m
    raise ValueError(op)


def solve_brillig(wm, brillig, backend, acir_index):
    """BrilligSolver::solve (acvm/src/pwg/brillig.rs:20-131).  Returns None, or foreign-call wait info."""
    pred = get_value(brillig["predicate"], wm) if brillig["predicate"] is not None else 1
    if pred == 0:  # brillig.rs:133-150
        for kind, out in brillig["outputs"]:
            for w in ([out] if kind == "Simple" else out):
                insert_value(w, 0, wm)
        return None
    regs, mem = [], []
    for kind, e in brillig["inputs"]:
        try:
            if kind == "Single":
                regs.append(get_value(e, wm))
            else:
                ptr = len(mem)
                for x in e:
                    mem.append(get_value(x, wm))
                regs.append(ptr)
        except ResolutionError:
            raise ResolutionError("OpcodeNotSolvable.ExpressionHasTooManyUnknowns")
    vm = VM(regs, mem, brillig["bytecode"], brillig["foreign_call_results"], backend)
    st = vm.process_opcodes()
    if st[0] == "Finished":
        for i, (kind, out) in enumerate(brillig["outputs"]):
            reg = vm.get(i)
            if kind == "Simple":
                insert_value(out, reg, wm)
            else:
                base = VM.to_usize(reg)
                for j, w in enumerate(out):
                    if base + j >= len(vm.memory):
                        raise ReferencePanic("brillig output memory index out of range")
                    insert_value(w, vm.memory[base + j], wm)
        return None
    if st[0] == "Failure":
        raise ResolutionError("BrilligFunctionFailed", message=st[1], index=st[2][-1])
    return {"function": st[1]}
