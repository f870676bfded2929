"""ACIR wire-format decoder: Circuit::read = gunzip . bincode (acir/src/circuit/mod.rs:155-161).

TEST INFRASTRUCTURE ONLY.  Independent of the product's C++ decoder and Python encoder
(acvm_b200/acir_builder.py), so a round trip through both is a real cross-check.

bincode 1.3.3 defaults: little-endian fixed-width ints, usize -> u64, enum tag u32 in
declaration order, Vec/String/BTreeSet = u64 length + items, Option = u8 tag, structs and
tuples = fields in order.  Struct/enum declaration order is cited per reader.
"""
import gzip
import struct
from dataclasses import dataclass, field as dfield
from typing import Any, List, Optional, Tuple

from . import field as F


class Reader:
    def __init__(self, data: bytes):
        self.d = data
        self.o = 0

    def u8(self):
        v = self.d[self.o]
        self.o += 1
        return v

    def u32(self):
        v = struct.unpack_from("<I", self.d, self.o)[0]
        self.o += 4
        return v

    def u64(self):
        v = struct.unpack_from("<Q", self.d, self.o)[0]
        self.o += 8
        return v

    def string(self):
        n = self.u64()
        s = self.d[self.o:self.o + n]
        if len(s) != n:
            raise ValueError("truncated string")
        self.o += n
        return s.decode("utf-8")

    def vec(self, item):
        return [item() for _ in range(self.u64())]

    def option(self, item):
        t = self.u8()
        if t == 0:
            return None
        if t == 1:
            return item()
        raise ValueError(f"bad Option tag {t}")

    # FieldElement serialises as its hex string (acir_field/src/generic_ark.rs:114-134)
    def fe(self):
        return F.from_hex(self.string())


@dataclass
class Expression:  # acir/src/native_types/expression/mod.rs:17-28
    mul_terms: List[Tuple[int, int, int]] = dfield(default_factory=list)
    linear_combinations: List[Tuple[int, int]] = dfield(default_factory=list)
    q_c: int = 0

    def to_const(self):  # expression/mod.rs is_const/to_const
        if not self.mul_terms and not self.linear_combinations:
            return self.q_c
        return None

    def to_witness(self):  # expression/mod.rs:to_witness -- exactly 1*w + 0
        if (not self.mul_terms and len(self.linear_combinations) == 1 and self.q_c == 0
                and self.linear_combinations[0][0] == 1):
            return self.linear_combinations[0][1]
        return None


@dataclass
class Opcode:
    kind: str
    body: Any


@dataclass
class Circuit:  # acir/src/circuit/mod.rs:18-41
    current_witness_index: int
    opcodes: List[Opcode]
    private_parameters: List[int]
    public_parameters: List[int]
    return_values: List[int]
    assert_messages: List[Tuple[Any, str]]


def r_expression(r: Reader) -> Expression:
    mul = r.vec(lambda: (r.fe(), r.u32(), r.u32()))
    lin = r.vec(lambda: (r.fe(), r.u32()))
    return Expression(mul, lin, r.fe())


def r_function_input(r):  # black_box_function_call.rs:8-11
    return (r.u32(), r.u32())


BLACKBOX_NAMES = ["AND", "XOR", "RANGE", "SHA256", "Blake2s", "SchnorrVerify", "Pedersen",
                  "HashToField128Security", "EcdsaSecp256k1", "EcdsaSecp256r1", "FixedBaseScalarMul",
                  "Keccak256", "Keccak256VariableLength", "RecursiveAggregation"]


def r_blackbox(r):  # black_box_function_call.rs:20-115 (declaration order)
    tag = r.u32()
    fi = lambda: r_function_input(r)
    vfi = lambda: r.vec(fi)
    name = BLACKBOX_NAMES[tag]
    if name in ("AND", "XOR"):
        return dict(name=name, lhs=fi(), rhs=fi(), output=r.u32())
    if name == "RANGE":
        return dict(name=name, input=fi())
    if name in ("SHA256", "Blake2s", "Keccak256"):
        return dict(name=name, inputs=vfi(), outputs=r.vec(r.u32))
    if name == "SchnorrVerify":
        return dict(name=name, public_key_x=fi(), public_key_y=fi(), signature=vfi(), message=vfi(), output=r.u32())
    if name == "Pedersen":
        return dict(name=name, inputs=vfi(), domain_separator=r.u32(), outputs=(r.u32(), r.u32()))
    if name == "HashToField128Security":
        return dict(name=name, inputs=vfi(), output=r.u32())
    if name in ("EcdsaSecp256k1", "EcdsaSecp256r1"):
        return dict(name=name, public_key_x=vfi(), public_key_y=vfi(), signature=vfi(), hashed_message=vfi(), output=r.u32())
    if name == "FixedBaseScalarMul":
        return dict(name=name, low=fi(), high=fi(), outputs=(r.u32(), r.u32()))
    if name == "Keccak256VariableLength":
        return dict(name=name, inputs=vfi(), var_message_size=fi(), outputs=r.vec(r.u32))
    if name == "RecursiveAggregation":
        return dict(name=name, verification_key=vfi(), proof=vfi(), public_inputs=vfi(), key_hash=fi(),
                    input_aggregation_object=r.option(vfi), output_aggregation_object=r.vec(r.u32))
    raise ValueError(tag)


def r_directive(r):  # acir/src/circuit/directives.rs:5-36
    tag = r.u32()
    if tag == 0:
        return dict(name="Quotient", a=r_expression(r), b=r_expression(r), q=r.u32(), r=r.u32(),
                    predicate=r.option(lambda: r_expression(r)))
    if tag == 1:
        return dict(name="ToLeRadix", a=r_expression(r), b=r.vec(r.u32), radix=r.u32())
    if tag == 2:
        return dict(name="PermutationSort", inputs=r.vec(lambda: r.vec(lambda: r_expression(r))), tuple=r.u32(),
                    bits=r.vec(r.u32), sort_by=r.vec(r.u32))
    raise ValueError(f"bad Directive tag {tag}")


def r_reg_or_mem(r):  # brillig/src/opcodes.rs:42-57
    tag = r.u32()
    if tag == 0:
        return ("Register", r.u64())
    if tag == 1:
        return ("HeapArray", r.u64(), r.u64())
    if tag == 2:
        return ("HeapVector", r.u64(), r.u64())
    raise ValueError(tag)


def r_brillig_blackbox(r):  # brillig/src/black_box.rs:7-48
    tag = r.u32()
    hv = lambda: (r.u64(), r.u64())
    ha = lambda: (r.u64(), r.u64())
    if tag in (0, 1, 2):
        return dict(name=["Sha256", "Blake2s", "Keccak256"][tag], message=hv(), output=ha())
    if tag == 3:
        return dict(name="HashToField128Security", message=hv(), output=r.u64())
    if tag in (4, 5):
        return dict(name=["EcdsaSecp256k1", "EcdsaSecp256r1"][tag - 4], hashed_msg=hv(), public_key_x=ha(),
                    public_key_y=ha(), signature=ha(), result=r.u64())
    if tag == 6:
        return dict(name="SchnorrVerify", public_key_x=r.u64(), public_key_y=r.u64(), message=hv(), signature=hv(),
                    result=r.u64())
    if tag == 7:
        return dict(name="Pedersen", inputs=hv(), domain_separator=r.u64(), output=ha())
    if tag == 8:
        return dict(name="FixedBaseScalarMul", low=r.u64(), high=r.u64(), result=ha())
    raise ValueError(tag)


def r_brillig_opcode(r):  # brillig/src/opcodes.rs:60-134
    tag = r.u32()
    if tag == 0:
        return dict(op="BinaryFieldOp", destination=r.u64(), bop=r.u32(), lhs=r.u64(), rhs=r.u64())
    if tag == 1:
        return dict(op="BinaryIntOp", destination=r.u64(), bop=r.u32(), bit_size=r.u32(), lhs=r.u64(), rhs=r.u64())
    if tag == 2:
        return dict(op="JumpIfNot", condition=r.u64(), location=r.u64())
    if tag == 3:
        return dict(op="JumpIf", condition=r.u64(), location=r.u64())
    if tag == 4:
        return dict(op="Jump", location=r.u64())
    if tag == 5:
        return dict(op="Call", location=r.u64())
    if tag == 6:
        return dict(op="Const", destination=r.u64(), value=r.fe())
    if tag == 7:
        return dict(op="Return")
    if tag == 8:
        return dict(op="ForeignCall", function=r.string(), destinations=r.vec(lambda: r_reg_or_mem(r)),
                    inputs=r.vec(lambda: r_reg_or_mem(r)))
    if tag == 9:
        return dict(op="Mov", destination=r.u64(), source=r.u64())
    if tag == 10:
        return dict(op="Load", destination=r.u64(), source_pointer=r.u64())
    if tag == 11:
        return dict(op="Store", destination_pointer=r.u64(), source=r.u64())
    if tag == 12:
        return dict(op="BlackBox", bb=r_brillig_blackbox(r))
    if tag == 13:
        return dict(op="Trap")
    if tag == 14:
        return dict(op="Stop")
    raise ValueError(f"bad brillig opcode tag {tag}")


def r_brillig(r):  # acir/src/circuit/brillig.rs:9-33
    def inp():
        t = r.u32()
        if t == 0:
            return ("Single", r_expression(r))
        if t == 1:
            return ("Array", r.vec(lambda: r_expression(r)))
        raise ValueError(t)

    def outp():
        t = r.u32()
        if t == 0:
            return ("Simple", r.u32())
        if t == 1:
            return ("Array", r.vec(r.u32))
        raise ValueError(t)

    def fco():  # brillig/src/foreign_call.rs:6-16
        t = r.u32()
        if t == 0:
            return ("Single", r.fe())
        if t == 1:
            return ("Array", r.vec(r.fe))
        raise ValueError(t)

    return dict(inputs=r.vec(inp), outputs=r.vec(outp), foreign_call_results=r.vec(lambda: r.vec(fco)),
                bytecode=r.vec(lambda: r_brillig_opcode(r)), predicate=r.option(lambda: r_expression(r)))


def r_opcode(r) -> Opcode:  # acir/src/circuit/opcodes.rs:15-34
    tag = r.u32()
    if tag == 0:
        return Opcode("Arithmetic", r_expression(r))
    if tag == 1:
        return Opcode("BlackBoxFuncCall", r_blackbox(r))
    if tag == 2:
        return Opcode("Directive", r_directive(r))
    if tag == 3:
        return Opcode("Brillig", r_brillig(r))
    if tag == 4:  # MemoryOp { block_id, op: MemOp{operation,index,value}, predicate }
        return Opcode("MemoryOp", dict(block_id=r.u32(), operation=r_expression(r), index=r_expression(r),
                                       value=r_expression(r), predicate=r.option(lambda: r_expression(r))))
    if tag == 5:
        return Opcode("MemoryInit", dict(block_id=r.u32(), init=r.vec(r.u32)))
    raise ValueError(f"bad Opcode tag {tag}")


def r_opcode_location(r):  # acir/src/circuit/mod.rs:57-60
    t = r.u32()
    if t == 0:
        return ("Acir", r.u64())
    if t == 1:
        return ("Brillig", r.u64(), r.u64())
    raise ValueError(t)


def decode_circuit(data: bytes) -> Circuit:
    raw = gzip.decompress(bytes(data))
    r = Reader(raw)
    cwi = r.u32()
    opcodes = r.vec(lambda: r_opcode(r))
    priv = r.vec(r.u32)
    pub = r.vec(r.u32)
    ret = r.vec(r.u32)
    msgs = r.vec(lambda: (r_opcode_location(r), r.string()))
    if r.o != len(raw):
        raise ValueError(f"trailing bytes: {len(raw) - r.o}")
    return Circuit(cwi, opcodes, priv, pub, ret, msgs)


def decode_witness_map(data: bytes):  # acir/src/native_types/witness_map.rs:108-146
    raw = gzip.decompress(bytes(data))
    r = Reader(raw)
    n = r.u64()
    out = {}
    for _ in range(n):
        k = r.u32()
        out[k] = r.fe()
    if r.o != len(raw):
        raise ValueError("trailing bytes")
    return out
