"""ctypes front-end of oracle/ref_solver.cpp (the C++ reference-algorithm restatement).

TEST / BENCH INFRASTRUCTURE ONLY.  Packs the oracle-decoded circuit into the flat u64 stream the
C++ solver parses, runs `threads` solver instances in parallel, returns statuses (+ witnesses).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import acir as oacir

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libref_solver.so")
ERR_NAMES = {0: None, 1: "OpcodeNotSolvable.MissingAssignment", 2: "OpcodeNotSolvable.ExpressionHasTooManyUnknowns",
             4: "UnsatisfiedConstrain", 6: "BlackBoxFunctionFailed", 8: "ReferencePanic"}


def build(force=False):
    src = os.path.join(_HERE, "ref_solver.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "_build/libref_solver.so"], check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.ref_solve_batch.restype = C.c_int
    return _lib


_ped_ready = False


def _ensure_pedersen_tables():
    """Hand the oracle's generators (oracle/pedersen.py; parity with barretenberg UNPINNED) to the C++ side, once."""
    global _ped_ready
    if _ped_ready:
        return
    from . import grumpkin, pedersen
    g = pedersen.generators()
    pts = [g[par][i] for par in range(2) for i in range(pedersen.NUM_WINDOWS)] + [grumpkin.G]
    arr = np.array([_limbs(c) for pt in pts for c in pt], dtype=np.uint64)
    lib().ref_set_pedersen_generators(arr.ctypes.data_as(C.c_void_p))
    _ped_ready = True


def _limbs(v):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def pack_circuit(circuit: oacir.Circuit) -> np.ndarray:
    words = []
    for op in circuit.opcodes:
        if op.kind == "Arithmetic":
            e = op.body
            words += [0, len(e.mul_terms), len(e.linear_combinations)]
            for (c, a, b) in e.mul_terms:
                words += _limbs(c) + [a, b]
            for (c, w) in e.linear_combinations:
                words += _limbs(c) + [w]
            words += _limbs(e.q_c)
        elif op.kind == "BlackBoxFuncCall" and op.body["name"] in ("AND", "XOR"):
            bb = op.body
            words += [1 if bb["name"] == "AND" else 2, bb["lhs"][0], bb["rhs"][0], bb["lhs"][1], bb["output"]]
        elif op.kind == "BlackBoxFuncCall" and op.body["name"] == "RANGE":
            words += [3, op.body["input"][0], op.body["input"][1]]
        elif op.kind == "BlackBoxFuncCall" and op.body["name"] in ("SHA256", "Keccak256"):
            bb = op.body
            words += [4 if bb["name"] == "SHA256" else 5, len(bb["inputs"])]
            for (w, nb) in bb["inputs"]:
                words += [w, nb]
            words += list(bb["outputs"])
        elif op.kind == "BlackBoxFuncCall" and op.body["name"] == "FixedBaseScalarMul":
            bb = op.body
            words += [6, bb["low"][0], bb["high"][0], bb["outputs"][0], bb["outputs"][1]]
        elif op.kind == "BlackBoxFuncCall" and op.body["name"] == "Pedersen":
            bb = op.body
            _ensure_pedersen_tables()
            words += [7, len(bb["inputs"])] + [w for (w, _) in bb["inputs"]] + [bb["domain_separator"], bb["outputs"][0], bb["outputs"][1]]
        else:
            raise NotImplementedError(f"ref_solver.cpp does not restate {op.kind}/{op.body.get('name') if isinstance(op.body, dict) else ''}")
    return np.array(words, dtype=np.uint64)


def solve_batch(circuit: oacir.Circuit, input_ids, inputs_be32: bytes, n_inst: int, n_witnesses: int, threads: int = 1,
                want_witness: bool = False, packed=None):
    stream = pack_circuit(circuit) if packed is None else packed
    n_in = len(input_ids)
    be = np.frombuffer(inputs_be32, dtype=np.uint8).reshape(n_inst * n_in, 32)
    limbs = np.ascontiguousarray(be[:, ::-1]).view("<u8").reshape(n_inst, n_in, 4).copy()
    ids = np.array(list(input_ids), dtype=np.uint32)
    res = np.zeros((n_inst, 4), dtype=np.uint32)
    outw = outp = None
    pw = pp = None
    if want_witness:
        outw = np.zeros((n_inst, n_witnesses, 4), dtype=np.uint64)
        outp = np.zeros((n_inst, n_witnesses), dtype=np.uint8)
        pw, pp = outw.ctypes.data_as(C.c_void_p), outp.ctypes.data_as(C.c_void_p)
    lib().ref_solve_batch(stream.ctypes.data_as(C.c_void_p), C.c_uint64(len(stream)), C.c_uint64(len(circuit.opcodes)),
                          ids.ctypes.data_as(C.c_void_p), C.c_uint32(n_in), limbs.ctypes.data_as(C.c_void_p), C.c_uint32(n_inst),
                          C.c_uint32(n_witnesses), pw, pp, res.ctypes.data_as(C.c_void_p), C.c_uint32(threads))
    return res, outw, outp


def solve_batch_optimized(circuit: oacir.Circuit, input_ids, inputs_be32: bytes, n_inst: int, n_witnesses: int, threads: int = 1,
                         want_witness: bool = False, packed=None):
    """The "optimised CPU" variant: dense witness vector, plan resolved once, inverses hoisted (ref_solver.cpp)."""
    stream = pack_circuit(circuit) if packed is None else packed
    n_in = len(input_ids)
    be = np.frombuffer(inputs_be32, dtype=np.uint8).reshape(n_inst * n_in, 32)
    limbs = np.ascontiguousarray(be[:, ::-1]).view("<u8").reshape(n_inst, n_in, 4).copy()
    ids = np.array(list(input_ids), dtype=np.uint32)
    res = np.zeros((n_inst, 4), dtype=np.uint32)
    outw = np.zeros((n_inst, n_witnesses, 4), dtype=np.uint64) if want_witness else None
    lib().ref_solve_batch_optimized.restype = C.c_int
    rc = lib().ref_solve_batch_optimized(stream.ctypes.data_as(C.c_void_p), C.c_uint64(len(stream)), C.c_uint64(len(circuit.opcodes)),
                                         ids.ctypes.data_as(C.c_void_p), C.c_uint32(n_in), limbs.ctypes.data_as(C.c_void_p),
                                         C.c_uint32(n_inst), C.c_uint32(n_witnesses),
                                         outw.ctypes.data_as(C.c_void_p) if want_witness else None,
                                         res.ctypes.data_as(C.c_void_p), C.c_uint32(threads))
    if rc != 0:
        raise NotImplementedError("optimised CPU variant does not cover this circuit (value-dependent or unsolvable gate)")
    return res, outw


def witness_dict(outw, outp, i):
    return {w: int.from_bytes(outw[i, w].tobytes(), "little") for w in range(outw.shape[1]) if outp[i, w]}
