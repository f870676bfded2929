"""Host-side mirror of the reference's solver surface, on top of the C ABI.

Reference interface mirrored (acvm/src/pwg/mod.rs):
  ACVM::new / solve / get_status / witness_map / instruction_pointer / finalize   :145-241
  ACVMStatus / OpcodeResolutionError                                             :33-127
plus the batch form this project exists for: the same circuit, many initial witnesses.
Everything numerical happens in libacvm_b200.so on the GPU; this file only marshals buffers.
"""
import ctypes as C
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

from . import _lib

STATUS_NAMES = {0: "Solved", 1: "InProgress", 2: "Failure", 3: "RequiresForeignCall"}
ERR_NAMES = {0: None, 1: "OpcodeNotSolvable.MissingAssignment", 2: "OpcodeNotSolvable.ExpressionHasTooManyUnknowns",
             3: "UnsupportedBlackBoxFunc", 4: "UnsatisfiedConstrain", 5: "IndexOutOfBounds", 6: "BlackBoxFunctionFailed",
             7: "BrilligFunctionFailed", 8: "ReferencePanic"}


class AcvmError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(f"acvm_b200 rc={rc}: {msg}")
        self.rc = rc


@dataclass
class InstanceStatus:
    status: str
    error: Optional[str]
    opcode_index: int
    aux: int


_lib_handle = None


def lib():
    global _lib_handle
    if _lib_handle is None:
        _lib_handle = _lib.load()
    return _lib_handle


def _check(rc):
    if rc != 0:
        raise AcvmError(rc, lib().acvmb_last_error().decode(errors="replace"))


def _u32_array(xs):
    return (C.c_uint32 * len(xs))(*xs) if len(xs) else None


class Context:
    """One CUDA device.  Fails loudly when there is no sm_100 GPU (no CPU fallback)."""

    def __init__(self, device=0, **options):
        """device: one CUDA device index, or a list of indices for a multi-device context (acvmb_ctx_create_multi): circuits are
        replicated to every device with one broadcast and solve_batch shards the batch over them."""
        self._h = C.c_void_p()
        self._children = weakref.WeakSet()   # circuits / VMs created on this context: closed before the context itself
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*device)
            _check(lib().acvmb_ctx_create_multi(devs, len(device), C.byref(self._h)))
        else:
            _check(lib().acvmb_ctx_create(device, C.byref(self._h)))
        for k, v in options.items():
            self.set_option(k, v)

    def n_devices(self):
        return lib().acvmb_ctx_n_devices(self._h)

    def broadcast_backend(self):
        return lib().acvmb_ctx_broadcast_backend(self._h).decode()

    def set_option(self, key, value):
        _check(lib().acvmb_ctx_set_option(self._h, key.encode(), int(value)))

    def device_name(self):
        buf = C.create_string_buffer(256)
        _check(lib().acvmb_device_name(self._h, buf, 256))
        return buf.value.decode()

    def imad_microbench(self):
        v = [C.c_double() for _ in range(4)]
        _check(lib().acvmb_imad_microbench(self._h, *[C.byref(x) for x in v]))
        f = (C.c_double * 5)()
        _check(lib().acvmb_frmul_microbench(self._h, f))
        return dict(imad32_per_s=v[0].value, imad_wide_per_s=v[1].value, imad_wide_carry_per_s=v[2].value,
                    sm_clock_mhz=v[3].value, fr_mul_per_s_by_split=list(f), fr_mul_per_s=max(f))

    def close(self):
        if self._h:
            for child in list(self._children):   # the C objects hold a pointer to the context: they must go first
                child.close()
            lib().acvmb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- BlackBoxFunctionSolver trait, batched (blackbox_solver/src/lib.rs:27-45) ----
    def fixed_base_scalar_mul(self, lows: Sequence[int], highs: Sequence[int]):
        n = len(lows)
        lo = b"".join(int(v).to_bytes(32, "big") for v in lows)
        hi = b"".join(int(v).to_bytes(32, "big") for v in highs)
        out = C.create_string_buffer(n * 64)
        st = (_lib.Status * n)()
        _check(lib().acvmb_fixed_base_scalar_mul(self._h, lo, hi, n, out, st))
        raw = out.raw
        pts = [(int.from_bytes(raw[i * 64:i * 64 + 32], "big"), int.from_bytes(raw[i * 64 + 32:i * 64 + 64], "big")) for i in range(n)]
        return pts, [_status(s) for s in st]

    def pedersen(self, inputs: Sequence[Sequence[int]], domain_separator: int = 0):
        n = len(inputs)
        k = len(inputs[0]) if n else 0
        buf = b"".join(int(v).to_bytes(32, "big") for row in inputs for v in row)
        out = C.create_string_buffer(n * 64)
        st = (_lib.Status * n)()
        _check(lib().acvmb_pedersen(self._h, buf, k, n, domain_separator, out, st))
        raw = out.raw
        pts = [(int.from_bytes(raw[i * 64:i * 64 + 32], "big"), int.from_bytes(raw[i * 64 + 32:i * 64 + 64], "big")) for i in range(n)]
        return pts, [_status(s) for s in st]

    def sha256(self, msgs: Sequence[bytes]):
        return self._hash("acvmb_sha256", msgs)

    def keccak256(self, msgs: Sequence[bytes]):
        return self._hash("acvmb_keccak256", msgs)

    def ecdsa_verify(self, curve: str, hashed_msgs: Sequence[bytes], public_keys_x: Sequence[bytes], public_keys_y: Sequence[bytes],
                     signatures: Sequence[bytes]):
        """ecdsa_secp256k1_verify / ecdsa_secp256r1_verify (blackbox_solver/src/lib.rs:67-83), batched.
        curve: "secp256k1" | "secp256r1".  Returns ([bool], [InstanceStatus]); reference panics show up in the status."""
        n = len(hashed_msgs)
        assert all(len(x) == 32 for x in list(hashed_msgs) + list(public_keys_x) + list(public_keys_y)) and all(len(x) == 64 for x in signatures)
        out = C.create_string_buffer(n)
        st = (_lib.Status * n)()
        fn = {"secp256k1": lib().acvmb_ecdsa_secp256k1_verify, "secp256r1": lib().acvmb_ecdsa_secp256r1_verify}[curve]
        _check(fn(self._h, b"".join(hashed_msgs), b"".join(public_keys_x), b"".join(public_keys_y), b"".join(signatures), n, out, st))
        return [bool(b) for b in out.raw], [_status(s) for s in st]

    def _hash(self, fn, msgs):
        n = len(msgs)
        ln = len(msgs[0]) if n else 0
        assert all(len(m) == ln for m in msgs), "one call hashes messages of one length"
        out = C.create_string_buffer(n * 32)
        _check(getattr(lib(), fn)(self._h, b"".join(msgs), ln, n, out))
        return [out.raw[i * 32:(i + 1) * 32] for i in range(n)]


def _status(s) -> InstanceStatus:
    return InstanceStatus(STATUS_NAMES[s.code], ERR_NAMES.get(s.err_kind), int(s.opcode_index), int(s.aux))


class CompiledCircuit:
    """Circuit::read + plan compilation for a fixed set of initial-witness indices."""

    def __init__(self, ctx: Context, acir_bytes: Optional[bytes], input_witnesses: Sequence[int] = (), _blob: Optional[bytes] = None):
        self.ctx = ctx
        self._h = C.c_void_p()
        self.input_witnesses = list(input_witnesses)
        if _blob is not None:
            buf = (C.c_uint8 * len(_blob)).from_buffer_copy(_blob)
            _check(lib().acvmb_circuit_deserialize(ctx._h, buf, len(_blob), C.byref(self._h)))
        else:
            _check(lib().acvmb_circuit_from_acir(ctx._h, acir_bytes, len(acir_bytes), _u32_array(self.input_witnesses),
                                                 len(self.input_witnesses), C.byref(self._h)))
        info = _lib.PlanInfo()
        _check(lib().acvmb_circuit_info(self._h, C.byref(info)))
        self.info = info.as_dict()
        self.num_witnesses = self.info["num_witnesses"]
        self._batches = weakref.WeakSet()
        ctx._children.add(self)

    @classmethod
    def from_blob(cls, ctx, blob: bytes, input_witnesses: Sequence[int]):
        return cls(ctx, None, input_witnesses, _blob=blob)

    def serialize(self) -> bytes:
        need = C.c_size_t()
        _check(lib().acvmb_circuit_serialize(self._h, None, 0, C.byref(need)))
        buf = (C.c_uint8 * need.value)()
        _check(lib().acvmb_circuit_serialize(self._h, buf, need.value, C.byref(need)))
        return bytes(buf)

    def assign_opcodes(self) -> List[int]:
        arr = (C.c_uint32 * self.num_witnesses)()
        _check(lib().acvmb_circuit_assign_opcodes(self._h, arr, self.num_witnesses))
        return list(arr)

    def run_info(self):
        ri = _lib.RunInfo()
        _check(lib().acvmb_last_run_info(self._h, C.byref(ri)))
        return {n: getattr(ri, n) for n, _ in ri._fields_}

    def solve_batch(self, inputs_be32, batch: int, out_ids: Optional[Sequence[int]] = None, want_witness=True, out_buffer=None,
                    want_present=False):
        """inputs_be32: bytes-like [batch][n_inputs][32].  Returns (witness bytes or None, [InstanceStatus]) and, with
        want_present, a third value: bytes [batch][n_out] (1 = the instance's witness map holds that witness)."""
        n_in = len(self.input_witnesses)
        assert len(inputs_be32) == batch * n_in * 32, "inputs must be [batch][n_inputs][32]"
        n_out = len(out_ids) if out_ids is not None else self.num_witnesses
        out = None
        outp = None
        if want_witness:
            if out_buffer is not None:
                outp = out_buffer
            else:
                out = C.create_string_buffer(batch * n_out * 32)
                outp = out
        st = (_lib.Status * batch)()
        inp = (C.c_uint8 * len(inputs_be32)).from_buffer_copy(inputs_be32) if not isinstance(inputs_be32, C.Array) else inputs_be32
        ids = _u32_array(list(out_ids)) if out_ids is not None else None
        pres = C.create_string_buffer(batch * n_out) if want_present else None
        _check(lib().acvmb_solve_batch_ex(self._h, batch, inp, ids, len(out_ids) if out_ids is not None else 0, outp, pres, st))
        if want_present:
            return (out.raw if out is not None else None), [_status(s) for s in st], pres.raw
        return (out.raw if out is not None else None), [_status(s) for s in st]

    def close(self):
        if self._h:
            for b in list(self._batches):
                b.close()
            lib().acvmb_circuit_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceBatch:
    """Witness columns of `n` instances resident in HBM (acvmb_batch_*)."""

    def __init__(self, circuit: CompiledCircuit, n: int):
        self.circuit = circuit
        self.n = n
        self._h = C.c_void_p()
        _check(lib().acvmb_batch_create(circuit._h, n, C.byref(self._h)))
        circuit._batches.add(self)

    def resize(self, n: int):
        _check(lib().acvmb_batch_resize(self._h, n))
        self.n = n

    def upload(self, inputs_be32):
        buf = (C.c_uint8 * len(inputs_be32)).from_buffer_copy(inputs_be32) if not isinstance(inputs_be32, (C.Array, int)) else inputs_be32
        _check(lib().acvmb_batch_upload(self._h, buf))

    def run(self) -> float:
        ms = C.c_float()
        _check(lib().acvmb_batch_run(self._h, C.byref(ms)))
        return ms.value

    def stage_inputs(self, slot: int, inputs_be32):
        buf = (C.c_uint8 * len(inputs_be32)).from_buffer_copy(inputs_be32) if not isinstance(inputs_be32, C.Array) else inputs_be32
        _check(lib().acvmb_batch_stage_inputs(self._h, slot, buf))

    def run_staged(self, slot: int):
        """status reset + input scatter + step-VM kernel from resident inputs; returns (total_ms, vm_kernel_ms)."""
        t, v = C.c_float(), C.c_float()
        _check(lib().acvmb_batch_run_staged(self._h, slot, C.byref(t), C.byref(v)))
        return t.value, v.value

    def status(self):
        st = (_lib.Status * self.n)()
        _check(lib().acvmb_batch_status(self._h, st))
        return [_status(s) for s in st]

    def checksums(self):
        """Per-instance checksum of the solved witness map, computed on the device (see witness_checksum())."""
        out = (C.c_uint64 * self.n)()
        _check(lib().acvmb_batch_checksum(self._h, out))
        return list(out)

    def download(self, first=0, n=None, out_ids=None, out_buffer=None):
        n = self.n - first if n is None else n
        n_out = len(out_ids) if out_ids is not None else self.circuit.num_witnesses
        out = out_buffer if out_buffer is not None else C.create_string_buffer(n * n_out * 32)
        ids = _u32_array(list(out_ids)) if out_ids is not None else None
        _check(lib().acvmb_batch_download(self._h, first, n, ids, len(out_ids) if out_ids is not None else 0, out))
        return out.raw if out_buffer is None else out_buffer

    def close(self):
        if self._h:
            lib().acvmb_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ACVM:
    """Single-instance mirror of acvm::pwg::ACVM (acvm/src/pwg/mod.rs:129-304) -- a batch of one."""

    def __init__(self, ctx: Context, acir_bytes: bytes, initial_witness: Dict[int, int]):
        self.ctx = ctx
        self._h = C.c_void_p()
        keys = sorted(initial_witness)
        vals = b"".join(int(initial_witness[k]).to_bytes(32, "big") for k in keys)
        _check(lib().acvmb_vm_new(ctx._h, acir_bytes, len(acir_bytes), _u32_array(keys), vals, len(keys), C.byref(self._h)))
        ctx._children.add(self)

    def solve(self) -> InstanceStatus:
        st = _lib.Status()
        _check(lib().acvmb_vm_solve(self._h, C.byref(st)))
        return _status(st)

    def solve_opcode(self) -> InstanceStatus:
        """ACVM::solve_opcode (acvm/src/pwg/mod.rs:243-303): one opcode per call."""
        st = _lib.Status()
        _check(lib().acvmb_vm_solve_opcode(self._h, C.byref(st)))
        return _status(st)

    def get_status(self) -> InstanceStatus:
        st = _lib.Status()
        _check(lib().acvmb_vm_status(self._h, C.byref(st)))
        return _status(st)

    def instruction_pointer(self) -> int:
        v = C.c_uint32()
        _check(lib().acvmb_vm_instruction_pointer(self._h, C.byref(v)))
        return v.value

    def witness_map(self) -> Dict[int, int]:
        n = C.c_uint32()
        _check(lib().acvmb_vm_num_witnesses(self._h, C.byref(n)))
        out = {}
        buf = C.create_string_buffer(32)
        present = C.c_int()
        for w in range(n.value):
            _check(lib().acvmb_vm_witness(self._h, w, buf, C.byref(present)))
            if present.value:
                out[w] = int.from_bytes(buf.raw, "big")
        return out

    def get_pending_foreign_call(self):
        """ACVM::get_pending_foreign_call: (function, [[values of input 0], ...]) or None."""
        if self.get_status().status != "RequiresForeignCall":
            return None
        name = C.create_string_buffer(256)
        n_in, n_val = C.c_uint32(), C.c_uint32()
        lens = (C.c_uint32 * 64)()
        vals = C.create_string_buffer(4096 * 32)
        _check(lib().acvmb_vm_pending_foreign_call(self._h, name, 256, C.byref(n_in), lens, 64, vals, 4096, C.byref(n_val)))
        out, k = [], 0
        for i in range(n_in.value):
            out.append([int.from_bytes(vals.raw[(k + j) * 32:(k + j + 1) * 32], "big") for j in range(lens[i])])
            k += lens[i]
        return name.value.decode(), out

    def resolve_pending_foreign_call(self, outputs):
        """outputs: list of int (ForeignCallOutput::Single) or list of ints (ForeignCallOutput::Array)."""
        lens, flat = [], []
        for o in outputs:
            if isinstance(o, int):
                lens.append(0xFFFFFFFF)
                flat.append(o)
            else:
                lens.append(len(o))
                flat += list(o)
        buf = b"".join(int(v).to_bytes(32, "big") for v in flat)
        _check(lib().acvmb_vm_resolve_foreign_call(self._h, len(lens), _u32_array(lens), buf))

    def finalize(self) -> Dict[int, int]:
        n = C.c_uint32()
        _check(lib().acvmb_vm_num_witnesses(self._h, C.byref(n)))
        vals = C.create_string_buffer(n.value * 32)
        pres = C.create_string_buffer(n.value)
        _check(lib().acvmb_vm_finalize(self._h, vals, pres, n.value))
        return {w: int.from_bytes(vals.raw[w * 32:(w + 1) * 32], "big") for w in range(n.value) if pres.raw[w]}

    def close(self):
        if self._h:
            lib().acvmb_vm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def witness_checksum(wm: Dict[int, int]) -> int:
    """Host-side definition of acvmb_batch_checksum for one witness map (used by tests to cross-check the device)."""
    M = (1 << 64) - 1
    total = 0
    for w, v in wm.items():
        h = ((w + 1) * 0x9E3779B97F4A7C15) & M
        for k in range(8):
            h = ((h ^ ((v >> (32 * k)) & 0xFFFFFFFF)) * 0x100000001B3) & M
        total = (total + h) & M
    return total


def compress_witness_map(wm: Dict[int, int]) -> bytes:
    """WitnessMap -> gzip(bincode) bytes the reference's `WitnessMap::try_from(&[u8])` accepts (witness_map.rs:108-146)."""
    keys = sorted(wm)
    vals = b"".join(int(wm[k]).to_bytes(32, "big") for k in keys)
    need = C.c_size_t()
    _check(lib().acvmb_witness_map_compress(_u32_array(keys), vals, len(keys), None, 0, C.byref(need)))
    buf = (C.c_uint8 * need.value)()
    _check(lib().acvmb_witness_map_compress(_u32_array(keys), vals, len(keys), buf, need.value, C.byref(need)))
    return bytes(buf[:need.value])


def decompress_witness_map(data: bytes) -> Dict[int, int]:
    n = C.c_uint32()
    _check(lib().acvmb_witness_map_decompress(data, len(data), None, None, 0, C.byref(n)))
    idx = (C.c_uint32 * max(1, n.value))()
    vals = C.create_string_buffer(max(1, n.value) * 32)
    _check(lib().acvmb_witness_map_decompress(data, len(data), idx, vals, n.value, C.byref(n)))
    return {idx[i]: int.from_bytes(vals.raw[i * 32:(i + 1) * 32], "big") for i in range(n.value)}


def compile_plan_host(acir_bytes: bytes, input_witnesses: Sequence[int], S: int = 16, temp_pool: int = 0,
                      pedersen_unpinned: bool = False, device_brillig: bool = True, scaled_columns: bool = True,
                      ring_slots: int = None, packed_hashes: bool = True, spread_heavy: bool = True,
                      slack_scheduling: bool = True):
    """Decode + compile on the host only (no device): returns (info dict, plan blob)."""
    info = _lib.PlanInfo()
    need = C.c_size_t()
    ids = _u32_array(list(input_witnesses))
    flags = (1 if pedersen_unpinned else 0) | (0 if device_brillig else 2) | (0 if scaled_columns else 4)
    flags |= 0 if packed_hashes else 8
    flags |= 0 if spread_heavy else 16
    flags |= 0 if slack_scheduling else 32
    if ring_slots is not None:   # entries of the shared-memory ring of recent values (0 = none)
        flags |= (0xFFFF if ring_slots == 0 else ring_slots) << 8
    _check(lib().acvmb_plan_compile_host_ex(acir_bytes, len(acir_bytes), ids, len(input_witnesses), S, temp_pool, flags,
                                            C.byref(info), None, 0, C.byref(need)))
    buf = (C.c_uint8 * need.value)()
    _check(lib().acvmb_plan_compile_host_ex(acir_bytes, len(acir_bytes), ids, len(input_witnesses), S, temp_pool, flags,
                                            None, buf, need.value, C.byref(need)))
    return info.as_dict(), bytes(buf)


def permutation_route_host(outputs: Sequence[int]):
    """sorting::route(0..n-1, outputs) on the host (acvm/src/pwg/directives/sorting.rs:164-235): list of switch bits."""
    n = len(outputs)
    arr = _u32_array(list(outputs))
    nb = C.c_uint32()
    _check(lib().acvmb_permutation_route_host(arr, n, None, 0, C.byref(nb)))
    buf = (C.c_uint8 * max(1, nb.value))()
    _check(lib().acvmb_permutation_route_host(arr, n, buf, nb.value, C.byref(nb)))
    return [bool(b) for b in buf[:nb.value]]
