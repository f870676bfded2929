"""ctypes binding of libacvm_b200.so (the C ABI declared in include/acvm_b200.h).

The library is the product; this module only loads it.  If the shared object is missing the import
fails loudly -- there is no Python / CPU implementation to fall back to.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libacvm_b200.so")


class Status(C.Structure):
    _fields_ = [("code", C.c_uint32), ("err_kind", C.c_uint32), ("opcode_index", C.c_uint32), ("aux", C.c_uint32)]


class RunInfo(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("scatter_ms", C.c_double), ("gather_ms", C.c_double),
                ("kernel_launches", C.c_uint64), ("T", C.c_uint32), ("S", C.c_uint32), ("n_tiles", C.c_uint32),
                ("threads_per_cta", C.c_uint32), ("resident_instances", C.c_uint32), ("n_subbatches", C.c_uint32)]


class PlanInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_opcodes", "n_micro_ops", "n_steps", "n_slots_filled", "n_gate_assign", "n_gate_check", "n_logic", "n_range",
        "n_hash", "n_curve", "ref_fr_mul", "ref_fr_inv", "dev_imad", "alg_bytes", "n_temps")] + [(n, C.c_uint32) for n in (
        "num_witnesses", "n_slots", "S", "needs_full_kernel", "static_fail_present", "static_fail_opcode",
        "static_fail_kind", "static_fail_aux", "n_segments", "n_host_segments", "n_brillig", "n_brillig_device")] + [("n_gate_one_reduction", C.c_uint64),
                                                                                  ("scaled_columns", C.c_uint32), ("ring_slots", C.c_uint32),
                                                                                  ("n_operand_reads", C.c_uint64), ("n_ring_reads", C.c_uint64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m acvm_b200.build` (nvcc, sm_100a). "
            "acvm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u8p, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    sig = {
        "acvmb_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "acvmb_ctx_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
        "acvmb_ctx_n_devices": (C.c_int, [vp]),
        "acvmb_ctx_broadcast_backend": (C.c_char_p, [vp]),
        "acvmb_ctx_destroy": (None, [vp]),
        "acvmb_last_error": (C.c_char_p, []),
        "acvmb_device_name": (C.c_int, [vp, C.c_char_p, C.c_size_t]),
        "acvmb_host_alloc": (vp, [C.c_size_t]),
        "acvmb_host_free": (None, [vp]),
        "acvmb_circuit_from_acir": (C.c_int, [vp, C.c_char_p, C.c_size_t, u32p, C.c_uint32, C.POINTER(vp)]),
        "acvmb_circuit_destroy": (None, [vp]),
        "acvmb_circuit_info": (C.c_int, [vp, C.POINTER(PlanInfo)]),
        "acvmb_circuit_assign_opcodes": (C.c_int, [vp, u32p, C.c_uint32]),
        "acvmb_circuit_serialize": (C.c_int, [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "acvmb_circuit_deserialize": (C.c_int, [vp, vp, C.c_size_t, C.POINTER(vp)]),
        "acvmb_solve_batch": (C.c_int, [vp, C.c_uint32, vp, u32p, C.c_uint32, vp, C.POINTER(Status)]),
        "acvmb_solve_batch_ex": (C.c_int, [vp, C.c_uint32, vp, u32p, C.c_uint32, vp, vp, C.POINTER(Status)]),
        "acvmb_batch_download_ex": (C.c_int, [vp, C.c_uint32, C.c_uint32, u32p, C.c_uint32, vp, vp]),
        "acvmb_last_run_info": (C.c_int, [vp, C.POINTER(RunInfo)]),
        "acvmb_batch_create": (C.c_int, [vp, C.c_uint32, C.POINTER(vp)]),
        "acvmb_batch_resize": (C.c_int, [vp, C.c_uint32]),
        "acvmb_batch_destroy": (None, [vp]),
        "acvmb_batch_upload": (C.c_int, [vp, vp]),
        "acvmb_batch_run": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "acvmb_batch_stage_inputs": (C.c_int, [vp, C.c_uint32, vp]),
        "acvmb_batch_run_staged": (C.c_int, [vp, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "acvmb_batch_status": (C.c_int, [vp, C.POINTER(Status)]),
        "acvmb_batch_download": (C.c_int, [vp, C.c_uint32, C.c_uint32, u32p, C.c_uint32, vp]),
        "acvmb_batch_checksum": (C.c_int, [vp, u64p]),
        "acvmb_vm_new": (C.c_int, [vp, C.c_char_p, C.c_size_t, u32p, C.c_char_p, C.c_uint32, C.POINTER(vp)]),
        "acvmb_vm_destroy": (None, [vp]),
        "acvmb_vm_solve": (C.c_int, [vp, C.POINTER(Status)]),
        "acvmb_vm_solve_opcode": (C.c_int, [vp, C.POINTER(Status)]),
        "acvmb_vm_status": (C.c_int, [vp, C.POINTER(Status)]),
        "acvmb_vm_instruction_pointer": (C.c_int, [vp, u32p]),
        "acvmb_vm_num_witnesses": (C.c_int, [vp, u32p]),
        "acvmb_vm_witness": (C.c_int, [vp, C.c_uint32, C.c_char_p, C.POINTER(C.c_int)]),
        "acvmb_vm_finalize": (C.c_int, [vp, vp, vp, C.c_uint32]),
        "acvmb_vm_pending_foreign_call": (C.c_int, [vp, C.c_char_p, C.c_size_t, u32p, u32p, C.c_uint32, vp, C.c_uint32, u32p]),
        "acvmb_vm_resolve_foreign_call": (C.c_int, [vp, C.c_uint32, u32p, C.c_char_p]),
        "acvmb_witness_map_compress": (C.c_int, [u32p, C.c_char_p, C.c_uint32, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "acvmb_witness_map_decompress": (C.c_int, [C.c_char_p, C.c_size_t, u32p, vp, C.c_uint32, u32p]),
        "acvmb_fixed_base_scalar_mul": (C.c_int, [vp, C.c_char_p, C.c_char_p, C.c_uint32, vp, C.POINTER(Status)]),
        "acvmb_pedersen": (C.c_int, [vp, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.POINTER(Status)]),
        "acvmb_sha256": (C.c_int, [vp, C.c_char_p, C.c_uint32, C.c_uint32, vp]),
        "acvmb_keccak256": (C.c_int, [vp, C.c_char_p, C.c_uint32, C.c_uint32, vp]),
        "acvmb_ecdsa_secp256k1_verify": (C.c_int, [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, vp, C.POINTER(Status)]),
        "acvmb_ecdsa_secp256r1_verify": (C.c_int, [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, vp, C.POINTER(Status)]),
        "acvmb_plan_compile_host": (C.c_int, [C.c_char_p, C.c_size_t, u32p, C.c_uint32, C.c_uint32, C.POINTER(PlanInfo), vp,
                                              C.c_size_t, C.POINTER(C.c_size_t)]),
        "acvmb_plan_compile_host_ex": (C.c_int, [C.c_char_p, C.c_size_t, u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                 C.POINTER(PlanInfo), vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "acvmb_d2h_microbench": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32, C.POINTER(C.c_double)]),
        "acvmb_imad_cc_microbench": (C.c_int, [vp, C.POINTER(C.c_double)]),
        "acvmb_pedersen_generator_host": (C.c_int, [C.c_uint32, C.c_char_p]),
        "acvmb_permutation_route_host": (C.c_int, [u32p, C.c_uint32, vp, C.c_uint32, u32p]),
        "acvmb_brillig_run_host": (C.c_int, [C.c_char_p, C.c_size_t, C.c_uint32, C.c_char_p, C.c_uint32, vp, C.c_uint32, u32p, u32p]),
        "acvmb_imad_microbench": (C.c_int, [vp] + [C.POINTER(C.c_double)] * 4),
        "acvmb_frmul_microbench": (C.c_int, [vp, C.POINTER(C.c_double)]),
        "acvmb_ctx_set_option": (C.c_int, [vp, C.c_char_p, C.c_uint64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    lib._acvmb_signatures = sig
    return lib
