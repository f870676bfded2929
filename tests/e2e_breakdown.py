"""Where the end-to-end time of one acvmb_solve_batch call goes (never a bench number): phases timed separately on the
headline circuit for one e2e chunk of instances."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acvm_b200
from acvm_b200 import acir_builder as ab

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2184
gates = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
ctx = acvm_b200.Context(0)
for kv in filter(None, os.environ.get("ACVMB_OPTS", "").split(",")):
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
data, inputs, nw = ab.synthetic_arith_circuit(gates, mode="local", coeffs="dense")
t = time.time(); circ = acvm_b200.CompiledCircuit(ctx, data, inputs); print("compile", round(time.time() - t, 2))
lib = acvm_b200.lib()
nw = circ.num_witnesses
inp = ab.synthetic_inputs(n, seed_id=1)
out_bytes = n * nw * 32
host = lib.acvmb_host_alloc(out_bytes)
st = (acvm_b200._lib.Status * n)()
buf = (C.c_uint8 * len(inp)).from_buffer_copy(inp)
def T(f):
    t = time.perf_counter(); r = f(); return time.perf_counter() - t, r
for rep in range(3):
    h = C.c_void_p()
    t_create, rc = T(lambda: lib.acvmb_batch_create(circ._h, n, C.byref(h))); assert rc == 0
    t_up, rc = T(lambda: lib.acvmb_batch_upload(h, buf)); assert rc == 0
    t_run, rc = T(lambda: lib.acvmb_batch_run(h, None)); assert rc == 0
    t_st, rc = T(lambda: lib.acvmb_batch_status(h, st)); assert rc == 0
    t_dl, rc = T(lambda: lib.acvmb_batch_download(h, 0, n, None, 0, C.c_void_p(host))); assert rc == 0
    t_destroy, _ = T(lambda: lib.acvmb_batch_destroy(h))
    t_all, rc = T(lambda: lib.acvmb_solve_batch(circ._h, n, buf, None, 0, C.c_void_p(host), st)); assert rc == 0
    print(f"rep {rep}: create {t_create*1e3:.1f} upload {t_up*1e3:.1f} run {t_run*1e3:.1f} status {t_st*1e3:.1f} download {t_dl*1e3:.1f} "
          f"({out_bytes/t_dl/1e9:.1f} GB/s) destroy {t_destroy*1e3:.1f} | solve_batch {t_all*1e3:.1f} ms", flush=True)
