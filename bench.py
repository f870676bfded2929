#!/usr/bin/env python3
"""bench.py -- witnesses solved / second on the BASELINE.json headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gates G] [--batch B]

Workload (config.workload): BASELINE.json configs[1] -- 2^20 arithmetic-only width-3 PLONK gates over
BN254 Fr ("local" operands, "dense" coefficients, every 16th opcode an all-known check), batch 8192
per GPU, synthetic splitmix64 inputs, serialised in the reference's own ACIR wire format.

One "step" = one pass of the hot path over the whole batch (all sub-batches).
  value : witnesses/s with the inputs resident in HBM; per step = status reset + input scatter +
          step-VM kernel for every sub-batch; timed with CUDA events on the library's stream.
  e2e   : the same metric through the C ABI with HOST buffers (acvmb_solve_batch): H2D of the
          inputs, solve, gather + D2H of the FULL dense witness map of every instance (what
          ACVM::finalize returns), wall-clock around the synchronous calls.
N > 1 : one process per GPU (torchrun), plan compiled on rank 0 and broadcast once over NCCL, batch
        sharded 8192 instances per GPU, no traffic during the solve ("scaling": "weak").
--impl reference : the CPU reference-algorithm restatement (oracle/ref_solver.cpp; the Rust reference
        cannot be built here) on all host cores, same circuit bytes, same JSON line.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_START = time.time()
METRIC = "witnesses_per_s_2^20_gate_bn254_acir"
# DRAM bytes per gate-instance of vm_kernel from the committed `ncu --set full` capture of the round-2 kernel AT FULL SIZE
# (profiles/r2_vm_kernel_ncu_full_v6.txt: dram read 28.97 GB + write 149.09 GB for one launch of 4736 instances x 1 049 616
# micro-gates): the write is the 32 B result, operand reads hit L2 in "local" mode.
NCU_DRAM_BYTES_PER_GATE_INSTANCE = (28.973338e9 + 149.086287e9) / (4736 * 1049616)
UNIT = "witnesses/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cached_circuit(gates, mode, coeffs):
    from acvm_b200 import acir_builder as ab
    cache = os.path.join("/tmp", f"acvmb_circuit_{gates}_{mode}_{coeffs}.bin")
    if os.path.exists(cache):
        with open(cache, "rb") as f:
            data = f.read()
        return data, list(range(ab.N_INPUTS))
    t = time.time()
    data, inputs, _ = ab.synthetic_arith_circuit(gates, seed_id=1, mode=mode, coeffs=coeffs)
    log(f"[bench] generated {gates}-gate circuit ({len(data) / 1e6:.1f} MB gz) in {time.time() - t:.1f}s")
    try:
        with open(cache, "wb") as f:
            f.write(data)
    except OSError:
        pass
    return data, inputs


_DECODED = {}


def cpu_reference_rate(data, inputs, gates, threads, n_inst=None, label="cpu_baseline", optimized=False):
    """Times the C++ reference-algorithm restatement: `n_inst` full solves spread over `threads` host threads.
    optimized=True times the "optimised CPU" variant instead (dense witness vector, inverses hoisted to plan time)."""
    from acvm_b200 import acir_builder as ab
    from oracle import acir as oacir, cref
    cref.build()
    t = time.time()
    key = (id(data), len(data))
    if _DECODED.get("key") != key:      # the two CPU comparators of one run share the decoded circuit (16 s for 2^20 gates)
        circ = oacir.decode_circuit(data)
        _DECODED.update(key=key, circ=circ, packed=cref.pack_circuit(circ))
    circ, packed = _DECODED["circ"], _DECODED["packed"]
    log(f"[{label}] oracle decode+pack {time.time() - t:.1f}s")
    n_inst = n_inst or threads
    inp = ab.synthetic_inputs(n_inst, seed_id=1)
    nw = circ.current_witness_index + 1

    def run():
        t0 = time.perf_counter()
        if optimized:
            res, _ = cref.solve_batch_optimized(circ, inputs, inp, n_inst, nw, threads=threads, packed=packed)
        else:
            res, _, _ = cref.solve_batch(circ, inputs, inp, n_inst, nw, threads=threads, packed=packed)
        dt = time.perf_counter() - t0
        assert (res[:, 0] == 0).all(), "cpu reference failed to solve the synthetic circuit"
        return dt

    return run, n_inst


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    data, inputs = cached_circuit(args.gates, args.mode, args.coeffs)
    run, n_inst = cpu_reference_rate(data, inputs, args.gates, threads, label="reference")
    for _ in range(args.warmup):
        run()
    times = [run() for _ in range(args.steps)]
    total = sum(times)
    value = n_inst * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u256 (4x64 Montgomery, BN254 Fr)", "data": "synthetic",
        "config": workload_config(args, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_inst} full solves of the {args.gates}-gate circuit per step, one solver instance per thread "
                                   "(oracle/ref_solver.cpp: std::map witness map, per-gate evaluate + inversion)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Rust acvm 0.27.0 cannot be built in this image (no cargo/rustc); this is the reference-algorithm restatement",
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs 2-4 at their stated sizes, reported in the `secondary` object of the JSON line.
# One timed pass each (they take 0.2-35 s per pass; the kernels are latency bound at these batch sizes), after a one-tile
# warm-up launch of the same kernel that absorbs module load and the lookup-table upload.
# ---------------------------------------------------------------------------------------------------------------------
PEDERSEN_ALG_FR_MUL = 57 * 3 * 11 + 3 * 380     # 57(k+1) mixed additions x 11 Fr-mul + (k+1) Fermat inversions, k = 2


def _byte_inputs(batch, n_inputs, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    out = np.zeros((batch, n_inputs, 32), dtype=np.uint8)
    out[:, :, 31] = rng.integers(0, 256, size=(batch, n_inputs), dtype=np.uint8)
    return out.tobytes()


def _field_inputs(batch, n_inputs, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 256, size=(batch, n_inputs, 32), dtype=np.uint8)
    out[:, :, 0] &= 0x1F   # < 2^253 < p
    return out.tobytes()


def secondary_circuit(which, scale):
    """(data, inputs, n_witnesses, units per instance, description, batch per GPU) of a BASELINE config at 1/scale size."""
    from acvm_b200 import acir_builder as ab
    if which == "config2":
        n = (1 << 16) // scale
        data, inputs, nw = ab.pedersen_chain_circuit(n)
        return data, inputs, nw, n, f"configs[2]: {n} chained Pedersen{{[prev.x, fresh_i], domain 0}} calls on Grumpkin, batch 4096", 4096
    if which == "config3":
        n = (1 << 14) // scale
        data, inputs, nw = ab.hash_chain_circuit(n)
        return data, inputs, nw, n, f"configs[3]: {n} chained hash calls alternating SHA256 / Keccak256 over 64 byte-witnesses, batch 4096", 4096
    n = (1 << 20) // scale
    data, inputs, nw, counts = ab.mixed_circuit(n)
    return data, inputs, nw, 1, (f"configs[4]: {n} mixed opcodes (93% dense arithmetic, 4% RANGE/AND/XOR(32), 2% SHA256/Keccak256(64 B), "
                                 f"1% Pedersen(2)/FixedBaseScalarMul), batch 8192 per GPU; counts {counts}"), 8192


def run_secondary(ctx, which, scale, hbm_peak, imad_peak, first_instance=0, cpu_threads=0):
    import acvm_b200
    from acvm_b200 import acir_builder as ab
    t0 = time.time()
    data, inputs, nw, units, desc, batch = secondary_circuit(which, scale)
    t_gen = time.time() - t0
    t0 = time.time()
    circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
    t_plan = time.time() - t0
    info = circ.info
    if which == "config3":
        make_inputs = lambda n, first: _byte_inputs(n, len(inputs), 3 + first)
    elif which == "config2":
        make_inputs = lambda n, first: _field_inputs(n, len(inputs), 2 + first)
    else:
        make_inputs = lambda n, first: ab.synthetic_inputs(n, seed_id=4, first_instance=first)
    import torch
    free_b, _ = torch.cuda.mem_get_info()
    per_inst = info["n_slots"] * 32 + 8 + len(inputs) * 32
    T = 32 if info["n_curve"] else max(1, 128 // info["S"])
    fit = max(T, (int(free_b * 0.92) // per_inst) // T * T)
    # equal sub-batches: these kernels are latency bound (a pass costs the same for 400 or 2700 instances), so a small
    # remainder pass would cost as much as a full one
    n_pass = -(-batch // fit)
    per = min(fit, -(-(-(-batch // n_pass)) // T) * T)
    sizes, left = [], batch
    while left > 0:
        take = min(left, per)
        sizes.append(take)
        left -= take
    b = acvm_b200.DeviceBatch(circ, max(sizes))
    off = 0
    for k, sz in enumerate(sizes):
        b.resize(sz)
        b.stage_inputs(k, make_inputs(sz, first_instance + off))
        off += sz
    b.resize(min(T, sizes[0]))          # warm-up: one tile through the same kernel
    b.stage_inputs(len(sizes), make_inputs(min(T, sizes[0]), first_instance))
    b.run_staged(len(sizes))
    tot = vm = 0.0
    solved = 0
    for k, sz in enumerate(sizes):
        b.resize(sz)
        t, v = b.run_staged(k)
        tot += t
        vm += v
        solved += sum(1 for s_ in b.status() if s_.status == "Solved")
    b.close()
    assert solved == batch, f"{which}: {batch - solved} instances did not solve"
    sec = tot * 1e-3
    line = {"workload": desc, "batch_per_gpu": batch, "sub_batches": sizes, "ms_per_step": tot, "steps": 1, "warmup": "one tile",
            "witnesses_per_s": batch / sec, "n_micro_ops": info["n_micro_ops"], "n_steps": info["n_steps"], "S": info["S"], "T": T,
            "plan_compile_s": round(t_plan, 1), "circuit_gen_s": round(t_gen, 1)}
    vm_launch_s = vm * 1e-3 / len(sizes)
    inst_per_launch = batch / len(sizes)
    if which == "config2":
        line.update(metric="pedersen_hashes_per_s", value=units * batch / sec, unit="hashes/s",
                    parity="UNPINNED vs barretenberg (opt-in kernel, structure-identical; DESIGN.md section 6)")
        alg = PEDERSEN_ALG_FR_MUL * 136 * units * inst_per_launch / vm_launch_s
        line["roofline"] = {"bound": "imad", "achieved": alg / 1e12, "peak": imad_peak / 1e12, "unit": "T IMAD/s", "frac": alg / imad_peak,
                            "traffic": None, "algorithmic": f"{PEDERSEN_ALG_FR_MUL} Fr-mul x 136 IMAD per Pedersen(2) call (SURVEY 8d)",
                            "kernel": "vm_kernel<32,8,FULL>"}
    elif which == "config3":
        line.update(metric="hashes_per_s", value=units * batch / sec, unit="hashes/s")
        alg = 3072.0 * units * inst_per_launch / vm_launch_s / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": alg, "peak": hbm_peak, "unit": "GB/s", "frac": alg / hbm_peak, "traffic": None,
                            "algorithmic": "3 KiB per hash call-instance: 64 x 32 B in + 32 x 32 B out (SURVEY 8d)", "kernel": "vm_kernel<8,16,FULL>"}
    else:
        line.update(metric="witnesses_per_s", value=batch / sec, unit="witnesses/s")
        alg = info["alg_bytes"] * inst_per_launch / vm_launch_s / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": alg, "peak": hbm_peak, "unit": "GB/s", "frac": alg / hbm_peak, "traffic": None,
                            "algorithmic": f"{info['alg_bytes']} B per instance (plan.stats.alg_bytes: 32 B per operand read / witness written)",
                            "kernel": "vm_kernel<32,8,FULL>"}
    if cpu_threads:
        # CPU comparator: the C++ reference-algorithm restatement on a bounded sample (one instance per host thread)
        from oracle import acir as oacir, cref
        cref.build()
        cpu_scale = {"config2": 64, "config3": 1, "config4": 8}[which]
        cdata, cinputs, cnw, cunits, _, _ = secondary_circuit(which, scale * cpu_scale) if cpu_scale > 1 else (data, inputs, nw, units, None, None)
        oc = oacir.decode_circuit(cdata)
        packed = cref.pack_circuit(oc)
        cinp = make_inputs(cpu_threads, 0)
        t0 = time.perf_counter()
        res, _, _ = cref.solve_batch(oc, cinputs, cinp, cpu_threads, cnw, threads=cpu_threads, packed=packed)
        dt = time.perf_counter() - t0
        assert (res[:, 0] == 0).all()
        per_s = cpu_threads * (cunits if which != "config4" else 1.0 / cpu_scale) / dt
        line["cpu_baseline"] = {"value": per_s, "unit": line["unit"], "cores": cpu_threads, "kind": "port",
                                "sample": f"{cpu_threads} instances (one per host thread) of the same circuit at 1/{cpu_scale} length "
                                          f"({dt:.1f}s; oracle/ref_solver.cpp); rate scaled by length for config 4"}
    circ.close()
    return line


def multi_device_c_abi_line(n_dev):
    """N > 1: the same sharded solve through ONE multi-device context of the C ABI (acvmb_ctx_create_multi) -- one process, one
    host thread per GPU inside the library, the plan broadcast once (NCCL or peer copies) -- on a 2^16-gate circuit with 8192
    instances per GPU, only the last 32 witnesses of every instance brought back.  Rank 0 runs it after the other ranks are done."""
    import acvm_b200
    from acvm_b200 import acir_builder as ab
    data, inputs, _ = ab.synthetic_arith_circuit(1 << 16, seed_id=1)
    ctx = acvm_b200.Context(list(range(n_dev)))
    try:
        t0 = time.time()
        circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
        t_circ = time.time() - t0
        batch = 8192 * n_dev
        inp = ab.synthetic_inputs(256, seed_id=1) * (batch // 256)
        tail = list(range(circ.num_witnesses - 32, circ.num_witnesses))
        out, st = circ.solve_batch(inp, batch, out_ids=tail)          # warm-up (allocations, module load on every device)
        t0 = time.perf_counter()
        out, st = circ.solve_batch(inp, batch, out_ids=tail)
        dt = time.perf_counter() - t0
        assert all(s_.status == "Solved" for s_ in st)
        assert out[:32 * 32] == out[256 * 32 * 32:257 * 32 * 32]      # instance 256 repeats instance 0: same witnesses, other shard
        ri = circ.run_info()
        return {"devices": ctx.n_devices(), "broadcast": ctx.broadcast_backend(), "gates": 1 << 16, "batch": batch,
                "witnesses_per_s": batch / dt, "wall_ms": 1e3 * dt, "max_device_kernel_ms": ri["kernel_ms"],
                "kernel_launches": ri["kernel_launches"], "circuit_create_and_broadcast_s": round(t_circ, 2)}
    finally:
        ctx.close()


def workload_config(args, extra):
    cfg = {"workload": f"configs[1]: {args.gates} arithmetic-only width-3 PLONK gates (BN254 Fr), batch {args.batch} per GPU, "
                       f"operands={args.mode}, coefficients={args.coeffs}, every 16th opcode an all-known check",
           "gates": args.gates, "batch_per_gpu": args.batch, "operand_mode": args.mode, "coefficients": args.coeffs,
           "l2_policy": "inputs larger than L2: witness columns of one sub-batch are >> 126 MB and are rewritten every step"}
    if extra:
        cfg.update(extra)
    return cfg


def bind_to_gpu_numa_node(gpu_index):
    """Multi-rank runs: pin the process to the CPUs NVML reports as local to its GPU, so the pinned host buffers of the
    end-to-end path are first-touched on the NUMA node next to the GPU's PCIe root (8 ranks x 50 GB/s of D2H otherwise
    cross the socket interconnect)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            log(f"[bench] gpu {gpu_index}: bound to {len(cpus)} local CPUs")
    except Exception as e:  # best effort only
        log(f"[bench] gpu {gpu_index}: NUMA binding skipped ({e})")


def run_ours(args):
    import acvm_b200
    from acvm_b200 import acir_builder as ab

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
    ctx = acvm_b200.Context(local_rank)
    if args.S:
        ctx.set_option("S", args.S)
    if args.T:
        ctx.set_option("T", args.T)
    for kv in filter(None, os.environ.get("ACVMB_OPTS", "").split(",")):   # A/B hook: context options, e.g. scaled_columns=0,ring_bytes=0
        k, v = kv.split("=")
        ctx.set_option(k, int(v))

    # ---- circuit: compiled on rank 0, broadcast once (the only collective on this path) ----
    data, inputs = (None, list(range(ab.N_INPUTS)))
    if rank == 0:
        data, inputs = cached_circuit(args.gates, args.mode, args.coeffs)
        t0 = time.time()   # decode + plan compile + upload only (the circuit generation above is the harness's)
        circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
        log(f"[bench] plan compiled in {time.time() - t0:.1f}s: {circ.info}")
    if world > 1:
        from acvm_b200.dist import broadcast_circuit
        circ = broadcast_circuit(ctx, circ if rank == 0 else None, inputs, src=0)
    info = circ.info
    nw = circ.num_witnesses

    # ---- sub-batching: witness columns of the whole batch do not fit in HBM ----
    import torch
    free_b, total_b = torch.cuda.mem_get_info(local_rank)
    per_inst = info["n_slots"] * 32 + 8 + len(inputs) * 32
    T = args.T or max(1, 128 // info["S"])
    budget = int(free_b * 0.88)
    fit = max(T, (budget // per_inst) // T * T)
    # sub-batch sizes: whole waves of CTAs (148 SMs x k CTAs x T instances) so no SM idles while another holds an
    # extra CTA; the remainder goes last
    wave = 148 * T
    sizes, left = [], args.batch
    while left > 0:
        take = min(left, fit)
        if take >= wave and left > take:
            take = take // wave * wave
        sizes.append(take)
        left -= take
    log(f"[bench] rank {rank}: free {free_b / 2**30:.1f} GiB, {per_inst / 2**20:.1f} MiB/instance -> sub-batches {sizes}")
    first_inst = rank * args.batch
    # the sub-batches share ONE column buffer in time (capacity = the largest); resize() selects the active count
    batch_obj = acvm_b200.DeviceBatch(circ, max(sizes))
    off = 0
    for k, sz in enumerate(sizes):
        batch_obj.resize(sz)
        batch_obj.stage_inputs(k, ab.synthetic_inputs(sz, seed_id=1, first_instance=first_inst + off))
        off += sz

    def step():
        tot = vm = 0.0
        for k in range(len(sizes)):
            batch_obj.resize(sizes[k])
            t, v = batch_obj.run_staged(k)
            tot += t
            vm += v
        return tot, vm

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    # correctness gate on the last warm-up pass: every instance solved
    st = batch_obj.status()
    assert len(st) == sizes[-1] and all(s.status == "Solved" for s in st), "synthetic circuit must solve for every instance"

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    dev_ms = vm_ms = 0.0
    for _ in range(args.steps):
        t, v = step()
        dev_ms += t
        vm_ms += v
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches_per_step = 3 * len(sizes)

    # ---- end-to-end through acvmb_solve_batch with host buffers (full dense witness map back) ----
    e2e = None
    if not args.no_e2e:
        import psutil
        avail = psutil.virtual_memory().available
        e2e_chunk = int(min(args.batch, max(1, min(args.e2e_chunk_gib * 2**30, avail * 0.4 / world) // (nw * 32))))
        n_calls = -(-args.batch // e2e_chunk)
        out_bytes = e2e_chunk * nw * 32
        batch_obj.close()
        lib = acvm_b200.lib()
        t0 = time.time()
        # the pinned pages are placed where the allocating thread runs: sit next to the GPU's PCIe root for the D2H path,
        # and give the cores back before the CPU baselines are timed
        saved_affinity = os.sched_getaffinity(0)
        if world == 1:
            bind_to_gpu_numa_node(local_rank)
        host_out = lib.acvmb_host_alloc(out_bytes)
        if not host_out:
            raise RuntimeError("pinned host allocation failed")
        log(f"[bench] e2e: {n_calls} calls of {e2e_chunk} instances, pinned {out_bytes / 2**30:.1f} GiB in {time.time() - t0:.1f}s")
        ins = []
        o = 0
        while o < args.batch:
            c = min(e2e_chunk, args.batch - o)
            ins.append((c, (C.c_uint8 * (c * len(inputs) * 32)).from_buffer_copy(ab.synthetic_inputs(c, seed_id=1, first_instance=first_inst + o))))
            o += c
        st_arr = (acvm_b200._lib.Status * e2e_chunk)()

        def e2e_step():
            for c, buf in ins:
                rc = lib.acvmb_solve_batch(circ._h, c, buf, None, 0, C.c_void_p(host_out), st_arr)
                if rc:
                    raise RuntimeError(lib.acvmb_last_error().decode())

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        e2e_step()  # warm-up (also faults in the pinned pages)
        # bare D2H rate of THIS box into the same pinned buffer, all ranks copying at once: the ceiling of the number below
        barrier()
        gbps = C.c_double()
        rc = lib.acvmb_d2h_microbench(ctx._h, C.c_void_p(host_out), min(out_bytes, 4 << 30), 4, C.byref(gbps))
        d2h_ceiling = gbps.value if rc == 0 else None
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_wall = (time.perf_counter() - w0) / e2e_steps
        # spot-check the last chunk against the kernel statuses
        assert all(st_arr[i].code == 0 for i in range(ins[-1][0]))
        lib.acvmb_host_free(C.c_void_p(host_out))
        os.sched_setaffinity(0, saved_affinity)
        ri = circ.run_info()
        e2e = {"wall_s_per_step": e2e_wall, "h2d": args.batch * len(inputs) * 32, "d2h": args.batch * (nw * 32 + 16),
               "calls_per_step": n_calls, "instances_per_call": e2e_chunk, "pieces_per_call": ri["n_subbatches"],
               "d2h_ceiling": d2h_ceiling}

    hbm_peak, peak_src = measured_peaks()
    imad = ctx.imad_microbench()
    # ---- BASELINE configs[4] as stated: 65536 instances of the 2^20-op mixed circuit sharded 8192 per GPU over 8 GPUs ----
    sec4 = None
    if world == 8 and args.secondary != "none":
        if e2e is None:
            batch_obj.close()
        ctx.set_option("pedersen_unpinned", 1)
        ctx.set_option("cache_batch", 0)   # release the column buffers the end-to-end calls left cached in the context
        ctx.set_option("cache_batch", 1)
        try:
            sec4 = run_secondary(ctx, "config4", 1 if args.secondary == "full" else 64, hbm_peak, imad["imad_wide_per_s"],
                                 first_instance=rank * 8192)
        except Exception as e_:
            sec4 = {"error": f"{type(e_).__name__}: {e_}", "ms_per_step": 0.0}
    # ---- reduce over ranks (max time) ----
    vals = [dev_ms, vm_ms, wall, e2e["wall_s_per_step"] if e2e else 0.0, sec4["ms_per_step"] if sec4 else 0.0,
            -(e2e["d2h_ceiling"] or 0.0) if e2e else 0.0]
    if dist is not None:
        t = torch.tensor(vals, dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = t.tolist()
    dev_ms, vm_ms, wall, e2e_wall, sec4_ms, neg_ceiling = vals
    d2h_ceiling_min = -neg_ceiling   # the slowest rank's bare rate (all ranks copied concurrently)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    total_inst = args.batch * world
    ms_per_step = dev_ms / args.steps
    value = total_inst / (ms_per_step * 1e-3)
    vm_launch_ms = vm_ms / (args.steps * len(sizes))          # average step-VM kernel launch
    inst_per_launch = sum(sizes) / len(sizes)
    alg_bytes_launch = info["alg_bytes"] * inst_per_launch
    achieved_gbs = alg_bytes_launch / (vm_launch_ms * 1e-3) / 1e9
    imad_achieved = info["dev_imad"] * inst_per_launch / (vm_launch_ms * 1e-3)
    launch_s = vm_launch_ms * 1e-3
    hbm_frac = achieved_gbs / hbm_peak
    imad_frac = imad_achieved / imad["imad_wide_per_s"]
    traffic = NCU_DRAM_BYTES_PER_GATE_INSTANCE * info["n_micro_ops"] * inst_per_launch
    hbm_obj = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_frac, "traffic": traffic,
               "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_launch,
               "algorithmic": "96 B per gate-instance: two 32 B operands + the 32 B result (SURVEY 8d; plan.stats.alg_bytes is the exact sum)"}
    imad_obj = {"bound": "imad", "achieved": imad_achieved / 1e12, "peak": imad["imad_wide_per_s"] / 1e12, "unit": "T IMAD/s",
                "frac": imad_frac, "traffic": traffic,
                "peak_source": "measured in this run: independent mad.wide.u32 chains on all SMs (profiles/IMAD_PEAKS.json holds the "
                               "committed copy)",
                "peak_imad32_per_s": imad["imad32_per_s"], "peak_wide_carry_per_s": imad["imad_wide_carry_per_s"],
                "frac_of_carry_peak": imad_achieved / imad["imad_wide_carry_per_s"],
                "imad_per_instance": info["dev_imad"], "executed": "32x32 multiply-accumulates the kernel executes (plan.stats.dev_imad)"}
    # the roofline that binds is the one the kernel sits closer to; the other one is reported beside it
    roof, other = (hbm_obj, imad_obj) if hbm_frac >= imad_frac else (imad_obj, hbm_obj)
    roof = dict(roof)
    roof.update({"kernel": "vm_kernel", "kernel_ms_per_launch": vm_launch_ms,
                 "traffic_source": "ncu --set full capture of this kernel at full size (profiles/r2_vm_kernel_ncu_full_v6.txt), "
                                   "35.8 B of DRAM traffic per gate-instance, scaled by gate-instances per launch",
                 "other": other,
                 "fr_mul": {"reference_fr_mul_per_instance": info["ref_fr_mul"], "reference_fr_inv_per_instance": info["ref_fr_inv"],
                            "algorithmic_fr_mul_per_s": info["ref_fr_mul"] * inst_per_launch / launch_s,
                            "device_montgomery_reductions_per_instance": info["n_gate_one_reduction"],
                            "note": "the reference performs 5 Fr-mul + 1 inversion per gate; scaled columns leave ONE Montgomery "
                                    "reduction (136 IMAD) per gate on the device"}})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (8x32-bit limbs, Montgomery, BN254 Fr)", "data": "synthetic",
        "config": workload_config(args, {"sub_batches": sizes, "T": T, "S": info["S"], "n_steps": info["n_steps"],
                                         "slot_fill": info["n_slots_filled"] / max(1, info["n_steps"] * info["S"]),
                                         "resident_warps_per_sm": [round(-(-sz // T) * (T * info["S"] // 32) / 148, 2) for sz in sizes],
                                         "scaled_columns": bool(info["scaled_columns"]),
                                         "d2h_overlapped_with_solve": bool(e2e and e2e.get("pieces_per_call", 1) > 1)}),
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "roofline": roof,
    }
    try:   # committed copy of the measured integer-multiply peaks (SURVEY 8d asks for them next to MEASURED_PEAKS.json)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "IMAD_PEAKS.json"), "w") as f:
            json.dump({"imad32_per_s": imad["imad32_per_s"], "imad_wide_per_s": imad["imad_wide_per_s"],
                       "imad_wide_carry_per_s": imad["imad_wide_carry_per_s"], "fr_mul_per_s": imad["fr_mul_per_s"],
                       "sm_clock_mhz": imad["sm_clock_mhz"], "how": "acvmb_imad_microbench / acvmb_frmul_microbench (vm_kernel.cu): "
                       "independent chains on 148 SMs x 8 CTAs x 256 threads, best of 3 after a warm-up"}, f, indent=1)
    except OSError:
        pass
    if e2e:
        line["e2e"] = {"value": total_inst / e2e_wall, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                       "calls_per_step": e2e["calls_per_step"], "instances_per_call": e2e["instances_per_call"],
                       "pieces_per_call": e2e["pieces_per_call"],
                       "d2h_ceiling_GBps_per_gpu": d2h_ceiling_min,
                       "d2h_achieved_GBps_per_gpu": e2e["d2h"] / e2e_wall / 1e9,
                       "frac_of_d2h_ceiling": (e2e["d2h"] / e2e_wall / 1e9) / d2h_ceiling_min if d2h_ceiling_min else None,
                       "d2h_ceiling_how": "cudaMemcpyAsync of 4 GiB x 4 from HBM into the same pinned buffer, every rank at the same "
                                          "time (acvmb_d2h_microbench); the slowest rank is reported",
                       "pipeline": "every call is cut into pieces: the VM kernel of piece k+1 runs while piece k drains (gather + D2H)",
                       "output": "full dense witness map of every instance (ACVM::finalize), pinned host buffer"}
    # ---- cpu baseline beside it (rank 0, N=1 only) ----
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        run, n_inst = cpu_reference_rate(data, inputs, args.gates, threads)
        dt = run()
        line["cpu_baseline"] = {"value": n_inst / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n_inst} full solves of the same {args.gates}-gate circuit, one per host thread "
                                          f"({dt:.1f}s; oracle/ref_solver.cpp reference-algorithm restatement)"}
        # the "optimised CPU" comparator SURVEY 8(d) asks for beside it: same results, none of the reference's map /
        # allocation / per-gate inversion overhead
        run_o, n_o = cpu_reference_rate(data, inputs, args.gates, threads, n_inst=threads * 8, label="cpu_optimized", optimized=True)
        dt_o = run_o()
        line["cpu_baseline_optimized"] = {"value": n_o / dt_o, "unit": UNIT, "cores": threads, "kind": "port",
                                          "sample": f"{n_o} full solves, dense witness vector + plan-time inverses + 4x64 Montgomery "
                                                    f"({dt_o:.1f}s incl. the one-time plan; oracle/ref_solver.cpp ref_solve_batch_optimized)"}
    if sec4 is not None and "error" not in sec4:
        sec4.update(n_gpus=world, batch_total=8192 * world, ms_per_step=sec4_ms, value=8192 * world / (sec4_ms * 1e-3),
                    witnesses_per_s=8192 * world / (sec4_ms * 1e-3), note="max over ranks of the per-GPU time; roofline object is rank 0's")
    if sec4 is not None:
        line["secondary"] = {"config4": sec4}
    if world > 1:
        try:
            batch_obj.close() if args.no_e2e else None
            circ.close()
            ctx.close()
            if dist is not None:
                dist.destroy_process_group()
                dist = None
            time.sleep(2.0)   # the other ranks are exiting: let their device memory go
            line["multi_gpu_c_abi"] = multi_device_c_abi_line(world)
        except Exception as e:
            line["multi_gpu_c_abi"] = {"error": f"{type(e).__name__}: {e}"}
    # ---- BASELINE configs 2-4 at their stated sizes (rank 0 at N = 1) ----
    if world == 1 and args.secondary != "none":
        scale = 1 if args.secondary == "full" else 64
        ctx.set_option("pedersen_unpinned", 1)   # configs 2 and 4 contain Pedersen calls: measured on the opt-in kernel, stated in the line
        line["secondary"] = {}
        threads = 0 if args.no_cpu_baseline else (os.cpu_count() or 1)
        batch_obj.close()
        ctx.set_option("cache_batch", 0)   # release the column buffers the end-to-end calls left cached in the context (tens of GB)
        ctx.set_option("cache_batch", 1)
        for which in ("config3", "config4", "config2"):
            t0 = time.time()
            if time.time() - T_START > args.time_budget_s:   # the headline line must come out: skip what no longer fits
                line["secondary"][which] = {"skipped": f"time budget of {args.time_budget_s} s used up before this configuration"}
                continue
            try:
                line["secondary"][which] = run_secondary(ctx, which, scale, hbm_peak, imad["imad_wide_per_s"], cpu_threads=threads)
            except Exception as e:   # a secondary measurement must never take the headline line down
                line["secondary"][which] = {"error": f"{type(e).__name__}: {e}"}
            log(f"[bench] secondary {which}: {time.time() - t0:.1f}s {line['secondary'][which].get('value')}")
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def _protect_stdout():
    """The driver reads ONE JSON line from stdout.  Native libraries (NCCL prints its version banner there) must not be able
    to add lines: fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real = os.fdopen(saved, "w")
    return real


_REAL_STDOUT = None


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gates", type=int, default=1 << 20)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--mode", default="local", choices=["local", "global"])
    ap.add_argument("--coeffs", default="dense", choices=["dense", "noir-like"])
    ap.add_argument("--S", type=int, default=0)
    ap.add_argument("--T", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--time-budget-s", type=float, default=420.0,
                    help="no further secondary configuration is started once the run is this old")
    ap.add_argument("--secondary", default="full", choices=["full", "quick", "none"],
                    help="BASELINE configs 2-4 in the `secondary` object: full = stated sizes, quick = 1/64 length, none = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-chunk-gib", type=float, default=64.0,
                    help="pinned host buffer per acvmb_solve_batch call; one call for the whole batch lets the library overlap sub-batches")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: fewer than 3 warm-up steps requested")
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
