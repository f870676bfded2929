"""Test-only Python interpreter of a compiled plan blob (acvmb_circuit_serialize / acvmb_plan_compile_host).

It executes the device record stream step by step on Python ints with the SAME hazards the kernel
has (all slots of a step read the state left by earlier steps), so the CPU suite can check the C++
decoder + plan compiler + record semantics against the oracle without a GPU.  Never shipped.
"""
import struct

P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
RINV = pow(1 << 256, -1, P)
R_MONT = (1 << 256) % P
NONE = 0xFFFFFFFF

MK = dict(NOP=0, GATE_ASSIGN=1, GATE_CHECK=2, AND=3, XOR=4, RANGE=5, SHA256=6, KECCAK256=7, FIXED_BASE=8, PEDERSEN=9,
          GATE_GENERAL=10, COPY_CHECK=11, REQUIRE=12, COPY=13, TO_LE_RADIX=14, QUOTIENT=15, MEM_READ=16, MEM_WRITE=17,
          BLAKE2S=18, HASH_TO_FIELD=19, ECDSA=20, CURVE_PART=21, JAC_ADD=22, JAC_FINAL=23, INT_OP=24,
          HASH_PACK=25, HASH_CORE=26, HASH_UNPACK=27)
EK_OOB = 5
EK_PANIC = 8
GF_MUL, GF_Y, GF_NLIN_SHIFT, GF_W1_IS_X, GF_OUT_CHECK = 1, 2, 2, 16, 32
EK_MISSING, EK_TOO_MANY, EK_UNSAT, EK_BB_FAILED = 1, 2, 4, 6


class PlanBlob:
    def __init__(self, blob: bytes):
        o = 0
        magic, = struct.unpack_from("<Q", blob, o)
        o += 8
        assert magic == 0x3430304E414C5042, "bad magic"
        (self.S, self.num_witnesses, self.n_slots, self.n_opcodes, self.chunk_steps, self.needs_full, self.n_steps,
         self.sf_present, self.sf_opcode, self.sf_kind, self.sf_aux, self.n_mu, self.ring_slots, _pad) = struct.unpack_from("<14I", blob, o)
        o += 56
        self.stats = struct.unpack_from("<23Q", blob, o)   # PlanStats (plan.hpp), a POD of 23 u64
        o += 184
        o = (o + 15) // 16 * 16

        def vec(fmt, size):
            nonlocal o
            n, = struct.unpack_from("<Q", blob, o)
            o += 8
            data = blob[o:o + n * size]
            o += n * size
            o = (o + 15) // 16 * 16
            return n, data

        n, d = vec("I", 4)
        self.input_witnesses = list(struct.unpack(f"<{n}I", d))
        n, d = vec("I", 4)
        self.input_scaled = list(struct.unpack(f"<{n}I", d))
        n, d = vec("I", 4)
        limbs = struct.unpack(f"<{n}I", d)
        # scaled columns: per witness (1/lambda_w)*R; canonical value = stored * that / R
        self.unscale = [sum(limbs[8 * w + j] << (32 * j) for j in range(8)) for w in range(n // 8)]
        n, d = vec("I", 4)
        self.assign_opcode = list(struct.unpack(f"<{n}I", d))
        n, d = vec("I", 4)
        self.mu_index_of = list(struct.unpack(f"<{n}I", d))
        n, d = vec("I", 4)
        self.payload = list(struct.unpack(f"<{n}I", d))
        n, d = vec("seg", 16)
        self.segments = [struct.unpack_from("<4I", d, 16 * i) for i in range(n)]
        n, d = vec("I", 4)
        self.host_desc = list(struct.unpack(f"<{n}I", d))
        n, d = vec("B", 1)
        self.acir_gz = bytes(d)
        n, d = vec("rec", 192)
        self.n_records = n
        self.stream = d
        assert n == self.n_steps * self.S

    def record(self, i):
        w = struct.unpack_from("<48I", self.stream, i * 192)
        hdr = w[:8]
        coefs = [sum(w[8 + 8 * k + j] << (32 * j) for j in range(8)) for k in range(5)]
        return hdr, coefs


def default_hooks():
    """Heavy micro-ops evaluated with the oracle's primitives (test code may use the oracle)."""
    from oracle import grumpkin, hashes
    GF_OUT_CHECK_, GF_OUT2_CHECK_ = 32, 128

    def hash_hook(fn):
        def run(cols, hdr, payload, record_fail):
            opcode, off = hdr[1], hdr[7]
            n_in, mask, var_w = payload[off], payload[off + 1], payload[off + 2]
            ins = payload[off + 4: off + 4 + 2 * n_in]
            outs = payload[off + 4 + 2 * n_in: off + 4 + 2 * n_in + 32]
            msg = bytearray()
            for k in range(n_in):
                nbytes = min(32, (ins[2 * k + 1] + 7) // 8)
                msg += cols[ins[2 * k]].to_bytes(32, "little")[:nbytes]
            if var_w != NONE:
                take = cols[var_w] & ((1 << 128) - 1)
                if take > len(msg):
                    record_fail(opcode, EK_BB_FAILED, 11)
                    return []
                msg = msg[:take]
            d = fn(bytes(msg))
            w = []
            for i in range(32):
                if (mask >> i) & 1:
                    if cols[outs[i]] != d[i]:
                        record_fail(opcode, EK_UNSAT)
                        w.append((outs[i], d[i]))
                else:
                    w.append((outs[i], d[i]))
            return w
        return run

    def hash_to_field(cols, hdr, payload, record_fail):
        opcode, off = hdr[1], hdr[7]
        n_in, mask = payload[off], payload[off + 1]
        ins = payload[off + 4: off + 4 + 2 * n_in]
        out = payload[off + 4 + 2 * n_in]
        msg = bytearray()
        for k in range(n_in):
            msg += cols[ins[2 * k]].to_bytes(32, "little")[:min(32, (ins[2 * k + 1] + 7) // 8)]
        v = int.from_bytes(hashes.blake2s(bytes(msg)), "big") % P
        if (mask & 1) and cols[out] != v:
            record_fail(opcode, EK_UNSAT)
        return [(out, v)]

    def fixed_base(cols, hdr, payload, record_fail):
        flags = hdr[0] >> 8
        opcode, ox, lo_w, hi_w, oy = hdr[1], hdr[2], hdr[3], hdr[4], hdr[5]
        try:
            x, y = grumpkin.fixed_base_scalar_mul(cols[lo_w], cols[hi_w])
        except grumpkin.BlackBoxFailed:
            record_fail(opcode, EK_BB_FAILED, 10)
            return []
        w = []
        for (slot, v, chk) in ((ox, x, flags & GF_OUT_CHECK_), (oy, y, flags & GF_OUT2_CHECK_)):
            if chk:
                if cols[slot] != v:
                    record_fail(opcode, EK_UNSAT)
                    w.append((slot, v))
            else:
                w.append((slot, v))
        return w

    def pedersen_hook(cols, hdr, payload, record_fail):
        from oracle import pedersen
        flags = hdr[0] >> 8
        opcode, ox, oy, off = hdr[1], hdr[2], hdr[5], hdr[7]
        n_in, iv = payload[off], payload[off + 1]
        x, y = pedersen.commit_native([cols[w] for w in payload[off + 2: off + 2 + n_in]], iv)
        w = []
        for (slot, v, chk) in ((ox, x, flags & GF_OUT_CHECK_), (oy, y, flags & GF_OUT2_CHECK_)):
            if chk:
                if cols[slot] != v:
                    record_fail(opcode, EK_UNSAT)
                    w.append((slot, v))
            else:
                w.append((slot, v))
        return w

    def ecdsa_hook(cols, hdr, payload, record_fail):
        from oracle import ecdsa
        flags = hdr[0] >> 8
        opcode, out, off = hdr[1], hdr[2], hdr[7]
        b = bytes(cols[w] & 0xFF for w in payload[off + 1: off + 161])
        try:
            v = 1 if ecdsa.verify(["EcdsaSecp256k1", "EcdsaSecp256r1"][payload[off]], b[128:160], b[0:32], b[32:64], b[64:128]) else 0
        except ecdsa.ReferencePanic:
            record_fail(opcode, EK_PANIC)
            return []
        if flags & GF_OUT_CHECK_:
            if cols[out] != v:
                record_fail(opcode, EK_UNSAT)
                return [(out, v)]
            return []
        return [(out, v)]

    # Split curve micro-ops.  The interpreter keeps a point in its three temporary slots as (x, y, 1) affine or (0, 0, 0):
    # the device's Jacobian/Montgomery encoding is not observable (temporaries never leave the device), only the
    # finaliser's canonical affine output is.
    def _pt_load(cols, base):
        return grumpkin.INF if cols[base + 2] == 0 else (cols[base], cols[base + 1])

    def _pt_store(base, pt):
        return [(base, 0), (base + 1, 0), (base + 2, 0)] if pt is grumpkin.INF else [(base, pt[0]), (base + 1, pt[1]), (base + 2, 1)]

    def curve_part(cols, hdr, payload, record_fail, consts=None):
        from oracle import pedersen
        mode, table, first, count, toff, imm = consts[0][:6]
        if mode == 0:
            v = cols[hdr[3]]
        elif mode == 1:
            v = pedersen.iv_point_x(imm)
        else:
            v = imm
        acc = grumpkin.INF
        for w in range(first, first + count):
            if table == 0:
                d = (v >> (8 * w)) & 0xFF
                if d:
                    acc = grumpkin.add(acc, grumpkin.mul(d << (8 * (toff + w)), grumpkin.G))
            else:
                s9 = (v >> (9 * w)) & 0x1FF
                par, i = divmod(toff + w, pedersen.NUM_WINDOWS)
                acc = grumpkin.add(acc, grumpkin.mul(s9 + 1, pedersen.generators()[par][i]))
        return _pt_store(hdr[2], acc)

    def jac_add(cols, hdr, payload, record_fail, consts=None):
        return _pt_store(hdr[2], grumpkin.add(_pt_load(cols, hdr[3]), _pt_load(cols, hdr[4])))

    def jac_final(cols, hdr, payload, record_fail, consts=None):
        flags = hdr[0] >> 8
        opcode = hdr[1]
        if consts[0][0]:
            try:
                grumpkin.fixed_base_scalar_mul(cols[hdr[6]], cols[hdr[7]])
            except grumpkin.BlackBoxFailed:
                record_fail(opcode, EK_BB_FAILED, 10)
                return []
        pt = _pt_load(cols, hdr[3])
        if hdr[4] != NONE:
            pt = grumpkin.add(pt, _pt_load(cols, hdr[4]))
        x, y = (0, 0) if pt is grumpkin.INF else pt
        if hdr[5] == NONE:
            return [(hdr[2], x)]
        w = []
        for (slot, v, chk) in ((hdr[2], x, flags & GF_OUT_CHECK_), (hdr[5], y, flags & GF_OUT2_CHECK_)):
            if chk:
                if cols[slot] != v:
                    record_fail(opcode, EK_UNSAT)
                    w.append((slot, v))
            else:
                w.append((slot, v))
        return w

    return {MK["CURVE_PART"]: curve_part, MK["JAC_ADD"]: jac_add, MK["JAC_FINAL"]: jac_final, MK["ECDSA"]: ecdsa_hook, MK["BLAKE2S"]: hash_hook(hashes.blake2s), MK["HASH_TO_FIELD"]: hash_to_field, MK["PEDERSEN"]: pedersen_hook, MK["SHA256"]: hash_hook(hashes.sha256), MK["KECCAK256"]: hash_hook(hashes.keccak256), MK["FIXED_BASE"]: fixed_base}


def run_plan(plan: PlanBlob, inputs, hooks=None, circuit=None):
    """inputs: dict witness->int.  Returns (status tuple, witness dict).

    status = ("Solved",) or ("Failure", err_kind, opcode_index, aux).  `hooks` maps heavy micro-op
    kinds to python callables (cols, hdr, payload) -> list of (slot, value) or raises."""
    if hooks is None:
        hooks = default_hooks()

    class Cols(dict):
        """a lane that already failed keeps executing on the device and reads garbage: model that as 0.
        Inside a step every read is logged with the slot thread that made it (`cur`): the kernel has ONE barrier per step, so
        no thread may read or write a column that another thread writes in the same step (checked at the end of each step)."""
        cur = None
        reads = {}

        def __missing__(self, k):
            return 0

        def __getitem__(self, k):
            if Cols.cur is not None:
                Cols.reads.setdefault(k, set()).add(Cols.cur)
            return dict.__getitem__(self, k)

        def get(self, k, default=None):
            if Cols.cur is not None:
                Cols.reads.setdefault(k, set()).add(Cols.cur)
            return dict.get(self, k, default)
    cols = Cols()
    for k_, w in enumerate(plan.input_witnesses):
        cols[w] = inputs[w] % P
        if plan.input_scaled and plan.input_scaled[k_]:
            cols[w] = cols[w] * R_MONT % P   # the scatter stores value * R for this input
    mu = {}      # mu index -> opcode that assigned it (this lane)
    fail = None  # (opcode, kind, aux)

    def record_fail(opcode, kind, aux=0):
        nonlocal fail
        key = (opcode, kind, aux)
        if fail is None or key < fail:
            fail = key

    def host_segment(opcode, off):
        """Brillig on the 'host' of the interpreter: the oracle's VM (test code) on the gathered slots."""
        from oracle import brillig_vm as obv, pwg as opwg
        nonlocal fail
        if fail is not None and fail[0] < opcode:
            return
        br = circuit.opcodes[opcode].body
        d = plan.host_desc
        p_ = off
        pred_slot = d[p_]; p_ += 1
        n_in = d[p_]; p_ += 1
        regs, mem = [], []
        for _ in range(n_in):
            arr, cnt = d[p_], d[p_ + 1]; p_ += 2
            vals = [cols[s_] for s_ in d[p_:p_ + cnt]]; p_ += cnt
            if arr:
                regs.append(len(mem)); mem += vals
            else:
                regs.append(vals[0])
        n_out = d[p_]; p_ += 1
        outs = []
        for _ in range(n_out):
            arr, cnt = d[p_], d[p_ + 1]; p_ += 2
            outs.append((arr, [(d[p_ + 2 * k], d[p_ + 2 * k + 1]) for k in range(cnt)])); p_ += 2 * cnt
        flat = [wk for _, ws in outs for wk in ws]
        if pred_slot != NONE and cols[pred_slot] == 0:
            result = [0] * len(flat)
        else:
            vm = obv.VM(regs, mem, br["bytecode"], br["foreign_call_results"], opwg.OracleBackend())
            try:
                st_ = vm.process_opcodes()
            except opwg.ReferencePanic:
                record_fail(opcode, EK_PANIC)
                return
            if st_[0] == "Failure":
                record_fail(opcode, 7, st_[2][-1])
                return
            if st_[0] == "ForeignCallWait":
                record_fail(opcode, 0xF)
                return
            result = []
            for i, (arr, ws) in enumerate(outs):
                reg = vm.get(i)
                if not arr:
                    result.append(reg)
                else:
                    if reg.bit_length() > 64 or reg + len(ws) > len(vm.memory):
                        record_fail(opcode, EK_PANIC)
                        return
                    result += vm.memory[reg:reg + len(ws)]
        for (w_, known_), v in zip(flat, result):
            if known_ and cols[w_] != v:
                record_fail(opcode, EK_UNSAT)
            cols[w_] = v

    def sort_segment(opcode, off):
        """PermutationSort host segment (kind 2): self-contained descriptor, oracle/sorting.py does the routing."""
        from oracle import sorting
        nonlocal fail
        if fail is not None and fail[0] < opcode:
            return
        d = plan.host_desc
        n, tup, n_sort = d[off:off + 3]
        p_ = off + 3
        sort_by = list(d[p_:p_ + n_sort]); p_ += n_sort
        vals = [cols[s_] for s_ in d[p_:p_ + n * tup]]; p_ += n * tup
        n_bits = d[p_]; p_ += 1
        outs = [(d[p_ + 2 * k], d[p_ + 2 * k + 1]) for k in range(n_bits)]
        try:
            bits = sorting.permutation_sort_bits([vals[i * tup:(i + 1) * tup] for i in range(n)], sort_by)
        except sorting.ReferencePanic:
            record_fail(opcode, EK_PANIC)
            return
        for (w_, known_), v in zip(outs, bits):
            v = int(v)
            if known_ and cols[w_] != v:
                record_fail(opcode, EK_UNSAT)
            cols[w_] = v

    step_plan = []
    if plan.segments:
        for (kind_, a_, b_, _c) in plan.segments:
            if kind_ == 0:
                step_plan += [("step", s_) for s_ in range(a_, a_ + b_)]
            else:
                step_plan.append(("host" if kind_ == 1 else "sort", a_, b_))
    else:
        step_plan = [("step", s_) for s_ in range(plan.n_steps)]
    # shared-memory ring of recent values: entry -> value; empty at every kernel launch (= device segment)
    ring = {}
    seg_first = {a_ for (kind_, a_, b_, _c) in plan.segments if kind_ == 0} if plan.segments else {0}

    class Operands:
        """cols[...] for the gate / logic / range micro-ops: a field with bit 31 set reads the ring (vm_kernel_impl.cuh load_op)"""
        def __init__(self):
            self.ring_reads = set()

        def __getitem__(self, field):
            if field & 0x80000000:
                self.ring_reads.add(field & 0x7FFFFFFF)
                return ring[field & 0x7FFFFFFF]   # KeyError = the plan points at an entry nothing wrote
            return cols[field]

    for item in step_plan:
        if item[0] == "host":
            host_segment(item[1], item[2])
            continue
        if item[0] == "sort":
            sort_segment(item[1], item[2])
            continue
        step = item[1]
        if step in seg_first:
            ring = {}
        writes = []
        ring_writes = []
        real_cols, cols = cols, cols   # (names kept: the gate code below reads operands through `ops`)
        ops = Operands()
        Cols.reads = {}
        owners = []          # slot thread of every entry of `writes`
        for s in range(plan.S):
            owners += [Cols.cur] * (len(writes) - len(owners))
            Cols.cur = s
            hdr, c = plan.record(step * plan.S + s)
            kind, flags = hdr[0] & 0xFF, hdr[0] >> 8
            opcode, out, x, y, w1, w2, aux = hdr[1:8]
            if kind == MK["NOP"]:
                continue
            if kind in (MK["GATE_ASSIGN"], MK["GATE_CHECK"]):
                # gate semantics on STORED values (plan.hpp OpRec, vm_kernel_impl.cuh exec_gate)
                nlin = (flags >> GF_NLIN_SHIFT) & 3
                nprod = (flags >> 13) & 3
                is_mul = bool(flags & GF_MUL)
                K = (1 if is_mul else 0) + nprod if (flags & GF_Y) else 0
                res = c[4] * RINV if K else c[4]   # with a reduction the constant is its initial accumulator, stored as const*R
                if flags & GF_Y:
                    if is_mul:
                        assert nlin <= 1
                        prod = (ops[x] + c[1]) * (ops[y] + c[2])
                        if flags & 4096:   # GF_ONE_RED
                            res += prod * RINV
                        else:
                            res += c[0] * RINV * RINV * prod
                        operands = [(w1, c[3], flags & 1024)][:nlin]
                    else:
                        operands = [(y, c[1], flags & 512), (w1, c[2], flags & 1024), (w2, c[3], flags & 2048)][:nlin + 1]
                    for i_, (slot_, coef_, neg_) in enumerate(operands):
                        if i_ < nprod:
                            res += coef_ * RINV * ops[slot_]
                        else:
                            res += -ops[slot_] if neg_ else ops[slot_]
                res %= P
                if kind == MK["GATE_ASSIGN"]:
                    if flags & GF_OUT_CHECK:
                        if cols[out] != res:
                            record_fail(opcode, EK_UNSAT)
                            writes.append((out, res))
                    else:
                        writes.append((out, res))
                    if plan.ring_slots and aux != NONE:
                        ring_writes.append((aux, res))
                elif res != 0:
                    record_fail(opcode, EK_UNSAT)
            elif kind in (MK["AND"], MK["XOR"]):
                m = (1 << aux) - 1 if aux < 256 else (1 << 256) - 1
                a, b = ops[x] & m, ops[y] & m
                res = ((a & b) if kind == MK["AND"] else (a ^ b)) % P
                if flags & GF_OUT_CHECK:
                    if cols[out] != res:
                        record_fail(opcode, EK_UNSAT)
                        writes.append((out, res))
                else:
                    writes.append((out, res))
                if plan.ring_slots and w1 != NONE:
                    ring_writes.append((w1, res))
            elif kind == MK["RANGE"]:
                if ops[x].bit_length() > aux:
                    record_fail(opcode, EK_UNSAT)
            elif kind == MK["GATE_GENERAL"]:
                pl = plan.payload
                off = aux
                n_mul, n_lin = pl[off], pl[off + 1]
                fe = lambda o_: sum(pl[o_ + j] << (32 * j) for j in range(8))
                q = fe(off + 2)
                p_ = off + 10
                known = lambda m: m == NONE or m in mu
                entries, mulrem = [], 0
                for _ in range(n_mul):
                    c = fe(p_) * RINV % P
                    w1_, m1, w2_, m2 = pl[p_ + 16:p_ + 20]
                    p_ += 20
                    k1, k2 = known(m1), known(m2)
                    if k1 and k2:
                        q = (q + c * cols[w1_] * cols[w2_]) % P
                    elif not k1 and not k2:
                        mulrem += c != 0
                    else:
                        v = c * cols[w1_ if k1 else w2_] % P
                        if v:
                            entries.append((v, w2_ if k1 else w1_, m2 if k1 else m1))
                for _ in range(n_lin):
                    c = fe(p_ + 8)
                    w_, m_ = pl[p_ + 16], pl[p_ + 17]
                    p_ += 18
                    if known(m_):
                        q = (q + c * cols[w_]) % P
                    elif c:
                        entries.append((c, w_, m_))
                if mulrem > 1:
                    record_fail(opcode, EK_PANIC)
                elif mulrem == 1 or len(entries) > 1:
                    record_fail(opcode, EK_TOO_MANY)
                elif not entries:
                    if q:
                        record_fail(opcode, EK_UNSAT)
                else:
                    coef, w_, m_ = entries[0]
                    writes.append((w_, (-q * pow(coef, P - 2, P)) % P))
                    mu[m_] = opcode
            elif kind == MK["COPY"]:
                writes.append((out, cols[x]))
            elif kind == MK["TO_LE_RADIX"]:
                pl = plan.payload
                n_b, radix = pl[aux], pl[aux + 1]
                nm = (n_b + 31) // 32
                mask = pl[aux + 2: aux + 2 + nm]
                outs = pl[aux + 2 + nm: aux + 2 + nm + n_b]
                v = cols[x]
                if n_b == 0:
                    record_fail(opcode, EK_UNSAT)
                else:
                    local = {}
                    for i in range(n_b):
                        d, v = (v % radix) & 0xFF, v // radix
                        if (mask[i // 32] >> (i % 32)) & 1:
                            old = local.get(outs[i], cols[outs[i]])
                            if old != d:
                                record_fail(opcode, EK_UNSAT)
                        local[outs[i]] = d
                    writes.extend(local.items())
                    if v:
                        record_fail(opcode, EK_UNSAT)
            elif kind == MK["QUOTIENT"]:
                a_, b_ = cols[x], cols[y]
                pred = cols[w1] != 0 if w1 != NONE else True
                qv, rv = (a_ // b_, a_ % b_) if (pred and b_) else (0, 0)
                local = {}
                for (slot, val, chk) in ((out, qv, flags & GF_OUT_CHECK), (w2, rv, flags & 128)):
                    if chk and local.get(slot, cols[slot]) != val:
                        record_fail(opcode, EK_UNSAT)
                    local[slot] = val
                writes.extend(local.items())
            elif kind in (MK["MEM_READ"], MK["MEM_WRITE"]):
                base, ln = plan.payload[aux], plan.payload[aux + 1]
                idx_v = cols[x]
                if idx_v.bit_length() > 64:
                    record_fail(opcode, EK_PANIC)
                else:
                    mi = idx_v & 0xFFFFFFFF
                    pred = cols[w1] != 0 if w1 != NONE else True
                    wants_read = cols[w2] == 0 if w2 != NONE else None   # witness-dependent selector (heavy_ops.cuh exec_mem)
                    if wants_read is False and kind == MK["MEM_READ"]:
                        record_fail(opcode, EK_MISSING, c[0] & 0xFFFFFFFF)
                    elif wants_read is True and kind == MK["MEM_WRITE"]:
                        record_fail(opcode, EK_PANIC)
                    elif kind == MK["MEM_READ"]:
                        if not pred:
                            writes.append((out, 0))
                        elif mi >= ln:
                            record_fail(opcode, EK_OOB, mi)
                        else:
                            writes.append((out, cols[base + mi]))
                    elif pred:
                        if mi >= ln:
                            record_fail(opcode, EK_OOB, mi)
                        else:
                            writes.append((base + mi, cols[y]))
            elif kind == MK["HASH_PACK"]:   # packed hash pipeline (heavy_ops.cuh exec_hash_pack / _core / _unpack)
                pl = plan.payload
                n_ = pl[aux]
                writes.append((out, sum((cols[pl[aux + 1 + k_]] & 0xFF) << (8 * k_) for k_ in range(n_))))
            elif kind == MK["HASH_CORE"]:
                from oracle import hashes as ohashes
                if w2 == 1:   # descriptor in the record's coefficient words
                    pl, aux = struct.unpack_from("<40I", plan.stream, (step * plan.S + s) * 192 + 32), 0
                else:
                    pl = plan.payload
                func_, n_, nch_ = pl[aux:aux + 3]
                msg = b"".join(cols[pl[aux + 3 + c_]].to_bytes(32, "little") for c_ in range(nch_))[:n_]
                dg_ = (ohashes.sha256, ohashes.keccak256, ohashes.blake2s)[func_](msg)
                writes.append((out, int.from_bytes(dg_, "little")))
            elif kind == MK["HASH_UNPACK"]:
                pl = plan.payload
                mask_ = pl[aux]
                dg_ = cols[x].to_bytes(32, "little")
                for i_ in range(32):
                    w_ = pl[aux + 1 + i_]
                    if (mask_ >> i_) & 1 and cols[w_] != dg_[i_]:
                        record_fail(opcode, EK_UNSAT)
                    writes.append((w_, dg_[i_]))
            elif kind == MK["INT_OP"]:   # a lowered Brillig BinaryIntOp (heavy_ops.cuh exec_int_op)
                from oracle import brillig_vm as obv, pwg as opwg
                try:
                    res = obv.bigint_op(aux & 0xFF, cols[x], cols[y], aux >> 8) % P
                except opwg.ReferencePanic:
                    record_fail(opcode, EK_PANIC)
                else:
                    if (flags & GF_OUT_CHECK) and cols[out] != res:
                        record_fail(opcode, EK_UNSAT)
                    writes.append((out, res))
            elif kind == MK["REQUIRE"]:
                pl = plan.payload
                for i in range(pl[aux]):
                    if pl[aux + 2 + 2 * i] not in mu:
                        record_fail(opcode, EK_MISSING, pl[aux + 1 + 2 * i])
                        break
            elif hooks and kind in hooks and kind in (MK["CURVE_PART"], MK["JAC_ADD"], MK["JAC_FINAL"]):
                words = [[(cv >> (32 * j)) & 0xFFFFFFFF for j in range(8)] for cv in c]
                writes.extend(hooks[kind](cols, hdr, plan.payload, record_fail, consts=words))
            elif hooks and kind in hooks:
                writes.extend(hooks[kind](cols, hdr, plan.payload, record_fail))
            else:
                raise NotImplementedError(f"plan_interp: micro-op kind {kind}")
        owners += [Cols.cur] * (len(writes) - len(owners))
        Cols.cur = None
        writer = {}
        for (slot, _v), own in zip(writes, owners):
            assert writer.setdefault(slot, own) == own, f"step {step}: column {slot} written by slot threads {writer[slot]} and {own}"
            others = Cols.reads.get(slot, set()) - {own}
            assert not others, f"step {step}: column {slot} written by slot thread {own} and read by {sorted(others)} in the same step"
        for (slot, v) in writes:
            cols[slot] = v
        # ring entries are written during the step with no barrier against its reads: no entry may be both
        assert not (ops.ring_reads & {e for e, _ in ring_writes}), f"step {step}: ring entry read and rewritten in the same step"
        assert len({e for e, _ in ring_writes}) == len(ring_writes) and all(e < plan.ring_slots for e, _ in ring_writes)
        for (e, v) in ring_writes:
            ring[e] = v
    fop = fail[0] if fail else 0xFFFFFFFF
    if plan.sf_present and plan.sf_opcode <= fop:
        status = ("Failure", plan.sf_kind, plan.sf_opcode, plan.sf_aux)
        fop = plan.sf_opcode
    elif fail and fail[1] == 0xF:
        status = ("RequiresForeignCall", 0, fail[0], 0)
    elif fail:
        status = ("Failure", fail[1], fail[0], fail[2])
    else:
        status = ("Solved",)
    wm = {}
    for w in range(plan.num_witnesses):
        ao = plan.assign_opcode[w]
        if ao == 0xFFFFFFFD:
            ao = mu.get(plan.mu_index_of[w], 0xFFFFFFFF)
        if ao == 0xFFFFFFFE or (ao != 0xFFFFFFFF and ao < fop):
            wm[w] = cols[w] * plan.unscale[w] * RINV % P if plan.unscale else cols[w]   # what the output gather does
    return status, wm
