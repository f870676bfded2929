// Plan compiler (see plan.hpp).  Static restatement of the reference's per-opcode decisions:
//   which witnesses are known when an opcode is reached   acvm/src/pwg/arithmetic.rs:212-239 (evaluate)
//   solvable / check / too-many-unknowns classification    acvm/src/pwg/arithmetic.rs:27-127,176-209
//   blackbox input pre-check + output insert_value          acvm/src/pwg/blackbox/mod.rs:32-62, pwg/mod.rs:338-357
#include "plan.hpp"
#include "brillig_host.hpp"

#include <algorithm>
#include <cstring>
#include <deque>
#include <map>
#include <stdexcept>
#include <thread>

namespace acvmb {

namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t ASSIGN_NEVER = 0xFFFFFFFFu;
constexpr uint32_t ASSIGN_INPUT = 0xFFFFFFFEu;
constexpr uint32_t ASSIGN_DYNAMIC = 0xFFFFFFFDu;
// static knowledge about a witness when an opcode is reached
enum : uint8_t { W_UNKNOWN = 0, W_KNOWN = 1, W_MAYBE = 2 };

struct Prod {
    U256 c;
    uint32_t a, b;
};
struct Lin {
    U256 c;
    uint32_t w;
};

class Scheduler {
   public:
    Scheduler(uint32_t S, uint32_t n_slots)
        : S_(S), ready_(n_slots, 0), war_(n_slots, 0), lvl_w_(n_slots, 0), lvl_r_(n_slots, 0), own_w_(n_slots, 0xFFFFFFFFu),
          own_r_(n_slots, 0xFFFFFFFFu), seen_(n_slots, 0) {}

    void grow_slots(uint32_t n) {
        if (n > ready_.size()) {
            ready_.resize(n, 0);
            war_.resize(n, 0);
            lvl_w_.resize(n, 0);
            lvl_r_.resize(n, 0);
            own_w_.resize(n, 0xFFFFFFFFu);
            own_r_.resize(n, 0xFFFFFFFFu);
            seen_.resize(n, 0);
        }
    }

    // everything placed so far completes before anything placed later starts (segment boundary)
    uint32_t barrier(uint32_t chunk_steps) {
        flush();
        floor_ = ((n_steps_ + chunk_steps - 1) / chunk_steps) * chunk_steps;
        n_steps_ = floor_;
        seg_starts_.push_back(floor_);   // a new kernel launch: the shared-memory ring starts empty
        return floor_;
    }

    static bool ring_kind(uint32_t kind) {
        return kind == MK_GATE_ASSIGN || kind == MK_GATE_CHECK || kind == MK_AND || kind == MK_XOR || kind == MK_RANGE;
    }

    // Curve micro-ops are 10..1000x slower than an arithmetic gate and a step costs as much as its slowest slot, so a
    // program-order list schedule that lets each curve call start wherever its operands happen to be ready scatters the slow
    // steps all over the stream.  From the first curve micro-op of a segment on, placement is deferred: flush() computes for
    // every buffered op its "curve depth" -- the number of curve CALLS (ACIR opcodes; a call may be many micro-ops) on the
    // longest slot-hazard chain (RAW/WAR/WAW) that ends in it -- orders the segment as
    //   [cheap depth 0] [curve depth 1] [cheap depth 1] [curve depth 2] ...
    // and list-schedules that order.  Curve calls of one depth are mutually independent and start on the same step, so
    // their partial sums, addition trees and finalisers line up in shared steps and the latencies overlap.  Any order
    // that respects the slot hazards computes the same values; failures are reported by lowest opcode index, not by step.
    static bool is_costly(uint32_t kind) {
        return kind == MK_FIXED_BASE || kind == MK_PEDERSEN || kind == MK_ECDSA || kind == MK_CURVE_PART || kind == MK_JAC_ADD ||
               kind == MK_JAC_FINAL;
    }

    bool slack_scheduling = true;   // PlanOptions::slack_scheduling
    bool dry = false;   // the inversion-recording pass of compile_plan(): its schedule is thrown away, do not build one

    void place(const OpRec& rec, const uint32_t* reads, size_t nr, const uint32_t* writes, size_t nw) {
        if (dry) return;
        const bool costly = is_costly(rec.w[0] & 0xFF);
        if (!buffering_ && !costly) {
            place_now(rec, reads, nr, writes, nw);
            return;
        }
        buffering_ = true;
        Pending p;
        p.rec = rec;
        p.off = (uint32_t)pool_.size();
        p.nr = (uint32_t)nr;
        p.nw = (uint32_t)nw;
        p.costly = costly;
        pool_.insert(pool_.end(), reads, reads + nr);
        pool_.insert(pool_.end(), writes, writes + nw);
        pend_.push_back(p);
    }

    void flush() {
        if (!buffering_) return;
        buffering_ = false;
        constexpr uint32_t NO_CALL = 0xFFFFFFFFu, MIXED = 0xFFFFFFFEu;
        std::vector<uint32_t> touched;
        std::vector<uint32_t> key(pend_.size());
        for (size_t i = 0; i < pend_.size(); ++i) {
            const Pending& p = pend_[i];
            const uint32_t* rd = pool_.data() + p.off;
            const uint32_t* wr = rd + p.nr;
            const uint32_t call = p.costly ? p.rec.w[1] : NO_CALL;   // ACIR opcode index of the curve call
            // a hazard on an op of another call (or on a cheap op) puts a curve micro-op one level deeper; hazards inside
            // its own call, and every hazard of a cheap op, keep the level
            auto via = [&](uint32_t lvl, uint32_t owner) { return (p.costly && owner != call) ? lvl + 1 : lvl; };
            uint32_t d = p.costly ? 1u : 0u;
            for (uint32_t k = 0; k < p.nr; ++k) d = std::max(d, via(lvl_w_[rd[k]], own_w_[rd[k]]));
            for (uint32_t k = 0; k < p.nw; ++k)
                d = std::max(d, std::max(via(lvl_w_[wr[k]], own_w_[wr[k]]), via(lvl_r_[wr[k]], own_r_[wr[k]])));
            for (uint32_t k = 0; k < p.nr; ++k) {
                const uint32_t s = rd[k];
                if (!seen_[s]) { seen_[s] = 1; touched.push_back(s); }
                if (d > lvl_r_[s]) { lvl_r_[s] = d; own_r_[s] = call; }
                else if (d == lvl_r_[s] && own_r_[s] != call) own_r_[s] = MIXED;
            }
            for (uint32_t k = 0; k < p.nw; ++k) {
                const uint32_t s = wr[k];
                if (!seen_[s]) { seen_[s] = 1; touched.push_back(s); }
                lvl_w_[s] = d; own_w_[s] = call;
                lvl_r_[s] = d; own_r_[s] = call;   // readers-since-last-write restarts
            }
            key[i] = 4 * d + (p.costly ? 0u : 2u);
        }
        // Slack.  A curve micro-op whose every successor (reader of its outputs, next writer of a slot it reads or writes) sits
        // two or more levels deeper need not run at its own level: the H1 sums over the fresh inputs of a Pedersen chain have
        // level 1 and are consumed at level k.  Running them all up front costs full-width steps of their own; moved to the
        // level just before their first successor -- listed after that level's own curve ops, key 4 d + 1 -- they fill the slots
        // that the addition trees and finalisers of the chain leave idle (place_now picks steps they do not lengthen).
        // Backward pass: bound(s) = the deepest level a predecessor through slot s may take, plus one.
        if (slack_scheduling) {
            constexpr uint32_t INF = 0xFFFFFFFFu;
            for (uint32_t s : touched) lvl_w_[s] = lvl_r_[s] = INF;   // reused as: bound by the next writer / by the readers before it
            for (size_t i = pend_.size(); i-- > 0;) {
                const Pending& p = pend_[i];
                const uint32_t* rd = pool_.data() + p.off;
                const uint32_t* wr = rd + p.nr;
                uint32_t d = key[i] >> 2, val;
                if (p.costly) {
                    uint32_t m = INF;
                    for (uint32_t k = 0; k < p.nw; ++k) m = std::min(m, std::min(lvl_r_[wr[k]], lvl_w_[wr[k]]));
                    for (uint32_t k = 0; k < p.nr; ++k) m = std::min(m, lvl_w_[rd[k]]);
                    if (m != INF && m >= d + 2) {
                        d = m - 1;
                        key[i] = 4 * d + 1;
                        val = d + 1;   // a predecessor may share the level of a slack successor (equal keys keep program order)
                    } else {
                        val = d;       // ... but must sort before a regular curve op: one level less
                    }
                } else {
                    val = d + 1;       // cheap ops of level d are listed after every curve op of level d
                }
                for (uint32_t k = 0; k < p.nw; ++k) { lvl_w_[wr[k]] = val; lvl_r_[wr[k]] = INF; }
                for (uint32_t k = 0; k < p.nr; ++k) lvl_r_[rd[k]] = std::min(lvl_r_[rd[k]], val);
            }
        }
        for (uint32_t s : touched) {
            lvl_w_[s] = lvl_r_[s] = 0;
            own_w_[s] = own_r_[s] = NO_CALL;
            seen_[s] = 0;
        }
        std::vector<uint32_t> order(pend_.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t)i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        uint32_t prev = 2;   // ops placed before buffering began are cheap, depth 0
        for (uint32_t i : order) {
            const uint32_t group = (key[i] & 3u) == 1u ? key[i] - 1 : key[i];   // slack ops share the steps of their level's curve ops
            if (group != prev) {
                floor_ = n_steps_;   // curve steps hold curve micro-ops only
                prev = group;
            }
            const Pending& p = pend_[i];
            const uint32_t* rd = pool_.data() + p.off;
            place_now(p.rec, rd, p.nr, rd + p.nr, p.nw, (key[i] & 3u) == 1u);
        }
        pend_.clear();
        pool_.clear();
    }

    // relative run time of a curve micro-op (profiles/r2_pedersen_chain_stalls_v7.txt), in 1/100 of the time of the
    // mixed additions of a 4-window partial sum: a step lasts as long as its slowest slot
    static uint32_t op_cost(const OpRec& rec) {
        switch (rec.w[0] & 0xFF) {
            case MK_CURVE_PART: return 15 + 48 * (rec.c[0][3] ? rec.c[0][3] - 1 : 0);
            case MK_JAC_ADD: return 68;
            case MK_JAC_FINAL: return 355;
            case MK_FIXED_BASE: case MK_PEDERSEN: case MK_ECDSA: return 2000;
            default: return 0;
        }
    }

    void place_now(const OpRec& rec, const uint32_t* reads, size_t nr, const uint32_t* writes, size_t nw, bool slack = false) {
        uint32_t e = floor_;
        for (size_t i = 0; i < nr; ++i) e = std::max(e, ready_[reads[i]]);
        for (size_t i = 0; i < nw; ++i) e = std::max(e, std::max(ready_[writes[i]], war_[writes[i]]));
        uint32_t s = find(e);
        const uint32_t cost = op_cost(rec);
        if (slack && cost) {
            // among the next steps with a free slot, the first one this op does not lengthen; failing that, the least lengthened
            uint32_t best = s, best_inc = cost > cost_[s] ? cost - cost_[s] : 0;
            for (uint32_t cand = s, n = 0; best_inc && n < 64; ++n) {
                cand = find(cand + 1);
                if (cand >= n_steps_) break;
                const uint32_t inc = cost > cost_[cand] ? cost - cost_[cand] : 0;
                if (inc < best_inc) { best = cand; best_inc = inc; }
            }
            s = best;
        }
        cost_[s] = std::max(cost_[s], cost);
        if (++fill_[s] == S_) next_[s] = s + 1;
        steps_.push_back(s);
        ops_.push_back(rec);
        if (!ring_kind(rec.w[0] & 0xFF)) {   // its writes bypass the ring: emit() must forget ring copies of those slots
            other_writes_at_.push_back((uint32_t)ops_.size() - 1);
            other_writes_off_.push_back((uint32_t)other_writes_.size());
            other_writes_.insert(other_writes_.end(), writes, writes + nw);
        }
        for (size_t i = 0; i < nw; ++i) ready_[writes[i]] = s + 1;
        for (size_t i = 0; i < nr; ++i) war_[reads[i]] = std::max(war_[reads[i]], s + 1);
        n_steps_ = std::max(n_steps_, s + 1);
    }

    uint32_t n_steps() { flush(); return n_steps_; }
    size_t n_ops() { flush(); return ops_.size(); }

    // Shared-memory ring of recent values (vm_kernel_impl.cuh): the k-th value written by a gate / logic micro-op of a segment
    // lives in ring entry k % W until the (k+W)-th one replaces it.  A read in step s may use the ring copy iff the value is
    // still there when step s ENDS (writes of step s land without a barrier).  Operand fields of such reads are rewritten to
    // RING_FLAG | entry; outputs get their entry in w[7] (gates) / w[5] (AND, XOR).
    void assign_ring(uint32_t W, uint64_t& n_reads, uint64_t& n_ring_reads) {
        flush();
        n_reads = n_ring_reads = 0;
        if (W == 0) return;
        const size_t n_ops = ops_.size();
        std::vector<uint32_t> order(n_ops);
        for (size_t i = 0; i < n_ops; ++i) order[i] = (uint32_t)i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return steps_[a] < steps_[b]; });
        std::vector<uint64_t> seq_of(ready_.size(), ~0ull);   // slot -> sequence number of its ring copy
        std::vector<uint32_t> ow_index(n_ops, 0xFFFFFFFFu);
        for (size_t k = 0; k < other_writes_at_.size(); ++k) ow_index[other_writes_at_[k]] = (uint32_t)k;
        std::vector<uint32_t> seg(seg_starts_);
        std::sort(seg.begin(), seg.end());
        size_t seg_pos = 0;
        uint64_t next_seq = 0;
        std::vector<uint32_t> touched;
        size_t i = 0;
        while (i < n_ops) {
            const uint32_t step = steps_[order[i]];
            size_t j = i;
            while (j < n_ops && steps_[order[j]] == step) ++j;
            while (seg_pos < seg.size() && seg[seg_pos] <= step) {   // new launch: forget everything
                for (uint32_t s : touched) seq_of[s] = ~0ull;
                touched.clear();
                ++seg_pos;
            }
            uint64_t n_writes = 0;
            for (size_t k = i; k < j; ++k) {
                const OpRec& r = ops_[order[k]];
                const uint32_t kind = r.w[0] & 0xFF;
                if ((kind == MK_GATE_ASSIGN || kind == MK_AND || kind == MK_XOR) && r.w[2] != 0xFFFFFFFFu) ++n_writes;
            }
            // more writes in one step than the ring holds: only the last W get an entry (two writes of one step must never
            // target the same entry -- nothing orders them)
            uint64_t skip = n_writes > W ? n_writes - W : 0;
            const uint64_t seq_end = next_seq + (n_writes - skip);
            auto rd = [&](uint32_t& field) {
                ++n_reads;
                const uint64_t q = field < seq_of.size() ? seq_of[field] : ~0ull;
                if (q != ~0ull && q + W >= seq_end) {
                    field = RING_FLAG | (uint32_t)(q % W);
                    ++n_ring_reads;
                }
            };
            for (size_t k = i; k < j; ++k) {
                OpRec& r = ops_[order[k]];
                const uint32_t kind = r.w[0] & 0xFF, flags = r.w[0] >> 8;
                if (kind == MK_GATE_ASSIGN || kind == MK_GATE_CHECK) {
                    if (flags & GF_Y) {
                        const uint32_t nlin = (flags >> GF_NLIN_SHIFT) & 3;
                        if (flags & GF_MUL) rd(r.w[3]);
                        rd(r.w[4]);
                        if (nlin >= 1) rd(r.w[5]);
                        if (nlin >= 2) rd(r.w[6]);
                    }
                } else if (kind == MK_AND || kind == MK_XOR) {
                    rd(r.w[3]);
                    rd(r.w[4]);
                } else if (kind == MK_RANGE) {
                    rd(r.w[3]);
                }
            }
            for (size_t k = i; k < j; ++k) {
                OpRec& r = ops_[order[k]];
                const uint32_t kind = r.w[0] & 0xFF;
                if (kind == MK_GATE_ASSIGN || kind == MK_AND || kind == MK_XOR) {
                    uint32_t& ring_field = kind == MK_GATE_ASSIGN ? r.w[7] : r.w[5];
                    ring_field = RING_NONE;
                    if (r.w[2] != 0xFFFFFFFFu && skip > 0) {
                        --skip;
                        if (r.w[2] < seq_of.size()) seq_of[r.w[2]] = ~0ull;
                    } else if (r.w[2] != 0xFFFFFFFFu) {
                        const uint64_t q = next_seq++;
                        ring_field = (uint32_t)(q % W);
                        if (seq_of[r.w[2]] == ~0ull) touched.push_back(r.w[2]);
                        seq_of[r.w[2]] = q;
                    }
                } else if (kind == MK_GATE_CHECK) {
                    r.w[7] = RING_NONE;
                } else if (ow_index[order[k]] != 0xFFFFFFFFu) {
                    const uint32_t o = ow_index[order[k]];
                    const uint32_t lo = other_writes_off_[o];
                    const uint32_t hi = o + 1 < other_writes_off_.size() ? other_writes_off_[o + 1] : (uint32_t)other_writes_.size();
                    for (uint32_t t = lo; t < hi; ++t)
                        if (other_writes_[t] < seq_of.size()) seq_of[other_writes_[t]] = ~0ull;
                }
            }
            i = j;
        }
    }

    // emit the dense [n_steps_padded][S] record array
    void emit(std::vector<OpRec>& stream, uint32_t& n_steps_padded, uint32_t chunk_steps, uint32_t slots_per_warp = 1) {
        flush();
        n_steps_padded = ((n_steps_ + chunk_steps - 1) / chunk_steps) * chunk_steps;
        if (n_steps_padded == 0) n_steps_padded = chunk_steps;
        stream.assign((size_t)n_steps_padded * S_, OpRec{});
        std::vector<uint32_t> cursor(n_steps_padded, 0);
        for (size_t i = 0; i < ops_.size(); ++i) {
            uint32_t s = steps_[i];
            stream[(size_t)s * S_ + cursor[s]++] = ops_[i];
        }
        // slots of one step are independent: order them so that the slots sharing a warp (32/T consecutive slots) run the same
        // code path -- by kind, then, for gates, by the width of their dot product (the lanes of a warp agree on the widest
        // one among them and the narrower lanes pad, vm_kernel_impl.cuh exec_gate), then by form; NOPs (kind 0) go last
        auto key = [](const OpRec& r) -> uint64_t {
            const uint32_t kind = r.w[0] & 0xFF, flags = r.w[0] >> 8;
            uint64_t k = kind == MK_GATE_CHECK ? MK_GATE_ASSIGN : kind;   // checks and assignments share the gate code
            uint32_t width = 0, two_red = 0;
            if (kind == MK_GATE_ASSIGN || kind == MK_GATE_CHECK) {
                const uint32_t mul = (flags & GF_MUL) ? 1u : 0u;
                width = (flags & GF_Y) ? mul + ((flags >> GF_NPROD_SHIFT) & 3) : 0;
                two_red = (mul && !(flags & GF_ONE_RED)) ? 1u : 0u;
            }
            return (k << 40) | ((uint64_t)width << 36) | ((uint64_t)two_red << 35) | (r.w[0] >> 8);
        };
        // Tiles narrower than a warp (slots_per_warp = 32 / T > 1): the heavy micro-ops of a step (hash, curve, general ops --
        // thousands of instructions each) go to DIFFERENT warps, one per warp before any warp gets a second one, so that a hash
        // core and the pack / unpack micro-ops beside it are not serialised by divergence inside one warp; the light ops keep
        // their sorted order in the slots that remain, warps without a heavy op first.
        const uint32_t spw = slots_per_warp > 1 && S_ % slots_per_warp == 0 ? slots_per_warp : 1;
        const uint32_t n_warps = S_ / spw;
        auto heavy = [](const OpRec& r) {
            const uint32_t kind = r.w[0] & 0xFF;
            return !(kind == MK_NOP || kind == MK_GATE_ASSIGN || kind == MK_GATE_CHECK || kind == MK_AND || kind == MK_XOR || kind == MK_RANGE);
        };
        std::vector<OpRec> tmp(S_);
        std::vector<uint8_t> taken(S_);
        for (uint32_t s = 0; s < n_steps_padded; ++s) {
            OpRec* first = &stream[(size_t)s * S_];
            std::stable_sort(first, first + cursor[s], [&](const OpRec& x, const OpRec& y) { return key(x) < key(y); });
            if (spw == 1 || n_warps < 2) continue;
            uint32_t n_heavy = 0;
            for (uint32_t j = 0; j < cursor[s]; ++j) n_heavy += heavy(first[j]) ? 1u : 0u;
            if (n_heavy == 0) continue;
            std::fill(tmp.begin(), tmp.end(), OpRec{});
            std::fill(taken.begin(), taken.end(), 0);
            uint32_t h = 0;
            for (uint32_t j = 0; j < cursor[s]; ++j)
                if (heavy(first[j])) {
                    const uint32_t pos = (h % n_warps) * spw + h / n_warps;
                    tmp[pos] = first[j];
                    taken[pos] = 1;
                    ++h;
                }
            // light ops: free slots of the warps that hold no heavy op first, then the rest, ascending
            std::vector<uint32_t> free_pos;
            for (int pass = 0; pass < 2; ++pass)
                for (uint32_t w = 0; w < n_warps; ++w) {
                    const bool has_heavy = w < std::min(n_heavy, n_warps);
                    if (has_heavy != (pass == 1)) continue;
                    for (uint32_t k = 0; k < spw; ++k)
                        if (!taken[w * spw + k]) free_pos.push_back(w * spw + k);
                }
            uint32_t f = 0;
            for (uint32_t j = 0; j < cursor[s]; ++j)
                if (!heavy(first[j])) tmp[free_pos[f++]] = first[j];
            std::copy(tmp.begin(), tmp.end(), first);
        }
    }

   private:
    uint32_t find(uint32_t s) {
        ensure(s);
        uint32_t r = s;
        while (true) {
            ensure(r);
            if (next_[r] == r) break;
            r = next_[r];
        }
        // path compression
        while (next_[s] != r && next_[s] != s) {
            uint32_t n = next_[s];
            next_[s] = r;
            s = n;
        }
        return r;
    }
    void ensure(uint32_t s) {
        while (fill_.size() <= s) {
            next_.push_back((uint32_t)fill_.size());
            fill_.push_back(0);
            cost_.push_back(0);
        }
    }
    struct Pending {
        OpRec rec;
        uint32_t off, nr, nw;
        bool costly;
    };
    std::vector<Pending> pend_;
    std::vector<uint32_t> pool_;            // reads then writes of every pending op
    std::vector<uint32_t> lvl_w_, lvl_r_;   // flush(): curve depth of the last writer / of the readers since
    std::vector<uint32_t> own_w_, own_r_;   // ... and the curve call they belong to
    std::vector<uint8_t> seen_;
    bool buffering_ = false;
    uint32_t S_;
    std::vector<uint32_t> ready_, war_;
    std::vector<uint32_t> fill_, next_, cost_;   // per step: used slots, next step with a free slot, cost of its slowest op
    std::vector<uint32_t> steps_;
    std::vector<OpRec> ops_;
    std::vector<uint32_t> other_writes_at_, other_writes_off_, other_writes_;   // write sets of the ops that bypass the ring
    std::vector<uint32_t> seg_starts_;
    uint32_t n_steps_ = 0;
    uint32_t floor_ = 0;
};

// thrown when a value-dependent gate meets a scaled column: compile_plan() starts over with canonical columns
struct NeedsCanonicalColumns {};

struct Compiler {
    const Circuit& c;
    PlanOptions opt;
    Plan plan;
    std::vector<uint8_t> known;
    Scheduler sched;
    uint32_t temp_base, temp_next = 0;
    uint32_t extra_slots_base = 0, extra_slots = 0;   // memory-block columns live after the temporaries

    Compiler(const Circuit& circ, const PlanOptions& o, uint32_t nw)
        : c(circ), opt(o), known(nw, 0), sched(o.S, nw + o.temp_pool), temp_base(nw), extra_slots_base(nw + o.temp_pool) {
        sched.slack_scheduling = o.slack_scheduling;
    }

    // Field inversions are the dominant plan-time cost (one or two per solving gate).  The compiler therefore runs twice:
    // a RECORD pass that only collects the values to invert (and gets a non-zero dummy back -- control flow never depends
    // on an inverse's value), one batched Montgomery-trick inversion of the whole list, and the real REPLAY pass.
    std::vector<U256>* inv_record = nullptr;
    const std::vector<U256>* inv_replay = nullptr;
    size_t inv_pos = 0;
    U256 inv(const U256& v) {
        if (inv_record) {
            inv_record->push_back(v);
            return hf::from_u64(1);
        }
        if (inv_replay) return (*inv_replay)[inv_pos++];
        return hf::inverse(v);
    }

    uint32_t new_temp() {
        uint32_t t = temp_base + (temp_next % opt.temp_pool);
        ++temp_next;
        set_canonical(t);   // whatever scale its previous value had
        return t;
    }

    // three consecutive temporaries: one Jacobian point
    uint32_t new_temp3() {
        if ((temp_next % opt.temp_pool) + 3 > opt.temp_pool) temp_next += opt.temp_pool - (temp_next % opt.temp_pool);
        uint32_t t = temp_base + (temp_next % opt.temp_pool);
        temp_next += 3;
        for (uint32_t i = 0; i < 3; ++i) set_canonical(t + i);
        return t;
    }

    // Dedicated (non-recycled) slots for values that must outlive the round-robin pool: the expression values a host
    // segment reads (Brillig inputs / predicate, PermutationSort tuples) stay live until that segment, the H1 points of a
    // Pedersen call until the chaining round that consumes them.  The round-robin pool never checks liveness, so anything
    // that is not consumed by the very next micro-ops of its own lowering lives here, after the pool, next to the memory
    // blocks.  Released slots are reused first-in-first-out, and only once `min_queue` others are waiting: immediate reuse
    // would chain independent curve calls through a false write-after-read dependency on the slot.
    std::deque<uint32_t> free1_, free3_;
    uint32_t new_pinned(uint32_t n, size_t min_queue) {
        auto& q = n == 1 ? free1_ : free3_;
        if (q.size() > min_queue) {
            uint32_t s = q.front();
            q.pop_front();
            for (uint32_t i = 0; i < n; ++i) set_canonical(s + i);
            return s;
        }
        uint32_t s = extra_slots_base + extra_slots;
        extra_slots += n;
        sched.grow_slots(s + n);
        return s;
    }
    void release_pinned(uint32_t s, uint32_t n) { (n == 1 ? free1_ : free3_).push_back(s); }

    static void put(uint32_t dst[8], const U256& v) { hf::to_limbs32(v, dst); }

    void fail_static(uint32_t opcode, uint32_t kind, uint32_t aux, const std::string& detail) {
        plan.static_fail.present = true;
        plan.static_fail.opcode = opcode;
        plan.static_fail.kind = kind;
        plan.static_fail.aux = aux;
        plan.static_fail.detail = detail;
    }

    void mark_assigned(uint32_t w, uint32_t opcode) {
        known[w] = 1;
        plan.assign_opcode[w] = opcode;
    }

    // ---- scaled columns ----------------------------------------------------------------------------------------------
    // A column does not have to hold the canonical value w: it holds s_w = lambda_w * w for a plan-time constant lambda_w
    // per slot -- Montgomery form (lambda = R for every value) generalised to one radix per column.  A gate
    //     out = M*(x+alpha)*(y+beta) + sum L_i*w_i + G
    // then needs ONE Montgomery reduction instead of two: with X' = s_x + alpha*lambda_x, Y' = s_y + beta*lambda_y,
    //     s_out = ( X'*Y' + sum s_i*(C_i*R) + (lambda_out*G)*R ) / R,   lambda_out = lambda_x*lambda_y / (M*R),
    //     C_i = lambda_out * L_i / lambda_i,
    // i.e. the multiplication by the gate's own coefficient M is absorbed into the scale of the output column, and a linear
    // term whose C_i is +-1 is a plain modular addition.  lambda and mu = 1/lambda are tracked with multiplications only
    // (the inverses needed are of circuit constants, which the record/replay batch inversion serves).  Canonical bytes are
    // produced where they are needed: the output gather multiplies by mu_w (Plan::unscale), and every column that a
    // non-arithmetic micro-op reads directly (blackbox inputs, memory blocks, host segments) is forced to lambda = 1.
    std::vector<U256> lam_, mu_;   // per slot; absent / never set = 1 (canonical)
    std::vector<uint8_t> need_canon;   // per witness: read directly by a non-gate micro-op
    bool scaled = false;
    const U256& lam_of(uint32_t s) const { return s < lam_.size() ? lam_[s] : hf::consts().one; }
    const U256& mu_of(uint32_t s) const { return s < mu_.size() ? mu_[s] : hf::consts().one; }
    void set_scale(uint32_t s, const U256& l, const U256& m) {
        if (!scaled) return;
        if (s >= lam_.size()) {
            if (l == hf::consts().one) return;
            lam_.resize((size_t)s + 1, hf::consts().one);
            mu_.resize((size_t)s + 1, hf::consts().one);
        }
        lam_[s] = l;
        mu_[s] = m;
    }
    void set_canonical(uint32_t s) { set_scale(s, hf::consts().one, hf::consts().one); }
    bool is_canonical(uint32_t s) const { return lam_of(s) == hf::consts().one; }

    // Lower  sum(products) + sum(linears) + constant  into chained micro-gates.
    // assign: write the value to `out`; otherwise CHECK it is zero.
    // out_check: `out` is already assigned -> compare instead of store (insert_value semantics).
    // `k` scales the whole sum (k = -1/coeff of the unknown for ASSIGN gates; `kinv` = 1/k = -coeff comes for free).  Terms
    // arrive UNSCALED so that the values handed to inv() never depend on another inverse (required by the record/replay
    // batching of inversions).  canonical_out: the value is read by something other than a gate -> lambda_out = 1.
    void lower_sum(std::vector<Prod> prods, std::vector<Lin> lins, const U256& constant, bool assign, uint32_t out,
                   uint32_t opcode, bool out_check, const U256* k = nullptr, const U256* kinv = nullptr, bool canonical_out = false) {
        uint32_t acc = NONE;
        const U256 one = hf::from_u64(1), pm1 = hf::neg(one);
        const U256 R = hf::consts().R, invR = hf::from_mont(one);
        auto sc = [&](const U256& v) { return k ? hf::mul(v, *k) : v; };
        struct Term {   // linear operand: true coefficient L (k applied), slot, and what is needed to invert L
            U256 L, Linv;
            uint32_t w;
        };
        bool first_gate = true;
        while (first_gate || !prods.empty() || !lins.empty()) {
            first_gate = false;
            OpRec r{};
            uint32_t flags = 0;
            uint32_t x = NONE, y = NONE;
            bool mul = false;
            U256 M, Minv, alpha, beta, gamma_part;
            std::vector<Term> terms;
            auto push_acc = [&] {
                terms.push_back(Term{one, one, acc});   // partial sums held in temporaries are already scaled by k
                acc = NONE;
            };
            // the inverse of a linear coefficient is only ever used to choose lambda_out of a gate without a product
            auto push_lin = [&](const Lin& l, bool want_inv) {
                Term t{sc(l.c), U256{}, l.w};
                if (want_inv) {
                    if (l.c == one) t.Linv = one;
                    else if (l.c == pm1) t.Linv = pm1;
                    else t.Linv = inv(l.c);
                    if (kinv) t.Linv = hf::mul(t.Linv, *kinv);
                }
                terms.push_back(t);
            };
            if (!prods.empty()) {
                // cM*x*y + cX*x + cY*y  ==  cM*(x + cY/cM)*(y + cX/cM) - cX*cY/cM : the x/y linear terms ride along as
                // plan-time constants
                Prod p = prods.back();
                prods.pop_back();
                mul = true;
                x = p.a;
                y = p.b;
                U256 cX, cY;
                for (size_t i = 0; i < lins.size(); ++i)
                    if (lins[i].w == y) {
                        cY = lins[i].c;
                        lins.erase(lins.begin() + i);
                        break;
                    }
                if (x != y)
                    for (size_t i = 0; i < lins.size(); ++i)
                        if (lins[i].w == x) {
                            cX = lins[i].c;
                            lins.erase(lins.begin() + i);
                            break;
                        }
                const bool fold = !cX.is_zero() || !cY.is_zero();
                U256 invM;
                if (fold || scaled) invM = inv(p.c);
                if (fold) {
                    alpha = hf::mul(cY, invM);
                    beta = hf::mul(cX, invM);
                    gamma_part = hf::neg(hf::mul(hf::mul(cX, cY), invM));
                }
                M = sc(p.c);
                Minv = kinv ? hf::mul(invM, *kinv) : invM;
                if (acc != NONE) push_acc();
                else if (!lins.empty()) {
                    push_lin(lins.back(), false);
                    lins.pop_back();
                }
            } else {
                if (acc != NONE) push_acc();
                while (terms.size() < 3 && !lins.empty()) {
                    push_lin(lins.back(), scaled);
                    lins.pop_back();
                }
            }
            const bool final_gate = prods.empty() && lins.empty() && acc == NONE;
            U256 G = sc(gamma_part);
            uint32_t kind, dst = NONE;
            if (final_gate) {
                kind = assign ? MK_GATE_ASSIGN : MK_GATE_CHECK;
                if (assign) {
                    dst = out;
                    if (out_check) flags |= GF_OUT_CHECK;
                }
                G = hf::add(G, sc(constant));
            } else {
                kind = MK_GATE_ASSIGN;
                dst = new_temp();
                ++plan.stats.n_temps;
            }
            if (inv_record) {
                // record pass of the batched inversion: only the SEQUENCE of inv() arguments matters, and none of them depends
                // on a column scale or a stored coefficient (they are circuit constants) -- skip that arithmetic
                if (!final_gate) acc = dst;
                continue;
            }
            // ---- scale of the output column ----
            bool forced = !scaled;
            U256 lo = one, mo = one;   // lambda_out, mu_out
            if (scaled && final_gate && assign) {
                if (out_check) {
                    forced = true;
                    lo = lam_of(out);
                    mo = mu_of(out);
                } else if (canonical_out || (out < need_canon.size() && need_canon[out])) {
                    forced = true;
                }
            }
            if (!forced) {
                if (mul) {   // absorb M: the product enters the reduction with coefficient 1
                    lo = hf::mul(hf::mul(hf::mul(lam_of(x), lam_of(y)), Minv), invR);
                    mo = hf::mul(hf::mul(hf::mul(mu_of(x), mu_of(y)), M), R);
                } else if (!terms.empty()) {
                    // lambda_out = lambda_i / L_i turns term i into a plain addition; keep the candidate that turns the most
                    // terms into additions (uniformly scaled operands with +-1 coefficients all become additions)
                    int best = -1;
                    for (size_t c = 0; c < terms.size(); ++c) {
                        U256 cl = hf::mul(lam_of(terms[c].w), terms[c].Linv);
                        int n_add = 0;
                        for (auto& t : terms) {
                            U256 C = hf::mul(hf::mul(cl, t.L), mu_of(t.w));
                            n_add += (C == one || C == pm1);
                        }
                        if (n_add > best) {
                            best = n_add;
                            lo = cl;
                            mo = hf::mul(mu_of(terms[c].w), terms[c].L);
                        }
                    }
                }
            }
            // ---- stored-value coefficients ----
            bool one_red = false;
            U256 coefM;
            if (mul) {
                coefM = hf::mul(hf::mul(lo, M), hf::mul(mu_of(x), mu_of(y)));
                one_red = scaled && coefM == invR;
                alpha = hf::mul(alpha, lam_of(x));
                beta = hf::mul(beta, lam_of(y));
            }
            struct Operand {
                U256 C;
                uint32_t w;
                bool add, neg;
            };
            std::vector<Operand> ops;
            for (auto& t : terms) {
                Operand o{hf::mul(hf::mul(lo, t.L), mu_of(t.w)), t.w, false, false};
                if (o.C == one) o.add = true;
                else if (o.C == pm1) o.add = o.neg = true;
                if (!o.C.is_zero()) ops.push_back(o);   // (a zero coefficient can only come from a dummy inverse of the record pass)
                else ops.push_back(Operand{one, t.w, true, false});
            }
            std::stable_sort(ops.begin(), ops.end(), [](const Operand& a, const Operand& b) { return !a.add && b.add; });
            uint32_t nprod = 0;
            for (auto& o : ops) nprod += !o.add;
            const uint32_t K = (mul ? 1u : 0u) + nprod;
            U256 Gs = hf::mul(lo, G);
            uint32_t w1 = NONE, w2 = NONE;
            if (mul) {
                flags |= GF_MUL | GF_Y;
                if (one_red) flags |= GF_ONE_RED;
                if (!ops.empty()) w1 = ops[0].w;
                flags |= (uint32_t)ops.size() << GF_NLIN_SHIFT;
                if (!ops.empty() && ops[0].neg) flags |= GF_NEG_W1;
                if (!one_red) put(r.c[0], hf::to_mont2(coefM));
                put(r.c[1], alpha);
                put(r.c[2], beta);
                if (!ops.empty() && !ops[0].add) put(r.c[3], hf::to_mont(ops[0].C));
            } else if (!ops.empty()) {
                flags |= GF_Y;
                y = ops[0].w;
                if (ops.size() > 1) w1 = ops[1].w;
                if (ops.size() > 2) w2 = ops[2].w;
                flags |= (uint32_t)(ops.size() - 1) << GF_NLIN_SHIFT;
                const uint32_t negbit[3] = {GF_NEG_Y, GF_NEG_W1, GF_NEG_W2};
                for (size_t i = 0; i < ops.size(); ++i) {
                    if (ops[i].neg) flags |= negbit[i];
                    if (!ops[i].add) put(r.c[1 + i], hf::to_mont(ops[i].C));
                }
                if (nprod == 0) flags |= GF_ADDSUB;
            }
            flags |= nprod << GF_NPROD_SHIFT;
            // with a reduction the constant is its initial accumulator (stored as const*R); otherwise it is used as is
            put(r.c[4], K ? hf::to_mont(Gs) : Gs);
            r.w[0] = kind | (flags << 8);
            r.w[1] = opcode;
            r.w[2] = dst;
            r.w[3] = x;
            r.w[4] = y;
            r.w[5] = w1;
            r.w[6] = w2;
            r.w[7] = 0;
            uint32_t reads[5];
            size_t nr = 0;
            if (mul) {
                reads[nr++] = x;
                reads[nr++] = y;
                if (w1 != NONE) reads[nr++] = w1;
            } else {
                if (y != NONE) reads[nr++] = y;
                if (w1 != NONE) reads[nr++] = w1;
                if (w2 != NONE) reads[nr++] = w2;
            }
            plan.stats.dev_imad += ((mul && !one_red) ? 136u : 0u) + (K ? 64u * K + 72u : 0u);
            {   // distinct operand reads
                uint32_t tmp[4];
                size_t nd = 0;
                for (size_t i = 0; i < nr; ++i) {
                    bool dup = false;
                    for (size_t j = 0; j < nd; ++j) dup |= tmp[j] == reads[i];
                    if (!dup) tmp[nd++] = reads[i];
                }
                plan.stats.alg_bytes += 32 * nd;
            }
            uint32_t writes[1];
            size_t nw = 0;
            if (dst != NONE) {
                if (flags & GF_OUT_CHECK) reads[nr++] = dst;   // compared, and REPLACED on mismatch (insert_value): ordered like a write
                writes[nw++] = dst;
                plan.stats.alg_bytes += 32;
                set_scale(dst, lo, mo);
            }
            sched.place(r, reads, nr, writes, nw);
            ++plan.stats.n_micro;
            if (kind == MK_GATE_ASSIGN) ++plan.stats.n_gate_assign; else ++plan.stats.n_gate_check;
            if (mul && one_red) ++plan.stats.n_gate_one_reduction;
            if (!final_gate) acc = dst;
        }
    }

    // returns false when compilation must stop (static failure)
    uint32_t mu_index(uint32_t w) {
        if (plan.mu_index_of[w] == NONE) plan.mu_index_of[w] = plan.n_mu++;
        return plan.mu_index_of[w];
    }

    void make_maybe(uint32_t w) {
        mu_index(w);
        known[w] = W_MAYBE;
        plan.assign_opcode[w] = ASSIGN_DYNAMIC;
    }

    // Value-dependent arithmetic opcode: which terms survive evaluate() (arithmetic.rs:212-239) depends on
    // per-instance VALUES (q_M * w_known == 0 drops the term) or on witnesses that only some instances have
    // assigned.  The whole reference decision procedure then runs per lane (exec_general, heavy_ops.cuh).
    // payload: n_mul, n_lin, qc[8], then per mul term {cR[8], cR2[8], w1, mu1, w2, mu2}, per lin term {cR[8], c[8], w, mu}
    // (mu = NONE for a statically known witness).
    void general_gate(uint32_t idx, const Expression& e) {
        if (scaled) {   // exec_general evaluates the reference's decision procedure on canonical values
            for (auto& t : e.mul_terms)
                if (!is_canonical(t.a) || !is_canonical(t.b)) throw NeedsCanonicalColumns{};
            for (auto& t : e.linear_combinations)
                if (!is_canonical(t.w)) throw NeedsCanonicalColumns{};
        }
        OpRec r{};
        std::vector<uint32_t> rd, wr;
        uint32_t off = (uint32_t)plan.payload.size();
        auto put_fe = [&](const U256& v) {
            uint32_t l[8];
            hf::to_limbs32(v, l);
            plan.payload.insert(plan.payload.end(), l, l + 8);
        };
        auto wref = [&](uint32_t w) {
            plan.payload.push_back(w);
            if (known[w] == W_KNOWN) {
                plan.payload.push_back(NONE);
                rd.push_back(w);
            } else {
                plan.payload.push_back(mu_index(w));
                rd.push_back(w);
                wr.push_back(w);
            }
        };
        plan.payload.push_back((uint32_t)e.mul_terms.size());
        plan.payload.push_back((uint32_t)e.linear_combinations.size());
        put_fe(e.q_c);
        for (auto& t : e.mul_terms) {
            put_fe(hf::to_mont(t.c));
            put_fe(hf::to_mont2(t.c));
            wref(t.a);
            wref(t.b);
            plan.stats.ref_fr_mul += 2;
        }
        for (auto& t : e.linear_combinations) {
            put_fe(hf::to_mont(t.c));
            put_fe(t.c);
            wref(t.w);
            plan.stats.ref_fr_mul += 1;
        }
        plan.stats.ref_fr_inv += 1;
        r.w[0] = MK_GATE_GENERAL | (GF_HEAVY << 8);
        r.w[1] = idx;
        r.w[2] = r.w[3] = r.w[4] = r.w[5] = r.w[6] = NONE;
        r.w[7] = off;
        std::sort(rd.begin(), rd.end());
        rd.erase(std::unique(rd.begin(), rd.end()), rd.end());
        std::sort(wr.begin(), wr.end());
        wr.erase(std::unique(wr.begin(), wr.end()), wr.end());
        sched.place(r, rd.data(), rd.size(), wr.data(), wr.size());
        for (uint32_t w : wr) make_maybe(w);
        plan.needs_full_kernel = true;
        ++plan.stats.n_micro;
        ++plan.stats.n_gate_general;
        plan.stats.alg_bytes += 32 * (rd.size() + 1);
    }

    // every input of a blackbox call must be assigned (blackbox/mod.rs:55-62); for value-dependent witnesses that is a
    // per-lane test, reported as MissingAssignment(first missing).  Survivors have them: they become statically known.
    void require_assigned(uint32_t idx, const std::vector<FunctionInput>& inputs) {
        std::vector<uint32_t> ws;
        for (auto& in : inputs)
            if (known[in.witness] == W_MAYBE) ws.push_back(in.witness);
        if (ws.empty()) return;
        OpRec r{};
        uint32_t off = (uint32_t)plan.payload.size();
        plan.payload.push_back((uint32_t)ws.size());
        for (uint32_t w : ws) {
            plan.payload.push_back(w);
            plan.payload.push_back(mu_index(w));
        }
        r.w[0] = MK_REQUIRE | (GF_HEAVY << 8);
        r.w[1] = idx;
        r.w[2] = r.w[3] = r.w[4] = r.w[5] = r.w[6] = NONE;
        r.w[7] = off;
        std::vector<uint32_t> rd(ws);
        std::sort(rd.begin(), rd.end());
        rd.erase(std::unique(rd.begin(), rd.end()), rd.end());
        sched.place(r, rd.data(), rd.size(), rd.data(), rd.size());   // ordered like a write: later readers wait for it
        for (uint32_t w : rd) known[w] = W_KNOWN;                     // presence in the output stays per-lane (ASSIGN_DYNAMIC)
        plan.needs_full_kernel = true;
        ++plan.stats.n_micro;
    }

    // get_value(expr) (pwg/mod.rs:321-332) for an expression whose witnesses are all statically known: returns the slot
    // that holds the value (the witness itself for `1*w`, else a temporary).  false => static MissingAssignment recorded.
    // `pinned`: the value is read by a host segment (or much later): give it a dedicated slot and list it for release.
    bool expr_to_slot(uint32_t idx, const Expression& e, uint32_t& slot, std::vector<uint32_t>* pinned = nullptr) {
        std::vector<Prod> prods;
        std::vector<Lin> lins;
        // get_value = evaluate + to_const, else MissingAssignment(any_witness_from_expression) (pwg/mod.rs:321-332,362-372):
        // the first entry of the evaluated linear_combinations (unknown linear terms, in order), else w1 of the first
        // surviving mul term.  A mul term with exactly one known operand turns into a linear entry only when
        // q_M * w_known != 0 -- per instance -- so it is refused here rather than decided statically.
        uint32_t missing_lin = NONE, missing_mul = NONE;
        for (auto& t : e.mul_terms) {
            if (known[t.a] == W_MAYBE || known[t.b] == W_MAYBE)
                throw std::runtime_error("opcode " + std::to_string(idx) + ": expression over a conditionally assigned witness is not supported here yet");
            if (known[t.a] && known[t.b]) {
                if (!t.c.is_zero()) prods.push_back({t.c, t.a, t.b});
            } else if (!t.c.is_zero()) {
                if (known[t.a] || known[t.b])
                    throw std::runtime_error("opcode " + std::to_string(idx) + ": directive / memory / Brillig input expression with a half-known "
                                             "multiplication term (value-dependent MissingAssignment) is not supported");
                if (missing_mul == NONE) missing_mul = t.a;
            }
        }
        for (auto& t : e.linear_combinations) {
            if (known[t.w] == W_MAYBE)
                throw std::runtime_error("opcode " + std::to_string(idx) + ": expression over a conditionally assigned witness is not supported here yet");
            if (known[t.w]) {
                if (!t.c.is_zero()) lins.push_back({t.c, t.w});
            } else if (!t.c.is_zero() && missing_lin == NONE) {
                missing_lin = t.w;
            }
        }
        const uint32_t missing = missing_lin != NONE ? missing_lin : missing_mul;
        if (missing != NONE) {
            fail_static(idx, EK_MISSING_ASSIGNMENT, missing, "missing assignment for witness index " + std::to_string(missing));
            return false;
        }
        if (prods.empty() && lins.size() == 1 && e.q_c.is_zero() && lins[0].c == hf::from_u64(1) && is_canonical(lins[0].w)) {
            slot = lins[0].w;
            return true;
        }
        if (pinned) {
            slot = new_pinned(1, 0);
            pinned->push_back(slot);
        } else {
            slot = new_temp();
        }
        ++plan.stats.n_temps;
        // the consumer is not a gate (directive, memory op, host segment): canonical value
        lower_sum(std::move(prods), std::move(lins), e.q_c, /*assign=*/true, slot, idx, false, nullptr, nullptr, /*canonical_out=*/true);
        return true;
    }

    static bool expr_is_const(const Expression& e, U256& v) {
        for (auto& t : e.mul_terms)
            if (!t.c.is_zero()) return false;
        for (auto& t : e.linear_combinations)
            if (!t.c.is_zero()) return false;
        v = e.q_c;
        return true;
    }

    void place_heavy(OpRec& r, std::vector<uint32_t>& rd, std::vector<uint32_t>& wr) {
        r.w[0] |= GF_HEAVY << 8;
        std::sort(rd.begin(), rd.end());
        rd.erase(std::unique(rd.begin(), rd.end()), rd.end());
        std::sort(wr.begin(), wr.end());
        wr.erase(std::unique(wr.begin(), wr.end()), wr.end());
        sched.place(r, rd.data(), rd.size(), wr.data(), wr.size());
        plan.needs_full_kernel = true;
        ++plan.stats.n_micro;
    }

    // Directive::Quotient / ToLeRadix (acvm/src/pwg/directives/mod.rs:28-87)
    bool directive(uint32_t idx, const Directive& d) {
        if (d.kind == DIR_Quotient) {
            uint32_t sa, sb, sp = NONE;
            if (!expr_to_slot(idx, d.a, sa)) return false;
            if (!expr_to_slot(idx, d.b, sb)) return false;
            if (d.predicate.present && !expr_to_slot(idx, d.predicate, sp)) return false;
            for (uint32_t w : {d.q, d.r})
                if (known[w] == W_MAYBE) throw std::runtime_error("opcode " + std::to_string(idx) + ": directive output is conditionally assigned; not supported yet");
            OpRec r{};
            uint32_t flags = 0;
            std::vector<uint32_t> rd = {sa, sb}, wr = {d.q, d.r};
            if (sp != NONE) rd.push_back(sp);
            if (known[d.q]) { flags |= GF_OUT_CHECK; rd.push_back(d.q); }
            if (known[d.r] || d.r == d.q) { flags |= GF_OUT2_CHECK; rd.push_back(d.r); }
            r.w[0] = MK_QUOTIENT | (flags << 8);
            r.w[1] = idx;
            r.w[2] = d.q;
            r.w[3] = sa;
            r.w[4] = sb;
            r.w[5] = sp;
            r.w[6] = d.r;
            r.w[7] = 0;
            place_heavy(r, rd, wr);
            if (!known[d.q]) mark_assigned(d.q, idx);
            if (!known[d.r]) mark_assigned(d.r, idx);
            ++plan.stats.n_directive;
            plan.stats.alg_bytes += 128;
            return true;
        }
        if (d.kind == DIR_ToLeRadix) {
            if (d.radix < 2 || d.radix > 256) {  // num-bigint to_radix_le asserts 2 <= radix <= 256
                fail_static(idx, EK_REFERENCE_PANIC, 0, "The radix must be within 2...256");
                return false;
            }
            uint32_t sa;
            if (!expr_to_slot(idx, d.a, sa)) return false;
            OpRec r{};
            std::vector<uint32_t> rd = {sa}, wr;
            uint32_t off = (uint32_t)plan.payload.size();
            uint32_t n_b = (uint32_t)d.out.size();
            plan.payload.push_back(n_b);
            plan.payload.push_back(d.radix);
            size_t mask_at = plan.payload.size();
            for (uint32_t i = 0; i < (n_b + 31) / 32; ++i) plan.payload.push_back(0);
            for (uint32_t i = 0; i < n_b; ++i) {
                uint32_t w = d.out[i];
                if (known[w] == W_MAYBE) throw std::runtime_error("opcode " + std::to_string(idx) + ": directive output is conditionally assigned; not supported yet");
                bool dup = false;   // the same witness listed twice: the second insert_value compares
                for (uint32_t j = 0; j < i; ++j) dup |= d.out[j] == w;
                if (known[w] || dup) {
                    plan.payload[mask_at + i / 32] |= 1u << (i % 32);
                    rd.push_back(w);
                }
                wr.push_back(w);
                plan.payload.push_back(w);
            }
            r.w[0] = MK_TO_LE_RADIX;
            r.w[1] = idx;
            r.w[2] = r.w[4] = r.w[5] = r.w[6] = NONE;
            r.w[3] = sa;
            r.w[7] = off;
            place_heavy(r, rd, wr);
            for (uint32_t w : d.out)
                if (!known[w]) mark_assigned(w, idx);
            ++plan.stats.n_directive;
            plan.stats.alg_bytes += 32 * (1 + n_b);
            return true;
        }
        // Directive::PermutationSort (directives/mod.rs:88-121): tuples are evaluated into slots on the device, the sort and
        // the switch routing run on the host between two device segments (sort_host.hpp), like Brillig.
        const uint32_t n = (uint32_t)d.sort_inputs.size();
        std::vector<uint32_t> slots, pinned;
        for (auto& element : d.sort_inputs) {
            if (element.size() != d.tuple) {
                fail_static(idx, EK_REFERENCE_PANIC, 0, "PermutationSort element does not have `tuple` entries");
                return false;
            }
            for (auto& e : element) {
                uint32_t s;
                if (!expr_to_slot(idx, e, s, &pinned)) return false;
                slots.push_back(s);
            }
        }
        if (n >= 2)
            for (uint32_t k : d.sort_by)
                if (k > d.tuple)
                    throw std::runtime_error("opcode " + std::to_string(idx) + ": PermutationSort sort_by index outside the tuple is not supported");
        uint32_t n_control = 0;   // switches of the network on n wires: sum of ceil(log2(i + 1))
        for (uint32_t i = 1; i < n; ++i) n_control += 32 - __builtin_clz(i);
        const uint32_t n_bits = std::min<uint32_t>(n_control, (uint32_t)d.out.size());   // bits.iter().zip(control)
        std::vector<uint32_t> desc = {n, d.tuple, (uint32_t)d.sort_by.size()};
        desc.insert(desc.end(), d.sort_by.begin(), d.sort_by.end());
        desc.insert(desc.end(), slots.begin(), slots.end());
        desc.push_back(n_bits);
        std::vector<uint32_t> outs;
        for (uint32_t k = 0; k < n_bits; ++k) {
            uint32_t w = d.out[k];
            if (known[w] == W_MAYBE)
                throw std::runtime_error("opcode " + std::to_string(idx) + ": PermutationSort output is conditionally assigned; not supported yet");
            bool dup = std::find(outs.begin(), outs.end(), w) != outs.end();
            desc.push_back(w);
            desc.push_back((known[w] || dup) ? 1u : 0u);
            outs.push_back(w);
        }
        close_device_segment();
        uint32_t off = (uint32_t)plan.host_desc.size();
        plan.host_desc.insert(plan.host_desc.end(), desc.begin(), desc.end());
        plan.segments.push_back(Segment{2, idx, off, 0});
        for (uint32_t s : pinned) release_pinned(s, 1);
        for (uint32_t w : outs)
            if (!known[w]) mark_assigned(w, idx);
        ++plan.stats.n_directive;
        return true;
    }

    // MemoryInit / MemoryOp (acvm/src/pwg/memory_op.rs:16-123).  A block is a run of extra columns; the dynamic index of
    // a MemoryOp is a per-lane column offset.
    struct Block {
        uint32_t base = 0, len = 0;
        bool inited = false;
    };
    std::vector<std::pair<uint32_t, Block>> blocks;
    std::vector<uint32_t> block_slots_scratch;

    Block& block_of(uint32_t id) {
        for (auto& b : blocks)
            if (b.first == id) return b.second;
        blocks.push_back({id, Block{}});
        return blocks.back().second;
    }

    bool memory_init(uint32_t idx, uint32_t block_id, const std::vector<uint32_t>& init) {
        for (uint32_t w : init) {
            if (!known[w]) {
                fail_static(idx, EK_MISSING_ASSIGNMENT, w, "missing assignment for witness index " + std::to_string(w));
                return false;
            }
            if (known[w] == W_MAYBE) throw std::runtime_error("opcode " + std::to_string(idx) + ": MemoryInit over a conditionally assigned witness; not supported yet");
        }
        Block& b = block_of(block_id);
        b.base = extra_slots_base + extra_slots;      // a re-init gets fresh columns
        b.len = (uint32_t)init.size();
        b.inited = true;
        extra_slots += b.len;
        sched.grow_slots(b.base + b.len);
        for (uint32_t i = 0; i < b.len; ++i) {
            OpRec r{};
            r.w[0] = MK_COPY;
            r.w[1] = idx;
            r.w[2] = b.base + i;
            r.w[3] = init[i];
            r.w[4] = r.w[5] = r.w[6] = NONE;
            std::vector<uint32_t> rd = {init[i]}, wr = {b.base + i};
            place_heavy(r, rd, wr);
        }
        ++plan.stats.n_memory;
        plan.stats.alg_bytes += 64ull * b.len;
        return true;
    }

    bool memory_op(uint32_t idx, const MemOp& m) {
        Block& b = block_of(m.block_id);
        // memory_op.rs:68-81 evaluates `operation` per instance.  Noir emits constants; a selector that depends on witnesses is
        // read from its column (canonical) and decided per lane.  Whether `value` is a readable witness or a writable value is
        // static, so such an opcode can succeed in one direction only and FAILS the lanes that take the other one exactly as
        // the reference does: a read lane of a known value panics ("Memory must be read into a specified witness index"), a
        // write lane of an unassigned witness is MissingAssignment(witness).
        U256 opv;
        uint32_t ssel = NONE;
        const bool const_sel = expr_is_const(m.operation, opv);
        if (!const_sel && !expr_to_slot(idx, m.operation, ssel)) return false;
        uint32_t si, sp = NONE;
        if (!expr_to_slot(idx, m.index, si)) return false;
        // to_witness(): exactly 1*w + 0 over a witness (after evaluate); anything else panics in the reference
        // (an already-known witness is folded into the constant by evaluate(), so to_witness() is None there too)
        std::vector<LinTerm> lin;
        bool has_mul = false;
        for (auto& t : m.value.mul_terms) has_mul |= !t.c.is_zero();
        for (auto& t : m.value.linear_combinations)
            if (!t.c.is_zero()) lin.push_back(t);
        const bool readable = !has_mul && lin.size() == 1 && lin[0].c == hf::from_u64(1) && m.value.q_c.is_zero() && !known[lin[0].w];
        bool is_read;
        if (const_sel) {
            is_read = opv.is_zero();
        } else {
            bool value_known = true;
            for (auto& t : m.value.mul_terms) value_known &= t.c.is_zero() || (known[t.a] == W_KNOWN && known[t.b] == W_KNOWN);
            for (auto& t : lin) value_known &= known[t.w] == W_KNOWN;
            if (readable && !m.predicate.present) is_read = true;     // write lanes: MissingAssignment(w)
            else if (value_known) is_read = false;                    // read lanes: the reference panics
            else
                throw std::runtime_error("opcode " + std::to_string(idx) + ": MemoryOp with a witness-dependent read/write selector whose "
                                         "value is neither one unassigned witness (without predicate) nor fully assigned is not supported");
        }
        // `value` is evaluated (not required) before the predicate (memory_op.rs:75-87)
        uint32_t sv = NONE, out_w = NONE;
        if (is_read) {
            if (!readable) {
                if (m.predicate.present && !expr_to_slot(idx, m.predicate, sp)) return false;
                fail_static(idx, EK_REFERENCE_PANIC, 0, "Memory must be read into a specified witness index, encountered an Expression");
                return false;
            }
            out_w = lin[0].w;
        }
        if (m.predicate.present && !expr_to_slot(idx, m.predicate, sp)) return false;
        if (!is_read) {
            // get_value(value) only happens when the predicate is non-zero; statically unknown witnesses there would be a
            // value-dependent MissingAssignment -- require them known at plan time
            if (!expr_to_slot(idx, m.value, sv)) return false;
        }
        OpRec r{};
        uint32_t off = (uint32_t)plan.payload.size();
        plan.payload.push_back(b.base);
        plan.payload.push_back(b.len);
        std::vector<uint32_t> rd = {si}, wr;
        if (sp != NONE) rd.push_back(sp);
        if (ssel != NONE) rd.push_back(ssel);
        for (uint32_t i = 0; i < b.len; ++i) rd.push_back(b.base + i);
        r.w[1] = idx;
        r.w[3] = si;
        r.w[5] = sp;
        r.w[6] = ssel;   // NONE: the direction is the micro-op kind; else the lanes whose selector disagrees with it fail
        r.w[7] = off;
        if (is_read) {
            r.w[0] = MK_MEM_READ;
            r.w[2] = out_w;
            r.w[4] = NONE;
            r.c[0][0] = out_w;   // MissingAssignment(witness) of a lane that wanted to write
            wr.push_back(out_w);
            place_heavy(r, rd, wr);
            mark_assigned(out_w, idx);
        } else {
            r.w[0] = MK_MEM_WRITE;
            r.w[2] = NONE;
            r.w[4] = sv;
            rd.push_back(sv);
            for (uint32_t i = 0; i < b.len; ++i) wr.push_back(b.base + i);
            place_heavy(r, rd, wr);
        }
        ++plan.stats.n_memory;
        plan.stats.alg_bytes += 96;
        return true;
    }

    uint32_t seg_start = 0;
    void close_device_segment() {
        uint32_t end = sched.barrier(opt.chunk_steps);
        if (end > seg_start) plan.segments.push_back(Segment{0, seg_start, end - seg_start, 0});
        seg_start = end;
    }

    // ---- Brillig on the device: plan-time symbolic execution of straight-line field bytecode ------------------------
    // A VM register / memory cell is a constant or a slot.  Supported: BinaryFieldOp Add/Sub/Mul, Const, Mov, Stop (and
    // running off the end of the bytecode, which the VM also treats as Finished, brillig_vm/src/lib.rs:140-151); no
    // predicate, no control flow, no memory opcodes; array inputs / outputs through constant pointers (brillig.rs:46-76,
    // 97-113).  Everything else keeps the host VM.  emit == false is a dry run that only decides feasibility.
    struct Sym {
        bool is_const = true;
        U256 c;
        uint32_t slot = NONE;
    };
    bool brillig_symbolic(uint32_t idx, const Brillig& br, bool emit, std::vector<uint32_t>& pinned) {
        if (br.bytecode.empty() || br.predicate.present) return false;
        // integer ops read canonical values: with any of them in the bytecode every intermediate value is kept canonical
        bool has_int = false;
        for (auto& o : br.bytecode) has_int |= o.tag == 1;
        std::vector<Sym> regs, mem;
        auto get = [&](uint64_t r) { return r < regs.size() ? regs[r] : Sym{}; };
        auto set = [&](uint64_t r, const Sym& v) {
            if (regs.size() <= r) regs.resize(r + 1);
            regs[r] = v;
        };
        auto input = [&](const Expression& e, Sym& out) {
            U256 cv;
            if (expr_is_const(e, cv)) {
                out = Sym{true, cv, NONE};
                return true;
            }
            out.is_const = false;
            out.slot = NONE;
            if (!emit) {   // feasibility only: every witness of the expression must be statically known
                for (auto& t : e.mul_terms)
                    if (!t.c.is_zero() && (known[t.a] != W_KNOWN || known[t.b] != W_KNOWN)) return false;
                for (auto& t : e.linear_combinations)
                    if (!t.c.is_zero() && known[t.w] != W_KNOWN) return false;
                return true;
            }
            return expr_to_slot(idx, e, out.slot, &pinned);
        };
        for (auto& in : br.inputs) {
            if (!in.is_array) {
                Sym v;
                if (!input(in.exprs[0], v)) return false;
                regs.push_back(v);
            } else {
                regs.push_back(Sym{true, hf::from_u64(mem.size()), NONE});
                for (auto& e : in.exprs) {
                    Sym v;
                    if (!input(e, v)) return false;
                    mem.push_back(v);
                }
            }
        }
        const U256 one = hf::from_u64(1);
        for (size_t pc = 0; pc < br.bytecode.size(); ++pc) {
            const BrilligOp& o = br.bytecode[pc];
            if (o.r0 >= (1u << 16) || o.r1 >= (1u << 16) || o.r2 >= (1u << 16)) return false;
            if (o.tag == 14) break;                                    // Stop
            if (o.tag == 6) { set(o.r0, Sym{true, o.value, NONE}); continue; }   // Const
            if (o.tag == 9) { set(o.r0, get(o.r1)); continue; }                  // Mov
            if (o.tag == 1) {   // BinaryIntOp (brillig_vm/src/arithmetic.rs:23-81): every op but SignedDiv, 1 <= bit_size <= 128
                if (o.bop == 3 || o.bop > 12 || o.bit_size < 1 || o.bit_size > 128) return false;
                Sym a = get(o.r1), b = get(o.r2), r;
                if (a.is_const && b.is_const) {
                    try {
                        r.c = hf::reduce(bvm::bigint_op(o.bop, a.c, b.c, o.bit_size));
                    } catch (const bvm::PanicEx&) {
                        return false;   // the reference panics for every instance: the host VM reports it
                    }
                } else {
                    r.is_const = false;
                    if (emit) {
                        auto slot_of = [&](const Sym& v) {   // a constant operand gets a (canonical) column of its own
                            if (!v.is_const) return v.slot;
                            uint32_t t = new_pinned(1, 0);
                            pinned.push_back(t);
                            ++plan.stats.n_temps;
                            lower_sum({}, {}, v.c, /*assign=*/true, t, idx, false, nullptr, nullptr, /*canonical_out=*/true);
                            return t;
                        };
                        const uint32_t sa = slot_of(a), sb = slot_of(b);
                        r.slot = new_pinned(1, 0);
                        pinned.push_back(r.slot);
                        OpRec rec{};
                        rec.w[0] = MK_INT_OP;
                        rec.w[1] = idx;
                        rec.w[2] = r.slot;
                        rec.w[3] = sa;
                        rec.w[4] = sb;
                        rec.w[5] = rec.w[6] = NONE;
                        rec.w[7] = o.bop | (o.bit_size << 8);
                        std::vector<uint32_t> rd = {sa, sb}, wr = {r.slot};
                        place_heavy(rec, rd, wr);
                        plan.stats.alg_bytes += 96;
                    }
                }
                set(o.r0, r);
                continue;
            }
            if (o.tag != 0 || o.bop > 2) return false;                 // BinaryFieldOp: Add / Sub / Mul only
            Sym a = get(o.r1), b = get(o.r2), r;
            if (a.is_const && b.is_const) {
                r.c = o.bop == 0 ? hf::add(a.c, b.c) : o.bop == 1 ? hf::sub(a.c, b.c) : hf::mul(a.c, b.c);
            } else {
                r.is_const = false;
                if (emit) {
                    r.slot = new_pinned(1, 0);
                    pinned.push_back(r.slot);
                    std::vector<Prod> prods;
                    std::vector<Lin> lins;
                    U256 cst;
                    if (o.bop == 2) {
                        if (!a.is_const && !b.is_const) prods.push_back({one, a.slot, b.slot});
                        else if (a.is_const) { if (!a.c.is_zero()) lins.push_back({a.c, b.slot}); }
                        else if (!b.c.is_zero()) lins.push_back({b.c, a.slot});
                    } else {
                        const U256 sb = o.bop == 0 ? one : hf::neg(one);
                        if (a.is_const) cst = a.c; else lins.push_back({one, a.slot});
                        if (b.is_const) cst = hf::add(cst, o.bop == 0 ? b.c : hf::neg(b.c));
                        else if (!a.is_const && a.slot == b.slot) {   // x + x / x - x: one linear term
                            lins.clear();
                            if (o.bop == 0) lins.push_back({hf::from_u64(2), a.slot});
                        } else lins.push_back({sb, b.slot});
                    }
                    ++plan.stats.n_temps;
                    lower_sum(std::move(prods), std::move(lins), cst, /*assign=*/true, r.slot, idx, false, nullptr, nullptr,
                              /*canonical_out=*/has_int);
                }
            }
            set(o.r0, r);
        }
        // outputs (brillig.rs:97-113): register i, or memory behind the pointer in register i; insert_value semantics
        std::vector<std::pair<uint32_t, Sym>> outs;
        for (size_t i = 0; i < br.outputs.size(); ++i) {
            Sym v = get(i);
            if (!br.outputs[i].is_array) {
                outs.emplace_back(br.outputs[i].witnesses[0], v);
                continue;
            }
            if (!v.is_const || v.c.l[1] || v.c.l[2] || v.c.l[3]) return false;
            for (size_t k = 0; k < br.outputs[i].witnesses.size(); ++k) {
                uint64_t p = v.c.l[0] + k;
                if (p >= mem.size()) return false;   // the reference panics on the index: leave it to the host VM's report
                outs.emplace_back(br.outputs[i].witnesses[k], mem[p]);
            }
        }
        for (auto& ov : outs)
            if (known[ov.first] == W_MAYBE) return false;
        if (!emit) return true;
        for (auto& ov : outs) {
            const uint32_t w = ov.first;
            const bool chk = known[w] != 0;
            if (ov.second.is_const) lower_sum({}, {}, ov.second.c, /*assign=*/true, w, idx, chk);
            else lower_sum({}, {Lin{one, ov.second.slot}}, U256{}, /*assign=*/true, w, idx, chk);
            if (!chk) mark_assigned(w, idx);
        }
        return true;
    }

    // Opcode::Brillig (acvm/src/pwg/brillig.rs:20-131): predicate and input expressions are evaluated on the device into
    // slots; the VM itself runs on the host between two device segments.
    bool brillig(uint32_t idx, const Brillig& br) {
        for (auto& o : br.bytecode)
            if (o.tag == 12 && (o.bb_tag == 6 || o.bb_tag == 7))   // SchnorrVerify / Pedersen need barretenberg (parity unpinned)
                throw std::runtime_error("opcode " + std::to_string(idx) + ": Brillig BlackBox op " + std::to_string(o.bb_tag) +
                                         " is not supported by the host VM yet");
        uint32_t sp = NONE;
        std::vector<uint32_t> pinned;
        if (opt.device_brillig && brillig_symbolic(idx, br, /*emit=*/false, pinned)) {
            if (!brillig_symbolic(idx, br, /*emit=*/true, pinned)) throw std::runtime_error("plan: device Brillig lowering diverged from its dry run");
            for (uint32_t s : pinned) release_pinned(s, 1);
            ++plan.stats.n_brillig;
            ++plan.stats.n_brillig_device;
            return true;
        }
        if (br.predicate.present && !expr_to_slot(idx, br.predicate, sp, &pinned)) return false;
        std::vector<uint32_t> desc;
        desc.push_back(sp);
        desc.push_back((uint32_t)br.inputs.size());
        for (auto& in : br.inputs) {
            desc.push_back(in.is_array ? 1u : 0u);
            desc.push_back((uint32_t)in.exprs.size());
            for (auto& e : in.exprs) {
                uint32_t s;
                if (!expr_to_slot(idx, e, s, &pinned)) {
                    // get_value() failure on an input is reported as ExpressionHasTooManyUnknowns (brillig.rs:49-55)
                    plan.static_fail.kind = EK_TOO_MANY_UNKNOWNS;
                    plan.static_fail.aux = 0;
                    return false;
                }
                desc.push_back(s);
            }
        }
        desc.push_back((uint32_t)br.outputs.size());
        std::vector<uint32_t> outs;
        for (auto& out : br.outputs) {
            desc.push_back(out.is_array ? 1u : 0u);
            desc.push_back((uint32_t)out.witnesses.size());
            for (uint32_t w : out.witnesses) {
                if (known[w] == W_MAYBE)
                    throw std::runtime_error("opcode " + std::to_string(idx) + ": Brillig output is conditionally assigned; not supported yet");
                bool dup = std::find(outs.begin(), outs.end(), w) != outs.end();
                desc.push_back(w);
                desc.push_back((known[w] || dup) ? 1u : 0u);
                outs.push_back(w);
            }
        }
        close_device_segment();
        uint32_t off = (uint32_t)plan.host_desc.size();
        plan.host_desc.insert(plan.host_desc.end(), desc.begin(), desc.end());
        plan.segments.push_back(Segment{1, idx, off, 0});
        for (uint32_t s : pinned) release_pinned(s, 1);
        for (uint32_t w : outs)
            if (!known[w]) mark_assigned(w, idx);
        ++plan.stats.n_brillig;
        return true;
    }

    bool arithmetic(uint32_t idx, const Expression& e) {
        {
            bool general = false;
            for (auto& t : e.mul_terms) {
                uint8_t ka = known[t.a], kb = known[t.b];
                if (ka == W_MAYBE || kb == W_MAYBE) general = true;
                if (!t.c.is_zero() && ((ka == W_KNOWN) != (kb == W_KNOWN))) general = true;
            }
            for (auto& t : e.linear_combinations)
                if (known[t.w] == W_MAYBE) general = true;
            if (general) {
                general_gate(idx, e);
                return true;
            }
        }
        std::vector<Prod> prods;
        std::vector<Lin> lins;
        // unknown linear entries after the reference's evaluate()
        std::vector<Lin> unknown;
        uint32_t n_both_unknown = 0;
        bool value_dependent = false;
        for (auto& t : e.mul_terms) {
            bool ka = known[t.a], kb = known[t.b];
            if (ka && kb) {
                plan.stats.ref_fr_mul += 2;
                if (!t.c.is_zero()) prods.push_back({t.c, t.a, t.b});
            } else if (!ka && !kb) {
                if (!t.c.is_zero()) ++n_both_unknown;
            } else {
                plan.stats.ref_fr_mul += 1;
                if (!t.c.is_zero()) value_dependent = true;  // coefficient q_M*w_known is per-instance
            }
        }
        for (auto& t : e.linear_combinations) {
            if (known[t.w]) {
                plan.stats.ref_fr_mul += 1;
                if (!t.c.is_zero()) lins.push_back({t.c, t.w});
            } else if (!t.c.is_zero()) {
                unknown.push_back({t.c, t.w});
            }
        }
        (void)value_dependent;  // handled by general_gate() above
        if (n_both_unknown > 1) {
            fail_static(idx, EK_REFERENCE_PANIC, 0, "Mul term in the arithmetic opcode must contain either zero or one term");
            return false;
        }
        if (n_both_unknown == 1 || unknown.size() > 1) {
            fail_static(idx, EK_TOO_MANY_UNKNOWNS, 0, "expression has too many unknowns");
            return false;
        }
        if (unknown.empty()) {
            lower_sum(std::move(prods), std::move(lins), e.q_c, /*assign=*/false, NONE, idx, false);
            return true;
        }
        // exactly one unknown (coeff != 0): w := -(sum)/coeff ; fold k = -1/coeff into every term
        U256 k = hf::neg(inv(unknown[0].c)), kinv = hf::neg(unknown[0].c);
        plan.stats.ref_fr_mul += 1;
        plan.stats.ref_fr_inv += 1;
        uint32_t w = unknown[0].w;
        lower_sum(std::move(prods), std::move(lins), e.q_c, /*assign=*/true, w, idx, false, &k, &kinv);
        mark_assigned(w, idx);
        return true;
    }

    // ---- plan-level parallel curve sums (device side: heavy_ops.cuh, MK_CURVE_PART / MK_JAC_ADD / MK_JAC_FINAL) ----
    bool split_curve_ops() const { return opt.split_curve && opt.S >= 8; }

    // partial sums over `n_windows` windows of one scalar, `per` windows per micro-op; returns the point slots
    void curve_parts(uint32_t idx, uint32_t mode, uint32_t src_slot, uint32_t imm, uint32_t table, uint32_t n_windows, uint32_t per,
                     uint32_t toff, std::vector<uint32_t>& points) {
        for (uint32_t first = 0; first < n_windows; first += per) {
            OpRec r{};
            uint32_t out = new_temp3();
            r.w[0] = MK_CURVE_PART;
            r.w[1] = idx;
            r.w[2] = out;
            r.w[3] = mode == 0 ? src_slot : NONE;
            r.w[4] = r.w[5] = r.w[6] = r.w[7] = NONE;
            r.c[0][0] = mode;
            r.c[0][1] = table;
            r.c[0][2] = first;
            r.c[0][3] = std::min(per, n_windows - first);
            r.c[0][4] = toff;
            r.c[0][5] = imm;
            std::vector<uint32_t> rd, wr = {out, out + 1, out + 2};
            if (mode == 0) rd.push_back(src_slot);
            place_heavy(r, rd, wr);
            points.push_back(out);
        }
    }
    // sum of the 29 table points a CONSTANT scalar selects (mode 1: Pedersen IV table entry `imm`, mode 2: the immediate `imm`):
    // computed by the first call that needs it, in a slot triple that is never handed out again
    std::map<uint64_t, uint32_t> const_points_;
    uint32_t const_curve_point(uint32_t idx, uint32_t mode, uint32_t imm, uint32_t toff) {
        const uint64_t key = ((uint64_t)mode << 56) | ((uint64_t)toff << 40) | imm;
        auto it = const_points_.find(key);
        if (it != const_points_.end()) return it->second;
        std::vector<uint32_t> pts;
        curve_parts(idx, mode, NONE, imm, 1, 29, 4, toff, pts);
        const uint32_t slot = new_pinned(3, 1u << 30);
        curve_reduce(idx, pts, 1, slot);
        const_points_[key] = slot;
        return slot;
    }
    // pairwise tree of Jacobian additions until at most `keep` points remain
    // `final_dst` (with keep == 1): the slot triple that receives the one remaining point
    void curve_reduce(uint32_t idx, std::vector<uint32_t>& points, size_t keep, uint32_t final_dst = NONE) {
        if (final_dst != NONE && (keep != 1 || points.size() < 2)) throw std::runtime_error("plan: curve_reduce final_dst misuse");
        while (points.size() > keep) {
            std::vector<uint32_t> next;
            size_t n_pairs = std::min(points.size() / 2, points.size() - keep);
            for (size_t i = 0; i < n_pairs; ++i) {
                OpRec r{};
                uint32_t a = points[2 * i], b = points[2 * i + 1];
                uint32_t out = (final_dst != NONE && points.size() == 2) ? final_dst : new_temp3();
                r.w[0] = MK_JAC_ADD;
                r.w[1] = idx;
                r.w[2] = out;
                r.w[3] = a;
                r.w[4] = b;
                r.w[5] = r.w[6] = r.w[7] = NONE;
                std::vector<uint32_t> rd = {a, a + 1, a + 2, b, b + 1, b + 2}, wr = {out, out + 1, out + 2};
                place_heavy(r, rd, wr);
                next.push_back(out);
            }
            for (size_t i = 2 * n_pairs; i < points.size(); ++i) next.push_back(points[i]);
            points.swap(next);
        }
    }
    // affine(points[0] (+ points[1])) -> out_x (, out_y); flags carry GF_OUT_CHECK / GF_OUT2_CHECK for witness outputs
    void curve_final(uint32_t idx, const std::vector<uint32_t>& points, uint32_t out_x, uint32_t out_y, uint32_t flags,
                     bool validate_fixed_base, uint32_t lo, uint32_t hi) {
        OpRec r{};
        r.w[0] = MK_JAC_FINAL | (flags << 8);
        r.w[1] = idx;
        r.w[2] = out_x;
        r.w[3] = points[0];
        r.w[4] = points.size() > 1 ? points[1] : NONE;
        r.w[5] = out_y;
        r.w[6] = validate_fixed_base ? lo : NONE;
        r.w[7] = validate_fixed_base ? hi : NONE;
        r.c[0][0] = validate_fixed_base ? 1u : 0u;
        std::vector<uint32_t> rd, wr = {out_x};
        for (uint32_t pt : points) { rd.push_back(pt); rd.push_back(pt + 1); rd.push_back(pt + 2); }
        if (validate_fixed_base) { rd.push_back(lo); rd.push_back(hi); }
        if (flags & GF_OUT_CHECK) rd.push_back(out_x);
        if (out_y != NONE) {
            if (flags & GF_OUT2_CHECK) rd.push_back(out_y);
            wr.push_back(out_y);
        }
        place_heavy(r, rd, wr);
    }

    // ---- packed hash pipeline ------------------------------------------------------------------------------------------
    // digest column of an earlier packed hash call, keyed by its first output witness
    struct DigestCol {
        uint32_t slot;
        std::vector<uint32_t> outs;
    };
    std::map<uint32_t, uint32_t> sha_pad_tables_;   // message length -> payload offset of the pad block's K + W table
    std::vector<std::pair<uint32_t, DigestCol>> digest_cols_;   // (first output, column), searched from the back (recent first)

    // returns false when the call does not qualify (the caller then emits the one-micro-op form)
    bool hash_packed(uint32_t idx, const BlackBoxCall& b) {
        if (!opt.packed_hashes) return false;
        const uint32_t func = b.func == BB_SHA256 ? 0u : b.func == BB_Keccak256 ? 1u : b.func == BB_Blake2s ? 2u : 3u;
        if (func > 2) return false;
        const uint32_t n = b.n_message_inputs;
        if (n == 0 || n > 4096) return false;
        if (func == 1 && n > 135) return false;            // one Keccak block: the core indexes its words statically
        for (uint32_t k = 0; k < n; ++k)
            if (b.inputs[k].num_bits == 0 || b.inputs[k].num_bits > 8) return false;
        for (uint32_t i = 0; i < 32; ++i)                   // 32 distinct outputs, none of them an input of this call
            for (uint32_t j = 0; j < i; ++j)
                if (b.outputs[i] == b.outputs[j]) return false;
        const uint32_t n_chunks = (n + 31) / 32;
        std::vector<uint32_t> chunks, packs;
        for (uint32_t c = 0; c < n_chunks; ++c) {
            const uint32_t cnt = std::min(32u, n - 32 * c);
            uint32_t slot = NONE;
            if (cnt == 32) {   // exactly the digest of an earlier call, in order: take its packed column
                for (size_t d = digest_cols_.size(); d-- > 0 && slot == NONE;) {
                    if (digest_cols_[d].first != b.inputs[32 * c].witness) continue;
                    bool same = true;
                    for (uint32_t k = 0; k < 32 && same; ++k) same = digest_cols_[d].second.outs[k] == b.inputs[32 * c + k].witness;
                    if (same) slot = digest_cols_[d].second.slot;
                }
            }
            if (slot == NONE) {
                slot = new_pinned(1, 64);
                packs.push_back(slot);
                OpRec r{};
                const uint32_t off = (uint32_t)plan.payload.size();
                plan.payload.push_back(cnt);
                std::vector<uint32_t> rd, wr = {slot};
                for (uint32_t k = 0; k < cnt; ++k) {
                    plan.payload.push_back(b.inputs[32 * c + k].witness);
                    rd.push_back(b.inputs[32 * c + k].witness);
                }
                r.w[0] = MK_HASH_PACK;
                r.w[1] = idx;
                r.w[2] = slot;
                r.w[3] = r.w[4] = r.w[5] = r.w[6] = NONE;
                r.w[7] = off;
                place_heavy(r, rd, wr);
            }
            chunks.push_back(slot);
        }
        const uint32_t dslot = new_pinned(1, 1u << 30);   // lives as long as a later call may take it as its message: never reused
        {
            OpRec r{};
            std::vector<uint32_t> rd(chunks), wr = {dslot};
            r.w[0] = MK_HASH_CORE;
            r.w[1] = idx;
            r.w[2] = dslot;
            r.w[3] = r.w[4] = r.w[5] = NONE;
            if (opt.sha_pad_table && func == 0 && n % 64 == 0) {
                // SHA-256 over a whole number of blocks: the last block is padding only and its K + W table is a constant
                auto it = sha_pad_tables_.find(n);
                if (it == sha_pad_tables_.end()) {
                    while (plan.payload.size() % 4) plan.payload.push_back(0);   // read with 128-bit loads
                    uint32_t kw[64];
                    bvm::sha256_pad_block_kw(n, kw);
                    it = sha_pad_tables_.emplace(n, (uint32_t)plan.payload.size()).first;
                    plan.payload.insert(plan.payload.end(), kw, kw + 64);
                }
                r.w[5] = it->second;
            }
            // descriptor = func, n_bytes, n_chunks, chunk columns.  Up to 37 chunks it rides in the record's coefficient words
            // (already in shared memory when the micro-op starts: one L2 round trip less on the chain of a hash chain).
            if (3 + n_chunks <= 40) {
                uint32_t* d = &r.c[0][0];
                d[0] = func;
                d[1] = n;
                d[2] = n_chunks;
                for (uint32_t c = 0; c < n_chunks; ++c) d[3 + c] = chunks[c];
                r.w[6] = 1;
                r.w[7] = NONE;
            } else {
                r.w[6] = 0;
                r.w[7] = (uint32_t)plan.payload.size();
                plan.payload.push_back(func);
                plan.payload.push_back(n);
                plan.payload.push_back(n_chunks);
                plan.payload.insert(plan.payload.end(), chunks.begin(), chunks.end());
            }
            place_heavy(r, rd, wr);
        }
        for (uint32_t s_ : packs) release_pinned(s_, 1);
        {
            OpRec r{};
            const uint32_t off = (uint32_t)plan.payload.size();
            uint32_t mask = 0;
            for (uint32_t i = 0; i < 32; ++i)
                if (known[b.outputs[i]]) mask |= 1u << i;
            plan.payload.push_back(mask);
            std::vector<uint32_t> rd = {dslot}, wr;
            for (uint32_t i = 0; i < 32; ++i) {
                plan.payload.push_back(b.outputs[i]);
                if (mask & (1u << i)) rd.push_back(b.outputs[i]);
                wr.push_back(b.outputs[i]);
            }
            r.w[0] = MK_HASH_UNPACK;
            r.w[1] = idx;
            r.w[2] = NONE;
            r.w[3] = dslot;
            r.w[4] = r.w[5] = r.w[6] = NONE;
            r.w[7] = off;
            place_heavy(r, rd, wr);
        }
        digest_cols_.push_back({b.outputs[0], DigestCol{dslot, std::vector<uint32_t>(b.outputs.begin(), b.outputs.begin() + 32)}});
        if (digest_cols_.size() > 4096) digest_cols_.erase(digest_cols_.begin(), digest_cols_.begin() + 2048);   // recent calls only
        for (uint32_t i = 0; i < 32; ++i)
            if (!known[b.outputs[i]]) mark_assigned(b.outputs[i], idx);
        plan.stats.alg_bytes += 32ull * (n + 32);
        ++plan.stats.n_hash;
        return true;
    }

    bool blackbox(uint32_t idx, const BlackBoxCall& b) {
        for (uint32_t w : b.outputs)
            if (known[w] == W_MAYBE && b.func != BB_RecursiveAggregation)
                throw std::runtime_error("opcode " + std::to_string(idx) + ": blackbox output witness " + std::to_string(w) +
                                         " is only conditionally assigned by an earlier value-dependent gate; not supported yet");
        for (auto& in : b.inputs)
            if (!known[in.witness]) {  // blackbox/mod.rs:55-62
                fail_static(idx, EK_MISSING_ASSIGNMENT, in.witness, "missing assignment for witness index " + std::to_string(in.witness));
                return false;
            }
        require_assigned(idx, b.inputs);
        OpRec r{};
        uint32_t reads[3];
        size_t nr = 0;
        uint32_t writes[2];
        size_t nw = 0;
        switch (b.func) {
            case BB_AND:
            case BB_XOR: {
                if (b.inputs[0].num_bits != b.inputs[1].num_bits) {
                    fail_static(idx, EK_REFERENCE_PANIC, 0, "number of bits specified for each input must be the same");
                    return false;
                }
                uint32_t out = b.outputs[0];
                uint32_t flags = 0;
                reads[nr++] = b.inputs[0].witness;
                reads[nr++] = b.inputs[1].witness;
                if (known[out]) {
                    flags |= GF_OUT_CHECK;
                    reads[nr++] = out;
                }
                writes[nw++] = out;
                r.w[0] = (b.func == BB_AND ? MK_AND : MK_XOR) | (flags << 8);
                r.w[1] = idx;
                r.w[2] = out;
                r.w[3] = b.inputs[0].witness;
                r.w[4] = b.inputs[1].witness;
                r.w[5] = r.w[6] = NONE;
                r.w[7] = b.inputs[0].num_bits;
                sched.place(r, reads, nr, writes, nw);
                if (!known[out]) mark_assigned(out, idx);
                ++plan.stats.n_micro;
                ++plan.stats.n_logic;
                plan.stats.alg_bytes += 96;
                return true;
            }
            case BB_RANGE: {
                reads[nr++] = b.inputs[0].witness;
                r.w[0] = MK_RANGE;
                r.w[1] = idx;
                r.w[2] = NONE;
                r.w[3] = b.inputs[0].witness;
                r.w[4] = r.w[5] = r.w[6] = NONE;
                r.w[7] = b.inputs[0].num_bits;
                sched.place(r, reads, nr, writes, nw);
                ++plan.stats.n_micro;
                ++plan.stats.n_range;
                plan.stats.alg_bytes += 32;
                return true;
            }
            case BB_HashToField128Security: {
                for (uint32_t k = 0; k < b.n_message_inputs; ++k)
                    if (b.inputs[k].num_bits > 256) {
                        fail_static(idx, EK_REFERENCE_PANIC, 0, "hash input wider than 256 bits");
                        return false;
                    }
                std::vector<uint32_t> rd, wr;
                uint32_t off = (uint32_t)plan.payload.size();
                uint32_t out = b.outputs[0];
                plan.payload.push_back(b.n_message_inputs);
                plan.payload.push_back(known[out] ? 1u : 0u);
                plan.payload.push_back(NONE);
                plan.payload.push_back([&] {   // 1: every message input is a single byte (num_bits 1..8) and no length cut: the device takes its packed-word path
                    if (b.func == BB_Keccak256VariableLength) return 0u;
                    for (uint32_t k = 0; k < b.n_message_inputs; ++k)
                        if (b.inputs[k].num_bits == 0 || b.inputs[k].num_bits > 8) return 0u;
                    return 1u;
                }());
                for (uint32_t k = 0; k < b.n_message_inputs; ++k) {
                    plan.payload.push_back(b.inputs[k].witness);
                    plan.payload.push_back(b.inputs[k].num_bits);
                    rd.push_back(b.inputs[k].witness);
                    plan.stats.alg_bytes += 32;
                }
                plan.payload.push_back(out);
                if (known[out]) rd.push_back(out);
                wr.push_back(out);
                r.w[0] = MK_HASH_TO_FIELD;
                r.w[1] = idx;
                r.w[2] = r.w[3] = r.w[4] = r.w[5] = r.w[6] = NONE;
                r.w[7] = off;
                place_heavy(r, rd, wr);
                if (!known[out]) mark_assigned(out, idx);
                ++plan.stats.n_hash;
                plan.stats.alg_bytes += 32;
                return true;
            }
            case BB_SHA256:
            case BB_Blake2s:
            case BB_Keccak256:
            case BB_Keccak256VariableLength: {
                if (b.outputs.size() != 32) {  // hash.rs:39-44
                    fail_static(idx, EK_BLACKBOX_FAILED, b.func, "Expected 32 outputs but encountered " + std::to_string(b.outputs.size()));
                    return false;
                }
                for (uint32_t k = 0; k < b.n_message_inputs; ++k)
                    if (b.inputs[k].num_bits > 256) {  // fetch_nearest_bytes slices past 32 bytes: the reference panics
                        fail_static(idx, EK_REFERENCE_PANIC, 0, "hash input wider than 256 bits");
                        return false;
                    }
                if (b.func != BB_Keccak256VariableLength && hash_packed(idx, b)) return true;
                std::vector<uint32_t> rd, wr;
                uint32_t off = (uint32_t)plan.payload.size();
                uint32_t mask = 0;
                for (uint32_t i = 0; i < 32; ++i)
                    if (known[b.outputs[i]]) mask |= 1u << i;
                plan.payload.push_back(b.n_message_inputs);
                plan.payload.push_back(mask);
                plan.payload.push_back(b.func == BB_Keccak256VariableLength ? b.inputs.back().witness : NONE);
                plan.payload.push_back([&] {   // 1: every message input is a single byte (num_bits 1..8) and no length cut: the device takes its packed-word path
                    if (b.func == BB_Keccak256VariableLength) return 0u;
                    for (uint32_t k = 0; k < b.n_message_inputs; ++k)
                        if (b.inputs[k].num_bits == 0 || b.inputs[k].num_bits > 8) return 0u;
                    return 1u;
                }());
                for (uint32_t k = 0; k < b.n_message_inputs; ++k) {
                    plan.payload.push_back(b.inputs[k].witness);
                    plan.payload.push_back(b.inputs[k].num_bits);
                    rd.push_back(b.inputs[k].witness);
                    plan.stats.alg_bytes += 32;
                }
                if (b.func == BB_Keccak256VariableLength) rd.push_back(b.inputs.back().witness);
                for (uint32_t i = 0; i < 32; ++i) {
                    plan.payload.push_back(b.outputs[i]);
                    if (mask & (1u << i)) rd.push_back(b.outputs[i]);
                    wr.push_back(b.outputs[i]);
                    plan.stats.alg_bytes += 32;
                }
                r.w[0] = (b.func == BB_SHA256 ? MK_SHA256 : b.func == BB_Blake2s ? MK_BLAKE2S : MK_KECCAK256) | (GF_HEAVY << 8);
                r.w[1] = idx;
                r.w[2] = r.w[3] = r.w[4] = r.w[5] = r.w[6] = NONE;
                r.w[7] = off;
                sched.place(r, rd.data(), rd.size(), wr.data(), wr.size());
                for (uint32_t i = 0; i < 32; ++i)
                    if (!known[b.outputs[i]]) mark_assigned(b.outputs[i], idx);
                plan.needs_full_kernel = true;
                ++plan.stats.n_micro;
                ++plan.stats.n_hash;
                return true;
            }
            case BB_FixedBaseScalarMul: {
                uint32_t ox = b.outputs[0], oy = b.outputs[1];
                uint32_t flags = GF_HEAVY;
                if (split_curve_ops()) {
                    // s*G = sum over 32 8-bit windows of table points: 8 partial sums of 4 windows, a 3-level addition tree
                    // whose last addition rides in the finaliser (which also validates the limbs, scalar_mul.rs:25-51)
                    if (known[ox]) flags |= GF_OUT_CHECK;
                    if (known[oy] || oy == ox) flags |= GF_OUT2_CHECK;
                    std::vector<uint32_t> pts;
                    curve_parts(idx, 0, b.inputs[0].witness, 0, 0, 16, 4, 0, pts);
                    curve_parts(idx, 0, b.inputs[1].witness, 0, 0, 16, 4, 16, pts);
                    curve_reduce(idx, pts, 2);
                    curve_final(idx, pts, ox, oy, flags, true, b.inputs[0].witness, b.inputs[1].witness);
                    if (!known[ox]) mark_assigned(ox, idx);
                    if (!known[oy]) mark_assigned(oy, idx);
                    ++plan.stats.n_curve;
                    plan.stats.alg_bytes += 128;
                    return true;
                }
                std::vector<uint32_t> rd = {b.inputs[0].witness, b.inputs[1].witness}, wr;
                if (known[ox]) { flags |= GF_OUT_CHECK; rd.push_back(ox); }
                wr.push_back(ox);
                if (known[oy] || oy == ox) { flags |= GF_OUT2_CHECK; rd.push_back(oy); }
                if (oy != ox) wr.push_back(oy);
                r.w[0] = MK_FIXED_BASE | (flags << 8);
                r.w[1] = idx;
                r.w[2] = ox;
                r.w[3] = b.inputs[0].witness;
                r.w[4] = b.inputs[1].witness;
                r.w[5] = oy;
                r.w[6] = NONE;
                r.w[7] = 0;
                sched.place(r, rd.data(), rd.size(), wr.data(), wr.size());
                if (!known[ox]) mark_assigned(ox, idx);
                if (!known[oy]) mark_assigned(oy, idx);
                plan.needs_full_kernel = true;
                ++plan.stats.n_micro;
                ++plan.stats.n_curve;
                plan.stats.alg_bytes += 128;
                return true;
            }
            case BB_Pedersen: {
                if (!opt.allow_unpinned_pedersen)
                    throw std::runtime_error("opcode " + std::to_string(idx) + ": BlackBoxFuncCall::Pedersen is not supported: barretenberg's "
                                             "generator tables are not reproducible here, results would differ from the reference "
                                             "(opt in with the context option pedersen_unpinned=1 to run the structurally identical kernel)");
                uint32_t ox = b.outputs[0], oy = b.outputs[1];
                uint32_t flags = GF_HEAVY;
                if (split_curve_ops() && !b.inputs.empty()) {
                    // r_0 = IV; r_{k+1} = (H0(r_k) + H1(v_k)).x; out = H0(r_n) + H1(n)   (oracle/pedersen.py).  Each H is a sum of
                    // 29 table points -> 8 partial sums.  Every H1 is independent of the chain and is reduced to one point ahead
                    // of it; a chaining round is then 8 partial sums, a 3-level addition tree and one finaliser.
                    if (known[ox]) flags |= GF_OUT_CHECK;
                    if (known[oy] || oy == ox) flags |= GF_OUT2_CHECK;
                    // H0(IV) and H1(n) hash plan-time constants: the same point for every instance and every call with that
                    // domain separator / input count, so each is computed ONCE per plan (const_curve_point) and kept.
                    const uint32_t n_in = (uint32_t)b.inputs.size();
                    std::vector<uint32_t> h1(n_in + 1);
                    for (uint32_t k = 0; k < n_in; ++k) {
                        std::vector<uint32_t> pts;
                        curve_parts(idx, 0, b.inputs[k].witness, 0, 1, 29, 4, 29, pts);   // num_bits is ignored (pedersen.rs:18-20)
                        h1[k] = new_pinned(3, 1024);   // consumed n_in + 1 - k chaining rounds later: not a pool slot
                        curve_reduce(idx, pts, 1, h1[k]);
                    }
                    h1[n_in] = const_curve_point(idx, 2, n_in, 29);   // the length block: an immediate scalar
                    uint32_t chain = NONE;   // slot holding r_k (canonical x of the previous round)
                    for (uint32_t k = 0; k <= n_in; ++k) {
                        std::vector<uint32_t> pts;
                        if (k == 0) pts.push_back(const_curve_point(idx, 1, b.domain_separator, 0));
                        else curve_parts(idx, 0, chain, 0, 1, 29, 4, 0, pts);
                        pts.push_back(h1[k]);
                        curve_reduce(idx, pts, 2);
                        if (k < n_in) {
                            chain = new_temp();
                            curve_final(idx, pts, chain, NONE, GF_HEAVY, false, NONE, NONE);
                            release_pinned(h1[k], 3);   // read for the last time by this round's finaliser
                        } else {
                            curve_final(idx, pts, ox, oy, flags, false, NONE, NONE);
                        }
                    }
                    if (!known[ox]) mark_assigned(ox, idx);
                    if (!known[oy]) mark_assigned(oy, idx);
                    ++plan.stats.n_curve;
                    plan.stats.alg_bytes += 32 * b.inputs.size() + 64;
                    return true;
                }
                std::vector<uint32_t> rd, wr;
                uint32_t off = (uint32_t)plan.payload.size();
                plan.payload.push_back((uint32_t)b.inputs.size());
                plan.payload.push_back(b.domain_separator);
                for (auto& in : b.inputs) {   // num_bits is ignored: the full field value is hashed (pedersen.rs:18-20)
                    plan.payload.push_back(in.witness);
                    rd.push_back(in.witness);
                }
                if (known[ox]) { flags |= GF_OUT_CHECK; rd.push_back(ox); }
                wr.push_back(ox);
                if (known[oy] || oy == ox) { flags |= GF_OUT2_CHECK; rd.push_back(oy); }
                if (oy != ox) wr.push_back(oy);
                r.w[0] = MK_PEDERSEN | (flags << 8);
                r.w[1] = idx;
                r.w[2] = ox;
                r.w[3] = r.w[4] = r.w[6] = NONE;
                r.w[5] = oy;
                r.w[7] = off;
                sched.place(r, rd.data(), rd.size(), wr.data(), wr.size());
                if (!known[ox]) mark_assigned(ox, idx);
                if (!known[oy]) mark_assigned(oy, idx);
                plan.needs_full_kernel = true;
                ++plan.stats.n_micro;
                ++plan.stats.n_curve;
                plan.stats.alg_bytes += 32 * b.inputs.size() + 64;
                return true;
            }
            case BB_EcdsaSecp256k1:
            case BB_EcdsaSecp256r1: {
                // signature/ecdsa.rs:12-97.  seg[] = sizes of public_key_x, public_key_y, signature, hashed_message.
                static const char* label[3] = {"pubkey_x", "pubkey_y", "signature"};
                static const uint32_t want[3] = {32, 32, 64};
                for (int k = 0; k < 3; ++k)
                    if (b.seg[k] != want[k]) {
                        fail_static(idx, EK_BLACKBOX_FAILED, b.func,
                                    std::string("expected ") + label[k] + " size " + std::to_string(want[k]) + " but received " +
                                        std::to_string(b.seg[k]));
                        return false;
                    }
                if (b.seg[3] != 32) {  // GenericArray::from_slice(hashed_msg) asserts the length (blackbox_solver/src/lib.rs:127)
                    fail_static(idx, EK_REFERENCE_PANIC, 0, "hashed message is not 32 bytes");
                    return false;
                }
                uint32_t out = b.outputs[0];
                uint32_t flags = GF_HEAVY;
                std::vector<uint32_t> rd, wr;
                uint32_t off = (uint32_t)plan.payload.size();
                plan.payload.push_back(b.func == BB_EcdsaSecp256k1 ? 0u : 1u);
                for (auto& in : b.inputs) {   // pkx[32] pky[32] sig[64] hashed_message[32]; num_bits unused (signature/mod.rs:5-18)
                    plan.payload.push_back(in.witness);
                    rd.push_back(in.witness);
                }
                if (known[out]) { flags |= GF_OUT_CHECK; rd.push_back(out); }
                wr.push_back(out);
                r.w[0] = MK_ECDSA | (flags << 8);
                r.w[1] = idx;
                r.w[2] = out;
                r.w[3] = r.w[4] = r.w[5] = r.w[6] = NONE;
                r.w[7] = off;
                place_heavy(r, rd, wr);
                if (!known[out]) mark_assigned(out, idx);
                ++plan.stats.n_curve;
                plan.stats.alg_bytes += 32 * b.inputs.size() + 32;
                return true;
            }
            case BB_RecursiveAggregation: {
                // blackbox/mod.rs:154-161: every output witness := 0 (insert_value semantics); the proof is the backend's job
                for (uint32_t w : b.outputs) {
                    if (known[w] == W_MAYBE) {   // per-lane presence: same rule as the gate `w = 0`
                        Expression e;
                        e.linear_combinations.push_back({hf::from_u64(1), w});
                        if (!arithmetic(idx, e)) return false;
                        continue;
                    }
                    lower_sum({}, {}, U256{}, /*assign=*/true, w, idx, /*out_check=*/known[w] != 0);
                    if (!known[w]) mark_assigned(w, idx);
                }
                return true;
            }
            default:
                throw std::runtime_error("opcode " + std::to_string(idx) + ": blackbox function " + blackbox_name(b.func) +
                                         " is not supported by the device plan yet");
        }
    }

    void run(const std::vector<uint32_t>& inputs) {
        plan.S = opt.S;
        plan.chunk_steps = opt.chunk_steps;
        plan.num_witnesses = (uint32_t)known.size();
        plan.n_opcodes = (uint32_t)c.opcodes.size();
        plan.input_witnesses = inputs;
        plan.assign_opcode.assign(known.size(), ASSIGN_NEVER);
        plan.mu_index_of.assign(known.size(), NONE);
        scaled = opt.scaled_columns;
        if (scaled) {
            // columns that something other than an arithmetic gate reads (or compares) directly stay canonical
            need_canon.assign(known.size(), 0);
            auto canon = [&](uint32_t w) { need_canon[w] = 1; };
            for (auto& op : c.opcodes) {
                switch (op.kind) {
                    case OP_BlackBox:
                        for (auto& in : op.bb().inputs) canon(in.witness);
                        for (uint32_t w : op.bb().outputs) canon(w);
                        break;
                    case OP_Directive:
                        canon(op.dir().q);
                        canon(op.dir().r);
                        for (uint32_t w : op.dir().out) canon(w);
                        break;
                    case OP_MemoryInit:
                        for (uint32_t w : op.init()) canon(w);
                        break;
                    case OP_MemoryOp:   // a read lands in the witness of `value` (memory_op.rs:89-101)
                        for (auto& t : op.mem().value.linear_combinations) canon(t.w);
                        break;
                    case OP_Brillig:
                        for (auto& out : op.brillig().outputs)
                            for (uint32_t w : out.witnesses) canon(w);
                        break;
                    default:
                        break;
                }
            }
        }
        plan.input_scaled.assign(inputs.size(), 0);
        for (size_t i = 0; i < inputs.size(); ++i) {
            const uint32_t w = inputs[i];
            known[w] = 1;
            plan.assign_opcode[w] = ASSIGN_INPUT;
            if (scaled && !need_canon[w]) {   // Montgomery form: products of inputs with unit coefficients stay uniformly scaled
                set_scale(w, hf::consts().R, hf::from_mont(hf::consts().one));
                plan.input_scaled[i] = 1;
            }
        }
        for (uint32_t i = 0; i < c.opcodes.size(); ++i) {
            const Opcode& op = c.opcodes[i];
            bool ok = true;
            switch (op.kind) {
                case OP_Arithmetic:
                    ok = arithmetic(i, op.expr);
                    break;
                case OP_BlackBox:
                    ok = blackbox(i, op.bb());
                    break;
                case OP_Directive:
                    ok = directive(i, op.dir());
                    break;
                case OP_Brillig:
                    ok = brillig(i, op.brillig());
                    break;
                case OP_MemoryInit:
                    ok = memory_init(i, op.block_id(), op.init());
                    break;
                case OP_MemoryOp:
                    ok = memory_op(i, op.mem());
                    break;
                default:
                    throw std::runtime_error("opcode " + std::to_string(i) + ": opcode kind " + std::to_string(op.kind) +
                                             " (Brillig) is not supported by the device plan yet");
            }
            if (!ok) break;
        }
        close_device_segment();
        plan.ring_slots = opt.ring_slots;
        sched.assign_ring(opt.ring_slots, plan.stats.n_operand_reads, plan.stats.n_ring_reads);
        {
            // tile width the runtime will pick (runtime.cu pick_T): 32 lanes for circuits with curve calls, else 128 / S
            uint32_t T = opt.tile_lanes ? opt.tile_lanes : (plan.stats.n_curve ? 32u : std::max(1u, 128u / opt.S));
            if (T > 32) T = 32;
            sched.emit(plan.stream, plan.n_steps, plan.chunk_steps, opt.spread_heavy && (32 % T) == 0 ? 32 / T : 1);
        }
        plan.n_slots = temp_base + opt.temp_pool + extra_slots;
        if (scaled) {
            bool any = false;
            for (uint32_t w = 0; w < plan.num_witnesses && !any; ++w) any = !is_canonical(w);
            if (any) {
                plan.unscale.resize((size_t)plan.num_witnesses * 8);
                for (uint32_t w = 0; w < plan.num_witnesses; ++w) put(&plan.unscale[(size_t)w * 8], hf::to_mont(mu_of(w)));
            } else {
                std::fill(plan.input_scaled.begin(), plan.input_scaled.end(), 0u);
            }
        }
        plan.stats.n_opcodes = c.opcodes.size();
        plan.stats.n_steps = sched.n_steps();
        plan.stats.n_slots_filled = sched.n_ops();
    }
};

uint32_t witness_span(const Circuit& c, const std::vector<uint32_t>& inputs) {
    uint32_t m = c.current_witness_index;
    auto upd = [&](uint32_t w) { m = std::max(m, w); };
    auto expr = [&](const Expression& e) {
        for (auto& t : e.mul_terms) {
            upd(t.a);
            upd(t.b);
        }
        for (auto& t : e.linear_combinations) upd(t.w);
    };
    for (uint32_t w : inputs) upd(w);
    for (auto& op : c.opcodes) {
        switch (op.kind) {
            case OP_Arithmetic:
                expr(op.expr);
                break;
            case OP_BlackBox:
                for (auto& in : op.bb().inputs) upd(in.witness);
                for (uint32_t w : op.bb().outputs) upd(w);
                break;
            case OP_Directive:
                expr(op.dir().a);
                expr(op.dir().b);
                if (op.dir().predicate.present) expr(op.dir().predicate);
                upd(op.dir().q);
                upd(op.dir().r);
                for (uint32_t w : op.dir().out) upd(w);
                break;
            case OP_MemoryInit:
                for (uint32_t w : op.init()) upd(w);
                break;
            case OP_Brillig:
                for (auto& in : op.brillig().inputs)
                    for (auto& e : in.exprs) expr(e);
                for (auto& out : op.brillig().outputs)
                    for (uint32_t w : out.witnesses) upd(w);
                if (op.brillig().predicate.present) expr(op.brillig().predicate);
                break;
            case OP_MemoryOp:
                expr(op.mem().operation);
                expr(op.mem().index);
                expr(op.mem().value);
                if (op.mem().predicate.present) expr(op.mem().predicate);
                break;
            default:
                break;
        }
    }
    return m + 1;
}

}  // namespace

// all-at-once inversion (Montgomery's trick): 3 multiplications per element + one real inversion per chunk, chunks in parallel
static std::vector<U256> batch_inverse(const std::vector<U256>& v) {
    std::vector<U256> out(v.size());
    const size_t n = v.size();
    unsigned n_thr = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)(n / 4096 + 1)));
    auto work = [&](size_t lo, size_t hi) {
        if (lo >= hi) return;
        std::vector<U256> prefix(hi - lo);
        U256 acc = hf::consts().R;   // Montgomery one
        for (size_t i = lo; i < hi; ++i) {   // zero stays zero (inverse(0) = 0) and is skipped in the running product
            prefix[i - lo] = acc;
            if (!v[i].is_zero()) acc = hf::mont_mul(acc, hf::to_mont(v[i]));
        }
        U256 inv_acc = hf::to_mont(hf::inverse(hf::from_mont(acc)));
        for (size_t i = hi; i-- > lo;) {
            if (v[i].is_zero()) continue;
            out[i] = hf::from_mont(hf::mont_mul(inv_acc, prefix[i - lo]));
            inv_acc = hf::mont_mul(inv_acc, hf::to_mont(v[i]));
        }
    };
    std::vector<std::thread> th;
    size_t per = (n + n_thr - 1) / n_thr;
    for (unsigned t = 1; t < n_thr; ++t) th.emplace_back(work, t * per, std::min(n, (t + 1) * per));
    work(0, std::min(n, per));
    for (auto& x : th) x.join();
    return out;
}

static Plan compile_plan_once(const Circuit& c, const std::vector<uint32_t>& input_witnesses, const PlanOptions& opt, uint32_t nw) {
    if (c.opcodes.size() < 2048) {   // small circuits: direct inversions
        Compiler comp(c, opt, nw);
        comp.run(input_witnesses);
        return std::move(comp.plan);
    }
    std::vector<U256> requests;
    {
        Compiler dry(c, opt, nw);
        dry.inv_record = &requests;
        dry.sched.dry = true;
        dry.run(input_witnesses);
    }
    std::vector<U256> inverses = batch_inverse(requests);
    Compiler comp(c, opt, nw);
    comp.inv_replay = &inverses;
    comp.run(input_witnesses);
    if (comp.inv_pos != inverses.size()) throw std::runtime_error("plan: inversion replay out of sync");
    return std::move(comp.plan);
}

Plan compile_plan(const Circuit& c, const std::vector<uint32_t>& input_witnesses, const PlanOptions& opt_in) {
    PlanOptions opt = opt_in;
    if (opt.S == 0 || opt.S > 64) throw std::runtime_error("plan: S must be in 1..64");
    uint32_t nw = witness_span(c, input_witnesses);
    // Jacobian points of the split curve micro-ops travel through temporaries (3 slots each, ~270 per Pedersen): a larger
    // round-robin pool keeps slot reuse (a false WAR dependency) from serialising independent curve operations
    if (opt.split_curve && opt.S >= 8)
        for (auto& op : c.opcodes)
            if (op.kind == OP_BlackBox && (op.bb().func == BB_Pedersen || op.bb().func == BB_FixedBaseScalarMul)) {
                opt.temp_pool = std::max<uint32_t>(opt.temp_pool, 32768);
                break;
            }
    try {
        return compile_plan_once(c, input_witnesses, opt, nw);
    } catch (const NeedsCanonicalColumns&) {
        // a value-dependent gate (arithmetic.rs:217-221) reads a column that an earlier gate left scaled: such circuits keep
        // every column canonical
        opt.scaled_columns = false;
        return compile_plan_once(c, input_witnesses, opt, nw);
    }
}

// ---- (de)serialisation: one flat blob for the multi-GPU broadcast -----------------------------
namespace {
template <typename T>
void put_pod(std::vector<uint8_t>& b, const T& v) {
    const uint8_t* p = (const uint8_t*)&v;
    b.insert(b.end(), p, p + sizeof(T));
}
template <typename T>
void put_vec(std::vector<uint8_t>& b, const std::vector<T>& v) {
    put_pod<uint64_t>(b, v.size());
    const uint8_t* p = (const uint8_t*)v.data();
    b.insert(b.end(), p, p + v.size() * sizeof(T));
    while (b.size() % 16) b.push_back(0);
}
struct Cursor {
    const uint8_t* d;
    size_t n, o = 0;
    template <typename T>
    T pod() {
        if (o + sizeof(T) > n) throw std::runtime_error("plan blob truncated");
        T v;
        memcpy(&v, d + o, sizeof(T));
        o += sizeof(T);
        return v;
    }
    template <typename T>
    std::vector<T> vec() {
        uint64_t k = pod<uint64_t>();
        if (o + k * sizeof(T) > n) throw std::runtime_error("plan blob truncated");
        std::vector<T> v(k);
        memcpy(v.data(), d + o, k * sizeof(T));
        o += k * sizeof(T);
        while (o % 16) ++o;
        return v;
    }
};
constexpr uint64_t kPlanMagic = 0x3430304e414c5042ULL;  // "BPLAN004"
}  // namespace

std::vector<uint8_t> serialize_plan(const Plan& p) {
    std::vector<uint8_t> b;
    put_pod<uint64_t>(b, kPlanMagic);
    put_pod<uint32_t>(b, p.S);
    put_pod<uint32_t>(b, p.num_witnesses);
    put_pod<uint32_t>(b, p.n_slots);
    put_pod<uint32_t>(b, p.n_opcodes);
    put_pod<uint32_t>(b, p.chunk_steps);
    put_pod<uint32_t>(b, p.needs_full_kernel ? 1u : 0u);
    put_pod<uint32_t>(b, p.n_steps);
    put_pod<uint32_t>(b, p.static_fail.present ? 1u : 0u);
    put_pod<uint32_t>(b, p.static_fail.opcode);
    put_pod<uint32_t>(b, p.static_fail.kind);
    put_pod<uint32_t>(b, p.static_fail.aux);
    put_pod<uint32_t>(b, p.n_mu);
    put_pod<uint32_t>(b, p.ring_slots);
    put_pod<uint32_t>(b, 0u);
    put_pod<PlanStats>(b, p.stats);
    while (b.size() % 16) b.push_back(0);
    put_vec(b, p.input_witnesses);
    put_vec(b, p.input_scaled);
    put_vec(b, p.unscale);
    put_vec(b, p.assign_opcode);
    put_vec(b, p.mu_index_of);
    put_vec(b, p.payload);
    put_vec(b, p.segments);
    put_vec(b, p.host_desc);
    put_vec(b, p.acir_gz);
    put_vec(b, p.stream);
    return b;
}

Plan deserialize_plan(const uint8_t* data, size_t len) {
    Cursor c{data, len};
    if (c.pod<uint64_t>() != kPlanMagic) throw std::runtime_error("not a plan blob");
    Plan p;
    p.S = c.pod<uint32_t>();
    p.num_witnesses = c.pod<uint32_t>();
    p.n_slots = c.pod<uint32_t>();
    p.n_opcodes = c.pod<uint32_t>();
    p.chunk_steps = c.pod<uint32_t>();
    p.needs_full_kernel = c.pod<uint32_t>() != 0;
    p.n_steps = c.pod<uint32_t>();
    p.static_fail.present = c.pod<uint32_t>() != 0;
    p.static_fail.opcode = c.pod<uint32_t>();
    p.static_fail.kind = c.pod<uint32_t>();
    p.static_fail.aux = c.pod<uint32_t>();
    p.n_mu = c.pod<uint32_t>();
    p.ring_slots = c.pod<uint32_t>();
    (void)c.pod<uint32_t>();
    p.stats = c.pod<PlanStats>();
    while (c.o % 16) ++c.o;
    p.input_witnesses = c.vec<uint32_t>();
    p.input_scaled = c.vec<uint32_t>();
    p.unscale = c.vec<uint32_t>();
    p.assign_opcode = c.vec<uint32_t>();
    p.mu_index_of = c.vec<uint32_t>();
    p.payload = c.vec<uint32_t>();
    p.segments = c.vec<Segment>();
    p.host_desc = c.vec<uint32_t>();
    p.acir_gz = c.vec<uint8_t>();
    p.stream = c.vec<OpRec>();
    if (p.stream.size() != (size_t)p.n_steps * p.S) throw std::runtime_error("plan blob: stream size mismatch");
    return p;
}

}  // namespace acvmb
