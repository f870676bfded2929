// Host runtime behind the C ABI (include/acvm_b200.h): context, compiled circuit, device-resident
// batch, sub-batching, I/O staging.  Mirrors the ownership model of the reference's ACVM struct
// (acvm/src/pwg/mod.rs:129-181): the circuit owns the opcodes/plan, a batch owns the witness columns.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/acvm_b200.h"
#include "acir.hpp"
#include "brillig_host.hpp"
#include "sort_host.hpp"
#include "curve_host.hpp"
#include "plan.hpp"
#include "vm_kernel.cuh"

using namespace acvmb;

static thread_local std::string g_last_error;
static int set_err(int rc, const std::string& msg) {
    g_last_error = msg;
    return rc;
}
#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return set_err(_e == cudaErrorMemoryAllocation ? ACVMB_ERR_OOM : ACVMB_ERR_CUDA,       \
                           std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
    } while (0)

struct acvmb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, gather_stream = nullptr;   // VM kernels | D2H copies | output gathers
    cudaDeviceProp prop{};
    uint32_t opt_T = 0;   // 0 = auto
    uint32_t opt_S = 0;   // 0 = auto: 16, or 8 for circuits with curve calls (see circuit_from_struct)
    uint32_t opt_chunk_steps = 4;   // steps per TMA stage (B200, full size: 4 is 2 % faster than 2 with the one-reduction gate kernel)
    uint32_t opt_split_curve = 1, opt_temp_pool = 0;
    uint32_t opt_slack_scheduling = 1;   // plan: curve micro-ops with slack fill idle slots of later levels
    uint32_t opt_spread_heavy = 1;   // plan: heavy micro-ops of one step go to different warps (tiles narrower than a warp)
    uint32_t opt_device_brillig = 1;
    uint32_t opt_scaled_columns = 1;
    uint32_t opt_packed_hashes = 1, opt_sha_pad_table = 1;
    // shared memory per CTA for the ring of recent values; 0 = no ring, the default: measured on B200 at full size the ring
    // is within 1 % of the plain kernel (operand loads hit L2 and four resident CTAs per SM hide that latency already)
    uint32_t opt_ring_bytes = 0;
    uint32_t opt_pedersen_unpinned = 0;   // 1: accept Pedersen opcodes / acvmb_pedersen although parity with barretenberg is unpinned
    int opt_split = -1;
    uint32_t opt_n_stage = 4;
    uint64_t max_resident_bytes = 0;  // 0 = auto (fraction of free memory)
    uint64_t staging_bytes = 512ull << 20;
    bool tables_ready = false;          // the per-device lookup tables (ensure_curve_tables) exist
    // acvmb_solve_batch keeps the column buffers of its last call (one per context): a caller that streams a large batch
    // through repeated calls otherwise pays a cudaMalloc + cudaFree of tens of GB per call (measured 20-290 ms)
    struct acvmb_batch* cached_batch = nullptr;
    struct acvmb_batch* pipe_batch[2] = {nullptr, nullptr};   // the two column buffers of the pipelined (solve | drain) path
    uint64_t cached_bytes = 0;
    uint32_t opt_cache_batch = 1;
    // multi-device context (acvmb_ctx_create_multi): this context is bound to devices[0], `peers` are the contexts of
    // devices[1..]; circuits created on it are replicated to every peer, acvmb_solve_batch shards the batch.
    std::vector<acvmb_ctx*> peers;
    void* nccl_comms = nullptr;      // ncclComm_t[1 + peers.size()], created with the first circuit
    const char* bcast_backend = "none";
};
static void drop_cached_batch(acvmb_ctx* ctx);
static acvmb_circuit* cached_batch_circuit(acvmb_ctx* ctx);

struct acvmb_circuit {
    acvmb_ctx* ctx = nullptr;
    Plan plan;
    Circuit circuit;            // kept only for plans with host (Brillig) segments
    bool has_circuit = false;
    uint8_t* d_stream = nullptr;
    uint32_t* d_payload = nullptr;
    uint32_t* d_assign = nullptr;
    uint32_t* d_input_slots = nullptr;
    uint32_t* d_mu_index_of = nullptr;
    uint32_t* d_unscale = nullptr;   // scaled columns: (1/lambda_w)*R per witness, for the output gather
    std::vector<acvmb_circuit*> shards;   // multi-device context: the replicas on the peer devices (owned)
    size_t stream_bytes = 0;
    acvmb_run_info run{};
    ~acvmb_circuit() {
        if (d_stream) cudaFree(d_stream);
        if (d_payload) cudaFree(d_payload);
        if (d_assign) cudaFree(d_assign);
        if (d_input_slots) cudaFree(d_input_slots);
        if (d_mu_index_of) cudaFree(d_mu_index_of);
        if (d_unscale) cudaFree(d_unscale);
    }
};
static void destroy_shards(acvmb_circuit* c);

struct acvmb_batch {
    acvmb_circuit* c = nullptr;
    uint32_t n_inst = 0, T = 0, n_tiles = 0;
    uint32_t capacity = 0;               // instances the buffers were sized for (n_inst <= capacity)
    uint4* d_cols = nullptr;
    unsigned long long* d_fail = nullptr;
    uint8_t* d_in = nullptr;
    std::vector<uint8_t*> d_staged_in;   // resident input sets (acvmb_batch_stage_inputs)
    uint8_t* d_stage[2] = {nullptr, nullptr};
    uint8_t* d_stage_present[2] = {nullptr, nullptr};
    size_t stage_present_bytes = 0;
    uint32_t* d_mu = nullptr;            // per-lane assignment table of value-dependent witnesses
    // grow-only staging of the host (Brillig) segments
    uint8_t* d_host_io = nullptr;
    uint32_t* d_host_ids = nullptr;
    size_t host_io_bytes = 0, host_ids_bytes = 0;
    std::vector<uint8_t> h_in, h_out;
    std::vector<unsigned long long> h_fail;
    // pending Brillig foreign call of instance 0 (kept for the single-instance ACVM mirror)
    bool fc_pending = false;
    std::string fc_function;
    std::vector<std::vector<U256>> fc_inputs;
    uint32_t* d_out_ids = nullptr;
    size_t out_ids_cap = 0;
    size_t stage_bytes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_gather[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_ready = nullptr, ev_dl_done = nullptr;   // columns final on the VM stream | last D2H of an async download
    bool dl_pending = false;
    ~acvmb_batch() {
        if (d_cols) cudaFree(d_cols);
        if (d_fail) cudaFree(d_fail);
        if (d_in) cudaFree(d_in);
        for (uint8_t* p : d_staged_in) if (p) cudaFree(p);
        if (d_stage[0]) cudaFree(d_stage[0]);
        if (d_stage[1]) cudaFree(d_stage[1]);
        if (d_stage_present[0]) cudaFree(d_stage_present[0]);
        if (d_stage_present[1]) cudaFree(d_stage_present[1]);
        if (d_mu) cudaFree(d_mu);
        if (d_host_io) cudaFree(d_host_io);
        if (d_host_ids) cudaFree(d_host_ids);
        if (d_out_ids) cudaFree(d_out_ids);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_ready) cudaEventDestroy(ev_ready);
        if (ev_dl_done) cudaEventDestroy(ev_dl_done);
        for (int i = 0; i < 2; ++i) {
            if (ev_gather[i]) cudaEventDestroy(ev_gather[i]);
            if (ev_copy[i]) cudaEventDestroy(ev_copy[i]);
        }
    }
};

struct acvmb_vm {
    acvmb_circuit* c = nullptr;
    acvmb_batch* b = nullptr;
    bool fc_pending = false;
    std::string fc_function;
    std::vector<std::vector<U256>> fc_inputs;
    std::vector<uint8_t> inputs;      // [n_initial][32]
    acvmb_status status{ACVMB_IN_PROGRESS, 0, 0, 0};
    std::vector<uint8_t> witness;     // dense [num_witnesses][32] after solve
    std::vector<uint8_t> present;     // dense [num_witnesses] after solve
    std::vector<uint32_t> assign;
    bool solved_once = false;
    // ACVM::solve_opcode (mod.rs:243-303) steps one opcode at a time.  The device executes a re-ordered schedule of the whole
    // circuit, so the mirror solves once and then walks `ip` over the already-solved state: a witness is visible once the
    // opcode that assigns it is behind ip, the recorded outcome is reported when ip reaches its opcode.
    acvmb_status final_status{ACVMB_IN_PROGRESS, 0, 0, 0};
    uint32_t ip = 0;
    std::vector<uint32_t> mu;         // instance 0 of the per-lane "assigned by opcode" table (value-dependent witnesses)
};

// ---------------------------------------------------------------------------------------------
extern "C" const char* acvmb_last_error(void) { return g_last_error.c_str(); }

extern "C" int acvmb_ctx_create(int device, acvmb_ctx** out) {
    if (!out) return set_err(ACVMB_ERR_INVALID_ARG, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_err(ACVMB_ERR_NO_DEVICE, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                                "); acvm_b200 has no CPU fallback");
    if (device < 0 || device >= n) return set_err(ACVMB_ERR_INVALID_ARG, "device index out of range");
    auto ctx = std::make_unique<acvmb_ctx>();
    ctx->device = device;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaGetDeviceProperties(&ctx->prop, device));
    if (ctx->prop.major < 10)
        return set_err(ACVMB_ERR_NO_DEVICE, std::string("device ") + ctx->prop.name + " is sm_" + std::to_string(ctx->prop.major) +
                                                std::to_string(ctx->prop.minor) + "; kernels are built for sm_100a only");
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->gather_stream, cudaStreamNonBlocking));
    *out = ctx.release();
    return ACVMB_OK;
}

static void destroy_nccl(acvmb_ctx* ctx);

extern "C" void acvmb_ctx_destroy(acvmb_ctx* ctx) {
    if (!ctx) return;
    destroy_nccl(ctx);
    for (acvmb_ctx* p : ctx->peers) acvmb_ctx_destroy(p);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    drop_cached_batch(ctx);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->gather_stream) cudaStreamDestroy(ctx->gather_stream);
    delete ctx;
}

extern "C" int acvmb_device_name(acvmb_ctx* ctx, char* buf, size_t len) {
    if (!ctx || !buf || !len) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    snprintf(buf, len, "%s", ctx->prop.name);
    return ACVMB_OK;
}

extern "C" void* acvmb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {   // pinned for every device of a multi-device context
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void acvmb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

extern "C" int acvmb_ctx_set_option(acvmb_ctx* ctx, const char* key, uint64_t value) {
    if (!ctx || !key) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    for (acvmb_ctx* p : ctx->peers) {
        int rc = acvmb_ctx_set_option(p, key, value);
        if (rc) return rc;
    }
    std::string k(key);
    if (k == "T") ctx->opt_T = (uint32_t)value;
    else if (k == "S") ctx->opt_S = (uint32_t)value;
    else if (k == "chunk_steps") ctx->opt_chunk_steps = (uint32_t)value;
    else if (k == "split_curve") ctx->opt_split_curve = (uint32_t)value;
    else if (k == "spread_heavy") ctx->opt_spread_heavy = value != 0;
    else if (k == "slack_scheduling") ctx->opt_slack_scheduling = value != 0;
    else if (k == "cache_batch") { ctx->opt_cache_batch = (uint32_t)value; if (!value) drop_cached_batch(ctx); }
    else if (k == "temp_pool") ctx->opt_temp_pool = (uint32_t)value;
    else if (k == "pedersen_unpinned") ctx->opt_pedersen_unpinned = value ? 1u : 0u;
    else if (k == "device_brillig") ctx->opt_device_brillig = value ? 1u : 0u;
    else if (k == "scaled_columns") ctx->opt_scaled_columns = value ? 1u : 0u;
    else if (k == "packed_hashes") ctx->opt_packed_hashes = value ? 1u : 0u;
    else if (k == "sha_pad_table") ctx->opt_sha_pad_table = value ? 1u : 0u;
    else if (k == "ring_bytes") ctx->opt_ring_bytes = (uint32_t)value;
    else if (k == "split") ctx->opt_split = (int)value;
    else if (k == "n_stage") ctx->opt_n_stage = (uint32_t)value;
    else if (k == "max_resident_bytes") ctx->max_resident_bytes = value;
    else if (k == "staging_bytes") ctx->staging_bytes = value;
    else return set_err(ACVMB_ERR_INVALID_ARG, "unknown option " + k);
    return ACVMB_OK;
}

// ---------------------------------------------------------------------------------------------
// The kernels find the lookup tables through a __constant__ pointer pair, which is per DEVICE (not per context), and the
// tables are the same data for every context: they are built and uploaded once per device for the life of the process and
// shared by all contexts of that device (a table owned by one context would dangle in the symbol once that context is gone).
#include <mutex>
static std::mutex g_tables_mutex;
static struct {
    uint32_t* fixed_base = nullptr;
    uint32_t* pedersen = nullptr;
    bool ready = false;
} g_tables[64];

static int ensure_curve_tables(acvmb_ctx* ctx) {
    if (ctx->tables_ready) return ACVMB_OK;
    const int d = ctx->device;
    if (d < 0 || d >= 64) return set_err(ACVMB_ERR_INVALID_ARG, "device index out of range");
    std::lock_guard<std::mutex> lk(g_tables_mutex);
    CUDA_TRY(cudaSetDevice(d));
    if (!g_tables[d].ready) {
        std::vector<uint32_t> fb = gk::build_fixed_base_table();
        CUDA_TRY(cudaMalloc(&g_tables[d].fixed_base, fb.size() * 4));
        CUDA_TRY(cudaMemcpy(g_tables[d].fixed_base, fb.data(), fb.size() * 4, cudaMemcpyHostToDevice));
        std::vector<uint32_t> pt = gk::build_pedersen_tables();
        CUDA_TRY(cudaMalloc(&g_tables[d].pedersen, pt.size() * 4));
        CUDA_TRY(cudaMemcpy(g_tables[d].pedersen, pt.data(), pt.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(set_curve_tables(g_tables[d].fixed_base, g_tables[d].pedersen));
        g_tables[d].ready = true;
    }
    ctx->tables_ready = true;
    return ACVMB_OK;
}

// device buffers of a plan: {field, bytes}
struct PlanBuf {
    void** dev;
    size_t bytes;
    const void* host;
};
static std::vector<PlanBuf> plan_buffers(acvmb_circuit* c, std::vector<uint32_t>& slots_scratch) {
    const Plan& p = c->plan;
    slots_scratch = p.input_witnesses;   // bit 31: the scatter stores value * R (scaled column)
    for (size_t i = 0; i < slots_scratch.size() && i < p.input_scaled.size(); ++i)
        if (p.input_scaled[i]) slots_scratch[i] |= 0x80000000u;
    std::vector<PlanBuf> b;
    b.push_back({(void**)&c->d_stream, c->stream_bytes, p.stream.data()});
    b.push_back({(void**)&c->d_payload, p.payload.size() * 4, p.payload.data()});
    b.push_back({(void**)&c->d_assign, p.assign_opcode.size() * 4, p.assign_opcode.data()});
    b.push_back({(void**)&c->d_input_slots, slots_scratch.size() * 4, slots_scratch.data()});
    b.push_back({(void**)&c->d_mu_index_of, p.mu_index_of.size() * 4, p.mu_index_of.data()});
    if (!p.unscale.empty()) b.push_back({(void**)&c->d_unscale, p.unscale.size() * 4, p.unscale.data()});
    return b;
}

static int upload_plan(acvmb_circuit* c) {
    const Plan& p = c->plan;
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    if (p.needs_full_kernel) {
        int rc = ensure_curve_tables(c->ctx);
        if (rc) return rc;
    }
    c->stream_bytes = p.stream.size() * sizeof(OpRec);
    std::vector<uint32_t> slots;
    for (PlanBuf& b : plan_buffers(c, slots)) {
        CUDA_TRY(cudaMalloc(b.dev, std::max<size_t>(b.bytes, 16)));
        if (b.bytes) CUDA_TRY(cudaMemcpy(*b.dev, b.host, b.bytes, cudaMemcpyHostToDevice));
    }
    return ACVMB_OK;
}

// ---------------------------------------------------------------------------------------------
// Multi-device context (SURVEY 8b `acvmb_create(devices, n)`, 8e): the batch shards contiguously over the devices, the
// compiled plan reaches them through ONE broadcast from devices[0] -- NCCL (ncclBroadcast over NVLink) when libnccl.so.2
// can be loaded, peer copies otherwise -- and nothing moves between devices during a solve.
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) {
            a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.lib, "ncclCommInitAll");
            a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
            a.GroupStart = (decltype(a.GroupStart))dlsym(a.lib, "ncclGroupStart");
            a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.lib, "ncclGroupEnd");
            a.Broadcast = (decltype(a.Broadcast))dlsym(a.lib, "ncclBroadcast");
            a.ok = a.CommInitAll && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Broadcast;
        }
        return a;
    }();
    return api;
}
}  // namespace

static void destroy_nccl(acvmb_ctx* ctx) {
    if (!ctx->nccl_comms) return;
    void** comms = (void**)ctx->nccl_comms;
    for (size_t i = 0; i < 1 + ctx->peers.size(); ++i)
        if (comms[i]) nccl_api().CommDestroy(comms[i]);
    delete[] comms;
    ctx->nccl_comms = nullptr;
}

extern "C" int acvmb_ctx_create_multi(const int* devices, int n, acvmb_ctx** out) {
    if (!devices || n < 1 || !out) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return set_err(ACVMB_ERR_INVALID_ARG, "a device is listed twice");
    acvmb_ctx* root = nullptr;
    int rc = acvmb_ctx_create(devices[0], &root);
    if (rc) return rc;
    for (int i = 1; i < n; ++i) {
        acvmb_ctx* p = nullptr;
        rc = acvmb_ctx_create(devices[i], &p);
        if (rc) {
            acvmb_ctx_destroy(root);
            return rc;
        }
        root->peers.push_back(p);
    }
    *out = root;
    return ACVMB_OK;
}

extern "C" int acvmb_ctx_n_devices(const acvmb_ctx* ctx) { return ctx ? 1 + (int)ctx->peers.size() : 0; }
extern "C" const char* acvmb_ctx_broadcast_backend(const acvmb_ctx* ctx) { return ctx ? ctx->bcast_backend : "none"; }

static void destroy_shards(acvmb_circuit* c) {
    for (acvmb_circuit* sh : c->shards) {
        cudaSetDevice(sh->ctx->device);
        if (cached_batch_circuit(sh->ctx) == sh) drop_cached_batch(sh->ctx);
        delete sh;
    }
    c->shards.clear();
}

// replicate the uploaded plan of `c` (on devices[0]) to every peer device: the one collective of the path
static int replicate_to_peers(acvmb_circuit* c) {
    acvmb_ctx* root = c->ctx;
    if (root->peers.empty()) return ACVMB_OK;
    const size_t n = 1 + root->peers.size();
    std::vector<acvmb_circuit*> all = {c};
    for (acvmb_ctx* p : root->peers) {
        auto sh = std::make_unique<acvmb_circuit>();
        sh->ctx = p;
        sh->plan = c->plan;
        sh->plan.stream.clear();           // host copy of the record stream stays with the root (serialisation)
        sh->plan.stream.shrink_to_fit();
        sh->stream_bytes = c->stream_bytes;
        sh->circuit = c->circuit;
        sh->has_circuit = c->has_circuit;
        c->shards.push_back(sh.release());
        all.push_back(c->shards.back());
    }
    std::vector<std::vector<uint32_t>> scratch(n);
    std::vector<std::vector<PlanBuf>> bufs(n);
    for (size_t i = 0; i < n; ++i) {
        bufs[i] = plan_buffers(all[i], scratch[i]);
        if (i == 0) continue;
        CUDA_TRY(cudaSetDevice(all[i]->ctx->device));
        if (c->plan.needs_full_kernel) {
            int rc = ensure_curve_tables(all[i]->ctx);
            if (rc) return rc;
        }
        for (PlanBuf& b : bufs[i]) CUDA_TRY(cudaMalloc(b.dev, std::max<size_t>(b.bytes, 16)));
    }
    NcclApi& api = nccl_api();
    bool use_nccl = api.ok;
    if (use_nccl && !root->nccl_comms) {
        std::vector<int> devs;
        for (acvmb_circuit* x : all) devs.push_back(x->ctx->device);
        void** comms = new void*[n]();
        if (api.CommInitAll(comms, (int)n, devs.data()) != 0) {
            delete[] comms;
            use_nccl = false;
        } else {
            root->nccl_comms = comms;
        }
    }
    if (use_nccl) {
        void** comms = (void**)root->nccl_comms;
        for (size_t k = 0; k < bufs[0].size(); ++k) {
            if (!bufs[0][k].bytes) continue;
            int e = api.GroupStart();
            for (size_t i = 0; i < n && e == 0; ++i) {
                cudaSetDevice(all[i]->ctx->device);
                e = api.Broadcast(*bufs[0][k].dev, *bufs[i][k].dev, bufs[0][k].bytes, /*ncclUint8*/ 1, 0, comms[i], all[i]->ctx->stream);
            }
            if (e == 0) e = api.GroupEnd(); else api.GroupEnd();
            if (e != 0) return set_err(ACVMB_ERR_CUDA, "ncclBroadcast of the plan failed (nccl error " + std::to_string(e) + ")");
        }
        root->bcast_backend = "nccl";
    } else {
        CUDA_TRY(cudaSetDevice(root->device));
        for (size_t i = 1; i < n; ++i)
            for (size_t k = 0; k < bufs[0].size(); ++k)
                if (bufs[0][k].bytes)
                    CUDA_TRY(cudaMemcpyPeerAsync(*bufs[i][k].dev, all[i]->ctx->device, *bufs[0][k].dev, root->device, bufs[0][k].bytes, root->stream));
        root->bcast_backend = "peer-copy";
    }
    for (size_t i = 0; i < n; ++i) {
        CUDA_TRY(cudaSetDevice(all[i]->ctx->device));
        CUDA_TRY(cudaStreamSynchronize(all[i]->ctx->stream));
    }
    CUDA_TRY(cudaSetDevice(root->device));
    return ACVMB_OK;
}

static int circuit_from_struct(acvmb_ctx* ctx, const Circuit& circ, const uint32_t* input_witnesses, uint32_t n_inputs,
                               acvmb_circuit** out) {
    auto c = std::make_unique<acvmb_circuit>();
    c->ctx = ctx;
    PlanOptions opt;
    opt.S = ctx->opt_S;
    if (!opt.S) {
        // Curve calls are the slow micro-ops by three orders of magnitude.  They want every lane of a warp busy (T = 32
        // instances per slot, pick_T) and their partial sums spread over slot threads (split lowering needs S >= 8);
        // measured on B200 (profiles/r1_curve_tile_shapes.txt) T=32,S=8 is ~2x the arithmetic-tuned T=8,S=16 on both the
        // Pedersen chain and the mixed circuit.
        opt.S = 16;
        size_t n_hash = 0;
        for (auto& op : circ.opcodes) {
            if (op.kind != OP_BlackBox) continue;
            if (op.bb().func == BB_Pedersen || op.bb().func == BB_FixedBaseScalarMul || op.bb().func == BB_EcdsaSecp256k1 ||
                op.bb().func == BB_EcdsaSecp256r1) {
                opt.S = 8;
                break;
            }
            n_hash += op.bb().func == BB_SHA256 || op.bb().func == BB_Keccak256 || op.bb().func == BB_Blake2s ||
                      op.bb().func == BB_Keccak256VariableLength || op.bb().func == BB_HashToField128Security;
        }
        // Hash-dominated circuits (a hash call is ~5 k instructions of one thread, a gate ~12): wider tiles, fewer CTAs, so that
        // the warps that run the hash cores do not share a scheduler -- T = 16 / S = 8 is 7 % faster than T = 8 / S = 16 on
        // the config-3 chain (profiles/r2_hash_packed_ab.txt)
        if (opt.S == 16 && n_hash * 8 >= circ.opcodes.size() && n_hash) opt.S = 8;
    }
    opt.chunk_steps = ctx->opt_chunk_steps;
    opt.split_curve = ctx->opt_split_curve != 0;
    if (ctx->opt_temp_pool) opt.temp_pool = ctx->opt_temp_pool;
    opt.allow_unpinned_pedersen = ctx->opt_pedersen_unpinned != 0;
    opt.device_brillig = ctx->opt_device_brillig != 0;
    opt.scaled_columns = ctx->opt_scaled_columns != 0;
    opt.packed_hashes = ctx->opt_packed_hashes != 0;
    opt.sha_pad_table = ctx->opt_sha_pad_table != 0;
    opt.spread_heavy = ctx->opt_spread_heavy != 0;
    opt.slack_scheduling = ctx->opt_slack_scheduling != 0;
    opt.tile_lanes = ctx->opt_T;
    {
        // the ring is [entries][T lanes][32 B]: size it for the tile width pick_T() will choose for this circuit
        bool curve = false;
        for (auto& op : circ.opcodes)
            curve |= op.kind == OP_BlackBox && (op.bb().func == BB_Pedersen || op.bb().func == BB_FixedBaseScalarMul ||
                                                op.bb().func == BB_EcdsaSecp256k1 || op.bb().func == BB_EcdsaSecp256r1);
        uint32_t T_guess = ctx->opt_T ? ctx->opt_T : (curve ? 32u : std::max(1u, 128u / opt.S));
        if (T_guess > 32) T_guess = 32;
        opt.ring_slots = std::min<uint32_t>(1024u, ctx->opt_ring_bytes / (T_guess * 32u));
    }
    std::vector<uint32_t> inputs(input_witnesses, input_witnesses + n_inputs);
    try {
        c->plan = compile_plan(circ, inputs, opt);
        if (c->plan.needs_full_kernel && c->plan.ring_slots) {   // the ring variant exists for the arithmetic / logic kernel only
            opt.ring_slots = 0;
            c->plan = compile_plan(circ, inputs, opt);
        }
    } catch (const std::exception& e) {
        return set_err(ACVMB_ERR_UNSUPPORTED, e.what());
    }
    int rc = upload_plan(c.get());
    if (rc) return rc;
    *out = c.release();
    return ACVMB_OK;
}

extern "C" int acvmb_circuit_from_acir(acvmb_ctx* ctx, const uint8_t* gz, size_t len, const uint32_t* input_witnesses,
                                       uint32_t n_inputs, acvmb_circuit** out) {
    if (!ctx || !gz || !out || (n_inputs && !input_witnesses)) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    Circuit circ;
    try {
        circ = decode_circuit(gz, len);
    } catch (const std::exception& e) {
        return set_err(ACVMB_ERR_DECODE, e.what());
    }
    int rc = circuit_from_struct(ctx, circ, input_witnesses, n_inputs, out);
    if (rc) return rc;
    bool host = false;
    for (auto& sg : (*out)->plan.segments) host |= sg.kind == 1;
    if (host) {
        (*out)->plan.acir_gz.assign(gz, gz + len);
        (*out)->circuit = std::move(circ);
        (*out)->has_circuit = true;
    }
    rc = replicate_to_peers(*out);
    if (rc) {
        acvmb_circuit_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

static acvmb_circuit* cached_batch_circuit(acvmb_ctx* ctx);

extern "C" void acvmb_circuit_destroy(acvmb_circuit* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    destroy_shards(c);
    cudaSetDevice(c->ctx->device);
    if (cached_batch_circuit(c->ctx) == c) drop_cached_batch(c->ctx);
    delete c;
}

extern "C" int acvmb_circuit_info(const acvmb_circuit* c, acvmb_plan_info* o) {
    if (!c || !o) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    const Plan& p = c->plan;
    memset(o, 0, sizeof(*o));
    o->n_opcodes = p.stats.n_opcodes;
    o->n_micro_ops = p.stats.n_micro;
    o->n_steps = p.stats.n_steps;
    o->n_slots_filled = p.stats.n_slots_filled;
    o->n_gate_assign = p.stats.n_gate_assign;
    o->n_gate_check = p.stats.n_gate_check;
    o->n_logic = p.stats.n_logic;
    o->n_range = p.stats.n_range;
    o->n_hash = p.stats.n_hash;
    o->n_curve = p.stats.n_curve;
    o->ref_fr_mul = p.stats.ref_fr_mul;
    o->ref_fr_inv = p.stats.ref_fr_inv;
    o->dev_imad = p.stats.dev_imad;
    o->alg_bytes = p.stats.alg_bytes;
    o->n_temps = p.stats.n_temps;
    o->num_witnesses = p.num_witnesses;
    o->n_slots = p.n_slots;
    o->S = p.S;
    o->needs_full_kernel = p.needs_full_kernel;
    o->static_fail_present = p.static_fail.present;
    o->static_fail_opcode = p.static_fail.opcode;
    o->static_fail_kind = p.static_fail.kind;
    o->static_fail_aux = p.static_fail.aux;
    o->n_segments = (uint32_t)p.segments.size();
    for (auto& sg : p.segments) o->n_host_segments += sg.kind != 0;
    o->n_brillig = (uint32_t)p.stats.n_brillig;
    o->n_brillig_device = (uint32_t)p.stats.n_brillig_device;
    o->n_gate_one_reduction = p.stats.n_gate_one_reduction;
    o->scaled_columns = p.unscale.empty() ? 0u : 1u;
    o->ring_slots = p.ring_slots;
    o->n_operand_reads = p.stats.n_operand_reads;
    o->n_ring_reads = p.stats.n_ring_reads;
    return ACVMB_OK;
}

extern "C" int acvmb_circuit_assign_opcodes(const acvmb_circuit* c, uint32_t* out, uint32_t n) {
    if (!c || !out) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    if (n != c->plan.num_witnesses) return set_err(ACVMB_ERR_INVALID_ARG, "n must equal num_witnesses");
    memcpy(out, c->plan.assign_opcode.data(), (size_t)n * 4);
    return ACVMB_OK;
}

extern "C" int acvmb_circuit_serialize(const acvmb_circuit* c, uint8_t* buf, size_t cap, size_t* needed) {
    if (!c || !needed) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    std::vector<uint8_t> blob = serialize_plan(c->plan);
    *needed = blob.size();
    if (!buf) return ACVMB_OK;
    if (cap < blob.size()) return set_err(ACVMB_ERR_INVALID_ARG, "buffer too small");
    memcpy(buf, blob.data(), blob.size());
    return ACVMB_OK;
}

extern "C" int acvmb_circuit_deserialize(acvmb_ctx* ctx, const uint8_t* blob, size_t len, acvmb_circuit** out) {
    if (!ctx || !blob || !out) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    auto c = std::make_unique<acvmb_circuit>();
    c->ctx = ctx;
    try {
        c->plan = deserialize_plan(blob, len);
        if (!c->plan.acir_gz.empty()) {
            c->circuit = decode_circuit(c->plan.acir_gz.data(), c->plan.acir_gz.size());
            c->has_circuit = true;
        }
    } catch (const std::exception& e) {
        return set_err(ACVMB_ERR_DECODE, e.what());
    }
    int rc = upload_plan(c.get());
    if (rc) return rc;
    rc = replicate_to_peers(c.get());
    if (rc) return rc;
    *out = c.release();
    return ACVMB_OK;
}

// ---------------------------------------------------------------------------------------------
static uint32_t pick_T(const acvmb_ctx* ctx, const Plan& p) {
    uint32_t T = ctx->opt_T ? ctx->opt_T : std::max(1u, 128u / p.S);
    if (!ctx->opt_T && p.stats.n_curve && vm_config_supported(32, (int)p.S, p.needs_full_kernel)) T = 32;   // full warps per curve micro-op
    if (T > 32) T = 32;
    while (T > 1 && !vm_config_supported((int)T, (int)p.S, p.needs_full_kernel)) T /= 2;
    return T;
}

extern "C" int acvmb_batch_create(acvmb_circuit* c, uint32_t n_instances, acvmb_batch** out) {
    if (!c || !out || n_instances == 0) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    auto b = std::make_unique<acvmb_batch>();
    b->c = c;
    b->n_inst = n_instances;
    b->T = pick_T(c->ctx, c->plan);
    if (!vm_config_supported((int)b->T, (int)c->plan.S, c->plan.needs_full_kernel))
        return set_err(ACVMB_ERR_INVALID_ARG, "no kernel instantiation for T=" + std::to_string(b->T) + " S=" + std::to_string(c->plan.S));
    b->capacity = n_instances;
    b->n_tiles = (n_instances + b->T - 1) / b->T;
    size_t col_bytes = (size_t)b->n_tiles * b->T * c->plan.n_slots * 32;
    CUDA_TRY(cudaMalloc(&b->d_cols, col_bytes));
    CUDA_TRY(cudaMalloc(&b->d_fail, (size_t)b->n_tiles * b->T * 8));
    if (c->plan.n_mu) CUDA_TRY(cudaMalloc(&b->d_mu, (size_t)b->n_tiles * b->T * c->plan.n_mu * 4));
    size_t in_bytes = (size_t)n_instances * c->plan.input_witnesses.size() * 32;
    CUDA_TRY(cudaMalloc(&b->d_in, std::max<size_t>(in_bytes, 16)));
    CUDA_TRY(cudaEventCreate(&b->ev0));
    CUDA_TRY(cudaEventCreate(&b->ev1));
    CUDA_TRY(cudaEventCreateWithFlags(&b->ev_ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&b->ev_dl_done, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&b->ev_gather[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&b->ev_copy[i], cudaEventDisableTiming));
    }
    *out = b.release();
    return ACVMB_OK;
}

// run fewer instances than the buffers were sized for (sub-batches of unequal size share one buffer)
extern "C" int acvmb_batch_resize(acvmb_batch* b, uint32_t n_instances) {
    if (!b || n_instances == 0 || n_instances > b->capacity) return set_err(ACVMB_ERR_INVALID_ARG, "resize beyond capacity");
    b->n_inst = n_instances;
    b->n_tiles = (n_instances + b->T - 1) / b->T;
    return ACVMB_OK;
}

extern "C" void acvmb_batch_destroy(acvmb_batch* b) {
    if (!b) return;
    cudaSetDevice(b->c->ctx->device);
    delete b;
}

static void drop_cached_batch(acvmb_ctx* ctx) {
    if (ctx->cached_batch) acvmb_batch_destroy(ctx->cached_batch);
    ctx->cached_batch = nullptr;
    for (int i = 0; i < 2; ++i) {
        if (ctx->pipe_batch[i]) acvmb_batch_destroy(ctx->pipe_batch[i]);
        ctx->pipe_batch[i] = nullptr;
    }
    ctx->cached_bytes = 0;
}
static acvmb_circuit* cached_batch_circuit(acvmb_ctx* ctx) {
    if (ctx->cached_batch) return ctx->cached_batch->c;
    return ctx->pipe_batch[0] ? ctx->pipe_batch[0]->c : nullptr;
}

extern "C" int acvmb_batch_upload(acvmb_batch* b, const uint8_t* inputs_be32) {
    if (!b) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    cudaGetLastError();   // a stale error of an unrelated earlier call must not be attributed to this one
    acvmb_circuit* c = b->c;
    cudaStream_t s = c->ctx->stream;
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    uint32_t n_in = (uint32_t)c->plan.input_witnesses.size();
    size_t in_bytes = (size_t)b->n_inst * n_in * 32;
    if (in_bytes && !inputs_be32) return set_err(ACVMB_ERR_INVALID_ARG, "inputs are NULL");
    CUDA_TRY(cudaEventRecord(b->ev0, s));
    if (in_bytes) CUDA_TRY(cudaMemcpyAsync(b->d_in, inputs_be32, in_bytes, cudaMemcpyHostToDevice, s));
    CUDA_TRY(launch_fill_u64(b->d_fail, (size_t)b->n_tiles * b->T, ~0ull, s));
    if (b->d_mu) CUDA_TRY(cudaMemsetAsync(b->d_mu, 0xFF, (size_t)b->n_tiles * b->T * c->plan.n_mu * 4, s));
    CUDA_TRY(launch_scatter_inputs(b->d_in, c->d_input_slots, n_in, b->d_cols, c->plan.n_slots, (int)b->T, b->n_inst, s));
    CUDA_TRY(cudaEventRecord(b->ev1, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, b->ev0, b->ev1);
    c->run.scatter_ms += ms;
    c->run.kernel_launches += 2;
    return ACVMB_OK;
}

static int run_host_brillig(acvmb_batch* b, const Segment& sg);
static int run_host_permutation_sort(acvmb_batch* b, const Segment& sg);

extern "C" int acvmb_batch_run(acvmb_batch* b, float* kernel_ms) {
    if (!b) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    cudaGetLastError();   // a stale error of an unrelated earlier call must not be attributed to this one
    acvmb_circuit* c = b->c;
    cudaStream_t s = c->ctx->stream;
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    VmArgs a{};
    a.stream = c->d_stream;
    a.payload = c->d_payload;
    a.cols = b->d_cols;
    a.fail = b->d_fail;
    a.chunk_steps = c->plan.chunk_steps;
    a.n_stage = c->ctx->opt_n_stage;
    a.n_slots = c->plan.n_slots;
    a.n_tiles = b->n_tiles;
    a.mu_assign = b->d_mu;
    a.n_mu = c->plan.n_mu;
    a.ring_slots = c->plan.ring_slots;
    KernelConfig cfg{(int)b->T, (int)c->plan.S, c->plan.needs_full_kernel, c->ctx->opt_split};
    float total = 0;
    for (const Segment& sg : c->plan.segments) {
        if (sg.kind == 0) {
            a.first_step = sg.a;
            a.n_steps = sg.b;
            CUDA_TRY(cudaEventRecord(b->ev0, s));
            CUDA_TRY(launch_vm(cfg, a, s));
            CUDA_TRY(cudaEventRecord(b->ev1, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
            total += ms;
            c->run.kernel_launches += 1;
        } else {
            int rc = sg.kind == 2 ? run_host_permutation_sort(b, sg) : run_host_brillig(b, sg);
            if (rc) return rc;
        }
    }
    if (kernel_ms) *kernel_ms = total;
    c->run.kernel_ms += total;
    c->run.T = b->T;
    c->run.S = c->plan.S;
    c->run.n_tiles = b->n_tiles;
    c->run.threads_per_cta = b->T * c->plan.S;
    return ACVMB_OK;
}

// Resident-input variant used for kernel-only throughput: inputs for several sub-batches are copied
// to HBM once (stage_inputs), then each run_staged = reset statuses + scatter + step-VM kernel, all on
// the library stream and timed with CUDA events there.
extern "C" int acvmb_batch_stage_inputs(acvmb_batch* b, uint32_t slot, const uint8_t* inputs_be32) {
    if (!b || !inputs_be32 || slot > 4096) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(b->c->ctx->device));
    size_t in_bytes = (size_t)b->n_inst * b->c->plan.input_witnesses.size() * 32;
    size_t cap_bytes = (size_t)b->capacity * b->c->plan.input_witnesses.size() * 32;
    if (b->d_staged_in.size() <= slot) b->d_staged_in.resize(slot + 1, nullptr);
    if (!b->d_staged_in[slot]) CUDA_TRY(cudaMalloc(&b->d_staged_in[slot], std::max<size_t>(cap_bytes, 16)));
    CUDA_TRY(cudaMemcpy(b->d_staged_in[slot], inputs_be32, in_bytes, cudaMemcpyHostToDevice));
    return ACVMB_OK;
}

extern "C" int acvmb_batch_run_staged(acvmb_batch* b, uint32_t slot, float* total_ms, float* vm_ms) {
    if (!b || slot >= b->d_staged_in.size() || !b->d_staged_in[slot]) return set_err(ACVMB_ERR_INVALID_ARG, "no staged inputs in this slot");
    acvmb_circuit* c = b->c;
    cudaStream_t s = c->ctx->stream;
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, s));
    CUDA_TRY(launch_fill_u64(b->d_fail, (size_t)b->n_tiles * b->T, ~0ull, s));
    if (b->d_mu) CUDA_TRY(cudaMemsetAsync(b->d_mu, 0xFF, (size_t)b->n_tiles * b->T * c->plan.n_mu * 4, s));
    CUDA_TRY(launch_scatter_inputs(b->d_staged_in[slot], c->d_input_slots, (uint32_t)c->plan.input_witnesses.size(), b->d_cols,
                                   c->plan.n_slots, (int)b->T, b->n_inst, s));
    c->run.kernel_launches += 2;
    float vm = 0;
    int rc = acvmb_batch_run(b, &vm);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(e1, s));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (total_ms) *total_ms = ms;
    if (vm_ms) *vm_ms = vm;
    return ACVMB_OK;
}

// ---------------------------------------------------------------------------------------------
// Host segment: one Brillig opcode for every live instance of the batch (acvm/src/pwg/brillig.rs:20-131).
// D2H of exactly the columns the opcode reads, VM runs on all host threads, H2D + scatter of the outputs, failures
// merged into the device status words.
// ---------------------------------------------------------------------------------------------
static inline unsigned long long fail_key(uint32_t opcode, uint32_t kind, uint32_t aux) {
    return ((unsigned long long)opcode << 32) | ((unsigned long long)(kind & 0xF) << 28) | (aux & 0x0FFFFFFFu);
}

// Shared I/O of a host segment: gather `in_slots` (+ already-assigned outputs), call per_instance(i, val, result) on all
// host threads for the instances that are still live at `opcode`, then insert_value the results and merge failures.
// per_instance returns ~0ull when the outputs are to be inserted, else the failure key to record.
template <class F>
static int run_host_io(acvmb_batch* b, uint32_t opcode, std::vector<uint32_t> in_slots, const std::vector<uint32_t>& out_w,
                       const std::vector<uint32_t>& out_known, F&& per_instance) {
    acvmb_circuit* c = b->c;
    acvmb_ctx* ctx = c->ctx;
    std::vector<int> known_pos(out_w.size(), -1);   // already-assigned outputs are read too (insert_value compares)
    for (size_t k = 0; k < out_w.size(); ++k)
        if (out_known[k]) {
            known_pos[k] = (int)in_slots.size();
            in_slots.push_back(out_w[k]);
        }
    const uint32_t n = b->n_inst, n_g = (uint32_t)in_slots.size(), n_o = (uint32_t)out_w.size();
    cudaStream_t s = ctx->stream;
    // ---- D2H: status words + the input columns ----
    std::vector<unsigned long long>& fail = b->h_fail;
    fail.resize(n);
    CUDA_TRY(cudaMemcpyAsync(fail.data(), b->d_fail, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    std::vector<uint8_t>& in_be = b->h_in;
    in_be.resize((size_t)n * n_g * 32);
    size_t io_bytes = std::max<size_t>((size_t)n * std::max(n_g, n_o) * 32, 16);
    size_t ids_bytes = std::max<size_t>((size_t)std::max(n_g, n_o) * 4, 16);
    if (b->host_io_bytes < io_bytes) {
        if (b->d_host_io) cudaFree(b->d_host_io);
        b->d_host_io = nullptr;
        b->host_io_bytes = 0;
        CUDA_TRY(cudaMalloc(&b->d_host_io, io_bytes));
        b->host_io_bytes = io_bytes;
    }
    if (b->host_ids_bytes < ids_bytes) {
        if (b->d_host_ids) cudaFree(b->d_host_ids);
        b->d_host_ids = nullptr;
        b->host_ids_bytes = 0;
        CUDA_TRY(cudaMalloc(&b->d_host_ids, ids_bytes));
        b->host_ids_bytes = ids_bytes;
    }
    uint8_t* d_io = b->d_host_io;
    uint32_t* d_ids = b->d_host_ids;
    if (n_g) {
        CUDA_TRY(cudaMemcpyAsync(d_ids, in_slots.data(), (size_t)n_g * 4, cudaMemcpyHostToDevice, s));
        GatherArgs g{};
        g.cols = b->d_cols;
        g.n_slots = c->plan.n_slots;
        g.T = (int)b->T;
        g.witness_ids = d_ids;
        g.n_out = n_g;
        g.first_inst = 0;
        g.n_inst = n;
        g.fail = b->d_fail;
        g.out_be = d_io;
        g.raw = 1;
        CUDA_TRY(launch_gather_outputs(g, s));
        CUDA_TRY(cudaMemcpyAsync(in_be.data(), d_io, (size_t)n * n_g * 32, cudaMemcpyDeviceToHost, s));
        c->run.kernel_launches += 1;
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    // ---- per instance on all host threads ----
    std::vector<uint8_t>& out_be = b->h_out;
    out_be.assign((size_t)n * n_o * 32, 0);
    unsigned n_thr = std::max(1u, std::thread::hardware_concurrency());
    n_thr = std::min<unsigned>(n_thr, n);
    auto work = [&](unsigned t) {
        for (uint32_t i = t; i < n; i += n_thr) {
            if ((uint32_t)(fail[i] >> 32) < opcode) continue;   // this instance stopped at an earlier opcode
            auto val = [&](uint32_t pos) { return hf::from_be_bytes_reduce(&in_be[((size_t)i * n_g + pos) * 32], 32); };
            std::vector<U256> result(n_o);
            unsigned long long key = per_instance(i, val, result);
            if (key != ~0ull) {
                fail[i] = std::min(fail[i], key);
                continue;
            }
            // insert_value in output order: an already-assigned witness is replaced, a mismatch is UnsatisfiedConstrain
            std::vector<std::pair<uint32_t, U256>> seen;
            for (uint32_t k = 0; k < n_o; ++k) {
                if (out_known[k]) {
                    U256 old;
                    bool found = false;
                    for (auto& kv : seen)
                        if (kv.first == out_w[k]) { old = kv.second; found = true; }
                    if (!found) old = val((uint32_t)known_pos[k]);
                    if (old != result[k]) fail[i] = std::min(fail[i], fail_key(opcode, EK_UNSATISFIED_CONSTRAIN, 0));
                }
                seen.emplace_back(out_w[k], result[k]);
                hf::to_be_bytes(result[k], &out_be[((size_t)i * n_o + k) * 32]);
            }
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < n_thr; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    // ---- H2D: outputs scattered into their columns, merged status words ----
    if (n_o) {
        CUDA_TRY(cudaMemcpyAsync(d_ids, out_w.data(), (size_t)n_o * 4, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(d_io, out_be.data(), (size_t)n * n_o * 32, cudaMemcpyHostToDevice, s));
        CUDA_TRY(launch_scatter_inputs(d_io, d_ids, n_o, b->d_cols, c->plan.n_slots, (int)b->T, n, s));
        c->run.kernel_launches += 1;
    }
    CUDA_TRY(cudaMemcpyAsync(b->d_fail, fail.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return ACVMB_OK;
}

static int run_host_brillig(acvmb_batch* b, const Segment& sg) {
    acvmb_circuit* c = b->c;
    if (!c->has_circuit) return set_err(ACVMB_ERR_STATE, "plan has a host segment but the circuit bytes are not attached");
    const Brillig& br = c->circuit.opcodes[sg.a].brillig();
    const uint32_t* d = c->plan.host_desc.data() + sg.b;
    const uint32_t opcode = sg.a;
    // ---- unpack the descriptor ----
    std::vector<uint32_t> in_slots;          // gathered columns, in order
    uint32_t pred_slot = *d++;
    if (pred_slot != 0xFFFFFFFFu) in_slots.push_back(pred_slot);
    struct In { bool arr; uint32_t first, count; };
    std::vector<In> ins;
    uint32_t n_in = *d++;
    for (uint32_t i = 0; i < n_in; ++i) {
        bool arr = *d++ != 0;
        uint32_t cnt = *d++;
        ins.push_back({arr, (uint32_t)in_slots.size(), cnt});
        for (uint32_t k = 0; k < cnt; ++k) in_slots.push_back(*d++);
    }
    struct Out { bool arr; uint32_t first, count; };
    std::vector<Out> outs;
    std::vector<uint32_t> out_w, out_known;
    uint32_t n_out = *d++;
    for (uint32_t i = 0; i < n_out; ++i) {
        bool arr = *d++ != 0;
        uint32_t cnt = *d++;
        outs.push_back({arr, (uint32_t)out_w.size(), cnt});
        for (uint32_t k = 0; k < cnt; ++k) {
            out_w.push_back(*d++);
            out_known.push_back(*d++);
        }
    }
    auto per_instance = [&](uint32_t i, auto& val, std::vector<U256>& result) -> unsigned long long {
        if (pred_slot != 0xFFFFFFFFu && val(0).is_zero()) return ~0ull;   // zero predicate: outputs are zeroed (brillig.rs:34-37,133-150)
        bvm::VM vm;
        for (auto& in : ins) {
            if (!in.arr) {
                vm.regs.push_back(val(in.first));
            } else {
                vm.regs.push_back(hf::from_u64(vm.mem.size()));
                for (uint32_t k = 0; k < in.count; ++k) vm.mem.push_back(val(in.first + k));
            }
        }
        bvm::Result r = vm.run(br);
        if (r.status == bvm::Status::Failure)
            return fail_key(opcode, EK_BRILLIG_FAILED, r.call_stack.empty() ? 0 : (uint32_t)r.call_stack.back());
        if (r.status == bvm::Status::Panic) return fail_key(opcode, EK_REFERENCE_PANIC, 0);
        if (r.status == bvm::Status::ForeignCallWait) {
            if (i == 0) {
                b->fc_pending = true;
                b->fc_function = r.message;
                b->fc_inputs = std::move(r.fc_inputs);
            }
            return fail_key(opcode, 0xF, 0);   // decoded as ACVMB_REQUIRES_FOREIGN_CALL
        }
        for (size_t oi = 0; oi < outs.size(); ++oi) {
            U256 reg = oi < vm.regs.size() ? vm.regs[oi] : U256{};
            if (!outs[oi].arr) {
                result[outs[oi].first] = reg;
            } else {
                if (reg.l[1] | reg.l[2] | reg.l[3]) return fail_key(opcode, EK_REFERENCE_PANIC, 0);
                for (uint32_t k = 0; k < outs[oi].count; ++k) {
                    size_t p = (size_t)reg.l[0] + k;
                    if (p >= vm.mem.size()) return fail_key(opcode, EK_REFERENCE_PANIC, 0);   // Vec index out of bounds panics
                    result[outs[oi].first + k] = vm.mem[p];
                }
            }
        }
        return ~0ull;
    };
    return run_host_io(b, opcode, std::move(in_slots), out_w, out_known, per_instance);
}

// Host segment kind 2: Directive::PermutationSort (directives/mod.rs:88-121).  Self-contained descriptor:
// n, tuple, n_sort_by, sort_by*, n*tuple value slots, n_bits, {witness, known}*.
static int run_host_permutation_sort(acvmb_batch* b, const Segment& sg) {
    const uint32_t* d = b->c->plan.host_desc.data() + sg.b;
    const uint32_t opcode = sg.a;
    const uint32_t n = *d++, tuple = *d++, n_sort = *d++;
    std::vector<uint32_t> sort_by(d, d + n_sort);
    d += n_sort;
    std::vector<uint32_t> in_slots(d, d + (size_t)n * tuple);
    d += (size_t)n * tuple;
    const uint32_t n_bits = *d++;
    std::vector<uint32_t> out_w, out_known;
    for (uint32_t k = 0; k < n_bits; ++k) {
        out_w.push_back(*d++);
        out_known.push_back(*d++);
    }
    auto per_instance = [&](uint32_t, auto& val, std::vector<U256>& result) -> unsigned long long {
        std::vector<U256> values((size_t)n * tuple);
        for (size_t k = 0; k < values.size(); ++k) values[k] = val((uint32_t)k);
        std::vector<uint8_t> bits;
        if (!psort::permutation_sort_bits(values, n, tuple, sort_by, bits)) return fail_key(opcode, EK_REFERENCE_PANIC, 0);
        for (uint32_t k = 0; k < n_bits; ++k) result[k] = hf::from_u64(k < bits.size() ? bits[k] : 0);
        return ~0ull;
    };
    return run_host_io(b, opcode, std::move(in_slots), out_w, out_known, per_instance);
}

static void decode_status(const Plan& p, unsigned long long word, acvmb_status* st) {
    uint32_t fop = (uint32_t)(word >> 32);
    if (p.static_fail.present && p.static_fail.opcode <= fop) {
        // a per-instance failure at the same opcode cannot exist: the static failure stops the plan there
        st->code = ACVMB_FAILURE;
        st->err_kind = p.static_fail.kind;
        st->opcode_index = p.static_fail.opcode;
        st->aux = p.static_fail.aux;
    } else if (word == ~0ull) {
        st->code = ACVMB_SOLVED;
        st->err_kind = ACVMB_E_NONE;
        st->opcode_index = p.n_opcodes;
        st->aux = 0;
    } else if (((word >> 28) & 0xF) == 0xF) {
        // a Brillig foreign call has no recorded result: ACVMStatus::RequiresForeignCall, ip NOT advanced (mod.rs:267)
        st->code = ACVMB_REQUIRES_FOREIGN_CALL;
        st->err_kind = ACVMB_E_NONE;
        st->opcode_index = fop;
        st->aux = 0;
    } else {
        st->code = ACVMB_FAILURE;
        st->err_kind = (uint32_t)((word >> 28) & 0xF);
        st->opcode_index = fop;
        st->aux = (uint32_t)(word & 0x0FFFFFFFu);
    }
}

extern "C" int acvmb_batch_status(acvmb_batch* b, acvmb_status* out_status) {
    if (!b || !out_status) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(b->c->ctx->device));
    std::vector<unsigned long long> words(b->n_inst);
    CUDA_TRY(cudaMemcpy(words.data(), b->d_fail, (size_t)b->n_inst * 8, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < b->n_inst; ++i) decode_status(b->c->plan, words[i], &out_status[i]);
    return ACVMB_OK;
}

// Output path: gather kernels (columns -> [inst][witness][32 B BE] staging, double-buffered) on the gather stream, D2H on the
// copy stream, both ordered by events only.  async: return once everything is enqueued (ev_dl_done marks the last copy), so
// the caller can run the NEXT sub-batch's VM kernel on the VM stream while this one drains over PCIe.
static int download_impl(acvmb_batch* b, uint32_t first, uint32_t n, const uint32_t* out_ids, uint32_t n_out_ids, uint8_t* out,
                         uint8_t* out_present, bool async);

static int wait_download(acvmb_batch* b) {
    if (b && b->dl_pending) {
        b->dl_pending = false;
        CUDA_TRY(cudaEventSynchronize(b->ev_dl_done));
    }
    return ACVMB_OK;
}

extern "C" int acvmb_batch_download_ex(acvmb_batch* b, uint32_t first, uint32_t n, const uint32_t* out_ids, uint32_t n_out_ids,
                                       uint8_t* out, uint8_t* out_present) {
    return download_impl(b, first, n, out_ids, n_out_ids, out, out_present, false);
}

static int download_impl(acvmb_batch* b, uint32_t first, uint32_t n, const uint32_t* out_ids, uint32_t n_out_ids, uint8_t* out,
                         uint8_t* out_present, bool async) {
    if (!b || (!out && !out_present) || first + n > b->n_inst) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    acvmb_circuit* c = b->c;
    acvmb_ctx* ctx = c->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint32_t n_out = out_ids ? n_out_ids : c->plan.num_witnesses;
    if (n == 0 || n_out == 0) return ACVMB_OK;
    if (out_ids) {
        for (uint32_t i = 0; i < n_out_ids; ++i)
            if (out_ids[i] >= c->plan.num_witnesses) return set_err(ACVMB_ERR_INVALID_ARG, "witness index out of range");
        if (b->out_ids_cap < n_out_ids) {
            cudaFree(b->d_out_ids);
            b->d_out_ids = nullptr;
            CUDA_TRY(cudaMalloc(&b->d_out_ids, (size_t)n_out_ids * 4));
            b->out_ids_cap = n_out_ids;
        }
        CUDA_TRY(cudaMemcpyAsync(b->d_out_ids, out_ids, (size_t)n_out_ids * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    size_t row = (size_t)n_out * 32;
    uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(n, ctx->staging_bytes / row));
    size_t need = (size_t)chunk * row;
    if (out && b->stage_bytes < need) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(b->d_stage[i]);
            b->d_stage[i] = nullptr;
        }
        b->stage_bytes = 0;
        CUDA_TRY(cudaMalloc(&b->d_stage[0], need));
        CUDA_TRY(cudaMalloc(&b->d_stage[1], need));
        b->stage_bytes = need;
    }
    size_t need_p = (size_t)chunk * n_out;
    if (out_present && b->stage_present_bytes < need_p) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(b->d_stage_present[i]);
            b->d_stage_present[i] = nullptr;
        }
        b->stage_present_bytes = 0;
        CUDA_TRY(cudaMalloc(&b->d_stage_present[0], need_p));
        CUDA_TRY(cudaMalloc(&b->d_stage_present[1], need_p));
        b->stage_present_bytes = need_p;
    }
    GatherArgs g{};
    g.cols = b->d_cols;
    g.n_slots = c->plan.n_slots;
    g.T = (int)b->T;
    g.witness_ids = out_ids ? b->d_out_ids : nullptr;
    g.n_out = n_out;
    g.fail = b->d_fail;
    g.assign_opcode = c->d_assign;
    g.mu_index_of = c->d_mu_index_of;
    g.mu_assign = b->d_mu;
    g.n_mu = c->plan.n_mu;
    g.static_fail_opcode = c->plan.static_fail.present ? c->plan.static_fail.opcode : 0xFFFFFFFFu;
    g.unscale = c->d_unscale;
    // everything queued on the VM stream so far (the solve, the id upload above) precedes the first gather
    CUDA_TRY(cudaEventRecord(b->ev_ready, ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->gather_stream, b->ev_ready, 0));
    CUDA_TRY(cudaEventRecord(b->ev0, ctx->gather_stream));
    uint32_t k = 0;
    for (uint32_t off = 0; off < n; off += chunk, ++k) {
        uint32_t cnt = std::min(chunk, n - off);
        int buf = k & 1;
        if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(ctx->gather_stream, b->ev_copy[buf], 0));  // staging buffer is free again
        g.first_inst = first + off;
        g.n_inst = cnt;
        g.out_be = out ? b->d_stage[buf] : nullptr;
        g.out_present = out_present ? b->d_stage_present[buf] : nullptr;
        CUDA_TRY(launch_gather_outputs(g, ctx->gather_stream));
        CUDA_TRY(cudaEventRecord(b->ev_gather[buf], ctx->gather_stream));
        CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, b->ev_gather[buf], 0));
        if (out)
            CUDA_TRY(cudaMemcpyAsync(out + (size_t)off * row, b->d_stage[buf], (size_t)cnt * row, cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (out_present)
            CUDA_TRY(cudaMemcpyAsync(out_present + (size_t)off * n_out, b->d_stage_present[buf], (size_t)cnt * n_out,
                                     cudaMemcpyDeviceToHost, ctx->copy_stream));
        CUDA_TRY(cudaEventRecord(b->ev_copy[buf], ctx->copy_stream));
        c->run.kernel_launches += 1;
    }
    CUDA_TRY(cudaEventRecord(b->ev1, ctx->gather_stream));
    if (async) {
        CUDA_TRY(cudaEventRecord(b->ev_dl_done, ctx->copy_stream));
        b->dl_pending = true;
        return ACVMB_OK;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->gather_stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, b->ev0, b->ev1);
    c->run.gather_ms += ms;
    return ACVMB_OK;
}

extern "C" int acvmb_batch_download(acvmb_batch* b, uint32_t first, uint32_t n, const uint32_t* out_ids, uint32_t n_out_ids,
                                    uint8_t* out) {
    if (!out) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    return acvmb_batch_download_ex(b, first, n, out_ids, n_out_ids, out, nullptr);
}

extern "C" int acvmb_batch_checksum(acvmb_batch* b, uint64_t* out) {
    if (!b || !out) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    acvmb_circuit* c = b->c;
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    GatherArgs g{};
    g.cols = b->d_cols;
    g.n_slots = c->plan.n_slots;
    g.T = (int)b->T;
    g.n_out = c->plan.num_witnesses;
    g.n_inst = b->n_inst;
    g.fail = b->d_fail;
    g.assign_opcode = c->d_assign;
    g.mu_index_of = c->d_mu_index_of;
    g.mu_assign = b->d_mu;
    g.n_mu = c->plan.n_mu;
    g.static_fail_opcode = c->plan.static_fail.present ? c->plan.static_fail.opcode : 0xFFFFFFFFu;
    g.unscale = c->d_unscale;
    unsigned long long* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, (size_t)b->n_inst * 8));
    cudaError_t e = launch_checksum(g, d_out, c->ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)b->n_inst * 8, cudaMemcpyDeviceToHost, c->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->ctx->stream);
    cudaFree(d_out);
    CUDA_TRY(e);
    c->run.kernel_launches += 1;
    return ACVMB_OK;
}

// ---------------------------------------------------------------------------------------------
static uint32_t resident_instances(acvmb_circuit* c, uint32_t batch, uint32_t T, uint32_t n_out) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    uint64_t budget = c->ctx->max_resident_bytes ? c->ctx->max_resident_bytes : (uint64_t)((free_b + c->ctx->cached_bytes) * 0.90);
    uint64_t fixed = 4 * std::min<uint64_t>(c->ctx->staging_bytes, (uint64_t)batch * n_out * 32) + (64ull << 20);   // two buffers x double staging
    uint64_t per_inst = (uint64_t)c->plan.n_slots * 32 + 8 + c->plan.input_witnesses.size() * 32;
    uint64_t fit = budget > fixed ? (budget - fixed) / per_inst : 0;
    fit = (fit / T) * T;
    if (fit < T) fit = T;
    return (uint32_t)std::min<uint64_t>(fit, batch);
}

extern "C" int acvmb_solve_batch_ex(acvmb_circuit* c, uint32_t batch, const uint8_t* inputs_be32, const uint32_t* out_ids,
                                    uint32_t n_out_ids, uint8_t* out_witness, uint8_t* out_present, acvmb_status* out_status);

extern "C" int acvmb_solve_batch(acvmb_circuit* c, uint32_t batch, const uint8_t* inputs_be32, const uint32_t* out_ids,
                                 uint32_t n_out_ids, uint8_t* out_witness, acvmb_status* out_status) {
    return acvmb_solve_batch_ex(c, batch, inputs_be32, out_ids, n_out_ids, out_witness, nullptr, out_status);
}

static int solve_batch_single(acvmb_circuit* c, uint32_t batch, const uint8_t* inputs_be32, const uint32_t* out_ids,
                              uint32_t n_out_ids, uint8_t* out_witness, uint8_t* out_present, acvmb_status* out_status);

extern "C" int acvmb_solve_batch_ex(acvmb_circuit* c, uint32_t batch, const uint8_t* inputs_be32, const uint32_t* out_ids,
                                    uint32_t n_out_ids, uint8_t* out_witness, uint8_t* out_present, acvmb_status* out_status) {
    if (!c) return set_err(ACVMB_ERR_INVALID_ARG, "circuit is NULL");
    if (batch == 0) return ACVMB_OK;
    if (c->shards.empty()) return solve_batch_single(c, batch, inputs_be32, out_ids, n_out_ids, out_witness, out_present, out_status);
    // multi-device context: contiguous instance ranges, one host thread per device, no traffic between devices
    std::vector<acvmb_circuit*> all = {c};
    all.insert(all.end(), c->shards.begin(), c->shards.end());
    const uint32_t n_dev = (uint32_t)all.size();
    const uint32_t T = pick_T(c->ctx, c->plan);
    uint32_t per = (batch + n_dev - 1) / n_dev;
    per = ((per + T - 1) / T) * T;   // whole tiles per device
    const size_t n_in = c->plan.input_witnesses.size();
    const uint32_t n_out = out_ids ? n_out_ids : c->plan.num_witnesses;
    std::vector<int> rcs(n_dev, ACVMB_OK);
    std::vector<std::string> errs(n_dev);
    auto work = [&](uint32_t d) {
        const uint32_t lo = std::min<uint64_t>((uint64_t)d * per, batch), hi = std::min<uint64_t>((uint64_t)(d + 1) * per, batch);
        memset(&all[d]->run, 0, sizeof(all[d]->run));
        if (hi == lo) return;
        rcs[d] = solve_batch_single(all[d], hi - lo, inputs_be32 ? inputs_be32 + (size_t)lo * n_in * 32 : nullptr, out_ids, n_out_ids,
                                    out_witness ? out_witness + (size_t)lo * n_out * 32 : nullptr,
                                    out_present ? out_present + (size_t)lo * n_out : nullptr, out_status ? out_status + lo : nullptr);
        if (rcs[d]) errs[d] = g_last_error;
    };
    std::vector<std::thread> th;
    for (uint32_t d = 1; d < n_dev; ++d) th.emplace_back(work, d);
    work(0);
    for (auto& t : th) t.join();
    cudaSetDevice(c->ctx->device);
    acvmb_run_info agg = c->run;
    for (uint32_t d = 1; d < n_dev; ++d) {   // run record: device times are the max over devices, launches the sum
        const acvmb_run_info& r = all[d]->run;
        agg.kernel_ms = std::max(agg.kernel_ms, r.kernel_ms);
        agg.scatter_ms = std::max(agg.scatter_ms, r.scatter_ms);
        agg.gather_ms = std::max(agg.gather_ms, r.gather_ms);
        agg.kernel_launches += r.kernel_launches;
        agg.n_tiles += r.n_tiles;
        agg.n_subbatches = std::max(agg.n_subbatches, r.n_subbatches);
    }
    c->run = agg;
    for (uint32_t d = 0; d < n_dev; ++d)
        if (rcs[d]) return set_err(rcs[d], "device " + std::to_string(all[d]->ctx->device) + ": " + errs[d]);
    return ACVMB_OK;
}

static int solve_batch_single(acvmb_circuit* c, uint32_t batch, const uint8_t* inputs_be32, const uint32_t* out_ids,
                              uint32_t n_out_ids, uint8_t* out_witness, uint8_t* out_present, acvmb_status* out_status) {
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    memset(&c->run, 0, sizeof(c->run));
    uint32_t T = pick_T(c->ctx, c->plan);
    uint32_t n_out = out_ids ? n_out_ids : c->plan.num_witnesses;
    const bool want_out = out_witness || out_present;
    uint32_t resident = resident_instances(c, batch, T, want_out ? n_out : 0);
    size_t n_in = c->plan.input_witnesses.size();
    int rc = ACVMB_OK;
    uint32_t n_sub = 0;
    const uint64_t out_bytes = (uint64_t)batch * n_out * (out_witness ? 32 : 1);
    if (want_out && resident >= 2 * T && batch >= 2 * T && (batch > resident || out_bytes >= (256ull << 20))) {
        // Every witness map goes back over PCIe (~55 GB/s), which takes ~10x longer than solving it.  Two column buffers:
        // while piece k drains (gather + D2H on their own streams), the VM kernel of piece k+1 runs on the VM stream.  A batch
        // that does not fit HBM is cut into halves of the resident size, one that does into four pieces.
        acvmb_ctx* ctx = c->ctx;
        uint32_t piece = batch > resident ? ((resident / 2) / T) * T : (((batch + 3) / 4 + T - 1) / T) * T;
        if (piece < T) piece = T;
        acvmb_batch** bufs = ctx->pipe_batch;
        const bool reuse = bufs[0] && bufs[1] && bufs[0]->c == c && bufs[0]->T == T && bufs[0]->capacity == piece && bufs[1]->capacity == piece;
        if (!reuse) drop_cached_batch(ctx);   // (resident_instances() budgeted every cached column buffer as free memory)
        else if (ctx->cached_batch) { acvmb_batch_destroy(ctx->cached_batch); ctx->cached_batch = nullptr; }
        for (uint32_t off = 0; off < batch && rc == ACVMB_OK; off += piece, ++n_sub) {
            const uint32_t cnt = std::min(piece, batch - off);
            const int i = n_sub & 1;
            rc = wait_download(bufs[i]);   // its previous piece has left the device
            if (rc) break;
            if (!bufs[i]) {
                rc = acvmb_batch_create(c, piece, &bufs[i]);
                if (rc) break;
            }
            acvmb_batch* b = bufs[i];
            rc = acvmb_batch_resize(b, cnt);
            if (rc) break;
            rc = acvmb_batch_upload(b, inputs_be32 ? inputs_be32 + (size_t)off * n_in * 32 : nullptr);
            if (rc) break;
            rc = acvmb_batch_run(b, nullptr);
            if (rc) break;
            if (out_status) {
                rc = acvmb_batch_status(b, out_status + off);
                if (rc) break;
            }
            rc = download_impl(b, 0, cnt, out_ids, n_out_ids, out_witness ? out_witness + (size_t)off * n_out * 32 : nullptr,
                               out_present ? out_present + (size_t)off * n_out : nullptr, /*async=*/true);
        }
        for (int i = 0; i < 2; ++i) {
            int rc2 = wait_download(bufs[i]);
            if (!rc) rc = rc2;
        }
        if (rc != ACVMB_OK || !ctx->opt_cache_batch) drop_cached_batch(ctx);
        else ctx->cached_bytes = 2ull * ((uint64_t)(piece + T - 1) / T) * T * c->plan.n_slots * 32;
        c->run.resident_instances = piece;
        c->run.n_subbatches = n_sub;
        return rc;
    }
    acvmb_batch* b = nullptr;
    uint32_t cap = 0;
    {
        acvmb_ctx* ctx = c->ctx;
        const uint32_t first_cnt = std::min(resident, batch);
        if (ctx->cached_batch && ctx->cached_batch->c == c && ctx->cached_batch->n_inst == first_cnt && ctx->cached_batch->T == T) {
            b = ctx->cached_batch;   // same circuit, same size as the previous call: reuse the columns
            cap = first_cnt;
            ctx->cached_batch = nullptr;
            ctx->cached_bytes = 0;
        } else {
            drop_cached_batch(ctx);
        }
        for (int i = 0; i < 2; ++i) {   // the pipelined path's buffers are not used here
            if (ctx->pipe_batch[i]) acvmb_batch_destroy(ctx->pipe_batch[i]);
            ctx->pipe_batch[i] = nullptr;
        }
    }
    // equal sub-batches: a launch of the step-VM kernel is latency bound (its time hardly depends on the instance count), so
    // [r, r, r, small remainder] costs a whole extra pass; ceil(batch / n) per pass does not
    if (batch > resident) {
        const uint32_t n_pass = (batch + resident - 1) / resident;
        const uint32_t even = (((batch + n_pass - 1) / n_pass + T - 1) / T) * T;
        if (even <= resident) resident = even;
    }
    for (uint32_t off = 0; off < batch && rc == ACVMB_OK; off += resident, ++n_sub) {
        uint32_t cnt = std::min(resident, batch - off);
        if (!b || cnt != cap) {
            if (b) acvmb_batch_destroy(b);
            b = nullptr;
            rc = acvmb_batch_create(c, cnt, &b);
            if (rc) break;
            cap = cnt;
        }
        rc = acvmb_batch_upload(b, inputs_be32 ? inputs_be32 + (size_t)off * n_in * 32 : nullptr);
        if (rc) break;
        rc = acvmb_batch_run(b, nullptr);
        if (rc) break;
        if (out_status) {
            rc = acvmb_batch_status(b, out_status + off);
            if (rc) break;
        }
        if (want_out)
            rc = acvmb_batch_download_ex(b, 0, cnt, out_ids, n_out_ids, out_witness ? out_witness + (size_t)off * n_out * 32 : nullptr,
                                         out_present ? out_present + (size_t)off * n_out : nullptr);
    }
    if (b && rc == ACVMB_OK && c->ctx->opt_cache_batch) {
        c->ctx->cached_batch = b;
        c->ctx->cached_bytes = (uint64_t)b->n_tiles * b->T * c->plan.n_slots * 32;
    } else if (b) {
        acvmb_batch_destroy(b);
    }
    c->run.resident_instances = resident;
    c->run.n_subbatches = n_sub;
    return rc;
}

extern "C" int acvmb_last_run_info(const acvmb_circuit* c, acvmb_run_info* out) {
    if (!c || !out) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    *out = c->run;
    return ACVMB_OK;
}

// ---------------------------------------------------------------------------------------------
// single-instance ACVM mirror (batch of 1)
// ---------------------------------------------------------------------------------------------
extern "C" int acvmb_vm_new(acvmb_ctx* ctx, const uint8_t* gz, size_t len, const uint32_t* widx, const uint8_t* wval,
                            uint32_t n_initial, acvmb_vm** out) {
    if (!ctx || !gz || !out || (n_initial && (!widx || !wval))) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    auto vm = std::make_unique<acvmb_vm>();
    int rc = acvmb_circuit_from_acir(ctx, gz, len, widx, n_initial, &vm->c);
    if (rc) return rc;
    vm->inputs.assign(wval, wval + (size_t)n_initial * 32);
    vm->assign = vm->c->plan.assign_opcode;
    // ACVM::new: Solved when there are no opcodes (mod.rs:147)
    vm->status.code = vm->c->plan.n_opcodes == 0 ? ACVMB_SOLVED : ACVMB_IN_PROGRESS;
    *out = vm.release();
    return ACVMB_OK;
}

extern "C" void acvmb_vm_destroy(acvmb_vm* vm) {
    if (!vm) return;
    if (vm->b) acvmb_batch_destroy(vm->b);
    if (vm->c) acvmb_circuit_destroy(vm->c);
    delete vm;
}

static int vm_ensure_solved(acvmb_vm* vm) {
    if (vm->solved_once) return ACVMB_OK;
    vm->witness.assign((size_t)vm->c->plan.num_witnesses * 32, 0);
    vm->present.assign((size_t)vm->c->plan.num_witnesses, 0);
    acvmb_batch* b = nullptr;
    int rc = acvmb_batch_create(vm->c, 1, &b);
    if (rc) return rc;
    rc = acvmb_batch_upload(b, vm->inputs.data());
    if (!rc) rc = acvmb_batch_run(b, nullptr);
    if (!rc) rc = acvmb_batch_status(b, &vm->final_status);
    if (!rc) rc = acvmb_batch_download_ex(b, 0, 1, nullptr, 0, vm->witness.data(), vm->present.data());
    vm->mu.assign(vm->c->plan.n_mu, 0xFFFFFFFFu);
    if (!rc && vm->c->plan.n_mu) {   // lane 0 of tile 0: stride T words
        std::vector<uint32_t> all((size_t)vm->c->plan.n_mu * b->T);
        cudaError_t e = cudaMemcpy(all.data(), b->d_mu, all.size() * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = set_err(ACVMB_ERR_CUDA, cudaGetErrorString(e));
        else for (uint32_t m = 0; m < vm->c->plan.n_mu; ++m) vm->mu[m] = all[(size_t)m * b->T];
    }
    vm->fc_pending = b->fc_pending && vm->final_status.code == ACVMB_REQUIRES_FOREIGN_CALL;
    vm->fc_function = b->fc_function;
    vm->fc_inputs = b->fc_inputs;
    acvmb_batch_destroy(b);
    if (rc) return rc;
    vm->solved_once = true;
    return ACVMB_OK;
}
// opcodes [0, vm_limit) execute successfully; the recorded outcome belongs to opcode vm_limit
static uint32_t vm_limit(const acvmb_vm* vm) {
    return vm->final_status.code == ACVMB_SOLVED ? vm->c->plan.n_opcodes : vm->final_status.opcode_index;
}

extern "C" int acvmb_vm_solve(acvmb_vm* vm, acvmb_status* out) {
    if (!vm) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    if (vm->c->plan.n_opcodes) {   // ACVM::new is already Solved without opcodes (mod.rs:147)
        int rc = vm_ensure_solved(vm);
        if (rc) return rc;
        vm->status = vm->final_status;
        vm->ip = vm_limit(vm);
    }
    if (out) *out = vm->status;
    return ACVMB_OK;
}

// ACVM::solve_opcode (acvm/src/pwg/mod.rs:243-303): execute the opcode at the instruction pointer.  Solved: the reference
// indexes past the end of `opcodes` and panics -> ACVMB_ERR_STATE.  Failure / an unresolved foreign call: the same opcode is
// attempted again with the same outcome.
extern "C" int acvmb_vm_solve_opcode(acvmb_vm* vm, acvmb_status* out) {
    if (!vm) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    if (vm->status.code == ACVMB_SOLVED) return set_err(ACVMB_ERR_STATE, "no opcode left to solve (the reference panics on the index)");
    int rc = vm_ensure_solved(vm);
    if (rc) return rc;
    const uint32_t limit = vm_limit(vm), n_op = vm->c->plan.n_opcodes;
    if (vm->ip < limit) {
        ++vm->ip;
        if (vm->ip == n_op) vm->status = vm->final_status;   // Solved (limit == n_opcodes only then)
        else vm->status = acvmb_status{ACVMB_IN_PROGRESS, ACVMB_E_NONE, vm->ip, 0};
    } else {
        vm->status = vm->final_status;
    }
    if (out) *out = vm->status;
    return ACVMB_OK;
}

extern "C" int acvmb_vm_status(const acvmb_vm* vm, acvmb_status* out) {
    if (!vm || !out) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    *out = vm->status;
    return ACVMB_OK;
}

extern "C" int acvmb_vm_instruction_pointer(const acvmb_vm* vm, uint32_t* out) {
    if (!vm || !out) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    *out = vm->ip;
    return ACVMB_OK;
}

extern "C" int acvmb_vm_num_witnesses(const acvmb_vm* vm, uint32_t* out) {
    if (!vm || !out) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    *out = vm->c->plan.num_witnesses;
    return ACVMB_OK;
}

static bool vm_present(const acvmb_vm* vm, uint32_t w) {
    uint32_t ao = vm->assign[w];
    if (ao == 0xFFFFFFFEu) return true;                       // initial witness
    if (!vm->solved_once || ao == 0xFFFFFFFFu) return false;
    if (vm->ip >= vm_limit(vm)) return vm->present[w] != 0;   // every opcode that runs has run: the device's own answer
    if (ao == 0xFFFFFFFDu) {                                  // value-dependent: this instance's "assigned by opcode" word
        uint32_t m = vm->c->plan.mu_index_of[w];
        ao = m < vm->mu.size() ? vm->mu[m] : 0xFFFFFFFFu;
        if (ao == 0xFFFFFFFFu) return false;
    }
    return ao < vm->ip;
}

extern "C" int acvmb_vm_witness(const acvmb_vm* vm, uint32_t w, uint8_t out[32], int* present) {
    if (!vm || !out || !present) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    if (w >= vm->c->plan.num_witnesses) {
        *present = 0;
        memset(out, 0, 32);
        return ACVMB_OK;
    }
    *present = vm_present(vm, w) ? 1 : 0;
    if (*present && !vm->solved_once) {
        // initial witness before solve(): answer from the inputs
        const auto& in = vm->c->plan.input_witnesses;
        for (size_t i = 0; i < in.size(); ++i)
            if (in[i] == w) memcpy(out, vm->inputs.data() + i * 32, 32);
        return ACVMB_OK;
    }
    if (*present) memcpy(out, vm->witness.data() + (size_t)w * 32, 32); else memset(out, 0, 32);
    return ACVMB_OK;
}

extern "C" int acvmb_vm_finalize(acvmb_vm* vm, uint8_t* out, uint8_t* present, uint32_t n) {
    if (!vm || !out || !present) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    if (vm->status.code != ACVMB_SOLVED) return set_err(ACVMB_ERR_STATE, "ACVM is not ready to be finalized");  // mod.rs:177-179
    if (n != vm->c->plan.num_witnesses) return set_err(ACVMB_ERR_INVALID_ARG, "n must equal num_witnesses");
    if (!vm->solved_once) {  // no opcodes: the map is the initial witness
        vm->witness.assign((size_t)n * 32, 0);
        vm->present.assign((size_t)n, 0);
        const auto& in = vm->c->plan.input_witnesses;
        for (size_t i = 0; i < in.size(); ++i) {
            memcpy(vm->witness.data() + (size_t)in[i] * 32, vm->inputs.data() + i * 32, 32);
            vm->present[in[i]] = 1;
        }
        vm->solved_once = true;
    }
    memcpy(out, vm->witness.data(), (size_t)n * 32);
    for (uint32_t w = 0; w < n; ++w) present[w] = vm_present(vm, w) ? 1 : 0;
    return ACVMB_OK;
}

// ---------------------------------------------------------------------------------------------
// BlackBoxFunctionSolver trait, batched: thin wrappers that build a one-opcode circuit
// ---------------------------------------------------------------------------------------------
static int run_single_opcode(acvmb_ctx* ctx, const Opcode& op, uint32_t n_witnesses, const std::vector<uint32_t>& inputs,
                             const std::vector<uint32_t>& outs, const uint8_t* in_be, uint32_t batch, uint8_t* out_be,
                             acvmb_status* st) {
    Circuit circ;
    circ.current_witness_index = n_witnesses - 1;
    circ.opcodes.push_back(op);
    acvmb_circuit* c = nullptr;
    int rc = circuit_from_struct(ctx, circ, inputs.data(), (uint32_t)inputs.size(), &c);
    if (rc) return rc;
    rc = acvmb_solve_batch(c, batch, in_be, outs.data(), (uint32_t)outs.size(), out_be, st);
    acvmb_circuit_destroy(c);
    return rc;
}

extern "C" int acvmb_fixed_base_scalar_mul(acvmb_ctx* ctx, const uint8_t* low, const uint8_t* high, uint32_t batch, uint8_t* out_xy,
                                           acvmb_status* st) {
    if (!ctx || !low || !high || !out_xy) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    Opcode op;
    op.kind = OP_BlackBox;
    op.bb().func = BB_FixedBaseScalarMul;
    op.bb().inputs = {{1, 128}, {2, 128}};
    op.bb().outputs = {3, 4};
    std::vector<uint8_t> in((size_t)batch * 64);
    for (uint32_t i = 0; i < batch; ++i) {
        memcpy(&in[(size_t)i * 64], low + (size_t)i * 32, 32);
        memcpy(&in[(size_t)i * 64 + 32], high + (size_t)i * 32, 32);
    }
    return run_single_opcode(ctx, op, 5, {1, 2}, {3, 4}, in.data(), batch, out_xy, st);
}

extern "C" int acvmb_pedersen(acvmb_ctx* ctx, const uint8_t* inputs, uint32_t n_inputs, uint32_t batch, uint32_t domain_separator,
                              uint8_t* out_xy, acvmb_status* st) {
    if (!ctx || (!inputs && n_inputs) || !out_xy) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    if (!ctx->opt_pedersen_unpinned)
        return set_err(ACVMB_ERR_UNSUPPORTED, "acvmb_pedersen: parity with barretenberg's generator tables is unpinned (both reference KATs "
                                              "fail); set the context option pedersen_unpinned=1 to run the structurally identical kernel");
    Opcode op;
    op.kind = OP_BlackBox;
    op.bb().func = BB_Pedersen;
    std::vector<uint32_t> in_ids;
    for (uint32_t i = 0; i < n_inputs; ++i) {
        op.bb().inputs.push_back({i + 1, 254});
        in_ids.push_back(i + 1);
    }
    op.bb().n_message_inputs = n_inputs;
    op.bb().domain_separator = domain_separator;
    op.bb().outputs = {n_inputs + 1, n_inputs + 2};
    return run_single_opcode(ctx, op, n_inputs + 3, in_ids, {n_inputs + 1, n_inputs + 2}, inputs, batch, out_xy, st);
}

static int hash_bytes(acvmb_ctx* ctx, uint32_t func, const uint8_t* msgs, uint32_t msg_len, uint32_t batch, uint8_t* digests) {
    if (!ctx || (!msgs && msg_len) || !digests) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    Opcode op;
    op.kind = OP_BlackBox;
    op.bb().func = func;
    std::vector<uint32_t> in_ids, out_ids;
    for (uint32_t i = 0; i < msg_len; ++i) {
        op.bb().inputs.push_back({i + 1, 8});
        in_ids.push_back(i + 1);
    }
    op.bb().n_message_inputs = msg_len;
    for (uint32_t i = 0; i < 32; ++i) {
        op.bb().outputs.push_back(msg_len + 1 + i);
        out_ids.push_back(msg_len + 1 + i);
    }
    // one byte per witness, like the ACIR opcode (hash.rs:51-66)
    std::vector<uint8_t> in((size_t)batch * msg_len * 32, 0);
    for (size_t i = 0; i < (size_t)batch * msg_len; ++i) in[i * 32 + 31] = msgs[i];
    std::vector<uint8_t> out((size_t)batch * 32 * 32);
    std::vector<acvmb_status> st(batch);
    int rc = run_single_opcode(ctx, op, msg_len + 33, in_ids, out_ids, in.data(), batch, out.data(), st.data());
    if (rc) return rc;
    for (size_t i = 0; i < (size_t)batch * 32; ++i) digests[i] = out[i * 32 + 31];
    return ACVMB_OK;
}

extern "C" int acvmb_sha256(acvmb_ctx* ctx, const uint8_t* msgs, uint32_t msg_len, uint32_t batch, uint8_t* digests) {
    return hash_bytes(ctx, BB_SHA256, msgs, msg_len, batch, digests);
}
extern "C" int acvmb_keccak256(acvmb_ctx* ctx, const uint8_t* msgs, uint32_t msg_len, uint32_t batch, uint8_t* digests) {
    return hash_bytes(ctx, BB_Keccak256, msgs, msg_len, batch, digests);
}

static int ecdsa_bytes(acvmb_ctx* ctx, uint32_t func, const uint8_t* hashed, const uint8_t* pkx, const uint8_t* pky, const uint8_t* sig,
                       uint32_t batch, uint8_t* out_valid, acvmb_status* st) {
    if (!ctx || !hashed || !pkx || !pky || !sig || !out_valid || !st) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    Opcode op;
    op.kind = OP_BlackBox;
    op.bb().func = func;
    std::vector<uint32_t> in_ids;
    for (uint32_t i = 0; i < 160; ++i) {   // get_inputs_vec order: pkx, pky, signature, hashed message
        op.bb().inputs.push_back({i + 1, 8});
        in_ids.push_back(i + 1);
    }
    op.bb().seg[0] = op.bb().seg[1] = op.bb().seg[3] = 32;
    op.bb().seg[2] = 64;
    op.bb().outputs = {161};
    std::vector<uint8_t> in((size_t)batch * 160 * 32, 0);
    for (size_t i = 0; i < batch; ++i) {
        uint8_t* row = &in[i * 160 * 32];
        for (int k = 0; k < 32; ++k) row[(size_t)k * 32 + 31] = pkx[i * 32 + k];
        for (int k = 0; k < 32; ++k) row[(size_t)(32 + k) * 32 + 31] = pky[i * 32 + k];
        for (int k = 0; k < 64; ++k) row[(size_t)(64 + k) * 32 + 31] = sig[i * 64 + k];
        for (int k = 0; k < 32; ++k) row[(size_t)(128 + k) * 32 + 31] = hashed[i * 32 + k];
    }
    std::vector<uint8_t> out((size_t)batch * 32);
    int rc = run_single_opcode(ctx, op, 162, in_ids, {161}, in.data(), batch, out.data(), st);
    if (rc) return rc;
    for (size_t i = 0; i < batch; ++i) out_valid[i] = st[i].code == ACVMB_SOLVED ? out[i * 32 + 31] : 0;
    return ACVMB_OK;
}
extern "C" int acvmb_ecdsa_secp256k1_verify(acvmb_ctx* ctx, const uint8_t* hashed, const uint8_t* pkx, const uint8_t* pky,
                                            const uint8_t* sig, uint32_t batch, uint8_t* out_valid, acvmb_status* st) {
    return ecdsa_bytes(ctx, BB_EcdsaSecp256k1, hashed, pkx, pky, sig, batch, out_valid, st);
}
extern "C" int acvmb_ecdsa_secp256r1_verify(acvmb_ctx* ctx, const uint8_t* hashed, const uint8_t* pkx, const uint8_t* pky,
                                            const uint8_t* sig, uint32_t batch, uint8_t* out_valid, acvmb_status* st) {
    return ecdsa_bytes(ctx, BB_EcdsaSecp256r1, hashed, pkx, pky, sig, batch, out_valid, st);
}

extern "C" int acvmb_imad_microbench(acvmb_ctx* ctx, double* a, double* b, double* c, double* mhz) {
    if (!ctx) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(imad_microbench(a, b, c, mhz));
    return ACVMB_OK;
}

// ---------------------------------------------------------------------------------------------
// host-only entry point (no device needed): decode + compile, return info and the plan blob.
// Used by the CPU test-suite to check the decoder and the plan compiler without a GPU.
// ---------------------------------------------------------------------------------------------
extern "C" int acvmb_plan_compile_host_ex(const uint8_t* gz, size_t len, const uint32_t* input_witnesses, uint32_t n_inputs, uint32_t S,
                                          uint32_t temp_pool, uint32_t flags, acvmb_plan_info* info, uint8_t* blob, size_t cap,
                                          size_t* needed) {
    if (!gz || (n_inputs && !input_witnesses)) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    Circuit circ;
    try {
        circ = decode_circuit(gz, len);
    } catch (const std::exception& e) {
        return set_err(ACVMB_ERR_DECODE, e.what());
    }
    acvmb_circuit tmp;
    PlanOptions opt;
    opt.S = S ? S : 16;
    if (temp_pool) opt.temp_pool = temp_pool;
    opt.allow_unpinned_pedersen = (flags & 1u) != 0;
    opt.device_brillig = (flags & 2u) == 0;
    opt.scaled_columns = (flags & 4u) == 0;
    opt.packed_hashes = (flags & 8u) == 0;
    opt.spread_heavy = (flags & 16u) == 0;
    opt.slack_scheduling = (flags & 32u) == 0;
    if (const uint32_t rs = (flags >> 8) & 0xFFFFu) opt.ring_slots = rs == 0xFFFFu ? 0u : rs;
    try {
        tmp.plan = compile_plan(circ, std::vector<uint32_t>(input_witnesses, input_witnesses + n_inputs), opt);
    } catch (const std::exception& e) {
        return set_err(ACVMB_ERR_UNSUPPORTED, e.what());
    }
    if (info) acvmb_circuit_info(&tmp, info);
    if (needed || blob) {
        std::vector<uint8_t> b = serialize_plan(tmp.plan);
        if (needed) *needed = b.size();
        if (blob) {
            if (cap < b.size()) return set_err(ACVMB_ERR_INVALID_ARG, "buffer too small");
            memcpy(blob, b.data(), b.size());
        }
    }
    return ACVMB_OK;
}
extern "C" int acvmb_plan_compile_host(const uint8_t* gz, size_t len, const uint32_t* input_witnesses, uint32_t n_inputs, uint32_t S,
                                       acvmb_plan_info* info, uint8_t* blob, size_t cap, size_t* needed) {
    return acvmb_plan_compile_host_ex(gz, len, input_witnesses, n_inputs, S, 0, 0, info, blob, cap, needed);
}

// measurement helper: bare device->host rate of this context's device into the caller's (pinned) buffer -- the ceiling of
// every end-to-end number that returns witness maps
extern "C" int acvmb_d2h_microbench(acvmb_ctx* ctx, void* host, size_t bytes, uint32_t reps, double* gb_per_s) {
    if (!ctx || !host || !bytes || !reps || !gb_per_s) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    void* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t e = cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream);   // warm-up
    if (e == cudaSuccess) e = cudaEventRecord(e0, ctx->copy_stream);
    for (uint32_t i = 0; i < reps && e == cudaSuccess; ++i) e = cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(e1, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float ms = 0;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    CUDA_TRY(e);
    *gb_per_s = (double)bytes * reps / (ms * 1e-3) / 1e9;
    return ACVMB_OK;
}

extern "C" int acvmb_imad_cc_microbench(acvmb_ctx* ctx, double* out3) {
    if (!ctx || !out3) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(run_imad_cc_microbench(out3));
    return ACVMB_OK;
}

extern "C" int acvmb_frmul_microbench(acvmb_ctx* ctx, double* fr_mul_per_s /*[5]*/) {
    if (!ctx) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(frmul_microbench(fr_mul_per_s));
    return ACVMB_OK;
}

// host-only test hook: generator `index` of the Pedersen tables (canonical x, y big-endian) -- lets the CPU
// suite check the C++ derivation against the oracle's without a device
extern "C" int acvmb_pedersen_generator_host(uint32_t index, uint8_t out_xy_be32[64]) {
    if (!out_xy_be32) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    gk::Pt p = gk::derive_pedersen_generator(index);
    hf::to_be_bytes(p.x, out_xy_be32);
    hf::to_be_bytes(p.y, out_xy_be32 + 32);
    return ACVMB_OK;
}

// host-only test hook: control bits of the permutation network that maps 0..n-1 onto `outputs` (sorting.rs:164-235)
extern "C" int acvmb_permutation_route_host(const uint32_t* outputs, uint32_t n, uint8_t* bits, uint32_t cap, uint32_t* n_bits) {
    if ((!outputs && n) || !n_bits) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    std::vector<uint32_t> base(n), out(outputs, outputs + n);
    for (uint32_t i = 0; i < n; ++i) base[i] = i;
    std::vector<uint8_t> b;
    if (!psort::route(base, out, n, b)) return set_err(ACVMB_ERR_STATE, "outputs are not a permutation of 0..n-1 (the reference panics)");
    *n_bits = (uint32_t)b.size();
    if (bits) memcpy(bits, b.data(), std::min<size_t>(cap, b.size()));
    return ACVMB_OK;
}

// host-only test hook: run Brillig opcode `opcode_index` of a circuit on the C++ host VM with explicit input values
// (flattened in input order) and return the flattened outputs -- lets the CPU suite compare the VM with the oracle's.
extern "C" int acvmb_brillig_run_host(const uint8_t* gz, size_t len, uint32_t opcode_index, const uint8_t* in_values_be32,
                                      uint32_t n_in_values, uint8_t* out_values_be32, uint32_t n_out_values, uint32_t* status,
                                      uint32_t* fail_pc) {
    if (!gz || !status) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    Circuit circ;
    try {
        circ = decode_circuit(gz, len);
    } catch (const std::exception& e) {
        return set_err(ACVMB_ERR_DECODE, e.what());
    }
    if (opcode_index >= circ.opcodes.size() || circ.opcodes[opcode_index].kind != OP_Brillig)
        return set_err(ACVMB_ERR_INVALID_ARG, "not a Brillig opcode");
    const Brillig& br = circ.opcodes[opcode_index].brillig();
    bvm::VM vm;
    uint32_t pos = 0;
    for (auto& in : br.inputs) {
        if (pos + in.exprs.size() > n_in_values) return set_err(ACVMB_ERR_INVALID_ARG, "too few input values");
        if (!in.is_array) {
            vm.regs.push_back(hf::from_be_bytes_reduce(in_values_be32 + (size_t)pos * 32, 32));
            ++pos;
        } else {
            vm.regs.push_back(hf::from_u64(vm.mem.size()));
            for (size_t k = 0; k < in.exprs.size(); ++k, ++pos) vm.mem.push_back(hf::from_be_bytes_reduce(in_values_be32 + (size_t)pos * 32, 32));
        }
    }
    bvm::Result r = vm.run(br);
    *status = (uint32_t)r.status;
    if (fail_pc) *fail_pc = r.call_stack.empty() ? 0 : (uint32_t)r.call_stack.back();
    if (r.status != bvm::Status::Finished) return ACVMB_OK;
    uint32_t k = 0;
    for (size_t oi = 0; oi < br.outputs.size(); ++oi) {
        U256 reg = oi < vm.regs.size() ? vm.regs[oi] : U256{};
        if (!br.outputs[oi].is_array) {
            if (k < n_out_values) hf::to_be_bytes(reg, out_values_be32 + (size_t)k * 32);
            ++k;
        } else {
            for (size_t j = 0; j < br.outputs[oi].witnesses.size(); ++j, ++k) {
                size_t p = (size_t)reg.l[0] + j;
                if ((reg.l[1] | reg.l[2] | reg.l[3]) || p >= vm.mem.size()) {
                    *status = (uint32_t)bvm::Status::Panic;
                    return ACVMB_OK;
                }
                if (k < n_out_values) hf::to_be_bytes(vm.mem[p], out_values_be32 + (size_t)k * 32);
            }
        }
    }
    return ACVMB_OK;
}

// ACVM::get_pending_foreign_call (acvm/src/pwg/mod.rs:197-203): name + resolved inputs of the call the VM is waiting on.
// input_lens[i] = number of values of input i; values are flattened in order.
extern "C" int acvmb_vm_pending_foreign_call(const acvmb_vm* vm, char* function, size_t function_cap, uint32_t* n_inputs,
                                             uint32_t* input_lens, uint32_t max_inputs, uint8_t* values_be32, uint32_t max_values,
                                             uint32_t* n_values) {
    if (!vm || !n_inputs || !n_values) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    if (!vm->fc_pending) return set_err(ACVMB_ERR_STATE, "no pending foreign call");
    if (function && function_cap) snprintf(function, function_cap, "%s", vm->fc_function.c_str());
    *n_inputs = (uint32_t)vm->fc_inputs.size();
    uint32_t k = 0;
    for (size_t i = 0; i < vm->fc_inputs.size(); ++i) {
        if (input_lens && i < max_inputs) input_lens[i] = (uint32_t)vm->fc_inputs[i].size();
        for (const U256& v : vm->fc_inputs[i]) {
            if (values_be32 && k < max_values) hf::to_be_bytes(v, values_be32 + (size_t)k * 32);
            ++k;
        }
    }
    *n_values = k;
    return ACVMB_OK;
}

// ACVM::resolve_pending_foreign_call (acvm/src/pwg/mod.rs:206-228): the result is appended to the Brillig opcode that made
// the call and execution resumes (the reference re-runs that opcode's VM from pc 0 and replays recorded results, :214-228;
// this mirror re-solves the whole instance, which is observationally the same).  out_lens[i] == 0xFFFFFFFF marks a
// ForeignCallOutput::Single, any other value an Array of that many values.
extern "C" int acvmb_vm_resolve_foreign_call(acvmb_vm* vm, uint32_t n_outputs, const uint32_t* out_lens, const uint8_t* values_be32) {
    if (!vm || (n_outputs && (!out_lens || !values_be32))) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    if (!vm->fc_pending || vm->status.code != ACVMB_REQUIRES_FOREIGN_CALL || !vm->c->has_circuit)
        return set_err(ACVMB_ERR_STATE, "ACVM is not expecting a foreign call response as no call was made");   // mod.rs:207-209 panics
    Opcode& op = vm->c->circuit.opcodes[vm->status.opcode_index];
    std::vector<ForeignCallOutput> res;
    size_t k = 0;
    for (uint32_t i = 0; i < n_outputs; ++i) {
        ForeignCallOutput o;
        o.is_array = out_lens[i] != 0xFFFFFFFFu;
        uint32_t cnt = o.is_array ? out_lens[i] : 1;
        for (uint32_t j = 0; j < cnt; ++j, ++k) o.values.push_back(hf::from_be_bytes_reduce(values_be32 + k * 32, 32));
        res.push_back(std::move(o));
    }
    op.brillig().foreign_call_results.push_back(std::move(res));
    vm->fc_pending = false;
    vm->solved_once = false;
    vm->ip = vm->status.opcode_index;   // the Brillig opcode is attempted again (mod.rs:214-228)
    vm->status = acvmb_status{ACVMB_IN_PROGRESS, 0, vm->status.opcode_index, 0};
    return ACVMB_OK;
}

// WitnessMap (de)compression: gzip(bincode(BTreeMap<Witness, FieldElement>)) (acir/src/native_types/witness_map.rs:108-146)
extern "C" int acvmb_witness_map_compress(const uint32_t* witness_idx, const uint8_t* values_be32, uint32_t n, uint8_t* out, size_t cap,
                                          size_t* needed) {
    if ((n && (!witness_idx || !values_be32)) || !needed) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    std::vector<std::pair<uint32_t, U256>> wm;
    for (uint32_t i = 0; i < n; ++i) wm.emplace_back(witness_idx[i], hf::from_be_bytes_reduce(values_be32 + (size_t)i * 32, 32));
    std::sort(wm.begin(), wm.end(), [](const auto& a, const auto& b) { return a.first < b.first; });   // BTreeMap order
    std::vector<uint8_t> gz;
    try {
        gz = encode_witness_map(wm);
    } catch (const std::exception& e) {
        return set_err(ACVMB_ERR_DECODE, e.what());
    }
    *needed = gz.size();
    if (!out) return ACVMB_OK;
    if (cap < gz.size()) return set_err(ACVMB_ERR_INVALID_ARG, "buffer too small");
    memcpy(out, gz.data(), gz.size());
    return ACVMB_OK;
}

extern "C" int acvmb_witness_map_decompress(const uint8_t* gz, size_t len, uint32_t* witness_idx, uint8_t* values_be32, uint32_t cap,
                                            uint32_t* n) {
    if (!gz || !n) return set_err(ACVMB_ERR_INVALID_ARG, "bad argument");
    std::vector<std::pair<uint32_t, U256>> wm;
    try {
        wm = decode_witness_map(gz, len);
    } catch (const std::exception& e) {
        return set_err(ACVMB_ERR_DECODE, e.what());
    }
    *n = (uint32_t)wm.size();
    if (!witness_idx || !values_be32) return ACVMB_OK;
    if (cap < wm.size()) return set_err(ACVMB_ERR_INVALID_ARG, "buffer too small");
    for (size_t i = 0; i < wm.size(); ++i) {
        witness_idx[i] = wm[i].first;
        hf::to_be_bytes(wm[i].second, values_be32 + i * 32);
    }
    return ACVMB_OK;
}
