// context.cu — contexts, the NCCL / single-process halo exchange, and the orchestration of one multiply:
//   pack (halo stream) -> grouped ncclSend/ncclRecv straight into the ghost segments of `gathered` (halo stream)
//   || interior row tiles (caller's stream)  ->  wait  ->  boundary row tiles + split long rows.
// Replaces execute_plan! (src/vectors.jl:394-463: host-staged Isend/Irecv with a D2H copy of x and an H2D copy of
// `gathered` per call) and the launch in Base.:*(A, x) / mul! (src/sparse.jl:2096-2128, 2019-2037).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>
#include <sched.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>

#include "device.h"

using namespace hpcla;

#define CU_TRY(expr)                                                                                         \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) return fail(HPCLA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// NCCL, resolved at run time: the process that loaded torch already has libnccl.so.2 mapped and dlopen returns that
// copy; a Julia process gets NCCL_jll's.  Nothing here needs NCCL until a multi-rank context asks for it.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;  // optional
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
#define LOAD(field, sym)                                          \
    api.field = (decltype(api.field))dlsym(api.handle, sym);      \
    if (!api.field) api.error = std::string("libnccl lacks ") + sym;
        LOAD(GetUniqueId, "ncclGetUniqueId")
        LOAD(CommInitRank, "ncclCommInitRank")
        api.CommInitRankConfig = (decltype(api.CommInitRankConfig))dlsym(api.handle, "ncclCommInitRankConfig");
        LOAD(CommDestroy, "ncclCommDestroy")
        LOAD(Send, "ncclSend")
        LOAD(Recv, "ncclRecv")
        LOAD(GroupStart, "ncclGroupStart")
        LOAD(GroupEnd, "ncclGroupEnd")
        LOAD(AllReduce, "ncclAllReduce")
        LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    });
    return &api;
}
}  // namespace

#define NCCL_TRY(expr)                                                                                             \
    do {                                                                                                           \
        ncclResult_t _r = (expr);                                                                                  \
        if (_r != ncclSuccess) return fail(HPCLA_ERR_NCCL, "%s failed: %s", #expr, nccl_api()->GetErrorString(_r)); \
    } while (0)

// Stream memory operations of the driver API (the runtime has no equivalent): a stream waits until a 32-bit word in
// device memory reaches a value — no kernel spins, no SM is held.  Resolved through the runtime (cudaGetDriverEntryPoint):
// the library does not link libcuda.
namespace {
typedef CUresult (*fn_stream_wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*fn_stream_write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
fn_stream_write32 stream_write32() {
    static fn_stream_write32 fn = []() -> fn_stream_write32 {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (fn_stream_write32)p;
    }();
    return fn;
}
fn_stream_wait32 stream_wait32() {
    static fn_stream_wait32 fn = []() -> fn_stream_wait32 {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (fn_stream_wait32)p;
    }();
    return fn;
}
}  // namespace

// Inside ncclGroupStart .. ncclGroupEnd: a failed call closes the group before the error is returned, so that later
// NCCL calls of this thread are not queued behind a group that never ends.
#define NCCL_TRY_IN_GROUP(expr)                                                                                    \
    do {                                                                                                           \
        ncclResult_t _r = (expr);                                                                                  \
        if (_r != ncclSuccess) {                                                                                   \
            nccl_api()->GroupEnd();                                                                                \
            return fail(HPCLA_ERR_NCCL, "%s failed: %s", #expr, nccl_api()->GetErrorString(_r));                   \
        }                                                                                                          \
    } while (0)

// NVTX ranges around the entry points that enqueue work: no-ops unless a tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ---------------------------------------------------------------------------------------------------------------
struct hpcla_group {
    std::mutex mu;  // rank-threads register and look up operators concurrently
    std::vector<hpcla_ctx*> ctxs;
    std::vector<std::vector<hpcla_spmv*>> ops;  // [sequence][rank]
    int refs = 0;
};

struct hpcla_ctx {
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t halo_stream = nullptr;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;  // staged multiply (created on first use)
    cudaStream_t cap_stream = nullptr;                        // graph capture (created on first use)
    ncclComm_t comm = nullptr;
    bool comm_owned = false;
    hpcla_group* group = nullptr;
    i64 op_seq = 0;
    double* d_red_scratch = nullptr;  // reduce_scratch_doubles()
    double* d_red_out = nullptr;      // 8 doubles
    double* h_red_out = nullptr;      // pinned, 8 doubles
    size_t l2_persist_max = 0, l2_window_max = 0, l2_persist_set = 0;  // persisting-L2 limits of the device / what has been set aside
};

struct hpcla_csr {
    hpcla_ctx* ctx = nullptr;
    int dtype = 0, itype = 0;
    i64 nrows = 0, ncc = 0, nnz = 0;
    const void *d_rowptr = nullptr, *d_colval = nullptr, *d_nzval = nullptr;  // borrowed
    TileShape shape{};
    TileDesc* d_tiles = nullptr;
    i64 ntiles = 0;
    std::vector<unsigned char> tile_class;  // per tile: 0 no rows, 1 row-walk kernel, 2 general kernel
    std::vector<TileDesc> h_tiles;          // host copy of the tile table [ntiles+1]
    unsigned char* d_hdrs = nullptr;        // tile headers of the direct row walk (ntiles * shape.hdr_bytes), or null
    i64 n_class[3] = {0, 0, 0};
    i64 long_threshold = 0, chunk_nnz = 0;
    i64 nlong = 0, nchunks = 0;
    i64 *d_long_rows = nullptr, *d_chunk_ptr = nullptr;
    void* d_partials = nullptr;
    FlatData flat;        // irregular matrices: the nnz-split multiply (flat.cu); n_chunks == 0 otherwise
    bool flat_keep_x = false;  // gather x with an L2 evict-last hint (HPCLA_FLAT_KEEP_X)
    bool flat_l2_window = false;  // pin the own segment of x in the persisting part of L2 during the launch (HPCLA_FLAT_L2_WINDOW=1; measured slower)
};

struct Seg {
    int peer;
    i64 start;   // recv: 1-based first position in gathered; send: offset (elements) into the packed send buffer
    i64 count;
    bool contiguous;  // send only: the requested local indices are src0, src0+1, ...
    i64 src0;         // send only: first 1-based local index
};

// Staged multiply (host x in, host y out): the rows are cut into blocks of consecutive tiles; block k may run once the
// prefix of x.v its interior tiles read has landed, and its slice of y goes back while later blocks compute.
struct HostPipe {
    bool usable = false;
    int nb = 0;
    std::vector<i64> row_at;     // [nb+1] first local row of each block
    std::vector<int> pos[3][2];  // [nb+1] positions of the block boundaries in each tile list
    std::vector<int> in_chunk;   // [nb] the x chunk that must have landed before the block runs (-1: none)
    std::vector<char> late;      // [nb] y slice complete only after the boundary tiles / split long rows
    std::vector<i64> xchunk;     // [nb+1] element offsets of the x chunks
    std::vector<cudaEvent_t> ev_in, ev_c;
    cudaEvent_t ev_start = nullptr, ev_tail = nullptr, ev_done = nullptr;
};

struct hpcla_spmv {
    hpcla_ctx* ctx = nullptr;
    hpcla_csr* csr = nullptr;
    hpcla_plan plan;  // private copy of the index fields
    i64 n_x_local = 0;
    i64 seq = 0;
    // own segment of gathered
    i64 own_lo = 1, own_n = 0, own_src0 = 1;
    bool x_in_place = true;  // own columns read straight from x.v
    bool has_ghost = false;
    bool has_peers = false;
    std::vector<Seg> sends, recvs;
    bool sends_contiguous = true;
    i64 total_send = 0;
    void* d_gathered = nullptr;
    void* d_sendbuf = nullptr;
    i64* d_send_idx = nullptr;                              // concatenated send_indices (all peers)
    i64 *d_local_src = nullptr, *d_local_dst = nullptr;     // only when needed (fallback / gather hook)
    // tile lists [class][0 interior, 1 boundary]; class 0: row walk, 1: general kernel, 2: row walk on compact tiles
    // (interior only: the tiles of class 0 for which compact data could be built)
    TileRec* d_list[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    int n_list[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    std::vector<int> h_list[3][2];  // host copies (block boundaries of the staged multiply)
    // every list as runs of consecutive tiles, found once: {first position in the list, first tile}, ascending
    std::vector<std::pair<int, int>> runs[3][2];
    // Direct halo (hpcla_spmv_halo_*): ghosts are pushed straight into the peers' `gathered` over NVLink by the copy
    // engine (peer memory mapped with CUDA IPC, or plain pointers inside one process) and announced with flags the
    // receiving stream waits on; no NCCL kernel, no rendezvous.  flags: [p] = last step whose data from rank p has landed
    // here; [nranks + p] = last step whose data rank p has finished reading (so its segment may be overwritten).
    struct DirectHalo {
        bool on = false;
        unsigned* d_flags = nullptr;
        std::vector<char*> peer_gathered;    // [nranks] base of rank p's `gathered` as seen from here (null: no exchange with p)
        std::vector<unsigned*> peer_flags;   // [nranks]
        std::vector<i64> peer_recv_start;    // [nranks] 1-based start, in rank p's gathered, of the segment I fill
        std::vector<void*> ipc_opened;       // mappings to close
        unsigned step = 0;
        bool memop_failed = false;           // cuStreamWriteValue32 refused the address once: flags are written by a kernel
        bool consumed_signalled = false;     // this step's "ghosts read" flags went out on the halo stream already
    } direct;
    // compact tiles (compact.cu): headers and 16-bit positions, by position in list [2][0]
    CompactShape csh{};
    unsigned char *d_chdr = nullptr, *d_cpos = nullptr;
    int ctail_q_min = 0x7fffffff;
    struct HostPipe* pipe = nullptr;
    // fused dot(x, A x) for CG: one partial per row-walk CTA, interior list first, then boundary list
    double* d_dot_partials = nullptr;
    bool dot_request = false;
    cudaEvent_t ev_x = nullptr, ev_packed = nullptr, ev_halo = nullptr;
    bool halo_recorded = false;
    // in-flight call
    const void* cur_x = nullptr;
    void* cur_y = nullptr;
    cudaStream_t cur_stream = nullptr;
    int phase = 0;  // 0 idle, 1 multiply begun, 2 gather begun
    // sparse x dense: compact row-major ghost rows / packed send rows for up to mm_cols columns
    void *d_ghost_rm = nullptr, *d_sendbuf_rm = nullptr;
    int mm_cols = 0;
    const void* mm_B = nullptr;
    void* mm_C = nullptr;
    i64 mm_ldb = 0, mm_ldc = 0;
    int mm_ncols = 0;
    std::atomic<long long> epoch{0};  // exchanges begun (a peer's finish checks that I have begun the matching one)
    i64 launches = 0;
    double* d_cg_scalars = nullptr;  // hpcla_cg: rr history + p.q (device)
    int cg_cap = 0;
    std::vector<double> cg_host;
    cudaEvent_t ev_last = nullptr;  // end of the most recent call on the caller's stream (what destroy waits for)
    // HPCLA_TIMELINE=1: timing events of the most recent multiply (hpcla_spmv_timeline)
    bool timeline = false;
    cudaEvent_t tl[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // x ready, exchange done, boundary done, interior done, end
    bool tl_rec[5] = {false, false, false, false, false};
    // CUDA graph of one multiply bound to fixed x / y (hpcla_spmv_graph_*)
    cudaGraphExec_t cg_graph = nullptr;  // HPCLA_CG_GRAPH=1: the whole hpcla_cg loop, keyed by its buffers and length
    const void *cg_key_b = nullptr, *cg_key_x = nullptr, *cg_key_w = nullptr;
    int cg_key_iters = -1;
    bool cg_key_fused = false;
    i64 cg_graph_launches = 0;
    cudaGraphExec_t graph = nullptr;
    const void* graph_x = nullptr;
    void* graph_y = nullptr;
    i64 graph_launches = 0;
};

// Handles under construction: destroyed on every early return, released on success.
template <class H, void (*Destroy)(H*)>
struct Building {
    H* h;
    explicit Building(H* p) : h(p) {}
    ~Building() {
        if (h) Destroy(h);
    }
    H* release() {
        H* p = h;
        h = nullptr;
        return p;
    }
};

static int set_device(const hpcla_ctx* ctx) {
    CU_TRY(cudaSetDevice(ctx->device));
    return HPCLA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------
extern "C" int hpcla_ctx_create(int device, int rank, int nranks, hpcla_ctx** out) {
    NvtxRange nvtx_range("hpcla_ctx_create");
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return fail(HPCLA_ERR_ARG, "hpcla_ctx_create: bad arguments");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(HPCLA_ERR_CUDA, "hpcla_ctx_create: no CUDA device is visible (%s); this backend has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(HPCLA_ERR_ARG, "hpcla_ctx_create: device %d out of range (%d visible)", device, ndev);
    CU_TRY(cudaSetDevice(device));
    hpcla_ctx* c = new hpcla_ctx();
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    // Highest priority: the small halo kernels (pack, NCCL send/recv) must get SMs ahead of the tens of thousands of
    // pending CTAs of the interior multiply, or the exchange only starts when the multiply has drained.
    int prio_least = 0, prio_greatest = 0;
    CU_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    CU_TRY(cudaStreamCreateWithPriority(&c->halo_stream, cudaStreamNonBlocking, prio_greatest));
    CU_TRY(cudaMalloc(&c->d_red_scratch, sizeof(double) * reduce_scratch_doubles()));
    CU_TRY(cudaMemset(c->d_red_scratch, 0, sizeof(double) * reduce_scratch_doubles()));
    CU_TRY(cudaMalloc(&c->d_red_out, sizeof(double) * 8));
    CU_TRY(cudaMallocHost(&c->h_red_out, sizeof(double) * 8));
    {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, device) == cudaSuccess) c->l2_persist_max = (size_t)v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, device) == cudaSuccess) c->l2_window_max = (size_t)v;
        cudaGetLastError();
    }
    *out = c;
    return HPCLA_OK;
}

extern "C" int hpcla_nccl_unique_id(void* id128) {
    NcclApi* api = nccl_api();
    if (!api->error.empty()) return fail(HPCLA_ERR_NCCL, "%s", api->error.c_str());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(api->GetUniqueId(&id));
    std::memcpy(id128, &id, 128);
    return HPCLA_OK;
}

extern "C" int hpcla_ctx_init_nccl(hpcla_ctx* ctx, const void* id128) {
    NvtxRange nvtx_range("hpcla_ctx_init_nccl");
    if (!ctx || !id128) return fail(HPCLA_ERR_ARG, "hpcla_ctx_init_nccl: null");
    if (ctx->comm || ctx->group) return fail(HPCLA_ERR_STATE, "hpcla_ctx_init_nccl: the context already has a world");
    NcclApi* api = nccl_api();
    if (!api->error.empty()) return fail(HPCLA_ERR_NCCL, "%s", api->error.c_str());
    int rc = set_device(ctx);
    if (rc) return rc;
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    // The halo messages are a few MB at most: a communicator capped at a few CTAs leaves the SMs to the multiply
    // (HPCLA_NCCL_MAX_CTAS, tuning hook; unset = NCCL's default).
    // (measured: a cap of 2 gains 1.4 % at 2 GPUs, profiles/r2m2_*, but at 8 GPUs the capped communicator was the slowest
    // configuration, profiles/r2m8_*: left to NCCL unless asked for)
    int max_ctas = 0;
    if (const char* e = getenv("HPCLA_NCCL_MAX_CTAS")) max_ctas = atoi(e);
    if (max_ctas > 0 && api->CommInitRankConfig) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = max_ctas;
        NCCL_TRY(api->CommInitRankConfig(&ctx->comm, ctx->nranks, id, ctx->rank, &cfg));
    } else {
        NCCL_TRY(api->CommInitRank(&ctx->comm, ctx->nranks, id, ctx->rank));
    }
    ctx->comm_owned = true;
    return HPCLA_OK;
}

extern "C" int hpcla_ctx_adopt_nccl(hpcla_ctx* ctx, void* nccl_comm) {
    if (!ctx || !nccl_comm) return fail(HPCLA_ERR_ARG, "hpcla_ctx_adopt_nccl: null");
    if (ctx->comm || ctx->group) return fail(HPCLA_ERR_STATE, "hpcla_ctx_adopt_nccl: the context already has a world");
    NcclApi* api = nccl_api();
    if (!api->error.empty()) return fail(HPCLA_ERR_NCCL, "%s", api->error.c_str());
    ctx->comm = (ncclComm_t)nccl_comm;
    ctx->comm_owned = false;
    return HPCLA_OK;
}

extern "C" int hpcla_ctx_form_group(hpcla_ctx* const* ctxs, int n) {
    if (!ctxs || n < 1) return fail(HPCLA_ERR_ARG, "hpcla_ctx_form_group: bad arguments");
    for (int r = 0; r < n; ++r) {
        if (!ctxs[r] || ctxs[r]->rank != r || ctxs[r]->nranks != n) return fail(HPCLA_ERR_ARG, "hpcla_ctx_form_group: contexts must be ranks 0..n-1 of an n-rank world, in order");
        if (ctxs[r]->comm || ctxs[r]->group) return fail(HPCLA_ERR_STATE, "hpcla_ctx_form_group: context %d already has a world", r);
    }
    for (int a = 0; a < n; ++a)  // peer access for cross-device copies (same device: nothing to do)
        for (int b = 0; b < n; ++b)
            if (ctxs[a]->device != ctxs[b]->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, ctxs[a]->device, ctxs[b]->device);
                if (can) {
                    cudaSetDevice(ctxs[a]->device);
                    cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[b]->device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(HPCLA_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
                    cudaGetLastError();
                }
            }
    hpcla_group* g = new hpcla_group();
    g->ctxs.assign(ctxs, ctxs + n);
    g->refs = n;
    for (int r = 0; r < n; ++r) ctxs[r]->group = g;
    return HPCLA_OK;
}

// Pinned host memory placed next to the GPU: the calling thread is moved onto the CPUs the GPU's PCIe root is attached
// to (sysfs local_cpulist) while cudaHostAlloc allocates and touches the pages, so that on a multi-socket host the
// staged multiply's H2D / D2H copies do not cross the socket interconnect.  Falls back to a plain pinned allocation
// when the topology cannot be read.
static bool parse_cpulist(const char* s, cpu_set_t* set) {
    CPU_ZERO(set);
    int n = 0;
    while (*s) {
        char* end = nullptr;
        long a = strtol(s, &end, 10);
        if (end == s) break;
        long b = a;
        s = end;
        if (*s == '-') {
            b = strtol(s + 1, &end, 10);
            s = end;
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET((int)c, set), ++n;
        if (*s == ',') ++s;
        else break;
    }
    return n > 0;
}

extern "C" int hpcla_host_alloc(hpcla_ctx* ctx, int64_t bytes, void** out, int* numa_node_out) {
    if (!ctx || !out || bytes < 0) return fail(HPCLA_ERR_ARG, "hpcla_host_alloc: bad arguments");
    int rc = set_device(ctx);
    if (rc) return rc;
    if (numa_node_out) *numa_node_out = -1;
    char busid[64] = {0};
    cpu_set_t old_set, gpu_set;
    bool moved = false;
    if (cudaDeviceGetPCIBusId(busid, sizeof busid, ctx->device) == cudaSuccess && sched_getaffinity(0, sizeof old_set, &old_set) == 0) {
        for (char* c = busid; *c; ++c) *c = (char)tolower(*c);
        char path[160], buf[4096];
        snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", busid);
        if (FILE* f = fopen(path, "r")) {
            if (fgets(buf, sizeof buf, f) && parse_cpulist(buf, &gpu_set)) {
                CPU_AND(&gpu_set, &gpu_set, &old_set);  // never leave the CPUs this process is allowed on
                if (CPU_COUNT(&gpu_set) > 0 && sched_setaffinity(0, sizeof gpu_set, &gpu_set) == 0) moved = true;
            }
            fclose(f);
        }
        snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", busid);
        if (FILE* f = fopen(path, "r")) {
            int node = -1;
            if (fscanf(f, "%d", &node) == 1 && numa_node_out) *numa_node_out = node;
            fclose(f);
        }
    }
    cudaError_t e = cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 16), cudaHostAllocDefault);
    if (e == cudaSuccess && bytes > 0) std::memset(*out, 0, (size_t)bytes);  // first touch from the GPU's own CPUs
    if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
    if (e != cudaSuccess) return fail(HPCLA_ERR_CUDA, "hpcla_host_alloc: cudaHostAlloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    return HPCLA_OK;
}

extern "C" int hpcla_host_free(hpcla_ctx* ctx, void* p) {
    if (!ctx) return fail(HPCLA_ERR_ARG, "hpcla_host_free: null");
    if (!p) return HPCLA_OK;
    int rc = set_device(ctx);
    if (rc) return rc;
    CU_TRY(cudaFreeHost(p));
    return HPCLA_OK;
}

extern "C" int hpcla_ctx_sync(hpcla_ctx* ctx) {
    if (!ctx) return fail(HPCLA_ERR_ARG, "hpcla_ctx_sync: null");
    int rc = set_device(ctx);
    if (rc) return rc;
    CU_TRY(cudaDeviceSynchronize());
    return HPCLA_OK;
}

extern "C" void hpcla_ctx_destroy(hpcla_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->comm && ctx->comm_owned) nccl_api()->CommDestroy(ctx->comm);
    if (ctx->group && --ctx->group->refs == 0) delete ctx->group;
    if (ctx->halo_stream) cudaStreamDestroy(ctx->halo_stream);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
    cudaFree(ctx->d_red_scratch);
    cudaFree(ctx->d_red_out);
    cudaFreeHost(ctx->h_red_out);
    delete ctx;
}

namespace hpcla {
int ctx_rank_info(const hpcla_ctx* ctx, int* device, int* rank, int* nranks, int* has_comm) {
    if (!ctx) return fail(HPCLA_ERR_ARG, "ctx_rank_info: null");
    *device = ctx->device;
    *rank = ctx->rank;
    *nranks = ctx->nranks;
    *has_comm = ctx->comm ? 1 : 0;
    return HPCLA_OK;
}

// max over the ranks of a small non-negative status (NCCL world; the value itself on a single rank): how a rank-local
// failure inside a collective set-up routine is made known to everybody before the next exchange is entered
int ctx_agree_max(hpcla_ctx* ctx, int local, int* out, cudaStream_t stream) {
    if (!ctx || !out) return fail(HPCLA_ERR_ARG, "ctx_agree_max: null");
    *out = local;
    if (ctx->nranks == 1 || !ctx->comm) return HPCLA_OK;
    int* d = reinterpret_cast<int*>(ctx->d_red_out + 4);  // (second half of the 8-double result slot)
    int* h = reinterpret_cast<int*>(ctx->h_red_out + 4);
    *h = local;
    CU_TRY(cudaMemcpyAsync(d, h, sizeof(int), cudaMemcpyHostToDevice, stream));
    NCCL_TRY(nccl_api()->AllReduce(d, d, 1, ncclInt32, ncclMax, ctx->comm, stream));
    CU_TRY(cudaMemcpyAsync(h, d, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaStreamSynchronize(stream));
    *out = *h;
    return HPCLA_OK;
}

int ctx_exchange_bytes(hpcla_ctx* ctx, const void* d_send, const i64* send_off, const i64* send_bytes, void* d_recv, const i64* recv_off,
                       const i64* recv_bytes, cudaStream_t stream) {
    if (!ctx) return fail(HPCLA_ERR_ARG, "ctx_exchange_bytes: null");
    const int me = ctx->rank;
    if (send_bytes[me] != recv_bytes[me]) return fail(HPCLA_ERR_STATE, "ctx_exchange_bytes: own range sizes differ");
    if (send_bytes[me] > 0)
        CU_TRY(cudaMemcpyAsync((char*)d_recv + recv_off[me], (const char*)d_send + send_off[me], (size_t)send_bytes[me], cudaMemcpyDeviceToDevice, stream));
    if (ctx->nranks == 1) return HPCLA_OK;
    if (!ctx->comm) return fail(HPCLA_ERR_STATE, "ctx_exchange_bytes: no NCCL communicator");
    NcclApi* api = nccl_api();
    NCCL_TRY(api->GroupStart());
    for (int q = 0; q < ctx->nranks; ++q) {
        if (q == me) continue;
        if (send_bytes[q] > 0) NCCL_TRY_IN_GROUP(api->Send((const char*)d_send + send_off[q], (size_t)send_bytes[q], ncclChar, q, ctx->comm, stream));
        if (recv_bytes[q] > 0) NCCL_TRY_IN_GROUP(api->Recv((char*)d_recv + recv_off[q], (size_t)recv_bytes[q], ncclChar, q, ctx->comm, stream));
    }
    NCCL_TRY(api->GroupEnd());
    return HPCLA_OK;
}
}  // namespace hpcla

// ---------------------------------------------------------------------------------------------------------------
// CSR view
// ---------------------------------------------------------------------------------------------------------------
extern "C" int hpcla_csr_create(hpcla_ctx* ctx, int dtype, int itype, int64_t nrows, int64_t ncc, int64_t nnz, const void* d_rowptr,
                                const void* d_colval, const void* d_nzval, hpcla_csr** out) {
    NvtxRange nvtx_range("hpcla_csr_create");
    if (!ctx || !out || nrows < 0 || ncc < 0 || nnz < 0 || !d_rowptr) return fail(HPCLA_ERR_ARG, "hpcla_csr_create: bad arguments");
    if (!dtype_size(dtype) || !itype_size(itype)) return fail(HPCLA_ERR_ARG, "hpcla_csr_create: unknown dtype/itype");
    if (nnz > 0 && (!d_colval || !d_nzval)) return fail(HPCLA_ERR_ARG, "hpcla_csr_create: null colval/nzval");
    if (((uintptr_t)d_colval & 15) || ((uintptr_t)d_nzval & 15) || ((uintptr_t)d_rowptr & 15))
        return fail(HPCLA_ERR_ARG, "hpcla_csr_create: rowptr, colval and nzval must be 16-byte aligned (128-bit loads / bulk copies)");
    if (itype == HPCLA_I32 && nnz >= (i64)INT32_MAX) return fail(HPCLA_ERR_ARG, "hpcla_csr_create: nnz does not fit Int32 row pointers");
    int rc = set_device(ctx);
    if (rc) return rc;
    hpcla_csr* A = new hpcla_csr();
    Building<hpcla_csr, hpcla_csr_destroy> guard(A);
    A->ctx = ctx;
    A->dtype = dtype;
    A->itype = itype;
    A->nrows = nrows;
    A->ncc = ncc;
    A->nnz = nnz;
    A->d_rowptr = d_rowptr;
    A->d_colval = d_colval;
    A->d_nzval = d_nzval;
    cudaStream_t st = ctx->halo_stream;
    // The arrays are borrowed as they are: check once that they describe a CSR matrix, so that a wrong row pointer or
    // column fails here and not as an out-of-range bulk copy or gather on the device (HPCLA_VALIDATE=0 skips the pass).
    {
        const char* e = getenv("HPCLA_VALIDATE");
        if (!(e && e[0] == '0')) {
            unsigned* d_bad = nullptr;
            unsigned bad = 0;
            CU_TRY(cudaMalloc(&d_bad, sizeof(unsigned)));
            CU_TRY(cudaMemsetAsync(d_bad, 0, sizeof(unsigned), st));
            CU_TRY(launch_validate_csr(itype, d_rowptr, d_colval, nrows, nnz, ncc, d_bad, st));
            CU_TRY(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            cudaFree(d_bad);
            if (bad) return fail(HPCLA_ERR_ARG, "hpcla_csr_create: the arrays are not a valid 1-based CSR view (%u violations of: rowptr[1] == 1, rowptr non-decreasing, rowptr[end] == nnz + 1, 1 <= colval <= ncols_compressed)", bad);
        }
    }
    // Tile shape from the mean row length, then the kernel class of every tile (row walk for balanced, fully staged
    // tiles; general otherwise).  A matrix whose tiles are mostly general is re-tiled with the general kernel's own
    // shape.  Tuning hooks: HPCLA_LANES, HPCLA_TILE_WINDOW, HPCLA_SPMV_KIND=general|rowwalk.
    int lanes_override = 0, window_override = 0, kind = 0;
    if (const char* e = getenv("HPCLA_LANES")) lanes_override = atoi(e);
    if (const char* e = getenv("HPCLA_TILE_WINDOW")) window_override = atoi(e);
    if (const char* e = getenv("HPCLA_SPMV_KIND")) kind = (e[0] == 'g') ? 2 : (e[0] == 'r') ? 1 : (e[0] == 'f') ? 3 : 0;
    int balance_pct = 50;  // a tile goes to the row walk when its mean row length is at least this share of its longest row
    if (const char* e = getenv("HPCLA_BALANCE_PCT")) balance_pct = std::max(0, std::min(100, atoi(e)));
    // typical row length: the most common one (a window of whole typical rows keeps tiles row-aligned), else the mean
    double avg_row = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
    if (nrows > 0 && nnz > 0) {
        unsigned long long* d_hist = nullptr;
        std::vector<unsigned long long> hist(1024, 0);
        CU_TRY(cudaMalloc(&d_hist, sizeof(unsigned long long) * 1024));
        CU_TRY(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * 1024, st));
        CU_TRY(launch_row_len_hist(itype, d_rowptr, nrows, d_hist, st));
        CU_TRY(cudaMemcpyAsync(hist.data(), d_hist, sizeof(unsigned long long) * 1024, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        cudaFree(d_hist);
        int mode = 1;
        for (int l = 1; l < 1023; ++l)
            if (hist[(size_t)l] > hist[(size_t)mode]) mode = l;
        if (2 * hist[(size_t)mode] >= (unsigned long long)nrows) avg_row = (double)mode;  // a clear majority of the rows
    }
    A->long_threshold = 16384;
    A->chunk_nnz = 16384;
    for (int pass = 0; pass < 2; ++pass) {
        const bool irregular = (pass == 1) || kind == 2 || kind == 3;
        A->shape = tile_shape(dtype, itype, avg_row, irregular, lanes_override, window_override);
        A->ntiles = nnz / A->shape.window + 1;
        if (A->ntiles >= (i64)INT32_MAX) return fail(HPCLA_ERR_ARG, "hpcla_csr_create: too many tiles");
        if (A->d_tiles) cudaFree(A->d_tiles);
        A->d_tiles = nullptr;
        CU_TRY(cudaMalloc(&A->d_tiles, sizeof(TileDesc) * (size_t)(A->ntiles + 1)));
        CU_TRY(launch_build_tiles(itype, d_rowptr, nrows, nnz, A->shape.window, A->d_tiles, A->ntiles, st));
        unsigned char* d_cls = nullptr;
        CU_TRY(cudaMalloc(&d_cls, (size_t)A->ntiles));
        CU_TRY(launch_tile_class(itype, d_rowptr, A->d_tiles, A->ntiles, A->shape.window, A->shape.cap, A->shape.rp_cap, balance_pct, d_cls, st));
        A->tile_class.assign((size_t)A->ntiles, 0);
        A->h_tiles.resize((size_t)A->ntiles + 1);
        CU_TRY(cudaMemcpyAsync(A->tile_class.data(), d_cls, (size_t)A->ntiles, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(A->h_tiles.data(), A->d_tiles, sizeof(TileDesc) * ((size_t)A->ntiles + 1), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        cudaFree(d_cls);
        A->n_class[0] = A->n_class[1] = A->n_class[2] = 0;
        for (unsigned char c : A->tile_class) A->n_class[c] += 1;
        if (irregular || kind == 1 || A->n_class[1] >= A->n_class[2]) break;
    }
    // Tile headers of the direct row walk.  Measured (profiles/r1h_tune_direct_walk.txt): it wins where its 16-bit row
    // offsets replace 8-byte row pointers (Int64 indices: 338 -> 327 us on Poisson 256^3) and for real-valued multi-lane
    // walks (27-point Float64: 405 -> 371 us); it loses where registers are the scarce resource (one lane per row at 8
    // CTAs per SM: 264 -> 267..282 us; ComplexF64: 640 -> 662 us).  HPCLA_DIRECT=0|1 overrides.
    {
        const char* e = getenv("HPCLA_DIRECT");
        bool direct = itype == HPCLA_I64 || (dtype != HPCLA_C128 && A->shape.lanes >= 2);
        if (e) direct = e[0] != '0';
        if (direct && A->shape.hdr_rows > 0 && A->n_class[1] > 0) {
            unsigned char* d_cls = nullptr;
            CU_TRY(cudaMalloc(&d_cls, (size_t)A->ntiles));
            CU_TRY(cudaMemcpyAsync(d_cls, A->tile_class.data(), (size_t)A->ntiles, cudaMemcpyHostToDevice, st));
            CU_TRY(cudaMalloc(&A->d_hdrs, (size_t)A->ntiles * (size_t)A->shape.hdr_bytes));
            CU_TRY(cudaMemsetAsync(A->d_hdrs, 0, (size_t)A->ntiles * (size_t)A->shape.hdr_bytes, st));
            CU_TRY(launch_build_tile_headers(itype, d_rowptr, A->d_tiles, d_cls, A->ntiles, A->shape.window, A->shape.hdr_bytes, A->d_hdrs, st));
            CU_TRY(cudaStreamSynchronize(st));
            cudaFree(d_cls);
        }
    }
    // Irregular matrices (most entries in tiles the row walk cannot take) multiply with the nnz-split kernel of flat.cu;
    // the tile table above still serves the sparse x dense product.  HPCLA_SPMV_KIND=general keeps the tile kernel.
    if (A->shape.lanes == 0 && kind != 2 && nnz > 0) {
        CU_TRY(flat_build(itype, d_rowptr, nrows, nnz, &A->flat, st));
        CU_TRY(cudaMalloc(&A->flat.d_heads, dtype_size(dtype) * (size_t)std::max<i64>(A->flat.n_wchunks, 1)));
        if (const char* e = getenv("HPCLA_FLAT_KEEP_X")) A->flat_keep_x = e[0] == '1';
        if (const char* e = getenv("HPCLA_FLAT_L2_WINDOW")) A->flat_l2_window = e[0] == '1';
    }
    // rows longer than the split threshold (rare: power-law tails)
    const i64 cap = nnz / A->long_threshold + 1;
    unsigned long long* d_count = nullptr;
    i64* d_rows = nullptr;
    CU_TRY(cudaMalloc(&d_count, sizeof(unsigned long long)));
    CU_TRY(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
    CU_TRY(cudaMalloc(&d_rows, sizeof(i64) * (size_t)cap));
    CU_TRY(launch_find_long_rows(itype, d_rowptr, nrows, A->long_threshold, d_rows, cap, d_count, st));
    unsigned long long count = 0;
    CU_TRY(cudaMemcpyAsync(&count, d_count, sizeof count, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    cudaFree(d_count);
    A->nlong = (i64)count;
    if (A->nlong > 0) {
        std::vector<i64> rows((size_t)A->nlong);
        CU_TRY(cudaMemcpy(rows.data(), d_rows, sizeof(i64) * rows.size(), cudaMemcpyDeviceToHost));
        std::sort(rows.begin(), rows.end());
        // row lengths: two row-pointer reads per long row
        std::vector<i64> chunk_ptr((size_t)A->nlong + 1, 0);
        const size_t is = itype_size(itype);
        for (i64 i = 0; i < A->nlong; ++i) {
            char two[16];
            CU_TRY(cudaMemcpy(two, (const char*)d_rowptr + (size_t)rows[(size_t)i] * is, 2 * is, cudaMemcpyDeviceToHost));
            i64 b = is == 4 ? (i64)((int32_t*)two)[0] : ((i64*)two)[0];
            i64 e = is == 4 ? (i64)((int32_t*)two)[1] : ((i64*)two)[1];
            chunk_ptr[(size_t)i + 1] = chunk_ptr[(size_t)i] + (e - b + A->chunk_nnz - 1) / A->chunk_nnz;
        }
        A->nchunks = chunk_ptr.back();
        CU_TRY(cudaMalloc(&A->d_long_rows, sizeof(i64) * rows.size()));
        CU_TRY(cudaMalloc(&A->d_chunk_ptr, sizeof(i64) * chunk_ptr.size()));
        CU_TRY(cudaMalloc(&A->d_partials, dtype_size(dtype) * (size_t)A->nchunks));
        CU_TRY(cudaMemcpy(A->d_long_rows, rows.data(), sizeof(i64) * rows.size(), cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(A->d_chunk_ptr, chunk_ptr.data(), sizeof(i64) * chunk_ptr.size(), cudaMemcpyHostToDevice));
    }
    cudaFree(d_rows);
    *out = guard.release();
    return HPCLA_OK;
}

extern "C" int hpcla_csr_info(const hpcla_csr* A, int64_t* ntiles_out, int64_t* nlong_out, int* variant_out) {
    if (!A) return fail(HPCLA_ERR_ARG, "hpcla_csr_info: null");
    if (ntiles_out) *ntiles_out = A->ntiles;
    if (nlong_out) *nlong_out = A->nlong;
    if (variant_out) *variant_out = A->shape.lanes;
    return HPCLA_OK;
}

extern "C" int hpcla_csr_tile_classes(const hpcla_csr* A, int64_t* n_rowwalk, int64_t* n_general, int64_t* n_empty, int* window_out) {
    if (!A) return fail(HPCLA_ERR_ARG, "hpcla_csr_tile_classes: null");
    if (n_rowwalk) *n_rowwalk = A->n_class[1];
    if (n_general) *n_general = A->n_class[2];
    if (n_empty) *n_empty = A->n_class[0];
    if (window_out) *window_out = A->shape.window;
    return HPCLA_OK;
}

extern "C" void hpcla_csr_destroy(hpcla_csr* A) {
    if (!A) return;
    cudaSetDevice(A->ctx->device);
    cudaFree(A->d_tiles);
    cudaFree(A->d_hdrs);
    cudaFree(A->d_long_rows);
    cudaFree(A->d_chunk_ptr);
    cudaFree(A->d_partials);
    flat_free(&A->flat);
    delete A;
}

// ---------------------------------------------------------------------------------------------------------------
// bound operator
// ---------------------------------------------------------------------------------------------------------------
static bool is_consecutive(const std::vector<i64>& v) {
    for (size_t k = 1; k < v.size(); ++k)
        if (v[k] != v[k - 1] + 1) return false;
    return true;
}

// Compact data for the interior row-walk tiles (compact.cu): analysis pass (x runs and staged x elements per tile), the
// staged-x capacity from it, then headers + 16-bit positions for the tiles that fit.  HPCLA_COMPACT=0 disables.
static int build_compact(hpcla_spmv* op, std::vector<int> (&lists)[3][2]) {
    const hpcla_csr* A = op->csr;
    const char* env = getenv("HPCLA_COMPACT");
    if (env && env[0] == '0') return HPCLA_OK;
    if (!(env && env[0] == '1')) {
        // Default: where it pays.  The compact walk trades the column indices for 16-bit positions and a longer start-up
        // chain per tile (header -> x runs); it wins when the multiply is bound by DRAM traffic and the indices are a fair
        // share of it.  Matrices that fit L2 (latency-bound: the direct walk has no dependent load at all) and element
        // types whose values dwarf the indices (ComplexF64 with Int32: 2 of 20 bytes) keep the plain / direct walk.
        const double per_nz = (double)(dtype_size(A->dtype) + itype_size(A->itype));
        const double saving = ((double)itype_size(A->itype) - 2.0) / per_nz;
        if (saving < 0.15 || (double)A->nnz * per_nz < 96.0e6) return HPCLA_OK;
    }
    if (A->shape.lanes <= 0 || !op->x_in_place || op->own_n >= (i64)INT32_MAX || lists[0][0].empty()) return HPCLA_OK;
    CompactShape sh = compact_shape(A->dtype, A->shape);
    if (sh.cw <= 0 || sh.cw > 8192) return HPCLA_OK;
    cudaStream_t st = op->ctx->halo_stream;
    const int n = (int)lists[0][0].size();
    int* d_ids = nullptr;
    int2* d_stats = nullptr;
    CU_TRY(cudaMalloc(&d_ids, sizeof(int) * (size_t)n));
    CU_TRY(cudaMalloc(&d_stats, sizeof(int2) * (size_t)n));
    struct Free2 {
        int*& a;
        int2*& b;
        ~Free2() { cudaFree(a), cudaFree(b); }
    } free2{d_ids, d_stats};
    CU_TRY(cudaMemcpyAsync(d_ids, lists[0][0].data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
    CU_TRY(launch_compact_tiles(false, A->dtype, A->itype, A->d_rowptr, A->d_colval, A->d_tiles, d_ids, n, A->shape.window, op->own_lo, op->own_n, sh, d_stats,
                                nullptr, nullptr, nullptr, st));
    std::vector<int2> stats((size_t)n);
    CU_TRY(cudaMemcpyAsync(stats.data(), d_stats, sizeof(int2) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    // staged x may take as many elements as the tile stages entries (the 7-point stencil needs ~0.72 of that, the 27-point ~0.34)
    const int limit = sh.cw;
    int xmax = 0, n_ok = 0;
    for (const int2& s2 : stats)
        if (s2.x >= 0 && s2.y <= limit) xmax = std::max(xmax, s2.y), ++n_ok;
    if (n_ok * 2 < n || xmax == 0) return HPCLA_OK;  // not a banded matrix: the plain row walk keeps its tiles
    sh.xcap = (xmax + 7) & ~7;
    std::vector<int> chosen, rest;
    for (int i = 0; i < n; ++i) (stats[(size_t)i].x >= 0 && stats[(size_t)i].y <= limit ? chosen : rest).push_back(lists[0][0][(size_t)i]);
    const int nc = (int)chosen.size();
    CU_TRY(cudaMemcpyAsync(d_ids, chosen.data(), sizeof(int) * (size_t)nc, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMalloc(&op->d_chdr, (size_t)nc * (size_t)sh.chdr_bytes));
    CU_TRY(cudaMalloc(&op->d_cpos, (size_t)nc * (size_t)sh.cp_bytes));
    CU_TRY(cudaMemsetAsync(op->d_chdr, 0, (size_t)nc * (size_t)sh.chdr_bytes, st));
    int* d_tail = reinterpret_cast<int*>(d_stats);  // (the stats are no longer needed)
    op->ctail_q_min = 0x7fffffff;
    CU_TRY(cudaMemcpyAsync(d_tail, &op->ctail_q_min, sizeof(int), cudaMemcpyHostToDevice, st));
    CU_TRY(launch_compact_tiles(true, A->dtype, A->itype, A->d_rowptr, A->d_colval, A->d_tiles, d_ids, nc, A->shape.window, op->own_lo, op->own_n, sh, nullptr,
                                op->d_chdr, op->d_cpos, d_tail, st));
    CU_TRY(cudaMemcpyAsync(&op->ctail_q_min, d_tail, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    op->csh = sh;
    lists[2][0].swap(chosen);
    lists[0][0].swap(rest);
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_create(hpcla_ctx* ctx, hpcla_csr* A, const hpcla_plan* plan, int64_t n_x_local, hpcla_spmv** out) {
    NvtxRange nvtx_range("hpcla_spmv_create");
    if (!ctx || !A || !plan || !out || n_x_local < 0) return fail(HPCLA_ERR_ARG, "hpcla_spmv_create: bad arguments");
    if (A->ctx != ctx) return fail(HPCLA_ERR_ARG, "hpcla_spmv_create: the matrix belongs to another context");
    if (plan->n_gathered != A->ncc) return fail(HPCLA_ERR_ARG, "hpcla_spmv_create: plan gathers %lld elements but A has %lld compressed columns", (long long)plan->n_gathered, (long long)A->ncc);
    if (plan->rank != ctx->rank || plan->nranks != ctx->nranks) return fail(HPCLA_ERR_ARG, "hpcla_spmv_create: plan was built for rank %d of %d", plan->rank, plan->nranks);
    if (plan->n_x_local >= 0 && plan->n_x_local != n_x_local) return fail(HPCLA_ERR_ARG, "hpcla_spmv_create: x.v has %lld elements, the plan expects %lld", (long long)n_x_local, (long long)plan->n_x_local);
    int rc = set_device(ctx);
    if (rc) return rc;
    const size_t es = dtype_size(A->dtype);
    hpcla_spmv* op = new hpcla_spmv();
    Building<hpcla_spmv, hpcla_spmv_destroy> guard(op);
    op->ctx = ctx;
    op->csr = A;
    op->plan = *plan;
    op->n_x_local = n_x_local;
    const hpcla_plan& P = op->plan;
    // validate what the kernels rely on
    for (i64 v : P.local_src)
        if (v < 1 || v > n_x_local) { return fail(HPCLA_ERR_ARG, "hpcla_spmv_create: local_src index out of range"); }
    for (auto& s : P.send_indices)
        for (i64 v : s)
            if (v < 1 || v > n_x_local) { return fail(HPCLA_ERR_ARG, "hpcla_spmv_create: send index out of range"); }
    if (!is_consecutive(P.local_dst)) { return fail(HPCLA_ERR_ARG, "hpcla_spmv_create: local_dst_indices must be a consecutive range (col_indices sorted, contiguous partition)"); }
    for (auto& s : P.recv_perm)
        if (s.empty() || !is_consecutive(s)) { return fail(HPCLA_ERR_ARG, "hpcla_spmv_create: every recv_perm must be a non-empty consecutive range"); }
    // every position of gathered must be produced exactly once
    {
        i64 covered = (i64)P.local_dst.size();
        for (auto& s : P.recv_perm) covered += (i64)s.size();
        if (covered != P.n_gathered) { return fail(HPCLA_ERR_ARG, "hpcla_spmv_create: plan covers %lld of %lld gathered positions", (long long)covered, (long long)P.n_gathered); }
    }
    op->own_n = (i64)P.local_dst.size();
    op->own_lo = op->own_n ? P.local_dst[0] : 1;
    op->own_src0 = op->own_n ? P.local_src[0] : 1;
    op->x_in_place = is_consecutive(P.local_src);
    op->has_ghost = !P.recv_perm.empty();
    op->has_peers = !P.recv_rank_ids.empty() || !P.send_rank_ids.empty();
    if (op->has_peers && !ctx->comm && !ctx->group && ctx->nranks > 1) { return fail(HPCLA_ERR_STATE, "hpcla_spmv_create: the plan exchanges data but the context has neither an NCCL communicator nor a single-process group"); }
    for (size_t i = 0; i < P.recv_rank_ids.size(); ++i) op->recvs.push_back(Seg{(int)P.recv_rank_ids[i], P.recv_perm[i][0], (i64)P.recv_perm[i].size(), true, 0});
    i64 off = 0;
    for (size_t i = 0; i < P.send_rank_ids.size(); ++i) {
        const bool contig = is_consecutive(P.send_indices[i]);
        op->sends.push_back(Seg{(int)P.send_rank_ids[i], off, (i64)P.send_indices[i].size(), contig, P.send_indices[i].empty() ? 1 : P.send_indices[i][0]});
        op->sends_contiguous = op->sends_contiguous && contig;
        off += (i64)P.send_indices[i].size();
    }
    op->total_send = off;
    if (op->has_ghost || !op->x_in_place) {
        CU_TRY(cudaMalloc(&op->d_gathered, es * (size_t)std::max<i64>(P.n_gathered, 1)));
        CU_TRY(cudaMemset(op->d_gathered, 0, es * (size_t)std::max<i64>(P.n_gathered, 1)));
    }
    if (op->total_send > 0) {
        CU_TRY(cudaMalloc(&op->d_sendbuf, es * (size_t)op->total_send));
        std::vector<i64> idx;
        idx.reserve((size_t)op->total_send);
        for (auto& s : P.send_indices) idx.insert(idx.end(), s.begin(), s.end());
        CU_TRY(cudaMalloc(&op->d_send_idx, sizeof(i64) * idx.size()));
        CU_TRY(cudaMemcpy(op->d_send_idx, idx.data(), sizeof(i64) * idx.size(), cudaMemcpyHostToDevice));
    }
    if (!op->x_in_place && op->own_n > 0) {
        CU_TRY(cudaMalloc(&op->d_local_src, sizeof(i64) * (size_t)op->own_n));
        CU_TRY(cudaMalloc(&op->d_local_dst, sizeof(i64) * (size_t)op->own_n));
        CU_TRY(cudaMemcpy(op->d_local_src, P.local_src.data(), sizeof(i64) * (size_t)op->own_n, cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(op->d_local_dst, P.local_dst.data(), sizeof(i64) * (size_t)op->own_n, cudaMemcpyHostToDevice));
    }
    CU_TRY(cudaEventCreateWithFlags(&op->ev_x, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&op->ev_packed, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&op->ev_halo, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&op->ev_last, cudaEventDisableTiming));
    if (const char* e = getenv("HPCLA_TIMELINE")) op->timeline = e[0] == '1';
    if (op->timeline)
        for (cudaEvent_t& e : op->tl) CU_TRY(cudaEventCreate(&e));
    // tile lists per kernel class, interior / boundary: a tile is boundary iff one of its stored columns is a ghost
    {
        std::vector<unsigned char> flags((size_t)A->ntiles, 0);
        if (op->has_ghost && A->ntiles > 0) {
            unsigned char* d_flags = nullptr;
            CU_TRY(cudaMalloc(&d_flags, (size_t)A->ntiles));
            CU_TRY(launch_classify_tiles(A->itype, A->d_colval, A->d_tiles, A->ntiles, op->own_lo, op->own_n, d_flags, ctx->halo_stream));
            CU_TRY(cudaMemcpyAsync(flags.data(), d_flags, flags.size(), cudaMemcpyDeviceToHost, ctx->halo_stream));
            CU_TRY(cudaStreamSynchronize(ctx->halo_stream));
            cudaFree(d_flags);
        }
        std::vector<int> (&lists)[3][2] = op->h_list;
        for (i64 t = 0; t < A->ntiles; ++t) {
            const int c = A->tile_class[(size_t)t];
            if (c) lists[c - 1][flags[(size_t)t] ? 1 : 0].push_back((int)t);
        }
        rc = build_compact(op, lists);  // moves the interior row-walk tiles that can be compacted to lists[2][0]
        if (rc) return rc;
        for (int c = 0; c < 3; ++c)
            for (int g = 0; g < 2; ++g) {
                op->n_list[c][g] = (int)lists[c][g].size();
                if (lists[c][g].empty()) continue;
                for (size_t q = 0; q < lists[c][g].size(); ++q)
                    if (q == 0 || lists[c][g][q] != lists[c][g][q - 1] + 1) op->runs[c][g].push_back({(int)q, lists[c][g][q]});
                std::vector<TileRec> recs(lists[c][g].size());
                for (size_t q = 0; q < recs.size(); ++q) {
                    const TileDesc &t0 = A->h_tiles[(size_t)lists[c][g][q]], &t1 = A->h_tiles[(size_t)lists[c][g][q] + 1];
                    recs[q] = TileRec{t0.row, t1.row, t0.nnz, t1.nnz};
                }
                CU_TRY(cudaMalloc(&op->d_list[c][g], sizeof(TileRec) * recs.size()));
                CU_TRY(cudaMemcpy(op->d_list[c][g], recs.data(), sizeof(TileRec) * recs.size(), cudaMemcpyHostToDevice));
            }
    }
    op->seq = ctx->op_seq++;
    if (ctx->group) {
        hpcla_group* g = ctx->group;
        std::lock_guard<std::mutex> lk(g->mu);
        if ((i64)g->ops.size() <= op->seq) g->ops.resize((size_t)op->seq + 1, std::vector<hpcla_spmv*>(g->ctxs.size(), nullptr));
        g->ops[(size_t)op->seq][(size_t)ctx->rank] = op;
    }
    *out = guard.release();
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_info(const hpcla_spmv* op, int64_t* n_int, int64_t* n_bnd, int* x_in_place, int* sends_contiguous) {
    if (!op) return fail(HPCLA_ERR_ARG, "hpcla_spmv_info: null");
    if (n_int) *n_int = op->n_list[0][0] + op->n_list[1][0] + op->n_list[2][0];
    if (n_bnd) *n_bnd = op->n_list[0][1] + op->n_list[1][1];
    if (x_in_place) *x_in_place = op->x_in_place ? 1 : 0;
    if (sends_contiguous) *sends_contiguous = op->sends_contiguous ? 1 : 0;
    return HPCLA_OK;
}
extern "C" int hpcla_spmv_tile_lists(const hpcla_spmv* op, int64_t* out6) {
    if (!op || !out6) return fail(HPCLA_ERR_ARG, "hpcla_spmv_tile_lists: null");
    out6[0] = op->n_list[0][0], out6[1] = op->n_list[0][1], out6[2] = op->n_list[1][0], out6[3] = op->n_list[1][1];
    out6[4] = op->n_list[2][0];
    out6[5] = op->csr->flat.n_chunks;
    return HPCLA_OK;
}
extern "C" int64_t hpcla_spmv_launch_count(const hpcla_spmv* op) { return op ? op->launches : -1; }

extern "C" void hpcla_spmv_destroy(hpcla_spmv* op) {
    if (!op) return;
    cudaSetDevice(op->ctx->device);
    // wait for this operator's own work (halo stream + the end of its last call on the caller's stream), not the device
    cudaStreamSynchronize(op->ctx->halo_stream);
    if (op->ev_last) cudaEventSynchronize(op->ev_last);
    if (op->pipe) {
        if (op->ctx->h2d_stream) cudaStreamSynchronize(op->ctx->h2d_stream);
        if (op->ctx->d2h_stream) cudaStreamSynchronize(op->ctx->d2h_stream);
    }
    if (op->graph) cudaGraphExecDestroy(op->graph);
    if (op->cg_graph) cudaGraphExecDestroy(op->cg_graph);
    cudaFree(op->d_cg_scalars);
    for (cudaEvent_t e : op->tl)
        if (e) cudaEventDestroy(e);
    if (op->ev_last) cudaEventDestroy(op->ev_last);
    if (op->ctx->group) {
        hpcla_group* g = op->ctx->group;
        std::lock_guard<std::mutex> lk(g->mu);
        if ((i64)g->ops.size() > op->seq && g->ops[(size_t)op->seq][(size_t)op->ctx->rank] == op) g->ops[(size_t)op->seq][(size_t)op->ctx->rank] = nullptr;
    }
    cudaFree(op->d_gathered);
    cudaFree(op->d_ghost_rm);
    cudaFree(op->d_sendbuf_rm);
    cudaFree(op->d_sendbuf);
    cudaFree(op->d_send_idx);
    cudaFree(op->d_local_src);
    cudaFree(op->d_local_dst);
    for (int c = 0; c < 3; ++c)
        for (int g = 0; g < 2; ++g) cudaFree(op->d_list[c][g]);
    cudaFree(op->d_chdr);
    cudaFree(op->d_cpos);
    for (void* q : op->direct.ipc_opened) cudaIpcCloseMemHandle(q);
    cudaFree(op->direct.d_flags);
    cudaFree(op->d_dot_partials);
    if (op->pipe) {
        for (cudaEvent_t e : op->pipe->ev_in) cudaEventDestroy(e);
        for (cudaEvent_t e : op->pipe->ev_c) cudaEventDestroy(e);
        if (op->pipe->ev_start) cudaEventDestroy(op->pipe->ev_start);
        if (op->pipe->ev_tail) cudaEventDestroy(op->pipe->ev_tail);
        if (op->pipe->ev_done) cudaEventDestroy(op->pipe->ev_done);
        delete op->pipe;
    }
    if (op->ev_x) cudaEventDestroy(op->ev_x);
    if (op->ev_packed) cudaEventDestroy(op->ev_packed);
    if (op->ev_halo) cudaEventDestroy(op->ev_halo);
    delete op;
}

static ncclDataType_t nccl_type(int dtype, size_t* per_elem) {
    if (dtype == HPCLA_F32) { *per_elem = 1; return ncclFloat32; }
    if (dtype == HPCLA_F64) { *per_elem = 1; return ncclFloat64; }
    *per_elem = 2;
    return ncclFloat64;  // ComplexF64 travels as interleaved (re, im) doubles
}

static hpcla_spmv* group_peer(const hpcla_spmv* op, int peer) {
    hpcla_group* g = op->ctx->group;
    if (!g) return nullptr;
    std::lock_guard<std::mutex> lk(g->mu);
    if ((i64)g->ops.size() <= op->seq) return nullptr;
    return g->ops[(size_t)op->seq][(size_t)peer];
}

// ---------------------------------------------------------------------------------------------------------------
// Direct halo: export / connect (collective through the host: every rank hands its blob to every other rank)
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct HaloBlobHead {
    cudaIpcMemHandle_t mem, flags;
    long long raw_gathered, raw_flags, pid, device;
};
static_assert(sizeof(HaloBlobHead) == 160, "blob header layout");
}  // namespace

extern "C" int hpcla_spmv_halo_blob_size(const hpcla_spmv* op, int64_t* bytes_out) {
    if (!op || !bytes_out) return fail(HPCLA_ERR_ARG, "hpcla_spmv_halo_blob_size: null");
    *bytes_out = (int64_t)sizeof(HaloBlobHead) + 8 * (int64_t)op->ctx->nranks;
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_halo_export(hpcla_spmv* op, void* blob_out) {
    if (!op || !blob_out) return fail(HPCLA_ERR_ARG, "hpcla_spmv_halo_export: null");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    const int P = op->ctx->nranks;
    if (!stream_wait32()) return fail(HPCLA_ERR_CUDA, "hpcla_spmv_halo_export: the driver offers no cuStreamWaitValue32");
    if (!op->direct.d_flags) {
        CU_TRY(cudaMalloc(&op->direct.d_flags, sizeof(unsigned) * 2 * (size_t)P));
        CU_TRY(cudaMemset(op->direct.d_flags, 0, sizeof(unsigned) * 2 * (size_t)P));
    }
    HaloBlobHead h;
    std::memset(&h, 0, sizeof h);
    if (op->d_gathered) CU_TRY(cudaIpcGetMemHandle(&h.mem, op->d_gathered));
    CU_TRY(cudaIpcGetMemHandle(&h.flags, op->direct.d_flags));
    h.raw_gathered = (long long)(uintptr_t)op->d_gathered;
    h.raw_flags = (long long)(uintptr_t)op->direct.d_flags;
    h.pid = (long long)getpid();
    h.device = op->ctx->device;
    std::memcpy(blob_out, &h, sizeof h);
    long long* starts = reinterpret_cast<long long*>((char*)blob_out + sizeof h);
    for (int p = 0; p < P; ++p) starts[p] = 0;
    for (const Seg& r : op->recvs) starts[r.peer] = r.start;  // where rank r.peer's data lands in my gathered
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_halo_connect(hpcla_spmv* op, const void* blobs) {
    NvtxRange nvtx_range("hpcla_spmv_halo_connect");
    if (!op || !blobs) return fail(HPCLA_ERR_ARG, "hpcla_spmv_halo_connect: null");
    if (!op->direct.d_flags) return fail(HPCLA_ERR_STATE, "hpcla_spmv_halo_connect: call hpcla_spmv_halo_export first");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    const int P = op->ctx->nranks, me = op->ctx->rank;
    const size_t stride = sizeof(HaloBlobHead) + 8 * (size_t)P;
    auto& D = op->direct;
    D.peer_gathered.assign((size_t)P, nullptr);
    D.peer_flags.assign((size_t)P, nullptr);
    D.peer_recv_start.assign((size_t)P, 0);
    // Everything the exchange paths allocate lazily is allocated now: once streams may sit in a flag wait, a call that
    // synchronises the device (cudaMalloc / cudaFree) on one rank-thread of a single-process world would wait for a peer's
    // halo stream, which waits for this thread's push — a deadlock.  (Between processes each has its own context.)
    CU_TRY(preload_halo_kernels());
    if (!op->d_local_src && op->own_n > 0) {
        const hpcla_plan& PL = op->plan;
        CU_TRY(cudaMalloc(&op->d_local_src, sizeof(i64) * (size_t)op->own_n));
        CU_TRY(cudaMalloc(&op->d_local_dst, sizeof(i64) * (size_t)op->own_n));
        CU_TRY(cudaMemcpy(op->d_local_src, PL.local_src.data(), sizeof(i64) * (size_t)op->own_n, cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(op->d_local_dst, PL.local_dst.data(), sizeof(i64) * (size_t)op->own_n, cudaMemcpyHostToDevice));
    }
    std::vector<char> need((size_t)P, 0);
    for (const Seg& sg : op->sends) need[(size_t)sg.peer] |= 1;  // I write its gathered and its arrival flag
    for (const Seg& r : op->recvs) need[(size_t)r.peer] |= 2;   // I write its consumed flag
    for (int p = 0; p < P; ++p) {
        if (!need[(size_t)p] || p == me) continue;
        HaloBlobHead h;
        std::memcpy(&h, (const char*)blobs + (size_t)p * stride, sizeof h);
        const long long* starts = reinterpret_cast<const long long*>((const char*)blobs + (size_t)p * stride + sizeof h);
        D.peer_recv_start[(size_t)p] = starts[me];
        if ((need[(size_t)p] & 1) && starts[me] <= 0) return fail(HPCLA_ERR_STATE, "hpcla_spmv_halo_connect: rank %d does not expect data from rank %d", p, me);
        if (h.pid == (long long)getpid()) {  // same process (rank-threads): plain pointers (peer access was enabled with the group)
            D.peer_gathered[(size_t)p] = (char*)(uintptr_t)h.raw_gathered;
            D.peer_flags[(size_t)p] = (unsigned*)(uintptr_t)h.raw_flags;
        } else {
            void* q = nullptr;
            if (need[(size_t)p] & 1) {
                CU_TRY(cudaIpcOpenMemHandle(&q, h.mem, cudaIpcMemLazyEnablePeerAccess));
                D.peer_gathered[(size_t)p] = (char*)q;
                D.ipc_opened.push_back(q);
            }
            CU_TRY(cudaIpcOpenMemHandle(&q, h.flags, cudaIpcMemLazyEnablePeerAccess));
            D.peer_flags[(size_t)p] = (unsigned*)q;
            D.ipc_opened.push_back(q);
        }
    }
    D.on = true;
    return HPCLA_OK;
}

// debugging aid: the direct halo's flags and step counter, read on a private stream (works while other streams are blocked)
extern "C" int hpcla_spmv_halo_debug(hpcla_spmv* op, unsigned* out /* [2 * nranks + 1] */) {
    if (!op || !out || !op->direct.d_flags) return fail(HPCLA_ERR_ARG, "hpcla_spmv_halo_debug: no direct halo");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    cudaStream_t st;
    CU_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const int P = op->ctx->nranks;
    CU_TRY(cudaMemcpyAsync(out, op->direct.d_flags, sizeof(unsigned) * 2 * (size_t)P, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    cudaStreamDestroy(st);
    out[2 * P] = op->direct.step;
    return HPCLA_OK;
}

// Raise a flag (possibly in a peer's memory) behind everything already in the stream: a stream memory operation — no SM, no
// launch — where the driver accepts one for the address, else a one-thread kernel (HPCLA_HALO_FLAG_KERNEL=1 forces it).
static int direct_write_flag(hpcla_spmv* op, unsigned* flag, unsigned value, cudaStream_t stream) {
    static const bool force_kernel = [] { const char* e = getenv("HPCLA_HALO_FLAG_KERNEL"); return e && e[0] == '1'; }();
    if (!force_kernel && !op->direct.memop_failed) {
        fn_stream_write32 w = stream_write32();
        if (w && w((CUstream)stream, (CUdeviceptr)(uintptr_t)flag, value, CU_STREAM_WRITE_VALUE_DEFAULT) == CUDA_SUCCESS) return HPCLA_OK;
        op->direct.memop_failed = true;  // (not supported for this memory: the kernel from now on)
    }
    CU_TRY(launch_write_flag(flag, value, stream));
    op->launches += 1;
    return HPCLA_OK;
}

// the receiving side of the back-pressure: tell every rank I received from that its data of this step has been read
static int direct_signal_consumed(hpcla_spmv* op, cudaStream_t stream) {
    auto& D = op->direct;
    if (!D.on) return HPCLA_OK;
    const int P = op->ctx->nranks, me = op->ctx->rank;
    for (const Seg& r : op->recvs) {
        int rc = direct_write_flag(op, D.peer_flags[(size_t)r.peer] + P + me, D.step, stream);
        if (rc) return rc;
    }
    return HPCLA_OK;
}

// Send half of the exchange, on the halo stream: wait for x, pack what is not contiguous, and (NCCL world) post the
// grouped send/recv pairs so that ghosts land directly in their segments of `gathered` (no unpack: SURVEY §0.7).
static int exchange_begin(hpcla_spmv* op, const void* d_x, cudaStream_t stream, cudaEvent_t x_ready = nullptr) {
    hpcla_ctx* ctx = op->ctx;
    const int dtype = op->csr->dtype;
    const size_t es = dtype_size(dtype);
    cudaStream_t hs = ctx->halo_stream;
    if (x_ready) {  // staged multiply: x.v is complete when the last upload chunk has landed
        CU_TRY(cudaStreamWaitEvent(hs, x_ready, 0));
    } else {
        CU_TRY(cudaEventRecord(op->ev_x, stream));
        CU_TRY(cudaStreamWaitEvent(hs, op->ev_x, 0));
    }
    const bool group = ctx->group != nullptr;
    if (group && !op->direct.on) {  // (pull model only: with the direct halo I push, nobody reads my packed buffer)
        // my packed buffer may still be read by a peer's copy of the previous multiply
        for (const Seg& s : op->sends) {
            hpcla_spmv* peer = group_peer(op, s.peer);
            if (peer && peer->halo_recorded) CU_TRY(cudaStreamWaitEvent(hs, peer->ev_halo, 0));
        }
    }
    if (op->total_send > 0) {
        if (group) {  // single-process world: peers copy out of my packed buffer, so everything is staged there
            if (op->sends_contiguous) {
                for (const Seg& s : op->sends)
                    CU_TRY(cudaMemcpyAsync((char*)op->d_sendbuf + (size_t)s.start * es, (const char*)d_x + (size_t)(s.src0 - 1) * es, (size_t)s.count * es, cudaMemcpyDeviceToDevice, hs));
            } else {
                CU_TRY(launch_pack(dtype, d_x, op->d_send_idx, op->total_send, op->d_sendbuf, hs));
                op->launches += 1;
            }
        } else if (!op->sends_contiguous) {
            CU_TRY(launch_pack(dtype, d_x, op->d_send_idx, op->total_send, op->d_sendbuf, hs));
            op->launches += 1;
        }
    }
    CU_TRY(cudaEventRecord(op->ev_packed, hs));
    op->epoch.fetch_add(1, std::memory_order_release);
    if (op->direct.on) {
        // push: every run goes straight into its segment of the peer's `gathered` (copy engine over NVLink), then the
        // peer's arrival flag is raised; my own receives are awaited as flag values, not as kernels
        auto& D = op->direct;
        const int P = ctx->nranks, me = ctx->rank;
        const unsigned step = ++D.step;
        fn_stream_wait32 wait32 = stream_wait32();
        for (const Seg& s : op->sends) {
            // the peer has finished reading what I sent it last time
            if (wait32((CUstream)hs, (CUdeviceptr)(uintptr_t)(D.d_flags + P + s.peer), step - 1, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                return fail(HPCLA_ERR_CUDA, "cuStreamWaitValue32 failed");
            const bool from_x = !group && op->sends_contiguous;
            const char* src = from_x ? (const char*)d_x + (size_t)(s.src0 - 1) * es : (const char*)op->d_sendbuf + (size_t)s.start * es;
            CU_TRY(cudaMemcpyAsync(D.peer_gathered[(size_t)s.peer] + (size_t)(D.peer_recv_start[(size_t)s.peer] - 1) * es, src, (size_t)s.count * es, cudaMemcpyDefault, hs));
            int frc = direct_write_flag(op, D.peer_flags[(size_t)s.peer] + me, step, hs);
            if (frc) return frc;
        }
        for (const Seg& r : op->recvs)
            if (wait32((CUstream)hs, (CUdeviceptr)(uintptr_t)(D.d_flags + r.peer), step, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                return fail(HPCLA_ERR_CUDA, "cuStreamWaitValue32 failed");
        CU_TRY(cudaEventRecord(op->ev_halo, hs));
        op->halo_recorded = true;
        if (op->timeline) {
            CU_TRY(cudaEventRecord(op->tl[1], hs));
            op->tl_rec[1] = true;
        }
        return HPCLA_OK;
    }
    if (!group && ctx->comm) {
        NcclApi* api = nccl_api();
        size_t per = 1;
        const ncclDataType_t nt = nccl_type(dtype, &per);
        NCCL_TRY(api->GroupStart());
        for (const Seg& s : op->sends) {
            // all runs contiguous in x.v (every stencil): send straight from x, no pack kernel, no copy
            const char* src = op->sends_contiguous ? (const char*)d_x + (size_t)(s.src0 - 1) * es : (const char*)op->d_sendbuf + (size_t)s.start * es;
            NCCL_TRY_IN_GROUP(api->Send(src, (size_t)s.count * per, nt, s.peer, ctx->comm, hs));
        }
        for (const Seg& r : op->recvs) NCCL_TRY_IN_GROUP(api->Recv((char*)op->d_gathered + (size_t)(r.start - 1) * es, (size_t)r.count * per, nt, r.peer, ctx->comm, hs));
        NCCL_TRY(api->GroupEnd());
        CU_TRY(cudaEventRecord(op->ev_halo, hs));
        op->halo_recorded = true;
        if (op->timeline) {
        CU_TRY(cudaEventRecord(op->tl[1], hs));
        op->tl_rec[1] = true;
    }
    }
    return HPCLA_OK;
}

// Receive half for a single-process world: copy each peer's packed run into my ghost segment.
static int exchange_finish_group(hpcla_spmv* op) {
    hpcla_ctx* ctx = op->ctx;
    const size_t es = dtype_size(op->csr->dtype);
    cudaStream_t hs = ctx->halo_stream;
    for (const Seg& r : op->recvs) {
        hpcla_spmv* peer = group_peer(op, r.peer);
        if (!peer || peer->epoch.load(std::memory_order_acquire) < op->epoch.load(std::memory_order_acquire)) return fail(HPCLA_ERR_STATE, "hpcla_spmv_finish: rank %d has not begun the matching multiply", r.peer);
        const Seg* ps = nullptr;
        for (const Seg& s : peer->sends)
            if (s.peer == ctx->rank) ps = &s;
        if (!ps || ps->count != r.count) return fail(HPCLA_ERR_STATE, "hpcla_spmv_finish: rank %d sends %lld elements to rank %d, which expects %lld", r.peer, ps ? (long long)ps->count : 0LL, ctx->rank, (long long)r.count);
        CU_TRY(cudaStreamWaitEvent(hs, peer->ev_packed, 0));
        CU_TRY(cudaMemcpyAsync((char*)op->d_gathered + (size_t)(r.start - 1) * es, (const char*)peer->d_sendbuf + (size_t)ps->start * es, (size_t)r.count * es, cudaMemcpyDefault, hs));
    }
    CU_TRY(cudaEventRecord(op->ev_halo, hs));
    op->halo_recorded = true;
    return HPCLA_OK;
}

static void fill_launch(const hpcla_spmv* op, SpmvLaunch& L, const void* d_x, void* d_y) {
    const hpcla_csr* A = op->csr;
    const size_t es = dtype_size(A->dtype);
    L.dtype = A->dtype;
    L.itype = A->itype;
    L.rowptr = A->d_rowptr;
    L.colval = A->d_colval;
    L.nzval = A->d_nzval;
    L.nrows = A->nrows;
    L.nnz = A->nnz;
    L.shape = A->shape;
    L.x_own = op->x_in_place ? (const void*)((const char*)d_x + (size_t)(op->own_src0 - 1) * es)
                             : (const void*)((const char*)op->d_gathered + (size_t)(op->own_lo - 1) * es);
    L.gathered = op->d_gathered;
    L.own_lo = op->own_lo;
    L.own_n = op->own_n;
    L.has_ghost = op->has_ghost;
    L.y = d_y;
    L.long_threshold = A->long_threshold;
    L.hdrs = A->d_hdrs;
}

// Positions [lo, hi) of an ascending tile list as at most 8 runs of consecutive tiles (what lets a CTA find its window
// by arithmetic).  False (n_runs = 0) when the slice is more fragmented than that.
template <class Launch>
static bool tile_runs(const std::vector<std::pair<int, int>>& runs, int lo, int hi, Launch& L) {
    L.n_runs = 0;
    if (hi <= lo || runs.empty()) return false;
    // the run holding position lo: last run starting at or before it
    size_t j = (size_t)(std::upper_bound(runs.begin(), runs.end(), std::make_pair(lo, INT32_MAX)) - runs.begin()) - 1;
    int n = 0;
    for (; j < runs.size() && runs[j].first < hi; ++j) {
        if (n == 8) {
            L.n_runs = 0;
            return false;
        }
        const int start = std::max(runs[j].first, lo);
        L.run_cta0[n] = start - lo;
        L.run_tile0[n] = runs[j].second + (start - runs[j].first);
        ++n;
    }
    for (int k = n; k <= 8; ++k) L.run_cta0[k] = hi - lo;
    for (int k = n; k < 8; ++k) L.run_tile0[k] = 0;
    L.n_runs = n;
    return n > 0;
}

// HPCLA_RING: bit 0 = the multiply, bit 1 = sparse x dense on compact tiles run as rings of persistent CTAs
static int ring_mode() {
    const char* e = getenv("HPCLA_RING");  // (read per call: a tuning hook, also switched by the tests)
    return e ? atoi(e) : 0;
}

// both kernel classes over the interior (which = 0) or boundary (which = 1) tiles
// (from, to: positions in the lists, per class; nullptr = the whole lists)
static int launch_tiles(hpcla_spmv* op, SpmvLaunch& L, int which, cudaStream_t stream, const int* from = nullptr, const int* to = nullptr) {
    for (int c = 0; c < 3; ++c) {
        const int lo = from ? from[c] : 0, hi = to ? to[c] : op->n_list[c][which];
        L.recs = op->d_list[c][which] + lo;
        L.n_launch = hi - lo;
        if (L.n_launch <= 0) continue;
        const bool runs = tile_runs(op->runs[c][which], lo, hi, L);
        L.dot_x = nullptr;
        if (c != 1 && op->dot_request) {  // whole lists only (hpcla_cg): partials of the interior lists (class 0, then 2), then the boundary list
            L.dot_x = op->cur_x;
            L.dot_out = op->d_dot_partials + (which == 1 ? op->n_list[0][0] + op->n_list[2][0] : c == 2 ? op->n_list[0][0] : 0);
        }
        if (c == 2 && runs && (((uintptr_t)L.x_own) & 15) == 0) {  // compact tiles: x runs are fetched with 16-byte bulk copies
            CWalkLaunch C;
            C.dtype = L.dtype;
            C.lanes = L.shape.lanes;
            C.window = L.shape.window;
            C.nzval = L.nzval;
            C.nnz = L.nnz;
            C.sh = op->csh;
            C.hdrs = op->d_chdr;
            C.colpos = op->d_cpos;
            C.q0 = lo;
            C.tail_q_min = op->ctail_q_min;
            C.n_runs = L.n_runs;
            for (int j = 0; j < 9; ++j) C.run_cta0[j] = L.run_cta0[j];
            for (int j = 0; j < 8; ++j) C.run_tile0[j] = L.run_tile0[j];
            C.n_launch = L.n_launch;
            C.x_own = L.x_own;
            C.y = L.y;
            C.dot_x = L.dot_x;
            C.dot_out = L.dot_out;
            if ((ring_mode() & 1) && !L.dot_x) {  // the same tiles behind a ring of persistent CTAs (HPCLA_RING, see compact.cu)
                CWalkMLaunch M;
                M.dtype = C.dtype, M.lanes = C.lanes, M.window = C.window, M.nzval = C.nzval, M.nnz = C.nnz, M.sh = C.sh;
                M.hdrs = C.hdrs, M.colpos = C.colpos, M.q0 = C.q0, M.tail_q_min = C.tail_q_min, M.n_runs = C.n_runs;
                for (int j = 0; j < 9; ++j) M.run_cta0[j] = C.run_cta0[j];
                for (int j = 0; j < 8; ++j) M.run_tile0[j] = C.run_tile0[j];
                M.n_launch = C.n_launch, M.b_own = C.x_own, M.ldb = 0, M.c = C.y, M.ldc = 0, M.k0 = 0, M.kn = 1;
                if (M.lanes <= 8) {
                    CU_TRY(launch_cring(M, stream));
                    op->launches += 1;
                    continue;
                }
            }
            CU_TRY(launch_spmv_cwalk(C, stream));
        } else if (c != 1 && runs && !L.has_ghost && op->csr->d_hdrs && !op->dot_request) CU_TRY(launch_spmv_direct(L, stream));
        else if (c != 1) CU_TRY(launch_spmv_rowwalk(L, stream));
        else CU_TRY(launch_spmv_general(L, stream));
        op->launches += 1;
    }
    return HPCLA_OK;
}

// irregular matrices: the nnz-split multiply over all stored entries, then the fix-up of rows that span warp chunks
static int launch_flat(hpcla_spmv* op, const SpmvLaunch& base, cudaStream_t stream) {
    const hpcla_csr* A = op->csr;
    FlatLaunch F;
    F.dtype = A->dtype;
    F.itype = A->itype;
    F.colval = A->d_colval;
    F.nzval = A->d_nzval;
    F.nnz = A->nnz;
    F.flat = &A->flat;
    F.x_own = base.x_own;
    F.gathered = base.gathered;
    F.own_lo = base.own_lo;
    F.own_n = base.own_n;
    F.has_ghost = base.has_ghost;
    F.keep_x = A->flat_keep_x;
    F.y = base.y;
    F.long_threshold = A->long_threshold;
    // x is the one array this multiply re-reads (every stored entry gathers from it) while 40x as many bytes of matrix
    // stream through L2: when the own segment of x fits, it is pinned in the persisting part of L2 for the duration of the
    // launch (access-policy window on the stream).  Opt-in, HPCLA_FLAT_L2_WINDOW=1: measured, it cuts the DRAM traffic of BASELINE
    // config 4 from 9.5 GB to 3.6 GB per multiply but makes the multiply SLOWER (3.1 ms against 2.3 ms: the set-aside part of
    // L2 serves the gathers at a lower rate), profiles/r2c_*.
    bool window = false;
    const size_t xbytes = (size_t)op->own_n * dtype_size(A->dtype);
    if (A->flat_l2_window && base.x_own && xbytes > 0 && xbytes <= op->ctx->l2_persist_max) {
        hpcla_ctx* ctx = op->ctx;
        if (ctx->l2_persist_set < xbytes) {
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(ctx->l2_persist_max, xbytes + (xbytes >> 3))) == cudaSuccess) ctx->l2_persist_set = xbytes;
            else cudaGetLastError();
        }
        if (ctx->l2_persist_set >= xbytes) {
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = const_cast<void*>(base.x_own);
            attr.accessPolicyWindow.num_bytes = std::min(xbytes, ctx->l2_window_max);
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess) window = true;
            else cudaGetLastError();
        }
    }
    const cudaError_t le = launch_spmv_flat(F, stream);
    if (window) {
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    }
    CU_TRY(le);
    op->launches += 2 + (A->flat.n_empty > 0 ? 1 : 0);
    return HPCLA_OK;
}

static int launch_long(hpcla_spmv* op, const SpmvLaunch& base, cudaStream_t stream) {
    const hpcla_csr* A = op->csr;
    if (A->nlong == 0) return HPCLA_OK;
    LongRowsLaunch LL;
    LL.dtype = A->dtype;
    LL.itype = A->itype;
    LL.rowptr = A->d_rowptr;
    LL.colval = A->d_colval;
    LL.nzval = A->d_nzval;
    LL.long_rows = A->d_long_rows;
    LL.chunk_ptr = A->d_chunk_ptr;
    LL.nlong = A->nlong;
    LL.nchunks = A->nchunks;
    LL.chunk_nnz = A->chunk_nnz;
    LL.x_own = base.x_own;
    LL.gathered = base.gathered;
    LL.own_lo = base.own_lo;
    LL.own_n = base.own_n;
    LL.has_ghost = base.has_ghost;
    LL.partials = A->d_partials;
    LL.y = base.y;
    CU_TRY(launch_long_rows(LL, stream));
    op->launches += 2;
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_begin(hpcla_spmv* op, const void* d_x, void* d_y, void* stream_) {
    NvtxRange nvtx_range("hpcla_spmv_begin");
    if (!op || (op->n_x_local > 0 && !d_x) || (op->csr->nrows > 0 && !d_y)) return fail(HPCLA_ERR_ARG, "hpcla_spmv_begin: null");
    if (op->phase != 0) return fail(HPCLA_ERR_STATE, "hpcla_spmv_begin: the previous call was not finished");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    op->cur_x = d_x;
    op->cur_y = d_y;
    op->cur_stream = stream;
    if (op->timeline) {
        for (bool& b : op->tl_rec) b = false;
        CU_TRY(cudaEventRecord(op->tl[0], stream));
        op->tl_rec[0] = true;
    }
    if (op->has_peers) {
        rc = exchange_begin(op, d_x, stream);
        if (rc) return rc;
    }
    if (!op->x_in_place && op->own_n > 0) {  // local copy of src/vectors.jl:426-428, only when own columns have gaps
        CU_TRY(launch_local_copy(op->csr->dtype, d_x, op->d_local_src, op->d_local_dst, op->own_n, op->d_gathered, stream));
        op->launches += 1;
        if (op->has_ghost) CU_TRY(cudaEventRecord(op->ev_x, stream));  // the boundary tiles (halo stream) read this copy
    }
    SpmvLaunch L;
    fill_launch(op, L, d_x, d_y);
    if (op->csr->flat.n_chunks > 0) {
        // irregular matrix: every chunk of the nonzero stream may read ghosts (random columns), so with ghosts the whole
        // multiply runs behind the halo (hpcla_spmv_finish); without, here
        if (!op->has_ghost) {
            L.has_ghost = false;
            rc = launch_flat(op, L, stream);
            if (rc) return rc;
            rc = launch_long(op, L, stream);
            if (rc) return rc;
        }
        if (op->timeline) {
        CU_TRY(cudaEventRecord(op->tl[3], stream));
        op->tl_rec[3] = true;
    }
        op->phase = 1;
        return HPCLA_OK;
    }
    // interior tiles (all tiles when there are no ghosts) read own columns only: the ghost-free kernels, run while the
    // halo is in flight
    L.has_ghost = false;
    rc = launch_tiles(op, L, 0, stream);
    if (rc) return rc;
    if (op->timeline) {
        CU_TRY(cudaEventRecord(op->tl[3], stream));
        op->tl_rec[3] = true;
    }
    if (!op->has_ghost) {
        rc = launch_long(op, L, stream);
        if (rc) return rc;
    }
    op->phase = 1;
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_finish(hpcla_spmv* op) {
    NvtxRange nvtx_range("hpcla_spmv_finish");
    if (!op) return fail(HPCLA_ERR_ARG, "hpcla_spmv_finish: null");
    if (op->phase != 1) return fail(HPCLA_ERR_STATE, "hpcla_spmv_finish: no multiply in flight");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    cudaStream_t stream = op->cur_stream;
    cudaStream_t hs = op->ctx->halo_stream;
    if (op->has_peers && op->ctx->group && !op->direct.on) {
        rc = exchange_finish_group(op);
        if (rc) return rc;
    }
    const bool flat = op->csr->flat.n_chunks > 0;
    if (op->has_ghost && !flat) {
        // Boundary tiles go on the (high-priority) halo stream, right behind the receives: they run next to the
        // interior tiles instead of after them, so a step costs max(interior, halo + boundary), not their sum.
        SpmvLaunch L;
        fill_launch(op, L, op->cur_x, op->cur_y);
        if (!op->x_in_place) CU_TRY(cudaStreamWaitEvent(hs, op->ev_x, 0));  // own columns come from the local copy
        rc = launch_tiles(op, L, 1, hs);
        if (rc) return rc;
        CU_TRY(cudaEventRecord(op->ev_halo, hs));
        op->halo_recorded = true;
        if (op->direct.on && op->csr->nlong == 0) {
            // (direct halo) the boundary tiles were the last readers of the ghosts: tell the senders from here, so that the
            // caller's stream ends with the interior tiles and not with flag writes
            rc = direct_signal_consumed(op, hs);
            if (rc) return rc;
            op->direct.consumed_signalled = true;
        }
        if (op->timeline) {
        CU_TRY(cudaEventRecord(op->tl[2], hs));
        op->tl_rec[2] = true;
    }
    }
    if (op->has_peers || op->has_ghost) {
        // x (contiguous sends) and the packed buffer are read on the halo stream: the caller's stream must not run
        // ahead of them, nor ahead of the boundary rows of y
        CU_TRY(cudaStreamWaitEvent(stream, op->ev_halo, 0));
        if (op->ctx->group && op->has_peers) CU_TRY(cudaStreamWaitEvent(stream, op->ev_packed, 0));
    }
    if (op->has_ghost) {
        SpmvLaunch L;
        fill_launch(op, L, op->cur_x, op->cur_y);
        if (flat) {
            if (!op->x_in_place) CU_TRY(cudaStreamWaitEvent(stream, op->ev_x, 0));
            rc = launch_flat(op, L, stream);
            if (rc) return rc;
        }
        rc = launch_long(op, L, stream);
        if (rc) return rc;
    }
    if (!op->direct.consumed_signalled) {
        rc = direct_signal_consumed(op, stream);  // (direct halo) the ghosts of this step have been read
        if (rc) return rc;
    }
    op->direct.consumed_signalled = false;
    if (op->timeline) {
        CU_TRY(cudaEventRecord(op->tl[4], stream));
        op->tl_rec[4] = true;
    }
    CU_TRY(cudaEventRecord(op->ev_last, stream));
    op->phase = 0;
    return HPCLA_OK;
}

// Timeline of the most recent multiply (HPCLA_TIMELINE=1 when the operator was created): milliseconds from the moment
// x was ready on the caller's stream to [0] the end of the halo exchange (halo stream), [1] the end of the boundary
// tiles (halo stream), [2] the end of the interior tiles (caller's stream), [3] the end of the call (caller's stream);
// -1 where the step does not exist (no peers / no ghosts).  Blocks until the multiply has finished.
extern "C" int hpcla_spmv_timeline(hpcla_spmv* op, double* ms4_out) {
    if (!op || !ms4_out) return fail(HPCLA_ERR_ARG, "hpcla_spmv_timeline: null");
    if (!op->timeline) return fail(HPCLA_ERR_STATE, "hpcla_spmv_timeline: the operator was created without HPCLA_TIMELINE=1");
    if (!op->tl_rec[0] || !op->tl_rec[4]) return fail(HPCLA_ERR_STATE, "hpcla_spmv_timeline: no multiply has run yet");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    CU_TRY(cudaEventSynchronize(op->tl[4]));
    CU_TRY(cudaStreamSynchronize(op->ctx->halo_stream));
    for (int k = 1; k <= 4; ++k) {
        float ms = -1.f;
        if (op->tl_rec[k]) CU_TRY(cudaEventElapsedTime(&ms, op->tl[0], op->tl[k]));
        ms4_out[k - 1] = (double)ms;
    }
    return HPCLA_OK;
}

// One multiply as a CUDA graph bound to fixed x.v / y.v: the event choreography between the caller's stream and the
// halo stream becomes graph dependencies, the grouped ncclSend/ncclRecv a captured kernel node, and a replay costs one
// launch instead of ~10 driver calls — what the latency-bound strong-scaling regime needs.
// NCCL world or a single rank.  capture: (re)builds the graph; launch: replays it on `stream`.
extern "C" int hpcla_spmv_graph_capture(hpcla_spmv* op, const void* d_x, void* d_y, void* stream_) {
    NvtxRange nvtx_range("hpcla_spmv_graph_capture");
    if (!op) return fail(HPCLA_ERR_ARG, "hpcla_spmv_graph_capture: null");
    if (op->ctx->group && op->ctx->nranks > 1 && op->has_peers) return fail(HPCLA_ERR_STATE, "hpcla_spmv_graph_capture: needs an NCCL world or a single rank");
    if (op->phase != 0) return fail(HPCLA_ERR_STATE, "hpcla_spmv_graph_capture: the previous call was not finished");
    if (op->direct.on) return fail(HPCLA_ERR_STATE, "hpcla_spmv_graph_capture: the direct halo counts steps in its flags and cannot be replayed from a graph");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    (void)stream_;  // the capture runs on a library-owned stream: the caller's may be the legacy default stream, which cannot be captured
    hpcla_ctx* ctx = op->ctx;
    if (!ctx->cap_stream) CU_TRY(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
    cudaStream_t stream = ctx->cap_stream;
    if (op->graph) {
        cudaGraphExecDestroy(op->graph);
        op->graph = nullptr;
    }
    const bool tl = op->timeline;
    op->timeline = false;  // timing events are not captured
    const i64 before = op->launches;
    cudaGraph_t g = nullptr;
    CU_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed));
    rc = hpcla_spmv_run(op, d_x, d_y, stream);
    cudaError_t ce = cudaStreamEndCapture(stream, &g);
    op->timeline = tl;
    op->phase = 0;
    if (rc || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return rc ? rc : fail(HPCLA_ERR_CUDA, "hpcla_spmv_graph_capture: cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
    }
    op->graph_launches = op->launches - before;
    op->launches = before;
    ce = cudaGraphInstantiate(&op->graph, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) return fail(HPCLA_ERR_CUDA, "hpcla_spmv_graph_capture: cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
    op->graph_x = d_x;
    op->graph_y = d_y;
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_graph_launch(hpcla_spmv* op, void* stream_) {
    NvtxRange nvtx_range("hpcla_spmv_graph_launch");
    if (!op || !op->graph) return fail(HPCLA_ERR_STATE, "hpcla_spmv_graph_launch: no captured graph");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    CU_TRY(cudaGraphLaunch(op->graph, (cudaStream_t)stream_));
    CU_TRY(cudaEventRecord(op->ev_last, (cudaStream_t)stream_));
    op->launches += op->graph_launches;
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_run(hpcla_spmv* op, const void* d_x, void* d_y, void* stream) {
    if (op && op->ctx->group && op->ctx->nranks > 1 && op->has_peers && !op->direct.on)
        return fail(HPCLA_ERR_STATE, "hpcla_spmv_run: in a single-process world call hpcla_spmv_begin on every rank, then hpcla_spmv_finish");
    int rc = hpcla_spmv_begin(op, d_x, d_y, stream);
    if (rc) return rc;
    return hpcla_spmv_finish(op);
}

// ---------------------------------------------------------------------------------------------------------------
// sparse x dense: C = A * B, B and C column-major local blocks (HPCMatrix.A), all columns share one halo exchange
// ---------------------------------------------------------------------------------------------------------------
static i64 ghost_number(const hpcla_spmv* op, i64 pos1 /* 1-based position in gathered */) {
    return pos1 < op->own_lo ? pos1 - 1 : pos1 - 1 - op->own_n;
}

static int spmm_tiles(hpcla_spmv* op, int which, bool ghost, cudaStream_t stream) {
    const hpcla_csr* A = op->csr;
    const size_t es = dtype_size(A->dtype);
    SpmmLaunch L;
    L.dtype = A->dtype;
    L.itype = A->itype;
    L.rowptr = A->d_rowptr;
    L.colval = A->d_colval;
    L.nzval = A->d_nzval;
    L.nrows = A->nrows;
    L.nnz = A->nnz;
    L.shape = A->shape;
    L.b_own = (const char*)op->mm_B + (size_t)(op->own_src0 - 1) * es;
    L.ldb = op->mm_ldb;
    L.ghost = op->d_ghost_rm;
    L.ncols = op->mm_ncols;
    L.own_lo = op->own_lo;
    L.own_n = op->own_n;
    L.has_ghost = ghost;
    L.c = op->mm_C;
    L.ldc = op->mm_ldc;
    const bool walk = spmm_supports_rowwalk(A->shape);
    // 8 columns per pass when there are that many (measured: 1 014 -> 918 us for 8 columns of Poisson 256^3); HPCLA_SPMM_K8=0 disables
    static const bool spmm_k8 = [] { const char* e = getenv("HPCLA_SPMM_K8"); return e ? e[0] != '0' : true; }();
    static const bool spmm_compact = [] { const char* e = getenv("HPCLA_SPMM_COMPACT"); return e ? e[0] != '0' : true; }();
    for (int k0 = 0; k0 < op->mm_ncols;) {
        L.k0 = k0;
        const int left = op->mm_ncols - k0;
        L.kn = (left >= 8 && spmm_k8) ? 8 : left >= 4 ? 4 : 1;
        for (int c = 0; c < 3; ++c) {
            L.recs = op->d_list[c][which];
            L.n_launch = op->n_list[c][which];
            if (L.n_launch <= 0) continue;
            const bool runs = tile_runs(op->runs[c][which], 0, L.n_launch, L);
            if (c == 2 && runs && which == 0 && spmm_compact) {  // interior compact tiles: x runs of all kn columns staged by bulk copies
                CWalkMLaunch M;
                M.dtype = L.dtype;
                M.lanes = L.shape.lanes;
                M.window = L.shape.window;
                M.nzval = L.nzval;
                M.nnz = L.nnz;
                M.sh = op->csh;
                M.hdrs = op->d_chdr;
                M.colpos = op->d_cpos;
                M.q0 = 0;
                M.tail_q_min = op->ctail_q_min;
                M.n_runs = L.n_runs;
                for (int j = 0; j < 9; ++j) M.run_cta0[j] = L.run_cta0[j];
                for (int j = 0; j < 8; ++j) M.run_tile0[j] = L.run_tile0[j];
                M.n_launch = L.n_launch;
                M.b_own = L.b_own;
                M.ldb = L.ldb;
                M.c = L.c;
                M.ldc = L.ldc;
                M.k0 = L.k0;
                M.kn = L.kn;
                if (spmm_cwalk_supported(M)) {
                    if (ring_mode() & 2) CU_TRY(launch_cring(M, stream));
                    else CU_TRY(launch_spmm_cwalk(M, stream));
                    op->launches += 1;
                    continue;
                }
            }
            if (c != 1 && walk) CU_TRY(launch_spmm_rowwalk(L, stream));
            else CU_TRY(launch_spmm_rows(L, stream));
            op->launches += 1;
        }
        k0 += L.kn;
    }
    return HPCLA_OK;
}

extern "C" int hpcla_spmm_begin(hpcla_spmv* op, const void* d_B, int64_t ldb, void* d_C, int64_t ldc, int ncols, void* stream_) {
    NvtxRange nvtx_range("hpcla_spmm_begin");
    if (!op || ncols < 0 || (op->n_x_local > 0 && ncols > 0 && !d_B) || (op->csr->nrows > 0 && ncols > 0 && !d_C)) return fail(HPCLA_ERR_ARG, "hpcla_spmm_begin: bad arguments");
    if (ldb < op->n_x_local || ldc < op->csr->nrows) return fail(HPCLA_ERR_ARG, "hpcla_spmm_begin: leading dimensions smaller than the local blocks");
    if (op->phase != 0) return fail(HPCLA_ERR_STATE, "hpcla_spmm_begin: the previous call was not finished");
    if (!op->x_in_place) return fail(HPCLA_ERR_STATE, "hpcla_spmm_begin: the own columns of A are not a contiguous run of B's local rows; multiply column by column");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    hpcla_ctx* ctx = op->ctx;
    cudaStream_t stream = (cudaStream_t)stream_, hs = ctx->halo_stream;
    const int dtype = op->csr->dtype;
    const size_t es = dtype_size(dtype);
    op->mm_B = d_B;
    op->mm_C = d_C;
    op->mm_ldb = ldb;
    op->mm_ldc = ldc;
    op->mm_ncols = ncols;
    op->cur_stream = stream;
    if (ncols == 0) {
        op->phase = 3;
        return HPCLA_OK;
    }
    const i64 n_ghost = op->plan.n_gathered - op->own_n;
    if (op->has_peers && ncols > op->mm_cols) {  // (re)size the exchange buffers
        CU_TRY(cudaDeviceSynchronize());
        cudaFree(op->d_ghost_rm);
        cudaFree(op->d_sendbuf_rm);
        op->d_ghost_rm = op->d_sendbuf_rm = nullptr;
        CU_TRY(cudaMalloc(&op->d_ghost_rm, es * (size_t)std::max<i64>(n_ghost, 1) * (size_t)ncols));
        CU_TRY(cudaMalloc(&op->d_sendbuf_rm, es * (size_t)std::max<i64>(op->total_send, 1) * (size_t)ncols));
        op->mm_cols = ncols;
    }
    if (op->has_peers) {
        CU_TRY(cudaEventRecord(op->ev_x, stream));
        CU_TRY(cudaStreamWaitEvent(hs, op->ev_x, 0));
        const bool group = ctx->group != nullptr;
        if (group)  // my packed rows may still be read by a peer's copy of the previous product
            for (const Seg& sg : op->sends) {
                hpcla_spmv* peer = group_peer(op, sg.peer);
                if (peer && peer->halo_recorded) CU_TRY(cudaStreamWaitEvent(hs, peer->ev_halo, 0));
            }
        if (op->total_send > 0) {
            CU_TRY(launch_pack_rows(dtype, d_B, ldb, op->d_send_idx, op->total_send, ncols, op->d_sendbuf_rm, hs));
            op->launches += 1;
        }
        CU_TRY(cudaEventRecord(op->ev_packed, hs));
        op->epoch.fetch_add(1, std::memory_order_release);
        if (!group && ctx->comm) {
            NcclApi* api = nccl_api();
            size_t per = 1;
            const ncclDataType_t nt = nccl_type(dtype, &per);
            NCCL_TRY(api->GroupStart());
            for (const Seg& sg : op->sends)  // one message per peer: its rows, all columns adjacent
                NCCL_TRY_IN_GROUP(api->Send((const char*)op->d_sendbuf_rm + (size_t)sg.start * ncols * es, (size_t)sg.count * ncols * per, nt, sg.peer, ctx->comm, hs));
            for (const Seg& r : op->recvs)
                NCCL_TRY_IN_GROUP(api->Recv((char*)op->d_ghost_rm + (size_t)ghost_number(op, r.start) * ncols * es, (size_t)r.count * ncols * per, nt, r.peer, ctx->comm, hs));
            NCCL_TRY(api->GroupEnd());
        }
    }
    rc = spmm_tiles(op, 0, false, stream);  // interior tiles while the halo is in flight
    if (rc) return rc;
    op->phase = 3;  // only now: a failure above leaves the operator idle, not wedged
    return HPCLA_OK;
}

extern "C" int hpcla_spmm_finish(hpcla_spmv* op) {
    NvtxRange nvtx_range("hpcla_spmm_finish");
    if (!op) return fail(HPCLA_ERR_ARG, "hpcla_spmm_finish: null");
    if (op->phase != 3) return fail(HPCLA_ERR_STATE, "hpcla_spmm_finish: no product in flight");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    hpcla_ctx* ctx = op->ctx;
    cudaStream_t stream = op->cur_stream, hs = ctx->halo_stream;
    const size_t es = dtype_size(op->csr->dtype);
    const int ncols = op->mm_ncols;
    op->phase = 0;
    if (ncols == 0 || !op->has_peers) return HPCLA_OK;
    if (ctx->group) {  // single-process world: copy each peer's packed rows into my ghost rows
        for (const Seg& r : op->recvs) {
            hpcla_spmv* peer = group_peer(op, r.peer);
            if (!peer || peer->epoch.load(std::memory_order_acquire) < op->epoch.load(std::memory_order_acquire) || peer->mm_ncols != ncols)
                return fail(HPCLA_ERR_STATE, "hpcla_spmm_finish: rank %d has not begun the matching product", r.peer);
            const Seg* ps = nullptr;
            for (const Seg& sg : peer->sends)
                if (sg.peer == ctx->rank) ps = &sg;
            if (!ps || ps->count != r.count) return fail(HPCLA_ERR_STATE, "hpcla_spmm_finish: rank %d and rank %d disagree on the halo size", r.peer, ctx->rank);
            CU_TRY(cudaStreamWaitEvent(hs, peer->ev_packed, 0));
            CU_TRY(cudaMemcpyAsync((char*)op->d_ghost_rm + (size_t)ghost_number(op, r.start) * ncols * es, (const char*)peer->d_sendbuf_rm + (size_t)ps->start * ncols * es,
                                   (size_t)r.count * ncols * es, cudaMemcpyDefault, hs));
        }
    }
    if (op->has_ghost) {  // boundary tiles behind the receives, on the halo stream
        rc = spmm_tiles(op, 1, true, hs);
        if (rc) return rc;
    }
    CU_TRY(cudaEventRecord(op->ev_halo, hs));
    op->halo_recorded = true;
    CU_TRY(cudaStreamWaitEvent(stream, op->ev_halo, 0));
    if (ctx->group) CU_TRY(cudaStreamWaitEvent(stream, op->ev_packed, 0));
    return HPCLA_OK;
}

extern "C" int hpcla_spmm_run(hpcla_spmv* op, const void* d_B, int64_t ldb, void* d_C, int64_t ldc, int ncols, void* stream) {
    if (op && op->ctx->group && op->ctx->nranks > 1 && op->has_peers)
        return fail(HPCLA_ERR_STATE, "hpcla_spmm_run: in a single-process world call hpcla_spmm_begin on every rank, then hpcla_spmm_finish");
    int rc = hpcla_spmm_begin(op, d_B, ldb, d_C, ldc, ncols, stream);
    if (rc) return rc;
    return hpcla_spmm_finish(op);
}

// ---------------------------------------------------------------------------------------------------------------
// staged multiply: host x -> x.v, y.v = A * x, y.v -> host y, pipelined over row blocks
// ---------------------------------------------------------------------------------------------------------------
static int build_pipe(hpcla_spmv* op) {
    hpcla_ctx* ctx = op->ctx;
    const hpcla_csr* A = op->csr;
    HostPipe* P = new HostPipe();
    op->pipe = P;
    if (!ctx->h2d_stream) CU_TRY(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    if (!ctx->d2h_stream) CU_TRY(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreateWithFlags(&P->ev_start, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&P->ev_tail, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&P->ev_done, cudaEventDisableTiming));
    const size_t es = dtype_size(A->dtype);
    i64 chunk_bytes = 8 << 20;  // measured: 8 MiB chunks give the best PCIe overlap (profiles/r1e_staged_chunks.txt)
    if (const char* e = getenv("HPCLA_STAGE_CHUNK_KB")) chunk_bytes = std::max<i64>(64, atoll(e)) << 10;
    const i64 bytes = std::max(op->n_x_local, A->nrows) * (i64)es;
    int nb = (int)std::min<i64>(64, std::max<i64>(1, (bytes + chunk_bytes - 1) / chunk_bytes));
    if ((i64)nb > A->ntiles) nb = (int)std::max<i64>(1, A->ntiles);
    // pipelining needs x.v read in place (own columns straight from x.v) and something to overlap
    P->usable = op->x_in_place && nb >= 2 && A->nrows > 0 && op->n_x_local > 0 && A->flat.n_chunks == 0;  // (a flat multiply needs all of x)
    if (!P->usable) return HPCLA_OK;
    P->nb = nb;
    // block boundaries in tiles, rows, and list positions
    const std::vector<TileDesc>& tiles = A->h_tiles;
    std::vector<i64> tb((size_t)nb + 1);
    P->row_at.resize((size_t)nb + 1);
    for (int k = 0; k <= nb; ++k) {
        tb[(size_t)k] = (i64)k * A->ntiles / nb;
        P->row_at[(size_t)k] = k == nb ? A->nrows : tiles[(size_t)tb[(size_t)k]].row;
    }
    for (int c = 0; c < 3; ++c)
        for (int g = 0; g < 2; ++g) {
            const std::vector<int>& l = op->h_list[c][g];
            P->pos[c][g].resize((size_t)nb + 1);
            for (int k = 0; k <= nb; ++k) P->pos[c][g][(size_t)k] = (int)(std::lower_bound(l.begin(), l.end(), (int)tb[(size_t)k]) - l.begin());
        }
    // x prefix needed by the interior tiles of each block
    std::vector<i64> maxcol((size_t)A->ntiles, -1);
    {
        i64* d_max = nullptr;
        CU_TRY(cudaMalloc(&d_max, sizeof(i64) * (size_t)A->ntiles));
        CU_TRY(launch_tile_maxcol(A->itype, A->d_colval, A->d_tiles, A->ntiles, op->own_lo, op->own_n, d_max, ctx->halo_stream));
        CU_TRY(cudaMemcpyAsync(maxcol.data(), d_max, sizeof(i64) * maxcol.size(), cudaMemcpyDeviceToHost, ctx->halo_stream));
        CU_TRY(cudaStreamSynchronize(ctx->halo_stream));
        cudaFree(d_max);
    }
    P->xchunk.resize((size_t)nb + 1);
    for (int k = 0; k <= nb; ++k) P->xchunk[(size_t)k] = (i64)k * op->n_x_local / nb;
    P->in_chunk.assign((size_t)nb, -1);
    P->late.assign((size_t)nb, 0);
    int running = -1;
    for (int k = 0; k < nb; ++k) {
        i64 need = -1;
        for (int c = 0; c < 3; ++c)
            for (int q = P->pos[c][0][(size_t)k]; q < P->pos[c][0][(size_t)k + 1]; ++q) need = std::max(need, maxcol[(size_t)op->h_list[c][0][(size_t)q]]);
        if (need >= 0) {
            const i64 xi = need + op->own_src0 - 1;  // 0-based index into x.v
            int j = (int)(std::upper_bound(P->xchunk.begin(), P->xchunk.end(), xi) - P->xchunk.begin()) - 1;
            j = std::min(std::max(j, 0), nb - 1);
            running = std::max(running, j);
        }
        P->in_chunk[(size_t)k] = running;
        for (int c = 0; c < 3; ++c)
            if (P->pos[c][1][(size_t)k + 1] > P->pos[c][1][(size_t)k]) P->late[(size_t)k] = 1;
    }
    if (A->nlong > 0) {
        std::vector<i64> rows((size_t)A->nlong);
        CU_TRY(cudaMemcpy(rows.data(), A->d_long_rows, sizeof(i64) * rows.size(), cudaMemcpyDeviceToHost));
        for (i64 r : rows) {
            const int k = (int)(std::upper_bound(P->row_at.begin(), P->row_at.end(), r) - P->row_at.begin()) - 1;
            if (k >= 0 && k < nb) P->late[(size_t)k] = 1;
        }
    }
    P->ev_in.resize((size_t)nb);
    P->ev_c.resize((size_t)nb);
    for (int k = 0; k < nb; ++k) {
        CU_TRY(cudaEventCreateWithFlags(&P->ev_in[(size_t)k], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&P->ev_c[(size_t)k], cudaEventDisableTiming));
    }
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_run_staged(hpcla_spmv* op, const void* h_x, void* d_x, void* d_y, void* h_y, void* stream_) {
    NvtxRange nvtx_range("hpcla_spmv_run_staged");
    if (!op || (op->n_x_local > 0 && (!h_x || !d_x)) || (op->csr->nrows > 0 && (!h_y || !d_y))) return fail(HPCLA_ERR_ARG, "hpcla_spmv_run_staged: null");
    if (op->phase != 0) return fail(HPCLA_ERR_STATE, "hpcla_spmv_run_staged: the previous call was not finished");
    hpcla_ctx* ctx = op->ctx;
    if (ctx->group && ctx->nranks > 1 && op->has_peers) return fail(HPCLA_ERR_STATE, "hpcla_spmv_run_staged: needs an NCCL world or a single rank");
    int rc = set_device(ctx);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const hpcla_csr* A = op->csr;
    const size_t es = dtype_size(A->dtype);
    if (!op->pipe) {
        rc = build_pipe(op);
        if (rc) return rc;
    }
    HostPipe* P = op->pipe;
    if (!P->usable) {  // nothing to overlap: copy, multiply, copy on the caller's stream
        if (op->n_x_local > 0) CU_TRY(cudaMemcpyAsync(d_x, h_x, es * (size_t)op->n_x_local, cudaMemcpyHostToDevice, stream));
        rc = hpcla_spmv_run(op, d_x, d_y, stream);
        if (rc) return rc;
        if (A->nrows > 0) CU_TRY(cudaMemcpyAsync(h_y, d_y, es * (size_t)A->nrows, cudaMemcpyDeviceToHost, stream));
        return HPCLA_OK;
    }
    const int nb = P->nb;
    cudaStream_t in = ctx->h2d_stream, out = ctx->d2h_stream;
    // x.v and y.v may still be in use by earlier work of the caller's stream
    CU_TRY(cudaEventRecord(P->ev_start, stream));
    CU_TRY(cudaStreamWaitEvent(in, P->ev_start, 0));
    CU_TRY(cudaStreamWaitEvent(out, P->ev_start, 0));
    for (int j = 0; j < nb; ++j) {
        const i64 b = P->xchunk[(size_t)j], e = P->xchunk[(size_t)j + 1];
        if (e > b) CU_TRY(cudaMemcpyAsync((char*)d_x + (size_t)b * es, (const char*)h_x + (size_t)b * es, (size_t)(e - b) * es, cudaMemcpyHostToDevice, in));
        CU_TRY(cudaEventRecord(P->ev_in[(size_t)j], in));
    }
    op->cur_x = d_x;
    op->cur_y = d_y;
    op->cur_stream = stream;
    if (op->has_peers) {
        rc = exchange_begin(op, d_x, stream, P->ev_in[(size_t)nb - 1]);
        if (rc) return rc;
    }
    SpmvLaunch L;
    fill_launch(op, L, d_x, d_y);
    L.has_ghost = false;
    int waited = -1;
    for (int k = 0; k < nb; ++k) {
        if (P->in_chunk[(size_t)k] > waited) {
            waited = P->in_chunk[(size_t)k];
            CU_TRY(cudaStreamWaitEvent(stream, P->ev_in[(size_t)waited], 0));
        }
        const int from[3] = {P->pos[0][0][(size_t)k], P->pos[1][0][(size_t)k], P->pos[2][0][(size_t)k]};
        const int to[3] = {P->pos[0][0][(size_t)k + 1], P->pos[1][0][(size_t)k + 1], P->pos[2][0][(size_t)k + 1]};
        rc = launch_tiles(op, L, 0, stream, from, to);
        if (rc) return rc;
        if (!P->late[(size_t)k]) {
            const i64 r0 = P->row_at[(size_t)k], r1 = P->row_at[(size_t)k + 1];
            if (r1 > r0) {
                CU_TRY(cudaEventRecord(P->ev_c[(size_t)k], stream));
                CU_TRY(cudaStreamWaitEvent(out, P->ev_c[(size_t)k], 0));
                CU_TRY(cudaMemcpyAsync((char*)h_y + (size_t)r0 * es, (const char*)d_y + (size_t)r0 * es, (size_t)(r1 - r0) * es, cudaMemcpyDeviceToHost, out));
            }
        }
    }
    // the tail: boundary tiles (need the ghosts) and split long rows (need all of x), then the late slices of y
    if (waited < nb - 1) CU_TRY(cudaStreamWaitEvent(stream, P->ev_in[(size_t)nb - 1], 0));
    if (op->has_peers) CU_TRY(cudaStreamWaitEvent(stream, op->ev_halo, 0));
    fill_launch(op, L, d_x, d_y);
    if (op->has_ghost) {
        rc = launch_tiles(op, L, 1, stream);
        if (rc) return rc;
    }
    rc = launch_long(op, L, stream);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(P->ev_tail, stream));
    CU_TRY(cudaStreamWaitEvent(out, P->ev_tail, 0));
    for (int k = 0; k < nb; ++k) {
        if (!P->late[(size_t)k]) continue;
        const i64 r0 = P->row_at[(size_t)k], r1 = P->row_at[(size_t)k + 1];
        if (r1 > r0) CU_TRY(cudaMemcpyAsync((char*)h_y + (size_t)r0 * es, (const char*)d_y + (size_t)r0 * es, (size_t)(r1 - r0) * es, cudaMemcpyDeviceToHost, out));
    }
    CU_TRY(cudaEventRecord(P->ev_done, out));
    CU_TRY(cudaStreamWaitEvent(stream, P->ev_done, 0));  // the caller's stream now orders after host y is complete
    rc = direct_signal_consumed(op, stream);
    if (rc) return rc;
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_gather(hpcla_spmv* op, const void* d_x, void* stream_, void** d_gathered_out) {
    NvtxRange nvtx_range("hpcla_spmv_gather");
    if (!op || !d_gathered_out) return fail(HPCLA_ERR_ARG, "hpcla_spmv_gather: null");
    if (op->phase != 0) return fail(HPCLA_ERR_STATE, "hpcla_spmv_gather: the previous call was not finished");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const hpcla_plan& P = op->plan;
    const size_t es = dtype_size(op->csr->dtype);
    if (!op->d_gathered) {
        CU_TRY(cudaMalloc(&op->d_gathered, es * (size_t)std::max<i64>(P.n_gathered, 1)));
        CU_TRY(cudaMemset(op->d_gathered, 0, es * (size_t)std::max<i64>(P.n_gathered, 1)));
    }
    if (!op->d_local_src && op->own_n > 0) {
        CU_TRY(cudaMalloc(&op->d_local_src, sizeof(i64) * (size_t)op->own_n));
        CU_TRY(cudaMalloc(&op->d_local_dst, sizeof(i64) * (size_t)op->own_n));
        CU_TRY(cudaMemcpy(op->d_local_src, P.local_src.data(), sizeof(i64) * (size_t)op->own_n, cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(op->d_local_dst, P.local_dst.data(), sizeof(i64) * (size_t)op->own_n, cudaMemcpyHostToDevice));
    }
    op->cur_x = d_x;
    op->cur_stream = stream;
    if (op->has_peers) {
        rc = exchange_begin(op, d_x, stream);
        if (rc) return rc;
    }
    if (op->own_n > 0) {
        CU_TRY(launch_local_copy(op->csr->dtype, d_x, op->d_local_src, op->d_local_dst, op->own_n, op->d_gathered, stream));
        op->launches += 1;
    }
    op->phase = 2;
    *d_gathered_out = op->d_gathered;
    if (!(op->ctx->group && op->ctx->nranks > 1 && op->has_peers)) return hpcla_spmv_gather_finish(op);
    return HPCLA_OK;
}

extern "C" int hpcla_spmv_gather_finish(hpcla_spmv* op) {
    if (!op) return fail(HPCLA_ERR_ARG, "hpcla_spmv_gather_finish: null");
    if (op->phase == 0) return HPCLA_OK;  // already completed by hpcla_spmv_gather (NCCL world / single rank)
    if (op->phase != 2) return fail(HPCLA_ERR_STATE, "hpcla_spmv_gather_finish: no gather in flight");
    int rc = set_device(op->ctx);
    if (rc) return rc;
    if (op->has_peers) {
        if (op->ctx->group && !op->direct.on) {
            rc = exchange_finish_group(op);
            if (rc) return rc;
        }
        CU_TRY(cudaStreamWaitEvent(op->cur_stream, op->ev_halo, 0));
        if (op->ctx->group) CU_TRY(cudaStreamWaitEvent(op->cur_stream, op->ev_packed, 0));
        // (direct halo) `gathered` stays valid for the caller until the next exchange begins; the peers may overwrite it then
        rc = direct_signal_consumed(op, op->cur_stream);
        if (rc) return rc;
    }
    op->phase = 0;
    return HPCLA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// reductions and updates
// ---------------------------------------------------------------------------------------------------------------
static int dot_to_host(hpcla_ctx* ctx, int dtype, i64 n, const void* d_x, const void* d_y, cudaStream_t stream, double out2[2]) {
    int rc = set_device(ctx);
    if (rc) return rc;
    CU_TRY(launch_dot(dtype, n, d_x, d_y, ctx->d_red_scratch, ctx->d_red_out, stream));
    if (ctx->comm && ctx->nranks > 1) NCCL_TRY(nccl_api()->AllReduce(ctx->d_red_out, ctx->d_red_out, 2, ncclFloat64, ncclSum, ctx->comm, stream));
    CU_TRY(cudaMemcpyAsync(ctx->h_red_out, ctx->d_red_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaStreamSynchronize(stream));
    out2[0] = ctx->h_red_out[0];
    out2[1] = ctx->h_red_out[1];
    return HPCLA_OK;
}

extern "C" int hpcla_dot(hpcla_ctx* ctx, int dtype, int64_t n, const void* d_x, const void* d_y, void* result_out, void* stream) {
    if (!ctx || !result_out || n < 0 || !dtype_size(dtype)) return fail(HPCLA_ERR_ARG, "hpcla_dot: bad arguments");
    double r[2];
    int rc = dot_to_host(ctx, dtype, n, d_x, d_y, (cudaStream_t)stream, r);
    if (rc) return rc;
    if (dtype == HPCLA_F32) *(float*)result_out = (float)r[0];
    else if (dtype == HPCLA_F64) *(double*)result_out = r[0];
    else ((double*)result_out)[0] = r[0], ((double*)result_out)[1] = r[1];
    return HPCLA_OK;
}

extern "C" int hpcla_nrm2(hpcla_ctx* ctx, int dtype, int64_t n, const void* d_x, void* result_out, void* stream) {
    if (!ctx || !result_out || n < 0 || !dtype_size(dtype)) return fail(HPCLA_ERR_ARG, "hpcla_nrm2: bad arguments");
    double r[2];
    int rc = dot_to_host(ctx, dtype, n, d_x, d_x, (cudaStream_t)stream, r);
    if (rc) return rc;
    const double v = ctx->group && ctx->nranks > 1 ? r[0] : std::sqrt(r[0]);
    if (dtype == HPCLA_F32) *(float*)result_out = (float)v;
    else *(double*)result_out = v;
    return HPCLA_OK;
}

extern "C" int hpcla_axpby(hpcla_ctx* ctx, int dtype, int64_t n, const void* alpha, const void* d_x, const void* beta, void* d_y, void* stream) {
    if (!ctx || !alpha || !beta || n < 0 || !dtype_size(dtype)) return fail(HPCLA_ERR_ARG, "hpcla_axpby: bad arguments");
    int rc = set_device(ctx);
    if (rc) return rc;
    CU_TRY(launch_axpby(dtype, n, alpha, d_x, beta, d_y, (cudaStream_t)stream));
    return HPCLA_OK;
}

// execute_plan!(plan::VectorRepartitionPlan, x) — src/vectors.jl:624-676 — on the device: the local overlap is one
// device-to-device copy, every other overlap one ncclSend / ncclRecv of a contiguous range straight between x.v and the
// result (the reference stages the whole vector through the host and packs per-peer buffers, tag 92).
extern "C" int hpcla_repartition_run(hpcla_ctx* ctx, int dtype, int64_t n_send, const int64_t* send_rank_ids, const int64_t* send_first,
                                     const int64_t* send_count, int64_t n_recv, const int64_t* recv_rank_ids, const int64_t* recv_count,
                                     const int64_t* recv_offset, const int64_t* local3, const void* d_src, void* d_dst, void* stream_) {
    NvtxRange nvtx_range("hpcla_repartition_run");
    if (!ctx || !dtype_size(dtype) || n_send < 0 || n_recv < 0 || !local3) return fail(HPCLA_ERR_ARG, "hpcla_repartition_run: bad arguments");
    if ((n_send > 0 || n_recv > 0) && !ctx->comm) return fail(HPCLA_ERR_STATE, "hpcla_repartition_run: the plan exchanges data but the context has no NCCL communicator");
    int rc = set_device(ctx);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t es = dtype_size(dtype);
    if (local3[1] > 0)
        CU_TRY(cudaMemcpyAsync((char*)d_dst + (size_t)(local3[2] - 1) * es, (const char*)d_src + (size_t)(local3[0] - 1) * es, (size_t)local3[1] * es, cudaMemcpyDeviceToDevice, stream));
    if (n_send > 0 || n_recv > 0) {
        NcclApi* api = nccl_api();
        size_t per = 1;
        const ncclDataType_t nt = nccl_type(dtype, &per);
        NCCL_TRY(api->GroupStart());
        for (i64 i = 0; i < n_send; ++i)
            NCCL_TRY_IN_GROUP(api->Send((const char*)d_src + (size_t)(send_first[i] - 1) * es, (size_t)send_count[i] * per, nt, (int)send_rank_ids[i], ctx->comm, stream));
        for (i64 i = 0; i < n_recv; ++i)
            NCCL_TRY_IN_GROUP(api->Recv((char*)d_dst + (size_t)(recv_offset[i] - 1) * es, (size_t)recv_count[i] * per, nt, (int)recv_rank_ids[i], ctx->comm, stream));
        NCCL_TRY(api->GroupEnd());
    }
    return HPCLA_OK;
}

// The device side of hpcla_cg: everything is enqueued, nothing is read back (capturable into a CUDA graph).
static int cg_enqueue(hpcla_spmv* op, const void* d_b, void* d_x, char* r, char* p, char* q, int iters, bool fused, i64 n_partials, double* d_s,
                      double* d_pq, cudaStream_t stream) {
    hpcla_ctx* ctx = op->ctx;
    const int dtype = op->csr->dtype;
    const i64 n = op->csr->nrows;
    const bool multi = ctx->comm && ctx->nranks > 1;
    NcclApi* api = multi ? nccl_api() : nullptr;
    CU_TRY(launch_cg_init(dtype, n, d_b, d_x, r, p, ctx->d_red_scratch, d_s, stream));
    op->launches += 1;
    if (multi) NCCL_TRY(api->AllReduce(d_s, d_s, 1, ncclFloat64, ncclSum, ctx->comm, stream));
    for (int k = 0; k < iters; ++k) {
        op->dot_request = fused;
        int rc = hpcla_spmv_run(op, p, q, stream);
        op->dot_request = false;
        if (rc) return rc;
        if (fused) CU_TRY(launch_dot_partials_sum(op->d_dot_partials, n_partials, d_pq, stream));
        else CU_TRY(launch_dot(dtype, n, p, q, ctx->d_red_scratch, d_pq, stream));
        if (multi) NCCL_TRY(api->AllReduce(d_pq, d_pq, 1, ncclFloat64, ncclSum, ctx->comm, stream));
        CU_TRY(launch_cg_update_xr(dtype, n, p, q, d_x, r, d_s + 2 * k, d_pq, ctx->d_red_scratch, d_s + 2 * (k + 1), stream));
        if (multi) NCCL_TRY(api->AllReduce(d_s + 2 * (k + 1), d_s + 2 * (k + 1), 1, ncclFloat64, ncclSum, ctx->comm, stream));
        CU_TRY(launch_cg_update_p(dtype, n, r, p, d_s + 2 * (k + 1), d_s + 2 * k, stream));
        op->launches += 3;
    }
    return HPCLA_OK;
}

extern "C" int hpcla_cg(hpcla_spmv* op, const void* d_b, void* d_x, void* d_work, int iters, double* rr_history_out, void* stream_) {
    NvtxRange nvtx_range("hpcla_cg");
    if (!op || !d_b || !d_x || !d_work || iters < 0) return fail(HPCLA_ERR_ARG, "hpcla_cg: bad arguments");
    hpcla_ctx* ctx = op->ctx;
    const int dtype = op->csr->dtype;
    if (dtype == HPCLA_C128) return fail(HPCLA_ERR_ARG, "hpcla_cg: real element types only");
    if (ctx->group && ctx->nranks > 1) return fail(HPCLA_ERR_STATE, "hpcla_cg: needs an NCCL world or a single rank");
    if (op->csr->nrows != op->n_x_local) return fail(HPCLA_ERR_ARG, "hpcla_cg: A must be square with x partitioned like its rows");
    int rc = set_device(ctx);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const i64 n = op->csr->nrows;
    const size_t es = dtype_size(dtype);
    char* r = (char*)d_work;
    char* p = r + (size_t)n * es;
    char* q = p + (size_t)n * es;
    if (op->cg_cap < iters) {  // rr_k at [2k], pq at the tail; kept by the operator, grown on demand
        CU_TRY(cudaStreamSynchronize(stream));
        cudaFree(op->d_cg_scalars);
        op->d_cg_scalars = nullptr;
        op->cg_cap = 0;
        const int cap = std::max(iters, 64);
        CU_TRY(cudaMalloc(&op->d_cg_scalars, sizeof(double) * (size_t)(2 * (cap + 1) + 2)));
        op->cg_cap = cap;
    }
    double* d_s = op->d_cg_scalars;
    double* d_pq = d_s + 2 * (op->cg_cap + 1);
    // p.q rides on the multiply when every row goes through the row-walk kernel (stencil-like matrices): the multiply
    // leaves one partial per CTA, a one-CTA kernel adds them in a fixed order — p and q are not read a second time
    const i64 n_partials = (i64)op->n_list[0][0] + op->n_list[2][0] + op->n_list[0][1];
    // (not with compact tiles: there the walk runs out of shared memory and the extra global load of p per row costs more
    // than the separate dot kernel saves — measured 570 vs 488 us per iteration on Poisson 256^3, profiles/r2e_*)
    const bool fused = op->n_list[1][0] + op->n_list[1][1] == 0 && op->n_list[2][0] == 0 && op->csr->nlong == 0 && op->csr->flat.n_chunks == 0 && op->x_in_place &&
                       op->own_src0 == 1 && n_partials > 0 && !getenv("HPCLA_CG_UNFUSED");
    if (fused && !op->d_dot_partials) CU_TRY(cudaMalloc(&op->d_dot_partials, sizeof(double) * (size_t)n_partials));
    // HPCLA_CG_GRAPH=1: the whole loop (iters x {multiply + halo, p.q, Allreduce, x/r update, Allreduce, p update}) as ONE
    // CUDA graph, re-used while the buffers and the iteration count stay the same: one launch instead of ~12 * iters
    // driver calls (the regime where the host, not the device, paces the loop).
    static const bool want_graph = [] { const char* e = getenv("HPCLA_CG_GRAPH"); return e && e[0] == '1'; }();
    const bool same = op->cg_graph && op->cg_key_b == d_b && op->cg_key_x == d_x && op->cg_key_w == d_work && op->cg_key_iters == iters && op->cg_key_fused == fused;
    if (want_graph && same) {
        CU_TRY(cudaGraphLaunch(op->cg_graph, stream));
        op->launches += op->cg_graph_launches;
    } else {
        const i64 before = op->launches;
        cudaStream_t es = stream;  // the stream the loop is enqueued on: the caller's, or the library's capture stream
        if (want_graph) {
            if (op->cg_graph) cudaGraphExecDestroy(op->cg_graph);
            op->cg_graph = nullptr;
            if (!ctx->cap_stream) CU_TRY(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
            es = ctx->cap_stream;
            CU_TRY(cudaStreamBeginCapture(es, cudaStreamCaptureModeRelaxed));
        }
        const bool tl = op->timeline;
        if (want_graph) op->timeline = false;
        rc = cg_enqueue(op, d_b, d_x, r, p, q, iters, fused, n_partials, d_s, d_pq, es);
        op->timeline = tl;
        if (want_graph) {
            cudaGraph_t g = nullptr;
            cudaError_t ce = cudaStreamEndCapture(es, &g);
            op->phase = 0;
            if (rc || ce != cudaSuccess || !g) {
                if (g) cudaGraphDestroy(g);
                cudaGetLastError();
                return rc ? rc : fail(HPCLA_ERR_CUDA, "hpcla_cg: graph capture failed: %s", cudaGetErrorString(ce));
            }
            ce = cudaGraphInstantiate(&op->cg_graph, g, 0);
            cudaGraphDestroy(g);
            if (ce != cudaSuccess) return fail(HPCLA_ERR_CUDA, "hpcla_cg: cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
            op->cg_key_b = d_b, op->cg_key_x = d_x, op->cg_key_w = d_work, op->cg_key_iters = iters, op->cg_key_fused = fused;
            op->cg_graph_launches = op->launches - before;
            CU_TRY(cudaGraphLaunch(op->cg_graph, stream));
        } else if (rc) {
            return rc;
        }
    }
    std::vector<double>& h = op->cg_host;
    h.resize((size_t)(2 * (iters + 1)));
    CU_TRY(cudaMemcpyAsync(h.data(), d_s, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaStreamSynchronize(stream));
    if (rr_history_out)
        for (int k = 0; k < iters; ++k) rr_history_out[k] = h[(size_t)(2 * (k + 1))];
    return HPCLA_OK;
}
