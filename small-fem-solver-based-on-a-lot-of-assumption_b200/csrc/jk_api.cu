// jk_api.cu -- C ABI of libjacket_b200.so (see include/jacket_b200.h).
// Host-side orchestration only: SoA upload, ordering (reverse Cuthill-McKee), CSR builds for the
// deterministic gathers, buffer management, kernel launches and CUDA-event stage timers.
// There is no CPU compute path: every jk_* numeric entry point launches kernels or fails.
#include "../../include/jacket_b200.h"

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <string>
#include <vector>

#include "jk_chol.cuh"
#include "jk_chol_cluster.cuh"
#include "jk_common.cuh"
#include "jk_fem.cuh"
#include "jk_morison.cuh"
#include "jk_sweep.cuh"

using namespace jk;

static thread_local std::string g_err;

// Run-time options (jk_set_option / jk_get_option).  The library never reads the environment.
enum JkOpt { OPT_DEBUG_MORISON_SMEM_PAD, OPT_DEBUG_FUSE_MODE, OPT_START_GATE, OPT_START_GATE2, OPT_POST_OVERLAP, OPT_EARLY_TOTALS, OPT_FACTOR_SPLIT, OPT_SPLIT_PCT, OPT_TWO_CHAINS,
             OPT_LEVEL_REGROUP, OPT_SUPPORT_ROOTED_RCM, OPT_TMA_SWEEP, OPT_BLOCKED_INVERSE, OPT_PROFILE_CHOL, OPT_PROFILE_SWEEP,
             OPT_DEBUG_FACTOR_DELAY, OPT_SWEEP_SLAB, OPT_CUDA_GRAPH, OPT_FUSED_LOADS, OPT_GATHER_BLOCKS, OPT_GATHER_ROWS, OPT_COUNT };
struct JkOptDesc { const char* key; int def, lo, hi; };
static const JkOptDesc g_opts[OPT_COUNT] = {
    {"debug_morison_smem_pad", 0, 0, 160},  // experiment: extra KB of dynamic shared memory per Morison block (occupancy probe)
    {"debug_fuse_mode", 0, 0, 2},           // experiment: 1 = fused Morison kernel without finalisation, 2 = without waiting (wrong results; timing only)
    {"start_gate", 1, 0, 1},            // main stream waits until the factor clusters are resident (asynchronous factorisation)
    {"start_gate2", 1, 0, 1},           // first forward sweep parts wait until the second factor segment is resident
    {"post_overlap", 1, 0, 1},          // member post of first-chain chunks beside the second chain's backward sweep
    {"early_totals", 1, 0, 1},          // Morison columns of the table reduced on a side stream behind the Morison kernel
    {"factor_split", 1, 0, 1},          // two factor segments, forward sweeps start on the first      [next jk_set_supports]
    {"split_pct", 70, 10, 95},          // split point in % of a chain's columns                         [next jk_set_supports]
    {"two_chains", 1, 0, 1},            // two-sided elimination                                         [next jk_set_supports]
    {"level_regroup", 1, 0, 1},         // BFS levels re-sorted by node degree (block-cost ordering)      [next jk_set_supports]
    {"support_rooted_rcm", 1, 0, 1},    // RCM rooted at the support-adjacent node set as a candidate    [next jk_set_supports]
    {"tma_sweep", 1, 0, 1},             // TMA / mbarrier sweep pipeline (0: cp.async slab sweep)        [next jk_set_supports]
    {"blocked_inverse", 1, 0, 1},       // diagonal-tile inverses from the factor's 8x8 block inverses
    {"profile_chol", 0, 0, 1},          // debug: clock stamps of the cluster factorisation to stderr
    {"profile_sweep", 0, 0, 2},         // debug: clock sums of the sweep warps to stderr (2: plus per-item clock stamps of two warps)
    {"debug_factor_delay", 0, 0, 2000000000},   // test aid: spin this many clocks in front of the factorisation
    {"sweep_slab", 0, 0, 32},           // right-hand sides per sweep CTA: 0 = auto (fill the SMs), 8, 16 or 32
    {"cuda_graph", 1, 0, 1},            // replay the resident scan (jk_phase_scan_dev) as a captured CUDA graph
    {"fused_loads", 0, 0, 1},           // Morison kernel lumps member end forces into nodal loads itself (no member-force round trip;
                                        // measured slower than the two-kernel path at c4: 1.81 vs 1.70 ms, see jk_morison.cuh)
    {"gather_blocks", 1, 1, 16},        // phase blocks of the Morison stage: the load gather of block b (HBM-bound, side stream) runs beside the
                                        // Morison kernel of block b + 1 (FP64-bound); 1 = one Morison launch, then one gather
    {"gather_rows", 0, 0, 65535},       // grid.y of the overlapped gather launches (0: one block row per node)
};

struct jk_handle_s {
    int opt[OPT_COUNT];
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t stream2 = nullptr;        // side stream: the factorisation runs here, concurrently with the Morison stage
    cudaStream_t stream3 = nullptr;        // load gather of one phase block while the Morison kernel works on the next
    cudaEvent_t ev_post_fork = nullptr, ev_post_join = nullptr;   // node-level post kernels run beside the member post on the side stream
    // pinned staging for the small per-scan host transfers (static load in, table + critical index + pivot flag out): one
    // asynchronous copy each and a single synchronisation per scan instead of pageable copies with their own syncs
    unsigned char* h_pin = nullptr; size_t pin_bytes = 0; cudaEvent_t ev_pin = nullptr; bool pin_busy = false;
    double* h_pin_t = nullptr; size_t pin_t_elems = 0; cudaEvent_t ev_pin_t = nullptr; bool pin_t_busy = false;   // phase times in
    // start gate of the factor clusters (k_band_chol_cluster): device counter + cuStreamWaitValue32 on the main stream
    unsigned* d_started = nullptr; unsigned started_target = 0;
    // early member post: chunks whose members only touch chain-0 / separator nodes are post-processed on a side stream
    // while the second chain's backward sweep runs (HBM-bound work on the SMs the sweep leaves idle)
    int* d_post_chunks = nullptr; int n_post_early = 0, n_post_late = 0, n_sm = 0;
    cudaEvent_t ev_bwd0 = nullptr, ev_post_early = nullptr, ev_mor = nullptr, ev_tot = nullptr;
    cudaEvent_t ev_blk[16] = {}, ev_gather = nullptr;      // Morison phase blocks -> overlapped load gathers (option gather_blocks)
    unsigned gate2_target = 0; bool gate2_armed = false;   // second factor segment resident (awaited before the first forward parts)
    CUresult (*wait_value32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    cudaEvent_t ev_seg1 = nullptr, ev_fwd1 = nullptr;   // first factor segment done / its forward tile streams built
    bool split_factor = false;
    cudaEvent_t ev_fork = nullptr, ev_factor = nullptr, ev_factor_bwd = nullptr;   // ev_factor: forward sweeps may start; ev_factor_bwd: backward tile streams built too
    bool factor_inflight = false;
    bool ev_bwd_external = false;   // ev_factor_bwd was recorded outside a graph capture and has not been joined by the main stream yet
    std::string err;
    int64_t launches = 0;

    // geometry
    int Nn = 0, M = 0, nsec = 0;
    std::vector<double> h_xyz;
    std::vector<int> h_conn, h_sec;
    double *d_xyz = nullptr, *d_secp = nullptr, *d_mc = nullptr, *d_Ke = nullptr, *d_Kl = nullptr, *d_pk = nullptr;
    int *d_conn = nullptr, *d_sec = nullptr, *d_adj_ptr = nullptr, *d_adj = nullptr;
    std::vector<int> h_adj_ptr, h_adj;

    // supports / ordering / storage
    bool have_supports = false;
    int n_fixed = 0, n_free_nodes = 0, n_free = 0, n_pad = 0, NT = 0, bw = 0, solver = 0, hb = 0;
    std::vector<int> h_node2slot, h_free_nodes, h_fixed;
    int *d_node2slot = nullptr, *d_fixed_nodes = nullptr, *d_free_nodes = nullptr;
    // The free-free system is stored as one or two tile-banded "chains".  Two chains = the band is split at a
    // separator S in the middle of the ordering: chain 0 = [A; S] in order, chain 1 = [B; S] in REVERSED order, so
    // both eliminate towards S and the sequential pivot chain of the factorisation is halved (no extra fill).
    struct Chain {
        int NT = 0, kS = 0, bw = 0, hb = 0;   // tile rows, first separator tile row (== NT: none), tile / DOF half-bandwidth
        int row0 = 0, n_rows = 0;             // first row in the slab, real (unpadded) rows
        KBlock* d_blocks = nullptr; int2* d_contrib = nullptr; int nblocks = 0;
        double *d_tiles = nullptr, *d_Linv = nullptr, *d_dinv = nullptr;
        size_t tiles_elems = 0;
        // TMA sweep programs (jk_sweep.cuh): [0] forward, [1] backward.  Item list (host-built) + tile stream (built
        // from L after every factorisation by k_sweep_build)
        struct Sweep { int n_items = 0, pre_row = 0, npre = 0, ktop = 0; uint4* d_prog = nullptr; double* d_stream = nullptr;
                       // forward programs can run as two launches: items [0, n_split) cover tile rows < k_split and only need
                       // the factor's columns < k_split, so they start while the factorisation finishes the rest
                       int n_split = 0, k_split = 0, npre2 = 0, xphase2 = 0; } sw[2];
    } ch[2];
    bool tma_sweep = false;   // narrow band: sweeps run as the TMA / mbarrier pipeline, otherwise the cp.async slab sweep
    int n_chains = 1, nS_nodes = 0;
    long long nnz_env_min = 0;   // non-zeros of L (lower triangle incl. diagonal) for the candidate ordering with the smallest envelope
    int factor_path = 0;   // 0 = auto (cluster kernel for narrow bands), 1 = per-column launches
    int* d_info = nullptr;          // pivot flag of the factorisation in flight
    int* d_info_sticky = nullptr;   // first non-zero pivot flag since it was last reported (survives re-assembly in a resident loop)
    bool assembled = false, factored = false;
    double E = 0, G = 0;

    // fused load lumping (jk_morison.cuh, LoadFuse): member processing order, chunk-local nodes, finalisation lists
    int fuse_n_pairs = 0, fuse_n_fin = 0, fuse_n_chunk = 0;
    int *d_fuse_order = nullptr, *d_fuse_pair_base = nullptr, *d_fuse_fin_ptr = nullptr, *d_fuse_fin_node = nullptr,
        *d_fuse_fin_rowptr = nullptr, *d_fuse_fin_rows = nullptr, *d_fuse_dep_lo = nullptr, *d_fuse_sync = nullptr, *d_fuse_fin_of_node = nullptr;
    unsigned* d_fuse_ends = nullptr;
    double* d_part = nullptr;          // [n_pairs][3][ldP] chunk-local nodal sums of the last fused scan
    bool last_fused = false;
    // loads / wave / morison
    double* d_Fstatic = nullptr;
    bool have_wave = false, have_morison = false, gp_valid = false;
    WaveAiry wv{};
    double rho = 0, Cd = 0, Cm = 0;
    int ng = 0;
    double *d_gsw = nullptr, *d_gp = nullptr;
    int wave_kind = 0;         // 0 = Airy closed form, 1 = Fourier series
    int n_harm = 0;
    double* d_four = nullptr;  // E[Nh], B[Nh], 1/cosh(j k d)[Nh]
    double* d_states = nullptr; long long* d_state_crit = nullptr; int cap_states = 0;   // ensemble: [5][S] parameters
    size_t gp_elems = 0;
    StressPts sp{};

    // scan buffers (capacity cap_ldP phases)
    int cap_ldP = 0;
    bool cap_details = false;
    double *d_t = nullptr, *d_trig = nullptr, *d_Fm = nullptr, *d_X = nullptr, *d_Z = nullptr, *d_Ffix = nullptr, *d_rows = nullptr;
    double *d_totpart = nullptr, *d_part_util = nullptr, *d_part_vm = nullptr, *d_part_disp = nullptr, *d_react = nullptr;
    double *d_table = nullptr, *d_details = nullptr, *d_Fload = nullptr, *d_argval = nullptr, *d_tmp = nullptr, *d_res = nullptr;
    size_t fload_elems = 0, tmp_elems = 0;
    int *d_part_mem = nullptr, *d_part_node = nullptr;
    long long* d_argidx = nullptr;
    int lastP = 0, last_ldP = 0;
    int sweep_slab_last = SLAB;   // right-hand sides per sweep CTA of the last solve
    bool last_morison = false, last_fem = false, last_fdir = false;
    // captured CUDA graph of one resident step (jk_step_dev / jk_step): valid while nothing it bakes in has changed
    cudaGraphExec_t graph_exec = nullptr;
    long long graph_epoch = 0, graph_epoch_built = -1, graph_launches = 0;      // epoch: bumped by every call that changes kernel arguments or buffers
    int graph_P = 0; double graph_E = 0, graph_G = 0, graph_fy = 0;
    unsigned graph_started_target = 0; bool graph_gate2_armed = false;
    bool capturing = false, graph_unsupported = false, graph_fused = false;
    double last_fy = 355.0;

    cudaEvent_t ev0[JK_NTIMERS], ev1[JK_NTIMERS];
    bool ev_set[JK_NTIMERS];
};

#define JK_FAIL(h, code, ...)                                    \
    do {                                                         \
        char _b[512];                                            \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                   \
        if (h) (h)->err = _b; else g_err = _b;                   \
        return (code);                                           \
    } while (0)

#define CUDA_TRY(h, expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) JK_FAIL(h, JK_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define LAUNCH_CHECK(h)                                                                          \
    do {                                                                                         \
        (h)->launches++;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) JK_FAIL(h, JK_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

template <typename T>
static cudaError_t dev_alloc(T** p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
}
template <typename T>
static void dev_free(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
// pinned staging buffer of at least n bytes, free for reuse (the previous asynchronous copy out of it has finished)
static unsigned char* pin_acquire(jk_handle_t h, size_t n) {
    if (h->pin_busy) { cudaEventSynchronize(h->ev_pin); h->pin_busy = false; }
    if (n > h->pin_bytes) {
        if (h->h_pin) { cudaFreeHost(h->h_pin); h->h_pin = nullptr; h->pin_bytes = 0; }
        if (cudaMallocHost((void**)&h->h_pin, n) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        h->pin_bytes = n;
    }
    return h->h_pin;
}
// stage timers: CUDA events on the launching stream; not recorded into a captured graph (a replay cannot be timed by them)
static void tic(jk_handle_t h, int id, cudaStream_t s = nullptr) { if (!h->capturing) cudaEventRecord(h->ev0[id], s ? s : h->stream); }
static void toc(jk_handle_t h, int id, cudaStream_t s = nullptr) { if (!h->capturing) { cudaEventRecord(h->ev1[id], s ? s : h->stream); h->ev_set[id] = true; } }

extern "C" int jk_version(void) { return 100; }

extern "C" const char* jk_last_error(jk_handle_t h) { return h ? h->err.c_str() : g_err.c_str(); }
extern "C" int64_t jk_launch_count(jk_handle_t h) { return h ? h->launches : 0; }
extern "C" void* jk_stream(jk_handle_t h) { return h ? (void*)h->stream : nullptr; }
extern "C" void* jk_table_dev(jk_handle_t h) { return h ? (void*)h->d_table : nullptr; }
extern "C" void* jk_critical_value_dev(jk_handle_t h) { return h ? (void*)h->d_argval : nullptr; }
extern "C" void* jk_critical_index_dev(jk_handle_t h) { return h ? (void*)h->d_argidx : nullptr; }

static int find_option(const char* key) {
    if (!key) return -1;
    for (int i = 0; i < OPT_COUNT; ++i) if (strcmp(key, g_opts[i].key) == 0) return i;
    return -1;
}
extern "C" int jk_set_option(jk_handle_t h, const char* key, int value) {
    if (!h) return JK_EINVAL;
    const int i = find_option(key);
    if (i < 0) JK_FAIL(h, JK_EINVAL, "jk_set_option: unknown option '%s'", key ? key : "(null)");
    if (value < g_opts[i].lo || value > g_opts[i].hi) JK_FAIL(h, JK_EINVAL, "jk_set_option: %s must be in %d..%d (got %d)", key, g_opts[i].lo, g_opts[i].hi, value);
    if (i == OPT_SWEEP_SLAB && value != 0 && value != 8 && value != 16 && value != 32) JK_FAIL(h, JK_EINVAL, "jk_set_option: sweep_slab must be 0 (auto), 8, 16 or 32");
    h->opt[i] = value;
    h->graph_epoch++;
    return JK_OK;
}
extern "C" int jk_get_option(jk_handle_t h, const char* key, int* value) {
    if (!h || !value) return JK_EINVAL;
    const int i = find_option(key);
    if (i < 0) JK_FAIL(h, JK_EINVAL, "jk_get_option: unknown option '%s'", key ? key : "(null)");
    *value = h->opt[i];
    return JK_OK;
}
extern "C" int jk_option_count(void) { return OPT_COUNT; }
extern "C" const char* jk_option_name(int i) { return (i >= 0 && i < OPT_COUNT) ? g_opts[i].key : nullptr; }

// ------------------------------------------------------------------------------------------------
extern "C" int jk_create(int device, void* stream, int n_nodes, const double* xyz, int n_members,
                         const int32_t* conn, const int32_t* sec_id, int n_sec, const double* sec_props,
                         jk_handle_t* out) {
    if (!out) JK_FAIL((jk_handle_t)nullptr, JK_EINVAL, "jk_create: out is NULL");
    *out = nullptr;
    if (n_nodes <= 0 || n_members <= 0 || n_sec <= 0 || !xyz || !conn || !sec_id || !sec_props)
        JK_FAIL((jk_handle_t)nullptr, JK_EINVAL, "jk_create: empty or NULL geometry (n_nodes=%d n_members=%d n_sec=%d)", n_nodes, n_members, n_sec);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        JK_FAIL((jk_handle_t)nullptr, JK_ENODEVICE, "jk_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) JK_FAIL((jk_handle_t)nullptr, JK_EINVAL, "jk_create: device %d out of range (%d devices)", device, ndev);
    for (int m = 0; m < n_members; ++m) {
        int a = conn[2 * m], b = conn[2 * m + 1];
        if (a < 0 || a >= n_nodes || b < 0 || b >= n_nodes || a == b)
            JK_FAIL((jk_handle_t)nullptr, JK_EINVAL, "jk_create: member %d has invalid end nodes (%d, %d)", m, a, b);
        if (sec_id[m] < 0 || sec_id[m] >= n_sec) JK_FAIL((jk_handle_t)nullptr, JK_EINVAL, "jk_create: member %d has invalid section %d", m, sec_id[m]);
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
        JK_FAIL((jk_handle_t)nullptr, JK_ENODEVICE, "jk_create: device %d is not sm_100-class (this library is built for sm_100a only)", device);

    jk_handle_t h = new jk_handle_s();
    h->device = device;
    cudaSetDevice(device);
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device);
    for (int i = 0; i < OPT_COUNT; ++i) h->opt[i] = g_opts[i].def;
    if (stream) { h->stream = (cudaStream_t)stream; h->own_stream = false; }
    else { if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; JK_FAIL((jk_handle_t)nullptr, JK_ECUDA, "jk_create: cudaStreamCreate failed"); } h->own_stream = true; }
    for (int i = 0; i < JK_NTIMERS; ++i) { cudaEventCreate(&h->ev0[i]); cudaEventCreate(&h->ev1[i]); h->ev_set[i] = false; }
    { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, hi); }
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_factor, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_factor_bwd, cudaEventDisableTiming);
    { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, hi); }
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn != nullptr
            && cudaMalloc((void**)&h->d_started, sizeof(unsigned)) == cudaSuccess && cudaMemset(h->d_started, 0, sizeof(unsigned)) == cudaSuccess)
            h->wait_value32 = reinterpret_cast<CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int)>(fn);
        cudaGetLastError();
    }
    cudaEventCreateWithFlags(&h->ev_pin, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_pin_t, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_seg1, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_fwd1, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_bwd0, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_mor, cudaEventDisableTiming);
    for (auto& e : h->ev_blk) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_gather, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_tot, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_post_early, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_post_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_post_join, cudaEventDisableTiming);
    h->Nn = n_nodes; h->M = n_members; h->nsec = n_sec;
    h->h_xyz.assign(xyz, xyz + 3 * (size_t)n_nodes);
    h->h_conn.assign(conn, conn + 2 * (size_t)n_members);
    h->h_sec.assign(sec_id, sec_id + n_members);
    // node -> (member, end) adjacency in member order
    h->h_adj_ptr.assign(n_nodes + 1, 0);
    for (int m = 0; m < n_members; ++m) { h->h_adj_ptr[conn[2 * m] + 1]++; h->h_adj_ptr[conn[2 * m + 1] + 1]++; }
    for (int i = 0; i < n_nodes; ++i) h->h_adj_ptr[i + 1] += h->h_adj_ptr[i];
    h->h_adj.resize(2 * (size_t)n_members);
    { std::vector<int> fill(h->h_adj_ptr.begin(), h->h_adj_ptr.end() - 1);
      for (int m = 0; m < n_members; ++m) for (int e = 0; e < 2; ++e) h->h_adj[fill[conn[2 * m + e]]++] = (m << 1) | e; }
    for (int i = 0; i < 8; ++i) { double rad = (45.0 * i) * (M_PI / 180.0); h->sp.cs[i] = cos(rad); h->sp.sn[i] = sin(rad); }

    *out = h;   // from here on errors are reported on the handle; caller destroys it
    CUDA_TRY(h, dev_alloc(&h->d_xyz, 3 * (size_t)n_nodes));
    CUDA_TRY(h, dev_alloc(&h->d_conn, 2 * (size_t)n_members));
    CUDA_TRY(h, dev_alloc(&h->d_sec, (size_t)n_members));
    CUDA_TRY(h, dev_alloc(&h->d_secp, (size_t)n_sec * JK_SEC_NPROP));
    CUDA_TRY(h, dev_alloc(&h->d_mc, (size_t)n_members * MC_STRIDE));
    CUDA_TRY(h, dev_alloc(&h->d_pk, (size_t)n_members * PK_STRIDE));
    CUDA_TRY(h, dev_alloc(&h->d_Ke, (size_t)n_members * 144));
    CUDA_TRY(h, dev_alloc(&h->d_Kl, (size_t)n_members * 144));
    CUDA_TRY(h, dev_alloc(&h->d_adj_ptr, (size_t)n_nodes + 1));
    CUDA_TRY(h, dev_alloc(&h->d_adj, 2 * (size_t)n_members));
    CUDA_TRY(h, dev_alloc(&h->d_Fstatic, 6 * (size_t)n_nodes));
    CUDA_TRY(h, dev_alloc(&h->d_info, 1));
    CUDA_TRY(h, dev_alloc(&h->d_info_sticky, 1));
    CUDA_TRY(h, cudaMemsetAsync(h->d_info_sticky, 0, sizeof(int), h->stream));
    CUDA_TRY(h, dev_alloc(&h->d_argval, 1));
    CUDA_TRY(h, dev_alloc(&h->d_argidx, 1));
    CUDA_TRY(h, dev_alloc(&h->d_res, 2));
    cudaStream_t s = h->stream;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_xyz, xyz, 3 * (size_t)n_nodes * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_conn, conn, 2 * (size_t)n_members * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_sec, sec_id, (size_t)n_members * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_secp, sec_props, (size_t)n_sec * JK_SEC_NPROP * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_adj_ptr, h->h_adj_ptr.data(), ((size_t)n_nodes + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_adj, h->h_adj.data(), 2 * (size_t)n_members * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemsetAsync(h->d_Fstatic, 0, 6 * (size_t)n_nodes * sizeof(double), s));
    // opt in to large dynamic shared memory once
    CUDA_TRY(h, cudaFuncSetAttribute(k_trailing_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPDATE_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_panel_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_tile_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INVERSE_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_tile_inverse_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INVERSE_BLOCKED_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_band_chol_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_CLUSTER_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_slab_sweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_slab_sweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_sweep<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_sweep<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_sweep<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_sweep<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_sweep<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_sweep<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(k_sweep_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWB_SMEM));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return JK_OK;
}

extern "C" int jk_destroy(jk_handle_t h) {
    if (!h) return JK_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    dev_free(h->d_xyz); dev_free(h->d_secp); dev_free(h->d_mc); dev_free(h->d_Ke); dev_free(h->d_Kl); dev_free(h->d_pk);
    dev_free(h->d_conn); dev_free(h->d_sec); dev_free(h->d_adj_ptr); dev_free(h->d_adj);
    dev_free(h->d_node2slot); dev_free(h->d_fixed_nodes); dev_free(h->d_free_nodes);
    for (auto& c : h->ch) { dev_free(c.d_blocks); dev_free(c.d_contrib); dev_free(c.d_tiles); dev_free(c.d_Linv); dev_free(c.d_dinv);
                            for (auto& w : c.sw) { dev_free(w.d_prog); dev_free(w.d_stream); } }
    dev_free(h->d_info); dev_free(h->d_info_sticky);
    dev_free(h->d_fuse_order); dev_free(h->d_fuse_pair_base); dev_free(h->d_fuse_fin_ptr); dev_free(h->d_fuse_fin_node);
    dev_free(h->d_fuse_fin_rowptr); dev_free(h->d_fuse_fin_rows); dev_free(h->d_fuse_dep_lo); dev_free(h->d_fuse_sync);
    dev_free(h->d_fuse_ends); dev_free(h->d_part); dev_free(h->d_fuse_fin_of_node);
    dev_free(h->d_Fstatic); dev_free(h->d_gsw); dev_free(h->d_gp); dev_free(h->d_four); dev_free(h->d_states); dev_free(h->d_state_crit);
    dev_free(h->d_t); dev_free(h->d_trig); dev_free(h->d_Fm); dev_free(h->d_X); dev_free(h->d_Z); dev_free(h->d_Ffix); dev_free(h->d_rows);
    dev_free(h->d_totpart); dev_free(h->d_part_util); dev_free(h->d_part_vm); dev_free(h->d_part_disp); dev_free(h->d_react);
    dev_free(h->d_table); dev_free(h->d_details); dev_free(h->d_Fload); dev_free(h->d_argval); dev_free(h->d_tmp); dev_free(h->d_res);
    dev_free(h->d_part_mem); dev_free(h->d_part_node); dev_free(h->d_argidx);
    for (int i = 0; i < JK_NTIMERS; ++i) { cudaEventDestroy(h->ev0[i]); cudaEventDestroy(h->ev1[i]); }
    if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_factor) cudaEventDestroy(h->ev_factor);
    if (h->ev_factor_bwd) cudaEventDestroy(h->ev_factor_bwd);
    if (h->stream3) { cudaStreamSynchronize(h->stream3); cudaStreamDestroy(h->stream3); }
    if (h->h_pin) cudaFreeHost(h->h_pin);
    if (h->h_pin_t) cudaFreeHost(h->h_pin_t);
    if (h->ev_pin_t) cudaEventDestroy(h->ev_pin_t);
    if (h->ev_pin) cudaEventDestroy(h->ev_pin);
    if (h->d_started) cudaFree(h->d_started);
    if (h->ev_seg1) cudaEventDestroy(h->ev_seg1);
    if (h->ev_fwd1) cudaEventDestroy(h->ev_fwd1);
    if (h->ev_bwd0) cudaEventDestroy(h->ev_bwd0);
    if (h->ev_mor) cudaEventDestroy(h->ev_mor);
    for (auto& e : h->ev_blk) if (e) cudaEventDestroy(e);
    if (h->ev_gather) cudaEventDestroy(h->ev_gather);
    if (h->ev_tot) cudaEventDestroy(h->ev_tot);
    if (h->ev_post_early) cudaEventDestroy(h->ev_post_early);
    dev_free(h->d_post_chunks);
    if (h->ev_post_fork) cudaEventDestroy(h->ev_post_fork);
    if (h->ev_post_join) cudaEventDestroy(h->ev_post_join);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return JK_OK;
}

// ------------------------------------------------------------------------------------------------
// reverse Cuthill-McKee over the free nodes (graph = members with both ends free)
// ------------------------------------------------------------------------------------------------
// member graph restricted to the free nodes
static void free_graph(int Nn, const std::vector<int>& conn, const std::vector<char>& is_fixed, std::vector<std::vector<int>>& nb) {
    int M = (int)conn.size() / 2;
    nb.assign(Nn, {});
    for (int m = 0; m < M; ++m) {
        int a = conn[2 * m], b = conn[2 * m + 1];
        if (is_fixed[a] || is_fixed[b]) continue;
        nb[a].push_back(b); nb[b].push_back(a);
    }
    for (auto& v : nb) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
}

// Cuthill-McKee from a root SET (in the given order) over the nodes with allowed[v] != 0 that are not yet visited;
// neighbours are queued by increasing degree.  Appends to out and marks visited.
static void cm_from(const std::vector<std::vector<int>>& nb, const std::vector<int>& roots, const std::vector<char>& allowed,
                    std::vector<char>& visited, std::vector<int>& out, std::vector<int>* level = nullptr) {
    std::queue<int> q;
    for (int r : roots) if (allowed[r] && !visited[r]) { visited[r] = 1; q.push(r); if (level) (*level)[r] = 0; }
    std::vector<int> nx;
    while (!q.empty()) {
        int u = q.front(); q.pop(); out.push_back(u);
        nx.clear();
        for (int v : nb[u]) if (allowed[v] && !visited[v]) { visited[v] = 1; nx.push_back(v); if (level) (*level)[v] = (*level)[u] + 1; }
        std::sort(nx.begin(), nx.end(), [&](int a, int b) { return nb[a].size() != nb[b].size() ? nb[a].size() < nb[b].size() : a < b; });
        for (int v : nx) q.push(v);
    }
}

// classic reverse Cuthill-McKee, one pseudo-peripheral root (George-Liu) per connected component
static void rcm_order(int Nn, const std::vector<std::vector<int>>& nb, const std::vector<char>& is_fixed, std::vector<int>& order) {
    std::vector<char> visited(Nn, 0), allowed(Nn, 0);
    for (int i = 0; i < Nn; ++i) allowed[i] = !is_fixed[i];
    std::vector<int> level(Nn, -1);
    order.clear();
    auto bfs_levels = [&](int root, std::vector<int>& comp) {   // returns eccentricity; comp = BFS order
        comp.clear();
        std::queue<int> q;
        level[root] = 0; q.push(root);
        int ecc = 0;
        while (!q.empty()) {
            int u = q.front(); q.pop(); comp.push_back(u);
            ecc = std::max(ecc, level[u]);
            for (int v : nb[u]) if (level[v] < 0) { level[v] = level[u] + 1; q.push(v); }
        }
        return ecc;
    };
    std::vector<int> comp, comp2;
    for (int seed = 0; seed < Nn; ++seed) {
        if (is_fixed[seed] || visited[seed]) continue;
        int root = seed;
        int ecc = bfs_levels(root, comp);
        for (int it = 0; it < 8; ++it) {
            int best = -1;
            for (int u : comp) if (level[u] == ecc && (best < 0 || nb[u].size() < nb[best].size())) best = u;
            for (int u : comp) level[u] = -1;
            int ecc2 = bfs_levels(best, comp2);
            if (ecc2 > ecc) { root = best; ecc = ecc2; comp.swap(comp2); }
            else { for (int u : comp2) level[u] = -1; bfs_levels(root, comp); break; }
        }
        for (int u : comp) level[u] = -1;
        std::vector<int> cm;
        cm_from(nb, {root}, allowed, visited, cm);
        order.insert(order.end(), cm.rbegin(), cm.rend());
    }
}

// reverse Cuthill-McKee rooted at the SET of free nodes attached to a support: for a tower the level sets are then the
// horizontal frames, which gives the minimum bandwidth (a single root makes slanted, wider fronts)
static void rcm_order_from_supports(int Nn, const std::vector<int>& conn, const std::vector<std::vector<int>>& nb,
                                    const std::vector<char>& is_fixed, std::vector<int>& order, std::vector<int>* cm_out = nullptr,
                                    std::vector<int>* level_out = nullptr) {
    std::vector<char> visited(Nn, 0), allowed(Nn, 0), is_root(Nn, 0);
    for (int i = 0; i < Nn; ++i) allowed[i] = !is_fixed[i];
    int M = (int)conn.size() / 2;
    for (int m = 0; m < M; ++m) {
        int a = conn[2 * m], b = conn[2 * m + 1];
        if (is_fixed[a] && !is_fixed[b]) is_root[b] = 1;
        if (is_fixed[b] && !is_fixed[a]) is_root[a] = 1;
    }
    std::vector<int> roots;
    for (int i = 0; i < Nn; ++i) if (is_root[i]) roots.push_back(i);
    std::sort(roots.begin(), roots.end(), [&](int a, int b) { return nb[a].size() != nb[b].size() ? nb[a].size() < nb[b].size() : a < b; });
    std::vector<int> cm, level(Nn, 0);
    cm_from(nb, roots, allowed, visited, cm, &level);
    for (int i = 0; i < Nn; ++i) if (allowed[i] && !visited[i]) cm_from(nb, {i}, allowed, visited, cm, &level);   // parts not tied to a support
    order.assign(cm.rbegin(), cm.rend());
    if (cm_out) *cm_out = cm;
    if (level_out) *level_out = level;
}

// A Cuthill-McKee order with its BFS levels kept contiguous but every level re-sorted by node degree (ascending or
// descending, ties in CM order).  In a frame the degree separates node families (jacket: brace hinges 4, leg joints 8)
// whose rows reach back differently far; grouping them makes the 8-row blocks of the sweep masks homogeneous.
static std::vector<int> regroup_levels_by_degree(const std::vector<int>& cm, const std::vector<int>& level,
                                                 const std::vector<std::vector<int>>& nb, bool ascending, bool reverse_ties = false) {
    std::vector<int> idx(cm.size());
    for (size_t i = 0; i < cm.size(); ++i) idx[i] = (int)i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
        const int la = level[cm[a]], lb = level[cm[b]];
        if (la != lb) return la < lb;
        const size_t da = nb[cm[a]].size(), db = nb[cm[b]].size();
        if (da != db) return ascending ? da < db : da > db;
        return reverse_ties ? a > b : a < b;
    });
    std::vector<int> out(cm.size());
    for (size_t i = 0; i < cm.size(); ++i) out[i] = cm[idx[i]];
    return out;
}

// What the triangular sweeps pay for an elimination order (one chain): 64x64 tile products per tile row, counted at the
// granularity of the zero-block masks -- forward tiles are dense in every 8-row block that reaches them, backward tiles
// in every 4-row group (jk_sweep.cuh).  Envelope-based (no fill left of a row's first coupled node).
static double sweep_cost(int Nn, const std::vector<int>& conn, const std::vector<int>& order) {
    const int n = (int)order.size();
    std::vector<int> pos(Nn, -1), first(n);
    for (int i = 0; i < n; ++i) { pos[order[i]] = i; first[i] = i; }
    const int M = (int)conn.size() / 2;
    for (int m = 0; m < M; ++m) {
        int a = pos[conn[2 * m]], b = pos[conn[2 * m + 1]];
        if (a < 0 || b < 0) continue;
        int hi = std::max(a, b), lo = std::min(a, b);
        first[hi] = std::min(first[hi], lo);
    }
    const int ndof = 6 * n;
    double cost = 0.0;
    auto reach_tile = [&](int r0, int r1) {          // first tile column touched by DOF rows [r0, r1)
        int f = r0;
        for (int nd = r0 / 6; nd <= (std::min(r1, ndof) - 1) / 6; ++nd) f = std::min(f, 6 * first[nd]);
        return f / NB;
    };
    for (int r0 = 0; r0 < ndof; r0 += 8) cost += std::max(0, r0 / NB - reach_tile(r0, r0 + 8)) / 8.0;
    for (int r0 = 0; r0 < ndof; r0 += 4) cost += std::max(0, r0 / NB - reach_tile(r0, r0 + 4)) / 16.0;
    return cost;
}

// node half-bandwidth and envelope size (sum of row widths) of an ordering
static void order_quality(int Nn, const std::vector<int>& conn, const std::vector<int>& order, int& hbn, long long& profile) {
    std::vector<int> pos(Nn, -1), first(order.size());
    for (size_t i = 0; i < order.size(); ++i) { pos[order[i]] = (int)i; first[i] = (int)i; }
    hbn = 0;
    int M = (int)conn.size() / 2;
    for (int m = 0; m < M; ++m) {
        int a = pos[conn[2 * m]], b = pos[conn[2 * m + 1]];
        if (a < 0 || b < 0) continue;
        int hi = std::max(a, b), lo = std::min(a, b);
        hbn = std::max(hbn, hi - lo);
        first[hi] = std::min(first[hi], lo);
    }
    profile = 0;
    for (size_t i = 0; i < order.size(); ++i) profile += (long long)i - first[i];
}

// Item list of one sweep of one chain (see jk_sweep.cuh).  kx = first partial (forward) / known (backward) tile row,
// NT for a plain sweep.  Ring slots follow the processing order: seq(k) = k (forward) or ktop - k (backward).
// jmin (nullable, [NT]): first tile column in which tile row k of the factor can hold a non-zero (row envelope of
// the chain, no fill outside it): tiles left of it are dropped from the program.  The tile next to the diagonal is
// always kept -- the last item of a row waiting on the row before it is what orders the ring and the slab updates.
static void build_sweep_program(int NT, int bw, int kx, bool backward, const int* jmin, std::vector<uint4>& prog, int& pre_row, int& npre, int& ktop) {
    prog.clear();
    auto push = [&](int row, int src, int flags, int xinfo, int next_row, int next_init) {
        prog.push_back(make_uint4((unsigned)row, (unsigned)src, (unsigned)flags, (unsigned)xinfo));
        prog.push_back(make_uint4((unsigned)next_row, (unsigned)next_init, 0u, 0u));
        prog.push_back(make_uint4(0u, 0u, 0u, 0u));
    };
    auto slotinfo = [&](int seq) { return (seq % SW_RING) | (((seq / SW_RING) & 1) << 8); };
    auto first_col = [&](int k) { return jmin ? std::max(0, std::min(jmin[k], k)) : 0; };
    // a consumer warp has to wait on an operand row's mbarrier only the first time the program uses that row (program
    // order = every warp's order, and a row solved earlier was waited on when the row after it was finished)
    std::vector<char> waited(NT, 0);
    auto need_wait = [&](int r) { if (waited[r]) return 0; waited[r] = 1; return SW_WAIT_X; };
    if (!backward) {
        pre_row = 0; npre = 0; ktop = 0;
        for (int k = 0; k < NT; ++k) {
            const bool partial = k >= kx, has_next = k + 1 < NT;
            const int hi = std::min(k - 1, kx - 1);
            int lo = std::max(std::max(0, k - bw), first_col(k));
            if (!partial) lo = std::min(lo, k - 1);                 // always keep (k, k-1)
            const int end_flags = SW_ROW_END | (partial ? SW_NO_RING : SW_OUT_FRAG);
            const int out = (k % SW_RING) << 16;
            if (hi < lo || hi < 0) { push(k, k, SW_ROW_BEGIN | SW_INIT_RHS | SW_NO_OPERAND | end_flags, out, has_next ? k + 1 : 0, has_next ? 1 : 0); continue; }
            lo = std::max(lo, 0);
            for (int j = lo; j <= hi; ++j) {
                int flags = (j == lo ? (SW_ROW_BEGIN | SW_INIT_RHS) : 0) | (j == hi ? end_flags : 0) | need_wait(j);
                push(k, j, flags, slotinfo(j) | out, has_next ? k + 1 : 0, (j == lo && has_next) ? 1 : 0);   // next row's RHS is fetched at ROW_BEGIN
            }
        }
    } else {
        npre = kx < NT ? std::min(bw, NT - kx) : 0;
        pre_row = kx;
        ktop = kx + npre - 1;
        for (int k = kx - 1; k >= 0; --k) {
            const int hi = std::min(NT - 1, std::min(k + bw, ktop));
            const int out = ((ktop - k) % SW_RING) << 16;
            push(k, k, SW_ROW_BEGIN | SW_DIAG | (hi < k + 1 ? SW_ROW_END : 0), out, 0, 0);
            for (int i = hi; i >= k + 1; --i) {
                if (i != k + 1 && first_col(i) > k) continue;       // tile (i, k) lies left of row i's envelope
                push(k, i, (i == k + 1 ? SW_ROW_END : 0) | need_wait(i), slotinfo(ktop - i) | out, 0, 0);
            }
        }
    }
}

// First tile column of every tile row of a chain's factor.  node_first[s] = lowest chain slot coupled to slot s (<= s);
// Cholesky fills the row envelope and nothing left of it.
static std::vector<int> chain_tile_reach(int NT, int n_nodes_chain, const std::vector<int>& node_first) {
    std::vector<int> jmin(NT);
    for (int k = 0; k < NT; ++k) {
        int reach = k * NB;                                        // padded rows: identity
        const int n_lo = (k * NB) / 6, n_hi = std::min(n_nodes_chain - 1, (k * NB + NB - 1) / 6);
        for (int nd = n_lo; nd <= n_hi; ++nd) reach = std::min(reach, 6 * node_first[nd]);
        jmin[k] = reach / NB;
    }
    return jmin;
}

// Metadata of the fused load lumping (LoadFuse): members sorted by their upper node (the end later in the solver's linear
// order; fixed nodes come first), chunks of MCHUNK consecutive members, one row per member (deposit of the lower end) and per
// run of members sharing the upper node inside a chunk, and, for every node, the rows to add -- listed with the LAST chunk
// that touches the node.
static int build_load_fusion(jk_handle_t h, const std::vector<int>& pos /* position of free nodes, -1 for fixed */) {
    const int M = h->M, Nn = h->Nn;
    auto key = [&](int node) { return pos[node] < 0 ? -1 : pos[node]; };
    // upper / lower end of every member: larger key wins, ties (two supports) go to end 2
    std::vector<int> up_end(M), up_node(M), lo_node(M);
    for (int m = 0; m < M; ++m) {
        const int a = h->h_conn[2 * m], b = h->h_conn[2 * m + 1];
        up_end[m] = key(b) >= key(a) ? 1 : 0;
        up_node[m] = up_end[m] ? b : a; lo_node[m] = up_end[m] ? a : b;
    }
    std::vector<int> order(M);
    for (int m = 0; m < M; ++m) order[m] = m;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        const int kx = key(up_node[x]), ky = key(up_node[y]);
        if (kx != ky) return kx < ky;
        if (up_node[x] != up_node[y]) return up_node[x] < up_node[y];
        return key(lo_node[x]) < key(lo_node[y]);
    });
    const int n_chunk = ceil_div(M, MCHUNK);
    std::vector<unsigned> ends(M);
    std::vector<int> pair_base(n_chunk + 1, 0);
    std::vector<std::vector<int>> rows_of_node(Nn);       // rows in a fixed order: ascending chunk, ascending position
    std::vector<int> last_chunk(Nn, -1), first_chunk(Nn, -1);
    int n_pairs = 0;
    for (int c = 0; c < n_chunk; ++c) {
        pair_base[c] = n_pairs;
        const int q0 = c * MCHUNK, q1 = std::min(M, (c + 1) * MCHUNK);
        int n_local = 0, run_row = -1;
        auto touch = [&](int node) { if (first_chunk[node] < 0) first_chunk[node] = c; last_chunk[node] = c; };
        for (int q = q0; q < q1; ++q) {
            const int m = order[q];
            const bool begin = q == q0 || up_node[order[q - 1]] != up_node[m];
            const bool end = q + 1 == q1 || up_node[order[q + 1]] != up_node[m];
            if (begin) run_row = n_local++;
            const int lo_row = n_local++;
            ends[q] = (unsigned)lo_row | ((unsigned)run_row << 8) | ((unsigned)up_end[m] << 16) | (begin ? 1u << 17 : 0u) | (end ? 1u << 18 : 0u);
            rows_of_node[lo_node[m]].push_back(n_pairs + lo_row); touch(lo_node[m]);
            if (end) { rows_of_node[up_node[m]].push_back(n_pairs + run_row); touch(up_node[m]); }
        }
        if (n_local > 255) JK_FAIL(h, JK_EINVAL, "internal: a member chunk needs more than 255 load rows");
        n_pairs += n_local;
    }
    pair_base[n_chunk] = n_pairs;
    // finalisation lists: node -> its last chunk (nodes without any member -- only supports can be -- go to chunk 0 with no rows)
    std::vector<std::vector<int>> fin(n_chunk);
    for (int node = 0; node < Nn; ++node) fin[last_chunk[node] < 0 ? 0 : last_chunk[node]].push_back(node);
    std::vector<int> fin_ptr(n_chunk + 1, 0), fin_node, fin_rowptr(1, 0), fin_rows, dep_lo(n_chunk);
    for (int c = 0; c < n_chunk; ++c) {
        fin_ptr[c] = (int)fin_node.size();
        int lo = c;
        for (int node : fin[c]) {
            fin_node.push_back(node);
            for (int r : rows_of_node[node]) fin_rows.push_back(r);
            fin_rowptr.push_back((int)fin_rows.size());
            if (first_chunk[node] >= 0) lo = std::min(lo, first_chunk[node]);
        }
        dep_lo[c] = lo;
    }
    fin_ptr[n_chunk] = (int)fin_node.size();
    h->fuse_n_pairs = n_pairs; h->fuse_n_fin = (int)fin_node.size(); h->fuse_n_chunk = n_chunk;
    auto up_i = [&](int*& d, const std::vector<int>& v) -> cudaError_t {
        cudaError_t e = dev_alloc(&d, v.size());
        return e != cudaSuccess ? e : cudaMemcpy(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice);
    };
    CUDA_TRY(h, up_i(h->d_fuse_order, order));
    CUDA_TRY(h, up_i(h->d_fuse_pair_base, pair_base));
    CUDA_TRY(h, up_i(h->d_fuse_fin_ptr, fin_ptr));
    CUDA_TRY(h, up_i(h->d_fuse_fin_node, fin_node));
    CUDA_TRY(h, up_i(h->d_fuse_fin_rowptr, fin_rowptr));
    CUDA_TRY(h, up_i(h->d_fuse_fin_rows, fin_rows));
    CUDA_TRY(h, up_i(h->d_fuse_dep_lo, dep_lo));
    std::vector<int> fin_of_node(Nn, 0);
    for (size_t e = 0; e < fin_node.size(); ++e) fin_of_node[fin_node[e]] = (int)e;
    CUDA_TRY(h, up_i(h->d_fuse_fin_of_node, fin_of_node));
    CUDA_TRY(h, dev_alloc(&h->d_fuse_ends, ends.size()));
    CUDA_TRY(h, cudaMemcpy(h->d_fuse_ends, ends.data(), ends.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    return JK_OK;
}

extern "C" int jk_set_supports(jk_handle_t h, int n_fixed, const int32_t* fixed_nodes, int ordering, int solver) {
    if (!h) return JK_EINVAL;
    if (n_fixed <= 0 || !fixed_nodes) JK_FAIL(h, JK_EINVAL, "jk_set_supports: at least one fixed node is required (K_ff would be singular)");
    cudaSetDevice(h->device);
    std::vector<char> is_fixed(h->Nn, 0);
    h->h_fixed.clear();
    for (int i = 0; i < n_fixed; ++i) {
        int f = fixed_nodes[i];
        if (f < 0 || f >= h->Nn) JK_FAIL(h, JK_EINVAL, "jk_set_supports: fixed node %d out of range", f);
        if (!is_fixed[f]) { is_fixed[f] = 1; h->h_fixed.push_back(f); }
    }
    h->n_fixed = (int)h->h_fixed.size();
    h->n_free_nodes = h->Nn - h->n_fixed;
    if (h->n_free_nodes <= 0) JK_FAIL(h, JK_EINVAL, "jk_set_supports: every node is fixed");
    std::vector<std::vector<int>> nbr;
    if (ordering == JK_ORDER_RCM) {
        // two candidates: single-root RCM and RCM rooted at the support-adjacent set; keep the narrower band (then the
        // smaller envelope)
        free_graph(h->Nn, h->h_conn, is_fixed, nbr);
        std::vector<int> o1, o2, cm2, lev2;
        rcm_order(h->Nn, nbr, is_fixed, o1);
        rcm_order_from_supports(h->Nn, h->h_conn, nbr, is_fixed, o2, &cm2, &lev2);
        int hb1 = 0, hb2 = 0; long long pr1 = 0, pr2 = 0;
        order_quality(h->Nn, h->h_conn, o1, hb1, pr1);
        order_quality(h->Nn, h->h_conn, o2, hb2, pr2);
        // envelope of L at DOF level: 36 entries per coupled node pair inside the row envelope + 21 per diagonal node block
        h->nnz_env_min = 36LL * std::min(pr1, o2.size() == o1.size() ? pr2 : pr1) + 21LL * (long long)o1.size();
        const bool second = h->opt[OPT_SUPPORT_ROOTED_RCM] && o2.size() == o1.size() && (hb2 < hb1 || (hb2 == hb1 && pr2 < pr1));
        h->h_free_nodes.swap(second ? o2 : o1);
        if (second && h->opt[OPT_LEVEL_REGROUP]) {
            // further candidates: same level structure, every level sorted by degree (either way, ties either way).  Kept when the band does not
            // widen and the sweeps get cheaper (more non-zeros in L, but fewer mask blocks to multiply).
            double best = sweep_cost(h->Nn, h->h_conn, h->h_free_nodes);
            for (int v = 0; v < 4; ++v) {
                std::vector<int> g = regroup_levels_by_degree(cm2, lev2, nbr, (v & 1) != 0, (v & 2) != 0);
                std::vector<int> cand(g.rbegin(), g.rend());
                int hbc = 0; long long prc = 0;
                order_quality(h->Nn, h->h_conn, cand, hbc, prc);
                h->nnz_env_min = std::min(h->nnz_env_min, 36LL * prc + 21LL * (long long)cand.size());
                const double c = sweep_cost(h->Nn, h->h_conn, cand);
                if (hbc <= hb2 && c < 0.98 * best) { best = c; h->h_free_nodes.swap(cand); }
            }
        }
    }
    else {
        h->h_free_nodes.clear(); for (int i = 0; i < h->Nn; ++i) if (!is_fixed[i]) h->h_free_nodes.push_back(i);
        int hb0 = 0; long long pr0 = 0;
        order_quality(h->Nn, h->h_conn, h->h_free_nodes, hb0, pr0);
        h->nnz_env_min = 36LL * pr0 + 21LL * (long long)h->h_free_nodes.size();
    }
    if ((int)h->h_free_nodes.size() != h->n_free_nodes) JK_FAIL(h, JK_EINVAL, "jk_set_supports: internal ordering error");
    h->n_free = 6 * h->n_free_nodes;
    h->solver = solver;
    const int N = h->n_free_nodes;
    std::vector<int> pos(h->Nn, -1);                       // position of a free node in the linear (RCM / natural) order
    for (int i = 0; i < N; ++i) pos[h->h_free_nodes[i]] = i;
    int hbn = 0;                                           // node half-bandwidth of that order
    for (int m = 0; m < h->M; ++m) {
        int a = pos[h->h_conn[2 * m]], b = pos[h->h_conn[2 * m + 1]];
        if (a >= 0 && b >= 0) hbn = std::max(hbn, std::abs(a - b));
    }
    // two chains when the band is narrow and long enough: |A|, |B| multiples of 32 nodes (= 3 tiles), separator >= hbn nodes
    int nA = N, nS = 0, nB = 0;
    h->n_chains = 1;
    if (solver == JK_SOLVER_BANDED && h->opt[OPT_TWO_CHAINS] && N - hbn >= 4 * 32 && 6 * hbn + 5 <= 12 * NB) {
        int avail = N - hbn;
        nA = 32 * (avail / 64);
        nB = 32 * ((avail - nA) / 32);
        nS = N - nA - nB;
        h->n_chains = 2;
    }
    h->nS_nodes = nS;
    if (h->n_chains == 2 && ordering == JK_ORDER_RCM) {
        // The second chain eliminates B in REVERSED order (from the far end towards S).  Re-number B by Cuthill-McKee
        // from the separator outwards: reversed, that is an RCM ordering rooted at S, whose envelope (and so the fill of
        // that chain's L) is as small as the first chain's.
        std::vector<char> inB(h->Nn, 0), visited(h->Nn, 0);
        for (int i = nA + nS; i < N; ++i) inB[h->h_free_nodes[i]] = 1;
        std::vector<int> roots, cm;
        for (int i = nA; i < nA + nS; ++i)
            for (int v : nbr[h->h_free_nodes[i]]) if (inB[v] && !visited[v]) { visited[v] = 1; roots.push_back(v); }
        for (int v : roots) visited[v] = 0;
        std::vector<int> levB(h->Nn, 0);
        cm_from(nbr, roots, inB, visited, cm, &levB);
        if ((int)cm.size() == nB) {
            int hb_old = 0; long long pr_old = 0;
            std::vector<int> rev_old(h->h_free_nodes.rbegin(), h->h_free_nodes.rbegin() + nB + nS);
            order_quality(h->Nn, h->h_conn, rev_old, hb_old, pr_old);
            double best = sweep_cost(h->Nn, h->h_conn, rev_old);
            std::vector<int> best_order;
            for (int variant = 0; variant < (h->opt[OPT_LEVEL_REGROUP] ? 5 : 1); ++variant) {
                std::vector<int> g = variant == 0 ? cm : regroup_levels_by_degree(cm, levB, nbr, ((variant - 1) & 1) != 0, ((variant - 1) & 2) != 0);
                std::vector<int> trial(h->h_free_nodes);
                std::copy(g.begin(), g.end(), trial.begin() + nA + nS);
                std::vector<int> rev_new(trial.rbegin(), trial.rbegin() + nB + nS);
                int hb_new = 0; long long pr_new = 0;
                order_quality(h->Nn, h->h_conn, rev_new, hb_new, pr_new);
                const double c = sweep_cost(h->Nn, h->h_conn, rev_new);
                if (hb_new <= hb_old && c < best) { best = c; best_order.swap(trial); }
            }
            if (!best_order.empty()) { h->h_free_nodes.swap(best_order); for (int i = 0; i < N; ++i) pos[h->h_free_nodes[i]] = i; }
        }
    }
    // local slot of every free node in its chain(s): chain 0 = [A; S] in order, chain 1 = [rev(B); rev(S)]
    std::vector<int> slot0(h->Nn, -1), slot1(h->Nn, -1);
    for (int i = 0; i < nA + nS; ++i) slot0[h->h_free_nodes[i]] = i;
    if (h->n_chains == 2) for (int i = nA; i < N; ++i) slot1[h->h_free_nodes[i]] = N - 1 - i;   // B: 0..nB-1, S: nB..nB+nS-1 (both reversed)
    auto& c0 = h->ch[0]; auto& c1 = h->ch[1];
    c0.n_rows = 6 * (nA + nS); c0.NT = ceil_div(c0.n_rows, NB); c0.kS = (h->n_chains == 2) ? (6 * nA) / NB : c0.NT; c0.row0 = 0;
    c1.n_rows = (h->n_chains == 2) ? 6 * (nB + nS) : 0; c1.NT = ceil_div(c1.n_rows, NB); c1.kS = (6 * nB) / NB; c1.row0 = c0.NT * NB;
    h->NT = c0.NT + c1.NT;
    h->n_pad = h->NT * NB;
    // node -> first row of the slab (solution lives in chain 0 for A and S nodes, in chain 1 for B nodes)
    h->h_node2slot.assign(h->Nn, 0);
    for (int i = 0; i < h->n_fixed; ++i) h->h_node2slot[h->h_fixed[i]] = -1 - i;
    for (int i = 0; i < N; ++i) {
        int node = h->h_free_nodes[i];
        h->h_node2slot[node] = (i < nA + nS) ? 6 * slot0[node] : c1.row0 + 6 * slot1[node];
    }

    // member chunks of the post-processing kernel: "early" = every member's nodes are solved once the first chain's
    // backward sweep is through (A and S nodes live in chain 0's rows, fixed nodes have no rows), "late" = the rest
    {
        const int n_mchunk = ceil_div(h->M, MCHUNK);
        std::vector<int> early, late;
        for (int ch = 0; ch < n_mchunk; ++ch) {
            bool e = h->n_chains == 2;
            for (int m = ch * MCHUNK; e && m < std::min(h->M, (ch + 1) * MCHUNK); ++m)
                for (int q = 0; q < 2; ++q) { const int sl = h->h_node2slot[h->h_conn[2 * m + q]]; if (sl >= c1.row0) e = false; }
            (e ? early : late).push_back(ch);
        }
        h->n_post_early = (int)early.size(); h->n_post_late = (int)late.size();
        early.insert(early.end(), late.begin(), late.end());
        dev_free(h->d_post_chunks);
        CUDA_TRY(h, dev_alloc(&h->d_post_chunks, early.size()));
        CUDA_TRY(h, cudaMemcpy(h->d_post_chunks, early.data(), early.size() * sizeof(int), cudaMemcpyHostToDevice));
    }

    // 6x6 block lists with contributions in member order (deterministic assembly), one list per chain.
    // Separator-separator blocks are assembled into chain 0 only (chain 1's trailing block collects -W_B W_B^T).
    struct Contrib { long long key; int member, quad; };
    std::vector<char> touched(h->Nn, 0);
    std::vector<KBlock> blocks[2];
    std::vector<int2> contribs[2];
    for (int c = 0; c < h->n_chains; ++c) {
        const std::vector<int>& sl = c == 0 ? slot0 : slot1;
        auto& chn = h->ch[c];
        const long long nn = (long long)N + 1;
        std::vector<Contrib> cs;
        cs.reserve(3 * (size_t)h->M);
        int bw = 0, hb = 0;
        auto span = [&](int rs, int cslot) { int I = (6 * rs + 5) / NB, J = (6 * cslot) / NB; bw = std::max(bw, I - J); hb = std::max(hb, 6 * (rs - cslot) + 5); };
        auto in_sep = [&](int node) { int q = pos[node]; return h->n_chains == 2 && q >= nA && q < nA + nS; };
        for (int m = 0; m < h->M; ++m) {
            int na = h->h_conn[2 * m], nb = h->h_conn[2 * m + 1];
            int s0 = sl[na], s1 = sl[nb];
            bool sepa = in_sep(na), sepb = in_sep(nb);
            if (s0 >= 0 && !(c == 1 && sepa)) { cs.push_back({(long long)s0 * nn + s0, m, 0}); span(s0, s0); touched[na] = 1; }
            if (s1 >= 0 && !(c == 1 && sepb)) { cs.push_back({(long long)s1 * nn + s1, m, 3}); span(s1, s1); touched[nb] = 1; }
            if (s0 >= 0 && s1 >= 0 && !(c == 1 && sepa && sepb)) {
                if (s0 > s1) { cs.push_back({(long long)s0 * nn + s1, m, (0 << 1) | 1}); span(s0, s1); }
                else { cs.push_back({(long long)s1 * nn + s0, m, (1 << 1) | 0}); span(s1, s0); }
            }
        }
        std::stable_sort(cs.begin(), cs.end(), [](const Contrib& a, const Contrib& b) { return a.key < b.key; });
        contribs[c].resize(cs.size());
        for (size_t i = 0; i < cs.size(); ++i) {
            contribs[c][i] = make_int2(cs[i].member, cs[i].quad);
            if (i == 0 || cs[i].key != cs[i - 1].key) {
                KBlock kb; kb.row_slot = (int)(cs[i].key / nn); kb.col_slot = (int)(cs[i].key % nn);
                kb.start = (int)i; kb.count = 0; blocks[c].push_back(kb);
            }
            blocks[c].back().count++;
        }
        chn.nblocks = (int)blocks[c].size();
        chn.bw = (solver == JK_SOLVER_DENSE) ? (chn.NT - 1) : std::min(bw, chn.NT - 1);
        chn.hb = hb;
        chn.tiles_elems = (size_t)chn.NT * (size_t)(chn.bw + 1) * NB * NB;
    }
    // a member that joins A and B directly would break the separator (cannot happen when nS >= hbn); free nodes
    // without any member would make K_ff singular: report both now
    for (int m = 0; m < h->M && h->n_chains == 2; ++m) {
        int a = pos[h->h_conn[2 * m]], b = pos[h->h_conn[2 * m + 1]];
        if (a >= 0 && b >= 0 && ((a < nA && b >= nA + nS) || (b < nA && a >= nA + nS))) JK_FAIL(h, JK_EINVAL, "jk_set_supports: internal separator error");
    }
    for (int i = 0; i < N; ++i) if (!touched[h->h_free_nodes[i]]) JK_FAIL(h, JK_EINVAL, "jk_set_supports: free node %d has no member attached", h->h_free_nodes[i]);
    h->bw = std::max(c0.bw, c1.bw);
    h->hb = std::max(c0.hb, c1.hb);
    // row order of the slab for jk_get_order: chain 0 nodes (A, S), then chain 1's own nodes (B reversed)
    { std::vector<int> ord2(h->h_free_nodes.begin(), h->h_free_nodes.begin() + nA + nS);
      for (int i = N - 1; i >= nA + nS; --i) ord2.push_back(h->h_free_nodes[i]);
      h->h_free_nodes.swap(ord2); }

    CUDA_TRY(h, dev_alloc(&h->d_node2slot, (size_t)h->Nn));
    CUDA_TRY(h, dev_alloc(&h->d_fixed_nodes, (size_t)h->n_fixed));
    CUDA_TRY(h, dev_alloc(&h->d_free_nodes, (size_t)h->n_free_nodes));
    cudaStream_t s = h->stream;
    for (int c = 0; c < 2; ++c) {
        auto& chn = h->ch[c];
        if (c >= h->n_chains) { dev_free(chn.d_blocks); dev_free(chn.d_contrib); dev_free(chn.d_tiles); dev_free(chn.d_Linv); dev_free(chn.d_dinv); chn.NT = 0; continue; }
        CUDA_TRY(h, dev_alloc(&chn.d_blocks, blocks[c].size()));
        CUDA_TRY(h, dev_alloc(&chn.d_contrib, contribs[c].size()));
        CUDA_TRY(h, dev_alloc(&chn.d_tiles, chn.tiles_elems));
        CUDA_TRY(h, dev_alloc(&chn.d_Linv, (size_t)chn.NT * NB * NB));
        CUDA_TRY(h, dev_alloc(&chn.d_dinv, (size_t)chn.NT * 512));
        CUDA_TRY(h, cudaMemcpyAsync(chn.d_blocks, blocks[c].data(), blocks[c].size() * sizeof(KBlock), cudaMemcpyHostToDevice, s));
        CUDA_TRY(h, cudaMemcpyAsync(chn.d_contrib, contribs[c].data(), contribs[c].size() * sizeof(int2), cudaMemcpyHostToDevice, s));
    }
    // row envelope of each chain's factor at tile granularity (symbolic: no fill left of a row's first coupled slot;
    // the separator block of the first chain is dense once the second chain's Schur complement is merged in)
    std::vector<int> tile_reach[2];
    for (int c = 0; c < h->n_chains; ++c) {
        const std::vector<int>& sl = c == 0 ? slot0 : slot1;
        const int nn = c == 0 ? nA + nS : nB + nS;
        std::vector<int> node_first(nn);
        for (int i = 0; i < nn; ++i) node_first[i] = i;
        for (int m = 0; m < h->M; ++m) {
            int s0 = sl[h->h_conn[2 * m]], s1 = sl[h->h_conn[2 * m + 1]];
            if (s0 < 0 || s1 < 0) continue;
            int hi = std::max(s0, s1), lo = std::min(s0, s1);
            node_first[hi] = std::min(node_first[hi], lo);
        }
        if (h->n_chains == 2) { const int sep0 = c == 0 ? nA : nB; for (int i = sep0; i < nn; ++i) node_first[i] = std::min(node_first[i], sep0); }
        tile_reach[c] = chain_tile_reach(h->ch[c].NT, nn, node_first);
    }
    // sweep programs: the TMA pipeline keeps the last SW_RING solved tiles in shared memory, so it needs a narrow band
    h->tma_sweep = h->opt[OPT_TMA_SWEEP] != 0;
    for (int c = 0; c < h->n_chains; ++c) if (h->ch[c].bw > SW_MAX_BW) h->tma_sweep = false;
    h->split_factor = false;
    for (int c = 0; c < 2; ++c)
        for (int d = 0; d < 2; ++d) {
            auto& w = h->ch[c].sw[d];
            dev_free(w.d_prog); dev_free(w.d_stream); w.n_items = 0;
            if (c >= h->n_chains || !h->tma_sweep) continue;
            auto& chn = h->ch[c];
            std::vector<uint4> prog;
            // the first chain sweeps all its rows (its separator rows are ordinary rows once the chains are merged);
            // the second chain's separator rows are partial (forward) / known (backward)
            build_sweep_program(chn.NT, chn.bw, c == 0 ? chn.NT : chn.kS, d == 1, tile_reach[c].data(), prog, w.pre_row, w.npre, w.ktop);
            w.n_items = (int)prog.size() / SW_ITEM_U4;
            w.n_split = w.n_items; w.k_split = 0; w.npre2 = 0; w.xphase2 = 0;
            if (d == 0 && w.n_items > 0) {
                // split point: 80 % of the chain's own columns (the separator columns of the first chain come last anyway)
                // 70 % of the chain's own columns: on c4 the first segment then ends well inside the Morison stage, so its
                // inverse / stream-build kernels and the relaunch of the cluster kernel do not collide with the gather and
                // the sweeps (60-80 % measure the same within 1 %; close to 90 % the second segment competes with the sweep
                // CTAs for whole SMs and the step degrades badly)
                const int split_pct = h->opt[OPT_SPLIT_PCT];
                const int kS = chn.kS, k1 = (int)((long long)split_pct * kS / 100);
                if (k1 >= 2 * SW_RING && kS - k1 >= 4) {
                    int n1 = 0;
                    while (n1 < w.n_items && (int)prog[(size_t)n1 * SW_ITEM_U4].x < k1) ++n1;
                    w.n_split = n1; w.k_split = k1; w.npre2 = std::min(chn.bw, k1);
                    // slot s is next filled (preloaded or computed) by the first row r >= k1 - npre2 with r % R == s; the
                    // item parities count fills from row 0, so the slot starts (r / R) & 1 phases ahead
                    for (int r = k1 - w.npre2; r < k1 - w.npre2 + SW_RING; ++r) if ((r / SW_RING) & 1) w.xphase2 |= 1 << (r % SW_RING);
                }
            }
            if (w.n_items == 0) continue;
            CUDA_TRY(h, dev_alloc(&w.d_prog, prog.size()));
            CUDA_TRY(h, dev_alloc(&w.d_stream, (size_t)w.n_items * SW_TILE));
            CUDA_TRY(h, cudaMemcpyAsync(w.d_prog, prog.data(), prog.size() * sizeof(uint4), cudaMemcpyHostToDevice, s));
            CUDA_TRY(h, cudaStreamSynchronize(s));   // prog is a local
        }
    // the factorisation runs in two segments (and the forward sweeps in two launches) when every chain has a split point
    h->split_factor = h->tma_sweep && h->opt[OPT_FACTOR_SPLIT];
    for (int c = 0; c < h->n_chains; ++c) if (h->ch[c].sw[0].k_split == 0) h->split_factor = false;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_node2slot, h->h_node2slot.data(), (size_t)h->Nn * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_fixed_nodes, h->h_fixed.data(), (size_t)h->n_fixed * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_free_nodes, h->h_free_nodes.data(), (size_t)h->n_free_nodes * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    { int rcf = build_load_fusion(h, pos); if (rcf != JK_OK) return rcf; }
    h->graph_epoch++;
    h->have_supports = true; h->assembled = false; h->factored = false;
    // buffers sized by n_pad / n_fixed must be rebuilt; resident results no longer match the new row numbering
    h->cap_ldP = 0; h->lastP = 0; h->last_fem = false;
    return JK_OK;
}

// the main stream joins a factorisation still running on the side streams (no host synchronisation).  After a scan the
// side streams have already been joined (run_fem waits for the backward tile streams), inside a captured step as well.
static int join_factor(jk_handle_t h, cudaStream_t s) {
    if (h->factor_inflight && h->ev_bwd_external && !h->capturing) CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_factor_bwd, 0));
    h->ev_bwd_external = false;
    return JK_OK;
}

static int assemble_launch(jk_handle_t h, double E, double G) {
    cudaStream_t s = h->stream;
    int rcj = join_factor(h, s);
    if (rcj != JK_OK) return rcj;
    h->factor_inflight = false;
    h->E = E; h->G = G;
    tic(h, JK_T_ASSEMBLE);
    k_member_setup<<<ceil_div(h->M, 128), 128, 0, s>>>(h->M, h->d_xyz, h->d_conn, h->d_sec, h->d_secp, JK_SEC_NPROP, E, G, h->d_mc, h->d_Ke, h->d_Kl);
    LAUNCH_CHECK(h);
    k_pack_member_consts<<<ceil_div(h->M, 128), 128, 0, s>>>(h->M, h->d_mc, h->sp, h->d_pk);
    LAUNCH_CHECK(h);
    for (int c = 0; c < h->n_chains; ++c) {
        auto& chn = h->ch[c];
        CUDA_TRY(h, cudaMemsetAsync(chn.d_tiles, 0, chn.tiles_elems * sizeof(double), s));
        k_assemble_blocks<<<chn.nblocks, 64, 0, s>>>(chn.nblocks, chn.d_blocks, chn.d_contrib, h->d_Ke, chn.d_tiles, chn.bw);
        LAUNCH_CHECK(h);
        if (chn.NT * NB > chn.n_rows) { k_pad_identity<<<1, NB, 0, s>>>(chn.n_rows, chn.NT * NB, chn.d_tiles, chn.bw); LAUNCH_CHECK(h); }
    }
    toc(h, JK_T_ASSEMBLE);
    h->assembled = true; h->factored = false;
    h->lastP = 0; h->last_fem = false;        // resident rows belong to the previous stiffness
    return JK_OK;
}

extern "C" int jk_assemble(jk_handle_t h, double E, double G) {
    if (!h) return JK_EINVAL;
    if (!h->have_supports) JK_FAIL(h, JK_ESTATE, "jk_assemble: call jk_set_supports first");
    if (!(E > 0) || !(G > 0)) JK_FAIL(h, JK_EINVAL, "jk_assemble: E and G must be positive");
    cudaSetDevice(h->device);
    return assemble_launch(h, E, G);
}

// end of a factorisation: latch a non-positive pivot into the sticky flag (reported by the next call that talks to the host;
// the on-device argmax poisons the critical pair meanwhile)
__global__ void k_latch_info(const int* __restrict__ info, int* __restrict__ sticky) { if (*info != 0 && *sticky == 0) *sticky = *info; }

__global__ void k_debug_spin(long long clocks) { const long long t0 = clock64(); while (clock64() - t0 < clocks) { } }

// part: 0 = whole program, 1 = items [0, n_split), 2 = items [n_split, n_items)   (forward programs of a split factor)
static int launch_sweep_build(jk_handle_t h, cudaStream_t s, int d, int part = 0) {
    if (!h->tma_sweep) return JK_OK;
    SweepBuildArgs a[2];
    for (int c = 0; c < 2; ++c) {
        auto& w = h->ch[c].sw[d];
        const int lo = part == 2 ? w.n_split : 0, hi = part == 1 ? w.n_split : w.n_items;
        a[c] = SweepBuildArgs{w.d_prog ? w.d_prog + (size_t)lo * SW_ITEM_U4 : nullptr, w.d_stream ? w.d_stream + (size_t)lo * SW_TILE : nullptr,
                              h->ch[c].d_tiles, h->ch[c].d_Linv, h->ch[c].bw, c < h->n_chains ? std::max(0, hi - lo) : 0};
    }
    if (a[0].n_items + a[1].n_items > 0) {
        k_sweep_build<<<a[0].n_items + a[1].n_items, 256, SWB_SMEM, s>>>(a[0], a[1], d);
        LAUNCH_CHECK(h);
    }
    return JK_OK;
}

// launches the factorisation on stream s (no host synchronisation)
static int launch_factor(jk_handle_t h, cudaStream_t s, cudaStream_t s_side = nullptr) {
    tic(h, JK_T_FACTOR, s);
    h->gate2_armed = false;
    CUDA_TRY(h, cudaMemsetAsync(h->d_info, 0, sizeof(int), s));
    // narrow band: persistent cluster kernel(s) (latency chain); wide band / dense: per-column launches
    const bool use_cluster = (h->factor_path == 0) && (h->bw <= 16);
    auto& c0 = h->ch[0]; auto& c1 = h->ch[1];
    if (use_cluster) {
        long long* prof = nullptr;
        const bool want_prof = h->opt[OPT_PROFILE_CHOL] != 0;
        if (want_prof) { CUDA_TRY(h, cudaMalloc((void**)&prof, (size_t)c0.NT * 8 * sizeof(long long))); CUDA_TRY(h, cudaMemsetAsync(prof, 0, (size_t)c0.NT * 8 * sizeof(long long), s)); }
        CholChain a{c0.d_tiles, c0.d_dinv, c0.NT, c0.bw, 0, c0.kS};
        unsigned* gate = (s_side != nullptr && h->wait_value32 && h->opt[OPT_START_GATE]) ? h->d_started : nullptr;   // asynchronous factorisation only
        if (h->opt[OPT_DEBUG_FACTOR_DELAY] > 0) k_debug_spin<<<1, 32, 0, s>>>((long long)h->opt[OPT_DEBUG_FACTOR_DELAY]);   // test aid: lose the race on purpose
        if (h->split_factor && s_side != nullptr) {
            // Two segments.  After the first (80 % of each chain's columns) a side stream inverts those diagonal tiles and
            // builds the forward tile streams of those rows, so the forward sweeps can start on them (run_fem) while this
            // stream factors the rest, the separator, and builds the remaining streams.
            auto& c1r = h->ch[h->n_chains == 2 ? 1 : 0];
            const int k0 = c0.sw[0].k_split, k1 = c1r.sw[0].k_split;
            CholChain a1{c0.d_tiles, c0.d_dinv, c0.NT, c0.bw, 0, k0}, a2{c0.d_tiles, c0.d_dinv, c0.NT, c0.bw, k0, c0.kS};
            CholChain b1{c1.d_tiles, c1.d_dinv, c1.NT, c1.bw, 0, k1}, b2{c1.d_tiles, c1.d_dinv, c1.NT, c1.bw, k1, c1.kS};
            const int ncl = h->n_chains == 2 ? 2 : 1;
            k_band_chol_cluster<<<ncl * CHOL_CLUSTER, CHOL_THREADS, CHOL_CLUSTER_SMEM, s>>>(a1, ncl == 2 ? b1 : a1, h->d_info, prof, gate);
            if (gate) h->started_target += (unsigned)(ncl * CHOL_CLUSTER);
            LAUNCH_CHECK(h);
            CUDA_TRY(h, cudaEventRecord(h->ev_seg1, s));
            CUDA_TRY(h, cudaStreamWaitEvent(s_side, h->ev_seg1, 0));
            k_tile_inverse_blocked<<<k0 + (ncl == 2 ? k1 : 0), 256, INVERSE_BLOCKED_SMEM, s_side>>>(c0.d_tiles, c0.d_dinv, c0.d_Linv, c0.bw, 0, k0,
                                                                                             c1.d_tiles, c1.d_dinv, c1.d_Linv, c1.bw, 0);
            LAUNCH_CHECK(h);
            int rc1 = launch_sweep_build(h, s_side, 0, 1);
            if (rc1 != JK_OK) return rc1;
            CUDA_TRY(h, cudaEventRecord(h->ev_fwd1, s_side));
            // second start gate: run_fem launches the first forward parts (one 200 KB CTA per SM on 128 SMs) only once these
            // clusters are resident.  With the Morison + load stage as short as the first segment the sweeps otherwise win
            // the race now and then, the clusters find no GPC with eight free SMs until both parts are through and the step
            // grows by ~1 ms (seen as 5.5 vs 6.7 ms per step between runs).
            const bool gate2_on = h->opt[OPT_START_GATE2] != 0;
            k_band_chol_cluster<<<ncl * CHOL_CLUSTER, CHOL_THREADS, CHOL_CLUSTER_SMEM, s>>>(a2, ncl == 2 ? b2 : a2, h->d_info, nullptr, gate2_on ? gate : nullptr);
            LAUNCH_CHECK(h);
            h->gate2_armed = false;
            if (gate && gate2_on) { h->gate2_target = h->started_target + (unsigned)(ncl * CHOL_CLUSTER); h->gate2_armed = true; }
            if (ncl == 2) {
                long long n = 36LL * h->nS_nodes * h->nS_nodes;
                k_sep_merge_tiles<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c0.d_tiles, c0.bw, c0.kS, c1.d_tiles, c1.bw, c1.kS, h->nS_nodes);
                LAUNCH_CHECK(h);
                CholChain f{c0.d_tiles, c0.d_dinv, c0.NT, c0.bw, c0.kS, c0.NT};
                k_band_chol_cluster<<<CHOL_CLUSTER, CHOL_THREADS, CHOL_CLUSTER_SMEM, s>>>(f, f, h->d_info, nullptr);
                LAUNCH_CHECK(h);
            }
            const int n0b = c0.NT - k0, n1b = ncl == 2 ? c1.kS - k1 : 0;
            k_tile_inverse_blocked<<<n0b + n1b, 256, INVERSE_BLOCKED_SMEM, s>>>(c0.d_tiles, c0.d_dinv, c0.d_Linv, c0.bw, k0, n0b,
                                                                                c1.d_tiles, c1.d_dinv, c1.d_Linv, c1.bw, k1);
            LAUNCH_CHECK(h);
            int rc2 = launch_sweep_build(h, s, 0, 2);
            if (rc2 != JK_OK) return rc2;
            CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_fwd1, 0));      // the factor stage is complete when both branches are
            k_latch_info<<<1, 1, 0, s>>>(h->d_info, h->d_info_sticky);
            LAUNCH_CHECK(h);
            toc(h, JK_T_FACTOR, s);
            if (want_prof) { CUDA_TRY(h, cudaStreamSynchronize(s)); cudaFree(prof); }
            return JK_OK;
        }
        if (h->n_chains == 2) {
            // stage 1: both chains eliminate towards the separator, concurrently on two clusters
            CholChain b{c1.d_tiles, c1.d_dinv, c1.NT, c1.bw, 0, c1.kS};
            k_band_chol_cluster<<<2 * CHOL_CLUSTER, CHOL_THREADS, CHOL_CLUSTER_SMEM, s>>>(a, b, h->d_info, prof, gate);
            if (gate) h->started_target += (unsigned)(2 * CHOL_CLUSTER);
            LAUNCH_CHECK(h);
            // stage 2: separator Schur complement = sum of both chains' trailing blocks
            long long n = 36LL * h->nS_nodes * h->nS_nodes;
            k_sep_merge_tiles<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c0.d_tiles, c0.bw, c0.kS, c1.d_tiles, c1.bw, c1.kS, h->nS_nodes);
            LAUNCH_CHECK(h);
            // stage 3: factor the separator (trailing columns of chain 0)
            CholChain f{c0.d_tiles, c0.d_dinv, c0.NT, c0.bw, c0.kS, c0.NT};
            k_band_chol_cluster<<<CHOL_CLUSTER, CHOL_THREADS, CHOL_CLUSTER_SMEM, s>>>(f, f, h->d_info, nullptr);
            LAUNCH_CHECK(h);
        } else {
            k_band_chol_cluster<<<CHOL_CLUSTER, CHOL_THREADS, CHOL_CLUSTER_SMEM, s>>>(a, a, h->d_info, prof, gate);
            if (gate) h->started_target += (unsigned)CHOL_CLUSTER;
            LAUNCH_CHECK(h);
        }
        if (want_prof) {   // debug aid: average clock deltas between the phase stamps of chain 0, CTA 0
            std::vector<long long> hp((size_t)c0.NT * 8);
            CUDA_TRY(h, cudaMemcpyAsync(hp.data(), prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(h, cudaStreamSynchronize(s));
            cudaFree(prof);
            double acc[6] = {0, 0, 0, 0, 0, 0}; int n = 0;
            for (int k = 8; k + 8 < c0.kS; ++k, ++n) for (int i = 0; i < 6; ++i) acc[i] += (double)(hp[(size_t)k * 8 + i + 1] - hp[(size_t)k * 8 + i]);
            if (n > 0) fprintf(stderr, "[jk chol profile] CTA0 clocks/column: load %.0f | trsm+store %.0f | arrive+syrk %.0f | potrf %.0f | store+wait(A) %.0f | barrier B %.0f\n",
                               acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n, acc[5] / n);
        }
    } else {
        // per-column launches (dense storage or very wide bands): single chain only
        for (int k = 0; k < c0.NT; ++k) {
            int w = std::min(c0.bw, c0.NT - 1 - k);
            k_potrf_tile<<<1, 256, 0, s>>>(c0.d_tiles, k, c0.bw, h->d_info);
            LAUNCH_CHECK(h);
            if (w > 0) {
                k_panel_trsm<<<w, NB, PANEL_SMEM, s>>>(c0.d_tiles, k, c0.bw);
                LAUNCH_CHECK(h);
                k_trailing_update<<<w * (w + 1) / 2, 128, UPDATE_SMEM, s>>>(c0.d_tiles, k, w, c0.bw);
                LAUNCH_CHECK(h);
            }
        }
    }
    { const int n1 = h->n_chains == 2 ? c1.kS : 0;     // second chain: its separator rows are factored in the first chain
      if (use_cluster && h->opt[OPT_BLOCKED_INVERSE])     // the cluster kernel left the 8x8 block inverses in d_dinv
          k_tile_inverse_blocked<<<c0.NT + n1, 256, INVERSE_BLOCKED_SMEM, s>>>(c0.d_tiles, c0.d_dinv, c0.d_Linv, c0.bw, 0, c0.NT,
                                                                               c1.d_tiles, c1.d_dinv, c1.d_Linv, c1.bw, 0);
      else
          k_tile_inverse<<<c0.NT + n1, 256, INVERSE_SMEM, s>>>(c0.d_tiles, c0.d_Linv, c0.bw, c0.NT, c1.d_tiles, c1.d_Linv, c1.bw);
      LAUNCH_CHECK(h); }
    // tile streams of the forward sweeps; the factor timer stops here (this is what the forward sweeps wait for)
    int rc = launch_sweep_build(h, s, 0);
    if (rc != JK_OK) return rc;
    k_latch_info<<<1, 1, 0, s>>>(h->d_info, h->d_info_sticky);
    LAUNCH_CHECK(h);
    toc(h, JK_T_FACTOR, s);
    return JK_OK;
}

// host-side completion of an asynchronous factorisation: pivot check.  Call only after the streams were synchronised.
static int finish_factor(jk_handle_t h) {
    if (h->factor_inflight) CUDA_TRY(h, cudaStreamSynchronize(h->stream2));
    h->factor_inflight = false;
    int info = 0;
    CUDA_TRY(h, cudaMemcpy(&info, h->d_info_sticky, sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0) {
        CUDA_TRY(h, cudaMemset(h->d_info_sticky, 0, sizeof(int)));
        h->factored = false;
        JK_FAIL(h, JK_ENOTSPD, "K_ff is not positive definite (pivot %d <= 0): structure is a mechanism or badly supported", info - 1);
    }
    return JK_OK;
}

extern "C" int jk_factor(jk_handle_t h) {
    if (!h) return JK_EINVAL;
    if (!h->assembled) JK_FAIL(h, JK_ESTATE, "jk_factor: call jk_assemble first");
    cudaSetDevice(h->device);
    int rc = launch_factor(h, h->stream);
    if (rc != JK_OK) return rc;
    if ((rc = launch_sweep_build(h, h->stream, 1)) != JK_OK) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev_fwd1, h->stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_factor, h->stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_factor_bwd, h->stream));
    h->ev_bwd_external = true;
    h->assembled = false;       // the tile storage now holds L
    h->factored = true;
    h->factor_inflight = true;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return finish_factor(h);
}

// Asynchronous variant: the factorisation is queued on the handle's side stream (after everything already queued on
// the main stream) and the call returns at once.  The next scan runs its Morison + load stage concurrently and joins
// before the triangular sweeps; a non-positive pivot is reported by that scan's jk_read_table / jk_phase_scan.
static int factor_begin_launch(jk_handle_t h) {
    // the gate counter restarts with every asynchronous factorisation: targets are then the same constants in every step,
    // which is what a captured graph replays
    if (h->d_started) { CUDA_TRY(h, cudaMemsetAsync(h->d_started, 0, sizeof(unsigned), h->stream)); h->started_target = 0; }
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    const unsigned target0 = h->started_target;
    int rc = launch_factor(h, h->stream2, h->stream3);
    if (rc != JK_OK) return rc;
    if (h->wait_value32 && h->started_target != target0) {
        // the main stream (Morison next) continues once every CTA of the first cluster launch is resident
        if (h->wait_value32((CUstream)h->stream, (CUdeviceptr)h->d_started, (cuuint32_t)h->started_target, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
            JK_FAIL(h, JK_ECUDA, "jk_factor_begin: cuStreamWaitValue32 failed");
    }
    if (h->gate2_armed) h->started_target = h->gate2_target;     // the counter also counts the second segment's CTAs
    CUDA_TRY(h, cudaEventRecord(h->ev_factor, h->stream2));
    // the backward tile streams are only needed after the forward sweeps: built behind the event, they overlap them
    if ((rc = launch_sweep_build(h, h->stream2, 1)) != JK_OK) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev_factor_bwd, h->stream2));
    h->assembled = false;
    h->factored = true;
    h->factor_inflight = true;
    h->ev_bwd_external = !h->capturing;
    return JK_OK;
}

extern "C" int jk_factor_begin(jk_handle_t h) {
    if (!h) return JK_EINVAL;
    if (!h->assembled) JK_FAIL(h, JK_ESTATE, "jk_factor_begin: call jk_assemble first");
    cudaSetDevice(h->device);
    return factor_begin_launch(h);
}

extern "C" int jk_set_static_load(jk_handle_t h, const double* F) {
    if (!h || !F) return JK_EINVAL;
    cudaSetDevice(h->device);
    const size_t nb = 6 * (size_t)h->Nn * sizeof(double);
    if (unsigned char* pin = pin_acquire(h, nb)) {
        memcpy(pin, F, nb);                                   // the caller's buffer is free again when this returns
        CUDA_TRY(h, cudaMemcpyAsync(h->d_Fstatic, pin, nb, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaEventRecord(h->ev_pin, h->stream));
        h->pin_busy = true;
        return JK_OK;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->d_Fstatic, F, nb, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return JK_OK;
}

extern "C" int jk_set_wave_airy(jk_handle_t h, double a, double k, double omega, double d, double U_c, double dt) {
    if (!h) return JK_EINVAL;
    if (!(k > 0) || !(omega > 0) || !(d > 0) || !(dt > 0)) JK_FAIL(h, JK_EINVAL, "jk_set_wave_airy: k, omega, d, dt must be positive");
    h->wv.a = a; h->wv.k = k; h->wv.omega = omega; h->wv.d = d; h->wv.Uc = U_c; h->wv.dt = dt; h->wv.inv_dt = 1.0 / dt;
    h->wave_kind = 0; h->n_harm = 0;
    h->have_wave = true; h->gp_valid = false; h->graph_epoch++;
    return JK_OK;
}

extern "C" int jk_set_wave_fourier(jk_handle_t h, double k, double omega, double d, double U_c, double dt,
                                   int n_harm, const double* E, const double* B) {
    if (!h || !E || !B) return JK_EINVAL;
    if (!(k > 0) || !(omega > 0) || !(d > 0) || !(dt > 0)) JK_FAIL(h, JK_EINVAL, "jk_set_wave_fourier: k, omega, d, dt must be positive");
    if (n_harm < 1 || n_harm > FOURIER_MAX_H) JK_FAIL(h, JK_EINVAL, "jk_set_wave_fourier: n_harm must be in 1..%d", FOURIER_MAX_H);
    cudaSetDevice(h->device);
    std::vector<double> four(3 * (size_t)n_harm);
    for (int j = 0; j < n_harm; ++j) { four[j] = E[j]; four[n_harm + j] = B[j]; four[2 * n_harm + j] = 1.0 / cosh((j + 1) * k * d); }
    CUDA_TRY(h, dev_alloc(&h->d_four, four.size()));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_four, four.data(), four.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->wv.a = E[0]; h->wv.k = k; h->wv.omega = omega; h->wv.d = d; h->wv.Uc = U_c; h->wv.dt = dt; h->wv.inv_dt = 1.0 / dt;
    h->wave_kind = 1; h->n_harm = n_harm;
    h->have_wave = true; h->gp_valid = false; h->graph_epoch++;
    return JK_OK;
}

extern "C" int jk_set_morison(jk_handle_t h, double theta_wave, double theta_current, double rho, double Cd, double Cm,
                              int n_gauss, const double* gauss_s, const double* gauss_w) {
    if (!h || !gauss_s || !gauss_w) return JK_EINVAL;
    if (n_gauss <= 0 || n_gauss > 64) JK_FAIL(h, JK_EINVAL, "jk_set_morison: n_gauss must be in 1..64");
    cudaSetDevice(h->device);
    h->rho = rho; h->Cd = Cd; h->Cm = Cm; h->ng = n_gauss;
    h->wv.cos_w = cos(theta_wave); h->wv.sin_w = sin(theta_wave);
    h->wv.uc_cos_c = cos(theta_current); h->wv.uc_sin_c = sin(theta_current);   // multiplied by U_c at launch
    std::vector<double> gsw(2 * (size_t)n_gauss);
    for (int i = 0; i < n_gauss; ++i) { gsw[i] = gauss_s[i]; gsw[n_gauss + i] = gauss_w[i]; }
    CUDA_TRY(h, dev_alloc(&h->d_gsw, gsw.size()));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_gsw, gsw.data(), gsw.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_morison = true; h->gp_valid = false; h->graph_epoch++;
    return JK_OK;
}

// ------------------------------------------------------------------------------------------------
// scan buffers
// ------------------------------------------------------------------------------------------------
static int ensure_buffers(jk_handle_t h, int P, bool need_fem, bool need_details) {
    int ldP = ceil_div(P, SLAB) * SLAB;
    bool need = ldP > h->cap_ldP || (need_details && !h->cap_details) || (need_fem && !h->d_X);
    if (!need) return JK_OK;
    h->graph_epoch++;                       // buffers move
    ldP = std::max(ldP, h->cap_ldP);
    size_t l = (size_t)ldP;
    int n_mchunk = ceil_div(h->M, MCHUNK), n_nchunk = ceil_div(h->Nn, NCHUNK);
    CUDA_TRY(h, dev_alloc(&h->d_t, l));
    CUDA_TRY(h, dev_alloc(&h->d_trig, 4 * l));
    CUDA_TRY(h, dev_alloc(&h->d_Fm, (size_t)h->M * 6 * l));
    CUDA_TRY(h, dev_alloc(&h->d_totpart, (size_t)n_mchunk * 9 * l));
    CUDA_TRY(h, dev_alloc(&h->d_table, l * JK_TABLE_NCOL));
    if (need_details || h->cap_details) { CUDA_TRY(h, dev_alloc(&h->d_details, (size_t)h->M * 4 * l)); h->cap_details = true; }
    if (need_fem || h->d_X) {
        if (!h->have_supports) JK_FAIL(h, JK_ESTATE, "supports not set");
        CUDA_TRY(h, dev_alloc(&h->d_X, (size_t)h->n_pad * l));
        CUDA_TRY(h, cudaMemsetAsync(h->d_X, 0, (size_t)h->n_pad * l * sizeof(double), h->stream));
        if (h->tma_sweep) CUDA_TRY(h, dev_alloc(&h->d_Z, (size_t)h->n_pad * l));     // forward intermediate of the TMA sweeps (fragment order)
        CUDA_TRY(h, dev_alloc(&h->d_part, (size_t)h->fuse_n_pairs * 3 * l));
        CUDA_TRY(h, dev_alloc(&h->d_fuse_sync, 1 + (size_t)(h->fuse_n_chunk + FUSE_LAG) * ceil_div(ldP, PH_TPB)));
        CUDA_TRY(h, dev_alloc(&h->d_Ffix, (size_t)h->n_fixed * 6 * l));
        CUDA_TRY(h, dev_alloc(&h->d_react, (size_t)h->n_fixed * 6 * l));
        CUDA_TRY(h, dev_alloc(&h->d_rows, (size_t)h->M * JK_MEMBER_NCOL * l));
        CUDA_TRY(h, dev_alloc(&h->d_part_util, (size_t)n_mchunk * l));
        CUDA_TRY(h, dev_alloc(&h->d_part_vm, (size_t)n_mchunk * l));
        CUDA_TRY(h, dev_alloc(&h->d_part_mem, (size_t)n_mchunk * l));
        CUDA_TRY(h, dev_alloc(&h->d_part_disp, (size_t)n_nchunk * l));
        CUDA_TRY(h, dev_alloc(&h->d_part_node, (size_t)n_nchunk * l));
    }
    h->cap_ldP = ldP;
    return JK_OK;
}

static int ensure_tmp(jk_handle_t h, size_t n) {
    if (n <= h->tmp_elems) return JK_OK;
    CUDA_TRY(h, dev_alloc(&h->d_tmp, n));
    h->tmp_elems = n;
    return JK_OK;
}

static WaveAiry launch_wave(jk_handle_t h) {
    WaveAiry w = h->wv;
    w.uc_cos_c = h->wv.Uc * h->wv.uc_cos_c;   // U_c * cos(theta_c), GUI.py:582
    w.uc_sin_c = h->wv.Uc * h->wv.uc_sin_c;
    return w;
}

// Gauss-point tables of the current wave (once per wave / Morison set-up; allocation inside, so never during a capture)
static int ensure_gauss_tables(jk_handle_t h) {
    if (h->gp_valid) return JK_OK;
    cudaStream_t s = h->stream;
    WaveAiry w = launch_wave(h);
    const int Nh = h->n_harm;
    const size_t gstride = h->wave_kind == 1 ? (size_t)(3 + 2 * Nh) : (size_t)GP_STRIDE;
    int n = h->M * h->ng;
    size_t need = (size_t)n * gstride;
    if (need > h->gp_elems) { CUDA_TRY(h, dev_alloc(&h->d_gp, need)); h->gp_elems = need; h->graph_epoch++; }
    if (h->wave_kind == 1)
        k_gauss_setup_fourier<<<ceil_div(n, 128), 128, 0, s>>>(h->M, h->ng, Nh, h->d_xyz, h->d_conn, h->d_gsw, w, h->d_four, h->d_gp);
    else
        k_gauss_setup_airy<<<ceil_div(n, 128), 128, 0, s>>>(h->M, h->ng, h->d_xyz, h->d_conn, h->d_gsw, w, h->d_gp);
    LAUNCH_CHECK(h);
    h->gp_valid = true;
    return JK_OK;
}

// Morison stage for the phases already in d_t: trig tables, Gauss-point tables (once per wave), K1
// tile0 / ntiles: phase tiles (PH_TPB phases each) of this launch; ntiles < 0 = all.  Tables are set up with the first tile.
static int run_morison(jk_handle_t h, int P, int ldP, bool details, bool fuse = false, int tile0 = 0, int ntiles = -1) {
    cudaStream_t s = h->stream;
    WaveAiry w = launch_wave(h);
    const int Nh = h->n_harm;
    const size_t gstride = h->wave_kind == 1 ? (size_t)(3 + 2 * Nh) : (size_t)GP_STRIDE;
    const int all_tiles = ceil_div(ldP, PH_TPB);
    const bool ranged = ntiles >= 0;
    if (!ranged) ntiles = all_tiles;
    if (tile0 == 0) {
        tic(h, JK_T_WAVE_SETUP);
        { int rcg = ensure_gauss_tables(h); if (rcg != JK_OK) return rcg; }
        k_phase_setup<<<ceil_div(ldP, 128), 128, 0, s>>>(P, ldP, h->d_t, w.omega, w.dt, h->d_trig);
        LAUNCH_CHECK(h);
        toc(h, JK_T_WAVE_SETUP);
        tic(h, JK_T_MORISON);
    }
    const int p_off = tile0 * PH_TPB;
    dim3 grid(ntiles, ceil_div(h->M, MCHUNK));
    double cD0 = 0.5 * h->rho * h->Cd, cI0 = h->rho * h->Cm;
    if (h->wave_kind == 1) {
        size_t smem = ((size_t)FCHUNK * h->ng * gstride + MCHUNK * 8 + 2 * h->ng + 3 * Nh) * sizeof(double);
        if (details) {
            CUDA_TRY(h, cudaFuncSetAttribute(k_morison_fourier<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_morison_fourier<true><<<grid, PH_TPB, smem, s>>>(h->M, h->ng, Nh, ldP, h->d_gp, h->d_mc, h->d_gsw, h->d_trig, h->d_four, w, cD0, cI0, h->d_Fm, h->d_totpart, h->d_details);
        } else {
            CUDA_TRY(h, cudaFuncSetAttribute(k_morison_fourier<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_morison_fourier<false><<<grid, PH_TPB, smem, s>>>(h->M, h->ng, Nh, ldP, h->d_gp, h->d_mc, h->d_gsw, h->d_trig, h->d_four, w, cD0, cI0, h->d_Fm, h->d_totpart, nullptr);
        }
    } else {
        size_t smem = ((size_t)MCHUNK * h->ng * (GP_STRIDE + MORISON_AIRY_SMEM_PER_POINT_EXTRA) + MCHUNK * MORISON_AIRY_SMEM_PER_MEMBER_EXTRA + 2 * h->ng + 1) * sizeof(double)
                      + (size_t)h->opt[OPT_DEBUG_MORISON_SMEM_PAD] * 1024;
        if (details) {
            CUDA_TRY(h, cudaFuncSetAttribute(k_morison_airy<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_morison_airy<true, 0><<<grid, PH_TPB, smem, s>>>(h->M, h->ng, ldP, h->d_gp, h->d_mc, h->d_gsw, h->d_trig, w, cD0, cI0, h->d_Fm, h->d_totpart, h->d_details, 0);
        } else {
            // the default 15-point rule may have its own instantiation with fully unrolled point loops (-DJK_MORISON_G15=1)
            if (fuse) {
                // load lumping inside the kernel (LoadFuse): ticket + flags start at zero
                const int n_tiles = (int)grid.x;
                CUDA_TRY(h, cudaMemsetAsync(h->d_fuse_sync, 0, (1 + (size_t)(h->fuse_n_chunk + FUSE_LAG) * n_tiles) * sizeof(int), s));
                LoadFuse lf{h->d_fuse_order, h->d_fuse_ends, h->d_fuse_pair_base, h->d_part, h->d_fuse_fin_ptr, h->d_fuse_fin_node,
                            h->d_fuse_fin_rowptr, h->d_fuse_fin_rows, h->d_fuse_dep_lo, h->d_fuse_sync, h->d_fuse_sync + 1,
                            h->d_node2slot, h->d_Fstatic, h->d_X, h->d_Ffix, h->n_pad, n_tiles, h->opt[OPT_DEBUG_FUSE_MODE]};
                auto kern = (JK_MORISON_G15 && h->ng == 15) ? k_morison_airy<false, 15, true> : k_morison_airy<false, 0, true>;
                CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kern<<<dim3(grid.x, grid.y + FUSE_LAG), PH_TPB, smem, s>>>(h->M, h->ng, ldP, h->d_gp, h->d_mc, h->d_gsw, h->d_trig, w, cD0, cI0, nullptr, h->d_totpart, nullptr, 0, lf);
            } else {
                auto kern = (JK_MORISON_G15 && h->ng == 15) ? k_morison_airy<false, 15> : k_morison_airy<false, 0>;
                CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kern<<<grid, PH_TPB, smem, s>>>(h->M, h->ng, ldP, h->d_gp, h->d_mc, h->d_gsw, h->d_trig, w, cD0, cI0, h->d_Fm, h->d_totpart, nullptr, p_off, LoadFuse{});
            }
        }
    }
    LAUNCH_CHECK(h);
    if (tile0 + ntiles >= all_tiles) toc(h, JK_T_MORISON);
    return JK_OK;
}

// member constants are produced by jk_assemble; Morison-only use needs them too
static int ensure_member_consts(jk_handle_t h) {
    if (h->E > 0) return JK_OK;   // jk_assemble ran
    // geometry-only constants: run the setup kernel with unit moduli (stiffness entries unused by Morison)
    k_member_setup<<<ceil_div(h->M, 128), 128, 0, h->stream>>>(h->M, h->d_xyz, h->d_conn, h->d_sec, h->d_secp, JK_SEC_NPROP, 1.0, 1.0, h->d_mc, h->d_Ke, h->d_Kl);
    LAUNCH_CHECK(h);
    return JK_OK;
}

// Morison columns of the table (t, totals) on a side stream right after the Morison kernel: 3/4 of the reduction's reads
// leave the tail of the scan.  The caller joins ev_tot before reduce_and_argmax(..., totals_done = true).
static int reduce_totals_early(jk_handle_t h, int P, int ldP) {
    cudaStream_t s = h->stream, s3 = h->stream3;
    CUDA_TRY(h, cudaEventRecord(h->ev_mor, s));
    CUDA_TRY(h, cudaStreamWaitEvent(s3, h->ev_mor, 0));
    k_phase_reduce<<<ceil_div(P, 32), 32 * RED_GROUPS, 0, s3>>>(P, ldP, h->d_t, ceil_div(h->M, MCHUNK), h->d_totpart, 0, nullptr, nullptr, nullptr,
                                                      0, nullptr, nullptr, 0, nullptr, h->d_table, JK_TABLE_NCOL, 1, h->wv.omega);
    LAUNCH_CHECK(h);
    CUDA_TRY(h, cudaEventRecord(h->ev_tot, s3));
    return JK_OK;
}

static int reduce_and_argmax(jk_handle_t h, int P, int ldP, bool morison, bool fem, bool totals_done = false,
                             const double* st_omega = nullptr, int n_phase = 1) {
    cudaStream_t s = h->stream;
    if (totals_done) CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_tot, 0));
    tic(h, JK_T_REDUCE);
    int n_mchunk = ceil_div(h->M, MCHUNK), n_nchunk = ceil_div(h->Nn, NCHUNK);
    k_phase_reduce<<<ceil_div(P, 32), 32 * RED_GROUPS, 0, s>>>(P, ldP, h->d_t, n_mchunk, (morison && !totals_done) ? h->d_totpart : nullptr,
                                                     n_mchunk, fem ? h->d_part_util : nullptr, h->d_part_vm, h->d_part_mem,
                                                     n_nchunk, fem ? h->d_part_disp : nullptr, h->d_part_node,
                                                     h->n_fixed, fem ? h->d_react : nullptr, h->d_table, JK_TABLE_NCOL, totals_done ? 0 : 1,
                                                     morison ? h->wv.omega : -1.0, st_omega, n_phase);
    LAUNCH_CHECK(h);
    k_argmax<<<1, 1024, 0, s>>>(P, h->d_table, JK_TABLE_NCOL, morison ? JK_COL_TOTAL_KN : JK_COL_MAX_UTIL, h->d_argval, h->d_argidx,
                                fem ? h->d_info_sticky : nullptr);
    LAUNCH_CHECK(h);
    toc(h, JK_T_REDUCE);
    return JK_OK;
}

// solve + post for the right-hand sides already in d_X / d_Ffix
static int run_fem(jk_handle_t h, int ldP, double fy) {
    cudaStream_t s = h->stream;
    int nslab = ldP / SLAB;
    if (h->factor_inflight && !h->tma_sweep) CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_factor, 0));   // join the side stream (the TMA path joins below)
    auto& c0 = h->ch[0]; auto& c1 = h->ch[1];
    const int nS6 = 6 * h->nS_nodes;
    dim3 gsep(ceil_div(ldP, 128), std::max(1, nS6));
    const bool sweep_prof = h->opt[OPT_PROFILE_SWEEP] != 0;
    long long* d_prof = nullptr;
    constexpr int TRACE_ITEMS = 1024;
    const size_t prof_elems = 4 * 64 + 4 * (size_t)TRACE_ITEMS * 8;
    const bool sweep_trace = h->opt[OPT_PROFILE_SWEEP] >= 2;     // per-item clock stamps of two warps of CTA 0 (stderr)
    if (sweep_prof && h->tma_sweep) { CUDA_TRY(h, cudaMalloc((void**)&d_prof, prof_elems * sizeof(long long))); CUDA_TRY(h, cudaMemsetAsync(d_prof, 0, prof_elems * sizeof(long long), s)); }
    // Right-hand sides per sweep CTA: a whole slab (32) when the slabs fill the GPU, otherwise half or a quarter of a slab per CTA
    // (option sweep_slab = 32 / 16 / 8 forces it).  Few phases make a sweep a latency chain per CTA; narrower CTAs shorten it.
    int ncb = 4;
    if (h->opt[OPT_SWEEP_SLAB] > 0) ncb = h->opt[OPT_SWEEP_SLAB] / 8;
    else if (2 * nslab <= h->n_sm) ncb = (4 * nslab <= h->n_sm + h->n_sm / 2) ? 1 : 2;
    const int sweep_ctas = nslab * (4 / ncb);
    h->sweep_slab_last = 8 * ncb;
    // part: 0 = whole program, 1 = items [0, n_split), 2 = the rest (continues from the rows part 1 left in the slab)
    auto sweep = [&](int c, int d, int part = 0, unsigned* started = nullptr) {
        auto& chn = h->ch[c]; auto& w = chn.sw[d];
        const int lo = part == 2 ? w.n_split : 0, hi = part == 1 ? w.n_split : w.n_items;
        if (hi <= lo) return false;
        const bool cont = part == 2;
        const uint4* prog = w.d_prog + (size_t)lo * SW_ITEM_U4;
        const double* strm = w.d_stream + (size_t)lo * SW_TILE;
        const int pre_row = cont ? w.k_split - w.npre2 : w.pre_row, npre = cont ? w.npre2 : w.npre, xph = cont ? w.xphase2 : 0;
        long long* pf = d_prof ? d_prof + (2 * c + d) * 64 : nullptr;
        long long* tr = (d_prof && sweep_trace && part == 0 && hi - lo <= TRACE_ITEMS) ? d_prof + 4 * 64 + (size_t)(2 * c + d) * TRACE_ITEMS * 8 : nullptr;
        auto launch = [&](auto kern, int nc) {
            kern<<<sweep_ctas, sw_threads(nc), SW_SMEM, s>>>(prog, strm, h->d_X, h->d_Z, hi - lo, h->n_pad, chn.row0, pre_row, npre, w.ktop, cont ? 1 : 0, xph, pf, started, tr);
        };
        if (pf) { if (ncb == 4) launch(k_sweep<4, true>, 4); else if (ncb == 2) launch(k_sweep<2, true>, 2); else launch(k_sweep<1, true>, 1); }
        else { if (ncb == 4) launch(k_sweep<4, false>, 4); else if (ncb == 2) launch(k_sweep<2, false>, 2); else launch(k_sweep<1, false>, 1); }
        return true;
    };
    // Early member post (see d_post_chunks): only when the backward sweep of the second chain leaves SMs idle (one CTA per
    // SM, fewer slabs than SMs) -- the post blocks are released once every sweep CTA is resident (same counter as the
    // factor's start gates), so they can only take what the sweep does not use.
    const bool post_overlap_on = h->opt[OPT_POST_OVERLAP] != 0;
    bool post_early = false;
    dim3 gm_all(ceil_div(h->M, MCHUNK), ceil_div(ldP, JK_POST_TPB));
    if (h->tma_sweep) {
        // same elimination-tree order as below; between its forward and backward sweep a chain's rows hold Z = L_kk Y_k
        // in fragment order (jk_sweep.cuh)
        const bool split = h->split_factor && h->factor_inflight && c0.sw[0].n_split < c0.sw[0].n_items;
        if (h->factor_inflight) CUDA_TRY(h, cudaStreamWaitEvent(s, split ? h->ev_fwd1 : h->ev_factor, 0));   // join the side stream(s)
        tic(h, JK_T_SOLVE_FWD);
        if (h->n_chains == 2)
            CUDA_TRY(h, cudaMemset2DAsync(h->d_X + ((size_t)c1.row0 + (size_t)c1.kS * NB) * SLAB, (size_t)h->n_pad * SLAB * sizeof(double), 0,
                                          (size_t)(c1.NT - c1.kS) * NB * SLAB * sizeof(double), (size_t)nslab, s));
        if (split) {
            // rows below the split point of both chains only need the first factor segment: they run while the
            // factorisation finishes (the chains do not depend on each other before the separator)
            if (h->gate2_armed) {                            // ... once the second segment's clusters hold their SMs
                h->gate2_armed = false;
                if (h->wait_value32((CUstream)s, (CUdeviceptr)h->d_started, (cuuint32_t)h->gate2_target, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                    JK_FAIL(h, JK_ECUDA, "run_fem: cuStreamWaitValue32 failed");
            }
            if (h->n_chains == 2) { sweep(1, 0, 1); LAUNCH_CHECK(h); }
            sweep(0, 0, 1); LAUNCH_CHECK(h);
            toc(h, JK_T_SOLVE_FWD);                           // first parts; the continuation is timed as SOLVE_FWD2
            CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_factor, 0));
            tic(h, JK_T_SOLVE_FWD2);
            if (h->n_chains == 2) { sweep(1, 0, 2); LAUNCH_CHECK(h); }
        } else if (h->n_chains == 2) { sweep(1, 0); LAUNCH_CHECK(h); }
        if (h->n_chains == 2) {
            k_sep_exchange<<<gsep, 128, 0, s>>>(h->d_X, h->n_pad, ldP, c0.row0 + c0.kS * NB, c1.row0 + c1.kS * NB, h->nS_nodes, 0);
            LAUNCH_CHECK(h);
        }
        sweep(0, 0, split ? 2 : 0); LAUNCH_CHECK(h);
        if (split) toc(h, JK_T_SOLVE_FWD2); else { toc(h, JK_T_SOLVE_FWD); h->ev_set[JK_T_SOLVE_FWD2] = false; }
        if (h->factor_inflight) { CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_factor_bwd, 0)); h->ev_bwd_external = false; }
        tic(h, JK_T_SOLVE_BWD);
        sweep(0, 1); LAUNCH_CHECK(h);
        if (h->n_chains == 2) {
            const bool overlap = post_overlap_on && h->wait_value32 && h->stream3 && h->d_started && h->n_post_early >= 8 && sweep_ctas + 2 <= h->n_sm;
            if (overlap) CUDA_TRY(h, cudaEventRecord(h->ev_bwd0, s));
            k_sep_exchange<<<gsep, 128, 0, s>>>(h->d_X, h->n_pad, ldP, c0.row0 + c0.kS * NB, c1.row0 + c1.kS * NB, h->nS_nodes, 1);
            LAUNCH_CHECK(h);
            const bool launched = sweep(1, 1, 0, overlap ? h->d_started : nullptr); LAUNCH_CHECK(h);
            if (overlap && launched) {
                h->started_target += (unsigned)sweep_ctas;
                cudaStream_t s3 = h->stream3;
                CUDA_TRY(h, cudaStreamWaitEvent(s3, h->ev_bwd0, 0));
                if (h->wait_value32((CUstream)s3, (CUdeviceptr)h->d_started, (cuuint32_t)h->started_target, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                    JK_FAIL(h, JK_ECUDA, "run_fem: cuStreamWaitValue32 failed");
                k_member_post<<<dim3(h->n_post_early, gm_all.y), JK_POST_TPB, 0, s3>>>(h->M, ldP, ldP, h->n_pad, h->d_X, h->d_node2slot, h->d_conn, h->d_mc, h->sp, fy,
                                                                                   h->d_rows, h->d_part_util, h->d_part_vm, h->d_part_mem, h->d_post_chunks, h->d_pk);
                LAUNCH_CHECK(h);
                CUDA_TRY(h, cudaEventRecord(h->ev_post_early, s3));
                post_early = true;
            }
        }
        toc(h, JK_T_SOLVE_BWD);
        if (d_prof) {   // debug aid: where the consumer warps of CTA 0 spend their clocks
            std::vector<long long> hpv(prof_elems);
            long long* hp = hpv.data();
            CUDA_TRY(h, cudaMemcpyAsync(hp, d_prof, prof_elems * sizeof(long long), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(h, cudaStreamSynchronize(s));
            cudaFree(d_prof);
            if (sweep_trace)
                for (int q = 0; q < 4; ++q) {
                    auto& w = h->ch[q / 2].sw[q % 2];
                    if (q / 2 >= h->n_chains || w.n_items == 0 || w.n_items > TRACE_ITEMS) continue;
                    std::vector<uint4> hprog((size_t)w.n_items * SW_ITEM_U4);
                    cudaMemcpy(hprog.data(), w.d_prog, hprog.size() * sizeof(uint4), cudaMemcpyDeviceToHost);
                    const long long* tr = hp + 4 * 64 + (size_t)q * TRACE_ITEMS * 8;
                    for (int n = 0; n < w.n_items; ++n) {
                        const uint4 it = hprog[(size_t)n * SW_ITEM_U4], mk = hprog[(size_t)n * SW_ITEM_U4 + 2];
                        const int cells = __builtin_popcount(mk.x) + __builtin_popcount(mk.y) + __builtin_popcount(mk.z) + __builtin_popcount(mk.w);
                        fprintf(stderr, "[jk sweep trace] %d %d row %u src %u flags %u cells %d |", q, n, it.x, it.y, it.z, cells);
                        for (int tw = 0; tw < 2; ++tw) fprintf(stderr, " %lld %lld %lld %lld", tr[(n * 2 + tw) * 4] - tr[0], tr[(n * 2 + tw) * 4 + 1] - tr[0], tr[(n * 2 + tw) * 4 + 2] - tr[0], tr[(n * 2 + tw) * 4 + 3] - tr[0]);
                        fprintf(stderr, "\n");
                    }
                }
            for (int q = 0; q < 4; ++q) {
                if (h->ch[q / 2].sw[q % 2].n_items == 0 || q / 2 >= h->n_chains) continue;
                fprintf(stderr, "[jk sweep profile] chain %d %s: %d items\n", q / 2, q % 2 ? "backward" : "forward", h->ch[q / 2].sw[q % 2].n_items);
                for (int w = 0; w < 8; ++w) {   // first eight consumer warps
                    const long long* p = hp + q * 64 + w * 8;
                    fprintf(stderr, "   warp %d: total %lld | wait tile %lld | row begin %lld | wait operand %lld | mma loop %lld | row store %lld | DMMAs %lld | items %lld\n",
                            w, p[6], p[0], p[7], p[1], p[2], p[3], p[4], p[5]);
                }
            }
        }
    } else {
    tic(h, JK_T_SOLVE_FWD);
    if (h->n_chains == 2) {
        // second chain first: its separator rows collect -L_SB y_B (they start at zero), merged into the first chain's
        CUDA_TRY(h, cudaMemset2DAsync(h->d_X + ((size_t)c1.row0 + (size_t)c1.kS * NB) * SLAB, (size_t)h->n_pad * SLAB * sizeof(double), 0,
                                      (size_t)(c1.NT - c1.kS) * NB * SLAB * sizeof(double), (size_t)nslab, s));
        k_slab_sweep<false><<<nslab, SOLVE_THREADS, SOLVE_SMEM, s>>>(c1.d_tiles, c1.d_Linv, h->d_X, c1.NT, c1.bw, h->n_pad, c1.row0, c1.kS);
        LAUNCH_CHECK(h);
        k_sep_exchange<<<gsep, 128, 0, s>>>(h->d_X, h->n_pad, ldP, c0.row0 + c0.kS * NB, c1.row0 + c1.kS * NB, h->nS_nodes, 0);
        LAUNCH_CHECK(h);
    }
    k_slab_sweep<false><<<nslab, SOLVE_THREADS, SOLVE_SMEM, s>>>(c0.d_tiles, c0.d_Linv, h->d_X, c0.NT, c0.bw, h->n_pad, c0.row0, c0.NT);
    LAUNCH_CHECK(h);
    toc(h, JK_T_SOLVE_FWD);
    tic(h, JK_T_SOLVE_BWD);
    k_slab_sweep<true><<<nslab, SOLVE_THREADS, SOLVE_SMEM, s>>>(c0.d_tiles, c0.d_Linv, h->d_X, c0.NT, c0.bw, h->n_pad, c0.row0, c0.NT);
    LAUNCH_CHECK(h);
    if (h->n_chains == 2) {
        k_sep_exchange<<<gsep, 128, 0, s>>>(h->d_X, h->n_pad, ldP, c0.row0 + c0.kS * NB, c1.row0 + c1.kS * NB, h->nS_nodes, 1);
        LAUNCH_CHECK(h);
        k_slab_sweep<true><<<nslab, SOLVE_THREADS, SOLVE_SMEM, s>>>(c1.d_tiles, c1.d_Linv, h->d_X, c1.NT, c1.bw, h->n_pad, c1.row0, c1.kS);
        LAUNCH_CHECK(h);
    }
    toc(h, JK_T_SOLVE_BWD);
    }
    tic(h, JK_T_POST);
    // the two small node-level kernels (max translation, reactions) only read the solution: they run on the side stream
    // beside the member post-processing
    cudaStream_t s2 = h->stream2 ? h->stream2 : s;
    if (s2 != s) { CUDA_TRY(h, cudaEventRecord(h->ev_post_fork, s)); CUDA_TRY(h, cudaStreamWaitEvent(s2, h->ev_post_fork, 0)); }
    dim3 gn(ceil_div(h->Nn, NCHUNK), ceil_div(ldP, PH_TPB));
    k_node_post<<<gn, PH_TPB, 0, s2>>>(h->Nn, ldP, h->n_pad, h->d_X, h->d_node2slot, h->d_part_disp, h->d_part_node);
    LAUNCH_CHECK(h);
    dim3 gr(ceil_div(ldP, PH_TPB), h->n_fixed);
    k_node_residual<<<gr, PH_TPB, 0, s2>>>(h->n_fixed, h->d_fixed_nodes, ldP, h->n_pad, h->d_X, h->d_node2slot, h->d_conn,
                                           h->d_adj_ptr, h->d_adj, h->d_Ke, h->d_Ffix, h->d_react);
    LAUNCH_CHECK(h);
    if (s2 != s) CUDA_TRY(h, cudaEventRecord(h->ev_post_join, s2));
    if (post_early) {
        if (h->n_post_late > 0) {
            k_member_post<<<dim3(h->n_post_late, gm_all.y), JK_POST_TPB, 0, s>>>(h->M, ldP, ldP, h->n_pad, h->d_X, h->d_node2slot, h->d_conn, h->d_mc, h->sp, fy,
                                                                              h->d_rows, h->d_part_util, h->d_part_vm, h->d_part_mem, h->d_post_chunks + h->n_post_early, h->d_pk);
            LAUNCH_CHECK(h);
        }
        CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_post_early, 0));
    } else {
        k_member_post<<<gm_all, JK_POST_TPB, 0, s>>>(h->M, ldP, ldP, h->n_pad, h->d_X, h->d_node2slot, h->d_conn, h->d_mc, h->sp, fy,
                                                 h->d_rows, h->d_part_util, h->d_part_vm, h->d_part_mem, nullptr, h->d_pk);
        LAUNCH_CHECK(h);
    }
    if (s2 != s) CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_post_join, 0));
    toc(h, JK_T_POST);
    return JK_OK;
}

static int scan_core(jk_handle_t h, int P, double fy, bool fem) {
    int ldP = ceil_div(P, SLAB) * SLAB;
    cudaStream_t s = h->stream;
    int rc;
    if ((rc = ensure_member_consts(h)) != JK_OK) return rc;
    const int nbx = ceil_div(ldP, PH_TPB);
    const bool fuse = fem && h->opt[OPT_FUSED_LOADS] && h->wave_kind == 0 && h->fuse_n_pairs > 0;
    // Phase blocks: the Morison kernel is FP64-bound, the load gather HBM-bound -- the gather of block b runs on the
    // high-priority side stream while the Morison kernel works on block b + 1.  Same kernels, same arithmetic per phase.
    int nblk = (fem && !fuse && h->wave_kind == 0 && h->stream3 && h->ev_gather) ? std::min(h->opt[OPT_GATHER_BLOCKS], std::min(16, nbx)) : 1;
    bool gathered = false;
    if (nblk > 1) {
        cudaStream_t s3 = h->stream3;
        const int gy = h->opt[OPT_GATHER_ROWS] > 0 ? std::min(h->opt[OPT_GATHER_ROWS], h->Nn) : std::min(h->Nn, 65535);
        for (int b = 0; b < nblk; ++b) {
            const int t0 = (int)((long long)nbx * b / nblk), t1 = (int)((long long)nbx * (b + 1) / nblk);
            if ((rc = run_morison(h, P, ldP, false, false, t0, t1 - t0)) != JK_OK) return rc;
            CUDA_TRY(h, cudaEventRecord(h->ev_blk[b], s));
            CUDA_TRY(h, cudaStreamWaitEvent(s3, h->ev_blk[b], 0));
            k_rhs_gather<<<dim3(t1 - t0, gy), PH_TPB, 0, s3>>>(h->Nn, ldP, h->n_pad, h->d_Fm, h->d_adj_ptr, h->d_adj, h->d_node2slot, h->d_Fstatic,
                                                               h->d_X, h->d_Ffix, nullptr, nullptr, nullptr, 0, 1, 0, t0 * PH_TPB);
            LAUNCH_CHECK(h);
        }
        CUDA_TRY(h, cudaEventRecord(h->ev_gather, s3));
        gathered = true;
    } else if ((rc = run_morison(h, P, ldP, false, fuse)) != JK_OK) return rc;
    const bool totals_early = fem && h->opt[OPT_EARLY_TOTALS] && h->stream3 != nullptr && h->ev_mor && h->ev_tot;
    if (totals_early && (rc = reduce_totals_early(h, P, ldP)) != JK_OK) return rc;
    h->last_fused = fuse;
    if (fem && fuse) {
        h->ev_set[JK_T_RHS] = false;                       // no separate load stage: the Morison kernel wrote the right-hand sides
        if ((rc = run_fem(h, ldP, fy)) != JK_OK) return rc;
    } else if (fem && gathered) {
        tic(h, JK_T_RHS);                                  // what is left of the last block's gather
        CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_gather, 0));
        toc(h, JK_T_RHS);
        if ((rc = run_fem(h, ldP, fy)) != JK_OK) return rc;
    } else if (fem) {
        tic(h, JK_T_RHS);
        dim3 g(nbx, std::min(h->Nn, 65535));
        k_rhs_gather<<<g, PH_TPB, 0, s>>>(h->Nn, ldP, h->n_pad, h->d_Fm, h->d_adj_ptr, h->d_adj, h->d_node2slot, h->d_Fstatic,
                                         h->d_X, h->d_Ffix, nullptr);
        LAUNCH_CHECK(h);
        toc(h, JK_T_RHS);
        if ((rc = run_fem(h, ldP, fy)) != JK_OK) return rc;
    }
    if ((rc = reduce_and_argmax(h, P, ldP, true, fem, totals_early)) != JK_OK) return rc;
    h->lastP = P; h->last_ldP = ldP; h->last_morison = true; h->last_fem = fem; h->last_fy = fy; h->last_fdir = false;
    return JK_OK;
}

static int check_scan_ready(jk_handle_t h, int P, const void* t, bool fem) {
    if (P <= 0 || !t) JK_FAIL(h, JK_EINVAL, "phase scan: P must be positive and t non-NULL (P=%d)", P);
    if (!h->have_wave || !h->have_morison) JK_FAIL(h, JK_ESTATE, "phase scan: call jk_set_wave_* and jk_set_morison first");
    if (fem && !h->factored) JK_FAIL(h, JK_ESTATE, "phase scan: call jk_assemble and jk_factor first");
    return JK_OK;
}

extern "C" int jk_read_table(jk_handle_t h, int P, double* table, int64_t* critical) {
    if (!h) return JK_EINVAL;
    if (P != h->lastP) JK_FAIL(h, JK_EINVAL, "jk_read_table: P=%d does not match the last scan (%d)", P, h->lastP);
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    long long idx = -1;
    const size_t tb = table ? (size_t)P * JK_TABLE_NCOL * sizeof(double) : 0;
    if (unsigned char* pin = pin_acquire(h, tb + 16)) {
        // table, critical index and the factorisation's pivot flag in one pinned buffer, one synchronisation
        { int rcj = join_factor(h, s); if (rcj != JK_OK) return rcj; }   // the pivot flag is final behind the factorisation
        tic(h, JK_T_D2H);
        if (table) CUDA_TRY(h, cudaMemcpyAsync(pin, h->d_table, tb, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaMemcpyAsync(pin + tb, h->d_argidx, sizeof(long long), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaMemcpyAsync(pin + tb + 8, h->d_info_sticky, sizeof(int), cudaMemcpyDeviceToHost, s));
        toc(h, JK_T_D2H);
        CUDA_TRY(h, cudaStreamSynchronize(s));
        if (table) memcpy(table, pin, tb);
        memcpy(&idx, pin + tb, sizeof(long long));
        int info = 0;
        memcpy(&info, pin + tb + 8, sizeof(int));
        if (critical) *critical = (int64_t)idx;
        if (h->factor_inflight) { h->factor_inflight = false; CUDA_TRY(h, cudaStreamSynchronize(h->stream2)); }
        {
            if (info != 0) {
                CUDA_TRY(h, cudaMemset(h->d_info_sticky, 0, sizeof(int)));
                h->factored = false;
                JK_FAIL(h, JK_ENOTSPD, "K_ff is not positive definite (pivot %d <= 0): structure is a mechanism or badly supported", info - 1);
            }
        }
        return JK_OK;
    }
    tic(h, JK_T_D2H);
    if (table) CUDA_TRY(h, cudaMemcpyAsync(table, h->d_table, (size_t)P * JK_TABLE_NCOL * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(&idx, h->d_argidx, sizeof(long long), cudaMemcpyDeviceToHost, s));
    toc(h, JK_T_D2H);
    CUDA_TRY(h, cudaStreamSynchronize(s));
    if (critical) *critical = (int64_t)idx;
    return finish_factor(h);
}

// phase times host -> d_t through the pinned staging buffer (asynchronous; the caller's array is free on return)
static int upload_times(jk_handle_t h, int P, const double* t) {
    tic(h, JK_T_H2D);
    if (h->pin_t_busy) { cudaEventSynchronize(h->ev_pin_t); h->pin_t_busy = false; }
    if ((size_t)P > h->pin_t_elems) {
        if (h->h_pin_t) { cudaFreeHost(h->h_pin_t); h->h_pin_t = nullptr; h->pin_t_elems = 0; }
        if (cudaMallocHost((void**)&h->h_pin_t, (size_t)P * sizeof(double)) == cudaSuccess) h->pin_t_elems = (size_t)P; else cudaGetLastError();
    }
    if (h->h_pin_t) {
        memcpy(h->h_pin_t, t, (size_t)P * sizeof(double));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_t, h->h_pin_t, (size_t)P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaEventRecord(h->ev_pin_t, h->stream));
        h->pin_t_busy = true;
    } else {
        CUDA_TRY(h, cudaMemcpyAsync(h->d_t, t, (size_t)P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    toc(h, JK_T_H2D);
    return JK_OK;
}

static int scan_host(jk_handle_t h, int P, const double* t, double fy, double* table, int64_t* critical, bool fem) {
    if (!h) return JK_EINVAL;
    int rc = check_scan_ready(h, P, t, fem);
    if (rc != JK_OK) return rc;
    cudaSetDevice(h->device);
    if ((rc = ensure_buffers(h, P, fem, false)) != JK_OK) return rc;
    tic(h, JK_T_SCAN_TOTAL);
    if ((rc = upload_times(h, P, t)) != JK_OK) return rc;
    if ((rc = scan_core(h, P, fy, fem)) != JK_OK) return rc;
    toc(h, JK_T_SCAN_TOTAL);
    if (!table && !critical) return JK_OK;      // asynchronous form: the results stay in HBM (jk_read_table / jk_table_dev)
    return jk_read_table(h, P, table, critical);
}

extern "C" int jk_morison_scan(jk_handle_t h, int P, const double* t, double* table, int64_t* critical) {
    return scan_host(h, P, t, 355.0, table, critical, false);
}

extern "C" int jk_phase_scan(jk_handle_t h, int P, const double* t, double fy, double* table, int64_t* critical) {
    if (h && !(fy > 0)) JK_FAIL(h, JK_EINVAL, "jk_phase_scan: fy must be positive");
    return scan_host(h, P, t, fy, table, critical, true);
}

extern "C" int jk_phase_scan_dev(jk_handle_t h, int P, const double* t_dev, double fy) {
    if (!h) return JK_EINVAL;
    int rc = check_scan_ready(h, P, t_dev, true);
    if (rc != JK_OK) return rc;
    if (!(fy > 0)) JK_FAIL(h, JK_EINVAL, "jk_phase_scan_dev: fy must be positive");
    cudaSetDevice(h->device);
    if ((rc = ensure_buffers(h, P, true, false)) != JK_OK) return rc;
    tic(h, JK_T_SCAN_TOTAL);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_t, t_dev, (size_t)P * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = scan_core(h, P, fy, true)) != JK_OK) return rc;
    toc(h, JK_T_SCAN_TOTAL);
    return JK_OK;
}

// ------------------------------------------------------------------------------------------------
// One whole analysis step of a resident loop: assemble + asynchronous factorisation + phase scan.  The ~30 launches on three
// streams (with their event joins and start gates) are captured once into a CUDA graph and replayed while nothing the graph
// bakes in has changed (geometry, supports, wave, Morison set-up, moduli, fy, phase count, options, buffers).
// ------------------------------------------------------------------------------------------------
static int step_launch(jk_handle_t h, int P, double fy, double E, double G) {
    int rc;
    if ((rc = assemble_launch(h, E, G)) != JK_OK) return rc;
    if ((rc = factor_begin_launch(h)) != JK_OK) return rc;
    return scan_core(h, P, fy, true);
}

static int step_core(jk_handle_t h, int P, double fy, double E, double G) {
    int rc;
    if ((rc = ensure_member_consts(h)) != JK_OK) return rc;
    if ((rc = ensure_gauss_tables(h)) != JK_OK) return rc;
    if ((rc = join_factor(h, h->stream)) != JK_OK) return rc;                 // a factorisation queued by an earlier eager call
    const bool want_graph = h->opt[OPT_CUDA_GRAPH] && !h->graph_unsupported && h->tma_sweep && h->wait_value32 != nullptr &&
                            !h->opt[OPT_PROFILE_CHOL] && !h->opt[OPT_PROFILE_SWEEP] && h->wave_kind == 0;
    if (!want_graph) return step_launch(h, P, fy, E, G);
    const bool hit = h->graph_exec != nullptr && h->graph_epoch_built == h->graph_epoch && h->graph_P == P && h->graph_E == E &&
                     h->graph_G == G && h->graph_fy == fy;
    if (!hit) {
        if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
        const long long launches0 = h->launches;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); h->graph_unsupported = true; return step_launch(h, P, fy, E, G); }
        h->capturing = true;
        rc = step_launch(h, P, fy, E, G);
        h->capturing = false;
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (rc != JK_OK || ce != cudaSuccess || graph == nullptr) {
            // something in the step cannot be captured on this driver (e.g. the stream memory operations of the start gates):
            // run the step the ordinary way from now on
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            h->graph_unsupported = true;
            h->launches = launches0;
            h->factor_inflight = false;
            return step_launch(h, P, fy, E, G);
        }
        ce = cudaGraphInstantiate(&h->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { cudaGetLastError(); h->graph_exec = nullptr; h->graph_unsupported = true; h->launches = launches0; h->factor_inflight = false; return step_launch(h, P, fy, E, G); }
        h->graph_epoch_built = h->graph_epoch; h->graph_P = P; h->graph_E = E; h->graph_G = G; h->graph_fy = fy;
        h->graph_launches = h->launches - launches0;
        h->graph_started_target = h->started_target;
        h->graph_fused = h->last_fused;
        h->launches = launches0;
    }
    CUDA_TRY(h, cudaGraphLaunch(h->graph_exec, h->stream));
    // host-side state exactly as the eager step leaves it
    h->launches += h->graph_launches;
    h->E = E; h->G = G;
    h->assembled = false; h->factored = true; h->factor_inflight = true; h->ev_bwd_external = false;
    h->gate2_armed = false; h->started_target = h->graph_started_target;
    const int ldP = ceil_div(P, SLAB) * SLAB;
    h->lastP = P; h->last_ldP = ldP; h->last_morison = true; h->last_fem = true; h->last_fy = fy; h->last_fdir = false;
    h->last_fused = h->graph_fused;
    return JK_OK;
}

static int check_step_ready(jk_handle_t h, int P, const void* t, double fy, double E, double G) {
    if (P <= 0 || !t) JK_FAIL(h, JK_EINVAL, "jk_step: P must be positive and t non-NULL (P=%d)", P);
    if (!h->have_supports) JK_FAIL(h, JK_ESTATE, "jk_step: call jk_set_supports first");
    if (!h->have_wave || !h->have_morison) JK_FAIL(h, JK_ESTATE, "jk_step: call jk_set_wave_* and jk_set_morison first");
    if (!(E > 0) || !(G > 0) || !(fy > 0)) JK_FAIL(h, JK_EINVAL, "jk_step: E, G and fy must be positive");
    return JK_OK;
}

extern "C" int jk_step_dev(jk_handle_t h, double E, double G, int P, const double* t_dev, double fy) {
    if (!h) return JK_EINVAL;
    int rc = check_step_ready(h, P, t_dev, fy, E, G);
    if (rc != JK_OK) return rc;
    cudaSetDevice(h->device);
    if ((rc = ensure_buffers(h, P, true, false)) != JK_OK) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_t, t_dev, (size_t)P * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return step_core(h, P, fy, E, G);
}

extern "C" int jk_step(jk_handle_t h, double E, double G, int P, const double* t, double fy, double* table, int64_t* critical) {
    if (!h) return JK_EINVAL;
    int rc = check_step_ready(h, P, t, fy, E, G);
    if (rc != JK_OK) return rc;
    cudaSetDevice(h->device);
    if ((rc = ensure_buffers(h, P, true, false)) != JK_OK) return rc;
    if ((rc = upload_times(h, P, t)) != JK_OK) return rc;
    if ((rc = step_core(h, P, fy, E, G)) != JK_OK) return rc;
    if (!table && !critical) return JK_OK;
    return jk_read_table(h, P, table, critical);
}

__global__ void k_nodal_single(int Nn, int ldP, int p, const double* __restrict__ Fm, const int* __restrict__ adj_ptr,
                               const int* __restrict__ adj, double* __restrict__ nodal) {
    int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= Nn) return;
    double f[3] = {0, 0, 0};
    for (int q = adj_ptr[node]; q < adj_ptr[node + 1]; ++q) {
        int m = adj[q] >> 1, end = adj[q] & 1;
        size_t o = ((size_t)m * 6 + 3 * end) * ldP + p;
        f[0] += Fm[o]; f[1] += Fm[o + ldP]; f[2] += Fm[o + 2 * (size_t)ldP];
    }
    nodal[3 * node] = f[0]; nodal[3 * node + 1] = f[1]; nodal[3 * node + 2] = f[2];
}

__global__ void k_gather_strided(int n, size_t stride, size_t offset, const double* __restrict__ src, double* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[(size_t)i * stride + offset];
}

__global__ void k_gather_U(int Nn, int p, int n_pad, const double* __restrict__ X, const int* __restrict__ node2slot, double* __restrict__ U) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * Nn) return;
    U[i] = load_u(X, node2slot, i / 6, i % 6, p, n_pad);
}

extern "C" int jk_morison_single(jk_handle_t h, double t, double* nodal_forces, double* totals, double* details) {
    if (!h) return JK_EINVAL;
    int rc = check_scan_ready(h, 1, &t, false);
    if (rc != JK_OK) return rc;
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    if ((rc = ensure_buffers(h, 1, false, details != nullptr)) != JK_OK) return rc;
    int ldP = SLAB;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_t, &t, sizeof(double), cudaMemcpyHostToDevice, s));
    if ((rc = ensure_member_consts(h)) != JK_OK) return rc;
    if ((rc = run_morison(h, 1, ldP, details != nullptr)) != JK_OK) return rc;
    int n_mchunk = ceil_div(h->M, MCHUNK);
    if ((rc = ensure_tmp(h, 3 * (size_t)h->Nn + (size_t)n_mchunk * 9 + 4 * (size_t)h->M)) != JK_OK) return rc;
    double* d_nodal = h->d_tmp;
    double* d_part = h->d_tmp + 3 * (size_t)h->Nn;
    double* d_det = d_part + (size_t)n_mchunk * 9;
    k_nodal_single<<<ceil_div(h->Nn, 128), 128, 0, s>>>(h->Nn, ldP, 0, h->d_Fm, h->d_adj_ptr, h->d_adj, d_nodal);
    LAUNCH_CHECK(h);
    // totals: chunk partials of phase 0, summed in chunk (= member) order on the host side of the ABI
    std::vector<double> part((size_t)n_mchunk * 9);
    k_gather_strided<<<ceil_div(n_mchunk * 9, 128), 128, 0, s>>>(n_mchunk * 9, (size_t)ldP, 0, h->d_totpart, d_part);
    LAUNCH_CHECK(h);
    CUDA_TRY(h, cudaMemcpyAsync(part.data(), d_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    if (totals) for (int k = 0; k < 9; ++k) { double v = 0.0; for (int c = 0; c < n_mchunk; ++c) v += part[(size_t)c * 9 + k]; totals[k] = v; }
    if (nodal_forces) CUDA_TRY(h, cudaMemcpyAsync(nodal_forces, d_nodal, 3 * (size_t)h->Nn * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (details) {
        k_gather_strided<<<ceil_div(h->M * 4, 128), 128, 0, s>>>(h->M * 4, (size_t)ldP, 0, h->d_details, d_det);
        LAUNCH_CHECK(h);
        CUDA_TRY(h, cudaMemcpyAsync(details, d_det, 4 * (size_t)h->M * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(h, cudaStreamSynchronize(s));
    h->lastP = 0;   // scan buffers no longer describe a scan
    return JK_OK;
}

extern "C" int jk_kinematics_points(jk_handle_t h, int n, const double* xyz, double t, double* out) {
    if (!h) return JK_EINVAL;
    if (n <= 0 || !xyz || !out) JK_FAIL(h, JK_EINVAL, "jk_kinematics_points: n must be positive, xyz and out non-NULL");
    if (!h->have_wave || !h->have_morison) JK_FAIL(h, JK_ESTATE, "jk_kinematics_points: call jk_set_wave_* and jk_set_morison first");
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    int rc;
    if ((rc = ensure_tmp(h, 13 * (size_t)n)) != JK_OK) return rc;
    double *d_xyz = h->d_tmp, *d_out = h->d_tmp + 3 * (size_t)n;
    CUDA_TRY(h, cudaMemcpyAsync(d_xyz, xyz, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    k_kinematics_points<<<ceil_div(n, 128), 128, 0, s>>>(n, d_xyz, t, launch_wave(h), h->n_harm, h->wave_kind == 1 ? h->d_four : nullptr, d_out);
    LAUNCH_CHECK(h);
    CUDA_TRY(h, cudaMemcpyAsync(out, d_out, 10 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return JK_OK;
}

// Sea-state ensemble: n_states Airy sea states x n_phase phases each = one batch of load cases on one factor.
// common front of the two ensemble entry points: argument checks, buffers, the optional direction loads
static int ensemble_prepare(jk_handle_t h, const char* who, int n_states, int n_phase, const double* F_dir, double fy, double** d_Fdir) {
    if (n_phase < 8 || n_phase > 4096) JK_FAIL(h, JK_EINVAL, "%s: n_phase must be in 8..4096 (got %d)", who, n_phase);
    if ((long long)n_states * n_phase > (1LL << 30)) JK_FAIL(h, JK_EINVAL, "%s: too many load cases", who);
    if (!h->have_wave || !h->have_morison || h->wave_kind != 0)
        JK_FAIL(h, JK_ESTATE, "%s: call jk_set_wave_airy (depth, current, dt) and jk_set_morison (current heading, coefficients) first", who);
    if (!h->factored) JK_FAIL(h, JK_ESTATE, "%s: call jk_assemble and jk_factor first", who);
    if (!(fy > 0)) JK_FAIL(h, JK_EINVAL, "%s: fy must be positive", who);
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    const int C = n_states * n_phase;
    int rc;
    if ((rc = ensure_buffers(h, C, true, false)) != JK_OK) return rc;
    if (n_states > h->cap_states) {
        CUDA_TRY(h, dev_alloc(&h->d_states, 5 * (size_t)n_states));
        CUDA_TRY(h, dev_alloc(&h->d_state_crit, (size_t)n_states));
        h->cap_states = n_states;
    }
    *d_Fdir = nullptr;
    if (F_dir) {
        size_t nF = 12 * (size_t)h->Nn;
        if (nF > h->fload_elems) { CUDA_TRY(h, dev_alloc(&h->d_Fload, nF)); h->fload_elems = nF; }
        *d_Fdir = h->d_Fload;
        CUDA_TRY(h, cudaMemcpyAsync(*d_Fdir, F_dir, nF * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    return JK_OK;
}

static int ensemble_run(jk_handle_t h, int n_states, int n_phase, double* d_Fdir, bool have_fdir, double fy, double* table,
                        int64_t* critical_per_state);

extern "C" int jk_ensemble_scan(jk_handle_t h, int n_states, int n_phase, const double* a, const double* k, const double* omega,
                                const double* theta_wave, const double* t, const double* F_dir, double fy, double* table,
                                int64_t* critical_per_state) {
    if (!h) return JK_EINVAL;
    if (n_states <= 0 || !a || !k || !omega || !theta_wave || !t) JK_FAIL(h, JK_EINVAL, "jk_ensemble_scan: empty or NULL sea-state arrays");
    int rc;
    double* d_Fdir;
    if ((rc = ensemble_prepare(h, "jk_ensemble_scan", n_states, n_phase, F_dir, fy, &d_Fdir)) != JK_OK) return rc;
    cudaStream_t s = h->stream;
    const int C = n_states * n_phase;
    std::vector<double> st(5 * (size_t)n_states);
    for (int i = 0; i < n_states; ++i) {
        if (!(k[i] > 0) || !(omega[i] > 0)) JK_FAIL(h, JK_EINVAL, "jk_ensemble_scan: state %d has non-positive k or omega", i);
        st[i] = a[i]; st[(size_t)n_states + i] = k[i]; st[2 * (size_t)n_states + i] = omega[i];
        st[3 * (size_t)n_states + i] = cos(theta_wave[i]); st[4 * (size_t)n_states + i] = sin(theta_wave[i]);
    }
    tic(h, JK_T_SCAN_TOTAL);
    tic(h, JK_T_H2D);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_states, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_t, t, (size_t)C * sizeof(double), cudaMemcpyHostToDevice, s));
    toc(h, JK_T_H2D);
    return ensemble_run(h, n_states, n_phase, d_Fdir, F_dir != nullptr, fy, table, critical_per_state);
}

extern "C" int jk_ensemble_scan_sea_states(jk_handle_t h, int n_states, int n_phase, const double* H, const double* T,
                                           const double* wave_dir_deg, double gravity, const double* F_dir, double fy,
                                           double* table, int64_t* critical_per_state, double* k_out) {
    if (!h) return JK_EINVAL;
    if (n_states <= 0 || !H || !T || !wave_dir_deg) JK_FAIL(h, JK_EINVAL, "jk_ensemble_scan_sea_states: empty or NULL sea-state arrays");
    if (!(gravity > 0)) JK_FAIL(h, JK_EINVAL, "jk_ensemble_scan_sea_states: gravity must be positive");
    int rc;
    double* d_Fdir;
    if ((rc = ensemble_prepare(h, "jk_ensemble_scan_sea_states", n_states, n_phase, F_dir, fy, &d_Fdir)) != JK_OK) return rc;
    cudaStream_t s = h->stream;
    const size_t S = (size_t)n_states;
    if ((rc = ensure_tmp(h, 3 * S + 1)) != JK_OK) return rc;
    double *dH = h->d_tmp, *dT = dH + S, *dD = dT + S;
    int* d_bad = reinterpret_cast<int*>(dD + S);
    tic(h, JK_T_SCAN_TOTAL);
    tic(h, JK_T_H2D);
    CUDA_TRY(h, cudaMemcpyAsync(dH, H, S * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(dT, T, S * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(dD, wave_dir_deg, S * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemsetAsync(d_bad, 0, sizeof(int), s));
    toc(h, JK_T_H2D);
    k_sea_state_setup<<<ceil_div(n_states, 128), 128, 0, s>>>(n_states, n_phase, dH, dT, dD, h->wv.d, gravity, h->d_states, h->d_t, d_bad);
    LAUNCH_CHECK(h);
    int bad = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (k_out) CUDA_TRY(h, cudaMemcpyAsync(k_out, h->d_states + S, S * sizeof(double), cudaMemcpyDeviceToHost, s));
    rc = ensemble_run(h, n_states, n_phase, d_Fdir, F_dir != nullptr, fy, table, critical_per_state);   // synchronises the stream
    if (bad) { h->lastP = 0; JK_FAIL(h, JK_EINVAL, "jk_ensemble_scan_sea_states: state %d has a negative height or a non-positive period", bad - 1); }
    return rc;
}

static int ensemble_run(jk_handle_t h, int n_states, int n_phase, double* d_Fdir, bool have_fdir, double fy, double* table,
                        int64_t* critical_per_state) {
    cudaStream_t s = h->stream;
    const int C = n_states * n_phase;
    const int ldC = ceil_div(C, SLAB) * SLAB;
    int rc;
    if ((rc = ensure_member_consts(h)) != JK_OK) return rc;
    WaveAiry w = launch_wave(h);
    tic(h, JK_T_MORISON);
    {
        dim3 grid(ceil_div(ldC, PH_TPB), ceil_div(h->M, MCHUNK));
        size_t smem = ens_smem_doubles(h->ng, n_phase) * sizeof(double);
        auto kern = (JK_ENSEMBLE_G15 && h->ng == 15) ? k_morison_ensemble<15> : k_morison_ensemble<0>;
        CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, PH_TPB, smem, s>>>(h->M, h->ng, C, ldC, n_states, n_phase, h->d_xyz, h->d_conn, h->d_mc, h->d_gsw,
                                                      h->d_states, h->d_t, w, 0.5 * h->rho * h->Cd, h->rho * h->Cm, h->d_Fm, h->d_totpart);
        LAUNCH_CHECK(h);
    }
    toc(h, JK_T_MORISON);
    tic(h, JK_T_RHS);
    {
        dim3 g(ceil_div(ldC, PH_TPB), std::min(h->Nn, 65535));
        k_rhs_gather<<<g, PH_TPB, 0, s>>>(h->Nn, ldC, h->n_pad, h->d_Fm, h->d_adj_ptr, h->d_adj, h->d_node2slot, h->d_Fstatic, h->d_X, h->d_Ffix, nullptr,
                                         d_Fdir, h->d_states, n_states, n_phase, C);
        LAUNCH_CHECK(h);
    }
    toc(h, JK_T_RHS);
    if ((rc = run_fem(h, ldC, fy)) != JK_OK) return rc;
    if ((rc = reduce_and_argmax(h, C, ldC, true, true, false, h->d_states + 2 * (size_t)n_states, n_phase)) != JK_OK) return rc;
    k_argmax_per_state<<<ceil_div(n_states, 128), 128, 0, s>>>(n_states, n_phase, h->d_table, JK_TABLE_NCOL, JK_COL_TOTAL_KN, h->d_state_crit);
    LAUNCH_CHECK(h);
    toc(h, JK_T_SCAN_TOTAL);
    h->lastP = C; h->last_ldP = ldC; h->last_morison = true; h->last_fem = true; h->last_fy = fy; h->last_fdir = have_fdir; h->last_fused = false;
    tic(h, JK_T_D2H);
    if (table) CUDA_TRY(h, cudaMemcpyAsync(table, h->d_table, (size_t)C * JK_TABLE_NCOL * sizeof(double), cudaMemcpyDeviceToHost, s));
    std::vector<long long> crit(n_states);
    if (critical_per_state) CUDA_TRY(h, cudaMemcpyAsync(crit.data(), h->d_state_crit, (size_t)n_states * sizeof(long long), cudaMemcpyDeviceToHost, s));
    toc(h, JK_T_D2H);
    CUDA_TRY(h, cudaStreamSynchronize(s));
    if (critical_per_state) for (int i = 0; i < n_states; ++i) critical_per_state[i] = (int64_t)crit[i];
    return finish_factor(h);
}

extern "C" int jk_solve(jk_handle_t h, int nrhs, const double* F, double fy) {
    if (!h) return JK_EINVAL;
    if (nrhs <= 0 || !F) JK_FAIL(h, JK_EINVAL, "jk_solve: nrhs must be positive and F non-NULL");
    if (!h->factored) JK_FAIL(h, JK_ESTATE, "jk_solve: call jk_assemble and jk_factor first");
    if (!(fy > 0)) JK_FAIL(h, JK_EINVAL, "jk_solve: fy must be positive");
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    int rc;
    if ((rc = ensure_buffers(h, nrhs, true, false)) != JK_OK) return rc;
    int ldP = ceil_div(nrhs, SLAB) * SLAB;
    size_t nF = (size_t)nrhs * 6 * h->Nn;
    if (nF > h->fload_elems) { CUDA_TRY(h, dev_alloc(&h->d_Fload, nF)); h->fload_elems = nF; }
    tic(h, JK_T_SCAN_TOTAL);
    tic(h, JK_T_H2D);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_Fload, F, nF * sizeof(double), cudaMemcpyHostToDevice, s));
    toc(h, JK_T_H2D);
    CUDA_TRY(h, cudaMemsetAsync(h->d_t, 0, (size_t)ldP * sizeof(double), s));
    tic(h, JK_T_RHS);
    dim3 g(ceil_div(ldP, PH_TPB), std::min(h->Nn, 65535));
    k_rhs_from_loads<<<g, PH_TPB, 0, s>>>(h->Nn, nrhs, ldP, h->n_pad, h->d_Fload, h->d_node2slot, h->d_X, h->d_Ffix);
    LAUNCH_CHECK(h);
    toc(h, JK_T_RHS);
    if ((rc = run_fem(h, ldP, fy)) != JK_OK) return rc;
    if ((rc = reduce_and_argmax(h, nrhs, ldP, false, true)) != JK_OK) return rc;
    toc(h, JK_T_SCAN_TOTAL);
    CUDA_TRY(h, cudaStreamSynchronize(s));
    h->lastP = nrhs; h->last_ldP = ldP; h->last_morison = false; h->last_fem = true; h->last_fy = fy; h->last_fdir = false; h->last_fused = false;
    return finish_factor(h);
}

extern "C" int jk_fetch_phase(jk_handle_t h, int phase, double* U, double* reactions, double* member_rows,
                              double* end_forces, double* nodal_forces) {
    if (!h) return JK_EINVAL;
    if (h->lastP <= 0) JK_FAIL(h, JK_ESTATE, "jk_fetch_phase: no scan or solve results are resident");
    if (phase < 0 || phase >= h->lastP) JK_FAIL(h, JK_EINVAL, "jk_fetch_phase: phase %d out of range (0..%d)", phase, h->lastP - 1);
    if ((U || reactions || member_rows || end_forces) && !h->last_fem) JK_FAIL(h, JK_ESTATE, "jk_fetch_phase: the last scan was Morison-only");
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    int ldP = h->last_ldP, rc;
    size_t nU = 6 * (size_t)h->Nn, nR = 6 * (size_t)h->n_fixed, nM = 7 * (size_t)h->M, nE = 12 * (size_t)h->M, nN = 3 * (size_t)h->Nn;
    if ((rc = ensure_tmp(h, nU + nR + nM + nE + nN)) != JK_OK) return rc;
    double *dU = h->d_tmp, *dR = dU + nU, *dM = dR + nR, *dE = dM + nM, *dN = dE + nE;
    if (U) {
        k_gather_U<<<ceil_div((int)nU, 128), 128, 0, s>>>(h->Nn, phase, h->n_pad, h->d_X, h->d_node2slot, dU); LAUNCH_CHECK(h);
        CUDA_TRY(h, cudaMemcpyAsync(U, dU, nU * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    if (reactions) {
        k_gather_strided<<<ceil_div((int)nR, 128), 128, 0, s>>>((int)nR, (size_t)ldP, (size_t)phase, h->d_react, dR); LAUNCH_CHECK(h);
        CUDA_TRY(h, cudaMemcpyAsync(reactions, dR, nR * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    if (member_rows || end_forces) {
        k_member_post_single<<<ceil_div(h->M, 128), 128, 0, s>>>(h->M, phase, h->n_pad, h->d_X, h->d_node2slot, h->d_conn, h->d_mc, h->sp,
                                                               h->last_fy, dM, dE);
        LAUNCH_CHECK(h);
        if (member_rows) CUDA_TRY(h, cudaMemcpyAsync(member_rows, dM, nM * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (end_forces) CUDA_TRY(h, cudaMemcpyAsync(end_forces, dE, nE * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    if (nodal_forces) {
        if (h->last_morison && h->last_fused) {
            k_nodal_from_partials<<<ceil_div(h->fuse_n_fin, 128), 128, 0, s>>>(h->fuse_n_fin, ldP, phase, h->d_fuse_fin_node, h->d_fuse_fin_rowptr, h->d_fuse_fin_rows, h->d_part, dN);
            LAUNCH_CHECK(h);
        } else if (h->last_morison) { k_nodal_single<<<ceil_div(h->Nn, 128), 128, 0, s>>>(h->Nn, ldP, phase, h->d_Fm, h->d_adj_ptr, h->d_adj, dN); LAUNCH_CHECK(h); }
        else CUDA_TRY(h, cudaMemsetAsync(dN, 0, nN * sizeof(double), s));
        CUDA_TRY(h, cudaMemcpyAsync(nodal_forces, dN, nN * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return JK_OK;
}

extern "C" int jk_fetch_member_column(jk_handle_t h, int member, int column, int P, double* out) {
    if (!h || !out) return JK_EINVAL;
    if (!h->last_fem || P != h->lastP) JK_FAIL(h, JK_ESTATE, "jk_fetch_member_column: no matching scan results are resident");
    if (member < 0 || member >= h->M || column < 0 || column >= JK_MEMBER_NCOL) JK_FAIL(h, JK_EINVAL, "jk_fetch_member_column: bad member/column");
    cudaSetDevice(h->device);
    CUDA_TRY(h, cudaMemcpyAsync(out, h->d_rows + ((size_t)member * JK_MEMBER_NCOL + column) * h->last_ldP, (size_t)P * sizeof(double),
                                cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return JK_OK;
}

// ------------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------------
extern "C" int jk_get_dims(jk_handle_t h, int32_t* out) {
    if (!h || !out) return JK_EINVAL;
    out[0] = h->Nn; out[1] = h->M; out[2] = h->n_fixed; out[3] = h->n_free; out[4] = h->n_pad; out[5] = NB; out[6] = h->bw; out[7] = h->NT;
    out[8] = h->hb; out[9] = h->n_chains; out[10] = h->ch[0].kS; out[11] = h->nS_nodes;
    return JK_OK;
}

extern "C" int jk_sweep_program(int n_tiles, int band_tiles, int kx, int backward, const int32_t* first_tile, int32_t* items, int cap_items,
                                int32_t* meta) {
    if (n_tiles <= 0 || band_tiles < 0 || band_tiles > SW_MAX_BW || kx < 0 || kx > n_tiles) return JK_EINVAL;
    std::vector<uint4> prog;
    int pre_row = 0, npre = 0, ktop = 0;
    build_sweep_program(n_tiles, band_tiles, kx, backward != 0, first_tile, prog, pre_row, npre, ktop);
    const int n = (int)prog.size() / SW_ITEM_U4;
    if (meta) { meta[0] = pre_row; meta[1] = npre; meta[2] = ktop; }
    if (items) {
        if (cap_items < n) return JK_EINVAL;
        for (int i = 0; i < n; ++i) {
            const uint4 a = prog[(size_t)i * SW_ITEM_U4], b = prog[(size_t)i * SW_ITEM_U4 + 1];
            items[6 * i + 0] = (int)a.x; items[6 * i + 1] = (int)a.y; items[6 * i + 2] = (int)a.z; items[6 * i + 3] = (int)a.w;
            items[6 * i + 4] = (int)b.x; items[6 * i + 5] = (int)b.y;
        }
    }
    return n;
}

// non-zeros of the factor: chain tiles (rows < row_end, columns < col_end DOFs), lower triangle incl. diagonal
__global__ void k_count_nnz(const double* __restrict__ tiles, int NT, int bw, int row_end, int col_end, unsigned long long* __restrict__ out) {
    const int I = blockIdx.x, dJ = blockIdx.y, J = I - dJ;
    if (J < 0) return;
    const double* g = tiles + tile_off(I, J, bw);
    unsigned cnt = 0;
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        int r = I * NB + idx / NB, c = J * NB + idx % NB;
        if (c <= r && r < row_end && c < col_end && g[idx] != 0.0) ++cnt;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, (unsigned long long)cnt);
}

extern "C" int jk_solver_stats(jk_handle_t h, double* out) {
    if (!h || !out) return JK_EINVAL;
    if (!h->factored) JK_FAIL(h, JK_ESTATE, "jk_solver_stats: factor first");
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    { int rcj = join_factor(h, s); if (rcj != JK_OK) return rcj; }
    unsigned long long* d_cnt = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&d_cnt, sizeof(unsigned long long)));
    CUDA_TRY(h, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), s));
    for (int c = 0; c < h->n_chains; ++c) {
        auto& chn = h->ch[c];
        // second chain: its separator columns belong to the first chain's factor
        const int col_end = (c == 1) ? chn.kS * NB : chn.n_rows;
        k_count_nnz<<<dim3(chn.NT, chn.bw + 1), 256, 0, s>>>(chn.d_tiles, chn.NT, chn.bw, chn.n_rows, col_end, d_cnt);
        LAUNCH_CHECK(h);
    }
    unsigned long long cnt = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    double exec = 0.0; long long items = 0;
    if (h->tma_sweep) {
        for (int c = 0; c < h->n_chains; ++c)
            for (int d = 0; d < 2; ++d) {
                auto& w = h->ch[c].sw[d];
                if (w.n_items == 0) continue;
                std::vector<uint4> prog((size_t)w.n_items * SW_ITEM_U4);
                CUDA_TRY(h, cudaMemcpyAsync(prog.data(), w.d_prog, prog.size() * sizeof(uint4), cudaMemcpyDeviceToHost, s));
                CUDA_TRY(h, cudaStreamSynchronize(s));
                for (int i = 0; i < w.n_items; ++i) {
                    const uint4 m = prog[(size_t)i * SW_ITEM_U4 + 2];
                    exec += 2.0 * 32.0 * (double)(__builtin_popcount(m.x) + __builtin_popcount(m.y) + __builtin_popcount(m.z) + __builtin_popcount(m.w));
                }
                items += w.n_items;
            }
    } else {
        for (int c = 0; c < h->n_chains; ++c) {
            auto& chn = h->ch[c];
            long long prods = 0;
            for (int k = 0; k < chn.NT; ++k) prods += 1 + std::min(k, chn.bw);
            exec += 2.0 * 2.0 * (double)prods * NB * NB;   // two sweeps, 2 flops per MAC, per right-hand-side column
        }
    }
    CUDA_TRY(h, cudaStreamSynchronize(s));
    cudaFree(d_cnt);
    out[0] = (double)cnt; out[1] = exec; out[2] = (double)items; out[3] = h->tma_sweep ? 1.0 : 0.0;
    out[4] = (double)h->nnz_env_min; out[5] = (double)h->sweep_slab_last;
    out[6] = h->graph_unsupported ? -1.0 : (h->graph_exec != nullptr ? 1.0 : 0.0); out[7] = (double)h->graph_launches;
    return JK_OK;
}

extern "C" int jk_get_order(jk_handle_t h, int32_t* free_nodes) {
    if (!h || !free_nodes) return JK_EINVAL;
    if (!h->have_supports) JK_FAIL(h, JK_ESTATE, "jk_get_order: supports not set");
    std::copy(h->h_free_nodes.begin(), h->h_free_nodes.end(), free_nodes);
    return JK_OK;
}

extern "C" int jk_get_K(jk_handle_t h, double* K) {
    if (!h || !K) return JK_EINVAL;
    if (!(h->E > 0)) JK_FAIL(h, JK_ESTATE, "jk_get_K: call jk_assemble first");
    cudaSetDevice(h->device);
    size_t n = 6 * (size_t)h->Nn;
    if (n * n > ((size_t)1 << 31)) JK_FAIL(h, JK_EINVAL, "jk_get_K: dense K of %zu x %zu is refused (use the tile storage)", n, n);
    int rc;
    if ((rc = ensure_tmp(h, n * n)) != JK_OK) return rc;
    cudaStream_t s = h->stream;
    CUDA_TRY(h, cudaMemsetAsync(h->d_tmp, 0, n * n * sizeof(double), s));
    size_t tot = (size_t)h->M * 144;
    k_dense_K<<<(unsigned)((tot + 127) / 128), 128, 0, s>>>(h->M, h->d_conn, h->d_Ke, h->d_tmp, (int)n);
    LAUNCH_CHECK(h);
    CUDA_TRY(h, cudaMemcpyAsync(K, h->d_tmp, n * n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return JK_OK;
}

extern "C" int jk_get_elements(jk_handle_t h, double* Ke, double* Kl, double* R, double* L) {
    if (!h) return JK_EINVAL;
    if (!(h->E > 0)) JK_FAIL(h, JK_ESTATE, "jk_get_elements: call jk_assemble first");
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    if (Ke) CUDA_TRY(h, cudaMemcpyAsync(Ke, h->d_Ke, (size_t)h->M * 144 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (Kl) CUDA_TRY(h, cudaMemcpyAsync(Kl, h->d_Kl, (size_t)h->M * 144 * sizeof(double), cudaMemcpyDeviceToHost, s));
    std::vector<double> mc;
    if (R || L) {
        mc.resize((size_t)h->M * MC_STRIDE);
        CUDA_TRY(h, cudaMemcpyAsync(mc.data(), h->d_mc, mc.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(h, cudaStreamSynchronize(s));
    for (int m = 0; m < h->M && (R || L); ++m) {
        if (R) for (int i = 0; i < 9; ++i) R[(size_t)m * 9 + i] = mc[(size_t)m * MC_STRIDE + MC_R + i];
        if (L) L[m] = mc[(size_t)m * MC_STRIDE + MC_L];
    }
    return JK_OK;
}

extern "C" int jk_get_timings(jk_handle_t h, double* ms) {
    if (!h || !ms) return JK_EINVAL;
    cudaSetDevice(h->device);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < JK_NTIMERS; ++i) {
        ms[i] = -1.0;
        if (h->ev_set[i]) { float f = 0; if (cudaEventElapsedTime(&f, h->ev0[i], h->ev1[i]) == cudaSuccess) ms[i] = f; }
    }
    return JK_OK;
}

// max |K u - F| over the free DOFs of all phases of the last scan, relative to max |F|
__global__ void __launch_bounds__(PH_TPB)
k_free_residual(int n_free_nodes, const int* __restrict__ free_nodes, int P, int ldP, int n_pad, const double* __restrict__ X,
                const int* __restrict__ node2slot, const int* __restrict__ conn, const int* __restrict__ adj_ptr,
                const int* __restrict__ adj, const double* __restrict__ Ke, const double* __restrict__ Fstatic,
                const double* __restrict__ Fm /* null: loads from Fload */, const double* __restrict__ Fload, int Nn,
                unsigned long long* __restrict__ out /* [2]: max |r|, max |F| as double bits */,
                const int* __restrict__ fin_of_node = nullptr /* fused scan: Morison loads = sum of the node's partial rows */,
                const int* __restrict__ fin_rowptr = nullptr, const int* __restrict__ fin_rows = nullptr, const double* __restrict__ part = nullptr) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    double rmax = 0.0, fmaxv = 0.0;
    for (int sidx = blockIdx.y; sidx < n_free_nodes && p < P; sidx += gridDim.y) {        // grid.y is capped at 65,535
        int node = free_nodes[sidx];
        double r[6] = {0, 0, 0, 0, 0, 0}, f[6];
        for (int c = 0; c < 6; ++c) f[c] = (Fm || part) ? Fstatic[6 * node + c] : Fload[(size_t)p * 6 * Nn + 6 * node + c];
        if (part) {
            const int e = fin_of_node[node];
            for (int r = fin_rowptr[e]; r < fin_rowptr[e + 1]; ++r)
                for (int c = 0; c < 3; ++c) f[c] += part[((size_t)fin_rows[r] * 3 + c) * ldP + p];
        }
        for (int q = adj_ptr[node]; q < adj_ptr[node + 1]; ++q) {
            int m = adj[q] >> 1, end = adj[q] & 1;
            double ue[12];
            for (int k = 0; k < 6; ++k) {
                ue[k] = load_u(X, node2slot, conn[2 * m], k, p, n_pad);
                ue[6 + k] = load_u(X, node2slot, conn[2 * m + 1], k, p, n_pad);
            }
            const double* ke = Ke + (size_t)m * 144 + (size_t)(6 * end) * 12;
            for (int i = 0; i < 6; ++i) { double s = 0.0; for (int j = 0; j < 12; ++j) s = fma(ke[i * 12 + j], ue[j], s); r[i] += s; }
            if (Fm) { size_t o = ((size_t)m * 6 + 3 * end) * ldP + p; f[0] += Fm[o]; f[1] += Fm[o + ldP]; f[2] += Fm[o + 2 * (size_t)ldP]; }
        }
        for (int i = 0; i < 6; ++i) { rmax = fmax(rmax, fabs(r[i] - f[i])); fmaxv = fmax(fmaxv, fabs(f[i])); }
    }
    for (int off = 16; off > 0; off >>= 1) {
        rmax = fmax(rmax, __shfl_down_sync(0xffffffffu, rmax, off));
        fmaxv = fmax(fmaxv, __shfl_down_sync(0xffffffffu, fmaxv, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&out[0], (unsigned long long)__double_as_longlong(rmax));
        atomicMax(&out[1], (unsigned long long)__double_as_longlong(fmaxv));
    }
}

extern "C" int jk_residual(jk_handle_t h, double* rel) {
    if (!h || !rel) return JK_EINVAL;
    if (!h->last_fem || h->lastP <= 0) JK_FAIL(h, JK_ESTATE, "jk_residual: no FEM results are resident");
    if (h->last_fdir) JK_FAIL(h, JK_ESTATE, "jk_residual: not available after an ensemble scan with heading-dependent loads (F_dir)");
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    CUDA_TRY(h, cudaMemsetAsync(h->d_res, 0, 2 * sizeof(double), s));
    dim3 g(ceil_div(h->lastP, PH_TPB), std::min(h->n_free_nodes, 65535));
    k_free_residual<<<g, PH_TPB, 0, s>>>(h->n_free_nodes, h->d_free_nodes, h->lastP, h->last_ldP, h->n_pad, h->d_X, h->d_node2slot,
                                         h->d_conn, h->d_adj_ptr, h->d_adj, h->d_Ke, h->d_Fstatic,
                                         (h->last_morison && !h->last_fused) ? h->d_Fm : nullptr, h->d_Fload, h->Nn, (unsigned long long*)h->d_res,
                                         h->d_fuse_fin_of_node, h->d_fuse_fin_rowptr, h->d_fuse_fin_rows, (h->last_morison && h->last_fused) ? h->d_part : nullptr);
    LAUNCH_CHECK(h);
    double v[2];
    CUDA_TRY(h, cudaMemcpyAsync(v, h->d_res, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    *rel = v[1] > 0 ? v[0] / v[1] : v[0];
    return JK_OK;
}
