// swmhd_api.cu — the C ABI of include/swmhd.h over the CUDA kernels.
// Host-side orchestration only: buffers, streams, launch sequencing, the NCCL halo exchange, clock.
//
// A context drives one or more y-slabs:
//   * one slab on one GPU (the default),
//   * n_gpus slabs on n_gpus devices of THIS process (cfg.n_gpus > 1; ncclCommInitAll), or
//   * one slab of a ring of `world` processes (cfg.world > 1; ncclCommInitRank in swmhd_comm_init).
// Per substage and slab: the tile rows next to the slab edges on a high-priority stream, x wrap of
// the rows to be sent, ncclSend/ncclRecv of 3 contiguous parent rows per field and direction (one
// NCCL group per substage), the interior concurrently on the main stream, event join.
#include "../../include/swmhd.h"
#include "kparams.h"
#include "nccl_dyn.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <utility>
#include <vector>

using namespace swmhd;

namespace {

struct Slab {
    int dev = 0, index = 0;     // CUDA ordinal; position in the y ring (0 = south)
    int j0 = 0, Ny = 0;         // global row offset, rows owned
    int rows[4] = {0, 0, 0, 0};
    size_t len[4] = {0, 0, 0, 0};
    double *U[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    double *G[4] = {nullptr, nullptr, nullptr, nullptr};
    double *O[4] = {nullptr, nullptr, nullptr, nullptr};   // staging of the asynchronous field writer (lazy)
    CUtensorMap tmap[2][4];     // TMA descriptors of U[b][k]: 2-D (P x rows) FP64, box (TX+6) x (TY+6)
    CUtensorMap tmap_rb[2][4];  // same arrays, box of the row-blocked kernel
    double *d_partials = nullptr, *d_diag = nullptr, *d_stage = nullptr, *d_red = nullptr;
    int nblocks_diag = 0, ntiles = 0;
    cudaStream_t main = nullptr, edge = nullptr, copy = nullptr;
    bool own_streams = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_edge = nullptr, ev_main = nullptr, ev_out_ready = nullptr, ev_out_done = nullptr;
    int ntr = 0, n_south = 0, n_north = 0;   // tile rows (units of the 8-row launch granularity); edge groups
    ncclComm_t comm = nullptr;
    int pending_slot = -1;
    bool out_in_flight = false;
    // SWMHD_GUARD=1 (test hook: compute-sanitizer is not available on every pool): every device array sits between two
    // guard zones filled with a sentinel; swmhd_check_guards verifies that no kernel wrote outside its arrays
    std::vector<std::pair<double *, size_t>> guarded;     // (user pointer, user doubles) of every guarded allocation
    bool guard = false;                                   // decided once, at create
};

} // namespace

struct swmhd_ctx {
    swmhd_config cfg;
    int Nx, P;
    int nslabs_total;           // slabs in the y ring (n_gpus or world)
    std::vector<Slab> slabs;    // the slabs this process drives
    int use_tma, use_rb;
    int tx, ty;                 // 8-row launch granularity of the tile rows
    int edge_tr;                // tile rows per edge group (row-blocked kernel: one 16-row tile)
    int cur;                    // U[cur] = current state (all slabs alike)
    int diag_slots;
    bool comm_ready;            // NCCL communicators exist (or are not needed)
    bool in_substage;
    int last_stage;             // last completed stage of the step in flight (0: between steps)
    double pending_dt;
    int armed_slot;             // >= 0: the next stage-1 substage also produces diagnostics into this slot
    double time;
    int64_t iter, launches;
    double last_ms;
    std::string err;
};

static thread_local std::string g_create_err;

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency,
// so the library still loads on a machine without a driver).
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool encode_field_map(CUtensorMap *tm, double *base, int P, int rows, int box_w, int box_h) {
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) {
            cudaGetLastError();
            return false;
        }
        fn = (encode_tiled_fn)ptr;
    }
    cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)P * sizeof(double)};      // bytes between rows (multiple of 16)
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char buf_[512];                                                                   \
            snprintf(buf_, sizeof buf_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            ctx->err = buf_;                                                                  \
            return SWMHD_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

#define NK(call)                                                                              \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != ncclSuccess) {                                                              \
            char buf_[512];                                                                   \
            snprintf(buf_, sizeof buf_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, nccl_api()->GetErrorString(r_)); \
            ctx->err = buf_;                                                                  \
            return SWMHD_ERR_NCCL;                                                            \
        }                                                                                     \
    } while (0)

// a batch of CUDA events that is destroyed on every exit path of the function that owns it
struct EventBatch {
    std::vector<cudaEvent_t> ev;
    cudaError_t create(size_t n, unsigned flags) {
        ev.assign(n, nullptr);
        for (auto &e : ev) {
            cudaError_t rc = cudaEventCreateWithFlags(&e, flags);
            if (rc != cudaSuccess) return rc;
        }
        return cudaSuccess;
    }
    cudaEvent_t operator[](size_t i) const { return ev[i]; }
    ~EventBatch() { for (auto e : ev) if (e) cudaEventDestroy(e); }
};

static int fail(swmhd_ctx *ctx, int code, const char *msg) {
    if (ctx) ctx->err = msg; else g_create_err = msg;
    return code;
}

extern "C" int swmhd_abi_version(void) { return SWMHD_ABI_VERSION; }

extern "C" const char *swmhd_last_error(const swmhd_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

extern "C" int swmhd_split_rows(int Ny, int nslabs, int index, int *j0, int *ny) {
    if (nslabs < 1 || index < 0 || index >= nslabs || Ny < nslabs || !j0 || !ny) return SWMHD_ERR_ARG;
    const int base = Ny / nslabs, rem = Ny % nslabs;       // the first `rem` slabs take one row more
    *j0 = index * base + (index < rem ? index : rem);
    *ny = base + (index < rem ? 1 : 0);
    return SWMHD_OK;
}

static bool multi(const swmhd_ctx *ctx) { return ctx->nslabs_total > 1; }

// ---- guarded allocations (SWMHD_GUARD) -----------------------------------------------------------------------
constexpr size_t GUARD_DOUBLES = 4096;                     // 32 KB on each side: keeps the 16-byte alignment TMA needs
static const unsigned long long GUARD_WORD = 0x7ff8dead7ff8beefull;   // a quiet NaN with a recognisable payload
static bool guard_on() { const char *e = getenv("SWMHD_GUARD"); return e && *e && *e != '0'; }

static cudaError_t dev_alloc(Slab &s, double **out, size_t n) {
    if (!s.guard) return cudaMalloc(out, n * sizeof(double));
    double *base = nullptr;
    cudaError_t e = cudaMalloc(&base, (n + 2 * GUARD_DOUBLES) * sizeof(double));
    if (e != cudaSuccess) return e;
    std::vector<unsigned long long> g(GUARD_DOUBLES, GUARD_WORD);
    if ((e = cudaMemcpy(base, g.data(), GUARD_DOUBLES * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(base + GUARD_DOUBLES + n, g.data(), GUARD_DOUBLES * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    *out = base + GUARD_DOUBLES;
    s.guarded.emplace_back(*out, n);
    return cudaSuccess;
}
static void dev_free(const Slab &s, double *p) {
    if (!p) return;
    cudaFree(s.guard ? p - GUARD_DOUBLES : p);
}

// ---------------------------------------------------------------------------
static int create_slab(swmhd_ctx *ctx, Slab &s, std::string &why) {
    const swmhd_config &c = ctx->cfg;
    cudaError_t e;
    auto bad = [&](const char *what, cudaError_t er) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s (device %d): %s", what, s.dev, cudaGetErrorString(er));
        why = buf;
        return SWMHD_ERR_CUDA;
    };
    if ((e = cudaSetDevice(s.dev)) != cudaSuccess) return bad("cudaSetDevice", e);
    s.guard = guard_on();
    for (int k = 0; k < 4; k++) {
        s.rows[k] = s.Ny + 6 + ((k == SWMHD_V && c.topo_y == SWMHD_BOUNDED) ? 1 : 0);
        s.len[k] = (size_t)ctx->P * s.rows[k];
    }
    s.ntr = (s.Ny + ctx->ty - 1) / ctx->ty;
    // edge groups: whole tiles of the kernel that runs (one 16-row tile of the row-blocked kernel = two
    // 8-row tile rows), the north group starts on a tile boundary and holds at least the 3 rows to be sent
    const int et = ctx->edge_tr;
    s.n_south = et;
    int north0 = ((s.Ny - 3) / (et * ctx->ty)) * et;
    if (north0 < 0) north0 = 0;
    s.n_north = s.ntr - north0;
    for (int k = 0; k < 4; k++) {
        const size_t bytes = s.len[k] * sizeof(double);
        for (int b = 0; b < 2; b++) {
            if ((e = dev_alloc(s, &s.U[b][k], s.len[k])) != cudaSuccess) return bad("cudaMalloc state", e);
            if ((e = cudaMemset(s.U[b][k], 0, bytes)) != cudaSuccess) return bad("cudaMemset", e);
        }
        if ((e = dev_alloc(s, &s.G[k], s.len[k])) != cudaSuccess) return bad("cudaMalloc tendency", e);
        if ((e = cudaMemset(s.G[k], 0, bytes)) != cudaSuccess) return bad("cudaMemset", e);
    }
    for (int b = 0; b < 2 && ctx->use_tma; b++)
        for (int k = 0; k < 4 && ctx->use_tma; k++)
            if (!encode_field_map(&s.tmap[b][k], s.U[b][k], ctx->P, s.rows[k], ctx->tx + 6, ctx->ty + 6)) ctx->use_tma = 0;
    if (!ctx->use_tma) ctx->use_rb = 0;
    if (ctx->use_rb) {
        int rtx, rty;
        substage_rb_tile(&rtx, &rty);
        for (int b = 0; b < 2 && ctx->use_rb; b++)
            for (int k = 0; k < 4 && ctx->use_rb; k++)
                if (!encode_field_map(&s.tmap_rb[b][k], s.U[b][k], ctx->P, s.rows[k], rtx + 6, rty + 6)) ctx->use_rb = 0;
    }
    // per-tile diagnostic partials: one slot per (8-row tile row, tile column) of the kernel that runs stage 1
    const int diag_tiles_x = (ctx->use_rb && (substage_rb_stage_mask() & 1)) ? substage_rb_tiles_x(c.formulation, ctx->Nx)
                                                                              : (ctx->Nx + ctx->tx - 1) / ctx->tx;
    s.nblocks_diag = diag_blocks(ctx->Nx, s.Ny);
    s.ntiles = diag_tiles_x * s.ntr;
    if (s.ntiles > s.nblocks_diag) s.nblocks_diag = s.ntiles;   // d_partials serves both diag paths
    if ((e = dev_alloc(s, &s.d_partials, (size_t)s.nblocks_diag * NDIAG)) != cudaSuccess) return bad("cudaMalloc diag", e);
    if ((e = dev_alloc(s, &s.d_diag, (size_t)ctx->diag_slots * NDIAG)) != cudaSuccess) return bad("cudaMalloc diag", e);
    if ((e = dev_alloc(s, &s.d_stage, (size_t)diag_stage_doubles())) != cudaSuccess) return bad("cudaMalloc diag", e);
    if ((e = cudaMemset(s.d_stage, 0, (size_t)diag_stage_doubles() * sizeof(double))) != cudaSuccess) return bad("cudaMemset", e);   // ticket word = 0
    if ((e = dev_alloc(s, &s.d_red, (size_t)2 * ctx->diag_slots * NDIAG)) != cudaSuccess) return bad("cudaMalloc diag", e);
    int lo, hi;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if ((e = cudaStreamCreateWithPriority(&s.main, cudaStreamNonBlocking, lo)) != cudaSuccess) return bad("stream", e);
    if ((e = cudaStreamCreateWithPriority(&s.edge, cudaStreamNonBlocking, hi)) != cudaSuccess) return bad("stream", e);
    if ((e = cudaStreamCreateWithPriority(&s.copy, cudaStreamNonBlocking, lo)) != cudaSuccess) return bad("stream", e);
    s.own_streams = true;
    if ((e = cudaEventCreate(&s.ev0)) != cudaSuccess) return bad("event", e);
    if ((e = cudaEventCreate(&s.ev1)) != cudaSuccess) return bad("event", e);
    if ((e = cudaEventCreateWithFlags(&s.ev_edge, cudaEventDisableTiming)) != cudaSuccess) return bad("event", e);
    if ((e = cudaEventCreateWithFlags(&s.ev_main, cudaEventDisableTiming)) != cudaSuccess) return bad("event", e);
    if ((e = cudaEventCreateWithFlags(&s.ev_out_ready, cudaEventDisableTiming)) != cudaSuccess) return bad("event", e);
    if ((e = cudaEventCreateWithFlags(&s.ev_out_done, cudaEventDisableTiming)) != cudaSuccess) return bad("event", e);
    return SWMHD_OK;
}

extern "C" int swmhd_create(const swmhd_config *c, swmhd_ctx **out) {
    if (!c || !out) return fail(nullptr, SWMHD_ERR_ARG, "null argument");
    *out = nullptr;
    if (c->abi_version != SWMHD_ABI_VERSION) return fail(nullptr, SWMHD_ERR_ARG, "abi_version mismatch");
    if (c->Hx != 3 || c->Hy != 3) return fail(nullptr, SWMHD_ERR_ARG, "halo must be 3 (WENO5 default)");
    if (c->Nx < 8 || c->Ny < 8) return fail(nullptr, SWMHD_ERR_ARG, "Nx, Ny must be >= 8");
    if (c->topo_x != SWMHD_PERIODIC) return fail(nullptr, SWMHD_ERR_ARG, "topo_x must be Periodic");
    if (c->topo_y != SWMHD_PERIODIC && c->topo_y != SWMHD_BOUNDED) return fail(nullptr, SWMHD_ERR_ARG, "bad topo_y");
    if (c->formulation != SWMHD_JACOBIAN && c->formulation != SWMHD_DIVERGENCE) return fail(nullptr, SWMHD_ERR_ARG, "bad formulation");
    if (c->arith != SWMHD_ARITH_FAST && c->arith != SWMHD_ARITH_STRICT) return fail(nullptr, SWMHD_ERR_ARG, "bad arith");
    if (c->flags != 0) return fail(nullptr, SWMHD_ERR_ARG, "non-default Appendix-C flags are oracle-only; the CUDA path implements the defaults");
    if (!(c->dx > 0) || !(c->dy > 0) || !(c->weno_eps > 0)) return fail(nullptr, SWMHD_ERR_ARG, "dx, dy, weno_eps must be positive");
    if (c->world < 1 || c->rank < 0 || c->rank >= c->world) return fail(nullptr, SWMHD_ERR_ARG, "bad rank/world");
    if (c->slab_ny < 8 || c->slab_j0 < 0 || c->slab_j0 + c->slab_ny > c->Ny) return fail(nullptr, SWMHD_ERR_ARG, "bad slab (need >= 8 rows)");
    if (c->world == 1 && (c->slab_j0 != 0 || c->slab_ny != c->Ny)) return fail(nullptr, SWMHD_ERR_ARG, "world == 1 needs the full domain");
    const int ngpu = c->n_gpus > 1 ? c->n_gpus : 1;
    if (c->n_gpus < 0 || ngpu > SWMHD_MAX_GPUS) return fail(nullptr, SWMHD_ERR_ARG, "n_gpus must be in 0..8");
    if (ngpu > 1 && c->world != 1) return fail(nullptr, SWMHD_ERR_ARG, "n_gpus > 1 (single process) and world > 1 (one process per GPU) are exclusive");
    if (ngpu > 1 && c->Ny / ngpu < 8) return fail(nullptr, SWMHD_ERR_ARG, "n_gpus > 1 needs at least 8 rows per slab");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, SWMHD_ERR_NODEVICE, "no CUDA device: libswmhd_cuda has no CPU fallback");
    }
    if (ngpu == 1 && (c->device < 0 || c->device >= ndev)) return fail(nullptr, SWMHD_ERR_ARG, "bad device ordinal");
    for (int d = 0; d < ngpu && ngpu > 1; d++) {
        if (c->device_ids[d] < 0 || c->device_ids[d] >= ndev) return fail(nullptr, SWMHD_ERR_ARG, "bad device_ids entry");
        for (int d2 = 0; d2 < d; d2++)
            if (c->device_ids[d2] == c->device_ids[d]) return fail(nullptr, SWMHD_ERR_ARG, "device_ids must be distinct");
    }

    swmhd_ctx *ctx = new (std::nothrow) swmhd_ctx();
    if (!ctx) return fail(nullptr, SWMHD_ERR_ARG, "out of host memory");
    ctx->cfg = *c;
    ctx->Nx = c->Nx; ctx->P = c->Nx + 6;
    ctx->nslabs_total = ngpu > 1 ? ngpu : c->world;
    ctx->cur = 0; ctx->time = 0; ctx->iter = 0; ctx->launches = 0; ctx->last_ms = 0;
    ctx->in_substage = false; ctx->last_stage = 0; ctx->pending_dt = 0; ctx->armed_slot = -1;
    ctx->diag_slots = 1024;
    ctx->comm_ready = (ctx->nslabs_total == 1);
    substage_tile(&ctx->tx, &ctx->ty);
    // TMA tile loads need a 16-byte row pitch (even Nx); otherwise the kernels use plain loads
    ctx->use_tma = (ctx->P % 2 == 0) ? 1 : 0;
    if (getenv("SWMHD_NO_TMA")) ctx->use_tma = 0;
    // row-blocked kernels (FAST arithmetic): taller box; their 8-row launch granularity is ctx->ty
    ctx->use_rb = (ctx->use_tma && ctx->ty == 8 && c->arith == SWMHD_ARITH_FAST) ? 1 : 0;
    if (getenv("SWMHD_NO_RB")) ctx->use_rb = 0;
    {
        int rtx = 0, rty = 8;
        if (ctx->use_rb) substage_rb_tile(&rtx, &rty);
        ctx->edge_tr = (ctx->use_rb && substage_rb_stage_mask() == 7) ? rty / ctx->ty : 1;
        if (ctx->edge_tr < 1) ctx->edge_tr = 1;
    }
    ctx->slabs.resize(ngpu);
    for (int d = 0; d < ngpu; d++) {
        Slab &s = ctx->slabs[d];
        if (ngpu > 1) {
            s.dev = c->device_ids[d]; s.index = d;
            swmhd_split_rows(c->Ny, ngpu, d, &s.j0, &s.Ny);
        } else {
            s.dev = c->device; s.index = c->rank; s.j0 = c->slab_j0; s.Ny = c->slab_ny;
        }
    }
    for (auto &s : ctx->slabs) {
        std::string why;
        int rc = create_slab(ctx, s, why);
        if (rc != SWMHD_OK) {
            g_create_err = why;
            swmhd_destroy(ctx);
            return rc;
        }
    }
    if (ngpu > 1) {     // one communicator per device of this process
        const NcclApi *api = nccl_api();
        if (!api) { g_create_err = std::string("n_gpus > 1 needs NCCL: ") + nccl_load_error(); swmhd_destroy(ctx); return SWMHD_ERR_NCCL; }
        std::vector<ncclComm_t> comms(ngpu);
        std::vector<int> devs(ngpu);
        for (int d = 0; d < ngpu; d++) devs[d] = ctx->slabs[d].dev;
        ncclResult_t r = api->CommInitAll(comms.data(), ngpu, devs.data());
        if (r != ncclSuccess) { g_create_err = std::string("ncclCommInitAll: ") + api->GetErrorString(r); swmhd_destroy(ctx); return SWMHD_ERR_NCCL; }
        for (int d = 0; d < ngpu; d++) ctx->slabs[d].comm = comms[d];
        ctx->comm_ready = true;
    }
    *out = ctx;
    return SWMHD_OK;
}

extern "C" void swmhd_destroy(swmhd_ctx *ctx) {
    if (!ctx) return;
    for (auto &s : ctx->slabs) {
        cudaSetDevice(s.dev);
        cudaDeviceSynchronize();
        if (s.comm && nccl_api()) nccl_api()->CommDestroy(s.comm);
        for (int k = 0; k < 4; k++) {
            dev_free(s, s.U[0][k]); dev_free(s, s.U[1][k]); dev_free(s, s.G[k]); dev_free(s, s.O[k]);
        }
        dev_free(s, s.d_partials); dev_free(s, s.d_diag); dev_free(s, s.d_stage); dev_free(s, s.d_red);
        if (s.own_streams) {
            if (s.main) cudaStreamDestroy(s.main);
            if (s.edge) cudaStreamDestroy(s.edge);
        }
        if (s.copy) cudaStreamDestroy(s.copy);
        for (cudaEvent_t ev : {s.ev0, s.ev1, s.ev_edge, s.ev_main, s.ev_out_ready, s.ev_out_done})
            if (ev) cudaEventDestroy(ev);
        cudaGetLastError();
    }
    delete ctx;
}

// ---- NCCL ------------------------------------------------------------------------------------
extern "C" int swmhd_comm_unique_id(void *id, size_t nbytes) {
    if (!id || nbytes < SWMHD_COMM_ID_BYTES) return fail(nullptr, SWMHD_ERR_ARG, "id buffer must hold SWMHD_COMM_ID_BYTES");
    const NcclApi *api = nccl_api();
    if (!api) { g_create_err = std::string("NCCL not available: ") + nccl_load_error(); return SWMHD_ERR_NCCL; }
    static_assert(sizeof(ncclUniqueId) == SWMHD_COMM_ID_BYTES, "unique id size");
    ncclUniqueId uid;
    ncclResult_t r = api->GetUniqueId(&uid);
    if (r != ncclSuccess) { g_create_err = std::string("ncclGetUniqueId: ") + api->GetErrorString(r); return SWMHD_ERR_NCCL; }
    memcpy(id, &uid, sizeof uid);
    return SWMHD_OK;
}

extern "C" int swmhd_comm_init(swmhd_ctx *ctx, const void *id, size_t nbytes) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (!id || nbytes < SWMHD_COMM_ID_BYTES) return fail(ctx, SWMHD_ERR_ARG, "id buffer must hold SWMHD_COMM_ID_BYTES");
    if (ctx->cfg.world < 2) return fail(ctx, SWMHD_ERR_STATE, "swmhd_comm_init is for world > 1 contexts");
    if (ctx->comm_ready) return fail(ctx, SWMHD_ERR_STATE, "communicator already initialised");
    const NcclApi *api = nccl_api();
    if (!api) { ctx->err = std::string("NCCL not available: ") + nccl_load_error(); return SWMHD_ERR_NCCL; }
    Slab &s = ctx->slabs[0];
    CK(cudaSetDevice(s.dev));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    NK(api->CommInitRank(&s.comm, ctx->cfg.world, uid, ctx->cfg.rank));
    ctx->comm_ready = true;
    return SWMHD_OK;
}

// neighbours of slab `index` in the y ring: -1 at a wall
static void ring_neighbours(const swmhd_ctx *ctx, int index, int *south, int *north) {
    const int n = ctx->nslabs_total;
    const bool per = (ctx->cfg.topo_y == SWMHD_PERIODIC);
    *south = index > 0 ? index - 1 : (per ? n - 1 : -1);
    *north = index < n - 1 ? index + 1 : (per ? 0 : -1);
}

// One halo exchange of `nf` fields: base[slab][k] are parent arrays laid out like the state.  My north
// edge rows go to the north neighbour's south halo and vice versa; sends and receives towards one peer
// are issued in matching order (north rows first), which also covers a ring of two.
static int exchange(swmhd_ctx *ctx, double *(*base)[4], int nf, bool on_edge_stream) {
    if (!multi(ctx)) return SWMHD_OK;
    const NcclApi *api = nccl_api();
    const size_t n = (size_t)3 * ctx->P;
    NK(api->GroupStart());
    for (size_t si = 0; si < ctx->slabs.size(); si++) {
        Slab &s = ctx->slabs[si];
        int so, no;
        ring_neighbours(ctx, s.index, &so, &no);
        cudaStream_t st = on_edge_stream ? s.edge : s.main;
        for (int k = 0; k < nf; k++) {
            double *a = base[si][k];
            double *north_send = a + (size_t)s.Ny * ctx->P, *south_send = a + (size_t)3 * ctx->P;
            double *south_halo = a, *north_halo = a + (size_t)(s.Ny + 3) * ctx->P;
            if (no >= 0) NK(api->Send(north_send, n, ncclFloat64, no, s.comm, st));
            if (so >= 0) NK(api->Recv(south_halo, n, ncclFloat64, so, s.comm, st));
            if (so >= 0) NK(api->Send(south_send, n, ncclFloat64, so, s.comm, st));
            if (no >= 0) NK(api->Recv(north_halo, n, ncclFloat64, no, s.comm, st));
        }
    }
    NK(api->GroupEnd());
    return SWMHD_OK;
}

static int exchange_state(swmhd_ctx *ctx, int buf, bool on_edge_stream) {
    double *base[SWMHD_MAX_GPUS][4];
    for (size_t si = 0; si < ctx->slabs.size(); si++)
        for (int k = 0; k < 4; k++) base[si][k] = ctx->slabs[si].U[buf][k];
    return exchange(ctx, base, 4, on_edge_stream);
}

static int need_comm(swmhd_ctx *ctx) {
    if (multi(ctx) && !ctx->comm_ready)
        return fail(ctx, SWMHD_ERR_STATE, "world > 1: call swmhd_comm_init first (or drive the exchange with substage_edges/interior/finish)");
    return SWMHD_OK;
}

// ---- fields -------------------------------------------------------------------------------------
static bool in_process_multi(const swmhd_ctx *ctx) { return ctx->slabs.size() > 1; }

static size_t host_len(const swmhd_ctx *ctx, int field) {
    if (!in_process_multi(ctx)) return ctx->slabs[0].len[field];
    const int rows = ctx->cfg.Ny + 6 + ((field == SWMHD_V && ctx->cfg.topo_y == SWMHD_BOUNDED) ? 1 : 0);
    return (size_t)ctx->P * rows;
}

extern "C" size_t swmhd_field_len(const swmhd_ctx *ctx, int field) {
    if (!ctx || field < 0 || field > 3) return 0;
    return host_len(ctx, field);
}

extern "C" int swmhd_set_field(swmhd_ctx *ctx, int field, const double *host, size_t n) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (field < 0 || field > 3 || !host) return fail(ctx, SWMHD_ERR_ARG, "bad field/host");
    if (n != host_len(ctx, field)) return fail(ctx, SWMHD_ERR_ARG, "host buffer is not the parent array of this field (length mismatch)");
    // every slab takes its rows with their 3 halo rows on each side out of the (global) parent array
    for (auto &s : ctx->slabs) {
        CK(cudaSetDevice(s.dev));
        const size_t off = in_process_multi(ctx) ? (size_t)s.j0 * ctx->P : 0;
        CK(cudaMemcpyAsync(s.U[ctx->cur][field], host + off, s.len[field] * sizeof(double), cudaMemcpyHostToDevice, s.main));
    }
    for (auto &s : ctx->slabs) { CK(cudaSetDevice(s.dev)); CK(cudaStreamSynchronize(s.main)); }
    return SWMHD_OK;
}

// rows [lo, hi) of slab s that it contributes to a gathered parent array: its interior rows, the end slabs
// also the global halo rows (and the wall row of v|vh)
static void gather_rows(const swmhd_ctx *ctx, const Slab &s, int field, int *lo, int *hi) {
    *lo = (s.index == 0) ? 0 : 3;
    *hi = (s.index == ctx->nslabs_total - 1) ? s.rows[field] : 3 + s.Ny;
}

extern "C" int swmhd_get_field(swmhd_ctx *ctx, int field, double *host, size_t n) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (field < 0 || field > 3 || !host) return fail(ctx, SWMHD_ERR_ARG, "bad field/host");
    if (n != host_len(ctx, field)) return fail(ctx, SWMHD_ERR_ARG, "host buffer is not the parent array of this field (length mismatch)");
    for (auto &s : ctx->slabs) {
        CK(cudaSetDevice(s.dev));
        if (!in_process_multi(ctx)) {
            CK(cudaMemcpyAsync(host, s.U[ctx->cur][field], n * sizeof(double), cudaMemcpyDeviceToHost, s.main));
        } else {
            int lo, hi;
            gather_rows(ctx, s, field, &lo, &hi);
            CK(cudaMemcpyAsync(host + (size_t)(s.j0 + lo) * ctx->P, s.U[ctx->cur][field] + (size_t)lo * ctx->P,
                               (size_t)(hi - lo) * ctx->P * sizeof(double), cudaMemcpyDeviceToHost, s.main));
        }
    }
    for (auto &s : ctx->slabs) { CK(cudaSetDevice(s.dev)); CK(cudaStreamSynchronize(s.main)); }
    return SWMHD_OK;
}

// ---------------------------------------------------------------------------
static HaloParams halo_params(swmhd_ctx *ctx, const Slab &s, double *const U[4], int j_lo, int j_hi, bool y_part) {
    const swmhd_config &c = ctx->cfg;
    HaloParams h;
    h.Nx = ctx->Nx; h.Ny = s.Ny; h.P = ctx->P;
    h.by = (c.topo_y == SWMHD_BOUNDED);
    h.first = (s.index == 0); h.last = (s.index == ctx->nslabs_total - 1);
    h.y_mode = 0;
    if (y_part) {
        if (h.by) h.y_mode = 2;
        else if (!multi(ctx)) h.y_mode = 1;
    }
    h.grad = c.A_gradient_bc; h.gs = c.A_grad_south; h.gn = c.A_grad_north; h.dy = c.dy;
    h.j_lo = j_lo; h.j_hi = j_hi;
    for (int k = 0; k < 4; k++) { h.U[k] = U[k]; h.rows[k] = s.rows[k]; }
    return h;
}

static KParams kparams(swmhd_ctx *ctx, const Slab &s, double dt, int stage) {
    const swmhd_config &c = ctx->cfg;
    KParams p;
    p.Nx = ctx->Nx; p.Ny = s.Ny; p.P = ctx->P;
    p.gj0 = s.j0; p.NyG = c.Ny; p.by = (c.topo_y == SWMHD_BOUNDED);
    p.tile_row0 = 0; p.tile_rows = s.ntr;
    for (int k = 0; k < 4; k++) p.rows[k] = s.rows[k];
    p.dx = c.dx; p.dy = c.dy; p.rdx = 1.0 / c.dx; p.rdy = 1.0 / c.dy; p.inv_az = 1.0 / (c.dx * c.dy);
    p.g = c.g; p.f = c.f; p.eps = c.weno_eps; p.h_ref = c.h_ref;
    p.dt = dt;
    p.gam = stage >= 1 ? RK_GAMMA[stage - 1] : 0.0;
    p.zet = stage >= 1 ? RK_ZETA[stage - 1] : 0.0;
    p.dtgam = dt * p.gam;
    p.dtzet = dt * p.zet;
    p.qrdxy = 0.25 * p.rdx * p.rdy;
    for (int k = 0; k < 4; k++) {
        p.Uo[k] = s.U[ctx->cur][k];
        p.Un[k] = s.U[1 - ctx->cur][k];
        p.G[k] = s.G[k];
    }
    p.diag = nullptr;
    p.use_tma = ctx->use_tma;
    if (ctx->use_tma)
        for (int k = 0; k < 4; k++) p.tm[k] = s.tmap[ctx->cur][k];
    p.use_rb = ctx->use_rb;
    p.row_begin = p.row_end = 0;
    if (ctx->use_rb)
        for (int k = 0; k < 4; k++) p.tm_rb[k] = s.tmap_rb[ctx->cur][k];
    return p;
}

static cudaError_t launch_substage(swmhd_ctx *ctx, const KParams &p, int stage, cudaStream_t st) {
    ctx->launches++;
    if (ctx->cfg.arith == SWMHD_ARITH_STRICT) return launch_substage_strict(p, ctx->cfg.formulation, stage, st);
    return launch_substage_fast(p, ctx->cfg.formulation, stage, st);
}

static void tick(swmhd_ctx *ctx, double dt, int stage) {
    // upstream tick!: first_stage_dt = g1*dt; then (g2+z2)*dt; then (g3+z3)*dt  
    double sdt = (stage == 1) ? RK_GAMMA[0] * dt : (RK_GAMMA[stage - 1] + RK_ZETA[stage - 1]) * dt;
    ctx->time += sdt;
    if (stage == 3) ctx->iter += 1;
    ctx->last_stage = stage % 3;
}

static int sync_all(swmhd_ctx *ctx) {
    for (auto &s : ctx->slabs) {
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.edge));
        CK(cudaStreamSynchronize(s.main));
    }
    return SWMHD_OK;
}

extern "C" int swmhd_fill_halos(swmhd_ctx *ctx) {
    if (!ctx) return SWMHD_ERR_ARG;
    for (auto &s : ctx->slabs) {
        CK(cudaSetDevice(s.dev));
        HaloParams h = halo_params(ctx, s, s.U[ctx->cur], 3, s.Ny + 2, true);
        ctx->launches++;
        CK(launch_halo(h, s.main));
    }
    if (multi(ctx) && ctx->comm_ready) {      // y halos from the neighbours (x-wrapped rows travel, so corners are right)
        int rc = exchange_state(ctx, ctx->cur, false);
        if (rc) return rc;
    }
    return sync_all(ctx);
}

// ---- one substage of a slab: edges, interior ---------------------------------------------------------
static int slab_edges(swmhd_ctx *ctx, Slab &s, double dt, int stage, cudaEvent_t e0) {
    CK(cudaSetDevice(s.dev));
    // the edge stream may not overwrite rows the previous substage's kernels still read
    CK(cudaEventRecord(s.ev_main, s.main));
    CK(cudaStreamWaitEvent(s.edge, s.ev_main, 0));
    if (e0) CK(cudaEventRecord(e0, s.edge));
    KParams p = kparams(ctx, s, dt, stage);
    s.pending_slot = -1;
    if (stage == 1 && ctx->armed_slot >= 0) {
        p.diag = s.d_partials;
        s.pending_slot = ctx->armed_slot;
    }
    const int ntr = s.ntr, ns = s.n_south, nn = s.n_north, ty = ctx->ty;
    double *const *Un = s.U[1 - ctx->cur];
    if (ntr <= ns + nn) {            // slab too thin to split: everything is "edge"
        CK(launch_substage(ctx, p, stage, s.edge));
        HaloParams h = halo_params(ctx, s, Un, 3, s.Ny + 2, true);
        ctx->launches++;
        CK(launch_halo(h, s.edge));
    } else {
        p.tile_row0 = 0; p.tile_rows = ns;
        CK(launch_substage(ctx, p, stage, s.edge));
        p.tile_row0 = ntr - nn; p.tile_rows = nn;
        CK(launch_substage(ctx, p, stage, s.edge));
        // x wrap of the rows that are about to be sent, and wall BCs on end slabs
        HaloParams h = halo_params(ctx, s, Un, 3, 3 + ns * ty - 1, true);
        ctx->launches++;
        CK(launch_halo(h, s.edge));
        HaloParams h2 = halo_params(ctx, s, Un, 3 + (ntr - nn) * ty, s.Ny + 2, false);
        ctx->launches++;
        CK(launch_halo(h2, s.edge));
    }
    return SWMHD_OK;
}

static int slab_interior(swmhd_ctx *ctx, Slab &s, double dt, int stage) {
    CK(cudaSetDevice(s.dev));
    const int ntr = s.ntr, ns = s.n_south, nn = s.n_north;
    if (ntr > ns + nn) {
        KParams p = kparams(ctx, s, dt, stage);
        if (stage == 1 && s.pending_slot >= 0) p.diag = s.d_partials;
        p.tile_row0 = ns; p.tile_rows = ntr - nn - ns;
        CK(launch_substage(ctx, p, stage, s.main));
        HaloParams h = halo_params(ctx, s, s.U[1 - ctx->cur], 3 + ns * ctx->ty, 3 + (ntr - nn) * ctx->ty - 1, false);
        ctx->launches++;
        CK(launch_halo(h, s.main));
    }
    return SWMHD_OK;
}

static int slab_join(swmhd_ctx *ctx, Slab &s, cudaEvent_t e1) {
    CK(cudaSetDevice(s.dev));
    CK(cudaEventRecord(s.ev_edge, s.edge));
    CK(cudaStreamWaitEvent(s.main, s.ev_edge, 0));
    if (s.pending_slot >= 0) {       // fold the per-tile partials of edges + interior (fixed order)
        ctx->launches++;
        CK(launch_diag_final(s.d_partials, s.ntiles, s.d_stage, s.d_diag + (size_t)s.pending_slot * NDIAG, s.main));
        s.pending_slot = -1;
    }
    if (e1) CK(cudaEventRecord(e1, s.main));
    return SWMHD_OK;
}

// one substage of every local slab, no host synchronisation.  Single slab: one launch on the main stream.
// Optional event pair around the fused substage kernel of slab 0 (single slab only).
static int substage_async(swmhd_ctx *ctx, double dt, int stage, cudaEvent_t e0 = nullptr, cudaEvent_t e1 = nullptr, int diag_slot = -1) {
    if (!multi(ctx)) {
        Slab &s = ctx->slabs[0];
        KParams p = kparams(ctx, s, dt, stage);
        if (diag_slot >= 0 && stage == 1) p.diag = s.d_partials;
        if (e0) CK(cudaEventRecord(e0, s.main));
        CK(launch_substage(ctx, p, stage, s.main));
        if (e1) CK(cudaEventRecord(e1, s.main));
        if (diag_slot >= 0 && stage == 1) {
            ctx->launches++;
            CK(launch_diag_final(s.d_partials, s.ntiles, s.d_stage, s.d_diag + (size_t)diag_slot * NDIAG, s.main));
        }
        HaloParams h = halo_params(ctx, s, s.U[1 - ctx->cur], 3, s.Ny + 2, true);
        ctx->launches++;
        CK(launch_halo(h, s.main));
    } else {
        ctx->armed_slot = (stage == 1) ? diag_slot : -1;
        for (auto &s : ctx->slabs) { int rc = slab_edges(ctx, s, dt, stage, nullptr); if (rc) return rc; }
        ctx->armed_slot = -1;
        int rc = exchange_state(ctx, 1 - ctx->cur, true);
        if (rc) return rc;
        for (auto &s : ctx->slabs) { rc = slab_interior(ctx, s, dt, stage); if (rc) return rc; }
        for (auto &s : ctx->slabs) { rc = slab_join(ctx, s, nullptr); if (rc) return rc; }
    }
    ctx->cur = 1 - ctx->cur;
    tick(ctx, dt, stage);
    return SWMHD_OK;
}

extern "C" int swmhd_substage(swmhd_ctx *ctx, double dt, int stage) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (stage < 1 || stage > 3) return fail(ctx, SWMHD_ERR_ARG, "stage must be 1, 2 or 3");
    if (ctx->in_substage) return fail(ctx, SWMHD_ERR_STATE, "a host-driven substage is in flight");
    int rc = need_comm(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->slabs[0].dev));
    rc = substage_async(ctx, dt, stage);
    if (rc) return rc;
    return sync_all(ctx);
}

// ---- diagnostics --------------------------------------------------------------------------------------
static void diag_fill(const swmhd_ctx *ctx, const double *r, swmhd_diag *o) {
    const swmhd_config &c = ctx->cfg;
    const double n = (double)c.Nx * (double)c.Ny, Lx = c.Nx * c.dx, Ly = c.Ny * c.dy;
    o->ke = r[0] / n * Lx * Ly; o->me = r[1] / n * Lx * Ly; o->pe = r[2] / n * Lx * Ly;
    o->total = o->ke + o->me + o->pe;
    o->sum_h = r[3]; o->max_abs_u = r[4]; o->max_abs_A = r[5]; o->min_h = -r[6]; o->max_abs_div_hB = r[7];
    o->all_finite = (r[8] == 0.0) ? 1 : 0; o->reserved = 0;
}

static inline bool diag_is_max(int q) { return q >= 4 && q <= 7; }

// Slots [first, first+count) of every local slab -> `count` combined raw records on the host.
// `reduce_ranks`: also combine across the processes of the ring (world > 1 with a communicator): two
// ncclAllReduce (sum, max) on copies of the slot block.  Sums add in slab order (in-process) / NCCL order.
static int fetch_diag(swmhd_ctx *ctx, int first, int count, bool reduce_ranks, std::vector<double> &out) {
    const size_t nd = (size_t)count * NDIAG;
    out.assign(nd, 0.0);
    const bool xproc = reduce_ranks && ctx->cfg.world > 1 && ctx->comm_ready;
    std::vector<double> tmp(2 * nd);
    bool first_slab = true;
    for (auto &s : ctx->slabs) {
        CK(cudaSetDevice(s.dev));
        const double *src = s.d_diag + (size_t)first * NDIAG;
        if (xproc) {
            const NcclApi *api = nccl_api();
            CK(cudaMemcpyAsync(s.d_red, src, nd * sizeof(double), cudaMemcpyDeviceToDevice, s.main));
            CK(cudaMemcpyAsync(s.d_red + nd, src, nd * sizeof(double), cudaMemcpyDeviceToDevice, s.main));
            NK(api->GroupStart());
            NK(api->AllReduce(s.d_red, s.d_red, nd, ncclFloat64, ncclSum, s.comm, s.main));
            NK(api->AllReduce(s.d_red + nd, s.d_red + nd, nd, ncclFloat64, ncclMax, s.comm, s.main));
            NK(api->GroupEnd());
            CK(cudaMemcpyAsync(tmp.data(), s.d_red, 2 * nd * sizeof(double), cudaMemcpyDeviceToHost, s.main));
            CK(cudaStreamSynchronize(s.main));
            for (size_t t = 0; t < nd; t++) out[t] = diag_is_max((int)(t % NDIAG)) ? tmp[nd + t] : tmp[t];
        } else {
            CK(cudaMemcpyAsync(tmp.data(), src, nd * sizeof(double), cudaMemcpyDeviceToHost, s.main));
            CK(cudaStreamSynchronize(s.main));
            for (size_t t = 0; t < nd; t++) {
                const bool mx = diag_is_max((int)(t % NDIAG));
                out[t] = first_slab ? tmp[t] : (mx ? fmax(out[t], tmp[t]) : out[t] + tmp[t]);
            }
        }
        first_slab = false;
    }
    return SWMHD_OK;
}

static int diag_async(swmhd_ctx *ctx, Slab &s, int slot) {
    const swmhd_config &c = ctx->cfg;
    CK(cudaSetDevice(s.dev));
    DiagParams d;
    d.Nx = ctx->Nx; d.Ny = s.Ny; d.P = ctx->P; d.form = c.formulation;
    d.dx = c.dx; d.dy = c.dy; d.g = c.g; d.h_ref = c.h_ref;
    for (int k = 0; k < 4; k++) d.U[k] = s.U[ctx->cur][k];
    d.partials = s.d_partials; d.nblocks = diag_blocks(ctx->Nx, s.Ny); d.stage = s.d_stage;
    ctx->launches += 2;
    CK(launch_diag(d, s.d_diag + (size_t)slot * NDIAG, s.main));
    return SWMHD_OK;
}

extern "C" int swmhd_diagnostics(swmhd_ctx *ctx, swmhd_diag *out) {
    if (!ctx || !out) return SWMHD_ERR_ARG;
    for (auto &s : ctx->slabs) { int rc = diag_async(ctx, s, 0); if (rc) return rc; }
    std::vector<double> r;
    int rc = fetch_diag(ctx, 0, 1, true, r);     // world > 1 without a communicator: this slab's partials
    if (rc) return rc;
    diag_fill(ctx, r.data(), out);
    return out->all_finite ? SWMHD_OK : fail(ctx, SWMHD_ERR_NONFINITE, "state contains NaN/Inf");
}

// ---- stepping ---------------------------------------------------------------------------------------------
// nsteps RK3 steps; step n uses dts[n] when dts is given, else dt
static int step_impl(swmhd_ctx *ctx, double dt, int nsteps, swmhd_diag *diags, const double *dts = nullptr) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (nsteps < 0) return fail(ctx, SWMHD_ERR_ARG, "nsteps < 0");
    if (ctx->in_substage) return fail(ctx, SWMHD_ERR_STATE, "a host-driven substage is in flight");
    if (ctx->last_stage != 0) return fail(ctx, SWMHD_ERR_STATE, "a step is in flight (finish its substages first)");
    int rc = need_comm(ctx);
    if (rc) return rc;
    for (auto &s : ctx->slabs) { CK(cudaSetDevice(s.dev)); CK(cudaEventRecord(s.ev0, s.main)); }
    std::vector<double> host;
    int done = 0;
    while (done < nsteps) {
        int chunk = nsteps - done;
        if (diags && chunk > ctx->diag_slots) chunk = ctx->diag_slots;
        for (int n = 0; n < chunk; n++) {
            // diagnostics of the state at the start of the step: fused into the stage-1 kernel
            const double dtn = dts ? dts[done + n] : dt;
            for (int st = 1; st <= 3; st++) { rc = substage_async(ctx, dtn, st, nullptr, nullptr, diags ? n : -1); if (rc) return rc; }
        }
        if (diags) {
            rc = fetch_diag(ctx, 0, chunk, true, host);
            if (rc) return rc;
            for (int n = 0; n < chunk; n++) diag_fill(ctx, &host[(size_t)n * NDIAG], &diags[done + n]);
        }
        done += chunk;
    }
    float ms_max = 0;
    for (auto &s : ctx->slabs) { CK(cudaSetDevice(s.dev)); CK(cudaEventRecord(s.ev1, s.main)); }
    rc = sync_all(ctx);
    if (rc) return rc;
    for (auto &s : ctx->slabs) {
        float ms = 0;
        CK(cudaSetDevice(s.dev));
        CK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
        if (ms > ms_max) ms_max = ms;
    }
    ctx->last_ms = ms_max;
    return SWMHD_OK;
}

extern "C" int swmhd_step(swmhd_ctx *ctx, double dt, int nsteps) { return step_impl(ctx, dt, nsteps, nullptr); }
extern "C" int swmhd_step_seq(swmhd_ctx *ctx, const double *dts, int nsteps, swmhd_diag *diags) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (!dts && nsteps > 0) return fail(ctx, SWMHD_ERR_ARG, "dts is null");
    for (int n = 0; n < nsteps; n++)
        if (!(dts[n] > 0.0) || !std::isfinite(dts[n])) return fail(ctx, SWMHD_ERR_ARG, "every dt of the sequence must be positive and finite");
    return step_impl(ctx, 0.0, nsteps, diags, dts);
}
extern "C" int swmhd_step_diag(swmhd_ctx *ctx, double dt, int nsteps, swmhd_diag *diags) {
    if (!diags) return ctx ? fail(ctx, SWMHD_ERR_ARG, "diags is null") : SWMHD_ERR_ARG;
    return step_impl(ctx, dt, nsteps, diags);
}

// set!(model, ...) + time_step!(model, dt) in one call, the upload pipelined with stage 1 by row bands.
// The four parent arrays travel band by band on the copy stream; as soon as band b and its neighbour band b+1 have
// arrived and their x halos are wrapped, stage 1 of band b runs on the main stream while later bands are still on the
// PCIe link.  The first and the last band also need the y halos (periodic images of the other end, or the wall BCs), so
// they run after the last band has landed.  Stages 2 and 3 follow as usual.  Single slab, between steps.
extern "C" int swmhd_upload_step(swmhd_ctx *ctx, const double *const host[4], size_t n_each, double dt, swmhd_diag *diag) {
    if (!ctx || !host) return SWMHD_ERR_ARG;
    if (multi(ctx)) return fail(ctx, SWMHD_ERR_STATE, "swmhd_upload_step is single-slab (upload with swmhd_set_field, then swmhd_step)");
    if (ctx->in_substage || ctx->last_stage != 0) return fail(ctx, SWMHD_ERR_STATE, "a step is in flight");
    if (!(dt > 0.0)) return fail(ctx, SWMHD_ERR_ARG, "dt must be positive");
    Slab &s = ctx->slabs[0];
    for (int k = 0; k < 4; k++) {
        if (!host[k]) return fail(ctx, SWMHD_ERR_ARG, "null host array");
        if (n_each < s.len[k]) return fail(ctx, SWMHD_ERR_ARG, "n_each is smaller than a field's parent array");
    }
    CK(cudaSetDevice(s.dev));
    const int ty = ctx->ty, ntr = s.ntr, P = ctx->P;
    // bands of whole tiles: 1/16 of the rows each on large grids (the last band to arrive holds back itself and its two
    // neighbours: 3/16 of stage 1), 1/8 on small ones, at least two row-blocked tiles
    const int want = (ntr >= 256) ? 16 : 8;
    int band_tr = ((ntr + want - 1) / want + ctx->edge_tr - 1) / ctx->edge_tr * ctx->edge_tr;
    if (band_tr < 2 * ctx->edge_tr) band_tr = 2 * ctx->edge_tr;
    const int nb = (ntr + band_tr - 1) / band_tr;
    if (nb < 3) {       // too small to pipeline: plain upload, then a step
        for (int k = 0; k < 4; k++) CK(cudaMemcpyAsync(s.U[ctx->cur][k], host[k], s.len[k] * sizeof(double), cudaMemcpyHostToDevice, s.main));
        HaloParams h = halo_params(ctx, s, s.U[ctx->cur], 3, s.Ny + 2, true);
        ctx->launches++;
        CK(launch_halo(h, s.main));
        return step_impl(ctx, dt, 1, diag);
    }
    EventBatch ev;
    CK(ev.create(nb, cudaEventDisableTiming));
    CK(cudaEventRecord(s.ev0, s.main));
    CK(cudaEventRecord(s.ev_main, s.main));                 // the copy stream may not overwrite what earlier work still reads
    CK(cudaStreamWaitEvent(s.copy, s.ev_main, 0));
    auto band_rows = [&](int b, int *r0, int *r1) { *r0 = b * band_tr * ty; *r1 = (b + 1) * band_tr * ty; if (*r1 > s.Ny || b == nb - 1) *r1 = s.Ny; };
    for (int b = 0; b < nb; b++) {                           // parent rows of band b; the end bands take the halo rows along
        int r0, r1;
        band_rows(b, &r0, &r1);
        for (int k = 0; k < 4; k++) {
            const size_t lo = (b == 0) ? 0 : (size_t)(3 + r0), hi = (b == nb - 1) ? (size_t)s.rows[k] : (size_t)(3 + r1);
            CK(cudaMemcpyAsync(s.U[ctx->cur][k] + lo * P, host[k] + lo * P, (hi - lo) * P * sizeof(double), cudaMemcpyHostToDevice, s.copy));
        }
        CK(cudaEventRecord(ev[b], s.copy));
    }
    KParams p = kparams(ctx, s, dt, 1);
    if (diag) p.diag = s.d_partials;
    auto stage1_band = [&](int b) -> int {
        p.tile_row0 = b * band_tr;
        p.tile_rows = (b == nb - 1) ? ntr - b * band_tr : band_tr;
        CK(launch_substage(ctx, p, 1, s.main));
        return SWMHD_OK;
    };
    for (int b = 0; b < nb; b++) {
        int r0, r1;
        band_rows(b, &r0, &r1);
        CK(cudaStreamWaitEvent(s.main, ev[b], 0));
        HaloParams h = halo_params(ctx, s, s.U[ctx->cur], 3 + r0, 3 + r1 - 1, false);        // x wrap of the band's rows
        ctx->launches++;
        CK(launch_halo(h, s.main));
        if (b >= 2) { int rc = stage1_band(b - 1); if (rc) return rc; }                          // bands b-2, b-1, b are complete
    }
    {   // y halos (periodic images / walls), then the two end bands
        HaloParams h = halo_params(ctx, s, s.U[ctx->cur], 3, 2, true);
        ctx->launches++;
        CK(launch_halo(h, s.main));
        int rc = stage1_band(0);
        if (rc) return rc;
        if ((rc = stage1_band(nb - 1))) return rc;
    }
    if (diag) {
        ctx->launches++;
        CK(launch_diag_final(s.d_partials, s.ntiles, s.d_stage, s.d_diag, s.main));
    }
    {
        HaloParams h = halo_params(ctx, s, s.U[1 - ctx->cur], 3, s.Ny + 2, true);
        ctx->launches++;
        CK(launch_halo(h, s.main));
    }
    ctx->cur = 1 - ctx->cur;
    tick(ctx, dt, 1);
    int rc = substage_async(ctx, dt, 2);
    if (rc == SWMHD_OK) rc = substage_async(ctx, dt, 3);
    if (rc == SWMHD_OK && diag) {
        std::vector<double> r;
        rc = fetch_diag(ctx, 0, 1, false, r);
        if (rc == SWMHD_OK) diag_fill(ctx, r.data(), diag);
    }
    if (rc == SWMHD_OK) {
        CK(cudaEventRecord(s.ev1, s.main));
        rc = sync_all(ctx);
        float ms = 0;
        if (rc == SWMHD_OK && cudaEventElapsedTime(&ms, s.ev0, s.ev1) == cudaSuccess) ctx->last_ms = ms;
    }
    return rc;
}

// nsteps RK3 steps with a CUDA-event pair around every substage-kernel launch (on the
// launching stream); out_ms[s] = mean device duration of the stage-(s+1) kernel.
static int step_profile_impl(swmhd_ctx *ctx, double dt, int nsteps, double out_ms[3], bool with_diag) {
    if (!ctx || !out_ms) return SWMHD_ERR_ARG;
    if (nsteps < 1 || nsteps > 512) return fail(ctx, SWMHD_ERR_ARG, "nsteps must be in 1..512");
    if (multi(ctx)) return fail(ctx, SWMHD_ERR_STATE, "single-slab only");
    if (ctx->last_stage != 0 || ctx->in_substage) return fail(ctx, SWMHD_ERR_STATE, "a step is in flight");
    Slab &s = ctx->slabs[0];
    CK(cudaSetDevice(s.dev));
    EventBatch ev;
    CK(ev.create((size_t)nsteps * 6, cudaEventDefault));
    int rc = SWMHD_OK;
    for (int n = 0; n < nsteps && rc == SWMHD_OK; n++)
        for (int st = 1; st <= 3 && rc == SWMHD_OK; st++)
            rc = substage_async(ctx, dt, st, ev[(size_t)n * 6 + 2 * (st - 1)], ev[(size_t)n * 6 + 2 * (st - 1) + 1], with_diag ? (n % ctx->diag_slots) : -1);
    if (rc == SWMHD_OK) {
        CK(cudaStreamSynchronize(s.main));
        for (int st = 0; st < 3; st++) {
            double acc = 0;
            for (int n = 0; n < nsteps; n++) {
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, ev[(size_t)n * 6 + 2 * st], ev[(size_t)n * 6 + 2 * st + 1]));
                acc += ms;
            }
            out_ms[st] = acc / nsteps;
        }
    }
    return rc;
}
extern "C" int swmhd_step_profile(swmhd_ctx *ctx, double dt, int nsteps, double out_ms[3]) { return step_profile_impl(ctx, dt, nsteps, out_ms, false); }
extern "C" int swmhd_step_profile_diag(swmhd_ctx *ctx, double dt, int nsteps, double out_ms[3]) { return step_profile_impl(ctx, dt, nsteps, out_ms, true); }

extern "C" int swmhd_tendencies(swmhd_ctx *ctx, double *const G_host[4], size_t n_each) {
    if (!ctx || !G_host) return SWMHD_ERR_ARG;
    if (multi(ctx)) return fail(ctx, SWMHD_ERR_STATE, "swmhd_tendencies is a single-slab test hook");
    Slab &s = ctx->slabs[0];
    for (int k = 0; k < 4; k++) {
        if (!G_host[k]) return fail(ctx, SWMHD_ERR_ARG, "null G_host entry");
        if (n_each < s.len[k]) return fail(ctx, SWMHD_ERR_ARG, "n_each is smaller than a field's parent array (v|vh has one row more in a Bounded-y grid)");
    }
    if (ctx->last_stage != 0 || ctx->in_substage)
        return fail(ctx, SWMHD_ERR_STATE, "tendencies overwrite G^-: call between steps, not after stage 1 or 2");
    CK(cudaSetDevice(s.dev));
    KParams p = kparams(ctx, s, 0.0, 0);
    CK(launch_substage(ctx, p, 0, s.main));
    for (int k = 0; k < 4; k++)
        CK(cudaMemcpyAsync(G_host[k], s.G[k], s.len[k] * sizeof(double), cudaMemcpyDeviceToHost, s.main));
    CK(cudaStreamSynchronize(s.main));
    return SWMHD_OK;
}

// ---- the field writer's outputs ---------------------------------------------------------------------------
// u, v, s computed on the device from the current state into dst[0..2] (arrays shaped like u, v, u), halos
// filled like u-, v- and u-located fields (x wrap + walls locally, y halos of a ring by exchange).
static int compute_outputs(swmhd_ctx *ctx, bool staging) {
    double *base[SWMHD_MAX_GPUS][4];
    for (size_t si = 0; si < ctx->slabs.size(); si++) {
        Slab &s = ctx->slabs[si];
        CK(cudaSetDevice(s.dev));
        double *const *dst = staging ? s.O : s.G;
        OutputParams o;
        o.Nx = ctx->Nx; o.Ny = s.Ny; o.P = ctx->P; o.form = ctx->cfg.formulation;
        for (int k = 0; k < 4; k++) o.U[k] = s.U[ctx->cur][k];
        o.out_u = dst[0]; o.out_v = dst[1]; o.out_s = dst[2];
        ctx->launches++;
        CK(launch_output(o, s.main));
        double *outs[4] = {dst[0], dst[1], dst[2], dst[2]};
        HaloParams h = halo_params(ctx, s, outs, 3, s.Ny + 2, true);
        h.grad = 0;
        h.rows[2] = s.rows[0]; h.rows[3] = s.rows[0];
        ctx->launches++;
        CK(launch_halo(h, s.main));
        for (int k = 0; k < 4; k++) base[si][k] = outs[k];
    }
    if (multi(ctx) && ctx->comm_ready) return exchange(ctx, base, 3, false);
    return SWMHD_OK;
}

static int copy_out(swmhd_ctx *ctx, Slab &s, double *host, const double *dev, int field, cudaStream_t st) {
    if (!in_process_multi(ctx)) {
        CK(cudaMemcpyAsync(host, dev, s.len[field] * sizeof(double), cudaMemcpyDeviceToHost, st));
    } else {
        int lo, hi;
        gather_rows(ctx, s, field, &lo, &hi);
        CK(cudaMemcpyAsync(host + (size_t)(s.j0 + lo) * ctx->P, dev + (size_t)lo * ctx->P,
                           (size_t)(hi - lo) * ctx->P * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    return SWMHD_OK;
}

extern "C" int swmhd_get_outputs(swmhd_ctx *ctx, double *u_host, double *v_host, double *s_host) {
    if (!ctx || !u_host || !v_host || !s_host) return SWMHD_ERR_ARG;
    if (ctx->in_substage || ctx->last_stage != 0) return fail(ctx, SWMHD_ERR_STATE, "outputs can only be taken between steps (they stage through G^-)");
    int rc = compute_outputs(ctx, false);
    if (rc) return rc;
    for (auto &s : ctx->slabs) {
        CK(cudaSetDevice(s.dev));
        if ((rc = copy_out(ctx, s, u_host, s.G[0], SWMHD_U, s.main))) return rc;
        if ((rc = copy_out(ctx, s, v_host, s.G[1], SWMHD_V, s.main))) return rc;
        if ((rc = copy_out(ctx, s, s_host, s.G[2], SWMHD_U, s.main))) return rc;
    }
    return sync_all(ctx);
}

extern "C" int swmhd_get_outputs_async(swmhd_ctx *ctx, double *u_host, double *v_host, double *s_host, double *A_host) {
    if (!ctx || !u_host || !v_host || !s_host || !A_host) return SWMHD_ERR_ARG;
    if (ctx->in_substage || ctx->last_stage != 0) return fail(ctx, SWMHD_ERR_STATE, "outputs can only be taken between steps");
    for (auto &s : ctx->slabs) {
        CK(cudaSetDevice(s.dev));
        const int shape[4] = {SWMHD_U, SWMHD_V, SWMHD_U, SWMHD_A};
        for (int k = 0; k < 4; k++)
            if (!s.O[k]) CK(dev_alloc(s, &s.O[k], s.len[shape[k]]));
        // the staging buffers may still be read by the previous asynchronous copy
        if (s.out_in_flight) CK(cudaStreamWaitEvent(s.main, s.ev_out_done, 0));
    }
    int rc = compute_outputs(ctx, true);
    if (rc) return rc;
    for (auto &s : ctx->slabs) {
        CK(cudaSetDevice(s.dev));
        CK(cudaMemcpyAsync(s.O[3], s.U[ctx->cur][SWMHD_A], s.len[SWMHD_A] * sizeof(double), cudaMemcpyDeviceToDevice, s.main));
        CK(cudaEventRecord(s.ev_out_ready, s.main));
        CK(cudaStreamWaitEvent(s.copy, s.ev_out_ready, 0));
        if ((rc = copy_out(ctx, s, u_host, s.O[0], SWMHD_U, s.copy))) return rc;
        if ((rc = copy_out(ctx, s, v_host, s.O[1], SWMHD_V, s.copy))) return rc;
        if ((rc = copy_out(ctx, s, s_host, s.O[2], SWMHD_U, s.copy))) return rc;
        if ((rc = copy_out(ctx, s, A_host, s.O[3], SWMHD_A, s.copy))) return rc;
        CK(cudaEventRecord(s.ev_out_done, s.copy));
        s.out_in_flight = true;
    }
    return SWMHD_OK;
}

extern "C" int swmhd_outputs_wait(swmhd_ctx *ctx) {
    if (!ctx) return SWMHD_ERR_ARG;
    for (auto &s : ctx->slabs) {
        if (!s.out_in_flight) continue;
        CK(cudaSetDevice(s.dev));
        CK(cudaEventSynchronize(s.ev_out_done));
        s.out_in_flight = false;
    }
    return SWMHD_OK;
}

extern "C" int swmhd_pin_host(void *ptr, size_t bytes) {
    if (!ptr || !bytes) return SWMHD_ERR_ARG;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { g_create_err = std::string("cudaHostRegister: ") + cudaGetErrorString(e); cudaGetLastError(); return SWMHD_ERR_CUDA; }
    return SWMHD_OK;
}
extern "C" int swmhd_unpin_host(void *ptr) {
    if (!ptr) return SWMHD_ERR_ARG;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { g_create_err = std::string("cudaHostUnregister: ") + cudaGetErrorString(e); cudaGetLastError(); return SWMHD_ERR_CUDA; }
    return SWMHD_OK;
}

extern "C" double swmhd_time(const swmhd_ctx *ctx) { return ctx ? ctx->time : NAN; }
extern "C" int64_t swmhd_iteration(const swmhd_ctx *ctx) { return ctx ? ctx->iter : -1; }
extern "C" int swmhd_set_clock(swmhd_ctx *ctx, double time, int64_t iteration) {
    if (!ctx) return SWMHD_ERR_ARG;
    ctx->time = time; ctx->iter = iteration;
    return SWMHD_OK;
}
extern "C" int64_t swmhd_launch_count(const swmhd_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" double swmhd_last_step_ms(const swmhd_ctx *ctx) { return ctx ? ctx->last_ms : NAN; }

// ---------------------------------------------------------------------------
// host-driven exchange (a host that owns its own transport): one slab per context
static int legacy_ok(swmhd_ctx *ctx) {
    if (in_process_multi(ctx)) return fail(ctx, SWMHD_ERR_STATE, "n_gpus > 1 contexts exchange internally: use swmhd_step / swmhd_substage");
    return SWMHD_OK;
}

extern "C" int swmhd_set_streams(swmhd_ctx *ctx, void *main_stream, void *edge_stream) {
    if (!ctx) return SWMHD_ERR_ARG;
    int rc = legacy_ok(ctx);
    if (rc) return rc;
    Slab &s = ctx->slabs[0];
    CK(cudaSetDevice(s.dev));
    CK(cudaDeviceSynchronize());
    if (s.own_streams) {
        cudaStreamDestroy(s.main); cudaStreamDestroy(s.edge);
        s.own_streams = false;
    }
    s.main = (cudaStream_t)main_stream;
    s.edge = (cudaStream_t)edge_stream;
    return SWMHD_OK;
}

extern "C" int swmhd_substage_edges(swmhd_ctx *ctx, double dt, int stage) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (stage < 1 || stage > 3) return fail(ctx, SWMHD_ERR_ARG, "stage must be 1, 2 or 3");
    if (ctx->in_substage) return fail(ctx, SWMHD_ERR_STATE, "previous substage not finished");
    int rc = legacy_ok(ctx);
    if (rc) return rc;
    rc = slab_edges(ctx, ctx->slabs[0], dt, stage, nullptr);
    if (rc) return rc;
    if (stage == 1) ctx->armed_slot = -1;
    CK(cudaEventRecord(ctx->slabs[0].ev_edge, ctx->slabs[0].edge));
    ctx->in_substage = true;
    ctx->pending_dt = dt;
    return SWMHD_OK;
}

extern "C" int swmhd_substage_interior(swmhd_ctx *ctx, double dt, int stage) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (!ctx->in_substage) return fail(ctx, SWMHD_ERR_STATE, "call swmhd_substage_edges first");
    return slab_interior(ctx, ctx->slabs[0], dt, stage);
}

extern "C" int swmhd_substage_finish(swmhd_ctx *ctx, int stage) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (!ctx->in_substage) return fail(ctx, SWMHD_ERR_STATE, "no substage in flight");
    if (stage < 1 || stage > 3) return fail(ctx, SWMHD_ERR_ARG, "stage must be 1, 2 or 3");
    int rc = slab_join(ctx, ctx->slabs[0], nullptr);
    if (rc) return rc;
    ctx->cur = 1 - ctx->cur;
    ctx->in_substage = false;
    tick(ctx, ctx->pending_dt, stage);
    return SWMHD_OK;
}

extern "C" int swmhd_exchange_rows(swmhd_ctx *ctx, int field, int which, void **dev_ptr, int *nrows, size_t *row_doubles) {
    if (!ctx || !dev_ptr || !nrows || !row_doubles) return SWMHD_ERR_ARG;
    if (field < 0 || field > 3 || which < 0 || which > 7) return fail(ctx, SWMHD_ERR_ARG, "bad field/which");
    int rc = legacy_ok(ctx);
    if (rc) return rc;
    Slab &s = ctx->slabs[0];
    int buf = (which >= 4) ? ctx->cur : 1 - ctx->cur;   // 0-3: state being written, 4-7: current state
    int w = which & 3;
    int row = (w == 0) ? 3 : (w == 1) ? s.Ny : (w == 2) ? 0 : s.Ny + 3;
    *dev_ptr = (void *)(s.U[buf][field] + (size_t)row * ctx->P);
    *nrows = 3;
    *row_doubles = (size_t)ctx->P;
    return SWMHD_OK;
}

extern "C" int swmhd_arm_diag(swmhd_ctx *ctx, int slot) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (slot < 0 || slot >= ctx->diag_slots) return fail(ctx, SWMHD_ERR_ARG, "diag slot out of range");
    ctx->armed_slot = slot;
    return SWMHD_OK;
}

extern "C" int swmhd_get_diag_slots(swmhd_ctx *ctx, int first, int count, swmhd_diag *out) {
    if (!ctx || !out) return SWMHD_ERR_ARG;
    if (first < 0 || count < 0 || first + count > ctx->diag_slots) return fail(ctx, SWMHD_ERR_ARG, "diag slots out of range");
    int rc = sync_all(ctx);
    if (rc) return rc;
    std::vector<double> host;
    rc = fetch_diag(ctx, first, count, false, host);     // this process's slabs only: the host combines ranks
    if (rc) return rc;
    for (int n = 0; n < count; n++) diag_fill(ctx, &host[(size_t)n * NDIAG], &out[n]);
    return SWMHD_OK;
}

// Test hook (SWMHD_GUARD=1 at create time): SWMHD_OK when every guard zone still holds its sentinel, SWMHD_ERR_STATE and a
// message naming the first damaged allocation otherwise; SWMHD_ERR_ARG when the context was created without guards.
extern "C" int swmhd_check_guards(swmhd_ctx *ctx) {
    if (!ctx) return SWMHD_ERR_ARG;
    int rc = sync_all(ctx);
    if (rc) return rc;
    std::vector<unsigned long long> g(GUARD_DOUBLES);
    size_t checked = 0;
    for (auto &s : ctx->slabs) {
        CK(cudaSetDevice(s.dev));
        for (size_t a = 0; a < s.guarded.size(); a++) {
            for (int side = 0; side < 2; side++) {
                const double *z = side ? s.guarded[a].first + s.guarded[a].second : s.guarded[a].first - GUARD_DOUBLES;
                CK(cudaMemcpy(g.data(), z, GUARD_DOUBLES * sizeof(double), cudaMemcpyDeviceToHost));
                for (size_t t = 0; t < GUARD_DOUBLES; t++)
                    if (g[t] != GUARD_WORD) {
                        char buf[200];
                        snprintf(buf, sizeof buf, "guard zone %s allocation %zu (slab %d, %zu doubles) overwritten at offset %zu",
                                 side ? "after" : "before", a, s.index, s.guarded[a].second, t);
                        return fail(ctx, SWMHD_ERR_STATE, buf);
                    }
                checked++;
            }
        }
    }
    if (!checked) return fail(ctx, SWMHD_ERR_ARG, "context was created without SWMHD_GUARD=1");
    return SWMHD_OK;
}

extern "C" int swmhd_sync(swmhd_ctx *ctx) {
    if (!ctx) return SWMHD_ERR_ARG;
    return sync_all(ctx);
}
