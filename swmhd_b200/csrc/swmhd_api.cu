// swmhd_api.cu — the C ABI of include/swmhd.h over the CUDA kernels.
// Host-side orchestration only: buffers, streams, launch sequencing, clock.
#include "../../include/swmhd.h"
#include "kparams.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

using namespace swmhd;

struct swmhd_ctx {
    swmhd_config cfg;
    int Nx, Ny, P;              // local slab
    int rows[4];
    size_t len[4];
    double *U[2][4];
    CUtensorMap tmap[2][4];     // TMA descriptors of U[b][k]: 2-D (P x rows) FP64, box (TX+6) x (TY+6)
    CUtensorMap tmap_rb[2][4];  // same arrays, box of the row-blocked kernel
    int use_tma, use_rb;
    double *G[4];
    int cur;                    // U[cur] = current state
    double *d_partials, *d_diag, *d_stage; // diag partials; d_diag holds NDIAG doubles per slot
    int diag_slots;
    int nblocks_diag, ntiles;
    cudaStream_t main, edge;
    bool own_streams;
    cudaEvent_t ev0, ev1, ev_edge, ev_main;
    double time;
    int64_t iter, launches;
    double last_ms;
    int tx, ty, ntr, n_last;    // tile geometry: tile rows, tile rows in the north edge group
    bool in_substage;
    double pending_dt;
    int armed_slot;             // >= 0: the next stage-1 slab substage also produces diagnostics into this slot
    int pending_slot;
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

static int fail(swmhd_ctx *ctx, int code, const char *msg) {
    if (ctx) ctx->err = msg; else g_create_err = msg;
    return code;
}

extern "C" int swmhd_abi_version(void) { return SWMHD_ABI_VERSION; }

extern "C" const char *swmhd_last_error(const swmhd_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
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

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, SWMHD_ERR_NODEVICE, "no CUDA device: libswmhd_cuda has no CPU fallback");
    }
    if (c->device < 0 || c->device >= ndev) return fail(nullptr, SWMHD_ERR_ARG, "bad device ordinal");

    swmhd_ctx *ctx = new (std::nothrow) swmhd_ctx();
    if (!ctx) return fail(nullptr, SWMHD_ERR_ARG, "out of host memory");
    ctx->cfg = *c;
    ctx->Nx = c->Nx; ctx->Ny = c->slab_ny; ctx->P = c->Nx + 6;
    for (int k = 0; k < 4; k++) {
        ctx->rows[k] = ctx->Ny + 6 + ((k == SWMHD_V && c->topo_y == SWMHD_BOUNDED) ? 1 : 0);
        ctx->len[k] = (size_t)ctx->P * ctx->rows[k];
        ctx->U[0][k] = ctx->U[1][k] = ctx->G[k] = nullptr;
    }
    ctx->cur = 0; ctx->time = 0; ctx->iter = 0; ctx->launches = 0; ctx->last_ms = 0; ctx->in_substage = false; ctx->armed_slot = -1; ctx->pending_slot = -1;
    ctx->d_partials = ctx->d_diag = ctx->d_stage = nullptr; ctx->main = ctx->edge = nullptr; ctx->own_streams = false;
    ctx->ev0 = ctx->ev1 = ctx->ev_edge = ctx->ev_main = nullptr;
    substage_tile(&ctx->tx, &ctx->ty);
    ctx->ntr = (ctx->Ny + ctx->ty - 1) / ctx->ty;
    ctx->n_last = (ctx->Ny - (ctx->ntr - 1) * ctx->ty >= 3) ? 1 : 2;

    auto bail = [&](const char *what, cudaError_t er) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(er));
        g_create_err = buf;
        swmhd_destroy(ctx);
        return SWMHD_ERR_CUDA;
    };
    if ((e = cudaSetDevice(c->device)) != cudaSuccess) return bail("cudaSetDevice", e);
    for (int k = 0; k < 4; k++) {
        size_t bytes = ctx->len[k] * sizeof(double);
        for (int b = 0; b < 2; b++) {
            if ((e = cudaMalloc(&ctx->U[b][k], bytes)) != cudaSuccess) return bail("cudaMalloc state", e);
            if ((e = cudaMemset(ctx->U[b][k], 0, bytes)) != cudaSuccess) return bail("cudaMemset", e);
        }
        if ((e = cudaMalloc(&ctx->G[k], bytes)) != cudaSuccess) return bail("cudaMalloc tendency", e);
        if ((e = cudaMemset(ctx->G[k], 0, bytes)) != cudaSuccess) return bail("cudaMemset", e);
    }
    // TMA tile loads need a 16-byte row pitch (even Nx); otherwise the kernels use plain loads
    ctx->use_tma = (ctx->P % 2 == 0) ? 1 : 0;
    for (int b = 0; b < 2 && ctx->use_tma; b++)
        for (int k = 0; k < 4 && ctx->use_tma; k++)
            if (!encode_field_map(&ctx->tmap[b][k], ctx->U[b][k], ctx->P, ctx->rows[k], ctx->tx + 6, ctx->ty + 6)) ctx->use_tma = 0;
    if (getenv("SWMHD_NO_TMA")) ctx->use_tma = 0;
    // row-blocked kernels (FAST arithmetic): taller box; their 8-row launch granularity is ctx->ty
    ctx->use_rb = (ctx->use_tma && ctx->ty == 8 && c->arith == SWMHD_ARITH_FAST) ? 1 : 0;
    if (ctx->use_rb) {
        int rtx, rty;
        substage_rb_tile(&rtx, &rty);
        for (int b = 0; b < 2 && ctx->use_rb; b++)
            for (int k = 0; k < 4 && ctx->use_rb; k++)
                if (!encode_field_map(&ctx->tmap_rb[b][k], ctx->U[b][k], ctx->P, ctx->rows[k], rtx + 6, rty + 6)) ctx->use_rb = 0;
    }
    if (getenv("SWMHD_NO_RB")) ctx->use_rb = 0;
    // per-tile diagnostic partials: one slot per (8-row tile row, tile column) of the kernel that runs stage 1
    const int diag_tiles_x = (ctx->use_rb && (substage_rb_stage_mask() & 1)) ? substage_rb_tiles_x(c->formulation, ctx->Nx)
                                                                             : (ctx->Nx + ctx->tx - 1) / ctx->tx;
    ctx->nblocks_diag = diag_blocks(ctx->Nx, ctx->Ny);
    ctx->ntiles = diag_tiles_x * ctx->ntr;
    if (ctx->ntiles > ctx->nblocks_diag) ctx->nblocks_diag = ctx->ntiles;   // d_partials serves both diag paths
    ctx->diag_slots = 1024;
    if ((e = cudaMalloc(&ctx->d_partials, (size_t)ctx->nblocks_diag * NDIAG * sizeof(double))) != cudaSuccess) return bail("cudaMalloc diag", e);
    if ((e = cudaMalloc(&ctx->d_diag, (size_t)ctx->diag_slots * NDIAG * sizeof(double))) != cudaSuccess) return bail("cudaMalloc diag", e);
    if ((e = cudaMalloc(&ctx->d_stage, (size_t)diag_stage_doubles() * sizeof(double))) != cudaSuccess) return bail("cudaMalloc diag", e);
    int lo, hi;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if ((e = cudaStreamCreateWithPriority(&ctx->main, cudaStreamNonBlocking, lo)) != cudaSuccess) return bail("stream", e);
    if ((e = cudaStreamCreateWithPriority(&ctx->edge, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail("stream", e);
    ctx->own_streams = true;
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail("event", e);
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail("event", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_edge, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    *out = ctx;
    return SWMHD_OK;
}

extern "C" void swmhd_destroy(swmhd_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    cudaDeviceSynchronize();
    for (int k = 0; k < 4; k++) {
        cudaFree(ctx->U[0][k]); cudaFree(ctx->U[1][k]); cudaFree(ctx->G[k]);
    }
    cudaFree(ctx->d_partials); cudaFree(ctx->d_diag); cudaFree(ctx->d_stage);
    if (ctx->own_streams) {
        if (ctx->main) cudaStreamDestroy(ctx->main);
        if (ctx->edge) cudaStreamDestroy(ctx->edge);
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_edge) cudaEventDestroy(ctx->ev_edge);
    if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
    cudaGetLastError();
    delete ctx;
}

extern "C" size_t swmhd_field_len(const swmhd_ctx *ctx, int field) {
    if (!ctx || field < 0 || field > 3) return 0;
    return ctx->len[field];
}

extern "C" int swmhd_set_field(swmhd_ctx *ctx, int field, const double *host, size_t n) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (field < 0 || field > 3 || !host) return fail(ctx, SWMHD_ERR_ARG, "bad field/host");
    if (n != ctx->len[field]) return fail(ctx, SWMHD_ERR_ARG, "host buffer is not the parent array of this field (length mismatch)");
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaMemcpyAsync(ctx->U[ctx->cur][field], host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->main));
    CK(cudaStreamSynchronize(ctx->main));
    return SWMHD_OK;
}

extern "C" int swmhd_get_field(swmhd_ctx *ctx, int field, double *host, size_t n) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (field < 0 || field > 3 || !host) return fail(ctx, SWMHD_ERR_ARG, "bad field/host");
    if (n != ctx->len[field]) return fail(ctx, SWMHD_ERR_ARG, "host buffer is not the parent array of this field (length mismatch)");
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaMemcpyAsync(host, ctx->U[ctx->cur][field], n * sizeof(double), cudaMemcpyDeviceToHost, ctx->main));
    CK(cudaStreamSynchronize(ctx->main));
    return SWMHD_OK;
}

// ---------------------------------------------------------------------------
static HaloParams halo_params(swmhd_ctx *ctx, double *const U[4], int j_lo, int j_hi, bool y_part) {
    const swmhd_config &c = ctx->cfg;
    HaloParams h;
    h.Nx = ctx->Nx; h.Ny = ctx->Ny; h.P = ctx->P;
    h.by = (c.topo_y == SWMHD_BOUNDED);
    h.first = (c.rank == 0); h.last = (c.rank == c.world - 1);
    h.y_mode = 0;
    if (y_part) {
        if (h.by) h.y_mode = 2;
        else if (c.world == 1) h.y_mode = 1;
    }
    h.grad = c.A_gradient_bc; h.gs = c.A_grad_south; h.gn = c.A_grad_north; h.dy = c.dy;
    h.j_lo = j_lo; h.j_hi = j_hi;
    for (int k = 0; k < 4; k++) { h.U[k] = U[k]; h.rows[k] = ctx->rows[k]; }
    return h;
}

static KParams kparams(swmhd_ctx *ctx, double dt, int stage) {
    const swmhd_config &c = ctx->cfg;
    KParams p;
    p.Nx = ctx->Nx; p.Ny = ctx->Ny; p.P = ctx->P;
    p.gj0 = c.slab_j0; p.NyG = c.Ny; p.by = (c.topo_y == SWMHD_BOUNDED);
    p.tile_row0 = 0; p.tile_rows = ctx->ntr;
    for (int k = 0; k < 4; k++) p.rows[k] = ctx->rows[k];
    p.dx = c.dx; p.dy = c.dy; p.rdx = 1.0 / c.dx; p.rdy = 1.0 / c.dy; p.inv_az = 1.0 / (c.dx * c.dy);
    p.g = c.g; p.f = c.f; p.eps = c.weno_eps; p.h_ref = c.h_ref;
    p.dt = dt;
    p.gam = stage >= 1 ? RK_GAMMA[stage - 1] : 0.0;
    p.zet = stage >= 1 ? RK_ZETA[stage - 1] : 0.0;
    p.dtgam = dt * p.gam;
    for (int k = 0; k < 4; k++) {
        p.Uo[k] = ctx->U[ctx->cur][k];
        p.Un[k] = ctx->U[1 - ctx->cur][k];
        p.G[k] = ctx->G[k];
    }
    p.diag = nullptr;
    p.use_tma = ctx->use_tma;
    if (ctx->use_tma)
        for (int k = 0; k < 4; k++) p.tm[k] = ctx->tmap[ctx->cur][k];
    p.use_rb = ctx->use_rb;
    p.row_begin = p.row_end = 0;
    if (ctx->use_rb)
        for (int k = 0; k < 4; k++) p.tm_rb[k] = ctx->tmap_rb[ctx->cur][k];
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
}

extern "C" int swmhd_fill_halos(swmhd_ctx *ctx) {
    if (!ctx) return SWMHD_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    HaloParams h = halo_params(ctx, ctx->U[ctx->cur], 3, ctx->Ny + 2, true);
    ctx->launches++;
    CK(launch_halo(h, ctx->main));
    CK(cudaStreamSynchronize(ctx->main));
    return SWMHD_OK;
}

// one substage on the main stream, no host synchronisation
static int substage_async(swmhd_ctx *ctx, double dt, int stage, cudaEvent_t e0 = nullptr, cudaEvent_t e1 = nullptr, int diag_slot = -1) {
    KParams p = kparams(ctx, dt, stage);
    if (diag_slot >= 0 && stage == 1) p.diag = ctx->d_partials;
    if (e0) CK(cudaEventRecord(e0, ctx->main));
    CK(launch_substage(ctx, p, stage, ctx->main));
    if (e1) CK(cudaEventRecord(e1, ctx->main));
    if (diag_slot >= 0 && stage == 1) {
        ctx->launches++;
        CK(launch_diag_final(ctx->d_partials, ctx->ntiles, ctx->d_stage, ctx->d_diag + (size_t)diag_slot * NDIAG, ctx->main));
    }
    HaloParams h = halo_params(ctx, ctx->U[1 - ctx->cur], 3, ctx->Ny + 2, true);
    ctx->launches++;
    CK(launch_halo(h, ctx->main));
    ctx->cur = 1 - ctx->cur;
    tick(ctx, dt, stage);
    return SWMHD_OK;
}

extern "C" int swmhd_substage(swmhd_ctx *ctx, double dt, int stage) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (stage < 1 || stage > 3) return fail(ctx, SWMHD_ERR_ARG, "stage must be 1, 2 or 3");
    if (ctx->cfg.world != 1) return fail(ctx, SWMHD_ERR_STATE, "swmhd_substage is single-slab; use substage_edges/interior/finish");
    CK(cudaSetDevice(ctx->cfg.device));
    int rc = substage_async(ctx, dt, stage);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->main));
    return SWMHD_OK;
}

static int diag_async(swmhd_ctx *ctx, int slot) {
    const swmhd_config &c = ctx->cfg;
    DiagParams d;
    d.Nx = ctx->Nx; d.Ny = ctx->Ny; d.P = ctx->P; d.form = c.formulation;
    d.dx = c.dx; d.dy = c.dy; d.g = c.g; d.h_ref = c.h_ref;
    for (int k = 0; k < 4; k++) d.U[k] = ctx->U[ctx->cur][k];
    d.partials = ctx->d_partials; d.nblocks = diag_blocks(ctx->Nx, ctx->Ny); d.stage = ctx->d_stage;
    ctx->launches += 2;
    CK(launch_diag(d, ctx->d_diag + (size_t)slot * NDIAG, ctx->main));
    return SWMHD_OK;
}

static void diag_fill(const swmhd_ctx *ctx, const double *r, swmhd_diag *o) {
    const swmhd_config &c = ctx->cfg;
    const double n = (double)c.Nx * (double)c.Ny, Lx = c.Nx * c.dx, Ly = c.Ny * c.dy;
    o->ke = r[0] / n * Lx * Ly; o->me = r[1] / n * Lx * Ly; o->pe = r[2] / n * Lx * Ly;
    o->total = o->ke + o->me + o->pe;
    o->sum_h = r[3]; o->max_abs_u = r[4]; o->max_abs_A = r[5]; o->min_h = -r[6]; o->max_abs_div_hB = r[7];
    o->all_finite = (r[8] == 0.0) ? 1 : 0; o->reserved = 0;
}

extern "C" int swmhd_diagnostics(swmhd_ctx *ctx, swmhd_diag *out) {
    if (!ctx || !out) return SWMHD_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    int rc = diag_async(ctx, 0);
    if (rc) return rc;
    double r[NDIAG];
    CK(cudaMemcpyAsync(r, ctx->d_diag, sizeof r, cudaMemcpyDeviceToHost, ctx->main));
    CK(cudaStreamSynchronize(ctx->main));
    diag_fill(ctx, r, out);
    return out->all_finite ? SWMHD_OK : fail(ctx, SWMHD_ERR_NONFINITE, "state contains NaN/Inf");
}

static int step_impl(swmhd_ctx *ctx, double dt, int nsteps, swmhd_diag *diags) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (nsteps < 0) return fail(ctx, SWMHD_ERR_ARG, "nsteps < 0");
    if (ctx->cfg.world != 1) return fail(ctx, SWMHD_ERR_STATE, "swmhd_step is single-slab; drive slabs with substage_edges/interior/finish");
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaEventRecord(ctx->ev0, ctx->main));
    std::vector<double> host;
    int done = 0;
    while (done < nsteps) {
        int chunk = nsteps - done;
        if (diags && chunk > ctx->diag_slots) chunk = ctx->diag_slots;
        for (int n = 0; n < chunk; n++) {
            // diagnostics of the state at the start of the step: fused into the stage-1 kernel
            for (int s = 1; s <= 3; s++) { int rc = substage_async(ctx, dt, s, nullptr, nullptr, diags ? n : -1); if (rc) return rc; }
        }
        if (diags) {
            host.resize((size_t)chunk * NDIAG);
            CK(cudaMemcpyAsync(host.data(), ctx->d_diag, host.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->main));
            CK(cudaStreamSynchronize(ctx->main));
            for (int n = 0; n < chunk; n++) diag_fill(ctx, &host[(size_t)n * NDIAG], &diags[done + n]);
        }
        done += chunk;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->main));
    CK(cudaStreamSynchronize(ctx->main));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    return SWMHD_OK;
}

extern "C" int swmhd_step(swmhd_ctx *ctx, double dt, int nsteps) { return step_impl(ctx, dt, nsteps, nullptr); }
extern "C" int swmhd_step_diag(swmhd_ctx *ctx, double dt, int nsteps, swmhd_diag *diags) {
    if (!diags) return ctx ? fail(ctx, SWMHD_ERR_ARG, "diags is null") : SWMHD_ERR_ARG;
    return step_impl(ctx, dt, nsteps, diags);
}

// nsteps RK3 steps with a CUDA-event pair around every substage-kernel launch (on the
// launching stream); out_ms[s] = mean device duration of the stage-(s+1) kernel.
static int step_profile_impl(swmhd_ctx *ctx, double dt, int nsteps, double out_ms[3], bool with_diag);
extern "C" int swmhd_step_profile(swmhd_ctx *ctx, double dt, int nsteps, double out_ms[3]) { return step_profile_impl(ctx, dt, nsteps, out_ms, false); }
extern "C" int swmhd_step_profile_diag(swmhd_ctx *ctx, double dt, int nsteps, double out_ms[3]) { return step_profile_impl(ctx, dt, nsteps, out_ms, true); }
static int step_profile_impl(swmhd_ctx *ctx, double dt, int nsteps, double out_ms[3], bool with_diag) {
    if (!ctx || !out_ms) return SWMHD_ERR_ARG;
    if (nsteps < 1 || nsteps > 512) return fail(ctx, SWMHD_ERR_ARG, "nsteps must be in 1..512");
    if (ctx->cfg.world != 1) return fail(ctx, SWMHD_ERR_STATE, "single-slab only");
    CK(cudaSetDevice(ctx->cfg.device));
    std::vector<cudaEvent_t> ev((size_t)nsteps * 6);
    for (auto &e : ev) CK(cudaEventCreate(&e));
    int rc = SWMHD_OK;
    for (int n = 0; n < nsteps && rc == SWMHD_OK; n++)
        for (int s = 1; s <= 3 && rc == SWMHD_OK; s++)
            rc = substage_async(ctx, dt, s, ev[(size_t)n * 6 + 2 * (s - 1)], ev[(size_t)n * 6 + 2 * (s - 1) + 1], with_diag ? (n % ctx->diag_slots) : -1);
    if (rc == SWMHD_OK) {
        CK(cudaStreamSynchronize(ctx->main));
        for (int s = 0; s < 3; s++) {
            double acc = 0;
            for (int n = 0; n < nsteps; n++) {
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, ev[(size_t)n * 6 + 2 * s], ev[(size_t)n * 6 + 2 * s + 1]));
                acc += ms;
            }
            out_ms[s] = acc / nsteps;
        }
    }
    for (auto &e : ev) cudaEventDestroy(e);
    return rc;
}

extern "C" int swmhd_tendencies(swmhd_ctx *ctx, double *const G_host[4], size_t n_each) {
    if (!ctx || !G_host) return SWMHD_ERR_ARG;
    (void)n_each;
    CK(cudaSetDevice(ctx->cfg.device));
    KParams p = kparams(ctx, 0.0, 0);
    CK(launch_substage(ctx, p, 0, ctx->main));
    for (int k = 0; k < 4; k++) {
        if (!G_host[k]) return fail(ctx, SWMHD_ERR_ARG, "null G_host entry");
        CK(cudaMemcpyAsync(G_host[k], ctx->G[k], ctx->len[k] * sizeof(double), cudaMemcpyDeviceToHost, ctx->main));
    }
    CK(cudaStreamSynchronize(ctx->main));
    return SWMHD_OK;
}

// u, v, s of the field writer, computed on the device from the current state.  The three results are
// staged in the tendency buffers G[0..2] (free between steps: stage 1 never reads G^-), halos filled
// like u-, v- and u-located fields, then copied to the host parent arrays.
extern "C" int swmhd_get_outputs(swmhd_ctx *ctx, double *u_host, double *v_host, double *s_host) {
    if (!ctx || !u_host || !v_host || !s_host) return SWMHD_ERR_ARG;
    if (ctx->in_substage) return fail(ctx, SWMHD_ERR_STATE, "outputs can only be taken between steps");
    CK(cudaSetDevice(ctx->cfg.device));
    OutputParams o;
    o.Nx = ctx->Nx; o.Ny = ctx->Ny; o.P = ctx->P; o.form = ctx->cfg.formulation;
    for (int k = 0; k < 4; k++) o.U[k] = ctx->U[ctx->cur][k];
    o.out_u = ctx->G[0]; o.out_v = ctx->G[1]; o.out_s = ctx->G[2];
    ctx->launches++;
    CK(launch_output(o, ctx->main));
    if (ctx->cfg.world == 1) {          // periodic / wall halos of the outputs: (u, v, s) behave like (u, v, u)
        double *outs[4] = {ctx->G[0], ctx->G[1], ctx->G[2], ctx->G[2]};
        HaloParams h = halo_params(ctx, outs, 3, ctx->Ny + 2, true);
        h.grad = 0;
        h.rows[2] = ctx->rows[0]; h.rows[3] = ctx->rows[0];
        ctx->launches++;
        CK(launch_halo(h, ctx->main));
    }
    CK(cudaMemcpyAsync(u_host, ctx->G[0], ctx->len[0] * sizeof(double), cudaMemcpyDeviceToHost, ctx->main));
    CK(cudaMemcpyAsync(v_host, ctx->G[1], ctx->len[1] * sizeof(double), cudaMemcpyDeviceToHost, ctx->main));
    CK(cudaMemcpyAsync(s_host, ctx->G[2], ctx->len[0] * sizeof(double), cudaMemcpyDeviceToHost, ctx->main));
    CK(cudaStreamSynchronize(ctx->main));
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
// y-slab plumbing
extern "C" int swmhd_set_streams(swmhd_ctx *ctx, void *main_stream, void *edge_stream) {
    if (!ctx) return SWMHD_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaDeviceSynchronize());
    if (ctx->own_streams) {
        cudaStreamDestroy(ctx->main); cudaStreamDestroy(ctx->edge);
        ctx->own_streams = false;
    }
    ctx->main = (cudaStream_t)main_stream;
    ctx->edge = (cudaStream_t)edge_stream;
    return SWMHD_OK;
}

extern "C" int swmhd_substage_edges(swmhd_ctx *ctx, double dt, int stage) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (stage < 1 || stage > 3) return fail(ctx, SWMHD_ERR_ARG, "stage must be 1, 2 or 3");
    if (ctx->in_substage) return fail(ctx, SWMHD_ERR_STATE, "previous substage not finished");
    CK(cudaSetDevice(ctx->cfg.device));
    // the edge stream may not overwrite rows the previous substage's kernels still read
    CK(cudaEventRecord(ctx->ev_main, ctx->main));
    CK(cudaStreamWaitEvent(ctx->edge, ctx->ev_main, 0));
    KParams p = kparams(ctx, dt, stage);
    ctx->pending_slot = -1;
    if (stage == 1 && ctx->armed_slot >= 0) {
        p.diag = ctx->d_partials;
        ctx->pending_slot = ctx->armed_slot;
        ctx->armed_slot = -1;
    }
    const int ntr = ctx->ntr, nl = ctx->n_last;
    if (ntr <= 1 + nl) {            // slab too thin to split: everything is "edge"
        CK(launch_substage(ctx, p, stage, ctx->edge));
    } else {
        p.tile_row0 = 0; p.tile_rows = 1;
        CK(launch_substage(ctx, p, stage, ctx->edge));
        p.tile_row0 = ntr - nl; p.tile_rows = nl;
        CK(launch_substage(ctx, p, stage, ctx->edge));
    }
    // x wrap of the rows that are about to be sent, and wall BCs on end ranks
    double *const *Un = ctx->U[1 - ctx->cur];
    int ty = ctx->ty;
    if (ntr <= 1 + nl) {
        HaloParams h = halo_params(ctx, Un, 3, ctx->Ny + 2, true);
        ctx->launches++;
        CK(launch_halo(h, ctx->edge));
    } else {
        HaloParams h = halo_params(ctx, Un, 3, 3 + ty - 1, true);
        ctx->launches++;
        CK(launch_halo(h, ctx->edge));
        HaloParams h2 = halo_params(ctx, Un, 3 + (ntr - nl) * ty, ctx->Ny + 2, false);
        ctx->launches++;
        CK(launch_halo(h2, ctx->edge));
    }
    CK(cudaEventRecord(ctx->ev_edge, ctx->edge));
    ctx->in_substage = true;
    ctx->pending_dt = dt;
    return SWMHD_OK;
}

extern "C" int swmhd_substage_interior(swmhd_ctx *ctx, double dt, int stage) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (!ctx->in_substage) return fail(ctx, SWMHD_ERR_STATE, "call swmhd_substage_edges first");
    CK(cudaSetDevice(ctx->cfg.device));
    const int ntr = ctx->ntr, nl = ctx->n_last;
    if (ntr > 1 + nl) {
        KParams p = kparams(ctx, dt, stage);
        if (stage == 1 && ctx->pending_slot >= 0) p.diag = ctx->d_partials;
        p.tile_row0 = 1; p.tile_rows = ntr - nl - 1;
        CK(launch_substage(ctx, p, stage, ctx->main));
        HaloParams h = halo_params(ctx, ctx->U[1 - ctx->cur], 3 + ctx->ty, 3 + (ntr - nl) * ctx->ty - 1, false);
        ctx->launches++;
        CK(launch_halo(h, ctx->main));
    }
    return SWMHD_OK;
}

extern "C" int swmhd_substage_finish(swmhd_ctx *ctx, int stage) {
    if (!ctx) return SWMHD_ERR_ARG;
    if (!ctx->in_substage) return fail(ctx, SWMHD_ERR_STATE, "no substage in flight");
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaStreamWaitEvent(ctx->main, ctx->ev_edge, 0));
    if (ctx->pending_slot >= 0) {       // fold the per-tile partials of edges + interior (fixed order)
        ctx->launches++;
        CK(launch_diag_final(ctx->d_partials, ctx->ntiles, ctx->d_stage, ctx->d_diag + (size_t)ctx->pending_slot * NDIAG, ctx->main));
        ctx->pending_slot = -1;
    }
    ctx->cur = 1 - ctx->cur;
    ctx->in_substage = false;
    if (stage < 1 || stage > 3) return fail(ctx, SWMHD_ERR_ARG, "stage must be 1, 2 or 3");
    tick(ctx, ctx->pending_dt, stage);
    return SWMHD_OK;
}

extern "C" int swmhd_exchange_rows(swmhd_ctx *ctx, int field, int which, void **dev_ptr, int *nrows, size_t *row_doubles) {
    if (!ctx || !dev_ptr || !nrows || !row_doubles) return SWMHD_ERR_ARG;
    if (field < 0 || field > 3 || which < 0 || which > 7) return fail(ctx, SWMHD_ERR_ARG, "bad field/which");
    int buf = (which >= 4) ? ctx->cur : 1 - ctx->cur;   // 0-3: state being written, 4-7: current state
    int w = which & 3;
    int row = (w == 0) ? 3 : (w == 1) ? ctx->Ny : (w == 2) ? 0 : ctx->Ny + 3;
    *dev_ptr = (void *)(ctx->U[buf][field] + (size_t)row * ctx->P);
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
    CK(cudaSetDevice(ctx->cfg.device));
    std::vector<double> host((size_t)count * NDIAG);
    CK(cudaStreamSynchronize(ctx->edge));
    CK(cudaMemcpyAsync(host.data(), ctx->d_diag + (size_t)first * NDIAG, host.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->main));
    CK(cudaStreamSynchronize(ctx->main));
    for (int n = 0; n < count; n++) diag_fill(ctx, &host[(size_t)n * NDIAG], &out[n]);
    return SWMHD_OK;
}

extern "C" int swmhd_sync(swmhd_ctx *ctx) {
    if (!ctx) return SWMHD_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaStreamSynchronize(ctx->edge));
    CK(cudaStreamSynchronize(ctx->main));
    return SWMHD_OK;
}
