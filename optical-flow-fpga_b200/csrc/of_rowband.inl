// Row-band (multi-GPU) pyramidal LK, native driver.  Included at the end of of_api.cu.
//
// One frame pair, every level's rows split over `world` ranks (one process or thread per GPU).  A
// rank computes only its rows of every pyramid level and of every refinement iteration; what the
// neighbours need travels through peer memory (peer.cu): the owned rows of each pyramid level and
// of each level's final flow are stored straight into every rank's arena, and the residual sums of
// every iteration are all-reduced through exchange slots, so one call enqueues the whole
// coarse-to-fine computation without a host round trip, a NCCL call or a Python loop.
//
// Row bookkeeping (same as distributed.py, which keeps the NCCL version of this driver):
//   * inside a level there is no halo exchange: iteration i of I runs on the band extended by
//     4 * (I - 1 - i) rows (3 rows of flow_in reach + 1 for an even band start), so after the last
//     iteration exactly the owned rows are valid;
//   * the upsampled start flow is produced on the band extended by 4 * (I + 1) rows;
//   * the warp gather has unbounded reach, hence the all-gather of every pyramid level.
// Results are bit-identical to the single-GPU driver (same kernels, same row pairing); only the
// float64 residual sums are associated differently (per rank, then rank order).

namespace {

// Rows by which a band's computed range shrinks per iteration: an iteration reads flow_in window_size / 2 + 1 rows
// beyond the rows it writes, one more row is left for the even alignment of the band start, rounded up to even
// (4 for windows up to 5, 6 for window 7).  distributed.py's band_grow is the same rule.
inline int rb_grow(int window) {
    const int g = (window / 2 + 2 + 1) & ~1;
    return g < 4 ? 4 : g;
}

struct RowbandCtx {
    int rank = 0, world = 1, H = 0, W = 0, L = 1, window = 5, iters = 3, mode = OF_MODE_FAST;
    int radius = 0;
    double gw[2 * OF_MAX_GAUSS_RADIUS + 1];
    std::vector<int> h, w;
    // arena offsets (bytes)
    size_t flag_off = 0, xchg_off = 0, err_off = 0, run_off = 0, sel_off = 0, done_off = 0, itx_off = 0, resid_off = 0;
    std::vector<size_t> prev_off, curr_off, au_off, av_off, bu_off, bv_off, gu_off, gv_off;
    size_t warped_off = 0, partial_off = 0, cnt_off = 0, total = 0;
    char* base = nullptr;
    char* peer[PEER_MAX_WORLD];
    bool ipc_opened[PEER_MAX_WORLD];
    bool peers_set = false;
    unsigned long long ops_per_run = 1;
    int device = 0;
    // Levels with at most this many pixels are computed whole on every rank ("replicated"): their
    // kernels are launch-latency bound at any band height, so splitting them buys nothing, while
    // replication removes their collectives (no all-reduce per iteration, no gathers).
    long long replicate_px = 600000;
    // how long a kernel waits for a peer before it gives up and sets the error word: far beyond any
    // step, short enough that a dead peer does not wedge the device.  Ranks must therefore enter
    // of_rowband_run within this time of each other.
    unsigned long long timeout_ns = 4000000000ULL;
    bool error_seen = false;  // the host has read a non-zero error word: the next run clears it
};

void rb_shard(int n, int rank, int world, int* a, int* b) {
    const int base = n / world, extra = n % world;
    *a = rank * base + (rank < extra ? rank : extra);
    *b = *a + base + (rank < extra ? 1 : 0);
}

PeerView rb_view(const RowbandCtx& c) {
    PeerView pv;
    for (int r = 0; r < PEER_MAX_WORLD; ++r) pv.peer[r] = r < c.world ? c.peer[r] : nullptr;
    pv.world = c.world;
    pv.rank = c.rank;
    pv.flag_off = c.flag_off;
    pv.xchg_off = c.xchg_off;
    pv.err = reinterpret_cast<int*>(c.base + c.err_off);
    pv.timeout_ns = c.timeout_ns;
    pv.run_id = reinterpret_cast<const unsigned long long*>(c.base + c.run_off);
    pv.ops_per_run = c.ops_per_run;
    return pv;
}

}  // namespace

extern "C" {

int of_rowband_create(of_rowband_t** out, int rank, int world, int height, int width, int levels, int window,
                      int iterations, int mode, const double* gauss_weights, int gauss_radius) {
    if (!out) return fail(OF_ERR_INVALID_ARGUMENT, "null pointer");
    *out = nullptr;
    if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world)
        return fail(OF_ERR_INVALID_ARGUMENT, "world must be in 1..8 and 0 <= rank < world");
    if (height < 1 || width < 1 || (long long)height * width > (1LL << 31) - 1)
        return fail(OF_ERR_INVALID_ARGUMENT, "bad frame size");
    OF_TRY(check_window(window));
    if (mode != OF_MODE_EXACT && mode != OF_MODE_FAST) return fail(OF_ERR_INVALID_ARGUMENT, "unknown mode");
    if (iterations < 0 || iterations > 1000) return fail(OF_ERR_INVALID_ARGUMENT, "num_iterations must be in 0..1000");
    if (levels < 1 || levels > 16) return fail(OF_ERR_INVALID_ARGUMENT, "num_levels must be in 1..16");
    if (levels > 1 && (!gauss_weights || gauss_radius < 0 || gauss_radius > OF_MAX_GAUSS_RADIUS))
        return fail(OF_ERR_INVALID_ARGUMENT, "gaussian weights missing or radius out of range");
    OF_TRY(need_device());
    RowbandCtx* c = new RowbandCtx();
    c->rank = rank;
    c->world = world;
    c->H = height;
    c->W = width;
    c->L = levels;
    c->window = window;
    c->iters = iterations;
    c->mode = mode;
    c->radius = gauss_radius;
    for (int i = 0; i < 2 * gauss_radius + 1 && gauss_weights; ++i) c->gw[i] = gauss_weights[i];
    for (int r = 0; r < PEER_MAX_WORLD; ++r) {
        c->peer[r] = nullptr;
        c->ipc_opened[r] = false;
    }
    c->h.assign(levels, 0);
    c->w.assign(levels, 0);
    c->h[0] = height;
    c->w[0] = width;
    for (int k = 1; k < levels; ++k) {
        c->h[k] = c->h[k - 1] / 2;
        c->w[k] = c->w[k - 1] / 2;
        if (c->h[k] < 1 || c->w[k] < 1) {
            delete c;
            return fail(OF_ERR_INVALID_ARGUMENT, "frame too small for this many pyramid levels");
        }
    }
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off += align_up(bytes);
        return o;
    };
    c->flag_off = take(sizeof(unsigned long long) * PEER_SLOTS * PEER_MAX_WORLD);
    c->xchg_off = take(sizeof(double) * PEER_SLOTS * PEER_MAX_WORLD * 2);
    c->err_off = take(sizeof(int));
    c->run_off = take(sizeof(unsigned long long));
    // collectives per run: one barrier per coarser pyramid level, one all-reduce per iteration and
    // level, one barrier per level's flow gather
    c->ops_per_run = (unsigned long long)(levels - 1) + (unsigned long long)levels * iterations + levels + 1;
    c->sel_off = take(sizeof(int) * levels);
    c->done_off = take(sizeof(int) * levels);
    c->itx_off = take(sizeof(int) * levels);
    c->resid_off = take(sizeof(float) * levels * (iterations > 0 ? iterations : 1) * 2);
    c->cnt_off = take(sizeof(unsigned));  // ticket counter of the fused iteration tail (zero between launches)
    for (auto* v : {&c->prev_off, &c->curr_off, &c->au_off, &c->av_off, &c->bu_off, &c->bv_off, &c->gu_off, &c->gv_off})
        v->assign(levels, 0);
    int max_blocks = 1;
    for (int k = 0; k < levels; ++k) {
        const size_t plane = (size_t)c->h[k] * c->w[k] * sizeof(float);
        if (k >= 1) {
            c->prev_off[k] = take(plane);
            c->curr_off[k] = take(plane);
        }
        c->au_off[k] = take(plane);
        c->av_off[k] = take(plane);
        c->bu_off[k] = take(plane);
        c->bv_off[k] = take(plane);
        c->gu_off[k] = take(plane);
        c->gv_off[k] = take(plane);
        const int nb = lk_tile_blocks_per_pair(c->h[k], c->w[k]);
        if (nb > max_blocks) max_blocks = nb;
    }
    c->warped_off = take((size_t)height * width * sizeof(float));
    c->partial_off = take((size_t)max_blocks * 2 * sizeof(double));
    c->total = off;
    cudaGetDevice(&c->device);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->base), c->total);
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete c;
        return fail(OF_ERR_OUT_OF_MEMORY, std::string("cudaMalloc of the row-band arena: ") + cudaGetErrorString(e));
    }
    // control words (flags, exchange slots, error) start at zero before any peer can see the arena
    e = cudaMemset(c->base, 0, c->au_off[0]);  // everything before the first plane
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(c->base);
        delete c;
        return fail(OF_ERR_CUDA, std::string("initialising the row-band arena: ") + cudaGetErrorString(e));
    }
    c->peer[rank] = c->base;
    if (world == 1) c->peers_set = true;
    *out = reinterpret_cast<of_rowband_t*>(c);
    return OF_OK;
}

size_t of_rowband_arena_bytes(const of_rowband_t* ctx) { return ctx ? reinterpret_cast<const RowbandCtx*>(ctx)->total : 0; }

void* of_rowband_arena(const of_rowband_t* ctx) { return ctx ? reinterpret_cast<const RowbandCtx*>(ctx)->base : nullptr; }

int of_rowband_ipc_handle(const of_rowband_t* ctx, void* handle_out) {
    if (!ctx || !handle_out) return fail(OF_ERR_INVALID_ARGUMENT, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == OF_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t hdl;
    OF_CUDA(cudaIpcGetMemHandle(&hdl, reinterpret_cast<const RowbandCtx*>(ctx)->base));
    memcpy(handle_out, &hdl, sizeof(hdl));
    return OF_OK;
}

int of_rowband_open_peers_ipc(of_rowband_t* ctx, const void* handles) {
    if (!ctx || !handles) return fail(OF_ERR_INVALID_ARGUMENT, "null pointer");
    RowbandCtx* c = reinterpret_cast<RowbandCtx*>(ctx);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t hdl;
        memcpy(&hdl, static_cast<const char*>(handles) + (size_t)r * OF_IPC_HANDLE_BYTES, sizeof(hdl));
        void* p = nullptr;
        OF_CUDA(cudaIpcOpenMemHandle(&p, hdl, cudaIpcMemLazyEnablePeerAccess));
        c->peer[r] = static_cast<char*>(p);
        c->ipc_opened[r] = true;
    }
    c->peers_set = true;
    return OF_OK;
}

int of_rowband_set_peers(of_rowband_t* ctx, void* const* bases) {
    if (!ctx || !bases) return fail(OF_ERR_INVALID_ARGUMENT, "null pointer");
    RowbandCtx* c = reinterpret_cast<RowbandCtx*>(ctx);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        if (!bases[r]) return fail(OF_ERR_INVALID_ARGUMENT, "null peer arena");
        c->peer[r] = static_cast<char*>(bases[r]);
    }
    c->peers_set = true;
    return OF_OK;
}

int of_rowband_set_replicate_pixels(of_rowband_t* ctx, long long pixels) {
    if (!ctx || pixels < 0) return fail(OF_ERR_INVALID_ARGUMENT, "bad argument");
    reinterpret_cast<RowbandCtx*>(ctx)->replicate_px = pixels;
    return OF_OK;
}

int of_rowband_set_timeout_ms(of_rowband_t* ctx, int milliseconds) {
    if (!ctx || milliseconds < 1 || milliseconds > 600000) return fail(OF_ERR_INVALID_ARGUMENT, "timeout must be in 1 ms .. 10 min");
    reinterpret_cast<RowbandCtx*>(ctx)->timeout_ns = (unsigned long long)milliseconds * 1000000ULL;
    return OF_OK;
}

int of_rowband_run(of_rowband_t* ctx, const float* prev, const float* curr, float* u, float* v, void* stream) {
    if (!ctx) return fail(OF_ERR_INVALID_ARGUMENT, "null context");
    RowbandCtx& c = *reinterpret_cast<RowbandCtx*>(ctx);
    if (!c.peers_set) return fail(OF_ERR_INVALID_ARGUMENT, "peer arenas have not been connected");
    OF_TRY(check_frame(prev, curr, c.H, c.W));
    if ((u == nullptr) != (v == nullptr)) return fail(OF_ERR_INVALID_ARGUMENT, "pass both u and v or neither");
    int dev_now = -1;
    OF_CUDA(cudaGetDevice(&dev_now));
    if (dev_now != c.device) return fail(OF_ERR_INVALID_ARGUMENT, "of_rowband_run must be called with the context's device current");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Counter cnt;
    const PeerView pv = rb_view(c);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(c.base + off); };
    int* sel = reinterpret_cast<int*>(c.base + c.sel_off);
    int* done = reinterpret_cast<int*>(c.base + c.done_off);
    int* itx = reinterpret_cast<int*>(c.base + c.itx_off);
    float* resid = reinterpret_cast<float*>(c.base + c.resid_off);
    double* partial = reinterpret_cast<double*>(c.base + c.partial_off);
    const int L = c.L, iters = c.iters, world = c.world, rank = c.rank;
    unsigned long long op = 0;  // index of the next collective of this run
    OF_CUDA(launch_peer_begin_run(pv, c.error_seen, &cnt.n, st));
    c.error_seen = false;

    // per-run control words (not the flags / exchange slots: their sequence numbers keep growing)
    OF_CUDA(cudaMemsetAsync(c.base + c.sel_off, 0, c.resid_off + align_up(sizeof(float) * L * (iters > 0 ? iters : 1) * 2) - c.sel_off, st));

    // ---- Gaussian pyramids: own rows of each level, stored into every rank's arena ---------
    std::vector<const float*> lp(L), lc(L);
    lp[0] = prev;
    lc[0] = curr;
    auto replicated = [&](int k) { return world > 1 && (long long)c.h[k] * c.w[k] <= c.replicate_px; };
    PeerView pv_self = pv;  // "gather" of a replicated level: a local copy, no peers
    pv_self.world = 1;
    pv_self.rank = 0;
    pv_self.peer[0] = c.base;
    for (int k = 1; k < L; ++k) {
        int a, b;
        rb_shard(c.h[k], rank, world, &a, &b);
        if (replicated(k)) {
            a = 0;
            b = c.h[k];
        }
        lp[k] = F(c.prev_off[k]);
        lc[k] = F(c.curr_off[k]);
        if (b > a) {
            OF_CUDA(launch_pyramid_down(lp[k - 1], F(c.prev_off[k]), 1, c.h[k - 1], c.w[k - 1], c.h[k], c.w[k], c.gw, c.radius,
                                        a, b, &cnt.n, st, c.mode == OF_MODE_FAST));
            OF_CUDA(launch_pyramid_down(lc[k - 1], F(c.curr_off[k]), 1, c.h[k - 1], c.w[k - 1], c.h[k], c.w[k], c.gw, c.radius,
                                        a, b, &cnt.n, st, c.mode == OF_MODE_FAST));
        }
        if (world > 1 && !replicated(k)) {
            const size_t first = (size_t)a * c.w[k], count = (size_t)(b - a) * c.w[k];
            OF_CUDA(launch_peer_push_rows(pv, lp[k], nullptr, nullptr, 0, c.prev_off[k], first, count, true, &cnt.n, st));
            OF_CUDA(launch_peer_push_rows(pv, lc[k], nullptr, nullptr, 0, c.curr_off[k], first, count, true, &cnt.n, st));
            OF_CUDA(launch_peer_sync(pv, op++, &cnt.n, st));
        }
    }

    // ---- coarse to fine --------------------------------------------------------------------
    const int start = iters & 1;  // a pair that runs all its iterations ends in buffer 0
    const int kc = L - 1;
    auto fu = [&](int k, int idx) { return F(idx ? c.bu_off[k] : c.au_off[k]); };
    auto fv = [&](int k, int idx) { return F(idx ? c.bv_off[k] : c.av_off[k]); };
    const size_t coarse_bytes = (size_t)c.h[kc] * c.w[kc] * sizeof(float);
    OF_CUDA(cudaMemsetAsync(fu(kc, start), 0, coarse_bytes, st));
    OF_CUDA(cudaMemsetAsync(fv(kc, start), 0, coarse_bytes, st));

    for (int k = kc; k >= 0; --k) {
        const int h = c.h[k], w = c.w[k];
        const int ref_level = kc - k;  // the reference counts levels from the coarsest
        int a, b;
        rb_shard(h, rank, world, &a, &b);
        const bool repl = replicated(k);
        if (repl) {
            a = 0;
            b = h;
        }
        int* sel_k = sel + k;
        int* done_k = done + k;
        if (k < kc && b > a) {
            const int reach = rb_grow(c.window) * (iters > 1 ? iters : 1) + rb_grow(c.window);
            const int lo = a - reach < 0 ? 0 : a - reach, hi = b + reach > h ? h : b + reach;
            if (replicated(k + 1))  // the coarser level lives whole in this rank's ping-pong buffers
                OF_CUDA(launch_upsample_flow(fu(k + 1, 0), fv(k + 1, 0), fu(k + 1, 1), fv(k + 1, 1), sel + (k + 1), start,
                                             fu(k, start), fv(k, start), 1, c.h[k + 1], c.w[k + 1], h, w, lo, hi, &cnt.n, st));
            else
                OF_CUDA(launch_upsample_flow(F(c.gu_off[k + 1]), F(c.gv_off[k + 1]), nullptr, nullptr, nullptr, 0,
                                             fu(k, start), fv(k, start), 1, c.h[k + 1], c.w[k + 1], h, w, lo, hi, &cnt.n, st));
        }
        RefineArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.prev = lp[k];
        ra.curr = lc[k];
        ra.flow_u[0] = fu(k, 0);
        ra.flow_v[0] = fv(k, 0);
        ra.flow_u[1] = fu(k, 1);
        ra.flow_v[1] = fv(k, 1);
        ra.sel = sel_k;
        ra.sel_xor = start;
        ra.done = done_k;
        ra.partial = partial;
        ra.H = h;
        ra.W = w;
        ra.window = c.window;
        ra.own_lo = a;
        ra.own_hi = b;
        const bool fast_level = (c.mode == OF_MODE_FAST) && lk_refine_supported(ra, c.window);
        for (int it = 0; it < iters; ++it) {
            const int ext = rb_grow(c.window) * (iters - 1 - it);
            int lo = a - ext < 0 ? 0 : a - ext;
            lo -= lo & 1;  // even start: rows pair up identically on every rank
            const int hi = b + ext > h ? h : b + ext;
            int blocks = 0;
            if (b > a) {
                ra.row_lo = lo;
                ra.row_hi = hi;
                if (fast_level && refine_split()) {
                    // the marching kernel's last warp reduces, all-reduces (split levels) and decides
                    ra.tail.counter = reinterpret_cast<unsigned*>(c.base + c.cnt_off);
                    ra.tail.peers = repl ? 0 : 1;
                    if (!repl) fill_peer_sync(ra.tail.sync, pv, op++);
                    ra.tail.n_pixels = (double)h * (double)w;
                    ra.tail.sel = sel_k;
                    ra.tail.done = done_k;
                    ra.tail.iters_executed = itx + ref_level;
                    ra.tail.iters_pair_stride = 0;
                    ra.tail.residuals = resid + (size_t)ref_level * (iters > 0 ? iters : 1) * 2;
                    ra.tail.resid_pair_stride = 0;
                    ra.tail.iteration = it;
                    OF_CUDA(launch_refine_split_form(ra, F(c.warped_off), 1, &cnt.n, st));
                    continue;
                }
                if (fast_level) {
                    if (refine_split())
                        OF_CUDA(launch_refine_split_form(ra, F(c.warped_off), 1, &cnt.n, st));
                    else
                        OF_CUDA(launch_lk_refine(ra, 1, &cnt.n, st));
                    blocks = lk_refine_units_per_pair(1, hi - lo, w);
                } else {
                    TileArgs t;
                    memset(&t, 0, sizeof(t));
                    t.in0 = lp[k];
                    t.in1 = lc[k];
                    for (int i = 0; i < 2; ++i) {
                        t.flow_u[i] = ra.flow_u[i];
                        t.flow_v[i] = ra.flow_v[i];
                    }
                    t.sel = sel_k;
                    t.sel_xor = start;
                    t.done = done_k;
                    t.partial = partial;
                    t.H = h;
                    t.W = w;
                    t.row_lo = lo;
                    t.row_hi = hi;
                    t.own_lo = a;
                    t.own_hi = b;
                    if (exact_refine_split()) {
                        // warp the rows the band's Sobel / window halo can touch, then the tile kernel on the plane
                        const int halo = c.window / 2 + 1;
                        OF_CUDA(launch_warp_rows(ra, F(c.warped_off), lo - halo < 0 ? 0 : lo - halo,
                                                 hi + halo > h ? h : hi + halo, true, 1, &cnt.n, st));
                        t.in1 = F(c.warped_off);
                        OF_CUDA(launch_lk_tile(SRC_WARPED, c.window, t, 1, &cnt.n, st));
                    } else {
                        OF_CUDA(launch_lk_tile(SRC_WARP, c.window, t, 1, &cnt.n, st));
                    }
                    blocks = lk_tile_blocks_per_pair(hi - lo, w);
                }
            }
            if (repl) {
                // whole level on this rank: the single-GPU driver's convergence kernel, no exchange
                IterFinalizeArgs f;
                f.partial = partial;
                f.blocks_per_pair = blocks;
                f.H = h;
                f.W = w;
                f.sel = sel_k;
                f.done = done_k;
                f.iters_executed = itx + ref_level;
                f.iters_pair_stride = L;
                f.residuals = resid + (size_t)ref_level * (iters > 0 ? iters : 1) * 2;
                f.resid_pair_stride = (size_t)L * (iters > 0 ? iters : 1) * 2;
                f.iteration = it;
                OF_CUDA(launch_iter_finalize(f, 1, &cnt.n, st));
            } else {
                OF_CUDA(launch_peer_allreduce_update(pv, op++, b > a ? partial : nullptr, blocks, (double)h * (double)w,
                                                     sel_k, done_k, itx + ref_level,
                                                     resid + (size_t)ref_level * (iters > 0 ? iters : 1) * 2, it, &cnt.n, st));
            }
        }
        // level done: the owned rows of the current ping-pong buffer -> the gathered plane of every rank
        // (a replicated level stays in its ping-pong buffers; only the finest one is copied, locally,
        // to where of_rowband_result points)
        const size_t first = (size_t)a * w, count = (size_t)(b - a) * w;
        if (!repl && k > 0) {
            // Flow of a coarser level: a rank needs only the coarse rows its upsample of level k - 1 reads -- its
            // own band plus the halo of its extended band -- so every rank receives just that part of my rows
            // (a halo exchange with the neighbours; all of it only when the bands are thinner than the halo).
            size_t lo_el[PEER_MAX_WORLD], hi_el[PEER_MAX_WORLD];
            const int hf = c.h[k - 1];
            const double sy = hf > 1 ? (double)(h - 1) / (double)(hf - 1) : 0.0;  // np.linspace step of upsample_flow
            const int reach = rb_grow(c.window) * (iters > 1 ? iters : 1) + rb_grow(c.window);
            for (int r = 0; r < PEER_MAX_WORLD; ++r) {
                lo_el[r] = hi_el[r] = 0;
                if (r >= world) continue;
                int fa, fb;
                rb_shard(hf, r, world, &fa, &fb);
                if (fb <= fa) continue;
                const int flo = fa - reach < 0 ? 0 : fa - reach, fhi = fb + reach > hf ? hf : fb + reach;
                // target rows [flo, fhi) blend coarse rows floor(y * sy) and the one below; one row of margin
                // on both sides (the product y * sy may round across an integer)
                long long need_lo = (long long)floor((double)flo * sy) - 1;
                long long need_hi = (long long)floor((double)(fhi - 1) * sy) + 3;
                if (need_lo < 0) need_lo = 0;
                if (need_hi > h) need_hi = h;
                lo_el[r] = (size_t)need_lo * w;
                hi_el[r] = (size_t)need_hi * w;
            }
            OF_CUDA(launch_peer_push_ranges(pv, fu(k, 0), fu(k, 1), sel_k, start, c.gu_off[k], first, count, lo_el, hi_el, &cnt.n, st));
            OF_CUDA(launch_peer_push_ranges(pv, fv(k, 0), fv(k, 1), sel_k, start, c.gv_off[k], first, count, lo_el, hi_el, &cnt.n, st));
            if (world > 1) OF_CUDA(launch_peer_sync(pv, op++, &cnt.n, st));
        } else if (!repl) {
            // the finest level: the final gather, every rank receives the whole flow
            OF_CUDA(launch_peer_push_rows(pv, fu(k, 0), fu(k, 1), sel_k, start, c.gu_off[k], first, count, false, &cnt.n, st));
            OF_CUDA(launch_peer_push_rows(pv, fv(k, 0), fv(k, 1), sel_k, start, c.gv_off[k], first, count, false, &cnt.n, st));
            if (world > 1) OF_CUDA(launch_peer_sync(pv, op++, &cnt.n, st));
        } else if (k == 0) {
            OF_CUDA(launch_peer_push_rows(pv_self, fu(k, 0), fu(k, 1), sel_k, start, c.gu_off[k], first, count, false, &cnt.n, st));
            OF_CUDA(launch_peer_push_rows(pv_self, fv(k, 0), fv(k, 1), sel_k, start, c.gv_off[k], first, count, false, &cnt.n, st));
        }
    }
    if (u) {
        const size_t bytes = (size_t)c.H * c.W * sizeof(float);
        OF_CUDA(cudaMemcpyAsync(u, F(c.gu_off[0]), bytes, cudaMemcpyDeviceToDevice, st));
        OF_CUDA(cudaMemcpyAsync(v, F(c.gv_off[0]), bytes, cudaMemcpyDeviceToDevice, st));
    }
    return OF_OK;
}

int of_rowband_result(const of_rowband_t* ctx, const float** u, const float** v) {
    if (!ctx || !u || !v) return fail(OF_ERR_INVALID_ARGUMENT, "null pointer");
    const RowbandCtx& c = *reinterpret_cast<const RowbandCtx*>(ctx);
    *u = reinterpret_cast<const float*>(c.base + c.gu_off[0]);
    *v = reinterpret_cast<const float*>(c.base + c.gv_off[0]);
    return OF_OK;
}

int of_rowband_trace(of_rowband_t* ctx, int* iters_executed, float* residuals, int* error, void* stream) {
    if (!ctx) return fail(OF_ERR_INVALID_ARGUMENT, "null context");
    RowbandCtx& c = *reinterpret_cast<RowbandCtx*>(ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OF_CUDA(cudaStreamSynchronize(st));
    if (iters_executed) OF_CUDA(cudaMemcpy(iters_executed, c.base + c.itx_off, sizeof(int) * c.L, cudaMemcpyDeviceToHost));
    if (residuals && c.iters > 0)
        OF_CUDA(cudaMemcpy(residuals, c.base + c.resid_off, sizeof(float) * c.L * c.iters * 2, cudaMemcpyDeviceToHost));
    int err = 0;
    OF_CUDA(cudaMemcpy(&err, c.base + c.err_off, sizeof(int), cudaMemcpyDeviceToHost));
    if (error) *error = err;
    if (err != 0) {
        c.error_seen = true;  // reported: the next of_rowband_run starts with a clean error word
        return fail(OF_ERR_PEER_TIMEOUT, "row-band run: a peer rank did not answer within the time-out; the flow of this run is invalid");
    }
    return OF_OK;
}

int of_rowband_status(of_rowband_t* ctx, void* stream) { return of_rowband_trace(ctx, nullptr, nullptr, nullptr, stream); }

int of_rowband_destroy(of_rowband_t* ctx) {
    if (!ctx) return OF_OK;
    RowbandCtx* c = reinterpret_cast<RowbandCtx*>(ctx);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (c->ipc_opened[r] && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->base) cudaFree(c->base);
    cudaGetLastError();
    delete c;
    return OF_OK;
}

}  // extern "C"
