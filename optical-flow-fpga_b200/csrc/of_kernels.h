// Internal launch interface between the C-ABI layer (of_api.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>

namespace ofb {

// a kernel launch (tests/host_emul/ runs the kernel's threads as OS threads instead)
#ifndef OF_LAUNCH
#define OF_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

// Opt-in to more than 48 KB of dynamic shared memory, once per (kernel, device): the attribute
// belongs to the device's context, so a process that drives several GPUs sets it on each.
struct SmemOptIn {
    std::atomic<unsigned long long> done{0};  // one bit per device ordinal
    template <typename Kernel>
    cudaError_t ensure(Kernel kernel, size_t bytes) {
#ifdef OF_HOST_EMULATION  // tests/host_emul/: no device, nothing to opt in to
        (void)kernel;
        (void)bytes;
        return cudaSuccess;
#else
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const unsigned long long bit = dev < 64 ? 1ULL << dev : 0ULL;
        if (bit && (done.load(std::memory_order_relaxed) & bit)) return cudaSuccess;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess && bit) done.fetch_or(bit, std::memory_order_relaxed);
        return e;
#endif
    }
};

// ---- tail of a refinement iteration, optionally fused into the marching kernel -----------
constexpr int PEER_MAX_WORLD = 8;  // one NVSwitch domain
constexpr int PEER_SLOTS = 64;     // flag / exchange slots, indexed by sequence number
// what a kernel needs to signal / wait for the other ranks (peer.cu, peer_device.cuh)
struct PeerSync {
    char* peer[PEER_MAX_WORLD];  // arena base of every rank as mapped in this process
    int world, rank;
    size_t flag_off;  // unsigned long long flags[PEER_SLOTS][PEER_MAX_WORLD] in every arena
    size_t xchg_off;  // double xchg[PEER_SLOTS][PEER_MAX_WORLD][2]
    // sequence number of this collective = *run_id * ops_per_run + op: the run counter lives in
    // device memory (bumped by the first kernel of every run), so a captured CUDA graph of a run
    // can be replayed and still produces fresh, growing sequence numbers
    const unsigned long long* run_id;
    unsigned long long ops_per_run, op;
    int* err;
    unsigned long long timeout_ns;
};
// Convergence bookkeeping of one iteration.  counter != nullptr asks the marching kernel to do it
// itself: the last warp of a pair to finish (ticket counter) reduces the pair's partials in a fixed
// order, all-reduces them over the ranks if `peers`, and applies the test -- no separate launch.
struct IterTail {
    unsigned* counter;  // [batch], zero between launches
    int peers;          // 0: this GPU holds the whole level; 1: all-reduce through `sync`
    PeerSync sync;
    double n_pixels;
    int* sel;
    int* done;
    int* iters_executed;  // nullable; pair b at iters_executed[b * iters_pair_stride]
    int iters_pair_stride;
    float* residuals;     // nullable; pair b, iteration i at residuals[b * resid_pair_stride + 2 * i]
    size_t resid_pair_stride;
    int iteration;
};

// ---- K1 fast: warp-marching fused single-scale LK (lk_march.cu) ------------------------
struct MarchArgs {
    const float* prev;
    const float* curr;
    float* u;
    float* v;
    int H, W;
    int n_strips, n_bands, band_rows;
    long long n_units;
    // REFINE variant (split refinement iteration): `curr` is the already warped current frame,
    // the result is flow[out] = flow[in] + d with in = sel[pair] ^ sel_xor, rows [row_lo, row_hi),
    // and sum|du|, sum|dv| over rows [own_lo, own_hi) go to partial[pair][unit][2].
    float* flow_u[2];
    float* flow_v[2];
    const int* sel;
    int sel_xor;
    const int* done;
    double* partial;
    int row_lo, row_hi, own_lo, own_hi;
    // fixed-point flavour: int16 S8.7 outputs, and whether to mirror the 9-bit average of gradient_compute.sv:116
    int16_t* u16;
    int16_t* v16;
    int fx_quirk;
    IterTail tail;  // REFINE: tail.counter != nullptr fuses the iteration's convergence step into the kernel
    // REFINE, optional: the kernel also warps the NEXT iteration's input -- warped_next[y][x] = bilinear(warp_src,
    // y + flow_out_v, x + flow_out_u) for every pixel it writes (the flow is in registers, warp_rows' arithmetic)
    const float* warp_src;  // the level's current frame, unwarped
    float* warped_next;     // nullable
};
bool lk_march_supported(int H, int W, int window);
// force_path: 0 = TMA when the pointers allow it, 1 = TMA or error, 2 = plain global loads
cudaError_t launch_lk_march(const float* prev, const float* curr, float* u, float* v, int batch, int H, int W, int window,
                            int force_path, int* launches, cudaStream_t stream);

// uint8 ingest of the same kernel (10 B per pixel): needs W % 16 == 0 and 16-byte aligned planes
bool lk_march_u8_supported(const uint8_t* prev, const uint8_t* curr, const float* u, const float* v, int H, int W, int window);
cudaError_t launch_lk_march_u8(const uint8_t* prev, const uint8_t* curr, float* u, float* v, int batch, int H, int W,
                               int window, int* launches, cudaStream_t stream);
// the RTL's fixed-point datapath on the same marching kernel (6 B per pixel); same frame requirements
bool lk_march_fx_supported(const uint8_t* prev, const uint8_t* curr, const int16_t* u, const int16_t* v, int H, int W);
cudaError_t launch_lk_march_fx(const uint8_t* prev, const uint8_t* curr, int16_t* u, int16_t* v, int batch, int H, int W,
                               int mirror_avg_quirk, int* launches, cudaStream_t stream);
cudaError_t launch_u8_to_f32(const uint8_t* src, float* dst, size_t n, int* launches, cudaStream_t stream);

// ---- K3 fast: warp-marching refinement iteration (lk_march.cu) --------------------------
struct RefineArgs {
    const float* prev;  // level image of the previous frame   [B][H][W]
    const float* curr;  // level image of the current frame (gathered through the flow)
    float* flow_u[2];   // ping-pong flow buffers, sel[pair] = index of the current one
    float* flow_v[2];
    const int* sel;
    int sel_xor;        // current buffer = sel[pair] ^ sel_xor (lets a level start in buffer 1)
    const int* done;
    double* partial;    // [pair][units_per_pair][2]
    int H, W;
    int window;         // 5 or 7 (lk_refine_supported)
    // Row-band mode: only rows [row_lo, row_hi) are computed (the frame and the flow buffers are
    // still full-size), and only rows [own_lo, own_hi) enter the |du|, |dv| sums.  Whole frame:
    // 0, H, 0, H.  row_lo must be even so that rows pair up the same way on every rank.
    int row_lo, row_hi, own_lo, own_hi;
    int n_strips, n_bands, band_rows;  // filled by the launcher
    long long n_units;
    IterTail tail;  // split form only: counter != nullptr -> the marching kernel finishes the iteration itself
    // split form only.  warped_ready: `warped` already holds warp(curr, flow_in) on the rows this launch reads (the
    // previous iteration's launch wrote it through its warped_next) -- no warp_rows launch.  warped_next (nullable):
    // where this launch leaves warp(curr, flow_out) of the rows it computes, for the next iteration.
    int warped_ready;
    float* warped_next;
};
bool lk_refine_supported(const RefineArgs& a, int window);
int lk_refine_units_per_pair(int batch, int rows, int W);
cudaError_t launch_lk_refine(const RefineArgs& a, int batch, int* launches, cudaStream_t stream);
// Split form of the same iteration: warp_rows (gather the current frame through the flow into
// `warped`, rows [row_lo - 3, row_hi + 3)) followed by the K1 marching kernel in REFINE mode.
cudaError_t launch_lk_refine_split(const RefineArgs& a, float* warped, int batch, int* launches, cudaStream_t stream);
// Warp-specialised form (window 5): producer warps gather the warped rows into the marching warps' ring stages; no
// warped plane, one launch.  Same bits as the split form; honours a.tail like it.
cudaError_t launch_lk_refine_ws(const RefineArgs& a, int batch, int* launches, cudaStream_t stream);
// The warp alone: warped[pair][y][x] = bilinear(curr, y + v, x + u) for rows [row_lo, row_hi) of every pair
// that has not converged, flow = the current ping-pong buffer of `a`.  exact = float64 sample fractions
// (warp_image's bits for any flow); false = the float32 fractions of the fast path.
cudaError_t launch_warp_rows(const RefineArgs& a, float* warped, int row_lo, int row_hi, bool exact, int batch,
                             int* launches, cudaStream_t stream);

// ---- K1/K3 exact: tile kernel in the reference's operation order (lk_tile.cu) ----------
enum TileSource { SRC_FRAMES = 0, SRC_WARP = 1, SRC_GRADS = 2, SRC_WARPED = 3 };
struct TileArgs {
    // SRC_FRAMES: in0 = prev, in1 = curr.  SRC_WARP: in0 = prev level, in1 = curr level (gathered
    // through the flow).  SRC_GRADS: in0 = Ix, in1 = Iy, in2 = It.  SRC_WARPED: SRC_WARP with
    // in1 = the current level already warped through the current flow (launch_warp_rows, exact).
    const float* in0;
    const float* in1;
    const float* in2;
    // SRC_WARP / SRC_WARPED: ping-pong flow buffers; sel[pair] says which one is current (0 = A).
    float* flow_u[2];
    float* flow_v[2];
    const int* sel;    // nullable: A is the input, B the output
    int sel_xor;       // current buffer = sel[pair] ^ sel_xor
    const int* done;   // nullable: pairs whose level has converged are skipped
    double* partial;   // [pair][blocks_per_pair][2] sums of |du|, |dv| (SRC_WARP / SRC_WARPED)
    // SRC_FRAMES / SRC_GRADS outputs
    float* out_u;
    float* out_v;
    int H, W;
    int row_lo, row_hi, own_lo, own_hi;  // SRC_WARP row-band mode (see RefineArgs); else 0, H, 0, H
    // SRC_WARPED on the marching exact kernel only (lk_tile_fuses_tail): counter != nullptr -> the pair's last unit
    // finishes the iteration itself (convergence test, ping-pong flip, trace) and no iter_finalize launch follows
    IterTail tail;
};
// whether launch_lk_tile(src, window, a, ...) runs a kernel that honours a.tail (the marching exact kernel)
bool lk_tile_fuses_tail(int src, int window, const TileArgs& a);
cudaError_t launch_lk_tile(int src, int window, const TileArgs& a, int batch, int* launches, cudaStream_t stream);
// window 5, SRC_FRAMES / SRC_WARPED: second version of the kernel (lk_tile5.cu); called through launch_lk_tile
cudaError_t launch_lk_tile5(int src, const TileArgs& a, int batch, cudaStream_t stream);
// window 5, SRC_FRAMES / SRC_WARPED: the marching form of the exact kernel (lk_exact_march.cu; packed pairs, warp-private
// shared-memory rings); called through launch_lk_tile.  Same bits; its per-unit partial sums fill the same
// lk_tile_blocks_per_pair(rows, W) slots the tile kernels use.
cudaError_t launch_lk_exact_march(int src, const TileArgs& a, int batch, cudaStream_t stream);
int lk_tile_blocks_per_pair(int rows, int W);
bool lk_tile_window_supported(int window);

// after one refinement iteration: reduce the per-block partial sums, decide convergence,
// flip the ping-pong selector (lucas_kanade_pyramidal.py:213-223)
struct IterFinalizeArgs {
    const double* partial;
    int blocks_per_pair;
    int H, W;
    int* sel;
    int* done;
    int* iters_executed;        // nullable; element of pair b is iters_executed[b * iters_pair_stride]
    int iters_pair_stride;
    float* residuals;           // nullable; (mean|du|, mean|dv|) of pair b, iteration i at
    size_t resid_pair_stride;   //   residuals[b * resid_pair_stride + 2 * i]
    int iteration;
};
cudaError_t launch_iter_finalize(const IterFinalizeArgs& a, int batch, int* launches, cudaStream_t stream);
// row-band mode: just the per-pair sums (sum|du|, sum|dv| over this rank's rows) -> sums[pair][2];
// the ranks all-reduce them and take the decision on the host side of the C ABI
// the same decision as iter_finalize, from sums that were already reduced (over blocks and ranks)
cudaError_t launch_convergence_update(const double* sums, int batch, double n_pixels, int* sel, int* done,
                                      int* iters_executed, float* residuals, int max_iters, int iteration,
                                      int* launches, cudaStream_t stream);
cudaError_t launch_sum_partials(const double* partial, int blocks_per_pair, double* sums, int batch, int* launches,
                                cudaStream_t stream);

// ---- helper kernels (pyramid.cu) -------------------------------------------------------
cudaError_t launch_gradients(const float* prev, const float* curr, float* ix, float* iy, float* it, int batch, int H,
                             int W, int* launches, cudaStream_t stream);
// fused Gaussian (separable, float64 accumulate, float32 store per axis, reflect) + bilinear
// resample on the np.linspace grid (lucas_kanade_pyramidal.py:44-59)
// fast = true (fast mode of the pyramidal drivers only): fused multiply-adds in the float64 filter
// (OF_B200_PYRAMID_FAST=f32: the marching kernel's float32 flavour, an experiment -- faster, less faithful)
cudaError_t launch_pyramid_down(const float* src, float* dst, int batch, int H, int W, int oh, int ow,
                                const double* weights, int radius, int row_lo, int row_hi, int* launches,
                                cudaStream_t stream, bool fast = false);
// the same level by the marching kernel (pyramid_march.cu): radius 8, decimation step in [1, 6]
bool pyramid_march_supported(int H, int W, int oh, int ow, int radius);
// flavour: 0 = SciPy's bits (separate float64 multiplies and adds), 1 = float64 with fused multiply-adds,
// 2 = float32 with fused multiply-adds (experiment, OF_B200_PYRAMID_FAST=f32: no conversions, a quarter of the pipe time)
cudaError_t launch_pyramid_march(const float* src, float* dst, int batch, int H, int W, int oh, int ow,
                                 const double* weights, int row_lo, int row_hi, int flavour, int* launches,
                                 cudaStream_t stream);
cudaError_t launch_warp(const float* img, const float* fu, const float* fv, float* out, int batch, int H, int W,
                        int* launches, cudaStream_t stream);
// coarse flow (selected ping-pong buffer) -> fine grid, scaled (lucas_kanade_pyramidal.py:100-138)
cudaError_t launch_upsample_flow(const float* cu0, const float* cv0, const float* cu1, const float* cv1,
                                 const int* sel, int sel_xor, float* fu, float* fv, int batch, int ch, int cw, int th,
                                 int tw, int row_lo, int row_hi, int* launches, cudaStream_t stream);
// copy the selected ping-pong buffer of every pair to the caller's output
cudaError_t launch_select_copy(const float* u0, const float* v0, const float* u1, const float* v1, const int* sel,
                               int sel_xor, float* out_u, float* out_v, int batch, size_t n, int* launches,
                               cudaStream_t stream);

// ---- peer-memory collectives of the row-band mode (peer.cu) ----------------------------
struct PeerView {
    char* peer[PEER_MAX_WORLD];  // arena base of every rank as mapped in this process (peer[rank] = own)
    int world, rank;
    size_t flag_off, xchg_off;   // unsigned long long flags[PEER_SLOTS][PEER_MAX_WORLD]; double xchg[PEER_SLOTS][PEER_MAX_WORLD][2]
    int* err;                    // error word in the local arena (a wait timed out); sticky until the host has read it
    unsigned long long timeout_ns;
    // sequence number of collective `op` of a run = *run_id * ops_per_run + op (run_id: device word)
    const unsigned long long* run_id;
    unsigned long long ops_per_run;
};
void fill_peer_sync(PeerSync& s, const PeerView& pv, unsigned long long op);
// first kernel of a run: ++*run_id; clear_error: also reset the error word (the host has seen it)
cudaError_t launch_peer_begin_run(const PeerView& pv, bool clear_error, int* launches, cudaStream_t stream);
// rows of a plane (current ping-pong buffer: sel[0] ^ sel_xor ? src1 : src0) -> byte offset dst_off
// of every rank's arena; `first` / `count` in elements
cudaError_t launch_peer_push_rows(const PeerView& pv, const float* src0, const float* src1, const int* sel, int sel_xor,
                                  size_t dst_off, size_t first, size_t count, bool skip_self, int* launches,
                                  cudaStream_t stream);
// the same with one element range [lo[r], hi[r]) per receiving rank (clipped to [first, first + count)): a halo push
cudaError_t launch_peer_push_ranges(const PeerView& pv, const float* src0, const float* src1, const int* sel, int sel_xor,
                                    size_t dst_off, size_t first, size_t count, const size_t* lo, const size_t* hi,
                                    int* launches, cudaStream_t stream);
// signal collective `op` of the current run to every rank and wait for every rank's
cudaError_t launch_peer_sync(const PeerView& pv, unsigned long long op, int* launches, cudaStream_t stream);
// all-reduce of (sum|du|, sum|dv|) + the reference's convergence test + ping-pong flip, in one kernel
cudaError_t launch_peer_allreduce_update(const PeerView& pv, unsigned long long op, const double* partial, int blocks,
                                         double n_pixels, int* sel, int* done, int* iters_executed, float* residuals,
                                         int iteration, int* launches, cudaStream_t stream);

// ---- flow-field error metrics over a rectangular test region (metrics.cu) ---------------
int metrics_blocks_per_pair(int rows, int cols);
// partial: [batch][metrics_blocks_per_pair][6] doubles; out: [batch][5] = mae_u, mae_v, rmse, epe, aae
cudaError_t launch_flow_metrics(const float* u, const float* v, const float* u_true, const float* v_true, int batch, int H,
                                int W, int y0, int y1, int x0, int x1, double* partial, double* out, int* launches,
                                cudaStream_t stream);

// ---- fixture generators' sub-pixel shift (motion.cu) ------------------------------------
cudaError_t launch_apply_motion(const uint8_t* src, uint8_t* dst, const double* dx, const double* dy, int batch, int H,
                                int W, double cval, int* launches, cudaStream_t stream);

// cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT) in OpenCV's fixed point; minv: [batch][6] inverted matrices (host)
cudaError_t launch_warp_affine(const uint8_t* src, uint8_t* dst, const double* minv, int batch, int H, int W, int cval,
                               int* launches, cudaStream_t stream);

// ---- fixed-point mode (lk_fixed.cu) ----------------------------------------------------
cudaError_t launch_lk_fixed(const uint8_t* prev, const uint8_t* curr, int16_t* u, int16_t* v, int batch, int H, int W,
                            int mirror_avg_quirk, int* launches, cudaStream_t stream);

}  // namespace ofb
