/*
 * of_b200.h -- C ABI of libof_b200.so, the B200 (sm_100a) backend for the Lucas-Kanade
 * hot path of rothej/optical-flow-fpga.
 *
 * The reference has no FFI: its hot path is a set of module-level Python functions
 * (python/lucas_kanade_core.py, python/lucas_kanade_pyramidal.py).  Each entry point
 * below replaces one of them and is what a ctypes binding of that function calls; the
 * drop-in Python modules in optical-flow-fpga_b200/ are exactly such bindings.
 * See INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - images are C-contiguous row-major [batch][height][width]; x is the column;
 *   - every function returns OF_OK (0) or an of_status error code, never throws;
 *     of_last_error() gives the message of the calling thread's last failure;
 *   - the caller owns every buffer; inputs are never written;
 *   - "host" entry points take host pointers and do H2D, kernels, D2H themselves;
 *     "_dev" entry points take device pointers and a cudaStream_t (as void*) and only
 *     enqueue work -- they do not synchronise;
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     OF_ERR_NO_DEVICE.
 */
#ifndef OF_B200_H
#define OF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum of_status {
    OF_OK = 0,
    OF_ERR_INVALID_ARGUMENT = 1,
    OF_ERR_CUDA = 2,
    OF_ERR_UNSUPPORTED = 3,
    OF_ERR_NO_DEVICE = 4,
    OF_ERR_OUT_OF_MEMORY = 5,
    OF_ERR_PEER_TIMEOUT = 6 /* row-band mode: a peer rank did not answer in time; the run's flow is invalid */
} of_status;

/* Arithmetic mode of the float path.
 * OF_MODE_EXACT  the reference's operation order (NumPy pairwise window sums, SciPy's
 *                float64 filters): every per-pixel operation is bit-identical to the Python reference on
 *                any float32 input.  One reduction is not: the pyramidal path's early-exit test compares
 *                mean|du|, mean|dv| of a level with 0.01, and those means are float64 sums rounded to
 *                float32 here, float32 pairwise sums (np.mean) in the reference -- they agree to ~1e-7
 *                relative, so an iteration count (and with it the flow) can differ only when a level's
 *                mean residual lies within that distance of the threshold.
 * OF_MODE_FAST   separable window sums in registers (the throughput kernels); single scale:
 *                bit-identical to the reference on uint8-valued frames (all partial sums exactly
 *                representable), tolerance-level otherwise.  Pyramidal: additionally fused
 *                multiply-adds in the float64 Gaussian (one float32 ulp on ~5 values in 10^9). */
#define OF_MODE_EXACT 0
#define OF_MODE_FAST 1

/* flags of the fixed-point mode */
#define OF_FX_MIRROR_AVG_QUIRK 1 /* 9-bit signed average of gradient_compute.sv:116 */

#define OF_MAX_WINDOW 11
#define OF_MAX_GAUSS_RADIUS 16

/* ---- library / device ---------------------------------------------------------------- */
int of_version(void);
const char* of_last_error(void);
int of_device_count(void);              /* >= 0; 0 when no CUDA device is usable        */
int of_set_device(int ordinal);
long long of_kernel_launches(void);     /* kernels launched by this library so far      */
int of_host_alloc_pinned(void** ptr, size_t bytes);
int of_host_free_pinned(void* ptr);
/* The host-buffer entry points keep grow-only device buffers and three streams PER DEVICE ORDINAL (the
 * device current at the call), so one process -- or several threads, each with its own current device --
 * can drive several GPUs.  of_release_host_buffers frees the current device's buffers (they are
 * re-allocated on demand); host calls on one device are serialised, calls on different devices are not. */
int of_release_host_buffers(void);

/* ---- host-buffer entry points ------------------------------------------------------- */

/* compute_gradients(frame_prev, frame_curr) -> (Ix, Iy, It)
 * replaces python/lucas_kanade_core.py:15-45 */
int of_gradients_f32(const float* prev, const float* curr, float* ix, float* iy, float* it,
                     int height, int width);

/* lucas_kanade_from_gradients(Ix, Iy, It, window_size) -> (u, v)
 * replaces python/lucas_kanade_core.py:73-135 */
int of_lk_from_gradients_f32(const float* ix, const float* iy, const float* it, float* u, float* v,
                             int height, int width, int window);

/* lucas_kanade_single_scale(frame_prev, frame_curr, window_size) -> (u, v), for a batch of
 * independent frame pairs.  replaces python/lucas_kanade_core.py:48-70 */
int of_lk_single_scale_f32(const float* prev, const float* curr, float* u, float* v, int batch,
                           int height, int width, int window, int mode);

/* One coarser level of build_gaussian_pyramid: gaussian_filter (taps `weights`, 2*radius+1
 * float64 values as scipy.ndimage builds them) then bilinear resampling on the
 * np.linspace(0, n-1, out_n) grid.  replaces python/lucas_kanade_pyramidal.py:44-59 */
int of_pyramid_down_f32(const float* src, float* dst, int height, int width, int out_height,
                        int out_width, const double* weights, int radius);

/* warp_image(image, flow_u, flow_v).  replaces python/lucas_kanade_pyramidal.py:66-97 */
int of_warp_f32(const float* image, const float* flow_u, const float* flow_v, float* out, int height,
                int width);

/* upsample_flow(flow_u, flow_v, target_shape).  replaces python/lucas_kanade_pyramidal.py:100-138 */
int of_upsample_flow_f32(const float* coarse_u, const float* coarse_v, float* u, float* v,
                         int coarse_height, int coarse_width, int target_height, int target_width);

/* lucas_kanade_pyramidal(frame_prev, frame_curr, num_levels, window_size, num_iterations)
 * for a batch of independent pairs; replaces python/lucas_kanade_pyramidal.py:141-228
 * (the numeric part; no prints, no plotting).
 *   iters_executed  optional [batch][levels] (level 0 = coarsest): iterations run before the
 *                   reference's early exit (mean|du| < 0.01 and mean|dv| < 0.01) fired
 *   residuals       optional [batch][levels][iterations][2]: mean|du|, mean|dv| */
int of_lk_pyramidal_f32(const float* prev, const float* curr, float* u, float* v, int batch,
                        int height, int width, int levels, int window, int iterations, int mode,
                        const double* gauss_weights, int gauss_radius, int* iters_executed,
                        float* residuals);

/* Fixed-point single-scale LK mirroring the reference RTL's integer datapath
 * (rtl/unopt/gradient_compute.sv, window_accumulator.sv, flow_solver.sv):
 * uint8 frames in, int16 S8.7 flow out (real = value / 128), 5x5 window. */
int of_lk_single_scale_fx(const uint8_t* prev, const uint8_t* curr, int16_t* u, int16_t* v,
                          int batch, int height, int width, int flags);

/* ---- device-buffer entry points (asynchronous on `stream`) --------------------------- */
int of_lk_single_scale_f32_dev(const float* prev, const float* curr, float* u, float* v, int batch,
                               int height, int width, int window, int mode, void* stream);

size_t of_lk_pyramidal_workspace_bytes(int batch, int height, int width, int levels, int iterations);

int of_lk_pyramidal_f32_dev(const float* prev, const float* curr, float* u, float* v, int batch,
                            int height, int width, int levels, int window, int iterations, int mode,
                            const double* gauss_weights, int gauss_radius, void* workspace,
                            size_t workspace_bytes, int* iters_executed_dev, float* residuals_dev,
                            void* stream);

/* ---- building blocks of the pyramidal path on device buffers (row-band multi-GPU mode) ---
 * A rank of a row-band job holds full-size frames / flow planes but computes only its rows.
 * All four calls only enqueue work on `stream`. */

/* one coarser pyramid level for a batch (device version of of_pyramid_down_f32).  mode:
 * OF_MODE_EXACT = SciPy's bits; OF_MODE_FAST = what the fast pyramidal drivers use (fused
 * multiply-adds in the float64 filter: about 5 values in 10^9 differ by one float32 ulp). */
int of_pyramid_down_f32_dev(const float* src, float* dst, int batch, int height, int width,
                            int out_height, int out_width, const double* weights, int radius,
                            int row_lo, int row_hi /* output rows to produce */, int mode, void* stream);

/* upsample_flow for target rows [row_lo, row_hi) only */
int of_upsample_flow_f32_dev(const float* coarse_u, const float* coarse_v, float* u, float* v,
                             int batch, int coarse_height, int coarse_width, int target_height,
                             int target_width, int row_lo, int row_hi, void* stream);

size_t of_lk_refine_workspace_bytes(int batch, int height, int width);

/* One refinement iteration of lucas_kanade_pyramidal (lucas_kanade_pyramidal.py:203-214) on rows
 * [row_lo, row_hi) of a level:  flow_out = flow_in + LK(prev, warp(curr, flow_in)).
 * flow_in must be valid on rows [row_lo - 3, row_hi + 3) (clipped to the frame); rows outside
 * [row_lo, row_hi) of flow_out are left untouched.  sums[pair][2] receives sum|du|, sum|dv|
 * over rows [own_lo, own_hi) (float64) -- the caller reduces them over ranks and applies the
 * reference's convergence test.  row_lo must be even. */
int of_lk_refine_f32_dev(const float* prev, const float* curr, const float* flow_in_u,
                         const float* flow_in_v, float* flow_out_u, float* flow_out_v, int batch,
                         int height, int width, int window, int mode, int row_lo, int row_hi,
                         int own_lo, int own_hi, double* sums, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Same iteration with device-side control, so that a rank never waits for the host inside a
 * level: the flow lives in two ping-pong buffer pairs, sel[pair] (0 / 1) says which one is
 * current, pairs with done[pair] != 0 are skipped.  After the ranks have all-reduced `sums`,
 * of_lk_convergence_update_dev applies the reference's test (mean|du| < 0.01 and mean|dv| <
 * 0.01 over n_pixels), flips sel, sets done and records the trace (optional arrays:
 * iters_executed[batch], residuals[batch][max_iterations][2]). */
int of_lk_refine_pingpong_f32_dev(const float* prev, const float* curr, float* flow0_u, float* flow0_v,
                                  float* flow1_u, float* flow1_v, const int* sel, const int* done,
                                  int batch, int height, int width, int window, int mode, int row_lo,
                                  int row_hi, int own_lo, int own_hi, double* sums, void* workspace,
                                  size_t workspace_bytes, void* stream);
int of_lk_convergence_update_dev(const double* sums, int batch, double n_pixels, int* sel, int* done,
                                 int* iters_executed, float* residuals, int max_iterations,
                                 int iteration, void* stream);

int of_lk_single_scale_fx_dev(const uint8_t* prev, const uint8_t* curr, int16_t* u, int16_t* v,
                              int batch, int height, int width, int flags, void* stream);

/* ---- uint8 ingest ---------------------------------------------------------------------------
 * The reference's frames are uint8 on disk (frame_00.bin raw bytes, frame_00.mem one hex byte per
 * line: python/generate_test_suite.py:259-271) and are widened to float32 by the callers
 * (python/lucas_kanade_reference.py:133-140, optical_flow_verifier.py:61-65) before
 * lucas_kanade_single_scale.  These entry points take the bytes directly: same float32 result, bit
 * for bit, with 10 instead of 16 bytes of HBM traffic per pixel in fast mode. */
int of_lk_single_scale_u8(const uint8_t* prev, const uint8_t* curr, float* u, float* v, int batch, int height,
                          int width, int window, int mode);
/* device buffers; fast mode only; needs window 5 or 7, width % 16 == 0, 16-byte aligned planes */
int of_lk_single_scale_u8_dev(const uint8_t* prev, const uint8_t* curr, float* u, float* v, int batch, int height,
                              int width, int window, void* stream);
/* read one [height][width] uint8 frame from a .bin (raw) or .mem (hex lines) file into host memory */
int of_load_frame_u8(const char* path, uint8_t* out, int height, int width);

/* ---- flow-field text export (host-side I/O) --------------------------------------------------
 * export_flow_field_txt of python/lucas_kanade_reference.py:78-103: header lines, then "x y u v" per
 * pixel (row-major, six decimals) -- the format scripts/visualize_flow.py and the RTL testbench
 * (tb/tb_optical_flow_top.sv:340-358) share.  x_min < 0: no "# Test region" line.  The _fx variant
 * takes the fixed-point mode's S8.7 flow and writes the testbench's header. */
int of_export_flow_txt(const char* path, const float* u, const float* v, int height, int width, int x_min, int x_max,
                       int y_min, int y_max);
int of_export_flow_fx_txt(const char* path, const int16_t* u, const int16_t* v, int height, int width, int x_min,
                          int x_max, int y_min, int y_max);

/* ---- apply_motion of the fixture generators -------------------------------------------------
 * apply_motion(frame, dx, dy) of python/generate_test_frames_natural.py:67-73, i.e.
 * scipy.ndimage.shift(frame, (dy, dx), order=1, mode="constant", cval=128) on uint8 frames, for a batch
 * with one (dx, dy) per frame; bit-identical to SciPy.  Produces the second frame of a synthetic pair
 * on the device.  dx / dy: [batch] float64 (host pointers for the host call, device pointers for _dev). */
int of_apply_motion_u8(const uint8_t* frames, uint8_t* out, int batch, int height, int width, const double* dx,
                       const double* dy, double cval);
int of_apply_motion_u8_dev(const uint8_t* frames, uint8_t* out, int batch, int height, int width, const double* dx,
                           const double* dy, double cval, void* stream);

/* apply_motion_opencv(frame, params) of python/generate_test_suite.py:165-204, i.e.
 * cv2.warpAffine(frame, M, (W, H), flags=INTER_LINEAR, borderMode=BORDER_CONSTANT, borderValue=cval) on uint8
 * frames, bit-identical to OpenCV's fixed-point bilinear warp (10-bit coordinates, 1/32-pixel weights,
 * 15-bit coefficients).  matrices: [batch][6] forward 2x3 matrices, row-major, HOST memory in both
 * flavours (they are inverted on the host in OpenCV's operation order and travel as kernel arguments). */
int of_warp_affine_u8(const uint8_t* frames, uint8_t* out, int batch, int height, int width, const double* matrices,
                      int cval);
int of_warp_affine_u8_dev(const uint8_t* frames, uint8_t* out, int batch, int height, int width,
                          const double* matrices, int cval, void* stream);

/* ---- flow-field error metrics on the device ----------------------------------------------
 * compute_all_metrics(u_pred, v_pred, u_true, v_true, mask) of python/flow_metrics.py:166-201 for a
 * batch of flow fields, the mask being the verifier's rectangular test region
 * (python/optical_flow_verifier.py:96-138): rows [y0, y1), columns [x0, x1).  u_true / v_true hold one
 * constant ground-truth flow per pair.  metrics[pair] = {mae_u, mae_v, rmse, epe, aae (degrees)}.
 * Per-pixel arithmetic is the reference's float32; the means are float64 sums (the reference's are
 * float32 pairwise sums), so values agree to float32 rounding of a mean, not bit for bit. */
size_t of_flow_metrics_workspace_bytes(int batch, int region_height, int region_width);
int of_flow_metrics_f32_dev(const float* u, const float* v, const float* u_true, const float* v_true, int batch,
                            int height, int width, int y0, int y1, int x0, int x1, double* metrics,
                            void* workspace, size_t workspace_bytes, void* stream);
int of_flow_metrics_f32(const float* u, const float* v, const float* u_true, const float* v_true, int batch, int height,
                        int width, int y0, int y1, int x0, int x1, double* metrics);

/* ---- row-band multi-GPU mode of lucas_kanade_pyramidal, native driver ---------------------
 * One frame pair, the rows of every pyramid level split over `world` ranks (one process -- or
 * thread -- per GPU of one NVLink / NVSwitch domain).  The reference has no multi-device path;
 * this is lucas_kanade_pyramidal (python/lucas_kanade_pyramidal.py:141-228) for frames too large
 * or too urgent for one GPU, with results bit-identical to of_lk_pyramidal_f32_dev.
 *
 * Every rank creates a context (which allocates its "arena" with one cudaMalloc), publishes the
 * arena to the other ranks -- of_rowband_ipc_handle / of_rowband_open_peers_ipc between
 * processes (the 64-byte handles travel over any channel, e.g. torch.distributed
 * all_gather_object), of_rowband_set_peers between threads of one process -- and then calls
 * of_rowband_run on a stream of its device: the call only enqueues kernels; pyramid rows, final
 * flow rows and the per-iteration residual sums move between the GPUs by peer stores and flag
 * words inside those kernels (no NCCL, no host round trip).  All ranks must issue the same
 * sequence of of_rowband_run calls. */
#define OF_IPC_HANDLE_BYTES 64
typedef struct of_rowband of_rowband_t;

int of_rowband_create(of_rowband_t** ctx, int rank, int world, int height, int width, int levels, int window,
                      int iterations, int mode, const double* gauss_weights, int gauss_radius);
size_t of_rowband_arena_bytes(const of_rowband_t* ctx);
void* of_rowband_arena(const of_rowband_t* ctx); /* device pointer of this rank's arena */
int of_rowband_ipc_handle(const of_rowband_t* ctx, void* handle_out /* OF_IPC_HANDLE_BYTES */);
int of_rowband_open_peers_ipc(of_rowband_t* ctx, const void* handles /* world x OF_IPC_HANDLE_BYTES, rank order */);
int of_rowband_set_peers(of_rowband_t* ctx, void* const* arenas /* world device pointers, rank order */);
/* Pyramid levels with at most `pixels` pixels are computed whole on every rank instead of in row
 * bands (their kernels are launch-latency bound either way; replication removes their collectives).
 * Default 600000; 0 = split every level.  Must be the same on every rank.  Same bits either way. */
int of_rowband_set_replicate_pixels(of_rowband_t* ctx, long long pixels);
/* How long a kernel spins on a peer's flag before it gives up, sets the error word (of_rowband_trace)
 * and lets the run drain: default 4000 ms.  The ranks must enter of_rowband_run within this time of
 * each other; raise it when one rank may be late (I/O), lower it to fail fast. */
int of_rowband_set_timeout_ms(of_rowband_t* ctx, int milliseconds);
/* prev, curr: the full [height][width] frames on this rank's device.  u, v: optional full-size
 * outputs (device); with NULL the result stays in the arena (of_rowband_result). */
int of_rowband_run(of_rowband_t* ctx, const float* prev, const float* curr, float* u, float* v, void* stream);
int of_rowband_result(const of_rowband_t* ctx, const float** u, const float** v);
/* waits for `stream`, then copies out (each optional) iters_executed[levels] (level 0 = coarsest),
 * residuals[levels][iterations][2] and the error word (non-zero: a peer did not answer in time).
 * Returns OF_ERR_PEER_TIMEOUT when the error word is set: the flow of that run is garbage.  The word stays
 * set until the host has read it here; the next of_rowband_run after that starts clean (a time-out can be
 * as harmless as ranks entering a cold first run too far apart). */
int of_rowband_trace(of_rowband_t* ctx, int* iters_executed, float* residuals, int* error, void* stream);
/* the same check without the trace: OF_OK, or OF_ERR_PEER_TIMEOUT (call it before trusting of_rowband_result) */
int of_rowband_status(of_rowband_t* ctx, void* stream);
int of_rowband_destroy(of_rowband_t* ctx);

#ifdef __cplusplus
}
#endif
#endif /* OF_B200_H */
