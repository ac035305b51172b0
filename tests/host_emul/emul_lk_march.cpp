// lk_march.cu -- the fast marching kernels (TMA ring, packed pairs, warp shuffles) -- run on the CPU from its own
// source through its own launchers.  Stand-ins for what only the device has:
//   * TMA: a tensor map is a small host record (base, extents, element size, box); cp.async.bulk.tensor is a
//     synchronous box copy with zero fill outside the tensor, followed by the barrier's complete_tx;
//   * mbarrier: a byte counter and a phase bit in the barrier's 8 bytes;
//   * packed f32x2 arithmetic: two IEEE float32 operations; rcp.approx: an exact reciprocal (the quotient
//     sequence built on it must give the IEEE quotient for ANY approximation within its error bound).
// TEST INFRASTRUCTURE; tests/test_kernel_host_emulation.py builds and drives it.
#include "cuda_on_host.h"

#include <cuda.h>  // CUtensorMap (128 opaque bytes)

namespace ofb {

// ---- tensor maps -------------------------------------------------------------------------------------------
struct HostTensorMap {
    const unsigned char* base;
    long long W, H, B;  // extents in elements
    int elem;           // bytes per element
    int box_w, box_h;   // box in elements
};
static_assert(sizeof(HostTensorMap) <= sizeof(CUtensorMap), "record must fit the opaque map");
typedef int (*EncodeTiledFn)();
static EncodeTiledFn get_encode_fn() { return [] { return 0; }; }  // "the driver has a tensor-map encoder"
static bool make_map(CUtensorMap* map, const void* base, int batch, int H, int W, int elem, int box_w, int box_h) {
    HostTensorMap m{static_cast<const unsigned char*>(base), W, H, batch, elem, box_w, box_h};
    std::memset(map, 0, sizeof(*map));
    std::memcpy(map, &m, sizeof(m));
    // what cuTensorMapEncodeTiled insists on: 16-byte aligned base and row pitch
    return (reinterpret_cast<uintptr_t>(base) & 15) == 0 && ((long long)W * elem) % 16 == 0;
}
constexpr int HOST_LOADW = 128, HOST_U8_BOX_W = 256;  // = LOADW, U8_BOX_W of lk_march.cu (checked below)
static bool make_frame_map(CUtensorMap* map, const float* base, int batch, int H, int W, int box_rows) {
    return make_map(map, base, batch, H, W, 4, HOST_LOADW, box_rows);
}
static bool make_frame_map_u8(CUtensorMap* map, const uint8_t* base, int batch, int H, int W, int box_rows) {
    return make_map(map, base, batch, H, W, 1, HOST_U8_BOX_W, box_rows);
}

// ---- shared-memory addresses, mbarriers, TMA -----------------------------------------------------------------
static inline uint32_t smem_u32(const void* p) {
    return (uint32_t)(static_cast<const char*>(p) - static_cast<const char*>(cuda_on_host::dynamic_smem())) + 128u;
}
static inline char* smem_ptr(uint32_t a) { return static_cast<char*>(cuda_on_host::dynamic_smem()) + (a - 128u); }
struct HostBarrier {
    int tx;                  // bytes still expected in the current phase
    unsigned short pending;  // arrivals still expected in the current phase
    unsigned char count;     // arrivals per phase
    unsigned char phase;     // parity of the current (incomplete) phase
};
static_assert(sizeof(HostBarrier) == 8, "an mbarrier is 8 bytes");
static pthread_mutex_t g_mbar_lock = PTHREAD_MUTEX_INITIALIZER;  // one lock for every barrier: speed is no concern here
static inline void mbar_complete_if_done(HostBarrier* b) {       // lock held
    if (b->pending == 0 && b->tx == 0) {
        b->pending = b->count;
        __atomic_store_n(&b->phase, (unsigned char)(b->phase ^ 1u), __ATOMIC_RELEASE);
    }
}
static inline void mbar_init(uint32_t bar, uint32_t count) {
    HostBarrier* b = reinterpret_cast<HostBarrier*>(smem_ptr(bar));
    b->tx = 0;
    b->pending = (unsigned short)count;
    b->count = (unsigned char)count;
    b->phase = 0;
}
// mbarrier.arrive.expect_tx: one arrival that also announces `bytes`
static inline void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    HostBarrier* b = reinterpret_cast<HostBarrier*>(smem_ptr(bar));
    pthread_mutex_lock(&g_mbar_lock);
    b->tx += (int)bytes;
    b->pending -= 1;
    mbar_complete_if_done(b);
    pthread_mutex_unlock(&g_mbar_lock);
}
static inline void mbar_arrive(uint32_t bar) {
    HostBarrier* b = reinterpret_cast<HostBarrier*>(smem_ptr(bar));
    pthread_mutex_lock(&g_mbar_lock);
    b->pending -= 1;
    mbar_complete_if_done(b);
    pthread_mutex_unlock(&g_mbar_lock);
}
static inline bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    const bool done = __atomic_load_n(&reinterpret_cast<HostBarrier*>(smem_ptr(bar))->phase, __ATOMIC_ACQUIRE) != parity;
    if (!done) sched_yield();
    return done;
}
template <int REGS>
static inline void setmaxnreg_inc() {}
template <int REGS>
static inline void setmaxnreg_dec() {}
static inline void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
    HostTensorMap m;
    std::memcpy(&m, map, sizeof(m));
    char* out = smem_ptr(dst);
    for (int r = 0; r < m.box_h; ++r)
        for (int c = 0; c < m.box_w; ++c) {
            const long long yy = (long long)y + r, xx = (long long)x + c;
            const bool in = z >= 0 && z < m.B && yy >= 0 && yy < m.H && xx >= 0 && xx < m.W;
            char* o = out + ((size_t)r * m.box_w + c) * m.elem;
            if (in)
                std::memcpy(o, m.base + (((size_t)z * m.H + yy) * m.W + xx) * m.elem, m.elem);
            else
                std::memset(o, 0, m.elem);  // CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE: zeros
        }
    HostBarrier* b = reinterpret_cast<HostBarrier*>(smem_ptr(bar));
    const int bytes = m.box_h * m.box_w * m.elem;
    pthread_mutex_lock(&g_mbar_lock);
    b->tx -= bytes;  // complete_tx; the phase's last byte (and arrival) completes it
    mbar_complete_if_done(b);
    pthread_mutex_unlock(&g_mbar_lock);
}
// cp.async (LDGSTS): a synchronous 4-byte copy; groups complete at once
static inline void cp_async4(void* smem_dst, const void* gsrc) { std::memcpy(smem_dst, gsrc, 4); }
static inline void cp_async_commit() {}
template <int PENDING>
static inline void cp_async_wait() {}
#define OF_FENCE_MBARRIER_INIT() __atomic_thread_fence(__ATOMIC_SEQ_CST)
#define OF_KEEP_IN_REGISTER_F(x) (void)(x)
#define OF_KEEP_ALIVE_L(x) (void)(x)
#define OF_PREFETCH_L2(p) (void)(p)

// ---- packed pairs: host stand-ins in csrc/f32x2.cuh (two IEEE float32 operations each) ------------------------
static inline float rcp_approx(float x) { volatile float r = 1.0f / x; return r; }
static inline double rcp_approx_f64(double x) { volatile double r = 1.0 / x; return r; }

}  // namespace ofb

#include "peer_on_host.h"  // peer_device.cuh's system-scope flag accesses

#include "lk_march.cu"

using namespace ofb;
static_assert(HOST_LOADW == LOADW && HOST_U8_BOX_W == U8_BOX_W, "box widths of the stand-in tensor maps");

extern "C" {
// force_path: 0 = TMA when the pointers allow it, 1 = TMA or error, 2 = plain global loads
int emul_lk_march(const float* prev, const float* curr, float* u, float* v, int batch, int H, int W, int force_path,
                  int window) {
    if (!lk_march_supported(H, W, window)) return -1;
    return (int)launch_lk_march(prev, curr, u, v, batch, H, W, window, force_path, nullptr, nullptr);
}
int emul_lk_march_u8(const uint8_t* prev, const uint8_t* curr, float* u, float* v, int batch, int H, int W, int window) {
    return (int)launch_lk_march_u8(prev, curr, u, v, batch, H, W, window, nullptr, nullptr);
}
// One fast-mode refinement iteration.  form 0: split (warp_rows_kernel<float> + lk_march_kernel<REFINE>), the
// per-unit sums go to `partial`; form 1: split with the iteration's tail fused into the marching kernel (ticket
// counter, convergence test, ping-pong flip); form 2: the fused lk_refine_kernel (gathers inside the marching warps).
int emul_lk_refine(int form, const float* prev, const float* curr, float* flow_u0, float* flow_v0, float* flow_u1,
                   float* flow_v1, int* sel, int sel_xor, int* done, double* partial, float* warped, unsigned* counter,
                   int* iters_executed, float* residuals, int max_iters, int iteration, int batch, int H, int W, int row_lo,
                   int row_hi, int own_lo, int own_hi, int window, float* warped_next, int warped_ready) {
    RefineArgs ra;
    std::memset(&ra, 0, sizeof(ra));
    ra.prev = prev;
    ra.curr = curr;
    ra.flow_u[0] = flow_u0;
    ra.flow_v[0] = flow_v0;
    ra.flow_u[1] = flow_u1;
    ra.flow_v[1] = flow_v1;
    ra.sel = sel;
    ra.sel_xor = sel_xor;
    ra.done = done;
    ra.partial = partial;
    ra.H = H;
    ra.W = W;
    ra.window = window;
    ra.warped_next = warped_next;  // split forms: the marching kernel's epilogue warps the next iteration's input
    ra.warped_ready = warped_ready;
    ra.row_lo = row_lo;
    ra.row_hi = row_hi;
    ra.own_lo = own_lo;
    ra.own_hi = own_hi;
    if (!lk_refine_supported(ra, window)) return -1;
    if (form == 1 || form == 4) {
        ra.tail.counter = counter;
        ra.tail.peers = 0;
        ra.tail.n_pixels = (double)H * (double)W;
        ra.tail.sel = sel;
        ra.tail.done = done;
        ra.tail.iters_executed = iters_executed;
        ra.tail.iters_pair_stride = 1;
        ra.tail.residuals = residuals;
        ra.tail.resid_pair_stride = (size_t)max_iters * 2;
        ra.tail.iteration = iteration;
    }
    if (form == 2) return (int)launch_lk_refine(ra, batch, nullptr, nullptr);
    if (form == 3 || form == 4) return (int)launch_lk_refine_ws(ra, batch, nullptr, nullptr);  // 4: with the fused tail
    return (int)launch_lk_refine_split(ra, warped, batch, nullptr, nullptr);
}
int emul_lk_refine_units_per_pair(int batch, int rows, int W) { return lk_refine_units_per_pair(batch, rows, W); }

int emul_lk_march_fx(const uint8_t* prev, const uint8_t* curr, int16_t* u, int16_t* v, int batch, int H, int W, int quirk) {
    return (int)launch_lk_march_fx(prev, curr, u, v, batch, H, W, quirk, nullptr, nullptr);
}
}
