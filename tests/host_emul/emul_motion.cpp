// motion.cu (apply_motion = scipy.ndimage.shift, warp_affine = cv2.warpAffine's fixed-point warp) on the CPU from
// its own source, through its launchers.  TEST INFRASTRUCTURE.
#include "cuda_on_host.h"
static inline long long __double2ll_rn(double x) { return std::llrint(x); }  // round to nearest even (default mode)

#include "motion.cu"

using namespace ofb;
extern "C" {
int emul_apply_motion(const uint8_t* src, uint8_t* dst, const double* dx, const double* dy, int batch, int H, int W, double cval) {
    return (int)launch_apply_motion(src, dst, dx, dy, batch, H, W, cval, nullptr, nullptr);
}
int emul_warp_affine(const uint8_t* src, uint8_t* dst, const double* minv, int batch, int H, int W, int cval) {
    return (int)launch_warp_affine(src, dst, minv, batch, H, W, cval, nullptr, nullptr);
}
}
