// lk_fixed.cu (the RTL's integer datapath, tile kernel) on the CPU from its own source.  TEST INFRASTRUCTURE.
#include "cuda_on_host.h"

#include "lk_fixed.cu"

using namespace ofb;
extern "C" int emul_lk_fixed(const uint8_t* prev, const uint8_t* curr, int16_t* u, int16_t* v, int batch, int H, int W,
                             int mirror_avg_quirk) {
    return (int)launch_lk_fixed(prev, curr, u, v, batch, H, W, mirror_avg_quirk, nullptr, nullptr);
}
