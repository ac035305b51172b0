// metrics.cu (compute_all_metrics over the verifier's test region) on the CPU from its own source.
// TEST INFRASTRUCTURE.
#include "cuda_on_host.h"

#include "metrics.cu"

using namespace ofb;
extern "C" {
int emul_metrics_blocks_per_pair(int rows, int cols) { return metrics_blocks_per_pair(rows, cols); }
int emul_flow_metrics(const float* u, const float* v, const float* u_true, const float* v_true, int batch, int H, int W,
                      int y0, int y1, int x0, int x1, double* partial, double* out) {
    return (int)launch_flow_metrics(u, v, u_true, v_true, batch, H, W, y0, y1, x0, x1, partial, out, nullptr, nullptr);
}
}
