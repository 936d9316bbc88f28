// Host entry points of the procgen height source; the algorithm lives in noise_impl.h (shared with the device kernels).
#include "uvt_host.h"

#include "noise_impl.h"

namespace {
const float *grad_table() {
    static float table[256];
    static const bool built = (uvt_noise::build_grad_table(table), true);
    (void)built;
    return table;
}
}  // namespace

extern "C" float uvt_noise2_fbm(float x, float y) { return uvt_noise::noise2_fbm(grad_table(), x, y); }

extern "C" void uvt_noise_grad_table(float out[256]) { uvt_noise::build_grad_table(out); }

extern "C" uint32_t uvt_procgen_height(uint32_t dim, uint32_t x, uint32_t z, float offset_x, float offset_y) {
    return uvt_noise::column_height(grad_table(), dim, x, z, offset_x, offset_y);
}
