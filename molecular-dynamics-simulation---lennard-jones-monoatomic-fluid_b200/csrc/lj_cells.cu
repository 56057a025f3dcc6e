// lj_cells.cu -- counting-sort cell list + 27-cell stencil force kernel (placeholder: planned).
#include <cmath>
#include "ljmd_internal.h"

namespace ljmd {

int cells_per_edge_host(const ljmd_params &p)
{
    int nc = (int)floor(p.box_length / p.rc);
    if (nc > 0 && (double)nc * p.rc * (1.0 + 1e-12) > p.box_length) nc -= 1;
    if (nc > 1024) nc = 1024;
    return nc;
}

int cellwork_plan(ljmd_ctx *) { return set_error(LJMD_ERR_UNSUPPORTED, "cell list not built yet"); }
void cellwork_free(ljmd_ctx *) {}
int launch_cells_force(ljmd_ctx *, int) { return set_error(LJMD_ERR_UNSUPPORTED, "cell list not built yet"); }

} // namespace ljmd

extern "C" int ljmd_ctx_download_cells(ljmd_ctx *c, int32_t *, int32_t *, int32_t *)
{
    if (!c) return ljmd::set_error(LJMD_ERR_NULL, "ljmd_ctx_download_cells(): ctx is NULL.");
    return ljmd::set_error(LJMD_ERR_UNSUPPORTED, "cell list not built yet");
}
