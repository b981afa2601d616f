/*
 * Sequential event -> voxel-grid accumulation: the bit-exact specification of
 * the deterministic binning mode.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py) -- never linked into the product.
 *
 * Restates, one event at a time, what the reference's two voxelisers do through
 * library scatter-adds (paths relative to the reference checkout):
 *
 *   flavour 0  utils/event_process.py:127-190  events_to_voxel_grid_pytorch
 *              fp64 time normalisation, floor, fp32 weights
 *              w_lo = p * (1.0f - (float)dt), w_hi = p * (float)dt,
 *              Tensor.index_add_ on CPU == sequential fp32 adds in event order.
 *   flavour 1  utils/event_process.py:15-72    events_to_voxel_grid
 *              fp64 weights, np.add.at into a float32 array == each add is
 *              done in fp64 and rounded to fp32, in event order.
 *   flavour 2  utils/event_process.py:75-123   events_to_voxel_grid_pol
 *              like flavour 1 but grid is [nb, 2, H, W], channel = polarity,
 *              weights positive.
 *
 * In every flavour ALL "left" contributions (bin ti) are added before ANY
 * "right" contribution (bin ti+1): the reference issues two scatter-adds.
 *
 * Build:  make -C oracle      (gcc -O2, no -ffast-math: IEEE semantics matter)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

int cf_oracle_voxel_seq(const double *ev, int64_t n, int nb, int height, int width,
                        int flavour, float *grid)
{
    const int64_t plane = (int64_t)height * width;
    const int64_t cells = plane * nb * (flavour == 2 ? 2 : 1);
    memset(grid, 0, sizeof(float) * (size_t)cells);
    if (n <= 0)
        return 0;
    if (flavour < 0 || flavour > 2)
        return -1;

    const double t0 = ev[0];
    double span = ev[(n - 1) * 4] - t0;
    if (span == 0.0)
        span = 1.0;
    const double scale = (double)(nb - 1);

    for (int pass = 0; pass < 2; ++pass) {
        for (int64_t i = 0; i < n; ++i) {
            const double *e = ev + i * 4;
            /* two roundings, in this order: multiply, then divide */
            volatile double num = scale * (e[0] - t0);
            const double tn = num / span;
            const double lo = floor(tn);
            if (!(lo >= 0.0))
                continue; /* unsorted / NaN input: dropped (torch path checks tis >= 0) */
            const double bin = lo + (double)pass;
            if (!(bin < (double)nb))
                continue;
            const double dt = tn - lo;
            const int64_t x = (int64_t)e[1];
            const int64_t y = (int64_t)e[2];
            const double p = e[3];
            int64_t cell;
            if (flavour == 2)
                cell = x + y * width + (int64_t)p * plane + (int64_t)bin * plane * 2;
            else
                cell = x + y * width + (int64_t)bin * plane;
            if (cell < 0 || cell >= cells)
                return -2; /* the reference would fault / wrap; callers pre-filter */
            if (flavour == 0) {
                const float s = (p == 0.0) ? -1.0f : (float)p;
                const float f = (float)dt;
                const float w = pass == 0 ? s * (1.0f - f) : s * f;
                volatile float acc = grid[cell] + w;
                grid[cell] = acc;
            } else {
                const double s = (p == 0.0) ? (flavour == 1 ? -1.0 : 1.0) : p;
                const double w = pass == 0 ? s * (1.0 - dt) : s * dt;
                grid[cell] = (float)((double)grid[cell] + w);
            }
        }
    }
    return 0;
}
