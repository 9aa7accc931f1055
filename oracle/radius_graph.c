/* CPU oracle for the canonical neighbour list.  TEST INFRASTRUCTURE ONLY (never linked
 * into the product library).
 *
 * Restates the call `radius_graph(pos, r=5.0, batch, loop=True, max_num_neighbors=32)`
 * made at /root/reference/src/ViSNet/model/utils.py:259-266.  The implementation of that
 * call lives in torch_cluster (un-vendored, version unpinned, absent from this image), so
 * this follows its *recalled* CUDA brute-force semantics -- parity unpinned for this
 * function: for each target i, scan sources j of the same graph in ascending index,
 * accumulate d2 += diff*diff over x,y,z (nvcc contracts this to an FMA chain), accept on
 * strict d2 < r*r, write into slot [i*K + count], stop after K hits.  The self pair is
 * accepted like any other (loop=True).
 *
 * Compile with -ffp-contract=off so that the only fused operations are the explicit fmaf().
 */
#include <math.h>
#include <stdint.h>

int64_t radius_graph_ref(const float *pos, const int64_t *batch, int64_t n, float cutoff,
                         int32_t max_nbr, int32_t *slots /* [n*max_nbr], -1 = empty */,
                         int32_t *deg /* [n] */)
{
    const float r2 = cutoff * cutoff;
    int64_t total = 0;
    int64_t g0 = 0;
    while (g0 < n) {
        int64_t g1 = g0;
        while (g1 < n && batch[g1] == batch[g0]) g1++;
        for (int64_t i = g0; i < g1; i++) {
            int32_t cnt = 0;
            const float xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
            for (int64_t j = g0; j < g1 && cnt < max_nbr; j++) {
                const float dx = pos[3 * j] - xi;
                const float dy = pos[3 * j + 1] - yi;
                const float dz = pos[3 * j + 2] - zi;
                float d2 = dx * dx;
                d2 = fmaf(dy, dy, d2);
                d2 = fmaf(dz, dz, d2);
                if (d2 < r2) slots[i * max_nbr + cnt++] = (int32_t)j;
            }
            for (int32_t k = cnt; k < max_nbr; k++) slots[i * max_nbr + k] = -1;
            deg[i] = cnt;
            total += cnt;
        }
        g0 = g1;
    }
    return total;
}
