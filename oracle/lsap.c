/* Oracle: rectangular linear-sum-assignment on the CPU.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference's Hungarian step is `scipy.optimize.linear_sum_assignment(cost)`
 * (thirdparty/mmdetection/mmdet/core/bbox/assigners/hungarian_assigner.py:136,
 * detr_ssod/models/dino_detr_ssod.py:279).  scipy is a third-party dependency that is
 * NOT vendored under /root/reference and is unpinned there
 * (thirdparty/mmdetection/requirements/optional.txt:4); the authoring container has
 * scipy 1.18.1.  This file restates the published algorithm scipy implements
 * (D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE TAES 2016:
 * shortest augmenting paths with float64 duals, tall matrices solved on the transpose,
 * unassigned-column preference on exact ties, candidate list filled in reverse order)
 * so that the GPU kernel has a step-for-step CPU twin to be compared with.
 *
 * Pinning: tests/test_oracle_lsap.py runs this against scipy itself on seeded random
 * float32 matrices, integer matrices full of exact ties, and the tie KATs of SURVEY.md
 * appendix B; tests/golden/lsap_golden.npz holds scipy's answers for the committed cases.
 *
 * Returns 0 ok, 1 = invalid entry (NaN / -inf), 2 = infeasible.
 * cost is row-major (nr x nc) float32 (what the reference hands scipy); arithmetic is
 * float64 like scipy's.  rows/cols receive min(nr,nc) pairs, rows ascending.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int64_t find_path(int64_t nc, const double *c, const double *u, const double *v,
                         int64_t *path, const int64_t *row4col, double *dist, int64_t i,
                         unsigned char *in_sr, unsigned char *in_sc, int64_t *cand,
                         double *out_min)
{
    double min_val = 0.0;
    int64_t left = nc;
    for (int64_t t = 0; t < nc; ++t) cand[t] = nc - t - 1;    /* reverse fill */
    for (int64_t j = 0; j < nc; ++j) dist[j] = INFINITY;
    int64_t sink = -1;
    while (sink < 0) {
        int64_t pick = -1;
        double low = INFINITY;
        in_sr[i] = 1;
        for (int64_t t = 0; t < left; ++t) {
            int64_t j = cand[t];
            double r = min_val + c[i * nc + j] - u[i] - v[j];
            if (r < dist[j]) { path[j] = i; dist[j] = r; }
            if (dist[j] < low || (dist[j] == low && row4col[j] < 0)) {
                low = dist[j];
                pick = t;
            }
        }
        min_val = low;
        if (min_val == INFINITY) return -1;
        int64_t j = cand[pick];
        if (row4col[j] < 0) sink = j; else i = row4col[j];
        in_sc[j] = 1;
        cand[pick] = cand[--left];
    }
    *out_min = min_val;
    return sink;
}

int oracle_lsap_f32(const float *cost, int64_t nr, int64_t nc, int64_t *rows, int64_t *cols)
{
    if (nr == 0 || nc == 0) return 0;
    const int transpose = nc < nr;
    const int64_t R = transpose ? nc : nr, C = transpose ? nr : nc;
    double *c = (double *)malloc(sizeof(double) * (size_t)(R * C));
    for (int64_t i = 0; i < nr; ++i)
        for (int64_t j = 0; j < nc; ++j) {
            double x = (double)cost[i * nc + j];
            if (transpose) c[j * nr + i] = x; else c[i * nc + j] = x;
        }
    for (int64_t k = 0; k < R * C; ++k)
        if (c[k] != c[k] || c[k] == -INFINITY) { free(c); return 1; }

    double *u = (double *)calloc((size_t)R, sizeof(double));
    double *v = (double *)calloc((size_t)C, sizeof(double));
    double *dist = (double *)malloc(sizeof(double) * (size_t)C);
    int64_t *path = (int64_t *)malloc(sizeof(int64_t) * (size_t)C);
    int64_t *col4row = (int64_t *)malloc(sizeof(int64_t) * (size_t)R);
    int64_t *row4col = (int64_t *)malloc(sizeof(int64_t) * (size_t)C);
    int64_t *cand = (int64_t *)malloc(sizeof(int64_t) * (size_t)C);
    unsigned char *in_sr = (unsigned char *)malloc((size_t)R);
    unsigned char *in_sc = (unsigned char *)malloc((size_t)C);
    for (int64_t j = 0; j < C; ++j) { path[j] = -1; row4col[j] = -1; }
    for (int64_t i = 0; i < R; ++i) col4row[i] = -1;
    int rc = 0;

    for (int64_t cur = 0; cur < R; ++cur) {
        memset(in_sr, 0, (size_t)R);
        memset(in_sc, 0, (size_t)C);
        double min_val;
        int64_t sink = find_path(C, c, u, v, path, row4col, dist, cur, in_sr, in_sc, cand, &min_val);
        if (sink < 0) { rc = 2; break; }
        u[cur] += min_val;
        for (int64_t i = 0; i < R; ++i)
            if (in_sr[i] && i != cur) u[i] += min_val - dist[col4row[i]];
        for (int64_t j = 0; j < C; ++j)
            if (in_sc[j]) v[j] -= min_val - dist[j];
        int64_t j = sink;
        for (;;) {
            int64_t i = path[j];
            row4col[j] = i;
            int64_t t = col4row[i]; col4row[i] = j; j = t;
            if (i == cur) break;
        }
    }
    if (rc == 0) {
        if (transpose) {
            /* pairs (col4row[k], k) sorted by col4row[k]: walk row4col over the original rows */
            int64_t n = 0;
            for (int64_t r = 0; r < C; ++r)
                if (row4col[r] >= 0) { rows[n] = r; cols[n] = row4col[r]; ++n; }
        } else {
            for (int64_t i = 0; i < R; ++i) { rows[i] = i; cols[i] = col4row[i]; }
        }
    }
    free(c); free(u); free(v); free(dist); free(path); free(col4row); free(row4col);
    free(cand); free(in_sr); free(in_sc);
    return rc;
}
