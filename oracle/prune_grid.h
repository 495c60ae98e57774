/*
 * prune_grid.h -- TEST INFRASTRUCTURE (oracle side only; the product never
 * includes this file).
 *
 * A CPU uniform grid used to skip pairs the reference's all-pairs loops
 * (reference ballbox.h:1986-1987, :2011-2012, :2050-2051) would test and
 * reject.  A pair survives when the bodies' bounding spheres overlap with a
 * safety margin, which is necessary for every narrowphase hit condition
 * (SURVEY.md appendix C): sphere-sphere (:457), sphere-box (:1587), SAT (:1076).
 * Queries return indices in ascending order, so a driver that walks i upward
 * and the returned j upward visits surviving pairs in the reference's order.
 */
#ifndef PRUNE_GRID_H
#define PRUNE_GRID_H
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct {
    int    n;
    const float* pos;      /* xyz, stride 3 floats */
    float* rad;            /* bounding radii (owned) */
    float  cell, ox, oy, oz;
    int    nx, ny, nz;
    int*   cell_start;     /* nx*ny*nz + 1 */
    int*   cell_items;     /* small bodies sorted by cell, ascending index inside a cell */
    int*   big;            /* indices of bodies too large for the grid */
    int    nbig;
} PruneGrid;

static int pg_cmp_int(const void* a, const void* b)
{
    int x = *(const int*)a, y = *(const int*)b;
    return (x > y) - (x < y);
}

static int pg_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void pg_cell_of(const PruneGrid* g, const float* p, int* cx, int* cy, int* cz)
{
    *cx = pg_clampi((int)floorf((p[0] - g->ox) / g->cell), 0, g->nx - 1);
    *cy = pg_clampi((int)floorf((p[1] - g->oy) / g->cell), 0, g->ny - 1);
    *cz = pg_clampi((int)floorf((p[2] - g->oz) / g->cell), 0, g->nz - 1);
}

/* rad[] is copied; small_cut = largest radius that still goes into the grid */
static PruneGrid* pg_build(int n, const float* pos, const float* rad, float small_cut)
{
    PruneGrid* g = (PruneGrid*)calloc(1, sizeof(PruneGrid));
    g->n = n; g->pos = pos;
    g->rad = (float*)malloc(sizeof(float) * (n > 0 ? n : 1));
    memcpy(g->rad, rad, sizeof(float) * n);
    g->big = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    int nsmall = 0;
    for (int i = 0; i < n; i++) {
        if (!(rad[i] <= small_cut)) { g->big[g->nbig++] = i; continue; }
        nsmall++;
        for (int k = 0; k < 3; k++) {
            float v = pos[3 * i + k];
            if (v < lo[k]) lo[k] = v;
            if (v > hi[k]) hi[k] = v;
        }
    }
    if (nsmall == 0) { lo[0] = lo[1] = lo[2] = 0; hi[0] = hi[1] = hi[2] = 1; }
    float cell = 2.0f * small_cut * 1.001f + 1e-4f;
    for (;;) {
        g->nx = (int)((hi[0] - lo[0]) / cell) + 2;
        g->ny = (int)((hi[1] - lo[1]) / cell) + 2;
        g->nz = (int)((hi[2] - lo[2]) / cell) + 2;
        if ((double)g->nx * g->ny * g->nz <= 8.0e6 && g->nx < 2000 && g->ny < 2000 && g->nz < 2000) break;
        cell *= 1.5f;
    }
    g->cell = cell; g->ox = lo[0] - 0.5f * cell; g->oy = lo[1] - 0.5f * cell; g->oz = lo[2] - 0.5f * cell;
    int ncell = g->nx * g->ny * g->nz;
    g->cell_start = (int*)calloc((size_t)ncell + 1, sizeof(int));
    g->cell_items = (int*)malloc(sizeof(int) * (nsmall > 0 ? nsmall : 1));
    int* cid = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        if (!(rad[i] <= small_cut)) { cid[i] = -1; continue; }
        int cx, cy, cz; pg_cell_of(g, pos + 3 * i, &cx, &cy, &cz);
        cid[i] = (cz * g->ny + cy) * g->nx + cx;
        g->cell_start[cid[i] + 1]++;
    }
    for (int c = 0; c < ncell; c++) g->cell_start[c + 1] += g->cell_start[c];
    int* cur = (int*)malloc(sizeof(int) * ((size_t)ncell + 1));
    memcpy(cur, g->cell_start, sizeof(int) * ((size_t)ncell + 1));
    for (int i = 0; i < n; i++) if (cid[i] >= 0) g->cell_items[cur[cid[i]]++] = i;   /* ascending i per cell */
    free(cur); free(cid);
    return g;
}

static void pg_free(PruneGrid* g)
{
    if (!g) return;
    free(g->rad); free(g->cell_start); free(g->cell_items); free(g->big); free(g);
}

static int pg_overlap(const float* a, float ra, const float* b, float rb)
{
    float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
    float d2 = dx * dx + dy * dy + dz * dz, rs = ra + rb;
    return d2 <= rs * rs * 1.001f + 1e-5f;
}

/* all bodies j of the grid (j >= jmin) whose bounding sphere may touch the
 * sphere (p, r); ascending order.  *out grows as needed. */
static int pg_query(const PruneGrid* g, const float* p, float r, int jmin, int** out, int* cap)
{
    int cnt = 0;
#define PG_PUSH(j_) do { if (cnt >= *cap) { *cap = *cap ? *cap * 2 : 64; *out = (int*)realloc(*out, sizeof(int) * (size_t)*cap); } (*out)[cnt++] = (j_); } while (0)
    for (int k = 0; k < g->nbig; k++) {
        int j = g->big[k];
        if (j >= jmin && pg_overlap(p, r, g->pos + 3 * j, g->rad[j])) PG_PUSH(j);
    }
    /* reach in cells: query radius + largest small radius (= cell/2) */
    float reach = r + 0.5f * g->cell;
    float lo[3] = {p[0] - reach, p[1] - reach, p[2] - reach}, hi[3] = {p[0] + reach, p[1] + reach, p[2] + reach};
    int x0, y0, z0, x1, y1, z1;
    pg_cell_of(g, lo, &x0, &y0, &z0); pg_cell_of(g, hi, &x1, &y1, &z1);
    for (int z = z0; z <= z1; z++) for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) {
        int c = (z * g->ny + y) * g->nx + x;
        for (int s = g->cell_start[c]; s < g->cell_start[c + 1]; s++) {
            int j = g->cell_items[s];
            if (j >= jmin && pg_overlap(p, r, g->pos + 3 * j, g->rad[j])) PG_PUSH(j);
        }
    }
#undef PG_PUSH
    if (cnt > 1) qsort(*out, (size_t)cnt, sizeof(int), pg_cmp_int);
    return cnt;
}
#endif /* PRUNE_GRID_H */
