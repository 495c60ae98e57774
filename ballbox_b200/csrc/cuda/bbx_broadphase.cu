// bbx_broadphase.cu -- uniform-grid broadphase (K2-K5).
//
// The reference has NO broadphase: CollectCollisions runs three all-pairs loops
// (ballbox.h:1986-1987, :2011-2012, :2050-2051).  This file replaces the loop
// bounds with a grid: bodies are binned by a one-pass radix (counting) sort on
// their cell key, every small body scans a half stencil of 14 cells, and bodies
// too large for the grid (static walls, ground) are tested against every body of
// their world.  A pair survives when the bounding spheres touch, which is
// necessary for each narrowphase hit condition, so the
// set of contacts is exactly the reference's.  Sphere-sphere pairs are resolved
// here (the test is 20 flops); sphere-box and box-box pairs are compacted into
// candidate lists for the narrowphase kernels.
#include "bbx_dev.cuh"

// ---------------------------------------------------------------------------
// exclusive scan of ints: three small kernels (tile reduce, spine, apply)
// ---------------------------------------------------------------------------
#define SCAN_TILE 2048           // 256 threads x 8 items

__global__ void __launch_bounds__(256) k_scan_reduce(const int* in, int* tile_sums, int n_max, const int* n_dev)
{
    int n = n_dev ? min(*n_dev, n_max) : n_max;
    int base = blockIdx.x * SCAN_TILE;
    if (base >= n) { if (threadIdx.x == 0) tile_sums[blockIdx.x] = 0; return; }
    int s = 0;
    for (int k = 0; k < 8; k++) { int i = base + k * 256 + threadIdx.x; if (i < n) s += in[i]; }
    __shared__ int ws[8];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < 8; k++) t += ws[k]; tile_sums[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(1024) k_scan_spine(int* tile_sums, int n_tiles, int* total_out)
{
    __shared__ int ws[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        int i = base + threadIdx.x;
        int v = (i < n_tiles) ? tile_sums[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = ws[threadIdx.x], z = w;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, z, o); if (threadIdx.x >= o) z += y; }
            ws[threadIdx.x] = z - w;
        }
        __syncthreads();
        int excl = carry + ws[threadIdx.x >> 5] + x - v;
        if (i < n_tiles) tile_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(256) k_scan_apply(const int* in, int* out, int* out2, const int* tile_sums, int n_max, const int* n_dev)
{
    int n = n_dev ? min(*n_dev, n_max) : n_max;
    int base = blockIdx.x * SCAN_TILE;
    if (base >= n) return;
    // each thread owns 8 consecutive items
    int i0 = base + threadIdx.x * 8;
    int v[8]; int s = 0;
    for (int k = 0; k < 8; k++) { v[k] = (i0 + k < n) ? in[i0 + k] : 0; s += v[k]; }
    __shared__ int ws[8];
    int x = s;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < (threadIdx.x >> 5); k++) woff += ws[k];
    int run = tile_sums[blockIdx.x] + woff + x - s;
    for (int k = 0; k < 8; k++) {
        if (i0 + k < n) { out[i0 + k] = run; if (out2) out2[i0 + k] = run; }
        run += v[k];
    }
}

// out[i] = sum(in[0..i)); n = min(*n_dev, n_max) when n_dev != NULL.  total_out
// (device pointer, optional) receives the grand total.
int bbx_scan_exclusive(BbxDev* d, const int* in, int* out, int* out2, int n_max, const int* n_dev, int* total_out)
{
    int tiles = (n_max + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles < 1) tiles = 1;
    if (tiles > (1 << 16)) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "scan too large", __FILE__, __LINE__);
    BBX_LAUNCH(d, k_scan_reduce, tiles, 256, 0, in, d->scan_tmp, n_max, n_dev);
    BBX_LAUNCH(d, k_scan_spine, 1, 1024, 0, d->scan_tmp, tiles, total_out);
    BBX_LAUNCH(d, k_scan_apply, tiles, 256, 0, in, out, out2, d->scan_tmp, n_max, n_dev);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// bounds of the small bodies + per-step counter reset
// ---------------------------------------------------------------------------
__global__ void k_step_reset(BbxDev dv)
{
    StepCounters* c = dv.ctr;
    c->n_sorted = 0; c->n_cand_sb = 0; c->n_cand_bb = 0; c->n_contacts = 0; c->n_solver = 0; c->n_touched = 0;
    c->n_levels = 0; c->n_ss_tests = 0; c->n_frontier0 = 0; c->n_mslots = 0; c->n_nodes = 0; c->fq_tail = 0; c->fq_head = 0; c->fq_ticket = 0; c->n_hit_bb = 0; c->n_hit_bb_edge = 0; c->n_surv_bb = 0; c->n_deferred = 0;
    c->bounds[0] = c->bounds[1] = c->bounds[2] = 0xffffffffu;
    c->bounds[3] = c->bounds[4] = c->bounds[5] = 0u;
}

__global__ void __launch_bounds__(256) k_bounds(BbxDev dv)
{
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dv.N; i += gridDim.x * blockDim.x) {
        unsigned char f = dv.flags[i];
        if (!(f & BF_VALID) || (f & BF_BIG)) continue;
        float4 p = dv.pos[i];
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    // one atomic per block and bound (six addresses for the whole grid: per-warp atomics were most of this kernel's time)
    __shared__ float s_lo[3][8], s_hi[3][8];
    if ((threadIdx.x & 31) == 0) for (int k = 0; k < 3; k++) { s_lo[k][threadIdx.x >> 5] = lo[k]; s_hi[k][threadIdx.x >> 5] = hi[k]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int k = threadIdx.x;
        float l = s_lo[k][0], h = s_hi[k][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) { l = fminf(l, s_lo[k][w]); h = fmaxf(h, s_hi[k][w]); }
        if (l < 3.0e38f) atomicMin(&dv.ctr->bounds[k], f2ord(l));
        if (h > -3.0e38f) atomicMax(&dv.ctr->bounds[3 + k], f2ord(h));
    }
}

// one thread: choose the cell size and the grid dimensions for this step
__global__ void k_grid_setup(BbxDev dv)
{
    GridParams* g = dv.grid;
    StepCounters* c = dv.ctr;
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) { lo[k] = ord2f(c->bounds[k]); hi[k] = ord2f(c->bounds[3 + k]); }
    if (c->bounds[0] == 0xffffffffu || !(hi[0] >= lo[0]) || !(hi[1] >= lo[1]) || !(hi[2] >= lo[2])) {
        lo[0] = lo[1] = lo[2] = 0.0f; hi[0] = hi[1] = hi[2] = 1.0f;       // no small bodies (or NaNs)
    }
    // clamp absurd extents (a body that flew away): bodies outside are binned into border cells
    for (int k = 0; k < 3; k++) { lo[k] = fmaxf(lo[k], -1.0e18f); hi[k] = fminf(hi[k], 1.0e18f); }
    float cell = 2.0f * g->small_cut * 1.01f + 1e-4f;      // 1% slack: cell-coordinate rounding can never split a touching pair by 2 cells
    long long budget = dv.cap_cells / dv.n_worlds;
    int nx, ny, nz;
    for (int it = 0; it < 200; it++) {
        float ex = (hi[0] - lo[0]) / cell, ey = (hi[1] - lo[1]) / cell, ez = (hi[2] - lo[2]) / cell;
        if (ex < 30000.0f && ey < 30000.0f && ez < 30000.0f) {
            nx = (int)ex + 2; ny = (int)ey + 2; nz = (int)ez + 2;
            if ((long long)nx * ny * nz <= budget) break;
        }
        cell *= 1.25f;
        nx = ny = nz = 1;
    }
    if ((long long)nx * ny * nz > budget) { nx = ny = nz = 1; g->overflow = 1; }
    g->cell = cell; g->inv_cell = 1.0f / cell;
    g->ox = lo[0] - 0.5f * cell; g->oy = lo[1] - 0.5f * cell; g->oz = lo[2] - 0.5f * cell;
    g->nx = nx; g->ny = ny; g->nz = nz;
    g->cells_per_world = nx * ny * nz;
    g->n_cells = g->cells_per_world * dv.n_worlds;
    g->n_scan = g->n_cells + 1;
}

__device__ __forceinline__ int cell_coord(float p, float o, float inv, int n)
{
    float f = floorf((p - o) * inv);
    // NaN and out-of-range positions land in a border cell (still conservative: see DESIGN.md)
    int c = (f >= 0.0f) ? ((f < (float)n) ? (int)f : n - 1) : 0;
    return c;
}

__device__ __forceinline__ int world_of(const BbxDev& dv, int g)
{
    return (g < dv.NS) ? g / max(dv.cap_s, 1) : (g - dv.NS) / max(dv.cap_b, 1);
}

__global__ void __launch_bounds__(256) k_cell_count(BbxDev dv)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dv.N) return;
    unsigned char f = dv.flags[i];
    if (!(f & BF_VALID) || (f & BF_BIG)) { dv.body_cell[i] = -1; return; }
    const GridParams g = *dv.grid;
    float4 p = dv.pos[i];
    int cx = cell_coord(p.x, g.ox, g.inv_cell, g.nx);
    int cy = cell_coord(p.y, g.oy, g.inv_cell, g.ny);
    int cz = cell_coord(p.z, g.oz, g.inv_cell, g.nz);
    int c = world_of(dv, i) * g.cells_per_world + (cz * g.ny + cy) * g.nx + cx;
    dv.body_cell[i] = c;
    atomicAdd(&dv.cell_count[c], 1);
}

__global__ void __launch_bounds__(256) k_cell_scatter(BbxDev dv)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dv.N) return;
    int c = dv.body_cell[i];
    if (c < 0) return;
    int slot = atomicAdd(&dv.cell_cursor[c], 1);
    dv.sorted_id[slot] = i;
    dv.sorted_pos[slot] = dv.pos[i];
}

// ---------------------------------------------------------------------------
// pair finding.  One thread per grid slot (cell order, so a warp's bodies are
// spatial neighbours and share most of their stencil in L1/L2).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void emit_ss_contact(const BbxDev& dv, int a, float4 pa, int b, float4 pb)
{
    // CheckSphereCollisionWithData, ballbox.h:451-471, with A = lower index
    float3 n = v3sub(xyz(pb), xyz(pa));
    float dist = v3len(n);
    float rs = pa.w + pb.w;
    bool hit = dist < rs;
    int slot = warp_append(&dv.ctr->n_contacts, hit);
    if (!hit) return;
    if (slot >= dv.cap_contacts) { atomicOr(&dv.ctr->overflow, 2); return; }
    float3 nrm = (dist > 1e-6f) ? v3scale(n, 1.0f / dist) : f3(1.0f, 0.0f, 0.0f);
    float3 pt = v3add(xyz(pa), v3scale(nrm, pa.w));
    dv.c_ids[slot] = make_int4(a, b, RUN_ENC(0, 1), PH_SS);
    dv.c_pt[slot] = make_float4(pt.x, pt.y, pt.z, rs - dist);
    dv.c_nrm[slot] = make_float4(nrm.x, nrm.y, nrm.z, 0.0f);
}

__device__ __forceinline__ bool bounds_touch(float4 a, float4 b)
{
    float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
    float d2 = dx * dx + dy * dy + dz * dz, rs = a.w + b.w;
    return d2 <= rs * rs * 1.001f + 1e-5f;
}

// classify one geometric neighbour pair (i, j) and route it
__device__ __forceinline__ void route_pair(const BbxDev& dv, int i, float4 pi, int j, float4 pj)
{
    bool ib = i >= dv.NS, jb = j >= dv.NS;
    if (!ib && !jb) {
        if (i < j) emit_ss_contact(dv, i, pi, j, pj); else emit_ss_contact(dv, j, pj, i, pi);
        return;
    }
    bool touch = bounds_touch(pi, pj);
    if (ib && jb) {
        int slot = warp_append(&dv.ctr->n_cand_bb, touch);
        if (!touch) return;
        if (slot >= dv.cap_cand) { atomicOr(&dv.ctr->overflow, 1); return; }
        dv.cand_bb[slot] = (i < j) ? make_int2(i, j) : make_int2(j, i);
    } else {
        int slot = warp_append(&dv.ctr->n_cand_sb, touch);
        if (!touch) return;
        if (slot >= dv.cap_cand) { atomicOr(&dv.ctr->overflow, 1); return; }
        dv.cand_sb[slot] = ib ? make_int2(j, i) : make_int2(i, j);       // (sphere, box)
    }
}

// Found pairs are buffered per thread and flushed with ONE global atomic per block
// and list (same-address atomics serialise in L2: one per found pair cost 2 ms at
// 1M bodies, see profiles/r01a).  Entry = partner gid | list << 28.
#define FP_CAP 40
#define FP_SS 0u
#define FP_SB 1u
#define FP_BB 2u

__device__ __forceinline__ void fp_consider(const BbxDev& dv, int i, float4 pi, int j, float4 pj, unsigned int* buf, int& nb)
{
    bool ib = i >= dv.NS, jb = j >= dv.NS;
    unsigned int list;
    if (!ib && !jb) {
        // the sphere-sphere hit condition itself (ballbox.h:457), symmetric in (i, j)
        float3 n = v3sub(xyz(pj), xyz(pi));
        if (!(v3len(n) < pi.w + pj.w)) return;
        list = FP_SS;
    } else {
        if (!bounds_touch(pi, pj)) return;
        list = (ib && jb) ? FP_BB : FP_SB;
    }
    if (nb < FP_CAP) buf[nb++] = (unsigned int)j | (list << 28);
    else route_pair(dv, i, pi, j, pj);                  // rare overflow: immediate append
}

__global__ void __launch_bounds__(128) k_find_pairs(BbxDev dv)
{
    __shared__ int s_cnt[3], s_base[3];
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    const GridParams g = *dv.grid;
    int n_sorted = dv.cell_start[g.n_cells];
    unsigned int buf[FP_CAP];
    int nb = 0, i = -1;
    float4 pi = make_float4(0, 0, 0, 0);
    if (p < n_sorted) {
        i = dv.sorted_id[p];
        pi = dv.sorted_pos[p];
        int c = dv.body_cell[i];
        int w = c / g.cells_per_world;
        int lc = c - w * g.cells_per_world;
        int cx = lc % g.nx, cy = (lc / g.nx) % g.ny, cz = lc / (g.nx * g.ny);
        int wbase = w * g.cells_per_world;
        // Candidate runs of the sorted order: the rest of the own cell and the 13 forward neighbour cells
        // ((dz,dy,dx) > (0,0,0) lexicographically; cells of one row are consecutive, so a row is one run).
        // The runs are walked by ONE loop: the lanes of a warp then diverge only by their total candidate
        // count, not by the length of every single run.
        int rs[6], re[6], nr = 0;
        rs[0] = p + 1; re[0] = dv.cell_start[c + 1]; nr = 1;
        for (int dz = 0; dz <= 1; dz++) {
            int z = cz + dz;
            if (z >= g.nz) continue;
            for (int dy = (dz ? -1 : 0); dy <= 1; dy++) {
                int y = cy + dy;
                if (y < 0 || y >= g.ny) continue;
                int x0 = (dz == 0 && dy == 0) ? cx + 1 : cx - 1;
                int x1 = cx + 1;
                if (x0 < 0) x0 = 0;
                if (x1 >= g.nx) x1 = g.nx - 1;
                if (x0 > x1) continue;
                int row = wbase + (z * g.ny + y) * g.nx;
                rs[nr] = dv.cell_start[row + x0]; re[nr] = dv.cell_start[row + x1 + 1]; nr++;
            }
        }
        {
            int r = 0, q = rs[0], e = re[0];
            while (true) {
                while (q >= e) { if (++r >= nr) break; q = rs[r]; e = re[r]; }
                if (r >= nr) break;
                fp_consider(dv, i, pi, dv.sorted_id[q], dv.sorted_pos[q], buf, nb);
                q++;
            }
        }
        // big bodies of this world
        int bs = dv.big_start[w], be = dv.big_start[w + 1];
        for (int k = bs; k < be; k++) {
            int j = dv.big_list[k];
            fp_consider(dv, i, pi, j, dv.pos[j], buf, nb);
        }
    }
    // ---- flush: block-aggregated reservation, one global atomic per list ----
    int cnt[3] = {0, 0, 0};
    for (int k = 0; k < nb; k++) cnt[buf[k] >> 28]++;
    int off[3];
    for (int t = 0; t < 3; t++) off[t] = cnt[t] ? atomicAdd(&s_cnt[t], cnt[t]) : 0;
    __syncthreads();
    if (threadIdx.x < 3 && s_cnt[threadIdx.x] > 0) {
        int* ctr = threadIdx.x == FP_SS ? &dv.ctr->n_contacts : (threadIdx.x == FP_SB ? &dv.ctr->n_cand_sb : &dv.ctr->n_cand_bb);
        s_base[threadIdx.x] = atomicAdd(ctr, s_cnt[threadIdx.x]);
    }
    __syncthreads();
    for (int k = 0; k < nb; k++) {
        unsigned int e = buf[k];
        int j = (int)(e & 0x0fffffffu);
        unsigned int t = e >> 28;
        int slot = s_base[t] + off[t]++;
        if (t == FP_SS) {
            if (slot >= dv.cap_contacts) { atomicOr(&dv.ctr->overflow, 2); continue; }
            // CheckSphereCollisionWithData, ballbox.h:451-471, with A = lower index
            int a = i < j ? i : j, b = i < j ? j : i;
            float4 pj = dv.pos[j];
            float4 pa = i < j ? pi : pj, pb = i < j ? pj : pi;
            float3 n = v3sub(xyz(pb), xyz(pa));
            float dist = v3len(n), rs = pa.w + pb.w;
            float3 nrm = (dist > 1e-6f) ? v3scale(n, 1.0f / dist) : f3(1.0f, 0.0f, 0.0f);
            float3 pt = v3add(xyz(pa), v3scale(nrm, pa.w));
            dv.c_ids[slot] = make_int4(a, b, RUN_ENC(0, 1), PH_SS);
            dv.c_pt[slot] = make_float4(pt.x, pt.y, pt.z, rs - dist);
            dv.c_nrm[slot] = make_float4(nrm.x, nrm.y, nrm.z, 0.0f);
        } else {
            if (slot >= dv.cap_cand) { atomicOr(&dv.ctr->overflow, 1); continue; }
            if (t == FP_BB) dv.cand_bb[slot] = (i < j) ? make_int2(i, j) : make_int2(j, i);
            else dv.cand_sb[slot] = (i >= dv.NS) ? make_int2(j, i) : make_int2(i, j);       // (sphere, box)
        }
    }
}

// ---------------------------------------------------------------------------
// The same walk with the candidates of a WARP flattened.  In k_find_pairs every lane walks the ~37 candidates of its own
// body; the lanes of a warp finish at different times (14 of 32 active on average, ncu).  Here every lane still sets up
// the runs of its own body, but the warp then walks the concatenation of all 32 candidate lists: position t of that
// list belongs to the lane o whose prefix interval contains t (binary search over the inclusive prefix by shuffles),
// the owner's runs, id and position are read from shared memory, and whichever lane holds t tests the pair.  Same pairs,
// all lanes busy until the warp's list ends.  Entries carry both body ids.
// ---------------------------------------------------------------------------
#define FPF_WARPS 4

__global__ void __launch_bounds__(FPF_WARPS * 32) k_find_pairs_flat(BbxDev dv)
{
    __shared__ int s_cnt[3], s_base[3];
    __shared__ int s_run[FPF_WARPS][32][12];          // per lane: 6 x (start, length)
    __shared__ float4 s_pos[FPF_WARPS][32];
    __shared__ int s_id[FPF_WARPS][32];
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const GridParams g = *dv.grid;
    const int n_sorted = dv.cell_start[g.n_cells];
    unsigned long long buf[FP_CAP];                   // (body i << 32) | body j | list << 28
    int nb = 0, i = -1, wld = 0;
    float4 pi = make_float4(0, 0, 0, 0);
    int cnt = 0;
    #pragma unroll
    for (int k = 0; k < 12; k++) s_run[wib][lane][k] = 0;
    if (p < n_sorted) {
        i = dv.sorted_id[p];
        pi = dv.sorted_pos[p];
        const int c = dv.body_cell[i];
        wld = c / g.cells_per_world;
        const int lc = c - wld * g.cells_per_world;
        const int cx = lc % g.nx, cy = (lc / g.nx) % g.ny, cz = lc / (g.nx * g.ny);
        const int wbase = wld * g.cells_per_world;
        // run 0: the rest of the own cell; runs 1-5: the forward neighbour rows (see k_find_pairs)
        { const int st = p + 1, ln = dv.cell_start[c + 1] - st; s_run[wib][lane][0] = st; s_run[wib][lane][1] = ln; cnt = ln; }
        #pragma unroll
        for (int k = 1; k < 6; k++) {
            const int dz = k >= 3 ? 1 : 0, dy = k == 1 ? 0 : (k == 2 ? 1 : k - 4);
            const int z = cz + dz, y = cy + dy;
            int x0 = (k == 1) ? cx + 1 : cx - 1, x1 = cx + 1;
            if (x0 < 0) x0 = 0;
            if (x1 >= g.nx) x1 = g.nx - 1;
            if (z < g.nz && y >= 0 && y < g.ny && x0 <= x1) {
                const int row = wbase + (z * g.ny + y) * g.nx;
                const int st = dv.cell_start[row + x0], ln = dv.cell_start[row + x1 + 1] - st;
                s_run[wib][lane][2 * k] = st; s_run[wib][lane][2 * k + 1] = ln; cnt += ln;
            }
        }
    }
    s_pos[wib][lane] = pi; s_id[wib][lane] = i;
    int incl = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    const int T = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        // owner = number of lanes whose inclusive prefix is <= t
        int o = 0;
        #pragma unroll
        for (int step = 16; step > 0; step >>= 1) { const int v = __shfl_sync(0xffffffffu, incl, o + step - 1); if (v <= t) o += step; }
        o = min(o, 31);
        const int cnt_o = __shfl_sync(0xffffffffu, cnt, o), incl_o = __shfl_sync(0xffffffffu, incl, o);
        if (t < T) {
            int li = t - (incl_o - cnt_o), q = 0;
            const int* run = s_run[wib][o];
            #pragma unroll
            for (int k = 0; k < 6; k++) { const int ln = run[2 * k + 1]; if (li >= 0 && li < ln) q = run[2 * k] + li; li -= ln; if (li < 0) li = -0x40000000; }
            const int io = s_id[wib][o], j = dv.sorted_id[q];
            const float4 pio = s_pos[wib][o], pj = dv.sorted_pos[q];
            const bool ib = io >= dv.NS, jb = j >= dv.NS;
            bool hit; unsigned int list;
            if (!ib && !jb) {
                // the sphere-sphere hit condition itself (ballbox.h:457), symmetric in (i, j)
                const float3 n3 = v3sub(xyz(pj), xyz(pio));
                hit = v3len(n3) < pio.w + pj.w; list = FP_SS;
            } else { hit = bounds_touch(pio, pj); list = (ib && jb) ? FP_BB : FP_SB; }
            if (hit) {
                if (nb < FP_CAP) buf[nb++] = ((unsigned long long)(unsigned int)io << 32) | (unsigned int)j | (list << 28);
                else route_pair(dv, io, pio, j, pj);              // rare overflow: immediate append
            }
        }
    }
    // big bodies of this world: every lane for its own body
    if (i >= 0) {
        const int bs = dv.big_start[wld], be = dv.big_start[wld + 1];
        for (int k = bs; k < be; k++) {
            const int j = dv.big_list[k];
            const float4 pj = dv.pos[j];
            const bool ib = i >= dv.NS, jb = j >= dv.NS;
            bool hit; unsigned int list;
            if (!ib && !jb) { const float3 n3 = v3sub(xyz(pj), xyz(pi)); hit = v3len(n3) < pi.w + pj.w; list = FP_SS; }
            else { hit = bounds_touch(pi, pj); list = (ib && jb) ? FP_BB : FP_SB; }
            if (hit) {
                if (nb < FP_CAP) buf[nb++] = ((unsigned long long)(unsigned int)i << 32) | (unsigned int)j | (list << 28);
                else route_pair(dv, i, pi, j, pj);
            }
        }
    }
    // ---- flush: block-aggregated reservation, one global atomic per list ----
    int cn[3] = {0, 0, 0};
    for (int k = 0; k < nb; k++) cn[((unsigned int)buf[k] >> 28) & 3u]++;
    int off[3];
    for (int t = 0; t < 3; t++) off[t] = cn[t] ? atomicAdd(&s_cnt[t], cn[t]) : 0;
    __syncthreads();
    if (threadIdx.x < 3 && s_cnt[threadIdx.x] > 0) {
        int* ctr = threadIdx.x == FP_SS ? &dv.ctr->n_contacts : (threadIdx.x == FP_SB ? &dv.ctr->n_cand_sb : &dv.ctr->n_cand_bb);
        s_base[threadIdx.x] = atomicAdd(ctr, s_cnt[threadIdx.x]);
    }
    __syncthreads();
    for (int k = 0; k < nb; k++) {
        const unsigned long long e = buf[k];
        const int bi = (int)(e >> 32), j = (int)((unsigned int)e & 0x0fffffffu);
        const unsigned int t = ((unsigned int)e >> 28) & 3u;
        const int slot = s_base[t] + off[t]++;
        if (t == FP_SS) {
            if (slot >= dv.cap_contacts) { atomicOr(&dv.ctr->overflow, 2); continue; }
            // CheckSphereCollisionWithData, ballbox.h:451-471, with A = lower index
            const int a = bi < j ? bi : j, b = bi < j ? j : bi;
            const float4 pa = dv.pos[a], pb = dv.pos[b];
            float3 n3 = v3sub(xyz(pb), xyz(pa));
            float dist = v3len(n3), rsum = pa.w + pb.w;
            float3 nrm = (dist > 1e-6f) ? v3scale(n3, 1.0f / dist) : f3(1.0f, 0.0f, 0.0f);
            float3 pt = v3add(xyz(pa), v3scale(nrm, pa.w));
            dv.c_ids[slot] = make_int4(a, b, RUN_ENC(0, 1), PH_SS);
            dv.c_pt[slot] = make_float4(pt.x, pt.y, pt.z, rsum - dist);
            dv.c_nrm[slot] = make_float4(nrm.x, nrm.y, nrm.z, 0.0f);
        } else {
            if (slot >= dv.cap_cand) { atomicOr(&dv.ctr->overflow, 1); continue; }
            if (t == FP_BB) dv.cand_bb[slot] = (bi < j) ? make_int2(bi, j) : make_int2(j, bi);
            else dv.cand_sb[slot] = (bi >= dv.NS) ? make_int2(j, bi) : make_int2(bi, j);       // (sphere, box)
        }
    }
}

// big-big pairs inside each world (few bodies: one thread per big body)
__global__ void k_find_pairs_big(BbxDev dv)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int nbig = dv.big_start[dv.n_worlds];
    if (k >= nbig) return;
    int i = dv.big_list[k];
    int w = world_of(dv, i);
    float4 pi = dv.pos[i];
    for (int m = dv.big_start[w]; m < dv.big_start[w + 1]; m++) {
        int j = dv.big_list[m];
        if (j <= i) continue;
        route_pair(dv, i, pi, j, dv.pos[j]);
    }
}

int bbx_stage_broadphase(BbxDev* d, const BbxStepParams* p)
{
    (void)p;
    int nb = bbx_blocks(d->N, 256);
    BBX_LAUNCH(d, k_step_reset, 1, 1, 0, *d);
    BBX_LAUNCH(d, k_bounds, BBX_NUM_SMS * 4, 256, 0, *d);
    BBX_LAUNCH(d, k_grid_setup, 1, 1, 0, *d);
    BBX_CUDA(d, cudaMemsetAsync(d->cell_count, 0, sizeof(int) * ((size_t)d->cap_cells + 1), d->stream));
    BBX_LAUNCH(d, k_cell_count, nb, 256, 0, *d);
    // n = n_cells + 1 so that cell_start[n_cells] is the number of binned bodies
    // (only the n_cells + 1 entries this step's grid uses are scanned, not the table's capacity: 42 -> 6 us at 1M bodies)
    int rc = bbx_scan_exclusive(d, d->cell_count, d->cell_start, d->cell_cursor, d->cap_cells + 1, &d->grid->n_scan, NULL);
    if (rc) return rc;
    BBX_LAUNCH(d, k_cell_scatter, nb, 256, 0, *d);
    // measured: the flattened walk wins on batches of small worlds (C5: broadphase 0.328 -> 0.283 ms; 512 worlds: 0.112 ->
    // 0.087 ms), the lane-per-body walk on one large world (C3: 0.390 vs 0.448 ms); BBX_FIND_PAIRS=0 / 1 forces one of them
    const int flat = d->find_pairs_mode < 0 ? (d->n_worlds >= 16) : d->find_pairs_mode;
    if (!flat) BBX_LAUNCH(d, k_find_pairs, bbx_blocks(d->N, 128), 128, 0, *d);
    else       BBX_LAUNCH(d, k_find_pairs_flat, bbx_blocks(d->N, FPF_WARPS * 32), FPF_WARPS * 32, 0, *d);
    BBX_LAUNCH(d, k_find_pairs_big, bbx_blocks(d->N, 128), 128, 0, *d);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}
