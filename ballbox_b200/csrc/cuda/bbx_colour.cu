// Coloured schedule (BBX_SOLVER_COLOURED, opt-in: REORDERS the Gauss-Seidel sweep, results differ from the reference's,
// SURVEY.md F3 / DESIGN.md 4.7): colour-major replay.
//
// Iteration 0 runs in k_solve_levels<true> (bbx_solver.cu): Kahn levels of the hash-oriented conflict graph as before, and
// every node takes the smallest colour free on both its bodies (greedy colouring in priority order; two nodes of one colour
// never share a dynamic body).  The kernels below rewrite the level-major queues, descriptors and records of that sweep in
// COLOUR-major order into the second buffer set; the host swaps the two sets, and k_replay_staged (unchanged) then replays
// a dozen fat colour classes per iteration instead of ~45 hash levels.  The record slots follow the same closed form as in
// k_solve_levels (RECORD SLOTS there) with the colour in the place of the level.
//
// Identity mode: when iteration 0 replayed the sweep itself (more than RS_MAXLEV levels) or a node found all 64 colours
// taken (colour_overflow), the buffers are copied unchanged, so that everything downstream (staged replay, export, warm
// start) sees one consistent set whatever happened.
#include "bbx_dev.cuh"
#include "bbx_solve.cuh"

#define COL_THREADS 256
#define COL_LC 256                     // levels of iteration 0 whose counts k_col_scatter keeps in shared memory

__device__ __forceinline__ bool col_identity(const BbxDev& dv) { return __ldcg(&dv.ctr->replay_done) != 0 || __ldcg(&dv.ctr->colour_overflow) != 0; }

// nodes per class over all levels; clears the per-(colour, class) counters
__global__ void __launch_bounds__(COL_THREADS) k_col_totals(BbxDev dv)
{
    __shared__ int s_tot[BBX_NCLS];
    if (threadIdx.x < BBX_NCLS) s_tot[threadIdx.x] = 0;
    __syncthreads();
    const int n_levels = __ldcg(&dv.ctr->n_levels);
    int mine[BBX_NCLS];
    for (int k = 0; k < BBX_NCLS; k++) mine[k] = 0;
    for (int l = threadIdx.x; l < n_levels; l += blockDim.x)
        for (int k = 0; k < BBX_NCLS; k++) mine[k] += __ldcg(&dv.level_count[l * BBX_NCLS + k]);
    for (int k = 0; k < BBX_NCLS; k++) if (mine[k]) atomicAdd(&s_tot[k], mine[k]);
    __syncthreads();
    if (threadIdx.x < BBX_NCLS) dv.col_tab[COL_TOT + threadIdx.x] = s_tot[threadIdx.x];
    for (int i = threadIdx.x; i < 2 * BBX_MAX_COLOURS * BBX_NCLS; i += blockDim.x) dv.col_tab[COL_CNT + i] = 0;      // counts and fill
}

// nodes per (colour, class)
__global__ void __launch_bounds__(COL_THREADS) k_col_hist(BbxDev dv)
{
    if (col_identity(dv)) return;
    __shared__ int s_h[BBX_MAX_COLOURS * BBX_NCLS];
    for (int i = threadIdx.x; i < BBX_MAX_COLOURS * BBX_NCLS; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    for (int k = 0; k < BBX_NCLS; k++) {
        const int tot = dv.col_tab[COL_TOT + k];
        // (one shared-memory atomic per distinct colour of a warp: a level's nodes mostly share a handful of colours, and
        //  same-address shared atomics serialise lane by lane)
        for (int e0 = blockIdx.x * blockDim.x; e0 < tot; e0 += gridDim.x * blockDim.x) {
            const int e = e0 + threadIdx.x;
            const bool live = e < tot;
            const int colr = live ? __ldcg(&dv.colq_k[k][e]) : -1;
            const unsigned int peers = __match_any_sync(0xffffffffu, colr);
            if (live && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_h[colr * BBX_NCLS + k], __popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BBX_MAX_COLOURS * BBX_NCLS; i += blockDim.x) if (s_h[i]) atomicAdd(&dv.col_tab[COL_CNT + i], s_h[i]);
}

// the colour-major level table: counts, queue heads and record bases per colour; n_levels becomes the number of colours
__global__ void __launch_bounds__(COL_THREADS) k_col_prefix(BbxDev dv)
{
    const int n_levels = __ldcg(&dv.ctr->n_levels);
    if (col_identity(dv)) {
        for (int i = threadIdx.x; i < (n_levels + 2) * BBX_NCLS; i += blockDim.x) dv.level_count2[i] = __ldcg(&dv.level_count[i]);
        return;
    }
    if (threadIdx.x == 0) {
        int ncol = 0;
        for (int c = 0; c < BBX_MAX_COLOURS; c++)
            for (int k = 0; k < BBX_NCLS; k++) if (dv.col_tab[COL_CNT + c * BBX_NCLS + k] > 0) ncol = c + 1;
        int head[BBX_NCLS], mbase = 0;
        for (int k = 0; k < BBX_NCLS; k++) head[k] = 0;
        for (int c = 0; c < ncol; c++) {
            const int* cnt = dv.col_tab + COL_CNT + c * BBX_NCLS;
            dv.col_tab[COL_MBASE + c] = mbase;
            for (int k = 0; k < BBX_NCLS; k++) {
                dv.col_tab[COL_HEAD + c * BBX_NCLS + k] = head[k];
                head[k] += cnt[k];
                dv.level_count2[c * BBX_NCLS + k] = cnt[k];
            }
            mbase += cnt[0] + cnt[1] + 2 * cnt[2] + 3 * cnt[3] + 4 * cnt[4];
        }
        for (int k = 0; k < 2 * BBX_NCLS; k++) dv.level_count2[ncol * BBX_NCLS + k] = 0;
        dv.col_tab[COL_MBASE + BBX_MAX_COLOURS] = n_levels;            // levels of iteration 0 (k_col_scatter walks them)
        dv.ctr->n_levels = ncol;
    }
}

// every node moves from its (level, class) position to its (colour, class) position: queue entry, descriptor and records
__global__ void __launch_bounds__(COL_THREADS) k_col_scatter(BbxDev dv)
{
    __shared__ int s_h[BBX_MAX_COLOURS], s_base[BBX_MAX_COLOURS];
    __shared__ int s_lc[COL_LC * BBX_NCLS];
    __shared__ int s_tab[COL_MBASE + BBX_MAX_COLOURS - COL_CNT];                      // counts, (fill), heads, record bases per colour
    const bool ident = col_identity(dv);
    const int n_levels0 = ident ? __ldcg(&dv.ctr->n_levels) : dv.col_tab[COL_MBASE + BBX_MAX_COLOURS];
    const size_t PL = (size_t)dv.cap_contacts;
    // the level table of iteration 0 is walked by every block: keep (the first COL_LC levels of) it in shared memory, a
    // dependent L2 round trip per (level, class) pair would cost more than the block's own tiles
    for (int i = threadIdx.x; i < min(n_levels0, COL_LC) * BBX_NCLS; i += blockDim.x) s_lc[i] = __ldcg(&dv.level_count[i]);
    for (int i = threadIdx.x; i < COL_MBASE + BBX_MAX_COLOURS - COL_CNT; i += blockDim.x) s_tab[i] = __ldcg(&dv.col_tab[COL_CNT + i]);
    __syncthreads();
    const int* t_cnt = s_tab; const int* t_head = s_tab + (COL_HEAD - COL_CNT); const int* t_mbase = s_tab + (COL_MBASE - COL_CNT);
    int head[BBX_NCLS];
    for (int k = 0; k < BBX_NCLS; k++) head[k] = 0;
    // tiles of COL_THREADS nodes, numbered through all (level, class) pairs, are dealt round-robin to the blocks
    int tmod = 0;                                                                     // number of the pair's first tile, modulo the grid
    for (int l = 0; l < n_levels0; l++) {
        for (int k = 0; k < BBX_NCLS; k++) {
            const int cnt0 = l < COL_LC ? s_lc[l * BBX_NCLS + k] : __ldcg(&dv.level_count[l * BBX_NCLS + k]);
            const int stride0 = k <= 4 ? cnt0 : 1;                                   // RECORD SLOTS (k_solve_levels)
            const int ntile = (cnt0 + COL_THREADS - 1) / COL_THREADS;
            const int first = ((int)blockIdx.x - tmod + (int)gridDim.x) % (int)gridDim.x;
            tmod = (int)(((long long)tmod + ntile) % (long long)gridDim.x);
            for (int tile = first; tile < ntile; tile += gridDim.x) {                 // (block-uniform)
                const int idx0 = tile * COL_THREADS + threadIdx.x;
                const bool live = idx0 < cnt0;
                const int e = head[k] + idx0;
                int colr = 0, rank = 0;
                int4 nd = make_int4(0, 0, -1, -1);
                int c = 0;
                if (live) { nd = __ldcg(&dv.Nd_k[k][e]); c = __ldcg(&dv.queue_k[k][e]); }
                if (!ident) {
                    if (threadIdx.x < BBX_MAX_COLOURS) s_h[threadIdx.x] = 0;
                    __syncthreads();
                    colr = live ? __ldcg(&dv.colq_k[k][e]) : -1;
                    {
                        const unsigned int peers = __match_any_sync(0xffffffffu, colr);
                        const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
                        int base = 0;
                        if (live && lane == leader) base = atomicAdd(&s_h[colr], __popc(peers));
                        base = __shfl_sync(0xffffffffu, base, leader);
                        rank = base + __popc(peers & ((1u << lane) - 1u));
                    }
                    __syncthreads();
                    if (threadIdx.x < BBX_MAX_COLOURS && s_h[threadIdx.x] > 0)
                        s_base[threadIdx.x] = atomicAdd(&dv.col_tab[COL_FILL + threadIdx.x * BBX_NCLS + k], s_h[threadIdx.x]);
                    __syncthreads();
                }
                if (live) {
                    int e2 = e, slot2 = nd.x, stride2 = stride0;
                    if (!ident) {
                        const int idx2 = s_base[colr] + rank;
                        const int* cnt = t_cnt + colr * BBX_NCLS;
                        e2 = t_head[colr * BBX_NCLS + k] + idx2;
                        if (k <= 4) {
                            const int off = k == 0 ? 0 : k == 1 ? cnt[0] : k == 2 ? cnt[0] + cnt[1] : k == 3 ? cnt[0] + cnt[1] + 2 * cnt[2]
                                                                                                          : cnt[0] + cnt[1] + 2 * cnt[2] + 3 * cnt[3];
                            slot2 = t_mbase[colr] + off + idx2;
                            stride2 = cnt[k];
                        }                                                             // (classes 5-7 keep their slots at the top of the array)
                    }
                    dv.queue2_k[k][e2] = c;
                    dv.Nd2_k[k][e2] = make_int4(slot2, nd.y, nd.z, nd.w);
                    for (int j = 0; j < nd.y; j++) {
                        const float4* M = dv.M + (size_t)(nd.x + j * stride0);
                        float4* M2 = dv.M2 + (size_t)(slot2 + j * stride2);
                        const float4 a0 = __ldcg(M), a1 = __ldcg(M + PL), a2 = __ldcg(M + 2 * PL), a3 = __ldcg(M + 3 * PL);
                        __stcg(M2, a0); __stcg(M2 + PL, a1); __stcg(M2 + 2 * PL, a2); __stcg(M2 + 3 * PL, a3);
                    }
                }
                if (!ident) __syncthreads();                                          // s_h / s_base are reused by the block's next tile
            }
            head[k] += cnt0;
        }
    }
}

template <typename T> static int col_alloc(BbxDev* d, T** p, size_t count)
{
    if (count == 0) count = 1;
    BBX_CUDA(d, cudaMalloc((void**)p, count * sizeof(T)));
    BBX_CUDA(d, cudaMemsetAsync(*p, 0, count * sizeof(T), d->stream));
    return 0;
}

int bbx_colour_ensure(BbxDev* d)
{
    if (d->M2) return 0;
    static const int min_len[BBX_NCLS] = {1, 1, 2, 3, 4, 5, 7, 9};                 // as for queue_k / Nd_k (bbx_state.cu)
    const size_t C = (size_t)d->cap_contacts;
    int rc;
    for (int k = 0; k < BBX_NCLS; k++) {
        if ((rc = col_alloc(d, &d->queue2_k[k], C / min_len[k] + 64))) return rc;
        if ((rc = col_alloc(d, &d->Nd2_k[k], C / min_len[k] + 64))) return rc;
        if ((rc = col_alloc(d, &d->colq_k[k], C / min_len[k] + 64))) return rc;
    }
    if ((rc = col_alloc(d, &d->level_count2, ((size_t)d->cap_levels + 4) * BBX_NCLS))) return rc;
    if ((rc = col_alloc(d, &d->col_used, (size_t)d->N))) return rc;
    if ((rc = col_alloc(d, &d->col_tab, (size_t)COL_INTS))) return rc;
    if ((rc = col_alloc(d, &d->M2, 4 * C))) return rc;
    return 0;
}

int bbx_colour_rebucket(BbxDev* d)
{
    BBX_LAUNCH(d, k_col_totals, 1, COL_THREADS, 0, *d);
    BBX_LAUNCH(d, k_col_hist, BBX_NUM_SMS * 4, COL_THREADS, 0, *d);
    BBX_LAUNCH(d, k_col_prefix, 1, COL_THREADS, 0, *d);
    BBX_LAUNCH(d, k_col_scatter, BBX_NUM_SMS * 8, COL_THREADS, 0, *d);
    BBX_CUDA(d, cudaGetLastError());
    // the colour-major set becomes the current one (every coloured solve rewrites both sets from the contact-order records)
    { float4* t = d->M; d->M = d->M2; d->M2 = t; }
    { int* t = d->level_count; d->level_count = d->level_count2; d->level_count2 = t; }
    for (int k = 0; k < BBX_NCLS; k++) {
        { int* t = d->queue_k[k]; d->queue_k[k] = d->queue2_k[k]; d->queue2_k[k] = t; }
        { int4* t = d->Nd_k[k]; d->Nd_k[k] = d->Nd2_k[k]; d->Nd2_k[k] = t; }
    }
    return 0;
}
