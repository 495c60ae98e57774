// bbx_replay.cu -- K11b: iterations 1..I-1 of the exact level schedule as a STAGED replay.
//
// Reference: SolveVelocityConstraints ballbox.h:1410-1489 (the sweep itself is solve_constraint, bbx_solve.cuh).
//
// After iteration 0 (k_solve_levels, bbx_solver.cu) the schedule is static: level l holds, per node class k, the
// descriptors Nd_k[k][head(l,k) .. +cnt(l,k)) = (first record slot, constraint count, body A, body B) and the
// plane-major records M.  A level is cut into 32-node chunks per class, so a warp is homogeneous.
//
// Round 1 replayed this inside k_solve_levels with plain loads: after every grid barrier every warp walked the chain
// descriptor (DRAM) -> bodies (L2) + record (DRAM) -> sweep -> store, all warps of the chip in the same phase, 6 us
// per chunk where the sweep alone needs 1.5 us (profiles/r01_notes.md).  Here every warp runs a small software
// pipeline through lane-private shared-memory slots filled by cp.async (LDGSTS, no registers held while in flight):
//   * everything static is fetched ahead of use: the descriptors of a warp's chunks r+2, the first record of chunk
//     r+1, and -- across the grid barrier -- the descriptors and the first record of its first chunk of the NEXT
//     level in which it has work (they do not depend on the level that is running);
//   * the body gather of chunk r+1 (velocities, box rotation / inertia; L2) is issued before chunk r is swept: nodes
//     of one level never share a dynamic body, so chunk r's stores cannot touch them;
//   * what remains exposed per level is ONE body round trip after the barrier.
// WORK DISTRIBUTION.  ncu on the round-1 replay and on a first staged version with the round-1 map (chunks dealt
// round-robin in class order) showed the same picture: ~48 % of the warp time in the grid barrier, long-scoreboard
// stalls only ~12 %: the warps wait for the SLOWEST warp of the level, not for memory.  A chunk costs ~2.4 us per
// constraint of its nodes (per-warp trace, BBX_TRACE_LEVEL) and the third round of a fat level consisted of the
// heaviest chunks (3-6 constraints per node): 8 constraints in a row on the slowest warp where the mean is 3.7.
// Now every level gets a packing, computed once per launch (one thread per level): class k may put q_k = T / cost_k
// chunks on a warp, classes take consecutive warp ranges from the heaviest down, and T is the smallest budget for
// which all chunks fit on the W warps of the grid (binary search).  A warp's chunks are consecutive chunks of ONE
// class, so its class and range follow from two table reads; warp indices go round-robin over the blocks, so every
// SM gets the same mix of classes.  Measured alternatives, both slower (profiles/r02_notes.md): an equal share of
// every class per block with the chunks dealt to its warps by longest-processing-time-first, and the same with
// chunks pulled through a shared-memory ticket counter (the look-ahead of the staging hands out all tickets at once).
// Per-body constraint order, operations and rounding are those of k_solve_levels: bit-identical results.
#include <stdlib.h>
#include <stdio.h>
#include <cooperative_groups.h>
#include "bbx_solve.cuh"
namespace cg = cooperative_groups;

#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_BLOCKS_PER_SM 2
// per warp: descriptor ring (2 x 512 B), first record (4 planes x 512 B), bodies (A vel 2, A box 2, B vel 2, B box 2 slots x 512 B)
#define RS_OFF_DESC 0
#define RS_OFF_REC  1024
#define RS_OFF_BODY 3072
#define RS_WARP_BYTES 7168
#define RS_TABLE_BYTES (RS_MAXLEV * (4 * BBX_NCLS + 1) * 4)
// measured cost of one 32-node chunk of class k in 0.1 us (C3, 16 warps per SM): m = 1 (no box), 1 (box), 2, 3, 4, 5-6, 7-8, 9+
// (per-warp trace of fat levels with the branch-free sweep, tools/profile_step.py with BBX_TRACE_LEVEL: (busy - first data
// wait) / chunks; BBX_RS_COST="c0,...,c7" overrides the table for measurements)
struct ReplayCost { int c[BBX_NCLS]; };
static const int rs_cost_default[BBX_NCLS] = { 21, 22, 37, 50, 59, 78, 105, 180 };

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(unsigned int saddr, const void* g, unsigned long long pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(saddr), "l"(g), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ float4 lds128(unsigned int saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ int4 lds128i(unsigned int saddr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns)); return ns; }

struct ReplayCtx {
    const int* t_cnt;      // [level][class] nodes
    const int* t_head;     // [level][class] queue position of the level's first node
    const int* t_q;        // [level][class] chunks of the class per warp
    const int* t_b;        // [level][class] first warp of the class's range (classes from the heaviest down)
    const int* t_nw;       // [level] warps with work
    unsigned int sbase;    // this lane's 16-byte column in the warp's staging area
    unsigned long long pol_keep, pol_stream;
    size_t PL;
};

// the chunks warp w owns in level l: descriptor pointer of this lane in the first one, number of chunks, and the
// number of nodes from this lane's first node to the end of the class (lane is live in chunk r iff r * 32 < left)
struct WarpWork { const int4* p; int nc; int left; int stride; int cls; };   // stride: slots between a node's consecutive records
__device__ __forceinline__ WarpWork level_work(const BbxDev& dv, const ReplayCtx& cx, int l, int w, int lane)
{
    const int* b = cx.t_b + l * BBX_NCLS;
    int k = BBX_NCLS - 1;
    #pragma unroll
    for (int kk = BBX_NCLS - 2; kk >= 0; kk--) if (w >= b[kk]) k = kk;       // ranges run from class 7 down to class 0
    const int q = cx.t_q[l * BBX_NCLS + k], cnt = cx.t_cnt[l * BBX_NCLS + k];
    const int first = (w - b[k]) * q;                                        // first chunk (of the class) of this warp
    const int nchunks = (cnt + 31) >> 5;
    WarpWork ww;
    ww.nc = min(q, nchunks - first);
    ww.left = cnt - (first * 32 + lane);
    ww.p = dv.Nd_k[k] + (cx.t_head[l * BBX_NCLS + k] + first * 32 + lane);
    ww.cls = k;
    ww.stride = k <= 4 ? cnt : 1;                                            // RECORD SLOTS, bbx_solver.cu
    return ww;
}
// the next (iteration, level) after (it, l) in which warp w owns chunks, searched 32 levels at a time by the warp's
// lanes (a lane-serial scan over the levels cost the warps at the upper end of a thin level's range ~1.5 us: they
// have no work in the following, smaller levels and walk to the next iteration -- and a thin level lasts as long as
// its slowest warp); nit = I when there is none
__device__ __forceinline__ void next_work(const ReplayCtx& cx, int n_levels, int I, int w, int lane, int it, int l, int& nit, int& nl)
{
    int base = l + 1;
    for (;;) {
        if (base >= n_levels) { base = 0; if (++it >= I) { nit = I; nl = 0; return; } }
        const int lv = base + lane;
        const unsigned int m = __ballot_sync(0xffffffffu, lv < n_levels && w < cx.t_nw[lv]);
        if (m) { nit = it; nl = base + __ffs(m) - 1; return; }
        base += 32;
    }
}

// descriptor of the warp's chunk r -> ring slot
__device__ __forceinline__ void issue_desc(const ReplayCtx& cx, const WarpWork& ww, int r, int slot)
{
    if (r * 32 < ww.left) cp16(cx.sbase + RS_OFF_DESC + slot * 512, ww.p + r * 32, cx.pol_stream);
}
// first record of a node (4 planes)
__device__ __forceinline__ void issue_rec0(const BbxDev& dv, const ReplayCtx& cx, int4 nd)
{
    const float4* M = dv.M + (size_t)nd.x;
    #pragma unroll
    for (int p = 0; p < 4; p++) cp16(cx.sbase + RS_OFF_REC + p * 512, M + p * cx.PL, cx.pol_stream);
}
// velocities (and, for boxes, rotation + inverse inertia) of the node's dynamic bodies
__device__ __forceinline__ void issue_bodies(const BbxDev& dv, const ReplayCtx& cx, int4 nd)
{
    #pragma unroll
    for (int side = 0; side < 2; side++) {
        const int g = side ? nd.w : nd.z;
        if (g < 0) continue;
        const unsigned int s = cx.sbase + RS_OFF_BODY + side * 2048;
        const float4* v = dv.vel + 2 * (size_t)g;
        cp16(s, v, cx.pol_keep); cp16(s + 512, v + 1, cx.pol_keep);
        if (g >= dv.NS) {
            const float4* q = dv.boxq + 2 * (size_t)(g - dv.NS);
            cp16(s + 1024, q, cx.pol_keep); cp16(s + 1536, q + 1, cx.pol_keep);
        }
    }
}
__device__ __forceinline__ void read_body(const BbxDev& dv, const ReplayCtx& cx, int side, int g, BodyState& s)
{
    const unsigned int a = cx.sbase + RS_OFF_BODY + side * 2048;
    const float4 l = lds128(a), w = lds128(a + 512);
    s.v = xyz(l); s.w = xyz(w); s.im = l.w; s.ii = w.w;
    s.box = g >= dv.NS;
    if (s.box) {
        const float4 q = lds128(a + 1024), iI = lds128(a + 1536);
        s.q = q; s.iI = xyz(iI);
        const float3 u = f3(q.x, q.y, q.z);
        s.k1 = q.w * q.w - v3dot(u, u); s.k2 = 2.0f * q.w;
    }
}

__global__ void __launch_bounds__(RS_THREADS, RS_BLOCKS_PER_SM) k_replay_staged(BbxDev dv, SolveParams sp, ReplayCost rc)
{
    const int* rs_cost = rc.c;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const int n_levels = dv.ctr->n_levels;
    if (dv.ctr->replay_done || n_levels <= 0 || n_levels > RS_MAXLEV) return;      // uniform over the grid
    int* t_cnt = reinterpret_cast<int*>(rs_smem);
    int* t_head = t_cnt + RS_MAXLEV * BBX_NCLS;
    int* t_q = t_head + RS_MAXLEV * BBX_NCLS;
    int* t_b = t_q + RS_MAXLEV * BBX_NCLS;
    int* t_nw = t_b + RS_MAXLEV * BBX_NCLS;
    unsigned char* stage = rs_smem + RS_TABLE_BYTES;
    const int W = gridDim.x * RS_WARPS;
    for (int i = threadIdx.x; i < n_levels * BBX_NCLS; i += blockDim.x) t_cnt[i] = __ldcg(&dv.level_count[i]);
    __syncthreads();
    if (threadIdx.x < BBX_NCLS) {
        int run = 0;
        for (int l = 0; l < n_levels; l++) { t_head[l * BBX_NCLS + threadIdx.x] = run; run += t_cnt[l * BBX_NCLS + threadIdx.x]; }
    }
    // packing of every level: smallest per-warp budget T for which the chunks fit on W warps
    for (int l = threadIdx.x; l < n_levels; l += blockDim.x) {
        int nch[BBX_NCLS], hi = 1;
        for (int k = 0; k < BBX_NCLS; k++) { nch[k] = (t_cnt[l * BBX_NCLS + k] + 31) >> 5; hi = max(hi, nch[k] * rs_cost[k]); }
        int lo = 1;                                               // invariant: budget hi fits (every class on one warp: <= 8 warps)
        while (lo < hi) {
            const int T = (lo + hi) >> 1;
            int bins = 0;
            for (int k = 0; k < BBX_NCLS; k++) { const int q = max(1, T / rs_cost[k]); bins += (nch[k] + q - 1) / q; }
            if (bins <= W) hi = T; else lo = T + 1;
        }
        int run = 0;
        for (int k = BBX_NCLS - 1; k >= 0; k--) {
            const int q = max(1, hi / rs_cost[k]);
            t_q[l * BBX_NCLS + k] = q; t_b[l * BBX_NCLS + k] = run;
            run += (nch[k] + q - 1) / q;
        }
        t_nw[l] = run;
    }
    __syncthreads();

    const bool big = dv.ctr->n_contacts > BBX_L2_STREAM_CONTACTS;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // consecutive warp indices go round-robin over the BLOCKS: every SM gets the same mix of classes, and a thin
    // level is spread over all SMs
    const int w = wib * gridDim.x + blockIdx.x;
    ReplayCtx cx;
    cx.t_cnt = t_cnt; cx.t_head = t_head; cx.t_q = t_q; cx.t_b = t_b; cx.t_nw = t_nw;
    cx.sbase = smem_u32(stage) + wib * RS_WARP_BYTES + lane * 16;
    cx.pol_keep = big ? l2_policy_keep() : l2_policy_normal();
    cx.pol_stream = big ? l2_policy_stream() : l2_policy_normal();
    cx.PL = (size_t)dv.cap_contacts;
    const int I = sp.iterations;

    // (nit, nl): the next (iteration, level) in which this warp owns chunks; the first two descriptors are staged as
    // soon as the warp is done with its previous level, the first record just before the barrier the warp arrives at
    int nit = 1, nl = 0;
    if (I > 1) next_work(cx, n_levels, I, w, lane, 1, -1, nit, nl); else nit = I;
    bool staged_rec = false;
    WarpWork nx;
    nx.p = NULL; nx.nc = 0; nx.left = 0; nx.stride = 1; nx.cls = 0;
    if (nit < I) {
        nx = level_work(dv, cx, nl, w, lane);
        issue_desc(cx, nx, 0, 0);
        if (nx.nc > 1) issue_desc(cx, nx, 1, 1);
        cp_commit();
    }

    for (int it = 1; it < I; it++) {
        for (int l = 0; l < n_levels; l++) {
            const bool trace = sp.gather == l + 1 && it == 1 && lane == 0 && w < BBX_TRACE_WARPS;
            unsigned long long* tr = dv.level_ns + 4104 + 4 * (size_t)w;
            if (trace) { tr[0] = gtime_ns(); tr[1] = 0; tr[2] = 0; tr[3] = 0; }
            if (nit == it && nl == l) {
                const WarpWork ww = nx;
                cp_wait_all();
                bool live = ww.left > 0;
                int4 nd = make_int4(0, 0, -1, -1);
                if (live) {
                    nd = lds128i(cx.sbase + RS_OFF_DESC);
                    if (!staged_rec) issue_rec0(dv, cx, nd);
                    issue_bodies(dv, cx, nd);                          // after the barrier: the previous level's results
                }
                cp_commit();
                int cost = 0;
                for (int r = 0; r < ww.nc; r++) {
                    cp_wait_all();
                    if (trace && r == 0) tr[1] = gtime_ns();
                    // ---- this chunk: staging slots -> registers
                    const bool hasA = live && nd.z >= 0, hasB = live && nd.w >= 0;
                    BodyState A = body_state_zero(), B = body_state_zero();       // (absent bodies must hold defined values for the branch-free sweep)
                    float4 c0, c1, c2, c3;
                    if (hasA) read_body(dv, cx, 0, nd.z, A);
                    if (hasB) read_body(dv, cx, 1, nd.w, B);
                    if (live) {
                        c0 = lds128(cx.sbase + RS_OFF_REC); c1 = lds128(cx.sbase + RS_OFF_REC + 512);
                        c2 = lds128(cx.sbase + RS_OFF_REC + 1024); c3 = lds128(cx.sbase + RS_OFF_REC + 1536);
                    }
                    // ---- next chunk of this level, or the head of the next level with work: start its loads
                    bool live_n = false;
                    int4 nd_n = make_int4(0, 0, -1, -1);
                    if (r + 1 < ww.nc) {
                        live_n = (r + 1) * 32 < ww.left;
                        if (live_n) {
                            nd_n = lds128i(cx.sbase + RS_OFF_DESC + ((r + 1) & 1) * 512);
                            issue_rec0(dv, cx, nd_n);
                            issue_bodies(dv, cx, nd_n);
                        }
                        if (r + 2 < ww.nc) issue_desc(cx, ww, r + 2, r & 1);
                    } else {
                        next_work(cx, n_levels, I, w, lane, it, l, nit, nl);
                        if (nit < I) {
                            nx = level_work(dv, cx, nl, w, lane);
                            issue_desc(cx, nx, 0, 0);
                            if (nx.nc > 1) issue_desc(cx, nx, 1, 1);
                        }
                    }
                    cp_commit();
                    // ---- sweep the node's constraints in reference order (:1414-1488)
                    if (live) {
                        float4* M = dv.M + (size_t)nd.x;
                        float4 n0, n1, n2, n3;
                        for (int j = 0; j < nd.y; j++, M += ww.stride) {
                            if (j + 1 < nd.y) {
                                const float4* Mn = M + ww.stride;
                                n0 = ld128_cg_pol(Mn, cx.pol_stream); n1 = ld128_cg_pol(Mn + cx.PL, cx.pol_stream);
                                n2 = ld128_cg_pol(Mn + 2 * cx.PL, cx.pol_stream); n3 = ld128_cg_pol(Mn + 3 * cx.PL, cx.pol_stream);
                            }
                            // class 0 = single constraints between spheres (and statics): the plain form skips the box
                            // arithmetic for the whole warp; every other class uses the branch-free form (bbx_solve.cuh)
                            if (ww.cls == 0) solve_constraint<false>(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                            else             solve_constraint<true>(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                            st128_cg_pol(M + 3 * cx.PL, c3, cx.pol_stream);
                            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
                        }
                        if (hasA) store_body_keep(dv, nd.z, A, cx.pol_keep);
                        if (hasB) store_body_keep(dv, nd.w, B, cx.pol_keep);
                    }
                    cost += nd.y;
                    live = live_n; nd = nd_n;
                }
                if (trace) { tr[2] = gtime_ns(); tr[3] = ((unsigned long long)ww.cls << 32) | ((unsigned long long)cost << 16) | (unsigned long long)ww.nc; }
                // the warp's next head chunk: its descriptor has landed during the last sweep; start its first record
                staged_rec = false;
                if (nit < I) {
                    cp_wait_all();
                    if (nx.left > 0) issue_rec0(dv, cx, lds128i(cx.sbase + RS_OFF_DESC));
                    cp_commit();
                    staged_rec = true;
                }
            }
            grid.sync();
            if (it == 1 && blockIdx.x == 0 && threadIdx.x == 0 && l < 4096) {
                unsigned long long ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
                dv.level_ns[l + 1] = ns;
                if (l == 0) dv.level_ns[0] = 0;
            }
        }
    }
}

int bbx_replay_staged(BbxDev* d, const SolveParams& sp_in)
{
    const int smem = RS_TABLE_BYTES + RS_WARPS * RS_WARP_BYTES;
    if (d->replay_blocks == 0) {
        int per_sm = 0;
        BBX_CUDA(d, cudaFuncSetAttribute(k_replay_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_replay_staged, RS_THREADS, smem));
        if (per_sm < 1) return bbx_fail(d, BBX_ERR_NO_DEVICE, "staged replay kernel does not fit on an SM", __FILE__, __LINE__);
        if (per_sm > RS_BLOCKS_PER_SM) per_sm = RS_BLOCKS_PER_SM;
        d->replay_blocks = per_sm * d->num_sms;
        if (d->tune_solve_blocks > 0 && d->tune_solve_blocks < d->replay_blocks) d->replay_blocks = d->tune_solve_blocks;
    }
    SolveParams sp = sp_in;
    BbxDev dv = *d;
    static ReplayCost rc;                                            // read once per process
    static int rc_set = 0;
    if (!rc_set) {
        for (int k = 0; k < BBX_NCLS; k++) rc.c[k] = rs_cost_default[k];
        const char* e = getenv("BBX_RS_COST");
        if (e) { int v[BBX_NCLS]; if (sscanf(e, "%d,%d,%d,%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]) == BBX_NCLS) for (int k = 0; k < BBX_NCLS; k++) if (v[k] > 0) rc.c[k] = v[k]; }
        rc_set = 1;
    }
    void* args[] = { (void*)&dv, (void*)&sp, (void*)&rc };
    BBX_CUDA(d, cudaLaunchCooperativeKernel((void*)k_replay_staged, dim3(d->replay_blocks), dim3(RS_THREADS), args, smem, d->stream));
    d->launches++;
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}
