// bbx_solver_flow.cu -- the alternative EXACT schedule (BBX_SOLVER_EXACT_FLOW): barrier-free dataflow solver.
// Kept as the one independent second implementation of the reference's per-body constraint order (the GPU suite
// checks it against the level schedule at full size); measured slower than the level schedule (profiles/r01_notes.md).
#include <stdlib.h>
#include <cooperative_groups.h>
#include "bbx_solve.cuh"
namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------
// FLOW SOLVER (single large world): the same exact schedule without levels and without barriers
//
// Every constraint is its own node.  A node may run as soon as the previous constraint in the chain
// of each of its two dynamic bodies has run (chains are in the reference's emission order, so the
// result is the sequential sweep's, bit for bit).  Body velocities travel from one constraint to the
// next through BbxDev::flow: 8-byte (value, version) words written with one vector store and polled
// with one vector load, version = number of visits the body has received so far, so a consumer needs
// neither a flag round trip nor a fence.  The expected version of constraint k (chain position pos,
// chain length len) in iteration it is it*len + pos: iterations overlap wherever the graph allows it.
//
// Iteration 0 discovers a topological order on the fly (Kahn): a constraint whose predecessors have
// run is appended to a ready queue, warps take the queue in 32-slot chunks dealt round-robin, lanes
// process their slot as soon as it is filled.  The queue order is kept (fnd, M): iterations 1..I-1
// walk the same chunks with the same warp -> lane assignment, so impulses accumulate thread-privately.
// In that order every constraint sits one whole frontier behind its predecessors, which keeps the
// waits short; the warp that owns the oldest unfinished chunk never waits for anything unfinished
// and lanes never wait for one another (ready lanes run while the others keep polling), so the
// schedule cannot deadlock.  One grid barrier separates iteration 0 from the replay (fnd and M are
// written by the slot's taker of iteration 0 and read by its owner afterwards).
// ---------------------------------------------------------------------------
#define FLOW_THREADS 256
#define FLOW_BLOCKS_PER_SM 2
#define FLOW_SPIN_LIMIT (1 << 17)      // passes over one chunk before giving up (never reached unless the schedule is corrupt)

// constraints with a dynamic body are counted; those without a predecessor seed the ready queue
__global__ void __launch_bounds__(256) k_frontier_flow(BbxDev dv)
{
    __shared__ int s_solver;
    if (threadIdx.x == 0) s_solver = 0;
    __syncthreads();
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    int n_round = (n + 31) & ~31;
    int mine = 0;
    if (blockIdx.x == 0 && threadIdx.x < 16) dv.level_ns[FLOW_DBG - 8 + threadIdx.x] = 0ull;     // debug trace of the last step
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_round; c += gridDim.x * blockDim.x) {
        bool seed = false;
        if (c < n) {
            if (dv.node[2 * (size_t)c].x >= 0 || dv.node[2 * (size_t)c + 1].x >= 0) { mine++; seed = dv.indeg[c] == 0; }
        }
        int slot = warp_append(&dv.ctr->fq_tail, seed);
        if (seed) dv.fq[slot] = c;
    }
    if (mine) atomicAdd(&s_solver, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_solver) { atomicAdd(&dv.ctr->n_solver, s_solver); atomicAdd(&dv.ctr->n_nodes, s_solver); }
}

// accumulated impulses of queue slot s: (jn,tag)(jt0,tag)(jt1,tag)(mT1,tag), tag = iterations applied so far
__device__ __forceinline__ void flow_store_imp(const BbxDev& dv, int s, float4 c3, int tag)
{
    unsigned long long* p = reinterpret_cast<unsigned long long*>(dv.fimp + 2 * (size_t)s);
    asm volatile("st.global.cg.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(ll_pack(c3.x, tag)), "l"(ll_pack(c3.y, tag)),
                 "l"(ll_pack(c3.z, tag)), "l"(ll_pack(c3.w, tag)) : "memory");
}
__device__ __forceinline__ bool flow_try_load_imp(const BbxDev& dv, int s, int tag, float4& c3)
{
    const unsigned long long* p = reinterpret_cast<const unsigned long long*>(dv.fimp + 2 * (size_t)s);
    unsigned long long x0, x1, x2, x3;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x0), "=l"(x1), "=l"(x2), "=l"(x3) : "l"(p) : "memory");
    const unsigned int t = (unsigned int)tag;
    c3 = make_float4(__uint_as_float((unsigned int)x0), __uint_as_float((unsigned int)x1), __uint_as_float((unsigned int)x2),
                     __uint_as_float((unsigned int)x3));
    return ((unsigned int)(x0 >> 32) == t) & ((unsigned int)(x1 >> 32) == t) & ((unsigned int)(x2 >> 32) == t) & ((unsigned int)(x3 >> 32) == t);
}

__device__ __forceinline__ float3 shfl_pair(unsigned int mask, float3 v)
{
    return f3(__shfl_xor_sync(mask, v.x, 1), __shfl_xor_sync(mask, v.y, 1), __shfl_xor_sync(mask, v.z, 1));
}

// One Gauss-Seidel visit of one constraint (:1414-1488) by a PAIR of lanes: the even lane holds body A, the odd lane
// body B; each lane evaluates its own body's velocity at the contact and applies the impulses to its own body, the
// scalar part (relative velocity, impulse magnitudes, clamps) is computed redundantly and identically on both lanes.
// Same operations in the same order as solve_constraint, per body.  `mask` = lanes of all pairs running this pass.
__device__ __forceinline__ void solve_half(unsigned int mask, bool side, bool has, float4 c0, float3 r, float mN, float mT0,
                                           float4& c3, BodyState& S, const SolveParams& sp)
{
    const float3 n = xyz(c0);
    const float bias = c0.w, mT1 = c3.w;
    const float3 zero = f3(0.0f, 0.0f, 0.0f);
    float3 vS = has ? v3add(S.v, v3cross(S.w, r)) : zero;                    // :1233-1261
    float3 vO = shfl_pair(mask, vS);
    float3 rel = side ? v3sub(vS, vO) : v3sub(vO, vS);                       // vB - vA
    float vn = v3dot(rel, n);
    float vbias = 0.0f;
    if (vn < -sp.vel_threshold) vbias = -sp.restitution * vn;                // :1443-1445
    float jn = mN * (-(vn + vbias) + bias);                                  // :1447
    float old = c3.x;
    c3.x = bb_fmaxf(0.0f, c3.x + jn);                                        // :1451
    jn = c3.x - old;
    float3 P = v3scale(n, jn);
    float3 J = side ? P : v3scale(P, -1.0f);                                 // A receives -P, B receives P (:1456-1459)
    if (has) apply_impulse(S, J, v3cross(r, J));
    if (sp.friction > 0.0f) {                                                // :1462-1487
        vS = has ? v3add(S.v, v3cross(S.w, r)) : zero;
        vO = shfl_pair(mask, vS);
        rel = side ? v3sub(vS, vO) : v3sub(vO, vS);
        float3 t1, t2;
        tangents_of(n, t1, t2);
        float maxF = sp.friction * c3.x;
        {
            float vt = v3dot(rel, t1);
            float jt = mT0 * (-vt);
            float o = c3.y;
            c3.y = bb_fmaxf(-maxF, bb_fminf(maxF, c3.y + jt));
            jt = c3.y - o;
            float3 T = v3scale(t1, jt);
            float3 JT = side ? T : v3scale(T, -1.0f);
            if (has) apply_impulse(S, JT, v3cross(r, JT));
        }
        {
            float vt = v3dot(rel, t2);                                       // same pre-loop rel (:1466 is outside the j loop)
            float jt = mT1 * (-vt);
            float o = c3.z;
            c3.z = bb_fmaxf(-maxF, bb_fminf(maxF, c3.z + jt));
            jt = c3.z - o;
            float3 T = v3scale(t2, jt);
            float3 JT = side ? T : v3scale(T, -1.0f);
            if (has) apply_impulse(S, JT, v3cross(r, JT));
        }
    }
}

__global__ void __launch_bounds__(FLOW_THREADS, FLOW_BLOCKS_PER_SM) k_solve_flow(BbxDev dv, SolveParams sp)
{
    cg::grid_group grid = cg::this_grid();
    const int n = dv.ctr->n_solver;                                     // scheduled constraints = queue length
    const int n_chunks = (n + 15) >> 4;                                 // 16 constraints per warp: one lane per body
    const int lane = threadIdx.x & 31;
    const bool side = lane & 1;                                         // 0: body A, 1: body B
    const size_t PL = (size_t)dv.cap_contacts;
    const int last_it = sp.iterations - 1;
    int depth_max = 0, udepth_max = 0;
    bool dead = false;
    unsigned long long dbg_chunks = 0, dbg_proc = 0, dbg_spin = 0, dbg_fail = 0, dbg_wait = 0;
    const long long t_begin = clock64();

    // phase 0 = iteration 0 (the ready queue is being built), phase 1 = iterations 1..I-1 (one ticket space: chunks of
    // iteration i+1 follow those of iteration i, so a free warp always takes the oldest chunk nobody has started)
    for (int phase = 0; phase < 2; phase++) {
        if (phase == 1) {
            if (sp.iterations <= 1) break;
            grid.sync();                  // fnd, M and the impulse records of iteration 0 are complete and visible
            if (dead || (__ldcg(&dv.ctr->overflow) & 64)) break;          // (every thread reaches the barrier, also after a failure)
            if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns)); dv.level_ns[FLOW_DBG + 6] = ns; }
        }
        const bool first = phase == 0;
        int* counter = first ? &dv.ctr->fq_head : &dv.ctr->fq_ticket;
        const long long limit = first ? (long long)n_chunks : (long long)(sp.iterations - 1) * n_chunks;
        // tickets are requested two chunks ahead, the descriptor of the next chunk one chunk ahead (replay), so that
        // neither the atomic nor the descriptor's DRAM latency sits between two chunks
        int ticket = 0, ticket2 = 0;
        if (lane == 0) { ticket = atomicAdd(counter, 1); ticket2 = atomicAdd(counter, 1); }
        int4 d_next = make_int4(-1, 0, 0, 0);
        {
            const int t0 = __shfl_sync(0xffffffffu, ticket, 0);
            if (!first && (long long)t0 < limit) {
                const int it0 = 1 + t0 / n_chunks, s0 = (t0 - (it0 - 1) * n_chunks) * 16 + (lane >> 1);
                if (s0 < n) d_next = __ldcs(&dv.fnd[2 * (size_t)s0 + side]);
            }
        }
        while (true) {
            const int t = __shfl_sync(0xffffffffu, ticket, 0);
            if ((long long)t >= limit || dead) break;
            const int tn = __shfl_sync(0xffffffffu, ticket2, 0);
            if (lane == 0) { ticket = ticket2; ticket2 = atomicAdd(counter, 1); }
            const int it = first ? 0 : 1 + t / n_chunks;
            const int g = first ? t : t - (it - 1) * n_chunks;
            const int s = g * 16 + (lane >> 1);
            bool pending = s < n;
            const int4 d_cur = d_next;
            if (!first && (long long)tn < limit) {
                const int itn = 1 + tn / n_chunks, sn = (tn - (itn - 1) * n_chunks) * 16 + (lane >> 1);
                if (sn < n) d_next = __ldcs(&dv.fnd[2 * (size_t)sn + side]);
            }
            bool have = false;                    // descriptor and record loaded
            int c = -1, gS = -1, need = 0, len = 0, succ = -1, dp = 0;
            float4 c0, c3;
            float3 r = f3(0.0f, 0.0f, 0.0f);
            float mS = 0.0f;                      // even lane: normal mass, odd lane: first tangent mass
            BodyState S;
            S.box = false;
            bool ok = false, okC = first;
            float4* M = dv.M + (size_t)s;
            if (!first && pending) {
                // replay: descriptor and record in queue order (coalesced), requested before the wait
                const int4 d0 = d_cur;
                gS = d0.x; len = d0.z; need = it * len + d0.y;
                c0 = __ldcs(M);
                float4 c1 = __ldcs(M + (side ? 2 * PL : PL));
                r = xyz(c1); mS = c1.w;
                if (gS >= 0) flow_load_box(dv, gS, S);
                ok = gS < 0;
                have = true;
            }
            int passes = 0, gathered = 0;
            long long t_wait = 0;
            while (__any_sync(0xffffffffu, pending)) {
                bool ready = false;
                if (pending) {
                    if (!have) {
                        // iteration 0: the slot is filled when the constraint's last predecessor has run
                        c = __ldcg(&dv.fq[s]);
                        if (c >= 0) {
                            int4 nd = __ldg(&dv.node[2 * (size_t)c + side]);
                            gS = nd.x; need = nd.y; len = nd.z; succ = nd.w;
                            const float4* L = dv.L + 4 * (size_t)c;
                            c0 = __ldg(L);
                            float4 c1 = __ldg(L + (side ? 2 : 1));
                            c3 = __ldg(L + 3);
                            r = xyz(c1); mS = c1.w;
                            if (gS >= 0) flow_load_box(dv, gS, S);
                            ok = gS < 0;
                            have = true;
                        }
                    }
                    if (have) {
                        if (!ok) ok = flow_try_load(dv, gS, need, S, dp);
                        if (!okC) okC = flow_try_load_imp(dv, s, it, c3);
                        ready = ok && okC;
                    }
                    if (!ready) dbg_fail++;
                }
                const bool other_ready = __shfl_xor_sync(0xffffffffu, ready, 1);      // (evaluated by every lane: no short circuit)
                ready = ready && other_ready;                                          // a constraint runs when both its bodies are there
                const unsigned int rdy = __ballot_sync(0xffffffffu, ready);
                const unsigned int pnd = __ballot_sync(0xffffffffu, pending);
                // stragglers: poll a few more times before running a partial warp (a partial pass costs a full one)
                if (rdy != 0u && rdy != pnd && gathered < sp.gather) { gathered++; dbg_spin++; continue; }
                if (rdy == 0u) {
                    if (t_wait == 0) t_wait = clock64();
                    ++passes; dbg_spin++;
                    // a corrupt schedule must not hang the device: give up, tell everybody else, report through the counters
                    if (passes > FLOW_SPIN_LIMIT) { if (pending) atomicOr(&dv.ctr->overflow, 64); pending = false; dead = true; }
                    else if ((passes & 1023) == 0 && (__ldcg(&dv.ctr->overflow) & 64)) { pending = false; dead = true; }
                    __nanosleep(64);
                    continue;
                }
                dbg_proc++;
                if (t_wait) { dbg_wait += clock64() - t_wait; t_wait = 0; }
                // the pairs that are ready run now; the others poll again in the next pass
                if (ready) {
                    const bool has = gS >= 0;
                    const float mN = __shfl_sync(rdy, mS, lane & ~1), mT0 = __shfl_sync(rdy, mS, lane | 1);
                    solve_half(rdy, side, has, c0, r, mN, mT0, c3, S, sp);
                    const int depth = max(dp, __shfl_xor_sync(rdy, dp, 1)) + 1;
                    udepth_max = max(udepth_max, depth);                       // depth of the unrolled (all-iterations) graph
                    if (!side) flow_store_imp(dv, s, c3, it + 1);
                    if (first) {
                        // keep the visiting order: record and descriptor in queue order for the replay and the export
                        if (!side) { __stcg(M, c0); dv.slot_of[c] = s; dv.fq[s] = -1; }     // the queue is empty again for the next step
                        __stcg(M + (side ? 2 * PL : PL), make_float4(r.x, r.y, r.z, mS));
                        __stcg(&dv.fnd[2 * (size_t)s + side], make_int4(gS, need, len, c));
                        depth_max = max(depth_max, depth);
                    }
                    if (has) {
                        if (it == last_it && need == (it + 1) * len - 1) store_body(dv, gS, S);       // the body's final visit
                        else flow_store(dv, gS, S, need + 1, depth);
                    }
                    if (first) {
                        // a successor whose other predecessor has already run is ready now
                        bool push = succ >= 0 && atomicSub(&dv.indeg[succ], 1) == 1;
                        int q = warp_append(&dv.ctr->fq_tail, push);
                        if (push) __stcg(&dv.fq[q], succ);
                    }
                    pending = false;
                }
            }
            dbg_chunks++;
            dead = __any_sync(0xffffffffu, dead);
        }
    }
    // statistics: dependency depth (reported as "levels"), trace
    for (int o = 16; o > 0; o >>= 1) depth_max = max(depth_max, __shfl_xor_sync(0xffffffffu, depth_max, o));
    for (int o = 16; o > 0; o >>= 1) udepth_max = max(udepth_max, __shfl_xor_sync(0xffffffffu, udepth_max, o));
    if (lane == 0) {
        if (udepth_max > 0) atomicMax(&dv.level_ns[FLOW_DBG - 8], (unsigned long long)udepth_max);   // debug trace (tools/profile_step.py)
        if (depth_max > 0) atomicMax(&dv.ctr->n_levels, depth_max);
        atomicAdd(&dv.level_ns[FLOW_DBG + 0], dbg_chunks); atomicAdd(&dv.level_ns[FLOW_DBG + 1], dbg_proc);
        atomicAdd(&dv.level_ns[FLOW_DBG + 2], dbg_spin); atomicAdd(&dv.level_ns[FLOW_DBG + 4], dbg_wait);
        atomicAdd(&dv.level_ns[FLOW_DBG + 5], (unsigned long long)(clock64() - t_begin));
        unsigned long long ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
        atomicMax(&dv.level_ns[FLOW_DBG + 7], ns);
    }
    atomicAdd(&dv.level_ns[FLOW_DBG + 3], dbg_fail);
}

// host side: frontier + the cooperative flow kernel (the caller has built the per-contact chains with flow records)
int bbx_flow_solve(BbxDev* d, const BbxStepParams* p, int iterations)
{
    const int G = d->num_sms * 8;
    BBX_LAUNCH(d, k_frontier_flow, G, 256, 0, *d);
    if (d->flow_blocks == 0) {
        int per_sm = 0, coop = 0;
        BBX_CUDA(d, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d->device));
        if (!coop) return bbx_fail(d, BBX_ERR_NO_DEVICE, "device lacks cooperative launch", __FILE__, __LINE__);
        BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_flow, FLOW_THREADS, 0));
        if (per_sm < 1) return bbx_fail(d, BBX_ERR_NO_DEVICE, "flow solver kernel does not fit on an SM", __FILE__, __LINE__);
        if (per_sm > FLOW_BLOCKS_PER_SM) per_sm = FLOW_BLOCKS_PER_SM;
        d->flow_blocks = per_sm * d->num_sms;
        if (d->tune_flow_blocks > 0 && d->tune_flow_blocks < d->flow_blocks) d->flow_blocks = d->tune_flow_blocks;
    }
    bbx_timing_mark(d, 4);
    SolveParams spf;
    spf.restitution = p->settings.restitution; spf.friction = p->settings.friction;
    spf.vel_threshold = p->settings.velocity_threshold; spf.iterations = iterations;
    spf.gather = d->tune_flow_gather; spf.first_only = 0;
    BbxDev dvf = *d;
    void* fargs[] = { (void*)&dvf, (void*)&spf };
    // cooperative launch: every block must be resident (the waits rely on it) and iteration 0 ends with a grid barrier
    BBX_CUDA(d, cudaLaunchCooperativeKernel((void*)k_solve_flow, dim3(d->flow_blocks), dim3(FLOW_THREADS), fargs, 0, d->stream));
    d->launches++;
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}
