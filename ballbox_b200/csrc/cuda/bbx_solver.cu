// bbx_solver.cu -- K9-K11: presolve, dependency graph and the level-scheduled
// Gauss-Seidel velocity solver.
//
// Reference: GenerateContactConstraints ballbox.h:1176-1211, PreSolveConstraints
// :1295-1356, SolveVelocityConstraints :1410-1489, ApplyConstraintImpulse
// :1359-1408, GetConstraintBodyVelocityAtPoint :1233-1261.
//
// EXACT SCHEDULE.  The reference sweeps the constraint array sequentially, so
// each dynamic body sees its constraints in emission order: (phase, partner
// index, manifold point), see SURVEY.md F4 and section 8a row 3.  Two constraints
// that share no dynamic body commute bit-for-bit.  We therefore give every
// constraint the level 1 + max(level of the previous constraint of body A, of
// body B) and run the levels in order: inside a level no two constraints touch
// the same dynamic body, so the level can run in parallel without atomics and
// the result is bit-identical to the sequential sweep.  Static bodies and the SDF
// never receive writes (:1361, :1379, :1407) and are ignored by the graph.
//
// NODES.  All contacts between the same two bodies (a box-box manifold of up to 8
// points, the up to 20 SDF probes of one box) are adjacent in both bodies' orders,
// so they form one scheduling node: one thread loads the two bodies once, sweeps
// the node's constraints in reference order in registers, and stores the bodies
// once.  This divides the graph depth and the body gather traffic by the manifold
// size without changing a single operation of the sweep.
//
// Levels are discovered on the fly during iteration 0 (Kahn's algorithm on the
// per-body successor links); the visiting order is recorded and the constraint
// records are rewritten in that level-major order (64-byte records, a node's
// constraints contiguous), so iterations 1..I-1 stream them.  One cooperative
// persistent kernel runs all iterations; levels are separated by grid-wide
// barriers, not kernel launches.
#include <stdlib.h>
#include <cooperative_groups.h>
#include "bbx_dev.cuh"
#include "bbx_solve.cuh"
namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------
// K9 presolve: contact -> 64-byte solver record (contact order); node heads also
// get their graph fields
// ---------------------------------------------------------------------------
// (inv_mass_sum: bbx_solve.cuh)

// mode 0: PhysicsStep (everything computed, impulses zeroed)            -- :1176-1211 + :1295-1356
// mode 1: PreSolveConstraints on an imported list (impulses kept)         -- :1295-1356
// mode 2: SolveVelocityConstraints on an imported list: only n, rA, rB are derived; effective
//         masses, bias and accumulated impulses are the ones found in the host records
// per_contact: every contact is its own scheduling node (flow solver), one half per body:
// node[2c] = (dynamic gid A or -1, position in A's chain, length of A's chain, successor on A), node[2c+1] = same for B
// per-body gather record of k_presolve: position, inverse mass, inverse inertia diagonal, flags
__global__ void __launch_bounds__(256) k_body_records(BbxDev dv)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= dv.N) return;
    const float4 p = dv.pos[g];
    const F8 vw = ld256_cg(&dv.vel[2 * g]);
    float3 iI;
    if (g < dv.NS) iI = f3(vw.b.w, vw.b.w, vw.b.w);
    else iI = xyz(dv.boxq[2 * (g - dv.NS) + 1]);
    st256_cg(&dv.brec[2 * (size_t)g], make_float4(p.x, p.y, p.z, vw.a.w), make_float4(iI.x, iI.y, iI.z, __int_as_float((int)dv.flags[g])));
}

// warm > 0 (extension, mode 0 only): a contact found in the previous step's table starts from warm x its final impulses
// and is flagged (c_nrm.w = 1) for k_warm_apply
// count_nodes (level path): the statistics n_solver / n_nodes are summed here (the other paths do it in their frontier kernels)
__global__ void __launch_bounds__(256) k_presolve(BbxDev dv, float baumgarte, float allowed, float dt, int mode, int per_contact, float warm,
                                                  int count_nodes)
{
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    const unsigned long long pol_links = l2_policy_normal();
    int my_solver = 0, my_nodes = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        int4 id = dv.c_ids[c];
        float4 pt4 = dv.c_pt[c];
        float3 pt = xyz(pt4), nrm = xyz(dv.c_nrm[c]);
        int a = id.x, b = id.y;
        float3 rA, rB, iA, iB; float imA, imB; bool dynA, dynB;
        {
            F8 ra = ld256_nc(&dv.brec[2 * (size_t)a]);          // (pos, 1/m) (1/I, flags): k_body_records
            dynA = !(__float_as_int(ra.b.w) & BF_STATIC);
            rA = v3sub(pt, xyz(ra.a));
            imA = ra.a.w;                           // 0 for static bodies (:1266)
            iA = xyz(ra.b);
        }
        if (b >= 0) {
            F8 rb = ld256_nc(&dv.brec[2 * (size_t)b]);
            dynB = !(__float_as_int(rb.b.w) & BF_STATIC);
            rB = v3sub(pt, xyz(rb.a));
            imB = rb.a.w;
            iB = xyz(rb.b);
        } else {                                    // SDF: "position" is the contact point (:1310-1311)
            dynB = false; rB = v3sub(pt, pt); imB = 0.0f; iB = f3(0.0f, 0.0f, 0.0f);
        }
        float s = inv_mass_sum(imA, imB, rA, rB, iA, iB, nrm);
        float mN = (s > 0.0f) ? 1.0f / s : 0.0f;
        float3 t1, t2;
        tangents_of(nrm, t1, t2);
        float s1 = inv_mass_sum(imA, imB, rA, rB, iA, iB, t1);
        float s2 = inv_mass_sum(imA, imB, rA, rB, iA, iB, t2);
        float mT0 = (s1 > 0.0f) ? 1.0f / s1 : 0.0f;
        float mT1 = (s2 > 0.0f) ? 1.0f / s2 : 0.0f;
        float bias = (baumgarte / dt) * bb_fmaxf(0.0f, pt4.w - allowed);               // :1352-1354
        float4* L = dv.L + 4 * (size_t)c;
        float4 keep = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (mode != 0) keep = L[3];                                   // PhysicsStep never reads the old record
        if (mode == 2) { bias = L[0].w; mN = L[1].w; mT0 = L[2].w; mT1 = keep.w; }
        L[0] = make_float4(nrm.x, nrm.y, nrm.z, bias);
        L[1] = make_float4(rA.x, rA.y, rA.z, mN);
        L[2] = make_float4(rB.x, rB.y, rB.z, mT0);
        if (mode == 0 && warm > 0.0f) {
            const unsigned long long key = warm_key_of(a, b, RUN_K(id.z));
            unsigned int h = warm_hash(key) & dv.warm_mask;
            float4 prev = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            bool hit = false;
            for (unsigned int probe = 0; probe <= dv.warm_mask; probe++, h = (h + 1) & dv.warm_mask) {
                const unsigned long long kk = dv.warm_key[h];
                if (kk == BBX_WARM_EMPTY) break;
                if (kk == key) {
                    prev = dv.warm_val[h];
                    const int wa = a < dv.NS ? a / dv.cap_s : (a - dv.NS) / dv.cap_b;
                    hit = __float_as_int(prev.w) == dv.warm_epoch[wa];              // entries of a world that was reset since are stale
                    break;
                }
            }
            L[3] = hit ? make_float4(warm * prev.x, warm * prev.y, warm * prev.z, mT1) : make_float4(0.0f, 0.0f, 0.0f, mT1);
            reinterpret_cast<float*>(&dv.c_nrm[c])[3] = hit ? 1.0f : 0.0f;
        }
        else if (mode == 0) L[3] = make_float4(0.0f, 0.0f, 0.0f, mT1);                  // impulses start at 0 (:1203-1206)
        else                L[3] = make_float4(keep.x, keep.y, keep.z, mT1);
        if (per_contact) {
            dv.slot_of[c] = -1;                     // (level path: bbx_ensure_slot_of clears and builds the table on demand)
            dv.node[2 * (size_t)c] = make_int4(dynA ? a : -1, 0, 0, -1);
            dv.node[2 * (size_t)c + 1] = make_int4(dynB ? b : -1, 0, 0, -1);
            st64_pol(&dv.link[c], make_int2(-1, -1), pol_links);        // group solver: successors in the dense array (see below)
            dv.indeg[c] = (dynA ? 1 : 0) + (dynB ? 1 : 0);
            if (dynA) atomicAdd(&dv.deg[a], 1);
            if (dynB) atomicAdd(&dv.deg[b], 1);
        } else if (RUN_K(id.z) != 0) {
            // level path: one 16-byte node record per CONTACT (dv.node[c]); a continuation of a run gets a dummy so that the
            // warp writes whole sectors (16 bytes at a 32-byte stride, heads only, was a DRAM read-modify-write per record)
            dv.node[c] = make_int4(-1, -1, 0, 0);
        } else {                                                                        // node head
            int m = RUN_LEN(id.z);
            dv.node[c] = make_int4(dynA ? a : -1, dynB ? b : -1, m, bbx_node_class(m, a >= dv.NS || b >= dv.NS));
            // successor links live in their own dense array: k_adj_link fills them with scattered 4-byte stores, which in the
            // 32-byte node records were DRAM read-modify-writes (669 MB of traffic for 70 MB of links); at 8 bytes per
            // contact the array is small enough for those partial writes to be merged in L2 (presolve+graph 0.81 -> 0.69 ms;
            // iteration 0 pays one more sector per node, +0.06 ms; an evict-last hint on the links measured the same)
            st64_pol(&dv.link[c], make_int2(-1, -1), pol_links);
            dv.indeg[c] = (dynA ? 1 : 0) + (dynB ? 1 : 0);
            if (dynA) atomicAdd(&dv.deg[a], 1);
            if (dynB) atomicAdd(&dv.deg[b], 1);
            if (dynA || dynB) { my_solver += m; my_nodes++; }
        }
    }
    if (count_nodes) {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { my_solver += __shfl_xor_sync(0xffffffffu, my_solver, o); my_nodes += __shfl_xor_sync(0xffffffffu, my_nodes, o); }
        if ((threadIdx.x & 31) == 0 && my_nodes) { atomicAdd(&dv.ctr->n_solver, my_solver); atomicAdd(&dv.ctr->n_nodes, my_nodes); }
    }
}

// ---------------------------------------------------------------------------
// K10 dependency graph: per-body ordered node lists -> successor links
// ---------------------------------------------------------------------------
// order of a node inside ONE body's list: (phase, partner index); the low bits carry the node's
// class (they cannot affect the order: (phase, partner) is unique within a body's list)
__device__ __forceinline__ unsigned int body_order_key(const BbxDev& dv, int4 id, bool for_a, int cls)
{
    int other = for_a ? id.y : id.x;
    unsigned int partner = 0;
    if (other >= 0) partner = (unsigned int)(other < dv.NS ? other : other - dv.NS);
    return ((unsigned int)id.w << 29) | (partner << 5) | (unsigned int)cls;
}

// COLOURED schedule (BBX_SOLVER_COLOURED): the per-body order is a hash of the node identity
// instead of the emission order.  Orienting every conflict (two nodes sharing a dynamic body)
// from the smaller to the larger hash gives a DAG whose Kahn levels are independent sets, i.e.
// a graph colouring with far fewer classes than the emission-order levels; the sweep visits the
// classes in order, so no atomics are needed inside a class.  This REORDERS the Gauss-Seidel
// sweep: results are deterministic but no longer equal to the reference's.
__device__ __forceinline__ unsigned int node_hash29(int4 id)
{
    unsigned int h = (unsigned int)id.x * 0x9E3779B1u ^ ((unsigned int)(id.y + 3) * 0x85EBCA77u) ^ ((unsigned int)id.w * 0xC2B2AE3Du);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h >> 3;
}
__device__ __forceinline__ unsigned int coloured_key(int4 id, int cls) { return (node_hash29(id) << 3) | (unsigned int)cls; }

// total order used to break hash ties deterministically: (phase, A gid, B gid)
__device__ __forceinline__ bool node_before(const BbxDev& dv, int c1, int c2)
{
    int4 a = dv.c_ids[c1], b = dv.c_ids[c2];
    if (a.w != b.w) return a.w < b.w;
    if (a.x != b.x) return a.x < b.x;
    return a.y < b.y;
}

// order: 0 emission order from (phase, partner); 1 coloured (hash); 2 imported list (array index)
__global__ void __launch_bounds__(256) k_adj_fill(BbxDev dv, int order, int per_contact)
{
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        int4 id = dv.c_ids[c];
        if (per_contact) {
            // order inside a body's chain: (phase, partner index, manifold point) = the reference's emission order
            int2 l = make_int2(dv.node[2 * (size_t)c].x, dv.node[2 * (size_t)c + 1].x);
            int k = RUN_K(id.z) & 31;
            if (l.x >= 0) {
                int slot = atomicAdd(&dv.adj_cursor[l.x], 1);
                unsigned int key = order == 2 ? (unsigned int)c : (body_order_key(dv, id, true, 0) | (unsigned int)k);
                dv.adj[slot] = ((unsigned long long)key << 32) | (unsigned int)c;                       // side A
            }
            if (l.y >= 0) {
                int slot = atomicAdd(&dv.adj_cursor[l.y], 1);
                unsigned int key = order == 2 ? (unsigned int)c : (body_order_key(dv, id, false, 0) | (unsigned int)k);
                dv.adj[slot] = ((unsigned long long)key << 32) | (unsigned int)c | 0x80000000u;         // side B
            }
            continue;
        }
        if (RUN_K(id.z) != 0) continue;
        int4 l = dv.node[c];
        if (l.x >= 0) {
            int slot = atomicAdd(&dv.adj_cursor[l.x], 1);
            unsigned int key = order == 1 ? coloured_key(id, l.w) : (order == 2 ? (((unsigned int)c << 3) | (unsigned int)l.w) : body_order_key(dv, id, true, l.w));
            dv.adj[slot] = ((unsigned long long)key << 32) | (unsigned int)c;
        }
        if (l.y >= 0) {
            int slot = atomicAdd(&dv.adj_cursor[l.y], 1);
            unsigned int key = order == 1 ? coloured_key(id, l.w) : (order == 2 ? (((unsigned int)c << 3) | (unsigned int)l.w) : body_order_key(dv, id, false, l.w));
            dv.adj[slot] = ((unsigned long long)key << 32) | (unsigned int)c | 0x80000000u;             // side B
        }
    }
}

// One thread per body: sort the body's node list into chain order and write the successor links.  An entry is
// (order key << 32) | side << 31 | node (side 0: the body is A of the node, 1: B), so that no node record has to be
// read back.  Lists of up to ADJ_LOCAL entries are sorted in a thread-private array (all entries are requested at
// once); longer ones (a big dynamic body) in place in global memory.
#define ADJ_LOCAL 24

__device__ __forceinline__ bool adj_after(const BbxDev& dv, unsigned long long a, unsigned long long v)
{
    // a sorts after v?  Equal 32-bit keys can only happen in the coloured schedule (hash collision): fall back to the
    // deterministic total order so that the result never depends on the contact numbering
    return (a >> 32) > (v >> 32) || ((a >> 32) == (v >> 32) && node_before(dv, ADJ_NODE(v), ADJ_NODE(a)));
}

__device__ __forceinline__ void adj_write_links(const BbxDev& dv, int x, const unsigned long long* lst, int d, int per_contact, int flow_records)
{
    if (per_contact && flow_records) {
        // flow solver: (position, length) and successor of every constraint in this body's chain
        for (int i = 0; i < d; i++) {
            const int cur = ADJ_NODE(lst[i]);
            const int nxt = (i + 1 < d) ? ADJ_NODE(lst[i + 1]) : -1;
            dv.node[2 * (size_t)cur + ADJ_SIDE(lst[i])] = make_int4(x, i, d, nxt);
        }
    } else if (per_contact) {
        // group solver: only the successor, into the dense link array (a 16-byte half of a node record per (constraint,
        // side) was a DRAM read-modify-write each)
        for (int i = 1; i < d; i++)
            st32_pol((int*)&dv.link[ADJ_NODE(lst[i - 1])] + ADJ_SIDE(lst[i - 1]), ADJ_NODE(lst[i]), l2_policy_normal());
    } else {
        for (int i = 1; i < d; i++) {
            const unsigned long long e2 = lst[i];
            const int link = ADJ_NODE(e2) | ((int)((e2 >> 32) & 7u) << 28);          // successor id | its class
            st32_pol((int*)&dv.link[ADJ_NODE(lst[i - 1])] + ADJ_SIDE(lst[i - 1]), link, l2_policy_normal());
        }
    }
}

// keep_sorted: the sorted list is written back (k_warm_apply walks it)
// push_frontier (level path): the thread whose decrement takes a node's in-degree to zero appends it to level 0 of
// its class queue (this used to be a pass over all contacts, k_frontier0)
__global__ void __launch_bounds__(128) k_adj_link(BbxDev dv, int per_contact, int flow_records, int keep_sorted, int push_frontier)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    bool has = false;
    int ready = -1, ready_cls = -1;
    const unsigned long long pol_links = l2_policy_normal();
    if (x < dv.N) {
        int s = dv.adj_start[x], e = dv.adj_start[x + 1];
        int d = e - s;
        has = d > 0;
        if (has) {
            unsigned long long* lst = dv.adj + s;
            if (d <= 8) {
                // most lists: sorted in REGISTERS with a fixed 19-comparator network (the thread-private array of the
                // general path lives in local memory: its dependent LDL/STL chain was half of this kernel's stall samples)
                unsigned long long e0, e1, e2, e3, e4, e5, e6, e7;
                e0 = lst[0];
                e1 = d > 1 ? lst[1] : ~0ull; e2 = d > 2 ? lst[2] : ~0ull; e3 = d > 3 ? lst[3] : ~0ull;
                e4 = d > 4 ? lst[4] : ~0ull; e5 = d > 5 ? lst[5] : ~0ull; e6 = d > 6 ? lst[6] : ~0ull; e7 = d > 7 ? lst[7] : ~0ull;
#define ADJ_CE(a, b) { if (b != ~0ull && (a == ~0ull || adj_after(dv, a, b))) { unsigned long long t_ = a; a = b; b = t_; } }   /* ~0: padding, stays last */
                ADJ_CE(e0, e1) ADJ_CE(e2, e3) ADJ_CE(e4, e5) ADJ_CE(e6, e7)
                ADJ_CE(e0, e2) ADJ_CE(e1, e3) ADJ_CE(e4, e6) ADJ_CE(e5, e7)
                ADJ_CE(e1, e2) ADJ_CE(e5, e6)
                ADJ_CE(e0, e4) ADJ_CE(e1, e5) ADJ_CE(e2, e6) ADJ_CE(e3, e7)
                ADJ_CE(e2, e4) ADJ_CE(e3, e5)
                ADJ_CE(e1, e2) ADJ_CE(e3, e4) ADJ_CE(e5, e6)
#undef ADJ_CE
                if (atomicSub(&dv.indeg[ADJ_NODE(e0)], 1) == 1) { ready = ADJ_NODE(e0); ready_cls = (int)((e0 >> 32) & 7u); }
#define ADJ_LINK(i, prev, cur, nxt_valid, nxt)                                                                         \
                if (i < d) {                                                                                            \
                    if (per_contact && flow_records) dv.node[2 * (size_t)ADJ_NODE(cur) + ADJ_SIDE(cur)] = make_int4(x, i, d, (nxt_valid) ? ADJ_NODE(nxt) : -1); \
                    else if (per_contact) { if (i > 0) st32_pol((int*)&dv.link[ADJ_NODE(prev)] + ADJ_SIDE(prev), ADJ_NODE(cur), pol_links); } \
                    else if (i > 0) st32_pol((int*)&dv.link[ADJ_NODE(prev)] + ADJ_SIDE(prev), ADJ_NODE(cur) | ((int)((cur >> 32) & 7u) << 28), pol_links); \
                    if (keep_sorted) lst[i] = cur;                                                                      \
                }
                ADJ_LINK(0, e0, e0, d > 1, e1) ADJ_LINK(1, e0, e1, d > 2, e2) ADJ_LINK(2, e1, e2, d > 3, e3) ADJ_LINK(3, e2, e3, d > 4, e4)
                ADJ_LINK(4, e3, e4, d > 5, e5) ADJ_LINK(5, e4, e5, d > 6, e6) ADJ_LINK(6, e5, e6, d > 7, e7) ADJ_LINK(7, e6, e7, false, e7)
#undef ADJ_LINK
            } else if (d <= ADJ_LOCAL) {
                unsigned long long loc[ADJ_LOCAL];
                for (int i = 0; i < d; i++) loc[i] = lst[i];
                for (int i = 1; i < d; i++) {                      // insertion sort: lists are short
                    unsigned long long v = loc[i];
                    int j = i - 1;
                    while (j >= 0 && adj_after(dv, loc[j], v)) { loc[j + 1] = loc[j]; j--; }
                    loc[j + 1] = v;
                }
                // first node of this body has no predecessor here
                if (atomicSub(&dv.indeg[ADJ_NODE(loc[0])], 1) == 1) { ready = ADJ_NODE(loc[0]); ready_cls = (int)((loc[0] >> 32) & 7u); }
                adj_write_links(dv, x, loc, d, per_contact, flow_records);
                if (keep_sorted) for (int i = 0; i < d; i++) lst[i] = loc[i];
            } else {
                for (int i = 1; i < d; i++) {
                    unsigned long long v = lst[i];
                    int j = i - 1;
                    while (j >= 0 && adj_after(dv, lst[j], v)) { lst[j + 1] = lst[j]; j--; }
                    lst[j + 1] = v;
                }
                if (atomicSub(&dv.indeg[ADJ_NODE(lst[0])], 1) == 1) { ready = ADJ_NODE(lst[0]); ready_cls = (int)((lst[0] >> 32) & 7u); }
                adj_write_links(dv, x, lst, d, per_contact, flow_records);
            }
            if (per_contact && flow_records && !keep_sorted) flow_init_body(dv, x);    // the body's versioned record starts at version 0
            wake_body(dv, x);              // :1373-1376, :1402-1405: every impulse application wakes the body, even a zero one
        }
    }
    unsigned int m = __ballot_sync(0xffffffffu, has);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&dv.ctr->n_touched, __popc(m));
    if (push_frontier) {
        const bool push = ready >= 0;
        const unsigned int active = __ballot_sync(0xffffffffu, push);
        if (push) {
            const unsigned int peers = __match_any_sync(active, ready_cls);
            const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&dv.level_count[ready_cls], __popc(peers));
            base = __shfl_sync(peers, base, leader);
            dv.queue_k[ready_cls][base + __popc(peers & ((1u << lane) - 1u))] = ready;
        }
    }
}

// Warm starting (extension, ballbox_ext.h): before the first sweep every contact that was matched with the previous
// step applies its starting impulses, A receiving -X and B +X for X = n*jn, then t1*jt0 and t2*jt1 when friction > 0,
// constraint after constraint in emission order.  An application only touches the body it is applied to, so walking
// every body's own chain (sorted by k_adj_link) performs, per body, exactly the operations of that sequential pass.
__global__ void __launch_bounds__(128) k_warm_apply(BbxDev dv, int per_contact, int flow_records, float friction)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= dv.N) return;
    const int s0 = dv.adj_start[x], d = dv.adj_start[x + 1] - s0;
    if (d <= 0) return;
    BodyState S;
    load_body(dv, x, S);
    for (int i = 0; i < d; i++) {
        const unsigned long long e = dv.adj[s0 + i];
        const int c0 = ADJ_NODE(e), side = ADJ_SIDE(e);
        const int m = per_contact ? 1 : RUN_LEN(dv.c_ids[c0].z);
        for (int j = 0; j < m; j++) {
            const int c = c0 + j;
            if (dv.c_nrm[c].w == 0.0f) continue;                       // not matched: starts from zero, nothing applied
            const float4* L = dv.L + 4 * (size_t)c;
            const float4 l0 = L[0], lr = L[side ? 2 : 1], l3 = L[3];
            const float3 n = xyz(l0), r = xyz(lr);
            float3 t1, t2;
            tangents_of(n, t1, t2);
            {
                float3 P = v3scale(n, l3.x);
                if (!side) P = v3scale(P, -1.0f);
                apply_impulse(S, P, v3cross(r, P));
            }
            if (friction > 0.0f) {
                float3 T = v3scale(t1, l3.y);
                if (!side) T = v3scale(T, -1.0f);
                apply_impulse(S, T, v3cross(r, T));
                T = v3scale(t2, l3.z);
                if (!side) T = v3scale(T, -1.0f);
                apply_impulse(S, T, v3cross(r, T));
            }
        }
    }
    store_body(dv, x, S);
    if (per_contact && flow_records) flow_init_body(dv, x);
}

// 3 blocks of 256 threads per SM (80 registers).  One 768-thread block per SM was measured
// slower (6.6 vs 5.8 ms per step at 1M bodies) although it makes the barrier cheaper.
#define LC_CACHE 512
#define SOLVE_THREADS 256
#define SOLVE_BLOCKS_PER_SM 2
#define PEND_ROUNDS 4                  // rounds of a level between two block-aggregated appends
#define DEFER_MIN_CLASS 3              // node classes 3.. (>= 3 constraints per node) may be deferred when terminal

__device__ __forceinline__ unsigned long long gtime0_ns() { unsigned long long ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns)); return ns; }

// COLOUR (coloured schedule with the staged replay): every node also takes the smallest colour not yet used on its two
// bodies (the per-body masks ride along the chains: a node is the only ready node of both its bodies, so the masks need no
// atomics; greedy colouring in the order of the hash priorities, <= deg(A) + deg(B) - 1 colours, ~max degree + 1 in a pile)
// and stores it at its queue position; bbx_colour_rebucket then rewrites the level-major order colour-major
template <bool COLOUR>
__global__ void __launch_bounds__(SOLVE_THREADS, SOLVE_BLOCKS_PER_SM) k_solve_levels(BbxDev dv, SolveParams sp)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ LevelMap lm;
    __shared__ int s_pcnt[BBX_NCLS], s_pbase[BBX_NCLS];
    const bool big = dv.ctr->n_contacts > BBX_L2_STREAM_CONTACTS;
    const unsigned long long pol_keep = big ? l2_policy_keep() : l2_policy_normal(), pol_stream = big ? l2_policy_stream() : l2_policy_normal();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    // work index of this thread: consecutive 32-node chunks go round-robin over the BLOCKS, so a level
    // with few nodes is spread over all SMs instead of filling the first few blocks
    const int wtid = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
    int* lc = dv.level_count;                       // lc[level * BBX_NCLS + class]
    const size_t PL = (size_t)dv.cap_contacts;

    // ---- iteration 0: discover levels (Kahn) while solving ----
    // RECORD SLOTS.  The level-major records of level l, class k <= 4 (1, 1, 2, 3, 4 constraints per node) sit at
    //     mbase(l) + off_k(l) + j * cnt(l, k) + idx      (node idx of the class, its constraint j):
    // a function of the level's class counts alone, so no counter is touched (the single slot counter used to take one
    // same-address atomic per warp and round, ~100 us serialised per step: iteration 0 1.24 -> 1.07 ms) and the lanes
    // of a warp read every j coalesced (node-contiguous slots, idx * m_k + j, measured equal: 4.31 vs 4.28 ms).
    // Classes 5-7 (5-6, 7-8, 9+ constraints: 0.3 % of the nodes) have no common length: their runs are
    // contiguous and handed out from the END of the array by a counter.  Bottom and top together hold one slot per
    // solver constraint <= cap_contacts, so they cannot meet.
    int lvl = 0, mbase = 0;
    // thin schedules only (fewer than 8 chunks per warp and level on average would still be "fat": there a level's time is
    // its total work, not its longest node, and an extra level only costs a barrier)
    const bool defer_on = __ldcg(&dv.ctr->n_nodes) < 256 * (int)(gridDim.x * (SOLVE_THREADS / 32));
    bool injected = false;
    if (threadIdx.x == 0) level_map_load(lm, lc, true);
    __syncthreads();
    for (;;) {
        if (lm.pref[BBX_NCLS] == 0) {
            // frontier empty: the deferred terminal nodes (if any) form the last level
            // (L2 reads: the counter shares a line with fields this kernel read at its start, and L1 is not coherent)
            const int nd = (defer_on && !injected) ? __ldcg(&dv.ctr->n_deferred) : 0;
            if (nd <= 0) break;
            if (lvl + 2 >= dv.cap_levels) { if (tid == 0) atomicOr(&dv.ctr->overflow, 16); break; }
            // every block must have read this level's (zero) counts before the first count is raised: a block that
            // loaded them later would take the level for an ordinary one and skip this branch
            grid.sync();
            for (int i = tid; i < nd; i += stride) {
                const int e = __ldcg(&dv.adj_cursor[i]), kk = (e >> 28) & 7;
                dv.queue_k[kk][lm.head[kk] + atomicAdd(&lc[lvl * BBX_NCLS + kk], 1)] = e & 0x0fffffff;
            }
            injected = true;
            grid.sync();
            if (threadIdx.x == 0) level_map_load(lm, lc + lvl * BBX_NCLS, false);
            __syncthreads();
            continue;
        }
        if (lvl + 2 >= dv.cap_levels) { if (tid == 0) atomicOr(&dv.ctr->overflow, 16); break; }
        if (tid < BBX_NCLS) lc[(lvl + 2) * BBX_NCLS + tid] = 0;
        int* lc_next = lc + (lvl + 1) * BBX_NCLS;
        const int total = lm.pref[BBX_NCLS];
        // Successors that became ready are collected per thread and appended once per BLOCK (and every PEND_ROUNDS
        // rounds): the class counters of the next level are the busiest addresses of this kernel (a same-address
        // atomic with a returned value costs 0.5-0.7 ns chip-wide, 1.7 us per dependent use: tools/micro/atomic_bench.cu),
        // one atomic per block and class instead of one per warp and class takes them off the critical path.
        int pend[2 * PEND_ROUNDS], np = 0;
        const int rounds = (total + stride - 1) / stride;            // uniform over the grid: every thread takes part in the flushes
        const int w0t = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
        const bool tr0 = sp.gather == lvl + 1 && (threadIdx.x & 31) == 0 && w0t < BBX_TRACE_WARPS / 2;
        unsigned long long* trp = dv.level_ns + 4104 + 8 * (size_t)w0t;
        if (tr0) { for (int q = 0; q < 8; q++) trp[q] = 0; trp[0] = gtime0_ns(); }
        for (int r = 0; r < rounds; r++) {
            const int t = wtid + r * stride;
            int k = BBX_NCLS - 1;
            #pragma unroll
            for (int kk = BBX_NCLS - 2; kk >= 0; kk--) if (t >= lm.pref[kk]) k = kk;
            int idx = t - lm.pref[k];
            bool live = t < total && idx < lm.cnt[k];
            int p = lm.head[k] + idx;
            int sa = -1, sb = -1, m = 0, c = 0;
            int4 nd0 = make_int4(-1, -1, 0, 0);
            if (live) {
                c = __ldcg(&dv.queue_k[k][p]);
                const int4 rec = ld128i_cg_pol(&dv.node[c], pol_stream);                         // ids, run length
                const int2 lk = ld64_pol(&dv.link[c], pol_stream);                                // successors (L2 resident, last use)
                nd0 = make_int4(rec.x, rec.y, rec.z, 0);
                sa = lk.x; sb = lk.y;
                m = nd0.z;
                if (tr0 && r == 0) trp[1] = gtime0_ns() + (unsigned long long)(m & 0);      // node + links have arrived
            }
            int slot0, mstride;
            if (k <= 4) {
                const int c0n = lm.cnt[0], c1n = lm.cnt[1], c2n = lm.cnt[2], c3n = lm.cnt[3];
                const int off = k == 0 ? 0 : k == 1 ? c0n : k == 2 ? c0n + c1n : k == 3 ? c0n + c1n + 2 * c2n : c0n + c1n + 2 * c2n + 3 * c3n;
                slot0 = mbase + off + idx; mstride = lm.cnt[k];
            } else {                                                  // (warp-uniform: a chunk holds one class)
                slot0 = dv.cap_contacts - (warp_reserve(&dv.ctr->n_mslots, m) + m); mstride = 1;
            }
            if (live) {
                __stcg(&dv.Nd_k[k][p], make_int4(slot0, m, nd0.x, nd0.y));
                bool hasA = nd0.x >= 0, hasB = nd0.y >= 0;
                if (COLOUR) {
                    const unsigned long long ua = hasA ? __ldcg(&dv.col_used[nd0.x]) : 0ull, ub = hasB ? __ldcg(&dv.col_used[nd0.y]) : 0ull;
                    const unsigned long long fr = ~(ua | ub);
                    int colr = fr == 0ull ? BBX_MAX_COLOURS : __ffsll((long long)fr) - 1;
                    if (colr >= dv.colour_limit) { atomicOr(&dv.ctr->colour_overflow, 1); colr = dv.colour_limit - 1; }      // (the step replays its levels)
                    if (hasA) __stcg(&dv.col_used[nd0.x], ua | (1ull << colr));
                    if (hasB) __stcg(&dv.col_used[nd0.y], ub | (1ull << colr));
                    __stcg(&dv.colq_k[k][p], colr);
                }
                BodyState A, B;
                if (k == 0) {                                          // class 0: spheres only, plain sweep below
                    if (hasA) load_body_keep(dv, nd0.x, A, pol_keep);
                    if (hasB) load_body_keep(dv, nd0.y, B, pol_keep);
                } else load_bodies_keep(dv, nd0.x, nd0.y, hasA, hasB, A, B, pol_keep);
                // contact-order records are gathered once (two 256-bit loads, the next record in flight while
                // the current constraint is swept), then written plane-major
                const float4* L = dv.L + 4 * (size_t)c;
                F8 lo = ld256_nc_pol(L, pol_stream), hi = ld256_nc_pol(L + 2, pol_stream);
                if (tr0 && r == 0) trp[2] = gtime0_ns() + (unsigned long long)(__float_as_int(lo.a.x) & 0) + (unsigned long long)(hasA ? (__float_as_int(A.v.x) & 0) : 0);   // record + body A have arrived
                for (int j = 0; j < m; j++, L += 4) {
                    float4 c0 = lo.a, c1 = lo.b, c2 = hi.a, c3 = hi.b;
                    if (j + 1 < m) { lo = ld256_nc_pol(L + 4, pol_stream); hi = ld256_nc_pol(L + 6, pol_stream); }
                    if (k == 0) solve_constraint<false>(c0, c1, c2, c3, hasA, hasB, A, B, sp);     // class 0: no box in the warp
                    else        solve_constraint<true>(c0, c1, c2, c3, hasA, hasB, A, B, sp);      // branch-free form (bbx_solve.cuh)
                    float4* M = dv.M + (size_t)(slot0 + j * mstride);
                    st128_cg_pol(M, c0, pol_stream); st128_cg_pol(M + PL, c1, pol_stream); st128_cg_pol(M + 2 * PL, c2, pol_stream); st128_cg_pol(M + 3 * PL, c3, pol_stream);
                }
                if (hasA) store_body_keep(dv, nd0.x, A, pol_keep);    // (the bodies were woken by k_adj_link: one coalesced pass over the
                if (hasB) store_body_keep(dv, nd0.y, B, pol_keep);    //  touched bodies instead of a random flag access per node here)
                if (tr0 && r == 0) trp[3] = gtime0_ns();                                     // sweeps and stores issued
                if (sa >= 0 && atomicSub(&dv.indeg[sa & 0x0fffffff], 1) != 1) sa = -1;
                if (sb >= 0 && atomicSub(&dv.indeg[sb & 0x0fffffff], 1) != 1) sb = -1;
                if (tr0 && r == 0) trp[4] = gtime0_ns() + (unsigned long long)((sa | sb) & 0);      // in-degree atomics have returned
                if (defer_on) {
                    // DEFERRED TERMINAL NODES: a heavy node (>= 3 constraints) without successors -- typically the probes of a box
                    // on the SDF, always last in the box's chain -- may run at any later time.  In a thin schedule every level
                    // waits for its longest node, so these are collected and swept together in ONE extra level at the end.
                    if (sa >= 0 && ((sa >> 28) & 7) >= DEFER_MIN_CLASS) {
                        const int2 l2 = ld64_pol(&dv.link[sa & 0x0fffffff], pol_stream);
                        if (l2.x < 0 && l2.y < 0) { dv.adj_cursor[atomicAdd(&dv.ctr->n_deferred, 1)] = sa; sa = -1; }
                    }
                    if (sb >= 0 && ((sb >> 28) & 7) >= DEFER_MIN_CLASS) {
                        const int2 l2 = ld64_pol(&dv.link[sb & 0x0fffffff], pol_stream);
                        if (l2.x < 0 && l2.y < 0) { dv.adj_cursor[atomicAdd(&dv.ctr->n_deferred, 1)] = sb; sb = -1; }
                    }
                }
            }
            // successors that became ready join the next level of their own class
            if (sa >= 0) pend[np++] = sa;
            if (sb >= 0) pend[np++] = sb;
            if ((r + 1) % PEND_ROUNDS == 0 || r + 1 == rounds) {
                // block-aggregated append: slot inside the block through shared-memory counters, one global atomic per class
                __syncthreads();
                if (threadIdx.x < BBX_NCLS) s_pcnt[threadIdx.x] = 0;
                __syncthreads();
                int pslot[2 * PEND_ROUNDS];
                for (int i = 0; i < np; i++) pslot[i] = atomicAdd(&s_pcnt[(pend[i] >> 28) & 7], 1);
                __syncthreads();
                if (threadIdx.x < BBX_NCLS && s_pcnt[threadIdx.x] > 0) s_pbase[threadIdx.x] = atomicAdd(&lc_next[threadIdx.x], s_pcnt[threadIdx.x]);
                __syncthreads();
                for (int i = 0; i < np; i++) {
                    const int cls = (pend[i] >> 28) & 7;
                    dv.queue_k[cls][lm.head[cls] + lm.cnt[cls] + s_pbase[cls] + pslot[i]] = pend[i] & 0x0fffffff;
                }
                np = 0;
            }
        }
        if (tr0) { trp[5] = gtime0_ns(); trp[6] = (unsigned long long)rounds; }               // all rounds and flushes done
        mbase += lm.cnt[0] + lm.cnt[1] + 2 * lm.cnt[2] + 3 * lm.cnt[3] + 4 * lm.cnt[4];
        grid.sync();
        if (tr0) trp[7] = gtime0_ns();
        if (tid == 0 && lvl < 2000) {                                 // debug trace: end of level `lvl` of iteration 0
            unsigned long long ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
            dv.level_ns[2048 + lvl] = ns;
        }
        if (threadIdx.x == 0) level_map_load(lm, lc + (lvl + 1) * BBX_NCLS, false);
        __syncthreads();
        lvl++;
    }
    const int n_levels = lvl;
    // iterations 1..I-1: in k_replay_staged (first_only) unless the schedule is deeper than its tables or THIN: below ~8k
    // nodes per level (less than one chunk for every ninth warp) a level is barrier + one chunk, and the plain replay
    // loop below has the shorter path per level (100k spheres, 4.2k nodes per level: 1.28 vs 1.61 ms; at 11k nodes per
    // level the two are equal, at 85k the staged replay wins by 0.55 ms)
    // (first_only == 2, BBX_REPLAY=2: the staged replay also for thin schedules -- test coverage)
    const bool own_replay = !sp.first_only || n_levels > RS_MAXLEV ||
                            (!COLOUR && sp.first_only == 1 && __ldcg(&dv.ctr->n_nodes) < RS_MIN_NODES_PER_LEVEL * n_levels);   // (colour-major replay: a dozen levels)
    if (tid == 0) { dv.ctr->n_levels = n_levels; dv.ctr->replay_done = own_replay ? 1 : 0; }
    if (!own_replay) return;
    if (tid == 0) { unsigned long long ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns)); dv.level_ns[4096 + 1] = ns; }

    // ---- iterations 1..I-1: stream the level-major records ----
    // (prefetching the next level's descriptors/records across the barrier was tried and measured
    //  slower: profiles/r01_notes.md)
    // the per-level class counts are static from here on: keep the first LC_CACHE levels in shared
    // memory so that a level does not start with a global round trip
    __shared__ int lc_cache[LC_CACHE * BBX_NCLS];
    // ... and, when all levels fit, the queue position of every (level, class) as well: a level then starts without the
    // two block barriers and the serial section that built its map
    __shared__ int head_cache[LC_CACHE * BBX_NCLS];
    const bool tables = n_levels <= LC_CACHE;
    for (int i = threadIdx.x; i < min(n_levels, LC_CACHE) * BBX_NCLS; i += blockDim.x) lc_cache[i] = __ldcg(&lc[i]);
    __syncthreads();
    if (tables && threadIdx.x < BBX_NCLS) {
        int run = 0;
        for (int l = 0; l < n_levels; l++) { head_cache[l * BBX_NCLS + threadIdx.x] = run; run += lc_cache[l * BBX_NCLS + threadIdx.x]; }
    }
    __syncthreads();
    // (with first_only the later iterations run in k_replay_staged, bbx_replay.cu)
    for (int it = 1; it < sp.iterations; it++) {
        for (int l = 0; l < n_levels; l++) {
            if (!tables) {
                __syncthreads();
                if (threadIdx.x == 0) {
                    for (int k = 0; k < BBX_NCLS; k++) {
                        if (l == 0) lm.head[k] = 0; else lm.head[k] += lm.cnt[k];
                        lm.cnt[k] = (l < LC_CACHE) ? lc_cache[l * BBX_NCLS + k] : __ldcg(&lc[l * BBX_NCLS + k]);
                    }
                }
                __syncthreads();
            }
            const int* cnt = tables ? &lc_cache[l * BBX_NCLS] : lm.cnt;
            const int* head = tables ? &head_cache[l * BBX_NCLS] : lm.head;
            int pref[BBX_NCLS + 1];
            pref[0] = 0;
            #pragma unroll
            for (int k = 0; k < BBX_NCLS; k++) pref[k + 1] = pref[k] + ((cnt[k] + 31) & ~31);
            const int total = pref[BBX_NCLS];
            for (int t = wtid; t < total; t += stride) {
                int k = 0, kbase = 0;
                #pragma unroll
                for (int kk = 1; kk < BBX_NCLS; kk++) if (t >= pref[kk]) { k = kk; kbase = pref[kk]; }
                int idx = t - kbase;
                if (idx >= cnt[k]) continue;
                int4 nd = ld128i_cg_pol(&dv.Nd_k[k][head[k] + idx], pol_stream);
                bool hasA = nd.z >= 0, hasB = nd.w >= 0;
                BodyState A, B;
                if (k == 0) {
                    if (hasA) load_body_keep(dv, nd.z, A, pol_keep);
                    if (hasB) load_body_keep(dv, nd.w, B, pol_keep);
                } else load_bodies_keep(dv, nd.z, nd.w, hasA, hasB, A, B, pol_keep);
                float4* M = dv.M + (size_t)nd.x;
                const int ms = k <= 4 ? cnt[k] : 1;                   // distance between a node's consecutive records (RECORD SLOTS above)
                // software pipeline: the record of constraint j+1 is in flight while constraint j is swept
                float4 n0 = ld128_cg_pol(M, pol_stream), n1 = ld128_cg_pol(M + PL, pol_stream), n2 = ld128_cg_pol(M + 2 * PL, pol_stream), n3 = ld128_cg_pol(M + 3 * PL, pol_stream);
                for (int j = 0; j < nd.y; j++, M += ms) {
                    float4 c0 = n0, c1 = n1, c2 = n2, c3 = n3;
                    if (j + 1 < nd.y) { n0 = ld128_cg_pol(M + ms, pol_stream); n1 = ld128_cg_pol(M + ms + PL, pol_stream); n2 = ld128_cg_pol(M + ms + 2 * PL, pol_stream); n3 = ld128_cg_pol(M + ms + 3 * PL, pol_stream); }
                    if (k == 0) solve_constraint<false>(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                    else        solve_constraint<true>(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                    st128_cg_pol(M + 3 * PL, c3, pol_stream);
                }
                if (hasA) store_body_keep(dv, nd.z, A, pol_keep);
                if (hasB) store_body_keep(dv, nd.w, B, pol_keep);
            }
            grid.sync();
            if (it == 1 && tid == 0 && l < 4096) {
                unsigned long long ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
                dv.level_ns[l + 1] = ns;
                if (l == 0) dv.level_ns[0] = 0;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Batched small worlds, second form: one BLOCK per GROUP of worlds, one thread per constraint.
//
// A 256-body world offers ~20 independent constraints per dependency level: a warp per world keeps a
// third of its lanes busy (and sweeps a manifold of m contacts in m dependent rounds).  Here a block of
// 128 threads owns a group of worlds (as many as it takes to run the whole batch in one wave) and sweeps
// level l of all of them together: every constraint is its own node (chains per constraint, as in the flow
// solver), so the lanes of a warp do equal work and levels are separated by __syncthreads().
// Iteration 0 discovers the group's levels (Kahn) and writes descriptors and records in visiting order;
// iterations 1..I-1 stream them.  Same per-body constraint order as everywhere else: bit-identical results.
// Two forms (chosen per launch by the host): body velocities in the packed global array (L2), which a
// block-wide barrier keeps coherent (any group), or -- groups of one or two worlds -- in shared memory
// (see the comment at the kernel).
// ---------------------------------------------------------------------------
#define GRP_THREADS 128
#ifndef GRP_FLAT
#define GRP_FLAT true                  // branch-free sweep (bbx_solve.cuh): a level of a small world is ONE warp, latency bound
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
#define GRP_BLOCKS_PER_SM 4
#define GRP_LMAX 2047                  // dependency levels of a group

// ---- shared-memory body state of a group (k_solve_groups, small groups) ----
__device__ __forceinline__ void grp_cp16(void* sdst, const void* g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned int)__cvta_generic_to_shared(sdst)), "l"(g) : "memory");
}
__device__ __forceinline__ void grp_cp_commit_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_all;" ::: "memory"); }
// both bodies of a constraint from the group's slots (ia / ib = 0xffff: absent); same values as load_bodies_keep
__device__ __forceinline__ void load_bodies_smem(const float4* s_vel, const float4* s_boxq, unsigned int nsl, unsigned int ia, unsigned int ib,
                                                 BodyState& A, BodyState& B)
{
    const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const bool hasA = ia != 0xffffu, hasB = ib != 0xffffu;
    const bool boxA = hasA && ia >= nsl, boxB = hasB && ib >= nsl;
    F8 va, vb, qa, qb;
    va.a = va.b = vb.a = vb.b = qa.a = qa.b = qb.a = qb.b = z4;
    if (hasA) { va.a = s_vel[2 * ia]; va.b = s_vel[2 * ia + 1]; }
    if (boxA) { qa.a = s_boxq[2 * (ia - nsl)]; qa.b = s_boxq[2 * (ia - nsl) + 1]; }
    if (hasB) { vb.a = s_vel[2 * ib]; vb.b = s_vel[2 * ib + 1]; }
    if (boxB) { qb.a = s_boxq[2 * (ib - nsl)]; qb.b = s_boxq[2 * (ib - nsl) + 1]; }
    A.v = xyz(va.a); A.w = xyz(va.b); A.im = va.a.w; A.ii = va.b.w; A.box = boxA; A.q = qa.a; A.iI = xyz(qa.b);
    B.v = xyz(vb.a); B.w = xyz(vb.b); B.im = vb.a.w; B.ii = vb.b.w; B.box = boxB; B.q = qb.a; B.iI = xyz(qb.b);
    { const float3 u = f3(A.q.x, A.q.y, A.q.z); A.k1 = A.q.w * A.q.w - v3dot(u, u); A.k2 = 2.0f * A.q.w; }
    { const float3 u = f3(B.q.x, B.q.y, B.q.z); B.k1 = B.q.w * B.q.w - v3dot(u, u); B.k2 = 2.0f * B.q.w; }
}
__device__ __forceinline__ void store_body_smem(float4* s_vel, unsigned int i, const BodyState& s)
{
    s_vel[2 * i] = make_float4(s.v.x, s.v.y, s.v.z, s.im); s_vel[2 * i + 1] = make_float4(s.w.x, s.w.y, s.w.z, s.ii);
}

__global__ void __launch_bounds__(256) k_frontier0_groups(BbxDev dv, int* gq, int* gq_count, int wpb, int gcap)
{
    __shared__ int s_solver;
    if (threadIdx.x == 0) s_solver = 0;
    __syncthreads();
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    int mine = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        int gA = dv.node[2 * (size_t)c].x, gB = dv.node[2 * (size_t)c + 1].x;
        if (gA < 0 && gB < 0) continue;
        mine++;
        if (dv.indeg[c] == 0) {
            int x = gA >= 0 ? gA : gB;
            int w = x < dv.NS ? x / dv.cap_s : (x - dv.NS) / dv.cap_b;
            int g = w / wpb;
            int slot = atomicAdd(&gq_count[g], 1);
            // the LAST group may hold fewer than wpb worlds: its slice ends with the allocation, not after gcap slots
            const long long glim = min((long long)gcap, (long long)dv.cap_contacts - (long long)g * gcap);
            if (slot < glim) gq[(size_t)g * gcap + slot] = c;
        }
    }
    if (mine) atomicAdd(&s_solver, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_solver) { atomicAdd(&dv.ctr->n_solver, s_solver); atomicAdd(&dv.ctr->n_nodes, s_solver); }
}

// SHARED-MEMORY FORM (smem_bodies, small groups).  A level of a small group is ONE warp walking one constraint per lane
// (~700 instructions, two dependent chains), and the step time of a small batch is the number of block-synchronised
// levels of its deepest world times that latency, whatever the number of worlds: the per-GPU share of a batch that is
// partitioned over many GPUs is bound by it.  ncu on the global-memory form at 512 worlds: per issued instruction 1.09
// cycles of long-scoreboard stall (queue entry -> node -> bodies in iteration 0, the body gather after every barrier in
// the replay) on top of 0.72 of fixed-latency waits.  The block is the only reader and writer of its worlds' velocities
// during the solve, so here they live in shared memory (32 bytes per body slot + 32 per box: rotation and inverse
// inertia; loaded once, written back once), together with the visiting order (iteration 0 reads its queue from shared
// memory) and the two body slots of every constraint (16 bits each); the next level's record is staged by cp.async
// into thread-private slots while the current level is swept, so that a replayed level starts from shared memory only.
// Same per-body order, same arithmetic: bit-identical.  Measured (512 worlds of config 5): profiles/r02_notes.md.
__global__ void __launch_bounds__(GRP_THREADS, GRP_BLOCKS_PER_SM) k_solve_groups(BbxDev dv, SolveParams sp, int* gq, const int* gq_count, int gcap,
                                                                                 int wpb, int pair_cap, int smem_bodies)
{
    __shared__ int level_start[GRP_LMAX + 2];
    __shared__ int s_next[3];                                         // pushes of level l are counted in s_next[l % 3] (see the Kahn loop)
    extern __shared__ __align__(16) int g_pair[];                     // shared-memory form: [pair_cap] body slots, [pair_cap] visiting order, bodies, record staging
    #define GRP_STAMP(i) do { if (sp.gather < 0 && tid == 0 && g < 256) { unsigned long long ns_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_)); dv.level_ns[g * 8 + (i)] = ns_; } } while (0)
    const int g = blockIdx.x, tid = threadIdx.x;
    const int vt = tid;                                               // (rotating the block's warps by a per-block offset measured SLOWER: profiles/r02_notes.md)
    // slots this group owns in queue / M / Nd: gcap, except for a ragged last group (fewer than wpb worlds), whose slice
    // ends with the allocations; every append and the overflow test are bounded by it
    const int glim = (int)min((long long)gcap, (long long)dv.cap_contacts - (long long)g * gcap);
    if (gq_count[g] > glim) { if (tid == 0) atomicOr(&dv.ctr->overflow, 32); return; }
    const int n0 = gq_count[g];
    if (n0 == 0) return;                                              // no dynamic contact in this group
    int* queue = gq + (size_t)g * gcap;                               // visiting order of the group
    const size_t base = (size_t)g * gcap;                             // the group's slice of M / Nd / slot numbers
    int4* gd = dv.Nd_k[0] + base;                                     // descriptors in visiting order: (gid A, gid B, contact, -)
    const size_t PL = (size_t)dv.cap_contacts;
    if (tid == 0) { s_next[0] = s_next[1] = s_next[2] = 0; level_start[0] = 0; }
    GRP_STAMP(0);
    const bool big = dv.ctr->n_contacts > BBX_L2_STREAM_CONTACTS;     // bodies stay in L2, records stream
    const unsigned long long pol_keep = big ? l2_policy_keep() : l2_policy_normal(), pol_stream = big ? l2_policy_stream() : l2_policy_normal();
    // shared-memory form: body slot of a gid = gid - s_off (spheres) / gid - b_off (boxes, after the nsl sphere slots)
    const int nsl = wpb * dv.cap_s, nbl = wpb * (dv.cap_s + dv.cap_b);
    const int s_off = g * wpb * dv.cap_s, b_off = dv.NS + g * wpb * dv.cap_b - nsl;
    int* s_queue = g_pair + pair_cap;                                   // visiting order: read by iteration 0 only, so that
    float4* s_rec = reinterpret_cast<float4*>(s_queue);                 // the replay's record staging [plane][thread] shares its place
    float4* s_vel = reinterpret_cast<float4*>(s_queue + max(pair_cap, 16 * GRP_THREADS));
    float4* s_boxq = s_vel + 2 * nbl;
    const int nwg = min(wpb, dv.n_worlds - g * wpb);                     // worlds of this group (the last group may be ragged)
    const int ns_valid = nwg * dv.cap_s, nb_valid = nwg * dv.cap_b;
    bool in_smem = smem_bodies != 0;                                    // (uniform over the block) the velocities live in s_vel
    if (in_smem) {
        for (int i = tid; i < nbl; i += GRP_THREADS) {
            if (!(i < nsl ? i < ns_valid : (i - nsl) < nb_valid)) continue;
            const int gid = i < nsl ? s_off + i : b_off + i;
            grp_cp16(&s_vel[2 * i], &dv.vel[2 * (size_t)gid]); grp_cp16(&s_vel[2 * i + 1], &dv.vel[2 * (size_t)gid + 1]);
            if (i >= nsl) { const size_t bx = (size_t)(gid - dv.NS); grp_cp16(&s_boxq[2 * (i - nsl)], &dv.boxq[2 * bx]); grp_cp16(&s_boxq[2 * (i - nsl) + 1], &dv.boxq[2 * bx + 1]); }
        }
        for (int p = tid; p < min(n0, pair_cap); p += GRP_THREADS) s_queue[p] = __ldcg(&queue[p]);
        grp_cp_commit_wait();
    }
    __syncthreads();
    // ---- iteration 0: Kahn levels ----
    int head = 0, tail = n0, lvl = 0, par = 0;
    bool bad = false;
    while (head < tail) {
        for (int p = head + vt; p < tail; p += GRP_THREADS) {
            const int c = (in_smem && p < pair_cap) ? s_queue[p] : __ldcg(&queue[p]);
            F8 nd = ld256_nc_pol((const float4*)&dv.node[2 * (size_t)c], pol_stream);    // (gid A, ...) (gid B, ...)
            const int2 lk = ld64_pol(&dv.link[c], pol_stream);                           // successors on A's and B's chain
            const int gA = __float_as_int(nd.a.x), gB = __float_as_int(nd.b.x);
            const int sa = lk.x, sb = lk.y;
            const float4* L = dv.L + 4 * (size_t)c;
            F8 lo = ld256_nc_pol(L, pol_stream), hi = ld256_nc_pol(L + 2, pol_stream);
            float4 c0 = lo.a, c1 = lo.b, c2 = hi.a, c3 = hi.b;
            const bool hasA = gA >= 0, hasB = gB >= 0;
            BodyState A, B;
            if (in_smem) {
                const unsigned int ia = gA < 0 ? 0xffffu : (unsigned int)(gA < dv.NS ? gA - s_off : gA - b_off);
                const unsigned int ib = gB < 0 ? 0xffffu : (unsigned int)(gB < dv.NS ? gB - s_off : gB - b_off);
                if (p < pair_cap) g_pair[p] = (int)(ia | (ib << 16));
                load_bodies_smem(s_vel, s_boxq, (unsigned int)nsl, ia, ib, A, B);
                solve_constraint<GRP_FLAT>(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                if (hasA) store_body_smem(s_vel, ia, A);
                if (hasB) store_body_smem(s_vel, ib, B);
            } else {
                load_bodies_keep(dv, gA, gB, hasA, hasB, A, B, pol_keep);
                solve_constraint<GRP_FLAT>(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                if (hasA) store_body_keep(dv, gA, A, pol_keep);
                if (hasB) store_body_keep(dv, gB, B, pol_keep);
            }
            float4* M = dv.M + base + p;
            st128_cg_pol(M, c0, pol_stream); st128_cg_pol(M + PL, c1, pol_stream); st128_cg_pol(M + 2 * PL, c2, pol_stream); st128_cg_pol(M + 3 * PL, c3, pol_stream);
            __stcg(&gd[p], make_int4(gA, gB, c, 0));
            dv.slot_of[c] = (int)(base + p);
            // both in-degree atomics are in flight before either result is used (one L2 round trip instead of two)
            const int ra = sa >= 0 ? atomicSub(&dv.indeg[sa], 1) : 0;
            const int rb = sb >= 0 ? atomicSub(&dv.indeg[sb], 1) : 0;
            if (ra == 1) { int q = tail + atomicAdd(&s_next[par], 1); if (q < glim) { __stcg(&queue[q], sa); if (in_smem && q < pair_cap) s_queue[q] = sa; } }
            if (rb == 1) { int q = tail + atomicAdd(&s_next[par], 1); if (q < glim) { __stcg(&queue[q], sb); if (in_smem && q < pair_cap) s_queue[q] = sb; } }
        }
        // ONE barrier per level: the counter of level l + 2 (= that of level l - 1, which every thread has read before it
        // arrived here) is cleared after this barrier and first used after the next one
        __syncthreads();
        const int pushed = s_next[par];
        head = tail; tail += pushed; lvl++;
        if (tail > glim || lvl > GRP_LMAX) { bad = true; break; }
        if (tid == 0) { s_next[par == 0 ? 2 : par - 1] = 0; level_start[lvl] = head; }
        par = par == 2 ? 0 : par + 1;
    }
    if (bad) { if (tid == 0) atomicOr(&dv.ctr->overflow, 32); return; }
    const int n_levels = lvl;
    if (tid == 0) { atomicMax(&dv.ctr->n_levels, n_levels); level_start[n_levels] = tail; }
    __syncthreads();
    GRP_STAMP(1);
    if (sp.gather < 0 && tid == 0 && g < 256) { dv.level_ns[g * 8 + 5] = n_levels; dv.level_ns[g * 8 + 7] = tail; }
    if (in_smem && sp.iterations > 1 && tail <= pair_cap) {
        // ---- iterations 1..I-1 from shared memory ----
        const bool stage = n_levels > 1;                                // (one level: the next record IS the current one, whose impulses are being updated)
        if (stage) { const int pn = level_start[0] + vt; if (pn < level_start[1]) { const float4* Mn = dv.M + base + pn; for (int r = 0; r < 4; r++) grp_cp16(&s_rec[r * GRP_THREADS + tid], Mn + r * PL); } }
        grp_cp_commit_wait();
        for (int it = 1; it < sp.iterations; it++) {
            for (int l = 0; l < n_levels; l++) {
                const int s0 = level_start[l], s1 = level_start[l + 1];
                float4 c0, c1, c2, c3;
                if (stage && s0 + vt < s1) { c0 = s_rec[tid]; c1 = s_rec[GRP_THREADS + tid]; c2 = s_rec[2 * GRP_THREADS + tid]; c3 = s_rec[3 * GRP_THREADS + tid]; }
                if (stage) {
                    const int ln = (l + 1 < n_levels) ? l + 1 : 0;
                    const int pn = level_start[ln] + vt;
                    if (pn < level_start[ln + 1]) { const float4* Mn = dv.M + base + pn; for (int r = 0; r < 4; r++) grp_cp16(&s_rec[r * GRP_THREADS + tid], Mn + r * PL); }
                }
                for (int p = s0 + vt; p < s1; p += GRP_THREADS) {
                    float4* M = dv.M + base + p;
                    if (!stage || p != s0 + vt) { c0 = ld128_cg_pol(M, pol_stream); c1 = ld128_cg_pol(M + PL, pol_stream); c2 = ld128_cg_pol(M + 2 * PL, pol_stream); c3 = ld128_cg_pol(M + 3 * PL, pol_stream); }   // (levels wider than the block)
                    const unsigned int pr = (unsigned int)g_pair[p];
                    const unsigned int ia = pr & 0xffffu, ib = pr >> 16;
                    const bool hasA = ia != 0xffffu, hasB = ib != 0xffffu;
                    BodyState A, B;
                    load_bodies_smem(s_vel, s_boxq, (unsigned int)nsl, ia, ib, A, B);
                    solve_constraint<GRP_FLAT>(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                    if (hasA) store_body_smem(s_vel, ia, A);
                    if (hasB) store_body_smem(s_vel, ib, B);
                    st128_cg_pol(M + 3 * PL, c3, pol_stream);
                }
                grp_cp_commit_wait();
                __syncthreads();
            }
        }
    }
    if (in_smem) {
        // the velocities go back to the packed array (also when the group turned out too large for the shared-memory
        // replay: the level replay below then continues on global memory)
        for (int i = tid; i < nbl; i += GRP_THREADS) {
            if (!(i < nsl ? i < ns_valid : (i - nsl) < nb_valid)) continue;
            const int gid = i < nsl ? s_off + i : b_off + i;
            st256_cg_pol(&dv.vel[2 * (size_t)gid], s_vel[2 * i], s_vel[2 * i + 1], pol_keep);
        }
        if (tail <= pair_cap || sp.iterations <= 1) { GRP_STAMP(4); return; }
        __syncthreads();
    }
    // ---- iterations 1..I-1: stream the visiting order ----
    // The descriptor of this thread's first constraint of the NEXT level is fetched while the current level is
    // swept (descriptors are static), so that after the barrier the body loads can be issued at once; the next
    // level's records are pulled into L2 at the same time (prefetch, no registers).
    int2 dn = make_int2(-1, -1);
    { const int pn = level_start[0] + vt; if (pn < level_start[1]) { int4 t4 = ld128i_cg_pol(&gd[pn], pol_stream); dn = make_int2(t4.x, t4.y); } }
    for (int it = 1; it < sp.iterations; it++) {
        for (int l = 0; l < n_levels; l++) {
            const int s0 = level_start[l], s1 = level_start[l + 1];
            const int2 dc = dn;
            {
                const int ln = (l + 1 < n_levels) ? l + 1 : 0;
                const int pn = level_start[ln] + vt;
                dn = make_int2(-1, -1);
                if (pn < level_start[ln + 1]) {
                    int4 t4 = ld128i_cg_pol(&gd[pn], pol_stream); dn = make_int2(t4.x, t4.y);
                    const float4* Mn = dv.M + base + pn;
                    prefetch_l2(Mn); prefetch_l2(Mn + PL); prefetch_l2(Mn + 2 * PL);
                    if (n_levels > 1) prefetch_l2(Mn + 3 * PL);
                }
            }
            for (int p = s0 + vt; p < s1; p += GRP_THREADS) {
                int2 d = dc;
                if (p != s0 + vt) { int4 t4 = ld128i_cg_pol(&gd[p], pol_stream); d = make_int2(t4.x, t4.y); }      // (levels wider than the block)
                float4* M = dv.M + base + p;
                float4 c0 = ld128_cg_pol(M, pol_stream), c1 = ld128_cg_pol(M + PL, pol_stream), c2 = ld128_cg_pol(M + 2 * PL, pol_stream), c3 = ld128_cg_pol(M + 3 * PL, pol_stream);
                const bool hasA = d.x >= 0, hasB = d.y >= 0;
                BodyState A, B;
                load_bodies_keep(dv, d.x, d.y, hasA, hasB, A, B, pol_keep);
                solve_constraint<GRP_FLAT>(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                if (hasA) store_body_keep(dv, d.x, A, pol_keep);
                if (hasB) store_body_keep(dv, d.y, B, pol_keep);
                st128_cg_pol(M + 3 * PL, c3, pol_stream);
            }
            __syncthreads();
        }
    }
    GRP_STAMP(4);
}

// Batches go to the group solver when there are enough worlds to fill the SMs with groups and a world is small enough
// for a block to sweep its levels in a round or two; anything else (a few large worlds) runs on the level kernel.
#define GRP_WORLD_BODIES_MAX 4096
static bool use_world_groups(const BbxDev* d, int imported, int schedule)
{
    return !imported && d->n_worlds >= 16 && schedule == BBX_SOLVER_EXACT && d->cap_s + d->cap_b <= GRP_WORLD_BODIES_MAX;
}

// contact -> record slot for the level path, built ON DEMAND (export of the constraint list, warm-start save): the level-major
// queues hold the node's head contact, the descriptors its first slot and length.  Iteration 0 used to scatter these 4 bytes
// per constraint itself: 24 MB of data, ~380 MB of DRAM read-modify-write traffic in the hottest kernel, for a table that
// resident stepping never reads.
__global__ void __launch_bounds__(256) k_build_slot_of(BbxDev dv)
{
    const int n_levels = dv.ctr->n_levels;
    for (int l = blockIdx.x; l < n_levels; l += gridDim.x) {
        for (int k = 0; k < BBX_NCLS; k++) {
            const int cnt = dv.level_count[l * BBX_NCLS + k];
            if (cnt == 0) continue;
            int head = 0;
            for (int l2 = 0; l2 < l; l2++) head += dv.level_count[l2 * BBX_NCLS + k];
            const int stride = k <= 4 ? cnt : 1;                                  // RECORD SLOTS (k_solve_levels)
            for (int idx = threadIdx.x; idx < cnt; idx += blockDim.x) {
                const int c = dv.queue_k[k][head + idx];
                const int4 nd = dv.Nd_k[k][head + idx];
                for (int j = 0; j < nd.y; j++) dv.slot_of[c + j] = nd.x + j * stride;
            }
        }
    }
}

int bbx_ensure_slot_of(BbxDev* d)
{
    if (!d->slot_of_stale) return 0;
    BBX_CUDA(d, cudaMemsetAsync(d->slot_of, 0xff, sizeof(int) * (size_t)d->cap_contacts, d->stream));      // -1: not scheduled
    BBX_LAUNCH(d, k_build_slot_of, BBX_NUM_SMS * 4, 256, 0, *d);
    BBX_CUDA(d, cudaGetLastError());
    d->slot_of_stale = 0;
    return 0;
}

// warm-start table, allocated at first use: at most half full
int bbx_warm_ensure(BbxDev* d)
{
    if (d->warm_key) return 0;
    unsigned long long cap = 1024;
    while (cap < 2ull * (unsigned long long)d->cap_contacts) cap <<= 1;
    if (cap > (1ull << 31)) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "warm starting: contact capacity too large", __FILE__, __LINE__);
    BBX_CUDA(d, cudaMalloc((void**)&d->warm_key, sizeof(unsigned long long) * cap));
    BBX_CUDA(d, cudaMalloc((void**)&d->warm_val, sizeof(float4) * cap));
    BBX_CUDA(d, cudaMalloc((void**)&d->warm_epoch, sizeof(int) * (size_t)d->n_worlds));
    d->warm_mask = (unsigned int)(cap - 1);
    BBX_CUDA(d, cudaMemsetAsync(d->warm_key, 0xff, sizeof(unsigned long long) * cap, d->stream));
    BBX_CUDA(d, cudaMemsetAsync(d->warm_epoch, 0, sizeof(int) * (size_t)d->n_worlds, d->stream));
    return 0;
}

// steps with joints replay one iteration per launch of k_replay_staged, whose tables hold RS_MAXLEV levels: a deeper
// contact graph cannot be replayed that way and is reported (overflow bit 16) instead of being skipped silently
__global__ void k_joint_guard(BbxDev dv) { if (dv.ctr->replay_done) atomicOr(&dv.ctr->overflow, 16); }

// touched-body statistics and sanity for the host
int bbx_stage_solve(BbxDev* d, const BbxStepParams* p) { return bbx_stage_solve_ex(d, p, 0, p->settings.solver_iterations, 0); }

// presolve_mode: see k_presolve; iterations: Gauss-Seidel sweeps; imported: the contact list came from
// the host in array order (public stage functions), so the per-body order is the array index
int bbx_stage_solve_ex(BbxDev* d, const BbxStepParams* p, int presolve_mode, int iterations, int imported)
{
    const int G = BBX_NUM_SMS * 8;
    // joints (extension, bbx_joints.cu): their rows are swept after the contacts in EVERY iteration, so a step with joints
    // runs the exact level schedule one iteration per launch with a joint sweep after each (PhysicsStep paths only)
    const int nj = (!imported && presolve_mode == 0) ? d->n_joints : 0;
    const int schedule = nj ? BBX_SOLVER_EXACT : p->solver_schedule;
    // opt-in (BBX_SOLVER_EXACT_FLOW or BBX_SOLVER_FLOW=1): barrier-free flow solver, every constraint its own node
    const int flow = (!nj && (schedule == BBX_SOLVER_EXACT_FLOW || (d->force_flow && schedule == BBX_SOLVER_EXACT))) ? 1 : 0;
    const int groups = (!nj && !flow && use_world_groups(d, imported, schedule)) ? 1 : 0;
    const int per_contact = flow | groups;                 // every constraint is its own node
    BBX_CUDA(d, cudaMemsetAsync(d->deg, 0, sizeof(int) * ((size_t)d->N + 1), d->stream));
    const float warm = (presolve_mode == 0 && !imported && p->warm_start > 0.0f) ? p->warm_start : 0.0f;
    if (warm > 0.0f) { int rcw = bbx_warm_ensure(d); if (rcw) return rcw; }
    // default path: level-scheduled kernels (single world, or batches the group solver does not take)
    const int levels_path = (!flow && !groups) ? 1 : 0;
    BBX_LAUNCH(d, k_body_records, bbx_blocks(d->N, 256), 256, 0, *d);
    BBX_LAUNCH(d, k_presolve, G, 256, 0, *d, p->settings.baumgarte_bias, p->settings.allowed_penetration, p->dt, presolve_mode, per_contact, warm, (levels_path && iterations > 0) ? 1 : 0);
    if (nj) { int rcj = bbx_joints_presolve(d, p, iterations > 0 ? 1 : 0); if (rcj) return rcj; }
    if (iterations <= 0) { bbx_timing_mark(d, 4); BBX_CUDA(d, cudaGetLastError()); return 0; }    // :1223 loop does not run
    int rc = bbx_scan_exclusive(d, d->deg, d->adj_start, d->adj_cursor, d->N + 1, NULL, NULL);
    if (rc) return rc;
    BBX_LAUNCH(d, k_adj_fill, G, 256, 0, *d, imported ? 2 : (schedule == BBX_SOLVER_COLOURED ? 1 : 0), per_contact);
    if (levels_path) BBX_CUDA(d, cudaMemsetAsync(d->level_count, 0, 2 * BBX_NCLS * sizeof(int), d->stream));
    BBX_LAUNCH(d, k_adj_link, bbx_blocks(d->N, 128), 128, 0, *d, per_contact, flow, warm > 0.0f ? 1 : 0, levels_path);
    if (warm > 0.0f) BBX_LAUNCH(d, k_warm_apply, bbx_blocks(d->N, 128), 128, 0, *d, per_contact, flow, p->settings.friction);
    d->last_flow = flow; d->last_per_contact = per_contact;
    d->slot_of_stale = levels_path;          // the level kernels do not write slot_of: bbx_ensure_slot_of builds it on demand
    if (flow) return bbx_flow_solve(d, p, iterations);
    SolveParams sp;
    sp.restitution = p->settings.restitution; sp.friction = p->settings.friction;
    sp.vel_threshold = p->settings.velocity_threshold; sp.iterations = iterations; sp.gather = 0; sp.first_only = 0;
    if (groups) {
        // batched small worlds: a block per group of worlds, sized so that the batch runs in one wave
        if (d->grp_blocks == 0) {
            int per_sm = 0;
            BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_groups, GRP_THREADS, 0));
            if (per_sm < 1) return bbx_fail(d, BBX_ERR_NO_DEVICE, "group solver kernel does not fit on an SM", __FILE__, __LINE__);
            d->grp_blocks = per_sm * d->num_sms;
            BBX_CUDA(d, cudaFuncSetAttribute(k_solve_groups, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        }
        int wpb = (d->n_worlds + d->grp_blocks - 1) / d->grp_blocks;
        if (wpb < 1) wpb = 1;
        if (d->grp_force_wpb > wpb) wpb = d->grp_force_wpb;           // BBX_GRP_WPB (tests): larger groups than the batch needs
        const int n_groups = (d->n_worlds + wpb - 1) / wpb;
        long long gcap_ll = (long long)wpb * d->cap_contacts_world;
        if (gcap_ll > d->cap_contacts) gcap_ll = d->cap_contacts;
        const int gcap = (int)gcap_ll;
        int* gq = d->queue_k[0];                       // capacity cap_contacts
        int* gq_count = d->adj_cursor;                 // free again after k_adj_link; N + 1 >= n_groups ints
        BBX_CUDA(d, cudaMemsetAsync(gq_count, 0, sizeof(int) * (size_t)n_groups, d->stream));
        BBX_LAUNCH(d, k_frontier0_groups, G, 256, 0, *d, gq, gq_count, wpb, gcap);
        bbx_timing_mark(d, 4);
        // small groups keep their body state, visiting order and body slots in shared memory (see k_solve_groups)
        int pair_cap = 0, smem_bodies = 0;
        size_t dyn_bytes = 0;
        {
            const size_t nbl = (size_t)wpb * (size_t)(d->cap_s + d->cap_b), nboxl = (size_t)wpb * (size_t)d->cap_b;
            if (iterations > 1 && wpb <= d->grp_smem_wpb && nbl < 65535) {
                pair_cap = gcap < 1408 * wpb ? gcap : 1408 * wpb;        // (a group with more constraints continues on the global-memory replay)
                if (d->grp_pair_cap > 0 && d->grp_pair_cap < pair_cap) pair_cap = d->grp_pair_cap;   // BBX_GRP_PAIR_CAP (tests): forces that hand-over
                pair_cap &= ~3;                                          // keeps what follows 16-byte aligned
                const size_t q_ints = pair_cap > 16 * GRP_THREADS ? (size_t)pair_cap : (size_t)(16 * GRP_THREADS);   // visiting order / record staging
                dyn_bytes = ((size_t)pair_cap + q_ints) * sizeof(int) + (2 * nbl + 2 * nboxl) * sizeof(float4);
                if (dyn_bytes <= 65536 && pair_cap > 0) smem_bodies = 1; else { pair_cap = 0; dyn_bytes = 0; }
            }
        }
        {   // the shared-memory form wants the SM's whole carve-out (4 blocks x ~55 KB); the global-memory form wants L1 (measured:
            // 1024 worlds in the global form 0.79 ms with the default carve-out, 0.85 ms with the maximum)
            static int carveout_set[64];                             // per device: 0 unknown, 1 default, 2 maximum
            const int want = smem_bodies ? 2 : 1, dev = d->device & 63;
            if (carveout_set[dev] != want) {
                BBX_CUDA(d, cudaFuncSetAttribute(k_solve_groups, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                 smem_bodies ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault));
                carveout_set[dev] = want;
            }
        }
        if (d->trace_level < 0) sp.gather = d->trace_level;         // BBX_TRACE_LEVEL=-1: per-block phase stamps in level_ns (tools/profile_batch.py)
        BBX_LAUNCH(d, k_solve_groups, n_groups, GRP_THREADS, dyn_bytes, *d, sp, gq, gq_count, gcap, wpb, pair_cap, smem_bodies);
        BBX_CUDA(d, cudaGetLastError());
        return 0;
    }
    // (level 0 of the class queues was filled by k_adj_link)
    // iteration 0 (level discovery fused with the first sweep) in k_solve_levels; iterations 1..I-1 replay the recorded
    // level-major order in k_replay_staged (BBX_REPLAY=0: inside k_solve_levels, the round-1 replay loop)
    sp.first_only = (iterations > 1 && d->replay_mode != 0) ? d->replay_mode : 0;
    if (nj) { sp.iterations = 1; sp.first_only = 2; }       // iteration 0 only; the later ones are single-iteration replays below
    // coloured schedule: colours are taken during iteration 0 and the replay runs colour-major (bbx_colour.cu)
    const int colour = (schedule == BBX_SOLVER_COLOURED && !imported && sp.first_only) ? 1 : 0;
    if (colour) {
        int rcc = bbx_colour_ensure(d); if (rcc) return rcc;
        BBX_CUDA(d, cudaMemsetAsync(d->col_used, 0, sizeof(unsigned long long) * (size_t)d->N, d->stream));
        BBX_CUDA(d, cudaMemsetAsync(&d->ctr->colour_overflow, 0, sizeof(int), d->stream));
    }
    int* coop_blocks = colour ? &d->coop_blocks_col : &d->coop_blocks;
    void* kernel = colour ? (void*)k_solve_levels<true> : (void*)k_solve_levels<false>;
    if (*coop_blocks == 0) {
        int per_sm = 0, coop = 0;
        BBX_CUDA(d, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d->device));
        if (!coop) return bbx_fail(d, BBX_ERR_NO_DEVICE, "device lacks cooperative launch", __FILE__, __LINE__);
        if (colour) BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_levels<true>, SOLVE_THREADS, 0));
        else        BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_levels<false>, SOLVE_THREADS, 0));
        if (per_sm < 1) return bbx_fail(d, BBX_ERR_NO_DEVICE, "solver kernel does not fit on an SM", __FILE__, __LINE__);
        if (per_sm > SOLVE_BLOCKS_PER_SM) per_sm = SOLVE_BLOCKS_PER_SM;
        *coop_blocks = per_sm * d->num_sms;
        if (d->tune_solve_blocks > 0 && d->tune_solve_blocks < *coop_blocks) *coop_blocks = d->tune_solve_blocks;
    }
    bbx_timing_mark(d, 4);
    if (d->trace_level > 0) { sp.gather = d->trace_level; BBX_CUDA(d, cudaMemsetAsync(d->level_ns + FLOW_DBG, 0, 8 * sizeof(unsigned long long), d->stream)); }
    BbxDev dv = *d;
    void* args[] = { (void*)&dv, (void*)&sp };
    BBX_CUDA(d, cudaLaunchCooperativeKernel(kernel, dim3(*coop_blocks), dim3(SOLVE_THREADS), args, 0, d->stream));
    d->launches++;
    BBX_CUDA(d, cudaGetLastError());
    if (nj) {
        // contacts of iteration it, then the joint rows of iteration it (oracle/bbx_oracle.c stage_solve)
        if (iterations > 1) BBX_LAUNCH(d, k_joint_guard, 1, 1, 0, *d);
        int rcj = bbx_joints_sweep(d);
        SolveParams one = sp;
        one.iterations = 2; one.gather = 0;                  // k_replay_staged runs iterations 1 .. I-1: exactly one sweep
        for (int it = 1; it < iterations && !rcj; it++) {
            rcj = bbx_replay_staged(d, one);
            if (!rcj) rcj = bbx_joints_sweep(d);
        }
        return rcj;
    }
    if (colour) { int rcc = bbx_colour_rebucket(d); if (rcc) return rcc; }
    if (sp.first_only) {
        if (d->timing_on && d->timing_steps < d->timing_cap) { bbx_timing_mark(d, 7); d->timing_has7[d->timing_steps] = 1; }
        return bbx_replay_staged(d, sp);
    }
    return 0;
}
