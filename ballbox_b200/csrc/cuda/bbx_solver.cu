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
namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------
// K9 presolve: contact -> 64-byte solver record (contact order); node heads also
// get their graph fields
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tangents_of(float3 n, float3& t1, float3& t2)          // :1329-1336
{
    if (fabsf(n.x) >= 0.57735f) t1 = f3(n.y, -n.x, 0.0f);
    else                        t1 = f3(0.0f, n.z, -n.y);
    t1 = v3norm(t1);
    t2 = v3cross(n, t1);
}

__device__ __forceinline__ float inv_mass_sum(float imA, float imB, float3 rA, float3 rB, float3 iA, float3 iB, float3 dir)
{
    float3 ra = v3cross(rA, dir), rb = v3cross(rB, dir);
    return imA + imB + v3dot(ra, v3mul(ra, iA)) + v3dot(rb, v3mul(rb, iB));             // :1323-1325
}

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
        dv.slot_of[c] = -1;
        if (per_contact) {
            dv.node[2 * (size_t)c] = make_int4(dynA ? a : -1, 0, 0, -1);
            dv.node[2 * (size_t)c + 1] = make_int4(dynB ? b : -1, 0, 0, -1);
            dv.indeg[c] = (dynA ? 1 : 0) + (dynB ? 1 : 0);
            if (dynA) atomicAdd(&dv.deg[a], 1);
            if (dynB) atomicAdd(&dv.deg[b], 1);
        } else if (RUN_K(id.z) == 0) {                                                  // node head
            int m = RUN_LEN(id.z);
            dv.node[2 * (size_t)c] = make_int4(dynA ? a : -1, dynB ? b : -1, m, bbx_node_class(m, a >= dv.NS || b >= dv.NS));
            dv.node[2 * (size_t)c + 1] = make_int4(-1, -1, 0, 0);
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
        int4 l = dv.node[2 * (size_t)c];
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

__device__ __forceinline__ void flow_init_body(const BbxDev& dv, int g);
__device__ __forceinline__ void wake_body(const BbxDev& dv, int g);

// One thread per body: sort the body's node list into chain order and write the successor links.  An entry is
// (order key << 32) | side << 31 | node (side 0: the body is A of the node, 1: B), so that no node record has to be
// read back.  Lists of up to ADJ_LOCAL entries are sorted in a thread-private array (all entries are requested at
// once); longer ones (a big dynamic body) in place in global memory.
#define ADJ_LOCAL 24
#define ADJ_NODE(e) ((int)((unsigned int)(e) & 0x7fffffffu))
#define ADJ_SIDE(e) ((int)(((unsigned int)(e)) >> 31))

__device__ __forceinline__ bool adj_after(const BbxDev& dv, unsigned long long a, unsigned long long v)
{
    // a sorts after v?  Equal 32-bit keys can only happen in the coloured schedule (hash collision): fall back to the
    // deterministic total order so that the result never depends on the contact numbering
    return (a >> 32) > (v >> 32) || ((a >> 32) == (v >> 32) && node_before(dv, ADJ_NODE(v), ADJ_NODE(a)));
}

// circular (unrolled group solver): the last constraint of a chain links back to the first one, flagged ADJ_WRAP
// ("the body's next visit belongs to the next sweep")
#define ADJ_WRAP 0x40000000
__device__ __forceinline__ void adj_write_links(const BbxDev& dv, int x, const unsigned long long* lst, int d, int per_contact, int circular)
{
    if (per_contact) {
        // (position, length) and successor of every constraint in this body's chain
        for (int i = 0; i < d; i++) {
            const int cur = ADJ_NODE(lst[i]);
            const int nxt = (i + 1 < d) ? ADJ_NODE(lst[i + 1]) : (circular ? (ADJ_NODE(lst[0]) | ADJ_WRAP) : -1);
            dv.node[2 * (size_t)cur + ADJ_SIDE(lst[i])] = make_int4(x, i, d, nxt);
        }
    } else {
        for (int i = 1; i < d; i++) {
            const unsigned long long e2 = lst[i];
            const int link = ADJ_NODE(e2) | ((int)((e2 >> 32) & 7u) << 28);          // successor id | its class
            int* nx = (int*)&dv.node[2 * (size_t)ADJ_NODE(lst[i - 1]) + 1];
            nx[ADJ_SIDE(lst[i - 1])] = link;
        }
    }
}

// keep_sorted: the sorted list is written back (k_warm_apply walks it)
// push_frontier (level path): the thread whose decrement takes a node's in-degree to zero appends it to level 0 of
// its class queue (this used to be a pass over all contacts, k_frontier0)
__global__ void __launch_bounds__(128) k_adj_link(BbxDev dv, int per_contact, int flow_records, int keep_sorted, int push_frontier, int circular)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    bool has = false;
    int ready = -1, ready_cls = -1;
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
                    if (per_contact) dv.node[2 * (size_t)ADJ_NODE(cur) + ADJ_SIDE(cur)] = make_int4(x, i, d, (nxt_valid) ? ADJ_NODE(nxt) : (circular ? (ADJ_NODE(e0) | ADJ_WRAP) : -1)); \
                    else if (i > 0) ((int*)&dv.node[2 * (size_t)ADJ_NODE(prev) + 1])[ADJ_SIDE(prev)] = ADJ_NODE(cur) | ((int)((cur >> 32) & 7u) << 28); \
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
                adj_write_links(dv, x, loc, d, per_contact, circular);
                if (keep_sorted) for (int i = 0; i < d; i++) lst[i] = loc[i];
            } else {
                for (int i = 1; i < d; i++) {
                    unsigned long long v = lst[i];
                    int j = i - 1;
                    while (j >= 0 && adj_after(dv, lst[j], v)) { lst[j + 1] = lst[j]; j--; }
                    lst[j + 1] = v;
                }
                if (atomicSub(&dv.indeg[ADJ_NODE(lst[0])], 1) == 1) { ready = ADJ_NODE(lst[0]); ready_cls = (int)((lst[0] >> 32) & 7u); }
                adj_write_links(dv, x, lst, d, per_contact, circular);
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

// ---------------------------------------------------------------------------
// K11 velocity solve
// ---------------------------------------------------------------------------
typedef unsigned long long P2;                   // two packed floats (f32x2)
struct PackConst { P2 neg0, one, mone, two, sgn; };   // (-0,-0) (1,1) (-1,-1) (2,2) (-1,1): opaque to ptxas on purpose
struct SolveParams { float restitution, friction, vel_threshold; int iterations; int gather; int first_only; int levels_known; int packed; PackConst pc; };

struct BodyState {
    float3 v, w;
    float im;           // 1/mass
    float ii;           // spheres: 1/inertia
    float3 iI;          // boxes: 1/diag(I)
    float4 q;           // boxes: rotation
    float k1, k2;       // boxes: q.w*q.w - |q.xyz|^2 and 2*q.w, the two scalars of QuatRotateVec3 (:635-637)
    bool box;
};

// QuatRotateVec3 (:630-639) with the body-constant scalars hoisted; `sgn` = -1 rotates by the conjugate
// (|u|^2 and s are the same for q and its conjugate, so k1 and k2 are shared)
__device__ __forceinline__ float3 qrot_pre(float3 u, float k1, float k2, float3 v)
{
    float3 a = v3scale(v, k1);
    float3 b = v3scale(u, 2.0f * v3dot(u, v));
    float3 c = v3scale(v3cross(u, v), k2);
    return v3add(v3add(a, b), c);
}

__device__ __forceinline__ void load_body(const BbxDev& dv, int g, BodyState& s)
{
    F8 vw = ld256_cg(&dv.vel[2 * g]);
    float4 l = vw.a, a = vw.b;
    s.v = xyz(l); s.w = xyz(a); s.im = l.w; s.ii = a.w;
    s.box = g >= dv.NS;
    if (s.box) {
        int b = g - dv.NS;
        F8 qi = ld256_nc(&dv.boxq[2 * b]);               // rotation and inverse inertia do not change during the solve
        s.q = qi.a; s.iI = xyz(qi.b);
        float3 u = f3(s.q.x, s.q.y, s.q.z);
        s.k1 = s.q.w * s.q.w - v3dot(u, u); s.k2 = 2.0f * s.q.w;
    }
}

__device__ __forceinline__ void store_body(const BbxDev& dv, int g, const BodyState& s)
{
    st256_cg(&dv.vel[2 * g], make_float4(s.v.x, s.v.y, s.v.z, s.im), make_float4(s.w.x, s.w.y, s.w.z, s.ii));
}

// level kernel: the same with an L2 "keep" policy
__device__ __forceinline__ void load_body_keep(const BbxDev& dv, int g, BodyState& s, unsigned long long pol)
{
    F8 vw = ld256_cg_pol(&dv.vel[2 * g], pol);
    float4 l = vw.a, a = vw.b;
    s.v = xyz(l); s.w = xyz(a); s.im = l.w; s.ii = a.w;
    s.box = g >= dv.NS;
    if (s.box) {
        int b = g - dv.NS;
        F8 qi = ld256_nc_pol(&dv.boxq[2 * b], pol);
        s.q = qi.a; s.iI = xyz(qi.b);
        float3 u = f3(s.q.x, s.q.y, s.q.z);
        s.k1 = s.q.w * s.q.w - v3dot(u, u); s.k2 = 2.0f * s.q.w;
    }
}
__device__ __forceinline__ void store_body_keep(const BbxDev& dv, int g, const BodyState& s, unsigned long long pol)
{
    st256_cg_pol(&dv.vel[2 * g], make_float4(s.v.x, s.v.y, s.v.z, s.im), make_float4(s.w.x, s.w.y, s.w.z, s.ii), pol);
}

__device__ __forceinline__ void wake_body(const BbxDev& dv, int g)                      // :1373-1376, :1402-1405
{
    unsigned char f = dv.flags[g];
    if (f & BF_SLEEP) { dv.flags[g] = f & (unsigned char)~BF_SLEEP; dv.timer[g] = 0.0f; }
}

// ---- flow solver: versioned body records (BbxDev::flow) ----
__device__ __forceinline__ unsigned long long ll_pack(float x, int ver)
{
    return ((unsigned long long)(unsigned int)ver << 32) | (unsigned long long)__float_as_uint(x);
}
// velocities of body g after its visit number `ver`; 1/m, 1/I and the dependency depth ride in the last two words
__device__ __forceinline__ void flow_store(const BbxDev& dv, int g, const BodyState& s, int ver, int depth)
{
    unsigned long long* p = reinterpret_cast<unsigned long long*>(dv.flow + 4 * (size_t)g);
    asm volatile("st.global.cg.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(ll_pack(s.v.x, ver)), "l"(ll_pack(s.v.y, ver)),
                 "l"(ll_pack(s.v.z, ver)), "l"(ll_pack(s.w.x, ver)) : "memory");
    asm volatile("st.global.cg.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p + 4), "l"(ll_pack(s.w.y, ver)), "l"(ll_pack(s.w.z, ver)),
                 "l"(ll_pack(s.im, __float_as_int(s.ii))), "l"((unsigned long long)(unsigned int)depth) : "memory");
}
__device__ __forceinline__ void flow_init_body(const BbxDev& dv, int g)
{
    F8 vw = ld256_cg(&dv.vel[2 * g]);
    BodyState s;
    s.v = xyz(vw.a); s.w = xyz(vw.b); s.im = vw.a.w; s.ii = vw.b.w;
    flow_store(dv, g, s, 0, 0);
}
// one poll of the whole record (L2 only): true when all six velocity words carry version `need`
__device__ __forceinline__ bool flow_try_load(const BbxDev& dv, int g, int need, BodyState& s, int& depth)
{
    const unsigned long long* p = reinterpret_cast<const unsigned long long*>(dv.flow + 4 * (size_t)g);
    unsigned long long x0, x1, x2, x3, x4, x5, x6, x7;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x0), "=l"(x1), "=l"(x2), "=l"(x3) : "l"(p) : "memory");
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x4), "=l"(x5), "=l"(x6), "=l"(x7) : "l"(p + 4) : "memory");
    const unsigned int nd = (unsigned int)need;
    bool ok = ((unsigned int)(x0 >> 32) == nd) & ((unsigned int)(x1 >> 32) == nd) & ((unsigned int)(x2 >> 32) == nd) &
              ((unsigned int)(x3 >> 32) == nd) & ((unsigned int)(x4 >> 32) == nd) & ((unsigned int)(x5 >> 32) == nd);
    s.v = f3(__uint_as_float((unsigned int)x0), __uint_as_float((unsigned int)x1), __uint_as_float((unsigned int)x2));
    s.w = f3(__uint_as_float((unsigned int)x3), __uint_as_float((unsigned int)x4), __uint_as_float((unsigned int)x5));
    s.im = __uint_as_float((unsigned int)x6); s.ii = __uint_as_float((unsigned int)(x6 >> 32));
    depth = (int)(unsigned int)x7;
    return ok;
}
// rotation and inverse inertia of a box (constant during the solve)
__device__ __forceinline__ void flow_load_box(const BbxDev& dv, int g, BodyState& s)
{
    s.box = g >= dv.NS;
    if (s.box) {
        int b = g - dv.NS;
        F8 qi = ld256_nc(&dv.boxq[2 * b]);
        s.q = qi.a; s.iI = xyz(qi.b);
        float3 u = f3(s.q.x, s.q.y, s.q.z);
        s.k1 = s.q.w * s.q.w - v3dot(u, u); s.k2 = 2.0f * s.q.w;
    }
}

// ApplyConstraintImpulse :1359-1408
__device__ __forceinline__ void apply_impulse(BodyState& s, float3 J, float3 L)
{
    s.v = v3add(s.v, v3scale(J, s.im));
    if (!s.box) {
        s.w = v3add(s.w, v3scale(L, s.ii));
    } else {
        float3 u = f3(s.q.x, s.q.y, s.q.z), um = f3(-s.q.x, -s.q.y, -s.q.z);
        float3 loc = qrot_pre(um, s.k1, s.k2, L);
        float3 dl = v3mul(loc, s.iI);
        s.w = v3add(s.w, qrot_pre(u, s.k1, s.k2, dl));
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

// one constraint, one Gauss-Seidel visit (:1414-1488) on bodies held in registers
__device__ __forceinline__ void solve_constraint(float4 c0, float4 c1, float4 c2, float4& c3, bool hasA, bool hasB,
                                                 BodyState& A, BodyState& B, const SolveParams& sp)
{
    float3 n = xyz(c0), rA = xyz(c1), rB = xyz(c2);
    float bias = c0.w, mN = c1.w, mT0 = c2.w, mT1 = c3.w;
    const float3 zero = f3(0.0f, 0.0f, 0.0f);
    float3 vA = hasA ? v3add(A.v, v3cross(A.w, rA)) : zero;                  // :1233-1261
    float3 vB = hasB ? v3add(B.v, v3cross(B.w, rB)) : zero;
    float3 rel = v3sub(vB, vA);
    float vn = v3dot(rel, n);
    float vbias = 0.0f;
    if (vn < -sp.vel_threshold) vbias = -sp.restitution * vn;                // :1443-1445
    float jn = mN * (-(vn + vbias) + bias);                                  // :1447
    float old = c3.x;
    c3.x = bb_fmaxf(0.0f, c3.x + jn);                                        // :1451
    jn = c3.x - old;
    float3 P = v3scale(n, jn);
    float3 Pm = v3scale(P, -1.0f);
    if (hasA) apply_impulse(A, Pm, v3cross(rA, Pm));                         // :1456-1459
    if (hasB) apply_impulse(B, P, v3cross(rB, P));
    if (sp.friction > 0.0f) {                                                // :1462-1487
        vA = hasA ? v3add(A.v, v3cross(A.w, rA)) : zero;
        vB = hasB ? v3add(B.v, v3cross(B.w, rB)) : zero;
        rel = v3sub(vB, vA);
        float3 t1, t2;
        tangents_of(n, t1, t2);
        float maxF = sp.friction * c3.x;
        {
            float vt = v3dot(rel, t1);
            float jt = mT0 * (-vt);
            float o = c3.y;
            c3.y = bb_fmaxf(-maxF, bb_fminf(maxF, c3.y + jt));
            jt = c3.y - o;
            float3 T = v3scale(t1, jt), Tm = v3scale(T, -1.0f);
            if (hasA) apply_impulse(A, Tm, v3cross(rA, Tm));
            if (hasB) apply_impulse(B, T, v3cross(rB, T));
        }
        {
            float vt = v3dot(rel, t2);                                       // same pre-loop rel (:1466 is outside the j loop)
            float jt = mT1 * (-vt);
            float o = c3.z;
            c3.z = bb_fmaxf(-maxF, bb_fminf(maxF, c3.z + jt));
            jt = c3.z - o;
            float3 T = v3scale(t2, jt), Tm = v3scale(T, -1.0f);
            if (hasA) apply_impulse(A, Tm, v3cross(rA, Tm));
            if (hasB) apply_impulse(B, T, v3cross(rB, T));
        }
    }
}

// ---------------------------------------------------------------------------
// The same visit with PACKED FP32x2 arithmetic (sm_100a FFMA2): body A in the low half, body B in the high half of
// every 64-bit operand.  The two bodies of a constraint go through structurally identical code, so velocity-at-point
// and the three impulse applications (~80 % of the ~700 instructions) run once for both.
// Bit-exactness: every operation is an fma whose result is the IEEE result of the scalar operation it stands for,
//   a * b = fma(a, b, -0)     a + b = fma(a, 1, b)     a - b = fma(b, -1, a)
// and the constants -0, 1, -1 come from the kernel parameters (SolveParams::pc).  With literal constants ptxas folds
// these back into FMUL2 / FADD2 and then CONTRACTS neighbouring pairs into one FFMA2 (seen in the SASS, and in 2220
// differing impulses on a 20k-sphere scene); it cannot do either with operands it does not know.
// Mixed pairs (sphere + box) compute the box form and the sphere form and select per half.
// (tools/micro/f32x2_bench.cu: one FFMA2 costs what one FFMA costs, in latency and in issue slots.)
// ---------------------------------------------------------------------------
struct P3 { P2 x, y, z; };
__device__ __forceinline__ P2 pk(float a, float b) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float plo(P2 p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); return a; }
__device__ __forceinline__ float phi(P2 p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); return b; }
__device__ __forceinline__ P2 pfma(P2 a, P2 b, P2 c) { P2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
#define PMUL(a, b) pfma((a), (b), K.neg0)
#define PADD(a, b) pfma((a), K.one, (b))
#define PSUB(a, b) pfma((b), K.mone, (a))
__device__ __forceinline__ P3 p3(float3 a, float3 b) { P3 r; r.x = pk(a.x, b.x); r.y = pk(a.y, b.y); r.z = pk(a.z, b.z); return r; }
__device__ __forceinline__ P3 p3add(const PackConst& K, P3 a, P3 b) { P3 r; r.x = PADD(a.x, b.x); r.y = PADD(a.y, b.y); r.z = PADD(a.z, b.z); return r; }
__device__ __forceinline__ P3 p3scale(const PackConst& K, P3 v, P2 s) { P3 r; r.x = PMUL(v.x, s); r.y = PMUL(v.y, s); r.z = PMUL(v.z, s); return r; }
__device__ __forceinline__ P3 p3mul(const PackConst& K, P3 a, P3 b) { P3 r; r.x = PMUL(a.x, b.x); r.y = PMUL(a.y, b.y); r.z = PMUL(a.z, b.z); return r; }
__device__ __forceinline__ P2 p3dot(const PackConst& K, P3 a, P3 b) { return PADD(PADD(PMUL(a.x, b.x), PMUL(a.y, b.y)), PMUL(a.z, b.z)); }
__device__ __forceinline__ P3 p3cross(const PackConst& K, P3 a, P3 b)
{
    P3 r;
    r.x = PSUB(PMUL(a.y, b.z), PMUL(a.z, b.y)); r.y = PSUB(PMUL(a.z, b.x), PMUL(a.x, b.z)); r.z = PSUB(PMUL(a.x, b.y), PMUL(a.y, b.x));
    return r;
}

struct PairState {          // bodies A (low halves) and B (high halves)
    P3 v, w;
    P2 im, ii, k1, k2;
    P3 iI, u, um;
    bool boxA, boxB;
};

__device__ __forceinline__ void pair_from(const BodyState& A, const BodyState& B, bool hasA, bool hasB, PairState& s)
{
    const float3 z = f3(0.0f, 0.0f, 0.0f);
    s.boxA = hasA && A.box; s.boxB = hasB && B.box;
    s.v = p3(hasA ? A.v : z, hasB ? B.v : z); s.w = p3(hasA ? A.w : z, hasB ? B.w : z);
    s.im = pk(hasA ? A.im : 0.0f, hasB ? B.im : 0.0f);
    s.ii = pk(hasA && !A.box ? A.ii : 0.0f, hasB && !B.box ? B.ii : 0.0f);
    if (s.boxA || s.boxB) {
        const float3 ua = s.boxA ? f3(A.q.x, A.q.y, A.q.z) : z, ub = s.boxB ? f3(B.q.x, B.q.y, B.q.z) : z;
        s.u = p3(ua, ub); s.um = p3(f3(-ua.x, -ua.y, -ua.z), f3(-ub.x, -ub.y, -ub.z));
        s.iI = p3(s.boxA ? A.iI : z, s.boxB ? B.iI : z);
        s.k1 = pk(s.boxA ? A.k1 : 0.0f, s.boxB ? B.k1 : 0.0f); s.k2 = pk(s.boxA ? A.k2 : 0.0f, s.boxB ? B.k2 : 0.0f);
    }
}

__device__ __forceinline__ P3 pqrot_pre(const PackConst& K, P3 u, P2 k1, P2 k2, P3 v)      // qrot_pre on both halves
{
    P3 a = p3scale(K, v, k1);
    P3 b = p3scale(K, u, PMUL(K.two, p3dot(K, u, v)));
    P3 c = p3scale(K, p3cross(K, u, v), k2);
    return p3add(K, p3add(K, a, b), c);
}

__device__ __forceinline__ P2 psel(bool lo_first, bool hi_first, P2 first, P2 second)
{
    return pk(lo_first ? plo(first) : plo(second), hi_first ? phi(first) : phi(second));
}

// ApplyConstraintImpulse on both bodies: A receives -X, B receives +X (X = n * jn or a tangent impulse)
__device__ __forceinline__ void pair_apply(const PackConst& K, PairState& s, P3 r, float3 X)
{
    // (-1) * X for A (v3scale(X, -1.0f)); X itself for B (1 * X is exact)
    P3 J; J.x = PMUL(pk(X.x, X.x), K.sgn); J.y = PMUL(pk(X.y, X.y), K.sgn); J.z = PMUL(pk(X.z, X.z), K.sgn);
    P3 L = p3cross(K, r, J);
    s.v = p3add(K, s.v, p3scale(K, J, s.im));
    P3 dw = p3scale(K, L, s.ii);                                                      // spheres
    if (s.boxA || s.boxB) {
        P3 loc = pqrot_pre(K, s.um, s.k1, s.k2, L);
        P3 dl = p3mul(K, loc, s.iI);
        P3 db = pqrot_pre(K, s.u, s.k1, s.k2, dl);
        dw.x = psel(s.boxA, s.boxB, db.x, dw.x); dw.y = psel(s.boxA, s.boxB, db.y, dw.y); dw.z = psel(s.boxA, s.boxB, db.z, dw.z);
    }
    s.w = p3add(K, s.w, dw);
}

__device__ __forceinline__ float3 pair_rel(const PackConst& K, const PairState& s, P3 r, bool hasA, bool hasB)   // vB - vA at the contact
{
    P3 vel = p3add(K, s.v, p3cross(K, s.w, r));
    const float3 z = f3(0.0f, 0.0f, 0.0f);
    float3 vA = hasA ? f3(plo(vel.x), plo(vel.y), plo(vel.z)) : z;
    float3 vB = hasB ? f3(phi(vel.x), phi(vel.y), phi(vel.z)) : z;
    return v3sub(vB, vA);
}

__device__ __forceinline__ void solve_constraint_packed(float4 c0, float4 c1, float4 c2, float4& c3, bool hasA, bool hasB,
                                                        PairState& s, const SolveParams& sp)
{
    const PackConst& K = sp.pc;
    float3 n = xyz(c0);
    float bias = c0.w, mN = c1.w, mT0 = c2.w, mT1 = c3.w;
    const P3 r = p3(xyz(c1), xyz(c2));
    float3 rel = pair_rel(K, s, r, hasA, hasB);
    float vn = v3dot(rel, n);
    float vbias = 0.0f;
    if (vn < -sp.vel_threshold) vbias = -sp.restitution * vn;                // :1443-1445
    float jn = mN * (-(vn + vbias) + bias);                                  // :1447
    float old = c3.x;
    c3.x = bb_fmaxf(0.0f, c3.x + jn);                                        // :1451
    jn = c3.x - old;
    pair_apply(K, s, r, v3scale(n, jn));                                     // :1456-1459
    if (sp.friction > 0.0f) {                                                // :1462-1487
        rel = pair_rel(K, s, r, hasA, hasB);
        float3 t1, t2;
        tangents_of(n, t1, t2);
        float maxF = sp.friction * c3.x;
        {
            float vt = v3dot(rel, t1);
            float jt = mT0 * (-vt);
            float o = c3.y;
            c3.y = bb_fmaxf(-maxF, bb_fminf(maxF, c3.y + jt));
            jt = c3.y - o;
            pair_apply(K, s, r, v3scale(t1, jt));
        }
        {
            float vt = v3dot(rel, t2);                                       // same pre-loop rel (:1466 is outside the j loop)
            float jt = mT1 * (-vt);
            float o = c3.z;
            c3.z = bb_fmaxf(-maxF, bb_fminf(maxF, c3.z + jt));
            jt = c3.z - o;
            pair_apply(K, s, r, v3scale(t2, jt));
        }
    }
}

__device__ __forceinline__ void pair_to(const PairState& s, bool hasA, bool hasB, BodyState& A, BodyState& B)
{
    if (hasA) { A.v = f3(plo(s.v.x), plo(s.v.y), plo(s.v.z)); A.w = f3(plo(s.w.x), plo(s.w.y), plo(s.w.z)); }
    if (hasB) { B.v = f3(phi(s.v.x), phi(s.v.y), phi(s.v.z)); B.w = f3(phi(s.w.x), phi(s.w.y), phi(s.w.z)); }
}

// Per level, the nodes of class k sit in queue_k[k][head_k .. head_k + cnt_k).  The flattened
// index space pads every class to a multiple of 32 so that a warp never mixes classes.
struct LevelMap { int head[BBX_NCLS]; int cnt[BBX_NCLS]; int pref[BBX_NCLS + 1]; };

__device__ __forceinline__ void level_map_load(LevelMap& lm, const int* lc_level, bool first_level)
{
    // called by one thread; heads advance by the previous level's counts
    int run = 0;
    for (int k = 0; k < BBX_NCLS; k++) {
        if (first_level) lm.head[k] = 0; else lm.head[k] += lm.cnt[k];
        lm.cnt[k] = __ldcg(&lc_level[k]);
        lm.pref[k] = run;
        run += (lm.cnt[k] + 31) & ~31;
    }
    lm.pref[BBX_NCLS] = run;
}

// 3 blocks of 256 threads per SM (80 registers).  One 768-thread block per SM was measured
// slower (6.6 vs 5.8 ms per step at 1M bodies) although it makes the barrier cheaper.
#define LC_CACHE 512
#define SOLVE_THREADS 256
#define SOLVE_BLOCKS_PER_SM 2
#define PEND_ROUNDS 4                  // rounds of a level between two block-aggregated appends

// Grid-wide barrier: cooperative-groups grid.sync() measured 1.2 us at 444 x 256 threads on B200,
// a hand-written arrive/poll barrier 1.7 us (tools/micro/barrier_bench.cu), so we use the former.
// append to the next level's class queues: lanes are grouped by class with one match_any
__device__ __forceinline__ void push_successor(const BbxDev& dv, int* lc_next, const LevelMap& lm, int link)
{
    bool has = link >= 0;
    unsigned int active = __ballot_sync(0xffffffffu, has);
    if (!has) return;
    int cls = (link >> 28) & 7, node = link & 0x0fffffff;
    unsigned int peers = __match_any_sync(active, cls);
    int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&lc_next[cls], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    int slot = base + __popc(peers & ((1u << lane) - 1u));
    dv.queue_k[cls][lm.head[cls] + lm.cnt[cls] + slot] = node;
}

// Level discovery WITHOUT the sweep (Kahn's algorithm on the successor links), at high occupancy: visits the
// nodes level by level, hands out their level-major record slots, copies their records from contact order (L)
// to the plane-major level order (M) and appends the successors that became ready.  The sweep of iteration 0
// then is an ordinary replay.  Opt-in (BBX_FUSED_KAHN=0): measured equal at 1M bodies (this kernel 0.94 ms, the
// solve kernel 0.91 ms shorter: a level of the discovery costs ~18 us with or without the sweep, it is the chain
// queue -> node -> in-degree atomics -> class-queue append) and slower on the 100k-200k configs.
#define KAHN_THREADS 256
#define KAHN_BLOCKS_PER_SM 4

__global__ void __launch_bounds__(KAHN_THREADS, KAHN_BLOCKS_PER_SM) k_kahn_levels(BbxDev dv)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ LevelMap lm;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    const int wtid = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
    int* lc = dv.level_count;
    const size_t PL = (size_t)dv.cap_contacts;
    int lvl = 0;
    if (threadIdx.x == 0) level_map_load(lm, lc, true);
    __syncthreads();
    while (lm.pref[BBX_NCLS] > 0) {
        if (lvl + 2 >= dv.cap_levels) { if (tid == 0) atomicOr(&dv.ctr->overflow, 16); break; }
        if (tid < BBX_NCLS) lc[(lvl + 2) * BBX_NCLS + tid] = 0;
        int* lc_next = lc + (lvl + 1) * BBX_NCLS;
        const int total = lm.pref[BBX_NCLS];
        for (int t = wtid; t < total; t += stride) {
            int k = 0;
            #pragma unroll
            for (int kk = 1; kk < BBX_NCLS; kk++) if (t >= lm.pref[kk]) k = kk;
            int idx = t - lm.pref[k];
            bool live = idx < lm.cnt[k];
            int p = lm.head[k] + idx;
            int sa = -1, sb = -1, m = 0, c = 0;
            int4 nd0 = make_int4(-1, -1, 0, 0);
            if (live) {
                c = __ldcg(&dv.queue_k[k][p]);
                F8 rec = ld256_nc((const float4*)&dv.node[2 * (size_t)c]);       // ids, run length, successors: one sector
                nd0 = make_int4(__float_as_int(rec.a.x), __float_as_int(rec.a.y), __float_as_int(rec.a.z), 0);
                sa = __float_as_int(rec.b.x); sb = __float_as_int(rec.b.y);
                m = nd0.z;
                // the successors first: their readiness is what the next level waits for
                if (sa >= 0 && atomicSub(&dv.indeg[sa & 0x0fffffff], 1) != 1) sa = -1;
                if (sb >= 0 && atomicSub(&dv.indeg[sb & 0x0fffffff], 1) != 1) sb = -1;
            }
            push_successor(dv, lc_next, lm, sa);
            push_successor(dv, lc_next, lm, sb);
            int mbase = warp_reserve(&dv.ctr->n_mslots, m);
            if (live) {
                __stcg(&dv.Nd_k[k][p], make_int4(mbase, m, nd0.x, nd0.y));
                const float4* L = dv.L + 4 * (size_t)c;
                for (int j = 0; j < m; j++, L += 4) {
                    F8 lo = ld256_nc(L), hi = ld256_nc(L + 2);
                    float4* M = dv.M + (size_t)(mbase + j);
                    __stcg(M, lo.a); __stcg(M + PL, lo.b); __stcg(M + 2 * PL, hi.a); __stcg(M + 3 * PL, hi.b);
                    dv.slot_of[c + j] = mbase + j;
                }
            }
        }
        grid.sync();
        if (threadIdx.x == 0) level_map_load(lm, lc + (lvl + 1) * BBX_NCLS, false);
        __syncthreads();
        lvl++;
    }
    if (tid == 0) dv.ctr->n_levels = lvl;
}

template <bool PACKED>
__global__ void __launch_bounds__(SOLVE_THREADS, SOLVE_BLOCKS_PER_SM) k_solve_levels(BbxDev dv, SolveParams sp)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ LevelMap lm;
    __shared__ int s_pcnt[BBX_NCLS], s_pbase[BBX_NCLS];
    constexpr bool PACKED0 = false;                 // iteration 0 keeps the scalar sweep (register budget of the discovery pass)
    const bool big = dv.ctr->n_contacts > BBX_L2_STREAM_CONTACTS;
    const unsigned long long pol_keep = big ? l2_policy_keep() : l2_policy_normal(), pol_stream = big ? l2_policy_stream() : l2_policy_normal();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    // work index of this thread: consecutive 32-node chunks go round-robin over the BLOCKS, so a level
    // with few nodes is spread over all SMs instead of filling the first few blocks
    const int wtid = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
    int* lc = dv.level_count;                       // lc[level * BBX_NCLS + class]
    const size_t PL = (size_t)dv.cap_contacts;

    // ---- iteration 0: discover levels (Kahn) while solving (only when k_kahn_levels has not run) ----
    int lvl = 0;
    if (threadIdx.x == 0) level_map_load(lm, lc, true);
    __syncthreads();
    while (!sp.levels_known && lm.pref[BBX_NCLS] > 0) {
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
        for (int r = 0; r < rounds; r++) {
            const int t = wtid + r * stride;
            int k = 0;
            #pragma unroll
            for (int kk = 1; kk < BBX_NCLS; kk++) if (t >= lm.pref[kk]) k = kk;
            int idx = t - lm.pref[k];
            bool live = t < total && idx < lm.cnt[k];
            int p = lm.head[k] + idx;
            int sa = -1, sb = -1, m = 0, c = 0;
            int4 nd0 = make_int4(-1, -1, 0, 0);
            if (live) {
                c = __ldcg(&dv.queue_k[k][p]);
                F8 rec = ld256_nc_pol((const float4*)&dv.node[2 * (size_t)c], pol_stream);       // ids, run length, successors: one sector
                nd0 = make_int4(__float_as_int(rec.a.x), __float_as_int(rec.a.y), __float_as_int(rec.a.z), 0);
                sa = __float_as_int(rec.b.x); sb = __float_as_int(rec.b.y);
                m = nd0.z;
            }
            int mbase = warp_reserve(&dv.ctr->n_mslots, m);
            if (live) {
                __stcg(&dv.Nd_k[k][p], make_int4(mbase, m, nd0.x, nd0.y));
                bool hasA = nd0.x >= 0, hasB = nd0.y >= 0;
                BodyState A, B;
                if (hasA) load_body_keep(dv, nd0.x, A, pol_keep);
                if (hasB) load_body_keep(dv, nd0.y, B, pol_keep);
                // contact-order records are gathered once (two 256-bit loads, the next record in flight while
                // the current constraint is swept), then written plane-major
                const float4* L = dv.L + 4 * (size_t)c;
                F8 lo = ld256_nc_pol(L, pol_stream), hi = ld256_nc_pol(L + 2, pol_stream);
                PairState ps;
                if (PACKED0) pair_from(A, B, hasA, hasB, ps);
                for (int j = 0; j < m; j++, L += 4) {
                    float4 c0 = lo.a, c1 = lo.b, c2 = hi.a, c3 = hi.b;
                    if (j + 1 < m) { lo = ld256_nc_pol(L + 4, pol_stream); hi = ld256_nc_pol(L + 6, pol_stream); }
                    if (PACKED0) solve_constraint_packed(c0, c1, c2, c3, hasA, hasB, ps, sp);
                    else         solve_constraint(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                    float4* M = dv.M + (size_t)(mbase + j);
                    st128_cg_pol(M, c0, pol_stream); st128_cg_pol(M + PL, c1, pol_stream); st128_cg_pol(M + 2 * PL, c2, pol_stream); st128_cg_pol(M + 3 * PL, c3, pol_stream);
                    dv.slot_of[c + j] = mbase + j;
                }
                if (PACKED0) pair_to(ps, hasA, hasB, A, B);
                if (hasA) store_body_keep(dv, nd0.x, A, pol_keep);    // (the bodies were woken by k_adj_link: one coalesced pass over the
                if (hasB) store_body_keep(dv, nd0.y, B, pol_keep);    //  touched bodies instead of a random flag access per node here)
                if (sa >= 0 && atomicSub(&dv.indeg[sa & 0x0fffffff], 1) != 1) sa = -1;
                if (sb >= 0 && atomicSub(&dv.indeg[sb & 0x0fffffff], 1) != 1) sb = -1;
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
        grid.sync();
        if (threadIdx.x == 0) level_map_load(lm, lc + (lvl + 1) * BBX_NCLS, false);
        __syncthreads();
        lvl++;
    }
    const int n_levels = sp.levels_known ? dv.ctr->n_levels : lvl;
    if (tid == 0 && !sp.levels_known) dv.ctr->n_levels = n_levels;
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
    // (with first_only the later iterations run in k_replay_pairs, two lanes per node)
    for (int it = (sp.levels_known ? 0 : 1); it < (sp.first_only ? 1 : sp.iterations); it++) {
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
                if (hasA) load_body_keep(dv, nd.z, A, pol_keep);
                if (hasB) load_body_keep(dv, nd.w, B, pol_keep);
                float4* M = dv.M + (size_t)nd.x;
                // software pipeline: the record of constraint j+1 is in flight while constraint j is swept
                float4 n0 = ld128_cg_pol(M, pol_stream), n1 = ld128_cg_pol(M + PL, pol_stream), n2 = ld128_cg_pol(M + 2 * PL, pol_stream), n3 = ld128_cg_pol(M + 3 * PL, pol_stream);
                PairState ps;
                if (PACKED) pair_from(A, B, hasA, hasB, ps);
                for (int j = 0; j < nd.y; j++, M++) {
                    if (PACKED && j > 0) { n0 = __ldcg(M); n1 = __ldcg(M + PL); n2 = __ldcg(M + 2 * PL); n3 = __ldcg(M + 3 * PL); }
                    float4 c0 = n0, c1 = n1, c2 = n2, c3 = n3;
                    if (!PACKED && j + 1 < nd.y) { n0 = ld128_cg_pol(M + 1, pol_stream); n1 = ld128_cg_pol(M + 1 + PL, pol_stream); n2 = ld128_cg_pol(M + 1 + 2 * PL, pol_stream); n3 = ld128_cg_pol(M + 1 + 3 * PL, pol_stream); }
                    if (PACKED) solve_constraint_packed(c0, c1, c2, c3, hasA, hasB, ps, sp);
                    else        solve_constraint(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                    st128_cg_pol(M + 3 * PL, c3, pol_stream);
                }
                if (PACKED) pair_to(ps, hasA, hasB, A, B);
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
#define FLOW_DBG 4088                  // level_ns[FLOW_DBG..+8): chunk visits, processing passes, spin passes, failed lane polls,
                                       // warp wait cycles, warp busy cycles, end of iteration 0 (ns), end of kernel (ns)
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

// ---------------------------------------------------------------------------
// Replay of the level-major order with TWO LANES PER NODE (iterations 1..I-1 of k_solve_levels; opt-in,
// BBX_REPLAY=2: measured equal to the one-lane replay, kept as the base for further work)
//
// The even lane of a pair holds body A, the odd lane body B (solve_half): a node's sweep is half as
// deep and a thread holds one body instead of two, which leaves registers to software-pipeline the
// warp's chunks of a level (next chunk's gather in flight during the current sweep).  Same level
// order, same per-body constraint order, same operations per body: bit-identical to the one-lane sweep.
// ---------------------------------------------------------------------------
#define RP_THREADS 256
#define RP_BLOCKS_PER_SM 2
struct PairRaw { F8 vw, qi; float4 c0, c1, c3; };

__global__ void __launch_bounds__(RP_THREADS, RP_BLOCKS_PER_SM) k_replay_pairs(BbxDev dv, SolveParams sp)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_head[BBX_NCLS], s_cnt[BBX_NCLS], s_pref[BBX_NCLS + 1];
    __shared__ int lc_cache[LC_CACHE * BBX_NCLS];
    const int n_levels = dv.ctr->n_levels;
    const int* lc = dv.level_count;
    for (int i = threadIdx.x; i < min(n_levels, LC_CACHE) * BBX_NCLS; i += blockDim.x) lc_cache[i] = lc[i];
    const int lane = threadIdx.x & 31;
    const bool side = lane & 1;
    const size_t PL = (size_t)dv.cap_contacts;
    // WORK DISTRIBUTION.  Chunks of different classes cost 1..8 sweeps, so dealing them round-robin leaves most
    // warps waiting at the barrier for the few that drew the long ones (ncu: half of all stall samples; a global
    // ticket counter balances perfectly but its atomics cost more than they save: 6.2 vs 5.5 ms).  Static instead:
    // the level's chunks are ordered from the longest class to the shortest and dealt in boustrophedon order
    // (round 0: warp 0..W-1, round 1: warp W-1..0, ...), so the warp that drew the longest chunk of one round
    // draws the shortest of the next.
    const int warp0 = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;      // consecutive chunks go to different SMs
    const int W = gridDim.x * (blockDim.x >> 5);
    __syncthreads();
    for (int it = (sp.levels_known ? 0 : 1); it < sp.iterations; it++) {
        for (int l = 0; l < n_levels; l++) {
            __syncthreads();
            if (threadIdx.x == 0) {
                // flattened chunk order of the level: classes from long to short runs (pref indexed by class)
                for (int k = 0; k < BBX_NCLS; k++) {
                    if (l == 0) s_head[k] = 0; else s_head[k] += s_cnt[k];
                    s_cnt[k] = (l < LC_CACHE) ? lc_cache[l * BBX_NCLS + k] : __ldcg(&lc[l * BBX_NCLS + k]);
                }
                int run = 0;
                for (int k = BBX_NCLS - 1; k >= 0; k--) { s_pref[k] = run; run += (s_cnt[k] + 15) >> 4; }     // 16 nodes per warp
                s_pref[BBX_NCLS] = run;
            }
            __syncthreads();
            const int total = s_pref[BBX_NCLS];
            const bool trace = (it == 1 && l == sp.gather && sp.gather > 0);       // debug: per-warp time inside one level
            unsigned long long t_in = 0; int my_chunks = 0;
            if (trace) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_in));
            // SOFTWARE PIPELINE over the warp's chunks of this level (they are independent of each other): while
            // chunk r is swept, the body, box constants and first record of chunk r+1 and the descriptor of chunk
            // r+2 are in flight, so only the first chunk of a level pays the gather latency.
            const int rounds = (total + W - 1) / W;
            auto desc_of = [&](int r) -> int4 {                  // (first M slot, constraints, gid A, gid B), m = 0 if none
                if (r >= rounds) return make_int4(0, 0, -1, -1);
                const int ch = r * W + ((r & 1) ? W - 1 - warp0 : warp0);
                if (ch >= total) return make_int4(0, 0, -1, -1);
                int k = BBX_NCLS - 1;
                #pragma unroll
                for (int kk = BBX_NCLS - 2; kk >= 0; kk--) if (ch >= s_pref[kk]) k = kk;
                const int idx = (ch - s_pref[k]) * 16 + (lane >> 1);
                if (idx >= s_cnt[k]) return make_int4(0, 0, -1, -1);
                return __ldcg(&dv.Nd_k[k][s_head[k] + idx]);
            };
            auto raw_of = [&](const int4& nd, PairRaw& o) {
                if (nd.y <= 0) return;
                const int g = side ? nd.w : nd.z;
                const float4* M = dv.M + (size_t)nd.x;
                o.c0 = __ldcg(M); o.c1 = __ldcg(M + (side ? 2 * PL : PL)); o.c3 = __ldcg(M + 3 * PL);
                if (g >= 0) { o.vw = ld256_cg(&dv.vel[2 * g]); if (g >= dv.NS) o.qi = ld256_nc(&dv.boxq[2 * (g - dv.NS)]); }
            };
            int4 nd0 = desc_of(0), nd1 = desc_of(1);
            PairRaw raw0;
            raw_of(nd0, raw0);
            for (int r = 0; r < rounds; r++) {
                const int4 nd2 = desc_of(r + 2);
                PairRaw raw1;
                raw_of(nd1, raw1);
                const int m = nd0.y;
                const int g = side ? nd0.w : nd0.z;
                const bool has = m > 0 && g >= 0;
                if (m > 0) my_chunks++;
                BodyState S;
                S.box = false;
                if (has) {
                    S.v = xyz(raw0.vw.a); S.w = xyz(raw0.vw.b); S.im = raw0.vw.a.w; S.ii = raw0.vw.b.w;
                    S.box = g >= dv.NS;
                    if (S.box) {
                        S.q = raw0.qi.a; S.iI = xyz(raw0.qi.b);
                        float3 u = f3(S.q.x, S.q.y, S.q.z);
                        S.k1 = S.q.w * S.q.w - v3dot(u, u); S.k2 = 2.0f * S.q.w;
                    }
                }
                int mmax = m;
                for (int o = 16; o > 0; o >>= 1) mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
                float4* M = dv.M + (size_t)nd0.x;
                float4 n0 = raw0.c0, n1 = raw0.c1, n3 = raw0.c3;
                for (int j = 0; j < mmax; j++) {
                    const bool on = j < m;
                    const unsigned int msk = __ballot_sync(0xffffffffu, on);
                    if (on) {
                        float4 c0 = n0, c1 = n1, c3 = n3;
                        if (j + 1 < m) { n0 = __ldcg(M + j + 1); n1 = __ldcg(M + j + 1 + (side ? 2 * PL : PL)); n3 = __ldcg(M + j + 1 + 3 * PL); }
                        const float mN = __shfl_sync(msk, c1.w, lane & ~1), mT0 = __shfl_sync(msk, c1.w, lane | 1);
                        solve_half(msk, side, has, c0, xyz(c1), mN, mT0, c3, S, sp);
                        if (!side) __stcg(M + j + 3 * PL, c3);
                    }
                }
                if (has) store_body(dv, g, S);
                nd0 = nd1; nd1 = nd2; raw0 = raw1;
            }
            if (trace && lane == 0) {
                unsigned long long t_out; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_out));
                unsigned long long d = t_out - t_in;
                unsigned long long* o = dv.level_ns + FLOW_DBG;
                atomicAdd(&o[0], d); atomicMax(&o[1], d); atomicAdd(&o[2], 1ull);
                if (my_chunks == 1) { atomicAdd(&o[3], d); atomicAdd(&o[4], 1ull); }
                else if (my_chunks == 2) { atomicAdd(&o[5], d); atomicAdd(&o[6], 1ull); }
                else if (my_chunks >= 3) atomicAdd(&o[7], d);
            }
            grid.sync();
        }
    }
}

// ---------------------------------------------------------------------------
// Batched small worlds: one WARP per world.
//
// Independent worlds never share a body, so a world can be swept by one warp with its body
// velocities in shared memory and __syncwarp() between levels: no grid barrier and no L2 round
// trip for body state.  A level of a 256-body world holds ~30 nodes, i.e. one warp's worth; eight
// worlds share a block and 24 worlds are resident per SM, so a 4096-world batch runs in about one
// wave with enough warps per scheduler to hide the instruction latency of the sweep (one block per
// world was measured slower than the grid-wide kernel: a single active warp per block).
// The schedule is the same order-preserving level schedule (Kahn on the same successor links), so
// results stay bit-identical to the single-world path and to the reference.  Impulses accumulate
// in place in the contact-order records L; a world's visiting order is written to its slice of the
// queue array and replayed by iterations 1..I-1.
// ---------------------------------------------------------------------------
#define WW_WARPS     8            // worlds per block
#define WW_BODY_MAX  512          // cap_s + cap_b of a world
#define WW_LMAX      512          // dependency levels of a world

__global__ void __launch_bounds__(256) k_frontier0_worlds(BbxDev dv, int* wq, int* wq_count)
{
    __shared__ int s_solver, s_nodes;
    if (threadIdx.x == 0) { s_solver = 0; s_nodes = 0; }
    __syncthreads();
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    int my_solver = 0, my_nodes = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        int4 id = dv.c_ids[c];
        if (RUN_K(id.z) != 0) continue;
        int4 l = dv.node[2 * (size_t)c];
        if (l.x < 0 && l.y < 0) continue;
        my_solver += l.z; my_nodes++;
        if (dv.indeg[c] == 0) {
            int w = id.x < dv.NS ? id.x / dv.cap_s : (id.x - dv.NS) / dv.cap_b;
            int slot = atomicAdd(&wq_count[w], 1);
            if (slot < dv.cap_contacts_world) wq[(size_t)w * dv.cap_contacts_world + slot] = c;
        }
    }
    if (my_nodes) { atomicAdd(&s_solver, my_solver); atomicAdd(&s_nodes, my_nodes); }
    __syncthreads();
    if (threadIdx.x == 0 && s_nodes) { atomicAdd(&dv.ctr->n_solver, s_solver); atomicAdd(&dv.ctr->n_nodes, s_nodes); }
}

__device__ __forceinline__ int ww_local(const BbxDev& dv, int w, int g)
{
    return g < dv.NS ? g - w * dv.cap_s : dv.cap_s + (g - dv.NS - w * dv.cap_b);
}

__device__ __forceinline__ void ww_load_body(const BbxDev& dv, const float4* vel, int w, int g, BodyState& s)
{
    int li = ww_local(dv, w, g);
    float4 l = vel[2 * li], a = vel[2 * li + 1];
    s.v = xyz(l); s.w = xyz(a); s.im = l.w; s.ii = a.w;
    s.box = g >= dv.NS;
    if (s.box) {
        int b = g - dv.NS;
        F8 qi = ld256_nc(&dv.boxq[2 * b]);
        s.q = qi.a; s.iI = xyz(qi.b);
        float3 u = f3(s.q.x, s.q.y, s.q.z);
        s.k1 = s.q.w * s.q.w - v3dot(u, u); s.k2 = 2.0f * s.q.w;
    }
}

__device__ __forceinline__ void ww_store_body(const BbxDev& dv, float4* vel, int w, int g, const BodyState& s)
{
    int li = ww_local(dv, w, g);
    vel[2 * li] = make_float4(s.v.x, s.v.y, s.v.z, s.im);
    vel[2 * li + 1] = make_float4(s.w.x, s.w.y, s.w.z, s.ii);
}

// sweep one node (all its constraints, in order) with bodies in shared memory, impulses in L
__device__ __forceinline__ void ww_sweep_node(const BbxDev& dv, float4* vel, int w, int c, int4 nd, const SolveParams& sp, bool wake)
{
    bool hasA = nd.x >= 0, hasB = nd.y >= 0;
    BodyState A, B;
    if (hasA) ww_load_body(dv, vel, w, nd.x, A);
    if (hasB) ww_load_body(dv, vel, w, nd.y, B);
    float4* L = dv.L + 4 * (size_t)c;
    F8 lo = ld256_cg(L), hi = ld256_cg(L + 2);
    for (int j = 0; j < nd.z; j++, L += 4) {
        float4 c0 = lo.a, c1 = lo.b, c2 = hi.a, c3 = hi.b;
        if (j + 1 < nd.z) { lo = ld256_cg(L + 4); hi = ld256_cg(L + 6); }       // next record in flight
        solve_constraint(c0, c1, c2, c3, hasA, hasB, A, B, sp);
        __stcg(L + 3, c3);
    }
    if (hasA) { ww_store_body(dv, vel, w, nd.x, A); if (wake) wake_body(dv, nd.x); }
    if (hasB) { ww_store_body(dv, vel, w, nd.y, B); if (wake) wake_body(dv, nd.y); }
}

// the same sweep with the node's first record already in registers (replay: it was requested one step ahead)
__device__ __forceinline__ void ww_sweep_node_pre(const BbxDev& dv, float4* vel, int w, int c, int gA, int gB, int m, F8 lo, F8 hi,
                                                  const SolveParams& sp)
{
    bool hasA = gA >= 0, hasB = gB >= 0;
    BodyState A, B;
    if (hasA) ww_load_body(dv, vel, w, gA, A);
    if (hasB) ww_load_body(dv, vel, w, gB, B);
    float4* L = dv.L + 4 * (size_t)c;
    for (int j = 0; j < m; j++, L += 4) {
        float4 c0 = lo.a, c1 = lo.b, c2 = hi.a, c3 = hi.b;
        if (j + 1 < m) { lo = ld256_cg(L + 4); hi = ld256_cg(L + 6); }       // next record in flight
        solve_constraint(c0, c1, c2, c3, hasA, hasB, A, B, sp);
        __stcg(L + 3, c3);
    }
    if (hasA) ww_store_body(dv, vel, w, gA, A);
    if (hasB) ww_store_body(dv, vel, w, gB, B);
}

// dynamic shared memory: per warp [ vel: 2*nb float4 | steps: WW_LMAX+1 ints | next_count, bad ]
__global__ void __launch_bounds__(WW_WARPS * 32, 2) k_solve_worlds(BbxDev dv, SolveParams sp, int* wq, const int* wq_count, int warp_bytes)
{
    extern __shared__ unsigned char smem_raw[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * WW_WARPS + wib;
    if (w >= dv.n_worlds) return;
    const int nb = dv.cap_s + dv.cap_b;
    unsigned char* base = smem_raw + (size_t)wib * warp_bytes;
    float4* vel = reinterpret_cast<float4*>(base);
    // steps of one sweep: 32 consecutive queue slots of one level each, entry = first slot | (count - 1) << 24
    int* steps = reinterpret_cast<int*>(base + (size_t)nb * 32);
    int* next_count = steps + WW_LMAX + 1;
    int n0 = min(wq_count[w], dv.cap_contacts_world);
    if (n0 == 0) return;                                              // no dynamic contact in this world
    int* queue = wq + (size_t)w * dv.cap_contacts_world;              // this world's visiting order
    const int qcap = dv.cap_contacts_world;
    // ---- stage the world's velocities ----
    for (int i = lane; i < nb; i += 32) {
        int g = i < dv.cap_s ? w * dv.cap_s + i : dv.NS + w * dv.cap_b + (i - dv.cap_s);
        F8 v = ld256_cg(&dv.vel[2 * g]);
        vel[2 * i] = v.a; vel[2 * i + 1] = v.b;
    }
    if (lane == 0) next_count[0] = 0;
    __syncwarp();
    // node descriptors in visiting order (contact, dynamic gid A, dynamic gid B, run length): the replay streams them
    int4* wd = dv.Nd_k[0] + (size_t)w * dv.cap_contacts_world;
    // ---- iteration 0: Kahn levels ----
    int head = 0, tail = n0, lvl = 0, ns = 0;
    bool bad = false;
    while (head < tail) {
        for (int p0 = head; p0 < tail; p0 += 32) {
            int p = p0 + lane;
            if (lane == 0 && ns < WW_LMAX) steps[ns] = p0 | ((min(32, tail - p0) - 1) << 24);
            ns++;
            if (p < tail) {
                int c = __ldcg(&queue[p]);
                F8 rec = ld256_nc((const float4*)&dv.node[2 * (size_t)c]);
                int4 nd = make_int4(__float_as_int(rec.a.x), __float_as_int(rec.a.y), __float_as_int(rec.a.z), 0);
                int sa = __float_as_int(rec.b.x), sb = __float_as_int(rec.b.y);
                __stcg(&wd[p], make_int4(c, nd.x, nd.y, nd.z));
                ww_sweep_node(dv, vel, w, c, nd, sp, true);
                if (sa >= 0 && atomicSub(&dv.indeg[sa & 0x0fffffff], 1) == 1) {
                    int q = tail + atomicAdd(next_count, 1);
                    if (q < qcap) __stcg(&queue[q], sa & 0x0fffffff);
                }
                if (sb >= 0 && atomicSub(&dv.indeg[sb & 0x0fffffff], 1) == 1) {
                    int q = tail + atomicAdd(next_count, 1);
                    if (q < qcap) __stcg(&queue[q], sb & 0x0fffffff);
                }
            }
        }
        __syncwarp();
        int pushed = next_count[0];
        __syncwarp();
        head = tail; tail += pushed; lvl++;
        if (tail > qcap || ns > WW_LMAX) { bad = true; break; }
        if (lane == 0) next_count[0] = 0;
        __syncwarp();
    }
    if (bad) { if (lane == 0) atomicOr(&dv.ctr->overflow, 32); return; }
    const int n_levels = lvl;
    if (lane == 0) atomicMax(&dv.ctr->n_levels, n_levels);
    // ---- iterations 1..I-1: replay the visiting order ----
    // One flattened sequence of steps; the warp is alone on its world, so what hides the latency of the record
    // gather is software pipelining: the descriptor of step j+2 and the first record of step j+1 are in flight while
    // step j is swept (a step never spans two levels; __syncwarp separates steps, which is all a level boundary needs).
    const int K = ns, total = (sp.iterations - 1) * K;
    auto desc_at = [&](int j) -> int4 {
        if (j >= total) return make_int4(-1, -1, -1, 0);
        int e = steps[j % K];
        return lane <= (e >> 24) ? __ldcg(&wd[(e & 0xffffff) + lane]) : make_int4(-1, -1, -1, 0);
    };
    int4 d0 = desc_at(0), d1 = desc_at(1);
    F8 lo0, hi0;
    lo0.a = lo0.b = hi0.a = hi0.b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (d0.x >= 0) { const float4* L = dv.L + 4 * (size_t)d0.x; lo0 = ld256_cg(L); hi0 = ld256_cg(L + 2); }
    // (with a single step per sweep the next step's record is the one this step is about to update: no early load)
    const bool early = K >= 2;
    for (int j = 0; j < total; j++) {
        int4 d2 = desc_at(j + 2);
        F8 lo1 = lo0, hi1 = hi0;
        if (early && d1.x >= 0) { const float4* L = dv.L + 4 * (size_t)d1.x; lo1 = ld256_cg(L); hi1 = ld256_cg(L + 2); }
        if (d0.x >= 0) ww_sweep_node_pre(dv, vel, w, d0.x, d0.y, d0.z, d0.w, lo0, hi0, sp);
        __syncwarp();
        if (!early && d1.x >= 0) { const float4* L = dv.L + 4 * (size_t)d1.x; lo1 = ld256_cg(L); hi1 = ld256_cg(L + 2); }
        d0 = d1; d1 = d2; lo0 = lo1; hi0 = hi1;
    }
    // ---- write the velocities back ----
    for (int i = lane; i < nb; i += 32) {
        int g = i < dv.cap_s ? w * dv.cap_s + i : dv.NS + w * dv.cap_b + (i - dv.cap_s);
        st256_cg(&dv.vel[2 * g], vel[2 * i], vel[2 * i + 1]);
    }
}

static int world_warp_bytes(const BbxDev* d) { return (d->cap_s + d->cap_b) * 32 + (WW_LMAX + 1 + 3) * 4; }

static bool use_world_warps(const BbxDev* d, int imported)
{
    return !imported && d->n_worlds >= 16 && d->cap_s + d->cap_b <= WW_BODY_MAX && d->cap_s > 0 && d->cap_b > 0 &&
           d->cap_contacts_world < (1 << 24) &&
           WW_WARPS * world_warp_bytes(d) <= 220 * 1024;
}

// ---------------------------------------------------------------------------
// Batched small worlds, second form: one BLOCK per GROUP of worlds, one thread per constraint.
//
// A 256-body world offers ~20 independent constraints per dependency level: a warp per world keeps a
// third of its lanes busy (and sweeps a manifold of m contacts in m dependent rounds).  Here a block of
// 256 threads owns a group of worlds (as many as it takes to run the whole batch in one wave) and sweeps
// level l of all of them together: every constraint is its own node (chains per constraint, as in the flow
// solver), so the lanes of a warp do equal work, levels are separated by __syncthreads() and body
// velocities stay in the packed global array (L2), which a block-wide barrier keeps coherent.
// Iteration 0 discovers the group's levels (Kahn) and writes descriptors and records in visiting order;
// iterations 1..I-1 stream them.  Same per-body constraint order as everywhere else: bit-identical results.
// ---------------------------------------------------------------------------
#define GRP_THREADS 128
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
#define GRP_BLOCKS_PER_SM 6
#define GRP_LMAX 2047                  // dependency levels of a group

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

__global__ void __launch_bounds__(GRP_THREADS, GRP_BLOCKS_PER_SM) k_solve_groups(BbxDev dv, SolveParams sp, int* gq, const int* gq_count, int gcap)
{
    __shared__ int level_start[GRP_LMAX + 2];
    __shared__ int s_next;
    const int g = blockIdx.x, tid = threadIdx.x;
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
    if (tid == 0) { s_next = 0; level_start[0] = 0; }
    __syncthreads();
    const bool big = dv.ctr->n_contacts > BBX_L2_STREAM_CONTACTS;     // bodies stay in L2, records stream
    const unsigned long long pol_keep = big ? l2_policy_keep() : l2_policy_normal(), pol_stream = big ? l2_policy_stream() : l2_policy_normal();
    // ---- iteration 0: Kahn levels ----
    int head = 0, tail = n0, lvl = 0;
    bool bad = false;
    while (head < tail) {
        for (int p = head + tid; p < tail; p += GRP_THREADS) {
            const int c = __ldcg(&queue[p]);
            F8 nd = ld256_nc_pol((const float4*)&dv.node[2 * (size_t)c], pol_stream);    // (gid A, pos, len, successor on A) (same for B)
            const int gA = __float_as_int(nd.a.x), gB = __float_as_int(nd.b.x);
            const int sa = __float_as_int(nd.a.w), sb = __float_as_int(nd.b.w);
            const float4* L = dv.L + 4 * (size_t)c;
            F8 lo = ld256_nc_pol(L, pol_stream), hi = ld256_nc_pol(L + 2, pol_stream);
            float4 c0 = lo.a, c1 = lo.b, c2 = hi.a, c3 = hi.b;
            const bool hasA = gA >= 0, hasB = gB >= 0;
            BodyState A, B;
            if (hasA) load_body_keep(dv, gA, A, pol_keep);
            if (hasB) load_body_keep(dv, gB, B, pol_keep);
            solve_constraint(c0, c1, c2, c3, hasA, hasB, A, B, sp);
            if (hasA) store_body_keep(dv, gA, A, pol_keep);
            if (hasB) store_body_keep(dv, gB, B, pol_keep);
            float4* M = dv.M + base + p;
            st128_cg_pol(M, c0, pol_stream); st128_cg_pol(M + PL, c1, pol_stream); st128_cg_pol(M + 2 * PL, c2, pol_stream); st128_cg_pol(M + 3 * PL, c3, pol_stream);
            __stcg(&gd[p], make_int4(gA, gB, c, 0));
            dv.slot_of[c] = (int)(base + p);
            if (sa >= 0 && atomicSub(&dv.indeg[sa], 1) == 1) { int q = tail + atomicAdd(&s_next, 1); if (q < glim) __stcg(&queue[q], sa); }
            if (sb >= 0 && atomicSub(&dv.indeg[sb], 1) == 1) { int q = tail + atomicAdd(&s_next, 1); if (q < glim) __stcg(&queue[q], sb); }
        }
        __syncthreads();
        const int pushed = s_next;
        __syncthreads();
        head = tail; tail += pushed; lvl++;
        if (tail > glim || lvl > GRP_LMAX) { bad = true; break; }
        if (tid == 0) { s_next = 0; level_start[lvl] = head; }
        __syncthreads();
    }
    if (bad) { if (tid == 0) atomicOr(&dv.ctr->overflow, 32); return; }
    const int n_levels = lvl;
    if (tid == 0) { atomicMax(&dv.ctr->n_levels, n_levels); level_start[n_levels] = tail; }
    __syncthreads();
    // ---- iterations 1..I-1: stream the visiting order ----
    // The descriptor of this thread's first constraint of the NEXT level is fetched while the current level is
    // swept (descriptors are static), so that after the barrier the body loads can be issued at once; the next
    // level's records are pulled into L2 at the same time (prefetch, no registers).
    int2 dn = make_int2(-1, -1);
    { const int pn = level_start[0] + tid; if (pn < level_start[1]) { int4 t4 = ld128i_cg_pol(&gd[pn], pol_stream); dn = make_int2(t4.x, t4.y); } }
    for (int it = 1; it < sp.iterations; it++) {
        for (int l = 0; l < n_levels; l++) {
            const int s0 = level_start[l], s1 = level_start[l + 1];
            const int2 dc = dn;
            {
                const int ln = (l + 1 < n_levels) ? l + 1 : 0;
                const int pn = level_start[ln] + tid;
                dn = make_int2(-1, -1);
                if (pn < level_start[ln + 1]) {
                    int4 t4 = ld128i_cg_pol(&gd[pn], pol_stream); dn = make_int2(t4.x, t4.y);
                    const float4* Mn = dv.M + base + pn;
                    prefetch_l2(Mn); prefetch_l2(Mn + PL); prefetch_l2(Mn + 2 * PL);
                    if (n_levels > 1) prefetch_l2(Mn + 3 * PL);
                }
            }
            for (int p = s0 + tid; p < s1; p += GRP_THREADS) {
                int2 d = dc;
                if (p != s0 + tid) { int4 t4 = ld128i_cg_pol(&gd[p], pol_stream); d = make_int2(t4.x, t4.y); }      // (levels wider than the block)
                float4* M = dv.M + base + p;
                float4 c0 = ld128_cg_pol(M, pol_stream), c1 = ld128_cg_pol(M + PL, pol_stream), c2 = ld128_cg_pol(M + 2 * PL, pol_stream), c3 = ld128_cg_pol(M + 3 * PL, pol_stream);
                const bool hasA = d.x >= 0, hasB = d.y >= 0;
                BodyState A, B;
                if (hasA) load_body_keep(dv, d.x, A, pol_keep);
                if (hasB) load_body_keep(dv, d.y, B, pol_keep);
                solve_constraint(c0, c1, c2, c3, hasA, hasB, A, B, sp);
                if (hasA) store_body_keep(dv, d.x, A, pol_keep);
                if (hasB) store_body_keep(dv, d.y, B, pol_keep);
                st128_cg_pol(M + 3 * PL, c3, pol_stream);
            }
            __syncthreads();
        }
    }
}

// The same group solver over the UNROLLED dependency graph: a constraint's visit in sweep it+1 only waits for the
// previous visits of its own two bodies, not for the end of sweep it, so the sweeps overlap wherever the graph allows
// it (measured with the flow solver on 4096 x 256 bodies: all 8 sweeps together are 285 constraints deep, against
// 8 x 56 = 448 level barriers one sweep after the other).  Every body's chain is circular (k_adj_link, ADJ_WRAP); the
// in-degree counter of a constraint is re-armed when a visit fires, and the visit that takes it to zero appends the
// constraint to the next wavefront.  Wavefronts are separated by __syncthreads(); nothing spins.  Records stay in
// contact order (L), impulses included.  Per-body order = the reference's, bit for bit.
__global__ void __launch_bounds__(GRP_THREADS, GRP_BLOCKS_PER_SM) k_solve_groups_unrolled(BbxDev dv, SolveParams sp, const int* gq, const int* gq_count, int gcap)
{
    __shared__ int s_next;
    const int g = blockIdx.x, tid = threadIdx.x;
    const int glim = (int)min((long long)gcap, (long long)dv.cap_contacts - (long long)g * gcap);      // ragged last group, see k_solve_groups
    if (gq_count[g] > glim) { if (tid == 0) atomicOr(&dv.ctr->overflow, 32); return; }
    int n_cur = gq_count[g];
    if (n_cur == 0) return;                                           // no dynamic contact in this group
    int2* cur = reinterpret_cast<int2*>(dv.Nd_k[0] + (size_t)g * gcap);   // two wavefront lists of (constraint, sweep) in the group's
    int2* nxt = cur + glim;                                               // descriptor slice (int4 per slot = room for both)
    for (int p = tid; p < n_cur; p += GRP_THREADS) cur[p] = make_int2(gq[(size_t)g * gcap + p], 0);
    if (tid == 0) s_next = 0;
    __syncthreads();
    int waves = 0;
    bool bad = false;
    while (n_cur > 0) {
        for (int p = tid; p < n_cur; p += GRP_THREADS) {
            const int2 e = __ldcg(&cur[p]);
            const int c = e.x, it = e.y;
            F8 nd = ld256_nc((const float4*)&dv.node[2 * (size_t)c]);    // (gid A, pos, len, successor on A) (same for B)
            const int gA = __float_as_int(nd.a.x), gB = __float_as_int(nd.b.x);
            const int sa = __float_as_int(nd.a.w), sb = __float_as_int(nd.b.w);
            float4* L = dv.L + 4 * (size_t)c;
            F8 lo = ld256_nc(L), hi = ld256_cg(L + 2);                    // the impulses (second half) change from visit to visit
            float4 c0 = lo.a, c1 = lo.b, c2 = hi.a, c3 = hi.b;
            const bool hasA = gA >= 0, hasB = gB >= 0;
            BodyState A, B;
            if (hasA) load_body(dv, gA, A);
            if (hasB) load_body(dv, gB, B);
            solve_constraint(c0, c1, c2, c3, hasA, hasB, A, B, sp);
            if (hasA) store_body(dv, gA, A);
            if (hasB) store_body(dv, gB, B);
            __stcg(L + 3, c3);
            __stcg(&dv.indeg[c], (hasA ? 1 : 0) + (hasB ? 1 : 0));        // re-armed for the next visit before anyone can signal it
            #pragma unroll
            for (int side = 0; side < 2; side++) {
                const int s = side ? sb : sa;
                if (s < 0) continue;
                const int tgt = s & (ADJ_WRAP - 1), tit = (s & ADJ_WRAP) ? it + 1 : it;
                if (tit < sp.iterations && atomicSub(&dv.indeg[tgt], 1) == 1) {
                    const int q = atomicAdd(&s_next, 1);
                    if (q < glim) __stcg(&nxt[q], make_int2(tgt, tit));
                }
            }
        }
        __syncthreads();
        const int pushed = s_next;
        __syncthreads();
        if (tid == 0) s_next = 0;
        if (pushed > glim) { bad = true; break; }
        int2* t = cur; cur = nxt; nxt = t;
        n_cur = pushed; waves++;
        __syncthreads();
    }
    if (tid == 0) {
        if (bad) atomicOr(&dv.ctr->overflow, 32);
        atomicMax(&dv.ctr->n_levels, waves);
    }
}

static bool use_world_groups(const BbxDev* d, int imported, int schedule)
{
    return !imported && d->n_worlds >= 16 && schedule == BBX_SOLVER_EXACT && !d->no_groups;
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

// touched-body statistics and sanity for the host
int bbx_stage_solve(BbxDev* d, const BbxStepParams* p) { return bbx_stage_solve_ex(d, p, 0, p->settings.solver_iterations, 0); }

// presolve_mode: see k_presolve; iterations: Gauss-Seidel sweeps; imported: the contact list came from
// the host in array order (public stage functions), so the per-body order is the array index
int bbx_stage_solve_ex(BbxDev* d, const BbxStepParams* p, int presolve_mode, int iterations, int imported)
{
    const int G = BBX_NUM_SMS * 8;
    // opt-in (BBX_SOLVER_EXACT_FLOW or BBX_SOLVER_FLOW=1): barrier-free flow solver, every constraint its own node
    const int flow = (p->solver_schedule == BBX_SOLVER_EXACT_FLOW || (d->force_flow && p->solver_schedule == BBX_SOLVER_EXACT)) ? 1 : 0;
    const int groups = (!flow && use_world_groups(d, imported, p->solver_schedule)) ? 1 : 0;
    const int per_contact = flow | groups;                 // every constraint is its own node
    BBX_CUDA(d, cudaMemsetAsync(d->deg, 0, sizeof(int) * ((size_t)d->N + 1), d->stream));
    const float warm = (presolve_mode == 0 && !imported && p->warm_start > 0.0f) ? p->warm_start : 0.0f;
    if (warm > 0.0f) { int rcw = bbx_warm_ensure(d); if (rcw) return rcw; }
    // default path: level-scheduled kernel (single world, or batches the world solvers do not take)
    const int levels_path = (!flow && !groups && !(use_world_warps(d, imported) && p->solver_schedule != BBX_SOLVER_EXACT_FLOW)) ? 1 : 0;
    BBX_LAUNCH(d, k_body_records, bbx_blocks(d->N, 256), 256, 0, *d);
    BBX_LAUNCH(d, k_presolve, G, 256, 0, *d, p->settings.baumgarte_bias, p->settings.allowed_penetration, p->dt, presolve_mode, per_contact, warm, (levels_path && iterations > 0) ? 1 : 0);
    if (iterations <= 0) { bbx_timing_mark(d, 4); BBX_CUDA(d, cudaGetLastError()); return 0; }    // :1223 loop does not run
    int rc = bbx_scan_exclusive(d, d->deg, d->adj_start, d->adj_cursor, d->N + 1, NULL, NULL);
    if (rc) return rc;
    BBX_LAUNCH(d, k_adj_fill, G, 256, 0, *d, imported ? 2 : (p->solver_schedule == BBX_SOLVER_COLOURED ? 1 : 0), per_contact);
    if (levels_path) BBX_CUDA(d, cudaMemsetAsync(d->level_count, 0, 2 * BBX_NCLS * sizeof(int), d->stream));
    BBX_LAUNCH(d, k_adj_link, bbx_blocks(d->N, 128), 128, 0, *d, per_contact, flow, warm > 0.0f ? 1 : 0, levels_path, (groups && d->grp_unrolled) ? 1 : 0);
    if (warm > 0.0f) BBX_LAUNCH(d, k_warm_apply, bbx_blocks(d->N, 128), 128, 0, *d, per_contact, flow, p->settings.friction);
    d->last_flow = flow; d->last_per_contact = per_contact;
    if (flow) {
        BBX_LAUNCH(d, k_frontier_flow, G, 256, 0, *d);
        if (d->flow_blocks == 0) {
            int per_sm = 0, sms = 0, coop = 0;
            BBX_CUDA(d, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d->device));
            if (!coop) return bbx_fail(d, BBX_ERR_NO_DEVICE, "device lacks cooperative launch", __FILE__, __LINE__);
            BBX_CUDA(d, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d->device));
            BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_flow, FLOW_THREADS, 0));
            if (per_sm < 1) return bbx_fail(d, BBX_ERR_NO_DEVICE, "flow solver kernel does not fit on an SM", __FILE__, __LINE__);
            if (per_sm > FLOW_BLOCKS_PER_SM) per_sm = FLOW_BLOCKS_PER_SM;
            d->flow_blocks = per_sm * sms;
            if (d->tune_flow_blocks > 0 && d->tune_flow_blocks < d->flow_blocks) d->flow_blocks = d->tune_flow_blocks;
        }
        bbx_timing_mark(d, 4);
        SolveParams spf;
        spf.restitution = p->settings.restitution; spf.friction = p->settings.friction;
        spf.vel_threshold = p->settings.velocity_threshold; spf.iterations = iterations;
        spf.gather = d->tune_flow_gather;
        BbxDev dvf = *d;
        void* fargs[] = { (void*)&dvf, (void*)&spf };
        // cooperative launch: every block must be resident (the waits rely on it) and iteration 0 ends with a grid barrier
        BBX_CUDA(d, cudaLaunchCooperativeKernel((void*)k_solve_flow, dim3(d->flow_blocks), dim3(FLOW_THREADS), fargs, 0, d->stream));
        d->launches++;
        BBX_CUDA(d, cudaGetLastError());
        return 0;
    }
    if (groups) {
        // batched small worlds: a block per group of worlds, sized so that the batch runs in one wave
        if (d->grp_blocks == 0) {
            int per_sm = 0, sms = 0;
            BBX_CUDA(d, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d->device));
            BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_groups, GRP_THREADS, 0));
            if (per_sm < 1) return bbx_fail(d, BBX_ERR_NO_DEVICE, "group solver kernel does not fit on an SM", __FILE__, __LINE__);
            d->grp_blocks = per_sm * sms;
        }
        int wpb = (d->n_worlds + d->grp_blocks - 1) / d->grp_blocks;
        if (wpb < 1) wpb = 1;
        const int n_groups = (d->n_worlds + wpb - 1) / wpb;
        long long gcap_ll = (long long)wpb * d->cap_contacts_world;
        if (gcap_ll > d->cap_contacts) gcap_ll = d->cap_contacts;
        const int gcap = (int)gcap_ll;
        int* gq = d->queue_k[0];                       // capacity cap_contacts
        int* gq_count = d->adj_cursor;                 // free again after k_adj_link; N + 1 >= n_groups ints
        BBX_CUDA(d, cudaMemsetAsync(gq_count, 0, sizeof(int) * (size_t)n_groups, d->stream));
        BBX_LAUNCH(d, k_frontier0_groups, G, 256, 0, *d, gq, gq_count, wpb, gcap);
        bbx_timing_mark(d, 4);
        SolveParams spg;
        spg.restitution = p->settings.restitution; spg.friction = p->settings.friction;
        spg.vel_threshold = p->settings.velocity_threshold; spg.iterations = iterations; spg.gather = 0;
        if (d->grp_unrolled) BBX_LAUNCH(d, k_solve_groups_unrolled, n_groups, GRP_THREADS, 0, *d, spg, gq, gq_count, gcap);
        else                 BBX_LAUNCH(d, k_solve_groups, n_groups, GRP_THREADS, 0, *d, spg, gq, gq_count, gcap);
        BBX_CUDA(d, cudaGetLastError());
        return 0;
    }
    if (use_world_warps(d, imported) && p->solver_schedule != BBX_SOLVER_EXACT_FLOW) {
        // batched small worlds: per-world frontier, then one warp per world (no grid barrier)
        int* wq = d->queue_k[0];                       // capacity cap_contacts = n_worlds * cap_contacts_world
        int* wq_count = d->adj_cursor;                 // free again after k_adj_link; N + 1 >= n_worlds ints
        BBX_CUDA(d, cudaMemsetAsync(wq_count, 0, sizeof(int) * (size_t)d->n_worlds, d->stream));
        BBX_LAUNCH(d, k_frontier0_worlds, G, 256, 0, *d, wq, wq_count);
        bbx_timing_mark(d, 4);
        SolveParams spw;
        spw.restitution = p->settings.restitution; spw.friction = p->settings.friction;
        spw.vel_threshold = p->settings.velocity_threshold; spw.iterations = iterations;
        int wbytes = world_warp_bytes(d), smem = WW_WARPS * wbytes;
        if (!d->wb_attr_set) {
            BBX_CUDA(d, cudaFuncSetAttribute(k_solve_worlds, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            d->wb_attr_set = 1;
        }
        BBX_LAUNCH(d, k_solve_worlds, (d->n_worlds + WW_WARPS - 1) / WW_WARPS, WW_WARPS * 32, smem, *d, spw, wq, wq_count, wbytes);
        BBX_CUDA(d, cudaGetLastError());
        return 0;
    }
    // (level 0 of the class queues was filled by k_adj_link)

    if (d->coop_blocks == 0) {
        int per_sm = 0, sms = 0, coop = 0;
        BBX_CUDA(d, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d->device));
        if (!coop) return bbx_fail(d, BBX_ERR_NO_DEVICE, "device lacks cooperative launch", __FILE__, __LINE__);
        BBX_CUDA(d, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d->device));
        BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, d->packed_sweep ? k_solve_levels<true> : k_solve_levels<false>, SOLVE_THREADS, 0));
        if (per_sm < 1) return bbx_fail(d, BBX_ERR_NO_DEVICE, "solver kernel does not fit on an SM", __FILE__, __LINE__);
        if (per_sm > SOLVE_BLOCKS_PER_SM) per_sm = SOLVE_BLOCKS_PER_SM;
        d->coop_blocks = per_sm * sms;
        if (d->tune_solve_blocks > 0 && d->tune_solve_blocks < d->coop_blocks) d->coop_blocks = d->tune_solve_blocks;
    }
    if (d->rp_blocks == 0) {
        int per_sm = 0, sms = 0;
        BBX_CUDA(d, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d->device));
        BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_replay_pairs, RP_THREADS, 0));
        if (per_sm < 1) return bbx_fail(d, BBX_ERR_NO_DEVICE, "replay kernel does not fit on an SM", __FILE__, __LINE__);
        if (per_sm > RP_BLOCKS_PER_SM) per_sm = RP_BLOCKS_PER_SM;
        d->rp_blocks = per_sm * sms;
        // measured neutral at 1M bodies (5.55 vs 5.48 ms: the fat levels are bound by their DRAM bursts, not by the
        // depth of the sweep; profiles/r01_notes.md), so the one-lane replay stays the default
        // 0 (default): all iterations in k_solve_levels; 2: iterations 1..I-1 in k_replay_pairs (measured equal, 5.45 ms)
        if (d->tune_replay_blocks > 0 && d->tune_replay_blocks < d->rp_blocks) d->rp_blocks = d->tune_replay_blocks;
    }
    if (d->kahn_blocks == 0) {
        int per_sm = 0, sms = 0;
        BBX_CUDA(d, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d->device));
        BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_kahn_levels, KAHN_THREADS, 0));
        if (per_sm > KAHN_BLOCKS_PER_SM) per_sm = KAHN_BLOCKS_PER_SM;
        d->kahn_blocks = per_sm * sms;                 // 0: does not fit, the fused iteration 0 is used
    }
    SolveParams sp;
    sp.restitution = p->settings.restitution; sp.friction = p->settings.friction;
    sp.vel_threshold = p->settings.velocity_threshold; sp.iterations = iterations; sp.gather = 0;
    sp.levels_known = (d->kahn_blocks > 0 && !d->fused_kahn) ? 1 : 0;
    sp.packed = d->packed_sweep;
    {   // f32x2 constants of the packed sweep, low half first
        const unsigned long long n0 = 0x80000000ull, one = 0x3f800000ull, mone = 0xbf800000ull, two = 0x40000000ull;
        sp.pc.neg0 = n0 | (n0 << 32); sp.pc.one = one | (one << 32); sp.pc.mone = mone | (mone << 32);
        sp.pc.two = two | (two << 32); sp.pc.sgn = mone | (one << 32);
    }
    if (sp.levels_known) {
        BbxDev dvk = *d;
        void* kargs[] = { (void*)&dvk };
        BBX_CUDA(d, cudaLaunchCooperativeKernel((void*)k_kahn_levels, dim3(d->kahn_blocks), dim3(KAHN_THREADS), kargs, 0, d->stream));
        d->launches++;
    }
    bbx_timing_mark(d, 4);
    if (d->trace_level > 0) { sp.gather = d->trace_level; BBX_CUDA(d, cudaMemsetAsync(d->level_ns + FLOW_DBG, 0, 8 * sizeof(unsigned long long), d->stream)); }
    // iteration 0 (level discovery) on one lane per node; the later iterations with two lanes per node
    sp.first_only = (iterations > 1 && d->replay_mode != 0 && !sp.levels_known) ? 1 : 0;
    const bool pairs_all = d->replay_mode == 2 && sp.levels_known;      // every sweep in k_replay_pairs
    BbxDev dv = *d;
    void* args[] = { (void*)&dv, (void*)&sp };
    if (!pairs_all) {
        BBX_CUDA(d, cudaLaunchCooperativeKernel(d->packed_sweep ? (void*)k_solve_levels<true> : (void*)k_solve_levels<false>, dim3(d->coop_blocks), dim3(SOLVE_THREADS), args, 0, d->stream));
        d->launches++;
    }
    if ((sp.first_only && d->replay_mode == 2) || pairs_all) {
        BBX_CUDA(d, cudaLaunchCooperativeKernel((void*)k_replay_pairs, dim3(d->rp_blocks), dim3(RP_THREADS), args, 0, d->stream));
        d->launches++;
    }
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}
