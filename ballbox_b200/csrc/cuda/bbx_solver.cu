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
// Levels are discovered on the fly during iteration 0 (Kahn's algorithm on the
// per-body successor links); the visiting order is recorded and the constraint
// records are rewritten in that level-major order, so iterations 1..I-1 stream
// them with coalesced float4 loads.  One cooperative persistent kernel runs all
// iterations; levels are separated by grid-wide barriers, not kernel launches.
#include <cooperative_groups.h>
#include "bbx_dev.cuh"
namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------
// K9 presolve: contact -> solver record (contact order)
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

__global__ void __launch_bounds__(256) k_presolve(BbxDev dv, float baumgarte, float allowed, float dt)
{
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        int4 id = dv.c_ids[c];
        float4 pt4 = dv.c_pt[c];
        float3 pt = xyz(pt4), nrm = xyz(dv.c_nrm[c]);
        int a = id.x, b = id.y;
        float3 rA, rB, iA, iB; float imA, imB; bool dynA, dynB;
        {
            unsigned char f = dv.flags[a];
            dynA = !(f & BF_STATIC);
            rA = v3sub(pt, xyz(dv.pos[a]));
            imA = dv.lin[a].w;                      // 0 for static bodies (:1266)
            if (a < dv.NS) { float ii = dv.ang[a].w; iA = f3(ii, ii, ii); }
            else iA = xyz(dv.invI[a - dv.NS]);
        }
        if (b >= 0) {
            unsigned char f = dv.flags[b];
            dynB = !(f & BF_STATIC);
            rB = v3sub(pt, xyz(dv.pos[b]));
            imB = dv.lin[b].w;
            if (b < dv.NS) { float ii = dv.ang[b].w; iB = f3(ii, ii, ii); }
            else iB = xyz(dv.invI[b - dv.NS]);
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
        dv.L0[c] = make_float4(nrm.x, nrm.y, nrm.z, bias);
        dv.L1[c] = make_float4(rA.x, rA.y, rA.z, mN);
        dv.L2[c] = make_float4(rB.x, rB.y, rB.z, mT0);
        dv.L3[c] = make_float4(0.0f, 0.0f, 0.0f, mT1);                                  // impulses start at 0 (:1203-1206)
        dv.Lid[c] = make_int2(dynA ? a : -1, dynB ? b : -1);
        dv.indeg[c] = (dynA ? 1 : 0) + (dynB ? 1 : 0);
        dv.next_a[c] = -1; dv.next_b[c] = -1;
        dv.slot_of[c] = -1;
        if (dynA) atomicAdd(&dv.deg[a], 1);
        if (dynB) atomicAdd(&dv.deg[b], 1);
    }
}

// ---------------------------------------------------------------------------
// K10 dependency graph: per-body ordered constraint lists -> successor links
// ---------------------------------------------------------------------------
// order of a constraint inside ONE body's list: (phase, partner index, point)
__device__ __forceinline__ unsigned int body_order_key(const BbxDev& dv, int4 id, bool for_a)
{
    int other = for_a ? id.y : id.x;
    unsigned int partner = 0;
    if (other >= 0) partner = (unsigned int)(other < dv.NS ? other : other - dv.NS);
    return ((unsigned int)id.w << 29) | (partner << 5) | (unsigned int)id.z;
}

__global__ void __launch_bounds__(256) k_adj_fill(BbxDev dv)
{
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        int2 l = dv.Lid[c];
        if (l.x < 0 && l.y < 0) continue;
        int4 id = dv.c_ids[c];
        if (l.x >= 0) {
            int slot = atomicAdd(&dv.adj_cursor[l.x], 1);
            dv.adj[slot] = ((unsigned long long)body_order_key(dv, id, true) << 32) | (unsigned int)c;
        }
        if (l.y >= 0) {
            int slot = atomicAdd(&dv.adj_cursor[l.y], 1);
            dv.adj[slot] = ((unsigned long long)body_order_key(dv, id, false) << 32) | (unsigned int)c;
        }
    }
}

__global__ void __launch_bounds__(128) k_adj_link(BbxDev dv)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    bool has = false;
    if (x < dv.N) {
        int s = dv.adj_start[x], e = dv.adj_start[x + 1];
        int d = e - s;
        has = d > 0;
        if (has) {
            unsigned long long* lst = dv.adj + s;
            // insertion sort: lists are short (a body touches a handful of neighbours)
            for (int i = 1; i < d; i++) {
                unsigned long long v = lst[i];
                int j = i - 1;
                while (j >= 0 && lst[j] > v) { lst[j + 1] = lst[j]; j--; }
                lst[j + 1] = v;
            }
            int prev = (int)(unsigned int)lst[0];
            atomicSub(&dv.indeg[prev], 1);                 // first constraint of this body has no predecessor here
            for (int i = 1; i < d; i++) {
                int cur = (int)(unsigned int)lst[i];
                if (dv.Lid[prev].x == x) dv.next_a[prev] = cur; else dv.next_b[prev] = cur;
                prev = cur;
            }
        }
    }
    unsigned int m = __ballot_sync(0xffffffffu, has);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&dv.ctr->n_touched, __popc(m));
}

__global__ void __launch_bounds__(256) k_frontier0(BbxDev dv)
{
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    int n_round = (n + 31) & ~31;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_round; c += gridDim.x * blockDim.x) {
        bool solver = false, ready = false;
        if (c < n) {
            int2 l = dv.Lid[c];
            solver = (l.x >= 0 || l.y >= 0);
            ready = solver && dv.indeg[c] == 0;
        }
        unsigned int ms = __ballot_sync(0xffffffffu, solver);
        if ((threadIdx.x & 31) == 0 && ms) atomicAdd(&dv.ctr->n_solver, __popc(ms));
        int slot = warp_append(&dv.ctr->n_frontier0, ready);
        if (ready) dv.queue[slot] = c;
    }
}

// ---------------------------------------------------------------------------
// K11 velocity solve
// ---------------------------------------------------------------------------
struct SolveParams { float restitution, friction, vel_threshold; int iterations; };

struct BodyState {
    float3 v, w;
    float im;           // 1/mass
    float ii;           // spheres: 1/inertia
    float3 iI;          // boxes: 1/diag(I)
    float4 q;           // boxes: rotation
    bool box;
};

__device__ __forceinline__ void load_body(const BbxDev& dv, int g, BodyState& s)
{
    float4 l = __ldcg(&dv.lin[g]), a = __ldcg(&dv.ang[g]);
    s.v = xyz(l); s.w = xyz(a); s.im = l.w; s.ii = a.w;
    s.box = g >= dv.NS;
    if (s.box) { s.q = dv.rot[g - dv.NS]; s.iI = xyz(dv.invI[g - dv.NS]); }
}

__device__ __forceinline__ void store_body(const BbxDev& dv, int g, const BodyState& s)
{
    __stcg(&dv.lin[g], make_float4(s.v.x, s.v.y, s.v.z, s.im));
    __stcg(&dv.ang[g], make_float4(s.w.x, s.w.y, s.w.z, s.ii));
}

// ApplyConstraintImpulse :1359-1408
__device__ __forceinline__ void apply_impulse(BodyState& s, float3 J, float3 L)
{
    s.v = v3add(s.v, v3scale(J, s.im));
    if (!s.box) {
        s.w = v3add(s.w, v3scale(L, s.ii));
    } else {
        float3 loc = qrot(qconj(s.q), L);
        float3 dl = v3mul(loc, s.iI);
        s.w = v3add(s.w, qrot(s.q, dl));
    }
}

// one constraint, one Gauss-Seidel visit (:1414-1488)
__device__ __forceinline__ void solve_constraint(const BbxDev& dv, float4 c0, float4 c1, float4 c2, float4& c3, int2 id,
                                                 const SolveParams& sp, bool wake)
{
    float3 n = xyz(c0), rA = xyz(c1), rB = xyz(c2);
    float bias = c0.w, mN = c1.w, mT0 = c2.w, mT1 = c3.w;
    bool hasA = id.x >= 0, hasB = id.y >= 0;
    BodyState A, B;
    if (hasA) load_body(dv, id.x, A);
    if (hasB) load_body(dv, id.y, B);
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
    if (hasA) {
        store_body(dv, id.x, A);
        if (wake) { unsigned char f = dv.flags[id.x]; if (f & BF_SLEEP) { dv.flags[id.x] = f & (unsigned char)~BF_SLEEP; dv.timer[id.x] = 0.0f; } }   // :1373-1376
    }
    if (hasB) {
        store_body(dv, id.y, B);
        if (wake) { unsigned char f = dv.flags[id.y]; if (f & BF_SLEEP) { dv.flags[id.y] = f & (unsigned char)~BF_SLEEP; dv.timer[id.y] = 0.0f; } }
    }
}

__global__ void __launch_bounds__(256) k_solve_levels(BbxDev dv, SolveParams sp)
{
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    int* lc = dv.level_count;

    // ---- iteration 0: discover levels (Kahn) while solving ----
    int head = 0, cnt = __ldcg(&dv.ctr->n_frontier0), lvl = 0;
    if (tid == 0) { lc[0] = cnt; }
    while (cnt > 0) {
        if (tid == 0) lc[lvl + 2] = 0;
        int end = head + cnt;
        int next_base = end;
        int p_round = head + ((cnt + 31) & ~31);
        for (int p = head + tid; p < p_round; p += stride) {
            int sa = -1, sb = -1;
            if (p < end) {
                int c = __ldcg(&dv.queue[p]);
                float4 c0 = dv.L0[c], c1 = dv.L1[c], c2 = dv.L2[c], c3 = dv.L3[c];
                int2 id = dv.Lid[c];
                solve_constraint(dv, c0, c1, c2, c3, id, sp, true);
                dv.M0[p] = c0; dv.M1[p] = c1; dv.M2[p] = c2; dv.M3[p] = c3; dv.Mid[p] = id;
                dv.slot_of[c] = p;
                sa = dv.next_a[c]; sb = dv.next_b[c];
                if (sa >= 0 && atomicSub(&dv.indeg[sa], 1) != 1) sa = -1;
                if (sb >= 0 && atomicSub(&dv.indeg[sb], 1) != 1) sb = -1;
            }
            int s1 = warp_append(&lc[lvl + 1], sa >= 0);
            if (sa >= 0) dv.queue[next_base + s1] = sa;
            int s2 = warp_append(&lc[lvl + 1], sb >= 0);
            if (sb >= 0) dv.queue[next_base + s2] = sb;
        }
        grid.sync();
        head = end;
        cnt = __ldcg(&lc[lvl + 1]);
        lvl++;
    }
    const int n_levels = lvl;
    if (tid == 0) dv.ctr->n_levels = n_levels;

    // ---- iterations 1..I-1: stream the level-major records ----
    for (int it = 1; it < sp.iterations; it++) {
        int h = 0;
        for (int l = 0; l < n_levels; l++) {
            int c_l = __ldcg(&lc[l]);
            for (int p = h + tid; p < h + c_l; p += stride) {
                float4 c0 = dv.M0[p], c1 = dv.M1[p], c2 = dv.M2[p], c3 = __ldcg(&dv.M3[p]);
                int2 id = dv.Mid[p];
                solve_constraint(dv, c0, c1, c2, c3, id, sp, false);
                dv.M3[p] = c3;
            }
            h += c_l;
            grid.sync();
        }
    }
}

// touched-body statistics and sanity for the host
int bbx_stage_solve(BbxDev* d, const BbxStepParams* p)
{
    const int G = BBX_NUM_SMS * 8;
    BBX_CUDA(d, cudaMemsetAsync(d->deg, 0, sizeof(int) * ((size_t)d->N + 1), d->stream));
    BBX_LAUNCH(d, k_presolve, G, 256, 0, *d, p->settings.baumgarte_bias, p->settings.allowed_penetration, p->dt);
    if (p->settings.solver_iterations <= 0) { bbx_timing_mark(d, 4); BBX_CUDA(d, cudaGetLastError()); return 0; }    // :1223 loop does not run
    int rc = bbx_scan_exclusive(d, d->deg, d->adj_start, d->adj_cursor, d->N + 1, NULL, NULL);
    if (rc) return rc;
    BBX_LAUNCH(d, k_adj_fill, G, 256, 0, *d);
    BBX_LAUNCH(d, k_adj_link, bbx_blocks(d->N, 128), 128, 0, *d);
    BBX_CUDA(d, cudaMemsetAsync(d->level_count, 0, 4 * sizeof(int), d->stream));
    BBX_LAUNCH(d, k_frontier0, G, 256, 0, *d);

    if (d->coop_blocks == 0) {
        int per_sm = 0, sms = 0, coop = 0;
        BBX_CUDA(d, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d->device));
        if (!coop) return bbx_fail(d, BBX_ERR_NO_DEVICE, "device lacks cooperative launch", __FILE__, __LINE__);
        BBX_CUDA(d, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d->device));
        BBX_CUDA(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_levels, 256, 0));
        if (per_sm < 1) return bbx_fail(d, BBX_ERR_NO_DEVICE, "solver kernel does not fit on an SM", __FILE__, __LINE__);
        if (per_sm > 4) per_sm = 4;
        d->coop_blocks = per_sm * sms;
    }
    bbx_timing_mark(d, 4);
    SolveParams sp;
    sp.restitution = p->settings.restitution; sp.friction = p->settings.friction;
    sp.vel_threshold = p->settings.velocity_threshold; sp.iterations = p->settings.solver_iterations;
    BbxDev dv = *d;
    void* args[] = { (void*)&dv, (void*)&sp };
    BBX_CUDA(d, cudaLaunchCooperativeKernel((void*)k_solve_levels, dim3(d->coop_blocks), dim3(256), args, 0, d->stream));
    d->launches++;
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}
