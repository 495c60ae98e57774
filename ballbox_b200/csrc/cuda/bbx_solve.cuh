// bbx_solve.cuh -- device code shared by the solver translation units (bbx_solver.cu: presolve, dependency
// graph, level discovery and the batched-world solver; bbx_replay.cu: the staged level replay; bbx_solver_flow.cu:
// the alternative barrier-free exact schedule).  Reference: SolveVelocityConstraints ballbox.h:1410-1489,
// ApplyConstraintImpulse :1359-1408, GetConstraintBodyVelocityAtPoint :1233-1261.
#pragma once
#include "bbx_dev.cuh"

__device__ __forceinline__ void tangents_of(float3 n, float3& t1, float3& t2)          // :1329-1336
{
    if (fabsf(n.x) >= 0.57735f) t1 = f3(n.y, -n.x, 0.0f);
    else                        t1 = f3(0.0f, n.z, -n.y);
    t1 = v3norm(t1);
    t2 = v3cross(n, t1);
}

// sum of the inverse masses seen along `dir` at the two lever arms (:1321-1326); the effective mass is its inverse
__device__ __forceinline__ float inv_mass_sum(float imA, float imB, float3 rA, float3 rB, float3 iA, float3 iB, float3 dir)
{
    float3 ra = v3cross(rA, dir), rb = v3cross(rB, dir);
    return imA + imB + v3dot(ra, v3mul(ra, iA)) + v3dot(rb, v3mul(rb, iB));             // :1323-1325
}

struct SolveParams { float restitution, friction, vel_threshold; int iterations; int gather; int first_only; };

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
// Both bodies of a node at once: ALL loads are issued before the first value is used.  With two calls of load_body_keep
// the box scalars of body A (k1, k2) are computed inside A's branch, i.e. the warp waits for A's rotation before it even
// requests body B: two (for a box-box node: two DRAM / L2 round trips) instead of one.
__device__ __forceinline__ void load_bodies_keep(const BbxDev& dv, int gA, int gB, bool hasA, bool hasB, BodyState& A, BodyState& B,
                                                 unsigned long long pol)
{
    const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    F8 va, vb, qa, qb;
    va.a = va.b = vb.a = vb.b = qa.a = qa.b = qb.a = qb.b = z4;
    const bool boxA = hasA && gA >= dv.NS, boxB = hasB && gB >= dv.NS;
    if (hasA) va = ld256_cg_pol(&dv.vel[2 * (size_t)gA], pol);
    if (boxA) qa = ld256_nc_pol(&dv.boxq[2 * (size_t)(gA - dv.NS)], pol);
    if (hasB) vb = ld256_cg_pol(&dv.vel[2 * (size_t)gB], pol);
    if (boxB) qb = ld256_nc_pol(&dv.boxq[2 * (size_t)(gB - dv.NS)], pol);
    A.v = xyz(va.a); A.w = xyz(va.b); A.im = va.a.w; A.ii = va.b.w; A.box = boxA; A.q = qa.a; A.iI = xyz(qa.b);
    B.v = xyz(vb.a); B.w = xyz(vb.b); B.im = vb.a.w; B.ii = vb.b.w; B.box = boxB; B.q = qb.a; B.iI = xyz(qb.b);
    { const float3 u = f3(A.q.x, A.q.y, A.q.z); A.k1 = A.q.w * A.q.w - v3dot(u, u); A.k2 = 2.0f * A.q.w; }
    { const float3 u = f3(B.q.x, B.q.y, B.q.z); B.k1 = B.q.w * B.q.w - v3dot(u, u); B.k2 = 2.0f * B.q.w; }
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

// The same operations without control flow: both angular updates are computed and one is selected.  A body that is absent
// (static) must hold defined values (zeros); its state is updated with garbage that nobody reads (velocities of an absent
// body enter the sweep as zero, and it is never stored).  With the `if (hasA) ... if (hasB) ...` of the plain form the two
// applications of an impulse sit in different basic blocks and a lone warp executes their dependent chains one after the
// other; here ptxas interleaves them.  Used where a sweep is latency bound (k_solve_groups: one warp per level of a world).
__device__ __forceinline__ void apply_impulse_flat(BodyState& s, float3 J, float3 L)
{
    s.v = v3add(s.v, v3scale(J, s.im));
    const float3 ws = v3add(s.w, v3scale(L, s.ii));
    const float3 u = f3(s.q.x, s.q.y, s.q.z), um = f3(-s.q.x, -s.q.y, -s.q.z);
    const float3 loc = qrot_pre(um, s.k1, s.k2, L);
    const float3 dl = v3mul(loc, s.iI);
    const float3 wb = v3add(s.w, qrot_pre(u, s.k1, s.k2, dl));
    s.w = s.box ? wb : ws;
}
__device__ __forceinline__ BodyState body_state_zero()
{
    BodyState s;
    s.v = s.w = s.iI = f3(0.0f, 0.0f, 0.0f); s.im = s.ii = s.k1 = s.k2 = 0.0f; s.q = make_float4(0.0f, 0.0f, 0.0f, 0.0f); s.box = false;
    return s;
}
// velocity of a body at the contact point (:1233-1261), zero for an absent body.  FLAT: computed unconditionally and
// masked (bit-and with all ones or zero: exactly v or +0), so that the A and B sides stay in one basic block
template <bool FLAT>
__device__ __forceinline__ float3 point_velocity(const BodyState& s, bool has, float3 r)
{
    if (!FLAT) return has ? v3add(s.v, v3cross(s.w, r)) : f3(0.0f, 0.0f, 0.0f);
    const float3 t = v3add(s.v, v3cross(s.w, r));
    const int m = has ? -1 : 0;
    return f3(__int_as_float(__float_as_int(t.x) & m), __int_as_float(__float_as_int(t.y) & m), __int_as_float(__float_as_int(t.z) & m));
}
template <bool FLAT>
__device__ __forceinline__ void apply_side(BodyState& s, bool has, float3 J, float3 L)
{
    if (FLAT) apply_impulse_flat(s, J, L);
    else if (has) apply_impulse(s, J, L);
}

// one constraint, one Gauss-Seidel visit (:1414-1488) on bodies held in registers
template <bool FLAT = false>
__device__ __forceinline__ void solve_constraint(float4 c0, float4 c1, float4 c2, float4& c3, bool hasA, bool hasB,
                                                 BodyState& A, BodyState& B, const SolveParams& sp)
{
    float3 n = xyz(c0), rA = xyz(c1), rB = xyz(c2);
    float bias = c0.w, mN = c1.w, mT0 = c2.w, mT1 = c3.w;
    float3 vA = point_velocity<FLAT>(A, hasA, rA);                           // :1233-1261
    float3 vB = point_velocity<FLAT>(B, hasB, rB);
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
    apply_side<FLAT>(A, hasA, Pm, v3cross(rA, Pm));                         // :1456-1459
    apply_side<FLAT>(B, hasB, P, v3cross(rB, P));
    if (sp.friction > 0.0f) {                                                // :1462-1487
        vA = point_velocity<FLAT>(A, hasA, rA);
        vB = point_velocity<FLAT>(B, hasB, rB);
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
            apply_side<FLAT>(A, hasA, Tm, v3cross(rA, Tm));
            apply_side<FLAT>(B, hasB, T, v3cross(rB, T));
        }
        {
            float vt = v3dot(rel, t2);                                       // same pre-loop rel (:1466 is outside the j loop)
            float jt = mT1 * (-vt);
            float o = c3.z;
            c3.z = bb_fmaxf(-maxF, bb_fminf(maxF, c3.z + jt));
            jt = c3.z - o;
            float3 T = v3scale(t2, jt), Tm = v3scale(T, -1.0f);
            apply_side<FLAT>(A, hasA, Tm, v3cross(rA, Tm));
            apply_side<FLAT>(B, hasB, T, v3cross(rB, T));
        }
    }
}

// Per level, the nodes of class k sit in queue_k[k][head_k .. head_k + cnt_k).  The flattened
// index space pads every class to a multiple of 32 so that a warp never mixes classes.
struct LevelMap { int head[BBX_NCLS]; int cnt[BBX_NCLS]; int pref[BBX_NCLS + 1]; };

__device__ __forceinline__ void level_map_load(LevelMap& lm, const int* lc_level, bool first_level)
{
    // called by one thread; heads advance by the previous level's counts.  The flattened index space starts with the
    // HEAVIEST class (pref[7] = 0 <= pref[6] <= ... <= pref[0]; pref[8] = total): the partial last round of a level then
    // consists of the lightest chunks
    int run = 0;
    for (int k = BBX_NCLS - 1; k >= 0; k--) {
        if (first_level) lm.head[k] = 0; else lm.head[k] += lm.cnt[k];
        lm.cnt[k] = __ldcg(&lc_level[k]);
        lm.pref[k] = run;
        run += (lm.cnt[k] + 31) & ~31;
    }
    lm.pref[BBX_NCLS] = run;
}

// adjacency entries (k_adj_fill / k_adj_link): (order key << 32) | side << 31 | node
#define ADJ_NODE(e) ((int)((unsigned int)(e) & 0x7fffffffu))
#define ADJ_SIDE(e) ((int)(((unsigned int)(e)) >> 31))

#define RS_MAXLEV 256                  // levels whose tables the staged replay (bbx_replay.cu) keeps in shared memory
#define RS_MIN_NODES_PER_LEVEL 8192     // thinner schedules replay inside k_solve_levels
#define FLOW_DBG 4088                  // level_ns[FLOW_DBG..+8): debug counters of the flow solver / level trace

// host entry points of the other solver translation units
int bbx_flow_solve(BbxDev* d, const BbxStepParams* p, int iterations);
int bbx_ensure_slot_of(BbxDev* d);
int bbx_replay_staged(BbxDev* d, const SolveParams& sp);
// coloured schedule (bbx_colour.cu)
#define BBX_MAX_COLOURS 64
#define COL_TOT   0                                   // [BBX_NCLS] nodes per class (all levels)
#define COL_CNT   8                                   // [64 * BBX_NCLS] nodes per (colour, class)
#define COL_FILL  (COL_CNT + BBX_MAX_COLOURS * BBX_NCLS)
#define COL_HEAD  (COL_FILL + BBX_MAX_COLOURS * BBX_NCLS)
#define COL_MBASE (COL_HEAD + BBX_MAX_COLOURS * BBX_NCLS)
#define COL_INTS  (COL_MBASE + BBX_MAX_COLOURS + 8)
int bbx_colour_ensure(BbxDev* d);
int bbx_colour_rebucket(BbxDev* d);
