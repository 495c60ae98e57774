// bbx_joints.cu -- joints and springs (EXTENSION: the reference lists "constraints, like hinges, springs" as planned,
// README.md:25, and has no code for them).  The sequential definition these kernels reproduce bit for bit is in
// oracle/bbx_oracle.c (stage_springs, the joint rows of stage_solve); the public calls are in ballbox_ext.h.
//
// SPRINGS (k_springs, before ApplyForces).  The definition adds every spring's force and torque to its two bodies in
// creation order.  A spring only READS positions and velocities, so the only order that matters is the order of the
// additions into one body's accumulators: one thread per body with spring ends walks that body's ends in creation
// order and recomputes each spring's force itself (same inputs, same operations -> same bits as the other end's
// thread).  No atomics, no levels.
//
// JOINTS.  Rows are presolved once per step (k_joint_presolve, one thread per joint: positions do not change during
// the solve) and swept after the contact constraints in every solver iteration, joints in creation order
// (k_joint_sweep).  That sweep is Gauss-Seidel: two joints that share a body must run in creation order.  The joint
// graph only changes when the user adds joints, so the schedule is built on the host when the list is set:
//   * connected components of the joint graph (union-find over body slots; the world is not a body) never interact:
//     each component belongs to ONE block, components are dealt to blocks largest first onto the least loaded block;
//   * inside a component, level(j) = 1 + max(level of the previous joint on body A, on body B): joints of one level
//     share no body.  A block sweeps its levels with __syncthreads() between them: no grid barrier, no atomics.
// Every body of a joint is treated as a dependency carrier, static or not (a body's static flag can change between
// steps; the schedule then stays valid).
#include <stdlib.h>
#include <algorithm>
#include <numeric>
#include <vector>
#include "bbx_solve.cuh"

#define JT_THREADS 128

// body frame -> world: boxes rotate (QuatRotateVec3 :630-639), spheres are attached at their centre, the world is fixed
__device__ __forceinline__ float3 anchor_world(const BbxDev& dv, int g, float3 l, float3& pos)
{
    if (g < 0) { pos = l; return l; }
    pos = xyz(dv.pos[g]);
    if (g < dv.NS) return v3add(pos, l);
    return v3add(pos, qrot(dv.boxq[2 * (size_t)(g - dv.NS)], l));
}

__global__ void __launch_bounds__(256) k_springs(BbxDev dv)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= dv.n_spring_bodies) return;
    const int g = dv.sp_body[t];
    const unsigned char fl = dv.flags[g];
    if (!(fl & BF_VALID) || (fl & BF_STATIC)) return;                       // only dynamic ends receive anything
    const bool is_box = g >= dv.NS;
    float* fp = is_box ? dv.r_b_force + 3 * (size_t)(g - dv.NS) : dv.r_s_force + 3 * (size_t)g;
    float* tp = is_box ? dv.r_b_torque + 3 * (size_t)(g - dv.NS) : dv.r_s_torque + 3 * (size_t)g;
    float3 F = f3(fp[0], fp[1], fp[2]), T = f3(tp[0], tp[1], tp[2]);
    bool any = false;
    for (int e = dv.sp_off[t]; e < dv.sp_off[t + 1]; e++) {
        const int code = dv.sp_list[e], s = code >> 1, end = code & 1;
        const int2 ab = dv.sp_ids[s];
        const float4 d0 = dv.sp_def[3 * (size_t)s], d1 = dv.sp_def[3 * (size_t)s + 1], d2 = dv.sp_def[3 * (size_t)s + 2];
        float3 posA, posB;
        const float3 pA = anchor_world(dv, ab.x, xyz(d0), posA), pB = anchor_world(dv, ab.y, xyz(d1), posB);
        const float3 zero = f3(0.0f, 0.0f, 0.0f);
        float3 vA = zero, oA = zero, vB = zero, oB = zero;
        if (!(dv.flags[ab.x] & BF_STATIC)) { vA = xyz(dv.vel[2 * (size_t)ab.x]); oA = xyz(dv.vel[2 * (size_t)ab.x + 1]); }
        if (ab.y >= 0 && !(dv.flags[ab.y] & BF_STATIC)) { vB = xyz(dv.vel[2 * (size_t)ab.y]); oB = xyz(dv.vel[2 * (size_t)ab.y + 1]); }
        const float3 d = v3sub(pB, pA);
        const float len = v3len(d);
        if (len < 1e-6f) continue;
        const float3 dir = v3scale(d, 1.0f / len);
        const float3 rA = v3sub(pA, posA), rB = v3sub(pB, posB);
        const float3 velA = v3add(vA, v3cross(oA, rA)), velB = v3add(vB, v3cross(oB, rB));
        const float f = d1.w * (len - d0.w) + d2.x * v3dot(v3sub(velB, velA), dir);
        const float3 Fs = v3scale(dir, f), Fm = v3scale(Fs, -1.0f);
        if (end == 0) { F = v3add(F, Fs); T = v3add(T, v3cross(rA, Fs)); }
        else          { F = v3add(F, Fm); T = v3add(T, v3cross(rB, Fm)); }
        any = true;
    }
    if (any) {
        fp[0] = F.x; fp[1] = F.y; fp[2] = F.z;
        tp[0] = T.x; tp[1] = T.y; tp[2] = T.z;
        if (fl & BF_SLEEP) { dv.flags[g] = fl & (unsigned char)~BF_SLEEP; dv.timer[g] = 0.0f; }
    }
}

// one thread per joint (schedule order): lever arms, effective masses, bias; accumulated impulses start at zero.
// wake: the sweeps will call ApplyConstraintImpulse on both bodies, which wakes them (:1373-1376), even with a zero impulse
__global__ void __launch_bounds__(256) k_joint_presolve(BbxDev dv, float beta_over_dt, int wake)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dv.n_joints) return;
    const int4 id = dv.j_ids[j];
    const float4 l0 = dv.j_local[2 * (size_t)j], l1 = dv.j_local[2 * (size_t)j + 1];
    float3 posA, posB;
    const float3 pA = anchor_world(dv, id.x, xyz(l0), posA), pB = anchor_world(dv, id.y, xyz(l1), posB);
    const float3 rA = v3sub(pA, posA), rB = v3sub(pB, posB);
    float imA = 0.0f, imB = 0.0f;
    float3 iA = f3(0.0f, 0.0f, 0.0f), iB = iA;
    {
        const F8 vw = ld256_cg(&dv.vel[2 * (size_t)id.x]);                   // 1/m and sphere 1/I are 0 for static bodies
        imA = vw.a.w;
        iA = id.x < dv.NS ? f3(vw.b.w, vw.b.w, vw.b.w) : xyz(dv.boxq[2 * (size_t)(id.x - dv.NS) + 1]);
        if (wake && !(dv.flags[id.x] & BF_STATIC)) wake_body(dv, id.x);
    }
    if (id.y >= 0) {
        const F8 vw = ld256_cg(&dv.vel[2 * (size_t)id.y]);
        imB = vw.a.w;
        iB = id.y < dv.NS ? f3(vw.b.w, vw.b.w, vw.b.w) : xyz(dv.boxq[2 * (size_t)(id.y - dv.NS) + 1]);
        if (wake && !(dv.flags[id.y] & BF_STATIC)) wake_body(dv, id.y);
    }
    const float3 C = v3sub(pB, pA);
    float4* R = dv.j_rec + 5 * (size_t)j;
    float m[3] = { 0.0f, 0.0f, 0.0f }, b[3] = { 0.0f, 0.0f, 0.0f };
    float3 dir = f3(0.0f, 0.0f, 0.0f);
    if (id.z == BBX_JOINT_ROD) {
        dir = v3norm(C);
        const float s = inv_mass_sum(imA, imB, rA, rB, iA, iB, dir);
        m[0] = (s > 0.0f) ? 1.0f / s : 0.0f;
        b[0] = -(beta_over_dt * (v3len(C) - l0.w));
    } else {
        const float c3[3] = { C.x, C.y, C.z };
        #pragma unroll
        for (int r = 0; r < 3; r++) {
            const float3 e = f3(r == 0 ? 1.0f : 0.0f, r == 1 ? 1.0f : 0.0f, r == 2 ? 1.0f : 0.0f);
            const float s = inv_mass_sum(imA, imB, rA, rB, iA, iB, e);
            m[r] = (s > 0.0f) ? 1.0f / s : 0.0f;
            b[r] = -(beta_over_dt * c3[r]);
        }
    }
    R[0] = make_float4(rA.x, rA.y, rA.z, m[0]);
    R[1] = make_float4(rB.x, rB.y, rB.z, m[1]);
    R[2] = make_float4(b[0], b[1], b[2], m[2]);
    R[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    R[4] = make_float4(dir.x, dir.y, dir.z, 0.0f);
}

// one bilateral row: no clamp, no restitution, no friction
__device__ __forceinline__ void joint_row(float3 e, float mass, float bias, float& imp, float3 rA, float3 rB, bool hasA, bool hasB,
                                          BodyState& A, BodyState& B)
{
    const float3 zero = f3(0.0f, 0.0f, 0.0f);
    const float3 velA = hasA ? v3add(A.v, v3cross(A.w, rA)) : zero;
    const float3 velB = hasB ? v3add(B.v, v3cross(B.w, rB)) : zero;
    const float vn = v3dot(v3sub(velB, velA), e);
    float jn = mass * (-(vn + 0.0f) + bias);
    const float old = imp;
    imp = imp + jn;
    jn = imp - old;
    const float3 P = v3scale(e, jn), Pm = v3scale(P, -1.0f);
    if (hasA) apply_impulse(A, Pm, v3cross(rA, Pm));
    if (hasB) apply_impulse(B, P, v3cross(rB, P));
}

// one Gauss-Seidel sweep over all joints: a block owns whole components and walks its levels
__global__ void __launch_bounds__(JT_THREADS) k_joint_sweep(BbxDev dv)
{
    const int k0 = dv.j_blk_off[blockIdx.x], k1 = dv.j_blk_off[blockIdx.x + 1];
    for (int k = k0; k < k1; k++) {
        const int j0 = dv.j_lev_off[k], j1 = dv.j_lev_off[k + 1];
        for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
            const int4 id = dv.j_ids[j];
            float4* R = dv.j_rec + 5 * (size_t)j;
            const float4 r0 = R[0], r1 = R[1], r2 = R[2];
            float4 r3 = __ldcg(&R[3]);
            const bool hasA = !(dv.flags[id.x] & BF_STATIC);
            const bool hasB = id.y >= 0 && !(dv.flags[id.y] & BF_STATIC);
            BodyState A, B;
            if (hasA) load_body(dv, id.x, A);
            if (hasB) load_body(dv, id.y, B);
            const float3 rA = xyz(r0), rB = xyz(r1);
            if (id.z == BBX_JOINT_ROD) {
                joint_row(xyz(R[4]), r0.w, r2.x, r3.x, rA, rB, hasA, hasB, A, B);
            } else {
                joint_row(f3(1.0f, 0.0f, 0.0f), r0.w, r2.x, r3.x, rA, rB, hasA, hasB, A, B);
                joint_row(f3(0.0f, 1.0f, 0.0f), r1.w, r2.y, r3.y, rA, rB, hasA, hasB, A, B);
                joint_row(f3(0.0f, 0.0f, 1.0f), r2.w, r2.z, r3.z, rA, rB, hasA, hasB, A, B);
            }
            __stcg(&R[3], r3);
            if (hasA) store_body(dv, id.x, A);
            if (hasB) store_body(dv, id.y, B);
        }
        __syncthreads();                                   // block-wide: orders the .cg (L2) stores and loads of consecutive levels
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static void free_joint_arrays(BbxDev* d)
{
    void* p[] = { d->j_blk_off, d->j_lev_off, d->j_ids, d->j_local, d->j_rec };
    for (size_t i = 0; i < sizeof(p) / sizeof(p[0]); i++) if (p[i]) cudaFree(p[i]);
    d->j_blk_off = d->j_lev_off = NULL; d->j_ids = NULL; d->j_local = d->j_rec = NULL;
    d->n_joints = d->n_jblocks = d->n_jlevels_max = 0;
}
static void free_spring_arrays(BbxDev* d)
{
    void* p[] = { d->sp_body, d->sp_off, d->sp_list, d->sp_ids, d->sp_def };
    for (size_t i = 0; i < sizeof(p) / sizeof(p[0]); i++) if (p[i]) cudaFree(p[i]);
    d->sp_body = d->sp_off = d->sp_list = NULL; d->sp_ids = NULL; d->sp_def = NULL;
    d->n_springs = d->n_spring_bodies = 0;
}
void bbx_joints_free(BbxDev* d) { free_joint_arrays(d); free_spring_arrays(d); }

static bool body_slot_ok(const BbxDev* d, int g, bool world_allowed) { return (g >= 0 && g < d->N) || (world_allowed && g == -1); }

template <typename T>
static int upload_vec(BbxDev* d, T** dst, const std::vector<T>& v)
{
    BBX_CUDA(d, cudaMalloc((void**)dst, sizeof(T) * (v.size() ? v.size() : 1)));
    if (!v.empty()) BBX_CUDA(d, cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int bbx_dev_set_joints(BbxDev* d, const BbxJointDef* joints, int n)
{
    if (!d || n < 0 || (n > 0 && !joints)) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));                           // no step may still be reading the old list
    free_joint_arrays(d);
    if (n == 0) return 0;
    for (int j = 0; j < n; j++)
        if (!body_slot_ok(d, joints[j].ga, false) || !body_slot_ok(d, joints[j].gb, true) || (joints[j].kind != BBX_JOINT_BALL && joints[j].kind != BBX_JOINT_ROD))
            return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "bbx_dev_set_joints: bad body slot or joint kind", __FILE__, __LINE__);
    // components (union-find over body slots) and levels (creation order)
    std::vector<int> parent((size_t)d->N);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (int j = 0; j < n; j++) if (joints[j].gb >= 0) { int a = find(joints[j].ga), b = find(joints[j].gb); if (a != b) parent[b] = a; }
    std::vector<int> last((size_t)d->N, 0), level((size_t)n), comp((size_t)n);
    std::vector<int> comp_of_root((size_t)d->N, -1), comp_size;
    for (int j = 0; j < n; j++) {
        int lv = last[joints[j].ga];
        if (joints[j].gb >= 0) lv = std::max(lv, last[joints[j].gb]);
        level[j] = lv + 1;
        last[joints[j].ga] = level[j];
        if (joints[j].gb >= 0) last[joints[j].gb] = level[j];
        const int r = find(joints[j].ga);
        if (comp_of_root[r] < 0) { comp_of_root[r] = (int)comp_size.size(); comp_size.push_back(0); }
        comp[j] = comp_of_root[r];
        comp_size[comp[j]]++;
    }
    // components -> blocks: largest first onto the least loaded block
    const int n_comp = (int)comp_size.size();
    const int n_blocks = std::min(n_comp, 4 * d->num_sms);
    std::vector<int> order((size_t)n_comp), blk_of_comp((size_t)n_comp), load((size_t)n_blocks, 0);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return comp_size[a] > comp_size[b]; });
    for (int c : order) {
        const int b = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        blk_of_comp[c] = b; load[b] += comp_size[c];
    }
    // schedule order: (block, level, creation order)
    std::vector<int> sched((size_t)n);
    std::iota(sched.begin(), sched.end(), 0);
    std::stable_sort(sched.begin(), sched.end(), [&](int a, int b) {
        const int ba = blk_of_comp[comp[a]], bb = blk_of_comp[comp[b]];
        if (ba != bb) return ba < bb;
        return level[a] < level[b];
    });
    std::vector<int> blk_off((size_t)n_blocks + 1, 0), lev_off;
    std::vector<int4> ids((size_t)n);
    std::vector<float4> local(2 * (size_t)n);
    int max_levels = 0, cur_blk = -1, cur_lev = -1, blk_levels = 0;
    for (int s = 0; s < n; s++) {
        const int j = sched[s], b = blk_of_comp[comp[j]];
        if (b != cur_blk) {
            for (int x = cur_blk + 1; x <= b; x++) blk_off[x] = (int)lev_off.size();
            max_levels = std::max(max_levels, blk_levels);
            cur_blk = b; cur_lev = -1; blk_levels = 0;
        }
        if (level[j] != cur_lev) { lev_off.push_back(s); cur_lev = level[j]; blk_levels++; }
        ids[s] = make_int4(joints[j].ga, joints[j].gb, joints[j].kind, j);
        local[2 * (size_t)s] = make_float4(joints[j].la[0], joints[j].la[1], joints[j].la[2], joints[j].rest);
        local[2 * (size_t)s + 1] = make_float4(joints[j].lb[0], joints[j].lb[1], joints[j].lb[2], 0.0f);
    }
    max_levels = std::max(max_levels, blk_levels);
    for (int x = cur_blk + 1; x <= n_blocks; x++) blk_off[x] = (int)lev_off.size();
    lev_off.push_back(n);
    int rc;
    if ((rc = upload_vec(d, &d->j_blk_off, blk_off)) || (rc = upload_vec(d, &d->j_lev_off, lev_off)) ||
        (rc = upload_vec(d, &d->j_ids, ids)) || (rc = upload_vec(d, &d->j_local, local))) return rc;
    BBX_CUDA(d, cudaMalloc((void**)&d->j_rec, sizeof(float4) * 5 * (size_t)n));
    BBX_CUDA(d, cudaMemset(d->j_rec, 0, sizeof(float4) * 5 * (size_t)n));
    d->n_joints = n; d->n_jblocks = n_blocks; d->n_jlevels_max = max_levels;
    return 0;
}

extern "C" int bbx_dev_set_springs(BbxDev* d, const BbxSpringDef* springs, int n)
{
    if (!d || n < 0 || (n > 0 && !springs)) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    free_spring_arrays(d);
    if (n == 0) return 0;
    for (int s = 0; s < n; s++)
        if (!body_slot_ok(d, springs[s].ga, false) || !body_slot_ok(d, springs[s].gb, true))
            return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "bbx_dev_set_springs: bad body slot", __FILE__, __LINE__);
    // per-body lists of spring ends in creation order
    std::vector<std::pair<int, int>> ends;                                   // (body, spring << 1 | end)
    ends.reserve(2 * (size_t)n);
    for (int s = 0; s < n; s++) {
        ends.push_back({ springs[s].ga, s << 1 });
        if (springs[s].gb >= 0) ends.push_back({ springs[s].gb, (s << 1) | 1 });
    }
    std::stable_sort(ends.begin(), ends.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
    std::vector<int> body, off, list;
    for (size_t e = 0; e < ends.size(); e++) {
        if (e == 0 || ends[e].first != ends[e - 1].first) { body.push_back(ends[e].first); off.push_back((int)e); }
        list.push_back(ends[e].second);
    }
    off.push_back((int)ends.size());
    std::vector<int2> ids((size_t)n);
    std::vector<float4> def(3 * (size_t)n);
    for (int s = 0; s < n; s++) {
        ids[s] = make_int2(springs[s].ga, springs[s].gb);
        def[3 * (size_t)s] = make_float4(springs[s].la[0], springs[s].la[1], springs[s].la[2], springs[s].rest);
        def[3 * (size_t)s + 1] = make_float4(springs[s].lb[0], springs[s].lb[1], springs[s].lb[2], springs[s].k);
        def[3 * (size_t)s + 2] = make_float4(springs[s].c, 0.0f, 0.0f, 0.0f);
    }
    int rc;
    if ((rc = upload_vec(d, &d->sp_body, body)) || (rc = upload_vec(d, &d->sp_off, off)) || (rc = upload_vec(d, &d->sp_list, list)) ||
        (rc = upload_vec(d, &d->sp_ids, ids)) || (rc = upload_vec(d, &d->sp_def, def))) return rc;
    d->n_springs = n; d->n_spring_bodies = (int)body.size();
    return 0;
}

int bbx_springs_apply(BbxDev* d)
{
    if (d->n_springs <= 0) return 0;
    BBX_LAUNCH(d, k_springs, bbx_blocks(d->n_spring_bodies, 256), 256, 0, *d);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

int bbx_joints_presolve(BbxDev* d, const BbxStepParams* p, int wake)
{
    if (d->n_joints <= 0) return 0;
    BBX_LAUNCH(d, k_joint_presolve, bbx_blocks(d->n_joints, 256), 256, 0, *d, p->settings.baumgarte_bias / p->dt, wake);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

int bbx_joints_sweep(BbxDev* d)
{
    if (d->n_joints <= 0) return 0;
    BBX_LAUNCH(d, k_joint_sweep, d->n_jblocks, JT_THREADS, 0, *d);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}
