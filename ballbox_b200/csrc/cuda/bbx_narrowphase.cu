// bbx_narrowphase.cu -- K6/K7: contact generation on the compacted candidate
// lists and SDF probing.  Every function restates the reference's arithmetic in
// the same operation order (see the ARITHMETIC CONTRACT in bbx_dev.cuh); the
// quirks listed in SURVEY.md section 8a rows 5-9 are reproduced, not fixed.
//
//   sphere-box            CheckSphereBoxCollisionWithData   ballbox.h:1565-1626
//   SAT + contact point   SATCollision                      ballbox.h:1040-1111
//                         CheckBoxCollisionWithData         ballbox.h:1117-1132
//   manifold              GenerateBoxBoxManifold            ballbox.h:2550-2724
//                         SelectOptimalContactPoints        ballbox.h:2726-2843
//   SDF                   sphere block :2074-2106, SDF_EstimateNormal :2318-2331,
//                         GenerateBoxSDFContacts :2413-2477
#include "bbx_dev.cuh"
#include "bbx_geom.cuh"

__device__ __forceinline__ void store_contact(const BbxDev& dv, int slot, int a, int b, int k, int phase,
                                              float3 pt, float pen, float3 nrm)
{
    dv.c_ids[slot] = make_int4(a, b, k, phase);
    dv.c_pt[slot] = make_float4(pt.x, pt.y, pt.z, pen);
    dv.c_nrm[slot] = make_float4(nrm.x, nrm.y, nrm.z, 0.0f);
}

__global__ void __launch_bounds__(256) k_narrow_sb(BbxDev dv)
{
    __shared__ int s_warp[32];
    int n = min(dv.ctr->n_cand_sb, dv.cap_cand);
    int n_round = (n + blockDim.x - 1) / blockDim.x * blockDim.x;          // block-uniform trip count (block_reserve)
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += gridDim.x * blockDim.x) {
        bool hit = false;
        float3 pt, nrm; float pen = 0.0f; int2 pr = make_int2(0, 0);
        if (t < n) {
            pr = dv.cand_sb[t];
            float4 sp = dv.pos[pr.x], bp = dv.pos[pr.y];
            int b = pr.y - dv.NS;
            hit = sphere_box_contact(xyz(sp), sp.w, xyz(bp), xyz(dv.half[b]), dv.boxq[2 * b], pt, nrm, pen);
        }
        int slot = block_reserve(&dv.ctr->n_contacts, hit ? 1 : 0, s_warp);
        if (hit) {
            if (slot < dv.cap_contacts) store_contact(dv, slot, pr.x, pr.y, RUN_ENC(0, 1), PH_SB, pt, pen, nrm);
            else atomicOr(&dv.ctr->overflow, 2);
        }
    }
}

// Pass 0: the six FACE axes of the SAT on every candidate.  Two thirds of the candidates (bounding spheres touch, boxes
// do not) are separated by a face axis; with the full 15-axis test in one kernel the lanes of those pairs sat idle while
// their neighbours ran the nine edge axes (16.6 of 32 lanes active, ncu).  Survivors are compacted into `cand_sb` (free by
// now: k_narrow_sb runs first), and pass 1 runs the complete SAT on them with dense warps; the face axes are tested
// again there (same function, same outcome), which costs less than carrying the partial minimum.
__global__ void __launch_bounds__(256) k_bb_face(BbxDev dv)
{
    __shared__ int s_warp[32];
    int n = min(dv.ctr->n_cand_bb, dv.cap_cand);
    int n_round = (n + blockDim.x - 1) / blockDim.x * blockDim.x;          // block-uniform trip count (block_reserve)
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += gridDim.x * blockDim.x) {
        bool keep = false;
        int2 pr = make_int2(0, 0);
        if (t < n) {
            pr = dv.cand_bb[t];
            int a = pr.x - dv.NS, b = pr.y - dv.NS;
            Obb A = make_obb(xyz(dv.pos[pr.x]), xyz(dv.half[a]), dv.boxq[2 * a]);
            Obb B = make_obb(xyz(dv.pos[pr.y]), xyz(dv.half[b]), dv.boxq[2 * b]);
            keep = sat_face_axes_overlap(A, B);
        }
        int slot = block_reserve(&dv.ctr->n_surv_bb, keep ? 1 : 0, s_warp);
        if (keep) dv.cand_sb[slot] = pr;                     // survivors <= candidates <= cap_cand
    }
}

// Pass 1: SAT on the survivors of pass 0; hits are compacted so that pass 2 runs with full warps.
// hit_ids = (box A gid, box B gid, axis type, -), hit_ax = (SAT axis, |min overlap|).
__global__ void __launch_bounds__(256) k_bb_sat(BbxDev dv)
{
    __shared__ int s_warp[32];
    int n = min(dv.ctr->n_surv_bb, dv.cap_cand);
    int n_round = (n + blockDim.x - 1) / blockDim.x * blockDim.x;          // block-uniform trip count (block_reserve)
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += gridDim.x * blockDim.x) {
        bool hit = false;
        int2 pr = make_int2(0, 0); float minPd = 0.0f; float3 ax = f3(0, 0, 0); int axisType = -1;
        if (t < n) {
            pr = dv.cand_sb[t];                              // survivor list of pass 0
            int a = pr.x - dv.NS, b = pr.y - dv.NS;
            Obb A = make_obb(xyz(dv.pos[pr.x]), xyz(dv.half[a]), dv.boxq[2 * a]);
            Obb B = make_obb(xyz(dv.pos[pr.y]), xyz(dv.half[b]), dv.boxq[2 * b]);
            hit = sat_axes(A, B, minPd, ax, axisType);
        }
        // face-axis hits (contact point from 96 vertex/face tests) and edge-axis hits (contact point = midpoint of
        // the centres) cost 5:1 in pass 2: they are kept apart, from the front and from the back of the hit list,
        // so that the warps of pass 2 are homogeneous
        const bool face = hit && axisType <= 5, edge = hit && axisType > 5;
        int slot_f = block_reserve(&dv.ctr->n_hit_bb, face ? 1 : 0, s_warp);
        int slot_e = block_reserve(&dv.ctr->n_hit_bb_edge, edge ? 1 : 0, s_warp);
        if (hit) {                                             // capacity: hits <= candidates <= cap_cand
            int slot = face ? slot_f : dv.cap_cand - 1 - slot_e;
            dv.hit_ids[slot] = make_int4(pr.x, pr.y, axisType, 0);
            dv.hit_ax[slot] = make_float4(ax.x, ax.y, ax.z, fabsf(minPd));
        }
    }
}

// Pass 2: contact point, manifold candidates and point selection for colliding pairs only.
__global__ void __launch_bounds__(128) k_bb_manifold(BbxDev dv, float tol, float pen_tol, int max_contacts)
{
    const int nf = dv.ctr->n_hit_bb, ne = dv.ctr->n_hit_bb_edge;          // nf + ne <= cap_cand by construction
    __shared__ int s_warp[32];
    const int nf_round = (nf + 31) & ~31;
    const int n_round = (nf_round + ((ne + 31) & ~31) + blockDim.x - 1) / blockDim.x * blockDim.x;      // block-uniform trip count
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += gridDim.x * blockDim.x) {
        int np = 0;
        float3 pts[8]; float pens[8]; float3 normal = f3(0, 0, 0);
        int4 h = make_int4(0, 0, 0, 0);
        const int src = t < nf_round ? (t < nf ? t : -1) : (t - nf_round < ne ? dv.cap_cand - 1 - (t - nf_round) : -1);
        if (src >= 0) {
            h = dv.hit_ids[src];
            float4 ax = dv.hit_ax[src];
            normal = xyz(ax);
            int a = h.x - dv.NS, b = h.y - dv.NS;
            np = box_box_manifold(xyz(dv.pos[h.x]), xyz(dv.half[a]), dv.boxq[2 * a], xyz(dv.pos[h.y]), xyz(dv.half[b]),
                                  dv.boxq[2 * b], tol, pen_tol, max_contacts, ax.w, h.z, pts, pens, normal);
        }
        int slot = block_reserve(&dv.ctr->n_contacts, np, s_warp);
        if (np > 0) {
            if (slot + np <= dv.cap_contacts) {
                for (int k = 0; k < np; k++) store_contact(dv, slot + k, h.x, h.y, RUN_ENC(k, np), PH_BB, pts[k], pens[k], normal);
            } else atomicOr(&dv.ctr->overflow, 2);
        }
    }
}

// ---------------------------------------------------------------------------
// SDF probing
// ---------------------------------------------------------------------------
struct SdfDesc { int id; float prm[8]; float eps; const float* grid; int ch; int nx, ny, nz; float ox, oy, oz, inv_cell; };

// baked host SDF: trilinear interpolation of the grid samples, coordinates clamped to the baked region.
// Plain f32 operations in a fixed order (compiled with -fmad=false like everything else): deterministic.
struct BakedCell { size_t base, sy, sz; float fx, fy, fz; };
__device__ __forceinline__ BakedCell baked_cell(const SdfDesc& s, float3 p)
{
    float ux = (p.x - s.ox) * s.inv_cell, uy = (p.y - s.oy) * s.inv_cell, uz = (p.z - s.oz) * s.inv_cell;
    ux = fminf(fmaxf(ux, 0.0f), (float)(s.nx - 1)); uy = fminf(fmaxf(uy, 0.0f), (float)(s.ny - 1)); uz = fminf(fmaxf(uz, 0.0f), (float)(s.nz - 1));
    int ix = min((int)ux, s.nx - 2), iy = min((int)uy, s.ny - 2), iz = min((int)uz, s.nz - 2);
    BakedCell c;
    c.fx = ux - (float)ix; c.fy = uy - (float)iy; c.fz = uz - (float)iz;
    c.base = ((size_t)iz * s.ny + iy) * s.nx + ix; c.sy = (size_t)s.nx; c.sz = (size_t)s.nx * s.ny;
    return c;
}
__device__ __forceinline__ float trilerp(const BakedCell& c, float v000, float v100, float v010, float v110,
                                         float v001, float v101, float v011, float v111)
{
    float a00 = v000 + (v100 - v000) * c.fx, a10 = v010 + (v110 - v010) * c.fx;
    float a01 = v001 + (v101 - v001) * c.fx, a11 = v011 + (v111 - v011) * c.fx;
    float b0 = a00 + (a10 - a00) * c.fy, b1 = a01 + (a11 - a01) * c.fy;
    return b0 + (b1 - b0) * c.fz;
}
__device__ __forceinline__ float sdf_baked(const SdfDesc& s, float3 p)
{
    const BakedCell c = baked_cell(s, p);
    const size_t k = (size_t)s.ch;                      // 1: distance only; 4: distance + gradient, distance first
    const float* g = s.grid + c.base * k;
    return trilerp(c, __ldg(g), __ldg(g + k), __ldg(g + c.sy * k), __ldg(g + (c.sy + 1) * k),
                   __ldg(g + c.sz * k), __ldg(g + (c.sz + 1) * k), __ldg(g + (c.sz + c.sy) * k), __ldg(g + (c.sz + c.sy + 1) * k));
}
// gradient channels (the reference's central differences of the host function, sampled at the grid nodes)
__device__ __forceinline__ float3 sdf_baked_gradient(const SdfDesc& s, float3 p)
{
    const BakedCell c = baked_cell(s, p);
    const float4* g = reinterpret_cast<const float4*>(s.grid) + c.base;
    float4 v000 = __ldg(g), v100 = __ldg(g + 1), v010 = __ldg(g + c.sy), v110 = __ldg(g + c.sy + 1);
    float4 v001 = __ldg(g + c.sz), v101 = __ldg(g + c.sz + 1), v011 = __ldg(g + c.sz + c.sy), v111 = __ldg(g + c.sz + c.sy + 1);
    return f3(trilerp(c, v000.y, v100.y, v010.y, v110.y, v001.y, v101.y, v011.y, v111.y),
              trilerp(c, v000.z, v100.z, v010.z, v110.z, v001.z, v101.z, v011.z, v111.z),
              trilerp(c, v000.w, v100.w, v010.w, v110.w, v001.w, v101.w, v011.w, v111.w));
}

__device__ __forceinline__ float sdf_at(const SdfDesc& s, float3 p)
{
    if (s.id == BBX_SDF_ID_BAKED) return sdf_baked(s, p);
    return bbx_sdf_eval(s.id, s.prm, p.x, p.y, p.z);
}

__device__ __forceinline__ float3 sdf_normal(const SdfDesc& s, float3 p)               // SDF_EstimateNormal :2318-2331
{
    if (s.id == BBX_SDF_ID_BAKED && s.ch == 4) return v3norm(sdf_baked_gradient(s, p));
    float e = s.eps;
    float3 ex = f3(e, 0.0f, 0.0f), ey = f3(0.0f, e, 0.0f), ez = f3(0.0f, 0.0f, e);
    float nx = (sdf_at(s, v3add(p, ex)) - sdf_at(s, v3sub(p, ex))) / (2.0f * e);
    float ny = (sdf_at(s, v3add(p, ey)) - sdf_at(s, v3sub(p, ey))) / (2.0f * e);
    float nz = (sdf_at(s, v3add(p, ez)) - sdf_at(s, v3sub(p, ez))) / (2.0f * e);
    return v3norm(f3(nx, ny, nz));
}

__global__ void __launch_bounds__(256) k_sdf_spheres(BbxDev dv, SdfDesc sdf)
{
    int n_round = (dv.NS + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool hit = false; float3 pt, nrm; float pen = 0.0f;
        if (i < dv.NS && (dv.flags[i] & BF_VALID)) {
            float4 p4 = dv.pos[i];
            float3 c = xyz(p4); float r = p4.w;
            float dd = sdf_at(sdf, c);
            if (dd < r) {                                           // :2086
                hit = true;
                pen = r - dd;
                nrm = v3scale(sdf_normal(sdf, c), -1.0f);
                pt = v3sub(c, v3scale(nrm, r));                     // :2094
            }
        }
        int slot = warp_reserve(&dv.ctr->n_contacts, hit ? 1 : 0);
        if (hit) {
            if (slot < dv.cap_contacts) store_contact(dv, slot, i, -2, RUN_ENC(0, 1), PH_SSDF, pt, pen, nrm);
            else atomicOr(&dv.ctr->overflow, 2);
        }
    }
}

__global__ void __launch_bounds__(128) k_sdf_boxes(BbxDev dv, SdfDesc sdf)
{
    const int edge_pairs[12][2] = { {0,1},{1,2},{2,3},{3,0}, {4,5},{5,6},{6,7},{7,4}, {0,4},{1,5},{2,6},{3,7} };   // :2454-2458
    int n_round = (dv.NB + 31) & ~31;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n_round; b += gridDim.x * blockDim.x) {
        int g = dv.NS + b;
        bool valid = b < dv.NB && (dv.flags[g] & BF_VALID);
        float3 vtx[8];
        unsigned int mask = 0;            // which of the 20 probes penetrate
        if (valid) {
            Obb o = make_obb(xyz(dv.pos[g]), xyz(dv.half[b]), dv.boxq[2 * b]);
            obb_vertices(o, vtx);
            for (int k = 0; k < 8; k++) if (sdf_at(sdf, vtx[k]) < -0.001f) mask |= 1u << k;
            for (int e = 0; e < 12; e++) {
                float3 mid = v3scale(v3add(vtx[edge_pairs[e][0]], vtx[edge_pairs[e][1]]), 0.5f);
                if (sdf_at(sdf, mid) < -0.001f) mask |= 1u << (8 + e);
            }
        }
        int np = __popc(mask);
        int slot = warp_reserve(&dv.ctr->n_contacts, np);
        if (np > 0) {
            if (slot + np > dv.cap_contacts) { atomicOr(&dv.ctr->overflow, 2); continue; }
            int k = 0;
            for (int pidx = 0; pidx < 20; pidx++) {
                if (!(mask & (1u << pidx))) continue;
                float3 p = (pidx < 8) ? vtx[pidx]
                                      : v3scale(v3add(vtx[edge_pairs[pidx - 8][0]], vtx[edge_pairs[pidx - 8][1]]), 0.5f);
                float dd = sdf_at(sdf, p);
                float3 nrm = v3scale(sdf_normal(sdf, p), -1.0f);
                store_contact(dv, slot + k, g, -2, RUN_ENC(k, np), PH_BSDF, p, -dd, nrm);
                k++;
            }
        }
    }
}

int bbx_stage_narrowphase(BbxDev* d, const BbxStepParams* p)
{
    int mc = p->settings.manifold_max_contacts;
    if (mc < 1) mc = 1;
    if (mc > MAX_MANIFOLD_CONTACTS) mc = MAX_MANIFOLD_CONTACTS;
    BBX_LAUNCH(d, k_narrow_sb, BBX_NUM_SMS * 8, 256, 0, *d);          // first: its candidate list is the scratch of the box-box passes
    BBX_LAUNCH(d, k_bb_face, BBX_NUM_SMS * 8, 256, 0, *d);
    BBX_LAUNCH(d, k_bb_sat, BBX_NUM_SMS * 8, 256, 0, *d);
    BBX_LAUNCH(d, k_bb_manifold, BBX_NUM_SMS * 8, 128, 0, *d, p->settings.manifold_contact_tolerance,
               p->settings.manifold_penetration_tolerance, mc);
    if (p->sdf_id != BBX_SDF_NONE) {
        SdfDesc s; s.id = p->sdf_id; s.eps = p->settings.sdf_normal_epsilon;
        for (int k = 0; k < 8; k++) s.prm[k] = p->sdf_params[k];
        s.grid = d->sdf_grid; s.ch = d->sdf_ch; s.nx = d->sdf_nx; s.ny = d->sdf_ny; s.nz = d->sdf_nz;
        s.ox = d->sdf_ox; s.oy = d->sdf_oy; s.oz = d->sdf_oz; s.inv_cell = d->sdf_inv_cell;
        if (s.id == BBX_SDF_ID_BAKED && !s.grid)
            return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "baked SDF selected but no grid was uploaded", __FILE__, __LINE__);
        if (d->NS > 0) BBX_LAUNCH(d, k_sdf_spheres, BBX_NUM_SMS * 8, 256, 0, *d, s);
        if (d->NB > 0) BBX_LAUNCH(d, k_sdf_boxes, BBX_NUM_SMS * 8, 128, 0, *d, s);
    }
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}
