// bbx_geom.cuh -- narrowphase geometry shared by the kernels (bbx_narrowphase.cu, bbx_broadphase.cu)
// and the host-side standalone queries of the public API (bbx_queries.cu): one source, the same
// IEEE binary32 operations in the reference's order on both sides (ARITHMETIC CONTRACT, bbx_dev.cuh;
// the host side is compiled without FMA contraction as well).
//
//   sphere-sphere         CheckSphereCollisionWithData      ballbox.h:451-471
//   sphere-box            CheckSphereBoxCollisionWithData   ballbox.h:1565-1626
//   SAT + contact point   SATCollision                      ballbox.h:1040-1111
//                         CheckBoxCollisionWithData         ballbox.h:1117-1132
//   manifold              GenerateBoxBoxManifold            ballbox.h:2550-2724
//                         SelectOptimalContactPoints        ballbox.h:2726-2843
#pragma once
#include "bbx_dev.cuh"

// ---------------------------------------------------------------------------
// sphere - sphere (:451-471): hit iff distance < rA + rB (strict)
// ---------------------------------------------------------------------------
BBX_HDI bool sphere_sphere_contact(float3 pa, float ra, float3 pb, float rb, float3& pt, float3& nrm, float& pen)
{
    float3 n = v3sub(pb, pa);
    float dist = v3len(n);
    float rs = ra + rb;
    if (!(dist < rs)) return false;
    nrm = (dist > 1e-6f) ? v3scale(n, 1.0f / dist) : f3(1.0f, 0.0f, 0.0f);
    pt = v3add(pa, v3scale(nrm, ra));
    pen = rs - dist;
    return true;
}

// ---------------------------------------------------------------------------
// sphere - box
// ---------------------------------------------------------------------------
BBX_HDI bool sphere_box_contact(float3 sp, float sr, float3 bp, float3 bh, float4 bq,
                                                   float3& pt, float3& nrm, float& pen)
{
    float3 rel = v3sub(sp, bp);
    float3 loc = qrot(qconj(bq), rel);
    float3 cp = f3(bb_fmaxf(-bh.x, bb_fminf(loc.x, bh.x)),
                   bb_fmaxf(-bh.y, bb_fminf(loc.y, bh.y)),
                   bb_fmaxf(-bh.z, bb_fminf(loc.z, bh.z)));
    float3 diff = v3sub(loc, cp);
    float d2 = v3dot(diff, diff);
    if (!(d2 <= sr * sr)) return false;
    float dist = sqrtf(d2);
    float3 n;
    if (dist > 1e-6f) {
        n = v3scale(diff, 1.0f / dist);
    } else {                                           // centre inside the box: axis of largest |local| (:1595-1603)
        float ax = fabsf(loc.x), ay = fabsf(loc.y), az = fabsf(loc.z);
        if (ax >= ay && ax >= az)      n = loc.x > 0 ? f3(1, 0, 0) : f3(-1, 0, 0);
        else if (ay >= az)             n = loc.y > 0 ? f3(0, 1, 0) : f3(0, -1, 0);
        else                           n = loc.z > 0 ? f3(0, 0, 1) : f3(0, 0, -1);
    }
    n = qrot(bq, n);
    float3 s2b = v3sub(bp, sp);
    if (v3dot(n, s2b) < 0.0f) n = v3scale(n, -1.0f);   // :1610-1614
    pt = v3add(sp, v3scale(n, -sr));                   // far side of the sphere (:1617-1619)
    nrm = n;
    pen = sr - dist;
    return true;
}

// ---------------------------------------------------------------------------
// box - box: SAT
// ---------------------------------------------------------------------------
struct Obb { float3 c; float3 h; float3 ax[3]; };

BBX_HDI float comp3(float3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

BBX_HDI Obb make_obb(float3 pos, float3 half, float4 q)            // CreateOBB :826-834
{
    Obb o; o.c = pos; o.h = half;
    o.ax[0] = qrot(q, f3(1, 0, 0)); o.ax[1] = qrot(q, f3(0, 1, 0)); o.ax[2] = qrot(q, f3(0, 0, 1));
    return o;
}

BBX_HDI void obb_vertices(const Obb& o, float3 v[8])               // GetOBBVertices :837-850
{
    float3 hX = v3scale(o.ax[0], o.h.x), hY = v3scale(o.ax[1], o.h.y), hZ = v3scale(o.ax[2], o.h.z);
    v[0] = v3add(v3add(v3add(o.c, hX), hY), hZ);
    v[1] = v3add(v3add(v3sub(o.c, hZ), hX), hY);
    v[2] = v3add(v3add(v3add(o.c, hX), hZ), v3scale(hY, -1.0f));
    v[3] = v3add(v3sub(v3add(o.c, hX), hY), v3scale(hZ, -1.0f));
    v[4] = v3add(v3add(v3add(o.c, hY), hZ), v3scale(hX, -1.0f));
    v[5] = v3add(v3sub(v3add(o.c, hY), hZ), v3scale(hX, -1.0f));
    v[6] = v3add(v3sub(v3add(o.c, hZ), hX), v3scale(hY, -1.0f));
    v[7] = v3sub(v3sub(v3sub(o.c, hX), hY), hZ);
}

BBX_HDI void project_obb(const Obb& o, float3 axis, float& mn, float& mx)   // :853-861
{
    float cpj = v3dot(o.c, axis);
    float ext = o.h.x * fabsf(v3dot(o.ax[0], axis)) + o.h.y * fabsf(v3dot(o.ax[1], axis)) + o.h.z * fabsf(v3dot(o.ax[2], axis));
    mn = cpj - ext; mx = cpj + ext;
}

BBX_HDI bool overlap_on_axis(const Obb& a, const Obb& b, float3 axis, float& pd)   // :864-884
{
    float min1, max1, min2, max2;
    project_obb(a, axis, min1, max1);
    project_obb(b, axis, min2, max2);
    if (max1 < min2 || max2 < min1) return false;
    float ov = bb_fminf(max1, max2) - bb_fmaxf(min1, min2);
    pd = ov;
    if (min1 < min2) pd *= -1.0f;
    return true;
}

// VertexFaceCollision :908-950 (with IsPointInFaceBounds :891-900)
static __host__ __device__ __noinline__ float3 vertex_face(const Obb& vo, const Obb& fo, float& smallest)
{
    float3 vtx[8];
    obb_vertices(vo, vtx);
    float3 best = f3(0, 0, 0);
    smallest = 999999.0f;
    for (int i = 0; i < 3; i++) {
        float3 fn = fo.ax[i];
        float3 U = v3norm(fo.ax[(i + 1) % 3]);
        float3 V = v3norm(fo.ax[(i + 2) % 3]);
        float uH = comp3(fo.h, (i + 1) % 3), vH = comp3(fo.h, (i + 2) % 3);
        for (int dir = -1; dir <= 1; dir += 2) {
            float3 dn = v3scale(fn, (float)dir);
            float off = comp3(fo.h, i) * (float)dir;
            float3 fc = v3add(fo.c, v3scale(fn, off));
            for (int j = 0; j < 8; j++) {
                float sd = v3dot(v3sub(vtx[j], fc), dn);
                if (sd < 0.0f) {
                    float pdep = -sd;
                    float3 rel = v3sub(vtx[j], fc);
                    float uc = v3dot(rel, U), vc = v3dot(rel, V);
                    if (fabsf(uc) <= uH && fabsf(vc) <= vH) {
                        if (pdep < smallest) { smallest = pdep; best = vtx[j]; }
                    }
                }
            }
        }
    }
    return best;
}

// SquaredDistanceBetweenEdges :953-990
BBX_HDI float sq_dist_edges(float3 p1, float3 q1, float3 p2, float3 q2)
{
    float3 d1 = v3sub(q1, p1), d2 = v3sub(q2, p2), r = v3sub(p1, p2);
    float a = v3dot(d1, d1), e = v3dot(d2, d2), f = v3dot(d2, r);
    float s, t;
    if (a <= 1e-6f && e <= 1e-6f) return v3dot(r, r);
    if (a <= 1e-6f) {
        s = 0.0f;
        t = bb_fmaxf(0.0f, bb_fminf(1.0f, f / e));
    } else {
        float c = v3dot(d1, r);
        if (e <= 1e-6f) {
            t = 0.0f;
            s = bb_fmaxf(0.0f, bb_fminf(1.0f, -c / a));
        } else {
            float b = v3dot(d1, d2);
            float denom = a * e - b * b;
            if (denom != 0.0f) s = bb_fmaxf(0.0f, bb_fminf(1.0f, (b * f - c * e) / denom));
            else s = 0.0f;
            t = bb_fmaxf(0.0f, bb_fminf(1.0f, (b * s + f) / e));
        }
    }
    float3 c1 = v3add(p1, v3scale(d1, s)), c2 = v3add(p2, v3scale(d2, t));
    float3 df = v3sub(c1, c2);
    return v3dot(df, df);
}

// ClosestPointBetweenEdges :993-1034: whatever edge wins, the result is the
// midpoint of the two centres -- unless every tested edge pair is >= 999999 apart.
static __host__ __device__ __noinline__ float3 closest_point_between_edges(const Obb& o1, const Obb& o2)
{
    float3 v1[8], v2[8];
    obb_vertices(o1, v1);
    obb_vertices(o2, v2);
    float3 e1[3] = { v3sub(v1[1], v1[0]), v3sub(v1[2], v1[0]), v3sub(v1[4], v1[0]) };
    float3 e2[3] = { v3sub(v2[1], v2[0]), v3sub(v2[2], v2[0]), v3sub(v2[4], v2[0]) };
    float mind = 999999.0f;
    float3 cp = f3(0, 0, 0);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float d = sq_dist_edges(v1[0], v3add(v1[0], e1[i]), v2[0], v3add(v2[0], e2[j]));
            // the first pair closer than 999999 fixes the result (later pairs can only assign the same midpoint)
            if (d < mind) return v3scale(v3add(o1.c, o2.c), 0.5f);
        }
    return cp;
}

// SATCollision :1040-1086: the 15-axis test.  Returns false at the first separating axis.
BBX_HDI bool sat_axes(const Obb& A, const Obb& B, float& minPd, float3& smallest, int& axisType)
{
    minPd = 999999.0f;
    smallest = f3(0, 0, 0);
    axisType = -1;
    int idx = 0;
#define SAT_TEST(AXIS) do { float3 ax_ = (AXIS); float pd_ = 0.0f; \
        if (!overlap_on_axis(A, B, ax_, pd_)) return false; \
        if (fabsf(pd_) < fabsf(minPd)) { minPd = pd_; smallest = v3scale(ax_, (pd_ < 0) ? -1.0f : 1.0f); axisType = idx; } \
        idx++; } while (0)
    // the reference first builds the (compacted) axis list, then tests in order; the
    // outcome of every test is independent of the others, so interleaving is equivalent
    for (int i = 0; i < 3; i++) SAT_TEST(A.ax[i]);
    for (int i = 0; i < 3; i++) SAT_TEST(B.ax[i]);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float3 cr = v3cross(A.ax[i], B.ax[j]);
            float l = v3len(cr);
            if (l > 1e-6f) SAT_TEST(v3scale(cr, 1.0f / l));
        }
#undef SAT_TEST
    return true;
}

// The six face-axis tests of SATCollision alone (same overlap_on_axis, same order): false = separated.  A pair that
// fails here fails sat_axes at the same axis, so filtering with it first changes nothing.
BBX_HDI bool sat_face_axes_overlap(const Obb& A, const Obb& B)
{
    float pd;
    for (int i = 0; i < 3; i++) if (!overlap_on_axis(A, B, A.ax[i], pd)) return false;
    for (int i = 0; i < 3; i++) if (!overlap_on_axis(A, B, B.ax[i], pd)) return false;
    return true;
}

// SATCollision :1088-1108 (contact point by axis type) + the normal flip of
// CheckBoxCollisionWithData :1121-1129
BBX_HDI void sat_point(const Obb& A, const Obb& B, int axisType, float3& nrm, float3& pt)
{
    if (axisType >= 0 && axisType <= 5) {
        float d1, d2;
        float3 p1 = vertex_face(A, B, d1);
        float3 p2 = vertex_face(B, A, d2);
        pt = (d1 < d2) ? p1 : p2;
    } else if (axisType >= 6) {
        pt = closest_point_between_edges(A, B);
    } else {
        pt = v3scale(v3add(A.c, B.c), 0.5f);
    }
    float3 c2c = v3sub(B.c, A.c);
    if (v3dot(nrm, c2c) < 0.0f) nrm = v3scale(nrm, -1.0f);
}

// is `p` inside box (pos, half, q) grown by tol?  returns the minimum face slack
BBX_HDI bool inside_box(float3 p, float3 pos, float3 half, float4 q, float tol, float& slack)
{
    float3 loc = qrot(qconj(q), v3sub(p, pos));
    if (fabsf(loc.x) <= half.x + tol && fabsf(loc.y) <= half.y + tol && fabsf(loc.z) <= half.z + tol) {
        float px = half.x - fabsf(loc.x), py = half.y - fabsf(loc.y), pz = half.z - fabsf(loc.z);
        slack = bb_fminf(px, bb_fminf(py, pz));
        return true;
    }
    return false;
}

// GenerateBoxBoxManifold + SelectOptimalContactPoints.  Returns the number of
// manifold points (0 = no collision).
// `normal` enters as the SAT axis (before the A->B flip), `spen` and `axisType` from sat_axes.
static __host__ __device__ inline int box_box_manifold(float3 posA, float3 hA, float4 qA, float3 posB, float3 hB, float4 qB,
                                float tol, float pen_tol, int max_contacts, float spen, int axisType,
                                float3 out_pt[8], float out_pen[8], float3& normal)
{
    Obb A = make_obb(posA, hA, qA), B = make_obb(posB, hB, qB);
    float3 cpt[24]; float cpen[24]; int cc = 0;
    float3 spt;
    sat_point(A, B, axisType, normal, spt);
    cpt[0] = spt; cpen[0] = spen; cc = 1;                          // :2583-2585
    float3 vA[8], vB[8];
    obb_vertices(A, vA);
    obb_vertices(B, vB);
    for (int i = 0; i < 8 && cc < 24; i++) {                       // :2595-2622
        float sl;
        if (inside_box(vA[i], posB, hB, qB, tol, sl) && sl > pen_tol) { cpt[cc] = vA[i]; cpen[cc] = sl; cc++; }
    }
    for (int i = 0; i < 8 && cc < 24; i++) {                       // :2625-2652
        float sl;
        if (inside_box(vB[i], posA, hA, qA, tol, sl) && sl > pen_tol) { cpt[cc] = vB[i]; cpen[cc] = sl; cc++; }
    }
    if (cc < 3) {                                                  // :2656-2714, edge table copied literally
        for (int e = 0; e < 12 && cc < 24; e++) {
            float3 s, t;
            if (e < 4)      { s = vA[e];     t = vA[(e + 1) % 4]; }
            else if (e < 8) { s = vA[e];     t = vA[4 + ((e - 4 + 1) % 4)]; }
            else            { s = vA[e - 8]; t = vA[e - 8 + 4]; }
            float3 mid = v3scale(v3add(s, t), 0.5f);
            float sl;
            if (inside_box(mid, posB, hB, qB, tol, sl) && sl > pen_tol) { cpt[cc] = mid; cpen[cc] = sl; cc++; }
        }
    }
    // ---- SelectOptimalContactPoints :2726-2843 ----
    if (cc <= 2) {
        for (int i = 0; i < cc; i++) { out_pt[i] = cpt[i]; out_pen[i] = cpen[i]; }
        return cc;
    }
    int deep = 0;
    for (int i = 1; i < cc; i++) if (cpen[i] > cpen[deep]) deep = i;
    out_pt[0] = cpt[deep]; out_pen[0] = cpen[deep];
    if (cc <= max_contacts) {
        int k = 1;
        for (int i = 0; i < cc; i++) if (i != deep) { out_pt[k] = cpt[i]; out_pen[k] = cpen[i]; k++; }
        return cc;
    }
    float3 centre = f3(0, 0, 0);
    for (int i = 0; i < cc; i++) centre = v3add(centre, cpt[i]);
    centre = v3scale(centre, 1.0f / (float)cc);
    int rem = max_contacts - 1;
    float bestd[8]; int besti[8];
    for (int j = 0; j < rem; j++) { bestd[j] = -1.0f; besti[j] = -1; }
    for (int i = 0; i < cc; i++) {
        if (i == deep) continue;
        float dist = v3len(v3sub(cpt[i], centre));
        for (int j = 0; j < rem; j++) {
            if (dist > bestd[j]) {
                for (int k = rem - 1; k > j; k--) { bestd[k] = bestd[k - 1]; besti[k] = besti[k - 1]; }
                bestd[j] = dist; besti[j] = i;
                break;
            }
        }
    }
    int np = 1;
    for (int j = 0; j < rem; j++) if (besti[j] >= 0) { out_pt[np] = cpt[besti[j]]; out_pen[np] = cpen[besti[j]]; np++; }
    return np;
}

