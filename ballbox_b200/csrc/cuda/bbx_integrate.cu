// bbx_integrate.cu -- K1 (forces + velocity integration + damping + force clear)
// and K12 (position integration + sleep-state update), one thread per body slot,
// float4 loads/stores, no inter-thread communication: pure HBM streaming.
//
// Reference: ApplyForces ballbox.h:1820-1838, IntegrateVelocities :1841-1925,
//            IntegratePositions :1928-1977, UpdateSleepStates :1696-1784.
#include "bbx_dev.cuh"

// ---------------------------------------------------------------------------
// K1.  The reference runs four passes (integrate, damp, clear) over the arrays;
// per body they compose to: F += g*m (non-static); if awake: v += (F*(1/m))*dt,
// w += alpha*dt, v *= lin_damp, w *= ang_damp; F = T = 0 for everybody.
// HAS_FORCES=false is the resident fast path: the force/torque arrays are known
// to hold zeros, so F = 0 + g*m and T = 0 without touching them (24+24 bytes per
// body saved each way).
// ---------------------------------------------------------------------------
// GRAVITY / INTEGRATE select the two public stages separately (ApplyForces, IntegrateVelocities);
// PhysicsStep runs both fused.
template <bool HAS_FORCES, bool GRAVITY, bool INTEGRATE>
__global__ void __launch_bounds__(256) k_integrate_velocities(BbxDev dv, float3 g, float dt, float lin_damp, float ang_damp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dv.N) return;
    unsigned char f = dv.flags[i];
    if (!(f & BF_VALID)) return;
    bool is_box = i >= dv.NS;
    int b = i - dv.NS;
    float3 F = f3(0.0f, 0.0f, 0.0f), T = f3(0.0f, 0.0f, 0.0f);
    if (HAS_FORCES) {
        const float* fp = is_box ? dv.r_b_force + 3 * b : dv.r_s_force + 3 * i;
        const float* tp = is_box ? dv.r_b_torque + 3 * b : dv.r_s_torque + 3 * i;
        F = f3(fp[0], fp[1], fp[2]);
        T = f3(tp[0], tp[1], tp[2]);
    }
    if (!(f & BF_STATIC)) {
        float m = dv.mass[i];
        // ApplyForces: forces += gravity * mass, sleeping bodies included (:1825-1827)
        if (GRAVITY) F = v3add(F, v3scale(g, m));
        if (INTEGRATE && !(f & BF_SLEEP)) {
            float4 v4 = dv.vel[2 * i], w4 = dv.vel[2 * i + 1];
            float3 v = xyz(v4), w = xyz(w4);
            float3 acc = v3scale(F, 1.0f / m);                       // :1852
            v = v3add(v, v3scale(acc, dt));                          // :1855
            float3 alpha;
            if (!is_box) {
                float inertia = HAS_FORCES ? dv.r_s_inertia[i] : 1.0f;
                alpha = HAS_FORCES ? v3scale(T, 1.0f / inertia) : f3(0.0f, 0.0f, 0.0f);   // :1860
            } else if (HAS_FORCES) {
                const float* ip = dv.r_b_inertia + 3 * b;            // world-axis, component-wise (:1886-1891)
                alpha.x = (ip[0] > 1e-6f) ? T.x / ip[0] : 0.0f;
                alpha.y = (ip[1] > 1e-6f) ? T.y / ip[1] : 0.0f;
                alpha.z = (ip[2] > 1e-6f) ? T.z / ip[2] : 0.0f;
            } else {
                alpha = f3(0.0f, 0.0f, 0.0f);
            }
            w = v3add(w, v3scale(alpha, dt));                        // :1863 / :1894
            v = v3scale(v, lin_damp);                                // :1905 / :1912
            w = v3scale(w, ang_damp);
            dv.vel[2 * i] = make_float4(v.x, v.y, v.z, v4.w);
            dv.vel[2 * i + 1] = make_float4(w.x, w.y, w.z, w4.w);
        }
    }
    if (HAS_FORCES && INTEGRATE) {                                   // :1917-1924
        float* fp = is_box ? dv.r_b_force + 3 * b : dv.r_s_force + 3 * i;
        float* tp = is_box ? dv.r_b_torque + 3 * b : dv.r_s_torque + 3 * i;
        fp[0] = 0.0f; fp[1] = 0.0f; fp[2] = 0.0f;
        tp[0] = 0.0f; tp[1] = 0.0f; tp[2] = 0.0f;
    }
    if (HAS_FORCES && !INTEGRATE) {                                  // ApplyForces alone: publish F
        float* fp = is_box ? dv.r_b_force + 3 * b : dv.r_s_force + 3 * i;
        fp[0] = F.x; fp[1] = F.y; fp[2] = F.z;
    }
}

int bbx_stage_integrate_velocities(BbxDev* d, const BbxStepParams* p)
{
    float3 g = make_float3(p->gravity[0], p->gravity[1], p->gravity[2]);
    int nb = bbx_blocks(d->N, 256);
    if (p->has_forces)
        BBX_LAUNCH(d, (k_integrate_velocities<true, true, true>), nb, 256, 0, *d, g, p->dt, p->settings.linear_damping, p->settings.angular_damping);
    else
        BBX_LAUNCH(d, (k_integrate_velocities<false, true, true>), nb, 256, 0, *d, g, p->dt, p->settings.linear_damping, p->settings.angular_damping);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

// the public stages ApplyForces (:1820) and IntegrateVelocities (:1841) on their own
int bbx_stage_apply_forces_only(BbxDev* d, const BbxStepParams* p)
{
    float3 g = make_float3(p->gravity[0], p->gravity[1], p->gravity[2]);
    BBX_LAUNCH(d, (k_integrate_velocities<true, true, false>), bbx_blocks(d->N, 256), 256, 0, *d, g, p->dt, 1.0f, 1.0f);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}
int bbx_stage_integrate_velocities_only(BbxDev* d, const BbxStepParams* p)
{
    float3 g = make_float3(0.0f, 0.0f, 0.0f);
    BBX_LAUNCH(d, (k_integrate_velocities<true, false, true>), bbx_blocks(d->N, 256), 256, 0, *d, g, p->dt,
               p->settings.linear_damping, p->settings.angular_damping);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// K12.  IntegratePositions then UpdateSleepStates, fused per body.
// ---------------------------------------------------------------------------
template <bool POSITIONS, bool SLEEP>
__global__ void __launch_bounds__(256) k_integrate_positions_sleep(BbxDev dv, float dt, float lin_thr, float ang_thr, float t_req)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dv.N) return;
    unsigned char f = dv.flags[i];
    if (!(f & BF_VALID) || (f & BF_STATIC)) return;
    float4 v4 = dv.vel[2 * i], w4 = dv.vel[2 * i + 1];
    float3 v = xyz(v4), w = xyz(w4);
    if (POSITIONS && !(f & BF_SLEEP)) {
        float4 p4 = dv.pos[i];
        float3 p = v3add(xyz(p4), v3scale(v, dt));                   // :1939 / :1950
        float R = p4.w;
        if (i >= dv.NS) {
            int b = i - dv.NS;
            float4 q = dv.boxq[2 * b];
            // q_dot = 0.5 * (w,0) (x) q  (QuatMultiply :543-550 with q1 = (w.x, w.y, w.z, 0))
            float dx = 0.0f * q.x + w.x * q.w + w.y * q.z - w.z * q.y;
            float dy = 0.0f * q.y - w.x * q.z + w.y * q.w + w.z * q.x;
            float dz = 0.0f * q.z + w.x * q.y - w.y * q.x + w.z * q.w;
            float dw = 0.0f * q.w - w.x * q.x - w.y * q.y - w.z * q.z;
            dx *= 0.5f; dy *= 0.5f; dz *= 0.5f; dw *= 0.5f;
            float4 n = make_float4(q.x + dx * dt, q.y + dy * dt, q.z + dz * dt, q.w + dw * dt);
            float len = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z + n.w * n.w);     // QuatNormalize :552-562
            if (len == 0.0f) n = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
            else n = make_float4(n.x / len, n.y / len, n.z / len, n.w / len);
            dv.boxq[2 * b] = n;
            R = box_bound_radius(xyz(dv.half[b]), n);
        }
        dv.pos[i] = make_float4(p.x, p.y, p.z, R);
    }
    if (!SLEEP) return;
    // UpdateSleepStates :1714-1742 / :1745-1783
    float ls = v3len(v), as = v3len(w);
    float t = dv.timer[i];
    bool low = (ls < lin_thr && as < ang_thr);
    if (low) {
        t += dt;
        if (t >= t_req && !(f & BF_SLEEP)) {
            f |= BF_SLEEP;
            dv.vel[2 * i] = make_float4(0.0f, 0.0f, 0.0f, v4.w);
            dv.vel[2 * i + 1] = make_float4(0.0f, 0.0f, 0.0f, w4.w);
        }
    } else {
        t = 0.0f;
        f &= ~BF_SLEEP;
    }
    dv.timer[i] = t;
    dv.flags[i] = f;
}

int bbx_stage_integrate_positions(BbxDev* d, const BbxStepParams* p)
{
    BBX_LAUNCH(d, (k_integrate_positions_sleep<true, true>), bbx_blocks(d->N, 256), 256, 0, *d, p->dt,
               p->settings.sleep_linear_threshold, p->settings.sleep_angular_threshold, p->settings.sleep_time_required);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

// the public stages IntegratePositions (:1928) and UpdateSleepStates (:1696) on their own
int bbx_stage_positions_only(BbxDev* d, const BbxStepParams* p)
{
    BBX_LAUNCH(d, (k_integrate_positions_sleep<true, false>), bbx_blocks(d->N, 256), 256, 0, *d, p->dt, 0.0f, 0.0f, 0.0f);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}
int bbx_stage_sleep_only(BbxDev* d, const BbxStepParams* p)
{
    BBX_LAUNCH(d, (k_integrate_positions_sleep<false, true>), bbx_blocks(d->N, 256), 256, 0, *d, p->dt,
               p->settings.sleep_linear_threshold, p->settings.sleep_angular_threshold, p->settings.sleep_time_required);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}
