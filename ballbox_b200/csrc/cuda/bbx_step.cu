// bbx_step.cu -- the step driver (reference PhysicsStep, ballbox.h:1787-1817) and
// the export of world->constraints in the reference's emission order.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "bbx_dev.cuh"
int bbx_ensure_slot_of(BbxDev* d);             // bbx_solver.cu
#include <nvtx3/nvToolsExt.h>          // header-only; ranges are pushed only with BBX_NVTX=1

void bbx_timing_mark(BbxDev* d, int mark)
{
    if (!d->timing_on || d->timing_steps >= d->timing_cap) return;
    cudaEventRecord(d->timing_ev[d->timing_steps * BBX_TIMING_MARKS + mark], d->stream);
}

// device-side failure flags of the steps so far (ctr_host must be current): a step that dropped contacts or could
// not schedule its constraints has produced wrong state, which every download path reports instead of handing it out
int bbx_check_overflow(BbxDev* d)
{
    const int f = d->ctr_host->overflow;
    if (f & 3)  return bbx_fail(d, BBX_ERR_OVERFLOW, "contact or candidate buffer overflow: raise max_collisions", __FILE__, __LINE__);
    if (f & 16) return bbx_fail(d, BBX_ERR_OVERFLOW, "solver dependency graph deeper than the level table", __FILE__, __LINE__);
    if (f & 32) return bbx_fail(d, BBX_ERR_OVERFLOW, "batched worlds: a world's dependency graph or contact list exceeds the per-group tables", __FILE__, __LINE__);
    if (f & 64) return bbx_fail(d, BBX_ERR_OVERFLOW, "flow solver gave up waiting for a predecessor (corrupt schedule)", __FILE__, __LINE__);
    return 0;
}

extern "C" void bbx_dev_set_timing(BbxDev* d, int on)
{
    if (!d) return;
    if (on && !d->timing_ev) {
        d->timing_cap = BBX_TIMING_STEPS;
        d->timing_ev = (cudaEvent_t*)calloc((size_t)d->timing_cap * BBX_TIMING_MARKS, sizeof(cudaEvent_t));
        d->timing_has7 = (unsigned char*)calloc((size_t)d->timing_cap, 1);
        cudaSetDevice(d->device);
        for (int i = 0; i < d->timing_cap * BBX_TIMING_MARKS; i++) cudaEventCreate(&d->timing_ev[i]);
    }
    d->timing_on = on; d->timing_steps = 0;
}

extern "C" int bbx_dev_get_timing(BbxDev* d, float* out6, int* steps)
{
    if (!d || !out6) return BBX_ERR_BAD_ARGUMENT;
    for (int k = 0; k < 6; k++) out6[k] = 0.0f;
    if (steps) *steps = 0;
    if (!d->timing_ev) return 0;
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    for (int s = 0; s < d->timing_steps; s++)
        for (int k = 0; k < 6; k++) {
            float ms = 0.0f;
            BBX_CUDA(d, cudaEventElapsedTime(&ms, d->timing_ev[s * BBX_TIMING_MARKS + k], d->timing_ev[s * BBX_TIMING_MARKS + k + 1]));
            out6[k] += ms;
        }
    d->replay_ms_sum = 0.0f; d->replay_steps = 0;
    for (int s = 0; s < d->timing_steps; s++)
        if (d->timing_has7[s]) {
            float ms = 0.0f;
            BBX_CUDA(d, cudaEventElapsedTime(&ms, d->timing_ev[s * BBX_TIMING_MARKS + 7], d->timing_ev[s * BBX_TIMING_MARKS + 5]));
            d->replay_ms_sum += ms; d->replay_steps++;
            d->timing_has7[s] = 0;
        }
    if (steps) *steps = d->timing_steps;
    d->timing_steps = 0;
    return 0;
}

// k_replay_staged alone (iterations 1..I-1 of the level schedule): ms summed over the steps of the last bbx_dev_get_timing call
extern "C" int bbx_dev_get_replay_timing(BbxDev* d, float* ms, int* steps)
{
    if (!d || !ms) return BBX_ERR_BAD_ARGUMENT;
    *ms = d->replay_ms_sum;
    if (steps) *steps = d->replay_steps;
    return 0;
}

extern "C" int bbx_dev_step(BbxDev* d, const BbxStepParams* p, int steps)
{
    if (!d || !p) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    if (steps <= 0) return 0;
    if (p->dt <= 0.0f || p->dt < 1e-6f) return 0;                    // ballbox.h:1791
    BBX_CUDA(d, cudaSetDevice(d->device));
    d->last_params = *p; d->have_params = 1;
    BbxStepParams q = *p;
    d->imported_list = 0;
    if (d->forces_on_device) { q.has_forces = 1; d->forces_on_device = 0; }
    // BBX_NVTX=1: one NVTX range per stage on the enqueueing thread (tracing)
    static const int nvtx_on = (getenv("BBX_NVTX") && getenv("BBX_NVTX")[0] == '1') ? 1 : 0;
#define BBX_RANGE(name) do { if (nvtx_on) { nvtxRangePop(); nvtxRangePushA(name); } } while (0)
    for (int s = 0; s < steps; s++) {
        int rc;
        if (nvtx_on) nvtxRangePushA("bbx:integrate_velocities");
        bbx_timing_mark(d, 0);
        if (d->n_springs > 0) {                                        // extension: spring forces join the accumulators before gravity
            if ((rc = bbx_springs_apply(d))) return rc;
            q.has_forces = 1;
        }
        if ((rc = bbx_stage_integrate_velocities(d, &q))) return rc;   // stages 1-2
        bbx_timing_mark(d, 1);
        BBX_RANGE("bbx:broadphase");
        if ((rc = bbx_stage_broadphase(d, &q))) return rc;             // stage 3: pair finding (+ sphere-sphere)
        bbx_timing_mark(d, 2);
        BBX_RANGE("bbx:narrowphase");
        if ((rc = bbx_stage_narrowphase(d, &q))) return rc;            // stage 3: contact generation
        if (q.want_callbacks && (rc = bbx_stage_collect_events(d, &q))) return rc;
        bbx_timing_mark(d, 3);
        BBX_RANGE("bbx:presolve_graph_solve");
        if ((rc = bbx_stage_solve(d, &q))) return rc;                  // stage 4 (marks 4 inside, before the solve kernel)
        if (q.warm_start > 0.0f && (rc = bbx_stage_warm_save(d, &q))) return rc;
        bbx_timing_mark(d, 5);
        BBX_RANGE("bbx:integrate_positions_sleep");
        if ((rc = bbx_stage_integrate_positions(d, &q))) return rc;    // stages 5-6
        bbx_timing_mark(d, 6);
        if (nvtx_on) nvtxRangePop();
        if (d->timing_on && d->timing_steps < d->timing_cap) d->timing_steps++;
        q.has_forces = 0;                                              // forces were consumed and cleared by step 0
        d->steps++;
        d->step_in_batch++;
        d->body_steps += (long long)d->n_valid_spheres + d->n_valid_boxes;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// export of the constraint list (reference layout, 104-byte records)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void split_gid(const BbxDev& dv, int g, int& type, int& world, int& local)
{
    if (g < 0) { type = 2; world = -1; local = 0; return; }
    if (g < dv.NS) { type = 0; world = g / dv.cap_s; local = g - world * dv.cap_s; }
    else { int b = g - dv.NS; type = 1; world = b / dv.cap_b; local = b - world * dv.cap_b; }
}

// key = world | phase | i | j | k packed into the fewest bits (fewer radix passes)
struct KeyBits { int wb, ib, jb; };
static int bits_for(int n) { int b = 1; while ((1ll << b) < n) b++; return b; }

// accumulated impulses of contact c as the step left them
__device__ __forceinline__ void final_impulses(const BbxDev& dv, int c, int4 id, float4 l3, int solved, float friction,
                                               float& jn, float& jt0, float& jt1)
{
    // not scheduled (or no sweep ran): the accumulated impulses are whatever the record held (0 in PhysicsStep)
    jn = l3.x; jt0 = l3.y; jt1 = l3.z;
    int slot = solved ? dv.slot_of[c] : -1;
    if (slot >= 0 && dv.last_flow) {
        // flow solver: (value, tag) words in queue order
        ulonglong2 i0 = dv.fimp[2 * (size_t)slot], i1 = dv.fimp[2 * (size_t)slot + 1];
        jn = __uint_as_float((unsigned int)i0.x); jt0 = __uint_as_float((unsigned int)i0.y); jt1 = __uint_as_float((unsigned int)i1.x);
    }
    else if (slot >= 0) { float4 m3 = dv.M[3 * (size_t)dv.cap_contacts + slot]; jn = m3.x; jt0 = m3.y; jt1 = m3.z; }
    else if (solved && friction > 0.0f) {
        // a constraint between two non-dynamic bodies is never scheduled, but the reference still
        // sweeps it: with zero effective masses it ends at fmaxf(-0, fminf(+0, +-0)) = -0 (:1477).
        // (In the block-per-world path scheduled constraints keep their impulses in L: slot < 0 too.)
        const int h = c - RUN_K(id.z);
        int4 hd = dv.last_per_contact ? dv.node[2 * (size_t)h] : dv.node[h];           // level path: one record per contact
        if (dv.last_per_contact) hd.y = dv.node[2 * (size_t)h + 1].x;                   // per-constraint layout: one half per body
        if (hd.x < 0 && hd.y < 0) { jt0 = -0.0f; jt1 = -0.0f; }
    }
}

__global__ void __launch_bounds__(256) k_export_constraints(BbxDev dv, ContactConstraint* out, unsigned long long* keys,
                                                            float restitution, float friction, int solved, KeyBits kb)
{
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        int4 id = dv.c_ids[c];
        float4 pt = dv.c_pt[c], nr = dv.c_nrm[c];
        float4 l0 = dv.L[4 * c], l1 = dv.L[4 * c + 1], l2 = dv.L[4 * c + 2], l3 = dv.L[4 * c + 3];
        int ta, wa, la, tb, wb, lb;
        split_gid(dv, id.x, ta, wa, la);
        split_gid(dv, id.y, tb, wb, lb);
        float3 n3 = make_float3(nr.x, nr.y, nr.z), t1, t2;
        if (fabsf(n3.x) >= 0.57735f) t1 = f3(n3.y, -n3.x, 0.0f); else t1 = f3(0.0f, n3.z, -n3.y);
        t1 = v3norm(t1);
        t2 = v3cross(n3, t1);
        float jn, jt0, jt1;
        final_impulses(dv, c, id, l3, solved, friction, jn, jt0, jt1);
        ContactConstraint r;
        r.bodyA_type = ta; r.bodyA_index = la; r.bodyB_type = tb; r.bodyB_index = lb;
        r.contact_point.x = pt.x; r.contact_point.y = pt.y; r.contact_point.z = pt.z;
        r.contact_normal.x = nr.x; r.contact_normal.y = nr.y; r.contact_normal.z = nr.z;
        r.penetration = pt.w;
        r.normal_mass = l1.w; r.tangent_mass[0] = l2.w; r.tangent_mass[1] = l3.w;
        r.tangent[0].x = t1.x; r.tangent[0].y = t1.y; r.tangent[0].z = t1.z;
        r.tangent[1].x = t2.x; r.tangent[1].y = t2.y; r.tangent[1].z = t2.z;
        r.normal_impulse = jn; r.tangent_impulse[0] = jt0; r.tangent_impulse[1] = jt1;
        r.position_bias = l0.w; r.restitution = restitution; r.friction = friction;
        out[c] = r;
        // emission order: world, phase, i, j, k  
        unsigned long long key = (unsigned long long)wa;
        key = (key << 3) | (unsigned long long)id.w;
        key = (key << kb.ib) | (unsigned long long)la;
        key = (key << kb.jb) | (unsigned long long)lb;
        key = (key << 5) | (unsigned long long)RUN_K(id.z);
        keys[c] = key;
    }
}

// ---------------------------------------------------------------------------
// warm starting: remember this step's final impulses by contact identity
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_warm_save(BbxDev dv, int solved, float friction)
{
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        int4 id = dv.c_ids[c];
        float jn, jt0, jt1;
        final_impulses(dv, c, id, dv.L[4 * (size_t)c + 3], solved, friction, jn, jt0, jt1);
        const int wa = id.x < dv.NS ? id.x / dv.cap_s : (id.x - dv.NS) / dv.cap_b;
        const unsigned long long key = warm_key_of(id.x, id.y, RUN_K(id.z));
        unsigned int h = warm_hash(key) & dv.warm_mask;
        for (unsigned int probe = 0; probe <= dv.warm_mask; probe++, h = (h + 1) & dv.warm_mask) {
            unsigned long long old = atomicCAS(&dv.warm_key[h], BBX_WARM_EMPTY, key);
            if (old == BBX_WARM_EMPTY || old == key) {
                dv.warm_val[h] = make_float4(jn, jt0, jt1, __int_as_float(dv.warm_epoch[wa]));
                break;
            }
        }
    }
}

int bbx_stage_warm_save(BbxDev* d, const BbxStepParams* p)
{
    int rc = bbx_warm_ensure(d);
    if (rc) return rc;
    BBX_CUDA(d, cudaMemsetAsync(d->warm_key, 0xff, sizeof(unsigned long long) * ((size_t)d->warm_mask + 1), d->stream));
    { int rcs = bbx_ensure_slot_of(d); if (rcs) return rcs; }
    BBX_LAUNCH(d, k_warm_save, BBX_NUM_SMS * 8, 256, 0, *d, p->settings.solver_iterations > 0 ? 1 : 0, p->settings.friction);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// callback events
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long emission_key(const BbxDev& dv, int4 id, KeyBits kb, int& wa, int& ta, int& la, int& tb, int& lb)
{
    int wb;
    split_gid(dv, id.x, ta, wa, la);
    split_gid(dv, id.y, tb, wb, lb);
    unsigned long long key = (unsigned long long)wa;
    key = (key << 3) | (unsigned long long)id.w;
    key = (key << kb.ib) | (unsigned long long)la;
    key = (key << kb.jb) | (unsigned long long)lb;
    key = (key << 5) | (unsigned long long)RUN_K(id.z);
    return key;
}

__global__ void __launch_bounds__(256) k_collect_events(BbxDev dv, KeyBits kb, int step)
{
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    int n_round = (n + 31) & ~31;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_round; c += gridDim.x * blockDim.x) {
        bool fire = false;
        int4 id = make_int4(0, 0, 0, 0);
        if (c < n) {
            id = dv.c_ids[c];
            bool cb = (dv.flags[id.x] & BF_HASCB) || (id.y >= 0 && (dv.flags[id.y] & BF_HASCB));
            // box-box: once per pair with the first manifold point (:2034-2044); box-SDF: never (:2413-2477)
            fire = cb && id.w != PH_BSDF && (id.w != PH_BB || RUN_K(id.z) == 0);
        }
        int slot = warp_append(&dv.ctr->n_events, fire);
        if (!fire) continue;
        if (slot >= dv.cap_events) { atomicOr(&dv.ctr->overflow, 8); continue; }
        int wa, ta, la, tb, lb;
        dv.ev_key[slot] = emission_key(dv, id, kb, wa, ta, la, tb, lb);
        dv.ev_step[slot] = step;
        float4 pt = dv.c_pt[c], nr = dv.c_nrm[c];
        CollisionContact k;
        k.bodyA_type = ta; k.bodyA_index = la; k.bodyB_type = tb; k.bodyB_index = lb;
        k.contact_point.x = pt.x; k.contact_point.y = pt.y; k.contact_point.z = pt.z;
        k.contact_normal.x = nr.x; k.contact_normal.y = nr.y; k.contact_normal.z = nr.z;
        k.penetration = pt.w;
        dv.ev_contact[slot] = k;
    }
}

static KeyBits key_bits_of(const BbxDev* d)
{
    KeyBits kb;
    kb.wb = bits_for(d->n_worlds); kb.ib = bits_for(d->cap_s > d->cap_b ? d->cap_s : d->cap_b); kb.jb = kb.ib;
    return kb;
}

int bbx_stage_collect_events(BbxDev* d, const BbxStepParams* p)
{
    (void)p;
    if (!d->ev_key) {
        long long cap = d->cap_contacts < (1 << 22) ? d->cap_contacts : (1 << 22);
        d->cap_events = (int)cap;
        BBX_CUDA(d, cudaMalloc((void**)&d->ev_key, sizeof(unsigned long long) * (size_t)cap));
        BBX_CUDA(d, cudaMalloc((void**)&d->ev_step, sizeof(int) * (size_t)cap));
        BBX_CUDA(d, cudaMalloc((void**)&d->ev_contact, sizeof(CollisionContact) * (size_t)cap));
    }
    BBX_LAUNCH(d, k_collect_events, BBX_NUM_SMS * 8, 256, 0, *d, key_bits_of(d), d->step_in_batch);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

extern "C" int bbx_dev_event_capacity(BbxDev* d) { return d ? (d->cap_events ? d->cap_events : (d->cap_contacts < (1 << 22) ? d->cap_contacts : (1 << 22))) : 0; }

__global__ void k_reset_events(BbxDev dv) { dv.ctr->n_events = 0; }

extern "C" int bbx_dev_fetch_events(BbxDev* d, CollisionContact* contacts, int* steps, int* worlds, int max_events, int* count)
{
    if (!d || !count) return BBX_ERR_BAD_ARGUMENT;
    *count = 0;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    d->step_in_batch = 0;
    if (!d->ev_key) return 0;
    BBX_CUDA(d, cudaMemcpyAsync(d->ctr_host, d->ctr, sizeof(StepCounters), cudaMemcpyDeviceToHost, d->stream));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    int n = d->ctr_host->n_events;
    bool overflow = (d->ctr_host->overflow & 8) != 0 || n > d->cap_events;
    if (n > d->cap_events) n = d->cap_events;
    if (n > max_events) { n = max_events; overflow = true; }
    if (n > 0) {
        std::vector<unsigned long long> keys((size_t)n);
        std::vector<int> st((size_t)n);
        std::vector<CollisionContact> recs((size_t)n);
        BBX_CUDA(d, cudaMemcpyAsync(keys.data(), d->ev_key, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToHost, d->stream));
        BBX_CUDA(d, cudaMemcpyAsync(st.data(), d->ev_step, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, d->stream));
        BBX_CUDA(d, cudaMemcpyAsync(recs.data(), d->ev_contact, sizeof(CollisionContact) * (size_t)n, cudaMemcpyDeviceToHost, d->stream));
        BBX_CUDA(d, cudaStreamSynchronize(d->stream));
        std::vector<int> order((size_t)n);
        for (int i = 0; i < n; i++) order[i] = i;
        // deterministic replay order: step, then the reference's emission order
        std::sort(order.begin(), order.end(), [&](int a, int b) { return st[a] != st[b] ? st[a] < st[b] : keys[a] < keys[b]; });
        KeyBits kb = key_bits_of(d);
        int wshift = 3 + kb.ib + kb.jb + 5;
        for (int t = 0; t < n; t++) {
            int e = order[t];
            contacts[t] = recs[e];
            if (steps) steps[t] = st[e];
            if (worlds) worlds[t] = (int)(keys[e] >> wshift);
        }
    }
    BBX_LAUNCH(d, k_reset_events, 1, 1, 0, *d);
    BBX_CUDA(d, cudaGetLastError());
    *count = n;
    if (overflow) return bbx_fail(d, BBX_ERR_OVERFLOW, "callback event buffer overflow: fetch events more often", __FILE__, __LINE__);
    return 0;
}

// sorted[t] = records[order[t]], 26 floats per 104-byte record, coalesced on the write side
__global__ void __launch_bounds__(256) k_gather_records(const float* src, float* dst, const unsigned int* order, int n)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n * 26;
    for (; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int t = (int)(idx / 26), f = (int)(idx - (long long)t * 26);
        dst[idx] = src[(size_t)order[t] * 26 + f];
    }
}

static int ensure_export_buffers(BbxDev* d);
int bbx_copy_stream(BbxDev* d);
int bbx_radix_sort(BbxDev* d, unsigned long long* keys_a, unsigned int* vals_a, unsigned long long* keys_b,
                   unsigned int* vals_b, int n, int bits, int* hist, int* hist_scan,
                   unsigned long long** keys_sorted, unsigned int** vals_sorted);

extern "C" int bbx_dev_download_constraints(BbxDev* d, ContactConstraint* out, int cap_per_world, int* counts)
{
    if (!d || !out || !counts) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    for (int w = 0; w < d->n_worlds; w++) counts[w] = 0;
    if (!d->have_params) return 0;
    KeyBits kb;
    kb.wb = bits_for(d->n_worlds); kb.ib = bits_for(d->cap_s > d->cap_b ? d->cap_s : d->cap_b); kb.jb = kb.ib;
    int total_bits = kb.wb + 3 + kb.ib + kb.jb + 5;
    if (total_bits > 64) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "export key does not fit 64 bits", __FILE__, __LINE__);
    { int rc0 = ensure_export_buffers(d); if (rc0) return rc0; }
    const BbxStepParams* p = &d->last_params;
    { int rcs = bbx_ensure_slot_of(d); if (rcs) return rcs; }
    BBX_LAUNCH(d, k_export_constraints, BBX_NUM_SMS * 8, 256, 0, *d, d->export_buf, d->sort_keys,
               p->settings.restitution, p->settings.friction, p->settings.solver_iterations > 0 ? 1 : 0, kb);
    BBX_CUDA(d, cudaGetLastError());
    BBX_CUDA(d, cudaMemcpyAsync(d->ctr_host, d->ctr, sizeof(StepCounters), cudaMemcpyDeviceToHost, d->stream));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    { int rco = bbx_check_overflow(d); if (rco) return rco; }
    int n = d->ctr_host->n_contacts;
    if (n > d->cap_contacts) n = d->cap_contacts;
    if (n == 0) return 0;
    unsigned long long* ks; unsigned int* order;
    int rc = bbx_radix_sort(d, d->sort_keys, d->sort_vals_a, d->sort_keys_b, d->sort_vals_b, n, total_bits,
                            d->sort_hist, d->sort_hist_scan, &ks, &order);
    if (rc) return rc;
    if (d->n_worlds == 1) {
        if (n > cap_per_world)
            return bbx_fail(d, BBX_ERR_OVERFLOW, "the step produced more contacts than max_collisions", __FILE__, __LINE__);
        // gather and copy in chunks: chunk i crosses PCIe (copy stream) while chunk i+1 is gathered (step stream)
        { int rcs = bbx_copy_stream(d); if (rcs) return rcs; }
        const int CH = 1 << 19;                                    // 512k records = 54.5 MB, ~1 ms on the link
        for (int off = 0; off < n; off += CH) {
            const int m = n - off < CH ? n - off : CH;
            BBX_LAUNCH(d, k_gather_records, BBX_NUM_SMS * 16, 256, 0, (const float*)d->export_buf,
                       (float*)(d->export_sorted + off), order + off, m);
            BBX_CUDA(d, cudaEventRecord(d->ev_chain, d->stream));
            BBX_CUDA(d, cudaStreamWaitEvent(d->copy_stream, d->ev_chain, 0));
            BBX_CUDA(d, cudaMemcpyAsync(out + off, d->export_sorted + off, sizeof(ContactConstraint) * (size_t)m,
                                        cudaMemcpyDeviceToHost, d->copy_stream));
        }
        BBX_CUDA(d, cudaGetLastError());
        BBX_CUDA(d, cudaStreamSynchronize(d->copy_stream));
        // the next step (or export) on the step stream must not overwrite export_sorted early: it is ordered after the
        // host-side wait above
        counts[0] = n;
        return 0;
    }
    BBX_LAUNCH(d, k_gather_records, BBX_NUM_SMS * 16, 256, 0, (const float*)d->export_buf, (float*)d->export_sorted, order, n);
    BBX_CUDA(d, cudaGetLastError());
    // batch: sorted by world first; cut the list at world boundaries on the host
    std::vector<unsigned long long> keys((size_t)n);
    std::vector<ContactConstraint> recs((size_t)n);
    BBX_CUDA(d, cudaMemcpyAsync(keys.data(), ks, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToHost, d->stream));
    BBX_CUDA(d, cudaMemcpyAsync(recs.data(), d->export_sorted, sizeof(ContactConstraint) * (size_t)n, cudaMemcpyDeviceToHost, d->stream));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    int wshift = 3 + kb.ib + kb.jb + 5, truncated = 0;
    for (int t = 0; t < n; t++) {
        int w = (int)(keys[t] >> wshift);
        if (w < 0 || w >= d->n_worlds) continue;
        if (counts[w] < cap_per_world) out[(size_t)w * cap_per_world + counts[w]++] = recs[t];
        else truncated = 1;
    }
    if (truncated)
        return bbx_fail(d, BBX_ERR_OVERFLOW, "a world produced more contacts than max_collisions", __FILE__, __LINE__);
    return 0;
}

// ---------------------------------------------------------------------------
// public stage functions (reference ballbox.h:217-229): import / export of host lists
// ---------------------------------------------------------------------------
__global__ void k_import_reset(BbxDev dv, int n)
{
    StepCounters* c = dv.ctr;
    c->n_contacts = n; c->n_cand_sb = 0; c->n_cand_bb = 0; c->n_hit_bb = 0; c->n_hit_bb_edge = 0; c->n_surv_bb = 0; c->n_deferred = 0; c->n_solver = 0; c->n_touched = 0;
    c->n_levels = 0; c->n_frontier0 = 0; c->n_mslots = 0; c->n_nodes = 0; c->fq_tail = 0; c->fq_head = 0; c->fq_ticket = 0;
}

struct ImportRec { int ta, ia, tb, ib; };
__device__ __forceinline__ ImportRec import_ids(const CollisionContact* cs, const ContactConstraint* ks, int i)
{
    ImportRec r;
    if (cs) { r.ta = cs[i].bodyA_type; r.ia = cs[i].bodyA_index; r.tb = cs[i].bodyB_type; r.ib = cs[i].bodyB_index; }
    else    { r.ta = ks[i].bodyA_type; r.ia = ks[i].bodyA_index; r.tb = ks[i].bodyB_type; r.ib = ks[i].bodyB_index; }
    return r;
}
__device__ __forceinline__ bool same_pair(ImportRec a, ImportRec b) { return a.ta == b.ta && a.ia == b.ia && a.tb == b.tb && a.ib == b.ib; }

// Host list (array order) -> device contact arrays.  Consecutive records of the same body pair form
// one solver node (they are adjacent in both bodies' orders by construction); runs are cut at 255.
__global__ void __launch_bounds__(256) k_import_list(BbxDev dv, const CollisionContact* cs, const ContactConstraint* ks, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        ImportRec me = import_ids(cs, ks, i);
        int back = 0;
        while (back < 4096 && i - back - 1 >= 0 && same_pair(me, import_ids(cs, ks, i - back - 1))) back++;
        int fwd = 0;
        while (fwd < 4096 && i + fwd + 1 < n && same_pair(me, import_ids(cs, ks, i + fwd + 1))) fwd++;
        int k = back % 255;
        int chunk_start = back - k;
        int len = min(255, back + 1 + fwd - chunk_start);
        int a = (me.ta == 0) ? me.ia : ((me.ta == 1) ? dv.NS + me.ia : -2);
        int b = (me.tb == 0) ? me.ib : ((me.tb == 1) ? dv.NS + me.ib : -2);
        int phase = (me.ta == 0 && me.tb == 0) ? PH_SS : (me.ta == 1 && me.tb == 1) ? PH_BB : (me.ta == 0 && me.tb == 1) ? PH_SB
                  : (me.ta == 0 && me.tb == 2) ? PH_SSDF : (me.ta == 1 && me.tb == 2) ? PH_BSDF : 5;
        Vec3 pt = cs ? cs[i].contact_point : ks[i].contact_point;
        Vec3 nr = cs ? cs[i].contact_normal : ks[i].contact_normal;
        float pen = cs ? cs[i].penetration : ks[i].penetration;
        dv.c_ids[i] = make_int4(a, b, RUN_ENC(k, len), phase);
        dv.c_pt[i] = make_float4(pt.x, pt.y, pt.z, pen);
        dv.c_nrm[i] = make_float4(nr.x, nr.y, nr.z, 0.0f);
        float4* L = dv.L + 4 * (size_t)i;
        if (ks) {
            L[0] = make_float4(0, 0, 0, ks[i].position_bias);
            L[1] = make_float4(0, 0, 0, ks[i].normal_mass);
            L[2] = make_float4(0, 0, 0, ks[i].tangent_mass[0]);
            L[3] = make_float4(ks[i].normal_impulse, ks[i].tangent_impulse[0], ks[i].tangent_impulse[1], ks[i].tangent_mass[1]);
        } else {
            L[3] = make_float4(0, 0, 0, 0);
        }
    }
}

static int ensure_export_buffers(BbxDev* d)
{
    size_t C = (size_t)d->cap_contacts;
    if (!d->export_buf) {
        size_t tiles = C / 2048 + 2;
        BBX_CUDA(d, cudaMalloc((void**)&d->export_buf, sizeof(ContactConstraint) * C));
        BBX_CUDA(d, cudaMalloc((void**)&d->export_sorted, sizeof(ContactConstraint) * C));
        BBX_CUDA(d, cudaMalloc((void**)&d->sort_keys_b, sizeof(unsigned long long) * C));
        BBX_CUDA(d, cudaMalloc((void**)&d->sort_vals_a, sizeof(unsigned int) * C));
        BBX_CUDA(d, cudaMalloc((void**)&d->sort_vals_b, sizeof(unsigned int) * C));
        BBX_CUDA(d, cudaMalloc((void**)&d->sort_hist, sizeof(int) * 256 * tiles));
        BBX_CUDA(d, cudaMalloc((void**)&d->sort_hist_scan, sizeof(int) * 256 * tiles));
    }
    return 0;
}

extern "C" int bbx_dev_import_list(BbxDev* d, const CollisionContact* contacts, const ContactConstraint* constraints, int n)
{
    if (!d || n < 0 || ((contacts != NULL) == (constraints != NULL) && n > 0)) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    if (d->n_worlds != 1) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "stage functions work on single worlds", __FILE__, __LINE__);
    if (n > d->cap_contacts) return bbx_fail(d, BBX_ERR_OVERFLOW, "imported list longer than max_collisions", __FILE__, __LINE__);
    BBX_CUDA(d, cudaSetDevice(d->device));
    int rc = ensure_export_buffers(d);
    if (rc) return rc;
    const CollisionContact* dcs = NULL; const ContactConstraint* dks = NULL;
    if (n > 0) {
        if (contacts) { BBX_CUDA(d, cudaMemcpyAsync(d->export_sorted, contacts, sizeof(CollisionContact) * (size_t)n, cudaMemcpyHostToDevice, d->stream));
                        dcs = (const CollisionContact*)d->export_sorted; }
        else          { BBX_CUDA(d, cudaMemcpyAsync(d->export_sorted, constraints, sizeof(ContactConstraint) * (size_t)n, cudaMemcpyHostToDevice, d->stream));
                        dks = d->export_sorted; }
    }
    BBX_LAUNCH(d, k_import_reset, 1, 1, 0, *d, n);
    if (n > 0) BBX_LAUNCH(d, k_import_list, bbx_blocks(n, 256), 256, 0, *d, dcs, dks, n);
    BBX_CUDA(d, cudaGetLastError());
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));          // the host list may be pageable user memory
    d->imported_list = 1;
    return 0;
}

extern "C" int bbx_dev_export_list(BbxDev* d, ContactConstraint* out, int n, float restitution, float friction, int solved)
{
    if (!d || (n > 0 && !out)) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    if (n <= 0) return 0;
    BBX_CUDA(d, cudaSetDevice(d->device));
    int rc = ensure_export_buffers(d);
    if (rc) return rc;
    KeyBits kb = key_bits_of(d);
    { int rcs = bbx_ensure_slot_of(d); if (rcs) return rcs; }
    BBX_LAUNCH(d, k_export_constraints, BBX_NUM_SMS * 8, 256, 0, *d, d->export_buf, d->sort_keys, restitution, friction, solved, kb);
    BBX_CUDA(d, cudaGetLastError());
    BBX_CUDA(d, cudaMemcpyAsync(out, d->export_buf, sizeof(ContactConstraint) * (size_t)n, cudaMemcpyDeviceToHost, d->stream));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    return 0;
}

__global__ void __launch_bounds__(256) k_export_contacts(BbxDev dv, CollisionContact* out, unsigned long long* keys, KeyBits kb)
{
    int n = min(dv.ctr->n_contacts, dv.cap_contacts);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        int4 id = dv.c_ids[c];
        int wa, ta, la, tb, lb;
        keys[c] = emission_key(dv, id, kb, wa, ta, la, tb, lb);
        float4 pt = dv.c_pt[c], nr = dv.c_nrm[c];
        CollisionContact k;
        k.bodyA_type = ta; k.bodyA_index = la; k.bodyB_type = tb; k.bodyB_index = lb;
        k.contact_point.x = pt.x; k.contact_point.y = pt.y; k.contact_point.z = pt.z;
        k.contact_normal.x = nr.x; k.contact_normal.y = nr.y; k.contact_normal.z = nr.z;
        k.penetration = pt.w;
        out[c] = k;
    }
}

__global__ void __launch_bounds__(256) k_gather_words(const float* src, float* dst, const unsigned int* order, int n, int words)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n * words;
    for (; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int t = (int)(idx / words), f = (int)(idx - (long long)t * words);
        dst[idx] = src[(size_t)order[t] * words + f];
    }
}

// after COLLECT: the contact list in the reference's emission order; like the reference it is cut at `cap`
extern "C" int bbx_dev_download_contacts(BbxDev* d, CollisionContact* out, int cap, int* count)
{
    if (!d || !count) return BBX_ERR_BAD_ARGUMENT;
    *count = 0;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    int rc = ensure_export_buffers(d);
    if (rc) return rc;
    BBX_CUDA(d, cudaMemcpyAsync(d->ctr_host, d->ctr, sizeof(StepCounters), cudaMemcpyDeviceToHost, d->stream));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    { int rco = bbx_check_overflow(d); if (rco) return rco; }
    int n = d->ctr_host->n_contacts;
    if (n <= 0) return 0;
    KeyBits kb = key_bits_of(d);
    BBX_LAUNCH(d, k_export_contacts, BBX_NUM_SMS * 8, 256, 0, *d, (CollisionContact*)d->export_buf, d->sort_keys, kb);
    unsigned long long* ks; unsigned int* order;
    rc = bbx_radix_sort(d, d->sort_keys, d->sort_vals_a, d->sort_keys_b, d->sort_vals_b, n, kb.wb + 3 + kb.ib + kb.jb + 5,
                        d->sort_hist, d->sort_hist_scan, &ks, &order);
    if (rc) return rc;
    BBX_LAUNCH(d, k_gather_words, BBX_NUM_SMS * 8, 256, 0, (const float*)d->export_buf, (float*)d->export_sorted, order, n, 11);
    BBX_CUDA(d, cudaGetLastError());
    int take = n < cap ? n : cap;                      // :1988, :2020, :2052: the reference stops adding at capacity
    if (take > 0 && out)
        BBX_CUDA(d, cudaMemcpyAsync(out, d->export_sorted, sizeof(CollisionContact) * (size_t)take, cudaMemcpyDeviceToHost, d->stream));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    *count = take;
    return 0;
}

extern "C" int bbx_dev_download_forces(BbxDev* d, const BbxHostBodies* hb)
{
    if (!d || !hb) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    size_t NS = d->NS, NB = d->NB;
    if (NS) { BBX_CUDA(d, cudaMemcpyAsync(hb->s_force, d->r_s_force, 12 * NS, cudaMemcpyDeviceToHost, d->stream));
              BBX_CUDA(d, cudaMemcpyAsync(hb->s_torque, d->r_s_torque, 12 * NS, cudaMemcpyDeviceToHost, d->stream)); }
    if (NB) { BBX_CUDA(d, cudaMemcpyAsync(hb->b_force, d->r_b_force, 12 * NB, cudaMemcpyDeviceToHost, d->stream));
              BBX_CUDA(d, cudaMemcpyAsync(hb->b_torque, d->r_b_torque, 12 * NB, cudaMemcpyDeviceToHost, d->stream)); }
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    return 0;
}

extern "C" int bbx_dev_run_stage(BbxDev* d, const BbxStepParams* p, int stage)
{
    if (!d || !p) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    if (d->n_worlds != 1) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "stage functions work on single worlds", __FILE__, __LINE__);
    BBX_CUDA(d, cudaSetDevice(d->device));
    d->last_params = *p; d->have_params = 1;
    int rc = 0;
    switch (stage) {
    case BBX_STAGE_APPLY_FORCES: rc = bbx_stage_apply_forces_only(d, p); break;
    case BBX_STAGE_INTEGRATE_V:  rc = bbx_stage_integrate_velocities_only(d, p); break;
    case BBX_STAGE_COLLECT:
        d->imported_list = 0;
        if ((rc = bbx_stage_broadphase(d, p))) break;
        rc = bbx_stage_narrowphase(d, p);
        break;
    case BBX_STAGE_PRESOLVE:
        if (!d->imported_list) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "PRESOLVE needs an imported list", __FILE__, __LINE__);
        rc = bbx_stage_solve_ex(d, p, 1, 0, 1); break;
    case BBX_STAGE_SOLVE_SWEEP:
        if (!d->imported_list) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "SOLVE_SWEEP needs an imported list", __FILE__, __LINE__);
        rc = bbx_stage_solve_ex(d, p, 2, 1, 1); break;
    case BBX_STAGE_SOLVE_ALL:
        if (!d->imported_list) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "SOLVE_ALL needs an imported list", __FILE__, __LINE__);
        rc = bbx_stage_solve_ex(d, p, 0, p->settings.solver_iterations, 1); break;
    case BBX_STAGE_INTEGRATE_P:  rc = bbx_stage_positions_only(d, p); break;
    case BBX_STAGE_SLEEP:        rc = bbx_stage_sleep_only(d, p); break;
    default: return BBX_ERR_BAD_ARGUMENT;
    }
    return rc;
}

// debug (BBX_TRACE_LEVEL=l+1): per warp of the staged replay, level l of iteration 1: start ns, data-ready ns, end ns, class << 16 | chunks
extern "C" int bbx_dev_debug_warp_trace(BbxDev* d, unsigned long long* out, int n_warps)
{
    if (!d || !out || n_warps < 0 || n_warps > BBX_TRACE_WARPS) return BBX_ERR_BAD_ARGUMENT;
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    BBX_CUDA(d, cudaMemcpy(out, d->level_ns + 4104, sizeof(unsigned long long) * 4 * (size_t)n_warps, cudaMemcpyDeviceToHost));
    return 0;
}

// debug: per-level node counts (per class) and end-of-level timestamps of iteration 1
extern "C" int bbx_dev_debug_levels(BbxDev* d, int* counts, unsigned long long* ns, int max_levels)
{
    if (!d || !counts || !ns) return BBX_ERR_BAD_ARGUMENT;
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    int n = max_levels < 4096 ? max_levels : 4096;
    BBX_CUDA(d, cudaMemcpy(counts, d->level_count, sizeof(int) * (size_t)n * BBX_NCLS, cudaMemcpyDeviceToHost));
    BBX_CUDA(d, cudaMemcpy(ns, d->level_ns, sizeof(unsigned long long) * (size_t)(n + 1), cudaMemcpyDeviceToHost));   // [0..n]: replay, [2048..]: iteration 0 (when n = 4096)
    return 0;
}

// Debug / tests: the visiting order of the last level-path solve.  out[i] = (level, constraints of the node, gid A, gid B)
// for every node, level by level and class by class; returns the number of nodes (or a negative error).
extern "C" int bbx_dev_debug_schedule(BbxDev* d, int* out4, int cap_nodes)
{
    if (!d || !out4) return BBX_ERR_BAD_ARGUMENT;
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    StepCounters c;
    BBX_CUDA(d, cudaMemcpy(&c, d->ctr, sizeof(c), cudaMemcpyDeviceToHost));
    if (d->last_flow || d->last_per_contact || c.n_levels <= 0) return 0;
    const int L = c.n_levels;
    int* lc = (int*)malloc(sizeof(int) * (size_t)L * BBX_NCLS);
    if (!lc) return BBX_ERR_BAD_ARGUMENT;
    cudaError_t e = cudaMemcpy(lc, d->level_count, sizeof(int) * (size_t)L * BBX_NCLS, cudaMemcpyDeviceToHost);
    int n = 0;
    for (int k = 0; k < BBX_NCLS && e == cudaSuccess; k++) {
        long long tot = 0;
        for (int l = 0; l < L; l++) tot += lc[l * BBX_NCLS + k];
        if (tot == 0) continue;
        int4* nd = (int4*)malloc(sizeof(int4) * (size_t)tot);
        if (!nd) { free(lc); return BBX_ERR_BAD_ARGUMENT; }
        e = cudaMemcpy(nd, d->Nd_k[k], sizeof(int4) * (size_t)tot, cudaMemcpyDeviceToHost);
        long long at = 0;
        for (int l = 0; l < L && e == cudaSuccess; l++)
            for (int i = 0; i < lc[l * BBX_NCLS + k]; i++, at++) {
                if (n < cap_nodes) { out4[4 * n] = l; out4[4 * n + 1] = nd[at].y; out4[4 * n + 2] = nd[at].z; out4[4 * n + 3] = nd[at].w; }
                n++;
            }
        free(nd);
    }
    free(lc);
    BBX_CUDA(d, e);
    return n;
}

