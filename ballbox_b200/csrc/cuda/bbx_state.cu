// bbx_state.cu -- device context lifetime, host<->HBM transfer of the public SoA
// arrays (reference ballbox.h:62-95) and the pack/unpack kernels that convert
// between the user-visible 12-byte Vec3 layout and the float4 working layout.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bbx_dev.cuh"

int bbx_fail(BbxDev* d, int code, const char* what, const char* file, int line)
{
    if (d && !d->has_err && d->err) {
        snprintf(d->err, 256, "ballbox-b200: %s (code %d) at %s:%d", what, code, file, line);
        d->has_err = 1;
        fprintf(stderr, "%s\n", d->err);
    }
    return code ? code : BBX_ERR_STICKY;
}

extern "C" int bbx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void* bbx_host_alloc(size_t bytes, int* pinned)
{
    void* p = NULL;
    if (bytes == 0) bytes = 16;
    if (pinned) *pinned = 0;
    if (bbx_device_count() > 0 && bytes <= ((size_t)3 << 30)) {
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) {
            memset(p, 0, bytes);
            if (pinned) *pinned = 1;
            return p;
        }
        cudaGetLastError();
    }
    return calloc(1, bytes);
}

extern "C" void bbx_host_free(void* p, int pinned)
{
    if (!p) return;
    if (pinned) cudaFreeHost(p); else free(p);
}

template <typename T>
static int dalloc(BbxDev* d, T** p, size_t count)
{
    if (count == 0) count = 1;
    BBX_CUDA(d, cudaMalloc((void**)p, count * sizeof(T)));
    BBX_CUDA(d, cudaMemsetAsync(*p, 0, count * sizeof(T), d->stream));
    return 0;
}
#define DALLOC(ptr, count) do { int rc_ = dalloc(d, &(ptr), (size_t)(count)); if (rc_) return rc_; } while (0)

extern "C" int bbx_dev_create(BbxDev** out, int device, int n_worlds, int cap_s, int cap_b, int cap_contacts_world)
{
    if (!out || n_worlds < 1 || cap_s < 0 || cap_b < 0 || cap_contacts_world < 0) return BBX_ERR_BAD_ARGUMENT;
    *out = NULL;
    if (bbx_device_count() <= 0) {
        fprintf(stderr, "ballbox-b200: no CUDA device visible; this library has no CPU fallback\n");
        return BBX_ERR_NO_DEVICE;
    }
    BbxDev* d = (BbxDev*)calloc(1, sizeof(BbxDev));
    if (!d) return BBX_ERR_BAD_ARGUMENT;
    *out = d;                      // caller destroys on failure
    d->err = (char*)calloc(256, 1);
    d->device = device;
    BBX_CUDA(d, cudaSetDevice(device));
    BBX_CUDA(d, cudaDeviceGetAttribute(&d->num_sms, cudaDevAttrMultiProcessorCount, device));
    d->stream = 0;
    d->n_worlds = n_worlds; d->cap_s = cap_s; d->cap_b = cap_b;
    long long NS = (long long)n_worlds * cap_s, NB = (long long)n_worlds * cap_b;
    long long CC = (long long)n_worlds * cap_contacts_world;
    if (NS + NB > (1ll << 28) || CC > (1ll << 28)) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "capacity too large", __FILE__, __LINE__);
    if (NS >= (1 << 24) || NB >= (1 << 24)) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "more than 2^24 bodies of one type", __FILE__, __LINE__);
    d->NS = (int)NS; d->NB = (int)NB; d->N = (int)(NS + NB);
    d->cap_contacts_world = cap_contacts_world;
    d->cap_contacts = (int)(CC < 1024 ? 1024 : CC);
    long long cc2 = 2ll * d->cap_contacts;
    d->cap_cand = (int)(cc2 < 65536 ? 65536 : cc2);
    long long cells = 4ll * d->N;
    if (cells < 65536) cells = 65536;
    if (cells > (1ll << 25)) cells = 1ll << 25;
    if (cells < 64ll * n_worlds) cells = 64ll * n_worlds;
    d->cap_cells = (int)cells;
    int N = d->N;

    DALLOC(d->r_s_pos, 3 * NS); DALLOC(d->r_s_radius, NS); DALLOC(d->r_s_vel, 3 * NS); DALLOC(d->r_s_force, 3 * NS);
    DALLOC(d->r_s_mass, NS); DALLOC(d->r_s_angvel, 3 * NS); DALLOC(d->r_s_torque, 3 * NS); DALLOC(d->r_s_inertia, NS);
    DALLOC(d->r_s_timer, NS); DALLOC(d->r_s_static, NS); DALLOC(d->r_s_sleep, NS); DALLOC(d->r_s_hascb, NS);
    DALLOC(d->r_b_pos, 3 * NB); DALLOC(d->r_b_size, 3 * NB); DALLOC(d->r_b_rot, 4 * NB); DALLOC(d->r_b_vel, 3 * NB);
    DALLOC(d->r_b_force, 3 * NB); DALLOC(d->r_b_mass, NB); DALLOC(d->r_b_angvel, 3 * NB); DALLOC(d->r_b_torque, 3 * NB);
    DALLOC(d->r_b_inertia, 3 * NB); DALLOC(d->r_b_timer, NB); DALLOC(d->r_b_static, NB); DALLOC(d->r_b_sleep, NB);
    DALLOC(d->r_b_hascb, NB);
    DALLOC(d->d_s_count, n_worlds); DALLOC(d->d_b_count, n_worlds);

    DALLOC(d->pos, N); DALLOC(d->vel, 2 * (size_t)N); DALLOC(d->mass, N); DALLOC(d->timer, N); DALLOC(d->flags, N);
    DALLOC(d->boxq, 2 * NB); DALLOC(d->half, NB); DALLOC(d->brec, 2 * N);

    DALLOC(d->grid, 1);
    DALLOC(d->cell_count, d->cap_cells + 1); DALLOC(d->cell_start, d->cap_cells + 1); DALLOC(d->cell_cursor, d->cap_cells + 1);
    DALLOC(d->body_cell, N); DALLOC(d->sorted_id, N); DALLOC(d->sorted_pos, N);
    DALLOC(d->big_list, N); DALLOC(d->big_start, n_worlds + 1);
    DALLOC(d->scan_tmp, 1 << 16);

    DALLOC(d->cand_sb, d->cap_cand); DALLOC(d->cand_bb, d->cap_cand);
    DALLOC(d->hit_ids, d->cap_cand); DALLOC(d->hit_ax, d->cap_cand);
    int C = d->cap_contacts;
    DALLOC(d->c_ids, C); DALLOC(d->c_pt, C); DALLOC(d->c_nrm, C);
    DALLOC(d->L, 4 * (size_t)C); DALLOC(d->node, 2 * (size_t)C); DALLOC(d->bar, 8);
    DALLOC(d->deg, N + 1); DALLOC(d->adj_start, N + 1); DALLOC(d->adj_cursor, N + 1);
    DALLOC(d->adj, 2 * (size_t)C);
    DALLOC(d->indeg, C);
    {
        static const int min_len[BBX_NCLS] = {1, 1, 2, 3, 4, 5, 7, 9};     // a class-k node holds >= min_len[k] contacts
        for (int k = 0; k < BBX_NCLS; k++) { DALLOC(d->queue_k[k], C / min_len[k] + 64); DALLOC(d->Nd_k[k], C / min_len[k] + 64); }
    }
    d->cap_levels = C < (1 << 20) ? C : (1 << 20);
    DALLOC(d->slot_of, C); DALLOC(d->link, C); DALLOC(d->level_count, ((size_t)d->cap_levels + 4) * BBX_NCLS);
    DALLOC(d->M, 4 * (size_t)C);
    DALLOC(d->flow, 4 * (size_t)N); DALLOC(d->fq, C); DALLOC(d->fnd, 2 * (size_t)C); DALLOC(d->fimp, 2 * (size_t)C);
    BBX_CUDA(d, cudaMemsetAsync(d->fq, 0xff, sizeof(int) * (size_t)C, d->stream));
    // measurement switches (A/B runs of tools/ and profiles/r01_notes.md), read once per device state
    { const char* e = getenv("BBX_COLOUR_LIMIT"); d->colour_limit = e ? atoi(e) : 64; if (d->colour_limit < 1 || d->colour_limit > 64) d->colour_limit = 64; }
    { const char* e = getenv("BBX_FIND_PAIRS"); d->find_pairs_mode = e ? (e[0] == '0' ? 0 : 1) : -1; }
    { const char* e = getenv("BBX_GRP_WPB"); d->grp_force_wpb = e ? atoi(e) : 0; }
    { const char* e = getenv("BBX_GRP_PAIR_CAP"); d->grp_pair_cap = e ? atoi(e) : 0; }
    { const char* e = getenv("BBX_GRP_SMEM"); d->grp_smem_wpb = e ? atoi(e) : 2; }
    { const char* e = getenv("BBX_SOLVER_FLOW"); d->force_flow = (e && e[0] == '1') ? 1 : 0; }
    { const char* e = getenv("BBX_REPLAY"); d->replay_mode = (e && e[0] == '0') ? 0 : ((e && e[0] == '2') ? 2 : 1); }   // 1 (default): staged replay kernel unless the schedule is thin; 2: always; 0: never
    { const char* e = getenv("BBX_SOLVE_BLOCKS"); d->tune_solve_blocks = e ? atoi(e) : 0; }
    { const char* e = getenv("BBX_FLOW_BLOCKS"); d->tune_flow_blocks = e ? atoi(e) : 0; }
    { const char* e = getenv("BBX_FLOW_GATHER"); d->tune_flow_gather = e ? atoi(e) : 0; }
    { const char* e = getenv("BBX_TRACE_LEVEL"); d->trace_level = e ? atoi(e) : 0; }
    DALLOC(d->level_ns, 4096 + 8 + 4 * BBX_TRACE_WARPS);      // level times, debug counters, per-warp trace of one level
    DALLOC(d->ctr, 1);
    DALLOC(d->sort_keys, C);
    d->export_buf = NULL;          // allocated on first constraint download
    BBX_CUDA(d, cudaHostAlloc((void**)&d->ctr_host, sizeof(StepCounters), cudaHostAllocDefault));
    memset(d->ctr_host, 0, sizeof(StepCounters));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    return 0;
}

extern "C" void bbx_dev_destroy(BbxDev* d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();
    void* ptrs[] = { d->r_s_pos, d->r_s_radius, d->r_s_vel, d->r_s_force, d->r_s_mass, d->r_s_angvel, d->r_s_torque,
        d->r_s_inertia, d->r_s_timer, d->r_s_static, d->r_s_sleep, d->r_s_hascb, d->r_b_pos, d->r_b_size, d->r_b_rot,
        d->r_b_vel, d->r_b_force, d->r_b_mass, d->r_b_angvel, d->r_b_torque, d->r_b_inertia, d->r_b_timer, d->r_b_static,
        d->r_b_sleep, d->r_b_hascb, d->d_s_count, d->d_b_count, d->pos, d->vel, d->mass, d->timer, d->flags,
        d->boxq, d->half, d->brec, d->grid, d->cell_count, d->cell_start, d->cell_cursor, d->body_cell, d->sorted_id,
        d->sorted_pos, d->big_list, d->big_start, d->scan_tmp, d->cand_sb, d->cand_bb, d->hit_ids, d->hit_ax, d->c_ids, d->c_pt, d->c_nrm,
        d->L, d->node, d->bar, d->deg, d->adj_start, d->adj_cursor, d->adj, d->indeg,
        d->sdf_grid, d->slot_of, d->link, d->level_count, d->flow, d->fq, d->fnd, d->fimp, d->level_ns, d->M, d->ctr, d->snap_pos, d->snap_vel, d->snap_boxq, d->snap_timer, d->snap_flags, d->reset_ids, d->wake_ids, d->ev_key, d->ev_step, d->ev_contact, d->sort_keys, d->export_buf, d->export_sorted, d->sort_keys_b, d->sort_vals_a, d->sort_vals_b,
        d->sort_hist, d->sort_hist_scan, d->warm_key, d->warm_val, d->warm_epoch, d->M2, d->level_count2, d->col_used, d->col_tab };
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) if (ptrs[i]) cudaFree(ptrs[i]);
    bbx_joints_free(d);
    for (int k = 0; k < BBX_NCLS; k++) { if (d->queue_k[k]) cudaFree(d->queue_k[k]); if (d->Nd_k[k]) cudaFree(d->Nd_k[k]); }
    for (int k = 0; k < BBX_NCLS; k++) { if (d->queue2_k[k]) cudaFree(d->queue2_k[k]); if (d->Nd2_k[k]) cudaFree(d->Nd2_k[k]); if (d->colq_k[k]) cudaFree(d->colq_k[k]); }
    if (d->ctr_host) cudaFreeHost(d->ctr_host);
    if (d->timing_ev) { for (int i = 0; i < d->timing_cap * BBX_TIMING_MARKS; i++) cudaEventDestroy(d->timing_ev[i]); free(d->timing_ev); free(d->timing_has7); }
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    if (d->ev_upload) cudaEventDestroy(d->ev_upload);
    if (d->ev_chain) cudaEventDestroy(d->ev_chain);
    cudaGetLastError();
    free(d->err);
    free(d);
}

extern "C" int bbx_dev_set_baked_sdf(BbxDev* d, const float* values, int channels, int nx, int ny, int nz, const float origin[3], float cell)
{
    if (!d || !values || !origin || nx < 2 || ny < 2 || nz < 2 || !(cell > 0.0f) || (channels != 1 && channels != 4)) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));           // no step may still be reading the old grid
    if (d->sdf_grid) { cudaFree(d->sdf_grid); d->sdf_grid = NULL; }
    size_t n = (size_t)nx * ny * nz * (size_t)channels;
    BBX_CUDA(d, cudaMalloc((void**)&d->sdf_grid, n * sizeof(float)));
    BBX_CUDA(d, cudaMemcpyAsync(d->sdf_grid, values, n * sizeof(float), cudaMemcpyHostToDevice, d->stream));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));           // `values` may be freed by the caller
    d->sdf_nx = nx; d->sdf_ny = ny; d->sdf_nz = nz; d->sdf_ch = channels;
    d->sdf_ox = origin[0]; d->sdf_oy = origin[1]; d->sdf_oz = origin[2]; d->sdf_inv_cell = 1.0f / cell;
    return 0;
}

extern "C" void bbx_dev_set_stream(BbxDev* d, void* s) { if (d) d->stream = (cudaStream_t)s; }
extern "C" const char* bbx_dev_error(BbxDev* d) { return (d && d->has_err) ? d->err : NULL; }
extern "C" void bbx_dev_clear_error(BbxDev* d) { if (d) { d->has_err = 0; if (d->err) d->err[0] = 0; } }

extern "C" int bbx_dev_synchronize(BbxDev* d)
{
    if (!d) return BBX_ERR_BAD_ARGUMENT;
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// pack: raw host-layout mirrors -> float4 working state
// ---------------------------------------------------------------------------
__global__ void k_pack(BbxDev dv)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= dv.N) return;
    if (g < dv.NS) {
        int w = g / dv.cap_s, i = g - w * dv.cap_s;
        bool valid = i < dv.d_s_count[w];
        if (!valid) { dv.flags[g] = 0; return; }
        bool st = dv.r_s_static[g] != 0;
        float r = dv.r_s_radius[g], m = dv.r_s_mass[g];
        dv.pos[g] = make_float4(dv.r_s_pos[3 * g], dv.r_s_pos[3 * g + 1], dv.r_s_pos[3 * g + 2], r);
        dv.vel[2 * g] = make_float4(dv.r_s_vel[3 * g], dv.r_s_vel[3 * g + 1], dv.r_s_vel[3 * g + 2], st ? 0.0f : 1.0f / m);
        dv.vel[2 * g + 1] = make_float4(dv.r_s_angvel[3 * g], dv.r_s_angvel[3 * g + 1], dv.r_s_angvel[3 * g + 2],
                                st ? 0.0f : 1.0f / dv.r_s_inertia[g]);
        dv.mass[g] = m;
        dv.timer[g] = dv.r_s_timer[g];
        dv.flags[g] = (unsigned char)(BF_VALID | (st ? BF_STATIC : 0) | (dv.r_s_sleep[g] ? BF_SLEEP : 0) |
                                      (dv.r_s_hascb[g] ? BF_HASCB : 0));
        unsigned int rb = __float_as_uint(fmaxf(r, 0.0f));
        if (!st) atomicMax(&dv.ctr->cut_dyn, rb);
        atomicMax(&dv.ctr->cut_all, rb);
    } else {
        int b = g - dv.NS;
        int w = b / dv.cap_b, i = b - w * dv.cap_b;
        bool valid = i < dv.d_b_count[w];
        if (!valid) { dv.flags[g] = 0; return; }
        bool st = dv.r_b_static[b] != 0;
        float m = dv.r_b_mass[b];
        float3 h = f3(dv.r_b_size[3 * b], dv.r_b_size[3 * b + 1], dv.r_b_size[3 * b + 2]);
        float4 q = make_float4(dv.r_b_rot[4 * b], dv.r_b_rot[4 * b + 1], dv.r_b_rot[4 * b + 2], dv.r_b_rot[4 * b + 3]);
        float R = box_bound_radius(h, q);
        dv.pos[g] = make_float4(dv.r_b_pos[3 * b], dv.r_b_pos[3 * b + 1], dv.r_b_pos[3 * b + 2], R);
        dv.vel[2 * g] = make_float4(dv.r_b_vel[3 * b], dv.r_b_vel[3 * b + 1], dv.r_b_vel[3 * b + 2], st ? 0.0f : 1.0f / m);
        dv.vel[2 * g + 1] = make_float4(dv.r_b_angvel[3 * b], dv.r_b_angvel[3 * b + 1], dv.r_b_angvel[3 * b + 2], 0.0f);
        dv.boxq[2 * b] = q;
        dv.half[b] = make_float4(h.x, h.y, h.z, 0.0f);
        dv.boxq[2 * b + 1] = st ? make_float4(0, 0, 0, 0)
                        : make_float4(1.0f / dv.r_b_inertia[3 * b], 1.0f / dv.r_b_inertia[3 * b + 1],
                                      1.0f / dv.r_b_inertia[3 * b + 2], 0.0f);
        dv.mass[g] = m;
        dv.timer[g] = dv.r_b_timer[b];
        dv.flags[g] = (unsigned char)(BF_VALID | (st ? BF_STATIC : 0) | (dv.r_b_sleep[b] ? BF_SLEEP : 0) |
                                      (dv.r_b_hascb[b] ? BF_HASCB : 0));
        unsigned int rb = __float_as_uint(R);
        if (!st) atomicMax(&dv.ctr->cut_dyn, rb);
        atomicMax(&dv.ctr->cut_all, rb);
    }
}

// bodies larger than the largest non-static body bypass the grid
__global__ void k_choose_cut(BbxDev dv)
{
    float cd = __uint_as_float(dv.ctr->cut_dyn), ca = __uint_as_float(dv.ctr->cut_all);
    float cut = cd > 0.0f ? cd : ca;
    if (!(cut > 0.0f)) cut = 1.0f;
    dv.grid->small_cut = cut;
}

__global__ void k_big_count(BbxDev dv, int* big_cnt)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= dv.N) return;
    unsigned char f = dv.flags[g];
    if (!(f & BF_VALID)) return;
    if ((f & BF_STATIC) && dv.pos[g].w > dv.grid->small_cut) {
        dv.flags[g] = f | BF_BIG;
        int w = (g < dv.NS) ? g / dv.cap_s : (g - dv.NS) / dv.cap_b;
        atomicAdd(&big_cnt[w], 1);
        atomicAdd(&dv.ctr->n_big, 1);
    }
}

__global__ void k_big_fill(BbxDev dv, int* cursor)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= dv.N) return;
    if (!(dv.flags[g] & BF_BIG)) return;
    int w = (g < dv.NS) ? g / dv.cap_s : (g - dv.NS) / dv.cap_b;
    int slot = atomicAdd(&cursor[w], 1);
    dv.big_list[slot] = g;
}

__global__ void k_reset_upload_counters(BbxDev dv)
{
    dv.ctr->cut_dyn = 0; dv.ctr->cut_all = 0; dv.ctr->n_big = 0;
}

int bbx_stage_pack(BbxDev* d)
{
    int nb = bbx_blocks(d->N, 256);
    BBX_LAUNCH(d, k_reset_upload_counters, 1, 1, 0, *d);
    BBX_LAUNCH(d, k_pack, nb, 256, 0, *d);
    BBX_LAUNCH(d, k_choose_cut, 1, 1, 0, *d);
    // big-body lists per world: count -> scan -> fill (cell_count/cell_cursor reused as scratch)
    BBX_CUDA(d, cudaMemsetAsync(d->cell_count, 0, sizeof(int) * (d->n_worlds + 1), d->stream));
    BBX_LAUNCH(d, k_big_count, nb, 256, 0, *d, d->cell_count);
    int rc = bbx_scan_exclusive(d, d->cell_count, d->big_start, d->cell_cursor, d->n_worlds + 1, NULL, NULL);
    if (rc) return rc;
    BBX_LAUNCH(d, k_big_fill, nb, 256, 0, *d, d->cell_cursor);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

extern "C" int bbx_dev_upload(BbxDev* d, const BbxHostBodies* hb)
{
    if (!d || !hb) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    if (hb->n_worlds != d->n_worlds || hb->cap_s != d->cap_s || hb->cap_b != d->cap_b)
        return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "upload: layout mismatch", __FILE__, __LINE__);
    BBX_CUDA(d, cudaSetDevice(d->device));
    size_t NS = d->NS, NB = d->NB;
    cudaStream_t s = d->stream;
#define UP(dst, src, bytes) do { if ((bytes) > 0 && (src)) BBX_CUDA(d, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, s)); } while (0)
    UP(d->d_s_count, hb->s_count, sizeof(int) * d->n_worlds);
    UP(d->d_b_count, hb->b_count, sizeof(int) * d->n_worlds);
    UP(d->r_s_pos, hb->s_pos, 12 * NS); UP(d->r_s_radius, hb->s_radius, 4 * NS); UP(d->r_s_vel, hb->s_vel, 12 * NS);
    UP(d->r_s_force, hb->s_force, 12 * NS); UP(d->r_s_mass, hb->s_mass, 4 * NS); UP(d->r_s_angvel, hb->s_angvel, 12 * NS);
    UP(d->r_s_torque, hb->s_torque, 12 * NS); UP(d->r_s_inertia, hb->s_inertia, 4 * NS); UP(d->r_s_timer, hb->s_timer, 4 * NS);
    UP(d->r_s_static, hb->s_static, NS); UP(d->r_s_sleep, hb->s_sleep, NS);
    if (hb->s_hascb) UP(d->r_s_hascb, hb->s_hascb, NS); else BBX_CUDA(d, cudaMemsetAsync(d->r_s_hascb, 0, NS ? NS : 1, s));
    UP(d->r_b_pos, hb->b_pos, 12 * NB); UP(d->r_b_size, hb->b_size, 12 * NB); UP(d->r_b_rot, hb->b_rot, 16 * NB);
    UP(d->r_b_vel, hb->b_vel, 12 * NB); UP(d->r_b_force, hb->b_force, 12 * NB); UP(d->r_b_mass, hb->b_mass, 4 * NB);
    UP(d->r_b_angvel, hb->b_angvel, 12 * NB); UP(d->r_b_torque, hb->b_torque, 12 * NB); UP(d->r_b_inertia, hb->b_inertia, 12 * NB);
    UP(d->r_b_timer, hb->b_timer, 4 * NB); UP(d->r_b_static, hb->b_static, NB); UP(d->r_b_sleep, hb->b_sleep, NB);
    if (hb->b_hascb) UP(d->r_b_hascb, hb->b_hascb, NB); else BBX_CUDA(d, cudaMemsetAsync(d->r_b_hascb, 0, NB ? NB : 1, s));
#undef UP
    if (!d->ev_upload) BBX_CUDA(d, cudaEventCreateWithFlags(&d->ev_upload, cudaEventDisableTiming));
    BBX_CUDA(d, cudaEventRecord(d->ev_upload, s));
    long long vs = 0, vb = 0;
    for (int w = 0; w < d->n_worlds; w++) { vs += hb->s_count[w]; vb += hb->b_count[w]; }
    d->n_valid_spheres = (int)vs; d->n_valid_boxes = (int)vb;
    return bbx_stage_pack(d);
}

extern "C" int bbx_dev_upload_forces(BbxDev* d, const BbxHostBodies* hb)
{
    if (!d || !hb) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    size_t NS = d->NS, NB = d->NB;
    if (NS) {
        BBX_CUDA(d, cudaMemcpyAsync(d->r_s_force, hb->s_force, 12 * NS, cudaMemcpyHostToDevice, d->stream));
        BBX_CUDA(d, cudaMemcpyAsync(d->r_s_torque, hb->s_torque, 12 * NS, cudaMemcpyHostToDevice, d->stream));
    }
    if (NB) {
        BBX_CUDA(d, cudaMemcpyAsync(d->r_b_force, hb->b_force, 12 * NB, cudaMemcpyHostToDevice, d->stream));
        BBX_CUDA(d, cudaMemcpyAsync(d->r_b_torque, hb->b_torque, 12 * NB, cudaMemcpyHostToDevice, d->stream));
    }
    // the host arrays are pinned: the copies run later, so the caller must bbx_dev_wait_upload() before it clears them
    if (!d->ev_upload) BBX_CUDA(d, cudaEventCreateWithFlags(&d->ev_upload, cudaEventDisableTiming));
    BBX_CUDA(d, cudaEventRecord(d->ev_upload, d->stream));
    return 0;
}

__global__ void k_set_cb_flags(BbxDev dv)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= dv.N) return;
    unsigned char f = dv.flags[g];
    if (!(f & BF_VALID)) return;
    bool cb = (g < dv.NS) ? dv.r_s_hascb[g] != 0 : dv.r_b_hascb[g - dv.NS] != 0;
    dv.flags[g] = cb ? (f | BF_HASCB) : (f & (unsigned char)~BF_HASCB);
}

extern "C" int bbx_dev_upload_callback_flags(BbxDev* d, const BbxHostBodies* hb)
{
    if (!d || !hb) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    if (d->NS && hb->s_hascb) BBX_CUDA(d, cudaMemcpyAsync(d->r_s_hascb, hb->s_hascb, d->NS, cudaMemcpyHostToDevice, d->stream));
    if (d->NB && hb->b_hascb) BBX_CUDA(d, cudaMemcpyAsync(d->r_b_hascb, hb->b_hascb, d->NB, cudaMemcpyHostToDevice, d->stream));
    BBX_LAUNCH(d, k_set_cb_flags, bbx_blocks(d->N, 256), 256, 0, *d);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// unpack: working state -> raw mirrors (only what a step can change)
// ---------------------------------------------------------------------------
__global__ void k_unpack(BbxDev dv)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= dv.N) return;
    unsigned char f = dv.flags[g];
    if (!(f & BF_VALID)) return;
    float4 p = dv.pos[g], v = dv.vel[2 * g], w = dv.vel[2 * g + 1];
    if (g < dv.NS) {
        dv.r_s_pos[3 * g] = p.x; dv.r_s_pos[3 * g + 1] = p.y; dv.r_s_pos[3 * g + 2] = p.z;
        dv.r_s_vel[3 * g] = v.x; dv.r_s_vel[3 * g + 1] = v.y; dv.r_s_vel[3 * g + 2] = v.z;
        dv.r_s_angvel[3 * g] = w.x; dv.r_s_angvel[3 * g + 1] = w.y; dv.r_s_angvel[3 * g + 2] = w.z;
        dv.r_s_sleep[g] = (f & BF_SLEEP) ? 1 : 0;
        dv.r_s_timer[g] = dv.timer[g];
    } else {
        int b = g - dv.NS;
        float4 q = dv.boxq[2 * b];
        dv.r_b_pos[3 * b] = p.x; dv.r_b_pos[3 * b + 1] = p.y; dv.r_b_pos[3 * b + 2] = p.z;
        dv.r_b_vel[3 * b] = v.x; dv.r_b_vel[3 * b + 1] = v.y; dv.r_b_vel[3 * b + 2] = v.z;
        dv.r_b_angvel[3 * b] = w.x; dv.r_b_angvel[3 * b + 1] = w.y; dv.r_b_angvel[3 * b + 2] = w.z;
        dv.r_b_rot[4 * b] = q.x; dv.r_b_rot[4 * b + 1] = q.y; dv.r_b_rot[4 * b + 2] = q.z; dv.r_b_rot[4 * b + 3] = q.w;
        dv.r_b_sleep[b] = (f & BF_SLEEP) ? 1 : 0;
        dv.r_b_timer[b] = dv.timer[g];
    }
}

int bbx_stage_unpack(BbxDev* d)
{
    BBX_LAUNCH(d, k_unpack, bbx_blocks(d->N, 256), 256, 0, *d);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

int bbx_copy_stream(BbxDev* d)
{
    if (!d->copy_stream) BBX_CUDA(d, cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    if (!d->ev_chain) BBX_CUDA(d, cudaEventCreateWithFlags(&d->ev_chain, cudaEventDisableTiming));
    return 0;
}

extern "C" int bbx_dev_wait_upload(BbxDev* d)
{
    if (!d) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    if (d->ev_upload) BBX_CUDA(d, cudaEventSynchronize(d->ev_upload));
    return 0;
}

// body arrays -> host on stream `s` (after the unpack kernel on the step stream)
static int download_bodies_on(BbxDev* d, const BbxHostBodies* hb, cudaStream_t s)
{
    size_t NS = d->NS, NB = d->NB;
#define DN(dst, src, bytes) do { if ((bytes) > 0 && (dst)) BBX_CUDA(d, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, s)); } while (0)
    DN(hb->s_pos, d->r_s_pos, 12 * NS); DN(hb->s_vel, d->r_s_vel, 12 * NS); DN(hb->s_angvel, d->r_s_angvel, 12 * NS);
    DN(hb->s_sleep, d->r_s_sleep, NS); DN(hb->s_timer, d->r_s_timer, 4 * NS);
    DN(hb->b_pos, d->r_b_pos, 12 * NB); DN(hb->b_vel, d->r_b_vel, 12 * NB); DN(hb->b_angvel, d->r_b_angvel, 12 * NB);
    DN(hb->b_rot, d->r_b_rot, 16 * NB); DN(hb->b_sleep, d->r_b_sleep, NB); DN(hb->b_timer, d->r_b_timer, 4 * NB);
#undef DN
    return 0;
}

extern "C" int bbx_dev_download(BbxDev* d, const BbxHostBodies* hb)
{
    if (!d || !hb) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    int rc = bbx_stage_unpack(d);
    if (rc) return rc;
    cudaStream_t s = d->stream;
    if ((rc = download_bodies_on(d, hb, s))) return rc;
    BBX_CUDA(d, cudaMemcpyAsync(d->ctr_host, d->ctr, sizeof(StepCounters), cudaMemcpyDeviceToHost, s));
    BBX_CUDA(d, cudaStreamSynchronize(s));
    return bbx_check_overflow(d);                // state computed from a truncated contact list is not handed out silently
}

extern "C" int bbx_dev_download_begin(BbxDev* d, const BbxHostBodies* hb)
{
    if (!d || !hb) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    int rc = bbx_stage_unpack(d);
    if (rc) return rc;
    if ((rc = bbx_copy_stream(d))) return rc;
    BBX_CUDA(d, cudaEventRecord(d->ev_chain, d->stream));
    BBX_CUDA(d, cudaStreamWaitEvent(d->copy_stream, d->ev_chain, 0));
    return download_bodies_on(d, hb, d->copy_stream);
}

extern "C" int bbx_dev_download_end(BbxDev* d)
{
    if (!d) return BBX_ERR_BAD_ARGUMENT;
    cudaSetDevice(d->device);
    if (d->copy_stream) cudaStreamSynchronize(d->copy_stream);     // also after a failure: the host arrays must be quiet
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaMemcpyAsync(d->ctr_host, d->ctr, sizeof(StepCounters), cudaMemcpyDeviceToHost, d->stream));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    return bbx_check_overflow(d);
}

extern "C" int bbx_dev_stats(BbxDev* d, BallboxStats* out)
{
    if (!d || !out) return BBX_ERR_BAD_ARGUMENT;
    memset(out, 0, sizeof(*out));
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_CUDA(d, cudaMemcpyAsync(d->ctr_host, d->ctr, sizeof(StepCounters), cudaMemcpyDeviceToHost, d->stream));
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));
    GridParams gp;
    BBX_CUDA(d, cudaMemcpy(&gp, d->grid, sizeof(gp), cudaMemcpyDeviceToHost));
    const StepCounters* c = d->ctr_host;
    out->steps = d->steps; out->body_steps = d->body_steps; out->kernel_launches = d->launches;
    out->spheres = d->n_valid_spheres; out->boxes = d->n_valid_boxes;
    out->candidate_pairs = c->n_ss_tests + c->n_cand_sb + c->n_cand_bb;
    out->contacts = c->n_contacts; out->solver_contacts = c->n_solver; out->touched_bodies = c->n_touched;
    out->levels = c->n_levels; out->overflow = c->overflow | (gp.overflow ? 4 : 0); out->grid_cells = gp.n_cells;
    out->reserved[0] = c->n_cand_sb; out->reserved[1] = c->n_cand_bb; out->reserved[2] = c->n_ss_tests;
    out->reserved[3] = c->n_big; out->reserved[4] = c->n_events;
    d->last_deferred = c->n_deferred;
    return 0;
}

// debug: heavy terminal nodes the level kernel held back for its last level in the step bbx_dev_stats last looked at
extern "C" int bbx_dev_debug_deferred(BbxDev* d) { return d ? d->last_deferred : 0; }

// ---------------------------------------------------------------------------
// zero-copy views, device-side wrench application, per-world reset
// ---------------------------------------------------------------------------
extern "C" int bbx_dev_views(BbxDev* d, BbxDeviceViews* out)
{
    if (!d || !out) return BBX_ERR_BAD_ARGUMENT;
    out->n_worlds = d->n_worlds; out->cap_spheres = d->cap_s; out->cap_boxes = d->cap_b;
    out->pos = d->pos; out->vel = d->vel; out->box_rot = d->boxq; out->flags = d->flags;
    out->sphere_force = d->r_s_force; out->sphere_torque = d->r_s_torque;
    out->box_force = d->r_b_force; out->box_torque = d->r_b_torque;
    return 0;
}

__global__ void __launch_bounds__(256) k_add_wrench(BbxDev dv, const float* sf, const float* st, const float* bf, const float* bt)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= dv.N) return;
    unsigned char f = dv.flags[g];
    if (!(f & BF_VALID)) return;
    bool box = g >= dv.NS;
    int i = box ? g - dv.NS : g;
    const float* fi = box ? bf : sf; const float* ti = box ? bt : st;
    float* fa = box ? dv.r_b_force : dv.r_s_force; float* ta = box ? dv.r_b_torque : dv.r_s_torque;
    bool nz = false;
    if (fi) { float x = fi[3 * i], y = fi[3 * i + 1], z = fi[3 * i + 2];
              if (x != 0.0f || y != 0.0f || z != 0.0f) { nz = true; fa[3 * i] += x; fa[3 * i + 1] += y; fa[3 * i + 2] += z; } }
    if (ti) { float x = ti[3 * i], y = ti[3 * i + 1], z = ti[3 * i + 2];
              if (x != 0.0f || y != 0.0f || z != 0.0f) { nz = true; ta[3 * i] += x; ta[3 * i + 1] += y; ta[3 * i + 2] += z; } }
    if (nz && (f & BF_SLEEP)) { dv.flags[g] = f & (unsigned char)~BF_SLEEP; dv.timer[g] = 0.0f; }
}

extern "C" int bbx_dev_add_wrench(BbxDev* d, const float* sf, const float* st, const float* bf, const float* bt)
{
    if (!d) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_LAUNCH(d, k_add_wrench, bbx_blocks(d->N, 256), 256, 0, *d, sf, st, bf, bt);
    BBX_CUDA(d, cudaGetLastError());
    d->forces_on_device = 1;
    return 0;
}

// Add*Force/Torque on a resident world (ballbox.h:420-423 & co.): the listed body slots wake up, whatever the force was
__global__ void __launch_bounds__(256) k_wake_list(BbxDev dv, const int* ids, int n)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int g = ids[t];
    if (g < 0 || g >= dv.N) return;
    const unsigned char f = dv.flags[g];
    if ((f & BF_VALID) && (f & BF_SLEEP)) { dv.flags[g] = f & (unsigned char)~BF_SLEEP; dv.timer[g] = 0.0f; }      // duplicates write the same values
}

extern "C" int bbx_dev_wake_bodies(BbxDev* d, const int* slots, int n)
{
    if (!d || (n > 0 && !slots)) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    if (n <= 0) return 0;
    BBX_CUDA(d, cudaSetDevice(d->device));
    if (n > d->wake_cap) {
        if (d->wake_ids) BBX_CUDA(d, cudaFree(d->wake_ids));
        d->wake_ids = NULL; d->wake_cap = 0;
        int cap = 1024;
        while (cap < n) cap *= 2;
        BBX_CUDA(d, cudaMalloc((void**)&d->wake_ids, sizeof(int) * (size_t)cap));
        d->wake_cap = cap;
    }
    BBX_CUDA(d, cudaMemcpyAsync(d->wake_ids, slots, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, d->stream));   // pageable source: staged before the call returns
    BBX_LAUNCH(d, k_wake_list, bbx_blocks(n, 256), 256, 0, *d, d->wake_ids, n);
    BBX_CUDA(d, cudaGetLastError());
    return 0;
}

extern "C" int bbx_dev_snapshot(BbxDev* d)
{
    if (!d) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    BBX_CUDA(d, cudaSetDevice(d->device));
    size_t N = d->N, NB = d->NB;
    if (!d->snap_pos) {
        BBX_CUDA(d, cudaMalloc((void**)&d->snap_pos, sizeof(float4) * (N ? N : 1)));
        BBX_CUDA(d, cudaMalloc((void**)&d->snap_vel, sizeof(float4) * 2 * (N ? N : 1)));
        BBX_CUDA(d, cudaMalloc((void**)&d->snap_boxq, sizeof(float4) * 2 * (NB ? NB : 1)));
        BBX_CUDA(d, cudaMalloc((void**)&d->snap_timer, sizeof(float) * (N ? N : 1)));
        BBX_CUDA(d, cudaMalloc((void**)&d->snap_flags, N ? N : 1));
        BBX_CUDA(d, cudaMalloc((void**)&d->reset_ids, sizeof(int) * (size_t)d->n_worlds));
    }
    cudaStream_t s = d->stream;
    BBX_CUDA(d, cudaMemcpyAsync(d->snap_pos, d->pos, sizeof(float4) * N, cudaMemcpyDeviceToDevice, s));
    BBX_CUDA(d, cudaMemcpyAsync(d->snap_vel, d->vel, sizeof(float4) * 2 * N, cudaMemcpyDeviceToDevice, s));
    BBX_CUDA(d, cudaMemcpyAsync(d->snap_boxq, d->boxq, sizeof(float4) * 2 * NB, cudaMemcpyDeviceToDevice, s));
    BBX_CUDA(d, cudaMemcpyAsync(d->snap_timer, d->timer, sizeof(float) * N, cudaMemcpyDeviceToDevice, s));
    BBX_CUDA(d, cudaMemcpyAsync(d->snap_flags, d->flags, N, cudaMemcpyDeviceToDevice, s));
    return 0;
}

// one block per (world in the list); threads stride over that world's sphere and box slots
__global__ void __launch_bounds__(256) k_restore_worlds(BbxDev dv, int n)
{
    int w = dv.reset_ids[blockIdx.x];
    if (blockIdx.x >= n || w < 0 || w >= dv.n_worlds) return;
    if (threadIdx.x == 0 && dv.warm_epoch) dv.warm_epoch[w]++;          // the world's remembered impulses are stale now
    for (int i = threadIdx.x; i < dv.cap_s + dv.cap_b; i += blockDim.x) {
        bool box = i >= dv.cap_s;
        int g = box ? dv.NS + w * dv.cap_b + (i - dv.cap_s) : w * dv.cap_s + i;
        dv.pos[g] = dv.snap_pos[g];
        dv.vel[2 * g] = dv.snap_vel[2 * g]; dv.vel[2 * g + 1] = dv.snap_vel[2 * g + 1];
        dv.timer[g] = dv.snap_timer[g];
        dv.flags[g] = dv.snap_flags[g];
        if (box) {
            int b = g - dv.NS;
            dv.boxq[2 * b] = dv.snap_boxq[2 * b]; dv.boxq[2 * b + 1] = dv.snap_boxq[2 * b + 1];
            for (int k = 0; k < 3; k++) { dv.r_b_force[3 * b + k] = 0.0f; dv.r_b_torque[3 * b + k] = 0.0f; }
        } else {
            for (int k = 0; k < 3; k++) { dv.r_s_force[3 * g + k] = 0.0f; dv.r_s_torque[3 * g + k] = 0.0f; }
        }
    }
}

extern "C" int bbx_dev_restore_worlds(BbxDev* d, const int* world_ids, int n)
{
    if (!d || (n > 0 && !world_ids)) return BBX_ERR_BAD_ARGUMENT;
    if (d->has_err) return BBX_ERR_STICKY;
    if (!d->snap_pos) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "restore without a snapshot", __FILE__, __LINE__);
    if (n <= 0) return 0;
    if (n > d->n_worlds) return bbx_fail(d, BBX_ERR_BAD_ARGUMENT, "restore: more ids than worlds", __FILE__, __LINE__);
    BBX_CUDA(d, cudaSetDevice(d->device));
    BBX_CUDA(d, cudaMemcpyAsync(d->reset_ids, world_ids, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, d->stream));
    BBX_LAUNCH(d, k_restore_worlds, n, 256, 0, *d, n);
    BBX_CUDA(d, cudaGetLastError());
    BBX_CUDA(d, cudaStreamSynchronize(d->stream));      // world_ids is caller memory
    return 0;
}
