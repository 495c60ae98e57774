// bbx_dev.cuh -- device-side data layout and exact-arithmetic helpers shared by
// the sm_100a kernels.
//
// ARITHMETIC CONTRACT.  Every kernel that produces physics results is compiled
// with -fmad=false and uses only IEEE binary32 + - * / sqrtf in the operation
// order of the reference (ballbox.h:501-639), so that results are bit-identical
// to the reference built with gcc -ffp-contract=off.  Do not
// "simplify" expressions in this directory: x + 0.0f, v * -1.0f and the
// association of sums are deliberate.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include "bbx_cuda.h"
#include "bbx_sdf_library.h"

// ---------------------------------------------------------------------------
// body flags (one byte per body slot)
// ---------------------------------------------------------------------------
#define BF_STATIC   1u
#define BF_SLEEP    2u
#define BF_VALID    4u
#define BF_HASCB    8u
#define BF_BIG     16u     // static body too large for the grid: tested against every body of its world

// contact phases, in the reference's emission order (ballbox.h:1985-2114)
#define PH_SS   0
#define PH_BB   1
#define PH_SB   2
#define PH_SSDF 3
#define PH_BSDF 4

// launch grids are sized from the device's SM count (BbxDev::num_sms, cudaDevAttrMultiProcessorCount); every use
// sits in a host function whose BbxDev* is called d
#define BBX_NUM_SMS (d->num_sms)
#define BBX_TIMING_MARKS 8       // step start, after K1, broadphase, narrowphase, presolve+graph, solve kernels, K12; [7] = between k_solve_levels and k_replay_staged
#define BBX_TIMING_STEPS 256
#define BBX_TRACE_WARPS 8192       // per-warp trace slots behind level_ns[4104] (BBX_TRACE_LEVEL)
#define BBX_NCLS 8               // solver node classes by run length: 1 (no box), 1 (box), 2, 3, 4, 5-6, 7-8, 9+

struct GridParams {
    float ox, oy, oz;        // origin of cell (0,0,0)
    float cell, inv_cell;
    int   nx, ny, nz;
    int   cells_per_world;
    int   n_cells;           // cells_per_world * n_worlds
    int   n_scan;            // n_cells + 1: entries of the cell table that this step's scan has to cover
    float small_cut;         // largest bounding radius that goes into the grid
    int   overflow;
};

// counters that live in device memory so that no stage needs a host round trip
struct StepCounters {
    int n_sorted;            // bodies placed in the grid
    int n_cand_sb;           // sphere-box candidates
    int n_cand_bb;           // box-box candidates
    int n_contacts;          // contacts appended (all phases)
    int n_solver;            // contacts with >= 1 dynamic body
    int n_touched;           // dynamic bodies with >= 1 contact
    int n_levels;
    int overflow;            // bit0 pairs, bit1 contacts, bit2 grid, bit3 events
    int n_ss_tests;
    int n_frontier0;
    int n_events;
    int n_mslots;            // M slots handed out (level-major constraint order)
    int n_nodes;             // solver nodes (runs of contacts between the same two bodies)
    int n_hit_bb;            // box-box candidates that passed SAT with a FACE axis (stored from the front of hit_ids)
    int fq_tail;             // flow solver: entries pushed to the ready queue
    int fq_head;             // flow solver: chunk tickets handed out in iteration 0
    unsigned int bounds[6];  // ordered-int encoded min xyz / max xyz of small bodies
    unsigned int cut_dyn;    // float bits: largest bounding radius among non-static bodies
    unsigned int cut_all;    // float bits: largest bounding radius among all bodies
    int n_big;
    int fq_ticket;           // flow solver: chunk tickets handed out in iterations 1..I-1
    int n_hit_bb_edge;       // ... with an EDGE-EDGE axis (stored from the back of hit_ids)
    int replay_done;         // k_solve_levels ran iterations 1..I-1 itself (more levels than the staged replay's tables hold)
    int n_surv_bb;           // box-box candidates not separated by a face axis (compacted into cand_sb by k_bb_face)
    int n_deferred;          // level kernel: heavy terminal nodes held back for the last level (list in adj_cursor)
    int colour_overflow;     // coloured schedule: a node found all 64 colours taken on its bodies (the step replays in level order)
};

struct BbxDev {
    int device;
    cudaStream_t stream;
    int n_worlds, cap_s, cap_b;          // per-world capacities
    int NS, NB, N;                       // total slots: NS = n_worlds*cap_s, NB = n_worlds*cap_b
    int cap_contacts_world;
    int cap_contacts;                    // total contact capacity
    int cap_cand;                        // capacity of each candidate list
    int cap_cells;
    int cap_levels;
    int coop_blocks;                     // co-resident blocks for the solver
    int n_big_host;

    // raw mirrors of the host arrays (12-byte Vec3 stride), the unit of H2D/D2H
    float *r_s_pos, *r_s_radius, *r_s_vel, *r_s_force, *r_s_mass, *r_s_angvel, *r_s_torque, *r_s_inertia, *r_s_timer;
    unsigned char *r_s_static, *r_s_sleep, *r_s_hascb;
    float *r_b_pos, *r_b_size, *r_b_rot, *r_b_vel, *r_b_force, *r_b_mass, *r_b_angvel, *r_b_torque, *r_b_inertia, *r_b_timer;
    unsigned char *r_b_static, *r_b_sleep, *r_b_hascb;
    int *d_s_count, *d_b_count;          // [n_worlds]

    // packed working state, one float4 per body slot (spheres first, then boxes)
    float4 *pos;        // xyz, w = bounding radius (sphere: its radius)
    float4 *vel;        // 2 per body, one 32-byte sector: [2g] = (linear velocity, 1/mass or 0 when static),
                        //                                  [2g+1] = (angular velocity, spheres: 1/inertia or 0)
    float  *mass;       // N
    float  *timer;      // N
    unsigned char *flags;   // N
    float4 *boxq;       // 2 per box, one 32-byte sector: [2b] = rotation, [2b+1] = (1/Ix, 1/Iy, 1/Iz) or 0 when static
    float4 *half;       // NB half extents, w unused
    float4 *brec;       // 2 per body, one 32-byte sector, rebuilt before every presolve: (pos, 1/m) (1/I diag, flags bits): what
                        // k_presolve gathers per contact side (one sector instead of four)

    // grid
    GridParams *grid;
    int *cell_count, *cell_start, *cell_cursor;   // cap_cells (+1)
    int *body_cell;                               // N
    int *sorted_id;  float4 *sorted_pos;          // N
    int *big_list;   int *big_start;              // N, n_worlds+1
    int *scan_tmp;                                // block sums for scans

    // candidates and contacts (unordered append buffers)
    int2 *cand_sb, *cand_bb;
    int4 *hit_ids; float4 *hit_ax;   // box-box pairs that passed SAT: (A gid, B gid, axis type, -), (axis, |overlap|)
    int4 *c_ids;        // a gid, b gid (-2 = SDF), k, phase
    float4 *c_pt;       // contact point, w = penetration
    float4 *c_nrm;      // contact normal A->B

    // solver records in contact order (written by presolve)
    float4 *L;                   // 4 per contact, 64-byte records: (n, bias) (rA, mN) (rB, mT0) (jn, jt0, jt1, mT1)
    int4   *node;                // 2 per node head (32-byte sector): [2c] = (dynamic gid A or -1, dynamic gid B or -1,
                                 // run length, class), [2c+1] = (successor on A, successor on B, -, -); a successor is
                                 // head contact id | class << 28, or -1
    int *deg, *adj_start, *adj_cursor;           // N (+1)
    unsigned long long *adj;                     // 2*cap_contacts
    int *indeg;                                  // cap_contacts (node heads)
    unsigned int *bar;                           // grid barrier counter of the solve kernel
    // level-major node queues, one per size class so that the lanes of a warp sweep equally long runs
    int  *queue_k[BBX_NCLS];                     // node ids (first contact of the run)
    int4 *Nd_k[BBX_NCLS];                        // (first M slot, constraint count, dynamic gid A, dynamic gid B)
    int *slot_of;                                // cap_contacts: contact -> M slot
    int2 *link;                                  // cap_contacts: (successor on body A's chain, on body B's chain) of the node headed by contact c,
                                                 // each = node | class << 28, -1 = none (level path)
    // warm starting (extension): open-addressing table (body A, body B, k) -> final impulses of the previous step
    unsigned long long *warm_key; float4 *warm_val; unsigned int warm_mask;
    // coloured schedule (bbx_colour.cu, allocated at first use): the second set of level-major buffers the colour-major
    // order is written to (swapped with the first set after every coloured solve), and the colouring scratch
    float4 *M2; int *queue2_k[BBX_NCLS]; int4 *Nd2_k[BBX_NCLS]; int *level_count2;
    int *colq_k[BBX_NCLS];                       // colour of the node at the same position of queue_k
    unsigned long long *col_used;                // per body: colours of the nodes swept so far in iteration 0
    int *col_tab;                                // class totals, counts / fill / heads per (colour, class), record bases (COL_* offsets)
    int coop_blocks_col;                         // co-resident blocks of k_solve_levels<true>
    int colour_limit;                            // colours a step may use (64; BBX_COLOUR_LIMIT lowers it to exercise the fallback)
    int *warm_epoch;                             // per world: bumped by a world reset, stored with every entry (stale entries miss)
    int *level_count;                            // (cap_levels + 4) * BBX_NCLS
    // ---- barrier-free flow solver (single large world, bbx_solver.cu k_solve_flow) ----
    // one 64-byte record per body, every 8-byte word = (float data, int version):
    //   [0] (vx,ver)(vy,ver) [1] (vz,ver)(wx,ver) [2] (wy,ver)(wz,ver) [3] (1/m | 1/I << 32)(dependency depth)
    // stored whole by the thread that visited the body, polled by the thread of the body's next constraint
    ulonglong2 *flow;                            // 4 per body
    int  *fq;                                    // ready queue = visiting order of iteration 0 (cap_contacts, -1 = empty)
    int4 *fnd;                                   // 2 per queue slot: (gid A, gid B, pos A, len A) (pos B, len B, contact, -)
    ulonglong2 *fimp;                            // 2 per queue slot: (jn,tag)(jt0,tag) (jt1,tag)(mT1,tag), tag = iterations applied
    int last_per_contact;                        // the last solve used one node per constraint (layout of `node`)
    int last_flow;                               // the last solve ran on the flow solver (impulses live in fimp)
    int flow_blocks;                             // co-resident blocks of k_solve_flow
    int num_sms;                                 // cudaDevAttrMultiProcessorCount of `device`
    int replay_mode;                             // BBX_REPLAY: 1 (default) staged replay kernel (bbx_replay.cu) unless the schedule is thin, 2 always, 0 replay inside k_solve_levels
    float replay_ms_sum; int replay_steps;       // k_replay_staged alone, summed by the last bbx_dev_get_timing
    unsigned char* timing_has7;                  // host array [BBX_TIMING_STEPS]: step s recorded mark 7
    int slot_of_stale;                           // the last solve ran on the level kernels and slot_of has not been built for it yet
    int last_deferred;                           // n_deferred as of the last bbx_dev_stats
    int replay_blocks;                           // co-resident blocks of k_replay_staged
    int tune_solve_blocks;                       // BBX_SOLVE_BLOCKS: smaller grid for k_solve_levels
    int tune_flow_blocks, tune_flow_gather, trace_level;   // BBX_FLOW_BLOCKS, BBX_FLOW_GATHER, BBX_TRACE_LEVEL
    int grp_blocks;                              // co-resident blocks of k_solve_groups (sizes the groups of worlds)
    int grp_force_wpb;                           // BBX_GRP_WPB: at least this many worlds per group (tests of multi-world groups on small batches)
    int grp_pair_cap;                            // BBX_GRP_PAIR_CAP: lowers the constraints a group may keep in shared memory (tests of the hand-over)
    int grp_smem_wpb;                            // BBX_GRP_SMEM: groups of up to this many worlds keep their body state in shared memory during the solve
    int find_pairs_mode;                         // BBX_FIND_PAIRS: 0 lane-per-body walk, 1 flattened warp walk, unset (-1): by batch size
    int force_flow;                              // BBX_SOLVER_FLOW=1: the exact schedule runs on the flow solver (A/B measurements)
    unsigned long long *level_ns;                // optional trace: globaltimer at the end of each level of iteration 1 (first 4096)
    // level-major copies streamed by iterations >= 1
    float4 *M;                   // same four float4 fields, plane-major: field r of slot s is M[r*cap_contacts + s]; a node's
                                 // slots are contiguous, so a warp of equal-length nodes loads each plane coalesced

    // callback events (allocated when the first step with want_callbacks runs)
    int cap_events;
    unsigned long long *ev_key;  // emission-order key (same packing as the export key)
    int *ev_step;
    CollisionContact *ev_contact;
    int step_in_batch;           // step index since the last event fetch

    // snapshot for per-world resets (allocated by bbx_dev_snapshot)
    float4 *snap_pos, *snap_vel, *snap_boxq; float *snap_timer; unsigned char *snap_flags; int *reset_ids;
    int forces_on_device;        // accumulators were written through the device views
    int *wake_ids; int wake_cap; // scratch of bbx_dev_wake_bodies
    int imported_list;           // the device contact list was imported from the host in array order
    int wb_attr_set;             // dynamic shared memory attribute of the per-world solve kernel was set

    // baked SDF (host function sampled on a grid, bbx_dev_set_baked_sdf)
    float *sdf_grid; int sdf_ch; int sdf_nx, sdf_ny, sdf_nz; float sdf_ox, sdf_oy, sdf_oz, sdf_inv_cell;

    // joints and springs (extension, bbx_joints.cu).  Joints are stored in SCHEDULE order: the joints of one connected
    // component of the joint graph belong to one block, sorted by dependency level inside it
    int n_joints, n_jblocks, n_jlevels_max;
    int *j_blk_off;              // [n_jblocks + 1] range of this block's entries in j_lev_off
    int *j_lev_off;              // ranges of schedule positions, one per (block, level), + 1
    int4 *j_ids;                 // (gid A, gid B or -1, kind, -)
    float4 *j_local;             // 2 per joint: (anchor in A's frame, rest) (anchor in B's frame, -)
    float4 *j_rec;               // 5 per joint, written by k_joint_presolve: (rA, m0) (rB, m1) (bias xyz, m2) (imp xyz, -) (rod direction, -)
    int n_springs, n_spring_bodies;
    int *sp_body;                // [n_spring_bodies] body slots with at least one spring end
    int *sp_off;                 // [n_spring_bodies + 1] range in sp_list
    int *sp_list;                // spring index << 1 | end (0 = A, 1 = B), creation order per body
    int2 *sp_ids;                // (gid A, gid B or -1)
    float4 *sp_def;              // 3 per spring: (anchor A, rest) (anchor B, k) (c, -, -, -)

    StepCounters *ctr;           // device
    StepCounters *ctr_host;      // pinned host copy
    unsigned long long *sort_keys;   // export: 64-bit reference-order keys
    ContactConstraint *export_buf;   // export staging (device), contact order
    ContactConstraint *export_sorted;   // ... and in emission order
    unsigned long long *sort_keys_b; unsigned int *sort_vals_a, *sort_vals_b; int *sort_hist, *sort_hist_scan;

    // bookkeeping
    long long launches, steps, body_steps;
    int n_valid_spheres, n_valid_boxes;
    BbxStepParams last_params;
    int have_params;
    // optional per-stage timing (CUDA events on the step stream)
    int timing_on, timing_steps, timing_cap;
    cudaEvent_t *timing_ev;      // timing_cap * BBX_TIMING_MARKS events
    // drop-in (host array) path: device->host copies run on their own stream so that they overlap the export kernels
    cudaStream_t copy_stream;
    cudaEvent_t ev_upload, ev_chain;
    char *err;                   // 256-byte host buffer (kept out of the by-value kernel argument)
    int has_err;
};

// ---------------------------------------------------------------------------
// warm-start table: key = (gid A, gid B + 2, index k of the contact inside its pair's run)
// ---------------------------------------------------------------------------
#define BBX_WARM_EMPTY 0xffffffffffffffffull
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long warm_key_of(int a, int b, int k)
{
    return ((unsigned long long)(unsigned int)a << 36) | ((unsigned long long)(unsigned int)(b + 2) << 8) | (unsigned long long)(k & 255);
}
__device__ __forceinline__ unsigned int warm_hash(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (unsigned int)x;
}
#endif

// ---------------------------------------------------------------------------
// error handling / launch accounting
// ---------------------------------------------------------------------------
int  bbx_fail(BbxDev* d, int code, const char* what, const char* file, int line);
#define BBX_CUDA(dev, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
        return bbx_fail((dev), (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define BBX_LAUNCH(dev, kernel, grid, block, smem, ...) do { \
        kernel<<<(grid), (block), (smem), (dev)->stream>>>(__VA_ARGS__); (dev)->launches++; } while (0)

static inline int bbx_blocks(long long n, int block) { long long b = (n + block - 1) / block; return (int)(b < 1 ? 1 : b); }

void bbx_timing_mark(BbxDev* d, int mark);
int  bbx_check_overflow(BbxDev* d);            // ctr_host must be current
// stage entry points (one translation unit each)
int bbx_stage_pack(BbxDev* d);
int bbx_stage_unpack(BbxDev* d);
int bbx_stage_integrate_velocities(BbxDev* d, const BbxStepParams* p);
int bbx_stage_broadphase(BbxDev* d, const BbxStepParams* p);
int bbx_stage_narrowphase(BbxDev* d, const BbxStepParams* p);
int bbx_stage_solve(BbxDev* d, const BbxStepParams* p);
int bbx_stage_solve_ex(BbxDev* d, const BbxStepParams* p, int presolve_mode, int iterations, int imported);
int bbx_stage_apply_forces_only(BbxDev* d, const BbxStepParams* p);
int bbx_stage_integrate_velocities_only(BbxDev* d, const BbxStepParams* p);
int bbx_stage_positions_only(BbxDev* d, const BbxStepParams* p);
int bbx_stage_sleep_only(BbxDev* d, const BbxStepParams* p);
int bbx_stage_integrate_positions(BbxDev* d, const BbxStepParams* p);
int bbx_stage_collect_events(BbxDev* d, const BbxStepParams* p);
int bbx_warm_ensure(BbxDev* d);
int bbx_stage_warm_save(BbxDev* d, const BbxStepParams* p);
int bbx_springs_apply(BbxDev* d);                                       // bbx_joints.cu
int bbx_joints_presolve(BbxDev* d, const BbxStepParams* p, int wake);
int bbx_joints_sweep(BbxDev* d);
void bbx_joints_free(BbxDev* d);
int bbx_scan_exclusive(BbxDev* d, const int* in, int* out, int* out2, int n_max, const int* n_dev, int* total_out);

#ifdef __CUDACC__
#define BBX_HDI __host__ __device__ __forceinline__     // geometry shared by the kernels and the host queries
// contact-run encoding in c_ids.z: low 8 bits = index k inside the run, bits 8.. = run length
#define RUN_K(z)   ((z) & 0xff)
#define RUN_LEN(z) ((z) >> 8)
#define RUN_ENC(k, len) ((k) | ((len) << 8))
__host__ __device__ __forceinline__ int bbx_node_class(int m, bool any_box)
{
    if (m <= 1) return any_box ? 1 : 0;
    if (m <= 4) return m;
    if (m <= 6) return 5;
    if (m <= 8) return 6;
    return 7;
}
// ---------------------------------------------------------------------------
// exact vector math (reference ballbox.h:501-535)
// ---------------------------------------------------------------------------
BBX_HDI float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
BBX_HDI float3 xyz(float4 v) { return make_float3(v.x, v.y, v.z); }
BBX_HDI float3 v3add(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
BBX_HDI float3 v3sub(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
BBX_HDI float3 v3scale(float3 v, float s) { return f3(v.x * s, v.y * s, v.z * s); }
BBX_HDI float  v3dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
BBX_HDI float3 v3cross(float3 a, float3 b)
{
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
BBX_HDI float  v3len(float3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
BBX_HDI float3 v3norm(float3 v)
{
    float l = v3len(v);
    if (l < 1e-6f) return f3(0.0f, 0.0f, 0.0f);
    return v3scale(v, 1.0f / l);
}
BBX_HDI float3 v3mul(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }

// glibc 2.39 fminf/fmaxf return the FIRST operand on ties (measured in this
// image: fminf(+0,-0) = +0, fminf(-0,+0) = -0) and the non-NaN operand otherwise.
BBX_HDI float bb_fminf(float x, float y) { return (x <= y || y != y) ? x : y; }
BBX_HDI float bb_fmaxf(float x, float y) { return (x >= y || y != y) ? x : y; }

// quaternion helpers (reference ballbox.h:543-562, :625-639)
BBX_HDI float4 qconj(float4 q) { return make_float4(-q.x, -q.y, -q.z, q.w); }
BBX_HDI float3 qrot(float4 q, float3 v)
{
    float3 u = f3(q.x, q.y, q.z);
    float  s = q.w;
    float3 a = v3scale(v, s * s - v3dot(u, u));
    float3 b = v3scale(u, 2.0f * v3dot(u, v));
    float3 c = v3scale(v3cross(u, v), 2.0f * s);
    return v3add(v3add(a, b), c);
}

// 256-bit global accesses (sm_100a LDG/STG.E.ENL2.256): one instruction and one 32-byte sector
// per lane for the packed body state, half the L1TEX wavefronts of two float4 accesses.
// .cg = cache in L2 only: data another block wrote before a grid barrier is never served stale.
struct __align__(32) F8 { float4 a, b; };
__device__ __forceinline__ F8 ld256_cg(const float4* p)
{
    F8 r;
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
    return r;
}
__device__ __forceinline__ F8 ld256_nc(const float4* p)          // read-only data
{
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256_cg(float4* p, float4 a, float4 b)
{
    asm volatile("st.global.cg.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// The same accesses with an L2 eviction policy (createpolicy): the body state is re-read at every level and should stay
// in L2 (evict_last), the level-major records stream through once per sweep (evict_first).
__device__ __forceinline__ unsigned long long l2_policy_keep() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ unsigned long long l2_policy_stream() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ unsigned long long l2_policy_normal() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p; }
// records stream only when they cannot stay in L2 anyway (64 B per contact against the 126 MB L2); in a small scene
// evict_first would throw away records that the next sweep finds in L2 today (measured: +3 % on 100k-200k bodies)
#define BBX_L2_STREAM_CONTACTS 1000000
__device__ __forceinline__ F8 ld256_cg_pol(const float4* p, unsigned long long pol)
{
    F8 r;
    asm volatile("ld.global.cg.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ F8 ld256_nc_pol(const float4* p, unsigned long long pol)
{
    F8 r;
    asm volatile("ld.global.nc.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st256_cg_pol(float4* p, float4 a, float4 b, unsigned long long pol)
{
    asm volatile("st.global.cg.L2::cache_hint.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %9;"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st32_pol(int* p, int v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st64_pol(int2* p, int2 v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1,%2}, %3;" :: "l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ int2 ld64_pol(const int2* p, unsigned long long pol)
{
    int2 v;
    asm volatile("ld.global.L2::cache_hint.v2.b32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld128_cg_pol(const float4* p, unsigned long long pol)
{
    float4 r;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ int4 ld128i_cg_pol(const int4* p, unsigned long long pol)
{
    int4 r;
    asm volatile("ld.global.cg.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st128_cg_pol(float4* p, float4 v, unsigned long long pol)
{
    asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// order-preserving float <-> uint encoding for atomicMin/atomicMax
__device__ __forceinline__ unsigned int f2ord(float f)
{
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bounding radius of a box whose rotation may be non-unit: QuatRotateVec3 scales
// by |q|^2 (ballbox.h:630-639) and AddBox never normalises (ballbox.h:716).
BBX_HDI float box_bound_radius(float3 half, float4 q)
{
    float q2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    if (!(q2 > 1.0f)) q2 = 1.0f;
    return v3len(half) * q2 * 1.0005f + 1e-5f;
}

// warp-aggregated append: every calling lane with pred reserves one slot
__device__ __forceinline__ int warp_append(int* counter, bool pred)
{
    unsigned int active = __activemask();
    unsigned int mask = __ballot_sync(active, pred);
    if (!pred) return -1;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}
// full-warp reservation of n slots per lane (n may be 0); all 32 lanes must call
__device__ __forceinline__ int warp_reserve(int* counter, int n)
{
    int lane = threadIdx.x & 31;
    int incl = n;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - n;
}

// block-wide reservation of n slots per thread (n may be 0) with ONE global atomic per block: every thread of the
// block must call (uniform control flow); `s_warp` is a shared int[32].  The global append counters are the
// busiest addresses of the detection kernels: a same-address atomic with a returned value costs 0.5-0.7 ns
// chip-wide and ~1.7 us per dependent use under contention (tools/micro/atomic_bench.cu).
__device__ __forceinline__ int block_reserve(int* counter, int n, int* s_warp)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int incl = n;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = lane < nw ? s_warp[lane] : 0, inc2 = v;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += y; }
        int tot = __shfl_sync(0xffffffffu, inc2, 31), base = 0;
        if (lane == 0 && tot > 0) base = atomicAdd(counter, tot);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < nw) s_warp[lane] = base + inc2 - v;
    }
    __syncthreads();
    return s_warp[wid] + incl - n;
}

#endif // __CUDACC__
