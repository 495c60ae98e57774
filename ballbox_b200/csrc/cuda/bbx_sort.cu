// bbx_sort.cu -- stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits
// per pass, used to put the exported constraint list into the reference's emission
// order (world, phase, i, j, k) -- SURVEY.md section 8a row 3.
//
// One pass = tile histogram -> exclusive scan over (digit, tile) -> stable scatter.
// A tile is 2048 keys handled by 8 warps; each warp owns a contiguous chunk of 256
// keys and ranks them with __match_any_sync, so the order inside a tile (warp-major,
// row-major, lane-minor) equals the input order and the sort is stable.
#include "bbx_dev.cuh"

#define RS_TILE   2048
#define RS_WARPS  8
#define RS_ROWS   8          // rows of 32 keys per warp

__device__ __forceinline__ unsigned int rs_digit(unsigned long long key, int shift) { return (unsigned int)(key >> shift) & 255u; }

// per-warp digit histogram of the warp's chunk, in shared memory wh[warp][256]
__device__ __forceinline__ void rs_warp_hist(const unsigned long long* keys, int n, int tile, int shift, int wh[RS_WARPS][256])
{
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int d = lane; d < 256; d += 32) wh[warp][d] = 0;
    __syncwarp();
    int base = tile * RS_TILE + warp * (RS_ROWS * 32);
    for (int r = 0; r < RS_ROWS; r++) {
        int i = base + r * 32 + lane;
        bool live = i < n;
        unsigned int dg = live ? rs_digit(keys[i], shift) : 256u + lane;      // dead lanes match nobody
        unsigned int peers = __match_any_sync(0xffffffffu, dg);
        if (live && (peers & ((1u << lane) - 1u)) == 0) wh[warp][dg] += __popc(peers);   // lowest lane of the group
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_rs_hist(const unsigned long long* keys, int n, int shift, int n_tiles, int* hist)
{
    __shared__ int wh[RS_WARPS][256];
    int tile = blockIdx.x;
    rs_warp_hist(keys, n, tile, shift, wh);
    __syncthreads();
    int d = threadIdx.x;                       // 256 threads, one digit each
    int s = 0;
    for (int w = 0; w < RS_WARPS; w++) s += wh[w][d];
    hist[d * n_tiles + tile] = s;              // digit-major so that one scan gives global offsets
}

__global__ void __launch_bounds__(256) k_rs_scatter(const unsigned long long* keys, const unsigned int* vals, int n, int shift,
                                                    int n_tiles, const int* offs, unsigned long long* keys_out, unsigned int* vals_out)
{
    __shared__ int wh[RS_WARPS][256];          // becomes the running offset of (warp, digit)
    int tile = blockIdx.x;
    rs_warp_hist(keys, n, tile, shift, wh);
    __syncthreads();
    {
        int d = threadIdx.x;
        int run = offs[d * n_tiles + tile];
        for (int w = 0; w < RS_WARPS; w++) { int c = wh[w][d]; wh[w][d] = run; run += c; }
    }
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int base = tile * RS_TILE + warp * (RS_ROWS * 32);
    for (int r = 0; r < RS_ROWS; r++) {
        int i = base + r * 32 + lane;
        bool live = i < n;
        unsigned long long k = live ? keys[i] : 0ull;
        unsigned int dg = live ? rs_digit(k, shift) : 256u + lane;
        unsigned int peers = __match_any_sync(0xffffffffu, dg);
        int rank = __popc(peers & ((1u << lane) - 1u));
        int dst = 0;
        if (live) dst = wh[warp][dg] + rank;
        __syncwarp();
        if (live && rank == 0) wh[warp][dg] += __popc(peers);
        __syncwarp();
        if (live) { keys_out[dst] = k; vals_out[dst] = vals[i]; }
    }
}

__global__ void k_rs_iota(unsigned int* vals, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vals[i] = (unsigned int)i;
}

// Sorts n pairs by the low `bits` bits of the key.  On return *keys_sorted / *vals_sorted
// point at whichever of the two buffers holds the result.  `hist` needs 256 * tiles ints.
int bbx_radix_sort(BbxDev* d, unsigned long long* keys_a, unsigned int* vals_a, unsigned long long* keys_b,
                   unsigned int* vals_b, int n, int bits, int* hist, int* hist_scan,
                   unsigned long long** keys_sorted, unsigned int** vals_sorted)
{
    int n_tiles = (n + RS_TILE - 1) / RS_TILE;
    if (n_tiles < 1) n_tiles = 1;
    BBX_LAUNCH(d, k_rs_iota, bbx_blocks(n, 256), 256, 0, vals_a, n);
    unsigned long long *kin = keys_a, *kout = keys_b;
    unsigned int *vin = vals_a, *vout = vals_b;
    for (int shift = 0; shift < bits; shift += 8) {
        BBX_LAUNCH(d, k_rs_hist, n_tiles, 256, 0, kin, n, shift, n_tiles, hist);
        int rc = bbx_scan_exclusive(d, hist, hist_scan, NULL, 256 * n_tiles, NULL, NULL);
        if (rc) return rc;
        BBX_LAUNCH(d, k_rs_scatter, n_tiles, 256, 0, kin, vin, n, shift, n_tiles, hist_scan, kout, vout);
        unsigned long long* tk = kin; kin = kout; kout = tk;
        unsigned int* tv = vin; vin = vout; vout = tv;
    }
    BBX_CUDA(d, cudaGetLastError());
    *keys_sorted = kin; *vals_sorted = vin;
    return 0;
}
